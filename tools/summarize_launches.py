"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name (share of the step)."""
import csv
import sys
from collections import defaultdict

rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if not l.startswith("==")]
rd = csv.DictReader(lines)
agg = defaultdict(lambda: [0, 0.0])
total = 0.0
for r in rd:
    try:
        v = float(r["Metric Value"].replace(",", ""))
    except Exception:
        continue
    unit = r.get("Metric Unit", "ns")
    us = v / 1e3 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1e3)
    name = r["Kernel Name"]
    name = name[:110]
    agg[name][0] += 1
    agg[name][1] += us
    total += us
own = sum(v[1] for k, v in agg.items() if "tfasr" in k or any(s in k for s in ("gemm_bf16_kernel", "rnnt_", "logmel", "layernorm_", "bn_", "dwconv", "joint_", "conv1_", "im2col", "col2im", "scale_mask", "glu_bwd", "ctc_", "rel_attention")))
print(f"total kernel time (serialised, cold cache) {total/1e3:.3f} ms over {sum(v[0] for v in agg.values())} launches; hand-written kernels {own/1e3:.3f} ms = {100*own/total:.1f} %")
print(f"{'share':>7} {'ms':>9} {'launches':>8}  kernel")
for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:45]:
    print(f"{100*us/total:6.2f}% {us/1e3:9.3f} {n:8d}  {k}")
