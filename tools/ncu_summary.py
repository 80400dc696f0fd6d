"""Key numbers of every kernel in an .ncu-rep (raw page): duration, DRAM bytes / throughput, tensor-pipe and issue utilisation, registers,
occupancy — plus, with --top N, the stall reasons and the most-stalled SASS lines of the first kernel (source page).
    python tools/ncu_summary.py report.ncu-rep [--top 12]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 0
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
want = [("gpu__time_duration.sum", "duration"), ("dram__bytes_read.sum", "dram read"), ("dram__bytes_write.sum", "dram write"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram % of peak"), ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm % of peak"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe % (active SMs' cycles)"),
        ("sm__inst_executed_pipe_tensor.avg.pct_of_peak_sustained_active", "tensor inst %"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"), ("smsp__inst_executed.sum", "warp instructions"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"), ("launch__registers_per_thread", "registers / thread"),
        ("launch__grid_size", "grid"), ("launch__block_size", "block"), ("launch__shared_mem_per_block_dynamic", "dynamic smem / block"),
        ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem wavefronts"), ("lts__t_bytes.sum", "L2 bytes")]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print("==", d.get("Kernel Name", "?")[:150])
    for k, label in want:
        if k in d and d[k] != "":
            print(f"   {label:38s} {d[k]} {units[hdr.index(k)]}")
if top:
    subprocess.run([sys.executable, __file__.replace("ncu_summary.py", "ncu_top.py"), rep, str(top)])
