"""Where does a label step of the cluster LSTM go?  Builds probe variants of csrc/lstm.cu (here: `python tools/lstm_probe.py build`) and times the
forward recurrent kernel of each on the GPU box (results of the probe variants are wrong by construction; only their time matters)."""
import ctypes
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
VARIANTS = {"full": [], "no_mma": ["-DLSTM_PROBE_NO_MMA"], "no_sync": ["-DLSTM_PROBE_NO_SYNC"], "no_push": ["-DLSTM_PROBE_NO_PUSH"],
            "no_mma_no_sync": ["-DLSTM_PROBE_NO_MMA", "-DLSTM_PROBE_NO_SYNC"]}


def build():
    from tensorflowasr_b200 import _build

    for name, flags in VARIANTS.items():
        out = os.path.join(ROOT, "tools", f"liblstm_{name}.so")
        subprocess.run([_build._nvcc(), *_build.NVCC_FLAGS, *flags, "-shared", "-o", out, os.path.join(_build.CSRC, "lstm.cu"), os.path.join(_build.CSRC, "api.cu"), "-lcudart"], check=True)
        print(out)


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "build":
        return build()
    import torch

    B, U, R = 32, 65, 320
    gx = torch.randn(B, U, 4 * R, device="cuda")
    whh = torch.randn(R, 4 * R, device="cuda") * 0.05
    y = torch.empty(B, U, R, dtype=torch.bfloat16, device="cuda")
    saved = torch.empty(B, U, 5, R, device="cuda")
    for name in VARIANTS:
        lib = ctypes.CDLL(os.path.join(ROOT, "tools", f"liblstm_{name}.so"))
        fn = lib.tfasr_lstm_rec_fwd
        fn.argtypes = [ctypes.c_void_p] * 9 + [ctypes.c_int32] * 3 + [ctypes.c_void_p]
        st = torch.cuda.current_stream().cuda_stream
        call = lambda: fn(gx.data_ptr(), whh.data_ptr(), None, None, None, y.data_ptr(), saved.data_ptr(), None, None, B, U, R, st)
        for _ in range(3):
            assert call() == 0
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            call()
        e1.record()
        torch.cuda.synchronize()
        print(f"{name:16s} {e0.elapsed_time(e1) / 10 * 1e3:8.1f} us per launch  ({e0.elapsed_time(e1) / 10 / U * 1e3:6.2f} us per label step)", flush=True)


if __name__ == "__main__":
    main()
