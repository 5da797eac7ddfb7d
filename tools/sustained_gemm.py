"""Power-capped (sustained) throughput of the ViT GEMM shapes: this library's tcgen05 kernels vs cuBLAS (torch.matmul) on the
same shapes, each looped for --seconds so the board sits at its power cap like it does inside a bench.py step.
Tells whether a gap to MEASURED_PEAKS' 8192^3 cuBLAS figure is the kernel or the shape (short K => more epilogue bytes/FLOP)."""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hirest_b200 import _lib  # noqa: E402


class Sampler(threading.Thread):
    def __init__(self):
        super().__init__(daemon=True)
        self.stop = False
        self.rows = []

    def run(self):
        while not self.stop:
            try:
                out = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader,nounits", "-i", "0"],
                                     capture_output=True, text=True, timeout=5).stdout.strip().split(",")
                self.rows.append((float(out[0]), float(out[1])))
            except Exception:
                pass
            time.sleep(0.2)


def sustained(fn, seconds, flops):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    smp = Sampler()
    smp.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 0
    t0 = time.time()
    e0.record()
    while time.time() - t0 < seconds:
        for _ in range(20):
            fn()
        n += 20
        torch.cuda.synchronize()
    e1.record()
    torch.cuda.synchronize()
    smp.stop = True
    smp.join()
    ms = e0.elapsed_time(e1) / n
    rows = smp.rows[len(smp.rows) // 3:] or [(0.0, 0.0)]
    clk = sorted(r[0] for r in rows)[len(rows) // 2]
    pw = sorted(r[1] for r in rows)[len(rows) // 2]
    return {"ms": ms, "tflops": flops / ms / 1e9, "sm_mhz": clk, "power_w": pw}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=float, default=3.0)
    ap.add_argument("--frames", type=int, default=1024)
    ap.add_argument("--cg", type=int, default=2)
    ap.add_argument("--only", default="")
    ap.add_argument("--no-cublas", action="store_true")
    ap.add_argument("--no-resid", action="store_true", help="fp32-output shapes without the residual read (isolates the epilogue's load side)")
    a = ap.parse_args()
    dev = "cuda:0"
    hb = _lib.init(0)
    _lib.check(hb.hb_debug_set(b"gemm_cta_group", a.cg))
    M = a.frames * 257
    shapes = [("qkv", 4224, 1408, 0), ("fc1", 6144, 1408, 1), ("proj", 1408, 1408, 2), ("fc2", 1408, 6144, 2), ("square8192", 8192, 8192, 0)]
    res = {}
    for name, N, K, epi in shapes:
        if a.only and name not in a.only.split(","):
            continue
        Mx = 8192 if name == "square8192" else M
        x = torch.randn(Mx, K, device=dev).bfloat16()
        w = (torch.randn(N, K, device=dev) * K ** -0.5).bfloat16()
        b = torch.randn(N, device=dev)
        flops = 2.0 * Mx * N * K
        out_t = torch.empty(Mx, N, device=dev, dtype=torch.float32 if epi == 2 else torch.bfloat16)
        resid = out_t if (epi == 2 and not a.no_resid) else None
        wt = w.t()
        out_c = torch.empty(Mx, N, device=dev, dtype=torch.bfloat16)
        s = _lib.stream_ptr()

        def ours():
            _lib.check(hb.hb_linear(x.data_ptr(), K, w.data_ptr(), K, b.data_ptr(), resid.data_ptr() if resid is not None else None,
                                    out_t.data_ptr(), N, Mx, N, K, epi, s), "hb_linear")

        def cublas():
            torch.matmul(x, wt, out=out_c)

        res[name] = {"M": Mx, "N": N, "K": K, "epilogue": ["bias->bf16", "bias+gelu->bf16", "bias+fp32 residual->fp32"][epi],
                     "cg": a.cg, "hirest_b200": sustained(ours, a.seconds, flops),
                     "cublas_plain_bf16": None if a.no_cublas else sustained(cublas, a.seconds, flops)}
        print(name, json.dumps(res[name]), flush=True)
        del x, w, out_t, out_c
        torch.cuda.empty_cache()
    print(json.dumps(res))


if __name__ == "__main__":
    main()
