"""Stock PyTorch-CUDA timing of the reference math (the north_star's ">= 6x" denominator): the CPU oracle's fp32 torch
restatement run eagerly on the GPU (TF32 matmul off = torch default), and the same in bf16 for context.
Informational only — not part of bench.py.   python tools/bench_stock_gpu.py [frames]"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from hirest_b200 import synthetic
from oracle import eva_oracle

B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
dev = torch.device("cuda:0")
cfg = synthetic.EVA_G14
sd = {k: v for k, v in synthetic.make_visual_state_dict(cfg, 0, device=dev).items()}
sd = {"visual." + k: v for k, v in sd.items()}
frames = synthetic.make_frames(B, 224, seed=1, device=dev)
res = {"frames": B, "allow_tf32": torch.backends.cuda.matmul.allow_tf32}
for name, cast in (("fp32", torch.float32), ("bf16", torch.bfloat16)):
    w = {k: v.to(cast) for k, v in sd.items()}
    x = frames.to(cast)
    with torch.no_grad():
        for _ in range(2):
            eva_oracle.encode_image(w, x, cfg)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        n = 3
        for _ in range(n):
            eva_oracle.encode_image(w, x, cfg)
        e1.record()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    res[name] = {"ms_per_batch": ms, "frames_per_s": B / (ms * 1e-3)}
    del w
print(json.dumps(res))
