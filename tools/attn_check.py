"""ViT attention kernel: numerics of every version against fp32 torch + time per launch (CUDA events).
    python tools/attn_check.py [B] [versions...]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from hirest_b200 import _lib
lib = _lib.init(0)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
versions = [int(v) for v in sys.argv[2:]] or [2, 3]
H, D = 16, 16 * 88
torch.manual_seed(0)
qkv = (torch.randn(B * 257, 3 * D, device="cuda") * 0.7).bfloat16()
out = torch.empty(B * 257, D, device="cuda", dtype=torch.bfloat16)
nb = min(B, 4)
q = qkv[:nb * 257].float().reshape(nb, 257, 3, H, 88).permute(2, 0, 3, 1, 4)
ref = ((q[0] @ q[1].transpose(-2, -1)).softmax(-1) @ q[2]).transpose(1, 2).reshape(nb * 257, D)
qe = qkv[(B - 1) * 257:].float().reshape(1, 257, 3, H, 88).permute(2, 0, 3, 1, 4)
ref_last = ((qe[0] @ qe[1].transpose(-2, -1)).softmax(-1) @ qe[2]).transpose(1, 2).reshape(257, D)
for v in versions:
    _lib.check(lib.hb_debug_set(b"attention_version", v))
    out.zero_()
    try:
        for _ in range(2):
            _lib.check(lib.hb_vit_attention(qkv.data_ptr(), out.data_ptr(), B, H, _lib.stream_ptr()))
        torch.cuda.synchronize()
    except Exception as ex:  # noqa: BLE001
        print(f"v{v}: FAILED {ex}")
        continue
    got = out[:nb * 257].float()
    rel = float((got - ref).norm() / ref.norm())
    err = (got - ref).abs().reshape(nb, 257, D)
    rel_last = float((out[(B - 1) * 257:].float() - ref_last).norm() / ref_last.norm())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 10
    e0.record()
    for _ in range(n):
        _lib.check(lib.hb_vit_attention(qkv.data_ptr(), out.data_ptr(), B, H, _lib.stream_ptr()))
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    gb = B * 257 * 4 * D * 2 / 1e9
    print(f"v{v}: B={B} rel {rel:.3e} (last frame {rel_last:.3e}) max|err| rows<256 {float(err[:, :256].max()):.4f} row256 {float(err[:, 256].max()):.4f}"
          f" | {ms:.3f} ms/launch = {gb / ms:.2f} TB/s algorithmic ({gb:.2f} GB)")
