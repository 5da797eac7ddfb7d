"""Run the ViT attention kernel a few times (for ncu)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from hirest_b200 import _lib
lib = _lib.init(0)
B, H = int(sys.argv[1]) if len(sys.argv) > 1 else 64, 16
if len(sys.argv) > 2:
    _lib.check(lib.hb_debug_set(b"attention_version", int(sys.argv[2])))
D = H * 88
qkv = (torch.randn(B * 257, 3 * D, device="cuda") * 0.7).bfloat16()
out = torch.empty(B * 257, D, device="cuda", dtype=torch.bfloat16)
for _ in range(3):
    _lib.check(lib.hb_vit_attention(qkv.data_ptr(), out.data_ptr(), B, H, _lib.stream_ptr()))
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    _lib.check(lib.hb_vit_attention(qkv.data_ptr(), out.data_ptr(), B, H, _lib.stream_ptr()))
e1.record(); torch.cuda.synchronize()
print(f"vit_attention B={B} H={H}: {e0.elapsed_time(e1)/5:.3f} ms per launch")
