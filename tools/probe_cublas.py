"""One cuBLAS bf16 GEMM and one hb_linear of the same shape, for an ncu capture (kernel names, L2->SM bytes, DRAM bytes)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hirest_b200 import _lib
hb = _lib.init(0)
for (M, N, K) in [(8192, 8192, 8192), (263168, 4224, 1408)]:
    x = torch.randn(M, K, device="cuda").bfloat16()
    w = (torch.randn(N, K, device="cuda") * K ** -0.5).bfloat16()
    b = torch.zeros(N, device="cuda")
    out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    for _ in range(2):
        torch.matmul(x, w.t(), out=out)
        _lib.check(hb.hb_linear(x.data_ptr(), K, w.data_ptr(), K, b.data_ptr(), None, out.data_ptr(), N, M, N, K, 0, _lib.stream_ptr()))
    torch.cuda.synchronize()
