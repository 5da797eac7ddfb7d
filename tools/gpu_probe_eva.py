"""Exploratory GPU probe: per-kernel and end-to-end errors of the CUDA path vs the CPU oracle (prints numbers)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import ctypes as C
from hirest_b200 import _lib, eva_clip
from oracle import weights, eva_oracle

dev = torch.device("cuda:0")
lib = _lib.init(0)

def rel(a, b):
    a = a.float().cpu(); b = b.float().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-12)).item(), (a - b).abs().max().item()

# ---- hb_linear
torch.manual_seed(0)
M, N, K = 300, 352, 352
x = torch.randn(M, K, device=dev).bfloat16(); w = (torch.randn(N, K, device=dev) * 0.05).bfloat16()
b = torch.randn(N, device=dev)
out = torch.empty(M, N, device=dev)
_lib.check(lib.hb_linear(x.data_ptr(), K, w.data_ptr(), K, b.data_ptr(), None, out.data_ptr(), N, M, N, K, 2, _lib.stream_ptr()))
torch.cuda.synchronize()
print("linear f32", rel(out, x.float() @ w.float().T + b))

# ---- vit attention
B, H = 3, 4
D = H * 88
qkv = (torch.randn(B * 257, 3 * D, device=dev) * 0.5).bfloat16()
o = torch.empty(B * 257, D, device=dev, dtype=torch.bfloat16)
_lib.check(lib.hb_vit_attention(qkv.data_ptr(), o.data_ptr(), B, H, _lib.stream_ptr()))
torch.cuda.synchronize()
q5 = qkv.float().reshape(B, 257, 3, H, 88).permute(2, 0, 3, 1, 4)
ref = ((q5[0] @ q5[1].transpose(-2, -1)).softmax(-1) @ q5[2]).transpose(1, 2).reshape(B * 257, D)
print("vit attn", rel(o, ref))
e = (o.float() - ref).abs().reshape(B, 257, H, 88)
print("  per-token-group max err: rows0-127 %.4g rows128-255 %.4g row256 %.4g" % (e[:, :128].max(), e[:, 128:256].max(), e[:, 256].max()))

# ---- small attention (causal)
Bq, Hs, T = 2, 2, 77
W = Hs * 64
qkv = (torch.randn(Bq * T, 3 * W, device=dev)).bfloat16()
o = torch.empty(Bq * T, W, device=dev, dtype=torch.bfloat16)
_lib.check(lib.hb_small_attention(qkv.data_ptr(), qkv.data_ptr() + 2 * W, qkv.data_ptr() + 4 * W, o.data_ptr(), Bq, Hs, T, T,
                                  3 * W, 3 * W, 3 * W, W, T * 3 * W, T * 3 * W, T * 3 * W, T * W, 0.125, 1, 0.0, 0, _lib.stream_ptr()))
torch.cuda.synchronize()
q4 = qkv.float().reshape(Bq, T, 3, Hs, 64).permute(2, 0, 3, 1, 4)
mask = torch.full((T, T), float("-inf"), device=dev).triu_(1)
ref = ((q4[0] @ q4[1].transpose(-2, -1) * 0.125 + mask).softmax(-1) @ q4[2]).transpose(1, 2).reshape(Bq * T, W)
print("small attn causal", rel(o, ref))

# ---- tiny EVA end to end with taps
for name, cfg, nf, nt in (("tiny", weights.EVA_TINY, 4, 6),):
    sd = weights.make_eva_state_dict(cfg, 0)
    model = eva_clip.EVA_CLIP(**cfg)
    print(model.load_state_dict(sd, strict=True))
    model = model.to(dev).eval()
    frames = weights.make_frames(nf, 224, 1); tokens = weights.make_tokens(nt, cfg, 2)
    taps = {0: None, 1: None, 3: None}
    with torch.no_grad():
        ref_im = eva_oracle.encode_image(sd, frames, cfg, taps=taps)
        ref_tx = eva_oracle.encode_text(sd, tokens, cfg)
    Dv = cfg["vision_cfg"]["width"]
    for layer in (0, 1, 3):
        tap = torch.empty(nf * 257, Dv, device=dev)
        got = model.visual(frames.to(dev), tap=(layer, tap))
        torch.cuda.synchronize()
        print(name, "tap", layer, rel(tap.reshape(nf, 257, Dv), taps[layer]))
    print(name, "encode_image", rel(got, ref_im))
    tx = model.encode_text(tokens.to(dev)); torch.cuda.synchronize()
    print(name, "encode_text", rel(tx, ref_tx))
    # stock torch bf16 drift of the same math, for the tolerance gate
    sdb = {k: v.to(dev).bfloat16() for k, v in sd.items()}
    with torch.no_grad():
        im_b = eva_oracle.encode_image(sdb, frames.to(dev).bfloat16(), cfg)
    print(name, "stock bf16 torch drift", rel(im_b, ref_im))
print("launches", lib.hb_launch_count())
