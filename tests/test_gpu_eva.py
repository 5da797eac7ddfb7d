"""End-to-end parity of encode_image / encode_text on the GPU against the CPU oracle and the reference-generated goldens.

Tolerances: GEMM operands are bf16 (north_star), the residual stream / LN / softmax are fp32.  SURVEY.md §7 measured that
stock PyTorch in bf16 drifts 0.8-1.8e-2 (relative L2) from the fp32 reference end to end, so the gate is
(ii) "our error vs the fp32 oracle <= stock torch-bf16's error on the same inputs", plus fixed caps written below."""
import os

import pytest
import torch

from hirest_b200 import _lib, eva_clip, retrieval, synthetic
from oracle import eva_oracle

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda:0")
# Caps = ~1.3x the measured values (fixed 2e-2 caps would have let a 2x regression through).
TOL_TINY_IMAGE = 5.7e-3  # relative L2, 3 layers; measured 4.3e-3 (bf16 GEMM operands, fp32 residual stream)
TOL_TINY_TEXT = 1e-4     # precise text tower (split-bf16 GEMMs, fp32 attention): measured ~1e-5; the bf16 tower measured 8.1e-3
TOL_G14_IMAGE = 9e-3     # relative L2, 40 layers; measured 6.6e-3 (stock bf16 torch on the same box: ~1.8e-2)
TOL_G14_TEXT = 1e-4      # measured ~1e-5


def rel(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return float((a - b).norm() / b.norm())


@pytest.fixture(scope="module")
def tiny(hb):
    cfg = synthetic.EVA_TINY
    sd = synthetic.make_eva_state_dict(cfg, seed=0)
    model = eva_clip.EVA_CLIP(**cfg)
    model.load_state_dict(sd, strict=True)
    return cfg, sd, model.to(DEV).eval()


def test_tiny_encode_image_vs_oracle_and_golden(tiny, golden_dir):
    cfg, sd, model = tiny
    g = torch.load(os.path.join(golden_dir, "eva_tiny.pt"))
    frames = synthetic.make_frames(4, 224, seed=1)
    with torch.no_grad():
        ref = eva_oracle.encode_image(sd, frames, cfg)
    got = model.encode_image(frames.to(DEV))
    assert got.shape == (4, cfg["embed_dim"]) and got.dtype == torch.float32 and got.device.type == "cuda"
    assert rel(got, ref) < TOL_TINY_IMAGE
    assert rel(got, g["image"]) < TOL_TINY_IMAGE
    # gate (ii): not worse than stock torch bf16 on the same math
    sdb = {k: v.to(DEV).bfloat16() for k, v in sd.items()}
    with torch.no_grad():
        stock = eva_oracle.encode_image(sdb, frames.to(DEV).bfloat16(), cfg)
    assert rel(got, ref) <= rel(stock, ref) * 1.05


def test_tiny_residual_stream_taps(tiny, golden_dir):
    cfg, sd, model = tiny
    g = torch.load(os.path.join(golden_dir, "eva_tiny.pt"))
    frames = synthetic.make_frames(4, 224, seed=1).to(DEV)
    D = cfg["vision_cfg"]["width"]
    for layer in (0, 1, 3):
        tap = torch.empty(4 * 257, D, device=DEV)
        model.visual(frames, tap=(layer, tap))
        torch.cuda.synchronize()
        assert rel(tap.reshape(4, 257, D), g[f"tap{layer}"]) < 5e-3, f"residual stream after {layer} blocks"


def test_tiny_encode_text(tiny, golden_dir):
    cfg, sd, model = tiny
    g = torch.load(os.path.join(golden_dir, "eva_tiny.pt"))
    tokens = synthetic.make_tokens(6, cfg, seed=2)
    got = model.encode_text(tokens.to(DEV))
    assert rel(got, g["text"]) < TOL_TINY_TEXT


def test_batch_independence_and_chunking(tiny):
    """Each frame's embedding is independent of what else is in the batch, and of how the batch is chunked."""
    cfg, sd, model = tiny
    frames = synthetic.make_frames(5, 224, seed=9).to(DEV)
    full = model.encode_image(frames)
    singles = torch.cat([model.encode_image(frames[i:i + 1]) for i in range(5)])
    assert torch.equal(full, singles)
    small = eva_clip.EVA_CLIP(**cfg, max_image_batch=2, max_text_batch=2)
    small.load_state_dict(sd, strict=True)
    small = small.to(DEV).eval()
    assert torch.equal(small.encode_image(frames), full)
    tokens = synthetic.make_tokens(5, cfg, seed=4).to(DEV)
    assert torch.equal(small.encode_text(tokens), model.encode_text(tokens))
    assert model.encode_image(frames[:0]).shape == (0, cfg["embed_dim"])  # empty batch


def test_engine_rebuilds_when_weights_change(tiny):
    cfg, sd, model = tiny
    frames = synthetic.make_frames(2, 224, seed=5).to(DEV)
    before = model.encode_image(frames)
    sd2 = synthetic.make_eva_state_dict(cfg, seed=123)
    model.load_state_dict(sd2, strict=True)
    after = model.encode_image(frames)
    assert not torch.allclose(before, after)
    model.load_state_dict(sd, strict=True)
    assert torch.equal(model.encode_image(frames), before)


def test_retrieval_pipeline_matches_oracle(tiny):
    cfg, sd, model = tiny
    frames = synthetic.make_frames(8, 224, seed=11)
    tokens = synthetic.make_tokens(6, cfg, seed=12)
    t_hat = retrieval.normalize(model.encode_text(tokens.to(DEV)))
    scores, v_hat = retrieval.encode_and_score(model, frames.to(DEV), 2, t_hat)
    with torch.no_grad():
        # scoring stage given identical embeddings: bit-stable ranking
        ref_from_ours = eva_oracle.similarity(t_hat.cpu(), v_hat.cpu())
        ref_full = eva_oracle.similarity(eva_oracle.normalize_text(eva_oracle.encode_text(sd, tokens, cfg)),
                                         eva_oracle.pool_normalize_video(eva_oracle.encode_image(sd, frames, cfg), 2))
    assert float((scores.cpu() - ref_from_ours).abs().max()) < 5e-7
    assert retrieval.topk(scores, None, 4) == [eva_oracle.rank_videos(r.tolist(), [f"{i:09d}" for i in range(4)])[:4]
                                               for r in ref_from_ours]
    assert float((scores.cpu() - ref_full).abs().max()) < 2e-2


def test_g14_encode_vs_reference_golden(hb, golden_dir):
    """BASELINE.json configs[0]: EVA-CLIP-g/14 on 8 random 224x224 frames — GPU path vs the reference's own CPU output."""
    cfg = synthetic.EVA_G14
    g = torch.load(os.path.join(golden_dir, "eva_g14.pt"))
    sd = synthetic.make_eva_state_dict(cfg, seed=0)
    model = eva_clip.EVA_CLIP(**cfg, max_image_batch=8, max_text_batch=8)
    model.load_state_dict(sd, strict=True)
    model = model.to(DEV).eval()
    frames = synthetic.make_frames(8, 224, seed=1)
    tokens = synthetic.make_tokens(4, cfg, seed=2)
    img = model.encode_image(frames.to(DEV))
    txt = model.encode_text(tokens.to(DEV))
    e_img, e_txt = rel(img, g["image"]), rel(txt, g["text"])
    cos = torch.nn.functional.cosine_similarity(img.cpu(), g["image"], dim=-1).min()
    print(f"g14: encode_image rel {e_img:.3e} (min cos {cos:.6f}), encode_text rel {e_txt:.3e}")
    assert e_img < TOL_G14_IMAGE and e_txt < TOL_G14_TEXT and cos > 0.9995
    # stock torch bf16 on the same weights, same box: we must not be worse
    sdb = {k: v.to(DEV).bfloat16() for k, v in sd.items() if k.startswith("visual.")}
    with torch.no_grad():
        stock = eva_oracle.encode_image(sdb, frames.to(DEV).bfloat16(), cfg)
    print(f"g14: stock torch bf16 rel {rel(stock, g['image']):.3e}")
    assert e_img <= rel(stock, g["image"]) * 1.05


def test_uint8_frames_match_cpu_preprocessing(tiny):
    """Raw uint8 frames: ToTensor + Normalize(mean, std) folded into the patch gather == the reference's CPU transform
    (EVA_clip/eva_clip.py:144-153) followed by encode_image — bit-identical embeddings."""
    cfg, sd, model = tiny
    g = torch.Generator().manual_seed(3)
    u8 = torch.randint(0, 256, (3, 3, 224, 224), generator=g, dtype=torch.uint8)
    mean = torch.tensor(model.visual.image_mean).view(1, 3, 1, 1)
    std = torch.tensor(model.visual.image_std).view(1, 3, 1, 1)
    f32 = (u8.float() / 255.0 - mean) / std          # ToTensor, then Normalize
    a = model.encode_image(u8.to(DEV))
    b = model.encode_image(f32.to(DEV))
    assert torch.equal(a, b)


def test_g14_full_size_properties(hb, golden_dir):
    """BASELINE configs[1] size (EVA-CLIP-g/14, 1024 frames).  (1) The 8 frames of the reference-generated golden (eva_g14.pt)
    sit INSIDE the 1024-frame batch, at the chunk / tile boundaries {0, 1, 383, 384, 511, 512, 1022, 1023}: their embeddings must
    match the reference's CPU output within the g/14 tolerance and be bit-identical to the 8-frame run.  (2) Size-independent
    properties: one chunk, chunks of 384 (ragged last chunk) and single frames give bit-identical embeddings; the retrieval
    scores of the pooled videos equal the CPU scoring of the same embeddings."""
    cfg = synthetic.EVA_G14
    sd = synthetic.make_eva_state_dict(cfg, seed=0)   # CPU generator: the weights the golden was made with
    sd = {k: v.to(DEV) for k, v in sd.items()}
    big = eva_clip.EVA_CLIP(**cfg, max_image_batch=1024, max_text_batch=8)
    big.load_state_dict(sd, strict=True)
    big = big.to(DEV).eval()
    frames = synthetic.make_frames(1024, 224, seed=77, device=DEV)
    g = torch.load(os.path.join(golden_dir, "eva_g14.pt"))
    gold_frames = synthetic.make_frames(8, 224, seed=1).to(DEV)
    pos = [0, 1, 383, 384, 511, 512, 1022, 1023]
    frames[pos] = gold_frames
    full = big.encode_image(frames)
    assert torch.isfinite(full).all()
    e_in_batch = rel(full[pos], g["image"])
    print(f"g14 @1024: golden frames inside the 1024-frame batch: rel {e_in_batch:.3e}")
    assert e_in_batch < TOL_G14_IMAGE
    assert torch.equal(full[pos], big.encode_image(gold_frames))
    small = eva_clip.EVA_CLIP(**cfg, max_image_batch=384, max_text_batch=8)
    small.load_state_dict(sd, strict=True)
    small = small.to(DEV).eval()
    del sd
    assert torch.equal(small.encode_image(frames), full)
    for i in (0, 383, 384, 1023):
        assert torch.equal(small.encode_image(frames[i:i + 1]), full[i:i + 1])
    v_hat = retrieval.pool_normalize(full, 32)
    t_hat = eva_oracle.normalize_text(torch.randn(64, cfg["embed_dim"], generator=torch.Generator().manual_seed(5)))
    scores = retrieval.similarity(t_hat.to(DEV), v_hat)
    ref = eva_oracle.similarity(t_hat, eva_oracle.pool_normalize_video(full.cpu(), 32))
    assert float((scores.cpu() - ref).abs().max()) < 5e-7


def test_layernorm_fold_matches_separate_layernorm(hb, golden_dir):
    """LayerNorm folded into the QKV / fc1 GEMM epilogues (default) vs separate LayerNorm kernels: both within tolerance of the
    reference golden, and the folded path is not less accurate (tiny config and EVA-CLIP-g/14)."""
    for cfg, gname, n in ((synthetic.EVA_TINY, "eva_tiny.pt", 4), (synthetic.EVA_G14, "eva_g14.pt", 8)):
        g = torch.load(os.path.join(golden_dir, gname))
        sd = synthetic.make_eva_state_dict(cfg, seed=0)
        frames = synthetic.make_frames(n, 224, seed=1).to(DEV)
        errs = {}
        for fold in (1, 0):
            _lib.check(hb.hb_debug_set(b"ln_fold", fold))
            model = eva_clip.EVA_CLIP(**cfg, max_image_batch=8, max_text_batch=8)
            model.load_state_dict(sd, strict=True)
            model = model.to(DEV).eval()
            errs[fold] = rel(model.encode_image(frames), g["image"])
            del model
        _lib.check(hb.hb_debug_set(b"ln_fold", 1))
        print(f"{gname}: rel err folded {errs[1]:.3e}, separate LayerNorm {errs[0]:.3e}")
        assert errs[1] < (TOL_TINY_IMAGE if cfg is synthetic.EVA_TINY else TOL_G14_IMAGE)
        assert errs[1] <= errs[0] * 1.25 + 1e-4


def test_schedule_and_tiling_knobs(hb):
    """Dynamic vs static tile hand-out only changes WHEN a tile runs: encode_image must be bit-identical.  Balanced vs plain
    N tiles also regroups the per-row LayerNorm partial sums (one slot per tile half), i.e. the fp32 summation order of the
    row statistics: results agree to bf16-rounding noise, each setting is reproducible run to run.  (EVA-CLIP-g/14 shapes,
    2 layers, 64 frames so that every GEMM has more tiles than workers and the atomic scheduler really hands tiles out.)"""
    cfg = dict(synthetic.EVA_G14)
    cfg["vision_cfg"] = dict(cfg["vision_cfg"], layers=2)
    sd = synthetic.make_eva_state_dict(cfg, seed=0)
    model = eva_clip.EVA_CLIP(**cfg, max_image_batch=64, max_text_batch=8)
    model.load_state_dict(sd, strict=True)
    model = model.to(DEV).eval()
    frames = synthetic.make_frames(64, 224, seed=4).to(DEV)
    outs = {}
    try:
        for dyn in (1, 0):
            for bal in (1, 0):
                _lib.check(hb.hb_debug_set(b"gemm_dynamic_schedule", dyn))
                _lib.check(hb.hb_debug_set(b"gemm_balanced_tiles", bal))
                outs[(dyn, bal)] = model.encode_image(frames).clone()
                assert torch.equal(outs[(dyn, bal)], model.encode_image(frames)), "not reproducible run to run"
    finally:
        _lib.check(hb.hb_debug_set(b"gemm_dynamic_schedule", 1))
        _lib.check(hb.hb_debug_set(b"gemm_balanced_tiles", 1))
    assert torch.isfinite(outs[(1, 1)]).all()
    assert torch.equal(outs[(1, 1)], outs[(0, 1)]) and torch.equal(outs[(1, 0)], outs[(0, 0)])
    assert rel(outs[(1, 0)], outs[(1, 1)]) < 3e-3
