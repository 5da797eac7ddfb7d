"""GPU parity of hb_resize_crop_u8 (through the C ABI) against the CPU oracle and the committed torchvision-on-PIL goldens:
integer work, so the bar is bit-exact."""
import hashlib
import os

import numpy as np
import pytest
import torch

from hirest_b200 import eva_clip, preprocess, synthetic
from oracle import make_golden_preprocess as mg
from oracle import preprocess_oracle as po

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(params=[3, 2, 1], ids=["words_extract_imad", "words_dp4a", "byte_loads"], autouse=True)
def resize_version(request):
    """Every test runs on all three kernels: v1 (byte loads + IMAD), v2 (planar word loads + dp4a on byte-plane weights) and v3 (planar
    word loads + byte extraction + IMAD)."""
    from hirest_b200 import _lib

    _lib.load()
    _lib.debug_set("resize_version", request.param)
    yield request.param
    _lib.debug_set("resize_version", 1)


def _sha(a):
    return np.frombuffer(hashlib.sha256(np.ascontiguousarray(a).tobytes()).digest(), np.uint8)


def test_matches_reference_goldens(hb, golden_dir):
    g = np.load(os.path.join(golden_dir, "preprocess.npz"))
    for i, (h, w, kind, seed) in enumerate(g["cases"]):
        img = mg.make_image(int(h), int(w), "noise" if kind == 0 else "smooth", int(seed))
        out = preprocess.resize_center_crop(torch.from_numpy(img).to(DEV), 224).cpu().numpy()
        assert out.shape == (3, 224, 224)
        assert np.array_equal(_sha(out), g[f"sha_u8_{i}"]), f"case {i} ({h}x{w})"
        if f"u8_{i}" in g:
            assert np.array_equal(out, g[f"u8_{i}"])


@pytest.mark.parametrize("h,w,size", [(97, 131, 224), (360, 640, 224), (500, 333, 224), (224, 224, 224), (50, 70, 32), (33, 20, 16),
                                      (2160, 3840, 224)])
def test_batch_matches_oracle(hb, h, w, size):
    rng = np.random.default_rng(h * 1000 + w)
    B = 3
    imgs = rng.integers(0, 256, (B, h, w, 3), dtype=np.uint8)
    out = preprocess.resize_center_crop(torch.from_numpy(imgs).to(DEV), size).cpu().numpy()
    for b in range(B):
        assert np.array_equal(out[b], po.resize_center_crop_u8(imgs[b], size)), f"frame {b}"


def test_unaligned_views_and_empty_batch(hb):
    """Rows of odd-width frames start at every byte alignment; a sliced batch starts at an odd address."""
    rng = np.random.default_rng(7)
    imgs = rng.integers(0, 256, (5, 121, 187, 3), dtype=np.uint8)
    flat = torch.zeros(imgs.size + 1, dtype=torch.uint8, device=DEV)
    flat[1:] = torch.from_numpy(imgs).to(DEV).flatten()
    view = flat[1:].view(5, 121, 187, 3)
    assert view.data_ptr() % 2 == 1
    out = preprocess.resize_center_crop(view, 224).cpu().numpy()
    for b in range(5):
        assert np.array_equal(out[b], po.resize_center_crop_u8(imgs[b], 224))
    assert preprocess.resize_center_crop(torch.zeros(0, 8, 8, 3, dtype=torch.uint8, device=DEV), 224).shape == (0, 3, 224, 224)


def test_full_size_properties(hb):
    """BASELINE-size batch (1024 frames of 360x640): batch independence (every copy of a frame gives the same bytes, equal
    to the oracle's), chunked == whole, identity when no resampling is needed, constant frames stay constant."""
    rng = np.random.default_rng(11)
    base = rng.integers(0, 256, (4, 360, 640, 3), dtype=np.uint8)
    frames = torch.from_numpy(base).to(DEV).repeat(256, 1, 1, 1)
    out = preprocess.resize_center_crop(frames, 224)
    assert out.shape == (1024, 3, 224, 224)
    ref = torch.stack([torch.from_numpy(po.resize_center_crop_u8(base[i], 224)) for i in range(4)]).to(DEV)
    assert torch.equal(out.view(256, 4, 3, 224, 224), ref.expand(256, -1, -1, -1, -1))
    assert torch.equal(preprocess.resize_center_crop(frames[100:300], 224), out[100:300])
    sq = torch.from_numpy(rng.integers(0, 256, (2, 224, 224, 3), dtype=np.uint8)).to(DEV)
    assert torch.equal(preprocess.resize_center_crop(sq, 224), sq.permute(0, 3, 1, 2))
    const = torch.full((1, 480, 854, 3), 201, dtype=torch.uint8, device=DEV)
    assert bool((preprocess.resize_center_crop(const, 224) == 201).all())


def test_encode_frames_equals_cpu_preprocessing_then_encode(hb):
    """Raw decoded frames -> GPU resize/crop -> uint8 encode_image  ==  reference CPU transform (fp32 tensor) -> encode_image,
    bit for bit (the normalisation folded into the patch gather is the same fp32 arithmetic)."""
    cfg = synthetic.EVA_TINY
    model = eva_clip.EVA_CLIP(**cfg)
    model.load_state_dict(synthetic.make_eva_state_dict(cfg, seed=0), strict=True)
    model = model.to(DEV).eval()
    model.visual.image_mean, model.visual.image_std = mg.OPENAI_DATASET_MEAN, mg.OPENAI_DATASET_STD
    rng = np.random.default_rng(3)
    imgs = rng.integers(0, 256, (3, 270, 480, 3), dtype=np.uint8)
    cpu = np.stack([po.to_tensor_normalize(po.resize_center_crop_u8(im, 224), mg.OPENAI_DATASET_MEAN, mg.OPENAI_DATASET_STD)
                    for im in imgs])
    a = preprocess.encode_frames(model, torch.from_numpy(imgs).to(DEV))
    b = model.encode_image(torch.from_numpy(cpu).to(DEV))
    assert torch.equal(a, b)
