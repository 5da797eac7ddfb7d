"""The real encode_text -> MomentModel chain on the GPU: hirest_b200.moment.MomentModel with hirest_b200.eva_clip.EVA_CLIP as its
clip_model (EVA-CLIP-g/14 text tower, seeded weights, real clip_text_ids), against tests/golden/chain.pt, which
oracle/make_golden_chain.py produced by running the UNMODIFIED reference MomentModel with the reference's own EVA text tower
(modeling.py:286,364,568 -> EVA_clip/eva_model.py:232-250).

Bar: text features within 1e-4 relative of the fp32 reference (precise text tower: split-bf16 GEMMs + fp32 attention); every INTEGER
output -- MR [start, end], MS boundary lists, caption token ids -- identical, including all 64 clips of BASELINE configs[3]
(64 x 300 frames)."""
import os

import pytest
import torch

from hirest_b200 import eva_clip, moment, synthetic

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda:0")
TEXT_REL_CAP = 1e-4   # measured 1.0e-5 .. 1.3e-5 (see profiles/r02_chain_parity.txt); the bf16 tower is at 7.5e-3


def rel(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return float((a - b).norm() / b.norm())


@pytest.fixture(scope="module")
def golden(golden_dir):
    return torch.load(os.path.join(golden_dir, "chain.pt"))


@pytest.fixture(scope="module")
def clip(hb):
    m = eva_clip.EVA_CLIP(**synthetic.CHAIN_CLIP)
    m.load_state_dict(synthetic.make_chain_clip_state_dict(), strict=True)
    return m.to(DEV).eval()


def _moment(clip, asr_dim):
    m = moment.MomentModel(-1, asr_dim, moment.default_args(), clip_model=clip, max_rows=64 * 300, max_batch=64)
    sd = synthetic.make_moment_state_dict(seed=3)
    if asr_dim <= 0:
        sd = {k: v for k, v in sd.items() if not k.startswith("asr_enc_layer.")}
    sd.update({"clip_model." + k: v for k, v in synthetic.make_chain_clip_state_dict().items()})
    m.load_state_dict(sd, strict=True)   # the reference key layout incl. clip_model.* loads strictly
    return m.to(DEV)


@pytest.fixture(scope="module")
def model(clip):
    return _moment(clip, 384)


def test_text_tower_precise_matches_reference(clip, golden):
    for case, (B, T, seed) in {"small": (3, 40, 11), "cfg4": (64, 300, 12)}.items():
        ids = synthetic.make_chain_batch(B, T, seed)["clip_text_ids"]
        got = clip.encode_text(ids.to(DEV))
        e = rel(got, golden[case]["text_feat"])
        print(f"chain text tower ({case}): rel err {e:.3e}")
        assert e < TEXT_REL_CAP, (case, e)


def test_repeated_prompts_are_encoded_once_with_identical_features(clip, model):
    """A prompt's embedding does not depend on what else is in the batch (same bits alone, repeated, shuffled among others), so
    MomentModel encodes each distinct token row of a batch once and gathers: identical text features, 5 rows through the tower
    instead of 40."""
    ids = synthetic.make_chain_batch(5, 40, 21)["clip_text_ids"]
    perm = torch.tensor([3, 0, 0, 4, 1, 3, 2, 2] * 5)
    rep = ids[perm]
    full = clip.encode_text(rep.to(DEV))
    each = clip.encode_text(ids.to(DEV))
    assert torch.equal(full, each[perm.to(DEV)])
    for i in range(5):
        assert torch.equal(clip.encode_text(ids[i:i + 1].to(DEV))[0], each[i])
    b = synthetic.make_chain_batch(40, 40, 22)
    b["clip_text_ids"] = rep
    calls = []
    orig = clip.encode_text
    try:
        clip.encode_text = lambda t: (calls.append(t.shape[0]), orig(t))[1]
        model.dedup_prompts = True
        tf_dedup = model._inputs(b)[3]
        model.dedup_prompts = False
        tf_all = model._inputs(b)[3]
    finally:
        model.dedup_prompts = True
        del clip.encode_text
    assert calls == [5, 40] and torch.equal(tf_dedup, tf_all)


def test_text_tower_bf16_mode_is_the_looser_one(hb, golden):
    """precise=False keeps the plain-bf16 tower of the north_star wording; it must still be within the bf16 budget."""
    m = eva_clip.EVA_CLIP(**synthetic.CHAIN_CLIP, precise_text=False)
    m.load_state_dict(synthetic.make_chain_clip_state_dict(), strict=True)
    m = m.to(DEV).eval()
    ids = synthetic.make_chain_batch(3, 40, 11)["clip_text_ids"]
    e = rel(m.encode_text(ids.to(DEV)), golden["small"]["text_feat"])
    print(f"chain text tower bf16 mode: rel err {e:.3e}")
    assert 1e-4 < e < 1.1e-2, e


def test_small_chain_all_tasks_identical_to_reference(model, golden):
    g = golden["small"]
    b = synthetic.make_chain_batch(3, 40, 11)
    b["tasks"] = ["moment_retrieval"] * 3
    assert model.test_step(b)["prediction"] == g["mr_pred"]
    b["tasks"] = ["moment_segmentation"] * 3
    assert model.test_step(b)["prediction"] == g["ms_pred"]
    b["tasks"] = ["step_captioning"] * 3
    assert model.caption_token_ids(b, num_beams=3) == [list(x) for x in g["caption_ids"]]   # ids -> text needs a vocab file


def test_cfg4_64x300_mr_and_ms_identical_to_reference(model, golden):
    """BASELINE configs[3] at full size: all 64 clips, text features from the repo's own text tower."""
    g = golden["cfg4"]
    b = synthetic.make_chain_batch(64, 300, 12)
    b["tasks"] = ["moment_retrieval"] * 64
    mr = model.test_step(b)["prediction"]
    bad = [i for i in range(64) if mr[i] != g["mr_pred"][i]]
    assert not bad, (bad, [mr[i] for i in bad], [g["mr_pred"][i] for i in bad])
    b["tasks"] = ["moment_segmentation"] * 64
    ms = model.test_step(b)["prediction"]
    bad = [i for i in range(64) if ms[i] != g["ms_pred"][i]]
    assert not bad, (bad, [ms[i] for i in bad], [g["ms_pred"][i] for i in bad])


def test_asr_free_variant(clip, golden):
    """asr_dim <= 0 (modeling.py:28-35): no asr_enc_layer; logits against the reference's forward_* methods."""
    m = _moment(clip, -1)
    assert not any(k.startswith("asr_enc_layer") for k in m.state_dict())
    g = golden["noasr"]
    b = synthetic.make_chain_batch(3, 40, 11)
    tf = clip.encode_text(b["clip_text_ids"].to(DEV))
    assert rel(tf, g["text_feat"]) < TEXT_REL_CAP
    v, vm, mm = b["vis_feats"].to(DEV), b["vis_mask"].to(DEV), b["moment_mask"].to(DEV)
    logits, _ = m._forward(v, tf, None, vm, mm)
    assert float((logits[..., 0].cpu() - g["start_logits"]).abs().max()) < 2e-3
    assert float((logits[..., 1].cpu() - g["end_logits"]).abs().max()) < 2e-3
    assert logits[..., 0].argmax(-1).tolist() == g["start_logits"].argmax(-1).tolist()
    bm = torch.zeros_like(b["moment_mask"])
    bm[:, 3] = 1
    logits, _ = m._forward(v, tf, None, vm, mm, bm.to(DEV))
    assert float((logits[..., 2].cpu() - g["ms_logits"]).abs().max()) < 2e-3
    # the reference's test_* methods crash without ASR (UnboundLocalError, modeling.py:280-289); ours run
    b["tasks"] = ["moment_retrieval"] * 3
    assert len(m.test_step(b)["prediction"]) == 3
