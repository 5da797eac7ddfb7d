"""MomentModel (shared encoder + MR / MS decoders + trim_feats) on the GPU vs the reference-generated golden and the CPU oracle.

The GEMMs run as 3-term split-bf16 tcgen05 GEMMs and attention in fp32, so the bar is: features / logits within 2e-4 relative
of the fp32 reference, and the INTEGER outputs (MR [start, end], MS boundary lists, trimmed frame selection) identical."""
import os

import pytest
import torch

from hirest_b200 import _lib, moment, synthetic
from oracle import moment_oracle as mo

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda:0")
CASES = {"small": (3, 40, 5), "t300": (2, 300, 6)}


class FixedText:
    """Stands in for clip_model: encode_text returns the batch's fixed text features (as oracle/make_golden_moment.py does)."""

    def __init__(self):
        self.feat = None

    def encode_text(self, ids):
        return self.feat.to(ids.device)


def rel(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return float((a - b).norm() / b.norm())


@pytest.fixture(scope="module")
def model(hb):
    clip = FixedText()
    m = moment.MomentModel(-1, 384, moment.default_args(), clip_model=clip, max_rows=1024, max_batch=8)
    m.load_state_dict(synthetic.make_moment_state_dict(seed=3), strict=True)
    return m.to(DEV), clip


@pytest.fixture(scope="module")
def golden(golden_dir):
    return torch.load(os.path.join(golden_dir, "moment.pt"))


@pytest.mark.parametrize("case", list(CASES))
def test_shared_encoder_and_logits(model, golden, case):
    m, clip = model
    B, T, seed = CASES[case]
    b = synthetic.make_moment_batch(B, T, seed=seed)
    g = golden[case]
    feats = m.foward_moment_shared(b["vis_feats"], b["text_feat"], b["vis_mask"], moment_mask=b["moment_mask"], asr_feats=b["asr_feats"])
    ref = g["shared"]
    got = feats.cpu() if case == "small" else feats.cpu()[:, :8]
    assert rel(got, ref) < 2e-4
    logits, _ = m._forward(b["vis_feats"].to(DEV), b["text_feat"].to(DEV), b["asr_feats"].to(DEV), b["vis_mask"].to(DEV),
                           b["moment_mask"].to(DEV))
    assert float((logits[..., 0].cpu() - g["start_logits"]).abs().max()) < 2e-3
    assert float((logits[..., 1].cpu() - g["end_logits"]).abs().max()) < 2e-3
    bm = torch.zeros_like(b["moment_mask"])
    bm[:, 3] = 1
    logits, _ = m._forward(b["vis_feats"].to(DEV), b["text_feat"].to(DEV), b["asr_feats"].to(DEV), b["vis_mask"].to(DEV),
                           b["moment_mask"].to(DEV), bm.to(DEV))
    assert float((logits[..., 2].cpu() - g["ms_logits"]).abs().max()) < 2e-3


@pytest.mark.parametrize("case", list(CASES))
def test_moment_retrieval_and_segmentation_predictions_match_reference(model, golden, case):
    m, clip = model
    B, T, seed = CASES[case]
    b = synthetic.make_moment_batch(B, T, seed=seed)
    clip.feat = b["text_feat"]
    b["tasks"] = ["moment_retrieval"] * B
    assert m.test_step(b)["prediction"] == golden[case]["mr_pred"]
    b["tasks"] = ["moment_segmentation"] * B
    out = m.test_step(b)
    assert out["prediction"] == golden[case]["ms_pred"]
    assert out["raw_predictions"] == out["prediction"]
    # the loop stops after an iteration that accepted no step (later iterations would repeat it): same predictions as all 20
    # forwards; in the 40-frame case iteration 10 is the first that changes nothing (CPU oracle), seen one iteration late
    early = m.ms_iterations_run
    try:
        m.ms_early_exit = False
        assert m.test_step(b)["prediction"] == golden[case]["ms_pred"] and m.ms_iterations_run == 20
    finally:
        m.ms_early_exit = True
    assert early == (12 if case == "small" else 20), early


def test_ms_step_kernel_matches_reference_control_flow(hb):
    """The on-device region growing (modeling.py:399-433) against the verbatim Python restatement, incl. edge cases."""
    g = torch.Generator().manual_seed(0)
    B, T = 6, 57
    logits = torch.randn(B, T, 3, generator=g) * 2
    mm = torch.ones(B, T, dtype=torch.long)
    mm[0, :10] = 0
    mm[1, 20:] = 0
    mm[2] = 0                      # empty moment: uniform softmax over -FLT_MAX -> max prob 1/T, accepted unless l == 0
    logits[3, 0, 2] = 50.0         # peak at frame 0 -> left bound 0 -> rejected (modeling.py:425-426)
    logits[4, :, 2] = 0.0          # flat -> grows to both ends
    bm = torch.zeros(B, T, dtype=torch.long)
    mm_d, bm_d = mm.to(DEV), bm.to(DEV)
    steps = torch.zeros(B, 4, 2, dtype=torch.int32, device=DEV)
    ns = torch.zeros(B, dtype=torch.int32, device=DEV)
    probs = torch.empty(B, T, device=DEV)
    _lib.check(hb.hb_moment_ms_step(logits.to(DEV).data_ptr(), mm_d.data_ptr(), bm_d.data_ptr(), steps.data_ptr(), ns.data_ptr(), 4,
                                    B, T, 0.5, probs.data_ptr(), _lib.stream_ptr()))
    torch.cuda.synchronize()
    lg = logits[..., 2].clone()
    lg[mm == 0] = -torch.finfo(torch.float32).max
    ref_p = lg.softmax(dim=1)
    assert float((probs.cpu() - ref_p).abs().max()) < 1e-6
    for b in range(B):
        sc = ref_p[b].tolist()
        mi = int(ref_p[b].argmax())
        exp_mm, exp_bm, exp_steps = mm[b].clone(), bm[b].clone(), []
        if not sc[mi] < 0.00001:
            l, r = mo.grow_region(sc, mi, 0.5)
            if not (l == 0 or r == 0):
                exp_mm[l:r + 1] = 0
                exp_bm[l] = 1
                exp_bm[r] = 1
                exp_steps = [[l, r]]
        assert torch.equal(mm_d[b].cpu(), exp_mm) and torch.equal(bm_d[b].cpu(), exp_bm), b
        assert steps[b, :int(ns[b])].cpu().tolist() == exp_steps, b


def test_mr_decode_masks_padded_frames(hb):
    B, T = 3, 33
    logits = torch.zeros(B, T, 3)
    logits[0, 30, 0] = 9.0   # padded frame wins unless masked
    logits[0, 5, 0] = 1.0
    logits[1, 7, 1] = 2.0
    logits[1, 9, 1] = 2.0    # tie -> first index
    vm = torch.ones(B, T, dtype=torch.long)
    vm[0, 20:] = 0
    pred = torch.empty(B, 2, dtype=torch.int64, device=DEV)
    _lib.check(hb.hb_moment_mr_decode(logits.to(DEV).data_ptr(), vm.to(DEV).data_ptr(), pred.data_ptr(), B, T, _lib.stream_ptr()))
    lg = logits.clone()
    lg[vm == 0] = -1e10
    assert pred.cpu().tolist() == torch.stack([lg[..., 0].argmax(1), lg[..., 1].argmax(1)], -1).tolist()


def test_trim_feats(model):
    m, _ = model
    g = torch.Generator().manual_seed(2)
    x = torch.randn(5, 50, 768, generator=g)
    mask = torch.zeros(5, 50, dtype=torch.long)
    mask[0, 3:10] = 1      # 7 frames  -> repeat-padded to 20
    mask[1, 0:45] = 1      # 45 frames -> first 20
    mask[2, 10:30] = 1     # exactly 20
    mask[3, 7] = 1         # a single frame
    #   [4]: empty moment -> zeros (modeling.py:537-549: `for j in range(N)` never runs with N = 0, x stays zeros)
    got = m.trim_feats(x, mask).cpu()
    assert torch.equal(got, mo.trim_feats(x, mask, 20))
    assert not got[4].any()


def test_long_clip_matches_oracle(hb):
    """T = 700 frames (> one 256-key attention tile, ragged second sample): shared features and MR prediction vs the CPU oracle."""
    clip = FixedText()
    m = moment.MomentModel(-1, 384, moment.default_args(), clip_model=clip, max_rows=1400, max_batch=2)
    sd = synthetic.make_moment_state_dict(seed=3)
    m.load_state_dict(sd, strict=True)
    m = m.to(DEV)
    b = synthetic.make_moment_batch(2, 700, seed=21)
    clip.feat = b["text_feat"]
    with torch.no_grad():
        ref = mo.moment_shared(sd, b["vis_feats"], b["text_feat"], b["vis_mask"], b["moment_mask"], b["asr_feats"])
        ref_pred, _, _ = mo.test_moment_retrieval(sd, b, b["text_feat"])
    got = m.foward_moment_shared(b["vis_feats"], b["text_feat"], b["vis_mask"], moment_mask=b["moment_mask"], asr_feats=b["asr_feats"])
    assert rel(got, ref) < 2e-4
    b["tasks"] = ["moment_retrieval"] * 2
    assert m.test_step(b)["prediction"] == ref_pred


def test_rejects_clips_beyond_position_embeddings(model):
    m, _ = model
    b = synthetic.make_moment_batch(1, 2049, seed=1, ragged=False)
    m2 = moment.MomentModel(-1, 384, moment.default_args(), clip_model=FixedText(), max_rows=2049, max_batch=1)
    m2.load_state_dict(synthetic.make_moment_state_dict(seed=3), strict=True)
    m2 = m2.to(DEV)
    with pytest.raises(RuntimeError, match="max_position_embeddings"):
        m2.foward_moment_shared(b["vis_feats"], b["text_feat"], b["vis_mask"], moment_mask=b["moment_mask"], asr_feats=b["asr_feats"])


def _write_vocab(tmp_path):
    vocab = [f"[unused{i}]" for i in range(30522)]
    vocab[0], vocab[100], vocab[101], vocab[102], vocab[103] = "[PAD]", "[UNK]", "[CLS]", "[SEP]", "[MASK]"
    for i in range(1000, 30522):
        vocab[i] = f"w{i}"
    p = tmp_path / "vocab.txt"
    p.write_text("\n".join(vocab) + "\n")
    return str(p)


def test_step_captioning_token_ids_match_reference(model, golden, tmp_path):
    """Beam-3 decode of the small batch: KV-cached on-device beam search vs the reference's full-prefix recompute
    (modeling.py:556-632): identical best-hypothesis token ids and strings."""
    m, clip = model
    b = synthetic.make_moment_batch(3, 40, seed=5)
    clip.feat = b["text_feat"]
    b["tasks"] = ["step_captioning"] * 3
    m.args.bert_vocab_path = _write_vocab(tmp_path)
    m._vocab_list = None
    out = m.test_step(b, num_beams=3)
    g = golden["caption"]
    assert out["token_ids"] == g["ids"]
    assert out["prediction"] == g["text"]


def test_step_captioning_early_finish(hb, golden, tmp_path):
    """[SEP] logit boosted: instances finish at different steps (48, 48, 2, 17 tokens in the reference) and are frozen."""
    clip = FixedText()
    m = moment.MomentModel(-1, 384, moment.default_args(bert_vocab_path=_write_vocab(tmp_path)), clip_model=clip, max_rows=1024, max_batch=8)
    sd = synthetic.make_moment_state_dict(seed=3)
    bias = sd["clip4cap_model.decoder.classifier.cls.predictions.bias"].clone()
    bias[102] += 2.0
    sd["clip4cap_model.decoder.classifier.cls.predictions.bias"] = bias
    m.load_state_dict(sd, strict=True)
    m = m.to(DEV)
    b = synthetic.make_moment_batch(4, 40, seed=7)
    clip.feat = b["text_feat"]
    b["tasks"] = ["step_captioning"] * 4
    out = m.test_step(b, num_beams=3)
    g = golden["caption_eos"]
    assert [len(x) for x in g["ids"]] == [48, 48, 2, 17]
    assert out["token_ids"] == g["ids"]
    assert out["prediction"] == g["text"]


@pytest.mark.parametrize("beam", [1, 5])
def test_step_captioning_other_beam_widths_match_oracle(hb, tmp_path, beam):
    """Beam widths other than the golden's 3 (run.py's default is 5; 1 = greedy): best-hypothesis token ids vs the CPU restatement
    of the reference beam search (oracle/caption_oracle.py) on the early-finishing weights."""
    from oracle import caption_oracle as co

    clip = FixedText()
    m = moment.MomentModel(-1, 384, moment.default_args(bert_vocab_path=_write_vocab(tmp_path)), clip_model=clip, max_rows=1024, max_batch=8)
    sd = synthetic.make_moment_state_dict(seed=3)
    bias = sd["clip4cap_model.decoder.classifier.cls.predictions.bias"].clone()
    bias[102] += 2.0
    sd["clip4cap_model.decoder.classifier.cls.predictions.bias"] = bias
    m.load_state_dict(sd, strict=True)
    m = m.to(DEV)
    b = synthetic.make_moment_batch(2, 40, seed=11)
    clip.feat = b["text_feat"]
    b["tasks"] = ["step_captioning"] * 2
    out = m.test_step(b, num_beams=beam)
    with torch.no_grad():
        ref_ids, _ = co.test_step_captioning(sd, b, b["text_feat"], beam=beam)
    assert out["token_ids"] == ref_ids


def test_decoder_graph_replay_and_kernel_variants_agree(hb, golden, tmp_path):
    """The same beam search run eagerly (first search of a shape), while its steps are captured into CUDA graphs (second) and as
    graph replays (third) gives the reference's token ids every time and the same kernel count; so do the variants without graphs
    and without the split-K GEMM + finish kernels (include/hirest_b200_debug.h)."""
    clip = FixedText()
    sd = synthetic.make_moment_state_dict(seed=3)
    bias = sd["clip4cap_model.decoder.classifier.cls.predictions.bias"].clone()
    bias[102] += 2.0
    sd["clip4cap_model.decoder.classifier.cls.predictions.bias"] = bias
    b = synthetic.make_moment_batch(4, 40, seed=7)
    clip.feat = b["text_feat"]
    b["tasks"] = ["step_captioning"] * 4
    g = golden["caption_eos"]

    def fresh():
        m = moment.MomentModel(-1, 384, moment.default_args(bert_vocab_path=_write_vocab(tmp_path)), clip_model=clip, max_rows=1024, max_batch=8)
        m.load_state_dict(sd, strict=True)
        return m.to(DEV)

    try:
        m = fresh()
        counts = []
        for _ in range(3):   # eager, capture + replay, replay
            n0 = hb.hb_launch_count()
            assert m.test_step(b, num_beams=3)["token_ids"] == g["ids"]
            counts.append(hb.hb_launch_count() - n0)
        assert counts[0] == counts[1] == counts[2], counts
        for key in ("decoder_graphs", "decoder_split_k", "decoder_kv_index"):
            _lib.debug_set(key, 0)   # (switches accumulate: the last variant runs with all three off)
            m = fresh()
            for _ in range(2):
                assert m.test_step(b, num_beams=3)["token_ids"] == g["ids"], key
    finally:
        _lib.debug_set("decoder_graphs", 1)
        _lib.debug_set("decoder_split_k", 6)
        _lib.debug_set("decoder_kv_index", 1)


def test_decoder_graph_sets_survive_alternating_shapes(hb, golden, tmp_path):
    """A job's full batches and its ragged last batch alternate between two (instances, beam, frames) shapes: each shape keeps its
    own step graphs (no re-capture after the second search of a shape: the kernel count per search stays constant), more shapes than
    the handle keeps evict the least recently used set, and every search still returns the reference's token ids."""
    clip = FixedText()
    sd = synthetic.make_moment_state_dict(seed=3)
    bias = sd["clip4cap_model.decoder.classifier.cls.predictions.bias"].clone()
    bias[102] += 2.0
    sd["clip4cap_model.decoder.classifier.cls.predictions.bias"] = bias
    b = synthetic.make_moment_batch(4, 40, seed=7)
    b["tasks"] = ["step_captioning"] * 4
    g = golden["caption_eos"]
    m = moment.MomentModel(-1, 384, moment.default_args(bert_vocab_path=_write_vocab(tmp_path)), clip_model=clip, max_rows=1024, max_batch=8)
    m.load_state_dict(sd, strict=True)
    m = m.to(DEV)

    def search(n):
        sub = {k: (v[:n] if torch.is_tensor(v) or isinstance(v, list) else v) for k, v in b.items()}
        clip.feat = b["text_feat"][:n]
        n0 = hb.hb_launch_count()
        ids = m.test_step(sub, num_beams=3)["token_ids"]
        assert ids == g["ids"][:n], n
        return hb.hb_launch_count() - n0

    counts = {4: [], 3: []}
    for _ in range(3):
        for n in (4, 3):
            counts[n].append(search(n))
    assert len(set(counts[4])) == 1 and len(set(counts[3])) == 1, counts
    # more shapes than a handle keeps (8): beam widths 1..5 x two batch sizes, twice over, then the first shapes again
    for _ in range(2):
        for beam in (1, 2, 4, 5):
            for n in (1, 2, 4):
                sub = {k: (v[:n] if torch.is_tensor(v) or isinstance(v, list) else v) for k, v in b.items()}
                clip.feat = b["text_feat"][:n]
                ids = m.test_step(sub, num_beams=beam)["token_ids"]
                assert len(ids) == n
    assert search(4) == counts[4][0] and search(3) == counts[3][0]


def test_config4_size_determinism_and_oracle_subset(hb):
    """BASELINE configs[3] size (64 clips x 300 frames): two runs are bit-identical, each clip's result is independent of the batch
    it is in, and the first two clips match the CPU oracle's moment-retrieval prediction."""
    clip = FixedText()
    m = moment.MomentModel(-1, 384, moment.default_args(), clip_model=clip, max_rows=64 * 300, max_batch=64)
    sd = synthetic.make_moment_state_dict(seed=3)
    m.load_state_dict(sd, strict=True)
    m = m.to(DEV)
    b = synthetic.make_moment_batch(64, 300, seed=31)
    a1 = m.foward_moment_shared(b["vis_feats"], b["text_feat"], b["vis_mask"], moment_mask=b["moment_mask"], asr_feats=b["asr_feats"])
    a2 = m.foward_moment_shared(b["vis_feats"], b["text_feat"], b["vis_mask"], moment_mask=b["moment_mask"], asr_feats=b["asr_feats"])
    assert torch.equal(a1, a2)
    sub = {k: (v[:2] if torch.is_tensor(v) else v) for k, v in b.items()}
    a3 = m.foward_moment_shared(sub["vis_feats"], sub["text_feat"], sub["vis_mask"], moment_mask=sub["moment_mask"], asr_feats=sub["asr_feats"])
    assert torch.equal(a3, a1[:2])
    clip.feat = b["text_feat"]
    b["tasks"] = ["moment_retrieval"] * 64
    pred = m.test_step(b)["prediction"]
    with torch.no_grad():
        ref_pred, _, _ = mo.test_moment_retrieval(sd, sub, sub["text_feat"])
    assert pred[:2] == ref_pred
