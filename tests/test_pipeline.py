"""In-memory end-to-end chain: hirest_b200.pipeline on the GPU vs the CPU restatement of the reference's
disk-based chain (oracle/pipeline_oracle.py, run.py:383-490).  Everything that crosses a task boundary is an integer or a
string, so the bar is equality."""
import os

import numpy as np
import pytest
import torch

from hirest_b200 import pipeline, synthetic


def test_frame_timestamp_conversions():
    """hirest_dataset.py:12-68 semantics, including the float truncation of linspace bins the reference inherits."""
    for dur in (1, 2, 40, 81, 200, 1234):
        bins = np.linspace(0, dur - 1, dur)
        for f in (0, dur // 3, dur - 1):
            assert pipeline.frame_index_to_timestamp(f, dur, n_frames=-1) == int(bins[f])
        for t in (0, 0.5, 3, dur / 2, dur - 1, dur + 5):
            assert pipeline.timestamp_to_frame_index(t, dur, n_frames=-1) == int(min(np.digitize(t, bins, right=True), dur - 1))
    # fixed 32-frame grid (README example in hirest_dataset.py:24-32)
    assert pipeline.timestamp_to_frame_index(100, 200, n_frames=32) == 16
    assert pipeline.frame_index_to_timestamp(31, 200, n_frames=32) == 199


@pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="reference checkout only exists in the build container")
def test_frame_conversions_match_reference():
    import importlib.util
    import sys
    import types

    # hirest_dataset.py imports heavy dependencies at module level; execute only its two pure functions
    src = open("/root/reference/hirest_dataset.py").read()
    start = src.index("def timestamp_to_frame_index")
    end = src.index("class ", start)
    mod = types.ModuleType("ref_conv")
    mod.__dict__["np"] = np
    exec(compile(src[start:end], "hirest_dataset_excerpt", "exec"), mod.__dict__)
    for dur in (7, 81, 200, 999):
        for n in (-1, 32):
            nf = dur if n < 0 else n
            for f in range(0, nf, max(1, nf // 9)):
                assert pipeline.frame_index_to_timestamp(f, dur, n_frames=n) == mod.frame_index_to_timestamp(f, dur, n_frames=n)
            for t in np.linspace(0, dur + 3, 23):
                assert pipeline.timestamp_to_frame_index(float(t), dur, n_frames=n) == mod.timestamp_to_frame_index(float(t), dur, n_frames=n)


def test_collate_pads_like_the_reference():
    items = []
    for i, T in enumerate((5, 9, 7)):
        items.append({"task": "moment_retrieval", "prompt": "p", "fname": f"v{i}", "video_duration": T, "vis_feats": torch.ones(T, 4) * (i + 1),
                      "asr_feats": torch.ones(T, 3), "clip_text_ids": torch.zeros(77, dtype=torch.long),
                      "video_mask": torch.ones(T, dtype=torch.long), "moment_mask": torch.ones(T, dtype=torch.long)})
    b = pipeline.collate(items)
    assert b["vis_feats"].shape == (3, 9, 4) and b["asr_feats"].shape == (3, 9, 3)
    assert b["vis_mask"].sum(1).tolist() == [5, 9, 7] and b["vis_mask"].dtype == torch.int64
    assert float(b["vis_feats"][0, 5:].abs().sum()) == 0.0 and b["video_fnames"] == ["v0", "v1", "v2"]


class TableText:
    """clip_model stand-in: encode_text looks the prompt's feature up by the id stored in token slot 1."""

    def __init__(self, table):
        self.table = table

    def encode_text(self, ids):
        return self.table[ids[:, 1].cpu()].to(ids.device)


def _write_vocab(tmp_path):
    vocab = [f"[unused{i}]" for i in range(30522)]
    vocab[0], vocab[100], vocab[101], vocab[102], vocab[103] = "[PAD]", "[UNK]", "[CLS]", "[SEP]", "[MASK]"
    for i in range(1000, 30522):
        vocab[i] = f"w{i}"
    p = tmp_path / "vocab.txt"
    p.write_text("\n".join(vocab) + "\n")
    return str(p), vocab


@pytest.mark.gpu
def test_end_to_end_chain_matches_reference_flow(hb, tmp_path):
    from hirest_b200 import moment
    from oracle import pipeline_oracle as po

    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(11)
    prompts = ["make tea", "fix bike"]
    table = torch.randn(len(prompts), 1024, generator=g)
    test, feats, videos = {}, {}, []
    k = 0
    for pi, p in enumerate(prompts):
        test[p] = {}
        for _ in range(3):
            T = int(torch.randint(40, 90, (1,), generator=g))
            fn = f"vid{k}"
            k += 1
            vis = torch.randn(T, 1024, generator=g)
            vis = vis / vis.norm(dim=-1, keepdim=True)
            asr = torch.randn(T, 384, generator=g)
            feats[fn] = {"vis_feats": vis, "asr_feats": asr, "text_feat": {p: table[pi]}}
            test[p][fn] = {"video_duration": T}
            ids = torch.zeros(77, dtype=torch.long)
            ids[0], ids[1], ids[2] = 49406, pi, 49407
            videos.append({"prompt": p, "fname": fn, "video_duration": T, "vis_feats": vis, "asr_feats": asr, "clip_text_ids": ids})
    vocab_path, vocab = _write_vocab(tmp_path)
    sd = synthetic.make_moment_state_dict(seed=3)
    m = moment.MomentModel(-1, 384, moment.default_args(bert_vocab_path=vocab_path), clip_model=TableText(table), max_rows=4 * 96, max_batch=4)
    m.load_state_dict(sd, strict=True)
    m = m.to(dev)
    got = pipeline.run_end_to_end(m, videos, batch_size=4, num_beams=3)
    with torch.no_grad():
        ref = po.run_end_to_end(sd, test, feats, vocab, batch_size=4, num_beams=3)
    assert got["moment_retrieval"] == ref["moment_retrieval"]
    assert {k: v["pred_bounds"] for k, v in got["moment_segmentation"].items()} == {k: v["pred_bounds"] for k, v in ref["moment_segmentation"].items()}
    for p in ref["final"]:
        for fn, a in ref["final"][p].items():
            assert got["final"][p][fn]["bounds"] == a["bounds"], (p, fn)
            assert got["final"][p][fn]["steps"] == a["steps"], (p, fn)   # step bounds (ints) and captions (strings)
    n_steps = sum(len(a["steps"]) for p in got["final"] for a in got["final"][p].values())
    assert n_steps >= 8 and all(s["heading"] for p in got["final"] for a in got["final"][p].values() for s in a["steps"])
    # captions do not depend on how the step items are batched: one search over all of them gives the same dictionaries
    m2 = moment.MomentModel(-1, 384, moment.default_args(bert_vocab_path=vocab_path), clip_model=TableText(table), max_rows=4 * 96, max_batch=64)
    m2.load_state_dict(sd, strict=True)
    assert pipeline.run_end_to_end(m2.to(dev), videos, batch_size=4, num_beams=3, caption_batch_size=64) == got
    # ... nor on whether the next batch is collated and copied by the worker thread (side stream) or inline
    assert pipeline.run_end_to_end(m, videos, batch_size=4, num_beams=3, prefetch=False) == got


def test_collate_into_supplied_storage_clears_the_padding():
    """feature_alloc (pinned host memory on the GPU path) hands out recycled, uninitialised storage: long items get their padding
    cleared row by row, short ones the whole batch — either way the batch equals the default zero-padded one."""
    g = torch.Generator().manual_seed(5)

    def dirty(shape, dtype):
        return torch.full(shape, float("nan"), dtype=dtype)

    for rows, width in ((range(3, 9), 16), ((40, 70, 55), 2048)):   # short items / items past the per-row clearing threshold
        items = [{"fname": f"v{i}", "prompt": "p", "video_duration": n, "task": "moment_retrieval",
                  "vis_feats": torch.randn(n, width, generator=g), "asr_feats": torch.randn(n, 8, generator=g),
                  "video_mask": torch.ones(n, dtype=torch.long), "moment_mask": torch.ones(n, dtype=torch.long),
                  "clip_text_ids": torch.zeros(77, dtype=torch.long)} for i, n in enumerate(rows)]
        a, b = pipeline.collate(items), pipeline.collate(items, feature_alloc=dirty)
        for k in ("vis_feats", "asr_feats", "vis_mask", "moment_mask"):
            assert torch.equal(a[k], b[k]), k


def test_prefetch_thread_changes_nothing_and_propagates_errors():
    """The worker thread that prepares batch i + 1 while the model runs batch i: same dictionaries as inline preparation, batches
    reach the model in order, and an exception raised while preparing a batch surfaces in the caller."""
    g = torch.Generator().manual_seed(3)
    vids = [{"prompt": f"p{i % 3}", "fname": f"v{i}", "video_duration": 30 + i, "vis_feats": torch.randn(30 + i, 8, generator=g),
             "asr_feats": torch.randn(30 + i, 4, generator=g), "clip_text_ids": torch.zeros(77, dtype=torch.long)} for i in range(11)]

    class Fake:
        def __init__(self):
            self.seen = []

        def test_step(self, b, **kw):
            task, B = b["tasks"][0], b["vis_feats"].shape[0]
            self.seen.append((task, tuple(b["video_fnames"])))
            if task == "moment_retrieval":
                return {"prediction": [[int(n) // 6, int(n) - 3] for n in b["vis_mask"].sum(1)]}
            if task == "moment_segmentation":
                out = []
                for i in range(B):
                    idx = b["moment_mask"][i].nonzero()[:, 0]
                    out.append(list(range(int(idx[0]), int(idx[-1]), 7)) + [int(idx[-1])])
                return {"prediction": out}
            return {"prediction": [f"c{float(b['vis_feats'][i].sum()):.4f}" for i in range(B)]}

    a, b = Fake(), Fake()
    ra = pipeline.run_end_to_end(a, vids, batch_size=4, num_beams=3, caption_batch_size=5, prefetch=True)
    rb = pipeline.run_end_to_end(b, vids, batch_size=4, num_beams=3, caption_batch_size=5, prefetch=False)
    assert ra == rb and a.seen == b.seen and len(a.seen) >= 3 + 3 + 2

    bad = [dict(v) for v in vids]
    bad[8]["vis_feats"] = torch.zeros(bad[8]["vis_feats"].shape[0], 9)   # third batch (items are grouped by prompt): collate fails inside the worker thread
    c = Fake()
    with pytest.raises(RuntimeError):
        pipeline.run_end_to_end(c, bad, batch_size=4, prefetch=True)
    assert len(c.seen) == 2   # the two good batches ran before the error surfaced

    class Boom(Fake):
        def test_step(self, b, **kw):
            if len(self.seen) == 1:
                raise RuntimeError("model failed")
            return super().test_step(b, **kw)

    with pytest.raises(RuntimeError, match="model failed"):   # the pending prepared batch is dropped, the pool shuts down
        pipeline.run_end_to_end(Boom(), vids, batch_size=4, prefetch=True)


def test_prompts_are_tokenized_when_ids_are_missing():
    """run_end_to_end fills clip_text_ids through the tokenize callable (once per distinct prompt) before the first model call."""
    calls = []

    def fake_tokenize(prompt):
        calls.append(prompt)
        ids = torch.zeros(1, 77, dtype=torch.long)
        ids[0, 0], ids[0, 1], ids[0, 2] = 49406, len(prompt), 49407
        return ids

    class Stop(Exception):
        pass

    class Probe:
        def test_step(self, batch, **kw):
            assert batch["clip_text_ids"].shape == (3, 77) and batch["clip_text_ids"][:, 1].tolist() == [5, 5, 8]
            raise Stop

    vids = [{"prompt": p, "fname": f"v{i}", "video_duration": 6, "vis_feats": torch.zeros(6, 4), "asr_feats": torch.zeros(6, 3)}
            for i, p in enumerate(["aaaaa", "aaaaa", "bbbbbbbb"])]
    with pytest.raises(Stop):
        pipeline.run_end_to_end(Probe(), vids, tokenize=fake_tokenize)
    assert calls == ["aaaaa", "bbbbbbbb"]
    with pytest.raises(ValueError):
        pipeline.run_end_to_end(Probe(), vids)


@pytest.mark.gpu
def test_cfg5_chain_eight_videos_match_oracle_with_real_text_tower(hb, tmp_path):
    """BASELINE configs[4] shapes (120-600 frame videos, beam 3, 48 words): 8 videos through the in-memory chain with the repo's own
    precise EVA text tower, against the CPU restatement of the reference flow fed with the ORACLE's text features — bounds and step
    boundaries must be identical, captions identical up to beam near-ties (see the comment at the end)."""
    from hirest_b200 import eva_clip, moment
    from oracle import eva_oracle, pipeline_oracle as po

    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(17)
    prompt_ids = synthetic.make_tokens(3, synthetic.EVA_G14, seed=77)
    clip_sd = synthetic.make_chain_clip_state_dict()
    with torch.no_grad():
        tfeat = eva_oracle.encode_text(clip_sd, prompt_ids, synthetic.CHAIN_CLIP)
    test, feats, videos = {}, {}, []
    for k in range(8):
        pi = k // 3                      # grouped by prompt, like the annotation file
        p = f"prompt {pi}"
        T = int(torch.randint(120, 601, (1,), generator=g))
        vis = torch.randn(T, 1024, generator=g)
        vis = vis / vis.norm(dim=-1, keepdim=True)
        asr = torch.randn(T, 384, generator=g)
        fn = f"vid{k:02d}"
        videos.append({"prompt": p, "fname": fn, "video_duration": T, "vis_feats": vis, "asr_feats": asr, "clip_text_ids": prompt_ids[pi]})
        test.setdefault(p, {})[fn] = {"video_duration": T}
        feats[fn] = {"vis_feats": vis, "asr_feats": asr, "text_feat": {p: tfeat[pi]}}
    vocab_path, vocab = _write_vocab(tmp_path)
    clip = eva_clip.EVA_CLIP(**synthetic.CHAIN_CLIP)
    clip.load_state_dict(clip_sd, strict=True)
    sd = synthetic.make_moment_state_dict(seed=3)
    m = moment.MomentModel(-1, 384, moment.default_args(bert_vocab_path=vocab_path), clip_model=clip.to(dev).eval(), max_rows=8 * 600, max_batch=8)
    m.load_state_dict(sd, strict=False)
    m = m.to(dev)
    got = pipeline.run_end_to_end(m, videos, batch_size=8, num_beams=3)
    with torch.no_grad():
        ref = po.run_end_to_end(sd, test, feats, vocab, batch_size=8, num_beams=3)
    assert got["moment_retrieval"] == ref["moment_retrieval"]
    assert {k: v["pred_bounds"] for k, v in got["moment_segmentation"].items()} == {k: v["pred_bounds"] for k, v in ref["moment_segmentation"].items()}
    bad = []
    for p in ref["final"]:
        for fn, a in ref["final"][p].items():
            assert got["final"][p][fn]["bounds"] == a["bounds"], (p, fn)
            assert [s["absolute_bounds"] for s in got["final"][p][fn]["steps"]] == [s["absolute_bounds"] for s in a["steps"]], (p, fn)
            for sg, sr in zip(got["final"][p][fn]["steps"], a["steps"]):
                if sg["heading"] != sr["heading"]:
                    wg, wr = sg["heading"].split(), sr["heading"].split()
                    first = next((i for i, (x, y) in enumerate(zip(wg, wr)) if x != y), min(len(wg), len(wr)))
                    bad.append((fn, sg["index"], first, len(wg), len(wr)))
    n_steps = sum(len(a["steps"]) for p in got["final"] for a in got["final"][p].values())
    print(f"cfg5 x 8 videos: {n_steps} steps captioned, captions differing from the CPU oracle: {bad}")
    # Moment bounds and step boundaries must be identical.  Captions: with RANDOM-INIT decoder weights the 30522-way distributions
    # are nearly flat, so beam search meets genuine near-ties; the fp32-accurate GPU path (1e-5 relative on the text feature, 2^-17 on
    # the GEMMs) may take the other branch of such a tie.  Measured: 96 of 97 captions identical (one diverges at word 11 of 48, with
    # the CUDA-core attention as well as with the tensor-core one); the gate allows 5 %.
    assert n_steps >= 16 and len(bad) <= 0.05 * n_steps
