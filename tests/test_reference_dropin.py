"""(1) retrieval.rank_videos / recall_at_k against results of the reference's own evaluate.evaluate_video_retrieval
(tests/golden/evaluate_retrieval.json, oracle/make_golden_evaluate.py) — travels, runs everywhere.
(2) Drop-in wiring: the UNMODIFIED reference modeling.MomentModel constructed with hirest_b200.eva_clip installed as the `eva_clip`
module (build container only: needs /root/reference).  There is no GPU there, so the proof stops where the reference's own
test_step calls clip_model.encode_text and lands in this package's "no CPU fallback" error."""
import json
import os
import sys

import pytest
import torch

from hirest_b200 import retrieval, synthetic


def test_rank_and_recall_match_reference_evaluate(golden_dir):
    with open(os.path.join(golden_dir, "evaluate_retrieval.json")) as f:
        g = json.load(f)
    names, prompts = g["video_names"], g["prompts"]
    scores = torch.tensor(g["scores"], dtype=torch.float64)
    for q in range(len(prompts)):
        assert [names[j] for j in retrieval.rank_videos(scores[q].numpy(), names)[:50]] == g["ranked_top50"][q]
    got = retrieval.recall_at_k(scores.numpy(), names, [g["gt"][p] for p in prompts])
    assert got == g["results"]["all"]
    for cat in ("cooking", "repair"):   # per-category results = the same function on the category's prompts
        idx = [q for q, p in enumerate(prompts) if g["categories"][p] == cat]
        assert retrieval.recall_at_k(scores[idx].numpy(), names, [g["gt"][prompts[q]] for q in idx]) == g["results"][cat]


@pytest.mark.skipif(not os.path.exists("/root/reference/modeling.py"), reason="the reference checkout exists only in the build container")
def test_reference_moment_model_runs_on_this_package_unchanged(tmp_path):
    import hirest_b200
    from hirest_b200 import eva_clip
    from oracle import ref_moment

    dst = ref_moment.prepare_copy("/tmp/hirest_ref_copy_dropin")
    ref_moment.install_stubs()
    saved = {k: sys.modules.get(k) for k in ("eva_clip", "modeling", "args")}
    cwd = os.getcwd()
    old_cfg = eva_clip.get_model_config("EVA_CLIP_g_14")
    try:
        # a tiny architecture under the g/14 name + its checkpoint where modeling.py:117 looks for it (the real one is a 4.5 GB download)
        eva_clip.add_model_config("EVA_CLIP_g_14", synthetic.EVA_TINY)
        torch.save({"model": synthetic.make_eva_state_dict(synthetic.EVA_TINY, seed=0)}, os.path.join(dst, "pretrained_weights", "eva_clip_psz14.pt"))
        hirest_b200.install_as_eva_clip()
        os.chdir(dst)
        for p in (dst, os.path.join(dst, "clip4caption")):
            if p not in sys.path:
                sys.path.insert(0, p)
        for k in ("modeling", "args"):
            sys.modules.pop(k, None)
        import args as ref_args
        import modeling as ref_modeling   # the reference file, unmodified: `from eva_clip import build_eva_model_and_transforms`

        a = ref_args.get_parser().parse_args(["--data_dir", "x", "--video_feature_dir", "x"])
        model = ref_modeling.MomentModel(n_frames=-1, asr_dim=384, args=a).eval()
        assert isinstance(model.clip_model, eva_clip.EVA_CLIP)
        assert callable(model.clip_preprocess)
        keys = set(model.state_dict())
        assert "clip_model.visual.cls_token" in keys and "clip_model.text.token_embedding.weight" in keys
        assert all(not p.requires_grad for p in model.clip_model.parameters())           # freeze_clip (modeling.py:127-130) worked
        batch = synthetic.make_moment_batch(2, 12, seed=1)
        batch["clip_text_ids"] = synthetic.make_tokens(2, synthetic.EVA_TINY, seed=3)
        batch["tasks"] = ["moment_retrieval"] * 2
        with pytest.raises(RuntimeError, match="no CPU fallback"):   # reached through modeling.py:286 -> our encode_text
            model.test_step(batch)
    finally:
        os.chdir(cwd)
        eva_clip.add_model_config("EVA_CLIP_g_14", old_cfg)
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
