"""Generate tests/golden/moment_*.pt by running the UNMODIFIED reference MomentModel (oracle/ref_moment.py) on the seeded
synthetic weights / batches of hirest_b200.synthetic.  Build container only:   python oracle/make_golden_moment.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from hirest_b200 import synthetic  # noqa: E402
from oracle import ref_moment  # noqa: E402

if __name__ == "__main__":
    model, args = ref_moment.build_reference_moment_model(num_beams=3)
    sd = synthetic.make_moment_state_dict(seed=3)
    print(model.load_state_dict(sd, strict=True))
    out = {"torch_version": str(torch.__version__)}
    for name, (B, T, seed) in {"small": (3, 40, 5), "t300": (2, 300, 6)}.items():
        batch = synthetic.make_moment_batch(B, T, seed=seed)
        tf = batch["text_feat"]
        model.clip_model.encode_text = lambda ids, tf=tf: tf
        with torch.no_grad():
            feats = model.foward_moment_shared(batch["vis_feats"], tf, batch["vis_mask"], moment_mask=batch["moment_mask"],
                                               asr_feats=batch["asr_feats"])
            mr = model.forward_moment_retrieval(batch["vis_feats"], tf, video_mask=batch["vis_mask"],
                                                moment_mask=batch["moment_mask"], asr_feats=batch["asr_feats"])
            b2 = dict(batch)
            b2["tasks"] = ["moment_retrieval"] * B
            b2["moment_retrieval_start_target"] = b2["moment_retrieval_end_target"] = None
            mr_pred = model.test_step(b2)["prediction"]
            b3 = dict(batch)
            b3["tasks"] = ["moment_segmentation"] * B
            ms_pred = model.test_step(b3)["prediction"]
            bm = torch.zeros_like(batch["moment_mask"])
            bm[:, 3] = 1
            ms_logits = model.forward_moment_segmentation(batch["vis_feats"], tf, batch["vis_mask"], batch["moment_mask"],
                                                          asr_feats=batch["asr_feats"], boundary_mask=bm)
            trimmed = model.trim_feats(batch["vis_feats"], batch["moment_mask"], B, "cpu")
        out[name] = {"shared": feats.clone(), "start_logits": mr["start_logits"].clone(), "end_logits": mr["end_logits"].clone(),
                     "mr_pred": mr_pred, "ms_pred": ms_pred, "ms_logits": ms_logits.clone(), "trimmed_sum": trimmed.sum(-1).clone()}
        print(name, "mr", mr_pred, "ms", ms_pred)
    # step captioning (beam 3) on the small batch: best-hypothesis token ids + strings
    import modeling as ref_modeling
    captured = {}
    orig = ref_modeling.collect_hypothesis_and_scores

    def capture(beams, n_best):
        hyp, sc = orig(beams, n_best)
        captured["hyp"] = [h[0] for h in hyp]
        captured["scores"] = [float(x[0]) for x in sc]
        return hyp, sc

    ref_modeling.collect_hypothesis_and_scores = capture
    batch = synthetic.make_moment_batch(3, 40, seed=5)
    tf = batch["text_feat"]
    model.clip_model.encode_text = lambda ids, tf=tf: tf
    b4 = dict(batch)
    b4["tasks"] = ["step_captioning"] * 3
    with torch.no_grad():
        cap = model.test_step(b4, num_beams=3)["prediction"]
    out["caption"] = {"text": cap, "ids": captured["hyp"], "scores": captured["scores"]}
    print("captions", cap, [len(h) for h in captured["hyp"]], captured["scores"])
    # same, with the [SEP] logit boosted so that beams finish early at different steps (exercises the "done" path)
    sd2 = dict(sd)
    bias = sd["clip4cap_model.decoder.classifier.cls.predictions.bias"].clone()
    bias[102] += 2.0
    sd2["clip4cap_model.decoder.classifier.cls.predictions.bias"] = bias
    model.load_state_dict(sd2, strict=True)
    batch = synthetic.make_moment_batch(4, 40, seed=7)
    tf = batch["text_feat"]
    model.clip_model.encode_text = lambda ids, tf=tf: tf
    b5 = dict(batch)
    b5["tasks"] = ["step_captioning"] * 4
    with torch.no_grad():
        cap = model.test_step(b5, num_beams=3)["prediction"]
    out["caption_eos"] = {"text": cap, "ids": captured["hyp"], "scores": captured["scores"]}
    print("captions (eos boosted)", [len(h) for h in captured["hyp"]], captured["scores"])
    model.load_state_dict(sd, strict=True)
    # keep the fixture small: shared features only for the small case, logits for both
    out["t300"]["shared"] = out["t300"]["shared"][:, :8].clone()
    torch.save(out, os.path.join(ROOT, "tests", "golden", "moment.pt"))
    print("saved", os.path.getsize(os.path.join(ROOT, "tests", "golden", "moment.pt")) / 1e6, "MB")
