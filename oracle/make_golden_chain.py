"""Generate tests/golden/chain.pt: the UNMODIFIED reference MomentModel (oracle/ref_moment.py) running its OWN EVA-CLIP text
tower (EVA_clip/eva_model.py:177-250, g/14 text config, seeded weights) on real clip_text_ids -- encode_text -> MomentModel
(modeling.py:286,364,568) -- for moment retrieval, moment segmentation and step captioning.  Build container only:

    python oracle/make_golden_chain.py

The visual tower of the clip model is the tiny config (MomentModel never calls encode_image); the text tower is EVA-CLIP-g/14's.
Cases: "small" 3 x 40 frames (all three tasks, beam 3), "cfg4" = BASELINE configs[3]: 64 clips x 300 frames (MR + MS),
"noasr" = the ASR-free model variant (modeling.py:28-35) on 3 x 40 (MR / MS logits through forward_*; the reference's own
test_* methods are broken without ASR).
Stored: reference text features, MR [start, end], MS boundary lists, best-hypothesis caption ids.  Weights / batches are
regenerated from seeds by the tests (hirest_b200.synthetic)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from hirest_b200 import synthetic  # noqa: E402
from oracle import ref_moment  # noqa: E402


def chain_clip_cfg():
    return {"embed_dim": 1024, "vision_cfg": dict(synthetic.EVA_TINY["vision_cfg"]), "text_cfg": dict(synthetic.EVA_G14["text_cfg"])}


def chain_clip_state_dict():
    return synthetic.make_chain_clip_state_dict()


def run_tasks(model, batch, tasks, num_beams=3, captured=None):
    out = {}
    B = batch["vis_feats"].shape[0]
    with torch.no_grad():
        out["text_feat"] = model.clip_model.encode_text(batch["clip_text_ids"]).float().clone()
        for task in tasks:
            b = dict(batch)
            b["tasks"] = [task] * B
            if task == "moment_retrieval":
                b["moment_retrieval_start_target"] = b["moment_retrieval_end_target"] = None
                out["mr_pred"] = model.test_step(b)["prediction"]
            elif task == "moment_segmentation":
                out["ms_pred"] = model.test_step(b)["prediction"]
            else:
                out["caption_text"] = model.test_step(b, num_beams=num_beams)["prediction"]
                out["caption_ids"] = list(captured["hyp"])
                out["caption_scores"] = list(captured["scores"])
    return out


if __name__ == "__main__":
    cfg = chain_clip_cfg()
    model, args = ref_moment.build_reference_moment_model(num_beams=3, clip_cfg=cfg)
    sd = synthetic.make_moment_state_dict(seed=3)
    clip_sd = chain_clip_state_dict()
    full = dict(sd)
    full.update({"clip_model." + k: v for k, v in clip_sd.items()})
    print(model.load_state_dict(full, strict=True))
    import modeling as ref_modeling

    captured = {}
    orig = ref_modeling.collect_hypothesis_and_scores

    def capture(beams, n_best):
        hyp, sc = orig(beams, n_best)
        captured["hyp"] = [h[0] for h in hyp]
        captured["scores"] = [float(x[0]) for x in sc]
        return hyp, sc

    ref_modeling.collect_hypothesis_and_scores = capture
    out = {"torch_version": str(torch.__version__)}
    small = synthetic.make_chain_batch(3, 40, seed=11)
    out["small"] = run_tasks(model, small, ("moment_retrieval", "moment_segmentation", "step_captioning"), 3, captured)
    print("small", out["small"]["mr_pred"], out["small"]["ms_pred"], [len(h) for h in out["small"]["caption_ids"]])
    cfg4 = synthetic.make_chain_batch(64, 300, seed=12)
    out["cfg4"] = run_tasks(model, cfg4, ("moment_retrieval", "moment_segmentation"))
    print("cfg4 mr[:4]", out["cfg4"]["mr_pred"][:4], "ms[0]", out["cfg4"]["ms_pred"][0])

    # ASR-free variant: same seeded weights minus asr_enc_layer.*
    model2, _ = ref_moment.build_reference_moment_model(num_beams=3, clip_cfg=cfg, asr_dim=-1)
    full2 = {k: v for k, v in full.items() if not k.startswith("asr_enc_layer.")}
    print(model2.load_state_dict(full2, strict=True))
    # the reference's test_* methods raise UnboundLocalError without ASR (modeling.py:280-289: asr_feats is only bound under
    # `if self.use_asr`), so the ASR-free variant is pinned through the forward_* methods its train path uses (asr_feats=None)
    with torch.no_grad():
        tf = model2.clip_model.encode_text(small["clip_text_ids"]).float()
        mr = model2.forward_moment_retrieval(small["vis_feats"], tf, video_mask=small["vis_mask"], moment_mask=small["moment_mask"],
                                             asr_feats=None)
        bm = torch.zeros_like(small["moment_mask"])
        bm[:, 3] = 1
        ms_logits = model2.forward_moment_segmentation(small["vis_feats"], tf, small["vis_mask"], small["moment_mask"],
                                                       asr_feats=None, boundary_mask=bm)
    out["noasr"] = {"text_feat": tf.clone(), "start_logits": mr["start_logits"].clone(), "end_logits": mr["end_logits"].clone(),
                    "ms_logits": ms_logits.clone()}
    print("noasr start argmax", mr["start_logits"].argmax(-1).tolist())
    path = os.path.join(ROOT, "tests", "golden", "chain.pt")
    torch.save(out, path)
    print("saved", os.path.getsize(path) / 1e6, "MB")
