"""Generate tests/golden/*.pt by running the UNMODIFIED reference (imported from /root/reference) on the
seeded synthetic weights / inputs of oracle/weights.py.  Run in the build container only:

    python oracle/make_golden.py [--full]

Outputs (small tensors only; inputs and weights are regenerated from seeds by the tests):
  tests/golden/eva_tiny.pt     EVA_TINY: encode_image(4 frames), residual taps, encode_text(6 queries)
  tests/golden/eva_g14.pt      EVA-CLIP-g/14 (BASELINE config #1): encode_image(8 frames), encode_text(4)   [--full]
"""
import argparse
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_shims, weights  # noqa: E402


def run_reference(cfg, n_frames, n_text, taps=()):
    model = ref_shims.reference_eva_clip(cfg)
    sd = weights.make_eva_state_dict(cfg, seed=0)
    missing = model.load_state_dict(sd, strict=True)
    print("load_state_dict:", missing)
    frames = weights.make_frames(n_frames, cfg["vision_cfg"]["image_size"], seed=1)
    tokens = weights.make_tokens(n_text, cfg, seed=2)
    out = {}
    with torch.no_grad():
        t0 = time.time()
        out["image"] = model.encode_image(frames).float().clone()
        out["image_seconds"] = time.time() - t0
        out["text"] = model.encode_text(tokens).float().clone()
        if taps:
            x = model.visual.patch_embed(frames)
            cls = model.visual.cls_token.expand(x.shape[0], -1, -1)
            x = torch.cat((cls, x), dim=1) + model.visual.pos_embed
            out["tap0"] = x.clone()
            for i, blk in enumerate(model.visual.blocks):
                x = blk(x)
                if (i + 1) in taps:
                    out[f"tap{i + 1}"] = x.clone()
    out["n_frames"], out["n_text"] = n_frames, n_text
    out["torch_version"] = str(torch.__version__)
    return out


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--full", action="store_true", help="also run EVA-CLIP-g/14 (4.5 GB weights, ~1 min)")
    args = ap.parse_args()
    torch.manual_seed(0)
    gdir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(gdir, exist_ok=True)
    out = run_reference(weights.EVA_TINY, 4, 6, taps=(1, 3))
    torch.save(out, os.path.join(gdir, "eva_tiny.pt"))
    print("tiny image", out["image"].shape, float(out["image"].abs().mean()), "text", float(out["text"].abs().mean()))
    if args.full:
        out = run_reference(weights.EVA_G14, 8, 4)
        torch.save(out, os.path.join(gdir, "eva_g14.pt"))
        print("g14 image", out["image"].shape, float(out["image"].abs().mean()), "seconds", out["image_seconds"])
