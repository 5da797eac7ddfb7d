"""Generates tests/golden/preprocess.npz by running the reference's own preprocessing stack — torchvision
Resize(224, BICUBIC) + CenterCrop(224) on PIL images, then ToTensor + Normalize with the EVA-CLIP mean/std
(EVA_clip/eva_clip.py:16-17, 120-153) — on seeded synthetic frames.  Run in the build container (Pillow + torchvision
are installed there; they are not needed on the GPU box):

    python oracle/make_golden_preprocess.py

Stored per case: the seed/shape that regenerate the input (``make_image``), the uint8 [3,224,224] output of
Resize+CenterCrop for the small cases, and a SHA-256 of it for every case (large inputs are regenerated, not stored).
"""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

OPENAI_DATASET_MEAN = (0.48145466, 0.4578275, 0.40821073)   # EVA_clip/eva_clip.py:16
OPENAI_DATASET_STD = (0.26862954, 0.26130258, 0.27577711)   # EVA_clip/eva_clip.py:17

# (height, width, kind): "noise" = uniform random bytes (worst case for the saturating shift), "smooth" = low-frequency image
CASES = [(224, 224, "noise"), (240, 320, "noise"), (360, 640, "smooth"), (720, 1280, "noise"), (480, 360, "smooth"),
         (225, 224, "noise"), (224, 500, "noise"), (1080, 1920, "smooth"), (300, 227, "noise"), (64, 48, "noise"),
         (223, 400, "smooth"), (1, 1, "noise"), (2, 900, "noise")]
STORE_FULL = {(240, 320), (480, 360), (64, 48), (225, 224)}


def make_image(h: int, w: int, kind: str, seed: int) -> np.ndarray:
    rng = np.random.default_rng(seed)
    if kind == "noise":
        return rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float64)
    ph = rng.uniform(0, 6.28, (3, 4))
    img = np.stack([127.5 + 80 * np.sin(xx / (17.0 + c) + ph[c, 0]) * np.cos(yy / (23.0 - c) + ph[c, 1])
                    + 47 * np.sin((xx + yy) / 5.0 + ph[c, 2]) for c in range(3)], axis=-1)
    return np.clip(np.rint(img), 0, 255).astype(np.uint8)


def main():
    import torch
    from PIL import Image
    from torchvision.transforms import CenterCrop, Compose, InterpolationMode, Normalize, Resize, ToTensor

    geom = Compose([Resize(224, interpolation=InterpolationMode.BICUBIC), CenterCrop(224)])
    full = Compose([geom, lambda im: im.convert("RGB"), ToTensor(), Normalize(OPENAI_DATASET_MEAN, OPENAI_DATASET_STD)])
    out = {"cases": np.array([(h, w, 0 if k == "noise" else 1, 100 + i) for i, (h, w, k) in enumerate(CASES)], np.int64)}
    for i, (h, w, kind) in enumerate(CASES):
        img = make_image(h, w, kind, 100 + i)
        u8 = np.ascontiguousarray(np.asarray(geom(Image.fromarray(img))).transpose(2, 0, 1))
        f32 = full(Image.fromarray(img)).numpy()
        out[f"sha_u8_{i}"] = np.frombuffer(hashlib.sha256(u8.tobytes()).digest(), np.uint8)
        out[f"sha_f32_{i}"] = np.frombuffer(hashlib.sha256(np.ascontiguousarray(f32).tobytes()).digest(), np.uint8)
        if (h, w) in STORE_FULL:
            out[f"u8_{i}"] = u8
        print(f"case {i}: {h}x{w} {kind}: u8 sha {hashlib.sha256(u8.tobytes()).hexdigest()[:16]}")
    import PIL
    import torchvision
    out["versions"] = np.array([f"Pillow {PIL.__version__}", f"torchvision {torchvision.__version__}", f"torch {torch.__version__}"])
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "preprocess.npz"), **out)


if __name__ == "__main__":
    main()
