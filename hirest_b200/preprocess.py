"""GPU frame preprocessing — the batch-level counterpart of the reference's per-image CPU transform.

Reference: ``image_transform`` EVA_clip/eva_clip.py:120-153 (torchvision ``Resize(224, BICUBIC)`` → ``CenterCrop(224)`` →
RGB → ``ToTensor`` → ``Normalize``), applied one PIL image at a time in the dataset workers
(inference_video_retrieval.py:43-49, extract_features.py:48-50).  At > 1700 frames/s per GPU that CPU path is the
bottleneck by two orders of magnitude, so here decoded frames go to the GPU as raw bytes and

* ``resize_center_crop`` (``hb_resize_crop_u8``) produces the uint8 ``[B,3,S,S]`` crop, **bit-identical** to
  torchvision-on-PIL (Pillow's 8-bit fixed-point bicubic resample, restated in hb_preproc.cu);
* ``ToTensor`` + ``Normalize`` are folded into the patch gather of ``encode_image`` (uint8 input path), which is bit-identical
  to normalising on the CPU and encoding the fp32 tensor.

No CPU fallback: a non-CUDA tensor raises.
"""
from __future__ import annotations

import ctypes as C
from typing import Tuple

import torch

from . import _lib


def resized_geometry(height: int, width: int, size: int = 224) -> Tuple[int, int, int, int]:
    """(new_h, new_w, top, left) of Resize(size) + CenterCrop(size) for a height x width frame (host only)."""
    lib = _lib.load()
    nh, nw, top, left = C.c_int(), C.c_int(), C.c_int(), C.c_int()
    _lib.check(lib.hb_resize_geometry(int(height), int(width), int(size), nh, nw, top, left), "hb_resize_geometry")
    return nh.value, nw.value, top.value, left.value


@torch.no_grad()
def resize_center_crop(frames: torch.Tensor, size: int = 224) -> torch.Tensor:
    """uint8 ``[B,H,W,3]`` (or ``[H,W,3]``) decoded RGB frames on a CUDA device → uint8 ``[B,3,size,size]``.

    Same bytes as ``CenterCrop(size)(Resize(size, BICUBIC)(PIL image))`` for every frame."""
    if frames.dtype != torch.uint8:
        raise TypeError(f"expected uint8 frames, got {frames.dtype}")
    if not frames.is_cuda:
        raise RuntimeError("hirest_b200.preprocess has no CPU path: move the decoded frames to the GPU first")
    squeeze = frames.dim() == 3
    if squeeze:
        frames = frames.unsqueeze(0)
    if frames.dim() != 4 or frames.shape[-1] != 3:
        raise ValueError(f"expected [B,H,W,3] frames, got {tuple(frames.shape)}")
    frames = frames.contiguous()
    B, H, W, _ = frames.shape
    dev = frames.device
    lib = _lib.init(dev.index if dev.index is not None else torch.cuda.current_device())
    out = torch.empty((B, 3, size, size), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        _lib.check(lib.hb_resize_crop_u8(frames.data_ptr(), B, H, W, int(size), out.data_ptr(), _lib.stream_ptr(dev)),
                   "hb_resize_crop_u8")
    return out[0] if squeeze else out


@torch.no_grad()
def encode_frames(model, frames: torch.Tensor) -> torch.Tensor:
    """Decoded uint8 ``[B,H,W,3]`` frames → ``[B,embed_dim]`` image embeddings: the reference's
    ``model.encode_image(torch.stack([preprocess(img) for img in frames]))`` with the preprocessing on the GPU."""
    return model.encode_image(resize_center_crop(frames, model.visual.image_size))
