"""Packed video-feature cache for the cached-feature retrieval path.

Reference: ``inference_video_retrieval.py:298-327`` loads one ``{video_id}.pt`` pickle per video with ``torch.load`` (64 s for
4282 files in the reference's own log), subsamples ``np.linspace(0, n-1, F).astype(int)`` frames, mean-pools and
L2-normalises on the CPU; ``extract_features.py:60-70`` / ``inference_video_retrieval.py:275-280`` write those pickles.

Here the same tensors live in ONE file — a small JSON header (ids, row offsets, dim) followed by the fp32 rows of all
videos back to back, 4096-byte aligned — that is memory-mapped, copied to the GPU in one piece, and reduced by one kernel
(``hb_subsample_pool_normalize``: on-the-fly linspace gather → mean → L2 normalise, one CTA per video).  Values are the
reference's fp32 features bit for bit; only the container changes.

File layout (little endian):
    bytes 0..7    magic  b"HBFEAT01"
    bytes 8..15   uint64 header length H
    bytes 16..16+H  JSON {"dim": E, "dtype": "float32", "video_ids": [...], "offsets": [0, T0, T0+T1, ...]}
    zero padding to the next multiple of 4096
    float32 [sum T, E] rows
"""
from __future__ import annotations

import json
import os
from typing import Dict, Iterable, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib

MAGIC = b"HBFEAT01"
ALIGN = 4096


def pack_features(items: Iterable[Tuple[str, torch.Tensor]], path: str) -> None:
    """Write ``(video_id, features[T, E])`` pairs (any float dtype; stored as fp32 like the reference's ``.float()``)."""
    ids: List[str] = []
    offsets = [0]
    chunks = []
    dim = None
    for vid, feats in items:
        a = feats.detach().to("cpu", torch.float32).contiguous().numpy()
        if a.ndim != 2 or a.shape[0] < 1:
            raise ValueError(f"{vid}: expected [T>=1, E] features, got {a.shape}")
        if dim is None:
            dim = a.shape[1]
        elif a.shape[1] != dim:
            raise ValueError(f"{vid}: feature dim {a.shape[1]} != {dim}")
        ids.append(str(vid))
        offsets.append(offsets[-1] + a.shape[0])
        chunks.append(a)
    header = json.dumps({"dim": dim or 0, "dtype": "float32", "video_ids": ids, "offsets": offsets}).encode()
    data_start = (16 + len(header) + ALIGN - 1) // ALIGN * ALIGN
    tmp = path + ".tmp"
    with open(tmp, "wb") as f:
        f.write(MAGIC)
        f.write(np.uint64(len(header)).tobytes())
        f.write(header)
        f.write(b"\0" * (data_start - 16 - len(header)))
        for a in chunks:
            f.write(a.tobytes())
    os.replace(tmp, path)


def pack_feature_dir(feature_dir: str, video_ids: Sequence[str], path: str) -> None:
    """Convert the reference's per-video ``{video_id}.pt`` files (inference_video_retrieval.py:275-280) into one blob."""
    pack_features(((v, torch.load(os.path.join(feature_dir, f"{v}.pt"), map_location="cpu")) for v in video_ids), path)


class FeatureStore:
    """Memory-mapped view of a packed feature file."""

    def __init__(self, path: str):
        with open(path, "rb") as f:
            if f.read(8) != MAGIC:
                raise ValueError(f"{path}: not a hirest_b200 feature store")
            hlen = int(np.frombuffer(f.read(8), np.uint64)[0])
            meta = json.loads(f.read(hlen).decode())
        self.path = path
        self.dim: int = int(meta["dim"])
        self.video_ids: List[str] = list(meta["video_ids"])
        self.offsets = np.asarray(meta["offsets"], np.int64)
        data_start = (16 + hlen + ALIGN - 1) // ALIGN * ALIGN
        rows = int(self.offsets[-1])
        self.data = np.memmap(path, np.float32, "r", offset=data_start, shape=(rows, self.dim)) if rows else np.zeros((0, self.dim), np.float32)
        self._index: Dict[str, int] = {v: i for i, v in enumerate(self.video_ids)}

    def __len__(self) -> int:
        return len(self.video_ids)

    def features(self, video_id: str) -> torch.Tensor:
        """The tensor ``torch.load(f"{video_id}.pt")`` would have returned (a copy)."""
        i = self._index[video_id]
        return torch.from_numpy(np.array(self.data[self.offsets[i]:self.offsets[i + 1]]))

    def to_device(self, device, video_range: Optional[Tuple[int, int]] = None):
        """(features [rows, E] fp32, offsets int64 [V+1]) on ``device`` for videos ``[lo, hi)`` (default: all).
        One host→device copy of the mapped rows through a pinned staging buffer."""
        lo, hi = video_range if video_range is not None else (0, len(self))
        r0, r1 = int(self.offsets[lo]), int(self.offsets[hi])
        staging = torch.empty((r1 - r0, self.dim), dtype=torch.float32).pin_memory() if torch.cuda.is_available() else \
            torch.empty((r1 - r0, self.dim), dtype=torch.float32)
        if r1 > r0:
            np.copyto(staging.numpy(), self.data[r0:r1])
        feats = staging.to(device, non_blocking=True)
        offs = torch.from_numpy(self.offsets[lo:hi + 1] - r0).to(device, non_blocking=True)
        return feats, offs


@torch.no_grad()
def pooled_video_embeddings(feats: torch.Tensor, offsets: torch.Tensor, n_model_frames: int) -> torch.Tensor:
    """Packed ``[sum T, E]`` fp32 features + ``[V+1]`` int64 offsets (both CUDA) → ``[V, E]`` L2-normalised video embeddings:
    the loop body of inference_video_retrieval.py:306-326 for every video in one launch.  ``n_model_frames <= 0`` pools all frames."""
    if not feats.is_cuda or not offsets.is_cuda:
        raise RuntimeError("hirest_b200.feature_store has no CPU path: call FeatureStore.to_device first")
    if feats.dtype != torch.float32 or offsets.dtype != torch.int64:
        raise TypeError("expected fp32 features and int64 offsets")
    feats, offsets = feats.contiguous(), offsets.contiguous()
    V = offsets.numel() - 1
    E = feats.shape[1]
    out = torch.empty((V, E), dtype=torch.float32, device=feats.device)
    lib = _lib.init(feats.device.index if feats.device.index is not None else torch.cuda.current_device())
    with torch.cuda.device(feats.device):
        _lib.check(lib.hb_subsample_pool_normalize(feats.data_ptr(), offsets.data_ptr(), V, int(n_model_frames), E, out.data_ptr(),
                                                   _lib.stream_ptr(feats.device)), "hb_subsample_pool_normalize")
    return out
