"""Packed video-feature cache for the cached-feature retrieval path.

Reference: ``inference_video_retrieval.py:298-327`` loads one ``{video_id}.pt`` pickle per video with ``torch.load`` (64 s for
4282 files in the reference's own log), subsamples ``np.linspace(0, n-1, F).astype(int)`` frames, mean-pools and
L2-normalises on the CPU; ``extract_features.py:60-70`` / ``inference_video_retrieval.py:275-280`` write those pickles.

Here the same tensors live in ONE file — a small JSON header (ids, row offsets, dim) followed by the fp32 rows of all
videos back to back, 4096-byte aligned — that is memory-mapped, copied to the GPU in one piece, and reduced by one kernel
(``hb_subsample_pool_normalize``: on-the-fly linspace gather → mean → L2 normalise, one CTA per video).  Values are the
reference's fp32 features bit for bit; only the container changes.  ``dtype="bfloat16"`` (SURVEY.md §8(f) N2's wording) stores
the rows rounded to bf16: half the file, half the H2D copy, embeddings within ~2e-3 relative instead of bit-identical (the
features are unit-norm CLIP embeddings, `extract_features.py:64`); the default stays fp32 because retrieval parity is defined on
bit-exact top-k.

The same container holds the ASR sentence features (``{video_id}.pt`` under ``asr_feature_dir``, one row per subtitle sentence,
hirest_dataset.py:366-368) with the sentences' start / end seconds in the header (``extra``), for ``hirest_b200.dataset.warp_asr``.

File layout (little endian):
    bytes 0..7    magic  b"HBFEAT01"
    bytes 8..15   uint64 header length H
    bytes 16..16+H  JSON {"dim": E, "dtype": "float32", "video_ids": [...], "offsets": [0, T0, T0+T1, ...]}
    zero padding to the next multiple of 4096
    float32 (or bfloat16) [sum T, E] rows
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


def pack_features(items: Iterable[Tuple[str, torch.Tensor]], path: str, dtype: str = "float32", extra: Optional[dict] = None,
                  allow_empty: bool = False) -> None:
    """Write ``(video_id, features[T, E])`` pairs (any float dtype; stored as fp32 like the reference's ``.float()``, or rounded to
    bf16 with ``dtype="bfloat16"``).  ``extra``: JSON-serialisable side data kept in the header (ASR stores: sentence seconds)."""
    if dtype not in ("float32", "bfloat16"):
        raise ValueError("dtype must be 'float32' or 'bfloat16'")
    ids: List[str] = []
    offsets = [0]
    chunks = []
    dim = None
    for vid, feats in items:
        t = feats.detach().to("cpu", torch.float32).contiguous()
        a = t.numpy() if dtype == "float32" else t.to(torch.bfloat16).view(torch.int16).numpy()
        if a.ndim != 2 or (a.shape[0] < 1 and not allow_empty):
            raise ValueError(f"{vid}: expected [T>=1, E] features, got {a.shape}")
        if dim is None:
            dim = a.shape[1]
        elif a.shape[1] != dim:
            raise ValueError(f"{vid}: feature dim {a.shape[1]} != {dim}")
        ids.append(str(vid))
        offsets.append(offsets[-1] + a.shape[0])
        chunks.append(a)
    header = json.dumps({"dim": dim or 0, "dtype": dtype, "video_ids": ids, "offsets": offsets, "extra": extra or {}}).encode()
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
        self.dtype: str = meta.get("dtype", "float32")
        self.extra: dict = meta.get("extra", {})
        self.video_ids: List[str] = list(meta["video_ids"])
        self.offsets = np.asarray(meta["offsets"], np.int64)
        data_start = (16 + hlen + ALIGN - 1) // ALIGN * ALIGN
        rows = int(self.offsets[-1])
        np_dtype = np.float32 if self.dtype == "float32" else np.int16   # bf16 rows are kept as raw 16-bit words on the host
        self.data = np.memmap(path, np_dtype, "r", offset=data_start, shape=(rows, self.dim)) if rows else np.zeros((0, self.dim), np_dtype)
        self._index: Dict[str, int] = {v: i for i, v in enumerate(self.video_ids)}

    def __len__(self) -> int:
        return len(self.video_ids)

    def features(self, video_id: str) -> torch.Tensor:
        """The tensor ``torch.load(f"{video_id}.pt")`` would have returned (a copy)."""
        i = self._index[video_id]
        t = torch.from_numpy(np.array(self.data[self.offsets[i]:self.offsets[i + 1]]))
        return t if self.dtype == "float32" else t.view(torch.bfloat16).float()

    def sentence_seconds(self):
        """ASR stores: (starts, ends) int32 ``[sum S]`` — the subtitle sentences' seconds, packed like the rows."""
        return (torch.tensor(self.extra.get("starts", []), dtype=torch.int32), torch.tensor(self.extra.get("ends", []), dtype=torch.int32))

    def to_device(self, device, video_range: Optional[Tuple[int, int]] = None):
        """(features [rows, E] fp32 — or bf16 for a bf16 store —, offsets int64 [V+1]) on ``device`` for videos ``[lo, hi)``
        (default: all).  One host→device copy of the mapped rows through a pinned staging buffer."""
        lo, hi = video_range if video_range is not None else (0, len(self))
        r0, r1 = int(self.offsets[lo]), int(self.offsets[hi])
        tdt = torch.float32 if self.dtype == "float32" else torch.int16
        staging = torch.empty((r1 - r0, self.dim), dtype=tdt)
        if torch.cuda.is_available():
            staging = staging.pin_memory()
        if r1 > r0:
            np.copyto(staging.numpy(), self.data[r0:r1])
        feats = staging.to(device, non_blocking=True)
        if self.dtype != "float32":
            feats = feats.view(torch.bfloat16)
        offs = torch.from_numpy(self.offsets[lo:hi + 1] - r0).to(device, non_blocking=True)
        return feats, offs


@torch.no_grad()
def pooled_video_embeddings(feats: torch.Tensor, offsets: torch.Tensor, n_model_frames: int) -> torch.Tensor:
    """Packed ``[sum T, E]`` fp32 features + ``[V+1]`` int64 offsets (both CUDA) → ``[V, E]`` L2-normalised video embeddings:
    the loop body of inference_video_retrieval.py:306-326 for every video in one launch.  ``n_model_frames <= 0`` pools all frames."""
    if not feats.is_cuda or not offsets.is_cuda:
        raise RuntimeError("hirest_b200.feature_store has no CPU path: call FeatureStore.to_device first")
    if feats.dtype not in (torch.float32, torch.bfloat16) or offsets.dtype != torch.int64:
        raise TypeError("expected fp32 (or bf16) features and int64 offsets")
    feats, offsets = feats.contiguous(), offsets.contiguous()
    V = offsets.numel() - 1
    E = feats.shape[1]
    out = torch.empty((V, E), dtype=torch.float32, device=feats.device)
    lib = _lib.init(feats.device.index if feats.device.index is not None else torch.cuda.current_device())
    fn = lib.hb_subsample_pool_normalize if feats.dtype == torch.float32 else lib.hb_subsample_pool_normalize_bf16
    with torch.cuda.device(feats.device):
        _lib.check(fn(feats.data_ptr(), offsets.data_ptr(), V, int(n_model_frames), E, out.data_ptr(), _lib.stream_ptr(feats.device)),
                   "hb_subsample_pool_normalize")
    return out


def pack_asr_features(items: Iterable[Tuple[str, torch.Tensor, Sequence[Tuple[int, int]]]], path: str) -> None:
    """ASR store: ``(video_id, sentence_features[S, C], [(start_s, end_s)] * S)`` — the ``{video_id}.pt`` tensors of ``asr_feature_dir``
    and the seconds of the matching ``.srt`` blocks (hirest_dataset.py:99-109, 366-379)."""
    starts: List[int] = []
    ends: List[int] = []

    def rows():
        for vid, feats, subs in items:
            n = min(feats.shape[0], len(subs))   # the reference indexes asr_features[i] per subtitle i
            if len(subs) > feats.shape[0]:
                raise ValueError(f"{vid}: {len(subs)} subtitle blocks but {feats.shape[0]} feature rows")
            for a, b in subs:
                starts.append(int(a))
                ends.append(int(b))
            yield vid, feats[:n]

    extra = {"starts": starts, "ends": ends}
    pack_features(rows(), path, extra=extra, allow_empty=True)
