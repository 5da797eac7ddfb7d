"""Inference-side data feed of the joint model: annotation JSON -> per-task items -> collated batches, with the cached video /
ASR features coming from packed stores and resampled / warped on the GPU (SURVEY.md §8(f) N2 + N4).

Reference: ``hirest_dataset.py`` — ``MomentDataset.__init__`` builds one datum per (prompt, video[, step]) for the task
(``:124-316``), ``__getitem__`` loads ``{fname}.pt`` / ``{video_id}.pt`` with ``torch.load`` and resamples / warps on the CPU
(``:323-407``), ``collate_fn`` pads and stacks (``:409-531``).  Here the item construction and the collate are the same rules on the
host (they are a few integers per item), while the feature work — linspace subsample or repeat-pad to ``n_model_frames``
(``:333-356``), ASR sentence features warped onto the 1-fps time axis (``:370-381``) and resampled like the video (``:383-403``) —
runs on the GPU over the packed blobs of ``hirest_b200.feature_store`` (``hb_resample_rows``, ``hb_asr_warp``).  Reference quirks
kept: with a fixed ``n_model_frames`` the ASR sentences are warped onto the already resampled video length (``:372``), and
``round()`` (banker's rounding) of the annotated duration (``:145``).

Only the inference branches are implemented (moment retrieval, the test branch of moment segmentation, step captioning);
the training-only segmentation items (``:196-238``) are out of scope.
"""
from __future__ import annotations

import functools
import re
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib

TASKS = ("moment_retrieval", "moment_segmentation", "step_captioning")


@functools.lru_cache(maxsize=4096)
def _bins(video_duration: int, n_frames: int) -> np.ndarray:
    """np.linspace(0, video_duration - 1, n_frames), hirest_dataset.py:27 / :57 (read-only: shared between calls)."""
    b = np.linspace(0, video_duration - 1, n_frames)
    b.setflags(write=False)
    return b


def timestamp_to_frame_index(timestamp, video_duration, n_frames: int = 32) -> int:
    """hirest_dataset.py:12-40: index of the linspace bin that contains the timestamp (right-closed), clipped."""
    video_duration = int(video_duration)
    if n_frames < 0:
        n_frames = video_duration
    bins = _bins(video_duration, n_frames)
    return int(min(np.searchsorted(bins, timestamp, side="left"), n_frames - 1))   # == np.digitize(timestamp, bins, right=True)


def frame_index_to_timestamp(frame_index: int, video_duration, n_frames: int = 32) -> int:
    """hirest_dataset.py:42-68."""
    video_duration = int(video_duration)
    if n_frames < 0:
        n_frames = video_duration
    return int(_bins(video_duration, n_frames)[frame_index])


def build_items(annotations: Dict[str, Dict[str, dict]], task: str, n_model_frames: int = -1, end_to_end: bool = False,
                caption_tokenizer=None, max_words: int = 48) -> List[dict]:
    """The data list ``MomentDataset.__init__`` builds for an evaluation split (hirest_dataset.py:124-316): one item per relevant
    clipped video (moment retrieval / segmentation) or per annotated step (step captioning), in annotation order.
    ``annotations``: ``{prompt: {video_fname: {"relevant", "clip", "v_duration", "bounds", "steps": [{"heading",
    "absolute_bounds"}]}}}`` (``all_data_{split}.json``).  ``caption_tokenizer``: a ``hirest_b200.wordpiece.WordPieceTokenizer`` to
    attach the step-captioning targets (``clip4cap_get_text``); omitted, items carry only ``target_text_raw``."""
    if task not in TASKS:
        raise ValueError(f"unknown task {task!r}")
    items: List[dict] = []
    for prompt, video_anns in annotations.items():
        for fname, ann in video_anns.items():
            if not ann["relevant"] or not ann["clip"]:
                continue
            duration = round(ann["v_duration"])                        # :145 — rounded once, used everywhere below
            n = n_model_frames if n_model_frames > 0 else duration
            base = {"fname": fname, "prompt": prompt, "video_duration": duration, "n_model_frames": n_model_frames, "task": task}

            def t2f(t):
                return timestamp_to_frame_index(t, video_duration=duration, n_frames=n)

            def f2t(f):
                return frame_index_to_timestamp(f, video_duration=duration, n_frames=n)

            if task == "moment_retrieval":                                 # :153-183
                s, e = ann["bounds"][0], ann["bounds"][1]
                sf, ef = t2f(s), t2f(e)
                items.append(dict(base, moment_retrieval_start_target=sf, moment_retrieval_end_target=ef, original_bounds=[[s, e]],
                                  approximate_bounds=[[f2t(sf), f2t(ef)]], video_mask=torch.ones(n, dtype=torch.long),
                                  moment_mask=torch.ones(n, dtype=torch.long)))
            elif task == "moment_segmentation":                            # :185-188, 240-266 (evaluation branch)
                if not end_to_end and len(ann["steps"]) == 0:
                    continue
                bounds = sorted({b for step in ann["steps"] for b in step["absolute_bounds"]})
                s, e = ann["bounds"][0], ann["bounds"][1]
                sf, ef = t2f(s), t2f(e)
                mm = torch.zeros(n, dtype=torch.long)
                mm[sf:ef + 1] = 1
                items.append(dict(base, moment_bound_timestamps=[s, e], moment_bound_frames=[sf, ef], moment_mask=mm,
                                  video_mask=torch.ones(n, dtype=torch.long), all_bound_frames=[t2f(b) for b in bounds]))
            else:                                                          # step captioning, :268-311
                if not end_to_end and len(ann["steps"]) == 0:
                    continue
                if len(ann["steps"]) == 0:
                    # the reference indexes steps[0] here and raises IndexError (:276); an end-to-end video whose segmentation found
                    # no step simply has nothing to caption
                    continue
                for step in ann["steps"]:
                    s, e = step["absolute_bounds"]
                    text = step["heading"].strip()
                    sf, ef = t2f(s), t2f(e)
                    mm = torch.zeros(n, dtype=torch.long)
                    mm[sf:ef] = 1
                    mm[ef] = 1
                    it = dict(base, target_text_raw=text, moment_mask=mm, video_mask=torch.ones(n, dtype=torch.long))
                    if caption_tokenizer is not None:
                        it["target_text"] = caption_tokenizer.encode_caption(text, max_words)
                    items.append(it)
    return items


def collate(items: Sequence[dict], n_model_frames: int = -1, tokenize: Optional[Callable] = None, feature_alloc: Optional[Callable] = None) -> dict:
    """``collate_fn`` (hirest_dataset.py:409-531): stack (fixed frame count) or zero-pad every per-frame tensor to the longest video of
    the batch; lists for the rest; ``clip_text_ids`` from the items if present, else ``tokenize(prompts)`` (``clip.tokenize``, :528).
    ``feature_alloc(shape, dtype) -> uninitialised tensor`` supplies the storage of the two feature batches (e.g. pinned host memory,
    what the reference's ``DataLoader(pin_memory=True)`` produces, :624); their padding is then zero-filled explicitly."""
    out: dict = {}
    if "target_text" in items[0]:
        out["target_text"] = [d["target_text"] for d in items]
    if "target_text_raw" in items[0]:
        out["target_text_raw"] = [d["target_text_raw"] for d in items]
    if "vis_feats" in items[0]:
        # (the reference pads every item with torch.cat and stacks, :439-470; one zero-filled batch tensor + row copies gives the
        # same tensors without 4 allocations per item — collating 64 x 600-frame videos is 150 MB of copies either way)
        max_len = items[0]["vis_feats"].shape[0] if n_model_frames > 0 else max(d["vis_feats"].shape[0] for d in items)

        def pad_stack(key, dtype, alloc=None):
            first = items[0][key]
            shape = (len(items), max_len) + tuple(first.shape[1:])
            # (torch.zeros of a large host tensor maps untouched zero pages, so "zero everything, then copy the rows" writes each
            # byte once; storage from `alloc` is recycled memory and gets its padding — or, for short items, all of it — cleared)
            out_t = torch.zeros(shape, dtype=dtype, device=first.device) if alloc is None else alloc(shape, dtype)
            tails = alloc is not None and max_len * int(first[0].numel() if first.shape[0] else 1) >= (64 << 10)
            if alloc is not None and not tails:
                out_t.zero_()
            for i, d in enumerate(items):
                x = d[key]
                n = x.shape[0]
                if n_model_frames > 0 and n != max_len:
                    raise RuntimeError(f"stack expects each tensor to be equal size, but got {n} and {max_len} rows ({key})")
                out_t[i, :n] = x
                if tails and n < max_len:
                    out_t[i, n:].zero_()
            return out_t

        out["vis_feats"] = pad_stack("vis_feats", torch.float32, feature_alloc)
        out["vis_mask"] = pad_stack("video_mask", torch.int64)
        out["moment_mask"] = pad_stack("moment_mask", torch.int64)
        for k in ("moment_retrieval_start_target", "moment_retrieval_end_target"):
            if k in items[0]:
                out[k] = torch.LongTensor([d[k] for d in items])
        if "prev_boundary_mask" in items[0]:
            out["prev_boundary_mask"] = pad_stack("prev_boundary_mask", torch.int64)
        if "asr_feats" in items[0]:
            out["asr_feats"] = pad_stack("asr_feats", torch.float32, feature_alloc)
    if "moment_segmentation_target" in items[0]:
        out["moment_segmentation_target"] = torch.LongTensor([d["moment_segmentation_target"] for d in items])
    for k in ("moment_bound_timestamps", "moment_bound_frames"):
        if k in items[0]:
            out[k] = torch.LongTensor([d[k] for d in items])
    if "all_bound_frames" in items[0]:
        out["all_bound_frames"] = [d["all_bound_frames"] for d in items]
    out["video_duration"] = [d["video_duration"] for d in items]
    out["video_fnames"] = [d["fname"] for d in items]
    out["tasks"] = [d["task"] for d in items]
    out["prompts"] = [d["prompt"] for d in items]
    if "clip_text_ids" in items[0]:
        out["clip_text_ids"] = torch.stack([d["clip_text_ids"] for d in items])
    elif tokenize is not None:
        out["clip_text_ids"] = tokenize(out["prompts"])
    return out


# ---------------------------------------------------------------------------------------------------------------------
# subtitles
# ---------------------------------------------------------------------------------------------------------------------
_SRT_TIME = re.compile(r"(\d+):(\d+):(\d+)[,.](\d+)\s*-->\s*(\d+):(\d+):(\d+)[,.](\d+)")


def parse_srt_seconds(text: str) -> List[Tuple[int, int]]:
    """(start, end) of every subtitle block in whole seconds — what the reference reads off ``srt.parse`` as ``sub.start.seconds`` /
    ``sub.end.seconds`` (hirest_dataset.py:105-109, 376-379; ``timedelta.seconds`` drops the fractional part)."""
    out = []
    for m in _SRT_TIME.finditer(text):
        h0, m0, s0, _, h1, m1, s1, _ = (int(x) for x in m.groups())
        out.append(((h0 * 3600 + m0 * 60 + s0) % 86400, (h1 * 3600 + m1 * 60 + s1) % 86400))
    return out


# ---------------------------------------------------------------------------------------------------------------------
# GPU feature feed
# ---------------------------------------------------------------------------------------------------------------------
def _dev_index(t: torch.Tensor) -> int:
    return t.device.index if t.device.index is not None else torch.cuda.current_device()


@torch.no_grad()
def resample_rows(feats: torch.Tensor, offsets: torch.Tensor, n_model_frames: int) -> torch.Tensor:
    """Packed ``[sum T, C]`` fp32 rows + ``[V+1]`` int64 offsets (CUDA) -> ``[V, n_model_frames, C]``: linspace subsample when a video
    has more rows, repeat-pad when it has fewer (hirest_dataset.py:333-356 / 383-403), every video in one launch."""
    if not feats.is_cuda or not offsets.is_cuda:
        raise RuntimeError("hirest_b200.dataset.resample_rows has no CPU path")
    if feats.dtype != torch.float32 or offsets.dtype != torch.int64:
        raise TypeError("expected fp32 rows and int64 offsets")
    feats, offsets = feats.contiguous(), offsets.contiguous()
    V, C = offsets.numel() - 1, feats.shape[1]
    out = torch.empty((V, int(n_model_frames), C), dtype=torch.float32, device=feats.device)
    lib = _lib.init(_dev_index(feats))
    with torch.cuda.device(feats.device):
        _lib.check(lib.hb_resample_rows(feats.data_ptr(), offsets.data_ptr(), V, int(n_model_frames), C, out.data_ptr(),
                                        _lib.stream_ptr(feats.device)), "hb_resample_rows")
    return out


@torch.no_grad()
def warp_asr(asr: torch.Tensor, sub_offsets: torch.Tensor, starts: torch.Tensor, ends: torch.Tensor, frame_offsets: torch.Tensor) -> torch.Tensor:
    """Sentence-level ASR features -> the 1-fps time axis of each video (hirest_dataset.py:370-381).  ``asr`` fp32 ``[sum S, C]`` packed
    per video (``sub_offsets`` int64 ``[V+1]``), ``starts`` / ``ends`` int32 ``[sum S]`` seconds, ``frame_offsets`` int64 ``[V+1]`` = packed
    row offsets of the output (video lengths); all CUDA.  Returns fp32 ``[sum len, C]``."""
    if not (asr.is_cuda and sub_offsets.is_cuda and starts.is_cuda and ends.is_cuda and frame_offsets.is_cuda):
        raise RuntimeError("hirest_b200.dataset.warp_asr has no CPU path")
    dev = asr.device
    V = frame_offsets.numel() - 1
    lens = (frame_offsets[1:] - frame_offsets[:-1])
    rows = int(frame_offsets[-1])
    row_video = torch.repeat_interleave(torch.arange(V, dtype=torch.int32, device=dev), lens)
    out = torch.empty((rows, asr.shape[1]), dtype=torch.float32, device=dev)
    lib = _lib.init(_dev_index(asr))
    with torch.cuda.device(dev):
        _lib.check(lib.hb_asr_warp(asr.float().contiguous().data_ptr(), sub_offsets.contiguous().data_ptr(),
                                   starts.to(torch.int32).contiguous().data_ptr(), ends.to(torch.int32).contiguous().data_ptr(),
                                   frame_offsets.contiguous().data_ptr(), row_video.data_ptr(), rows, asr.shape[1], out.data_ptr(),
                                   _lib.stream_ptr(dev)), "hb_asr_warp")
    return out


class FeatureFeed:
    """Video + ASR features of a split, resident on the GPU as two packed blobs, served per batch in the layout ``collate_fn`` gives
    the model: ``vis_feats [B, T, 1024]`` / ``asr_feats [B, T, 384]`` already on the device (T = ``n_model_frames`` or the longest
    video of the batch, zero padded).  Replaces one ``torch.load`` per item per task plus the CPU loops of ``__getitem__``."""

    def __init__(self, video_store, asr_store, device):
        self.device = torch.device(device)
        self.vfeats, self.voffs = video_store.to_device(self.device)
        self.vindex = {v: i for i, v in enumerate(video_store.video_ids)}
        self.afeats, self.aoffs = asr_store.to_device(self.device)
        st, en = asr_store.sentence_seconds()
        self.starts, self.ends = st.to(self.device), en.to(self.device)
        self.aindex = {v: i for i, v in enumerate(asr_store.video_ids)}
        self._voffs_host = video_store.offsets

    @torch.no_grad()
    def batch(self, fnames: Sequence[str], n_model_frames: int = -1):
        dev = self.device
        vi = torch.tensor([self.vindex[f] for f in fnames], dtype=torch.int64)
        ai = torch.tensor([self.aindex[f.replace(".mp4", "")] for f in fnames], dtype=torch.int64)
        lens = torch.from_numpy(self._voffs_host[vi.numpy() + 1] - self._voffs_host[vi.numpy()])
        B = len(fnames)
        # gather this batch's rows into a packed [sum len, C] block (index_select of row ranges), then warp / resample / pad
        foffs = torch.zeros(B + 1, dtype=torch.int64)
        foffs[1:] = torch.cumsum(lens, 0)
        row_ids = torch.cat([torch.arange(int(self._voffs_host[v]), int(self._voffs_host[v + 1])) for v in vi.tolist()]).to(dev)
        vis = self.vfeats.index_select(0, row_ids)
        a0 = self.aoffs.cpu()
        s_lens = (a0[ai + 1] - a0[ai])
        soffs = torch.zeros(B + 1, dtype=torch.int64)
        soffs[1:] = torch.cumsum(s_lens, 0)
        s_ids = torch.cat([torch.arange(int(a0[a]), int(a0[a + 1])) for a in ai.tolist()] or [torch.zeros(0, dtype=torch.int64)]).to(dev)
        if n_model_frames > 0:
            # reference quirk (hirest_dataset.py:372): the sentences are warped onto the length of the ALREADY resampled video
            # features, i.e. onto n_model_frames "seconds" (later seconds are dropped), after which the ASR resampling is the identity
            aoffs = (torch.arange(B + 1, dtype=torch.int64) * n_model_frames).to(dev)
            asr = warp_asr(self.afeats.index_select(0, s_ids), soffs.to(dev), self.starts.index_select(0, s_ids),
                           self.ends.index_select(0, s_ids), aoffs)
            return resample_rows(vis, foffs.to(dev), n_model_frames), asr.view(B, n_model_frames, -1)
        asr = warp_asr(self.afeats.index_select(0, s_ids), soffs.to(dev), self.starts.index_select(0, s_ids),
                       self.ends.index_select(0, s_ids), foffs.to(dev))
        T = int(lens.max())
        pad_idx = (foffs[:-1, None] + torch.arange(T)[None, :])
        valid = torch.arange(T)[None, :] < lens[:, None]
        pad_idx = torch.where(valid, pad_idx, torch.zeros_like(pad_idx)).to(dev)
        valid = valid.to(dev)[..., None]
        return (vis.index_select(0, pad_idx.reshape(-1)).view(B, T, -1) * valid, asr.index_select(0, pad_idx.reshape(-1)).view(B, T, -1) * valid)
