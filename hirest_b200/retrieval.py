"""Video-retrieval scoring on the GPU — the math of inference_video_retrieval.py:203-334 without the file I/O.

    text:   t_hat = t / ||t||                                   (:210-212)
    video:  v     = mean over the F frames of a video           (:283 raw frames / :323 cached features)
            v_hat = v / ||v||                                   (:285 / :326)
    score:  S     = T_hat @ V_hat.T   (no logit scale)          (:334)
    rank:   descending by the tuple (score, video_name)         (evaluate.py:58-60)

Multi-GPU: whole videos are sharded over ranks (contiguous blocks, so the mean-pool stays local and the global
video order is preserved); ONE all-gather of the L2-normalised [V/R, E] embeddings over NCCL precedes the
similarity GEMM.  All device math goes through libhirest_b200.so; nothing here falls back to torch ops.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import numpy as np
import torch

from . import _lib


def _stream(dev):
    return _lib.stream_ptr(dev)


def pool_normalize(frame_embeds: torch.Tensor, n_frames: int) -> torch.Tensor:
    """[V*F, E] or [V, F, E] fp32 frame embeddings -> [V, E] L2-normalised video embeddings."""
    E = frame_embeds.shape[-1]
    x = frame_embeds.float().contiguous().view(-1, n_frames, E)
    V = x.shape[0]
    out = torch.empty((V, E), dtype=torch.float32, device=x.device)
    lib = _lib.init(x.device.index or 0)
    with torch.cuda.device(x.device):
        _lib.check(lib.hb_pool_normalize(x.data_ptr(), V, n_frames, E, out.data_ptr(), _stream(x.device)), "hb_pool_normalize")
    return out


def normalize(embeds: torch.Tensor) -> torch.Tensor:
    """Row-wise L2 normalisation (text side, inference_video_retrieval.py:210-212)."""
    return pool_normalize(embeds, 1)


def similarity(text_hat: torch.Tensor, video_hat: torch.Tensor, exact: bool = True) -> torch.Tensor:
    """S = T_hat @ V_hat.T as one tcgen05 GEMM.  exact=True uses 3-way split-bf16 operands (K = 6E) so the result is fp32-accurate
    and the ranking matches the reference's fp32 CPU matmul; exact=False is the plain single bf16 GEMM."""
    Q, E = text_hat.shape
    V = video_hat.shape[0]
    dev = text_hat.device
    Vp = (V + 15) // 16 * 16
    vh = video_hat.float().contiguous()
    if Vp != V:
        pad = torch.zeros((Vp - V, E), dtype=torch.float32, device=dev)
        vh = torch.cat((vh, pad), dim=0)
    scores = torch.empty((Q, Vp), dtype=torch.float32, device=dev)
    lib = _lib.init(dev.index or 0)
    with torch.cuda.device(dev):
        _lib.check(lib.hb_similarity(text_hat.float().contiguous().data_ptr(), Q, vh.data_ptr(), Vp, E, scores.data_ptr(), Vp,
                                     1 if exact else 0, _stream(dev)), "hb_similarity")
    return scores[:, :V]


def shard_range(n_items: int, rank: int, world: int):
    """Contiguous block of whole videos owned by `rank`."""
    per = (n_items + world - 1) // world
    lo = min(n_items, rank * per)
    return lo, min(n_items, lo + per)


class PendingGather:
    """Handle of an all-gather in flight (see all_gather_embeddings(..., async_op=True)); ``wait()`` makes the current stream wait
    for it and returns the gathered [n_total, E] embeddings in global video order."""

    def __init__(self, work, out, sizes, per, out_dtype=None):
        self._work, self._out, self._sizes, self._per, self._out_dtype = work, out, sizes, per, out_dtype

    def wait(self) -> torch.Tensor:
        if self._work is not None:
            self._work.wait()
            self._work = None
        if self._out_dtype is not None and self._out.dtype != self._out_dtype:
            self._out = self._out.to(self._out_dtype)   # wire dtype -> the caller's dtype, once
        out, sizes, per = self._out, self._sizes, self._per
        total = sum(sizes)
        if all(n == per for n in sizes[:-1]) or total == 0:
            # full blocks followed by at most one short block, then empty ones: the valid rows are a prefix
            full = 0
            for n in sizes:
                if n != per:
                    break
                full += 1
            if sum(sizes[full + 1:]) == 0:
                return out[:total]
        return torch.cat([out[r * per:r * per + n] for r, n in enumerate(sizes) if n > 0], dim=0)


def all_gather_embeddings(local_hat: torch.Tensor, group=None, n_total: Optional[int] = None, async_op: bool = False,
                          wire_dtype: Optional[torch.dtype] = None):
    """The one collective of the path: all-gather of the L2-normalised video embeddings over NCCL (gloo in the CPU tests).

    Shards may be UNEQUAL (4282 validation videos over 8 ranks = 7 x 536 + 530, or trailing empty shards): every rank pads its
    block to the common size, one ``all_gather_into_tensor`` moves the padded blocks, and the padding is dropped afterwards.
    ``n_total`` = the global number of videos when the shards are the contiguous blocks of ``shard_range`` (no extra
    communication); without it the per-rank sizes are exchanged first (one tiny all-gather + a host sync).
    ``async_op=True`` returns a PendingGather: the collective runs on the communicator's stream while the caller keeps
    enqueueing compute (bench.py overlaps the gather of step i with the encoder of step i+1).
    ``wire_dtype=torch.bfloat16`` sends the blocks as bf16 (SURVEY.md §8(e)'s wording: half the bytes of a transfer that is
    already < 0.01 % of a step) and returns them in ``local_hat``'s dtype; EVERY block, the local one included, then carries
    bf16-rounded values, so all ranks still score identical embeddings — but not the fp32 ones a single GPU would, which is
    why the default keeps fp32 on the wire and bench.py's sharded-equals-single-GPU check stays bit-exact."""
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        if wire_dtype is not None:
            local_hat = local_hat.to(wire_dtype).to(local_hat.dtype)   # same values a multi-rank run would score
        return PendingGather(None, local_hat, [local_hat.shape[0]], local_hat.shape[0]) if async_op else local_hat
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    n_local, E = local_hat.shape
    if n_total is not None:
        sizes = [shard_range(n_total, r, world)[1] - shard_range(n_total, r, world)[0] for r in range(world)]
        if sizes[rank] != n_local:
            raise ValueError(f"rank {rank} holds {n_local} embeddings, shard_range({n_total}, {rank}, {world}) says {sizes[rank]}")
    else:
        cnt = torch.tensor([n_local], dtype=torch.int64, device=local_hat.device)
        all_cnt = torch.empty((world,), dtype=torch.int64, device=local_hat.device)
        dist.all_gather_into_tensor(all_cnt, cnt, group=group)
        sizes = [int(x) for x in all_cnt.tolist()]
    per = max(sizes)
    wire = wire_dtype or local_hat.dtype
    block = local_hat.contiguous().to(wire)
    if n_local != per:
        block = torch.zeros((per, E), dtype=wire, device=local_hat.device)
        block[:n_local].copy_(local_hat)
    out = torch.empty((world * per, E), dtype=wire, device=local_hat.device)
    work = dist.all_gather_into_tensor(out, block, group=group, async_op=async_op)
    pend = PendingGather(work if async_op else None, out, sizes, per, local_hat.dtype)
    return pend if async_op else pend.wait()


def rank_videos(scores_row: Sequence[float], video_names: Sequence[str]) -> List[int]:
    """Indices into video_names, best first, with the reference's tie-break: sort (score, name) ascending and reverse
    (evaluate.py:58-60)."""
    s = np.asarray(scores_row, dtype=np.float64)
    name_rank = np.argsort(np.argsort(np.asarray(video_names, dtype=object), kind="stable"), kind="stable")
    order = np.lexsort((name_rank, s))  # ascending by score, ties by name
    return order[::-1].tolist()


def recall_at_k(scores, video_names: Sequence[str], gt_videos: Sequence[Sequence[str]], ks: Sequence[int] = (1, 5, 10, 50)) -> dict:
    """``evaluate_video_retrieval`` (evaluate.py:33-81) for the "all" category: per prompt, rank the videos by (score, name) as
    ``rank_videos`` does and count a hit at k if ANY of the top-k videos is one of the prompt's ground-truth videos.
    ``scores`` [Q, V] (tensor / array), ``gt_videos[q]`` = names of prompt q's ground-truth videos.  Returns ``{"R@k": percent}``
    plus ``total_prompt_count`` like the reference's result dictionary."""
    sc = scores.detach().float().cpu().numpy() if torch.is_tensor(scores) else np.asarray(scores, dtype=np.float64)
    names = list(video_names)
    hits = {k: 0 for k in ks}
    for q in range(sc.shape[0]):
        order = rank_videos(sc[q], names)
        gt = set(gt_videos[q])
        for k in ks:
            if any(names[j] in gt for j in order[:k]):
                hits[k] += 1
    total = sc.shape[0]
    out = {"total_prompt_count": total}
    out.update({f"R@{k}": (hits[k] / total) * 100 for k in ks} if total else {})
    return out


def topk(scores: torch.Tensor, video_names: Optional[Sequence[str]], k: int) -> List[List[int]]:
    """Top-k video indices per query under the reference ranking rule."""
    sc = scores.detach().float().cpu().numpy()
    names = list(video_names) if video_names is not None else [f"{i:09d}" for i in range(sc.shape[1])]
    return [rank_videos(sc[i], names)[:k] for i in range(sc.shape[0])]


@torch.no_grad()
def encode_and_score(model, frames: torch.Tensor, n_frames: int, text_hat: torch.Tensor, exact: bool = True, group=None,
                     n_total: Optional[int] = None):
    """One retrieval step: encode this rank's frames, mean-pool per video, L2-normalise, all-gather, score.
    frames: [V_local * n_frames, 3, S, S]; text_hat: [Q, E] normalised queries. Returns (scores [Q, V_total], v_hat)."""
    v_all = all_gather_embeddings(encode_videos(model, frames, n_frames), group, n_total=n_total)
    return similarity(text_hat, v_all, exact=exact), v_all


@torch.no_grad()
def encode_videos(model, frames: torch.Tensor, n_frames: int) -> torch.Tensor:
    """This rank's frames [V_local * n_frames, 3, S, S] -> L2-normalised video embeddings [V_local, E]
    (inference_video_retrieval.py:270-285: encode_image, view(B, F, E), mean over F, / norm)."""
    return pool_normalize(model.encode_image(frames), n_frames)


@torch.no_grad()
def retrieve(model, frames: torch.Tensor, n_frames: int, text_hat: torch.Tensor, n_total: Optional[int] = None, group=None,
             chunk_videos: int = 32, exact: bool = True, ks: Sequence[int] = (1, 5, 10, 50)):
    """BASELINE configs[2] as ONE job (inference_video_retrieval.py:257-334 without the file I/O): this rank encodes its contiguous
    block of whole videos chunk by chunk, ONE all-gather of the [V/R, E] embeddings at the end, then the similarity GEMM and the
    top-k lists.  frames: uint8 or fp32 [V_local * n_frames, 3, S, S] (host pinned or device; host chunks are copied on the
    current stream).  Returns (scores [Q, V_total], topk {k: int64 [Q, k]}, v_all)."""
    dev = text_hat.device
    V_local = frames.shape[0] // n_frames
    E = text_hat.shape[1]
    v_hat = torch.empty((V_local, E), dtype=torch.float32, device=dev)
    for v0 in range(0, V_local, chunk_videos):
        v1 = min(V_local, v0 + chunk_videos)
        chunk = frames[v0 * n_frames:v1 * n_frames]
        if chunk.device != dev:
            chunk = chunk.to(dev, non_blocking=True)
        v_hat[v0:v1] = encode_videos(model, chunk, n_frames)
    v_all = all_gather_embeddings(v_hat, group, n_total=n_total)
    scores = similarity(text_hat, v_all, exact=exact)
    kmax = min(max(ks), scores.shape[1])
    order = torch.topk(scores, kmax, dim=1, largest=True, sorted=True).indices   # ranking only; ties: see rank_videos / topk()
    return scores, {k: order[:, :min(k, kmax)] for k in ks}, v_all
