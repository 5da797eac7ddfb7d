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


def all_gather_embeddings(local_hat: torch.Tensor, group=None) -> torch.Tensor:
    """The one collective of the path: all-gather of L2-normalised video embeddings (equal shard sizes)."""
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local_hat
    world = dist.get_world_size(group)
    out = torch.empty((world * local_hat.shape[0], local_hat.shape[1]), dtype=local_hat.dtype, device=local_hat.device)
    dist.all_gather_into_tensor(out, local_hat.contiguous(), group=group)
    return out


def rank_videos(scores_row: Sequence[float], video_names: Sequence[str]) -> List[int]:
    """Indices into video_names, best first, with the reference's tie-break: sort (score, name) ascending and reverse
    (evaluate.py:58-60)."""
    s = np.asarray(scores_row, dtype=np.float64)
    name_rank = np.argsort(np.argsort(np.asarray(video_names, dtype=object), kind="stable"), kind="stable")
    order = np.lexsort((name_rank, s))  # ascending by score, ties by name
    return order[::-1].tolist()


def topk(scores: torch.Tensor, video_names: Optional[Sequence[str]], k: int) -> List[List[int]]:
    """Top-k video indices per query under the reference ranking rule."""
    sc = scores.detach().float().cpu().numpy()
    names = list(video_names) if video_names is not None else [f"{i:09d}" for i in range(sc.shape[1])]
    return [rank_videos(sc[i], names)[:k] for i in range(sc.shape[0])]


@torch.no_grad()
def encode_and_score(model, frames: torch.Tensor, n_frames: int, text_hat: torch.Tensor, exact: bool = True, group=None):
    """One retrieval step: encode this rank's frames, mean-pool per video, L2-normalise, all-gather, score.
    frames: [V_local * n_frames, 3, S, S]; text_hat: [Q, E] normalised queries. Returns (scores [Q, V_total], v_hat)."""
    emb = model.encode_image(frames)
    v_hat = pool_normalize(emb, n_frames)
    v_all = all_gather_embeddings(v_hat, group)
    return similarity(text_hat, v_all, exact=exact), v_all
