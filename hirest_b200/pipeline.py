"""In-memory end-to-end chain: moment retrieval → moment segmentation → step captioning.

Reference: ``run.py:383-490`` (``--end_to_end``) runs the three tasks one after the other and hands results from one to the
next THROUGH DISK: it dumps each task's predictions to JSON, rewrites ``all_data_test.json`` with the predicted bounds /
steps, and rebuilds dataset + loader for the next task (``hirest_dataset.py:150-300`` turns the timestamps back into frame
masks, ``:409-531`` collates).  Here the same hand-off happens in memory: predictions stay Python lists for the few integers
that cross task boundaries, features stay tensors, and each task is one ``MomentModel.test_step`` on the GPU.

The glue below restates, with citations, exactly what the reference computes between the model calls:

* timestamp ↔ frame conversions (``hirest_dataset.py:12-68``),
* the per-task item construction and masks (``hirest_dataset.py:153-300``),
* ``collate_fn`` padding (``hirest_dataset.py:409-531``),
* the result dictionaries of ``evaluate`` (``run.py:704-830``) and the JSON updates of ``run.py:399-474``.

Output: the structure the reference writes to ``final_end_to_end_results.json``.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import numpy as np
import torch


from .dataset import build_items, collate as _collate, frame_index_to_timestamp, timestamp_to_frame_index  # noqa: F401  (re-exported)


def collate(items: Sequence[dict], n_model_frames: int = -1, feature_alloc=None) -> dict:
    """``collate_fn`` (hirest_dataset.py:409-531) for inference items; see ``hirest_b200.dataset.collate``."""
    return _collate(items, n_model_frames, feature_alloc=feature_alloc)


def shard_indices(n_items: int, rank: int, world: int) -> List[int]:
    """Item indices of `rank`: torch's DistributedSampler(dataset, shuffle=False) as hirest_dataset.py:604-606 builds it
    (round-robin rank::world over the index list padded to a multiple of world by wrapping around)."""
    if world <= 1:
        return list(range(n_items))
    if n_items == 0:
        return []
    per = (n_items + world - 1) // world
    total = per * world
    idx = list(range(n_items))
    while len(idx) < total:
        idx += idx[:total - len(idx)]
    return idx[rank:total:world]


def all_gather_objects(obj, group=None) -> list:
    """dist_utils.all_gather (dist_utils.py:145-179): every rank receives the list of every rank's picklable object, in rank order."""
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return [obj]
    out = [None] * dist.get_world_size(group)
    dist.all_gather_object(out, obj, group=group)
    return out


def _run_sharded(items: List[dict], run_batch, batch_size: int, rank: int, world: int, gather, prepare=None, prefetch: bool = True) -> List:
    """Run `run_batch(chunk) -> list of per-item results` over this rank's items, gather every rank's (index, result) pairs and
    return the per-item results in ITEM order (padding duplicates dropped).  (The reference extends the gathered lists in rank
    order, run.py:645-662, which is equivalent for its dictionaries keyed by video and wrong for the caption lists; item order
    is what the single-process run produces.)

    With ``prepare`` the work of a chunk is split into ``prepare(chunk) -> batch`` (host collate + host-to-device copies) and
    ``run_batch(batch)`` (the model call): one worker thread prepares chunk i + 1 while the model runs chunk i — what the
    reference's ``DataLoader(num_workers=..., pin_memory=True)`` does with processes (hirest_dataset.py:620-630).  Results do not depend on it."""
    mine = shard_indices(len(items), rank, world)
    chunks = [mine[c0:c0 + batch_size] for c0 in range(0, len(mine), batch_size)]
    local = []
    if prepare is None:
        for ids in chunks:
            local += list(zip(ids, run_batch([items[i] for i in ids])))
    elif not prefetch:
        for ids in chunks:
            local += list(zip(ids, run_batch(prepare([items[i] for i in ids]))))
    elif chunks:
        from concurrent.futures import ThreadPoolExecutor

        with ThreadPoolExecutor(max_workers=1) as pool:
            fut = pool.submit(prepare, [items[i] for i in chunks[0]])
            for k, ids in enumerate(chunks):
                batch = fut.result()
                if k + 1 < len(chunks):
                    fut = pool.submit(prepare, [items[i] for i in chunks[k + 1]])
                local += list(zip(ids, run_batch(batch)))
    merged = {}
    for part in gather(local):
        for i, r in part:
            merged.setdefault(i, r)
    return [merged[i] for i in range(len(items))]


def run_end_to_end(model, videos: Sequence[dict], batch_size: int = 64, num_beams: int = 5, n_model_frames: int = -1,
                   tokenize=None, rank: int = 0, world: int = 1, group=None, gather=None,
                   caption_batch_size: Optional[int] = None, prefetch: bool = True) -> Dict:
    """Chain the three tasks over ``videos`` (dicts with ``prompt``, ``fname``, ``video_duration``, ``vis_feats [T,1024]``,
    ``asr_feats [T,384]``, ``clip_text_ids [77]``; order = dataset order).  Videos without ``clip_text_ids`` get them from
    ``tokenize(prompt) -> LongTensor[1, 77]`` (e.g. ``hirest_b200.tokenizer.tokenize``), as ``collate_fn`` does with
    ``clip.tokenize`` (hirest_dataset.py:528).  Returns
    ``{"final": {prompt: {fname: {"bounds", "steps": [{"index", "heading", "absolute_bounds"}]}}},
    "moment_retrieval": ..., "moment_segmentation": ..., "step_captioning": ...}`` with the per-task dictionaries the
    reference dumps next to the final file.

    Multi-GPU: with ``world > 1`` every task's items are sharded over the ranks with DistributedSampler
    semantics (``shard_indices``), each rank runs its share, and the per-item results are gathered as Python objects
    (``all_gather_objects`` = dist_utils.all_gather; ``gather`` overrides it, e.g. to simulate ranks in one process), so every rank
    returns the full dictionaries.  There is no tensor exchange on this path.

    ``caption_batch_size`` (default: ``8 * batch_size`` — a video yields about that many steps) batches the step-captioning
    items separately: a step item is at most 20 trimmed frames, and the 48 decode steps of a beam search are latency-bound, so
    several hundred steps per search cost little more than 64 (256-video job, captioning stage: 0.53 s in searches of 256 items,
    0.32 s with 512, 0.28 s with 1024).  A caption does not depend on what else is in its batch (trimmed items carry no padding; every decoder kernel
    computes a row from that row's inputs only), so this is a throughput knob, not a semantic one.

    ``prefetch`` (default on): one worker thread collates the next batch and copies its features to the GPU on a side stream while
    the model runs the current one (``_run_sharded``); off, every batch is prepared inline.  Same results either way."""
    nmf = n_model_frames
    if gather is None:
        gather = (lambda obj: all_gather_objects(obj, group)) if world > 1 else (lambda obj: [obj])
    # (hirest_dataset.py:145 rounds the annotated duration once, inside build_items; everything downstream uses the rounded value)
    if any("clip_text_ids" not in v for v in videos):
        if tokenize is None:
            raise ValueError("videos without 'clip_text_ids' need a tokenize callable (hirest_b200.tokenizer.tokenize)")
        cache: Dict[str, torch.Tensor] = {}
        filled = []
        for v in videos:
            if "clip_text_ids" not in v:
                if v["prompt"] not in cache:
                    cache[v["prompt"]] = tokenize(v["prompt"])[0]
                v = dict(v, clip_text_ids=cache[v["prompt"]])
            filled.append(v)
        videos = filled
    by_name = {v["fname"]: v for v in videos}

    def annotations(bounds_of, steps_of):
        """The test JSON the reference (re)writes between tasks (run.py:399-452), grouped by prompt in first-appearance order."""
        ann: Dict[str, Dict[str, dict]] = {}
        for v in videos:
            ann.setdefault(v["prompt"], {})[v["fname"]] = {"relevant": True, "clip": True, "v_duration": v["video_duration"],
                                                         "bounds": bounds_of(v), "steps": steps_of(v)}
        return ann

    def with_features(items, slice_to_moment=False):
        out = []
        for it in items:
            v = by_name[it["fname"]]
            n = it["video_mask"].numel()
            if v["vis_feats"].shape[0] != n:
                raise ValueError(f"{v['fname']}: {v['vis_feats'].shape[0]} feature rows for {n} frames")
            it = dict(it, clip_text_ids=v["clip_text_ids"])
            if slice_to_moment:
                # Step captioning only ever reads the frames inside the step (trim_feats, modeling.py:529-554, keeps the rows with
                # moment_mask == 1, in order), so the item carries just those rows with an all-ones mask instead of the whole video
                # padded to the longest one in the batch: same trimmed tensor, ~30x fewer bytes collated and copied to the GPU.
                rows = it["moment_mask"].nonzero(as_tuple=True)[0]
                k = rows.numel()
                if k > 0 and int(rows[-1]) - int(rows[0]) + 1 == k:   # a step is one run of frames (dataset :287-289): views, no copy
                    r0 = int(rows[0])
                    vis, asr = v["vis_feats"][r0:r0 + k], v["asr_feats"][r0:r0 + k]
                else:
                    vis, asr = v["vis_feats"][rows], v["asr_feats"][rows]
                it.update(vis_feats=vis, asr_feats=asr, video_mask=torch.ones(k, dtype=torch.long), moment_mask=torch.ones(k, dtype=torch.long))
            else:
                it.update(vis_feats=v["vis_feats"], asr_feats=v["asr_feats"])
            out.append(it)
        return out

    # ---- 1. moment retrieval (dataset :153-183, evaluate :704-744) -------------------------------------------------
    # (items carry masks and bounds; the features are attached per batch inside the prepare step, i.e. for this rank's items only
    # and, with prefetch, by the worker thread)
    items = build_items(annotations(lambda v: [0, 0], lambda v: []), "moment_retrieval", nmf, end_to_end=True)
    mr: Dict[str, Dict[str, dict]] = {}
    # Retrieval and segmentation see the same videos in the same batches (one item per video, same order, same shards): the padded
    # feature tensors of a batch are collated and copied to the GPU once and reused by the segmentation pass (the reference
    # re-reads every feature file and re-collates, run.py:419-435); only the masks / bounds of the items differ.
    feat_cache: Dict[tuple, dict] = {}
    cache_budget = [8 << 30]   # bytes of device memory the cache may hold; beyond it batches are collated again
    try:
        dev = next(model.parameters()).device
    except (AttributeError, StopIteration, TypeError):
        dev = None   # not a torch module (tests drive the glue with stand-ins): no device cache

    on_gpu = dev is not None and dev.type == "cuda"
    side = torch.cuda.Stream(dev) if on_gpu else None   # the worker thread's copies run beside the model's kernels

    def pinned(shape, dtype):
        """Feature batches are collated straight into pinned host memory (torch's caching host allocator recycles the blocks and
        holds a block back until the copy that reads it has finished): pageable batches went to the GPU at ~1.4 GB/s."""
        return torch.empty(shape, dtype=dtype, pin_memory=True)

    feature_alloc = pinned if on_gpu else None

    def to_device(*tensors):
        """Host-to-device copies on the side stream; the model's stream waits for them in `ready`."""
        with torch.cuda.stream(side):
            out = [t.to(dev, non_blocking=True) for t in tensors]
            ev = torch.cuda.Event()
            ev.record(side)
        return out, ev

    def ready(b):
        ev = b.pop("_copied", None)
        if ev is not None:
            cur = torch.cuda.current_stream(dev)
            cur.wait_event(ev)
            for k in ("vis_feats", "asr_feats"):
                b[k].record_stream(cur)
        return b

    def prepare_video_batch(chunk):
        key = tuple(it["fname"] for it in chunk)
        hit = feat_cache.get(key)
        chunk = with_features(chunk)
        if hit is None:
            b = collate(chunk, nmf, feature_alloc)
            nbytes = (b["vis_feats"].numel() + b["asr_feats"].numel()) * 4
            if on_gpu and nbytes <= cache_budget[0]:
                cache_budget[0] -= nbytes
                (b["vis_feats"], b["asr_feats"]), b["_copied"] = to_device(b["vis_feats"], b["asr_feats"])
                feat_cache[key] = {"vis_feats": b["vis_feats"], "asr_feats": b["asr_feats"]}
        else:   # masks and bounds from the items, features from the cache (placeholders keep collate's padding logic in one place)
            light = [dict(it, vis_feats=it["vis_feats"].new_empty((it["vis_feats"].shape[0], 0)),
                          asr_feats=it["asr_feats"].new_empty((it["asr_feats"].shape[0], 0))) for it in chunk]
            b = collate(light, nmf)
            b.update(hit)
        return b

    def run_video_batch(b):
        return model.test_step(ready(b))["prediction"]

    preds = _run_sharded(items, run_video_batch, batch_size, rank, world, gather, prepare=prepare_video_batch, prefetch=prefetch)
    for it, (s, e) in zip(items, preds):
        d = it["video_duration"]
        mr.setdefault(it["prompt"], {})[it["fname"]] = {
            "bounds": [frame_index_to_timestamp(s, d, n_frames=nmf), frame_index_to_timestamp(e, d, n_frames=nmf)],
            "video_duration": d}
    # the rewritten test file (run.py:399-417): predicted bounds + five placeholder steps
    state: Dict[str, Dict[str, dict]] = {}
    for v in videos:
        state.setdefault(v["prompt"], {})[v["fname"]] = {
            "bounds": mr[v["prompt"]][v["fname"]]["bounds"],
            "steps": [{"index": i, "heading": "", "absolute_bounds": [i, i + 1]} for i in range(5)]}
    # ---- 2. moment segmentation (dataset :239-266 test branch, evaluate :746-782) ----------------------------------
    items = build_items(annotations(lambda v: state[v["prompt"]][v["fname"]]["bounds"],
                                    lambda v: state[v["prompt"]][v["fname"]]["steps"]), "moment_segmentation", nmf, end_to_end=True)
    ms: Dict[str, dict] = {}
    preds = _run_sharded(items, run_video_batch, batch_size, rank, world, gather, prepare=prepare_video_batch, prefetch=prefetch)
    feat_cache.clear()
    for it, raw in zip(items, preds):
        d = it["video_duration"]
        bounds = [[frame_index_to_timestamp(raw[j], d, n_frames=nmf), frame_index_to_timestamp(raw[j + 1], d, n_frames=nmf)]
                  for j in range(len(raw) - 1)]
        ms[it["fname"]] = {"bounds": bounds, "video_duration": d, "pred_bounds": raw}   # keyed by video only (run.py:757)
    for prompt in state:                                                                     # run.py:437-452
        for fname in state[prompt]:
            state[prompt][fname]["steps"] = [{"index": i, "heading": "", "absolute_bounds": b}
                                             for i, b in enumerate(ms[fname]["bounds"])] if fname in ms else []
    # ---- 3. step captioning (dataset :268-312, evaluate :787-830) --------------------------------------------------
    # (a video whose segmentation produced no step contributes nothing; the reference indexes steps[0] and raises)
    items = build_items(annotations(lambda v: state[v["prompt"]][v["fname"]]["bounds"],
                                    lambda v: state[v["prompt"]][v["fname"]]["steps"]), "step_captioning", nmf, end_to_end=True)
    sc: Dict[str, dict] = {}
    def prepare_step_batch(chunk):
        b = collate(with_features(chunk, slice_to_moment=True), -1, feature_alloc)   # ragged items: pad path
        if on_gpu:
            (b["vis_feats"], b["asr_feats"]), b["_copied"] = to_device(b["vis_feats"], b["asr_feats"])
        return b

    preds = _run_sharded(items, lambda b: model.test_step(ready(b), num_beams=num_beams)["prediction"],
                         caption_batch_size or 8 * batch_size, rank, world, gather, prepare=prepare_step_batch, prefetch=prefetch)
    for it, sent in zip(items, preds):
        e = sc.setdefault(it["fname"], {"captions": []})
        e["captions"].append({"sentence": sent})
        e["video_duration"] = it["video_duration"]
    for prompt in state:                                                                     # run.py:466-472
        for fname in state[prompt]:
            if fname in sc:
                for i, sent in enumerate(sc[fname]["captions"]):
                    if i < len(state[prompt][fname]["steps"]):
                        state[prompt][fname]["steps"][i]["heading"] = sent["sentence"]
    return {"final": state, "moment_retrieval": mr, "moment_segmentation": ms, "step_captioning": sc}
