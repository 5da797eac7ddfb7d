"""world_size-2 gloo test of the multi-GPU host logic: contiguous video shards + one all-gather reproduce the
single-process score matrix (the GPU path uses the same code with the nccl backend)."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from hirest_b200 import retrieval
from oracle import eva_oracle


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = torch.Generator().manual_seed(0)
    V, F, E, Q = 8, 4, 32, 5
    frame_emb = torch.randn(V * F, E, generator=g)
    text = torch.randn(Q, E, generator=g)
    lo, hi = retrieval.shard_range(V, rank, world)
    local = eva_oracle.pool_normalize_video(frame_emb[lo * F:hi * F], F)  # stands in for the GPU pool kernel
    gathered = retrieval.all_gather_embeddings(local)
    scores = eva_oracle.similarity(eva_oracle.normalize_text(text), gathered)
    torch.save(scores, os.path.join(out_dir, f"scores_{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_scores_equal_single_process(tmp_path):
    world = 2
    port = 29500 + os.getpid() % 1000
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    g = torch.Generator().manual_seed(0)
    frame_emb = torch.randn(8 * 4, 32, generator=g)
    text = torch.randn(5, 32, generator=g)
    ref = eva_oracle.similarity(eva_oracle.normalize_text(text), eva_oracle.pool_normalize_video(frame_emb, 4))
    for r in range(world):
        got = torch.load(os.path.join(str(tmp_path), f"scores_{r}.pt"))
        assert torch.equal(got, ref)


UNEQUAL_CASES = [(V, wt, ua) for V in (7, 1, 4) for wt in (True, False) for ua in (False, True)]


def _worker_unequal(rank, world, port, out_dir, Vs):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    for n, (V, with_total, use_async) in enumerate(Vs):
        full = torch.randn(V, 16, generator=torch.Generator().manual_seed(1))
        lo, hi = retrieval.shard_range(V, rank, world)
        got = retrieval.all_gather_embeddings(full[lo:hi].clone(), n_total=V if with_total else None, async_op=use_async)
        if use_async:
            got = got.wait()
        torch.save(got.clone(), os.path.join(out_dir, f"g_{n}_{rank}.pt"))
    # bf16 on the wire (SURVEY.md section 8(e) wording): every rank gets the bf16-rounded embeddings back in fp32
    full = torch.randn(7, 16, generator=torch.Generator().manual_seed(1))
    lo, hi = retrieval.shard_range(7, rank, world)
    got = retrieval.all_gather_embeddings(full[lo:hi].clone(), n_total=7, wire_dtype=torch.bfloat16)
    torch.save(got.clone(), os.path.join(out_dir, f"gb_{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


def _check_unequal(tmp_path, world, cases):
    for n, (V, with_total, use_async) in enumerate(cases):
        full = torch.randn(V, 16, generator=torch.Generator().manual_seed(1))
        for r in range(world):
            assert torch.equal(torch.load(os.path.join(str(tmp_path), f"g_{n}_{r}.pt")), full), (V, with_total, use_async, r)
    full = torch.randn(7, 16, generator=torch.Generator().manual_seed(1))
    for r in range(world):
        got = torch.load(os.path.join(str(tmp_path), f"gb_{r}.pt"))
        assert got.dtype == torch.float32 and torch.equal(got, full.bfloat16().float()), r


def test_all_gather_with_unequal_and_empty_shards(tmp_path):
    """V % world != 0 (4282 videos on 8 ranks = 7 x 536 + 530) and trailing empty shards (V = 1 on 2 ranks)."""
    mp.spawn(_worker_unequal, args=(2, 30500 + os.getpid() % 1000, str(tmp_path), UNEQUAL_CASES), nprocs=2, join=True)
    _check_unequal(tmp_path, 2, UNEQUAL_CASES)


def test_all_gather_three_ranks_short_last_block(tmp_path):
    cases = [(5, True, False), (5, False, True), (2, True, False)]   # per = 2: shards 2, 2, 1; per = 1: shards 1, 1, 0
    mp.spawn(_worker_unequal, args=(3, 31500 + os.getpid() % 1000, str(tmp_path), cases), nprocs=3, join=True)
    _check_unequal(tmp_path, 3, cases)


def test_feature_store_shards_reassemble(tmp_path):
    """Every rank reads only its own contiguous block of videos from the packed store (retrieval.shard_range); the blocks tile
    the store exactly, in order — host logic, no GPU."""
    import numpy as np
    import torch

    from hirest_b200 import feature_store, retrieval

    g = torch.Generator().manual_seed(0)
    vids = [(f"v{i}", torch.randn(int(torch.randint(1, 9, (1,), generator=g)), 16, generator=g)) for i in range(11)]
    path = str(tmp_path / "s.hbf")
    feature_store.pack_features(vids, path)
    st = feature_store.FeatureStore(path)
    for world in (1, 2, 3, 4, 8, 16):
        rows, names = [], []
        for rank in range(world):
            lo, hi = retrieval.shard_range(len(st), rank, world)
            feats, offs = st.to_device("cpu", video_range=(lo, hi))
            assert offs[0] == 0 and offs.numel() == hi - lo + 1
            rows.append(feats)
            names += st.video_ids[lo:hi]
        assert names == st.video_ids
        assert torch.equal(torch.cat(rows), torch.from_numpy(np.array(st.data)))


class _FakeMoment:
    """Deterministic, batch-independent stand-in for MomentModel.test_step (host logic test: sharding + gathering only)."""

    def test_step(self, batch, **kw):
        task = batch["tasks"][0]
        n = batch["vis_mask"].sum(1).tolist()
        key = [int(round(float(v[:k].sum()) * 1000)) for v, k in zip(batch["vis_feats"], n)]
        if task == "moment_retrieval":
            return {"prediction": [[1 + k % 3, nn - 2 - k % 2] for k, nn in zip(key, n)]}
        if task == "moment_segmentation":
            out = []
            for k, (a, b) in zip(key, batch["moment_bound_frames"].tolist()):
                step = 5 + k % 4
                out.append(list(range(a, b + 1, step)) or [a])
            return {"prediction": out}
        return {"prediction": [f"cap-{k}-{nn}" for k, nn in zip(key, n)]}


def _fake_videos(n=7):
    g = torch.Generator().manual_seed(3)
    vids = []
    for i in range(n):
        T = int(torch.randint(30, 60, (1,), generator=g))
        vids.append({"prompt": f"p{i % 3}", "fname": f"v{i}", "video_duration": T + 0.4, "vis_feats": torch.randn(T, 8, generator=g),
                     "asr_feats": torch.randn(T, 4, generator=g), "clip_text_ids": torch.zeros(77, dtype=torch.long)})
    return vids


def _worker_pipeline(rank, world, port, out_dir):
    from hirest_b200 import pipeline

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    out = pipeline.run_end_to_end(_FakeMoment(), _fake_videos(), batch_size=2, num_beams=3, rank=rank, world=world)
    torch.save(out, os.path.join(out_dir, f"e2e_{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_end_to_end_chain_equals_single_process(tmp_path):
    """SURVEY.md §8(e), configs[3]/[4]: items sharded with DistributedSampler semantics (7 videos on 2 ranks: one padded duplicate),
    results gathered as Python objects -> every rank holds the single-process result."""
    from hirest_b200 import pipeline

    assert pipeline.shard_indices(7, 0, 2) == [0, 2, 4, 6] and pipeline.shard_indices(7, 1, 2) == [1, 3, 5, 0]
    assert pipeline.shard_indices(2, 2, 4) == [0] and pipeline.shard_indices(0, 1, 4) == [] and pipeline.shard_indices(5, 0, 1) == [0, 1, 2, 3, 4]
    import torch.utils.data as tud

    class _DS(tud.Dataset):
        def __len__(self):
            return 7

    for r in range(3):   # the reference's sampler (hirest_dataset.py:604-606)
        assert list(tud.distributed.DistributedSampler(_DS(), num_replicas=3, rank=r, shuffle=False)) == pipeline.shard_indices(7, r, 3)
    ref = pipeline.run_end_to_end(_FakeMoment(), _fake_videos(), batch_size=2, num_beams=3)
    assert all(isinstance(v["video_duration"], int) for p in ref["moment_retrieval"].values() for v in p.values())   # round(), dataset :145
    mp.spawn(_worker_pipeline, args=(2, 32500 + os.getpid() % 1000, str(tmp_path)), nprocs=2, join=True)
    for r in range(2):
        assert torch.load(os.path.join(str(tmp_path), f"e2e_{r}.pt")) == ref
