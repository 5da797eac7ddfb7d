"""Packed feature cache: container round trip on the CPU, ragged subsample + pool kernel on the GPU
against the oracle restatement of inference_video_retrieval.py:298-327."""
import os

import numpy as np
import pytest
import torch

from hirest_b200 import feature_store, retrieval
from oracle import eva_oracle


def _videos(n, dim, seed, tmin=1, tmax=700):
    g = torch.Generator().manual_seed(seed)
    lens = torch.randint(tmin, tmax, (n,), generator=g).tolist()
    lens[0] = 1            # single-frame video: linspace(0, 0, F) -> all zeros
    if n > 2:
        lens[1] = 2
        lens[2] = 32       # exactly n_model_frames
    return [(f"vid_{i:04d}", torch.randn(t, dim, generator=g)) for i, t in enumerate(lens)]


def test_pack_and_reopen_roundtrip(tmp_path):
    vids = _videos(9, 64, seed=0)
    path = str(tmp_path / "feats.hbf")
    feature_store.pack_features(vids, path)
    st = feature_store.FeatureStore(path)
    assert st.video_ids == [v for v, _ in vids] and st.dim == 64 and len(st) == 9
    assert st.offsets.tolist() == np.cumsum([0] + [f.shape[0] for _, f in vids]).tolist()
    for vid, f in vids:
        assert torch.equal(st.features(vid), f)


def test_pack_feature_dir_reads_reference_pickles(tmp_path):
    """The per-video .pt files the reference writes (inference_video_retrieval.py:275-280) convert without change of value."""
    vids = _videos(4, 32, seed=1)
    for vid, f in vids:
        torch.save(f, str(tmp_path / f"{vid}.pt"))
    path = str(tmp_path / "feats.hbf")
    feature_store.pack_feature_dir(str(tmp_path), [v for v, _ in vids], path)
    st = feature_store.FeatureStore(path)
    for vid, f in vids:
        assert torch.equal(st.features(vid), torch.load(str(tmp_path / f"{vid}.pt")))


def test_rejects_bad_input(tmp_path):
    with pytest.raises(ValueError):
        feature_store.pack_features([("a", torch.zeros(0, 8))], str(tmp_path / "x.hbf"))
    with pytest.raises(ValueError):
        feature_store.pack_features([("a", torch.zeros(2, 8)), ("b", torch.zeros(2, 9))], str(tmp_path / "x.hbf"))
    bad = tmp_path / "bad.hbf"
    bad.write_bytes(b"not a store")
    with pytest.raises(ValueError):
        feature_store.FeatureStore(str(bad))
    with pytest.raises(RuntimeError):   # no CPU path
        feature_store.pooled_video_embeddings(torch.zeros(4, 8), torch.tensor([0, 4]), 2)


@pytest.mark.gpu
@pytest.mark.parametrize("n_frames", [32, 1, 7, -1])
def test_pooled_embeddings_match_reference_loop(hb, tmp_path, n_frames):
    vids = _videos(40, 1024, seed=2)
    path = str(tmp_path / "feats.hbf")
    feature_store.pack_features(vids, path)
    st = feature_store.FeatureStore(path)
    feats, offs = st.to_device("cuda:0")
    got = feature_store.pooled_video_embeddings(feats, offs, n_frames).cpu()
    ref = torch.cat([eva_oracle.cached_video_embedding(f, n_frames) for _, f in vids], dim=0)
    # fp32 sums in a different order: 1e-6 relative (the gathered rows themselves are identical)
    assert torch.allclose(got, ref, rtol=0, atol=2e-7 * 8)
    assert float((got - ref).norm() / ref.norm()) < 1e-6
    # a shard of the store gives the same rows (multi-GPU video sharding reads only its own block)
    f2, o2 = st.to_device("cuda:0", video_range=(10, 25))
    assert torch.equal(feature_store.pooled_video_embeddings(f2, o2, n_frames).cpu(), got[10:25])


@pytest.mark.gpu
def test_cached_path_ranking_at_reference_size(hb, tmp_path):
    """Reference split size (4282 videos, 546 prompts): top-10 lists of the packed-store path equal the
    reference loop's except where its fp32 scores tie within rounding."""
    g = torch.Generator().manual_seed(5)
    V, Q, E = 4282, 546, 1024
    lens = torch.randint(20, 400, (V,), generator=g).tolist()
    vids = [(f"v{i:05d}", torch.randn(t, E, generator=g)) for i, t in enumerate(lens)]
    path = str(tmp_path / "val.hbf")
    feature_store.pack_features(vids, path)
    st = feature_store.FeatureStore(path)
    feats, offs = st.to_device("cuda:0")
    v_hat = feature_store.pooled_video_embeddings(feats, offs, 32)
    t = torch.randn(Q, E, generator=g)
    scores = retrieval.similarity(retrieval.normalize(t.cuda()), v_hat, exact=True).cpu()
    ref_v = torch.cat([eva_oracle.cached_video_embedding(f, 32) for _, f in vids], dim=0)
    ref = eva_oracle.similarity(eva_oracle.normalize_text(t), ref_v)
    assert float((scores - ref).abs().max()) < 5e-7
    names = st.video_ids
    TIE = 1e-6
    tie_rows, bad_rows = set(), set()
    for q in range(Q):
        b = eva_oracle.rank_videos(ref[q].tolist(), names)[:11]
        if any(abs(float(ref[q, b[j]]) - float(ref[q, b[j + 1]])) < TIE for j in range(10)):
            tie_rows.add(q)
        a = retrieval.topk(scores[q:q + 1], names, 10)[0]
        if a != b[:10]:
            bad_rows.add(q)
            for x, y in zip(a, b):   # any disagreement must be between near-tied scores
                assert abs(float(ref[q, x]) - float(ref[q, y])) < TIE
    print(f"top-10 identical for {Q - len(bad_rows)}/{Q} queries; near-tie rows {len(tie_rows)}; differing rows {sorted(bad_rows)}")
    assert bad_rows <= tie_rows
