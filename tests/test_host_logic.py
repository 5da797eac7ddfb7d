"""Host-side mirror of the reference interface: state_dict layout, error behaviour, ranking, sharding (no GPU)."""
import numpy as np
import pytest
import torch

from hirest_b200 import eva_clip, retrieval, synthetic
from oracle import eva_oracle


def test_state_dict_layout_matches_reference_keys():
    cfg = synthetic.EVA_TINY
    model = eva_clip.EVA_CLIP(**cfg)
    sd = synthetic.make_eva_state_dict(cfg, seed=0)
    assert sorted(model.state_dict().keys()) == sorted(sd.keys())
    assert {k: tuple(v.shape) for k, v in model.state_dict().items()} == {k: tuple(v.shape) for k, v in sd.items()}
    res = model.load_state_dict(sd, strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    assert torch.equal(model.state_dict()["visual.blocks.1.attn.qkv.weight"], sd["visual.blocks.1.attn.qkv.weight"])


def test_g14_parameter_counts_match_survey():
    cfg = eva_clip.get_model_config("EVA_CLIP_g_14")
    vis = sum(int(np.prod(s)) for s in eva_clip._visual_shapes(cfg["embed_dim"], cfg["vision_cfg"]).values())
    txt = sum(int(np.prod(s)) for s in eva_clip._text_shapes(cfg["embed_dim"], cfg["text_cfg"]).values())
    assert round(vis / 1e6, 2) == 1012.59 and round(txt / 1e6, 2) == 123.85


def test_no_cpu_fallback_and_shape_asserts():
    cfg = synthetic.EVA_TINY
    model = eva_clip.EVA_CLIP(**cfg).eval()
    with pytest.raises(AssertionError, match="doesn't match model"):  # vit_model.py:203-204
        model.encode_image(torch.zeros(1, 3, 200, 224))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        model.encode_image(torch.zeros(1, 3, 224, 224))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        model.encode_text(torch.zeros(1, 77, dtype=torch.int64))
    with pytest.raises(RuntimeError, match="not found"):
        eva_clip.create_model("no_such_model", "x.pt")


def test_rank_videos_equals_reference_rule_with_ties():
    rng = np.random.default_rng(0)
    names = [f"v{int(i):03d}" for i in rng.permutation(50)]
    scores = np.round(rng.normal(size=50), 1)  # many exact ties
    assert retrieval.rank_videos(scores, names) == eva_oracle.rank_videos(scores.tolist(), names)


def test_shard_range_covers_everything_in_order():
    for n, w in ((4096, 8), (10, 4), (3, 8), (0, 2)):
        spans = [retrieval.shard_range(n, r, w) for r in range(w)]
        flat = [i for lo, hi in spans for i in range(lo, hi)]
        assert flat == list(range(n))


def test_synthetic_tokens_have_eot_as_row_max():
    cfg = synthetic.EVA_G14
    t = synthetic.make_tokens(16, cfg, seed=3)
    assert t.shape == (16, 77) and (t[:, 0] == 49406).all()
    am = t.argmax(-1)
    assert (t[torch.arange(16), am] == 49407).all() and (am >= 4).all()


def test_moment_model_state_dict_layout_and_errors():
    from hirest_b200 import moment

    class Stub:
        def encode_text(self, ids):
            return None

    m = moment.MomentModel(-1, 384, moment.default_args(), clip_model=Stub())
    sd = synthetic.make_moment_state_dict(seed=3)
    assert {k: tuple(v.shape) for k, v in m.state_dict().items()} == {k: tuple(v.shape) for k, v in sd.items()}
    assert sum(v.numel() for v in sd.values()) == 86_627_901  # 86.63 M state-dict elements
    res = m.load_state_dict(sd, strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    assert m.args.max_frames == 20 and m.args.d_model == 512  # args mutated like modeling.py:103-105
    with pytest.raises(NotImplementedError):
        m.train_step({"tasks": ["moment_retrieval"]})
    b = synthetic.make_moment_batch(1, 16, seed=1)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m.foward_moment_shared(b["vis_feats"], b["text_feat"], b["vis_mask"], moment_mask=b["moment_mask"], asr_feats=b["asr_feats"])


def test_bench_reference_arm_json_contract():
    """`bench.py --impl reference` (CPU oracle port) prints one JSON line with the contract keys; tiny config keeps it fast."""
    import json
    import os
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--tiny", "--steps", "2", "--warmup", "1",
                          "--cpu-frames", "4"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-500:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "dtype",
              "data", "config", "cpu_baseline", "e2e"):
        assert k in line, k
    assert line["impl"] == "reference" and line["unit"] == "frames/s" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["e2e"]["h2d_bytes_per_step"] == 0
    # non-zero ranks of a torchrun launch exit quietly
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--tiny"], capture_output=True, text=True,
                         timeout=120, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_gemm_column_tiling_partitions_n():
    """Host view of the kernel's N tiling (hb_gemm_n_tiling): tiles partition [0, N), are <= 256 wide, multiples of 16 * cta_group,
    balanced tiles differ by at most one unit, and the ViT shapes come out as documented in DESIGN.md §3.1."""
    import numpy as np

    from hirest_b200 import _lib

    lib = _lib.load()

    def tiles(N, cg, bal):
        n0 = np.zeros(256, np.int32)
        w = np.zeros(256, np.int32)
        n = lib.hb_gemm_n_tiling(N, cg, bal, n0.ctypes.data, w.ctypes.data, 256)
        assert n > 0
        return n0[:n].tolist(), w[:n].tolist()

    for cg in (1, 2):
        unit = 16 * cg
        for N in [unit * k for k in (1, 2, 3, 7, 8, 9, 15, 16, 17, 44, 88, 132, 192, 954)]:
            for bal in (0, 1):
                n0, w = tiles(N, cg, bal)
                assert n0[0] == 0 and sum(w) == N and all(a + b == c for a, b, c in zip(n0, w, n0[1:] + [N]))
                assert max(w) <= 256 and all(x % unit == 0 and x > 0 for x in w)
                assert len(w) == (N + 255) // 256
                if bal:
                    assert max(w) - min(w) <= unit
    assert tiles(1408, 2, 1)[1] == [256, 256, 224, 224, 224, 224]
    assert tiles(1408, 2, 0)[1] == [256] * 5 + [128]
    assert sorted(tiles(4224, 2, 1)[1], reverse=True) == [256] * 13 + [224] * 4
    assert tiles(6144, 2, 1)[1] == [256] * 24
    assert lib.hb_gemm_n_tiling(0, 2, 1, None, None, 0) < 0
