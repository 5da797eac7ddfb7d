"""BASELINE.json configs[4]-shaped timing: the in-memory moment-retrieval -> segmentation -> step-captioning chain
(hirest_b200.pipeline) on synthetic videos, next to the CPU restatement of the reference's
disk-based chain (oracle/pipeline_oracle.py) on a bounded sample.  Not the bench.py headline.

    python tools/bench_pipeline.py [--videos 256] [--tmin 120] [--tmax 600] [--beam 3] [--cpu-videos 2]
"""
import argparse
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hirest_b200 import moment, pipeline, synthetic  # noqa: E402


class TableText:
    def __init__(self, table):
        self.table = table

    def encode_text(self, ids):
        return self.table[ids[:, 1].cpu()].to(ids.device)


def vocab():
    v = [f"[unused{i}]" for i in range(30522)]
    v[0], v[100], v[101], v[102], v[103] = "[PAD]", "[UNK]", "[CLS]", "[SEP]", "[MASK]"
    for i in range(1000, 30522):
        v[i] = f"w{i}"
    p = "/tmp/hb_vocab.txt"
    open(p, "w").write("\n".join(v) + "\n")
    return p, v


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--videos", type=int, default=256)
    ap.add_argument("--tmin", type=int, default=120)
    ap.add_argument("--tmax", type=int, default=600)
    ap.add_argument("--beam", type=int, default=3)
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--cpu-videos", type=int, default=2)
    ap.add_argument("--caption-batch", type=int, default=0, help="step items per beam search (0 = the pipeline default, 8 x --batch)")
    ap.add_argument("--profile", action="store_true", help="cProfile of the timed job (host side), top functions to stderr")
    ap.add_argument("--ab", type=int, default=0, help="A/B of the prefetch thread: N alternating pairs of whole jobs (inline, prefetch), seconds each")
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(17)
    n_prompts = max(1, a.videos // 4)
    table = torch.randn(n_prompts, 1024, generator=g)
    videos, test, feats = [], {}, {}
    for k in range(a.videos):
        pi = k % n_prompts
        p = f"prompt {pi}"
        T = int(torch.randint(a.tmin, a.tmax + 1, (1,), generator=g))
        vis = torch.randn(T, 1024, generator=g)
        vis = vis / vis.norm(dim=-1, keepdim=True)
        asr = torch.randn(T, 384, generator=g)
        ids = torch.zeros(77, dtype=torch.long)
        ids[0], ids[1], ids[2] = 49406, pi, 49407
        fn = f"vid{k:04d}"
        videos.append({"prompt": p, "fname": fn, "video_duration": T, "vis_feats": vis, "asr_feats": asr, "clip_text_ids": ids})
        if k < a.cpu_videos:
            test.setdefault(p, {})[fn] = {"video_duration": T}
            feats[fn] = {"vis_feats": vis, "asr_feats": asr, "text_feat": {p: table[pi]}}
    vpath, vlist = vocab()
    sd = synthetic.make_moment_state_dict(seed=3)
    m = moment.MomentModel(-1, 384, moment.default_args(bert_vocab_path=vpath), clip_model=TableText(table), max_rows=a.batch * a.tmax,
                           max_batch=max(a.batch, a.caption_batch or 8 * a.batch))
    m.load_state_dict(sd, strict=True)
    m = m.to(dev)
    # warm-up: the whole job once (engine / decoder builds for every batch shape, CUDA-graph capture of the decode steps)
    pipeline.run_end_to_end(m, videos, batch_size=a.batch, num_beams=a.beam, caption_batch_size=a.caption_batch or None)
    torch.cuda.synchronize()
    if a.ab > 0:
        pairs = []
        for _ in range(a.ab):
            row = []
            for pf in (False, True):
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                pipeline.run_end_to_end(m, videos, batch_size=a.batch, num_beams=a.beam, caption_batch_size=a.caption_batch or None, prefetch=pf)
                torch.cuda.synchronize()
                row.append(round(time.perf_counter() - t0, 4))
            pairs.append(row)
        print(json.dumps({"op": "pipeline.run_end_to_end, prefetch thread A/B", "videos": a.videos, "seconds_inline_prefetch": pairs,
                          "median_inline": sorted(r[0] for r in pairs)[len(pairs) // 2], "median_prefetch": sorted(r[1] for r in pairs)[len(pairs) // 2]}))
        return
    stage = {}
    orig = m.test_step

    def timed(batch, **kw):   # per-task time inside the model calls (the rest of the wall time is host glue: collate, H2D)
        torch.cuda.synchronize()
        t = time.perf_counter()
        r = orig(batch, **kw)
        torch.cuda.synchronize()
        stage[batch["tasks"][0]] = stage.get(batch["tasks"][0], 0.0) + time.perf_counter() - t
        return r

    m.test_step = timed
    prof = None
    if a.profile:
        import cProfile
        prof = cProfile.Profile()
        prof.enable()
    t0 = time.perf_counter()
    out = pipeline.run_end_to_end(m, videos, batch_size=a.batch, num_beams=a.beam, caption_batch_size=a.caption_batch or None)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    if prof is not None:
        import pstats
        prof.disable()
        pstats.Stats(prof, stream=sys.stderr).sort_stats("cumulative").print_stats(45)
    m.test_step = orig
    n_steps = sum(len(x["steps"]) for p in out["final"].values() for x in p.values())
    res = {"op": "pipeline.run_end_to_end (MR -> MS -> SC, in memory)", "videos": a.videos, "frames": [a.tmin, a.tmax], "beam": a.beam,
           "batch": a.batch, "caption_batch": a.caption_batch or 8 * a.batch, "seconds": dt, "videos_per_s": a.videos / dt, "steps_captioned": n_steps,
           "model_seconds": {k: round(v, 3) for k, v in stage.items()}}
    if a.cpu_videos > 0:
        from oracle import pipeline_oracle as po
        torch.set_num_threads(os.cpu_count())
        t0 = time.perf_counter()
        with torch.no_grad():
            ref = po.run_end_to_end(sd, test, feats, vlist, batch_size=a.batch, num_beams=a.beam)
        cdt = time.perf_counter() - t0
        same = all(out["final"][p][v]["steps"] == x["steps"] and out["final"][p][v]["bounds"] == x["bounds"]
                   for p, vs in ref["final"].items() for v, x in vs.items())
        res["cpu_baseline"] = {"value": a.cpu_videos / cdt, "unit": "videos/s", "cores": os.cpu_count(), "kind": "port",
                               "sample": f"first {a.cpu_videos} videos", "identical_to_gpu": same}
    print(json.dumps(res))


if __name__ == "__main__":
    main()
