"""In-process A/B of one MomentModel / pipeline switch on the BASELINE configs[4] chain (the bench.py --config e2e job: 256 synthetic
videos, real EVA text tower): alternating whole jobs with the switch off / on, seconds each.  Box-to-box variation of the chain
is ~7 %, so a switch is judged inside one process.

    python tools/ab_chain.py --toggle dedup_prompts|ms_early_exit|prefetch [--pairs 4] [--videos 256]
"""
import argparse
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
from hirest_b200 import pipeline, synthetic  # noqa: E402
import bench_extra  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--toggle", default="dedup_prompts", choices=["dedup_prompts", "ms_early_exit", "prefetch"])
    ap.add_argument("--pairs", type=int, default=4)
    ap.add_argument("--videos", type=int, default=256)
    ap.add_argument("--beam", type=int, default=3)
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    n, tmin, tmax, bs, cbs = a.videos, 120, 600, 64, 512
    vpath, _ = bench_extra._vocab_file()
    model, _ = bench_extra._chain_model(dev, bs * tmax, cbs, vpath)
    g = torch.Generator().manual_seed(17)
    n_prompts = max(1, n // 4)
    prompt_ids = synthetic.make_tokens(n_prompts, synthetic.EVA_G14, seed=77)
    videos = []
    for k in range(n):
        pi = k % n_prompts
        T = int(torch.randint(tmin, tmax + 1, (1,), generator=g))
        vis = torch.randn(T, 1024, generator=g)
        videos.append({"prompt": f"prompt {pi}", "fname": f"vid{k:04d}", "video_duration": T, "vis_feats": vis / vis.norm(dim=-1, keepdim=True),
                       "asr_feats": torch.randn(T, 384, generator=g), "clip_text_ids": prompt_ids[pi]})

    def job(on):
        kw = {}
        if a.toggle == "prefetch":
            kw["prefetch"] = on
        else:
            setattr(model, a.toggle, on)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out = pipeline.run_end_to_end(model, videos, batch_size=bs, num_beams=a.beam, caption_batch_size=cbs, **kw)
        torch.cuda.synchronize()
        return time.perf_counter() - t0, out

    _, ref = job(True)   # warm-up: engines, decoder graphs
    job(False)
    rows, same = [], True
    for _ in range(a.pairs):
        t_off, o_off = job(False)
        t_on, o_on = job(True)
        same = same and o_off == ref and o_on == ref
        rows.append([round(t_off, 4), round(t_on, 4)])
    med = lambda xs: sorted(xs)[len(xs) // 2]  # noqa: E731
    print(json.dumps({"op": f"run_end_to_end A/B: {a.toggle} off / on", "videos": n, "seconds_off_on": rows,
                      "median_off": med([r[0] for r in rows]), "median_on": med([r[1] for r in rows]), "identical_results": same}))


if __name__ == "__main__":
    main()
