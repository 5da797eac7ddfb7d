"""Per-category device time (library profiler: CUDA events around every launch) of MomentModel.test_step for the three tasks.
    python tools/profile_moment.py [B] [T]"""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from hirest_b200 import _lib, synthetic
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__))))
import bench_extra

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
T = int(sys.argv[2]) if len(sys.argv) > 2 else 300
dev = torch.device("cuda:0")
lib = _lib.init(0)
vpath, _ = bench_extra._vocab_file()
model, sd = bench_extra._chain_model(dev, B * T, B, vpath)
batch = synthetic.make_chain_batch(B, T, seed=12)
res = {}
for task, kw in (("moment_retrieval", {}), ("moment_segmentation", {}), ("step_captioning", {"num_beams": 3})):
    b = dict(batch)
    b["tasks"] = [task] * B
    for _ in range(3):   # the second search of a shape records the decoder's CUDA graphs: keep that out of the timing
        model.test_step(b, **kw)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(3):
        model.test_step(b, **kw)
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) / 3 * 1e3
    lib.hb_profile_start()
    n0 = lib.hb_launch_count()
    model.test_step(b, **kw)
    prof = _lib.HbProfileSummary()
    _lib.check(lib.hb_profile_stop(prof))
    res[task] = {"wall_ms": round(wall, 2), "launches": int(lib.hb_launch_count() - n0),
                 "device_ms": {c: round(prof.ms[i], 2) for i, c in enumerate(_lib.PROF_CATEGORIES) if prof.launches[i]},
                 "device_ms_total": round(sum(prof.ms), 2)}
print(json.dumps(res, indent=1))
