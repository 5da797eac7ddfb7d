"""One MomentModel.test_step per task (after a warm-up), for ncu launch lists.  python tools/moment_once.py [B] [T] [task]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch
from hirest_b200 import _lib, synthetic
import bench_extra
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
T = int(sys.argv[2]) if len(sys.argv) > 2 else 300
task = sys.argv[3] if len(sys.argv) > 3 else "moment_segmentation"
dev = torch.device("cuda:0")
_lib.init(0)
vpath, _ = bench_extra._vocab_file()
model, sd = bench_extra._chain_model(dev, B * T, B, vpath)
b = synthetic.make_chain_batch(B, T, seed=12)
b["tasks"] = [task] * B
kw = {"num_beams": 3} if task == "step_captioning" else {}
model.test_step(b, **kw)
torch.cuda.synchronize()
torch.cuda.profiler.start()
model.test_step(b, **kw)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
