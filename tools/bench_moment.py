"""Auxiliary timings for BASELINE.json configs[3] / [4]: MomentModel test_step on synthetic clips (GPU path vs CPU oracle).
Not the bench.py headline; results are recorded in DESIGN.md §6.   python tools/bench_moment.py [--B 64] [--T 300]"""
import argparse, os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from hirest_b200 import moment, synthetic

ap = argparse.ArgumentParser()
ap.add_argument("--B", type=int, default=64)
ap.add_argument("--T", type=int, default=300)
ap.add_argument("--beam", type=int, default=3)
ap.add_argument("--cpu-B", type=int, default=4)
ap.add_argument("--no-cpu", action="store_true")
a = ap.parse_args()
dev = torch.device("cuda:0")


class FixedText:
    feat = None
    def encode_text(self, ids): return self.feat.to(ids.device)


def vocab_file():
    p = "/tmp/hb_vocab.txt"
    v = [f"[unused{i}]" for i in range(30522)]
    v[0], v[100], v[101], v[102], v[103] = "[PAD]", "[UNK]", "[CLS]", "[SEP]", "[MASK]"
    for i in range(1000, 30522): v[i] = f"w{i}"
    open(p, "w").write("\n".join(v) + "\n")
    return p

clip = FixedText()
m = moment.MomentModel(-1, 384, moment.default_args(bert_vocab_path=vocab_file()), clip_model=clip, max_rows=a.B * a.T, max_batch=a.B)
sd = synthetic.make_moment_state_dict(seed=3)
m.load_state_dict(sd, strict=True)
m = m.to(dev)
batch = synthetic.make_moment_batch(a.B, a.T, seed=11, ragged=False)
batch["moment_bound_frames"] = torch.tensor([[30, a.T - 30]] * a.B)
batch["moment_mask"] = ((torch.arange(a.T)[None] >= 30) & (torch.arange(a.T)[None] <= a.T - 30)).long().expand(a.B, -1).contiguous()
clip.feat = batch["text_feat"]
res = {"B": a.B, "T": a.T, "beam": a.beam}
for task in ("moment_retrieval", "moment_segmentation", "step_captioning"):
    batch["tasks"] = [task] * a.B
    kw = {"num_beams": a.beam} if task == "step_captioning" else {}
    m.test_step(batch, **kw); torch.cuda.synchronize()
    t0 = time.perf_counter(); n = 3
    for _ in range(n): out = m.test_step(batch, **kw)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / n
    res[task] = {"gpu_s_per_batch": dt, "gpu_clips_per_s": a.B / dt}
if not a.no_cpu:
    from oracle import moment_oracle as mo, caption_oracle as co
    torch.set_num_threads(os.cpu_count())
    cb = {k: (v[:a.cpu_B] if torch.is_tensor(v) else v) for k, v in batch.items()}
    tf = cb["text_feat"]
    with torch.no_grad():
        t0 = time.perf_counter(); mo.test_moment_retrieval(sd, cb, tf); res["moment_retrieval"]["cpu_clips_per_s"] = a.cpu_B / (time.perf_counter() - t0)
        t0 = time.perf_counter(); mo.test_moment_segmentation(sd, cb, tf); res["moment_segmentation"]["cpu_clips_per_s"] = a.cpu_B / (time.perf_counter() - t0)
        t0 = time.perf_counter(); co.test_step_captioning(sd, cb, tf, beam=a.beam); res["step_captioning"]["cpu_clips_per_s"] = a.cpu_B / (time.perf_counter() - t0)
    res["cpu_cores"] = os.cpu_count(); res["cpu_sample_clips"] = a.cpu_B
print(json.dumps(res))
