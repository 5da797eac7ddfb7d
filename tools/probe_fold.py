import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from hirest_b200 import _lib, eva_clip, synthetic
lib = _lib.init(0)
dev = torch.device("cuda:0")
cfg = synthetic.EVA_TINY
sd = synthetic.make_eva_state_dict(cfg, 0)
frames = synthetic.make_frames(5, 224, seed=9).to(dev)
D = cfg["vision_cfg"]["width"]
for fold in (1, 0):
    _lib.check(lib.hb_debug_set(b"ln_fold", fold))
    m = eva_clip.EVA_CLIP(**cfg, max_image_batch=8, max_text_batch=8); m.load_state_dict(sd); m = m.to(dev).eval()
    a = m.encode_image(frames); b = m.encode_image(frames); c = m.encode_image(frames[:1])
    print("fold", fold, "run-to-run max diff", float((a - b).abs().max()), "single-vs-batch", float((a[:1] - c).abs().max()))
    for layer in (0, 1, 2, 3):
        t1 = torch.empty(5 * 257, D, device=dev); t2 = torch.empty(5 * 257, D, device=dev)
        m.visual(frames, tap=(layer, t1)); m.visual(frames, tap=(layer, t2)); torch.cuda.synchronize()
        print("   tap", layer, "run-to-run max diff", float((t1 - t2).abs().max()), "abs mean", float(t1.abs().mean()))
