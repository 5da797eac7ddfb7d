"""Generates tests/golden/dataset.pt by running the UNMODIFIED reference MomentDataset (hirest_dataset.py: __init__ item construction,
__getitem__ feature loading / subsample / repeat-pad / ASR warping, collate_fn) on a synthetic data directory.
Build container only:    python oracle/make_golden_dataset.py

Shims (no arithmetic of the path): `srt.parse` (a 20-line SubRip reader returning objects with timedelta .start / .end — the pip
package is not installed), `clip.tokenize` (a deterministic stand-in: the BPE tokenizer is pinned separately by tokenizer.json),
boto3 & co.  The BERT vocabulary is the synthetic WordPiece vocabulary of tests/golden/wordpiece.json.  np.long is aliased to
np.int64 (removed from numpy 1.24; the reference's clip4cap_get_text still uses it)."""
import datetime
import json
import os
import re
import shutil
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_moment  # noqa: E402

WORK = "/tmp/hb_dataset_golden"
C_VIS, C_ASR = 8, 4


def fake_clip_tokenize(prompts):
    out = torch.zeros((len(prompts), 77), dtype=torch.long)
    for i, p in enumerate(prompts):
        out[i, 0], out[i, 1], out[i, 2] = 49406, len(p), 49407
    return out


def srt_shim():
    m = types.ModuleType("srt")
    pat = re.compile(r"(\d+):(\d+):(\d+)[,.](\d+)\s*-->\s*(\d+):(\d+):(\d+)[,.](\d+)")

    def parse(text):
        for mm in pat.finditer(text):
            h0, m0, s0, ms0, h1, m1, s1, ms1 = (int(x) for x in mm.groups())
            yield types.SimpleNamespace(start=datetime.timedelta(hours=h0, minutes=m0, seconds=s0, milliseconds=ms0),
                                        end=datetime.timedelta(hours=h1, minutes=m1, seconds=s1, milliseconds=ms1))

    m.parse = parse
    return m


def annotations():
    def steps(*bounds):
        return [{"index": i, "heading": h, "absolute_bounds": [a, b]} for i, (a, b, h) in enumerate(bounds)]

    return {
        "make iced coffee": {
            "vidA.mp4": {"relevant": True, "clip": True, "v_duration": 20.4, "bounds": [3, 17],
                         "steps": steps((3, 7, "boil the water"), (7, 12, " pour over ice "), (12, 17, "add milk & sugar"))},
            "vidB.mp4": {"relevant": True, "clip": True, "v_duration": 75.6, "bounds": [10, 60],
                         "steps": steps((10, 25, "chop the onions"), (25, 26, "fry"), (40, 60, "season; serve!"))},
            "vidX.mp4": {"relevant": False, "clip": False, "v_duration": 50.0, "bounds": [0, 0], "steps": []},
        },
        "fix a flat tire": {
            "vidC.mp4": {"relevant": True, "clip": True, "v_duration": 32.0, "bounds": [0, 31], "steps": steps((0, 31, "remove the wheel"))},
            "vidD.mp4": {"relevant": True, "clip": True, "v_duration": 47.5, "bounds": [5, 40], "steps": []},
            "vidY.mp4": {"relevant": True, "clip": False, "v_duration": 10.0, "bounds": [0, 0], "steps": []},
        },
    }


SUBS = {   # (start, end) seconds incl. overlaps, zero-length, beyond-the-end and out-of-order blocks
    "vidA": [(0, 4), (2, 6), (10, 10), (15, 30)],
    "vidB": [(5, 9), (9, 20), (70, 90), (1, 3), (100, 120)],
    "vidC": [(0, 32)],
    "vidD": [],
}


def build_workdir():
    shutil.rmtree(WORK, ignore_errors=True)
    for d in ("data", "vis", "asr", "asr_feats"):
        os.makedirs(os.path.join(WORK, d))
    ann = annotations()
    with open(os.path.join(WORK, "data", "all_data_test.json"), "w") as f:
        json.dump(ann, f)
    g = torch.Generator().manual_seed(21)
    feats = {}
    for prompt in ann.values():
        for fname, a in prompt.items():
            if not (a["relevant"] and a["clip"]):
                continue
            T = round(a["v_duration"])
            vid = fname.replace(".mp4", "")
            v = torch.randn(T, C_VIS, generator=g)
            torch.save(v, os.path.join(WORK, "vis", f"{fname}.pt"))
            subs = SUBS[vid]
            af = torch.randn(len(subs), C_ASR, generator=g)
            torch.save(af, os.path.join(WORK, "asr_feats", f"{vid}.pt"))

            def ts(s):
                return f"{s // 3600:02d}:{(s % 3600) // 60:02d}:{s % 60:02d},{(s * 137) % 1000:03d}"

            with open(os.path.join(WORK, "asr", f"{vid}.srt"), "w") as f:
                for i, (a0, b0) in enumerate(subs):
                    f.write(f"{i + 1}\n{ts(a0)} --> {ts(b0)}\nsentence {i}\n\n")
            feats[fname] = {"vis": v, "asr": af, "subs": subs}
    return ann, feats


def to_plain(item):
    out = {}
    for k, v in item.items():
        if isinstance(v, tuple):   # target_text: tuple of numpy arrays
            out[k] = [torch.from_numpy(np.asarray(x)) if isinstance(x, np.ndarray) else x for x in v]
        else:
            out[k] = v
    return out


def main():
    with open(os.path.join(ROOT, "tests", "golden", "wordpiece.json"), encoding="utf-8") as f:
        vocab = json.load(f)["vocab"]
    ann, feats = build_workdir()
    dst = ref_moment.prepare_copy("/tmp/hirest_ref_copy_ds")
    with open(os.path.join(dst, "clip4caption", "modules", "bert-base-uncased", "vocab.txt"), "w", encoding="utf-8") as f:
        f.write("\n".join(vocab) + "\n")
    ref_moment.install_stubs()
    sys.modules["srt"] = srt_shim()
    sys.modules["clip"] = types.SimpleNamespace(tokenize=fake_clip_tokenize, clip=types.SimpleNamespace(_transform=lambda n: None))
    if not hasattr(np, "long"):
        np.long = np.int64
    os.chdir(dst)
    sys.path.insert(0, dst)
    import hirest_dataset as ref_ds

    out = {"annotations": ann, "features": feats, "c_vis": C_VIS, "c_asr": C_ASR, "cases": {}}
    for nmf in (-1, 32):
        for e2e in (False, True):
            for task in ("moment_retrieval", "moment_segmentation", "step_captioning"):
                args = types.SimpleNamespace(end_to_end=e2e, max_words=48)
                try:
                    ds = ref_ds.MomentDataset(args, os.path.join(WORK, "data", "all_data_test.json"), video_feature_dir=os.path.join(WORK, "vis"),
                                              asr_dir=os.path.join(WORK, "asr"), asr_feature_dir=os.path.join(WORK, "asr_feats"),
                                              n_model_frames=nmf, task=task)
                except IndexError as ex:   # step captioning + end_to_end + a video without steps: steps[0] (hirest_dataset.py:276)
                    out["cases"][(nmf, e2e, task)] = {"raises": "IndexError"}
                    print(nmf, e2e, task, "-> IndexError", ex)
                    continue
                items = [ds[i] for i in range(len(ds))]
                batches = [ds.collate_fn(items)] + [ds.collate_fn(items[i:i + 2]) for i in range(0, len(items), 2)]
                out["cases"][(nmf, e2e, task)] = {"items": [to_plain(x) for x in items], "batches": batches}
                print(nmf, e2e, task, len(items), "items", [tuple(b["vis_feats"].shape) for b in batches])
    path = os.path.join(ROOT, "tests", "golden", "dataset.pt")
    torch.save(out, path)
    print("saved", os.path.getsize(path) / 1e3, "KB")


if __name__ == "__main__":
    main()
