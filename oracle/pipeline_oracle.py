"""TEST INFRASTRUCTURE ONLY — CPU restatement of the reference's ``--end_to_end`` chain (run.py:383-490), following its
data flow literally: every task's predictions go into the dictionaries ``evaluate`` returns (run.py:704-830), the test
annotation dict is rewritten the way run.py:399-417 / :437-452 / :466-472 rewrite ``all_data_test.json``, and the next task's
dataset items are rebuilt from that dict (hirest_dataset.py:149-312) and collated (:409-531).  The model calls are the CPU
oracle's (oracle/moment_oracle.py, oracle/caption_oracle.py).

Pinning: the conversions are checked against the imported reference functions where the reference can be imported
(tests/test_pipeline.py::test_frame_conversions_match_reference runs only when /root/reference exists); the model calls are
pinned by tests/golden/moment.pt.
"""
from copy import deepcopy

import numpy as np
import torch

from oracle import caption_oracle as co
from oracle import moment_oracle as mo


def t2f(timestamp, video_duration, n_frames=32):
    """hirest_dataset.py:12-40."""
    video_duration = int(video_duration)
    if n_frames < 0:
        n_frames = video_duration
    bins = np.linspace(0, video_duration - 1, n_frames)
    bin_index = np.digitize(timestamp, bins, right=True)
    bin_index = min(bin_index, n_frames - 1)
    return int(bin_index)


def f2t(frame_index, video_duration, n_frames=32):
    """hirest_dataset.py:42-68."""
    video_duration = int(video_duration)
    if n_frames < 0:
        n_frames = video_duration
    bins = np.linspace(0, video_duration - 1, n_frames)
    return int(bins[frame_index])


def build_items(test, feats, task):
    """hirest_dataset.py:120-312 for split 'test', end_to_end=True, n_model_frames=-1."""
    data = []
    for prompt in test:
        for fname, ann in test[prompt].items():
            video_duration = ann["video_duration"]
            n_frames = int(video_duration)
            datum = {"prompt": prompt, "fname": fname, "video_duration": video_duration, "task": task,
                     "vis_feats": feats[fname]["vis_feats"], "asr_feats": feats[fname]["asr_feats"],
                     "text_feat": feats[fname]["text_feat"][prompt]}
            if task == "moment_retrieval":
                d = dict(datum)
                d["video_mask"] = torch.ones(n_frames, dtype=torch.long)
                d["moment_mask"] = torch.ones(n_frames, dtype=torch.long)
                data.append(d)
            elif task == "moment_segmentation":
                d = dict(datum)
                s = t2f(ann["bounds"][0], video_duration=video_duration, n_frames=n_frames)
                e = t2f(ann["bounds"][1], video_duration=video_duration, n_frames=n_frames)
                d["moment_bound_frames"] = [s, e]
                mm = torch.zeros(n_frames, dtype=torch.long)
                mm[s:e + 1] = 1
                d["moment_mask"] = mm
                d["video_mask"] = torch.ones(n_frames, dtype=torch.long)
                data.append(d)
            elif task == "step_captioning":
                for step in ann["steps"]:
                    ss, se = step["absolute_bounds"]
                    sf = t2f(ss, video_duration=video_duration, n_frames=n_frames)
                    ef = t2f(se, video_duration=video_duration, n_frames=n_frames)
                    d = dict(datum)
                    mm = torch.zeros(n_frames, dtype=torch.long)
                    mm[sf:ef] = 1
                    mm[ef] = 1
                    d["moment_mask"] = mm
                    d["video_mask"] = torch.ones(n_frames, dtype=torch.long)
                    data.append(d)
    return data


def collate(batch):
    """hirest_dataset.py:409-531, n_model_frames <= 0 branch."""
    lens = [d["vis_feats"].shape[0] for d in batch]
    mx = max(lens)
    out = {}
    out["vis_feats"] = torch.stack([torch.cat([d["vis_feats"], torch.zeros(mx - n, d["vis_feats"].shape[1])], dim=0) for d, n in zip(batch, lens)])
    out["vis_mask"] = torch.stack([torch.cat([d["video_mask"], torch.zeros(mx - n)], dim=0) for d, n in zip(batch, lens)]).long()
    out["moment_mask"] = torch.stack([torch.cat([d["moment_mask"], torch.zeros(mx - n)], dim=0) for d, n in zip(batch, lens)]).long()
    out["asr_feats"] = torch.stack([torch.cat([d["asr_feats"], torch.zeros(mx - n, d["asr_feats"].shape[1])], dim=0) for d, n in zip(batch, lens)]).float()
    if "moment_bound_frames" in batch[0]:
        out["moment_bound_frames"] = torch.LongTensor([d["moment_bound_frames"] for d in batch])
    out["text_feat"] = torch.stack([d["text_feat"] for d in batch])
    return out


def run_end_to_end(sd, test, feats, vocab, batch_size=64, num_beams=5, ms_threshold=0.5, ms_iterations=20):
    """``test``: {prompt: {fname: {"video_duration": d}}}; ``feats``: {fname: {"vis_feats", "asr_feats", "text_feat": {prompt: [1024]}}}."""
    test = deepcopy(test)
    # --- moment retrieval, run.py:388-417
    items = build_items(test, feats, "moment_retrieval")
    moments = {}
    for i in range(0, len(items), batch_size):
        chunk = items[i:i + batch_size]
        b = collate(chunk)
        pred, _, _ = mo.test_moment_retrieval(sd, b, b["text_feat"])
        for d, p in zip(chunk, pred):
            moments.setdefault(d["prompt"], {})[d["fname"]] = {
                "bounds": [f2t(p[0], d["video_duration"], n_frames=-1), f2t(p[1], d["video_duration"], n_frames=-1)],
                "video_duration": d["video_duration"]}
    mr = moments
    for prompt in test:
        for video in test[prompt]:
            test[prompt][video]["bounds"] = moments[prompt][video]["bounds"]
            test[prompt][video]["steps"] = [{"index": i, "heading": "", "absolute_bounds": [i, i + 1]} for i in range(5)]
    # --- moment segmentation, run.py:427-452
    items = build_items(test, feats, "moment_segmentation")
    moments = {}
    for i in range(0, len(items), batch_size):
        chunk = items[i:i + batch_size]
        b = collate(chunk)
        pred = mo.test_moment_segmentation(sd, b, b["text_feat"], threshold=ms_threshold, max_iterations=ms_iterations)
        for d, raw in zip(chunk, pred):
            bounds = [[f2t(raw[j], d["video_duration"], n_frames=-1), f2t(raw[j + 1], d["video_duration"], n_frames=-1)]
                      for j in range(len(raw) - 1)]
            moments[d["fname"]] = {"bounds": bounds, "video_duration": d["video_duration"], "pred_bounds": raw}
    ms = moments
    for prompt in test:
        for video in test[prompt]:
            test[prompt][video]["steps"] = []
            if video not in moments:
                continue
            for i, bound in enumerate(moments[video]["bounds"]):
                test[prompt][video]["steps"].append({"index": i, "heading": "", "absolute_bounds": bound})
    # --- step captioning, run.py:462-474
    items = build_items(test, feats, "step_captioning")
    moments = {}
    for i in range(0, len(items), batch_size):
        chunk = items[i:i + batch_size]
        b = collate(chunk)
        ids, _ = co.test_step_captioning(sd, b, b["text_feat"], beam=num_beams)
        for d, seq in zip(chunk, ids):
            moments.setdefault(d["fname"], {"captions": []})["captions"].append({"sentence": co.ids_to_text(seq, vocab)})
    sc = moments
    for prompt in test:
        for video in test[prompt]:
            if video in moments:
                for i, sent in enumerate(moments[video]["captions"]):
                    test[prompt][video]["steps"][i]["heading"] = sent["sentence"]
    return {"final": test, "moment_retrieval": mr, "moment_segmentation": ms, "step_captioning": sc}
