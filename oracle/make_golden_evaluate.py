"""Generates tests/golden/evaluate_retrieval.json by running the reference's OWN evaluate.evaluate_video_retrieval (evaluate.py:33-81)
on a synthetic ground truth / prediction pair with many exactly tied scores.  Build container only:
    python oracle/make_golden_evaluate.py
`language_evaluation` (caption metrics, not on this path) is shimmed with an empty module; the module-level globals
PROMPT_TO_CAT / PROMPT_CATEGORIES that evaluate.py's __main__ builds from data/evaluation/categories.json are set by hand."""
import importlib.util
import json
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    sys.modules["language_evaluation"] = types.ModuleType("language_evaluation")
    spec = importlib.util.spec_from_file_location("ref_evaluate", "/root/reference/evaluate.py")
    ev = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ev)
    rng = np.random.default_rng(0)
    V, Q = 70, 24
    names = [f"vid{int(i):03d}.mp4" for i in rng.permutation(V)]
    scores = np.round(rng.normal(size=(Q, V)), 1)          # one decimal: many exact ties, broken by the name
    prompts = [f"prompt {q}" for q in range(Q)]
    gt = {p: {names[int(j)]: {} for j in rng.choice(V, size=int(rng.integers(1, 4)), replace=False)} for p in prompts}
    pred = {p: {"videos": names, "scores": scores[q].tolist()} for q, p in enumerate(prompts)}
    ev.PROMPT_TO_CAT = {p: ("cooking" if q % 2 else "repair") for q, p in enumerate(prompts)}
    ev.PROMPT_CATEGORIES = ["cooking", "repair", "all"]
    ev.tqdm = lambda x: x
    res = ev.evaluate_video_retrieval(gt, pred)
    # the reference's ranking itself, per prompt (evaluate.py:58-60)
    ranked = []
    for p in prompts:
        sc, vs = zip(*sorted(zip(pred[p]["scores"], pred[p]["videos"])))
        ranked.append(list(vs[::-1][:50]))
    out = {"video_names": names, "scores": scores.tolist(), "prompts": prompts, "gt": {p: list(v) for p, v in gt.items()},
           "categories": ev.PROMPT_TO_CAT, "results": res, "ranked_top50": ranked}
    with open(os.path.join(ROOT, "tests", "golden", "evaluate_retrieval.json"), "w") as f:
        json.dump(out, f)
    print(res)


if __name__ == "__main__":
    main()
