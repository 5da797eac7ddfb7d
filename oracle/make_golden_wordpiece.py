"""Generates tests/golden/wordpiece.json by running the reference's OWN BertTokenizer (clip4caption/modules/tokenization.py) and
MomentDataset.clip4cap_get_text (hirest_dataset.py:533-580) on a synthetic WordPiece vocabulary and a fixed list of captions.
Build container only:    python oracle/make_golden_wordpiece.py

The real bert-base-uncased vocab.txt is a download (not available offline); the vocabulary below has the real special-token ids
([PAD] 0, [UNK] 100, [CLS] 101, [SEP] 102, [MASK] 103) and enough whole words / '##' continuations / single characters to exercise
greedy longest-match, [UNK] fallbacks, accents, CJK isolation, punctuation splitting and the 100-character word limit.
numpy >= 1.24 has no np.long (the reference still uses it): it is aliased to np.int64 for the clip4cap_get_text call."""
import importlib.util
import json
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"

WORDS = ("the a an and of to in on with for into from until then it is are be add cut chop slice dice mix stir pour heat boil fry bake "
         "whisk fold knead roll season drain rinse peel grate serve place put remove let rest cool onion onions garlic butter flour sugar "
         "salt pepper water oil egg eggs milk cream cheese chicken beef rice pasta sauce pan pot bowl oven knife board minutes minute "
         "degrees cup cups tablespoon teaspoon how make change replace install tire bike wheel screw bolt step first next finally "
         "cafe creme brulee jalapeno naive un aff able ed er est ly super win do dows").split()
SUFFIXES = "s ing ed er est ly able aff tion ment ness ful less ize ise al ic ous ive en ry y e es d n t r l".split()
CAPTIONS = [
    "add the chopped onions and garlic to the pan",
    "Cut the ONIONS; then FRY them (5-10 minutes)!",
    "crème brûlée & jalapeño — naïve café",
    "unaffable unaffordable supercalifragilistic",
    "step 12: bake at 350 degrees for 25-30 minutes",
    "中文 recipe: 炒饭 with eggs",
    "  leading,   trailing\tand\nmixed   whitespace  ",
    "[CLS] literal special [SEP] tokens [MASK] [PAD] [UNK] [cls]",
    "it's the chef's knife, isn't it?",
    "x" * 101 + " short " + "y" * 100,
    "emoji 🍕 and control \x00\x07 chars� here",
    "",
    "a",
    "install windows 11 on a 2tb ssd (uefi/gpt)",
    "whisk " * 60,
    "Ünïcödé ÀÉÎÕÜ straße ǅ ﬁ",
    "semi-colon;colon:dash-underscore_caret^dollar$backtick`tilde~",
]


def make_vocab():
    v = ["[PAD]"] + [f"[unused{i}]" for i in range(1, 100)] + ["[UNK]", "[CLS]", "[SEP]", "[MASK]"]
    chars = list("abcdefghijklmnopqrstuvwxyz0123456789") + list("!\"#$%&'()*+,-./:;<=>?@[\\]^_`{|}~") + ["—", "中", "文", "炒", "ß", "ﬁ"]
    v += chars + ["##" + c for c in "abcdefghijklmnopqrstuvwxyz0123456789"]
    seen = set(v)
    for w in WORDS + ["##" + s for s in SUFFIXES]:
        if w not in seen:
            v.append(w)
            seen.add(w)
    return v


def main():
    vocab = make_vocab()
    tmp = "/tmp/hb_wordpiece_vocab.txt"
    with open(tmp, "w", encoding="utf-8") as f:
        f.write("\n".join(vocab) + "\n")
    sys.path.insert(0, ROOT)
    from oracle import ref_moment

    ref_moment.install_stubs()   # boto3 / botocore (file_utils.py:20) carry no tokenizer logic
    sys.path.insert(0, os.path.join(REF, "clip4caption"))
    from modules.tokenization import BertTokenizer  # the reference's class

    tok = BertTokenizer(tmp, do_lower_case=True, max_len=512, never_split=("[UNK]", "[SEP]", "[PAD]", "[CLS]", "[MASK]"))
    # clip4cap_get_text, unbound, on a stand-in `self` (the method only reads self.args.max_words and self.tokenizer)
    for name in ("srt", "clip"):
        sys.modules.setdefault(name, types.ModuleType(name))
    spec = importlib.util.spec_from_file_location("ref_hirest_dataset", os.path.join(REF, "hirest_dataset.py"))
    ds = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ds)
    if not hasattr(np, "long"):
        np.long = np.int64
    fake = types.SimpleNamespace(args=types.SimpleNamespace(max_words=48), tokenizer=tok)
    out = {"note": "reference clip4caption/modules/tokenization.py BertTokenizer + hirest_dataset.py clip4cap_get_text; synthetic vocab",
           "vocab": vocab, "captions": CAPTIONS, "tokens": [], "ids": [], "input_ids": [], "output_ids": [], "decoder_mask": []}
    for c in CAPTIONS:
        t = tok.tokenize(c)
        out["tokens"].append(t)
        out["ids"].append(tok.convert_tokens_to_ids(t))
        r = ds.MomentDataset.clip4cap_get_text(fake, c)
        out["input_ids"].append(r[5][0].tolist())
        out["decoder_mask"].append(r[6][0].tolist())
        out["output_ids"].append(r[7][0].tolist())
    with open(os.path.join(ROOT, "tests", "golden", "wordpiece.json"), "w", encoding="utf-8") as f:
        json.dump(out, f, ensure_ascii=False, indent=0)
    print(len(CAPTIONS), "captions;", sum(len(t) for t in out["tokens"]), "pieces; vocab", len(vocab))


if __name__ == "__main__":
    main()
