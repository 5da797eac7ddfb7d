"""Generates tests/golden/tokenizer.json by running the reference's OWN tokenizer (EVA_clip/simple_tokenizer.py +
EVA_clip/clip.py:196-232 `tokenize`) on a fixed list of prompts.  Run in the build container:

    python oracle/make_golden_tokenizer.py

ftfy is not installed there; it is shimmed with the identity (`fix_text` only repairs mojibake, the prompts below contain
none), which the header of the fixture records."""
import importlib.util
import json
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/EVA_clip"

PROMPTS = [
    "how to make a paper airplane",
    "How to Change a Flat Tire on a Bicycle?",
    "make   iced   coffee\tat home\n",
    "it's the chef's knife, isn't it? they've said we'll & I'd",
    "step 12: bake at 350 degrees for 25-30 minutes!!!",
    "crème brûlée &amp; jalapeño poppers — naïve café",
    "install windows 11 on a 2TB ssd (uefi/gpt)",
    "日本語のレシピ: お好み焼き",
    "emoji 🍕🍔 test",
    "&lt;b&gt;bold&lt;/b&gt; &amp;amp; html",
    "a",
    "",
    "supercalifragilisticexpialidocious antidisestablishmentarianism pneumonoultramicroscopicsilicovolcanoconiosis",
    "<|startoftext|> literal special tokens <|endoftext|>",
    "MiXeD CaSe WITH NUMBERS 1234567890 and_underscores #hashtag @mention",
]
LONG = "repair the kitchen sink faucet " * 20


def train_synthetic_merges(n_merges=600):
    """A small merge table that TRAVELS with the repo (the real one is CLIP's data file): plain BPE training over the prompts
    above, in the reference's symbol alphabet (bytes -> printable code points, '</w>' on the last symbol of a word)."""
    import collections
    import html

    import regex

    sys.path.insert(0, ROOT)
    from hirest_b200 import tokenizer as ours   # only its byte table / pre-tokenisation pattern (data, not the algorithm under test)

    byte_syms = ours._byte_symbols()
    words = collections.Counter()
    for p in PROMPTS[:8] + [LONG]:   # words of the other prompts are only PARTLY covered by the merges
        text = regex.sub(r"\s+", " ", html.unescape(html.unescape(p)).strip()).strip().lower()
        for piece in ours._PATTERN.findall(text):
            if piece in (ours.SOT, ours.EOT):
                continue
            syms = [byte_syms[b] for b in piece.encode("utf-8")]
            syms[-1] += "</w>"
            words[tuple(syms)] += 1
    merges = []
    for _ in range(n_merges):
        pairs = collections.Counter()
        for w, c in words.items():
            for a, b in zip(w, w[1:]):
                pairs[(a, b)] += c
        if not pairs:
            break
        best = max(sorted(pairs), key=lambda k: pairs[k])
        merges.append(best)
        new_words = collections.Counter()
        for w, c in words.items():
            out, i = [], 0
            while i < len(w):
                if i + 1 < len(w) and (w[i], w[i + 1]) == best:
                    out.append(w[i] + w[i + 1])
                    i += 2
                else:
                    out.append(w[i])
                    i += 1
            new_words[tuple(out)] += c
        words = new_words
    return merges


def main():
    sys.modules["ftfy"] = types.SimpleNamespace(fix_text=lambda t: t)
    spec = importlib.util.spec_from_file_location("ref_simple_tokenizer", os.path.join(REF, "simple_tokenizer.py"))
    st = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(st)
    tok = st.SimpleTokenizer(os.path.join(REF, "bpe_simple_vocab_16e6.txt.gz"))
    sot, eot = tok.encoder["<|startoftext|>"], tok.encoder["<|endoftext|>"]
    out = {"note": "reference EVA_clip/simple_tokenizer.py, ftfy shimmed with identity", "sot": sot, "eot": eot,
           "vocab_size": len(tok.encoder), "prompts": PROMPTS, "ids": [tok.encode(p) for p in PROMPTS],
           "long_prompt": LONG, "long_ids": tok.encode(LONG),
           "decoded": [tok.decode(tok.encode(p)) for p in PROMPTS]}
    # the same prompts through the reference tokenizer on the synthetic merge table committed next to the golden
    import gzip

    merges = train_synthetic_merges()
    syn_path = os.path.join(ROOT, "tests", "golden", "bpe_synthetic.txt.gz")
    with gzip.open(syn_path, "wt", encoding="utf-8") as f:
        f.write("#version: synthetic (oracle/make_golden_tokenizer.py)\n" + "\n".join(" ".join(m) for m in merges))
    tok2 = st.SimpleTokenizer(syn_path)
    out["synthetic"] = {"n_merges": len(merges), "vocab_size": len(tok2.encoder), "sot": tok2.encoder["<|startoftext|>"],
                        "eot": tok2.encoder["<|endoftext|>"], "ids": [tok2.encode(p) for p in PROMPTS], "long_ids": tok2.encode(LONG),
                        "decoded": [tok2.decode(tok2.encode(p)) for p in PROMPTS]}
    with open(os.path.join(ROOT, "tests", "golden", "tokenizer.json"), "w") as f:
        json.dump(out, f, ensure_ascii=False, indent=0)
    print(len(PROMPTS), "prompts;", sum(len(x) for x in out["ids"]), "tokens;", len(merges), "synthetic merges ->",
          sum(len(x) for x in out["synthetic"]["ids"]), "tokens")


if __name__ == "__main__":
    main()
