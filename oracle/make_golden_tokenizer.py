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
    with open(os.path.join(ROOT, "tests", "golden", "tokenizer.json"), "w") as f:
        json.dump(out, f, ensure_ascii=False, indent=0)
    print(len(PROMPTS), "prompts;", sum(len(x) for x in out["ids"]), "tokens")


if __name__ == "__main__":
    main()
