"""The library's native batch tokenisers (csrc/hb_tokenize.cu, ASCII fast path) against the Python
implementations — which are themselves pinned on the reference's tokenisers (tests/golden/wordpiece.json, tokenizer.json) — on the
goldens and on random ASCII text; non-ASCII rows must be flagged and filled by the Python path.  Host only, no GPU."""
import json
import os
import random

import numpy as np
import pytest
import torch

from hirest_b200 import _lib, tokenizer, wordpiece

ALPHABET = "abcdefghijklmnopqrstuvwxyzABCDEFGHIJ0123456789      \t\n.,;:!?'\"()[]{}<>|-_/\\@#$%^*+=~`"


def _random_texts(n, seed, extra=""):
    rng = random.Random(seed)
    out = []
    for _ in range(n):
        k = rng.randint(0, 90)
        s = "".join(rng.choice(ALPHABET + extra) for _ in range(k))
        if rng.random() < 0.3:
            s += rng.choice([" it's", " don't", " we're", " i've", " i'm", " they'll", " he'd", " rock'n'roll", "[CLS]", " [SEP] ", "[MASK].",
                             " <|startoftext|>", "<|endoftext|> "])
        out.append(s)
    return out


@pytest.fixture(scope="module")
def lib():
    return _lib.load()


def test_wordpiece_native_equals_python_on_golden_and_random(lib, golden_dir):
    with open(os.path.join(golden_dir, "wordpiece.json"), encoding="utf-8") as f:
        g = json.load(f)
    tok = wordpiece.WordPieceTokenizer(g["vocab"])
    assert tok._native_handle() is not None, "libhirest_b200.so must export the native tokeniser"
    caps = list(g["captions"]) + _random_texts(400, 3) + ["", " ", "a" * 101, "x" * 100, "café crème", "中文 mix", "tab\x07bell"]
    for mw in (48, 8, 2):
        nat = tok.encode_captions(caps, max_words=mw)
        ref = tok.encode_captions(caps, max_words=mw, native=False)
        for a, b in zip(nat, ref):
            assert a.dtype == np.int64 and np.array_equal(a, b), mw
    a, b, m = tok.encode_captions(g["captions"], max_words=48)
    assert a.tolist() == g["input_ids"] and b.tolist() == g["output_ids"] and m.tolist() == g["decoder_mask"]


def test_wordpiece_native_flags_non_ascii(lib, golden_dir):
    import ctypes as C

    with open(os.path.join(golden_dir, "wordpiece.json"), encoding="utf-8") as f:
        g = json.load(f)
    tok = wordpiece.WordPieceTokenizer(g["vocab"])
    h = tok._native_handle()
    caps = ["plain ascii", "naïve", "ctrl\x01char", "del\x7f"]
    arr = (C.c_char_p * len(caps))(*[c.encode("utf-8") for c in caps])
    bufs = [np.full((len(caps), 6), -7, dtype=np.int64) for _ in range(3)]
    flags = np.zeros(len(caps), dtype=np.uint8)
    assert h[0].hb_wordpiece_encode_captions(h[1], arr, len(caps), 6, bufs[0].ctypes.data, bufs[1].ctypes.data, bufs[2].ctypes.data,
                                             flags.ctypes.data) == 0
    assert flags.tolist() == [0, 1, 1, 1]
    assert (bufs[0][1:] == -7).all() and (bufs[0][0] != -7).all()   # flagged rows are left to the caller


def test_bpe_native_equals_python(lib, golden_dir):
    path = os.path.join(golden_dir, "bpe_synthetic.txt.gz")
    tok = tokenizer.ClipBpeTokenizer(path)
    if tokenizer._ftfy is not None:
        pytest.skip("ftfy installed: the native path is disabled")
    assert tok._native_handle() is not None
    texts = _random_texts(500, 11) + ["", "   ", "Tom &amp; Jerry", "café", "a photo of a cat", "  MIXED   Case\tTabs\nnewlines  ",
                                      "\U0001f600 emoji", "semi;colon's", "x" * 60]
    for ctx in (77, 16):
        nat = tok.tokenize(texts, context_length=ctx, truncate=True)
        ref = tok.tokenize(texts, context_length=ctx, truncate=True, native=False)
        assert nat.dtype == torch.long and torch.equal(nat, ref), ctx
    with pytest.raises(RuntimeError, match="too long"):
        tok.tokenize(["word " * 100], context_length=77)
    with pytest.raises(RuntimeError, match="too long"):
        tok.tokenize(["word " * 100], context_length=77, native=False)


def test_bpe_native_on_reference_golden_prompts(lib, golden_dir):
    """The 16 prompts whose ids were produced by the reference's SimpleTokenizer (tests/golden/tokenizer.json) need the real CLIP merge
    table, which is user data: runs where HIREST_BPE_PATH (or the reference checkout) provides it."""
    path = os.environ.get("HIREST_BPE_PATH") or "/root/reference/EVA_clip/bpe_simple_vocab_16e6.txt.gz"
    if not os.path.exists(path):
        pytest.skip("CLIP merge table not available")
    with open(os.path.join(golden_dir, "tokenizer.json"), encoding="utf-8") as f:
        g = json.load(f)
    if tokenizer._ftfy is not None:
        pytest.skip("ftfy installed: the native path is disabled")
    tok = tokenizer.get_tokenizer(path)
    prompts = g["prompts"] if "prompts" in g else [c["text"] for c in g["cases"]]
    nat = tok.tokenize(prompts, truncate=True)
    ref = tok.tokenize(prompts, truncate=True, native=False)
    assert torch.equal(nat, ref)
    assert tok.sot_token == 49406 and tok.eot_token == 49407
