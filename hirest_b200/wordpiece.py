"""BERT WordPiece tokenizer for the caption side of the path.

Reference: ``clip4caption/modules/tokenization.py`` (``BertTokenizer`` = basic tokenizer + greedy longest-match WordPiece, lower-cased,
``never_split`` = the five special tokens) as used by ``hirest_dataset.py:119-121`` and ``clip4cap_get_text`` (``:533-580``: caption →
``[CLS]`` + pieces / pieces + ``[SEP]`` padded to ``max_words``), and by ``modeling.py:615-626`` to turn generated ids back into text.
The vocabulary file (``bert-base-uncased/vocab.txt``, one token per line) is user-supplied data, like the CLIP merge table.

Same outputs, different construction: one pass classifies every character once (drop / space / punctuation / CJK / word), the
word-piece search is bounded by the longest vocabulary entry instead of starting from the whole remainder, and batches can be
encoded straight into padded id arrays.  Pinned against the reference class on the strings of ``tests/golden/wordpiece.json``
(``oracle/make_golden_wordpiece.py``).
"""
from __future__ import annotations

import unicodedata
from typing import Dict, Iterable, List, Sequence, Tuple

import numpy as np

SPECIAL_TOKENS = ("[UNK]", "[SEP]", "[PAD]", "[CLS]", "[MASK]")
_ASCII_PUNCT = frozenset(chr(c) for r in ((33, 48), (58, 65), (91, 97), (123, 127)) for c in range(*r))
_CJK_RANGES = ((0x4E00, 0x9FFF), (0x3400, 0x4DBF), (0x20000, 0x2A6DF), (0x2A700, 0x2B73F), (0x2B740, 0x2B81F), (0x2B820, 0x2CEAF),
               (0xF900, 0xFAFF), (0x2F800, 0x2FA1F))
_DROP, _SPACE, _PUNCT, _CJK, _WORD = range(5)


def _char_class(ch: str) -> int:
    """Input cleaning (tokenization.py `_clean_text`, `_is_whitespace`, `_is_control`), CJK isolation and punctuation classes."""
    if ch in " \t\n\r":
        return _SPACE
    cp = ord(ch)
    if cp == 0 or cp == 0xFFFD:
        return _DROP
    cat = unicodedata.category(ch)
    if cat[0] == "C":
        return _DROP
    if cat == "Zs":
        return _SPACE
    if any(lo <= cp <= hi for lo, hi in _CJK_RANGES):
        return _CJK
    if ch in _ASCII_PUNCT or cat[0] == "P":
        return _PUNCT
    return _WORD


def _is_punct(ch: str) -> bool:
    return ch in _ASCII_PUNCT or unicodedata.category(ch)[0] == "P"


def load_vocab(path: str) -> Dict[str, int]:
    """One token per line, index = line number; a later duplicate overwrites an earlier one (as the reference's OrderedDict does)."""
    vocab: Dict[str, int] = {}
    with open(path, "r", encoding="utf-8") as f:
        for i, line in enumerate(f):
            vocab[line.strip()] = i
    return vocab


class WordPieceTokenizer:
    def __init__(self, vocab, do_lower_case: bool = True, never_split: Sequence[str] = SPECIAL_TOKENS, unk_token: str = "[UNK]",
                 max_chars_per_word: int = 100):
        self.vocab: Dict[str, int] = load_vocab(vocab) if isinstance(vocab, str) else (
            dict(vocab) if isinstance(vocab, dict) else {t: i for i, t in enumerate(vocab)})
        self.ids_to_tokens = {i: t for t, i in self.vocab.items()}
        self.do_lower_case = do_lower_case
        self.never_split = frozenset(never_split)
        self.unk_token = unk_token
        self.max_chars_per_word = max_chars_per_word
        self._longest = max((len(t) for t in self.vocab), default=1)

    # ------------------------------------------------------------------ basic tokenizer
    def _words(self, text: str) -> Iterable[str]:
        """Whitespace-delimited tokens after cleaning / CJK isolation, each lower-cased + accent-stripped (unless it is a special
        token) and split at punctuation."""
        buf: List[str] = []
        raw: List[str] = []
        for ch in text:
            k = _char_class(ch)
            if k == _DROP:
                continue
            if k == _SPACE:
                if buf:
                    raw.append("".join(buf))
                    buf = []
            elif k == _CJK:
                if buf:
                    raw.append("".join(buf))
                    buf = []
                raw.append(ch)
            else:
                buf.append(ch)
        if buf:
            raw.append("".join(buf))
        for tok in raw:
            if self.do_lower_case and tok not in self.never_split:
                tok = "".join(c for c in unicodedata.normalize("NFD", tok.lower()) if unicodedata.category(c) != "Mn")
            if tok in self.never_split:
                yield tok
                continue
            start = None
            for i, c in enumerate(tok):
                if _is_punct(c):
                    if start is not None:
                        yield tok[start:i]
                        start = None
                    yield c
                elif start is None:
                    start = i
            if start is not None:
                yield tok[start:]

    # ------------------------------------------------------------------ word pieces
    def _pieces(self, word: str) -> List[str]:
        n = len(word)
        if n > self.max_chars_per_word:
            return [self.unk_token]
        out: List[str] = []
        pos = 0
        while pos < n:
            prefix = "##" if pos else ""
            hit = None
            for end in range(min(n, pos + self._longest), pos, -1):   # longest match first
                cand = prefix + word[pos:end]
                if cand in self.vocab:
                    hit = (cand, end)
                    break
            if hit is None:
                return [self.unk_token]
            out.append(hit[0])
            pos = hit[1]
        return out

    def tokenize(self, text: str) -> List[str]:
        out: List[str] = []
        for word in self._words(text):
            for w in word.split():   # accent stripping can surface whitespace-like characters; the reference re-splits here
                out += self._pieces(w)
        return out

    def convert_tokens_to_ids(self, tokens: Sequence[str]) -> List[int]:
        unk = self.vocab[self.unk_token]
        return [self.vocab.get(t, unk) for t in tokens]

    def convert_ids_to_tokens(self, ids: Sequence[int]) -> List[str]:
        return [self.ids_to_tokens[int(i)] for i in ids]

    # ------------------------------------------------------------------ caption side of the dataset (hirest_dataset.py:533-580)
    def encode_caption(self, caption: str, max_words: int = 48) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
        """``clip4cap_get_text``: (decoder input ids ``[CLS] w1 ..``, decoder target ids ``w1 .. [SEP]``, decoder mask), each
        ``int64 [max_words]``, pieces truncated to ``max_words - 1``."""
        words = self.tokenize(caption)[:max_words - 1]
        inp = self.convert_tokens_to_ids(["[CLS]"] + words)
        tgt = self.convert_tokens_to_ids(words + ["[SEP]"])
        a, b, m = (np.zeros(max_words, dtype=np.int64) for _ in range(3))
        a[:len(inp)] = inp
        b[:len(tgt)] = tgt
        m[:len(inp)] = 1
        return a, b, m

    def _native_handle(self):
        """The library's batch tokeniser (csrc/hb_tokenize.cu: ASCII fast path) for the reference's configuration, or None."""
        h = getattr(self, "_native", False)
        if h is not False:
            return h
        self._native = None
        if (self.never_split == frozenset(SPECIAL_TOKENS) and self.unk_token == "[UNK]" and self.max_chars_per_word == 100
                and all(t in self.vocab for t in ("[UNK]", "[CLS]", "[SEP]"))):
            try:
                import ctypes as C

                from . import _lib

                lib = _lib.load()
                toks = [t for t in self.vocab if t and "\0" not in t]
                blob = b"".join(t.encode("utf-8") + b"\0" for t in toks)
                ids = np.asarray([self.vocab[t] for t in toks], dtype=np.int64)
                handle = C.c_void_p()
                if lib.hb_wordpiece_create(blob, len(blob), ids.ctypes.data, len(toks), int(self.do_lower_case), C.byref(handle)) == 0:
                    self._native = (lib, handle)
            except (OSError, RuntimeError, AttributeError):
                self._native = None   # library not built: the Python implementation below is complete on its own
        return self._native

    def __del__(self):
        h = getattr(self, "_native", None)
        if h:
            try:
                h[0].hb_wordpiece_destroy(h[1])
            except Exception:  # noqa: BLE001  (interpreter shutdown)
                pass

    def encode_captions(self, captions: Sequence[str], max_words: int = 48, native: bool = True):
        """Batch form: three ``int64 [n, max_words]`` arrays.  ASCII captions go through the library's native batch tokeniser in one
        call when it is available; captions it flags (non-ASCII, control characters) and everything else use the code above."""
        n = len(captions)
        out = tuple(np.zeros((n, max_words), dtype=np.int64) for _ in range(3))
        todo = range(n)
        h = self._native_handle() if (native and n > 0 and max_words >= 2) else None
        if h is not None:
            import ctypes as C

            lib, handle = h
            arr = (C.c_char_p * n)(*[c.encode("utf-8") if "\0" not in c else None for c in captions])
            flags = np.ones(n, dtype=np.uint8)
            if lib.hb_wordpiece_encode_captions(handle, arr, n, int(max_words), out[0].ctypes.data, out[1].ctypes.data, out[2].ctypes.data,
                                                flags.ctypes.data) == 0:
                todo = np.nonzero(flags)[0].tolist()
        for i in todo:
            a, b, m = self.encode_caption(captions[i], max_words)
            out[0][i], out[1][i], out[2][i] = a, b, m
        return out

    def decode(self, ids: Sequence[int]) -> str:
        """``modeling.py:615-626``: ids → tokens, cut at the first [SEP] / [PAD], join, merge ``##`` continuations."""
        toks = self.convert_ids_to_tokens(ids)
        for stop in ("[SEP]", "[PAD]"):
            if stop in toks:
                toks = toks[:toks.index(stop)]
        return str(" ".join(toks).replace(" ##", "").strip("##").strip())
