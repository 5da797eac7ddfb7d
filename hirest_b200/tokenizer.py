"""CLIP byte-pair tokenizer (host side) — produces the ``clip_text_ids`` the text tower consumes.

Reference: ``clip.tokenize(prompts)`` in ``hirest_dataset.py:528`` / ``inference_video_retrieval.py:203-206`` (the pip
``clip`` package; the reference vendors an equivalent at ``EVA_clip/clip.py:196-232`` + ``EVA_clip/simple_tokenizer.py``).
Neither ``clip`` nor ``ftfy`` is a dependency here.  Behaviour reproduced:

* text clean-up: (``ftfy.fix_text`` when ftfy is installed) → ``html.unescape`` twice → strip → runs of whitespace to one
  space → lower-case (``simple_tokenizer.py:47-57,118``);
* pre-tokenisation with the CLIP pattern (special tokens, English contractions, letter runs, single digits, punctuation runs);
* bytes → printable code points (the GPT-2 table), then greedy lowest-rank-first pair merging with an end-of-word marker on
  the last symbol; merges are lines 1 .. 48894 of the BPE file (``simple_tokenizer.py:64-66``);
* vocabulary order: 256 byte symbols, the same with ``</w>``, one entry per merge, ``<|startoftext|>``, ``<|endoftext|>``
  → ids 49406 / 49407;
* ``tokenize``: ``[SOT] + ids + [EOT]`` zero-padded to ``context_length`` (77); longer inputs raise unless ``truncate``.

The merge table is data the user supplies (``bpe_simple_vocab_16e6.txt.gz`` ships with CLIP and with the reference under
``EVA_clip/``); pass its path or set ``HIREST_BPE_PATH``.
"""
from __future__ import annotations

import gzip
import html
import os
from typing import Dict, Iterable, List, Sequence, Tuple, Union

import torch

try:
    import regex as _re
    _PATTERN = _re.compile(r"<\|startoftext\|>|<\|endoftext\|>|'s|'t|'re|'ve|'m|'ll|'d|[\p{L}]+|[\p{N}]|[^\s\p{L}\p{N}]+", _re.IGNORECASE)
except ImportError as e:  # pragma: no cover - the image ships `regex`
    raise ImportError("hirest_b200.tokenizer needs the `regex` package (Unicode property classes)") from e

try:
    import ftfy as _ftfy
except ImportError:
    _ftfy = None

SOT, EOT = "<|startoftext|>", "<|endoftext|>"
N_MERGES = 49152 - 256 - 2   # 48894


def _byte_symbols() -> List[str]:
    """One printable character per byte value: printable Latin-1 bytes map to themselves, the other 68 to U+0100 onwards."""
    keep = set(range(0x21, 0x7F)) | set(range(0xA1, 0xAD)) | set(range(0xAE, 0x100))
    table, extra = [""] * 256, 0
    for b in range(256):
        if b in keep:
            table[b] = chr(b)
        else:
            table[b] = chr(256 + extra)
            extra += 1
    return table


def _vocab_order(byte_syms: Sequence[str]) -> List[str]:
    """The reference lists the self-mapped bytes first (in byte order), then the remapped ones (simple_tokenizer.py:28-37)."""
    direct = [s for b, s in enumerate(byte_syms) if ord(s) == b]
    remapped = [s for b, s in enumerate(byte_syms) if ord(s) != b]
    return direct + remapped


class ClipBpeTokenizer:
    def __init__(self, bpe_path: str = None):
        bpe_path = bpe_path or os.environ.get("HIREST_BPE_PATH")
        if not bpe_path or not os.path.exists(bpe_path):
            raise FileNotFoundError("CLIP BPE merge table not found: pass bpe_path or set HIREST_BPE_PATH "
                                    "(bpe_simple_vocab_16e6.txt.gz, shipped with CLIP and under the reference's EVA_clip/)")
        with gzip.open(bpe_path, "rt", encoding="utf-8") as f:
            lines = f.read().split("\n")
        merges: List[Tuple[str, str]] = [tuple(ln.split()) for ln in lines[1:1 + N_MERGES]]
        self._byte = _byte_symbols()
        base = _vocab_order(self._byte)
        vocab = base + [s + "</w>" for s in base] + ["".join(m) for m in merges] + [SOT, EOT]
        self.encoder: Dict[str, int] = {tok: i for i, tok in enumerate(vocab)}
        self.decoder: Dict[int, str] = {i: tok for tok, i in self.encoder.items()}
        self._rank: Dict[Tuple[str, str], int] = {m: r for r, m in enumerate(merges)}
        self._unbyte = {s: b for b, s in enumerate(self._byte)}
        self._memo: Dict[str, List[int]] = {}
        self.sot_token, self.eot_token = self.encoder[SOT], self.encoder[EOT]
        self._merge_lines = "\n".join(" ".join(m) for m in merges)
        self._native = False   # built on first use (csrc/hb_tokenize.cu: ASCII fast path of the same algorithm)

    def _native_handle(self):
        if self._native is False:
            self._native = None
            try:
                import ctypes as C

                from . import _lib

                lib = _lib.load()
                blob = self._merge_lines.encode("utf-8")
                handle = C.c_void_p()
                if lib.hb_bpe_create(blob, len(blob), C.byref(handle)) == 0:
                    self._native = (lib, handle)
            except (OSError, RuntimeError, AttributeError):
                self._native = None   # library not built: the Python implementation is complete on its own
        return self._native

    def __del__(self):
        h = getattr(self, "_native", None)
        if h:
            try:
                h[0].hb_bpe_destroy(h[1])
            except Exception:  # noqa: BLE001  (interpreter shutdown)
                pass

    # ------------------------------------------------------------------ BPE
    def _merge_word(self, symbols: List[str]) -> List[str]:
        """Repeatedly fuse the adjacent pair with the lowest merge rank (every occurrence, left to right) until none is ranked."""
        while len(symbols) > 1:
            best, best_rank = None, None
            for pair in zip(symbols, symbols[1:]):
                r = self._rank.get(pair)
                if r is not None and (best_rank is None or r < best_rank):
                    best, best_rank = pair, r
            if best is None:
                break
            fused, out, i = best[0] + best[1], [], 0
            while i < len(symbols):
                if i + 1 < len(symbols) and symbols[i] == best[0] and symbols[i + 1] == best[1]:
                    out.append(fused)
                    i += 2
                else:
                    out.append(symbols[i])
                    i += 1
            symbols = out
        return symbols

    def _encode_piece(self, piece: str) -> List[int]:
        ids = self._memo.get(piece)
        if ids is None:
            if piece in (SOT, EOT):
                ids = [self.encoder[piece]]
            else:
                syms = [self._byte[b] for b in piece.encode("utf-8")]
                syms[-1] += "</w>"
                ids = [self.encoder[s] for s in self._merge_word(syms)]
            self._memo[piece] = ids
        return ids

    # ------------------------------------------------------------------ public
    @staticmethod
    def clean(text: str) -> str:
        if _ftfy is not None:
            text = _ftfy.fix_text(text)
        text = html.unescape(html.unescape(text)).strip()
        return _re.sub(r"\s+", " ", text).strip().lower()

    def encode(self, text: str) -> List[int]:
        out: List[int] = []
        for piece in _PATTERN.findall(self.clean(text)):
            out.extend(self._encode_piece(piece))
        return out

    def decode(self, ids: Iterable[int]) -> str:
        text = "".join(self.decoder[int(i)] for i in ids)
        chunks, buf = [], bytearray()
        i = 0
        while i < len(text):   # byte symbols back to bytes; "</w>" marks a word end
            if text.startswith("</w>", i):
                buf.append(0x20)
                i += 4
            elif text.startswith(SOT, i) or text.startswith(EOT, i):
                tok = SOT if text.startswith(SOT, i) else EOT
                chunks.append(buf.decode("utf-8", errors="replace") + tok)
                buf = bytearray()
                i += len(tok)
            else:
                buf.append(self._unbyte[text[i]])
                i += 1
        chunks.append(buf.decode("utf-8", errors="replace"))
        return "".join(chunks)

    def tokenize(self, texts: Union[str, Sequence[str]], context_length: int = 77, truncate: bool = False,
                 native: bool = True) -> torch.LongTensor:
        """``clip.tokenize``: ``[len(texts), context_length]`` int64, ``[SOT] ids [EOT] 0 0 …``.  ASCII prompts are tokenised by the
        library's native batch tokeniser in one call when it is available (and ftfy, which could rewrite them, is not installed);
        prompts it flags (non-ASCII, ``&`` entities, control characters) go through the Python code of this class."""
        if isinstance(texts, str):
            texts = [texts]
        n = len(texts)
        out = torch.zeros((n, context_length), dtype=torch.long)
        todo = range(n)
        h = self._native_handle() if (native and n > 0 and context_length >= 2 and _ftfy is None) else None
        if h is not None:
            import ctypes as C

            import numpy as np

            lib, handle = h
            arr = (C.c_char_p * n)(*[t.encode("utf-8") if "\0" not in t else None for t in texts])
            status = np.ones(n, dtype=np.uint8)
            if lib.hb_bpe_tokenize(handle, arr, n, int(context_length), int(bool(truncate)), out.data_ptr(), status.ctypes.data) == 0:
                for i in np.nonzero(status == 2)[0].tolist():
                    raise RuntimeError(f"Input {texts[i]} is too long for context length {context_length}")
                todo = np.nonzero(status == 1)[0].tolist()
        for i in todo:
            text = texts[i]
            ids = [self.sot_token] + self.encode(text) + [self.eot_token]
            if len(ids) > context_length:
                if not truncate:
                    raise RuntimeError(f"Input {text} is too long for context length {context_length}")
                ids = ids[:context_length]
                ids[-1] = self.eot_token
            out[i, :len(ids)] = torch.tensor(ids, dtype=torch.long)
        return out


_cache: Dict[str, ClipBpeTokenizer] = {}


def get_tokenizer(bpe_path: str = None) -> ClipBpeTokenizer:
    """One tokenizer per merge-table path per process (building one reads the gzip and indexes 48k merges)."""
    key = os.path.abspath(bpe_path or os.environ.get("HIREST_BPE_PATH") or "")
    tok = _cache.get(key)
    if tok is None:
        tok = _cache[key] = ClipBpeTokenizer(bpe_path)
    return tok


def tokenize(texts: Union[str, Sequence[str]], context_length: int = 77, truncate: bool = False, bpe_path: str = None) -> torch.LongTensor:
    """Module-level ``clip.tokenize`` replacement."""
    return get_tokenizer(bpe_path).tokenize(texts, context_length, truncate)
