// hb_tokenize.cu — host-only batch tokenisers behind the C ABI: BERT WordPiece for the caption side
// (clip4caption/modules/tokenization.py via hirest_dataset.py:119-121, 533-580) and CLIP byte-pair encoding for the prompts
// (EVA_clip/simple_tokenizer.py, clip.tokenize at hirest_dataset.py:528 / inference_video_retrieval.py:203-206).
//
// Scope: the ASCII fast path.  Unicode normalisation (NFD accent stripping, category tables, ftfy, HTML entities) lives in Python's
// standard library and stays there: a text with a byte >= 0x80, an '&' (entities) or a control character is FLAGGED and the Python
// caller runs its full-Unicode path for that row (hirest_b200/wordpiece.py, tokenizer.py).  For ASCII input the two paths are
// identical by construction and are tested against each other and against the reference's own tokenisers' goldens.
// No CUDA in this file; it is compiled by nvcc only because the library has one build recipe.
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <memory>
#include <new>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/hirest_b200.h"

namespace {

inline bool is_space(unsigned char c) { return c == ' ' || c == '\t' || c == '\n' || c == '\r'; }
inline bool is_ascii_punct(unsigned char c) { return (c >= 33 && c <= 47) || (c >= 58 && c <= 64) || (c >= 91 && c <= 96) || (c >= 123 && c <= 126); }
inline bool is_alpha(unsigned char c) { return (c >= 'a' && c <= 'z') || (c >= 'A' && c <= 'Z'); }
inline bool is_digit(unsigned char c) { return c >= '0' && c <= '9'; }
inline unsigned char lower(unsigned char c) { return (c >= 'A' && c <= 'Z') ? static_cast<unsigned char>(c + 32) : c; }

// UTF-8 of one code point < 0x800 (all the byte-symbol table needs)
std::string utf8(unsigned cp) {
  std::string s;
  if (cp < 0x80) s.push_back(static_cast<char>(cp));
  else { s.push_back(static_cast<char>(0xC0 | (cp >> 6))); s.push_back(static_cast<char>(0x80 | (cp & 0x3F))); }
  return s;
}

}  // namespace

// =====================================================================================================================
// WordPiece
// =====================================================================================================================
struct HbWordPiece {
  std::unordered_map<std::string, int64_t> vocab;
  size_t longest = 1;
  bool lower_case = true;
  int64_t unk = 0, cls = 0, sep = 0;
  static bool never_split(const std::string& t) { return t == "[UNK]" || t == "[SEP]" || t == "[PAD]" || t == "[CLS]" || t == "[MASK]"; }

  // tokenization.py WordpieceTokenizer.tokenize: greedy longest match, "##" on continuations, whole word -> [UNK] on a miss
  void pieces(const std::string& w, std::vector<int64_t>& out) const {
    if (w.size() > 100) { out.push_back(unk); return; }
    const size_t first = out.size();
    size_t pos = 0;
    std::string cand;
    while (pos < w.size()) {
      bool hit = false;
      const size_t max_end = std::min(w.size(), pos + longest);
      for (size_t end = max_end; end > pos; --end) {
        cand.assign(pos ? "##" : "");
        cand.append(w, pos, end - pos);
        auto it = vocab.find(cand);
        if (it != vocab.end()) { out.push_back(it->second); pos = end; hit = true; break; }
      }
      if (!hit) { out.resize(first); out.push_back(unk); return; }
    }
  }

  // BasicTokenizer (clean -> whitespace split -> lower -> split at punctuation) + WordPiece on an ASCII text
  void encode(const char* text, std::vector<int64_t>& out) const {
    std::string raw, tok;
    auto flush = [&]() {
      if (raw.empty()) return;
      if (never_split(raw)) { pieces(raw, out); raw.clear(); return; }   // kept intact, not lower-cased
      if (lower_case) for (auto& c : raw) c = static_cast<char>(lower(static_cast<unsigned char>(c)));
      if (never_split(raw)) { pieces(raw, out); raw.clear(); return; }   // (cannot happen after lower-casing; mirrors the reference's second test)
      tok.clear();
      for (char ch : raw) {
        if (is_ascii_punct(static_cast<unsigned char>(ch))) {
          if (!tok.empty()) { pieces(tok, out); tok.clear(); }
          pieces(std::string(1, ch), out);
        } else {
          tok.push_back(ch);
        }
      }
      if (!tok.empty()) pieces(tok, out);
      raw.clear();
    };
    for (const unsigned char* p = reinterpret_cast<const unsigned char*>(text); *p; ++p) {
      if (is_space(*p)) flush();
      else raw.push_back(static_cast<char>(*p));   // control characters were excluded by the caller-visible ASCII test
    }
    flush();
  }
};

// pure ASCII without control characters other than tab / newline / carriage return (those are dropped or are Unicode business)
static bool plain_ascii(const char* s) {
  for (const unsigned char* p = reinterpret_cast<const unsigned char*>(s); *p; ++p)
    if (*p >= 0x7F || (*p < 0x20 && !is_space(*p))) return false;
  return true;
}

extern "C" {

int hb_wordpiece_create(const char* tokens, int64_t tokens_bytes, const int64_t* ids, int n_tokens, int do_lower_case, HbWordPiece** out) {
  if (!tokens || !ids || !out || n_tokens <= 0 || tokens_bytes <= 0) return HB_ERR_INVALID;
  std::unique_ptr<HbWordPiece> w(new (std::nothrow) HbWordPiece);
  if (!w) return HB_ERR_NOMEM;
  w->lower_case = do_lower_case != 0;
  const char* p = tokens;
  const char* end = tokens + tokens_bytes;
  for (int i = 0; i < n_tokens; ++i) {
    if (p >= end) return HB_ERR_INVALID;
    const size_t len = strnlen(p, static_cast<size_t>(end - p));
    if (p + len >= end) return HB_ERR_INVALID;   // every token is NUL-terminated inside the blob
    if (len > 0) {
      w->vocab[std::string(p, len)] = ids[i];    // a later duplicate overwrites an earlier one, like the reference's dict
      w->longest = std::max(w->longest, len);
    }
    p += len + 1;
  }
  auto need = [&](const char* t, int64_t& dst) { auto it = w->vocab.find(t); if (it == w->vocab.end()) return false; dst = it->second; return true; };
  if (!need("[UNK]", w->unk) || !need("[CLS]", w->cls) || !need("[SEP]", w->sep)) return HB_ERR_INVALID;
  *out = w.release();
  return HB_OK;
}

void hb_wordpiece_destroy(HbWordPiece* w) { delete w; }

int hb_wordpiece_encode_captions(const HbWordPiece* w, const char* const* captions, int n, int max_words, int64_t* input_ids,
                                 int64_t* target_ids, int64_t* mask, uint8_t* fallback) {
  if (!w || !captions || !input_ids || !target_ids || !mask || !fallback || n < 0 || max_words < 2) return HB_ERR_INVALID;
  std::vector<int64_t> ids;
  for (int i = 0; i < n; ++i) {
    int64_t* a = input_ids + static_cast<size_t>(i) * max_words;
    int64_t* b = target_ids + static_cast<size_t>(i) * max_words;
    int64_t* m = mask + static_cast<size_t>(i) * max_words;
    if (!captions[i] || !plain_ascii(captions[i])) { fallback[i] = 1; continue; }
    fallback[i] = 0;
    ids.clear();
    w->encode(captions[i], ids);
    const size_t nw = std::min(ids.size(), static_cast<size_t>(max_words - 1));   // clip4cap_get_text: pieces truncated to max_words - 1
    std::memset(a, 0, sizeof(int64_t) * max_words);
    std::memset(b, 0, sizeof(int64_t) * max_words);
    std::memset(m, 0, sizeof(int64_t) * max_words);
    a[0] = w->cls;
    for (size_t k = 0; k < nw; ++k) { a[k + 1] = ids[k]; b[k] = ids[k]; }
    b[nw] = w->sep;
    for (size_t k = 0; k <= nw; ++k) m[k] = 1;
  }
  return HB_OK;
}

}  // extern "C"

// =====================================================================================================================
// CLIP BPE
// =====================================================================================================================
struct HbBpe {
  std::unordered_map<std::string, int32_t> encoder;   // symbol string (UTF-8) -> id
  std::unordered_map<std::string, int32_t> rank;      // "a\x01b" -> merge rank
  std::unordered_map<std::string, std::vector<int32_t>> memo;
  std::string byte_sym[256];
  int32_t sot = 0, eot = 0;

  const std::vector<int32_t>& encode_piece(const std::string& piece) {
    auto it = memo.find(piece);
    if (it != memo.end()) return it->second;
    std::vector<int32_t> ids;
    if (piece == "<|startoftext|>") ids.push_back(sot);
    else if (piece == "<|endoftext|>") ids.push_back(eot);
    else {
      std::vector<std::string> sym;
      sym.reserve(piece.size());
      for (unsigned char c : piece) sym.push_back(byte_sym[c]);
      sym.back() += "</w>";
      std::string key;
      while (sym.size() > 1) {   // fuse the adjacent pair with the lowest merge rank, every occurrence left to right
        int best = -1;
        size_t best_i = 0;
        for (size_t i = 0; i + 1 < sym.size(); ++i) {
          key.assign(sym[i]); key.push_back('\x01'); key.append(sym[i + 1]);
          auto r = rank.find(key);
          if (r != rank.end() && (best < 0 || r->second < best)) { best = r->second; best_i = i; }
        }
        if (best < 0) break;
        const std::string a = sym[best_i], b = sym[best_i + 1];
        std::vector<std::string> out;
        out.reserve(sym.size());
        for (size_t i = 0; i < sym.size();) {
          if (i + 1 < sym.size() && sym[i] == a && sym[i + 1] == b) { out.push_back(a + b); i += 2; }
          else { out.push_back(sym[i]); i += 1; }
        }
        sym.swap(out);
      }
      for (const auto& s : sym) {
        auto e = encoder.find(s);
        ids.push_back(e == encoder.end() ? -1 : e->second);   // cannot miss: every merge product is a vocabulary entry
      }
    }
    return memo.emplace(piece, std::move(ids)).first->second;
  }

  // simple_tokenizer.py encode() on a cleaned, lower-cased ASCII text: the CLIP pattern's alternatives in order
  void encode(const std::string& t, std::vector<int32_t>& out) {
    static const char* kContr[] = {"'s", "'t", "'re", "'ve", "'m", "'ll", "'d"};
    static const std::string kSot = "<|startoftext|>", kEot = "<|endoftext|>";
    size_t i = 0;
    const size_t n = t.size();
    std::string piece;
    while (i < n) {
      const unsigned char c = static_cast<unsigned char>(t[i]);
      size_t len = 0;
      if (t.compare(i, kSot.size(), kSot) == 0) len = kSot.size();
      else if (t.compare(i, kEot.size(), kEot) == 0) len = kEot.size();
      else if (c == '\'') {
        for (const char* k : kContr) {
          const size_t kl = std::strlen(k);
          if (t.compare(i, kl, k) == 0) { len = kl; break; }
        }
      }
      if (len == 0) {
        if (is_alpha(c)) { size_t j = i; while (j < n && is_alpha(static_cast<unsigned char>(t[j]))) ++j; len = j - i; }
        else if (is_digit(c)) len = 1;
        else if (is_space(c)) { ++i; continue; }
        else {
          size_t j = i;
          while (j < n) {
            const unsigned char d = static_cast<unsigned char>(t[j]);
            if (is_space(d) || is_alpha(d) || is_digit(d)) break;
            ++j;
          }
          len = j - i;
        }
      }
      piece.assign(t, i, len);
      const auto& ids = encode_piece(piece);
      out.insert(out.end(), ids.begin(), ids.end());
      i += len;
    }
  }
};

extern "C" {

int hb_bpe_create(const char* merges, int64_t merges_bytes, HbBpe** out) {
  if (!merges || !out || merges_bytes <= 0) return HB_ERR_INVALID;
  std::unique_ptr<HbBpe> b(new (std::nothrow) HbBpe);
  if (!b) return HB_ERR_NOMEM;
  // bytes -> printable code points (the GPT-2 table, simple_tokenizer.py:14-37): printable Latin-1 bytes map to themselves,
  // the other 68 to U+0100 onwards; vocabulary order = the self-mapped ones first, then the remapped ones
  std::vector<std::string> direct, remapped;
  unsigned extra = 0;
  for (unsigned v = 0; v < 256; ++v) {
    const bool keep = (v >= 0x21 && v <= 0x7E) || (v >= 0xA1 && v <= 0xAC) || (v >= 0xAE);
    b->byte_sym[v] = utf8(keep ? v : 256 + extra);
    if (keep) direct.push_back(b->byte_sym[v]); else { remapped.push_back(b->byte_sym[v]); ++extra; }
  }
  int32_t id = 0;
  std::vector<std::string> base(direct);
  base.insert(base.end(), remapped.begin(), remapped.end());
  for (const auto& s : base) b->encoder[s] = id++;
  for (const auto& s : base) b->encoder[s + "</w>"] = id++;
  const char* p = merges;
  const char* end = merges + merges_bytes;
  int32_t r = 0;
  while (p < end) {
    const char* nl = static_cast<const char*>(memchr(p, '\n', static_cast<size_t>(end - p)));
    const char* le = nl ? nl : end;
    const char* sp = static_cast<const char*>(memchr(p, ' ', static_cast<size_t>(le - p)));
    if (!sp || sp == p || sp + 1 >= le) return HB_ERR_INVALID;   // every line is "left right"
    const std::string a(p, sp), c(sp + 1, le);
    std::string key(a); key.push_back('\x01'); key.append(c);
    // vocab = base + base</w> + [one entry PER MERGE] + specials; encoder = dict(zip(vocab, range)) and the rank table are Python
    // dicts in the reference, so a repeated key keeps its LAST value
    b->rank[key] = r++;
    b->encoder[a + c] = id++;
    p = nl ? nl + 1 : end;
  }
  b->sot = id++;
  b->eot = id++;
  b->encoder["<|startoftext|>"] = b->sot;
  b->encoder["<|endoftext|>"] = b->eot;
  *out = b.release();
  return HB_OK;
}

void hb_bpe_destroy(HbBpe* b) { delete b; }

int hb_bpe_tokenize(HbBpe* b, const char* const* texts, int n, int context_length, int truncate, int64_t* out, uint8_t* status) {
  if (!b || !texts || !out || !status || n < 0 || context_length < 2) return HB_ERR_INVALID;
  std::vector<int32_t> ids;
  std::string clean;
  for (int i = 0; i < n; ++i) {
    int64_t* row = out + static_cast<size_t>(i) * context_length;
    std::memset(row, 0, sizeof(int64_t) * context_length);
    const char* t = texts[i];
    bool ok = t != nullptr;
    if (ok) for (const unsigned char* p = reinterpret_cast<const unsigned char*>(t); *p; ++p)
      if (*p >= 0x7F || *p == '&' || (*p < 0x20 && !is_space(*p))) { ok = false; break; }   // Unicode / entities / control: Python path
    if (!ok) { status[i] = 1; continue; }
    // clean(): strip, runs of whitespace -> one space, lower-case
    clean.clear();
    bool pending_space = false;
    for (const unsigned char* p = reinterpret_cast<const unsigned char*>(t); *p; ++p) {
      if (is_space(*p)) { pending_space = !clean.empty(); continue; }
      if (pending_space) { clean.push_back(' '); pending_space = false; }
      clean.push_back(static_cast<char>(lower(*p)));
    }
    ids.clear();
    ids.push_back(b->sot);
    b->encode(clean, ids);
    ids.push_back(b->eot);
    for (int32_t v : ids) if (v < 0) { ok = false; break; }
    if (!ok) { status[i] = 1; continue; }                      // a symbol outside the table: let the Python path raise as it would
    if (ids.size() > static_cast<size_t>(context_length)) {
      if (!truncate) { status[i] = 2; continue; }              // caller raises "Input ... is too long for context length"
      ids.resize(static_cast<size_t>(context_length));
      ids.back() = b->eot;
    }
    for (size_t k = 0; k < ids.size(); ++k) row[k] = ids[k];
    status[i] = 0;
  }
  return HB_OK;
}

}  // extern "C"
