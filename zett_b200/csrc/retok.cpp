// libzett_b200.so -- surface-form half of the C ABI (include/zett_b200.h): a native, multi-threaded
// get_surface_form_matrix (reference zett/utils.py:651-689).
//
// The reference calls `tokenizer_to_use._tokenizer.model.tokenize(token)` once per target token from a Python loop
// (zett/utils.py:681); the model is HF `tokenizers`' Unigram or BPE (third-party Rust, not vendored in the reference).
// Both algorithms are implemented here from their published behaviour:
//   Unigram  Viterbi over byte offsets advancing by whole UTF-8 chars, f64 scores, unk_score = min_score - 10, a
//            candidate replaces the incumbent only when strictly greater, an unk candidate covers one char when no
//            piece of exactly that char's length matched, consecutive unk nodes are fused, piece -> id with
//            byte fallback (<0xXX>) or unk.
//   BPE      chars (+ continuing_subword_prefix / end_of_word_suffix) -> ids (byte fallback / unk with optional
//            fusing / dropped), then repeatedly the lowest (rank, position) merge via a heap with stale-entry checks.
// Bit-exactness is pinned by tests against goldens produced by the reference itself (tests/golden/make_golden.py).
#include <algorithm>
#include <atomic>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <limits>
#include <queue>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#include "../../include/zett_b200.h"

namespace {

thread_local std::string g_tok_error;

inline int utf8_len(unsigned char c) {
  if (c < 0x80) return 1;
  if (c >= 0xF0) return 4;
  if (c >= 0xE0) return 3;
  return 2;
}

// decode one UTF-8 char starting at s[i] (no validation beyond bounds), returns code point
inline uint32_t utf8_decode(const std::string& s, size_t i, int len) {
  const unsigned char* p = reinterpret_cast<const unsigned char*>(s.data()) + i;
  switch (len) {
    case 1: return p[0];
    case 2: return ((p[0] & 0x1Fu) << 6) | (p[1] & 0x3Fu);
    case 3: return ((p[0] & 0x0Fu) << 12) | ((p[1] & 0x3Fu) << 6) | (p[2] & 0x3Fu);
    default: return ((p[0] & 0x07u) << 18) | ((p[1] & 0x3Fu) << 12) | ((p[2] & 0x3Fu) << 6) | (p[3] & 0x3Fu);
  }
}

// membership in the GPT-2 byte alphabet CHARS_TO_BYTES (reference zett/utils.py:351-609): the 188 printable bytes map
// to themselves, the other 68 bytes to U+0100 .. U+0143
inline bool in_byte_alphabet(uint32_t cp) {
  return (cp >= 33 && cp <= 126) || (cp >= 161 && cp <= 172) || (cp >= 174 && cp <= 255) || (cp >= 256 && cp < 256 + 68);
}

// byte trie with an open-addressing edge table: (node, byte) -> child
class ByteTrie {
 public:
  ByteTrie() { value_.push_back(-1); }
  void reserve_edges(size_t n) {
    size_t cap = 16;
    while (cap < n * 2) cap <<= 1;
    keys_.assign(cap, kEmpty);
    vals_.assign(cap, 0);
    mask_ = cap - 1;
  }
  void insert(const std::string& s, int32_t id) {
    int32_t node = 0;
    for (unsigned char c : s) {
      int32_t nxt = child(node, c);
      if (nxt < 0) {
        nxt = static_cast<int32_t>(value_.size());
        value_.push_back(-1);
        put(node, c, nxt);
      }
      node = nxt;
    }
    value_[node] = id;  // later duplicates win
  }
  inline int32_t child(int32_t node, unsigned char c) const {
    const uint64_t key = (static_cast<uint64_t>(node) << 8) | c;
    size_t i = hash(key) & mask_;
    while (true) {
      const uint64_t k = keys_[i];
      if (k == key) return vals_[i];
      if (k == kEmpty) return -1;
      i = (i + 1) & mask_;
    }
  }
  inline int32_t value(int32_t node) const { return value_[node]; }
  int32_t find(const char* s, size_t n) const {
    int32_t node = 0;
    for (size_t i = 0; i < n; ++i) {
      node = child(node, static_cast<unsigned char>(s[i]));
      if (node < 0) return -1;
    }
    return value_[node];
  }

 private:
  static constexpr uint64_t kEmpty = ~0ull;
  static inline uint64_t hash(uint64_t x) {
    x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33;
    return x;
  }
  void put(int32_t node, unsigned char c, int32_t v) {
    if ((used_ + 1) * 2 > keys_.size()) grow();
    const uint64_t key = (static_cast<uint64_t>(node) << 8) | c;
    size_t i = hash(key) & mask_;
    while (keys_[i] != kEmpty) i = (i + 1) & mask_;
    keys_[i] = key;
    vals_[i] = v;
    ++used_;
  }
  void grow() {
    std::vector<uint64_t> ok;
    std::vector<int32_t> ov;
    ok.swap(keys_);
    ov.swap(vals_);
    const size_t cap = std::max<size_t>(16, ok.size() * 2);
    keys_.assign(cap, kEmpty);
    vals_.assign(cap, 0);
    mask_ = cap - 1;
    used_ = 0;
    for (size_t i = 0; i < ok.size(); ++i)
      if (ok[i] != kEmpty) {
        size_t j = hash(ok[i]) & mask_;
        while (keys_[j] != kEmpty) j = (j + 1) & mask_;
        keys_[j] = ok[i];
        vals_[j] = ov[i];
        ++used_;
      }
  }
  std::vector<uint64_t> keys_ = std::vector<uint64_t>(16, kEmpty);
  std::vector<int32_t> vals_ = std::vector<int32_t>(16, 0);
  size_t mask_ = 15, used_ = 0;
  std::vector<int32_t> value_;
};

// (left id, right id) -> (rank, merged id): open addressing, one probe on average -- the BPE loop does ~2 look-ups per
// symbol and std::unordered_map made them the largest item of a retokenisation
class PairMap {
 public:
  void reserve(size_t n) {
    size_t cap = 16;
    while (cap < n * 2) cap <<= 1;
    keys_.assign(cap, kEmpty);
    vals_.assign(cap, {0, 0});
    mask_ = cap - 1;
    used_ = 0;
  }
  void set(uint64_t key, std::pair<int32_t, int32_t> v) {  // later duplicates win
    if ((used_ + 1) * 2 > keys_.size()) grow();
    size_t i = hash(key) & mask_;
    while (keys_[i] != kEmpty && keys_[i] != key) i = (i + 1) & mask_;
    if (keys_[i] == kEmpty) { keys_[i] = key; ++used_; }
    vals_[i] = v;
  }
  inline const std::pair<int32_t, int32_t>* find(uint64_t key) const {
    size_t i = hash(key) & mask_;
    while (true) {
      const uint64_t k = keys_[i];
      if (k == key) return &vals_[i];
      if (k == kEmpty) return nullptr;
      i = (i + 1) & mask_;
    }
  }

 private:
  static constexpr uint64_t kEmpty = ~0ull;  // no pair of non-negative ids packs to this
  static inline uint64_t hash(uint64_t x) {
    x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33;
    return x;
  }
  void grow() {
    std::vector<uint64_t> ok;
    std::vector<std::pair<int32_t, int32_t>> ov;
    ok.swap(keys_);
    ov.swap(vals_);
    reserve(ok.size());
    for (size_t i = 0; i < ok.size(); ++i)
      if (ok[i] != kEmpty) set(ok[i], ov[i]);
  }
  std::vector<uint64_t> keys_ = std::vector<uint64_t>(16, kEmpty);
  std::vector<std::pair<int32_t, int32_t>> vals_ = std::vector<std::pair<int32_t, int32_t>>(16, {0, 0});
  size_t mask_ = 15, used_ = 0;
};

struct TokError {
  int code = 0;
};

}  // namespace

struct zett_tok {
  enum Kind { kUnigram, kBpe } kind = kUnigram;
  // shared
  ByteTrie trie;  // piece / vocab string -> id
  int64_t unk_id = -1;
  bool byte_fallback = false;
  int32_t byte_ids[256];
  bool all_byte_ids = false;
  // unigram
  std::vector<double> scores;
  double min_score = std::numeric_limits<double>::infinity();
  // bpe
  PairMap merges;  // (a << 32 | b) -> (rank, new id)
  std::string prefix, suffix;
  bool has_prefix = false, has_suffix = false, fuse_unk = false, ignore_merges = false;

  void index_byte_pieces() {
    all_byte_ids = true;
    for (int b = 0; b < 256; ++b) {
      char buf[8];
      snprintf(buf, sizeof buf, "<0x%02X>", b);
      byte_ids[b] = trie.find(buf, 6);
      if (byte_ids[b] < 0) all_byte_ids = false;
    }
  }

  // ---- Unigram -------------------------------------------------------------------------------------------------
  int unigram(const std::string& s, std::vector<int32_t>& out) const {
    const size_t n = s.size();
    out.clear();
    if (n == 0) return ZETT_OK;
    const double unk_score = min_score - 10.0;
    // per-thread scratch: a vocabulary is ~50k short strings, the allocations would cost as much as the search
    thread_local std::vector<double> best;
    thread_local std::vector<int32_t> start, node_id;
    thread_local std::vector<std::pair<int32_t, int32_t>> spans;  // [start, end)
    best.assign(n + 1, 0.0);
    start.assign(n + 1, -1);
    node_id.assign(n + 1, 0);
    spans.clear();
    size_t pos = 0;
    while (pos < n) {
      const double here = best[pos];
      const size_t mblen = std::min<size_t>(utf8_len(static_cast<unsigned char>(s[pos])), n - pos);
      bool single = false;
      int32_t node = 0;
      for (size_t end = pos; end < n;) {
        node = trie.child(node, static_cast<unsigned char>(s[end]));
        if (node < 0) break;
        ++end;
        const int32_t tid = trie.value(node);
        if (tid < 0) continue;
        const double cand = scores[tid] + here;
        if (start[end] < 0 || cand > best[end]) {
          best[end] = cand; start[end] = static_cast<int32_t>(pos); node_id[end] = tid;
        }
        if (end - pos == mblen) single = true;
      }
      if (!single) {
        const size_t end = pos + mblen;
        const double cand = unk_score + here;
        if (start[end] < 0 || cand > best[end]) {
          if (unk_id < 0) return ZETT_ERR_MISSING_UNK;
          best[end] = cand; start[end] = static_cast<int32_t>(pos); node_id[end] = static_cast<int32_t>(unk_id);
        }
      }
      pos += mblen;
    }
    // backtrack, fusing consecutive unk nodes; pieces come out right-to-left
    int32_t pend_start = -1, pend_end = -1;
    size_t end = n;
    while (end > 0) {
      const int32_t st = start[end];
      if (st < 0) return ZETT_ERR_INVALID;  // unreachable for well-formed UTF-8
      if (unk_id >= 0 && node_id[end] == unk_id) {
        if (pend_end < 0) pend_end = static_cast<int32_t>(end);
        pend_start = st;
      } else {
        if (pend_end >= 0) { spans.emplace_back(pend_start, pend_end); pend_end = -1; }
        spans.emplace_back(st, static_cast<int32_t>(end));
      }
      end = static_cast<size_t>(st);
    }
    if (pend_end >= 0) spans.emplace_back(pend_start, pend_end);
    for (auto it = spans.rbegin(); it != spans.rend(); ++it) {
      const int32_t tid = trie.find(s.data() + it->first, static_cast<size_t>(it->second - it->first));
      if (tid >= 0) { out.push_back(tid); continue; }
      if (byte_fallback) {
        bool ok = true;
        for (int32_t i = it->first; i < it->second; ++i) ok &= byte_ids[static_cast<unsigned char>(s[i])] >= 0;
        if (ok) {
          for (int32_t i = it->first; i < it->second; ++i) out.push_back(byte_ids[static_cast<unsigned char>(s[i])]);
          continue;
        }
      }
      if (unk_id < 0) return ZETT_ERR_MISSING_UNK;
      out.push_back(static_cast<int32_t>(unk_id));
    }
    return ZETT_OK;
  }

  // ---- BPE -----------------------------------------------------------------------------------------------------
  int bpe(const std::string& w, std::vector<int32_t>& out) const {
    out.clear();
    if (w.empty()) return ZETT_OK;
    if (ignore_merges) {
      const int32_t tid = trie.find(w.data(), w.size());
      if (tid >= 0) { out.push_back(tid); return ZETT_OK; }
    }
    thread_local std::vector<int32_t> c;
    thread_local std::string sym;
    c.clear();
    int32_t unk = -1;
    for (size_t i = 0; i < w.size();) {
      const size_t len = std::min<size_t>(utf8_len(static_cast<unsigned char>(w[i])), w.size() - i);
      const bool first = i == 0, last = i + len >= w.size();
      sym.clear();
      if (!first && has_prefix) sym += prefix;
      sym.append(w, i, len);
      if (last && has_suffix) sym += suffix;
      i += len;
      const int32_t tid = trie.find(sym.data(), sym.size());
      if (tid >= 0) {
        if (unk >= 0) { c.push_back(unk); unk = -1; }
        c.push_back(tid);
        continue;
      }
      if (byte_fallback) {
        bool ok = true;
        for (unsigned char b : sym) ok &= byte_ids[b] >= 0;
        if (ok) {
          for (unsigned char b : sym) c.push_back(byte_ids[b]);
          continue;
        }
      }
      if (unk_id >= 0) {
        if (unk >= 0 && !fuse_unk) c.push_back(unk);
        if (unk < 0 || !fuse_unk) unk = static_cast<int32_t>(unk_id);
      }
    }
    if (unk >= 0) c.push_back(unk);
    const int n = static_cast<int>(c.size());
    if (n == 0) return ZETT_OK;
    thread_local std::vector<char> alive;
    thread_local std::vector<int> prev, nxt;
    alive.assign(n, 1);
    prev.resize(n);
    nxt.resize(n);
    for (int i = 0; i < n; ++i) { prev[i] = i - 1; nxt[i] = (i + 1 < n) ? i + 1 : -1; }
    struct Item { int32_t rank; int pos; int32_t new_id; };
    auto cmp = [](const Item& a, const Item& b) {  // "a comes out after b": min-heap on (rank, pos, new id)
      if (a.rank != b.rank) return a.rank > b.rank;
      if (a.pos != b.pos) return a.pos > b.pos;
      return a.new_id > b.new_id;
    };
    thread_local std::vector<Item> heap;  // std::push_heap / pop_heap over per-thread storage: no allocation per token
    heap.clear();
    auto push = [&](Item it) { heap.push_back(it); std::push_heap(heap.begin(), heap.end(), cmp); };
    auto lookup = [&](int32_t a, int32_t b) -> const std::pair<int32_t, int32_t>* {
      return merges.find((static_cast<uint64_t>(static_cast<uint32_t>(a)) << 32) | static_cast<uint32_t>(b));
    };
    for (int i = 0; i + 1 < n; ++i)
      if (auto* m = lookup(c[i], c[i + 1])) push({m->first, i, m->second});
    while (!heap.empty()) {
      std::pop_heap(heap.begin(), heap.end(), cmp);
      const Item top = heap.back();
      heap.pop_back();
      const int pos = top.pos;
      if (!alive[pos] || nxt[pos] == -1) continue;
      const int r = nxt[pos];
      const auto* m = lookup(c[pos], c[r]);
      if (!m || m->second != top.new_id) continue;  // stale entry
      c[pos] = top.new_id;
      alive[r] = 0;
      nxt[pos] = nxt[r];
      if (nxt[r] != -1) prev[nxt[r]] = pos;
      if (prev[pos] >= 0)
        if (auto* m2 = lookup(c[prev[pos]], c[pos])) push({m2->first, prev[pos], m2->second});
      if (nxt[pos] != -1)
        if (auto* m3 = lookup(c[pos], c[nxt[pos]])) push({m3->first, pos, m3->second});
    }
    for (int i = 0; i < n; ++i)
      if (alive[i]) out.push_back(c[i]);
    return ZETT_OK;
  }

  int tokenize(const std::string& s, std::vector<int32_t>& out) const { return kind == kUnigram ? unigram(s, out) : bpe(s, out); }
};

namespace {
int tok_fail(int code, const std::string& msg);
}

extern "C" {

// the error string is shared with the hypernet half through zett_last_error(); declared in hypernet.cu
void zett_set_last_error_(const char* msg);

int zett_tok_create_unigram(const char* const* pieces, const double* scores, int64_t n, int64_t unk_id, int byte_fallback,
                            zett_tok** out) {
  if (!pieces || !scores || !out || n <= 0) return tok_fail(ZETT_ERR_INVALID, "bad argument");
  if (unk_id >= n) return tok_fail(ZETT_ERR_INVALID, "unk_id out of range");
  auto* t = new zett_tok();
  t->kind = zett_tok::kUnigram;
  t->unk_id = unk_id;
  t->byte_fallback = byte_fallback != 0;
  t->scores.assign(scores, scores + n);
  size_t total = 0;
  for (int64_t i = 0; i < n; ++i) total += strlen(pieces[i]);
  t->trie.reserve_edges(total);
  for (int64_t i = 0; i < n; ++i) {
    t->trie.insert(pieces[i], static_cast<int32_t>(i));
    t->min_score = std::min(t->min_score, scores[i]);
  }
  t->index_byte_pieces();
  *out = t;
  return ZETT_OK;
}

int zett_tok_create_bpe(const char* const* vocab, int64_t n, const int32_t* merges, int64_t m, int64_t unk_id,
                        const char* continuing_subword_prefix, const char* end_of_word_suffix, int fuse_unk,
                        int byte_fallback, int ignore_merges, zett_tok** out) {
  if (!vocab || !out || n <= 0 || m < 0 || (m > 0 && !merges)) return tok_fail(ZETT_ERR_INVALID, "bad argument");
  if (unk_id >= n) return tok_fail(ZETT_ERR_INVALID, "unk_id out of range");
  auto* t = new zett_tok();
  t->kind = zett_tok::kBpe;
  t->unk_id = unk_id;
  t->byte_fallback = byte_fallback != 0;
  t->fuse_unk = fuse_unk != 0;
  t->ignore_merges = ignore_merges != 0;
  if (continuing_subword_prefix && *continuing_subword_prefix) { t->prefix = continuing_subword_prefix; t->has_prefix = true; }
  if (end_of_word_suffix && *end_of_word_suffix) { t->suffix = end_of_word_suffix; t->has_suffix = true; }
  size_t total = 0;
  for (int64_t i = 0; i < n; ++i) total += strlen(vocab[i]);
  t->trie.reserve_edges(total);
  for (int64_t i = 0; i < n; ++i) t->trie.insert(vocab[i], static_cast<int32_t>(i));
  t->merges.reserve(static_cast<size_t>(m));
  const size_t plen = t->prefix.size();
  for (int64_t r = 0; r < m; ++r) {
    const int32_t a = merges[2 * r], b = merges[2 * r + 1];
    if (a < 0 || b < 0 || a >= n || b >= n) { delete t; return tok_fail(ZETT_ERR_INVALID, "merge id out of range"); }
    std::string joined = vocab[a];
    const char* bs = vocab[b];
    const size_t bl = strlen(bs);
    joined.append(bs + std::min(plen, bl), bl - std::min(plen, bl));
    const int32_t nid = t->trie.find(joined.data(), joined.size());
    if (nid < 0) { delete t; return tok_fail(ZETT_ERR_INVALID, "merge result not in vocab: " + joined); }
    t->merges.set((static_cast<uint64_t>(static_cast<uint32_t>(a)) << 32) | static_cast<uint32_t>(b), {static_cast<int32_t>(r), nid});
  }
  t->index_byte_pieces();
  *out = t;
  return ZETT_OK;
}

int64_t zett_tok_tokenize(const zett_tok* t, const char* token, int32_t* out_ids, int64_t cap) {
  if (!t || !token) return tok_fail(ZETT_ERR_INVALID, "bad argument");
  std::vector<int32_t> ids;
  const int rc = t->tokenize(token, ids);
  if (rc != ZETT_OK) return tok_fail(rc, rc == ZETT_ERR_MISSING_UNK ? "MissingUnkId" : "tokenize failed");
  for (int64_t i = 0; i < std::min<int64_t>(cap, static_cast<int64_t>(ids.size())); ++i) out_ids[i] = ids[i];
  return static_cast<int64_t>(ids.size());
}

}  // extern "C"

namespace {

using SpecialMap = std::unordered_map<std::string, int32_t>;

// shared body of the two entry points: special tokens are given per token (special_ids) or as a map matched in the workers
int surface_forms_impl(const zett_tok* t, const char* const* tokens, int64_t v, const int32_t* special_ids,
                       const SpecialMap* special_map, int32_t maxlen, int32_t pad_id, int64_t padding, int32_t* out,
                       int64_t* n_truncated, int n_threads) {
  if (!t || (!tokens && v > 0) || !out || v < 0 || maxlen <= 0 || padding < 0) return tok_fail(ZETT_ERR_INVALID, "bad argument");
  size_t max_special = 0;
  if (special_map)
    for (const auto& kv : *special_map) max_special = std::max(max_special, kv.first.size());
  std::fill(out, out + (v + padding) * maxlen, pad_id);
  int nt = n_threads > 0 ? n_threads : static_cast<int>(std::thread::hardware_concurrency());
  nt = std::max(1, std::min<int>(nt, static_cast<int>(std::max<int64_t>(1, v / 256))));
  std::atomic<int64_t> truncated{0};
  std::atomic<int64_t> first_bad{std::numeric_limits<int64_t>::max()};
  std::vector<int> bad_code(nt, 0);
  auto work = [&](int tid) {
    const int64_t lo = v * tid / nt, hi = v * (tid + 1) / nt;
    std::vector<int32_t> ids;
    std::string s;
    int64_t trunc = 0;
    for (int64_t i = lo; i < hi; ++i) {
      if (special_ids && special_ids[i] >= 0) {  // token in hn_tokenizer.all_special_tokens (utils.py:671-673)
        out[i * maxlen] = special_ids[i];
        continue;
      }
      s.assign(tokens[i]);
      if (special_map && s.size() <= max_special) {
        auto it = special_map->find(s);
        if (it != special_map->end()) { out[i * maxlen] = it->second; continue; }
      }
      int rc = ZETT_OK;
      for (size_t p = 0; p < s.size();) {  // bytes([CHARS_TO_BYTES[c] for c in token]) raises KeyError (utils.py:675)
        const int len = std::min<size_t>(utf8_len(static_cast<unsigned char>(s[p])), s.size() - p);
        if (!in_byte_alphabet(utf8_decode(s, p, len))) { rc = ZETT_ERR_KEY; break; }
        p += len;
      }
      if (rc == ZETT_OK) rc = t->tokenize(s, ids);
      if (rc != ZETT_OK) {
        int64_t cur = first_bad.load();
        while (i < cur && !first_bad.compare_exchange_weak(cur, i)) {}
        bad_code[tid] = rc;
        return;
      }
      size_t n = ids.size();
      if (n > static_cast<size_t>(maxlen)) { n = maxlen; ++trunc; }
      std::copy(ids.begin(), ids.begin() + n, out + i * maxlen);
    }
    truncated += trunc;
  };
  if (nt == 1) {
    work(0);
  } else {
    std::vector<std::thread> th;
    for (int i = 0; i < nt; ++i) th.emplace_back(work, i);
    for (auto& x : th) x.join();
  }
  const int64_t fb = first_bad.load();
  if (fb != std::numeric_limits<int64_t>::max()) {
    const int tid = static_cast<int>(std::min<int64_t>(nt - 1, (fb * nt + nt - 1) / std::max<int64_t>(v, 1)));
    int code = 0;
    for (int i = 0; i < nt; ++i)
      if (v * i / nt <= fb && fb < v * (i + 1) / nt) code = bad_code[i];
    (void)tid;
    if (code == ZETT_ERR_KEY) return tok_fail(code, std::string("token ") + std::to_string(fb) + " has a char outside the byte alphabet: " + tokens[fb]);
    if (code == ZETT_ERR_MISSING_UNK) return tok_fail(code, "MissingUnkId (token " + std::to_string(fb) + ")");
    return tok_fail(code ? code : ZETT_ERR_INVALID, "tokenize failed at token " + std::to_string(fb));
  }
  if (n_truncated) *n_truncated = truncated.load();
  return ZETT_OK;
}

}  // namespace

extern "C" {

int zett_surface_forms(const zett_tok* t, const char* const* tokens, int64_t v, const int32_t* special_ids, int32_t maxlen,
                       int32_t pad_id, int64_t padding, int32_t* out, int64_t* n_truncated, int n_threads) {
  return surface_forms_impl(t, tokens, v, special_ids, nullptr, maxlen, pad_id, padding, out, n_truncated, n_threads);
}

int zett_surface_forms_blob(const zett_tok* t, const char* tokens_blob, int64_t blob_bytes, int64_t v,
                            const char* special_blob, int64_t special_bytes, const int32_t* special_token_ids,
                            int64_t n_special, int32_t maxlen, int32_t pad_id, int64_t padding, int32_t* out,
                            int64_t* n_truncated, int n_threads) {
  if (!t || v < 0 || blob_bytes < 0 || (!tokens_blob && v > 0) || n_special < 0 || (n_special > 0 && (!special_blob || !special_token_ids)))
    return tok_fail(ZETT_ERR_INVALID, "bad argument");
  // token starts: one pass over the buffer (a few hundred KB for a whole vocabulary)
  std::vector<const char*> ptrs(static_cast<size_t>(v));
  const char* p = tokens_blob;
  const char* end = tokens_blob + blob_bytes;
  for (int64_t i = 0; i < v; ++i) {
    if (p > end) return tok_fail(ZETT_ERR_INVALID, "tokens_blob holds fewer than v NUL-terminated strings");
    ptrs[static_cast<size_t>(i)] = p;
    const void* z = p < end ? memchr(p, 0, static_cast<size_t>(end - p)) : nullptr;
    p = z ? static_cast<const char*>(z) + 1 : end + 1;  // the last string may rely on the buffer's own terminator
  }
  SpecialMap special;
  const char* q = special_blob;
  const char* qend = special_blob + special_bytes;
  for (int64_t i = 0; i < n_special; ++i) {
    if (q > qend) return tok_fail(ZETT_ERR_INVALID, "special_blob holds fewer than n_special strings");
    const void* z = q < qend ? memchr(q, 0, static_cast<size_t>(qend - q)) : nullptr;
    const char* stop = z ? static_cast<const char*>(z) : qend;
    special.emplace(std::string(q, stop), special_token_ids[i]);  // first occurrence wins, like list.index
    q = stop + 1;
  }
  return surface_forms_impl(t, ptrs.data(), v, nullptr, special.empty() ? nullptr : &special, maxlen, pad_id, padding, out,
                            n_truncated, n_threads);
}

void zett_tok_destroy(zett_tok* t) { delete t; }

}  // extern "C"

namespace {
int tok_fail(int code, const std::string& msg) {
  zett_set_last_error_(msg.c_str());
  return code;
}
}  // namespace
