// libzett_b200.so -- TokenizerSampler: the seed vocabulary of a freshly sampled Unigram tokenizer, host-side C++.
//
// Replaces rust_utils.TokenizerSampler.sample_tokenizer (reference rust_utils/src/lib.rs:69-250), the one native component
// of the reference, used by the TRAINING collator (zett/collator.py:341-452) to mint a new target tokenizer every step.
// Not on the inference path; built because SURVEY section 8f lists it as the last "next" row.
//
//   * pre-tokenisation = the reference's Sequence[Split(GPT-2 regex, Removed, invert), ByteLevel(no prefix space, no
//     regex)] (lib.rs:27, 83-93).  The regex needs Unicode classes and a look-ahead; it is a hand-written scanner here
//     over generated category tables (unicode_tables.inc), alternatives tried in the regex's order:
//         's|'t|'re|'ve|'m|'ll|'d | ?\p{L}+ | ?\p{N}+ | ?[^\s\p{L}\p{N}]+ |\s+(?!\S)|\s+
//   * for every pre-token, every listed start and every length < max_length: score[substring] += count * byte length
//     (lib.rs:119-160), with the reference's start list as it is -- a duplicate 0 for the first pre-token and
//     end-of-character offsets for multi-byte characters (lib.rs:123-130);
//   * the cache of the last batches' tables and the assembly of the seed list (lib.rs:163-245).
// What the reference leaves to chance is pinned (and documented in oracle/sampler_oracle.py, which this file is tested
// against at noise_std = 0): the noise generator takes a seed, the byte alphabet is emitted in byte order, ties in the
// score order are broken by the piece's bytes.  Scores accumulate in 64 bits (the reference's u32 wraps in release builds).
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <random>
#include <string>
#include <unordered_map>
#include <utility>
#include <vector>

#include "../../include/zett_b200.h"

extern "C" void zett_set_last_error_(const char* msg);

namespace {

#include "unicode_tables.inc"

template <size_t N>
bool in_ranges(const uint32_t (&r)[N][2], uint32_t cp) {
  size_t lo = 0, hi = N;
  while (lo < hi) {
    const size_t mid = (lo + hi) / 2;
    if (cp < r[mid][0]) hi = mid;
    else if (cp > r[mid][1]) lo = mid + 1;
    else return true;
  }
  return false;
}
bool is_letter(uint32_t cp) { return in_ranges(kUnicodeLetterRanges, cp); }
bool is_number(uint32_t cp) { return in_ranges(kUnicodeNumberRanges, cp); }
bool is_space(uint32_t cp) {  // Unicode White_Space, what \s means to the reference's regex engine
  return (cp >= 0x9 && cp <= 0xD) || cp == 0x20 || cp == 0x85 || cp == 0xA0 || cp == 0x1680 || (cp >= 0x2000 && cp <= 0x200A) ||
         cp == 0x2028 || cp == 0x2029 || cp == 0x202F || cp == 0x205F || cp == 0x3000;
}
bool is_other(uint32_t cp) { return !is_space(cp) && !is_letter(cp) && !is_number(cp); }

struct Char {
  uint32_t cp;
  uint32_t byte_begin, byte_end;  // in the sentence
};

// strict-enough UTF-8 decoder; returns false on malformed input
bool decode_utf8(const std::string& s, std::vector<Char>* out) {
  size_t i = 0;
  while (i < s.size()) {
    const unsigned char c = static_cast<unsigned char>(s[i]);
    uint32_t cp;
    size_t n;
    if (c < 0x80) { cp = c; n = 1; }
    else if ((c >> 5) == 0x6) { cp = c & 0x1F; n = 2; }
    else if ((c >> 4) == 0xE) { cp = c & 0x0F; n = 3; }
    else if ((c >> 3) == 0x1E) { cp = c & 0x07; n = 4; }
    else return false;
    if (i + n > s.size()) return false;
    for (size_t k = 1; k < n; ++k) {
      const unsigned char d = static_cast<unsigned char>(s[i + k]);
      if ((d >> 6) != 0x2) return false;
      cp = (cp << 6) | (d & 0x3F);
    }
    out->push_back(Char{cp, static_cast<uint32_t>(i), static_cast<uint32_t>(i + n)});
    i += n;
  }
  return true;
}

// one match of the GPT-2 split regex starting at character `p`; returns the end (exclusive, > p)
size_t match_at(const std::vector<Char>& t, size_t p) {
  const size_t n = t.size();
  auto cp = [&](size_t i) { return i < n ? t[i].cp : 0xFFFFFFFFu; };
  if (cp(p) == '\'') {  // 's|'t|'re|'ve|'m|'ll|'d
    const uint32_t a = cp(p + 1), b = cp(p + 2);
    if (a == 's' || a == 't') return p + 2;
    if ((a == 'r' && b == 'e') || (a == 'v' && b == 'e')) return p + 3;
    if (a == 'm') return p + 2;
    if (a == 'l' && b == 'l') return p + 3;
    if (a == 'd') return p + 2;
  }
  const size_t q = (cp(p) == ' ') ? p + 1 : p;  // " ?": tried with the space first; without it the classes below cannot match a space
  auto run = [&](size_t from, bool (*pred)(uint32_t)) {
    size_t e = from;
    while (e < n && pred(t[e].cp)) ++e;
    return e;
  };
  if (q < n && is_letter(t[q].cp)) return run(q, is_letter);
  if (q < n && is_number(t[q].cp)) return run(q, is_number);
  if (q < n && is_other(t[q].cp)) return run(q, is_other);
  // here t[p] is white space (anything else matched above)
  const size_t e = run(p, is_space);
  if (e == n) return e;          // \s+(?!\S) at the end of the text
  if (e - p >= 2) return e - 1;  // \s+(?!\S): give back one character so that white space follows
  return e;                      // \s+
}

// GPT-2 bytes_to_unicode: byte -> code point of its byte-level character
const uint32_t* byte_level_table() {
  static uint32_t table[256];
  static bool ready = false;
  if (!ready) {
    int extra = 0;
    for (int b = 0; b < 256; ++b) {
      const bool keep = (b >= '!' && b <= '~') || (b >= 0xA1 && b <= 0xAC) || (b >= 0xAE && b <= 0xFF);
      table[b] = keep ? static_cast<uint32_t>(b) : static_cast<uint32_t>(256 + extra++);
    }
    ready = true;
  }
  return table;
}
void append_utf8(std::string* s, uint32_t cp) {  // byte-level characters are below U+0800
  if (cp < 0x80) s->push_back(static_cast<char>(cp));
  else { s->push_back(static_cast<char>(0xC0 | (cp >> 6))); s->push_back(static_cast<char>(0x80 | (cp & 0x3F))); }
}

using Table = std::unordered_map<std::string, uint64_t>;

int fail(int code, const std::string& msg) {
  zett_set_last_error_(msg.c_str());
  return code;
}

// lib.rs:95-161
int substring_scores(const std::string& text, uint64_t count, size_t max_length, size_t stride, Table* index) {
  const std::string sentence = " " + text;  // prefix space (lib.rs:99)
  std::vector<Char> chars;
  if (!decode_utf8(sentence, &chars)) return fail(ZETT_ERR_INVALID, "TokenizerSampler: text is not valid UTF-8");
  const uint32_t* bl = byte_level_table();
  size_t p = 0;
  bool first = true;
  std::vector<size_t> starts, char_pos;
  std::string pretoken;
  while (p < chars.size()) {
    const size_t e = match_at(chars, p);
    // byte-level spelling of the piece: one character per byte of the original
    pretoken.clear();
    char_pos.clear();
    for (uint32_t b = chars[p].byte_begin; b < chars[e - 1].byte_end; ++b) {
      char_pos.push_back(pretoken.size());
      append_utf8(&pretoken, bl[static_cast<unsigned char>(sentence[b])]);
    }
    const size_t nb = char_pos.size();
    // start list (lib.rs:123-130): cumulative byte length up to and including character j, minus the same for the piece's
    // first character -- i.e. end-of-character offsets relative to the END of the first character -- used as indices into
    // the byte-level characters; the first piece gets an extra 0 in front
    starts.clear();
    if (first) starts.push_back(0);
    for (size_t j = p; j < e; ++j) starts.push_back(chars[j].byte_end - chars[p].byte_end);
    for (size_t si = 0; si < starts.size(); si += stride) {
      const size_t s = starts[si];
      for (size_t k = 1; k < max_length; ++k) {
        if (s + k > nb) break;
        const size_t b0 = char_pos[s], b1 = (s + k == nb) ? pretoken.size() : char_pos[s + k];
        if (b1 == b0) continue;
        (*index)[pretoken.substr(b0, b1 - b0)] += count * static_cast<uint64_t>(b1 - b0);
      }
    }
    first = false;
    p = e;
  }
  return ZETT_OK;
}

size_t count_chars(const std::string& s) {
  size_t n = 0;
  for (unsigned char c : s) n += (c & 0xC0) != 0x80;
  return n;
}

}  // namespace

struct zett_sampler {
  std::deque<Table> seed_cache;
};

extern "C" {

int zett_sampler_create(zett_sampler** out) {
  if (!out) return fail(ZETT_ERR_INVALID, "null argument");
  *out = new zett_sampler();
  return ZETT_OK;
}

void zett_sampler_destroy(zett_sampler* s) { delete s; }

void zett_sampler_free(void* p) { free(p); }

int zett_sampler_sample(zett_sampler* s, const char* texts_blob, int64_t blob_bytes, const uint32_t* counts, int64_t n_texts,
                        int64_t seed_size, int64_t max_length, int64_t stride, double noise_std, uint64_t noise_seed, int pop_prev,
                        int push_current, char** out_pieces_blob, int64_t* out_blob_bytes, double** out_scores, int64_t* out_n) {
  if (!s || (!texts_blob && n_texts > 0) || (!counts && n_texts > 0) || !out_pieces_blob || !out_blob_bytes || !out_scores || !out_n)
    return fail(ZETT_ERR_INVALID, "null argument");
  if (n_texts < 0 || seed_size < 0 || max_length < 1 || stride < 1 || noise_std < 0) return fail(ZETT_ERR_INVALID, "bad argument");
  Table current;
  const char* p = texts_blob;
  const char* end = texts_blob + blob_bytes;
  for (int64_t i = 0; i < n_texts; ++i) {
    if (p > end) return fail(ZETT_ERR_INVALID, "texts blob shorter than n_texts strings");
    const char* q = static_cast<const char*>(memchr(p, 0, static_cast<size_t>(end - p)));
    const size_t len = q ? static_cast<size_t>(q - p) : static_cast<size_t>(end - p);
    const int rc = substring_scores(std::string(p, len), counts[i], static_cast<size_t>(max_length), static_cast<size_t>(stride), &current);
    if (rc != ZETT_OK) return rc;
    p += len + 1;
  }
  // lib.rs:163-176
  bool have_prev = false;
  Table prev;
  if (pop_prev && !s->seed_cache.empty()) {
    prev = std::move(s->seed_cache.back());
    s->seed_cache.pop_back();
    have_prev = true;
  }
  s->seed_cache.push_front(std::move(current));
  std::vector<std::pair<std::string, double>> seed;
  if (pop_prev) {
    Table merged;
    for (const Table& t : s->seed_cache)
      for (const auto& kv : t) merged[kv.first] += kv.second;
    double score_sum = 0;
    uint64_t min_score = 0xFFFFFFFFull;
    for (const auto& kv : merged) {
      score_sum += static_cast<double>(kv.second);
      min_score = std::min(min_score, kv.second);
    }
    const double min_log_prob = std::log(static_cast<double>(min_score) / score_sum);
    const uint32_t* bl = byte_level_table();
    for (int b = 0; b < 256; ++b) {  // ByteLevel::alphabet(), in byte order
      std::string c;
      append_utf8(&c, bl[b]);
      seed.emplace_back(c, min_log_prob);
    }
    std::vector<std::pair<std::string, uint64_t>> items(merged.begin(), merged.end());
    std::sort(items.begin(), items.end());  // a fixed order before the noise is drawn
    std::mt19937_64 rng(noise_seed);
    std::normal_distribution<double> normal(0.0, noise_std > 0 ? noise_std : 1.0);
    std::vector<std::pair<std::string, double>> scored;
    scored.reserve(items.size());
    for (auto& kv : items) {
      const double noised = static_cast<double>(kv.second) / score_sum + (noise_std > 0 ? normal(rng) : 0.0);
      scored.emplace_back(std::move(kv.first), noised > 0.0 ? std::log(noised) : -100000.0);
    }
    std::sort(scored.begin(), scored.end(), [](const auto& a, const auto& b) { return a.second != b.second ? a.second > b.second : a.first < b.first; });
    const char* extra[3] = {"\xC4\xA0", "\xC4\x8A", "\xC4\x89"};  // the byte-level spellings of space, newline, tab
    for (int c1 = 0; c1 < 3; ++c1)
      for (int64_t i = 1; i < max_length; ++i)
        for (int c2 = 0; c2 < 3; ++c2) {
          std::string piece = extra[c2];
          for (int64_t r = 0; r < i; ++r) piece += extra[c1];
          seed.emplace_back(piece, 0.0);
        }
    for (auto& kv : scored) {
      const std::string& piece = kv.first;
      size_t ws = 0;
      for (size_t i = 0; i + 1 < piece.size(); ++i)
        if (static_cast<unsigned char>(piece[i]) == 0xC4 && (static_cast<unsigned char>(piece[i + 1]) == 0xA0 ||
                                                              static_cast<unsigned char>(piece[i + 1]) == 0x8A ||
                                                              static_cast<unsigned char>(piece[i + 1]) == 0x89))
          ++ws;
      if (count_chars(piece) == 1 || ws >= 2) continue;  // already added
      seed.emplace_back(piece, kv.second);
      if (static_cast<int64_t>(seed.size()) >= seed_size) break;
    }
  }
  if (!push_current) {  // lib.rs:240-245
    s->seed_cache.pop_front();
    if (have_prev) s->seed_cache.push_back(std::move(prev));
  }
  size_t bytes = 0;
  for (const auto& kv : seed) bytes += kv.first.size() + 1;
  char* blob = static_cast<char*>(malloc(bytes ? bytes : 1));
  double* scores = static_cast<double*>(malloc(sizeof(double) * (seed.empty() ? 1 : seed.size())));
  if (!blob || !scores) { free(blob); free(scores); return fail(ZETT_ERR_STATE, "out of memory"); }
  size_t off = 0;
  for (size_t i = 0; i < seed.size(); ++i) {
    memcpy(blob + off, seed[i].first.c_str(), seed[i].first.size() + 1);
    off += seed[i].first.size() + 1;
    scores[i] = seed[i].second;
  }
  *out_pieces_blob = blob;
  *out_blob_bytes = static_cast<int64_t>(bytes);
  *out_scores = scores;
  *out_n = static_cast<int64_t>(seed.size());
  return ZETT_OK;
}

}  // extern "C"
