// text_host.h — host-side text routines of the sparse path: the Snowball "english" (Porter2) stemmer and the
// ASCII fast path of bm25s' tokeniser.  Pure C++, no CUDA; exported through the C ABI (vfi_stem_english,
// vfi_tokenize_ascii) so the BM25 path is self-contained.
//
// Reference call sites (relative to /root/reference/): src/utils/bm25Retriever.py:14-15 (index build) and :47,67
// (query): `Stemmer.Stemmer('english')` handed to `bm25s.tokenize(..., stopwords="english", stemmer=...)`.
// PyStemmer wraps the Snowball C library; neither is vendored or pinned by the reference, so this restates the
// published algorithm (snowball/algorithms/english.sbl as shipped with Snowball 2.0 - 2.2, the releases PyStemmer
// 2.x wraps) from its definition, in UTF-8 mode: bytes >= 0x80 belong to non-vowel characters, and the few places
// where the algorithm counts characters (hop 3, hop 2, next) count UTF-8 characters, not bytes.
#pragma once
#include <cstdint>
#include <cstring>
#include <string>

namespace vfi_text {

inline bool is_vowel(unsigned char c) { return c == 'a' || c == 'e' || c == 'i' || c == 'o' || c == 'u' || c == 'y'; }
inline bool is_cont(unsigned char c) { return (c & 0xC0) == 0x80; }

class EnglishStemmer {
 public:
  // stems `in` (UTF-8, expected lower-case) into w_; returns the result view
  const std::string& stem(const char* in, size_t len) {
    w_.assign(in, len);
    if (exception1()) return w_;
    if (char_count(0, w_.size()) < 3) return w_;
    prelude();
    mark_regions();
    step_1a();
    if (!exception2()) {
      step_1b();
      step_1c();
      step_2();
      step_3();
      step_4();
      step_5();
    }
    for (char& c : w_)
      if (c == 'Y') c = 'y';
    return w_;
  }

 private:
  std::string w_;
  size_t p1_ = 0, p2_ = 0;

  size_t char_count(size_t a, size_t b) const {
    size_t n = 0;
    for (size_t i = a; i < b; ++i) n += !is_cont(static_cast<unsigned char>(w_[i]));
    return n;
  }
  // start of the character that ends at byte position `end` (end > 0)
  size_t prev(size_t end) const {
    size_t i = end - 1;
    while (i > 0 && is_cont(static_cast<unsigned char>(w_[i]))) --i;
    return i;
  }
  bool vowel_at(size_t i) const { return is_vowel(static_cast<unsigned char>(w_[i])); }
  bool ends(const char* s, size_t n) const { return w_.size() >= n && std::memcmp(w_.data() + w_.size() - n, s, n) == 0; }
  bool ends(const char* s) const { return ends(s, std::strlen(s)); }
  bool is_word(const char* s) const { return w_.size() == std::strlen(s) && w_ == s; }
  void replace_suffix(size_t n, const char* by) {
    w_.resize(w_.size() - n);
    w_.append(by);
  }
  // a vowel anywhere in [0, end)
  bool has_vowel_before(size_t end) const {
    for (size_t i = 0; i < end; ++i)
      if (vowel_at(i)) return true;
    return false;
  }
  // does a short syllable end at byte position `end`?
  bool short_syllable(size_t end) const {
    if (end < 2) return false;
    const size_t c = prev(end);                       // the final non-vowel
    if (vowel_at(c)) return false;
    if (c == 0) return false;
    const size_t v = c - 1;                           // vowels are single bytes
    if (!vowel_at(v)) return false;
    if (v == 0) return true;                          // (b) vowel at the beginning followed by a non-vowel
    const unsigned char last = static_cast<unsigned char>(w_[c]);
    if (last == 'w' || last == 'x' || last == 'Y') return false;
    return !vowel_at(prev(v));                        // (a) non-vowel, vowel, non-vowel other than w x Y
  }

  bool exception1() {
    static const char* const kMap[][2] = {
        {"skis", "ski"},   {"skies", "sky"},   {"dying", "die"},   {"lying", "lie"},   {"tying", "tie"},   {"idly", "idl"},
        {"gently", "gentl"}, {"ugly", "ugli"}, {"early", "earli"}, {"only", "onli"},   {"singly", "singl"}, {"sky", "sky"},
        {"news", "news"},  {"howe", "howe"},   {"atlas", "atlas"}, {"cosmos", "cosmos"}, {"bias", "bias"}, {"andes", "andes"}};
    for (const auto& e : kMap)
      if (is_word(e[0])) {
        w_ = e[1];
        return true;
      }
    return false;
  }
  bool exception2() const {
    static const char* const kWords[] = {"inning", "outing", "canning", "herring", "earring", "proceed", "exceed", "succeed"};
    for (const char* s : kWords)
      if (is_word(s)) return true;
    return false;
  }

  void prelude() {
    if (!w_.empty() && w_[0] == '\'') w_.erase(0, 1);
    if (!w_.empty() && w_[0] == 'y') w_[0] = 'Y';
    for (size_t i = 0; i + 1 < w_.size(); ++i)
      if (vowel_at(i) && w_[i + 1] == 'y') w_[i + 1] = 'Y';
  }

  // Snowball's `gopast v gopast non-v` from pos: on success pos is just after the first non-vowel that follows a vowel
  bool gopast_vowel_consonant(size_t& pos) const {
    size_t i = pos;
    while (i < w_.size() && !vowel_at(i)) ++i;
    if (i >= w_.size()) return false;
    ++i;
    while (i < w_.size() && vowel_at(i)) ++i;
    if (i >= w_.size()) return false;
    ++i;
    while (i < w_.size() && is_cont(static_cast<unsigned char>(w_[i]))) ++i;   // the non-vowel may be multi-byte
    pos = i;
    return true;
  }
  void mark_regions() {
    p1_ = p2_ = w_.size();
    size_t pos = 0;
    if (w_.compare(0, 5, "gener") == 0) pos = 5;
    else if (w_.compare(0, 6, "commun") == 0) pos = 6;
    else if (w_.compare(0, 5, "arsen") == 0) pos = 5;
    else if (!gopast_vowel_consonant(pos)) return;
    p1_ = pos;
    if (gopast_vowel_consonant(pos)) p2_ = pos;
  }

  bool in_r1(size_t suffix_len) const { return w_.size() - suffix_len >= p1_; }
  bool in_r2(size_t suffix_len) const { return w_.size() - suffix_len >= p2_; }

  void step_1a() {
    if (ends("'s'")) w_.resize(w_.size() - 3);
    else if (ends("'s")) w_.resize(w_.size() - 2);
    else if (ends("'")) w_.resize(w_.size() - 1);
    if (ends("sses")) { replace_suffix(4, "ss"); return; }
    if (ends("ied") || ends("ies")) {
      replace_suffix(3, char_count(0, w_.size() - 3) > 1 ? "i" : "ie");
      return;
    }
    if (ends("us") || ends("ss")) return;
    if (ends("s")) {
      const size_t before_s = w_.size() - 1;
      if (before_s == 0) return;
      if (has_vowel_before(prev(before_s))) w_.resize(before_s);
    }
  }

  void step_1b() {
    size_t n = 0;
    if (ends("eedly")) n = 5;
    else if (ends("eed")) n = 3;
    if (n != 0) {
      if (in_r1(n)) replace_suffix(n, "ee");
      return;
    }
    if (ends("ingly")) n = 5;
    else if (ends("edly")) n = 4;
    else if (ends("ing")) n = 3;
    else if (ends("ed")) n = 2;
    if (n == 0) return;
    if (!has_vowel_before(w_.size() - n)) return;
    w_.resize(w_.size() - n);
    if (ends("at") || ends("bl") || ends("iz")) { w_.push_back('e'); return; }
    if (w_.size() >= 2) {
      const char a = w_[w_.size() - 1], b = w_[w_.size() - 2];
      if (a == b && a != '\0' && std::strchr("bdfgmnprt", a) != nullptr) { w_.pop_back(); return; }
    }
    if (w_.size() == p1_ && short_syllable(w_.size())) w_.push_back('e');
  }

  void step_1c() {
    const size_t n = w_.size();
    if (n < 2 || (w_[n - 1] != 'y' && w_[n - 1] != 'Y')) return;
    const size_t c = prev(n - 1);
    if (vowel_at(c) || c == 0) return;
    w_[n - 1] = 'i';
  }

  struct Rule { const char* suffix; const char* by; };
  // longest matching suffix of a table sorted by decreasing length; nullptr when none matches
  template <size_t N>
  const Rule* longest(const Rule (&rules)[N]) const {
    for (const Rule& r : rules)
      if (ends(r.suffix)) return &r;
    return nullptr;
  }

  void step_2() {
    static const Rule kRules[] = {
        {"ization", "ize"}, {"ational", "ate"}, {"fulness", "ful"}, {"ousness", "ous"}, {"iveness", "ive"}, {"tional", "tion"},
        {"biliti", "ble"},  {"lessli", "less"}, {"entli", "ent"},   {"ation", "ate"},   {"alism", "al"},    {"aliti", "al"},
        {"ousli", "ous"},   {"iviti", "ive"},   {"fulli", "ful"},   {"enci", "ence"},   {"anci", "ance"},   {"abli", "able"},
        {"izer", "ize"},    {"ator", "ate"},    {"alli", "al"},     {"bli", "ble"},     {"ogi", nullptr},   {"li", nullptr}};
    const Rule* r = longest(kRules);
    if (r == nullptr) return;
    const size_t n = std::strlen(r->suffix);
    if (!in_r1(n)) return;
    if (r->by != nullptr) { replace_suffix(n, r->by); return; }
    const size_t at = w_.size() - n;
    if (n == 3) {                                     // ogi: preceded by l
      if (at > 0 && w_[at - 1] == 'l') replace_suffix(3, "og");
    } else {                                          // li: preceded by a valid li-ending
      if (at > 0 && w_[at - 1] != '\0' && std::strchr("cdeghkmnrt", w_[at - 1]) != nullptr) w_.resize(at);
    }
  }

  void step_3() {
    static const Rule kRules[] = {{"ational", "ate"}, {"tional", "tion"}, {"alize", "al"}, {"icate", "ic"}, {"iciti", "ic"},
                                  {"ative", nullptr}, {"ical", "ic"},     {"ness", ""},    {"ful", ""}};
    const Rule* r = longest(kRules);
    if (r == nullptr) return;
    const size_t n = std::strlen(r->suffix);
    if (!in_r1(n)) return;
    if (r->by != nullptr) replace_suffix(n, r->by);
    else if (in_r2(n)) w_.resize(w_.size() - n);
  }

  void step_4() {
    static const Rule kRules[] = {{"ement", ""}, {"ance", ""}, {"ence", ""}, {"able", ""}, {"ible", ""}, {"ment", ""},
                                  {"ant", ""},   {"ent", ""},  {"ism", ""},  {"ate", ""},  {"iti", ""},  {"ous", ""},
                                  {"ive", ""},   {"ize", ""},  {"ion", nullptr}, {"al", ""}, {"er", ""}, {"ic", ""}};
    const Rule* r = longest(kRules);
    if (r == nullptr) return;
    const size_t n = std::strlen(r->suffix);
    if (!in_r2(n)) return;
    if (r->by == nullptr) {
      const size_t at = w_.size() - n;
      if (at == 0 || (w_[at - 1] != 's' && w_[at - 1] != 't')) return;
    }
    w_.resize(w_.size() - n);
  }

  void step_5() {
    if (w_.empty()) return;
    const char last = w_.back();
    if (last == 'e') {
      if (in_r2(1) || (in_r1(1) && !short_syllable(w_.size() - 1))) w_.pop_back();
    } else if (last == 'l') {
      if (in_r2(1) && w_.size() >= 2 && w_[w_.size() - 2] == 'l') w_.pop_back();
    }
  }
};

// bm25s' token pattern r"(?u)\b\w\w+\b" on ASCII text = maximal runs of [0-9A-Za-z_] of length >= 2.
// Returns the number of tokens found (which may exceed cap; only the first cap are written), or -1 when the text
// holds a byte >= 0x80 (the caller then uses the Unicode-aware regex on the host).
inline int64_t tokenize_ascii(const char* text, int64_t len, int64_t* starts, int64_t* lens, int64_t cap) {
  auto word = [](unsigned char c) { return (c >= '0' && c <= '9') || (c >= 'a' && c <= 'z') || (c >= 'A' && c <= 'Z') || c == '_'; };
  int64_t n = 0, i = 0;
  while (i < len) {
    const unsigned char c = static_cast<unsigned char>(text[i]);
    if (c >= 0x80) return -1;
    if (!word(c)) { ++i; continue; }
    int64_t j = i + 1;
    while (j < len && word(static_cast<unsigned char>(text[j]))) ++j;
    if (j < len && static_cast<unsigned char>(text[j]) >= 0x80) return -1;
    if (j - i >= 2) {
      if (n < cap) { starts[n] = i; lens[n] = j - i; }
      ++n;
    }
    i = j;
  }
  return n;
}

}  // namespace vfi_text
