// Dictionary preprocessing: the reference's offline word-replacing text transform (src/preprocess/dictionary.{h,cpp},
// tool src/runner/dictionary-prep.cpp). Host-side only - it runs once over a file before compression and
// after decompression, never on the per-bit path - and byte-compatible with the reference in both directions: a file encoded
// here decodes with the reference's tool and vice versa (tests/test_dictionary_prep.py).
//
// Format (dictionary.cpp:11-37, 41-75): the dictionary lists lower-case words by rank; rank r gets a 1-, 2- or 3-byte code whose
// bytes are all >= 0x80 (first byte 0x80-0xCF: one byte; 0xD0-0xEF / 0xF0-0xFF lead longer codes, a second byte > 0xCF announces a
// third). In the text a maximal run [A-Z]*[a-z]* that is lower case, Capitalized or UPPER is one word: its code (or, for words of
// 8+ letters, the code of a dictionary word that is its suffix or prefix of 7+ letters next to the remaining letters, or the bare
// letters) behind a case marker (0x40 Capitalized, 0x07 UPPER, closed by 0x06 when a lower-case letter follows directly). Bytes
// that collide with the markers or are >= 0x80 are escaped with 0x0C; the six characters `&quot;` become `&` + 0x08.
#ifndef GMIX_B200_HOST_DICTIONARY_H_
#define GMIX_B200_HOST_DICTIONARY_H_
#include <stdint.h>

#include <string>
#include <unordered_map>
#include <vector>

namespace gmixb {

class WordTransform {
 public:
  // dictionary file bytes: words are the maximal runs of [a-z], everything else separates (dictionary.cpp:48-74)
  explicit WordTransform(const std::vector<uint8_t>& dictionary) {
    std::string w;
    uint32_t rank = 0;
    auto close = [&]() {
      if (w.empty()) return;
      if (w.size() > longest_) longest_ = w.size();
      const uint32_t code = CodeOfRank(rank++);
      code_of_[w] = code;        // a repeated word keeps its LAST code for encoding ...
      word_of_[code] = w;        // ... and every code decodes (ranks past the table share code 0, as in the reference)
      w.clear();
    };
    for (uint8_t c : dictionary) {
      if (c >= 'a' && c <= 'z') w += (char)c; else close();
    }
    // (a word that runs into the end of the file is not registered: the reference only closes a word on a separator)
  }

  std::vector<uint8_t> Encode(const std::vector<uint8_t>& in) const {
    Encoder e{*this};
    for (size_t i = 0; i < in.size(); ++i) e.Push(in[i], i + 1 == in.size());
    return std::move(e.out);
  }

  std::vector<uint8_t> Decode(const std::vector<uint8_t>& in) const {
    std::vector<uint8_t> out;
    bool upper = false, capital = false;
    size_t i = 0;
    auto next = [&]() -> uint8_t { return i < in.size() ? in[i++] : (uint8_t)0xFF; };   // getc() == EOF stored in a byte
    while (i < in.size()) {
      uint8_t c = in[i++];
      if (c == kEscape) { upper = false; out.push_back(next()); }
      else if (c == kQuote) out.insert(out.end(), {'q', 'u', 'o', 't', ';'});
      else if (c == kUpper) upper = true;
      else if (c == kCapital) capital = true;
      else if (c == kEndUpper) upper = false;
      else if (c >= 0x80) {
        uint32_t code = c;
        if (c > 0xCF) {
          c = next(); code += (uint32_t)c << 8;
          if (c > 0xCF) { c = next(); code += (uint32_t)c << 16; }
        }
        auto it = word_of_.find(code);
        if (it == word_of_.end()) continue;
        for (size_t k = 0; k < it->second.size(); ++k) {
          uint8_t ch = (uint8_t)it->second[k];
          if (k == 0 && capital) { ch = (uint8_t)(ch - 'a' + 'A'); capital = false; }
          if (upper) ch = (uint8_t)(ch - 'a' + 'A');   // (after a capital first letter this shifts once more, as the reference does)
          out.push_back(ch);
        }
      } else {
        const bool letter = (c >= 'a' && c <= 'z') || (c >= 'A' && c <= 'Z');
        if (!letter) upper = false;
        if (capital || upper) c = (uint8_t)(c - 'a' + 'A');
        capital = false;
        out.push_back(c);
      }
    }
    return out;
  }

  size_t words() const { return code_of_.size(); }

 private:
  static constexpr uint8_t kCapital = 0x40, kUpper = 0x07, kEndUpper = 0x06, kEscape = 0x0C, kQuote = 0x08;

  // rank -> code bytes, least significant byte first in the stream (dictionary.cpp:45-68)
  static uint32_t CodeOfRank(uint32_t r) {
    const uint32_t b1 = 80, b2 = b1 + 3840, b3 = b2 + 40960, b4 = b3 + 81920;
    if (r < b1) return 0x80 + r;
    if (r < b2) return (0xD0 + (r - b1) / 80) + ((0x80 + (r - b1) % 80) << 8);
    const uint32_t q = (r - b2) / 80, low = (0x80 + (r - b2) % 80) << 16;
    if (r < b3) return (0xF0 + q / 32) + ((0xD0 + q % 32) << 8) + low;
    if (r < b4) return (0xD0 + q / 32) + ((0xD0 + q % 32) << 8) + low;
    return 0;   // the reference leaves `bytes` unset past the fourth boundary; no dictionary in use is that long
  }

  struct Encoder {
    const WordTransform& T;
    std::vector<uint8_t> out;
    std::string word;            // the run so far, lower-cased
    int upper = 0, lower = 0;    // its upper- / lower-case letters
    int quote = 0;               // matched prefix of "&quot;"

    void Raw(uint8_t c) {
      if (c == kEndUpper || c == kEscape || c == kUpper || c == kCapital || c == kQuote || c >= 0x80) out.push_back(kEscape);
      out.push_back(c);
    }
    void Code(uint32_t code) {
      out.push_back((uint8_t)code);
      if (!(code & 0xFF00)) return;
      out.push_back((uint8_t)(code >> 8));
      if (code & 0xFF0000) out.push_back((uint8_t)(code >> 16));
    }
    void Letters(const std::string& s, size_t from, size_t to) { for (size_t i = from; i < to; ++i) out.push_back((uint8_t)s[i]); }
    // a dictionary word of 7+ letters as proper suffix, else as proper prefix, of a word of 8+ letters (dictionary.cpp:158-191)
    bool Affix() {
      if (word.size() <= 7) return false;
      size_t n = word.size() - 1;
      if (n > T.longest_) n = T.longest_;
      for (size_t len = n; len >= 7; --len) {
        auto it = T.code_of_.find(word.substr(word.size() - len));
        if (it != T.code_of_.end()) { Letters(word, 0, word.size() - len); Code(it->second); return true; }
      }
      for (size_t len = n; len >= 7; --len) {
        auto it = T.code_of_.find(word.substr(0, len));
        if (it != T.code_of_.end()) { Code(it->second); Letters(word, len, word.size()); return true; }
      }
      return false;
    }
    void Flush(bool next_lower) {
      if (upper > 1) out.push_back(kUpper); else if (upper == 1) out.push_back(kCapital);
      auto it = T.code_of_.find(word);
      if (it != T.code_of_.end()) Code(it->second);
      else if (!Affix()) Letters(word, 0, word.size());
      if (upper > 1 && next_lower) out.push_back(kEndUpper);
      word.clear(); upper = lower = 0;
    }
    void Push(uint8_t c, bool last) {
      static const char kQuoteStr[] = "&quot;";
      if (c == (uint8_t)kQuoteStr[quote]) {
        if (++quote == 6) {   // the '&' went out as itself, the letters q-u-o-t gathered since are dropped
          out.push_back(kQuote);
          word.clear(); upper = lower = 0;
          return;             // (quote stays 6: kQuoteStr[6] is the terminator, the next byte resets it unless it is 0)
        }
      } else {
        quote = 0;
      }
      const bool is_lower = c >= 'a' && c <= 'z', is_upper = c >= 'A' && c <= 'Z';
      // does c extend the run? lower case after at most one capital, or upper case while no lower-case letter has been seen
      const bool extends = word.size() <= T.longest_ && ((is_lower && upper <= 1) || (is_upper && lower == 0));
      if (extends) {
        if (is_lower) { ++lower; word += (char)c; } else { ++upper; word += (char)(c - 'A' + 'a'); }
        if (last) Flush(false);
        return;
      }
      if (word.empty()) { Raw(c); return; }
      Flush(is_lower);
      if (is_lower) { ++lower; word += (char)c; }
      else if (is_upper) { ++upper; word += (char)(c - 'A' + 'a'); }
      else Raw(c);
      if (last && !word.empty()) Flush(false);
    }
  };

  std::unordered_map<std::string, uint32_t> code_of_;
  std::unordered_map<uint32_t, std::string> word_of_;
  size_t longest_ = 0;
};

}  // namespace gmixb
#endif
