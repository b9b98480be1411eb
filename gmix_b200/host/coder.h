// Binary arithmetic coder used with the Predictor facade on the host: 32-bit carry-less range coder
// with 16-bit probabilities, behaviour of the reference's Encoder/Decoder (reference
// src/coder/encoder.cpp:8-34, src/coder/decoder.cpp:3-39). The batch kernels carry their own copy on
// the device; this one exists so that host code written against Predictor can still code streams.
#ifndef GMIX_B200_HOST_CODER_H_
#define GMIX_B200_HOST_CODER_H_
#include <stdint.h>
#include <string.h>

#include <fstream>
#include <string>
#include <vector>

namespace gmixb {

inline uint32_t Discretize(float p) {  // encoder.cpp:8: 1 + 65534 * p, truncated
  const float scaled = 65534.0f * p;
  return (uint32_t)(1.0f + scaled);
}

inline uint32_t Split(uint32_t x1, uint32_t x2, uint32_t p16) {
  const uint32_t range = x2 - x1;
  return x1 + (range >> 16) * p16 + (((range & 0xffff) * p16) >> 16);
}

class Encoder {
 public:
  explicit Encoder(std::vector<uint8_t>* out) : out_(out) {}
  void Encode(int bit, float p) {
    const uint32_t mid = Split(x1_, x2_, Discretize(p));
    if (bit) x2_ = mid; else x1_ = mid + 1;
    Normalize();
  }
  void Flush() { Normalize(); out_->push_back((uint8_t)(x2_ >> 24)); }  // encoder.cpp:27-34
  // encoder.cpp:36-51: the file holds x1, x2 as raw little-endian u32 (a file that cannot be opened is silently ignored)
  void WriteCheckpoint(const std::string& path) const {
    std::ofstream f(path, std::ios::out | std::ios::binary);
    if (!f.is_open()) return;
    f.write((const char*)&x1_, 4); f.write((const char*)&x2_, 4);
  }
  void ReadCheckpoint(const std::string& path) {
    std::ifstream f(path, std::ios::in | std::ios::binary);
    if (!f.is_open()) return;
    f.read((char*)&x1_, 4); f.read((char*)&x2_, 4);
  }
  uint32_t x1() const { return x1_; }
  uint32_t x2() const { return x2_; }
  void Set(uint32_t x1, uint32_t x2) { x1_ = x1; x2_ = x2; }

 private:
  void Normalize() {
    while (((x1_ ^ x2_) & 0xff000000u) == 0) { out_->push_back((uint8_t)(x2_ >> 24)); x1_ <<= 8; x2_ = (x2_ << 8) + 255; }
  }
  std::vector<uint8_t>* out_;
  uint32_t x1_ = 0, x2_ = 0xffffffffu;
};

class Decoder {
 public:
  Decoder(const uint8_t* in, uint64_t n) : in_(in), n_(n) {
    for (int i = 0; i < 4; ++i) x_ = (x_ << 8) + Next();
  }
  int Decode(float p) {  // decoder.cpp:19-39 (the caller runs Predict/Perceive/Learn around it)
    const uint32_t mid = Split(x1_, x2_, Discretize(p));
    int bit = 0;
    if (x_ <= mid) { bit = 1; x2_ = mid; } else x1_ = mid + 1;
    while (((x1_ ^ x2_) & 0xff000000u) == 0) { x1_ <<= 8; x2_ = (x2_ << 8) + 255; x_ = (x_ << 8) + Next(); }
    return bit;
  }
  // decoder.cpp:41-57: x1, x2, x as raw little-endian u32. The read position of the coded stream is the caller's
  // business, as in the reference (its decoder reads from a std::ifstream that simply keeps its position).
  void WriteCheckpoint(const std::string& path) const {
    std::ofstream f(path, std::ios::out | std::ios::binary);
    if (!f.is_open()) return;
    f.write((const char*)&x1_, 4); f.write((const char*)&x2_, 4); f.write((const char*)&x_, 4);
  }
  void ReadCheckpoint(const std::string& path) {
    std::ifstream f(path, std::ios::in | std::ios::binary);
    if (!f.is_open()) return;
    f.read((char*)&x1_, 4); f.read((char*)&x2_, 4); f.read((char*)&x_, 4);
  }
  uint64_t position() const { return pos_; }       // coded bytes consumed so far
  void Seek(uint64_t pos) { pos_ = pos; }

 private:
  uint32_t Next() { return pos_ < n_ ? in_[pos_++] : (pos_++, 0u); }  // reads as 0 past the end
  const uint8_t* in_;
  uint64_t n_, pos_ = 0;
  uint32_t x1_ = 0, x2_ = 0xffffffffu, x_ = 0;
};

}  // namespace gmixb
#endif
