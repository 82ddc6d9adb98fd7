// hevc_bits.h — NAL payload unescaping and an MSB-first bit reader for parameter sets / slice headers.
// Host front-end of the B200 HEVC-intra engine (new code, written from ITU-T H.265 §7.2/§7.3/§9.2).
// Reference counterpart: third-party/libde265/libde265/nal-parser.cc:116-140 (remove_stuffing_bytes),
// bitstream.cc (get_bits/get_uvlc/get_svlc).
#pragma once
#include <cstring>
#include <cstdint>
#include <cstddef>
#include <vector>

namespace hc {

// Removes emulation-prevention bytes (00 00 03 -> 00 00). `skipped` receives, for every removed
// byte, its position in the *escaped* input; slice-header entry points are given in escaped bytes
// and are corrected with it (reference: decctx.cc:674-680).
inline void nal_unescape(const uint8_t* in, size_t n, std::vector<uint8_t>& out,
                         std::vector<uint32_t>* skipped) {
  // emulation prevention bytes are rare: find candidate 0x03 bytes with memchr and copy the spans between them
  out.resize(n);
  uint8_t* dst = out.data();
  size_t from = 0, i = 2;
  while (i < n) {
    const uint8_t* hit = (const uint8_t*)memchr(in + i, 3, n - i);
    if (!hit) break;
    i = (size_t)(hit - in);
    if (in[i - 1] == 0 && in[i - 2] == 0) {
      memcpy(dst, in + from, i - from);
      dst += i - from;
      if (skipped) skipped->push_back((uint32_t)i);
      from = i + 1;
      i += 3;      // the two bytes after an escape cannot complete another 00 00 03 with it
    } else {
      i++;
    }
  }
  memcpy(dst, in + from, n - from);
  dst += n - from;
  out.resize((size_t)(dst - out.data()));
}

struct BitReader {
  const uint8_t* p = nullptr;
  size_t size = 0;     // bytes
  size_t pos = 0;      // bit position
  bool overrun = false;

  BitReader() {}
  BitReader(const uint8_t* data, size_t n) : p(data), size(n) {}

  size_t bits_left() const { return size * 8 > pos ? size * 8 - pos : 0; }

  uint32_t u(int n) {  // n <= 32
    uint32_t v = 0;
    for (int i = 0; i < n; i++) {
      size_t byte = pos >> 3;
      uint32_t bit = 0;
      if (byte < size) bit = (p[byte] >> (7 - (pos & 7))) & 1u;
      else overrun = true;
      v = (v << 1) | bit;
      pos++;
    }
    return v;
  }
  uint32_t flag() { return u(1); }
  void skip(int n) { pos += n; if (pos > size * 8) overrun = true; }

  // Exp-Golomb. Returns UINT32_MAX on malformed input (more than 31 leading zeros).
  uint32_t ue() {
    int zeros = 0;
    while (u(1) == 0) {
      zeros++;
      if (zeros > 31 || overrun) { overrun = true; return 0xFFFFFFFFu; }
    }
    if (zeros == 0) return 0;
    return ((1u << zeros) - 1) + u(zeros);
  }
  int32_t se() {
    uint32_t k = ue();
    if (k == 0xFFFFFFFFu) return 0;
    return (k & 1) ? (int32_t)((k + 1) >> 1) : -(int32_t)(k >> 1);
  }
  bool byte_aligned() const { return (pos & 7) == 0; }
  size_t byte_pos() const { return (pos + 7) >> 3; }
};

inline int ceil_log2(uint32_t v) {
  int r = 0;
  while ((1u << r) < v) r++;
  return r;
}

}  // namespace hc
