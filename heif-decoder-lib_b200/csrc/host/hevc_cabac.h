// hevc_cabac.h — CABAC arithmetic decoding engine + I-slice context set (H.265 §9.3).
// New implementation: the 9-bit arithmetic offset and up to 47 look-ahead bits share one 64-bit
// register, so renormalisation is a counter decrement and the input is refilled 32 bits at a time;
// fixed-length bypass strings are decoded by one division. Tables 9-5..9-37 (context
// initialisation, initType 0 only: this front-end decodes I slices) and 9-46/9-47 (state
// transition, LPS range) are normative data.
// Reference counterpart: third-party/libde265/libde265/cabac.cc:181-655, contextmodel.cc:234-354.
#pragma once
#include <cstdint>
#include <cstddef>
#include <cstring>

namespace hc {

// Context layout. CBF_CHROMA has 4 entries directly followed by SPLIT_TRANSFORM so that
// cbf_cb/cbf_cr at trafoDepth 4 (4:4:4 only) aliases split_transform_flag[0] exactly like the
// reference's table does (contextmodel.h:58-60).
enum CtxIdx {
  CTX_SAO_MERGE = 0,
  CTX_SAO_TYPE = CTX_SAO_MERGE + 1,
  CTX_SPLIT_CU = CTX_SAO_TYPE + 1,
  CTX_TQ_BYPASS = CTX_SPLIT_CU + 3,
  CTX_PART_MODE = CTX_TQ_BYPASS + 1,
  CTX_PREV_INTRA_LUMA = CTX_PART_MODE + 1,
  CTX_INTRA_CHROMA = CTX_PREV_INTRA_LUMA + 1,
  CTX_CBF_LUMA = CTX_INTRA_CHROMA + 1,
  CTX_CBF_CHROMA = CTX_CBF_LUMA + 2,
  CTX_SPLIT_TRANSFORM = CTX_CBF_CHROMA + 4,
  CTX_CU_QP_DELTA = CTX_SPLIT_TRANSFORM + 3,
  CTX_TSKIP = CTX_CU_QP_DELTA + 2,
  CTX_LAST_X = CTX_TSKIP + 2,
  CTX_LAST_Y = CTX_LAST_X + 18,
  CTX_CSBF = CTX_LAST_Y + 18,
  CTX_SIG = CTX_CSBF + 4,
  CTX_G1 = CTX_SIG + 44,
  CTX_G2 = CTX_G1 + 24,
  CTX_CHROMA_QP_OFFSET_FLAG = CTX_G2 + 6,
  CTX_CHROMA_QP_OFFSET_IDX = CTX_CHROMA_QP_OFFSET_FLAG + 1,
  CTX_RES_SCALE_ABS = CTX_CHROMA_QP_OFFSET_IDX + 1,
  CTX_RES_SCALE_SIGN = CTX_RES_SCALE_ABS + 8,
  CTX_COUNT = CTX_RES_SCALE_SIGN + 2
};

struct CtxSet {
  uint8_t s[CTX_COUNT];  // (pStateIdx << 1) | valMps
};

namespace detail {
static const uint8_t kRangeLps[64][4] = {
    {128, 176, 208, 240}, {128, 167, 197, 227}, {128, 158, 187, 216}, {123, 150, 178, 205},
    {116, 142, 169, 195}, {111, 135, 160, 185}, {105, 128, 152, 175}, {100, 122, 144, 166},
    {95, 116, 137, 158},  {90, 110, 130, 150},  {85, 104, 123, 142},  {81, 99, 117, 135},
    {77, 94, 111, 128},   {73, 89, 105, 122},   {69, 85, 100, 116},   {66, 80, 95, 110},
    {62, 76, 90, 104},    {59, 72, 86, 99},     {56, 69, 81, 94},     {53, 65, 77, 89},
    {51, 62, 73, 85},     {48, 59, 69, 80},     {46, 56, 66, 76},     {43, 53, 63, 72},
    {41, 50, 59, 69},     {39, 48, 56, 65},     {37, 45, 54, 62},     {35, 43, 51, 59},
    {33, 41, 48, 56},     {32, 39, 46, 53},     {30, 37, 43, 50},     {29, 35, 41, 48},
    {27, 33, 39, 45},     {26, 31, 37, 43},     {24, 30, 35, 41},     {23, 28, 33, 39},
    {22, 27, 32, 37},     {21, 26, 30, 35},     {20, 24, 29, 33},     {19, 23, 27, 31},
    {18, 22, 26, 30},     {17, 21, 25, 28},     {16, 20, 23, 27},     {15, 19, 22, 25},
    {14, 18, 21, 24},     {14, 17, 20, 23},     {13, 16, 19, 22},     {12, 15, 18, 21},
    {12, 14, 17, 20},     {11, 14, 16, 19},     {11, 13, 15, 18},     {10, 12, 15, 17},
    {10, 12, 14, 16},     {9, 11, 13, 15},      {9, 11, 12, 14},      {8, 10, 12, 14},
    {8, 9, 11, 13},       {7, 9, 11, 12},       {7, 9, 10, 12},       {7, 8, 10, 11},
    {6, 8, 9, 11},        {6, 7, 9, 10},        {6, 7, 8, 9},         {2, 2, 2, 2}};
struct Transitions {
  uint8_t mps[128], lps[128];
  uint8_t next[128][2];   // [state][is_lps]
  constexpr Transitions() : mps(), lps(), next() {
    constexpr uint8_t next_mps[64] = {1,  2,  3,  4,  5,  6,  7,  8,  9,  10, 11, 12, 13, 14, 15, 16, 17, 18, 19, 20, 21, 22,
                                      23, 24, 25, 26, 27, 28, 29, 30, 31, 32, 33, 34, 35, 36, 37, 38, 39, 40, 41, 42, 43, 44,
                                      45, 46, 47, 48, 49, 50, 51, 52, 53, 54, 55, 56, 57, 58, 59, 60, 61, 62, 62, 63};
    constexpr uint8_t next_lps[64] = {0,  0,  1,  2,  2,  4,  4,  5,  6,  7,  8,  9,  9,  11, 11, 12, 13, 13, 15, 15, 16, 16,
                                      18, 18, 19, 19, 21, 21, 22, 22, 23, 24, 24, 25, 26, 26, 27, 27, 28, 29, 29, 30, 30, 30,
                                      31, 32, 32, 33, 33, 33, 34, 34, 35, 35, 35, 36, 36, 36, 37, 37, 37, 38, 38, 63};
    for (int st = 0; st < 64; st++)
      for (int m = 0; m < 2; m++) {
        mps[(st << 1) | m] = (uint8_t)((next_mps[st] << 1) | m);
        int nm = (st == 0) ? 1 - m : m;
        lps[(st << 1) | m] = (uint8_t)((next_lps[st] << 1) | nm);
        next[(st << 1) | m][0] = mps[(st << 1) | m];
        next[(st << 1) | m][1] = lps[(st << 1) | m];
      }
  }
};
static constexpr Transitions kTransitions{};
}  // namespace detail

inline uint8_t ctx_init_state(int init_value, int slice_qp) {
  int slope = init_value >> 4, offs = init_value & 15;
  int m = slope * 5 - 45, n = (offs << 3) - 16;
  int q = slice_qp < 0 ? 0 : (slice_qp > 51 ? 51 : slice_qp);
  int pre = ((m * q) >> 4) + n;
  pre = pre < 1 ? 1 : (pre > 126 ? 126 : pre);
  int mps = pre <= 63 ? 0 : 1;
  int st = mps ? pre - 64 : 63 - pre;
  return (uint8_t)((st << 1) | mps);
}

// initValue of every context (Tables 9-5..9-37, initType 0)
inline void ctx_init_values(uint8_t v[CTX_COUNT]) {
  auto set = [&](int first, const int* src, int n) {
    for (int i = 0; i < n; i++) v[first + i] = (uint8_t)src[i];
  };
  static const int sao_merge[1] = {153}, sao_type[1] = {200}, split_cu[3] = {139, 141, 157};
  static const int bypass[1] = {154}, part_mode[1] = {184}, prev_luma[1] = {184}, chroma_mode[1] = {63};
  static const int cbf_luma[2] = {111, 141}, cbf_chroma[4] = {94, 138, 182, 154};
  static const int split_tr[3] = {153, 138, 138}, qp_delta[2] = {154, 154}, tskip[2] = {139, 139};
  static const int last[18] = {110, 110, 124, 125, 140, 153, 125, 127, 140, 109, 111, 143, 127, 111, 79, 108, 123, 63};
  static const int csbf[4] = {91, 171, 134, 141};
  static const int sig[44] = {111, 111, 125, 110, 110, 94,  124, 108, 124, 107, 125, 141, 179, 153, 125,
                              107, 125, 141, 179, 153, 125, 107, 125, 141, 179, 153, 125, 140, 139, 182,
                              182, 152, 136, 152, 136, 153, 136, 139, 111, 136, 139, 111, /*skip-mode*/ 141, 111};
  static const int g1[24] = {140, 92,  137, 138, 140, 152, 138, 139, 153, 74,  149, 92,
                             139, 107, 122, 152, 140, 179, 166, 182, 140, 227, 122, 197};
  static const int g2[6] = {138, 153, 136, 167, 152, 152};
  static const int c154[8] = {154, 154, 154, 154, 154, 154, 154, 154};
  set(CTX_SAO_MERGE, sao_merge, 1);
  set(CTX_SAO_TYPE, sao_type, 1);
  set(CTX_SPLIT_CU, split_cu, 3);
  set(CTX_TQ_BYPASS, bypass, 1);
  set(CTX_PART_MODE, part_mode, 1);
  set(CTX_PREV_INTRA_LUMA, prev_luma, 1);
  set(CTX_INTRA_CHROMA, chroma_mode, 1);
  set(CTX_CBF_LUMA, cbf_luma, 2);
  set(CTX_CBF_CHROMA, cbf_chroma, 4);
  set(CTX_SPLIT_TRANSFORM, split_tr, 3);
  set(CTX_CU_QP_DELTA, qp_delta, 2);
  set(CTX_TSKIP, tskip, 2);
  set(CTX_LAST_X, last, 18);
  set(CTX_LAST_Y, last, 18);
  set(CTX_CSBF, csbf, 4);
  set(CTX_SIG, sig, 44);
  set(CTX_G1, g1, 24);
  set(CTX_G2, g2, 6);
  set(CTX_CHROMA_QP_OFFSET_FLAG, c154, 1);
  set(CTX_CHROMA_QP_OFFSET_IDX, c154, 1);
  set(CTX_RES_SCALE_ABS, c154, 8);
  set(CTX_RES_SCALE_SIGN, c154, 2);
}

inline void ctx_init_all(CtxSet& c, int slice_qp) {
  uint8_t v[CTX_COUNT];
  ctx_init_values(v);
  for (int i = 0; i < CTX_COUNT; i++) c.s[i] = ctx_init_state(v[i], slice_qp);
}

// Arithmetic decoder. `value` = (ivlOffset << avail) | the next `avail` bits of the stream, so
// "ivlOffset < ivlCurrRange" is `value < (range << avail)` and a renormalisation shift by n is
// `avail -= n`. Invariant between calls: 16 <= avail <= 47 (every call may consume up to 16 bits).
// The input must be readable for 8 bytes past `end` (the parser pads its RBSP buffer).
struct Cabac {
  const uint8_t* start = nullptr;
  const uint8_t* cur = nullptr;   // next byte to load
  const uint8_t* end = nullptr;
  uint64_t value = 0;
  uint32_t range = 510;
  int avail = 0;

  static inline uint32_t load_be32(const uint8_t* p) {
    uint32_t v;
    memcpy(&v, p, 4);
    return __builtin_bswap32(v);
  }
  inline void refill() {
    if (avail < 16) {
      value = (value << 32) | load_be32(cur);
      cur += 4;
      avail += 32;
    }
  }

  void init(const uint8_t* p, const uint8_t* e) {
    start = p;
    end = e;
    range = 510;
    value = load_be32(p);
    cur = p + 4;
    avail = 32 - 9;
  }

  // Byte position "after the last byte fetched" of a decoder that loads 2 bytes at start-up and
  // one more every 8 renormalisation shifts: where the next substream / the PCM samples begin
  // (H.265 9.3.2.5; the reference uses its fetch pointer the same way, slice.cc:5306-5312).
  inline const uint8_t* position() const {
    const long shifts = (long)(cur - start) * 8 - 9 - avail;
    return start + 2 + (shifts >> 3);
  }
  inline bool overrun() const { return position() > end; }

  // Branch-free bin decode: the MPS/LPS outcome of a well-used context is close to a coin flip, so
  // it is turned into a mask instead of a (mispredicted) branch. Both paths renormalise by
  // clz(range) - 23 (MPS: range >= 128 after the subtraction, i.e. 0 or 1 shifts).
  inline int decode_bin(uint8_t& state) {
    const uint32_t st = state;
    const uint32_t lps = detail::kRangeLps[st >> 1][(range >> 6) & 3];
    const uint32_t rmps = range - lps;
    const uint64_t scaled = (uint64_t)rmps << avail;
    const uint32_t is_lps = value >= scaled;                 // 0 / 1
    const uint32_t mask = 0u - is_lps;
    value -= scaled & (uint64_t)(int64_t)(int32_t)mask;
    const uint32_t r = rmps + ((lps - rmps) & mask);
    const int n = __builtin_clz(r) - 23;
    range = r << n;
    avail -= n;
    state = detail::kTransitions.next[st][is_lps];
    refill();
    return (int)((st & 1) ^ is_lps);
  }

  inline int decode_bypass() {
    avail--;
    const uint64_t scaled = (uint64_t)range << avail;
    int bin = 0;
    if (value >= scaled) {
      value -= scaled;
      bin = 1;
    }
    refill();
    return bin;
  }

  // n <= 16 bypass bins at once: the repeated compare/subtract of n bypass decodes is the long division
  // of the top (9 + n) bits of `value` by the range (a 32-bit division); the quotient bits are the bins
  inline uint32_t decode_bypass_bits(int n) {
    uint32_t out = 0;
    while (n > 0) {
      const int k = n > 16 ? 16 : n;
      avail -= k;
      const uint32_t q = (uint32_t)(value >> avail) / range;
      value -= (uint64_t)(q * range) << avail;
      out = (out << k) | q;
      refill();
      n -= k;
    }
    return out;
  }
  // The next 16 bypass bins without consuming them (bit 15 = first bin) ...
  inline uint32_t peek_bypass16() const { return (uint32_t)(value >> (avail - 16)) / range; }
  // ... and the commit of the first `n` of them (`bins` = peek >> (16 - n))
  inline void consume_bypass(int n, uint32_t bins) {
    avail -= n;
    value -= (uint64_t)(bins * range) << avail;
    refill();
  }

  inline int decode_terminate() {
    range -= 2;
    const uint64_t scaled = (uint64_t)range << avail;
    if (value >= scaled) return 1;
    if (range < 256) {
      range <<= 1;
      avail--;
    }
    refill();
    return 0;
  }
};

}  // namespace hc
