// k1_core.cuh — per-thread arithmetic of K1 (dequantisation + 1-D inverse transforms).
//
// Register-resident partial butterflies: a thread holds one column (pass 1) or one row (pass 2)
// of a transform block in registers and runs an N-point inverse transform on it. The even/odd
// decomposition is exact integer arithmetic, so results equal the reference's plain matrix
// products (fallback-dct.cc:593-733) bit for bit.  All loops have compile-time bounds and are
// fully unrolled; every matrix coefficient becomes an immediate operand of an IMAD.
//
// Reference semantics: third-party/libde265/libde265/transform.cc:473-546 (dequantisation),
// fallback-dct.cc:309-378 (DST 4x4), :593-733 (DCT), :84-108 (transform skip).
#pragma once
#include "common.cuh"

namespace hc {

// First column of the standard's 32x32 matrix (k -> T[k][0]); every other entry follows from the
// cosine symmetries, see dct_coef().
HC_HD constexpr int dct_col0(int k) {
  constexpr int t[32] = {64, 90, 90, 90, 89, 88, 87, 85, 83, 82, 80, 78, 75, 73, 70, 67,
                         64, 61, 57, 54, 50, 46, 43, 38, 36, 31, 25, 22, 18, 13, 9,  4};
  return t[k];
}

// T32[k][n]: row k (frequency), column n (sample) of the 32-point matrix.
HC_HD constexpr int dct_coef(int k, int n) {
  if (k == 0) return 64;
  int a = (k * (2 * n + 1)) % 128;  // angle in units of pi/64
  int s = 1;
  if (a > 64) a = 128 - a;
  if (a > 32) { a = 64 - a; s = -1; }
  return a == 32 ? 0 : s * dct_col0(a);
}

// N-point inverse DCT of in[0..N) -> out[0..N):  out[i] = sum_j T_N[j][i] * in[j],
// with T_N[j][i] = T32[j * (32/N)][i]. STEP is the row stride into T32 at this recursion level.
template <int N, int STEP>
struct InvDct {
  HC_HD static void run(const int* in, int* out) {
    int e_in[N / 2], e_out[N / 2];
#pragma unroll
    for (int j = 0; j < N / 2; j++) e_in[j] = in[2 * j];
    InvDct<N / 2, STEP * 2>::run(e_in, e_out);
#pragma unroll
    for (int i = 0; i < N / 2; i++) {
      int o = 0;
#pragma unroll
      for (int j = 0; j < N / 2; j++) o += dct_coef((2 * j + 1) * STEP, i) * in[2 * j + 1];
      out[i] = e_out[i] + o;
      out[N - 1 - i] = e_out[i] - o;
    }
  }
};
template <int STEP>
struct InvDct<2, STEP> {
  HC_HD static void run(const int* in, int* out) {
    // T[0][*] = 64, T[16][0] = 64, T[16][1] = -64
    out[0] = 64 * in[0] + 64 * in[1];
    out[1] = 64 * in[0] - 64 * in[1];
  }
};

// 4-point inverse DST-VII (luma intra 4x4): out[i] = sum_j M[j][i] * in[j]
HC_HD void inv_dst4(const int* in, int* out) {
  out[0] = 29 * in[0] + 74 * in[1] + 84 * in[2] + 55 * in[3];
  out[1] = 55 * in[0] + 74 * in[1] - 29 * in[2] - 84 * in[3];
  out[2] = 74 * in[0] + 0 * in[1] - 74 * in[2] + 74 * in[3];
  out[3] = 84 * in[0] - 74 * in[1] + 55 * in[2] - 29 * in[3];
}

HC_HD int level_scale(int r) {
  // H.265 eq. (8-309): levelScale[] = {40,45,51,57,64,72}
  return r == 0 ? 40 : r == 1 ? 45 : r == 2 ? 51 : r == 3 ? 57 : r == 4 ? 64 : 72;
}

// Flat dequantisation of one level (transform.cc:487-503): 32-bit arithmetic with wrap, the
// m=16 factor folded into the shift.
HC_HD int dequant_flat(int level, int qp, int bd_shift_minus4) {
  const unsigned fact = (unsigned)(level_scale(qp % 6) << (qp / 6));
  const unsigned offset = 1u << (bd_shift_minus4 - 1);
  const int v = (int)((unsigned)level * fact + offset) >> bd_shift_minus4;
  return sat16(v);
}

// Scaling-list dequantisation (transform.cc:509-545): 64-bit product.
HC_HD int dequant_scaled(int level, int qp, int m, int bd_shift) {
  const int fact = (m * level_scale(qp % 6)) << (qp / 6);
  const long long offset = 1ll << (bd_shift - 1);
  long long v = ((long long)level * fact + offset) >> bd_shift;
  return (int)(v < -32768 ? -32768 : (v > 32767 ? 32767 : v));
}

}  // namespace hc
