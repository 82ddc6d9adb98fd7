// k1_transform.cu — K1: dequantisation + inverse DST-4 / DCT-4/8/16/32 (+ transform-skip, bypass,
// implicit RDPCM, rotation) of every coded transform block of a batch.
//
// Replaces scale_coefficients_internal (transform.cc:386-689) and the SIMD transforms behind
// acceleration_functions::transform_add / transform_4x4_dst_add / transform_dc_add
// (x86_new/x86_idct.cc), minus the "+= prediction" which belongs to K2.
//
// Mapping: transform blocks are bucketed by size on the host. A block of size N is processed by N
// consecutive lanes (32/N blocks per warp): sparse (pos,level) records are loaded coalesced,
// dequantised and scattered into a shared-memory tile; pass 1 gives every lane one column in
// registers, pass 2 one row; the row is written as 16-byte vectors into the block's contiguous
// int16 residual tile. No tensor cores: the stages are small exact-integer butterflies with an
// int16 clip in between.
//
// Algorithmic bytes per block: 4*ncoeff (records) + 16 (hc_tb) read, 2*N*N written.
#include <algorithm>
#include "launch.h"
#include "k1_core.cuh"

namespace hc {

constexpr int K1_WARPS = 4;

template <int LOG2>
__device__ __forceinline__ void k1_block(const BatchView& bv, const uint32_t* __restrict__ tb_index, long long slot,
                                         int16_t* tile, int grp, int t, unsigned grp_mask) {
  constexpr int N = 1 << LOG2;
  constexpr int S = N + 2;  // padded row stride (int16): conflict-free column and row access
  const hc_tb tb = bv.tbs[tb_index[slot]];
  const hc_pic& pic = bv.pics[tb.pic];
  const int cidx = tb.type & HC_TB_CIDX_MASK;
  const int bit_depth = cidx == 0 ? pic.bit_depth_y : pic.bit_depth_c;
  const hc_coeff* __restrict__ co = bv.coeffs + pic.coeff_base + tb.coeff_off;
  int16_t* __restrict__ out = bv.resid + pic.resid_base + tb.resid_off;

  // ---- zero the tile, then scatter the dequantised coefficients -------------------------------
#pragma unroll
  for (int j = 0; j < N; j++) tile[t * S + j] = 0;
  __syncwarp(grp_mask);

  const bool bypass = tb.type & HC_TB_BYPASS;
  const bool rotate = tb.type & HC_TB_ROTATE;
  const int bd_shift = bit_depth + LOG2 - 5;
  const uint8_t* sc = nullptr;
  if (!bypass && (pic.flags & HC_PIC_SCALING_LIST)) {
    const uint8_t* base = bv.scaling + pic.scaling_base;
    sc = LOG2 == 2 ? base + tb.matrix_id * 16
       : LOG2 == 3 ? base + 96 + tb.matrix_id * 64
       : LOG2 == 4 ? base + 96 + 384 + tb.matrix_id * 256
                   : base + 96 + 384 + 1536 + (tb.matrix_id ? 1024 : 0);
  }
  for (int i = t; i < tb.ncoeff; i += N) {
    const hc_coeff c = co[i];
    int v;
    if (bypass) v = c.level;
    else if (sc) v = dequant_scaled(c.level, tb.qp, sc[c.pos], bd_shift);
    else v = dequant_flat(c.level, tb.qp, bd_shift - 4);
    int x = c.pos & (N - 1), y = c.pos >> LOG2;
    if (rotate) { x = N - 1 - x; y = N - 1 - y; }
    tile[y * S + x] = (int16_t)v;
  }
  __syncwarp(grp_mask);

  int row[N];
  if (!(tb.type & (HC_TB_BYPASS | HC_TB_TSKIP))) {
    // ---- pass 1: lane t owns column t ---------------------------------------------------------
    int in[N], o[N];
#pragma unroll
    for (int j = 0; j < N; j++) in[j] = tile[j * S + t];
    if (LOG2 == 2 && (tb.type & HC_TB_DST)) inv_dst4(in, o);
    else InvDct<N, 32 / N>::run(in, o);
    __syncwarp(grp_mask);
#pragma unroll
    for (int i = 0; i < N; i++) tile[i * S + t] = (int16_t)sat16((o[i] + 64) >> 7);
    __syncwarp(grp_mask);
    // ---- pass 2: lane t owns row t ------------------------------------------------------------
#pragma unroll
    for (int j = 0; j < N; j++) in[j] = tile[t * S + j];
    if (LOG2 == 2 && (tb.type & HC_TB_DST)) inv_dst4(in, o);
    else InvDct<N, 32 / N>::run(in, o);
    const int shift2 = 20 - bit_depth, rnd2 = 1 << (shift2 - 1);
#pragma unroll
    for (int i = 0; i < N; i++) row[i] = sat16((o[i] + rnd2) >> shift2);
  } else {
    // ---- transform skip / transquant bypass, optional implicit RDPCM (fallback-dct.cc:84-260) ---
    const bool tskip = tb.type & HC_TB_TSKIP;
    int bd2 = 20 - bit_depth;
    if (bd2 < 0) bd2 = 0;
    const int ts_shift = 5 + LOG2, rnd = bd2 > 0 ? 1 << (bd2 - 1) : 0;
    auto conv = [&](int c) -> int { return tskip ? (((c << ts_shift) + rnd) >> bd2) : c; };
    if (tb.type & HC_TB_RDPCM_V) {
      // lane t accumulates down column t, result goes back through the tile
      int sum = 0;
      int col[N];
#pragma unroll
      for (int y = 0; y < N; y++) { sum += conv(tile[y * S + t]); col[y] = sat16(sum); }
      __syncwarp(grp_mask);
#pragma unroll
      for (int y = 0; y < N; y++) tile[y * S + t] = (int16_t)col[y];
      __syncwarp(grp_mask);
#pragma unroll
      for (int x = 0; x < N; x++) row[x] = tile[t * S + x];
    } else if (tb.type & HC_TB_RDPCM_H) {
      int sum = 0;
#pragma unroll
      for (int x = 0; x < N; x++) { sum += conv(tile[t * S + x]); row[x] = sat16(sum); }
    } else {
#pragma unroll
      for (int x = 0; x < N; x++) row[x] = sat16(conv(tile[t * S + x]));
    }
  }

  // ---- write row t of the residual tile: N int16, contiguous -----------------------------------
  int16_t* dst = out + t * N;
  if (N == 4) {
    uint2 v;
    v.x = (uint32_t)(uint16_t)row[0] | ((uint32_t)(uint16_t)row[1] << 16);
    v.y = (uint32_t)(uint16_t)row[2] | ((uint32_t)(uint16_t)row[3] << 16);
    *reinterpret_cast<uint2*>(dst) = v;
  } else {
#pragma unroll
    for (int k = 0; k < N / 8; k++) {
      uint4 v;
      v.x = (uint32_t)(uint16_t)row[8 * k + 0] | ((uint32_t)(uint16_t)row[8 * k + 1] << 16);
      v.y = (uint32_t)(uint16_t)row[8 * k + 2] | ((uint32_t)(uint16_t)row[8 * k + 3] << 16);
      v.z = (uint32_t)(uint16_t)row[8 * k + 4] | ((uint32_t)(uint16_t)row[8 * k + 5] << 16);
      v.w = (uint32_t)(uint16_t)row[8 * k + 6] | ((uint32_t)(uint16_t)row[8 * k + 7] << 16);
      reinterpret_cast<uint4*>(dst)[k] = v;
    }
  }
}

// The launch list holds `count` entries, or *count_ptr when that is non-null: lists built on the device by K0 are
// consumed without a host round trip, by a grid sized for the list capacity whose CTAs stride over the real count.
// Resident CTAs the register allocation must allow: the kernels wait on their sparse record loads (long-scoreboard
// stalls), so occupancy is worth a few spills — 32x32: 156 -> 96 registers (5 CTAs), 16x16: 56 -> 40 (12 CTAs).
template <int LOG2>
__global__ void __launch_bounds__(K1_WARPS * 32, LOG2 == 5 ? 8 : 16)
k1_transform_kernel(BatchView bv, const uint32_t* __restrict__ tb_index, int count, const unsigned* __restrict__ count_ptr) {
  constexpr int N = 1 << LOG2;
  constexpr int PER_WARP = 32 / N;
  constexpr int S = N + 2;
  __shared__ int16_t tiles[K1_WARPS][PER_WARP][N * S];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int grp = lane / N, t = lane % N;
  const unsigned grp_mask = (N == 32) ? 0xffffffffu : (((1u << N) - 1u) << (grp * N));
  const long long cnt = count_ptr ? (long long)*count_ptr : (long long)count;
  const long long step = (long long)gridDim.x * K1_WARPS * PER_WARP;
  for (long long slot = ((long long)blockIdx.x * K1_WARPS + warp) * PER_WARP + grp; slot < cnt; slot += step) {  // whole groups
    k1_block<LOG2>(bv, tb_index, slot, tiles[warp][grp], grp, t, grp_mask);
    __syncwarp(grp_mask);   // the tile is reused by the next block of this group
  }
}

// 4x4 blocks (three quarters of all transform blocks of typical content): ONE LANE per block. The generic kernel
// above spends 4 lanes, a shared-memory tile and 4 warp barriers on 16 samples and is instruction-bound (ncu: 85 %
// issue utilisation at 17 % DRAM throughput); here the block lives in 16 registers of its lane, only the scatter of the
// sparse (pos, level) records goes through shared memory (word w of thread t at tiles[w][t]: conflict-free for the
// zeroing and the read-back), and a warp writes 32 consecutive 32-byte residual tiles.
constexpr int K1_44_THREADS = 128;
__global__ void __launch_bounds__(K1_44_THREADS)
k1_transform4x4_kernel(BatchView bv, const uint32_t* __restrict__ tb_index, int count, const unsigned* __restrict__ count_ptr) {
  __shared__ uint32_t tiles[8][K1_44_THREADS];
  const int t = threadIdx.x;
  const long long cnt = count_ptr ? (long long)*count_ptr : (long long)count;
  const long long step = (long long)gridDim.x * K1_44_THREADS;
  for (long long slot = (long long)blockIdx.x * K1_44_THREADS + t; slot < cnt; slot += step) {
    const hc_tb tb = bv.tbs[tb_index[slot]];
    const hc_pic& pic = bv.pics[tb.pic];
    const int cidx = tb.type & HC_TB_CIDX_MASK;
    const int bit_depth = cidx == 0 ? pic.bit_depth_y : pic.bit_depth_c;
    const hc_coeff* __restrict__ co = bv.coeffs + pic.coeff_base + tb.coeff_off;
    int16_t* __restrict__ out = bv.resid + pic.resid_base + tb.resid_off;
    const bool bypass = tb.type & HC_TB_BYPASS;
    const bool rotate = tb.type & HC_TB_ROTATE;
    const int bd_shift = bit_depth + 2 - 5;
    const uint8_t* sc = nullptr;
    if (!bypass && (pic.flags & HC_PIC_SCALING_LIST)) sc = bv.scaling + pic.scaling_base + tb.matrix_id * 16;

#pragma unroll
    for (int w = 0; w < 8; w++) tiles[w][t] = 0;
    for (int i = 0; i < tb.ncoeff; i++) {
      const hc_coeff c = co[i];
      int v;
      if (bypass) v = c.level;
      else if (sc) v = dequant_scaled(c.level, tb.qp, sc[c.pos], bd_shift);
      else v = dequant_flat(c.level, tb.qp, bd_shift - 4);
      int p = c.pos & 15;
      if (rotate) p = 15 - p;                       // (x, y) -> (3 - x, 3 - y)
      reinterpret_cast<int16_t*>(&tiles[p >> 1][t])[p & 1] = (int16_t)v;
    }
    int m[16];
#pragma unroll
    for (int w = 0; w < 8; w++) {
      const uint32_t u = tiles[w][t];
      m[2 * w] = (int)(int16_t)(u & 0xffff);
      m[2 * w + 1] = (int)(int16_t)(u >> 16);
    }

    int res[16];
    if (!(tb.type & (HC_TB_BYPASS | HC_TB_TSKIP))) {
      const bool dst = tb.type & HC_TB_DST;
      int tmp[16];
#pragma unroll
      for (int x = 0; x < 4; x++) {                 // pass 1: columns
        int in[4] = {m[x], m[4 + x], m[8 + x], m[12 + x]}, o[4];
        if (dst) inv_dst4(in, o);
        else InvDct<4, 8>::run(in, o);
#pragma unroll
        for (int i = 0; i < 4; i++) tmp[i * 4 + x] = sat16((o[i] + 64) >> 7);
      }
      const int shift2 = 20 - bit_depth, rnd2 = 1 << (shift2 - 1);
#pragma unroll
      for (int y = 0; y < 4; y++) {                 // pass 2: rows
        int in[4] = {tmp[4 * y], tmp[4 * y + 1], tmp[4 * y + 2], tmp[4 * y + 3]}, o[4];
        if (dst) inv_dst4(in, o);
        else InvDct<4, 8>::run(in, o);
#pragma unroll
        for (int i = 0; i < 4; i++) res[4 * y + i] = sat16((o[i] + rnd2) >> shift2);
      }
    } else {
      // transform skip / transquant bypass, optional implicit RDPCM (fallback-dct.cc:84-260)
      const bool tskip = tb.type & HC_TB_TSKIP;
      int bd2 = 20 - bit_depth;
      if (bd2 < 0) bd2 = 0;
      const int ts_shift = 5 + 2, rnd = bd2 > 0 ? 1 << (bd2 - 1) : 0;
      auto conv = [&](int c) -> int { return tskip ? (((c << ts_shift) + rnd) >> bd2) : c; };
      if (tb.type & HC_TB_RDPCM_V) {
#pragma unroll
        for (int x = 0; x < 4; x++) {
          int sum = 0;
#pragma unroll
          for (int y = 0; y < 4; y++) { sum += conv(m[4 * y + x]); res[4 * y + x] = sat16(sum); }
        }
      } else if (tb.type & HC_TB_RDPCM_H) {
#pragma unroll
        for (int y = 0; y < 4; y++) {
          int sum = 0;
#pragma unroll
          for (int x = 0; x < 4; x++) { sum += conv(m[4 * y + x]); res[4 * y + x] = sat16(sum); }
        }
      } else {
#pragma unroll
        for (int k = 0; k < 16; k++) res[k] = sat16(conv(m[k]));
      }
    }
    uint4 v0, v1;
    v0.x = (uint32_t)(uint16_t)res[0] | ((uint32_t)(uint16_t)res[1] << 16);   v0.y = (uint32_t)(uint16_t)res[2] | ((uint32_t)(uint16_t)res[3] << 16);
    v0.z = (uint32_t)(uint16_t)res[4] | ((uint32_t)(uint16_t)res[5] << 16);   v0.w = (uint32_t)(uint16_t)res[6] | ((uint32_t)(uint16_t)res[7] << 16);
    v1.x = (uint32_t)(uint16_t)res[8] | ((uint32_t)(uint16_t)res[9] << 16);   v1.y = (uint32_t)(uint16_t)res[10] | ((uint32_t)(uint16_t)res[11] << 16);
    v1.z = (uint32_t)(uint16_t)res[12] | ((uint32_t)(uint16_t)res[13] << 16); v1.w = (uint32_t)(uint16_t)res[14] | ((uint32_t)(uint16_t)res[15] << 16);
    reinterpret_cast<uint4*>(out)[0] = v0;
    reinterpret_cast<uint4*>(out)[1] = v1;
  }
}

// Host-side launcher. counts[l] blocks of log2 size l+2, indices in tb_index[l] (device pointers).
void launch_k1(const BatchView& bv, const uint32_t* const tb_index[4], const int counts[4], cudaStream_t stream) {
  if (counts[0] > 0)
    k1_transform4x4_kernel<<<(counts[0] + K1_44_THREADS - 1) / K1_44_THREADS, K1_44_THREADS, 0, stream>>>(bv, tb_index[0], counts[0], nullptr);
  if (counts[1] > 0) {
    int per_cta = K1_WARPS * 4;
    k1_transform_kernel<3><<<(counts[1] + per_cta - 1) / per_cta, K1_WARPS * 32, 0, stream>>>(bv, tb_index[1], counts[1], nullptr);
  }
  if (counts[2] > 0) {
    int per_cta = K1_WARPS * 2;
    k1_transform_kernel<4><<<(counts[2] + per_cta - 1) / per_cta, K1_WARPS * 32, 0, stream>>>(bv, tb_index[2], counts[2], nullptr);
  }
  if (counts[3] > 0) {
    int per_cta = K1_WARPS;
    k1_transform_kernel<5><<<(counts[3] + per_cta - 1) / per_cta, K1_WARPS * 32, 0, stream>>>(bv, tb_index[3], counts[3], nullptr);
  }
}

// Lists whose lengths only the device knows (d_counts[4], written by K0): a fixed grid of a few CTAs per SM strides
// over each list.
void launch_k1_indirect(const BatchView& bv, const uint32_t* const tb_index[4], const long long capacity[4], const unsigned* d_counts,
                        int sm_count, cudaStream_t stream) {
  const int per_cta[4] = {K1_WARPS * 8, K1_WARPS * 4, K1_WARPS * 2, K1_WARPS};
  auto grid = [&](int l) { return (int)std::max<long long>(1, std::min<long long>((capacity[l] + per_cta[l] - 1) / per_cta[l], (long long)sm_count * 16)); };
  if (capacity[0] > 0) {
    const int g = (int)std::max<long long>(1, std::min<long long>((capacity[0] + K1_44_THREADS - 1) / K1_44_THREADS, (long long)sm_count * 16));
    k1_transform4x4_kernel<<<g, K1_44_THREADS, 0, stream>>>(bv, tb_index[0], 0, d_counts + 0);
  }
  if (capacity[1] > 0) k1_transform_kernel<3><<<grid(1), K1_WARPS * 32, 0, stream>>>(bv, tb_index[1], 0, d_counts + 1);
  if (capacity[2] > 0) k1_transform_kernel<4><<<grid(2), K1_WARPS * 32, 0, stream>>>(bv, tb_index[2], 0, d_counts + 2);
  if (capacity[3] > 0) k1_transform_kernel<5><<<grid(3), K1_WARPS * 32, 0, stream>>>(bv, tb_index[3], 0, d_counts + 3);
}

}  // namespace hc
