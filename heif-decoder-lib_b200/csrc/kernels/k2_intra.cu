// k2_intra.cu — K2: intra prediction + residual add as a CTB wavefront over every picture of a
// batch at once.
//
// Replaces decode_intra_prediction (intrapred.cc:337-362: border gather intrapred.h:838-984,
// reference smoothing :192-266, planar :269-293, DC :296-330, angular :338-441) and the
// "+= residual, clip" tail of the reference's transform functions.
//
// Parallel structure (nothing like the reference's per-TU call chain):
//   * one warp owns one CTB row of one colour component of one picture (RowTask); luma, Cb and Cr
//     are independent dependency chains and run concurrently;
//   * a row may process CTB x once the row above has finished CTB x+1 (left / top-left / top /
//     top-right neighbours): the classic 2-CTB wavefront, tracked with one acquire/release
//     counter per task in global memory;
//   * tasks are ordered row-major across ALL pictures, so every picture / grid tile of the batch
//     advances its own wavefront simultaneously and a task only ever waits on a task with a
//     smaller index (scheduled no later than itself -> forward progress);
//   * inside a block the 32 lanes gather the 4nT+1 reference samples, substitute, smooth and
//     predict cooperatively; slice/tile/z-scan availability was resolved by the host
//     (hc_blk::avail_*), so the kernel carries no bitstream-structure logic.
//
// Bound: dependency latency (L2 round trips per block), not HBM; see DESIGN.md.
#include "launch.h"

namespace hc {

constexpr int K2_WARPS = 8;

__device__ __constant__ int8_t c_intra_angle[35] = {0,   0,   32,  26,  21,  17,  13,  9,   5,  2,  0,  -2,
                                                    -5,  -9,  -13, -17, -21, -26, -32, -26, -21, -17, -13, -9,
                                                    -5,  -2,  0,   2,   5,   9,   13,  17,  21,  26,  32};
__device__ __constant__ int16_t c_inv_angle[15] = {-4096, -1638, -910, -630, -482, -390, -315, -256,
                                                   -315,  -390,  -482, -630, -910, -1638, -4096};

// Per-warp scratch: reference samples p[-64..64] (index + REF_OFF) in two buffers.
constexpr int REF_OFF = 66;
constexpr int REF_LEN = 136;

template <typename Pixel>
__device__ void process_block(const hc_pic& pic, const hc_blk blk, Pixel* __restrict__ plane, int stride,
                              const int16_t* __restrict__ resid, int16_t* refA, int16_t* refB, int lane) {
  const int log2 = blk.log2, nT = 1 << log2, cidx = blk.cidx;
  const int bit_depth = cidx == 0 ? pic.bit_depth_y : pic.bit_depth_c;
  const int x0 = blk.x, y0 = blk.y;
  const int16_t* __restrict__ res = resid + blk.resid_off;

  if (blk.flags & HC_BLK_PCM) {
    for (int s = lane; s < nT * nT; s += 32) {
      int x = s & (nT - 1), y = s >> log2;
      plane[(size_t)(y0 + y) * stride + x0 + x] = (Pixel)(uint16_t)res[s];
    }
    __syncwarp();
    return;
  }

  // ---- 1. gather reference samples (L2-coherent loads) ------------------------------------------
  const unsigned availL = blk.avail_left, availT = blk.avail_top;
  const bool availTL = blk.flags & HC_BLK_AVAIL_TL;
  const int nref = 4 * nT + 1;
  for (int idx = lane; idx < nref; idx += 32) {
    const int i = idx - 2 * nT;
    int v = 0;
    if (i < 0) {
      const int y = -i - 1;
      if ((availL >> (y >> 2)) & 1) v = (int)ld_sample_cg(plane + (size_t)(y0 + y) * stride + x0 - 1);
    } else if (i == 0) {
      if (availTL) v = (int)ld_sample_cg(plane + (size_t)(y0 - 1) * stride + x0 - 1);
    } else {
      const int x = i - 1;
      if ((availT >> (x >> 2)) & 1) v = (int)ld_sample_cg(plane + (size_t)(y0 - 1) * stride + x0 + x);
    }
    refA[REF_OFF + i] = (int16_t)v;
  }
  __syncwarp();

  // ---- 2. substitution (intrapred.h:944-984) ---------------------------------------------------
  // Slots in scan order: left units bottom->top, top-left, top units left->right.
  const int nu = nT >> 1;  // 2*nT/4 units per side
  unsigned long long A = 0;
  {
    // bit s (s < nu)  = left unit (nu-1-s);  bit nu = TL;  bit nu+1+u = top unit u
    unsigned revL = __brev(availL) >> (32 - nu);
    A = (unsigned long long)revL | ((unsigned long long)(availTL ? 1 : 0) << nu) |
        ((unsigned long long)(availT & ((1u << nu) - 1u)) << (nu + 1));
    if (nu == 16) A = (unsigned long long)(__brev(availL) >> 16) | ((unsigned long long)(availTL ? 1 : 0) << 16) |
                      ((unsigned long long)availT << 17);
  }
  const bool all_avail = (A == ((1ull << (2 * nu + 1)) - 1ull));
  if (!all_avail) {
    int vals[5];
#pragma unroll
    for (int k = 0; k < 5; k++) {
      const int idx = lane + 32 * k;
      int v = 0;
      if (idx < nref) {
        const int i = idx - 2 * nT;
        if (A == 0) {
          v = 1 << (bit_depth - 1);
        } else {
          int s;
          if (i < 0) s = nu - 1 - ((-i - 1) >> 2);
          else if (i == 0) s = nu;
          else s = nu + 1 + ((i - 1) >> 2);
          int src = i;
          if (!((A >> s) & 1)) {
            const unsigned long long below = A & ((1ull << s) - 1ull);
            if (below) {
              const int ps = 63 - __clzll((long long)below);  // nearest available slot before s
              // last sample (highest index) of slot ps
              if (ps < nu) src = -(4 * (nu - 1 - ps)) - 1;
              else if (ps == nu) src = 0;
              else src = 4 * (ps - nu - 1) + 4;
            } else {
              const int fs = __ffsll((long long)A) - 1;  // first available slot
              // first sample (lowest index) of slot fs
              if (fs < nu) src = -(4 * (nu - 1 - fs) + 3) - 1;
              else if (fs == nu) src = 0;
              else src = 4 * (fs - nu - 1) + 1;
            }
          }
          v = refA[REF_OFF + src];
        }
      }
      vals[k] = v;
    }
    __syncwarp();
#pragma unroll
    for (int k = 0; k < 5; k++) {
      const int idx = lane + 32 * k;
      if (idx < nref) refA[REF_OFF + idx - 2 * nT] = (int16_t)vals[k];
    }
    __syncwarp();
  }

  // ---- 3. reference smoothing (intrapred.h:192-266) -------------------------------------------
  const int mode = blk.mode;
  int16_t* p = refA + REF_OFF;
  if (!(pic.flags & HC_PIC_NO_INTRA_SMOOTH) && (cidx == 0 || pic.chroma_format == 3) && mode != 1 && nT != 4) {
    const int d1 = iabs(mode - 26), d2 = iabs(mode - 10);
    const int minDist = d1 < d2 ? d1 : d2;
    const bool filter = nT == 8 ? minDist > 7 : nT == 16 ? minDist > 1 : minDist > 0;
    if (filter) {
      bool bi = false;
      if ((pic.flags & HC_PIC_STRONG_INTRA) && cidx == 0 && nT == 32) {
        const int thr = 1 << (pic.bit_depth_y - 5);
        bi = iabs(p[0] + p[64] - 2 * p[32]) < thr && iabs(p[0] + p[-64] - 2 * p[-32]) < thr;
      }
      int16_t* q = refB + REF_OFF;
      for (int idx = lane; idx < nref; idx += 32) {
        const int i = idx - 2 * nT;
        int v;
        if (i == -2 * nT || i == 2 * nT) v = p[i];
        else if (bi) {
          if (i == 0) v = p[0];
          else if (i < 0) v = p[0] + (((-i) * (p[-64] - p[0]) + 32) >> 6);
          else v = p[0] + ((i * (p[64] - p[0]) + 32) >> 6);
        } else {
          v = (p[i + 1] + 2 * p[i] + p[i - 1] + 2) >> 2;
        }
        q[i] = (int16_t)v;
      }
      __syncwarp();
      p = q;
    }
  }

  // ---- 4. prediction + residual + store ---------------------------------------------------------
  const bool has_res = blk.flags & HC_BLK_HAS_RESID;
  const int maxv = (1 << bit_depth) - 1;
  int dc = 0;
  if (mode == 1) {
    int sum = 0;
    for (int i = lane; i < nT; i += 32) sum += p[i + 1] + p[-i - 1];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    dc = (sum + nT) >> (log2 + 1);
  }
  const int angle = mode >= 2 ? c_intra_angle[mode] : 0;
  const int inv = (mode >= 11 && mode <= 25) ? c_inv_angle[mode - 11] : 0;
  const bool edge_ok = cidx == 0 && nT < 32;
  const bool no_edge_flt = blk.flags & HC_BLK_NO_EDGE_FLT;

  for (int s = lane; s < nT * nT; s += 32) {
    const int x = s & (nT - 1), y = s >> log2;
    int v;
    if (mode == 0) {
      v = ((nT - 1 - x) * p[-1 - y] + (x + 1) * p[1 + nT] + (nT - 1 - y) * p[1 + x] + (y + 1) * p[-1 - nT] + nT) >>
          (log2 + 1);
    } else if (mode == 1) {
      v = dc;
      if (edge_ok) {
        if (x == 0 && y == 0) v = (p[-1] + 2 * dc + p[1] + 2) >> 2;
        else if (y == 0) v = (p[x + 1] + 3 * dc + 2) >> 2;
        else if (x == 0) v = (p[-y - 1] + 3 * dc + 2) >> 2;
      }
    } else if (mode >= 18) {
      const int iIdx = ((y + 1) * angle) >> 5, iFact = ((y + 1) * angle) & 31;
      const int k0 = x + iIdx + 1;
      // ref[k] = p[k] for k >= 0, projected left column for k < 0
      const int a = k0 >= 0 ? p[k0] : p[-((k0 * inv + 128) >> 8)];
      if (iFact) {
        const int k1 = k0 + 1;
        const int b = k1 >= 0 ? p[k1] : p[-((k1 * inv + 128) >> 8)];
        v = ((32 - iFact) * a + iFact * b + 16) >> 5;
      } else {
        v = a;
      }
      if (mode == 26 && edge_ok && !no_edge_flt && x == 0) v = clip3i(0, maxv, p[1] + ((p[-1 - y] - p[0]) >> 1));
    } else {
      const int iIdx = ((x + 1) * angle) >> 5, iFact = ((x + 1) * angle) & 31;
      const int k0 = y + iIdx + 1;
      // ref[k] = p[-k] for k >= 0, projected top row for k < 0
      const int a = k0 >= 0 ? p[-k0] : p[(k0 * inv + 128) >> 8];
      if (iFact) {
        const int k1 = k0 + 1;
        const int b = k1 >= 0 ? p[-k1] : p[(k1 * inv + 128) >> 8];
        v = ((32 - iFact) * a + iFact * b + 16) >> 5;
      } else {
        v = a;
      }
      if (mode == 10 && edge_ok && !no_edge_flt && y == 0) v = clip3i(0, maxv, p[-1] + ((p[1 + x] - p[0]) >> 1));
    }
    if (has_res) v = clip3i(0, maxv, v + res[s]);
    plane[(size_t)(y0 + y) * stride + x0 + x] = (Pixel)v;
  }
  // make the block visible to the lanes that gather the next block's references
  __syncwarp();
}

template <typename Pixel>
__device__ void run_row(const BatchView& bv, const hc_pic& pic, const RowTask task, int* progress, int my_index,
                        int16_t* refA, int16_t* refB, int lane) {
  const int comp = task.comp;
  Pixel* plane = reinterpret_cast<Pixel*>(bv.planes + pic.rec_off[comp]);
  const int stride = (int)pic.rec_stride[comp];
  const int16_t* resid = bv.resid + pic.resid_base;
  const hc_blk* blks = bv.blks + pic.blk_base;
  const hc_ctu* ctus = bv.ctus + pic.ctu_base + (size_t)task.row * pic.ctbs_w;
  const int W = pic.ctbs_w;

  for (int cx = 0; cx < W; cx++) {
    if (task.dep >= 0) {
      const int need = cx + 2 < W ? cx + 2 : W;
      if (lane == 0) {
        while (ld_acquire_s32(progress + task.dep) < need) __nanosleep(64);
      }
      __syncwarp();
    }
    const uint32_t first = ctus[cx].blk_first[comp];
    const int n = ctus[cx].blk_count[comp];
    for (int k = 0; k < n; k++) process_block<Pixel>(pic, blks[first + k], plane, stride, resid, refA, refB, lane);
    __threadfence();
    __syncwarp();
    if (lane == 0) st_release_s32(progress + my_index, cx + 1);
  }
}

__global__ void __launch_bounds__(K2_WARPS * 32)
k2_intra_kernel(BatchView bv, const RowTask* __restrict__ tasks, int ntasks, int* progress) {
  __shared__ int16_t ref[K2_WARPS][2][REF_LEN];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int index = blockIdx.x * K2_WARPS + warp;
  if (index >= ntasks) return;
  const RowTask task = tasks[index];
  const hc_pic& pic = bv.pics[task.pic];
  const int bd = task.comp == 0 ? pic.bit_depth_y : pic.bit_depth_c;
  // planes of a picture share one sample type: 8-bit pictures use bytes, everything else uint16
  if (pic.bit_depth_y == 8 && pic.bit_depth_c == 8)
    run_row<uint8_t>(bv, pic, task, progress, index, ref[warp][0], ref[warp][1], lane);
  else
    run_row<uint16_t>(bv, pic, task, progress, index, ref[warp][0], ref[warp][1], lane);
  (void)bd;
}

void launch_k2(const BatchView& bv, const RowTask* tasks, int ntasks, int* progress, cudaStream_t stream) {
  if (ntasks <= 0) return;
  const int grid = (ntasks + K2_WARPS - 1) / K2_WARPS;
  k2_intra_kernel<<<grid, K2_WARPS * 32, 0, stream>>>(bv, tasks, ntasks, progress);
}

}  // namespace hc
