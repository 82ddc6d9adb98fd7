// k3_deblock.cu — K3: HEVC deblocking filter for intra pictures as two fully parallel passes
// (all vertical edges of every picture, then all horizontal edges), luma and chroma in one launch.
//
// Replaces apply_deblocking_filter (deblock.cc:1921-1959): edge_filtering_luma_internal
// (:708-792) + loop_filter_luma (fallback-postfilter.h:31-135 / x86_new/x86_dbk.cc:295) and
// edge_filtering_chroma_internal (:1607-1772) + loop_filter_chroma (:138-179 / x86_dbk.cc:551).
// Edge flags (bS == 2 on every marked edge of an intra picture, deblock.cc:275-277), QP_Y per 8x8
// and the slice's beta/tc offsets come from the host records.
//
// Mapping: one thread filters one 4-line unit of one edge (the unit on which the standard takes
// its filter decision). Threads are ordered so that a warp touches consecutive addresses of the
// same rows. Filtering along an edge only modifies 3 samples on either side and edges lie 8
// samples apart, so all units of one direction are independent and the pass runs in place.
//
// Algorithmic bytes: read + write every sample once per pass (2 * s * c B/px per direction),
// plus 1/16 B/px edge map and 1/64 B/px QP map.
#include "launch.h"

namespace hc {

__device__ __constant__ uint8_t c_beta_tab[52] = {0,  0,  0,  0,  0,  0,  0,  0,  0,  0,  0,  0,  0,  0,  0,  0,  6,  7,
                                                  8,  9,  10, 11, 12, 13, 14, 15, 16, 17, 18, 20, 22, 24, 26, 28, 30, 32,
                                                  34, 36, 38, 40, 42, 44, 46, 48, 50, 52, 54, 56, 58, 60, 62, 64};
__device__ __constant__ uint8_t c_tc_tab[54] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0,  0,  0,  0,  0,
                                                1, 1, 1, 1, 1, 1, 1, 1, 1, 2, 2, 2, 2, 3,  3,  3,  3,  4,
                                                4, 4, 5, 5, 6, 6, 7, 8, 9, 10, 11, 13, 14, 16, 18, 20, 22, 24};

HC_D int qpc_from_qpi_420(int qPi) {
  // H.265 Table 8-10
  if (qPi < 30) return qPi;
  if (qPi >= 44) return qPi - 6;
  // 30..43 -> 29,30,31,32,33,33,34,34,35,35,36,36,37,37
  const int d = qPi - 30;
  return d < 4 ? 29 + d : 33 + ((d - 4) >> 1);
}

// Luma: 4 lines x 8 samples (p3 p2 p1 p0 | q0 q1 q2 q3) held in registers.
template <typename Pixel>
__device__ void deblock_luma_unit(Pixel* __restrict__ pix, ptrdiff_t xs, ptrdiff_t ys, int beta, int tc, bool no_p,
                                  bool no_q, int bit_depth) {
  int s[4][8];
#pragma unroll
  for (int k = 0; k < 4; k++)
#pragma unroll
    for (int i = 0; i < 8; i++) s[k][i] = pix[(ptrdiff_t)(i - 4) * xs + (ptrdiff_t)k * ys];

  const int dp0 = iabs(s[0][1] - 2 * s[0][2] + s[0][3]), dq0 = iabs(s[0][6] - 2 * s[0][5] + s[0][4]);
  const int dp3 = iabs(s[3][1] - 2 * s[3][2] + s[3][3]), dq3 = iabs(s[3][6] - 2 * s[3][5] + s[3][4]);
  const int d0 = dp0 + dq0, d3 = dp3 + dq3;
  if (d0 + d3 >= beta) return;

  const int beta_3 = beta >> 3, beta_2 = beta >> 2, tc25 = (tc * 5 + 1) >> 1;
  const bool strong = iabs(s[0][0] - s[0][3]) + iabs(s[0][7] - s[0][4]) < beta_3 && iabs(s[0][3] - s[0][4]) < tc25 &&
                      iabs(s[3][0] - s[3][3]) + iabs(s[3][7] - s[3][4]) < beta_3 && iabs(s[3][3] - s[3][4]) < tc25 &&
                      (d0 << 1) < beta_2 && (d3 << 1) < beta_2;
  const int maxv = (1 << bit_depth) - 1;
  if (strong) {
    const int tc2 = tc << 1;
#pragma unroll
    for (int k = 0; k < 4; k++) {
      const int p3 = s[k][0], p2 = s[k][1], p1 = s[k][2], p0 = s[k][3];
      const int q0 = s[k][4], q1 = s[k][5], q2 = s[k][6], q3 = s[k][7];
      Pixel* l = pix + (ptrdiff_t)k * ys;
      if (!no_p) {
        l[-1 * xs] = (Pixel)(p0 + clip3i(-tc2, tc2, ((p2 + 2 * p1 + 2 * p0 + 2 * q0 + q1 + 4) >> 3) - p0));
        l[-2 * xs] = (Pixel)(p1 + clip3i(-tc2, tc2, ((p2 + p1 + p0 + q0 + 2) >> 2) - p1));
        l[-3 * xs] = (Pixel)(p2 + clip3i(-tc2, tc2, ((2 * p3 + 3 * p2 + p1 + p0 + q0 + 4) >> 3) - p2));
      }
      if (!no_q) {
        l[0] = (Pixel)(q0 + clip3i(-tc2, tc2, ((p1 + 2 * p0 + 2 * q0 + 2 * q1 + q2 + 4) >> 3) - q0));
        l[1 * xs] = (Pixel)(q1 + clip3i(-tc2, tc2, ((p0 + q0 + q1 + q2 + 2) >> 2) - q1));
        l[2 * xs] = (Pixel)(q2 + clip3i(-tc2, tc2, ((2 * q3 + 3 * q2 + q1 + q0 + p0 + 4) >> 3) - q2));
      }
    }
  } else {
    const int side_thr = (beta + (beta >> 1)) >> 3;
    const bool two_p = dp0 + dp3 < side_thr, two_q = dq0 + dq3 < side_thr;
    const int tc_2 = tc >> 1;
#pragma unroll
    for (int k = 0; k < 4; k++) {
      const int p2 = s[k][1], p1 = s[k][2], p0 = s[k][3];
      const int q0 = s[k][4], q1 = s[k][5], q2 = s[k][6];
      int delta = (9 * (q0 - p0) - 3 * (q1 - p1) + 8) >> 4;
      if (iabs(delta) >= 10 * tc) continue;
      delta = clip3i(-tc, tc, delta);
      Pixel* l = pix + (ptrdiff_t)k * ys;
      if (!no_p) {
        l[-1 * xs] = (Pixel)clip3i(0, maxv, p0 + delta);
        if (two_p) l[-2 * xs] = (Pixel)clip3i(0, maxv, p1 + clip3i(-tc_2, tc_2, (((p2 + p0 + 1) >> 1) - p1 + delta) >> 1));
      }
      if (!no_q) {
        l[0] = (Pixel)clip3i(0, maxv, q0 - delta);
        if (two_q) l[1 * xs] = (Pixel)clip3i(0, maxv, q1 + clip3i(-tc_2, tc_2, (((q2 + q0 + 1) >> 1) - q1 - delta) >> 1));
      }
    }
  }
}

// plane: 0 luma, 1 Cb, 2 Cr (blockIdx.z); picture: blockIdx.y
template <typename Pixel>
__device__ void deblock_picture(const BatchView& bv, const hc_pic& pic, int plane_idx, bool vertical, long long tid) {
  const int W = pic.width, H = pic.height;
  const int w4 = W >> 2, w8 = W >> 3;
  const uint8_t* __restrict__ edge = bv.edge_map + pic.edge_base;
  const int8_t* __restrict__ qp = bv.qp_map + pic.qp_base;
  const hc_ctu* __restrict__ ctus = bv.ctus + pic.ctu_base;
  Pixel* plane = reinterpret_cast<Pixel*>(bv.planes + pic.rec_off[plane_idx]);
  const int stride = (int)pic.rec_stride[plane_idx];
  const int mask = vertical ? HC_EDGE_V : HC_EDGE_H;

  if (plane_idx == 0) {
    int x, y;  // luma position of the unit's first q0 sample
    if (vertical) {
      const int nex = W >> 3;
      const long long total = (long long)nex * (H >> 2);
      if (tid >= total) return;
      x = (int)(tid % nex) << 3;
      y = (int)(tid / nex) << 2;
      if (x == 0) return;
    } else {
      const int nux = W >> 2;
      const long long total = (long long)nux * (H >> 3);
      if (tid >= total) return;
      x = (int)(tid % nux) << 2;
      y = (int)(tid / nux) << 3;
      if (y == 0) return;
    }
    const int e = edge[(x >> 2) + (size_t)(y >> 2) * w4];
    if (!(e & mask)) return;
    // QP / offsets are taken at the first unit of the 8-sample segment (deblock.cc:731-752)
    const int sx = vertical ? x : (x & ~7), sy = vertical ? (y & ~7) : y;
    const int QP_Q = qp[(sx >> 3) + (size_t)(sy >> 3) * w8];
    const int QP_P = vertical ? qp[((sx - 1) >> 3) + (size_t)(sy >> 3) * w8] : qp[(sx >> 3) + (size_t)((sy - 1) >> 3) * w8];
    const int qPL = (QP_Q + QP_P + 1) >> 1;
    const hc_ctu& ctu = ctus[(sx >> pic.log2_ctb) + (sy >> pic.log2_ctb) * pic.ctbs_w];
    const int bd = pic.bit_depth_y;
    const int beta = c_beta_tab[clip3i(0, 51, qPL + ctu.beta_offset)] * (1 << (bd - 8));
    const int tc = c_tc_tab[clip3i(0, 53, qPL + 2 + ctu.tc_offset)] * (1 << (bd - 8));
    // Streams with pcm(+loop filter disabled) / transquant bypass: mirror of the reference's
    // special path as its default build behaves (deblock.cc:755-790, see oracle/hevc_recon_oracle.c)
    bool no_p = false, no_q = false;
    if (pic.flags & HC_PIC_PCMF) {
      bool normal[2][2];
#pragma unroll
      for (int u = 0; u < 2; u++) {
        const int qx = vertical ? sx : sx + 4 * u, qy = vertical ? sy + 4 * u : sy;
        const int px = vertical ? qx - 1 : qx, py = vertical ? qy : qy - 1;
        normal[u][0] = !(edge[(px >> 2) + (size_t)(py >> 2) * w4] & (HC_EDGE_PCM | HC_EDGE_BYPASS));
        normal[u][1] = !(edge[(qx >> 2) + (size_t)(qy >> 2) * w4] & (HC_EDGE_PCM | HC_EDGE_BYPASS));
      }
      const int j = vertical ? ((y >> 2) & 1) : ((x >> 2) & 1);
      if (normal[0][0] && normal[0][1] && normal[1][0] && normal[1][1]) no_p = no_q = bd > 8;
      else { no_p = normal[j][0]; no_q = normal[j][1]; }
    }
    Pixel* pix = plane + x + (size_t)y * stride;
    if (vertical) deblock_luma_unit<Pixel>(pix, 1, stride, beta, tc, no_p, no_q, bd);
    else deblock_luma_unit<Pixel>(pix, stride, 1, beta, tc, no_p, no_q, bd);
    return;
  }

  // ---- chroma: edges on the 8-sample chroma grid, bS == 2 only, 1 sample each side ----------------
  if (pic.chroma_format == 0) return;
  const int SubW = (pic.chroma_format == 1 || pic.chroma_format == 2) ? 2 : 1;
  const int SubH = pic.chroma_format == 1 ? 2 : 1;
  const int CW = W / SubW, CH = H / SubH;
  int xc, yc;
  if (vertical) {
    const int nex = (CW + 7) >> 3;
    const long long total = (long long)nex * (CH >> 2);
    if (tid >= total) return;
    xc = (int)(tid % nex) << 3;
    yc = (int)(tid / nex) << 2;
    if (xc == 0) return;
  } else {
    const int nux = CW >> 2;
    const long long total = (long long)nux * ((CH + 7) >> 3);
    if (tid >= total) return;
    xc = (int)(tid % nux) << 2;
    yc = (int)(tid / nux) << 3;
    if (yc == 0) return;
  }
  const int lx = xc * SubW, ly = yc * SubH;  // luma position of this 4-sample unit
  const int e = edge[(lx >> 2) + (size_t)(ly >> 2) * w4];
  if (!(e & mask)) return;
  const int QP_Q = qp[(lx >> 3) + (size_t)(ly >> 3) * w8];
  const int QP_P = vertical ? qp[((lx - 1) >> 3) + (size_t)(ly >> 3) * w8] : qp[(lx >> 3) + (size_t)((ly - 1) >> 3) * w8];
  const int cQpPicOffset = plane_idx == 1 ? pic.pps_cb_qp_offset : pic.pps_cr_qp_offset;
  const int qPi = ((QP_Q + QP_P + 1) >> 1) + cQpPicOffset;
  const int QpC = pic.chroma_format == 1 ? qpc_from_qpi_420(qPi) : (qPi < 51 ? qPi : 51);
  // tc offset of the slice at the start of the 8-sample chroma segment (deblock.cc:1700-1701)
  const int sxc = vertical ? xc : (xc & ~7), syc = vertical ? (yc & ~7) : yc;
  const hc_ctu& ctu = ctus[((sxc * SubW) >> pic.log2_ctb) + ((syc * SubH) >> pic.log2_ctb) * pic.ctbs_w];
  const int bd = pic.bit_depth_c;
  const int tc = c_tc_tab[clip3i(0, 53, QpC + 2 + ctu.tc_offset)] * (1 << (bd - 8));
  const int maxv = (1 << bd) - 1;
  Pixel* pix = plane + xc + (size_t)yc * stride;
  const ptrdiff_t xs = vertical ? 1 : stride, ys = vertical ? stride : 1;
  bool no_p = false, no_q = false;
  if (pic.flags & HC_PIC_PCMF) {
    // deblock.cc:1716-1755 + loop_filter_chroma_c (fallback-postfilter.h:138-179)
    bool normal[2][2];
    const bool lfd = pic.flags & HC_PIC_PCM_LF_DISABLED;
    const int slx = sxc * SubW, sly = syc * SubH;  // luma position of the segment start
#pragma unroll
    for (int u = 0; u < 2; u++) {
      const int ux = vertical ? slx : slx + 4 * u * SubW, uy = vertical ? sly + 4 * u * SubH : sly;
      if (ux >= W || uy >= H) { normal[u][0] = normal[u][1] = true; continue; }
      const int upx = vertical ? ux - 1 : ux, upy = vertical ? uy : uy - 1;
      const int ep = edge[(upx >> 2) + (size_t)(upy >> 2) * w4], eq = edge[(ux >> 2) + (size_t)(uy >> 2) * w4];
      normal[u][0] = !((lfd && (ep & HC_EDGE_PCM)) || (ep & HC_EDGE_BYPASS));
      normal[u][1] = !((lfd && (eq & HC_EDGE_PCM)) || (eq & HC_EDGE_BYPASS));
    }
    if (!(normal[0][0] && normal[0][1] && normal[1][0] && normal[1][1])) {
      const int j = vertical ? ((yc >> 2) & 1) : ((xc >> 2) & 1);
      no_p = !normal[j][0];
      no_q = vertical ? !normal[j][0] : !normal[j][1];
    }
  }
#pragma unroll
  for (int k = 0; k < 4; k++) {
    Pixel* l = pix + (ptrdiff_t)k * ys;
    const int p1 = l[-2 * xs], p0 = l[-1 * xs], q0 = l[0], q1 = l[xs];
    const int delta = clip3i(-tc, tc, (((q0 - p0) * 4) + p1 - q1 + 4) >> 3);
    if (!no_p) l[-1 * xs] = (Pixel)clip3i(0, maxv, p0 + delta);
    if (!no_q) l[0] = (Pixel)clip3i(0, maxv, q0 - delta);
  }
}

template <bool VERTICAL>
__global__ void __launch_bounds__(256, 8) k3_deblock_kernel(BatchView bv) {
  const hc_pic& pic = bv.pics[blockIdx.y];
  if (!(pic.flags & HC_PIC_HAS_DEBLOCK)) return;
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (pic.bit_depth_y == 8 && pic.bit_depth_c == 8) deblock_picture<uint8_t>(bv, pic, blockIdx.z, VERTICAL, tid);
  else deblock_picture<uint16_t>(bv, pic, blockIdx.z, VERTICAL, tid);
}

// max_units = max over pictures of (width/8)*(height/4) (== (width/4)*(height/8))
void launch_k3(const BatchView& bv, long long max_units, int planes, cudaStream_t stream) {
  if (max_units <= 0 || bv.npics <= 0) return;
  dim3 grid((unsigned)((max_units + 255) / 256), (unsigned)bv.npics, (unsigned)planes);
  k3_deblock_kernel<true><<<grid, 256, 0, stream>>>(bv);
  k3_deblock_kernel<false><<<grid, 256, 0, stream>>>(bv);
}

}  // namespace hc
