// k5_csc.cu — K5: libheif's YCbCr -> interleaved RGB colour-conversion chain fused into one
// output-writing kernel (chroma upsampling, matrix, range expansion, interleave, endianness,
// alpha merge).
//
// Replaces, depending on hc_csc_params::mode (selection mirrors ColorConversionPipeline::
// construct_pipeline, colorconversion.cc:266-420, for the default decoding options):
//   HC_CSC_INT420 : Op_YCbCr420_to_RGB24 / Op_YCbCr420_to_RGB32            yuv2rgb.cc:260-495
//   HC_CSC_FLOAT  : Op_YCbCr_to_RGB<uint8_t|uint16_t> (+ Op_RGB_to_RGB24_32, Op_RGB_HDR_to_RRGGBBaa_BE,
//                   Op_RRGGBBaa_swap_endianness) and Op_YCbCr420_to_RRGGBBaa  yuv2rgb.cc:28-254,498-643,
//                   rgb2rgb.cc:28-272,613-729 — each of which is a separate full-image pass with
//                   its own allocation in the reference
//   HC_CSC_GBR / HC_CSC_YCGCO : the matrix_coefficients 0 / 8 branches        yuv2rgb.cc:197-226
// Float arithmetic uses explicit round-to-nearest mul/add (no FMA contraction) and the
// reference's (long)(x + 0.5f) rounding (common_utils.h:64-70) so results are bit-exact.
//
// Mapping: one thread converts 4 horizontally adjacent pixels and writes them with 32-bit / 128-bit
// stores; a warp writes 384 (RGB) .. 1024 (RRGGBBAA) contiguous bytes. Pure streaming:
// algorithmic bytes = s*c (+s alpha) read + bytes-per-pixel written.
#include "launch.h"

namespace hc {

HC_D int clip_f(float fx, int maxi) {
  const long long x = (long long)__fadd_rn(fx, 0.5f);
  return x < 0 ? 0 : (x > maxi ? maxi : (int)x);
}

template <typename Pixel>
__device__ void convert_px(const CscArgs& a, int yv, int cbv, int crv, int& r, int& g, int& b) {
  const hc_csc_params& p = a.p;
  const int bpp = p.bit_depth;
  const int maxv = (1 << bpp) - 1, half = 1 << (bpp - 1);
  if (p.mode == HC_CSC_INT420) {
    const int cb = cbv - 128, cr = crv - 128;
    r = clip3i(0, 255, yv + ((p.r_cr_i * cr + 128) >> 8));
    g = clip3i(0, 255, yv + ((p.g_cb_i * cb + p.g_cr_i * cr + 128) >> 8));
    b = clip3i(0, 255, yv + ((p.b_cb_i * cb + 128) >> 8));
  } else if (p.mode == HC_CSC_FLOAT) {
    float fy = (float)yv, cb = (float)(cbv - half), cr = (float)(crv - half);
    if (!p.full_range) {
      fy = __fmul_rn(__fsub_rn(fy, (float)(16 << (bpp - 8))), 1.1689f);
      cb = __fmul_rn(cb, 1.1429f);
      cr = __fmul_rn(cr, 1.1429f);
    }
    r = clip_f(__fadd_rn(fy, __fmul_rn(p.r_cr, cr)), maxv);
    g = clip_f(__fadd_rn(__fadd_rn(fy, __fmul_rn(p.g_cb, cb)), __fmul_rn(p.g_cr, cr)), maxv);
    b = clip_f(__fadd_rn(fy, __fmul_rn(p.b_cb, cb)), maxv);
  } else if (p.mode == HC_CSC_GBR) {
    if (p.full_range) { r = crv; g = yv; b = cbv; }
    else {
      const float off = (float)(16 << (bpp - 8));
      r = clip_f(__fmul_rn(__fsub_rn((float)crv, off), 1.1429f), maxv);
      g = clip_f(__fmul_rn(__fsub_rn((float)yv, off), 1.1689f), maxv);
      b = clip_f(__fmul_rn(__fsub_rn((float)cbv, off), 1.1429f), maxv);
    }
  } else {  // HC_CSC_YCGCO (clip_int_u8 in the reference regardless of depth, yuv2rgb.cc:218-226)
    const int cb = cbv - half, cr = crv - half;
    r = clip3i(0, 255, yv - cb + cr);
    g = clip3i(0, 255, yv + cb);
    b = clip3i(0, 255, yv - cb - cr);
  }
}

template <typename Pixel>
__global__ void __launch_bounds__(256) k5_csc_kernel(CscArgs a) {
  const int nq = (a.width + 3) >> 2;
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (tid >= (long long)nq * a.height) return;
  const int y = (int)(tid / nq), x0 = (int)(tid % nq) << 2;
  const int shiftH = (a.chroma_format == 1 || a.chroma_format == 2) ? 1 : 0;
  const int shiftV = a.chroma_format == 1 ? 1 : 0;
  const Pixel* __restrict__ py = reinterpret_cast<const Pixel*>(a.y) + (size_t)y * a.y_stride;
  const Pixel* __restrict__ pcb = a.chroma_format ? reinterpret_cast<const Pixel*>(a.cb) + (size_t)(y >> shiftV) * a.c_stride : nullptr;
  const Pixel* __restrict__ pcr = a.chroma_format ? reinterpret_cast<const Pixel*>(a.cr) + (size_t)(y >> shiftV) * a.c_stride : nullptr;
  const Pixel* __restrict__ pa = a.a ? reinterpret_cast<const Pixel*>(a.a) + (size_t)y * a.a_stride : nullptr;
  const int bpp = a.p.bit_depth;
  const int fmt = a.p.out_format;
  const int half = 1 << (bpp - 1);

  int R[4], G[4], B[4], A[4];
#pragma unroll
  for (int k = 0; k < 4; k++) {
    const int x = x0 + k;
    const int yv = py[x];
    int cbv = half, crv = half;
    if (a.chroma_format) { cbv = pcb[x >> shiftH]; crv = pcr[x >> shiftH]; }
    if (a.chroma_format) convert_px<Pixel>(a, yv, cbv, crv, R[k], G[k], B[k]);
    else { R[k] = G[k] = B[k] = yv; }
    A[k] = pa ? (int)pa[x] : ((1 << bpp) - 1);
  }

  uint8_t* orow = a.out + (size_t)y * a.out_stride;
  if (fmt == HC_OUT_RGB) {
    uint32_t w0 = R[0] | (G[0] << 8) | (B[0] << 16) | (R[1] << 24);
    uint32_t w1 = G[1] | (B[1] << 8) | (R[2] << 16) | (G[2] << 24);
    uint32_t w2 = B[2] | (R[3] << 8) | (G[3] << 16) | (B[3] << 24);
    uint32_t* o = reinterpret_cast<uint32_t*>(orow + (size_t)x0 * 3);
    o[0] = w0; o[1] = w1; o[2] = w2;
  } else if (fmt == HC_OUT_RGBA) {
    uint4 v;
    v.x = R[0] | (G[0] << 8) | (B[0] << 16) | ((uint32_t)A[0] << 24);
    v.y = R[1] | (G[1] << 8) | (B[1] << 16) | ((uint32_t)A[1] << 24);
    v.z = R[2] | (G[2] << 8) | (B[2] << 16) | ((uint32_t)A[2] << 24);
    v.w = R[3] | (G[3] << 8) | (B[3] << 16) | ((uint32_t)A[3] << 24);
    *reinterpret_cast<uint4*>(orow + (size_t)x0 * 4) = v;
  } else {
    const bool le = (fmt == HC_OUT_RRGGBB_LE || fmt == HC_OUT_RRGGBBAA_LE);
    const bool alpha = (fmt == HC_OUT_RRGGBBAA_BE || fmt == HC_OUT_RRGGBBAA_LE);
    auto h16 = [&](int v) -> uint32_t { return le ? (uint32_t)v : (uint32_t)(((v & 0xff) << 8) | (v >> 8)); };
    if (alpha) {
      uint4* o = reinterpret_cast<uint4*>(orow + (size_t)x0 * 8);
#pragma unroll
      for (int k = 0; k < 2; k++) {
        uint4 v;
        v.x = h16(R[2 * k]) | (h16(G[2 * k]) << 16);
        v.y = h16(B[2 * k]) | (h16(A[2 * k]) << 16);
        v.z = h16(R[2 * k + 1]) | (h16(G[2 * k + 1]) << 16);
        v.w = h16(B[2 * k + 1]) | (h16(A[2 * k + 1]) << 16);
        o[k] = v;
      }
    } else {
      uint32_t* o = reinterpret_cast<uint32_t*>(orow + (size_t)x0 * 6);
      o[0] = h16(R[0]) | (h16(G[0]) << 16);
      o[1] = h16(B[0]) | (h16(R[1]) << 16);
      o[2] = h16(G[1]) | (h16(B[1]) << 16);
      o[3] = h16(R[2]) | (h16(G[2]) << 16);
      o[4] = h16(B[2]) | (h16(R[3]) << 16);
      o[5] = h16(G[3]) | (h16(B[3]) << 16);
    }
  }
}

void launch_k5(const CscArgs& a, bool sixteen_bit, cudaStream_t stream) {
  const long long n = (long long)((a.width + 3) >> 2) * a.height;
  if (n <= 0) return;
  const unsigned grid = (unsigned)((n + 255) / 256);
  if (sixteen_bit) k5_csc_kernel<uint16_t><<<grid, 256, 0, stream>>>(a);
  else k5_csc_kernel<uint8_t><<<grid, 256, 0, stream>>>(a);
}

}  // namespace hc
