// k34_postfilter.cu — K3+K4 fused: deblocking (vertical edges, then horizontal edges), SAO, conformance-window crop and
// paste into the destination, one tile at a time in shared memory. The reconstruction is read ONCE and the finished
// samples are written ONCE (the split kernels read and write every plane three times).
//
// Replaces apply_deblocking_filter (deblock.cc:1921-1959), apply_sample_adaptive_offset_sequential (sao.cc:552-625), the
// conformance-window copy (decoder_libde265.cc:88-157) and decode_and_paste_tile_image (context.cc:2407-2539); the
// per-unit arithmetic is postfilter_core.cuh, shared with the split kernels.
//
// Why a tile is self-contained. A tile is a 128 x 64 luma rectangle O (origin a multiple of 8) with the matching chroma
// rectangles. SAO of O reads deblocked samples of O grown by one. A deblocked sample depends on the reconstruction within
// 4 samples of the nearest 8-grid edge in x (vertical-edge pass) and then in y (horizontal-edge pass over the
// vertical-filtered samples). So the CTA loads R = O grown by 4 on every side, filters the vertical edges X, X+8, ..,
// X+128 over all rows of R (each edge's 8-sample footprint lies inside R), then the horizontal edges Y, Y+8, .., Y+64
// over all columns of R: afterwards every sample of O grown by ONE (by three, in fact) holds its final deblocked value,
// whatever the neighbouring tiles do. The halo edges are filtered by both neighbours — 20 % more filter work and
// reads that hit L2, against two full passes over HBM less.
//
// Mapping: one CTA of 256 threads per tile and picture, all colour planes in the same four phases (load, V, H, SAO+store),
// unit lists of the planes concatenated so that the chroma units fill the lanes the luma units leave over.
// Algorithmic bytes: read s + write s per sample (+ 1/16 B/px edge map, 1/64 B/px QP map).
#include <cstdlib>
#include "postfilter_core.cuh"

namespace hc {

constexpr int PF_TW = 128, PF_TH = 64;             // luma tile
constexpr int PF_PITCH = PF_TW + 8;                // samples per shared-memory row (halo of 4 on either side)
constexpr int PF_ROWS = PF_TH + 8;
constexpr int PF_THREADS = 256;

struct PfGeom {
  int X, Y;          // plane position of the tile's first output sample
  int tw, th;        // output tile size in this plane
  int PW, PH;        // coded plane size
  int on;            // plane present and wanted at the destination
};

template <typename Pixel>
__device__ __forceinline__ void copy4(Pixel* dst, const Pixel* src);
template <>
__device__ __forceinline__ void copy4<uint8_t>(uint8_t* dst, const uint8_t* src) {
  *reinterpret_cast<uint32_t*>(dst) = __ldg(reinterpret_cast<const uint32_t*>(src));
}
template <>
__device__ __forceinline__ void copy4<uint16_t>(uint16_t* dst, const uint16_t* src) {
  *reinterpret_cast<uint2*>(dst) = __ldg(reinterpret_cast<const uint2*>(src));
}

template <typename Pixel>
__device__ __forceinline__ void postfilter_tile(const BatchView& bv, const hc_pic& pic, const PfGeom* G, SaoPlane<Pixel>* SP, uint8_t* raw,
                                                int nplanes) {
  const int tid = threadIdx.x;
  Pixel* const sm = reinterpret_cast<Pixel*>(raw);
  constexpr int PLANE_ELEMS = PF_ROWS * PF_PITCH;
  if (tid < nplanes && G[tid].on) SP[tid].init(bv, pic, tid);

  // ---- phase 1: R (tile grown by 4) of every plane -> shared memory, 4 samples per access -------------------------
  // unit lists of the planes, concatenated: [0, e1) plane 0, [e1, e2) plane 1, [e2, e3) plane 2
  int e1, e2, e3;
  auto count = [&](auto units_of) {
    e1 = (nplanes > 0 && G[0].on) ? units_of(G[0]) : 0;
    e2 = e1 + ((nplanes > 1 && G[1].on) ? units_of(G[1]) : 0);
    e3 = e2 + ((nplanes > 2 && G[2].on) ? units_of(G[2]) : 0);
  };
  count([](const PfGeom& g) { return (g.th + 8) * ((g.tw + 8) >> 2); });
  for (int i = tid; i < e3; i += PF_THREADS) {
    const int p = (i >= e1) + (i >= e2);
    const PfGeom& g = G[p];
    const int upr = (g.tw + 8) >> 2, l = i - (p == 0 ? 0 : (p == 1 ? e1 : e2));
    const int r = l / upr, u = l - r * upr;
    const int x = g.X - 4 + 4 * u, y = g.Y - 4 + r;
    if (x < 0 || x >= g.PW || y < 0 || y >= g.PH) continue;
    const Pixel* src = reinterpret_cast<const Pixel*>(bv.planes + pic.rec_off[p]) + (size_t)y * pic.rec_stride[p] + x;
    copy4<Pixel>(sm + p * PLANE_ELEMS + r * PF_PITCH + 4 * u, src);
  }
  __syncthreads();

  if ((pic.flags & HC_PIC_HAS_DEBLOCK) && !(bv.flags & HC_VIEW_NO_DEBLOCK)) {
    // ---- phase 2: vertical edges x = X + 8e of R, 4-row units ------------------------------------------------------
    count([](const PfGeom& g) { return ((g.tw >> 3) + 1) * ((g.th + 8) >> 2); });
    for (int i = tid; i < e3; i += PF_THREADS) {
      const int p = (i >= e1) + (i >= e2);
      const PfGeom& g = G[p];
      const int ne = (g.tw >> 3) + 1, l = i - (p == 0 ? 0 : (p == 1 ? e1 : e2));
      const int r = l / ne, e = l - r * ne;
      const int x = g.X + 8 * e, y = g.Y - 4 + 4 * r;
      if (x == 0 || x >= g.PW || y < 0 || y >= g.PH) continue;
      deblock_unit_at<Pixel>(bv, pic, p, true, x, y, sm + p * PLANE_ELEMS + (4 * r) * PF_PITCH + 4 + 8 * e, PF_PITCH);
    }
    __syncthreads();
    // ---- phase 3: horizontal edges y = Y + 8e of R, 4-column units -------------------------------------------------
    count([](const PfGeom& g) { return ((g.th >> 3) + 1) * ((g.tw + 8) >> 2); });
    for (int i = tid; i < e3; i += PF_THREADS) {
      const int p = (i >= e1) + (i >= e2);
      const PfGeom& g = G[p];
      const int nu = (g.tw + 8) >> 2, l = i - (p == 0 ? 0 : (p == 1 ? e1 : e2));
      const int e = l / nu, u = l - e * nu;
      const int x = g.X - 4 + 4 * u, y = g.Y + 8 * e;
      if (y == 0 || y >= g.PH || x < 0 || x >= g.PW) continue;
      deblock_unit_at<Pixel>(bv, pic, p, false, x, y, sm + p * PLANE_ELEMS + (4 + 8 * e) * PF_PITCH + 4 * u, PF_PITCH);
    }
    __syncthreads();
  }

  // ---- phase 4: SAO + crop + paste of O, 8 samples per unit -----------------------------------------------------------
  count([](const PfGeom& g) { return (g.tw >> 3) * g.th; });
  for (int i = tid; i < e3; i += PF_THREADS) {
    const int p = (i >= e1) + (i >= e2);
    const PfGeom& g = G[p];
    const int upr = g.tw >> 3, l = i - (p == 0 ? 0 : (p == 1 ? e1 : e2));
    const int r = l / upr, u = l - r * upr;
    const int x0 = g.X + 8 * u, y = g.Y + r;
    if (x0 >= g.PW || y >= g.PH) continue;
    sao_unit<Pixel, true>(bv, pic, SP[p], x0, y, sm + p * PLANE_ELEMS + (4 + r) * PF_PITCH + 4 + 8 * u, PF_PITCH);
  }
}

template <int MINB>
__global__ void __launch_bounds__(PF_THREADS, MINB) k34_postfilter_kernel(BatchView bv, int nplanes_max) {
  extern __shared__ __align__(16) uint8_t pf_smem[];
  __shared__ hc_pic spic;
  __shared__ PfGeom G[3];
  __shared__ __align__(8) uint8_t sp_raw[3 * sizeof(SaoPlane<uint16_t>)];
  {
    const uint32_t* src = reinterpret_cast<const uint32_t*>(bv.pics + blockIdx.y);
    uint32_t* dst = reinterpret_cast<uint32_t*>(&spic);
    for (int i = threadIdx.x; i < (int)(sizeof(hc_pic) / 4); i += blockDim.x) dst[i] = src[i];
  }
  __syncthreads();
  const hc_pic& pic = spic;
  const int tiles_x = (pic.width + PF_TW - 1) / PF_TW, tiles_y = (pic.height + PF_TH - 1) / PF_TH;
  if ((int)blockIdx.x >= tiles_x * tiles_y) return;
  const int tiy = blockIdx.x / tiles_x, tix = blockIdx.x - tiy * tiles_x;
  const int nplanes = pic.chroma_format == 0 ? 1 : min(3, nplanes_max);
  if (threadIdx.x < 3) {
    const int p = threadIdx.x;
    const int sw = (p && (pic.chroma_format == 1 || pic.chroma_format == 2)) ? 1 : 0, sh = (p && pic.chroma_format == 1) ? 1 : 0;
    PfGeom g;
    g.tw = PF_TW >> sw; g.th = PF_TH >> sh;
    g.X = tix * g.tw; g.Y = tiy * g.th;
    g.PW = pic.width >> sw; g.PH = pic.height >> sh;
    g.on = p < nplanes && !(pic.dst_flags & (HC_DST_SKIP_Y << p));   // component not wanted at the destination
    G[p] = g;
  }
  __syncthreads();
  if (pic.bit_depth_y == 8 && pic.bit_depth_c == 8)
    postfilter_tile<uint8_t>(bv, pic, G, reinterpret_cast<SaoPlane<uint8_t>*>(sp_raw), pf_smem, nplanes);
  else
    postfilter_tile<uint16_t>(bv, pic, G, reinterpret_cast<SaoPlane<uint16_t>*>(sp_raw), pf_smem, nplanes);
}

// max_tiles = max over pictures of ceil(width / 128) * ceil(height / 64); sixteen_bit: some picture of the batch has
// samples of more than 8 bits (sizes the shared-memory tile)
void launch_k34(const BatchView& bv, long long max_tiles, int planes, bool sixteen_bit, cudaStream_t stream) {
  if (max_tiles <= 0 || bv.npics <= 0) return;
  const int smem = 3 * PF_ROWS * PF_PITCH * (sixteen_bit ? 2 : 1);
  // above the 48 KB default only for 16-bit samples; the attribute belongs to the current device, so it is set per launch
  static const int occ = getenv("HEIFCUDA_PF_OCC") ? atoi(getenv("HEIFCUDA_PF_OCC")) : 3;
  dim3 grid((unsigned)max_tiles, (unsigned)bv.npics);
  if (occ == 4) {
    if (smem > 48 * 1024) cudaFuncSetAttribute(k34_postfilter_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    k34_postfilter_kernel<4><<<grid, PF_THREADS, smem, stream>>>(bv, planes);
  } else if (occ == 2) {
    if (smem > 48 * 1024) cudaFuncSetAttribute(k34_postfilter_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    k34_postfilter_kernel<2><<<grid, PF_THREADS, smem, stream>>>(bv, planes);
  } else {
    if (smem > 48 * 1024) cudaFuncSetAttribute(k34_postfilter_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    k34_postfilter_kernel<3><<<grid, PF_THREADS, smem, stream>>>(bv, planes);
  }
}

}  // namespace hc
