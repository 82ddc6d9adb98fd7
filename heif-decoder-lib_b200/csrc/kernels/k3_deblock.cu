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
#include "postfilter_core.cuh"

namespace hc {

// plane: 0 luma, 1 Cb, 2 Cr (blockIdx.z); picture: blockIdx.y; tid -> unit in raster order
template <typename Pixel>
__device__ void deblock_picture(const BatchView& bv, const hc_pic& pic, int plane_idx, bool vertical, long long tid) {
  const int W = pic.width, H = pic.height;
  if (plane_idx > 0 && pic.chroma_format == 0) return;
  const int SubW = (plane_idx && (pic.chroma_format == 1 || pic.chroma_format == 2)) ? 2 : 1;
  const int SubH = (plane_idx && pic.chroma_format == 1) ? 2 : 1;
  const int PW = W / SubW, PH = H / SubH;
  int x, y;  // plane position of the unit's first q0 sample
  if (vertical) {
    const int nex = (PW + 7) >> 3;
    const long long total = (long long)nex * (PH >> 2);
    if (tid >= total) return;
    x = (int)(tid % nex) << 3;
    y = (int)(tid / nex) << 2;
    if (x == 0) return;
  } else {
    const int nux = PW >> 2;
    const long long total = (long long)nux * ((PH + 7) >> 3);
    if (tid >= total) return;
    x = (int)(tid % nux) << 2;
    y = (int)(tid / nux) << 3;
    if (y == 0) return;
  }
  Pixel* plane = reinterpret_cast<Pixel*>(bv.planes + pic.rec_off[plane_idx]);
  const int stride = (int)pic.rec_stride[plane_idx];
  deblock_unit_at<Pixel>(bv, pic, plane_idx, vertical, x, y, plane + x + (size_t)y * stride, stride);
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
