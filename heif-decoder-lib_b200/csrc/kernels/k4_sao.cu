// k4_sao.cu — K4: sample adaptive offset + conformance-window crop + paste into the destination
// image (own picture or HEIF grid canvas) in ONE pass.
//
// Replaces apply_sample_adaptive_offset_sequential (sao.cc:552-625: full-plane copy, then
// apply_sao_internal :261-488 with sao_band_filter / sao_edge_filter, fallback-postfilter.h:216-315),
// the conformance-window copy of convert_libde265_image_to_heif_image (decoder_libde265.cc:88-157)
// and the tile paste of decode_and_paste_tile_image (context.cc:2407-2539) including its
// limited->full range rescale of limited-range tiles (:2504-2528).
//
// The reference needs a deblocked copy of every plane because it filters in place; here the
// deblocked reconstruction planes are read-only input and the result is written straight to its
// final position, so the copy, the crop and the paste cost no extra memory pass.
//
// Mapping: one thread per 4 horizontally adjacent samples; a warp covers 128 consecutive samples
// of one row. Algorithmic bytes: read s + write s per sample (neighbour rows hit L1/L2).
#include "launch.h"

namespace hc {

HC_D int sign3(int v) { return (v > 0) - (v < 0); }

template <typename Pixel>
__device__ void sao_picture(const BatchView& bv, const hc_pic& pic, int c, long long tid) {
  const int SubW = (c && (pic.chroma_format == 1 || pic.chroma_format == 2)) ? 2 : 1;
  const int SubH = (c && pic.chroma_format == 1) ? 2 : 1;
  const int width = pic.width / SubW, height = pic.height / SubH;  // coded plane size
  const int nq = (width + 3) >> 2;
  if (tid >= (long long)nq * height) return;
  const int y = (int)(tid / nq), xq = (int)(tid % nq) << 2;

  // crop window and destination clip, in samples of this plane (context.cc:2467-2497)
  const int cx0 = pic.crop_x / SubW, cy0 = pic.crop_y / SubH;
  const int cw = (pic.crop_w + SubW - 1) / SubW, ch = (pic.crop_h + SubH - 1) / SubH;
  const int dx0 = (pic.dst_x + SubW - 1) / SubW, dy0 = (pic.dst_y + SubH - 1) / SubH;
  const int dw = (pic.dst_w + SubW - 1) / SubW, dh = (pic.dst_h + SubH - 1) / SubH;
  const int copy_w = min(cw, dw - dx0), copy_h = min(ch, dh - dy0);
  const int oy = y - cy0;
  if (oy < 0 || oy >= copy_h) return;

  const Pixel* __restrict__ src = reinterpret_cast<const Pixel*>(bv.planes + pic.rec_off[c]);
  const int sstride = (int)pic.rec_stride[c];
  Pixel* __restrict__ dst = reinterpret_cast<Pixel*>(bv.planes + pic.dst_off[c]);
  const int dstride = (int)pic.dst_stride[c];
  const int bit_depth = c == 0 ? pic.bit_depth_y : pic.bit_depth_c;
  const int maxv = (1 << bit_depth) - 1;
  const int ctb_w = (1 << pic.log2_ctb) / SubW, ctb_h = (1 << pic.log2_ctb) / SubH;
  const int log2w = pic.log2_ctb - (SubW == 2), log2h = pic.log2_ctb - (SubH == 2);
  const hc_ctu* __restrict__ ctus = bv.ctus + pic.ctu_base;
  const uint8_t* __restrict__ edge = bv.edge_map + pic.edge_base;
  const int w4 = pic.width >> 2;
  const bool rescale = pic.dst_flags & HC_DST_RESCALE_LIMITED;
  (void)ctb_w; (void)ctb_h;

  const Pixel* row = src + (size_t)y * sstride;
#pragma unroll
  for (int k = 0; k < 4; k++) {
    const int x = xq + k;
    const int ox = x - cx0;
    if (x >= width || ox < 0 || ox >= copy_w) continue;
    const int v0 = row[x];
    int v = v0;
    const int ctbx = x >> log2w, ctby = y >> log2h;
    const hc_ctu& ctu = ctus[ctbx + ctby * pic.ctbs_w];
    const int type = (bv.flags & HC_VIEW_NO_SAO) ? 0 : ctu.sao_type[c];
    if (type) {
      bool skip = false;
      if (ctu.flags & HC_CTU_HAS_NOFILTER) {
        const int e = edge[((x * SubW) >> 2) + (size_t)((y * SubH) >> 2) * w4];
        skip = ((pic.flags & HC_PIC_PCM_LF_DISABLED) && (e & HC_EDGE_PCM)) || (e & HC_EDGE_BYPASS);
      }
      if (!skip) {
        if (type == 1) {
          // bandShift >= 8 leaves the sample untouched in the reference (sao.cc:461)
          if (bit_depth - 5 < 8) {
            const int band = v0 >> (bit_depth - 5);
            const int k4 = (band - ctu.sao_band_or_class[c]) & 31;
            if (k4 < 4) v = clip3i(0, maxv, v0 + ctu.sao_offset[c][k4]);
          }
        } else {
          const int cls = ctu.sao_band_or_class[c];
          const int hx = cls == 1 ? 0 : (cls == 3 ? 1 : -1);   // first neighbour
          const int vy = cls == 0 ? 0 : -1;
          // neighbours: (x+hx, y+vy) and (x-hx, y-vy)
          bool ok = true;
#pragma unroll
          for (int n = 0; n < 2; n++) {
            const int xs = n == 0 ? x + hx : x - hx, ys = n == 0 ? y + vy : y - vy;
            if (xs < 0 || ys < 0 || xs >= width || ys >= height) { ok = false; continue; }
            const int dxc = (xs >> log2w) - ctbx, dyc = (ys >> log2h) - ctby;
            if (dxc | dyc) {
              int bit;
              if (dyc == 0) bit = dxc < 0 ? HC_NB_L : HC_NB_R;
              else if (dxc == 0) bit = dyc < 0 ? HC_NB_T : HC_NB_B;
              else if (dyc < 0) bit = dxc < 0 ? HC_NB_TL : HC_NB_TR;
              else bit = dxc < 0 ? HC_NB_BL : HC_NB_BR;
              if (!((c ? ctu.sao_nb_c : ctu.sao_nb) & bit)) ok = false;
            } else if (c && (ctu.flags & HC_CTU_SAO_C_SELF)) {
              // reference quirk (sao.cc:283): border samples of this CTB also lose their in-CTB neighbours
              const int lx = x & ((1 << log2w) - 1), ly = y & ((1 << log2h) - 1);
              const int cwid = min(1 << log2w, width - (ctbx << log2w)), chei = min(1 << log2h, height - (ctby << log2h));
              if (lx == 0 || ly == 0 || lx == cwid - 1 || ly == chei - 1) ok = false;
            }
          }
          if (ok) {
            const int a = src[(size_t)(y + vy) * sstride + x + hx];
            const int b = src[(size_t)(y - vy) * sstride + x - hx];
            const int e = sign3(v0 - a) + sign3(v0 - b);   // -2..2
            // edgeIdx -2,-1,1,2 -> offsets 0,1,2,3 (sao.cc:312-317)
            if (e) v = clip3i(0, maxv, v0 + ctu.sao_offset[c][e < 0 ? e + 2 : e + 1]);
          }
        }
      }
    }
    if (rescale) {
      // context.cc:2504-2528: bytewise float rescale of limited-range tiles, no FMA contraction
      const float ratio = c == 0 ? 1.1689f : 1.1429f;
      const float full = __fmul_rn(__fsub_rn((float)v, (float)(16 << (bit_depth - 8))), ratio);
      const long r = (long)__fadd_rn(full, 0.5f);
      v = r < 0 ? 0 : (r > 255 ? 255 : (int)r);
    }
    dst[(size_t)(dy0 + oy) * dstride + dx0 + ox] = (Pixel)v;
  }
}

__global__ void __launch_bounds__(256) k4_sao_kernel(BatchView bv) {
  const hc_pic& pic = bv.pics[blockIdx.y];
  const int c = blockIdx.z;
  if (c > 0 && pic.chroma_format == 0) return;
  if (pic.dst_flags & (HC_DST_SKIP_Y << c)) return;   // component not wanted at the destination
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (pic.bit_depth_y == 8 && pic.bit_depth_c == 8) sao_picture<uint8_t>(bv, pic, c, tid);
  else sao_picture<uint16_t>(bv, pic, c, tid);
}

// max_quads = max over pictures of ceil(width/4)*height
void launch_k4(const BatchView& bv, long long max_quads, int planes, cudaStream_t stream) {
  if (max_quads <= 0 || bv.npics <= 0) return;
  dim3 grid((unsigned)((max_quads + 255) / 256), (unsigned)bv.npics, (unsigned)planes);
  k4_sao_kernel<<<grid, 256, 0, stream>>>(bv);
}

}  // namespace hc
