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
// Mapping: one thread per 8 horizontally adjacent samples (8-aligned, so a unit never straddles a
// CTB: CTBs are >= 8 samples wide in every component); the unit's row and, for the edge classes,
// the rows above / below are fetched with one 8- or 16-byte load each. A warp covers a
// (CTB width) x (256 / CTB width) patch of ONE CTB, so SAO type, class and offsets are warp-uniform.
// Algorithmic bytes: read s + write s per sample (neighbour rows hit L1/L2).
#include "launch.h"

namespace hc {

HC_D int sign3(int v) { return (v > 0) - (v < 0); }

template <typename Pixel>
HC_D void load8(const Pixel* p, int v[8]);
template <>
HC_D void load8<uint8_t>(const uint8_t* p, int v[8]) {
  const uint2 w = *reinterpret_cast<const uint2*>(p);
#pragma unroll
  for (int k = 0; k < 4; k++) { v[k] = (w.x >> (8 * k)) & 0xff; v[4 + k] = (w.y >> (8 * k)) & 0xff; }
}
template <>
HC_D void load8<uint16_t>(const uint16_t* p, int v[8]) {
  const uint4 w = *reinterpret_cast<const uint4*>(p);
  v[0] = w.x & 0xffff; v[1] = w.x >> 16; v[2] = w.y & 0xffff; v[3] = w.y >> 16;
  v[4] = w.z & 0xffff; v[5] = w.z >> 16; v[6] = w.w & 0xffff; v[7] = w.w >> 16;
}
HC_D void store8(uint8_t* p, const int v[8]) {
  uint2 w;
  w.x = v[0] | (v[1] << 8) | (v[2] << 16) | (v[3] << 24);
  w.y = v[4] | (v[5] << 8) | (v[6] << 16) | (v[7] << 24);
  *reinterpret_cast<uint2*>(p) = w;
}
HC_D void store8(uint16_t* p, const int v[8]) {
  uint4 w;
  w.x = v[0] | (v[1] << 16); w.y = v[2] | (v[3] << 16); w.z = v[4] | (v[5] << 16); w.w = v[6] | (v[7] << 16);
  *reinterpret_cast<uint4*>(p) = w;
}

// One CTB of one colour plane by one warp: the geometry (crop / paste windows, CTB position, SAO parameters) is set up
// once, then the warp walks the CTB's row groups. lane -> (8-sample unit in the CTB row, row inside the group).
template <typename Pixel>
__device__ __forceinline__ void sao_ctb(const BatchView& bv, const hc_pic& pic, int c, unsigned ctb, int lane) {
  const int sw = (c && (pic.chroma_format == 1 || pic.chroma_format == 2)) ? 1 : 0;   // log2 subsampling
  const int sh = (c && pic.chroma_format == 1) ? 1 : 0;
  const int SubW = 1 << sw, SubH = 1 << sh;
  const int width = pic.width >> sw, height = pic.height >> sh;  // coded plane size (multiples of 4)
  const int log2w = pic.log2_ctb - sw, log2h = pic.log2_ctb - sh;
  // a warp pass covers (CTB width / 8) units x (256 / CTB width) rows of ONE CTB, so the SAO type / class /
  // offsets are warp-uniform and only the CTB-border handling differs between lanes
  const int lu = log2w - 3;                          // log2 units per CTB row (0..3)
  const int rows = 32 >> lu;                         // rows per warp pass
  const int ctby = (int)(ctb / pic.ctbs_w), ctbx = (int)(ctb - (unsigned)ctby * pic.ctbs_w);
  const int x0 = (ctbx << log2w) + ((lane & ((1 << lu) - 1)) << 3);
  if (x0 >= width) return;

  // crop window and destination clip, in samples of this plane (context.cc:2467-2497)
  const int cx0 = pic.crop_x >> sw, cy0 = pic.crop_y >> sh;
  const int cw = (pic.crop_w + SubW - 1) >> sw, ch = (pic.crop_h + SubH - 1) >> sh;
  const int dx0 = (pic.dst_x + SubW - 1) >> sw, dy0 = (pic.dst_y + SubH - 1) >> sh;
  const int dw = (pic.dst_w + SubW - 1) >> sw, dh = (pic.dst_h + SubH - 1) >> sh;
  const int copy_w = min(cw, dw - dx0), copy_h = min(ch, dh - dy0);
  const int ox0 = x0 - cx0;                         // destination column of sample 0 of the unit
  if (ox0 + 8 <= 0 || ox0 >= copy_w) return;

  const Pixel* __restrict__ src = reinterpret_cast<const Pixel*>(bv.planes + pic.rec_off[c]);
  const int sstride = (int)pic.rec_stride[c];
  Pixel* __restrict__ dst = reinterpret_cast<Pixel*>(bv.planes + pic.dst_off[c]);
  const int dstride = (int)pic.dst_stride[c];
  const int bit_depth = c == 0 ? pic.bit_depth_y : pic.bit_depth_c;
  const int maxv = (1 << bit_depth) - 1;
  const hc_ctu& ctu = bv.ctus[pic.ctu_base + ctbx + ctby * pic.ctbs_w];
  const int nvalid = min(8, width - x0);            // 4 or 8
  const int y_end = min(height, (ctby + 1) << log2h);

  for (int y = (ctby << log2h) + (lane >> lu); y < y_end; y += rows) {
  const int oy = y - cy0;
  if (oy < 0 || oy >= copy_h) continue;
  const Pixel* row = src + (size_t)y * sstride + x0;
  int v[8];
  if (nvalid == 8) load8<Pixel>(row, v);
  else {
#pragma unroll
    for (int k = 0; k < 8; k++) v[k] = k < nvalid ? (int)row[k] : 0;
  }

  const int type = (bv.flags & HC_VIEW_NO_SAO) ? 0 : ctu.sao_type[c];
  if (type) {
    // samples of pcm (with pcm_loop_filter_disabled) / transquant-bypass CUs are left alone (sao.cc:288-300)
    unsigned skip = 0;
    if (ctu.flags & HC_CTU_HAS_NOFILTER) {
      const uint8_t* __restrict__ edge = bv.edge_map + pic.edge_base;
      const int w4 = pic.width >> 2;
#pragma unroll
      for (int k = 0; k < 8; k++) {
        if (k >= nvalid) continue;
        const int e = edge[(((x0 + k) * SubW) >> 2) + (size_t)((y * SubH) >> 2) * w4];
        if (((pic.flags & HC_PIC_PCM_LF_DISABLED) && (e & HC_EDGE_PCM)) || (e & HC_EDGE_BYPASS)) skip |= 1u << k;
      }
    }
    const int8_t* offs = ctu.sao_offset[c];
    const int o0 = offs[0], o1 = offs[1], o2 = offs[2], o3 = offs[3];
    if (type == 1) {
      // bandShift >= 8 leaves the sample untouched in the reference (sao.cc:461)
      if (bit_depth - 5 < 8) {
        const int pos = ctu.sao_band_or_class[c];
#pragma unroll
        for (int k = 0; k < 8; k++) {
          const int k4 = ((v[k] >> (bit_depth - 5)) - pos) & 31;
          if (k4 < 4 && !((skip >> k) & 1)) v[k] = clip3i(0, maxv, v[k] + (k4 == 0 ? o0 : k4 == 1 ? o1 : k4 == 2 ? o2 : o3));
        }
      }
    } else {
      const int cls = ctu.sao_band_or_class[c];
      const int hx = cls == 1 ? 0 : (cls == 3 ? 1 : -1);   // first neighbour (x+hx, y+vy), second (x-hx, y-vy)
      const int vy = cls == 0 ? 0 : -1;
      // rows of the two neighbours, columns x0-1 .. x0+8 (index + 1)
      int ra[10], rb[10];
      const bool have_up = y > 0, have_dn = y + 1 < height;
      const Pixel* rowa = vy ? row - sstride : row;
      const Pixel* rowb = vy ? row + sstride : row;
      const bool oka = vy ? have_up : true, okb = vy ? have_dn : true;
#pragma unroll
      for (int k = 0; k < 10; k++) { ra[k] = 0; rb[k] = 0; }
      if (vy == 0) {
#pragma unroll
        for (int k = 0; k < 8; k++) { ra[k + 1] = v[k]; rb[k + 1] = v[k]; }
      } else {
        if (oka) {
          if (nvalid == 8) load8<Pixel>(rowa, ra + 1);
          else for (int k = 0; k < nvalid; k++) ra[k + 1] = rowa[k];
        }
        if (okb) {
          if (nvalid == 8) load8<Pixel>(rowb, rb + 1);
          else for (int k = 0; k < nvalid; k++) rb[k + 1] = rowb[k];
        }
      }
      if (hx) {
        if (x0 > 0) { if (oka) ra[0] = rowa[-1]; if (okb) rb[0] = rowb[-1]; }
        if (x0 + 8 < width) { if (oka) ra[9] = rowa[8]; if (okb) rb[9] = rowb[8]; }   // nvalid == 8 here
      }
      const unsigned nb = c ? ctu.sao_nb_c : ctu.sao_nb;
      const bool self_quirk = c && (ctu.flags & HC_CTU_SAO_C_SELF);
      const int mw = (1 << log2w) - 1, mh = (1 << log2h) - 1;
      const int orig[8] = {v[0], v[1], v[2], v[3], v[4], v[5], v[6], v[7]};
      // offsets for edgeIdx -2,-1,(0),1,2 packed as bytes (sao.cc:312-317)
      const unsigned long long packed = (unsigned long long)(uint8_t)o0 | ((unsigned long long)(uint8_t)o1 << 8) |
                                        ((unsigned long long)(uint8_t)o2 << 24) | ((unsigned long long)(uint8_t)o3 << 32);
      // interior unit: no sample's neighbour leaves the CTB (or the picture) in the direction of this class
      const bool in_x = hx == 0 || ((x0 & mw) != 0 && ((x0 + 8) & mw) != 0 && x0 + 8 < width);
      const bool in_y = vy == 0 || ((y & mh) != 0 && ((y + 1) & mh) != 0 && y + 1 < height);
      // One arithmetic path for every unit of the warp: an interior unit uses all eight samples, a border unit only those
      // whose neighbours are usable (okmask). The first version ran a fast loop for interior units and a second, branchy
      // loop for border units — a warp holds both kinds (units 0 and 7 of every 64-sample CTB row are border units for
      // three of the four edge classes), so it executed both loops one after the other.
      unsigned okmask = 0;
      if (nvalid == 8 && skip == 0 && in_x && in_y && !self_quirk) {
        okmask = 0xffu;
      } else {
        const int lwid = min(1 << log2w, width - (ctbx << log2w)), lhei = min(1 << log2h, height - (ctby << log2h));
        const int ly = y & mh;
        // may sample x0 + k use both of its neighbours? (picture edge, neighbouring CTB not usable, reference quirk)
        auto sample_ok = [&](int k) -> bool {
          const int x = x0 + k;
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
              if (!(nb & bit)) ok = false;
            } else if (self_quirk) {
              // reference quirk (sao.cc:283): border samples of this CTB also lose their in-CTB neighbours
              const int lx = x & mw;
              if (lx == 0 || ly == 0 || lx == lwid - 1 || ly == lhei - 1) ok = false;
            }
          }
          return ok;
        };
        if (nvalid == 8 && !self_quirk) {
          // A full unit lies inside one CTB column. Which CTB does a neighbour fall into? 3 x 3 bits, (dy + 1) * 3 + dx + 1,
          // the centre (this CTB) always usable, the others from the CTB's neighbour mask (a CTB beyond the picture edge
          // has no bit; the edge of a partial CTB counts as its border). Samples 1..6 can only leave the CTB vertically,
          // sample 0 also to the left, sample 7 also to the right.
          const unsigned grid9 = 0x10u | ((nb & HC_NB_TL) ? 0x001u : 0u) | ((nb & HC_NB_T) ? 0x002u : 0u) | ((nb & HC_NB_TR) ? 0x004u : 0u) |
                                 ((nb & HC_NB_L) ? 0x008u : 0u) | ((nb & HC_NB_R) ? 0x020u : 0u) | ((nb & HC_NB_BL) ? 0x040u : 0u) |
                                 ((nb & HC_NB_B) ? 0x080u : 0u) | ((nb & HC_NB_BR) ? 0x100u : 0u);
          const int cy1 = (vy && ly == 0) ? -1 : 0, cy2 = (vy && ly == lhei - 1) ? 1 : 0;   // first neighbour looks up, second down
          auto okbit = [&](int cx, int cy) -> unsigned { return (grid9 >> ((cy + 1) * 3 + cx + 1)) & 1u; };
          okmask = (okbit(0, cy1) & okbit(0, cy2)) ? 0xffu : 0u;
          if (hx) {
            const int lx0 = x0 & mw;
            if (lx0 == 0) {
              const unsigned o0 = hx < 0 ? (okbit(-1, cy1) & okbit(0, cy2)) : (okbit(0, cy1) & okbit(-1, cy2));
              okmask = (okmask & ~1u) | o0;
            }
            if (lx0 + 8 == lwid) {
              const unsigned o7 = hx > 0 ? (okbit(1, cy1) & okbit(0, cy2)) : (okbit(0, cy1) & okbit(1, cy2));
              okmask = (okmask & ~0x80u) | (o7 << 7);
            }
          }
        } else {
#pragma unroll
          for (int k = 0; k < 8; k++)
            if (k < nvalid && sample_ok(k)) okmask |= 1u << k;
        }
        okmask &= ~skip;
      }
#pragma unroll
      for (int k = 0; k < 8; k++) {
        const int a = hx < 0 ? ra[k] : (hx == 0 ? ra[k + 1] : ra[k + 2]);
        const int b = hx < 0 ? rb[k + 2] : (hx == 0 ? rb[k + 1] : rb[k]);
        const int e = sign3(orig[k] - a) + sign3(orig[k] - b);   // -2..2; the packed table holds 0 for e == 0
        const int off = (int)(int8_t)(packed >> (8 * (e + 2)));
        const int r = clip3i(0, maxv, orig[k] + off);
        v[k] = ((okmask >> k) & 1) ? r : orig[k];
      }
    }
  }
  if (pic.dst_flags & HC_DST_RESCALE_LIMITED) {
    // context.cc:2504-2528: bytewise float rescale of limited-range tiles, no FMA contraction
    const float ratio = c == 0 ? 1.1689f : 1.1429f;
#pragma unroll
    for (int k = 0; k < 8; k++) {
      const float full = __fmul_rn(__fsub_rn((float)v[k], (float)(16 << (bit_depth - 8))), ratio);
      const long r = (long)__fadd_rn(full, 0.5f);
      v[k] = r < 0 ? 0 : (r > 255 ? 255 : (int)r);
    }
  }
  Pixel* out = dst + (size_t)(dy0 + oy) * dstride + dx0 + ox0;
  if (nvalid == 8 && ox0 >= 0 && ox0 + 8 <= copy_w && ((dx0 + ox0) & 7) == 0) {
    store8(out, v);
  } else {
#pragma unroll
    for (int k = 0; k < 8; k++)
      if (k < nvalid && ox0 + k >= 0 && ox0 + k < copy_w) out[k] = (Pixel)v[k];
  }
  }
}

// One warp per CTB (it walks the CTB's row groups), 8 CTBs per CTA, the picture descriptor staged in shared memory:
// a CTA moves 8 x 4 KB instead of 2 KB, so the pass is no longer bound by CTA launch rate and descriptor loads
// (the first version was: 147 K CTAs per 8 x 12 MP step, half of them empty chroma CTAs).
__global__ void __launch_bounds__(256, 5) k4_sao_kernel(BatchView bv) {
  __shared__ hc_pic spic;
  {
    const uint32_t* src = reinterpret_cast<const uint32_t*>(bv.pics + blockIdx.y);
    uint32_t* dst = reinterpret_cast<uint32_t*>(&spic);
    for (int i = threadIdx.x; i < (int)(sizeof(hc_pic) / 4); i += blockDim.x) dst[i] = src[i];
  }
  __syncthreads();
  const hc_pic& pic = spic;
  const int c = blockIdx.z;
  if (c > 0 && pic.chroma_format == 0) return;
  if (pic.dst_flags & (HC_DST_SKIP_Y << c)) return;   // component not wanted at the destination
  const unsigned ctb = blockIdx.x * 8u + (threadIdx.x >> 5);
  if (ctb >= (unsigned)pic.ctbs_w * pic.ctbs_h) return;
  const int lane = threadIdx.x & 31;
  if (pic.bit_depth_y == 8 && pic.bit_depth_c == 8) sao_ctb<uint8_t>(bv, pic, c, ctb, lane);
  else sao_ctb<uint16_t>(bv, pic, c, ctb, lane);
}

// max_ctbs = max over pictures of the number of CTBs
void launch_k4(const BatchView& bv, long long max_ctbs, int planes, cudaStream_t stream) {
  if (max_ctbs <= 0 || bv.npics <= 0) return;
  dim3 grid((unsigned)((max_ctbs + 7) / 8), (unsigned)bv.npics, (unsigned)planes);
  k4_sao_kernel<<<grid, 256, 0, stream>>>(bv);
}

}  // namespace hc
