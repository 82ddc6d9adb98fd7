// hevc_params.cc — SPS / PPS / slice-segment-header parsing for the intra-only front-end.
// Written from ITU-T H.265 §7.3.2.2 (SPS), §7.3.2.3 (PPS), §7.3.6.1 (slice header), §E.2.1 (VUI),
// §7.3.4 (scaling lists), §6.5.1/6.5.2 (CTB raster/tile scan, min-TB z-scan).
// Reference behaviour mirrored where it deviates from the text (cited inline).
#include "hevc_params.h"
#include "hevc_scan.h"
#include <algorithm>

namespace hc {

// ---------------------------------------------------------------------------------------------
static void skip_profile_tier_level(BitReader& br, int max_sub_layers_minus1) {
  br.skip(2 + 1 + 5);  // profile_space, tier, profile_idc
  br.skip(32);         // compatibility flags
  br.skip(4);          // progressive, interlaced, non_packed, frame_only
  br.skip(43);         // reserved / profile-specific flags
  br.skip(1);          // inbld / reserved
  br.skip(8);          // general_level_idc
  bool prof[8] = {false}, lvl[8] = {false};
  for (int i = 0; i < max_sub_layers_minus1; i++) {
    prof[i] = br.flag();
    lvl[i] = br.flag();
  }
  if (max_sub_layers_minus1 > 0)
    for (int i = max_sub_layers_minus1; i < 8; i++) br.skip(2);
  for (int i = 0; i < max_sub_layers_minus1; i++) {
    if (prof[i]) br.skip(88);
    if (lvl[i]) br.skip(8);
  }
}

static const uint8_t kDefault4x4[16] = {16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16};
// H.265 Table 7-6 (default 8x8 lists, up-right diagonal order)
static const uint8_t kDefault8x8Intra[64] = {
    16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 17, 16, 17, 16, 17, 18, 17, 18, 18, 17, 18, 21,
    19, 20, 21, 20, 19, 21, 24, 22, 22, 24, 24, 22, 22, 24, 25, 25, 27, 30, 27, 25, 25, 29,
    31, 35, 35, 31, 29, 36, 41, 44, 41, 36, 47, 54, 54, 47, 65, 70, 65, 88, 88, 115};
static const uint8_t kDefault8x8Inter[64] = {
    16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 17, 17, 17, 17, 17, 18, 18, 18, 18, 18, 18, 20,
    20, 20, 20, 20, 20, 20, 24, 24, 24, 24, 24, 24, 24, 24, 25, 25, 25, 25, 25, 25, 25, 28,
    28, 28, 28, 28, 28, 33, 33, 33, 33, 33, 41, 41, 41, 41, 54, 54, 54, 71, 71, 91};

// Expands a diagonal-order list to an n x n ScalingFactor array (8x8 lists are replicated
// 2x2 / 4x4 for 16x16 / 32x32, §7.4.5).
static void fill_factor(uint8_t* dst, const uint8_t* list, int sizeId) {
  const ScanTables& st = scan_tables();
  if (sizeId == 0) {
    const ScanPos* sc = st.order[2][0];
    for (int i = 0; i < 16; i++) dst[sc[i].x + 4 * sc[i].y] = list[i];
    return;
  }
  const ScanPos* sc = st.order[3][0];
  int rep = 1 << (sizeId - 1);
  int n = 8 * rep;
  for (int i = 0; i < 64; i++)
    for (int dy = 0; dy < rep; dy++)
      for (int dx = 0; dx < rep; dx++) dst[(sc[i].x * rep + dx) + n * (sc[i].y * rep + dy)] = list[i];
}

static uint8_t* factor_ptr(ScalingLists& s, int sizeId, int m) {
  switch (sizeId) {
    case 0: return s.s4[m];
    case 1: return s.s8[m];
    case 2: return s.s16[m];
    default: return s.s32[m];
  }
}

static void set_default_scaling(ScalingLists& s) {
  for (int m = 0; m < 6; m++) {
    fill_factor(s.s4[m], kDefault4x4, 0);
    fill_factor(s.s8[m], m < 3 ? kDefault8x8Intra : kDefault8x8Inter, 1);
    fill_factor(s.s16[m], m < 3 ? kDefault8x8Intra : kDefault8x8Inter, 2);
    fill_factor(s.s32[m], m < 3 ? kDefault8x8Intra : kDefault8x8Inter, 3);
  }
}

// scaling_list_data(): 32x32 carries two matrices, kept at matrixId 0 and 3 like the reference
// (sps.cc:960-975, which always steps by 3 for sizeId 3, also for 4:4:4).
static std::string read_scaling_lists(BitReader& br, ScalingLists& out) {
  int dc[4][6];
  for (int sizeId = 0; sizeId < 4; sizeId++) {
    uint8_t lists[6][64];
    for (int m = 0; m < 6; m += (sizeId == 3 ? 3 : 1)) {
      uint8_t* cur = lists[m];
      int dc_coef = 16;
      int n = sizeId == 0 ? 16 : 64;
      if (!br.flag()) {  // scaling_list_pred_mode_flag == 0
        uint32_t delta = br.ue();
        if (delta == 0xFFFFFFFFu) return "scaling list: bad pred_matrix_id_delta";
        if (sizeId == 3) delta *= 3;
        if ((int)delta > m) return "scaling list: pred_matrix_id_delta out of range";
        dc[sizeId][m] = 16;
        if (delta == 0) {
          if (sizeId == 0) memcpy(cur, kDefault4x4, 16);
          else memcpy(cur, m < 3 ? kDefault8x8Intra : kDefault8x8Inter, 64);
        } else {
          int ref = m - (int)delta;
          memcpy(cur, lists[ref], n);
          dc_coef = dc[sizeId][ref];
          dc[sizeId][m] = dc_coef;
        }
      } else {
        int next = 8;
        if (sizeId > 1) {
          int v = br.se();
          if (v < -7 || v > 247) return "scaling list: dc coefficient out of range";
          dc_coef = v + 8;
          next = dc_coef;
          dc[sizeId][m] = dc_coef;
        }
        for (int i = 0; i < n; i++) {
          int d = br.se();
          if (d < -128 || d > 127) return "scaling list: delta out of range";
          next = (next + d + 256) % 256;
          cur[i] = (uint8_t)next;
        }
      }
      uint8_t* f = factor_ptr(out, sizeId, m);
      fill_factor(f, cur, sizeId);
      if (sizeId >= 2) f[0] = (uint8_t)dc_coef;
    }
  }
  // chroma 32x32 matrices derived from the 8x8 factors (sps.cc:1074-1100); unreachable from
  // intra blocks in the reference's dequantiser but kept for completeness.
  for (int m = 0; m < 6; m++)
    if (m != 0 && m != 3) {
      for (int y = 0; y < 8; y++)
        for (int x = 0; x < 8; x++) {
          uint8_t v = out.s8[m][x + 8 * y];
          for (int dy = 0; dy < 4; dy++)
            for (int dx = 0; dx < 4; dx++) out.s32[m][(4 * x + dx) + 32 * (4 * y + dy)] = v;
        }
      out.s32[m][0] = out.s8[m][0];
    }
  if (br.overrun) return "scaling list: truncated";
  return "";
}

// st_ref_pic_set(): only parsed so that the following syntax elements are found at the right bit
// (HEIC stills normally carry none). Derivation per §7.4.8 eq. (7-59)..(7-71).
static std::string read_st_rps(BitReader& br, int idx, int num_sets, const std::vector<ShortTermRps>& sets,
                               bool in_slice_header, ShortTermRps& out) {
  bool inter = false;
  if (idx != 0) inter = br.flag();
  if (inter) {
    int delta_idx = 1;
    if (in_slice_header) {
      uint32_t v = br.ue();
      if (v == 0xFFFFFFFFu || (int)v + 1 > idx) return "st_rps: delta_idx out of range";
      delta_idx = (int)v + 1;
    }
    int ref_idx = idx - delta_idx;
    if (ref_idx < 0 || ref_idx >= (int)sets.size()) return "st_rps: bad reference set";
    const ShortTermRps& ref = sets[ref_idx];
    int sign = br.flag();
    uint32_t absd = br.ue();
    if (absd == 0xFFFFFFFFu || absd > 32767) return "st_rps: abs_delta_rps out of range";
    int deltaRps = (1 - 2 * sign) * ((int)absd + 1);
    int nref = ref.num_delta_pocs();
    bool used[33], use_delta[33];
    for (int j = 0; j <= nref; j++) {
      used[j] = br.flag();
      use_delta[j] = true;
      if (!used[j]) use_delta[j] = br.flag();
    }
    int i = 0;
    for (int j = ref.num_positive - 1; j >= 0; j--) {
      int d = ref.delta_poc_s1[j] + deltaRps;
      if (d < 0 && use_delta[ref.num_negative + j] && i < 16) out.delta_poc_s0[i++] = d;
    }
    if (deltaRps < 0 && use_delta[nref] && i < 16) out.delta_poc_s0[i++] = deltaRps;
    for (int j = 0; j < ref.num_negative; j++) {
      int d = ref.delta_poc_s0[j] + deltaRps;
      if (d < 0 && use_delta[j] && i < 16) out.delta_poc_s0[i++] = d;
    }
    out.num_negative = i;
    i = 0;
    for (int j = ref.num_negative - 1; j >= 0; j--) {
      int d = ref.delta_poc_s0[j] + deltaRps;
      if (d > 0 && use_delta[j] && i < 16) out.delta_poc_s1[i++] = d;
    }
    if (deltaRps > 0 && use_delta[nref] && i < 16) out.delta_poc_s1[i++] = deltaRps;
    for (int j = 0; j < ref.num_positive; j++) {
      int d = ref.delta_poc_s1[j] + deltaRps;
      if (d > 0 && use_delta[ref.num_negative + j] && i < 16) out.delta_poc_s1[i++] = d;
    }
    out.num_positive = i;
  } else {
    uint32_t nn = br.ue(), np = br.ue();
    if (nn > 16 || np > 16 || nn + np > 16) return "st_rps: too many pictures";
    out.num_negative = (int)nn;
    out.num_positive = (int)np;
    int poc = 0;
    for (uint32_t i = 0; i < nn; i++) {
      poc -= (int)br.ue() + 1;
      br.flag();
      out.delta_poc_s0[i] = poc;
    }
    poc = 0;
    for (uint32_t i = 0; i < np; i++) {
      poc += (int)br.ue() + 1;
      br.flag();
      out.delta_poc_s1[i] = poc;
    }
  }
  (void)num_sets;
  if (br.overrun) return "st_rps: truncated";
  return "";
}

static void skip_sub_layer_hrd(BitReader& br, int cpb_cnt, bool sub_pic) {
  for (int i = 0; i < cpb_cnt; i++) {
    br.ue(); br.ue();
    if (sub_pic) { br.ue(); br.ue(); }
    br.flag();
  }
}

static void skip_hrd(BitReader& br, bool common, int max_sub_layers_minus1) {
  bool nal = false, vcl = false, sub_pic = false;
  if (common) {
    nal = br.flag();
    vcl = br.flag();
    if (nal || vcl) {
      sub_pic = br.flag();
      if (sub_pic) br.skip(8 + 5 + 1 + 5);
      br.skip(4 + 4);
      if (sub_pic) br.skip(4);
      br.skip(5 + 5 + 5);
    }
  }
  for (int i = 0; i <= max_sub_layers_minus1; i++) {
    bool fixed_general = br.flag();
    bool fixed_cvs = true;
    if (!fixed_general) fixed_cvs = br.flag();
    bool low_delay = false;
    if (fixed_cvs) br.ue();
    else low_delay = br.flag();
    int cpb_cnt = 1;
    if (!low_delay) cpb_cnt = (int)br.ue() + 1;
    if (cpb_cnt > 32) { br.overrun = true; return; }
    if (nal) skip_sub_layer_hrd(br, cpb_cnt, sub_pic);
    if (vcl) skip_sub_layer_hrd(br, cpb_cnt, sub_pic);
  }
}

static void read_vui(BitReader& br, Sps& sps) {
  if (br.flag()) {                     // aspect_ratio_info_present_flag
    if (br.u(8) == 255) br.skip(32);
  }
  if (br.flag()) br.flag();            // overscan
  if (br.flag()) {                     // video_signal_type_present_flag
    br.u(3);
    sps.video_full_range = br.flag();
    if (br.flag()) {
      // value sanitising as in the reference (vui.cc:296-312)
      int cp = br.u(8), tc = br.u(8), mc = br.u(8);
      if (cp == 0 || cp == 3 || cp >= 11) cp = 2;
      if (tc == 0 || tc == 3 || tc >= 18) tc = 2;
      if (mc >= 11) mc = 2;
      sps.colour_primaries = cp;
      sps.transfer_characteristics = tc;
      sps.matrix_coeffs = mc;
    }
  }
  if (br.flag()) { br.ue(); br.ue(); } // chroma_loc_info
  br.skip(3);                          // neutral_chroma, field_seq, frame_field_info
  if (br.flag()) { br.ue(); br.ue(); br.ue(); br.ue(); }  // default display window
  if (br.flag()) {                     // vui_timing_info_present_flag
    br.skip(32); br.skip(32);
    if (br.flag()) br.ue();
    if (br.flag()) skip_hrd(br, true, sps.max_sub_layers - 1);
  }
  if (br.flag()) {                     // bitstream_restriction_flag
    br.skip(3);
    br.ue(); br.ue(); br.ue(); br.ue(); br.ue();
  }
}

std::string parse_sps(const uint8_t* rbsp, size_t n, Sps& sps) {
  sps = Sps();
  BitReader br(rbsp, n);
  br.skip(4);  // sps_video_parameter_set_id
  int max_sub_layers_minus1 = br.u(3);
  if (max_sub_layers_minus1 > 6) return "SPS: sps_max_sub_layers_minus1 out of range";
  sps.max_sub_layers = max_sub_layers_minus1 + 1;
  br.skip(1);  // temporal_id_nesting
  skip_profile_tier_level(br, max_sub_layers_minus1);
  uint32_t id = br.ue();
  if (id > 15) return "SPS: id out of range";
  sps.sps_id = (int)id;
  uint32_t cf = br.ue();
  if (cf > 3) return "SPS: chroma_format_idc out of range";
  sps.chroma_format_idc = (int)cf;
  if (cf == 3) sps.separate_colour_plane = br.flag();
  sps.ChromaArrayType = sps.separate_colour_plane ? 0 : sps.chroma_format_idc;
  sps.SubWidthC = (cf == 1 || cf == 2) ? 2 : 1;
  sps.SubHeightC = (cf == 1) ? 2 : 1;
  uint32_t w = br.ue(), h = br.ue();
  if (w == 0 || h == 0 || w > 65535 || h > 65535) return "SPS: picture size out of range";
  // the reference's security limit (heif_limits.h:37-38, context.cc:547-560 check_resolution): 32768 x 32768 samples in total
  if ((uint64_t)w * h > (uint64_t)32768 * 32768) return "SPS: picture size exceeds the maximum image size";
  sps.width = (int)w;
  sps.height = (int)h;
  if (br.flag()) {
    sps.conf_left = (int)br.ue();
    sps.conf_right = (int)br.ue();
    sps.conf_top = (int)br.ue();
    sps.conf_bottom = (int)br.ue();
  }
  uint32_t bdy = br.ue(), bdc = br.ue();
  if (bdy > 8 || bdc > 8) return "SPS: bit depth out of range";
  sps.bit_depth_y = 8 + (int)bdy;
  sps.bit_depth_c = 8 + (int)bdc;
  sps.qp_bd_offset_y = 6 * (int)bdy;
  sps.qp_bd_offset_c = 6 * (int)bdc;
  uint32_t lp = br.ue();
  if (lp > 12) return "SPS: log2_max_pic_order_cnt_lsb out of range";
  sps.log2_max_poc_lsb = (int)lp + 4;
  bool sub_layer_ordering = br.flag();
  for (int i = sub_layer_ordering ? 0 : max_sub_layers_minus1; i <= max_sub_layers_minus1; i++) {
    br.ue(); br.ue(); br.ue();
  }
  uint32_t l2mincb = br.ue(), l2diffcb = br.ue(), l2mintb = br.ue(), l2difftb = br.ue();
  if (l2mincb > 3 || l2diffcb > 3 || l2mintb > 3 || l2difftb > 3) return "SPS: block size out of range";
  sps.log2_min_cb = (int)l2mincb + 3;
  sps.log2_ctb = sps.log2_min_cb + (int)l2diffcb;
  sps.log2_min_tb = (int)l2mintb + 2;
  sps.log2_max_tb = sps.log2_min_tb + (int)l2difftb;
  if (sps.log2_ctb > 6 || sps.log2_ctb < 4) return "SPS: CTB size unsupported";
  if (sps.log2_max_tb > 5 || sps.log2_max_tb > sps.log2_ctb) return "SPS: max TB size out of range";
  if (sps.log2_min_tb >= sps.log2_min_cb) return "SPS: min TB size must be below min CB size";
  sps.max_th_depth_inter = (int)br.ue();
  sps.max_th_depth_intra = (int)br.ue();
  if (sps.max_th_depth_intra > sps.log2_ctb - sps.log2_min_tb) return "SPS: transform hierarchy depth out of range";
  sps.scaling_list_enabled = br.flag();
  if (sps.scaling_list_enabled) {
    if (br.flag()) {
      std::string e = read_scaling_lists(br, sps.scaling);
      if (!e.empty()) return "SPS: " + e;
    } else {
      set_default_scaling(sps.scaling);
    }
  }
  sps.amp_enabled = br.flag();
  sps.sao_enabled = br.flag();
  sps.pcm_enabled = br.flag();
  if (sps.pcm_enabled) {
    sps.pcm_bit_depth_y = br.u(4) + 1;
    sps.pcm_bit_depth_c = br.u(4) + 1;
    sps.log2_min_pcm_cb = (int)br.ue() + 3;
    sps.log2_max_pcm_cb = sps.log2_min_pcm_cb + (int)br.ue();
    sps.pcm_loop_filter_disabled = br.flag();
    if (sps.pcm_bit_depth_y > sps.bit_depth_y || sps.pcm_bit_depth_c > sps.bit_depth_c)
      return "SPS: PCM bit depth exceeds sample bit depth";
  }
  uint32_t nst = br.ue();
  if (nst > 64) return "SPS: num_short_term_ref_pic_sets out of range";
  sps.st_rps.resize(nst);
  for (uint32_t i = 0; i < nst; i++) {
    std::string e = read_st_rps(br, (int)i, (int)nst, sps.st_rps, false, sps.st_rps[i]);
    if (!e.empty()) return "SPS: " + e;
  }
  sps.long_term_ref_pics_present = br.flag();
  if (sps.long_term_ref_pics_present) {
    uint32_t nlt = br.ue();
    if (nlt > 32) return "SPS: num_long_term_ref_pics_sps out of range";
    sps.num_long_term_ref_pics_sps = (int)nlt;
    for (uint32_t i = 0; i < nlt; i++) { br.skip(sps.log2_max_poc_lsb); br.flag(); }
  }
  sps.temporal_mvp_enabled = br.flag();
  sps.strong_intra_smoothing = br.flag();
  sps.vui_present = br.flag();
  if (sps.vui_present) read_vui(br, sps);
  if (br.flag()) {  // sps_extension_present_flag
    bool range_ext = br.flag();
    br.skip(7);     // multilayer, 3d, scc, 4 bits
    if (range_ext) {
      sps.transform_skip_rotation_enabled = br.flag();
      sps.transform_skip_context_enabled = br.flag();
      sps.implicit_rdpcm_enabled = br.flag();
      sps.explicit_rdpcm_enabled = br.flag();
      sps.extended_precision_processing = br.flag();
      sps.intra_smoothing_disabled = br.flag();
      sps.high_precision_offsets_enabled = br.flag();
      sps.persistent_rice_adaptation_enabled = br.flag();
      sps.cabac_bypass_alignment_enabled = br.flag();
    }
  }
  if (br.overrun) return "SPS: truncated";

  int min_cb = 1 << sps.log2_min_cb;
  if (sps.width % min_cb || sps.height % min_cb) return "SPS: picture size not a multiple of the min CB size";
  int ctb = 1 << sps.log2_ctb;
  sps.ctbs_w = (sps.width + ctb - 1) / ctb;
  sps.ctbs_h = (sps.height + ctb - 1) / ctb;
  sps.pic_size_in_ctbs = sps.ctbs_w * sps.ctbs_h;
  sps.min_cb_w = sps.width >> sps.log2_min_cb;
  sps.min_cb_h = sps.height >> sps.log2_min_cb;
  sps.tbs_w = sps.ctbs_w << (sps.log2_ctb - sps.log2_min_tb);
  sps.tbs_h = sps.ctbs_h << (sps.log2_ctb - sps.log2_min_tb);
  if (sps.conf_left * sps.SubWidthC + sps.conf_right * sps.SubWidthC >= sps.width ||
      sps.conf_top * sps.SubHeightC + sps.conf_bottom * sps.SubHeightC >= sps.height)
    return "SPS: conformance window larger than the picture";
  // A monochrome stream still codes bit_depth_chroma; nothing decodes with it, but the pixel type of a picture is chosen
  // from both depths further down, so it follows luma here (the reference ignores it for chroma_format_idc 0).
  if (sps.chroma_format_idc == 0) { sps.bit_depth_c = sps.bit_depth_y; sps.qp_bd_offset_c = sps.qp_bd_offset_y; }
  sps.valid = true;
  return "";
}

// ---------------------------------------------------------------------------------------------
static void derive_pps(const Sps& sps, Pps& pps) {
  // tile boundaries
  int W = sps.ctbs_w, H = sps.ctbs_h;
  if (pps.uniform_spacing || !pps.tiles_enabled) {
    pps.col_bd.assign(pps.num_tile_cols + 1, 0);
    pps.row_bd.assign(pps.num_tile_rows + 1, 0);
    for (int i = 0; i < pps.num_tile_cols; i++)
      pps.col_bd[i + 1] = ((i + 1) * W) / pps.num_tile_cols;
    for (int i = 0; i < pps.num_tile_rows; i++)
      pps.row_bd[i + 1] = ((i + 1) * H) / pps.num_tile_rows;
  }
  int n = W * H;
  pps.ctb_addr_rs_to_ts.assign(n, 0);
  pps.ctb_addr_ts_to_rs.assign(n, 0);
  pps.tile_id_rs.assign(n, 0);
  pps.tile_start_ctb.assign(n, 0);
  // §6.5.1
  for (int rs = 0; rs < n; rs++) {
    int tbX = rs % W, tbY = rs / W;
    int tileX = 0, tileY = 0;
    for (int i = 0; i < pps.num_tile_cols; i++)
      if (tbX >= pps.col_bd[i]) tileX = i;
    for (int j = 0; j < pps.num_tile_rows; j++)
      if (tbY >= pps.row_bd[j]) tileY = j;
    int ts = 0;
    for (int i = 0; i < tileX; i++)
      ts += (pps.row_bd[tileY + 1] - pps.row_bd[tileY]) * (pps.col_bd[i + 1] - pps.col_bd[i]);
    for (int j = 0; j < tileY; j++) ts += W * (pps.row_bd[j + 1] - pps.row_bd[j]);
    ts += (tbY - pps.row_bd[tileY]) * (pps.col_bd[tileX + 1] - pps.col_bd[tileX]) + tbX - pps.col_bd[tileX];
    pps.ctb_addr_rs_to_ts[rs] = ts;
    pps.ctb_addr_ts_to_rs[ts] = rs;
    pps.tile_id_rs[rs] = tileY * pps.num_tile_cols + tileX;
    if (tbX == pps.col_bd[tileX] && tbY == pps.row_bd[tileY]) pps.tile_start_ctb[rs] = 1;
  }
  // §6.5.2 MinTbAddrZs: derive_min_tb_addr_zs(), on demand (only the host slice-data parser needs it)
  pps.min_tb_addr_zs.clear();
  pps.log2_min_cu_qp_delta_size = sps.log2_ctb - pps.diff_cu_qp_delta_depth;
  pps.log2_min_cu_chroma_qp_offset_size = sps.log2_ctb - pps.diff_cu_chroma_qp_offset_depth;
}

// §6.5.2 MinTbAddrZs. One entry per minimum transform block of the picture (16 K for a 512 x 512 tile): this table was
// 95 % of the host work per device-parsed item while it was built with every PPS, so it is built when a picture that the
// HOST parser decodes starts (HevcIntraParser::Impl::start_picture) — K0 computes z-scan addresses itself.
void derive_min_tb_addr_zs(const Sps& sps, Pps& pps) {
  const int W = sps.ctbs_w, shift = sps.log2_ctb - sps.log2_min_tb, mask = (1 << shift) - 1;
  int morton[16][16];   // shift <= 4 (CTB 64, minimum TB 4)
  for (int y = 0; y <= mask; y++)
    for (int x = 0; x <= mask; x++) {
      int v = 0;
      for (int i = 0; i < shift; i++) {
        const int m = 1 << i;
        v += (m & x ? m * m : 0) + (m & y ? 2 * m * m : 0);
      }
      morton[y][x] = v;
    }
  pps.min_tb_addr_zs.assign((size_t)sps.tbs_w * sps.tbs_h, 0);
  for (int y = 0; y < sps.tbs_h; y++)
    for (int x = 0; x < sps.tbs_w; x++)
      pps.min_tb_addr_zs[x + (size_t)y * sps.tbs_w] = (pps.ctb_addr_rs_to_ts[W * (y >> shift) + (x >> shift)] << (shift * 2)) + morton[y & mask][x & mask];
}

std::string parse_pps(const uint8_t* rbsp, size_t n, const Sps* sps_table, Pps& pps) {
  pps = Pps();
  BitReader br(rbsp, n);
  uint32_t id = br.ue();
  if (id > 63) return "PPS: id out of range";
  pps.pps_id = (int)id;
  uint32_t sid = br.ue();
  if (sid > 15 || !sps_table[sid].valid) return "PPS: refers to a missing SPS";
  pps.sps_id = (int)sid;
  const Sps& sps = sps_table[sid];
  pps.dependent_slice_segments_enabled = br.flag();
  pps.output_flag_present = br.flag();
  pps.num_extra_slice_header_bits = br.u(3);
  pps.sign_data_hiding = br.flag();
  pps.cabac_init_present = br.flag();
  br.ue(); br.ue();  // num_ref_idx defaults
  int iq = br.se();
  pps.init_qp = 26 + iq;
  if (pps.init_qp < -sps.qp_bd_offset_y || pps.init_qp > 51) return "PPS: init_qp out of range";
  pps.constrained_intra_pred = br.flag();
  pps.transform_skip_enabled = br.flag();
  pps.cu_qp_delta_enabled = br.flag();
  if (pps.cu_qp_delta_enabled) {
    uint32_t d = br.ue();
    if ((int)d > sps.log2_ctb - sps.log2_min_cb) return "PPS: diff_cu_qp_delta_depth out of range";
    pps.diff_cu_qp_delta_depth = (int)d;
  }
  pps.cb_qp_offset = br.se();
  pps.cr_qp_offset = br.se();
  if (pps.cb_qp_offset < -12 || pps.cb_qp_offset > 12 || pps.cr_qp_offset < -12 || pps.cr_qp_offset > 12)
    return "PPS: chroma qp offset out of range";
  pps.slice_chroma_qp_offsets_present = br.flag();
  pps.weighted_pred = br.flag();
  pps.weighted_bipred = br.flag();
  pps.transquant_bypass_enabled = br.flag();
  pps.tiles_enabled = br.flag();
  pps.entropy_coding_sync_enabled = br.flag();
  if (pps.tiles_enabled) {
    uint32_t nc = br.ue(), nr = br.ue();
    if ((int)nc >= sps.ctbs_w || (int)nr >= sps.ctbs_h || nc > 19 || nr > 21) return "PPS: tile count out of range";
    pps.num_tile_cols = (int)nc + 1;
    pps.num_tile_rows = (int)nr + 1;
    pps.uniform_spacing = br.flag();
    if (!pps.uniform_spacing) {
      pps.col_bd.assign(pps.num_tile_cols + 1, 0);
      pps.row_bd.assign(pps.num_tile_rows + 1, 0);
      for (int i = 0; i < pps.num_tile_cols - 1; i++) {
        uint32_t wv = br.ue();
        if (wv > 65535) return "PPS: tile column width out of range";
        pps.col_bd[i + 1] = pps.col_bd[i] + (int)wv + 1;
      }
      for (int i = 0; i < pps.num_tile_rows - 1; i++) {
        uint32_t hv = br.ue();
        if (hv > 65535) return "PPS: tile row height out of range";
        pps.row_bd[i + 1] = pps.row_bd[i] + (int)hv + 1;
      }
      pps.col_bd[pps.num_tile_cols] = sps.ctbs_w;
      pps.row_bd[pps.num_tile_rows] = sps.ctbs_h;
      if (pps.col_bd[pps.num_tile_cols - 1] >= sps.ctbs_w || pps.row_bd[pps.num_tile_rows - 1] >= sps.ctbs_h)
        return "PPS: tile sizes exceed the picture";
    }
    pps.loop_filter_across_tiles = br.flag();
  }
  pps.loop_filter_across_slices = br.flag();
  pps.deblocking_control_present = br.flag();
  if (pps.deblocking_control_present) {
    pps.deblocking_override_enabled = br.flag();
    pps.deblocking_disabled = br.flag();
    if (!pps.deblocking_disabled) {
      int b = br.se(), t = br.se();
      if (b < -6 || b > 6 || t < -6 || t > 6) return "PPS: deblocking offsets out of range";
      pps.beta_offset = 2 * b;
      pps.tc_offset = 2 * t;
    }
  }
  pps.scaling_list_data_present = br.flag();
  if (pps.scaling_list_data_present) {
    if (!sps.scaling_list_enabled) return "PPS: scaling list data without SPS scaling lists";
    std::string e = read_scaling_lists(br, pps.scaling);
    if (!e.empty()) return "PPS: " + e;
  } else if (sps.scaling_list_enabled) {
    pps.scaling = sps.scaling;
  }
  pps.lists_modification_present = br.flag();
  pps.log2_parallel_merge_level = (int)br.ue() + 2;
  pps.slice_header_extension_present = br.flag();
  if (br.flag()) {  // pps_extension_present_flag
    bool range_ext = br.flag();
    br.skip(7);
    if (range_ext) {
      if (pps.transform_skip_enabled) {
        uint32_t v = br.ue();
        if (v > 3) return "PPS: log2_max_transform_skip_block_size out of range";
        pps.log2_max_transform_skip_size = (int)v + 2;
      }
      pps.cross_component_prediction_enabled = br.flag();
      pps.chroma_qp_offset_list_enabled = br.flag();
      if (pps.chroma_qp_offset_list_enabled) {
        uint32_t d = br.ue();
        if ((int)d > sps.log2_ctb - sps.log2_min_cb) return "PPS: diff_cu_chroma_qp_offset_depth out of range";
        pps.diff_cu_chroma_qp_offset_depth = (int)d;
        uint32_t len = br.ue();
        if (len > 5) return "PPS: chroma_qp_offset_list_len out of range";
        pps.chroma_qp_offset_list_len = (int)len + 1;
        for (int i = 0; i < pps.chroma_qp_offset_list_len; i++) {
          pps.cb_qp_offset_list[i] = br.se();
          pps.cr_qp_offset_list[i] = br.se();
        }
      }
      pps.log2_sao_offset_scale_luma = (int)br.ue();
      pps.log2_sao_offset_scale_chroma = (int)br.ue();
      if (pps.log2_sao_offset_scale_luma > std::max(0, sps.bit_depth_y - 10) ||
          pps.log2_sao_offset_scale_chroma > std::max(0, sps.bit_depth_c - 10))
        return "PPS: sao offset scale out of range";
    }
  }
  if (br.overrun) return "PPS: truncated";
  derive_pps(sps, pps);
  pps.valid = true;
  return "";
}

// ---------------------------------------------------------------------------------------------
std::string parse_slice_header(const uint8_t* rbsp, size_t n, int nal_type, const Sps* sps_table,
                               const Pps* pps_table, const SliceHeader* prev_independent,
                               const std::vector<uint32_t>& skipped, SliceHeader& sh) {
  BitReader br(rbsp, n);
  br.skip(16);  // NAL header
  sh = SliceHeader();
  sh.first_slice_segment_in_pic = br.flag();
  if (nal_type >= NAL_BLA_W_LP && nal_type <= NAL_RSV_IRAP_VCL23) br.flag();  // no_output_of_prior_pics
  uint32_t pid = br.ue();
  if (pid > 63 || !pps_table[pid].valid) return "slice: refers to a missing PPS";
  sh.pps_id = (int)pid;
  const Pps& pps = pps_table[pid];
  const Sps& sps = sps_table[pps.sps_id];
  if (!sps.valid) return "slice: refers to a missing SPS";
  if (!sh.first_slice_segment_in_pic) {
    if (pps.dependent_slice_segments_enabled) sh.dependent = br.flag();
    int bits = ceil_log2((uint32_t)sps.pic_size_in_ctbs);
    int addr = (int)br.u(bits);
    if (addr >= sps.pic_size_in_ctbs) return "slice: segment address out of range";
    if (sh.dependent) {
      if (addr == 0 || !prev_independent) return "slice: dependent slice segment without a preceding slice";
      bool dep = true;
      sh = *prev_independent;
      sh.first_slice_segment_in_pic = false;
      sh.dependent = dep;
      sh.entry_points.clear();
    }
    sh.segment_address = addr;
  }
  if (!sh.dependent) {
    sh.slice_addr_rs = sh.segment_address;
    br.skip(pps.num_extra_slice_header_bits);
    uint32_t st = br.ue();
    if (st > 2) return "slice: slice_type out of range";
    sh.slice_type = (int)st;
    if (sh.slice_type != 2) return "slice: only intra (I) slices are supported by this decoder";
    if (pps.output_flag_present) br.flag();
    if (sps.separate_colour_plane) br.skip(2);
    if (nal_type != NAL_IDR_W_RADL && nal_type != NAL_IDR_N_LP) {
      br.skip(sps.log2_max_poc_lsb);
      bool sps_rps = br.flag();
      int NumDeltaPocs = 0;
      if (!sps_rps) {
        ShortTermRps tmp;
        std::string e = read_st_rps(br, (int)sps.st_rps.size(), (int)sps.st_rps.size(), sps.st_rps, true, tmp);
        if (!e.empty()) return "slice: " + e;
        NumDeltaPocs = tmp.num_delta_pocs();
      } else {
        int bits = ceil_log2((uint32_t)sps.st_rps.size());
        int idx = bits > 0 ? (int)br.u(bits) : 0;
        if (idx >= (int)sps.st_rps.size()) return "slice: short_term_ref_pic_set_idx out of range";
        NumDeltaPocs = sps.st_rps[idx].num_delta_pocs();
      }
      (void)NumDeltaPocs;
      if (sps.long_term_ref_pics_present) {
        uint32_t num_lt_sps = 0;
        if (sps.num_long_term_ref_pics_sps > 0) num_lt_sps = br.ue();
        uint32_t num_lt_pics = br.ue();
        if (num_lt_sps > 32 || num_lt_pics > 32) return "slice: too many long-term pictures";
        for (uint32_t i = 0; i < num_lt_sps + num_lt_pics; i++) {
          if (i < num_lt_sps) {
            int bits = ceil_log2((uint32_t)sps.num_long_term_ref_pics_sps);
            if (bits > 0) br.skip(bits);
          } else {
            br.skip(sps.log2_max_poc_lsb);
            br.flag();
          }
          if (br.flag()) br.ue();
        }
      }
      if (sps.temporal_mvp_enabled) br.flag();
    }
    if (sps.sao_enabled) {
      sh.sao_luma = br.flag();
      if (sps.ChromaArrayType != 0) sh.sao_chroma = br.flag();
    }
    sh.slice_qp_delta = br.se();
    if (pps.slice_chroma_qp_offsets_present) {
      sh.cb_qp_offset = br.se();
      sh.cr_qp_offset = br.se();
      if (sh.cb_qp_offset < -12 || sh.cb_qp_offset > 12 || sh.cr_qp_offset < -12 || sh.cr_qp_offset > 12)
        return "slice: chroma qp offset out of range";
    }
    if (pps.chroma_qp_offset_list_enabled) sh.cu_chroma_qp_offset_enabled = br.flag();
    bool override_flag = false;
    if (pps.deblocking_override_enabled) override_flag = br.flag();
    sh.deblocking_disabled = pps.deblocking_disabled;
    sh.beta_offset = pps.beta_offset;
    sh.tc_offset = pps.tc_offset;
    if (override_flag) {
      sh.deblocking_disabled = br.flag();
      if (!sh.deblocking_disabled) {
        int b = br.se(), t = br.se();
        if (b < -6 || b > 6 || t < -6 || t > 6) return "slice: deblocking offsets out of range";
        sh.beta_offset = 2 * b;
        sh.tc_offset = 2 * t;
      }
    }
    sh.loop_filter_across_slices = pps.loop_filter_across_slices;
    if (pps.loop_filter_across_slices && (sh.sao_luma || sh.sao_chroma || !sh.deblocking_disabled))
      sh.loop_filter_across_slices = br.flag();
    sh.slice_qp_y = pps.init_qp + sh.slice_qp_delta;
    if (sh.slice_qp_y < -sps.qp_bd_offset_y || sh.slice_qp_y > 51) return "slice: SliceQpY out of range";
  }
  if (pps.tiles_enabled || pps.entropy_coding_sync_enabled) {
    uint32_t num = br.ue();
    if (num > (uint32_t)sps.pic_size_in_ctbs) return "slice: num_entry_point_offsets out of range";
    if (num > 0) {
      uint32_t len = br.ue() + 1;
      if (len > 32) return "slice: offset_len_minus1 out of range";
      sh.entry_points.resize(num);
      uint32_t acc = 0;
      for (uint32_t i = 0; i < num; i++) {
        acc += br.u((int)len) + 1;
        sh.entry_points[i] = acc;  // escaped-domain offset of substream i+1 from the slice data start
      }
    }
  }
  if (pps.slice_header_extension_present) {
    uint32_t len = br.ue();
    if (len > 256) return "slice: header extension too long";
    br.skip(8 * (int)len);
  }
  // byte_alignment(): a one bit followed by zero bits
  br.flag();
  while (!br.byte_aligned()) br.flag();
  if (br.overrun) return "slice: header truncated";
  sh.data_byte_offset = br.pos >> 3;

  // Convert entry points from escaped to unescaped byte offsets (reference: decctx.cc:674-680).
  // `skipped` holds positions in the escaped payload; the slice data starts at unescaped offset
  // data_byte_offset, i.e. escaped offset data_byte_offset + (#skipped bytes before it).
  if (!sh.entry_points.empty()) {
    size_t hdr_skipped = 0;
    // escaped position of the first slice-data byte
    size_t esc_start = sh.data_byte_offset;
    for (size_t k = 0; k < skipped.size(); k++) {
      if (skipped[k] < esc_start + 1) { esc_start++; hdr_skipped++; }
      else break;
    }
    (void)hdr_skipped;
    for (size_t i = 0; i < sh.entry_points.size(); i++) {
      size_t esc_pos = esc_start + sh.entry_points[i];  // escaped position of the substream start
      size_t removed = 0;
      for (size_t k = 0; k < skipped.size(); k++)
        if (skipped[k] < esc_pos) removed++;
      size_t unesc = esc_pos - removed;
      sh.entry_points[i] = (uint32_t)(unesc - sh.data_byte_offset);
    }
  }
  return "";
}

}  // namespace hc
