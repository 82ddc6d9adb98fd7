// hevc_params.h — VPS/SPS/PPS/slice-segment-header state for the intra-only HEVC front-end.
// New code written from ITU-T H.265 §7.3.2 / §7.3.6 / §7.4; the reference's equivalents are
// third-party/libde265/libde265/sps.cc:198-560, pps.cc:270-780, slice.cc:356-900, vui.cc:170-420.
#pragma once
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>
#include <memory>
#include "hevc_bits.h"

namespace hc {

enum NalType {
  NAL_TRAIL_N = 0, NAL_BLA_W_LP = 16, NAL_BLA_W_RADL = 17, NAL_BLA_N_LP = 18,
  NAL_IDR_W_RADL = 19, NAL_IDR_N_LP = 20, NAL_CRA = 21, NAL_RSV_IRAP_VCL23 = 23,
  NAL_VPS = 32, NAL_SPS = 33, NAL_PPS = 34, NAL_AUD = 35, NAL_EOS = 36, NAL_EOB = 37,
  NAL_FD = 38, NAL_SEI_PREFIX = 39, NAL_SEI_SUFFIX = 40
};

// Scaling factors expanded to block size, [matrixId][y*n+x] (libde265 keeps the second 32x32
// matrix at matrixId 3; only matrixId 0 is reachable from intra blocks: transform.cc:512-522).
struct ScalingLists {
  uint8_t s4[6][16];
  uint8_t s8[6][64];
  uint8_t s16[6][256];
  uint8_t s32[6][1024];
};

struct ShortTermRps {
  int num_negative = 0, num_positive = 0;
  int delta_poc_s0[16], delta_poc_s1[16];
  int num_delta_pocs() const { return num_negative + num_positive; }
};

struct Sps {
  bool valid = false;
  int sps_id = 0;
  int max_sub_layers = 1;
  int chroma_format_idc = 1;
  bool separate_colour_plane = false;
  int ChromaArrayType = 1;
  int width = 0, height = 0;           // pic_width/height_in_luma_samples
  int conf_left = 0, conf_right = 0, conf_top = 0, conf_bottom = 0;  // in chroma units
  int bit_depth_y = 8, bit_depth_c = 8;
  int log2_max_poc_lsb = 4;
  int log2_min_cb = 3, log2_ctb = 4, log2_min_tb = 2, log2_max_tb = 5;
  int max_th_depth_inter = 0, max_th_depth_intra = 0;
  bool scaling_list_enabled = false;
  ScalingLists scaling;                 // valid if scaling_list_enabled
  bool amp_enabled = false, sao_enabled = false;
  bool pcm_enabled = false;
  int pcm_bit_depth_y = 8, pcm_bit_depth_c = 8, log2_min_pcm_cb = 3, log2_max_pcm_cb = 3;
  bool pcm_loop_filter_disabled = false;
  std::vector<ShortTermRps> st_rps;
  bool long_term_ref_pics_present = false;
  int num_long_term_ref_pics_sps = 0;
  bool temporal_mvp_enabled = false;
  bool strong_intra_smoothing = false;
  // VUI (defaults per vui.cc:93-97)
  bool vui_present = false;
  int video_full_range = 0, colour_primaries = 2, transfer_characteristics = 2, matrix_coeffs = 2;
  // range extension
  bool transform_skip_rotation_enabled = false, transform_skip_context_enabled = false;
  bool implicit_rdpcm_enabled = false, explicit_rdpcm_enabled = false;
  bool extended_precision_processing = false, intra_smoothing_disabled = false;
  bool high_precision_offsets_enabled = false, persistent_rice_adaptation_enabled = false;
  bool cabac_bypass_alignment_enabled = false;
  // derived
  int SubWidthC = 2, SubHeightC = 2;
  int ctbs_w = 0, ctbs_h = 0, pic_size_in_ctbs = 0;
  int min_cb_w = 0, min_cb_h = 0;       // picture size in min coding blocks
  int tbs_w = 0, tbs_h = 0;             // picture size in min transform blocks
  int qp_bd_offset_y = 0, qp_bd_offset_c = 0;
};

struct Pps {
  bool valid = false;
  int pps_id = 0, sps_id = 0;
  bool dependent_slice_segments_enabled = false, output_flag_present = false;
  int num_extra_slice_header_bits = 0;
  bool sign_data_hiding = false, cabac_init_present = false;
  int init_qp = 26;
  bool constrained_intra_pred = false, transform_skip_enabled = false;
  bool cu_qp_delta_enabled = false;
  int diff_cu_qp_delta_depth = 0;
  int cb_qp_offset = 0, cr_qp_offset = 0;
  bool slice_chroma_qp_offsets_present = false;
  bool weighted_pred = false, weighted_bipred = false;
  bool transquant_bypass_enabled = false;
  bool tiles_enabled = false, entropy_coding_sync_enabled = false;
  int num_tile_cols = 1, num_tile_rows = 1;
  bool uniform_spacing = true;
  std::vector<int> col_bd, row_bd;      // tile boundaries in CTBs (size cols+1 / rows+1)
  bool loop_filter_across_tiles = true;
  bool loop_filter_across_slices = false;
  bool deblocking_control_present = false, deblocking_override_enabled = false;
  bool deblocking_disabled = false;
  int beta_offset = 0, tc_offset = 0;   // already *2
  bool scaling_list_data_present = false;
  ScalingLists scaling;                 // effective lists (own or copied from SPS)
  bool lists_modification_present = false;
  int log2_parallel_merge_level = 2;
  bool slice_header_extension_present = false;
  // range extension
  int log2_max_transform_skip_size = 2;
  bool cross_component_prediction_enabled = false;
  bool chroma_qp_offset_list_enabled = false;
  int diff_cu_chroma_qp_offset_depth = 0, chroma_qp_offset_list_len = 0;
  int cb_qp_offset_list[6] = {0}, cr_qp_offset_list[6] = {0};
  int log2_sao_offset_scale_luma = 0, log2_sao_offset_scale_chroma = 0;
  // derived (need the SPS)
  int log2_min_cu_qp_delta_size = 0, log2_min_cu_chroma_qp_offset_size = 0;
  std::vector<int> ctb_addr_rs_to_ts, ctb_addr_ts_to_rs, tile_id_rs;   // tile_id indexed by RS
  std::vector<int> min_tb_addr_zs;      // [x + y*tbs_w] in min-TB units
  std::vector<uint8_t> tile_start_ctb;  // indexed RS: CTB is the first CTB of a tile
};

struct SliceHeader {
  bool first_slice_segment_in_pic = true;
  bool dependent = false;
  int pps_id = 0;
  int segment_address = 0;     // CTB address (RS) of the first CTB of this slice segment
  int slice_addr_rs = 0;       // SliceAddrRs: address of the first CTB of the (independent) slice
  int slice_type = 2;
  bool sao_luma = false, sao_chroma = false;
  int slice_qp_delta = 0, cb_qp_offset = 0, cr_qp_offset = 0;
  bool cu_chroma_qp_offset_enabled = false;
  bool deblocking_disabled = false;
  int beta_offset = 0, tc_offset = 0;   // already *2
  bool loop_filter_across_slices = false;
  std::vector<uint32_t> entry_points;   // byte offsets (unescaped) of substreams 1.. relative to slice data start
  int slice_qp_y = 26;
  size_t data_byte_offset = 0;          // byte offset of slice_segment_data in the unescaped NAL payload
};

// Parses parameter sets. All functions return an empty string on success, else an error text.
std::string parse_sps(const uint8_t* rbsp, size_t n, Sps& sps);
// fills pps.min_tb_addr_zs (left empty by parse_pps)
void derive_min_tb_addr_zs(const Sps& sps, Pps& pps);
std::string parse_pps(const uint8_t* rbsp, size_t n, const Sps* sps_table /*[16]*/, Pps& pps);
// `skipped` are the escaped-input positions of removed emulation-prevention bytes (relative to the
// NAL payload start incl. the 2-byte NAL header), used to convert entry points.
std::string parse_slice_header(const uint8_t* rbsp, size_t n, int nal_type,
                               const Sps* sps_table, const Pps* pps_table,
                               const SliceHeader* prev_independent,
                               const std::vector<uint32_t>& skipped, SliceHeader& sh);

}  // namespace hc
