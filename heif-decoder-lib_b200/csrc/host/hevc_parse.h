// hevc_parse.h — host front-end: HEVC-intra NAL stream -> packed reconstruction records.
//
// This is the serial part of the hot path that BASELINE.json's north_star keeps on the CPU:
// NAL framing, parameter sets, slice headers and the CABAC slice-data parse.  It performs NO
// sample reconstruction; its output (PictureRecords) is the input of the sm_100a kernels.
// Reference counterpart (whose interleaved parse+reconstruct loop this splits in two):
//   third-party/libde265/libde265/slice.cc:2886-5346, transform.cc:31-210, decctx.cc:1209-1290.
#pragma once
#include <cstdint>
#include <memory>
#include <string>
#include <vector>
#include "../../../include/heifcuda_records.h"
#include "hevc_params.h"

namespace hc {

// All records of one coded picture (one HEVC item: a single image or one grid tile).
struct PictureRecords {
  hc_pic pic;                      // bases / placement fields are filled later by the engine
  std::vector<hc_ctu> ctus;        // ctbs_w * ctbs_h, raster order
  std::vector<hc_blk> blks;        // grouped per CTB, per component
  std::vector<hc_tb> tbs;
  std::vector<hc_coeff> coeffs;
  std::vector<uint8_t> edge_map;   // (width/4)*(height/4)
  std::vector<int8_t> qp_map;      // (width/8)*(height/8)
  std::vector<uint8_t> scaling;    // HC_SCALING_BLOB_BYTES or empty
  uint64_t resid_count = 0;        // int16 elements needed in the residual buffer
  uint32_t tbs_by_size[4] = {0, 0, 0, 0};   // coded transform blocks per log2 size 2..5 (K1 launch lists)
  std::vector<std::string> warnings;
};

class HevcIntraParser {
 public:
  HevcIntraParser();
  ~HevcIntraParser();

  // Feeds one NAL unit (2-byte NAL header first, no start code / length prefix).
  // Returns "" or an error text. Parameter sets are remembered; slice segments are parsed at once.
  std::string push_nal(const uint8_t* nal, size_t size);

  // Feeds a buffer of NAL units each prefixed by a 4-byte big-endian length, which is what
  // libheif hands to heif_decoder_plugin::push_data (libheif/plugins/decoder_libde265.cc:269-303).
  std::string push_length_prefixed(const uint8_t* data, size_t size);

  // Feeds an Annex-B byte stream (00 00 01 start codes), e.g. the *.265 test files.
  std::string push_annexb(const uint8_t* data, size_t size);

  // K0 (device parser) preparation: with collect-only set, slice segments are NOT parsed; their headers and
  // RBSP bytes are kept and take_k0() turns them into the inputs of kernels/k0_core.cuh.
  void set_collect_only(bool on);
  std::string take_k0(struct K0HostPicture& out);

  // Forgets every parameter set and any picture in progress, so that one parser object (its tables are ~700 KB) can
  // serve many independent coded items — a grid file is 48 of them — without being rebuilt for each.
  void reset();

  // True when every CTB of the current picture has been parsed.
  bool picture_complete() const;
  // True when at least one slice of a picture has been seen.
  bool picture_started() const;

  // Moves the finished picture out. If the picture is incomplete (missing slices) an error text
  // is returned through `err` and the records must not be used.
  std::unique_ptr<PictureRecords> take_picture(std::string* err);

 private:
  struct Impl;
  Impl* impl_;
};

}  // namespace hc
