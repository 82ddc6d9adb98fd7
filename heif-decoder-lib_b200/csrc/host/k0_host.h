// k0_host.h — host-side preparation of a picture for K0, the device CABAC parser (kernels/k0_core.cuh):
// parameter sets and slice headers are parsed here (cheap, bit-serial), the slice DATA is only unescaped and
// shipped as bytes. Also hosts the CPU execution of the K0 core that the tests use to validate it.
#pragma once
#include <memory>
#include <string>
#include <vector>
#include "../kernels/k0_core.cuh"
#include "hevc_parse.h"

namespace hc {

struct K0HostPicture {
  bool eligible = false;
  std::string why_not;               // why the host parser has to take this picture
  hc_pic hpic;                       // picture header exactly as the host parser would emit it
  k0::Pic pic;                       // parameters + capacities; every pointer is left null here
  std::vector<k0::Slice> slices;
  std::vector<int32_t> ctb_slice;
  std::vector<uint8_t> ctu_static;   // 4 bytes per CTB
  std::vector<uint8_t> bytes;        // RBSP of all slice segments (each followed by 16 zero bytes)
  std::vector<k0::Sub> subs;         // Sub::pic is 0 here
  std::vector<k0::Chain> chains;     // indices into subs
  std::vector<uint8_t> scaling;      // HC_SCALING_BLOB_BYTES or empty
};

// Parses the headers of one coded picture given as 4-byte-length-prefixed / Annex-B NAL units.
std::string k0_prepare(const uint8_t* data, size_t size, int stream_format, K0HostPicture& out);

// The read-only tables of the K0 core (built once).
const k0::Tables& k0_tables();

// Runs the K0 core on the CPU and compacts its fixed-capacity output into the host parser's record form.
std::unique_ptr<PictureRecords> k0_parse_on_cpu(const K0HostPicture& hp, std::string* err);

}  // namespace hc
