// capi_internal.h — definitions shared by the C-ABI translation units.
#pragma once
#include <memory>
#include <string>
#include "../../../include/heifcuda.h"
#include "../host/hevc_parse.h"
#include "../host/heif_reader.h"
#include "../host/k0_host.h"

struct hc_parser {
  hc::HevcIntraParser parser;
};
struct hc_records {
  std::unique_ptr<hc::PictureRecords> rec;
};
struct hc_k0_picture {
  hc::K0HostPicture hp;
};
struct hc_heif {
  hc::HeifFile file;
};

namespace hc {
void set_last_error(const std::string& s);
}
