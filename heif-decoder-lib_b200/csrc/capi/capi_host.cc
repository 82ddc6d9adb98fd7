// capi_host.cc — C ABI over the host front-end (include/heifcuda.h, section "host front-end").
#include "../../../include/heifcuda.h"
#include "../host/hevc_parse.h"
#include <string>

namespace {
thread_local std::string g_last_error;
}

namespace hc {
void set_last_error(const std::string& s) { g_last_error = s; }
}  // namespace hc

struct hc_parser {
  hc::HevcIntraParser parser;
};
struct hc_records {
  std::unique_ptr<hc::PictureRecords> rec;
};

extern "C" {

const char* hc_last_error(void) { return g_last_error.c_str(); }

hc_parser* hc_parser_new(void) { return new (std::nothrow) hc_parser; }
void hc_parser_free(hc_parser* p) { delete p; }

int hc_parser_push(hc_parser* p, const uint8_t* data, size_t size, int stream_format) {
  if (!p || (!data && size)) {
    g_last_error = "hc_parser_push: null argument";
    return HC_ERR_ARGUMENT;
  }
  std::string e;
  try {
    switch (stream_format) {
      case HC_STREAM_LENGTH_PREFIXED: e = p->parser.push_length_prefixed(data, size); break;
      case HC_STREAM_ANNEXB: e = p->parser.push_annexb(data, size); break;
      case HC_STREAM_SINGLE_NAL: e = p->parser.push_nal(data, size); break;
      default: g_last_error = "hc_parser_push: unknown stream format"; return HC_ERR_ARGUMENT;
    }
  } catch (const std::bad_alloc&) {
    g_last_error = "out of memory while parsing";
    return HC_ERR_MEMORY;
  }
  if (!e.empty()) {
    g_last_error = e;
    return HC_ERR_BITSTREAM;
  }
  return HC_OK;
}

hc_records* hc_parser_take_picture(hc_parser* p) {
  if (!p) {
    g_last_error = "hc_parser_take_picture: null parser";
    return nullptr;
  }
  std::string e;
  std::unique_ptr<hc::PictureRecords> rec = p->parser.take_picture(&e);
  if (!rec) {
    g_last_error = e;
    return nullptr;
  }
  hc_records* r = new (std::nothrow) hc_records;
  if (!r) {
    g_last_error = "out of memory";
    return nullptr;
  }
  r->rec = std::move(rec);
  return r;
}

void hc_records_free(hc_records* r) { delete r; }
const hc_pic* hc_records_pic(const hc_records* r) { return &r->rec->pic; }
const hc_ctu* hc_records_ctus(const hc_records* r, size_t* n) { if (n) *n = r->rec->ctus.size(); return r->rec->ctus.data(); }
const hc_blk* hc_records_blks(const hc_records* r, size_t* n) { if (n) *n = r->rec->blks.size(); return r->rec->blks.data(); }
const hc_tb* hc_records_tbs(const hc_records* r, size_t* n) { if (n) *n = r->rec->tbs.size(); return r->rec->tbs.data(); }
const hc_coeff* hc_records_coeffs(const hc_records* r, size_t* n) { if (n) *n = r->rec->coeffs.size(); return r->rec->coeffs.data(); }
const uint8_t* hc_records_edge_map(const hc_records* r, size_t* n) { if (n) *n = r->rec->edge_map.size(); return r->rec->edge_map.data(); }
const int8_t* hc_records_qp_map(const hc_records* r, size_t* n) { if (n) *n = r->rec->qp_map.size(); return r->rec->qp_map.data(); }
const uint8_t* hc_records_scaling(const hc_records* r, size_t* n) {
  if (n) *n = r->rec->scaling.size();
  return r->rec->scaling.empty() ? nullptr : r->rec->scaling.data();
}
size_t hc_records_upload_bytes(const hc_records* r) {
  const hc::PictureRecords& p = *r->rec;
  return sizeof(hc_pic) + p.ctus.size() * sizeof(hc_ctu) + p.blks.size() * sizeof(hc_blk) +
         p.tbs.size() * sizeof(hc_tb) + p.coeffs.size() * sizeof(hc_coeff) + p.edge_map.size() + p.qp_map.size() +
         p.scaling.size();
}

hc_records* hc_parse_picture(const uint8_t* data, size_t size, int stream_format) {
  hc_parser p;
  if (hc_parser_push(&p, data, size, stream_format) != HC_OK) return nullptr;
  return hc_parser_take_picture(&p);
}

}  // extern "C"
