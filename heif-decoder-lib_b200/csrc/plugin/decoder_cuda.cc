// decoder_cuda.cc — libheif-cuda.so: a heif_decoder_plugin for HEVC that reconstructs on the B200.
//
// Drop-in counterpart of the reference's libheif/plugins/decoder_libde265.cc (the ABI is
// libheif/api/libheif/heif_plugin.h:53-112, re-declared in include/heifcuda_plugin.h):
//   new_decoder   (decoder_libde265.cc:160-190)  -> a parser handle; the engine is process-wide
//   push_data     (:269-303)  4-byte big-endian length-prefixed NAL units -> host CABAC parse
//   decode_image  (:306-369)  records -> K1..K4 on the GPU -> D2H straight into heif_image planes,
//                             conformance-window sized, nclx from the VUI (convert :88-157)
//   free_decoder / set_strict_decoding (:193-207, :372-376)
// libheif creates one decoder per coded item and may run several concurrently (std::async per grid
// tile, context.cc:2385-2387): decoder handles are cheap, the engine (device memory pools, streams)
// is a thread-safe singleton created on first use.
//
// There is no CPU fallback: without a CUDA device decode_image fails with
// heif_error_Decoder_plugin_error and libheif reports the error to the caller.
#include <dlfcn.h>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include "../../../include/heifcuda.h"
#include "../../../include/heifcuda_plugin.h"

namespace {

// ---- callbacks into the libheif that loaded us (heif.h:2342,2376,1731,1781,1453-1456,1434-1440,1754,1776)
struct HeifApi {
  hcp_error (*image_create)(int, int, int, int, void**) = nullptr;
  hcp_error (*image_add_plane)(void*, int, int, int, int) = nullptr;
  uint8_t* (*image_get_plane)(void*, int, int*) = nullptr;
  void (*image_release)(const void*) = nullptr;
  hcp_nclx* (*nclx_alloc)(void) = nullptr;
  void (*nclx_free)(hcp_nclx*) = nullptr;
  hcp_error (*nclx_set_primaries)(hcp_nclx*, uint16_t) = nullptr;
  hcp_error (*nclx_set_transfer)(hcp_nclx*, uint16_t) = nullptr;
  hcp_error (*nclx_set_matrix)(hcp_nclx*, uint16_t) = nullptr;
  hcp_error (*image_set_nclx)(void*, const hcp_nclx*) = nullptr;
  void (*image_add_warning)(void*, hcp_error) = nullptr;
  bool ok = false;
};

HeifApi g_api;
std::once_flag g_api_once, g_engine_once;
hc_engine* g_engine = nullptr;
std::string g_engine_error;

void resolve_api() {
  void* h = RTLD_DEFAULT;
  if (const char* path = getenv("HEIFCUDA_LIBHEIF")) {
    void* lib = dlopen(path, RTLD_NOW | RTLD_LOCAL);   // usually already loaded: this only fetches its handle
    if (lib) h = lib;
  }
  auto sym = [&](const char* name) -> void* {
    void* p = dlsym(h, name);
    if (!p && h != RTLD_DEFAULT) p = dlsym(RTLD_DEFAULT, name);
    return p;
  };
#define HCP_BIND(field, name) g_api.field = reinterpret_cast<decltype(g_api.field)>(sym(name))
  HCP_BIND(image_create, "heif_image_create");
  HCP_BIND(image_add_plane, "heif_image_add_plane");
  HCP_BIND(image_get_plane, "heif_image_get_plane");
  HCP_BIND(image_release, "heif_image_release");
  HCP_BIND(nclx_alloc, "heif_nclx_color_profile_alloc");
  HCP_BIND(nclx_free, "heif_nclx_color_profile_free");
  HCP_BIND(nclx_set_primaries, "heif_nclx_color_profile_set_color_primaries");
  HCP_BIND(nclx_set_transfer, "heif_nclx_color_profile_set_transfer_characteristics");
  HCP_BIND(nclx_set_matrix, "heif_nclx_color_profile_set_matrix_coefficients");
  HCP_BIND(image_set_nclx, "heif_image_set_nclx_color_profile");
  HCP_BIND(image_add_warning, "heif_image_add_decoding_warning");
#undef HCP_BIND
  g_api.ok = g_api.image_create && g_api.image_add_plane && g_api.image_get_plane && g_api.image_release && g_api.nclx_alloc &&
             g_api.nclx_free && g_api.nclx_set_primaries && g_api.nclx_set_transfer && g_api.nclx_set_matrix && g_api.image_set_nclx &&
             g_api.image_add_warning;
}

void create_engine() {
  int device = 0;
  if (const char* d = getenv("HEIFCUDA_DEVICE")) device = atoi(d);
  g_engine = hc_engine_create(device);
  if (!g_engine) g_engine_error = hc_last_error();
}

const char kSuccess[] = "Success";
const hcp_error kOk = {HCP_ERROR_OK, HCP_SUBERROR_UNSPECIFIED, kSuccess};

// error text must outlive the call: one buffer per thread (libheif copies it at once, error.cc)
hcp_error plugin_error(const std::string& text, int sub = HCP_SUBERROR_UNSPECIFIED) {
  thread_local std::string keep;
  keep = text.empty() ? std::string("decoder_cuda: unknown error") : text;
  return {HCP_ERROR_DECODER_PLUGIN, sub, keep.c_str()};
}

struct Decoder {
  hc_parser* parser = nullptr;
  bool strict = false;
  bool pushed = false;
};

const char* cuda_plugin_name() { return "B200 CUDA HEVC-intra decoder (heifcuda), no CPU fallback"; }
void cuda_init_plugin() {}
void cuda_deinit_plugin() {
  // the engine stays alive until process exit: decoded images never reference device memory,
  // but other threads may still be inside decode_image when libheif unloads plugins (init.cc:141-180)
}

int cuda_does_support_format(int format) { return format == HCP_COMPRESSION_HEVC ? 200 : 0; }

hcp_error cuda_new_decoder(void** dec, int /*nthreads: the GPU engine has no worker threads*/) {
  if (!dec) return plugin_error("new_decoder: null argument");
  Decoder* d = new (std::nothrow) Decoder;
  if (!d) return {HCP_ERROR_MEMORY_ALLOCATION, HCP_SUBERROR_UNSPECIFIED, "decoder_cuda: out of memory"};
  d->parser = hc_parser_new();
  if (!d->parser) {
    delete d;
    return {HCP_ERROR_MEMORY_ALLOCATION, HCP_SUBERROR_UNSPECIFIED, "decoder_cuda: out of memory"};
  }
  *dec = d;
  return kOk;
}

void cuda_free_decoder(void* raw) {
  Decoder* d = static_cast<Decoder*>(raw);
  if (!d) return;
  hc_parser_free(d->parser);
  delete d;
}

void cuda_set_strict_decoding(void* raw, int flag) {
  if (raw) static_cast<Decoder*>(raw)->strict = flag != 0;
}

hcp_error cuda_push_data(void* raw, const void* data, size_t size) {
  Decoder* d = static_cast<Decoder*>(raw);
  if (!d) return plugin_error("push_data: null decoder");
  // the reference rejects truncated NAL length fields with End_of_data (decoder_libde265.cc:276-290)
  const uint8_t* p = static_cast<const uint8_t*>(data);
  size_t pos = 0;
  while (pos < size) {
    if (size - pos < 4) return plugin_error("decoder_cuda: truncated NAL length", HCP_SUBERROR_END_OF_DATA);
    const size_t n = ((size_t)p[pos] << 24) | ((size_t)p[pos + 1] << 16) | ((size_t)p[pos + 2] << 8) | p[pos + 3];
    pos += 4;
    if (n > size - pos) return plugin_error("decoder_cuda: NAL unit exceeds the pushed data", HCP_SUBERROR_END_OF_DATA);
    pos += n;
  }
  if (hc_parser_push(d->parser, p, size, HC_STREAM_LENGTH_PREFIXED) != HC_OK) return plugin_error(hc_last_error());
  d->pushed = true;
  return kOk;
}

struct BatchGuard {
  hc_batch* b = nullptr;
  hc_records* r = nullptr;
  ~BatchGuard() {
    if (b) hc_batch_destroy(b);
    if (r) hc_records_free(r);
  }
};

hcp_error cuda_decode_image(void* raw, void** out_img) {
  Decoder* d = static_cast<Decoder*>(raw);
  if (!d || !out_img) return plugin_error("decode_image: null argument");
  *out_img = nullptr;
  std::call_once(g_api_once, resolve_api);
  if (!g_api.ok) return plugin_error("decoder_cuda: libheif callbacks (heif_image_create ...) are not visible; set HEIFCUDA_LIBHEIF");
  std::call_once(g_engine_once, create_engine);
  if (!g_engine) return plugin_error("decoder_cuda: " + g_engine_error);

  BatchGuard g;
  g.r = hc_parser_take_picture(d->parser);
  if (!g.r) return plugin_error(std::string("decoder_cuda: ") + hc_last_error());
  const hc_pic* pic = hc_records_pic(g.r);
  const bool mono = pic->chroma_format == 0;
  if (!mono && pic->bit_depth_c != pic->bit_depth_y)   // decoder_libde265.cc:115-121
    return plugin_error("Channels with different number of bits per pixel are not supported");

  g.b = hc_batch_create(g_engine);
  if (!g.b) return plugin_error(std::string("decoder_cuda: ") + hc_last_error());
  const int canvas = hc_batch_add_canvas(g.b, pic->crop_w, pic->crop_h, pic->chroma_format, pic->bit_depth_y, 0);
  if (canvas < 0 || hc_batch_add_picture(g.b, g.r, canvas, 0, 0, HC_ROLE_COLOUR, 0) < 0 || hc_batch_upload(g.b) != HC_OK ||
      hc_batch_reconstruct(g.b, HC_STAGE_ALL) != HC_OK)
    return plugin_error(std::string("decoder_cuda: ") + hc_last_error());

  void* img = nullptr;
  hcp_error err = g_api.image_create(pic->crop_w, pic->crop_h, mono ? HCP_COLORSPACE_MONOCHROME : HCP_COLORSPACE_YCBCR,
                                     pic->chroma_format /* heif_chroma == de265_chroma numerically */, &img);
  if (err.code) return err;
  const int SubW = (pic->chroma_format == 1 || pic->chroma_format == 2) ? 2 : 1, SubH = pic->chroma_format == 1 ? 2 : 1;
  void* plane_dst[3] = {nullptr, nullptr, nullptr};
  size_t plane_stride[3] = {0, 0, 0};
  const int nplanes = mono ? 1 : 3;
  for (int c = 0; c < nplanes; c++) {
    const int w = c ? (pic->crop_w + SubW - 1) / SubW : pic->crop_w, h = c ? (pic->crop_h + SubH - 1) / SubH : pic->crop_h;
    err = g_api.image_add_plane(img, HCP_CHANNEL_Y + c, w, h, pic->bit_depth_y);
    if (err.code) { g_api.image_release(img); return err; }
    int stride = 0;
    plane_dst[c] = g_api.image_get_plane(img, HCP_CHANNEL_Y + c, &stride);
    plane_stride[c] = (size_t)stride;
    if (!plane_dst[c]) { g_api.image_release(img); return plugin_error("decoder_cuda: heif_image_get_plane failed"); }
  }
  // one synchronisation for all planes (and the first point where a GPU error of this tile can surface)
  if (hc_batch_read_planes(g.b, canvas, nplanes, plane_dst, plane_stride) != HC_OK) {
    g_api.image_release(img);
    return plugin_error(std::string("decoder_cuda: ") + hc_last_error());
  }

  // nclx from the VUI, defaults 2/2/2 + limited range when absent (decoder_libde265.cc:339-362, vui.cc:93-97)
  hcp_nclx* nclx = g_api.nclx_alloc();
  if (nclx) {
    const hcp_error e[3] = {g_api.nclx_set_primaries(nclx, pic->colour_primaries), g_api.nclx_set_transfer(nclx, pic->transfer_characteristics),
                            g_api.nclx_set_matrix(nclx, pic->matrix_coeffs)};
    for (const hcp_error& x : e) {
      if (x.code == HCP_ERROR_OK) continue;
      if (d->strict) {   // HEIF_WARN_OR_FAIL, heif_plugin.h:290-301
        g_api.nclx_free(nclx);
        g_api.image_release(img);
        return x;
      }
      g_api.image_add_warning(img, x);
    }
    nclx->full_range_flag = pic->full_range ? 1 : 0;
    g_api.image_set_nclx(img, nclx);
    g_api.nclx_free(nclx);
  }
  *out_img = img;
  return kOk;
}

}  // namespace

extern "C" {

const hcp_decoder_plugin heifcuda_decoder_plugin = {
    3,
    cuda_plugin_name,
    cuda_init_plugin,
    cuda_deinit_plugin,
    cuda_does_support_format,
    cuda_new_decoder,
    cuda_free_decoder,
    cuda_push_data,
    cuda_decode_image,
    cuda_set_strict_decoding,
    "cuda",
};

hcp_plugin_info plugin_info = {1, HCP_PLUGIN_TYPE_DECODER, &heifcuda_decoder_plugin, nullptr};

}  // extern "C"
