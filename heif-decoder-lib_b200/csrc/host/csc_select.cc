// csc_select.cc — which colour conversion the reference would run, and with which coefficients.
//
// Mirrors, for the default heif_decoding_options, the outcome of the reference's Dijkstra search
// over its ColorConversionOperation pool (libheif/color-conversion/colorconversion.cc:266-420;
// resolved table in SURVEY.md §3.5) and the coefficient derivation of libheif/nclx.cc:82-171.
// The float expressions are written in the same order as the reference's so that fp32 rounding is
// identical; this file is compiled with -ffp-contract=off.
#include "../capi/capi_internal.h"
#include <cmath>

namespace {

struct KrKb { float Kr = 0.f, Kb = 0.f; };

// nclx.cc:93-140 (matrix 12/13 need the colour-primaries tables and are not provided)
KrKb kr_kb(int matrix) {
  KrKb r;
  switch (matrix) {
    case 1: r.Kr = 0.2126f; r.Kb = 0.0722f; break;
    case 4: r.Kr = 0.30f; r.Kb = 0.11f; break;
    case 5:
    case 6: r.Kr = 0.299f; r.Kb = 0.114f; break;
    case 7: r.Kr = 0.212f; r.Kb = 0.087f; break;
    case 9:
    case 10: r.Kr = 0.2627f; r.Kb = 0.0593f; break;
    default: break;
  }
  return r;
}

}  // namespace

extern "C" int hc_csc_select(int matrix, int primaries, int full_range, int chroma_format, int bit_depth,
                             int has_alpha, int out_format, hc_csc_params* out) {
  (void)primaries;
  (void)has_alpha;
  if (!out || out_format < HC_OUT_RGB || out_format > HC_OUT_RRGGBBAA_LE || bit_depth < 8 || bit_depth > 16) {
    hc::set_last_error("hc_csc_select: bad argument");
    return HC_ERR_ARGUMENT;
  }
  // matrix 2 (unspecified) reaches the conversion ops unchanged: Kr = Kb = 0 selects the literal
  // BT.601 defaults (nclx.cc:140-149,159-169), which differ in the last ulp from the values
  // computed for matrix 6.
  if (matrix == 11 || matrix == 14) {
    hc::set_last_error("matrix_coefficients 11/14 are not convertible (the reference rejects them too)");
    return HC_ERR_UNSUPPORTED;
  }
  if (matrix == 12 || matrix == 13) {
    hc::set_last_error("chromaticity-derived matrix_coefficients 12/13 are not implemented");
    return HC_ERR_UNSUPPORTED;
  }
  const bool wants8 = out_format == HC_OUT_RGB || out_format == HC_OUT_RGBA;
  if (wants8 != (bit_depth == 8)) {
    hc::set_last_error("8-bit images convert to RGB/RGBA, deeper images to RRGGBB(AA)");
    return HC_ERR_UNSUPPORTED;
  }
  hc_csc_params p;
  p.out_format = out_format;
  p.full_range = full_range ? 1 : 0;
  p.bit_depth = bit_depth;

  // nclx.cc:151-171 get_YCbCr_to_RGB_coefficients (defaults :140-149 when Kr = Kb = 0)
  const KrKb k = kr_kb(matrix);
  if (k.Kb != 0 || k.Kr != 0) {
    p.r_cr = 2 * (-k.Kr + 1);
    p.g_cb = 2 * k.Kb * (-k.Kb + 1) / (k.Kb + k.Kr - 1);
    p.g_cr = 2 * k.Kr * (-k.Kr + 1) / (k.Kb + k.Kr - 1);
    p.b_cb = 2 * (-k.Kb + 1);
  } else {
    p.r_cr = 1.402f;
    p.g_cb = -0.344136f;
    p.g_cr = -0.714136f;
    p.b_cb = 1.772f;
  }
  // yuv2rgb.cc:331-334
  p.r_cr_i = (int)std::lround(256 * p.r_cr);
  p.g_cr_i = (int)std::lround(256 * p.g_cr);
  p.g_cb_i = (int)std::lround(256 * p.g_cb);
  p.b_cb_i = (int)std::lround(256 * p.b_cb);

  if (matrix == 0) p.mode = HC_CSC_GBR;
  else if (matrix == 8) p.mode = HC_CSC_YCGCO;
  else if (bit_depth == 8 && chroma_format == 1 && full_range) p.mode = HC_CSC_INT420;  // with or without alpha
  else p.mode = HC_CSC_FLOAT;
  *out = p;
  return HC_OK;
}
