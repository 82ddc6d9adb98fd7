/*
 * csc_oracle.c — CPU restatement of the reference's YCbCr -> interleaved RGB conversion chain.
 *
 * TEST INFRASTRUCTURE ONLY (see hevc_recon_oracle.c). Parity status: PINNED against the unmodified
 * reference's heif_decode_image() output on its bundled HEIC files (golden MD5s) and
 * against oracle/_ref on synthetic grids (tests/test_oracle_golden.py).
 *
 * Follows, with paths relative to /root/reference/libheif/:
 *   nclx.cc:82-171                 Kr/Kb and matrix coefficients
 *   colorconversion.cc:266-420     which op chain is chosen (resolved table: SURVEY.md §3.5)
 *   color-conversion/yuv2rgb.cc:260-495   Op_YCbCr420_to_RGB24 / _RGB32 (8-bit fixed point)
 *   color-conversion/yuv2rgb.cc:28-254    Op_YCbCr_to_RGB<Pixel> (fp32)
 *   color-conversion/yuv2rgb.cc:498-643   Op_YCbCr420_to_RRGGBBaa (fp32)
 *   color-conversion/rgb2rgb.cc:28-272,613-729  interleave / endianness ops
 *   common_utils.h:56-79           clip helpers
 * Compile with -ffp-contract=off: every float product and sum is rounded separately, as in the
 * reference's SSE2 build.
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>

static uint16_t clip_f_u16(float fx, int32_t maxi) { /* common_utils.h:64-70 */
  long x = (long)(fx + 0.5f);
  if (x < 0) return 0;
  if (x > maxi) return (uint16_t)maxi;
  return (uint16_t)x;
}
static uint8_t clip_int_u8(int x) { return x < 0 ? 0 : (x > 255 ? 255 : (uint8_t)x); }

typedef struct { float r_cr, g_cb, g_cr, b_cb; } coeffs_t;

static coeffs_t get_coeffs(int matrix) { /* nclx.cc:82-171 */
  float Kr = 0.f, Kb = 0.f;
  switch (matrix) {
    case 1: Kr = 0.2126f; Kb = 0.0722f; break;
    case 4: Kr = 0.30f; Kb = 0.11f; break;
    case 5: case 6: Kr = 0.299f; Kb = 0.114f; break;
    case 7: Kr = 0.212f; Kb = 0.087f; break;
    case 9: case 10: Kr = 0.2627f; Kb = 0.0593f; break;
    default: break;
  }
  coeffs_t c;
  if (Kb != 0 || Kr != 0) {
    c.r_cr = 2 * (-Kr + 1);
    c.g_cb = 2 * Kb * (-Kb + 1) / (Kb + Kr - 1);
    c.g_cr = 2 * Kr * (-Kr + 1) / (Kb + Kr - 1);
    c.b_cb = 2 * (-Kb + 1);
  } else {
    c.r_cr = 1.402f; c.g_cb = -0.344136f; c.g_cr = -0.714136f; c.b_cb = 1.772f;
  }
  return c;
}

/* out_format: 0 RGB, 1 RGBA, 2 RRGGBB_BE, 3 RRGGBBAA_BE, 4 RRGGBB_LE, 5 RRGGBBAA_LE.
 * Planes are uint16 arrays (any bit depth), strides in samples. `a` may be NULL.
 * Returns 0, or -1 for combinations the reference cannot convert. */
int hc_oracle_csc(const uint16_t* y, const uint16_t* cb, const uint16_t* cr, const uint16_t* a, int y_stride,
                  int c_stride, int a_stride, int width, int height, int chroma_format, int bit_depth, int matrix,
                  int full_range, int out_format, uint8_t* out, size_t out_stride) {
  /* matrix 2 (unspecified) reaches the ops unchanged: Kr = Kb = 0 -> the literal BT.601 defaults
   * (nclx.cc:140-149,159-169), NOT the values computed from Kr/Kb of matrix 6 */
  if (matrix == 11 || matrix == 14 || matrix == 12 || matrix == 13) return -1;
  const int to_alpha = out_format == 1 || out_format == 3 || out_format == 5;
  const int has_alpha = a != NULL;
  const coeffs_t k = get_coeffs(matrix);
  const int shiftH = (chroma_format == 1 || chroma_format == 2) ? 1 : 0;
  const int shiftV = chroma_format == 1 ? 1 : 0;
  const int maxv = (1 << bit_depth) - 1;
  const int half = 1 << (bit_depth - 1);
  /* the fixed-point ops win only for 8-bit 4:2:0 full-range input */
  /* (also when an alpha plane is present but not wanted: the pipeline drops it and still takes the
   * fixed-point op — checked against the reference with tests/golden/heic/alpha_420_8.heic) */
  const int use_int = bit_depth == 8 && chroma_format == 1 && full_range && matrix != 0 && matrix != 8;
  const int r_cr = (int)lround(256 * k.r_cr), g_cr = (int)lround(256 * k.g_cr);
  const int g_cb = (int)lround(256 * k.g_cb), b_cb = (int)lround(256 * k.b_cb);
  const float lro = (float)(16 << (bit_depth - 8));

  for (int yy = 0; yy < height; yy++)
    for (int x = 0; x < width; x++) {
      const int yv = y[(size_t)yy * y_stride + x];
      int cbv = half, crv = half;
      if (chroma_format) {
        cbv = cb[(size_t)(yy >> shiftV) * c_stride + (x >> shiftH)];
        crv = cr[(size_t)(yy >> shiftV) * c_stride + (x >> shiftH)];
      }
      int R, G, B;
      if (!chroma_format) {
        R = G = B = yv;
      } else if (use_int) { /* yuv2rgb.cc:349-361 */
        const int cbi = cbv - 128, cri = crv - 128;
        R = clip_int_u8(yv + ((r_cr * cri + 128) >> 8));
        G = clip_int_u8(yv + ((g_cb * cbi + g_cr * cri + 128) >> 8));
        B = clip_int_u8(yv + ((b_cb * cbi + 128) >> 8));
      } else if (matrix == 0) { /* yuv2rgb.cc:197-211 */
        if (full_range) { R = crv; G = yv; B = cbv; }
        else {
          R = clip_f_u16((crv - lro) * 1.1429f, maxv);
          G = clip_f_u16((yv - lro) * 1.1689f, maxv);
          B = clip_f_u16((cbv - lro) * 1.1429f, maxv);
        }
      } else if (matrix == 8) { /* yuv2rgb.cc:212-226 */
        const int c1 = cbv - half, c2 = crv - half;
        R = clip_int_u8(yv - c1 + c2);
        G = clip_int_u8(yv + c1);
        B = clip_int_u8(yv - c1 - c2);
      } else { /* yuv2rgb.cc:227-244 */
        float fy = (float)yv, fcb = (float)(cbv - half), fcr = (float)(crv - half);
        if (!full_range) {
          fy = (fy - lro) * 1.1689f;
          fcb = fcb * 1.1429f;
          fcr = fcr * 1.1429f;
        }
        R = clip_f_u16(fy + k.r_cr * fcr, maxv);
        G = clip_f_u16(fy + k.g_cb * fcb + k.g_cr * fcr, maxv);
        B = clip_f_u16(fy + k.b_cb * fcb, maxv);
      }
      const int A = has_alpha ? a[(size_t)yy * a_stride + x] : maxv;
      uint8_t* o = out + (size_t)yy * out_stride;
      if (out_format == 0) {
        o[3 * x + 0] = (uint8_t)R; o[3 * x + 1] = (uint8_t)G; o[3 * x + 2] = (uint8_t)B;
      } else if (out_format == 1) {
        o[4 * x + 0] = (uint8_t)R; o[4 * x + 1] = (uint8_t)G; o[4 * x + 2] = (uint8_t)B; o[4 * x + 3] = (uint8_t)A;
      } else {
        const int le = out_format >= 4, ps = to_alpha ? 8 : 6;
        const int v[4] = {R, G, B, A};
        for (int c = 0; c < (to_alpha ? 4 : 3); c++) {
          o[ps * x + 2 * c + le] = (uint8_t)(v[c] >> 8);
          o[ps * x + 2 * c + 1 - le] = (uint8_t)(v[c] & 0xff);
        }
      }
    }
  return 0;
}
