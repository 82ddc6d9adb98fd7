// csc_pipeline_probe.cc — asks the UNMODIFIED reference which colour-conversion ops its pipeline search
// (ColorConversionPipeline::construct_pipeline, libheif/color-conversion/colorconversion.cc:266-420) picks for every
// input class x interleaved output format, with the default decoding options. Test infrastructure: its output is committed
// as tests/golden/csc_pipelines.json (tests/golden/make_csc_pipelines.py builds and runs it against oracle/_ref) and pins
// hc_csc_select (csrc/host/csc_select.cc), the product's static mirror of that search. Not part of any product library.
//
// The colour states are prepared exactly like convert_colorspace does (colorconversion.cc:521-587): unspecified nclx values
// replaced by the sRGB defaults, interleaved RGB(A) forced to 8 bit, RRGGBB(AA) to the input depth or 10 bit.
#include <cstdio>
#include <string>
#include "libheif/color-conversion/colorconversion.h"

int main(int argc, char** argv) {
  // argument "bilinear": the options heif-dec -C bilinear sets (examples/heif_dec.cc:502-509): bilinear chroma upsampling only
  const bool only_bilinear = argc > 1 && std::string(argv[1]) == "bilinear";
  ColorConversionPipeline::init_ops();
  const heif_chroma chromas[4] = {heif_chroma_monochrome, heif_chroma_420, heif_chroma_422, heif_chroma_444};
  const int depths[3] = {8, 10, 12};
  const int matrices[] = {0, 1, 2, 5, 6, 8, 9, 12};
  const heif_chroma outs[6] = {heif_chroma_interleaved_RGB, heif_chroma_interleaved_RGBA, heif_chroma_interleaved_RRGGBB_BE,
                               heif_chroma_interleaved_RRGGBBAA_BE, heif_chroma_interleaved_RRGGBB_LE, heif_chroma_interleaved_RRGGBBAA_LE};
  printf("[\n");
  bool first = true;
  for (int ci = 0; ci < 4; ci++)
    for (int bpp : depths)
      for (int full = 0; full < 2; full++)
        for (int matrix : matrices)
          for (int alpha = 0; alpha < 2; alpha++)
            for (int oi = 0; oi < 6; oi++) {
              ColorState a, b;
              a.colorspace = ci == 0 ? heif_colorspace_monochrome : heif_colorspace_YCbCr;
              a.chroma = chromas[ci];
              a.has_alpha = alpha != 0;
              a.bits_per_pixel = bpp;
              a.nclx_profile.set_matrix_coefficients((uint16_t)matrix);
              a.nclx_profile.set_full_range_flag(full != 0);
              a.nclx_profile.replace_undefined_values_with_sRGB_defaults();
              b = a;
              b.colorspace = heif_colorspace_RGB;
              b.chroma = outs[oi];
              b.has_alpha = (oi & 1) != 0;
              b.bits_per_pixel = oi < 2 ? 8 : (bpp > 8 ? bpp : 10);
              heif_color_conversion_options opt;
              opt.version = 1;
              opt.preferred_chroma_downsampling_algorithm = heif_chroma_downsampling_average;
              opt.preferred_chroma_upsampling_algorithm = heif_chroma_upsampling_bilinear;
              opt.only_use_preferred_chroma_algorithm = only_bilinear;
              ColorConversionPipeline p;
              std::string ops;
              if (p.construct_pipeline(a, b, opt)) {
                // "final pipeline has N steps:\n> <typeid name>\n> ..."
                const std::string d = p.debug_dump_pipeline();
                size_t pos = 0;
                while ((pos = d.find("> ", pos)) != std::string::npos) {
                  const size_t e = d.find('\n', pos);
                  std::string name = d.substr(pos + 2, e == std::string::npos ? std::string::npos : e - pos - 2);
                  size_t k = 0;
                  while (k < name.size() && name[k] >= '0' && name[k] <= '9') k++;   // mangled name: length prefix
                  if (!ops.empty()) ops += ",";
                  ops += name.substr(k);
                  pos = e == std::string::npos ? d.size() : e;
                }
              } else {
                ops = "NONE";
              }
              printf("%s {\"chroma\": %d, \"depth\": %d, \"full\": %d, \"matrix\": %d, \"alpha\": %d, \"out\": %d, \"ops\": \"%s\"}", first ? "" : ",\n", ci,
                     bpp, full, matrix, alpha, oi, ops.c_str());
              first = false;
            }
  printf("\n]\n");
  return 0;
}
