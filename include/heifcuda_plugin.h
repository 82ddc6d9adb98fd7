/*
 * heifcuda_plugin.h — the drop-in boundary: libheif's decoder-plugin ABI as implemented by
 * libheif-cuda.so (heif-decoder-lib_b200/csrc/plugin/decoder_cuda.cc).
 *
 * libheif-cuda.so is a `heif_decoder_plugin` for heif_compression_HEVC that sits next to the
 * reference's plugins/decoder_libde265.cc.  An unchanged libheif finds it through
 *   dlopen(file) + dlsym("plugin_info")        libheif/plugins_unix.cc:95-110, libheif/init.cc:211-267
 * when the file lies in a directory of $LIBHEIF_PLUGIN_PATH (init.cc:48-63), or an application hands
 * &heifcuda_decoder_plugin to heif_register_decoder_plugin() (heif.h:2408).  It outranks libde265
 * (priority 100, decoder_libde265.cc:43) with priority 200 and can be selected explicitly with
 * heif_decoding_options::decoder_id = "cuda" (plugin_registry.cc:231-255).
 *
 * This header re-declares, layout-compatibly, ONLY the part of the reference ABI the plugin
 * touches, so that the plugin builds without the reference tree.  Every declaration names the
 * reference declaration it must stay binary compatible with.  The plugin calls back into libheif
 * (heif_image_create, ...) through symbols resolved at run time from the libheif that loaded it.
 */
#ifndef HEIFCUDA_PLUGIN_H
#define HEIFCUDA_PLUGIN_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif
/* the libraries are built with -fvisibility=hidden -fno-gnu-unique: only this C ABI is exported, so that
 * no C++ symbol of the engine can bind to (or be bound by) another library of the host process */
#if defined(__GNUC__)
#pragma GCC visibility push(default)
#endif

/* struct heif_error                                   libheif/api/libheif/heif.h:373-384 */
typedef struct hcp_error {
  int code;            /* enum heif_error_code         heif.h:101-139 */
  int subcode;         /* enum heif_suberror_code      heif.h:142-370 */
  const char* message; /* never NULL                                  */
} hcp_error;

#define HCP_ERROR_OK 0                      /* heif_error_Ok                       heif.h:104 */
#define HCP_ERROR_MEMORY_ALLOCATION 6       /* heif_error_Memory_allocation_error  heif.h:122 */
#define HCP_ERROR_DECODER_PLUGIN 7          /* heif_error_Decoder_plugin_error     heif.h:125 */
#define HCP_SUBERROR_UNSPECIFIED 0          /* heif_suberror_Unspecified           heif.h:144 */
#define HCP_SUBERROR_END_OF_DATA 100        /* heif_suberror_End_of_data           heif.h:149 */
#define HCP_COMPRESSION_HEVC 1              /* heif_compression_HEVC               heif.h:413 */
#define HCP_COLORSPACE_YCBCR 0              /* heif_colorspace_YCbCr                          */
#define HCP_COLORSPACE_MONOCHROME 2         /* heif_colorspace_monochrome                     */
#define HCP_CHANNEL_Y 0                     /* heif_channel_Y / _Cb / _Cr = 0 / 1 / 2         */
#define HCP_PLUGIN_TYPE_DECODER 1           /* heif_plugin_type_decoder            heif.h:584-588 */

/* struct heif_decoder_plugin (the Aliyun fork's variant: new_decoder takes nthreads)
 *                                                     libheif/api/libheif/heif_plugin.h:53-112 */
typedef struct hcp_decoder_plugin {
  int plugin_api_version;                                            /* 3 */
  const char* (*get_plugin_name)(void);
  void (*init_plugin)(void);
  void (*deinit_plugin)(void);
  int (*does_support_format)(int /* enum heif_compression_format */ format);
  hcp_error (*new_decoder)(void** decoder, int nthreads);           /* heif_plugin.h:76 */
  void (*free_decoder)(void* decoder);
  hcp_error (*push_data)(void* decoder, const void* data, size_t size);
  hcp_error (*decode_image)(void* decoder, void /* struct heif_image */** out_img);
  void (*set_strict_decoding)(void* decoder, int flag);             /* api version 2 */
  const char* id_name;                                               /* api version 3 */
} hcp_decoder_plugin;

/* struct heif_plugin_info                             libheif/api/libheif/heif.h:590-596 */
typedef struct hcp_plugin_info {
  int version;         /* 1 */
  int type;            /* HCP_PLUGIN_TYPE_DECODER */
  const void* plugin;  /* -> hcp_decoder_plugin */
  void* internal_handle;
} hcp_plugin_info;

/* struct heif_color_profile_nclx (version 1 fields)   libheif/api/libheif/heif.h:1414-1431 */
typedef struct hcp_nclx {
  uint8_t version;
  int color_primaries;
  int transfer_characteristics;
  int matrix_coefficients;
  uint8_t full_range_flag;
  float primaries_xy[8];
} hcp_nclx;

/* The two symbols libheif-cuda.so exports. `plugin_info` is the name libheif's loader looks up
 * (plugins_unix.cc:103; pattern decoder_libde265.cc:416-422). */
extern hcp_plugin_info plugin_info;
extern const hcp_decoder_plugin heifcuda_decoder_plugin;

/* Environment:
 *   HEIFCUDA_DEVICE   CUDA device ordinal used by the plugin's engine (default 0)
 *   HEIFCUDA_LIBHEIF  path of the libheif shared object to call back into, for hosts that loaded
 *                     libheif with RTLD_LOCAL (default: symbols already visible in the process) */

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif /* HEIFCUDA_PLUGIN_H */
