/*
 * heifcuda.h — C ABI of the B200 HEVC-intra reconstruction + colour-conversion engine.
 *
 * Plain C, no torch / CUDA types in any signature.  Two layers:
 *
 *  (1) host front-end (CPU, serial CABAC parse): hc_parser_*  -> hc_records
 *      replaces the parse half of  third-party/libde265/libde265/slice.cc:5179-5346
 *      (decode_substream / read_coding_tree_unit) that libde265_v1_push_data / de265_decode
 *      (libheif/plugins/decoder_libde265.cc:269-369) drive in the reference.
 *
 *  (2) device engine (sm_100a kernels): hc_engine_*  reconstructs batches of parsed pictures and
 *      converts them to interleaved RGB.  Replaces the reconstruct half of the same loop
 *      (decode_TU slice.cc:3741, scale_coefficients transform.cc:692, decode_intra_prediction
 *      intrapred.cc:337, apply_deblocking_filter deblock.cc:1921,
 *      apply_sample_adaptive_offset_sequential sao.cc:552) and libheif's colour conversion
 *      (convert_colorspace colorconversion.cc:487, Op_YCbCr420_to_RGB24 yuv2rgb.cc:260 ...).
 *
 * The libheif decoder-plugin (struct heif_decoder_plugin, libheif/api/libheif/heif_plugin.h:53-112)
 * built on top of this ABI lives in libheif-cuda.so and is declared in heifcuda_plugin.h.
 *
 * All functions returning int return 0 on success and a negative HC_ERR_* code otherwise;
 * hc_last_error() gives the text for the calling thread.
 */
#ifndef HEIFCUDA_H
#define HEIFCUDA_H

#include <stddef.h>
#include <stdint.h>
#include "heifcuda_records.h"

#ifdef __cplusplus
extern "C" {
#endif

#define HC_OK 0
#define HC_ERR_BITSTREAM (-1)   /* malformed / unsupported bitstream                           */
#define HC_ERR_ARGUMENT (-2)
#define HC_ERR_NO_DEVICE (-3)   /* no CUDA device / engine not built: there is NO CPU fallback */
#define HC_ERR_CUDA (-4)
#define HC_ERR_MEMORY (-5)
#define HC_ERR_UNSUPPORTED (-6)

const char* hc_last_error(void);
/* 1 if this build of the library contains the CUDA engine (hc_engine_*), else 0. */
int hc_has_cuda_engine(void);

/* ------------------------------------------------------------------ host front-end -------- */
typedef struct hc_parser hc_parser;
typedef struct hc_records hc_records;

#define HC_STREAM_LENGTH_PREFIXED 0 /* 4-byte big-endian NAL lengths: what libheif's push_data gets */
#define HC_STREAM_ANNEXB 1          /* 00 00 01 start codes (.265 / .bit files)                     */
#define HC_STREAM_SINGLE_NAL 2      /* exactly one NAL unit                                         */

hc_parser* hc_parser_new(void);
void hc_parser_free(hc_parser* p);
/* Feeds data; parameter sets are kept, slice segments are parsed immediately. */
int hc_parser_push(hc_parser* p, const uint8_t* data, size_t size, int stream_format);
/* Returns the records of the picture parsed so far and resets the parser for the next picture
 * (parameter sets stay). NULL + HC_ERR_BITSTREAM text if no complete picture is available. */
hc_records* hc_parser_take_picture(hc_parser* p);

void hc_records_free(hc_records* r);
const hc_pic* hc_records_pic(const hc_records* r);
const hc_ctu* hc_records_ctus(const hc_records* r, size_t* count);
const hc_blk* hc_records_blks(const hc_records* r, size_t* count);
const hc_tb* hc_records_tbs(const hc_records* r, size_t* count);
const hc_coeff* hc_records_coeffs(const hc_records* r, size_t* count);
const uint8_t* hc_records_edge_map(const hc_records* r, size_t* bytes);
const int8_t* hc_records_qp_map(const hc_records* r, size_t* bytes);
const uint8_t* hc_records_scaling(const hc_records* r, size_t* bytes); /* NULL if flat */
/* total bytes of all record arrays (the `R` term of the roofline: bytes uploaded per picture) */
size_t hc_records_upload_bytes(const hc_records* r);

/* One-call convenience: parse one coded picture (parameter sets + slice NALs). */
hc_records* hc_parse_picture(const uint8_t* data, size_t size, int stream_format);

/* ------------------------------------------------------------------ HEIF container -------- */
/* Minimal ISO-BMFF/HEIF item resolver used by the batch decode API (the plugin path gets NAL
 * units from libheif instead). Mirrors what libheif resolves before it calls the decoder plugin:
 * libheif/file.cc:1246-1536, libheif/context.cc:172-221,710-1240. */
typedef struct hc_heif hc_heif;

typedef struct hc_heif_image_info {
  uint32_t id;
  int32_t is_grid;
  int32_t width, height;       /* output size: grid output size or ispe of a single image       */
  int32_t rows, cols;          /* grid layout (1,1 for a single image)                          */
  uint32_t alpha_id;           /* alpha auxiliary image item, 0 if none                         */
  int32_t rot;                 /* irot: anti-clockwise quarter turns                            */
  int32_t mirror;              /* imir: -1 none, 0 vertical axis, 1 horizontal axis             */
  int32_t nclx_present;        /* container colr(nclx) overrides the bitstream VUI              */
  int32_t primaries, transfer, matrix, full_range;
} hc_heif_image_info;

/* `data` must stay valid until hc_heif_close. NULL + error text on malformed files. */
hc_heif* hc_heif_open(const uint8_t* data, size_t size);
void hc_heif_close(hc_heif* f);
uint32_t hc_heif_primary_id(const hc_heif* f);
/* writes up to `max` ids, returns the total number of top-level images */
int hc_heif_top_level_ids(const hc_heif* f, uint32_t* ids, int max);
int hc_heif_get_image_info(const hc_heif* f, uint32_t id, hc_heif_image_info* info);
/* row-major tile item ids of a grid; returns rows*cols or a negative error */
int hc_heif_grid_tiles(const hc_heif* f, uint32_t id, uint32_t* tiles, int max);
/* hvcC parameter sets + item data as 4-byte-length-prefixed NAL units (what libheif gives
 * heif_decoder_plugin::push_data). *out is malloc'ed; release with hc_free. */
int hc_heif_coded_stream(const hc_heif* f, uint32_t id, uint8_t** out, size_t* size);
void hc_free(void* p);

#ifdef __cplusplus
}
#endif
#endif /* HEIFCUDA_H */
