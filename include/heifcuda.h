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

#ifdef __cplusplus
}
#endif
#endif /* HEIFCUDA_H */
