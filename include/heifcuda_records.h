/*
 * heifcuda_records.h — packed per-picture / per-CTU records that the host HEVC-intra front-end
 * (CABAC slice parse, stays serial on the CPU) hands to the sm_100a reconstruction kernels.
 *
 * This is the data contract named in BASELINE.json `north_star`: "packed per-CTU coefficient,
 * mode and QP records".  It mirrors what the reference keeps as intermediate decoder state:
 *   - per-TB sparse coefficient list      third-party/libde265/libde265/decctx.h:88-96
 *   - per-TB qP', transform flags         third-party/libde265/libde265/decctx.h:73-112
 *   - per-4x4 deblock edge flags / bS     third-party/libde265/libde265/image.h:71-75
 *   - per-min-CB QP_Y                     third-party/libde265/libde265/image.h:225
 *   - per-CTB sao_info                    third-party/libde265/libde265/slice.h:458-465
 * but laid out for a GPU: flat little-endian POD arrays, picture-relative offsets, neighbour
 * availability resolved on the host so the kernels carry no slice/tile/z-scan logic.
 *
 * Plain C header: included by the CUDA kernels, the host parser, the C-ABI and by oracle/.
 */
#ifndef HEIFCUDA_RECORDS_H
#define HEIFCUDA_RECORDS_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* One non-zero quantised coefficient (level before dequantisation).
 * pos = x + y*nT inside its transform block.  (reference: coeffList/coeffPos, decctx.h:88-90) */
typedef struct hc_coeff {
  uint16_t pos;
  int16_t  level;
} hc_coeff;

/* hc_tb.type bits */
#define HC_TB_CIDX_MASK   0x03u  /* colour component 0..2                                   */
#define HC_TB_DST         0x04u  /* 4x4 luma intra: inverse DST-VII instead of DCT           */
#define HC_TB_TSKIP       0x08u  /* transform_skip_flag                                      */
#define HC_TB_BYPASS      0x10u  /* cu_transquant_bypass_flag                                */
#define HC_TB_RDPCM_H     0x20u  /* implicit RDPCM, horizontal accumulation                  */
#define HC_TB_RDPCM_V     0x40u  /* implicit RDPCM, vertical accumulation                    */
#define HC_TB_ROTATE      0x80u  /* transform_skip_rotation (4x4)                            */

/* One coded transform block (cbf=1).  K1 (dequant + inverse transform) consumes these;
 * output is nT*nT int16 residuals at resid_off (picture-relative, in int16 elements). */
typedef struct hc_tb {
  uint32_t coeff_off;   /* picture-relative index into hc_coeff[]                             */
  uint32_t resid_off;   /* picture-relative int16 element offset into the residual buffer     */
  uint16_t ncoeff;      /* 1..1024                                                            */
  uint16_t pic;         /* picture index inside the batch                                     */
  uint8_t  log2;        /* log2 of block size, 2..5                                           */
  uint8_t  qp;          /* qP' = QP + QpBdOffset for this component (transform.cc:396-402)    */
  uint8_t  type;        /* HC_TB_* bits                                                       */
  uint8_t  matrix_id;   /* scaling-list matrixId (0..5) when the picture uses scaling lists   */
} hc_tb;

/* hc_blk.flags bits */
#define HC_BLK_AVAIL_TL     0x01u  /* top-left reference sample available                      */
#define HC_BLK_HAS_RESID    0x02u  /* add residual at resid_off after prediction               */
#define HC_BLK_NO_EDGE_FLT  0x04u  /* disableIntraBoundaryFilter (intrapred.cc:318-320)        */
#define HC_BLK_PCM          0x08u  /* block is PCM: samples are in the residual buffer verbatim */

/* One intra prediction + reconstruction block (a transform-tree leaf of one component).
 * Neighbour availability (slice, tile, picture edge, z-scan order: intrapred.h:443-940) is
 * resolved by the host in units of 4 component samples:
 *   avail_left bit k : rows  y+4k .. y+4k+3 of column x-1   (k < 2*nT/4, top to bottom)
 *   avail_top  bit k : cols  x+4k .. x+4k+3 of row    y-1   (k < 2*nT/4, left to right) */
typedef struct hc_blk {
  uint16_t x, y;        /* top-left, in samples of this component's plane                     */
  uint8_t  log2;        /* 2..5 (6 is never produced: max TB is 32)                           */
  uint8_t  mode;        /* IntraPredMode 0..34 (chroma already derived / 4:2:2-remapped)      */
  uint8_t  flags;       /* HC_BLK_*                                                           */
  uint8_t  cidx;        /* 0..2                                                               */
  uint16_t avail_left;
  uint16_t avail_top;
  uint32_t resid_off;   /* picture-relative int16 element offset                              */
} hc_blk;

/* hc_ctu.sao_nb bits: neighbouring CTB usable by the SAO edge classifier
 * (picture edge, slice_loop_filter_across_slices, loop_filter_across_tiles: sao.cc:349-423) */
#define HC_NB_L  0x01u
#define HC_NB_R  0x02u
#define HC_NB_T  0x04u
#define HC_NB_B  0x08u
#define HC_NB_TL 0x10u
#define HC_NB_TR 0x20u
#define HC_NB_BL 0x40u
#define HC_NB_BR 0x80u

#define HC_CTU_HAS_NOFILTER 0x01u  /* CTB contains pcm or transquant-bypass CUs                    */
#define HC_CTU_DEBLOCK_OFF  0x02u  /* slice_deblocking_filter_disabled_flag for this CTB's slice  */
#define HC_CTU_SAO_C_SELF   0x04u  /* chroma SAO: neighbours inside this CTB count as unavailable for
                                      samples on the CTB border (reference quirk, sao.cc:283,377-389:
                                      the "current" slice is looked up at chroma coordinates)      */

/* Per-CTB record, indexed by raster CTB address inside the picture. */
typedef struct hc_ctu {
  uint32_t blk_first[3];   /* picture-relative index of this CTB's first hc_blk, per component */
  uint16_t blk_count[3];
  uint8_t  sao_type[3];    /* 0 off, 1 band, 2 edge (already gated by slice_sao_luma/chroma)   */
  uint8_t  sao_band_or_class[3]; /* band position 0..31 or edge class 0..3                     */
  int8_t   sao_offset[3][4];     /* SaoOffsetVal (already << log2OffsetScale)                  */
  int8_t   beta_offset;    /* slice_beta_offset_div2*2 of the slice covering this CTB          */
  int8_t   tc_offset;      /* slice_tc_offset_div2*2                                           */
  uint8_t  sao_nb;         /* HC_NB_* for luma                                                 */
  uint8_t  flags;          /* HC_CTU_* */
  uint8_t  sao_nb_c;       /* HC_NB_* for chroma (differs from sao_nb only through the quirk above) */
  uint8_t  pad[3];
} hc_ctu;

/* hc_pic.edge_map byte per 4x4 luma unit */
#define HC_EDGE_V       0x01u  /* filter the vertical edge at the left of this unit (bS=2)      */
#define HC_EDGE_H       0x02u  /* filter the horizontal edge at the top of this unit (bS=2)     */
#define HC_EDGE_PCM     0x04u  /* unit belongs to a pcm coding unit                             */
#define HC_EDGE_BYPASS  0x08u  /* unit belongs to a cu_transquant_bypass coding unit            */

#define HC_PIC_STRONG_INTRA      0x0001u /* sps strong_intra_smoothing_enabled_flag              */
#define HC_PIC_NO_INTRA_SMOOTH   0x0002u /* sps range-ext intra_smoothing_disabled_flag          */
#define HC_PIC_HAS_DEBLOCK       0x0004u /* at least one edge is flagged                         */
#define HC_PIC_HAS_SAO           0x0008u /* at least one CTB component has sao_type != 0         */
#define HC_PIC_SCALING_LIST      0x0010u /* scaling_list_enabled: hc_pic.scaling_off is valid    */
#define HC_PIC_LIMITED_RANGE     0x0020u /* VUI video_full_range_flag == 0 (or VUI absent)       */
#define HC_PIC_PCMF              0x0040u /* (pcm_enabled && pcm_loop_filter_disabled) || transquant_bypass_enabled:
                                            the reference's special deblocking path (deblock.cc:724,755-790) */
#define HC_PIC_PCM_LF_DISABLED   0x0080u /* pcm_loop_filter_disabled_flag                         */

#define HC_DST_RESCALE_LIMITED 0x01u /* limited->full range rescale while pasting (context.cc:2504-2528) */
#define HC_DST_SKIP_Y  0x02u          /* component not written to the destination                  */
#define HC_DST_SKIP_CB 0x04u
#define HC_DST_SKIP_CR 0x08u

/* Per-picture header. All *_base are batch-global indices of this picture's first element;
 * record-internal offsets are relative to them. Plane/destination fields are filled by the
 * engine when it places the picture in device memory (host parser leaves them zero). */
typedef struct hc_pic {
  int32_t  width, height;        /* coded luma size (multiple of MinCbSize)                     */
  int32_t  crop_x, crop_y;       /* conformance window origin, luma samples                     */
  int32_t  crop_w, crop_h;       /* conformance window size = output size, luma samples         */
  uint8_t  chroma_format;        /* 0 mono, 1 4:2:0, 2 4:2:2, 3 4:4:4 (== heif_chroma / de265_chroma) */
  uint8_t  bit_depth_y, bit_depth_c;
  uint8_t  log2_ctb;
  uint16_t ctbs_w, ctbs_h;
  uint16_t flags;                /* HC_PIC_*                                                    */
  int8_t   pps_cb_qp_offset, pps_cr_qp_offset;   /* deblock chroma cQpPicOffset (deblock.cc:1652) */
  /* VUI colour description (defaults 2/2/2 when absent: vui.cc:93-97) */
  uint8_t  colour_primaries, transfer_characteristics, matrix_coeffs, full_range;
  uint8_t  dst_flags;            /* HC_DST_* (engine-filled)                                     */
  uint8_t  pad0;
  /* batch-global bases */
  uint32_t ctu_base;             /* hc_ctu[] : ctbs_w*ctbs_h entries                            */
  uint32_t blk_base;             /* hc_blk[]                                                    */
  uint32_t blk_count;
  uint32_t tb_base;              /* hc_tb[]                                                     */
  uint32_t tb_count;
  uint32_t coeff_base;           /* hc_coeff[]                                                  */
  uint32_t coeff_count;
  uint32_t edge_base;            /* uint8 edge_map[]: (width/4)*(height/4) bytes                 */
  uint32_t qp_base;              /* int8 qp_map[]: (width/8)*(height/8) bytes (QP_Y per 8x8)     */
  uint32_t scaling_base;         /* uint8 scaling factors (6*16+6*64+6*256+2*1024) if flagged    */
  uint64_t resid_base;           /* int16 element index into the residual buffer                */
  uint64_t resid_count;
  /* device placement (bytes from the plane-pool base), per component */
  uint64_t rec_off[3];           /* reconstruction / deblock planes (in place)                   */
  uint32_t rec_stride[3];        /* in samples                                                   */
  uint64_t dst_off[3];           /* final (post-SAO) planes: own picture or a grid canvas        */
  uint32_t dst_stride[3];        /* in samples                                                   */
  int32_t  dst_x, dst_y;         /* paste position in the destination, luma samples              */
  int32_t  dst_w, dst_h;         /* clip size available at the destination, luma samples         */
} hc_pic;

/* Sizes of the scaling-factor blob per picture: ScalingFactor for 4x4,8x8,16x16 (6 matrices
 * each) and 32x32 (2 matrices), row-major [matrix][y][x] (sps.h:54-57). */
#define HC_SCALING_BLOB_BYTES (6*16 + 6*64 + 6*256 + 2*1024)

#ifdef __cplusplus
}
#endif
#endif /* HEIFCUDA_RECORDS_H */
