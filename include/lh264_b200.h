/*
 * lh264_b200.h — C ABI of the B200-native hot path of losslessh264.
 *
 * This is the flat shim between the host decoder (NAL/slice/CAVLC parse, which stays on the
 * CPU) and the sm_100a CUDA kernels that do macroblock reconstruction.  Plain pointers and
 * sizes only: no C++ types, no torch types.  Everything here replaces a specific piece of the
 * reference (paths relative to the reference root, OpenH264 v1.4 fork "losslessh264"):
 *
 *   lh264_mb_rec / lh264_slice_rec    <- SDqLayer per-MB SoA hand-off record
 *                                        (codec/decoder/core/inc/dec_frame.h:61-127,
 *                                         allocated at decoder_core.cpp:1226-1348) and the
 *                                        slice-header fields the reconstruction reads
 *                                        (slice.h: SSliceHeader; deblocking.cpp:872-905).
 *   lh264_recon_submit()              <- WelsTargetSliceConstruction  decode_slice.cpp:110-206
 *                                        (= per MB WelsTargetMbConstruction :353-373, then
 *                                        WelsDeblockingFilterSlice deblocking.cpp:872-933),
 *                                        for ALL slices of one picture, for a BATCH of pictures
 *                                        belonging to independent streams.
 *   lh264_stream_open()/pictures      <- PPicture pool / DPB storage (pic_queue.cpp:60-107);
 *                                        border handling replaces ExpandReferencingPicture
 *                                        (expand_pic.cpp:37-174) by clamp-at-read.
 *   lh264_picture_read()              <- DecodeFrameConstruction output pointers
 *                                        (decoder_core.cpp:56-217): cropped I420 to host.
 *   lh264_k_*()                       <- the function-pointer kernel tables of
 *                                        SWelsDecoderContext (decoder_context.h:391-407):
 *                                        pIdctResAddPredFunc, sMcFunc, sDeblockingFunc,
 *                                        pGetI*PredFunc, sExpandPicFunc.  Batched, for
 *                                        kernel-level parity tests.
 *
 * Error convention: every function returns LH264_OK (0) or a negative LH264_E_* code; no
 * exceptions cross this boundary.  There is NO CPU fallback: with no usable CUDA device
 * lh264_ctx_create() fails with LH264_E_NO_DEVICE.
 */
#ifndef LH264_B200_H_
#define LH264_B200_H_

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LH264_ABI_VERSION 3

/* ---- error codes ------------------------------------------------------------------------ */
#define LH264_OK             0
#define LH264_E_INVALID     -1   /* bad argument / malformed record                         */
#define LH264_E_NO_DEVICE   -2   /* no CUDA device / driver; there is no CPU fallback        */
#define LH264_E_CUDA        -3   /* a CUDA call failed; see lh264_last_error()               */
#define LH264_E_NOMEM       -4
#define LH264_E_UNSUPPORTED -5

/* ---- macroblock record -------------------------------------------------------------------- */
/* mb_type (reference bitmask values wels_common_defs.h:264-276 are folded into an enum)     */
enum {
  LH264_MB_NONE   = 0,  /* not decoded in this picture: samples left untouched               */
  LH264_MB_I4x4   = 1,
  LH264_MB_I8x8   = 2,
  LH264_MB_I16x16 = 3,
  LH264_MB_IPCM   = 4,  /* 384 samples carried as int16 in the coefficient pool              */
  LH264_MB_P16x16 = 5,
  LH264_MB_P16x8  = 6,
  LH264_MB_P8x16  = 7,
  LH264_MB_P8x8   = 8,  /* also P_8x8ref0                                                    */
  LH264_MB_PSKIP  = 9
};
#define LH264_MB_IS_INTRA(t) ((t) >= LH264_MB_I4x4 && (t) <= LH264_MB_IPCM)
#define LH264_MB_IS_INTER(t) ((t) >= LH264_MB_P16x16)

enum { LH264_SUB_8x8 = 0, LH264_SUB_8x4 = 1, LH264_SUB_4x8 = 2, LH264_SUB_4x4 = 3 };

/* Resolved intra 4x4 / 8x8 predictor index.  0..8 are the H.264 modes; 9..13 are the
 * availability variants the reference resolves at parse time (pIntra4x4FinalMode,
 * wels_common_defs.h:314-331).                                                              */
enum {
  LH264_I4_V = 0, LH264_I4_H = 1, LH264_I4_DC = 2, LH264_I4_DDL = 3, LH264_I4_DDR = 4,
  LH264_I4_VR = 5, LH264_I4_HD = 6, LH264_I4_VL = 7, LH264_I4_HU = 8,
  LH264_I4_DC_L = 9, LH264_I4_DC_T = 10, LH264_I4_DC_128 = 11,
  LH264_I4_DDL_TOP = 12, /* top-right unavailable: replicate top[3] (top[7] for 8x8)       */
  LH264_I4_VL_TOP = 13
};
/* Resolved intra 16x16 luma predictor (wels_common_defs.h:302-311).                          */
enum { LH264_I16_V = 0, LH264_I16_H = 1, LH264_I16_DC = 2, LH264_I16_P = 3,
       LH264_I16_DC_L = 4, LH264_I16_DC_T = 5, LH264_I16_DC_128 = 6 };
/* Resolved intra chroma predictor (wels_common_defs.h:334-343).                              */
enum { LH264_C_DC = 0, LH264_C_H = 1, LH264_C_V = 2, LH264_C_P = 3,
       LH264_C_DC_L = 4, LH264_C_DC_T = 5, LH264_C_DC_128 = 6 };

#define LH264_MBF_TRANSFORM8x8 0x01

/* One macroblock, 128 bytes, 128-byte aligned in the batch (one L2 line / one warp-wide
 * 128-bit-per-lane... load).  Raster MB order inside a picture.
 *
 * Block index conventions: "z" = H.264 4x4 block decode order (8x8 quadrant major);
 * "r" = raster order of the 4x4 blocks inside the MB (r = 4*by + bx).                       */
typedef struct lh264_mb_rec {
  uint8_t  mb_type;           /* LH264_MB_*                                                  */
  uint8_t  flags;             /* LH264_MBF_*                                                 */
  uint8_t  cbp;               /* coded_block_pattern: luma bits 0..3, chroma bits 4..5       */
  uint8_t  qp_y;              /* luma QP (pLumaQp)                                           */
  uint8_t  qp_c[2];           /* Cb, Cr QP (pChromaQp[2])                                    */
  uint8_t  i16_mode;          /* LH264_I16_* (pIntraPredMode[7])                             */
  uint8_t  chroma_mode;       /* LH264_C_*   (pChromaPredMode)                               */
  uint16_t slice_idx;         /* index into the picture's slice table (pSliceIdc)            */
  uint8_t  nxn_avail;         /* I8x8 neighbour flags, bit0 top, bit1 top-left, bit2 left,
                                 bit3 top-right  (pIntraNxNAvailFlag, rec_mb.cpp:88-97)      */
  uint8_t  reserved0;
  uint32_t nz_mask;           /* bit r (r<16): luma 4x4 block r has non-zero count (pNzc);
                                 bits 16..19 Cb 2x2 raster, 20..23 Cr.  Used for bS.         */
  uint32_t coded_mask;        /* bit z (z<16 luma in z order; 16..19 Cb; 20..23 Cr): the
                                 16-coefficient block is present in the coefficient pool     */
  uint32_t coef_off;          /* index (in 16-coefficient blocks) of this MB's first present
                                 block in the picture's coefficient pool                     */
  int8_t   ref_idx[4];        /* ref_idx_l0 per 8x8 quadrant (raw index: bS compares these)  */
  uint8_t  sub_mb_type[4];    /* LH264_SUB_* per 8x8 quadrant (P8x8 only)                    */
  uint8_t  intra_mode[16];    /* LH264_I4_* per 4x4 block in z order; I8x8 uses [0,4,8,12]   */
  int16_t  mv[16][2];         /* final motion vectors (quarter-pel) per 4x4 block, r order   */
  uint8_t  reserved1[16];
} lh264_mb_rec;

/* Coefficient pool: int16 *quantised levels* (the reference's pScaledTCoeffQuant,
 * decoder_core.cpp:1226+; dequantisation moves to the GPU).  Present blocks of one MB are
 * consecutive, ordered by increasing bit index of coded_mask; each is 16 int16 in raster order
 * of the 4x4 block.  With LH264_MBF_TRANSFORM8x8 the four bits of a luma quadrant are set or
 * clear together and their 64 int16 are one raster 8x8 block.  Intra16x16 luma DC and chroma DC
 * levels sit in slot 0 of their block (same as the reference).  IPCM: coded_mask = 0xFFFFFF and
 * the 384 int16 are the samples: Y 16x16 raster, Cb 8x8 raster, Cr 8x8 raster.               */

#define LH264_MAX_REFS 16

typedef struct lh264_slice_rec {
  int32_t  first_mb;
  int32_t  num_mbs;
  uint8_t  slice_type;              /* 0 = P, 2 = I (H.264 slice_type % 5)                    */
  uint8_t  disable_deblocking_idc;  /* 0, 1 or 2                                              */
  int8_t   alpha_c0_offset;         /* already doubled (slice_alpha_c0_offset_div2 * 2)       */
  int8_t   beta_offset;             /* already doubled                                        */
  uint8_t  num_ref_idx_l0;
  uint8_t  weighted_pred;           /* explicit WP on (rec_mb.cpp:276-341)                    */
  uint8_t  luma_log2_denom;
  uint8_t  chroma_log2_denom;
  int16_t  ref_slot[LH264_MAX_REFS];/* picture slot of RefPicList0[i] after reordering; the
                                       host has already applied the "missing reference ->
                                       list entry 0" fallback (rec_mb.cpp:229-233)            */
  int16_t  luma_weight[LH264_MAX_REFS];
  int16_t  luma_offset[LH264_MAX_REFS];
  int16_t  chroma_weight[LH264_MAX_REFS][2];
  int16_t  chroma_offset[LH264_MAX_REFS][2];
  /* Scaling lists (High profile; WelsCalcDeqCoeffScalingList, decode_slice.cpp:1231-1277).  The
   * lists are the ones IN FORCE for the slice after the SPS / PPS fall-back rules
   * (ParseScalingList, au_parser.cpp:1502-1562), in raster position order.  The multiplier of a
   * level is list[pos] * LevelScale(qp, pos), applied as the reference applies it:
   * (level * m) >> 4 for 4x4 blocks (parse_mb_syn_cavlc.cpp:945,993), the 8x8 rounding formula
   * (parse_mb_syn_cavlc.cpp:1088-1090) for 8x8 blocks, all zero at qp 51 (the reference fills
   * its tables for qp 0..50 only).                                                           */
  uint8_t  scaling_flags;           /* LH264_SLF_*                                            */
  uint8_t  reserved[15];
  uint8_t  scale4[6][16];           /* Y intra, Cb intra, Cr intra, Y inter, Cb inter, Cr inter */
  uint8_t  scale8[2][64];           /* Y intra, Y inter                                       */
} lh264_slice_rec;

#define LH264_SLF_SCALING      0x01 /* scaling lists in use (bUseScalingList)                 */
#define LH264_SLF_SCALING_CAVLC 0x02 /* CAVLC slice: the reference indexes its 8x8 multiplier
                                       tables with the 4x4 residual property (decode_slice.cpp:3385,
                                       :3782 pass LUMA_DC_AC_INTRA / _INTER to
                                       WelsResidualBlockCavlc8x8, whose table index is property - 6),
                                       i.e. it reads pDequant_coeff4x4[0] / [3] as a 64-wide table.
                                       With this bit the 8x8 multiplier of (qp, pos) is element
                                       base + qp * 64 + pos of the flat 4x4 tables [6][52][16]
                                       followed by the flat 8x8 tables [6][52][64], base = 0 for an
                                       intra and 3 * 52 * 16 for an inter macroblock.           */

/* One picture to reconstruct (all of its slices).  All pointers are HOST pointers for
 * lh264_recon_submit(); the library copies what it needs before returning only if
 * LH264_SUBMIT_COPY is set, otherwise the memory must stay valid until lh264_sync().        */
typedef struct lh264_picture_desc {
  int32_t  stream;                  /* handle from lh264_stream_open                          */
  int32_t  cur_slot;                /* picture slot to reconstruct into                       */
  int32_t  n_slices;
  int32_t  n_coef_blocks;           /* blocks of 16 int16 in coefs                            */
  const lh264_slice_rec* slices;
  const lh264_mb_rec*    mbs;       /* mb_w*mb_h records, raster order                        */
  const int16_t*         coefs;
} lh264_picture_desc;

typedef struct lh264_ctx lh264_ctx;

/* ---- context / streams ---------------------------------------------------------------------- */
int  lh264_abi_version(void);
const char* lh264_last_error(void);

/* device: CUDA ordinal.  cuda_stream: a cudaStream_t to run on (e.g. the torch current
 * stream), or NULL for a library-owned stream.                                              */
int  lh264_ctx_create(lh264_ctx** out, int device, void* cuda_stream);
void lh264_ctx_destroy(lh264_ctx* ctx);

/* Opens a stream: allocates n_slots pictures of mb_w x mb_h macroblocks in HBM, pre-filled
 * with 128 (pic_queue.cpp:96).  Returns the stream handle (>= 0) or an error.               */
int  lh264_stream_open(lh264_ctx* ctx, int mb_w, int mb_h, int n_slots);
int  lh264_stream_close(lh264_ctx* ctx, int stream);

/* ---- the hot path ----------------------------------------------------------------------------- */
/* Reconstructs a batch of pictures (at most one per stream per call; streams are
 * independent).  Host buffers in, asynchronous on the context's CUDA stream:
 * H2D of records + coefficient pools, then the reconstruction kernels.                      */
int  lh264_recon_submit(lh264_ctx* ctx, const lh264_picture_desc* pics, int n_pics);

/* Resident variant used by the roofline bench: stage a batch in HBM once, replay it.        */
typedef struct lh264_resident_batch lh264_resident_batch;
int  lh264_batch_upload(lh264_ctx* ctx, const lh264_picture_desc* pics, int n_pics,
                        lh264_resident_batch** out);
int  lh264_batch_recon(lh264_ctx* ctx, lh264_resident_batch* batch);
void lh264_batch_free(lh264_ctx* ctx, lh264_resident_batch* batch);
/* number of kernel launches issued by this context so far                                   */
uint64_t lh264_kernel_launches(const lh264_ctx* ctx);
/* accumulated device time (ms, CUDA events on the context stream) of one kernel family
 * since the last reset: 0 k_decode_rows (reconstruction + deblocking, one fused kernel), 1-2 unused (0), 3 = 0, 4 model generation
 * kernels (k_model_frame/count/scan/emit), 5 k_model_code, 6 k_model_decode; family = -1 resets and switches the
 * (synchronising) per-family timing on, family = -2 switches it off                          */
float lh264_kernel_ms(lh264_ctx* ctx, int family);

/* Copies the cropped I420 picture to host memory (asynchronous unless sync != 0).
 * crop_* in luma samples, even.  Strides in bytes.                                          */
int  lh264_picture_read(lh264_ctx* ctx, int stream, int slot,
                        int crop_x, int crop_y, int width, int height,
                        uint8_t* dst_y, int stride_y,
                        uint8_t* dst_u, uint8_t* dst_v, int stride_c, int sync);
/* Uploads a full (uncropped, mb_w*16 x mb_h*16) I420 picture into a slot — used by tests and
 * by host-side error concealment.                                                           */
int  lh264_picture_write(lh264_ctx* ctx, int stream, int slot,
                         const uint8_t* src_y, int stride_y,
                         const uint8_t* src_u, const uint8_t* src_v, int stride_c);
/* Device pointers of a slot (for zero-copy consumers such as torch); pitch in bytes.        */
int  lh264_picture_device_ptrs(lh264_ctx* ctx, int stream, int slot,
                               void** y, void** u, void** v, int* pitch_y, int* pitch_c);
int  lh264_sync(lh264_ctx* ctx);
/* Orders the context stream after all model work submitted so far (the model runs on its own CUDA stream); for callers
 * that time or chain work with events on the stream they passed to lh264_ctx_create.                         */
int  lh264_join(lh264_ctx* ctx);
/* Waits until the host buffers handed to every lh264_recon_submit / lh264_model_submit so far have been copied to the device
 * (they may then be overwritten), without waiting for the kernels.                                              */
int  lh264_inputs_consumed(lh264_ctx* ctx);
/* Gathers the cropped I420 pictures of n (stream, slot) pairs on the device and brings them to
 * the host with ONE copy: for each picture Y (width x height) then U, V ((w+1)/2 x (h+1)/2),
 * tightly packed, pictures back to back.  dst should be pinned (lh264_host_alloc).  With sync == 0 the call returns once
 * the gather kernel has run; the copy to the host proceeds on a copy stream of its own, concurrently with later submits,
 * and is complete after lh264_output_wait() (or lh264_sync()).                                */
int  lh264_pictures_read_packed(lh264_ctx* ctx, const int* streams, const int* slots, int n,
                                int crop_x, int crop_y, int width, int height,
                                uint8_t* dst, size_t dst_bytes, int sync);
int  lh264_output_wait(lh264_ctx* ctx);
/* Finer grained: lh264_read_ticket() names the lh264_pictures_read_packed() call made last on this context (two may be in flight: the
 * output buffers inside the library alternate), lh264_output_wait_ticket() waits for THAT call's copy only - a consumer thread can
 * take the pictures of step t while the copy of step t+1 is still running.                                                    */
int  lh264_read_ticket(const lh264_ctx* ctx);
int  lh264_output_wait_ticket(lh264_ctx* ctx, int ticket);
/* Pinned host memory for hand-off records and output pictures (async copies need it).       */
void* lh264_host_alloc(size_t bytes);
/* The same, write-combined: for staging memory the host only WRITES (record arrays on their way to lh264_recon_submit /
 * lh264_submit_compact).  Stores stream out without read-for-ownership and the DMA reads do not snoop the caches; host reads
 * of such memory are uncached and very slow.  Freed with lh264_host_free.                                                  */
void* lh264_host_alloc_wc(size_t bytes);
void  lh264_host_free(void* p);
/* Makes `device` the calling THREAD's CUDA device.  The allocators above work on the calling thread's device: a helper thread
 * of a process that drives GPU n > 0 (one process per GPU) binds itself first, or its first allocation would open a context on
 * GPU 0.  Returns LH264_OK or LH264_E_NO_DEVICE.                                                                              */
int   lh264_thread_bind(int device);

/* ---- kernel-level entry points (parity tests against the reference kernel tables) ---------- */
/* n independent 4x4 blocks: pred/out u8[n][16], levels int16[n][16] already dequantised.
 * = IdctResAddPred_c (decode_mb_aux.cpp:42-77).                                              */
int  lh264_k_idct4x4_add(lh264_ctx* ctx, uint8_t* pix /*[n][16] in/out*/, const int16_t* coef, int n);
/* = IdctResAddPred8x8_c (decode_mb_aux.cpp:79-167): pix u8[n][64], coef int16[n][64].       */
int  lh264_k_idct8x8_add(lh264_ctx* ctx, uint8_t* pix, const int16_t* coef, int n);
/* = McLuma_c (mc.cpp:334-346) on a w x h reference plane (border replicated by clamping);
 * n blocks: x,y (integer position of the block), mvx,mvy (quarter-pel), bw,bh in
 * int32[n][6]; out u8[n][256] (row stride 16).                                              */
int  lh264_k_mc_luma(lh264_ctx* ctx, const uint8_t* ref, int w, int h, const int32_t* blocks, int n, uint8_t* out);
/* = McChroma_c (mc.cpp:368-378); blocks as above in chroma samples with 1/8-pel mv.        */
int  lh264_k_mc_chroma(lh264_ctx* ctx, const uint8_t* ref, int w, int h, const int32_t* blocks, int n, uint8_t* out);
/* = ExpandPictureLuma_c/Chroma_c (expand_pic.cpp:37-115): in w x h, out (w+2p) x (h+2p).    */
int  lh264_k_expand_plane(lh264_ctx* ctx, const uint8_t* in, int w, int h, int pad, uint8_t* out);

/* ==== recompression model ("Pied Piper" layer) ================================================
 * Replaces, for the compress direction, the per-macroblock modelling of the reference:
 *   lh264_model_mb / lh264_model_slice  <- DecodedMacroblock + the emission decisions that
 *                                          WelsDecodeSliceForNonRecoding takes from decoder state
 *                                          (decoded_macroblock.h:4-104, decode_slice.cpp:2174-2474)
 *   lh264_model_submit()                <- WelsDecodeSlice's MB loop for every slice of a picture
 *                                          (decode_slice.cpp:2959-3171): FreqImage neighbour
 *                                          contexts (decoded_macroblock.h:106-191, macroblock_model.cpp:9-44),
 *                                          prior selection (macroblock_model.cpp:375-645), the
 *                                          binarisations (compression_stream.h:441-591), DynProb
 *                                          adaptation (:87-115) and the VP8 bool coder (bitwriter.h:35-105),
 *                                          for a BATCH of pictures of independent streams.
 *   lh264_model_finish()                <- vpx_stop_encode + ArithmeticCodedOutput::flushToWriter
 *                                          (bitwriter.cpp:28-38, compression_stream.cpp:211-234): the
 *                                          bytes of one `x.pip.<tag>` side stream.
 * Model state (priors, previous-frame contexts, coder states) lives in HBM per model stream.       */

/* tag ids = the reference's billing enum (billing.h:7-58); file suffix of the side stream          */
enum {
  LH264_TAG_SKIP = 1, LH264_TAG_SKIP_END = 2, LH264_TAG_CBPL = 4, LH264_TAG_QPL = 6, LH264_TAG_MB_TYPE = 7,
  LH264_TAG_TRANSFORM_8x8 = 8, LH264_TAG_REF = 9, LH264_TAG_8x8 = 10, LH264_TAG_16x16 = 11,
  LH264_TAG_PRED_MODE = 13, LH264_TAG_SUB_MB = 14, LH264_TAG_MVX = 15, LH264_TAG_MVY = 16,
  LH264_TAG_LDC = 17, LH264_TAG_CRDC = 18,
  LH264_TAG_LAC_0_EOB = 19, LH264_TAG_LAC_0_BITMASK = 20, LH264_TAG_LAC_0_EXP = 21, LH264_TAG_LAC_0_RES = 22, LH264_TAG_LAC_0_SIGN = 23,
  LH264_TAG_LAC_N_BITMASK = 25, LH264_TAG_LAC_N_EXP = 26, LH264_TAG_LAC_N_RES = 27, LH264_TAG_LAC_N_SIGN = 28,
  LH264_TAG_CRAC_EOB = 29, LH264_TAG_CRAC_BITMASK = 30, LH264_TAG_CRAC_EXP = 31, LH264_TAG_CRAC_RES = 32, LH264_TAG_CRAC_SIGN = 33,
  LH264_TAG_PADBYTE = 69,
  LH264_NUM_TAGS = 70
};

#define LH264_MM_EMIT_SKIP_RUN 0x01  /* an mb_skip_run was read in front of this MB: code skip_run (decode_slice.cpp:2210-2217) */
#define LH264_MM_STOP_IN_SKIP  0x02  /* last MB of a skip run: code the stop bit (:2223-2226)                                  */
#define LH264_MM_WRITE_BLOCK   0x04  /* a coded macroblock (not skipped): code the whole record (:2244-2434)                     */
#define LH264_MM_STOP_BIT      0x08  /* value of the stop bit: the slice data ends right after this MB (:2200-2201)            */
#define LH264_MM_T8_CODED      0x10  /* transform_size_8x8_flag is part of the MB syntax (needParseTransformSize8x8)            */
#define LH264_MM_T8            0x20  /* its value                                                                               */
#define LH264_MM_LASTDQP_SHIFT 6     /* bits 6-7: sign class of the last non-zero luma QP delta of the slice before this MB
                                        (0 negative, 1 none yet, 2 positive): getQPLPrior, macroblock_model.cpp:441-444          */

/* One macroblock in slice-data order (skipped MBs included), 128 bytes.                                                       */
typedef struct lh264_model_mb {
  uint32_t mb_addr;          /* raster address in the picture                                                                   */
  uint16_t flags;            /* LH264_MM_*                                                                                      */
  uint16_t skip_run;         /* value coded when LH264_MM_EMIT_SKIP_RUN                                                         */
  uint16_t mb_in_slice;      /* macroblockIndexInSlice (decode_slice.cpp:3063)                                                  */
  uint8_t  type_code;        /* MacroblockModel::encodeMacroblockType: 0 I4x4 1 I16x16 2 I8x8 3 16x16 4 16x8 5 8x16 6 8x8
                                7 8x8ref0 8 I_PCM 11 skip (macroblock_model.cpp:647-679)                                        */
  uint8_t  cbp;              /* pCbp[mb] as the reference holds it: luma bits 0-3, chroma bits 4-5                              */
  uint8_t  luma_qp;          /* pLumaQp[mb]                                                                                     */
  uint8_t  chroma_mode;      /* pChromaPredMode[mb] (availability-resolved; retained from an earlier picture for inter MBs)    */
  uint8_t  i16_mode;         /* pIntraPredMode[mb][7] (likewise)                                                                */
  int8_t   qp_delta;         /* luma_qp minus the luma_qp of the previous coded MB of the slice (0 before the first one):
                                deltaLumaQp, decode_slice.cpp:2267-2274                                                         */
  uint8_t  sub_mb_type[4];   /* pSubMbType[mb][]: 1 8x8, 2 8x4, 4 4x8, 8 4x4 (retained values for I8x8)                          */
  int8_t   ref_idx[4];       /* rtd.iRefIdx[]: per partition (16x8 / 8x16: [0],[1]) or per 8x8 quadrant                           */
  uint8_t  pred_mode[16];    /* I4x4: low nibble = coded mode (iBestIntra4x4PredMode), high nibble = predicted mode;
                                I8x8: entries 0..3, low nibble only                                                            */
  uint8_t  pred_avail[16];   /* I4x4: (left << 2) | (top << 1) | top-left sample availability of each block (:2310-2316)       */
  uint32_t coded_mask;       /* as lh264_mb_rec (0 for I_PCM: the reference models no coefficients there)                       */
  uint32_t coef_off;
  int16_t  mvd[16][2];       /* motion vector differences as parsed, at the raster index of each partition's first 4x4 block   */
} lh264_model_mb;

typedef struct lh264_model_slice {
  int32_t  first_rec;        /* index of the slice's first lh264_model_mb in the picture's array                                */
  int32_t  n_recs;
  int32_t  frame_num;        /* FreqImage::updateFrame key (decode_slice.cpp:3034)                                              */
  uint8_t  slice_type;       /* eSliceType: 0 P, 2 I                                                                            */
  uint8_t  num_ref_idx_l0;   /* uiRefCount[0]                                                                                   */
  uint8_t  n_pad_bits;       /* slice trailing alignment bits coded into the PADBYTE tag (:3133-3147)                            */
  uint8_t  pad_bits;
  uint8_t  chroma_format_idc;
  uint8_t  reserved[3];
} lh264_model_slice;

typedef struct lh264_model_picture {
  int32_t  model_stream;     /* handle from lh264_model_open                                                                    */
  int32_t  mb_w, mb_h;
  int32_t  n_slices;
  int32_t  n_recs;
  int32_t  n_coef_blocks;
  const lh264_model_slice* slices;
  const lh264_model_mb*    recs;
  const int16_t*           coefs;   /* the picture's coefficient pool (same pool as lh264_picture_desc.coefs)                   */
} lh264_model_picture;

/* Opens the model state of one stream in HBM.  max_mbs bounds the picture size; with_8x8 != 0 also allocates the 8x8
 * coefficient priors (High profile streams).  Returns a handle >= 0 or an error.                                             */
int  lh264_model_open(lh264_ctx* ctx, int max_mbs, int with_8x8);
int  lh264_model_close(lh264_ctx* ctx, int model_stream);
/* Codes one picture per stream for a batch of independent streams (host buffers in; the buffers may be reused when the
 * call returns).  Runs on a CUDA stream of its own, concurrently with reconstruction work on the context stream;
 * lh264_sync() waits for both, lh264_join() orders the context stream after it.  A picture
 * whose `coefs` pointer equals that of a picture of the immediately preceding lh264_recon_submit() (the same pool, as the
 * host decoder hands it over) reads that batch's device copy instead of being copied again.  Output bytes accumulate on
 * the device and are read with lh264_model_read().                                                                          */
int  lh264_model_submit(lh264_ctx* ctx, const lh264_model_picture* pics, int n_pics);
/* Resident variant used by the roofline bench: stage a batch of model pictures in HBM once, code it (again) later.   */
typedef struct lh264_model_batch lh264_model_batch;
int  lh264_model_batch_upload(lh264_ctx* ctx, const lh264_model_picture* pics, int n_pics, lh264_model_batch** out);
int  lh264_model_batch_run(lh264_ctx* ctx, lh264_model_batch* batch);
void lh264_model_batch_free(lh264_ctx* ctx, lh264_model_batch* batch);
/* Ends every tag stream of a model stream (32 trailing zero decisions etc.) and makes all bytes available.                  */
int  lh264_model_finish(lh264_ctx* ctx, int model_stream);
/* Bytes produced so far for one tag: copies up to cap bytes starting at offset into dst (host), returns the total
 * number of bytes available for that tag (>= 0) or an error; a tag never used returns 0 and has no side file.               */
long long lh264_model_read(lh264_ctx* ctx, int model_stream, int tag, uint8_t* dst, size_t offset, size_t cap);
/* decisions coded so far by a model stream (for throughput reporting)                                                         */
unsigned long long lh264_model_decisions(lh264_ctx* ctx, int model_stream);

/* ---- decompress direction -----------------------------------------------------------------------------
 * The reference's round trip (h264dec x.pip out.264): every arithmetic-coded side stream is read back through the same
 * adaptive priors (decode4x4 / readMv / scanInt / scanUEGkInt, codec/decoder/core/inc/compression_stream.h:601-689,
 * src/decode_slice.cpp:2476-2956; bool decoder inc/bitreader.h:77-136) and each macroblock is re-serialised as CAVLC
 * (EncoderState, decode_slice.cpp:1442-1835 -> encoder/core/src/svc_set_mb_syn_cavlc.cpp) or, for slices of a CABAC picture
 * parameter set, as CABAC (decode_slice.cpp:1512-1539 -> svc_set_mb_syn_cabac.cpp; flags bit 2 of lh264_decode_slice).  Here one job = one stream: its
 * side streams and the parameters of its slices (from the slice headers kept in the default stream) go in, the slice_data()
 * bits of every slice come out.  Decoding is strictly serial inside a stream; a call runs all of its jobs concurrently.   */
typedef struct lh264_decode_slice {
  uint32_t first_mb;         /* first_mb_in_slice                                                                           */
  int32_t  frame_num;        /* FreqImage::updateFrame key (decode_slice.cpp:3034)                                          */
  uint16_t mb_w, mb_h;       /* picture size in macroblocks (active SPS)                                                    */
  uint8_t  slice_type;       /* 0 P, 2 I                                                                                    */
  uint8_t  num_ref_idx_l0;   /* active L0 reference count of the slice (PPS default for I slices)                           */
  uint8_t  slice_qp;         /* SliceQPY                                                                                    */
  uint8_t  flags;            /* bit0 PPS transform_8x8_mode_flag, bit1 PPS constrained_intra_pred_flag, bit2 PPS
                                entropy_coding_mode_flag (CABAC slice data), bits 3-4 cabac_init_idc                           */
  uint32_t header_bits;      /* bit position of slice_data() in the slice RBSP (trailing alignment bits, :3133-3147)        */
} lh264_decode_slice;

typedef struct lh264_decode_slice_out {
  uint64_t byte_offset;      /* where this slice's bits start in the job's `out` buffer (byte aligned)                     */
  uint32_t n_bits;           /* slice_data() length                                                                         */
  uint32_t n_mbs;            /* macroblocks in the slice, skipped ones included                                            */
  uint8_t  n_pad_bits;       /* alignment bits that follow the stop bit ...                                                */
  uint8_t  pad_bits;         /* ... and their value (PADBYTE stream)                                                       */
  uint16_t reserved;
  uint32_t reserved2;
} lh264_decode_slice_out;

typedef struct lh264_decode_job {
  int32_t  max_mbs;          /* largest picture of the stream, in macroblocks                                              */
  int32_t  with_8x8;         /* != 0: the stream may use the 8x8 transform                                                 */
  const uint8_t* tag_data[LH264_NUM_TAGS];   /* the `.pip.<tag>` side streams (NULL = no such file)                        */
  uint32_t tag_len[LH264_NUM_TAGS];
  const lh264_decode_slice* slices;
  int32_t  n_slices;
  int32_t  error;            /* out: 0, or 21 I_PCM (not representable, as in the reference), 22 `out` too small,
                                23 corrupt side streams, 24 picture larger than max_mbs, 25 context outside the prior tables */
  uint8_t* out;              /* host buffer receiving the slice_data() bits of all slices                                   */
  uint64_t out_capacity;
  lh264_decode_slice_out* slice_out;         /* [n_slices]                                                                 */
} lh264_decode_job;

/* Synchronous: uploads the jobs, runs one warp per job, downloads the bits.  Returns LH264_OK when every job decoded.    */
int  lh264_model_decode(lh264_ctx* ctx, lh264_decode_job* jobs, int n_jobs);

/* ---- compact transfer form of a picture's records -------------------------------------------------------------------------
 * lh264_mb_rec and lh264_model_mb are laid out for the kernels (128 bytes, one warp-wide load); most of their bytes repeat
 * (a 16x16 partition's motion vector sixteen times) or are unused by the macroblock's type.  Where the records cross PCIe at
 * the rate of a thousand 1080p streams, host memory bandwidth - parser stores + DMA reads + the decoded pictures coming back -
 * is what limits the path, so a host may hand them over in this form instead: 32 + 48 bytes per macroblock plus 16-byte
 * "extra" units for the types that need more, expanded on the device into exactly the records above (lh264_expand_mb /
 * lh264_expand_mm in losslessh264_b200/csrc/compact_records.h state the mapping, and its inverse for the host side).          */
typedef struct lh264_cmb {            /* compact lh264_mb_rec, 32 bytes                                                          */
  uint8_t  mb_type, flags, cbp, qp_y;
  uint8_t  qp_c[2], i16_mode, chroma_mode;
  uint16_t slice_idx;
  uint8_t  nxn_avail;
  uint8_t  kind;                      /* 0: mv below holds for all sixteen blocks (P16x16, P_Skip; intra without 4x4 modes: 0)
                                         1: extra = intra_mode[16]                      (1 unit;  I4x4, I8x8)
                                         2: extra = mv[16][2]                           (4 units; P16x8, P8x16)
                                         3: extra = mv[16][2], sub_mb_type[4], 12 pad   (5 units; P8x8)                          */
  uint32_t nz_mask, coded_mask, coef_off;
  int8_t   ref_idx[4];
  union { int16_t mv[2]; uint32_t extra_off; } u;   /* kind != 0: first unit of the extra data in the picture's mb extra array  */
} lh264_cmb;

typedef struct lh264_cmm {            /* compact lh264_model_mb, 48 bytes                                                        */
  uint32_t mb_addr;
  uint16_t flags, skip_run, mb_in_slice;
  uint8_t  type_code, cbp;
  uint8_t  luma_qp, chroma_mode, i16_mode;
  int8_t   qp_delta;
  uint8_t  sub_mb_type[4];
  int8_t   ref_idx[4];
  uint32_t coded_mask, coef_off;
  uint8_t  kind;                      /* 0: mvd below: [0] the first partition's, [1] the second's (16x8: block 8, 8x16: block 2)
                                         1: extra = pred_mode[16], pred_avail[16]      (2 units; I4x4, I8x8)
                                         2: extra = mvd[16][2]                          (4 units; 8x8 types)                     */
  uint8_t  reserved[3];
  union { int16_t mvd[2][2]; uint32_t extra_off; } u;
  uint32_t reserved2;
} lh264_cmm;

/* One picture in compact form: reconstruction and model hand-off together (they share the coefficient pool).                  */
typedef struct lh264_compact_picture {
  int32_t  stream;                    /* lh264_stream_open handle (ignored without LH264_COMPACT_RECON)                          */
  int32_t  model_stream;              /* lh264_model_open handle  (ignored without LH264_COMPACT_MODEL)                          */
  int32_t  cur_slot, mb_w, mb_h;
  int32_t  n_slices, n_model_slices, n_recs, n_coef_blocks;
  int32_t  n_mb_extra, n_mm_extra;    /* 16-byte units                                                                           */
  int32_t  reserved;
  const lh264_slice_rec*   slices;
  const lh264_model_slice* model_slices;
  const lh264_cmb*  cmbs;             /* mb_w * mb_h, raster order                                                               */
  const lh264_cmm*  cmms;             /* n_recs, slice-data order                                                                */
  const uint8_t*    mb_extra;
  const uint8_t*    mm_extra;
  const int16_t*    coefs;
} lh264_compact_picture;
#define LH264_COMPACT_RECON 1
#define LH264_COMPACT_MODEL 2
/* = lh264_recon_submit (flags & LH264_COMPACT_RECON) and lh264_model_submit (flags & LH264_COMPACT_MODEL) for a batch of pictures
 * handed over in compact form: the records are expanded on the device (k_expand_mb / k_expand_mm) in front of the same kernels.
 * Host buffers may be reused after lh264_inputs_consumed().                                                                   */
int  lh264_submit_compact(lh264_ctx* ctx, const lh264_compact_picture* pics, int n_pics, int flags);

#ifdef __cplusplus
}
#endif
#endif /* LH264_B200_H_ */
