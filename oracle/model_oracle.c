/*
 * oracle/model_oracle.c — TEST INFRASTRUCTURE ONLY.  Never linked, imported or executed by the
 * product (losslessh264_b200/); only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may use it, and only as the checker / the thing timed as CPU baseline.
 *
 * A plain-C, single-threaded, strictly sequential CPU restatement of the reference's "Pied Piper"
 * recompression model in the compress direction: for every macroblock in slice-data order it walks
 * the same symbols in the same order as the reference, selects the same adaptive priors from the same
 * neighbour contexts, and drives one VP8 bool coder per tag.  Input = the model hand-off records of
 * include/lh264_b200.h (lh264_model_picture); output = the bytes of each `x.pip.<tag>` side stream.
 *
 * Parity is PINNED: tests/test_model_oracle.py compares, tag by tag, the bytes produced here (from the
 * host decoder's records) with the `.pip.<tag>` files written by the compiled, unmodified reference
 * (oracle/_ref/h264dec) — SHA1s of those files are committed in tests/golden/pip_tag_sha1.json with
 * the script that made them.
 *
 * Reference code followed (paths relative to the reference root, codec/decoder/core/):
 *   DynProb, Branch<n>, *IntPrior, emitBits/emitBitsZeroToPow2Inclusive/emitUnary, emitInt, emitUEGkInt
 *                                              inc/compression_stream.h:87-244, :407-487, :523-591
 *   vpx_write / vpx_start_encode / vpx_stop_encode     inc/bitwriter.h:35-105, src/bitwriter.cpp:17-38
 *   FreqImage, Neighbors::init                  inc/decoded_macroblock.h:106-191, src/macroblock_model.cpp:9-44
 *   prior selectors                             src/macroblock_model.cpp:375-645
 *   per-MB emission order                       src/decode_slice.cpp:2174-2474 (WelsDecodeSliceForNonRecoding)
 *   encode4x4<16|64>, writeMv                   src/decode_slice.cpp:2060-2094, :2128-2135
 *   slice level (updateFrame, pad bits)         src/decode_slice.cpp:3032-3046, :3133-3147
 * The 9.4 GB dense prior table of the reference (inc/macroblock_model.h:36-136) is replaced by
 * zero-initialised arrays over the index ranges that can occur (a fresh DynProb is stored as all
 * zero bits: prob is kept XOR 0x80) — observationally identical.
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "../include/lh264_b200.h"

/* ---- DynProb (compression_stream.h:87-115) ------------------------------------------------------ */
typedef struct { uint8_t p; uint16_t c0, c1; } dynprob;          /* p = prob ^ 0x80 */
static inline int dp_prob (const dynprob* d) { return d->p ^ 0x80; }
static inline void dp_update (dynprob* d, int bit) {
  unsigned c0 = d->c0, c1 = d->c1;
  if (bit) c1++; else c0++;
  d->p = (uint8_t) (((256 * (c0 + 1)) / (c0 + c1 + 2)) ^ 0x80);
  if (c0 + c1 > 512) { c0 = (c0 + 1) >> 1; c1 = (c1 + 1) >> 1; }
  d->c0 = (uint16_t)c0; d->c1 = (uint16_t)c1;
}

/* ---- VP8 bool coder (bitwriter.h:35-105) --------------------------------------------------------- */
typedef struct { unsigned lowvalue, range; int count; unsigned pos; uint8_t* buf; size_t cap; int used; uint64_t decisions; } writer;

static unsigned norm_shift (unsigned range) {        /* vpx_norm[range]: left shifts to bring range into [128, 255] */
  unsigned s = 0;
  if (range == 0) return 0;
  while ((range << s) < 128) s++;
  return s;
}
static void w_start (writer* w) {                     /* vpx_start_encode, bitwriter.cpp:17-26 */
  w->lowvalue = 0; w->range = 255; w->count = -24; w->pos = 0; w->used = 1;
  if (!w->buf) { w->cap = 4096; w->buf = (uint8_t*)malloc (w->cap); }
}
static void w_write (writer* w, int bit, int probability) {
  if (w->pos + 16 > w->cap) { w->cap *= 2; w->buf = (uint8_t*)realloc (w->buf, w->cap); }
  unsigned split = 1 + (((w->range - 1) * (unsigned)probability) >> 8);
  unsigned range = split, lowvalue = w->lowvalue;
  int count = w->count;
  if (bit) { lowvalue += split; range = w->range - split; }
  unsigned shift = norm_shift (range);
  range <<= shift;
  count += (int)shift;
  if (count >= 0) {
    int offset = (int)shift - count;
    if ((lowvalue << (offset - 1)) & 0x80000000u) {
      int x = (int)w->pos - 1;
      while (x >= 0 && w->buf[x] == 0xff) { w->buf[x] = 0; x--; }
      if (x >= 0) w->buf[x] += 1;
    }
    w->buf[w->pos++] = (uint8_t) (lowvalue >> (24 - offset));
    lowvalue <<= offset;
    shift = (unsigned)count;
    lowvalue &= 0xffffff;
    count -= 8;
  }
  lowvalue <<= shift;
  w->count = count; w->lowvalue = lowvalue; w->range = range;
  w->decisions++;
}
static void w_stop (writer* w) {                      /* vpx_stop_encode, bitwriter.cpp:28-38 */
  for (int i = 0; i < 32; i++) w_write (w, 0, 128);
  w->decisions -= 32;
  if ((w->buf[w->pos - 1] & 0xe0) == 0xc0) w->buf[w->pos++] = 0;
}

/* ---- prior shapes (compression_stream.h:172-244, macroblock_model.h:40-53) --------------------- */
typedef struct { dynprob exponent[3], mantissa[4], zero; } nz_prior;                 /* UnsignedIntPrior<3,4>      */
typedef struct { dynprob exponent[3], mantissa[4], zero, sign; } dc_prior;           /* IntPrior<3,4>              */
typedef struct { dynprob sign, first[4], exponent[2], mantissa[4], zero2, zero; } ac_prior; /* UEGkIntPrior<14,4,2,4,0> */
typedef struct { dynprob sign, first[4], exponent[3], mantissa[4], zero2, zero; } mv_prior; /* UEGkIntPrior<9,4,3,4,3>  */

/* ---- per-position record of the two-frame context image (decoded_macroblock.h:4-104) ---------- */
typedef struct {
  int16_t lumaAC[256], chromaAC[128];
  uint8_t initialized, type_code, cbpC, cbpL, num_ref, chroma_mode, i16_mode, isSkipped;
  uint16_t cachedSkips;
} dmb;

#define N_TYPE 16
typedef struct mo_stream {
  /* FreqImage */
  dmb* frame[2]; size_t frame_len; int width, height, cur, lastFrameId, priorValid;
  /* priors; dimension comments are the reference's */
  dynprob (*mbTypePriors)[2][15];                              /* [32][2][15]                             */
  mv_prior (*mvd)[16];                                         /* [200 raw mb type][16]                   */
  dynprob chromaI8x8ModePriors[8][8];
  dc_prior (*lumaDC)[5][N_TYPE];                               /* [16][5][16]                             */
  dc_prior (*chromaDC)[5][N_TYPE];                             /* [8][5][16]                              */
  nz_prior* nonzeros;                                          /* [5][16][3][3][3][3]                     */
  nz_prior* nonzeros8x8;
  ac_prior* ac;                                                /* [5][16][3][16] x [5][5][5] (left/above fixed at 2) */
  ac_prior* ac8x8;                                             /* [5][16][1][64] x [5][5][5]              */
  dynprob* skipRun;                                            /* [512][16][511]                          */
  dynprob mbQPL[2][3][128];
  dynprob (*subMb)[255];                                       /* [16][255]                               */
  dynprob (*numRef)[N_TYPE][15];                               /* [16][16][15]                            */
  dynprob CbpC[4][N_TYPE][3];
  dynprob CbpL[16][N_TYPE][15];
  dynprob stopBit[2048];
  dynprob t8[N_TYPE][128];
  dynprob predMode[N_TYPE][8][9][15];
  dynprob TEST_PROB;                                           /* ArithmeticCodedOutput::TEST_PROB, shared by all tags */
  writer tags[LH264_NUM_TAGS];
  int error;
} mo_stream;

static writer* tagw (mo_stream* s, int tag) {         /* CompressionStream::tag(), compression_stream.h:510-519 */
  writer* w = &s->tags[tag];
  if (!w->used) w_start (w);
  return w;
}
static int mo_trace = -1;                                               /* MO_TRACE=1: every decision as "tag prob bit" on stderr */
static void emit_bit (mo_stream* s, int tag, int bit, dynprob* p) {   /* ArithmeticCodedOutput::emitBit :407-435 */
  if (mo_trace < 0) mo_trace = getenv ("MO_TRACE") != NULL;
  if (mo_trace) fprintf (stderr, "%d %d %d\n", tag, (int)dp_prob (p), bit);
  w_write (tagw (s, tag), bit, dp_prob (p));
  dp_update (p, bit);
}
static void emit_raw (mo_stream* s, int tag, unsigned data, int nbits) { /* emitBits(data, nBits) :441-448 */
  for (int i = 0; i < nbits; i++) emit_bit (s, tag, (data >> (nbits - 1 - i)) & 1, &s->TEST_PROB);
}
/* emitBits(data, Branch<n>) :477-487 with the pre-order tree layout of Branch<n>::selectBranch :131-139 */
static void emit_branch (mo_stream* s, int tag, unsigned data, int nbits, dynprob* tree) {
  for (int n = nbits; n >= 1; n--) {
    const int bit = (data >> (n - 1)) & 1;
    emit_bit (s, tag, bit, &tree[0]);
    const int children = (1 << (n - 1)) - 1;
    tree += bit ? children + 1 : 1;
    data &= (1u << (n - 1)) - 1;
  }
}
/* emitBitsZeroToPow2Inclusive<n> :455-463 */
static void emit_zero_to_pow2 (mo_stream* s, int tag, unsigned data, int nbits, dynprob* priors, unsigned preferred) {
  emit_bit (s, tag, data != preferred, &priors[0]);
  if (data != preferred) emit_branch (s, tag, data > preferred ? data - 1 : data, nbits, priors + 1);
}
/* emitUnary :465-474 with UnaryIntPrior<N>::at :175-177 */
static void emit_unary (mo_stream* s, int tag, int data, dynprob* pri, int n, int early_termination) {
  for (int i = 0; i < data; i++) {
    emit_bit (s, tag, 1, &pri[i < n - 1 ? i : n - 1]);
    if (i == early_termination - 1) return;
  }
  emit_bit (s, tag, 0, &pri[data < n - 1 ? data : n - 1]);
}
/* CompressionStream::emitInt :523-572 */
static void emit_int (mo_stream* s, int data, dynprob* exponent, int n_exp, dynprob* mantissa, int n_mant, int order,
                      dynprob* zero, dynprob* sign, int tagExp, int tagMant, int tagZero, int tagSign) {
  if (zero) { emit_bit (s, tagZero, data == 0, zero); if (data == 0) return; }
  if (sign) { emit_bit (s, tagSign, data > 0, sign); if (data < 0) data = -data; }
  data--;
  int log2 = 0;
  const int data_high = 1 + (data >> order);
  while ((2 << log2) <= data_high) log2++;
  emit_unary (s, tagExp, log2, exponent, n_exp, -1);
  int lo = 0, hi = n_mant;
  for (int k = 0; k < log2 + order; k++) {
    const int bit = k < log2 ? (data_high >> (log2 - 1 - k)) & 1 : (data >> (order - 1 - (k - log2))) & 1;
    if (hi > lo) {
      const int mid = (hi + lo) / 2;
      emit_bit (s, tagMant, bit, &mantissa[mid]);
      if (bit) lo = mid + 1; else hi = mid;
    } else emit_bit (s, tagMant, bit, &s->TEST_PROB);
  }
}
/* CompressionStream::emitUEGkInt :575-591 (Prior = UEGkIntPrior<threshold, 4, n_exp, 4, order>) */
static void emit_uegk (mo_stream* s, int data, int threshold, dynprob* zero, dynprob* sign, dynprob* first,
                       dynprob* exponent, int n_exp, dynprob* mantissa, int order, dynprob* zero2,
                       int tagExp, int tagMant, int tagZero, int tagSign) {
  emit_bit (s, tagZero, data == 0, zero);
  if (data == 0) return;
  emit_bit (s, tagSign, data < 0, sign);
  if (data < 0) data = -data;
  emit_unary (s, tagMant, data - 1, first, 4, threshold);
  if (data - 1 >= threshold)
    emit_int (s, data - 1 - threshold, exponent, n_exp, mantissa, 4, order, zero2, NULL, tagExp, tagMant, tagZero, tagSign);
}

/* ---- FreqImage (decoded_macroblock.h:106-191) --------------------------------------------------- */
static void fi_resize (mo_stream* s, int w, int h) {  /* std::vector::resize semantics, decode_slice.cpp:3043-3044 */
  const size_t n = (size_t)w * h;
  for (int k = 0; k < 2; k++) {
    dmb* nf = (dmb*)calloc (n ? n : 1, sizeof (dmb));
    if (s->frame[k]) { memcpy (nf, s->frame[k], sizeof (dmb) * (n < s->frame_len ? n : s->frame_len)); free (s->frame[k]); }
    s->frame[k] = nf;
  }
  s->frame_len = n; s->width = w; s->height = h;
}
static void fi_update_frame (mo_stream* s, int frame) {   /* FreqImage::updateFrame :120-165 */
  if (frame != s->lastFrameId) { s->cur ^= 1; s->lastFrameId = frame; }
  dmb* f = s->frame[1 - s->cur];
  unsigned contiguous = 0;
  for (size_t i = 0; i < s->frame_len; i++) {
    int zeroed = 1;
    for (int j = 0; j < 256 && zeroed; j++) if (f[i].lumaAC[j]) zeroed = 0;
    for (int j = 0; j < 128 && zeroed; j++) if (f[i].chromaAC[j]) zeroed = 0;
    if (zeroed) { f[i].isSkipped = 1; contiguous = (contiguous + 1) & 0xffff; }
    else {
      f[i].isSkipped = 0;
      for (unsigned j = 0; j < contiguous; j++) f[i - j].cachedSkips = (uint16_t)contiguous;
      contiguous = 0;
    }
  }
}

static int count16 (const int16_t* ac) { int n = 0; for (int i = 0; i < 16; i++) n += ac[i] != 0; return n; }
static int count64 (const int16_t* ac) { int n = 0; for (int i = 0; i < 64; i++) n += ac[i] != 0; return n; }
static const int16_t* get_ac (const dmb* m, int color, int idx) { return (color == 0 ? m->lumaAC : m->chromaAC) + 16 * idx; }
static inline int min_i (int a, int b) { return a < b ? a : b; }
static inline int clamp04 (int v) { return v < 0 ? 0 : (v > 4 ? 4 : v); }

typedef struct { const dmb* left; const dmb* above; const dmb* past; const dmb* mb; int slice_type, type_code; } mbctx;

/* getNonzerosPrior4x4, macroblock_model.cpp:478-527 */
static nz_prior* nonzeros_prior_4x4 (mo_stream* s, const mbctx* c, int color, int idx) {
  int past = 0, left = 0, above = 0;
  if (c->past) past = count16 (get_ac (c->past, color, idx));
  if (color == 0) {
    if ((idx & 3) == 0) { if (c->left) left = count16 (get_ac (c->left, color, idx + 3)); }
    else left = count16 (get_ac (c->mb, color, idx - 1));
    if (idx < 4) { if (c->above) above = count16 (get_ac (c->above, color, idx + 12)); }
    else above = count16 (get_ac (c->mb, color, idx - 4));
  } else {
    if ((idx & 1) == 0) { if (c->left) left = count16 (get_ac (c->left, color, idx + 1)); }
    else left = count16 (get_ac (c->mb, color, idx - 1));
    if ((idx & 2) == 0) { if (c->above) above = count16 (get_ac (c->above, color, idx + 2)); }
    else above = count16 (get_ac (c->mb, color, idx - 2));
  }
  return &s->nonzeros[((((c->slice_type * N_TYPE + c->type_code) * 3 + color) * 3 + min_i (2, past)) * 3 + min_i (2, left)) * 3 + min_i (2, above)];
}
/* getNonzerosPrior8x8, macroblock_model.cpp:529-560 */
static nz_prior* nonzeros_prior_8x8 (mo_stream* s, const mbctx* c, int idx) {
  int past = 0, left = 0, above = 0;
  if (c->past) past = count64 (get_ac (c->past, 0, idx));
  const int q = idx / 4;
  if ((q & 1) == 0) { if (c->left) left = count64 (get_ac (c->left, 0, (q + 1) * 4)); }
  else left = count64 (get_ac (c->mb, 0, (q - 1) * 4));
  if ((q & 2) == 0) { if (c->above) above = count64 (get_ac (c->above, 0, (q + 2) * 4)); }
  else above = count64 (get_ac (c->mb, 0, (q - 2) * 4));
  return &s->nonzeros8x8[((((c->slice_type * N_TYPE + c->type_code) * 3 + 0) * 3 + min_i (2, past)) * 3 + min_i (2, left)) * 3 + min_i (2, above)];
}

static const uint8_t kzz16[16] = { 0, 1, 4, 8, 5, 2, 3, 6, 9, 12, 13, 10, 7, 11, 14, 15 };   /* decode_slice.cpp:2035-2040 */
static const uint8_t kzz64[64] = {                                                            /* decode_slice.cpp:2041-2050 */
  0, 1, 5, 6, 14, 15, 27, 28, 2, 4, 7, 13, 16, 26, 29, 42, 3, 8, 12, 17, 25, 30, 41, 43, 9, 11, 18, 24, 31, 40, 44, 53,
  10, 19, 23, 32, 39, 45, 52, 54, 20, 22, 33, 38, 46, 51, 55, 60, 21, 34, 37, 47, 50, 56, 59, 61, 35, 36, 48, 49, 57, 58, 62, 63 };

/* encode4x4<n>, decode_slice.cpp:2060-2094, with getACPrior4x4/8x8 (macroblock_model.cpp:562-604): the neighbour
 * terms of the AC prior index are constant 2 because priorCoef only ever fills `past` (:196-262). */
static void encode_block (mo_stream* s, const mbctx* c, const int16_t* ac, int n, int index, int emit_dc, int color) {
  const uint8_t* zz = n == 16 ? kzz16 : kzz64;
  int nonzeros = 0;
  for (int i = n - 1; i >= (emit_dc ? 0 : 1); --i) if (ac[zz[i]] != 0) nonzeros++;
  nz_prior* np = n == 16 ? nonzeros_prior_4x4 (s, c, color, index) : nonzeros_prior_8x8 (s, c, index);
  const int tEob = color ? LH264_TAG_CRAC_EOB : LH264_TAG_LAC_0_EOB;
  emit_int (s, nonzeros, np->exponent, 3, np->mantissa, 4, 0, &np->zero, NULL, tEob, tEob, tEob, tEob);
  int nonzeros_left = nonzeros, emitted = 0, prev = 0, prev2 = 0;
  for (int i = 0; i < n; i++) {
    if (i == 0 && !emit_dc) continue;
    if (nonzeros_left == 0) continue;
    ac_prior* base = n == 16 ? s->ac : s->ac8x8;
    const size_t outer = n == 16 ? (size_t) (((c->slice_type * N_TYPE + c->type_code) * 3 + color) * 16 + emitted)
                                 : (size_t) (((c->slice_type * N_TYPE + c->type_code) * 1 + 0) * 64 + emitted);
    ac_prior* p = &base[(outer * 5 + min_i (4, nonzeros_left)) * 25 + clamp04 (prev + 2) * 5 + clamp04 (prev2 + 2)];
    const int coefficient = ac[zz[i]];
    const int first = i == 0 && !color;
    /* the reference touches (and thereby creates) the EXP tag stream for billing even when no bit goes there (:2081-2083) */
    tagw (s, color ? LH264_TAG_CRAC_EXP : first ? LH264_TAG_LAC_0_EXP : LH264_TAG_LAC_N_EXP);
    emit_uegk (s, coefficient, 14, &p->zero, &p->sign, p->first, p->exponent, 2, p->mantissa, 0, &p->zero2,
               color ? LH264_TAG_CRAC_EXP : first ? LH264_TAG_LAC_0_EXP : LH264_TAG_LAC_N_EXP,
               color ? LH264_TAG_CRAC_RES : first ? LH264_TAG_LAC_0_RES : LH264_TAG_LAC_N_RES,
               color ? LH264_TAG_CRAC_BITMASK : first ? LH264_TAG_LAC_0_BITMASK : LH264_TAG_LAC_N_BITMASK,
               color ? LH264_TAG_CRAC_SIGN : first ? LH264_TAG_LAC_0_SIGN : LH264_TAG_LAC_N_SIGN);
    prev2 = prev; prev = coefficient; emitted++;
    if (coefficient != 0) nonzeros_left--;
  }
}

/* writeMv, decode_slice.cpp:2128-2135: the prior is indexed by the RAW mb type bitmask (macroblock_model.cpp:607) */
static void write_mv (mo_stream* s, const lh264_model_mb* r, int raw_type, int i) {
  mv_prior* p = &s->mvd[raw_type][i];
  for (int xy = 0; xy < 2; xy++) {
    const int t = xy ? LH264_TAG_MVY : LH264_TAG_MVX;
    emit_uegk (s, r->mvd[i][xy], 9, &p->zero, &p->sign, p->first, p->exponent, 3, p->mantissa, 3, &p->zero2, t, t, t, t);
  }
}

static unsigned swizzle_sign (int v) { return v >= 0 ? ((unsigned)v << 1) & 0xffff : ((((unsigned) (-v - 1)) << 1) | 1) & 0xffff; }
static const uint8_t kScan4[16] = { 0, 1, 4, 5, 2, 3, 6, 7, 8, 9, 12, 13, 10, 11, 14, 15 };   /* g_kuiScan4 */
static const int kRawType[12] = { 1, 2, 4, 8, 16, 32, 64, 128, 512, 1024, 0x4000, 256 };       /* wels_common_defs.h:264-276 */

static void expand_coefs (dmb* m, const lh264_model_mb* r, const int16_t* pool) {
  const int16_t* src = pool + (size_t)r->coef_off * 16;
  for (int b = 0; b < 24; b++)
    if (r->coded_mask & (1u << b)) {
      memcpy ((b < 16 ? m->lumaAC : m->chromaAC) + 16 * (b < 16 ? b : b - 16), src, 32);
      src += 16;
    }
}

/* one macroblock: WelsDecodeSliceForNonRecoding decode_slice.cpp:2174-2474 + the FreqImage store of
 * WelsDecodeSlice :3098-3109 */
static void model_mb (mo_stream* s, const lh264_model_slice* sl, const lh264_model_mb* r, const int16_t* pool) {
  const int mbx = (int) (r->mb_addr % (unsigned)s->width), mby = (int) (r->mb_addr / (unsigned)s->width);
  dmb* cur = s->frame[s->cur];
  dmb* last = s->frame[1 - s->cur];
  /* Neighbors::init macroblock_model.cpp:9-44 */
  const dmb* L = mbx > 0 && cur[r->mb_addr - 1].initialized ? &cur[r->mb_addr - 1] : NULL;
  const dmb* A = mby > 0 && cur[r->mb_addr - s->width].initialized ? &cur[r->mb_addr - s->width] : NULL;
  const dmb* P = s->priorValid && last[r->mb_addr].initialized ? &last[r->mb_addr] : NULL;
  dmb rtd;
  memset (&rtd, 0, sizeof (rtd));
  rtd.initialized = 1;
  const int stop_idx = r->mb_in_slice < 2048 ? r->mb_in_slice : 2047;

  if (r->flags & LH264_MM_EMIT_SKIP_RUN) {            /* :2210-2217; rtd.uiMbType is still 0 here => type code 11 */
    const int prior = P ? P->cachedSkips / 8 + (P->cachedSkips % 8 ? 1 : 0) : 0;
    if (prior >= 512) { s->error = 1; return; }       /* the reference's array_nd bounds assert */
    emit_branch (s, LH264_TAG_SKIP, r->skip_run, 9, &s->skipRun[((size_t)prior * 16 + 11) * 511]);
  }
  if (r->flags & LH264_MM_STOP_IN_SKIP)               /* :2223-2226 */
    emit_bit (s, LH264_TAG_SKIP_END, !! (r->flags & LH264_MM_STOP_BIT), &s->stopBit[stop_idx]);

  if (r->flags & LH264_MM_WRITE_BLOCK) {
    const int tc = r->type_code;
    const int st = sl->slice_type;
    rtd.type_code = (uint8_t)tc; rtd.cbpC = r->cbp >> 4; rtd.cbpL = r->cbp & 15;
    rtd.num_ref = sl->num_ref_idx_l0; rtd.chroma_mode = r->chroma_mode; rtd.i16_mode = r->i16_mode;
    expand_coefs (&rtd, r, pool);
    mbctx c = { L, A, P, &rtd, st, tc };

    emit_bit (s, LH264_TAG_SKIP_END, !! (r->flags & LH264_MM_STOP_BIT), &s->stopBit[stop_idx]);        /* :2252 */
    {                                                 /* getMacroblockTypePrior macroblock_model.cpp:415-438 */
      int prior = 15, prevType = 15;
      if (A) prior = A->type_code;
      if (L) prior = L->type_code;
      if (P) prevType = P->type_code;
      prior += prevType;
      emit_branch (s, LH264_TAG_MB_TYPE, (unsigned)tc, 4, s->mbTypePriors[prior][st == 0]);
    }
    if (sl->chroma_format_idc != 0)
      emit_branch (s, LH264_TAG_CBPL, rtd.cbpC, 2, s->CbpC[P ? P->cbpC : 0][tc]);
    emit_branch (s, LH264_TAG_CBPL, rtd.cbpL, 4, s->CbpL[P ? P->cbpL : 0][tc]);
    emit_zero_to_pow2 (s, LH264_TAG_QPL, swizzle_sign (r->qp_delta), 7,
                       s->mbQPL[r->mb_in_slice == 0][(r->flags >> LH264_MM_LASTDQP_SHIFT) & 3], 0);
    {
      const int pr = P ? P->num_ref : 0;
      if (pr >= 16) { s->error = 1; return; }
      emit_branch (s, LH264_TAG_REF, rtd.num_ref, 4, s->numRef[pr][tc]);
    }
    int refBits = 0;                                  /* ceil(log2(n)) :2285 */
    while ((1 << refBits) < rtd.num_ref) refBits++;
    {
      const int pc = P ? (P->chroma_mode >= 6 ? 6 : P->chroma_mode) : 7;
      emit_zero_to_pow2 (s, LH264_TAG_8x8, r->chroma_mode, 3, s->chromaI8x8ModePriors[pc], (unsigned)pc);
      const int pl = P ? (P->i16_mode >= 6 ? 6 : P->i16_mode) : 7;       /* uses the CHROMA table: macroblock_model.cpp:632 */
      emit_zero_to_pow2 (s, LH264_TAG_16x16, r->i16_mode, 3, s->chromaI8x8ModePriors[pl], (unsigned)pl);
    }
    if (tc == 0) {                                    /* Intra4x4 :2301-2319 */
      for (int i = 0; i < 16; i++)
        emit_branch (s, LH264_TAG_PRED_MODE, r->pred_mode[i] & 15, 4, s->predMode[tc][r->pred_avail[i] & 7][r->pred_mode[i] >> 4]);
    }
    if (tc == 2) {                                    /* Intra8x8 :2320-2335 */
      for (int i = 0; i < 4; i++) emit_branch (s, LH264_TAG_PRED_MODE, r->pred_mode[i] & 15, 4, s->predMode[tc][6][1]);
      for (int i = 0; i < 4; i++) emit_branch (s, LH264_TAG_SUB_MB, r->sub_mb_type[i], 8, s->subMb[tc]);
      for (int i = 0; i < 4; i++) emit_raw (s, LH264_TAG_REF, (uint8_t)r->ref_idx[i], refBits);
    } else if (tc == 6 || tc == 7) {                  /* P8x8 :2336-2374 */
      for (int i = 0; i < 4; i++) emit_branch (s, LH264_TAG_SUB_MB, r->sub_mb_type[i], 8, s->subMb[tc]);
      if (tc == 6) for (int i = 0; i < 4; i++) emit_raw (s, LH264_TAG_REF, (uint8_t)r->ref_idx[i], refBits);
      for (int i = 0; i < 4; i++) {
        switch (r->sub_mb_type[i]) {
        case 1: write_mv (s, r, kRawType[tc], kScan4[i << 2]); break;
        case 2: for (int j = 0; j < 2; j++) write_mv (s, r, kRawType[tc], kScan4[ (i << 2) + (j << 1)]); break;
        case 4: for (int j = 0; j < 2; j++) write_mv (s, r, kRawType[tc], kScan4[ (i << 2) + j]); break;
        case 8: for (int j = 0; j < 4; j++) write_mv (s, r, kRawType[tc], kScan4[ (i << 2) + j]); break;
        default: s->error = 1; return;
        }
      }
    } else if (tc == 4 || tc == 5) {                  /* 16x8 / 8x16 :2375-2385 */
      for (int i = 0; i < 2; i++) emit_raw (s, LH264_TAG_REF, (unsigned) (int)r->ref_idx[i], refBits);
      for (int i = 0; i < 2; i++) write_mv (s, r, kRawType[tc], tc == 4 ? i * 8 : i * 2);
    } else if (tc == 3) {                             /* 16x16 :2386-2389 */
      emit_raw (s, LH264_TAG_REF, (unsigned) (int)r->ref_idx[0], refBits);
      write_mv (s, r, kRawType[tc], 0);
    }
    if (r->flags & LH264_MM_T8_CODED)                 /* :2390-2393 */
      emit_bit (s, LH264_TAG_TRANSFORM_8x8, !! (r->flags & LH264_MM_T8), &s->t8[tc][r->luma_qp]);
    int emitted_luma_dc = 0, emitted_chroma_dc = 0;
    if (tc == 1) {                                    /* :2396-2402 */
      emitted_luma_dc = 1;
      for (int i = 0; i < 16; i++) {
        dc_prior* p = &s->lumaDC[i][st][tc];
        emit_int (s, rtd.lumaAC[16 * i], p->exponent, 3, p->mantissa, 4, 0, &p->zero, &p->sign, LH264_TAG_LDC, LH264_TAG_LDC, LH264_TAG_LDC, LH264_TAG_LDC);
      }
    }
    if (rtd.cbpC == 1 || rtd.cbpC == 2) {             /* :2403-2409 */
      emitted_chroma_dc = 1;
      for (int i = 0; i < 8; i++) {
        dc_prior* p = &s->chromaDC[i][st][tc];
        emit_int (s, rtd.chromaAC[16 * i], p->exponent, 3, p->mantissa, 4, 0, &p->zero, &p->sign, LH264_TAG_CRDC, LH264_TAG_CRDC, LH264_TAG_CRDC, LH264_TAG_CRDC);
      }
    }
    const int t8 = !! (r->flags & LH264_MM_T8);
    for (int q = 0; q < 4; q++)                       /* :2410-2426 */
      if (rtd.cbpL & (1 << q)) {
        if (t8) encode_block (s, &c, &rtd.lumaAC[q * 64], 64, q * 4, !emitted_luma_dc, 0);
        else for (int k = 0; k < 4; k++) encode_block (s, &c, &rtd.lumaAC[ (q * 4 + k) * 16], 16, q * 4 + k, !emitted_luma_dc, 0);
      }
    if (rtd.cbpC == 2)                                /* :2427-2435 */
      for (int i = 0; i < 8; i++) encode_block (s, &c, &rtd.chromaAC[i * 16], 16, i, !emitted_chroma_dc, i < 4 ? 1 : 2);
    cur[r->mb_addr] = rtd;                            /* :3107 */
  } else {
    cur[r->mb_addr] = last[r->mb_addr];               /* :3105 (skipped: copy the co-located record of the other frame) */
  }
}

/* ---- public (ctypes) ------------------------------------------------------------------------------ */
mo_stream* mo_open (void) {
  mo_stream* s = (mo_stream*)calloc (1, sizeof (mo_stream));
  s->mbTypePriors = calloc (32, sizeof (*s->mbTypePriors));
  s->mvd = calloc (200, sizeof (*s->mvd));
  s->lumaDC = calloc (16, sizeof (*s->lumaDC));
  s->chromaDC = calloc (8, sizeof (*s->chromaDC));
  s->nonzeros = calloc (5 * N_TYPE * 3 * 27, sizeof (nz_prior));
  s->nonzeros8x8 = calloc (5 * N_TYPE * 3 * 27, sizeof (nz_prior));
  s->ac = calloc ((size_t)5 * N_TYPE * 3 * 16 * 125, sizeof (ac_prior));
  s->ac8x8 = calloc ((size_t)5 * N_TYPE * 64 * 125, sizeof (ac_prior));
  s->skipRun = calloc ((size_t)512 * 16 * 511, sizeof (dynprob));
  s->subMb = calloc (16, sizeof (*s->subMb));
  s->numRef = calloc (16, sizeof (*s->numRef));
  return s;
}
void mo_close (mo_stream* s) {
  if (!s) return;
  free (s->mbTypePriors); free (s->mvd); free (s->lumaDC); free (s->chromaDC); free (s->nonzeros); free (s->nonzeros8x8);
  free (s->ac); free (s->ac8x8); free (s->skipRun); free (s->subMb); free (s->numRef);
  free (s->frame[0]); free (s->frame[1]);
  for (int t = 0; t < LH264_NUM_TAGS; t++) free (s->tags[t].buf);
  free (s);
}
/* one picture = all of its slices; mirrors WelsDecodeSlice decode_slice.cpp:2959-3171 per slice */
int mo_picture (mo_stream* s, int mb_w, int mb_h, int n_slices, const lh264_model_slice* slices, const lh264_model_mb* recs, const int16_t* coefs) {
  for (int i = 0; i < n_slices; i++) {
    const lh264_model_slice* sl = &slices[i];
    s->priorValid = 1;                                /* :3033-3046 */
    fi_update_frame (s, sl->frame_num);
    if (s->width != mb_w || s->height != mb_h) { s->priorValid = 0; fi_resize (s, mb_w, mb_h); }
    for (int k = 0; k < sl->n_recs && !s->error; k++) model_mb (s, sl, &recs[sl->first_rec + k], coefs);
    if (sl->n_pad_bits) emit_raw (s, LH264_TAG_PADBYTE, sl->pad_bits, sl->n_pad_bits);   /* :3133-3147 */
    if (s->error) return -1;
  }
  return 0;
}
int mo_finish (mo_stream* s) {                        /* ArithmeticCodedOutput::flushToWriter compression_stream.cpp:211 */
  for (int t = 0; t < LH264_NUM_TAGS; t++) if (s->tags[t].used) w_stop (&s->tags[t]);
  return 0;
}
long mo_tag_bytes (mo_stream* s, int tag, const uint8_t** p) {
  if (tag < 0 || tag >= LH264_NUM_TAGS || !s->tags[tag].used) return -1;
  *p = s->tags[tag].buf;
  return (long)s->tags[tag].pos;
}
unsigned long long mo_decisions (mo_stream* s) {
  unsigned long long n = 0;
  for (int t = 0; t < LH264_NUM_TAGS; t++) n += s->tags[t].decisions;
  return n;
}
