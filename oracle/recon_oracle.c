/*
 * oracle/recon_oracle.c — TEST INFRASTRUCTURE ONLY.  Never linked, imported or executed by the
 * product (losslessh264_b200/); only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline
 * leg may use it, and only as the checker.
 *
 * A plain-C, single-threaded CPU restatement of the reference's macroblock reconstruction for
 * one picture, consuming the hand-off records of include/lh264_b200.h.  Parity is PINNED: the
 * output is checked (tests/test_oracle_golden.py) against pictures dumped from the compiled
 * reference itself (oracle/_ref/h264dec_dump, fixtures in tests/golden/ .lhfb files) and, through the
 * host decoder, against the SHA1 table of the reference's test/api/decoder_test.cpp:90-128.
 *
 * Every function cites the reference code it follows (paths relative to the reference root).
 * Differences in structure that do not change results:
 *   - dequantisation is applied here (the reference fuses it into the CAVLC/CABAC parse,
 *     codec/decoder/core/src/parse_mb_syn_cavlc.cpp:938-1005, :1088-1095);
 *   - reference pictures are read with coordinate clamping instead of from a 32-pixel
 *     replicated border (codec/common/src/expand_pic.cpp:37-115) — identical samples because
 *     BaseMC clamps the motion vector into the border (codec/decoder/core/src/rec_mb.cpp:251-252).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include "../include/lh264_b200.h"

typedef struct ora_picture {
  uint8_t* plane[3];   /* Y, Cb, Cr; tightly packed */
  int w, h;            /* luma size = mb_w*16, mb_h*16 */
} ora_picture;

static inline int clip3 (int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }
static inline uint8_t clip1 (int v) { return (uint8_t)clip3 (v, 0, 255); }
static inline int iabs (int v) { return v < 0 ? -v : v; }

/* ------------------------------------------------------------------------------------------------
 * Dequantisation.  g_kuiDequantCoeff[52][8] (codec/common/src/common_tables.cpp:208-235) is
 * LevelScale4x4[qp%6][class] << (qp/6), indexed by (raster position & 7);
 * g_kuiDequantCoeff8x8[52][64] (:237+) is 16*normAdjust8x8[qp%6][pos] without the qp/6 shift. */
static const uint8_t kScale4[6][3] = { {10, 16, 13}, {11, 18, 14}, {13, 20, 16}, {14, 23, 18}, {16, 25, 20}, {18, 29, 23} };
static inline int dq4 (int qp, int pos) {
  static const uint8_t cls[8] = {0, 2, 0, 2, 2, 1, 2, 1};
  return kScale4[qp % 6][cls[pos & 7]] << (qp / 6);
}
static const uint8_t kScale8[6][6] = { {20, 18, 32, 19, 25, 24}, {22, 19, 35, 21, 28, 26}, {26, 23, 42, 24, 33, 31},
  {28, 25, 45, 26, 35, 33}, {32, 28, 51, 30, 40, 38}, {36, 32, 58, 34, 46, 43} };
static inline int dq8 (int qp, int pos) {
  int i = pos >> 3, j = pos & 7, c;
  if ((i & 3) == 0 && (j & 3) == 0) c = 0;
  else if ((i & 1) && (j & 1)) c = 1;
  else if ((i & 3) == 2 && (j & 3) == 2) c = 2;
  else if (((i & 3) == 0 && (j & 1)) || ((i & 1) && (j & 3) == 0)) c = 3;
  else if (((i & 3) == 0 && (j & 3) == 2) || ((i & 3) == 2 && (j & 3) == 0)) c = 4;
  else c = 5;
  return 16 * kScale8[qp % 6][c];
}

/* Scaling lists (WelsCalcDeqCoeffScalingList, decode_slice.cpp:1231-1277): the reference's per-qp tables are
 * list[pos] * g_kuiDequantCoeff[q][pos & 7] (4x4) and list[pos] * g_kuiMatrixV[q % 6][pos] (8x8), filled for q < 51 only
 * (the q == 51 row keeps the zeros of the context's allocation).  `sc` = the slice record when LH264_SLF_SCALING is set. */
static inline int sl_dq4 (const lh264_slice_rec* sc, int list, int q, int pos) { return q < 51 ? (uint16_t) (sc->scale4[list][pos] * dq4 (q, pos)) : 0; }   /* uint16_t tables */
static inline int sl_dq8 (const lh264_slice_rec* sc, int list, int q, int pos) { return q < 51 ? sc->scale8[list][pos] * (dq8 (q % 6, pos) >> 4) : 0; }
/* CAVLC slices: WelsResidualBlockCavlc8x8 is handed the 4x4 property (decode_slice.cpp:3385, :3782), so its table index
 * (property - 6) is -6 / -3: pDequant_coeff8x8[-6] IS pDequant_coeff4x4[0] (adjacent pointer arrays, decoder_context.h:448-449),
 * read as a 64-wide table over pDequant_coeff_buffer4x4[6][52][16] and, past its end, pDequant_coeff_buffer8x8[6][52][64]. */
static inline int sl_dq8_cavlc (const lh264_slice_rec* sc, int intra, int qp, int pos) {
  const int n = (intra ? 0 : 3 * 52 * 16) + qp * 64 + pos;
  if (n < 6 * 52 * 16) return sl_dq4 (sc, n / (52 * 16), (n % (52 * 16)) / 16, n % 16);
  const int m = n - 6 * 52 * 16;
  return sl_dq8 (sc, m / (52 * 64), (m % (52 * 64)) / 64, m % 64);
}

/* parse_mb_syn_cavlc.cpp:989-993: pTCoeff[j] = level * dq[j & 7], stored as int16_t; with scaling lists
 * (level * table[j]) >> 4 (same lines). */
static void dequant_ac4x4 (int16_t* out, const int16_t* lev, int qp, int first, const lh264_slice_rec* sc, int list) {
  if (sc) { for (int j = first; j < 16; j++) out[j] = (int16_t) ((lev[j] * sl_dq4 (sc, list, qp, j)) >> 4); return; }
  for (int j = first; j < 16; j++) out[j] = (int16_t) (lev[j] * dq4 (qp, j));
}
/* parse_mb_syn_cavlc.cpp:1088-1095 */
static void dequant_8x8 (int16_t* out, const int16_t* lev, int qp, const lh264_slice_rec* sc, int intra) {
  for (int j = 0; j < 64; j++) {
    const int mul = !sc ? dq8 (qp, j) : ((sc->scaling_flags & LH264_SLF_SCALING_CAVLC) ? sl_dq8_cavlc (sc, intra, qp, j) : sl_dq8 (sc, intra ? 0 : 1, qp, j));
    int v = lev[j] * mul;
    out[j] = (int16_t) (qp >= 36 ? (v << (qp / 6 - 6)) : ((v + (1 << (5 - qp / 6))) >> (6 - qp / 6)));
  }
}

/* IdctResAddPred_c, codec/decoder/core/src/decode_mb_aux.cpp:42-77 (int16 row intermediates). */
static void idct4x4_add (uint8_t* p, int stride, const int16_t* c) {
  int16_t t[16];
  for (int i = 0; i < 4; i++) {
    const int16_t* r = c + 4 * i;
    int e0 = r[0] + r[2], e1 = r[0] - r[2], e2 = (r[1] >> 1) - r[3], e3 = r[1] + (r[3] >> 1);
    t[4 * i] = (int16_t) (e0 + e3); t[4 * i + 1] = (int16_t) (e1 + e2);
    t[4 * i + 2] = (int16_t) (e1 - e2); t[4 * i + 3] = (int16_t) (e0 - e3);
  }
  for (int i = 0; i < 4; i++) {
    int e0 = t[i] + t[i + 8], e1 = t[i] - t[i + 8], e2 = (t[i + 4] >> 1) - t[i + 12], e3 = t[i + 4] + (t[i + 12] >> 1);
    p[i] = clip1 (p[i] + ((32 + e0 + e3) >> 6));
    p[i + stride] = clip1 (p[i + stride] + ((32 + e1 + e2) >> 6));
    p[i + 2 * stride] = clip1 (p[i + 2 * stride] + ((32 + e1 - e2) >> 6));
    p[i + 3 * stride] = clip1 (p[i + 3 * stride] + ((32 + e0 - e3) >> 6));
  }
}

/* one 8-point pass of IdctResAddPred8x8_c (decode_mb_aux.cpp:79-167); everything int16. */
static void idct8_1d (const int16_t* in, int istride, int16_t* out, int ostride) {
  int16_t p[8], a[4], b[8];
  for (int j = 0; j < 8; j++) p[j] = in[j * istride];
  a[0] = (int16_t) (p[0] + p[4]); a[1] = (int16_t) (p[0] - p[4]);
  a[2] = (int16_t) (p[6] - (p[2] >> 1)); a[3] = (int16_t) (p[2] + (p[6] >> 1));
  b[0] = (int16_t) (a[0] + a[3]); b[2] = (int16_t) (a[1] - a[2]); b[4] = (int16_t) (a[1] + a[2]); b[6] = (int16_t) (a[0] - a[3]);
  a[0] = (int16_t) (-p[3] + p[5] - p[7] - (p[7] >> 1));
  a[1] = (int16_t) (p[1] + p[7] - p[3] - (p[3] >> 1));
  a[2] = (int16_t) (-p[1] + p[7] + p[5] + (p[5] >> 1));
  a[3] = (int16_t) (p[3] + p[5] + p[1] + (p[1] >> 1));
  b[1] = (int16_t) (a[0] + (a[3] >> 2)); b[3] = (int16_t) (a[1] + (a[2] >> 2));
  b[5] = (int16_t) (a[2] - (a[1] >> 2)); b[7] = (int16_t) (a[3] - (a[0] >> 2));
  out[0 * ostride] = (int16_t) (b[0] + b[7]); out[1 * ostride] = (int16_t) (b[2] - b[5]);
  out[2 * ostride] = (int16_t) (b[4] + b[3]); out[3 * ostride] = (int16_t) (b[6] + b[1]);
  out[4 * ostride] = (int16_t) (b[6] - b[1]); out[5 * ostride] = (int16_t) (b[4] - b[3]);
  out[6 * ostride] = (int16_t) (b[2] + b[5]); out[7 * ostride] = (int16_t) (b[0] - b[7]);
}
static void idct8x8_add (uint8_t* p, int stride, const int16_t* c) {
  int16_t t[64], r[64];
  for (int i = 0; i < 8; i++) idct8_1d (c + 8 * i, 1, t + 8 * i, 1);
  for (int i = 0; i < 8; i++) idct8_1d (t + i, 8, r + i, 8);
  for (int i = 0; i < 8; i++) for (int j = 0; j < 8; j++)
      p[i * stride + j] = clip1 (p[i * stride + j] + ((32 + r[8 * i + j]) >> 6));
}

/* WelsLumaDcDequantIdct, decode_slice.cpp:271-311.  dc[z] = level in slot 0 of block z (z order);
 * result written back (int16) to slot 0 of each block. */
static void luma_dc_dequant_idct (int16_t* blk /*[16][16]*/, int qp, const lh264_slice_rec* sc) {
  static const int xoff[4] = {0, 1, 4, 5}, yoff[4] = {0, 2, 8, 10};   /* block index = yoff[i] + xoff[j] */
  const int qmul = sc ? sl_dq4 (sc, 0, qp, 0) >> 4 : dq4 (qp, 0);      /* decode_slice.cpp:272 */
  int t[16];
  for (int i = 0; i < 4; i++) {
    int o = yoff[i];
    int a = blk[16 * (o + 0)], b = blk[16 * (o + 1)], c = blk[16 * (o + 4)], d = blk[16 * (o + 5)];
    int z0 = a + c, z1 = a - c, z2 = b - d, z3 = b + d;
    t[4 * i] = z0 + z3; t[4 * i + 1] = z1 + z2; t[4 * i + 2] = z1 - z2; t[4 * i + 3] = z0 - z3;
  }
  for (int i = 0; i < 4; i++) {
    int o = xoff[i];
    int z0 = t[i] + t[8 + i], z1 = t[i] - t[8 + i], z2 = t[4 + i] - t[12 + i], z3 = t[4 + i] + t[12 + i];
    blk[16 * (o + yoff[0])] = (int16_t) (((z0 + z3) * qmul + 2) >> 2);
    blk[16 * (o + yoff[1])] = (int16_t) (((z1 + z2) * qmul + 2) >> 2);
    blk[16 * (o + yoff[2])] = (int16_t) (((z1 - z2) * qmul + 2) >> 2);
    blk[16 * (o + yoff[3])] = (int16_t) (((z0 - z3) * qmul + 2) >> 2);
  }
}
/* WelsChromaDcIdct, decode_slice.cpp:375-396 (input already multiplied by dq[qp][0], int16). */
static void chroma_dc_idct (int16_t* blk /*[4][16]*/) {
  int a = blk[0], b = blk[16], c = blk[32], d = blk[48];
  int e = a - b; a += b; b = c - d; c += d;
  blk[0] = (int16_t) ((a + c) >> 1); blk[16] = (int16_t) ((e + b) >> 1);
  blk[32] = (int16_t) ((a - c) >> 1); blk[48] = (int16_t) ((e - b) >> 1);
}

/* ------------------------------------------------------------------------------------------------
 * Inter prediction: McLuma_c / McChroma_c, codec/common/src/mc.cpp:142-378. */
static inline int refpix (const uint8_t* p, int w, int h, int x, int y) {
  return p[(size_t)clip3 (y, 0, h - 1) * w + clip3 (x, 0, w - 1)];
}
static inline int tap6 (int a, int b, int c, int d, int e, int f) { return a + f - 5 * (b + e) + 20 * (c + d); }
static int luma_h (const uint8_t* p, int w, int h, int x, int y) {   /* McHorVer20_c */
  return clip1 ((tap6 (refpix (p, w, h, x - 2, y), refpix (p, w, h, x - 1, y), refpix (p, w, h, x, y), refpix (p, w, h, x + 1, y),
                       refpix (p, w, h, x + 2, y), refpix (p, w, h, x + 3, y)) + 16) >> 5);
}
static int luma_v_raw (const uint8_t* p, int w, int h, int x, int y) {
  return tap6 (refpix (p, w, h, x, y - 2), refpix (p, w, h, x, y - 1), refpix (p, w, h, x, y), refpix (p, w, h, x, y + 1),
               refpix (p, w, h, x, y + 2), refpix (p, w, h, x, y + 3));
}
static int luma_v (const uint8_t* p, int w, int h, int x, int y) { return clip1 ((luma_v_raw (p, w, h, x, y) + 16) >> 5); } /* McHorVer02_c */
static int luma_c (const uint8_t* p, int w, int h, int x, int y) {   /* McHorVer22_c: vertical first into int16 */
  int16_t t[6];
  for (int k = 0; k < 6; k++) t[k] = (int16_t)luma_v_raw (p, w, h, x - 2 + k, y);
  return clip1 ((tap6 (t[0], t[1], t[2], t[3], t[4], t[5]) + 512) >> 10);
}
static int luma_sample (const uint8_t* p, int w, int h, int x, int y, int fx, int fy) {
#define AVG(a, b) (((a) + (b) + 1) >> 1)
  switch (fx * 4 + fy) {   /* pWelsMcFunc[x][y], mc.cpp:337-343 */
  case 0: return refpix (p, w, h, x, y);
  case 1: return AVG (refpix (p, w, h, x, y), luma_v (p, w, h, x, y));
  case 2: return luma_v (p, w, h, x, y);
  case 3: return AVG (refpix (p, w, h, x, y + 1), luma_v (p, w, h, x, y));
  case 4: return AVG (refpix (p, w, h, x, y), luma_h (p, w, h, x, y));
  case 5: return AVG (luma_h (p, w, h, x, y), luma_v (p, w, h, x, y));
  case 6: return AVG (luma_v (p, w, h, x, y), luma_c (p, w, h, x, y));
  case 7: return AVG (luma_h (p, w, h, x, y + 1), luma_v (p, w, h, x, y));
  case 8: return luma_h (p, w, h, x, y);
  case 9: return AVG (luma_h (p, w, h, x, y), luma_c (p, w, h, x, y));
  case 10: return luma_c (p, w, h, x, y);
  case 11: return AVG (luma_h (p, w, h, x, y + 1), luma_c (p, w, h, x, y));
  case 12: return AVG (refpix (p, w, h, x + 1, y), luma_h (p, w, h, x, y));
  case 13: return AVG (luma_h (p, w, h, x, y), luma_v (p, w, h, x + 1, y));
  case 14: return AVG (luma_v (p, w, h, x + 1, y), luma_c (p, w, h, x, y));
  default: return AVG (luma_h (p, w, h, x, y + 1), luma_v (p, w, h, x + 1, y));
  }
#undef AVG
}
void ora_mc_luma (const uint8_t* ref, int w, int h, int x, int y, int mvx, int mvy, int bw, int bh, uint8_t* dst, int dstride) {
  int ix = x + (mvx >> 2), iy = y + (mvy >> 2);
  for (int j = 0; j < bh; j++) for (int i = 0; i < bw; i++)
      dst[j * dstride + i] = (uint8_t)luma_sample (ref, w, h, ix + i, iy + j, mvx & 3, mvy & 3);
}
void ora_mc_chroma (const uint8_t* ref, int w, int h, int x, int y, int mvx, int mvy, int bw, int bh, uint8_t* dst, int dstride) {
  int ix = x + (mvx >> 3), iy = y + (mvy >> 3), dx = mvx & 7, dy = mvy & 7;
  int A = (8 - dx) * (8 - dy), B = dx * (8 - dy), C = (8 - dx) * dy, D = dx * dy;
  for (int j = 0; j < bh; j++) for (int i = 0; i < bw; i++)
      dst[j * dstride + i] = (uint8_t) ((A * refpix (ref, w, h, ix + i, iy + j) + B * refpix (ref, w, h, ix + i + 1, iy + j) +
                                         C * refpix (ref, w, h, ix + i, iy + j + 1) + D * refpix (ref, w, h, ix + i + 1, iy + j + 1) + 32) >> 6);
}

/* BaseMC + WeightPrediction, rec_mb.cpp:247-341.  (x,y) = partition origin in the picture. */
static void base_mc (const ora_picture* cur, const ora_picture* ref, const lh264_slice_rec* sl, int refidx,
                     int x, int y, int bw, int bh, int mvx, int mvy) {
  int fx = clip3 ((x << 2) + mvx, (-32 + 2) * 4, (cur->w + 32 - 19) << 2);
  int fy = clip3 ((y << 2) + mvy, (-32 + 2) * 4, (cur->h + 32 - 19) << 2);
  const int cw = cur->w >> 1, ch = cur->h >> 1;
  uint8_t* dy = cur->plane[0] + (size_t)y * cur->w + x;
  ora_mc_luma (ref->plane[0], cur->w, cur->h, 0, 0, fx, fy, bw, bh, dy, cur->w);
  for (int p = 1; p < 3; p++) {
    uint8_t* dc = cur->plane[p] + (size_t) (y >> 1) * cw + (x >> 1);
    ora_mc_chroma (ref->plane[p], cw, ch, 0, 0, fx, fy, bw >> 1, bh >> 1, dc, cw);
  }
  if (sl->weighted_pred) {
    int ld = sl->luma_log2_denom, wgt = sl->luma_weight[refidx], off = sl->luma_offset[refidx];
    for (int j = 0; j < bh; j++) for (int i = 0; i < bw; i++) {
        int v = dy[j * cur->w + i];
        dy[j * cur->w + i] = clip1 (ld >= 1 ? (((v * wgt + (1 << (ld - 1))) >> ld) + off) : (v * wgt + off));
      }
    /* the reference's chroma loops run over (bw>>2) x (bh>>2) samples only (rec_mb.cpp:309-311) */
    int cbw = bw >> 2, cbh = bh >> 2;
    ld = sl->chroma_log2_denom;
    for (int p = 1; p < 3; p++) {
      uint8_t* dc = cur->plane[p] + (size_t) (y >> 1) * cw + (x >> 1);
      wgt = sl->chroma_weight[refidx][p - 1]; off = sl->chroma_offset[refidx][p - 1];
      for (int j = 0; j < cbh; j++) for (int i = 0; i < cbw; i++) {
          int v = dc[j * cw + i];
          dc[j * cw + i] = clip1 (ld >= 1 ? (((v * wgt + (1 << (ld - 1))) >> ld) + off) : (v * wgt + off));
        }
    }
  }
}

/* GetInterPred, rec_mb.cpp:344-545 */
static void inter_pred (const ora_picture* cur, const ora_picture* slots, int n_slots, const lh264_slice_rec* sl,
                        const lh264_mb_rec* m, int mbx, int mby) {
  const int x0 = mbx * 16, y0 = mby * 16;
#define REF(q) (&slots[clip3 (sl->ref_slot[m->ref_idx[q] & 15], 0, n_slots - 1)])
#define WIDX(q) (m->ref_idx[q] & 15)
  switch (m->mb_type) {
  case LH264_MB_PSKIP: case LH264_MB_P16x16:
    base_mc (cur, REF (0), sl, WIDX (0), x0, y0, 16, 16, m->mv[0][0], m->mv[0][1]);
    break;
  case LH264_MB_P16x8:
    base_mc (cur, REF (0), sl, WIDX (0), x0, y0, 16, 8, m->mv[0][0], m->mv[0][1]);
    base_mc (cur, REF (2), sl, WIDX (2), x0, y0 + 8, 16, 8, m->mv[8][0], m->mv[8][1]);
    break;
  case LH264_MB_P8x16:
    base_mc (cur, REF (0), sl, WIDX (0), x0, y0, 8, 16, m->mv[0][0], m->mv[0][1]);
    base_mc (cur, REF (1), sl, WIDX (1), x0 + 8, y0, 8, 16, m->mv[2][0], m->mv[2][1]);
    break;
  case LH264_MB_P8x8:
    for (int q = 0; q < 4; q++) {
      int bx = (q & 1) * 8, by = (q >> 1) * 8, r = (q >> 1) * 8 + (q & 1) * 2;
      /* with explicit WP the reference indexes the weights by ref idx, else by 0 (rec_mb.cpp:449) */
      int widx = sl->weighted_pred ? WIDX (q) : 0;
      switch (m->sub_mb_type[q]) {
      case LH264_SUB_8x8:
        base_mc (cur, REF (q), sl, widx, x0 + bx, y0 + by, 8, 8, m->mv[r][0], m->mv[r][1]);
        break;
      case LH264_SUB_8x4:
        base_mc (cur, REF (q), sl, widx, x0 + bx, y0 + by, 8, 4, m->mv[r][0], m->mv[r][1]);
        base_mc (cur, REF (q), sl, widx, x0 + bx, y0 + by + 4, 8, 4, m->mv[r + 4][0], m->mv[r + 4][1]);
        break;
      case LH264_SUB_4x8:
        base_mc (cur, REF (q), sl, widx, x0 + bx, y0 + by, 4, 8, m->mv[r][0], m->mv[r][1]);
        base_mc (cur, REF (q), sl, widx, x0 + bx + 4, y0 + by, 4, 8, m->mv[r + 1][0], m->mv[r + 1][1]);
        break;
      default:
        for (int j = 0; j < 4; j++) {
          int rr = r + (j >> 1) * 4 + (j & 1);
          base_mc (cur, REF (q), sl, widx, x0 + bx + (j & 1) * 4, y0 + by + (j >> 1) * 4, 4, 4, m->mv[rr][0], m->mv[rr][1]);
        }
      }
    }
    break;
  default: break;
  }
#undef REF
#undef WIDX
}

/* ------------------------------------------------------------------------------------------------
 * Intra prediction: codec/decoder/core/src/get_intra_predictor.cpp:54-1155 (in place). */
static void pred4x4 (uint8_t* p, int s, int mode) {
  int t[8], l[4], tl = 0;
  const int need_top = ! (mode == LH264_I4_H || mode == LH264_I4_HU || mode == LH264_I4_DC_L || mode == LH264_I4_DC_128);
  const int need_left = ! (mode == LH264_I4_V || mode == LH264_I4_DDL || mode == LH264_I4_VL || mode == LH264_I4_DC_T ||
                           mode == LH264_I4_DC_128 || mode == LH264_I4_DDL_TOP || mode == LH264_I4_VL_TOP);
  memset (t, 0, sizeof (t)); memset (l, 0, sizeof (l));
  if (need_top) {
    for (int i = 0; i < 4; i++) t[i] = p[-s + i];
    if (mode == LH264_I4_DDL || mode == LH264_I4_VL) for (int i = 4; i < 8; i++) t[i] = p[-s + i];
    else for (int i = 4; i < 8; i++) t[i] = t[3];
  }
  if (need_left) for (int i = 0; i < 4; i++) l[i] = p[i * s - 1];
  if (mode == LH264_I4_DDR || mode == LH264_I4_VR || mode == LH264_I4_HD) tl = p[-s - 1];
#define P(x, y) p[(y) * s + (x)]
#define F3(a, b, c) (((a) + 2 * (b) + (c) + 2) >> 2)
#define F2(a, b) (((a) + (b) + 1) >> 1)
  switch (mode) {
  case LH264_I4_V: for (int y = 0; y < 4; y++) for (int x = 0; x < 4; x++) P (x, y) = (uint8_t)t[x]; break;
  case LH264_I4_H: for (int y = 0; y < 4; y++) for (int x = 0; x < 4; x++) P (x, y) = (uint8_t)l[y]; break;
  case LH264_I4_DC: case LH264_I4_DC_L: case LH264_I4_DC_T: case LH264_I4_DC_128: {
    int dc = 128;
    if (mode == LH264_I4_DC) dc = (t[0] + t[1] + t[2] + t[3] + l[0] + l[1] + l[2] + l[3] + 4) >> 3;
    else if (mode == LH264_I4_DC_L) dc = (l[0] + l[1] + l[2] + l[3] + 2) >> 2;
    else if (mode == LH264_I4_DC_T) dc = (t[0] + t[1] + t[2] + t[3] + 2) >> 2;
    for (int y = 0; y < 4; y++) for (int x = 0; x < 4; x++) P (x, y) = (uint8_t)dc;
    break;
  }
  case LH264_I4_DDL: case LH264_I4_DDL_TOP:
    for (int y = 0; y < 4; y++) for (int x = 0; x < 4; x++)
        P (x, y) = (uint8_t) ((x == 3 && y == 3) ? F3 (t[6], t[7], t[7]) : F3 (t[x + y], t[x + y + 1], t[x + y + 2]));
    break;
  case LH264_I4_DDR: {
    int e[9] = {l[3], l[2], l[1], l[0], tl, t[0], t[1], t[2], t[3]};   /* edge from bottom-left to top-right */
    for (int y = 0; y < 4; y++) for (int x = 0; x < 4; x++) { int k = 4 + x - y; P (x, y) = (uint8_t)F3 (e[k - 1], e[k], e[k + 1]); }
    break;
  }
  case LH264_I4_VR: {
    int e[9] = {l[3], l[2], l[1], l[0], tl, t[0], t[1], t[2], t[3]};
    for (int y = 0; y < 4; y++) for (int x = 0; x < 4; x++) {
        int z = 2 * x - y;
        if (z >= 0 && ! (z & 1)) P (x, y) = (uint8_t)F2 (e[4 + x - (y >> 1)], e[5 + x - (y >> 1)]);
        else if (z >= 0) P (x, y) = (uint8_t)F3 (e[3 + x - (y >> 1)], e[4 + x - (y >> 1)], e[5 + x - (y >> 1)]);
        else if (z == -1) P (x, y) = (uint8_t)F3 (e[3], e[4], e[5]);
        else P (x, y) = (uint8_t)F3 (e[4 - y + 2], e[4 - y + 1], e[4 - y]);   /* l[y-1], l[y-2], l[y-3] with l[-1] = tl */
      }
    break;
  }
  case LH264_I4_HD: {
    int e[9] = {l[3], l[2], l[1], l[0], tl, t[0], t[1], t[2], t[3]};
    for (int y = 0; y < 4; y++) for (int x = 0; x < 4; x++) {
        int z = 2 * y - x;
        if (z >= 0 && ! (z & 1)) P (x, y) = (uint8_t)F2 (e[4 - (y - (x >> 1))], e[3 - (y - (x >> 1))]);
        else if (z >= 0) P (x, y) = (uint8_t)F3 (e[5 - (y - (x >> 1))], e[4 - (y - (x >> 1))], e[3 - (y - (x >> 1))]);
        else if (z == -1) P (x, y) = (uint8_t)F3 (e[3], e[4], e[5]);
        else P (x, y) = (uint8_t)F3 (e[4 + x], e[4 + x - 1], e[4 + x - 2]);   /* t[x-1], t[x-2], t[x-3] with t[-1] = tl */
      }
    break;
  }
  case LH264_I4_VL: case LH264_I4_VL_TOP:
    for (int y = 0; y < 4; y++) for (int x = 0; x < 4; x++) {
        int k = x + (y >> 1);
        P (x, y) = (uint8_t) ((y & 1) ? F3 (t[k], t[k + 1], t[k + 2]) : F2 (t[k], t[k + 1]));
      }
    break;
  case LH264_I4_HU:
    for (int y = 0; y < 4; y++) for (int x = 0; x < 4; x++) {
        int z = x + 2 * y;
        if (z > 5) P (x, y) = (uint8_t)l[3];
        else if (z == 5) P (x, y) = (uint8_t)F3 (l[2], l[3], l[3]);
        else if (z & 1) P (x, y) = (uint8_t)F3 (l[z >> 1], l[(z >> 1) + 1], l[(z >> 1) + 2]);
        else P (x, y) = (uint8_t)F2 (l[z >> 1], l[(z >> 1) + 1]);
      }
    break;
  default: break;
  }
}

/* 8x8 predictors with reference-sample filtering, get_intra_predictor.cpp:382-880. */
static void pred8x8l (uint8_t* p, int s, int mode, int has_tl, int has_tr) {
  int ft[16], fl[8], ftl = 0;
  const int use_top = ! (mode == LH264_I4_H || mode == LH264_I4_HU || mode == LH264_I4_DC_L || mode == LH264_I4_DC_128);
  const int use_left = ! (mode == LH264_I4_V || mode == LH264_I4_DDL || mode == LH264_I4_VL || mode == LH264_I4_DC_T ||
                          mode == LH264_I4_DC_128 || mode == LH264_I4_DDL_TOP || mode == LH264_I4_VL_TOP);
  memset (ft, 0, sizeof (ft)); memset (fl, 0, sizeof (fl));
  if (mode == LH264_I4_DDR || mode == LH264_I4_VR || mode == LH264_I4_HD) {
    has_tl = 1;                                   /* "TopLeft, Top, Left are all available under this mode" */
    ftl = (p[-1] + 2 * p[-s - 1] + p[-s] + 2) >> 2;
  }
  if (use_top) {
    int t[16];
    for (int i = 0; i < 8; i++) t[i] = p[-s + i];
    int real_tr;
    if (mode == LH264_I4_DDL || mode == LH264_I4_VL) real_tr = 1;
    else if (mode == LH264_I4_DDL_TOP || mode == LH264_I4_VL_TOP) real_tr = 0;
    else real_tr = has_tr;
    for (int i = 8; i < 16; i++) t[i] = real_tr ? p[-s + i] : t[7];
    ft[0] = has_tl ? ((p[-s - 1] + 2 * t[0] + t[1] + 2) >> 2) : ((3 * t[0] + t[1] + 2) >> 2);
    for (int i = 1; i < 15; i++) ft[i] = (t[i - 1] + 2 * t[i] + t[i + 1] + 2) >> 2;
    ft[15] = (t[14] + 3 * t[15] + 2) >> 2;
  }
  if (use_left) {
    int l[8];
    for (int i = 0; i < 8; i++) l[i] = p[i * s - 1];
    fl[0] = has_tl ? ((p[-s - 1] + 2 * l[0] + l[1] + 2) >> 2) : ((3 * l[0] + l[1] + 2) >> 2);
    for (int i = 1; i < 7; i++) fl[i] = (l[i - 1] + 2 * l[i] + l[i + 1] + 2) >> 2;
    fl[7] = (l[6] + 3 * l[7] + 2) >> 2;
  }
  switch (mode) {
  case LH264_I4_V: for (int y = 0; y < 8; y++) for (int x = 0; x < 8; x++) P (x, y) = (uint8_t)ft[x]; break;
  case LH264_I4_H: for (int y = 0; y < 8; y++) for (int x = 0; x < 8; x++) P (x, y) = (uint8_t)fl[y]; break;
  case LH264_I4_DC: case LH264_I4_DC_L: case LH264_I4_DC_T: case LH264_I4_DC_128: {
    int st = 0, sl = 0, dc = 128;
    for (int i = 0; i < 8; i++) { st += ft[i]; sl += fl[i]; }
    if (mode == LH264_I4_DC) dc = (st + sl + 8) >> 4;
    else if (mode == LH264_I4_DC_L) dc = (sl + 4) >> 3;
    else if (mode == LH264_I4_DC_T) dc = (st + 4) >> 3;
    for (int y = 0; y < 8; y++) for (int x = 0; x < 8; x++) P (x, y) = (uint8_t)dc;
    break;
  }
  case LH264_I4_DDL: case LH264_I4_DDL_TOP:
    for (int y = 0; y < 8; y++) for (int x = 0; x < 8; x++)
        P (x, y) = (uint8_t) ((x == 7 && y == 7) ? ((ft[14] + 3 * ft[15] + 2) >> 2) : F3 (ft[x + y], ft[x + y + 1], ft[x + y + 2]));
    break;
  case LH264_I4_DDR: case LH264_I4_VR: case LH264_I4_HD: {
    int e[17];                                   /* fl[7]..fl[0], ftl, ft[0]..ft[7] */
    for (int i = 0; i < 8; i++) { e[7 - i] = fl[i]; e[9 + i] = ft[i]; }
    e[8] = ftl;
    for (int y = 0; y < 8; y++) for (int x = 0; x < 8; x++) {
        if (mode == LH264_I4_DDR) { int k = 8 + x - y; P (x, y) = (uint8_t)F3 (e[k - 1], e[k], e[k + 1]); }
        else if (mode == LH264_I4_VR) {
          int z = 2 * x - y, k = 8 + x - (y >> 1);
          if (z >= 0 && ! (z & 1)) P (x, y) = (uint8_t)F2 (e[k], e[k + 1]);
          else if (z >= 0) P (x, y) = (uint8_t)F3 (e[k - 1], e[k], e[k + 1]);
          else if (z == -1) P (x, y) = (uint8_t)F3 (e[7], e[8], e[9]);
          else { int q = 8 - y + 2 * x; P (x, y) = (uint8_t)F3 (e[q + 2], e[q + 1], e[q]); }  /* fl[y-2x-1], [y-2x-2], [y-2x-3] */
        } else {
          int z = 2 * y - x, k = 8 - (y - (x >> 1));
          if (z >= 0 && ! (z & 1)) P (x, y) = (uint8_t)F2 (e[k], e[k - 1]);
          else if (z >= 0) P (x, y) = (uint8_t)F3 (e[k + 1], e[k], e[k - 1]);
          else if (z == -1) P (x, y) = (uint8_t)F3 (e[7], e[8], e[9]);
          else { int q = 8 + x - 2 * y; P (x, y) = (uint8_t)F3 (e[q], e[q - 1], e[q - 2]); }  /* ft[x-2y-1], [x-2y-2], [x-2y-3] */
        }
      }
    break;
  }
  case LH264_I4_VL: case LH264_I4_VL_TOP:
    for (int y = 0; y < 8; y++) for (int x = 0; x < 8; x++) {
        int k = x + (y >> 1);
        P (x, y) = (uint8_t) ((y & 1) ? F3 (ft[k], ft[k + 1], ft[k + 2]) : F2 (ft[k], ft[k + 1]));
      }
    break;
  case LH264_I4_HU:
    for (int y = 0; y < 8; y++) for (int x = 0; x < 8; x++) {
        int z = x + 2 * y;
        if (z > 13) P (x, y) = (uint8_t)fl[7];
        else if (z == 13) P (x, y) = (uint8_t) ((fl[6] + 3 * fl[7] + 2) >> 2);
        else if (z & 1) P (x, y) = (uint8_t)F3 (fl[z >> 1], fl[(z >> 1) + 1], fl[(z >> 1) + 2]);
        else P (x, y) = (uint8_t)F2 (fl[z >> 1], fl[(z >> 1) + 1]);
      }
    break;
  default: break;
  }
}
#undef F2
#undef F3

static void pred16x16 (uint8_t* p, int s, int mode) {
  switch (mode) {
  case LH264_I16_V: for (int y = 0; y < 16; y++) for (int x = 0; x < 16; x++) P (x, y) = p[-s + x]; break;
  case LH264_I16_H: for (int y = 0; y < 16; y++) for (int x = 0; x < 16; x++) P (x, y) = p[y * s - 1]; break;
  case LH264_I16_P: {
    int H = 0, V = 0;
    for (int i = 0; i < 8; i++) { H += (i + 1) * (p[-s + 8 + i] - p[-s + 6 - i]); V += (i + 1) * (p[(8 + i) * s - 1] - p[(6 - i) * s - 1]); }
    int a = (p[15 * s - 1] + p[-s + 15]) << 4, b = (5 * H + 32) >> 6, c = (5 * V + 32) >> 6;
    for (int y = 0; y < 16; y++) for (int x = 0; x < 16; x++) P (x, y) = clip1 ((a + b * (x - 7) + c * (y - 7) + 16) >> 5);
    break;
  }
  default: {
    int st = 0, sl = 0, dc = 128;
    if (mode == LH264_I16_DC || mode == LH264_I16_DC_T) for (int i = 0; i < 16; i++) st += p[-s + i];
    if (mode == LH264_I16_DC || mode == LH264_I16_DC_L) for (int i = 0; i < 16; i++) sl += p[i * s - 1];
    if (mode == LH264_I16_DC) dc = (st + sl + 16) >> 5;
    else if (mode == LH264_I16_DC_T) dc = (st + 8) >> 4;
    else if (mode == LH264_I16_DC_L) dc = (sl + 8) >> 4;
    for (int y = 0; y < 16; y++) for (int x = 0; x < 16; x++) P (x, y) = (uint8_t)dc;
  }
  }
}
static void pred_chroma (uint8_t* p, int s, int mode) {
  switch (mode) {
  case LH264_C_V: for (int y = 0; y < 8; y++) for (int x = 0; x < 8; x++) P (x, y) = p[-s + x]; break;
  case LH264_C_H: for (int y = 0; y < 8; y++) for (int x = 0; x < 8; x++) P (x, y) = p[y * s - 1]; break;
  case LH264_C_P: {
    int H = 0, V = 0;
    for (int i = 0; i < 4; i++) { H += (i + 1) * (p[-s + 4 + i] - p[-s + 2 - i]); V += (i + 1) * (p[(4 + i) * s - 1] - p[(2 - i) * s - 1]); }
    int a = (p[7 * s - 1] + p[-s + 7]) << 4, b = (17 * H + 16) >> 5, c = (17 * V + 16) >> 5;
    for (int y = 0; y < 8; y++) for (int x = 0; x < 8; x++) P (x, y) = clip1 ((a + b * (x - 3) + c * (y - 3) + 16) >> 5);
    break;
  }
  default: {
    int dc[4] = {128, 128, 128, 128};   /* quadrants: TL, TR, BL, BR */
    int t0 = 0, t1 = 0, l0 = 0, l1 = 0;
    if (mode == LH264_C_DC || mode == LH264_C_DC_T) for (int i = 0; i < 4; i++) { t0 += p[-s + i]; t1 += p[-s + 4 + i]; }
    if (mode == LH264_C_DC || mode == LH264_C_DC_L) for (int i = 0; i < 4; i++) { l0 += p[i * s - 1]; l1 += p[(4 + i) * s - 1]; }
    if (mode == LH264_C_DC) { dc[0] = (t0 + l0 + 4) >> 3; dc[1] = (t1 + 2) >> 2; dc[2] = (l1 + 2) >> 2; dc[3] = (t1 + l1 + 4) >> 3; }
    else if (mode == LH264_C_DC_T) { dc[0] = dc[2] = (t0 + 2) >> 2; dc[1] = dc[3] = (t1 + 2) >> 2; }
    else if (mode == LH264_C_DC_L) { dc[0] = dc[1] = (l0 + 2) >> 2; dc[2] = dc[3] = (l1 + 2) >> 2; }
    for (int y = 0; y < 8; y++) for (int x = 0; x < 8; x++) P (x, y) = (uint8_t)dc[(y >> 2) * 2 + (x >> 2)];
  }
  }
}
#undef P

/* ------------------------------------------------------------------------------------------------
 * One macroblock: WelsTargetMbConstruction, decode_slice.cpp:353-373 and callees. */
static const uint8_t kBlkX[16] = {0, 4, 0, 4, 8, 12, 8, 12, 0, 4, 0, 4, 8, 12, 8, 12};
static const uint8_t kBlkY[16] = {0, 0, 4, 4, 0, 0, 4, 4, 8, 8, 12, 12, 8, 8, 12, 12};

/* expands the MB's present blocks into a dense 24x16 level array */
static void gather_levels (const lh264_mb_rec* m, const int16_t* pool, int16_t lev[384]) {
  const int16_t* src = pool + (size_t)m->coef_off * 16;
  memset (lev, 0, 384 * sizeof (int16_t));
  for (int k = 0; k < 24; k++) if (m->coded_mask & (1u << k)) { memcpy (lev + 16 * k, src, 32); src += 16; }
}

static void residual_chroma (const ora_picture* cur, const lh264_mb_rec* m, const int16_t* lev, int mbx, int mby, const lh264_slice_rec* sc) {
  /* RecChroma (rec_mb.cpp:547-575) / WelsMbInterSampleConstruction chroma part (decode_slice.cpp:213-247) */
  const int cw = cur->w >> 1;
  for (int p = 0; p < 2; p++) {
    int16_t c[64];
    const int qp = m->qp_c[p];
    for (int b = 0; b < 4; b++) {
      const int16_t* l = lev + 16 * (16 + 4 * p + b);
      const int list = (LH264_MB_IS_INTRA (m->mb_type) ? 1 : 4) + p;
      c[16 * b] = sc ? (int16_t) ((l[0] * sl_dq4 (sc, list, qp, 0)) >> 4) : (int16_t) (l[0] * dq4 (qp, 0));           /* parse_mb_syn_cavlc.cpp:944 */
      dequant_ac4x4 (c + 16 * b, l, qp, 1, sc, list);
    }
    chroma_dc_idct (c);
    for (int b = 0; b < 4; b++) {
      uint8_t* d = cur->plane[1 + p] + (size_t) (mby * 8 + (b >> 1) * 4) * cw + mbx * 8 + (b & 1) * 4;
      idct4x4_add (d, cw, c + 16 * b);
    }
  }
}

static void recon_mb (const ora_picture* cur, const ora_picture* slots, int n_slots, const lh264_slice_rec* sl,
                      const lh264_mb_rec* m, const int16_t* pool, int mbx, int mby) {
  const int w = cur->w, cw = w >> 1;
  uint8_t* py = cur->plane[0] + (size_t)mby * 16 * w + mbx * 16;
  int16_t lev[384], c[64];
  const int t8 = m->flags & LH264_MBF_TRANSFORM8x8;
  const lh264_slice_rec* sc = (sl->scaling_flags & LH264_SLF_SCALING) ? sl : NULL;
  if (m->mb_type == LH264_MB_NONE) return;
  gather_levels (m, pool, lev);
  if (m->mb_type == LH264_MB_IPCM) {            /* decode_slice.cpp:3213-3263: samples copied at parse time */
    for (int y = 0; y < 16; y++) for (int x = 0; x < 16; x++) py[y * w + x] = (uint8_t)lev[y * 16 + x];
    for (int p = 0; p < 2; p++) for (int y = 0; y < 8; y++) for (int x = 0; x < 8; x++)
          cur->plane[1 + p][(size_t) (mby * 8 + y) * cw + mbx * 8 + x] = (uint8_t)lev[256 + 64 * p + y * 8 + x];
    return;
  }
  if (LH264_MB_IS_INTER (m->mb_type)) {
    inter_pred (cur, slots, n_slots, sl, m, mbx, mby);
    if (m->cbp == 0) return;                    /* WelsMbInterPrediction only, decode_slice.cpp:362-364 */
    if (t8) {
      for (int q = 0; q < 4; q++) { dequant_8x8 (c, lev + 64 * q, m->qp_y, sc, 0); idct8x8_add (py + kBlkY[4 * q] * w + kBlkX[4 * q], w, c); }
    } else {
      for (int z = 0; z < 16; z++) { dequant_ac4x4 (c, lev + 16 * z, m->qp_y, 0, sc, 3); idct4x4_add (py + kBlkY[z] * w + kBlkX[z], w, c); }
    }
    residual_chroma (cur, m, lev, mbx, mby, sc);
    return;
  }
  /* intra: WelsMbIntraPredictionConstruction, decode_slice.cpp:313-334 */
  if (m->mb_type == LH264_MB_I16x16) {          /* RecI16x16Mb, rec_mb.cpp:179-219 */
    int16_t blk[256];
    for (int z = 0; z < 16; z++) { blk[16 * z] = lev[16 * z]; dequant_ac4x4 (blk + 16 * z, lev + 16 * z, m->qp_y, 1, sc, 0); }
    luma_dc_dequant_idct (blk, m->qp_y, sc);
    pred16x16 (py, w, m->i16_mode);
    for (int z = 0; z < 16; z++) idct4x4_add (py + kBlkY[z] * w + kBlkX[z], w, blk + 16 * z);
  } else if (m->mb_type == LH264_MB_I8x8) {     /* RecI8x8Luma, rec_mb.cpp:70-115 */
    const int f = m->nxn_avail;
    const int tl[4] = { !! (f & 2), !! (f & 1), !! (f & 4), 1 }, tr[4] = { !! (f & 1), !! (f & 8), 1, 0 };
    for (int q = 0; q < 4; q++) {
      uint8_t* d = py + kBlkY[4 * q] * w + kBlkX[4 * q];
      pred8x8l (d, w, m->intra_mode[4 * q], tl[q], tr[q]);
      dequant_8x8 (c, lev + 64 * q, m->qp_y, sc, 1);
      idct8x8_add (d, w, c);
    }
  } else {                                      /* RecI4x4Luma, rec_mb.cpp:124-158 */
    for (int z = 0; z < 16; z++) {
      uint8_t* d = py + kBlkY[z] * w + kBlkX[z];
      pred4x4 (d, w, m->intra_mode[z]);
      dequant_ac4x4 (c, lev + 16 * z, m->qp_y, 0, sc, 0);
      idct4x4_add (d, w, c);
    }
  }
  pred_chroma (cur->plane[1] + (size_t)mby * 8 * cw + mbx * 8, cw, m->chroma_mode);
  pred_chroma (cur->plane[2] + (size_t)mby * 8 * cw + mbx * 8, cw, m->chroma_mode);
  { int cc = m->cbp >> 4; if (cc == 1 || cc == 2) residual_chroma (cur, m, lev, mbx, mby, sc); }
}

/* ------------------------------------------------------------------------------------------------
 * Deblocking: codec/decoder/core/src/deblocking.cpp:160-933 + codec/common/src/deblocking_common.cpp:5-246 */
static const uint8_t kAlpha[52] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 4, 4, 5, 6, 7, 8, 9, 10, 12, 13, 15, 17, 20, 22,
                                   25, 28, 32, 36, 40, 45, 50, 56, 63, 71, 80, 90, 101, 113, 127, 144, 162, 182, 203, 226, 255, 255
                                  };
static const uint8_t kBeta[52] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 6, 6, 7, 7,
                                  8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13, 14, 14, 15, 15, 16, 16, 17, 17, 18, 18
                                 };
static const int8_t kTc0[52][4] = {
  { -1, 0, 0, 0}, { -1, 0, 0, 0}, { -1, 0, 0, 0}, { -1, 0, 0, 0}, { -1, 0, 0, 0}, { -1, 0, 0, 0}, { -1, 0, 0, 0}, { -1, 0, 0, 0},
  { -1, 0, 0, 0}, { -1, 0, 0, 0}, { -1, 0, 0, 0}, { -1, 0, 0, 0}, { -1, 0, 0, 0}, { -1, 0, 0, 0}, { -1, 0, 0, 0}, { -1, 0, 0, 0},
  { -1, 0, 0, 0}, { -1, 0, 0, 1}, { -1, 0, 0, 1}, { -1, 0, 0, 1}, { -1, 0, 0, 1}, { -1, 0, 1, 1}, { -1, 0, 1, 1}, { -1, 1, 1, 1},
  { -1, 1, 1, 1}, { -1, 1, 1, 1}, { -1, 1, 1, 1}, { -1, 1, 1, 2}, { -1, 1, 1, 2}, { -1, 1, 1, 2}, { -1, 1, 1, 2}, { -1, 1, 2, 3},
  { -1, 1, 2, 3}, { -1, 2, 2, 3}, { -1, 2, 2, 4}, { -1, 2, 3, 4}, { -1, 2, 3, 4}, { -1, 3, 3, 5}, { -1, 3, 4, 6}, { -1, 3, 4, 6},
  { -1, 4, 5, 7}, { -1, 4, 5, 8}, { -1, 4, 6, 9}, { -1, 5, 7, 10}, { -1, 6, 8, 11}, { -1, 6, 8, 13}, { -1, 7, 10, 14}, { -1, 8, 11, 16},
  { -1, 9, 12, 18}, { -1, 10, 13, 20}, { -1, 11, 15, 23}, { -1, 13, 17, 25}
};
/* the reference tables are offset by 12 and saturate at both ends (deblocking.cpp:89-125) */
static inline int tab_idx (int v) { return clip3 (v, 0, 51); }

static void filter_luma_lt4 (uint8_t* p, int sx, int sy, int alpha, int beta, const int8_t tc4[4]) {
  for (int i = 0; i < 16; i++, p += sy) {       /* DeblockLumaLt4_c */
    int tc0 = tc4[i >> 2];
    if (tc0 < 0) continue;
    int p0 = p[-sx], p1 = p[-2 * sx], p2 = p[-3 * sx], q0 = p[0], q1 = p[sx], q2 = p[2 * sx];
    if (iabs (p0 - q0) < alpha && iabs (p1 - p0) < beta && iabs (q1 - q0) < beta) {
      int tc = tc0;
      if (iabs (p2 - p0) < beta) { p[-2 * sx] = (uint8_t) (p1 + clip3 ((p2 + ((p0 + q0 + 1) >> 1) - (p1 << 1)) >> 1, -tc0, tc0)); tc++; }
      if (iabs (q2 - q0) < beta) { p[sx] = (uint8_t) (q1 + clip3 ((q2 + ((p0 + q0 + 1) >> 1) - (q1 << 1)) >> 1, -tc0, tc0)); tc++; }
      int d = clip3 ((((q0 - p0) << 2) + (p1 - q1) + 4) >> 3, -tc, tc);
      p[-sx] = clip1 (p0 + d); p[0] = clip1 (q0 - d);
    }
  }
}
static void filter_luma_eq4 (uint8_t* p, int sx, int sy, int alpha, int beta) {
  for (int i = 0; i < 16; i++, p += sy) {       /* DeblockLumaEq4_c */
    int p0 = p[-sx], p1 = p[-2 * sx], p2 = p[-3 * sx], q0 = p[0], q1 = p[sx], q2 = p[2 * sx];
    int d = iabs (p0 - q0);
    if (d < alpha && iabs (p1 - p0) < beta && iabs (q1 - q0) < beta) {
      if (d < ((alpha >> 2) + 2)) {
        if (iabs (p2 - p0) < beta) {
          int p3 = p[-4 * sx];
          p[-sx] = (uint8_t) ((p2 + 2 * p1 + 2 * p0 + 2 * q0 + q1 + 4) >> 3);
          p[-2 * sx] = (uint8_t) ((p2 + p1 + p0 + q0 + 2) >> 2);
          p[-3 * sx] = (uint8_t) ((2 * p3 + 3 * p2 + p1 + p0 + q0 + 4) >> 3);
        } else p[-sx] = (uint8_t) ((2 * p1 + p0 + q1 + 2) >> 2);
        if (iabs (q2 - q0) < beta) {
          int q3 = p[3 * sx];
          p[0] = (uint8_t) ((p1 + 2 * p0 + 2 * q0 + 2 * q1 + q2 + 4) >> 3);
          p[sx] = (uint8_t) ((p0 + q0 + q1 + q2 + 2) >> 2);
          p[2 * sx] = (uint8_t) ((2 * q3 + 3 * q2 + q1 + q0 + p0 + 4) >> 3);
        } else p[0] = (uint8_t) ((2 * q1 + q0 + p1 + 2) >> 2);
      } else {
        p[-sx] = (uint8_t) ((2 * p1 + p0 + q1 + 2) >> 2);
        p[0] = (uint8_t) ((2 * q1 + q0 + p1 + 2) >> 2);
      }
    }
  }
}
static void filter_chroma_lt4 (uint8_t* p, int sx, int sy, int alpha, int beta, const int8_t tc4[4]) {
  for (int i = 0; i < 8; i++, p += sy) {        /* DeblockChromaLt4_c (one plane) */
    int tc0 = tc4[i >> 1];
    if (tc0 <= 0) continue;
    int p0 = p[-sx], p1 = p[-2 * sx], q0 = p[0], q1 = p[sx];
    if (iabs (p0 - q0) < alpha && iabs (p1 - p0) < beta && iabs (q1 - q0) < beta) {
      int d = clip3 ((((q0 - p0) << 2) + (p1 - q1) + 4) >> 3, -tc0, tc0);
      p[-sx] = clip1 (p0 + d); p[0] = clip1 (q0 - d);
    }
  }
}
static void filter_chroma_eq4 (uint8_t* p, int sx, int sy, int alpha, int beta) {
  for (int i = 0; i < 8; i++, p += sy) {        /* DeblockChromaEq4_c (one plane) */
    int p0 = p[-sx], p1 = p[-2 * sx], q0 = p[0], q1 = p[sx];
    if (iabs (p0 - q0) < alpha && iabs (p1 - p0) < beta && iabs (q1 - q0) < beta) {
      p[-sx] = (uint8_t) ((2 * p1 + p0 + q1 + 2) >> 2); p[0] = (uint8_t) ((2 * q1 + q0 + p1 + 2) >> 2);
    }
  }
}

typedef struct { const lh264_slice_rec* sl; const ora_picture* pic; } dbk_ctx;

/* FilteringEdgeLuma{V,H} / FilteringEdgeLumaIntra{V,H}: dir 0 = vertical edge (filter across x) */
static void edge_luma (const dbk_ctx* c, uint8_t* pix, int dir, int qp, const uint8_t bs[4]) {
  int ia = tab_idx (qp + c->sl->alpha_c0_offset), alpha = kAlpha[ia], beta = kBeta[tab_idx (qp + c->sl->beta_offset)];
  int sx = dir ? c->pic->w : 1, sy = dir ? 1 : c->pic->w;
  if (! (alpha | beta)) return;
  if (bs[0] == 4) filter_luma_eq4 (pix, sx, sy, alpha, beta);
  else { int8_t tc[4]; for (int i = 0; i < 4; i++) tc[i] = kTc0[ia][bs[i]]; filter_luma_lt4 (pix, sx, sy, alpha, beta, tc); }
}
/* FilteringEdgeChroma{V,H}: both planes, each with its own averaged QP */
static void edge_chroma (const dbk_ctx* c, uint8_t* cb, uint8_t* cr, int dir, const int qp[2], const uint8_t bs[4]) {
  int cw = c->pic->w >> 1, sx = dir ? cw : 1, sy = dir ? 1 : cw;
  uint8_t* pl[2] = {cb, cr};
  for (int p = 0; p < 2; p++) {
    int ia = tab_idx (qp[p] + c->sl->alpha_c0_offset), alpha = kAlpha[ia], beta = kBeta[tab_idx (qp[p] + c->sl->beta_offset)];
    if (! (alpha | beta)) continue;
    if (bs[0] == 4) filter_chroma_eq4 (pl[p], sx, sy, alpha, beta);
    else { int8_t tc[4]; for (int i = 0; i < 4; i++) tc[i] = (int8_t) (kTc0[ia][bs[i]] + 1); filter_chroma_lt4 (pl[p], sx, sy, alpha, beta, tc); }
  }
}

static inline int nzb (const lh264_mb_rec* m, int r) { return (m->nz_mask >> r) & 1; }
static inline int mvdiff (const lh264_mb_rec* a, int ra, const lh264_mb_rec* b, int rb) {
  return (a->ref_idx[(ra >> 3) * 2 + ((ra & 3) >> 1)] != b->ref_idx[(rb >> 3) * 2 + ((rb & 3) >> 1)]) ||
         iabs (a->mv[ra][0] - b->mv[rb][0]) >= 4 || iabs (a->mv[ra][1] - b->mv[rb][1]) >= 4;
}
/* nz of the 8x8 quadrant that contains raster block r */
static inline int nz8 (const lh264_mb_rec* m, int r) {
  int b = (r & 8) + ((r & 2));   /* top-left raster block of the quadrant */
  return nzb (m, b) | nzb (m, b + 1) | nzb (m, b + 4) | nzb (m, b + 5);
}

/* DeblockingBsMarginalMBAvcbase (deblocking.cpp:280-353): bS of the 4 segments of an MB edge. */
static void bs_mb_edge (const lh264_mb_rec* cur, const lh264_mb_rec* nb, int dir, uint8_t bs[4]) {
  const int c8 = cur->flags & LH264_MBF_TRANSFORM8x8, n8 = nb->flags & LH264_MBF_TRANSFORM8x8;
  for (int i = 0; i < 4; i++) {
    int rc = dir ? i : 4 * i;                   /* current MB: first column / first row */
    int rn = dir ? 12 + i : 4 * i + 3;          /* neighbour: last row / last column */
    int nz, rcm = rc, rnm = rn;
    if (c8 && n8) {
      nz = nz8 (cur, rc) | nz8 (nb, rn);
      /* MVs/refs compared at the first block of each 8x8 pair (g_kuiTableB8x8Idx) */
      rcm = dir ? (i & 2) : 4 * (i & 2); rnm = dir ? 8 + (i & 2) : 4 * (i & 2) + 2;
    } else if (c8) {
      nz = nz8 (cur, rc) | nzb (nb, rn);
      rcm = dir ? (i & 2) : 4 * (i & 2);
    } else if (n8) {
      nz = nzb (cur, rc) | nz8 (nb, rn);
      rnm = dir ? 8 + (i & 2) : 4 * (i & 2) + 2;
    } else nz = nzb (cur, rc) | nzb (nb, rn);
    bs[i] = (uint8_t) (nz ? 2 : mvdiff (cur, rcm, nb, rnm));
  }
}

static void deblock_mb (const dbk_ctx* c, const lh264_mb_rec* mbs, int mb_w, int mbx, int mby) {
  const lh264_mb_rec* m = &mbs[mby * mb_w + mbx];
  const lh264_slice_rec* sl = c->sl;
  const int w = c->pic->w, cw = w >> 1, idc = sl->disable_deblocking_idc;
  uint8_t* py = c->pic->plane[0] + (size_t)mby * 16 * w + mbx * 16;
  uint8_t* pcb = c->pic->plane[1] + (size_t)mby * 8 * cw + mbx * 8;
  uint8_t* pcr = c->pic->plane[2] + (size_t)mby * 8 * cw + mbx * 8;
  /* DeblockingAvailableNoInterlayer, deblocking.cpp:354-369 */
  int left = mbx > 0, top = mby > 0;
  if (idc == 2) { left = left && m[-1].slice_idx == m->slice_idx; top = top && m[-mb_w].slice_idx == m->slice_idx; }
  const int t8 = m->flags & LH264_MBF_TRANSFORM8x8;
  const int intra = LH264_MB_IS_INTRA (m->mb_type);
  uint8_t bs[2][4][4];
  memset (bs, 0, sizeof (bs));
  if (intra) {                                  /* DeblockingIntraMb: MB edges 4, inner edges 3 */
    memset (bs, 3, sizeof (bs));
    memset (bs[0][0], 4, 4); memset (bs[1][0], 4, 4);
  } else {                                      /* WelsDeblockingMb default branch, deblocking.cpp:829-860 */
    if (left) { if (LH264_MB_IS_INTRA (m[-1].mb_type)) memset (bs[0][0], 4, 4); else bs_mb_edge (m, m - 1, 0, bs[0][0]); }
    if (top) { if (LH264_MB_IS_INTRA (m[-mb_w].mb_type)) memset (bs[1][0], 4, 4); else bs_mb_edge (m, m - mb_w, 1, bs[1][0]); }
    if (m->mb_type != LH264_MB_PSKIP) {
      const int use_mv = m->mb_type != LH264_MB_P16x16;   /* 16x16: DeblockingBSInsideMBAvsbase ignores MVs */
      for (int e = 1; e < 4; e++) for (int i = 0; i < 4; i++) {
          if (t8 && e != 2) continue;
          /* vertical edge e, row i: blocks (4i+e) and (4i+e-1); horizontal edge e, col i: (4e+i) and (4e-4+i) */
          int a0 = 4 * i + e, b0 = a0 - 1, a1 = 4 * e + i, b1 = a1 - 4;
          int nz0 = t8 ? (nz8 (m, a0) | nz8 (m, b0)) : (nzb (m, a0) | nzb (m, b0));
          int nz1 = t8 ? (nz8 (m, a1) | nz8 (m, b1)) : (nzb (m, a1) | nzb (m, b1));
          /* in the 8x8-transform case MVs are compared at the quadrant origins (deblocking.cpp:216-230) */
          int ma0 = t8 ? ((a0 & 8) + (a0 & 2)) : a0, mb0 = t8 ? ((b0 & 8) + (b0 & 2)) : b0;
          int ma1 = t8 ? ((a1 & 8) + (a1 & 2)) : a1, mb1 = t8 ? ((b1 & 8) + (b1 & 2)) : b1;
          bs[0][e][i] = (uint8_t) (nz0 ? 2 : (use_mv ? mvdiff (m, ma0, m, mb0) : 0));
          bs[1][e][i] = (uint8_t) (nz1 ? 2 : (use_mv ? mvdiff (m, ma1, m, mb1) : 0));
        }
    }
  }
  for (int dir = 0; dir < 2; dir++) {
    const int avail = dir ? top : left;
    const lh264_mb_rec* nb = dir ? m - mb_w : m - 1;
    for (int e = 0; e < 4; e++) {
      const uint8_t* b = bs[dir][e];
      if (e == 0 && !avail) continue;
      if (e != 0 && (e & 1) && t8) continue;
      if (! (b[0] | b[1] | b[2] | b[3])) continue;
      int qp = m->qp_y, qpc[2] = {m->qp_c[0], m->qp_c[1]};
      if (e == 0) { qp = (qp + nb->qp_y + 1) >> 1; qpc[0] = (qpc[0] + nb->qp_c[0] + 1) >> 1; qpc[1] = (qpc[1] + nb->qp_c[1] + 1) >> 1; }
      edge_luma (c, py + (dir ? 4 * e * w : 4 * e), dir, qp, b);
      if (! (e & 1)) edge_chroma (c, pcb + (dir ? 2 * e * cw : 2 * e), pcr + (dir ? 2 * e * cw : 2 * e), dir, qpc, b);
    }
  }
}

/* ------------------------------------------------------------------------------------------------
 * One picture = every slice in arrival order: reconstruct the slice's MBs, then deblock them
 * (WelsTargetSliceConstruction, decode_slice.cpp:110-206).  slots[] are the stream's pictures. */
int ora_recon_picture (const lh264_picture_desc* pd, int mb_w, int mb_h, ora_picture* slots, int n_slots) {
  if (!pd || pd->cur_slot < 0 || pd->cur_slot >= n_slots) return LH264_E_INVALID;
  ora_picture* cur = &slots[pd->cur_slot];
  for (int s = 0; s < pd->n_slices; s++) {
    const lh264_slice_rec* sl = &pd->slices[s];
    /* the slice's macroblocks in address order; with slice groups (FMO) they are every n-th one, not a range (FmoNextMb,
     * decode_slice.cpp:173-177) */
    const int first = sl->first_mb, n = mb_w * mb_h;
    if (first < 0 || first > n || sl->num_mbs < 0 || sl->num_mbs > n) return LH264_E_INVALID;
    int done = 0;
    for (int xy = first; xy < n && done < sl->num_mbs; xy++)
      if (pd->mbs[xy].mb_type != LH264_MB_NONE && pd->mbs[xy].slice_idx == s) { recon_mb (cur, slots, n_slots, sl, &pd->mbs[xy], pd->coefs, xy % mb_w, xy / mb_w); done++; }
    if (sl->disable_deblocking_idc != 1) {
      dbk_ctx c = { sl, cur };
      done = 0;
      for (int xy = first; xy < n && done < sl->num_mbs; xy++)
        if (pd->mbs[xy].mb_type != LH264_MB_NONE && pd->mbs[xy].slice_idx == s) { deblock_mb (&c, pd->mbs, mb_w, xy % mb_w, xy / mb_w); done++; }
    }
  }
  return LH264_OK;
}

/* kernel-level entry points for tests */
void ora_idct4x4_add (uint8_t* pix, int stride, const int16_t* coef) { idct4x4_add (pix, stride, coef); }
void ora_idct8x8_add (uint8_t* pix, int stride, const int16_t* coef) { idct8x8_add (pix, stride, coef); }
void ora_expand_plane (const uint8_t* in, int w, int h, int pad, uint8_t* out) {
  /* ExpandPictureLuma_c / ExpandPictureChroma_c, codec/common/src/expand_pic.cpp:37-115 */
  int ow = w + 2 * pad;
  for (int y = -pad; y < h + pad; y++) for (int x = -pad; x < w + pad; x++)
      out[(size_t) (y + pad) * ow + x + pad] = in[(size_t)clip3 (y, 0, h - 1) * w + clip3 (x, 0, w - 1)];
}
int ora_sizeof_mb_rec (void) { return (int)sizeof (lh264_mb_rec); }
int ora_sizeof_slice_rec (void) { return (int)sizeof (lh264_slice_rec); }
