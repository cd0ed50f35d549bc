// recon_kernels.cuh — building blocks of the sm_100a reconstruction kernels.  Included by lh264_shim.cu.
//
// Shared device code of the reconstruction kernels: the MB staging record, dequant + inverse
// transforms of one coefficient block, the intra 4x4 / 8x8 predictors, the deblocking line filters,
// output packing and the kernel-level test entry points.  The persistent row kernels that drive
// them are in recon_rows.cuh.
#pragma once
#include "recon_device.cuh"

namespace lh264 {

struct PicDev {
  const lh264_mb_rec* mbs;
  const int16_t* coefs;
  const lh264_slice_rec* slices;
  uint8_t* base;                  // picture pool of the stream
  unsigned long long slot_stride; // bytes per picture slot
  int cur_slot, n_slots, mb_w, mb_h;
  uint8_t* lines;                 // per-stream line buffer: unfiltered bottom row of every MB row (mb_h x 2 x mb_w*16 bytes), decode_rows.cuh
  int n_slices, n_coef_blocks;    // bounds of `slices` / `coefs`, for record validation on the device
};

__device__ __forceinline__ uint8_t* plane_ptr (const PicDev& p, int slot, int plane) {
  const size_t ysz = (size_t)p.mb_w * p.mb_h * 256;
  return p.base + (size_t)slot * p.slot_stride + (plane == 0 ? 0 : (plane == 1 ? ysz : ysz + (ysz >> 2)));
}

// ---- shared staging of the MB record and its coefficient levels ------------------------------------------
struct MbShared {
  lh264_mb_rec mb;          // 128 B
  int16_t lev[384];         // dense quantised levels, 24 blocks x 16
  int16_t res[384];         // residual: luma 16x16 raster, then Cb 8x8, Cr 8x8
  uint8_t pred[384];        // prediction (same layout)
};

// Residual of coefficient block k (0..15 luma in z order, 16..19 Cb, 20..23 Cr) into s.res.
// Dequantisation as fused into the reference parse (parse_mb_syn_cavlc.cpp:938-1005,:1088-1095),
// DC transforms WelsLumaDcDequantIdct / WelsChromaDcIdct (decode_slice.cpp:271-311,:375-396).
__device__ void residual_block (MbShared& s, int k) {
  const bool i16 = s.mb.mb_type == LH264_MB_I16x16;
  // A block none of whose levels is present has a zero residual (the reference skips IdctResAddPred for it, gated by
  // pNzc || coef[0], decode_slice.cpp:237-243 / rec_mb.cpp:206,567): write the zeros and skip dequant + transform.  Most inter
  // MBs carry one or two coded blocks out of 24.
  {
    const uint32_t mask = s.mb.mb_type == LH264_MB_IPCM ? 0xffffffu : s.mb.coded_mask;
    bool zero;
    if (k < 16) zero = i16 ? (mask & 0xffffu) == 0 : ((s.mb.flags & LH264_MBF_TRANSFORM8x8) ? ((mask >> (k & ~3)) & 15u) == 0 : ((mask >> k) & 1u) == 0);
    else zero = ((mask >> (16 + 4 * ((k - 16) >> 2))) & 15u) == 0;
    if (zero) {
      if (k < 16) {
        if ((s.mb.flags & LH264_MBF_TRANSFORM8x8) && (k & 3)) return;       // the quadrant's first lane clears all of it
        const int n = (s.mb.flags & LH264_MBF_TRANSFORM8x8) ? 8 : 4;
        int16_t* res = s.res + blk_y (k) * 16 + blk_x (k);
        for (int y = 0; y < n; y++) for (int x = 0; x < n; x += 4) *reinterpret_cast<uint2*> (res + 16 * y + x) = make_uint2 (0u, 0u);
      } else {
        const int p = (k - 16) >> 2, b = (k - 16) & 3;
        int16_t* res = s.res + 256 + 64 * p + (b >> 1) * 32 + (b & 1) * 4;
#pragma unroll
        for (int y = 0; y < 4; y++) *reinterpret_cast<uint2*> (res + 8 * y) = make_uint2 (0u, 0u);
      }
      return;
    }
  }
  if (k < 16) {
    const int qp = s.mb.qp_y;
    if (s.mb.flags & LH264_MBF_TRANSFORM8x8) {
      if (k & 3) return;
      const int16_t* lev = s.lev + 16 * k;
      int t[64];
      const int sh = qp / 6;
#pragma unroll
      for (int i = 0; i < 8; i++) {
        int p[8], o[8];
#pragma unroll
        for (int j = 0; j < 8; j++) {
          const int v = lev[8 * i + j] * dq8 (qp, 8 * i + j);
          p[j] = S16 (qp >= 36 ? (v << (sh - 6)) : ((v + (1 << (5 - sh))) >> (6 - sh)));
        }
        idct8_1d (p, o);
#pragma unroll
        for (int j = 0; j < 8; j++) t[8 * i + j] = o[j];
      }
      int16_t* res = s.res + blk_y (k) * 16 + blk_x (k);
#pragma unroll
      for (int i = 0; i < 8; i++) {
        int p[8], o[8];
#pragma unroll
        for (int j = 0; j < 8; j++) p[j] = t[8 * j + i];
        idct8_1d (p, o);
#pragma unroll
        for (int j = 0; j < 8; j++) res[16 * j + i] = (int16_t) ((32 + o[j]) >> 6);
      }
      return;
    }
    int c[16];
    const int16_t* lev = s.lev + 16 * k;
#pragma unroll
    for (int j = 0; j < 16; j++) c[j] = S16 (lev[j] * dq4 (qp, j));
    if (i16) {
      // 4x4 Hadamard over the 16 DC levels, then ((x * dq[qp][0] + 2) >> 2)
      const int qmul = dq4 (qp, 0);
      int m[16], t[16];
      // DC matrix row i / col j = block yoff[i] + xoff[j], yoff = {0,2,8,10}, xoff = {0,1,4,5}
#pragma unroll
      for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) m[4 * i + j] = s.lev[16 * (((i & 1) << 1) + ((i & 2) << 2) + (j & 1) + ((j & 2) << 1))];
#pragma unroll
      for (int i = 0; i < 4; i++) {
        const int a = m[4 * i], b = m[4 * i + 1], cc = m[4 * i + 2], d = m[4 * i + 3];
        const int z0 = a + cc, z1 = a - cc, z2 = b - d, z3 = b + d;
        t[4 * i] = z0 + z3; t[4 * i + 1] = z1 + z2; t[4 * i + 2] = z1 - z2; t[4 * i + 3] = z0 - z3;
      }
      // this block's position in the DC matrix: row from bits (1,3) of k, col from bits (0,2)
      const int col = (k & 1) + ((k & 4) >> 1), row = ((k & 2) >> 1) + ((k & 8) >> 2);
      const int z0 = t[col] + t[8 + col], z1 = t[col] - t[8 + col], z2 = t[4 + col] - t[12 + col], z3 = t[4 + col] + t[12 + col];
      const int v = row == 0 ? z0 + z3 : (row == 1 ? z1 + z2 : (row == 2 ? z1 - z2 : z0 - z3));
      c[0] = S16 ((v * qmul + 2) >> 2);
    }
    idct4x4_res (c, s.res + blk_y (k) * 16 + blk_x (k), 16);
  } else {
    const int p = (k - 16) >> 2, b = (k - 16) & 3, qp = s.mb.qp_c[p];
    const int16_t* lev = s.lev + 16 * k;
    int c[16];
#pragma unroll
    for (int j = 1; j < 16; j++) c[j] = S16 (lev[j] * dq4 (qp, j));
    const int q0 = dq4 (qp, 0);
    const int16_t* dc = s.lev + 16 * (16 + 4 * p);
    int a = S16 (dc[0] * q0), bb = S16 (dc[16] * q0), cc = S16 (dc[32] * q0), d = S16 (dc[48] * q0);
    const int e = a - bb; a += bb; bb = cc - d; cc += d;
    c[0] = S16 (b == 0 ? (a + cc) >> 1 : (b == 1 ? (e + bb) >> 1 : (b == 2 ? (a - cc) >> 1 : (e - bb) >> 1)));
    idct4x4_res (c, s.res + 256 + 64 * p + (b >> 1) * 32 + (b & 1) * 4, 8);
  }
}

// ---- the same with scaling lists (High profile), WelsCalcDeqCoeffScalingList decode_slice.cpp:1231-1277 -----------------------
// multiplier of a level = list[pos] * LevelScale (qp, pos); the reference fills its tables for qp 0..50 only, so qp 51 gives 0
__device__ __forceinline__ int sl_dq4 (const lh264_slice_rec* sl, int list, int qp, int pos) {
  return qp < 51 ? ((int)__ldg (&sl->scale4[list][pos]) * dq4 (qp, pos)) & 0xffff : 0;      // the reference's tables are uint16_t
}
__device__ __forceinline__ int sl_dq8 (const lh264_slice_rec* sl, int list, int qp, int pos) {
  return qp < 51 ? (int)__ldg (&sl->scale8[list][pos]) * (dq8 (qp % 6, pos) >> 4) : 0;
}
// 8x8 multiplier of a CAVLC slice: element base + qp * 64 + pos of the reference's flat tables (LH264_SLF_SCALING_CAVLC)
__device__ __forceinline__ int sl_dq8_cavlc (const lh264_slice_rec* sl, bool intra, int qp, int pos) {
  const int n = (intra ? 0 : 3 * 52 * 16) + qp * 64 + pos;
  if (n < 6 * 52 * 16) return sl_dq4 (sl, n / (52 * 16), (n % (52 * 16)) >> 4, n & 15);
  const int m = n - 6 * 52 * 16;
  return sl_dq8 (sl, m / (52 * 64), (m % (52 * 64)) >> 6, m & 63);
}
// residual_block() for a slice with LH264_SLF_SCALING.  Cold path: kept out of line.
#ifdef LH264_CPU_EMU
#define LH264_COLD __attribute__ ((noinline))
#else
#define LH264_COLD __noinline__
#endif
__device__ LH264_COLD void residual_block_scaled (MbShared& s, int k, const lh264_slice_rec* sl) {
  const bool i16 = s.mb.mb_type == LH264_MB_I16x16, intra = LH264_MB_IS_INTRA (s.mb.mb_type);
  if (s.mb.mb_type == LH264_MB_IPCM) { residual_block (s, k); return; }
  if (k < 16) {
    const int qp = s.mb.qp_y;
    if (s.mb.flags & LH264_MBF_TRANSFORM8x8) {
      if (k & 3) return;
      int16_t* res = s.res + blk_y (k) * 16 + blk_x (k);
      if (((s.mb.coded_mask >> k) & 15u) == 0) {
        for (int y = 0; y < 8; y++) for (int x = 0; x < 8; x += 4) *reinterpret_cast<uint2*> (res + 16 * y + x) = make_uint2 (0u, 0u);
        return;
      }
      const bool cavlc = (__ldg (&sl->scaling_flags) & LH264_SLF_SCALING_CAVLC) != 0;
      const int16_t* lev = s.lev + 16 * k;
      int t[64];
      const int sh = qp / 6;
#pragma unroll 1
      for (int i = 0; i < 8; i++) {
        int p[8], o[8];
        for (int j = 0; j < 8; j++) {
          const int m = cavlc ? sl_dq8_cavlc (sl, intra, qp, 8 * i + j) : sl_dq8 (sl, intra ? 0 : 1, qp, 8 * i + j);
          const int v = lev[8 * i + j] * m;
          p[j] = S16 (qp >= 36 ? (v << (sh - 6)) : ((v + (1 << (5 - sh))) >> (6 - sh)));
        }
        idct8_1d (p, o);
        for (int j = 0; j < 8; j++) t[8 * i + j] = o[j];
      }
#pragma unroll 1
      for (int i = 0; i < 8; i++) {
        int p[8], o[8];
        for (int j = 0; j < 8; j++) p[j] = t[8 * j + i];
        idct8_1d (p, o);
        for (int j = 0; j < 8; j++) res[16 * j + i] = (int16_t) ((32 + o[j]) >> 6);
      }
      return;
    }
    const uint32_t mask = s.mb.coded_mask;
    if (i16 ? (mask & 0xffffu) == 0 : ((mask >> k) & 1u) == 0) { residual_block (s, k); return; }     // writes the zeros
    const int list = intra ? 0 : 3;
    int c[16];
    const int16_t* lev = s.lev + 16 * k;
    for (int j = 0; j < 16; j++) c[j] = S16 ((lev[j] * sl_dq4 (sl, list, qp, j)) >> 4);
    if (i16) {                                    // WelsLumaDcDequantIdct with kiQMul = table[0] >> 4 (decode_slice.cpp:272)
      const int qmul = sl_dq4 (sl, 0, qp, 0) >> 4;
      int m[16], t[16];
      for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++) m[4 * i + j] = s.lev[16 * (((i & 1) << 1) + ((i & 2) << 2) + (j & 1) + ((j & 2) << 1))];
      for (int i = 0; i < 4; i++) {
        const int a = m[4 * i], b = m[4 * i + 1], cc = m[4 * i + 2], d = m[4 * i + 3];
        const int z0 = a + cc, z1 = a - cc, z2 = b - d, z3 = b + d;
        t[4 * i] = z0 + z3; t[4 * i + 1] = z1 + z2; t[4 * i + 2] = z1 - z2; t[4 * i + 3] = z0 - z3;
      }
      const int col = (k & 1) + ((k & 4) >> 1), row = ((k & 2) >> 1) + ((k & 8) >> 2);
      const int z0 = t[col] + t[8 + col], z1 = t[col] - t[8 + col], z2 = t[4 + col] - t[12 + col], z3 = t[4 + col] + t[12 + col];
      const int v = row == 0 ? z0 + z3 : (row == 1 ? z1 + z2 : (row == 2 ? z1 - z2 : z0 - z3));
      c[0] = S16 ((v * qmul + 2) >> 2);
    }
    idct4x4_res (c, s.res + blk_y (k) * 16 + blk_x (k), 16);
  } else {
    const int p = (k - 16) >> 2, b = (k - 16) & 3, qp = s.mb.qp_c[p];
    if (((s.mb.coded_mask >> (16 + 4 * p)) & 15u) == 0) { residual_block (s, k); return; }
    const int list = (intra ? 1 : 4) + p;
    const int16_t* lev = s.lev + 16 * k;
    int c[16];
    for (int j = 1; j < 16; j++) c[j] = S16 ((lev[j] * sl_dq4 (sl, list, qp, j)) >> 4);
    const int q0 = sl_dq4 (sl, list, qp, 0);
    const int16_t* dc = s.lev + 16 * (16 + 4 * p);
    int a = S16 ((dc[0] * q0) >> 4), bb = S16 ((dc[16] * q0) >> 4), cc = S16 ((dc[32] * q0) >> 4), d = S16 ((dc[48] * q0) >> 4);
    const int e = a - bb; a += bb; bb = cc - d; cc += d;
    c[0] = S16 (b == 0 ? (a + cc) >> 1 : (b == 1 ? (e + bb) >> 1 : (b == 2 ? (a - cc) >> 1 : (e - bb) >> 1)));
    idct4x4_res (c, s.res + 256 + 64 * p + (b >> 1) * 32 + (b & 1) * 4, 8);
  }
}

// =====================================================================================================
// Intra wavefront.  WelsMbIntraPredictionConstruction (decode_slice.cpp:313-334), RecI4x4Mb / RecI8x8Mb /
// RecI16x16Mb / RecChroma (rec_mb.cpp:64-219,:547-575), predictors get_intra_predictor.cpp:54-1155.
// grid = (longest anti-diagonal, pictures with intra MBs), one launch per wave.
// =====================================================================================================
#define TS 32   // luma tile row stride; tile(x,y) with x in [-1,24), y in [-1,16)
#define TILE(x, y) tile[((y) + 1) * TS + (x) + 1]
#define F3(a, b, c) (((a) + 2 * (b) + (c) + 2) >> 2)
#define F2(a, b) (((a) + (b) + 1) >> 1)

// one sample (x,y) of a 4x4 predictor at tile position (X0,Y0); modes as resolved by the host
__device__ int pred4x4_sample (const uint8_t* tile, int X0, int Y0, int mode, int x, int y) {
#define T(i) ((int)TILE (X0 + (((mode == LH264_I4_DDL_TOP || mode == LH264_I4_VL_TOP) && (i) > 3) ? 3 : (i)), Y0 - 1))
#define L(i) ((int)TILE (X0 - 1, Y0 + (i)))
#define E(k) ((k) < 4 ? L (3 - (k)) : ((k) == 4 ? (int)TILE (X0 - 1, Y0 - 1) : T ((k) - 5)))   /* l3..l0, tl, t0..t3 */
  switch (mode) {
  case LH264_I4_V: return T (x);
  case LH264_I4_H: return L (y);
  case LH264_I4_DC: return (T (0) + T (1) + T (2) + T (3) + L (0) + L (1) + L (2) + L (3) + 4) >> 3;
  case LH264_I4_DC_L: return (L (0) + L (1) + L (2) + L (3) + 2) >> 2;
  case LH264_I4_DC_T: return (T (0) + T (1) + T (2) + T (3) + 2) >> 2;
  case LH264_I4_DC_128: return 128;
  case LH264_I4_DDL: case LH264_I4_DDL_TOP:
    return (x == 3 && y == 3) ? F3 (T (6), T (7), T (7)) : F3 (T (x + y), T (x + y + 1), T (x + y + 2));
  case LH264_I4_DDR: { const int k = 4 + x - y; return F3 (E (k - 1), E (k), E (k + 1)); }
  case LH264_I4_VR: {
    const int z = 2 * x - y, k = 4 + x - (y >> 1);
    if (z >= 0 && ! (z & 1)) return F2 (E (k), E (k + 1));
    if (z >= 0) return F3 (E (k - 1), E (k), E (k + 1));
    if (z == -1) return F3 (E (3), E (4), E (5));
    return F3 (E (6 - y), E (5 - y), E (4 - y));
  }
  case LH264_I4_HD: {
    const int z = 2 * y - x, k = 4 - (y - (x >> 1));
    if (z >= 0 && ! (z & 1)) return F2 (E (k), E (k - 1));
    if (z >= 0) return F3 (E (k + 1), E (k), E (k - 1));
    if (z == -1) return F3 (E (3), E (4), E (5));
    return F3 (E (4 + x), E (3 + x), E (2 + x));
  }
  case LH264_I4_VL: case LH264_I4_VL_TOP: {
    const int k = x + (y >> 1);
    return (y & 1) ? F3 (T (k), T (k + 1), T (k + 2)) : F2 (T (k), T (k + 1));
  }
  default: {   // HU
    const int z = x + 2 * y;
    if (z > 5) return L (3);
    if (z == 5) return F3 (L (2), L (3), L (3));
    return (z & 1) ? F3 (L (z >> 1), L ((z >> 1) + 1), L ((z >> 1) + 2)) : F2 (L (z >> 1), L ((z >> 1) + 1));
  }
  }
#undef T
#undef L
#undef E
}

// one sample of an 8x8 predictor (reference-sample filtering included), get_intra_predictor.cpp:382-880
__device__ int pred8x8_sample (const uint8_t* tile, int X0, int Y0, int mode, int has_tl, int has_tr, int x, int y) {
  const bool use_top = ! (mode == LH264_I4_H || mode == LH264_I4_HU || mode == LH264_I4_DC_L || mode == LH264_I4_DC_128);
  const bool use_left = ! (mode == LH264_I4_V || mode == LH264_I4_DDL || mode == LH264_I4_VL || mode == LH264_I4_DC_T ||
                           mode == LH264_I4_DC_128 || mode == LH264_I4_DDL_TOP || mode == LH264_I4_VL_TOP);
  int ft[16], fl[8], ftl = 0;
  if (mode == LH264_I4_DDR || mode == LH264_I4_VR || mode == LH264_I4_HD) {
    has_tl = 1;
    ftl = (TILE (X0 - 1, Y0) + 2 * TILE (X0 - 1, Y0 - 1) + TILE (X0, Y0 - 1) + 2) >> 2;
  }
#pragma unroll
  for (int i = 0; i < 16; i++) ft[i] = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) fl[i] = 0;
  if (use_top) {
    int t[16];
    const bool real_tr = (mode == LH264_I4_DDL || mode == LH264_I4_VL) ? true :
                         ((mode == LH264_I4_DDL_TOP || mode == LH264_I4_VL_TOP) ? false : (has_tr != 0));
#pragma unroll
    for (int i = 0; i < 8; i++) t[i] = TILE (X0 + i, Y0 - 1);
#pragma unroll
    for (int i = 8; i < 16; i++) t[i] = real_tr ? (int)TILE (X0 + i, Y0 - 1) : t[7];
    ft[0] = has_tl ? ((TILE (X0 - 1, Y0 - 1) + 2 * t[0] + t[1] + 2) >> 2) : ((3 * t[0] + t[1] + 2) >> 2);
#pragma unroll
    for (int i = 1; i < 15; i++) ft[i] = (t[i - 1] + 2 * t[i] + t[i + 1] + 2) >> 2;
    ft[15] = (t[14] + 3 * t[15] + 2) >> 2;
  }
  if (use_left) {
    int l[8];
#pragma unroll
    for (int i = 0; i < 8; i++) l[i] = TILE (X0 - 1, Y0 + i);
    fl[0] = has_tl ? ((TILE (X0 - 1, Y0 - 1) + 2 * l[0] + l[1] + 2) >> 2) : ((3 * l[0] + l[1] + 2) >> 2);
#pragma unroll
    for (int i = 1; i < 7; i++) fl[i] = (l[i - 1] + 2 * l[i] + l[i + 1] + 2) >> 2;
    fl[7] = (l[6] + 3 * l[7] + 2) >> 2;
  }
#define E8(k) ((k) < 8 ? fl[7 - (k)] : ((k) == 8 ? ftl : ft[(k) - 9]))
  switch (mode) {
  case LH264_I4_V: return ft[x];
  case LH264_I4_H: return fl[y];
  case LH264_I4_DC: case LH264_I4_DC_L: case LH264_I4_DC_T: case LH264_I4_DC_128: {
    int st = 0, sl = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) { st += ft[i]; sl += fl[i]; }
    if (mode == LH264_I4_DC) return (st + sl + 8) >> 4;
    if (mode == LH264_I4_DC_L) return (sl + 4) >> 3;
    if (mode == LH264_I4_DC_T) return (st + 4) >> 3;
    return 128;
  }
  case LH264_I4_DDL: case LH264_I4_DDL_TOP:
    return (x == 7 && y == 7) ? ((ft[14] + 3 * ft[15] + 2) >> 2) : F3 (ft[x + y], ft[x + y + 1], ft[x + y + 2]);
  case LH264_I4_DDR: { const int k = 8 + x - y; return F3 (E8 (k - 1), E8 (k), E8 (k + 1)); }
  case LH264_I4_VR: {
    const int z = 2 * x - y, k = 8 + x - (y >> 1);
    if (z >= 0 && ! (z & 1)) return F2 (E8 (k), E8 (k + 1));
    if (z >= 0) return F3 (E8 (k - 1), E8 (k), E8 (k + 1));
    if (z == -1) return F3 (E8 (7), E8 (8), E8 (9));
    const int q = 8 - y + 2 * x;
    return F3 (E8 (q + 2), E8 (q + 1), E8 (q));
  }
  case LH264_I4_HD: {
    const int z = 2 * y - x, k = 8 - (y - (x >> 1));
    if (z >= 0 && ! (z & 1)) return F2 (E8 (k), E8 (k - 1));
    if (z >= 0) return F3 (E8 (k + 1), E8 (k), E8 (k - 1));
    if (z == -1) return F3 (E8 (7), E8 (8), E8 (9));
    const int q = 8 + x - 2 * y;
    return F3 (E8 (q), E8 (q - 1), E8 (q - 2));
  }
  case LH264_I4_VL: case LH264_I4_VL_TOP: {
    const int k = x + (y >> 1);
    return (y & 1) ? F3 (ft[k], ft[k + 1], ft[k + 2]) : F2 (ft[k], ft[k + 1]);
  }
  default: {   // HU
    const int z = x + 2 * y;
    if (z > 13) return fl[7];
    if (z == 13) return (fl[6] + 3 * fl[7] + 2) >> 2;
    return (z & 1) ? F3 (fl[z >> 1], fl[(z >> 1) + 1], fl[(z >> 1) + 2]) : F2 (fl[z >> 1], fl[(z >> 1) + 1]);
  }
  }
#undef E8
}

#undef F2
#undef F3

// =====================================================================================================
// Deblocking wavefront.  WelsDeblockingFilterSlice / WelsDeblockingMb (deblocking.cpp:815-933), bS
// derivation (:160-353), edge filters DeblockLuma/Chroma{Lt4,Eq4}_c (deblocking_common.cpp:5-246).
// One warp per MB; 4 warps per CTA.  grid = (ceil(longest diagonal / 4), pictures).
// =====================================================================================================
__device__ __forceinline__ int nzb (uint32_t nz, int r) { return (nz >> r) & 1; }
__device__ __forceinline__ int nz8 (uint32_t nz, int r) { const int b = (r & 8) + (r & 2); return ((nz >> b) | (nz >> (b + 1)) | (nz >> (b + 4)) | (nz >> (b + 5))) & 1; }
__device__ __forceinline__ int quad_of (int r) { return ((r >> 3) << 1) + ((r & 3) >> 1); }
__device__ __forceinline__ int mv_differs (const lh264_mb_rec* a, int ra, const lh264_mb_rec* b, int rb) {
  return (a->ref_idx[quad_of (ra)] != b->ref_idx[quad_of (rb)]) || abs (a->mv[ra][0] - b->mv[rb][0]) >= 4 || abs (a->mv[ra][1] - b->mv[rb][1]) >= 4;
}

__device__ __forceinline__ void luma_line (uint8_t* p, int sx, int bS, int ia, int alpha, int beta) {
  const int p0 = p[-sx], p1 = p[-2 * sx], p2 = p[-3 * sx], q0 = p[0], q1 = p[sx], q2 = p[2 * sx];
  const int d = abs (p0 - q0);
  if (! (d < alpha && abs (p1 - p0) < beta && abs (q1 - q0) < beta)) return;
  if (bS == 4) {                              // DeblockLumaEq4_c
    if (d < ((alpha >> 2) + 2)) {
      if (abs (p2 - p0) < beta) {
        const int p3 = p[-4 * sx];
        p[-sx] = (uint8_t) ((p2 + 2 * p1 + 2 * p0 + 2 * q0 + q1 + 4) >> 3);
        p[-2 * sx] = (uint8_t) ((p2 + p1 + p0 + q0 + 2) >> 2);
        p[-3 * sx] = (uint8_t) ((2 * p3 + 3 * p2 + p1 + p0 + q0 + 4) >> 3);
      } else p[-sx] = (uint8_t) ((2 * p1 + p0 + q1 + 2) >> 2);
      if (abs (q2 - q0) < beta) {
        const int q3 = p[3 * sx];
        p[0] = (uint8_t) ((p1 + 2 * p0 + 2 * q0 + 2 * q1 + q2 + 4) >> 3);
        p[sx] = (uint8_t) ((p0 + q0 + q1 + q2 + 2) >> 2);
        p[2 * sx] = (uint8_t) ((2 * q3 + 3 * q2 + q1 + q0 + p0 + 4) >> 3);
      } else p[0] = (uint8_t) ((2 * q1 + q0 + p1 + 2) >> 2);
    } else {
      p[-sx] = (uint8_t) ((2 * p1 + p0 + q1 + 2) >> 2);
      p[0] = (uint8_t) ((2 * q1 + q0 + p1 + 2) >> 2);
    }
  } else {                                    // DeblockLumaLt4_c
    const int tc0 = c_tc0[ia][bS];
    int tc = tc0;
    if (abs (p2 - p0) < beta) { p[-2 * sx] = (uint8_t) (p1 + clip3i ((p2 + ((p0 + q0 + 1) >> 1) - (p1 << 1)) >> 1, -tc0, tc0)); tc++; }
    if (abs (q2 - q0) < beta) { p[sx] = (uint8_t) (q1 + clip3i ((q2 + ((p0 + q0 + 1) >> 1) - (q1 << 1)) >> 1, -tc0, tc0)); tc++; }
    const int dl = clip3i ((((q0 - p0) << 2) + (p1 - q1) + 4) >> 3, -tc, tc);
    p[-sx] = (uint8_t)clip255 (p0 + dl); p[0] = (uint8_t)clip255 (q0 - dl);
  }
}
__device__ __forceinline__ void chroma_line (uint8_t* p, int sx, int bS, int ia, int alpha, int beta) {
  const int p0 = p[-sx], p1 = p[-2 * sx], q0 = p[0], q1 = p[sx];
  if (! (abs (p0 - q0) < alpha && abs (p1 - p0) < beta && abs (q1 - q0) < beta)) return;
  if (bS == 4) {                              // DeblockChromaEq4_c
    p[-sx] = (uint8_t) ((2 * p1 + p0 + q1 + 2) >> 2); p[0] = (uint8_t) ((2 * q1 + q0 + p1 + 2) >> 2);
  } else {                                    // DeblockChromaLt4_c, tc = tc0 + 1
    const int tc = c_tc0[ia][bS] + 1;
    const int dl = clip3i ((((q0 - p0) << 2) + (p1 - q1) + 4) >> 3, -tc, tc);
    p[-sx] = (uint8_t)clip255 (p0 + dl); p[0] = (uint8_t)clip255 (q0 - dl);
  }
}

// =====================================================================================================
// Output packing: cropped I420 of every picture of the batch into one contiguous buffer (one D2H).
// DecodeFrameConstruction (decoder_core.cpp:56-217) hands out cropped plane pointers; here the
// cropped planes are gathered so the host gets them with a single copy.
// =====================================================================================================
struct PackDev { const uint8_t* src; int pitch; int width; int height; unsigned long long dst_off; };
__global__ void k_pack (const PackDev* __restrict__ jobs, uint8_t* __restrict__ out) {
  const PackDev j = jobs[blockIdx.y];
  for (int row = blockIdx.x; row < j.height; row += gridDim.x) {
    const uint8_t* s = j.src + (size_t)row * j.pitch;
    uint8_t* d = out + j.dst_off + (size_t)row * j.width;
    if ((((uintptr_t)s | (uintptr_t)d | (unsigned)j.width) & 15) == 0) {
      for (int x = threadIdx.x; x < (j.width >> 4); x += blockDim.x) reinterpret_cast<uint4*> (d)[x] = __ldg (reinterpret_cast<const uint4*> (s) + x);
    } else {
      for (int x = threadIdx.x; x < j.width; x += blockDim.x) d[x] = s[x];
    }
  }
}

// ---- kernel-level entry points -------------------------------------------------------------------------
__global__ void k_test_idct4 (uint8_t* pix, const int16_t* coef, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int c[16]; int16_t r[16];
  for (int j = 0; j < 16; j++) c[j] = coef[16 * i + j];
  idct4x4_res (c, r, 4);
  for (int j = 0; j < 16; j++) pix[16 * i + j] = (uint8_t)clip255 (pix[16 * i + j] + r[j]);
}
__global__ void k_test_idct8 (uint8_t* pix, const int16_t* coef, int n) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= n) return;
  int t[64];
  for (int i = 0; i < 8; i++) { int p[8], o[8]; for (int j = 0; j < 8; j++) p[j] = coef[64 * b + 8 * i + j]; idct8_1d (p, o); for (int j = 0; j < 8; j++) t[8 * i + j] = o[j]; }
  for (int i = 0; i < 8; i++) {
    int p[8], o[8];
    for (int j = 0; j < 8; j++) p[j] = t[8 * j + i];
    idct8_1d (p, o);
    for (int j = 0; j < 8; j++) pix[64 * b + 8 * j + i] = (uint8_t)clip255 (pix[64 * b + 8 * j + i] + ((32 + o[j]) >> 6));
  }
}
// blocks: int32[n][6] = x, y, mvx, mvy, bw, bh ; one CTA of 256 threads per block, out u8[n][256]
__global__ void k_test_mc_luma (const uint8_t* ref, int w, int h, const int32_t* blocks, uint8_t* out) {
  __shared__ uint8_t win[21 * 24];
  const int32_t* b = blocks + 6 * blockIdx.x;
  const int ix = b[0] + (b[2] >> 2) - 2, iy = b[1] + (b[3] >> 2) - 2, bw = b[4], bh = b[5];
  for (int k = threadIdx.x; k < 21 * 21; k += blockDim.x) {
    const int wy = k / 21, wx = k % 21;
    win[wy * 24 + wx] = ref[(size_t)clip3i (iy + wy, 0, h - 1) * w + clip3i (ix + wx, 0, w - 1)];
  }
  __syncthreads();
  const int x = threadIdx.x & 15, y = threadIdx.x >> 4;
  if (x < bw && y < bh) out[256 * blockIdx.x + y * 16 + x] = (uint8_t)mc_luma_sample (win + (2 + y) * 24 + 2 + x, 24, ((b[2] & 3) << 2) | (b[3] & 3));
}
__global__ void k_test_mc_chroma (const uint8_t* ref, int w, int h, const int32_t* blocks, uint8_t* out) {
  const int32_t* b = blocks + 6 * blockIdx.x;
  const int cx = b[0] + (b[2] >> 3), cy = b[1] + (b[3] >> 3), dx = b[2] & 7, dy = b[3] & 7, bw = b[4], bh = b[5];
  const int x = threadIdx.x & 15, y = threadIdx.x >> 4;
  if (x >= bw || y >= bh) return;
  const int A = (8 - dx) * (8 - dy), B = dx * (8 - dy), C = (8 - dx) * dy, D = dx * dy;
#define RP(xx, yy) ref[(size_t)clip3i (yy, 0, h - 1) * w + clip3i (xx, 0, w - 1)]
  out[256 * blockIdx.x + y * 16 + x] = (uint8_t) ((A * RP (cx + x, cy + y) + B * RP (cx + x + 1, cy + y) + C * RP (cx + x, cy + y + 1) + D * RP (cx + x + 1, cy + y + 1) + 32) >> 6);
#undef RP
}
__global__ void k_test_expand (const uint8_t* in, int w, int h, int pad, uint8_t* out) {
  const int ow = w + 2 * pad, oh = h + 2 * pad;
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < ow * oh; k += gridDim.x * blockDim.x) {
    const int y = k / ow - pad, x = k % ow - pad;
    out[k] = in[(size_t)clip3i (y, 0, h - 1) * w + clip3i (x, 0, w - 1)];
  }
}

} // namespace lh264
