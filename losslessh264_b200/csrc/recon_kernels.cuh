// recon_kernels.cuh — the sm_100a reconstruction kernels (inter pass, intra wavefront, deblocking
// wavefront, output packing).  Included by lh264_shim.cu.
//
// Work decomposition per picture batch (one picture per independent stream):
//   k_inter    all inter MBs of all pictures, fully parallel, one 128-thread CTA per MB:
//              64 threads = 16 4x4 luma blocks x 4 rows (9x9 reference window per block staged in
//              shared memory with clamp-at-read borders), 32 threads = chroma 2x2 blocks,
//              24 threads = dequant + inverse transform of the 24 coefficient blocks.
//   k_intra    anti-diagonal wave d = x + 2y (left, top, top-left and top-right neighbours are
//              complete), one launch per wave over all pictures of the batch.
//   k_deblock  same wave order (the top-right MB filters before the current one in raster
//              order), one warp per MB: lanes 0-15 = luma lines, 16-31 = Cb/Cr lines.
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
};

__device__ __forceinline__ uint8_t* plane_ptr (const PicDev& p, int slot, int plane) {
  const size_t ysz = (size_t)p.mb_w * p.mb_h * 256;
  return p.base + (size_t)slot * p.slot_stride + (plane == 0 ? 0 : (plane == 1 ? ysz : ysz + (ysz >> 2)));
}

// wave d -> k-th MB on the anti-diagonal x + 2y = d; returns false when k is past its end
__device__ __forceinline__ bool wave_mb (int d, int k, int mb_w, int mb_h, int& x, int& y) {
  const int y_min = max (0, (d - (mb_w - 1) + 1) >> 1), y_max = min (mb_h - 1, d >> 1);
  y = y_min + k;
  x = d - 2 * y;
  return y <= y_max;
}

// ---- shared staging of the MB record and its coefficient levels ------------------------------------------
struct MbShared {
  lh264_mb_rec mb;          // 128 B
  int16_t lev[384];         // dense quantised levels, 24 blocks x 16
  int16_t res[384];         // residual: luma 16x16 raster, then Cb 8x8, Cr 8x8
  uint8_t pred[384];        // prediction (same layout)
};

__device__ __forceinline__ void load_mb_and_levels (MbShared& s, const lh264_mb_rec* m, const int16_t* coefs, int tid, int nthreads) {
  if (tid < 32) reinterpret_cast<uint32_t*> (&s.mb)[tid] = __ldg (reinterpret_cast<const uint32_t*> (m) + tid);
  __syncthreads();
  const uint32_t mask = s.mb.coded_mask;
  const uint32_t* src = reinterpret_cast<const uint32_t*> (coefs) + (size_t)s.mb.coef_off * 8;
  for (int w = tid; w < 192; w += nthreads) {
    const int blk = w >> 3;
    uint32_t v = 0;
    if ((mask >> blk) & 1) v = __ldg (src + __popc (mask & ((1u << blk) - 1)) * 8 + (w & 7));
    reinterpret_cast<uint32_t*> (s.lev)[w] = v;
  }
  __syncthreads();
}

// Residual of coefficient block k (0..15 luma in z order, 16..19 Cb, 20..23 Cr) into s.res.
// Dequantisation as fused into the reference parse (parse_mb_syn_cavlc.cpp:938-1005,:1088-1095),
// DC transforms WelsLumaDcDequantIdct / WelsChromaDcIdct (decode_slice.cpp:271-311,:375-396).
__device__ void residual_block (MbShared& s, int k) {
  const bool i16 = s.mb.mb_type == LH264_MB_I16x16;
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

// final add + clip + store of 4 horizontally adjacent samples (thread t < 96)
__device__ __forceinline__ void store_quad (const MbShared& s, const PicDev& pic, int mbx, int mby, int t, bool add_res) {
  const int W = pic.mb_w * 16;
  int idx;
  uint8_t* dst;
  if (t < 64) {
    const int y = t >> 2, x4 = (t & 3) << 2;
    idx = y * 16 + x4;
    dst = plane_ptr (pic, pic.cur_slot, 0) + (size_t) (mby * 16 + y) * W + mbx * 16 + x4;
  } else {
    const int c = t - 64, p = c >> 4, y = (c & 15) >> 1, x4 = (c & 1) << 2;
    idx = 256 + 64 * p + y * 8 + x4;
    dst = plane_ptr (pic, pic.cur_slot, 1 + p) + (size_t) (mby * 8 + y) * (W >> 1) + mbx * 8 + x4;
  }
  uchar4 o;
  if (add_res) {
    o.x = (uint8_t)clip255 (s.pred[idx] + s.res[idx]); o.y = (uint8_t)clip255 (s.pred[idx + 1] + s.res[idx + 1]);
    o.z = (uint8_t)clip255 (s.pred[idx + 2] + s.res[idx + 2]); o.w = (uint8_t)clip255 (s.pred[idx + 3] + s.res[idx + 3]);
  } else {
    o.x = s.pred[idx]; o.y = s.pred[idx + 1]; o.z = s.pred[idx + 2]; o.w = s.pred[idx + 3];
  }
  *reinterpret_cast<uchar4*> (dst) = o;
}

// =====================================================================================================
// Inter pass.  GetInterPred/BaseMC/WeightPrediction (rec_mb.cpp:247-545) + WelsMbInterSampleConstruction
// (decode_slice.cpp:208-250).  grid = (mbs of the largest picture, pictures).
// =====================================================================================================
#define WIN_STRIDE 12
__global__ void __launch_bounds__ (128) k_inter (const PicDev* __restrict__ pics) {
  __shared__ MbShared s;
  __shared__ uint8_t s_win[16][9 * WIN_STRIDE];
  const PicDev pic = pics[blockIdx.y];
  const int mb = blockIdx.x;
  if (mb >= pic.mb_w * pic.mb_h) return;
  const int mb_type = __ldg (&pic.mbs[mb].mb_type);
  if (!LH264_MB_IS_INTER (mb_type)) return;
  const int tid = threadIdx.x;
  load_mb_and_levels (s, pic.mbs + mb, pic.coefs, tid, 128);
  const int mbx = mb % pic.mb_w, mby = mb / pic.mb_w;
  const int W = pic.mb_w * 16, H = pic.mb_h * 16;
  const lh264_slice_rec* sl = pic.slices + s.mb.slice_idx;
  const bool has_res = s.mb.cbp != 0;

  if (tid < 96) {
    const bool luma = tid < 64;
    const int c = tid - 64;
    const int r = luma ? (tid >> 2) : (c & 15);            // raster 4x4 block this thread predicts for
    const int bx = r & 3, by = r >> 2, q = (by >> 1) * 2 + (bx >> 1);
    int ox, oy, bw, bh;
    part_geom (mb_type, s.mb.sub_mb_type[q], bx, by, ox, oy, bw, bh);
    const int refidx = s.mb.ref_idx[q] & 15;
    const int slot = clip3i (__ldg (&sl->ref_slot[refidx]), 0, pic.n_slots - 1);
    // BaseMC: clamp the partition's full-sample MV into the (virtual) 32-pixel border
    const int fx = clip3i (((mbx * 16 + ox) << 2) + s.mb.mv[r][0], (-32 + 2) * 4, (W + 32 - 19) << 2);
    const int fy = clip3i (((mby * 16 + oy) << 2) + s.mb.mv[r][1], (-32 + 2) * 4, (H + 32 - 19) << 2);
    const int wp = __ldg (&sl->weighted_pred);
    if (luma) {
      const int row = tid & 3;
      const uint8_t* ref = plane_ptr (pic, slot, 0);
      const int ix = (fx >> 2) + (bx * 4 - ox) - 2, iy = (fy >> 2) + (by * 4 - oy) - 2;
      uint8_t* win = s_win[r];
      for (int k = row; k < 81; k += 4) {
        const int wy = k / 9, wx = k - wy * 9;
        win[wy * WIN_STRIDE + wx] = __ldg (ref + (size_t)clip3i (iy + wy, 0, H - 1) * W + clip3i (ix + wx, 0, W - 1));
      }
      __syncwarp();
      const int frac = ((fx & 3) << 2) | (fy & 3);
      const uint8_t* w0 = win + (2 + row) * WIN_STRIDE + 2;
      uint8_t* out = s.pred + (by * 4 + row) * 16 + bx * 4;
      int lw = 0, lo = 0, ld = 0;
      if (wp) { lw = __ldg (&sl->luma_weight[refidx]); lo = __ldg (&sl->luma_offset[refidx]); ld = __ldg (&sl->luma_log2_denom); }
#pragma unroll
      for (int i = 0; i < 4; i++) {
        int v = mc_luma_sample (w0 + i, WIN_STRIDE, frac);
        if (wp) v = wp_sample (v, lw, lo, ld);
        out[i] = (uint8_t)v;
      }
    } else {
      // McChroma_c on the 2x2 chroma block that belongs to luma block r
      const int p = c >> 4, CW = W >> 1, CH = H >> 1;
      const uint8_t* ref = plane_ptr (pic, slot, 1 + p);
      const int cx = (fx >> 3) + ((bx * 4 - ox) >> 1), cy = (fy >> 3) + ((by * 4 - oy) >> 1);
      const int dx = fx & 7, dy = fy & 7;
      int v[3][3];
#pragma unroll
      for (int j = 0; j < 3; j++)
#pragma unroll
        for (int i = 0; i < 3; i++) v[j][i] = __ldg (ref + (size_t)clip3i (cy + j, 0, CH - 1) * CW + clip3i (cx + i, 0, CW - 1));
      const int A = (8 - dx) * (8 - dy), B = dx * (8 - dy), C = (8 - dx) * dy, D = dx * dy;
      int cw = 0, co = 0, cd = 0;
      // the reference weights only the top-left (bw>>2) x (bh>>2) samples of a partition's chroma (rec_mb.cpp:309-311)
      if (wp) { cw = __ldg (&sl->chroma_weight[refidx][p]); co = __ldg (&sl->chroma_offset[refidx][p]); cd = __ldg (&sl->chroma_log2_denom); }
#pragma unroll
      for (int j = 0; j < 2; j++)
#pragma unroll
        for (int i = 0; i < 2; i++) {
          int o = (A * v[j][i] + B * v[j][i + 1] + C * v[j + 1][i] + D * v[j + 1][i + 1] + 32) >> 6;
          if (wp && ((bx * 4 - ox) >> 1) + i < (bw >> 2) && ((by * 4 - oy) >> 1) + j < (bh >> 2)) o = wp_sample (o, cw, co, cd);
          s.pred[256 + 64 * p + (by * 2 + j) * 8 + bx * 2 + i] = (uint8_t)o;
        }
    }
  } else if (tid < 120 && has_res) {
    residual_block (s, tid - 96);
  }
  __syncthreads();
  if (tid < 96) store_quad (s, pic, mbx, mby, tid, has_res);
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

__global__ void __launch_bounds__ (128) k_intra (const PicDev* __restrict__ pics, const int* __restrict__ pic_ids, int wave) {
  __shared__ MbShared s;
  __shared__ uint8_t tile[17 * TS];
  __shared__ uint8_t cedge[2][2][9];      // [plane][0 = top incl. top-left at [0], 1 = left incl. top-left][9]
  const PicDev pic = pics[pic_ids[blockIdx.y]];
  int mbx, mby;
  if (!wave_mb (wave, blockIdx.x, pic.mb_w, pic.mb_h, mbx, mby)) return;
  const int mb = mby * pic.mb_w + mbx;
  const int mb_type = __ldg (&pic.mbs[mb].mb_type);
  if (!LH264_MB_IS_INTRA (mb_type)) return;
  const int tid = threadIdx.x;
  load_mb_and_levels (s, pic.mbs + mb, pic.coefs, tid, 128);
  const int W = pic.mb_w * 16, H = pic.mb_h * 16, CW = W >> 1;
  uint8_t* Y = plane_ptr (pic, pic.cur_slot, 0);

  if (mb_type == LH264_MB_IPCM) {           // samples travel in the coefficient pool
    for (int i = tid; i < 384; i += 128) s.pred[i] = (uint8_t)s.lev[i];
    __syncthreads();
    if (tid < 96) store_quad (s, pic, mbx, mby, tid, false);
    return;
  }

  // phase A: neighbour samples -> shared, residuals of all 24 blocks
  if (tid < 96) {
    // luma: top row x = -1..23 (25 samples), left column y = 0..15; coordinates clamped (unavailable
    // neighbours are never used by a resolved mode)
    if (tid < 25) {
      const int x = mbx * 16 - 1 + tid, y = mby * 16 - 1;
      tile[tid] = Y[(size_t)max (y, 0) * W + clip3i (x, 0, W - 1)];
    } else if (tid < 41) {
      const int y = mby * 16 + tid - 25, x = mbx * 16 - 1;
      tile[(tid - 25 + 1) * TS] = Y[(size_t)y * W + max (x, 0)];
    } else if (tid < 41 + 36) {
      const int c = tid - 41, p = c / 18, k = c % 18;        // k 0..8 top (x=-1..7), 9..17 left (y=-1..7)
      const uint8_t* C = plane_ptr (pic, pic.cur_slot, 1 + p);
      int x, y;
      if (k < 9) { x = mbx * 8 - 1 + k; y = mby * 8 - 1; } else { x = mbx * 8 - 1; y = mby * 8 - 1 + (k - 9); }
      cedge[p][k < 9 ? 0 : 1][k < 9 ? k : k - 9] = C[(size_t)max (y, 0) * CW + clip3i (x, 0, CW - 1)];
    }
  } else if (tid < 120) {
    residual_block (s, tid - 96);
  }
  __syncthreads();
  (void)H;

  // chroma prediction (threads 64..95), WelsIChromaPred*_c
  if (tid >= 64 && tid < 96) {
    const int c = tid - 64, p = c >> 4, y = (c & 15) >> 1, x4 = (c & 1) << 2;
    const uint8_t* top = &cedge[p][0][1];   // top[-1] = top-left
    const uint8_t* left = &cedge[p][1][1];
    const int mode = s.mb.chroma_mode;
    int v[4];
    if (mode == LH264_C_V) { for (int i = 0; i < 4; i++) v[i] = top[x4 + i]; }
    else if (mode == LH264_C_H) { for (int i = 0; i < 4; i++) v[i] = left[y]; }
    else if (mode == LH264_C_P) {
      int Hs = 0, Vs = 0;
#pragma unroll
      for (int i = 0; i < 4; i++) { Hs += (i + 1) * (top[4 + i] - top[2 - i]); Vs += (i + 1) * (left[4 + i] - left[2 - i]); }
      const int a = (left[7] + top[7]) << 4, b = (17 * Hs + 16) >> 5, cc = (17 * Vs + 16) >> 5;
      for (int i = 0; i < 4; i++) v[i] = clip255 ((a + b * (x4 + i - 3) + cc * (y - 3) + 16) >> 5);
    } else {
      int t0 = 0, t1 = 0, l0 = 0, l1 = 0, dc = 128;
#pragma unroll
      for (int i = 0; i < 4; i++) { t0 += top[i]; t1 += top[4 + i]; l0 += left[i]; l1 += left[4 + i]; }
      const int quad = (y >> 2) * 2 + (x4 >> 2);
      if (mode == LH264_C_DC) dc = quad == 0 ? (t0 + l0 + 4) >> 3 : (quad == 1 ? (t1 + 2) >> 2 : (quad == 2 ? (l1 + 2) >> 2 : (t1 + l1 + 4) >> 3));
      else if (mode == LH264_C_DC_T) dc = ((x4 ? t1 : t0) + 2) >> 2;
      else if (mode == LH264_C_DC_L) dc = ((y >> 2 ? l1 : l0) + 2) >> 2;
      for (int i = 0; i < 4; i++) v[i] = dc;
    }
    for (int i = 0; i < 4; i++) s.pred[256 + 64 * p + y * 8 + x4 + i] = (uint8_t)v[i];
  }

  if (mb_type == LH264_MB_I16x16) {
    if (tid < 64) {
      const int y = tid >> 2, x4 = (tid & 3) << 2, mode = s.mb.i16_mode;
      int v[4];
      if (mode == LH264_I16_V) { for (int i = 0; i < 4; i++) v[i] = TILE (x4 + i, -1); }
      else if (mode == LH264_I16_H) { for (int i = 0; i < 4; i++) v[i] = TILE (-1, y); }
      else if (mode == LH264_I16_P) {
        int Hs = 0, Vs = 0;
#pragma unroll
        for (int i = 0; i < 8; i++) { Hs += (i + 1) * (TILE (8 + i, -1) - TILE (6 - i, -1)); Vs += (i + 1) * (TILE (-1, 8 + i) - TILE (-1, 6 - i)); }
        const int a = (TILE (-1, 15) + TILE (15, -1)) << 4, b = (5 * Hs + 32) >> 6, cc = (5 * Vs + 32) >> 6;
        for (int i = 0; i < 4; i++) v[i] = clip255 ((a + b * (x4 + i - 7) + cc * (y - 7) + 16) >> 5);
      } else {
        int st = 0, sl = 0, dc = 128;
#pragma unroll
        for (int i = 0; i < 16; i++) { st += TILE (i, -1); sl += TILE (-1, i); }
        if (mode == LH264_I16_DC) dc = (st + sl + 16) >> 5;
        else if (mode == LH264_I16_DC_T) dc = (st + 8) >> 4;
        else if (mode == LH264_I16_DC_L) dc = (sl + 8) >> 4;
        for (int i = 0; i < 4; i++) v[i] = dc;
      }
      for (int i = 0; i < 4; i++) s.pred[y * 16 + x4 + i] = (uint8_t)v[i];
    }
    __syncthreads();
    if (tid < 96) store_quad (s, pic, mbx, mby, tid, true);
    return;
  }

  if (mb_type == LH264_MB_I4x4) {
    if (tid < 16) {                          // 16 dependent 4x4 steps, one lane per sample
      const int x = tid & 3, y = tid >> 2;
      for (int z = 0; z < 16; z++) {
        const int X0 = blk_x (z), Y0 = blk_y (z);
        const int v = pred4x4_sample (tile, X0, Y0, s.mb.intra_mode[z], x, y);
        __syncwarp (0xffff);
        TILE (X0 + x, Y0 + y) = (uint8_t)clip255 (v + s.res[(Y0 + y) * 16 + X0 + x]);
        __syncwarp (0xffff);
      }
    }
  } else {                                    // I8x8: 4 dependent 8x8 steps, one thread per sample
    const int f = s.mb.nxn_avail;
    for (int q = 0; q < 4; q++) {
      const int X0 = (q & 1) * 8, Y0 = (q >> 1) * 8;
      const int tl = q == 0 ? (f >> 1) & 1 : (q == 1 ? f & 1 : (q == 2 ? (f >> 2) & 1 : 1));
      const int tr = q == 0 ? f & 1 : (q == 1 ? (f >> 3) & 1 : (q == 2 ? 1 : 0));
      int v = 0;
      if (tid < 64) v = pred8x8_sample (tile, X0, Y0, s.mb.intra_mode[4 * q], tl, tr, tid & 7, tid >> 3);
      __syncthreads();
      if (tid < 64) TILE (X0 + (tid & 7), Y0 + (tid >> 3)) = (uint8_t)clip255 (v + s.res[(Y0 + (tid >> 3)) * 16 + X0 + (tid & 7)]);
      __syncthreads();
    }
  }
  __syncthreads();
  if (tid < 64) {                            // luma: the tile already holds pred + residual
    const int y = tid >> 2, x4 = (tid & 3) << 2;
    uchar4 o = make_uchar4 (TILE (x4, y), TILE (x4 + 1, y), TILE (x4 + 2, y), TILE (x4 + 3, y));
    *reinterpret_cast<uchar4*> (Y + (size_t) (mby * 16 + y) * W + mbx * 16 + x4) = o;
  } else if (tid < 96) {
    const int cc = s.mb.cbp >> 4;
    store_quad (s, pic, mbx, mby, tid, cc == 1 || cc == 2);
  }
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

__global__ void __launch_bounds__ (128) k_deblock (const PicDev* __restrict__ pics, int wave) {
  const PicDev pic = pics[blockIdx.y];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int mbx, mby;
  if (!wave_mb (wave, blockIdx.x * 4 + warp, pic.mb_w, pic.mb_h, mbx, mby)) return;
  const int mb_w = pic.mb_w, mbi = mby * mb_w + mbx;
  const lh264_mb_rec* m = pic.mbs + mbi;
  const int mb_type = m->mb_type;
  if (mb_type == LH264_MB_NONE) return;
  const lh264_slice_rec* sl = pic.slices + m->slice_idx;
  const int idc = sl->disable_deblocking_idc;
  if (idc == 1) return;
  const int aoff = sl->alpha_c0_offset, boff = sl->beta_offset;
  // DeblockingAvailableNoInterlayer, deblocking.cpp:354-369
  bool left = mbx > 0, top = mby > 0;
  if (idc == 2) { left = left && m[-1].slice_idx == m->slice_idx; top = top && m[-mb_w].slice_idx == m->slice_idx; }
  const bool t8 = m->flags & LH264_MBF_TRANSFORM8x8;
  const bool intra = LH264_MB_IS_INTRA (mb_type);

  // ---- boundary strength: lane = dir*16 + edge*4 + segment ----
  const int dir = lane >> 4, e = (lane >> 2) & 3, i = lane & 3;
  int bs = 0;
  if (intra) bs = e == 0 ? 4 : 3;
  else if (e == 0) {
    const bool avail = dir ? top : left;
    if (avail) {
      const lh264_mb_rec* nb = dir ? m - mb_w : m - 1;
      if (LH264_MB_IS_INTRA (nb->mb_type)) bs = 4;
      else {                                  // DeblockingBsMarginalMBAvcbase
        const uint32_t nzc = m->nz_mask, nzn = nb->nz_mask;
        const bool n8 = nb->flags & LH264_MBF_TRANSFORM8x8;
        const int rc = dir ? i : 4 * i, rn = dir ? 12 + i : 4 * i + 3;
        const int rc8 = dir ? (i & 2) : 4 * (i & 2), rn8 = dir ? 8 + (i & 2) : 4 * (i & 2) + 2;
        const int nz = (t8 ? nz8 (nzc, rc) : nzb (nzc, rc)) | (n8 ? nz8 (nzn, rn) : nzb (nzn, rn));
        bs = nz ? 2 : mv_differs (m, t8 ? rc8 : rc, nb, n8 ? rn8 : rn);
      }
    }
  } else if (mb_type != LH264_MB_PSKIP && ! (t8 && e != 2)) {
    const uint32_t nzc = m->nz_mask;
    const int a = dir ? 4 * e + i : 4 * i + e, b = dir ? a - 4 : a - 1;
    const int nz = t8 ? (nz8 (nzc, a) | nz8 (nzc, b)) : (nzb (nzc, a) | nzb (nzc, b));
    if (nz) bs = 2;
    else if (mb_type != LH264_MB_P16x16) {
      const int ma = t8 ? ((a & 8) + (a & 2)) : a, mb2 = t8 ? ((b & 8) + (b & 2)) : b;
      bs = mv_differs (m, ma, m, mb2);
    }
  }

  const int W = mb_w * 16, CW = W >> 1;
  uint8_t* Y = plane_ptr (pic, pic.cur_slot, 0) + (size_t)mby * 16 * W + mbx * 16;
  const int qp_cur = m->qp_y;
  const int cplane = (lane - 16) >> 3;         // chroma lanes: 16..23 Cb, 24..31 Cr
  uint8_t* C = nullptr;
  int qpc_cur = 0;
  if (lane >= 16) { C = plane_ptr (pic, pic.cur_slot, 1 + cplane) + (size_t)mby * 8 * CW + mbx * 8; qpc_cur = m->qp_c[cplane]; }

#pragma unroll 1
  for (int d = 0; d < 2; d++) {
    const bool avail = d ? top : left;
    const lh264_mb_rec* nb = d ? m - mb_w : m - 1;
#pragma unroll 1
    for (int ed = 0; ed < 4; ed++) {
      // bS of this lane's line: luma line l -> segment l>>2 ; chroma line l -> segment l>>1
      const int seg = lane < 16 ? (lane >> 2) : ((lane & 7) >> 1);
      const int my_bs = __shfl_sync (0xffffffffu, bs, d * 16 + ed * 4 + seg);
      const unsigned any = __ballot_sync (0xffffffffu, my_bs != 0);
      if (ed == 0 && !avail) continue;
      if ((ed & 1) && t8) continue;
      if (!any) continue;
      if (lane < 16) {
        int qp = qp_cur;
        if (ed == 0) qp = (qp + nb->qp_y + 1) >> 1;
        const int ia = clip3i (qp + aoff, 0, 51), alpha = c_alpha[ia], beta = c_beta[clip3i (qp + boff, 0, 51)];
        if ((alpha | beta) && my_bs) {
          uint8_t* p = d ? Y + (size_t)4 * ed * W + lane : Y + (size_t)lane * W + 4 * ed;
          luma_line (p, d ? W : 1, my_bs, ia, alpha, beta);
        }
      } else if (! (ed & 1)) {
        int qp = qpc_cur;
        if (ed == 0) qp = (qp + nb->qp_c[cplane] + 1) >> 1;
        const int ia = clip3i (qp + aoff, 0, 51), alpha = c_alpha[ia], beta = c_beta[clip3i (qp + boff, 0, 51)];
        if ((alpha | beta) && my_bs) {
          const int l = lane & 7;
          uint8_t* p = d ? C + (size_t)2 * ed * CW + l : C + (size_t)l * CW + 2 * ed;
          chroma_line (p, d ? CW : 1, my_bs, ia, alpha, beta);
        }
      }
      __syncwarp();
    }
    __syncwarp();
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
