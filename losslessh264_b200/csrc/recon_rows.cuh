// recon_rows.cuh — persistent macroblock-row kernels (the product path of lh264_recon_submit).
//
// Work decomposition.  A batch holds one picture per independent stream.  A work item is one
// macroblock ROW of one picture; items are handed out by an atomic ticket in (row, picture) order
// and each item is owned by ONE WARP which walks its row left to right:
//
//   k_recon_rows    inter MBs (GetInterPred/BaseMC/WeightPrediction rec_mb.cpp:247-545 +
//                   WelsMbInterSampleConstruction decode_slice.cpp:208-250) need no neighbour and are
//                   processed as the warp reaches them; intra MBs (WelsMbIntraPredictionConstruction
//                   decode_slice.cpp:313-334, rec_mb.cpp:64-219) first wait until the row above has
//                   finished MB x+1 (top-right neighbour) — the classic 2-MB-lag row wavefront.
//   k_deblock_rows  WelsDeblockingFilterSlice / WelsDeblockingMb (deblocking.cpp:815-933) in the same
//                   row order; MB (x,y) waits for (x+1,y-1), which preserves the standard's edge
//                   order (raster MB order, vertical edges before horizontal ones inside an MB).
//
// Progress is published per row with st.release.gpu / polled with ld.acquire.gpu.  A warp only ever
// waits on a row with a SMALLER ticket, and a ticket is taken by a warp that is already running, so
// the lowest outstanding ticket never waits: no deadlock regardless of how many warps are resident.
// One launch per phase per batch replaces the per-anti-diagonal launches of the first version.
#pragma once
#include "recon_kernels.cuh"

namespace lh264 {

struct RowSync {            // device-resident, zeroed before every batch
  unsigned int ticket[2];   // [0] recon, [1] deblock
  unsigned int pad[2];
  // followed by int progress[2][n_pics][max_rows]
};

__device__ __forceinline__ int ld_acquire (const int* p) {
  int v;
  asm volatile ("ld.acquire.gpu.global.s32 %0, [%1];" : "=r" (v) : "l" (p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release (int* p, int v) {
  asm volatile ("st.release.gpu.global.s32 [%0], %1;" :: "l" (p), "r" (v) : "memory");
}

// k / d for k < 2048, d <= 32 without an integer division: (k * ceil(65536 / d)) >> 16
__constant__ uint16_t c_rcp[33] = { 0, 65535, 32768, 21846, 16384, 13108, 10923, 9363, 8192, 7282, 6554, 5958, 5462, 5042, 4682, 4370, 4096,
  3856, 3641, 3450, 3277, 3121, 2979, 2850, 2731, 2622, 2521, 2428, 2341, 2260, 2185, 2115, 2048 };
__device__ __forceinline__ int fdiv (int k, int d) { return d == 1 ? k : (int) (((unsigned)k * c_rcp[d]) >> 16); }

#define LW_STRIDE 28
#define CW_STRIDE 16
struct __align__ (16) WarpCtx {
  lh264_mb_rec mb;              // 128 B
  int16_t lev[384];             // dense quantised levels, 24 blocks x 16
  int16_t res[384];             // residual: luma 16x16 raster, then Cb 8x8, Cr 8x8
  uint8_t pred[384];            // prediction (same layout)
  uint8_t tile[17 * TS];        // intra luma tile with its top row / left column
  uint8_t cedge[2][2][12];      // chroma neighbours
  __align__ (4) uint8_t lwin[21 * LW_STRIDE + 4]; // luma reference window of one partition (word-aligned rows)
  __align__ (4) uint8_t cwin[2][9 * CW_STRIDE];   // chroma reference windows of one partition (word-aligned rows)
};

// residual_block() works on an MbShared; WarpCtx starts with the same four members.
static_assert (sizeof (MbShared) == 128 + 768 + 768 + 384, "layout");

__device__ __forceinline__ void w_load_mb (WarpCtx& s, const lh264_mb_rec* m, int lane) {
  reinterpret_cast<uint32_t*> (&s.mb)[lane] = __ldg (reinterpret_cast<const uint32_t*> (m) + lane);
  __syncwarp();
}
// the same in two halves, so that the global loads are in flight while the warp does something else
__device__ __forceinline__ void w_issue_levels (const WarpCtx& s, const int16_t* coefs, int lane, uint32_t v[6]) {
  const uint32_t mask = s.mb.coded_mask;
  const uint32_t* src = reinterpret_cast<const uint32_t*> (coefs) + (size_t)s.mb.coef_off * 8;
#pragma unroll
  for (int i = 0; i < 6; i++) {
    const int w = lane + 32 * i, blk = w >> 3;
    v[i] = 0;
    if ((mask >> blk) & 1) v[i] = __ldg (src + __popc (mask & ((1u << blk) - 1)) * 8 + (w & 7));
  }
}
__device__ __forceinline__ void w_store_levels (WarpCtx& s, int lane, const uint32_t v[6]) {
#pragma unroll
  for (int i = 0; i < 6; i++) reinterpret_cast<uint32_t*> (s.lev)[lane + 32 * i] = v[i];
  __syncwarp();
}
__device__ __forceinline__ void w_load_levels (WarpCtx& s, const int16_t* coefs, int lane) {
  const uint32_t mask = s.mb.coded_mask;
  const uint32_t* src = reinterpret_cast<const uint32_t*> (coefs) + (size_t)s.mb.coef_off * 8;
#pragma unroll
  for (int i = 0; i < 6; i++) {
    const int w = lane + 32 * i, blk = w >> 3;
    uint32_t v = 0;
    if ((mask >> blk) & 1) v = __ldg (src + __popc (mask & ((1u << blk) - 1)) * 8 + (w & 7));
    reinterpret_cast<uint32_t*> (s.lev)[w] = v;
  }
  __syncwarp();
}

// final add + clip + store, 96 quads of 4 samples, 3 per lane
__device__ __forceinline__ void w_store (const WarpCtx& s, const PicDev& pic, int mbx, int mby, int lane, bool add_luma, bool add_chroma, bool luma_from_tile) {
  const int W = pic.mb_w * 16;
#pragma unroll
  for (int i = 0; i < 3; i++) {
    const int t = lane + 32 * i;
    int idx;
    uint8_t* dst;
    bool add;
    if (t < 64) {
      const int y = t >> 2, x4 = (t & 3) << 2;
      idx = y * 16 + x4; add = add_luma;
      dst = plane_ptr (pic, pic.cur_slot, 0) + (size_t) (mby * 16 + y) * W + mbx * 16 + x4;
      if (luma_from_tile) {
        *reinterpret_cast<uchar4*> (dst) = make_uchar4 (s.tile[(y + 1) * TS + x4 + 1], s.tile[(y + 1) * TS + x4 + 2], s.tile[(y + 1) * TS + x4 + 3], s.tile[(y + 1) * TS + x4 + 4]);
        continue;
      }
    } else {
      const int c = t - 64, p = c >> 4, y = (c & 15) >> 1, x4 = (c & 1) << 2;
      idx = 256 + 64 * p + y * 8 + x4; add = add_chroma;
      dst = plane_ptr (pic, pic.cur_slot, 1 + p) + (size_t) (mby * 8 + y) * (W >> 1) + mbx * 8 + x4;
    }
    uchar4 o;
    if (add) {
      o.x = (uint8_t)clip255 (s.pred[idx] + s.res[idx]); o.y = (uint8_t)clip255 (s.pred[idx + 1] + s.res[idx + 1]);
      o.z = (uint8_t)clip255 (s.pred[idx + 2] + s.res[idx + 2]); o.w = (uint8_t)clip255 (s.pred[idx + 3] + s.res[idx + 3]);
    } else {
      o.x = s.pred[idx]; o.y = s.pred[idx + 1]; o.z = s.pred[idx + 2]; o.w = s.pred[idx + 3];
    }
    *reinterpret_cast<uchar4*> (dst) = o;
  }
}

// ---- inter prediction of one partition by the whole warp ---------------------------------------------------
// BaseMC (rec_mb.cpp:247-274) + McLuma_c/McChroma_c (mc.cpp:334-378) + WeightPrediction (rec_mb.cpp:276-341)
__device__ void w_inter_partition (WarpCtx& s, const PicDev& pic, const lh264_slice_rec* sl, int mbx, int mby, int lane,
                                   int ox, int oy, int bw, int bh, int r, int q) {
  const int W = pic.mb_w * 16, H = pic.mb_h * 16, CW = W >> 1, CH = H >> 1;
  const int refidx = s.mb.ref_idx[q] & 15;
  const int slot = clip3i (__ldg (&sl->ref_slot[refidx]), 0, pic.n_slots - 1);
  const int fx = clip3i (((mbx * 16 + ox) << 2) + s.mb.mv[r][0], (-32 + 2) * 4, (W + 32 - 19) << 2);
  const int fy = clip3i (((mby * 16 + oy) << 2) + s.mb.mv[r][1], (-32 + 2) * 4, (H + 32 - 19) << 2);
  const int wp = __ldg (&sl->weighted_pred);
  // ---- luma window: (bw + 5) x (bh + 5) samples starting at (ix, iy), rows clamped ----
  const int ix = (fx >> 2) - 2, iy = (fy >> 2) - 2, ww = bw + 5, wh = bh + 5;
  const uint8_t* ref = plane_ptr (pic, slot, 0);
  const int cx = fx >> 3, cy = fy >> 3, cw = (bw >> 1) + 1, chh = (bh >> 1) + 1;
  int woff, coff;                                  // byte offsets of sample (ix, iy) / (cx, cy) inside a window row
  const bool lfast = ix >= 0 && ix + ww <= W, cfast = cx >= 0 && cx + cw <= CW;
  if (lfast && cfast) {
    // fast path: both windows as aligned 32-bit words, every global load issued before the first shared store so that
    // the warp pays one memory latency per partition instead of one per window
    const int ax = ix & ~3, nw = ((ix + ww + 3) >> 2) - (ax >> 2), nl = nw * wh;          // <= 6 x 21 words
    const int acx = cx & ~3, nwc = ((cx + cw + 3) >> 2) - (acx >> 2), per = nwc * chh;    // <= 3 x 9 words per plane
    woff = ix - ax; coff = cx - acx;
    uint32_t lv[4], cv[2];
    int lo_[4], co_[2];
#pragma unroll
    for (int i = 0; i < 4; i++) {
      const int k = lane + 32 * i;
      lo_[i] = -1; lv[i] = 0;
      if (k < nl) {
        const int wy = fdiv (k, nw), wx = k - wy * nw;
        lv[i] = __ldg (reinterpret_cast<const uint32_t*> (ref + (size_t)clip3i (iy + wy, 0, H - 1) * W + ax) + wx);
        lo_[i] = wy * LW_STRIDE + 4 * wx;
      }
    }
#pragma unroll
    for (int i = 0; i < 2; i++) {
      const int k = lane + 32 * i;
      co_[i] = -1; cv[i] = 0;
      if (k < 2 * per) {
        const int p = k >= per, kk = p ? k - per : k;
        const int wy = fdiv (kk, nwc), wx = kk - wy * nwc;
        cv[i] = __ldg (reinterpret_cast<const uint32_t*> (plane_ptr (pic, slot, 1 + p) + (size_t)clip3i (cy + wy, 0, CH - 1) * CW + acx) + wx);
        co_[i] = p * (9 * CW_STRIDE) + wy * CW_STRIDE + 4 * wx;
      }
    }
#pragma unroll
    for (int i = 0; i < 4; i++) if (lo_[i] >= 0) *reinterpret_cast<uint32_t*> (s.lwin + lo_[i]) = lv[i];
#pragma unroll
    for (int i = 0; i < 2; i++) if (co_[i] >= 0) *reinterpret_cast<uint32_t*> (&s.cwin[0][0] + co_[i]) = cv[i];
  } else {                                         // picture border: clamp every sample (= ExpandPicture, expand_pic.cpp:37-115)
    woff = 0; coff = 0;
    for (int k = lane; k < ww * wh; k += 32) {
      const int wy = fdiv (k, ww), wx = k - wy * ww;
      s.lwin[wy * LW_STRIDE + wx] = __ldg (ref + (size_t)clip3i (iy + wy, 0, H - 1) * W + clip3i (ix + wx, 0, W - 1));
    }
    for (int k = lane; k < 2 * cw * chh; k += 32) {
      const int p = k >= cw * chh, kk = p ? k - cw * chh : k;
      const int wy = fdiv (kk, cw), wx = kk - wy * cw;
      s.cwin[p][wy * CW_STRIDE + wx] = __ldg (plane_ptr (pic, slot, 1 + p) + (size_t)clip3i (cy + wy, 0, CH - 1) * CW + clip3i (cx + wx, 0, CW - 1));
    }
  }
  __syncwarp();
  // ---- luma samples: quads of 4 ----
  const int frac = ((fx & 3) << 2) | (fy & 3);
  int lw = 0, lo = 0, ld = 0;
  if (wp) { lw = __ldg (&sl->luma_weight[refidx]); lo = __ldg (&sl->luma_offset[refidx]); ld = __ldg (&sl->luma_log2_denom); }
  const int qpr = bw >> 2;                          // quads per row
  for (int k = lane; k < qpr * bh; k += 32) {
    const int y = fdiv (k, qpr), x4 = (k - y * qpr) << 2;
    const uint8_t* w0 = s.lwin + (2 + y) * LW_STRIDE + woff + 2 + x4;
    int v[4];
    mc_luma_quad (w0, LW_STRIDE, frac, v);
    if (wp) {
#pragma unroll
      for (int i = 0; i < 4; i++) v[i] = wp_sample (v[i], lw, lo, ld);
    }
    *reinterpret_cast<uint32_t*> (s.pred + (oy + y) * 16 + ox + x4) = (uint32_t)v[0] | ((uint32_t)v[1] << 8) | ((uint32_t)v[2] << 16) | ((uint32_t)v[3] << 24);
  }
  // ---- chroma samples: pairs of 2, both planes ----
  const int dx = fx & 7, dy = fy & 7;
  const int A = (8 - dx) * (8 - dy), B = dx * (8 - dy), C = (8 - dx) * dy, D = dx * dy;
  const int ppr = bw >> 2, cbh = bh >> 1;           // pairs per chroma row, chroma rows
  for (int k = lane; k < 2 * ppr * cbh; k += 32) {
    const int p = k >= ppr * cbh, kk = p ? k - ppr * cbh : k;
    const int y = fdiv (kk, ppr), x2 = (kk - y * ppr) << 1;
    int cwt = 0, co = 0, cd = 0;
    if (wp) { cwt = __ldg (&sl->chroma_weight[refidx][p]); co = __ldg (&sl->chroma_offset[refidx][p]); cd = __ldg (&sl->chroma_log2_denom); }
    const uint8_t* c0 = s.cwin[p] + y * CW_STRIDE + coff + x2;
#pragma unroll
    for (int i = 0; i < 2; i++) {
      int o = (A * c0[i] + B * c0[i + 1] + C * c0[CW_STRIDE + i] + D * c0[CW_STRIDE + i + 1] + 32) >> 6;
      // the reference weights only the top-left (bw>>2) x (bh>>2) chroma samples of a partition (rec_mb.cpp:309-311)
      if (wp && x2 + i < (bw >> 2) && y < (bh >> 2)) o = wp_sample (o, cwt, co, cd);
      s.pred[256 + 64 * p + ((oy >> 1) + y) * 8 + (ox >> 1) + x2 + i] = (uint8_t)o;
    }
  }
  __syncwarp();
}

__device__ void w_inter_mb (WarpCtx& s, const PicDev& pic, int mbx, int mby, int lane) {
  const lh264_slice_rec* sl = pic.slices + s.mb.slice_idx;
  const int t = s.mb.mb_type;
  const bool has_res = s.mb.cbp != 0;
  uint32_t lv[6];
  if (has_res) w_issue_levels (s, pic.coefs, lane, lv);          // in flight during the first partition
  bool res_done = !has_res;
  auto residual = [&] () {
    if (res_done) return;
    res_done = true;
    w_store_levels (s, lane, lv);
    if (lane < 24) residual_block (reinterpret_cast<MbShared&> (s), lane);
  };
  if (t == LH264_MB_P16x16 || t == LH264_MB_PSKIP) { w_inter_partition (s, pic, sl, mbx, mby, lane, 0, 0, 16, 16, 0, 0); residual(); }
  else if (t == LH264_MB_P16x8) { w_inter_partition (s, pic, sl, mbx, mby, lane, 0, 0, 16, 8, 0, 0); residual(); w_inter_partition (s, pic, sl, mbx, mby, lane, 0, 8, 16, 8, 8, 2); }
  else if (t == LH264_MB_P8x16) { w_inter_partition (s, pic, sl, mbx, mby, lane, 0, 0, 8, 16, 0, 0); residual(); w_inter_partition (s, pic, sl, mbx, mby, lane, 8, 0, 8, 16, 2, 1); }
  else {
#pragma unroll 1
    for (int q = 0; q < 4; q++) {
      const int qx = (q & 1) << 3, qy = (q >> 1) << 3, r0 = (qy >> 2) * 4 + (qx >> 2), sub = s.mb.sub_mb_type[q];
      if (sub == LH264_SUB_8x8) w_inter_partition (s, pic, sl, mbx, mby, lane, qx, qy, 8, 8, r0, q);
      else if (sub == LH264_SUB_8x4) { w_inter_partition (s, pic, sl, mbx, mby, lane, qx, qy, 8, 4, r0, q); w_inter_partition (s, pic, sl, mbx, mby, lane, qx, qy + 4, 8, 4, r0 + 4, q); }
      else if (sub == LH264_SUB_4x8) { w_inter_partition (s, pic, sl, mbx, mby, lane, qx, qy, 4, 8, r0, q); w_inter_partition (s, pic, sl, mbx, mby, lane, qx + 4, qy, 4, 8, r0 + 1, q); }
      else {
        w_inter_partition (s, pic, sl, mbx, mby, lane, qx, qy, 4, 4, r0, q); w_inter_partition (s, pic, sl, mbx, mby, lane, qx + 4, qy, 4, 4, r0 + 1, q);
        w_inter_partition (s, pic, sl, mbx, mby, lane, qx, qy + 4, 4, 4, r0 + 4, q); w_inter_partition (s, pic, sl, mbx, mby, lane, qx + 4, qy + 4, 4, 4, r0 + 5, q);
      }
      residual();
    }
  }
  __syncwarp();
  w_store (s, pic, mbx, mby, lane, has_res, has_res, false);
}

// ---- intra MB by one warp ----------------------------------------------------------------------------------
#define F3(a, b, c) (((a) + 2 * (b) + (c) + 2) >> 2)
#define F2(a, b) (((a) + (b) + 1) >> 1)
__device__ void w_intra_mb (WarpCtx& s, const PicDev& pic, int mbx, int mby, int lane) {
  uint8_t* tile = s.tile;
  const int mb_type = s.mb.mb_type;
  const int W = pic.mb_w * 16, CW = W >> 1;
  const uint8_t* Y = plane_ptr (pic, pic.cur_slot, 0);
  w_load_levels (s, pic.coefs, lane);
  if (mb_type == LH264_MB_IPCM) {            // samples travel in the coefficient pool
    for (int i = lane; i < 384; i += 32) s.pred[i] = (uint8_t)s.lev[i];
    __syncwarp();
    w_store (s, pic, mbx, mby, lane, false, false, false);
    return;
  }
  // neighbour samples -> shared (coordinates clamped: a resolved mode never uses an unavailable neighbour)
  for (int t = lane; t < 77; t += 32) {
    if (t < 25) {
      const int x = mbx * 16 - 1 + t, y = mby * 16 - 1;
      tile[t] = Y[(size_t)max (y, 0) * W + clip3i (x, 0, W - 1)];
    } else if (t < 41) {
      const int y = mby * 16 + t - 25, x = mbx * 16 - 1;
      tile[(t - 25 + 1) * TS] = Y[(size_t)y * W + max (x, 0)];
    } else {
      const int c = t - 41, p = c / 18, k = c % 18;        // k 0..8 top (x=-1..7), 9..17 left (y=-1..7)
      const uint8_t* C = plane_ptr (pic, pic.cur_slot, 1 + p);
      int x, y;
      if (k < 9) { x = mbx * 8 - 1 + k; y = mby * 8 - 1; } else { x = mbx * 8 - 1; y = mby * 8 - 1 + (k - 9); }
      s.cedge[p][k < 9 ? 0 : 1][k < 9 ? k : k - 9] = C[(size_t)max (y, 0) * CW + clip3i (x, 0, CW - 1)];
    }
  }
  if (lane < 24) residual_block (reinterpret_cast<MbShared&> (s), lane);
  __syncwarp();

  // chroma prediction, WelsIChromaPred*_c: 32 quads, one per lane
  {
    const int c = lane, p = c >> 4, y = (c & 15) >> 1, x4 = (c & 1) << 2;
    const uint8_t* top = &s.cedge[p][0][1];   // top[-1] = top-left
    const uint8_t* left = &s.cedge[p][1][1];
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
    const int mode = s.mb.i16_mode;
    int b = 0, cc = 0, a = 0, dc = 128;
    if (mode == LH264_I16_P) {
      int Hs = 0, Vs = 0;
#pragma unroll
      for (int i = 0; i < 8; i++) { Hs += (i + 1) * (TILE (8 + i, -1) - TILE (6 - i, -1)); Vs += (i + 1) * (TILE (-1, 8 + i) - TILE (-1, 6 - i)); }
      a = (TILE (-1, 15) + TILE (15, -1)) << 4; b = (5 * Hs + 32) >> 6; cc = (5 * Vs + 32) >> 6;
    } else if (mode >= LH264_I16_DC) {
      int st = 0, sl = 0;
#pragma unroll
      for (int i = 0; i < 16; i++) { st += TILE (i, -1); sl += TILE (-1, i); }
      if (mode == LH264_I16_DC) dc = (st + sl + 16) >> 5;
      else if (mode == LH264_I16_DC_T) dc = (st + 8) >> 4;
      else if (mode == LH264_I16_DC_L) dc = (sl + 8) >> 4;
    }
#pragma unroll
    for (int h = 0; h < 2; h++) {
      const int t = lane + 32 * h, y = t >> 2, x4 = (t & 3) << 2;
      for (int i = 0; i < 4; i++) {
        int v;
        if (mode == LH264_I16_V) v = TILE (x4 + i, -1);
        else if (mode == LH264_I16_H) v = TILE (-1, y);
        else if (mode == LH264_I16_P) v = clip255 ((a + b * (x4 + i - 7) + cc * (y - 7) + 16) >> 5);
        else v = dc;
        s.pred[y * 16 + x4 + i] = (uint8_t)v;
      }
    }
    __syncwarp();
    w_store (s, pic, mbx, mby, lane, true, true, false);
    return;
  }

  if (mb_type == LH264_MB_I4x4) {              // 16 dependent 4x4 steps, one lane per sample
    const int x = lane & 3, y = (lane >> 2) & 3;
#pragma unroll 1
    for (int z = 0; z < 16; z++) {
      const int X0 = blk_x (z), Y0 = blk_y (z);
      int v = 0;
      if (lane < 16) v = pred4x4_sample (tile, X0, Y0, s.mb.intra_mode[z], x, y);
      __syncwarp();
      if (lane < 16) TILE (X0 + x, Y0 + y) = (uint8_t)clip255 (v + s.res[(Y0 + y) * 16 + X0 + x]);
      __syncwarp();
    }
  } else {                                      // I8x8: 4 dependent 8x8 steps, two samples per lane
    const int f = s.mb.nxn_avail;
#pragma unroll 1
    for (int q = 0; q < 4; q++) {
      const int X0 = (q & 1) * 8, Y0 = (q >> 1) * 8;
      const int tl = q == 0 ? (f >> 1) & 1 : (q == 1 ? f & 1 : (q == 2 ? (f >> 2) & 1 : 1));
      const int tr = q == 0 ? f & 1 : (q == 1 ? (f >> 3) & 1 : (q == 2 ? 1 : 0));
      const int mode = s.mb.intra_mode[4 * q];
      const int v0 = pred8x8_sample (tile, X0, Y0, mode, tl, tr, lane & 7, lane >> 3);
      const int v1 = pred8x8_sample (tile, X0, Y0, mode, tl, tr, lane & 7, 4 + (lane >> 3));
      __syncwarp();
      TILE (X0 + (lane & 7), Y0 + (lane >> 3)) = (uint8_t)clip255 (v0 + s.res[(Y0 + (lane >> 3)) * 16 + X0 + (lane & 7)]);
      TILE (X0 + (lane & 7), Y0 + 4 + (lane >> 3)) = (uint8_t)clip255 (v1 + s.res[(Y0 + 4 + (lane >> 3)) * 16 + X0 + (lane & 7)]);
      __syncwarp();
    }
  }
  const int cc = s.mb.cbp >> 4;
  w_store (s, pic, mbx, mby, lane, false, cc == 1 || cc == 2, true);
}
#undef F2
#undef F3

#define ROWS_WARPS 8
__global__ void __launch_bounds__ (ROWS_WARPS * 32, 2) k_recon_rows (const PicDev* __restrict__ pics, int n_pics, int max_rows, RowSync* sync) {
  __shared__ WarpCtx ctx[ROWS_WARPS];
  const int lane = threadIdx.x & 31;
  WarpCtx& s = ctx[threadIdx.x >> 5];
  int* progress = reinterpret_cast<int*> (sync + 1);
  const unsigned total = (unsigned)n_pics * max_rows;
  for (;;) {
    unsigned t = 0;
    if (lane == 0) t = atomicAdd (&sync->ticket[0], 1u);
    t = __shfl_sync (0xffffffffu, t, 0);
    if (t >= total) return;
    const int mby = t / n_pics, pi = t - mby * n_pics;
    const PicDev pic = pics[pi];
    if (mby >= pic.mb_h) continue;
    int* my = progress + (size_t)pi * max_rows + mby;
    const int* above = my - 1;
    int seen = 0;                                   // progress of the row above known so far
    const lh264_mb_rec* row = pic.mbs + mby * pic.mb_w;
    uint32_t recw = __ldg (reinterpret_cast<const uint32_t*> (row) + lane);      // the record of the next MB is always in flight
#pragma unroll 1
    for (int mbx = 0; mbx < pic.mb_w; mbx++) {
      __syncwarp();
      reinterpret_cast<uint32_t*> (&s.mb)[lane] = recw;
      __syncwarp();
      if (mbx + 1 < pic.mb_w) recw = __ldg (reinterpret_cast<const uint32_t*> (row + mbx + 1) + lane);
      const int mb_type = s.mb.mb_type;
      if (LH264_MB_IS_INTER (mb_type)) w_inter_mb (s, pic, mbx, mby, lane);
      else if (LH264_MB_IS_INTRA (mb_type)) {
        if (mby > 0) {
          const int need = min (mbx + 2, pic.mb_w);
          if (seen < need) {
            if (lane == 0) { while ((seen = ld_acquire (above)) < need) __nanosleep (64); }
            seen = __shfl_sync (0xffffffffu, seen, 0);
          }
        }
        w_intra_mb (s, pic, mbx, mby, lane);
      }
      __syncwarp();
      // publish: every MB for the last rows' consumers would be wasteful; intra consumers below need x+2
      if (lane == 0 && ((mbx & 3) == 3 || mbx == pic.mb_w - 1)) { __threadfence(); st_release (my, mbx + 1); }
    }
  }
}

// ---- deblocking, one warp per MB row ---------------------------------------------------------------------------
// m / ml / mt: the records of this MB, its left and its top neighbour, staged in shared memory by the caller (the record
// of the next MB and of its top neighbour are loaded while this MB is filtered)
__device__ void w_deblock_mb (const PicDev& pic, int mbx, int mby, int lane, const lh264_mb_rec* m, const lh264_mb_rec* ml, const lh264_mb_rec* mt) {
  const int mb_w = pic.mb_w;
  const int mb_type = m->mb_type;
  if (mb_type == LH264_MB_NONE) return;
  const lh264_slice_rec* sl = pic.slices + m->slice_idx;
  const int idc = sl->disable_deblocking_idc;
  if (idc == 1) return;
  const int aoff = sl->alpha_c0_offset, boff = sl->beta_offset;
  bool left = mbx > 0, top = mby > 0;               // DeblockingAvailableNoInterlayer, deblocking.cpp:354-369
  if (idc == 2) { left = left && ml->slice_idx == m->slice_idx; top = top && mt->slice_idx == m->slice_idx; }
  const bool t8 = m->flags & LH264_MBF_TRANSFORM8x8;
  const bool intra = LH264_MB_IS_INTRA (mb_type);
  // boundary strength: lane = dir*16 + edge*4 + segment
  const int dir = lane >> 4, e = (lane >> 2) & 3, i = lane & 3;
  int bs = 0;
  if (intra) bs = e == 0 ? 4 : 3;
  else if (e == 0) {
    const bool avail = dir ? top : left;
    if (avail) {
      const lh264_mb_rec* nb = dir ? mt : ml;
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
  if (!left && dir == 0 && e == 0) bs = 0;
  if (!top && dir == 1 && e == 0) bs = 0;
  if (__ballot_sync (0xffffffffu, bs != 0) == 0) return;     // nothing to filter in this MB

  const int W = mb_w * 16, CW = W >> 1;
  uint8_t* Y = plane_ptr (pic, pic.cur_slot, 0) + (size_t)mby * 16 * W + mbx * 16;
  const int qp_cur = m->qp_y;
  const int cplane = (lane - 16) >> 3;         // chroma lanes: 16..23 Cb, 24..31 Cr
  uint8_t* C = nullptr;
  int qpc_cur = 0;
  if (lane >= 16) { C = plane_ptr (pic, pic.cur_slot, 1 + cplane) + (size_t)mby * 8 * CW + mbx * 8; qpc_cur = m->qp_c[cplane]; }
#pragma unroll 1
  for (int d = 0; d < 2; d++) {
    const lh264_mb_rec* nb = d ? mt : ml;
#pragma unroll 1
    for (int ed = 0; ed < 4; ed++) {
      const int seg = lane < 16 ? (lane >> 2) : ((lane & 7) >> 1);
      const int my_bs = __shfl_sync (0xffffffffu, bs, d * 16 + ed * 4 + seg);
      const unsigned any = __ballot_sync (0xffffffffu, my_bs != 0);
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
  }
}

__global__ void __launch_bounds__ (ROWS_WARPS * 32, 4) k_deblock_rows (const PicDev* __restrict__ pics, int n_pics, int max_rows, RowSync* sync) {
  __shared__ __align__ (16) lh264_mb_rec recs[ROWS_WARPS][4];       // [0],[1]: this MB / its left neighbour (alternating), [2]: top
  const int lane = threadIdx.x & 31;
  lh264_mb_rec* my_recs = recs[threadIdx.x >> 5];
  int* progress = reinterpret_cast<int*> (sync + 1) + (size_t)n_pics * max_rows;
  const unsigned total = (unsigned)n_pics * max_rows;
  for (;;) {
    unsigned t = 0;
    if (lane == 0) t = atomicAdd (&sync->ticket[1], 1u);
    t = __shfl_sync (0xffffffffu, t, 0);
    if (t >= total) return;
    const int mby = t / n_pics, pi = t - mby * n_pics;
    const PicDev pic = pics[pi];
    if (mby >= pic.mb_h) continue;
    int* my = progress + (size_t)pi * max_rows + mby;
    const int* above = my - 1;
    int seen = 0;
    const uint32_t* row = reinterpret_cast<const uint32_t*> (pic.mbs + (size_t)mby * pic.mb_w);
    const uint32_t* row_top = mby > 0 ? reinterpret_cast<const uint32_t*> (pic.mbs + (size_t) (mby - 1) * pic.mb_w) : row;
    uint32_t w_cur = __ldg (row + lane), w_top = __ldg (row_top + lane);
#pragma unroll 1
    for (int mbx = 0; mbx < pic.mb_w; mbx++) {
      lh264_mb_rec* mc = &my_recs[mbx & 1];
      const lh264_mb_rec* ml = &my_recs[(mbx & 1) ^ 1];
      __syncwarp();
      reinterpret_cast<uint32_t*> (mc)[lane] = w_cur;
      reinterpret_cast<uint32_t*> (&my_recs[2])[lane] = w_top;
      __syncwarp();
      if (mbx + 1 < pic.mb_w) { w_cur = __ldg (row + (mbx + 1) * 32 + lane); w_top = __ldg (row_top + (mbx + 1) * 32 + lane); }
      if (mby > 0) {
        const int need = min (mbx + 2, pic.mb_w);
        if (seen < need) {
          if (lane == 0) { while ((seen = ld_acquire (above)) < need) __nanosleep (64); }
          seen = __shfl_sync (0xffffffffu, seen, 0);
        }
      }
      w_deblock_mb (pic, mbx, mby, lane, mc, ml, &my_recs[2]);
      __syncwarp();
      if (lane == 0 && ((mbx & 3) == 3 || mbx == pic.mb_w - 1)) { __threadfence(); st_release (my, mbx + 1); }
    }
  }
}

} // namespace lh264
