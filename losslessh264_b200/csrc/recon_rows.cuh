// recon_rows.cuh — warp-level building blocks of the row kernel (decode_rows.cuh): the per-warp shared-memory context, the
// acquire / release progress accessors, level staging and the GENERAL inter-prediction routine (any partition shape, explicit
// weighting, windows that cross the picture border).  The fused reconstruction + deblocking kernel and its fast paths are in
// decode_rows.cuh.
#pragma once
#include "recon_kernels.cuh"

namespace lh264 {

#ifdef LH264_CPU_EMU
__device__ __forceinline__ int ld_acquire (const int* p) { return *reinterpret_cast<const volatile int*> (p); }
__device__ __forceinline__ void st_release (int* p, int v) { *reinterpret_cast<volatile int*> (p) = v; }
#else
__device__ __forceinline__ int ld_acquire (const int* p) {
  int v;
  asm volatile ("ld.acquire.gpu.global.s32 %0, [%1];" : "=r" (v) : "l" (p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release (int* p, int v) {
  asm volatile ("st.release.gpu.global.s32 [%0], %1;" :: "l" (p), "r" (v) : "memory");
}
#endif

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

#define ROWS_WARPS 8

} // namespace lh264
