// decode_rows.cuh — k_decode_rows: macroblock reconstruction AND in-loop deblocking in ONE persistent row kernel
// (the product path of lh264_recon_submit / lh264_batch_recon).
//
// Replaces, per picture:  WelsTargetSliceConstruction (decode_slice.cpp:110-206) = per MB WelsTargetMbConstruction (:353-373:
// GetInterPred / BaseMC / WeightPrediction rec_mb.cpp:247-545, WelsMbInterSampleConstruction :208-250,
// WelsMbIntraPredictionConstruction :313-334) followed by WelsDeblockingFilterSlice / WelsDeblockingMb (deblocking.cpp:815-933).
//
// Work decomposition.  A batch holds one picture per independent stream.  A work item is one macroblock ROW of one picture,
// handed out by an atomic ticket in (row, picture) order and owned by ONE WARP that walks it left to right.  For every
// macroblock x of its row y the warp
//   1. waits until the row above has finished macroblock x + 1 (progress counter, acquire / release);
//   2. reconstructs the UNFILTERED macroblock U(x, y) into a shared-memory tile ("cur"): prediction + residual.  Intra
//      prediction reads unfiltered neighbours only: the left column of U(x - 1, y) kept in shared memory, and the bottom rows
//      of U(x - 1 .. x + 1, y - 1) from a per-stream line buffer in HBM that every row writes for the row below;
//   3. deblocks macroblock x in the standard's order (vertical edges left to right, then horizontal edges top to bottom):
//      the left edge works across the boundary of the "prev" tile (macroblock x - 1, already filtered by its own edges, still
//      in shared memory at columns 0..15) and "cur" (columns 16..31); the top edge works on rows 12..15 of the macroblock above,
//      read back from the picture (they are final once the row above has finished macroblock x + 1), which sit right above
//      "cur" in the same shared array, and writes the three modified rows back;
//   4. stores macroblock x - 1 (now final but for the edge the row below will filter) with 8-byte stores, then shifts cur -> prev.
// The picture is written once and the reference picture read once: the separate deblocking pass (one more read + write of
// every picture) of the first version is gone, and so is the second persistent launch.
//
// Fast path (the bulk of real streams: P_L0_16x16 / P_Skip, no explicit weighting, window inside the picture): reference
// samples go from global memory straight to registers as aligned 32-bit words + funnel shifts (integer-pel luma is a copy,
// chroma is a 4-tap bilinear done with dp4a on byte pairs); residual blocks are dequantised and transformed only where
// coded_mask says a block exists, their levels read as two 128-bit loads per block.  Everything else (partitions, explicit
// weighting, windows crossing the picture border, 8x8 transform, intra) takes the general routines of recon_rows.cuh.
#pragma once
#include "recon_rows.cuh"

#ifdef LH264_CPU_EMU
#define LH264_NOINLINE __attribute__ ((noinline))
#else
#define LH264_NOINLINE                 // (measured: __noinline__ here puts the PicDev copy and 264 bytes of spills into local memory, +65 % kernel time)
#endif

namespace lh264 {

#define RY 36                 // row stride of the luma tile array: prev MB at columns 0..15, cur MB at 16..31, 4 bytes of padding
#define RC 20                 // row stride of a chroma tile array: prev at 0..7, cur at 8..15, padding
#define RY_TOP 4              // rows 0..3 of the luma array = rows 12..15 of the macroblock above (columns 16..31)
#define RC_TOP 2

struct __align__ (16) RowCtx : WarpCtx {
  lh264_mb_rec left, top;                       // records of the left / top neighbour (boundary strength)
  __align__ (4) uint8_t ry[(RY_TOP + 16) * RY];
  __align__ (4) uint8_t rc[2][(RC_TOP + 8) * RC];
  __align__ (4) uint8_t ucol[32];               // unfiltered right column of the previous MB: luma 16, Cb 8, Cr 8
  // what the hot path reads of the current slice's record, refreshed when slice_idx changes (most pictures: once per row)
  int16_t sl_slot[LH264_MAX_REFS];              // picture slot of RefPicList0[i], clamped to the stream's pool
  uint8_t sl_wp, sl_idc; int8_t sl_aoff, sl_boff;
  uint8_t sl_scal;                              // LH264_SLF_*: scaling lists in use (cold path: residual_block_scaled)
  // Per-lane addresses of the row being decoded, for macroblock 0 (a lane adds mbx * 16 or mbx * 8): kept here because the
  // compiler, short of registers, recomputed them from the picture descriptor at every use (a tenth of the kernel's instructions
  // were loop invariants).  [0] luma store (row lane / 2, half lane & 1), [1] chroma store, [2] the words of the rows above
  // (lanes 0..15 luma rows -4..-1, lanes 16..23 chroma rows -2..-1), [3] line-buffer words (lanes 0..3 luma, 4..7 chroma)
  unsigned long long lp[4][32];
  // the picture descriptor of the row being decoded: 18 registers as a by-value copy; here its fields cost one LDS where they are
  // used and the kernel fits its 128 registers without spills or recomputation (P pictures 14.3 -> 13.2 ms per 1024)
  PicDev pic;
};

struct DecodeSync {          // device-resident, zeroed before every batch
  unsigned int ticket;
  unsigned int pad[3];
  // followed by int progress[n_pics][max_rows]: macroblocks of that row that are stored (final but for the bottom edge)
};

__device__ __forceinline__ uint32_t dp4a_uu (uint32_t a, uint32_t b, uint32_t c) {
#ifdef LH264_CPU_EMU
  for (int i = 0; i < 4; i++) c += ((a >> (8 * i)) & 0xffu) * ((b >> (8 * i)) & 0xffu);
  return c;
#else
  return __dp4a (a, b, c);
#endif
}

// pointer to the line buffer of MB row y: unfiltered bottom row of every MB of that row (luma W, Cb W/2, Cr W/2)
__device__ __forceinline__ uint8_t* line_ptr (const PicDev& pic, int y, int plane) {
  const int W = pic.mb_w * 16;
  return pic.lines + (size_t)y * 2 * W + (plane == 0 ? 0 : (plane == 1 ? W : W + (W >> 1)));
}

// four samples + four residuals, clipped
__device__ __forceinline__ uint32_t add_res4 (uint32_t pix, const int16_t* r) {
  const uint2 v = *reinterpret_cast<const uint2*> (r);
  const int r0 = (int16_t) (v.x & 0xffffu), r1 = (int16_t) (v.x >> 16), r2 = (int16_t) (v.y & 0xffffu), r3 = (int16_t) (v.y >> 16);
  const uint32_t b0 = (uint32_t)clip255 ((int) (pix & 0xffu) + r0), b1 = (uint32_t)clip255 ((int) ((pix >> 8) & 0xffu) + r1);
  const uint32_t b2 = (uint32_t)clip255 ((int) ((pix >> 16) & 0xffu) + r2), b3 = (uint32_t)clip255 ((int) (pix >> 24) + r3);
  return b0 | (b1 << 8) | (b2 << 16) | (b3 << 24);
}

// ---- residual of an inter macroblock without the 8x8 transform, levels read straight from the coefficient pool ---------------
// Same arithmetic as residual_block(): dequantisation as fused into the reference parse (parse_mb_syn_cavlc.cpp:938-1005),
// WelsChromaDcIdct (decode_slice.cpp:375-396), IdctResAddPred_c (decode_mb_aux.cpp:42-77).  Lane k < 24 owns coefficient
// block k; only blocks that exist are transformed (luma: coded_mask bit; chroma: any block of the plane present).
__device__ __forceinline__ void w_residual_direct (RowCtx& s, const int16_t* coefs, int lane) {
  const uint32_t mask = s.mb.coded_mask;
  const bool present = lane < 24 && ((mask >> lane) & 1u);
  uint4 v0 = make_uint4 (0u, 0u, 0u, 0u), v1 = v0;
  if (present) {
    const uint4* src = reinterpret_cast<const uint4*> (coefs + ((size_t)s.mb.coef_off + __popc (mask & ((1u << lane) - 1u))) * 16);
    v0 = __ldg (src); v1 = __ldg (src + 1);
  }
  // the four DC levels of a chroma plane travel by shuffle (absent blocks contribute 0)
  const int dc_own = (int16_t) (v0.x & 0xffffu);
  const int cb = lane >= 16 && lane < 24 ? (lane & ~3) : 16;
  const int dc0 = __shfl_sync (0xffffffffu, dc_own, cb), dc1 = __shfl_sync (0xffffffffu, dc_own, cb + 1);
  const int dc2 = __shfl_sync (0xffffffffu, dc_own, cb + 2), dc3 = __shfl_sync (0xffffffffu, dc_own, cb + 3);
  if (lane >= 24) return;
  const uint32_t w[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
  // A block whose only non-zero coefficient is the DC one transforms to a constant: with c[1..15] = 0 both passes of
  // IdctResAddPred_c reduce to (32 + c[0]) >> 6 for all sixteen residuals (most blocks of a typical inter picture).
  const bool ac_zero = ((w[0] & 0xffff0000u) | w[1] | w[2] | w[3] | w[4] | w[5] | w[6] | w[7]) == 0;
  int c[16];
  int16_t* dst; int stride;
  if (lane < 16) {
    if (!present) return;
    const int qp = s.mb.qp_y;
    dst = s.res + blk_y (lane) * 16 + blk_x (lane); stride = 16;
    c[0] = S16 (dc_own * dq4 (qp, 0));
    if (!ac_zero) {
#pragma unroll
      for (int j = 1; j < 16; j++) c[j] = S16 ((int) (int16_t) (w[j >> 1] >> ((j & 1) * 16)) * dq4 (qp, j));
    }
  } else {
    const int p = (lane - 16) >> 2, b = (lane - 16) & 3;
    if (((mask >> (16 + 4 * p)) & 15u) == 0) return;
    const int qp = s.mb.qp_c[p];
    dst = s.res + 256 + 64 * p + (b >> 1) * 32 + (b & 1) * 4; stride = 8;
    const int q0 = dq4 (qp, 0);
    int a = S16 (dc0 * q0), bb = S16 (dc1 * q0), cc = S16 (dc2 * q0), d = S16 (dc3 * q0);
    const int e = a - bb; a += bb; bb = cc - d; cc += d;
    c[0] = S16 (b == 0 ? (a + cc) >> 1 : (b == 1 ? (e + bb) >> 1 : (b == 2 ? (a - cc) >> 1 : (e - bb) >> 1)));
    if (!ac_zero) {
#pragma unroll
      for (int j = 1; j < 16; j++) c[j] = S16 ((int) (int16_t) (w[j >> 1] >> ((j & 1) * 16)) * dq4 (qp, j));
    }
  }
  if (ac_zero) {
    const uint32_t v = (uint32_t) (uint16_t) (int16_t) ((32 + c[0]) >> 6);
    const uint2 vv = make_uint2 (v | (v << 16), v | (v << 16));
#pragma unroll
    for (int y = 0; y < 4; y++) *reinterpret_cast<uint2*> (dst + y * stride) = vv;
  } else idct4x4_res (c, dst, stride);
}

// ---- commit of the unfiltered macroblock into the "cur" tile ----------------------------------------------------------------
// lane -> luma row r = lane >> 1, columns 8h .. 8h + 7 (h = lane & 1); chroma plane lane >> 4, row (lane >> 1) & 7, columns 4h .. 4h + 3
__device__ __forceinline__ void commit_words (RowCtx& s, int lane, uint32_t y0, uint32_t y1, uint32_t c0, bool add0, bool add1, bool addc) {
  const int r = lane >> 1, h = lane & 1;
  if (add0) y0 = add_res4 (y0, s.res + r * 16 + 8 * h);
  if (add1) y1 = add_res4 (y1, s.res + r * 16 + 8 * h + 4);
  uint32_t* dy = reinterpret_cast<uint32_t*> (s.ry + (RY_TOP + r) * RY + 16 + 8 * h);
  dy[0] = y0; dy[1] = y1;
  const int p = lane >> 4, cr = (lane >> 1) & 7;
  if (addc) c0 = add_res4 (c0, s.res + 256 + 64 * p + cr * 8 + 4 * h);
  *reinterpret_cast<uint32_t*> (s.rc[p] + (RC_TOP + cr) * RC + 8 + 4 * h) = c0;
}

// prediction (+ residual) from s.pred / s.res, or luma from the intra work tile
__device__ __forceinline__ void commit_from_pred (RowCtx& s, int lane, bool add_luma, bool add_chroma, bool luma_from_tile) {
  const int r = lane >> 1, h = lane & 1, p = lane >> 4, cr = (lane >> 1) & 7;
  uint32_t y0, y1;
  if (luma_from_tile) {
    const uint8_t* t = s.tile + (r + 1) * TS + 8 * h + 1;
    y0 = (uint32_t)t[0] | ((uint32_t)t[1] << 8) | ((uint32_t)t[2] << 16) | ((uint32_t)t[3] << 24);
    y1 = (uint32_t)t[4] | ((uint32_t)t[5] << 8) | ((uint32_t)t[6] << 16) | ((uint32_t)t[7] << 24);
  } else {
    const uint32_t* q = reinterpret_cast<const uint32_t*> (s.pred + r * 16 + 8 * h);
    y0 = q[0]; y1 = q[1];
  }
  const uint32_t c0 = *reinterpret_cast<const uint32_t*> (s.pred + 256 + 64 * p + cr * 8 + 4 * h);
  const bool al = add_luma && !luma_from_tile;
  commit_words (s, lane, y0, y1, c0, al, al, add_chroma);
}

// ---- P_L0_16x16 / P_Skip fast path ---------------------------------------------------------------------------------------------
// Returns false (nothing done) when the macroblock needs the general routine.
__device__ __forceinline__ bool w_inter16_fast (RowCtx& s, const PicDev& pic, int mbx, int mby, int lane) {
  if (s.sl_wp) return false;
  const int W = pic.mb_w * 16, H = pic.mb_h * 16, CW = W >> 1, CH = H >> 1;
  // BaseMC: the motion vector is clamped into the 32-pixel border (rec_mb.cpp:251-252); rows are clamped at read below
  const int fx = clip3i (((mbx * 16) << 2) + s.mb.mv[0][0], (-32 + 2) * 4, (W + 32 - 19) << 2);
  const int fy = clip3i (((mby * 16) << 2) + s.mb.mv[0][1], (-32 + 2) * 4, (H + 32 - 19) << 2);
  const int x0 = fx >> 2, y0 = fy >> 2, cx = fx >> 3, cy = fy >> 3, dx = fx & 7, dy = fy & 7;
  const bool frac = ((fx | fy) & 3) != 0;
  if (cx < 0 || cx + 8 + (dx != 0) > CW) return false;
  if (frac ? (x0 - 2 < 0 || x0 + 19 > W) : (x0 < 0 || x0 + 16 > W)) return false;
  const int slot = s.sl_slot[s.mb.ref_idx[0] & 15];
  const uint8_t* ref = plane_ptr (pic, slot, 0);
  const int r = lane >> 1, h = lane & 1;
  // ---- issue every global load first ----
  // chroma: two rows x two aligned words
  const int p = lane >> 4, cr = (lane >> 1) & 7;
  const uint8_t* cref = plane_ptr (pic, slot, 1 + p);
  const int ccol = cx + 4 * h, cal = ccol & ~3;
  const unsigned csh = (unsigned) (ccol & 3) * 8u;
  const uint32_t* ca = reinterpret_cast<const uint32_t*> (cref + (size_t)clip3i (cy + cr, 0, CH - 1) * CW + cal);
  const uint32_t* cb = reinterpret_cast<const uint32_t*> (cref + (size_t)clip3i (cy + cr + 1, 0, CH - 1) * CW + cal);
  const uint32_t c00 = __ldg (ca), c01 = __ldg (ca + 1);
  uint32_t c10 = c00, c11 = c01;
  if (dy) { c10 = __ldg (cb); c11 = __ldg (cb + 1); }
  uint32_t ya = 0, yb = 0;
  if (!frac) {                                    // integer-pel luma: a copy, row clamped
    const int col = x0 + 8 * h, al = col & ~3;
    const unsigned sh = (unsigned) (col & 3) * 8u;
    const uint32_t* q = reinterpret_cast<const uint32_t*> (ref + (size_t)clip3i (y0 + r, 0, H - 1) * W + al);
    const uint32_t w0 = __ldg (q), w1 = __ldg (q + 1), w2 = __ldg (q + 2);
    ya = __funnelshift_r (w0, w1, sh); yb = __funnelshift_r (w1, w2, sh);
  } else {                                        // 21 x 21 window as aligned words into shared memory (rows clamped)
    const int ix = x0 - 2, iy = y0 - 2, ax = ix & ~3;
    uint32_t lv[4];
#pragma unroll
    for (int i = 0; i < 4; i++) {                  // 21 rows x 6 words = 126 words
      const int k = lane + 32 * i, wy = (k * 43) >> 8, wx = k - wy * 6;      // k / 6 for k < 128
      lv[i] = 0;
      if (k < 126) lv[i] = __ldg (reinterpret_cast<const uint32_t*> (ref + (size_t)clip3i (iy + wy, 0, H - 1) * W + ax) + wx);
    }
#pragma unroll
    for (int i = 0; i < 4; i++) {
      const int k = lane + 32 * i, wy = (k * 43) >> 8, wx = k - wy * 6;
      if (k < 126) *reinterpret_cast<uint32_t*> (s.lwin + wy * LW_STRIDE + 4 * wx) = lv[i];
    }
  }
  // ---- residual while the loads are in flight ----
  const uint32_t mask = s.mb.coded_mask;
  const bool has_res = s.mb.cbp != 0 && mask != 0;
  if (has_res) w_residual_direct (s, pic.coefs, lane);
  __syncwarp();
  // ---- chroma: bilinear, McChroma_c (mc.cpp:368-378), four samples ----
  uint32_t cw;
  {
    const uint32_t R0 = __funnelshift_r (c00, c01, csh), R1 = __funnelshift_r (c10, c11, csh);
    if ((dx | dy) == 0) cw = R0;
    else {
      const uint32_t R0s = __funnelshift_rc (c00, c01, csh + 8u), R1s = __funnelshift_rc (c10, c11, csh + 8u);
      const uint32_t A = (uint32_t) ((8 - dx) * (8 - dy)), B = (uint32_t) (dx * (8 - dy)), C = (uint32_t) ((8 - dx) * dy), D = (uint32_t) (dx * dy);
      const uint32_t wab = A | (B << 8), wcd = C | (D << 8);
      const uint32_t X01 = __byte_perm (R0, R0s, 0x5140), X23 = __byte_perm (R0, R0s, 0x7362);
      const uint32_t Y01 = __byte_perm (R1, R1s, 0x5140), Y23 = __byte_perm (R1, R1s, 0x7362);
      const uint32_t o0 = dp4a_uu (Y01, wcd, dp4a_uu (X01, wab, 32u)) >> 6, o1 = dp4a_uu (Y01, wcd << 16, dp4a_uu (X01, wab << 16, 32u)) >> 6;
      const uint32_t o2 = dp4a_uu (Y23, wcd, dp4a_uu (X23, wab, 32u)) >> 6, o3 = dp4a_uu (Y23, wcd << 16, dp4a_uu (X23, wab << 16, 32u)) >> 6;
      cw = o0 | (o1 << 8) | (o2 << 16) | (o3 << 24);
    }
  }
  // ---- luma ----
  if (frac) {
    const int fr = ((fx & 3) << 2) | (fy & 3), woff = (x0 - 2) & 3;
    int v[4];
    mc_luma_quad (s.lwin + (2 + r) * LW_STRIDE + woff + 2 + 8 * h, LW_STRIDE, fr, v);
    ya = (uint32_t)v[0] | ((uint32_t)v[1] << 8) | ((uint32_t)v[2] << 16) | ((uint32_t)v[3] << 24);
    mc_luma_quad (s.lwin + (2 + r) * LW_STRIDE + woff + 2 + 8 * h + 4, LW_STRIDE, fr, v);
    yb = (uint32_t)v[0] | ((uint32_t)v[1] << 8) | ((uint32_t)v[2] << 16) | ((uint32_t)v[3] << 24);
  }
  // ---- residual add where a block exists, into the cur tile ----
  bool a0 = false, a1 = false, ac = false;
  if (has_res) {
    const int by = r >> 2, z0 = ((by & 1) << 1) | ((by >> 1) << 3) | (h << 2);     // z index of block (bx = 2h, by); bx + 1 -> z + 1
    a0 = (mask >> z0) & 1u; a1 = (mask >> (z0 + 1)) & 1u;
    ac = ((mask >> (16 + 4 * p)) & 15u) != 0;
  }
  commit_words (s, lane, ya, yb, cw, a0, a1, ac);
  return true;
}

// ---- inter macroblock, general routine: partitions / weighting / border windows (recon_rows.cuh) -----------------------------
__device__ LH264_NOINLINE void w_inter_general (RowCtx& s, const PicDev& pic, const lh264_slice_rec* sl, int mbx, int mby, int lane) {
  const int t = s.mb.mb_type;
  const bool has_res = s.mb.cbp != 0;
  uint32_t lv[6];
  if (has_res) w_issue_levels (s, pic.coefs, lane, lv);          // in flight during the first partition
  bool res_done = !has_res;
  auto residual = [&] () {
    if (res_done) return;
    res_done = true;
    w_store_levels (s, lane, lv);
    if (lane < 24) { if (s.sl_scal & LH264_SLF_SCALING) residual_block_scaled (reinterpret_cast<MbShared&> (s), lane, sl); else residual_block (reinterpret_cast<MbShared&> (s), lane); }
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
  commit_from_pred (s, lane, has_res, has_res, false);
}

// ---- intra macroblock ----------------------------------------------------------------------------------------------------------
// WelsMbIntraPredictionConstruction (decode_slice.cpp:313-334), RecI4x4Mb / RecI8x8Mb / RecI16x16Mb / RecChroma
// (rec_mb.cpp:64-219,:547-575), predictors get_intra_predictor.cpp:54-1155.  Neighbours are UNFILTERED samples: the left
// column from s.ucol, the row above (and the corner) from the line buffer of row y - 1.
#define F3(a, b, c) (((a) + 2 * (b) + (c) + 2) >> 2)
#define F2(a, b) (((a) + (b) + 1) >> 1)
__device__ LH264_NOINLINE void w_intra_row (RowCtx& s, const PicDev& pic, int mbx, int mby, int lane) {
  uint8_t* tile = s.tile;
  const int mb_type = s.mb.mb_type;
  const int W = pic.mb_w * 16, CW = W >> 1;
  w_load_levels (s, pic.coefs, lane);
  if (mb_type == LH264_MB_IPCM) {            // samples travel in the coefficient pool
    for (int i = lane; i < 384; i += 32) s.pred[i] = (uint8_t)s.lev[i];
    __syncwarp();
    commit_from_pred (s, lane, false, false, false);
    return;
  }
  // neighbour samples -> shared (coordinates clamped: a resolved mode never uses an unavailable neighbour)
  const int ly = mby > 0 ? mby - 1 : 0;
  const uint8_t* LY = line_ptr (pic, ly, 0);
  for (int t = lane; t < 77; t += 32) {
    if (t < 25) tile[t] = __ldcg (LY + clip3i (mbx * 16 - 1 + t, 0, W - 1));
    else if (t < 41) tile[(t - 25 + 1) * TS] = s.ucol[t - 25];
    else {
      const int c = t - 41, p = c / 18, k = c % 18;        // k 0..8 top (x = -1..7), 9..17 left (y = -1..7)
      const uint8_t* LC = line_ptr (pic, ly, 1 + p);
      if (k < 9) s.cedge[p][0][k] = __ldcg (LC + clip3i (mbx * 8 - 1 + k, 0, CW - 1));
      else if (k == 9) s.cedge[p][1][0] = __ldcg (LC + max (mbx * 8 - 1, 0));
      else s.cedge[p][1][k - 9] = s.ucol[16 + 8 * p + (k - 10)];
    }
  }
  if (lane < 24) { if (s.sl_scal & LH264_SLF_SCALING) residual_block_scaled (reinterpret_cast<MbShared&> (s), lane, pic.slices + s.mb.slice_idx); else residual_block (reinterpret_cast<MbShared&> (s), lane); }
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
    commit_from_pred (s, lane, true, true, false);
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
  commit_from_pred (s, lane, false, cc == 1 || cc == 2, true);
}
#undef F2
#undef F3

// ---- deblocking of the macroblock in the cur tile ----------------------------------------------------------------------------
// WelsDeblockingMb (deblocking.cpp:815-870), bS derivation (:160-353), FilteringEdge* and DeblockLuma/Chroma{Lt4,Eq4}_c
// (deblocking_common.cpp:5-246).  Samples: s.ry / s.rc (prev | cur, rows of the macroblock above on top of cur).
// Returns true when the top edge was filtered (the rows above must be written back).
// One line across an edge, luma and chroma in ONE instruction stream (the chroma lanes of the warp run the same code with
// `chroma` set instead of a divergent second routine): DeblockLumaLt4_c / DeblockLumaEq4_c / DeblockChromaLt4_c /
// DeblockChromaEq4_c, deblocking_common.cpp:5-246.
__device__ __forceinline__ void edge_line (uint8_t* p, int sx, int bS, int ia, int alpha, int beta, bool chroma) {
  const int p0 = p[-sx], p1 = p[-2 * sx], q0 = p[0], q1 = p[sx];
  const int d = abs (p0 - q0);
  if (! (d < alpha && abs (p1 - p0) < beta && abs (q1 - q0) < beta)) return;
  if (bS == 4) {
    if (chroma || d >= ((alpha >> 2) + 2)) {
      p[-sx] = (uint8_t) ((2 * p1 + p0 + q1 + 2) >> 2); p[0] = (uint8_t) ((2 * q1 + q0 + p1 + 2) >> 2);
      return;
    }
    const int p2 = p[-3 * sx], q2 = p[2 * sx];
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
    return;
  }
  const int tc0 = c_tc0[ia][bS];
  int tc = tc0 + (chroma ? 1 : 0);
  if (!chroma) {
    const int p2 = p[-3 * sx], q2 = p[2 * sx];
    if (abs (p2 - p0) < beta) { p[-2 * sx] = (uint8_t) (p1 + clip3i ((p2 + ((p0 + q0 + 1) >> 1) - (p1 << 1)) >> 1, -tc0, tc0)); tc++; }
    if (abs (q2 - q0) < beta) { p[sx] = (uint8_t) (q1 + clip3i ((q2 + ((p0 + q0 + 1) >> 1) - (q1 << 1)) >> 1, -tc0, tc0)); tc++; }
  }
  const int dl = clip3i ((((q0 - p0) << 2) + (p1 - q1) + 4) >> 3, -tc, tc);
  p[-sx] = (uint8_t)clip255 (p0 + dl); p[0] = (uint8_t)clip255 (q0 - dl);
}

__device__ __forceinline__ bool w_deblock_tile (RowCtx& s, int mbx, int mby, int lane) {
  const lh264_mb_rec* m = &s.mb; const lh264_mb_rec* ml = &s.left; const lh264_mb_rec* mt = &s.top;
  const int mb_type = m->mb_type;
  if (mb_type == LH264_MB_NONE) return false;
  const int idc = s.sl_idc;
  if (idc == 1) return false;
  const int aoff = s.sl_aoff, boff = s.sl_boff;
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
  } else if (mb_type != LH264_MB_PSKIP && ! (mb_type == LH264_MB_P16x16 && m->nz_mask == 0) && ! (t8 && e != 2)) {      // (one vector, no coefficients: no inner edge)
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
  // edges with anything to filter, in the standard's order: bit 4 * k of `live` for edge k = dir * 4 + edge
  unsigned live = __ballot_sync (0xffffffffu, bs != 0);
  if (live == 0) return false;
  live = (live | (live >> 1) | (live >> 2) | (live >> 3)) & 0x11111111u;
  if (t8) live &= 0x01010101u;                         // 8x8 transform: edges 1 and 3 do not exist
  const bool top_done = (live >> 16) & 1u;
  const bool chroma = lane >= 16;
  const int cplane = (lane >> 3) & 1;                  // chroma lanes: 16..23 Cb, 24..31 Cr
  const int qp_own = chroma ? m->qp_c[cplane] : m->qp_y;
  const int seg = chroma ? ((lane & 7) >> 1) : (lane >> 2);
  const int l = chroma ? (lane & 7) : lane;            // line index along the edge
  uint8_t* const plane = chroma ? s.rc[cplane] + RC_TOP * RC + 8 : s.ry + RY_TOP * RY + 16;   // sample (0, 0) of the cur tile
  const int stride = chroma ? RC : RY;
#pragma unroll 1
  while (live) {
    const int k = (__ffs (live) - 1) >> 2;             // dir * 4 + edge
    live &= live - 1;
    const int d = k >> 2, ed = k & 3;
    const int my_bs = __shfl_sync (0xffffffffu, bs, 4 * k + seg);
    if (my_bs && ! (chroma && (ed & 1))) {
      const lh264_mb_rec* nb = d ? mt : ml;
      int qp = qp_own;
      if (ed == 0) qp = (qp + (chroma ? nb->qp_c[cplane] : nb->qp_y) + 1) >> 1;
      const int ia = clip3i (qp + aoff, 0, 51), alpha = c_alpha[ia], beta = c_beta[clip3i (qp + boff, 0, 51)];
      if (alpha | beta) {
        const int pos = chroma ? 2 * ed : 4 * ed;      // edge position inside the tile
        uint8_t* p = d ? plane + pos * stride + l : plane + l * stride + pos;
        edge_line (p, d ? stride : 1, my_bs, ia, alpha, beta, chroma);
      }
    }
    __syncwarp();
  }
  return top_done;
}

// ---- the kernel ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__ (ROWS_WARPS * 32, 2) k_decode_rows (const PicDev* __restrict__ pics, int n_pics, int max_rows, DecodeSync* sync, unsigned int* error) {
  __shared__ RowCtx ctx[ROWS_WARPS];
  const int lane = threadIdx.x & 31;
  RowCtx& s = ctx[threadIdx.x >> 5];
  int* progress = reinterpret_cast<int*> (sync + 1);
  const unsigned total = (unsigned)n_pics * max_rows;
  const int r = lane >> 1, h = lane & 1, cp = lane >> 4, cr = (lane >> 1) & 7;       // store mapping (see commit_words)
  for (;;) {
    unsigned t = 0;
    if (lane == 0) t = atomicAdd (&sync->ticket, 1u);
    t = __shfl_sync (0xffffffffu, t, 0);
    if (t >= total) return;
    const int mby = t / n_pics, pi = t - mby * n_pics;
    __syncwarp();
    if (lane < (int) (sizeof (PicDev) / 4)) reinterpret_cast<uint32_t*> (&s.pic)[lane] = reinterpret_cast<const uint32_t*> (pics + pi)[lane];
    __syncwarp();
    const PicDev& pic = s.pic;
    if (mby >= pic.mb_h) continue;
    const int mb_w = pic.mb_w, W = mb_w * 16, CW = W >> 1;
    int* my = progress + (size_t)pi * max_rows + mby;
    const int* above = my - 1;
    int seen = 0;                                   // progress of the row above known so far
    int cached_slice = -1;                          // slice whose fields are in s.sl_*
    const uint32_t* row = reinterpret_cast<const uint32_t*> (pic.mbs + (size_t)mby * mb_w);
    const uint32_t* row_top = mby > 0 ? reinterpret_cast<const uint32_t*> (pic.mbs + (size_t) (mby - 1) * mb_w) : row;
    uint32_t recw = __ldg (row + lane), topw = __ldg (row_top + lane);   // the records of the next MB are always in flight
    uint8_t* Y = plane_ptr (pic, pic.cur_slot, 0) + (size_t)mby * 16 * W;
    uint8_t* C = plane_ptr (pic, pic.cur_slot, 1 + cp) + (size_t)mby * 8 * CW;
    {
      __syncwarp();
      s.lp[0][lane] = reinterpret_cast<unsigned long long> (Y + (size_t)r * W + 8 * h);
      s.lp[1][lane] = reinterpret_cast<unsigned long long> (C + (size_t)cr * CW + 4 * h);
      uint8_t* tp = Y;
      if (mby > 0) {
        if (lane < 16) tp = Y - (size_t) (4 - (lane >> 2)) * W + 4 * (lane & 3);
        else if (lane < 24) { const int q = lane - 16; tp = plane_ptr (pic, pic.cur_slot, 1 + (q >> 2)) + ((size_t)mby * 8 - 2 + ((q >> 1) & 1)) * CW + 4 * (q & 1); }
      }
      s.lp[2][lane] = reinterpret_cast<unsigned long long> (tp);
      uint8_t* lb = line_ptr (pic, mby, 0) + 4 * (lane & 3);
      if (lane >= 4) { const int q = (lane - 4) & 3; lb = line_ptr (pic, mby, 1 + (q >> 1)) + 4 * (q & 1); }
      s.lp[3][lane] = reinterpret_cast<unsigned long long> (lb);
      __syncwarp();
    }
#pragma unroll 1
    for (int mbx = 0; mbx < mb_w; mbx++) {
      __syncwarp();
      if (mbx > 0) reinterpret_cast<uint32_t*> (&s.left)[lane] = reinterpret_cast<const uint32_t*> (&s.mb)[lane];
      __syncwarp();
      reinterpret_cast<uint32_t*> (&s.mb)[lane] = recw;
      reinterpret_cast<uint32_t*> (&s.top)[lane] = topw;
      __syncwarp();
      if (mbx + 1 < mb_w) { recw = __ldg (row + (mbx + 1) * 32 + lane); topw = __ldg (row_top + (mbx + 1) * 32 + lane); }
      // a malformed record (C-ABI caller, parser bug) must not become an out-of-bounds access: validate what indexes memory
      {
        const uint32_t cm = s.mb.coded_mask;
        // (a macroblock that was not decoded keeps whatever its record held before: only its type is looked at)
        const bool bad = s.mb.mb_type > LH264_MB_PSKIP || (s.mb.mb_type != LH264_MB_NONE && ((int)s.mb.slice_idx >= pic.n_slices || (cm >> 24) != 0 ||
                         (cm != 0 && (unsigned long long)s.mb.coef_off + (unsigned)__popc (cm) > (unsigned long long)pic.n_coef_blocks)));
        if (bad) {
          __syncwarp();
          if (lane == 0) { s.mb.mb_type = LH264_MB_NONE; atomicOr (error, 1u); }
          __syncwarp();
        }
      }
      const int mb_type = s.mb.mb_type;
      if (mb_type != LH264_MB_NONE && (int)s.mb.slice_idx != cached_slice) {      // uniform: every lane reads the same record
        cached_slice = s.mb.slice_idx;
        const lh264_slice_rec* sl = pic.slices + cached_slice;
        __syncwarp();
        if (lane < LH264_MAX_REFS) s.sl_slot[lane] = (int16_t)clip3i (__ldg (&sl->ref_slot[lane]), 0, pic.n_slots - 1);
        else if (lane == 16) { s.sl_wp = __ldg (&sl->weighted_pred); s.sl_idc = __ldg (&sl->disable_deblocking_idc); s.sl_aoff = __ldg (&sl->alpha_c0_offset); s.sl_boff = __ldg (&sl->beta_offset); s.sl_scal = __ldg (&sl->scaling_flags); }
        __syncwarp();
      }
      // the row above must have stored macroblock x (its final state but for the edge filtered here) and written the
      // unfiltered bottom row of macroblock x + 1: both are true once it has finished macroblock x + 1, i.e. published x + 1
      if (mby > 0) {
        const int need = mbx + 1;
        if (seen < need) {
          if (lane == 0) { while ((seen = ld_acquire (above)) < need) __nanosleep (64); }
          seen = __shfl_sync (0xffffffffu, seen, 0);
        }
      }
      // rows 12..15 (6..7) of the macroblock above, for the top edge: issued now, parked in the tile arrays after reconstruction
      uint32_t tw = 0;
      if (mby > 0 && lane < 24) tw = __ldcg (reinterpret_cast<const uint32_t*> (reinterpret_cast<const uint8_t*> (s.lp[2][lane]) + mbx * (lane < 16 ? 16 : 8)));
      // ---- unfiltered reconstruction into the cur tile ----
      if (LH264_MB_IS_INTER (mb_type)) {
        const lh264_slice_rec* sl = pic.slices + s.mb.slice_idx;
        bool done = false;
        if ((mb_type == LH264_MB_P16x16 || mb_type == LH264_MB_PSKIP) && ! (s.mb.flags & LH264_MBF_TRANSFORM8x8) && !s.sl_scal) done = w_inter16_fast (s, pic, mbx, mby, lane);
        if (!done) w_inter_general (s, pic, sl, mbx, mby, lane);
      } else if (LH264_MB_IS_INTRA (mb_type)) w_intra_row (s, pic, mbx, mby, lane);
      else {                                        // not decoded in this picture: the samples stay as they are
        const uint32_t* q = reinterpret_cast<const uint32_t*> (Y + (size_t)r * W + mbx * 16 + 8 * h);
        const uint32_t y0 = __ldcg (q), y1 = __ldcg (q + 1);
        const uint32_t c0 = __ldcg (reinterpret_cast<const uint32_t*> (C + (size_t)cr * CW + mbx * 8 + 4 * h));
        commit_words (s, lane, y0, y1, c0, false, false, false);
      }
      if (mby > 0) {
        if (lane < 16) *reinterpret_cast<uint32_t*> (s.ry + (lane >> 2) * RY + 16 + 4 * (lane & 3)) = tw;
        else if (lane < 24) { const int q = lane - 16; *reinterpret_cast<uint32_t*> (s.rc[q >> 2] + ((q >> 1) & 1) * RC + 8 + 4 * (q & 1)) = tw; }
      }
      __syncwarp();
      // ---- what the neighbours' intra prediction needs of U(x, y): right column (shared), bottom row (line buffer) ----
      if (lane < 16) s.ucol[lane] = s.ry[(RY_TOP + lane) * RY + 31];
      else s.ucol[lane] = s.rc[(lane - 16) >> 3][(RC_TOP + (lane & 7)) * RC + 15];
      if (lane < 4) __stcg (reinterpret_cast<uint32_t*> (reinterpret_cast<uint8_t*> (s.lp[3][lane]) + mbx * 16), *reinterpret_cast<const uint32_t*> (s.ry + (RY_TOP + 15) * RY + 16 + 4 * lane));
      else if (lane < 8) {
        const int q = lane - 4;
        __stcg (reinterpret_cast<uint32_t*> (reinterpret_cast<uint8_t*> (s.lp[3][lane]) + mbx * 8), *reinterpret_cast<const uint32_t*> (s.rc[q >> 1] + (RC_TOP + 7) * RC + 8 + 4 * (q & 1)));
      }
      __syncwarp();
      // ---- deblocking of macroblock x: left edge across prev | cur, top edge across the rows above | cur ----
      const bool top_done = w_deblock_tile (s, mbx, mby, lane);
      __syncwarp();
      // ---- macroblock x - 1 is final (but for its bottom edge): store it; rows of the macroblock above go back; cur -> prev ----
      {
        uint32_t* py = reinterpret_cast<uint32_t*> (s.ry + (RY_TOP + r) * RY + 8 * h);
        uint32_t* pc = reinterpret_cast<uint32_t*> (s.rc[cp] + (RC_TOP + cr) * RC + 4 * h);
        if (mbx > 0) {
          *reinterpret_cast<uint2*> (reinterpret_cast<uint8_t*> (s.lp[0][lane]) + (mbx - 1) * 16) = make_uint2 (py[0], py[1]);
          *reinterpret_cast<uint32_t*> (reinterpret_cast<uint8_t*> (s.lp[1][lane]) + (mbx - 1) * 8) = pc[0];
        }
        if (top_done) {
          // the words of the rows above go back where they were read from: luma rows -3..-1 = the loads of lanes 4..15, chroma row -1
          // of plane q >> 1 = the loads of lanes 16 + 4 (q >> 1) + 2 + (q & 1)
          if (lane < 12) { const int tr = 1 + (lane >> 2); *reinterpret_cast<uint32_t*> (reinterpret_cast<uint8_t*> (s.lp[2][lane + 4]) + mbx * 16) = *reinterpret_cast<const uint32_t*> (s.ry + tr * RY + 16 + 4 * (lane & 3)); }
          else if (lane < 16) {
            const int q = lane - 12;
            *reinterpret_cast<uint32_t*> (reinterpret_cast<uint8_t*> (s.lp[2][16 + 4 * (q >> 1) + 2 + (q & 1)]) + mbx * 8) = *reinterpret_cast<const uint32_t*> (s.rc[q >> 1] + 1 * RC + 8 + 4 * (q & 1));
          }
        }
        const uint32_t n0 = py[4], n1 = py[5], nc = pc[2];      // cur = prev + 16 bytes (luma) / + 8 bytes (chroma)
        __syncwarp();
        py[0] = n0; py[1] = n1; pc[0] = nc;
      }
      __syncwarp();
      // progress = macroblocks of this row stored so far; a consumer at macroblock x needs x + 1.  A release store is a fence:
      // every fourth macroblock is often enough (rows of one picture are picked up ~n_pics tickets apart)
      if (lane == 0 && mbx > 0 && (mbx & 3) == 0) st_release (my, mbx);
    }
    // the last macroblock of the row
    __syncwarp();
    {
      const uint32_t* py = reinterpret_cast<const uint32_t*> (s.ry + (RY_TOP + r) * RY + 8 * h);
      const uint32_t* pc = reinterpret_cast<const uint32_t*> (s.rc[cp] + (RC_TOP + cr) * RC + 4 * h);
      *reinterpret_cast<uint2*> (reinterpret_cast<uint8_t*> (s.lp[0][lane]) + (mb_w - 1) * 16) = make_uint2 (py[0], py[1]);
      *reinterpret_cast<uint32_t*> (reinterpret_cast<uint8_t*> (s.lp[1][lane]) + (mb_w - 1) * 8) = pc[0];
    }
    __syncwarp();
    if (lane == 0) st_release (my, mb_w);
  }
}

} // namespace lh264
