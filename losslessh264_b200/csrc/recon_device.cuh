// recon_device.cuh — device-side arithmetic of the sm_100a macroblock-reconstruction kernels.
//
// Each routine names the reference code whose results it must reproduce bit-exactly (paths
// relative to the reference root).  The decomposition is GPU-first: per-4x4-block reference
// windows staged in shared memory, one thread per 4-pixel row for motion compensation, one
// thread per coefficient block for dequant+IDCT, warp-lane-per-line deblocking.
#pragma once
#include <stdint.h>
#ifndef LH264_CPU_EMU          // tests/emu compiles these sources with g++ behind a CUDA-model shim (test infrastructure)
#include <cuda_runtime.h>
#endif
#include "../../include/lh264_b200.h"

namespace lh264 {

__device__ __forceinline__ int clip3i (int v, int lo, int hi) { return min (max (v, lo), hi); }
__device__ __forceinline__ int clip255 (int v) { return min (max (v, 0), 255); }

// ---- tables ---------------------------------------------------------------------------------------
// LevelScale4x4 (g_kuiDequantCoeff = this << qp/6, codec/common/src/common_tables.cpp:208-235)
__constant__ uint8_t c_scale4[6][3] = { {10, 16, 13}, {11, 18, 14}, {13, 20, 16}, {14, 23, 18}, {16, 25, 20}, {18, 29, 23} };
// normAdjust8x8 (g_kuiDequantCoeff8x8 = 16 * this, common_tables.cpp:237+)
__constant__ uint8_t c_scale8[6][6] = { {20, 18, 32, 19, 25, 24}, {22, 19, 35, 21, 28, 26}, {26, 23, 42, 24, 33, 31},
  {28, 25, 45, 26, 35, 33}, {32, 28, 51, 30, 40, 38}, {36, 32, 58, 34, 46, 43} };
// alpha / beta / tc0 of the loop filter (codec/decoder/core/src/deblocking.cpp:89-125, without the +12 guard band)
__constant__ uint8_t c_alpha[52] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 4, 4, 5, 6, 7, 8, 9, 10, 12, 13, 15, 17, 20, 22,
                                    25, 28, 32, 36, 40, 45, 50, 56, 63, 71, 80, 90, 101, 113, 127, 144, 162, 182, 203, 226, 255, 255
                                   };
__constant__ uint8_t c_beta[52] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 6, 6, 7, 7,
                                   8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13, 14, 14, 15, 15, 16, 16, 17, 17, 18, 18
                                  };
__constant__ int8_t c_tc0[52][4] = {
  { -1, 0, 0, 0}, { -1, 0, 0, 0}, { -1, 0, 0, 0}, { -1, 0, 0, 0}, { -1, 0, 0, 0}, { -1, 0, 0, 0}, { -1, 0, 0, 0}, { -1, 0, 0, 0},
  { -1, 0, 0, 0}, { -1, 0, 0, 0}, { -1, 0, 0, 0}, { -1, 0, 0, 0}, { -1, 0, 0, 0}, { -1, 0, 0, 0}, { -1, 0, 0, 0}, { -1, 0, 0, 0},
  { -1, 0, 0, 0}, { -1, 0, 0, 1}, { -1, 0, 0, 1}, { -1, 0, 0, 1}, { -1, 0, 0, 1}, { -1, 0, 1, 1}, { -1, 0, 1, 1}, { -1, 1, 1, 1},
  { -1, 1, 1, 1}, { -1, 1, 1, 1}, { -1, 1, 1, 1}, { -1, 1, 1, 2}, { -1, 1, 1, 2}, { -1, 1, 1, 2}, { -1, 1, 1, 2}, { -1, 1, 2, 3},
  { -1, 1, 2, 3}, { -1, 2, 2, 3}, { -1, 2, 2, 4}, { -1, 2, 3, 4}, { -1, 2, 3, 4}, { -1, 3, 3, 5}, { -1, 3, 4, 6}, { -1, 3, 4, 6},
  { -1, 4, 5, 7}, { -1, 4, 5, 8}, { -1, 4, 6, 9}, { -1, 5, 7, 10}, { -1, 6, 8, 11}, { -1, 6, 8, 13}, { -1, 7, 10, 14}, { -1, 8, 11, 16},
  { -1, 9, 12, 18}, { -1, 10, 13, 20}, { -1, 11, 15, 23}, { -1, 13, 17, 25}
};

// 4x4 block z (decode order) -> pixel offset inside the MB
__device__ __forceinline__ int blk_x (int z) { return ((z & 1) << 2) | ((z & 4) << 1); }
__device__ __forceinline__ int blk_y (int z) { return ((z & 2) << 1) | ((z & 8)); }

__device__ __forceinline__ int dq4 (int qp, int pos) {
  // class by (pos & 7): a c a c / c b c b
  const int cls = (0x12122020 >> ((pos & 7) * 4)) & 3;
  return (int)c_scale4[qp % 6][cls] << (qp / 6);
}
__device__ __forceinline__ int dq8 (int qp, int pos) {
  const int i = pos >> 3, j = pos & 7;
  int c;
  if ((i & 3) == 0 && (j & 3) == 0) c = 0;
  else if ((i & 1) && (j & 1)) c = 1;
  else if ((i & 3) == 2 && (j & 3) == 2) c = 2;
  else if (((i & 3) == 0 && (j & 1)) || ((i & 1) && (j & 3) == 0)) c = 3;
  else if (((i & 3) == 0 && (j & 3) == 2) || ((i & 3) == 2 && (j & 3) == 0)) c = 4;
  else c = 5;
  return 16 * (int)c_scale8[qp % 6][c];
}

// ---- inverse transforms ------------------------------------------------------------------------------
// IdctResAddPred_c (codec/decoder/core/src/decode_mb_aux.cpp:42-77): c[16] dequantised (int16
// values), row pass stored as int16, residual = (x + 32) >> 6.  Writes the 16 residuals to res[].
__device__ __forceinline__ void idct4x4_res (const int c[16], int16_t* res, int rstride) {
  int t[16];
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const int a = c[4 * i], b = c[4 * i + 1], d = c[4 * i + 2], e = c[4 * i + 3];
    const int e0 = a + d, e1 = a - d, e2 = (b >> 1) - e, e3 = b + (e >> 1);
    t[4 * i] = (int16_t) (e0 + e3); t[4 * i + 1] = (int16_t) (e1 + e2);
    t[4 * i + 2] = (int16_t) (e1 - e2); t[4 * i + 3] = (int16_t) (e0 - e3);
  }
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const int e0 = t[i] + t[i + 8], e1 = t[i] - t[i + 8], e2 = (t[i + 4] >> 1) - t[i + 12], e3 = t[i + 4] + (t[i + 12] >> 1);
    res[i] = (int16_t) ((32 + e0 + e3) >> 6);
    res[i + rstride] = (int16_t) ((32 + e1 + e2) >> 6);
    res[i + 2 * rstride] = (int16_t) ((32 + e1 - e2) >> 6);
    res[i + 3 * rstride] = (int16_t) ((32 + e0 - e3) >> 6);
  }
}

// one 8-point pass of IdctResAddPred8x8_c (decode_mb_aux.cpp:79-167); every intermediate is int16
#define S16(x) ((int)(int16_t)(x))
__device__ __forceinline__ void idct8_1d (const int p[8], int o[8]) {
  int a0 = S16 (p[0] + p[4]), a1 = S16 (p[0] - p[4]), a2 = S16 (p[6] - (p[2] >> 1)), a3 = S16 (p[2] + (p[6] >> 1));
  const int b0 = S16 (a0 + a3), b2 = S16 (a1 - a2), b4 = S16 (a1 + a2), b6 = S16 (a0 - a3);
  a0 = S16 (-p[3] + p[5] - p[7] - (p[7] >> 1));
  a1 = S16 (p[1] + p[7] - p[3] - (p[3] >> 1));
  a2 = S16 (-p[1] + p[7] + p[5] + (p[5] >> 1));
  a3 = S16 (p[3] + p[5] + p[1] + (p[1] >> 1));
  const int b1 = S16 (a0 + (a3 >> 2)), b3 = S16 (a1 + (a2 >> 2)), b5 = S16 (a2 - (a1 >> 2)), b7 = S16 (a3 - (a0 >> 2));
  o[0] = S16 (b0 + b7); o[1] = S16 (b2 - b5); o[2] = S16 (b4 + b3); o[3] = S16 (b6 + b1);
  o[4] = S16 (b6 - b1); o[5] = S16 (b4 - b3); o[6] = S16 (b2 + b5); o[7] = S16 (b0 - b7);
}

// ---- luma / chroma interpolation on a shared-memory window ----------------------------------------------
// McLuma_c family, codec/common/src/mc.cpp:142-346.  w points at the sample co-located with the
// output pixel inside a window of row stride ws that holds 2 extra samples left/above and 4
// right/below.
__device__ __forceinline__ int tap6 (int a, int b, int c, int d, int e, int f) { return a + f - 5 * (b + e) + 20 * (c + d); }
__device__ __forceinline__ int mc_h (const uint8_t* w) { return clip255 ((tap6 (w[-2], w[-1], w[0], w[1], w[2], w[3]) + 16) >> 5); }
__device__ __forceinline__ int mc_vraw (const uint8_t* w, int ws) { return tap6 (w[-2 * ws], w[-ws], w[0], w[ws], w[2 * ws], w[3 * ws]); }
__device__ __forceinline__ int mc_v (const uint8_t* w, int ws) { return clip255 ((mc_vraw (w, ws) + 16) >> 5); }
__device__ __forceinline__ int mc_c (const uint8_t* w, int ws) {
  // McHorVer22_c: vertical 6-tap first (int16 intermediates), then horizontal, (x + 512) >> 10
  return clip255 ((tap6 (mc_vraw (w - 2, ws), mc_vraw (w - 1, ws), mc_vraw (w, ws), mc_vraw (w + 1, ws), mc_vraw (w + 2, ws),
                         mc_vraw (w + 3, ws)) + 512) >> 10);
}
__device__ __forceinline__ int mc_luma_sample (const uint8_t* w, int ws, int frac /* fx*4+fy */) {
#define AVG2(a, b) (((a) + (b) + 1) >> 1)
  switch (frac) {            // pWelsMcFunc[x][y], mc.cpp:337-343
  case 0: return w[0];
  case 1: return AVG2 (w[0], mc_v (w, ws));
  case 2: return mc_v (w, ws);
  case 3: return AVG2 (w[ws], mc_v (w, ws));
  case 4: return AVG2 (w[0], mc_h (w));
  case 5: return AVG2 (mc_h (w), mc_v (w, ws));
  case 6: return AVG2 (mc_v (w, ws), mc_c (w, ws));
  case 7: return AVG2 (mc_h (w + ws), mc_v (w, ws));
  case 8: return mc_h (w);
  case 9: return AVG2 (mc_h (w), mc_c (w, ws));
  case 10: return mc_c (w, ws);
  case 11: return AVG2 (mc_h (w + ws), mc_c (w, ws));
  case 12: return AVG2 (w[1], mc_h (w));
  case 13: return AVG2 (mc_h (w), mc_v (w + 1, ws));
  case 14: return AVG2 (mc_v (w + 1, ws), mc_c (w, ws));
  default: return AVG2 (mc_h (w + ws), mc_v (w + 1, ws));
  }
#undef AVG2
}

// Four horizontally adjacent luma samples at once.  p points at the sample co-located with the first output pixel in a
// window whose rows start 4-byte aligned (row stride ws, a multiple of 4).  Each of the 6 window rows is fetched as three
// 32-bit words and funnel-shifted into place (18 word loads instead of up to 144 byte loads for a centre position); the
// vertical 6-tap sums of the 9 columns are shared by the four outputs.  Same arithmetic as mc_luma_sample.
__device__ __forceinline__ void mc_luma_quad (const uint8_t* p, int ws, int frac, int out[4]) {
  const int fx = frac >> 2, fy = frac & 3;
  const unsigned addr = (unsigned)__cvta_generic_to_shared (p - 2);
  const unsigned sh = (addr & 3u) * 8u;
  const uint32_t* q = reinterpret_cast<const uint32_t*> (p - 2 - (addr & 3u));
  const int wsw = ws >> 2;
  const bool needC = (fy == 2 && fx != 0) || (fx == 2 && fy != 0);
  const bool needV = fy != 0 && fx != 2;
  const bool needH = fx != 0 && fy != 2;
  int vr[9];
#pragma unroll
  for (int j = 0; j < 9; j++) vr[j] = 0;
  uint32_t r0a = 0, r0b = 0, r0c = 0, r1a = 0, r1b = 0, r1c = 0;      // rows y and y+1, bytes -2..1 | 2..5 | 6
#pragma unroll
  for (int r = 0; r < 6; r++) {
    if (r != 2 && r != 3 && !needC && !needV) continue;
    const uint32_t* rq = q + (r - 2) * wsw;
    const uint32_t w0 = rq[0], w1 = rq[1], w2 = rq[2];
    const uint32_t a = __funnelshift_r (w0, w1, sh), b = __funnelshift_r (w1, w2, sh), c = (w2 >> sh) & 0xffu;
    if (r == 2) { r0a = a; r0b = b; r0c = c; }
    if (r == 3) { r1a = a; r1b = b; r1c = c; }
    if (needC || needV) {
      const int kc = (r == 0 || r == 5) ? 1 : ((r == 1 || r == 4) ? -5 : 20);
#pragma unroll
      for (int j = 0; j < 4; j++) { vr[j] += kc * (int) ((a >> (8 * j)) & 0xffu); vr[4 + j] += kc * (int) ((b >> (8 * j)) & 0xffu); }
      vr[8] += kc * (int)c;
    }
  }
  // bytes of the row the horizontal filter runs on (row y, or y+1 when fy == 3)
  const uint32_t ha = fy == 3 ? r1a : r0a, hb = fy == 3 ? r1b : r0b, hc = fy == 3 ? r1c : r0c;
  int hrow[9];
#pragma unroll
  for (int j = 0; j < 4; j++) { hrow[j] = (int) ((ha >> (8 * j)) & 0xffu); hrow[4 + j] = (int) ((hb >> (8 * j)) & 0xffu); }
  hrow[8] = (int)hc;
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const int G = (int) (((i < 2 ? r0a >> (8 * (i + 2)) : r0b >> (8 * (i - 2)))) & 0xffu);
    int H = 0, V = 0, C = 0;
    if (needH) H = clip255 ((tap6 (hrow[i], hrow[i + 1], hrow[i + 2], hrow[i + 3], hrow[i + 4], hrow[i + 5]) + 16) >> 5);
    if (needV) V = clip255 (((fx == 3 ? vr[i + 3] : vr[i + 2]) + 16) >> 5);
    if (needC) C = clip255 ((tap6 (vr[i], vr[i + 1], vr[i + 2], vr[i + 3], vr[i + 4], vr[i + 5]) + 512) >> 10);
    int o;
    if (fx == 0) {
      const int Gb = (int) (((i < 2 ? r1a >> (8 * (i + 2)) : r1b >> (8 * (i - 2)))) & 0xffu);
      o = fy == 0 ? G : (fy == 2 ? V : (((fy == 1 ? G : Gb) + V + 1) >> 1));
    } else if (fx == 2) {
      o = fy == 0 ? H : (fy == 2 ? C : ((H + C + 1) >> 1));
    } else {
      const int Gn = fx == 1 ? G : (int) (((i + 1 < 2 ? r0a >> (8 * (i + 3)) : r0b >> (8 * (i - 1)))) & 0xffu);   // fx == 3: the pixel to the right
      o = fy == 0 ? ((Gn + H + 1) >> 1) : (fy == 2 ? ((V + C + 1) >> 1) : ((H + V + 1) >> 1));
    }
    out[i] = o;
  }
}

// partition geometry of the 4x4 block (bx,by) (block units) inside an inter MB — GetInterPred's
// partition walk, codec/decoder/core/src/rec_mb.cpp:344-545
__device__ __forceinline__ void part_geom (int mb_type, int sub, int bx, int by, int& ox, int& oy, int& bw, int& bh) {
  const int qx = (bx >> 1) << 3, qy = (by >> 1) << 3;
  switch (mb_type) {
  case LH264_MB_P16x8: ox = 0; oy = qy; bw = 16; bh = 8; break;
  case LH264_MB_P8x16: ox = qx; oy = 0; bw = 8; bh = 16; break;
  case LH264_MB_P8x8:
    switch (sub) {
    case LH264_SUB_8x8: ox = qx; oy = qy; bw = 8; bh = 8; break;
    case LH264_SUB_8x4: ox = qx; oy = by << 2; bw = 8; bh = 4; break;
    case LH264_SUB_4x8: ox = bx << 2; oy = qy; bw = 4; bh = 8; break;
    default: ox = bx << 2; oy = by << 2; bw = 4; bh = 4; break;
    }
    break;
  default: ox = 0; oy = 0; bw = 16; bh = 16; break;
  }
}

// WeightPrediction sample op, rec_mb.cpp:291-300
__device__ __forceinline__ int wp_sample (int v, int w, int o, int ld) {
  return clip255 (ld >= 1 ? (((v * w + (1 << (ld - 1))) >> ld) + o) : (v * w + o));
}

} // namespace lh264
