// compact_records.h — the mapping between the kernel-side records (lh264_mb_rec, lh264_model_mb: include/lh264_b200.h) and
// their compact transfer form (lh264_cmb, lh264_cmm).  Expansion runs on the device (k_expand_mb / k_expand_mm) and, for the
// host's self-check, on the CPU; compaction runs on the host (runtime.cpp) when a picture is handed over.  expand (compact (r))
// == r byte for byte for every record the host parser produces (tests/test_compact_records.py checks it on the whole corpus).
#pragma once
#include <stdint.h>
#include <string.h>
#include "../../include/lh264_b200.h"

#if defined(__CUDACC__)
#define LH264_HD __host__ __device__ __forceinline__
#else
#define LH264_HD static inline
#endif

// ---- device / host: compact -> full -------------------------------------------------------------------------------------------
LH264_HD void lh264_expand_mb (const lh264_cmb& c, const uint8_t* extra, lh264_mb_rec& m) {
  uint32_t* w = reinterpret_cast<uint32_t*> (&m);
  for (int i = 0; i < 32; i++) w[i] = 0;
  m.mb_type = c.mb_type; m.flags = c.flags; m.cbp = c.cbp; m.qp_y = c.qp_y;
  m.qp_c[0] = c.qp_c[0]; m.qp_c[1] = c.qp_c[1]; m.i16_mode = c.i16_mode; m.chroma_mode = c.chroma_mode;
  m.slice_idx = c.slice_idx; m.nxn_avail = c.nxn_avail;
  m.nz_mask = c.nz_mask; m.coded_mask = c.coded_mask; m.coef_off = c.coef_off;
  for (int i = 0; i < 4; i++) m.ref_idx[i] = c.ref_idx[i];
  if (c.kind == 0) {
    const uint32_t v = (uint32_t) (uint16_t)c.u.mv[0] | ((uint32_t) (uint16_t)c.u.mv[1] << 16);
    uint32_t* mv = reinterpret_cast<uint32_t*> (&m.mv[0][0]);
    for (int i = 0; i < 16; i++) mv[i] = v;
    return;
  }
  const uint8_t* x = extra + (size_t)c.u.extra_off * 16;
  if (c.kind == 1) { for (int i = 0; i < 16; i++) m.intra_mode[i] = x[i]; return; }
  const uint32_t* xv = reinterpret_cast<const uint32_t*> (x);
  uint32_t* mv = reinterpret_cast<uint32_t*> (&m.mv[0][0]);
  for (int i = 0; i < 16; i++) mv[i] = xv[i];
  if (c.kind == 3) for (int i = 0; i < 4; i++) m.sub_mb_type[i] = x[64 + i];
}

LH264_HD void lh264_expand_mm (const lh264_cmm& c, const uint8_t* extra, lh264_model_mb& m) {
  uint32_t* w = reinterpret_cast<uint32_t*> (&m);
  for (int i = 0; i < 32; i++) w[i] = 0;
  m.mb_addr = c.mb_addr; m.flags = c.flags; m.skip_run = c.skip_run; m.mb_in_slice = c.mb_in_slice;
  m.type_code = c.type_code; m.cbp = c.cbp; m.luma_qp = c.luma_qp; m.chroma_mode = c.chroma_mode; m.i16_mode = c.i16_mode; m.qp_delta = c.qp_delta;
  for (int i = 0; i < 4; i++) { m.sub_mb_type[i] = c.sub_mb_type[i]; m.ref_idx[i] = c.ref_idx[i]; }
  m.coded_mask = c.coded_mask; m.coef_off = c.coef_off;
  if (c.kind == 0) {
    // the second partition's difference sits at the raster index of its first 4x4 block: 16x8 (type_code 4) block 8, 8x16 (5) block 2
    m.mvd[0][0] = c.u.mvd[0][0]; m.mvd[0][1] = c.u.mvd[0][1];
    const int second = c.type_code == 4 ? 8 : 2;
    if (c.type_code == 4 || c.type_code == 5) { m.mvd[second][0] = c.u.mvd[1][0]; m.mvd[second][1] = c.u.mvd[1][1]; }
    return;
  }
  const uint8_t* x = extra + (size_t)c.u.extra_off * 16;
  if (c.kind == 1) { for (int i = 0; i < 16; i++) { m.pred_mode[i] = x[i]; m.pred_avail[i] = x[16 + i]; } return; }
  const uint32_t* xv = reinterpret_cast<const uint32_t*> (x);
  uint32_t* mvd = reinterpret_cast<uint32_t*> (&m.mvd[0][0]);
  for (int i = 0; i < 16; i++) mvd[i] = xv[i];
}

// ---- host: full -> compact ------------------------------------------------------------------------------------------------------
#if !defined(__CUDA_ARCH__)
// Writes the record's extra units (if any) at extra + 16 * *n_extra and advances *n_extra.  `extra` must have room for 5 units.
static inline void lh264_compact_mb (const lh264_mb_rec& m, lh264_cmb& c, uint8_t* extra, uint32_t* n_extra) {
  memset (&c, 0, sizeof (c));
  c.mb_type = m.mb_type; c.flags = m.flags; c.cbp = m.cbp; c.qp_y = m.qp_y;
  c.qp_c[0] = m.qp_c[0]; c.qp_c[1] = m.qp_c[1]; c.i16_mode = m.i16_mode; c.chroma_mode = m.chroma_mode;
  c.slice_idx = m.slice_idx; c.nxn_avail = m.nxn_avail;
  c.nz_mask = m.nz_mask; c.coded_mask = m.coded_mask; c.coef_off = m.coef_off;
  memcpy (c.ref_idx, m.ref_idx, 4);
  uint8_t* x = extra + (size_t)*n_extra * 16;
  if (m.mb_type == LH264_MB_I4x4 || m.mb_type == LH264_MB_I8x8) {
    c.kind = 1; c.u.extra_off = *n_extra;
    memcpy (x, m.intra_mode, 16);
    *n_extra += 1;
  } else if (m.mb_type == LH264_MB_P16x8 || m.mb_type == LH264_MB_P8x16 || m.mb_type == LH264_MB_P8x8) {
    const bool sub = m.mb_type == LH264_MB_P8x8;
    c.kind = sub ? 3 : 2; c.u.extra_off = *n_extra;
    memcpy (x, m.mv, 64);
    if (sub) { memset (x + 64, 0, 16); memcpy (x + 64, m.sub_mb_type, 4); }
    *n_extra += sub ? 5 : 4;
  } else {
    c.kind = 0; c.u.mv[0] = m.mv[0][0]; c.u.mv[1] = m.mv[0][1];
  }
}
static inline void lh264_compact_mm (const lh264_model_mb& m, lh264_cmm& c, uint8_t* extra, uint32_t* n_extra) {
  memset (&c, 0, sizeof (c));
  c.mb_addr = m.mb_addr; c.flags = m.flags; c.skip_run = m.skip_run; c.mb_in_slice = m.mb_in_slice;
  c.type_code = m.type_code; c.cbp = m.cbp; c.luma_qp = m.luma_qp; c.chroma_mode = m.chroma_mode; c.i16_mode = m.i16_mode; c.qp_delta = m.qp_delta;
  memcpy (c.sub_mb_type, m.sub_mb_type, 4); memcpy (c.ref_idx, m.ref_idx, 4);
  c.coded_mask = m.coded_mask; c.coef_off = m.coef_off;
  uint8_t* x = extra + (size_t)*n_extra * 16;
  if (m.type_code == 0 || m.type_code == 2) {        // I4x4 / I8x8: signalled and predicted modes, sample availability
    c.kind = 1; c.u.extra_off = *n_extra;
    memcpy (x, m.pred_mode, 16); memcpy (x + 16, m.pred_avail, 16);
    *n_extra += 2;
  } else if (m.type_code == 6 || m.type_code == 7) { // 8x8 types: a difference per sub-partition
    c.kind = 2; c.u.extra_off = *n_extra;
    memcpy (x, m.mvd, 64);
    *n_extra += 4;
  } else {
    c.kind = 0;
    c.u.mvd[0][0] = m.mvd[0][0]; c.u.mvd[0][1] = m.mvd[0][1];
    const int second = m.type_code == 4 ? 8 : 2;
    if (m.type_code == 4 || m.type_code == 5) { c.u.mvd[1][0] = m.mvd[second][0]; c.u.mvd[1][1] = m.mvd[second][1]; }
  }
}
#endif
