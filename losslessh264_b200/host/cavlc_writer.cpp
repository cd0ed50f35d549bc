// cavlc_writer.cpp — see cavlc_writer.h.
#include "cavlc_writer.h"
#include "cavlc.h"
#include <string.h>

namespace lh264host {

// code tables of cavlc.cpp
void cavlc_tables (const uint8_t (**tok_len)[4][17], const uint8_t (**tok_code)[4][17], const uint8_t (**tokc_len)[5], const uint8_t (**tokc_code)[5],
                   const uint8_t (**tz_len)[16], const uint8_t (**tz_code)[16], const uint8_t (**tzc_len)[4], const uint8_t (**tzc_code)[4],
                   const uint8_t (**run_len)[15], const uint8_t (**run_code)[15]);

void cavlc_write_residual_block (BitWriter& bw, int nC, int max_coeff, const int16_t* c) {
  const uint8_t (*tok_len)[4][17], (*tok_code)[4][17], (*tokc_len)[5], (*tokc_code)[5], (*tz_len)[16], (*tz_code)[16], (*tzc_len)[4], (*tzc_code)[4], (*run_len)[15], (*run_code)[15];
  cavlc_tables (&tok_len, &tok_code, &tokc_len, &tokc_code, &tz_len, &tz_code, &tzc_len, &tzc_code, &run_len, &run_code);
  // non-zero levels from the highest frequency down, runs of zeros before each
  int level[16], run[16], total = 0, last = -1;
  for (int i = max_coeff - 1; i >= 0; i--) if (c[i]) { last = i; break; }
  int zeros = 0;
  for (int i = last; i >= 0; i--) {
    if (c[i]) { level[total] = c[i]; run[total] = 0; total++; }
    else { run[total - 1]++; zeros++; }
  }
  int t1 = 0;
  while (t1 < total && t1 < 3 && (level[t1] == 1 || level[t1] == -1)) t1++;
  // coeff_token
  if (nC < 0) bw.put (tokc_code[t1][total], tokc_len[t1][total]);
  else if (nC >= 8) bw.put (total == 0 ? 3u : (uint32_t) (((total - 1) << 2) | t1), 6);
  else { const int t = nC < 2 ? 0 : (nC < 4 ? 1 : 2); bw.put (tok_code[t][t1][total], tok_len[t][t1][total]); }
  if (total == 0) return;
  for (int i = 0; i < t1; i++) bw.bit (level[i] < 0);
  int suffix_len = (total > 10 && t1 < 3) ? 1 : 0;
  for (int i = t1; i < total; i++) {
    const int lv = level[i];
    int code = lv > 0 ? 2 * lv - 2 : -2 * lv - 1;
    if (i == t1 && t1 < 3) code -= 2;
    if (suffix_len == 0) {
      if (code < 14) { bw.put (1, code + 1); }
      else if (code < 30) { bw.put (1, 15); bw.put ((uint32_t) (code - 14), 4); }
      else {
        const int rest = code - 30;
        int prefix;
        if (rest < 4096) { bw.put (1, 16); bw.put ((uint32_t)rest, 12); }
        else {
          prefix = 16;
          while (rest - ((1 << (prefix - 3)) - 4096) >= (1 << (prefix - 3))) prefix++;
          bw.put (1, prefix + 1); bw.put ((uint32_t) (rest - ((1 << (prefix - 3)) - 4096)), prefix - 3);
        }
      }
    } else {
      if (code < (15 << suffix_len)) { bw.put (1, (code >> suffix_len) + 1); bw.put ((uint32_t) (code & ((1 << suffix_len) - 1)), suffix_len); }
      else {
        const int rest = code - (15 << suffix_len);
        if (rest < 4096) { bw.put (1, 16); bw.put ((uint32_t)rest, 12); }
        else {
          int prefix = 16;
          while (rest - ((1 << (prefix - 3)) - 4096) >= (1 << (prefix - 3))) prefix++;
          bw.put (1, prefix + 1); bw.put ((uint32_t) (rest - ((1 << (prefix - 3)) - 4096)), prefix - 3);
        }
      }
    }
    if (suffix_len == 0) suffix_len = 1;
    if ((lv < 0 ? -lv : lv) > (3 << (suffix_len - 1)) && suffix_len < 6) suffix_len++;
  }
  if (total < max_coeff) {
    if (nC < 0) bw.put (tzc_code[total - 1][zeros], tzc_len[total - 1][zeros]);
    else bw.put (tz_code[total - 1][zeros], tz_len[total - 1][zeros]);
  }
  int zeros_left = zeros;
  for (int i = 0; i + 1 < total && zeros_left > 0; i++) {
    const int k = (zeros_left > 7 ? 7 : zeros_left) - 1;
    bw.put (run_code[k][run[i]], run_len[k][run[i]]);
    zeros_left -= run[i];
  }
}

PictureSyntaxWriter::PictureSyntaxWriter (int mb_w, int mb_h) : mb_w_ (mb_w), mb_h_ (mb_h), tc_ ((size_t)mb_w * mb_h * 24, 0), slice_of_ ((size_t)mb_w * mb_h, -1) {}

void PictureSyntaxWriter::begin_slice (const SliceWriteParams& p) { p_ = p; slice_no_++; qp_ = p.slice_qp; }

int PictureSyntaxWriter::nC (int mb, int r, int plane) const {
  const int mbx = mb % mb_w_, mby = mb / mb_w_;
  int nA = -1, nB = -1;
  if (plane == 0) {
    const int bx = r & 3, by = r >> 2;
    if (bx > 0) nA = tc_[(size_t)mb * 24 + r - 1];
    else if (mbx > 0 && avail (mb - 1)) nA = tc_[(size_t) (mb - 1) * 24 + r + 3];
    if (by > 0) nB = tc_[(size_t)mb * 24 + r - 4];
    else if (mby > 0 && avail (mb - mb_w_)) nB = tc_[(size_t) (mb - mb_w_) * 24 + r + 12];
  } else {
    const int base = 16 + 4 * (plane - 1), bx = r & 1, by = r >> 1;
    if (bx > 0) nA = tc_[(size_t)mb * 24 + base + r - 1];
    else if (mbx > 0 && avail (mb - 1)) nA = tc_[(size_t) (mb - 1) * 24 + base + r + 1];
    if (by > 0) nB = tc_[(size_t)mb * 24 + base + r - 2];
    else if (mby > 0 && avail (mb - mb_w_)) nB = tc_[(size_t) (mb - mb_w_) * 24 + base + r + 2];
  }
  if (nA >= 0 && nB >= 0) return (nA + nB + 1) >> 1;
  return nA >= 0 ? nA : (nB >= 0 ? nB : 0);
}

bool PictureSyntaxWriter::write_mb (BitWriter& bw, const lh264_model_mb& r, const int16_t* pool) {
  const int mb = (int)r.mb_addr;
  if (mb < 0 || mb >= mb_w_ * mb_h_) return false;
  uint8_t* tc = &tc_[(size_t)mb * 24];
  const bool p_slice = p_.slice_type == 0;
  if (p_slice && (r.flags & LH264_MM_EMIT_SKIP_RUN)) bw.ue (r.skip_run);      // mb_skip_run
  slice_of_[mb] = slice_no_;
  if (!(r.flags & LH264_MM_WRITE_BLOCK)) { memset (tc, 0, 24); return true; }  // P_Skip
  const int tcode = r.type_code;
  if (tcode == 8 || tcode > 8) return false;                                  // I_PCM / unknown: not expressible from model records
  // dense levels
  int16_t lev[384];
  memset (lev, 0, sizeof (lev));
  {
    const int16_t* src = pool + (size_t)r.coef_off * 16;
    for (int b = 0; b < 24; b++) if (r.coded_mask & (1u << b)) { memcpy (lev + 16 * b, src, 32); src += 16; }
  }
  const int cbp = r.cbp, cbpL = cbp & 15, cbpC = cbp >> 4;
  const bool t8 = r.flags & LH264_MM_T8;
  const bool intra = tcode <= 2, i16 = tcode == 1;
  // ---- mb_type and prediction syntax (7.3.5, 7.3.5.1, 7.3.5.2) ----
  if (intra) {
    uint32_t t;
    if (i16) { const int raw = r.i16_mode > 3 ? 2 : r.i16_mode; t = 1u + raw + 4 * cbpC + (cbpL == 15 ? 12 : 0); }
    else t = 0;
    bw.ue (p_slice ? t + 5 : t);
    if (!i16) {
      if (p_.transform_8x8_mode) bw.bit (tcode == 2);
      const int n = tcode == 2 ? 4 : 16;
      for (int i = 0; i < n; i++) {
        const int mode = r.pred_mode[i] & 15, pred = r.pred_mode[i] >> 4;
        if (mode == pred) bw.bit (1);
        else { bw.bit (0); bw.put ((uint32_t) (mode < pred ? mode : mode - 1), 3); }
      }
    }
    bw.ue (r.chroma_mode > 3 ? 0u : r.chroma_mode);                            // the record holds the availability-resolved predictor: 4..6 are DC
  } else {
    bw.ue ((uint32_t) (tcode - 3));
    const int nref = p_.num_ref_idx_l0;
    auto mvd = [&] (int b) { bw.se (r.mvd[b][0]); bw.se (r.mvd[b][1]); };
    if (tcode == 6 || tcode == 7) {
      for (int q = 0; q < 4; q++) { const int s = r.sub_mb_type[q]; bw.ue (s == 1 ? 0u : (s == 2 ? 1u : (s == 4 ? 2u : 3u))); }
      if (tcode == 6 && nref > 1) for (int q = 0; q < 4; q++) bw.te (nref - 1, (uint32_t)r.ref_idx[q]);
      for (int q = 0; q < 4; q++) {
        const int s = r.sub_mb_type[q], qx = (q & 1) * 2, qy = (q >> 1) * 2;
        const int cnt = s == 1 ? 1 : (s == 8 ? 4 : 2);
        for (int j = 0; j < cnt; j++) {
          int bx = qx, by = qy;
          if (s == 2) by += j; else if (s == 4) bx += j; else if (s == 8) { bx += j & 1; by += j >> 1; }
          mvd (by * 4 + bx);
        }
      }
    } else if (tcode == 4 || tcode == 5) {
      if (nref > 1) for (int i = 0; i < 2; i++) bw.te (nref - 1, (uint32_t)r.ref_idx[i]);
      for (int i = 0; i < 2; i++) mvd (tcode == 4 ? i * 8 : i * 2);
    } else {
      if (nref > 1) bw.te (nref - 1, (uint32_t)r.ref_idx[0]);
      mvd (0);
    }
  }
  if (!i16) {
    const uint8_t* table = intra ? kCbpIntra : kCbpInter;
    int code = 0;
    while (code < 48 && table[code] != cbp) code++;
    if (code >= 48) return false;
    bw.ue ((uint32_t)code);
    if (!intra && (r.flags & LH264_MM_T8_CODED)) bw.bit (t8);
  }
  if (cbp == 0 && !i16) { memset (tc, 0, 24); return true; }
  // ---- mb_qp_delta ----
  {
    int d = (int)r.luma_qp - qp_;
    if (d > 25) d -= 52; else if (d < -26) d += 52;
    bw.se (d);
    qp_ = r.luma_qp;
  }
  // ---- residual (7.3.5.3) ----
  int16_t blk[16];
  if (i16) {
    for (int i = 0; i < 16; i++) blk[i] = lev[16 * kRtoZ[kZigzag4x4[i]]];
    cavlc_write_residual_block (bw, nC (mb, 0, 0), 16, blk);
  }
  for (int q = 0; q < 4; q++) {
    if (!(cbpL & (1 << q))) { for (int k = 0; k < 4; k++) tc[kZtoR[4 * q + k]] = 0; continue; }
    for (int k = 0; k < 4; k++) {
      const int z = 4 * q + k, rr = kZtoR[z];
      int n = 0, maxc = 16;
      if (t8) { for (int i = 0; i < 16; i++) blk[i] = lev[64 * q + kZigzag8x8[4 * i + k]]; }
      else if (i16) { maxc = 15; for (int i = 0; i < 15; i++) blk[i] = lev[16 * z + kZigzag4x4[i + 1]]; }
      else { for (int i = 0; i < 16; i++) blk[i] = lev[16 * z + kZigzag4x4[i]]; }
      for (int i = 0; i < maxc; i++) n += blk[i] != 0;
      cavlc_write_residual_block (bw, nC (mb, rr, 0), maxc, blk);
      tc[rr] = (uint8_t)n;
    }
  }
  if (cbpC) {
    for (int p = 0; p < 2; p++) {
      for (int b = 0; b < 4; b++) blk[b] = lev[16 * (16 + 4 * p + b)];
      cavlc_write_residual_block (bw, -1, 4, blk);
    }
  }
  for (int p = 0; p < 2; p++) for (int b = 0; b < 4; b++) {
      if (cbpC & 2) {
        int n = 0;
        for (int i = 0; i < 15; i++) { blk[i] = lev[16 * (16 + 4 * p + b) + kZigzag4x4[i + 1]]; n += blk[i] != 0; }
        cavlc_write_residual_block (bw, nC (mb, b, 1 + p), 15, blk);
        tc[16 + 4 * p + b] = (uint8_t)n;
      } else tc[16 + 4 * p + b] = 0;
    }
  return true;
}

} // namespace lh264host
