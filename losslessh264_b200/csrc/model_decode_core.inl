// model_decode_core.inl — the DECOMPRESS direction of the recompression model, one stream = one strictly serial decoder.
//
// Replaces, per stream, the reference's round-trip path: reading the arithmetic-coded side streams back through the same
// adaptive priors (decode4x4 / readMv / scanInt / scanUEGkInt / scanBits*, codec/decoder/core/inc/compression_stream.h
// :601-689, src/decode_slice.cpp:2476-2956; bool decoder inc/bitreader.h:77-136), rebuilding every macroblock's syntax
// elements, and re-serialising them as the original CAVLC slice_data() (EncoderState::writeMbSynCavlc, decode_slice.cpp
// :1442-1835 -> encoder/core/src/svc_set_mb_syn_cavlc.cpp).  Every binary decision depends on all earlier ones of the stream
// (adaptive probabilities AND the symbol-dependent choice of the next prior), so the only parallel axis is the stream: the
// CUDA kernel (model_decode.cuh) runs one warp per stream, lane 0 walking this code.
//
// This file is plain C++ behind three macros so that the tests can compile the very same code for the CPU as a checker:
//   MD_FN     function qualifier   (`__device__` in the CUDA build, `static inline` on the CPU)
//   MD_FN_NOINLINE  the same, never inlined (keeps the kernel small enough for the instruction cache)
//   MD_CLZ    count-leading-zeros of a non-zero 32-bit value;  MD_DIV(n, d) exact unsigned n / d for n < 2^18, d < 2^11
//   MD_READER(c, tag) / MD_TEST(c)  lvalues of a tag's bool-decoder state and of TEST_PROB (shared memory in the CUDA build)
//   tables    host/cavlc_tables.h is included by the includer with LH264_CAVLC_TABLE set accordingly
// The includer is in namespace scope; model_layout.h (tags, prior layout, CtxRec) must already be included.

namespace lh264md {
using namespace lh264m;
using namespace lh264host;              // CAVLC tables

// ---- DynProb word (compression_stream.h:87-115); identical packing to the compress kernels ---------------------------
MD_FN uint32_t md_prob (uint32_t w) { return (w >> 24) ^ 0x80u; }
MD_FN uint32_t md_adapt (uint32_t w, uint32_t bit) {
  uint32_t c0 = w & 0x3ffu, c1 = (w >> 10) & 0x3ffu;
  if (bit) c1++; else c0++;
  const uint32_t prob = MD_DIV (256u * (c0 + 1u), c0 + c1 + 2u);
  if (c0 + c1 > 512u) { c0 = (c0 + 1u) >> 1; c1 = (c1 + 1u) >> 1; }
  return c0 | (c1 << 10) | ((prob ^ 0x80u) << 24);
}

// ---- VP8 bool decoder (bitreader.h:77-136, bitreader.cpp:45-110), 32-bit window; bytes past the end read as zero ------
struct MdReader { const uint8_t* buf; uint32_t len, pos, value, range; int32_t count; };
#ifdef MD_SHARED_STATE
__shared__ MdReader s_rd[CT_COUNT];
__shared__ uint32_t s_test;
#endif
MD_FN void md_reader_init (MdReader& r, const uint8_t* buf, uint32_t len) { r.buf = buf; r.len = len; r.pos = 0; r.value = 0; r.range = 255; r.count = -8; }
MD_FN uint32_t md_read (MdReader& r, uint32_t prob) {
  // the state is worked on in registers and stored once: `buf` is a byte pointer, so every store through `r` inside the
  // refill loop would otherwise have to be assumed to alias it
  uint32_t value = r.value, range = r.range;
  int32_t count = r.count;
  const uint32_t split = 1u + (((range - 1u) * prob) >> 8);
  if (count < 0) {
    uint32_t pos = r.pos;
    const uint32_t len = r.len;
    const uint8_t* buf = r.buf;
    int shift = 32 - 8 - (count + 8);                 // 17..24: three bytes, four when the window is empty
    for (int k = 0; k < 4; k++) {
      if (shift < 0) break;
      const uint32_t b = pos < len ? buf[pos] : 0u;
      pos++;
      value |= b << shift;
      shift -= 8; count += 8;
    }
    r.pos = pos;
  }
  const uint32_t bigsplit = split << 24;
  uint32_t bit = 0;
  if (value >= bigsplit) { range -= split; value -= bigsplit; bit = 1; } else range = split;
  const int shift = MD_CLZ (range) - 24;
  r.range = range << shift; r.value = value << shift; r.count = count - shift;
  return bit;
}

// ---- per-stream state ------------------------------------------------------------------------------------------
#define MD_NB_STRIDE 48
struct MdStream {
  uint32_t* priors;               // pruned prior table (model_layout.h), zero = fresh
  CtxRec* frame[2];               // FreqImage ring (decoded_macroblock.h:106-191)
  uint8_t* ipm;                   // [mb][16] Intra4x4/8x8PredMode per raster 4x4 block (2 for every other MB type)
  int32_t* mb_slice;              // [mb] number of the slice that last coded the MB (availability, 6.4.8)
  uint8_t* mb_intra;              // [mb] 1 = intra MB (constrained_intra_pred)
  uint8_t* tcoef;                 // [mb][MD_NB_STRIDE] what the entropy coder's context selection reads from a neighbouring macroblock.
                                  // CAVLC slices: total_coeff per 4x4 block, 16 luma (raster), 4 Cb, 4 Cr (nC, 9.2.1); CABAC: see cb_mb
  int32_t width, height, frame_len, cur, lastFrameId, max_mbs, slice_counter, error;
  MdReader rd[CT_COUNT];
  uint32_t present;               // compact tags that have a side stream
  uint32_t pad_;
};
// error codes
enum { MD_ERR_PCM = 21, MD_ERR_OUT_FULL = 22, MD_ERR_SYNTAX = 23, MD_ERR_DIMS = 24, MD_ERR_CONTEXT = 25 };

// Decoder context.  Priors are read-modified-written in place in the table: their loads hit L1 (99.6 % in the ncu capture,
// profiles/r1_decode_ncu.md) and account for ~8 % of the stall samples.  Two software caches of the table in shared memory were
// measured and rejected (a one-record write-back buffer, and a 256-record direct-mapped cache with a 95 % hit rate): the
// decoder is bound by the length of its dependent instruction chain (~170 instructions per decision), not by these loads.
// TEST_PROB (shared by all tags, interleaved with everything) lives next to the bool-decoder states.
struct MdCtx { uint32_t* priors; MdReader* rd; uint32_t test; };
MD_FN void md_ctx_init (MdCtx& c, uint32_t* priors, MdReader* rd) { c.priors = priors; c.rd = rd; }
MD_FN uint32_t* md_prior (MdCtx& c, uint32_t prior) { return &c.priors[prior]; }

MD_FN_NOINLINE uint32_t md_bit (MdCtx& c, int tag, uint32_t prior) {     // ArithmeticCodedInput::scanBit + DynProb update
  if (prior == P_TEST) {
    const uint32_t t = MD_TEST (c);
    const uint32_t bit = md_read (MD_READER (c, tag), md_prob (t));
    MD_TEST (c) = md_adapt (t, bit);
    return bit;
  }
  uint32_t* w = md_prior (c, prior);
  const uint32_t v = *w;
  const uint32_t bit = md_read (MD_READER (c, tag), md_prob (v));
  *w = md_adapt (v, bit);
  return bit;
}
// scanBits(nBits, Branch<n>): compression_stream.h pre-order tree, mirror of emitBits :477-487
MD_FN_NOINLINE uint32_t md_branch (MdCtx& c, int tag, int nbits, uint32_t tree) {
  uint32_t data = 0;
  for (int n = nbits; n >= 1; n--) {
    const uint32_t bit = md_bit (c, tag, tree);
    data = (data << 1) | bit;
    tree += bit ? (1u << (n - 1)) : 1u;
  }
  return data;
}
MD_FN uint32_t md_raw (MdCtx& c, int tag, int nbits) {          // scanBits(nBits) through TEST_PROB
  uint32_t data = 0;
  for (int i = 0; i < nbits; i++) data = (data << 1) | md_bit (c, tag, P_TEST);
  return data;
}
MD_FN uint32_t md_zero_to_pow2 (MdCtx& c, int tag, int nbits, uint32_t priors, uint32_t preferred) {   // scanBitsZeroToPow2Inclusive
  if (!md_bit (c, tag, priors)) return preferred;
  const uint32_t d = md_branch (c, tag, nbits, priors + 1);
  return d >= preferred ? d + 1 : d;
}
MD_FN_NOINLINE int md_unary (MdCtx& c, int tag, uint32_t base, int n, int early) {                              // scanUnary
  int i = 0;
  for (;;) {
    if (!md_bit (c, tag, base + (uint32_t) (i < n - 1 ? i : n - 1))) return i;
    i++;
    if (i == early || i > 64) return i;
  }
}
// scanInt: mirror of CompressionStream::emitInt :523-572 (mantissa size 4)
MD_FN_NOINLINE int md_int (MdCtx& c, uint32_t expo, int n_exp, uint32_t mant, int order, uint32_t zero, uint32_t sign, int tagExp, int tagMant, int tagZero, int tagSign) {
  if (zero != NO_PRIOR && md_bit (c, tagZero, zero)) return 0;
  uint32_t positive = 1;
  if (sign != NO_PRIOR) positive = md_bit (c, tagSign, sign);
  const int log2 = md_unary (c, tagExp, expo, n_exp, -1);
  if (log2 > 30) return 0;
  int lo = 0, hi = 4;
  uint32_t data_high = 1, low = 0;
  for (int k = 0; k < log2 + order; k++) {
    uint32_t bit;
    if (hi > lo) {
      const int mid = (hi + lo) >> 1;
      bit = md_bit (c, tagMant, mant + (uint32_t)mid);
      if (bit) lo = mid + 1; else hi = mid;
    } else bit = md_bit (c, tagMant, P_TEST);
    if (k < log2) data_high = (data_high << 1) | bit; else low = (low << 1) | bit;
  }
  const int data = (int) ((((data_high - 1u) << order) | low) + 1u);
  return positive ? data : -data;
}
// scanUEGkInt: mirror of emitUEGkInt :575-591
MD_FN_NOINLINE int md_uegk (MdCtx& c, int threshold, uint32_t p, int n_exp, int order, int tagExp, int tagMant, int tagZero, int tagSign) {
  if (md_bit (c, tagZero, p + U_ZERO)) return 0;
  const uint32_t neg = md_bit (c, tagSign, p + U_SIGN);
  const int u = md_unary (c, tagMant, p + U_FIRST, 4, threshold);
  int mag = u + 1;
  if (u >= threshold)
    mag = 1 + threshold + md_int (c, p + U_EXP, n_exp, p + U_EXP + n_exp, order, p + U_EXP + n_exp + 4, NO_PRIOR, tagExp, tagMant, tagZero, tagSign);
  return neg ? -mag : mag;
}

MD_TABLE uint8_t md_zz16[16] = { 0, 1, 4, 8, 5, 2, 3, 6, 9, 12, 13, 10, 7, 11, 14, 15 };     // decode_slice.cpp:2035-2040
MD_TABLE uint8_t md_zz64[64] = {                                                              // decode_slice.cpp:2041-2050
  0, 1, 5, 6, 14, 15, 27, 28, 2, 4, 7, 13, 16, 26, 29, 42, 3, 8, 12, 17, 25, 30, 41, 43, 9, 11, 18, 24, 31, 40, 44, 53,
  10, 19, 23, 32, 39, 45, 52, 54, 20, 22, 33, 38, 46, 51, 55, 60, 21, 34, 37, 47, 50, 56, 59, 61, 35, 36, 48, 49, 57, 58, 62, 63 };
MD_TABLE uint8_t md_scan4[16] = { 0, 1, 4, 5, 2, 3, 6, 7, 8, 9, 12, 13, 10, 11, 14, 15 };    // g_kuiScan4 (z order <-> raster)
// H.264 scans (Figure 8-8 frame scan; 8x8 zig-zag as read by CAVLC: 4 interleaved 4x4 blocks)
MD_TABLE uint8_t md_zigzag4[16] = { 0, 1, 4, 8, 5, 2, 3, 6, 9, 12, 13, 10, 7, 11, 14, 15 };
MD_TABLE uint8_t md_zigzag8[64] = {
  0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6, 7, 14, 21, 28,
  35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63 };
MD_TABLE uint8_t md_i16dc_scan[16] = { 0, 16, 32, 128, 48, 64, 80, 96, 144, 160, 176, 192, 112, 208, 224, 240 };   // 16 * z(raster zigzag4[i]): Intra16x16 DC levels sit in slot 0 of block z
MD_TABLE uint8_t md_cdc_scan[4] = { 0, 16, 32, 48 };                                                                  // chroma DC k -> slot 0 of block k
MD_TABLE uint8_t md_cbp_intra[48] = { 47, 31, 15, 0, 23, 27, 29, 30, 7, 11, 13, 14, 39, 43, 45, 46, 16, 3, 5, 10, 12, 19, 21, 26, 28, 35, 37, 42, 44, 1, 2, 4,
                                      8, 17, 18, 20, 24, 6, 9, 22, 25, 32, 33, 34, 36, 40, 38, 41 };   // Table 9-4
MD_TABLE uint8_t md_cbp_inter[48] = { 0, 16, 1, 2, 4, 8, 32, 3, 5, 10, 12, 15, 47, 7, 11, 13, 14, 6, 9, 31, 35, 37, 42, 44, 33, 34, 36, 40, 39, 43, 45, 46,
                                      17, 18, 20, 24, 19, 21, 26, 28, 23, 27, 29, 30, 22, 25, 38, 41 };

// decode4x4<n>: mirror of encode4x4 decode_slice.cpp:2060-2094 (see gen_block in model_kernels.cuh for the prior indexing)
// cnt = non-zero counts of the block's 16-coefficient chunks (DC slot included), kept current as levels are written
MD_FN_NOINLINE void md_block (MdCtx& c, int16_t* ac, uint8_t* cnt, int n, bool emit_dc, int color, uint32_t nzp, uint32_t acbase) {
  const uint8_t* zz = n == 16 ? md_zz16 : md_zz64;
  const int tEob = color ? CT_C_EOB : CT_L0_EOB;
  const int nonzeros = md_int (c, nzp + I_EXP, 3, nzp + I_MANT, 0, nzp + I_ZERO, NO_PRIOR, tEob, tEob, tEob, tEob);
  int left = nonzeros, emitted = 0, prev = 0, prev2 = 0;
  for (int i = emit_dc ? 0 : 1; i < n && left > 0; i++) {
    const int tb = color ? CT_C_BITMASK : (i == 0 ? CT_L0_BITMASK : CT_LN_BITMASK);
    const int l4 = left < 4 ? left : 4;
    const int a = prev + 2 < 0 ? 0 : (prev + 2 > 4 ? 4 : prev + 2), b = prev2 + 2 < 0 ? 0 : (prev2 + 2 > 4 ? 4 : prev2 + 2);
    const uint32_t p = acbase + ((uint32_t) (emitted * 5 + l4) * 25u + (uint32_t)a * 5u + (uint32_t)b) * 16u;
    const int v = md_uegk (c, 14, p, 2, 0, tb + 1, tb + 2, tb, tb + 3);
    ac[zz[i]] = (int16_t)v;
    if (v) cnt[zz[i] >> 4]++;
    prev2 = prev; prev = v; emitted++;
    if (v) left--;
  }
}

// Per-stream scratch (shared memory in the CUDA build: a lone lane's stack costs a 128-byte cache line per word)
struct MdScratch {
  lh264_model_mb r;               // the macroblock being decoded
  int16_t lev[384];               // its levels, layout of the coefficient pool (24 chunks of 16)
  int32_t level[16], run[16];     // cw_residual_block
  uint8_t cnt[24];
  // the scan tables of the slice-data writers, staged here (CUDA build: shared memory) so that the per-coefficient look-ups of
  // cw_residual_block / cb_block stay off the global-load path (md_scratch_init)
  uint8_t t_zigzag4[16], t_zigzag8[64], t_i16dc[16], t_cdc[4], t_scan4[16];
  // CABAC slices: arithmetic encoder (9.3.4.2) and the context variables of the slice, (pStateIdx << 1) | valMPS
  uint32_t cb_low, cb_range, cb_outstanding, cb_first;
  uint8_t cb_state[460];
};
MD_FN void md_scratch_init (MdScratch& sc, int lane, int n_lanes) {
  for (int i = lane; i < 64; i += n_lanes) {
    sc.t_zigzag8[i] = md_zigzag8[i];
    if (i < 16) { sc.t_zigzag4[i] = md_zigzag4[i]; sc.t_i16dc[i] = md_i16dc_scan[i]; sc.t_scan4[i] = md_scan4[i]; }
    if (i < 4) sc.t_cdc[i] = md_cdc_scan[i];
  }
}

// ---- bit sink of the re-serialiser ------------------------------------------------------------------------------------
struct MdBits {
  uint8_t* out; uint64_t cap_bits, n_bits; uint32_t acc; int n_acc; int overflow;
  int trim;                       // bits at the end that belong to the caller's trailer (CABAC: the flush ends with the rbsp_stop_one_bit)
};
MD_FN void mb_init (MdBits& b, uint8_t* out, uint64_t cap_bytes) { b.out = out; b.cap_bits = cap_bytes * 8; b.n_bits = 0; b.acc = 0; b.n_acc = 0; b.overflow = 0; b.trim = 0; }
MD_FN void mb_put (MdBits& b, uint32_t v, int n) {              // n <= 24
  if (n <= 0) return;
  b.acc = (b.acc << n) | (v & ((1u << n) - 1u));
  b.n_acc += n;
  while (b.n_acc >= 8) {
    if (b.n_bits + 8 > b.cap_bits) { b.overflow = 1; b.n_acc -= 8; continue; }
    b.out[b.n_bits >> 3] = (uint8_t) (b.acc >> (b.n_acc - 8));
    b.n_bits += 8; b.n_acc -= 8;
  }
}
MD_FN void mb_put_long (MdBits& b, uint32_t v, int n) { if (n > 16) { mb_put (b, v >> 16, n - 16); mb_put (b, v & 0xffffu, 16); } else mb_put (b, v, n); }
MD_FN uint64_t mb_bits (const MdBits& b) { return b.n_bits + (uint64_t)b.n_acc; }
MD_FN void mb_flush (MdBits& b) {                               // left-aligns what is pending into one more byte (not counted in n_bits)
  if (b.n_acc > 0 && b.n_bits + 8 <= b.cap_bits) b.out[b.n_bits >> 3] = (uint8_t) (b.acc << (8 - b.n_acc));
}
MD_FN_NOINLINE void mb_ue (MdBits& b, uint32_t v) {
  const uint32_t x = v + 1u;
  const int len = 31 - MD_CLZ (x);
  mb_put (b, 0, len > 24 ? 24 : len);
  if (len > 24) mb_put (b, 0, len - 24);
  mb_put_long (b, x, len + 1);
}
MD_FN void mb_se (MdBits& b, int v) { mb_ue (b, v > 0 ? (uint32_t) (2 * v - 1) : (uint32_t) (-2 * v)); }
MD_FN void mb_te (MdBits& b, int range, uint32_t v) { if (range > 1) mb_ue (b, v); else mb_put (b, v ^ 1u, 1); }

// residual_block_cavlc() (7.3.5.3.2 / 9.2) in the writing direction, canonical choices.  The block's levels are read in place:
// coefficient i of the scan is base[scan[i * step]] (one pass from the highest frequency down collects levels, runs and
// total_zeros).  Returns total_coeff.
MD_FN_NOINLINE int cw_residual_block (MdBits& bw, MdScratch& sc, int nC, int max_coeff, const int16_t* base, const uint8_t* scan, int step) {
  int32_t* level = sc.level;
  int32_t* run = sc.run;
  int total = 0, zeros = 0;
  for (int i = max_coeff - 1; i >= 0; i--) {
    const int v = base[scan[i * step]];
    if (v) { level[total] = v; run[total] = 0; total++; }
    else if (total) { run[total - 1]++; zeros++; }
  }
  int t1 = 0;
  while (t1 < total && t1 < 3 && (level[t1] == 1 || level[t1] == -1)) t1++;
  if (nC < 0) mb_put (bw, kTokCodeC[t1][total], kTokLenC[t1][total]);
  else if (nC >= 8) mb_put (bw, total == 0 ? 3u : (uint32_t) (((total - 1) << 2) | t1), 6);
  else { const int t = nC < 2 ? 0 : (nC < 4 ? 1 : 2); mb_put (bw, kTokCode[t][t1][total], kTokLen[t][t1][total]); }
  if (total == 0) return 0;
  for (int i = 0; i < t1; i++) mb_put (bw, level[i] < 0, 1);
  int suffix_len = (total > 10 && t1 < 3) ? 1 : 0;
  for (int i = t1; i < total; i++) {
    const int lv = level[i];
    int code = lv > 0 ? 2 * lv - 2 : -2 * lv - 1;
    if (i == t1 && t1 < 3) code -= 2;
    int rest = -1;                                     // >= 0: escape with level_prefix >= 15
    if (suffix_len == 0) {
      if (code < 14) mb_put (bw, 1, code + 1);
      else if (code < 30) { mb_put (bw, 1, 15); mb_put (bw, (uint32_t) (code - 14), 4); }
      else rest = code - 30;
    } else {
      if (code < (15 << suffix_len)) { mb_put (bw, 1, (code >> suffix_len) + 1); mb_put (bw, (uint32_t)code & ((1u << suffix_len) - 1u), suffix_len); }
      else rest = code - (15 << suffix_len);
    }
    if (rest >= 0) {
      if (rest < 4096) { mb_put (bw, 1, 16); mb_put (bw, (uint32_t)rest, 12); }
      else {
        int prefix = 16;
        while (rest - ((1 << (prefix - 3)) - 4096) >= (1 << (prefix - 3))) prefix++;
        mb_put (bw, 1, prefix + 1); mb_put_long (bw, (uint32_t) (rest - ((1 << (prefix - 3)) - 4096)), prefix - 3);
      }
    }
    if (suffix_len == 0) suffix_len = 1;
    if ((lv < 0 ? -lv : lv) > (3 << (suffix_len - 1)) && suffix_len < 6) suffix_len++;
  }
  if (total < max_coeff) {
    if (nC < 0) mb_put (bw, kTzCodeC[total - 1][zeros], kTzLenC[total - 1][zeros]);
    else mb_put (bw, kTzCode[total - 1][zeros], kTzLen[total - 1][zeros]);
  }
  int zeros_left = zeros;
  for (int i = 0; i + 1 < total && zeros_left > 0; i++) {
    const int k = (zeros_left > 7 ? 7 : zeros_left) - 1;
    mb_put (bw, kRunCode[k][run[i]], kRunLen[k][run[i]]);
    zeros_left -= run[i];
  }
  return total;
}

// per-slice state of the syntax writer
struct CwSlice { int slice_type /*0 P, 2 I*/, num_ref_idx_l0, qp, transform_8x8_mode, cabac, last_dqp_nz; };

MD_FN int cw_nC (int nA, int nB) { if (nA >= 0 && nB >= 0) return (nA + nB + 1) >> 1; return nA >= 0 ? nA : (nB >= 0 ? nB : 0); }

// macroblock_layer() (7.3.5) for one record; lev = the MB's 384 levels in the layout of the coefficient pool (24 chunks of 16).
// tc = this MB's total_coeff[24] (written), tcL / tcT = the left / top neighbour's (nullptr when not available, 6.4.8).
// Returns false for what cannot be expressed (I_PCM).
MD_FN_NOINLINE bool cw_mb (MdBits& bw, MdScratch& sc, CwSlice& s, uint8_t* tc, const uint8_t* tcL, const uint8_t* tcT) {
  const lh264_model_mb& r = sc.r;
  const int16_t* lev = sc.lev;
  const bool p_slice = s.slice_type == 0;
  if (p_slice && (r.flags & LH264_MM_EMIT_SKIP_RUN)) mb_ue (bw, r.skip_run);
  if (!(r.flags & LH264_MM_WRITE_BLOCK)) { for (int i = 0; i < 24; i++) tc[i] = 0; return true; }
  const int tcode = r.type_code;
  if (tcode >= 8) return false;
  const int cbp = r.cbp, cbpL = cbp & 15, cbpC = cbp >> 4;
  const bool t8 = (r.flags & LH264_MM_T8) != 0;
  const bool intra = tcode <= 2, i16 = tcode == 1;
  if (intra) {
    uint32_t t = 0;
    if (i16) { const int raw = r.i16_mode > 3 ? 2 : r.i16_mode; t = 1u + (uint32_t)raw + 4u * (uint32_t)cbpC + (cbpL == 15 ? 12u : 0u); }
    mb_ue (bw, p_slice ? t + 5u : t);
    if (!i16) {
      if (s.transform_8x8_mode) mb_put (bw, tcode == 2, 1);
      const int n = tcode == 2 ? 4 : 16;
      for (int i = 0; i < n; i++) {
        const int mode = r.pred_mode[i] & 15, pred = r.pred_mode[i] >> 4;
        if (mode == pred) mb_put (bw, 1, 1);
        else { mb_put (bw, 0, 1); mb_put (bw, (uint32_t) (mode < pred ? mode : mode - 1), 3); }
      }
    }
    mb_ue (bw, r.chroma_mode > 3 ? 0u : r.chroma_mode);      // 4..6 are the availability-resolved DC predictors
  } else {
    mb_ue (bw, (uint32_t) (tcode - 3));
    const int nref = s.num_ref_idx_l0;
    if (tcode == 6 || tcode == 7) {
      for (int q = 0; q < 4; q++) { const int sb = r.sub_mb_type[q]; mb_ue (bw, sb == 1 ? 0u : (sb == 2 ? 1u : (sb == 4 ? 2u : 3u))); }
      if (tcode == 6 && nref > 1) for (int q = 0; q < 4; q++) mb_te (bw, nref - 1, (uint32_t)r.ref_idx[q]);
      for (int q = 0; q < 4; q++) {
        const int sb = r.sub_mb_type[q], qx = (q & 1) * 2, qy = (q >> 1) * 2;
        const int cnt = sb == 1 ? 1 : (sb == 8 ? 4 : 2);
        for (int j = 0; j < cnt; j++) {
          int bx = qx, by = qy;
          if (sb == 2) by += j; else if (sb == 4) bx += j; else if (sb == 8) { bx += j & 1; by += j >> 1; }
          mb_se (bw, r.mvd[by * 4 + bx][0]); mb_se (bw, r.mvd[by * 4 + bx][1]);
        }
      }
    } else if (tcode == 4 || tcode == 5) {
      if (nref > 1) for (int i = 0; i < 2; i++) mb_te (bw, nref - 1, (uint32_t)r.ref_idx[i]);
      for (int i = 0; i < 2; i++) { const int b = tcode == 4 ? i * 8 : i * 2; mb_se (bw, r.mvd[b][0]); mb_se (bw, r.mvd[b][1]); }
    } else {
      if (nref > 1) mb_te (bw, nref - 1, (uint32_t)r.ref_idx[0]);
      mb_se (bw, r.mvd[0][0]); mb_se (bw, r.mvd[0][1]);
    }
  }
  if (!i16) {
    const uint8_t* table = intra ? md_cbp_intra : md_cbp_inter;
    int code = 0;
    while (code < 48 && table[code] != cbp) code++;
    if (code >= 48) return false;
    mb_ue (bw, (uint32_t)code);
    if (!intra && (r.flags & LH264_MM_T8_CODED)) mb_put (bw, t8, 1);
  }
  if (cbp == 0 && !i16) { for (int i = 0; i < 24; i++) tc[i] = 0; return true; }
  {
    int d = (int)r.luma_qp - s.qp;
    if (d > 25) d -= 52; else if (d < -26) d += 52;
    mb_se (bw, d);
    s.qp = r.luma_qp;
  }
  if (i16) cw_residual_block (bw, sc, cw_nC (tcL ? tcL[3] : -1, tcT ? tcT[12] : -1), 16, lev, sc.t_i16dc, 1);
  for (int q = 0; q < 4; q++) {
    if (!(cbpL & (1 << q))) { for (int k = 0; k < 4; k++) tc[sc.t_scan4[4 * q + k]] = 0; continue; }
    for (int k = 0; k < 4; k++) {
      const int z = 4 * q + k, rr = sc.t_scan4[z], bx = rr & 3, by = rr >> 2;
      const int nA = bx > 0 ? tc[rr - 1] : (tcL ? tcL[rr + 3] : -1), nB = by > 0 ? tc[rr - 4] : (tcT ? tcT[rr + 12] : -1);
      int n;
      if (t8) n = cw_residual_block (bw, sc, cw_nC (nA, nB), 16, lev + 64 * q, sc.t_zigzag8 + k, 4);       // 4 interleaved sub-blocks
      else if (i16) n = cw_residual_block (bw, sc, cw_nC (nA, nB), 15, lev + 16 * z, sc.t_zigzag4 + 1, 1);
      else n = cw_residual_block (bw, sc, cw_nC (nA, nB), 16, lev + 16 * z, sc.t_zigzag4, 1);
      tc[rr] = (uint8_t)n;
    }
  }
  if (cbpC)
    for (int p = 0; p < 2; p++) cw_residual_block (bw, sc, -1, 4, lev + 16 * (16 + 4 * p), sc.t_cdc, 1);
  for (int p = 0; p < 2; p++)
    for (int b = 0; b < 4; b++) {
      const int base = 16 + 4 * p;
      if (cbpC & 2) {
        const int nA = (b & 1) ? tc[base + b - 1] : (tcL ? tcL[base + b + 1] : -1), nB = (b & 2) ? tc[base + b - 2] : (tcT ? tcT[base + b + 2] : -1);
        tc[base + b] = (uint8_t)cw_residual_block (bw, sc, cw_nC (nA, nB), 15, lev + 16 * (base + b), sc.t_zigzag4 + 1, 1);
      } else tc[base + b] = 0;
    }
  return true;
}


// ---- CABAC slice_data() writer (decompress direction of CABAC streams) ---------------------------------------------------
// Replaces EncoderState::writeMbSynCabac (decode_slice.cpp:1536-1539 -> encoder/core/src/svc_set_mb_syn_cabac.cpp, set_mb_syn_cabac.cpp):
// the arithmetic encoder of 9.3.4.2, and for every syntax element the binarisation and ctxIdx rule the host parser reads with
// (host/decoder.cpp cabac_*).  Per macroblock, MD_NB_STRIDE bytes of st.tcoef hold what context selection needs from a neighbour:
//   [0..3] coded_block_flag word (bits 0..15 luma raster, 16..19 Cb AC, 20..23 Cr AC, 24 luma DC, 25 Cb DC, 26 Cr DC)
//   [4] coded_block_pattern  [5] intra_chroma_pred_mode  [6] kind (1 skip, 2 inter, 3 I_NxN, 4 I16x16)  [7] transform 8x8
//   [8..11] ref_idx per 8x8  [12..43] |mvd| per raster 4x4 block and component, saturated at 128
enum { CB_SKIP = 1, CB_INTER = 2, CB_INXN = 3, CB_I16 = 4 };
MD_TABLE uint8_t cb_sig8[63] = { 0, 1, 2, 3, 4, 5, 5, 4, 4, 3, 3, 4, 4, 4, 5, 5, 4, 4, 4, 4, 3, 3, 6, 7, 7, 7, 8, 9, 10, 9, 8, 7,
                                 7, 6, 11, 12, 13, 11, 6, 7, 8, 9, 14, 10, 9, 8, 6, 11, 12, 13, 11, 6, 9, 14, 10, 9, 11, 12, 13, 11, 14, 10, 12 };
MD_TABLE uint8_t cb_last8[63] = { 0, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2,
                                  3, 3, 3, 3, 3, 3, 3, 3, 4, 4, 4, 4, 4, 4, 4, 4, 5, 5, 5, 5, 6, 6, 6, 6, 7, 7, 7, 7, 8, 8, 8 };
MD_FN uint32_t cb_word (const uint8_t* nb) { return (uint32_t)nb[0] | ((uint32_t)nb[1] << 8) | ((uint32_t)nb[2] << 16) | ((uint32_t)nb[3] << 24); }
MD_FN void cb_set_word (uint8_t* nb, uint32_t w) { nb[0] = (uint8_t)w; nb[1] = (uint8_t) (w >> 8); nb[2] = (uint8_t) (w >> 16); nb[3] = (uint8_t) (w >> 24); }

MD_FN void cb_slice_begin (MdBits& bw, MdScratch& sc, int cls, int slice_qp, uint32_t header_bits) {
  for (uint32_t i = header_bits & 7u; i & 7u; i++) mb_put (bw, 1, 1);       // cabac_alignment_one_bit
  const int qp = slice_qp < 0 ? 0 : (slice_qp > 51 ? 51 : slice_qp);
  for (int i = 0; i < 460; i++) {                                            // 9.3.1.1
    const int v = kCabacInitMN[cls][i], m = v >> 8, n = (int8_t) (v & 0xff);
    int pre = ((m * qp) >> 4) + n;
    pre = pre < 1 ? 1 : (pre > 126 ? 126 : pre);
    sc.cb_state[i] = pre <= 63 ? (uint8_t) ((63 - pre) << 1) : (uint8_t) (((pre - 64) << 1) | 1);
  }
  sc.cb_low = 0; sc.cb_range = 510; sc.cb_outstanding = 0; sc.cb_first = 1;
}
MD_FN void cb_put (MdBits& bw, MdScratch& sc, uint32_t bit) {                // PutBit
  if (sc.cb_first) sc.cb_first = 0; else mb_put (bw, bit, 1);
  while (sc.cb_outstanding) { const uint32_t n = sc.cb_outstanding > 16u ? 16u : sc.cb_outstanding; mb_put (bw, bit ? 0u : 0xffffu, (int)n); sc.cb_outstanding -= n; }
}
MD_FN void cb_renorm (MdBits& bw, MdScratch& sc) {                           // RenormE
  while (sc.cb_range < 256u) {
    if (sc.cb_low < 256u) cb_put (bw, sc, 0);
    else if (sc.cb_low >= 512u) { sc.cb_low -= 512u; cb_put (bw, sc, 1); }
    else { sc.cb_low -= 256u; sc.cb_outstanding++; }
    sc.cb_range <<= 1; sc.cb_low <<= 1;
  }
}
MD_FN_NOINLINE void cb_bin (MdBits& bw, MdScratch& sc, int ctx, uint32_t bin) {   // EncodeDecision
  uint8_t& s = sc.cb_state[ctx];
  const uint32_t idx = s >> 1;
  uint32_t mps = s & 1u;
  const uint32_t lps = kCabacRangeLps[(sc.cb_range >> 6) & 3u][idx];
  sc.cb_range -= lps;
  if (bin != mps) {
    sc.cb_low += sc.cb_range; sc.cb_range = lps;
    if (idx == 0) mps ^= 1u;
    s = (uint8_t) ((kCabacNextLps[idx] << 1) | mps);
  } else s = (uint8_t) (((idx < 62u ? idx + 1u : idx) << 1) | mps);
  cb_renorm (bw, sc);
}
MD_FN_NOINLINE void cb_bypass (MdBits& bw, MdScratch& sc, uint32_t bin) {         // EncodeBypass
  sc.cb_low <<= 1;
  if (bin) sc.cb_low += sc.cb_range;
  if (sc.cb_low >= 1024u) { cb_put (bw, sc, 1); sc.cb_low -= 1024u; }
  else if (sc.cb_low < 512u) cb_put (bw, sc, 0);
  else { sc.cb_low -= 512u; sc.cb_outstanding++; }
}
MD_FN void cb_terminate (MdBits& bw, MdScratch& sc, uint32_t bin) {               // EncodeTerminate + EncodeFlush: the last bit written is the
  sc.cb_range -= 2u;                                                               // rbsp_stop_one_bit
  if (bin) {
    sc.cb_low += sc.cb_range; sc.cb_range = 2u;
    cb_renorm (bw, sc);
    cb_put (bw, sc, (sc.cb_low >> 9) & 1u);
    mb_put (bw, ((sc.cb_low >> 7) & 3u) | 1u, 2);
  } else cb_renorm (bw, sc);
}

// residual_block_cabac(): coefficient i of the block is base[scan[i * 1 + off]]; returns the number of non-zero coefficients
MD_FN_NOINLINE int cb_block (MdBits& bw, MdScratch& sc, int cat, int cbf_inc, const int16_t* base, const uint8_t* scan, int off) {
  const int n_max = cat == 5 ? 64 : (cat == 3 ? 4 : ((cat == 1 || cat == 4) ? 15 : 16));
  const int sig0 = cat == 5 ? 402 : 105 + (cat == 0 ? 0 : (cat == 1 ? 15 : (cat == 2 ? 29 : (cat == 3 ? 44 : 47))));
  const int last0 = cat == 5 ? 417 : sig0 + 61;
  const int abs0 = cat == 5 ? 426 : 227 + (cat == 0 ? 0 : (cat == 1 ? 10 : (cat == 2 ? 20 : (cat == 3 ? 30 : 39))));
  int n = 0, last = -1;
  for (int i = 0; i < n_max; i++) if (base[scan[off + i]]) { n++; last = i; }
  if (cbf_inc >= 0) cb_bin (bw, sc, 85 + 4 * cat + cbf_inc, n != 0);
  if (!n) return 0;
  for (int i = 0; i < n_max - 1; i++) {
    const int inc_s = cat == 5 ? cb_sig8[i] : (cat == 3 ? (i < 2 ? i : 2) : i);
    const uint32_t sig = base[scan[off + i]] != 0;
    cb_bin (bw, sc, sig0 + inc_s, sig);
    if (!sig) continue;
    const int inc_l = cat == 5 ? cb_last8[i] : (cat == 3 ? (i < 2 ? i : 2) : i);
    cb_bin (bw, sc, last0 + inc_l, i == last);
    if (i == last) break;
  }
  int eq1 = 0, gt1 = 0;
  const int cap = cat == 3 ? 3 : 4;
  for (int i = last; i >= 0; i--) {
    const int v = base[scan[off + i]];
    if (!v) continue;
    const int a = v < 0 ? -v : v;
    const int c0 = abs0 + (gt1 ? 0 : (1 + eq1 < 4 ? 1 + eq1 : 4));
    if (a == 1) { cb_bin (bw, sc, c0, 0); eq1++; }
    else {
      cb_bin (bw, sc, c0, 1);
      const int c1 = abs0 + 5 + (gt1 < cap ? gt1 : cap);
      for (int k = 2; k < a && k < 15; k++) cb_bin (bw, sc, c1, 1);
      if (a < 15) cb_bin (bw, sc, c1, 0);
      else {                                           // Exp-Golomb (k = 0) suffix of a - 15
        int rest = a - 15, e = 0;
        while (rest >= (1 << e)) { cb_bypass (bw, sc, 1); rest -= 1 << e; e++; }
        cb_bypass (bw, sc, 0);
        while (e--) cb_bypass (bw, sc, (uint32_t) (rest >> e) & 1u);
      }
      gt1++;
    }
    cb_bypass (bw, sc, v < 0);
  }
  return n;
}
MD_FN int cb_cbf_inc (const uint8_t* cur, const uint8_t* L, const uint8_t* T, bool intra, int bit_cur_a, int bit_left, int bit_cur_b, int bit_top) {
  const int dflt = intra ? 1 : 0;
  const int a = bit_cur_a >= 0 ? (int) ((cb_word (cur) >> bit_cur_a) & 1u) : (L ? (int) ((cb_word (L) >> bit_left) & 1u) : dflt);
  const int b = bit_cur_b >= 0 ? (int) ((cb_word (cur) >> bit_cur_b) & 1u) : (T ? (int) ((cb_word (T) >> bit_top) & 1u) : dflt);
  return a + 2 * b;
}
MD_FN_NOINLINE void cb_mvd (MdBits& bw, MdScratch& sc, uint8_t* cur, const uint8_t* L, const uint8_t* T, int bx, int by, int bw4, int bh4, int comp, int v) {
  int sum = 0;
  if (bx > 0) sum += cur[12 + 2 * (by * 4 + bx - 1) + comp]; else if (L) sum += L[12 + 2 * (by * 4 + 3) + comp];
  if (by > 0) sum += cur[12 + 2 * ((by - 1) * 4 + bx) + comp]; else if (T) sum += T[12 + 2 * (12 + bx) + comp];
  int ctx = comp ? 47 : 40;
  const int a = v < 0 ? -v : v;
  cb_bin (bw, sc, ctx + (sum < 3 ? 0 : (sum > 32 ? 2 : 1)), a != 0);
  if (a) {
    ctx += 3;
    int k = 1;
    for (; k < a && k < 9; k++) { cb_bin (bw, sc, ctx, 1); if (k < 4) ctx++; }
    if (a < 9) cb_bin (bw, sc, ctx, 0);
    else {                                             // UEG3 suffix of a - 9
      int rest = a - 9, e = 3;
      while (rest >= (1 << e)) { cb_bypass (bw, sc, 1); rest -= 1 << e; e++; }
      cb_bypass (bw, sc, 0);
      while (e--) cb_bypass (bw, sc, (uint32_t) (rest >> e) & 1u);
    }
    cb_bypass (bw, sc, v < 0);
  }
  const uint8_t s8 = (uint8_t) (a > 128 ? 128 : a);
  for (int y = by; y < by + bh4; y++) for (int x = bx; x < bx + bw4; x++) cur[12 + 2 * (y * 4 + x) + comp] = s8;
}
MD_FN void cb_ref (MdBits& bw, MdScratch& sc, const uint8_t* L, const uint8_t* T, int q, const int8_t* vis, int ref) {
  const bool a = (q & 1) ? vis[q - 1] > 0 : (L && (int8_t)L[8 + q + 1] > 0);
  const bool b = (q & 2) ? vis[q - 2] > 0 : (T && (int8_t)T[8 + q + 2] > 0);
  int ctx = (a ? 1 : 0) + (b ? 2 : 0);
  for (int i = 0; i < ref; i++) { cb_bin (bw, sc, 54 + ctx, 1); ctx = (ctx >> 2) + 4; }
  cb_bin (bw, sc, 54 + ctx, 0);
}
MD_FN void cb_intra_type (MdBits& bw, MdScratch& sc, int base, bool i_slice, int inc0, int t) {   // t: 0 I_NxN, 1..24 Intra16x16
  cb_bin (bw, sc, i_slice ? base + inc0 : base, t != 0);
  if (!t) return;
  if (i_slice) base += 2;
  cb_terminate (bw, sc, 0);
  const int v = t - 1, pred = v & 3, chroma = (v >> 2) % 3, luma = v >= 12;
  cb_bin (bw, sc, base + 1, (uint32_t)luma);
  cb_bin (bw, sc, base + 2, chroma != 0);
  if (chroma) cb_bin (bw, sc, base + 2 + (i_slice ? 1 : 0), chroma == 2);
  cb_bin (bw, sc, base + 3 + (i_slice ? 1 : 0), (uint32_t) (pred >> 1));
  cb_bin (bw, sc, base + 3 + (i_slice ? 2 : 0), (uint32_t) (pred & 1));
}

// macroblock_layer() with mb_skip_flag in front and end_of_slice_flag behind, for one record.  cur = this macroblock's neighbour
// record (written), L / T = the left / top macroblock's (nullptr when not available).  Returns false for I_PCM.
MD_FN_NOINLINE bool cb_mb (MdBits& bw, MdScratch& sc, CwSlice& s, uint8_t* cur, const uint8_t* L, const uint8_t* T) {
  const lh264_model_mb& r = sc.r;
  const int16_t* lev = sc.lev;
  const bool p_slice = s.slice_type == 0;
  const bool coded = (r.flags & LH264_MM_WRITE_BLOCK) != 0;
  for (int i = 0; i < MD_NB_STRIDE; i++) cur[i] = 0;
  if (p_slice) cb_bin (bw, sc, 11 + (L && L[6] != CB_SKIP ? 1 : 0) + (T && T[6] != CB_SKIP ? 1 : 0), !coded);
  if (!coded) {
    cur[6] = CB_SKIP;
    s.last_dqp_nz = 0;
    cb_terminate (bw, sc, (r.flags & LH264_MM_STOP_BIT) != 0);
    return true;
  }
  const int tcode = r.type_code;
  if (tcode >= 8) return false;
  const int cbp = r.cbp, cbpL = cbp & 15, cbpC = cbp >> 4;
  const bool intra = tcode <= 2, i16 = tcode == 1;
  bool t8 = (r.flags & LH264_MM_T8) != 0;
  for (int q = 0; q < 4; q++) cur[8 + q] = (uint8_t) (intra ? 0xff : 0);
  auto t8_flag = [&] (bool v) { cb_bin (bw, sc, 399 + (L && L[7] ? 1 : 0) + (T && T[7] ? 1 : 0), v); };
  if (intra) {
    int t = 0;
    if (i16) { const int raw = r.i16_mode > 3 ? 2 : r.i16_mode; t = 1 + raw + 4 * cbpC + (cbpL == 15 ? 12 : 0); }
    if (p_slice) { cb_bin (bw, sc, 14, 1); cb_intra_type (bw, sc, 17, false, 0, t); }
    else cb_intra_type (bw, sc, 3, true, (L && L[6] != CB_INXN ? 1 : 0) + (T && T[6] != CB_INXN ? 1 : 0), t);
    cur[6] = (uint8_t) (i16 ? CB_I16 : CB_INXN);
    if (!i16) {
      if (s.transform_8x8_mode) t8_flag (tcode == 2);
      t8 = tcode == 2;
      const int n = tcode == 2 ? 4 : 16;
      for (int i = 0; i < n; i++) {
        const int mode = r.pred_mode[i] & 15, pred = r.pred_mode[i] >> 4;
        cb_bin (bw, sc, 68, mode == pred);
        if (mode != pred) { const int rem = mode < pred ? mode : mode - 1; cb_bin (bw, sc, 69, rem & 1); cb_bin (bw, sc, 69, (rem >> 1) & 1); cb_bin (bw, sc, 69, (rem >> 2) & 1); }
      }
    }
    const int cm = r.chroma_mode > 3 ? 0 : r.chroma_mode;
    auto cnd = [] (const uint8_t* n) { return n && (n[6] == CB_INXN || n[6] == CB_I16) && n[5] != 0; };
    cb_bin (bw, sc, 64 + (cnd (L) ? 1 : 0) + (cnd (T) ? 1 : 0), cm != 0);
    if (cm) { cb_bin (bw, sc, 67, cm != 1); if (cm != 1) cb_bin (bw, sc, 67, cm == 3); }
    cur[5] = (uint8_t)cm;
  } else {
    cur[6] = CB_INTER;
    cb_bin (bw, sc, 14, 0);
    if (tcode == 3) { cb_bin (bw, sc, 15, 0); cb_bin (bw, sc, 16, 0); }
    else if (tcode == 6 || tcode == 7) { cb_bin (bw, sc, 15, 0); cb_bin (bw, sc, 16, 1); }
    else { cb_bin (bw, sc, 15, 1); cb_bin (bw, sc, 17, tcode == 4); }       // 16x8 -> 1, 8x16 -> 0
    const int nref = s.num_ref_idx_l0;
    int8_t vis[4] = { -1, -1, -1, -1 };
    if (tcode == 6 || tcode == 7) {
      for (int q = 0; q < 4; q++) {
        const int sb = r.sub_mb_type[q];
        if (sb == 1) cb_bin (bw, sc, 21, 1);
        else { cb_bin (bw, sc, 21, 0); if (sb == 2) cb_bin (bw, sc, 22, 0); else { cb_bin (bw, sc, 22, 1); cb_bin (bw, sc, 23, sb == 4); } }
      }
      for (int q = 0; q < 4; q++) { if (nref > 1) cb_ref (bw, sc, L, T, q, vis, r.ref_idx[q]); vis[q] = r.ref_idx[q]; }
      for (int q = 0; q < 4; q++) cur[8 + q] = (uint8_t)vis[q];
      for (int q = 0; q < 4; q++) {
        const int sb = r.sub_mb_type[q], qx = (q & 1) * 2, qy = (q >> 1) * 2;
        const int cnt = sb == 1 ? 1 : (sb == 8 ? 4 : 2);
        for (int j = 0; j < cnt; j++) {
          int bx = qx, by = qy, w4 = 2, h4 = 2;
          if (sb == 2) { by += j; h4 = 1; } else if (sb == 4) { bx += j; w4 = 1; } else if (sb == 8) { bx += j & 1; by += j >> 1; w4 = h4 = 1; }
          cb_mvd (bw, sc, cur, L, T, bx, by, w4, h4, 0, r.mvd[by * 4 + bx][0]);
          cb_mvd (bw, sc, cur, L, T, bx, by, w4, h4, 1, r.mvd[by * 4 + bx][1]);
        }
      }
    } else {
      const int nparts = tcode == 3 ? 1 : 2;
      for (int i = 0; i < nparts; i++) {
        const int q0 = tcode == 4 ? 2 * i : i;
        if (nref > 1) cb_ref (bw, sc, L, T, q0, vis, r.ref_idx[i]);
        if (tcode == 3) vis[0] = vis[1] = vis[2] = vis[3] = r.ref_idx[0];
        else if (tcode == 4) vis[2 * i] = vis[2 * i + 1] = r.ref_idx[i];
        else vis[i] = vis[i + 2] = r.ref_idx[i];
      }
      for (int q = 0; q < 4; q++) cur[8 + q] = (uint8_t)vis[q];
      for (int i = 0; i < nparts; i++) {
        int bx = 0, by = 0, w4 = 4, h4 = 4;
        if (tcode == 4) { by = 2 * i; h4 = 2; } else if (tcode == 5) { bx = 2 * i; w4 = 2; }
        cb_mvd (bw, sc, cur, L, T, bx, by, w4, h4, 0, r.mvd[by * 4 + bx][0]);
        cb_mvd (bw, sc, cur, L, T, bx, by, w4, h4, 1, r.mvd[by * 4 + bx][1]);
      }
    }
  }
  if (!i16) {                                          // coded_block_pattern (9.3.2.6, 9.3.3.1.1.4)
    const int ca = L ? L[4] : 0x0f, cb = T ? T[4] : 0x0f;
    cb_bin (bw, sc, 73 + !(ca & 2) + 2 * !(cb & 4), cbpL & 1);
    cb_bin (bw, sc, 73 + !(cbpL & 1) + 2 * !(cb & 8), (cbpL >> 1) & 1);
    cb_bin (bw, sc, 73 + !(ca & 8) + 2 * !(cbpL & 1), (cbpL >> 2) & 1);
    cb_bin (bw, sc, 73 + !(cbpL & 4) + 2 * !(cbpL & 2), (cbpL >> 3) & 1);
    const int xa = ca >> 4, xb = cb >> 4;
    cb_bin (bw, sc, 77 + (xa > 0 ? 1 : 0) + (xb > 0 ? 2 : 0), cbpC != 0);
    if (cbpC) cb_bin (bw, sc, 77 + 4 + (xa == 2 ? 1 : 0) + (xb == 2 ? 2 : 0), cbpC == 2);
    if (!intra && (r.flags & LH264_MM_T8_CODED)) t8_flag (t8);
  }
  cur[4] = (uint8_t)cbp; cur[7] = t8 ? 1 : 0;
  if (cbp == 0 && !i16) { s.last_dqp_nz = 0; cb_terminate (bw, sc, (r.flags & LH264_MM_STOP_BIT) != 0); return true; }
  {
    int d = (int)r.luma_qp - s.qp;
    if (d > 25) d -= 52; else if (d < -26) d += 52;
    const int v = d > 0 ? 2 * d - 1 : -2 * d;          // mapped value, unary
    int ctx = s.last_dqp_nz ? 1 : 0;
    for (int i = 0; i < v; i++) { cb_bin (bw, sc, 60 + ctx, 1); ctx = 2 + (ctx >> 1); }
    cb_bin (bw, sc, 60 + ctx, 0);
    s.last_dqp_nz = d != 0;
    s.qp = r.luma_qp;
  }
  uint32_t cbf = 0;
  if (i16) { if (cb_block (bw, sc, 0, cb_cbf_inc (cur, L, T, intra, -1, 24, -1, 24), lev, sc.t_i16dc, 0)) { cbf |= 1u << 24; cb_set_word (cur, cbf); } }
  for (int q = 0; q < 4; q++) {
    if (!(cbpL & (1 << q))) continue;
    if (t8) {
      cb_block (bw, sc, 5, -1, lev + 64 * q, sc.t_zigzag8, 0);
      for (int k = 0; k < 4; k++) cbf |= 1u << sc.t_scan4[4 * q + k];
      cb_set_word (cur, cbf);
      continue;
    }
    for (int k = 0; k < 4; k++) {
      const int z = 4 * q + k, rr = sc.t_scan4[z], bx = rr & 3, by = rr >> 2;
      const int inc = cb_cbf_inc (cur, L, T, intra, bx > 0 ? rr - 1 : -1, rr + 3, by > 0 ? rr - 4 : -1, rr + 12);
      const int n = i16 ? cb_block (bw, sc, 1, inc, lev + 16 * z, sc.t_zigzag4, 1) : cb_block (bw, sc, 2, inc, lev + 16 * z, sc.t_zigzag4, 0);
      if (n) { cbf |= 1u << rr; cb_set_word (cur, cbf); }
    }
  }
  if (cbpC)
    for (int p = 0; p < 2; p++)
      if (cb_block (bw, sc, 3, cb_cbf_inc (cur, L, T, intra, -1, 25 + p, -1, 25 + p), lev + 16 * (16 + 4 * p), sc.t_cdc, 0)) { cbf |= 1u << (25 + p); cb_set_word (cur, cbf); }
  if (cbpC & 2)
    for (int p = 0; p < 2; p++)
      for (int b = 0; b < 4; b++) {
        const int bit = 16 + 4 * p + b;
        const int inc = cb_cbf_inc (cur, L, T, intra, (b & 1) ? bit - 1 : -1, bit + 1, (b & 2) ? bit - 2 : -1, bit + 2);
        if (cb_block (bw, sc, 4, inc, lev + 16 * bit, sc.t_zigzag4, 1)) { cbf |= 1u << bit; cb_set_word (cur, cbf); }
      }
  cb_terminate (bw, sc, (r.flags & LH264_MM_STOP_BIT) != 0);
  return true;
}

// ---- one slice ------------------------------------------------------------------------------------------------------
// FreqImage::updateFrame (decoded_macroblock.h:120-165), serial form; the CUDA build replaces it by a warp-wide one.
MD_FN void md_update_frame_serial (CtxRec* f, int len) {
  int run_start = 0;
  for (int i = 0; i < len; i++) {
    bool zeroed = true;
    for (int k = 0; k < 24; k++) if (f[i].cnt[k]) { zeroed = false; break; }
    f[i].flags = (uint8_t) (zeroed ? (f[i].flags | 2) : (f[i].flags & ~2));
    if (!zeroed) {
      const int run = i - run_start;
      for (int j = i - run + 1; j <= i && run > 0; j++) f[j].cachedSkips = (uint16_t)run;
      run_start = i + 1;
    }
  }
}
// frame flip + resize of WelsDecodeSlice decode_slice.cpp:3032-3046; returns priorValid.  (updateFrame itself runs between
// md_slice_begin_flip and the rest so that the CUDA build can do it with all lanes.)
MD_FN CtxRec* md_slice_flip (MdStream& st, int frame_num) {
  if (frame_num != st.lastFrameId) { st.cur ^= 1; st.lastFrameId = frame_num; }
  return st.frame[1 - st.cur];
}
MD_FN bool md_slice_resize (MdStream& st, int mb_w, int mb_h) {
  if (st.width == mb_w && st.height == mb_h) return true;
  if (mb_w <= 0 || mb_h <= 0) { st.error = MD_ERR_SYNTAX; return false; }   // a crafted default stream must not reach the index arithmetic below
  const int n = mb_w * mb_h;
  if (n > st.max_mbs) { st.error = MD_ERR_DIMS; return false; }
  for (int k = 0; k < 2; k++)
    for (int j = st.frame_len; j < n; j++) { uint32_t* q = reinterpret_cast<uint32_t*> (&st.frame[k][j]); for (int w = 0; w < 8; w++) q[w] = 0; }
  for (int j = 0; j < n; j++) st.mb_slice[j] = -1;
  st.width = mb_w; st.height = mb_h; st.frame_len = n;
  return false;
}

struct MdSliceIn {                 // = lh264_decode_slice (include/lh264_b200.h)
  uint32_t first_mb; int32_t frame_num; uint16_t mb_w, mb_h; uint8_t slice_type, num_ref_idx_l0, slice_qp, flags; uint32_t header_bits;
};

// Decodes the slice's macroblocks from the side streams and appends their CAVLC syntax to bw.  sc = per-stream scratch,
// c = the stream's decoder context.
// Returns the number of macroblocks (skipped ones included); st.error is set on failure.
MD_FN int md_slice (MdStream& st, MdCtx& c, const MdSliceIn& sl, bool prior_valid, MdBits& bw, MdScratch& sc, uint8_t* pad_bits_out, uint8_t* n_pad_out) {
  lh264_model_mb& r = sc.r;
  int16_t* lev = sc.lev;
  uint8_t* cnt = sc.cnt;
  for (int i = 0; i < 384; i++) lev[i] = 0;
  for (int i = 0; i < 24; i++) cnt[i] = 0;
  const int W = st.width, n_mbs = st.frame_len;
  CtxRec* cur = st.frame[st.cur];
  CtxRec* last = st.frame[1 - st.cur];
  const uint32_t stype = sl.slice_type == 0 ? 0u : 1u;
  const bool p_slice = sl.slice_type == 0;
  const bool t8mode = sl.flags & 1, cip = sl.flags & 2, cabac = sl.flags & 4;
  const int slice_no = ++st.slice_counter;
  CwSlice cs = { sl.slice_type, sl.num_ref_idx_l0, sl.slice_qp, t8mode ? 1 : 0, cabac ? 1 : 0, 0 };
  if (cabac) cb_slice_begin (bw, sc, p_slice ? 1 + ((sl.flags >> 3) & 3) : 0, sl.slice_qp, sl.header_bits);
  // first_mb_in_slice comes from an untrusted default stream: outside the picture it would index last[], tcoef[] and mb_slice[] out of bounds
  if (W <= 0 || n_mbs <= 0 || sl.first_mb >= (uint32_t)n_mbs) { st.error = MD_ERR_SYNTAX; return 0; }
  int mb_addr = (int)sl.first_mb, mb_in_slice = 0, cached_qp = 0, last_dqp = 0, pending_skips = 0;
  bool run_known = false, stop = false;
  auto avail = [&] (int mb) -> bool { return mb >= 0 && st.mb_slice[mb] == slice_no; };
  auto avail_intra = [&] (int mb) -> bool { return avail (mb) && (!cip || st.mb_intra[mb]); };
  while (!stop && mb_addr < n_mbs && !st.error) {
    const int mbx = mb_addr % W, mby = mb_addr / W;
    const CtxRec* P = prior_valid && (last[mb_addr].flags & 1) ? &last[mb_addr] : nullptr;
    { uint32_t* q = reinterpret_cast<uint32_t*> (&r); for (unsigned w = 0; w < sizeof (r) / 4; w++) q[w] = 0; }
    r.mb_addr = (uint32_t)mb_addr; r.mb_in_slice = (uint16_t) (mb_in_slice < 65535 ? mb_in_slice : 65535);
    const uint32_t stop_idx = P_STOP + (uint32_t) (mb_in_slice < 2047 ? mb_in_slice : 2047);
    bool skipped = false;
    if (pending_skips > 0) skipped = true;
    else if (!run_known) {                             // mb_skip_run: decode_slice.cpp:2210-2217, getSkipRunPrior macroblock_model.cpp:375-390
      const int csk = P ? P->cachedSkips : 0;
      const int prior = csk / 8 + (csk % 8 ? 1 : 0);
      if (prior >= 512) { st.error = MD_ERR_CONTEXT; break; }
      const uint32_t run = md_branch (c, CT_SKIP, 9, P_SKIP + (uint32_t)prior * 512u);
      r.flags |= LH264_MM_EMIT_SKIP_RUN; r.skip_run = (uint16_t)run;
      if (p_slice && run > 0) {
        if ((int)run > n_mbs - mb_addr) { st.error = MD_ERR_SYNTAX; break; }
        pending_skips = (int)run; skipped = true;
      }
    }
    run_known = false;
    uint8_t* tc = &st.tcoef[(size_t)mb_addr * MD_NB_STRIDE];
    const bool hasLs = mbx > 0 && avail (mb_addr - 1), hasTs = mby > 0 && avail (mb_addr - W);   // before this MB joins the slice
    if (skipped) {
      pending_skips--;
      if (pending_skips == 0) {                        // last MB of the run: stop bit (:2223-2226); a coded MB follows without a new run
        r.flags |= LH264_MM_STOP_IN_SKIP;
        stop = md_bit (c, CT_SKIP_END, stop_idx) != 0;
        if (stop) r.flags |= LH264_MM_STOP_BIT;
        run_known = !cabac;                            // CABAC: a one-macroblock run (0 or 1) is coded in front of EVERY macroblock (:2207-2209)
      }
      cur[mb_addr] = last[mb_addr];                    // :3105
      if (cabac) cb_mb (bw, sc, cs, tc, hasLs ? &st.tcoef[(size_t) (mb_addr - 1) * MD_NB_STRIDE] : nullptr, hasTs ? &st.tcoef[(size_t) (mb_addr - W) * MD_NB_STRIDE] : nullptr);
      else cw_mb (bw, sc, cs, tc, nullptr, nullptr);
      uint8_t* pm = &st.ipm[(size_t)mb_addr * 16];
      for (int i = 0; i < 16; i++) pm[i] = 2;
      st.mb_intra[mb_addr] = 0; st.mb_slice[mb_addr] = slice_no;
      mb_addr++; mb_in_slice++;
      continue;
    }
    // ---- a coded macroblock: decode_slice.cpp:2244-2434 in the reading direction ----
    r.flags |= LH264_MM_WRITE_BLOCK;
    const CtxRec* L = mbx > 0 && (cur[mb_addr - 1].flags & 1) ? &cur[mb_addr - 1] : nullptr;          // Neighbors::init macroblock_model.cpp:9-44
    const CtxRec* A = mby > 0 && (cur[mb_addr - W].flags & 1) ? &cur[mb_addr - W] : nullptr;
    stop = md_bit (c, CT_SKIP_END, stop_idx) != 0;                                                   // :2252
    if (stop) r.flags |= LH264_MM_STOP_BIT;
    int tcd;
    {
      int prior = 15, prevType = 15;
      if (A) prior = A->type_code;
      if (L) prior = L->type_code;
      if (P) prevType = P->type_code;
      tcd = (int)md_branch (c, CT_MB_TYPE, 4, P_MBTYPE + ((uint32_t) (prior + prevType) * 2u + (stype == 0)) * 16u);
    }
    if (tcd >= (int)NT) { st.error = MD_ERR_SYNTAX; break; }
    r.type_code = (uint8_t)tcd;
    const int cbpC = (int)md_branch (c, CT_CBPL, 2, P_CBPC + ((uint32_t) (P ? P->cbp >> 4 : 0) * NT + tcd) * 4u);
    const int cbpL = (int)md_branch (c, CT_CBPL, 4, P_CBPL + ((uint32_t) (P ? P->cbp & 15 : 0) * NT + tcd) * 16u);
    r.cbp = (uint8_t) (cbpL | (cbpC << 4));
    {
      const uint32_t cls = last_dqp < 0 ? 0u : (last_dqp == 0 ? 1u : 2u);
      const uint32_t sw = md_zero_to_pow2 (c, CT_QPL, 7, P_QPL + ((mb_in_slice == 0 ? 3u : 0u) + cls) * 128u, 0);
      const int d = (sw & 1u) ? - (int) (sw >> 1) - 1 : (int) (sw >> 1);
      r.qp_delta = (int8_t)d;
      r.flags |= (uint16_t) (cls << LH264_MM_LASTDQP_SHIFT);
      r.luma_qp = (uint8_t) (cached_qp + d);
      cached_qp = r.luma_qp;
      if (d) last_dqp = d;
    }
    int num_ref;
    {
      const int pr = P ? P->num_ref : 0;
      if (pr >= 16) { st.error = MD_ERR_CONTEXT; break; }
      num_ref = (int)md_branch (c, CT_REF, 4, P_NUMREF + ((uint32_t)pr * NT + tcd) * 16u);
    }
    int refBits = 0;
    while ((1 << refBits) < num_ref) refBits++;
    {
      const uint32_t pc = P ? (uint32_t) (P->chroma_mode < 6 ? P->chroma_mode : 6) : 7u;
      r.chroma_mode = (uint8_t)md_zero_to_pow2 (c, CT_8x8, 3, P_CMODE + pc * 8u, pc);
      const uint32_t pl = P ? (uint32_t) (P->i16_mode < 6 ? P->i16_mode : 6) : 7u;
      r.i16_mode = (uint8_t)md_zero_to_pow2 (c, CT_16x16, 3, P_CMODE + pl * 8u, pl);
    }
    uint8_t* pm = &st.ipm[(size_t)mb_addr * 16];
    for (int i = 0; i < 16; i++) pm[i] = 2;
    if (tcd == 0 || tcd == 2) {                        // Intra4x4 / Intra8x8 prediction modes: 8.3.1.1 / 8.3.2.1
      const int nL = mbx > 0 ? mb_addr - 1 : -1, nT = mby > 0 ? mb_addr - W : -1, nTL = (mbx > 0 && mby > 0) ? mb_addr - W - 1 : -1;
      const bool aL = avail_intra (nL), aT = avail_intra (nT), aTL = avail_intra (nTL);
      const bool i8 = tcd == 2;
      const int n = i8 ? 4 : 16;
      for (int i = 0; i < n; i++) {
        const int rr = i8 ? md_scan4[4 * i] : md_scan4[i], bx = rr & 3, by = rr >> 2;
        int a, b;
        bool okA = true, okB = true;
        if (bx > 0) a = pm[rr - 1]; else if (aL) a = st.ipm[(size_t)nL * 16 + rr + 3]; else { a = 2; okA = false; }
        if (by > 0) b = pm[rr - 4]; else if (aT) b = st.ipm[(size_t)nT * 16 + rr + 12]; else { b = 2; okB = false; }
        const int pred = (!okA || !okB) ? 2 : (a < b ? a : b);
        int mode;
        if (!i8) {
          const bool l = bx > 0 ? true : aL, tp = by > 0 ? true : aT;
          const bool tl = (bx > 0 && by > 0) ? true : (bx > 0 ? aT : (by > 0 ? aL : aTL));
          const uint32_t av = (l ? 4u : 0u) | (tp ? 2u : 0u) | (tl ? 1u : 0u);
          r.pred_avail[i] = (uint8_t)av;
          mode = (int)md_branch (c, CT_PRED_MODE, 4, P_PREDMODE + (((uint32_t)tcd * 8u + av) * 9u + (uint32_t) (pred < 8 ? pred : 8)) * 16u);
          pm[rr] = (uint8_t)mode;
        } else {
          mode = (int)md_branch (c, CT_PRED_MODE, 4, P_PREDMODE + (((uint32_t)tcd * 8u + 6u) * 9u + 1u) * 16u);
          pm[rr] = pm[rr + 1] = pm[rr + 4] = pm[rr + 5] = (uint8_t)mode;
        }
        if (mode > 8) { st.error = MD_ERR_SYNTAX; break; }
        r.pred_mode[i] = (uint8_t) (mode | (pred << 4));
      }
      if (st.error) break;
    }
    auto read_mv = [&] (int b) {
      const uint32_t p = P_MVD + ((uint32_t) (tcd - 3) * 16u + (uint32_t)b) * 16u;
      r.mvd[b][0] = (int16_t)md_uegk (c, 9, p, 3, 3, CT_MVX, CT_MVX, CT_MVX, CT_MVX);
      r.mvd[b][1] = (int16_t)md_uegk (c, 9, p, 3, 3, CT_MVY, CT_MVY, CT_MVY, CT_MVY);
    };
    bool sub_lt8 = false;
    if (tcd == 2) {
      for (int i = 0; i < 4; i++) r.sub_mb_type[i] = (uint8_t)md_branch (c, CT_SUB_MB, 8, P_SUBMB + (uint32_t)tcd * 256u);
      for (int i = 0; i < 4; i++) r.ref_idx[i] = (int8_t)md_raw (c, CT_REF, refBits);
    } else if (tcd == 6 || tcd == 7) {
      for (int i = 0; i < 4; i++) { r.sub_mb_type[i] = (uint8_t)md_branch (c, CT_SUB_MB, 8, P_SUBMB + (uint32_t)tcd * 256u); if (r.sub_mb_type[i] != 1) sub_lt8 = true; }
      if (tcd == 6) for (int i = 0; i < 4; i++) r.ref_idx[i] = (int8_t)md_raw (c, CT_REF, refBits);
      for (int i = 0; i < 4 && !st.error; i++) {
        const int sub = r.sub_mb_type[i];
        if (sub != 1 && sub != 2 && sub != 4 && sub != 8) { st.error = MD_ERR_SYNTAX; break; }
        const int cnt = sub == 1 ? 1 : (sub == 8 ? 4 : 2), step = sub == 2 ? 2 : 1;
        for (int j = 0; j < cnt; j++) read_mv (md_scan4[(i << 2) + j * step]);
      }
      if (st.error) break;
    } else if (tcd == 4 || tcd == 5) {
      for (int i = 0; i < 2; i++) r.ref_idx[i] = (int8_t)md_raw (c, CT_REF, refBits);
      for (int i = 0; i < 2; i++) read_mv (tcd == 4 ? i * 8 : i * 2);
    } else if (tcd == 3) {
      r.ref_idx[0] = (int8_t)md_raw (c, CT_REF, refBits);
      read_mv (0);
    }
    bool t8 = tcd == 2;
    if (tcd >= 3 && tcd <= 7 && cbpL && t8mode && !((tcd == 6 || tcd == 7) && sub_lt8)) {               // needParseTransformSize8x8
      r.flags |= LH264_MM_T8_CODED;
      t8 = md_bit (c, CT_T8, P_T8 + (uint32_t)tcd * 64u + (uint32_t) (r.luma_qp < 63 ? r.luma_qp : 63)) != 0;
    }
    if (t8) r.flags |= LH264_MM_T8;
    // ---- coefficients ----
    for (int b = 0; b < 24; b++)                       // only the chunks the previous coded macroblock left non-zero
      if (cnt[b]) { for (int i = 0; i < 16; i++) lev[16 * b + i] = 0; cnt[b] = 0; }
    if (tcd == 1)
      for (int i = 0; i < 16; i++) {
        const uint32_t p = P_LDC + (((uint32_t)i * 2u + stype) * NT + (uint32_t)tcd) * 16u;
        lev[16 * i] = (int16_t)md_int (c, p + I_EXP, 3, p + I_MANT, 0, p + I_ZERO, p + I_SIGN, CT_LDC, CT_LDC, CT_LDC, CT_LDC);
        cnt[i] = lev[16 * i] != 0;
      }
    if (cbpC == 1 || cbpC == 2)
      for (int i = 0; i < 8; i++) {
        const uint32_t p = P_CDC + (((uint32_t)i * 2u + stype) * NT + (uint32_t)tcd) * 16u;
        lev[16 * (16 + i)] = (int16_t)md_int (c, p + I_EXP, 3, p + I_MANT, 0, p + I_ZERO, p + I_SIGN, CT_CRDC, CT_CRDC, CT_CRDC, CT_CRDC);
        cnt[16 + i] = lev[16 * (16 + i)] != 0;
      }
    auto m2 = [] (int v) { return (uint32_t) (v < 2 ? v : 2); };
    const bool emit_ldc = tcd != 1;
    for (int q = 0; q < 4; q++) {
      if (!(cbpL & (1 << q))) continue;
      if (t8) {                                         // getNonzerosPrior8x8 macroblock_model.cpp:529-560
        const int idx = q * 4;
        int past = 0, left = 0, above = 0;
        if (P) past = P->cnt[idx] + P->cnt[idx + 1] + P->cnt[idx + 2] + P->cnt[idx + 3];
        if ((q & 1) == 0) { if (L) { const int k = (q + 1) * 4; left = L->cnt[k] + L->cnt[k + 1] + L->cnt[k + 2] + L->cnt[k + 3]; } }
        else { const int k = (q - 1) * 4; left = cnt[k] + cnt[k + 1] + cnt[k + 2] + cnt[k + 3]; }
        if ((q & 2) == 0) { if (A) { const int k = (q + 2) * 4; above = A->cnt[k] + A->cnt[k + 1] + A->cnt[k + 2] + A->cnt[k + 3]; } }
        else { const int k = (q - 2) * 4; above = cnt[k] + cnt[k + 1] + cnt[k + 2] + cnt[k + 3]; }
        if (tcd > 7) { st.error = MD_ERR_CONTEXT; break; }
        const uint32_t nzp = P_NZ8 + (((stype * NT + (uint32_t)tcd) * 3u + m2 (past)) * 3u + m2 (left)) * 3u * 8u + m2 (above) * 8u;
        const uint32_t acb = P_AC8 + ((stype * 8u + (uint32_t)tcd) * 64u) * 125u * 16u;
        md_block (c, lev + 64 * q, cnt + idx, 64, emit_ldc, 0, nzp, acb);
      } else {
        for (int k = 0; k < 4; k++) {                   // getNonzerosPrior4x4 macroblock_model.cpp:478-527
          const int idx = q * 4 + k;
          int past = 0, left = 0, above = 0;
          if (P) past = P->cnt[idx];
          if ((idx & 3) == 0) { if (L) left = L->cnt[idx + 3]; } else left = cnt[idx - 1];
          if (idx < 4) { if (A) above = A->cnt[idx + 12]; } else above = cnt[idx - 4];
          const uint32_t nzp = P_NZ + ((((stype * NT + (uint32_t)tcd) * 3u + 0u) * 3u + m2 (past)) * 3u + m2 (left)) * 3u * 8u + m2 (above) * 8u;
          const uint32_t acb = P_AC + (((stype * 8u + (uint32_t) (tcd < 7 ? tcd : 7)) * 3u + 0u) * 16u) * 125u * 16u;
          md_block (c, lev + 16 * idx, cnt + idx, 16, emit_ldc, 0, nzp, acb);
        }
      }
    }
    if (st.error) break;
    if (cbpC == 2)
      for (int i = 0; i < 8; i++) {
        const uint32_t color = i < 4 ? 1u : 2u;
        int past = 0, left = 0, above = 0;
        if (P) past = P->cnt[16 + i];
        if ((i & 1) == 0) { if (L) left = L->cnt[16 + i + 1]; } else left = cnt[16 + i - 1];
        if ((i & 2) == 0) { if (A) above = A->cnt[16 + i + 2]; } else above = cnt[16 + i - 2];
        const uint32_t nzp = P_NZ + ((((stype * NT + (uint32_t)tcd) * 3u + color) * 3u + m2 (past)) * 3u + m2 (left)) * 3u * 8u + m2 (above) * 8u;
        const uint32_t acb = P_AC + (((stype * 8u + (uint32_t) (tcd < 7 ? tcd : 7)) * 3u + color) * 16u) * 125u * 16u;
        md_block (c, lev + 16 * (16 + i), cnt + 16 + i, 16, false, (int)color, nzp, acb);
      }
    // ---- FreqImage store (:3107) ----
    {
      CtxRec& o = cur[mb_addr];
      o.flags = 1; o.type_code = (uint8_t)tcd; o.cbp = r.cbp; o.num_ref = (uint8_t)num_ref; o.chroma_mode = r.chroma_mode; o.i16_mode = r.i16_mode;
      o.cachedSkips = 0;
      for (int i = 0; i < 24; i++) o.cnt[i] = cnt[i];
    }
    // ---- CAVLC ----
    if (tcd >= 8) { st.error = MD_ERR_PCM; break; }
    {
      const uint8_t* nbL = hasLs ? &st.tcoef[(size_t) (mb_addr - 1) * MD_NB_STRIDE] : nullptr;
      const uint8_t* nbT = hasTs ? &st.tcoef[(size_t) (mb_addr - W) * MD_NB_STRIDE] : nullptr;
      if (!(cabac ? cb_mb (bw, sc, cs, tc, nbL, nbT) : cw_mb (bw, sc, cs, tc, nbL, nbT))) { st.error = MD_ERR_SYNTAX; break; }
    }
    if (bw.overflow) { st.error = MD_ERR_OUT_FULL; break; }
    st.mb_intra[mb_addr] = tcd <= 2; st.mb_slice[mb_addr] = slice_no;
    mb_addr++; mb_in_slice++;
  }
  // slice trailing alignment bits (decode_slice.cpp:3133-3147): coded through TEST_PROB into the PADBYTE stream
  if (cabac) {
    // the reference stores the low seven bits of the slice's last byte (decoder.cpp parse_slice_data_cabac); the arithmetic
    // encoder's flush ended with the stop bit, which the caller's trailer writes again: hand it back
    const uint32_t pad7 = st.error ? 0u : md_raw (c, CT_PAD, 7);
    if (!stop && !st.error) st.error = MD_ERR_SYNTAX;  // a CABAC slice ends with end_of_slice_flag = 1
    bw.trim = stop ? 1 : 0;
    const int n_pad = 7 - (int) ((sl.header_bits + mb_bits (bw) - (uint64_t)bw.trim) & 7u);
    *n_pad_out = (uint8_t)n_pad;
    *pad_bits_out = (uint8_t) (pad7 & ((1u << n_pad) - 1u));
    return mb_addr - (int)sl.first_mb;
  }
  const int n_pad = 7 - (int) ((sl.header_bits + mb_bits (bw)) & 7u);
  *n_pad_out = (uint8_t)n_pad;
  *pad_bits_out = (uint8_t) (n_pad && !st.error ? md_raw (c, CT_PAD, n_pad) : 0u);
  return mb_addr - (int)sl.first_mb;
}

} // namespace lh264md
