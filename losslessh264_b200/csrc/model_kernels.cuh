// model_kernels.cuh — the recompression model ("Pied Piper" layer) of losslessh264 on sm_100a,
// compress direction.  Replaces, for a batch of pictures of independent streams, the per-macroblock
// modelling of WelsDecodeSlice / WelsDecodeSliceForNonRecoding (codec/decoder/core/src/decode_slice.cpp
// :2959-3171, :2174-2474), the prior selection of MacroblockModel (src/macroblock_model.cpp:375-645),
// the binarisations of CompressionStream (inc/compression_stream.h:441-591), DynProb adaptation
// (:87-115) and the VP8 bool coder (inc/bitwriter.h:35-105).
//
// Decomposition (B200-first, not the reference's serial per-MB loop).  The reference interleaves three
// things per binary decision: (1) binarise a symbol and pick a prior from neighbour context, (2) read and
// adapt that prior's probability, (3) push (bit, probability) through the tag's bool coder.  Only (2) and
// (3) are serial, and they are serial along different axes: (2) per prior, (3) per tag.  (1) depends on
// data alone.  So:
//
//   k_model_frame   one warp per stream: FreqImage::updateFrame (inc/decoded_macroblock.h:120-165) on the
//                   previous frame's context records, frame flip / resolution change.
//   k_model_count   one warp per macroblock record, all streams at once: context record of the MB
//                   (what the reference stores in FreqImage) + number of decisions per lane role.
//   k_model_scan    exclusive scan of the decision counts -> position of every MB in its stream's
//                   decision list.
//   k_model_emit    one warp per macroblock record: lanes = {MB header, luma DC, chroma DC, 16 luma
//                   blocks, 8 chroma blocks, slice tail}; each lane binarises its symbols and writes
//                   32-bit decision words (prior index | tag | bit) in the reference's decision order.
//   k_model_code    one warp per stream walks the stream's decision list 32 decisions at a time:
//                   lanes gather the 32 priors, resolve same-prior chains inside the chunk in order,
//                   write the adapted priors back; then lanes turn into tag owners (<= 30 live tags)
//                   and each owner pushes its tag's (bit, probability) pairs through its bool coder.
//
// Model state per stream lives in HBM: a pruned prior table of 32-bit DynProbs (the reference's 9.4 GB
// dense table, inc/macroblock_model.h:36-136, restricted to the index ranges that can occur; a fresh
// DynProb is all zero bits), two frames of 32-byte context records, 32 coder states and paged output.
#pragma once
#ifndef LH264_CPU_EMU          // tests/emu compiles these sources with g++ behind a CUDA-model shim (test infrastructure)
#include <cuda_runtime.h>
#define LH264_SPIN_HINT()
#define LH264_DYN_SHARED(name) extern __shared__ __align__ (16) unsigned char name[]
#else
#define LH264_DYN_SHARED(name) unsigned char* name = emu::dyn_shared()
#endif
#include <stdint.h>
#include "../../include/lh264_b200.h"
#include "model_layout.h"

namespace lh264m {

// ---- device structures -----------------------------------------------------------------------------------


struct Coder { uint32_t lowvalue, range; int32_t count; uint32_t pos; };   // vpx_writer, bitwriter.h:22-28

constexpr int PAGE_SHIFT = 14;                          // 16 KB output pages
constexpr uint32_t PAGE_SIZE = 1u << PAGE_SHIFT;
constexpr int MAX_PAGES_PER_TAG = 1024;                 // 16 MB per tag per stream

struct MState {                       // per model stream, device resident
  uint32_t* priors;
  CtxRec* frame[2];
  CtxRec* newrec;                     // records of the picture being coded, committed to frame[cur] at its end
  int* order;                         // slice-data position of every MB address in the picture being coded
  uint32_t* pagelist;                 // [CT_COUNT][MAX_PAGES_PER_TAG] page index + 1
  int width, height, frame_len, cur, lastFrameId, max_mbs;
  int no_past_slice0;
  int error;                          // 11-14 a context index outside the reference's prior tables, 15-16 internal, 2 out of output pages
  uint32_t used;                      // compact tags that have a stream (CompressionStream::tag creates it)
  uint32_t finished;
  unsigned long long decisions;
  Coder coder[32];
};

struct MPicDev {                      // one picture of the batch
  const lh264_model_mb* recs;
  const lh264_model_slice* slices;
  const int16_t* coefs;
  MState* st;
  uint32_t* rec_cnt;                  // [n_recs] decisions of the record
  uint32_t* rec_off;                  // [n_recs] position of the record's first decision inside the picture
  int n_recs, n_slices, mb_w, mb_h;
  int rec_base;                       // index of the picture's first record in the batch-wide record numbering
  int pad;
};

struct ArenaDev { uint8_t** page_ptrs; unsigned int* next; unsigned int n_pages; };

// Coding is DEFERRED: the generation kernels of a step (one picture per stream) append the step's decisions to a pool and leave
// one CodeJob per picture; k_model_code later codes the jobs of up to T steps in ONE launch, the block of a stream walking that
// stream's pictures in step order.  (With one launch per step the slowest picture of the step - an I picture among P pictures -
// set the step's time while every other stream's coder sat idle: streams whose I pictures fall in different steps now overlap.)
struct CodeJob { MState* st; unsigned long long base; uint32_t n; uint32_t pad; };   // st == nullptr: nothing to code
// decisions in the pool half being filled; output pages promised to the jobs of this half and of the other half (whose coder may
// still be running: the pages it has not taken yet are not in *ArenaDev::next)
struct PoolDev { unsigned long long* used; unsigned int* pages_pending; const unsigned int* pages_pending_other; };

// Written by k_model_plan, read by the host when it next touches the batch.  `skip` makes k_model_emit / k_model_code return
// without touching any state, so that the host can grow what was too small and run the two kernels again.
struct BatchCtl { unsigned long long total; unsigned int skip /* 1 too small, 2 behind a skipped batch */; unsigned int pages_needed; unsigned int pages_in_use; unsigned int pad;
                  unsigned long long pool_base /* position of the step's decisions in the pool */; };

// ---- binarisation: identical decision sequences to the reference, prior = index into the table ---------
template <bool W> struct Em {
  uint32_t* out; uint32_t n;
  __device__ __forceinline__ void put (uint32_t tag, uint32_t prior, uint32_t bit) {
    if (W) out[n] = (prior << 6) | (tag << 1) | (bit & 1u);
    n++;
  }
};
// emitBits(data, Branch<n>): compression_stream.h:477-487; pre-order tree layout of Branch<n>::selectBranch :131-139
template <bool W> __device__ __forceinline__ void e_branch (Em<W>& e, uint32_t tag, uint32_t data, int nbits, uint32_t tree) {
  for (int n = nbits; n >= 1; n--) {
    const uint32_t bit = (data >> (n - 1)) & 1u;
    e.put (tag, tree, bit);
    tree += bit ? (1u << (n - 1)) : 1u;
  }
}
// emitBits(data, nBits) through TEST_PROB: :441-448
template <bool W> __device__ __forceinline__ void e_raw (Em<W>& e, uint32_t tag, uint32_t data, int nbits) {
  for (int i = 0; i < nbits; i++) e.put (tag, P_TEST, (data >> (nbits - 1 - i)) & 1u);
}
// emitBitsZeroToPow2Inclusive<n>: :455-463
template <bool W> __device__ __forceinline__ void e_zero_to_pow2 (Em<W>& e, uint32_t tag, uint32_t data, int nbits, uint32_t priors, uint32_t preferred) {
  e.put (tag, priors, data != preferred);
  if (data != preferred) e_branch (e, tag, data > preferred ? data - 1 : data, nbits, priors + 1);
}
// emitUnary: :465-474, UnaryIntPrior<N>::at :175-177
template <bool W> __device__ __forceinline__ void e_unary (Em<W>& e, uint32_t tag, int data, uint32_t base, int n, int early) {
  for (int i = 0; i < data; i++) {
    e.put (tag, base + min (i, n - 1), 1);
    if (i == early - 1) return;
  }
  e.put (tag, base + min (data, n - 1), 0);
}
// CompressionStream::emitInt: :523-572 (mantissa size 4)
template <bool W> __device__ __forceinline__ void e_int (Em<W>& e, int data, uint32_t expo, int n_exp, uint32_t mant, int order, uint32_t zero, uint32_t sign,
                                                         uint32_t tagExp, uint32_t tagMant, uint32_t tagZero, uint32_t tagSign) {
  if (zero != NO_PRIOR) { e.put (tagZero, zero, data == 0); if (data == 0) return; }
  if (sign != NO_PRIOR) { e.put (tagSign, sign, data > 0); if (data < 0) data = -data; }
  data--;
  const int data_high = 1 + (data >> order);
  const int log2 = 31 - __clz (data_high);
  e_unary (e, tagExp, log2, expo, n_exp, -1);
  int lo = 0, hi = 4;
  for (int k = 0; k < log2 + order; k++) {
    const uint32_t bit = k < log2 ? (uint32_t) (data_high >> (log2 - 1 - k)) & 1u : (uint32_t) (data >> (order - 1 - (k - log2))) & 1u;
    if (hi > lo) {
      const int mid = (hi + lo) >> 1;
      e.put (tagMant, mant + mid, bit);
      if (bit) lo = mid + 1; else hi = mid;
    } else e.put (tagMant, P_TEST, bit);
  }
}
// CompressionStream::emitUEGkInt: :575-591; p = base of a UEGk prior with n_exp exponent slots
template <bool W> __device__ __forceinline__ void e_uegk (Em<W>& e, int data, int threshold, uint32_t p, int n_exp, int order,
                                                          uint32_t tagExp, uint32_t tagMant, uint32_t tagZero, uint32_t tagSign) {
  e.put (tagZero, p + U_ZERO, data == 0);
  if (data == 0) return;
  e.put (tagSign, p + U_SIGN, data < 0);
  if (data < 0) data = -data;
  e_unary (e, tagMant, data - 1, p + U_FIRST, 4, threshold);
  if (data - 1 >= threshold)
    e_int (e, data - 1 - threshold, p + U_EXP, n_exp, p + U_EXP + n_exp, order, p + U_EXP + n_exp + 4, NO_PRIOR, tagExp, tagMant, tagZero, tagSign);
}

__constant__ uint8_t c_zz16[16] = { 0, 1, 4, 8, 5, 2, 3, 6, 9, 12, 13, 10, 7, 11, 14, 15 };     // decode_slice.cpp:2035-2040
__constant__ uint8_t c_zz64[64] = {                                                              // decode_slice.cpp:2041-2050
  0, 1, 5, 6, 14, 15, 27, 28, 2, 4, 7, 13, 16, 26, 29, 42, 3, 8, 12, 17, 25, 30, 41, 43, 9, 11, 18, 24, 31, 40, 44, 53,
  10, 19, 23, 32, 39, 45, 52, 54, 20, 22, 33, 38, 46, 51, 55, 60, 21, 34, 37, 47, 50, 56, 59, 61, 35, 36, 48, 49, 57, 58, 62, 63 };
__constant__ uint8_t c_scan4[16] = { 0, 1, 4, 5, 2, 3, 6, 7, 8, 9, 12, 13, 10, 11, 14, 15 };    // g_kuiScan4

// encode4x4<n>: decode_slice.cpp:2060-2094 with getACPrior4x4/8x8 (macroblock_model.cpp:562-604; the neighbour terms of
// the index are the constant 2 because priorCoef only fills `past`, :196-262).  nzp = nonzeros prior base.
// cnt_all = non-zero levels of the block including position 0 (from the chunk counts); ac may be nullptr when it is 0.
template <bool W> __device__ void gen_block (Em<W>& e, const int16_t* ac, int cnt_all, int n, bool emit_dc, int color, uint32_t nzp, uint32_t acbase) {
  const uint8_t* zz = n == 16 ? c_zz16 : c_zz64;
  const int nonzeros = cnt_all - ((!emit_dc && cnt_all && ac[0] != 0) ? 1 : 0);
  const uint32_t tEob = color ? CT_C_EOB : CT_L0_EOB;
  e_int (e, nonzeros, nzp + I_EXP, 3, nzp + I_MANT, 0, nzp + I_ZERO, NO_PRIOR, tEob, tEob, tEob, tEob);
  int left = nonzeros, emitted = 0, prev = 0, prev2 = 0;
  for (int i = emit_dc ? 0 : 1; i < n && left > 0; i++) {
    const int c = ac[zz[i]];
    const uint32_t tb = color ? CT_C_BITMASK : (i == 0 ? CT_L0_BITMASK : CT_LN_BITMASK);
    const uint32_t p = acbase + ((uint32_t) (emitted * 5 + min (4, left)) * 25u + (uint32_t)min (4, max (0, prev + 2)) * 5u + (uint32_t)min (4, max (0, prev2 + 2))) * 16u;
    e_uegk (e, c, 14, p, 2, 0, tb + 1, tb + 2, tb, tb + 3);
    prev2 = prev; prev = c; emitted++;
    if (c) left--;
  }
}
// which EXP tag streams a block creates without necessarily coding into them (tag().billTo(), decode_slice.cpp:2081-2083)
__device__ __forceinline__ uint32_t block_touch (const int16_t* ac, int cnt_all, bool emit_dc, int color) {
  const int nonzeros = cnt_all - ((!emit_dc && cnt_all && ac[0] != 0) ? 1 : 0);
  if (!nonzeros) return 0;
  if (color) return 1u << CT_C_EXP;
  uint32_t m = emit_dc ? 1u << CT_L0_EXP : 0;
  // a second coded position exists unless the only non-zero level is the first one coded
  if (!emit_dc || nonzeros > 1 || ac[0] == 0) m |= 1u << CT_LN_EXP;
  return m;
}

struct CtxRecS {                      // CtxRec without the 16-byte alignment, for the shared-memory staging
  uint8_t flags, type_code, cbp, num_ref, chroma_mode, i16_mode;
  uint16_t cachedSkips;
  uint8_t cnt[24];
};
struct WarpMb {                       // staging of one macroblock in shared memory
  lh264_model_mb rec;                 // 128 B
  CtxRecS L, A, P;
  uint32_t flagsLAP;                  // bit0 L, bit1 A, bit2 P present
};
// 16-coefficient chunk b of the macroblock in the picture's coefficient pool (present chunks are consecutive, ordered by
// bit index of coded_mask); nullptr = not coded = all zero.  The levels are read in place: a chunk is one 32-byte sector
// that stays in L1 while its thread walks it.
__device__ __forceinline__ const int16_t* chunk_ptr (const lh264_model_mb& r, const int16_t* pool, int b) {
  const uint32_t mask = (r.flags & LH264_MM_WRITE_BLOCK) ? r.coded_mask : 0u;
  if (!((mask >> b) & 1u)) return nullptr;
  return pool + ((size_t)r.coef_off + __popc (mask & ((1u << b) - 1u))) * 16;
}

__device__ __forceinline__ int chunk_count (const int16_t* a) {
  int n = 0;
#pragma unroll
  for (int i = 0; i < 16; i++) n += a[i] != 0;
  return n;
}

// lane roles of a macroblock warp, in the reference's emission order
enum { ROLE_HEADER = 0, ROLE_LDC = 1, ROLE_CDC = 2, ROLE_LUMA0 = 3, ROLE_CHROMA0 = 19, ROLE_TAIL = 27 };

// Decisions of one lane role.  Returns through e; touch = EXP tags created.
template <bool W> __device__ void gen_lane (Em<W>& e, const WarpMb& s, const int16_t* pool, const lh264_model_slice& sl, int lane, bool last_of_slice,
                                            const uint8_t* own_cnt, uint32_t* touch, int* error) {
  const lh264_model_mb& r = s.rec;
  const int tc = r.type_code;
  const uint32_t st = sl.slice_type == 0 ? 0u : 1u;
  const bool write = r.flags & LH264_MM_WRITE_BLOCK;
  const bool hasL = s.flagsLAP & 1, hasA = s.flagsLAP & 2, hasP = s.flagsLAP & 4;
  const int cbpL = r.cbp & 15, cbpC = r.cbp >> 4;
  if (lane == ROLE_HEADER) {
    const uint32_t stop_idx = P_STOP + min ((int)r.mb_in_slice, 2047);
    if (r.flags & LH264_MM_EMIT_SKIP_RUN) {         // decode_slice.cpp:2210-2217; getSkipRunPrior macroblock_model.cpp:375-390
      const int cs = hasP ? s.P.cachedSkips : 0;
      const int prior = cs / 8 + (cs % 8 ? 1 : 0);
      if (prior >= 512) { *error = 11; return; }
      e_branch (e, CT_SKIP, r.skip_run, 9, P_SKIP + (uint32_t)prior * 512u);
    }
    if (r.flags & LH264_MM_STOP_IN_SKIP) e.put (CT_SKIP_END, stop_idx, (r.flags & LH264_MM_STOP_BIT) != 0);
    if (!write) return;
    e.put (CT_SKIP_END, stop_idx, (r.flags & LH264_MM_STOP_BIT) != 0);                     // :2252
    {                                                 // getMacroblockTypePrior macroblock_model.cpp:415-438
      int prior = 15, prevType = 15;
      if (hasA) prior = s.A.type_code;
      if (hasL) prior = s.L.type_code;
      if (hasP) prevType = s.P.type_code;
      e_branch (e, CT_MB_TYPE, (uint32_t)tc, 4, P_MBTYPE + ((uint32_t) (prior + prevType) * 2u + (st == 0)) * 16u);
    }
    if (sl.chroma_format_idc != 0) e_branch (e, CT_CBPL, (uint32_t)cbpC, 2, P_CBPC + ((uint32_t) (hasP ? s.P.cbp >> 4 : 0) * NT + tc) * 4u);
    e_branch (e, CT_CBPL, (uint32_t)cbpL, 4, P_CBPL + ((uint32_t) (hasP ? s.P.cbp & 15 : 0) * NT + tc) * 16u);
    {                                                 // :2267-2270, swizzle_sign macroblock_model.cpp:715-721
      const int d = r.qp_delta;
      const uint32_t sw = d >= 0 ? (uint32_t)d << 1 : (((uint32_t) (-d - 1)) << 1) | 1u;
      e_zero_to_pow2 (e, CT_QPL, sw, 7, P_QPL + ((r.mb_in_slice == 0 ? 3u : 0u) + ((r.flags >> LH264_MM_LASTDQP_SHIFT) & 3u)) * 128u, 0);
    }
    const int num_ref = sl.num_ref_idx_l0;
    {
      const int pr = hasP ? s.P.num_ref : 0;
      if (pr >= 16) { *error = 12; return; }
      e_branch (e, CT_REF, (uint32_t)num_ref, 4, P_NUMREF + ((uint32_t)pr * NT + tc) * 16u);
    }
    int refBits = 0;
    while ((1 << refBits) < num_ref) refBits++;
    {
      const uint32_t pc = hasP ? min ((int)s.P.chroma_mode, 6) : 7;
      e_zero_to_pow2 (e, CT_8x8, r.chroma_mode, 3, P_CMODE + pc * 8u, pc);
      const uint32_t pl = hasP ? min ((int)s.P.i16_mode, 6) : 7;       // the CHROMA table: macroblock_model.cpp:632
      e_zero_to_pow2 (e, CT_16x16, r.i16_mode, 3, P_CMODE + pl * 8u, pl);
    }
    if (tc == 0)
      for (int i = 0; i < 16; i++)
        e_branch (e, CT_PRED_MODE, r.pred_mode[i] & 15u, 4, P_PREDMODE + (((uint32_t)tc * 8u + (r.pred_avail[i] & 7u)) * 9u + min (r.pred_mode[i] >> 4, 8)) * 16u);
    if (tc == 2) {
      for (int i = 0; i < 4; i++) e_branch (e, CT_PRED_MODE, r.pred_mode[i] & 15u, 4, P_PREDMODE + (((uint32_t)tc * 8u + 6u) * 9u + 1u) * 16u);
      for (int i = 0; i < 4; i++) e_branch (e, CT_SUB_MB, r.sub_mb_type[i], 8, P_SUBMB + (uint32_t)tc * 256u);
      for (int i = 0; i < 4; i++) e_raw (e, CT_REF, (uint8_t)r.ref_idx[i], refBits);
    } else if (tc == 6 || tc == 7) {
      for (int i = 0; i < 4; i++) e_branch (e, CT_SUB_MB, r.sub_mb_type[i], 8, P_SUBMB + (uint32_t)tc * 256u);
      if (tc == 6) for (int i = 0; i < 4; i++) e_raw (e, CT_REF, (uint8_t)r.ref_idx[i], refBits);
      for (int i = 0; i < 4; i++) {
        const int sub = r.sub_mb_type[i];
        const int cnt = sub == 1 ? 1 : (sub == 8 ? 4 : 2), step = sub == 2 ? 2 : 1;
        if (sub != 1 && sub != 2 && sub != 4 && sub != 8) { *error = 13; return; }
        for (int j = 0; j < cnt; j++) {
          const int b = c_scan4[(i << 2) + j * step];
          const uint32_t p = P_MVD + ((uint32_t) (tc - 3) * 16u + b) * 16u;
          e_uegk (e, r.mvd[b][0], 9, p, 3, 3, CT_MVX, CT_MVX, CT_MVX, CT_MVX);
          e_uegk (e, r.mvd[b][1], 9, p, 3, 3, CT_MVY, CT_MVY, CT_MVY, CT_MVY);
        }
      }
    } else if (tc == 4 || tc == 5) {
      for (int i = 0; i < 2; i++) e_raw (e, CT_REF, (uint32_t) (int)r.ref_idx[i], refBits);
      for (int i = 0; i < 2; i++) {
        const int b = tc == 4 ? i * 8 : i * 2;
        const uint32_t p = P_MVD + ((uint32_t) (tc - 3) * 16u + b) * 16u;
        e_uegk (e, r.mvd[b][0], 9, p, 3, 3, CT_MVX, CT_MVX, CT_MVX, CT_MVX);
        e_uegk (e, r.mvd[b][1], 9, p, 3, 3, CT_MVY, CT_MVY, CT_MVY, CT_MVY);
      }
    } else if (tc == 3) {
      e_raw (e, CT_REF, (uint32_t) (int)r.ref_idx[0], refBits);
      const uint32_t p = P_MVD;
      e_uegk (e, r.mvd[0][0], 9, p, 3, 3, CT_MVX, CT_MVX, CT_MVX, CT_MVX);
      e_uegk (e, r.mvd[0][1], 9, p, 3, 3, CT_MVY, CT_MVY, CT_MVY, CT_MVY);
    }
    if (r.flags & LH264_MM_T8_CODED) e.put (CT_T8, P_T8 + (uint32_t)tc * 64u + min ((int)r.luma_qp, 63), (r.flags & LH264_MM_T8) != 0);
    return;
  }
  if (lane == ROLE_TAIL) {                            // slice trailing bits, decode_slice.cpp:3133-3147
    if (last_of_slice && sl.n_pad_bits) e_raw (e, CT_PAD, sl.pad_bits, sl.n_pad_bits);
    return;
  }
  if (!write) return;
  if (lane == ROLE_LDC) {                             // :2396-2402
    if (tc != 1) return;
    for (int i = 0; i < 16; i++) {
      const uint32_t p = P_LDC + (((uint32_t)i * 2u + st) * NT + tc) * 16u;
      const int16_t* dcp = chunk_ptr (r, pool, i);
      e_int (e, dcp ? dcp[0] : 0, p + I_EXP, 3, p + I_MANT, 0, p + I_ZERO, p + I_SIGN, CT_LDC, CT_LDC, CT_LDC, CT_LDC);
    }
    return;
  }
  if (lane == ROLE_CDC) {                             // :2403-2409
    if (cbpC != 1 && cbpC != 2) return;
    for (int i = 0; i < 8; i++) {
      const uint32_t p = P_CDC + (((uint32_t)i * 2u + st) * NT + tc) * 16u;
      const int16_t* dcp = chunk_ptr (r, pool, 16 + i);
      e_int (e, dcp ? dcp[0] : 0, p + I_EXP, 3, p + I_MANT, 0, p + I_ZERO, p + I_SIGN, CT_CRDC, CT_CRDC, CT_CRDC, CT_CRDC);
    }
    return;
  }
  const bool t8 = r.flags & LH264_MM_T8;
  if (lane >= ROLE_LUMA0 && lane < ROLE_CHROMA0) {    // :2410-2426
    const int idx = lane - ROLE_LUMA0;
    if (!(cbpL & (1 << (idx >> 2)))) return;
    if (tc > 7 && !t8) {                              // I_PCM with retained cbp: all levels are zero, only the non-zero count is coded
    }
    const bool emit_dc = tc != 1;
    if (t8) {
      if (idx & 3) return;
      // getNonzerosPrior8x8 macroblock_model.cpp:529-560
      const int q = idx >> 2;
      int past = 0, left = 0, above = 0;
      if (hasP) past = s.P.cnt[idx] + s.P.cnt[idx + 1] + s.P.cnt[idx + 2] + s.P.cnt[idx + 3];
      if ((q & 1) == 0) { if (hasL) { const int k = (q + 1) * 4; left = s.L.cnt[k] + s.L.cnt[k + 1] + s.L.cnt[k + 2] + s.L.cnt[k + 3]; } }
      else { const int k = (q - 1) * 4; left = own_cnt[k] + own_cnt[k + 1] + own_cnt[k + 2] + own_cnt[k + 3]; }
      if ((q & 2) == 0) { if (hasA) { const int k = (q + 2) * 4; above = s.A.cnt[k] + s.A.cnt[k + 1] + s.A.cnt[k + 2] + s.A.cnt[k + 3]; } }
      else { const int k = (q - 2) * 4; above = own_cnt[k] + own_cnt[k + 1] + own_cnt[k + 2] + own_cnt[k + 3]; }
      const uint32_t nzp = P_NZ8 + (((st * NT + tc) * 3u + min (2, past)) * 3u + min (2, left)) * 3u * 8u + (uint32_t)min (2, above) * 8u;
      if (tc > 7) { *error = 14; return; }
      const uint32_t acb = P_AC8 + ((st * 8u + tc) * 64u) * 125u * 16u;
      const int16_t* ac = chunk_ptr (r, pool, idx);
      const int call = own_cnt[idx] + own_cnt[idx + 1] + own_cnt[idx + 2] + own_cnt[idx + 3];
      if (touch) *touch |= block_touch (ac, call, emit_dc, 0);
      gen_block (e, ac, call, 64, emit_dc, 0, nzp, acb);
    } else {
      // getNonzerosPrior4x4 macroblock_model.cpp:478-527 (colour 0)
      int past = 0, left = 0, above = 0;
      if (hasP) past = s.P.cnt[idx];
      if ((idx & 3) == 0) { if (hasL) left = s.L.cnt[idx + 3]; } else left = own_cnt[idx - 1];
      if (idx < 4) { if (hasA) above = s.A.cnt[idx + 12]; } else above = own_cnt[idx - 4];
      const uint32_t nzp = P_NZ + ((((st * NT + tc) * 3u + 0u) * 3u + min (2, past)) * 3u + min (2, left)) * 3u * 8u + (uint32_t)min (2, above) * 8u;
      const uint32_t acb = P_AC + (((st * 8u + min (tc, 7)) * 3u + 0u) * 16u) * 125u * 16u;
      const int16_t* ac = chunk_ptr (r, pool, idx);
      if (touch) *touch |= block_touch (ac, own_cnt[idx], emit_dc, 0);
      gen_block (e, ac, own_cnt[idx], 16, emit_dc, 0, nzp, acb);
    }
    return;
  }
  if (lane >= ROLE_CHROMA0 && lane < ROLE_TAIL) {     // :2427-2435
    if (cbpC != 2) return;
    const int i = lane - ROLE_CHROMA0;                // 0..7: Cb 0-3, Cr 4-7
    const int color = i < 4 ? 1 : 2;
    const bool emit_dc = !(cbpC == 1 || cbpC == 2);
    int past = 0, left = 0, above = 0;
    if (hasP) past = s.P.cnt[16 + i];
    if ((i & 1) == 0) { if (hasL) left = s.L.cnt[16 + i + 1]; } else left = own_cnt[16 + i - 1];
    if ((i & 2) == 0) { if (hasA) above = s.A.cnt[16 + i + 2]; } else above = own_cnt[16 + i - 2];
    const uint32_t nzp = P_NZ + ((((st * NT + tc) * 3u + color) * 3u + min (2, past)) * 3u + min (2, left)) * 3u * 8u + (uint32_t)min (2, above) * 8u;
    const uint32_t acb = P_AC + (((st * 8u + min (tc, 7)) * 3u + color) * 16u) * 125u * 16u;
    const int16_t* ac = chunk_ptr (r, pool, 16 + i);
    if (touch) *touch |= block_touch (ac, own_cnt[16 + i], emit_dc, color);
    gen_block (e, ac, own_cnt[16 + i], 16, emit_dc, color, nzp, acb);
    return;
  }
}

// ---- k_model_frame: per-stream frame bookkeeping -------------------------------------------------------------
// WelsDecodeSlice decode_slice.cpp:3032-3046 for the first slice of the picture (later slices of the same picture repeat
// the same updateFrame on unchanged data) and FreqImage::updateFrame decoded_macroblock.h:120-165.
__global__ void __launch_bounds__ (32) k_model_frame (const MPicDev* pics, int n_pics, const unsigned int* poison) {
  const int p = blockIdx.x, lane = threadIdx.x;
  if (p >= n_pics || *poison) return;
  const MPicDev& pic = pics[p];
  MState* st = pic.st;
  int cur = st->cur, len = st->frame_len;
  const int frame = pic.slices[0].frame_num;
  if (frame != st->lastFrameId) cur ^= 1;
  CtxRec* f = st->frame[1 - cur];
  // isSkipped / cachedSkips of the other frame: run of all-zero MBs followed by a non-zero one gets its length written into
  // the LAST `run` records ending at the non-zero one (the reference's off-by-one is kept)
  int run_start = 0;                                 // first index of the current zero run
  for (int base = 0; base < len; base += 32) {
    const int i = base + lane;
    bool zeroed = false, valid = i < len;
    if (valid) {
      const uint4* q = reinterpret_cast<const uint4*> (&f[i]);
      const uint4 a = q[0], b = q[1];
      zeroed = ((a.z | a.w | b.x | b.y | b.z | b.w) == 0);
      uint8_t fl = (uint8_t) (a.x & 0xff);
      fl = zeroed ? (fl | 2) : (fl & ~2);
      f[i].flags = fl;
    }
    unsigned nzmask = __ballot_sync (0xffffffffu, valid && !zeroed);
    while (nzmask) {
      const int l = __ffs (nzmask) - 1;
      nzmask &= nzmask - 1;
      const int at = base + l, run = at - run_start;
      if (run > 0) for (int j = at - run + 1 + lane; j <= at; j += 32) f[j].cachedSkips = (uint16_t)run;
      run_start = at + 1;
    }
    __syncwarp();
  }
  if (lane == 0) {
    st->cur = cur; st->lastFrameId = frame;
    st->no_past_slice0 = 0;
  }
  if (st->width != pic.mb_w || st->height != pic.mb_h) {     // std::vector::resize semantics, decode_slice.cpp:3036-3045
    const int n = pic.mb_w * pic.mb_h;
    for (int k = 0; k < 2; k++)
      for (int j = len + lane; j < n; j += 32) { uint4* q = reinterpret_cast<uint4*> (&st->frame[k][j]); q[0] = make_uint4 (0, 0, 0, 0); q[1] = make_uint4 (0, 0, 0, 0); }
    __syncwarp();
    if (lane == 0) { st->no_past_slice0 = 1; st->width = pic.mb_w; st->height = pic.mb_h; st->frame_len = n; }
  }
  for (int j = lane; j < pic.mb_w * pic.mb_h; j += 32) st->order[j] = 0x7fffffff;
}

// ---- k_model_count / k_model_emit: ONE THREAD per macroblock record ---------------------------------------------------
// A typical inter macroblock produces ~25 header decisions and a handful of block decisions; a warp per macroblock left
// 31 lanes idle most of the time.  Here every thread owns one record: it stages the record and its three context records
// in a private slice of shared memory (stride 63 words: odd, so equal-field accesses of the 32 lanes hit 32 different
// banks), reads coefficient chunks in place, and walks the lane roles of gen_lane() in the reference's emission order.
constexpr int GEN_THREADS = 256;
struct ThreadMb { WarpMb m; uint8_t own[24]; };
static_assert (sizeof (ThreadMb) == 252, "ThreadMb stride must stay an odd number of words");

// slices of a picture are in arrival order, first_rec increasing
__device__ __forceinline__ int find_slice (const MPicDev& pic, int rec) {
  int lo = 0, hi = pic.n_slices - 1;
  while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (pic.slices[mid].first_rec <= rec) lo = mid; else hi = mid - 1; }
  const lh264_model_slice& sl = pic.slices[lo];
  return rec >= sl.first_rec && rec < sl.first_rec + sl.n_recs ? lo : -1;
}
// batch-wide record numbering -> (picture, record): binary search over MPicDev::rec_base
__device__ __forceinline__ int find_pic (const MPicDev* pics, int n_pics, int g) {
  int lo = 0, hi = n_pics - 1;
  while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (pics[mid].rec_base <= g) lo = mid; else hi = mid - 1; }
  return lo;
}

// Which picture / record every thread of the block owns.  The block's records are consecutive in the batch-wide numbering,
// so the binary search runs once per block (thread 0) and the other threads only step over picture boundaries.
__device__ __forceinline__ bool block_locate (const MPicDev* pics, int n_pics, int total_recs, int* p_out, int* ri_out) {
  __shared__ int s_first_pic;
  const int g = blockIdx.x * GEN_THREADS + threadIdx.x;
  if (threadIdx.x == 0) s_first_pic = find_pic (pics, n_pics, blockIdx.x * GEN_THREADS);
  __syncthreads();
  if (g >= total_recs) return false;
  int p = s_first_pic;
  while (p + 1 < n_pics && pics[p + 1].rec_base <= g) p++;
  *p_out = p; *ri_out = g - pics[p].rec_base;
  return true;
}
// The warp copies its 32 records (4 KB, contiguous unless a picture boundary intervenes) with coalesced loads.
__device__ __forceinline__ void warp_stage_recs (ThreadMb* tm, const lh264_model_mb* my_rec, bool active) {
  const int lane = threadIdx.x & 31, w0 = threadIdx.x & ~31;
  for (int i = 0; i < 32; i++) {
    const unsigned long long a = __shfl_sync (0xffffffffu, (unsigned long long)my_rec, i);
    const int on = __shfl_sync (0xffffffffu, (int)active, i);
    if (on) reinterpret_cast<uint32_t*> (&tm[w0 + i].m.rec)[lane] = __ldg (reinterpret_cast<const uint32_t*> (a) + lane);
  }
  __syncwarp();
}
// non-zero levels per 16-coefficient chunk (DC slot included): what the neighbour contexts and FreqImage::updateFrame need
__device__ __forceinline__ void t_count_chunks (ThreadMb& t, const int16_t* coefs) {
  const uint32_t mask = (t.m.rec.flags & LH264_MM_WRITE_BLOCK) ? t.m.rec.coded_mask : 0u;
  const uint4* pool = reinterpret_cast<const uint4*> (coefs) + (size_t)t.m.rec.coef_off * 2;
  int k = 0;
  for (int b = 0; b < 24; b++) {
    int n = 0;
    if ((mask >> b) & 1u) {
      const uint4 x = __ldg (pool + 2 * k), y = __ldg (pool + 2 * k + 1);
      k++;
      const uint32_t w[8] = { x.x, x.y, x.z, x.w, y.x, y.y, y.z, y.w };
#pragma unroll
      for (int i = 0; i < 8; i++) n += ((w[i] & 0xffffu) != 0) + ((w[i] >> 16) != 0);
    }
    t.own[b] = (uint8_t)n;
  }
}
__device__ __forceinline__ void t_load_ctx (CtxRecS& dst, const CtxRec* src) {
  uint32_t* d = reinterpret_cast<uint32_t*> (&dst);
  if (src) {
    const uint4 a = reinterpret_cast<const uint4*> (src)[0], b = reinterpret_cast<const uint4*> (src)[1];
    d[0] = a.x; d[1] = a.y; d[2] = a.z; d[3] = a.w; d[4] = b.x; d[5] = b.y; d[6] = b.z; d[7] = b.w;
  } else {
#pragma unroll
    for (int i = 0; i < 8; i++) d[i] = 0;
  }
}

__global__ void __launch_bounds__ (GEN_THREADS) k_model_count (const MPicDev* pics, int n_pics, int total_recs, const unsigned int* poison) {
  LH264_DYN_SHARED (gen_smem);
  if (*poison) return;
  ThreadMb* tm = reinterpret_cast<ThreadMb*> (gen_smem);
  ThreadMb& t = tm[threadIdx.x];
  int p = 0, ri = 0;
  const bool active = block_locate (pics, n_pics, total_recs, &p, &ri);
  warp_stage_recs (tm, active ? &pics[p].recs[ri] : nullptr, active);
  if (!active) return;
  const MPicDev& pic = pics[p];
  t_count_chunks (t, pic.coefs);
  const int sidx = find_slice (pic, ri);
  if (sidx < 0) { pic.rec_cnt[ri] = 0; return; }
  const lh264_model_slice& sl = pic.slices[sidx];
  MState* st = pic.st;
  const lh264_model_mb& r = t.m.rec;
  const bool write = r.flags & LH264_MM_WRITE_BLOCK;
  const bool in_range = (int)r.mb_addr < pic.mb_w * pic.mb_h;
  if (!in_range) { pic.rec_cnt[ri] = 0; return; }
  // context record of this MB: a coded MB stores its own data, a skipped one copies the co-located record of the other frame
  // (decode_slice.cpp:3098-3109)
  const CtxRec* past = &st->frame[1 - st->cur][r.mb_addr];
  {
    uint4* dst = reinterpret_cast<uint4*> (&st->newrec[r.mb_addr]);
    if (write) {
      CtxRec n;
      n.flags = 1; n.type_code = r.type_code; n.cbp = r.cbp; n.num_ref = sl.num_ref_idx_l0;
      n.chroma_mode = r.chroma_mode; n.i16_mode = r.i16_mode; n.cachedSkips = 0;
#pragma unroll
      for (int i = 0; i < 24; i++) n.cnt[i] = t.own[i];
      dst[0] = reinterpret_cast<const uint4*> (&n)[0]; dst[1] = reinterpret_cast<const uint4*> (&n)[1];
    } else {
      dst[0] = reinterpret_cast<const uint4*> (past)[0]; dst[1] = reinterpret_cast<const uint4*> (past)[1];
    }
    st->order[r.mb_addr] = ri;
  }
  // The binarisation is context free except for the two intra mode symbols, whose "preferred" value comes from the
  // co-located record of the other frame (PAST), so only that neighbour is loaded for counting.
  t_load_ctx (t.m.L, nullptr); t_load_ctx (t.m.A, nullptr);
  t_load_ctx (t.m.P, (st->no_past_slice0 && sidx == 0) ? nullptr : past);
  t.m.flagsLAP = (t.m.P.flags & 1u) << 2;
  Em<false> e; e.out = nullptr; e.n = 0;
  int err = 0;
  uint32_t touch = 0;
  const bool last = ri == sl.first_rec + sl.n_recs - 1;
  for (int role = 0; role <= ROLE_TAIL; role++) gen_lane<false> (e, t.m, pic.coefs, sl, role, last, t.own, &touch, &err);
  pic.rec_cnt[ri] = e.n;
  if (touch) atomicOr (&st->used, touch);
  if (err) st->error = err;
}

// ---- k_model_scan: one block per picture ---------------------------------------------------------------------------------
__global__ void __launch_bounds__ (1024) k_model_scan (const MPicDev* pics, int n_pics, uint32_t* pic_total, const unsigned int* poison) {
  __shared__ uint32_t wsum[32];
  __shared__ uint32_t carry_s;
  const int p = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (p >= n_pics || *poison) return;
  const MPicDev& pic = pics[p];
  if (tid == 0) carry_s = 0;
  __syncthreads();
  for (int base = 0; base < pic.n_recs; base += 1024) {
    const int i = base + tid;
    const uint32_t v = i < pic.n_recs ? pic.rec_cnt[i] : 0u;
    uint32_t x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync (0xffffffffu, x, o); if (lane >= o) x += y; }
    if (lane == 31) wsum[warp] = x;
    __syncthreads();
    if (warp == 0) {
      uint32_t w = wsum[lane], z = w;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync (0xffffffffu, z, o); if (lane >= o) z += y; }
      wsum[lane] = z - w;
    }
    __syncthreads();
    const uint32_t excl = carry_s + wsum[warp] + x - v;
    if (i < pic.n_recs) pic.rec_off[i] = excl;
    __syncthreads();
    if (tid == 1023) carry_s = excl + v;
    __syncthreads();
  }
  if (tid == 0) pic_total[p] = carry_s;
}

// ---- k_model_plan: one block; positions of the pictures in the decision list, and the capacity checks the host used to
// make between two launches (it no longer waits in the middle of a submit) ------------------------------------------
__global__ void __launch_bounds__ (1024) k_model_plan (int n_pics, const uint32_t* pic_total, unsigned long long* dec_base, BatchCtl* ctl,
                                                       unsigned long long dec_capacity, ArenaDev ar, PoolDev pool, unsigned int* poison) {
  __shared__ unsigned long long wsum[32];
  __shared__ unsigned long long carry_s;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // an earlier batch skipped its coding: nothing of this one has run (every kernel checks), the host runs both again in order
  if (*poison) { if (tid == 0) { ctl->skip = 2; ctl->total = 0; ctl->pages_needed = 0; ctl->pages_in_use = *ar.next; ctl->pool_base = *pool.used; } return; }
  if (tid == 0) carry_s = 0;
  __syncthreads();
  for (int base = 0; base < n_pics; base += 1024) {
    const int i = base + tid;
    const unsigned long long v = i < n_pics ? pic_total[i] : 0ull;
    unsigned long long x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const unsigned long long y = __shfl_up_sync (0xffffffffu, x, o); if (lane >= o) x += y; }
    if (lane == 31) wsum[warp] = x;
    __syncthreads();
    if (warp == 0) {
      unsigned long long w = wsum[lane], z = w;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const unsigned long long y = __shfl_up_sync (0xffffffffu, z, o); if (lane >= o) z += y; }
      wsum[lane] = z - w;
    }
    __syncthreads();
    const unsigned long long excl = carry_s + wsum[warp] + x - v;
    if (i < n_pics) dec_base[i] = excl;
    __syncthreads();
    if (tid == 1023) carry_s = excl + v;
    __syncthreads();
  }
  if (tid == 0) {
    const unsigned long long total = carry_s;
    // every decision shifts out at most 7 bits, and a tag needs at most one partly filled page more than its bytes
    const unsigned long long pages = total * 7 / 8 / PAGE_SIZE + (unsigned long long)n_pics * CT_COUNT + 1;
    // pages already promised to steps that are generated but not coded yet count as in use
    const unsigned int in_use = *ar.next + *pool.pages_pending + *pool.pages_pending_other;
    const unsigned long long base = *pool.used;
    ctl->total = total; ctl->pages_needed = (unsigned int)min (pages, 0xffffffffull); ctl->pages_in_use = in_use; ctl->pool_base = base;
    ctl->skip = (base + total + 32 > dec_capacity || in_use + pages > ar.n_pages) ? 1u : 0u;
    if (ctl->skip) *poison = 1u;
    else { *pool.used = base + total; *pool.pages_pending += (unsigned int)min (pages, 0xffffffffull); }
  }
}

__global__ void __launch_bounds__ (GEN_THREADS) k_model_emit (const MPicDev* pics, int n_pics, int total_recs, uint32_t* dec, const unsigned long long* dec_base, const BatchCtl* ctl) {
  LH264_DYN_SHARED (gen_smem);
  if (ctl->skip) return;
  ThreadMb* tm = reinterpret_cast<ThreadMb*> (gen_smem);
  ThreadMb& t = tm[threadIdx.x];
  int p = 0, ri = 0;
  bool active = block_locate (pics, n_pics, total_recs, &p, &ri);
  const uint32_t cnt = active ? pics[p].rec_cnt[ri] : 0u;
  active = active && cnt != 0;
  warp_stage_recs (tm, active ? &pics[p].recs[ri] : nullptr, active);
  if (!active) return;
  const MPicDev& pic = pics[p];
  t_count_chunks (t, pic.coefs);
  const int sidx = find_slice (pic, ri);
  if (sidx < 0) return;
  const lh264_model_slice& sl = pic.slices[sidx];
  MState* st = pic.st;
  const lh264_model_mb& r = t.m.rec;
  const int mbx = (int) (r.mb_addr % (unsigned)pic.mb_w), mby = (int) (r.mb_addr / (unsigned)pic.mb_w);
  // Neighbors::init macroblock_model.cpp:9-44.  LEFT/ABOVE come from the frame being written: the record of this picture if
  // that MB precedes this one in slice-data order, else whatever the buffer still holds from two frames ago.
  {
    const CtxRec* l = nullptr; const CtxRec* a = nullptr;
    if (mbx > 0) { const int nb = (int)r.mb_addr - 1; l = st->order[nb] < ri ? &st->newrec[nb] : &st->frame[st->cur][nb]; }
    if (mby > 0) { const int nb = (int)r.mb_addr - pic.mb_w; a = st->order[nb] < ri ? &st->newrec[nb] : &st->frame[st->cur][nb]; }
    t_load_ctx (t.m.L, l); t_load_ctx (t.m.A, a);
    t_load_ctx (t.m.P, (st->no_past_slice0 && sidx == 0) ? nullptr : &st->frame[1 - st->cur][r.mb_addr]);
    t.m.flagsLAP = (t.m.L.flags & 1u) | ((t.m.A.flags & 1u) << 1) | ((t.m.P.flags & 1u) << 2);
  }
  Em<true> e;
  e.out = dec + ctl->pool_base + dec_base[p] + pic.rec_off[ri];
  e.n = 0;
  int err = 0;
  const bool last = ri == sl.first_rec + sl.n_recs - 1;
  for (int role = 0; role <= ROLE_TAIL; role++) gen_lane<true> (e, t.m, pic.coefs, sl, role, last, t.own, nullptr, &err);
  if (err) st->error = err; else if (e.n != cnt) st->error = 16;
}

// ---- k_model_code ------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t dp_prob (uint32_t w) { return (w >> 24) ^ 0x80u; }
__device__ __forceinline__ uint32_t dp_update (uint32_t w, uint32_t bit) {   // DynProb::updateProb compression_stream.h:101-110
  uint32_t c0 = w & 0x3ffu, c1 = (w >> 10) & 0x3ffu;
  if (bit) c1++; else c0++;
  // (256 * (c0 + 1)) / (c0 + c1 + 2), exactly: numerator < 2^18, so the float quotient is off by at most one
  const uint32_t num = 256u * (c0 + 1u), den = c0 + c1 + 2u;
  uint32_t prob = (uint32_t)__fdividef ((float)num, (float)den);
  const int rem = (int)num - (int) (prob * den);
  prob += rem >= (int)den ? 1u : 0u;
  prob -= rem < 0 ? 1u : 0u;
  if (c0 + c1 > 512u) { c0 = (c0 + 1u) >> 1; c1 = (c1 + 1u) >> 1; }
  return c0 | (c1 << 10) | ((prob ^ 0x80u) << 24);
}

struct OutCursor { uint8_t* page; };

__device__ __forceinline__ uint8_t* out_ptr (MState* st, const ArenaDev& ar, int tag, uint32_t pos, bool alloc) {
  uint32_t* slot = &st->pagelist[tag * MAX_PAGES_PER_TAG + (pos >> PAGE_SHIFT)];
  uint32_t pg = *slot;
  if (pg == 0) {
    if (!alloc || (pos >> PAGE_SHIFT) >= (uint32_t)MAX_PAGES_PER_TAG) return nullptr;
    const unsigned int idx = atomicAdd (ar.next, 1u);
    if (idx >= ar.n_pages) return nullptr;
    pg = idx + 1;
    *slot = pg;
  }
  return ar.page_ptrs[pg - 1] + (pos & (PAGE_SIZE - 1));
}

// vpx_write, bitwriter.h:35-105
__device__ __forceinline__ void coder_write (Coder& c, uint8_t*& page, MState* st, const ArenaDev& ar, int tag, uint32_t bit, uint32_t prob, int* error) {
  const uint32_t split = 1u + (((c.range - 1u) * prob) >> 8);
  uint32_t range = split, low = c.lowvalue;
  if (bit) { low += split; range = c.range - split; }
  int shift = __clz (range) - 24;
  range <<= shift;
  int count = c.count + shift;
  if (count >= 0) {
    const int offset = shift - count;
    if ((low << (offset - 1)) & 0x80000000u) {      // carry into the bytes already written
      int x = (int)c.pos - 1;
      while (x >= 0) {
        uint8_t* q = out_ptr (st, ar, tag, (uint32_t)x, false);
        if (!q) break;
        if (*q == 0xff) { *q = 0; x--; } else { *q += 1; break; }
      }
    }
    if ((c.pos & (PAGE_SIZE - 1)) == 0 || !page) {
      uint8_t* q = out_ptr (st, ar, tag, c.pos, true);
      if (!q) { *error = 2; return; }
      page = q - (c.pos & (PAGE_SIZE - 1));
    }
    page[c.pos & (PAGE_SIZE - 1)] = (uint8_t) (low >> (24 - offset));
    c.pos++;
    low <<= offset;
    shift = count;
    low &= 0xffffffu;
    count -= 8;
  }
  low <<= shift;
  c.count = count; c.lowvalue = low; c.range = range;
}

// One block of two warps per stream.  Warp 0 (producer) walks the decision list 32 decisions at a time: gathers the
// priors, resolves same-prior chains inside the chunk in decision order, scatters the adapted priors, and appends
// (probability, bit) to the shared-memory ring of the decision's tag.  Warp 1 (consumer): lane t owns compact tag t and
// drains ring t through its bool coder.  The two serial chains (per prior, per tag) thus run concurrently and the coder
// lanes never wait for each other.
#ifndef LH264_CODE_RING
#define LH264_CODE_RING 256
#endif
constexpr int CODE_RING = LH264_CODE_RING;    // entries per tag ring (power of two)
struct CodeShared {
  uint16_t ring[32][CODE_RING];
  volatile uint32_t prod[32];               // entries ever appended to ring t (producer writes)
  volatile uint32_t cons[32];               // entries consumed from ring t (consumer lane t writes)
  volatile int done;
  volatile int skip;
};

// ---- k_model_commit: end of a picture's generation -------------------------------------------------------------------------------
// Commits the picture's context records to the stream's current frame (decode_slice.cpp:3104-3108) - the generation kernels of the
// stream's NEXT picture read them - and leaves the picture's CodeJob for the deferred coder.  One block per picture.
__global__ void __launch_bounds__ (128) k_model_commit (const MPicDev* pics, int n_pics, const unsigned long long* dec_base, const uint32_t* pic_total,
                                                        const BatchCtl* ctl, CodeJob* jobs) {
  const int p = blockIdx.x;
  if (p >= n_pics || ctl->skip) return;
  const MPicDev& pic = pics[p];
  MState* st = pic.st;
  if (st->error) return;
  CtxRec* dst = st->frame[st->cur];
  for (int i = threadIdx.x; i < pic.n_recs; i += blockDim.x) {
    const uint32_t a = pic.recs[i].mb_addr;
    if ((int)a < pic.mb_w * pic.mb_h) {
      const uint4* q = reinterpret_cast<const uint4*> (&st->newrec[a]);
      uint4* o = reinterpret_cast<uint4*> (&dst[a]);
      o[0] = q[0]; o[1] = q[1];
    }
  }
  if (threadIdx.x == 0) { CodeJob j; j.st = st; j.base = ctl->pool_base + dec_base[p]; j.n = pic_total[p]; j.pad = 0; jobs[p] = j; }
}

// `index[b * n_steps + j]` = position of stream-block b's picture in step j's job row (jobs + j * stride), or -1.
#ifndef LH264_CODE_BOUNDS
#define LH264_CODE_BOUNDS __launch_bounds__ (64)
#endif
__global__ void LH264_CODE_BOUNDS k_model_code (const CodeJob* jobs, const int* index, int n_steps, int stride, const uint32_t* dec, ArenaDev ar) {
  __shared__ CodeShared sh;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll 1
  for (int step_j = 0; step_j < n_steps; step_j++) {
  const int ji = index[(size_t)blockIdx.x * n_steps + step_j];
  if (ji < 0) continue;
  const CodeJob job = jobs[(size_t)step_j * stride + ji];
  MState* st = job.st;
  if (!st) continue;
  __syncthreads();                                   // the previous picture's rings are drained, its coder states and error flag stored
  if (threadIdx.x < 32) { sh.prod[threadIdx.x] = 0; sh.cons[threadIdx.x] = 0; }
  if (threadIdx.x == 0) { sh.done = 0; sh.skip = st->error; }
  __syncthreads();
  if (sh.skip) continue;                             // (one read for the whole block: the two warps must agree)
  const uint32_t n = job.n;
  if (warp == 0) {
    // ---- producer ---------------------------------------------------------------------------------------------------
    uint32_t* priors = st->priors;
    const uint32_t* d = dec + job.base;
    const uint32_t lt = (1u << lane) - 1u;
    // software pipeline: decision words are loaded two chunks ahead and the priors of chunk k+1 are gathered right after
    // the scatter of chunk k
    uint32_t w1 = lane < n ? __ldg (d + lane) : 0xffffffffu;
    uint32_t w2 = 32u + lane < n ? __ldg (d + 32 + lane) : 0xffffffffu;
    uint32_t sv1 = w1 != 0xffffffffu ? priors[w1 >> 6] : 0u;
    for (uint32_t k = 0; k < n; k += 32) {
      const uint32_t w = w1;
      uint32_t sv = sv1;
      w1 = w2;
      w2 = k + 64u + lane < n ? __ldg (d + k + 64 + lane) : 0xffffffffu;
      const bool valid = w != 0xffffffffu;
      const uint32_t idx = w >> 6, tag = valid ? (w >> 1) & 31u : 31u, bit = w & 1u;
      const unsigned peers = __match_any_sync (0xffffffffu, idx);
      const unsigned bits = __ballot_sync (0xffffffffu, bit);
      // Earlier decisions of this chunk on the same prior, in decision order.  Until the counts are halved (their sum passes 512)
      // an update only increments one of them and recomputes the probability from the new counts, so a run of updates equals
      // ONE update applied to counts advanced by the run's zeros and ones (minus its last decision): hot priors, which can
      // occupy most lanes of a chunk, cost two updates per lane instead of one per predecessor.
      unsigned m = peers & lt;
      if (m) {
        const int jl = 31 - __clz (m);
        const unsigned head = m & ~(1u << jl);
        const uint32_t ones = (uint32_t)__popc (head & bits), zeros = (uint32_t)__popc (head) - ones;
        const uint32_t c0 = sv & 0x3ffu, c1 = (sv >> 10) & 0x3ffu;
        if (c0 + c1 + zeros + ones <= 512u) sv = dp_update ((sv & 0xff000000u) | (c0 + zeros) | ((c1 + ones) << 10), (bits >> jl) & 1u);
        else while (m) { const int j = __ffs (m) - 1; m &= m - 1; sv = dp_update (sv, (bits >> j) & 1u); }
      }
      const uint32_t prob = dp_prob (sv);
      sv = dp_update (sv, bit);
      if (valid && (peers >> lane) <= 1u) priors[idx] = sv;      // the last decision of the chain stores
      __syncwarp();                                                // orders the scatter before the next chunk's gather
      sv1 = w1 != 0xffffffffu ? priors[w1 >> 6] : 0u;
      // append to the tag rings in decision order
      const unsigned tpeers = __match_any_sync (0xffffffffu, tag);
      if (valid) {
        const uint32_t pos = sh.prod[tag] + (uint32_t)__popc (tpeers & lt);
        while (pos - sh.cons[tag] >= (uint32_t)CODE_RING) { LH264_SPIN_HINT(); }
        sh.ring[tag][pos & (CODE_RING - 1)] = (uint16_t) ((prob << 1) | bit);
      }
      __threadfence_block();
      __syncwarp();
      if (valid && (tpeers >> lane) <= 1u) sh.prod[tag] = sh.prod[tag] + (uint32_t)__popc (tpeers);
      __syncwarp();
    }
    __threadfence_block();
    if (lane == 0) { sh.done = 1; st->decisions += n; }
  } else {
    // ---- consumer: lane t = bool coder of compact tag t -----------------------------------------------------------------
    Coder c = st->coder[lane];
    uint8_t* page = nullptr;
    int err = 0;
    uint32_t consumed = 0, avail = 0;
    const uint16_t* myring = sh.ring[lane];
    // The loop body is the bare vpx_write dependency chain, predicated instead of branched: with one warp per scheduler
    // every taken branch costs more than the arithmetic it skips.  Only the byte-output path (once per 8 shifted bits)
    // branches.
    // one coder step; the ring entry for the NEXT step is fetched before this step's arithmetic so that the shared-memory
    // latency is off the dependency chain (vnext is only trusted when that entry had been published at fetch time)
    uint32_t vnext = 0;
    bool vnext_ok = false;
    uint32_t lastb = 0;                                 // last byte this lane wrote in this launch (carry fast path)
    bool have_last = false;
    auto step = [&] () -> bool {
      const bool has = consumed != avail;
      uint32_t v = vnext;
      if (has && !vnext_ok) v = myring[consumed & (CODE_RING - 1)];
      vnext_ok = has && (consumed + 1u != avail);
      vnext = myring[(consumed + 1u) & (CODE_RING - 1)];
      const uint32_t prob = v >> 1, bit = v & 1u;
      const uint32_t split = 1u + (((c.range - 1u) * prob) >> 8);
      uint32_t range = bit ? c.range - split : split;
      uint32_t low = c.lowvalue + (bit ? split : 0u);
      int shift = __clz (range) - 24;
      range <<= shift;
      int count = c.count + shift;
      if (has && count >= 0) {
        const int offset = shift - count;
        if ((low << (offset - 1)) & 0x80000000u) {      // carry into the bytes already written
          if (have_last && lastb != 0xffu) {            // the common case touches only the byte this lane wrote last: it is still in a
            lastb++;                                    // register and `page` is the page it went to - no loads on the coder chain
            page[(c.pos - 1u) & (PAGE_SIZE - 1)] = (uint8_t)lastb;
          } else {
            int x = (int)c.pos - 1;
            while (x >= 0) {
              uint8_t* q = out_ptr (st, ar, lane, (uint32_t)x, false);
              if (!q) break;
              if (*q == 0xff) { *q = 0; x--; } else { *q += 1; break; }
            }
          }
        }
        if (((c.pos & (PAGE_SIZE - 1)) == 0 || !page) && !err) {
          uint8_t* q = out_ptr (st, ar, lane, c.pos, true);
          if (!q) err = 2; else page = q - (c.pos & (PAGE_SIZE - 1));
        }
        lastb = (low >> (24 - offset)) & 0xffu;
        have_last = !err;
        if (!err) page[c.pos & (PAGE_SIZE - 1)] = (uint8_t)lastb;
        c.pos++;
        low <<= offset;
        shift = count;
        low &= 0xffffffu;
        count -= 8;
      }
      low <<= shift;
      if (has) { c.range = range; c.lowvalue = low; c.count = count; consumed++; }
      return has;
    };
    while (true) {
      const bool h0 = step();
      const bool h1 = step();
      sh.cons[lane] = consumed;
      if (!h1) { avail = sh.prod[lane]; vnext_ok = false; }
      if (!__any_sync (0xffffffffu, h0 || h1)) {
        const int dn = sh.done;
        avail = sh.prod[lane];
        vnext_ok = false;
        if (dn && __all_sync (0xffffffffu, avail == consumed)) break;
      }
    }
    st->coder[lane] = c;
    const unsigned usedm = __ballot_sync (0xffffffffu, consumed != 0);
    const unsigned errm = __ballot_sync (0xffffffffu, err != 0);
    if (lane == 0) {
      if (usedm) atomicOr (&st->used, usedm);
      if (errm) st->error = 2;
    }
  }
  }                                                  // next picture of this stream
}

// vpx_stop_encode (bitwriter.cpp:28-38) for every tag stream that exists; one warp per stream
__global__ void __launch_bounds__ (32) k_model_finish (MState* st, ArenaDev ar) {
  const int lane = threadIdx.x;
  if (st->finished) return;
  Coder c = st->coder[lane];
  uint8_t* page = nullptr;
  int err = 0;
  if ((st->used >> lane) & 1u) {
    for (int i = 0; i < 32; i++) coder_write (c, page, st, ar, lane, 0, 128, &err);
    if (c.pos > 0 && !err) {
      const uint8_t* lastb = out_ptr (st, ar, lane, c.pos - 1, false);
      if (lastb && (*lastb & 0xe0) == 0xc0) {
        uint8_t* q = out_ptr (st, ar, lane, c.pos, true);
        if (q) { *q = 0; c.pos++; } else err = 2;
      }
    }
  }
  st->coder[lane] = c;
  const unsigned errm = __ballot_sync (0xffffffffu, err != 0);
  __syncwarp();
  if (lane == 0) { st->finished = 1; if (errm) st->error = 2; }
}

__global__ void k_model_init_state (MState* st) {
  const int lane = threadIdx.x;
  if (lane < 32) { st->coder[lane].lowvalue = 0; st->coder[lane].range = 255; st->coder[lane].count = -24; st->coder[lane].pos = 0; }
}

} // namespace lh264m
