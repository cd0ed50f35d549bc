// decoder.cpp — see decoder.h.  Clause numbers refer to ITU-T H.264.
#include "decoder.h"
#include "cavlc.h"
#include "cabac.h"
#include "cavlc_writer.h"
#include <memory>
#include <string.h>
#include <stdio.h>
#include <stdlib.h>
#include <algorithm>
#if defined(__SSE2__)
#include <emmintrin.h>
#endif

// The slice-data parser is the host-side bottleneck of the bitstream-in path: its hot functions are compiled twice, for
// x86-64-v3 (AVX2 / BMI1 / BMI2 / LZCNT / MOVBE: variable shifts, bit counts and the byte swap of the bit reader are single
// instructions there) and for the baseline ISA; the loader picks per CPU.  Same source, same results.
#if defined(__x86_64__) && defined(__GNUC__) && !defined(__clang__) && !defined(LH264_NO_CLONES)
#define LH264_HOT __attribute__ ((target_clones ("arch=x86-64-v3", "default")))
#else
#define LH264_HOT
#endif

namespace lh264host {

// Table 9-4: coded_block_pattern me(v) mapping for ChromaArrayType 1/2
const uint8_t kCbpIntra[48] = {47, 31, 15, 0, 23, 27, 29, 30, 7, 11, 13, 14, 39, 43, 45, 46, 16, 3, 5, 10, 12, 19, 21, 26, 28, 35, 37, 42, 44, 1, 2, 4,
                               8, 17, 18, 20, 24, 6, 9, 22, 25, 32, 33, 34, 36, 40, 38, 41
                              };
const uint8_t kCbpInter[48] = {0, 16, 1, 2, 4, 8, 32, 3, 5, 10, 12, 15, 47, 7, 11, 13, 14, 6, 9, 31, 35, 37, 42, 44, 33, 34, 36, 40, 39, 43, 45, 46,
                               17, 18, 20, 24, 19, 21, 26, 28, 23, 27, 29, 30, 22, 25, 38, 41
                              };
// z (decode order) <-> raster index of the 4x4 luma blocks
const uint8_t kZtoR[16] = {0, 1, 4, 5, 2, 3, 6, 7, 8, 9, 12, 13, 10, 11, 14, 15};
const uint8_t kRtoZ[16] = {0, 1, 4, 5, 2, 3, 6, 7, 8, 9, 12, 13, 10, 11, 14, 15};

namespace {

// LH264_DEBUG=1: per-slice / per-picture trace of the host decoder on stderr (reference lists, completeness); read once
bool debug_trace() { static const bool on = getenv ("LH264_DEBUG") != nullptr; return on; }

// Table 8-15: QPc as a function of qPI
const uint8_t kChromaQp[52] = {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17, 18, 19, 20, 21, 22, 23, 24, 25, 26, 27, 28, 29,
                               29, 30, 31, 32, 32, 33, 34, 34, 35, 35, 36, 36, 37, 37, 37, 38, 38, 38, 39, 39, 39, 39
                              };
const uint8_t kChromaDcScan[4] = {0, 16, 32, 48};   // DC k -> slot 0 of block k

struct RefPic {
  int slot = -1, frame_num = 0, pic_num = 0, long_idx = -1;
  bool is_long = false, exists = true;
};

struct Nb { bool avail; int ref; int mvx, mvy; };

// Records the parser never reads back (model records, coefficient pools) leave with non-temporal stores: no read-for-ownership
// of ~180 bytes per macroblock of staging memory that is only written here and read by the DMA engine.  With sixteen parse
// threads the host memory system, not the parse, sets the pace.
inline void stream_copy (void* dst, const void* src, size_t bytes) {          // bytes: multiple of 16
#if defined(__SSE2__)
  if ((reinterpret_cast<uintptr_t> (dst) & 15) == 0) {
    __m128i* d = static_cast<__m128i*> (dst);
    const __m128i* q = static_cast<const __m128i*> (src);
    for (size_t i = 0; i < bytes / 16; i++) _mm_stream_si128 (d + i, _mm_loadu_si128 (q + i));
    return;
  }
#endif
  memcpy (dst, src, bytes);
}
inline void stream_fence() {
#if defined(__SSE2__)
  _mm_sfence();
#endif
}

// Growable array of hand-off records that lives either in heap memory of the decoder or in a RecordArena (decoder.h).  grow()
// leaves the new elements uninitialised (the parser overwrites every one of them).
template <class T, int KIND> class RecBuf {
 public:
  RecBuf() {}
  RecBuf (const RecBuf&) = delete;
  RecBuf& operator= (const RecBuf&) = delete;
  ~RecBuf() { free (own_); }
  T* data() const { return p_; }
  size_t size() const { return n_; }
  T& operator[] (size_t i) { return p_[i]; }
  const T& operator[] (size_t i) const { return p_[i]; }
  T& back() { return p_[n_ - 1]; }
  void clear() { n_ = 0; if (arena_) { p_ = nullptr; cap_ = 0; } }       // (arena storage belongs to ONE picture)
  void pop_back() { n_--; }
  void truncate (size_t n) { n_ = n; }
  T* grow (size_t k) {
    if (n_ + k > cap_ && !reserve (n_ + k)) return nullptr;
    T* r = p_ + n_;
    n_ += k;
    return r;
  }
  bool reserve (size_t need) {
    if (need <= cap_) return true;
    if (arena_) {
      size_t capb = 0;
      uint8_t* q = arena_->storage (KIND, need * sizeof (T), &capb, reinterpret_cast<const uint8_t*> (p_), n_ * sizeof (T));
      if (!q) return false;
      p_ = reinterpret_cast<T*> (q); cap_ = capb / sizeof (T);
      return true;
    }
    const size_t ncap = std::max (need, std::max<size_t> (2 * own_cap_, 1024));
    T* q = static_cast<T*> (realloc (own_, ncap * sizeof (T)));
    if (!q) return false;
    own_ = q; own_cap_ = ncap; p_ = q; cap_ = ncap;
    return true;
  }
  bool set_arena (RecordArena* a) {
    if (!a && !arena_) return true;
    if (n_ == 0) {
      arena_ = a;
      if (a) { p_ = nullptr; cap_ = 0; } else { p_ = own_; cap_ = own_cap_; }
      return true;
    }
    if (a) {                                          // the records so far move into the arena
      size_t capb = 0;
      uint8_t* q = a->storage (KIND, n_ * sizeof (T), &capb, reinterpret_cast<const uint8_t*> (p_), n_ * sizeof (T));
      if (!q) return false;
      arena_ = a; p_ = reinterpret_cast<T*> (q); cap_ = capb / sizeof (T);
      return true;
    }
    if (n_ > own_cap_) {
      T* q = static_cast<T*> (realloc (own_, n_ * sizeof (T)));
      if (!q) return false;
      own_ = q; own_cap_ = n_;
    }
    memcpy (own_, p_, n_ * sizeof (T));
    arena_ = nullptr; p_ = own_; cap_ = own_cap_;
    return true;
  }
 private:
  T* p_ = nullptr;
  size_t n_ = 0, cap_ = 0;
  RecordArena* arena_ = nullptr;
  T* own_ = nullptr;
  size_t own_cap_ = 0;
};

inline int median3 (int a, int b, int c) { return a > b ? (b > c ? b : (a > c ? c : a)) : (a > c ? a : (b > c ? c : b)); }

} // namespace

struct Decoder::Impl {
  PictureSink* sink = nullptr;
  Sps sps[32];
  Pps pps[256];
  int active_sps = -1;
  int mb_w = 0, mb_h = 0, n_mbs = 0, n_slots = 0;
  std::vector<uint8_t> rbsp;

  // decoded picture buffer (reference marking, 8.2.5)
  std::vector<RefPic> refs;
  int max_long_idx = -1;                 // "no long-term frame indices" = -1
  int prev_ref_frame_num = 0;
  int serial = 0;

  // picture being decoded
  bool open = false;
  SliceHeader sh0;                       // first slice header of the picture
  int cur_slot = 0, decoded = 0;
  bool ltr_marking = false; int ltr_marked_frame_num = 0;   // set by mark_current() for the picture being finished
  bool alloc_failed = false;
  bool stream_coefs = true, stream_mrecs = true;   // Decoder::set_streaming_stores             // record storage could not be had (RecordArena / heap): the picture is dropped
  // the reference holds its output back until the first IDR picture has been decoded (bFreezeOutput under the freeze error
  // concealment modes h264dec selects, decoder_core.cpp:56-217): pictures of a stream whose leading IDR was lost are parsed,
  // modelled and reconstructed, but marked `output = false` for the sink and not handed out
  // Error concealment (error_concealment.cpp:43-471, DecodeFrameConstruction decoder_core.cpp:56-217).  `ec_mode` = ERROR_CON_IDC;
  // `freeze` = bFreezeOutput as h264dec leaves it (Initialize with ERROR_CON_SLICE_COPY clears it, the SetOption that follows
  // does not set it again): set by the two FREEZE_RES_CHANGE modes when a picture with an error changes the output size, cleared
  // by a complete error-free IDR picture.  `slot_incomplete` = !bIsComplete per picture slot (concealed, or predicted from such
  // a picture); `prev_slot` = pPreviousDecodedPictureInDpb.
  int ec_mode = 7;
  bool freeze = false;
  int last_out_w = 0, last_out_h = 0;
  uint32_t slot_incomplete = 0;
  int prev_slot = -1;
  int slot_poc[32];
  // DecodeFrameNoDelay (what h264dec calls) loses a picture that is only given up - concealed - when the NEXT access unit arrives:
  // its second DecodeFrame2 clears the output structure of the first (welsDecoderExt.cpp:408-430).  On for users of this class
  // that stand in for h264dec (lh264_rt_*, lh264_parse_dump); the API layer does the two halves itself and switches it off.
  bool nodelay_output = true, by_flush = false;
  uint32_t slots_touched = 0;            // picture slots that have been reconstructed into since the stream's storage was made
  bool freeze_output = true;             // Decoder::set_freeze_output
  unsigned long long ts = 0, ts_first = 0;
  bool ref_lost = false, ref_lost_first = false;   // a reference picture of the current picture was never received (frame_num gap / no reference at all)
  RecBuf<lh264_mb_rec, RecordArena::MBS> mbs;      // n_mbs records while a picture is open
  RecBuf<int16_t, RecordArena::COEFS> coefs;
  std::vector<lh264_slice_rec> slices;
  std::vector<uint8_t> nzc;              // [mb][24] total_coeff (luma raster 0..15, Cb 16..19, Cr 20..23)
  std::vector<int8_t> ipm;               // [mb][16] Intra4x4/8x8PredMode per raster block; 2 for other MB types
  std::vector<int16_t> mb_slice;         // slice number of each MB, -1 = not decoded in this picture
  std::vector<ModelMbSide> side;
  std::vector<uint8_t> slice_rbsp;
  // recompression-model hand-off (include/lh264_b200.h, lh264_model_*): records in slice-data order
  bool want_model = false;
  bool check_rewrite = false;
  std::unique_ptr<PictureSyntaxWriter> rewriter;
  RecBuf<lh264_model_mb, RecordArena::MRECS> mrecs;
  std::vector<lh264_model_slice> mslices;
  // per-position decoder state the reference keeps across pictures and reads back when it fills its
  // DecodedMacroblock (initRTDFromDecoderState, decode_slice.cpp:81-108): pCbp, pChromaPredMode,
  // pIntraPredMode[][7], pSubMbType are only written by the MB types that carry them
  std::vector<uint8_t> st_cbp, st_chroma, st_i16, st_sub;
  int st_alloc_w = 0, st_alloc_h = 0;

  // CABAC (9.3): the engine, and what context selection reads from neighbouring macroblocks of the picture
  struct CabMb {
    uint32_t cbf;                        // coded_block_flag: bits 0..15 luma 4x4 (raster), 16..19 Cb AC, 20..23 Cr AC, 24 luma DC, 25 Cb DC, 26 Cr DC
    uint8_t cbp, chroma;                 // coded_block_pattern as 9.3.3.1.1.4 sees it (I_PCM: 0x2f), intra_chroma_pred_mode as coded
  };
  CabacDecoder cab;
  int16_t lev_[384] = {0};               // residual scratch of parse_mb
  bool cabac_on = false, last_dqp_nz = false;
  // FMO (slice groups, map types 0 and 1 - the ones the reference decodes, fmo.cpp:53-104): slice group of every macroblock of
  // the picture for the slice being parsed; empty = one slice group, the walk is mb_addr + 1
  std::vector<uint8_t> fmo_map;
  void build_fmo_map (const Pps& pps) {
    fmo_map.clear();
    if (pps.num_slice_groups < 2) return;
    fmo_map.assign ((size_t)n_mbs, 0);
    const int g = pps.num_slice_groups;
    if (pps.slice_group_map_type == 0) {             // interleaved: runs of run_length[group] macroblocks, group after group
      int i = 0;
      while (i < n_mbs)
        for (int grp = 0; grp < g && i < n_mbs; grp++) {
          const int run = pps.run_length[grp];
          for (int j = 0; j < run && i + j < n_mbs; j++) fmo_map[(size_t) (i + j)] = (uint8_t)grp;
          i += run;
        }
    } else                                            // dispersed (8.2.2.2)
      for (int i = 0; i < n_mbs; i++) fmo_map[(size_t)i] = (uint8_t) (((i % mb_w) + (((i / mb_w) * g) >> 1)) % g);
  }
  // next macroblock of the slice: the next one of the same slice group (FmoNextMb, fmo.cpp:308-330); n_mbs = none
  int next_mb (int a) const {
    if (fmo_map.empty()) return a + 1;
    const uint8_t grp = fmo_map[(size_t)a];
    for (int n = a + 1; n < n_mbs; n++) if (fmo_map[(size_t)n] == grp) return n;
    return n_mbs;
  }
  // scaling lists: what the reference's multiplier tables were last built from (see the slice record set-up)
  bool scal_init = false, sps_late_pps = false;
  int scal_ppsid = -1;
  uint8_t scal4[6][16], scal8[2][64];
  std::vector<CabMb> cabmb;
  std::vector<uint8_t> mvd_abs;          // [mb][16][2] |mvd| per raster 4x4 block, saturated at 128

  // slice being parsed
  const Sps* S = nullptr;
  const Pps* P = nullptr;
  SliceHeader sh;
  int slice_no = 0, qp = 26;
  RefPic* list0[33];
  bool cur_done[16];                     // per raster 4x4 block of the current MB: MV already derived
  int cur_mb = 0, cur_mbx = 0, cur_mby = 0;        // macroblock being parsed and its position (set by parse_mb: no division per neighbour look-up)

  // ---------------------------------------------------------------------------------------------
  bool mb_avail (int mb) const { return mb >= 0 && mb < n_mbs && mb_slice[mb] == slice_no; }
  bool mb_avail_intra (int mb) const {   // usable for intra prediction samples / mode prediction
    return mb_avail (mb) && (!P->constrained_intra_pred || LH264_MB_IS_INTRA (mbs[mb].mb_type));
  }
  int nb_mb (int mbx, int mby, int dx, int dy) const {
    const int x = mbx + dx, y = mby + dy;
    if (x < 0 || x >= mb_w || y < 0 || y >= mb_h) return -1;
    return y * mb_w + x;
  }
  // neighbouring 4x4 block at block coordinates (bx,by) relative to the current MB's origin; may be
  // negative / >= 4.  Returns motion data as 6.4.11.7 + 8.4.1.3.2 want it.
  Nb nb_block (int bx, int by) const {
    Nb r{false, -2, 0, 0};
    int tx = cur_mbx, ty = cur_mby;
    if (bx < 0) { tx--; bx += 4; } else if (bx > 3) { tx++; bx -= 4; }
    if (by < 0) { ty--; by += 4; } else if (by > 3) return r;
    if (tx < 0 || tx >= mb_w || ty < 0) return r;
    const int t = ty * mb_w + tx;
    const int rr = by * 4 + bx;
    if (t == cur_mb) {
      if (!cur_done[rr]) return r;
    } else {
      if (t > cur_mb || !mb_avail (t)) return r;
    }
    const lh264_mb_rec& m = mbs[t];
    r.avail = true;
    if (!LH264_MB_IS_INTER (m.mb_type)) { r.ref = -1; return r; }
    r.ref = m.ref_idx[(by >> 1) * 2 + (bx >> 1)];
    r.mvx = m.mv[rr][0]; r.mvy = m.mv[rr][1];
    return r;
  }
  // 8.4.1.3: luma motion vector prediction for the partition at block (bx,by), size (bw,bh) blocks
  void pred_mv (int bx, int by, int bw, int bh, int ref, int shape /*0 general, 1 16x8, 2 8x16*/, int& px, int& py) const {
    const Nb A = nb_block (bx - 1, by), B = nb_block (bx, by - 1);
    Nb C = nb_block (bx + bw, by - 1);
    if (!C.avail) C = nb_block (bx - 1, by - 1);
    if (shape == 1) {                                   // 16x8: upper uses B, lower uses A when refs match
      if (by == 0) { if (B.ref == ref) { px = B.mvx; py = B.mvy; return; } }
      else if (A.ref == ref) { px = A.mvx; py = A.mvy; return; }
    } else if (shape == 2) {                            // 8x16: left uses A, right uses C
      if (bx == 0) { if (A.ref == ref) { px = A.mvx; py = A.mvy; return; } }
      else if (C.ref == ref) { px = C.mvx; py = C.mvy; return; }
    }
    // 8.4.1.3.1 median
    if (!B.avail && !C.avail && A.avail) { px = A.mvx; py = A.mvy; return; }
    const int match = (A.ref == ref) + (B.ref == ref) + (C.ref == ref);
    if (match == 1) {
      const Nb& n = A.ref == ref ? A : (B.ref == ref ? B : C);
      px = n.mvx; py = n.mvy;
      return;
    }
    px = median3 (A.mvx, B.mvx, C.mvx);
    py = median3 (A.mvy, B.mvy, C.mvy);
  }

  // ---- reference lists / marking -----------------------------------------------------------------
  int free_slot() const {
    for (int s = 0; s < n_slots; s++) {
      bool used = false;
      for (const RefPic& r : refs) if (r.slot == s) used = true;
      if (!used) return s;
    }
    return 0;
  }
  void build_list0() {
    const int max_fn = 1 << S->log2_max_frame_num;
    std::vector<RefPic*> sh_refs, lg_refs;
    for (RefPic& r : refs) {
      if (r.is_long) { r.pic_num = r.long_idx; lg_refs.push_back (&r); }
      else { r.pic_num = r.frame_num > sh.frame_num ? r.frame_num - max_fn : r.frame_num; sh_refs.push_back (&r); }
    }
    std::sort (sh_refs.begin(), sh_refs.end(), [] (RefPic * a, RefPic * b) { return a->pic_num > b->pic_num; });
    std::sort (lg_refs.begin(), lg_refs.end(), [] (RefPic * a, RefPic * b) { return a->pic_num < b->pic_num; });
    const int n = sh.num_ref_idx_l0;
    for (int i = 0; i < 33; i++) list0[i] = nullptr;
    int k = 0;
    for (RefPic* r : sh_refs) if (k < 32) list0[k++] = r;
    for (RefPic* r : lg_refs) if (k < 32) list0[k++] = r;
    // Reordering exactly as the reference performs it (WelsReorderRefList, manage_dec_ref.cpp:174-252):
    // the picture is searched from the END of the whole initial list down to the current
    // reordering index (by frame_num modulo MaxFrameNum, or long-term index) and rotated into
    // place; when it is not found the remaining commands are dropped and the list keeps its
    // current order.  This differs from 8.2.4.3 for streams that list one picture twice
    // (x264 weightp duplicates, roundtriptest/tibby8x8cavlc.264) and must be mirrored for parity.
    const int ref_count = k;
    int pred = sh.frame_num, idx = 0;
    for (const RefPicListMod& m : sh.mods) {
      int i = -1;
      if (m.op == 0 || m.op == 1) {
        const int abs_diff = (int)m.val + 1;
        pred = (m.op == 0 ? pred - abs_diff : pred + abs_diff) & (max_fn - 1);
        for (i = ref_count - 1; i >= idx; i--) if (list0[i]->frame_num == pred && !list0[i]->is_long) break;
      } else {
        for (i = ref_count - 1; i >= idx; i--) if (list0[i]->is_long && list0[i]->long_idx == (int)m.val) break;
      }
      if (i < idx) break;
      RefPic* pic = list0[i];
      for (int c = i; c > idx; c--) list0[c] = list0[c - 1];
      list0[idx++] = pic;
    }
    (void)n;
  }
  void mark_current() {                                  // 8.2.5 after the picture is decoded
    ltr_marking = false; ltr_marked_frame_num = 0;
    if (sh0.nal_ref_idc == 0) return;
    const int max_fn = 1 << S->log2_max_frame_num;
    RefPic cur;
    cur.slot = cur_slot; cur.frame_num = sh0.frame_num;
    if (sh0.idr) {
      refs.clear();
      if (sh0.long_term_reference_flag) { cur.is_long = true; cur.long_idx = 0; max_long_idx = 0; }
      else max_long_idx = -1;
      refs.push_back (cur);
      prev_ref_frame_num = sh0.frame_num;
      return;
    }
    bool cur_long = false, had5 = false;
    if (sh0.adaptive_marking) {
      auto pic_num_of = [&] (const RefPic & r) { return r.frame_num > sh0.frame_num ? r.frame_num - max_fn : r.frame_num; };
      for (const Mmco& m : sh0.mmcos) {
        switch (m.op) {
        case 1: {
          const int pn = sh0.frame_num - ((int)m.a + 1);
          for (size_t i = 0; i < refs.size(); i++) if (!refs[i].is_long && pic_num_of (refs[i]) == pn) { refs.erase (refs.begin() + i); break; }
          break;
        }
        case 2:
          for (size_t i = 0; i < refs.size(); i++) if (refs[i].is_long && refs[i].long_idx == (int)m.a) { refs.erase (refs.begin() + i); break; }
          break;
        case 3: {
          const int pn = sh0.frame_num - ((int)m.a + 1);
          for (size_t i = 0; i < refs.size(); i++) if (refs[i].is_long && refs[i].long_idx == (int)m.b) { refs.erase (refs.begin() + i); break; }
          for (RefPic& r : refs) if (!r.is_long && pic_num_of (r) == pn) { r.is_long = true; r.long_idx = (int)m.b; ltr_marking = true; ltr_marked_frame_num = pn & (max_fn - 1); break; }
          break;
        }
        case 4:
          max_long_idx = (int)m.a - 1;
          for (size_t i = 0; i < refs.size();) { if (refs[i].is_long && refs[i].long_idx > max_long_idx) refs.erase (refs.begin() + i); else i++; }
          break;
        case 5: refs.clear(); max_long_idx = -1; had5 = true; break;
        case 6:
          for (size_t i = 0; i < refs.size(); i++) if (refs[i].is_long && refs[i].long_idx == (int)m.b) { refs.erase (refs.begin() + i); break; }
          cur.is_long = true; cur.long_idx = (int)m.b; cur_long = true;
          ltr_marking = true; ltr_marked_frame_num = sh0.frame_num;
          break;
        }
      }
    } else {
      sliding_window();
    }
    (void)cur_long;
    if (had5) cur.frame_num = 0;
    refs.push_back (cur);
    prev_ref_frame_num = had5 ? 0 : sh0.frame_num;
  }
  void sliding_window() {
    const int max_refs = std::max (S->max_num_ref_frames, 1);
    while ((int)refs.size() >= max_refs) {
      int best = -1, best_num = 0;
      const int max_fn = 1 << S->log2_max_frame_num;
      for (size_t i = 0; i < refs.size(); i++) {
        if (refs[i].is_long) continue;
        const int w = refs[i].frame_num > sh0.frame_num ? refs[i].frame_num - max_fn : refs[i].frame_num;
        if (best < 0 || w < best_num) { best = (int)i; best_num = w; }
      }
      if (best < 0) break;
      refs.erase (refs.begin() + best);
    }
  }

  // ---- picture life cycle --------------------------------------------------------------------------
  int activate (const Sps& s) {
    if (active_sps == s.id && mb_w == s.mb_w && mb_h == s.mb_h && n_slots >= s.max_num_ref_frames + 2) return 0;
    active_sps = s.id;
    mb_w = s.mb_w; mb_h = s.mb_h; n_mbs = mb_w * mb_h;
    n_slots = std::max (3, s.max_num_ref_frames + 2);
    refs.clear(); max_long_idx = -1;
    slots_touched = 0;                     // (new picture storage: configure() below)
    mbs.clear();
    nzc.assign ((size_t)n_mbs * 24, 0);
    ipm.assign ((size_t)n_mbs * 16, 2);
    mb_slice.assign (n_mbs, -1);
    side.assign (n_mbs, ModelMbSide());
    // the reference reallocates (zeroed) its per-macroblock arrays only when the picture GROWS beyond what it has
    // (InitialDqLayersContext, decoder_core.cpp:1233-1235, :1344-1345); a smaller picture keeps the old bytes, indexed by the
    // new macroblock addresses
    if (mb_w * 16 > st_alloc_w || mb_h * 16 > st_alloc_h) {
      st_alloc_w = mb_w * 16; st_alloc_h = mb_h * 16;
      st_cbp.assign (n_mbs, 0); st_chroma.assign (n_mbs, 0); st_i16.assign (n_mbs, 0); st_sub.assign ((size_t)n_mbs * 4, 0);
    }
    cabmb.assign (n_mbs, CabMb()); mvd_abs.assign ((size_t)n_mbs * 32, 0);
    return sink->configure (mb_w, mb_h, n_slots);
  }
  bool is_new_picture (const SliceHeader& a, const SliceHeader& b) const {   // 7.4.1.2.4
    return a.frame_num != b.frame_num || a.pps_id != b.pps_id || (a.nal_ref_idc == 0) != (b.nal_ref_idc == 0) ||
           a.idr != b.idr || (a.idr && a.idr_pic_id != b.idr_pic_id) || a.poc_lsb != b.poc_lsb ||
           a.delta_poc_bottom != b.delta_poc_bottom || a.delta_poc[0] != b.delta_poc[0] || a.delta_poc[1] != b.delta_poc[1];
  }
  void start_picture (Decoder* self) {
    (void)self;
    open = true; sh0 = sh; decoded = 0; slice_no = 0;
    ts_first = ts;
    ref_lost_first = false;
    if (sh.idr) ref_lost = false;
    else {
      const int max_fn0 = 1 << S->log2_max_frame_num;
      const bool gap = !S->gaps_in_frame_num_allowed && sh.frame_num != prev_ref_frame_num && sh.frame_num != (prev_ref_frame_num + 1) % max_fn0;
      if ((refs.empty() || gap) && !ref_lost) { ref_lost = true; ref_lost_first = true; }
    }
    cur_mb = -2;                           // (cur_mbx / cur_mby are recomputed for the first macroblock)
    if (sh.idr) { /* refs dropped at marking time; lists are empty for an I picture anyway */ }
    else if (S->gaps_in_frame_num_allowed) {
      // 8.2.5.2: fill frame_num gaps with non-existing short-term frames
      const int max_fn = 1 << S->log2_max_frame_num;
      if (sh.frame_num != prev_ref_frame_num && sh.frame_num != (prev_ref_frame_num + 1) % max_fn) {
        int fn = (prev_ref_frame_num + 1) % max_fn, guard = 0;
        const SliceHeader keep = sh0;
        while (fn != sh.frame_num && guard++ < max_fn) {
          sh0.frame_num = fn;
          sliding_window();
          RefPic r; r.slot = refs.empty() ? 0 : refs.back().slot; r.frame_num = fn; r.exists = false;
          refs.push_back (r);
          prev_ref_frame_num = fn;
          fn = (fn + 1) % max_fn;
        }
        sh0 = keep;
      }
    }
    cur_slot = free_slot();
    slots_touched |= 1u << (cur_slot & 31);
    slices.clear(); coefs.clear(); mrecs.clear(); mslices.clear(); mbs.clear();
    if (!mbs.grow (n_mbs)) { open = false; alloc_failed = true; return; }
    coefs.reserve ((size_t)n_mbs * 32);               // two coefficient blocks per macroblock to start with
    // (records of macroblocks that stay undecoded are marked LH264_MB_NONE when the picture is handed over: no pass over
    // ~1 MB of records for the usual, complete picture)
    std::fill (mb_slice.begin(), mb_slice.end(), (int16_t) -1);
  }
  int finish_picture (Decoder* self) {
    if (!open) return 0;
    open = false;
    const bool all_decoded = decoded >= n_mbs;
    bool concealed = false;
    if (!all_decoded) {
      // ImplementErrorCon runs for a picture that has SOME macroblocks (decoder_core.cpp:2352, :2497: iTotalNumMbRec != 0)
      if (ec_mode != 0 && decoded > 0 && !slices.empty()) concealed = conceal();
      if (!concealed) for (int i = 0; i < n_mbs; i++) if (mb_slice[i] < 0) mbs[i].mb_type = LH264_MB_NONE;
    }
    stream_fence();                                   // the streamed records are visible before anybody is told about them
    lh264_picture_desc d;
    d.stream = 0; d.cur_slot = cur_slot; d.n_slices = (int)slices.size(); d.n_coef_blocks = (int) (coefs.size() / 16);
    d.slices = slices.data(); d.mbs = mbs.data(); d.coefs = coefs.data();
    DecodedPictureInfo info;
    info.mb_w = mb_w; info.mb_h = mb_h;
    info.crop_x = S->crop_left; info.crop_y = S->crop_top;
    info.width = mb_w * 16 - S->crop_left - S->crop_right; info.height = mb_h * 16 - S->crop_top - S->crop_bottom;
    info.frame_num = sh0.frame_num; info.poc_lsb = sh0.poc_lsb; info.idr_pic_id = sh0.idr_pic_id; info.nal_ref_idc = sh0.nal_ref_idc;
    info.slot = cur_slot; info.serial = serial++; info.complete = all_decoded; info.timestamp = ts; info.timestamp_first = ts_first;
    info.ref_lost = ref_lost; info.ref_lost_first = ref_lost_first;
    int rc = 0;
    // bIsComplete: every macroblock decoded AND every picture it predicts from complete (CheckRefPicturesComplete,
    // decoder_core.cpp:2529-2565: the reference index of every partition of every inter macroblock)
    bool ref_incomplete = false;
    if (slot_incomplete && !sh0.idr)
      for (int i = 0; i < n_mbs && !ref_incomplete; i++) {
        const lh264_mb_rec& m = mbs[i];
        if (mb_slice[i] < 0 || !LH264_MB_IS_INTER (m.mb_type) || (int)m.slice_idx >= (int)slices.size()) continue;
        const lh264_slice_rec& sl = slices[m.slice_idx];
        const int nq = (m.mb_type == LH264_MB_P16x16 || m.mb_type == LH264_MB_PSKIP) ? 1 : 4;
        for (int q = 0; q < nq; q++) if (m.ref_idx[q] >= 0 && ((slot_incomplete >> (sl.ref_slot[m.ref_idx[q] & 15] & 31)) & 1u)) ref_incomplete = true;
      }
    const bool pic_error = !all_decoded || ref_incomplete || ref_lost;
    info.error = pic_error;
    if (pic_error) slot_incomplete |= 1u << (cur_slot & 31); else slot_incomplete &= ~(1u << (cur_slot & 31));
    slot_poc[cur_slot & 31] = sh0.poc_lsb;
    if (all_decoded && sh0.idr && !pic_error) freeze = false;          // complete non-ECed IDR frame done (decoder_core.cpp:163-166)
    if (!all_decoded && !concealed) info.output = false;               // nothing to hand out (EC disabled, or no macroblock at all)
    else {
      const bool res_change = last_out_w != info.width || last_out_h != info.height;
      last_out_w = info.width; last_out_h = info.height;
      if (ec_mode == 0) info.output = !pic_error;                      // decoder_core.cpp:192-194
      else {
        if ((ec_mode == 5 || ec_mode == 7) && pic_error && res_change) freeze = true;   // :195-198
        info.output = !freeze;
      }
      if (concealed && nodelay_output && !by_flush) info.output = false;   // (constructed in the first half of DecodeFrameNoDelay: lost)
    }
    if (debug_trace()) fprintf (stderr, "picture: %dx%d idr %d frame_num %d decoded %d of %d slices %zu\n", mb_w, mb_h, (int)sh0.idr, sh0.frame_num, decoded, n_mbs, slices.size());
    if (!slices.empty()) rc = sink->submit (d, info);
    if (want_model && !mslices.empty()) {
      lh264_model_picture mp;
      mp.model_stream = 0; mp.mb_w = mb_w; mp.mb_h = mb_h; mp.n_slices = (int)mslices.size(); mp.n_recs = (int)mrecs.size();
      mp.n_coef_blocks = d.n_coef_blocks; mp.slices = mslices.data(); mp.recs = mrecs.data(); mp.coefs = coefs.data();
      const int rm = sink->submit_model (mp, info);
      if (!rc) rc = rm;
    }
    mark_current();
    if (!slices.empty()) prev_slot = cur_slot;           // pPreviousDecodedPictureInDpb (decoder_core.cpp:2365, :2505)
    info.idr = sh0.idr; info.ltr_marking = ltr_marking; info.ltr_marked_frame_num = ltr_marked_frame_num;
    self->stats_.pictures++;
    self->stats_.coef_blocks += coefs.size() / 16;
    if (info.output) self->out_.push_back (info);
    self->fin_.push_back (info);
    mbs.clear(); coefs.clear(); mrecs.clear();       // handed over: arena storage is the sink's now
    return rc;
  }

  // ---- error concealment inside a picture (ImplementErrorCon, error_concealment.cpp:443-471) ---------------------------
  // The macroblocks no slice brought are handed to the reconstruction kernel as ordinary records of a synthetic slice without
  // deblocking: a 16x16 motion-compensated block from the previous decoded picture (zero vector = DoErrorConSliceCopy, the
  // average vector of the picture's inter macroblocks = DoErrorConSliceMVCopy), I_PCM grey when there is no such picture; the
  // FRAME_COPY modes replace every macroblock.  (A decoded macroblock whose left / top neighbour is missing was deblocked by the
  // reference against whatever the buffer held there before; that is not reproduced - see DESIGN.md.)
  bool conceal() {
    if (slices.size() >= 65535) return false;
    int src = prev_slot;
    if ((ec_mode == 1 || ec_mode == 2) && sh0.idr) src = -1;           // no cross-IDR copy: fill
    const bool frame_copy = ec_mode == 1 || ec_mode == 3;
    const bool same_buffer = src >= 0 && src == cur_slot;             // (the reference copies a buffer onto itself / skips the MC)
    // GetAvilInfoFromCorrectMb: mean motion vector of the correctly decoded inter partitions that use reference index 0
    int mvx = 0, mvy = 0;
    bool use_mv = false;
    if ((ec_mode == 6 || ec_mode == 7) && src >= 0 && !sh0.idr) {
      long long sx = 0, sy = 0; int cnt = 0;
      auto add = [&] (const lh264_mb_rec& m, int q, int idx) { if (m.ref_idx[q] == 0) { sx += m.mv[idx][0]; sy += m.mv[idx][1]; cnt++; } };
      for (int i = 0; i < n_mbs; i++) {
        if (mb_slice[i] < 0) continue;
        const lh264_mb_rec& m = mbs[i];
        switch (m.mb_type) {
        case LH264_MB_PSKIP: case LH264_MB_P16x16: add (m, 0, 0); break;
        case LH264_MB_P16x8: add (m, 0, 0); add (m, 2, 8); break;
        case LH264_MB_P8x16: add (m, 0, 0); add (m, 1, 2); break;
        case LH264_MB_P8x8:
          for (int q = 0; q < 4; q++) {
            const int b = ((q >> 1) << 3) + ((q & 1) << 1);
            switch (m.sub_mb_type[q]) {
            case LH264_SUB_8x8: add (m, q, b); break;
            case LH264_SUB_8x4: add (m, q, b); add (m, q, b + 4); break;
            case LH264_SUB_4x8: add (m, q, b); add (m, q, b + 1); break;
            default: add (m, q, b); add (m, q, b + 1); add (m, q, b + 4); add (m, q, b + 5); break;
            }
          }
          break;
        default: break;
        }
      }
      if (cnt > 0) {
        mvx = (int) (sx / cnt); mvy = (int) (sy / cnt);
        use_mv = true;
        const int ref0 = slices.back().ref_slot[0];                   // pECRefPic[0]: entry 0 of the list in force (the last slice's)
        if (ref0 != src) {                                            // scaled by the distance in iFramePoc (= pic_order_cnt_lsb)
          const int s0 = slot_poc[ref0 & 31] - sh0.poc_lsb, s1 = slot_poc[src & 31] - sh0.poc_lsb;
          mvx = s0 == 0 ? 0 : mvx * s1 / s0; mvy = s0 == 0 ? 0 : mvy * s1 / s0;
        }
      }
    }
    lh264_slice_rec sr;
    memset (&sr, 0, sizeof (sr));
    sr.slice_type = SLICE_P; sr.disable_deblocking_idc = 1; sr.num_ref_idx_l0 = 1;
    for (int i = 0; i < LH264_MAX_REFS; i++) sr.ref_slot[i] = (int16_t) (src >= 0 ? src : cur_slot);
    const int sidx = (int)slices.size();
    int first = -1, count = 0;
    const int W = mb_w * 16, H = mb_h * 16;
    for (int i = 0; i < n_mbs; i++) {
      if (!frame_copy && mb_slice[i] >= 0) continue;
      if (first < 0) first = i;
      count++;
      lh264_mb_rec& m = mbs[i];
      memset (&m, 0, sizeof (m));
      m.slice_idx = (uint16_t)sidx;
      if (same_buffer) { m.mb_type = LH264_MB_NONE; continue; }
      if (src < 0) {                                                  // memset 128 (error_concealment.cpp:139-156)
        m.mb_type = LH264_MB_IPCM; m.coded_mask = 0xffffffu; m.cbp = 0x2f; m.nz_mask = 0xffffffu;
        m.coef_off = (uint32_t) (coefs.size() / 16);
        int16_t* o = coefs.grow (384);
        if (!o) { m.mb_type = LH264_MB_NONE; m.coded_mask = 0; continue; }
        for (int k = 0; k < 384; k++) o[k] = 128;
        continue;
      }
      m.mb_type = LH264_MB_P16x16;
      int vx = 0, vy = 0;
      if (use_mv) {                                                   // DoMbECMvCopy: the block stays inside the (cropped) picture
        const int px = (i % mb_w) << 4, py = (i / mb_w) << 4;
        int fx = (px << 2) + mvx, fy = (py << 2) + mvy;
        int l = 0, t = 0, r = W, b = H;
        if (S->crop_left || S->crop_right || S->crop_top || S->crop_bottom) { l = S->crop_left; r = W - S->crop_right; t = S->crop_top; b = H - S->crop_top; }
        if (fx < ((l + 2) << 2)) { fx &= ~3; fx = std::max (l, fx); }
        else if (fx > ((r - 19) << 2)) { fx &= ~3; fx = std::min ((r - 17) << 2, fx); }
        if (fy < ((t + 2) << 2)) { fy &= ~3; fy = std::max (t, fy); }
        else if (fy > ((b - 19) << 2)) { fy &= ~3; fy = std::min ((b - 17) << 2, fy); }
        vx = fx - (px << 2); vy = fy - (py << 2);
      }
      for (int k = 0; k < 16; k++) { m.mv[k][0] = (int16_t)vx; m.mv[k][1] = (int16_t)vy; }
    }
    if (count == 0) return false;
    sr.first_mb = first; sr.num_mbs = count;
    slices.push_back (sr);
    return true;
  }

  // ---- slice data ------------------------------------------------------------------------------------
  int parse_slice_data (BitReader& br, Decoder* self);
  int parse_mb (BitReader& br, int mb_addr, bool skipped, lh264_model_mb* mr);
  int parse_residual (BitReader& br, lh264_mb_rec& m, int mb_addr, int16_t* lev, uint32_t& coded, bool i16, int cbp);
  void resolve_intra (lh264_mb_rec& m, int mb_addr, const int8_t* modes, bool i8x8);
  int nC_luma (int mb_addr, int r) const;
  int nC_chroma (int mb_addr, int plane, int b) const;
  // CABAC twins of the syntax element readers (9.3.2 binarisations, 9.3.3.1 ctxIdx derivation)
  int parse_slice_data_cabac (BitReader& br, Decoder* self);
  int parse_residual_cabac (lh264_mb_rec& m, int mb_addr, int16_t* lev, uint32_t& coded, bool i16, int cbp);
  int cabac_block (int cat, int cbf_inc, int16_t* out, const uint8_t* scan, int scan_off);
  int cabac_intra_mb_type (int base, bool i_slice, int mb_addr);
  int cabac_mb_type (int mb_addr);
  int cabac_sub_mb_type();
  int cabac_ref_idx (int mb_addr, int q, const int8_t* cur_refs);
  int cabac_mvd (int mb_addr, int bx, int by, int comp);
  int cabac_cbp (int mb_addr);
  int cabac_chroma_mode (int mb_addr);
  int cabac_dqp();
  int cabac_t8_flag (int mb_addr);
  int cbf_inc (int mb_addr, int bit_cur_a, int bit_left, int bit_cur_b, int bit_top) const;
};

// ---- nC derivation (9.2.1) -----------------------------------------------------------------------------
int Decoder::Impl::nC_luma (int mb_addr, int r) const {
  const int mbx = cur_mbx, mby = cur_mby, bx = r & 3, by = r >> 2;        // (mb_addr == cur_mb)
  int nA = -1, nB = -1;
  if (bx > 0) nA = nzc[(size_t)mb_addr * 24 + r - 1];
  else { const int n = nb_mb (mbx, mby, -1, 0); if (mb_avail (n)) nA = nzc[(size_t)n * 24 + r + 3]; }
  if (by > 0) nB = nzc[(size_t)mb_addr * 24 + r - 4];
  else { const int n = nb_mb (mbx, mby, 0, -1); if (mb_avail (n)) nB = nzc[(size_t)n * 24 + r + 12]; }
  if (nA >= 0 && nB >= 0) return (nA + nB + 1) >> 1;
  return nA >= 0 ? nA : (nB >= 0 ? nB : 0);
}
int Decoder::Impl::nC_chroma (int mb_addr, int plane, int b) const {
  const int mbx = cur_mbx, mby = cur_mby, bx = b & 1, by = b >> 1, base = 16 + 4 * plane;
  int nA = -1, nB = -1;
  if (bx > 0) nA = nzc[(size_t)mb_addr * 24 + base + b - 1];
  else { const int n = nb_mb (mbx, mby, -1, 0); if (mb_avail (n)) nA = nzc[(size_t)n * 24 + base + b + 1]; }
  if (by > 0) nB = nzc[(size_t)mb_addr * 24 + base + b - 2];
  else { const int n = nb_mb (mbx, mby, 0, -1); if (mb_avail (n)) nB = nzc[(size_t)n * 24 + base + b + 2]; }
  if (nA >= 0 && nB >= 0) return (nA + nB + 1) >> 1;
  return nA >= 0 ? nA : (nB >= 0 ? nB : 0);
}

// ---- intra mode resolution: what the reference does at parse time (CheckIntraNxNPredMode etc.,
// parse_mb_syn_cavlc.cpp:519-612) expressed on neighbour availability -----------------------------------
void Decoder::Impl::resolve_intra (lh264_mb_rec& m, int mb_addr, const int8_t* modes, bool i8x8) {
  const int mbx = cur_mbx, mby = cur_mby;            // (mb_addr == cur_mb)
  const bool L = mb_avail_intra (nb_mb (mbx, mby, -1, 0)), T = mb_avail_intra (nb_mb (mbx, mby, 0, -1));
  const bool TL = mb_avail_intra (nb_mb (mbx, mby, -1, -1)), TR = mb_avail_intra (nb_mb (mbx, mby, 1, -1));
  m.nxn_avail = (uint8_t) ((T ? 1 : 0) | (TL ? 2 : 0) | (L ? 4 : 0) | (TR ? 8 : 0));
  auto fix = [] (int mode, bool l, bool t, bool tr) -> uint8_t {
    if (mode == LH264_I4_DC) return (uint8_t) (l && t ? LH264_I4_DC : (l ? LH264_I4_DC_L : (t ? LH264_I4_DC_T : LH264_I4_DC_128)));
    if (mode == LH264_I4_DDL && !tr) return LH264_I4_DDL_TOP;
    if (mode == LH264_I4_VL && !tr) return LH264_I4_VL_TOP;
    return (uint8_t)mode;
  };
  if (modes) {
    if (i8x8) {
      for (int q = 0; q < 4; q++) {
        const bool l = (q & 1) ? true : L, t = (q >> 1) ? true : T;
        const bool tr = q == 0 ? T : (q == 1 ? TR : (q == 2));
        const uint8_t f = fix (modes[q], l, t, tr);
        for (int k = 0; k < 4; k++) m.intra_mode[4 * q + k] = f;
      }
    } else {
      for (int z = 0; z < 16; z++) {
        const int r = kZtoR[z], bx = r & 3, by = r >> 2;
        const bool l = bx > 0 ? true : L, t = by > 0 ? true : T;
        bool tr;
        if (by == 0) tr = bx < 3 ? T : TR;
        else tr = bx < 3 && kRtoZ[r - 3] < z;          // block (bx+1, by-1) decoded earlier?
        m.intra_mode[z] = fix (modes[z], l, t, tr);
      }
    }
  }
  // Intra16x16 / chroma DC variants (CheckIntra16x16PredMode / CheckIntraChromaPredMode)
  if (m.mb_type == LH264_MB_I16x16 && m.i16_mode == LH264_I16_DC)
    m.i16_mode = (uint8_t) (L && T ? LH264_I16_DC : (L ? LH264_I16_DC_L : (T ? LH264_I16_DC_T : LH264_I16_DC_128)));
  if (m.chroma_mode == LH264_C_DC)
    m.chroma_mode = (uint8_t) (L && T ? LH264_C_DC : (L ? LH264_C_DC_L : (T ? LH264_C_DC_T : LH264_C_DC_128)));
}

// ---- residual( ) for CAVLC (7.3.5.3) --------------------------------------------------------------------
LH264_HOT int Decoder::Impl::parse_residual (BitReader& br, lh264_mb_rec& m, int mb_addr, int16_t* lev, uint32_t& coded, bool i16, int cbp) {
  uint8_t* nz = &nzc[(size_t)mb_addr * 24];
  const bool t8 = m.flags & LH264_MBF_TRANSFORM8x8;
  if (i16) {                                        // Intra16x16DCLevel: 16 DC levels, zig-zag over the 4x4 DC matrix
    int16_t dc[16];
    memset (dc, 0, sizeof (dc));
    const int tc = cavlc_residual_block (br, nC_luma (mb_addr, 0), 16, kZigzag4x4, 1, 0, dc);
    if (tc < 0) return -1;
    for (int r = 0; r < 16; r++) if (dc[r]) { const int z = kRtoZ[r]; lev[16 * z] = dc[r]; coded |= 1u << z; }
  }
  for (int q = 0; q < 4; q++) {
    if (! (cbp & (1 << q))) { uint8_t* z = nz + (q >> 1) * 8 + (q & 1) * 2; z[0] = z[1] = z[4] = z[5] = 0; continue; }
    for (int k = 0; k < 4; k++) {
      const int z = 4 * q + k, r = kZtoR[z];
      int tc;
      if (t8) tc = cavlc_residual_block (br, nC_luma (mb_addr, r), 16, kZigzag8x8, 4, k, lev + 64 * q);
      else if (i16) tc = cavlc_residual_block (br, nC_luma (mb_addr, r), 15, kZigzag4x4, 1, 1, lev + 16 * z);
      else tc = cavlc_residual_block (br, nC_luma (mb_addr, r), 16, kZigzag4x4, 1, 0, lev + 16 * z);
      if (tc < 0) return -1;
      nz[r] = (uint8_t)tc;
      if (tc) coded |= t8 ? (0xFu << (4 * q)) : (1u << z);
    }
  }
  const int cc = cbp >> 4;
  if (cc) {
    for (int p = 0; p < 2; p++) {                   // ChromaDCLevel
      int16_t dc[4] = {0, 0, 0, 0};
      static const uint8_t kDcScan[4] = {0, 1, 2, 3};
      const int tc = cavlc_residual_block (br, -1, 4, kDcScan, 1, 0, dc);
      if (tc < 0) return -1;
      if (tc) for (int b = 0; b < 4; b++) if (dc[b]) { lev[16 * (16 + 4 * p + b)] = dc[b]; coded |= 1u << (16 + 4 * p + b); }
    }
  }
  for (int p = 0; p < 2; p++) for (int b = 0; b < 4; b++) {
      if (cc & 2) {                                 // ChromaACLevel
        const int tc = cavlc_residual_block (br, nC_chroma (mb_addr, p, b), 15, kZigzag4x4, 1, 1, lev + 16 * (16 + 4 * p + b));
        if (tc < 0) return -1;
        nz[16 + 4 * p + b] = (uint8_t)tc;
        if (tc) coded |= 1u << (16 + 4 * p + b);
      } else nz[16 + 4 * p + b] = 0;
    }
  return 0;
}

// ---- macroblock_layer( ) (7.3.5) -----------------------------------------------------------------------
LH264_HOT int Decoder::Impl::parse_mb (BitReader& br, int mb_addr, bool skipped, lh264_model_mb* mr) {
  lh264_mb_rec& m = mbs[mb_addr];
  memset (&m, 0, sizeof (m));
  cur_mb = mb_addr;
  m.slice_idx = (uint16_t)slice_no;
  uint8_t* nz = &nzc[(size_t)mb_addr * 24];
  int8_t* pm = &ipm[(size_t)mb_addr * 16];
  memset (pm, 2, 16);
  if (mb_addr == cur_mb + 1 && cur_mbx + 1 < mb_w) cur_mbx++;                        // the usual step: next macroblock of the row
  else { cur_mbx = mb_addr % mb_w; cur_mby = mb_addr / mb_w; }
  const int mbx = cur_mbx, mby = cur_mby;
  auto set_qp = [&] (lh264_mb_rec & mm) {
    mm.qp_y = (uint8_t)qp;
    mm.qp_c[0] = kChromaQp[std::min (51, std::max (0, qp + P->chroma_qp_index_offset))];
    mm.qp_c[1] = kChromaQp[std::min (51, std::max (0, qp + P->second_chroma_qp_index_offset))];
  };
  memset (cur_done, 0, sizeof (cur_done));
  memset (m.ref_idx, -1, 4);
  uint8_t* const mvda = cabac_on ? &mvd_abs[(size_t)mb_addr * 32] : nullptr;
  if (cabac_on) { cabmb[mb_addr] = CabMb(); memset (mvda, 0, 32); }
  auto put_mvd = [&] (int bx, int by, int bw, int bh, int dx, int dy) {
    const uint8_t ax = (uint8_t)std::min (128, std::abs (dx)), ay = (uint8_t)std::min (128, std::abs (dy));
    for (int y = by; y < by + bh; y++) for (int x = bx; x < bx + bw; x++) { mvda[2 * (y * 4 + x)] = ax; mvda[2 * (y * 4 + x) + 1] = ay; }
  };

  if (skipped) {                                    // P_Skip: 8.4.1.1
    m.mb_type = LH264_MB_PSKIP;
    set_qp (m);
    memset (nz, 0, 24);
    memset (m.ref_idx, 0, 4);
    int px = 0, py = 0;
    const Nb A = nb_block (-1, 0), B = nb_block (0, -1);
    if (A.avail && B.avail && ! (A.ref == 0 && A.mvx == 0 && A.mvy == 0) && ! (B.ref == 0 && B.mvx == 0 && B.mvy == 0))
      pred_mv (0, 0, 4, 4, 0, 0, px, py);
    {
      const uint32_t v = (uint32_t) (uint16_t) (int16_t)px | ((uint32_t) (uint16_t) (int16_t)py << 16);
      uint32_t* mvw = reinterpret_cast<uint32_t*> (&m.mv[0][0]);
      for (int r = 0; r < 16; r++) mvw[r] = v;
    }
    mb_slice[mb_addr] = (int16_t)slice_no;
    st_cbp[mb_addr] = 0;                             // pCbp[iMbXy] = 0, decode_slice.cpp:3945
    if (mr) { mr->type_code = 11; mr->luma_qp = (uint8_t)qp; }
    last_dqp_nz = false;
    return 0;
  }

  uint32_t t = cabac_on ? (uint32_t)cabac_mb_type (mb_addr) : br.ue();
  const bool p_slice = sh.slice_type == SLICE_P;
  bool intra = true;
  if (p_slice) {
    if (t < 5) intra = false;
    else t -= 5;
  }
  if (intra && t > 25) return -1;
  int16_t* const lev = lev_;                         // all zero between macroblocks (see below)
  uint32_t coded = 0;
  int cbp = 0;
  bool i16 = false;

  if (!intra) {
    static const uint8_t kind[5] = {LH264_MB_P16x16, LH264_MB_P16x8, LH264_MB_P8x16, LH264_MB_P8x8, LH264_MB_P8x8};
    m.mb_type = kind[t];
    const int nref = sh.num_ref_idx_l0;
    bool sub_lt8 = false;
    auto put_mv = [&] (int bx, int by, int bw, int bh, int mvx, int mvy) {
      const uint32_t v = (uint32_t) (uint16_t) (int16_t)mvx | ((uint32_t) (uint16_t) (int16_t)mvy << 16);
      uint32_t* mvw = reinterpret_cast<uint32_t*> (&m.mv[0][0]);                   // mv[r] = {x, y}: one 32-bit word per block
      for (int y = by; y < by + bh; y++) for (int x = bx; x < bx + bw; x++) { mvw[y * 4 + x] = v; cur_done[y * 4 + x] = true; }
    };
    if (m.mb_type == LH264_MB_P8x8) {
      for (int q = 0; q < 4; q++) {
        const uint32_t s = cabac_on ? (uint32_t)cabac_sub_mb_type() : br.ue();
        if (s > 3) return -1;
        m.sub_mb_type[q] = (uint8_t)s;
        if (s) sub_lt8 = true;
        st_sub[(size_t)mb_addr * 4 + q] = (uint8_t) (1u << s);   // g_ksInterSubMbTypeInfo[].iType
      }
      for (int q = 0; q < 4; q++) {
        int ref = 0;
        if (t != 4 && nref > 1) { ref = cabac_on ? cabac_ref_idx (mb_addr, q, m.ref_idx) : (int)br.te (nref - 1); if (ref < 0 || ref >= nref) return -1; }
        m.ref_idx[q] = (int8_t)ref;
        if (mr) mr->ref_idx[q] = (int8_t)ref;
      }
      // the standard makes a block "available" only once its motion has been derived; the ref_idx of
      // later quadrants must therefore stay hidden from nb_block(): cur_done gates them.
      for (int q = 0; q < 4; q++) {
        const int bx0 = (q & 1) * 2, by0 = (q >> 1) * 2, ref = m.ref_idx[q];
        static const uint8_t npart[4] = {1, 2, 2, 4};
        for (int s = 0; s < npart[m.sub_mb_type[q]]; s++) {
          int bx = bx0, by = by0, bw = 2, bh = 2;
          switch (m.sub_mb_type[q]) {
          case LH264_SUB_8x4: bh = 1; by += s; break;
          case LH264_SUB_4x8: bw = 1; bx += s; break;
          case LH264_SUB_4x4: bw = bh = 1; bx += s & 1; by += s >> 1; break;
          default: break;
          }
          int px, py;
          pred_mv (bx, by, bw, bh, ref, 0, px, py);
          int dx, dy;
          if (cabac_on) { dx = cabac_mvd (mb_addr, bx, by, 0); dy = cabac_mvd (mb_addr, bx, by, 1); put_mvd (bx, by, bw, bh, dx, dy); }
          else { dx = br.se(); dy = br.se(); }
          if (mr) { mr->mvd[by * 4 + bx][0] = (int16_t)dx; mr->mvd[by * 4 + bx][1] = (int16_t)dy; }
          put_mv (bx, by, bw, bh, px + dx, py + dy);
        }
      }
    } else {
      const int nparts = m.mb_type == LH264_MB_P16x16 ? 1 : 2;
      int refs_[2] = {0, 0};
      int8_t vis[4] = {-1, -1, -1, -1};              // ref_idx of the partitions parsed so far, per 8x8 (CABAC context selection)
      for (int p = 0; p < nparts; p++) {
        const int q0 = m.mb_type == LH264_MB_P16x8 ? 2 * p : p;
        if (nref > 1) { refs_[p] = cabac_on ? cabac_ref_idx (mb_addr, q0, vis) : (int)br.te (nref - 1); if (refs_[p] < 0 || refs_[p] >= nref) return -1; }
        if (m.mb_type == LH264_MB_P16x8) vis[2 * p] = vis[2 * p + 1] = (int8_t)refs_[p];
        else if (m.mb_type == LH264_MB_P8x16) vis[p] = vis[p + 2] = (int8_t)refs_[p];
      }
      for (int p = 0; p < nparts; p++) {
        int bx = 0, by = 0, bw = 4, bh = 4, shape = 0;
        if (m.mb_type == LH264_MB_P16x8) { bh = 2; by = 2 * p; shape = 1; }
        else if (m.mb_type == LH264_MB_P8x16) { bw = 2; bx = 2 * p; shape = 2; }
        // ref_idx of this partition becomes visible together with its MV
        for (int y = by; y < by + bh; y += 2) for (int x = bx; x < bx + bw; x += 2) m.ref_idx[(y >> 1) * 2 + (x >> 1)] = (int8_t)refs_[p];
        int px, py;
        pred_mv (bx, by, bw, bh, refs_[p], shape, px, py);
        int dx, dy;
        if (cabac_on) { dx = cabac_mvd (mb_addr, bx, by, 0); dy = cabac_mvd (mb_addr, bx, by, 1); put_mvd (bx, by, bw, bh, dx, dy); }
        else { dx = br.se(); dy = br.se(); }
        if (mr) { mr->ref_idx[p] = (int8_t)refs_[p]; mr->mvd[by * 4 + bx][0] = (int16_t)dx; mr->mvd[by * 4 + bx][1] = (int16_t)dy; }
        put_mv (bx, by, bw, bh, px + dx, py + dy);
      }
    }
    if (mr) mr->type_code = (uint8_t) (3 + t);          // 16x16 3, 16x8 4, 8x16 5, 8x8 6, 8x8ref0 7
    if (cabac_on) cbp = cabac_cbp (mb_addr);
    else {
      const uint32_t c = br.ue();
      if (c > 47) return -1;
      cbp = kCbpInter[c];
    }
    if ((cbp & 15) && P->transform_8x8_mode && ! (m.mb_type == LH264_MB_P8x8 && sub_lt8)) {
      if (cabac_on ? cabac_t8_flag (mb_addr) : (int)br.u1()) m.flags |= LH264_MBF_TRANSFORM8x8;
      if (mr) mr->flags |= LH264_MM_T8_CODED;
    }
  } else if (t == 25) {                               // I_PCM
    m.mb_type = LH264_MB_IPCM;
    br.align();
    const size_t at = coefs.size();
    int16_t* o = coefs.grow (384);
    if (!o) return -1;
    if (br.pos() + 384 * 8 > br.size_bits()) { coefs.truncate (at); return -1; }   // samples cut off (damaged stream)
    const uint8_t* src = br.data() + (br.pos() >> 3);
    for (int i = 0; i < 384; i++) o[i] = src[i];
    br.skip (384 * 8);
    m.coded_mask = 0xFFFFFF; m.coef_off = (uint32_t) (at / 16);
    m.nz_mask = 0xFFFFFF;
    memset (nz, 16, 24);
    m.qp_y = 0; m.qp_c[0] = m.qp_c[1] = 0;          // decode_slice.cpp:3258-3259
    mb_slice[mb_addr] = (int16_t)slice_no;
    if (mr) {                                         // the reference models no coefficients for I_PCM and reads back retained state
      mr->type_code = 8; mr->luma_qp = 0; mr->cbp = st_cbp[mb_addr]; mr->chroma_mode = st_chroma[mb_addr]; mr->i16_mode = st_i16[mb_addr];
      mr->coded_mask = 0; mr->coef_off = 0;
    }
    if (cabac_on) {                                   // 9.3.1.2: the engine starts again behind the samples
      cab.start (&br);
      cabmb[mb_addr].cbf = 0x7FFFFFFu; cabmb[mb_addr].cbp = 0x2f;
      last_dqp_nz = false;
    }
    return br.overrun() ? -1 : 0;
  } else if (t == 0) {                                // I_NxN
    bool i8 = false;
    if (P->transform_8x8_mode) i8 = cabac_on ? cabac_t8_flag (mb_addr) : (int)br.u1();
    m.mb_type = i8 ? LH264_MB_I8x8 : LH264_MB_I4x4;
    if (i8) m.flags |= LH264_MBF_TRANSFORM8x8;
    int8_t modes[16];
    const int nL = nb_mb (mbx, mby, -1, 0), nT = nb_mb (mbx, mby, 0, -1);
    const bool aL = mb_avail_intra (nL), aT = mb_avail_intra (nT);
    const int n = i8 ? 4 : 16;
    for (int i = 0; i < n; i++) {
      // 8.3.1.1 / 8.3.2.1: predicted mode = min(left, top), 2 when either neighbour is unavailable
      const int r = i8 ? kZtoR[4 * i] : kZtoR[i], bx = r & 3, by = r >> 2;
      int a, b;
      bool okA = true, okB = true;
      if (bx > 0) a = pm[r - 1]; else if (aL) a = ipm[(size_t)nL * 16 + r + 3]; else { a = 2; okA = false; }
      if (by > 0) b = pm[r - 4]; else if (aT) b = ipm[(size_t)nT * 16 + r + 12]; else { b = 2; okB = false; }
      if (i8 && bx == 0 && aL && mbs[nL].mb_type == LH264_MB_I4x4) a = ipm[(size_t)nL * 16 + r + 3];          // 4x4 block 8x8idxA*4+1 = (x=3, y=by)
      if (i8 && by == 0 && aT && mbs[nT].mb_type == LH264_MB_I4x4) b = ipm[(size_t)nT * 16 + 12 + bx];        // 4x4 block 8x8idxB*4+2 = (x=bx, y=3)
      int pred = (!okA || !okB) ? 2 : std::min (a, b);
      int mode = pred;
      if (cabac_on) {                                 // prev_intra4x4_pred_mode_flag ctxIdx 68, rem_intra4x4_pred_mode 3 bins of ctxIdx 69, LSB first
        if (!cab.decision (68)) { int rem = cab.decision (69); rem |= cab.decision (69) << 1; rem |= cab.decision (69) << 2; mode = rem < pred ? rem : rem + 1; }
      } else if (!br.u1()) { const int rem = (int)br.u (3); mode = rem < pred ? rem : rem + 1; }
      modes[i] = (int8_t)mode;
      if (mr) {
        mr->pred_mode[i] = (uint8_t) (mode | (pred << 4));
        if (!i8) {                                    // sample availability as CheckIntraNxNPredMode sees it (iSampleAvail grid)
          const int nTL = nb_mb (mbx, mby, -1, -1);
          const bool l = bx > 0 ? true : aL, tp = by > 0 ? true : aT;
          const bool tl = (bx > 0 && by > 0) ? true : (bx > 0 ? aT : (by > 0 ? aL : mb_avail_intra (nTL)));
          mr->pred_avail[i] = (uint8_t) ((l ? 4 : 0) | (tp ? 2 : 0) | (tl ? 1 : 0));
        }
      }
      if (i8) { pm[r] = pm[r + 1] = pm[r + 4] = pm[r + 5] = (int8_t)mode; }
      else pm[r] = (int8_t)mode;
    }
    const uint32_t cm = cabac_on ? (uint32_t)cabac_chroma_mode (mb_addr) : br.ue();
    if (cm > 3) return -1;
    m.chroma_mode = (uint8_t)cm;
    if (cabac_on) cabmb[mb_addr].chroma = (uint8_t)cm;
    resolve_intra (m, mb_addr, modes, i8);
    st_chroma[mb_addr] = m.chroma_mode;
    if (mr) mr->type_code = i8 ? 2 : 0;
    if (cabac_on) cbp = cabac_cbp (mb_addr);
    else {
      const uint32_t c = br.ue();
      if (c > 47) return -1;
      cbp = kCbpIntra[c];
    }
  } else {                                            // Intra16x16
    m.mb_type = LH264_MB_I16x16;
    i16 = true;
    m.i16_mode = (uint8_t) ((t - 1) & 3);
    cbp = (((t - 1) >> 2) % 3) << 4 | ((t - 1) >= 12 ? 15 : 0);
    const uint32_t cm = cabac_on ? (uint32_t)cabac_chroma_mode (mb_addr) : br.ue();
    if (cm > 3) return -1;
    m.chroma_mode = (uint8_t)cm;
    if (cabac_on) cabmb[mb_addr].chroma = (uint8_t)cm;
    resolve_intra (m, mb_addr, nullptr, false);
    st_chroma[mb_addr] = m.chroma_mode; st_i16[mb_addr] = m.i16_mode;
    if (mr) mr->type_code = 1;
  }
  m.cbp = (uint8_t)cbp;
  st_cbp[mb_addr] = (uint8_t)cbp;
  if (cabac_on) cabmb[mb_addr].cbp = (uint8_t)cbp;
  if (cbp > 0 || i16) {
    const int dqp = cabac_on ? cabac_dqp() : br.se();
    last_dqp_nz = dqp != 0;
    if (dqp < -26 || dqp > 25) return -1;
    qp = (qp + dqp + 52) % 52;
    set_qp (m);
    // the residual parsers write levels only into blocks they flag in `coded`, and the copy below zeroes those again: the
    // 768-byte scratch never needs a full clear (only after a parse error, which may stop in the middle of a block)
    if ((cabac_on ? parse_residual_cabac (m, mb_addr, lev, coded, i16, cbp) : parse_residual (br, m, mb_addr, lev, coded, i16, cbp)) < 0) { memset (lev_, 0, sizeof (lev_)); return -1; }
    if (coded) {
      m.coded_mask = coded;
      m.coef_off = (uint32_t) (coefs.size() / 16);
      int16_t* o = coefs.grow (16 * (size_t)__builtin_popcount (coded));
      if (!o) { memset (lev_, 0, sizeof (lev_)); return -1; }
      for (uint32_t mk = coded; mk; mk &= mk - 1) { const int k = __builtin_ctz (mk); if (stream_coefs) stream_copy (o, lev + 16 * k, 32); else memcpy (o, lev + 16 * k, 32); memset (lev + 16 * k, 0, 32); o += 16; }
    }
  } else {
    set_qp (m);
    memset (nz, 0, 24);
    last_dqp_nz = false;
  }
  uint32_t nzm = 0;
  for (int w = 0; w < 3; w++) {                       // 24 total_coeff bytes -> 24 flag bits, eight at a time
    uint64_t v;
    memcpy (&v, nz + 8 * w, 8);
    v = ((v & 0x7f7f7f7f7f7f7f7full) + 0x7f7f7f7f7f7f7f7full) | v;      // bit 7 of every byte = byte != 0
    nzm |= (uint32_t) (((v & 0x8080808080808080ull) >> 7) * 0x0102040810204080ull >> 56) << (8 * w);
  }
  m.nz_mask = nzm;
  mb_slice[mb_addr] = (int16_t)slice_no;
  if (mr) {
    mr->cbp = (uint8_t)cbp; mr->luma_qp = m.qp_y;
    mr->chroma_mode = st_chroma[mb_addr]; mr->i16_mode = st_i16[mb_addr];
    for (int q = 0; q < 4; q++) mr->sub_mb_type[q] = st_sub[(size_t)mb_addr * 4 + q];
    if (m.flags & LH264_MBF_TRANSFORM8x8) mr->flags |= LH264_MM_T8;
    mr->coded_mask = m.coded_mask; mr->coef_off = m.coef_off;
  }
  return br.overrun() ? -1 : 0;
}


// ---- CABAC syntax elements (9.3.2 binarisation + 9.3.3.1 ctxIdx) -----------------------------------------------------
// Replaces parse_mb_syn_cabac.cpp (ParseMBTypeISliceCabac :189, ParseMBTypePSliceCabac :227, ParseSubMBTypeCabac :301,
// ParseIntraPredModeLumaCabac :322, ParseIntraPredModeChromaCabac :346, ParseInterMotionInfoCabac :393, ParseRefIdxCabac :600,
// ParseMvdInfoCabac :651, ParseCbpInfoCabac :687, ParseDeltaQpCabac :760, ParseCbfInfoCabac :792, ParseSignificantMapCabac
// :844, ParseSignificantCoeffCabac :889, ParseResidualBlockCabac :929-1040, ParseIPCMInfoCabac :1044) on the records of
// this decoder.  Only what frame macroblocks of I and P slices need (no MBAFF, no B slices: as the reference).
namespace {
const uint8_t kSig8x8[63] = {0, 1, 2, 3, 4, 5, 5, 4, 4, 3, 3, 4, 4, 4, 5, 5, 4, 4, 4, 4, 3, 3, 6, 7, 7, 7, 8, 9, 10, 9, 8, 7,
                             7, 6, 11, 12, 13, 11, 6, 7, 8, 9, 14, 10, 9, 8, 6, 11, 12, 13, 11, 6, 9, 14, 10, 9, 11, 12, 13, 11, 14, 10, 12
                            };                        // Table 9-43, frame coded: significant_coeff_flag ctxIdxInc for ctxBlockCat 5
const uint8_t kLast8x8[63] = {0, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2,
                              3, 3, 3, 3, 3, 3, 3, 3, 4, 4, 4, 4, 4, 4, 4, 4, 5, 5, 5, 5, 6, 6, 6, 6, 7, 7, 7, 7, 8, 8, 8
                             };                       // Table 9-43: last_significant_coeff_flag
}

int Decoder::Impl::cabac_intra_mb_type (int base, bool i_slice, int mb_addr) {
  if (i_slice) {                                      // ctxIdxInc of bin 0: neighbours that are not I_NxN
    const int mbx = cur_mbx, mby = cur_mby;            // (mb_addr == cur_mb)
    const int nA = nb_mb (mbx, mby, -1, 0), nB = nb_mb (mbx, mby, 0, -1);
    auto cond = [&] (int n) { return mb_avail (n) && mbs[n].mb_type != LH264_MB_I4x4 && mbs[n].mb_type != LH264_MB_I8x8; };
    if (!cab.decision (base + (cond (nA) ? 1 : 0) + (cond (nB) ? 1 : 0))) return 0;
    base += 2;
  } else if (!cab.decision (base)) return 0;
  if (cab.terminate()) return 25;
  int t = 1 + 12 * cab.decision (base + 1);
  if (cab.decision (base + 2)) t += 4 + 4 * cab.decision (base + 2 + (i_slice ? 1 : 0));
  t += 2 * cab.decision (base + 3 + (i_slice ? 1 : 0));
  t += cab.decision (base + 3 + (i_slice ? 2 : 0));
  return t;
}

int Decoder::Impl::cabac_mb_type (int mb_addr) {       // value of mb_type as Tables 7-11 / 7-13 number it (P: intra types follow 5)
  if (sh.slice_type == SLICE_I) return cabac_intra_mb_type (3, true, mb_addr);
  if (!cab.decision (14)) {
    if (!cab.decision (15)) return 3 * cab.decision (16);
    return 2 - cab.decision (17);
  }
  return 5 + cabac_intra_mb_type (17, false, mb_addr);
}

int Decoder::Impl::cabac_sub_mb_type() {
  if (cab.decision (21)) return 0;
  if (!cab.decision (22)) return 1;
  return cab.decision (23) ? 2 : 3;
}

int Decoder::Impl::cabac_ref_idx (int mb_addr, int q, const int8_t* cur_refs) {
  const int mbx = cur_mbx, mby = cur_mby;            // (mb_addr == cur_mb)
  auto gt0 = [&] (int n, int nq) { return mb_avail (n) && !LH264_MB_IS_INTRA (mbs[n].mb_type) && mbs[n].ref_idx[nq] > 0; };
  const bool a = (q & 1) ? cur_refs[q - 1] > 0 : gt0 (nb_mb (mbx, mby, -1, 0), q + 1);
  const bool b = (q & 2) ? cur_refs[q - 2] > 0 : gt0 (nb_mb (mbx, mby, 0, -1), q + 2);
  int ctx = (a ? 1 : 0) + (b ? 2 : 0), ref = 0;
  while (cab.decision (54 + ctx)) {
    ref++;
    ctx = (ctx >> 2) + 4;
    if (ref >= 32) return -1;
  }
  return ref;
}

int Decoder::Impl::cabac_mvd (int mb_addr, int bx, int by, int comp) {
  const int mbx = cur_mbx, mby = cur_mby;            // (mb_addr == cur_mb)
  int sum = 0;
  if (bx > 0) sum += mvd_abs[(size_t)mb_addr * 32 + 2 * (by * 4 + bx - 1) + comp];
  else { const int n = nb_mb (mbx, mby, -1, 0); if (mb_avail (n)) sum += mvd_abs[(size_t)n * 32 + 2 * (by * 4 + 3) + comp]; }
  if (by > 0) sum += mvd_abs[(size_t)mb_addr * 32 + 2 * ((by - 1) * 4 + bx) + comp];
  else { const int n = nb_mb (mbx, mby, 0, -1); if (mb_avail (n)) sum += mvd_abs[(size_t)n * 32 + 2 * (12 + bx) + comp]; }
  int ctx = (comp ? 47 : 40);
  if (!cab.decision (ctx + (sum < 3 ? 0 : (sum > 32 ? 2 : 1)))) return 0;
  int v = 1;
  ctx += 3;
  while (v < 9 && cab.decision (ctx)) { if (v < 4) ctx++; v++; }
  if (v >= 9) {                                       // UEG3 suffix
    int k = 3;
    while (cab.bypass()) { v += 1 << k; if (++k > 24) return 0; }
    while (k--) v += cab.bypass() << k;
  }
  return cab.bypass() ? -v : v;
}

int Decoder::Impl::cabac_cbp (int mb_addr) {
  const int mbx = cur_mbx, mby = cur_mby;            // (mb_addr == cur_mb)
  const int nA = nb_mb (mbx, mby, -1, 0), nB = nb_mb (mbx, mby, 0, -1);
  // unavailable neighbours count as "all luma blocks coded, no chroma"
  const int ca = mb_avail (nA) ? cabmb[nA].cbp : 0x0f, cb = mb_avail (nB) ? cabmb[nB].cbp : 0x0f;
  int cbp = 0;
  cbp |= cab.decision (73 + ! (ca & 2) + 2 * ! (cb & 4));
  cbp |= cab.decision (73 + ! (cbp & 1) + 2 * ! (cb & 8)) << 1;
  cbp |= cab.decision (73 + ! (ca & 8) + 2 * ! (cbp & 1)) << 2;
  cbp |= cab.decision (73 + ! (cbp & 4) + 2 * ! (cbp & 2)) << 3;
  const int xa = ca >> 4, xb = cb >> 4;
  if (cab.decision (77 + (xa > 0 ? 1 : 0) + (xb > 0 ? 2 : 0)))
    cbp |= (1 + cab.decision (77 + 4 + (xa == 2 ? 1 : 0) + (xb == 2 ? 2 : 0))) << 4;
  return cbp;
}

int Decoder::Impl::cabac_chroma_mode (int mb_addr) {
  const int mbx = cur_mbx, mby = cur_mby;            // (mb_addr == cur_mb)
  auto cond = [&] (int n) { return mb_avail (n) && LH264_MB_IS_INTRA (mbs[n].mb_type) && mbs[n].mb_type != LH264_MB_IPCM && cabmb[n].chroma != 0; };
  if (!cab.decision (64 + (cond (nb_mb (mbx, mby, -1, 0)) ? 1 : 0) + (cond (nb_mb (mbx, mby, 0, -1)) ? 1 : 0))) return 0;
  if (!cab.decision (64 + 3)) return 1;
  return 2 + cab.decision (64 + 3);
}

int Decoder::Impl::cabac_dqp() {
  int ctx = last_dqp_nz ? 1 : 0, v = 0;
  while (cab.decision (60 + ctx)) {
    ctx = 2 + (ctx >> 1);
    if (++v > 104) return 99;                         // out of range: rejected by the caller
  }
  return (v & 1) ? (v + 1) >> 1 : - (v >> 1);
}

int Decoder::Impl::cabac_t8_flag (int mb_addr) {
  const int mbx = cur_mbx, mby = cur_mby;            // (mb_addr == cur_mb)
  auto cond = [&] (int n) { return mb_avail (n) && (mbs[n].flags & LH264_MBF_TRANSFORM8x8); };
  return cab.decision (399 + (cond (nb_mb (mbx, mby, -1, 0)) ? 1 : 0) + (cond (nb_mb (mbx, mby, 0, -1)) ? 1 : 0));
}

// 9.3.3.1.1.9: ctxIdxInc of coded_block_flag.  bit_cur_* >= 0: the neighbouring block lies in the current macroblock (that bit
// of its cbf word); otherwise the bit of the left / top macroblock's word.  Unavailable macroblock: 1 for intra, 0 for inter.
int Decoder::Impl::cbf_inc (int mb_addr, int bit_cur_a, int bit_left, int bit_cur_b, int bit_top) const {
  const int mbx = cur_mbx, mby = cur_mby;            // (mb_addr == cur_mb)
  const int dflt = LH264_MB_IS_INTRA (mbs[mb_addr].mb_type) ? 1 : 0;
  int a, b;
  if (bit_cur_a >= 0) a = (cabmb[mb_addr].cbf >> bit_cur_a) & 1;
  else { const int n = nb_mb (mbx, mby, -1, 0); a = mb_avail (n) ? (int) ((cabmb[n].cbf >> bit_left) & 1) : dflt; }
  if (bit_cur_b >= 0) b = (cabmb[mb_addr].cbf >> bit_cur_b) & 1;
  else { const int n = nb_mb (mbx, mby, 0, -1); b = mb_avail (n) ? (int) ((cabmb[n].cbf >> bit_top) & 1) : dflt; }
  return a + 2 * b;
}

// residual_block_cabac( ) (7.3.5.3.3).  cat = ctxBlockCat 0..5; cbf_inc < 0: no coded_block_flag is read (8x8 blocks).  Level of
// coefficient i goes to out[scan[scan_off + i]]; returns the number of non-zero coefficients.
int Decoder::Impl::cabac_block (int cat, int cbf_ctx_inc, int16_t* out, const uint8_t* scan, int scan_off) {
  static const int kCbfBase[5] = {85, 85 + 4, 85 + 8, 85 + 12, 85 + 16};
  static const int kSigBase[6] = {105, 105 + 15, 105 + 29, 105 + 44, 105 + 47, 402};
  static const int kLastBase[6] = {166, 166 + 15, 166 + 29, 166 + 44, 166 + 47, 417};
  static const int kAbsBase[6] = {227, 227 + 10, 227 + 20, 227 + 30, 227 + 39, 426};
  static const int kMax[6] = {16, 15, 16, 4, 15, 64};
  if (cbf_ctx_inc >= 0 && !cab.decision (kCbfBase[cat] + cbf_ctx_inc)) return 0;
  const int n_max = kMax[cat];
  uint8_t pos[64];
  int n = 0, i;
  for (i = 0; i < n_max - 1; i++) {
    const int inc_s = cat == 5 ? kSig8x8[i] : (cat == 3 ? std::min (i, 2) : i);
    if (!cab.decision (kSigBase[cat] + inc_s)) continue;
    pos[n++] = (uint8_t)i;
    const int inc_l = cat == 5 ? kLast8x8[i] : (cat == 3 ? std::min (i, 2) : i);
    if (cab.decision (kLastBase[cat] + inc_l)) break;
  }
  if (i == n_max - 1) pos[n++] = (uint8_t) (n_max - 1);   // the last position is significant by inference
  int eq1 = 0, gt1 = 0;
  const int gt1_cap = cat == 3 ? 3 : 4;
  for (int k = n - 1; k >= 0; k--) {
    int a = 1;
    if (cab.decision (kAbsBase[cat] + (gt1 ? 0 : std::min (4, 1 + eq1)))) {
      const int ctx = kAbsBase[cat] + 5 + std::min (gt1_cap, gt1);
      a = 2;
      while (a < 15 && cab.decision (ctx)) a++;
      if (a >= 15) {                                  // Exp-Golomb (k = 0) suffix, bypass coded
        int e = 0;
        while (cab.bypass()) { a += 1 << e; if (++e > 20) return -1; }
        while (e--) a += cab.bypass() << e;
      }
      gt1++;
    } else eq1++;
    out[scan[scan_off + pos[k]]] = (int16_t) (cab.bypass() ? -a : a);
  }
  return n;
}

int Decoder::Impl::parse_residual_cabac (lh264_mb_rec& m, int mb_addr, int16_t* lev, uint32_t& coded, bool i16, int cbp) {
  uint8_t* nz = &nzc[(size_t)mb_addr * 24];
  uint32_t& cbf = cabmb[mb_addr].cbf;
  const bool t8 = m.flags & LH264_MBF_TRANSFORM8x8;
  if (i16) {                                          // Intra16x16DCLevel
    int16_t dc[16];
    memset (dc, 0, sizeof (dc));
    const int tc = cabac_block (0, cbf_inc (mb_addr, -1, 24, -1, 24), dc, kZigzag4x4, 0);
    if (tc < 0) return -1;
    if (tc) cbf |= 1u << 24;
    for (int r = 0; r < 16; r++) if (dc[r]) { const int z = kRtoZ[r]; lev[16 * z] = dc[r]; coded |= 1u << z; }
  }
  for (int q = 0; q < 4; q++) {
    if (! (cbp & (1 << q))) { uint8_t* z = nz + (q >> 1) * 8 + (q & 1) * 2; z[0] = z[1] = z[4] = z[5] = 0; continue; }
    if (t8) {
      const int tc = cabac_block (5, -1, lev + 64 * q, kZigzag8x8, 0);
      if (tc < 0) return -1;
      for (int k = 0; k < 4; k++) { const int r = kZtoR[4 * q + k]; nz[r] = (uint8_t)tc; cbf |= 1u << r; }   // coded_block_flag inferred 1
      if (tc) coded |= 0xFu << (4 * q);
      continue;
    }
    for (int k = 0; k < 4; k++) {
      const int z = 4 * q + k, r = kZtoR[z], bx = r & 3, by = r >> 2;
      const int inc = cbf_inc (mb_addr, bx > 0 ? r - 1 : -1, r + 3, by > 0 ? r - 4 : -1, r + 12);
      const int tc = i16 ? cabac_block (1, inc, lev + 16 * z, kZigzag4x4, 1) : cabac_block (2, inc, lev + 16 * z, kZigzag4x4, 0);
      if (tc < 0) return -1;
      nz[r] = (uint8_t)tc;
      if (tc) { coded |= 1u << z; cbf |= 1u << r; }
    }
  }
  const int cc = cbp >> 4;
  if (cc) {
    for (int p = 0; p < 2; p++) {                     // ChromaDCLevel
      int16_t dc[64];
      memset (dc, 0, sizeof (dc));
      const int tc = cabac_block (3, cbf_inc (mb_addr, -1, 25 + p, -1, 25 + p), dc, kChromaDcScan, 0);
      if (tc < 0) return -1;
      if (tc) cbf |= 1u << (25 + p);
      for (int b = 0; b < 4; b++) if (dc[16 * b]) { lev[16 * (16 + 4 * p + b)] = dc[16 * b]; coded |= 1u << (16 + 4 * p + b); }
    }
  }
  for (int p = 0; p < 2; p++) for (int b = 0; b < 4; b++) {
      const int bit = 16 + 4 * p + b;
      if (cc & 2) {                                   // ChromaACLevel
        const int inc = cbf_inc (mb_addr, (b & 1) ? bit - 1 : -1, bit + 1, (b & 2) ? bit - 2 : -1, bit + 2);
        const int tc = cabac_block (4, inc, lev + 16 * bit, kZigzag4x4, 1);
        if (tc < 0) return -1;
        nz[bit] = (uint8_t)tc;
        if (tc) { coded |= 1u << bit; cbf |= 1u << bit; }
      } else nz[bit] = 0;
    }
  return 0;
}

// ---- slice_data( ) (7.3.4), CABAC ----------------------------------------------------------------------------
int Decoder::Impl::parse_slice_data_cabac (BitReader& br, Decoder* self) {
  int mb_addr = sh.first_mb;
  int n_done = 0;
  qp = sh.slice_qp;
  br.align();                                         // cabac_alignment_one_bit
  cab.init_contexts (sh.slice_type == SLICE_I ? 0 : 1 + sh.cabac_init_idc, sh.slice_qp);
  cab.start (&br);
  cabac_on = true;
  last_dqp_nz = false;
  int rc = 0;
  // model records: the reference codes a one-macroblock "skip run" (0 or 1) in front of every macroblock of a CABAC slice and
  // takes the stop bit from end_of_slice_flag (WelsDecodeSliceForNonRecoding decode_slice.cpp:2197-2226, WelsDecodeMbCabacPSlice
  // :1164-1228)
  const size_t rec0 = mrecs.size();
  int m_cached_qp = 0, m_last_dqp = 0;
  while (mb_addr < n_mbs) {
    bool skip = false;
    if (sh.slice_type != SLICE_I) {
      const int mbx = mb_addr % mb_w, mby = mb_addr / mb_w;
      auto cond = [&] (int n) { return mb_avail (n) && mbs[n].mb_type != LH264_MB_PSKIP; };
      st_cbp[mb_addr] = 0; st_chroma[mb_addr] = 0;    // pCbp / pChromaPredMode are reset before the skip flag is read (:1172-1174)
      skip = cab.decision (11 + (cond (nb_mb (mbx, mby, -1, 0)) ? 1 : 0) + (cond (nb_mb (mbx, mby, 0, -1)) ? 1 : 0));
    }
    lh264_model_mb* mr = nullptr;
    if (want_model) {
      mr = mrecs.grow (1);
      if (!mr) { rc = -1; break; }
      memset (mr, 0, sizeof (*mr));
      mr->mb_addr = (uint32_t)mb_addr; mr->mb_in_slice = (uint16_t)std::min<size_t> (mrecs.size() - 1 - rec0, 65535);
      mr->flags |= LH264_MM_EMIT_SKIP_RUN | (skip ? LH264_MM_STOP_IN_SKIP : LH264_MM_WRITE_BLOCK);
      mr->skip_run = skip ? 1 : 0;
    }
    if (parse_mb (br, mb_addr, skip, mr) < 0 || cab.overrun()) { rc = -1; self->stats_.errors++; if (mr) mrecs.pop_back(); break; }
    if (mr && !skip) {
      const int d = (int)mr->luma_qp - m_cached_qp;
      mr->qp_delta = (int8_t)d;
      mr->flags |= (uint16_t) ((m_last_dqp < 0 ? 0 : (m_last_dqp == 0 ? 1 : 2)) << LH264_MM_LASTDQP_SHIFT);
      m_cached_qp = mr->luma_qp;
      if (d) m_last_dqp = d;
    }
    mb_addr = next_mb (mb_addr); n_done++;
    const bool eos = cab.terminate();                 // end_of_slice_flag
    if (mr && eos) mr->flags |= LH264_MM_STOP_BIT;
    if (eos) break;
  }
  cabac_on = false;
  lh264_slice_rec& sr = slices.back();
  sr.num_mbs = n_done;
  decoded += n_done;
  self->stats_.mbs += n_done;
  if (want_model) {
    lh264_model_slice ms;
    memset (&ms, 0, sizeof (ms));
    ms.first_rec = (int32_t)rec0; ms.n_recs = (int32_t) (mrecs.size() - rec0);
    ms.frame_num = sh.frame_num;
    ms.slice_type = (uint8_t) (sh.slice_type == SLICE_I ? 2 : 0);
    ms.num_ref_idx_l0 = (uint8_t) (sh.slice_type == SLICE_P ? sh.num_ref_idx_l0 : P->num_ref_idx_l0_default);
    // the reference hands its bit reader back byte aligned (RestoreCabacDecEngineToBS), so `7 - (iLeftBits & 7)` is always 7:
    // the low seven bits of the slice's last byte go to the PADBYTE stream, whatever the arithmetic decoder consumed of them
    ms.n_pad_bits = 7;
    ms.pad_bits = (uint8_t) (br.size_bits() ? (br.data()[br.payload_end() >> 3] & 0x7fu) : 0);
    ms.chroma_format_idc = 1;
    mslices.push_back (ms);
  }
  return rc;
}

// ---- slice_data( ) (7.3.4), CAVLC --------------------------------------------------------------------------
LH264_HOT int Decoder::Impl::parse_slice_data (BitReader& br, Decoder* self) {
  int mb_addr = sh.first_mb;
  qp = sh.slice_qp;
  const int first = mb_addr;
  int n_done = 0;
  const size_t data_begin = br.pos();
  bool more = true;
  int rc = 0;
  // model records: one per MB in slice-data order, with the reference's skip-run / stop-bit bookkeeping
  // (WelsDecodeMbCavlcPSlice decode_slice.cpp:3880-3960, WelsDecodeSliceForNonRecoding :2192-2252)
  const size_t rec0 = mrecs.size();
  alignas (16) lh264_model_mb rec_local;              // assembled here, streamed out when the macroblock is done
  auto new_rec = [&] (int addr) -> lh264_model_mb* {
    if (!want_model) return nullptr;
    memset (&rec_local, 0, sizeof (rec_local));
    rec_local.mb_addr = (uint32_t)addr; rec_local.mb_in_slice = (uint16_t)std::min<size_t> (mrecs.size() - rec0, 65535);
    return &rec_local;
  };
  auto commit_rec = [&] () -> bool {
    lh264_model_mb* dst = mrecs.grow (1);
    if (!dst) return false;
    if (stream_mrecs) stream_copy (dst, &rec_local, sizeof (rec_local)); else memcpy (dst, &rec_local, sizeof (rec_local));
    return true;
  };
  int m_cached_qp = 0, m_last_dqp = 0;                // uiCachedLumaQp / uiLastNonzeroDeltaLumaQp, decode_slice.cpp:3061-3062
  while (more && mb_addr < n_mbs) {
    bool run_read = false;
    if (sh.slice_type != SLICE_I) {
      const uint32_t run = br.ue();
      if (run > (uint32_t) (n_mbs - mb_addr)) { rc = -1; break; }
      run_read = true;
      for (uint32_t i = 0; i < run; i++) {
        lh264_model_mb* mr = new_rec (mb_addr);
        parse_mb (br, mb_addr, true, mr);
        if (mr) {
          if (i == 0) { mr->flags |= LH264_MM_EMIT_SKIP_RUN; mr->skip_run = (uint16_t)std::min<uint32_t> (run, 65535); }
          if (i + 1 == run) { mr->flags |= LH264_MM_STOP_IN_SKIP; if (!br.more_rbsp_data()) mr->flags |= LH264_MM_STOP_BIT; }
          if (!commit_rec()) { rc = -1; break; }
        }
        mb_addr = next_mb (mb_addr); n_done++;
        if (mb_addr >= n_mbs && i + 1 < run) { rc = -1; break; }      // the run leaves the slice group
      }
      if (rc < 0) break;
      if (run > 0) { more = br.more_rbsp_data(); run_read = false; }
      if (!more || mb_addr >= n_mbs) break;
    }
    lh264_model_mb* mr = new_rec (mb_addr);
    if (mr) {
      mr->flags |= LH264_MM_WRITE_BLOCK;
      if (run_read || sh.slice_type == SLICE_I) mr->flags |= LH264_MM_EMIT_SKIP_RUN;   // value 0
    }
    if (parse_mb (br, mb_addr, false, mr) < 0) { rc = -1; self->stats_.errors++; break; }       // (its record is dropped)
    if (mr) {                                         // decode_slice.cpp:2267-2274
      const int d = (int)mr->luma_qp - m_cached_qp;
      mr->qp_delta = (int8_t)d;
      mr->flags |= (uint16_t) ((m_last_dqp < 0 ? 0 : (m_last_dqp == 0 ? 1 : 2)) << LH264_MM_LASTDQP_SHIFT);
      m_cached_qp = mr->luma_qp;
      if (d) m_last_dqp = d;
    }
    mb_addr = next_mb (mb_addr); n_done++;
    more = br.more_rbsp_data();
    if (mr && !more) mr->flags |= LH264_MM_STOP_BIT;
    if (mr && !commit_rec()) { rc = -1; break; }
  }
  lh264_slice_rec& sr = slices.back();
  sr.num_mbs = n_done;
  decoded += n_done;
  self->stats_.mbs += n_done;
  if (want_model) {
    lh264_model_slice ms;
    memset (&ms, 0, sizeof (ms));
    ms.first_rec = (int32_t)rec0; ms.n_recs = (int32_t) (mrecs.size() - rec0);
    ms.frame_num = sh.frame_num;
    ms.slice_type = (uint8_t) (sh.slice_type == SLICE_I ? 2 : 0);
    ms.num_ref_idx_l0 = (uint8_t) (sh.slice_type == SLICE_P ? sh.num_ref_idx_l0 : P->num_ref_idx_l0_default);
    // alignment bits after the stop bit (decode_slice.cpp:3133-3147): 7 - (bits used & 7), taken from the byte that
    // holds the stop bit
    const size_t used = br.pos();
    ms.n_pad_bits = (uint8_t) (7 - (used & 7));
    ms.pad_bits = (uint8_t) (br.size_bits() ? (br.data()[br.payload_end() >> 3] & ((1u << ms.n_pad_bits) - 1)) : 0);
    ms.chroma_format_idc = 1;
    mslices.push_back (ms);
    if (check_rewrite && rc == 0) {
      if (mslices.size() == 1 || !rewriter) rewriter.reset (new PictureSyntaxWriter (mb_w, mb_h));
      SliceWriteParams wp;
      wp.slice_type = ms.slice_type; wp.num_ref_idx_l0 = ms.num_ref_idx_l0; wp.slice_qp = sh.slice_qp; wp.transform_8x8_mode = P->transform_8x8_mode;
      rewriter->begin_slice (wp);
      BitWriter bw;
      bool ok = true;
      for (size_t i = rec0; i < mrecs.size() && ok; i++) ok = rewriter->write_mb (bw, mrecs[i], coefs.data());
      const size_t n = br.pos() - data_begin;
      ok = ok && bw.bits() == n;
      for (size_t i = 0; ok && i < n; i++) {
        const size_t a = data_begin + i;
        ok = (((br.data()[a >> 3] >> (7 - (a & 7))) ^ (bw.bytes()[i >> 3] >> (7 - (i & 7)))) & 1) == 0;
      }
      self->stats_.rewritten_slices++;
      if (!ok) {
        self->stats_.rewrite_mismatches++;
        if (debug_trace()) fprintf (stderr, "rewrite mismatch: slice first_mb %d type %d: %zu bits parsed, %zu written\n", first, ms.slice_type, n, bw.bits());
      }
    }
  }
  return rc;
}

// ---- Decoder ---------------------------------------------------------------------------------------------------
Decoder::Decoder (PictureSink* sink) : d_ (new Impl()) { d_->sink = sink; cavlc_init(); }
Decoder::~Decoder() {}
const std::vector<ModelMbSide>& Decoder::model_side() const { return d_->side; }
void Decoder::enable_model (bool on) { d_->want_model = on; }
void Decoder::enable_rewrite_check (bool on) { d_->check_rewrite = on; if (on) d_->want_model = true; }

int Decoder::flush() { d_->by_flush = true; const int rc = d_->finish_picture (this); d_->by_flush = false; return rc; }
void Decoder::set_nodelay_output (bool on) { d_->nodelay_output = on; }
bool Decoder::picture_open() const { return d_->open; }
void Decoder::set_freeze_output (bool on) { d_->freeze_output = on; }
// InitErrorCon (error_concealment.cpp:43-52): the three non-freezing slice modes clear bFreezeOutput, nothing sets it here
bool Decoder::open_picture_error() const {
  const Impl& D = *d_;
  if (!D.open || D.sh0.idr) return false;
  if (D.ref_lost) return true;
  if (!D.slot_incomplete || D.slices.empty()) return false;
  const lh264_slice_rec& sl = D.slices.back();
  if (sl.slice_type != SLICE_P) return false;
  for (int i = 0; i < sl.num_ref_idx_l0 && i < LH264_MAX_REFS; i++) if ((D.slot_incomplete >> (sl.ref_slot[i] & 31)) & 1u) return true;
  return false;
}
void Decoder::set_ec_mode (int idc, bool at_initialize) {
  if (at_initialize) d_->freeze = true;                // WelsDecoderDefaults, decoder.cpp:300
  d_->ec_mode = idc;
  if (idc == 2 || idc == 4 || idc == 6) d_->freeze = false;
}
void Decoder::set_streaming_stores (bool coefs, bool model_recs) { d_->stream_coefs = coefs; d_->stream_mrecs = model_recs; }
bool Decoder::set_arena (RecordArena* a) {
  const bool ok = d_->mbs.set_arena (a) & d_->coefs.set_arena (a) & d_->mrecs.set_arena (a);
  if (!ok) d_->alloc_failed = true;
  return ok;
}

int Decoder::decode (const uint8_t* data, size_t len, unsigned long long timestamp) {
  Impl& D = *d_;
  std::vector<NalSpan> nals;
  split_annexb (data, len, nals);
  int last_err = 0;
  last_slice_ = LastSlice();
  for (const NalSpan& n : nals) {
    stats_.nals++;
    if (n.len < 1) continue;
    const int ref_idc = (n.p[0] >> 5) & 3, type = n.p[0] & 31;
    const size_t plen = unescape_nal (n.p + 1, n.len - 1, D.rbsp);
    BitReader br (D.rbsp.data(), plen);
    if (type == NAL_SPS) {
      Sps s;
      if (parse_sps (br, s)) {
        if (D.open && D.active_sps == s.id) D.finish_picture (this);
        D.sps[s.id] = s;
        sps_seen_ = true;
      } else { last_err = -10; stats_.errors++; }
    } else if (type == NAL_PPS) {
      Pps p;
      if (parse_pps (br, p, D.sps)) { D.pps[p.id] = p; if (p.scaling_matrix_present && p.sps_was_missing) D.sps_late_pps = true; } else { last_err = -11; stats_.errors++; }
    } else if (type == NAL_SLICE || type == NAL_IDR) {
      SliceHeader sh;
      const int rc = parse_slice_header (br, ref_idc, type, D.sps, D.pps, sh);
      if (rc < 0) { last_err = rc; stats_.errors++; continue; }
      const Pps& pps = D.pps[sh.pps_id];
      const Sps& sps = D.sps[pps.sps_id];
      if (D.open && D.is_new_picture (D.sh0, sh)) D.finish_picture (this);
      if (!D.open) {
        if (D.activate (sps) < 0) { last_err = -30; continue; }
        D.S = &D.sps[pps.sps_id]; D.P = &D.pps[sh.pps_id]; D.sh = sh; D.ts = timestamp;
        D.start_picture (this);
        if (!D.open) { last_err = -31; stats_.errors++; continue; }        // no storage for the picture's records
      } else { D.S = &D.sps[pps.sps_id]; D.P = &D.pps[sh.pps_id]; D.sh = sh; D.ts = timestamp; }   // (a picture carries the time stamp of its last slice's call)
      if (sh.first_mb >= D.n_mbs) { last_err = -1; stats_.errors++; continue; }
      if (pps.num_slice_groups > 7) { last_err = -4; stats_.errors++; continue; }   // FmoGenerateMbAllocMapType0/1 refuse 8 groups (fmo.cpp:61, :92)
      D.build_fmo_map (pps);
      D.slice_no = (int)D.slices.size();
      // slice record
      lh264_slice_rec sr;
      memset (&sr, 0, sizeof (sr));
      sr.first_mb = sh.first_mb;
      sr.slice_type = (uint8_t)sh.slice_type;
      sr.disable_deblocking_idc = (uint8_t)sh.disable_deblocking_idc;
      sr.alpha_c0_offset = (int8_t)sh.alpha_c0_offset; sr.beta_offset = (int8_t)sh.beta_offset;
      sr.num_ref_idx_l0 = (uint8_t)sh.num_ref_idx_l0;
      if (sh.slice_type == SLICE_P) {
        // No reference picture at all (the stream's IDR picture was lost): the reference puts a fresh grey picture with frame_num 0
        // into its short-term list and predicts from it (WelsInitRefList, manage_dec_ref.cpp:112-150; the previous picture is
        // copied instead only when there is one, which a stream without any reference does not have).  The slots of a stream
        // start out grey (pic_queue.cpp:96); one that has been written since is filled again by the sink.
        if (D.refs.empty()) {
          int gs = -1;
          for (int s2 = 0; s2 < D.n_slots && gs < 0; s2++) if (s2 != D.cur_slot) gs = s2;
          if (gs >= 0) {
            RefPic r;
            r.slot = gs; r.frame_num = 0;
            D.refs.push_back (r);
            if ((D.slots_touched >> gs) & 1u) D.sink->fill_slot (gs, 128);
          }
        }
        D.build_list0();
        if (debug_trace()) { fprintf(stderr, "slice fn=%d nref=%d refs=%zu mods:", sh.frame_num, sh.num_ref_idx_l0, D.refs.size()); for (auto& mm : sh.mods) fprintf(stderr, " (%d,%u)", mm.op, mm.val); fprintf(stderr, " | refs:"); for (auto& r : D.refs) fprintf(stderr, " [slot %d fn %d long %d idx %d]", r.slot, r.frame_num, r.is_long, r.long_idx); fprintf(stderr, " | mmco:"); for (auto& mm : sh.mmcos) fprintf(stderr, " (%d,%u,%u)", mm.op, mm.a, mm.b); fprintf(stderr, "\n"); }
        const int fallback = D.list0[0] ? D.list0[0]->slot : (D.refs.empty() ? 0 : D.refs.back().slot);
        for (int i = 0; i < LH264_MAX_REFS; i++) sr.ref_slot[i] = (int16_t) (D.list0[i] ? D.list0[i]->slot : fallback);
        // explicit WP is "on" in the reference only when some weight differs from the default
        // (bUseWeightPredictionFlag); identical results either way, so keep it simple: on iff signalled
        if (sh.has_wp) {
          sr.weighted_pred = 1;
          sr.luma_log2_denom = (uint8_t)sh.luma_log2_denom; sr.chroma_log2_denom = (uint8_t)sh.chroma_log2_denom;
          for (int i = 0; i < LH264_MAX_REFS && i < sh.num_ref_idx_l0; i++) {
            sr.luma_weight[i] = (int16_t)sh.luma_weight[i]; sr.luma_offset[i] = (int16_t)sh.luma_offset[i];
            for (int c = 0; c < 2; c++) { sr.chroma_weight[i][c] = (int16_t)sh.chroma_weight[i][c]; sr.chroma_offset[i][c] = (int16_t)sh.chroma_offset[i][c]; }
          }
        }
      }
      // Scaling lists in force (WelsCalcDeqCoeffScalingList, decode_slice.cpp:1231-1277, called once per slice): the reference
      // rebuilds its multiplier tables only when the PPS id differs from the one they were built for, so a slice keeps the
      // lists of the FIRST parameter sets seen under that id until another id comes by - reproduced here.
      if (sps.scaling_matrix_present || pps.scaling_matrix_present) {
        if (!D.scal_init || D.scal_ppsid != pps.id) {
          Pps& wp = D.pps[sh.pps_id];
          if (D.sps_late_pps)                        // a PPS parsed before its SPS: absent lists are patched from the SPS now
            for (int i = 0; i < 12; i++) if (!wp.list_present[i]) {
                if (i < 6) memcpy (wp.scale4[i], (i == 0 || i == 3) ? sps.scale4[i] : wp.scale4[i - 1], 16);
                else memcpy (wp.scale8[i - 6], (i == 6 || i == 7) ? sps.scale8[i - 6] : wp.scale8[i - 8], 64);
              }
          for (int i = 0; i < 6; i++) memcpy (D.scal4[i], wp.scaling_matrix_present ? wp.scale4[i] : sps.scale4[i], 16);
          for (int i = 0; i < 2; i++) memcpy (D.scal8[i], wp.scaling_matrix_present ? wp.scale8[i] : sps.scale8[i], 64);
          D.scal_init = true; D.scal_ppsid = pps.id;
        }
        sr.scaling_flags = (uint8_t) (LH264_SLF_SCALING | (pps.cabac ? 0 : LH264_SLF_SCALING_CAVLC));
        memcpy (sr.scale4, D.scal4, sizeof (sr.scale4));
        memcpy (sr.scale8, D.scal8, sizeof (sr.scale8));
      }
      D.slices.push_back (sr);
      stats_.slices++;
      br.seek (sh.header_bits);
      D.slice_rbsp.assign (D.rbsp.begin(), D.rbsp.begin() + std::min (D.rbsp.size(), (sh.header_bits + 7) / 8));   // for last_slice()
      last_slice_.decoded = true; last_slice_.cabac = pps.cabac; last_slice_.rbsp = D.slice_rbsp.data(); last_slice_.header_bits = sh.header_bits;
      last_slice_.idr = sh.idr; last_slice_.idr_pic_id = sh.idr_pic_id; last_slice_.frame_num = sh.frame_num;
      if ((pps.cabac ? D.parse_slice_data_cabac (br, this) : D.parse_slice_data (br, this)) < 0) last_err = -40;
      if (D.decoded >= D.n_mbs) D.finish_picture (this);
    } else if (type == NAL_END_SEQ || type == NAL_END_STREAM) {
      D.finish_picture (this);
    }
  }
  return last_err;
}

} // namespace lh264host
