// runtime.cpp — the multi-stream runtime of include/lh264_runtime.h (lh264_rt_*): many independent Annex-B streams through the
// GPU path at once, bitstream in, pictures + `.pip` out.
//
// The reference is one stream per process and strictly single-threaded (its model state is process-global,
// compression_stream.cpp:18-30); its driver is h264dec's unit loop (h264dec.cpp:246-364).  This is the host runtime the B200
// path wants instead (SURVEY.md §7 "host thread per stream", §8f N2):
//
//   parse threads ── the unit loop of h264dec → host decoder (NAL / slice / CAVLC / CABAC parse).  Step k = the next picture
//        │           of every stream; the threads DRAW (step, stream) tickets from one counter per pipeline, so an intra
//        │           picture (several times the parse work of a P picture) delays nobody but the thread that drew it - with
//        │           streams owned by threads (w, w + T, ...) any period in the streams that divides T put all the intra
//        │           pictures of a step on one thread (measured: parse threads idle 57 % of a 1024-stream run).  A stream's
//        │           pictures are parsed in order (a ticket waits for the stream's previous one).  The hand-off records of the
//        │           tickets a thread draws for step k go into ITS pinned staging buffer k mod D (tightly packed regions: MB
//        │           records, coefficient pools, model records, slice records, model slice records), so the records of a step
//        │           leave for the device with a handful of H2D copies per thread.  A thread runs up to D - 1 steps ahead of
//        │           the device.
//   submit thread ── per pipeline: waits until every parse thread has finished step k, hands the step's pictures to the GPU
//                    with ONE lh264_recon_submit + ONE lh264_model_submit, starts the packed D2H copy of the decoded pictures,
//                    delivers the pictures of step k - 1 (callback / SHA-1) while step k runs, and releases staging buffer
//                    k mod D once its H2D copies are done.
//   at the end    ── lh264_model_finish per stream; the `.pip` default stream (host) + tag side streams (GPU) are kept for
//                    lh264_rt_pip_read().
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <deque>
#include <future>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>
#include "../../include/lh264_runtime.h"
#include "decoder.h"
#include "pip_container.h"
#include "sha1.h"
#include "../csrc/compact_records.h"
#if defined(__SSE2__)
#include <emmintrin.h>
#endif

using namespace lh264host;

namespace {

double now() { return std::chrono::duration<double> (std::chrono::steady_clock::now().time_since_epoch()).count(); }

// Pinning memory is slow (cudaHostAlloc page-locks and maps every page: well under 1 GB/s when sixteen threads ask at once), and a
// 1024-stream run stages ~8 GB.  Blocks are therefore kept for the life of the process and handed to the next run
// (lh264_rt_trim() gives them back).
class PinnedCache {
 public:
  explicit PinnedCache (bool wc) : wc_ (wc) {}
  uint8_t* get (size_t need, size_t* got) {
    {
      std::lock_guard<std::mutex> g (mu_);
      auto it = free_.lower_bound (need);
      if (it != free_.end() && it->first <= 2 * need + ((size_t)4 << 20)) { uint8_t* p = it->second; *got = it->first; free_.erase (it); return p; }
    }
    uint8_t* p = (uint8_t*) (wc_ ? lh264_host_alloc_wc (need) : lh264_host_alloc (need));
    *got = p ? need : 0;
    return p;
  }
  void put (uint8_t* p, size_t cap) { if (!p) return; std::lock_guard<std::mutex> g (mu_); free_.emplace (cap, p); }
  void trim() { std::lock_guard<std::mutex> g (mu_); for (auto& kv : free_) lh264_host_free (kv.second); free_.clear(); }
 private:
  bool wc_;
  std::mutex mu_;
  std::multimap<size_t, uint8_t*> free_;
};
PinnedCache& pinned_cache() { static PinnedCache* c = new PinnedCache (false); return *c; }   // (never destroyed: no CUDA calls at exit)
// Staging memory - what the parse threads fill and the copy engine reads - is write-combined: the parser only ever WRITES it
// (compact records and coefficient pools leave the cache-resident scratch for good), so its stores need no read-for-ownership
// and the DMA reads need not snoop sixteen cores' caches.  Measured on the pool's box, 1024 1080p streams: 3.5 k -> 5.7 k
// pictures/s with pictures coming back, 6.5 k -> 8.2 k without (LH264_RT_NO_WC=1 switches it off).  Output buffers stay cacheable:
// the client reads them.
PinnedCache& staging_cache() {
  static PinnedCache* c = new PinnedCache (getenv ("LH264_RT_NO_WC") == nullptr);
  return *c;
}

struct PinRegion {                      // growable pinned buffer, bump-allocated within a step
  uint8_t* p = nullptr; size_t cap = 0, used = 0;
  double* alloc_s = nullptr;            // where the time spent in pinned allocation is accounted
  ~PinRegion() { staging_cache().put (p, cap); }
  // returns the offset of n fresh bytes (contents so far are kept when the buffer has to grow), or (size_t)-1;
  // `room` = what the region is expected to hold in the end (pinning memory is slow: grow in few, large steps)
  size_t take (size_t n, size_t room = 0) {
    const size_t at = used, need = used + n;
    if (need > cap) {
      const double t0 = now();
      size_t ncap = std::max (std::max (need + need / 2, room), (size_t)1 << 16);
      uint8_t* q = staging_cache().get (ncap, &ncap);
      if (!q) return (size_t) -1;
      if (used) memcpy (q, p, used);                   // (a read of write-combined memory: slow, and rare - regions grow in few, large steps)
      staging_cache().put (p, cap);
      p = q; cap = ncap;
      if (alloc_s) *alloc_s += now() - t0;
    }
    used = need;
    return at;
  }
};

struct Placed {                         // one picture of a step: where its arrays sit in the thread's staging buffer
  int stream;
  size_t o_sl, o_cmb, o_xmb, o_co, o_msl, o_cmm, o_xmm;
  int n_slices, n_mbs, n_blocks, n_mslices, n_mrecs, n_xmb, n_xmm;      // (n_xmb / n_xmm: 16-byte units)
  int mb_w, mb_h, n_slots, geo_serial;
  DecodedPictureInfo info;
};

// Regions of a step buffer (pinned): compact MB records + their extra units, coefficient pools, compact model records + extra
// units, slice records, model slice records.
enum { R_CMB = 0, R_XMB, R_COEFS, R_CMM, R_XMM, R_SLICES, R_MSLICES, R_COUNT };
struct StepBuf {
  PinRegion r[R_COUNT];
  size_t expect = 1;                    // pictures a step of the owning thread usually holds (its number of streams)
  std::vector<Placed> placed;
  bool failed = false;                  // pinned allocation failed
  void reset() { for (PinRegion& x : r) x.used = 0; placed.clear(); failed = false; }
};

// Where the parser writes the records of the picture in progress (decoder.h RecordArena).  Coefficient pools go STRAIGHT into the
// pinned step buffer of the owning thread: the open picture is the last allocation of the region and grows in place.  The
// 128-byte MB records and model records go into scratch memory of the thread (one picture at a time: it stays in the core's
// cache, the parser reads its neighbours back from there) and are handed over in compact form when the picture is done
// (StageSink): what crosses host memory and PCIe is 32 + 48 bytes per macroblock instead of 256.
class StepArena : public RecordArena {
 public:
  void bind (StepBuf* b) { buf_ = b; open_ = false; }
  void picture_done() { open_ = false; }
  uint8_t* storage (int kind, size_t need, size_t* cap, const uint8_t* old, size_t keep) override {
    if (kind != COEFS) {
      std::vector<uint8_t>& sc = scratch_[kind == MBS ? 0 : 1];
      if (old && old != sc.data() && keep) {          // records arriving from the decoder's own heap (a picture left open by the previous step)
        if (sc.size() < need) sc.resize (need + need / 4);
        memcpy (sc.data(), old, keep);
      } else if (sc.size() < need) sc.resize (need + need / 2 + 4096);        // (vector::resize keeps the bytes so far)
      *cap = sc.size();
      return sc.data();
    }
    PinRegion& r = buf_->r[R_COEFS];
    const bool mine = open_ && old == r.p + start_;
    // headroom: the parser asks again only when it runs out, so hand out generously (the sink trims to the final size)
    const size_t want = need + std::max<size_t> (need / 2, (size_t)64 << 10);
    size_t kept = 0;
    if (mine) { kept = keep; r.used = start_ + keep; }           // grow in place (a region that has to move takes its bytes along)
    else { start_ = r.used; open_ = true; }                      // a new picture, or records arriving from the heap
    const size_t room = start_ + want * buf_->expect + want * buf_->expect / 8;      // room for the thread's whole step
    if (want > kept && r.take (want - kept, room) == (size_t) -1) { buf_->failed = true; return nullptr; }
    uint8_t* q = r.p + start_;
    if (!mine && old && keep) memcpy (q, old, keep);
    *cap = want;
    return q;
  }
 private:
  StepBuf* buf_ = nullptr;
  size_t start_ = 0;
  bool open_ = false;
  std::vector<uint8_t> scratch_[2];
};

inline void stream16 (void* dst, const void* src, size_t bytes) {   // non-temporal copy, multiples of 16 bytes to 16-byte aligned memory
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

// Receives the pictures a Decoder completes: compacts their MB / model records into the step buffer (lh264_cmb / lh264_cmm,
// include/lh264_b200.h; csrc/compact_records.h), notes where the coefficient pool sits (parsed in place) and copies the small
// slice records behind them.
class StageSink : public PictureSink {
 public:
  int stream = 0;
  int mb_w = 0, mb_h = 0, n_slots = 0, geo_serial = 0;
  StepBuf* buf = nullptr;               // where pictures completed now are placed
  int placed_now = 0;                   // pictures placed since the owner last looked
  bool verify = false;                  // LH264_RT_VERIFY_COMPACT: expand every compact record again and compare with its source
  int configure (int w, int h, int slots) override { mb_w = w; mb_h = h; n_slots = slots; geo_serial++; return 0; }
  int submit (const lh264_picture_desc& pic, const DecodedPictureInfo& info) override {
    Placed pl = Placed();
    pl.stream = stream; pl.info = info;
    pl.n_slices = pic.n_slices; pl.n_mbs = mb_w * mb_h; pl.n_blocks = pic.n_coef_blocks;
    pl.mb_w = mb_w; pl.mb_h = mb_h; pl.n_slots = n_slots; pl.geo_serial = geo_serial;
    const size_t b_sl = (size_t)pic.n_slices * sizeof (lh264_slice_rec), b_co = (size_t)pic.n_coef_blocks * 32, n = (size_t)pl.n_mbs;
    PinRegion& rc = buf->r[R_COEFS];
    pl.o_sl = buf->r[R_SLICES].take (b_sl);
    if (b_co && inside (rc, pic.coefs)) { pl.o_co = (size_t) ((const uint8_t*)pic.coefs - rc.p); rc.used = pl.o_co + b_co; }      // parsed in place
    else { pl.o_co = rc.take (b_co); if (pl.o_co != (size_t) -1 && b_co) memcpy (rc.p + pl.o_co, pic.coefs, b_co); }
    pl.o_cmb = buf->r[R_CMB].take (n * sizeof (lh264_cmb), n * sizeof (lh264_cmb) * buf->expect * 9 / 8);
    pl.o_xmb = buf->r[R_XMB].take (n * 80 + 16);                                  // worst case: five units per macroblock; trimmed below
    if (pl.o_sl == (size_t) -1 || pl.o_co == (size_t) -1 || pl.o_cmb == (size_t) -1 || pl.o_xmb == (size_t) -1) { buf->failed = true; return LH264_E_NOMEM; }
    memcpy (buf->r[R_SLICES].p + pl.o_sl, pic.slices, b_sl);
    lh264_cmb* out = reinterpret_cast<lh264_cmb*> (buf->r[R_CMB].p + pl.o_cmb);
    uint8_t* extra = buf->r[R_XMB].p + pl.o_xmb;
    uint32_t nx = 0;
    lh264_mb_rec none;
    memset (&none, 0, sizeof (none));                                            // (LH264_MB_NONE: whatever the scratch record held is dropped)
    for (size_t i = 0; i < n; i++) {
      const lh264_mb_rec& m = pic.mbs[i].mb_type == LH264_MB_NONE ? none : pic.mbs[i];
      lh264_cmb c;
      lh264_compact_mb (m, c, extra, &nx);
      stream16 (out + i, &c, sizeof (c));
      if (verify) { lh264_mb_rec back; lh264_expand_mb (c, extra, back); if (memcmp (&back, &m, sizeof (m))) { fprintf (stderr, "compact MB record %zu does not expand to its source (type %d)\n", i, m.mb_type); abort(); } }
    }
    pl.n_xmb = (int)nx;
    buf->r[R_XMB].used = pl.o_xmb + (size_t)nx * 16;
    if (arena) arena->picture_done();
    buf->placed.push_back (pl);
    placed_now++;
    cur_ = (int)buf->placed.size() - 1;
    return 0;
  }
  int submit_model (const lh264_model_picture& pic, const DecodedPictureInfo&) override {
    if (cur_ < 0 || cur_ >= (int)buf->placed.size()) return 0;
    Placed& pl = buf->placed[cur_];
    const size_t b_sl = (size_t)pic.n_slices * sizeof (lh264_model_slice), n = (size_t)pic.n_recs;
    pl.o_msl = buf->r[R_MSLICES].take (b_sl);
    pl.o_cmm = buf->r[R_CMM].take (n * sizeof (lh264_cmm), n * sizeof (lh264_cmm) * buf->expect * 9 / 8);
    pl.o_xmm = buf->r[R_XMM].take (n * 64 + 16);                                  // worst case: four units per record; trimmed below
    if (pl.o_msl == (size_t) -1 || pl.o_cmm == (size_t) -1 || pl.o_xmm == (size_t) -1) { buf->failed = true; return LH264_E_NOMEM; }
    memcpy (buf->r[R_MSLICES].p + pl.o_msl, pic.slices, b_sl);
    lh264_cmm* out = reinterpret_cast<lh264_cmm*> (buf->r[R_CMM].p + pl.o_cmm);
    uint8_t* extra = buf->r[R_XMM].p + pl.o_xmm;
    uint32_t nx = 0;
    for (size_t i = 0; i < n; i++) {
      lh264_cmm c;
      lh264_compact_mm (pic.recs[i], c, extra, &nx);
      stream16 (out + i, &c, sizeof (c));
      if (verify) { lh264_model_mb back; lh264_expand_mm (c, extra, back); if (memcmp (&back, &pic.recs[i], sizeof (back))) { fprintf (stderr, "compact model record %zu does not expand to its source (type code %d)\n", i, pic.recs[i].type_code); abort(); } }
    }
    pl.n_xmm = (int)nx;
    buf->r[R_XMM].used = pl.o_xmm + (size_t)nx * 16;
    pl.n_mslices = pic.n_slices; pl.n_mrecs = pic.n_recs;
    return 0;
  }
  void begin_step (StepBuf* b, StepArena* a) { buf = b; arena = a; cur_ = -1; placed_now = 0; }
  StepArena* arena = nullptr;
 private:
  static bool inside (const PinRegion& r, const void* p) { return r.p && (const uint8_t*)p >= r.p && (const uint8_t*)p < r.p + r.cap; }
  int cur_ = -1;
};

// offset and size of a placed picture's part in region k of its step buffer (0 bytes: the picture has no such part)
inline size_t part_off (const Placed& pl, int k) {
  switch (k) { case R_CMB: return pl.o_cmb; case R_XMB: return pl.o_xmb; case R_COEFS: return pl.o_co; case R_CMM: return pl.o_cmm; case R_XMM: return pl.o_xmm;
               case R_SLICES: return pl.o_sl; default: return pl.o_msl; }
}
inline size_t part_size (const Placed& pl, int k) {
  switch (k) { case R_CMB: return (size_t)pl.n_mbs * sizeof (lh264_cmb); case R_XMB: return (size_t)pl.n_xmb * 16; case R_COEFS: return (size_t)pl.n_blocks * 32;
               case R_CMM: return (size_t)pl.n_mrecs * sizeof (lh264_cmm); case R_XMM: return (size_t)pl.n_xmm * 16;
               case R_SLICES: return (size_t)pl.n_slices * sizeof (lh264_slice_rec); default: return (size_t)pl.n_mslices * sizeof (lh264_model_slice); }
}
inline void set_part_off (Placed& pl, int k, size_t o) {
  switch (k) { case R_CMB: pl.o_cmb = o; break; case R_XMB: pl.o_xmb = o; break; case R_COEFS: pl.o_co = o; break; case R_CMM: pl.o_cmm = o; break;
               case R_XMM: pl.o_xmm = o; break; case R_SLICES: pl.o_sl = o; break; default: pl.o_msl = o; break; }
}

struct Stream {
  const uint8_t* data = nullptr;        // Annex-B bytes, followed by the 00 00 00 01 sentinel
  size_t size = 0, pos = 0;
  std::vector<uint8_t> own;             // when the caller's bytes had no sentinel
  StageSink sink;
  std::unique_ptr<Decoder> dec;
  PipBoilerplate pip;
  bool flushed = false;
  int loops_left = 0;
  long parsed = 0;                      // pictures placed so far
  int last_error = 0;
  std::atomic<long> steps_parsed {0};   // steps this stream has been taken through (its tickets are served in order)
  bool live = true;                     // more pictures may come (touched by the holder of the stream's current ticket only)
  // a damaged stream can complete more than one picture in one unit; a step takes one picture per stream, the others wait in
  // a private step buffer (heap, not pinned: the rare path) and are moved into the staging of the following steps
  struct Spill { Placed pl; std::vector<uint8_t> part[R_COUNT]; };
  std::vector<Spill> spill;
  // submit-thread side
  int gpu_stream = -1, model_stream = -1, model_mbs = 0, geo_seen = 0;
  long pictures = 0, pictures_output = 0;
  Sha1 yuv;
  // results
  std::vector<uint8_t> pip_default;
  std::vector<std::vector<uint8_t> > pip_tags;
  std::string yuv_hex = "-", pip_hex = "-";
  size_t pip_bytes = 0;
};

} // namespace

struct lh264_rt {
  lh264_rt_config cfg;
  std::vector<std::unique_ptr<Stream> > streams;
  std::string error;
  lh264_rt_stats stats;
  bool ran = false;
  std::mutex err_mu;
  void fail (const std::string& e) { std::lock_guard<std::mutex> g (err_mu); if (error.empty()) error = e; }
};

namespace {

struct Pipeline;

struct Out {                             // pictures on their way to the host
  bool valid = false;
  int ticket = 0;                       // lh264_read_ticket of the packed read
  std::vector<Placed> pics;
  size_t per = 0;
  int idx = 0;
};


struct Worker {
  Pipeline* pl = nullptr;
  std::vector<StepBuf> bufs;
  long steps_done = 0;                  // guarded by Pipeline::mu
  bool finished = false;                // no picture after step steps_done - 1
  double busy = 0, wait = 0, alloc = 0;
  std::thread th;
};

struct Pipeline {
  lh264_rt* rt = nullptr;
  int index = 0;
  lh264_ctx* ctx = nullptr;
  std::vector<std::unique_ptr<Worker> > workers;
  std::vector<int> streams;             // indices into rt->streams
  std::atomic<long> ticket {0};         // next (step, stream) pair to parse: step = ticket / n, stream = streams[ticket % n]
  std::atomic<int> live {0};            // streams that may still bring pictures
  std::mutex mu;
  std::condition_variable cv_parsed, cv_released;
  long released = 0;                    // steps whose staging buffers may be overwritten again
  std::atomic<bool> abort {false};
  int rc = 0;
  long pictures = 0;
  double t_wait_parse = 0, t_wait_gpu = 0, t_end = 0, t_first = 0, t_open = 0;
  double t_submit = 0, t_read = 0, t_desc = 0;     // LH264_STATS: where the submit thread's own time goes
  long first_pictures = 0;
  size_t h2d = 0, d2h = 0;
  unsigned long long launches = 0;
  lh264_rt_picture_fn fn = nullptr; void* user = nullptr;
  std::thread th;
  // decoded pictures on their way to the host: the submit thread issues the packed reads, an output thread waits for each copy
  // and delivers it (SHA-1 / callback), so that the next step's submit never waits behind a device-to-host copy
  std::mutex out_mu;
  std::condition_variable out_cv;
  std::deque<Out> out_q;
  int out_in_flight = 0;                // reads issued and not yet delivered
  bool out_quit = false;
  uint8_t* out_pin[2] = {nullptr, nullptr}; size_t out_cap[2] = {0, 0};   // (pinned: 3.2 GB each for 1024 1080p streams)
  std::future<std::pair<uint8_t*, size_t> > out_prefetch;                  // out_pin[1] on its way (helper thread)
  long out_issued = 0;
};

// Parses until the stream has placed one picture into `buf` (or ended).  h264dec's unit loop, h264dec.cpp:246-364.
void advance (Stream& s, StepBuf* buf, StepArena* arena, bool want_pip, long max_pictures) {
  s.sink.begin_step (buf, arena);
  if (max_pictures > 0 && s.parsed >= max_pictures) return;
  if (!s.spill.empty()) {                // a picture that was completed together with an earlier one
    Stream::Spill& sp = s.spill.front();
    Placed pl = sp.pl;
    for (int k = 0; k < R_COUNT; k++) {
      const size_t o = buf->r[k].take (sp.part[k].size());
      if (o == (size_t) -1) { buf->failed = true; return; }
      if (!sp.part[k].empty()) memcpy (buf->r[k].p + o, sp.part[k].data(), sp.part[k].size());
      set_part_off (pl, k, o);
    }
    buf->placed.push_back (pl);
    s.spill.erase (s.spill.begin());
    s.parsed++;
    return;
  }
  arena->bind (buf);
  if (!s.dec->set_arena (arena)) { buf->failed = true; return; }   // (a picture left open by the previous step moves into this buffer)
  const size_t first = buf->placed.size();
  size_t m0[R_COUNT];
  for (int k = 0; k < R_COUNT; k++) m0[k] = buf->r[k].used;
  while (s.sink.placed_now == 0 && !s.flushed) {
    if (s.pos >= s.size) {
      if (s.loops_left > 0) { s.loops_left--; s.pos = 0; continue; }
      s.dec->flush(); s.flushed = true; break;
    }
    const size_t len = PipBoilerplate::next_unit (s.data, s.pos, s.size);
    if (len < 4) { if (want_pip) s.pip.raw (s.data + s.pos, len); s.pos += len; continue; }
    if (want_pip) s.pip.unit_begin (s.data + s.pos, len, s.dec->sps_seen());
    const int rc = s.dec->decode (s.data + s.pos, len);
    if (rc < 0) s.last_error = rc;
    if (want_pip) {
      UnitSliceInfo u;
      u.slice_decoded = s.dec->last_slice().decoded; u.cabac = s.dec->last_slice().cabac; u.rbsp = s.dec->last_slice().rbsp; u.header_bits = s.dec->last_slice().header_bits;
      s.pip.unit_end (u);
    }
    s.pos += len;
  }
  // more than one picture completed by the last unit: the first stays in this step, the others move to the spill list
  if (buf->placed.size() > first + 1) {
    for (size_t i = first + 1; i < buf->placed.size(); i++) {
      const Placed& pl = buf->placed[i];
      Stream::Spill sp;
      sp.pl = pl;
      for (int k = 0; k < R_COUNT; k++) { const uint8_t* q = buf->r[k].p + part_off (pl, k); sp.part[k].assign (q, q + part_size (pl, k)); }
      s.spill.push_back (std::move (sp));
    }
    // give the bytes back: everything behind the first picture of this call
    const Placed& f = buf->placed[first];
    for (int k = 0; k < R_COUNT; k++) buf->r[k].used = part_size (f, k) ? part_off (f, k) + part_size (f, k) : m0[k];
    buf->placed.resize (first + 1);
  }
  // a picture still in progress (damaged streams: completed only by the arrival of the next one) must not stay in this buffer
  if (s.dec->picture_open() && !s.dec->set_arena (nullptr)) buf->failed = true;
  (void)m0;
  if (buf->placed.size() > first) s.parsed++;
}

void worker_main (Worker* w) {
  Pipeline* P = w->pl;
  lh264_rt* rt = P->rt;
  lh264_thread_bind (rt->cfg.device);    // this thread page-locks staging memory: on the pipeline's GPU, not on GPU 0
  const int D = (int)w->bufs.size();
  const long n = (long)P->streams.size();
  const bool want_pip = rt->cfg.want_pip != 0;
  StepArena arena;
  long cur = -1;                         // the step whose staging buffer this thread is filling
  auto leave = [&] {                     // nothing behind steps 0 .. cur from this thread
    { std::lock_guard<std::mutex> g (P->mu); w->steps_done = cur + 1; w->finished = true; }
    P->cv_parsed.notify_all();
  };
  for (;;) {
    const long t = P->ticket.fetch_add (1, std::memory_order_relaxed);
    const long k = t / n;
    if (P->live.load (std::memory_order_acquire) == 0 || P->abort.load()) { leave(); return; }
    // close the steps in front of the ticket's (a step this thread drew nothing of leaves an empty buffer) and take buffer k mod D
    while (cur < k) {
      if (cur >= 0) {
        { std::lock_guard<std::mutex> g (P->mu); w->steps_done = cur + 1; }
        P->cv_parsed.notify_all();
      }
      const double t0 = now();
      {
        std::unique_lock<std::mutex> g (P->mu);
        P->cv_released.wait (g, [&] { return P->abort.load() || cur + 1 < P->released + D; });
      }
      w->wait += now() - t0;
      if (P->abort.load()) { leave(); return; }
      cur++;
      w->bufs[cur % D].reset();
    }
    Stream& s = *rt->streams[P->streams[t % n]];
    if (s.steps_parsed.load (std::memory_order_acquire) < k) {          // the stream's previous picture is still being parsed
      const double t0 = now();
      while (s.steps_parsed.load (std::memory_order_acquire) < k && !P->abort.load()) std::this_thread::yield();
      w->wait += now() - t0;
      if (P->abort.load()) { leave(); return; }
    }
    if (s.live) {
      const double t0 = now();
      advance (s, &w->bufs[k % D], &arena, want_pip, rt->cfg.max_pictures);
      w->busy += now() - t0;
      if (!(!s.spill.empty() || (!s.flushed && !(rt->cfg.max_pictures > 0 && s.parsed >= rt->cfg.max_pictures)))) {
        s.live = false;
        P->live.fetch_sub (1, std::memory_order_acq_rel);
      }
    }
    s.steps_parsed.store (k + 1, std::memory_order_release);
  }
}

void pipeline_main (Pipeline* P) {
  lh264_rt* rt = P->rt;
  const lh264_rt_config& cfg = rt->cfg;
  lh264_thread_bind (cfg.device);
  const bool want_yuv = cfg.want_yuv != 0, want_hash = want_yuv && cfg.want_hash != 0, want_pip = cfg.want_pip != 0;
  lh264_ctx* ctx = P->ctx;
  const int D = (int)P->workers[0]->bufs.size();
  std::vector<uint8_t> out_host;
  std::vector<lh264_compact_picture> cdescs;
  std::vector<int> rd_streams, rd_slots;
  std::vector<Placed> step;
  std::vector<const StepBuf*> step_buf;
  std::vector<long> done_snapshot;
#define FAIL(msg) do { rt->fail (std::string (msg) + ": " + lh264_last_error()); P->rc = -1; goto done; } while (0)

  auto deliver = [&] (const Out& o) {    // pictures of a finished step: SHA-1 + callback
    const int n = (int)o.pics.size();
    const uint8_t* base = P->out_pin[o.idx];
    if (want_hash && n > 0) {
      const int T = std::min (n, std::max (1, std::min (8, (int)std::thread::hardware_concurrency())));
      std::atomic<int> next (0);
      auto job = [&] { for (int i; (i = next.fetch_add (1)) < n;) if (o.pics[i].info.output) rt->streams[o.pics[i].stream]->yuv.update (base + o.per * i, o.per); };
      if (T <= 1) job();
      else { std::vector<std::thread> th; for (int t = 0; t < T; t++) th.emplace_back (job); for (auto& t : th) t.join(); }
    }
    for (int i = 0; i < n; i++) {
      Stream& s = *rt->streams[o.pics[i].stream];
      if (!o.pics[i].info.output) continue;
      if (P->fn) P->fn (P->user, o.pics[i].stream, s.pictures_output, o.pics[i].info.width, o.pics[i].info.height, base + o.per * i);
      s.pictures_output++;
    }
  };
  // the output thread: one packed read at a time, oldest first
  std::thread out_thread ([&] {
    for (;;) {
      Out o;
      {
        std::unique_lock<std::mutex> g (P->out_mu);
        P->out_cv.wait (g, [&] { return P->out_quit || !P->out_q.empty(); });
        if (P->out_q.empty()) return;
        o = std::move (P->out_q.front());
        P->out_q.pop_front();
      }
      if (lh264_output_wait_ticket (ctx, o.ticket) != LH264_OK) { rt->fail (std::string ("lh264_output_wait_ticket: ") + lh264_last_error()); P->rc = -1; }   // this read's copy only: the next one may still be running
      else deliver (o);
      { std::lock_guard<std::mutex> g (P->out_mu); P->out_in_flight--; }
      P->out_cv.notify_all();
    }
  });
  auto drain_output = [&] { std::unique_lock<std::mutex> g (P->out_mu); P->out_cv.wait (g, [&] { return P->out_in_flight == 0; }); };

  for (long k = 0;; k++) {
    // ---- wait for step k of every parse thread ----
    bool all_finished = true;
    {
      const double t0 = now();
      std::unique_lock<std::mutex> g (P->mu);
      P->cv_parsed.wait (g, [&] { for (auto& w : P->workers) if (w->steps_done <= k && !w->finished) return false; return true; });
      P->t_wait_parse += now() - t0;
      done_snapshot.clear();
      for (auto& w : P->workers) { if (!(w->finished && w->steps_done <= k + 1)) all_finished = false; done_snapshot.push_back (w->steps_done); }
    }
    step.clear(); step_buf.clear();
    for (size_t wi = 0; wi < P->workers.size(); wi++) {
      if (done_snapshot[wi] <= k) continue;
      StepBuf& b = P->workers[wi]->bufs[k % D];
      if (b.failed) { rt->fail ("cannot allocate pinned staging memory"); P->rc = LH264_E_NOMEM; goto done; }
      for (const Placed& pl : b.placed) { step.push_back (pl); step_buf.push_back (&b); }
    }
    if (step.empty()) {
      if (all_finished) break;
      std::lock_guard<std::mutex> g (P->mu);
      P->released = k + 1;
      P->cv_released.notify_all();
      continue;
    }
    // ---- device streams follow the geometry of their stream (SPS change = new picture storage) ----
    {
      bool reopen = false;
      for (const Placed& pl : step) if (rt->streams[pl.stream]->geo_seen != pl.geo_serial) reopen = true;
      if (reopen) {
        drain_output();
        if (lh264_sync (ctx) != LH264_OK) FAIL ("lh264_sync");
      }
      const double to0 = now();
      for (const Placed& pl : step) {
        Stream& s = *rt->streams[pl.stream];
        if (s.geo_seen != pl.geo_serial) {
          if (s.gpu_stream >= 0) lh264_stream_close (ctx, s.gpu_stream);
          s.gpu_stream = -1;
          if (want_yuv) { s.gpu_stream = lh264_stream_open (ctx, pl.mb_w, pl.mb_h, pl.n_slots); if (s.gpu_stream < 0) FAIL ("lh264_stream_open"); }
          s.geo_seen = pl.geo_serial;
        }
        if (want_pip && pl.n_mrecs > 0 && s.model_stream < 0) {
          // the model's priors live across a change of resolution: opened once, for the largest picture a level allows
          s.model_mbs = std::max (pl.mb_w * pl.mb_h, 36864);
          s.model_stream = lh264_model_open (ctx, s.model_mbs, 1);
          if (s.model_stream < 0) FAIL ("lh264_model_open");
        }
      }
      P->t_open += now() - to0;
    }
    // ---- descriptors ----
    cdescs.clear(); rd_streams.clear(); rd_slots.clear();
    {
      int out_w = 0, out_h = 0, crop_x = 0, crop_y = 0;
      bool same_geometry = true, any_model = false;
      for (size_t pi = 0; pi < step.size(); pi++) {
        const Placed& pl = step[pi];
        const StepBuf* b = step_buf[pi];
        Stream& s = *rt->streams[pl.stream];
        lh264_compact_picture d;
        memset (&d, 0, sizeof (d));
        d.stream = s.gpu_stream; d.model_stream = s.model_stream; d.cur_slot = pl.info.slot; d.mb_w = pl.mb_w; d.mb_h = pl.mb_h;
        d.n_slices = pl.n_slices; d.n_model_slices = pl.n_mslices; d.n_recs = want_pip ? pl.n_mrecs : 0; d.n_coef_blocks = pl.n_blocks;
        d.n_mb_extra = pl.n_xmb; d.n_mm_extra = pl.n_xmm;
        d.slices = reinterpret_cast<const lh264_slice_rec*> (b->r[R_SLICES].p + pl.o_sl);
        d.model_slices = reinterpret_cast<const lh264_model_slice*> (b->r[R_MSLICES].p + pl.o_msl);
        d.cmbs = reinterpret_cast<const lh264_cmb*> (b->r[R_CMB].p + pl.o_cmb);
        d.cmms = reinterpret_cast<const lh264_cmm*> (b->r[R_CMM].p + pl.o_cmm);
        d.mb_extra = b->r[R_XMB].p + pl.o_xmb; d.mm_extra = b->r[R_XMM].p + pl.o_xmm;
        d.coefs = reinterpret_cast<const int16_t*> (b->r[R_COEFS].p + pl.o_co);
        cdescs.push_back (d);
        if (want_yuv) P->h2d += part_size (pl, R_SLICES) + part_size (pl, R_CMB) + part_size (pl, R_XMB);
        P->h2d += part_size (pl, R_COEFS);
        if (want_pip && pl.n_mrecs > 0) {
          if (pl.mb_w * pl.mb_h > s.model_mbs) { rt->fail ("a stream's picture grew beyond level 5.2"); P->rc = LH264_E_UNSUPPORTED; goto done; }
          any_model = true;
          P->h2d += part_size (pl, R_MSLICES) + part_size (pl, R_CMM) + part_size (pl, R_XMM);
        }
        rd_streams.push_back (s.gpu_stream); rd_slots.push_back (pl.info.slot);
        if (out_w == 0) { out_w = pl.info.width; out_h = pl.info.height; crop_x = pl.info.crop_x; crop_y = pl.info.crop_y; }
        else if (out_w != pl.info.width || out_h != pl.info.height || crop_x != pl.info.crop_x || crop_y != pl.info.crop_y) same_geometry = false;
      }
      // ---- the device ----
      static const bool parse_only = getenv ("LH264_RT_PARSE_ONLY") != nullptr;   // diagnostic: how fast the parse threads alone are
      const int sflags = (want_yuv ? LH264_COMPACT_RECON : 0) | (any_model ? LH264_COMPACT_MODEL : 0);
      const double ts0 = now();
      if (!parse_only && sflags && lh264_submit_compact (ctx, cdescs.data(), (int)cdescs.size(), sflags) != LH264_OK) FAIL ("lh264_submit_compact");
      P->t_submit += now() - ts0;
      if (want_yuv && !parse_only) {
        if (same_geometry) {
          const size_t per = (size_t)out_w * out_h + 2 * (size_t) ((out_w + 1) / 2) * ((out_h + 1) / 2);
          const int oi = (int) (P->out_issued % 2);
          {                                                // output buffer oi is free once the read of two steps ago is delivered
            const double tg0 = now();
            std::unique_lock<std::mutex> g (P->out_mu);
            P->out_cv.wait (g, [&] { return P->out_in_flight < 2; });
            P->t_wait_gpu += now() - tg0;
          }
          if (P->out_prefetch.valid() && oi == 1) {          // the second output buffer, page-locked by the helper thread since step 0
            const std::pair<uint8_t*, size_t> got = P->out_prefetch.get();
            if (got.first) { pinned_cache().put (P->out_pin[1], P->out_cap[1]); P->out_pin[1] = got.first; P->out_cap[1] = got.second; }
          }
          if (P->out_issued == 0 && per * step.size() > P->out_cap[1] && !P->out_prefetch.valid()) {
            // page-locking 3 GB takes about a second: the buffer of step 1 is made beside step 0 instead of inside step 1
            const size_t want = per * step.size();
            const int dev = cfg.device;
            P->out_prefetch = std::async (std::launch::async, [want, dev] { lh264_thread_bind (dev); size_t cap = 0; uint8_t* q = pinned_cache().get (want, &cap); return std::make_pair (q, q ? cap : (size_t)0); });
          }
          if (per * step.size() > P->out_cap[oi]) {
            pinned_cache().put (P->out_pin[oi], P->out_cap[oi]);
            P->out_pin[oi] = pinned_cache().get (per * step.size(), &P->out_cap[oi]);
            if (!P->out_pin[oi]) { P->out_cap[oi] = 0; rt->fail ("cannot allocate pinned output memory"); P->rc = LH264_E_NOMEM; goto done; }
          }
          const double tr0 = now();
          if (lh264_pictures_read_packed (ctx, rd_streams.data(), rd_slots.data(), (int)step.size(), crop_x, crop_y, out_w, out_h, P->out_pin[oi], P->out_cap[oi], 0) != LH264_OK)
            FAIL ("lh264_pictures_read_packed");
          P->t_read += now() - tr0;
          P->d2h += per * step.size();
          P->out_issued++;
          {
            std::lock_guard<std::mutex> g (P->out_mu);
            Out o; o.valid = true; o.pics = step; o.per = per; o.idx = oi; o.ticket = lh264_read_ticket (ctx);
            P->out_q.push_back (std::move (o));
            P->out_in_flight++;
          }
          P->out_cv.notify_all();
        } else {                                           // mixed geometry in one step: picture by picture, after what is on its way
          drain_output();
          for (const Placed& pl : step) {
            Stream& s = *rt->streams[pl.stream];
            const int w = pl.info.width, h = pl.info.height, cw = (w + 1) / 2, ch = (h + 1) / 2;
            out_host.resize ((size_t)w * h + 2 * (size_t)cw * ch);
            uint8_t* y = out_host.data();
            if (lh264_picture_read (ctx, s.gpu_stream, pl.info.slot, pl.info.crop_x, pl.info.crop_y, w, h, y, w, y + (size_t)w * h, y + (size_t)w * h + (size_t)cw * ch, cw, 1) != LH264_OK)
              FAIL ("lh264_picture_read");
            P->d2h += out_host.size();
            if (!pl.info.output) continue;
            if (want_hash) s.yuv.update (y, out_host.size());
            if (P->fn) P->fn (P->user, pl.stream, s.pictures_output, w, h, y);
            s.pictures_output++;
          }
        }
      }
      for (const Placed& pl : step) { rt->streams[pl.stream]->pictures++; P->pictures++; }
      // ---- staging buffer k mod D is free again once its H2D copies are done ----
      const double tc0 = now();
      if (lh264_inputs_consumed (ctx) != LH264_OK) FAIL ("lh264_inputs_consumed");
      P->t_wait_gpu += now() - tc0;
    }
    if (P->t_first == 0) { P->t_first = now(); P->first_pictures = (long)step.size(); }
    {
      std::lock_guard<std::mutex> g (P->mu);
      P->released = k + 1;
    }
    P->cv_released.notify_all();
  }
  if (lh264_sync (ctx) != LH264_OK) FAIL ("lh264_sync");
done:
#undef FAIL
  if (getenv ("LH264_STATS")) fprintf (stderr, "lh264 stats: submit thread: lh264_submit_compact %.3f s, lh264_pictures_read_packed %.3f s, waiting for parse %.3f s, for the device %.3f s, opening streams %.3f s\n", P->t_submit, P->t_read, P->t_wait_parse, P->t_wait_gpu, P->t_open);
  drain_output();
  { std::lock_guard<std::mutex> g (P->out_mu); P->out_quit = true; }
  P->out_cv.notify_all();
  out_thread.join();
  P->t_end = now();
  {
    std::lock_guard<std::mutex> g (P->mu);
    P->abort = true;                     // (also the normal end: no parse thread is waiting any more)
  }
  P->cv_released.notify_all();
  if (P->out_prefetch.valid()) { const std::pair<uint8_t*, size_t> got = P->out_prefetch.get(); pinned_cache().put (got.first, got.second); }   // (a run that ended after one step)
  for (int k = 0; k < 2; k++) { pinned_cache().put (P->out_pin[k], P->out_cap[k]); P->out_pin[k] = nullptr; P->out_cap[k] = 0; }
}

// the `.pip` file set of every stream of a pipeline: default stream (host) + tag side streams (GPU)
void collect_pip (Pipeline* P) {
  lh264_rt* rt = P->rt;
  for (int si : P->streams) {
    Stream& s = *rt->streams[si];
    Sha1 h;
    s.pip_default = s.pip.finish();
    h.update (s.pip_default.data(), s.pip_default.size());
    s.pip_bytes = s.pip_default.size();
    s.pip_tags.assign (LH264_NUM_TAGS, std::vector<uint8_t>());
    if (s.model_stream >= 0) {
      if (lh264_model_finish (P->ctx, s.model_stream) != LH264_OK) { rt->fail (std::string ("lh264_model_finish: ") + lh264_last_error()); P->rc = -1; return; }
      for (int tag = 1; tag < LH264_NUM_TAGS; tag++) {
        const long long n = lh264_model_read (P->ctx, s.model_stream, tag, nullptr, 0, 0);
        if (n <= 0) continue;
        std::vector<uint8_t>& b = s.pip_tags[tag];
        b.resize ((size_t)n);
        lh264_model_read (P->ctx, s.model_stream, tag, b.data(), 0, b.size());
        const uint8_t t8 = (uint8_t)tag;
        h.update (&t8, 1); h.update (b.data(), b.size());
        s.pip_bytes += b.size();
      }
    }
    s.pip_hex = h.hex();
  }
}

} // namespace

extern "C" {

int lh264_rt_create (lh264_rt** out, const lh264_rt_config* cfg) {
  if (!out) return LH264_E_INVALID;
  lh264_rt* rt = new (std::nothrow) lh264_rt();
  if (!rt) return LH264_E_NOMEM;
  memset (&rt->cfg, 0, sizeof (rt->cfg));
  memset (&rt->stats, 0, sizeof (rt->stats));
  if (cfg) rt->cfg = *cfg;
  else { rt->cfg.want_yuv = rt->cfg.want_hash = rt->cfg.want_pip = 1; }
  if (rt->cfg.threads <= 0) rt->cfg.threads = std::max (1, (int)std::thread::hardware_concurrency());
  if (rt->cfg.pipelines <= 0) rt->cfg.pipelines = 1;
  if (rt->cfg.loops <= 0) rt->cfg.loops = 1;
  if (rt->cfg.depth <= 0) rt->cfg.depth = 3;
  rt->cfg.depth = std::max (2, std::min (rt->cfg.depth, 8));
  if (!rt->cfg.want_yuv) rt->cfg.want_hash = 0;
  *out = rt;
  return LH264_OK;
}

void lh264_rt_destroy (lh264_rt* rt) { delete rt; }

int lh264_rt_add_stream (lh264_rt* rt, const uint8_t* annexb, size_t len, int trailing_start_code) {
  if (!rt || (!annexb && len) || rt->ran) return LH264_E_INVALID;
  std::unique_ptr<Stream> s (new (std::nothrow) Stream());
  if (!s) return LH264_E_NOMEM;
  if (trailing_start_code) s->data = annexb;
  else {
    static const uint8_t kStart[4] = {0, 0, 0, 1};
    s->own.assign (annexb, annexb + len);
    s->own.insert (s->own.end(), kStart, kStart + 4);
    s->data = s->own.data();
  }
  s->size = len;
  s->loops_left = rt->cfg.loops - 1;
  s->sink.stream = (int)rt->streams.size();
  s->dec.reset (new Decoder (&s->sink));
  s->dec->enable_model (rt->cfg.want_pip != 0);
  s->dec->set_streaming_stores (true, false);         // coefficient pools go to pinned staging, model records to cache-resident scratch
  s->sink.verify = getenv ("LH264_RT_VERIFY_COMPACT") != nullptr;
  rt->streams.push_back (std::move (s));
  return (int)rt->streams.size() - 1;
}

int lh264_rt_run (lh264_rt* rt, lh264_rt_picture_fn fn, void* user) {
  if (!rt || rt->ran) return LH264_E_INVALID;
  rt->ran = true;
  const int S = (int)rt->streams.size();
  if (S == 0) return LH264_OK;
  const int G = std::max (1, std::min (rt->cfg.pipelines, S));
  const int T = std::max (G, std::min (rt->cfg.threads, S));
  std::vector<std::unique_ptr<Pipeline> > pipes;
  const double t0 = now();
  for (int g = 0; g < G; g++) {
    std::unique_ptr<Pipeline> P (new Pipeline());
    P->rt = rt; P->index = g; P->fn = fn; P->user = user;
    if (lh264_ctx_create (&P->ctx, rt->cfg.device, nullptr) != LH264_OK) {
      rt->fail (std::string ("no usable CUDA device (") + lh264_last_error() + "): this runtime has no CPU fallback");
      for (auto& q : pipes) lh264_ctx_destroy (q->ctx);
      return LH264_E_NO_DEVICE;
    }
    const int first = (int) ((long long)S * g / G), last = (int) ((long long)S * (g + 1) / G);
    const int tw = std::max (1, std::min (last - first, T * (g + 1) / G - T * g / G));
    for (int i = first; i < last; i++) P->streams.push_back (i);
    P->live.store (last - first);
    for (int w = 0; w < tw; w++) {
      std::unique_ptr<Worker> W (new Worker());
      W->pl = P.get();
      W->bufs.resize (rt->cfg.depth);
      for (StepBuf& b : W->bufs) {
        b.expect = (size_t) ((last - first + tw - 1) / tw);
        for (PinRegion& x : b.r) x.alloc_s = &W->alloc;
      }
      P->workers.push_back (std::move (W));
    }
    pipes.push_back (std::move (P));
  }
  for (auto& P : pipes) {
    for (auto& w : P->workers) w->th = std::thread (worker_main, w.get());
    P->th = std::thread (pipeline_main, P.get());
  }
  int rc = 0;
  double t_end = t0;
  for (auto& P : pipes) {
    P->th.join();
    for (auto& w : P->workers) w->th.join();
    t_end = std::max (t_end, P->t_end);
    if (P->rc && !rc) rc = P->rc;
  }
  lh264_rt_stats& st = rt->stats;
  st.streams = S; st.threads = 0; st.pipelines = G;
  st.seconds = t_end - t0;
  for (auto& P : pipes) {
    if (!rc && rt->cfg.want_pip) { collect_pip (P.get()); if (P->rc && !rc) rc = P->rc; }
    st.pictures += P->pictures; st.submit_wait_parse_s += P->t_wait_parse; st.submit_wait_gpu_s += P->t_wait_gpu;
    st.h2d_bytes += P->h2d; st.d2h_bytes += P->d2h;
    if (P->t_first > 0) { st.setup_s = std::max (st.setup_s, P->t_first - t0); st.setup_pictures += P->first_pictures; }
    for (auto& w : P->workers) { st.parse_busy_s += w->busy; st.parse_wait_s += w->wait; st.pinned_alloc_s += w->alloc; st.threads++; }
    st.open_streams_s += P->t_open;
    st.kernel_launches += lh264_kernel_launches (P->ctx);
  }
  for (auto& sp : rt->streams) {
    Stream& s = *sp;
    st.input_bytes += s.size * (size_t)rt->cfg.loops;
    st.pip_bytes += s.pip_bytes;
    if (rt->cfg.want_hash) s.yuv_hex = s.yuv.hex();
  }
  for (auto& P : pipes) {
    for (int si : P->streams) {
      Stream& s = *rt->streams[si];
      if (s.model_stream >= 0) lh264_model_close (P->ctx, s.model_stream);
      if (s.gpu_stream >= 0) lh264_stream_close (P->ctx, s.gpu_stream);
      s.dec.reset();
    }
    lh264_ctx_destroy (P->ctx);
  }
  st.seconds_total = now() - t0;
  return rc;
}

int lh264_rt_stream_result_get (const lh264_rt* rt, int stream, lh264_rt_stream_result* out) {
  if (!rt || !out || stream < 0 || stream >= (int)rt->streams.size()) return LH264_E_INVALID;
  const Stream& s = *rt->streams[stream];
  memset (out, 0, sizeof (*out));
  out->pictures = s.pictures; out->pictures_output = s.pictures_output;
  snprintf (out->yuv_sha1, sizeof (out->yuv_sha1), "%s", s.yuv_hex.c_str());
  snprintf (out->pip_sha1, sizeof (out->pip_sha1), "%s", s.pip_hex.c_str());
  out->pip_bytes = s.pip_bytes; out->error = s.last_error;
  return LH264_OK;
}

long long lh264_rt_pip_read (const lh264_rt* rt, int stream, int tag, uint8_t* dst, size_t cap) {
  if (!rt || stream < 0 || stream >= (int)rt->streams.size() || tag < 0 || tag >= LH264_NUM_TAGS) return LH264_E_INVALID;
  const Stream& s = *rt->streams[stream];
  const std::vector<uint8_t>* b = tag == 0 ? &s.pip_default : (tag < (int)s.pip_tags.size() ? &s.pip_tags[tag] : nullptr);
  if (!b) return 0;
  if (dst && cap) memcpy (dst, b->data(), std::min (cap, b->size()));
  return (long long)b->size();
}

int lh264_rt_stats_get (const lh264_rt* rt, lh264_rt_stats* out) {
  if (!rt || !out) return LH264_E_INVALID;
  *out = rt->stats;
  return LH264_OK;
}

const char* lh264_rt_last_error (const lh264_rt* rt) { return rt ? rt->error.c_str() : "no runtime"; }

void lh264_rt_trim (void) { pinned_cache().trim(); staging_cache().trim(); }

} // extern "C"
