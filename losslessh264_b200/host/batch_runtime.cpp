// batch_runtime.cpp — lh264_batch: many independent streams through ONE GPU context, bitstream in, pictures + .pip out.
//
// The reference is one stream per process (its model state is process-global, compression_stream.cpp:18-30) and strictly
// single-threaded.  This is the host runtime the B200 path wants instead (SURVEY.md §7 "host thread per stream", §8f N2):
//
//   parse threads ── each owns a share of the streams: Annex-B unit loop of h264dec (h264dec.cpp:246-364) → host decoder
//        │           (NAL / slice / CAVLC parse) → the picture's hand-off records copied into a pinned staging arena
//        │           (double buffered: step t+1 is parsed while step t is on the GPU)
//   main thread  ── one lh264_recon_submit + one lh264_model_submit per step for ALL streams that produced a picture
//                   (contiguous staging ⇒ a handful of H2D copies), one packed D2H of the decoded pictures, SHA-1 per stream
//   at the end   ── lh264_model_finish per stream; `.pip` default stream + tag side streams written or hashed
//
// usage: lh264_batch [--copies N] [--threads T] [--device D] [--no-yuv] [--no-hash] [--no-model] [--pip-dir DIR] [--max-pictures P] in.264...
// prints one line per distinct input (pictures, YUV SHA1, SHA1 over all .pip bytes) and a JSON summary line.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <atomic>
#include <functional>
#include <memory>
#include <chrono>
#include <condition_variable>
#include <mutex>
#include <string>
#include <thread>
#include <vector>
#include "decoder.h"
#include "pip_container.h"
#include "sha1.h"

using namespace lh264host;

namespace {

struct Staged {                         // one finished picture of one stream, waiting to be placed in the staging arena
  bool have = false;
  DecodedPictureInfo info;
  // The records live either in the decoder itself (a complete picture ends the parse of its stream for this step, so they
  // stay valid until the stream is advanced again) or in `own` (pictures finished by the arrival of the next one).
  const lh264_slice_rec* slices = nullptr; const lh264_mb_rec* mbs = nullptr; const int16_t* coefs = nullptr;
  const lh264_model_slice* mslices = nullptr; const lh264_model_mb* mrecs = nullptr;
  int n_slices = 0, n_mbs = 0, n_blocks = 0, n_mslices = 0, n_mrecs = 0;
  std::shared_ptr<std::vector<uint8_t> > own, mown;
};

struct Placed {                         // a live picture of a step: where its arrays sit in the arena, and what the submit needs
  int stream; size_t o_sl, o_mb, o_co, o_msl, o_mr;
  int n_slices, n_blocks, n_mslices, n_mrecs;
  DecodedPictureInfo info;
};

struct Arena {                          // pinned; five tightly packed regions per step (slices, mbs, coefs, model slices, model recs)
  uint8_t* p = nullptr; size_t cap = 0;
  std::vector<Placed> placed;
  bool ensure (size_t need) {
    if (need <= cap) return true;
    if (p) lh264_host_free (p);
    cap = need + need / 4;
    p = (uint8_t*)lh264_host_alloc (cap);
    if (!p) cap = 0;
    return p != nullptr;
  }
};

class CollectSink : public PictureSink {
 public:
  int mb_w = 0, mb_h = 0, n_slots = 0;
  bool reconfigured = false;
  Staged* out = nullptr;
  std::vector<Staged> spill;            // pictures completed beyond the first one of a step (damaged streams), FIFO
  int configure (int w, int h, int slots) override { mb_w = w; mb_h = h; n_slots = slots; reconfigured = true; return 0; }
  int submit (const lh264_picture_desc& pic, const DecodedPictureInfo& info) override {
    if (!out->have) cur_ = out; else { spill.emplace_back (); cur_ = &spill.back(); }
    Staged& s = *cur_;
    s.info = info; s.have = true;
    s.n_slices = pic.n_slices; s.n_mbs = mb_w * mb_h; s.n_blocks = pic.n_coef_blocks;
    if (info.complete && cur_ == out) { s.slices = pic.slices; s.mbs = pic.mbs; s.coefs = pic.coefs; return 0; }
    const size_t b_sl = (size_t)pic.n_slices * sizeof (lh264_slice_rec), b_mb = (size_t)s.n_mbs * sizeof (lh264_mb_rec), b_co = (size_t)pic.n_coef_blocks * 32;
    s.own = std::make_shared<std::vector<uint8_t> > (b_mb + b_co + b_sl + 64);
    uint8_t* q = s.own->data();
    memcpy (q, pic.mbs, b_mb); s.mbs = reinterpret_cast<const lh264_mb_rec*> (q); q += b_mb;
    if (b_co) memcpy (q, pic.coefs, b_co);
    s.coefs = reinterpret_cast<const int16_t*> (q); q += (b_co + 15) & ~(size_t)15;
    memcpy (q, pic.slices, b_sl); s.slices = reinterpret_cast<const lh264_slice_rec*> (q);
    return 0;
  }
  int submit_model (const lh264_model_picture& pic, const DecodedPictureInfo&) override {
    if (!cur_) return 0;
    Staged& s = *cur_;
    s.n_mslices = pic.n_slices; s.n_mrecs = pic.n_recs;
    if (!s.own) { s.mslices = pic.slices; s.mrecs = pic.recs; return 0; }
    const size_t b_sl = (size_t)pic.n_slices * sizeof (lh264_model_slice), b_rc = (size_t)pic.n_recs * sizeof (lh264_model_mb);
    s.mown = std::make_shared<std::vector<uint8_t> > (b_rc + b_sl + 16);
    uint8_t* q = s.mown->data();
    memcpy (q, pic.recs, b_rc); s.mrecs = reinterpret_cast<const lh264_model_mb*> (q);
    memcpy (q + b_rc, pic.slices, b_sl); s.mslices = reinterpret_cast<const lh264_model_slice*> (q + b_rc);
    return 0;
  }
 private:
  Staged* cur_ = nullptr;
};

struct Stream {
  const std::vector<uint8_t>* data = nullptr;     // file bytes + 4-byte sentinel
  size_t size = 0, pos = 0;
  int input = 0;                                  // index of the distinct input file
  CollectSink sink;
  Decoder* dec = nullptr;
  PipBoilerplate pip;
  bool eof = false, flushed = false;
  int gpu_stream = -1, model_stream = -1;
  Staged staged[2];
  Sha1 yuv;
  long pictures = 0, parsed = 0;
  int loops_left = 0;                             // replay the input this many more times (one long concatenated stream)
  // parses until one picture is complete (or the stream ends); h264dec's unit loop
  void advance (int buf, bool want_pip) {
    Staged& st = staged[buf];
    st = Staged();
    if (!sink.spill.empty()) { st = sink.spill.front(); sink.spill.erase (sink.spill.begin()); parsed++; return; }
    sink.out = &st;
    while (!st.have && !flushed) {
      if (pos >= size) {
        if (loops_left > 0) { loops_left--; pos = 0; continue; }
        dec->flush(); flushed = true; break;
      }
      const size_t len = PipBoilerplate::next_unit (data->data(), pos, size);
      if (len < 4) { if (want_pip) pip.raw (data->data() + pos, len); pos += len; continue; }
      if (want_pip) pip.unit_begin (data->data() + pos, len, dec->sps_seen());
      dec->decode (data->data() + pos, len);
      if (want_pip) {
        UnitSliceInfo u;
        u.slice_decoded = dec->last_slice().decoded; u.cabac = dec->last_slice().cabac; u.rbsp = dec->last_slice().rbsp; u.header_bits = dec->last_slice().header_bits;
        pip.unit_end (u);
      }
      pos += len;
    }
    if (st.have) parsed++;
  }
};

class Pool {                            // fixed worker threads running "for i in [0, n): fn(i)"; one Job object per start()
 public:
  explicit Pool (int n) { for (int i = 0; i < n; i++) th_.emplace_back ([this] { loop(); }); }
  ~Pool() { { std::lock_guard<std::mutex> g (m_); quit_ = true; } cv_.notify_all(); for (auto& t : th_) t.join(); }
  void start (int n, std::function<void (int)> fn) {
    auto j = std::make_shared<Job>();
    j->fn = std::move (fn); j->n = n; j->left = n;
    std::lock_guard<std::mutex> g (m_);
    job_ = j; gen_++;
    cv_.notify_all();
  }
  void wait() { std::unique_lock<std::mutex> g (m_); done_.wait (g, [this] { return !job_ || job_->left == 0; }); }
 private:
  struct Job { std::function<void (int)> fn; int n = 0; std::atomic<int> next{0}; int left = 0; };
  void loop() {
    unsigned seen = 0;
    for (;;) {
      std::shared_ptr<Job> j;
      {
        std::unique_lock<std::mutex> g (m_);
        cv_.wait (g, [&] { return quit_ || gen_ != seen; });
        if (quit_) return;
        seen = gen_; j = job_;
      }
      for (;;) {
        const int i = j->next.fetch_add (1);
        if (i >= j->n) break;
        j->fn (i);
        std::lock_guard<std::mutex> g (m_);
        if (--j->left == 0) done_.notify_all();
      }
    }
  }
  std::vector<std::thread> th_;
  std::mutex m_;
  std::condition_variable cv_, done_;
  std::shared_ptr<Job> job_;
  unsigned gen_ = 0;
  bool quit_ = false;
};

double now() { return std::chrono::duration<double> (std::chrono::steady_clock::now().time_since_epoch()).count(); }

} // namespace

struct Options {
  int copies = 1, threads = (int)std::thread::hardware_concurrency(), device = 0, max_pictures = 1 << 30, groups = 1, loops = 1;
  bool want_yuv = true, want_model = true, want_hash = true;
  std::string pip_dir;
  std::vector<std::string> files;
};
struct StreamResult { long pictures = 0; std::string yuv, pip; };
struct GroupResult {
  int rc = 0;
  long pictures = 0;
  size_t pip_bytes = 0;
  double t_steps = 0, t_all = 0, t_main = 0, t_open = 0, t_hash = 0, t_parse_wait = 0;
  unsigned long long launches = 0;
  std::vector<StreamResult> streams;
};

// One pipeline: its own GPU context (CUDA streams), parse pool and staging; streams [first, first + n) of the run.  Several
// groups run side by side so that the H2D copies, kernels and D2H copies of different groups overlap.
static void run_group (const Options& o, const std::vector<std::vector<uint8_t> >& inputs, int first, int S, int threads, GroupResult* res) {
  lh264_ctx* ctx = nullptr;
  if (lh264_ctx_create (&ctx, o.device, nullptr) != LH264_OK) { fprintf (stderr, "no CUDA device (%s): this tool has no CPU fallback\n", lh264_last_error()); res->rc = 1; return; }
  const bool want_yuv = o.want_yuv, want_model = o.want_model, want_hash = o.want_hash, want_pip = o.want_model;
  const int n_files = (int)o.files.size();
  std::vector<Stream> streams (S);
  for (int i = 0; i < S; i++) {
    Stream& s = streams[i];
    s.input = (first + i) % n_files;
    s.data = &inputs[s.input]; s.size = inputs[s.input].size() - 4;
    s.loops_left = o.loops - 1;
    s.dec = new Decoder (&s.sink);
    s.dec->enable_model (want_model);
  }
  Arena arena[2];
  Pool pool (threads);
  std::vector<uint8_t> out_host;
  uint8_t* out_pin[2] = {nullptr, nullptr}; size_t out_cap[2] = {0, 0};
  std::function<void ()> last_consume;
  std::vector<lh264_picture_desc> descs;
  std::vector<lh264_model_picture> mdescs;
  std::vector<int> live, rd_streams, rd_slots;
  long total_pictures = 0;
  double t_parse_wait = 0, t_gpu = 0, t_hash = 0, t_open = 0, t_place = 0, t_call[4] = {0, 0, 0, 0};
  const int max_pictures = o.max_pictures;
  const bool serial_model = getenv ("LH264_BATCH_FREE") == nullptr;
#define FAIL(...) do { fprintf (stderr, __VA_ARGS__); res->rc = 1; pool.wait(); return; } while (0)

  // Lays the pictures parsed into staged[k] out in arena[k]: five tightly packed regions in stream order, so that every array
  // type of the whole batch goes up with ONE copy (the library merges contiguous host ranges); the copies into the arena run
  // on the pool.
  auto place = [&] (int k) -> bool {
    Arena& a = arena[k];
    a.placed.clear();
    size_t n_sl = 0, n_mb = 0, n_co = 0, n_msl = 0, n_mr = 0;
    for (int i = 0; i < S; i++) {
      const Staged& st = streams[i].staged[k];
      if (!st.have) continue;
      a.placed.push_back (Placed{ i, n_sl, n_mb, n_co, n_msl, n_mr, st.n_slices, st.n_blocks, st.n_mslices, st.n_mrecs, st.info });
      n_sl += (size_t)st.n_slices * sizeof (lh264_slice_rec); n_mb += (size_t)st.n_mbs * sizeof (lh264_mb_rec); n_co += (size_t)st.n_blocks * 32;
      n_msl += (size_t)st.n_mslices * sizeof (lh264_model_slice); n_mr += (size_t)st.n_mrecs * sizeof (lh264_model_mb);
    }
    const size_t b_mb = 0, b_co = (n_mb + 255) & ~(size_t)255, b_mr = b_co + ((n_co + 255) & ~(size_t)255), b_sl = b_mr + ((n_mr + 255) & ~(size_t)255),
                 b_msl = b_sl + ((n_sl + 255) & ~(size_t)255), total = b_msl + n_msl + 256;
    if (!a.ensure (total)) return false;
    for (Placed& pl : a.placed) { pl.o_mb += b_mb; pl.o_co += b_co; pl.o_mr += b_mr; pl.o_sl += b_sl; pl.o_msl += b_msl; }
    pool.start ((int)a.placed.size(), [&, k] (int j) {
      const Placed& pl = arena[k].placed[j];
      const Staged& st = streams[pl.stream].staged[k];
      uint8_t* base = arena[k].p;
      memcpy (base + pl.o_mb, st.mbs, (size_t)st.n_mbs * sizeof (lh264_mb_rec));
      if (st.n_blocks) memcpy (base + pl.o_co, st.coefs, (size_t)st.n_blocks * 32);
      memcpy (base + pl.o_sl, st.slices, (size_t)st.n_slices * sizeof (lh264_slice_rec));
      if (st.n_mrecs) { memcpy (base + pl.o_mr, st.mrecs, (size_t)st.n_mrecs * sizeof (lh264_model_mb)); memcpy (base + pl.o_msl, st.mslices, (size_t)st.n_mslices * sizeof (lh264_model_slice)); }
    });
    pool.wait();
    return true;
  };
  const bool ktimes = getenv ("LH264_BATCH_KTIMES") != nullptr;   // diagnostic: per-family kernel times (serialises the launches)
  if (ktimes) lh264_kernel_ms (ctx, -1);
  const double t0 = now();
  int buf = 0;
  pool.start (S, [&] (int i) { streams[i].advance (0, want_pip); });
  pool.wait();
  if (!place (0)) FAIL ("cannot allocate pinned staging\n");
  auto start_parse = [&] (int k) {                      // parse every stream's next picture into staged[k] (pool, asynchronous)
    pool.start (S, [&, k] (int i) { if (streams[i].parsed < max_pictures) streams[i].advance (k, want_pip); else streams[i].staged[k] = Staged(); });
  };
  start_parse (1);
  for (int step = 0;; step++) {
    live.clear();
    for (const Placed& pl : arena[buf].placed) live.push_back (pl.stream);
    if (live.empty()) { pool.wait(); break; }
    const int nb = buf ^ 1;
    const double tg0 = now();
    descs.clear(); mdescs.clear(); rd_streams.clear(); rd_slots.clear();
    int out_w = 0, out_h = 0, crop_x = 0, crop_y = 0;
    bool same_geometry = true;
    for (const Placed& pl : arena[buf].placed) {
      const int i = pl.stream;
      Stream& s = streams[i];
      const Placed& st = pl;                            // (staged[] of this step may already be reused by the parse of step t+2)
      const double to = now();
      if (s.sink.reconfigured) {
        if (s.gpu_stream >= 0) lh264_stream_close (ctx, s.gpu_stream);
        s.gpu_stream = lh264_stream_open (ctx, s.sink.mb_w, s.sink.mb_h, s.sink.n_slots);
        s.sink.reconfigured = false;
        if (s.gpu_stream < 0) FAIL ("lh264_stream_open: %s\n", lh264_last_error());
      }
      if (want_model && s.model_stream < 0) {
        s.model_stream = lh264_model_open (ctx, s.sink.mb_w * s.sink.mb_h, 1);
        if (s.model_stream < 0) FAIL ("lh264_model_open: %s\n", lh264_last_error());
      }
      t_open += now() - to;
      const uint8_t* base = arena[buf].p;
      lh264_picture_desc d;
      d.stream = s.gpu_stream; d.cur_slot = st.info.slot; d.n_slices = st.n_slices; d.n_coef_blocks = st.n_blocks;
      d.slices = reinterpret_cast<const lh264_slice_rec*> (base + pl.o_sl);
      d.mbs = reinterpret_cast<const lh264_mb_rec*> (base + pl.o_mb);
      d.coefs = reinterpret_cast<const int16_t*> (base + pl.o_co);
      descs.push_back (d);
      if (want_model && st.n_mrecs > 0) {
        lh264_model_picture m;
        m.model_stream = s.model_stream; m.mb_w = s.sink.mb_w; m.mb_h = s.sink.mb_h; m.n_slices = st.n_mslices; m.n_recs = st.n_mrecs; m.n_coef_blocks = st.n_blocks;
        m.slices = reinterpret_cast<const lh264_model_slice*> (base + pl.o_msl);
        m.recs = reinterpret_cast<const lh264_model_mb*> (base + pl.o_mr);
        m.coefs = d.coefs;
        mdescs.push_back (m);
      }
      rd_streams.push_back (s.gpu_stream); rd_slots.push_back (st.info.slot);
      if (out_w == 0) { out_w = st.info.width; out_h = st.info.height; crop_x = st.info.crop_x; crop_y = st.info.crop_y; }
      else if (out_w != st.info.width || out_h != st.info.height || crop_x != st.info.crop_x || crop_y != st.info.crop_y) same_geometry = false;
    }
    double tq = now();
    if (lh264_recon_submit (ctx, descs.data(), (int)descs.size()) != LH264_OK) FAIL ("lh264_recon_submit: %s\n", lh264_last_error());
    t_call[0] += now() - tq; tq = now();
    if (!mdescs.empty() && lh264_model_submit (ctx, mdescs.data(), (int)mdescs.size()) != LH264_OK) FAIL ("lh264_model_submit: %s\n", lh264_last_error());
    if (serial_model) lh264_join (ctx);               // next reconstruction after this model (they slow each other down)
    t_call[1] += now() - tq;
    // host work of the NEXT step while this one is on the GPU: wait for its parse, copy its records into the other arena
    { const double tw = now(); pool.wait(); t_parse_wait += now() - tw; }
    if (lh264_inputs_consumed (ctx) != LH264_OK) FAIL ("lh264_inputs_consumed: %s\n", lh264_last_error());   // arena[nb] fed step t-1
    { const double tp = now(); if (!place (nb)) FAIL ("cannot allocate pinned staging\n"); t_place += now() - tp; }
    start_parse (buf);                                  // step t+2: its records are copied out of staged[] only after the next place()
    tq = now();
    // The decoded pictures of step t-1 are still on their way to the host (copy stream) while step t's records went up and its
    // kernels ran: wait for them now, start this step's copy, then hash the previous pictures while it runs.
    struct Out { bool valid = false; std::vector<int> lv; std::vector<char> shown; size_t per = 0; int idx = 0; };
    static thread_local Out prev;
    auto consume_prev = [&] () {
      if (!prev.valid) return;
      prev.valid = false;
      if (!want_hash) return;
      pool.wait();                                      // (the pool runs one job at a time: the parse of step t+2 first)
      const double th = now();
      const Out p = prev;
      pool.start ((int)p.lv.size(), [&, p] (int k) { if (p.shown[k]) streams[p.lv[k]].yuv.update (out_pin[p.idx] + p.per * k, p.per); });
      pool.wait();
      t_hash += now() - th;
    };
    if (want_yuv && lh264_output_wait (ctx) != LH264_OK) FAIL ("lh264_output_wait: %s\n", lh264_last_error());
    t_call[2] += now() - tq; tq = now();
    if (want_yuv) {
      if (same_geometry) {
        const size_t per = (size_t)out_w * out_h + 2 * (size_t) ((out_w + 1) / 2) * ((out_h + 1) / 2);
        const int oi = step & 1;
        if (per * live.size() > out_cap[oi]) { if (out_pin[oi]) lh264_host_free (out_pin[oi]); out_cap[oi] = per * live.size(); out_pin[oi] = (uint8_t*)lh264_host_alloc (out_cap[oi]); }
        if (!out_pin[oi] || lh264_pictures_read_packed (ctx, rd_streams.data(), rd_slots.data(), (int)live.size(), crop_x, crop_y, out_w, out_h, out_pin[oi], out_cap[oi], 0) != LH264_OK)
          FAIL ("lh264_pictures_read_packed: %s\n", lh264_last_error());
        t_call[3] += now() - tq;
        consume_prev();
        prev.valid = true; prev.lv = live; prev.per = per; prev.idx = oi;
        prev.shown.clear();
        for (const Placed& st : arena[buf].placed) prev.shown.push_back (st.info.output ? 1 : 0);   // pictures before a stream's first IDR are held back
      } else {                                           // mixed geometry: picture by picture
        consume_prev();
        for (const Placed& st : arena[buf].placed) {
          const int i = st.stream;
          const int w = st.info.width, h = st.info.height, cw = (w + 1) / 2, ch = (h + 1) / 2;
          out_host.resize ((size_t)w * h + 2 * (size_t)cw * ch);
          uint8_t* y = out_host.data();
          if (lh264_picture_read (ctx, streams[i].gpu_stream, st.info.slot, st.info.crop_x, st.info.crop_y, w, h, y, w, y + (size_t)w * h, y + (size_t)w * h + (size_t)cw * ch, cw, 1) != LH264_OK)
            FAIL ("lh264_picture_read: %s\n", lh264_last_error());
          if (want_hash && st.info.output) streams[i].yuv.update (y, out_host.size());
        }
      }
    } else {
      if (lh264_sync (ctx) != LH264_OK) FAIL ("lh264_sync: %s\n", lh264_last_error());
    }
    last_consume = consume_prev;
    for (int i : live) { streams[i].pictures++; total_pictures++; }
    t_gpu += now() - tg0;
    buf = nb;
  }
  lh264_sync (ctx);
  if (last_consume) last_consume();                     // the pictures of the last step
  res->t_steps = now() - t0;

  // the .pip file set of every stream: default stream (host) + tag side streams (GPU); hashed, and written when asked
  res->streams.resize (S);
  for (int i = 0; i < S; i++) {
    Stream& s = streams[i];
    res->streams[i].pictures = s.pictures;
    res->streams[i].yuv = want_yuv && want_hash ? s.yuv.hex() : "-";
    res->streams[i].pip = "-";
    if (!want_model) continue;
    Sha1 h;
    const std::vector<uint8_t>& def = s.pip.finish();
    h.update (def.data(), def.size());
    res->pip_bytes += def.size();
    const std::string pre = o.pip_dir + "/s" + std::to_string (first + i) + ".pip";
    if (!o.pip_dir.empty()) { FILE* f = fopen (pre.c_str(), "wb"); if (f) { fwrite (def.data(), 1, def.size(), f); fclose (f); } }
    if (s.model_stream >= 0) {
      if (lh264_model_finish (ctx, s.model_stream) != LH264_OK) FAIL ("lh264_model_finish: %s\n", lh264_last_error());
      std::vector<uint8_t> b;
      for (int tag = 1; tag < LH264_NUM_TAGS; tag++) {
        const long long n = lh264_model_read (ctx, s.model_stream, tag, nullptr, 0, 0);
        if (n <= 0) continue;
        b.resize ((size_t)n);
        lh264_model_read (ctx, s.model_stream, tag, b.data(), 0, b.size());
        const uint8_t t8 = (uint8_t)tag;
        h.update (&t8, 1); h.update (b.data(), b.size());
        res->pip_bytes += b.size();
        if (!o.pip_dir.empty()) { const std::string p = pre + "." + std::to_string (tag); FILE* f = fopen (p.c_str(), "wb"); if (f) { fwrite (b.data(), 1, b.size(), f); fclose (f); } }
      }
    }
    res->streams[i].pip = h.hex();
  }
  res->t_all = now() - t0;
  res->pictures = total_pictures; res->t_main = t_gpu; res->t_open = t_open; res->t_hash = t_hash; res->t_parse_wait = t_parse_wait;
  res->launches = lh264_kernel_launches (ctx);
  if (ktimes) fprintf (stderr, "group %d kernels: recon %.1f ms, deblock %.1f ms, model generation %.1f ms, model coder %.1f ms\n", first, lh264_kernel_ms (ctx, 0), lh264_kernel_ms (ctx, 2), lh264_kernel_ms (ctx, 4), lh264_kernel_ms (ctx, 5));
  if (getenv ("LH264_BATCH_TIMES")) fprintf (stderr, "group %d: recon_submit %.3f s, model_submit %.3f s, output_wait %.3f s, read_packed %.3f s, open %.3f s, parse wait %.3f s, staging copy %.3f s\n", first, t_call[0], t_call[1], t_call[2], t_call[3], t_open, t_parse_wait, t_place);
  for (auto& s : streams) delete s.dec;
  for (int k = 0; k < 2; k++) if (out_pin[k]) lh264_host_free (out_pin[k]);
  for (int k = 0; k < 2; k++) if (arena[k].p) lh264_host_free (arena[k].p);
  lh264_ctx_destroy (ctx);
#undef FAIL
}

int main (int argc, char** argv) {
  Options o;
  for (int i = 1; i < argc; i++) {
    const std::string a = argv[i];
    if (a == "--copies" && i + 1 < argc) o.copies = atoi (argv[++i]);
    else if (a == "--threads" && i + 1 < argc) o.threads = atoi (argv[++i]);
    else if (a == "--groups" && i + 1 < argc) o.groups = atoi (argv[++i]);
    else if (a == "--loops" && i + 1 < argc) o.loops = atoi (argv[++i]);
    else if (a == "--device" && i + 1 < argc) o.device = atoi (argv[++i]);
    else if (a == "--max-pictures" && i + 1 < argc) o.max_pictures = atoi (argv[++i]);
    else if (a == "--pip-dir" && i + 1 < argc) o.pip_dir = argv[++i];
    else if (a == "--no-yuv") o.want_yuv = false;
    else if (a == "--no-hash") o.want_hash = false;      // still brings every decoded picture to the host, but does not SHA-1 it
    else if (a == "--no-model") o.want_model = false;
    else o.files.push_back (a);
  }
  if (o.files.empty() || o.copies < 1) {
    fprintf (stderr, "usage: %s [--copies N] [--threads T] [--groups G] [--loops L] [--device D] [--no-yuv] [--no-hash] [--no-model] [--pip-dir DIR] [--max-pictures P] in.264...\n", argv[0]);
    return 2;
  }
  const int S = o.copies * (int)o.files.size();
  o.groups = std::max (1, std::min (o.groups, S));
  o.threads = std::max (o.threads, o.groups);
  std::vector<std::vector<uint8_t> > inputs (o.files.size());
  size_t in_bytes = 0;
  for (size_t f = 0; f < o.files.size(); f++) {
    FILE* fi = fopen (o.files[f].c_str(), "rb");
    if (!fi) { perror (o.files[f].c_str()); return 1; }
    uint8_t tmp[65536]; size_t n;
    while ((n = fread (tmp, 1, sizeof (tmp), fi)) > 0) inputs[f].insert (inputs[f].end(), tmp, tmp + n);
    fclose (fi);
    in_bytes += inputs[f].size() * (size_t)o.copies * o.loops;
    static const uint8_t kStart[4] = {0, 0, 0, 1};
    inputs[f].insert (inputs[f].end(), kStart, kStart + 4);
  }
  std::vector<GroupResult> results (o.groups);
  std::vector<std::thread> th;
  const double t0 = now();
  for (int g = 0; g < o.groups; g++) {
    const int first = (int) ((long long)S * g / o.groups), last = (int) ((long long)S * (g + 1) / o.groups);
    th.emplace_back (run_group, std::cref (o), std::cref (inputs), first, last - first, std::max (1, o.threads / o.groups), &results[g]);
  }
  for (auto& t : th) t.join();
  const double wall = now() - t0;
  std::vector<StreamResult> all;
  GroupResult sum;
  double t_steps = 0;
  for (auto& r : results) {
    if (r.rc) return r.rc;
    all.insert (all.end(), r.streams.begin(), r.streams.end());
    sum.pictures += r.pictures; sum.pip_bytes += r.pip_bytes; sum.launches += r.launches;
    sum.t_main += r.t_main; sum.t_open += r.t_open; sum.t_hash += r.t_hash; sum.t_parse_wait += r.t_parse_wait;
    t_steps = std::max (t_steps, r.t_steps);
  }
  // every copy of an input must give identical results; report one line per distinct input
  int mismatches = 0;
  const int nf = (int)o.files.size();
  for (int f = 0; f < nf; f++) {
    for (int i = f + nf; i < S; i += nf)
      if (all[i].yuv != all[f].yuv || all[i].pip != all[f].pip || all[i].pictures != all[f].pictures) mismatches++;
    printf ("%s pictures %ld yuv_sha1 %s pip_sha1 %s\n", o.files[f].c_str(), all[f].pictures, all[f].yuv.c_str(), all[f].pip.c_str());
  }
  printf ("{\"streams\": %d, \"threads\": %d, \"groups\": %d, \"loops\": %d, \"pictures\": %ld, \"seconds\": %.4f, \"frames_per_s\": %.1f, \"input_mb_per_s\": %.2f, \"pip_bytes\": %zu, "
          "\"seconds_incl_finish\": %.4f, \"main_threads_s\": %.4f, \"open_streams_s\": %.4f, \"hash_s\": %.4f, \"wait_for_parse_s\": %.4f, \"copy_mismatches\": %d, \"kernel_launches\": %llu}\n",
          S, o.threads, o.groups, o.loops, sum.pictures, t_steps, sum.pictures / t_steps, in_bytes / t_steps / 1e6, sum.pip_bytes, wall, sum.t_main, sum.t_open, sum.t_hash, sum.t_parse_wait,
          mismatches, sum.launches);
  return mismatches ? 3 : 0;
}
