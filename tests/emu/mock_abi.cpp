// mock_abi.cpp — TEST INFRASTRUCTURE ONLY.  A CPU stand-in for the subset of the C ABI (include/lh264_b200.h) that the multi-stream
// runtime (losslessh264_b200/host/runtime.cpp) calls, so that the runtime's threading, staging and bookkeeping can be run — and
// its output checked against the golden SHA1s — in the build container, which has no GPU.  Behind every entry point the
// product's kernel SOURCES run under tests/emu/cuda_emu.h (k_decode_rows; the k_model_* sequence of model_emu_core.h).
// Never linked into the product: libopenh264_lh264.so links liblh264_b200.so, which has no CPU path.
#include "cuda_emu.h"
#include "../../losslessh264_b200/csrc/decode_rows.cuh"
#include "../../losslessh264_b200/csrc/model_kernels.cuh"
using namespace lh264m;
#include "model_emu_core.h"
#include "../../losslessh264_b200/csrc/compact_records.h"
#include <memory>
#include <mutex>

using namespace lh264;

namespace {
std::mutex g_emu;                       // the fiber emulator is single-threaded: one launch at a time, process-wide
thread_local char g_err[256] = "";
struct StreamEmu { bool open = false; int mb_w = 0, mb_h = 0, n_slots = 0; size_t stride = 0; std::vector<uint8_t> pool, lines; };
}

struct lh264_ctx {
  std::vector<StreamEmu> streams;
  std::vector<std::unique_ptr<ModelEmu> > models;
  uint64_t launches = 0;
  bool invalid = false;
};

static int fail (int code, const char* msg) { snprintf (g_err, sizeof (g_err), "%s", msg); return code; }

extern "C" {

int lh264_abi_version (void) { return LH264_ABI_VERSION; }
const char* lh264_last_error (void) { return g_err; }
int lh264_ctx_create (lh264_ctx** out, int, void*) {
  if (getenv ("LH264_MOCK_NO_DEVICE")) return fail (LH264_E_NO_DEVICE, "mock: no device");
  *out = new lh264_ctx();
  return LH264_OK;
}
void lh264_ctx_destroy (lh264_ctx* c) { delete c; }
int lh264_stream_open (lh264_ctx* c, int mb_w, int mb_h, int n_slots) {
  StreamEmu s;
  s.open = true; s.mb_w = mb_w; s.mb_h = mb_h; s.n_slots = n_slots;
  s.stride = (size_t)mb_w * mb_h * 384;
  s.pool.assign (s.stride * n_slots + 64, 128);      // pic_queue.cpp:96
  s.lines.assign ((size_t)mb_h * 2 * mb_w * 16, 0);
  for (size_t i = 0; i < c->streams.size(); i++) if (!c->streams[i].open) { c->streams[i] = std::move (s); return (int)i; }
  c->streams.push_back (std::move (s));
  return (int)c->streams.size() - 1;
}
int lh264_stream_close (lh264_ctx* c, int s) {
  if (s < 0 || s >= (int)c->streams.size() || !c->streams[s].open) return fail (LH264_E_INVALID, "bad stream");
  c->streams[s] = StreamEmu();
  return LH264_OK;
}
int lh264_recon_submit (lh264_ctx* c, const lh264_picture_desc* pics, int n) {
  std::lock_guard<std::mutex> g (g_emu);
  std::vector<PicDev> d (n);
  int max_rows = 0;
  std::vector<char> seen (c->streams.size(), 0);
  for (int i = 0; i < n; i++) {
    const lh264_picture_desc& p = pics[i];
    if (p.stream < 0 || p.stream >= (int)c->streams.size() || !c->streams[p.stream].open) return fail (LH264_E_INVALID, "picture names a closed stream");
    if (seen[p.stream]) return fail (LH264_E_INVALID, "two pictures of one stream in a batch");
    seen[p.stream] = 1;
    StreamEmu& s = c->streams[p.stream];
    memset (&d[i], 0, sizeof (PicDev));
    d[i].mbs = p.mbs; d[i].coefs = p.coefs; d[i].slices = p.slices;
    d[i].base = s.pool.data(); d[i].slot_stride = s.stride; d[i].cur_slot = p.cur_slot; d[i].n_slots = s.n_slots; d[i].mb_w = s.mb_w; d[i].mb_h = s.mb_h;
    d[i].n_slices = p.n_slices; d[i].n_coef_blocks = p.n_coef_blocks; d[i].lines = s.lines.data();
    max_rows = std::max (max_rows, s.mb_h);
  }
  std::vector<int> sync (4 + (size_t)n * max_rows + 16, 0);
  DecodeSync* ds = reinterpret_cast<DecodeSync*> (sync.data());
  const PicDev* pd = d.data();
  unsigned int err = 0;
  emu::launch (1, 1, 3 * 32, [&] { k_decode_rows (pd, n, max_rows, ds, &err); });
  c->launches++;
  if (err) c->invalid = true;
  return LH264_OK;
}
int lh264_sync (lh264_ctx* c) { if (c->invalid) { c->invalid = false; return fail (LH264_E_INVALID, "malformed record"); } return LH264_OK; }
int lh264_join (lh264_ctx*) { return LH264_OK; }
int lh264_inputs_consumed (lh264_ctx*) { return LH264_OK; }
int lh264_output_wait (lh264_ctx*) { return LH264_OK; }
int lh264_read_ticket (const lh264_ctx*) { return 0; }
int lh264_output_wait_ticket (lh264_ctx*, int) { return LH264_OK; }
uint64_t lh264_kernel_launches (const lh264_ctx* c) { return c->launches; }
void* lh264_host_alloc (size_t n) { return malloc (n ? n : 1); }
void* lh264_host_alloc_wc (size_t n) { return malloc (n ? n : 1); }
int lh264_thread_bind (int) { return LH264_OK; }
void lh264_host_free (void* p) { free (p); }

int lh264_picture_read (lh264_ctx* c, int stream, int slot, int crop_x, int crop_y, int width, int height,
                        uint8_t* dy, int sy, uint8_t* du, uint8_t* dv, int sc, int) {
  if (stream < 0 || stream >= (int)c->streams.size() || !c->streams[stream].open) return fail (LH264_E_INVALID, "bad stream");
  const StreamEmu& s = c->streams[stream];
  const int W = s.mb_w * 16, H = s.mb_h * 16;
  const uint8_t* y = s.pool.data() + (size_t)slot * s.stride;
  const uint8_t* u = y + (size_t)W * H;
  const uint8_t* v = u + (size_t)W * H / 4;
  for (int r = 0; r < height; r++) memcpy (dy + (size_t)r * sy, y + (size_t) (crop_y + r) * W + crop_x, width);
  const int cw = (width + 1) >> 1, ch = (height + 1) >> 1;
  for (int r = 0; r < ch; r++) {
    memcpy (du + (size_t)r * sc, u + (size_t) ((crop_y >> 1) + r) * (W >> 1) + (crop_x >> 1), cw);
    memcpy (dv + (size_t)r * sc, v + (size_t) ((crop_y >> 1) + r) * (W >> 1) + (crop_x >> 1), cw);
  }
  return LH264_OK;
}
int lh264_pictures_read_packed (lh264_ctx* c, const int* streams, const int* slots, int n, int crop_x, int crop_y, int width, int height,
                                uint8_t* dst, size_t dst_bytes, int) {
  const int cw = (width + 1) >> 1, ch = (height + 1) >> 1;
  const size_t per = (size_t)width * height + 2 * (size_t)cw * ch;
  if (per * n > dst_bytes) return fail (LH264_E_INVALID, "dst too small");
  for (int i = 0; i < n; i++) {
    uint8_t* y = dst + per * i;
    const int rc = lh264_picture_read (c, streams[i], slots[i], crop_x, crop_y, width, height, y, width, y + (size_t)width * height, y + (size_t)width * height + (size_t)cw * ch, cw, 1);
    if (rc) return rc;
  }
  return LH264_OK;
}

int lh264_picture_write (lh264_ctx* c, int stream, int slot, const uint8_t* sy, int stride_y, const uint8_t* su, const uint8_t* sv, int stride_c) {
  if (stream < 0 || stream >= (int)c->streams.size() || !c->streams[stream].open) return fail (LH264_E_INVALID, "bad stream");
  StreamEmu& s = c->streams[stream];
  const int W = s.mb_w * 16, H = s.mb_h * 16;
  uint8_t* y = s.pool.data() + (size_t)slot * s.stride;
  uint8_t* u = y + (size_t)W * H;
  uint8_t* v = u + (size_t)W * H / 4;
  for (int r = 0; r < H; r++) memcpy (y + (size_t)r * W, sy + (size_t)r * stride_y, W);
  for (int r = 0; r < H / 2; r++) { memcpy (u + (size_t)r * (W / 2), su + (size_t)r * stride_c, W / 2); memcpy (v + (size_t)r * (W / 2), sv + (size_t)r * stride_c, W / 2); }
  return LH264_OK;
}

int lh264_model_open (lh264_ctx* c, int, int) { c->models.emplace_back (new ModelEmu()); return (int)c->models.size() - 1; }
int lh264_model_close (lh264_ctx* c, int m) { if (m < 0 || m >= (int)c->models.size()) return fail (LH264_E_INVALID, "bad model stream"); c->models[m].reset(); return LH264_OK; }
int lh264_model_submit (lh264_ctx* c, const lh264_model_picture* pics, int n) {
  std::lock_guard<std::mutex> g (g_emu);
  for (int i = 0; i < n; i++) {
    const int m = pics[i].model_stream;
    if (m < 0 || m >= (int)c->models.size() || !c->models[m]) return fail (LH264_E_INVALID, "bad model stream");
    if (c->models[m]->submit (pics[i])) return fail (LH264_E_INVALID, "model submit failed");
    c->launches += 7;
  }
  return LH264_OK;
}
int lh264_model_finish (lh264_ctx* c, int m) {
  std::lock_guard<std::mutex> g (g_emu);
  if (m < 0 || m >= (int)c->models.size() || !c->models[m]) return fail (LH264_E_INVALID, "bad model stream");
  return c->models[m]->finish() ? fail (LH264_E_INVALID, "model finish failed") : LH264_OK;
}
long long lh264_model_read (lh264_ctx* c, int m, int tag, uint8_t* dst, size_t offset, size_t cap) {
  if (m < 0 || m >= (int)c->models.size() || !c->models[m]) return fail (LH264_E_INVALID, "bad model stream");
  std::vector<uint8_t> b;
  if (!c->models[m]->tag_bytes (tag, b)) return fail (LH264_E_INVALID, "broken page list");
  if (dst && offset < b.size()) memcpy (dst, b.data() + offset, std::min (cap, b.size() - offset));
  return (long long)b.size();
}

// = lh264_submit_compact (lh264_shim.cu / model_host.inl): the records are expanded with the very functions k_expand_mb / k_expand_mm
// run (csrc/compact_records.h), then take the same path as full records
int lh264_submit_compact (lh264_ctx* c, const lh264_compact_picture* pics, int n, int flags) {
  std::vector<std::vector<lh264_mb_rec> > mbs (n);
  std::vector<std::vector<lh264_model_mb> > recs (n);
  std::vector<lh264_picture_desc> rd;
  std::vector<lh264_model_picture> md;
  for (int i = 0; i < n; i++) {
    const lh264_compact_picture& p = pics[i];
    if (flags & LH264_COMPACT_RECON) {
      mbs[i].resize ((size_t)p.mb_w * p.mb_h);
      for (size_t k = 0; k < mbs[i].size(); k++) lh264_expand_mb (p.cmbs[k], p.mb_extra, mbs[i][k]);
      lh264_picture_desc d;
      d.stream = p.stream; d.cur_slot = p.cur_slot; d.n_slices = p.n_slices; d.n_coef_blocks = p.n_coef_blocks;
      d.slices = p.slices; d.mbs = mbs[i].data(); d.coefs = p.coefs;
      rd.push_back (d);
    }
    if ((flags & LH264_COMPACT_MODEL) && p.n_recs > 0) {
      recs[i].resize ((size_t)p.n_recs);
      for (size_t k = 0; k < recs[i].size(); k++) lh264_expand_mm (p.cmms[k], p.mm_extra, recs[i][k]);
      lh264_model_picture q;
      q.model_stream = p.model_stream; q.mb_w = p.mb_w; q.mb_h = p.mb_h; q.n_slices = p.n_model_slices; q.n_recs = p.n_recs; q.n_coef_blocks = p.n_coef_blocks;
      q.slices = p.model_slices; q.recs = recs[i].data(); q.coefs = p.coefs;
      md.push_back (q);
    }
  }
  int rc = LH264_OK;
  if (!rd.empty()) rc = lh264_recon_submit (c, rd.data(), (int)rd.size());
  if (rc == LH264_OK && !md.empty()) rc = lh264_model_submit (c, md.data(), (int)md.size());
  return rc;
}

} // extern "C"
