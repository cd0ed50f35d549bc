// lh264_shim.cu — the C-ABI of include/lh264_b200.h over the sm_100a kernels.
//
// Host side of the hot-path boundary: owns the per-stream picture pools in HBM (the reference's
// PPicture pool, codec/decoder/core/src/pic_queue.cpp:60-107), stages a batch of hand-off records
// (SDqLayer, dec_frame.h:61-127) into HBM and issues the reconstruction launches that replace
// WelsTargetSliceConstruction (decode_slice.cpp:110-206) for every picture of the batch.
// There is no CPU fallback in this file: without a CUDA device every entry point fails.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <string.h>
#include <stdarg.h>
#include <map>
#include <vector>
#include <algorithm>
#include <utility>
#include <new>

#include "recon_kernels.cuh"
#include "recon_rows.cuh"
#include "decode_rows.cuh"
#include "model_kernels.cuh"
#include "compact_records.h"

// ---- expansion of compact records (include/lh264_b200.h: lh264_cmb / lh264_cmm) into the kernels' 128-byte records ------------
struct ExpandJob { const void* src; const uint8_t* extra; void* dst; int n; int pad; };    // one picture
__global__ void __launch_bounds__ (128) k_expand_mb (const ExpandJob* __restrict__ jobs) {
  const ExpandJob j = jobs[blockIdx.y];
  const lh264_cmb* src = static_cast<const lh264_cmb*> (j.src);
  lh264_mb_rec* dst = static_cast<lh264_mb_rec*> (j.dst);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < j.n; i += gridDim.x * blockDim.x) {
    lh264_mb_rec m;
    lh264_expand_mb (src[i], j.extra, m);
    const uint4* q = reinterpret_cast<const uint4*> (&m);
    uint4* o = reinterpret_cast<uint4*> (dst + i);
#pragma unroll
    for (int k = 0; k < 8; k++) o[k] = q[k];
  }
}
__global__ void __launch_bounds__ (128) k_expand_mm (const ExpandJob* __restrict__ jobs) {
  const ExpandJob j = jobs[blockIdx.y];
  const lh264_cmm* src = static_cast<const lh264_cmm*> (j.src);
  lh264_model_mb* dst = static_cast<lh264_model_mb*> (j.dst);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < j.n; i += gridDim.x * blockDim.x) {
    lh264_model_mb m;
    lh264_expand_mm (src[i], j.extra, m);
    const uint4* q = reinterpret_cast<const uint4*> (&m);
    uint4* o = reinterpret_cast<uint4*> (dst + i);
#pragma unroll
    for (int k = 0; k < 8; k++) o[k] = q[k];
  }
}

using namespace lh264;

static thread_local char g_err[512] = "";
static int fail (int code, const char* fmt, ...) {
  va_list ap; va_start (ap, fmt); vsnprintf (g_err, sizeof (g_err), fmt, ap); va_end (ap);
  return code;
}
#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return fail (LH264_E_CUDA, "%s: %s (%s:%d)", #call, cudaGetErrorString (e_), __FILE__, __LINE__); } while (0)

namespace {

struct StreamRec {
  bool open = false;
  uint64_t batch_stamp = 0;                  // serial number of the last batch that named this stream (duplicate check, O(n))
  int mb_w = 0, mb_h = 0, n_slots = 0;
  uint8_t* base = nullptr;
  size_t slot_stride = 0;
  uint8_t* lines = nullptr;                  // unfiltered bottom rows, one line set per MB row (decode_rows.cuh)
};

template <typename T> struct DevBuf {       // grow-only device buffer
  T* p = nullptr; size_t cap = 0;
  cudaError_t reserve (size_t n) {
    if (n <= cap) return cudaSuccess;
    if (p) cudaFree (p);
    p = nullptr; cap = 0;
    size_t want = std::max (n, (size_t)1024);
    want += want / 4;
    cudaError_t e = cudaMalloc (&p, want * sizeof (T));
    if (e == cudaSuccess) cap = want;
    return e;
  }
  void release() { if (p) cudaFree (p); p = nullptr; cap = 0; }
};
template <typename T> struct PinBuf {       // grow-only pinned host buffer
  T* p = nullptr; size_t cap = 0;
  cudaError_t reserve (size_t n) {
    if (n <= cap) return cudaSuccess;
    if (p) cudaFreeHost (p);
    p = nullptr; cap = 0;
    size_t want = std::max (n, (size_t)256);
    want += want / 4;
    cudaError_t e = cudaMallocHost (&p, want * sizeof (T));
    if (e == cudaSuccess) cap = want;
    return e;
  }
  void release() { if (p) cudaFreeHost (p); p = nullptr; cap = 0; }
};

struct BatchDev {                            // device-resident batch
  DevBuf<lh264_mb_rec> mbs;
  DevBuf<int16_t> coefs;
  DevBuf<lh264_slice_rec> slices;
  DevBuf<PicDev> pics;
  DevBuf<int> row_sync;                      // DecodeSync header + progress[n_pics][max_rows]
  DevBuf<lh264_cmb> cmb; DevBuf<uint8_t> xmb; DevBuf<ExpandJob> xjobs;   // compact hand-off: staged here, expanded into mbs
  int n_pics = 0, max_rows = 0, max_mbs = 0;
  bool compact = false;
  void release() { mbs.release(); coefs.release(); slices.release(); pics.release(); row_sync.release(); cmb.release(); xmb.release(); xjobs.release(); }
};

} // namespace

struct lh264_resident_batch { BatchDev dev; };

// Device memory of the per-stream buffers (picture slots, line buffers, priors, context records ...).  Opening a stream used to cost
// seven cudaMalloc calls - a millisecond or so each for the multi-megabyte ones: 4 s for the 1024 streams of a batch run before
// the first picture moved.  Slabs are taken 1 GB at a time and handed out by a bump pointer; a block given back is kept on a
// free list by size and reused by the next stream of the same geometry (streams of one run mostly have one geometry); the
// slabs go back to the driver when the context is destroyed.
struct DevSlabs {
  static constexpr size_t SLAB = (size_t)1 << 30, ALIGN = 512;
  struct Slab { char* p; size_t cap, used; };
  std::vector<Slab> slabs;
  std::multimap<size_t, void*> free_;
  std::vector<void*> big;                    // requests that do not fit a slab: their own cudaMalloc
  cudaError_t take (void** out, size_t bytes) {
    const size_t n = (bytes + ALIGN - 1) / ALIGN * ALIGN;
    auto it = free_.find (n);
    if (it != free_.end()) { *out = it->second; free_.erase (it); return cudaSuccess; }
    if (n > SLAB / 4) { cudaError_t e = cudaMalloc (out, n); if (e == cudaSuccess) big.push_back (*out); return e; }
    if (slabs.empty() || slabs.back().used + n > slabs.back().cap) {
      char* q = nullptr;
      size_t cap = SLAB;
      cudaError_t e = cudaMalloc (&q, cap);
      while (e != cudaSuccess && cap / 2 >= n && cap > ((size_t)32 << 20)) { cudaGetLastError(); cap /= 2; e = cudaMalloc (&q, cap); }   // memory is short: smaller slabs
      if (e != cudaSuccess) return e;
      slabs.push_back (Slab{q, cap, 0});
    }
    *out = slabs.back().p + slabs.back().used;
    slabs.back().used += n;
    return cudaSuccess;
  }
  void give (void* p, size_t bytes) {
    if (!p) return;
    const size_t n = (bytes + ALIGN - 1) / ALIGN * ALIGN;
    for (size_t i = 0; i < big.size(); i++) if (big[i] == p) { cudaFree (p); big.erase (big.begin() + (long)i); return; }
    free_.emplace (n, p);
  }
  void release() { for (auto& sl : slabs) cudaFree (sl.p); for (void* q : big) cudaFree (q); slabs.clear(); big.clear(); free_.clear(); }
};

struct lh264_ctx {
  int device = 0, n_sm = 148, occ_recon = 2;
  DevSlabs slabs;
  unsigned int* d_error = nullptr;           // records that failed validation on the device (sticky until lh264_sync reports it)
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  std::vector<StreamRec> streams;
  BatchDev batch[2];                         // staging for lh264_recon_submit, alternating: the model may still read the
                                             // coefficient pools of the previous batch (shared, see staged_coefs)
  int batch_cur = 0;
  cudaEvent_t batch_free[2] = {nullptr, nullptr};   // recorded on the model stream after the model run that shared batch[k]
  bool batch_shared[2] = {false, false};
  void* batch_sharer[2] = {nullptr, nullptr};   // the MBatch (model_host.inl) that shares batch[k]'s pools
  PinBuf<PicDev> h_pics;
  PinBuf<ExpandJob> h_xjobs;                 // (same life cycle as h_pics: guarded by `staged`)
  cudaEvent_t staged = nullptr;              // h_pics consumed by the H2D copy
  bool staged_pending = false;
  // packed read-back: job tables and output buffers alternate, so that a call neither waits for the kernels in front of it on the
  // context stream nor holds them up until the previous call's copy to the host is over
  DevBuf<PackDev> pack_jobs; PinBuf<PackDev> h_pack_jobs[2]; DevBuf<uint8_t> pack_out[2];
  cudaEvent_t pack_copied[2] = {nullptr, nullptr}, pack_read[2] = {nullptr, nullptr};
  bool pack_read_pending[2] = {false, false};
  int pack_cur = 0;
  // host-to-device copies of a submit run on a stream of their own: the records of step t+1 go up while step t's kernels run
  cudaStream_t h2d_stream = nullptr;
  cudaEvent_t uploaded = nullptr, decode_done[2] = {nullptr, nullptr};
  bool decode_done_pending[2] = {false, false};
  DevBuf<uint8_t> scratch_a, scratch_b, scratch_c;
  uint64_t launches = 0;
  uint64_t batch_serial = 0;
  // per-family device timing
  cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
  float ms[3] = {0, 0, 0};
  bool timing = false;
  void* model = nullptr;                     // ModelCtx (model_host.inl), created by the first lh264_model_open
  void* decode = nullptr;                    // DecodeCtx (model_decode.cuh), created by the first lh264_model_decode
  // host pointer -> device copy of the coefficient pools staged by the last lh264_recon_submit: a model picture that names
  // the same pool (it is the same data) reads the device copy instead of crossing PCIe twice
  struct StagedPool { const int16_t* host; const int16_t* dev; int n_blocks; };
  std::vector<StagedPool> staged_coefs;        // valid for ONE lh264_model_submit: consumed (cleared) by it
  cudaEvent_t recon_staged = nullptr;        // the H2D copies of that batch are done
  // decoded pictures leave on a copy stream of their own, so that this batch's D2H overlaps the next batch's H2D (PCIe is
  // full duplex) and kernels
  cudaStream_t d2h_stream = nullptr;
  cudaEvent_t packed = nullptr, d2h_done = nullptr;
  bool d2h_pending = false;
};
static void model_destroy (lh264_ctx* c);
static void decode_destroy (lh264_ctx* c);
static float decode_ms (lh264_ctx* c, bool reset);
static float model_ms (lh264_ctx* c, int which, bool reset);
static cudaError_t model_sync (lh264_ctx* c);
static cudaError_t model_join (lh264_ctx* c);
static int model_settle_sharer (lh264_ctx* c, int k);
static cudaError_t model_inputs_consumed (lh264_ctx* c);
static cudaError_t model_order_recon (lh264_ctx* c);

extern "C" {

int lh264_abi_version (void) { return LH264_ABI_VERSION; }
const char* lh264_last_error (void) { return g_err; }

int lh264_ctx_create (lh264_ctx** out, int device, void* cuda_stream) {
  if (!out) return fail (LH264_E_INVALID, "null out");
  *out = nullptr;
  int n = 0;
  cudaError_t e = cudaGetDeviceCount (&n);
  if (e != cudaSuccess || n <= 0) return fail (LH264_E_NO_DEVICE, "no CUDA device (%s); this library has no CPU fallback", cudaGetErrorString (e));
  if (device < 0 || device >= n) return fail (LH264_E_INVALID, "device %d out of range (%d devices)", device, n);
  CK (cudaSetDevice (device));
  lh264_ctx* c = new (std::nothrow) lh264_ctx();
  if (!c) return fail (LH264_E_NOMEM, "out of host memory");
  c->device = device;
  cudaDeviceGetAttribute (&c->n_sm, cudaDevAttrMultiProcessorCount, device);
  cudaOccupancyMaxActiveBlocksPerMultiprocessor (&c->occ_recon, k_decode_rows, ROWS_WARPS * 32, 0);
  if (c->occ_recon < 1) c->occ_recon = 1;
  CK (cudaMalloc (&c->d_error, sizeof (unsigned int)));
  CK (cudaMemset (c->d_error, 0, sizeof (unsigned int)));
  if (cuda_stream) c->stream = (cudaStream_t)cuda_stream;
  else { CK (cudaStreamCreateWithFlags (&c->stream, cudaStreamNonBlocking)); c->own_stream = true; }
  CK (cudaEventCreateWithFlags (&c->staged, cudaEventDisableTiming));
  CK (cudaEventCreateWithFlags (&c->recon_staged, cudaEventDisableTiming));
  CK (cudaStreamCreateWithFlags (&c->d2h_stream, cudaStreamNonBlocking));
  CK (cudaStreamCreateWithFlags (&c->h2d_stream, cudaStreamNonBlocking));
  CK (cudaEventCreateWithFlags (&c->uploaded, cudaEventDisableTiming));
  for (int i = 0; i < 2; i++) {
    CK (cudaEventCreateWithFlags (&c->pack_copied[i], cudaEventDisableTiming)); CK (cudaEventCreateWithFlags (&c->pack_read[i], cudaEventDisableTiming));
    CK (cudaEventCreateWithFlags (&c->decode_done[i], cudaEventDisableTiming));
  }
  CK (cudaEventCreateWithFlags (&c->packed, cudaEventDisableTiming));
  CK (cudaEventCreateWithFlags (&c->d2h_done, cudaEventDisableTiming));
  for (int i = 0; i < 2; i++) CK (cudaEventCreateWithFlags (&c->batch_free[i], cudaEventDisableTiming));
  for (int i = 0; i < 4; i++) CK (cudaEventCreate (&c->ev[i]));
  *out = c;
  return LH264_OK;
}

void lh264_ctx_destroy (lh264_ctx* c) {
  if (!c) return;
  cudaSetDevice (c->device);
  cudaStreamSynchronize (c->stream);
  model_destroy (c);
  decode_destroy (c);
  c->slabs.release();                        // (picture slots, line buffers and the model streams' buffers live there)
  if (c->d_error) cudaFree (c->d_error);
  c->batch[0].release(); c->batch[1].release(); c->h_pics.release(); c->h_xjobs.release();
  for (int i = 0; i < 2; i++) if (c->batch_free[i]) cudaEventDestroy (c->batch_free[i]);
  c->pack_jobs.release();
  for (int i = 0; i < 2; i++) {
    c->h_pack_jobs[i].release(); c->pack_out[i].release();
    if (c->pack_copied[i]) cudaEventDestroy (c->pack_copied[i]);
    if (c->pack_read[i]) cudaEventDestroy (c->pack_read[i]);
    if (c->decode_done[i]) cudaEventDestroy (c->decode_done[i]);
  }
  if (c->uploaded) cudaEventDestroy (c->uploaded);
  if (c->h2d_stream) { cudaStreamSynchronize (c->h2d_stream); cudaStreamDestroy (c->h2d_stream); }
  c->scratch_a.release(); c->scratch_b.release(); c->scratch_c.release();
  if (c->staged) cudaEventDestroy (c->staged);
  if (c->recon_staged) cudaEventDestroy (c->recon_staged);
  if (c->d2h_stream) { cudaStreamSynchronize (c->d2h_stream); cudaStreamDestroy (c->d2h_stream); }
  if (c->packed) cudaEventDestroy (c->packed);
  if (c->d2h_done) cudaEventDestroy (c->d2h_done);
  for (int i = 0; i < 4; i++) if (c->ev[i]) cudaEventDestroy (c->ev[i]);
  if (c->own_stream) cudaStreamDestroy (c->stream);
  delete c;
}

int lh264_stream_open (lh264_ctx* c, int mb_w, int mb_h, int n_slots) {
  if (!c || mb_w <= 0 || mb_h <= 0 || n_slots <= 0 || n_slots > 64) return fail (LH264_E_INVALID, "bad stream geometry");
  CK (cudaSetDevice (c->device));
  StreamRec s;
  s.open = true; s.mb_w = mb_w; s.mb_h = mb_h; s.n_slots = n_slots;
  s.slot_stride = (size_t)mb_w * mb_h * 384;
  // + 64: the aligned word loads of the reconstruction kernel may touch up to 3 bytes past the last row of the last plane
  CK (c->slabs.take ((void**)&s.base, s.slot_stride * n_slots + 64));
  CK (cudaMemsetAsync (s.base, 128, s.slot_stride * n_slots + 64, c->stream));   // pic_queue.cpp:96
  if (c->slabs.take ((void**)&s.lines, (size_t)mb_h * 2 * mb_w * 16) != cudaSuccess) { cudaGetLastError(); c->slabs.give (s.base, s.slot_stride * n_slots + 64); return fail (LH264_E_NOMEM, "out of device memory (line buffer)"); }
  for (size_t i = 0; i < c->streams.size(); i++) if (!c->streams[i].open) { c->streams[i] = s; return (int)i; }
  c->streams.push_back (s);
  return (int)c->streams.size() - 1;
}

int lh264_stream_close (lh264_ctx* c, int stream) {
  if (!c || stream < 0 || stream >= (int)c->streams.size() || !c->streams[stream].open) return fail (LH264_E_INVALID, "bad stream");
  CK (cudaStreamSynchronize (c->stream));
  {
    const StreamRec& s = c->streams[stream];
    c->slabs.give (s.base, s.slot_stride * s.n_slots + 64);
    c->slabs.give (s.lines, (size_t)s.mb_h * 2 * s.mb_w * 16);
  }
  c->streams[stream] = StreamRec();
  return LH264_OK;
}

static int check_pic (lh264_ctx* c, const lh264_picture_desc& p, bool need_mbs = true) {
  if (p.stream < 0 || p.stream >= (int)c->streams.size() || !c->streams[p.stream].open) return fail (LH264_E_INVALID, "picture names a closed stream");
  const StreamRec& s = c->streams[p.stream];
  if (p.cur_slot < 0 || p.cur_slot >= s.n_slots) return fail (LH264_E_INVALID, "cur_slot out of range");
  if (p.n_slices <= 0 || !p.slices || (need_mbs && !p.mbs) || (p.n_coef_blocks > 0 && !p.coefs) || p.n_coef_blocks < 0) return fail (LH264_E_INVALID, "malformed picture_desc");
  return LH264_OK;
}

// Host -> device copies of a batch are laid out back to back on the device; when the caller's buffers are back to back on the
// host too (a staging arena), consecutive copies are merged into one cudaMemcpyAsync.
struct CopyMerger {
  cudaStream_t st; char* d = nullptr; const char* h = nullptr; size_t n = 0; cudaError_t err = cudaSuccess; int issued = 0;
  explicit CopyMerger (cudaStream_t s) : st (s) {}
  void add (void* dev, const void* host, size_t bytes) {
    if (!bytes) return;
    if (n && d + n == (char*)dev && h + n == (const char*)host) { n += bytes; return; }
    flush();
    d = (char*)dev; h = (const char*)host; n = bytes;
  }
  void flush() {
    if (n && err == cudaSuccess) { err = cudaMemcpyAsync (d, h, n, cudaMemcpyHostToDevice, st); issued++; }
    n = 0;
  }
};

// Copies the records of a batch into HBM (async) and fills the device picture table.
static int stage_batch (lh264_ctx* c, BatchDev& b, PinBuf<PicDev>& h_pics, const lh264_picture_desc* pics, int n, const lh264_compact_picture* cp = nullptr) {
  size_t tot_mbs = 0, tot_blk = 0, tot_sl = 0, tot_x = 0;
  for (int i = 0; i < n; i++) {
    int rc = check_pic (c, pics[i], cp == nullptr);
    if (rc) return rc;
    if (cp) {
      if (!cp[i].cmbs || cp[i].n_mb_extra < 0 || (cp[i].n_mb_extra > 0 && !cp[i].mb_extra)) return fail (LH264_E_INVALID, "malformed compact picture");
      tot_x += (size_t)cp[i].n_mb_extra;
    }
    const StreamRec& s = c->streams[pics[i].stream];
    tot_mbs += (size_t)s.mb_w * s.mb_h; tot_blk += pics[i].n_coef_blocks; tot_sl += pics[i].n_slices;
  }
  CK (b.mbs.reserve (tot_mbs)); CK (b.coefs.reserve (tot_blk * 16 + 16)); CK (b.slices.reserve (tot_sl));
  CK (b.pics.reserve (n));
  b.compact = cp != nullptr;
  if (cp) { CK (b.cmb.reserve (tot_mbs)); CK (b.xmb.reserve (tot_x * 16 + 16)); CK (b.xjobs.reserve (n)); }
  if (c->staged_pending) { CK (cudaEventSynchronize (c->staged)); c->staged_pending = false; }
  CK (h_pics.reserve (n));
  if (cp) CK (c->h_xjobs.reserve (n));
  size_t o_mb = 0, o_blk = 0, o_sl = 0, o_x = 0;
  b.n_pics = n; b.max_rows = 0; b.max_mbs = 0;
  CopyMerger cm_cmb (c->h2d_stream), cm_x (c->h2d_stream);
  CopyMerger cm_mb (c->h2d_stream), cm_co (c->h2d_stream), cm_sl (c->h2d_stream);
  for (int i = 0; i < n; i++) {
    const lh264_picture_desc& p = pics[i];
    const StreamRec& s = c->streams[p.stream];
    const size_t nmb = (size_t)s.mb_w * s.mb_h;
    if (!cp) cm_mb.add (b.mbs.p + o_mb, p.mbs, nmb * sizeof (lh264_mb_rec));
    else {
      cm_cmb.add (b.cmb.p + o_mb, cp[i].cmbs, nmb * sizeof (lh264_cmb));
      cm_x.add (b.xmb.p + o_x * 16, cp[i].mb_extra, (size_t)cp[i].n_mb_extra * 16);
      c->h_xjobs.p[i] = ExpandJob{ b.cmb.p + o_mb, b.xmb.p + o_x * 16, b.mbs.p + o_mb, (int)nmb, 0 };
      o_x += (size_t)cp[i].n_mb_extra;
      b.max_mbs = std::max (b.max_mbs, (int)nmb);
    }
    cm_co.add (b.coefs.p + o_blk * 16, p.coefs, (size_t)p.n_coef_blocks * 32);
    cm_sl.add (b.slices.p + o_sl, p.slices, (size_t)p.n_slices * sizeof (lh264_slice_rec));
    PicDev& d = h_pics.p[i];
    d.mbs = b.mbs.p + o_mb; d.coefs = b.coefs.p + o_blk * 16; d.slices = b.slices.p + o_sl;
    d.base = s.base; d.slot_stride = s.slot_stride; d.cur_slot = p.cur_slot; d.n_slots = s.n_slots; d.mb_w = s.mb_w; d.mb_h = s.mb_h;
    d.lines = s.lines; d.n_slices = p.n_slices; d.n_coef_blocks = p.n_coef_blocks;
    b.max_rows = std::max (b.max_rows, s.mb_h);
    o_mb += nmb; o_blk += p.n_coef_blocks; o_sl += p.n_slices;
  }
  cm_mb.flush(); cm_co.flush(); cm_sl.flush(); cm_cmb.flush(); cm_x.flush();
  CK (cm_mb.err); CK (cm_co.err); CK (cm_sl.err); CK (cm_cmb.err); CK (cm_x.err);
  if (cp) CK (cudaMemcpyAsync (b.xjobs.p, c->h_xjobs.p, n * sizeof (ExpandJob), cudaMemcpyHostToDevice, c->h2d_stream));
  CK (cudaMemcpyAsync (b.pics.p, h_pics.p, n * sizeof (PicDev), cudaMemcpyHostToDevice, c->h2d_stream));
  CK (cudaEventRecord (c->staged, c->h2d_stream));
  c->staged_pending = true;
  CK (cudaEventRecord (c->uploaded, c->h2d_stream));
  CK (cudaStreamWaitEvent (c->stream, c->uploaded, 0));          // the kernels of this batch start when its records are in HBM
  CK (b.row_sync.reserve (4 + (size_t)n * b.max_rows));
  return LH264_OK;
}

static int launch_recon (lh264_ctx* c, BatchDev& b) {
  if (b.n_pics == 0) return LH264_OK;
  // one persistent launch: enough warps to fill the machine, never more than there are rows
  const size_t sync_bytes = (4 + (size_t)b.n_pics * b.max_rows) * sizeof (int);
  const long long rows = (long long)b.n_pics * b.max_rows;
  const int blocks = (int)std::max (1LL, std::min ((long long)c->n_sm * c->occ_recon, (rows + ROWS_WARPS - 1) / ROWS_WARPS));
  CK (model_order_recon (c));
  if (c->timing) CK (cudaEventRecord (c->ev[0], c->stream));
  if (b.compact) {                            // the records arrived in compact form: expand them where the row kernel reads them
    k_expand_mb<<<dim3 ((unsigned)std::min ((b.max_mbs + 127) / 128, 64), (unsigned)b.n_pics), 128, 0, c->stream>>> (b.xjobs.p);
    c->launches++;
    b.compact = false;
  }
  CK (cudaMemsetAsync (b.row_sync.p, 0, sync_bytes, c->stream));
  k_decode_rows<<<blocks, ROWS_WARPS * 32, 0, c->stream>>> (b.pics.p, b.n_pics, b.max_rows, reinterpret_cast<DecodeSync*> (b.row_sync.p), c->d_error);
  c->launches++;
  if (c->timing) {
    CK (cudaEventRecord (c->ev[1], c->stream));
    CK (cudaEventSynchronize (c->ev[1]));
    float t = 0; cudaEventElapsedTime (&t, c->ev[0], c->ev[1]); c->ms[0] += t;
  }
  CK (cudaGetLastError());
  return LH264_OK;
}

static int recon_submit_impl (lh264_ctx* c, const lh264_picture_desc* pics, int n_pics, const lh264_compact_picture* cp) {
  if (!c || (n_pics > 0 && !pics) || n_pics < 0) return fail (LH264_E_INVALID, "bad arguments");
  CK (cudaSetDevice (c->device));
  // two pictures of one stream in a batch would race on cur_slot (and on the stream's references): refuse, at every batch size
  c->batch_serial++;
  for (int i = 0; i < n_pics; i++) {
    const int s = pics[i].stream;
    if (s < 0 || s >= (int)c->streams.size() || !c->streams[s].open) return fail (LH264_E_INVALID, "picture names a closed stream");
    if (c->streams[s].batch_stamp == c->batch_serial) return fail (LH264_E_INVALID, "two pictures of one stream in a batch");
    c->streams[s].batch_stamp = c->batch_serial;
  }
  c->staged_coefs.clear();
  c->batch_cur ^= 1;
  const int k = c->batch_cur;
  if (c->batch_shared[k]) {
    // the model batch that reads this staging buffer's coefficient pools must have run for good (not skipped for lack of
    // space and waiting to be run again) before the pools are overwritten
    int rs = model_settle_sharer (c, k);
    if (rs) return rs;
    CK (cudaStreamWaitEvent (c->h2d_stream, c->batch_free[k], 0));
    c->batch_shared[k] = false;
  }
  // ... and so must the reconstruction kernel that read this staging buffer two submits ago
  if (c->decode_done_pending[k]) { CK (cudaStreamWaitEvent (c->h2d_stream, c->decode_done[k], 0)); c->decode_done_pending[k] = false; }
  int rc = stage_batch (c, c->batch[k], c->h_pics, pics, n_pics, cp);
  if (rc) return rc;
  for (int i = 0; i < n_pics; i++) if (pics[i].n_coef_blocks) c->staged_coefs.push_back (lh264_ctx::StagedPool{pics[i].coefs, c->h_pics.p[i].coefs, pics[i].n_coef_blocks});
  CK (cudaEventRecord (c->recon_staged, c->h2d_stream));
  rc = launch_recon (c, c->batch[k]);
  if (rc) return rc;
  CK (cudaEventRecord (c->decode_done[k], c->stream));
  c->decode_done_pending[k] = true;
  return LH264_OK;
}

int lh264_recon_submit (lh264_ctx* c, const lh264_picture_desc* pics, int n_pics) { return recon_submit_impl (c, pics, n_pics, nullptr); }

int lh264_batch_upload (lh264_ctx* c, const lh264_picture_desc* pics, int n_pics, lh264_resident_batch** out) {
  if (!c || !pics || n_pics <= 0 || !out) return fail (LH264_E_INVALID, "bad arguments");
  CK (cudaSetDevice (c->device));
  lh264_resident_batch* rb = new (std::nothrow) lh264_resident_batch();
  if (!rb) return fail (LH264_E_NOMEM, "out of host memory");
  int rc = stage_batch (c, rb->dev, c->h_pics, pics, n_pics);
  if (rc) { rb->dev.release(); delete rb; return rc; }
  CK (cudaStreamSynchronize (c->h2d_stream));
  CK (cudaStreamSynchronize (c->stream));
  c->staged_pending = false;
  *out = rb;
  return LH264_OK;
}
int lh264_batch_recon (lh264_ctx* c, lh264_resident_batch* b) {
  if (!c || !b) return fail (LH264_E_INVALID, "bad arguments");
  CK (cudaSetDevice (c->device));
  return launch_recon (c, b->dev);
}
void lh264_batch_free (lh264_ctx* c, lh264_resident_batch* b) {
  if (!b) return;
  if (c) { cudaSetDevice (c->device); cudaStreamSynchronize (c->stream); }
  b->dev.release();
  delete b;
}

uint64_t lh264_kernel_launches (const lh264_ctx* c) { return c ? c->launches : 0; }
float lh264_kernel_ms (lh264_ctx* c, int family) {
  if (!c) return 0.f;
  if (family == -2) { c->timing = false; return 0.f; }
  if (family < 0) { c->ms[0] = c->ms[1] = c->ms[2] = 0.f; model_ms (c, 0, true); decode_ms (c, true); c->timing = true; return 0.f; }
  if (family == 6) return decode_ms (c, false);
  if (family < 3) return c->ms[family];
  if (family == 4 || family == 5) return model_ms (c, family - 4, false);
  return c->ms[0] + c->ms[1] + c->ms[2];
}

static int plane_geom (lh264_ctx* c, int stream, int slot, const StreamRec** sp) {
  if (!c || stream < 0 || stream >= (int)c->streams.size() || !c->streams[stream].open) return fail (LH264_E_INVALID, "bad stream");
  if (slot < 0 || slot >= c->streams[stream].n_slots) return fail (LH264_E_INVALID, "bad slot");
  *sp = &c->streams[stream];
  return LH264_OK;
}

int lh264_picture_read (lh264_ctx* c, int stream, int slot, int crop_x, int crop_y, int width, int height,
                        uint8_t* dst_y, int stride_y, uint8_t* dst_u, uint8_t* dst_v, int stride_c, int sync) {
  const StreamRec* s;
  int rc = plane_geom (c, stream, slot, &s);
  if (rc) return rc;
  const int W = s->mb_w * 16, H = s->mb_h * 16;
  if (crop_x < 0 || crop_y < 0 || width <= 0 || height <= 0 || crop_x + width > W || crop_y + height > H || (crop_x & 1) || (crop_y & 1))
    return fail (LH264_E_INVALID, "bad crop window");
  CK (cudaSetDevice (c->device));
  const uint8_t* y = s->base + (size_t)slot * s->slot_stride;
  const uint8_t* u = y + (size_t)W * H;
  const uint8_t* v = u + (size_t)W * H / 4;
  const int cw = (width + 1) >> 1, ch = (height + 1) >> 1;
  CK (cudaMemcpy2DAsync (dst_y, stride_y, y + (size_t)crop_y * W + crop_x, W, width, height, cudaMemcpyDeviceToHost, c->stream));
  CK (cudaMemcpy2DAsync (dst_u, stride_c, u + (size_t) (crop_y >> 1) * (W >> 1) + (crop_x >> 1), W >> 1, cw, ch, cudaMemcpyDeviceToHost, c->stream));
  CK (cudaMemcpy2DAsync (dst_v, stride_c, v + (size_t) (crop_y >> 1) * (W >> 1) + (crop_x >> 1), W >> 1, cw, ch, cudaMemcpyDeviceToHost, c->stream));
  if (sync) CK (cudaStreamSynchronize (c->stream));
  return LH264_OK;
}

int lh264_picture_write (lh264_ctx* c, int stream, int slot, const uint8_t* src_y, int stride_y,
                         const uint8_t* src_u, const uint8_t* src_v, int stride_c) {
  const StreamRec* s;
  int rc = plane_geom (c, stream, slot, &s);
  if (rc) return rc;
  const int W = s->mb_w * 16, H = s->mb_h * 16;
  CK (cudaSetDevice (c->device));
  uint8_t* y = s->base + (size_t)slot * s->slot_stride;
  uint8_t* u = y + (size_t)W * H;
  uint8_t* v = u + (size_t)W * H / 4;
  CK (cudaMemcpy2DAsync (y, W, src_y, stride_y, W, H, cudaMemcpyHostToDevice, c->stream));
  CK (cudaMemcpy2DAsync (u, W >> 1, src_u, stride_c, W >> 1, H >> 1, cudaMemcpyHostToDevice, c->stream));
  CK (cudaMemcpy2DAsync (v, W >> 1, src_v, stride_c, W >> 1, H >> 1, cudaMemcpyHostToDevice, c->stream));
  CK (cudaStreamSynchronize (c->stream));
  return LH264_OK;
}

int lh264_picture_device_ptrs (lh264_ctx* c, int stream, int slot, void** y, void** u, void** v, int* pitch_y, int* pitch_c) {
  const StreamRec* s;
  int rc = plane_geom (c, stream, slot, &s);
  if (rc) return rc;
  const int W = s->mb_w * 16, H = s->mb_h * 16;
  uint8_t* py = s->base + (size_t)slot * s->slot_stride;
  if (y) *y = py;
  if (u) *u = py + (size_t)W * H;
  if (v) *v = py + (size_t)W * H + (size_t)W * H / 4;
  if (pitch_y) *pitch_y = W;
  if (pitch_c) *pitch_c = W >> 1;
  return LH264_OK;
}

// Gathers the cropped I420 pictures of n (stream, slot) pairs into one pinned host buffer with a
// single D2H copy: dst layout = for each picture Y (width x height), U, V ((w+1)/2 x (h+1)/2).
int lh264_pictures_read_packed (lh264_ctx* c, const int* streams, const int* slots, int n, int crop_x, int crop_y,
                                int width, int height, uint8_t* dst, size_t dst_bytes, int sync) {
  if (!c || !streams || !slots || n <= 0 || !dst) return fail (LH264_E_INVALID, "bad arguments");
  CK (cudaSetDevice (c->device));
  const int cw = (width + 1) >> 1, ch = (height + 1) >> 1;
  const size_t per = (size_t)width * height + 2 * (size_t)cw * ch;
  if (dst_bytes < per * n) return fail (LH264_E_INVALID, "destination too small");
  const int k = (c->pack_cur ^= 1);
  CK (cudaEventSynchronize (c->pack_copied[k]));      // the job table of two calls ago has long gone up
  CK (c->pack_jobs.reserve (3 * n)); CK (c->h_pack_jobs[k].reserve (3 * n)); CK (c->pack_out[k].reserve (per * n));
  for (int i = 0; i < n; i++) {
    const StreamRec* s;
    int rc = plane_geom (c, streams[i], slots[i], &s);
    if (rc) return rc;
    const int W = s->mb_w * 16, H = s->mb_h * 16;
    if (crop_x < 0 || crop_y < 0 || crop_x + width > W || crop_y + height > H || (crop_x & 1) || (crop_y & 1)) return fail (LH264_E_INVALID, "bad crop window");
    const uint8_t* y = s->base + (size_t)slots[i] * s->slot_stride;
    const uint8_t* u = y + (size_t)W * H;
    const uint8_t* v = u + (size_t)W * H / 4;
    PackDev* j = c->h_pack_jobs[k].p + 3 * i;
    j[0] = PackDev{ y + (size_t)crop_y * W + crop_x, W, width, height, per * i };
    j[1] = PackDev{ u + (size_t) (crop_y >> 1) * (W >> 1) + (crop_x >> 1), W >> 1, cw, ch, per * i + (size_t)width * height };
    j[2] = PackDev{ v + (size_t) (crop_y >> 1) * (W >> 1) + (crop_x >> 1), W >> 1, cw, ch, per * i + (size_t)width * height + (size_t)cw * ch };
  }
  // pack_out[k] was read by the copy to the host of two calls ago
  if (c->pack_read_pending[k]) { CK (cudaStreamWaitEvent (c->stream, c->pack_read[k], 0)); c->pack_read_pending[k] = false; }
  CK (cudaMemcpyAsync (c->pack_jobs.p, c->h_pack_jobs[k].p, 3 * n * sizeof (PackDev), cudaMemcpyHostToDevice, c->stream));
  CK (cudaEventRecord (c->pack_copied[k], c->stream));
  k_pack<<<dim3 (std::min (height, 64), 3 * n), 128, 0, c->stream>>> (c->pack_jobs.p, c->pack_out[k].p);
  c->launches++;
  CK (cudaEventRecord (c->packed, c->stream));
  CK (cudaStreamWaitEvent (c->d2h_stream, c->packed, 0));
  CK (cudaMemcpyAsync (dst, c->pack_out[k].p, per * n, cudaMemcpyDeviceToHost, c->d2h_stream));
  CK (cudaEventRecord (c->pack_read[k], c->d2h_stream));
  c->pack_read_pending[k] = true;
  CK (cudaEventRecord (c->d2h_done, c->d2h_stream));
  c->d2h_pending = true;
  if (sync) CK (cudaStreamSynchronize (c->d2h_stream));
  return LH264_OK;
}

int lh264_output_wait (lh264_ctx* c) {
  if (!c) return fail (LH264_E_INVALID, "null ctx");
  CK (cudaSetDevice (c->device));
  CK (cudaStreamSynchronize (c->d2h_stream));
  return LH264_OK;
}

int lh264_read_ticket (const lh264_ctx* c) { return c ? c->pack_cur : LH264_E_INVALID; }
int lh264_output_wait_ticket (lh264_ctx* c, int ticket) {
  if (!c || ticket < 0 || ticket > 1) return fail (LH264_E_INVALID, "bad ticket");
  CK (cudaSetDevice (c->device));
  CK (cudaEventSynchronize (c->pack_read[ticket]));
  return LH264_OK;
}

int lh264_inputs_consumed (lh264_ctx* c) {
  if (!c) return fail (LH264_E_INVALID, "null ctx");
  CK (cudaSetDevice (c->device));
  if (c->staged_pending) { CK (cudaEventSynchronize (c->staged)); c->staged_pending = false; }
  CK (model_inputs_consumed (c));
  return LH264_OK;
}

int lh264_join (lh264_ctx* c) {
  if (!c) return fail (LH264_E_INVALID, "null ctx");
  CK (cudaSetDevice (c->device));
  CK (model_join (c));
  return LH264_OK;
}

int lh264_sync (lh264_ctx* c) {
  if (!c) return fail (LH264_E_INVALID, "null ctx");
  CK (cudaSetDevice (c->device));
  CK (model_sync (c));
  unsigned int bad = 0;
  CK (cudaMemcpyAsync (&bad, c->d_error, sizeof (bad), cudaMemcpyDeviceToHost, c->stream));
  CK (cudaStreamSynchronize (c->h2d_stream));
  CK (cudaStreamSynchronize (c->stream));
  CK (cudaStreamSynchronize (c->d2h_stream));
  c->staged_pending = false;
  if (bad) {
    CK (cudaMemsetAsync (c->d_error, 0, sizeof (bad), c->stream));
    return fail (LH264_E_INVALID, "malformed macroblock record(s) in a submitted picture (slice_idx, coded_mask / coef_off or mb_type out of range): those macroblocks were left untouched");
  }
  return LH264_OK;
}

void* lh264_host_alloc (size_t bytes) { void* p = nullptr; return cudaMallocHost (&p, bytes) == cudaSuccess ? p : nullptr; }
void* lh264_host_alloc_wc (size_t bytes) { void* p = nullptr; return cudaHostAlloc (&p, bytes, cudaHostAllocWriteCombined) == cudaSuccess ? p : nullptr; }
void lh264_host_free (void* p) { if (p) cudaFreeHost (p); }
int lh264_thread_bind (int device) { return cudaSetDevice (device) == cudaSuccess ? LH264_OK : fail (LH264_E_NO_DEVICE, "cudaSetDevice failed"); }

// ---- kernel-level entry points ---------------------------------------------------------------------------
static int up (lh264_ctx* c, DevBuf<uint8_t>& b, const void* src, size_t bytes) {
  CK (b.reserve (bytes + 16));
  if (src) CK (cudaMemcpyAsync (b.p, src, bytes, cudaMemcpyHostToDevice, c->stream));
  return LH264_OK;
}
int lh264_k_idct4x4_add (lh264_ctx* c, uint8_t* pix, const int16_t* coef, int n) {
  if (!c || !pix || !coef || n <= 0) return fail (LH264_E_INVALID, "bad arguments");
  CK (cudaSetDevice (c->device));
  int rc;
  if ((rc = up (c, c->scratch_a, pix, (size_t)n * 16))) return rc;
  if ((rc = up (c, c->scratch_b, coef, (size_t)n * 32))) return rc;
  k_test_idct4<<<(n + 127) / 128, 128, 0, c->stream>>> (c->scratch_a.p, (const int16_t*)c->scratch_b.p, n);
  c->launches++;
  CK (cudaMemcpyAsync (pix, c->scratch_a.p, (size_t)n * 16, cudaMemcpyDeviceToHost, c->stream));
  CK (cudaStreamSynchronize (c->stream));
  return LH264_OK;
}
int lh264_k_idct8x8_add (lh264_ctx* c, uint8_t* pix, const int16_t* coef, int n) {
  if (!c || !pix || !coef || n <= 0) return fail (LH264_E_INVALID, "bad arguments");
  CK (cudaSetDevice (c->device));
  int rc;
  if ((rc = up (c, c->scratch_a, pix, (size_t)n * 64))) return rc;
  if ((rc = up (c, c->scratch_b, coef, (size_t)n * 128))) return rc;
  k_test_idct8<<<(n + 63) / 64, 64, 0, c->stream>>> (c->scratch_a.p, (const int16_t*)c->scratch_b.p, n);
  c->launches++;
  CK (cudaMemcpyAsync (pix, c->scratch_a.p, (size_t)n * 64, cudaMemcpyDeviceToHost, c->stream));
  CK (cudaStreamSynchronize (c->stream));
  return LH264_OK;
}
static int mc_common (lh264_ctx* c, const uint8_t* ref, int w, int h, const int32_t* blocks, int n, uint8_t* out, bool luma) {
  if (!c || !ref || !blocks || !out || n <= 0 || w <= 0 || h <= 0) return fail (LH264_E_INVALID, "bad arguments");
  CK (cudaSetDevice (c->device));
  int rc;
  if ((rc = up (c, c->scratch_a, ref, (size_t)w * h))) return rc;
  if ((rc = up (c, c->scratch_b, blocks, (size_t)n * 24))) return rc;
  if ((rc = up (c, c->scratch_c, nullptr, (size_t)n * 256))) return rc;
  CK (cudaMemsetAsync (c->scratch_c.p, 0, (size_t)n * 256, c->stream));
  if (luma) k_test_mc_luma<<<n, 256, 0, c->stream>>> (c->scratch_a.p, w, h, (const int32_t*)c->scratch_b.p, c->scratch_c.p);
  else k_test_mc_chroma<<<n, 256, 0, c->stream>>> (c->scratch_a.p, w, h, (const int32_t*)c->scratch_b.p, c->scratch_c.p);
  c->launches++;
  CK (cudaMemcpyAsync (out, c->scratch_c.p, (size_t)n * 256, cudaMemcpyDeviceToHost, c->stream));
  CK (cudaStreamSynchronize (c->stream));
  return LH264_OK;
}
int lh264_k_mc_luma (lh264_ctx* c, const uint8_t* ref, int w, int h, const int32_t* blocks, int n, uint8_t* out) { return mc_common (c, ref, w, h, blocks, n, out, true); }
int lh264_k_mc_chroma (lh264_ctx* c, const uint8_t* ref, int w, int h, const int32_t* blocks, int n, uint8_t* out) { return mc_common (c, ref, w, h, blocks, n, out, false); }
int lh264_k_expand_plane (lh264_ctx* c, const uint8_t* in, int w, int h, int pad, uint8_t* out) {
  if (!c || !in || !out || w <= 0 || h <= 0 || pad < 0) return fail (LH264_E_INVALID, "bad arguments");
  CK (cudaSetDevice (c->device));
  int rc;
  const size_t ob = (size_t) (w + 2 * pad) * (h + 2 * pad);
  if ((rc = up (c, c->scratch_a, in, (size_t)w * h))) return rc;
  if ((rc = up (c, c->scratch_c, nullptr, ob))) return rc;
  k_test_expand<<<148 * 4, 256, 0, c->stream>>> (c->scratch_a.p, w, h, pad, c->scratch_c.p);
  c->launches++;
  CK (cudaMemcpyAsync (out, c->scratch_c.p, ob, cudaMemcpyDeviceToHost, c->stream));
  CK (cudaStreamSynchronize (c->stream));
  return LH264_OK;
}

} // extern "C"

#include "model_host.inl"
#include "model_decode.cuh"
#include "model_decode_host.inl"

static float model_ms (lh264_ctx* c, int which, bool reset) {
  ModelCtx* m = static_cast<ModelCtx*> (c->model);
  if (!m) return 0.f;
  model_coder_idle (c, m);                            // steps still waiting for the coder are coded now (timed if timing is on)
  if (reset) { m->ms_gen = m->ms_code = 0.f; return 0.f; }
  return which == 0 ? m->ms_gen : m->ms_code;
}
