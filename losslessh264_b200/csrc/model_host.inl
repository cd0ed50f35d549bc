// model_host.inl — host side of the lh264_model_* entry points (included by lh264_shim.cu).
// Owns the per-stream model state in HBM (prior table, two frames of context records, coder states,
// output page lists), the output page arena and the staging of a batch of model pictures.

namespace {

using namespace lh264m;

struct MStreamHost {
  bool open = false;
  int max_mbs = 0, with_8x8 = 0;
  MState* d_state = nullptr;
  uint32_t* d_priors = nullptr;
  CtxRec* d_frames = nullptr;      // frame[0], frame[1], newrec
  int* d_order = nullptr;
  uint32_t* d_pagelist = nullptr;
  unsigned long long decisions = 0;
  uint64_t batch_stamp = 0;        // serial number of the last batch that named this stream (duplicate check, O(n))
};

struct MBatch {                    // one batch of model pictures in HBM
  DevBuf<lh264_model_mb> recs; DevBuf<lh264_model_slice> slices; DevBuf<int16_t> coefs;
  DevBuf<MPicDev> pics; PinBuf<MPicDev> h_pics;
  DevBuf<lh264_cmm> cmm; DevBuf<uint8_t> xmm; DevBuf<ExpandJob> xjobs; PinBuf<ExpandJob> h_xjobs;   // compact hand-off, expanded into recs
  bool compact = false; int max_recs = 0;
  DevBuf<uint32_t> rec_cnt; DevBuf<uint32_t> rec_off; DevBuf<uint32_t> pic_total; DevBuf<unsigned long long> dec_base;
  PinBuf<uint32_t> h_total; PinBuf<unsigned long long> h_base;
  DevBuf<BatchCtl> ctl; PinBuf<BatchCtl> h_ctl;
  bool pending = false;            // launched, totals not yet read back (model_settle)
  bool resident = false;           // staged once by lh264_model_batch_upload and replayed
  cudaEvent_t done = nullptr;      // recorded after the batch's last kernel
  cudaEvent_t staged = nullptr;    // recorded after the batch's host-to-device copies
  std::vector<int> stream_of;
  int n = 0, total_recs = 0;
  int shares_recon_batch = -1;     // index of the recon staging batch whose coefficient pools this batch reads
  void release() { recs.release(); slices.release(); coefs.release(); pics.release(); h_pics.release(); rec_cnt.release(); rec_off.release();
                   cmm.release(); xmm.release(); xjobs.release(); h_xjobs.release();
                   pic_total.release(); dec_base.release(); h_total.release(); h_base.release(); ctl.release(); h_ctl.release();
                   if (done) { cudaEventDestroy (done); done = nullptr; }
                   if (staged) { cudaEventDestroy (staged); staged = nullptr; } }
};

struct ModelCtx {
  std::vector<MStreamHost> streams;
  // output arena
  std::vector<uint8_t*> segments;
  std::vector<uint8_t*> h_page_ptrs;
  uint8_t** d_page_ptrs = nullptr; size_t page_ptr_cap = 0;
  unsigned int* d_next = nullptr;
  unsigned int n_pages = 0, reserved_pages = 0;
  uint64_t batch_serial = 0;
  MBatch batch[2];                 // staging of lh264_model_submit, alternating: batch t is staged while t-1 still runs
  int batch_cur = 0;
  std::vector<MBatch*> inflight;   // launched batches whose totals the host has not read yet, oldest first (at most 2)
  unsigned int* d_poison = nullptr;  // set by a batch that skipped its coding: later batches skip entirely (model_settle)
  cudaStream_t aux = nullptr;      // small D2H reads that must not queue behind running model kernels
  // decision pool, in two halves: the generation kernels fill one while the coder (on code_stream) works through the other
  DevBuf<uint32_t> dec[2];
  // deferred coding (model_kernels.cuh CodeJob): up to code_steps steps are generated before ONE k_model_code launch codes them
  int code_steps = 16;             // LH264_CODE_STEPS
  size_t dec_cap_limit = 0;        // decisions the pool may grow to for the sake of deferral (a single step may exceed it)
  DevBuf<CodeJob> jobs[2]; int job_stride = 0;     // per half: [code_steps][job_stride]
  int cur = 0;                                     // the half being filled
  cudaStream_t code_stream = nullptr;              // k_model_code runs here, beside the generation of the following steps
  cudaEvent_t gen_done = nullptr, coder_done[2] = {nullptr, nullptr}, recon_idle = nullptr, coder_start = nullptr;
  bool coder_pending[2] = {false, false};
  // The coder launch (one block of two latency-bound warps per stream, resident for the whole window) and the persistent
  // reconstruction kernel (two 256-thread blocks per SM: the whole register file) do not fit on an SM together: left to the
  // block scheduler they starve each other (measured, profiles/r2_summary.md: 58 ms per step together, 52 ms one after the
  // other).  LH264_CODE_EXCLUSIVE=0 lets them overlap again.
  // When the steps come from resident batches (nothing but the GPU sets the pace) the coder therefore never shares the GPU with
  // the reconstruction kernel: it waits for the reconstruction kernels submitted so far and those of the following steps wait for it
  // (the generation kernels of the following window run beside it: mode 3 below).  When the steps are fed from host buffers the
  // GPU has slack and a 330 ms bubble per window would only stall the host pipeline: there they share the GPU.
  // LH264_CODE_EXCLUSIVE=0 / 1 (reconstruction waits) / 2 (reconstruction and generation wait) / 3 (reconstruction waits; the
  // generation kernels of the NEXT window wait for the coder to START and run beside it - they are HBM-bound, the coder is a
  // dependency chain, and neither reads what the other writes - so a window costs reconstruction + max (coder, generation)
  // instead of the sum of the three) overrides.
  int exclusive_env = -1;
  bool gen_with_coder = false;
  size_t max_inflight_resident = 2;
  bool beside_coder = false;                       // mode 3: a coder launch is in front of the window being generated (else the window is chained as in mode 2)
  bool last_resident = false;                      // the step submitted last came from a resident batch
  bool exclusive = true;
  bool exclusive_gen = true;
  bool partitioned = false;                        // code_stream belongs to an SM partition (model_partition_stream)
  int pend_steps = 0;
  std::vector<std::vector<int> > pend_streams;     // model stream of every job of a pending step
  std::vector<int> slot_of;                        // scratch: model stream -> block of the coder launch
  unsigned long long* d_pool_used = nullptr; unsigned int* d_pages_pending = nullptr;   // [2] each
  DevBuf<int> job_index; PinBuf<int> h_index[2]; cudaEvent_t index_copied[2] = {nullptr, nullptr}; int index_cur = 0;
  unsigned long long est_pool_used = 0, last_step_total = 0;
  unsigned long long max_step_total = 0;
  // LH264_STATS=1: one line per coder launch / pool change on stderr (how long the windows are is what sets the coder's cost)
  bool stats = false; unsigned long long n_flush = 0, n_grow = 0, n_redo = 0;    // largest step seen: what a pool half is sized for (code_steps of them, memory permitting)
  unsigned long long pool_step_total = 0;   // the step size the pool was last sized for (model_grow_pool)
  cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
  float ms_gen = 0, ms_code = 0;
  // The model runs on its own CUDA stream: it consumes syntax elements, not samples, so the model of picture t overlaps
  // the reconstruction of picture t.  `done` makes the context stream (and whoever records events on it) wait for it.
  cudaStream_t stream = nullptr;
  cudaEvent_t done = nullptr;
  bool done_recorded = false;
};

constexpr unsigned int SEG_PAGES = 2048;         // 32 MB arena segments
constexpr size_t MAX_ARENA_PAGES = 1u << 21;     // 32 GB of output pages (grown in segments on demand; the table of page pointers is 16 MB)

int model_grow_arena (lh264_ctx* c, ModelCtx* m, unsigned int need_pages, unsigned int wish_pages = 0);
int model_grow_pool (lh264_ctx* c, ModelCtx* m, unsigned long long step_total);

} // namespace

struct lh264_model_batch { MBatch b; };

// A CUDA stream whose kernels run on a partition of `n_sms` SMs only (driver green contexts, CUDA 12.4+; fetched through
// cudaGetDriverEntryPoint, so nothing links against libcuda).  The deferred coder - one block of two latency-bound warps per stream,
// resident for a whole window - is confined to a quarter of the GPU this way, and the reconstruction / generation kernels of the
// following steps keep the other SMs to themselves instead of being starved SM by SM.  nullptr = not
// available: the caller falls back to an ordinary stream and to running coder and reconstruction one after the other.
static cudaStream_t model_partition_stream (int device, int n_sms) {
  typedef CUresult (*fn_dev_get) (CUdevice*, int);
  typedef CUresult (*fn_get_res) (CUdevice, CUdevResource*, CUdevResourceType);
  typedef CUresult (*fn_split) (CUdevResource*, unsigned int*, const CUdevResource*, CUdevResource*, unsigned int, unsigned int);
  typedef CUresult (*fn_desc) (CUdevResourceDesc*, CUdevResource*, unsigned int);
  typedef CUresult (*fn_gctx) (CUgreenCtx*, CUdevResourceDesc, CUdevice, unsigned int);
  typedef CUresult (*fn_gstream) (CUstream*, CUgreenCtx, unsigned int, int);
  auto drv = [] (const char* name) -> void* {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint (name, &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) { cudaGetLastError(); return nullptr; }
    return p;
  };
  fn_dev_get dev_get = (fn_dev_get)drv ("cuDeviceGet");
  fn_get_res get_res = (fn_get_res)drv ("cuDeviceGetDevResource");
  fn_split split = (fn_split)drv ("cuDevSmResourceSplitByCount");
  fn_desc gen_desc = (fn_desc)drv ("cuDevResourceGenerateDesc");
  fn_gctx gctx_create = (fn_gctx)drv ("cuGreenCtxCreate");
  fn_gstream gstream_create = (fn_gstream)drv ("cuGreenCtxStreamCreate");
  if (!dev_get || !get_res || !split || !gen_desc || !gctx_create || !gstream_create) return nullptr;
  CUdevice dev;
  CUdevResource all, part, rest;
  unsigned int groups = 1;
  CUdevResourceDesc desc;
  CUgreenCtx g;
  CUstream st = nullptr;
  if (dev_get (&dev, device) != CUDA_SUCCESS) return nullptr;
  if (get_res (dev, &all, CU_DEV_RESOURCE_TYPE_SM) != CUDA_SUCCESS) return nullptr;
  if (split (&part, &groups, &all, &rest, 0, (unsigned int)n_sms) != CUDA_SUCCESS || groups < 1) return nullptr;
  if (gen_desc (&desc, &part, 1) != CUDA_SUCCESS) return nullptr;
  if (gctx_create (&g, desc, dev, CU_GREEN_CTX_DEFAULT_STREAM) != CUDA_SUCCESS) return nullptr;
  if (gstream_create (&st, g, CU_STREAM_NON_BLOCKING, 0) != CUDA_SUCCESS) return nullptr;
  if (getenv ("LH264_DEBUG")) fprintf (stderr, "lh264: coder partition of %u SMs (asked for %d)\n", part.sm.smCount, n_sms);
  return (cudaStream_t)st;                            // (the green context lives as long as the process)
}

static ModelCtx* model_ctx (lh264_ctx* c) {
  if (!c->model) {
    ModelCtx* m = new (std::nothrow) ModelCtx();
    if (!m) return nullptr;
    for (int i = 0; i < 4; i++) cudaEventCreate (&m->ev[i]);
    cudaStreamCreateWithFlags (&m->stream, cudaStreamNonBlocking);
    cudaStreamCreateWithFlags (&m->aux, cudaStreamNonBlocking);
    cudaMalloc (&m->d_poison, sizeof (unsigned int));
    cudaMemset (m->d_poison, 0, sizeof (unsigned int));
    cudaFuncSetAttribute (k_model_count, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) (GEN_THREADS * sizeof (ThreadMb)));
    cudaFuncSetAttribute (k_model_emit, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) (GEN_THREADS * sizeof (ThreadMb)));
    cudaEventCreateWithFlags (&m->done, cudaEventDisableTiming);
    cudaMalloc (&m->d_pool_used, 2 * sizeof (unsigned long long));
    cudaMalloc (&m->d_pages_pending, 2 * sizeof (unsigned int));
    cudaMemset (m->d_pool_used, 0, 2 * sizeof (unsigned long long));
    cudaMemset (m->d_pages_pending, 0, 2 * sizeof (unsigned int));
    // LH264_CODER_SMS=n: the coder gets a partition of n SMs and runs beside the other kernels (0 / unset: no partition)
    if (const char* e = getenv ("LH264_CODER_SMS")) {
      const int n = atoi (e);
      if (n > 0) { m->code_stream = model_partition_stream (c->device, n); if (m->code_stream) m->partitioned = true; }
    }
    if (!m->code_stream) cudaStreamCreateWithFlags (&m->code_stream, cudaStreamNonBlocking);
    cudaEventCreateWithFlags (&m->gen_done, cudaEventDisableTiming);
    cudaEventCreateWithFlags (&m->recon_idle, cudaEventDisableTiming);
    cudaEventCreateWithFlags (&m->coder_start, cudaEventDisableTiming);
    for (int k = 0; k < 2; k++) { cudaEventCreateWithFlags (&m->index_copied[k], cudaEventDisableTiming); cudaEventCreateWithFlags (&m->coder_done[k], cudaEventDisableTiming); }
    if (const char* e = getenv ("LH264_CODE_EXCLUSIVE")) m->exclusive_env = atoi (e);
    if (const char* e = getenv ("LH264_MODEL_INFLIGHT")) m->max_inflight_resident = (size_t)std::max (2, std::min (atoi (e), 12));
    if (const char* e = getenv ("LH264_STATS")) m->stats = atoi (e) != 0;
    if (const char* e = getenv ("LH264_CODE_STEPS")) m->code_steps = std::max (1, std::min (atoi (e), 64));
    {
      size_t free_b = 0, total_b = 0;
      cudaMemGetInfo (&free_b, &total_b);
      size_t limit_mb = 65536;                       // LH264_DEC_POOL_MB: memory the decision pool (both halves) may take for deferral
      if (const char* e = getenv ("LH264_DEC_POOL_MB")) limit_mb = (size_t)std::max (1, atoi (e));
      m->dec_cap_limit = std::min (limit_mb << 20, free_b / 3) / sizeof (uint32_t) / 2;     // per half
    }
    m->pend_streams.resize (m->code_steps);
    c->model = m;
  }
  return static_cast<ModelCtx*> (c->model);
}

// the buffers of a model stream come from the context's slabs (DevSlabs, lh264_shim.cu): sizes as lh264_model_open asked for them
static size_t model_priors_bytes (const MStreamHost& s) { return (s.with_8x8 ? P_END : P_END_NO8) * sizeof (uint32_t); }
static void model_free_stream (lh264_ctx* c, MStreamHost& s) {
  c->slabs.give (s.d_state, sizeof (MState));
  c->slabs.give (s.d_priors, model_priors_bytes (s));
  c->slabs.give (s.d_frames, (size_t)3 * s.max_mbs * sizeof (CtxRec));
  c->slabs.give (s.d_order, (size_t)s.max_mbs * sizeof (int));
  c->slabs.give (s.d_pagelist, (size_t)CT_COUNT * MAX_PAGES_PER_TAG * sizeof (uint32_t));
  s = MStreamHost();
}

static void model_destroy (lh264_ctx* c) {
  ModelCtx* m = static_cast<ModelCtx*> (c->model);
  if (!m) return;
  for (auto& s : m->streams) if (s.open) model_free_stream (c, s);
  for (uint8_t* p : m->segments) cudaFree (p);
  if (m->d_page_ptrs) cudaFree (m->d_page_ptrs);
  if (m->d_next) cudaFree (m->d_next);
  m->batch[0].release(); m->batch[1].release();
  if (m->code_stream) { cudaStreamSynchronize (m->code_stream); cudaStreamDestroy (m->code_stream); }
  for (int k = 0; k < 2; k++) { m->dec[k].release(); m->jobs[k].release(); if (m->coder_done[k]) cudaEventDestroy (m->coder_done[k]); }
  if (m->gen_done) cudaEventDestroy (m->gen_done);
  if (m->recon_idle) cudaEventDestroy (m->recon_idle);
  if (m->coder_start) cudaEventDestroy (m->coder_start);
  m->job_index.release(); m->h_index[0].release(); m->h_index[1].release();
  if (m->d_pool_used) cudaFree (m->d_pool_used);
  if (m->d_pages_pending) cudaFree (m->d_pages_pending);
  for (int k = 0; k < 2; k++) if (m->index_copied[k]) cudaEventDestroy (m->index_copied[k]);
  if (m->d_poison) cudaFree (m->d_poison);
  if (m->aux) cudaStreamDestroy (m->aux);
  for (int i = 0; i < 4; i++) if (m->ev[i]) cudaEventDestroy (m->ev[i]);
  if (m->stream) { cudaStreamSynchronize (m->stream); cudaStreamDestroy (m->stream); }
  if (m->done) cudaEventDestroy (m->done);
  delete m;
  c->model = nullptr;
}

namespace {
int model_grow_arena (lh264_ctx* c, ModelCtx* m, unsigned int need_pages, unsigned int wish_pages) {
  if (!m->d_page_ptrs) {
    CK (cudaMalloc (&m->d_page_ptrs, MAX_ARENA_PAGES * sizeof (uint8_t*)));
    CK (cudaMalloc (&m->d_next, sizeof (unsigned int)));
    CK (cudaMemsetAsync (m->d_next, 0, sizeof (unsigned int), m->stream));
  }
  bool pool_released = false;
  // need_pages: what the step cannot do without; wish_pages: room for a whole window of promises (given up when memory is short)
  const unsigned int target = std::max (need_pages, wish_pages);
  while (m->n_pages < target) {
    if (m->n_pages + SEG_PAGES > MAX_ARENA_PAGES) { if (m->n_pages >= need_pages) break; return fail (LH264_E_NOMEM, "model output arena exhausted"); }
    uint8_t* seg = nullptr;
    cudaError_t e = cudaMalloc (&seg, (size_t)SEG_PAGES * PAGE_SIZE);
    if (e == cudaErrorMemoryAllocation && m->n_pages >= need_pages) { cudaGetLastError(); break; }
    if (e == cudaErrorMemoryAllocation && !pool_released && m->pend_steps == 0 && (m->dec[0].p || m->dec[1].p)) {
      // output pages are a need, the decision pool's room for deferral is a wish: with no step waiting in the pool, hand the
      // pool's memory back, take the pages, and let the pool come back as large as what is left allows (model_grow_pool)
      cudaGetLastError();
      CK (cudaStreamSynchronize (m->stream)); CK (cudaStreamSynchronize (m->code_stream));
      m->dec[0].release(); m->dec[1].release();
      pool_released = true;
      if (getenv ("LH264_DEBUG")) fprintf (stderr, "lh264: output arena out of memory at %u pages: decision pool released\n", m->n_pages);
      continue;
    }
    if (e != cudaSuccess) return fail (LH264_E_CUDA, "cudaMalloc of a model output segment: %s (%u pages of %u)", cudaGetErrorString (e), m->n_pages, need_pages);
    m->segments.push_back (seg);
    const size_t first = m->h_page_ptrs.size();
    for (unsigned int i = 0; i < SEG_PAGES; i++) m->h_page_ptrs.push_back (seg + (size_t)i * PAGE_SIZE);
    CK (cudaMemcpyAsync (m->d_page_ptrs + first, m->h_page_ptrs.data() + first, SEG_PAGES * sizeof (uint8_t*), cudaMemcpyHostToDevice, m->stream));
    CK (cudaStreamSynchronize (m->stream));
    m->n_pages += SEG_PAGES;
  }
  if (pool_released && m->pool_step_total) return model_grow_pool (c, m, m->pool_step_total);
  return LH264_OK;
}
} // namespace

extern "C" {

static int model_settle (lh264_ctx* c, ModelCtx* m);
static int model_settle_batch (lh264_ctx* c, ModelCtx* m, MBatch* only);
static int model_flush_coder (lh264_ctx* c, ModelCtx* m);
static int model_coder_idle (lh264_ctx* c, ModelCtx* m);

int lh264_model_open (lh264_ctx* c, int max_mbs, int with_8x8) {
  if (!c || max_mbs <= 0 || max_mbs > (1 << 20)) return fail (LH264_E_INVALID, "bad model stream geometry");
  CK (cudaSetDevice (c->device));
  ModelCtx* m = model_ctx (c);
  if (!m) return fail (LH264_E_NOMEM, "out of host memory");
  MStreamHost s;
  s.open = true; s.max_mbs = max_mbs; s.with_8x8 = with_8x8 ? 1 : 0;
  const size_t n_priors = with_8x8 ? P_END : P_END_NO8;
  if (c->slabs.take ((void**)&s.d_state, sizeof (MState)) != cudaSuccess || c->slabs.take ((void**)&s.d_priors, n_priors * sizeof (uint32_t)) != cudaSuccess ||
      c->slabs.take ((void**)&s.d_frames, (size_t)3 * max_mbs * sizeof (CtxRec)) != cudaSuccess || c->slabs.take ((void**)&s.d_order, (size_t)max_mbs * sizeof (int)) != cudaSuccess ||
      c->slabs.take ((void**)&s.d_pagelist, (size_t)CT_COUNT * MAX_PAGES_PER_TAG * sizeof (uint32_t)) != cudaSuccess) {
    cudaGetLastError();
    model_free_stream (c, s);                        // (gives back what was had: give() ignores null pointers)
    return fail (LH264_E_NOMEM, "out of device memory (model stream)");
  }
  CK (cudaMemsetAsync (s.d_priors, 0, n_priors * sizeof (uint32_t), m->stream));          // fresh DynProb = all zero bits
  CK (cudaMemsetAsync (s.d_frames, 0, (size_t)3 * max_mbs * sizeof (CtxRec), m->stream));
  CK (cudaMemsetAsync (s.d_pagelist, 0, (size_t)CT_COUNT * MAX_PAGES_PER_TAG * sizeof (uint32_t), m->stream));
  MState h;
  memset (&h, 0, sizeof (h));
  h.priors = s.d_priors; h.frame[0] = s.d_frames; h.frame[1] = s.d_frames + max_mbs; h.newrec = s.d_frames + 2 * (size_t)max_mbs;
  h.order = s.d_order; h.pagelist = s.d_pagelist; h.max_mbs = max_mbs;
  CK (cudaMemcpyAsync (s.d_state, &h, sizeof (h), cudaMemcpyHostToDevice, m->stream));
  k_model_init_state<<<1, 32, 0, m->stream>>> (s.d_state);
  CK (cudaStreamSynchronize (m->stream));                                                  // h is on the stack
  for (size_t i = 0; i < m->streams.size(); i++) if (!m->streams[i].open) { m->streams[i] = s; return (int)i; }
  m->streams.push_back (s);
  return (int)m->streams.size() - 1;
}

static int model_stream_ok (lh264_ctx* c, int ms, ModelCtx** pm) {
  if (!c || !c->model) return fail (LH264_E_INVALID, "no model stream is open");
  ModelCtx* m = static_cast<ModelCtx*> (c->model);
  if (ms < 0 || ms >= (int)m->streams.size() || !m->streams[ms].open) return fail (LH264_E_INVALID, "bad model stream");
  *pm = m;
  return LH264_OK;
}

int lh264_model_close (lh264_ctx* c, int ms) {
  ModelCtx* m;
  int rc = model_stream_ok (c, ms, &m);
  if (rc) return rc;
  CK (cudaSetDevice (c->device));
  { int rs = model_settle (c, m); if (rs) return rs; }
  // pending jobs name the stream's state, and a coder launch may still be working on it
  if ((rc = model_coder_idle (c, m))) return rc;
  CK (cudaStreamSynchronize (m->stream));
  // the stream's output pages stay in the arena until the context is destroyed (bump allocator)
  model_free_stream (c, m->streams[ms]);
  return LH264_OK;
}

// Validates a batch and copies its records into HBM (async); fills the device picture table.
static int model_stage (lh264_ctx* c, ModelCtx* m, MBatch& b, const lh264_model_picture* pics, int n, bool share_coefs, const lh264_compact_picture* cp = nullptr) {
  size_t tot_rec = 0, tot_sl = 0, tot_blk = 0, tot_x = 0;
  m->batch_serial++;
  for (int i = 0; i < n; i++) {
    const lh264_model_picture& p = pics[i];
    ModelCtx* mm;
    int rc = model_stream_ok (c, p.model_stream, &mm);
    if (rc) return rc;
    const MStreamHost& s = m->streams[p.model_stream];
    if (p.mb_w <= 0 || p.mb_h <= 0 || (long long)p.mb_w * p.mb_h > s.max_mbs) return fail (LH264_E_INVALID, "picture larger than the model stream was opened for");
    if (p.n_slices <= 0 || p.n_recs <= 0 || !p.slices || (!cp && !p.recs) || p.n_coef_blocks < 0 || (p.n_coef_blocks > 0 && !p.coefs)) return fail (LH264_E_INVALID, "malformed model picture");
    if (cp) {
      if (!cp[i].cmms || cp[i].n_mm_extra < 0 || (cp[i].n_mm_extra > 0 && !cp[i].mm_extra)) return fail (LH264_E_INVALID, "malformed compact picture");
      tot_x += (size_t)cp[i].n_mm_extra;
    }
    for (int k = 0; k < p.n_slices; k++) {
      const lh264_model_slice& sl = p.slices[k];
      if (sl.first_rec < 0 || sl.n_recs < 0 || sl.first_rec + sl.n_recs > p.n_recs) return fail (LH264_E_INVALID, "slice record range out of bounds");
      if (sl.slice_type != 0 && sl.slice_type != 2) return fail (LH264_E_UNSUPPORTED, "only P and I slices are modelled");
    }
    if (m->streams[p.model_stream].batch_stamp == m->batch_serial) return fail (LH264_E_INVALID, "two pictures of one model stream in a batch");
    m->streams[p.model_stream].batch_stamp = m->batch_serial;
    tot_rec += p.n_recs; tot_sl += p.n_slices; tot_blk += p.n_coef_blocks;
  }
  CK (b.recs.reserve (tot_rec)); CK (b.slices.reserve (tot_sl)); CK (b.coefs.reserve (tot_blk * 16 + 16));
  CK (b.pics.reserve (n)); CK (b.rec_cnt.reserve (tot_rec)); CK (b.rec_off.reserve (tot_rec));
  CK (b.pic_total.reserve (n)); CK (b.dec_base.reserve (n)); CK (b.ctl.reserve (1));
  { int rs = model_settle_batch (c, m, &b); if (rs) return rs; }   // the batch that used this staging before is done and read
  CK (b.h_pics.reserve (n)); CK (b.h_total.reserve (n + 1)); CK (b.h_base.reserve (n)); CK (b.h_ctl.reserve (1));
  b.compact = cp != nullptr; b.max_recs = 0;
  if (cp) { CK (b.cmm.reserve (tot_rec)); CK (b.xmm.reserve (tot_x * 16 + 16)); CK (b.xjobs.reserve (n)); CK (b.h_xjobs.reserve (n)); }
  b.stream_of.assign (n, 0);
  bool any_shared = false;
  size_t o_rec = 0, o_sl = 0, o_blk = 0, o_x = 0;
  CopyMerger cm_rec (m->stream), cm_sl (m->stream), cm_co (m->stream), cm_x (m->stream);
  for (int i = 0; i < n; i++) {
    const lh264_model_picture& p = pics[i];
    const MStreamHost& s = m->streams[p.model_stream];
    if (!cp) cm_rec.add (b.recs.p + o_rec, p.recs, (size_t)p.n_recs * sizeof (lh264_model_mb));
    else {
      cm_rec.add (b.cmm.p + o_rec, cp[i].cmms, (size_t)p.n_recs * sizeof (lh264_cmm));
      cm_x.add (b.xmm.p + o_x * 16, cp[i].mm_extra, (size_t)cp[i].n_mm_extra * 16);
      b.h_xjobs.p[i] = ExpandJob{ b.cmm.p + o_rec, b.xmm.p + o_x * 16, b.recs.p + o_rec, p.n_recs, 0 };
      o_x += (size_t)cp[i].n_mm_extra;
      b.max_recs = std::max (b.max_recs, p.n_recs);
    }
    cm_sl.add (b.slices.p + o_sl, p.slices, (size_t)p.n_slices * sizeof (lh264_model_slice));
    const int16_t* dev_coefs = nullptr;
    // same host pool AND same length as a picture of the preceding recon submit: it is that picture's pool, read its device copy
    if (share_coefs && p.n_coef_blocks && (size_t)i < c->staged_coefs.size() && c->staged_coefs[i].host == p.coefs && c->staged_coefs[i].n_blocks == p.n_coef_blocks) { dev_coefs = c->staged_coefs[i].dev; any_shared = true; }
    else if (share_coefs && p.n_coef_blocks)
      for (const auto& sc : c->staged_coefs) if (sc.host == p.coefs && sc.n_blocks == p.n_coef_blocks) { dev_coefs = sc.dev; any_shared = true; break; }
    if (p.n_coef_blocks && !dev_coefs) cm_co.add (b.coefs.p + o_blk * 16, p.coefs, (size_t)p.n_coef_blocks * 32);
    MPicDev& d = b.h_pics.p[i];
    d.recs = b.recs.p + o_rec; d.slices = b.slices.p + o_sl; d.coefs = dev_coefs ? dev_coefs : b.coefs.p + o_blk * 16;
    d.st = s.d_state; d.rec_cnt = b.rec_cnt.p + o_rec; d.rec_off = b.rec_off.p + o_rec;
    d.n_recs = p.n_recs; d.n_slices = p.n_slices; d.mb_w = p.mb_w; d.mb_h = p.mb_h; d.rec_base = (int)o_rec; d.pad = 0;
    b.stream_of[i] = p.model_stream;
    o_rec += p.n_recs; o_sl += p.n_slices; o_blk += p.n_coef_blocks;
  }
  cm_rec.flush(); cm_sl.flush(); cm_co.flush(); cm_x.flush();
  CK (cm_rec.err); CK (cm_sl.err); CK (cm_co.err); CK (cm_x.err);
  if (cp) {                                           // expanded right behind the copies, on the same stream: the generation kernels see full records
    CK (cudaMemcpyAsync (b.xjobs.p, b.h_xjobs.p, n * sizeof (ExpandJob), cudaMemcpyHostToDevice, m->stream));
    k_expand_mm<<<dim3 ((unsigned)std::min ((b.max_recs + 127) / 128, 64), (unsigned)n), 128, 0, m->stream>>> (b.xjobs.p);
    c->launches++;
  }
  CK (cudaMemcpyAsync (b.pics.p, b.h_pics.p, n * sizeof (MPicDev), cudaMemcpyHostToDevice, m->stream));
  if (!b.staged) CK (cudaEventCreateWithFlags (&b.staged, cudaEventDisableTiming));
  CK (cudaEventRecord (b.staged, m->stream));
  b.shares_recon_batch = any_shared ? c->batch_cur : -1;
  // the device copies are named once: a later model submit with the same host pointers (a reused or overwritten buffer, a
  // model-only caller) must upload its own data instead of silently coding the previous picture's
  if (share_coefs) c->staged_coefs.clear();
  if (any_shared) CK (cudaStreamWaitEvent (m->stream, c->recon_staged, 0));      // the shared pools are in HBM
  b.n = n; b.total_recs = (int)tot_rec;
  return LH264_OK;
}

// Codes every pending step with ONE launch: block b = one model stream, walking its pictures in step order.
#ifndef LH264_RESIDENT_MODE
#define LH264_RESIDENT_MODE 3      // measured, 1024 staggered 1080p streams: 43.7 ms per step in mode 2, 40.2 in mode 3 (profiles/r2_summary.md section 11)
#endif
static int model_mode (const ModelCtx* m) { return m->partitioned ? 0 : (m->exclusive_env >= 0 ? m->exclusive_env : (m->last_resident ? LH264_RESIDENT_MODE : 0)); }
static bool model_serial (const ModelCtx* m) { const int mode = model_mode (m); return mode == 2 || (mode >= 3 && !m->beside_coder); }

static int model_flush_coder (lh264_ctx* c, ModelCtx* m) {
  const int T = m->pend_steps;
  if (T == 0) return LH264_OK;
  const int h = m->cur;
  m->n_flush++;
  if (m->stats) fprintf (stderr, "lh264 stats: coder launch %llu: %d steps, half %d (cap %zu M, est used %llu M, last step %llu M)\n", m->n_flush, T, h, m->dec[h].cap >> 20, m->est_pool_used >> 20, m->last_step_total >> 20);
  {
    const int mode = model_mode (m);
    m->exclusive = mode >= 1 && mode <= 3; m->exclusive_gen = mode == 2; m->gen_with_coder = mode >= 3;     // (4: experiment - as 3, but the reconstruction kernels do not wait either)
  }
  m->slot_of.assign (m->streams.size(), -1);
  int nb = 0;
  for (int j = 0; j < T; j++) for (int s : m->pend_streams[j]) if (m->slot_of[s] < 0) m->slot_of[s] = nb++;
  const size_t n_idx = (size_t)nb * T;
  const int k = (m->index_cur ^= 1);
  CK (cudaEventSynchronize (m->index_copied[k]));      // (the copy of two flushes ago: long done)
  CK (m->h_index[k].reserve (n_idx));
  if (n_idx > m->job_index.cap) { CK (cudaStreamSynchronize (m->code_stream)); CK (m->job_index.reserve (n_idx)); }
  for (size_t i = 0; i < n_idx; i++) m->h_index[k].p[i] = -1;
  for (int j = 0; j < T; j++) for (size_t i = 0; i < m->pend_streams[j].size(); i++) m->h_index[k].p[(size_t)m->slot_of[m->pend_streams[j][i]] * T + j] = (int)i;
  // the coder starts when the generation kernels of its steps are through (and, in exclusive mode, when the reconstruction
  // kernels submitted so far are)
  CK (cudaEventRecord (m->gen_done, m->stream));
  CK (cudaStreamWaitEvent (m->code_stream, m->gen_done, 0));
  if (m->exclusive) { CK (cudaEventRecord (m->recon_idle, c->stream)); CK (cudaStreamWaitEvent (m->code_stream, m->recon_idle, 0)); }
  CK (cudaMemcpyAsync (m->job_index.p, m->h_index[k].p, n_idx * sizeof (int), cudaMemcpyHostToDevice, m->code_stream));
  CK (cudaEventRecord (m->index_copied[k], m->code_stream));
  ArenaDev ar{ m->d_page_ptrs, m->d_next, m->n_pages };
  if (m->gen_with_coder) CK (cudaEventRecord (m->coder_start, m->code_stream));    // everything this launch waits for is through
  if (c->timing) CK (cudaEventRecord (m->ev[2], m->code_stream));
  k_model_code<<<nb, 64, 0, m->code_stream>>> (m->jobs[h].p, m->job_index.p, T, m->job_stride, m->dec[h].p, ar);
  c->launches++;
  if (c->timing) CK (cudaEventRecord (m->ev[3], m->code_stream));
  CK (cudaMemsetAsync (m->d_pool_used + h, 0, sizeof (unsigned long long), m->code_stream));
  CK (cudaMemsetAsync (m->d_pages_pending + h, 0, sizeof (unsigned int), m->code_stream));
  CK (cudaEventRecord (m->coder_done[h], m->code_stream));
  m->coder_pending[h] = true;
  // exclusive: neither the reconstruction kernels nor the generation kernels of the following steps start before the coder is through
  if (m->exclusive) { CK (cudaStreamWaitEvent (c->stream, m->coder_done[h], 0)); if (m->exclusive_gen) CK (cudaStreamWaitEvent (m->stream, m->coder_done[h], 0)); }
  // the generation kernels of the following window run BESIDE this coder launch, not beside the reconstruction kernels in front of it
  if (m->gen_with_coder) CK (cudaStreamWaitEvent (m->stream, m->coder_start, 0));
  m->beside_coder = m->gen_with_coder;
  m->pend_steps = 0; m->est_pool_used = 0;
  // generation goes on in the other half, once the coder launch that read it (two flushes ago) is over
  m->cur ^= 1;
  if (m->coder_pending[m->cur]) { CK (cudaStreamWaitEvent (m->stream, m->coder_done[m->cur], 0)); m->coder_pending[m->cur] = false; }
  if (c->timing) {
    CK (cudaEventSynchronize (m->ev[3]));
    float t = 0;
    cudaEventElapsedTime (&t, m->ev[2], m->ev[3]); m->ms_code += t;
  }
  CK (cudaGetLastError());
  return LH264_OK;
}

// every coder launch is over: the stream states, priors and output pages are final
static int model_coder_idle (lh264_ctx* c, ModelCtx* m) {
  int rc = model_flush_coder (c, m);
  if (rc) return rc;
  CK (cudaStreamSynchronize (m->code_stream));
  m->beside_coder = false;
  return LH264_OK;
}

// The generation launches of one step (one picture per stream): its decisions go into the pool, its jobs into row pend_steps.
static int model_launch (lh264_ctx* c, ModelCtx* m, MBatch& b, bool head, bool timed) {
  const int n = b.n, total_recs = b.total_recs;
  const int mb_blocks = (total_recs + GEN_THREADS - 1) / GEN_THREADS;
  const size_t gen_smem = (size_t)GEN_THREADS * sizeof (ThreadMb);
  if (m->pend_steps >= m->code_steps) { int rc = model_flush_coder (c, m); if (rc) return rc; }
  if (n > m->job_stride) {                            // a wider step than any before: the job table is re-laid
    int rc = model_flush_coder (c, m);
    if (rc) return rc;
    CK (cudaStreamSynchronize (m->stream)); CK (cudaStreamSynchronize (m->code_stream));
    m->job_stride = n + n / 4 + 16;
    for (int k = 0; k < 2; k++) CK (m->jobs[k].reserve ((size_t)m->job_stride * m->code_steps));
  }
  const int h = m->cur;
  ArenaDev ar{ m->d_page_ptrs, m->d_next, m->n_pages };
  PoolDev pool{ m->d_pool_used + h, m->d_pages_pending + h, m->d_pages_pending + (h ^ 1) };
  CodeJob* row = m->jobs[h].p + (size_t)m->pend_steps * m->job_stride;
  CK (cudaMemsetAsync (row, 0, (size_t)n * sizeof (CodeJob), m->stream));       // st == nullptr: a step that skips leaves no jobs
  if (head) {
    if (timed) CK (cudaEventRecord (m->ev[0], m->stream));
    k_model_frame<<<n, 32, 0, m->stream>>> (b.pics.p, n, m->d_poison);
    k_model_count<<<mb_blocks, GEN_THREADS, gen_smem, m->stream>>> (b.pics.p, n, total_recs, m->d_poison);
    k_model_scan<<<n, 1024, 0, m->stream>>> (b.pics.p, n, b.pic_total.p, m->d_poison);
    c->launches += 3;
  }
  k_model_plan<<<1, 1024, 0, m->stream>>> (n, b.pic_total.p, b.dec_base.p, b.ctl.p, (unsigned long long)m->dec[h].cap, ar, pool, m->d_poison);
  k_model_emit<<<mb_blocks, GEN_THREADS, gen_smem, m->stream>>> (b.pics.p, n, total_recs, m->dec[h].p, b.dec_base.p, b.ctl.p);
  k_model_commit<<<n, 128, 0, m->stream>>> (b.pics.p, n, b.dec_base.p, b.pic_total.p, b.ctl.p, row);
  c->launches += 3;
  if (timed) CK (cudaEventRecord (m->ev[1], m->stream));
  m->pend_streams[m->pend_steps] = b.stream_of;
  m->pend_steps++;
  m->est_pool_used += m->last_step_total;
  if (!b.done) CK (cudaEventCreateWithFlags (&b.done, cudaEventDisableTiming));
  CK (cudaEventRecord (b.done, m->stream));
  return LH264_OK;
}

// The pool must hold one step of `step_total` decisions; for the sake of deferral it is given room for code_steps such steps
// as long as that stays below dec_cap_limit.  Call with the pool empty (after model_flush_coder).
extern "C++" {
namespace {
int model_grow_pool (lh264_ctx* c, ModelCtx* m, unsigned long long step_total) {
  (void)c;
  m->pool_step_total = std::max (m->pool_step_total, step_total);
  const size_t one = (size_t)step_total + (size_t)step_total / 4 + 64;
  if (m->dec[m->cur].cap >= one * (size_t)m->code_steps) return LH264_OK;      // room for a full window already
  // what the device can spare NOW (other contexts, picture pools and output pages have come and gone since the context was
  // made): both halves together leave a tenth of the device (at least 4 GB) free for output pages and staging
  size_t budget = m->dec_cap_limit;
  {
    size_t free_b = 0, total_b = 0;
    if (cudaMemGetInfo (&free_b, &total_b) == cudaSuccess) {
      const size_t mine = (m->dec[0].cap + m->dec[1].cap) * sizeof (uint32_t);
      const size_t keep = std::max (total_b / 10, (size_t)4 << 30);
      const size_t avail = free_b + mine > keep ? free_b + mine - keep : 0;
      budget = std::min (budget, avail / sizeof (uint32_t) / 2 * 4 / 5);     // per half; DevBuf::reserve adds a quarter
    }
  }
  size_t want = std::max (one, std::min (one * (size_t)m->code_steps, budget));
  DevBuf<uint32_t>& dec = m->dec[m->cur];
  if (want <= dec.cap) return LH264_OK;
  CK (cudaStreamSynchronize (m->stream)); CK (cudaStreamSynchronize (m->code_stream));
  m->n_grow++;
  if (m->stats || getenv ("LH264_DEBUG")) fprintf (stderr, "lh264: decision pool half %d: %zu -> %zu M decisions (step %llu M, budget %zu M)\n", m->cur, dec.cap >> 20, want >> 20, step_total >> 20, budget >> 20);
  // room for code_steps steps is a wish, room for one step a need: halve the wish until the device has it
  for (;;) {
    if (dec.reserve (want) == cudaSuccess) return LH264_OK;
    cudaGetLastError();
    if (want == one) return fail (LH264_E_NOMEM, "out of device memory for the model's decision pool");
    want = std::max (one, want / 2);
  }
}
} // namespace
} // extern "C++"

// Output pages are PROMISED to a step when it is generated (worst case: 7 bits per decision plus a page per tag and picture) and
// the promise is released when the coder has run, so a window of code_steps pending steps needs code_steps promises at once.
// Growing by one or two steps' worth at a time cut every window short until the arena had crept up to that size (fifteen
// re-runs and as many short coder launches at 1024 streams): ask for the whole window the first time a step does not fit.
static unsigned int model_arena_want (const ModelCtx* m, unsigned int in_use, unsigned int needed) {
  const unsigned long long must = (unsigned long long)in_use + 2ull * needed;
  // (mode 3 generates a window while the coder works on the one before: two windows' promises are out at once)
  const int steps = model_mode (m) >= 3 ? 2 * m->code_steps + 1 : m->code_steps + 1;
  const unsigned long long window = (unsigned long long)in_use + (unsigned long long)std::max (2, steps) * needed;
  const unsigned long long room = MAX_ARENA_PAGES - SEG_PAGES;                     // (segments are added whole)
  return (unsigned int)std::max (must, std::min (window, (unsigned long long)room));
}

// A step skipped itself because the decision pool or the promised output pages did not fit; the pending steps have been handed
// to the coder (model_flush_coder).  Pool and arena are made large enough for it - and, memory permitting, for a whole window.
static int model_make_room (lh264_ctx* c, ModelCtx* m, const BatchCtl& ctl) {
  int rc = model_grow_pool (c, m, ctl.total);
  if (rc) return rc;
  unsigned int in_use = ctl.pages_in_use;
  if ((unsigned long long)in_use + 2ull * ctl.pages_needed > MAX_ARENA_PAGES - SEG_PAGES) {
    // the arena cannot grow that far: the promises counted in `in_use` are released as the coder runs, so wait for it and count
    // the pages that really hold output
    CK (cudaStreamSynchronize (m->code_stream));
    unsigned int used = 0;
    CK (cudaMemcpyAsync (&used, m->d_next, sizeof (used), cudaMemcpyDeviceToHost, m->stream));
    CK (cudaStreamSynchronize (m->stream));
    in_use = used;
  }
  return model_grow_arena (c, m, in_use + 2 * ctl.pages_needed, model_arena_want (m, in_use, ctl.pages_needed));
}

// blocking: runs a batch to completion, growing the decision list / output arena until it fits
static int model_complete (lh264_ctx* c, ModelCtx* m, MBatch& b, bool head) {
  for (int attempt = 0; attempt < 4; attempt++) {
    int rc = model_launch (c, m, b, head, false);
    if (rc) return rc;
    CK (cudaMemcpyAsync (b.h_ctl.p, b.ctl.p, sizeof (BatchCtl), cudaMemcpyDeviceToHost, m->stream));
    CK (cudaMemcpyAsync (b.h_total.p, b.pic_total.p, b.n * sizeof (uint32_t), cudaMemcpyDeviceToHost, m->stream));
    CK (cudaStreamSynchronize (m->stream));
    const BatchCtl ctl = *b.h_ctl.p;
    if (!ctl.skip) return LH264_OK;
    head = false;                                   // frame / count / scan did run: the batch itself was the one that did not fit
    CK (cudaMemsetAsync (m->d_poison, 0, sizeof (unsigned int), m->stream));
    if ((rc = model_flush_coder (c, m))) return rc;  // empties the pool (the step that skipped left no jobs)
    if ((rc = model_make_room (c, m, ctl))) return rc;
  }
  return fail (LH264_E_NOMEM, "model batch does not fit after growing its buffers");
}

static void model_account (ModelCtx* m, MBatch& b) {
  unsigned long long tot = 0;
  for (int i = 0; i < b.n; i++) { tot += b.h_total.p[i]; if (m->streams[b.stream_of[i]].open) m->streams[b.stream_of[i]].decisions += b.h_total.p[i]; }
  m->last_step_total = tot;
  m->max_step_total = std::max (m->max_step_total, tot);
}

// Reads back what a launched batch reported (decision totals, capacity verdict); `only` = settle up to and including that
// batch (nullptr: all).  In steady state the batch finished long ago and this costs two small D2H copies on a side stream.
// If a batch skipped its coding because the decision list or the output arena was too small, every later batch skipped
// entirely (poison flag): grow, then run them again in launch order.
static int model_settle_batch (lh264_ctx* c, ModelCtx* m, MBatch* only) {
  if (only && !only->pending) return LH264_OK;
  while (!m->inflight.empty()) {
    MBatch& b = *m->inflight.front();
    CK (cudaEventSynchronize (b.done));
    CK (cudaMemcpyAsync (b.h_ctl.p, b.ctl.p, sizeof (BatchCtl), cudaMemcpyDeviceToHost, m->aux));
    CK (cudaMemcpyAsync (b.h_total.p, b.pic_total.p, b.n * sizeof (uint32_t), cudaMemcpyDeviceToHost, m->aux));
    CK (cudaStreamSynchronize (m->aux));
    if (b.h_ctl.p->skip) {
      CK (cudaStreamSynchronize (m->stream));       // the later batches have skipped themselves too by now
      std::vector<MBatch*> redo = m->inflight;
      m->n_redo++;
      if (m->stats) fprintf (stderr, "lh264 stats: arena %u pages; ", m->n_pages);
      if (m->stats) fprintf (stderr, "lh264 stats: a step did not fit (skip %d, total %llu M): %zu step(s) run again\n", (int)b.h_ctl.p->skip, (unsigned long long)b.h_ctl.p->total >> 20, redo.size());
      m->inflight.clear();
      bool head = b.h_ctl.p->skip == 2;             // 2 cannot be first in the list, but be safe
      CK (cudaMemsetAsync (m->d_poison, 0, sizeof (unsigned int), m->stream));
      {                                              // what the first one needs is known; later ones report when they run
        const BatchCtl ctl = *b.h_ctl.p;
        int rc = model_flush_coder (c, m);           // the steps generated before it are coded; the pool is empty again
        if (rc) return rc;
        if ((rc = model_make_room (c, m, ctl))) return rc;
      }
      for (MBatch* r : redo) {
        int rc = model_complete (c, m, *r, head);
        if (rc) return rc;
        r->pending = false;
        model_account (m, *r);
        head = true;
      }
      return LH264_OK;
    }
    b.pending = false;
    model_account (m, b);
    m->inflight.erase (m->inflight.begin());
    if (only == &b) return LH264_OK;
  }
  return LH264_OK;
}
static int model_settle (lh264_ctx* c, ModelCtx* m) { return model_settle_batch (c, m, nullptr); }

// The launches of one batch; nothing here waits for the device unless two batches are already in flight.
static int model_run (lh264_ctx* c, ModelCtx* m, MBatch& b) {
  const int n = b.n;
  if (n == 0) return LH264_OK;
  int rc = model_settle_batch (c, m, &b);          // a resident batch that is run again while still pending
  if (rc) return rc;
  // (LH264_MODEL_INFLIGHT: how many steps of RESIDENT batches the host may run ahead of the device before it settles the oldest;
  //  default 2.  An experiment hook for the schedule outlier of profiles/r2_summary.md section 15 - not measured yet)
  while (m->inflight.size() >= (b.resident ? m->max_inflight_resident : (size_t)2)) if ((rc = model_settle_batch (c, m, m->inflight.front()))) return rc;
  for (int i = 0; i < n; i++) if (!m->streams[b.stream_of[i]].open) return fail (LH264_E_INVALID, "batch names a closed model stream");
  if (!m->d_page_ptrs && (rc = model_grow_arena (c, m, (unsigned int)n * CT_COUNT + 64))) return rc;
  if (!m->dec[m->cur].p && (rc = model_grow_pool (c, m, std::max<unsigned long long> ((unsigned long long)b.total_recs * 96 + 1024, m->last_step_total)))) return rc;   // first guess; model_settle grows it when needed
  // An empty half that could not hold code_steps steps of the size seen so far is re-made larger (it was sized from an early,
  // smaller step): a window cut short by the pool costs a coder launch that is as long as a full one (the longest picture sets
  // its time), so short windows halve the coder's throughput.
  if (m->pend_steps == 0 && m->dec[m->cur].p && m->max_step_total > 0 && (rc = model_grow_pool (c, m, m->max_step_total))) return rc;
  // a step about as large as the last one would not fit behind the pending ones: code those first (saves a skip + re-run)
  if (m->pend_steps > 0 && m->est_pool_used + m->last_step_total + m->last_step_total / 2 + 64 > m->dec[m->cur].cap && (rc = model_flush_coder (c, m))) return rc;
  m->last_resident = b.resident;
  // resident steps: one in-order chain on the device - this step's generation behind the reconstruction kernels submitted so far,
  // the next reconstruction behind this generation (model_order_recon) - because side by side the persistent reconstruction kernel
  // and the generation kernels cost more than one after the other (measured: 33-53 ms per step together, 27.6 ms in a row)
  if (model_serial (m)) { CK (cudaEventRecord (m->recon_idle, c->stream)); CK (cudaStreamWaitEvent (m->stream, m->recon_idle, 0)); }
  if ((rc = model_launch (c, m, b, true, c->timing))) return rc;
  b.pending = true; m->inflight.push_back (&b);
  if (c->timing) {
    CK (cudaEventSynchronize (m->ev[1]));
    float t = 0;
    cudaEventElapsedTime (&t, m->ev[0], m->ev[1]); m->ms_gen += t;
    if ((rc = model_settle (c, m))) return rc;     // (a retry after growth is not part of the timed figures)
  }
  CK (cudaEventRecord (m->done, m->stream));
  m->done_recorded = true;
  if (b.shares_recon_batch >= 0) {                 // that staging batch may be overwritten once this run is over
    CK (cudaEventRecord (c->batch_free[b.shares_recon_batch], m->stream));
    c->batch_shared[b.shares_recon_batch] = true;
    c->batch_sharer[b.shares_recon_batch] = &b;
  }
  CK (cudaGetLastError());
  return LH264_OK;
}

static int model_submit_impl (lh264_ctx* c, const lh264_model_picture* pics, int n, const lh264_compact_picture* cp) {
  if (!c || (n > 0 && !pics) || n < 0) return fail (LH264_E_INVALID, "bad arguments");
  if (n == 0) return LH264_OK;
  CK (cudaSetDevice (c->device));
  ModelCtx* m = model_ctx (c);
  if (!m) return fail (LH264_E_NOMEM, "out of host memory");
  m->batch_cur ^= 1;
  MBatch& b = m->batch[m->batch_cur];
  int rc = model_stage (c, m, b, pics, n, true, cp);
  if (rc) return rc;
  return model_run (c, m, b);
}
int lh264_model_submit (lh264_ctx* c, const lh264_model_picture* pics, int n) { return model_submit_impl (c, pics, n, nullptr); }

int lh264_submit_compact (lh264_ctx* c, const lh264_compact_picture* pics, int n, int flags) {
  if (!c || (n > 0 && !pics) || n < 0) return fail (LH264_E_INVALID, "bad arguments");
  if (n == 0) return LH264_OK;
  int rc = LH264_OK;
  if (flags & LH264_COMPACT_RECON) {
    std::vector<lh264_picture_desc> d (n);
    for (int i = 0; i < n; i++) {
      d[i].stream = pics[i].stream; d[i].cur_slot = pics[i].cur_slot; d[i].n_slices = pics[i].n_slices; d[i].n_coef_blocks = pics[i].n_coef_blocks;
      d[i].slices = pics[i].slices; d[i].mbs = nullptr; d[i].coefs = pics[i].coefs;
    }
    if ((rc = recon_submit_impl (c, d.data(), n, pics))) return rc;
  }
  if (flags & LH264_COMPACT_MODEL) {
    std::vector<lh264_model_picture> d;
    std::vector<lh264_compact_picture> cp;
    for (int i = 0; i < n; i++) {
      if (pics[i].n_recs <= 0) continue;              // (a picture without model records: nothing to code)
      lh264_model_picture q;
      q.model_stream = pics[i].model_stream; q.mb_w = pics[i].mb_w; q.mb_h = pics[i].mb_h; q.n_slices = pics[i].n_model_slices; q.n_recs = pics[i].n_recs;
      q.n_coef_blocks = pics[i].n_coef_blocks; q.slices = pics[i].model_slices; q.recs = nullptr; q.coefs = pics[i].coefs;
      d.push_back (q); cp.push_back (pics[i]);
    }
    if (!d.empty() && (rc = model_submit_impl (c, d.data(), (int)d.size(), cp.data()))) return rc;
  }
  return rc;
}

int lh264_model_batch_upload (lh264_ctx* c, const lh264_model_picture* pics, int n, lh264_model_batch** out) {
  if (!c || !pics || n <= 0 || !out) return fail (LH264_E_INVALID, "bad arguments");
  CK (cudaSetDevice (c->device));
  ModelCtx* m = model_ctx (c);
  if (!m) return fail (LH264_E_NOMEM, "out of host memory");
  lh264_model_batch* rb = new (std::nothrow) lh264_model_batch();
  if (!rb) return fail (LH264_E_NOMEM, "out of host memory");
  int rc = model_stage (c, m, rb->b, pics, n, false);
  if (rc) { rb->b.release(); delete rb; return rc; }
  rb->b.resident = true;
  CK (cudaStreamSynchronize (m->stream));
  *out = rb;
  return LH264_OK;
}
int lh264_model_batch_run (lh264_ctx* c, lh264_model_batch* b) {
  if (!c || !b || !c->model) return fail (LH264_E_INVALID, "bad arguments");
  CK (cudaSetDevice (c->device));
  return model_run (c, static_cast<ModelCtx*> (c->model), b->b);
}
void lh264_model_batch_free (lh264_ctx* c, lh264_model_batch* b) {
  if (!b) return;
  if (c) { cudaSetDevice (c->device); model_sync (c); }
  b->b.release();
  delete b;
}

static int model_check_error (lh264_ctx* c, MStreamHost& s) {
  ModelCtx* m = static_cast<ModelCtx*> (c->model);
  { int rs = model_settle (c, m); if (rs) return rs; }
  { int rs = model_coder_idle (c, m); if (rs) return rs; }
  int err = 0;
  CK (cudaMemcpyAsync (&err, &s.d_state->error, sizeof (int), cudaMemcpyDeviceToHost, m->stream));
  CK (cudaStreamSynchronize (m->stream));
  if (err >= 11 && err <= 14) return fail (LH264_E_UNSUPPORTED, "a context index left the reference's prior tables (the reference asserts on this input; site %d)", err);
  if (err) { if (err != 2) return fail (LH264_E_CUDA, "model kernels: internal inconsistency %d", err); }
  if (err == 2) return fail (LH264_E_NOMEM, "model output pages exhausted");
  return LH264_OK;
}

int lh264_model_finish (lh264_ctx* c, int ms) {
  ModelCtx* m;
  int rc = model_stream_ok (c, ms, &m);
  if (rc) return rc;
  CK (cudaSetDevice (c->device));
  if ((rc = model_settle (c, m))) return rc;
  if ((rc = model_coder_idle (c, m))) return rc;
  unsigned int used = 0;
  if (m->d_next) { CK (cudaMemcpyAsync (&used, m->d_next, sizeof (used), cudaMemcpyDeviceToHost, m->stream)); CK (cudaStreamSynchronize (m->stream)); }
  if ((rc = model_grow_arena (c, m, used + CT_COUNT))) return rc;
  ArenaDev ar{ m->d_page_ptrs, m->d_next, m->n_pages };
  k_model_finish<<<1, 32, 0, m->stream>>> (m->streams[ms].d_state, ar);
  c->launches++;
  CK (cudaGetLastError());
  return model_check_error (c, m->streams[ms]);
}

long long lh264_model_read (lh264_ctx* c, int ms, int tag, uint8_t* dst, size_t offset, size_t cap) {
  ModelCtx* m;
  int rc = model_stream_ok (c, ms, &m);
  if (rc) return rc;
  int ct = -1;
  for (int i = 0; i < CT_COUNT; i++) if (kTagOfCompact[i] == tag) ct = i;
  if (ct < 0) return 0;                              // a tag the model never writes: no side stream
  CK (cudaSetDevice (c->device));
  MStreamHost& s = m->streams[ms];
  if ((rc = model_check_error (c, s))) return rc;
  struct { uint32_t used; uint32_t finished; } uf;
  Coder cd;
  CK (cudaMemcpy (&uf, &s.d_state->used, sizeof (uf), cudaMemcpyDeviceToHost));
  CK (cudaMemcpy (&cd, &s.d_state->coder[ct], sizeof (cd), cudaMemcpyDeviceToHost));
  if (!((uf.used >> ct) & 1u)) return 0;
  const size_t total = cd.pos;
  if (!dst || cap == 0 || offset >= total) return (long long)total;
  const size_t end = std::min (total, offset + cap);
  const size_t p0 = offset >> PAGE_SHIFT, p1 = (end - 1) >> PAGE_SHIFT;
  std::vector<uint32_t> pl (p1 - p0 + 1);
  CK (cudaMemcpy (pl.data(), s.d_pagelist + (size_t)ct * MAX_PAGES_PER_TAG + p0, pl.size() * sizeof (uint32_t), cudaMemcpyDeviceToHost));
  size_t at = offset;
  while (at < end) {
    const size_t pg = at >> PAGE_SHIFT, in = at & (PAGE_SIZE - 1), nb = std::min ((size_t)PAGE_SIZE - in, end - at);
    const uint32_t idx = pl[pg - p0];
    if (idx == 0 || idx > m->h_page_ptrs.size()) return fail (LH264_E_CUDA, "model output page list is inconsistent");
    CK (cudaMemcpy (dst + (at - offset), m->h_page_ptrs[idx - 1] + in, nb, cudaMemcpyDeviceToHost));
    at += nb;
  }
  return (long long)total;
}

unsigned long long lh264_model_decisions (lh264_ctx* c, int ms) {
  ModelCtx* m;
  if (model_stream_ok (c, ms, &m)) return 0;
  if (model_settle (c, m)) return 0;
  return m->streams[ms].decisions;
}

} // extern "C"

static cudaError_t model_sync (lh264_ctx* c) {
  ModelCtx* m = static_cast<ModelCtx*> (c->model);
  if (m && !m->inflight.empty() && model_settle (c, m) != LH264_OK) return cudaErrorUnknown;
  if (m && model_coder_idle (c, m) != LH264_OK) return cudaErrorUnknown;
  return m && m->stream ? cudaStreamSynchronize (m->stream) : cudaSuccess;
}

// the context stream waits for everything submitted to the model stream so far
static cudaError_t model_join (lh264_ctx* c) {
  ModelCtx* m = static_cast<ModelCtx*> (c->model);
  if (!m || !m->stream) return cudaSuccess;
  if (model_flush_coder (c, m) != LH264_OK) return cudaErrorUnknown;
  m->beside_coder = false;             // what follows a join is ordered behind it: the next window is chained, not run beside the launch in front of the join
  cudaError_t e = cudaEventRecord (m->done, m->stream);
  if (e == cudaSuccess) e = cudaStreamWaitEvent (c->stream, m->done, 0);
  for (int k = 0; k < 2 && e == cudaSuccess; k++) if (m->coder_pending[k]) e = cudaStreamWaitEvent (c->stream, m->coder_done[k], 0);
  return e;
}

static int model_settle_sharer (lh264_ctx* c, int k) {
  ModelCtx* m = static_cast<ModelCtx*> (c->model);
  MBatch* b = static_cast<MBatch*> (c->batch_sharer[k]);
  c->batch_sharer[k] = nullptr;
  if (!m || !b) return LH264_OK;
  for (MBatch* f : m->inflight) if (f == b) {
      const int rc = model_settle_batch (c, m, b);
      if (rc) return rc;
      // a re-run recorded nothing on batch_free: order the context stream after the model stream as it is now
      CK (cudaEventRecord (c->batch_free[k], m->stream));
      return LH264_OK;
    }
  return LH264_OK;
}

// the caller's host buffers of every model submit so far have been copied to the device
// called in front of a reconstruction launch: in the in-order mode it waits for the model work submitted so far
static cudaError_t model_order_recon (lh264_ctx* c) {
  ModelCtx* m = static_cast<ModelCtx*> (c->model);
  if (!m || !model_serial (m) || !m->done_recorded) return cudaSuccess;
  return cudaStreamWaitEvent (c->stream, m->done, 0);
}

static cudaError_t model_inputs_consumed (lh264_ctx* c) {
  ModelCtx* m = static_cast<ModelCtx*> (c->model);
  if (!m) return cudaSuccess;
  for (int k = 0; k < 2; k++) if (m->batch[k].staged) { cudaError_t e = cudaEventSynchronize (m->batch[k].staged); if (e != cudaSuccess) return e; }
  return cudaSuccess;
}
