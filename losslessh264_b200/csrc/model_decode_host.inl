// model_decode_host.inl — host side of lh264_model_decode(): lays the jobs out in one device slab, uploads the side streams and
// slice parameters through one pinned staging buffer, launches k_model_decode (model_decode.cuh) and brings the bits back.
// Included by lh264_shim.cu (uses its lh264_ctx, DevBuf / PinBuf, CK and fail).
namespace lh264md {

struct DecodeCtx {                     // grow-only buffers of lh264_model_decode, kept in the context
  DevBuf<uint8_t> slab;
  PinBuf<uint8_t> h_in;
  cudaEvent_t ev[2] = {nullptr, nullptr};
  float ms = 0.f;
  ~DecodeCtx() { slab.release(); h_in.release(); for (auto& e : ev) if (e) cudaEventDestroy (e); }
};


} // namespace lh264md

static void decode_destroy (lh264_ctx* c) { delete static_cast<lh264md::DecodeCtx*> (c->decode); c->decode = nullptr; }
static float decode_ms (lh264_ctx* c, bool reset) {
  lh264md::DecodeCtx* d = static_cast<lh264md::DecodeCtx*> (c->decode);
  if (!d) return 0.f;
  if (reset) d->ms = 0.f;
  return d->ms;
}

extern "C" int lh264_model_decode (lh264_ctx* c, lh264_decode_job* jobs, int n_jobs) {
  using namespace lh264md;
  if (!c || !jobs || n_jobs <= 0) return fail (LH264_E_INVALID, "bad arguments");
  CK (cudaSetDevice (c->device));
  if (!c->decode) c->decode = new (std::nothrow) DecodeCtx();
  DecodeCtx* d = static_cast<DecodeCtx*> (c->decode);
  if (!d) return fail (LH264_E_NOMEM, "out of host memory");
  if (!d->ev[0]) { CK (cudaEventCreate (&d->ev[0])); CK (cudaEventCreate (&d->ev[1])); }
  // ---- layout: [MdJobDev x n] [inputs: slices + side streams, staged through one pinned buffer] [state + outputs] ----
  auto up = [] (size_t v) { return (v + 255) & ~(size_t)255; };
  struct Lay { size_t slices, tags[CT_COUNT], priors, frame[2], ipm, mb_slice, intra, tcoef, sout, out; };
  std::vector<Lay> lay (n_jobs);
  size_t in_off = up (sizeof (MdJobDev) * (size_t)n_jobs);
  const size_t in_begin = 0;
  for (int j = 0; j < n_jobs; j++) {
    const lh264_decode_job& job = jobs[j];
    if (job.max_mbs <= 0 || job.n_slices < 0 || (job.n_slices && (!job.slices || !job.slice_out)) || !job.out) return fail (LH264_E_INVALID, "job %d: bad fields", j);
    lay[j].slices = in_off; in_off = up (in_off + sizeof (lh264_decode_slice) * (size_t)job.n_slices);
    for (int t = 0; t < CT_COUNT; t++) {
      const int tag = lh264m::kTagOfCompact[t];
      lay[j].tags[t] = in_off;
      if (job.tag_data[tag]) in_off = up (in_off + job.tag_len[tag] + 8);
    }
  }
  const size_t in_end = in_off;
  size_t off = in_end;
  for (int j = 0; j < n_jobs; j++) {
    const lh264_decode_job& job = jobs[j];
    const size_t n = (size_t)job.max_mbs;
    lay[j].priors = off; off = up (off + 4 * (size_t) (job.with_8x8 ? lh264m::P_END : lh264m::P_END_NO8));
    for (int k = 0; k < 2; k++) { lay[j].frame[k] = off; off = up (off + n * sizeof (lh264m::CtxRec)); }
    lay[j].ipm = off; off = up (off + n * 16);
    lay[j].mb_slice = off; off = up (off + n * 4);
    lay[j].intra = off; off = up (off + n);
    lay[j].tcoef = off; off = up (off + n * MD_NB_STRIDE);
    lay[j].sout = off; off = up (off + sizeof (lh264_decode_slice_out) * (size_t)job.n_slices);
    lay[j].out = off; off = up (off + job.out_capacity);
  }
  CK (d->slab.reserve (off));
  CK (d->h_in.reserve (in_end));
  uint8_t* base = d->slab.p;
  uint8_t* h = d->h_in.p;
  memset (h, 0, up (sizeof (MdJobDev) * (size_t)n_jobs));
  for (int j = 0; j < n_jobs; j++) {
    const lh264_decode_job& job = jobs[j];
    MdJobDev& jd = reinterpret_cast<MdJobDev*> (h)[j];
    jd.st.priors = reinterpret_cast<uint32_t*> (base + lay[j].priors);
    jd.st.frame[0] = reinterpret_cast<lh264m::CtxRec*> (base + lay[j].frame[0]);
    jd.st.frame[1] = reinterpret_cast<lh264m::CtxRec*> (base + lay[j].frame[1]);
    jd.st.ipm = base + lay[j].ipm; jd.st.mb_slice = reinterpret_cast<int32_t*> (base + lay[j].mb_slice);
    jd.st.mb_intra = base + lay[j].intra; jd.st.tcoef = base + lay[j].tcoef;
    jd.st.max_mbs = job.max_mbs;
    jd.slices = reinterpret_cast<const MdSliceIn*> (base + lay[j].slices);
    jd.out = base + lay[j].out; jd.cap = job.out_capacity;
    jd.sout = reinterpret_cast<lh264_decode_slice_out*> (base + lay[j].sout);
    jd.n_slices = job.n_slices;
    if (job.n_slices) memcpy (h + lay[j].slices, job.slices, sizeof (lh264_decode_slice) * (size_t)job.n_slices);
    for (int t = 0; t < CT_COUNT; t++) {
      const int tag = lh264m::kTagOfCompact[t];
      if (!job.tag_data[tag]) continue;
      jd.tag_buf[t] = base + lay[j].tags[t]; jd.tag_len[t] = job.tag_len[tag];
      memcpy (h + lay[j].tags[t], job.tag_data[tag], job.tag_len[tag]);
    }
  }
  cudaStream_t s = c->stream;
  CK (cudaMemcpyAsync (base + in_begin, h, in_end, cudaMemcpyHostToDevice, s));
  for (int j = 0; j < n_jobs; j++) CK (cudaMemsetAsync (base + lay[j].priors, 0, lay[j].frame[0] - lay[j].priors, s));
  CK (cudaEventRecord (d->ev[0], s));
  k_model_decode<<<MD_GRID (n_jobs), MD_BLOCK, 0, s>>> (reinterpret_cast<MdJobDev*> (base), n_jobs);
  c->launches++;
  CK (cudaGetLastError());
  CK (cudaEventRecord (d->ev[1], s));
  // results: job headers (error), per-slice records, then exactly the bytes written
  CK (cudaMemcpyAsync (h, base, sizeof (MdJobDev) * (size_t)n_jobs, cudaMemcpyDeviceToHost, s));
  for (int j = 0; j < n_jobs; j++)
    if (jobs[j].n_slices) CK (cudaMemcpyAsync (jobs[j].slice_out, base + lay[j].sout, sizeof (lh264_decode_slice_out) * (size_t)jobs[j].n_slices, cudaMemcpyDeviceToHost, s));
  CK (cudaStreamSynchronize (s));
  float ms = 0.f;
  if (cudaEventElapsedTime (&ms, d->ev[0], d->ev[1]) == cudaSuccess) d->ms += ms;
  int bad = 0, first_err = 0;
  for (int j = 0; j < n_jobs; j++) {
    lh264_decode_job& job = jobs[j];
    job.error = reinterpret_cast<MdJobDev*> (h)[j].error;
    if (job.error) { if (!bad) first_err = job.error; bad++; continue; }
    if (!job.n_slices) continue;
    const lh264_decode_slice_out& last = job.slice_out[job.n_slices - 1];
    const size_t used = (size_t)last.byte_offset + (last.n_bits + 7) / 8 + 1;
    CK (cudaMemcpyAsync (job.out, base + lay[j].out, std::min<size_t> (used, job.out_capacity), cudaMemcpyDeviceToHost, s));
  }
  CK (cudaStreamSynchronize (s));
  if (bad) return fail (LH264_E_INVALID, "%d of %d decode jobs failed (first error code %d: 21 I_PCM, 22 output buffer too small, 23 corrupt side streams, 24 picture larger than max_mbs, 25 context outside the prior tables)", bad, n_jobs, first_err);
  return LH264_OK;
}
