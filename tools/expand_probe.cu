// tools/expand_probe.cu — development harness for k_model_decode: compiles csrc/model_decode.cuh alone (seconds instead of the
// minutes of the whole library) so that kernel variants (-DMD_...) can be timed against each other in one GPU session.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o /tmp/expand_probe tools/expand_probe.cu \
//        losslessh264_b200/host/pip_expand.cpp losslessh264_b200/host/pip_container.cpp losslessh264_b200/host/h264_syntax.cpp
//   /tmp/expand_probe x.pip original.264 <streams> [repeats]
// Not part of the product or the tests.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <string>
#include <vector>
#include "../include/lh264_b200.h"
#include "../losslessh264_b200/csrc/model_layout.h"
#include "../losslessh264_b200/csrc/model_decode.cuh"
#include "../losslessh264_b200/host/pip_expand.h"

using namespace lh264md;
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf (stderr, "%s: %s\n", #x, cudaGetErrorString (e_)); exit (1); } } while (0)

static bool slurp (const std::string& path, std::vector<uint8_t>& out) {
  FILE* f = fopen (path.c_str(), "rb");
  if (!f) return false;
  uint8_t tmp[65536]; size_t n; out.clear();
  while ((n = fread (tmp, 1, sizeof (tmp), f)) > 0) out.insert (out.end(), tmp, tmp + n);
  fclose (f);
  return true;
}

int main (int argc, char** argv) {
  if (argc < 4) { fprintf (stderr, "usage: %s x.pip original.264 streams [repeats [max slices]]\n", argv[0]); return 2; }
  const int S = atoi (argv[3]), reps = argc > 4 ? atoi (argv[4]) : 2;
  std::vector<uint8_t> def, orig;
  if (!slurp (argv[1], def) || !slurp (argv[2], orig)) { perror ("input"); return 1; }
  std::vector<std::vector<uint8_t>> tags (LH264_NUM_TAGS);
  size_t total = def.size();
  for (int t = 0; t < LH264_NUM_TAGS; t++) if (slurp (std::string (argv[1]) + "." + std::to_string (t), tags[t])) total += tags[t].size();
  lh264host::ExpandPlan plan;
  if (!lh264host::expand_plan (def.data(), def.size(), plan)) { fprintf (stderr, "plan: %s\n", plan.error.c_str()); return 3; }
  const int max_slices = argc > 5 ? atoi (argv[5]) : 0;      // profile runs: only the first few slices (no verification then)
  if (max_slices > 0 && (size_t)max_slices < plan.slices.size()) plan.slices.resize (max_slices);
  const size_t n = (size_t)plan.max_mbs, ns = plan.slices.size();
  const size_t cap = total * 4 + (1u << 20);
  auto up = [] (size_t v) { return (v + 255) & ~(size_t)255; };
  // shared inputs (one copy), per-stream state
  size_t off = 0;
  const size_t o_slices = off; off = up (off + ns * sizeof (lh264_decode_slice));
  size_t o_tag[CT_COUNT];
  for (int t = 0; t < CT_COUNT; t++) { o_tag[t] = off; off = up (off + tags[kTagOfCompact[t]].size() + 8); }
  const size_t o_jobs = off; off = up (off + sizeof (MdJobDev) * (size_t)S);
  const size_t per_priors = up (4 * (size_t) (plan.with_8x8 ? P_END : P_END_NO8));
  const size_t per = per_priors + 2 * up (n * 32) + up (n * 16) + up (n * 4) + up (n) + up (n * 24) + up (ns * sizeof (lh264_decode_slice_out)) + up (cap);
  const size_t o_state = off; off += per * (size_t)S;
  uint8_t* base;
  CK (cudaMalloc (&base, off));
  std::vector<uint8_t> h (o_state, 0);
  memcpy (h.data() + o_slices, plan.slices.data(), ns * sizeof (lh264_decode_slice));
  for (int t = 0; t < CT_COUNT; t++) if (!tags[kTagOfCompact[t]].empty()) memcpy (h.data() + o_tag[t], tags[kTagOfCompact[t]].data(), tags[kTagOfCompact[t]].size());
  for (int j = 0; j < S; j++) {
    MdJobDev& jd = reinterpret_cast<MdJobDev*> (h.data() + o_jobs)[j];
    uint8_t* p = base + o_state + per * (size_t)j;
    jd.st.priors = (uint32_t*)p; p += per_priors;
    jd.st.frame[0] = (CtxRec*)p; p += up (n * 32); jd.st.frame[1] = (CtxRec*)p; p += up (n * 32);
    jd.st.ipm = p; p += up (n * 16); jd.st.mb_slice = (int32_t*)p; p += up (n * 4); jd.st.mb_intra = p; p += up (n); jd.st.tcoef = p; p += up (n * MD_NB_STRIDE);
    jd.sout = (lh264_decode_slice_out*)p; p += up (ns * sizeof (lh264_decode_slice_out));
    jd.out = p; jd.cap = cap;
    jd.st.max_mbs = plan.max_mbs; jd.slices = (const MdSliceIn*) (base + o_slices); jd.n_slices = (int)ns;
    for (int t = 0; t < CT_COUNT; t++) if (!tags[kTagOfCompact[t]].empty() || slurp (std::string (argv[1]) + "." + std::to_string (kTagOfCompact[t]), tags[kTagOfCompact[t]])) { jd.tag_buf[t] = base + o_tag[t]; jd.tag_len[t] = (uint32_t)tags[kTagOfCompact[t]].size(); }
  }
  cudaEvent_t e0, e1;
  CK (cudaEventCreate (&e0)); CK (cudaEventCreate (&e1));
  float best = 1e30f;
  for (int r = 0; r < reps; r++) {
    CK (cudaMemcpy (base, h.data(), o_state, cudaMemcpyHostToDevice));
    for (int j = 0; j < S; j++) CK (cudaMemsetAsync (base + o_state + per * (size_t)j, 0, per_priors, 0));
    CK (cudaEventRecord (e0));
    k_model_decode<<<MD_GRID (S), MD_BLOCK>>> ((MdJobDev*) (base + o_jobs), S);
    CK (cudaEventRecord (e1));
    CK (cudaDeviceSynchronize());
    float ms; CK (cudaEventElapsedTime (&ms, e0, e1));
    if (ms < best) best = ms;
  }
  // verify the last stream
  std::vector<MdJobDev> jobs (S);
  CK (cudaMemcpy (jobs.data(), base + o_jobs, sizeof (MdJobDev) * (size_t)S, cudaMemcpyDeviceToHost));
  int bad = 0;
  for (int j = 0; j < S; j++) if (jobs[j].error) bad++;
  std::vector<lh264_decode_slice_out> outs (ns + 1);
  std::vector<uint8_t> bits (cap);
  CK (cudaMemcpy (outs.data(), jobs[S - 1].sout, ns * sizeof (lh264_decode_slice_out), cudaMemcpyDeviceToHost));
  CK (cudaMemcpy (bits.data(), jobs[S - 1].out, cap, cudaMemcpyDeviceToHost));
  std::vector<uint8_t> out;
  if (max_slices > 0) { printf ("%s S=%d slices=%zu: k_model_decode %.1f ms (%d jobs with errors)\n", MD_VARIANT_NAME, S, ns, best, bad); return bad ? 1 : 0; }
  const bool ok = !bad && lh264host::expand_assemble (plan, def.data(), bits.data(), outs.data(), out) && out == orig;
  printf ("%s S=%d slices=%zu: k_model_decode %.1f ms, %s (%d jobs with errors)\n", MD_VARIANT_NAME, S, ns, best, ok ? "IDENTICAL" : "DIFFERENT", bad);
  return ok ? 0 : 1;
}
