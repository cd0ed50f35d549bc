/*
 * oracle/model_decode_cpu.cpp — TEST INFRASTRUCTURE ONLY.  Never linked, imported or executed by the product; only tests/
 * may build and run it, as the checker of the decompress direction.
 *
 * CPU build of losslessh264_b200/csrc/model_decode_core.inl (the code the CUDA kernel k_model_decode runs, compiled here
 * with MD_FN = static inline) behind the signature of lh264_model_decode(), plus a small CLI:
 *     expand_cpu x.pip out.264 [--intra-as-cabac]
 * which reads x.pip (default stream) and x.pip.<tag>, expands them with host/pip_expand.cpp and writes the .264.
 * Parity is PINNED by the data itself: the expanded file must equal the original .264 byte for byte, and the side streams
 * fed in are those of the compiled, unmodified reference (oracle/_ref/h264dec) or of this repository's compress direction,
 * which tests/golden/pip_tag_sha1.json pins to the reference's.
 * Follows: inc/compression_stream.h:601-689 (scan*), inc/bitreader.h:77-136, src/decode_slice.cpp:2476-2956, :1442-1835.
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <string>
#include <vector>
#include "../include/lh264_b200.h"
#include "../losslessh264_b200/csrc/model_layout.h"
#include "../losslessh264_b200/host/cavlc_tables.h"
#include "../losslessh264_b200/host/cabac_tables.h"
#define MD_FN static inline
#define MD_FN_NOINLINE static
#define MD_TABLE static const
#define MD_CLZ(x) __builtin_clz (x)
#define MD_DIV(n, d) ((n) / (d))
#define MD_READER(c, tag) ((c).rd[tag])
#define MD_TEST(c) ((c).test)
#include "../losslessh264_b200/csrc/model_decode_core.inl"
#include "../losslessh264_b200/host/pip_expand.h"

using namespace lh264md;

static_assert (sizeof (MdSliceIn) == sizeof (lh264_decode_slice), "slice descriptor layout");

extern "C" int mdc_model_decode (lh264_decode_job* jobs, int n_jobs) {
  int bad = 0;
  for (int j = 0; j < n_jobs; j++) {
    lh264_decode_job& job = jobs[j];
    if (job.max_mbs <= 0 || !job.out) { job.error = MD_ERR_DIMS; bad++; continue; }   // as lh264_model_decode (model_decode_host.inl)
    const size_t n = (size_t)job.max_mbs;
    std::vector<uint32_t> priors (job.with_8x8 ? P_END : P_END_NO8, 0);
    std::vector<CtxRec> f0 (n), f1 (n);
    memset (f0.data(), 0, n * sizeof (CtxRec)); memset (f1.data(), 0, n * sizeof (CtxRec));
    std::vector<uint8_t> ipm (n * 16, 2), intra (n, 0), tcoef (n * 48, 0);
    std::vector<int32_t> mb_slice (n, -1);
    MdStream st;
    memset (&st, 0, sizeof (st));
    st.priors = priors.data(); st.frame[0] = f0.data(); st.frame[1] = f1.data();
    st.ipm = ipm.data(); st.mb_slice = mb_slice.data(); st.mb_intra = intra.data(); st.tcoef = tcoef.data();
    st.max_mbs = job.max_mbs;
    for (int t = 0; t < CT_COUNT; t++) {
      const int tag = kTagOfCompact[t];
      md_reader_init (st.rd[t], job.tag_data[tag], job.tag_data[tag] ? job.tag_len[tag] : 0);
      if (job.tag_data[tag]) st.present |= 1u << t;
    }
    MdScratch sc;
    md_scratch_init (sc, 0, 1);
    MdCtx c;
    md_ctx_init (c, st.priors, st.rd);
    c.test = priors[P_TEST];
    uint64_t off = 0;
    for (int s = 0; s < job.n_slices && !st.error; s++) {
      MdSliceIn sl;
      memcpy (&sl, &job.slices[s], sizeof (sl));
      CtxRec* f = md_slice_flip (st, sl.frame_num);
      md_update_frame_serial (f, st.frame_len);
      const bool prior_valid = md_slice_resize (st, sl.mb_w, sl.mb_h);
      if (st.error) break;
      MdBits bw;
      mb_init (bw, job.out + off, job.out_capacity - off);
      lh264_decode_slice_out& so = job.slice_out[s];
      memset (&so, 0, sizeof (so));
      so.n_mbs = (uint32_t)md_slice (st, c, sl, prior_valid, bw, sc, &so.pad_bits, &so.n_pad_bits);
      mb_flush (bw);
      so.byte_offset = off; so.n_bits = (uint32_t) (mb_bits (bw) - (uint64_t)bw.trim);
      off += (mb_bits (bw) + 7) / 8;
    }
    job.error = st.error;
    if (st.error) bad++;
  }
  return bad ? -1 : 0;
}

static bool slurp (const std::string& path, std::vector<uint8_t>& out) {
  FILE* f = fopen (path.c_str(), "rb");
  if (!f) return false;
  uint8_t tmp[65536]; size_t n;
  out.clear();
  while ((n = fread (tmp, 1, sizeof (tmp), f)) > 0) out.insert (out.end(), tmp, tmp + n);
  fclose (f);
  return true;
}

int main (int argc, char** argv) {
  if (argc < 3) { fprintf (stderr, "usage: %s x.pip out.264 [--intra-as-cabac]\n", argv[0]); return 2; }
  std::vector<uint8_t> def;
  if (!slurp (argv[1], def)) { perror (argv[1]); return 1; }
  std::vector<std::vector<uint8_t>> tags (LH264_NUM_TAGS);
  lh264_decode_job job;
  memset (&job, 0, sizeof (job));
  size_t total = def.size();
  for (int t = 0; t < LH264_NUM_TAGS; t++)
    if (slurp (std::string (argv[1]) + "." + std::to_string (t), tags[t])) {
      tags[t].push_back (0);                           // keeps data() non-null for an empty file
      job.tag_data[t] = tags[t].data(); job.tag_len[t] = (uint32_t)tags[t].size() - 1; total += tags[t].size();
    }
  lh264host::ExpandPlan plan;
  if (!lh264host::expand_plan (def.data(), def.size(), plan)) { fprintf (stderr, "cannot expand: %s\n", plan.error.c_str()); return 3; }
  // test hook (tests/test_cabac_transcode.py): write the slice data of I slices with the CABAC writer instead of the CAVLC one.
  // The header of an I slice is the same for both entropy coders, so flipping entropy_coding_mode_flag in the PPS of the result
  // gives a valid CABAC stream - a way to exercise CABAC paths (8x8 residual blocks) no stream of the corpus has.
  if (argc > 3 && !strcmp (argv[3], "--intra-as-cabac"))
    for (lh264_decode_slice& sl : plan.slices) if (sl.slice_type == 2) sl.flags |= 4;
  std::vector<lh264_decode_slice_out> outs (plan.slices.size() + 1);
  std::vector<uint8_t> bits;
  for (size_t cap = total * 4 + (1u << 20);; cap *= 2) {
    bits.assign (cap, 0);
    job.max_mbs = plan.max_mbs; job.with_8x8 = plan.with_8x8;
    job.slices = plan.slices.data(); job.n_slices = (int)plan.slices.size();
    job.out = bits.data(); job.out_capacity = bits.size(); job.slice_out = outs.data();
    mdc_model_decode (&job, 1);
    if (job.error != MD_ERR_OUT_FULL) break;
  }
  if (job.error) { fprintf (stderr, "model decode failed: error %d\n", job.error); return 4; }
  std::vector<uint8_t> out;
  if (!lh264host::expand_assemble (plan, def.data(), bits.data(), outs.data(), out)) { fprintf (stderr, "assembly failed\n"); return 5; }
  FILE* fo = fopen (argv[2], "wb");
  if (!fo) { perror (argv[2]); return 1; }
  fwrite (out.data(), 1, out.size(), fo);
  fclose (fo);
  fprintf (stderr, "%zu slices, %zu bytes\n", plan.slices.size(), out.size());
  return 0;
}
