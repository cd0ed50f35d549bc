// main_batch.cpp — lh264_batch: command-line front end of the multi-stream runtime (include/lh264_runtime.h, runtime.cpp).
//
// usage: lh264_batch [--copies N] [--threads T] [--groups G] [--loops L] [--device D] [--depth K] [--no-yuv] [--no-hash] [--no-model]
//                    [--pip-dir DIR] [--max-pictures P] in.264...
// prints one line per distinct input (pictures, YUV SHA1, SHA1 over all .pip bytes) and a JSON summary line.
// --no-yuv --> model only: `h264dec in.264 out.pip` without a YUV sink (no reconstruction, no picture comes back).
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <string>
#include <vector>
#include "../../include/lh264_runtime.h"
#include "../../include/lh264_b200.h"

int main (int argc, char** argv) {
  lh264_rt_config cfg;
  memset (&cfg, 0, sizeof (cfg));
  cfg.want_yuv = cfg.want_hash = cfg.want_pip = 1;
  cfg.pipelines = 1; cfg.loops = 1;
  int copies = 1;
  std::string pip_dir;
  std::vector<std::string> files;
  for (int i = 1; i < argc; i++) {
    const std::string a = argv[i];
    if (a == "--copies" && i + 1 < argc) copies = atoi (argv[++i]);
    else if (a == "--threads" && i + 1 < argc) cfg.threads = atoi (argv[++i]);
    else if (a == "--groups" && i + 1 < argc) cfg.pipelines = atoi (argv[++i]);
    else if (a == "--loops" && i + 1 < argc) cfg.loops = atoi (argv[++i]);
    else if (a == "--device" && i + 1 < argc) cfg.device = atoi (argv[++i]);
    else if (a == "--depth" && i + 1 < argc) cfg.depth = atoi (argv[++i]);
    else if (a == "--max-pictures" && i + 1 < argc) cfg.max_pictures = atoi (argv[++i]);
    else if (a == "--pip-dir" && i + 1 < argc) pip_dir = argv[++i];
    else if (a == "--no-yuv") cfg.want_yuv = 0;
    else if (a == "--no-hash") cfg.want_hash = 0;      // still brings every decoded picture to the host, but does not SHA-1 it
    else if (a == "--no-model") cfg.want_pip = 0;
    else files.push_back (a);
  }
  if (files.empty() || copies < 1) {
    fprintf (stderr, "usage: %s [--copies N] [--threads T] [--groups G] [--loops L] [--device D] [--depth K] [--no-yuv] [--no-hash] [--no-model] [--pip-dir DIR] [--max-pictures P] in.264...\n", argv[0]);
    return 2;
  }
  const int nf = (int)files.size(), S = copies * nf;
  std::vector<std::vector<uint8_t> > inputs (nf);
  for (int f = 0; f < nf; f++) {
    FILE* fi = fopen (files[f].c_str(), "rb");
    if (!fi) { perror (files[f].c_str()); return 1; }
    uint8_t tmp[65536]; size_t n;
    while ((n = fread (tmp, 1, sizeof (tmp), fi)) > 0) inputs[f].insert (inputs[f].end(), tmp, tmp + n);
    fclose (fi);
    static const uint8_t kStart[4] = {0, 0, 0, 1};
    inputs[f].insert (inputs[f].end(), kStart, kStart + 4);     // sentinel, h264dec.cpp:238
  }
  lh264_rt* rt = nullptr;
  if (lh264_rt_create (&rt, &cfg) != 0) { fprintf (stderr, "lh264_rt_create failed\n"); return 1; }
  for (int i = 0; i < S; i++) if (lh264_rt_add_stream (rt, inputs[i % nf].data(), inputs[i % nf].size() - 4, 1) < 0) { fprintf (stderr, "lh264_rt_add_stream failed\n"); return 1; }
  const int rc = lh264_rt_run (rt, nullptr, nullptr);
  if (rc != 0) { fprintf (stderr, "lh264_rt_run: %s\n", lh264_rt_last_error (rt)); lh264_rt_destroy (rt); return 1; }
  std::vector<lh264_rt_stream_result> res (S);
  for (int i = 0; i < S; i++) lh264_rt_stream_result_get (rt, i, &res[i]);
  if (!pip_dir.empty() && cfg.want_pip) {
    std::vector<uint8_t> b;
    for (int i = 0; i < S; i++) for (int tag = 0; tag < LH264_NUM_TAGS; tag++) {
      const long long n = lh264_rt_pip_read (rt, i, tag, nullptr, 0);
      if (n <= 0 && tag != 0) continue;
      b.resize ((size_t)std::max (0LL, n));
      if (n > 0) lh264_rt_pip_read (rt, i, tag, b.data(), b.size());
      const std::string p = pip_dir + "/s" + std::to_string (i) + ".pip" + (tag ? "." + std::to_string (tag) : std::string());
      FILE* f = fopen (p.c_str(), "wb");
      if (f) { fwrite (b.data(), 1, b.size(), f); fclose (f); }
    }
  }
  // every copy of an input must give identical results; report one line per distinct input
  int mismatches = 0;
  for (int f = 0; f < nf; f++) {
    for (int i = f + nf; i < S; i += nf)
      if (strcmp (res[i].yuv_sha1, res[f].yuv_sha1) || strcmp (res[i].pip_sha1, res[f].pip_sha1) || res[i].pictures != res[f].pictures) mismatches++;
    printf ("%s pictures %ld yuv_sha1 %s pip_sha1 %s\n", files[f].c_str(), res[f].pictures, res[f].yuv_sha1, cfg.want_pip ? res[f].pip_sha1 : "-");
  }
  lh264_rt_stats st;
  lh264_rt_stats_get (rt, &st);
  printf ("{\"streams\": %ld, \"threads\": %d, \"groups\": %d, \"loops\": %d, \"pictures\": %ld, \"seconds\": %.4f, \"frames_per_s\": %.1f, \"input_mb_per_s\": %.2f, \"pip_bytes\": %zu, "
          "\"seconds_incl_finish\": %.4f, \"submit_wait_parse_s\": %.4f, \"submit_wait_gpu_s\": %.4f, \"parse_busy_s\": %.4f, \"parse_wait_s\": %.4f, \"h2d_bytes\": %zu, \"d2h_bytes\": %zu, \"setup_s\": %.4f, \"pinned_alloc_s\": %.4f, \"open_streams_s\": %.4f, \"steady_frames_per_s\": %.1f, "
          "\"copy_mismatches\": %d, \"kernel_launches\": %llu}\n",
          st.streams, st.threads, st.pipelines, cfg.loops, st.pictures, st.seconds, st.pictures / st.seconds, st.input_bytes / st.seconds / 1e6, st.pip_bytes, st.seconds_total,
          st.submit_wait_parse_s, st.submit_wait_gpu_s, st.parse_busy_s, st.parse_wait_s, st.h2d_bytes, st.d2h_bytes, st.setup_s, st.pinned_alloc_s, st.open_streams_s,
          st.seconds > st.setup_s ? (st.pictures - st.setup_pictures) / (st.seconds - st.setup_s) : 0.0, mismatches, st.kernel_launches);
  lh264_rt_destroy (rt);
  return mismatches ? 3 : 0;
}
