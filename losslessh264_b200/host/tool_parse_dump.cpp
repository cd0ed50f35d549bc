// tool_parse_dump.cpp — CPU-only tool: parses an Annex-B stream with the host decoder and writes
// the hand-off records as .lhfb (no reconstruction).  Used by the CPU tests to compare the host
// parser with the reference's own records and to feed the oracle.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <chrono>
#include <vector>
#include "decoder.h"
#include "sinks.h"
#include "cavlc.h"
#include "pip_container.h"

using namespace lh264host;

namespace {
class NullSink : public PictureSink {   // timing iterations: the records are produced and dropped
 public:
  int configure (int, int, int) override { return 0; }
  int submit (const lh264_picture_desc&, const DecodedPictureInfo&) override { return 0; }
};
}

int main (int argc, char** argv) {
  if (argc < 3) { fprintf (stderr, "usage: %s in.264 out.lhfb [repeat] [out.lhmd] [out.pip (default stream only)]\n", argv[0]); return 2; }
  if (!cavlc_selfcheck()) { fprintf (stderr, "CAVLC tables failed the Kraft/prefix self-check\n"); return 3; }
  FILE* f = fopen (argv[1], "rb");
  if (!f) { perror (argv[1]); return 1; }
  std::vector<uint8_t> buf;
  uint8_t tmp[65536];
  size_t n;
  while ((n = fread (tmp, 1, sizeof (tmp), f)) > 0) buf.insert (buf.end(), tmp, tmp + n);
  fclose (f);
  const int repeat = argc > 3 ? atoi (argv[3]) : 1;
  double best = 1e30;
  DecoderStats st;
  for (int it = 0; it < repeat; it++) {
    LhfbSink fsink (it == 0 ? argv[2] : "/dev/null");
    NullSink nsink;
    PictureSink& sink = it == 0 ? static_cast<PictureSink&> (fsink) : static_cast<PictureSink&> (nsink);
    Decoder dec (&sink);
    if (getenv ("LH264_CHECK_REWRITE")) dec.enable_rewrite_check (true);   // decompress-direction self-check (cavlc_writer.h)
    if (argc > 4) { if (it == 0) fsink.set_model_path (argv[4]); dec.enable_model (true); }   // (timed with the model hand-off, as the runtime parses)
    const auto t0 = std::chrono::steady_clock::now();
    if (argc > 5 && it == 0) {                          // the unit loop of h264dec, writing the `.pip` default stream
      PipBoilerplate pip;
      const size_t size = buf.size();
      static const uint8_t kStart[4] = {0, 0, 0, 1};
      buf.insert (buf.end(), kStart, kStart + 4);       // sentinel, h264dec.cpp:238
      for (size_t pos = 0; pos < size;) {
        const size_t len = PipBoilerplate::next_unit (buf.data(), pos, size);
        if (len < 4) { pip.raw (buf.data() + pos, len); pos += len; continue; }
        pip.unit_begin (buf.data() + pos, len, dec.sps_seen());
        dec.decode (buf.data() + pos, len);
        UnitSliceInfo u;
        u.slice_decoded = dec.last_slice().decoded; u.cabac = dec.last_slice().cabac; u.rbsp = dec.last_slice().rbsp; u.header_bits = dec.last_slice().header_bits;
        pip.unit_end (u);
        pos += len;
      }
      buf.resize (size);
      const std::vector<uint8_t>& d = pip.finish();
      FILE* fo = fopen (argv[5], "wb");
      if (fo) { fwrite (d.data(), 1, d.size(), fo); fclose (fo); }
    } else dec.decode (buf.data(), buf.size());
    dec.flush();
    const double dt = std::chrono::duration<double> (std::chrono::steady_clock::now() - t0).count();
    if (dt < best) best = dt;
    st = dec.stats();
    if (it == 0 && !dec.output().empty()) {
      const DecodedPictureInfo& o = dec.output().front();
      fprintf (stderr, "crop %d %d %d %d\n", o.crop_x, o.crop_y, o.width, o.height);
    }
  }
  fprintf (stderr, "nals %llu slices %llu pictures %llu mbs %llu coef_blocks %llu errors %llu  parse %.4f s (%.1f pictures/s)\n",
           (unsigned long long)st.nals, (unsigned long long)st.slices, (unsigned long long)st.pictures, (unsigned long long)st.mbs,
           (unsigned long long)st.coef_blocks, (unsigned long long)st.errors, best, st.pictures / best);
  if (getenv ("LH264_CHECK_REWRITE")) {
    fprintf (stderr, "rewrite check: %llu slices re-serialised, %llu mismatches\n", (unsigned long long)st.rewritten_slices, (unsigned long long)st.rewrite_mismatches);
    if (st.rewrite_mismatches) return 5;
  }
  return st.errors ? 4 : 0;
}
