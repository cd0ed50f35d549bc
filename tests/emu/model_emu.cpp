// model_emu.cpp — TEST INFRASTRUCTURE ONLY.  Compiles the recompression-model kernels of the product
// (losslessh264_b200/csrc/model_kernels.cuh, UNCHANGED sources) with g++ behind tests/emu/cuda_emu.h and runs them for whole
// streams:   model_emu in.264 out.pip
// Host decoder (parse) -> model hand-off records -> emulated k_model_frame / count / scan / plan / emit / code / finish, launched
// as model_host.inl launches them -> out.pip.<tag> side streams.  tests/test_emu_model.py compares them, byte for byte, with
// the side streams of the compiled reference (tests/golden/pip_tag_sha1.json).
#include "cuda_emu.h"
#include "../../losslessh264_b200/csrc/model_kernels.cuh"
#include "../../losslessh264_b200/host/decoder.h"
#include "../../losslessh264_b200/host/pip_container.h"
#include <string>
#include <vector>

using namespace lh264m;
using namespace lh264host;
#include "model_emu_core.h"

class ModelEmuSink : public PictureSink {
 public:
  int configure (int, int, int) override { return 0; }
  int submit (const lh264_picture_desc&, const DecodedPictureInfo&) override { return 0; }
  int submit_model (const lh264_model_picture& pic, const DecodedPictureInfo&) override { return m_.submit (pic); }
  // = lh264_model_finish + lh264_model_read for every tag
  int finish (const std::string& pip) {
    const int rc = m_.finish();
    if (rc) return rc;
    std::vector<uint8_t> b;
    for (int tag = 1; tag < LH264_NUM_TAGS; tag++) {
      if (!m_.tag_bytes (tag, b)) return -4;
      if (b.empty()) continue;
      FILE* f = fopen ((pip + "." + std::to_string (tag)).c_str(), "wb");
      if (!f) return -1;
      fwrite (b.data(), 1, b.size(), f);
      fclose (f);
    }
    return 0;
  }
 private:
  ModelEmu m_;
};

int main (int argc, char** argv) {
  if (argc < 3) { fprintf (stderr, "usage: %s in.264 out.pip\n", argv[0]); return 2; }
  FILE* f = fopen (argv[1], "rb");
  if (!f) { perror (argv[1]); return 1; }
  std::vector<uint8_t> buf;
  uint8_t tmp[65536];
  size_t n;
  while ((n = fread (tmp, 1, sizeof (tmp), f)) > 0) buf.insert (buf.end(), tmp, tmp + n);
  fclose (f);
  ModelEmuSink sink;
  Decoder dec (&sink);
  dec.enable_model (true);
  PipBoilerplate pip;
  const size_t size = buf.size();
  static const uint8_t kStart[4] = {0, 0, 0, 1};
  buf.insert (buf.end(), kStart, kStart + 4);          // sentinel of the unit loop, h264dec.cpp:238
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
  dec.flush();
  const std::vector<uint8_t>& d = pip.finish();
  FILE* fo = fopen (argv[2], "wb");
  if (!fo) { perror (argv[2]); return 1; }
  fwrite (d.data(), 1, d.size(), fo);
  fclose (fo);
  const int rc = sink.finish (argv[2]);
  if (rc) { fprintf (stderr, "model error %d\n", rc); return 3; }
  return 0;
}
