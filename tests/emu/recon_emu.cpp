// recon_emu.cpp — TEST INFRASTRUCTURE ONLY.  Compiles the reconstruction kernels of the product
// (losslessh264_b200/csrc/recon_rows.cuh and what it includes, UNCHANGED sources) with g++ behind tests/emu/cuda_emu.h and runs
// them for whole streams:   recon_emu in.264 out.yuv [warps]
// Host decoder (parse) -> hand-off records -> emulated k_decode_rows -> cropped I420, as the CLI of the
// product does through the C ABI on the GPU.  tests/test_emu_recon.py compares the SHA1 of out.yuv with the reference's.
#include "cuda_emu.h"
#include "../../losslessh264_b200/csrc/decode_rows.cuh"
#include "../../losslessh264_b200/host/decoder.h"
#include "../../losslessh264_b200/host/pip_container.h"
#include <vector>

using namespace lh264;
using namespace lh264host;

static int g_warps = 2;

class EmuSink : public PictureSink {
 public:
  int configure (int mb_w, int mb_h, int n_slots) override {
    mb_w_ = mb_w; mb_h_ = mb_h; n_slots_ = n_slots;
    stride_ = (size_t)mb_w * mb_h * 384;
    pool_.assign (stride_ * n_slots + 64, 128);      // pic_queue.cpp:96
    lines_.assign ((size_t)mb_h * 2 * mb_w * 16, 0);
    return 0;
  }
  int submit (const lh264_picture_desc& pic, const DecodedPictureInfo&) override {
    PicDev d;
    memset (&d, 0, sizeof (d));
    d.mbs = pic.mbs; d.coefs = pic.coefs; d.slices = pic.slices;
    d.base = pool_.data(); d.slot_stride = stride_; d.cur_slot = pic.cur_slot; d.n_slots = n_slots_; d.mb_w = mb_w_; d.mb_h = mb_h_;
    d.n_slices = pic.n_slices; d.n_coef_blocks = pic.n_coef_blocks;
    d.lines = lines_.data();
    std::vector<lh264_mb_rec> corrupt;                 // LH264_EMU_CORRUPT: one record per picture names tables it must not reach
    if (getenv ("LH264_EMU_CORRUPT")) {
      corrupt.assign (pic.mbs, pic.mbs + (size_t)mb_w_ * mb_h_);
      lh264_mb_rec& m = corrupt[corrupt.size() / 2];
      if (serial_++ & 1) m.slice_idx = 60000; else { m.coded_mask = 0xffffffu; m.coef_off = 0x7fffffffu; m.cbp = 47; }
      d.mbs = corrupt.data();
    }
    const int max_rows = mb_h_;
    std::vector<int> sync (4 + (size_t)max_rows + 16, 0);
    DecodeSync* ds = reinterpret_cast<DecodeSync*> (sync.data());
    const PicDev* pd = &d;
    unsigned int err = 0;
    emu::launch (1, 1, (unsigned)g_warps * 32, [&] { k_decode_rows (pd, 1, max_rows, ds, &err); });
    if (err) { fprintf (stderr, "record validation failed\n"); return -1; }
    return 0;
  }
  void read (const DecodedPictureInfo& o, FILE* f) const {
    const int W = mb_w_ * 16, H = mb_h_ * 16;
    const uint8_t* y = pool_.data() + (size_t)o.slot * stride_;
    const uint8_t* u = y + (size_t)W * H;
    const uint8_t* v = u + (size_t)W * H / 4;
    for (int r = 0; r < o.height; r++) fwrite (y + (size_t) (o.crop_y + r) * W + o.crop_x, 1, o.width, f);
    const int cw = (o.width + 1) >> 1, ch = (o.height + 1) >> 1;
    for (int r = 0; r < ch; r++) fwrite (u + (size_t) ((o.crop_y >> 1) + r) * (W >> 1) + (o.crop_x >> 1), 1, cw, f);
    for (int r = 0; r < ch; r++) fwrite (v + (size_t) ((o.crop_y >> 1) + r) * (W >> 1) + (o.crop_x >> 1), 1, cw, f);
  }
 private:
  int mb_w_ = 0, mb_h_ = 0, n_slots_ = 0;
  size_t stride_ = 0;
  std::vector<uint8_t> pool_, lines_;
  int serial_ = 0;
};

int main (int argc, char** argv) {
  if (argc < 3) { fprintf (stderr, "usage: %s in.264 out.yuv [warps per CTA]\n", argv[0]); return 2; }
  if (argc > 3) g_warps = atoi (argv[3]);
  FILE* f = fopen (argv[1], "rb");
  if (!f) { perror (argv[1]); return 1; }
  std::vector<uint8_t> buf;
  uint8_t tmp[65536];
  size_t n;
  while ((n = fread (tmp, 1, sizeof (tmp), f)) > 0) buf.insert (buf.end(), tmp, tmp + n);
  fclose (f);
  FILE* out = fopen (argv[2], "wb");
  if (!out) { perror (argv[2]); return 1; }
  EmuSink sink;
  Decoder dec (&sink);
  const size_t size = buf.size();
  static const uint8_t kStart[4] = {0, 0, 0, 1};
  buf.insert (buf.end(), kStart, kStart + 4);          // sentinel of the unit loop, h264dec.cpp:238
  int frames = 0;
  auto drain = [&] () {
    for (const DecodedPictureInfo& o : dec.output()) { sink.read (o, out); frames++; }
    dec.clear_output();
  };
  for (size_t pos = 0; pos < size;) {
    const size_t len = PipBoilerplate::next_unit (buf.data(), pos, size);
    if (len >= 4) { dec.decode (buf.data() + pos, len); drain(); }
    pos += len;
  }
  dec.flush();
  drain();
  fclose (out);
  fprintf (stderr, "frames %d\n", frames);
  return 0;
}
