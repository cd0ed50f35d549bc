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

class ModelEmuSink : public PictureSink {
 public:
  int configure (int, int, int) override { return 0; }
  int submit (const lh264_picture_desc&, const DecodedPictureInfo&) override { return 0; }
  int submit_model (const lh264_model_picture& pic, const DecodedPictureInfo&) override {
    if (!open_) {                                     // = lh264_model_open (model_host.inl), level-5.2 sized like the API sink
      max_mbs_ = 36864;
      priors_.assign (P_END, 0u);
      frames_.assign ((size_t)3 * max_mbs_, CtxRec());
      memset (frames_.data(), 0, frames_.size() * sizeof (CtxRec));
      order_.assign (max_mbs_, 0);
      pagelist_.assign ((size_t)CT_COUNT * MAX_PAGES_PER_TAG, 0u);
      memset (&st_, 0, sizeof (st_));
      st_.priors = priors_.data(); st_.frame[0] = frames_.data(); st_.frame[1] = frames_.data() + max_mbs_; st_.newrec = frames_.data() + 2 * (size_t)max_mbs_;
      st_.order = order_.data(); st_.pagelist = pagelist_.data(); st_.max_mbs = max_mbs_;
      MState* st = &st_;
      emu::launch (1, 1, 32, [&] { k_model_init_state (st); });
      open_ = true;
    }
    if (pic.mb_w * pic.mb_h > max_mbs_) return -1;
    MPicDev d;
    memset (&d, 0, sizeof (d));
    std::vector<uint32_t> rec_cnt (pic.n_recs, 0u), rec_off (pic.n_recs, 0u);
    d.recs = pic.recs; d.slices = pic.slices; d.coefs = pic.coefs; d.st = &st_;
    d.rec_cnt = rec_cnt.data(); d.rec_off = rec_off.data();
    d.n_recs = pic.n_recs; d.n_slices = pic.n_slices; d.mb_w = pic.mb_w; d.mb_h = pic.mb_h; d.rec_base = 0;
    const MPicDev* pd = &d;
    unsigned int poison = 0;
    uint32_t pic_total = 0;
    unsigned long long dec_base = 0;
    BatchCtl ctl;
    memset (&ctl, 0, sizeof (ctl));
    const int total_recs = pic.n_recs, blocks = (total_recs + GEN_THREADS - 1) / GEN_THREADS;
    const size_t gen_smem = (size_t)GEN_THREADS * sizeof (ThreadMb);
    emu::launch (1, 1, 32, [&] { k_model_frame (pd, 1, &poison); });
    emu::launch ((unsigned)blocks, 1, GEN_THREADS, [&] { k_model_count (pd, 1, total_recs, &poison); }, gen_smem);
    emu::launch (1, 1, 1024, [&] { k_model_scan (pd, 1, &pic_total, &poison); });
    for (int attempt = 0; attempt < 2; attempt++) {
      ArenaDev ar{ page_ptrs_.data(), &next_, (unsigned int)page_ptrs_.size() };
      emu::launch (1, 1, 1024, [&] { k_model_plan (1, &pic_total, &dec_base, &ctl, (unsigned long long)dec_.size(), ar, &poison); });
      if (!ctl.skip) break;
      poison = 0;                                     // grow and plan again (model_complete, model_host.inl)
      if (ctl.total + 32 > dec_.size()) dec_.assign ((size_t)ctl.total + ctl.total / 4 + 64, 0u);
      while (page_ptrs_.size() < (size_t)ctl.pages_in_use + 2 * ctl.pages_needed) { pages_.emplace_back (PAGE_SIZE, 0); page_ptrs_.push_back (pages_.back().data()); }
    }
    if (ctl.skip) return -2;
    ArenaDev ar{ page_ptrs_.data(), &next_, (unsigned int)page_ptrs_.size() };
    uint32_t* dec = dec_.data();
    emu::launch ((unsigned)blocks, 1, GEN_THREADS, [&] { k_model_emit (pd, 1, total_recs, dec, &dec_base, &ctl); }, gen_smem);
    emu::launch (1, 1, 64, [&] { k_model_code (pd, 1, dec, &dec_base, &pic_total, ar, &ctl); });
    return st_.error ? -3 : 0;
  }
  // = lh264_model_finish + lh264_model_read for every tag
  int finish (const std::string& pip) {
    if (!open_) return 0;
    while (page_ptrs_.size() < (size_t)next_ + CT_COUNT) { pages_.emplace_back (PAGE_SIZE, 0); page_ptrs_.push_back (pages_.back().data()); }
    ArenaDev ar{ page_ptrs_.data(), &next_, (unsigned int)page_ptrs_.size() };
    MState* st = &st_;
    emu::launch (1, 1, 32, [&] { k_model_finish (st, ar); });
    if (st_.error) return st_.error;
    for (int ct = 0; ct < CT_COUNT; ct++) {
      if (! ((st_.used >> ct) & 1u)) continue;
      FILE* f = fopen ((pip + "." + std::to_string (kTagOfCompact[ct])).c_str(), "wb");
      if (!f) return -1;
      for (uint32_t at = 0; at < st_.coder[ct].pos;) {
        const uint32_t idx = pagelist_[(size_t)ct * MAX_PAGES_PER_TAG + (at >> PAGE_SHIFT)];
        const uint32_t nb = std::min<uint32_t> (PAGE_SIZE - (at & (PAGE_SIZE - 1)), st_.coder[ct].pos - at);
        if (idx == 0) { fclose (f); return -4; }
        fwrite (page_ptrs_[idx - 1] + (at & (PAGE_SIZE - 1)), 1, nb, f);
        at += nb;
      }
      fclose (f);
    }
    return 0;
  }
 private:
  bool open_ = false;
  int max_mbs_ = 0;
  MState st_;
  std::vector<uint32_t> priors_, pagelist_, dec_;
  std::vector<CtxRec> frames_;
  std::vector<int> order_;
  std::vector<std::vector<uint8_t> > pages_;
  std::vector<uint8_t*> page_ptrs_;
  unsigned int next_ = 0;
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
