// model_emu_core.h — TEST INFRASTRUCTURE ONLY.  The emulated launch sequence of the recompression-model kernels (as model_host.inl
// launches them on the GPU) for ONE model stream, shared by model_emu.cpp (whole-stream tool) and mock_abi.cpp (CPU stand-in of
// the C ABI for runtime tests).  Include after cuda_emu.h and model_kernels.cuh.
#pragma once
#include <string>
#include <vector>

class ModelEmu {
 public:
  int submit (const lh264_model_picture& pic) {
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
    // The generation launches of model_launch (model_host.inl).  The picture's records and coefficients must outlive this call only
    // until it returns: the coder, which runs later (flush), reads the decision pool and the CodeJob, nothing of the picture.
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
    if ((int)jobs_.size() >= code_steps_) { const int rc = flush(); if (rc) return rc; }
    PoolDev pool{ &pool_used_, &pages_pending_, &pages_pending_other_ };
    emu::launch (1, 1, 32, [&] { k_model_frame (pd, 1, &poison); });
    emu::launch ((unsigned)blocks, 1, GEN_THREADS, [&] { k_model_count (pd, 1, total_recs, &poison); }, gen_smem);
    emu::launch (1, 1, 1024, [&] { k_model_scan (pd, 1, &pic_total, &poison); });
    for (int attempt = 0; attempt < 3; attempt++) {
      ArenaDev ar{ page_ptrs_.data(), &next_, (unsigned int)page_ptrs_.size() };
      emu::launch (1, 1, 1024, [&] { k_model_plan (1, &pic_total, &dec_base, &ctl, (unsigned long long)dec_.size(), ar, pool, &poison); });
      if (!ctl.skip) break;
      poison = 0;                                     // code what is pending (empties the pool), grow, plan again (model_settle_batch)
      const int rc = flush();
      if (rc) return rc;
      if (ctl.total + 32 > dec_.size()) dec_.assign ((size_t)ctl.total + ctl.total / 4 + 64, 0u);
      while (page_ptrs_.size() < (size_t)ctl.pages_in_use + 2 * ctl.pages_needed) { pages_.emplace_back (PAGE_SIZE, 0); page_ptrs_.push_back (pages_.back().data()); }
    }
    if (ctl.skip) return -2;
    uint32_t* dec = dec_.data();
    emu::launch ((unsigned)blocks, 1, GEN_THREADS, [&] { k_model_emit (pd, 1, total_recs, dec, &dec_base, &ctl); }, gen_smem);
    CodeJob job;
    memset (&job, 0, sizeof (job));
    emu::launch (1, 1, 128, [&] { k_model_commit (pd, 1, &dec_base, &pic_total, &ctl, &job); });
    jobs_.push_back (job);
    return st_.error ? -3 : 0;
  }
  // = model_flush_coder: ONE k_model_code launch codes the pending pictures of the stream in order
  int flush() {
    if (jobs_.empty()) return 0;
    const int T = (int)jobs_.size();
    std::vector<int> index (T, 0);                    // one stream = one block; step j's job row holds one job, at position 0
    ArenaDev ar{ page_ptrs_.data(), &next_, (unsigned int)page_ptrs_.size() };
    const CodeJob* jobs = jobs_.data();
    const int* idx = index.data();
    const uint32_t* dec = dec_.data();
    emu::launch (1, 1, 64, [&] { k_model_code (jobs, idx, T, 1, dec, ar); });
    jobs_.clear();
    pool_used_ = 0; pages_pending_ = 0;
    return st_.error ? -3 : 0;
  }
  void set_code_steps (int t) { code_steps_ = t < 1 ? 1 : t; }
  // = lh264_model_finish
  int finish() {
    if (!open_) return 0;
    { const int rc = flush(); if (rc) return rc; }
    while (page_ptrs_.size() < (size_t)next_ + CT_COUNT) { pages_.emplace_back (PAGE_SIZE, 0); page_ptrs_.push_back (pages_.back().data()); }
    ArenaDev ar{ page_ptrs_.data(), &next_, (unsigned int)page_ptrs_.size() };
    MState* st = &st_;
    emu::launch (1, 1, 32, [&] { k_model_finish (st, ar); });
    return st_.error;
  }
  // = lh264_model_read: bytes of the side stream of reference tag `tag` (empty: the tag has no file); false = broken page list
  bool tag_bytes (int tag, std::vector<uint8_t>& out) const {
    out.clear();
    if (!open_) return true;
    for (int ct = 0; ct < CT_COUNT; ct++) {
      if (kTagOfCompact[ct] != tag || ! ((st_.used >> ct) & 1u)) continue;
      for (uint32_t at = 0; at < st_.coder[ct].pos;) {
        const uint32_t idx = pagelist_[(size_t)ct * MAX_PAGES_PER_TAG + (at >> PAGE_SHIFT)];
        const uint32_t nb = std::min<uint32_t> (PAGE_SIZE - (at & (PAGE_SIZE - 1)), st_.coder[ct].pos - at);
        if (idx == 0) return false;
        out.insert (out.end(), page_ptrs_[idx - 1] + (at & (PAGE_SIZE - 1)), page_ptrs_[idx - 1] + (at & (PAGE_SIZE - 1)) + nb);
        at += nb;
      }
    }
    return true;
  }
 private:
  std::vector<CodeJob> jobs_;                          // pending: generated, not coded yet
  unsigned long long pool_used_ = 0;
  unsigned int pages_pending_ = 0, pages_pending_other_ = 0;
  int code_steps_ = getenv ("LH264_CODE_STEPS") ? atoi (getenv ("LH264_CODE_STEPS")) : 4;
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

