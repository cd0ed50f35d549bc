// sinks_file.cpp — LhfbSink, see sinks.h.
#include "sinks.h"
#include <vector>
#include <string.h>

namespace lh264host {

LhfbSink::LhfbSink (const char* path) : f_ (fopen (path, "wb")) {}
LhfbSink::~LhfbSink() { if (f_) fclose (f_); if (fm_) fclose (fm_); }
void LhfbSink::set_model_path (const char* path) { if (fm_) fclose (fm_); fm_ = path ? fopen (path, "wb") : nullptr; }
int LhfbSink::submit_model (const lh264_model_picture& pic, const DecodedPictureInfo&) {
  if (!fm_) return 0;
  const int32_t hdr[6] = {0x444d484c, pic.mb_w, pic.mb_h, pic.n_slices, pic.n_recs, pic.n_coef_blocks};
  fwrite (hdr, 4, 6, fm_);
  if (pic.n_slices) fwrite (pic.slices, sizeof (lh264_model_slice), pic.n_slices, fm_);     // (a damaged picture may bring none: no null pointer to fwrite)
  if (pic.n_recs) fwrite (pic.recs, sizeof (lh264_model_mb), pic.n_recs, fm_);
  if (pic.n_coef_blocks) fwrite (pic.coefs, 32, pic.n_coef_blocks, fm_);
  return 0;
}
int LhfbSink::configure (int mb_w, int mb_h, int) { mb_w_ = mb_w; mb_h_ = mb_h; return 0; }
int LhfbSink::submit (const lh264_picture_desc& pic, const DecodedPictureInfo& info) {
  if (!f_) return -1;
  // A picture that is not handed out is still written when it was concealed (info.error with every macroblock present): later
  // pictures predict from it.  Bit 8 of the nal_ref_idc word marks "decoded, not output".  (A picture held back behind a lost IDR
  // picture is left out as before: nothing that is output depends on it.)
  if (!info.output && ! (info.error && !info.complete)) return 0;
  auto w32 = [&] (int32_t v) { fwrite (&v, 4, 1, f_); };
  for (int s = 0; s < pic.n_slices; s++) {
    const lh264_slice_rec& sl = pic.slices[s];
    w32 (0x4246484c); w32 (1);
    w32 (mb_w_); w32 (mb_h_); w32 (pic.cur_slot); w32 (info.frame_num); w32 (info.poc_lsb); w32 (info.idr_pic_id); w32 (info.nal_ref_idc | (info.output ? 0 : 0x100)); w32 (info.serial);
    fwrite (&sl, sizeof (sl), 1, f_);
    // the slice's macroblocks in address order (with slice groups they are not a contiguous range)
    std::vector<int> addr;
    for (int a = sl.first_mb; a < mb_w_ * mb_h_ && (int)addr.size() < sl.num_mbs; a++)
      if (pic.mbs[a].mb_type != LH264_MB_NONE && pic.mbs[a].slice_idx == s) addr.push_back (a);
    w32 ((int32_t)addr.size());
    uint32_t lo = 0xffffffffu, hi = 0;
    for (int a : addr) {
      const lh264_mb_rec& m = pic.mbs[a];
      if (m.coded_mask) { lo = lo < m.coef_off ? lo : m.coef_off; const uint32_t e = m.coef_off + __builtin_popcount (m.coded_mask); hi = hi > e ? hi : e; }
    }
    if (lo == 0xffffffffu) lo = hi = 0;
    for (int a : addr) {
      lh264_mb_rec m = pic.mbs[a];
      m.slice_idx = 0;
      if (m.coded_mask) m.coef_off -= lo; else m.coef_off = 0;
      w32 (a);
      fwrite (&m, sizeof (m), 1, f_);
    }
    w32 ((int32_t) (hi - lo));
    if (hi > lo) {
      fwrite (pic.coefs + (size_t)lo * 16, 32, hi - lo, f_);
      std::vector<int16_t> z ((size_t) (hi - lo) * 16, 0);
      fwrite (z.data(), 32, hi - lo, f_);
    }
    w32 (0);
  }
  return 0;
}

} // namespace lh264host
