// pip_expand.cpp — see pip_expand.h.
#include "pip_expand.h"
#include "pip_container.h"
#include "h264_syntax.h"
#include <string.h>

namespace lh264host {

namespace {
struct Bits {                                          // RBSP under construction
  std::vector<uint8_t> buf; size_t n = 0;
  void bit (uint32_t b) { if ((n & 7) == 0) buf.push_back (0); if (b) buf.back() |= (uint8_t) (0x80u >> (n & 7)); n++; }
  void copy (const uint8_t* src, size_t first_bit, size_t count) {
    size_t i = 0;
    if ((n & 7) == 0 && (first_bit & 7) == 0) {        // byte-aligned fast path
      const size_t bytes = count >> 3;
      buf.insert (buf.end(), src + (first_bit >> 3), src + (first_bit >> 3) + bytes);
      n += bytes * 8; i = bytes * 8;
    }
    for (; i < count; i++) { const size_t a = first_bit + i; bit ((src[a >> 3] >> (7 - (a & 7))) & 1u); }
  }
};
size_t start_code_end (const uint8_t* u, size_t len) {
  for (size_t k = 0; k + 3 <= len; k++) if (u[k] == 0 && u[k + 1] == 0 && u[k + 2] == 1) return k + 3;
  return 0;
}
} // namespace

bool expand_plan (const uint8_t* def_in, size_t n, ExpandPlan& plan) {
  plan = ExpandPlan();
  std::vector<uint8_t> def (def_in, def_in + n);
  static const uint8_t kStart[4] = {0, 0, 0, 1};
  def.insert (def.end(), kStart, kStart + 4);           // sentinel for next_unit (h264dec.cpp:238)
  std::vector<Sps> sps (32);
  std::vector<Pps> pps (256);
  std::vector<uint8_t> nal;
  auto literal = [&] (size_t off, size_t len) {
    if (!len) return;
    if (!plan.pieces.empty() && plan.pieces.back().slice < 0 && plan.pieces.back().def_off + plan.pieces.back().def_len == off) plan.pieces.back().def_len += len;
    else { ExpandPiece p; p.def_off = off; p.def_len = len; plan.pieces.push_back (p); }
  };
  for (size_t pos = 0; pos < n;) {
    const size_t len = PipBoilerplate::next_unit (def.data(), pos, n);
    const uint8_t* u = def.data() + pos;
    const size_t off = len >= 4 ? start_code_end (u, len) : 0;
    if (off == 0 || off >= len) { literal (pos, len); pos += len; continue; }
    const size_t plen = unescape_nal (u + off, len - off, nal);
    const int type = nal[0] & 31, ref_idc = (nal[0] >> 5) & 3;
    const bool forbidden = nal[0] & 0x80;
    if (forbidden || plen < 1) { literal (pos, len); pos += len; continue; }
    if (type == NAL_SPS) {
      BitReader br (nal.data() + 1, plen - 1);
      Sps s;
      if (parse_sps (br, s)) { sps[s.id] = s; if (s.mb_w * s.mb_h > plan.max_mbs) plan.max_mbs = s.mb_w * s.mb_h; }
      literal (pos, len);
    } else if (type == NAL_PPS) {
      BitReader br (nal.data() + 1, plen - 1);
      Pps p;
      if (parse_pps (br, p, sps.data())) { pps[p.id] = p; if (p.transform_8x8_mode) plan.with_8x8 = true; }
      literal (pos, len);
    } else if (type == NAL_SLICE || type == NAL_IDR) {
      if (plen < 2) { plan.error = "a slice was dropped at compression time (only its NAL header is in the default stream)"; return false; }
      // The unit holds <slice header bits><stop bit (CAVLC only)><zero padding to the byte> (pip_container.cpp: unit_end) and its last
      // two bytes are written unescaped (stop_escape).  When a CABAC header's last byte is 0x00 that byte runs into the zeros of the
      // next start code prefix and next_unit() - the reference's own unit split, h264dec.cpp:238 - hands it to the NEXT unit
      // (test_cif_I_CABAC_slice.264: `.. 21 00 02 00 | 00 00 01 65` is read as `.. 21 00 02 | 00 00 00 01 65`).  The header's length is
      // known from its own syntax, so parse it over zero-extended bytes and take the missing zero bytes back from the next unit.
      std::vector<uint8_t> ext (nal.begin(), nal.begin() + plen);
      ext.resize (plen + 24, 0);
      BitReader br (ext.data() + 1, plen - 1 + 8);
      SliceHeader sh;
      if (parse_slice_header (br, ref_idc, type, sps.data(), pps.data(), sh) < 0) { plan.error = "slice header does not parse"; return false; }
      if (pps[sh.pps_id].num_slice_groups > 1) { plan.error = "slice groups (FMO): the decompress direction walks macroblocks in raster order"; return false; }
      const Pps& pp = pps[sh.pps_id];
      const Sps& sp = sps[pp.sps_id];
      const uint8_t* rb = ext.data() + 1;
      const size_t hb = sh.header_bits;
      size_t rl = plen - 1, missing = 0;
      if (pp.cabac && (hb + 7) / 8 > rl) {
        missing = (hb + 7) / 8 - rl;
        // the bytes taken back must be zeros in front of the next start code, and a start code (00 00 01) must remain
        size_t z = 0;
        while (pos + len + z < n + 4 && def[pos + len + z] == 0) z++;
        if (missing + 2 > z || def[pos + len + z] != 1) { plan.error = "slice header runs past its unit"; return false; }
        rl += missing;
      }
      bool ok;
      if (pp.cabac) {
        ok = (hb + 7) / 8 == rl || ((hb & 7) == 0 && hb / 8 == rl);
        if (ok && (hb & 7)) ok = (rb[hb >> 3] & ((1u << (8 - (hb & 7))) - 1u)) == 0;
      } else {
        ok = (hb >> 3) < rl && ((rb[hb >> 3] >> (7 - (hb & 7))) & 1u) && (hb + 8) / 8 == rl;
        if (ok) ok = (rb[hb >> 3] & ((1u << (7 - (hb & 7))) - 1u)) == 0;
      }
      if (!ok) { plan.error = "slice unit of the default stream is not <header bits><stop bit> (trailing zero bytes in the original NAL?)"; return false; }
      if (sh.slice_type != SLICE_P && sh.slice_type != SLICE_I) { plan.error = "B/SP/SI slice"; return false; }
      lh264_decode_slice ds;
      memset (&ds, 0, sizeof (ds));
      ds.first_mb = (uint32_t)sh.first_mb; ds.frame_num = sh.frame_num;
      ds.mb_w = (uint16_t)sp.mb_w; ds.mb_h = (uint16_t)sp.mb_h;
      ds.slice_type = (uint8_t) (sh.slice_type == SLICE_I ? 2 : 0);
      ds.num_ref_idx_l0 = (uint8_t) (sh.slice_type == SLICE_P ? sh.num_ref_idx_l0 : pp.num_ref_idx_l0_default);
      ds.slice_qp = (uint8_t)sh.slice_qp;
      ds.flags = (uint8_t) ((pp.transform_8x8_mode ? 1 : 0) | (pp.constrained_intra_pred ? 2 : 0) | (pp.cabac ? 4 | ((sh.cabac_init_idc & 3) << 3) : 0));
      ds.header_bits = (uint32_t)hb;
      literal (pos, off);
      ExpandPiece p; p.def_off = pos + off; p.def_len = len - off; p.slice = (int)plan.slices.size();
      plan.pieces.push_back (p);
      plan.slices.push_back (ds);
      pos += missing;                                   // those zero bytes are part of this slice's header, not of the next start code
    } else literal (pos, len);
    pos += len;
  }
  return true;
}

bool expand_assemble (const ExpandPlan& plan, const uint8_t* def, const uint8_t* slice_bits, const lh264_decode_slice_out* outs, std::vector<uint8_t>& out) {
  std::vector<uint8_t> nal;
  for (const ExpandPiece& p : plan.pieces) {
    if (p.slice < 0) { out.insert (out.end(), def + p.def_off, def + p.def_off + p.def_len); continue; }
    const lh264_decode_slice& ds = plan.slices[p.slice];
    const lh264_decode_slice_out& so = outs[p.slice];
    unescape_nal (def + p.def_off, p.def_len, nal);
    Bits b;
    b.buf.reserve (ds.header_bits / 8 + so.n_bits / 8 + 8);
    b.copy (nal.data() + 1, 0, ds.header_bits);
    b.copy (slice_bits + so.byte_offset, 0, so.n_bits);
    b.bit (1);
    for (int i = so.n_pad_bits - 1; i >= 0; i--) b.bit ((so.pad_bits >> i) & 1u);
    if (b.n & 7) return false;
    out.push_back (nal[0]);
    int zeros = 0;                                     // emulation prevention (7.4.1)
    for (uint8_t x : b.buf) {
      if (zeros >= 2 && x <= 3) { out.push_back (3); zeros = 0; }
      out.push_back (x);
      zeros = x == 0 ? zeros + 1 : 0;
    }
  }
  return true;
}

} // namespace lh264host

// ---- C entry points (include/lh264_codec_api.h) ---------------------------------------------------------------------
struct lh264_pip_plan_s { lh264host::ExpandPlan plan; };
extern "C" {
lh264_pip_plan_s* lh264_pip_plan (const unsigned char* def, size_t len, char* err, size_t err_cap) {
  lh264_pip_plan_s* p = new lh264_pip_plan_s();
  if (!lh264host::expand_plan (def, len, p->plan)) {
    if (err && err_cap) { strncpy (err, p->plan.error.c_str(), err_cap - 1); err[err_cap - 1] = 0; }
    delete p;
    return nullptr;
  }
  return p;
}
void lh264_pip_plan_free (lh264_pip_plan_s* p) { delete p; }
const lh264_decode_slice* lh264_pip_plan_slices (const lh264_pip_plan_s* p, int* n_slices, int* max_mbs, int* with_8x8) {
  if (!p) return nullptr;
  if (n_slices) *n_slices = (int)p->plan.slices.size();
  if (max_mbs) *max_mbs = p->plan.max_mbs;
  if (with_8x8) *with_8x8 = p->plan.with_8x8 ? 1 : 0;
  return p->plan.slices.data();
}
// single-file container (pip_container.h): x.pip + x.pip.<tag>  <->  x.pipx.  0 on success.
int lh264_pip_pack (const char* pip_path, const char* archive_path) {
  lh264host::PipFileSet fs;
  if (!pip_path || !archive_path || !lh264host::pip_set_read (pip_path, fs)) return -1;
  return lh264host::pip_archive_write (archive_path, fs) ? 0 : -2;
}
int lh264_pip_unpack (const char* archive_path, const char* pip_path) {
  lh264host::PipFileSet fs;
  if (!pip_path || !archive_path || !lh264host::pip_archive_read (archive_path, fs)) return -1;
  return lh264host::pip_set_write (pip_path, fs) ? 0 : -2;
}
long long lh264_pip_assemble (const lh264_pip_plan_s* p, const unsigned char* def, const unsigned char* bits, const lh264_decode_slice_out* outs,
                              unsigned char* dst, size_t cap) {
  if (!p) return -1;
  std::vector<uint8_t> out;
  if (!lh264host::expand_assemble (p->plan, def, bits, outs, out)) return -1;
  if (dst && out.size() <= cap) memcpy (dst, out.data(), out.size());
  return (long long)out.size();
}
}
