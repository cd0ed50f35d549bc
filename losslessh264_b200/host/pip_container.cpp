// pip_container.cpp — see pip_container.h.
#include "pip_container.h"
#include <stdio.h>
#include <string.h>
#include <string>
#include "h264_syntax.h"

namespace lh264host {

void DefaultStream::append_byte (uint8_t x) {          // BitStream::appendByte, compression_stream.cpp:51-69
  if (!escaping_) { buf_.push_back (x); return; }
  if (x <= 3 && esc_n_ == 2 && esc_[0] == 0 && esc_[1] == 0) {
    buf_.push_back (0); buf_.push_back (0); buf_.push_back (3); buf_.push_back (x);
    esc_n_ = 0;
  } else if (esc_n_ == 2) {
    buf_.push_back (esc_[0]);
    esc_[0] = esc_[1]; esc_[1] = x;
  } else esc_[esc_n_++] = x;
}
void DefaultStream::append_bytes (const uint8_t* p, size_t n) { for (size_t i = 0; i < n; i++) append_byte (p[i]); }
void DefaultStream::emit_bits (uint32_t bits, uint32_t n) {      // BitStream::emitBits :89-112
  n += n_bits_;
  bits <<= 32 - n;
  bits |= bits_;
  while (n >= 8) { append_byte ((uint8_t) (bits >> 24)); bits <<= 8; n -= 8; }
  bits_ = bits; n_bits_ = (uint8_t)n;
}
void DefaultStream::pad_to_byte (int val) { for (int i = n_bits_; (i & 7) != 0; ++i) emit_bit ((uint32_t)val); }
void DefaultStream::stop_escape() {                    // compression_stream.h:71-84 (flushBits pads with zero bits)
  while (n_bits_ > 0) emit_bits (0, 1);
  escaping_ = false;
  const uint8_t tmp[2] = {esc_[0], esc_[1]};
  const uint32_t n = esc_n_;
  esc_n_ = 0;
  append_bytes (tmp, n);
}

size_t PipBoilerplate::next_unit (const uint8_t* b, size_t pos, size_t file_size) {
  size_t i = 0;
  // the unit ends at the next start code - at the latest the sentinel behind the last byte (b[file_size ..]): jump from zero byte
  // to zero byte inside that range, and fall back to the plain walk only if nothing was found there
  if (file_size > pos) {
    const size_t limit = file_size - pos;               // candidates i = 1 .. limit
    for (size_t k = 1; k <= limit; k++) {
      const uint8_t* z = static_cast<const uint8_t*> (memchr (b + pos + k, 0, limit - k + 1));
      if (!z) break;
      k = (size_t) (z - (b + pos));
      const uint8_t* q = z;
      if (q[1] == 0 && ((q[2] == 0 && q[3] == 1) || q[2] == 1)) return k;
    }
    i = limit + 1;
  }
  for (; i < file_size; i++) {
    const uint8_t* q = b + pos + i;
    if (i > 0 && q[0] == 0 && q[1] == 0 && ((q[2] == 0 && q[3] == 1) || q[2] == 1)) break;   // at the latest: the sentinel
  }
  return i;
}

void PipBoilerplate::unit_begin (const uint8_t* u, size_t len, bool sps_seen) {
  // DetectStartCodePrefix: everything up to and including the first 00 00 01
  size_t off = 0;
  for (size_t k = 0; k + 3 <= len; k++) if (u[k] == 0 && u[k + 1] == 0 && u[k + 2] == 1) { off = k + 3; break; }
  if (off == 0) return;                                // no start code: the reference rejects the buffer
  def_.append_bytes (u, off);
  std::vector<uint8_t>& nal = nal_scratch_;            // (reused: no allocation and no clearing per unit)
  const size_t n = unescape_nal (u + off, len - off, nal);
  def_.escape_on();
  if (n == 0) return;
  size_t zeros = 0;
  while (zeros < n && nal[n - 1 - zeros] == 0) zeros++;
  if (zeros < n && !(nal[0] & 0x80)) {                 // forbidden_zero_bit set: ParseNalHeader returns before the append
    def_.append_byte (nal[0]);
    const int type = nal[0] & 31;
    const bool param = type == NAL_SPS || type == NAL_PPS || type == 15 /* subset SPS */;
    if ((type == NAL_SEI || param) && (type == NAL_SEI || type == NAL_SPS || sps_seen) && n - zeros > 1)
      def_.append_bytes (nal.data() + 1, n - zeros - 1);
  }
  for (size_t i = 0; i < zeros; i++) def_.append_byte (0);
}

void PipBoilerplate::unit_end (const UnitSliceInfo& s) {
  if (s.slice_decoded) {
    for (size_t i = 0; i < s.header_bits; i++) def_.emit_bit ((uint32_t) (s.rbsp[i >> 3] >> (7 - (i & 7))) & 1u);
    if (!s.cabac) def_.emit_bit (1);
  }
  def_.stop_escape();
}


// ---- file sets and the single-file container --------------------------------------------------------------------
namespace {
bool slurp_file (const std::string& path, std::vector<uint8_t>& out) {
  FILE* f = fopen (path.c_str(), "rb");
  if (!f) return false;
  uint8_t tmp[65536];
  size_t n;
  out.clear();
  while ((n = fread (tmp, 1, sizeof (tmp), f)) > 0) out.insert (out.end(), tmp, tmp + n);
  fclose (f);
  return true;
}
bool spill_file (const std::string& path, const std::vector<uint8_t>& b) {
  FILE* f = fopen (path.c_str(), "wb");
  if (!f) return false;
  const bool ok = b.empty() || fwrite (b.data(), 1, b.size(), f) == b.size();
  return fclose (f) == 0 && ok;
}
void put32 (std::vector<uint8_t>& o, uint32_t v) { for (int i = 0; i < 4; i++) o.push_back ((uint8_t) (v >> (8 * i))); }
void put64 (std::vector<uint8_t>& o, uint64_t v) { for (int i = 0; i < 8; i++) o.push_back ((uint8_t) (v >> (8 * i))); }
uint64_t get_le (const uint8_t* p, int n) { uint64_t v = 0; for (int i = 0; i < n; i++) v |= (uint64_t)p[i] << (8 * i); return v; }
const int kNumTags = 70;               // LH264_NUM_TAGS (include/lh264_b200.h)
}

bool pip_set_read (const char* pip_path, PipFileSet& out) {
  out.tags.assign (kNumTags, std::vector<uint8_t>());
  out.present.assign (kNumTags, false);
  if (!slurp_file (pip_path, out.def)) return false;
  for (int t = 0; t < kNumTags; t++) out.present[t] = slurp_file (std::string (pip_path) + "." + std::to_string (t), out.tags[t]);
  return true;
}
bool pip_set_write (const char* pip_path, const PipFileSet& in) {
  if (!spill_file (pip_path, in.def)) return false;
  for (int t = 0; t < kNumTags && t < (int)in.tags.size(); t++)
    if (in.present[t] && !spill_file (std::string (pip_path) + "." + std::to_string (t), in.tags[t])) return false;
  return true;
}
bool pip_archive_write (const char* archive_path, const PipFileSet& in) {
  std::vector<uint8_t> o;
  uint32_t n = 1;
  for (int t = 0; t < (int)in.tags.size(); t++) if (in.present[t]) n++;
  o.insert (o.end(), {'L', 'H', 'P', 'X'});
  put32 (o, 1); put32 (o, n);
  put32 (o, 0xffffffffu); put64 (o, in.def.size());
  for (int t = 0; t < (int)in.tags.size(); t++) if (in.present[t]) { put32 (o, (uint32_t)t); put64 (o, in.tags[t].size()); }
  o.insert (o.end(), in.def.begin(), in.def.end());
  for (int t = 0; t < (int)in.tags.size(); t++) if (in.present[t]) o.insert (o.end(), in.tags[t].begin(), in.tags[t].end());
  return spill_file (archive_path, o);
}
bool PipStreamWriter::open (const char* path) {
  if (f_) fclose (f_);
  f_ = fopen (path, "wb");
  if (!f_) return false;
  std::vector<uint8_t> h = {'L', 'H', 'P', 'X'};
  put32 (h, 2);
  return fwrite (h.data(), 1, h.size(), f_) == h.size();
}
bool PipStreamWriter::chunk (uint32_t tag, const uint8_t* p, size_t n) {
  if (!f_) return false;
  while (n > 0) {                                      // (a chunk length is 32 bits)
    const size_t part = n > 0x40000000u ? 0x40000000u : n;
    std::vector<uint8_t> h;
    put32 (h, tag); put32 (h, (uint32_t)part);
    if (fwrite (h.data(), 1, h.size(), f_) != h.size() || fwrite (p, 1, part, f_) != part) return false;
    p += part; n -= part;
  }
  return fflush (f_) == 0;
}
bool PipStreamWriter::close() {
  if (!f_) return false;
  std::vector<uint8_t> h;
  put32 (h, 0xfffffffeu); put32 (h, 0);
  const bool ok = fwrite (h.data(), 1, h.size(), f_) == h.size();
  const bool ok2 = fclose (f_) == 0;
  f_ = nullptr;
  return ok && ok2;
}
bool pip_archive_read (const char* archive_path, PipFileSet& out) {
  std::vector<uint8_t> b;
  if (!slurp_file (archive_path, b) || b.size() < 12 || memcmp (b.data(), "LHPX", 4)) return false;
  if (get_le (b.data() + 4, 4) == 2) {                 // chunked form (PipStreamWriter)
    out.tags.assign (kNumTags, std::vector<uint8_t>());
    out.present.assign (kNumTags, false);
    out.def.clear();
    size_t at = 8;
    for (;;) {
      if (b.size() - at < 8) return false;             // no closing chunk: the writer was cut off
      const uint32_t tag = (uint32_t)get_le (b.data() + at, 4);
      const uint64_t len = get_le (b.data() + at + 4, 4);
      at += 8;
      if (tag == 0xfffffffeu) return len == 0 && at == b.size();
      if (len > b.size() - at) return false;
      if (tag == 0xffffffffu) out.def.insert (out.def.end(), b.begin() + at, b.begin() + at + len);
      else if (tag < (uint32_t)kNumTags) { out.present[tag] = true; out.tags[tag].insert (out.tags[tag].end(), b.begin() + at, b.begin() + at + len); }
      else return false;
      at += len;
    }
  }
  if (get_le (b.data() + 4, 4) != 1) return false;
  const uint64_t n = get_le (b.data() + 8, 4);
  if (n < 1 || n > (uint64_t)kNumTags + 1 || b.size() < 12 + 12 * n) return false;
  out.tags.assign (kNumTags, std::vector<uint8_t>());
  out.present.assign (kNumTags, false);
  out.def.clear();
  uint64_t at = 12 + 12 * n;
  bool have_def = false;
  for (uint64_t i = 0; i < n; i++) {
    const uint32_t tag = (uint32_t)get_le (b.data() + 12 + 12 * i, 4);
    const uint64_t len = get_le (b.data() + 16 + 12 * i, 8);
    if (len > b.size() - at) return false;             // truncated or corrupt table
    if (tag == 0xffffffffu) { if (have_def) return false; out.def.assign (b.begin() + at, b.begin() + at + len); have_def = true; }
    else if (tag < (uint32_t)kNumTags && !out.present[tag]) { out.tags[tag].assign (b.begin() + at, b.begin() + at + len); out.present[tag] = true; }
    else return false;
    at += len;
  }
  return have_def && at == b.size();
}

} // namespace lh264host
