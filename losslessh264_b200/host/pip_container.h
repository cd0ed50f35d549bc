// pip_container.h — the `.pip` container, compress direction.
//
// `x.pip` (no suffix) is the reference's CompressionStream::DEFAULT_STREAM: the "boilerplate" Annex-B
// skeleton of the input — start codes, NAL header bytes, re-escaped SPS/PPS/SEI RBSP, the bits of every
// slice HEADER followed by a stop bit, trailing zero bytes — with the slice DATA removed (it lives in the
// arithmetic-coded `x.pip.<tag>` side streams produced by the recompression model on the GPU).
// This file mirrors, byte for byte, what the reference writes through oMovie().def():
//   BitStream (appendByte / emitBits / padToByte / escape00xWith003x / stopEscape)
//                                   codec/decoder/core/src/compression_stream.cpp:51-133, inc/compression_stream.h:41-85
//   the unit loop of h264dec        codec/console/dec/src/h264dec.cpp:246-364 (units < 4 bytes are appended raw, :268-271)
//   WelsDecodeBs                    codec/decoder/core/src/decoder.cpp:656-859 (start code :681, trailing zeros :610-627,
//                                   stop bit + stopEscape on the flush call :836-846)
//   ParseNalHeader / ParseNonVclNal codec/decoder/core/src/au_parser.cpp:143 (header byte), :588 (parameter-set RBSP)
//   WelsDecodeSlice                 codec/decoder/core/src/decode_slice.cpp:2975-2980 (slice header bits)
//   CompressionStream::flushToWriter src/compression_stream.cpp:269-287 (final padToByte)
// Host code: byte shuffling that follows the parse, not a hot path.
#pragma once
#include <stdint.h>
#include <stdio.h>
#include <stddef.h>
#include <string.h>
#include <vector>

namespace lh264host {

class DefaultStream {                  // = BitStream
 public:
  void append_byte (uint8_t x);
  void append_bytes (const uint8_t* p, size_t n);
  void emit_bits (uint32_t bits, uint32_t n_bits);
  void emit_bit (uint32_t bit) { emit_bits (bit & 1u, 1); }
  void pad_to_byte (int val = 0);
  void escape_on() { escaping_ = true; }
  void stop_escape();
  const std::vector<uint8_t>& bytes() const { return buf_; }
 private:
  std::vector<uint8_t> buf_;
  bool escaping_ = false;
  uint8_t esc_[2] = {0, 0};
  uint32_t esc_n_ = 0, bits_ = 0;
  uint8_t n_bits_ = 0;
};

// What one decode unit (one h264dec Annex-B unit = one DecodeFrameNoDelay call) contributes.
struct UnitSliceInfo {
  bool slice_decoded = false;          // the unit held a slice NAL that was parsed and decoded
  bool cabac = false;
  const uint8_t* rbsp = nullptr;       // the slice NAL's RBSP (after the NAL header byte, emulation prevention removed)
  size_t header_bits = 0;              // bit position of slice_data() in it
};

class PipBoilerplate {
 public:
  // h264dec.cpp:262-267: length of the unit starting at buf[pos] (buf must be followed by 00 00 00 01)
  static size_t next_unit (const uint8_t* buf, size_t pos, size_t file_size);
  void raw (const uint8_t* p, size_t n) { def_.append_bytes (p, n); }
  // first half of DecodeFrameNoDelay: start code, NAL header, parameter-set / SEI RBSP, trailing zeros
  void unit_begin (const uint8_t* unit, size_t len, bool sps_seen);
  // the flush half: slice header bits and stop bit of the slice decoded from this unit, then stopEscape
  void unit_end (const UnitSliceInfo& s);
  const std::vector<uint8_t>& finish() { def_.pad_to_byte(); return def_.bytes(); }
  // the bytes written so far: the buffer only ever grows, so any prefix handed out stays valid (streaming flush)
  const std::vector<uint8_t>& so_far() const { return def_.bytes(); }
 private:
  DefaultStream def_;
  std::vector<uint8_t> nal_scratch_;   // unit_begin: the unescaped unit
};

// ---- single-file container (not in the reference, which writes <= ~35 side files per stream) -----------------
// `x.pipx` = one file holding the default stream and every tag stream of a `.pip` file set:
//   "LHPX", u32 version (1), u32 n_entries, n_entries x { u32 tag (0xffffffff = default stream), u64 length }, then the payloads
//   in table order.  Little endian.  The payload bytes are exactly the files of the set, so packing and unpacking are lossless
//   and a set unpacked from an archive still expands with the unmodified reference.
struct PipFileSet {
  std::vector<uint8_t> def;
  std::vector<std::vector<uint8_t> > tags;   // [LH264_NUM_TAGS]
  std::vector<bool> present;                 // a tag that was never used has no side file
};
bool pip_set_read (const char* pip_path, PipFileSet& out);            // x.pip + x.pip.<tag>
bool pip_set_write (const char* pip_path, const PipFileSet& in);
bool pip_archive_read (const char* archive_path, PipFileSet& out);
// Version 2 of the archive, written WHILE the stream is being compressed (streaming flush instead of the reference's exit-time
// FlushOnClose, h264dec.cpp:105-121): "LHPX", u32 version (2), then chunks { u32 tag (0xffffffff = default stream), u32 length,
// payload } in the order they became final, closed by the chunk { 0xfffffffe, 0 }.  The payloads of one tag, concatenated in file
// order, are that tag's file of the set; an archive without the closing chunk is a stream that was cut off (refused).
class PipStreamWriter {
 public:
  bool open (const char* path);
  bool chunk (uint32_t tag, const uint8_t* p, size_t n);
  bool close();                           // writes the closing chunk
  bool is_open() const { return f_ != nullptr; }
  ~PipStreamWriter() { if (f_) fclose (f_); }
 private:
  FILE* f_ = nullptr;
};
bool pip_archive_write (const char* archive_path, const PipFileSet& in);
inline bool is_pip_archive_name (const char* path) { const char* d = strrchr (path, '.'); return d && !strcmp (d, ".pipx"); }

} // namespace lh264host
