// pip_expand.h — the `.pip` container, decompress direction.
//
// The default stream of a `.pip` is the Annex-B skeleton of the original file with every slice's DATA removed
// (pip_container.h).  The reference expands it by running its decoder over the skeleton and letting the slice-data parser
// pull symbols from the arithmetic-coded side streams (h264dec.cpp:246-364 in round-trip mode, WelsDecodeSlice
// decode_slice.cpp:2959-3171, slice data through EncoderState :1442-1835).  Here the same happens in three steps:
//   1. expand_plan():     walk the skeleton, keep parameter sets, parse every slice header -> lh264_decode_slice list;
//   2. lh264_model_decode() on the GPU (csrc/model_decode.cuh): side streams -> slice_data() bits of every slice;
//   3. expand_assemble(): header bits + slice data + stop bit + alignment bits, emulation prevention, Annex-B framing.
// Host code on both sides of the hot path: byte shuffling, not a hot path.
#pragma once
#include <stdint.h>
#include <stddef.h>
#include <string>
#include <vector>
#include "../../include/lh264_b200.h"

namespace lh264host {

struct ExpandPiece {
  size_t def_off = 0, def_len = 0;     // literal bytes of the default stream ...
  int slice = -1;                      // ... or (>= 0) a slice NAL: def_off/def_len = its escaped payload (NAL header byte included)
};

struct ExpandPlan {
  std::vector<ExpandPiece> pieces;
  std::vector<lh264_decode_slice> slices;
  int max_mbs = 0;
  bool with_8x8 = false;
  std::string error;                   // non-empty: the skeleton cannot be expanded (CABAC slices, trailing zero bytes inside a slice unit ...)
};

bool expand_plan (const uint8_t* def, size_t n, ExpandPlan& plan);
// slice_bits / outs: what lh264_model_decode() returned for plan.slices
bool expand_assemble (const ExpandPlan& plan, const uint8_t* def, const uint8_t* slice_bits, const lh264_decode_slice_out* outs, std::vector<uint8_t>& out264);

} // namespace lh264host
