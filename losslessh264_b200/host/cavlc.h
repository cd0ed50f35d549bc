// cavlc.h — CAVLC residual block parsing (H.264 9.2).  Host side; replaces the reference's
// WelsResidualBlockCavlc (codec/decoder/core/src/parse_mb_syn_cavlc.cpp:863-1011) and its VLC
// lookup tables (vlc_decoder.h, decoder_data_tables.cpp).  The code tables below are the
// specification's Tables 9-5, 9-7, 9-8, 9-9 and 9-10 written as (length, code) pairs; decode
// LUTs are generated from them at start-up and checked for completeness (Kraft equality).
#pragma once
#include <stdint.h>
#include "bits.h"

namespace lh264host {

void cavlc_init();                 // idempotent, thread-safe after first call returns
bool cavlc_selfcheck();            // Kraft equality + prefix-freeness of every table

// Parses one residual_block_cavlc().  nC: -1 = ChromaDCLevel.  max_coeff 4, 15 or 16.
// Writes levels in scan order position -> out[scan[start + idx]] and returns total_coeff
// (0..16) or -1 on a malformed block.  `out` must be zeroed by the caller.
int cavlc_residual_block (BitReader& br, int nC, int max_coeff, const uint8_t* scan, int scan_step, int scan_off, int16_t* out);

extern const uint8_t kZigzag4x4[16];
extern const uint8_t kZigzag8x8[64];
extern const uint8_t kCbpIntra[48], kCbpInter[48];   // Table 9-4 (decoder.cpp)
extern const uint8_t kZtoR[16], kRtoZ[16];           // decode order <-> raster index of the 4x4 luma blocks

} // namespace lh264host
