// h264_syntax.h — parameter sets, slice header and NAL framing of the host decoder.
//
// Host side of the hot path (stays on the CPU by design): this is what the reference does in
// codec/decoder/core/src/au_parser.cpp (ParseSps :~860, ParsePps :~1270, ParseSliceHeaderSyntaxes
// in decoder_core.cpp:~600-1100) and decoder.cpp:656-859 (NAL split + emulation prevention).
// Written from the H.264 specification (clauses 7.3.2.1, 7.3.2.2, 7.3.3); only what a
// frame-only I/P decoder needs is kept, the rest is parsed and dropped.
#pragma once
#include <stdint.h>
#include <vector>
#include "bits.h"

namespace lh264host {

enum { NAL_SLICE = 1, NAL_IDR = 5, NAL_SEI = 6, NAL_SPS = 7, NAL_PPS = 8, NAL_AUD = 9, NAL_END_SEQ = 10, NAL_END_STREAM = 11 };
enum { SLICE_P = 0, SLICE_B = 1, SLICE_I = 2, SLICE_SP = 3, SLICE_SI = 4 };

struct Sps {
  bool valid = false;
  int id = 0, profile_idc = 0, level_idc = 0, constraint_flags = 0;
  int chroma_format_idc = 1;
  int log2_max_frame_num = 4;
  int poc_type = 0, log2_max_poc_lsb = 4;
  bool delta_pic_order_always_zero = false;
  int offset_for_non_ref_pic = 0, offset_for_top_to_bottom = 0, num_ref_frames_in_poc_cycle = 0;
  int offset_for_ref_frame[256];
  int max_num_ref_frames = 1;
  bool gaps_in_frame_num_allowed = false;
  int mb_w = 0, mb_h = 0;
  bool frame_mbs_only = true, direct_8x8_inference = false;
  int crop_left = 0, crop_right = 0, crop_top = 0, crop_bottom = 0;   // in luma samples
  bool scaling_matrix_present = false;
  // scaling lists in raster position order, every list filled in by the fall-back rules (ParseScalingList, au_parser.cpp:1502-1562)
  uint8_t scale4[6][16], scale8[6][64];
};

struct Pps {
  bool valid = false;
  int id = 0, sps_id = 0;
  bool cabac = false, bottom_field_pic_order = false;
  int num_slice_groups = 1, slice_group_map_type = 0;
  int run_length[8] = {0, 0, 0, 0, 0, 0, 0, 0};        // slice_group_map_type 0: run_length_minus1 + 1 per group
  int num_ref_idx_l0_default = 1, num_ref_idx_l1_default = 1;
  bool weighted_pred = false;
  int weighted_bipred_idc = 0;
  int pic_init_qp = 26, pic_init_qs = 26;
  int chroma_qp_index_offset = 0, second_chroma_qp_index_offset = 0;
  bool deblocking_filter_control_present = false, constrained_intra_pred = false, redundant_pic_cnt_present = false;
  bool transform_8x8_mode = false, scaling_matrix_present = false;
  bool list_present[12] = {false, false, false, false, false, false, false, false, false, false, false, false};
  uint8_t scale4[6][16], scale8[6][64];              // as in Sps; fall-back from the SPS known when the PPS arrived
  bool sps_was_missing = false;                      // parsed before its SPS (the reference then reads 12 lists and patches later)
};

struct RefPicListMod { int op; uint32_t val; };
struct Mmco { int op; uint32_t a, b; };

struct SliceHeader {
  int nal_ref_idc = 0, nal_type = 0;
  bool idr = false;
  int first_mb = 0, slice_type = SLICE_P, pps_id = 0;
  int frame_num = 0, idr_pic_id = 0, poc_lsb = 0, delta_poc_bottom = 0, delta_poc[2] = {0, 0};
  int redundant_pic_cnt = 0;
  int num_ref_idx_l0 = 1;
  std::vector<RefPicListMod> mods;
  bool has_wp = false;
  int luma_log2_denom = 0, chroma_log2_denom = 0;
  int luma_weight[32], luma_offset[32], chroma_weight[32][2], chroma_offset[32][2];
  bool no_output_of_prior_pics = false, long_term_reference_flag = false, adaptive_marking = false;
  std::vector<Mmco> mmcos;
  int slice_qp = 26;
  int disable_deblocking_idc = 0, alpha_c0_offset = 0, beta_offset = 0;   // offsets already doubled
  int cabac_init_idc = 0;
  size_t header_bits = 0;     // bit position of slice_data()
};

// Annex B: splits [data, data+len) into NAL payloads (start code stripped).
struct NalSpan { const uint8_t* p; size_t len; };
void split_annexb (const uint8_t* data, size_t len, std::vector<NalSpan>& out);
// removes emulation_prevention_three_byte; appends 8 zero bytes of padding after the payload.
// Returns the payload length (without padding).
size_t unescape_nal (const uint8_t* p, size_t len, std::vector<uint8_t>& out);

bool parse_sps (BitReader& br, Sps& sps);
bool parse_pps (BitReader& br, Pps& pps, const Sps* sps_table /*[32]*/);
// returns 0 on success, negative on unsupported / malformed
int parse_slice_header (BitReader& br, int nal_ref_idc, int nal_type, const Sps* sps_table, const Pps* pps_table, SliceHeader& sh);

} // namespace lh264host
