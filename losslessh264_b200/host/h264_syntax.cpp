// h264_syntax.cpp — see h264_syntax.h.  H.264 clauses 7.3.1 (NAL), 7.3.2.1 (SPS), 7.3.2.2 (PPS),
// 7.3.3 (slice header), B.1 (byte stream).
#include "h264_syntax.h"
#include <string.h>

namespace lh264host {

void split_annexb (const uint8_t* d, size_t len, std::vector<NalSpan>& out) {
  out.clear();
  size_t i = 0;
  // find first start code
  auto find_sc = [&] (size_t from, size_t& sc_len) -> size_t {
    for (size_t k = from; k + 3 <= len; k++) {
      // entropy-coded bytes are rarely zero: jump from zero byte to zero byte (memchr runs at tens of bytes per cycle)
      const uint8_t* z = static_cast<const uint8_t*> (memchr (d + k, 0, len - 2 - k));
      if (!z) break;
      k = (size_t) (z - d);
      if (d[k + 1] == 0) {
        if (d[k + 2] == 1) { sc_len = 3; return k; }
        if (k + 4 <= len && d[k + 2] == 0 && d[k + 3] == 1) { sc_len = 4; return k; }
      }
    }
    sc_len = 0;
    return len;
  };
  size_t scl = 0;
  i = find_sc (0, scl);
  while (i < len) {
    const size_t start = i + scl;
    size_t nscl = 0;
    const size_t next = find_sc (start, nscl);
    size_t end = next;
    while (end > start && d[end - 1] == 0) end--;       // trailing_zero_8bits
    if (end > start) out.push_back (NalSpan{d + start, end - start});
    i = next; scl = nscl;
  }
}

size_t unescape_nal (const uint8_t* p, size_t len, std::vector<uint8_t>& out) {
  if (out.size() < len + 8) out.resize (len + 8);       // (a reused vector is not cleared again: every byte up to o + 8 is written below)
  uint8_t* dst = out.data();
  size_t o = 0, i = 0;
  int zeros = 0;
  while (i < len) {
    if (zeros == 0) {                                    // no zero run open: everything up to the next zero byte is copied as it is
      const uint8_t* z = static_cast<const uint8_t*> (memchr (p + i, 0, len - i));
      const size_t n = z ? (size_t) (z - (p + i)) : len - i;
      memcpy (dst + o, p + i, n);
      o += n; i += n;
      if (i >= len) break;
    }
    const uint8_t b = p[i++];
    if (zeros >= 2 && b == 3) { zeros = 0; continue; }   // emulation_prevention_three_byte
    dst[o++] = b;
    zeros = b == 0 ? zeros + 1 : 0;
  }
  memset (dst + o, 0, 8);
  out.resize (o + 8);
  return o;
}

// ---- scaling lists (7.3.2.1.1.1, fall-back rules of 7.4.2.1.1 / 7.4.2.2) as the reference applies them ------------------------
// ParseScalingList / SetScalingListValue, au_parser.cpp:1475-1562.  Lists are kept in raster position order.
static const uint8_t kZig4[16] = {0, 1, 4, 8, 5, 2, 3, 6, 9, 12, 13, 10, 7, 11, 14, 15};
static uint8_t g_zig8[64];
static uint8_t g_default4[2][16], g_default8[2][64];       // Default_4x4_Intra / _Inter, Default_8x8_Intra / _Inter (Tables 7-3, 7-4)
static bool init_scaling_tables() {
  // 8x8 frame zig-zag: walk the anti-diagonals
  int n = 0;
  for (int d = 0; d < 15; d++)
    for (int k = 0; k <= d; k++) {
      const int i = (d & 1) ? k : d - k, j = d - i;       // odd diagonals run down-left, even ones up-right
      if (i < 8 && j < 8) g_zig8[n++] = (uint8_t) (8 * i + j);
    }
  static const uint8_t z4[2][16] = { {6, 13, 13, 20, 20, 20, 28, 28, 28, 28, 32, 32, 32, 37, 37, 42}, {10, 14, 14, 20, 20, 20, 24, 24, 24, 24, 27, 27, 27, 30, 30, 34} };
  static const uint8_t z8[2][64] = {
    {6, 10, 10, 13, 11, 13, 16, 16, 16, 16, 18, 18, 18, 18, 18, 23, 23, 23, 23, 23, 23, 25, 25, 25, 25, 25, 25, 25, 27, 27, 27, 27, 27, 27, 27, 27,
     29, 29, 29, 29, 29, 29, 29, 31, 31, 31, 31, 31, 31, 33, 33, 33, 33, 33, 36, 36, 36, 36, 38, 38, 38, 40, 40, 42},
    {9, 13, 13, 15, 13, 15, 17, 17, 17, 17, 19, 19, 19, 19, 19, 21, 21, 21, 21, 21, 21, 22, 22, 22, 22, 22, 22, 22, 24, 24, 24, 24, 24, 24, 24, 24,
     25, 25, 25, 25, 25, 25, 25, 27, 27, 27, 27, 27, 27, 28, 28, 28, 28, 28, 30, 30, 30, 30, 32, 32, 32, 33, 33, 35} };
  for (int t = 0; t < 2; t++) {
    for (int k = 0; k < 16; k++) g_default4[t][kZig4[k]] = z4[t][k];
    for (int k = 0; k < 64; k++) g_default8[t][g_zig8[k]] = z8[t][k];
  }
  return true;
}
static const bool g_scaling_tables_ready = init_scaling_tables();

// one scaling_list(): returns true when the list asks for the default matrix (first delta brings next_scale to 0)
static bool read_scaling_list (BitReader& br, uint8_t* list, int size, bool& bad) {
  int last = 8, next = 8;
  for (int j = 0; j < size; j++) {
    if (next != 0) {
      const int delta = br.se();
      if (delta < -128 || delta > 127) { bad = true; return false; }
      next = (last + delta + 256) % 256;
      if (j == 0 && next == 0) return true;
    }
    const int idx = size == 16 ? kZig4[j] : g_zig8[j];
    list[idx] = (uint8_t) (next == 0 ? last : next);
    last = list[idx];
  }
  return false;
}

// n_lists = 8 (4:2:0) or 12; `sps` = the sequence parameter set whose lists are fall-back rule B for a PPS (nullptr: rule A)
static bool parse_scaling_lists (BitReader& br, int n_lists, const Sps* sps, bool* present, uint8_t (*l4)[16], uint8_t (*l8)[64]) {
  (void)g_scaling_tables_ready;
  const uint8_t* fall[4] = { sps ? sps->scale4[0] : g_default4[0], sps ? sps->scale4[3] : g_default4[1],
                             sps ? sps->scale8[0] : g_default8[0], sps ? sps->scale8[1] : g_default8[1] };
  for (int i = 0; i < n_lists; i++) {
    const bool here = br.u1();
    if (present) present[i] = here;
    bool bad = false;
    if (here) {
      if (i < 6) { if (read_scaling_list (br, l4[i], 16, bad)) memcpy (l4[i], g_default4[i / 3], 16); }
      else if (read_scaling_list (br, l8[i - 6], 64, bad)) memcpy (l8[i - 6], g_default8[(i - 6) & 1], 64);
      if (bad) return false;
    } else if (i < 6) {
      if (i != 0 && i != 3) memcpy (l4[i], l4[i - 1], 16); else memcpy (l4[i], fall[i / 3], 16);
    } else {
      if (i == 6 || i == 7) memcpy (l8[i - 6], fall[(i & 1) + 2], 64); else memcpy (l8[i - 6], l8[i - 8], 64);
    }
  }
  return !br.overrun();
}

bool parse_sps (BitReader& br, Sps& s) {
  s = Sps();
  s.profile_idc = br.u (8);
  s.constraint_flags = br.u (8);
  s.level_idc = br.u (8);
  s.id = br.ue();
  if (s.id < 0 || s.id > 31) return false;
  const int p = s.profile_idc;
  if (p == 100 || p == 110 || p == 122 || p == 244 || p == 44 || p == 83 || p == 86 || p == 118 || p == 128) {
    s.chroma_format_idc = br.ue();
    if (s.chroma_format_idc == 3) br.u1();
    const int bdl = br.ue(), bdc = br.ue();
    br.u1();                                       // qpprime_y_zero_transform_bypass_flag
    if (br.u1()) {                                 // seq_scaling_matrix_present_flag
      s.scaling_matrix_present = true;
      if (!parse_scaling_lists (br, s.chroma_format_idc != 3 ? 8 : 12, nullptr, nullptr, s.scale4, s.scale8)) return false;
    }
    if (s.chroma_format_idc != 1 || bdl != 0 || bdc != 0) return false;   // 4:2:0 8-bit only
  }
  s.log2_max_frame_num = br.ue() + 4;
  s.poc_type = br.ue();
  if (s.poc_type == 0) s.log2_max_poc_lsb = br.ue() + 4;
  else if (s.poc_type == 1) {
    s.delta_pic_order_always_zero = br.u1();
    s.offset_for_non_ref_pic = br.se();
    s.offset_for_top_to_bottom = br.se();
    s.num_ref_frames_in_poc_cycle = br.ue();
    if (s.num_ref_frames_in_poc_cycle > 255) return false;
    for (int i = 0; i < s.num_ref_frames_in_poc_cycle; i++) s.offset_for_ref_frame[i] = br.se();
  } else if (s.poc_type != 2) return false;
  s.max_num_ref_frames = br.ue();
  s.gaps_in_frame_num_allowed = br.u1();
  s.mb_w = br.ue() + 1;
  s.mb_h = br.ue() + 1;
  s.frame_mbs_only = br.u1();
  if (!s.frame_mbs_only) { br.u1(); return false; }   // interlace is outside the reference decoder's scope too
  s.direct_8x8_inference = br.u1();
  if (br.u1()) {                                      // frame_cropping_flag (4:2:0: units of 2 samples)
    s.crop_left = 2 * br.ue(); s.crop_right = 2 * br.ue(); s.crop_top = 2 * br.ue(); s.crop_bottom = 2 * br.ue();
  }
  // vui_parameters: nothing the reconstruction path needs
  if (br.overrun() || s.mb_w <= 0 || s.mb_h <= 0 || s.mb_w > 1024 || s.mb_h > 1024 || s.max_num_ref_frames > 16) return false;
  s.valid = true;
  return true;
}

bool parse_pps (BitReader& br, Pps& p, const Sps* sps_table) {
  p = Pps();
  const uint32_t pid = br.ue(), sid = br.ue();
  if (pid > 255 || sid > 31) return false;
  p.id = (int)pid; p.sps_id = (int)sid;
  p.cabac = br.u1();
  p.bottom_field_pic_order = br.u1();
  {
    const uint32_t nsg = br.ue();                      // num_slice_groups_minus1: 0..7 (7.4.2.2; ParsePps: MAX_SLICEGROUP_IDS, au_parser.cpp:1315)
    if (nsg > 7) return false;
    p.num_slice_groups = (int)nsg + 1;
  }
  if (p.num_slice_groups > 1) {
    // The reference decodes slice_group_map_type 0 (interleaved) and 1 (dispersed) only and drops a parameter set with any other
    // type (au_parser.cpp:1322-1326); the same here, which also keeps crafted type-6 maps (2^32 map units) out of the parser.
    const uint32_t type = br.ue();
    if (type > 1) return false;
    p.slice_group_map_type = (int)type;
    if (type == 0) for (int i = 0; i < p.num_slice_groups; i++) {
        const uint32_t rl = br.ue();
        if (rl >= (1u << 20)) return false;
        p.run_length[i] = (int)rl + 1;
      }
  }
  {
    const uint32_t l0 = br.ue(), l1 = br.ue();         // ParsePps drops the set when either exceeds MAX_REF_PIC_COUNT = 16 (au_parser.cpp:1345-1348)
    if (l0 > 15 || l1 > 15) return false;
    p.num_ref_idx_l0_default = (int)l0 + 1; p.num_ref_idx_l1_default = (int)l1 + 1;
  }
  p.weighted_pred = br.u1();
  p.weighted_bipred_idc = br.u (2);
  if (p.weighted_bipred_idc != 0) return false;      // B-slice tool: ParsePps drops the set (au_parser.cpp:1352-1358)
  p.pic_init_qp = 26 + br.se();
  p.pic_init_qs = 26 + br.se();
  p.chroma_qp_index_offset = br.se();
  p.second_chroma_qp_index_offset = p.chroma_qp_index_offset;
  p.deblocking_filter_control_present = br.u1();
  p.constrained_intra_pred = br.u1();
  p.redundant_pic_cnt_present = br.u1();
  if (br.more_rbsp_data()) {
    p.transform_8x8_mode = br.u1();
    if (br.u1()) {                                    // pic_scaling_matrix_present_flag
      // ParsePps reads 8 lists (12 when the SPS has not arrived yet or is 4:4:4) whatever transform_8x8_mode_flag says, and
      // falls back on the lists of an SPS that carries a matrix of its own (au_parser.cpp:1384-1395, :1512-1522)
      p.scaling_matrix_present = true;
      const Sps* sps = sps_table && sps_table[p.sps_id].valid ? &sps_table[p.sps_id] : nullptr;
      p.sps_was_missing = sps == nullptr;
      if (!parse_scaling_lists (br, sps && sps->chroma_format_idc != 3 ? 8 : 12, sps && sps->scaling_matrix_present ? sps : nullptr, p.list_present, p.scale4, p.scale8)) return false;
    }
    p.second_chroma_qp_index_offset = br.se();
  }
  if (br.overrun()) return false;
  p.valid = true;
  return true;
}

int parse_slice_header (BitReader& br, int nal_ref_idc, int nal_type, const Sps* sps_table, const Pps* pps_table, SliceHeader& sh) {
  sh = SliceHeader();
  sh.nal_ref_idc = nal_ref_idc; sh.nal_type = nal_type; sh.idr = nal_type == NAL_IDR;
  const uint32_t first_mb = br.ue();                   // unsigned: 32 leading zero bits give 0xffffffff, which must not become -1
  const uint32_t st = br.ue();
  if (st > 9) return -1;
  sh.slice_type = st % 5;
  const uint32_t pps_id = br.ue();
  if (pps_id > 255 || !pps_table[pps_id].valid) return -2;
  sh.pps_id = (int)pps_id;
  const Pps& pps = pps_table[sh.pps_id];
  if (!sps_table[pps.sps_id].valid) return -2;
  const Sps& sps = sps_table[pps.sps_id];
  if (first_mb >= (uint32_t) (sps.mb_w * sps.mb_h)) return -1;   // 7.4.3: first_mb_in_slice < PicSizeInMbs
  sh.first_mb = (int)first_mb;
  if (sh.slice_type != SLICE_P && sh.slice_type != SLICE_I) return -3;     // B/SP/SI: not in the reference decoder either
  sh.frame_num = br.u (sps.log2_max_frame_num);
  if (sh.idr) { const uint32_t v = br.ue(); if (v > 65535) return -1; sh.idr_pic_id = (int)v; }   // 7.4.3: 0..65535
  if (sps.poc_type == 0) {
    sh.poc_lsb = br.u (sps.log2_max_poc_lsb);
    if (pps.bottom_field_pic_order) sh.delta_poc_bottom = br.se();
  } else if (sps.poc_type == 1 && !sps.delta_pic_order_always_zero) {
    sh.delta_poc[0] = br.se();
    if (pps.bottom_field_pic_order) sh.delta_poc[1] = br.se();
  }
  if (pps.redundant_pic_cnt_present) { const uint32_t v = br.ue(); if (v > 127) return -1; sh.redundant_pic_cnt = (int)v; }   // 0..127
  sh.num_ref_idx_l0 = pps.num_ref_idx_l0_default;
  if (sh.slice_type == SLICE_P) {
    if (br.u1()) { const uint32_t v = br.ue(); if (v > 31) return -1; sh.num_ref_idx_l0 = (int)v + 1; }   // num_ref_idx_active_override_flag
    // frame-only decoding: 7.4.3 limits num_ref_idx_l0_active_minus1 to 0..15; lh264_slice_rec carries LH264_MAX_REFS = 16 entries
    if (sh.num_ref_idx_l0 < 1 || sh.num_ref_idx_l0 > 16) return -1;
    if (br.u1()) {                                    // ref_pic_list_modification_flag_l0
      for (;;) {
        const uint32_t op = br.ue();
        if (op == 3) break;
        if (op > 3 || sh.mods.size() > 64) return -1;
        sh.mods.push_back (RefPicListMod{(int)op, br.ue()});
      }
    }
    if (pps.weighted_pred) {
      sh.has_wp = true;
      const uint32_t ld = br.ue(), cd = br.ue();
      if (ld > 7 || cd > 7) return -1;                // 7.4.3.2: 0..7
      sh.luma_log2_denom = (int)ld; sh.chroma_log2_denom = (int)cd;
      for (int i = 0; i < sh.num_ref_idx_l0; i++) {
        sh.luma_weight[i] = 1 << sh.luma_log2_denom; sh.luma_offset[i] = 0;
        for (int c = 0; c < 2; c++) { sh.chroma_weight[i][c] = 1 << sh.chroma_log2_denom; sh.chroma_offset[i][c] = 0; }
        if (br.u1()) { sh.luma_weight[i] = br.se(); sh.luma_offset[i] = br.se(); }
        if (br.u1()) for (int c = 0; c < 2; c++) { sh.chroma_weight[i][c] = br.se(); sh.chroma_offset[i][c] = br.se(); }
      }
    }
  } else sh.num_ref_idx_l0 = 0;
  if (nal_ref_idc != 0) {                             // dec_ref_pic_marking()
    if (sh.idr) { sh.no_output_of_prior_pics = br.u1(); sh.long_term_reference_flag = br.u1(); }
    else if ((sh.adaptive_marking = br.u1())) {
      for (;;) {
        const uint32_t op = br.ue();
        if (op == 0) break;
        if (op > 6 || sh.mmcos.size() > 64) return -1;
        Mmco m{(int)op, 0, 0};
        if (op == 1 || op == 3) m.a = br.ue();        // difference_of_pic_nums_minus1
        if (op == 2) m.a = br.ue();                   // long_term_pic_num
        if (op == 3 || op == 6) m.b = br.ue();        // long_term_frame_idx
        if (op == 4) m.a = br.ue();                   // max_long_term_frame_idx_plus1
        sh.mmcos.push_back (m);
      }
    }
  }
  if (pps.cabac && sh.slice_type != SLICE_I) { const uint32_t idc = br.ue(); if (idc > 2) return -1; sh.cabac_init_idc = (int)idc; }   // 7.4.3: 0..2
  sh.slice_qp = pps.pic_init_qp + br.se();
  if (sh.slice_qp < 0 || sh.slice_qp > 51) return -1;
  if (pps.deblocking_filter_control_present) {
    const uint32_t idc = br.ue();
    if (idc > 2) return -1;
    sh.disable_deblocking_idc = (int)idc;
    if (sh.disable_deblocking_idc != 1) { sh.alpha_c0_offset = 2 * br.se(); sh.beta_offset = 2 * br.se(); }
  }
  if (br.overrun()) return -1;
  sh.header_bits = br.pos();
  return 0;
}

} // namespace lh264host
