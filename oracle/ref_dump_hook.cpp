// oracle/ref_dump_hook.cpp — TEST INFRASTRUCTURE ONLY (never linked into the product).
//
// Builds into oracle/_ref/h264dec_dump: the UNMODIFIED reference decoder with one translation
// unit swapped.  This file #includes the reference's decode_slice.cpp from where it lies under
// /root/reference (no reference source is copied into this repository) with
// WelsTargetSliceConstruction renamed, and supplies a wrapper that
//   1. converts the reference's hand-off record (SDqLayer per-MB SoA, dec_frame.h:61-127) of
//      the slice about to be reconstructed into the lh264_mb_rec / lh264_slice_rec layout of
//      include/lh264_b200.h and appends it to the file named by $LH264_DUMP,
//   2. calls the real WelsTargetSliceConstruction (decode_slice.cpp:110-206),
//   3. appends the picture planes before and after that slice's deblocking pass.
// The dumps are the golden vectors for the hot-path boundary: tests/golden/*.lhfb are produced
// by tests/golden/make_fixtures.py from this binary.
//
// File format (little endian), a sequence of slice records:
//   u32 magic 'LHFB'  u32 version=1
//   i32 mb_w, mb_h, cur_slot, frame_num, poc_lsb, idr_pic_id, nal_ref_idc, pic_serial
//   lh264_slice_rec
//   i32 n_mbs; n_mbs x { i32 mb_xy; lh264_mb_rec (slice_idx = 0, coef_off relative to slice) }
//   i32 n_coef_blocks; n x int16[16] quantised levels; n x int16[16] reference-dequantised
//   i32 has_pix; if set: pre-deblock Y,U,V planes then post-deblock Y,U,V planes
//                (mb_w*16 x mb_h*16 and two mb_w*8 x mb_h*8, tightly packed)

#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <map>
#include <vector>

#include "../include/lh264_b200.h"

// Rename the definition (and its prototype in decode_slice.h) inside the included TU.
#define WelsTargetSliceConstruction WelsTargetSliceConstruction_reference
#define WelsDeblockingFilterSlice WelsDeblockingFilterSlice_hooked
#define LH264_STR2(x) #x
#define LH264_STR(x) LH264_STR2(x)
#include LH264_STR(REF_ROOT/codec/decoder/core/src/decode_slice.cpp)
#undef WelsTargetSliceConstruction
#undef WelsDeblockingFilterSlice

namespace WelsDec {

// the real deblocking pass lives in deblocking.cpp under its original name
void WelsDeblockingFilterSlice (PWelsDecoderContext pCtx, PDeblockingFilterMbFunc pDeblockMb);

namespace {
FILE* g_dump = NULL;
bool g_dump_pix = false;
bool g_init = false;
int g_pic_serial = -1;
std::map<void*, int> g_slots;
std::vector<uint8_t> g_pre;   // pre-deblock snapshot of the current slice call

void dump_init() {
  if (g_init) return;
  g_init = true;
  const char* path = getenv ("LH264_DUMP");
  if (path && *path) g_dump = fopen (path, "wb");
  const char* pix = getenv ("LH264_DUMP_PIX");
  g_dump_pix = pix && *pix == '1';
}
int slot_of (PPicture p) {
  if (!p) return -1;
  std::map<void*, int>::iterator it = g_slots.find ((void*)p);
  if (it != g_slots.end()) return it->second;
  int id = (int)g_slots.size();
  g_slots[(void*)p] = id;
  return id;
}
void w32 (int32_t v) { fwrite (&v, 4, 1, g_dump); }

void snapshot (PPicture pic, int mb_w, int mb_h, std::vector<uint8_t>& out) {
  int w = mb_w * 16, h = mb_h * 16;
  out.resize ((size_t)w * h * 3 / 2);
  uint8_t* o = out.data();
  for (int y = 0; y < h; y++, o += w) memcpy (o, pic->pData[0] + (size_t)y * pic->iLinesize[0], w);
  for (int p = 1; p < 3; p++)
    for (int y = 0; y < h / 2; y++, o += w / 2) memcpy (o, pic->pData[p] + (size_t)y * pic->iLinesize[1], w / 2);
}

int map_mb_type (int t) {
  switch (t) {
  case MB_TYPE_INTRA4x4: return LH264_MB_I4x4;
  case MB_TYPE_INTRA8x8: return LH264_MB_I8x8;
  case MB_TYPE_INTRA16x16: return LH264_MB_I16x16;
  case MB_TYPE_INTRA_PCM: return LH264_MB_IPCM;
  case MB_TYPE_16x16: return LH264_MB_P16x16;
  case MB_TYPE_16x8: return LH264_MB_P16x8;
  case MB_TYPE_8x16: return LH264_MB_P8x16;
  case MB_TYPE_8x8: case MB_TYPE_8x8_REF0: return LH264_MB_P8x8;
  case MB_TYPE_SKIP: return LH264_MB_PSKIP;
  default: return LH264_MB_NONE;
  }
}
int map_sub_type (int t) {
  switch (t) {
  case SUB_MB_TYPE_8x4: return LH264_SUB_8x4;
  case SUB_MB_TYPE_4x8: return LH264_SUB_4x8;
  case SUB_MB_TYPE_4x4: return LH264_SUB_4x4;
  default: return LH264_SUB_8x8;
  }
}

void dump_slice_pre (PWelsDecoderContext pCtx) {
  PDqLayer L = pCtx->pCurDqLayer;
  PSlice S = &L->sLayerInfo.sSliceInLayer;
  PSliceHeader H = &S->sSliceHeaderExt.sSliceHeader;
  const int mb_w = L->iMbWidth, mb_h = L->iMbHeight;
  if (H->iFirstMbInSlice == 0 || g_pic_serial < 0) g_pic_serial++;

  w32 (0x4246484c); w32 (1);
  w32 (mb_w); w32 (mb_h); w32 (slot_of (pCtx->pDec)); w32 (H->iFrameNum); w32 (H->iPicOrderCntLsb);
  w32 (H->uiIdrPicId); w32 (pCtx->pCurDqLayer->sLayerInfo.sNalHeaderExt.sNalUnitHeader.uiNalRefIdc); w32 (g_pic_serial);

  lh264_slice_rec sr;
  memset (&sr, 0, sizeof (sr));
  sr.first_mb = H->iFirstMbInSlice;
  sr.num_mbs = S->iTotalMbInCurSlice;
  sr.slice_type = (S->eSliceType == I_SLICE) ? 2 : 0;
  sr.disable_deblocking_idc = (uint8_t)H->uiDisableDeblockingFilterIdc;
  sr.alpha_c0_offset = (int8_t)H->iSliceAlphaC0Offset;
  sr.beta_offset = (int8_t)H->iSliceBetaOffset;
  sr.num_ref_idx_l0 = (uint8_t)H->uiRefCount[0];
  sr.weighted_pred = L->bUseWeightPredictionFlag ? 1 : 0;
  if (L->pPredWeightTable) {
    sr.luma_log2_denom = (uint8_t)L->pPredWeightTable->uiLumaLog2WeightDenom;
    sr.chroma_log2_denom = (uint8_t)L->pPredWeightTable->uiChromaLog2WeightDenom;
  }
  int fallback = slot_of (pCtx->sRefPic.pRefList[LIST_0][0]);
  for (int i = 0; i < LH264_MAX_REFS; i++) {
    PPicture r = pCtx->sRefPic.pRefList[LIST_0][i];
    sr.ref_slot[i] = (int16_t) (r ? slot_of (r) : fallback);
    if (L->pPredWeightTable) {
      sr.luma_weight[i] = (int16_t)L->pPredWeightTable->sPredList[LIST_0].iLumaWeight[i];
      sr.luma_offset[i] = (int16_t)L->pPredWeightTable->sPredList[LIST_0].iLumaOffset[i];
      for (int c = 0; c < 2; c++) {
        sr.chroma_weight[i][c] = (int16_t)L->pPredWeightTable->sPredList[LIST_0].iChromaWeight[i][c];
        sr.chroma_offset[i][c] = (int16_t)L->pPredWeightTable->sPredList[LIST_0].iChromaOffset[i][c];
      }
    }
  }
  fwrite (&sr, sizeof (sr), 1, g_dump);

  // walk the slice exactly like WelsTargetSliceConstruction does (decode_slice.cpp:136-183)
  std::vector<int32_t> mbxy;
  {
    int total = H->pSps->uiTotalMbCount, next = H->iFirstMbInSlice, cnt = 0;
    while (cnt < S->iTotalMbInCurSlice) {
      mbxy.push_back (next);
      ++cnt;
      if (H->pPps->uiNumSliceGroups > 1) next = FmoNextMb (pCtx->pFmo, next); else ++next;
      if (next == -1 || next >= total) break;
    }
  }
  std::vector<int16_t> quant, scaled;
  w32 ((int32_t)mbxy.size());
  static const int kNzcChroma[8] = {16, 17, 20, 21, 18, 19, 22, 23};
  for (size_t n = 0; n < mbxy.size(); n++) {
    const int xy = mbxy[n];
    lh264_mb_rec m;
    memset (&m, 0, sizeof (m));
    const int t = L->pMbType[xy];
    m.mb_type = (uint8_t)map_mb_type (t);
    m.flags = L->pTransformSize8x8Flag[xy] ? LH264_MBF_TRANSFORM8x8 : 0;
    m.cbp = (uint8_t)L->pCbp[xy];
    m.qp_y = (uint8_t)L->pLumaQp[xy];
    m.qp_c[0] = (uint8_t)L->pChromaQp[xy][0];
    m.qp_c[1] = (uint8_t)L->pChromaQp[xy][1];
    m.i16_mode = (uint8_t)L->pIntraPredMode[xy][7];
    m.chroma_mode = (uint8_t)L->pChromaPredMode[xy];
    m.nxn_avail = L->pIntraNxNAvailFlag[xy];
    for (int r = 0; r < 16; r++) if (L->pNzc[xy][r]) m.nz_mask |= 1u << r;
    for (int c = 0; c < 8; c++) if (L->pNzc[xy][kNzcChroma[c]]) m.nz_mask |= 1u << (16 + c);
    for (int q = 0; q < 4; q++) {
      m.ref_idx[q] = L->pRefIndex[0][xy][ (q >> 1) * 8 + (q & 1) * 2];
      m.sub_mb_type[q] = (uint8_t)map_sub_type (L->pSubMbType[xy][q]);
    }
    for (int z = 0; z < 16; z++) m.intra_mode[z] = (uint8_t)L->pIntra4x4FinalMode[xy][g_kuiScan4[z]];
    for (int r = 0; r < 16; r++) { m.mv[r][0] = L->pMv[0][xy][r][0]; m.mv[r][1] = L->pMv[0][xy][r][1]; }
    m.coef_off = (uint32_t) (quant.size() / 16);
    if (m.mb_type == LH264_MB_IPCM) {
      // samples were written into the picture during parse (decode_slice.cpp:3213-3263)
      m.coded_mask = 0xFFFFFF;
      PPicture pic = pCtx->pDec;
      int mx = xy % mb_w, my = xy / mb_w;
      for (int y = 0; y < 16; y++) for (int x = 0; x < 16; x++) {
        int16_t v = pic->pData[0][ (size_t) (my * 16 + y) * pic->iLinesize[0] + mx * 16 + x];
        quant.push_back (v); scaled.push_back (v);
      }
      for (int p = 1; p < 3; p++) for (int y = 0; y < 8; y++) for (int x = 0; x < 8; x++) {
        int16_t v = pic->pData[p][ (size_t) (my * 8 + y) * pic->iLinesize[1] + mx * 8 + x];
        quant.push_back (v); scaled.push_back (v);
      }
    } else if (m.mb_type != LH264_MB_NONE && ! (IS_INTER (t) && m.cbp == 0)) {
      const int16_t* q = L->pScaledTCoeffQuant[xy];
      const int16_t* s = L->pScaledTCoeff[xy];
      for (int k = 0; k < 24; k++) {
        bool any = false;
        if (k < 16 && (m.flags & LH264_MBF_TRANSFORM8x8)) {
          for (int i = 0; i < 64; i++) any |= q[ (k & ~3) * 16 + i] != 0;
        } else {
          for (int i = 0; i < 16; i++) any |= q[k * 16 + i] != 0;
        }
        if (any) {
          m.coded_mask |= 1u << k;
          quant.insert (quant.end(), q + k * 16, q + k * 16 + 16);
          scaled.insert (scaled.end(), s + k * 16, s + k * 16 + 16);
        }
      }
    }
    w32 (xy);
    fwrite (&m, sizeof (m), 1, g_dump);
  }
  w32 ((int32_t) (quant.size() / 16));
  if (!quant.empty()) {
    fwrite (quant.data(), 2, quant.size(), g_dump);
    fwrite (scaled.data(), 2, scaled.size(), g_dump);
  }
}

} // namespace

void WelsDeblockingFilterSlice_hooked (PWelsDecoderContext pCtx, PDeblockingFilterMbFunc pDeblockMb) {
  if (g_dump && g_dump_pix) snapshot (pCtx->pDec, pCtx->pCurDqLayer->iMbWidth, pCtx->pCurDqLayer->iMbHeight, g_pre);
  WelsDeblockingFilterSlice (pCtx, pDeblockMb);
}

int32_t WelsTargetSliceConstruction (PWelsDecoderContext pCtx) {
  dump_init();
  if (!g_dump) return WelsTargetSliceConstruction_reference (pCtx);
  dump_slice_pre (pCtx);
  g_pre.clear();
  int32_t rc = WelsTargetSliceConstruction_reference (pCtx);
  if (g_dump_pix) {
    std::vector<uint8_t> post;
    snapshot (pCtx->pDec, pCtx->pCurDqLayer->iMbWidth, pCtx->pCurDqLayer->iMbHeight, post);
    if (g_pre.empty()) g_pre = post;      // deblocking disabled for this slice
    w32 (1);
    fwrite (g_pre.data(), 1, g_pre.size(), g_dump);
    fwrite (post.data(), 1, post.size(), g_dump);
  } else {
    w32 (0);
  }
  fflush (g_dump);
  return rc;
}

} // namespace WelsDec
