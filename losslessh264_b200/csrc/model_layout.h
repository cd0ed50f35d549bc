// model_layout.h — layout shared by the two directions of the recompression model: compact tag ids, the pruned prior table
// and the per-macroblock context record.  Plain C++ (no CUDA): included by model_kernels.cuh (compress), model_decode_core.inl
// (decompress) and the CPU build of the latter that the tests use as checker.
#pragma once
#include <stdint.h>
#include "../../include/lh264_b200.h"

namespace lh264m {

// ---- compact tag ids (lane = tag owner in k_model_code) ---------------------------------------------------
enum {
  CT_SKIP = 0, CT_SKIP_END, CT_CBPL, CT_QPL, CT_MB_TYPE, CT_T8, CT_REF, CT_8x8, CT_16x16, CT_PRED_MODE, CT_SUB_MB,
  CT_MVX, CT_MVY, CT_LDC, CT_CRDC,
  CT_L0_EOB, CT_L0_BITMASK, CT_L0_EXP, CT_L0_RES, CT_L0_SIGN,
  CT_LN_BITMASK, CT_LN_EXP, CT_LN_RES, CT_LN_SIGN,
  CT_C_EOB, CT_C_BITMASK, CT_C_EXP, CT_C_RES, CT_C_SIGN, CT_PAD, CT_COUNT
};
static const int kTagOfCompact[CT_COUNT] = {   // billing.h:7-58
  LH264_TAG_SKIP, LH264_TAG_SKIP_END, LH264_TAG_CBPL, LH264_TAG_QPL, LH264_TAG_MB_TYPE, LH264_TAG_TRANSFORM_8x8, LH264_TAG_REF,
  LH264_TAG_8x8, LH264_TAG_16x16, LH264_TAG_PRED_MODE, LH264_TAG_SUB_MB, LH264_TAG_MVX, LH264_TAG_MVY, LH264_TAG_LDC, LH264_TAG_CRDC,
  LH264_TAG_LAC_0_EOB, LH264_TAG_LAC_0_BITMASK, LH264_TAG_LAC_0_EXP, LH264_TAG_LAC_0_RES, LH264_TAG_LAC_0_SIGN,
  LH264_TAG_LAC_N_BITMASK, LH264_TAG_LAC_N_EXP, LH264_TAG_LAC_N_RES, LH264_TAG_LAC_N_SIGN,
  LH264_TAG_CRAC_EOB, LH264_TAG_CRAC_BITMASK, LH264_TAG_CRAC_EXP, LH264_TAG_CRAC_RES, LH264_TAG_CRAC_SIGN, LH264_TAG_PADBYTE };

// ---- prior table layout (units: one 32-bit DynProb) ---------------------------------------------------------
// DynProb word: bits 0-9 counts[0], bits 10-19 counts[1], bits 24-31 prob ^ 0x80 (compression_stream.h:87-115).
constexpr uint32_t NT = 12;                                  // mb type codes 0..11 (macroblock_model.cpp:647-679)
constexpr uint32_t P_TEST = 0;                               // ArithmeticCodedOutput::TEST_PROB (shared by all tags)
constexpr uint32_t P_MBTYPE = 16;                            // [32][2][16]        mbTypePriors
constexpr uint32_t P_CBPC = P_MBTYPE + 32 * 2 * 16;          // [4][NT][4]         CbpCPrior
constexpr uint32_t P_CBPL = P_CBPC + 4 * NT * 4;             // [16][NT][16]       CbpLPrior
constexpr uint32_t P_QPL = P_CBPL + 16 * NT * 16;            // [2][3][128]        mbQPLPrior
constexpr uint32_t P_NUMREF = P_QPL + 2 * 3 * 128;           // [16][NT][16]       numRefIdxL0ActivePrior
constexpr uint32_t P_CMODE = P_NUMREF + 16 * NT * 16;        // [8][8]             chromaI8x8ModePriors
constexpr uint32_t P_PREDMODE = P_CMODE + 64;                // [NT][8][9][16]     predictionModePriors
constexpr uint32_t P_SUBMB = P_PREDMODE + NT * 8 * 9 * 16;   // [NT][256]          subMbPriors
constexpr uint32_t P_STOP = P_SUBMB + NT * 256;              // [2048]             stopBitPriors
constexpr uint32_t P_T8 = P_STOP + 2048;                     // [NT][64]           transform8x8FlagPriors
constexpr uint32_t P_LDC = P_T8 + NT * 64;                   // [16][2][NT][16]    lumaDCIntPriors   (IntPrior<3,4>)
constexpr uint32_t P_CDC = P_LDC + 16 * 2 * NT * 16;         // [8][2][NT][16]     chromaDCIntPriors
constexpr uint32_t P_MVD = P_CDC + 8 * 2 * NT * 16;          // [5][16][16]        motionVectorDifferencePriors (types 3..7)
constexpr uint32_t P_NZ = P_MVD + 5 * 16 * 16;               // [2][NT][3][27][8]  nonzerosPriors    (UnsignedIntPrior<3,4>)
constexpr uint32_t P_NZ8 = P_NZ + 2 * NT * 3 * 27 * 8;       // [2][NT][27][8]     nonzerosPriors8x8
constexpr uint32_t P_SKIP = P_NZ8 + 2 * NT * 27 * 8;         // [512][512]         mbSkipRunPrior (type code 11 only)
constexpr uint32_t P_AC = P_SKIP + 512 * 512;                // [2][8][3][16][125][16]  acPriors (UEGkIntPrior<14,4,2,4,0>)
constexpr uint32_t P_AC8 = P_AC + 2 * 8 * 3 * 16 * 125 * 16; // [2][8][64][125][16]     acPriors8x8
constexpr uint32_t P_END_NO8 = P_AC8;
constexpr uint32_t P_END = P_AC8 + 2 * 8 * 64 * 125 * 16;
// slots inside the integer priors
enum { I_EXP = 0, I_MANT = 3, I_ZERO = 7, I_SIGN = 8 };                      // UnsignedIntPrior<3,4> / IntPrior<3,4>
enum { U_SIGN = 0, U_FIRST = 1, U_EXP = 5, U_ZERO = 13 };                    // UEGkIntPrior<.,4,.,4,.>: mantissa follows exponent
#define NO_PRIOR 0xffffffffu

struct alignas (16) CtxRec {      // what the priors read from a DecodedMacroblock (decoded_macroblock.h:4-104), 32 bytes
  uint8_t flags;                    // bit0 odata.initialized, bit1 isSkipped
  uint8_t type_code, cbp, num_ref, chroma_mode, i16_mode;
  uint16_t cachedSkips;
  uint8_t cnt[24];                  // non-zero levels per 16-coefficient chunk (16 luma, 4 Cb, 4 Cr), DC slot included
};
static_assert (sizeof (CtxRec) == 32, "CtxRec");

} // namespace lh264m
