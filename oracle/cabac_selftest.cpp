/*
 * oracle/cabac_selftest.cpp — TEST INFRASTRUCTURE ONLY (built and run by tests/test_cabac_coder.py).
 *
 * Known-answer and round-trip checks of the CABAC arithmetic coder pair, in the spirit of the reference's start-up self-test of
 * its bool coder (codec/console/dec/src/test_arithmetic_coding.cpp:18-71):
 *   encoder = the CPU build of the code k_model_decode runs (csrc/model_decode_core.inl: cb_bin / cb_bypass / cb_terminate)
 *   decoder = the host parser's engine (host/cabac.h: CabacDecoder)
 * Random sequences of context-coded, bypass and terminate bins with skewed and adversarial statistics (long runs that pile up
 * outstanding bits, carries through them) are encoded, decoded again, and must come back bin for bin with the decoder's bit
 * position landing exactly behind the flush's last bit.
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include "../include/lh264_b200.h"
#include "../losslessh264_b200/csrc/model_layout.h"
#include "../losslessh264_b200/host/cavlc_tables.h"
#include "../losslessh264_b200/host/cabac.h"
#define MD_FN static inline
#define MD_FN_NOINLINE static
#define MD_TABLE static const
#define MD_CLZ(x) __builtin_clz (x)
#define MD_DIV(n, d) ((n) / (d))
#define MD_READER(c, tag) ((c).rd[tag])
#define MD_TEST(c) ((c).test)
#include "../losslessh264_b200/csrc/model_decode_core.inl"

using namespace lh264md;

static uint64_t rng_state = 0x9E3779B97F4A7C15ull;
static uint32_t rnd() { rng_state ^= rng_state << 13; rng_state ^= rng_state >> 7; rng_state ^= rng_state << 17; return (uint32_t) (rng_state >> 11); }

struct Bin { uint8_t kind; uint16_t ctx; uint8_t val; };   // kind 0 decision, 1 bypass, 2 terminate(0)

static int run_case (int n_bins, int mode, int cls, int qp, int header_bits) {
  std::vector<Bin> bins (n_bins);
  for (int i = 0; i < n_bins; i++) {
    Bin b;
    const uint32_t r = rnd();
    b.kind = (uint8_t) (mode == 3 ? 1 : (r % 16 == 0 ? 1 : (r % 61 == 0 ? 2 : 0)));
    b.ctx = (uint16_t) ((r >> 8) % (mode == 2 ? 3 : 460));
    if (b.kind == 2) b.val = 0;
    else if (mode == 0) b.val = (r >> 20) & 1;                        // fair
    else if (mode == 1) b.val = ((r >> 20) % 64) == 0;                // heavily skewed: long MPS runs, range stays small
    else if (mode == 2) b.val = ((r >> 20) % 7) != 0;                 // few contexts, skewed the other way
    else b.val = (i / 97) & 1 ? 1 : ((r >> 20) & 1);                  // bypass only, with runs of ones (carry chains)
    bins[i] = b;
  }
  std::vector<uint8_t> out ((size_t)n_bins / 2 + 64 + 16, 0);
  MdBits bw;
  mb_init (bw, out.data(), out.size() - 16);
  static MdScratch sc;
  mb_put (bw, 0x2aaaaa, header_bits);                                 // pretend slice header
  cb_slice_begin (bw, sc, cls, qp, (uint32_t)header_bits);
  for (const Bin& b : bins) {
    if (b.kind == 0) cb_bin (bw, sc, b.ctx, b.val);
    else if (b.kind == 1) cb_bypass (bw, sc, b.val);
    else cb_terminate (bw, sc, 0);
  }
  cb_terminate (bw, sc, 1);
  const uint64_t total = mb_bits (bw);
  mb_flush (bw);
  if (bw.overflow) { fprintf (stderr, "encoder overflow\n"); return 1; }
  lh264host::BitReader br (out.data(), (size_t) ((total + 7) / 8));
  br.skip (header_bits);
  br.align();
  lh264host::CabacDecoder dec;
  dec.init_contexts (cls, qp);
  dec.start (&br);
  for (int i = 0; i < n_bins; i++) {
    const Bin& b = bins[i];
    const int got = b.kind == 0 ? dec.decision (b.ctx) : (b.kind == 1 ? dec.bypass() : dec.terminate());
    if (got != b.val) { fprintf (stderr, "mode %d: bin %d of %d (kind %d ctx %d) decoded %d, coded %d\n", mode, i, n_bins, b.kind, b.ctx, got, b.val); return 1; }
  }
  if (dec.terminate() != 1) { fprintf (stderr, "mode %d: final terminate bin not seen\n", mode); return 1; }
  if (br.pos() != total) { fprintf (stderr, "mode %d: decoder stands at bit %zu, encoder wrote %llu\n", mode, br.pos(), (unsigned long long)total); return 1; }
  if (!((out[(total - 1) >> 3] >> (7 - ((total - 1) & 7))) & 1)) { fprintf (stderr, "mode %d: the flush does not end with a 1 bit\n", mode); return 1; }
  return 0;
}

int main() {
  int cases = 0;
  for (int mode = 0; mode < 4; mode++)
    for (int rep = 0; rep < 60; rep++) {
      const int n = rep < 8 ? rep : 1 + (int) (rnd() % 20000);
      if (run_case (n, mode, (int) (rnd() % 4), (int) (rnd() % 52), (int) (rnd() % 24))) return 1;
      cases++;
    }
  printf ("cabac selftest ok: %d cases\n", cases);
  return 0;
}
