// cuda_emu.h — TEST INFRASTRUCTURE ONLY: a CPU emulation of the slice of the CUDA execution model that this repository's
// kernels use, so that the kernel SOURCES (losslessh264_b200/csrc/*.cuh, unchanged) can be compiled with g++ and checked against
// the golden SHA1s in the build container, which has no GPU.  Never part of the product: the product has no CPU path.
//
// Model: every CUDA thread is a fiber (own stack, hand-written x86-64 context switch) inside ONE OS thread, scheduled round
// robin; __syncwarp / __syncthreads / shuffles / ballots are barriers over the fibers of a warp / CTA; atomics and acquire /
// release accesses are plain accesses (no real concurrency); a spinning thread yields in __nanosleep (or LH264_SPIN_HINT).
// CTAs of a launch run one after the other (so `__shared__` can be a function-local static); persistent kernels whose CTAs wait
// on one another are launched with a single CTA of several warps, which exercises the same inter-warp protocol.
// What this does NOT check: memory-model races, bank conflicts, alignment faults — the GPU test suite stays the judge of those.
#pragma once
#define LH264_CPU_EMU 1
#include <stdint.h>
#include <stddef.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <functional>
#include <vector>
#include <sys/mman.h>

#define __device__
#define __global__
#define __host__
#define __forceinline__ inline __attribute__ ((always_inline))
#define __constant__ static const
#define __shared__ static
#define __launch_bounds__(...)
#define __align__(n) __attribute__ ((aligned (n)))

struct uint3_emu { unsigned x, y, z; };
struct uchar4 { uint8_t x, y, z, w; };
struct __attribute__ ((aligned (8))) uint2 { uint32_t x, y; };
struct __attribute__ ((aligned (16))) uint4 { uint32_t x, y, z, w; };
struct __attribute__ ((aligned (8))) int2 { int x, y; };
struct __attribute__ ((aligned (16))) int4 { int x, y, z, w; };
static inline uchar4 make_uchar4 (uint8_t x, uint8_t y, uint8_t z, uint8_t w) { return uchar4{x, y, z, w}; }
static inline uint2 make_uint2 (uint32_t x, uint32_t y) { return uint2{x, y}; }
static inline uint4 make_uint4 (uint32_t x, uint32_t y, uint32_t z, uint32_t w) { return uint4{x, y, z, w}; }
static inline int2 make_int2 (int x, int y) { return int2{x, y}; }

namespace emu {

struct Barrier { unsigned count = 0, gen = 0; };
struct WarpState { Barrier bar; uint64_t xch[32]; };
struct Fiber {
  void* sp = nullptr; char* stack = nullptr;
  uint3_emu tid{0, 0, 0};
  int lane = 0;
  WarpState* ws = nullptr;
  bool done = false;
  Barrier* wait_bar = nullptr; unsigned wait_gen = 0;
};
struct Launch {
  uint3_emu grid{1, 1, 1}, block{1, 1, 1}, bid{0, 0, 0};
  Barrier cta_bar;
  unsigned n_threads = 0;
  std::vector<Fiber> fibers;
  std::vector<WarpState> warps;
  std::function<void()> body;
  void* sched_sp = nullptr;
  unsigned long long progress = 0, spins = 0;
};
extern Fiber* cur;
extern Launch* L;
extern "C" void emu_switch (void** save_sp, void* load_sp);
void yield();
unsigned char* dyn_shared();           // the launch's dynamic shared memory (extern __shared__)
void barrier_wait (Barrier& b, unsigned expected);
// runs `body` once per CUDA thread of a grid x block launch (CTAs sequentially)
void launch (unsigned grid_x, unsigned grid_y, unsigned block_x, std::function<void()> body, size_t dyn_smem_bytes = 0);

template <class T> static inline T xchg (unsigned mask, T v, int src) {
  WarpState& w = *cur->ws;
  uint64_t t = 0;
  memcpy (&t, &v, sizeof (T));
  w.xch[cur->lane] = t;
  barrier_wait (w.bar, (unsigned)__builtin_popcount (mask));
  T r;
  memcpy (&r, &w.xch[src & 31], sizeof (T));
  barrier_wait (w.bar, (unsigned)__builtin_popcount (mask));
  return r;
}

} // namespace emu

#define threadIdx (emu::cur->tid)
#define blockIdx (emu::L->bid)
#define blockDim (emu::L->block)
#define gridDim (emu::L->grid)
#define LH264_SPIN_HINT() emu::yield()

// ---- warp / CTA primitives ------------------------------------------------------------------------------------------
static inline void __syncwarp (unsigned mask = 0xffffffffu) { emu::barrier_wait (emu::cur->ws->bar, (unsigned)__builtin_popcount (mask)); }
static inline void __syncthreads() { emu::barrier_wait (emu::L->cta_bar, emu::L->n_threads); }
template <class T> static inline T __shfl_sync (unsigned mask, T v, int src) { return emu::xchg (mask, v, src); }
template <class T> static inline T __shfl_xor_sync (unsigned mask, T v, int x) { return emu::xchg (mask, v, emu::cur->lane ^ x); }
template <class T> static inline T __shfl_up_sync (unsigned mask, T v, unsigned d) { const int s = emu::cur->lane - (int)d; return emu::xchg (mask, v, s < 0 ? emu::cur->lane : s); }
template <class T> static inline T __shfl_down_sync (unsigned mask, T v, unsigned d) { const int s = emu::cur->lane + (int)d; return emu::xchg (mask, v, s > 31 ? emu::cur->lane : s); }
static inline unsigned __ballot_sync (unsigned mask, int pred) {
  emu::WarpState& w = *emu::cur->ws;
  w.xch[emu::cur->lane] = pred ? 1 : 0;
  emu::barrier_wait (w.bar, (unsigned)__builtin_popcount (mask));
  unsigned r = 0;
  for (int i = 0; i < 32; i++) if (((mask >> i) & 1u) && w.xch[i]) r |= 1u << i;
  emu::barrier_wait (w.bar, (unsigned)__builtin_popcount (mask));
  return r;
}
static inline int __any_sync (unsigned mask, int pred) { return __ballot_sync (mask, pred) != 0; }
static inline int __all_sync (unsigned mask, int pred) { return __ballot_sync (mask, pred) == mask; }
static inline unsigned __match_any_sync (unsigned mask, uint32_t v) {
  emu::WarpState& w = *emu::cur->ws;
  w.xch[emu::cur->lane] = v;
  emu::barrier_wait (w.bar, (unsigned)__builtin_popcount (mask));
  unsigned r = 0;
  for (int i = 0; i < 32; i++) if (((mask >> i) & 1u) && (uint32_t)w.xch[i] == v) r |= 1u << i;
  emu::barrier_wait (w.bar, (unsigned)__builtin_popcount (mask));
  return r;
}
static inline void __nanosleep (unsigned) { emu::yield(); }
static inline void __threadfence() {}
static inline void __threadfence_block() {}

// ---- memory / atomics (single OS thread: plain accesses) ------------------------------------------------------------
template <class T> static inline T __ldg (const T* p) { return *p; }
template <class T> static inline T __ldcg (const T* p) { return *p; }
template <class T> static inline void __stcg (T* p, T v) { *p = v; }
static inline unsigned atomicAdd (unsigned* p, unsigned v) { const unsigned o = *p; *p = o + v; return o; }
static inline int atomicAdd (int* p, int v) { const int o = *p; *p = o + v; return o; }
static inline unsigned long long atomicAdd (unsigned long long* p, unsigned long long v) { const unsigned long long o = *p; *p = o + v; return o; }
static inline unsigned atomicOr (unsigned* p, unsigned v) { const unsigned o = *p; *p = o | v; return o; }
static inline int atomicOr (int* p, int v) { const int o = *p; *p = o | v; return o; }
static inline unsigned atomicMax (unsigned* p, unsigned v) { const unsigned o = *p; if (v > o) *p = v; return o; }
static inline size_t __cvta_generic_to_shared (const void* p) { return (size_t)p; }

// ---- integer intrinsics -----------------------------------------------------------------------------------------------
static inline int __popc (unsigned v) { return __builtin_popcount (v); }
static inline int __clz (int v) { return v == 0 ? 32 : __builtin_clz ((unsigned)v); }
static inline int __ffs (int v) { return __builtin_ffs (v); }
static inline unsigned __funnelshift_r (unsigned lo, unsigned hi, unsigned sh) { return (unsigned) ((((uint64_t)hi << 32) | lo) >> (sh & 31u)); }
static inline unsigned __funnelshift_rc (unsigned lo, unsigned hi, unsigned sh) { return sh >= 32u ? hi : (unsigned) ((((uint64_t)hi << 32) | lo) >> sh); }
static inline unsigned __funnelshift_l (unsigned lo, unsigned hi, unsigned sh) { return (unsigned) (((((uint64_t)hi << 32) | lo) << (sh & 31u)) >> 32); }
static inline unsigned __byte_perm (unsigned x, unsigned y, unsigned s) {
  const uint64_t v = ((uint64_t)y << 32) | x;
  unsigned r = 0;
  for (int i = 0; i < 4; i++) {
    const unsigned sel = (s >> (4 * i)) & 0xfu;
    unsigned b = (unsigned) (v >> (8 * (sel & 7u))) & 0xffu;
    if (sel & 8u) b = (b & 0x80u) ? 0xffu : 0u;
    r |= b << (8 * i);
  }
  return r;
}
static inline float __fdividef (float a, float b) { return a / b; }
template <class T> static inline T min (T a, T b) { return b < a ? b : a; }
template <class T> static inline T max (T a, T b) { return a < b ? b : a; }
