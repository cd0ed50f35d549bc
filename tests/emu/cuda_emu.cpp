// cuda_emu.cpp — runtime of tests/emu/cuda_emu.h (TEST INFRASTRUCTURE ONLY): fibers, barriers, launch.
#include "cuda_emu.h"

namespace emu {

Fiber* cur = nullptr;
Launch* L = nullptr;

// x86-64 System V: callee-saved registers + stack pointer.  A new fiber's stack is primed so that the first switch "returns"
// into fiber_entry.
asm (R"(
.text
.globl emu_switch
.type emu_switch,@function
emu_switch:
  pushq %rbp
  pushq %rbx
  pushq %r12
  pushq %r13
  pushq %r14
  pushq %r15
  movq %rsp, (%rdi)
  movq %rsi, %rsp
  popq %r15
  popq %r14
  popq %r13
  popq %r12
  popq %rbx
  popq %rbp
  ret
.size emu_switch,.-emu_switch
)");

static const size_t kStack = 512 * 1024;

static void fiber_entry() {
  L->body();
  cur->done = true;
  L->progress++;
  for (;;) emu_switch (&cur->sp, L->sched_sp);
}

void yield() {
  L->spins++;
  emu_switch (&cur->sp, L->sched_sp);
}

void barrier_wait (Barrier& b, unsigned expected) {
  L->progress++;
  if (++b.count >= expected) { b.count = 0; b.gen++; return; }
  cur->wait_bar = &b; cur->wait_gen = b.gen;
  emu_switch (&cur->sp, L->sched_sp);
}

static void prime (Fiber& f) {
  if (!f.stack) {
    f.stack = (char*)mmap (nullptr, kStack, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
    if (f.stack == (char*)MAP_FAILED) { perror ("mmap"); abort(); }
  }
  // layout popped by emu_switch: r15 r14 r13 r12 rbx rbp, then the return address; at entry of fiber_entry rsp % 16 must be 8
  uintptr_t top = ((uintptr_t)f.stack + kStack) & ~(uintptr_t)15;
  void** sp = (void**)(top - 8);          // so that after `ret` pops the address, rsp = top - 8  (== 8 mod 16)
  *--sp = (void*)fiber_entry;
  for (int i = 0; i < 6; i++) *--sp = nullptr;
  f.sp = sp;
  f.done = false; f.wait_bar = nullptr;
}

static std::vector<unsigned char> g_dyn;
unsigned char* dyn_shared() { return g_dyn.data(); }

void launch (unsigned grid_x, unsigned grid_y, unsigned block_x, std::function<void()> body, size_t dyn_smem_bytes) {
  g_dyn.assign (dyn_smem_bytes + 64, 0);
  static Launch launch_state;            // fiber stacks are reused across launches
  Launch& l = launch_state;
  Launch* outer = L;
  if (outer) { fprintf (stderr, "emu: nested launch\n"); abort(); }
  L = &l;
  l.grid = uint3_emu{grid_x, grid_y, 1}; l.block = uint3_emu{block_x, 1, 1};
  l.n_threads = block_x;
  l.body = body;
  if (l.fibers.size() < block_x) l.fibers.resize (block_x);
  l.warps.assign ((block_x + 31) / 32, WarpState());
  for (unsigned by = 0; by < grid_y; by++) for (unsigned bx = 0; bx < grid_x; bx++) {
      l.bid = uint3_emu{bx, by, 0};
      l.cta_bar = Barrier();
      for (auto& w : l.warps) w = WarpState();
      for (unsigned t = 0; t < block_x; t++) {
        Fiber& f = l.fibers[t];
        prime (f);
        f.tid = uint3_emu{t, 0, 0}; f.lane = (int) (t & 31); f.ws = &l.warps[t >> 5];
      }
      unsigned alive = block_x;
      unsigned long long idle_rounds = 0;
      while (alive) {
        const unsigned long long before = l.progress;
        bool ran = false;
        for (unsigned t = 0; t < block_x; t++) {
          Fiber& f = l.fibers[t];
          if (f.done) continue;
          if (f.wait_bar) { if (f.wait_bar->gen == f.wait_gen) continue; f.wait_bar = nullptr; }
          cur = &f;
          ran = true;
          emu_switch (&l.sched_sp, f.sp);
          if (f.done) alive--;
        }
        cur = nullptr;
        if (!ran) { fprintf (stderr, "emu: deadlock - every live thread of CTA (%u,%u) waits at a barrier that cannot complete (divergent __syncwarp / shuffle?)\n", bx, by); abort(); }
        idle_rounds = l.progress == before ? idle_rounds + 1 : 0;
        if (idle_rounds > 2000000ull) { fprintf (stderr, "emu: live-lock - threads of CTA (%u,%u) spin without progress (a wait on a flag nobody sets?)\n", bx, by); abort(); }
      }
    }
  L = nullptr;
}

} // namespace emu
