// Lock-step warp emulator for tests: compiles warp-synchronous CUDA device code (a kernel that only
// uses 32-lane shuffles with the full mask, __syncwarp, shared memory and global atomics) for the
// CPU, so that a kernel variant can be checked against the oracle where no GPU is available.
//
// One warp = 32 fibers (ucontext) on one OS thread, resumed round-robin.  A warp-wide (or, with a partial
// mask, group-wide) primitive is a barrier among the lanes of its mask: a lane that reaches it yields until
// the others have arrived, so lanes interleave only at
// the points where the hardware would also let them observe each other, and code that lane 0 runs
// alone between two __syncwarp() calls runs to completion while the others wait, as on the device.
// Memory is sequentially consistent here, so this checks the algorithm (indexing, tie-breaks, which
// lane holds what), not fences or races.  Test infrastructure only.
#pragma once
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

// Fiber switch: on x86-64 a dozen instructions (callee-saved registers + stack pointer); elsewhere ucontext,
// whose swapcontext makes two system calls per switch.
#if defined(__x86_64__) && !defined(EMU_USE_UCONTEXT)
#define EMU_FAST_SWITCH 1
extern "C" void emu_switch(void** save_sp, void* const* load_sp);
asm(".text\n"
    ".p2align 4\n"
    ".type emu_switch,@function\n"
    "emu_switch:\n"
    "  pushq %rbp\n  pushq %rbx\n  pushq %r12\n  pushq %r13\n  pushq %r14\n  pushq %r15\n"
    "  movq %rsp, (%rdi)\n"
    "  movq (%rsi), %rsp\n"
    "  popq %r15\n  popq %r14\n  popq %r13\n  popq %r12\n  popq %rbx\n  popq %rbp\n"
    "  ret\n"
    ".size emu_switch,.-emu_switch\n");
#else
#include <ucontext.h>
#endif

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __shared__

struct EmuDim3 { unsigned x, y, z; };
static EmuDim3 threadIdx = {0, 0, 0}, blockIdx = {0, 0, 0}, blockDim = {32, 1, 1}, gridDim = {1, 1, 1};

namespace emu {

struct Warp {
#ifdef EMU_FAST_SWITCH
  void* sched = nullptr;   // saved stack pointers
  void* ctx[32];
#else
  ucontext_t sched;
  ucontext_t ctx[32];
#endif
  std::vector<char> stack[32];
  bool done[32];
  int cur = 0;
  // one barrier per group of lanes (a group = the set bits of the mask its members pass; keyed by the
  // lowest lane of the mask): sub-warp groups with their own control flow synchronise independently
  int arrived[32];
  unsigned long generation[32];
  uint32_t slots[2][32];
  void (*body)(void*) = nullptr;
  void* arg = nullptr;
};
static Warp* g_warp = nullptr;

#ifdef EMU_FAST_SWITCH
inline void to_sched() { emu_switch(&g_warp->ctx[g_warp->cur], &g_warp->sched); }
inline void to_lane(int l) { emu_switch(&g_warp->sched, &g_warp->ctx[l]); }
#else
inline void to_sched() { swapcontext(&g_warp->ctx[g_warp->cur], &g_warp->sched); }
inline void to_lane(int l) { swapcontext(&g_warp->sched, &g_warp->ctx[l]); }
#endif
inline void yield_lane() { to_sched(); }

inline void barrier(unsigned mask = 0xffffffffu) {
  Warp* w = g_warp;
  const int key = __builtin_ctz(mask), need = __builtin_popcount(mask);
  if (!((mask >> w->cur) & 1u)) { fprintf(stderr, "warp_emul: lane %d is not in the mask %08x it passed\n", w->cur, mask); abort(); }
  const unsigned long my = w->generation[key];
  if (++w->arrived[key] == need) { w->arrived[key] = 0; ++w->generation[key]; }
  else while (w->generation[key] == my) yield_lane();
}

inline uint32_t exchange(unsigned mask, uint32_t v, int src_lane) {
  Warp* w = g_warp;
  const int key = __builtin_ctz(mask);
  const int par = (int)(w->generation[key] & 1);
  w->slots[par][w->cur] = v;
  barrier(mask);
  return w->slots[par][src_lane & 31];
}

static void trampoline() {
  Warp* w = g_warp;
  w->body(w->arg);
  w->done[w->cur] = true;
  to_sched();
  abort();   // a finished lane is never resumed
}

// runs body(arg) on 32 lanes (threadIdx.x = lane) until all have returned; false if the warp deadlocks
// (some lanes wait at a barrier the others never reach: divergent use of a full-mask primitive)
inline bool run_warp(void (*body)(void*), void* arg, size_t stack_bytes = 1 << 20) {
  Warp w;
  g_warp = &w;
  w.body = body; w.arg = arg;
  for (int l = 0; l < 32; ++l) {
    w.done[l] = false; w.arrived[l] = 0; w.generation[l] = 0;
    w.stack[l].resize(stack_bytes);
#ifdef EMU_FAST_SWITCH
    // a fresh stack that emu_switch "returns" into: six zeroed callee-saved slots, the entry point as the return
    // address, one pad word so that the entry sees the alignment of a called function
    uintptr_t top = ((uintptr_t)w.stack[l].data() + stack_bytes) & ~(uintptr_t)15;
    void** sp = (void**)top - 8;
    for (int k = 0; k < 6; ++k) sp[k] = nullptr;
    sp[6] = (void*)trampoline;
    sp[7] = nullptr;
    w.ctx[l] = (void*)sp;
#else
    getcontext(&w.ctx[l]);
    w.ctx[l].uc_stack.ss_sp = w.stack[l].data();
    w.ctx[l].uc_stack.ss_size = stack_bytes;
    w.ctx[l].uc_link = &w.sched;
    makecontext(&w.ctx[l], (void (*)())trampoline, 0);
#endif
  }
  int live = 32;
  unsigned long idle_rounds = 0, last_sum = 0;
  while (live > 0) {
    for (int l = 0; l < 32; ++l) {
      if (w.done[l]) continue;
      w.cur = l;
      threadIdx.x = (unsigned)l;
      to_lane(l);
      if (w.done[l]) --live;
    }
    unsigned long sum = (unsigned long)(32 - live);
    for (int l = 0; l < 32; ++l) sum += w.generation[l];
    // a full round in which no barrier completed and no lane finished: the remaining lanes wait for lanes that
    // will never arrive (divergent use of a masked primitive)
    if (sum == last_sum && live > 0) { if (++idle_rounds > 4) { g_warp = nullptr; return false; } }
    else idle_rounds = 0;
    last_sum = sum;
  }
  g_warp = nullptr;
  return true;
}

}  // namespace emu

namespace emu {
template <class T> inline T xchg(unsigned mask, T v, int src) { static_assert(sizeof(T) == 4, "32-bit shuffles only"); uint32_t u; __builtin_memcpy(&u, &v, 4); u = exchange(mask, u, src); __builtin_memcpy(&v, &u, 4); return v; }
}
// `width` as in CUDA: lanes are cut into segments of `width`, source lanes are relative to the caller's segment
template <class T> inline T __shfl_sync(unsigned m, T v, int src, int width = 32) { const int lane = emu::g_warp->cur; return emu::xchg(m, v, (lane & ~(width - 1)) | (src & (width - 1))); }
template <class T> inline T __shfl_up_sync(unsigned m, T v, unsigned d, int width = 32) { const int lane = emu::g_warp->cur; return emu::xchg(m, v, (lane & (width - 1)) >= (int)d ? lane - (int)d : lane); }
template <class T> inline T __shfl_down_sync(unsigned m, T v, unsigned d, int width = 32) { const int lane = emu::g_warp->cur; return emu::xchg(m, v, (lane & (width - 1)) + (int)d < width ? lane + (int)d : lane); }
template <class T> inline T __shfl_xor_sync(unsigned m, T v, int x, int width = 32) { const int lane = emu::g_warp->cur; const int t = lane ^ x; return emu::xchg(m, v, (t & ~(width - 1)) == (lane & ~(width - 1)) ? t : lane); }
inline void __syncwarp(unsigned m = 0xffffffffu) { emu::barrier(m); }
// REDUX (sm_80+): reduction over the lanes of the mask, every lane gets the result
inline int __reduce_max_sync(unsigned m, int v) { int r = v; bool any = false; for (int l = 0; l < 32; ++l) if ((m >> l) & 1u) { const int u = emu::xchg(m, v, l); r = any ? (u > r ? u : r) : u; any = true; } return r; }
inline int __reduce_min_sync(unsigned m, int v) { int r = v; bool any = false; for (int l = 0; l < 32; ++l) if ((m >> l) & 1u) { const int u = emu::xchg(m, v, l); r = any ? (u < r ? u : r) : u; any = true; } return r; }
inline int __ffs(int x) { return __builtin_ffs(x); }
struct int4 { int x, y, z, w; };
inline int4 make_int4(int x, int y, int z, int w) { int4 r = {x, y, z, w}; return r; }
inline unsigned __ballot_sync(unsigned m, int pred) { unsigned r = 0; for (int l = 0; l < 32; ++l) if ((m >> l) & 1u) r |= (emu::exchange(m, pred ? 1u : 0u, l) & 1u) << l; return r; }
inline unsigned atomicAdd(unsigned* p, unsigned v) { const unsigned o = *p; *p = o + v; return o; }
inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) { const unsigned long long o = *p; *p = o + v; return o; }
inline long long clock64() { return 0; }
inline int min(int a, int b) { return a < b ? a : b; }
inline int max(int a, int b) { return a > b ? a : b; }
inline unsigned min(unsigned a, unsigned b) { return a < b ? a : b; }
inline unsigned max(unsigned a, unsigned b) { return a > b ? a : b; }
inline long long min(long long a, long long b) { return a < b ? a : b; }
inline long long max(long long a, long long b) { return a > b ? a : b; }
