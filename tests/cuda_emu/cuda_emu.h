// cuda_emu.h -- TEST INFRASTRUCTURE, never part of the product.
//
// A minimal CUDA execution model on host threads, just large enough to compile the product's own kernel
// sources (splat_b200/csrc/*.cuh, splat_api.cu) with g++ and run them on a box without a GPU, so that the
// C-ABI parity tests can exercise the real kernel logic and the real frame orchestration on the CPU
// (tests/test_emu_parity.py; built by tests/cuda_emu/emu_build.py into tests/cuda_emu/_build/).  The product
// never loads that build: splat_b200/_lib.py knows one path, libsplat_b200.so, and fails without it.
//
// Model: a block's threads are cooperative fibers of one OS worker thread, several workers run different blocks
// of a launch concurrently (`__shared__` = static thread_local).  __syncthreads / named barriers / warp
// collectives are barriers whose participant count follows the threads that have not yet returned, as on the
// hardware; a barrier that can never complete is reported as a deadlock.  mbarriers (arrive / expect_tx /
// complete_tx / parity test) live in the 8 bytes the kernel gives them.  Device memory is host memory; streams
// are in-order and synchronous, so every stream dependency of the host code holds trivially.  IEEE arithmetic:
// compile with -ffp-contract=off -frounding-math (the product is built with --fmad=false); __f*_rn map to the
// C operators, fmaf is the correctly rounded one.
// What this cannot show: performance, and races between the threads of a block (fibers do not pre-empt each other).
#pragma once
#include <sys/mman.h>
#include <ucontext.h>

#include <atomic>
#include <cfenv>
#include <chrono>
#include <cmath>
#include <condition_variable>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <mutex>
#include <thread>
#include <tuple>
#include <type_traits>
#include <vector>

#define SPLAT_EMU 1
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __noinline__ __attribute__((noinline))
#define __restrict__
#define __shared__ static thread_local
#define __launch_bounds__(...)
#define __grid_constant__
#define __align__(n) __attribute__((aligned(n)))

// ------------------------------------------------------------------------------------------------ vector types
struct alignas(8) uint2 { unsigned int x, y; };
struct alignas(16) uint4 { unsigned int x, y, z, w; };
struct alignas(8) float2 { float x, y; };
struct alignas(16) float4 { float x, y, z, w; };
struct alignas(8) int2 { int x, y; };
inline uint2 make_uint2(unsigned x, unsigned y) { return {x, y}; }
inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { return {x, y, z, w}; }
inline float2 make_float2(float x, float y) { return {x, y}; }
inline float4 make_float4(float x, float y, float z, float w) { return {x, y, z, w}; }
struct dim3 {
  unsigned x, y, z;
  constexpr dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};

// ------------------------------------------------------------------------------------------------ execution engine
namespace emu {

// A block's threads are FIBERS (ucontext) of one OS worker thread, switched cooperatively at barriers, warp
// collectives and mbarrier polls; several workers run different blocks of a launch at the same time, which is
// why `__shared__` is `static thread_local` (a block never leaves its worker).  A round of the scheduler in
// which nothing can run while some fibers still wait is a deadlock (e.g. a barrier not reached by every live
// thread) and aborts with a description -- on the hardware that would be a hang.
struct Thread {
  dim3 threadIdx, blockIdx;
  int tid = 0, lane = 0, warp = 0;
};

struct Barrier {               // participants: the fibers that have not returned yet (or a fixed number)
  int count = 0;
  unsigned gen = 0;
};

struct Warp {
  int live = 0;
  unsigned live_mask = 0;
  Barrier bar;
  unsigned long long slot[32];
};

enum FiberState { READY = 0, BLOCKED = 1, DONE = 2 };
#if defined(__x86_64__)
// callee-saved registers on the fiber's own stack, one pointer to switch (a swapcontext costs a system call)
extern "C" void emu_switch(void **save_sp, void *load_sp);
asm(R"(
.text
.weak emu_switch
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
struct Context { void *sp = nullptr; const void *bottom = nullptr; size_t size = 0; };
#if defined(__SANITIZE_ADDRESS__)
// AddressSanitizer build (EMU_ASAN=1: out-of-bounds accesses of "device" and shared memory become reports): tell
// the sanitizer about every stack switch.  `final`: the fiber being left will never run again.
extern "C" void __sanitizer_start_switch_fiber(void **fake_stack_save, const void *bottom, size_t size);
extern "C" void __sanitizer_finish_switch_fiber(void *fake_stack_save, const void **bottom_old, size_t *size_old);
inline void ctx_switch(Context &from, Context &to, bool final = false) {
  void *fake = nullptr;
  __sanitizer_start_switch_fiber(final ? nullptr : &fake, to.bottom, to.size);
  emu_switch(&from.sp, to.sp);
  __sanitizer_finish_switch_fiber(fake, nullptr, nullptr);
}
inline void ctx_entered(Context &came_from) {        // first instruction of a new fiber: learn the scheduler's stack
  const void *b = nullptr;
  size_t n = 0;
  __sanitizer_finish_switch_fiber(nullptr, &b, &n);
  came_from.bottom = b;
  came_from.size = n;
}
#else
inline void ctx_switch(Context &from, Context &to, bool = false) { emu_switch(&from.sp, to.sp); }
inline void ctx_entered(Context &) {}
#endif
inline void ctx_make(Context &c, void *stack, size_t size, void (*entry)()) {
  c.bottom = stack;
  c.size = size;
  uintptr_t top = (reinterpret_cast<uintptr_t>(stack) + size) & ~(uintptr_t)15;
  void **sp = reinterpret_cast<void **>(top);
  *--sp = nullptr;                                  // return address of entry (it never returns)
  *--sp = reinterpret_cast<void *>(entry);          // popped by emu_switch's ret
  for (int i = 0; i < 6; ++i) *--sp = nullptr;      // rbp rbx r12 r13 r14 r15
  c.sp = sp;
}
#else
struct Context { ucontext_t uc; };
inline void ctx_switch(Context &from, Context &to, bool = false) { swapcontext(&from.uc, &to.uc); }
inline void ctx_entered(Context &) {}
inline void ctx_make(Context &c, void *stack, size_t size, void (*entry)()) {
  getcontext(&c.uc);
  c.uc.uc_stack.ss_sp = stack;
  c.uc.uc_stack.ss_size = size;
  c.uc.uc_link = nullptr;
  makecontext(&c.uc, entry, 0);
}
#endif

struct Fiber {
  Context ctx;
  Thread t;
  FiberState st = DONE;
  const Barrier *waits_on = nullptr;
  unsigned waits_gen = 0;
  void *stack = nullptr;
};

#if defined(__SANITIZE_ADDRESS__)
constexpr size_t FIBER_STACK = 64u << 10;    // the sanitizer touches the shadow of a whole stack at every fiber start
#else
constexpr size_t FIBER_STACK = 256u << 10;
#endif

struct Worker {                 // one per OS thread; runs one block at a time
  std::vector<Fiber> fibers;
  Context sched;
  Fiber *cur = nullptr;
  int live = 0;
  Barrier bars[16];             // 0 = __syncthreads, others = bar.sync id, n
  int acc = 0;
  std::vector<Warp> warps;
  std::vector<unsigned char> dyn_store;
  unsigned char *dyn = nullptr;
  const std::function<void()> *body = nullptr;
  ~Worker() { for (auto &f : fibers) if (f.stack) munmap(f.stack, FIBER_STACK); }
};

inline Worker &worker() { static thread_local Worker w; return w; }
inline dim3 &block_dim() { static dim3 d; return d; }
inline dim3 &grid_dim() { static dim3 d; return d; }
inline Thread &self() { return worker().cur->t; }
inline int &last_error() { static int e = 0; return e; }
inline std::mutex &launch_mutex() { static std::mutex m; return m; }   // one launch at a time, process-wide

struct cfg {
  dim3 grid, block;
  size_t smem;
  template <class S = void *>
  cfg(dim3 g, dim3 b, size_t s = 0, S = S()) : grid(g), block(b), smem(s) {}
};

inline void *dyn_smem() { return worker().dyn; }

inline void yield() {           // back to the scheduler; returns when this fiber is picked again
  Worker &w = worker();
  Fiber *f = w.cur;
  ctx_switch(f->ctx, w.sched);
}

// EMU_ORDER=random: besides the random order at barriers, a thread may lose its turn at any global load through
// __ldg and at any atomic -- more interleavings of the threads of a block than barriers alone produce
inline void maybe_yield() {
  static const bool on = [] { const char *e = std::getenv("EMU_ORDER"); return e && std::strcmp(e, "random") == 0; }();
  if (!on) return;
  static thread_local unsigned long long r = 0xD1B54A32D192ED03ull;
  r ^= r << 13; r ^= r >> 7; r ^= r << 17;
  if ((r & 15u) == 0u && worker().cur) yield();
}

inline void release(Worker &w, Barrier &b) {
  b.count = 0;
  b.gen += 1;
  for (auto &f : w.fibers)
    if (f.st == BLOCKED && f.waits_on == &b) { f.st = READY; f.waits_on = nullptr; }
}
inline void barrier_wait(Barrier &b, int expected) {
  Worker &w = worker();
  if (++b.count >= expected) { release(w, b); return; }
  Fiber *f = w.cur;
  f->st = BLOCKED;
  f->waits_on = &b;
  yield();
}
inline void fiber_main() {
  Worker &w = worker();
  ctx_entered(w.sched);
  (*w.body)();
  // the thread has returned: it no longer takes part in any barrier of this block
  Fiber *f = w.cur;
  Warp &wp = w.warps[f->t.warp];
  wp.live_mask &= ~(1u << f->t.lane);
  wp.live -= 1;
  if (wp.bar.count > 0 && wp.bar.count >= wp.live) release(w, wp.bar);
  w.live -= 1;
  if (w.bars[0].count > 0 && w.bars[0].count >= w.live) release(w, w.bars[0]);
  f->st = DONE;
  ctx_switch(f->ctx, w.sched, true);
  std::abort();   // a finished fiber is never resumed
}

inline void run_block(Worker &w, const cfg &c, unsigned long long b, int B, const std::function<void()> &body) {
  const int nw = (B + 31) / 32;
  if ((int)w.fibers.size() < B) {
    const size_t old = w.fibers.size();
    w.fibers.resize((size_t)B);       // (no fiber is live while the vector grows)
    for (size_t i = old; i < (size_t)B; ++i) {
      w.fibers[i].stack = mmap(nullptr, FIBER_STACK, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE | MAP_STACK, -1, 0);
      if (w.fibers[i].stack == MAP_FAILED) { std::perror("cuda_emu: mmap"); std::abort(); }
    }
  }
  if ((int)w.warps.size() < nw) w.warps.resize((size_t)nw);
  if (w.dyn_store.size() < c.smem + 256) w.dyn_store.resize(c.smem + 256);
  w.dyn = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(w.dyn_store.data()) + 127) & ~(uintptr_t)127);
  w.body = &body;
  w.live = B;
  w.acc = 0;
  for (auto &br : w.bars) br = Barrier{};
  for (int k = 0; k < nw; ++k) {
    const int lanes = std::min(32, B - 32 * k);
    w.warps[k].live = lanes;
    w.warps[k].live_mask = lanes == 32 ? 0xFFFFFFFFu : ((1u << lanes) - 1u);
    w.warps[k].bar = Barrier{};
  }
  const dim3 bidx((unsigned)(b % c.grid.x), (unsigned)((b / c.grid.x) % c.grid.y), (unsigned)(b / ((unsigned long long)c.grid.x * c.grid.y)));
  for (int i = 0; i < B; ++i) {
    Fiber &f = w.fibers[i];
    f.t.tid = i; f.t.lane = i & 31; f.t.warp = i >> 5;
    f.t.threadIdx = dim3((unsigned)i % c.block.x, ((unsigned)i / c.block.x) % c.block.y, (unsigned)i / (c.block.x * c.block.y));
    f.t.blockIdx = bidx;
    f.st = READY;
    f.waits_on = nullptr;
    ctx_make(f.ctx, f.stack, FIBER_STACK, fiber_main);
  }
  // order in which the runnable threads of a block get their turn: 0 = ascending (default), 1 = descending,
  // 2 = a new random permutation every round (EMU_ORDER=reverse|random).  Code that relies on a barrier it does not
  // have -- a shared-memory hand-over without __syncthreads, warp lockstep without __syncwarp -- works under at
  // most one of the fixed orders.
  static const int order_mode = [] {
    const char *e = std::getenv("EMU_ORDER");
    return !e ? 0 : (std::strcmp(e, "reverse") == 0 ? 1 : (std::strcmp(e, "random") == 0 ? 2 : 0));
  }();
  static thread_local std::vector<int> perm;
  static thread_local unsigned long long rng = 0x9E3779B97F4A7C15ull ^ (unsigned long long)(uintptr_t)&perm;
  if (order_mode) {
    perm.resize((size_t)B);
    for (int i = 0; i < B; ++i) perm[(size_t)i] = order_mode == 1 ? B - 1 - i : i;
  }
  int done = 0;
  while (done < B) {
    bool ran = false;
    if (order_mode == 2)
      for (int i = B - 1; i > 0; --i) {
        rng ^= rng << 13; rng ^= rng >> 7; rng ^= rng << 17;
        std::swap(perm[(size_t)i], perm[(size_t)(rng % (unsigned long long)(i + 1))]);
      }
    for (int ii = 0; ii < B; ++ii) {
      const int i = order_mode ? perm[(size_t)ii] : ii;
      Fiber &f = w.fibers[i];
      if (f.st != READY) continue;
      ran = true;
      w.cur = &f;
      ctx_switch(w.sched, f.ctx);
      if (f.st == DONE) ++done;
    }
    if (!ran && done < B) {
      std::fprintf(stderr, "cuda_emu: DEADLOCK in block %llu: %d of %d threads wait on a barrier nobody else will reach\n", b, B - done, B);
      for (int i = 0; i < B; ++i)
        if (w.fibers[i].st == BLOCKED) {
          const Barrier *wb = w.fibers[i].waits_on;
          const char *kind = (wb >= w.bars && wb < w.bars + 16) ? "block barrier" : "warp barrier";
          std::fprintf(stderr, "  thread %d: %s %d\n", i, kind, (wb >= w.bars && wb < w.bars + 16) ? (int)(wb - w.bars) : i >> 5);
        }
      std::abort();
    }
  }
  w.cur = nullptr;
}

// persistent OS worker threads (their fiber stacks are kept between launches)
class Pool {
 public:
  static Pool &get(unsigned n) { static Pool *p = new Pool(n); return *p; }   // leaked on purpose: its threads outlive main()
  void run(unsigned nworkers, const std::function<void()> &job) {
    if (nworkers <= 1) { job(); return; }        // on the calling thread
    std::unique_lock<std::mutex> lk(m_);
    job_ = &job;
    want_ = std::min<unsigned>(nworkers, (unsigned)threads_.size());
    taken_ = 0;
    running_ = want_;
    gen_ += 1;
    cv_.notify_all();
    done_cv_.wait(lk, [&] { return running_ == 0; });
    job_ = nullptr;
  }
 private:
  explicit Pool(unsigned n) {
    for (unsigned i = 0; i < n; ++i) threads_.emplace_back([this] { loop(); });
    for (auto &t : threads_) t.detach();         // they live as long as the process
  }
  void loop() {
    unsigned seen = 0;
    for (;;) {
      const std::function<void()> *job = nullptr;
      {
        std::unique_lock<std::mutex> lk(m_);
        cv_.wait(lk, [&] { return gen_ != seen; });
        seen = gen_;
        if (taken_ >= want_) continue;
        taken_ += 1;
        job = job_;
      }
      (*job)();
      {
        std::unique_lock<std::mutex> lk(m_);
        if (--running_ == 0) done_cv_.notify_all();
      }
    }
  }
  std::mutex m_;
  std::condition_variable cv_, done_cv_;
  std::vector<std::thread> threads_;
  const std::function<void()> *job_ = nullptr;
  unsigned want_ = 0, taken_ = 0, running_ = 0, gen_ = 0;
};

template <class Body>
void run_grid(const cfg &c, Body body_fn, const char *name = "") {
  std::lock_guard<std::mutex> launch_lock(launch_mutex());
  static const bool trace = std::getenv("EMU_TRACE") != nullptr;
  const auto t_start = std::chrono::steady_clock::now();
  const int B = (int)(c.block.x * c.block.y * c.block.z);
  const unsigned long long G = (unsigned long long)c.grid.x * c.grid.y * c.grid.z;
  if (B <= 0 || B > 1024 || G == 0 || c.grid.x > 0x7FFFFFFFu || c.grid.y > 65535u || c.grid.z > 65535u || c.smem > 227u * 1024u) {
    // what the runtime answers with cudaErrorInvalidConfiguration: a launch that silently did nothing here would hide it
    std::fprintf(stderr, "cuda_emu: invalid launch configuration for %s: grid (%u,%u,%u) block (%u,%u,%u) smem %zu\n", name, c.grid.x, c.grid.y,
                 c.grid.z, c.block.x, c.block.y, c.block.z, c.smem);
    last_error() = 9;   // cudaErrorInvalidConfiguration
    return;
  }
  block_dim() = c.block;
  grid_dim() = c.grid;
  const std::function<void()> body = body_fn;
  static const unsigned max_workers = [] {
    const char *e = std::getenv("EMU_WORKERS");
    const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
    return e ? (unsigned)std::max(1, std::atoi(e)) : std::min(hw, 32u);
  }();
  const unsigned nworkers = (unsigned)std::min<unsigned long long>(max_workers, G);
  std::atomic<unsigned long long> next{0};
  const std::function<void()> work = [&] {
    Worker &w = worker();
    for (;;) {
      const unsigned long long b = next.fetch_add(1);
      if (b >= G) break;
      run_block(w, c, b, B, body);
    }
  };
  Pool::get(max_workers).run(nworkers, work);
  if (trace)
    std::fprintf(stderr, "[emu] %-28s grid %6llu x %4d threads  %8.1f ms\n", name, G, B,
                 std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_start).count());
}

inline void warp_sync() {
  Worker &w = worker();
  Warp &wp = w.warps[self().warp];
  barrier_wait(wp.bar, wp.live);
}
// every live lane publishes a 64-bit value; f(slots, live_mask) computes this lane's result
template <class F>
auto warp_exchange(unsigned long long mine, F f) {
  Warp &wp = worker().warps[self().warp];
  wp.slot[self().lane] = mine;
  warp_sync();
  auto r = f(wp.slot, wp.live_mask);
  warp_sync();
  return r;
}
template <class T>
unsigned long long to_bits(T v) { unsigned long long b = 0; static_assert(sizeof(T) <= 8); std::memcpy(&b, &v, sizeof(T)); return b; }
template <class T>
T from_bits(unsigned long long b) { T v; std::memcpy(&v, &b, sizeof(T)); return v; }

// ---- mbarrier in the caller's 8 bytes: [0,20) pending arrivals, [20,40) arrival count per phase,
// [40,63) pending transaction bytes + 2^22 bias, bit 63 = phase parity
constexpr unsigned long long MB_TX_BIAS = 1ull << 22;
inline unsigned long long mb_pack(unsigned long long pend, unsigned long long init, unsigned long long tx, unsigned long long phase) {
  return pend | (init << 20) | (tx << 40) | (phase << 63);
}
inline void mb_update(uint64_t *bar, long long d_pend, long long d_tx) {
  std::atomic_ref<uint64_t> a(*bar);
  uint64_t old = a.load(std::memory_order_acquire);
  for (;;) {
    long long pend = (long long)(old & 0xFFFFF), init = (long long)((old >> 20) & 0xFFFFF);
    long long tx = (long long)((old >> 40) & 0x7FFFFF);
    unsigned long long phase = old >> 63;
    pend += d_pend; tx += d_tx;
    if (pend < 0) { std::fprintf(stderr, "cuda_emu: mbarrier over-arrival\n"); std::abort(); }
    if (pend == 0 && tx == (long long)MB_TX_BIAS) { phase ^= 1ull; pend = init; }
    const uint64_t neu = mb_pack((unsigned long long)pend, (unsigned long long)init, (unsigned long long)tx, phase);
    if (a.compare_exchange_weak(old, neu, std::memory_order_acq_rel, std::memory_order_acquire)) return;
  }
}
inline void mb_init(uint64_t *bar, unsigned count) { std::atomic_ref<uint64_t>(*bar).store(mb_pack(count, count, MB_TX_BIAS, 0), std::memory_order_release); }
inline bool mb_test(uint64_t *bar, unsigned parity) { return (std::atomic_ref<uint64_t>(*bar).load(std::memory_order_acquire) >> 63) != (parity & 1u); }
inline void mb_spin(uint64_t *bar, unsigned parity) {
  while (!mb_test(bar, parity)) yield();
}

struct RoundTowardZero {
  int old;
  RoundTowardZero() : old(std::fegetround()) { std::fesetround(FE_TOWARDZERO); }
  ~RoundTowardZero() { std::fesetround(old); }
};

}  // namespace emu

#define threadIdx (emu::self().threadIdx)
#define blockIdx (emu::self().blockIdx)
#define blockDim (emu::block_dim())
#define gridDim (emu::grid_dim())

// ------------------------------------------------------------------------------------------------ block / warp intrinsics
inline void __syncthreads() {
  emu::Worker &w = emu::worker();
  emu::barrier_wait(w.bars[0], w.live);
}
inline void emu_bar_sync(int id, unsigned nthreads) { emu::barrier_wait(emu::worker().bars[id], (int)nthreads); }
inline int __syncthreads_count(int pred) {
  emu::Worker &w = emu::worker();
  if (pred) w.acc += 1;
  __syncthreads();
  const int r = w.acc;
  __syncthreads();
  w.acc = 0;
  __syncthreads();
  return r;
}
inline int __syncthreads_and(int pred) {
  emu::Worker &w = emu::worker();
  if (!pred) w.acc += 1;
  __syncthreads();
  const int r = w.acc == 0;
  __syncthreads();
  w.acc = 0;
  __syncthreads();
  return r;
}
inline int __syncthreads_or(int pred) { return __syncthreads_count(pred) != 0; }
inline void __syncwarp(unsigned = 0xFFFFFFFFu) { emu::warp_sync(); }
inline unsigned __ballot_sync(unsigned, int pred) {
  return emu::warp_exchange(pred ? 1ull : 0ull, [](const unsigned long long *s, unsigned live) {
    unsigned r = 0;
    for (int l = 0; l < 32; ++l) if (((live >> l) & 1u) && s[l]) r |= 1u << l;
    return r;
  });
}
inline int __any_sync(unsigned m, int pred) { return __ballot_sync(m, pred) != 0u; }
inline int __all_sync(unsigned m, int pred) { return __ballot_sync(m, !pred) == 0u; }
template <class T>
T __shfl_sync(unsigned, T v, int src, int width = 32) {
  const int lane = emu::self().lane;
  const int s = (lane & ~(width - 1)) | (src & (width - 1));
  return emu::warp_exchange(emu::to_bits(v), [=](const unsigned long long *sl, unsigned live) {
    return ((live >> s) & 1u) ? emu::from_bits<T>(sl[s]) : v;
  });
}
template <class T>
T __shfl_up_sync(unsigned, T v, unsigned delta, int width = 32) {
  const int lane = emu::self().lane;
  const int s = lane - (int)delta;
  const bool ok = s >= (lane & ~(width - 1));
  return emu::warp_exchange(emu::to_bits(v), [=](const unsigned long long *sl, unsigned live) {
    return (ok && ((live >> s) & 1u)) ? emu::from_bits<T>(sl[s]) : v;
  });
}
template <class T>
T __shfl_down_sync(unsigned, T v, unsigned delta, int width = 32) {
  const int lane = emu::self().lane;
  const int s = lane + (int)delta;
  const bool ok = s < (lane & ~(width - 1)) + width;
  return emu::warp_exchange(emu::to_bits(v), [=](const unsigned long long *sl, unsigned live) {
    return (ok && ((live >> s) & 1u)) ? emu::from_bits<T>(sl[s]) : v;
  });
}
template <class T>
T __shfl_xor_sync(unsigned, T v, int mask, int width = 32) {
  const int lane = emu::self().lane;
  const int s = lane ^ mask;
  const bool ok = (s & ~(width - 1)) == (lane & ~(width - 1));
  return emu::warp_exchange(emu::to_bits(v), [=](const unsigned long long *sl, unsigned live) {
    return (ok && ((live >> s) & 1u)) ? emu::from_bits<T>(sl[s]) : v;
  });
}

// ------------------------------------------------------------------------------------------------ atomics, loads, fences
template <class T>
T atomicAdd(T *p, T v) { emu::maybe_yield(); return __atomic_fetch_add(p, v, __ATOMIC_ACQ_REL); }
inline unsigned atomicAdd(unsigned *p, int v) { return __atomic_fetch_add(p, (unsigned)v, __ATOMIC_ACQ_REL); }
inline float atomicAdd(float *p, float v) {
  std::atomic_ref<float> a(*p);
  float old = a.load();
  while (!a.compare_exchange_weak(old, old + v)) {}
  return old;
}
template <class T> T atomicMax(T *p, T v) { emu::maybe_yield(); T old = __atomic_load_n(p, __ATOMIC_ACQUIRE); while (old < v && !__atomic_compare_exchange_n(p, &old, v, true, __ATOMIC_ACQ_REL, __ATOMIC_ACQUIRE)) {} return old; }
template <class T> T atomicMin(T *p, T v) { T old = __atomic_load_n(p, __ATOMIC_ACQUIRE); while (old > v && !__atomic_compare_exchange_n(p, &old, v, true, __ATOMIC_ACQ_REL, __ATOMIC_ACQUIRE)) {} return old; }
template <class T> T atomicOr(T *p, T v) { return __atomic_fetch_or(p, v, __ATOMIC_ACQ_REL); }
template <class T> T atomicAnd(T *p, T v) { return __atomic_fetch_and(p, v, __ATOMIC_ACQ_REL); }
template <class T> T atomicExch(T *p, T v) { return __atomic_exchange_n(p, v, __ATOMIC_ACQ_REL); }
template <class T> T atomicCAS(T *p, T cmp, T v) { emu::maybe_yield(); __atomic_compare_exchange_n(p, &cmp, v, false, __ATOMIC_ACQ_REL, __ATOMIC_ACQUIRE); return cmp; }
template <class T> T __ldg(const T *p) { emu::maybe_yield(); return *p; }
inline void __threadfence() { std::atomic_thread_fence(std::memory_order_seq_cst); }
inline void __threadfence_block() { std::atomic_thread_fence(std::memory_order_seq_cst); }
inline void __threadfence_system() { std::atomic_thread_fence(std::memory_order_seq_cst); }
inline void __nanosleep(unsigned) { emu::yield(); }
[[noreturn]] inline void __trap() { std::fprintf(stderr, "cuda_emu: __trap() in block %u thread %u\n", blockIdx.x, threadIdx.x); std::abort(); }
inline size_t __cvta_generic_to_shared(const void *p) { return reinterpret_cast<size_t>(p); }

// ------------------------------------------------------------------------------------------------ scalar intrinsics
inline int __popc(unsigned v) { return __builtin_popcount(v); }
inline int __ffs(int v) { return __builtin_ffs(v); }
inline int __clz(int v) { return v ? __builtin_clz((unsigned)v) : 32; }
inline unsigned __float_as_uint(float f) { unsigned u; std::memcpy(&u, &f, 4); return u; }
inline int __float_as_int(float f) { int u; std::memcpy(&u, &f, 4); return u; }
inline float __uint_as_float(unsigned u) { float f; std::memcpy(&f, &u, 4); return f; }
inline float __int_as_float(int u) { float f; std::memcpy(&f, &u, 4); return f; }
inline float __fadd_rn(float a, float b) { return a + b; }
inline float __fsub_rn(float a, float b) { return a - b; }
inline float __fmul_rn(float a, float b) { return a * b; }
inline float __fdiv_rn(float a, float b) { return a / b; }
inline float __fsqrt_rn(float a) { return std::sqrt(a); }
inline float __fmaf_rn(float a, float b, float c) { return std::fmaf(a, b, c); }
inline float __fadd_rz(float a, float b) { emu::RoundTowardZero rz; volatile float x = a, y = b; volatile float s = x + y; return s; }
inline float __frcp_rz(float a) { emu::RoundTowardZero rz; volatile float x = a; volatile float s = 1.0f / x; return s; }
inline float __uint2float_rz(unsigned v) { emu::RoundTowardZero rz; volatile unsigned x = v; volatile float s = (float)x; return s; }
inline float __saturatef(float v) { return v != v ? 0.0f : (v < 0.0f ? 0.0f : (v > 1.0f ? 1.0f : v)); }
inline float __log2f(float v) { return std::log2(v); }
template <class A, class B> constexpr std::common_type_t<A, B> min(A a, B b) { using C = std::common_type_t<A, B>; return (C)a < (C)b ? (C)a : (C)b; }
template <class A, class B> constexpr std::common_type_t<A, B> max(A a, B b) { using C = std::common_type_t<A, B>; return (C)a > (C)b ? (C)a : (C)b; }

// ------------------------------------------------------------------------------------------------ PTX that the product issues as inline asm
// (emu_build.py replaces every asm statement of the sources by a call of the function of the same meaning)
namespace emu::ptx {
inline unsigned long long pack(float lo, float hi) { return (unsigned long long)__float_as_uint(lo) | ((unsigned long long)__float_as_uint(hi) << 32); }
inline void unpack(unsigned long long v, float &lo, float &hi) { lo = __uint_as_float((unsigned)v); hi = __uint_as_float((unsigned)(v >> 32)); }
template <class F> unsigned long long lanewise(unsigned long long a, unsigned long long b, unsigned long long c, F f) {
  float a0, a1, b0, b1, c0, c1;
  unpack(a, a0, a1); unpack(b, b0, b1); unpack(c, c0, c1);
  return pack(f(a0, b0, c0), f(a1, b1, c1));
}
inline unsigned long long fma_rn_f32x2(unsigned long long a, unsigned long long b, unsigned long long c) { return lanewise(a, b, c, [](float x, float y, float z) { return std::fmaf(x, y, z); }); }
inline unsigned long long mul_rn_f32x2(unsigned long long a, unsigned long long b) { return lanewise(a, b, 0, [](float x, float y, float) { return x * y; }); }
inline unsigned long long add_rn_f32x2(unsigned long long a, unsigned long long b) { return lanewise(a, b, 0, [](float x, float y, float) { return x + y; }); }
inline unsigned long long sub_rn_f32x2(unsigned long long a, unsigned long long b) { return lanewise(a, b, 0, [](float x, float y, float) { return x - y; }); }
inline unsigned long long add_rz_f32x2(unsigned long long a, unsigned long long b) { return lanewise(a, b, 0, [](float x, float y, float) { return __fadd_rz(x, y); }); }
inline void cp_async_bulk(void *dst, const void *src, unsigned bytes, uint64_t *bar) {
  std::memcpy(dst, src, bytes);
  emu::mb_update(bar, 0, -(long long)bytes);      // complete_tx
}
}  // namespace emu::ptx

// ------------------------------------------------------------------------------------------------ runtime API (synchronous, in-order)
typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorMemoryAllocation = 2, cudaErrorNotReady = 600, cudaErrorInvalidValue = 1 };
typedef struct emu_stream *cudaStream_t;
struct emu_event { std::chrono::steady_clock::time_point t; };
typedef emu_event *cudaEvent_t;
enum cudaMemcpyKind { cudaMemcpyHostToHost, cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyDefault };
enum { cudaStreamNonBlocking = 1, cudaEventDisableTiming = 2, cudaHostAllocMapped = 2, cudaHostRegisterDefault = 0 };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
inline const char *cudaGetErrorString(cudaError_t e) { return e == cudaSuccess ? "no error" : e == 9 ? "invalid configuration argument" : (e == cudaErrorMemoryAllocation ? "out of memory" : "emulated CUDA error"); }
inline cudaError_t cudaGetLastError() { const int e = emu::last_error(); emu::last_error() = 0; return e; }
inline cudaError_t cudaSetDevice(int d) { return d >= 0 && d < 64 ? cudaSuccess : cudaErrorInvalidValue; }   // "devices" are just ordinals
inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
// EMU_GUARD=1: every "device" allocation ends 0..15 bytes before an inaccessible page and starts right after one,
// so a kernel that reads or writes outside a buffer faults at the instruction that does it (a memcheck for global
// memory at full speed).  Default: plain aligned allocations.
namespace emu {
struct GuardMap {
  std::mutex m;
  std::vector<std::tuple<void *, void *, size_t>> live;     // user pointer, mapping base, mapping length
};
inline GuardMap &guards() { static GuardMap *g = new GuardMap(); return *g; }
inline bool guard_mode() { static const bool on = std::getenv("EMU_GUARD") != nullptr; return on; }
inline void *guarded_alloc(size_t bytes) {
  const size_t page = 4096, body = (bytes + 15) & ~(size_t)15, pages = (body + page - 1) / page + 2;
  unsigned char *base = static_cast<unsigned char *>(mmap(nullptr, pages * page, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0));
  if (base == MAP_FAILED) return nullptr;
  mprotect(base, page, PROT_NONE);
  mprotect(base + (pages - 1) * page, page, PROT_NONE);
  unsigned char *user = base + (pages - 1) * page - body;
  std::lock_guard<std::mutex> lk(guards().m);
  guards().live.emplace_back(user, base, pages * page);
  return user;
}
inline bool guarded_free(void *p) {
  std::lock_guard<std::mutex> lk(guards().m);
  auto &v = guards().live;
  for (size_t i = 0; i < v.size(); ++i)
    if (std::get<0>(v[i]) == p) {
      munmap(std::get<1>(v[i]), std::get<2>(v[i]));
      v[i] = v.back();
      v.pop_back();
      return true;
    }
  return false;
}
}  // namespace emu
template <class T> cudaError_t cudaMalloc(T **p, size_t bytes) {
  void *q = nullptr;
  if (emu::guard_mode()) {
    q = emu::guarded_alloc(bytes ? bytes : 16);
    if (!q) return cudaErrorMemoryAllocation;
  } else if (posix_memalign(&q, 256, bytes ? bytes : 256)) {
    return cudaErrorMemoryAllocation;
  }
  std::memset(q, 0xCD, bytes);                  // device memory is not zeroed: make reliance on that visible
  *p = static_cast<T *>(q);
  return cudaSuccess;
}
inline cudaError_t cudaFree(void *p) {
  if (p && emu::guard_mode() && emu::guarded_free(p)) return cudaSuccess;
  std::free(p);
  return cudaSuccess;
}
template <class T> cudaError_t cudaMallocHost(T **p, size_t bytes) { return cudaMalloc(p, bytes); }
template <class T> cudaError_t cudaHostAlloc(T **p, size_t bytes, unsigned) { return cudaMalloc(p, bytes); }
inline cudaError_t cudaFreeHost(void *p) { return cudaFree(p); }
template <class T, class U> cudaError_t cudaHostGetDevicePointer(T **d, U *h, unsigned) { *d = reinterpret_cast<T *>(h); return cudaSuccess; }
inline cudaError_t cudaHostRegister(void *, size_t, unsigned) { return cudaSuccess; }
inline cudaError_t cudaHostUnregister(void *) { return cudaSuccess; }
inline cudaError_t cudaMemcpy(void *d, const void *s, size_t n, cudaMemcpyKind) { std::memmove(d, s, n); return cudaSuccess; }
inline cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, cudaMemcpyKind, cudaStream_t = nullptr) { std::memmove(d, s, n); return cudaSuccess; }
inline cudaError_t cudaMemset(void *d, int v, size_t n) { std::memset(d, v, n); return cudaSuccess; }
inline cudaError_t cudaMemsetAsync(void *d, int v, size_t n, cudaStream_t = nullptr) { std::memset(d, v, n); return cudaSuccess; }
inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned) { static char dummy[64]; static int k = 0; *s = reinterpret_cast<cudaStream_t>(&dummy[(k++) % 64]); return cudaSuccess; }
inline cudaError_t cudaStreamDestroy(cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaStreamQuery(cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return cudaSuccess; }
inline cudaError_t cudaEventCreate(cudaEvent_t *e) { *e = new emu_event{std::chrono::steady_clock::now()}; return cudaSuccess; }
inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t *e, unsigned) { return cudaEventCreate(e); }
inline cudaError_t cudaEventDestroy(cudaEvent_t e) { delete e; return cudaSuccess; }
inline cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t = nullptr) { e->t = std::chrono::steady_clock::now(); return cudaSuccess; }
inline cudaError_t cudaEventQuery(cudaEvent_t) { return cudaSuccess; }
inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
inline cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t a, cudaEvent_t b) { *ms = std::chrono::duration<float, std::milli>(b->t - a->t).count(); return cudaSuccess; }
template <class F> cudaError_t cudaFuncSetAttribute(F, cudaFuncAttribute, int) { return cudaSuccess; }
