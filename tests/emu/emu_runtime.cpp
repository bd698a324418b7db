// Fiber scheduler of the CPU emulation shim (see cuda_emu.h) -- test infrastructure only.
#include "cuda_emu.h"

#include <mutex>

namespace emu {

thread_local Cta* g_cta = nullptr;
static bool g_concurrent = false;
void set_concurrent(bool on) { g_concurrent = on; }
static unsigned g_max_workers = 0;  // 0 = one per host core; 1 = CTAs run one after the other in blockIdx order (deterministic fp32 atomics)
void set_max_workers(int n) { g_max_workers = n > 0 ? (unsigned)n : 0u; }
bool concurrent() { return g_concurrent; }
static int g_sm_count = 4;
int sm_count() { return g_sm_count; }
void set_sm_count(int n) { g_sm_count = n < 1 ? 1 : n; }

static void fiber_entry() {
  Cta* c = g_cta;
  c->body();
  Fiber& f = c->fibers[c->cur];
  f.done = true;
  c->alive -= 1;
  // a thread that exits no longer takes part in CTA barriers (CUDA semantics): release waiters it was holding up
  if (c->bar_arrived > 0 && c->bar_arrived >= c->alive) {
    c->bar_arrived = 0;
    c->bar_gen += 1;
  }
  swapcontext(&f.ctx, &c->main);
}

static void run_cta(Cta& c, unsigned bx, dim3 grid, dim3 block, size_t dyn_bytes, const std::function<void()>& body) {
  if (dyn_bytes) {
    if (posix_memalign(&c.dyn_smem, 128, dyn_bytes)) abort();
    memset(c.dyn_smem, 0xcd, dyn_bytes);
  }
  const int n = (int)(block.x * block.y * block.z);
  if (n > kMaxThreads) {
    fprintf(stderr, "[emu] block too large\n");
    abort();
  }
  if ((int)c.fibers.size() < n) c.fibers.resize(n);
  c.nthreads = c.alive = n;
  c.bar_gen = 0;
  c.bar_arrived = 0;
  for (auto& w : c.warps) w = WarpSlots();
  c.bid = make_uint3(bx, 0, 0);
  c.bdim = block;
  c.gdim = grid;
  c.body = body;
  g_cta = &c;
  for (int t = 0; t < n; ++t) {
    Fiber& f = c.fibers[t];
    if (!f.stack) f.stack = static_cast<char*>(malloc(kStackBytes));
    f.done = false;
    f.tid = make_uint3((unsigned)t, 0, 0);
    getcontext(&f.ctx);
    f.ctx.uc_stack.ss_sp = f.stack;
    f.ctx.uc_stack.ss_size = kStackBytes;
    f.ctx.uc_link = &c.main;
    makecontext(&f.ctx, fiber_entry, 0);
  }
  long long idle_rounds = 0;
  while (c.alive > 0) {
    const int before = c.alive;
    for (int t = 0; t < n; ++t) {
      if (c.fibers[t].done) continue;
      c.cur = t;
      swapcontext(&c.main, &c.fibers[t].ctx);
    }
    // a full round in which nobody finished is normal (fibers waiting on other CTAs); give the OS a chance
    if (c.alive == before && ((++idle_rounds) & 63) == 0) std::this_thread::yield();
  }
  g_cta = nullptr;
  free(c.dyn_smem);
  c.dyn_smem = nullptr;
}

void run_grid(dim3 grid, dim3 block, size_t dyn_bytes, bool all_concurrent, const std::function<void()>& body) {
  const unsigned G = grid.x * grid.y * grid.z;
  if (G == 0) return;
  unsigned workers = all_concurrent ? G : std::min<unsigned>(G, std::max(2u, std::thread::hardware_concurrency()));
  if (!all_concurrent && g_max_workers) workers = std::min(workers, g_max_workers);
  std::atomic<unsigned> next{0};
  auto work = [&]() {
    Cta* c = new Cta();
    for (;;) {
      const unsigned bx = next.fetch_add(1);
      if (bx >= G) break;
      run_cta(*c, bx, grid, block, dyn_bytes, body);
    }
    for (auto& f : c->fibers) free(f.stack);
    delete c;
  };
  if (workers == 1) {
    work();
    return;
  }
  std::vector<std::thread> pool;
  for (unsigned i = 0; i < workers; ++i) pool.emplace_back(work);
  for (auto& t : pool) t.join();
}

}  // namespace emu
