// tests/emu/cuda_emu.cpp -- TEST INFRASTRUCTURE ONLY (see cuda_emu.h).
// Fiber scheduler of the CPU CUDA emulator (x86-64 System V only).
#include "cuda_emu.h"

namespace emu {

Block* g_block = nullptr;
Fiber* g_cur = nullptr;
void* g_sched_sp = nullptr;
uint3 g_blockIdx{0, 0, 0};
dim3 g_blockDim, g_gridDim;
unsigned char* g_dyn_smem = nullptr;
uint64_t g_progress = 0;

static const size_t kStack = 256 * 1024;
static unsigned char* g_stacks = nullptr;
static size_t g_stacks_n = 0;
static unsigned char* g_smem_buf = nullptr;

asm(".text\n"
    ".globl emu_switch\n"
    ".type emu_switch,@function\n"
    "emu_switch:\n"
    "  pushq %rbp\n  pushq %rbx\n  pushq %r12\n  pushq %r13\n  pushq %r14\n  pushq %r15\n"
    "  movq %rsp, (%rdi)\n"
    "  movq %rsi, %rsp\n"
    "  popq %r15\n  popq %r14\n  popq %r13\n  popq %r12\n  popq %rbx\n  popq %rbp\n"
    "  ret\n");

void yield() {
  Fiber* f = g_cur;
  emu_switch(&f->sp, g_sched_sp);
}

static void trampoline() {
  g_block->body();
  Fiber* f = g_cur;
  f->done = true;
  Block& b = *g_block;
  b.alive--;
  ++g_progress;
  // a thread that exits releases a barrier the remaining threads are waiting on
  if (b.alive > 0 && b.bar_arrived >= b.alive) {
    b.bar_arrived = 0;
    b.bar_gen++;
  }
  emu_switch(&f->sp, g_sched_sp);
  abort();
}

static void run_block(Block& b) {
  static const bool shuffle = getenv("LACB_EMU_SHUFFLE") && atoi(getenv("LACB_EMU_SHUFFLE")) != 0;
  static uint32_t rng = 12345u;
  const unsigned n = b.nthreads;
  if ((size_t)n > g_stacks_n) {
    if (g_stacks) munmap(g_stacks, g_stacks_n * kStack);
    g_stacks_n = n;
    g_stacks = (unsigned char*)mmap(nullptr, g_stacks_n * kStack, PROT_READ | PROT_WRITE,
                                    MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
    if (g_stacks == MAP_FAILED) { perror("emu mmap"); abort(); }
  }
  b.fibers.assign(n, Fiber{});
  b.warps.assign((n + 31) / 32, WarpX{});
  b.alive = n;
  for (unsigned t = 0; t < n; ++t) {
    unsigned char* top = g_stacks + (size_t)(t + 1) * kStack;
    void** sp = (void**)top;
    *--sp = nullptr;              // fake return address of trampoline (keeps rsp%16==8 at entry)
    *--sp = (void*)&trampoline;   // `ret` target
    for (int r = 0; r < 6; ++r) *--sp = nullptr;
    b.fibers[t].sp = sp;
    b.fibers[t].tid = t;
  }
  std::vector<unsigned> order(n);
  for (unsigned t = 0; t < n; ++t) order[t] = t;
  uint64_t last_progress = g_progress;
  unsigned idle_rounds = 0;
  while (b.alive > 0) {
    if (shuffle) {
      for (unsigned t = n; t > 1; --t) {
        rng = rng * 1664525u + 1013904223u;
        std::swap(order[t - 1], order[(rng >> 8) % t]);
      }
    }
    for (unsigned oi = 0; oi < n; ++oi) {
      Fiber& f = b.fibers[order[oi]];
      if (f.done) continue;
      g_cur = &f;
      emu_switch(&g_sched_sp, f.sp);
    }
    if (g_progress == last_progress) {
      if (++idle_rounds > 4) {
        fprintf(stderr, "[cuda_emu] deadlock: block (%u,%u) alive=%u bar_arrived=%u\n", g_blockIdx.x,
                g_blockIdx.y, b.alive, b.bar_arrived);
        abort();
      }
    } else {
      idle_rounds = 0;
      last_progress = g_progress;
    }
  }
}

void launch(dim3 grid, dim3 block, size_t smem, const std::function<void()>& body) {
  if (!g_smem_buf) g_smem_buf = (unsigned char*)aligned_alloc(128, 256 * 1024);
  if (smem > 232448) { fprintf(stderr, "[cuda_emu] dynamic smem %zu too large\n", smem); abort(); }
  g_dyn_smem = g_smem_buf;
  Block b;
  b.nthreads = block.x * block.y * block.z;
  b.body = body;
  g_blockDim = block;
  g_gridDim = grid;
  Block* saved = g_block;
  g_block = &b;
  for (unsigned by = 0; by < grid.y; ++by)
    for (unsigned bx = 0; bx < grid.x; ++bx) {
      g_blockIdx = uint3{bx, by, 0};
      b.bar_arrived = 0;
      b.bar_gen = 0;
      b.bar_acc_or[0] = b.bar_acc_or[1] = 0;
      b.bar_acc_and[0] = b.bar_acc_and[1] = 1;
      b.bar_acc_cnt[0] = b.bar_acc_cnt[1] = 0;
      memset(g_smem_buf, 0xCD, smem);  // poison: uninitialised shared memory reads show up
      run_block(b);
    }
  g_block = saved;
}

}  // namespace emu
