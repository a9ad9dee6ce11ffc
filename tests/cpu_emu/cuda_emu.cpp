// TEST INFRASTRUCTURE (see cuda_emu.hpp): block scheduler on fibers + an immediately-executing CUDA runtime.
#include "cuda_emu.hpp"

#include <omp.h>
#if !defined(__x86_64__)
#include <ucontext.h>
#endif

#include <cstdio>
#include <cstdlib>
#include <vector>

thread_local uint3 threadIdx, blockIdx;
thread_local dim3 blockDim, gridDim;

namespace emu
{
namespace
{
constexpr size_t STACK_BYTES = 256 << 10;
constexpr size_t DYN_SMEM_BYTES = 256 << 10;

struct Wait
{
  const unsigned* gen = nullptr;  // resume when *gen != seen
  unsigned seen = 0;
};

// Context switch.  x86-64: six callee-saved registers and the stack pointer, no system call (glibc's swapcontext
// saves the signal mask with a syscall per switch, which dominated the run time); elsewhere: ucontext.
#if defined(__x86_64__)
extern "C" void emu_switch(void** save_sp, void* load_sp);
asm(R"(
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
struct Context
{
  void* sp = nullptr;
};
inline void ctx_switch(Context& from, Context& to) { emu_switch(&from.sp, to.sp); }
inline void ctx_make(Context& c, char* stack, size_t bytes, void (*entry)())
{
  uintptr_t top = ((uintptr_t)stack + bytes) & ~(uintptr_t)15;
  void** p = (void**)(top - 16);
  *p = (void*)entry;            // emu_switch's `ret` lands here with rsp = top - 8, as after a call
  for (int i = 0; i < 6; i++) *--p = nullptr;
  c.sp = p;
}
#else
struct Context
{
  ucontext_t uc;
};
inline void ctx_switch(Context& from, Context& to) { swapcontext(&from.uc, &to.uc); }
inline void ctx_make(Context& c, char* stack, size_t bytes, void (*entry)())
{
  getcontext(&c.uc);
  c.uc.uc_stack.ss_sp = stack;
  c.uc.uc_stack.ss_size = bytes;
  c.uc.uc_link = nullptr;
  makecontext(&c.uc, entry, 0);
}
#endif

struct Fiber
{
  Context ctx;
  char* stack = nullptr;
  uint3 tid{};
  int lin = 0;
  bool finished = false;
  Wait wait;
};

struct WarpState
{
  int alive = 0, arrived = 0;
  unsigned gen = 0;
  uint64_t buf[2][32];
};

struct Block
{
  const std::function<void()>* body = nullptr;
  dim3 dim;
  int total = 0, next = 0;      // threads of the block, next one not yet started
  int alive = 0, arrived = 0;   // block barrier
  unsigned gen = 0;
  std::vector<WarpState> warps;
  std::vector<int> order;       // the order threads start (and resume) in: EMU_ORDER = forward | reverse | shuffle[:seed]
  Context main_ctx;
  Fiber* current = nullptr;
  std::vector<Fiber*> pool, running;
};

thread_local Block* tl_block = nullptr;
thread_local std::vector<Fiber*>* tl_pool = nullptr;
alignas(128) thread_local unsigned char tl_dyn_smem[DYN_SMEM_BYTES];

void yield_to_main(Block& b) { ctx_switch(b.current->ctx, b.main_ctx); }

void thread_exit(Block& b, int lin)
{
  // a thread that returns no longer takes part in barriers (CUDA: exited threads count as arrived)
  b.alive--;
  if (b.alive > 0 && b.arrived == b.alive)
  {
    b.arrived = 0;
    b.gen++;
  }
  WarpState& w = b.warps[lin / 32];
  w.alive--;
  if (w.alive > 0 && w.arrived == w.alive)
  {
    w.arrived = 0;
    w.gen++;
  }
}

void fiber_main()
{
  Block& b = *tl_block;
  Fiber* f = b.current;
  while (b.next < b.total)
  {
    const int lin = b.order[b.next++];
    f->lin = lin;
    f->tid.x = lin % b.dim.x;
    f->tid.y = (lin / b.dim.x) % b.dim.y;
    f->tid.z = lin / (b.dim.x * b.dim.y);
    threadIdx = f->tid;
    (*b.body)();
    thread_exit(b, lin);
  }
  f->finished = true;
  ctx_switch(f->ctx, b.main_ctx);
  std::abort();  // a finished fiber is never resumed
}

// Thread order inside a block.  A kernel that is correct on the device does not depend on it; a missing barrier
// around a shared-memory ring usually does, so the parity tests are run under several orders (tests/test_emu.py).
void fill_order(Block& b)
{
  static const char* mode = std::getenv("EMU_ORDER");
  b.order.resize(b.total);
  for (int i = 0; i < b.total; i++) b.order[i] = i;
  if (!mode || !std::strncmp(mode, "forward", 7)) return;
  if (!std::strncmp(mode, "reverse", 7))
  {
    for (int i = 0; i < b.total; i++) b.order[i] = b.total - 1 - i;
    return;
  }
  // shuffle[:seed] — warps stay together only by chance, as on the device nothing orders them
  static thread_local uint64_t state = 0;
  if (state == 0) state = 0x9E3779B97F4A7C15ull ^ (std::strlen(mode) > 8 ? std::strtoull(mode + 8, nullptr, 10) : 1ull);
  for (int i = b.total - 1; i > 0; i--)
  {
    state ^= state << 13; state ^= state >> 7; state ^= state << 17;
    std::swap(b.order[i], b.order[(int)(state % (uint64_t)(i + 1))]);
  }
}

Fiber* take_fiber(Block& b)
{
  Fiber* f;
  if (!b.pool.empty())
  {
    f = b.pool.back();
    b.pool.pop_back();
  }
  else
  {
    f = new Fiber;
    f->stack = (char*)std::malloc(STACK_BYTES);
  }
  ctx_make(f->ctx, f->stack, STACK_BYTES, fiber_main);
  f->finished = false;
  f->wait = Wait{};
  return f;
}

void resume(Block& b, Fiber* f)
{
  b.current = f;
  threadIdx = f->tid;
  ctx_switch(b.main_ctx, f->ctx);
  b.current = nullptr;
}

void run_block(Block& b)
{
  b.next = 0;
  b.alive = b.total;
  b.arrived = 0;
  b.gen = 0;
  const int nw = (b.total + 31) / 32;
  b.warps.assign(nw, WarpState{});
  for (int w = 0; w < nw; w++) b.warps[w].alive = std::min(32, b.total - 32 * w);
  b.running.clear();
  fill_order(b);
  while (true)
  {
    bool progressed = false;
    if (b.next < b.total)
    {
      Fiber* f = take_fiber(b);
      b.running.push_back(f);
      resume(b, f);
      progressed = true;
    }
    else
    {
      for (size_t i = 0; i < b.running.size(); i++)
      {
        Fiber* f = b.running[i];
        if (f->finished) continue;
        if (f->wait.gen && *f->wait.gen == f->wait.seen) continue;
        resume(b, f);
        progressed = true;
      }
    }
    size_t keep = 0;
    for (size_t i = 0; i < b.running.size(); i++)
    {
      if (b.running[i]->finished) b.pool.push_back(b.running[i]);
      else b.running[keep++] = b.running[i];
    }
    b.running.resize(keep);
    if (b.running.empty() && b.next >= b.total) break;
    if (!progressed)
    {
      std::fprintf(stderr, "cuda_emu: deadlock in block (%u,%u,%u): %zu threads wait on a barrier the others never reach\n", blockIdx.x,
                   blockIdx.y, blockIdx.z, b.running.size());
      std::abort();
    }
  }
}
}  // namespace

void* dyn_smem() { return tl_dyn_smem; }

void sync_threads()
{
  Block& b = *tl_block;
  if (++b.arrived == b.alive)
  {
    b.arrived = 0;
    b.gen++;
    return;
  }
  Fiber* f = b.current;
  f->wait.gen = &b.gen;
  f->wait.seen = b.gen;
  while (b.gen == f->wait.seen) yield_to_main(b);
  f->wait.gen = nullptr;
}

// a thread that polls for something another thread of its block will do (an mbarrier phase): let the others run
void spin_yield() { yield_to_main(*tl_block); }

static void warp_barrier(Block& b, WarpState& w)
{
  if (++w.arrived == w.alive)
  {
    w.arrived = 0;
    w.gen++;
    return;
  }
  Fiber* f = b.current;
  f->wait.gen = &w.gen;
  f->wait.seen = w.gen;
  while (w.gen == f->wait.seen) yield_to_main(b);
  f->wait.gen = nullptr;
}

uint64_t warp_exchange(uint64_t mine, int delta, bool up, int width)
{
  Block& b = *tl_block;
  Fiber* f = b.current;
  WarpState& w = b.warps[f->lin / 32];
  const int lane = f->lin % 32;
  const unsigned parity = w.gen & 1u;  // the generation this exchange completes in; the next one uses the other buffer
  w.buf[parity][lane] = mine;
  warp_barrier(b, w);
  const int seg = lane / width * width;
  const int src = up ? lane - delta : lane + delta;
  const int lanes = std::min(32, b.total - 32 * (f->lin / 32));
  if (src < seg || src >= seg + width || src >= lanes) return mine;  // out of range: own value (CUDA semantics)
  return w.buf[parity][src];
}

void run_grid(const LaunchCfg& cfg, const std::function<void()>& body)
{
  const long long nblocks = (long long)cfg.grid.x * cfg.grid.y * cfg.grid.z;
  const int total = (int)(cfg.block.x * cfg.block.y * cfg.block.z);
  if (nblocks <= 0 || total <= 0) return;
  if (cfg.smem > DYN_SMEM_BYTES)
  {
    std::fprintf(stderr, "cuda_emu: %zu bytes of dynamic shared memory requested\n", cfg.smem);
    std::abort();
  }
  const bool parallel = nblocks > 1 && !omp_in_parallel();
#pragma omp parallel if (parallel)
  {
    static thread_local Block blk;  // keeps its fiber pool between launches
    Block& b = blk;
    b.body = &body;
    b.dim = cfg.block;
    b.total = total;
    tl_block = &b;
    blockDim = cfg.block;
    gridDim = cfg.grid;
#pragma omp for schedule(dynamic, 1)
    for (long long i = 0; i < nblocks; i++)
    {
      blockIdx.x = (unsigned)(i % cfg.grid.x);
      blockIdx.y = (unsigned)((i / cfg.grid.x) % cfg.grid.y);
      blockIdx.z = (unsigned)(i / ((long long)cfg.grid.x * cfg.grid.y));
      run_block(b);
    }
    tl_block = nullptr;
  }
}
}  // namespace emu

// ------------------------------------------------------------------------------------------------
// CUDA runtime: one in-order device, everything executes at the call
// ------------------------------------------------------------------------------------------------
extern "C"
{
cudaError_t cudaGetDeviceCount(int* n) { *n = 2; return cudaSuccess; }
cudaError_t cudaSetDevice(int) { return cudaSuccess; }
cudaError_t cudaGetDevice(int* d) { *d = 0; return cudaSuccess; }
cudaError_t cudaDeviceSynchronize(void) { return cudaSuccess; }
cudaError_t cudaGetLastError(void) { return cudaSuccess; }
cudaError_t cudaPeekAtLastError(void) { return cudaSuccess; }
const char* cudaGetErrorString(cudaError_t e) { return e == cudaSuccess ? "no error" : "emulated CUDA runtime error"; }
cudaError_t cudaMalloc(void** p, size_t n)
{
  void* q = nullptr;
  if (posix_memalign(&q, 256, n ? n : 1) != 0) return cudaErrorMemoryAllocation;
  std::memset(q, 0xFF, n);  // NaN doubles, -1 ints: reads of never-written device memory show up in the parity tests
  *p = q;
  return cudaSuccess;
}
cudaError_t cudaFree(void* p) { std::free(p); return cudaSuccess; }
cudaError_t cudaMallocHost(void** p, size_t n) { return posix_memalign(p, 256, n ? n : 1) == 0 ? cudaSuccess : cudaErrorMemoryAllocation; }
cudaError_t cudaHostAlloc(void** p, size_t n, unsigned) { return cudaMallocHost(p, n); }
cudaError_t cudaFreeHost(void* p) { std::free(p); return cudaSuccess; }
cudaError_t cudaHostRegister(void*, size_t, unsigned) { return cudaSuccess; }
cudaError_t cudaHostUnregister(void*) { return cudaSuccess; }
cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { std::memmove(d, s, n); return cudaSuccess; }
cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t) { std::memmove(d, s, n); return cudaSuccess; }
cudaError_t cudaMemcpyPeerAsync(void* d, int, const void* s, int, size_t n, cudaStream_t) { std::memmove(d, s, n); return cudaSuccess; }
cudaError_t cudaMemcpy2DAsync(void* d, size_t dp, const void* s, size_t sp, size_t w, size_t h, cudaMemcpyKind, cudaStream_t)
{
  for (size_t r = 0; r < h; r++) std::memmove((char*)d + r * dp, (const char*)s + r * sp, w);
  return cudaSuccess;
}
cudaError_t cudaMemset(void* d, int v, size_t n) { std::memset(d, v, n); return cudaSuccess; }
cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t) { std::memset(d, v, n); return cudaSuccess; }
cudaError_t cudaStreamCreate(cudaStream_t* s) { *s = (cudaStream_t) new int(0); return cudaSuccess; }
cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { return cudaStreamCreate(s); }
cudaError_t cudaStreamCreateWithPriority(cudaStream_t* s, unsigned, int) { return cudaStreamCreate(s); }
cudaError_t cudaStreamDestroy(cudaStream_t s) { delete (int*)s; return cudaSuccess; }
cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return cudaSuccess; }
cudaError_t cudaDeviceGetStreamPriorityRange(int* lo, int* hi) { *lo = 0; *hi = -1; return cudaSuccess; }
cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = (cudaEvent_t) new int(0); return cudaSuccess; }
cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { return cudaEventCreate(e); }
cudaError_t cudaEventDestroy(cudaEvent_t e) { delete (int*)e; return cudaSuccess; }
cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t) { return cudaSuccess; }
cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
cudaError_t cudaEventQuery(cudaEvent_t) { return cudaSuccess; }
cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t, cudaEvent_t) { *ms = 0.0f; return cudaSuccess; }
cudaError_t cudaDeviceCanAccessPeer(int* can, int, int) { *can = 1; return cudaSuccess; }
cudaError_t cudaDeviceEnablePeerAccess(int, unsigned) { return cudaSuccess; }
cudaError_t cudaFuncSetAttribute(const void*, cudaFuncAttribute, int) { return cudaSuccess; }
// a small "device" on purpose: the band-height rule (pick_band_rows) then cuts even the tests' grids into several bands
// a small "device" on purpose: persistent launches then stride over their tiles even on the tests' grids
cudaError_t cudaOccupancyMaxActiveBlocksPerMultiprocessor(int* n, const void*, int, size_t) { *n = 3; return cudaSuccess; }
cudaError_t cudaDeviceGetAttribute(int* v, cudaDeviceAttr a, int)
{
  *v = a == cudaDevAttrMultiProcessorCount ? 3 : a == cudaDevAttrMaxRegistersPerMultiprocessor ? 65536 : a == cudaDevAttrMaxSharedMemoryPerMultiprocessor ? 233472 : 0;
  return cudaSuccess;
}
cudaError_t cudaFuncGetAttributes(cudaFuncAttributes* fa, const void*)
{
  std::memset(fa, 0, sizeof(*fa));
  fa->numRegs = 160;
  return cudaSuccess;
}
// stream capture is not emulated: lbm_use_graph reports the error, plain stepping is what the emulated tests run
cudaError_t cudaStreamBeginCapture(cudaStream_t, cudaStreamCaptureMode) { return cudaErrorNotSupported; }
cudaError_t cudaStreamEndCapture(cudaStream_t, cudaGraph_t* g) { *g = nullptr; return cudaErrorNotSupported; }
cudaError_t cudaGraphInstantiate(cudaGraphExec_t* e, cudaGraph_t, unsigned long long) { *e = nullptr; return cudaErrorNotSupported; }
cudaError_t cudaGraphLaunch(cudaGraphExec_t, cudaStream_t) { return cudaErrorNotSupported; }
cudaError_t cudaGraphDestroy(cudaGraph_t) { return cudaSuccess; }
cudaError_t cudaGraphExecDestroy(cudaGraphExec_t) { return cudaSuccess; }
}
