// TEST INFRASTRUCTURE ONLY: scheduler of the cooperative-fiber CUDA emulator (see include/cuda_runtime.h).
#include <cuda_runtime.h>

// the runtime entry points the host code of the kernels' translation units calls
extern "C" cudaError_t cudaGetLastError(void) { return cudaSuccess; }
extern "C" const char* cudaGetErrorString(cudaError_t) { return "emulated"; }
extern "C" cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, enum cudaMemcpyKind, cudaStream_t) { memcpy(d, s, n); return cudaSuccess; }
extern "C" cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t) { memset(d, v, n); return cudaSuccess; }
extern "C" cudaError_t cudaGetDevice(int* d) { *d = 0; return cudaSuccess; }
extern "C" cudaError_t cudaDeviceGetAttribute(int* v, enum cudaDeviceAttr, int) { *v = 4; return cudaSuccess; }

uint3 threadIdx, blockIdx;
dim3 blockDim, gridDim;

namespace emu {
ucontext_t g_sched;
Fiber* g_cur = nullptr;
std::vector<Fiber>* g_fibers = nullptr;
Barrier g_warp_bar[64], g_block_bar;
unsigned long long g_slot[64][32];
int g_block_threads = 0;
char* dyn_smem = nullptr;
std::function<void()> g_body;
long long g_progress = 0;
unsigned g_mma_a[64][32][4], g_mma_b[64][32][2];
float g_tmem[128][512];
Barrier g_named_bar[16];

static void trampoline() {
  g_body();
  g_cur->done = true;
  ++g_progress;
  swapcontext(&g_cur->ctx, &g_sched);
}

void run_block(std::function<void()> body, dim3 block) {
  const int n = (int)(block.x * block.y * block.z);
  if (n > 2048) { fprintf(stderr, "emu: block of %d threads\n", n); abort(); }
  g_block_threads = n;
  g_body = body;
  for (auto& b : g_warp_bar) b = Barrier();
  g_block_bar = Barrier();
  for (auto& b : g_named_bar) b = Barrier();
  std::vector<Fiber> fibers(n);
  g_fibers = &fibers;
  for (int t = 0; t < n; ++t) {
    Fiber& f = fibers[t];
    f.stack.resize(256 * 1024);
    f.tid = uint3{(unsigned)(t % block.x), (unsigned)((t / block.x) % block.y), (unsigned)(t / (block.x * block.y))};
    getcontext(&f.ctx);
    f.ctx.uc_stack.ss_sp = f.stack.data();
    f.ctx.uc_stack.ss_size = f.stack.size();
    f.ctx.uc_link = &g_sched;
    makecontext(&f.ctx, trampoline, 0);
  }
  int left = n, stalled = 0;
  while (left > 0) {
    const long long before = g_progress;
    for (int t = 0; t < n; ++t) {
      Fiber& f = fibers[t];
      if (f.done) continue;
      g_cur = &f;
      threadIdx = f.tid;
      swapcontext(&g_sched, &f.ctx);
      if (f.done) --left;
    }
    // A pass without a single arrival / completion is not yet a deadlock (a thread may have moved silently from one wait to the
    // next), but thousands in a row are: everybody spins on something nobody will ever signal.
    stalled = (g_progress == before) ? stalled + 1 : 0;
    if (left > 0 && stalled > 4000) {
      fprintf(stderr, "emu: deadlock -- %d threads of block (%u,%u,%u) wait for something the others never signal\n", left, blockIdx.x, blockIdx.y, blockIdx.z);
      abort();
    }
  }
  g_fibers = nullptr;
}
}  // namespace emu
