// Host emulation of the small CUDA subset used by the SIMT (non-TMA, non-tcgen05) kernels -- TEST INFRASTRUCTURE.
// tests/test_kernel_emulation.py compiles a kernel source with -DZS3_HOST_EMULATION against this header (g++),
// runs ONE thread block as 256 host threads (pthread barriers for __syncthreads / warp shuffles) on host buffers and
// compares the result with the oracle.  It checks a kernel's indexing, phase ordering and arithmetic without a GPU;
// it says nothing about memory-model or performance behaviour, and nothing in zs3_b200/ uses it.
#pragma once
#include <pthread.h>
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <cmath>
#include <vector>

#include "../../include/zs3b200.h"

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __shared__ static
#define __align__(n) __attribute__((aligned(n)))
#define __launch_bounds__(...)

struct emul_dim3 { unsigned x, y, z; };
struct float4 { float x, y, z, w; } __attribute__((aligned(16)));
struct float2 { float x, y; } __attribute__((aligned(8)));
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }

struct EmulWarp { pthread_barrier_t bar; float fbuf[32]; int ibuf[32]; };
struct EmulBlock { pthread_barrier_t bar; std::vector<EmulWarp> warps; };

static thread_local emul_dim3 threadIdx, blockIdx, blockDim, gridDim;
static thread_local EmulBlock* emul_block;

static inline void __syncthreads() { pthread_barrier_wait(&emul_block->bar); }
static inline void __threadfence() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
static inline unsigned atomicAdd(unsigned* p, unsigned v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
static inline int atomicAdd(int* p, int v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
static inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) {
  return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST);
}
static inline int atomicMin(int* p, int v) {
  int old = __atomic_load_n(p, __ATOMIC_SEQ_CST);
  while (old > v && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {
  }
  return old;
}
template <class T> static inline T __ldcg(const T* p) { return *p; }
template <class T> static inline T __ldg(const T* p) { return *p; }
static inline int min(int a, int b) { return a < b ? a : b; }
static inline float __shfl_xor_sync(unsigned, float v, int off) {
  EmulWarp& w = emul_block->warps[threadIdx.x >> 5];
  const int lane = threadIdx.x & 31;
  w.fbuf[lane] = v;
  pthread_barrier_wait(&w.bar);
  const float r = w.fbuf[lane ^ off];
  pthread_barrier_wait(&w.bar);
  return r;
}
static inline unsigned emul_ld_acquire(const unsigned* p) { return __atomic_load_n(p, __ATOMIC_ACQUIRE); }

// Runs block `block` of a grid of `nblocks` as `threads` host threads.
template <class P>
static void emul_run_block(void (*kernel)(const P), const P& param, int threads, int block, int nblocks) {
  EmulBlock blk;
  pthread_barrier_init(&blk.bar, nullptr, threads);
  blk.warps.resize((threads + 31) / 32);
  for (size_t w = 0; w < blk.warps.size(); ++w)
    pthread_barrier_init(&blk.warps[w].bar, nullptr, std::min(32, threads - (int)w * 32));
  struct Arg { void (*k)(const P); const P* p; EmulBlock* b; int tid, n, blk, nblk; };
  std::vector<Arg> args(threads);
  std::vector<pthread_t> th(threads);
  auto entry = +[](void* v) -> void* {
    Arg* a = static_cast<Arg*>(v);
    threadIdx = emul_dim3{(unsigned)a->tid, 0, 0};
    blockIdx = emul_dim3{(unsigned)a->blk, 0, 0};
    blockDim = emul_dim3{(unsigned)a->n, 1, 1};
    gridDim = emul_dim3{(unsigned)a->nblk, 1, 1};
    emul_block = a->b;
    a->k(*a->p);
    return nullptr;
  };
  pthread_attr_t attr;
  pthread_attr_init(&attr);
  pthread_attr_setstacksize(&attr, 1 << 20);
  for (int t = 0; t < threads; ++t) {
    args[t] = Arg{kernel, &param, &blk, t, threads, block, nblocks};
    pthread_create(&th[t], &attr, entry, &args[t]);
  }
  for (int t = 0; t < threads; ++t) pthread_join(th[t], nullptr);
  pthread_attr_destroy(&attr);
}

// Runs kernel(param) on a grid of `nblocks` blocks.  One block runs in this process; several blocks run as forked
// child processes (their `__shared__` statics are then private per block, as on the device), so every buffer the
// kernel WRITES must live in MAP_SHARED memory (torch: tensor.share_memory_()).  Returns 0 on success.
#include <sys/wait.h>
#include <unistd.h>
template <class P>
static int emul_launch_grid(void (*kernel)(const P), const P& param, int threads, int nblocks) {
  if (nblocks <= 1) {
    emul_run_block<P>(kernel, param, threads, 0, 1);
    return 0;
  }
  std::vector<pid_t> pids;
  for (int b = 0; b < nblocks; ++b) {
    const pid_t pid = fork();
    if (pid < 0) return -1;
    if (pid == 0) {
      emul_run_block<P>(kernel, param, threads, b, nblocks);
      _exit(0);
    }
    pids.push_back(pid);
  }
  int rc = 0;
  for (pid_t pid : pids) {
    int status = 0;
    if (waitpid(pid, &status, 0) < 0 || !WIFEXITED(status) || WEXITSTATUS(status) != 0) rc = -1;
  }
  return rc;
}
