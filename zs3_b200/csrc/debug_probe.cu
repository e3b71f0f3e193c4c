// Debug-only entry point: issue ONE im2col TMA load and dump the raw (still swizzled) shared-memory tile.
// tests/ use it to pin the TMA im2col addressing semantics the conv kernels rely on.
#include "common.cuh"
#include "ptx.cuh"

namespace zs3 {

struct ProbeParams {
  CUtensorMap map;
  int c, w, h, n, off_w, off_h;
  int bytes;
  uint8_t* out;
};

__global__ void im2col_probe_kernel(const __grid_constant__ ProbeParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    fence_mbar_init();
  }
  for (int i = threadIdx.x; i < p.bytes / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0xFFFFFFFFu;
  fence_proxy_async_smem();
  __syncthreads();
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(&bar, p.bytes);
    tma_load_im2col_4d(smem, &p.map, &bar, p.c, p.w, p.h, p.n, (uint16_t)p.off_w, (uint16_t)p.off_h);
  }
  mbar_wait(&bar, 0);
  __syncthreads();
  for (int i = threadIdx.x; i < p.bytes / 4; i += blockDim.x)
    reinterpret_cast<uint32_t*>(p.out)[i] = reinterpret_cast<uint32_t*>(smem)[i];
}

}  // namespace zs3

using namespace zs3;

// x: bf16 NHWC [N][H][W][C]; loads `ppc` pixels x `cpp` channels starting at base pixel (w,h,n), channel c,
// with tap offset (off_w, off_h); out receives ppc*cpp*2 raw bytes of shared memory.
extern "C" int zs3_debug_im2col_probe(const void* x, int N, int H, int W, int C, int pad, int upper, int stride, int cpp,
                                      int ppc, int c, int w, int h, int n, int off_w, int off_h, void* out,
                                      void* stream) {
  ProbeParams p;
  memset(&p, 0, sizeof(p));
  int rc = encode_im2col_bf16(&p.map, x, N, H, W, C, pad, upper, stride, cpp, ppc);
  if (rc) return rc;
  p.c = c;
  p.w = w;
  p.h = h;
  p.n = n;
  p.off_w = off_w;
  p.off_h = off_h;
  p.bytes = ppc * cpp * 2;
  p.out = static_cast<uint8_t*>(out);
  const int smem_bytes = p.bytes + 1024;
  cudaFuncSetAttribute(im2col_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
  im2col_probe_kernel<<<1, 128, smem_bytes, static_cast<cudaStream_t>(stream)>>>(p);
  ZS3_CHECK_LAUNCH("im2col_probe");
  return ZS3_OK;
}
