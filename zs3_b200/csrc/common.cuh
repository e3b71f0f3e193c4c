// Host-side helpers shared by the C-ABI translation units: error reporting, launch checks,
// TMA tensor-map encoding through the driver entry points (no link-time libcuda dependency).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

#include "../../include/zs3b200.h"

namespace zs3 {

void set_error(const char* fmt, ...);

#define ZS3_CHECK_ARG(cond, ...)           \
  do {                                     \
    if (!(cond)) {                         \
      zs3::set_error(__VA_ARGS__);         \
      return ZS3_ERR_INVALID_ARG;          \
    }                                      \
  } while (0)

extern unsigned long long g_launch_count;  // kernels launched by this library (bench.py's gpu_launches)

#define ZS3_CHECK_LAUNCH(name)                                                  \
  do {                                                                          \
    ++zs3::g_launch_count;                                                      \
    cudaError_t e__ = cudaGetLastError();                                       \
    if (e__ != cudaSuccess) {                                                   \
      zs3::set_error("%s: launch failed: %s", name, cudaGetErrorString(e__));   \
      return ZS3_ERR_LAUNCH;                                                    \
    }                                                                           \
  } while (0)

// Programmatic dependent launch: a kernel launched through launch_pdl() may be scheduled while its predecessor in
// the stream is still draining (the predecessor executes griddepcontrol.launch_dependents, or finishes); the kernel
// itself must execute griddepcontrol.wait (ptx.cuh: griddep_sync()) before its first global-memory access.  This
// hides launch latency and kernel prologues (barrier init, TMEM allocation, descriptor prefetch) behind the previous
// kernel's tail; the edges survive CUDA-graph capture.  Measured on B200 inside the captured training step
// (profiles/r01_pdl_experiment.md): +1.9 % with the attribute on the conv kernels (long prologues), -1.1 % on the
// BatchNorm streaming kernels (no prologue to hide), so the default is ZS3_PDL=1: 1 = conv kernels, 2 = BatchNorm
// kernels, 3 = both, 0 = off.
constexpr int PDL_CONV = 1, PDL_BN = 2;
bool pdl_enabled(int kind);

template <typename... KArgs, typename... Args>
static inline cudaError_t launch_pdl(int kind, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem,
                                     cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled(kind) ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
static inline long long ceil_div_ll(long long a, long long b) { return (a + b - 1) / b; }

// bf16, SWIZZLE_128B tensor maps.  All return 0 on success.
// im2col map over an NHWC activation tensor [N][H][W][C] (C = channel stride, elements).
// oob_nan (reserved, no caller passes it yet): out-of-bounds elements (zero padding, rows past the tensor) are filled with
// NaN instead of zero -- meant for an operand prologue that applies max(scale*x + shift, 0) in shared memory, where
// max(NaN, 0) = 0 restores the padding
int encode_im2col_bf16(CUtensorMap* out, const void* base, int N, int H, int W, int C, int pad_lo, int upper_corner,
                       int stride, int channels_per_pixel, int pixels_per_column, int oob_nan = 0);
// tiled 2-D map over a row-major [rows][cols] bf16 matrix with row stride ld (elements); box = [box_rows][box_cols].
int encode_tiled2d_bf16(CUtensorMap* out, const void* base, long long rows, int cols, long long ld, int box_rows,
                        int box_cols);
// same with SWIZZLE_64B and a 32-column (64-byte) box: the epilogue's TMA store tiles
int encode_tiled2d_bf16_sw64(CUtensorMap* out, const void* base, long long rows, int cols, long long ld, int box_rows,
                             int box_cols);
// fp32 3-D map (SWIZZLE_128B) over a strided [d2][d1][d0] view (strides in ELEMENTS, d0 contiguous): the weight-gradient
// tensor the wgrad epilogue reduces into with cp.reduce.async.bulk (.add)
int encode_tiled3d_f32(CUtensorMap* out, const void* base, int d2, int d1, int d0, long long stride2, long long stride1,
                       int b2, int b1, int b0);
// tiled 3-D map over [d2][d1][d0] bf16 (d0 contiguous); box = [b2][b1][b0].
int encode_tiled3d_bf16(CUtensorMap* out, const void* base, int d2, int d1, int d0, int b2, int b1, int b0);

}  // namespace zs3
