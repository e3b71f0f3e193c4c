// Error plumbing + TMA tensor-map encoders (driver entry points resolved at run time).
#include "common.cuh"

#include <cudaTypedefs.h>
#include <mutex>
#include <stdlib.h>
#include <string.h>

namespace zs3 {

static thread_local char g_err[512] = {0};
unsigned long long g_launch_count = 0;

bool pdl_enabled(int kind) {
  static int pref = -1;
  if (pref < 0) {
    const char* e = getenv("ZS3_PDL");
    pref = e ? atoi(e) : PDL_CONV;
    if (pref < 0 || pref > 3) pref = PDL_CONV;
  }
  return (pref & kind) != 0;
}

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
typedef CUresult (*EncodeIm2colFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const int*, const int*, cuuint32_t, cuuint32_t,
                                   const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn g_encode_tiled = nullptr;
static EncodeIm2colFn g_encode_im2col = nullptr;
static int g_driver_version = 0;
static std::once_flag g_once;

static void resolve_entry_points() {
  cudaDriverEntryPointQueryResult q;
  void* fn = nullptr;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) == cudaSuccess &&
      q == cudaDriverEntryPointSuccess)
    g_encode_tiled = reinterpret_cast<EncodeTiledFn>(fn);
  fn = nullptr;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeIm2col", &fn, cudaEnableDefault, &q) == cudaSuccess &&
      q == cudaDriverEntryPointSuccess)
    g_encode_im2col = reinterpret_cast<EncodeIm2colFn>(fn);
  cudaDriverGetVersion(&g_driver_version);
}

static int ensure_driver() {
  std::call_once(g_once, resolve_entry_points);
  if (!g_encode_tiled || !g_encode_im2col) {
    set_error("cuTensorMapEncode{Tiled,Im2col} driver entry points unavailable (no CUDA driver?)");
    return ZS3_ERR_DRIVER;
  }
  return ZS3_OK;
}

int encode_im2col_bf16(CUtensorMap* out, const void* base, int N, int H, int W, int C, int pad_lo, int upper_corner,
                       int stride, int channels_per_pixel, int pixels_per_column, int oob_nan) {
  int rc = ensure_driver();
  if (rc) return rc;
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
  int lower[2] = {-pad_lo, -pad_lo};                // {W, H}
  int upper[2] = {upper_corner, upper_corner};      // pad_hi - (filter-1)*dil
  cuuint32_t estr[4] = {1, (cuuint32_t)stride, (cuuint32_t)stride, 1};
  CUresult r = g_encode_im2col(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides,
                               lower, upper, (cuuint32_t)channels_per_pixel, (cuuint32_t)pixels_per_column, estr,
                               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                               CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                               oob_nan ? CU_TENSOR_MAP_FLOAT_OOB_FILL_NAN_REQUEST_ZERO_FMA : CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeIm2col failed (%d): N=%d H=%d W=%d C=%d pad=%d upper=%d stride=%d cpp=%d ppc=%d",
              (int)r, N, H, W, C, pad_lo, upper_corner, stride, channels_per_pixel, pixels_per_column);
    return ZS3_ERR_DRIVER;
  }
  // Known driver issue (<= 13.1) for im2col maps over tensors smaller than 128 KiB: bit 21 of the second
  // 64-bit descriptor word must be cleared (same workaround as CUTLASS's make_im2col_tma_copy_desc).
  if (g_driver_version <= 13010) {
    unsigned long long bytes = (unsigned long long)N * H * W * C * 2ull;
    if (bytes < 131072ull) reinterpret_cast<uint64_t*>(out)[1] &= ~(1ull << 21);
  }
  return ZS3_OK;
}

int encode_tiled2d_bf16(CUtensorMap* out, const void* base, long long rows, int cols, long long ld, int box_rows,
                        int box_cols) {
  int rc = ensure_driver();
  if (rc) return rc;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = g_encode_tiled(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box,
                              estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                              CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(2d) failed (%d): rows=%lld cols=%d ld=%lld box=%dx%d", (int)r, rows, cols, ld,
              box_rows, box_cols);
    return ZS3_ERR_DRIVER;
  }
  return ZS3_OK;
}

int encode_tiled2d_bf16_sw64(CUtensorMap* out, const void* base, long long rows, int cols, long long ld, int box_rows,
                             int box_cols) {
  int rc = ensure_driver();
  if (rc) return rc;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = g_encode_tiled(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box,
                              estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B,
                              CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(2d, sw64) failed (%d): rows=%lld cols=%d ld=%lld box=%dx%d", (int)r, rows, cols,
              ld, box_rows, box_cols);
    return ZS3_ERR_DRIVER;
  }
  return ZS3_OK;
}

int encode_tiled3d_f32(CUtensorMap* out, const void* base, int d2, int d1, int d0, long long stride2, long long stride1,
                       int b2, int b1, int b0) {
  int rc = ensure_driver();
  if (rc) return rc;
  cuuint64_t dims[3] = {(cuuint64_t)d0, (cuuint64_t)d1, (cuuint64_t)d2};
  cuuint64_t strides[2] = {(cuuint64_t)stride1 * 4, (cuuint64_t)stride2 * 4};
  cuuint32_t box[3] = {(cuuint32_t)b0, (cuuint32_t)b1, (cuuint32_t)b2};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = g_encode_tiled(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<void*>(base), dims, strides, box,
                              estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                              CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(3d f32) failed (%d): dims=%d,%d,%d strides=%lld,%lld box=%d,%d,%d", (int)r, d2, d1,
              d0, stride2, stride1, b2, b1, b0);
    return ZS3_ERR_DRIVER;
  }
  return ZS3_OK;
}

int encode_tiled3d_bf16(CUtensorMap* out, const void* base, int d2, int d1, int d0, int b2, int b1, int b0) {
  int rc = ensure_driver();
  if (rc) return rc;
  cuuint64_t dims[3] = {(cuuint64_t)d0, (cuuint64_t)d1, (cuuint64_t)d2};
  cuuint64_t strides[2] = {(cuuint64_t)d0 * 2, (cuuint64_t)d0 * d1 * 2};
  cuuint32_t box[3] = {(cuuint32_t)b0, (cuuint32_t)b1, (cuuint32_t)b2};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = g_encode_tiled(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box,
                              estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                              CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(3d) failed (%d): dims=%d,%d,%d box=%d,%d,%d", (int)r, d2, d1, d0, b2, b1, b0);
    return ZS3_ERR_DRIVER;
  }
  return ZS3_OK;
}

}  // namespace zs3

extern "C" {

const char* zs3_last_error(void) { return zs3::g_err; }

int zs3_abi_version(void) { return 1; }

unsigned long long zs3_sizeof(int which) {
  switch (which) {
    case ZS3_STRUCT_CONV_ARGS: return sizeof(zs3_conv_args);
    case ZS3_STRUCT_WGRAD_ARGS: return sizeof(zs3_wgrad_args);
    case ZS3_STRUCT_BN_APPLY_ARGS: return sizeof(zs3_bn_apply_args);
    case ZS3_STRUCT_BN_BWD_ARGS: return sizeof(zs3_bn_bwd_args);
    case ZS3_STRUCT_SGEMM_ARGS: return sizeof(zs3_sgemm_args);
    case ZS3_STRUCT_GMMN_ITEM: return sizeof(zs3_gmmn_item);
    case ZS3_STRUCT_GMMN_TRAIN_ARGS: return sizeof(zs3_gmmn_train_args);
    case ZS3_STRUCT_COMPONENTS_ARGS: return sizeof(zs3_components_args);
    case ZS3_STRUCT_CONV_SEGMENT: return sizeof(zs3_conv_segment);
    case ZS3_STRUCT_ROW_SOURCE: return sizeof(zs3_row_source);
    case ZS3_STRUCT_BN_ACT_F32_ARGS: return sizeof(zs3_bn_act_f32_args);
    case ZS3_STRUCT_BN_BWD_F32_ARGS: return sizeof(zs3_bn_bwd_f32_args);
    case ZS3_STRUCT_AUG_ITEM: return sizeof(zs3_aug_item);
    case ZS3_STRUCT_AUGMENT_ARGS: return sizeof(zs3_augment_args);
    default: return 0;
  }
}

unsigned long long zs3_launch_count(void) { return zs3::g_launch_count; }

int zs3_device_supported(void) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  int major = 0;
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return 0;
  return major == 10 ? 1 : 0;
}
}
