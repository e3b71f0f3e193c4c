// fp32-grade "parity mode" support kernels (forward only).
//
// The tcgen05 conv kernel multiplies bf16 operands exactly and accumulates in fp32, so an fp32-grade convolution is
// obtained by splitting both operands into three bf16 pieces (x = hi + mid + lo, 24 mantissa bits in total) and
// reducing the six significant cross products as six K-segments of ONE accumulator (zs3_conv_fprop's segment
// mechanism).  Activations then have to stay fp32 between layers; this file provides the fp32-I/O versions of the
// HBM-bound glue kernels (BatchNorm apply, max-pool, bilinear, pooling, logits upsample, stem im2col) and the
// splitter.  Used by tests/test_deeplab_parity.py to meet the 1e-3 logits tolerance of BASELINE.json's north star;
// the bf16 kernels remain the throughput path.
#include "common.cuh"
#include "ptx.cuh"

namespace zs3 {

static int pblocks(long long items, int threads, int cap = 148 * 16) {
  long long b = (items + threads - 1) / threads;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

// x = hi + mid + lo with each piece exactly representable in bf16 (|lo| <= 2^-16 |x|, dropped tail <= 2^-24 |x|)
__global__ void split3_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ hi,
                              __nv_bfloat16* __restrict__ mid, __nv_bfloat16* __restrict__ lo, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float v = x[i];
    const __nv_bfloat16 h = __float2bfloat16(v);
    const float r1 = v - __bfloat162float(h);
    const __nv_bfloat16 m = __float2bfloat16(r1);
    const float r2 = r1 - __bfloat162float(m);
    hi[i] = h;
    mid[i] = m;
    lo[i] = __float2bfloat16(r2);
  }
}

// one bf16 component (0 hi, 1 mid, 2 lo) of an fp32 OIHW/KRSC weight, written in the packed fprop layout
__global__ void pack_weight_component_kernel(const float* __restrict__ w, int Cout, int Cin, int taps, int ci_begin,
                                             int ci_count, __nv_bfloat16* __restrict__ dst, int cout_pad, int cin_pad,
                                             int src_krsc, int component) {
  const long long total = (long long)cout_pad * taps * cin_pad;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int ci = (int)(i % cin_pad);
    const int tap = (int)((i / cin_pad) % taps);
    const int co = (int)(i / ((long long)cin_pad * taps));
    float v = 0.f;
    if (co < Cout && ci < ci_count)
      v = src_krsc ? w[((long long)co * taps + tap) * Cin + ci_begin + ci]
                   : w[((long long)co * Cin + ci_begin + ci) * taps + tap];
    __nv_bfloat16 c = __float2bfloat16(v);
    for (int k = 0; k < component; ++k) {
      v -= __bfloat162float(c);
      c = __float2bfloat16(v);
    }
    dst[i] = c;
  }
}

__global__ void bn_apply_f32_kernel(const float* __restrict__ y, long long y_cs, const float* __restrict__ res,
                                    long long res_cs, float* __restrict__ out, long long out_cs,
                                    const float* __restrict__ scale, const float* __restrict__ shift, long long M, int C,
                                    int relu) {
  const long long total = M * C;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long m = i / C;
    const int c = (int)(i - m * C);
    float v = fmaf(y[m * y_cs + c], scale[c], shift[c]);
    if (res) v += res[m * res_cs + c];
    if (relu) v = fmaxf(v, 0.f);
    out[m * out_cs + c] = v;
  }
}

__global__ void maxpool_f32_kernel(const float* __restrict__ x, float* __restrict__ y, int N, int H, int W, int C,
                                   int Ho, int Wo, int k, int stride, int pad) {
  const long long total = (long long)N * Ho * Wo * C;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const long long m = i / C;
    const int q = (int)(m % Wo), p = (int)((m / Wo) % Ho), n = (int)(m / ((long long)Wo * Ho));
    float best = -INFINITY;
    for (int r = 0; r < k; ++r) {
      const int ih = p * stride - pad + r;
      if (ih < 0 || ih >= H) continue;
      for (int s = 0; s < k; ++s) {
        const int iw = q * stride - pad + s;
        if (iw < 0 || iw >= W) continue;
        best = fmaxf(best, x[(((long long)n * H + ih) * W + iw) * C + c]);
      }
    }
    y[i] = best;
  }
}

__device__ __forceinline__ void blc(int o, float scale, int in_size, int& i0, int& i1, float& l1) {
  const float r = scale * o;
  i0 = (int)r;
  if (i0 > in_size - 1) i0 = in_size - 1;
  i1 = i0 + (i0 < in_size - 1 ? 1 : 0);
  l1 = r - i0;
}

// NHWC fp32 -> NHWC fp32 (to_nchw = 0) or NCHW fp32 (to_nchw = 1, only channels < C_out are written)
__global__ void bilinear_f32_kernel(const float* __restrict__ x, float* __restrict__ y, int N, int Hi, int Wi, int Ho,
                                    int Wo, int Cs, int C_out, float sh, float sw, int to_nchw) {
  const long long total = (long long)N * Ho * Wo * C_out;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    int c, ow, oh, n;
    if (to_nchw) {
      ow = (int)(i % Wo); oh = (int)((i / Wo) % Ho); c = (int)((i / ((long long)Wo * Ho)) % C_out);
      n = (int)(i / ((long long)Wo * Ho * C_out));
    } else {
      c = (int)(i % C_out); ow = (int)((i / C_out) % Wo); oh = (int)((i / ((long long)C_out * Wo)) % Ho);
      n = (int)(i / ((long long)C_out * Wo * Ho));
    }
    int y0, y1, x0, x1;
    float ly, lx;
    blc(oh, sh, Hi, y0, y1, ly);
    blc(ow, sw, Wi, x0, x1, lx);
    const float* b = x + (long long)n * Hi * Wi * Cs + c;
    const float v00 = b[((long long)y0 * Wi + x0) * Cs], v01 = b[((long long)y0 * Wi + x1) * Cs];
    const float v10 = b[((long long)y1 * Wi + x0) * Cs], v11 = b[((long long)y1 * Wi + x1) * Cs];
    // same operation order as ATen's upsample_bilinear2d: h0lambda*(w0lambda*a + w1lambda*b) + h1lambda*(...)
    const float v = (1.f - ly) * ((1.f - lx) * v00 + lx * v01) + ly * ((1.f - lx) * v10 + lx * v11);
    if (to_nchw)
      y[i] = v;
    else
      y[(((long long)n * Ho + oh) * Wo + ow) * Cs + c] = v;
  }
}

// y[n][c] = scale * sum_hw x[n][hw][c]
__global__ void spatial_sum_f32_kernel(const float* __restrict__ x, float* __restrict__ y, int HW, int C, float scale) {
  const int n = blockIdx.y, c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  double acc = 0.0;
  for (int p = 0; p < HW; ++p) acc += (double)x[((long long)n * HW + p) * C + c];
  y[(long long)n * C + c] = (float)(acc * scale);
}

__global__ void spatial_broadcast_f32_kernel(const float* __restrict__ x, float* __restrict__ y, int N, int HW, int C) {
  const long long total = (long long)N * HW * C;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const int n = (int)(i / ((long long)HW * C));
    y[i] = x[(long long)n * C + c];
  }
}

// NCHW fp32 -> NHWC fp32 with zero-padded channel stride, and the stem im2col in fp32
__global__ void nchw_to_nhwc_f32_kernel(const float* __restrict__ src, float* __restrict__ dst, int N, int C,
                                        long long HW, int cs) {
  const long long total = (long long)N * HW * cs;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % cs);
    const long long hw = (i / cs) % HW;
    const int n = (int)(i / ((long long)cs * HW));
    dst[i] = c < C ? src[((long long)n * C + c) * HW + hw] : 0.f;
  }
}

__global__ void stem_im2col_f32_kernel(const float* __restrict__ x, float* __restrict__ cols, int N, int C, int H,
                                       int W, int R, int stride, int pad, int Ho, int Wo, int kpad, int krsc) {
  const int K = C * R * R;
  const long long total = (long long)N * Ho * Wo * kpad;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int k = (int)(i % kpad);
    const long long m = i / kpad;
    float v = 0.f;
    if (k < K) {
      const int c = krsc ? k % C : k / (R * R);
      const int rs = krsc ? k / C : k - c * R * R;
      const int r = rs / R, s = rs - r * R;
      const int q = (int)(m % Wo), p = (int)((m / Wo) % Ho), n = (int)(m / ((long long)Wo * Ho));
      const int ih = p * stride - pad + r, iw = q * stride - pad + s;
      if (ih >= 0 && ih < H && iw >= 0 && iw < W) v = x[(((long long)n * C + c) * H + ih) * W + iw];
    }
    cols[i] = v;
  }
}

}  // namespace zs3

using namespace zs3;
#define ST(s) static_cast<cudaStream_t>(s)
#define BF(p) static_cast<__nv_bfloat16*>(p)

extern "C" int zs3_split3_f32(const float* x, void* hi, void* mid, void* lo, long long n, void* stream) {
  ZS3_CHECK_ARG(x && hi && mid && lo && n >= 0, "split3: bad args");
  if (n == 0) return ZS3_OK;
  split3_kernel<<<pblocks(n, 256), 256, 0, ST(stream)>>>(x, BF(hi), BF(mid), BF(lo), n);
  ZS3_CHECK_LAUNCH("split3");
  return ZS3_OK;
}

extern "C" int zs3_pack_weight_component(const float* w, int Cout, int Cin, int R, int S, int ci_begin, int ci_count,
                                         void* dst, int cout_pad, int cin_pad, int src_krsc, int component,
                                         void* stream) {
  ZS3_CHECK_ARG(w && dst && component >= 0 && component <= 2 && Cout <= cout_pad && ci_count <= cin_pad &&
                    ci_begin >= 0 && ci_begin + ci_count <= Cin,
                "pack_weight_component: bad args");
  const long long total = (long long)cout_pad * R * S * cin_pad;
  pack_weight_component_kernel<<<pblocks(total, 256), 256, 0, ST(stream)>>>(w, Cout, Cin, R * S, ci_begin, ci_count,
                                                                            BF(dst), cout_pad, cin_pad, src_krsc,
                                                                            component);
  ZS3_CHECK_LAUNCH("pack_weight_component");
  return ZS3_OK;
}

extern "C" int zs3_bn_apply_f32(const float* y, int y_cs, const float* residual, int res_cs, float* out, int out_cs,
                                const float* scale, const float* shift, long long M, int C, int relu, void* stream) {
  ZS3_CHECK_ARG(y && out && scale && shift && C > 0, "bn_apply_f32: bad args");
  if (M <= 0) return ZS3_OK;
  bn_apply_f32_kernel<<<pblocks(M * C, 256), 256, 0, ST(stream)>>>(y, y_cs, residual, res_cs, out, out_cs, scale,
                                                                   shift, M, C, relu);
  ZS3_CHECK_LAUNCH("bn_apply_f32");
  return ZS3_OK;
}

extern "C" int zs3_maxpool_f32(const float* x, float* y, int N, int H, int W, int C, int Ho, int Wo, int k, int stride,
                               int pad, void* stream) {
  ZS3_CHECK_ARG(x && y, "maxpool_f32: bad args");
  maxpool_f32_kernel<<<pblocks((long long)N * Ho * Wo * C, 256), 256, 0, ST(stream)>>>(x, y, N, H, W, C, Ho, Wo, k,
                                                                                       stride, pad);
  ZS3_CHECK_LAUNCH("maxpool_f32");
  return ZS3_OK;
}

static float blscale(int in, int out) { return out > 1 ? (float)(in - 1) / (float)(out - 1) : 0.f; }

extern "C" int zs3_bilinear_f32(const float* x, float* y, int N, int Hi, int Wi, int Ho, int Wo, int Cs, int C_out,
                                int to_nchw, void* stream) {
  ZS3_CHECK_ARG(x && y && C_out <= Cs, "bilinear_f32: bad args");
  bilinear_f32_kernel<<<pblocks((long long)N * Ho * Wo * C_out, 256), 256, 0, ST(stream)>>>(
      x, y, N, Hi, Wi, Ho, Wo, Cs, C_out, blscale(Hi, Ho), blscale(Wi, Wo), to_nchw);
  ZS3_CHECK_LAUNCH("bilinear_f32");
  return ZS3_OK;
}

extern "C" int zs3_spatial_sum_f32(const float* x, float* y, int N, int HW, int C, float scale, void* stream) {
  ZS3_CHECK_ARG(x && y, "spatial_sum_f32: bad args");
  spatial_sum_f32_kernel<<<dim3((C + 127) / 128, N), 128, 0, ST(stream)>>>(x, y, HW, C, scale);
  ZS3_CHECK_LAUNCH("spatial_sum_f32");
  return ZS3_OK;
}

extern "C" int zs3_spatial_broadcast_f32(const float* x, float* y, int N, int HW, int C, void* stream) {
  ZS3_CHECK_ARG(x && y, "spatial_broadcast_f32: bad args");
  spatial_broadcast_f32_kernel<<<pblocks((long long)N * HW * C, 256), 256, 0, ST(stream)>>>(x, y, N, HW, C);
  ZS3_CHECK_LAUNCH("spatial_broadcast_f32");
  return ZS3_OK;
}

extern "C" int zs3_nchw_to_nhwc_f32(const float* src, float* dst, int N, int C, long long HW, int cs, void* stream) {
  ZS3_CHECK_ARG(src && dst && cs >= C, "nchw_to_nhwc_f32: bad args");
  nchw_to_nhwc_f32_kernel<<<pblocks((long long)N * HW * cs, 256), 256, 0, ST(stream)>>>(src, dst, N, C, HW, cs);
  ZS3_CHECK_LAUNCH("nchw_to_nhwc_f32");
  return ZS3_OK;
}

extern "C" int zs3_stem_im2col_f32(const float* x, float* cols, int N, int C, int H, int W, int R, int stride, int pad,
                                   int Ho, int Wo, int kpad, int krsc, void* stream) {
  ZS3_CHECK_ARG(x && cols && kpad >= C * R * R, "stem_im2col_f32: bad args");
  stem_im2col_f32_kernel<<<pblocks((long long)N * Ho * Wo * kpad, 256), 256, 0, ST(stream)>>>(
      x, cols, N, C, H, W, R, stride, pad, Ho, Wo, kpad, krsc);
  ZS3_CHECK_LAUNCH("stem_im2col_f32");
  return ZS3_OK;
}
