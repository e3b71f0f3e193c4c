// BatchNorm2d training/eval passes fused with ReLU, residual add and Dropout (NHWC bf16, fp32 math).
// HBM-bound kernels: 16-byte vector accesses, one pass over each tensor, per-channel reductions kept in
// registers/shared memory and flushed with one fp64 atomic per channel per CTA.
//
// Reference semantics: F.batch_norm as called from zs3/modeling/sync_batchnorm/batchnorm.py:48-58
// (biased variance for normalisation, unbiased for running_var, momentum 0.1, eps 1e-5),
// nn.ReLU, `out += residual` (zs3/modeling/backbone/resnet.py:50), nn.Dropout (aspp.py:101, decoder.py:19,23).
#include <stdlib.h>

#include "common.cuh"
#include "ptx.cuh"

namespace zs3 {

__device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}

// keep decisions for 8 consecutive logical elements starting at linear index `idx` (multiple of 8)
__device__ __forceinline__ uint32_t dropout_keep8(uint64_t seed, uint64_t offset, uint64_t idx, uint32_t thresh16) {
  const uint64_t h0 = splitmix64(seed ^ splitmix64(offset + idx));
  const uint64_t h1 = splitmix64(h0 ^ 0xD1B54A32D192ED03ull);
  uint32_t keep = 0;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    keep |= (uint32_t)(((h0 >> (16 * j)) & 0xFFFF) >= thresh16) << j;
    keep |= (uint32_t)(((h1 >> (16 * j)) & 0xFFFF) >= thresh16) << (4 + j);
  }
  return keep;
}

__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
  f[0] = bf16_lo(u.x); f[1] = bf16_hi(u.x); f[2] = bf16_lo(u.y); f[3] = bf16_hi(u.y);
  f[4] = bf16_lo(u.z); f[5] = bf16_hi(u.z); f[6] = bf16_lo(u.w); f[7] = bf16_hi(u.w);
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  uint4 u;
  u.x = pack_bf16x2(f[0], f[1]); u.y = pack_bf16x2(f[2], f[3]);
  u.z = pack_bf16x2(f[4], f[5]); u.w = pack_bf16x2(f[6], f[7]);
  return u;
}
__device__ __forceinline__ void load8f(const float* p, float (&f)[8]) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(p));
  const float4 b = __ldg(reinterpret_cast<const float4*>(p) + 1);
  f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
}

__device__ __forceinline__ void lds8f(const float* p, float (&f)[8]) {  // shared-memory flavour (no __ldg)
  const float4 a = *reinterpret_cast<const float4*>(p);
  const float4 b = *(reinterpret_cast<const float4*>(p) + 1);
  f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
}

// ------------------------------------------------------------------------------- finalize
__global__ void bn_finalize_kernel(double* sum, double* sqsum, long long count, const float* gamma, const float* beta,
                                   float eps, float momentum, float* rmean, float* rvar, float* scale, float* shift,
                                   float* mean_out, float* invstd_out, int C, int Cpad, int reset) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= Cpad) return;
  if (c < C) {
    const double n = (double)count;
    const double m = sum[c] / n;
    double var = sqsum[c] / n - m * m;  // biased
    if (var < 0.0) var = 0.0;
    const float invstd = (float)(1.0 / sqrt(var + (double)eps));
    const float g = gamma ? gamma[c] : 1.f;
    const float b = beta ? beta[c] : 0.f;
    const float sc = g * invstd;
    scale[c] = sc;
    shift[c] = b - (float)m * sc;
    if (mean_out) mean_out[c] = (float)m;
    if (invstd_out) invstd_out[c] = invstd;
    if (rmean) {
      const double unbiased = count > 1 ? var * n / (n - 1.0) : var;
      rmean[c] = (1.f - momentum) * rmean[c] + momentum * (float)m;
      rvar[c] = (1.f - momentum) * rvar[c] + momentum * (float)unbiased;
    }
  } else {
    scale[c] = 0.f;
    shift[c] = 0.f;
    if (mean_out) mean_out[c] = 0.f;
    if (invstd_out) invstd_out[c] = 0.f;
  }
  if (reset) {
    sum[c] = 0.0;
    sqsum[c] = 0.0;
  }
}

__global__ void bn_eval_coeffs_kernel(const float* gamma, const float* beta, const float* rmean, const float* rvar,
                                      float eps, float* scale, float* shift, float* mean_out, float* invstd_out, int C,
                                      int Cpad) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= Cpad) return;
  float sc = 0.f, sh = 0.f, m = 0.f, is = 0.f;
  if (c < C) {
    m = rmean[c];
    is = 1.f / sqrtf(rvar[c] + eps);
    sc = (gamma ? gamma[c] : 1.f) * is;
    sh = (beta ? beta[c] : 0.f) - m * sc;
  }
  scale[c] = sc;
  shift[c] = sh;
  if (mean_out) mean_out[c] = m;
  if (invstd_out) invstd_out[c] = is;
}

// ---------------------------------------------------------------------------------- apply
// Streaming pattern shared by the kernels below: a thread owns ONE 8-channel slice (16 bytes of every pixel row) for
// the whole kernel, so the per-channel coefficients live in registers and there is no index arithmetic in the loop;
// rows are visited with a grid stride and UNROLL rows are in flight per thread (all loads issued before any use) to
// keep enough bytes outstanding for HBM.  A 256-thread CTA covers rows_per_block = 256 / (C/8) rows at a time.
constexpr int UNROLL = 4;

struct ApplyP {
  const __nv_bfloat16* y; long long y_cs;
  const __nv_bfloat16* res; long long res_cs;
  __nv_bfloat16* out; long long out_cs;
  const float* scale; const float* shift;
  long long M; int C; int relu; int drop_mode; uint32_t thresh16; float keep_scale;
  uint64_t seed, offset; const unsigned char* mask; const unsigned long long* offset_dev;
  int rows_per_block;
  // fused finalize (training mode): coefficients are derived from the raw batch statistics inside this kernel
  const double* stat_sum; const double* stat_sqsum; long long count; double inv_count; const float* gamma; const float* beta;
  float eps, momentum; float* rmean; float* rvar; int C_real;
  float* mean_out; float* invstd_out; float* scale_out; float* shift_out;
  double* reset_a; double* reset_b;  // the OTHER statistics buffer, zeroed for the next layer
  int reset_count;
  unsigned char* relu_bits;  // optional [M][C/8] ReLU bit mask for the backward
  int sync_clamp;            // invstd = rsqrt(max(var, eps)) (reference SyncBN) instead of rsqrt(var + eps)
};

__device__ __forceinline__ void apply_one(const ApplyP& p, long long m, int c, const float (&sc)[8],
                                          const float (&sh)[8], uint64_t rng_offset, const uint4& yraw,
                                          const uint4& rraw) {
  float f[8];
  unpack8(yraw, f);
#pragma unroll
  for (int j = 0; j < 8; ++j) f[j] = fmaf(f[j], sc[j], sh[j]);
  if (p.res) {
    float r[8];
    unpack8(rraw, r);
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] += r[j];
  }
  if (p.relu) {
    if (p.relu_bits) {
      unsigned bits = 0;
#pragma unroll
      for (int j = 0; j < 8; ++j) bits |= (f[j] > 0.f ? 1u : 0u) << j;
      p.relu_bits[m * (p.C >> 3) + (c >> 3)] = (unsigned char)bits;
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] = fmaxf(f[j], 0.f);
  }
  if (p.drop_mode == 1) {
    const uint32_t keep = dropout_keep8(p.seed, rng_offset, (uint64_t)(m * p.C + c), p.thresh16);
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] = ((keep >> j) & 1) ? f[j] * p.keep_scale : 0.f;
  } else if (p.drop_mode == 2) {
    const uint2 mk = *reinterpret_cast<const uint2*>(p.mask + m * p.C + c);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const uint32_t b = ((j < 4 ? mk.x : mk.y) >> (8 * (j & 3))) & 0xFF;
      f[j] = b ? f[j] * p.keep_scale : 0.f;
    }
  }
  *reinterpret_cast<uint4*>(p.out + m * p.out_cs + c) = pack8(f);
}

__global__ void __launch_bounds__(256) bn_apply_kernel(const ApplyP p) {
  griddep_sync();
  extern __shared__ float s_coef[];  // [2][C] scale, shift (fused-finalize launches only)
  const int vpc = p.C >> 3;
  const int c = (threadIdx.x % vpc) << 3;
  const int rl = threadIdx.x / vpc;
  const bool active = rl < p.rows_per_block;
  if (blockIdx.x == 0 && p.reset_a != nullptr) {
    for (int i = threadIdx.x; i < p.reset_count; i += blockDim.x) {
      p.reset_a[i] = 0.0;
      p.reset_b[i] = 0.0;
    }
  }
  const uint64_t rng_offset = p.offset + (p.offset_dev ? *p.offset_dev : 0ull);
  const long long stride = (long long)gridDim.x * p.rows_per_block;
  long long m = (long long)blockIdx.x * p.rows_per_block + rl;
  const uint4 zero = make_uint4(0, 0, 0, 0);
  // the first batch of rows is requested BEFORE the coefficients are derived, so that the statistics loads + math
  // and the first activation loads share one memory round trip (most launches of this kernel are 5-20 us long)
  const bool first = active && m + (UNROLL - 1) * stride < p.M;
  uint4 yr0[UNROLL], rr0[UNROLL];
  if (first) {
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      yr0[u] = *reinterpret_cast<const uint4*>(p.y + (m + u * stride) * p.y_cs + c);
      rr0[u] = p.res ? *reinterpret_cast<const uint4*>(p.res + (m + u * stride) * p.res_cs + c) : zero;
    }
  }
  float sc[8], sh[8];
  if (p.stat_sum != nullptr) {
    // fused bn_finalize, once per CTA: thread t derives the affine of channel t (t, t+256, ...) from the fp64 sums
    // into shared memory -- NOT every thread for its own 8 channels, which had all 300k threads of the grid hammer
    // the same 2*C doubles in L2 and cost 8-15 us per launch.  The cancellation E[x^2] - mean^2 stays in fp64, the
    // rest is fp32.  CTA 0 also publishes mean/invstd/scale/shift for the backward pass and updates the running
    // statistics.
    const bool publish = blockIdx.x == 0;
    for (int ch = threadIdx.x; ch < p.C; ch += blockDim.x) {
      float s = 0.f, b = 0.f, mu = 0.f, is = 0.f;
      if (ch < p.C_real) {
        const double mean = p.stat_sum[ch] * p.inv_count;
        const double dvar = fma(-mean, mean, p.stat_sqsum[ch] * p.inv_count);
        const float var = fmaxf((float)dvar, 0.f);
        is = p.sync_clamp ? rsqrtf(fmaxf(var, p.eps)) : rsqrtf(var + p.eps);
        mu = (float)mean;
        s = (p.gamma ? p.gamma[ch] : 1.f) * is;
        b = (p.beta ? p.beta[ch] : 0.f) - mu * s;
        if (publish && p.rmean) {
          const float unbiased = p.count > 1 ? var * ((float)p.count / (float)(p.count - 1)) : var;
          p.rmean[ch] = (1.f - p.momentum) * p.rmean[ch] + p.momentum * mu;
          p.rvar[ch] = (1.f - p.momentum) * p.rvar[ch] + p.momentum * unbiased;
        }
      }
      s_coef[ch] = s;
      s_coef[p.C + ch] = b;
      if (publish) {
        p.scale_out[ch] = s;
        p.shift_out[ch] = b;
        p.mean_out[ch] = mu;
        p.invstd_out[ch] = is;
      }
    }
    __syncthreads();
    lds8f(s_coef + c, sc);
    lds8f(s_coef + p.C + c, sh);
  } else {
    load8f(p.scale + c, sc);
    load8f(p.shift + c, sh);
  }
  if (!active) return;
  if (first) {
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) apply_one(p, m + u * stride, c, sc, sh, rng_offset, yr0[u], rr0[u]);
    m += UNROLL * stride;
  }
  for (; m + (UNROLL - 1) * stride < p.M; m += UNROLL * stride) {
    uint4 yr[UNROLL], rr[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      yr[u] = *reinterpret_cast<const uint4*>(p.y + (m + u * stride) * p.y_cs + c);
      rr[u] = p.res ? *reinterpret_cast<const uint4*>(p.res + (m + u * stride) * p.res_cs + c) : zero;
    }
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) apply_one(p, m + u * stride, c, sc, sh, rng_offset, yr[u], rr[u]);
  }
  for (; m < p.M; m += stride) {
    const uint4 yr = *reinterpret_cast<const uint4*>(p.y + m * p.y_cs + c);
    const uint4 rr = p.res ? *reinterpret_cast<const uint4*>(p.res + m * p.res_cs + c) : zero;
    apply_one(p, m, c, sc, sh, rng_offset, yr, rr);
  }
}

// ------------------------------------------------------------------------------- backward
struct BwdP {
  const __nv_bfloat16* dout; long long dout_cs;
  const __nv_bfloat16* out; long long out_cs;
  const __nv_bfloat16* y; long long y_cs;
  const float* mean; const float* invstd; const float* scale; const float* shift;
  long long M; int C; int relu; float grad_scale; int training;
  double* sum_dz; double* sum_dzx;
  __nv_bfloat16* dy; long long dy_cs;
  int sp_stride, sp_HoWo, sp_Wo; long long dy_img, dy_row;
  __nv_bfloat16* dres; long long dres_cs; int dres_acc;
  float* dgamma; float* dbeta; int C_real; int param_acc;
  int rows_per_block;  // pixel rows handled concurrently by one CTA
  double* reset_a; double* reset_b; int reset_count;  // the other sums buffer, zeroed by the apply phase
  const unsigned char* relu_bits;  // relu == 3
  long long stat_count;            // divisor of the batch sums (0: M); Synchronised BatchNorm passes the global count
};

// the "forward activation" operand of make_dz: the saved output row (relu == 1), the ReLU mask byte (relu == 3) or nothing
__device__ __forceinline__ uint4 load_mask_or_out(const BwdP& p, long long m, int c, const uint4& zero) {
  if (p.relu == 1) return *reinterpret_cast<const uint4*>(p.out + m * p.out_cs + c);
  if (p.relu == 3) return make_uint4(p.relu_bits[m * (p.C >> 3) + (c >> 3)], 0u, 0u, 0u);
  return zero;
}

// dz = dout * [forward activation > 0] * grad_scale for one 8-channel slice.
// relu == 1: mask from the saved forward output (residual / dropout layers); relu == 2: mask recomputed as
// scale*y + shift > 0 (plain conv->BN->ReLU layers: saves reading `out` in both backward passes).
__device__ __forceinline__ void make_dz(const BwdP& p, const uint4& draw, const uint4& oraw, const float (&yv)[8],
                                        const float (&sc)[8], const float (&sh)[8], float (&dz)[8]) {
  unpack8(draw, dz);
  if (p.relu == 2) {
#pragma unroll
    for (int j = 0; j < 8; ++j) dz[j] = fmaf(yv[j], sc[j], sh[j]) > 0.f ? dz[j] * p.grad_scale : 0.f;
  } else if (p.relu == 1) {
    float o[8];
    unpack8(oraw, o);
#pragma unroll
    for (int j = 0; j < 8; ++j) dz[j] = o[j] > 0.f ? dz[j] * p.grad_scale : 0.f;
  } else if (p.relu == 3) {  // oraw.x carries the mask byte of this 8-channel slice
#pragma unroll
    for (int j = 0; j < 8; ++j) dz[j] = ((oraw.x >> j) & 1u) ? dz[j] * p.grad_scale : 0.f;
  } else if (p.grad_scale != 1.f) {
#pragma unroll
    for (int j = 0; j < 8; ++j) dz[j] *= p.grad_scale;
  }
}

__global__ void __launch_bounds__(256) bn_bwd_reduce_kernel(const BwdP p) {
  griddep_sync();
  extern __shared__ float red[];  // [rows_per_block][C] x 2
  const int vpc = p.C >> 3;
  const int c = (threadIdx.x % vpc) << 3;
  const int rl = threadIdx.x / vpc;
  float a1[8], a2[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) a1[j] = a2[j] = 0.f;
  if (rl < p.rows_per_block) {
    float mu[8], is[8], sc[8], sh[8];
    load8f(p.mean + c, mu);
    load8f(p.invstd + c, is);
    load8f(p.scale + c, sc);
    if (p.relu == 2) load8f(p.shift + c, sh);
    const long long stride = (long long)gridDim.x * p.rows_per_block;
    long long m = (long long)blockIdx.x * p.rows_per_block + rl;
    const uint4 zero = make_uint4(0, 0, 0, 0);
    auto accumulate = [&](const uint4& yr, const uint4& dr, const uint4& orw) {
      float yv[8], dz[8];
      unpack8(yr, yv);
      make_dz(p, dr, orw, yv, sc, sh, dz);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        a1[j] += dz[j];
        a2[j] = fmaf(dz[j], (yv[j] - mu[j]) * is[j], a2[j]);
      }
    };
    for (; m + (UNROLL - 1) * stride < p.M; m += UNROLL * stride) {
      uint4 yr[UNROLL], dr[UNROLL], orw[UNROLL];
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) {
        yr[u] = *reinterpret_cast<const uint4*>(p.y + (m + u * stride) * p.y_cs + c);
        dr[u] = *reinterpret_cast<const uint4*>(p.dout + (m + u * stride) * p.dout_cs + c);
        orw[u] = load_mask_or_out(p, m + u * stride, c, zero);
      }
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) accumulate(yr[u], dr[u], orw[u]);
    }
    for (; m < p.M; m += stride) {
      const uint4 yr = *reinterpret_cast<const uint4*>(p.y + m * p.y_cs + c);
      const uint4 dr = *reinterpret_cast<const uint4*>(p.dout + m * p.dout_cs + c);
      const uint4 orw = load_mask_or_out(p, m, c, zero);
      accumulate(yr, dr, orw);
    }
    float* r1 = red + (size_t)rl * p.C + c;
    float* r2 = red + (size_t)p.rows_per_block * p.C + (size_t)rl * p.C + c;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      r1[j] = a1[j];
      r2[j] = a2[j];
    }
  }
  __syncthreads();
  for (int ch = threadIdx.x; ch < p.C; ch += blockDim.x) {
    float s1 = 0.f, s2 = 0.f;
    for (int r = 0; r < p.rows_per_block; ++r) {
      s1 += red[(size_t)r * p.C + ch];
      s2 += red[(size_t)p.rows_per_block * p.C + (size_t)r * p.C + ch];
    }
    atomicAdd(p.sum_dz + ch, (double)s1);
    atomicAdd(p.sum_dzx + ch, (double)s2);
  }
}

// dy = A*dz + B*y + C per channel, with A = scale, B = -scale*m2*invstd, C = -scale*m1 + scale*m2*invstd*mean
// (m1 = sum_dz/M, m2 = sum_dzx/M): the fp64 sums are folded into three fp32 coefficients per channel once per
// thread, so the streaming loop is 2 FMAs per element and touches no fp64.
__device__ __forceinline__ void bwd_apply_one(const BwdP& p, long long m, int c, const float (&cA)[8],
                                              const float (&cB)[8], const float (&cC)[8], const float (&sc)[8],
                                              const float (&sh)[8], const uint4& yr, const uint4& dr, const uint4& orw) {
  float yv[8], dz[8], g[8];
  unpack8(yr, yv);
  make_dz(p, dr, orw, yv, sc, sh, dz);
  if (p.dres) {
    uint4* dst = reinterpret_cast<uint4*>(p.dres + m * p.dres_cs + c);
    if (p.dres_acc) {
      float r[8];
      unpack8(*dst, r);
#pragma unroll
      for (int j = 0; j < 8; ++j) r[j] += dz[j];
      *dst = pack8(r);
    } else {
      *dst = pack8(dz);
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) g[j] = fmaf(cA[j], dz[j], fmaf(cB[j], yv[j], cC[j]));
  long long pix = m;
  if (p.sp_stride > 1) {
    const long long img = m / p.sp_HoWo;
    const int rem = (int)(m - img * p.sp_HoWo);
    const int op = rem / p.sp_Wo;
    const int oq = rem - op * p.sp_Wo;
    pix = img * p.dy_img + (long long)op * p.dy_row + (long long)oq * p.sp_stride;
  }
  *reinterpret_cast<uint4*>(p.dy + pix * p.dy_cs + c) = pack8(g);
}

__global__ void __launch_bounds__(256) bn_bwd_apply_kernel(const BwdP p) {
  griddep_sync();
  extern __shared__ float s_coef[];  // [3][C]: A, B, C of dy = A*dz + B*y + C
  const int vpc = p.C >> 3;
  const int c = (threadIdx.x % vpc) << 3;
  const int rl = threadIdx.x / vpc;
  const bool active = rl < p.rows_per_block;
  if (blockIdx.x == 0 && p.reset_a != nullptr) {
    for (int i = threadIdx.x; i < p.reset_count; i += blockDim.x) {
      p.reset_a[i] = 0.0;
      p.reset_b[i] = 0.0;
    }
  }
  const long long stride = (long long)gridDim.x * p.rows_per_block;
  long long m = (long long)blockIdx.x * p.rows_per_block + rl;
  const uint4 zero = make_uint4(0, 0, 0, 0);
  // first batch requested before the coefficient loads/math (see bn_apply_kernel)
  const bool first = active && m + (UNROLL - 1) * stride < p.M;
  uint4 yr0[UNROLL], dr0[UNROLL], or0[UNROLL];
  if (first) {
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      yr0[u] = *reinterpret_cast<const uint4*>(p.y + (m + u * stride) * p.y_cs + c);
      dr0[u] = *reinterpret_cast<const uint4*>(p.dout + (m + u * stride) * p.dout_cs + c);
      or0[u] = load_mask_or_out(p, m + u * stride, c, zero);
    }
  }
  // per-channel coefficients once per CTA (thread t -> channel t, t+256, ...), through shared memory; CTA 0 also
  // writes the affine-parameter gradients
  {
    const float inv_m = 1.f / (float)(p.stat_count > 0 ? p.stat_count : p.M);
    for (int ch = threadIdx.x; ch < p.C; ch += blockDim.x) {
      const float scl = p.scale[ch];
      float B = 0.f, Cc = 0.f;
      const bool params = blockIdx.x == 0 && p.dgamma != nullptr && ch < p.C_real;
      if (p.training || params) {
        const float sdz = (float)p.sum_dz[ch], sdzx = (float)p.sum_dzx[ch];
        if (p.training) {
          const float m1 = sdz * inv_m, m2 = sdzx * inv_m, is = p.invstd[ch];
          B = -scl * m2 * is;
          Cc = -scl * m1 + scl * m2 * is * p.mean[ch];
        }
        if (params) {
          p.dgamma[ch] = p.param_acc ? p.dgamma[ch] + sdzx : sdzx;
          p.dbeta[ch] = p.param_acc ? p.dbeta[ch] + sdz : sdz;
        }
      }
      s_coef[ch] = scl;
      s_coef[p.C + ch] = B;
      s_coef[2 * p.C + ch] = Cc;
    }
  }
  __syncthreads();
  if (!active) return;
  float cA[8], cB[8], cC[8], sc[8], sh[8];
  lds8f(s_coef + c, cA);
  lds8f(s_coef + p.C + c, cB);
  lds8f(s_coef + 2 * p.C + c, cC);
#pragma unroll
  for (int j = 0; j < 8; ++j) sc[j] = cA[j];
  if (p.relu == 2) load8f(p.shift + c, sh);
  if (first) {
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) bwd_apply_one(p, m + u * stride, c, cA, cB, cC, sc, sh, yr0[u], dr0[u], or0[u]);
    m += UNROLL * stride;
  }
  for (; m + (UNROLL - 1) * stride < p.M; m += UNROLL * stride) {
    uint4 yr[UNROLL], dr[UNROLL], orw[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      yr[u] = *reinterpret_cast<const uint4*>(p.y + (m + u * stride) * p.y_cs + c);
      dr[u] = *reinterpret_cast<const uint4*>(p.dout + (m + u * stride) * p.dout_cs + c);
      orw[u] = load_mask_or_out(p, m + u * stride, c, zero);
    }
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) bwd_apply_one(p, m + u * stride, c, cA, cB, cC, sc, sh, yr[u], dr[u], orw[u]);
  }
  for (; m < p.M; m += stride) {
    const uint4 yr = *reinterpret_cast<const uint4*>(p.y + m * p.y_cs + c);
    const uint4 dr = *reinterpret_cast<const uint4*>(p.dout + m * p.dout_cs + c);
    const uint4 orw = load_mask_or_out(p, m, c, zero);
    bwd_apply_one(p, m, c, cA, cB, cC, sc, sh, yr, dr, orw);
  }
}

// ----------------------------------------------------------------------- standalone statistics
// sum[c] += sum_m y[m][c], sqsum[c] += sum_m y[m][c]^2 (used instead of the conv epilogue's fused statistics for
// short-K convolutions whose epilogue would otherwise be the bottleneck; costs one extra read of y)
__global__ void __launch_bounds__(256) bn_stats_kernel(const __nv_bfloat16* __restrict__ y, long long y_cs, long long M,
                                                       int C, int rows_per_block, double* __restrict__ sum,
                                                       double* __restrict__ sqsum) {
  griddep_sync();
  extern __shared__ float red[];  // [rows_per_block][C] x 2
  const int vpc = C >> 3;
  const int c = (threadIdx.x % vpc) << 3;
  const int rl = threadIdx.x / vpc;
  float a1[8], a2[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) a1[j] = a2[j] = 0.f;
  if (rl < rows_per_block) {
    const long long stride = (long long)gridDim.x * rows_per_block;
    long long m = (long long)blockIdx.x * rows_per_block + rl;
    auto acc = [&](const uint4& raw) {
      float v[8];
      unpack8(raw, v);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        a1[j] += v[j];
        a2[j] = fmaf(v[j], v[j], a2[j]);
      }
    };
    for (; m + (2 * UNROLL - 1) * stride < M; m += 2 * UNROLL * stride) {
      uint4 r[2 * UNROLL];
#pragma unroll
      for (int u = 0; u < 2 * UNROLL; ++u) r[u] = *reinterpret_cast<const uint4*>(y + (m + u * stride) * y_cs + c);
#pragma unroll
      for (int u = 0; u < 2 * UNROLL; ++u) acc(r[u]);
    }
    for (; m < M; m += stride) acc(*reinterpret_cast<const uint4*>(y + m * y_cs + c));
    float* r1 = red + (size_t)rl * C + c;
    float* r2 = red + (size_t)rows_per_block * C + (size_t)rl * C + c;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      r1[j] = a1[j];
      r2[j] = a2[j];
    }
  }
  __syncthreads();
  for (int ch = threadIdx.x; ch < C; ch += blockDim.x) {
    float s1 = 0.f, s2 = 0.f;
    for (int r = 0; r < rows_per_block; ++r) {
      s1 += red[(size_t)r * C + ch];
      s2 += red[(size_t)rows_per_block * C + (size_t)r * C + ch];
    }
    atomicAdd(sum + ch, (double)s1);
    atomicAdd(sqsum + ch, (double)s2);
  }
}

// ------------------------------------------------------------------------- layout changes
// NCHW fp32 -> NHWC bf16 through a 32x32 shared-memory transpose (coalesced on both sides).
__global__ void nchw_to_nhwc_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, int C, long long HW,
                                    int cs) {
  __shared__ float tile[32][33];
  const int n = blockIdx.z;
  const long long hw0 = (long long)blockIdx.x * 32;
  const int c0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i;
    const long long hw = hw0 + threadIdx.x;
    tile[i][threadIdx.x] = (c < C && hw < HW) ? src[((long long)n * C + c) * HW + hw] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const long long hw = hw0 + i;
    const int c = c0 + threadIdx.x;
    if (hw < HW && c < cs) dst[((long long)n * HW + hw) * cs + c] = __float2bfloat16(tile[threadIdx.x][i]);
  }
}

__global__ void nhwc_to_nchw_kernel(const __nv_bfloat16* __restrict__ src, float* __restrict__ dst, int C, long long HW,
                                    int cs) {
  __shared__ float tile[32][33];
  const int n = blockIdx.z;
  const long long hw0 = (long long)blockIdx.x * 32;
  const int c0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const long long hw = hw0 + i;
    const int c = c0 + threadIdx.x;
    tile[i][threadIdx.x] = (hw < HW && c < C) ? __bfloat162float(src[((long long)n * HW + hw) * cs + c]) : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i;
    const long long hw = hw0 + threadIdx.x;
    if (c < C && hw < HW) dst[((long long)n * C + c) * HW + hw] = tile[threadIdx.x][i];
  }
}

// grids for the row-streaming kernels: every thread should see at least UNROLL rows.  The map kernels (apply,
// bwd_apply) run up to 8 CTAs per SM.  The reducing kernels (stats, bwd_reduce) end with one fp64 atomic per channel
// and CTA onto the SAME 2*C addresses -- measured ~19 ns per same-address atomic, i.e. 20 us of tail at 1184 CTAs --
// so they use fewer, longer-running CTAs.  ZS3_BN_MAP_CTAS / ZS3_BN_REDUCE_CTAS (per SM) override for experiments.
static int env_int(const char* name, int dflt) {
  const char* e = getenv(name);
  const int v = e ? atoi(e) : dflt;
  return v > 0 ? v : dflt;  // non-positive / unset -> default
}

static int stream_grid_cap(long long M, int rows_per_block, int ctas_per_sm, int rows_per_thread) {
  long long b = (M + (long long)rows_per_block * rows_per_thread - 1) / ((long long)rows_per_block * rows_per_thread);
  const long long cap = 148ll * ctas_per_sm;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

// map kernels: measured on the full training step (profiles/r01_bn_grid_experiment.md) 2 CTAs/SM beat 1, 3, 4 and 8
// (854 -> 874 img/s from 8 to 2): fewer CTAs mean fewer per-CTA coefficient prologues and a shorter tail, and
// 2 x 256 threads x 4 rows in flight already cover the L2/HBM latency; only the > 96 MB tensors (stem, layer1's
// 256-channel outputs) keep 4 CTAs/SM, which streams them 10 % faster (tools/bn_bench.py)
static int stream_grid(long long M, int rows_per_block, long long bytes = 0) {
  static const int forced = env_int("ZS3_BN_MAP_CTAS", -1);
  int per_sm = bytes >= (96ll << 20) ? 4 : 2;
  if (forced > 0) per_sm = forced;
  return stream_grid_cap(M, rows_per_block, per_sm, UNROLL);
}

// measured (tools/bn_bench.py, profiles/r01_bn_bench.md): 2 CTAs/SM is best for the backward reduction at every
// size; the lighter statistics kernel prefers 4 CTAs/SM on tensors that do not fit L2 and 1 on the smallest ones
static int reduce_grid(long long M, int rows_per_block, long long bytes = 0, bool stats = false) {
  static const int forced = env_int("ZS3_BN_REDUCE_CTAS", -1);
  int per_sm = 2;
  if (stats) per_sm = bytes >= (48ll << 20) ? 4 : (bytes <= (12ll << 20) ? 1 : 2);
  if (forced > 0) per_sm = forced;
  return stream_grid_cap(M, rows_per_block, per_sm, 2 * UNROLL);
}

static int ew_grid(long long work_items, int threads) {
  long long b = (work_items + threads - 1) / threads;
  const long long cap = 148ll * 16;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

}  // namespace zs3

using namespace zs3;

extern "C" int zs3_bn_finalize(double* stat_sum, double* stat_sqsum, long long count, const float* gamma,
                               const float* beta, float eps, float momentum, float* running_mean, float* running_var,
                               float* scale, float* shift, float* mean, float* invstd, int C, int Cpad,
                               int reset_stats, void* stream) {
  ZS3_CHECK_ARG(stat_sum && stat_sqsum && scale && shift, "bn_finalize: null pointer");
  ZS3_CHECK_ARG(count > 0 && C > 0 && Cpad >= C, "bn_finalize: bad sizes count=%lld C=%d Cpad=%d", count, C, Cpad);
  ZS3_CHECK_ARG((running_mean == nullptr) == (running_var == nullptr), "bn_finalize: running stats mismatch");
  bn_finalize_kernel<<<ceil_div(Cpad, 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(
      stat_sum, stat_sqsum, count, gamma, beta, eps, momentum, running_mean, running_var, scale, shift, mean, invstd,
      C, Cpad, reset_stats);
  ZS3_CHECK_LAUNCH("bn_finalize");
  return ZS3_OK;
}

extern "C" int zs3_bn_stats(const void* y, int y_cstride, long long M, int C, double* stat_sum, double* stat_sqsum,
                            void* stream) {
  ZS3_CHECK_ARG(y && stat_sum && stat_sqsum, "bn_stats: null pointer");
  ZS3_CHECK_ARG(C > 0 && C % 8 == 0 && C <= 2048 && y_cstride % 8 == 0 && y_cstride >= C,
                "bn_stats: C=%d must be a multiple of 8 and <= 2048", C);
  if (M <= 0) return ZS3_OK;
  const int vpc = C / 8;
  int rpb = 256 / vpc;
  if (rpb < 1) rpb = 1;
  const size_t smem = (size_t)2 * rpb * C * sizeof(float);
  launch_pdl(PDL_BN, bn_stats_kernel, dim3(reduce_grid(M, rpb, M * C * 2, true)), dim3(256), smem, static_cast<cudaStream_t>(stream),
             static_cast<const __nv_bfloat16*>(y), (long long)y_cstride, M, C, rpb, stat_sum, stat_sqsum);
  ZS3_CHECK_LAUNCH("bn_stats");
  return ZS3_OK;
}

extern "C" int zs3_bn_eval_coeffs(const float* gamma, const float* beta, const float* running_mean,
                                  const float* running_var, float eps, float* scale, float* shift, float* mean,
                                  float* invstd, int C, int Cpad, void* stream) {
  ZS3_CHECK_ARG(running_mean && running_var && scale && shift, "bn_eval_coeffs: null pointer");
  ZS3_CHECK_ARG(C > 0 && Cpad >= C, "bn_eval_coeffs: bad sizes");
  bn_eval_coeffs_kernel<<<ceil_div(Cpad, 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(
      gamma, beta, running_mean, running_var, eps, scale, shift, mean, invstd, C, Cpad);
  ZS3_CHECK_LAUNCH("bn_eval_coeffs");
  return ZS3_OK;
}

extern "C" int zs3_bn_apply(const zs3_bn_apply_args* a, void* stream) {
  ZS3_CHECK_ARG(a && a->y && a->out && ((a->scale && a->shift) || a->stat_sum), "bn_apply: null pointer");
  ZS3_CHECK_ARG(a->stat_sum == nullptr ||
                    (a->stat_sqsum && a->count > 0 && a->mean_out && a->invstd_out && a->scale_out && a->shift_out &&
                     a->C_real > 0 && a->C_real <= a->C && (a->running_mean == nullptr) == (a->running_var == nullptr) &&
                     (a->reset_sum == nullptr) == (a->reset_sqsum == nullptr)),
                "bn_apply: incomplete fused-finalize arguments");
  ZS3_CHECK_ARG(a->C > 0 && a->C % 8 == 0 && a->y_cstride % 8 == 0 && a->out_cstride % 8 == 0 &&
                    a->y_cstride >= a->C && a->out_cstride >= a->C,
                "bn_apply: C=%d strides %d/%d must be multiples of 8", a->C, a->y_cstride, a->out_cstride);
  ZS3_CHECK_ARG(a->residual == nullptr || (a->res_cstride % 8 == 0 && a->res_cstride >= a->C),
                "bn_apply: residual stride");
  ZS3_CHECK_ARG(a->drop_mode >= 0 && a->drop_mode <= 2 && a->drop_p >= 0.f && a->drop_p < 1.f, "bn_apply: dropout");
  ZS3_CHECK_ARG(a->drop_mode != 2 || a->keep_mask != nullptr, "bn_apply: drop_mode 2 needs keep_mask");
  if (a->M <= 0) return ZS3_OK;
  ApplyP p;
  p.y = static_cast<const __nv_bfloat16*>(a->y); p.y_cs = a->y_cstride;
  p.res = static_cast<const __nv_bfloat16*>(a->residual); p.res_cs = a->res_cstride;
  p.out = static_cast<__nv_bfloat16*>(a->out); p.out_cs = a->out_cstride;
  p.scale = a->scale; p.shift = a->shift;
  p.M = a->M; p.C = a->C; p.relu = a->relu;
  p.drop_mode = (a->drop_p > 0.f) ? a->drop_mode : 0;
  p.thresh16 = (uint32_t)(a->drop_p * 65536.0f + 0.5f);
  p.keep_scale = 1.f / (1.f - a->drop_p);
  p.seed = a->seed; p.offset = a->offset; p.mask = a->keep_mask; p.offset_dev = a->offset_dev;
  p.stat_sum = a->stat_sum; p.stat_sqsum = a->stat_sqsum; p.count = a->count;
  p.inv_count = a->count > 0 ? 1.0 / (double)a->count : 0.0; p.gamma = a->gamma; p.beta = a->beta;
  p.eps = a->eps; p.momentum = a->momentum; p.rmean = a->running_mean; p.rvar = a->running_var; p.C_real = a->C_real;
  p.mean_out = a->mean_out; p.invstd_out = a->invstd_out; p.scale_out = a->scale_out; p.shift_out = a->shift_out;
  p.reset_a = a->reset_sum; p.reset_b = a->reset_sqsum; p.reset_count = a->reset_count;
  p.relu_bits = a->relu ? a->relu_mask_out : nullptr;
  p.sync_clamp = a->sync_clamp;
  ZS3_CHECK_ARG(a->C <= 2048, "bn_apply: C=%d > 2048", a->C);
  p.rows_per_block = 256 / (a->C / 8) > 0 ? 256 / (a->C / 8) : 1;
  const size_t smem = a->stat_sum ? (size_t)2 * a->C * sizeof(float) : 0;
  launch_pdl(PDL_BN, bn_apply_kernel, dim3(stream_grid(a->M, p.rows_per_block, a->M * a->C * 2)), dim3(256), smem, static_cast<cudaStream_t>(stream),
             p);
  ZS3_CHECK_LAUNCH("bn_apply");
  return ZS3_OK;
}

static int fill_bwd(const zs3_bn_bwd_args* a, BwdP& p, const char* who) {
  ZS3_CHECK_ARG(a && a->dout && a->y && a->mean && a->invstd && a->scale && a->sum_dz && a->sum_dzx,
                "%s: null pointer", who);
  ZS3_CHECK_ARG(a->C > 0 && a->C % 8 == 0 && a->C <= 2048 && a->dout_cstride % 8 == 0 && a->y_cstride % 8 == 0,
                "%s: C=%d must be a multiple of 8 and <= 2048", who, a->C);
  ZS3_CHECK_ARG(a->relu != 1 || (a->out && a->out_cstride % 8 == 0), "%s: relu=1 needs the forward output", who);
  ZS3_CHECK_ARG(a->relu != 2 || a->shift != nullptr, "%s: relu=2 (mask recomputed from y) needs shift", who);
  ZS3_CHECK_ARG(a->relu != 3 || a->relu_mask != nullptr, "%s: relu=3 needs the ReLU bit mask", who);
  ZS3_CHECK_ARG(a->relu >= 0 && a->relu <= 3, "%s: relu mode %d", who, a->relu);
  p.dout = static_cast<const __nv_bfloat16*>(a->dout); p.dout_cs = a->dout_cstride;
  p.out = static_cast<const __nv_bfloat16*>(a->out); p.out_cs = a->out_cstride;
  p.y = static_cast<const __nv_bfloat16*>(a->y); p.y_cs = a->y_cstride;
  p.mean = a->mean; p.invstd = a->invstd; p.scale = a->scale; p.shift = a->shift;
  p.M = a->M; p.C = a->C; p.relu = a->relu; p.grad_scale = a->grad_scale; p.training = a->training;
  p.sum_dz = a->sum_dz; p.sum_dzx = a->sum_dzx;
  p.dy = static_cast<__nv_bfloat16*>(a->dy); p.dy_cs = a->dy_cstride;
  p.sp_stride = a->dy_sp_stride; p.sp_HoWo = a->sp_Ho * a->sp_Wo; p.sp_Wo = a->sp_Wo;
  p.dy_img = (long long)a->dy_H * a->dy_W; p.dy_row = (long long)a->dy_W * a->dy_sp_stride;
  p.dres = static_cast<__nv_bfloat16*>(a->dres); p.dres_cs = a->dres_cstride; p.dres_acc = a->dres_accumulate;
  p.dgamma = a->dgamma; p.dbeta = a->dbeta; p.C_real = a->C_real; p.param_acc = a->param_accumulate;
  p.reset_a = a->reset_sum_dz; p.reset_b = a->reset_sum_dzx; p.reset_count = a->reset_count;
  p.relu_bits = a->relu_mask;
  p.stat_count = a->stat_count;
  p.rows_per_block = 1;
  return ZS3_OK;
}

extern "C" int zs3_bn_bwd_reduce(const zs3_bn_bwd_args* a, void* stream) {
  BwdP p;
  int rc = fill_bwd(a, p, "bn_bwd_reduce");
  if (rc) return rc;
  if (a->M <= 0) return ZS3_OK;
  const int vpc = a->C / 8;
  p.rows_per_block = 256 / vpc;  // C <= 2048 -> vpc <= 256
  if (p.rows_per_block < 1) p.rows_per_block = 1;
  const size_t smem = (size_t)2 * p.rows_per_block * a->C * sizeof(float);
  launch_pdl(PDL_BN, bn_bwd_reduce_kernel, dim3(reduce_grid(a->M, p.rows_per_block)), dim3(256), smem,
             static_cast<cudaStream_t>(stream), p);
  ZS3_CHECK_LAUNCH("bn_bwd_reduce");
  return ZS3_OK;
}

extern "C" int zs3_bn_bwd_apply(const zs3_bn_bwd_args* a, void* stream) {
  BwdP p;
  int rc = fill_bwd(a, p, "bn_bwd_apply");
  if (rc) return rc;
  ZS3_CHECK_ARG(a->dy && a->dy_cstride % 8 == 0 && a->dy_cstride >= a->C, "bn_bwd_apply: bad dy");
  ZS3_CHECK_ARG(a->dres == nullptr || (a->dres_cstride % 8 == 0 && a->dres_cstride >= a->C), "bn_bwd_apply: bad dres");
  ZS3_CHECK_ARG(a->dy_sp_stride <= 1 || (a->sp_Ho > 0 && a->sp_Wo > 0 && a->dy_H >= (a->sp_Ho - 1) * a->dy_sp_stride + 1 &&
                                         a->dy_W >= (a->sp_Wo - 1) * a->dy_sp_stride + 1),
                "bn_bwd_apply: bad scatter geometry");
  ZS3_CHECK_ARG((a->dgamma == nullptr) == (a->dbeta == nullptr) && (a->dgamma == nullptr || a->C_real <= a->C),
                "bn_bwd_apply: bad parameter gradient buffers");
  if (a->M <= 0) return ZS3_OK;
  p.rows_per_block = 256 / (a->C / 8) > 0 ? 256 / (a->C / 8) : 1;
  launch_pdl(PDL_BN, bn_bwd_apply_kernel, dim3(stream_grid(a->M, p.rows_per_block, a->M * a->C * 2)), dim3(256), (size_t)3 * a->C * sizeof(float),
             static_cast<cudaStream_t>(stream), p);
  ZS3_CHECK_LAUNCH("bn_bwd_apply");
  return ZS3_OK;
}

extern "C" int zs3_nchw_f32_to_nhwc_bf16(const float* src, void* dst, int N, int C, long long HW, int cs,
                                         void* stream) {
  ZS3_CHECK_ARG(src && dst && N > 0 && C > 0 && HW > 0 && cs >= C, "nchw_to_nhwc: bad args");
  dim3 grid((unsigned)((HW + 31) / 32), (unsigned)((cs + 31) / 32), (unsigned)N);
  nchw_to_nhwc_kernel<<<grid, dim3(32, 8), 0, static_cast<cudaStream_t>(stream)>>>(
      src, static_cast<__nv_bfloat16*>(dst), C, HW, cs);
  ZS3_CHECK_LAUNCH("nchw_to_nhwc");
  return ZS3_OK;
}

extern "C" int zs3_nhwc_bf16_to_nchw_f32(const void* src, float* dst, int N, int C, long long HW, int cs,
                                         void* stream) {
  ZS3_CHECK_ARG(src && dst && N > 0 && C > 0 && HW > 0 && cs >= C, "nhwc_to_nchw: bad args");
  dim3 grid((unsigned)((HW + 31) / 32), (unsigned)((C + 31) / 32), (unsigned)N);
  nhwc_to_nchw_kernel<<<grid, dim3(32, 8), 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(src), dst, C, HW, cs);
  ZS3_CHECK_LAUNCH("nhwc_to_nchw");
  return ZS3_OK;
}
