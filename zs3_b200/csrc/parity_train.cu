// fp32-activation ("split precision") TRAINING support kernels.
//
// The throughput path stores activations in bf16, which bounds end-to-end agreement with the fp32 reference at
// ~1e-2 (DESIGN.md "Numerics").  The split-precision mode keeps every activation and gradient tensor in fp32 and
// feeds the tcgen05 conv kernel with P bf16 pieces per operand (x = p0 + p1 [+ p2], 8 mantissa bits each); the
// P(P+1)/2 significant cross products are reduced into one fp32 TMEM accumulator (zs3_conv_fprop K-segments, or
// repeated zs3_conv_wgrad launches that reduce-add into the same gradient).  P = 1 is plain bf16 operands with fp32
// storage, P = 2 carries 16 mantissa bits per operand (better than cuDNN's TF32, 10 bits), P = 3 all 24.
//
// This file holds the HBM-bound glue of that mode, forward AND backward, all fp32 I/O, NHWC with a channel stride
// that is a multiple of 64 (padding channels are zero):
//   * BatchNorm apply (+residual, ReLU, Dropout) fused with the operand split of its output,
//   * BatchNorm backward (two passes: per-channel sums, then dy written directly as bf16 pieces),
//   * max-pool with argmax + backward, bilinear backward (NHWC and from NCHW logits), pooled-branch helpers.
// Reference semantics: F.batch_norm (zs3/modeling/sync_batchnorm/batchnorm.py:48-58), nn.ReLU, `out += residual`
// (zs3/modeling/backbone/resnet.py:50), nn.Dropout (aspp.py:101, decoder.py:19,23), nn.MaxPool2d (resnet.py:82),
// F.interpolate(bilinear, align_corners=True) (decoder.py:33-35, deeplab.py:44).
#include "common.cuh"
#include "ptx.cuh"

namespace zs3 {

static int ptblocks(long long items, int threads, int cap = 148 * 16) {
  long long b = (items + threads - 1) / threads;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

__device__ __forceinline__ uint64_t pt_splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}

// same stream as bn.cu's dropout_keep8: keep decisions of 8 consecutive elements starting at idx (multiple of 8)
__device__ __forceinline__ uint32_t pt_dropout_keep8(uint64_t seed, uint64_t offset, uint64_t idx, uint32_t thresh16) {
  const uint64_t h0 = pt_splitmix64(seed ^ pt_splitmix64(offset + idx));
  const uint64_t h1 = pt_splitmix64(h0 ^ 0xD1B54A32D192ED03ull);
  uint32_t keep = 0;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    keep |= (uint32_t)(((h0 >> (16 * j)) & 0xFFFF) >= thresh16) << j;
    keep |= (uint32_t)(((h1 >> (16 * j)) & 0xFFFF) >= thresh16) << (4 + j);
  }
  return keep;
}

// v -> up to three bf16 pieces, written as 8 consecutive channels (16-byte stores)
struct Pieces {
  __nv_bfloat16* p[3];
  int n;
};

__device__ __forceinline__ void store_pieces8(const Pieces& pc, long long off, const float (&f)[8]) {
  float r[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) r[j] = f[j];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    if (k >= pc.n) break;
    uint32_t w[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const __nv_bfloat16 a = __float2bfloat16(r[2 * j]), b = __float2bfloat16(r[2 * j + 1]);
      r[2 * j] -= __bfloat162float(a);
      r[2 * j + 1] -= __bfloat162float(b);
      w[j] = (uint32_t)__bfloat16_as_ushort(a) | ((uint32_t)__bfloat16_as_ushort(b) << 16);
    }
    *reinterpret_cast<uint4*>(pc.p[k] + off) = make_uint4(w[0], w[1], w[2], w[3]);
  }
}

__device__ __forceinline__ void ld8(const float* p, float (&f)[8]) {
  const float4 a = *reinterpret_cast<const float4*>(p);
  const float4 b = *(reinterpret_cast<const float4*>(p) + 1);
  f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
}
__device__ __forceinline__ void st8(float* p, const float (&f)[8]) {
  *reinterpret_cast<float4*>(p) = make_float4(f[0], f[1], f[2], f[3]);
  *(reinterpret_cast<float4*>(p) + 1) = make_float4(f[4], f[5], f[6], f[7]);
}

// ------------------------------------------------------------------------------------------ split
__global__ void split_f32_kernel(const float* __restrict__ x, Pieces pc, long long n8) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
    float f[8];
    ld8(x + i * 8, f);
    store_pieces8(pc, i * 8, f);
  }
}

// ------------------------------------------------------------------------------------------ BN apply (+ split)
struct ActP {
  const float* y; long long y_cs;
  const float* res; long long res_cs;
  const float* scale; const float* shift;
  float* out; long long out_cs;       // fp32 output (nullable)
  Pieces pc; long long pc_cs;         // bf16 pieces of the output (pc.n may be 0)
  long long M; int C;
  int relu, drop_mode; uint32_t thresh16; float keep_scale;
  unsigned long long seed, offset; const long long* offset_dev; const unsigned char* mask;
};

__global__ void __launch_bounds__(256) bn_act_f32_kernel(const ActP p) {
  const int cg = p.C >> 3;  // 8-channel groups per row
  const long long total = p.M * cg;
  unsigned long long rng_off = p.offset;
  if (p.drop_mode == 1 && p.offset_dev) rng_off += (unsigned long long)*p.offset_dev;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long m = i / cg;
    const int c = (int)(i - m * cg) << 3;
    float f[8], sc[8], sh[8];
    ld8(p.y + m * p.y_cs + c, f);
    ld8(p.scale + c, sc);
    ld8(p.shift + c, sh);
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] = fmaf(f[j], sc[j], sh[j]);
    if (p.res) {
      float r[8];
      ld8(p.res + m * p.res_cs + c, r);
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] += r[j];
    }
    if (p.relu) {
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] = fmaxf(f[j], 0.f);
    }
    if (p.drop_mode == 1) {
      const uint32_t keep = pt_dropout_keep8(p.seed, rng_off, (uint64_t)(m * p.C + c), p.thresh16);
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
    if (p.out) st8(p.out + m * p.out_cs + c, f);
    if (p.pc.n) store_pieces8(p.pc, m * p.pc_cs + c, f);
  }
}

// ------------------------------------------------------------------------------------------ BN backward
// dz = dout * [act > 0] * grad_scale  (act = the layer's fp32 output after ReLU/Dropout, or the hi piece of it: the
// sign of a bf16 rounding is the sign of the value; relu = 0: no mask).  xhat = (y - mean) * invstd.
struct BwdP {
  const float* dout; long long dout_cs;
  const float* act; long long act_cs;            // fp32 mask source (nullable)
  const __nv_bfloat16* act_hi; long long hi_cs;  // bf16 mask source (nullable)
  const float* y; long long y_cs;
  const float* mean; const float* invstd; const float* scale;
  long long M; int C; int relu; float grad_scale; int training;
  double* sum_dz; double* sum_dzx;
  float* dy; long long dy_cs;                    // fp32 dy (nullable)
  Pieces pc; long long pc_cs;                    // bf16 pieces of dy (pc.n may be 0)
  float* dres; long long dres_cs;                // fp32 gradient of the residual input (= dz), nullable
  float* dgamma; float* dbeta; int C_real; int param_accumulate;
};

__device__ __forceinline__ void load_dz8(const BwdP& p, long long m, int c, float (&dz)[8]) {
  ld8(p.dout + m * p.dout_cs + c, dz);
  if (p.relu) {
    if (p.act) {
      float a[8];
      ld8(p.act + m * p.act_cs + c, a);
#pragma unroll
      for (int j = 0; j < 8; ++j) dz[j] = a[j] > 0.f ? dz[j] : 0.f;
    } else {
      const uint4 u = *reinterpret_cast<const uint4*>(p.act_hi + m * p.hi_cs + c);
      const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const uint32_t h = (w[j >> 1] >> (16 * (j & 1))) & 0xFFFFu;  // bf16 bits: positive iff sign clear and non-zero
        dz[j] = (h != 0 && (h & 0x8000u) == 0) ? dz[j] : 0.f;
      }
    }
  }
  if (p.grad_scale != 1.f) {
#pragma unroll
    for (int j = 0; j < 8; ++j) dz[j] *= p.grad_scale;
  }
}

// One block = 256 threads = G channel groups (8 channels each, G = min(C/8, 32)) x 256/G row lanes; grid = (C/8/G, slabs).
// A warp reads G*32 contiguous bytes of a row per tensor.  y == nullptr: plain per-channel sums of dout (bias gradients).
__global__ void __launch_bounds__(256) bn_bwd_reduce_f32_kernel(const BwdP p, int G) {
  __shared__ double red[256][17];
  const int gi = threadIdx.x % G, lane = threadIdx.x / G, lanes = 256 / G;
  const int c = (blockIdx.x * G + gi) << 3;
  float s1[8], s2[8], mu[8], is[8];
  double d1[8], d2[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { s1[j] = s2[j] = 0.f; d1[j] = d2[j] = 0.0; mu[j] = 0.f; is[j] = 1.f; }
  if (p.y) {
    ld8(p.mean + c, mu);
    ld8(p.invstd + c, is);
  }
  int run = 0;
  for (long long m = blockIdx.y * (long long)lanes + lane; m < p.M; m += (long long)gridDim.y * lanes) {
    float dz[8];
    load_dz8(p, m, c, dz);
    if (p.y) {
      float yv[8];
      ld8(p.y + m * p.y_cs + c, yv);
#pragma unroll
      for (int j = 0; j < 8; ++j) s2[j] = fmaf(dz[j], (yv[j] - mu[j]) * is[j], s2[j]);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) s1[j] += dz[j];
    if (++run == 32) {  // short fp32 runs, long accumulation in fp64 (the sums cancel)
#pragma unroll
      for (int j = 0; j < 8; ++j) { d1[j] += s1[j]; d2[j] += s2[j]; s1[j] = s2[j] = 0.f; }
      run = 0;
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    red[threadIdx.x][j] = d1[j] + s1[j];
    red[threadIdx.x][8 + j] = d2[j] + s2[j];
  }
  __syncthreads();
  for (int v = threadIdx.x; v < G * 16; v += blockDim.x) {
    const int g = v / 16, k = v % 16;
    double a = 0.0;
    for (int l = 0; l < lanes; ++l) a += red[l * G + g][k];
    const int ch = ((blockIdx.x * G + g) << 3) + (k & 7);
    if (k < 8) atomicAdd(p.sum_dz + ch, a);
    else if (p.y) atomicAdd(p.sum_dzx + ch, a);
  }
}

__global__ void __launch_bounds__(256) bn_bwd_apply_f32_kernel(const BwdP p) {
  const int cg = p.C >> 3;
  const long long total = p.M * cg;
  const float inv_m = 1.f / (float)p.M;
  if (blockIdx.x == 0 && p.dgamma) {
    for (int c = threadIdx.x; c < p.C_real; c += blockDim.x) {
      const float dg = (float)p.sum_dzx[c], db = (float)p.sum_dz[c];
      p.dgamma[c] = p.param_accumulate ? p.dgamma[c] + dg : dg;
      p.dbeta[c] = p.param_accumulate ? p.dbeta[c] + db : db;
    }
  }
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long m = i / cg;
    const int c = (int)(i - m * cg) << 3;
    float dz[8], sc[8];
    load_dz8(p, m, c, dz);
    if (p.dres) st8(p.dres + m * p.dres_cs + c, dz);
    ld8(p.scale + c, sc);
    float dy[8];
    if (p.training) {
      float yv[8], mu[8], is[8];
      ld8(p.y + m * p.y_cs + c, yv);
      ld8(p.mean + c, mu);
      ld8(p.invstd + c, is);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float xh = (yv[j] - mu[j]) * is[j];
        const float a = (float)p.sum_dz[c + j] * inv_m, b = (float)p.sum_dzx[c + j] * inv_m;
        dy[j] = sc[j] * (dz[j] - a - xh * b);
      }
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) dy[j] = sc[j] * dz[j];
    }
    if (p.dy) st8(p.dy + m * p.dy_cs + c, dy);
    if (p.pc.n) store_pieces8(p.pc, m * p.pc_cs + c, dy);
  }
}

// ------------------------------------------------------------------------------------------ max-pool
__global__ void maxpool_arg_f32_kernel(const float* __restrict__ x, float* __restrict__ y, unsigned char* __restrict__ arg,
                                       int N, int H, int W, int C, int Ho, int Wo, int k, int stride, int pad) {
  const long long total = (long long)N * Ho * Wo * C;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const long long m = i / C;
    const int q = (int)(m % Wo), pp = (int)((m / Wo) % Ho), n = (int)(m / ((long long)Wo * Ho));
    float best = -INFINITY;
    int slot = 255;
    for (int r = 0; r < k; ++r) {
      const int ih = pp * stride - pad + r;
      if (ih < 0 || ih >= H) continue;
      for (int s = 0; s < k; ++s) {
        const int iw = q * stride - pad + s;
        if (iw < 0 || iw >= W) continue;
        const float v = x[(((long long)n * H + ih) * W + iw) * C + c];
        if (v > best || slot == 255) { best = v; slot = r * k + s; }  // first maximum in window order, like ATen
      }
    }
    y[i] = best;
    arg[i] = (unsigned char)slot;
  }
}

__global__ void maxpool_bwd_f32_kernel(const float* __restrict__ dy, const unsigned char* __restrict__ arg,
                                       float* __restrict__ dx, int N, int H, int W, int C, int Ho, int Wo, int k,
                                       int stride, int pad) {
  const long long total = (long long)N * H * W * C;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const long long m = i / C;
    const int w = (int)(m % W), h = (int)((m / W) % H), n = (int)(m / ((long long)W * H));
    float a = 0.f;
    for (int r = 0; r < k; ++r) {
      const int t = h + pad - r;
      if (t < 0 || t % stride) continue;
      const int pp = t / stride;
      if (pp >= Ho) continue;
      for (int s = 0; s < k; ++s) {
        const int u = w + pad - s;
        if (u < 0 || u % stride) continue;
        const int q = u / stride;
        if (q >= Wo) continue;
        const long long o = (((long long)n * Ho + pp) * Wo + q) * C + c;
        if (arg[o] == r * k + s) a += dy[o];
      }
    }
    dx[i] = a;
  }
}

// ------------------------------------------------------------------------------------------ bilinear backward
__device__ __forceinline__ void ptblc(int o, float scale, int in_size, int& i0, int& i1, float& l1) {
  const float r = scale * o;
  i0 = (int)r;
  if (i0 > in_size - 1) i0 = in_size - 1;
  i1 = i0 + (i0 < in_size - 1 ? 1 : 0);
  l1 = r - i0;
}

// dx[n][h][w][c] (+)= sum over the output pixels whose footprint holds (h, w).  from_nchw: dy is [N][C][Ho][Wo] (the
// logits gradient) and threads run over w fastest; otherwise dy is NHWC and threads run over c fastest.
__global__ void bilinear_bwd_f32_kernel(const float* __restrict__ dy, float* __restrict__ dx, int N, int Hi, int Wi,
                                        int Ho, int Wo, int C, int dy_cs, int dx_cs, float sh, float sw, int from_nchw) {
  const long long total = (long long)N * Hi * Wi * C;
  const float ish = sh > 0.f ? 1.f / sh : 0.f, isw = sw > 0.f ? 1.f / sw : 0.f;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int c, w, h, n;
    if (from_nchw) {
      w = (int)(i % Wi); h = (int)((i / Wi) % Hi); c = (int)((i / ((long long)Wi * Hi)) % C);
      n = (int)(i / ((long long)Wi * Hi * C));
    } else {
      c = (int)(i % C); w = (int)((i / C) % Wi); h = (int)((i / ((long long)C * Wi)) % Hi);
      n = (int)(i / ((long long)C * Wi * Hi));
    }
    int oh_lo = sh > 0.f ? (int)floorf((h - 1) * ish) - 1 : 0, oh_hi = sh > 0.f ? (int)ceilf((h + 1) * ish) + 1 : Ho - 1;
    int ow_lo = sw > 0.f ? (int)floorf((w - 1) * isw) - 1 : 0, ow_hi = sw > 0.f ? (int)ceilf((w + 1) * isw) + 1 : Wo - 1;
    oh_lo = max(oh_lo, 0); oh_hi = min(oh_hi, Ho - 1);
    ow_lo = max(ow_lo, 0); ow_hi = min(ow_hi, Wo - 1);
    float a = 0.f;
    for (int oh = oh_lo; oh <= oh_hi; ++oh) {
      int y0, y1; float ly;
      ptblc(oh, sh, Hi, y0, y1, ly);
      const float wy = (y0 == h ? 1.f - ly : 0.f) + (y1 == h ? ly : 0.f);
      if (wy == 0.f) continue;
      float rowacc = 0.f;
      for (int ow = ow_lo; ow <= ow_hi; ++ow) {
        int x0, x1; float lx;
        ptblc(ow, sw, Wi, x0, x1, lx);
        const float wx = (x0 == w ? 1.f - lx : 0.f) + (x1 == w ? lx : 0.f);
        if (wx == 0.f) continue;
        const float g = from_nchw ? dy[(((long long)n * C + c) * Ho + oh) * Wo + ow]
                                  : dy[(((long long)n * Ho + oh) * Wo + ow) * dy_cs + c];
        rowacc = fmaf(wx, g, rowacc);
      }
      a = fmaf(wy, rowacc, a);
    }
    dx[(((long long)n * Hi + h) * Wi + w) * dx_cs + c] = a;
  }
}

// y[n][p][c] (+)= scale * x[n][c]
__global__ void spatial_broadcast_acc_f32_kernel(const float* __restrict__ x, float* __restrict__ y, int N, int HW, int C,
                                                 float scale, int accumulate) {
  const long long total = (long long)N * HW * C;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const int n = (int)(i / ((long long)HW * C));
    const float v = scale * x[(long long)n * C + c];
    y[i] = accumulate ? y[i] + v : v;
  }
}

}  // namespace zs3

using namespace zs3;
#define ST(s) static_cast<cudaStream_t>(s)
#define BF(p) static_cast<__nv_bfloat16*>(p)

static int fill_pieces(Pieces& pc, void* const* ptrs, int n, const char* what) {
  ZS3_CHECK_ARG(n >= 0 && n <= 3, "%s: 0..3 pieces", what);
  pc.n = n;
  for (int k = 0; k < 3; ++k) pc.p[k] = nullptr;
  for (int k = 0; k < n; ++k) {
    ZS3_CHECK_ARG(ptrs && ptrs[k] && (reinterpret_cast<uintptr_t>(ptrs[k]) & 15) == 0, "%s: piece %d null/unaligned", what, k);
    pc.p[k] = BF(ptrs[k]);
  }
  return ZS3_OK;
}

extern "C" int zs3_split_f32(const float* x, void* const* pieces, int n_pieces, long long n, void* stream) {
  ZS3_CHECK_ARG(x && n >= 0 && n % 8 == 0 && n_pieces >= 1 && (reinterpret_cast<uintptr_t>(x) & 15) == 0, "split_f32: bad args");
  Pieces pc;
  int rc = fill_pieces(pc, pieces, n_pieces, "split_f32");
  if (rc) return rc;
  if (n == 0) return ZS3_OK;
  split_f32_kernel<<<ptblocks(n / 8, 256), 256, 0, ST(stream)>>>(x, pc, n / 8);
  ZS3_CHECK_LAUNCH("split_f32");
  return ZS3_OK;
}

extern "C" int zs3_bn_act_f32(const zs3_bn_act_f32_args* a, void* stream) {
  ZS3_CHECK_ARG(a != nullptr, "bn_act_f32: null args");
  ZS3_CHECK_ARG(a->y && a->scale && a->shift && a->C > 0 && a->C % 8 == 0 && a->y_cstride % 4 == 0, "bn_act_f32: bad args");
  ZS3_CHECK_ARG(a->out || a->n_pieces > 0, "bn_act_f32: nothing to write");
  ZS3_CHECK_ARG(a->drop_mode >= 0 && a->drop_mode <= 2 && a->drop_p >= 0.f && a->drop_p < 1.f, "bn_act_f32: dropout");
  ZS3_CHECK_ARG(a->drop_mode != 2 || a->keep_mask != nullptr, "bn_act_f32: drop_mode 2 needs keep_mask");
  ActP p;
  memset(&p, 0, sizeof(p));
  int rc = fill_pieces(p.pc, a->pieces, a->n_pieces, "bn_act_f32");
  if (rc) return rc;
  p.y = a->y; p.y_cs = a->y_cstride;
  p.res = a->residual; p.res_cs = a->res_cstride;
  p.scale = a->scale; p.shift = a->shift;
  p.out = a->out; p.out_cs = a->out_cstride;
  p.pc_cs = a->piece_cstride;
  p.M = a->M; p.C = a->C; p.relu = a->relu;
  p.drop_mode = a->drop_p > 0.f ? a->drop_mode : 0;
  p.thresh16 = (uint32_t)(a->drop_p * 65536.0f + 0.5f);
  p.keep_scale = 1.f / (1.f - a->drop_p);
  p.seed = a->seed; p.offset = a->offset;
  p.offset_dev = static_cast<const long long*>(a->offset_dev);
  p.mask = static_cast<const unsigned char*>(a->keep_mask);
  if (a->M <= 0) return ZS3_OK;
  bn_act_f32_kernel<<<ptblocks(a->M * (a->C / 8), 256), 256, 0, ST(stream)>>>(p);
  ZS3_CHECK_LAUNCH("bn_act_f32");
  return ZS3_OK;
}

static int fill_bwd(BwdP& p, const zs3_bn_bwd_f32_args* a) {
  memset(&p, 0, sizeof(p));
  int rc = fill_pieces(p.pc, a->dy_pieces, a->n_pieces, "bn_bwd_f32");
  if (rc) return rc;
  p.dout = a->dout; p.dout_cs = a->dout_cstride;
  p.act = a->act; p.act_cs = a->act_cstride;
  p.act_hi = static_cast<const __nv_bfloat16*>(a->act_hi); p.hi_cs = a->act_hi_cstride;
  p.y = a->y; p.y_cs = a->y_cstride;
  p.mean = a->mean; p.invstd = a->invstd; p.scale = a->scale;
  p.M = a->M; p.C = a->C; p.relu = a->relu; p.grad_scale = a->grad_scale; p.training = a->training;
  p.sum_dz = a->sum_dz; p.sum_dzx = a->sum_dzx;
  p.dy = a->dy; p.dy_cs = a->dy_cstride; p.pc_cs = a->piece_cstride;
  p.dres = a->dres; p.dres_cs = a->dres_cstride;
  p.dgamma = a->dgamma; p.dbeta = a->dbeta; p.C_real = a->C_real; p.param_accumulate = a->param_accumulate;
  return ZS3_OK;
}

extern "C" int zs3_bn_bwd_f32(const zs3_bn_bwd_f32_args* a, void* stream) {
  ZS3_CHECK_ARG(a != nullptr, "bn_bwd_f32: null args");
  ZS3_CHECK_ARG(a->dout && a->y && a->mean && a->invstd && a->scale && a->sum_dz && a->sum_dzx, "bn_bwd_f32: null pointer");
  ZS3_CHECK_ARG(a->C >= 64 && (a->C & (a->C - 1)) == 0, "bn_bwd_f32: C=%d must be a power of two >= 64", a->C);
  ZS3_CHECK_ARG(!a->relu || a->act || a->act_hi, "bn_bwd_f32: relu needs the forward output (fp32 or its hi piece)");
  ZS3_CHECK_ARG(a->dy || a->n_pieces > 0, "bn_bwd_f32: nothing to write");
  ZS3_CHECK_ARG((a->dgamma == nullptr) == (a->dbeta == nullptr) && (!a->dgamma || a->C_real <= a->C), "bn_bwd_f32: dgamma/dbeta");
  BwdP p;
  int rc = fill_bwd(p, a);
  if (rc) return rc;
  if (a->M <= 0) return ZS3_OK;
  cudaStream_t st = ST(stream);
  cudaMemsetAsync(a->sum_dz, 0, sizeof(double) * a->C, st);
  cudaMemsetAsync(a->sum_dzx, 0, sizeof(double) * a->C, st);
  const int cg = a->C / 8, G = cg < 32 ? cg : 32, lanes = 256 / G;
  long long slabs = (a->M + lanes * 16 - 1) / (lanes * 16);  // >= 16 rows per row lane
  const long long cap = (148 * 4) / (cg / G) > 0 ? (148 * 4) / (cg / G) : 1;
  if (slabs > cap) slabs = cap;
  if (slabs < 1) slabs = 1;
  bn_bwd_reduce_f32_kernel<<<dim3(cg / G, (unsigned)slabs), 256, 0, st>>>(p, G);
  ZS3_CHECK_LAUNCH("bn_bwd_reduce_f32");
  bn_bwd_apply_f32_kernel<<<ptblocks(a->M * cg, 256), 256, 0, st>>>(p);
  ZS3_CHECK_LAUNCH("bn_bwd_apply_f32");
  return ZS3_OK;
}

extern "C" int zs3_channel_sums_f32(const float* x, long long x_cstride, long long M, int C, double* out, void* stream) {
  ZS3_CHECK_ARG(x && out && C >= 64 && (C & (C - 1)) == 0 && x_cstride >= C, "channel_sums_f32: bad args (C a power of two >= 64)");
  BwdP p;
  memset(&p, 0, sizeof(p));
  p.dout = x; p.dout_cs = x_cstride; p.M = M; p.C = C; p.grad_scale = 1.f; p.sum_dz = out; p.sum_dzx = out;
  cudaStream_t st = ST(stream);
  cudaMemsetAsync(out, 0, sizeof(double) * C, st);
  if (M <= 0) return ZS3_OK;
  const int cg = C / 8, G = cg < 32 ? cg : 32, lanes = 256 / G;
  long long slabs = (M + lanes * 16 - 1) / (lanes * 16);
  const long long cap = (148 * 4) / (cg / G) > 0 ? (148 * 4) / (cg / G) : 1;
  if (slabs > cap) slabs = cap;
  bn_bwd_reduce_f32_kernel<<<dim3(cg / G, (unsigned)slabs), 256, 0, st>>>(p, G);
  ZS3_CHECK_LAUNCH("channel_sums_f32");
  return ZS3_OK;
}

extern "C" int zs3_maxpool_arg_f32(const float* x, float* y, unsigned char* argmax, int N, int H, int W, int C, int Ho,
                                   int Wo, int k, int stride, int pad, void* stream) {
  ZS3_CHECK_ARG(x && y && argmax && k >= 1 && k * k < 255 && stride >= 1, "maxpool_arg_f32: bad args");
  maxpool_arg_f32_kernel<<<ptblocks((long long)N * Ho * Wo * C, 256), 256, 0, ST(stream)>>>(x, y, argmax, N, H, W, C, Ho,
                                                                                            Wo, k, stride, pad);
  ZS3_CHECK_LAUNCH("maxpool_arg_f32");
  return ZS3_OK;
}

extern "C" int zs3_maxpool_bwd_f32(const float* dy, const unsigned char* argmax, float* dx, int N, int H, int W, int C,
                                   int Ho, int Wo, int k, int stride, int pad, void* stream) {
  ZS3_CHECK_ARG(dy && dx && argmax && k >= 1 && stride >= 1, "maxpool_bwd_f32: bad args");
  maxpool_bwd_f32_kernel<<<ptblocks((long long)N * H * W * C, 256), 256, 0, ST(stream)>>>(dy, argmax, dx, N, H, W, C, Ho, Wo,
                                                                                          k, stride, pad);
  ZS3_CHECK_LAUNCH("maxpool_bwd_f32");
  return ZS3_OK;
}

static float ptblscale(int in, int out) { return out > 1 ? (float)(in - 1) / (float)(out - 1) : 0.f; }

extern "C" int zs3_bilinear_bwd_f32(const float* dy, float* dx, int N, int Hi, int Wi, int Ho, int Wo, int C, int dy_cs,
                                    int dx_cs, int from_nchw, void* stream) {
  ZS3_CHECK_ARG(dy && dx && C > 0 && C <= dx_cs && (from_nchw || C <= dy_cs), "bilinear_bwd_f32: bad args");
  bilinear_bwd_f32_kernel<<<ptblocks((long long)N * Hi * Wi * C, 256), 256, 0, ST(stream)>>>(
      dy, dx, N, Hi, Wi, Ho, Wo, C, dy_cs, dx_cs, ptblscale(Hi, Ho), ptblscale(Wi, Wo), from_nchw);
  ZS3_CHECK_LAUNCH("bilinear_bwd_f32");
  return ZS3_OK;
}

extern "C" int zs3_spatial_broadcast_acc_f32(const float* x, float* y, int N, int HW, int C, float scale, int accumulate,
                                             void* stream) {
  ZS3_CHECK_ARG(x && y, "spatial_broadcast_acc_f32: bad args");
  spatial_broadcast_acc_f32_kernel<<<ptblocks((long long)N * HW * C, 256), 256, 0, ST(stream)>>>(x, y, N, HW, C, scale,
                                                                                                 accumulate);
  ZS3_CHECK_LAUNCH("spatial_broadcast_acc_f32");
  return ZS3_OK;
}
