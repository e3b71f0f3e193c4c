// HBM-bound glue kernels of the DeepLab path (NHWC bf16 activations, fp32 math):
// stem im2col, max-pool, bilinear resampling, global average pool, cross-entropy, fused optimizers.
// Each entry point cites the reference op it replaces in include/zs3b200.h.
#include "common.cuh"
#include "ptx.cuh"

namespace zs3 {

__device__ __forceinline__ void unpack8m(const uint4& u, float (&f)[8]) {
  f[0] = bf16_lo(u.x); f[1] = bf16_hi(u.x); f[2] = bf16_lo(u.y); f[3] = bf16_hi(u.y);
  f[4] = bf16_lo(u.z); f[5] = bf16_hi(u.z); f[6] = bf16_lo(u.w); f[7] = bf16_hi(u.w);
}
__device__ __forceinline__ uint4 pack8m(const float (&f)[8]) {
  uint4 u;
  u.x = pack_bf16x2(f[0], f[1]); u.y = pack_bf16x2(f[2], f[3]);
  u.z = pack_bf16x2(f[4], f[5]); u.w = pack_bf16x2(f[6], f[7]);
  return u;
}

static int ew_blocks(long long items, int threads, int cap = 148 * 16) {
  long long b = (items + threads - 1) / threads;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

// ------------------------------------------------------------------------------ stem im2col
// cols[m][k], k = c*R*R + r*R + s (the OIHW flattening of the weight), zero padded to kpad columns.
// One CTA per (image, output row): the R input rows it needs are staged in shared memory (zero-padded borders),
// then every thread owns one PAIR of k columns and walks along the output row: coalesced 4-byte stores.
__global__ void __launch_bounds__(192) stem_im2col_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ cols,
                                                          int C, int H, int W, int R, int stride, int pad, int Ho,
                                                          int Wo, int kpad, int krsc) {
  extern __shared__ float rows[];  // [C][R][W + 2*pad]
  const int Wp = W + 2 * pad;
  const int p = blockIdx.x % Ho, n = blockIdx.x / Ho;
  const int K = C * R * R;
  for (int i = threadIdx.x; i < C * R * Wp; i += blockDim.x) {
    const int wv = i % Wp, r = (i / Wp) % R, c = i / (Wp * R);
    const int ih = p * stride - pad + r, iw = wv - pad;
    rows[i] = (ih >= 0 && ih < H && iw >= 0 && iw < W) ? __ldg(x + (((long long)n * C + c) * H + ih) * W + iw) : 0.f;
  }
  __syncthreads();
  const int pairs = kpad >> 1;                  // k pairs per pixel
  const int ppi = blockDim.x / pairs;           // pixels per iteration
  const int kp = threadIdx.x % pairs, ql = threadIdx.x / pairs;
  int off[2];
  bool live[2];
#pragma unroll
  for (int e = 0; e < 2; ++e) {
    const int k = 2 * kp + e;
    live[e] = k < K;
    // k = c*R*R + r*R + s (OIHW flattening) or (r*R + s)*C + c (KRSC / channels_last flattening)
    const int c = krsc ? k % C : k / (R * R);
    const int rs = krsc ? k / C : k - c * R * R;
    const int r = rs / R, sft = rs - r * R;
    off[e] = live[e] ? (c * R + r) * Wp + sft : 0;
  }
  if (ql >= ppi) return;
  __nv_bfloat16* out = cols + ((long long)n * Ho + p) * Wo * kpad;
  for (int q = ql; q < Wo; q += ppi) {
    const float v0 = live[0] ? rows[off[0] + q * stride] : 0.f;
    const float v1 = live[1] ? rows[off[1] + q * stride] : 0.f;
    *reinterpret_cast<uint32_t*>(out + (long long)q * kpad + 2 * kp) = pack_bf16x2(v0, v1);
  }
}

// KRSC flavour (k = (r*R + s)*C + c, the order of channels_last weights -- what the model uses): with the staged input
// rows laid out [r][w][c], the R*C... S*C values of one tap row r of an output pixel are CONTIGUOUS in shared memory
// (start (q*stride)*C), so a thread produces 8 consecutive k (one 16-byte store) from precomputed offsets.
__global__ void __launch_bounds__(256) stem_im2col_krsc_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ cols,
                                                               int C, int H, int W, int R, int stride, int pad, int Ho,
                                                               int Wo, int kpad) {
  extern __shared__ float rows[];  // [R][W + 2*pad][C]
  const int Wp = W + 2 * pad;
  const int p = blockIdx.x % Ho, n = blockIdx.x / Ho;
  const int K = C * R * R, SC = R * C;
  // one (channel, filter row) source line at a time, consecutive threads on consecutive columns: coalesced reads and
  // no per-element index arithmetic (the first version derived (c, r, w) from a flat index with three divisions per
  // element, which made the staging loop 6x the instruction count of the im2col expansion itself)
  for (int cr = 0; cr < C * R; ++cr) {
    const int r = cr % R, c = cr / R;
    const int ih = p * stride - pad + r;
    const bool row_ok = ih >= 0 && ih < H;
    const float* src = x + (((long long)n * C + c) * H + (row_ok ? ih : 0)) * W;
    float* dst = rows + (long long)r * Wp * C + c;
    for (int wv = threadIdx.x; wv < Wp; wv += blockDim.x) {
      const int iw = wv - pad;
      dst[wv * C] = (row_ok && iw >= 0 && iw < W) ? __ldg(src + iw) : 0.f;
    }
  }
  __syncthreads();
  const int groups = kpad >> 3;                 // 8-column groups per pixel
  const int ppi = blockDim.x / groups;          // pixels per iteration
  const int kg = threadIdx.x % groups, ql = threadIdx.x / groups;
  if (ql >= ppi) return;
  int off[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const int k = 8 * kg + e;
    const int r = k / SC;
    off[e] = k < K ? r * Wp * C + (k - r * SC) : -1;
  }
  __nv_bfloat16* out = cols + ((long long)n * Ho + p) * Wo * kpad + 8 * kg;
  const int qstep = stride * C;
  for (int q = ql; q < Wo; q += ppi) {
    const float* base = rows + q * qstep;
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] = off[e] >= 0 ? base[off[e]] : 0.f;
    *reinterpret_cast<uint4*>(out + (long long)q * kpad) = pack8m(v);
  }
}

// --------------------------------------------------------------------------------- max-pool
// first maximum in row-major window order wins (ATen max_pool2d semantics); argmax stores the window slot.
// IdxT = int when the flat item count fits in 31 bits (always on this path): the index arithmetic is 32-bit divisions
// instead of emulated 64-bit ones, which were most of the instructions of these map kernels.
template <typename IdxT>
__global__ void maxpool_fwd_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y,
                                   unsigned char* __restrict__ arg, int N, int H, int W, int C, int Ho, int Wo, int k,
                                   int stride, int pad) {
  const int vpc = C >> 3;
  const IdxT total = (IdxT)N * Ho * Wo * vpc;
  for (IdxT i = blockIdx.x * (IdxT)blockDim.x + threadIdx.x; i < total; i += (IdxT)gridDim.x * blockDim.x) {
    const int c = (int)(i % vpc) << 3;
    const IdxT m = i / vpc;
    const int q = (int)(m % Wo), p = (int)((m / Wo) % Ho), n = (int)(m / ((IdxT)Wo * Ho));
    float best[8];
    int bi[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { best[j] = -INFINITY; bi[j] = 0; }
    for (int r = 0; r < k; ++r) {
      const int ih = p * stride - pad + r;
      if (ih < 0 || ih >= H) continue;
      for (int s = 0; s < k; ++s) {
        const int iw = q * stride - pad + s;
        if (iw < 0 || iw >= W) continue;
        float v[8];
        unpack8m(*reinterpret_cast<const uint4*>(x + (((long long)n * H + ih) * W + iw) * C + c), v);
#pragma unroll
        for (int j = 0; j < 8; ++j)
          if (v[j] > best[j]) { best[j] = v[j]; bi[j] = r * k + s; }
      }
    }
    *reinterpret_cast<uint4*>(y + m * C + c) = pack8m(best);
    uint2 a;
    a.x = bi[0] | (bi[1] << 8) | (bi[2] << 16) | (bi[3] << 24);
    a.y = bi[4] | (bi[5] << 8) | (bi[6] << 16) | (bi[7] << 24);
    *reinterpret_cast<uint2*>(arg + m * C + c) = a;
  }
}

// gather form: every input pixel sums the output gradients of the windows that selected it.  One CTA per input row:
// the (at most 8) output rows whose windows cover it are resolved once and STAGED in shared memory (gradient row +
// argmax-slot row, coalesced 16-byte copies), so every output pixel is fetched once per CTA instead of once per input
// pixel it covers (k/stride * k/stride times through L1/L2 in the first version: 187 us for the 135 MB stem gradient).
__global__ void __launch_bounds__(256) maxpool_bwd_kernel(const __nv_bfloat16* __restrict__ dy,
                                                          const unsigned char* __restrict__ arg,
                                                          __nv_bfloat16* __restrict__ dx, int N, int H, int W, int C,
                                                          int Ho, int Wo, int k, int stride, int pad, int max_rows) {
  extern __shared__ __align__(16) unsigned char mp_smem[];
  const int vpc = C >> 3;
  const int h = blockIdx.x % H, n = blockIdx.x / H;
  int prow[8], rsel[8], np = 0;
  for (int r = 0; r < k; ++r) {
    const int t = h + pad - r;
    if (t < 0 || t % stride) continue;
    const int p = t / stride;
    if (p < Ho && np < max_rows) {
      prow[np] = p;
      rsel[np] = r;
      ++np;
    }
  }
  const int row_bytes_g = Wo * C * 2, row_bytes_a = Wo * C;                       // multiples of 16 and 8 (C % 8 == 0)
  uint4* sg = reinterpret_cast<uint4*>(mp_smem);                                   // [max_rows][Wo*C] bf16
  uint2* sa = reinterpret_cast<uint2*>(mp_smem + (size_t)max_rows * row_bytes_g);  // [max_rows][Wo*C] bytes
  for (int a = 0; a < np; ++a) {
    const long long obase = ((long long)n * Ho + prow[a]) * Wo * C;
    const uint4* gsrc = reinterpret_cast<const uint4*>(dy + obase);
    const uint2* asrc = reinterpret_cast<const uint2*>(arg + obase);
    for (int i = threadIdx.x; i < row_bytes_g / 16; i += blockDim.x) sg[a * (row_bytes_g / 16) + i] = gsrc[i];
    for (int i = threadIdx.x; i < row_bytes_a / 8; i += blockDim.x) sa[a * (row_bytes_a / 8) + i] = asrc[i];
  }
  __syncthreads();
  const int rowlen = W * vpc;
  __nv_bfloat16* drow = dx + ((long long)n * H + h) * W * C;
  for (int i = threadIdx.x; i < rowlen; i += blockDim.x) {
    const int w = i / vpc, cv = i - w * vpc;
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    for (int s = 0; s < k; ++s) {
      const int u = w + pad - s;
      if (u < 0 || u % stride) continue;
      const int q = u / stride;
      if (q >= Wo) continue;
      for (int a = 0; a < np; ++a) {
        const int e = q * vpc + cv;  // 8-channel group index inside the staged row
        const uint2 sel2 = sa[a * (row_bytes_a / 8) + e];
        float g[8];
        unpack8m(sg[a * (row_bytes_g / 16) + e], g);
        const unsigned slot = (unsigned)(rsel[a] * k + s);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const unsigned sel = ((j < 4 ? sel2.x : sel2.y) >> (8 * (j & 3))) & 0xFFu;
          if (sel == slot) acc[j] += g[j];
        }
      }
    }
    *reinterpret_cast<uint4*>(drow + (long long)w * C + (cv << 3)) = pack8m(acc);
  }
}

// ------------------------------------------------------------------ bilinear (align_corners)
__device__ __forceinline__ void bl_coord(int o, float scale, int in_size, int& i0, int& i1, float& l1) {
  const float r = scale * o;  // ATen: area_pixel_compute_source_index(align_corners=True)
  i0 = (int)r;
  if (i0 > in_size - 1) i0 = in_size - 1;
  i1 = i0 + (i0 < in_size - 1 ? 1 : 0);
  l1 = r - i0;
}

template <typename IdxT>
__global__ void bilinear_fwd_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y, int N, int Hi,
                                    int Wi, int Ho, int Wo, int C, int x_cs, int y_cs, float sh, float sw) {
  const int vpc = C >> 3;
  const IdxT total = (IdxT)N * Ho * Wo * vpc;
  for (IdxT i = blockIdx.x * (IdxT)blockDim.x + threadIdx.x; i < total; i += (IdxT)gridDim.x * blockDim.x) {
    const int c = (int)(i % vpc) << 3;
    const IdxT m = i / vpc;
    const int ow = (int)(m % Wo), oh = (int)((m / Wo) % Ho), n = (int)(m / ((IdxT)Wo * Ho));
    int y0, y1, x0, x1;
    float ly, lx;
    bl_coord(oh, sh, Hi, y0, y1, ly);
    bl_coord(ow, sw, Wi, x0, x1, lx);
    const __nv_bfloat16* base = x + (long long)n * Hi * Wi * x_cs + c;
    float a[8], b[8], d[8], e[8], o[8];
    unpack8m(*reinterpret_cast<const uint4*>(base + ((long long)y0 * Wi + x0) * x_cs), a);
    unpack8m(*reinterpret_cast<const uint4*>(base + ((long long)y0 * Wi + x1) * x_cs), b);
    unpack8m(*reinterpret_cast<const uint4*>(base + ((long long)y1 * Wi + x0) * x_cs), d);
    unpack8m(*reinterpret_cast<const uint4*>(base + ((long long)y1 * Wi + x1) * x_cs), e);
#pragma unroll
    for (int j = 0; j < 8; ++j)
      o[j] = (1.f - ly) * ((1.f - lx) * a[j] + lx * b[j]) + ly * ((1.f - lx) * d[j] + lx * e[j]);
    *reinterpret_cast<uint4*>(y + m * y_cs + c) = pack8m(o);
  }
}

// gather form of the transpose: dx[h][w] = sum over the output pixels whose stencil touches (h,w)
template <typename IdxT>
__global__ void bilinear_bwd_kernel(const __nv_bfloat16* __restrict__ dy, __nv_bfloat16* __restrict__ dx, int N,
                                    int Hi, int Wi, int Ho, int Wo, int C, int dy_cs, int dx_cs, float sh, float sw,
                                    int accumulate) {
  const int vpc = C >> 3;
  const IdxT total = (IdxT)N * Hi * Wi * vpc;
  const float ish = sh > 0.f ? 1.f / sh : 0.f, isw = sw > 0.f ? 1.f / sw : 0.f;
  for (IdxT i = blockIdx.x * (IdxT)blockDim.x + threadIdx.x; i < total; i += (IdxT)gridDim.x * blockDim.x) {
    const int c = (int)(i % vpc) << 3;
    const IdxT m = i / vpc;
    const int w = (int)(m % Wi), h = (int)((m / Wi) % Hi), n = (int)(m / ((IdxT)Wi * Hi));
    int oh_lo = sh > 0.f ? (int)floorf((h - 1) * ish) - 1 : 0, oh_hi = sh > 0.f ? (int)ceilf((h + 1) * ish) + 1 : Ho - 1;
    int ow_lo = sw > 0.f ? (int)floorf((w - 1) * isw) - 1 : 0, ow_hi = sw > 0.f ? (int)ceilf((w + 1) * isw) + 1 : Wo - 1;
    oh_lo = max(oh_lo, 0); oh_hi = min(oh_hi, Ho - 1);
    ow_lo = max(ow_lo, 0); ow_hi = min(ow_hi, Wo - 1);
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    for (int oh = oh_lo; oh <= oh_hi; ++oh) {
      int y0, y1; float ly;
      bl_coord(oh, sh, Hi, y0, y1, ly);
      const float wy = (y0 == h ? 1.f - ly : 0.f) + (y1 == h ? ly : 0.f);
      if (wy == 0.f) continue;
      for (int ow = ow_lo; ow <= ow_hi; ++ow) {
        int x0, x1; float lx;
        bl_coord(ow, sw, Wi, x0, x1, lx);
        const float wx = (x0 == w ? 1.f - lx : 0.f) + (x1 == w ? lx : 0.f);
        if (wx == 0.f) continue;
        float g[8];
        unpack8m(*reinterpret_cast<const uint4*>(dy + (((long long)n * Ho + oh) * Wo + ow) * dy_cs + c), g);
        const float wgt = wy * wx;
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] += wgt * g[j];
      }
    }
    uint4* dst = reinterpret_cast<uint4*>(dx + m * dx_cs + c);
    if (accumulate) {
      float o[8];
      unpack8m(*dst, o);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += o[j];
    }
    *dst = pack8m(acc);
  }
}

// final x4 upsample of the class scores: NHWC bf16 [N][Hi][Wi][cs] -> NCHW fp32 [N][C][Ho][Wo]
// one CTA per output row; the vertically blended input row is staged in shared memory as [c][w_in].
__global__ void upsample_logits_fwd_kernel(const __nv_bfloat16* __restrict__ x, float* __restrict__ y, int C, int Hi,
                                           int Wi, int cs, int Ho, int Wo, float sh, float sw) {
  extern __shared__ float rb[];  // [C][Wi+1]
  const int oh = blockIdx.x % Ho, n = blockIdx.x / Ho;
  int y0, y1; float ly;
  bl_coord(oh, sh, Hi, y0, y1, ly);
  const int ld = Wi + 1;
  const __nv_bfloat16* r0 = x + ((long long)n * Hi + y0) * Wi * cs;
  const __nv_bfloat16* r1 = x + ((long long)n * Hi + y1) * Wi * cs;
  for (int i = threadIdx.x; i < Wi * C; i += blockDim.x) {
    const int c = i % C, w = i / C;
    rb[c * ld + w] = (1.f - ly) * __bfloat162float(r0[(long long)w * cs + c]) + ly * __bfloat162float(r1[(long long)w * cs + c]);
  }
  __syncthreads();
  float* out = y + ((long long)n * C * Ho + oh) * Wo;
  for (int i = threadIdx.x; i < C * Wo; i += blockDim.x) {
    const int ow = i % Wo, c = i / Wo;
    int x0, x1; float lx;
    bl_coord(ow, sw, Wi, x0, x1, lx);
    out[(long long)c * Ho * Wo + ow] = (1.f - lx) * rb[c * ld + x0] + lx * rb[c * ld + x1];
  }
}

// backward: NCHW fp32 dlogits -> NHWC bf16 [N][Hi][Wi][cs]; one CTA per input row:
// stage 1 blends the contributing output rows vertically into shared memory v[c][ow], stage 2 reduces horizontally.
__global__ void upsample_logits_bwd_kernel(const float* __restrict__ dy, __nv_bfloat16* __restrict__ dx, int C, int Hi,
                                           int Wi, int cs, int Ho, int Wo, float sh, float sw) {
  extern __shared__ float v[];  // [C][Wo+1]
  const int h = blockIdx.x % Hi, n = blockIdx.x / Hi;
  const int ld = Wo + 1;
  const float ish = sh > 0.f ? 1.f / sh : 0.f, isw = sw > 0.f ? 1.f / sw : 0.f;
  int oh_lo = sh > 0.f ? (int)floorf((h - 1) * ish) - 1 : 0, oh_hi = sh > 0.f ? (int)ceilf((h + 1) * ish) + 1 : Ho - 1;
  oh_lo = max(oh_lo, 0); oh_hi = min(oh_hi, Ho - 1);
  for (int i = threadIdx.x; i < C * Wo; i += blockDim.x) {
    const int ow = i % Wo, c = i / Wo;
    float acc = 0.f;
    for (int oh = oh_lo; oh <= oh_hi; ++oh) {
      int y0, y1; float ly;
      bl_coord(oh, sh, Hi, y0, y1, ly);
      const float wy = (y0 == h ? 1.f - ly : 0.f) + (y1 == h ? ly : 0.f);
      if (wy != 0.f) acc += wy * __ldg(dy + (((long long)n * C + c) * Ho + oh) * Wo + ow);
    }
    v[c * ld + ow] = acc;
  }
  __syncthreads();
  __nv_bfloat16* out = dx + ((long long)n * Hi + h) * Wi * cs;
  for (int i = threadIdx.x; i < Wi * cs; i += blockDim.x) {
    const int c = i % cs, w = i / cs;
    float acc = 0.f;
    if (c < C) {
      int ow_lo = sw > 0.f ? (int)floorf((w - 1) * isw) - 1 : 0, ow_hi = sw > 0.f ? (int)ceilf((w + 1) * isw) + 1 : Wo - 1;
      ow_lo = max(ow_lo, 0); ow_hi = min(ow_hi, Wo - 1);
      for (int ow = ow_lo; ow <= ow_hi; ++ow) {
        int x0, x1; float lx;
        bl_coord(ow, sw, Wi, x0, x1, lx);
        const float wx = (x0 == w ? 1.f - lx : 0.f) + (x1 == w ? lx : 0.f);
        if (wx != 0.f) acc += wx * v[c * ld + ow];
      }
    }
    out[(long long)w * cs + c] = __float2bfloat16(acc);
  }
}

// ------------------------------------------------------------------ global average pooling
// y[n][c] = scale * sum_hw x[n][hw][c]   (one CTA per (n, 64-channel group))
__global__ void spatial_sum_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y, int HW, int x_cs,
                                   int y_cs, float scale) {
  __shared__ float red[4][64];
  const int n = blockIdx.y, c = blockIdx.x * 64 + (threadIdx.x & 63);
  const int lane_row = threadIdx.x >> 6;  // 4 row lanes with 256 threads
  float acc = 0.f;
  for (int p = lane_row; p < HW; p += 4) acc += __bfloat162float(x[((long long)n * HW + p) * x_cs + c]);
  red[lane_row][threadIdx.x & 63] = acc;
  __syncthreads();
  if (threadIdx.x < 64) {
    const float s = red[0][threadIdx.x] + red[1][threadIdx.x] + red[2][threadIdx.x] + red[3][threadIdx.x];
    y[(long long)n * y_cs + c] = __float2bfloat16(s * scale);
  }
}

// y[n][hw][c] (+)= scale * x[n][c]
__global__ void spatial_broadcast_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y, int N,
                                         int HW, int C, int x_cs, int y_cs, float scale, int accumulate) {
  const int vpc = C >> 3;
  const long long total = (long long)N * HW * vpc;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % vpc) << 3;
    const long long m = i / vpc;
    const int n = (int)(m / HW);
    float f[8];
    unpack8m(*reinterpret_cast<const uint4*>(x + (long long)n * x_cs + c), f);
    uint4* dst = reinterpret_cast<uint4*>(y + m * y_cs + c);
    float o[8];
    if (accumulate) unpack8m(*dst, o);
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] = f[j] * scale + (accumulate ? o[j] : 0.f);
    *dst = pack8m(f);
  }
}

// ----------------------------------------------------------------------------- cross-entropy
// accum[0] += sum_i w[t_i] * (-log softmax(logit_i)[t_i]),  accum[1] += sum_i w[t_i]   (t_i != ignore)
__global__ void ce_fwd_kernel(const float* __restrict__ logit, const float* __restrict__ target,
                              const float* __restrict__ weight, int C, long long HW, long long total, int ignore,
                              double* __restrict__ accum) {
  float num = 0.f, den = 0.f;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int t = (int)target[i];
    if (t == ignore || t < 0 || t >= C) continue;
    const long long n = i / HW, hw = i - n * HW;
    const float* lp = logit + n * C * HW + hw;
    float mx = -INFINITY, s = 0.f, lt = 0.f;
    for (int c = 0; c < C; ++c) {
      const float v = __ldg(lp + (long long)c * HW);
      if (c == t) lt = v;
      if (v > mx) { s = s * __expf(mx - v); mx = v; }
      s += __expf(v - mx);
    }
    const float w = weight ? weight[t] : 1.f;
    num += w * (mx + __logf(s) - lt);
    den += w;
  }
  __shared__ float sn[32], sd[32];
  for (int off = 16; off; off >>= 1) {
    num += __shfl_xor_sync(0xffffffffu, num, off);
    den += __shfl_xor_sync(0xffffffffu, den, off);
  }
  if ((threadIdx.x & 31) == 0) { sn[threadIdx.x >> 5] = num; sd[threadIdx.x >> 5] = den; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0, b = 0;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) { a += sn[i]; b += sd[i]; }
    atomicAdd(accum, a);
    atomicAdd(accum + 1, b);
  }
}

__global__ void ce_finalize_kernel(const double* accum, float div, float* loss) {
  *loss = (float)(accum[0] / accum[1] / (double)div);
}

// dlogit = gout * w[t] * (softmax - onehot) / (sum_w * div); 0 for ignored pixels
__global__ void ce_bwd_kernel(const float* __restrict__ logit, const float* __restrict__ target,
                              const float* __restrict__ weight, int C, long long HW, long long total, int ignore,
                              const double* __restrict__ accum, float div, const float* __restrict__ gout,
                              float* __restrict__ dlogit) {
  const float g = gout[0] / ((float)accum[1] * div);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int t = (int)target[i];
    const long long n = i / HW, hw = i - n * HW;
    const float* lp = logit + n * C * HW + hw;
    float* dp = dlogit + n * C * HW + hw;
    if (t == ignore || t < 0 || t >= C) {
      for (int c = 0; c < C; ++c) dp[(long long)c * HW] = 0.f;
      continue;
    }
    float mx = -INFINITY, s = 0.f;
    for (int c = 0; c < C; ++c) {
      const float v = __ldg(lp + (long long)c * HW);
      if (v > mx) { s = s * __expf(mx - v); mx = v; }
      s += __expf(v - mx);
    }
    const float inv = 1.f / s;
    const float gw = g * (weight ? weight[t] : 1.f);
    for (int c = 0; c < C; ++c) {
      const float pr = __expf(__ldg(lp + (long long)c * HW) - mx) * inv;
      dp[(long long)c * HW] = gw * (pr - (c == t ? 1.f : 0.f));
    }
  }
}

// ------------------------------------------------- fused x4 upsample + cross-entropy (training loss)
// The training step never needs the full-resolution class scores themselves: loss = CE(upsample(x), target).  These
// two kernels evaluate the bilinear upsample on the fly (same blend order as upsample_logits_fwd_kernel: vertical
// into fp32, then horizontal), so the [N][C][Ho][Wo] fp32 logits and their gradient (2 x 354 MB at bs=16, 513x513)
// are never written to or read from HBM.  The backward keeps one accumulator per class in registers: instantiated for
// C <= 24 (VOC: 21; softmax terms also stay in registers) and C <= 64 (Pascal-Context: 60; softmax terms re-read from
// the blended row in shared memory).
constexpr int CE_MAX_C = 64;

__global__ void __launch_bounds__(256) upsample_ce_fwd_kernel(const __nv_bfloat16* __restrict__ x,
                                                              const float* __restrict__ target,
                                                              const float* __restrict__ weight, int C, int Hi, int Wi,
                                                              int cs, int Ho, int Wo, float sh, float sw, int ignore,
                                                              double* __restrict__ accum) {
  extern __shared__ float rb[];  // [C][Wi+1] vertically blended input row
  const int oh = blockIdx.x % Ho, n = blockIdx.x / Ho;
  int y0, y1; float ly;
  bl_coord(oh, sh, Hi, y0, y1, ly);
  const int ld = Wi + 1;
  const __nv_bfloat16* r0 = x + ((long long)n * Hi + y0) * Wi * cs;
  const __nv_bfloat16* r1 = x + ((long long)n * Hi + y1) * Wi * cs;
  for (int i = threadIdx.x; i < Wi * C; i += blockDim.x) {
    const int c = i % C, w = i / C;
    rb[c * ld + w] = (1.f - ly) * __bfloat162float(r0[(long long)w * cs + c]) + ly * __bfloat162float(r1[(long long)w * cs + c]);
  }
  __syncthreads();
  const float* trow = target + ((long long)n * Ho + oh) * Wo;
  float num = 0.f, den = 0.f;
  for (int ow = threadIdx.x; ow < Wo; ow += blockDim.x) {
    const int t = (int)trow[ow];
    if (t == ignore || t < 0 || t >= C) continue;
    int x0, x1; float lx;
    bl_coord(ow, sw, Wi, x0, x1, lx);
    float mx = -INFINITY, s = 0.f, lt = 0.f;
    for (int c = 0; c < C; ++c) {
      const float v = (1.f - lx) * rb[c * ld + x0] + lx * rb[c * ld + x1];
      if (c == t) lt = v;
      if (v > mx) { s = s * __expf(mx - v); mx = v; }
      s += __expf(v - mx);
    }
    const float w = weight ? weight[t] : 1.f;
    num += w * (mx + __logf(s) - lt);
    den += w;
  }
  __shared__ float sn[8], sd[8];
  for (int off = 16; off; off >>= 1) {
    num += __shfl_xor_sync(0xffffffffu, num, off);
    den += __shfl_xor_sync(0xffffffffu, den, off);
  }
  if ((threadIdx.x & 31) == 0) { sn[threadIdx.x >> 5] = num; sd[threadIdx.x >> 5] = den; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0, b = 0;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) { a += sn[i]; b += sd[i]; }
    if (b != 0.0) {
      atomicAdd(accum, a);
      atomicAdd(accum + 1, b);
    }
  }
}

// backward: dx[n][h][w][c] = sum over the output pixels (oh, ow) whose bilinear footprint contains (h, w) of
// wy * wx * g * weight[t] * (softmax(logits(oh, ow))[c] - [c == t]).  One CTA per input row h: the three input rows
// h-1, h, h+1 are staged in shared memory (every contributing output row blends two of them); stage 1 recomputes the
// softmax per (oh, ow) and accumulates vertically into v[c][ow], stage 2 reduces horizontally (as in
// upsample_logits_bwd_kernel).
template <int MAXC>
__global__ void __launch_bounds__(640) upsample_ce_bwd_kernel(const __nv_bfloat16* __restrict__ x,
                                                               const float* __restrict__ target,
                                                               const float* __restrict__ weight, int C, int Hi, int Wi,
                                                               int cs, int Ho, int Wo, float sh, float sw, int ignore,
                                                               const double* __restrict__ accum, float div,
                                                               const float* __restrict__ gout,
                                                               __nv_bfloat16* __restrict__ dx, int max_rows) {
  // blockDim.x >= Wo (one output column per thread, so the per-class accumulators stay in registers across rows)
  extern __shared__ float smem_f[];
  const int ldi = Wi + 1, ldo = Wo + 1;
  float* xs = smem_f;                 // [3][C][Wi+1]: input rows h-1, h, h+1
  float* v = smem_f;                  // [C][Wo+1]: ALIASES xs (only written after the row loop is done with it)
  const int fl_words = max(4 * C * ldi, C * ldo);
  unsigned char* tg = reinterpret_cast<unsigned char*>(smem_f + fl_words);  // [max_rows][Wo] labels (255 = no gradient)
  const int h = blockIdx.x % Hi, n = blockIdx.x / Hi;
  const float ish = sh > 0.f ? 1.f / sh : 0.f, isw = sw > 0.f ? 1.f / sw : 0.f;
  int oh_lo = sh > 0.f ? (int)floorf((h - 1) * ish) - 1 : 0, oh_hi = sh > 0.f ? (int)ceilf((h + 1) * ish) + 1 : Ho - 1;
  oh_lo = max(oh_lo, 0);
  oh_hi = min(min(oh_hi, Ho - 1), oh_lo + max_rows - 1);
  const int rows = oh_hi - oh_lo + 1;
  for (int i = threadIdx.x; i < rows * Wo; i += blockDim.x) {  // contiguous rows: fully coalesced
    const int t = (int)__ldg(target + ((long long)n * Ho + oh_lo) * Wo + i);
    tg[i] = (t == ignore || t < 0 || t >= C) ? 255 : (unsigned char)t;
  }
  for (int i = threadIdx.x; i < 3 * Wi * C; i += blockDim.x) {
    const int c = i % C, w = (i / C) % Wi, slot = i / (C * Wi);
    const int row = h - 1 + slot;
    xs[(slot * C + c) * ldi + w] =
        (row >= 0 && row < Hi) ? __bfloat162float(x[(((long long)n * Hi + row) * Wi + w) * cs + c]) : 0.f;
  }
  const float g = gout[0] / ((float)accum[1] * div);
  const int ow = threadIdx.x;
  int x0 = 0, x1 = 0; float lx = 0.f;
  if (ow < Wo) bl_coord(ow, sw, Wi, x0, x1, lx);
  constexpr bool KEEP = MAXC <= 24;  // softmax terms in registers next to the accumulators
  float acc[MAXC];
#pragma unroll
  for (int c = 0; c < MAXC; ++c) acc[c] = 0.f;
  __syncthreads();  // xs and tg staged
  // No block-wide step inside the row loop: every thread blends its own two input columns vertically in registers
  // (same order as the forward: vertical into fp32, then horizontal), so the warps of the CTA run the ~7 contributing
  // output rows independently.  (The first version staged the vertically blended row in shared memory behind two
  // __syncthreads per output row; with one 17-warp CTA per SM -- 96 registers x 544 threads -- that serialisation made
  // this kernel 0.64 ms of the 18.4 ms step, profiles/r02_launches_step.md.)
  if (ow < Wo) {
    for (int r = 0; r < rows; ++r) {
      int y0, y1; float ly;
      bl_coord(oh_lo + r, sh, Hi, y0, y1, ly);
      const float wy = (y0 == h ? 1.f - ly : 0.f) + (y1 == h ? ly : 0.f);
      if (wy == 0.f) continue;  // CTA-uniform
      const int t = tg[r * Wo + ow];
      if (t == 255) continue;
      const float* s0 = xs + (y0 - (h - 1)) * C * ldi;
      const float* s1 = xs + (y1 - (h - 1)) * C * ldi;
      const float my = 1.f - ly, mxw = 1.f - lx;
      float lg[KEEP ? MAXC : 1];
      float mx = -INFINITY;
#pragma unroll
      for (int c = 0; c < MAXC; ++c) {
        if (c < C) {
          const float a = my * s0[c * ldi + x0] + ly * s1[c * ldi + x0];
          const float b = my * s0[c * ldi + x1] + ly * s1[c * ldi + x1];
          const float l = mxw * a + lx * b;
          if (KEEP) lg[c] = l;
          mx = fmaxf(mx, l);
        }
      }
      float s = 0.f;
#pragma unroll
      for (int c = 0; c < MAXC; ++c) {
        if (c < C) {
          float l;
          if (KEEP) {
            l = lg[c];
          } else {
            const float a = my * s0[c * ldi + x0] + ly * s1[c * ldi + x0];
            const float b = my * s0[c * ldi + x1] + ly * s1[c * ldi + x1];
            l = mxw * a + lx * b;
          }
          const float e = __expf(l - mx);
          if (KEEP) lg[c] = e;
          s += e;
        }
      }
      const float coef = wy * g * (weight ? weight[t] : 1.f);
      const float inv = coef / s;
#pragma unroll
      for (int c = 0; c < MAXC; ++c) {
        if (c < C) {
          float e;
          if (KEEP) {
            e = lg[c];
          } else {
            const float a = my * s0[c * ldi + x0] + ly * s1[c * ldi + x0];
            const float b = my * s0[c * ldi + x1] + ly * s1[c * ldi + x1];
            e = __expf(mxw * a + lx * b - mx);
          }
          acc[c] += e * inv - (c == t ? coef : 0.f);
        }
      }
    }
  }
  __syncthreads();  // every reader of xs is done: v (same shared memory) may be written
  if (ow < Wo) {
#pragma unroll
    for (int c = 0; c < MAXC; ++c) {
      if (c < C) v[c * ldo + ow] = acc[c];
    }
  }
  __syncthreads();
  __nv_bfloat16* out = dx + ((long long)n * Hi + h) * Wi * cs;
  for (int i = threadIdx.x; i < Wi * cs; i += blockDim.x) {
    const int c = i % cs, w = i / cs;
    float a = 0.f;
    if (c < C) {
      int ow_lo = sw > 0.f ? (int)floorf((w - 1) * isw) - 1 : 0, ow_hi = sw > 0.f ? (int)ceilf((w + 1) * isw) + 1 : Wo - 1;
      ow_lo = max(ow_lo, 0); ow_hi = min(ow_hi, Wo - 1);
      for (int o = ow_lo; o <= ow_hi; ++o) {
        int xa, xb; float l;
        bl_coord(o, sw, Wi, xa, xb, l);
        const float wx = (xa == w ? 1.f - l : 0.f) + (xb == w ? l : 0.f);
        if (wx != 0.f) a += wx * v[c * ldo + o];
      }
    }
    out[(long long)w * cs + c] = __float2bfloat16(a);
  }
}

// ---- x4 fast path of the fused upsample + cross-entropy backward (DeepLab: 129 -> 513, 17 -> 65: Ho = 4(Hi-1)+1).
// With the exact 1/4 scale, output row oh blends input rows y0 = oh/4 and y0+1 with ly = (oh%4)/4 and output column ow
// blends columns x0 = ow/4 and x0+1 with lx = (ow%4)/4.  One CTA per input-row INTERVAL h (output rows 4h .. 4h+3), one
// thread per output column: the softmax of every output pixel is evaluated exactly ONCE (the generic kernel above
// evaluates it from both neighbouring input rows), nothing persists in registers across rows (plain loops over the
// classes, online softmax), the four output columns of an input interval are reduced with a two-step butterfly per class
// and lane j of the quad accumulates into plane j of four per-CTA shared-memory planes -- A0/A1: input row h, columns
// x0 / x0+1; B0/B1: input row h+1 -- each address owned by exactly one thread (no atomics).  Interval h leaves its sums for row h in pa and for row h+1 in pb; the combine
// kernel adds the two partials of every input row in a fixed order: deterministic, no read-modify-write in HBM.
__global__ void __launch_bounds__(544, 2) upsample4_ce_bwd_kernel(const __nv_bfloat16* __restrict__ x,
                                                                  const float* __restrict__ target,
                                                                  const float* __restrict__ weight, int C, int Hi, int Wi,
                                                                  int cs, int Ho, int Wo, int ignore,
                                                                  const double* __restrict__ accum, float div,
                                                                  const float* __restrict__ gout, float* __restrict__ pa,
                                                                  float* __restrict__ pb, int pc) {
  extern __shared__ float smem_f[];
  const int ldi = Wi + 1, plane = C * ldi;
  float* xs0 = smem_f;            // [C][Wi+1] input row h
  float* xs1 = xs0 + plane;       // input row h+1 (row h again for the last interval)
  float* bl = xs1 + plane;        // the two rows blended for the current output row
  float* acc = bl + plane;        // four planes: A0, B0, A1, B1
  const int h = blockIdx.x % Hi, n = blockIdx.x / Hi;
  const int h1 = min(h + 1, Hi - 1);
  for (int i = threadIdx.x; i < Wi * C; i += blockDim.x) {
    const int c = i % C, w = i / C;
    xs0[c * ldi + w] = __bfloat162float(x[(((long long)n * Hi + h) * Wi + w) * cs + c]);
    xs1[c * ldi + w] = __bfloat162float(x[(((long long)n * Hi + h1) * Wi + w) * cs + c]);
  }
  for (int i = threadIdx.x; i < 4 * plane; i += blockDim.x) acc[i] = 0.f;
  const float g = gout[0] / ((float)accum[1] * div);
  const int ow = threadIdx.x;
  const bool col = ow < Wo;
  const int x0 = min(ow >> 2, Wi - 1), x1 = min(x0 + 1, Wi - 1);
  const float lx = col ? 0.25f * (float)(ow & 3) : 0.f;
  // after the butterfly every lane of a quad holds both sums; lane j of the quad owns plane j:
  //   0: A0 (row h, column x0)   1: B0 (row h+1, column x0)   2: A1 (row h, column x0+1)   3: B1 (row h+1, column x0+1)
  const int q = ow & 3;
  const bool mine = (ow & ~3) < Wo && (q < 2 || x0 + 1 < Wi);
  float* my = acc + q * plane + (q < 2 ? x0 : x0 + 1);
  for (int r = 0; r < 4; ++r) {
    const int oh = 4 * h + r;
    if (oh >= Ho) break;  // CTA-uniform (last interval: one row)
    const float ly = 0.25f * (float)r;
    __syncthreads();      // staging done (r = 0) / the previous row's readers of bl are done
    for (int i = threadIdx.x; i < plane; i += blockDim.x) bl[i] = (1.f - ly) * xs0[i] + ly * xs1[i];
    __syncthreads();
    int t = 255;
    if (col) {
      const int tv = (int)__ldg(target + ((long long)n * Ho + oh) * Wo + ow);
      t = (tv == ignore || tv < 0 || tv >= C) ? 255 : tv;
    }
    const float* b0 = bl + x0;
    const float* b1 = bl + x1;
    float mx = -INFINITY, s = 0.f;
    if (t != 255) {  // online softmax: one pass for the maximum and the normaliser
#pragma unroll 4
      for (int c = 0; c < C; ++c) {
        const float l = (1.f - lx) * b0[c * ldi] + lx * b1[c * ldi];
        if (l > mx) {
          s *= __expf(mx - l);
          mx = l;
        }
        s += __expf(l - mx);
      }
    }
    const float coef = t != 255 ? g * (weight ? weight[t] : 1.f) : 0.f;
    const float inv = t != 255 ? coef / s : 0.f;
    const float wq = (q & 1) ? ly : 1.f - ly;  // planes 1 and 3 belong to input row h+1
#pragma unroll 4
    for (int c = 0; c < C; ++c) {
      float d = 0.f;
      if (t != 255) {
        const float l = (1.f - lx) * b0[c * ldi] + lx * b1[c * ldi];
        d = __expf(l - mx) * inv - (c == t ? coef : 0.f);
      }
      float u0 = (1.f - lx) * d, u1 = lx * d;
      u0 += __shfl_xor_sync(0xffffffffu, u0, 1);
      u1 += __shfl_xor_sync(0xffffffffu, u1, 1);
      u0 += __shfl_xor_sync(0xffffffffu, u0, 2);
      u1 += __shfl_xor_sync(0xffffffffu, u1, 2);
      if (mine) my[c * ldi] += wq * (q < 2 ? u0 : u1);
    }
  }
  __syncthreads();
  const float* A0 = acc;
  const float* B0 = acc + plane;
  const float* A1 = acc + 2 * plane;
  const float* B1 = acc + 3 * plane;
  float* oa = pa + ((long long)n * Hi + h) * Wi * pc;
  float* ob = pb + ((long long)n * Hi + h1) * Wi * pc;
  const bool has_b = h + 1 < Hi;
  for (int i = threadIdx.x; i < Wi * C; i += blockDim.x) {
    const int c = i % C, w = i / C;
    oa[(long long)w * pc + c] = A0[c * ldi + w] + A1[c * ldi + w];
    if (has_b) ob[(long long)w * pc + c] = B0[c * ldi + w] + B1[c * ldi + w];
  }
}

// dx[n][h][w][c] = pa[n][h][w][c] + pb[n][h][w][c] (row 0 has no pb), padding channels zero.  One thread per 8
// channels: float4 partial loads (pc % 4 == 0), one 16-byte store.
__global__ void upsample4_ce_combine_kernel(const float* __restrict__ pa, const float* __restrict__ pb,
                                            __nv_bfloat16* __restrict__ dx, int pixels, int Hi, int Wi, int C, int pc,
                                            int cs) {
  const int vpc = cs >> 3;
  const int total = pixels * vpc;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int px = i / vpc, c0 = (i - px * vpc) << 3;
    const bool with_b = (px / Wi) % Hi != 0;
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; j += 4) {
      float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
      if (c0 + j < pc) {  // pc is a multiple of 4: the whole vector is inside the partial row
        a = *reinterpret_cast<const float4*>(pa + (long long)px * pc + c0 + j);
        if (with_b) {
          const float4 b = *reinterpret_cast<const float4*>(pb + (long long)px * pc + c0 + j);
          a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
        }
      }
      v[j] = c0 + j < C ? a.x : 0.f;
      v[j + 1] = c0 + j + 1 < C ? a.y : 0.f;
      v[j + 2] = c0 + j + 2 < C ? a.z : 0.f;
      v[j + 3] = c0 + j + 3 < C ? a.w : 0.f;
    }
    *reinterpret_cast<uint4*>(dx + (long long)px * cs + c0) = pack8m(v);
  }
}

// bf16 shadow of the flat fp32 parameter buffer (one launch per step instead of one pack kernel per layer)
__global__ void cast_f32_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, long long n) {
  const long long n4 = n >> 2;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 v = reinterpret_cast<const float4*>(src)[i];
    uint2 o;
    o.x = pack_bf16x2(v.x, v.y);
    o.y = pack_bf16x2(v.z, v.w);
    reinterpret_cast<uint2*>(dst)[i] = o;
  }
  for (long long i = (n4 << 2) + blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x)
    dst[i] = __float2bfloat16(src[i]);
}

// -------------------------------------------------------------------------------- optimizers
// torch.optim.SGD: d = g*gscale + wd*p; buf = first ? d : mom*buf + d; p -= lr * (nesterov ? d + mom*buf : buf)
__global__ void sgd_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ buf, long long n,
                           float lr, float mom, float wd, int nesterov, int first, float gscale,
                           const float* __restrict__ lr_dev) {
  if (lr_dev) lr = *lr_dev;  // learning rate read at RUN time: a captured CUDA graph follows the LR schedule
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float w = p[i];
    float d = g[i] * gscale + wd * w;
    float b = first ? d : mom * buf[i] + d;
    buf[i] = b;
    p[i] = w - lr * (nesterov ? d + mom * b : b);
  }
}

// torch.optim.Adam (no amsgrad, no weight decay)
__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                            float* __restrict__ v, long long n, float lr, float b1, float b2, float eps, float bc1,
                            float bc2_sqrt, float gscale) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float gi = g[i] * gscale;
    const float mi = b1 * m[i] + (1.f - b1) * gi;
    const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    p[i] -= (lr / bc1) * mi / (sqrtf(vi) / bc2_sqrt + eps);
  }
}

}  // namespace zs3

using namespace zs3;
#define ST(s) static_cast<cudaStream_t>(s)
#define BF(p) static_cast<__nv_bfloat16*>(p)
#define CBF(p) static_cast<const __nv_bfloat16*>(p)

extern "C" int zs3_stem_im2col(const float* x, void* cols, int N, int C, int H, int W, int R, int stride, int pad,
                               int Ho, int Wo, int kpad, int krsc, void* stream) {
  ZS3_CHECK_ARG(x && cols && kpad >= C * R * R && kpad % 8 == 0, "stem_im2col: bad args");
  ZS3_CHECK_ARG(kpad / 2 <= 192, "stem_im2col: kpad=%d too large", kpad);
  const size_t smem = (size_t)C * R * (W + 2 * pad) * sizeof(float);
  ZS3_CHECK_ARG(smem <= 200 * 1024, "stem_im2col: input rows do not fit in shared memory");
  if (smem > 48 * 1024) cudaFuncSetAttribute(stem_im2col_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (krsc) {
    if (smem > 48 * 1024)
      cudaFuncSetAttribute(stem_im2col_krsc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const int threads = (256 / (kpad / 8)) * (kpad / 8);  // whole pixels per iteration
    stem_im2col_krsc_kernel<<<N * Ho, threads, smem, ST(stream)>>>(x, BF(cols), C, H, W, R, stride, pad, Ho, Wo, kpad);
  } else {
    stem_im2col_kernel<<<N * Ho, 192, smem, ST(stream)>>>(x, BF(cols), C, H, W, R, stride, pad, Ho, Wo, kpad, krsc);
  }
  ZS3_CHECK_LAUNCH("stem_im2col");
  return ZS3_OK;
}

extern "C" int zs3_maxpool_fwd(const void* x, void* y, unsigned char* argmax, int N, int H, int W, int C, int Ho,
                               int Wo, int k, int stride, int pad, void* stream) {
  ZS3_CHECK_ARG(x && y && argmax && C % 8 == 0 && k * k <= 255, "maxpool_fwd: bad args");
  const long long items = (long long)N * Ho * Wo * (C / 8);
  if (items * 2 < (1ll << 31) && (long long)N * H * W * C < (1ll << 31))
    maxpool_fwd_kernel<int><<<ew_blocks(items, 256), 256, 0, ST(stream)>>>(CBF(x), BF(y), argmax, N, H, W, C, Ho, Wo, k,
                                                                         stride, pad);
  else
    maxpool_fwd_kernel<long long><<<ew_blocks(items, 256), 256, 0, ST(stream)>>>(CBF(x), BF(y), argmax, N, H, W, C, Ho, Wo,
                                                                               k, stride, pad);
  ZS3_CHECK_LAUNCH("maxpool_fwd");
  return ZS3_OK;
}

extern "C" int zs3_maxpool_bwd(const void* dy, const unsigned char* argmax, void* dx, int N, int H, int W, int C,
                               int Ho, int Wo, int k, int stride, int pad, void* stream) {
  ZS3_CHECK_ARG(dy && dx && argmax && C % 8 == 0, "maxpool_bwd: bad args");
  ZS3_CHECK_ARG(k >= 1 && k <= 8 && stride >= 1, "maxpool_bwd: kernel size %d outside [1, 8]", k);
  // output rows whose windows can cover one input row: ceil(k / stride), staged in shared memory (gradient + slots)
  const int max_rows = (k + stride - 1) / stride < 8 ? (k + stride - 1) / stride : 8;
  const size_t smem = (size_t)max_rows * Wo * C * 3;
  ZS3_CHECK_ARG(smem <= 200 * 1024, "maxpool_bwd: %d output rows of %d x %d do not fit in shared memory", max_rows, Wo, C);
  if (smem > 48 * 1024)
    cudaFuncSetAttribute(maxpool_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  maxpool_bwd_kernel<<<N * H, 256, smem, ST(stream)>>>(CBF(dy), argmax, BF(dx), N, H, W, C, Ho, Wo, k, stride, pad, max_rows);
  ZS3_CHECK_LAUNCH("maxpool_bwd");
  return ZS3_OK;
}

static float bl_scale(int in, int out) { return out > 1 ? (float)(in - 1) / (float)(out - 1) : 0.f; }

extern "C" int zs3_bilinear_fwd(const void* x, void* y, int N, int Hi, int Wi, int Ho, int Wo, int C, int x_cs,
                                int y_cs, void* stream) {
  ZS3_CHECK_ARG(x && y && C % 8 == 0 && x_cs % 8 == 0 && y_cs % 8 == 0 && x_cs >= C && y_cs >= C, "bilinear_fwd: bad args");
  const long long items = (long long)N * Ho * Wo * (C / 8);
  if (items * 2 < (1ll << 31) && (long long)N * Ho * Wo * y_cs < (1ll << 31))
    bilinear_fwd_kernel<int><<<ew_blocks(items, 256), 256, 0, ST(stream)>>>(CBF(x), BF(y), N, Hi, Wi, Ho, Wo, C, x_cs, y_cs,
                                                                          bl_scale(Hi, Ho), bl_scale(Wi, Wo));
  else
    bilinear_fwd_kernel<long long><<<ew_blocks(items, 256), 256, 0, ST(stream)>>>(CBF(x), BF(y), N, Hi, Wi, Ho, Wo, C, x_cs,
                                                                                y_cs, bl_scale(Hi, Ho), bl_scale(Wi, Wo));
  ZS3_CHECK_LAUNCH("bilinear_fwd");
  return ZS3_OK;
}

extern "C" int zs3_bilinear_bwd(const void* dy, void* dx, int N, int Hi, int Wi, int Ho, int Wo, int C, int dy_cs,
                                int dx_cs, int accumulate, void* stream) {
  ZS3_CHECK_ARG(dy && dx && C % 8 == 0 && dy_cs % 8 == 0 && dx_cs % 8 == 0, "bilinear_bwd: bad args");
  const long long items = (long long)N * Hi * Wi * (C / 8);
  if (items * 2 < (1ll << 31) && (long long)N * Hi * Wi * dx_cs < (1ll << 31))
    bilinear_bwd_kernel<int><<<ew_blocks(items, 256), 256, 0, ST(stream)>>>(
        CBF(dy), BF(dx), N, Hi, Wi, Ho, Wo, C, dy_cs, dx_cs, bl_scale(Hi, Ho), bl_scale(Wi, Wo), accumulate);
  else
    bilinear_bwd_kernel<long long><<<ew_blocks(items, 256), 256, 0, ST(stream)>>>(
        CBF(dy), BF(dx), N, Hi, Wi, Ho, Wo, C, dy_cs, dx_cs, bl_scale(Hi, Ho), bl_scale(Wi, Wo), accumulate);
  ZS3_CHECK_LAUNCH("bilinear_bwd");
  return ZS3_OK;
}

extern "C" int zs3_upsample_logits_fwd(const void* x, float* y, int N, int C, int Hi, int Wi, int cs, int Ho, int Wo,
                                       void* stream) {
  ZS3_CHECK_ARG(x && y && C <= cs, "upsample_logits_fwd: bad args");
  const size_t smem = (size_t)C * (Wi + 1) * sizeof(float);
  ZS3_CHECK_ARG(smem <= 200 * 1024, "upsample_logits_fwd: C*Wi too large for shared memory");
  if (smem > 48 * 1024) cudaFuncSetAttribute(upsample_logits_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  upsample_logits_fwd_kernel<<<N * Ho, 256, smem, ST(stream)>>>(CBF(x), y, C, Hi, Wi, cs, Ho, Wo, bl_scale(Hi, Ho),
                                                                bl_scale(Wi, Wo));
  ZS3_CHECK_LAUNCH("upsample_logits_fwd");
  return ZS3_OK;
}

extern "C" int zs3_upsample_logits_bwd(const float* dy, void* dx, int N, int C, int Hi, int Wi, int cs, int Ho, int Wo,
                                       void* stream) {
  ZS3_CHECK_ARG(dy && dx && C <= cs, "upsample_logits_bwd: bad args");
  const size_t smem = (size_t)C * (Wo + 1) * sizeof(float);
  ZS3_CHECK_ARG(smem <= 200 * 1024, "upsample_logits_bwd: C*Wo too large for shared memory");
  if (smem > 48 * 1024) cudaFuncSetAttribute(upsample_logits_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  upsample_logits_bwd_kernel<<<N * Hi, 256, smem, ST(stream)>>>(dy, BF(dx), C, Hi, Wi, cs, Ho, Wo, bl_scale(Hi, Ho),
                                                                bl_scale(Wi, Wo));
  ZS3_CHECK_LAUNCH("upsample_logits_bwd");
  return ZS3_OK;
}

extern "C" int zs3_spatial_sum(const void* x, void* y, int N, int HW, int C, int x_cs, int y_cs, float scale,
                               void* stream) {
  ZS3_CHECK_ARG(x && y && C % 64 == 0 && x_cs >= C && y_cs >= C, "spatial_sum: C must be a multiple of 64");
  spatial_sum_kernel<<<dim3(C / 64, N), 256, 0, ST(stream)>>>(CBF(x), BF(y), HW, x_cs, y_cs, scale);
  ZS3_CHECK_LAUNCH("spatial_sum");
  return ZS3_OK;
}

extern "C" int zs3_spatial_broadcast(const void* x, void* y, int N, int HW, int C, int x_cs, int y_cs, float scale,
                                     int accumulate, void* stream) {
  ZS3_CHECK_ARG(x && y && C % 8 == 0 && x_cs % 8 == 0 && y_cs % 8 == 0, "spatial_broadcast: bad args");
  spatial_broadcast_kernel<<<ew_blocks((long long)N * HW * (C / 8), 256), 256, 0, ST(stream)>>>(
      CBF(x), BF(y), N, HW, C, x_cs, y_cs, scale, accumulate);
  ZS3_CHECK_LAUNCH("spatial_broadcast");
  return ZS3_OK;
}

extern "C" int zs3_ce_fwd(const float* logit, const float* target, const float* weight, int N, int C, long long HW,
                          int ignore_index, float div, double* accum2, float* loss, void* stream) {
  ZS3_CHECK_ARG(logit && target && accum2 && loss && C > 0, "ce_fwd: bad args");
  cudaMemsetAsync(accum2, 0, 2 * sizeof(double), ST(stream));
  const long long total = (long long)N * HW;
  ce_fwd_kernel<<<ew_blocks(total, 256, 148 * 8), 256, 0, ST(stream)>>>(logit, target, weight, C, HW, total,
                                                                        ignore_index, accum2);
  ce_finalize_kernel<<<1, 1, 0, ST(stream)>>>(accum2, div, loss);
  ZS3_CHECK_LAUNCH("ce_fwd");
  return ZS3_OK;
}

extern "C" int zs3_ce_bwd(const float* logit, const float* target, const float* weight, int N, int C, long long HW,
                          int ignore_index, float div, const double* accum2, const float* grad_out, float* dlogit,
                          void* stream) {
  ZS3_CHECK_ARG(logit && target && accum2 && grad_out && dlogit, "ce_bwd: bad args");
  const long long total = (long long)N * HW;
  ce_bwd_kernel<<<ew_blocks(total, 256, 148 * 8), 256, 0, ST(stream)>>>(logit, target, weight, C, HW, total,
                                                                        ignore_index, accum2, div, grad_out, dlogit);
  ZS3_CHECK_LAUNCH("ce_bwd");
  return ZS3_OK;
}

extern "C" int zs3_upsample_ce_fwd(const void* x, const float* target, const float* weight, int N, int C, int Hi, int Wi,
                                   int cs, int Ho, int Wo, int ignore_index, float div, double* accum2, float* loss,
                                   void* stream) {
  ZS3_CHECK_ARG(x && target && accum2 && loss && C > 0 && C <= cs && C <= CE_MAX_C, "upsample_ce_fwd: bad args (C <= %d)",
                CE_MAX_C);
  const size_t smem = (size_t)C * (Wi + 1) * sizeof(float);
  ZS3_CHECK_ARG(smem <= 200 * 1024, "upsample_ce_fwd: C*Wi too large for shared memory");
  if (smem > 48 * 1024) cudaFuncSetAttribute(upsample_ce_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaMemsetAsync(accum2, 0, 2 * sizeof(double), ST(stream));
  upsample_ce_fwd_kernel<<<N * Ho, 256, smem, ST(stream)>>>(CBF(x), target, weight, C, Hi, Wi, cs, Ho, Wo, bl_scale(Hi, Ho),
                                                            bl_scale(Wi, Wo), ignore_index, accum2);
  ce_finalize_kernel<<<1, 1, 0, ST(stream)>>>(accum2, div, loss);
  ZS3_CHECK_LAUNCH("upsample_ce_fwd");
  return ZS3_OK;
}

extern "C" int zs3_upsample_ce_bwd(const void* x, const float* target, const float* weight, int N, int C, int Hi, int Wi,
                                   int cs, int Ho, int Wo, int ignore_index, float div, const double* accum2,
                                   const float* grad_out, void* dx, void* stream) {
  ZS3_CHECK_ARG(x && target && accum2 && grad_out && dx && C > 0 && C <= cs && C <= CE_MAX_C,
                "upsample_ce_bwd: bad args (C <= %d)", CE_MAX_C);
  // output rows whose bilinear footprint can touch one input row: (h-1)/sh - 1 .. (h+1)/sh + 1
  const float shf = bl_scale(Hi, Ho);
  const int max_rows = shf > 0.f ? (int)ceilf(2.f / shf) + 6 : Ho;
  // staging rows (xs, bl) and the vertical accumulators v share one region (the kernel aliases them)
  const size_t fl_words = (size_t)4 * C * (Wi + 1) > (size_t)C * (Wo + 1) ? (size_t)4 * C * (Wi + 1) : (size_t)C * (Wo + 1);
  const size_t smem = fl_words * sizeof(float) + (size_t)max_rows * Wo;
  ZS3_CHECK_ARG(smem <= 220 * 1024, "upsample_ce_bwd: rows do not fit in shared memory");
  ZS3_CHECK_ARG(Wo <= 640, "upsample_ce_bwd: output width %d > 640 (use the unfused upsample + CE kernels)", Wo);
  const int threads = ((Wo + 31) / 32) * 32;  // one output column per thread
  if (C <= 24) {
    if (smem > 48 * 1024)
      cudaFuncSetAttribute(upsample_ce_bwd_kernel<24>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    upsample_ce_bwd_kernel<24><<<N * Hi, threads, smem, ST(stream)>>>(CBF(x), target, weight, C, Hi, Wi, cs, Ho, Wo, shf,
                                                                      bl_scale(Wi, Wo), ignore_index, accum2, div,
                                                                      grad_out, BF(dx), max_rows);
  } else {
    if (smem > 48 * 1024)
      cudaFuncSetAttribute(upsample_ce_bwd_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    upsample_ce_bwd_kernel<64><<<N * Hi, threads, smem, ST(stream)>>>(CBF(x), target, weight, C, Hi, Wi, cs, Ho, Wo, shf,
                                                                      bl_scale(Wi, Wo), ignore_index, accum2, div,
                                                                      grad_out, BF(dx), max_rows);
  }
  ZS3_CHECK_LAUNCH("upsample_ce_bwd");
  return ZS3_OK;
}

extern "C" unsigned long long zs3_upsample4_ce_bwd_workspace_size(int N, int C, int Hi, int Wi) {
  if (N <= 0 || C <= 0 || Hi <= 0 || Wi <= 0) return 0;
  const unsigned long long pc = (unsigned long long)((C + 3) / 4 * 4);
  return 2ull * N * Hi * Wi * pc * sizeof(float);
}

/* x4 fast path of zs3_upsample_ce_bwd (Ho == 4*(Hi-1)+1, Wo == 4*(Wi-1)+1, Wo <= 544); returns ZS3_ERR_UNSUPPORTED for
 * other geometries (the caller then uses zs3_upsample_ce_bwd).  workspace: zs3_upsample4_ce_bwd_workspace_size bytes. */
extern "C" int zs3_upsample4_ce_bwd(const void* x, const float* target, const float* weight, int N, int C, int Hi, int Wi,
                                    int cs, int Ho, int Wo, int ignore_index, float div, const double* accum2,
                                    const float* grad_out, void* dx, void* workspace, unsigned long long workspace_bytes,
                                    void* stream) {
  ZS3_CHECK_ARG(x && target && accum2 && grad_out && dx && workspace && C > 0 && C <= cs && C <= CE_MAX_C,
                "upsample4_ce_bwd: bad args (C <= %d)", CE_MAX_C);
  if (Ho != 4 * (Hi - 1) + 1 || Wo != 4 * (Wi - 1) + 1 || Wo > 544 || Hi < 2 || Wi < 2) {
    zs3::set_error("upsample4_ce_bwd: geometry %dx%d -> %dx%d is not an exact x4 upsample", Hi, Wi, Ho, Wo);
    return ZS3_ERR_UNSUPPORTED;
  }
  ZS3_CHECK_ARG(workspace_bytes >= zs3_upsample4_ce_bwd_workspace_size(N, C, Hi, Wi) &&
                    (reinterpret_cast<uintptr_t>(workspace) & 15) == 0,
                "upsample4_ce_bwd: workspace too small or misaligned");
  const int pc = (C + 3) / 4 * 4;
  const size_t smem = (size_t)7 * C * (Wi + 1) * sizeof(float);
  ZS3_CHECK_ARG(smem <= 224 * 1024, "upsample4_ce_bwd: %d classes x %d columns do not fit in shared memory", C, Wi);
  float* pa = static_cast<float*>(workspace);
  float* pb = pa + (size_t)N * Hi * Wi * pc;
  const int threads = ((Wo + 31) / 32) * 32;
  if (smem > 48 * 1024)
    cudaFuncSetAttribute(upsample4_ce_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  upsample4_ce_bwd_kernel<<<N * Hi, threads, smem, ST(stream)>>>(CBF(x), target, weight, C, Hi, Wi, cs, Ho, Wo, ignore_index,
                                                                 accum2, div, grad_out, pa, pb, pc);
  ZS3_CHECK_LAUNCH("upsample4_ce_bwd");
  const long long pixels = (long long)N * Hi * Wi;
  ZS3_CHECK_ARG(cs % 8 == 0 && pixels * (cs / 8) < (1ll << 31), "upsample4_ce_bwd: bad channel stride / too many pixels");
  upsample4_ce_combine_kernel<<<ew_blocks(pixels * (cs / 8), 256), 256, 0, ST(stream)>>>(pa, pb, BF(dx), (int)pixels, Hi, Wi,
                                                                                          C, pc, cs);
  ZS3_CHECK_LAUNCH("upsample4_ce_bwd(combine)");
  return ZS3_OK;
}

extern "C" int zs3_cast_f32_to_bf16(const float* src, void* dst, long long n, void* stream) {
  ZS3_CHECK_ARG(src && dst && n >= 0 && (reinterpret_cast<uintptr_t>(src) & 15) == 0 &&
                    (reinterpret_cast<uintptr_t>(dst) & 7) == 0,
                "cast_f32_to_bf16: bad args");
  if (n == 0) return ZS3_OK;
  cast_f32_bf16_kernel<<<ew_blocks(n / 4 + 1, 256), 256, 0, ST(stream)>>>(src, BF(dst), n);
  ZS3_CHECK_LAUNCH("cast_f32_to_bf16");
  return ZS3_OK;
}

extern "C" int zs3_sgd_step(float* p, const float* g, float* buf, long long n, float lr, float momentum,
                            float weight_decay, int nesterov, int first_step, float grad_scale, void* stream) {
  ZS3_CHECK_ARG(p && g && buf && n >= 0, "sgd_step: bad args");
  if (n == 0) return ZS3_OK;
  sgd_kernel<<<ew_blocks(n, 256), 256, 0, ST(stream)>>>(p, g, buf, n, lr, momentum, weight_decay, nesterov, first_step,
                                                        grad_scale, nullptr);
  ZS3_CHECK_LAUNCH("sgd_step");
  return ZS3_OK;
}

extern "C" int zs3_sgd_step_lrdev(float* p, const float* g, float* buf, long long n, const float* lr_dev,
                                  float momentum, float weight_decay, int nesterov, int first_step, float grad_scale,
                                  void* stream) {
  ZS3_CHECK_ARG(p && g && buf && lr_dev && n >= 0, "sgd_step_lrdev: bad args");
  if (n == 0) return ZS3_OK;
  sgd_kernel<<<ew_blocks(n, 256), 256, 0, ST(stream)>>>(p, g, buf, n, 0.f, momentum, weight_decay, nesterov,
                                                        first_step, grad_scale, lr_dev);
  ZS3_CHECK_LAUNCH("sgd_step_lrdev");
  return ZS3_OK;
}

extern "C" int zs3_adam_step(float* p, const float* g, float* m, float* v, long long n, float lr, float beta1,
                             float beta2, float eps, int step, float grad_scale, void* stream) {
  ZS3_CHECK_ARG(p && g && m && v && n >= 0 && step >= 1, "adam_step: bad args");
  if (n == 0) return ZS3_OK;
  const float bc1 = 1.f - powf(beta1, (float)step);
  const float bc2s = sqrtf(1.f - powf(beta2, (float)step));
  adam_kernel<<<ew_blocks(n, 256), 256, 0, ST(stream)>>>(p, g, m, v, n, lr, beta1, beta2, eps, bc1, bc2s, grad_scale);
  ZS3_CHECK_LAUNCH("adam_step");
  return ZS3_OK;
}
