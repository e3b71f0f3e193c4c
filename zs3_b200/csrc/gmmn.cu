// GMMN generator + multi-bandwidth MMD loss kernels (fp32 math: the MMD exponent cancels catastrophically
// in reduced precision; the whole generator is < 0.2 GFLOP per trained class).
//
//   zs3_sgemm            generic fp32 tiled GEMM with optional row gathers and a device-side dynamic extent
//                        (nn.Linear forward / backward of zs3/modeling/gmmn.py:17-21,28-34 and the dense
//                        adj @ (x @ W) of pygcn's GraphConvolution, gmmn.py:55-67)
//   zs3_find_active_rows compacts the rows of a gradient that are non-zero (only the 128 sampled rows of
//                        zs3/train_pascal_GMMN.py:229-237 carry gradient; everything else is skipped)
//   zs3_leaky_dropout_*  LeakyReLU(0.2) + Dropout(0.5) (gmmn.py:19-20)
//   zs3_mmd_*            GMMNLoss.moment_loss (zs3/utils/loss.py:99-115) forward and analytic backward
#include "common.cuh"
#include "ptx.cuh"

namespace zs3 {

constexpr int GT = 64;   // tile rows / cols
constexpr int GK = 16;   // k chunk
constexpr int GLD = 68;  // padded leading dimension of the shared tiles

struct GemmP {
  const float* A; long long lda; int transA; const int* idxA;
  const float* B; long long ldb; int transB; const int* idxB;
  float* C; long long ldc;
  int M, N, K;
  const float* bias;  // [N] or null
  int accumulate;
  const int* dyn_count; int dyn_dim;  // 1: M = *dyn_count, 2: K = *dyn_count
};

__global__ void __launch_bounds__(256) sgemm_kernel(const GemmP p) {
  __shared__ __align__(16) float As[GK][GLD];
  __shared__ __align__(16) float Bs[GK][GLD];
  int M = p.M, K = p.K;
  if (p.dyn_dim == 1) M = min(M, *p.dyn_count);
  if (p.dyn_dim == 2) K = min(K, *p.dyn_count);
  const int i0 = blockIdx.y * GT, j0 = blockIdx.x * GT;
  if (i0 >= M) return;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  float acc[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;

  for (int k0 = 0; k0 < K; k0 += GK) {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int l = threadIdx.x + e * 256;
      {  // A tile: op(A)(i, k)
        int i, k;
        if (p.transA) { i = l & 63; k = l >> 6; } else { k = l & 15; i = l >> 4; }
        float v = 0.f;
        if (i0 + i < M && k0 + k < K) {
          long long row = p.transA ? (k0 + k) : (i0 + i);
          const long long col = p.transA ? (i0 + i) : (k0 + k);
          if (p.idxA) row = p.idxA[row];
          v = __ldg(p.A + row * p.lda + col);
        }
        As[k][i] = v;
      }
      {  // B tile: op(B)(k, j)
        int j, k;
        if (p.transB) { k = l & 15; j = l >> 4; } else { j = l & 63; k = l >> 6; }
        float v = 0.f;
        if (j0 + j < p.N && k0 + k < K) {
          long long row = p.transB ? (j0 + j) : (k0 + k);
          const long long col = p.transB ? (k0 + k) : (j0 + j);
          if (p.idxB) row = p.idxB[row];
          v = __ldg(p.B + row * p.ldb + col);
        }
        Bs[k][j] = v;
      }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < GK; ++k) {
      const float4 a = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int x = 0; x < 4; ++x)
#pragma unroll
        for (int y = 0; y < 4; ++y) acc[x][y] = fmaf(av[x], bv[y], acc[x][y]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int x = 0; x < 4; ++x) {
    const int i = i0 + ty * 4 + x;
    if (i >= M) continue;
#pragma unroll
    for (int y = 0; y < 4; ++y) {
      const int j = j0 + tx * 4 + y;
      if (j >= p.N) continue;
      float v = acc[x][y] + (p.bias ? p.bias[j] : 0.f);
      float* c = p.C + (long long)i * p.ldc + j;
      *c = p.accumulate ? *c + v : v;
    }
  }
}

// rows[0..count) = indices of rows of g[n][f] with any non-zero entry (order not deterministic)
__global__ void find_active_rows_kernel(const float* __restrict__ g, int n, int f, int* __restrict__ rows,
                                        int* __restrict__ count) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= n) return;
  bool nz = false;
  for (int j = lane; j < f; j += 32) nz |= (g[(long long)warp * f + j] != 0.f);
  nz = __any_sync(0xffffffffu, nz);
  if (nz && lane == 0) rows[atomicAdd(count, 1)] = warp;
}

// out[j] (+)= sum_r A[idx[r]][j], r < *count (or n if count == null)
__global__ void col_sum_kernel(const float* __restrict__ A, long long lda, const int* idx, const int* count, int n,
                               int ncols, float* __restrict__ out, int accumulate) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= ncols) return;
  const int R = count ? min(n, *count) : n;
  float s = 0.f;
  for (int r = 0; r < R; ++r) s += A[(long long)(idx ? idx[r] : r) * lda + j];
  out[j] = accumulate ? out[j] + s : s;
}

__device__ __forceinline__ uint64_t splitmix64g(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}

// y = dropout(leaky_relu(x, slope), p); drop_mode 0 none, 1 counter RNG, 2 explicit keep mask (bytes)
__global__ void leaky_dropout_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, long long n, float slope,
                                         int drop_mode, float p, uint64_t seed, uint64_t offset,
                                         const unsigned char* __restrict__ mask) {
  const float ks = 1.f / (1.f - p);
  const uint32_t thresh = (uint32_t)(p * 65536.0f + 0.5f);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float v = x[i];
    v = v > 0.f ? v : v * slope;
    if (drop_mode == 1) {
      const uint64_t h = splitmix64g(seed ^ splitmix64g(offset + (uint64_t)(i >> 2)));
      const uint32_t u = (uint32_t)(h >> (16 * (i & 3))) & 0xFFFF;
      v = u >= thresh ? v * ks : 0.f;
    } else if (drop_mode == 2) {
      v = mask[i] ? v * ks : 0.f;
    }
    y[i] = v;
  }
}

// dx = dy * d/dx[dropout(leaky(x))] reconstructed from the forward OUTPUT h: h>0 -> ks, h<0 -> slope*ks, h==0 -> 0.
// rows of h are gathered through idx when given (dy is compact [R][cols], R = *count)
__global__ void leaky_dropout_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ h,
                                         float* __restrict__ dx, int rows, int cols, const int* idx, const int* count,
                                         float slope, float p) {
  const float ks = 1.f / (1.f - p);
  const int R = count ? min(rows, *count) : rows;
  const long long total = (long long)R * cols;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(i / cols), c = (int)(i - (long long)r * cols);
    const float hv = h[(long long)(idx ? idx[r] : r) * cols + c];
    dx[i] = dy[i] * (hv > 0.f ? ks : (hv < 0.f ? slope * ks : 0.f));
  }
}

// ------------------------------------------------------------------------------------ MMD
// X = [gen (M rows); real (N rows)], L = M + N, D columns.  s_i = +1/N for i < N... the reference's
// get_scale_matrix quirk: the FIRST N rows get +1/N and the last M rows -1/M (zs3/utils/loss.py:92-97).
// P[i][j] = s_i s_j sum_sigma exp(e_ij/sigma)/sigma,  loss2 += sum_ij s_i s_j sum_sigma exp(e_ij/sigma),
// e_ij = x_i.x_j - |x_i|^2/2 - |x_j|^2/2 = -|x_i - x_j|^2 / 2.
struct MmdP {
  const float* gen; const float* real; int M, N, D;
  float sigma[8]; int nsigma;
  float* P; double* loss2;
};

__device__ __forceinline__ const float* mmd_row(const MmdP& p, int i) {
  return i < p.M ? p.gen + (long long)i * p.D : p.real + (long long)(i - p.M) * p.D;
}

__global__ void __launch_bounds__(256) mmd_pairs_kernel(const MmdP p) {
  __shared__ float Xi[32][33], Xj[32][33];
  __shared__ float red[8];
  const int L = p.M + p.N;
  const int i0 = blockIdx.y * 32, j0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // ty in 0..7: rows ty, ty+8, ty+16, ty+24
  float d2[4] = {0.f, 0.f, 0.f, 0.f};
  for (int d0 = 0; d0 < p.D; d0 += 32) {
    for (int e = ty; e < 32; e += 8) {
      const int i = i0 + e, j = j0 + e, d = d0 + tx;
      Xi[e][tx] = (i < L && d < p.D) ? mmd_row(p, i)[d] : 0.f;
      Xj[e][tx] = (j < L && d < p.D) ? mmd_row(p, j)[d] : 0.f;
    }
    __syncthreads();
#pragma unroll 8
    for (int d = 0; d < 32; ++d) {
      const float xj = Xj[tx][d];
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        const float df = Xi[ty + 8 * a][d] - xj;
        d2[a] = fmaf(df, df, d2[a]);
      }
    }
    __syncthreads();
  }
  float part = 0.f;
  const int j = j0 + tx;
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    const int i = i0 + ty + 8 * a;
    if (i < L && j < L) {
      const float si = i < p.N ? 1.f / p.N : -1.f / p.M;
      const float sj = j < p.N ? 1.f / p.N : -1.f / p.M;
      const float e = -0.5f * d2[a];
      float kv = 0.f, kp = 0.f;
      for (int s = 0; s < p.nsigma; ++s) {
        const float ex = expf(e / p.sigma[s]);
        kv += ex;
        kp += ex / p.sigma[s];
      }
      part += si * sj * kv;
      p.P[(long long)i * L + j] = si * sj * kp;
    }
  }
  for (int off = 16; off; off >>= 1) part += __shfl_xor_sync(0xffffffffu, part, off);
  if (tx == 0) red[ty] = part;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0;
    for (int a = 0; a < 8; ++a) s += red[a];
    atomicAdd(p.loss2, s);
  }
}

__global__ void mmd_finalize_kernel(const double* loss2, float* loss) { *loss = sqrtf((float)*loss2); }

// grad_i = gout / loss * sum_j P_ij (x_j - x_i)   for rows [row_begin, row_begin + rows) of X
__global__ void mmd_grad_kernel(const MmdP p, const float* __restrict__ loss, const float* __restrict__ gout,
                                int row_begin, int rows, float* __restrict__ grad) {
  const int L = p.M + p.N;
  const int d = blockIdx.x * blockDim.x + threadIdx.x;
  const int i = row_begin + blockIdx.y;
  if (d >= p.D || blockIdx.y >= rows) return;
  const float xi = mmd_row(p, i)[d];
  const float* Pi = p.P + (long long)i * L;
  float acc = 0.f;
  for (int j = 0; j < L; ++j) acc = fmaf(Pi[j], mmd_row(p, j)[d] - xi, acc);
  grad[(long long)blockIdx.y * p.D + d] = gout[0] / loss[0] * acc;
}

// y[r][0:c1] = a[r], y[r][c1:c1+c2] = b[r]   (torch.cat((embd, noise), 1), gmmn.py:44)
__global__ void concat2_kernel(const float* __restrict__ a, int c1, const float* __restrict__ b, int c2,
                               float* __restrict__ y, long long rows) {
  const int c = c1 + c2;
  const long long total = rows * c;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / c;
    const int j = (int)(i - r * c);
    y[i] = j < c1 ? a[r * c1 + j] : b[r * c2 + (j - c1)];
  }
}

}  // namespace zs3

using namespace zs3;
#define ST(s) static_cast<cudaStream_t>(s)

extern "C" int zs3_sgemm(const zs3_sgemm_args* a, void* stream) {
  ZS3_CHECK_ARG(a && a->A && a->B && a->C && a->M >= 0 && a->N > 0 && a->K >= 0, "sgemm: bad args");
  ZS3_CHECK_ARG(a->dyn_dim == 0 || a->dyn_count != nullptr, "sgemm: dyn_dim needs dyn_count");
  if (a->M == 0) return ZS3_OK;
  GemmP p;
  p.A = a->A; p.lda = a->lda; p.transA = a->transA; p.idxA = a->idxA;
  p.B = a->B; p.ldb = a->ldb; p.transB = a->transB; p.idxB = a->idxB;
  p.C = a->C; p.ldc = a->ldc; p.M = a->M; p.N = a->N; p.K = a->K;
  p.bias = a->bias; p.accumulate = a->accumulate; p.dyn_count = a->dyn_count; p.dyn_dim = a->dyn_dim;
  dim3 grid((a->N + GT - 1) / GT, (a->M + GT - 1) / GT);
  sgemm_kernel<<<grid, 256, 0, ST(stream)>>>(p);
  ZS3_CHECK_LAUNCH("sgemm");
  return ZS3_OK;
}

extern "C" int zs3_find_active_rows(const float* g, int n, int f, int* rows, int* count, void* stream) {
  ZS3_CHECK_ARG(g && rows && count && n >= 0 && f > 0, "find_active_rows: bad args");
  cudaMemsetAsync(count, 0, sizeof(int), ST(stream));
  if (n == 0) return ZS3_OK;
  find_active_rows_kernel<<<(n * 32 + 255) / 256, 256, 0, ST(stream)>>>(g, n, f, rows, count);
  ZS3_CHECK_LAUNCH("find_active_rows");
  return ZS3_OK;
}

extern "C" int zs3_col_sum(const float* A, long long lda, const int* idx, const int* count, int n, int ncols,
                           float* out, int accumulate, void* stream) {
  ZS3_CHECK_ARG(A && out && ncols > 0, "col_sum: bad args");
  col_sum_kernel<<<(ncols + 127) / 128, 128, 0, ST(stream)>>>(A, lda, idx, count, n, ncols, out, accumulate);
  ZS3_CHECK_LAUNCH("col_sum");
  return ZS3_OK;
}

extern "C" int zs3_leaky_dropout_fwd(const float* x, float* y, long long n, float slope, int drop_mode, float p,
                                     unsigned long long seed, unsigned long long offset, const unsigned char* mask,
                                     void* stream) {
  ZS3_CHECK_ARG(x && y && p >= 0.f && p < 1.f && (drop_mode != 2 || mask), "leaky_dropout_fwd: bad args");
  if (n == 0) return ZS3_OK;
  long long b = (n + 255) / 256;
  if (b > 148 * 8) b = 148 * 8;
  leaky_dropout_fwd_kernel<<<(int)b, 256, 0, ST(stream)>>>(x, y, n, slope, p > 0.f ? drop_mode : 0, p, seed, offset,
                                                            mask);
  ZS3_CHECK_LAUNCH("leaky_dropout_fwd");
  return ZS3_OK;
}

extern "C" int zs3_leaky_dropout_bwd(const float* dy, const float* h, float* dx, int rows, int cols, const int* idx,
                                     const int* count, float slope, float p, void* stream) {
  ZS3_CHECK_ARG(dy && h && dx && cols > 0, "leaky_dropout_bwd: bad args");
  if (rows == 0) return ZS3_OK;
  long long b = ((long long)rows * cols + 255) / 256;
  if (b > 148 * 8) b = 148 * 8;
  leaky_dropout_bwd_kernel<<<(int)b, 256, 0, ST(stream)>>>(dy, h, dx, rows, cols, idx, count, slope, p);
  ZS3_CHECK_LAUNCH("leaky_dropout_bwd");
  return ZS3_OK;
}

static int fill_mmd(MmdP& p, const float* gen, const float* real, int M, int N, int D, const float* sigma, int nsigma,
                    float* P, double* loss2) {
  ZS3_CHECK_ARG(gen && real && P && loss2 && M > 0 && N > 0 && D > 0 && nsigma > 0 && nsigma <= 8, "mmd: bad args");
  p.gen = gen; p.real = real; p.M = M; p.N = N; p.D = D; p.nsigma = nsigma; p.P = P; p.loss2 = loss2;
  for (int i = 0; i < nsigma; ++i) p.sigma[i] = sigma[i];
  return ZS3_OK;
}

// sigma is a HOST array
extern "C" int zs3_mmd_fwd(const float* gen, const float* real, int M, int N, int D, const float* sigma, int nsigma,
                           float* P, double* loss2, float* loss, void* stream) {
  MmdP p;
  int rc = fill_mmd(p, gen, real, M, N, D, sigma, nsigma, P, loss2);
  if (rc) return rc;
  ZS3_CHECK_ARG(loss != nullptr, "mmd_fwd: null loss");
  cudaMemsetAsync(loss2, 0, sizeof(double), ST(stream));
  const int L = M + N;
  dim3 grid((L + 31) / 32, (L + 31) / 32);
  mmd_pairs_kernel<<<grid, 256, 0, ST(stream)>>>(p);
  mmd_finalize_kernel<<<1, 1, 0, ST(stream)>>>(loss2, loss);
  ZS3_CHECK_LAUNCH("mmd_fwd");
  return ZS3_OK;
}

extern "C" int zs3_mmd_bwd(const float* gen, const float* real, int M, int N, int D, const float* P,
                           const float* loss, const float* grad_out, float* dgen, float* dreal, void* stream) {
  MmdP p;
  static double dummy;
  int rc = fill_mmd(p, gen, real, M, N, D, reinterpret_cast<const float*>(&dummy), 1, const_cast<float*>(P), &dummy);
  if (rc) return rc;
  ZS3_CHECK_ARG(loss && grad_out, "mmd_bwd: null pointer");
  if (dgen) {
    dim3 grid((D + 127) / 128, M);
    mmd_grad_kernel<<<grid, 128, 0, ST(stream)>>>(p, loss, grad_out, 0, M, dgen);
  }
  if (dreal) {
    dim3 grid((D + 127) / 128, N);
    mmd_grad_kernel<<<grid, 128, 0, ST(stream)>>>(p, loss, grad_out, M, N, dreal);
  }
  ZS3_CHECK_LAUNCH("mmd_bwd");
  return ZS3_OK;
}

extern "C" int zs3_concat2(const float* a, int c1, const float* b, int c2, float* y, long long rows, void* stream) {
  ZS3_CHECK_ARG(a && b && y && c1 > 0 && c2 > 0, "concat2: bad args");
  if (rows == 0) return ZS3_OK;
  long long blocks = (rows * (c1 + c2) + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  concat2_kernel<<<(int)blocks, 256, 0, ST(stream)>>>(a, c1, b, c2, y, rows);
  ZS3_CHECK_LAUNCH("concat2");
  return ZS3_OK;
}
