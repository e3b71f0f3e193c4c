// Fused GMMN generator update: ONE persistent launch runs a whole work list of (image, class) iterations of
// zs3/train_pascal_GMMN.py:211-240 -- gather of the batch_size_generator (=128) sampled rows, GMMNnetwork.forward
// (zs3/modeling/gmmn.py:43-49: Linear -> LeakyReLU(0.2) -> Dropout(0.5) -> Linear), GMMNLoss.moment_loss
// (zs3/utils/loss.py:92-115), the analytic backward and the Adam step (torch.optim.Adam, lr 2e-4) -- back to back,
// with no host round trip between the sequentially dependent updates.
//
// The reference generates features for ALL n_c pixels of the class and then samples 128 rows with replacement
// for the loss (train_pascal_GMMN.py:220-237); only the sampled rows reach the loss, so evaluating the MLP on the
// gathered rows (same noise rows, same Dropout-mask rows) is the same function ("only 128 rows/class
// are algorithmically required").
//
// Layout: all intermediates ([128 x 600] input, [128 x 256] hidden, [256 x 256] sample matrix X = [gen; real],
// the [256 x 256] kernel matrix P, dY, dH) live in a 1.3 MB workspace that stays L2-resident; every operand tile
// is staged through shared memory ([k][32] panels) and consumed by 8 warps that split the k range of the tile
// (in-CTA split-K, fixed-order reduction through shared memory => bit-reproducible results).  The phases of one
// iteration are separated by a grid-wide barrier (the kernel is launched cooperatively, all CTAs co-resident):
//   P1 Hd = dropout(leaky(Xin W1^T + b1)) + gather of item w+1 (idle CTAs)     P4 dY_i = 1/loss * sum_j P_ij (x_j - x_i)
//   P2 Y  = Hd W2^T + b2  -> X[0:B]                 P5 dH  = (dY W2) * f'(Hd)
//   P3 P_ij, loss^2 partials (difference form)      P6 dW1, db1, dW2, db2 -> Adam in the epilogue
// Graph generator (config 5, GMMNnetwork_GCN, zs3/modeling/gmmn.py:52-67): an item with an adjacency matrix A runs both
// layers as pygcn GraphConvolutions, A (x W) + b: P1 / P2 then produce the products without bias, two extra phases
// multiply by A (P1b applies bias, LeakyReLU, Dropout; P2b the bias), and the backward inserts dT2 = A^T dY before P5
// and dT1 = A^T dZ1 before P6 (weights stored [in][out]: `weights_in_out`).  Items flagged ZS3_GMMN_FORWARD_ONLY stop
// after the forward (features of images holding an unseen class, train_context_GMMN_GCNcontext.py:421-428).
// fp32 SIMT on purpose (the MMD exponent cancels catastrophically in reduced precision); the whole
// iteration is 90 MFMA -- the cost is latency (6 barriers + operand staging), not arithmetic.
//
// tests/test_kernel_emulation.py also compiles this file for the host (-DZS3_HOST_EMULATION, tests/emul/cuda_emul.h)
// to check the indexing and phase logic against the oracle without a GPU; the product build never defines it.
#ifdef ZS3_HOST_EMULATION
#include "cuda_emul.h"
#define ZS3_CHECK_ARG(cond, ...) \
  do {                           \
    if (!(cond)) return -1;      \
  } while (0)
#else
#include "common.cuh"
#endif

namespace zs3 {

constexpr int FT = 32;          // output tile edge
constexpr int FTR = 16;         // rows of the tiles of P1-P5 (their 32-row tiling leaves >= half of the grid idle)
constexpr int FKC = 128;        // k extent staged per chunk: 8 warps x 16
constexpr int FLD = 36;         // pitch of a staged [k][32] panel (floats); keeps rows 16-byte aligned
constexpr int FTHREADS = 256;
constexpr int FMAXB = 128;      // max sampled rows per item
constexpr int FGATHER_ROWS = 2;   // rows per gather unit (64 units for a 128-row item)

struct Opnd {
  const float* p;
  int ld;
  int kcontig;  // element(i, k) = kcontig ? p[i*ld + k] : p[k*ld + i]
};

struct FusedP {
  const zs3_gmmn_item* items;
  int n_items;
  int E, Z, H, F;
  float* W1; float* b1; float* W2; float* b2;
  float* mW1; float* mb1; float* mW2; float* mb2;
  float* vW1; float* vb1; float* vW2; float* vb2;
  float* gW1; float* gb1; float* gW2; float* gb2;
  int apply_adam;
  float lr, beta1, beta2, eps;
  long long step0;
  float inv_sigma[8];
  int nsigma;
  float slope, drop_p;
  unsigned long long seed, offset;
  float* losses;
  float* Xin;     // [2][FMAXB][E+Z]
  float* X;       // [2][2*FMAXB][F]
  float* Hd;      // [FMAXB][H]
  float* P;       // [2*FMAXB][2*FMAXB]
  float* dY;      // [FMAXB][F]
  float* dH;      // [FMAXB][H]
  float* T1;      // [FMAXB][H]  graph items: Xin W1 before the adjacency product
  float* T2;      // [FMAXB][F]  graph items: Hd W2 before the adjacency product
  float* dT2;     // [FMAXB][F]  graph items: A^T dY
  float* dT1;     // [FMAXB][H]  graph items: A^T dZ1
  int w_in_out;   // weights stored [in][out] (pygcn) instead of [out][in] (nn.Linear)
  double* lossp;  // [128] per-tile partial sums of loss^2 (P3: 16 x 8 tiles at most)
  unsigned* barrier;
  unsigned long long* stamps;  // optional [n_items][8] %globaltimer at the phase boundaries
};

enum { OP_DOT = 0, OP_DIST = 1, OP_WDIFF = 2 };

__device__ __forceinline__ uint64_t splitmix64f(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}

// All CTAs of the (cooperatively launched) grid arrive; `target` is the running arrival count to wait for.
// bar.sync orders the CTA's writes before thread 0's release-add (cumulativity), the acquire poll orders the other
// CTAs' writes before everything after the second bar.sync; cross-CTA data is read with ld.cg (L2) throughout.
__device__ __forceinline__ void grid_barrier(unsigned* ctr, unsigned& target) {
  target += gridDim.x;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned v;
#ifdef ZS3_HOST_EMULATION
    atomicAdd(ctr, 1u);
    do {
      v = emul_ld_acquire(ctr);
    } while ((int)(v - target) < 0);
#else
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(ctr) : "memory");
    do {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory");
    } while ((int)(v - target) < 0);
#endif
  }
  __syncthreads();
}

__device__ __forceinline__ void phase_stamp(unsigned long long* stamps, int w, int slot) {
#ifndef ZS3_HOST_EMULATION
  if (stamps && blockIdx.x == 0 && threadIdx.x == 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    stamps[w * 8 + slot] = t;
  }
#endif
}

// Panel staging, split in two halves so that the global loads of chunk c+1 are in flight while chunk c is being
// consumed: panel_load reads this thread's PR/2 elements of the [FKC k][PR rows] panel (rows r0.., k range k0..; zero
// outside [R) x [K)) into registers, panel_store writes them to shared memory as S[k][row].  PR = 32 or 16.
template <int PR>
__device__ __forceinline__ void panel_load(float (&v)[PR / 2], const Opnd o, int r0, int R, int k0, int K) {
  if (o.kcontig) {
    const int k = threadIdx.x & (FKC - 1);
    const int rb = (threadIdx.x >> 7) * (PR / 2);  // PR/2 consecutive rows (lanes of a warp: consecutive k -> coalesced)
    const bool kin = (k0 + k) < K;
#pragma unroll
    for (int e = 0; e < PR / 2; ++e) {
      const int r = rb + e;
      v[e] = (kin && (r0 + r) < R) ? __ldcg(o.p + (long long)(r0 + r) * o.ld + (k0 + k)) : 0.f;
    }
  } else {
    const int r = threadIdx.x & (PR - 1);
    const int kb = threadIdx.x / PR;  // 0 .. 256/PR - 1
    const bool rin = (r0 + r) < R;
#pragma unroll
    for (int e = 0; e < PR / 2; ++e) {
      const int k = kb + (FTHREADS / PR) * e;
      v[e] = (rin && (k0 + k) < K) ? __ldcg(o.p + (long long)(k0 + k) * o.ld + (r0 + r)) : 0.f;
    }
  }
}

template <int PR>
__device__ __forceinline__ void panel_store(float* __restrict__ S, const float (&v)[PR / 2], const Opnd o) {
  if (o.kcontig) {
    // one thread = PR/2 consecutive rows of one k: 128-bit stores; a quarter warp (8 consecutive k, pitch 36)
    // covers all 32 banks exactly once
    const int k = threadIdx.x & (FKC - 1);
    const int rb = (threadIdx.x >> 7) * (PR / 2);
#pragma unroll
    for (int e = 0; e < PR / 8; ++e)
      *reinterpret_cast<float4*>(S + k * FLD + rb + 4 * e) = make_float4(v[4 * e], v[4 * e + 1], v[4 * e + 2], v[4 * e + 3]);
  } else {
    const int r = threadIdx.x & (PR - 1);
    const int kb = threadIdx.x / PR;
#pragma unroll
    for (int e = 0; e < PR / 2; ++e) S[(kb + (FTHREADS / PR) * e) * FLD + r] = v[e];
  }
}

// One RT x 32 output tile (RT = 32 or 16 rows): out(i, j) = sum_k f(A(i,k), B(j,k)) with
//   OP_DOT   f = a*b            OP_DIST  f = (a-b)^2            OP_WDIFF f = a*(b - Cmat[i][j])
// 8 warps split every staged k chunk (16 k each); warp partials are summed in warp order through shared memory and
// epi(i, j, value) is called once per in-range output element (consecutive threads = consecutive j).  The 16-row
// variant doubles the number of tiles of a phase whose 32-row tiling would leave most of the grid without work.
template <int OP, int RT, class Epi>
__device__ __forceinline__ void tile32(float* __restrict__ sm, const Opnd A, int i0, int M, const Opnd B, int j0, int N,
                                       int K, const float* __restrict__ Cmat, int ldc, Epi epi) {
  constexpr int RX = RT / 8;  // rows per lane
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int li = lane >> 2, lj = lane & 3;
  float* As = sm;
  float* Bs = sm + FKC * FLD;
  float acc[RX][8], c[RX][8];
#pragma unroll
  for (int x = 0; x < RX; ++x)
#pragma unroll
    for (int y = 0; y < 8; ++y) {
      acc[x][y] = 0.f;
      c[x][y] = 0.f;
    }
  if (OP == OP_WDIFF) {
#pragma unroll
    for (int x = 0; x < RX; ++x)
#pragma unroll
      for (int y = 0; y < 8; ++y) {
        const int i = i0 + li * RX + x, j = j0 + lj * 8 + y;
        if (i < M && j < N) c[x][y] = __ldcg(Cmat + (long long)i * ldc + j);
      }
  }
  float va[RT / 2], vb[16];
  panel_load<RT>(va, A, i0, M, 0, K);
  panel_load<32>(vb, B, j0, N, 0, K);
  for (int k0 = 0; k0 < K; k0 += FKC) {
    __syncthreads();  // the previous chunk (or the previous tile's reduction buffer) is no longer being read
    panel_store<RT>(As, va, A);
    panel_store<32>(Bs, vb, B);
    __syncthreads();
    if (k0 + FKC < K) {  // next chunk's loads overlap this chunk's arithmetic
      panel_load<RT>(va, A, i0, M, k0 + FKC, K);
      panel_load<32>(vb, B, j0, N, k0 + FKC, K);
    }
    if (k0 + warp * 16 < K) {
#pragma unroll 4
      for (int kk = 0; kk < 16; ++kk) {
        const int k = warp * 16 + kk;
        float av[RX];
        if (RX == 4) {
          const float4 a4 = *reinterpret_cast<const float4*>(As + k * FLD + li * 4);
          av[0] = a4.x; av[1] = a4.y; av[RX - 2] = a4.z; av[RX - 1] = a4.w;
        } else {
          const float2 a2 = *reinterpret_cast<const float2*>(As + k * FLD + li * 2);
          av[0] = a2.x; av[1] = a2.y;
        }
        const float4 b0 = *reinterpret_cast<const float4*>(Bs + k * FLD + lj * 8);
        const float4 b1 = *reinterpret_cast<const float4*>(Bs + k * FLD + lj * 8 + 4);
        const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
        for (int x = 0; x < RX; ++x)
#pragma unroll
          for (int y = 0; y < 8; ++y) {
            if (OP == OP_DOT) {
              acc[x][y] = fmaf(av[x], bv[y], acc[x][y]);
            } else if (OP == OP_DIST) {
              const float df = av[x] - bv[y];
              acc[x][y] = fmaf(df, df, acc[x][y]);
            } else {
              acc[x][y] = fmaf(av[x], bv[y] - c[x][y], acc[x][y]);
            }
          }
      }
    }
  }
  __syncthreads();  // every warp is done with the staged panels: reuse the buffer for the reduction
  float* red = sm;  // [8 warps][RT][32]
#pragma unroll
  for (int x = 0; x < RX; ++x) {
    float* dst = red + warp * (RT * FT) + (li * RX + x) * FT + lj * 8;
    *reinterpret_cast<float4*>(dst) = make_float4(acc[x][0], acc[x][1], acc[x][2], acc[x][3]);
    *reinterpret_cast<float4*>(dst + 4) = make_float4(acc[x][4], acc[x][5], acc[x][6], acc[x][7]);
  }
  __syncthreads();
#pragma unroll
  for (int q = 0; q < RT / 8; ++q) {
    const int e = threadIdx.x + FTHREADS * q;
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += red[w * (RT * FT) + e];
    const int i = i0 + (e >> 5), j = j0 + (e & 31);
    if (i < M && j < N) epi(i, j, s);
  }
  // the next tile32 call starts with __syncthreads() before it overwrites `red`
}

// Gather unit `u`: rows [u*FGATHER_ROWS, +FGATHER_ROWS) of item `it` into the packed input / sample buffers.  Every
// thread first issues all of its (independent) loads, then stores: the row indirection + strided feature reads
// are latency, not bandwidth.
__device__ __forceinline__ void gather_rows(const FusedP& p, const zs3_gmmn_item& it, int u, float* __restrict__ Xin,
                                            float* __restrict__ X) {
  const int K1 = p.E + p.Z, B = min(it.rows, FMAXB), F = p.F;
  const int r0 = u * FGATHER_ROWS, r1 = min(r0 + FGATHER_ROWS, B);
  for (int r = r0; r < r1; ++r) {
    const long long er = it.emb.rows ? (long long)__ldg(it.emb.rows + r) : (long long)r;
    const long long zr = it.noise.rows ? (long long)__ldg(it.noise.rows + r) : (long long)r;
    const long long fr = it.real.rows ? (long long)__ldg(it.real.rows + r) : (long long)r;
    const float* eb = it.emb.base + er * it.emb.row_stride;
    const float* zb = it.noise.base + zr * it.noise.row_stride;
    const float* fb = it.real.base + fr * it.real.row_stride;
    for (int k0 = 0; k0 < K1; k0 += 4 * FTHREADS) {
      float v[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int k = k0 + e * FTHREADS + threadIdx.x;
        v[e] = k < p.E ? __ldg(eb + (long long)k * it.emb.col_stride)
                       : (k < K1 ? __ldg(zb + (long long)(k - p.E) * it.noise.col_stride) : 0.f);
      }
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int k = k0 + e * FTHREADS + threadIdx.x;
        if (k < K1) Xin[(long long)r * K1 + k] = v[e];
      }
    }
    for (int d = threadIdx.x; d < F; d += FTHREADS)
      X[(long long)(B + r) * F + d] = __ldg(fb + (long long)d * it.real.col_stride);
  }
}

// torch.optim.Adam update of one element (or a plain gradient store when apply_adam == 0)
__device__ __forceinline__ void apply_grad(const FusedP& p, float* __restrict__ w, float* __restrict__ m,
                                           float* __restrict__ v, float* __restrict__ gout, long long idx, float g,
                                           float bc1, float bc2s) {
  if (p.apply_adam) {
    const float mi = p.beta1 * m[idx] + (1.f - p.beta1) * g;
    const float vi = p.beta2 * v[idx] + (1.f - p.beta2) * g * g;
    m[idx] = mi;
    v[idx] = vi;
    w[idx] -= (p.lr / bc1) * mi / (sqrtf(vi) / bc2s + p.eps);
  } else {
    gout[idx] = g;
  }
}

__global__ void __launch_bounds__(FTHREADS) gmmn_train_fused_kernel(const FusedP p) {
  __shared__ __align__(16) float sm[2 * FKC * FLD];
  __shared__ zs3_gmmn_item sh_item[2];  // [0] current item, [1] next item (gathered during P6)
  __shared__ float sh_red[8];
  __shared__ float sh_scalar[4];        // loss, bc1, sqrt(bc2)

  const int K1 = p.E + p.Z, H = p.H, F = p.F;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  unsigned bar_target = 0;
  const float ks = p.drop_p > 0.f ? 1.f / (1.f - p.drop_p) : 1.f;
  const uint32_t thresh = (uint32_t)(p.drop_p * 65536.0f + 0.5f);

  if (p.n_items <= 0) return;
  if (threadIdx.x == 0) sh_item[0] = p.items[0];
  __syncthreads();
  {
    const int ng0 = (min(sh_item[0].rows, FMAXB) + FGATHER_ROWS - 1) / FGATHER_ROWS;
    for (int u = blockIdx.x; u < ng0; u += gridDim.x) gather_rows(p, sh_item[0], u, p.Xin, p.X);
  }
  grid_barrier(p.barrier, bar_target);

  int n_updates = 0;  // Adam steps taken so far in this launch (forward-only items take none)
  for (int w = 0; w < p.n_items; ++w) {
    __syncthreads();
    if (threadIdx.x == 0) {
      sh_item[0] = p.items[w];
      if (w + 1 < p.n_items) sh_item[1] = p.items[w + 1];
      const double t = (double)(p.step0 + n_updates + 1);
      sh_scalar[1] = (float)(1.0 - pow((double)p.beta1, t));
      sh_scalar[2] = (float)sqrt(1.0 - pow((double)p.beta2, t));
    }
    __syncthreads();
    phase_stamp(p.stamps, w, 0);
    const zs3_gmmn_item& it = sh_item[0];
    const int B = min(it.rows, FMAXB), L = 2 * B;
    const float* adj = it.adj;                    // graph item: both layers are A (x W) + b
    const bool fwd_only = (it.flags & ZS3_GMMN_FORWARD_ONLY) != 0;
    float* Xin = p.Xin + (size_t)(w & 1) * FMAXB * K1;
    float* X = p.X + (size_t)(w & 1) * 2 * FMAXB * F;
    const int tL = (L + FT - 1) / FT, tH = (H + FT - 1) / FT, tF = (F + FT - 1) / FT;
    const int tK1 = (K1 + FT - 1) / FT;
    const int tBr = (B + FTR - 1) / FTR, tLr = (L + FTR - 1) / FTR;  // row tiles of P1-P5
    // weight operands: element (out j, in k) of W1 / W2 in either storage order
    const Opnd W1_jk = p.w_in_out ? Opnd{p.W1, H, 0} : Opnd{p.W1, K1, 1};   // (h, k1)
    const Opnd W2_jk = p.w_in_out ? Opnd{p.W2, F, 0} : Opnd{p.W2, H, 1};    // (f, h)
    const Opnd W2_kj = p.w_in_out ? Opnd{p.W2, F, 1} : Opnd{p.W2, H, 0};    // (h, f): reduction over f

    // epilogue of the hidden layer: bias, LeakyReLU, Dropout -> Hd
    auto hidden_epilogue = [&](int i, int j, float v) {
      v += __ldcg(p.b1 + j);
      v = v > 0.f ? v : v * p.slope;
      if (p.drop_p > 0.f) {
        const long long krow = it.keep_rows ? (long long)__ldg(it.keep_rows + i) : (long long)i;
        bool keep;
        if (it.keep_mask) {
          keep = __ldg(it.keep_mask + krow * H + j) != 0;
        } else {
          const uint64_t idx = (uint64_t)krow * (uint64_t)H + (uint64_t)j;
          const uint64_t hsh = splitmix64f(p.seed ^ splitmix64f(p.offset + ((uint64_t)w << 40) + (idx >> 2)));
          keep = ((uint32_t)(hsh >> (16 * (idx & 3))) & 0xFFFFu) >= thresh;
        }
        v = keep ? v * ks : 0.f;
      }
      p.Hd[(long long)i * H + j] = v;
    };
    auto output_epilogue = [&](int i, int j, float v) {
      v += __ldcg(p.b2 + j);
      X[(long long)i * F + j] = v;
      if (it.out) it.out[(long long)i * F + j] = v;
    };

    // ---- P1: Hd = dropout(leaky(Xin W1^T + b1)) (graph items: T1 = Xin W1 only); the CTAs without a tile gather the
    //          rows of item w+1 into the other Xin / X buffers (last read in P6 of item w-1, i.e. before the previous
    //          grid barrier)
    const int ng = (w + 1 < p.n_items) ? (min(sh_item[1].rows, FMAXB) + FGATHER_ROWS - 1) / FGATHER_ROWS : 0;
    for (int u = blockIdx.x; u < tBr * tH + ng; u += gridDim.x) {
      if (u >= tBr * tH) {
        gather_rows(p, sh_item[1], u - tBr * tH, p.Xin + (size_t)((w + 1) & 1) * FMAXB * K1,
                    p.X + (size_t)((w + 1) & 1) * 2 * FMAXB * F);
        continue;
      }
      const int i0 = (u / tH) * FTR, j0 = (u % tH) * FT;
      if (adj)
        tile32<OP_DOT, FTR>(sm, Opnd{Xin, K1, 1}, i0, B, W1_jk, j0, H, K1, nullptr, 0,
                            [&](int i, int j, float v) { p.T1[(long long)i * H + j] = v; });
      else
        tile32<OP_DOT, FTR>(sm, Opnd{Xin, K1, 1}, i0, B, W1_jk, j0, H, K1, nullptr, 0, hidden_epilogue);
    }
    grid_barrier(p.barrier, bar_target);
    if (adj) {  // ---- P1b: Hd = dropout(leaky(A T1 + b1))
      for (int u = blockIdx.x; u < tBr * tH; u += gridDim.x) {
        const int i0 = (u / tH) * FTR, j0 = (u % tH) * FT;
        tile32<OP_DOT, FTR>(sm, Opnd{adj, B, 1}, i0, B, Opnd{p.T1, H, 0}, j0, H, B, nullptr, 0, hidden_epilogue);
      }
      grid_barrier(p.barrier, bar_target);
    }
    phase_stamp(p.stamps, w, 1);

    // ---- P2: Y = Hd W2^T + b2 -> X[0:B] (graph items: T2 = Hd W2, then P2b: Y = A T2 + b2)
    for (int u = blockIdx.x; u < tBr * tF; u += gridDim.x) {
      const int i0 = (u / tF) * FTR, j0 = (u % tF) * FT;
      if (adj)
        tile32<OP_DOT, FTR>(sm, Opnd{p.Hd, H, 1}, i0, B, W2_jk, j0, F, H, nullptr, 0,
                            [&](int i, int j, float v) { p.T2[(long long)i * F + j] = v; });
      else
        tile32<OP_DOT, FTR>(sm, Opnd{p.Hd, H, 1}, i0, B, W2_jk, j0, F, H, nullptr, 0, output_epilogue);
    }
    grid_barrier(p.barrier, bar_target);
    if (adj) {
      for (int u = blockIdx.x; u < tBr * tF; u += gridDim.x) {
        const int i0 = (u / tF) * FTR, j0 = (u % tF) * FT;
        tile32<OP_DOT, FTR>(sm, Opnd{adj, B, 1}, i0, B, Opnd{p.T2, F, 0}, j0, F, B, nullptr, 0, output_epilogue);
      }
      grid_barrier(p.barrier, bar_target);
    }
    phase_stamp(p.stamps, w, 2);
    if (fwd_only) {  // uniform over the grid: every CTA skips the same barriers
      if (blockIdx.x == 0 && threadIdx.x == 0) p.losses[w] = 0.f;
      continue;
    }

    // ---- P3: P_ij = s_i s_j sum_sigma exp(e_ij/sigma)/sigma, loss^2 partial per tile; e_ij = -|x_i - x_j|^2 / 2.
    // get_scale_matrix quirk (loss.py:92-97): the FIRST N rows carry +1/N, the last M rows -1/M (M = N = B here).
    for (int u = blockIdx.x; u < tLr * tL; u += gridDim.x) {
      const int i0 = (u / tL) * FTR, j0 = (u % tL) * FT;
      float part = 0.f;
      tile32<OP_DIST, FTR>(sm, Opnd{X, F, 1}, i0, L, Opnd{X, F, 1}, j0, L, F, nullptr, 0, [&](int i, int j, float d2) {
        const float si = i < B ? 1.f / B : -1.f / B;
        const float sj = j < B ? 1.f / B : -1.f / B;
        const float e = -0.5f * d2;
        float kv = 0.f, kp = 0.f;
        for (int s = 0; s < p.nsigma; ++s) {
          const float ex = expf(e * p.inv_sigma[s]);
          kv += ex;
          kp += ex * p.inv_sigma[s];
        }
        part += si * sj * kv;
        p.P[(long long)i * L + j] = si * sj * kp;
      });
      for (int off = 16; off; off >>= 1) part += __shfl_xor_sync(0xffffffffu, part, off);
      if (lane == 0) sh_red[warp] = part;
      __syncthreads();
      if (threadIdx.x == 0) {
        double s = 0;
        for (int a = 0; a < 8; ++a) s += (double)sh_red[a];
        p.lossp[u] = s;
      }
      __syncthreads();
    }
    grid_barrier(p.barrier, bar_target);
    phase_stamp(p.stamps, w, 3);

    // ---- P4: loss = sqrt(sum of partials) (NaN for a negative sum, like the reference: loss.py:114);
    //          dY_i = 1/loss * sum_j P_ij (x_j - x_i), i < B
    {  // the <= 128 partials are fetched by 128 threads at once (one L2 round trip instead of 128 dependent ones)
       // and summed from shared memory in index order, so every CTA derives the same bits
      double* part = reinterpret_cast<double*>(sm);  // the staging buffer is free between phases
      if (threadIdx.x < 128) part[threadIdx.x] = threadIdx.x < tLr * tL ? __ldcg(p.lossp + threadIdx.x) : 0.0;
      __syncthreads();
      if (threadIdx.x == 0) {
        double s = 0;
        for (int t = 0; t < tLr * tL; ++t) s += part[t];
        const float loss = sqrtf((float)s);
        sh_scalar[0] = loss;
        if (blockIdx.x == 0) p.losses[w] = loss;
      }
      __syncthreads();
    }
    {
      const float inv_loss = 1.f / sh_scalar[0];
      for (int u = blockIdx.x; u < tBr * tF; u += gridDim.x) {
        const int i0 = (u / tF) * FTR, j0 = (u % tF) * FT;
        tile32<OP_WDIFF, FTR>(sm, Opnd{p.P, L, 1}, i0, B, Opnd{X, F, 0}, j0, F, L, X, F,
                         [&](int i, int j, float v) { p.dY[(long long)i * F + j] = inv_loss * v; });
      }
    }
    grid_barrier(p.barrier, bar_target);
    if (adj) {  // ---- P4b: dT2 = A^T dY (gradient of T2 = Hd W2 through Y = A T2 + b2)
      for (int u = blockIdx.x; u < tBr * tF; u += gridDim.x) {
        const int i0 = (u / tF) * FTR, j0 = (u % tF) * FT;
        tile32<OP_DOT, FTR>(sm, Opnd{adj, B, 0}, i0, B, Opnd{p.dY, F, 0}, j0, F, B, nullptr, 0,
                            [&](int i, int j, float v) { p.dT2[(long long)i * F + j] = v; });
      }
      grid_barrier(p.barrier, bar_target);
    }
    phase_stamp(p.stamps, w, 4);
    const float* G2 = adj ? p.dT2 : p.dY;   // gradient of the output layer's matrix product

    // ---- P5: dZ1 = (G2 W2) * d/dh[dropout(leaky(h))], reconstructed from the forward output Hd
    for (int u = blockIdx.x; u < tBr * tH; u += gridDim.x) {
      const int i0 = (u / tH) * FTR, j0 = (u % tH) * FT;
      tile32<OP_DOT, FTR>(sm, Opnd{G2, F, 1}, i0, B, W2_kj, j0, H, F, nullptr, 0, [&](int i, int j, float v) {
        const float hv = __ldcg(p.Hd + (long long)i * H + j);
        p.dH[(long long)i * H + j] = v * (hv > 0.f ? ks : (hv < 0.f ? p.slope * ks : 0.f));
      });
    }
    grid_barrier(p.barrier, bar_target);
    if (adj) {  // ---- P5b: dT1 = A^T dZ1
      for (int u = blockIdx.x; u < tBr * tH; u += gridDim.x) {
        const int i0 = (u / tH) * FTR, j0 = (u % tH) * FT;
        tile32<OP_DOT, FTR>(sm, Opnd{adj, B, 0}, i0, B, Opnd{p.dH, H, 0}, j0, H, B, nullptr, 0,
                            [&](int i, int j, float v) { p.dT1[(long long)i * H + j] = v; });
      }
      grid_barrier(p.barrier, bar_target);
    }
    phase_stamp(p.stamps, w, 5);
    const float* G1 = adj ? p.dT1 : p.dH;   // gradient of the hidden layer's matrix product

    // ---- P6: dW1 = G1^T Xin, dW2 = G2^T Hd, db1 = colsum(dZ1), db2 = colsum(dY) -> Adam
    {
      const float bc1 = sh_scalar[1], bc2s = sh_scalar[2];
      const int n1 = tH * tK1, n2 = tF * tH;
      const int total = 2 + n1 + n2;
      // units 0 / 1 = the bias gradients (column sums over the B rows: 32 independent loads in flight per thread,
      // 4 round trips instead of B dependent ones); they come first so that they never queue behind a tile
      for (int u = blockIdx.x; u < total; u += gridDim.x) {
        if (u == 0) {
          for (int h = threadIdx.x; h < H; h += FTHREADS) {
            float s = 0.f;
#pragma unroll 32
            for (int r = 0; r < B; ++r) s += __ldcg(p.dH + (long long)r * H + h);
            apply_grad(p, p.b1, p.mb1, p.vb1, p.gb1, h, s, bc1, bc2s);
          }
        } else if (u == 1) {
          for (int o = threadIdx.x; o < F; o += FTHREADS) {
            float s = 0.f;
#pragma unroll 32
            for (int r = 0; r < B; ++r) s += __ldcg(p.dY + (long long)r * F + o);
            apply_grad(p, p.b2, p.mb2, p.vb2, p.gb2, o, s, bc1, bc2s);
          }
        } else if (u < 2 + n1) {
          const int t = u - 2;
          const int i0 = (t / tK1) * FT, j0 = (t % tK1) * FT;
          tile32<OP_DOT, FT>(sm, Opnd{G1, H, 0}, i0, H, Opnd{Xin, K1, 0}, j0, K1, B, nullptr, 0,
                         [&](int i, int j, float v) {
                           apply_grad(p, p.W1, p.mW1, p.vW1, p.gW1,
                                      p.w_in_out ? (long long)j * H + i : (long long)i * K1 + j, v, bc1, bc2s);
                         });
        } else {
          const int t = u - 2 - n1;
          const int i0 = (t / tH) * FT, j0 = (t % tH) * FT;
          tile32<OP_DOT, FT>(sm, Opnd{G2, F, 0}, i0, F, Opnd{p.Hd, H, 0}, j0, H, B, nullptr, 0,
                         [&](int i, int j, float v) {
                           apply_grad(p, p.W2, p.mW2, p.vW2, p.gW2,
                                      p.w_in_out ? (long long)j * F + i : (long long)i * H + j, v, bc1, bc2s);
                         });
        }
      }
    }
    grid_barrier(p.barrier, bar_target);
    phase_stamp(p.stamps, w, 6);
    ++n_updates;
  }
}

static size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

struct FusedLayout {
  size_t xin, x, hd, pm, dy, dh, t1, t2, dt2, dt1, lossp, barrier, total;
};

static FusedLayout fused_layout(int E, int Z, int H, int F) {
  FusedLayout l;
  size_t o = 0;
  l.xin = o; o += align256(sizeof(float) * 2 * FMAXB * (size_t)(E + Z));
  l.x = o; o += align256(sizeof(float) * 2 * 2 * FMAXB * (size_t)F);
  l.hd = o; o += align256(sizeof(float) * FMAXB * (size_t)H);
  l.pm = o; o += align256(sizeof(float) * 4 * FMAXB * FMAXB);
  l.dy = o; o += align256(sizeof(float) * FMAXB * (size_t)F);
  l.dh = o; o += align256(sizeof(float) * FMAXB * (size_t)H);
  l.t1 = o; o += align256(sizeof(float) * FMAXB * (size_t)H);
  l.t2 = o; o += align256(sizeof(float) * FMAXB * (size_t)F);
  l.dt2 = o; o += align256(sizeof(float) * FMAXB * (size_t)F);
  l.dt1 = o; o += align256(sizeof(float) * FMAXB * (size_t)H);
  l.lossp = o; o += align256(sizeof(double) * 128);
  l.barrier = o; o += 256;
  l.total = o;
  return l;
}

}  // namespace zs3

using namespace zs3;

#ifdef ZS3_HOST_EMULATION
extern "C" unsigned long long zs3_emul_gmmn_train_workspace_size(
#else
extern "C" unsigned long long zs3_gmmn_train_workspace_size(
#endif
int embed_dim, int noise_dim, int hidden, int feat) {
  if (embed_dim <= 0 || noise_dim < 0 || hidden <= 0 || feat <= 0) return 0;
  return (unsigned long long)fused_layout(embed_dim, noise_dim, hidden, feat).total;
}

#ifdef ZS3_HOST_EMULATION
extern "C" int zs3_emul_gmmn_train_fused(const zs3_gmmn_train_args* a, void* stream) {
#else
extern "C" int zs3_gmmn_train_fused(const zs3_gmmn_train_args* a, void* stream) {
#endif
  ZS3_CHECK_ARG(a != nullptr, "gmmn_train_fused: null args");
  ZS3_CHECK_ARG(a->n_items >= 0 && (a->n_items == 0 || a->items), "gmmn_train_fused: bad item list");
  ZS3_CHECK_ARG(a->embed_dim > 0 && a->noise_dim >= 0 && a->hidden > 0 && a->feat > 0, "gmmn_train_fused: bad dims");
  ZS3_CHECK_ARG(a->w1 && a->b1 && a->w2 && a->b2 && a->losses, "gmmn_train_fused: null parameter / loss pointer");
  ZS3_CHECK_ARG(a->nsigma > 0 && a->nsigma <= 8, "gmmn_train_fused: nsigma=%d", a->nsigma);
  ZS3_CHECK_ARG(a->drop_p >= 0.f && a->drop_p < 1.f, "gmmn_train_fused: drop_p=%f", a->drop_p);
  ZS3_CHECK_ARG(a->max_rows > 0 && a->max_rows <= FMAXB, "gmmn_train_fused: max_rows=%d (1..%d)", a->max_rows, FMAXB);
  if (a->apply_adam) {
    for (int i = 0; i < 4; ++i)
      ZS3_CHECK_ARG(a->adam_m[i] && a->adam_v[i], "gmmn_train_fused: apply_adam needs exp_avg / exp_avg_sq buffers");
    ZS3_CHECK_ARG(a->step0 >= 0, "gmmn_train_fused: step0 < 0");
  } else {
    for (int i = 0; i < 4; ++i) ZS3_CHECK_ARG(a->grad[i], "gmmn_train_fused: apply_adam = 0 needs gradient outputs");
    ZS3_CHECK_ARG(a->n_items <= 1, "gmmn_train_fused: gradient output mode takes one item");
  }
  const FusedLayout l = fused_layout(a->embed_dim, a->noise_dim, a->hidden, a->feat);
  ZS3_CHECK_ARG(a->workspace && a->workspace_bytes >= l.total, "gmmn_train_fused: workspace too small (%llu < %llu)",
                (unsigned long long)a->workspace_bytes, (unsigned long long)l.total);
  ZS3_CHECK_ARG((reinterpret_cast<uintptr_t>(a->workspace) & 15) == 0, "gmmn_train_fused: workspace not 16-byte aligned");
  if (a->n_items == 0) return ZS3_OK;

  FusedP p;
  p.items = a->items; p.n_items = a->n_items;
  p.E = a->embed_dim; p.Z = a->noise_dim; p.H = a->hidden; p.F = a->feat;
  p.W1 = a->w1; p.b1 = a->b1; p.W2 = a->w2; p.b2 = a->b2;
  p.mW1 = a->adam_m[0]; p.mb1 = a->adam_m[1]; p.mW2 = a->adam_m[2]; p.mb2 = a->adam_m[3];
  p.vW1 = a->adam_v[0]; p.vb1 = a->adam_v[1]; p.vW2 = a->adam_v[2]; p.vb2 = a->adam_v[3];
  p.gW1 = a->grad[0]; p.gb1 = a->grad[1]; p.gW2 = a->grad[2]; p.gb2 = a->grad[3];
  p.apply_adam = a->apply_adam;
  p.lr = a->lr; p.beta1 = a->beta1; p.beta2 = a->beta2; p.eps = a->eps; p.step0 = a->step0;
  for (int i = 0; i < 8; ++i) p.inv_sigma[i] = i < a->nsigma ? 1.f / a->sigma[i] : 1.f;
  p.nsigma = a->nsigma;
  p.slope = a->slope; p.drop_p = a->drop_p; p.seed = a->seed; p.offset = a->offset;
  p.losses = a->losses;
  char* ws = static_cast<char*>(a->workspace);
  p.Xin = reinterpret_cast<float*>(ws + l.xin);
  p.X = reinterpret_cast<float*>(ws + l.x);
  p.Hd = reinterpret_cast<float*>(ws + l.hd);
  p.P = reinterpret_cast<float*>(ws + l.pm);
  p.dY = reinterpret_cast<float*>(ws + l.dy);
  p.dH = reinterpret_cast<float*>(ws + l.dh);
  p.T1 = reinterpret_cast<float*>(ws + l.t1);
  p.T2 = reinterpret_cast<float*>(ws + l.t2);
  p.dT2 = reinterpret_cast<float*>(ws + l.dt2);
  p.dT1 = reinterpret_cast<float*>(ws + l.dt1);
  p.w_in_out = a->weights_in_out ? 1 : 0;
  p.lossp = reinterpret_cast<double*>(ws + l.lossp);
  p.barrier = reinterpret_cast<unsigned*>(ws + l.barrier);
  p.stamps = a->phase_stamps;

#ifdef ZS3_HOST_EMULATION
  // `stream` carries the number of emulated thread blocks (tests/test_kernel_emulation.py)
  const int nblocks = stream ? (int)reinterpret_cast<intptr_t>(stream) : 1;
  *p.barrier = 0;
  return emul_launch_grid<FusedP>(gmmn_train_fused_kernel, p, FTHREADS, nblocks);
#else
  // all CTAs must be co-resident (grid-wide barrier): size the grid from the occupancy of this device
  static int grid_cached = 0;
  if (grid_cached == 0) {
    int dev = 0, sms = 0, per_sm = 0;
    if (cudaGetDevice(&dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess ||
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, gmmn_train_fused_kernel, FTHREADS, 0) != cudaSuccess ||
        sms <= 0 || per_sm <= 0) {
      cudaGetLastError();
      zs3::set_error("gmmn_train_fused: cannot query the device occupancy");
      return ZS3_ERR_DRIVER;
    }
    grid_cached = sms < 128 ? sms : 128;  // one CTA per SM; 128 >= the widest phase that matters (P6 runs 2 rounds)
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (cudaMemsetAsync(p.barrier, 0, sizeof(unsigned), st) != cudaSuccess) {
    zs3::set_error("gmmn_train_fused: cudaMemsetAsync failed: %s", cudaGetErrorString(cudaGetLastError()));
    return ZS3_ERR_LAUNCH;
  }
  void* kargs[] = {&p};
  cudaError_t e = cudaLaunchCooperativeKernel(reinterpret_cast<void*>(gmmn_train_fused_kernel), dim3(grid_cached),
                                              dim3(FTHREADS), kargs, 0, st);
  if (e != cudaSuccess) {
    cudaGetLastError();
    zs3::set_error("gmmn_train_fused: cooperative launch failed: %s", cudaGetErrorString(e));
    return ZS3_ERR_LAUNCH;
  }
  ZS3_CHECK_LAUNCH("gmmn_train_fused");
  return ZS3_OK;
#endif
}
