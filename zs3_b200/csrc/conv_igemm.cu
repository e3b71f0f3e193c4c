// Implicit-GEMM convolution on the 5th-gen tensor cores (sm_100a).
//
//   fprop / dgrad :  Y[m][n]  = sum_k A[m][k] * Wt[n][k]     m = output pixel, n = output channel,
//                                                            k = (segment, tap, input channel)
//   wgrad         :  dW[n][c] = sum_p dY[p][n] * X[p@tap][c] p = output pixel (the reduction)
//
// Data movement: A tiles are fetched by TMA in im2col mode straight from the NHWC activation tensor
// (zero padding, stride and dilation are resolved by the TMA unit: no index math in the SM), weight
// tiles by tiled TMA; both land in 128B-swizzled shared memory that tcgen05.mma consumes through
// shared-memory descriptors.  Accumulators live in TMEM (double buffered), epilogue warps read them back
// with tcgen05.ld, fuse bias / BatchNorm batch statistics / accumulate, and store NHWC.
//
// Warp roles (384 threads, 1 persistent CTA per SM):
//   warp 0: TMA producer (one elected lane)      warp 1: MMA issuer (one elected lane)
//   warp 2: TMEM allocator                       warps 4-11: epilogue (TMEM lane quarter = warp%4, two warps
//                                                per quarter split the 32-column chunks between them)
//
// Reference call sites replaced: every nn.Conv2d on the DeepLab path (see include/zs3b200.h).
#include <stdlib.h>

#include "common.cuh"
#include "ptx.cuh"

namespace zs3 {

constexpr int BLOCK_M = 128;      // output pixels per tile (= TMEM lanes)
constexpr int BLOCK_K = 64;       // bf16 elements per k-block = one 128-byte swizzle row
constexpr int A_TILE_BYTES = BLOCK_M * BLOCK_K * 2;  // 16 KiB
constexpr int NUM_THREADS = 384;   // fprop: 4 control warps + 8 epilogue warps
constexpr int WG_THREADS = 384;    // wgrad: 4 control warps + 8 epilogue warps
constexpr int EPI_WARP0 = 4;
constexpr int NUM_EPI_WARPS = 8;
constexpr int MAX_STAT_CH = 2048;  // per-CTA shared-memory BatchNorm statistic accumulators

struct alignas(64) FpropSegment {
  CUtensorMap a;  // im2col map over the segment's activation tensor
  CUtensorMap b;  // tiled 3-D map over the segment's packed weights [cout_pad][taps][cin_pad]
  int num_cblk;   // cin_pad / 64
};

struct FpropParams {
  FpropSegment seg[ZS3_MAX_SEGMENTS];
  int b_mn;            // 1: data-gradient mode on forward-packed weights W[k][tap][n] (MN-major B tiles, flipped taps)
  CUtensorMap ymap;    // tiled 2-D map over y [M][cout_pad] (box 32 rows x 32 channels, SWIZZLE_64B) for the TMA-store epilogue
  int use_tma_store;   // dense bf16 output: stage through shared memory and store with TMA (accumulate: reduce-add)
  int num_segments;
  int M;          // N*Ho*Wo
  int HoWo, Wo;
  int stride, pad, dil, R, S;
  int cout_pad;
  int num_m_tiles, num_n_tiles;
  int num_m_groups;  // ceil(num_m_tiles / cluster size): the CTAs of a cluster take consecutive m-tiles of one n-tile
  int kb_per_tile;   // total k-blocks per output tile
  int kb_per_row;    // k-blocks per filter row r (kb_per_tile / R)
  int cull;          // skip the filter rows whose taps lie in the zero padding for EVERY pixel of a tile
  int H, Ho;         // input / output rows (tap culling)
  void* y;
  long long y_cstride;
  // output pixel (n,p,q) is stored at pixel index n*y_img + p*y_row + q*y_pix of y
  long long y_img, y_row, y_pix;
  int y_is_f32;
  int accumulate;
  const float* bias;
  double* stat_sum;
  double* stat_sqsum;
  // inference epilogue (frozen / eval-mode BatchNorm folded in): out = relu?(acc*scale[n] + shift[n] (+ residual))
  const float* ep_scale;
  const float* ep_shift;
  const __nv_bfloat16* ep_res;
  long long ep_res_cs;
  int ep_relu;
};

template <int BN, int STAGES>
struct FpropSmem {
  static constexpr int B_TILE_BYTES = BN * BLOCK_K * 2;
  static constexpr int STAGE_BYTES = A_TILE_BYTES + B_TILE_BYTES;
  static constexpr int STAGING_OFFSET = STAGES * STAGE_BYTES;      // 1024-byte aligned: the 64B swizzle is address based
  static constexpr int STAGING_BYTES = NUM_EPI_WARPS * 32 * 64;    // one [32 rows][32 ch] bf16 box (2 KiB) per warp
  // the statistics partials (2 x 2048 floats) double as a SECOND set of staging boxes when no statistics are fused
  // (short-K layers, where the epilogue is the bottleneck and store latency must be overlapped)
  static constexpr int STAT_OFFSET = STAGING_OFFSET + STAGING_BYTES;
  static constexpr int BAR_OFFSET = STAT_OFFSET + 2 * MAX_STAT_CH * 4;
  static constexpr int TOTAL = BAR_OFFSET + 256 /*barriers + tmem slot*/ + 1024 /*alignment slack*/;
  static_assert(2 * MAX_STAT_CH * 4 >= STAGING_BYTES, "statistics region must hold the second staging set");
};

// Filter rows r whose taps touch at least one real input row for some output pixel of the tile starting at pixel m0
// (bit r of the result).  A filter row outside that set reads nothing but zero padding for the whole tile -- at 33x33
// with dilation 12/18 (ASPP, aspp.py:57-80) or 4/8 (layer4's multi-grid blocks) that is up to a third of the k-blocks
// of a tile -- so the producer does not fetch it and the MMA issuer does not multiply it.  Output row op reads input
// row op*stride - pad + r*dil; the tile covers rows [op0, Ho) of its first image, every row of the images in between
// and rows [0, op1] of its last image.
__device__ __forceinline__ uint32_t valid_filter_rows(const FpropParams& p, int m0) {
  const uint32_t all = (1u << p.R) - 1u;
  if (!p.cull) return all;
  int m_last = m0 + BLOCK_M;
  if (m_last > p.M) m_last = p.M;
  m_last -= 1;
  if (m_last < m0) return 1u;  // a tile past the last pixel (cluster padding): nothing but TMA zero fill
  const int img0 = m0 / p.HoWo, img1 = m_last / p.HoWo;
  const int op0 = (m0 - img0 * p.HoWo) / p.Wo, op1 = (m_last - img1 * p.HoWo) / p.Wo;
  uint32_t mask = 0;
  for (int r = 0; r < p.R; ++r) {
    const int t = p.pad - r * p.dil;
    const int lo = t > 0 ? (t + p.stride - 1) / p.stride : 0;  // first output row whose tap r is inside the input
    const int u = p.H - 1 + t;
    int hi = u >= 0 ? u / p.stride : -1;                       // last such output row
    if (hi > p.Ho - 1) hi = p.Ho - 1;
    bool v;
    if (img0 == img1)
      v = (lo > op0 ? lo : op0) <= (hi < op1 ? hi : op1);
    else if (img1 - img0 >= 2)
      v = lo <= hi;
    else
      v = ((lo > op0 ? lo : op0) <= hi) || (lo <= (hi < op1 ? hi : op1));
    mask |= (v ? 1u : 0u) << r;
  }
  return mask ? mask : 1u;
}

// Column sums over the 32 lanes of a warp: lane j ends up with sum over lanes of v[j].
// Recursive halving: 31 shuffles instead of 32*5.
__device__ __forceinline__ float warp_column_sums(float (&v)[32], int lane) {
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    const bool upper = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < off; ++i) {
      float keep = upper ? v[i + off] : v[i];
      float send = upper ? v[i] : v[i + off];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
  return v[0];
}

// MODE selects the epilogue that is compiled in (the hot variants stay small enough for the instruction cache):
//   0  bf16 output through the staged TMA store (or TMA reduce-add when accumulating), nothing else;
//   1  same + BatchNorm batch statistics kept in REGISTERS across the CTA's tiles (needs one n-tile per CTA, i.e.
//      gridDim.x % num_n_tiles == 0) and flushed once at the end;
//   2  generic: bias, fp32 / strided / read-modify-write outputs, the folded inference epilogue, statistics through
//      per-CTA shared-memory partials;
//   3  variant 0 + the folded inference epilogue (frozen / eval-mode BatchNorm: out = relu?(acc*scale[n] + shift[n]
//      (+ residual))) on the pipelined TMEM-read / staged TMA-store path: the frozen-backbone feature extraction of
//      ZS3Net's step 2 (train_pascal_GMMN.py:154-157) and validation run every conv -> BN -> ReLU layer as this ONE kernel.
template <int BN, int STAGES, int MODE>
__global__ void __launch_bounds__(NUM_THREADS, 1) conv_fprop_kernel(const __grid_constant__ FpropParams p) {
  using L = FpropSmem<BN, STAGES>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::BAR_OFFSET);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* acc_full = empty_bar + STAGES;   // [2]
  uint64_t* acc_empty = acc_full + 2;        // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
  float* s_sum = reinterpret_cast<float*>(smem + L::STAT_OFFSET);  // [MAX_STAT_CH] per-CTA partial sums
  float* s_sq = s_sum + MAX_STAT_CH;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  // Thread-block cluster (1, 2 or 4 CTAs): the CTAs of a cluster work on consecutive m-tiles of the SAME n-tile in
  // lock step; each loads 1/csize of the weight tile and multicasts it to all, cutting L2->SM weight traffic.
  const int csize = (int)cluster_nctarank();
  const int crank = (int)cluster_ctarank();
  const uint16_t cmask = (uint16_t)((1u << csize) - 1u);
  const int num_groups = p.num_m_groups * p.num_n_tiles;
  const int cluster_id = blockIdx.x / csize;
  const int num_clusters = gridDim.x / csize;
  if (MODE == 2 && p.stat_sum != nullptr) {
    for (int i = threadIdx.x; i < p.cout_pad; i += NUM_THREADS) {
      s_sum[i] = 0.f;
      s_sq[i] = 0.f;
    }
  }

  if (threadIdx.x == 0) {
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], csize);  // one tcgen05.commit per CTA of the cluster
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&acc_full[i], 1);
      mbar_init(&acc_empty[i], NUM_EPI_WARPS);  // one arrive per epilogue warp
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, 2 * BN);
    tmem_relinquish();
  }
  if (warp == 0 && lane == 0) {
    for (int s = 0; s < p.num_segments; ++s) {
      tma_prefetch_desc(&p.seg[s].a);
      tma_prefetch_desc(&p.seg[s].b);
    }
    if (p.use_tma_store) tma_prefetch_desc(&p.ymap);
  }
  tc_fence_before();
  __syncthreads();
  if (csize > 1) cluster_sync_all();  // peers' barriers must be initialised before any multicast reaches them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  griddep_sync();  // everything above overlapped the previous kernel's tail; its results are visible from here on

  if (warp == 0) {
    // ===================================================================== TMA producer
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      const int b_rows = BN / csize;  // rows of the weight tile this CTA fetches (and multicasts)
      for (int g = cluster_id; g < num_groups; g += num_clusters) {
        const int n_tile = g % p.num_n_tiles;
        const int m_tile = (g / p.num_n_tiles) * csize + crank;  // may be >= num_m_tiles: TMA zero-fills, epilogue masks
        const int m0 = m_tile * BLOCK_M;
        const int img = m0 / p.HoWo;
        const int rem = m0 - img * p.HoWo;
        const int op = rem / p.Wo;
        const int oq = rem - op * p.Wo;
        const int base_w = oq * p.stride - p.pad;
        const int base_h = op * p.stride - p.pad;
        const uint32_t rows = valid_filter_rows(p, m0);
        for (int s = 0; s < p.num_segments; ++s) {
          const FpropSegment& sg = p.seg[s];
          for (int r = 0; r < p.R; ++r) {
            if (!((rows >> r) & 1u)) continue;
            for (int q = 0; q < p.S; ++q) {
              const int tap = r * p.S + q;
              for (int cb = 0; cb < sg.num_cblk; ++cb) {
                mbar_wait(&empty_bar[stage], phase ^ 1);
                uint8_t* sa = smem + stage * L::STAGE_BYTES;
                uint8_t* sb = sa + A_TILE_BYTES;
                mbar_arrive_expect_tx(&full_bar[stage], L::STAGE_BYTES);
                tma_load_im2col_4d(sa, &sg.a, &full_bar[stage], cb * BLOCK_K, base_w, base_h, img,
                                   (uint16_t)(q * p.dil), (uint16_t)(r * p.dil));
                if (p.b_mn) {
                  // weights stay in their forward layout [k = forward cout][tap][n = forward cin]: BN/64 boxes of
                  // [64 k rows][64 n] form MN-major B tiles; the data gradient uses the spatially flipped tap
                  const int ftap = (p.R - 1 - r) * p.S + (p.S - 1 - q);
#pragma unroll
                  for (int j = 0; j < BN / 64; ++j)
                    tma_load_3d(sb + j * 8192, &sg.b, &full_bar[stage], n_tile * BN + j * 64, ftap, cb * BLOCK_K);
                } else if (csize == 1)
                  tma_load_3d(sb, &sg.b, &full_bar[stage], cb * BLOCK_K, tap, n_tile * BN);
                else
                  tma_load_3d_mc(sb + crank * b_rows * 128, &sg.b, &full_bar[stage], cb * BLOCK_K, tap,
                                 n_tile * BN + crank * b_rows, cmask);
                if (++stage == STAGES) {
                  stage = 0;
                  phase ^= 1;
                }
              }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ======================================================================= MMA issuer
    if (elect_one()) {
      const uint32_t idesc = make_idesc_bf16(BLOCK_M, BN, 0, p.b_mn);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int g = cluster_id; g < num_groups; g += num_clusters, ++it) {
        const int as = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        mbar_wait(&acc_empty[as], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + as * BN;
        // the issuer only needs the NUMBER of k-blocks the producer fetches for this tile
        const int m_tile_i = (g / p.num_n_tiles) * csize + crank;
        const int nkb = p.cull ? __popc(valid_filter_rows(p, m_tile_i * BLOCK_M)) * p.kb_per_row : p.kb_per_tile;
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * L::STAGE_BYTES);
          const uint32_t sb = sa + A_TILE_BYTES;
          const uint64_t adesc = make_smem_desc_sw128(sa, 16, 1024);
          if (p.b_mn) {
            // MN-major B: LBO = 8192 B between 64-column sub-tiles, SBO = 1024 B per 8 k rows, 2048 B per UMMA_K
#pragma unroll
            for (int k = 0; k < BLOCK_K / 16; ++k)
              umma_f16(tmem_d, adesc + 2 * k, make_smem_desc_sw128(sb + k * 2048, 8192, 1024), idesc, (kb | k) != 0);
          } else {
            const uint64_t bdesc = make_smem_desc_sw128(sb, 16, 1024);
#pragma unroll
            for (int k = 0; k < BLOCK_K / 16; ++k) {
              // advance 16 bf16 = 32 bytes inside the 128-byte swizzle row: +2 in the (addr >> 4) field
              umma_f16(tmem_d, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
            }
          }
          if (csize == 1)
            umma_commit(&empty_bar[stage]);
          else
            umma_commit_mc(&empty_bar[stage], cmask);  // the stage is free once EVERY CTA of the cluster has read it
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit(&acc_full[as]);
      }
    }
  } else if (warp >= EPI_WARP0) {
    // ========================================================================= epilogue
    const int quarter = warp & 3;
    const int half = (warp - EPI_WARP0) >> 2;  // which 32-column chunks this warp owns (even / odd)
    const int row = quarter * 32 + lane;
    constexpr int CHUNKS = BN / 64;            // chunks per warp and tile
    uint32_t store_seq = 0;
    int it = 0;
    if constexpr (MODE != 2) {
      // ---- hot variants: TMEM -> registers -> bf16 -> 64B-swizzled staging box -> TMA store / reduce-add.
      // Two staging boxes per warp (the second set lives in the statistics region, unused here), the TMEM read of
      // chunk i+1 is in flight while chunk i is converted and staged (two register buffers, fully unrolled), and the
      // accumulator goes back to the MMA warp as soon as its last read has landed.
      float st[MODE == 1 ? CHUNKS : 1][4];     // MODE 1: running sum / sum of squares (lo, hi channel of this lane's word)
#pragma unroll
      for (int i = 0; i < (MODE == 1 ? CHUNKS : 1); ++i) st[i][0] = st[i][1] = st[i][2] = st[i][3] = 0.f;
      const int sh = lane & 1, sw = lane >> 1;  // statistics: row parity and bf16x2 word column of this lane
      for (int g = cluster_id; g < num_groups; g += num_clusters, ++it) {
        const int n_tile = g % p.num_n_tiles;
        const int m_tile = (g / p.num_n_tiles) * csize + crank;
        const int as = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        const bool valid = m_tile * BLOCK_M + row < p.M;
        mbar_wait(&acc_full[as], acc_phase);
        tc_fence_after();
        const uint32_t tacc = tmem_base + (uint32_t(quarter * 32) << 16) + as * BN + half * 32;
        uint32_t raw2[2][32];
        tmem_ld_32x32(tacc, raw2[0]);
#pragma unroll
        for (int ci = 0; ci < CHUNKS; ++ci) {
          const int n = n_tile * BN + (half + 2 * ci) * 32;
          uint32_t(&raw)[32] = raw2[ci & 1];
          uint4 rres[4];
          if constexpr (MODE == 3) {
            // the residual row segment of this chunk (64 contiguous bytes per lane) is requested before the TMEM read
            // is waited for, so both latencies overlap
            if (p.ep_res != nullptr && valid && n < p.cout_pad) {
              const uint4* rp =
                  reinterpret_cast<const uint4*>(p.ep_res + (long long)(m_tile * BLOCK_M + row) * p.ep_res_cs + n);
#pragma unroll
              for (int j = 0; j < 4; ++j) rres[j] = __ldg(rp + j);
            } else {
#pragma unroll
              for (int j = 0; j < 4; ++j) rres[j] = make_uint4(0u, 0u, 0u, 0u);
            }
          }
          tmem_ld_wait_for(raw);
          if (ci + 1 < CHUNKS) {
            tmem_ld_32x32(tacc + (ci + 1) * 64, raw2[(ci + 1) & 1]);
          } else {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_empty[as]);
          }
          if (n >= p.cout_pad) continue;  // warp-uniform
          if constexpr (MODE == 3) {
            const float4* sc4 = reinterpret_cast<const float4*>(p.ep_scale + n);
            const float4* sh4 = reinterpret_cast<const float4*>(p.ep_shift + n);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 a = __ldg(sc4 + j), b = __ldg(sh4 + j);  // same address in every lane: one broadcast each
              raw[4 * j + 0] = __float_as_uint(fmaf(__uint_as_float(raw[4 * j + 0]), a.x, b.x));
              raw[4 * j + 1] = __float_as_uint(fmaf(__uint_as_float(raw[4 * j + 1]), a.y, b.y));
              raw[4 * j + 2] = __float_as_uint(fmaf(__uint_as_float(raw[4 * j + 2]), a.z, b.z));
              raw[4 * j + 3] = __float_as_uint(fmaf(__uint_as_float(raw[4 * j + 3]), a.w, b.w));
            }
            if (p.ep_res != nullptr) {
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                raw[8 * j + 0] = __float_as_uint(__uint_as_float(raw[8 * j + 0]) + bf16_lo(rres[j].x));
                raw[8 * j + 1] = __float_as_uint(__uint_as_float(raw[8 * j + 1]) + bf16_hi(rres[j].x));
                raw[8 * j + 2] = __float_as_uint(__uint_as_float(raw[8 * j + 2]) + bf16_lo(rres[j].y));
                raw[8 * j + 3] = __float_as_uint(__uint_as_float(raw[8 * j + 3]) + bf16_hi(rres[j].y));
                raw[8 * j + 4] = __float_as_uint(__uint_as_float(raw[8 * j + 4]) + bf16_lo(rres[j].z));
                raw[8 * j + 5] = __float_as_uint(__uint_as_float(raw[8 * j + 5]) + bf16_hi(rres[j].z));
                raw[8 * j + 6] = __float_as_uint(__uint_as_float(raw[8 * j + 6]) + bf16_lo(rres[j].w));
                raw[8 * j + 7] = __float_as_uint(__uint_as_float(raw[8 * j + 7]) + bf16_hi(rres[j].w));
              }
            }
            if (p.ep_relu) {
#pragma unroll
              for (int j = 0; j < 32; ++j) raw[j] = __float_as_uint(fmaxf(__uint_as_float(raw[j]), 0.f));
            }
          }
          if (MODE == 1 && !valid) {
#pragma unroll
            for (int j = 0; j < 32; ++j) raw[j] = 0u;  // rows past M must not reach the batch statistics
          }
          uint8_t* stg = smem + ((store_seq & 1) ? L::STAT_OFFSET : L::STAGING_OFFSET) + (warp - EPI_WARP0) * (32 * 64);
          tma_store_wait_read1();  // the store issued two chunks ago (same box) has been read out
          ++store_seq;
          __syncwarp();
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            uint4 o;
            o.x = pack_bf16x2(__uint_as_float(raw[8 * j + 0]), __uint_as_float(raw[8 * j + 1]));
            o.y = pack_bf16x2(__uint_as_float(raw[8 * j + 2]), __uint_as_float(raw[8 * j + 3]));
            o.z = pack_bf16x2(__uint_as_float(raw[8 * j + 4]), __uint_as_float(raw[8 * j + 5]));
            o.w = pack_bf16x2(__uint_as_float(raw[8 * j + 6]), __uint_as_float(raw[8 * j + 7]));
            // SWIZZLE_64B: 16-byte chunk index ^= (byte address bits [7,9)) = (row >> 1) & 3
            *reinterpret_cast<uint4*>(stg + lane * 64 + ((j ^ ((lane >> 1) & 3)) << 4)) = o;
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            if (p.accumulate)  // residual join of the data gradient: y += tile, added in bf16 by the L2 (no SM read)
              tma_reduce_add_2d(&p.ymap, stg, n, m_tile * BLOCK_M + quarter * 32);
            else
              tma_store_2d(&p.ymap, stg, n, m_tile * BLOCK_M + quarter * 32);
            tma_store_commit();
          }
          if constexpr (MODE == 1) {
            // BatchNorm batch statistics straight from the staged tile (i.e. of the bf16 values the normalisation
            // pass will read): the box is [32 rows][16 bf16x2 words]; lane (sw, sh) walks the 16 rows of parity sh of
            // word column sw -- 16 conflict-free LDS.32 -- and keeps the sums in registers until the CTA is done.
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const uint32_t word = *reinterpret_cast<const uint32_t*>(stg + (2 * i + sh) * 64 +
                                                                       (((sw >> 2) ^ (i & 3)) << 4) + ((sw & 3) << 2));
              const float lo = bf16_lo(word), hi = bf16_hi(word);
              st[ci][0] += lo;
              st[ci][1] = fmaf(lo, lo, st[ci][1]);
              st[ci][2] += hi;
              st[ci][3] = fmaf(hi, hi, st[ci][3]);
            }
          }
        }
      }
      if (lane == 0) tma_store_wait_all();
      if constexpr (MODE == 1) {
        // every store of this CTA has completed once all epilogue warps are past this point: the staging region is
        // free and becomes the [2][BN] cross-warp reduction buffer (zeroed below, after the CTA-wide barrier)
        __syncwarp();
        tc_fence_before();
        __syncthreads();  // (A) matched by the non-epilogue warps below
        float* red = reinterpret_cast<float*>(smem + L::STAGING_OFFSET);
        for (int i = threadIdx.x - EPI_WARP0 * 32; i < 2 * BN; i += NUM_EPI_WARPS * 32) red[i] = 0.f;
        named_barrier_sync(1, NUM_EPI_WARPS * 32);
#pragma unroll
        for (int ci = 0; ci < CHUNKS; ++ci) {
          float v0 = st[ci][0], v1 = st[ci][1], v2 = st[ci][2], v3 = st[ci][3];
          v0 += __shfl_xor_sync(0xffffffffu, v0, 1);
          v1 += __shfl_xor_sync(0xffffffffu, v1, 1);
          v2 += __shfl_xor_sync(0xffffffffu, v2, 1);
          v3 += __shfl_xor_sync(0xffffffffu, v3, 1);
          const int col = (half + 2 * ci) * 32 + 2 * sw;
          // parity 0 owns the sums, parity 1 the squares
          atomicAdd(sh == 0 ? &red[col] : &red[BN + col], sh == 0 ? v0 : v1);
          atomicAdd(sh == 0 ? &red[col + 1] : &red[BN + col + 1], sh == 0 ? v2 : v3);
        }
        named_barrier_sync(1, NUM_EPI_WARPS * 32);
        if (it > 0) {  // this CTA processed tiles, all of n-tile (cluster_id % num_n_tiles)
          const int n0 = (cluster_id % p.num_n_tiles) * BN;
          for (int i = threadIdx.x - EPI_WARP0 * 32; i < BN; i += NUM_EPI_WARPS * 32) {
            const float b2 = red[BN + i];
            if (n0 + i < p.cout_pad && b2 != 0.f) {
              atomicAdd(p.stat_sum + n0 + i, (double)red[i]);
              atomicAdd(p.stat_sqsum + n0 + i, (double)b2);
            }
          }
        }
      }
    } else {
      // ---- generic variant
      for (int g = cluster_id; g < num_groups; g += num_clusters, ++it) {
        const int n_tile = g % p.num_n_tiles;
        const int m_tile = (g / p.num_n_tiles) * csize + crank;
        const int as = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        const int m = m_tile * BLOCK_M + row;
        const bool valid = m < p.M;
        long long pix = 0;
        if (valid) {
          const int img = m / p.HoWo;
          const int rem = m - img * p.HoWo;
          const int op = rem / p.Wo;
          const int oq = rem - op * p.Wo;
          pix = img * p.y_img + op * p.y_row + oq * p.y_pix;
        }
        mbar_wait(&acc_full[as], acc_phase);
        tc_fence_after();
#pragma unroll 1
        for (int chunk = half; chunk < BN / 32; chunk += 2) {
          const int n = n_tile * BN + chunk * 32;
          uint32_t raw[32];
          tmem_ld_32x32(tmem_base + (uint32_t(quarter * 32) << 16) + as * BN + chunk * 32, raw);
          tmem_ld_wait();
          if (n >= p.cout_pad) continue;  // warp-uniform
          float f[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(raw[j]);
          if (p.bias != nullptr) {
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] += __ldg(p.bias + n + j);
          }
          if (p.ep_scale != nullptr) {
            // eval-mode BatchNorm (+residual, +ReLU) folded into the epilogue: no separate normalisation pass
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] = fmaf(f[j], __ldg(p.ep_scale + n + j), __ldg(p.ep_shift + n + j));
            if (p.ep_res != nullptr && valid) {
              const uint4* rp = reinterpret_cast<const uint4*>(p.ep_res + pix * p.ep_res_cs + n);
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const uint4 o = __ldg(rp + j);
                f[8 * j + 0] += bf16_lo(o.x);
                f[8 * j + 1] += bf16_hi(o.x);
                f[8 * j + 2] += bf16_lo(o.y);
                f[8 * j + 3] += bf16_hi(o.y);
                f[8 * j + 4] += bf16_lo(o.z);
                f[8 * j + 5] += bf16_hi(o.z);
                f[8 * j + 6] += bf16_lo(o.w);
                f[8 * j + 7] += bf16_hi(o.w);
              }
            }
            if (p.ep_relu) {
#pragma unroll
              for (int j = 0; j < 32; ++j) f[j] = fmaxf(f[j], 0.f);
            }
          }
          if (p.stat_sum != nullptr && !valid) {
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] = 0.f;  // rows past M must not reach the batch statistics
          }
          if (p.use_tma_store) {
            // staged TMA store; a single box per warp here (the second set's region holds the statistics partials)
            uint8_t* stg = smem + L::STAGING_OFFSET + (warp - EPI_WARP0) * (32 * 64);
            tma_store_wait_read();
            __syncwarp();
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              uint4 o;
              o.x = pack_bf16x2(f[8 * j + 0], f[8 * j + 1]);
              o.y = pack_bf16x2(f[8 * j + 2], f[8 * j + 3]);
              o.z = pack_bf16x2(f[8 * j + 4], f[8 * j + 5]);
              o.w = pack_bf16x2(f[8 * j + 6], f[8 * j + 7]);
              *reinterpret_cast<uint4*>(stg + lane * 64 + ((j ^ ((lane >> 1) & 3)) << 4)) = o;
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
              if (p.accumulate)
                tma_reduce_add_2d(&p.ymap, stg, n, m_tile * BLOCK_M + quarter * 32);
              else
                tma_store_2d(&p.ymap, stg, n, m_tile * BLOCK_M + quarter * 32);
              tma_store_commit();
            }
          } else if (valid) {
            if (p.y_is_f32) {
              float4* dst = reinterpret_cast<float4*>(static_cast<float*>(p.y) + pix * p.y_cstride + n);
              if (p.accumulate) {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                  float4 o = dst[j];
                  f[4 * j + 0] += o.x;
                  f[4 * j + 1] += o.y;
                  f[4 * j + 2] += o.z;
                  f[4 * j + 3] += o.w;
                }
              }
#pragma unroll
              for (int j = 0; j < 8; ++j) dst[j] = make_float4(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]);
            } else {
              uint4* dst = reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(p.y) + pix * p.y_cstride + n);
              if (p.accumulate) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  uint4 o = dst[j];
                  f[8 * j + 0] += bf16_lo(o.x);
                  f[8 * j + 1] += bf16_hi(o.x);
                  f[8 * j + 2] += bf16_lo(o.y);
                  f[8 * j + 3] += bf16_hi(o.y);
                  f[8 * j + 4] += bf16_lo(o.z);
                  f[8 * j + 5] += bf16_hi(o.z);
                  f[8 * j + 6] += bf16_lo(o.w);
                  f[8 * j + 7] += bf16_hi(o.w);
                }
              }
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                uint4 o;
                o.x = pack_bf16x2(f[8 * j + 0], f[8 * j + 1]);
                o.y = pack_bf16x2(f[8 * j + 2], f[8 * j + 3]);
                o.z = pack_bf16x2(f[8 * j + 4], f[8 * j + 5]);
                o.w = pack_bf16x2(f[8 * j + 6], f[8 * j + 7]);
                dst[j] = o;
              }
            }
          }
          if (p.stat_sum != nullptr) {
            // batch statistics of the fp32 conv output (F.batch_norm training path,
            // zs3/modeling/sync_batchnorm/batchnorm.py:48-58): per-channel sum and sum of squares, shuffle-tree
            // transpose-reduce per 32 columns -> per-CTA shared-memory partials
            float sq[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) sq[j] = f[j] * f[j];
            const float s1 = warp_column_sums(f, lane);
            const float s2 = warp_column_sums(sq, lane);
            atomicAdd(&s_sum[n + lane], s1);
            atomicAdd(&s_sq[n + lane], s2);
          }
        }
        // all tcgen05.ld of this warp have completed (wait::ld above): release the accumulator
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&acc_empty[as]);
      }
      if (p.use_tma_store && lane == 0) tma_store_wait_all();
    }
  }

  tc_fence_before();
  if (MODE != 1 || warp < EPI_WARP0) __syncthreads();  // MODE 1: the epilogue warps arrived at (A) above
  if (csize > 1) cluster_sync_all();  // no CTA may exit while a peer can still signal its barriers
  tc_fence_after();
  if (warp == 2) tmem_dealloc(tmem_base, 2 * BN);
  if (MODE == 2 && p.stat_sum != nullptr) {
    for (int i = threadIdx.x; i < p.cout_pad; i += NUM_THREADS) {
      const float a = s_sum[i], b = s_sq[i];
      if (b != 0.f) {  // channels of n-tiles this CTA never visited (and all-zero channels) add nothing
        atomicAdd(p.stat_sum + i, (double)a);
        atomicAdd(p.stat_sqsum + i, (double)b);
      }
    }
  }
}

// ------------------------------------------------------------------------------------ wgrad
// dW[co][tap][ci] += sum_p dY[p][co] * X[p@tap][ci]; both operands are MN-major for the MMA
// (the reduction index p runs over 128-byte rows of the TMA tiles).
constexpr int WG_BLOCK_P = 64;  // pixels (reduction) per k-block
constexpr int WG_CHUNK_BYTES = WG_BLOCK_P * 128;  // one [64 pixels][64 channels] bf16 sub-tile = 8 KiB

struct WgradParams {
  CUtensorMap dy;  // tiled 2-D over [M][dy_cstride], box [64 pixels][64 channels]
  CUtensorMap x;   // im2col over the input activation, 64 pixels x 64 channels
  int M, HoWo, Wo;
  int stride, pad, dil, R, S;
  int cout_pad, cin_pad;
  int num_co_tiles, num_ci_tiles;
  int k_splits;
  int pblocks_total;  // ceil(M / 64)
  float* dw;
  long long dw_ld;       // elements between consecutive (co, tap) rows of dw
  int dw_ci_offset;      // first input channel of this segment inside a dw row
  int cout_valid, cin_valid;  // logical extents (rows / columns beyond them are not written)
  CUtensorMap dwmap;          // fp32 3-D map (ci, tap, co) over the valid region of dw; box 32 ci x 1 x 32 co
  int use_tma_reduce;
};

// MB = number of 128-row output-channel blocks per CTA (1 or 2): MB = 2 reuses every X tile for 256 output
// channels, raising the arithmetic intensity of the L2->SM stream from 85 to 128 FLOP/B.
template <int CN, int STAGES, int MB>
struct WgradSmem {
  static constexpr int A_BYTES = MB * 2 * WG_CHUNK_BYTES;     // MB * 128 output channels
  static constexpr int B_BYTES = (CN / 64) * WG_CHUNK_BYTES;  // CN input channels
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int BAR_OFFSET = STAGES * STAGE_BYTES;
  static constexpr int TOTAL = BAR_OFFSET + 256 + 1024;
};

template <int CN, int STAGES, int MB>
__global__ void __launch_bounds__(WG_THREADS, 1) conv_wgrad_kernel(const __grid_constant__ WgradParams p) {
  using L = WgradSmem<CN, STAGES, MB>;
  constexpr int TMEM_COLS = (MB * CN) < 32 ? 32 : (MB * CN);
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::BAR_OFFSET);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* acc_full = empty_bar + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  // blockIdx.x -> (ci_tile, co_tile, tap); blockIdx.y -> pixel split
  int t = blockIdx.x;
  const int ci_tile = t % p.num_ci_tiles;
  t /= p.num_ci_tiles;
  const int co_tile = t % p.num_co_tiles;
  const int tap = t / p.num_co_tiles;
  const int tap_r = tap / p.S;
  const int tap_s = tap - tap_r * p.S;
  const int pb_per_split = (p.pblocks_total + p.k_splits - 1) / p.k_splits;
  const int pb_begin = blockIdx.y * pb_per_split;
  const int pb_end = min(p.pblocks_total, pb_begin + pb_per_split);
  const int num_kb = pb_end - pb_begin;

  if (threadIdx.x == 0) {
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    mbar_init(acc_full, 1);
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  griddep_sync();

  if (num_kb > 0) {
    if (warp == 0) {
      if (elect_one()) {
        int stage = 0;
        uint32_t phase = 0;
        for (int kb = 0; kb < num_kb; ++kb) {
          const int p0 = (pb_begin + kb) * WG_BLOCK_P;
          const int img = p0 / p.HoWo;
          const int rem = p0 - img * p.HoWo;
          const int op = rem / p.Wo;
          const int oq = rem - op * p.Wo;
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * L::STAGE_BYTES;
          uint8_t* sb = sa + L::A_BYTES;
          mbar_arrive_expect_tx(&full_bar[stage], L::STAGE_BYTES);
#pragma unroll
          for (int c = 0; c < 2 * MB; ++c)
            tma_load_2d(sa + c * WG_CHUNK_BYTES, &p.dy, &full_bar[stage], co_tile * (128 * MB) + c * 64, p0);
#pragma unroll
          for (int c = 0; c < CN / 64; ++c)
            tma_load_im2col_4d(sb + c * WG_CHUNK_BYTES, &p.x, &full_bar[stage], ci_tile * CN + c * 64,
                               oq * p.stride - p.pad, op * p.stride - p.pad, img, (uint16_t)(tap_s * p.dil),
                               (uint16_t)(tap_r * p.dil));
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    } else if (warp == 1) {
      if (elect_one()) {
        constexpr uint32_t idesc = make_idesc_bf16(128, CN, 1, 1);
        int stage = 0;
        uint32_t phase = 0;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * L::STAGE_BYTES);
          const uint32_t sb = sa + L::A_BYTES;
#pragma unroll
          for (int k = 0; k < WG_BLOCK_P / 16; ++k) {
            // MN-major SW128: LBO = byte distance between 64-channel sub-tiles, SBO = 8 pixel rows = 1024 B;
            // one UMMA_K step = 16 pixel rows = 2048 B.
            const uint64_t bdesc = make_smem_desc_sw128(sb + k * 2048, WG_CHUNK_BYTES, 1024);
#pragma unroll
            for (int mb = 0; mb < MB; ++mb) {
              const uint64_t adesc = make_smem_desc_sw128(sa + mb * 2 * WG_CHUNK_BYTES + k * 2048, WG_CHUNK_BYTES, 1024);
              umma_f16(tmem_base + mb * CN, adesc, bdesc, idesc, (kb | k) != 0);
            }
          }
          umma_commit(&empty_bar[stage]);
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit(acc_full);
      }
    } else if (warp >= EPI_WARP0) {
      const int quarter = warp & 3;
      mbar_wait(acc_full, 0);
      tc_fence_after();
      const int taps = p.R * p.S;
      const bool vec_ok = (p.dw_ld % 4 == 0) && (p.dw_ci_offset % 4 == 0) &&
                          ((reinterpret_cast<uintptr_t>(p.dw) & 15) == 0);
      // TMA-reduce path: the main loop is over, so the pipeline stages are free: two 4 KiB fp32 boxes per warp
      // two epilogue warps per 32-row quarter take alternate 32-column chunks; the TMEM read of the next chunk is in
      // flight while the current one is staged and reduced (two register buffers, fully unrolled loop) -- this
      // epilogue is NOT overlapped with a main loop (one tile per CTA), so its latency is on the critical path
      uint8_t* stg_base = smem + (warp - EPI_WARP0) * 8192;
      uint32_t seq = 0;
      const int half = (warp - EPI_WARP0) >> 2;
      constexpr int CHUNKS = MB * CN / 64;  // per warp
      const uint32_t tacc = tmem_base + (uint32_t(quarter * 32) << 16) + half * 32;
      uint32_t raw2[2][32];
      tmem_ld_32x32(tacc, raw2[0]);
#pragma unroll
      for (int it = 0; it < CHUNKS; ++it) {
        const int chunk = half + 2 * it;
        uint32_t(&raw)[32] = raw2[it & 1];
        tmem_ld_wait_for(raw);
        if (it + 1 < CHUNKS) tmem_ld_32x32(tacc + (it + 1) * 64, raw2[(it + 1) & 1]);
        const int mb = chunk / (CN / 32);
        const int co = co_tile * (128 * MB) + mb * 128 + quarter * 32 + lane;
        const int ci = ci_tile * CN + (chunk - mb * (CN / 32)) * 32;
        if (p.use_tma_reduce) {
          // [32 co rows][32 ci] fp32 box, 128-byte rows with the 128B swizzle (16-byte chunk ^= row & 7), then ONE
          // cp.reduce.async.bulk (.add): rows/columns outside the valid extents are clipped by the tensor map
          uint8_t* stg = stg_base + (seq & 1) * 4096;
          tma_store_wait_read1();
          ++seq;
          __syncwarp();
#pragma unroll
          for (int j = 0; j < 8; ++j)
            *reinterpret_cast<uint4*>(stg + lane * 128 + ((j ^ (lane & 7)) << 4)) =
                make_uint4(raw[4 * j], raw[4 * j + 1], raw[4 * j + 2], raw[4 * j + 3]);
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            tma_reduce_add_3d(&p.dwmap, stg, ci, tap, co_tile * (128 * MB) + mb * 128 + quarter * 32);
            tma_store_commit();
          }
        } else if (co < p.cout_valid && ci < p.cin_valid) {
          float* row = p.dw + ((long long)co * taps + tap) * p.dw_ld + p.dw_ci_offset + ci;
          if (vec_ok && ci + 32 <= p.cin_valid) {
            float4* dst = reinterpret_cast<float4*>(row);
#pragma unroll
            for (int j = 0; j < 8; ++j)  // 16-byte vector reductions (sm_90+): 8 instead of 32 atomics per row chunk
              atomicAdd(dst + j, make_float4(__uint_as_float(raw[4 * j]), __uint_as_float(raw[4 * j + 1]),
                                             __uint_as_float(raw[4 * j + 2]), __uint_as_float(raw[4 * j + 3])));
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (ci + j < p.cin_valid) atomicAdd(row + j, __uint_as_float(raw[j]));
          }
        }
      }
      if (p.use_tma_reduce && lane == 0) tma_store_wait_all();
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 2) tmem_dealloc(tmem_base, TMEM_COLS);
}

// ---------------------------------------------------------------------------- weight packing
// Tiled through shared memory so that both the fp32 OIHW reads and the bf16 packed writes are contiguous runs:
// one CTA repacks a 32 (co) x 32 (ci) x taps block.
constexpr int PK_T = 32;
__global__ void __launch_bounds__(256) pack_weight_kernel(const float* __restrict__ w, int Cout, int Cin, int R, int S,
                                                          int ci_begin, int ci_count, __nv_bfloat16* __restrict__ dst,
                                                          int cout_pad, int cin_pad, int mode, int src_krsc) {
  extern __shared__ float tile[];  // [PK_T co][PK_T ci][taps] (+1 padding per co row)
  const int taps = R * S;
  const int ld = PK_T * taps + 1;
  const int co0 = blockIdx.y * PK_T, ci0 = blockIdx.x * PK_T;
  if (!src_krsc) {
    // OIHW source: for each co a contiguous run of (PK_T ci x taps) floats
    for (int i = threadIdx.x; i < PK_T * PK_T * taps; i += blockDim.x) {
      const int co = i / (PK_T * taps), rem = i - co * (PK_T * taps);
      const int ci = rem / taps;
      float v = 0.f;
      if (co0 + co < Cout && ci0 + ci < ci_count)
        v = w[((long long)(co0 + co) * Cin + ci_begin + ci0) * taps + rem];
      tile[co * ld + rem] = v;
    }
  } else {
    // KRSC (torch channels_last) source: for each (co, tap) a contiguous run of PK_T input channels
    for (int i = threadIdx.x; i < PK_T * taps * PK_T; i += blockDim.x) {
      const int ci = i % PK_T, tap = (i / PK_T) % taps, co = i / (PK_T * taps);
      float v = 0.f;
      if (co0 + co < Cout && ci0 + ci < ci_count)
        v = w[((long long)(co0 + co) * taps + tap) * Cin + ci_begin + ci0 + ci];
      tile[co * ld + ci * taps + tap] = v;
    }
  }
  __syncthreads();
  if (mode == 0) {  // dst[co][tap][ci]
    for (int i = threadIdx.x; i < PK_T * taps * PK_T; i += blockDim.x) {
      const int ci = i % PK_T, tap = (i / PK_T) % taps, co = i / (PK_T * taps);
      if (co0 + co < cout_pad && ci0 + ci < cin_pad)
        dst[((long long)(co0 + co) * taps + tap) * cin_pad + ci0 + ci] = __float2bfloat16(tile[co * ld + ci * taps + tap]);
    }
  } else {  // dst[ci][taps-1-tap][co]
    for (int i = threadIdx.x; i < PK_T * taps * PK_T; i += blockDim.x) {
      const int co = i % PK_T, tap = (i / PK_T) % taps, ci = i / (PK_T * taps);
      if (co0 + co < cout_pad && ci0 + ci < cin_pad)
        dst[((long long)(ci0 + ci) * taps + (taps - 1 - tap)) * cout_pad + co0 + co] =
            __float2bfloat16(tile[co * ld + ci * taps + tap]);
    }
  }
}

// KRSC (channels_last) sources are one cast / one 2-D transpose per tap away from the kernel layouts:
// mode 0: dst[co][tap][ci] = src[co][tap][ci_begin + ci]   (contiguous runs on both sides)
__global__ void __launch_bounds__(256) pack_krsc_fprop_kernel(const float* __restrict__ w, int Cout, int Cin, int taps,
                                                              int ci_begin, int ci_count,
                                                              __nv_bfloat16* __restrict__ dst, int cout_pad,
                                                              int cin_pad) {
  const int vpr = cin_pad >> 2;  // 4-element vectors per (co, tap) row
  const long long total = (long long)cout_pad * taps * vpr;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int ci = (int)(i % vpr) << 2;
    const long long row = i / vpr;  // co * taps + tap
    const int co = (int)(row / taps);
    float v[4] = {0.f, 0.f, 0.f, 0.f};
    if (co < Cout) {
      const float* src = w + row * Cin + ci_begin + ci;
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (ci + j < ci_count) v[j] = src[j];
    }
    uint2 o;
    o.x = pack_bf16x2(v[0], v[1]);
    o.y = pack_bf16x2(v[2], v[3]);
    *reinterpret_cast<uint2*>(dst + row * cin_pad + ci) = o;
  }
}

// mode 1: dst[ci][taps-1-tap][co] = src[co][tap][ci_begin + ci]: a 32x32 shared-memory transpose per tap
__global__ void __launch_bounds__(256) pack_krsc_dgrad_kernel(const float* __restrict__ w, int Cout, int Cin, int taps,
                                                              int ci_begin, int ci_count,
                                                              __nv_bfloat16* __restrict__ dst, int cout_pad,
                                                              int cin_pad) {
  __shared__ float tile[32][33];
  const int tap = blockIdx.z, co0 = blockIdx.y * 32, ci0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  for (int r = ty; r < 32; r += 8) {
    const int co = co0 + r, ci = ci0 + tx;
    tile[r][tx] = (co < Cout && ci < ci_count) ? w[((long long)co * taps + tap) * Cin + ci_begin + ci] : 0.f;
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    const int ci = ci0 + r, co = co0 + tx;
    if (ci < cin_pad && co < cout_pad)
      dst[((long long)ci * taps + (taps - 1 - tap)) * cout_pad + co] = __float2bfloat16(tile[tx][r]);
  }
}

// grad_oihw[co][ci_begin+ci][tap] (+)= dw[co][tap][ci]: same tiling, reversed direction
__global__ void __launch_bounds__(256) unpack_wgrad_kernel(const float* __restrict__ dw, int cout_pad, int cin_pad,
                                                           float* __restrict__ g, int Cout, int Cin, int R, int S,
                                                           int ci_begin, int ci_count, int accumulate) {
  extern __shared__ float tile[];
  const int taps = R * S;
  const int ld = PK_T * taps + 1;
  const int co0 = blockIdx.y * PK_T, ci0 = blockIdx.x * PK_T;
  for (int i = threadIdx.x; i < PK_T * taps * PK_T; i += blockDim.x) {
    const int ci = i % PK_T, tap = (i / PK_T) % taps, co = i / (PK_T * taps);
    float v = 0.f;
    if (co0 + co < Cout && ci0 + ci < ci_count) v = dw[((long long)(co0 + co) * taps + tap) * cin_pad + ci0 + ci];
    tile[co * ld + ci * taps + tap] = v;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < PK_T * PK_T * taps; i += blockDim.x) {
    const int co = i / (PK_T * taps), rem = i - co * (PK_T * taps);
    const int ci = rem / taps;
    if (co0 + co < Cout && ci0 + ci < ci_count) {
      float* d = g + ((long long)(co0 + co) * Cin + ci_begin + ci0) * taps + rem;
      *d = accumulate ? (*d + tile[co * ld + rem]) : tile[co * ld + rem];
    }
  }
}

static int g_num_sms = 0;
static int num_sms() {
  if (g_num_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (g_num_sms <= 0) g_num_sms = 148;
  }
  return g_num_sms;
}

template <int BN, int STAGES, int MODE>
static int launch_fprop_mode(const FpropParams& p, int csize, int clusters, cudaStream_t st);

// ZS3_FOLD_HOT=0 sends the folded inference epilogue back to the generic variant (A/B knob, profiles/r02_*)
static bool fold_hot_enabled() {
  static int pref = -1;
  if (pref < 0) {
    const char* env = getenv("ZS3_FOLD_HOT");
    pref = env ? atoi(env) : 1;
  }
  return pref != 0;
}

// epilogue variant + grid: see conv_fprop_kernel.  The register-statistics variant needs every CTA to stay on one
// n-tile, i.e. a grid that is a multiple of num_n_tiles (148 already is for 1, 2 and 4 n-tiles; 8 n-tiles run on 144).
template <int BN, int STAGES>
static int launch_fprop(const FpropParams& p, int csize, cudaStream_t st) {
  const int groups = p.num_m_groups * p.num_n_tiles;
  const int max_clusters = num_sms() / csize;
  int clusters = groups < max_clusters ? groups : max_clusters;
  // the folded inference epilogue on the hot path: dense bf16 output, residual rows addressed by the output pixel index
  if (p.use_tma_store && p.bias == nullptr && p.ep_scale != nullptr && p.stat_sum == nullptr && !p.accumulate &&
      fold_hot_enabled())
    return launch_fprop_mode<BN, STAGES, 3>(p, csize, clusters, st);
  const bool hot = p.use_tma_store && p.bias == nullptr && p.ep_scale == nullptr;
  if (!hot) return launch_fprop_mode<BN, STAGES, 2>(p, csize, clusters, st);
  if (p.stat_sum == nullptr) return launch_fprop_mode<BN, STAGES, 0>(p, csize, clusters, st);
  if (csize == 1 && clusters >= p.num_n_tiles) {
    clusters -= clusters % p.num_n_tiles;
    return launch_fprop_mode<BN, STAGES, 1>(p, csize, clusters, st);
  }
  return launch_fprop_mode<BN, STAGES, 2>(p, csize, clusters, st);
}

template <int BN, int STAGES, int MODE>
static int launch_fprop_mode(const FpropParams& p, int csize, int clusters, cudaStream_t st) {
  using L = FpropSmem<BN, STAGES>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(conv_fprop_kernel<BN, STAGES, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         L::TOTAL);
    if (e != cudaSuccess) {
      set_error("conv_fprop: cudaFuncSetAttribute(%d bytes) failed: %s", L::TOTAL, cudaGetErrorString(e));
      return ZS3_ERR_LAUNCH;
    }
    configured = true;
  }
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(clusters * csize);
  cfg.blockDim = dim3(NUM_THREADS);
  cfg.dynamicSmemBytes = L::TOTAL;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = csize;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = (csize == 1 && pdl_enabled(PDL_CONV)) ? 2 : 1;  // PDL only for the (default) non-cluster launch
  cudaError_t e = cudaLaunchKernelEx(&cfg, conv_fprop_kernel<BN, STAGES, MODE>, p);
  if (e != cudaSuccess) {
    set_error("conv_fprop: cudaLaunchKernelEx(cluster=%d) failed: %s", csize, cudaGetErrorString(e));
    return ZS3_ERR_LAUNCH;
  }
  ZS3_CHECK_LAUNCH("conv_fprop");
  return ZS3_OK;
}

template <int CN, int STAGES, int MB>
static int launch_wgrad(const WgradParams& p, cudaStream_t st) {
  using L = WgradSmem<CN, STAGES, MB>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(conv_wgrad_kernel<CN, STAGES, MB>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         L::TOTAL);
    if (e != cudaSuccess) {
      set_error("conv_wgrad: cudaFuncSetAttribute(%d bytes) failed: %s", L::TOTAL, cudaGetErrorString(e));
      return ZS3_ERR_LAUNCH;
    }
    configured = true;
  }
  dim3 grid(p.num_ci_tiles * p.num_co_tiles * p.R * p.S, p.k_splits);
  cudaError_t le = launch_pdl(PDL_CONV, conv_wgrad_kernel<CN, STAGES, MB>, grid, dim3(WG_THREADS), L::TOTAL, st, p);
  if (le != cudaSuccess) {
    set_error("conv_wgrad: launch failed: %s", cudaGetErrorString(le));
    return ZS3_ERR_LAUNCH;
  }
  ZS3_CHECK_LAUNCH("conv_wgrad");
  return ZS3_OK;
}

}  // namespace zs3

using namespace zs3;

extern "C" int zs3_conv_fprop(const zs3_conv_args* a, void* stream) {
  ZS3_CHECK_ARG(a != nullptr, "conv_fprop: null args");
  ZS3_CHECK_ARG(a->num_segments >= 1 && a->num_segments <= ZS3_MAX_SEGMENTS, "conv_fprop: num_segments=%d",
                a->num_segments);
  ZS3_CHECK_ARG(a->cout_pad > 0 && a->cout_pad % 64 == 0, "conv_fprop: cout_pad=%d must be a multiple of 64",
                a->cout_pad);
  ZS3_CHECK_ARG(a->N > 0 && a->H > 0 && a->W > 0 && a->Ho > 0 && a->Wo > 0, "conv_fprop: bad spatial dims");
  ZS3_CHECK_ARG(a->R >= 1 && a->S >= 1 && a->stride >= 1 && a->stride <= 8 && a->dil >= 1 && a->pad >= 0,
                "conv_fprop: bad filter geometry");
  ZS3_CHECK_ARG(a->pad <= 128 && (a->R - 1) * a->dil <= 255 && (a->S - 1) * a->dil <= 255,
                "conv_fprop: pad/dilation outside the TMA im2col range");
  ZS3_CHECK_ARG(a->y != nullptr && a->y_cstride >= a->cout_pad, "conv_fprop: bad output");
  ZS3_CHECK_ARG((a->stat_sum == nullptr) == (a->stat_sqsum == nullptr), "conv_fprop: stat_sum/stat_sqsum mismatch");
  ZS3_CHECK_ARG(a->stat_sum == nullptr || a->cout_pad <= MAX_STAT_CH, "conv_fprop: statistics need cout_pad <= %d",
                MAX_STAT_CH);
  const long long M = (long long)a->N * a->Ho * a->Wo;
  ZS3_CHECK_ARG(M < (1ll << 31), "conv_fprop: too many output pixels");
  for (int s = 0; s < ZS3_MAX_SEGMENTS; ++s)
    ZS3_CHECK_ARG(a->pre_scale[s] == nullptr && a->pre_shift[s] == nullptr && a->pre_relu == 0,
                  "conv_fprop: the operand prologue (pre_scale / pre_shift / pre_relu) is reserved and not implemented");

  FpropParams p;
  memset(&p, 0, sizeof(p));
  int BN = a->cout_pad >= 256 ? 256 : (a->cout_pad >= 128 ? 128 : 64);
  {
    // A 256-channel output whose 128-row tiles fit in ONE wave (layer3's 33x33 maps: 137 tiles on 148 SMs) gives every
    // CTA a single tile: prologue, pipeline fill and the whole epilogue are exposed.  With 128-column tiles each CTA
    // runs two half-width tiles and the epilogue of the first overlaps the main loop of the second (the A tile is
    // read twice, from L2).  ZS3_FPROP_SPLIT_N=0/1; see profiles/ for the measurement behind the default.
    static int split_n = -1;
    if (split_n < 0) {
      const char* env = getenv("ZS3_FPROP_SPLIT_N");
      split_n = env ? atoi(env) : 0;
    }
    if (split_n && a->cout_pad == 256 && ceil_div_ll(M, BLOCK_M) <= num_sms()) BN = 128;
  }
  // cluster size: CTAs of a cluster share (multicast) the weight tile.  Measured on B200 (profiles/r01_cluster_*):
  // multicast at cluster sizes <= 4 does not reduce L2->SM traffic and the lock step costs 8-60 %, so the default is
  // 1; ZS3_CLUSTER=2|4 keeps the path testable.
  static int cluster_pref = -1;
  if (cluster_pref < 0) {
    const char* env = getenv("ZS3_CLUSTER");
    cluster_pref = env ? atoi(env) : 1;
    if (cluster_pref != 1 && cluster_pref != 2 && cluster_pref != 4) cluster_pref = 1;
  }
  const int m_tiles_total = (int)ceil_div_ll(M, BLOCK_M);
  int csize = a->w_forward_layout ? 1 : cluster_pref;
  while (csize > 1 && m_tiles_total < 2 * csize) csize >>= 1;
  int kb = 0;
  for (int s = 0; s < a->num_segments; ++s) {
    const zs3_conv_segment& sg = a->seg[s];
    ZS3_CHECK_ARG(sg.x != nullptr && sg.w != nullptr, "conv_fprop: segment %d null pointer", s);
    ZS3_CHECK_ARG(sg.cin_pad > 0 && sg.cin_pad % 64 == 0 && sg.x_cstride >= sg.cin_pad && sg.x_cstride % 8 == 0,
                  "conv_fprop: segment %d cin_pad=%d x_cstride=%d", s, sg.cin_pad, sg.x_cstride);
    // bounding box of filter base positions: [-pad, W-1 + pad - (S-1)*dil]
    const int upper_w = a->pad - (a->S - 1) * a->dil;
    const int upper_h = a->pad - (a->R - 1) * a->dil;
    ZS3_CHECK_ARG(upper_w == upper_h || a->H == a->W, "conv_fprop: non-square filters need square geometry");
    int rc = encode_im2col_bf16(&p.seg[s].a, sg.x, a->N, a->H, a->W, sg.x_cstride, a->pad, upper_w, a->stride, BLOCK_K,
                                BLOCK_M);
    if (rc) return rc;
    if (a->w_forward_layout)  // w[k = sg.cin_pad][taps][n = cout_pad]
      rc = encode_tiled3d_bf16(&p.seg[s].b, sg.w, sg.cin_pad, a->R * a->S, a->cout_pad, 64, 1, 64);
    else
      rc = encode_tiled3d_bf16(&p.seg[s].b, sg.w, a->cout_pad, a->R * a->S, sg.cin_pad, BN / csize, 1, BLOCK_K);
    if (rc) return rc;
    p.seg[s].num_cblk = sg.cin_pad / BLOCK_K;
    kb += a->R * a->S * p.seg[s].num_cblk;
  }
  p.num_segments = a->num_segments;
  p.M = (int)M;
  p.HoWo = a->Ho * a->Wo;
  p.Wo = a->Wo;
  p.stride = a->stride;
  p.pad = a->pad;
  p.dil = a->dil;
  p.R = a->R;
  p.S = a->S;
  p.cout_pad = a->cout_pad;
  p.num_m_tiles = m_tiles_total;
  p.num_m_groups = ceil_div(m_tiles_total, csize);
  p.num_n_tiles = ceil_div(a->cout_pad, BN);
  p.kb_per_tile = kb;
  p.kb_per_row = kb / a->R;
  p.H = a->H;
  p.Ho = a->Ho;
  {
    // ZS3_TAP_CULL=0 multiplies the zero padding like any other tap (A/B knob, profiles/r02_tap_culling.md)
    static int cull_pref = -1;
    if (cull_pref < 0) {
      const char* env = getenv("ZS3_TAP_CULL");
      cull_pref = env ? atoi(env) : 1;
    }
    // clusters run their CTAs in lock step on a shared weight stage: every CTA must fetch the same k-blocks.
    // Only worth the per-tile index arithmetic when a noticeable share of the (output row, filter row) pairs reads
    // padding: dilated 3x3 at 33x33 (layer4 d >= 4, ASPP) qualify, the pad-1 convs do not.
    p.cull = 0;
    if (cull_pref && a->R > 1 && a->R <= 16 && csize == 1) {
      long long skippable = 0;
      for (int r = 0; r < a->R; ++r) {
        const int t = a->pad - r * a->dil;
        const int lo = t > 0 ? (t + a->stride - 1) / a->stride : 0;
        const int u = a->H - 1 + t;
        int hi = u >= 0 ? u / a->stride : -1;
        if (hi > a->Ho - 1) hi = a->Ho - 1;
        const int valid = hi >= lo ? hi - lo + 1 : 0;
        skippable += a->Ho - (valid < a->Ho ? valid : a->Ho);
      }
      p.cull = (cull_pref >= 2 || skippable * 20 >= (long long)a->R * a->Ho) ? 1 : 0;  // >= 5 % (2 = always, for tests)
    }
  }
  p.y = a->y;
  p.y_cstride = a->y_cstride;
  if (a->y_sp_stride > 1) {
    ZS3_CHECK_ARG(a->y_H >= (a->Ho - 1) * a->y_sp_stride + 1 && a->y_W >= (a->Wo - 1) * a->y_sp_stride + 1,
                  "conv_fprop: scatter target %dx%d too small", a->y_H, a->y_W);
    p.y_img = (long long)a->y_H * a->y_W;
    p.y_row = (long long)a->y_W * a->y_sp_stride;
    p.y_pix = a->y_sp_stride;
  } else {
    p.y_img = (long long)a->Ho * a->Wo;
    p.y_row = a->Wo;
    p.y_pix = 1;
  }
  p.y_is_f32 = a->y_is_f32;
  p.accumulate = a->accumulate;
  p.b_mn = a->w_forward_layout ? 1 : 0;
  p.use_tma_store = 0;
  static int tma_store_pref = -1;
  if (tma_store_pref < 0) {
    const char* env = getenv("ZS3_TMA_STORE");
    tma_store_pref = env ? atoi(env) : 1;
  }
  static int tma_acc_pref = -1;
  if (tma_acc_pref < 0) {
    const char* env = getenv("ZS3_TMA_ACCUMULATE");
    tma_acc_pref = env ? atoi(env) : 1;
  }
  if (tma_store_pref && !a->y_is_f32 && (!a->accumulate || tma_acc_pref) && a->y_sp_stride <= 1 && a->y_cstride % 8 == 0 &&
      (reinterpret_cast<uintptr_t>(a->y) & 15) == 0) {
    int rc2 = encode_tiled2d_bf16_sw64(&p.ymap, a->y, M, a->cout_pad, a->y_cstride, 32, 32);
    if (rc2) return rc2;
    p.use_tma_store = 1;
  }
  p.bias = a->bias;
  p.stat_sum = a->stat_sum;
  p.stat_sqsum = a->stat_sqsum;
  ZS3_CHECK_ARG((a->ep_scale == nullptr) == (a->ep_shift == nullptr), "conv_fprop: ep_scale/ep_shift mismatch");
  ZS3_CHECK_ARG(a->ep_residual == nullptr || (a->ep_scale != nullptr && a->ep_res_cstride >= a->cout_pad &&
                                              a->ep_res_cstride % 8 == 0 && a->y_sp_stride <= 1),
                "conv_fprop: bad epilogue residual");
  p.ep_scale = a->ep_scale;
  p.ep_shift = a->ep_shift;
  p.ep_res = static_cast<const __nv_bfloat16*>(a->ep_residual);
  p.ep_res_cs = a->ep_res_cstride;
  p.ep_relu = a->ep_relu;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (BN == 256) return launch_fprop<256, 4>(p, csize, st);
  if (BN == 128) return launch_fprop<128, 6>(p, csize, st);
  return launch_fprop<64, 8>(p, csize, st);
}

// Output-tile shape of the weight gradient.  The grid is (output tiles) x (pixel splits), one wave of CTAs; every CTA
// ends with an fp32 reduce-add of its whole tile into dW, so the split-K reduction moves splits x |dW| bytes through
// the L2 atomic units.  On the 33x33 layers (17 k pixels, 1 MB of dW) the largest tile (256 x 256) leaves 4-9 output
// tiles and therefore 16-37 splits: 37 MB of reduce traffic per launch (22-29 us measured, profiles/r02_layer_table.md).
// Smaller tiles re-read the operands more often but cut the splits; the choice is made by a three-term model (operand
// bytes per CTA at the L2->SM rate, MMA time, total reduce bytes at the L2 reduce rate).  In practice it moves the
// 1x1 1024<->256 layers of layer3 to 128 x 128 tiles (9 splits instead of 37) and leaves everything else on the
// largest tile.  ZS3_WGRAD_TILE=<MB>x<CN> pins the tile (experiments).
static void choose_wgrad_tile(long long M, int cout_pad, int cin_pad, int taps, int k_splits_arg, int* MB, int* CN) {
  static int pin_mb = -1, pin_cn = 0;
  if (pin_mb < 0) {
    pin_mb = 0;
    const char* env = getenv("ZS3_WGRAD_TILE");
    if (env) {
      int mb = 0, cn = 0;
      if (sscanf(env, "%dx%d", &mb, &cn) == 2 && (mb == 1 || mb == 2) && (cn == 64 || cn == 128 || cn == 256)) {
        pin_mb = mb;
        pin_cn = cn;
      }
    }
  }
  const int max_cn = cin_pad >= 256 ? 256 : (cin_pad >= 128 ? 128 : 64);
  const int max_mb = cout_pad >= 256 ? 2 : 1;
  if (pin_mb > 0) {
    *MB = pin_mb <= max_mb ? pin_mb : max_mb;
    *CN = pin_cn <= max_cn ? pin_cn : max_cn;
    return;
  }
  // rates fitted to four B200 measurements (3x3 256->256 and 1x1 1024->256 at 33x33, tiles 256x256 and 128x128,
  // profiles/r02_wgrad_tile.md): T = operand bytes per CTA / kLoad + total reduce bytes / kReduce + const
  const double kLoadBytesPerUs = 97e3;    // L2 -> one SM, TMA operand stream
  const double kMacPerUs = 4096.0 * 1900; // one SM, dense bf16
  const double kReduceBytesPerUs = 4.0e6; // all SMs -> L2 fp32 reduce-add
  const long long pblocks = (M + WG_BLOCK_P - 1) / WG_BLOCK_P;
  double best = 1e30, t_default = 1e30;
  int best_mb = max_mb, best_cn = max_cn;
  for (int mb = 1; mb <= max_mb; ++mb) {
    for (int cn = 64; cn <= max_cn; cn *= 2) {
      const long long tiles = (long long)ceil_div(cout_pad, 128 * mb) * ceil_div(cin_pad, cn) * taps;
      long long ks = k_splits_arg > 0 ? k_splits_arg : num_sms() / tiles;
      if (ks < 1) ks = 1;
      const long long max_ks = pblocks / 4 > 0 ? pblocks / 4 : 1;
      if (ks > max_ks) ks = max_ks;
      const double waves = (double)((tiles * ks + num_sms() - 1) / num_sms());
      const double px = (double)((pblocks + ks - 1) / ks) * WG_BLOCK_P;
      const double t_load = px * (128.0 * mb + cn) * 2.0 / kLoadBytesPerUs;
      const double t_mma = px * 128.0 * mb * cn / kMacPerUs;
      const double t_red = (double)tiles * ks * (128.0 * mb * cn * 4.0) / kReduceBytesPerUs;
      const double t = waves * ((t_load > t_mma ? t_load : t_mma) + 1.5) + t_red;
      if (mb == max_mb && cn == max_cn) t_default = t;
      if (t < best) {
        best = t;
        best_mb = mb;
        best_cn = cn;
      }
    }
  }
  // the largest tile is the measured default; leave it only for a predicted gain of at least 10 %
  if (best < 0.9 * t_default) {
    *MB = best_mb;
    *CN = best_cn;
  } else {
    *MB = max_mb;
    *CN = max_cn;
  }
}

extern "C" int zs3_conv_wgrad(const zs3_wgrad_args* a, void* stream) {
  ZS3_CHECK_ARG(a != nullptr, "conv_wgrad: null args");
  ZS3_CHECK_ARG(a->cout_pad > 0 && a->cout_pad % 64 == 0 && a->cin_pad > 0 && a->cin_pad % 64 == 0,
                "conv_wgrad: cout_pad=%d cin_pad=%d must be multiples of 64", a->cout_pad, a->cin_pad);
  ZS3_CHECK_ARG(a->x && a->dy && a->dw, "conv_wgrad: null pointer");
  ZS3_CHECK_ARG(a->x_cstride >= a->cin_pad && a->x_cstride % 8 == 0 && a->dy_cstride >= a->cout_pad &&
                    a->dy_cstride % 8 == 0,
                "conv_wgrad: bad channel strides");
  ZS3_CHECK_ARG(a->pad <= 128 && (a->R - 1) * a->dil <= 255 && (a->S - 1) * a->dil <= 255,
                "conv_wgrad: pad/dilation outside the TMA im2col range");
  const long long M = (long long)a->N * a->Ho * a->Wo;
  ZS3_CHECK_ARG(M > 0 && M < (1ll << 31), "conv_wgrad: bad pixel count");
  WgradParams p;
  memset(&p, 0, sizeof(p));
  int CN = a->cin_pad >= 256 ? 256 : (a->cin_pad >= 128 ? 128 : 64);
  int MB = a->cout_pad >= 256 ? 2 : 1;  // 256 output channels per CTA when the layer has them
  choose_wgrad_tile(M, a->cout_pad, a->cin_pad, a->R * a->S, a->k_splits, &MB, &CN);
  int rc = encode_tiled2d_bf16(&p.dy, a->dy, M, a->cout_pad, a->dy_cstride, WG_BLOCK_P, 64);
  if (rc) return rc;
  rc = encode_im2col_bf16(&p.x, a->x, a->N, a->H, a->W, a->x_cstride, a->pad, a->pad - (a->S - 1) * a->dil, a->stride,
                          64, WG_BLOCK_P);
  if (rc) return rc;
  p.M = (int)M;
  p.HoWo = a->Ho * a->Wo;
  p.Wo = a->Wo;
  p.stride = a->stride;
  p.pad = a->pad;
  p.dil = a->dil;
  p.R = a->R;
  p.S = a->S;
  p.cout_pad = a->cout_pad;
  p.cin_pad = a->cin_pad;
  p.num_co_tiles = ceil_div(a->cout_pad, 128 * MB);
  p.num_ci_tiles = ceil_div(a->cin_pad, CN);
  p.pblocks_total = (int)ceil_div_ll(M, WG_BLOCK_P);
  const int out_tiles = p.num_co_tiles * p.num_ci_tiles * a->R * a->S;
  int ks = a->k_splits;
  if (ks <= 0) {
    // Split the pixel reduction so that the grid is (at most) ONE full wave of CTAs: every extra split costs a
    // whole fp32 atomic pass over the dW tile, and a grid of waves*SMs + 1 CTAs costs a whole extra wave
    // (ncu, profiles/r01: 297 CTAs on 148 SMs ran as three waves).
    ks = num_sms() / out_tiles;
    if (ks < 1) ks = 1;
    const int max_ks = p.pblocks_total / 4 > 0 ? p.pblocks_total / 4 : 1;  // >= 4 k-blocks of 64 pixels per CTA
    if (ks > max_ks) ks = max_ks;
    static int cap = -1;  // experiment knob: upper bound on the pixel splits (fewer splits = less fp32 reduce traffic)
    if (cap < 0) {
      const char* e = getenv("ZS3_WGRAD_MAX_SPLITS");
      cap = e ? atoi(e) : 0;
    }
    if (cap > 0 && ks > cap) ks = cap;
  }
  if (ks > p.pblocks_total) ks = p.pblocks_total;
  p.k_splits = ks;
  p.dw = a->dw;
  p.dw_ld = a->dw_ld > 0 ? a->dw_ld : a->cin_pad;
  p.dw_ci_offset = a->dw_ci_offset;
  p.cout_valid = a->cout_valid > 0 ? a->cout_valid : a->cout_pad;
  p.cin_valid = a->cin_valid > 0 ? a->cin_valid : a->cin_pad;
  p.use_tma_reduce = 0;
  {
    static int pref = -1;
    if (pref < 0) {
      const char* env = getenv("ZS3_TMA_REDUCE");
      pref = env ? atoi(env) : 1;
    }
    if (pref && p.dw_ld % 4 == 0 && p.dw_ci_offset % 4 == 0 && (reinterpret_cast<uintptr_t>(a->dw) & 15) == 0) {
      int rc3 = encode_tiled3d_f32(&p.dwmap, a->dw + p.dw_ci_offset, p.cout_valid, a->R * a->S, p.cin_valid,
                                   (long long)a->R * a->S * p.dw_ld, p.dw_ld, 32, 1, 32);
      if (rc3) return rc3;
      p.use_tma_reduce = 1;
    }
  }
  ZS3_CHECK_ARG(p.cout_valid <= a->cout_pad && p.cin_valid <= a->cin_pad && p.dw_ci_offset >= 0 &&
                    p.dw_ld >= p.dw_ci_offset + p.cin_valid,
                "conv_wgrad: bad dw view (ld=%lld offset=%d cin_valid=%d cout_valid=%d)", p.dw_ld, p.dw_ci_offset,
                p.cin_valid, p.cout_valid);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (MB == 2) {
    if (CN == 256) return launch_wgrad<256, 3, 2>(p, st);
    if (CN == 128) return launch_wgrad<128, 4, 2>(p, st);
    return launch_wgrad<64, 5, 2>(p, st);
  }
  if (CN == 256) return launch_wgrad<256, 4, 1>(p, st);
  if (CN == 128) return launch_wgrad<128, 6, 1>(p, st);
  return launch_wgrad<64, 8, 1>(p, st);
}

extern "C" int zs3_debug_wgrad_tile(long long M, int cout_pad, int cin_pad, int taps, int* mb, int* cn) {
  ZS3_CHECK_ARG(M > 0 && cout_pad > 0 && cout_pad % 64 == 0 && cin_pad > 0 && cin_pad % 64 == 0 && taps > 0 && mb && cn,
                "debug_wgrad_tile: bad args");
  *cn = cin_pad >= 256 ? 256 : (cin_pad >= 128 ? 128 : 64);
  *mb = cout_pad >= 256 ? 2 : 1;
  choose_wgrad_tile(M, cout_pad, cin_pad, taps, 0, mb, cn);
  return ZS3_OK;
}

extern "C" int zs3_pack_weight(const float* w_oihw, int Cout, int Cin, int R, int S, int ci_begin, int ci_count,
                               void* dst_bf16, int cout_pad, int cin_pad, int mode, void* stream) {
  ZS3_CHECK_ARG(w_oihw && dst_bf16, "pack_weight: null pointer");
  ZS3_CHECK_ARG(Cout <= cout_pad && ci_count <= cin_pad && ci_begin >= 0 && ci_begin + ci_count <= Cin,
                "pack_weight: bad channel ranges");
  ZS3_CHECK_ARG(mode >= 0 && mode <= 3, "pack_weight: mode=%d", mode);
  const int src_krsc = mode >> 1;
  mode &= 1;
  if (src_krsc) {
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    __nv_bfloat16* d = static_cast<__nv_bfloat16*>(dst_bf16);
    if (mode == 0) {
      const long long total = (long long)cout_pad * R * S * (cin_pad / 4);
      int blocks = (int)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
      pack_krsc_fprop_kernel<<<blocks, 256, 0, st>>>(w_oihw, Cout, Cin, R * S, ci_begin, ci_count, d, cout_pad, cin_pad);
    } else {
      dim3 grid(ceil_div(cin_pad, 32), ceil_div(cout_pad, 32), R * S);
      pack_krsc_dgrad_kernel<<<grid, 256, 0, st>>>(w_oihw, Cout, Cin, R * S, ci_begin, ci_count, d, cout_pad, cin_pad);
    }
    ZS3_CHECK_LAUNCH("pack_weight(krsc)");
    return ZS3_OK;
  }
  ZS3_CHECK_ARG(R * S <= 9, "pack_weight: at most 9 taps (the 7x7 stem goes through its [Cout][147][1][1] view)");
  const size_t smem = (size_t)PK_T * (PK_T * R * S + 1) * sizeof(float);
  dim3 grid(ceil_div(cin_pad, PK_T), ceil_div(cout_pad, PK_T));
  pack_weight_kernel<<<grid, 256, smem, static_cast<cudaStream_t>(stream)>>>(
      w_oihw, Cout, Cin, R, S, ci_begin, ci_count, static_cast<__nv_bfloat16*>(dst_bf16), cout_pad, cin_pad, mode,
      src_krsc);
  ZS3_CHECK_LAUNCH("pack_weight");
  return ZS3_OK;
}

extern "C" int zs3_unpack_wgrad(const float* dw, int cout_pad, int cin_pad, float* grad_oihw, int Cout, int Cin, int R,
                                int S, int ci_begin, int ci_count, int accumulate, void* stream) {
  ZS3_CHECK_ARG(dw && grad_oihw, "unpack_wgrad: null pointer");
  ZS3_CHECK_ARG(Cout <= cout_pad && ci_count <= cin_pad && ci_begin >= 0 && ci_begin + ci_count <= Cin,
                "unpack_wgrad: bad channel ranges");
  ZS3_CHECK_ARG(R * S <= 9, "unpack_wgrad: at most 9 taps");
  const size_t smem = (size_t)PK_T * (PK_T * R * S + 1) * sizeof(float);
  dim3 grid(ceil_div(ci_count, PK_T), ceil_div(Cout, PK_T));
  unpack_wgrad_kernel<<<grid, 256, smem, static_cast<cudaStream_t>(stream)>>>(dw, cout_pad, cin_pad, grad_oihw, Cout,
                                                                              Cin, R, S, ci_begin, ci_count, accumulate);
  ZS3_CHECK_LAUNCH("unpack_wgrad");
  return ZS3_OK;
}
