// Validation metrics on the device: argmax over the class logits + confusion-matrix accumulation,
// replacing the reference's per-batch device->host copy of the full logits (354 MB at bs=16, 21 classes, 513x513),
// numpy argmax and np.bincount (zs3/train_pascal_GMMN.py:371-375 -> zs3/utils/metrics.py:73-82).
//
//   pred[b][p]            = first index of the maximum over c of logits[b][c][p]      (np.argmax, axis=1)
//   conf[gt][pred]       += 1 for every pixel with 0 <= gt < num_class                (Evaluator._generate_matrix)
//
// HBM-bound: C*4 B read + 4 B label read + 1 B prediction written per pixel.  Two pixels per thread, class planes
// read coalesced; counts go to a per-CTA shared-memory histogram (C*C ints) and are flushed with one 64-bit atomic
// per non-empty bin.  Integer outputs: bit-exact against the oracle.
//
// tests/test_kernel_emulation.py also compiles this file for the host (-DZS3_HOST_EMULATION, tests/emul/cuda_emul.h).
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

constexpr int AM_THREADS = 256;
constexpr int AM_MAXC = 64;
constexpr int AM_PIX = 2;

struct ArgmaxP {
  const float* logits;   // [B][C][HW], or null when pred_in is given
  const int* pred_in;    // [B][HW] predictions computed elsewhere (Evaluator.add_batch), or null
  const float* target;   // [B][HW] float class ids, or null
  unsigned char* pred;   // [B][HW] or null
  unsigned long long* conf;  // [C][C] or null
  int B, C;
  long long HW;
};

__global__ void __launch_bounds__(AM_THREADS) argmax_confusion_kernel(const ArgmaxP p) {
  __shared__ int hist[AM_MAXC * AM_MAXC];
  const int C = p.C;
  const bool count = p.conf != nullptr && p.target != nullptr;
  if (count) {
    for (int i = threadIdx.x; i < C * C; i += AM_THREADS) hist[i] = 0;
    __syncthreads();
  }
  // AM_PIX pixels per thread (block-strided, so every class-plane read stays coalesced) x 8 unrolled class steps
  // = 16 independent loads in flight per thread; small iterations (512 pixels per CTA) keep the 1184-CTA grid-stride
  // loop balanced to within one iteration in ~7
  const long long total = (long long)p.B * p.HW;
  for (long long base = (long long)blockIdx.x * (AM_THREADS * AM_PIX); base < total;
       base += (long long)gridDim.x * (AM_THREADS * AM_PIX)) {
    long long q[AM_PIX];
    const float* src[AM_PIX];
    bool ok[AM_PIX];
    int arg[AM_PIX];
#pragma unroll
    for (int e = 0; e < AM_PIX; ++e) {
      q[e] = base + e * AM_THREADS + threadIdx.x;
      ok[e] = q[e] < total;
      const long long qq = ok[e] ? q[e] : 0;
      const long long b = qq / p.HW, px = qq - b * p.HW;
      src[e] = p.logits ? p.logits + b * C * p.HW + px : nullptr;
      arg[e] = 0;
    }
    if (p.logits) {
      float best[AM_PIX];
#pragma unroll
      for (int e = 0; e < AM_PIX; ++e) best[e] = __ldg(src[e]);
#pragma unroll 8
      for (int c = 1; c < C; ++c) {
        float v[AM_PIX];
#pragma unroll
        for (int e = 0; e < AM_PIX; ++e) v[e] = __ldg(src[e] + (long long)c * p.HW);
#pragma unroll
        for (int e = 0; e < AM_PIX; ++e)
          if (v[e] > best[e]) {   // strict: the first maximum wins, like np.argmax
            best[e] = v[e];
            arg[e] = c;
          }
      }
#pragma unroll
      for (int e = 0; e < AM_PIX; ++e)
        if (ok[e] && p.pred) p.pred[q[e]] = (unsigned char)arg[e];
    } else {
#pragma unroll
      for (int e = 0; e < AM_PIX; ++e)
        if (ok[e]) arg[e] = __ldg(p.pred_in + q[e]);
    }
    if (count) {
#pragma unroll
      for (int e = 0; e < AM_PIX; ++e) {
        if (!ok[e]) continue;
        const float g = __ldg(p.target + q[e]);
        // metrics.py:74-75: mask = (gt >= 0) & (gt < num_class); gt.astype(int) truncates; a prediction outside
        // [0, C) cannot come from an argmax and is skipped here (np.bincount would spill it into the next row)
        if (g >= 0.f && g < (float)C && arg[e] >= 0 && arg[e] < C) atomicAdd(&hist[(int)g * C + arg[e]], 1);
      }
    }
  }
  if (count) {
    __syncthreads();
    for (int i = threadIdx.x; i < C * C; i += AM_THREADS)
      if (hist[i]) atomicAdd(p.conf + i, (unsigned long long)hist[i]);
  }
}

}  // namespace zs3

using namespace zs3;

#ifdef ZS3_HOST_EMULATION
extern "C" int zs3_emul_argmax_confusion(const float* logits, const float* target, int B, int C, long long HW,
                                         unsigned char* pred, unsigned long long* conf, void* stream) {
#else
extern "C" int zs3_argmax_confusion(const float* logits, const float* target, int B, int C, long long HW,
                                    unsigned char* pred, unsigned long long* conf, void* stream) {
#endif
  ZS3_CHECK_ARG(logits && B >= 0 && C >= 1 && C <= AM_MAXC && HW > 0, "argmax_confusion: bad args (1 <= C <= 64)");
  ZS3_CHECK_ARG(pred || (conf && target), "argmax_confusion: nothing to compute");
  ZS3_CHECK_ARG(!conf || target, "argmax_confusion: the confusion matrix needs the target labels");
  if (B == 0) return ZS3_OK;
  ArgmaxP p;
  p.logits = logits; p.pred_in = nullptr; p.target = target; p.pred = pred; p.conf = conf; p.B = B; p.C = C; p.HW = HW;
  const long long total = (long long)B * HW;
  long long blocks = (total + AM_THREADS * AM_PIX - 1) / (AM_THREADS * AM_PIX);
#ifdef ZS3_HOST_EMULATION
  const int nb = stream ? (int)reinterpret_cast<intptr_t>(stream) : 1;
  for (int b = 0; b < nb; ++b) emul_run_block<ArgmaxP>(argmax_confusion_kernel, p, AM_THREADS, b, nb);
  (void)blocks;
  return ZS3_OK;
#else
  if (blocks > 148 * 8) blocks = 148 * 8;  // 8 resident CTAs per SM (16 KB histogram each), grid-stride beyond
  argmax_confusion_kernel<<<(int)blocks, AM_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(p);
  ZS3_CHECK_LAUNCH("argmax_confusion");
  return ZS3_OK;
#endif
}

#ifdef ZS3_HOST_EMULATION
extern "C" int zs3_emul_confusion_from_pred(const int* pred, const float* target, long long n, int C,
                                            unsigned long long* conf, void* stream) {
#else
extern "C" int zs3_confusion_from_pred(const int* pred, const float* target, long long n, int C,
                                       unsigned long long* conf, void* stream) {
#endif
  ZS3_CHECK_ARG(pred && target && conf && n >= 0 && C >= 1 && C <= AM_MAXC, "confusion_from_pred: bad args (1 <= C <= 64)");
  if (n == 0) return ZS3_OK;
  ArgmaxP p;
  p.logits = nullptr; p.pred_in = pred; p.target = target; p.pred = nullptr; p.conf = conf; p.B = 1; p.C = C; p.HW = n;
#ifdef ZS3_HOST_EMULATION
  (void)stream;
  emul_run_block<ArgmaxP>(argmax_confusion_kernel, p, AM_THREADS, 0, 1);
  return ZS3_OK;
#else
  long long blocks = (n + AM_THREADS * AM_PIX - 1) / (AM_THREADS * AM_PIX);
  if (blocks > 148 * 8) blocks = 148 * 8;
  argmax_confusion_kernel<<<(int)blocks, AM_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(p);
  ZS3_CHECK_LAUNCH("confusion_from_pred");
  return ZS3_OK;
#endif
}
