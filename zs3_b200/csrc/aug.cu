// Training / validation input transforms on the device: replaces the per-sample PIL pipeline of
// zs3/dataloaders/custom_transforms.py (RandomHorizontalFlip :47-56, RandomScaleCrop :69-104, RandomGaussianBlur :58-66,
// FixScale :107-124, Normalize :8-27, ToTensor :30-44; composed in datasets/pascal.py:120-144) for a whole batch.
// Input: the decoded pictures as bytes (RGB HWC + label map) and the random draws of the reference's transforms; output:
// the [n,3,H,W] float image batch and [n,H,W] float label batch the trainer consumes.  Byte / integer work, HBM-bound,
// BIT-EXACT against Pillow (the algorithms restated in oracle/zs3_transforms_oracle.py):
//   * antialiased triangle-filter resize as two 8-bit passes with 22-bit fixed-point coefficients (Resample.c); only the
//     crop window is ever computed: horizontal pass over the source rows the window's vertical taps touch, then the
//     vertical pass straight into the crop; right/bottom padding (image 0, label `fill`) is written in the same pass;
//   * nearest label resize with Pillow's ACCUMULATED double coordinate (Geometry.c ImagingScaleAffine);
//   * GaussianBlur = 3 + 3 box-blur passes on bytes (BoxBlur.c), rows then columns, each group in shared memory;
//   * /255, -mean, /std through a 3x256 table built by the host with numpy's own arithmetic; HWC -> CHW on the store.
// The flip is a mirrored source index, so no flipped copy exists.  The coefficient tables are computed on the device in
// IEEE double with explicit round-to-nearest operations (no FMA contraction), one plan block per image and axis.
//
// tests/test_kernel_emulation.py also compiles this file for the host (-DZS3_HOST_EMULATION, tests/emul/cuda_emul.h).
#ifdef ZS3_HOST_EMULATION
#include "cuda_emul.h"
#define ZS3_CHECK_ARG(cond, ...) \
  do {                           \
    if (!(cond)) return -1;      \
  } while (0)
static inline double __dadd_rn(double a, double b) { return a + b; }
static inline double __dsub_rn(double a, double b) { return a - b; }
static inline double __dmul_rn(double a, double b) { return a * b; }
static inline double __ddiv_rn(double a, double b) { return a / b; }
static inline double __dsqrt_rn(double a) { return std::sqrt(a); }
static inline float __fadd_rn(float a, float b) { return a + b; }
static inline float __fsub_rn(float a, float b) { return a - b; }
static inline float __fmul_rn(float a, float b) { return a * b; }
static inline float __fdiv_rn(float a, float b) { return a / b; }
static inline int __double2int_rz(double a) { return (int)a; }
static inline unsigned __float2uint_rz(float a) { return (unsigned)a; }
#else
#include "common.cuh"
#endif

namespace zs3 {

constexpr int AUG_THREADS = 256;
constexpr int AUG_KMAX = 17;       // taps per output = 2 * ceil(max(scale, 1)) + 1: down-scaling up to 8x
constexpr int AUG_STRIP = 16;      // pixels per column strip of the vertical blur
constexpr int AUG_HEADER = 8;      // ints: first source row of the horizontal pass, its row count, blur (on, r, ww, fw)

struct AugP {
  const zs3_aug_item* items;
  int n, max_src_h, out_w, out_h, fill_label;
  const float* lut;
  float* out_image;
  float* out_label;
  int* plan;            // [n][plan_ints]
  unsigned char* tmp;   // [n][max_src_h][out_w][3] horizontal-pass rows
  unsigned char* crop;  // [n][out_h][out_w][3]
  long long plan_ints;
};

// per-image plan: header | hmin[W] hcnt[W] hk[W][KMAX] | vmin[H] vcnt[H] vk[H][KMAX] | xidx[W] | yidx[H]
__device__ __forceinline__ int* plan_hmin(const AugP& p, int* pl) { return pl + AUG_HEADER; }
__device__ __forceinline__ int* plan_hcnt(const AugP& p, int* pl) { return plan_hmin(p, pl) + p.out_w; }
__device__ __forceinline__ int* plan_hk(const AugP& p, int* pl) { return plan_hcnt(p, pl) + p.out_w; }
__device__ __forceinline__ int* plan_vmin(const AugP& p, int* pl) { return plan_hk(p, pl) + (long long)p.out_w * AUG_KMAX; }
__device__ __forceinline__ int* plan_vcnt(const AugP& p, int* pl) { return plan_vmin(p, pl) + p.out_h; }
__device__ __forceinline__ int* plan_vk(const AugP& p, int* pl) { return plan_vcnt(p, pl) + p.out_h; }
__device__ __forceinline__ int* plan_xidx(const AugP& p, int* pl) { return plan_vk(p, pl) + (long long)p.out_h * AUG_KMAX; }
__device__ __forceinline__ int* plan_yidx(const AugP& p, int* pl) { return plan_xidx(p, pl) + p.out_w; }

// Resample.c precompute_coeffs + normalize_coeffs_8bpc for output coordinate xx of an axis in_size -> rs_size
__device__ __forceinline__ void triangle_coeffs(int in_size, int rs_size, int xx, int* kmin, int* kcnt, int* kk) {
  const double scale = __ddiv_rn((double)(float)in_size, (double)rs_size);
  const double filterscale = scale < 1.0 ? 1.0 : scale;
  const double support = filterscale;                       // triangle filter: support 1.0 * filterscale
  const double ss = __ddiv_rn(1.0, filterscale);
  const double center = __dmul_rn(__dadd_rn((double)xx, 0.5), scale);
  int lo = __double2int_rz(__dadd_rn(__dsub_rn(center, support), 0.5));
  if (lo < 0) lo = 0;
  int hi = __double2int_rz(__dadd_rn(__dadd_rn(center, support), 0.5));
  if (hi > in_size) hi = in_size;
  int n = hi - lo;
  if (n > AUG_KMAX) n = AUG_KMAX;                           // unreachable: the host entry bounds the scale
  double w[AUG_KMAX];
  double ww = 0.0;
  for (int x = 0; x < n; ++x) {
    double t = __dmul_rn(__dadd_rn(__dsub_rn((double)(x + lo), center), 0.5), ss);
    if (t < 0.0) t = -t;
    w[x] = t < 1.0 ? __dsub_rn(1.0, t) : 0.0;
    ww = __dadd_rn(ww, w[x]);
  }
  for (int x = 0; x < n; ++x) {
    const double v = ww != 0.0 ? __ddiv_rn(w[x], ww) : w[x];
    kk[x] = __double2int_rz(__dadd_rn(0.5, __dmul_rn(v, 4194304.0)));   // 1 << (32 - 8 - 2); v >= 0
  }
  *kmin = lo;
  *kcnt = n;
}

// One block per (image, axis): axis 0 = columns, axis 1 = rows.
__global__ void __launch_bounds__(AUG_THREADS) aug_plan_kernel(const AugP p) {
  const int img = blockIdx.x >> 1, axis = blockIdx.x & 1;
  const zs3_aug_item it = p.items[img];
  int* pl = p.plan + (long long)img * p.plan_ints;
  const int in_size = axis ? it.h : it.w, rs_size = axis ? it.rh : it.rw;
  const int origin = axis ? it.y1 : it.x1, out_len = axis ? p.out_h : p.out_w;
  int* kmin = axis ? plan_vmin(p, pl) : plan_hmin(p, pl);
  int* kcnt = axis ? plan_vcnt(p, pl) : plan_hcnt(p, pl);
  int* kk = axis ? plan_vk(p, pl) : plan_hk(p, pl);
  int* nidx = axis ? plan_yidx(p, pl) : plan_xidx(p, pl);
  for (int o = threadIdx.x; o < out_len; o += AUG_THREADS) {
    const int r = origin + o;
    if (r < rs_size) {
      triangle_coeffs(in_size, rs_size, r, kmin + o, kcnt + o, kk + (long long)o * AUG_KMAX);
    } else {
      kmin[o] = 0;
      kcnt[o] = 0;        // right / bottom padding of ImageOps.expand
      nidx[o] = -2;
    }
  }
  if (threadIdx.x == 0) {
    // Geometry.c ImagingScaleAffine: the source coordinate is accumulated in double from output pixel 0
    const double a = __ddiv_rn((double)in_size, (double)rs_size);
    double xo = __dmul_rn(a, 0.5);
    const int last = (origin + out_len < rs_size) ? origin + out_len : rs_size;
    for (int x = 0; x < last; ++x) {
      if (x >= origin) {
        int xin = xo < 0.0 ? -1 : __double2int_rz(xo);
        nidx[x - origin] = (xin >= 0 && xin < in_size) ? xin : -1;
      }
      xo = __dadd_rn(xo, a);
    }
    if (axis == 0) {
      // BoxBlur.c _gaussian_blur_radius (3 passes) and the fixed-point weights of ImagingHorizontalBoxBlur
      int on = 0, br = 0;
      unsigned ww = 0, fw = 0;
      if (it.blur_radius > 0.f) {
        const float radius = it.blur_radius;
        const float sigma2 = __fdiv_rn(__fmul_rn(radius, radius), 3.f);
        const float L = (float)__dsqrt_rn(__dadd_rn(__dmul_rn(12.0, (double)sigma2), 1.0));
        const float l = (float)floor(__ddiv_rn(__dsub_rn((double)L, 1.0), 2.0));
        float av = __fmul_rn(__fadd_rn(__fmul_rn(2.f, l), 1.f),
                             __fsub_rn(__fmul_rn(l, __fadd_rn(l, 1.f)), __fmul_rn(3.f, sigma2)));
        av = __fdiv_rn(av, __fmul_rn(6.f, __fsub_rn(sigma2, __fmul_rn(__fadd_rn(l, 1.f), __fadd_rn(l, 1.f)))));
        const float fr = __fadd_rn(l, av);
        if (fr != 0.f) {
          on = 1;
          br = (int)fr;
          ww = __float2uint_rz(__fdiv_rn(16777216.f, __fadd_rn(__fmul_rn(fr, 2.f), 1.f)));
          fw = ((1u << 24) - (unsigned)(br * 2 + 1) * ww) / 2u;
        }
      }
      pl[2] = on; pl[3] = br; pl[4] = (int)ww; pl[5] = (int)fw;
    }
  }
  if (axis == 1) {
    // rows the horizontal pass has to produce: the union of the vertical taps of the window's in-picture rows
    __syncthreads();
    if (threadIdx.x == 0) {
      int n_in = it.rh - origin;
      if (n_in > out_len) n_in = out_len;
      int lo = 0, hi = 0;
      if (n_in > 0) {
        lo = kmin[0];
        hi = kmin[n_in - 1] + kcnt[n_in - 1];
      }
      pl[0] = lo;
      pl[1] = hi - lo;
    }
  }
}

// Resample.c ImagingResampleHorizontal_8bpc over the rows the vertical pass needs, columns of the crop window only.
// grid: n * ceil(max_src_h * out_w / AUG_THREADS)
__global__ void __launch_bounds__(AUG_THREADS) aug_hpass_kernel(const AugP p) {
  const int per_img = (int)(((long long)p.max_src_h * p.out_w + AUG_THREADS - 1) / AUG_THREADS);
  const int img = blockIdx.x / per_img;
  const long long e = (long long)(blockIdx.x - img * per_img) * AUG_THREADS + threadIdx.x;
  int* pl = p.plan + (long long)img * p.plan_ints;
  const int row = (int)(e / p.out_w), x = (int)(e - (long long)row * p.out_w);
  if (row >= pl[1]) return;
  const int n = plan_hcnt(p, pl)[x];
  if (n == 0) return;
  const zs3_aug_item it = p.items[img];
  const int lo = plan_hmin(p, pl)[x];
  const int* k = plan_hk(p, pl) + (long long)x * AUG_KMAX;
  const unsigned char* src = it.image + (long long)(pl[0] + row) * it.w * 3;
  int s0 = 1 << 21, s1 = 1 << 21, s2 = 1 << 21;
  for (int j = 0; j < n; ++j) {
    const int sx = it.flip ? it.w - 1 - (lo + j) : lo + j;
    const int kv = k[j];
    s0 += src[sx * 3 + 0] * kv;
    s1 += src[sx * 3 + 1] * kv;
    s2 += src[sx * 3 + 2] * kv;
  }
  unsigned char* dst = p.tmp + (((long long)img * p.max_src_h + row) * p.out_w + x) * 3;
  s0 >>= 22; s1 >>= 22; s2 >>= 22;
  dst[0] = (unsigned char)(s0 < 0 ? 0 : (s0 > 255 ? 255 : s0));
  dst[1] = (unsigned char)(s1 < 0 ? 0 : (s1 > 255 ? 255 : s1));
  dst[2] = (unsigned char)(s2 < 0 ? 0 : (s2 > 255 ? 255 : s2));
}

// Vertical pass + padding + label gather (+ Normalize/ToTensor when the image is not blurred).
// grid: n * ceil(out_h * out_w / AUG_THREADS)
__global__ void __launch_bounds__(AUG_THREADS) aug_vpass_kernel(const AugP p) {
  const int per_img = (int)(((long long)p.out_h * p.out_w + AUG_THREADS - 1) / AUG_THREADS);
  const int img = blockIdx.x / per_img;
  const long long e = (long long)(blockIdx.x - img * per_img) * AUG_THREADS + threadIdx.x;
  if (e >= (long long)p.out_h * p.out_w) return;
  int* pl = p.plan + (long long)img * p.plan_ints;
  const int y = (int)(e / p.out_w), x = (int)(e - (long long)y * p.out_w);
  const zs3_aug_item it = p.items[img];
  const int nv = plan_vcnt(p, pl)[y], nh = plan_hcnt(p, pl)[x];
  int s0 = 0, s1 = 0, s2 = 0;
  if (nv > 0 && nh > 0) {
    const int* k = plan_vk(p, pl) + (long long)y * AUG_KMAX;
    const unsigned char* src =
        p.tmp + (((long long)img * p.max_src_h + (plan_vmin(p, pl)[y] - pl[0])) * p.out_w + x) * 3;
    s0 = s1 = s2 = 1 << 21;
    for (int j = 0; j < nv; ++j) {
      const int kv = k[j];
      s0 += src[0] * kv;
      s1 += src[1] * kv;
      s2 += src[2] * kv;
      src += (long long)p.out_w * 3;
    }
    s0 >>= 22; s1 >>= 22; s2 >>= 22;
    s0 = s0 < 0 ? 0 : (s0 > 255 ? 255 : s0);
    s1 = s1 < 0 ? 0 : (s1 > 255 ? 255 : s1);
    s2 = s2 < 0 ? 0 : (s2 > 255 ? 255 : s2);
  }
  const long long plane = (long long)p.out_h * p.out_w;
  if (pl[2]) {
    unsigned char* dst = p.crop + ((long long)img * plane + e) * 3;
    dst[0] = (unsigned char)s0; dst[1] = (unsigned char)s1; dst[2] = (unsigned char)s2;
  } else {
    float* o = p.out_image + (long long)img * 3 * plane + e;
    o[0] = p.lut[s0];
    o[plane] = p.lut[256 + s1];
    o[2 * plane] = p.lut[512 + s2];
  }
  if (p.out_label) {
    const int yi = plan_yidx(p, pl)[y], xi = plan_xidx(p, pl)[x];
    int lab = 0;                                   // ImagingScaleAffine leaves pixels without a source at 0
    if (yi == -2 || xi == -2) lab = p.fill_label;  // ImageOps.expand(mask, fill)
    else if (yi >= 0 && xi >= 0) lab = it.label[(long long)yi * it.w + (it.flip ? it.w - 1 - xi : xi)];
    p.out_label[(long long)img * plane + e] = (float)lab;
  }
}

#ifdef ZS3_HOST_EMULATION
#define AUG_DYN_SMEM(name) static unsigned char name[2 * 4096 * 3 * AUG_STRIP]
#else
#define AUG_DYN_SMEM(name) extern __shared__ unsigned char name[]
#endif

// BoxBlur.c ImagingLineBoxBlur on one line held in shared memory: len elements of `stride` bytes apart per channel
__device__ __forceinline__ unsigned char box_tap(const unsigned char* line, int i, int len, int step, int r, unsigned ww,
                                                 unsigned fw) {
  unsigned acc = 0;
  for (int j = -r; j <= r; ++j) {
    int q = i + j;
    q = q < 0 ? 0 : (q > len - 1 ? len - 1 : q);
    acc += line[q * step];
  }
  int ql = i - r - 1, qr = i + r + 1;
  ql = ql < 0 ? 0 : ql;
  qr = qr > len - 1 ? len - 1 : qr;
  const unsigned bulk = acc * ww + ((unsigned)line[ql * step] + (unsigned)line[qr * step]) * fw;
  return (unsigned char)((bulk + (1u << 23)) >> 24);
}

// three horizontal box passes of one row; grid: n * out_h, dynamic shared memory 2 * out_w * 3
__global__ void __launch_bounds__(AUG_THREADS) aug_blur_rows_kernel(const AugP p) {
  AUG_DYN_SMEM(sm);
  const int img = blockIdx.x / p.out_h, y = blockIdx.x - img * p.out_h;
  const int* pl = p.plan + (long long)img * p.plan_ints;
  if (!pl[2]) return;
  const int r = pl[3];
  const unsigned ww = (unsigned)pl[4], fw = (unsigned)pl[5];
  const int nb = p.out_w * 3;
  unsigned char* row = p.crop + ((long long)img * p.out_h + y) * nb;
  unsigned char* a = sm;
  unsigned char* b = sm + nb;
  for (int i = threadIdx.x; i < nb; i += AUG_THREADS) a[i] = row[i];
  __syncthreads();
  for (int pass = 0; pass < 3; ++pass) {
    for (int i = threadIdx.x; i < nb; i += AUG_THREADS) {
      const int x = i / 3, c = i - x * 3;
      b[i] = box_tap(a + c, x, p.out_w, 3, r, ww, fw);
    }
    __syncthreads();
    unsigned char* t = a; a = b; b = t;
  }
  for (int i = threadIdx.x; i < nb; i += AUG_THREADS) row[i] = a[i];
}

// three vertical box passes of a strip of AUG_STRIP columns, then Normalize/ToTensor of the strip;
// grid: n * ceil(out_w / AUG_STRIP), dynamic shared memory 2 * out_h * AUG_STRIP * 3
__global__ void __launch_bounds__(AUG_THREADS) aug_blur_cols_kernel(const AugP p) {
  AUG_DYN_SMEM(sm);
  const int strips = (p.out_w + AUG_STRIP - 1) / AUG_STRIP;
  const int img = blockIdx.x / strips, x0 = (blockIdx.x - img * strips) * AUG_STRIP;
  const int* pl = p.plan + (long long)img * p.plan_ints;
  if (!pl[2]) return;
  const int r = pl[3];
  const unsigned ww = (unsigned)pl[4], fw = (unsigned)pl[5];
  const int wpx = (p.out_w - x0 < AUG_STRIP) ? p.out_w - x0 : AUG_STRIP;
  const int wb = wpx * 3, nb = p.out_h * wb;
  const unsigned char* src = p.crop + ((long long)img * p.out_h * p.out_w + x0) * 3;
  unsigned char* a = sm;
  unsigned char* b = sm + (long long)p.out_h * AUG_STRIP * 3;
  for (int i = threadIdx.x; i < nb; i += AUG_THREADS) {
    const int y = i / wb, j = i - y * wb;
    a[i] = src[(long long)y * p.out_w * 3 + j];
  }
  __syncthreads();
  for (int pass = 0; pass < 3; ++pass) {
    for (int i = threadIdx.x; i < nb; i += AUG_THREADS) {
      const int y = i / wb, j = i - y * wb;
      b[i] = box_tap(a + j, y, p.out_h, wb, r, ww, fw);
    }
    __syncthreads();
    unsigned char* t = a; a = b; b = t;
  }
  const long long plane = (long long)p.out_h * p.out_w;
  for (int i = threadIdx.x; i < p.out_h * wpx * 3; i += AUG_THREADS) {
    const int c = i / (p.out_h * wpx), q = i - c * (p.out_h * wpx);
    const int y = q / wpx, x = q - y * wpx;
    p.out_image[((long long)img * 3 + c) * plane + (long long)y * p.out_w + x0 + x] = p.lut[c * 256 + a[y * wb + x * 3 + c]];
  }
}

static inline long long aug_plan_ints(int out_w, int out_h) {
  return AUG_HEADER + (long long)(out_w + out_h) * (3 + AUG_KMAX);
}
static inline unsigned long long align16(unsigned long long v) { return (v + 15ull) & ~15ull; }

}  // namespace zs3

using namespace zs3;

#ifdef ZS3_HOST_EMULATION
extern "C" unsigned long long zs3_emul_augment_workspace_size(int n, int max_src_h, int out_w, int out_h) {
#else
extern "C" unsigned long long zs3_augment_workspace_size(int n, int max_src_h, int out_w, int out_h) {
#endif
  if (n <= 0 || max_src_h <= 0 || out_w <= 0 || out_h <= 0) return 0;
  return align16((unsigned long long)n * aug_plan_ints(out_w, out_h) * sizeof(int)) +
         align16((unsigned long long)n * max_src_h * out_w * 3) + align16((unsigned long long)n * out_h * out_w * 3);
}

#ifdef ZS3_HOST_EMULATION
extern "C" int zs3_emul_augment_batch(const zs3_augment_args* a, void* stream) {
#else
extern "C" int zs3_augment_batch(const zs3_augment_args* a, void* stream) {
#endif
  ZS3_CHECK_ARG(a != nullptr, "augment_batch: null args");
  ZS3_CHECK_ARG(a->n >= 0 && a->max_src_h > 0 && a->out_w > 0 && a->out_h > 0, "augment_batch: bad dims");
  ZS3_CHECK_ARG(a->out_w <= 4096 && a->out_h <= 2048, "augment_batch: output larger than 4096 x 2048");
  if (a->n == 0) return ZS3_OK;
  ZS3_CHECK_ARG(a->items && a->items_host && a->lut && a->out_image && a->workspace, "augment_batch: null pointer");
  int any_blur = 0;
  for (int i = 0; i < a->n; ++i) {
    const zs3_aug_item& it = a->items_host[i];
    ZS3_CHECK_ARG(it.image && it.w > 0 && it.h > 0 && it.rw > 0 && it.rh > 0 && it.x1 >= 0 && it.y1 >= 0,
                  "augment_batch: item %d: bad geometry", i);
    ZS3_CHECK_ARG(it.label || !a->out_label, "augment_batch: item %d: label map missing", i);
    ZS3_CHECK_ARG(it.h <= a->max_src_h, "augment_batch: item %d: %d rows > max_src_h %d", i, it.h, a->max_src_h);
    ZS3_CHECK_ARG((long long)it.w <= 8ll * it.rw && (long long)it.h <= 8ll * it.rh,
                  "augment_batch: item %d: down-scaling by more than 8x is not supported", i);
    any_blur |= it.blur_radius > 0.f;
  }
  const unsigned long long need =
#ifdef ZS3_HOST_EMULATION
      zs3_emul_augment_workspace_size(a->n, a->max_src_h, a->out_w, a->out_h);
#else
      zs3_augment_workspace_size(a->n, a->max_src_h, a->out_w, a->out_h);
#endif
  ZS3_CHECK_ARG(a->workspace_bytes >= need, "augment_batch: workspace %llu < %llu bytes", a->workspace_bytes, need);
  ZS3_CHECK_ARG((reinterpret_cast<uintptr_t>(a->workspace) & 15) == 0, "augment_batch: workspace not 16-byte aligned");
  AugP p;
  p.items = a->items; p.n = a->n; p.max_src_h = a->max_src_h; p.out_w = a->out_w; p.out_h = a->out_h;
  p.fill_label = a->fill_label; p.lut = a->lut; p.out_image = a->out_image; p.out_label = a->out_label;
  p.plan_ints = aug_plan_ints(a->out_w, a->out_h);
  unsigned char* ws = static_cast<unsigned char*>(a->workspace);
  p.plan = reinterpret_cast<int*>(ws);
  ws += align16((unsigned long long)a->n * p.plan_ints * sizeof(int));
  p.tmp = ws;
  ws += align16((unsigned long long)a->n * a->max_src_h * a->out_w * 3);
  p.crop = ws;
  const long long hblocks = (long long)a->n * (((long long)a->max_src_h * a->out_w + AUG_THREADS - 1) / AUG_THREADS);
  const long long vblocks = (long long)a->n * (((long long)a->out_h * a->out_w + AUG_THREADS - 1) / AUG_THREADS);
  const int strips = (a->out_w + AUG_STRIP - 1) / AUG_STRIP;
  const size_t smem_rows = 2 * (size_t)a->out_w * 3, smem_cols = 2 * (size_t)a->out_h * AUG_STRIP * 3;
  ZS3_CHECK_ARG(hblocks < (1ll << 31) && vblocks < (1ll << 31), "augment_batch: batch too large for one launch");
#ifdef ZS3_HOST_EMULATION
  (void)stream; (void)smem_rows; (void)smem_cols;
  for (int b = 0; b < 2 * a->n; ++b) emul_run_block<AugP>(aug_plan_kernel, p, AUG_THREADS, b, 2 * a->n);
  for (int b = 0; b < (int)hblocks; ++b) emul_run_block<AugP>(aug_hpass_kernel, p, AUG_THREADS, b, (int)hblocks);
  for (int b = 0; b < (int)vblocks; ++b) emul_run_block<AugP>(aug_vpass_kernel, p, AUG_THREADS, b, (int)vblocks);
  if (any_blur) {
    for (int b = 0; b < a->n * a->out_h; ++b) emul_run_block<AugP>(aug_blur_rows_kernel, p, AUG_THREADS, b, a->n * a->out_h);
    for (int b = 0; b < a->n * strips; ++b) emul_run_block<AugP>(aug_blur_cols_kernel, p, AUG_THREADS, b, a->n * strips);
  }
  return ZS3_OK;
#else
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(aug_blur_cols_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) !=
        cudaSuccess) {
      cudaGetLastError();
      zs3::set_error("augment_batch: cannot raise the dynamic shared-memory limit");
      return ZS3_ERR_DRIVER;
    }
    attr_set = true;
  }
  aug_plan_kernel<<<2 * a->n, AUG_THREADS, 0, st>>>(p);
  ZS3_CHECK_LAUNCH("augment_batch(plan)");
  aug_hpass_kernel<<<(int)hblocks, AUG_THREADS, 0, st>>>(p);
  ZS3_CHECK_LAUNCH("augment_batch(horizontal pass)");
  aug_vpass_kernel<<<(int)vblocks, AUG_THREADS, 0, st>>>(p);
  ZS3_CHECK_LAUNCH("augment_batch(vertical pass)");
  if (any_blur) {
    aug_blur_rows_kernel<<<a->n * a->out_h, AUG_THREADS, smem_rows, st>>>(p);
    ZS3_CHECK_LAUNCH("augment_batch(blur rows)");
    aug_blur_cols_kernel<<<a->n * strips, AUG_THREADS, smem_cols, st>>>(p);
    ZS3_CHECK_LAUNCH("augment_batch(blur columns)");
  }
  return ZS3_OK;
#endif
}
