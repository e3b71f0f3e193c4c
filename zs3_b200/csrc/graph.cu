// Semantic-cluster graph of a label map on the device (BASELINE configs[4]): replaces the
// pure-Python depth-first search of construct_adj_mat (zs3/train_context_GMMN_GCNcontext.py:33-102, ~0.5 s per
// 129x129 image on the host plus three device->host copies, `:315-320`).
//
// What the reference computes, restated: pixels are visited in raster order; an unvisited pixel seeds a new cluster
// (node id = number of clusters found so far) that is grown over the 8-connected pixels carrying the same label
// (`:55-77`; 255 is a label like any other); the node's label, embedding and feature are those of the SEED pixel
// (`:58-61,72-74`); two nodes are joined by an (undirected, unit-weight) edge when any of their pixels are
// 8-neighbours (`:78-88`; neighbouring nodes necessarily differ in label).  Hence:
//   node id   = rank of the component's smallest flat pixel index among all components' smallest indices,
//   adjacency = OR over 8-neighbour pixel pairs with different node ids.
// One thread block per image: labels and the union-find forest live in shared memory (2 x 4 B per pixel), roots are
// hooked with atomicMin so that every component's root ends up being its smallest pixel index, node ids come from a
// block-wide prefix sum over the root flags.  Integer work, bit-exact against the oracle.
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

constexpr int CC_THREADS = 1024;

struct CompP {
  const float* labels;
  const int* src_index;
  long long image_stride;
  int h, w, max_nodes;
  int* n_nodes;
  int* node_label;
  int* node_seed;
  int* node_map;
  float* adj;
};

__device__ __forceinline__ int cc_find(const volatile int* parent, int x) {
  int p;
  while ((p = parent[x]) != x) x = p;
  return x;
}

// hook the larger root under the smaller one (lock-free; the smallest index of a component ends up as its root)
__device__ __forceinline__ void cc_unite(int* parent, int a, int b) {
  while (true) {
    a = cc_find(parent, a);
    b = cc_find(parent, b);
    if (a == b) return;
    if (a < b) {
      const int t = a;
      a = b;
      b = t;
    }
    const int old = atomicMin(&parent[a], b);
    if (old == a) return;
    a = old;  // `a` had been hooked elsewhere in the meantime: its former parent still has to meet b
  }
}

#ifdef ZS3_HOST_EMULATION
static int cc_smem_emul[2 * 40000 + 2 * CC_THREADS];
#endif

__global__ void __launch_bounds__(CC_THREADS) label_components_kernel(const CompP p) {
#ifdef ZS3_HOST_EMULATION
  int* smem = cc_smem_emul;
#else
  extern __shared__ int smem[];
#endif
  const int h = p.h, w = p.w, npix = h * w;
  int* lab = smem;               // label of every pixel; later: node id stored at the root pixels
  int* parent = smem + npix;     // union-find forest; later: node id of every pixel
  int* scan = smem + 2 * npix;   // [2][CC_THREADS]
  const int b = blockIdx.x, tid = threadIdx.x;
  const float* src = p.labels + (long long)b * p.image_stride;

  for (int q = tid; q < npix; q += CC_THREADS) {
    lab[q] = (int)__ldg(src + (p.src_index ? __ldg(p.src_index + q) : q));
    parent[q] = q;
  }
  float* adj = p.adj + (long long)b * p.max_nodes * p.max_nodes;
  for (int q = tid; q < p.max_nodes * p.max_nodes; q += CC_THREADS) adj[q] = 0.f;
  __syncthreads();

  // horizontal runs first: link every pixel to its left neighbour of equal label, then pointer-jump until each
  // pixel points at the first pixel of its run (ceil(log2 w) rounds; unsynchronised reads only ever see a pointer
  // that is already further up the same chain).  The unions below then start from trees of depth <= 1 instead
  // of w-long chains.
  for (int q = tid; q < npix; q += CC_THREADS) {
    const int j = q % w;
    if (j > 0 && lab[q - 1] == lab[q]) parent[q] = q - 1;
  }
  __syncthreads();
  for (int span = 1; span < w; span <<= 1) {
    for (int q = tid; q < npix; q += CC_THREADS) {
      const volatile int* vp = parent;
      parent[q] = vp[vp[q]];
    }
    __syncthreads();
  }
  // union with the previous row's part of the 8-neighbourhood (NW, N, NE).  Runs are already connected, so one
  // union per maximal overlap of two runs suffices; a pixel skips the unions that a row neighbour performs:
  //   N  is implied when the left neighbour and NW carry the label too (the left pixel's N-union + the two runs),
  //   NW / NE only matter when N differs, and are implied when the left / right neighbour carries the label
  //   (that pixel sees the same upper-row pixel as its own N).
  for (int q = tid; q < npix; q += CC_THREADS) {
    const int i = q / w, j = q - i * w, l = lab[q];
    if (i == 0) continue;
    const bool left = j > 0 && lab[q - 1] == l, right = j + 1 < w && lab[q + 1] == l;
    const bool up = lab[q - w] == l;
    const bool upl = j > 0 && lab[q - w - 1] == l, upr = j + 1 < w && lab[q - w + 1] == l;
    if (up) {
      if (!(left && upl)) cc_unite(parent, q, q - w);
    } else {
      if (upl && !left) cc_unite(parent, q, q - w - 1);
      if (upr && !right) cc_unite(parent, q, q - w + 1);
    }
  }
  __syncthreads();
  // flatten: a few pointer-jumping rounds take the bulk of the depth out, the exact walk finishes
  for (int r = 0; r < 6; ++r) {
    for (int q = tid; q < npix; q += CC_THREADS) {
      const volatile int* vp = parent;
      parent[q] = vp[vp[q]];
    }
    __syncthreads();
  }
  for (int q = tid; q < npix; q += CC_THREADS) parent[q] = cc_find(parent, q);
  __syncthreads();

  // node ids: exclusive prefix sum of the root flags in raster order (contiguous chunk per thread)
  const int chunk = (npix + CC_THREADS - 1) / CC_THREADS;
  const int q0 = min(tid * chunk, npix), q1 = min(q0 + chunk, npix);
  int cnt = 0;
  for (int q = q0; q < q1; ++q) cnt += (parent[q] == q);
  scan[tid] = cnt;
  __syncthreads();
  int cur = 0;
  for (int off = 1; off < CC_THREADS; off <<= 1) {  // Hillis-Steele inclusive scan, double buffered
    const int v = scan[cur * CC_THREADS + tid] + (tid >= off ? scan[cur * CC_THREADS + tid - off] : 0);
    scan[(cur ^ 1) * CC_THREADS + tid] = v;
    cur ^= 1;
    __syncthreads();
  }
  const int total = scan[cur * CC_THREADS + CC_THREADS - 1];
  int id = scan[cur * CC_THREADS + tid] - cnt;  // exclusive
  int* node_label = p.node_label + (long long)b * p.max_nodes;
  int* node_seed = p.node_seed + (long long)b * p.max_nodes;
  for (int q = q0; q < q1; ++q) {
    if (parent[q] == q) {
      if (id < p.max_nodes) {
        node_label[id] = lab[q];
        node_seed[id] = q;
      }
      lab[q] = id;  // the label of a root pixel has been emitted: reuse the slot for its node id
      ++id;
    }
  }
  if (tid == 0) p.n_nodes[b] = total;
  __syncthreads();
  for (int q = tid; q < npix; q += CC_THREADS) {
    const int node = lab[parent[q]];
    if (p.node_map) p.node_map[(long long)b * npix + q] = node;
    parent[q] = node;  // own entry only; roots are read through lab[]
  }
  __syncthreads();

  // edges: forward half of the 8-neighbourhood (E, SW, S, SE); identical racing stores are benign
  const int mn = p.max_nodes;
  for (int q = tid; q < npix; q += CC_THREADS) {
    const int i = q / w, j = q - i * w, a = parent[q];
    int nb[4];
    nb[0] = (j + 1 < w) ? parent[q + 1] : a;
    nb[1] = (i + 1 < h && j > 0) ? parent[q + w - 1] : a;
    nb[2] = (i + 1 < h) ? parent[q + w] : a;
    nb[3] = (i + 1 < h && j + 1 < w) ? parent[q + w + 1] : a;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int c = nb[e];
      if (c != a && a < mn && c < mn) {
        adj[(long long)a * mn + c] = 1.f;
        adj[(long long)c * mn + a] = 1.f;
      }
    }
  }
}

}  // namespace zs3

using namespace zs3;

#ifdef ZS3_HOST_EMULATION
extern "C" int zs3_emul_label_components(const zs3_components_args* a, void* stream) {
#else
extern "C" int zs3_label_components(const zs3_components_args* a, void* stream) {
#endif
  ZS3_CHECK_ARG(a != nullptr, "label_components: null args");
  ZS3_CHECK_ARG(a->labels && a->n_nodes && a->node_label && a->node_seed && a->adj, "label_components: null pointer");
  ZS3_CHECK_ARG(a->B >= 0 && a->h > 0 && a->w > 0 && a->max_nodes > 0, "label_components: bad dims");
  const long long npix = (long long)a->h * a->w;
  const size_t smem = sizeof(int) * (2 * (size_t)npix + 2 * CC_THREADS);
  ZS3_CHECK_ARG(smem <= 220 * 1024, "label_components: %d x %d label map does not fit in shared memory", a->h, a->w);
  if (a->B == 0) return ZS3_OK;
  CompP p;
  p.labels = a->labels; p.src_index = a->src_index; p.image_stride = a->image_stride;
  p.h = a->h; p.w = a->w; p.max_nodes = a->max_nodes;
  p.n_nodes = a->n_nodes; p.node_label = a->node_label; p.node_seed = a->node_seed; p.node_map = a->node_map;
  p.adj = a->adj;
#ifdef ZS3_HOST_EMULATION
  (void)stream;
  if (npix > 40000) return -1;
  for (int b = 0; b < a->B; ++b) emul_run_block<CompP>(label_components_kernel, p, CC_THREADS, b, a->B);
  return ZS3_OK;
#else
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(label_components_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024) !=
        cudaSuccess) {
      cudaGetLastError();
      zs3::set_error("label_components: cannot raise the dynamic shared-memory limit");
      return ZS3_ERR_DRIVER;
    }
    attr_set = true;
  }
  label_components_kernel<<<a->B, CC_THREADS, smem, static_cast<cudaStream_t>(stream)>>>(p);
  ZS3_CHECK_LAUNCH("label_components");
  return ZS3_OK;
#endif
}
