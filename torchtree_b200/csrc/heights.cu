// Node-height reparameterisation on the device.
//
// Replaces the Python loop of GeneralNodeHeightTransform._call
// (torchtree/evolution/tree_height_transform.py:58-66: T-2 dependent in-place tensor
// updates per evaluation, taped by autograd) and its backward:
//     h[root] = x[root],   h[c] = b[c] + x[c] (h[parent(c)] - b[c])
// for every internal node c, where b[c] is the latest sampling time below c
// (update_bounds, :36-56) and x holds ratios in (0,1) plus the root height.
//
// One CTA per draw walks the tree level by level (levels = height above the tips,
// the engine's post-order schedule read backwards): the nodes of one level are
// independent, __syncthreads() separates levels, the output array itself is the
// working storage.  The backward pass is the exact reverse sweep
//     t[n] = g[n] + sum_{internal children c} t[c] x[c],   dx[c] = t[c] (h[n] - b[c]),   dx[root] = t[root]
// parent-centric, so every sum has a fixed order (deterministic, no atomics).
#include <algorithm>
#include <vector>

#include "engine.cuh"

struct ttb2_heights {
  int T = 0, I = 0, device = 0;
  int nLevels = 0;
  int* ops = nullptr;        // [I][3] (node, left, right) sorted by level, device
  int* levelOff = nullptr;   // [nLevels + 1], device
  double* bounds = nullptr;  // [I], device
  double *x = nullptr, *h = nullptr, *g = nullptr, *gx = nullptr;  // staging for host callers
  int stagedDraws = 0;
  cudaStream_t stream = nullptr;
};

namespace ttb2 {
namespace {

constexpr int HT_THREADS = 256;

__global__ void __launch_bounds__(HT_THREADS)
heights_fwd_kernel(const int* __restrict__ ops, const int* __restrict__ levelOff, int nLevels,
                   const double* __restrict__ bounds, const double* __restrict__ x,
                   double* __restrict__ h, int T, int I) {
  const double* xd = x + (size_t)blockIdx.x * I;
  double* hd = h + (size_t)blockIdx.x * I;
  if (threadIdx.x == 0) {
    const int root = ops[(levelOff[nLevels] - 1) * 3];   // the top level holds the root alone
    hd[root - T] = xd[root - T];
  }
  __syncthreads();
  for (int l = nLevels - 1; l >= 0; --l) {
    for (int j = levelOff[l] + threadIdx.x; j < levelOff[l + 1]; j += blockDim.x) {
      const int n = ops[j * 3] - T;
      const double hn = hd[n];
#pragma unroll
      for (int s = 1; s <= 2; ++s) {
        const int c = ops[j * 3 + s] - T;
        if (c >= 0) hd[c] = fma(xd[c], hn - bounds[c], bounds[c]);
      }
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(HT_THREADS)
heights_bwd_kernel(const int* __restrict__ ops, const int* __restrict__ levelOff, int nLevels,
                   const double* __restrict__ bounds, const double* __restrict__ x,
                   const double* __restrict__ h, const double* __restrict__ g,
                   double* __restrict__ gx, int T, int I) {
  const double* xd = x + (size_t)blockIdx.x * I;
  const double* hd = h + (size_t)blockIdx.x * I;
  const double* gd = g + (size_t)blockIdx.x * I;
  double* t = gx + (size_t)blockIdx.x * I;
  for (int l = 0; l < nLevels; ++l) {
    for (int j = levelOff[l] + threadIdx.x; j < levelOff[l + 1]; j += blockDim.x) {
      const int n = ops[j * 3] - T;
      const double hn = hd[n];
      double tn = gd[n];
#pragma unroll
      for (int s = 1; s <= 2; ++s) {
        const int c = ops[j * 3 + s] - T;
        if (c >= 0) {
          const double tc = t[c];              // finished at a lower level
          tn = fma(tc, xd[c], tn);
          t[c] = tc * (hn - bounds[c]);        // d/dx[c]
        }
      }
      t[n] = tn;                               // the root keeps it: dx[root] = t[root]
    }
    __syncthreads();
  }
}

int stage(ttb2_heights& p, int draws) {
  if (draws <= p.stagedDraws) return TTB2_OK;
  for (double** q : {&p.x, &p.h, &p.g, &p.gx}) {
    if (*q) cudaFree(*q);
    *q = nullptr;
    TTB2_CUDA_CHECK(cudaMalloc((void**)q, (size_t)draws * p.I * sizeof(double)));
  }
  p.stagedDraws = draws;
  return TTB2_OK;
}

}  // namespace
}  // namespace ttb2

using ttb2::set_error;

extern "C" int ttb2_heights_create(int32_t tip_count, const int32_t* postorder,
                                   const double* bounds, int32_t device, ttb2_heights** out) {
  if (!postorder || !bounds || !out || tip_count < 2) {
    set_error("ttb2_heights_create: null argument or fewer than 2 tips");
    return TTB2_E_INVALID;
  }
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) {
    cudaGetLastError();
    set_error("ttb2_heights_create: no CUDA device available (this library has no CPU fallback)");
    return TTB2_E_CUDA;
  }
  if (device < 0 || device >= count) {
    set_error("ttb2_heights_create: device index out of range");
    return TTB2_E_INVALID;
  }
  const int T = tip_count, I = T - 1;
  // level(n) = 1 + max(level(children)), tips = 0; validate the triples on the way
  std::vector<int> level(2 * T - 1, -1);
  for (int t = 0; t < T; ++t) level[t] = 0;
  int maxLevel = 0;
  for (int j = 0; j < I; ++j) {
    const int n = postorder[j * 3], l = postorder[j * 3 + 1], r = postorder[j * 3 + 2];
    if (n < T || n > 2 * T - 2 || l < 0 || r < 0 || l > 2 * T - 2 || r > 2 * T - 2 ||
        level[l] < 0 || level[r] < 0 || level[n] >= 0) {
      set_error("ttb2_heights_create: postorder is not a valid post-order list of (node, left, right)");
      return TTB2_E_INVALID;
    }
    level[n] = 1 + std::max(level[l], level[r]);
    maxLevel = std::max(maxLevel, level[n]);
  }
  std::vector<int> order(I);
  for (int j = 0; j < I; ++j) order[j] = j;
  std::stable_sort(order.begin(), order.end(), [&](int a, int b) {
    return level[postorder[a * 3]] < level[postorder[b * 3]];
  });
  std::vector<int> ops(3 * I), off(maxLevel + 1, 0);
  for (int j = 0; j < I; ++j) {
    for (int s = 0; s < 3; ++s) ops[j * 3 + s] = postorder[order[j] * 3 + s];
    off[level[ops[j * 3]]] = j + 1;   // levels 1..maxLevel -> off[1..maxLevel]
  }
  for (int l = 1; l <= maxLevel; ++l)
    if (off[l] < off[l - 1]) off[l] = off[l - 1];
  if (off[maxLevel] - off[maxLevel - 1] != 1) {
    set_error("ttb2_heights_create: the last post-order entry must be the root");
    return TTB2_E_INVALID;
  }
  auto* p = new ttb2_heights();
  p->T = T; p->I = I; p->device = device; p->nLevels = maxLevel;
  auto fail = [&](int rc) { ttb2_heights_destroy(p); return rc; };
#define HT_TRY(expr)                                                                    \
  do {                                                                                  \
    cudaError_t err__ = (expr);                                                         \
    if (err__ != cudaSuccess) {                                                         \
      set_error(std::string(#expr) + ": " + cudaGetErrorString(err__));                 \
      return fail(TTB2_E_CUDA);                                                         \
    }                                                                                   \
  } while (0)
  HT_TRY(cudaSetDevice(device));
  HT_TRY(cudaStreamCreateWithFlags(&p->stream, cudaStreamNonBlocking));
  HT_TRY(cudaMalloc((void**)&p->ops, ops.size() * sizeof(int)));
  HT_TRY(cudaMalloc((void**)&p->levelOff, off.size() * sizeof(int)));
  HT_TRY(cudaMalloc((void**)&p->bounds, (size_t)I * sizeof(double)));
  HT_TRY(cudaMemcpy(p->ops, ops.data(), ops.size() * sizeof(int), cudaMemcpyHostToDevice));
  HT_TRY(cudaMemcpy(p->levelOff, off.data(), off.size() * sizeof(int), cudaMemcpyHostToDevice));
  HT_TRY(cudaMemcpy(p->bounds, bounds, (size_t)I * sizeof(double), cudaMemcpyHostToDevice));
#undef HT_TRY
  *out = p;
  return TTB2_OK;
}

extern "C" int ttb2_heights_destroy(ttb2_heights* p) {
  if (!p) return TTB2_OK;
  cudaSetDevice(p->device);
  for (void* q : {(void*)p->ops, (void*)p->levelOff, (void*)p->bounds, (void*)p->x, (void*)p->h,
                  (void*)p->g, (void*)p->gx})
    if (q) cudaFree(q);
  if (p->stream) cudaStreamDestroy(p->stream);
  delete p;
  return TTB2_OK;
}

extern "C" int ttb2_heights_forward(ttb2_heights* p, int32_t draws, const double* x,
                                    double* heights, int32_t where) {
  if (!p || !x || !heights || draws < 1) {
    set_error("ttb2_heights_forward: null argument or draws < 1");
    return TTB2_E_INVALID;
  }
  TTB2_CUDA_CHECK(cudaSetDevice(p->device));
  const size_t bytes = (size_t)draws * p->I * sizeof(double);
  const double* dx = x;
  double* dh = heights;
  if (where == TTB2_HOST) {
    int rc = ttb2::stage(*p, draws);
    if (rc) return rc;
    TTB2_CUDA_CHECK(cudaMemcpyAsync(p->x, x, bytes, cudaMemcpyHostToDevice, p->stream));
    dx = p->x;
    dh = p->h;
  }
  ttb2::heights_fwd_kernel<<<draws, ttb2::HT_THREADS, 0, p->stream>>>(
      p->ops, p->levelOff, p->nLevels, p->bounds, dx, dh, p->T, p->I);
  TTB2_CUDA_CHECK(cudaGetLastError());
  if (where == TTB2_HOST)
    TTB2_CUDA_CHECK(cudaMemcpyAsync(heights, p->h, bytes, cudaMemcpyDeviceToHost, p->stream));
  TTB2_CUDA_CHECK(cudaStreamSynchronize(p->stream));
  return TTB2_OK;
}

extern "C" int ttb2_heights_backward(ttb2_heights* p, int32_t draws, const double* x,
                                     const double* heights, const double* grad_heights,
                                     double* grad_x, int32_t where) {
  if (!p || !x || !heights || !grad_heights || !grad_x || draws < 1) {
    set_error("ttb2_heights_backward: null argument or draws < 1");
    return TTB2_E_INVALID;
  }
  TTB2_CUDA_CHECK(cudaSetDevice(p->device));
  const size_t bytes = (size_t)draws * p->I * sizeof(double);
  const double *dx = x, *dh = heights, *dg = grad_heights;
  double* dgx = grad_x;
  if (where == TTB2_HOST) {
    int rc = ttb2::stage(*p, draws);
    if (rc) return rc;
    TTB2_CUDA_CHECK(cudaMemcpyAsync(p->x, x, bytes, cudaMemcpyHostToDevice, p->stream));
    TTB2_CUDA_CHECK(cudaMemcpyAsync(p->h, heights, bytes, cudaMemcpyHostToDevice, p->stream));
    TTB2_CUDA_CHECK(cudaMemcpyAsync(p->g, grad_heights, bytes, cudaMemcpyHostToDevice, p->stream));
    dx = p->x; dh = p->h; dg = p->g; dgx = p->gx;
  }
  ttb2::heights_bwd_kernel<<<draws, ttb2::HT_THREADS, 0, p->stream>>>(
      p->ops, p->levelOff, p->nLevels, p->bounds, dx, dh, dg, dgx, p->T, p->I);
  TTB2_CUDA_CHECK(cudaGetLastError());
  if (where == TTB2_HOST)
    TTB2_CUDA_CHECK(cudaMemcpyAsync(grad_x, p->gx, bytes, cudaMemcpyDeviceToHost, p->stream));
  TTB2_CUDA_CHECK(cudaStreamSynchronize(p->stream));
  return TTB2_OK;
}
