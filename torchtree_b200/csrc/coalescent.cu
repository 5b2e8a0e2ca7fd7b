// Constant-population coalescent log-density of a batch of time trees on the device
//  -- replaces ConstantCoalescent.log_prob,
// torchtree/evolution/coalescent.py:112-134:
//     sort the 2T-1 node heights; lineages k_i after event i = cumsum(+1 tip / -1 coalescence);
//     log p = - sum_i C(k_i, 2) (s_{i+1} - s_i) / theta - (T - 1) log theta
// and its autograd backward, which has a closed form once the sorted order is known:
//     d log p / d s_j    = - (C(k_{j-1}, 2) - C(k_j, 2)) / theta      (C(k_{-1}) = C(k_{n-1}) = 0)
//     d log p / d theta  =   sum_i C(k_i, 2) (s_{i+1} - s_i) / theta^2 - (T - 1) / theta
// One CTA per draw: bitonic sort of (height, node index) pairs in shared memory (ties broken
// by index, i.e. like a stable sort: intervals between tied events have zero length and do
// not contribute), an integer scan for the lineage counts, and a fixed-order block reduction
// (bit-wise reproducible).  Up to 4096 tips the 16 bytes per node live in shared memory; larger
// trees run the same kernel on a global-memory scratch area (any T that fits in memory).
#include <climits>
#include <math_constants.h>
#include <string>

#include "engine.cuh"

namespace ttb2 {

namespace {

constexpr int CO_THREADS = 256;

__global__ void __launch_bounds__(CO_THREADS)
coalescent_constant_kernel(const double* __restrict__ heights, const double* __restrict__ theta,
                           int thetaDraws, double* __restrict__ logp,
                           double* __restrict__ dHeights, double* __restrict__ dTheta, int T,
                           int np2, double* __restrict__ scratch) {
  extern __shared__ double sm[];
  const int n = 2 * T - 1;
  // working arrays: shared memory, or this draw's slice of the global scratch (large trees)
  double* key = scratch ? scratch + (size_t)blockIdx.x * np2 * 2 : sm;   // [np2]
  int* idx = reinterpret_cast<int*>(key + np2);           // [np2]
  int* cnt = idx + np2;                                   // [np2] lineage counts k_i
  __shared__ int chunkSum[CO_THREADS];
  __shared__ double red[CO_THREADS / 32];
  __shared__ double total;
  __shared__ int bad;
  const int d = blockIdx.x, tid = threadIdx.x;
  const double* h = heights + (size_t)d * n;
  if (tid == 0) bad = 0;
  __syncthreads();
  for (int i = tid; i < np2; i += CO_THREADS) {
    const double v = i < n ? h[i] : CUDART_INF;
    if (i < n && !(fabs(v) < CUDART_INF)) bad = 1;   // NaN or infinite height
    key[i] = v;
    idx[i] = i < n ? i : INT_MAX;
  }
  __syncthreads();
  if (bad) {
    // NaN in -> NaN out (a NaN has no place in the sorted order; the reference's log_prob is
    // NaN as well), for the value and every derivative
    if (tid == 0) {
      logp[d] = CUDART_NAN;
      if (dTheta) dTheta[d] = CUDART_NAN;
    }
    if (dHeights)
      for (int j = tid; j < n; j += CO_THREADS) dHeights[(size_t)d * n + j] = CUDART_NAN;
    return;
  }
  // bitonic sort, ascending in (height, index)
  for (int k = 2; k <= np2; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = tid; i < np2; i += CO_THREADS) {
        const int p = i ^ j;
        if (p > i) {
          const double a = key[i], b = key[p];
          const int ia = idx[i], ib = idx[p];
          const bool gt = a > b || (a == b && ia > ib);
          if (gt == ((i & k) == 0)) {
            key[i] = b; key[p] = a;
            idx[i] = ib; idx[p] = ia;
          }
        }
      }
      __syncthreads();
    }
  }
  // lineage counts: inclusive scan of +1 (tip) / -1 (coalescence) over the sorted events
  const int per = (np2 + CO_THREADS - 1) / CO_THREADS;
  const int lo = tid * per, hi = min(lo + per, n);
  int s = 0;
  for (int i = lo; i < hi; ++i) s += idx[i] < T ? 1 : -1;
  chunkSum[tid] = s;
  __syncthreads();
  int base = 0;
  for (int t = 0; t < tid; ++t) base += chunkSum[t];
  for (int i = lo; i < hi; ++i) {
    base += idx[i] < T ? 1 : -1;
    cnt[i] = base;
  }
  __syncthreads();
  // sum_i C(k_i, 2) (s_{i+1} - s_i), i = 0 .. n-2
  double part = 0.0;
  for (int i = tid; i < n - 1; i += CO_THREADS) {
    const double k = (double)cnt[i];
    part += 0.5 * k * (k - 1.0) * (key[i + 1] - key[i]);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
  if ((tid & 31) == 0) red[tid >> 5] = part;
  __syncthreads();
  if (tid == 0) {
    double t = 0.0;
    for (int w = 0; w < CO_THREADS / 32; ++w) t += red[w];
    total = t;
  }
  __syncthreads();
  const double th = theta[thetaDraws == 1 ? 0 : d];
  if (tid == 0) {
    logp[d] = -total / th - (double)(T - 1) * log(th);
    if (dTheta) dTheta[d] = total / (th * th) - (double)(T - 1) / th;
  }
  if (dHeights) {
    double* g = dHeights + (size_t)d * n;
    for (int j = tid; j < n; j += CO_THREADS) {
      double cPrev = 0.0, cHere = 0.0;
      if (j > 0) {
        const double k = (double)cnt[j - 1];
        cPrev = 0.5 * k * (k - 1.0);
      }
      if (j < n - 1) {
        const double k = (double)cnt[j];
        cHere = 0.5 * k * (k - 1.0);
      }
      g[idx[j]] = -(cPrev - cHere) / th;
    }
  }
}

// per-thread staging buffers for host-pointer calls
struct CoScratch {
  int device = -1;
  size_t cap = 0;
  double* buf = nullptr;
};
thread_local CoScratch co_scratch;

}  // namespace
}  // namespace ttb2

extern "C" int ttb2_coalescent_constant(int32_t device, int32_t draws, int32_t tip_count,
                                        const double* node_heights, const double* theta,
                                        int32_t theta_draws, double* log_prob, double* d_heights,
                                        double* d_theta, int32_t where) {
  using namespace ttb2;
  if (!node_heights || !theta || !log_prob || draws < 1 || tip_count < 2 ||
      (theta_draws != 1 && theta_draws != draws)) {
    set_error("ttb2_coalescent_constant: null argument, draws < 1, fewer than 2 tips, or a "
              "theta_draws that is neither 1 nor draws");
    return TTB2_E_INVALID;
  }
  const int n = 2 * tip_count - 1;
  int np2 = 1;
  while (np2 < n) np2 <<= 1;
  const bool big = np2 > 8192;   // beyond 4096 tips the working arrays do not fit in shared memory
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    set_error("ttb2_coalescent_constant: no CUDA device available (sm_100a required; there is no CPU fallback)");
    return TTB2_E_CUDA;
  }
  if (device < 0 || device >= ndev) {
    set_error("ttb2_coalescent_constant: device index out of range");
    return TTB2_E_INVALID;
  }
  TTB2_CUDA_CHECK(cudaSetDevice(device));
  const size_t nh = (size_t)draws * n;
  const double* dH = node_heights;
  const double* dT = theta;
  double *dL = log_prob, *dGH = d_heights, *dGT = d_theta;
  if (where == TTB2_HOST) {
    // layout: heights | theta | logp | d_heights | d_theta | [sort scratch of large trees]
    const size_t need = nh + theta_draws + draws + nh + draws + (big ? (size_t)draws * np2 * 2 : 0);
    CoScratch& sc = co_scratch;
    if (sc.device != device || sc.cap < need) {
      if (sc.buf) {
        cudaSetDevice(sc.device);
        cudaFree(sc.buf);
        cudaSetDevice(device);
        sc.buf = nullptr;
        sc.cap = 0;
      }
      TTB2_CUDA_CHECK(cudaMalloc((void**)&sc.buf, need * sizeof(double)));
      sc.cap = need;
      sc.device = device;
    }
    double* b = sc.buf;
    TTB2_CUDA_CHECK(cudaMemcpyAsync(b, node_heights, nh * sizeof(double), cudaMemcpyHostToDevice, 0));
    TTB2_CUDA_CHECK(cudaMemcpyAsync(b + nh, theta, theta_draws * sizeof(double), cudaMemcpyHostToDevice, 0));
    dH = b;
    dT = b + nh;
    dL = b + nh + theta_draws;
    dGH = d_heights ? dL + draws : nullptr;
    dGT = d_theta ? dL + draws + nh : nullptr;
  }
  double* sortScratch = nullptr;
  if (big) {
    if (where == TTB2_HOST) {
      sortScratch = co_scratch.buf + (nh + theta_draws + draws + nh + draws);
    } else {
      TTB2_CUDA_CHECK(cudaMalloc((void**)&sortScratch, (size_t)draws * np2 * 2 * sizeof(double)));
    }
  }
  const size_t smem = big ? 0 : (size_t)np2 * (sizeof(double) + 2 * sizeof(int));
  if (smem > 48 * 1024)
    TTB2_CUDA_CHECK(cudaFuncSetAttribute(coalescent_constant_kernel,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  coalescent_constant_kernel<<<draws, CO_THREADS, smem, 0>>>(dH, dT, theta_draws, dL, dGH, dGT,
                                                            tip_count, np2, sortScratch);
  {
    const cudaError_t lerr = cudaGetLastError();
    if (lerr != cudaSuccess) {
      if (big && where != TTB2_HOST) cudaFree(sortScratch);
      set_error(std::string("coalescent_constant_kernel: ") + cudaGetErrorString(lerr));
      return TTB2_E_CUDA;
    }
  }
  if (big && where != TTB2_HOST) {
    TTB2_CUDA_CHECK(cudaStreamSynchronize(0));
    cudaFree(sortScratch);
  }
  if (where == TTB2_HOST) {
    TTB2_CUDA_CHECK(cudaMemcpyAsync(log_prob, dL, draws * sizeof(double), cudaMemcpyDeviceToHost, 0));
    if (d_heights)
      TTB2_CUDA_CHECK(cudaMemcpyAsync(d_heights, dGH, nh * sizeof(double), cudaMemcpyDeviceToHost, 0));
    if (d_theta)
      TTB2_CUDA_CHECK(cudaMemcpyAsync(d_theta, dGT, draws * sizeof(double), cudaMemcpyDeviceToHost, 0));
  }
  TTB2_CUDA_CHECK(cudaStreamSynchronize(0));
  return TTB2_OK;
}
