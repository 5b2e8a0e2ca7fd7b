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

// =============================================================================================
// Piecewise-constant coalescents on the device: skyride (one population size per
// inter-coalescent interval) and skygrid (one per segment of a fixed time grid) -- replace
// PiecewiseConstantCoalescent.log_prob (torchtree/evolution/coalescent.py:311-396) and
// PiecewiseConstantCoalescentGrid.log_prob (:459-549) with their autograd backward.
//
// Same structure as the constant kernel: one CTA per draw sorts the events -- the 2T-1 node
// heights plus, for the skygrid, the G grid points (mask 0: no lineage change) -- scans the
// lineage counts k_i and the population-size index of every interval
//     skyride: #coalescent events among events 0..i      (:379-388, cumsum(...)[..., :-1])
//     skygrid: #grid points among events 0..i            (:527-531)
// and returns
//     log p = - sum_i C(k_i, 2) (s_{i+1} - s_i) / theta[ix_i] - sum_{log terms} log theta[.]
// (skyride: every theta once, :393-395; skygrid: theta[ix_j] once per coalescent event j, :541-548)
// with d/ds_j = -(C(k_{j-1},2)/theta[ix_{j-1}] - C(k_j,2)/theta[ix_j]) for the node events and
// d/dtheta_m = sum_{i: ix_i = m} C(k_i,2) (s_{i+1}-s_i) / theta_m^2 - (#log terms of m) / theta_m.
// ix is non-decreasing along the sorted events, so every theta_m owns a contiguous run of
// intervals: one thread per m sums its run in order (deterministic, no atomics).
// =============================================================================================
namespace ttb2 {
namespace {

__global__ void __launch_bounds__(CO_THREADS)
coalescent_piecewise_kernel(const double* __restrict__ heights, const double* __restrict__ grid,
                            const double* __restrict__ theta, int thetaDraws,
                            double* __restrict__ logp, double* __restrict__ dHeights,
                            double* __restrict__ dTheta, int T, int G, int M, int np2,
                            double* __restrict__ scratch) {
  extern __shared__ double sm[];
  const int n = 2 * T - 1;      // node events
  const int ne = n + G;         // all events
  // working arrays (20 bytes per event): shared memory or this draw's slice of the scratch area
  double* key = scratch ? scratch + (size_t)blockIdx.x * np2 * 3 : sm;   // [np2]
  int* idx = reinterpret_cast<int*>(key + np2);                           // [np2]
  int* cnt = idx + np2;                                                   // [np2] lineage counts
  int* tix = cnt + np2;                                                   // [np2] theta index
  __shared__ int chunkSumK[CO_THREADS], chunkSumT[CO_THREADS];
  __shared__ double red[CO_THREADS / 32];
  __shared__ double total;
  __shared__ int bad;
  const int d = blockIdx.x, tid = threadIdx.x;
  const double* h = heights + (size_t)d * n;
  const double* th = theta + (size_t)(thetaDraws == 1 ? 0 : d) * M;
  const bool skygrid = G > 0;
  if (tid == 0) bad = 0;
  __syncthreads();
  for (int i = tid; i < np2; i += CO_THREADS) {
    const double v = i < n ? h[i] : (i < ne ? grid[i - n] : CUDART_INF);
    if (i < ne && !(fabs(v) < CUDART_INF)) bad = 1;
    key[i] = v;
    idx[i] = i < ne ? i : INT_MAX;
  }
  __syncthreads();
  if (bad) {
    if (tid == 0) logp[d] = CUDART_NAN;
    if (dTheta)
      for (int m = tid; m < M; m += CO_THREADS) dTheta[(size_t)d * M + m] = CUDART_NAN;
    if (dHeights)
      for (int j = tid; j < n; j += CO_THREADS) dHeights[(size_t)d * n + j] = CUDART_NAN;
    return;
  }
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
  // inclusive scans over the sorted events: lineage count and theta index
  auto dk = [&](int id) { return id < T ? 1 : (id < n ? -1 : 0); };
  auto dt = [&](int id) { return skygrid ? (id >= n ? 1 : 0) : ((id >= T && id < n) ? 1 : 0); };
  const int per = (np2 + CO_THREADS - 1) / CO_THREADS;
  const int lo = tid * per, hi = min(lo + per, ne);
  int sk = 0, st = 0;
  for (int i = lo; i < hi; ++i) {
    sk += dk(idx[i]);
    st += dt(idx[i]);
  }
  chunkSumK[tid] = sk;
  chunkSumT[tid] = st;
  __syncthreads();
  int bk = 0, bt = 0;
  for (int t = 0; t < tid; ++t) {
    bk += chunkSumK[t];
    bt += chunkSumT[t];
  }
  for (int i = lo; i < hi; ++i) {
    bk += dk(idx[i]);
    bt += dt(idx[i]);
    cnt[i] = bk;
    tix[i] = bt < M ? bt : M - 1;
  }
  __syncthreads();
  // value: interval terms + log terms
  double part = 0.0;
  for (int i = tid; i < ne - 1; i += CO_THREADS) {
    const double k = (double)cnt[i];
    part -= 0.5 * k * (k - 1.0) * (key[i + 1] - key[i]) / th[tix[i]];
  }
  if (skygrid) {
    for (int j = 1 + tid; j < ne; j += CO_THREADS) {
      const int id = idx[j];
      if (id >= T && id < n) part -= log(th[tix[j]]);
    }
  } else {
    for (int m = tid; m < M; m += CO_THREADS) part -= log(th[m]);
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
  if (tid == 0) logp[d] = total;
  if (dHeights) {
    double* g = dHeights + (size_t)d * n;
    for (int j = tid; j < ne; j += CO_THREADS) {
      const int id = idx[j];
      if (id >= n) continue;   // grid points are constants
      double cPrev = 0.0, cHere = 0.0;
      if (j > 0) {
        const double k = (double)cnt[j - 1];
        cPrev = 0.5 * k * (k - 1.0) / th[tix[j - 1]];
      }
      if (j < ne - 1) {
        const double k = (double)cnt[j];
        cHere = 0.5 * k * (k - 1.0) / th[tix[j]];
      }
      g[id] = -(cPrev - cHere);
    }
  }
  if (dTheta) {
    // theta_m owns the sorted positions [first with tix >= m, first with tix >= m+1)
    for (int m = tid; m < M; m += CO_THREADS) {
      int a = 0, b = ne;   // lower bound of m in tix[0 .. ne-1]
      while (a < b) {
        const int mid = (a + b) >> 1;
        if (tix[mid] < m) a = mid + 1; else b = mid;
      }
      double s = 0.0;
      int logs = skygrid ? 0 : 1;
      for (int pos = a; pos < ne && tix[pos] == m; ++pos) {
        if (pos < ne - 1) {
          const double k = (double)cnt[pos];
          s += 0.5 * k * (k - 1.0) * (key[pos + 1] - key[pos]);
        }
        if (skygrid && pos >= 1 && idx[pos] >= T && idx[pos] < n) ++logs;   // log theta of a coalescence
      }
      const double t = th[m];
      dTheta[(size_t)d * M + m] = s / (t * t) - (double)logs / t;
    }
  }
}

}  // namespace
}  // namespace ttb2

extern "C" int ttb2_coalescent_piecewise(int32_t device, int32_t draws, int32_t tip_count,
                                         const double* node_heights, const double* theta,
                                         int32_t theta_draws, int32_t theta_count,
                                         const double* grid, int32_t grid_count, double* log_prob,
                                         double* d_heights, double* d_theta, int32_t where) {
  using namespace ttb2;
  if (!node_heights || !theta || !log_prob || draws < 1 || tip_count < 2 || grid_count < 0 ||
      (grid_count > 0 && !grid) || (theta_draws != 1 && theta_draws != draws)) {
    set_error("ttb2_coalescent_piecewise: null argument, draws < 1, fewer than 2 tips, or a "
              "theta_draws that is neither 1 nor draws");
    return TTB2_E_INVALID;
  }
  const int M = grid_count > 0 ? grid_count + 1 : tip_count - 1;
  if (theta_count != M) {
    set_error("ttb2_coalescent_piecewise: theta_count must be tip_count - 1 (skyride) or "
              "grid_count + 1 (skygrid)");
    return TTB2_E_INVALID;
  }
  const int n = 2 * tip_count - 1, ne = n + grid_count;
  int np2 = 1;
  while (np2 < ne) np2 <<= 1;
  const bool big = (size_t)np2 * 20 > 200 * 1024;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    set_error("ttb2_coalescent_piecewise: no CUDA device available (sm_100a required; there is no CPU fallback)");
    return TTB2_E_CUDA;
  }
  if (device < 0 || device >= ndev) {
    set_error("ttb2_coalescent_piecewise: device index out of range");
    return TTB2_E_INVALID;
  }
  TTB2_CUDA_CHECK(cudaSetDevice(device));
  const size_t nh = (size_t)draws * n, nt = (size_t)theta_draws * M, ndt = (size_t)draws * M;
  // device layout: heights | theta | grid | logp | d_heights | d_theta | [sort scratch]
  const size_t scratchN = big ? (size_t)draws * np2 * 3 : 0;
  const size_t need = nh + nt + grid_count + draws + nh + ndt + scratchN;
  double* buf = nullptr;
  const double *dH = node_heights, *dT = theta, *dG = grid;
  double *dL = log_prob, *dGH = d_heights, *dGT = d_theta, *sortScratch = nullptr;
  if (where == TTB2_HOST) {
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
    buf = sc.buf;
    TTB2_CUDA_CHECK(cudaMemcpyAsync(buf, node_heights, nh * sizeof(double), cudaMemcpyHostToDevice, 0));
    TTB2_CUDA_CHECK(cudaMemcpyAsync(buf + nh, theta, nt * sizeof(double), cudaMemcpyHostToDevice, 0));
    if (grid_count)
      TTB2_CUDA_CHECK(cudaMemcpyAsync(buf + nh + nt, grid, grid_count * sizeof(double),
                                      cudaMemcpyHostToDevice, 0));
    dH = buf;
    dT = buf + nh;
    dG = buf + nh + nt;
    dL = buf + nh + nt + grid_count;
    dGH = d_heights ? dL + draws : nullptr;
    dGT = d_theta ? dL + draws + nh : nullptr;
    if (big) sortScratch = dL + draws + nh + ndt;
  } else if (big) {
    TTB2_CUDA_CHECK(cudaMalloc((void**)&sortScratch, scratchN * sizeof(double)));
  }
  const size_t smem = big ? 0 : (size_t)np2 * 20;
  if (smem > 48 * 1024)
    TTB2_CUDA_CHECK(cudaFuncSetAttribute(coalescent_piecewise_kernel,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  coalescent_piecewise_kernel<<<draws, CO_THREADS, smem, 0>>>(dH, dG, dT, theta_draws, dL, dGH, dGT,
                                                             tip_count, grid_count, M, np2,
                                                             sortScratch);
  {
    const cudaError_t lerr = cudaGetLastError();
    if (lerr != cudaSuccess) {
      if (big && where != TTB2_HOST) cudaFree(sortScratch);
      set_error(std::string("coalescent_piecewise_kernel: ") + cudaGetErrorString(lerr));
      return TTB2_E_CUDA;
    }
  }
  if (where == TTB2_HOST) {
    TTB2_CUDA_CHECK(cudaMemcpyAsync(log_prob, dL, draws * sizeof(double), cudaMemcpyDeviceToHost, 0));
    if (d_heights)
      TTB2_CUDA_CHECK(cudaMemcpyAsync(d_heights, dGH, nh * sizeof(double), cudaMemcpyDeviceToHost, 0));
    if (d_theta)
      TTB2_CUDA_CHECK(cudaMemcpyAsync(d_theta, dGT, ndt * sizeof(double), cudaMemcpyDeviceToHost, 0));
  }
  TTB2_CUDA_CHECK(cudaStreamSynchronize(0));
  if (big && where != TTB2_HOST) cudaFree(sortScratch);
  return TTB2_OK;
}
