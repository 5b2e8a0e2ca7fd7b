// Batched matrix exponential on the device for general (non-reversible) generators:
// P[d][b][k] = exp(Q_d * t_{d,b} * r_{d,k}) for every branch x category x draw in one launch,
// and the exact adjoint for the gradient -- SURVEY 8(f) row f4, second half.
//
// Replaces: NonSymmetricSubstitutionModel.p_t = torch.matrix_exp(Q t)
// (torchtree/evolution/substitution_model/abstract.py:89-94) and its autograd backward, the
// route of GeneralNonSymmetricSubstitutionModel (general.py:203-291), i.e. the discrete-trait /
// phylogeography likelihoods the CLI builds at cli/evolution.py:540-611.  A non-reversible
// generator has a complex spectrum, so there is no real eigen route; scaling-and-squaring works
// for any matrix.
//
// Forward, one CTA per (draw, branch, category), everything in shared memory:
//   A = Q tau;  s = smallest integer with ||A||_inf / 2^s <= 1/2;  X = A / 2^s
//   E = sum_{j<=18} X^j / j!   (Horner; remainder < 2^-19 / 19! ~ 1.6e-23)
//   E <- E^2, s times.
// Backward: d<G, exp(A)>/dA = L(A^T, G), the Frechet derivative of exp at A^T applied to
// G = d lnL / d P, obtained WITHOUT storing any intermediate of the forward pass by running
// the same recurrence in "dual" arithmetic on the block matrix [[A^T, G], [0, A^T]]:
//   (E, dE) <- (I + X E / j, (X dE + dX E) / j)      X = A^T / 2^s, dX = G / 2^s
//   (E, dE) <- (E E, E dE + dE E)                     s times;   L = dE.
// Then d lnL / d tau = <L, Q> and d lnL / d Q += tau L, reduced over branches and categories
// by the kernels the eigen route uses (kernels_small.cu).
#include <cmath>

#include "engine.cuh"

namespace ttb2 {

namespace {

constexpr int EX_THREADS = 256;
constexpr int EX_TAYLOR = 18;

// C = A . B  (row-major S x S in shared memory; C must not alias A or B)
__device__ __forceinline__ void ex_matmul(double* C, const double* A, const double* B, int S) {
  for (int idx = threadIdx.x; idx < S * S; idx += blockDim.x) {
    const int i = idx / S, j = idx - i * S;
    double acc = 0.0;
    for (int l = 0; l < S; ++l) acc = fma(A[i * S + l], B[l * S + j], acc);
    C[idx] = acc;
  }
}

// C += A . B
__device__ __forceinline__ void ex_matmul_acc(double* C, const double* A, const double* B, int S) {
  for (int idx = threadIdx.x; idx < S * S; idx += blockDim.x) {
    const int i = idx / S, j = idx - i * S;
    double acc = C[idx];
    for (int l = 0; l < S; ++l) acc = fma(A[i * S + l], B[l * S + j], acc);
    C[idx] = acc;
  }
}

// number of halvings that bring the infinity norm of A (S x S in shared memory) to <= 1/2
__device__ int ex_scaling(const double* A, int S, double* red) {
  double rowmax = 0.0;
  bool bad = false;
  for (int i = threadIdx.x; i < S; i += blockDim.x) {
    double sum = 0.0;
    for (int j = 0; j < S; ++j) sum += fabs(A[i * S + j]);
    if (!(sum == sum) || isinf(sum)) bad = true;
    rowmax = fmax(rowmax, sum);
  }
  __syncthreads();
  if (threadIdx.x < 32) red[threadIdx.x] = 0.0;
  __syncthreads();
  // S <= 64 rows: the first 64 threads hold the row sums
  if (threadIdx.x < 64) {
    // max is order-independent: a shared-memory atomic on the bit pattern of a non-negative
    // double is exact and deterministic
    atomicMax(reinterpret_cast<unsigned long long*>(red),
              (unsigned long long)__double_as_longlong(bad ? INFINITY : rowmax));
  }
  __syncthreads();
  const double norm = red[0];
  __syncthreads();
  if (!(norm > 0.5) || isinf(norm)) return 0;   // small, zero, or not finite (NaN/inf propagate)
  int s = 0;
  double x = norm;
  while (x > 0.5 && s < 1000) {
    x *= 0.5;
    ++s;
  }
  return s;
}

// forward: mats[d][b][k] = exp(Q[dq] * bl[d][b] * rates[dr][k])
__global__ void __launch_bounds__(EX_THREADS)
expm_forward_kernel(const double* __restrict__ q, int qDraws, const double* __restrict__ bl,
                    const double* __restrict__ rates, int rateDraws, double* __restrict__ mats,
                    int B, int K, int S) {
  extern __shared__ double sm[];
  __shared__ double red[32];
  const int SS = S * S;
  double* X = sm;
  double* E = X + SS;
  double* T = E + SS;
  const int bk = blockIdx.x, b = bk / K, k = bk - b * K, d = blockIdx.y;
  const double tau = bl[(size_t)d * B + b] * rates[(size_t)(rateDraws > 1 ? d : 0) * K + k];
  const double* Q = q + (size_t)(qDraws > 1 ? d : 0) * SS;
  for (int idx = threadIdx.x; idx < SS; idx += blockDim.x) X[idx] = Q[idx] * tau;
  __syncthreads();
  const int s = ex_scaling(X, S, red);
  const double scale = scalbn(1.0, -s);
  for (int idx = threadIdx.x; idx < SS; idx += blockDim.x) {
    X[idx] *= scale;
    E[idx] = (idx / S == idx % S) ? 1.0 : 0.0;
  }
  __syncthreads();
  for (int j = EX_TAYLOR; j >= 1; --j) {
    ex_matmul(T, X, E, S);
    __syncthreads();
    const double inv = 1.0 / j;
    for (int idx = threadIdx.x; idx < SS; idx += blockDim.x)
      E[idx] = ((idx / S == idx % S) ? 1.0 : 0.0) + T[idx] * inv;
    __syncthreads();
  }
  for (int i = 0; i < s; ++i) {
    ex_matmul(T, E, E, S);
    __syncthreads();
    for (int idx = threadIdx.x; idx < SS; idx += blockDim.x) E[idx] = T[idx];
    __syncthreads();
  }
  double* out = mats + (((size_t)d * B + b) * K + k) * SS;
  for (int idx = threadIdx.x; idx < SS; idx += blockDim.x) out[idx] = E[idx];
}

// backward: L = Frechet adjoint of exp at A = Q tau applied to G = dmat[d][b][k];
//   gscal[item] = <L, Q>  (d lnL / d tau),  hpart[item] = tau L  (this item's share of d lnL / d Q)
__global__ void __launch_bounds__(EX_THREADS)
expm_backward_kernel(const double* __restrict__ q, int qDraws, const double* __restrict__ bl,
                     const double* __restrict__ rates, int rateDraws,
                     const double* __restrict__ dmat, double* __restrict__ hpart,
                     double* __restrict__ gscal, int B, int K, int S) {
  extern __shared__ double sm[];
  __shared__ double red[32];
  const int SS = S * S;
  double* X = sm;          // A^T / 2^s
  double* dX = X + SS;     // G / 2^s
  double* E = dX + SS;
  double* dE = E + SS;
  double* T1 = dE + SS;
  double* T2 = T1 + SS;
  const int bk = blockIdx.x, b = bk / K, k = bk - b * K, d = blockIdx.y;
  const size_t item = ((size_t)d * B + b) * K + k;
  const double tau = bl[(size_t)d * B + b] * rates[(size_t)(rateDraws > 1 ? d : 0) * K + k];
  const double* Q = q + (size_t)(qDraws > 1 ? d : 0) * SS;
  const double* G = dmat + item * SS;
  for (int idx = threadIdx.x; idx < SS; idx += blockDim.x) {
    const int i = idx / S, j = idx - i * S;
    X[idx] = Q[j * S + i] * tau;     // transposed
    dX[idx] = G[idx];
  }
  __syncthreads();
  const int s = ex_scaling(X, S, red);   // ||A^T||_inf = ||A||_1: any consistent norm serves
  const double scale = scalbn(1.0, -s);
  for (int idx = threadIdx.x; idx < SS; idx += blockDim.x) {
    X[idx] *= scale;
    dX[idx] *= scale;
    E[idx] = (idx / S == idx % S) ? 1.0 : 0.0;
    dE[idx] = 0.0;
  }
  __syncthreads();
  for (int j = EX_TAYLOR; j >= 1; --j) {
    ex_matmul(T2, X, dE, S);         // X dE
    __syncthreads();
    ex_matmul_acc(T2, dX, E, S);     // + dX E
    ex_matmul(T1, X, E, S);          // X E   (E is only read in this phase)
    __syncthreads();
    const double inv = 1.0 / j;
    for (int idx = threadIdx.x; idx < SS; idx += blockDim.x) {
      E[idx] = ((idx / S == idx % S) ? 1.0 : 0.0) + T1[idx] * inv;
      dE[idx] = T2[idx] * inv;
    }
    __syncthreads();
  }
  for (int i = 0; i < s; ++i) {
    ex_matmul(T2, E, dE, S);
    __syncthreads();
    ex_matmul_acc(T2, dE, E, S);
    ex_matmul(T1, E, E, S);
    __syncthreads();
    for (int idx = threadIdx.x; idx < SS; idx += blockDim.x) {
      E[idx] = T1[idx];
      dE[idx] = T2[idx];
    }
    __syncthreads();
  }
  // dE = L(A^T, G): gradient with respect to A (same index order as Q)
  double part = 0.0;
  double* H = hpart + item * SS;
  for (int idx = threadIdx.x; idx < SS; idx += blockDim.x) {
    const double l = dE[idx];
    part = fma(l, Q[idx], part);
    H[idx] = tau * l;
  }
  // fixed-order block sum
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = part;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[w];
    gscal[item] = t;
  }
}

}  // namespace

size_t expm_smem_bytes(int S, bool backward) {
  return (size_t)(backward ? 6 : 3) * S * S * sizeof(double);
}

int small_expm_forward(Engine& e, int draws) {
  const Dims& m = e.dm;
  const size_t smem = expm_smem_bytes(m.S, false);
  if (smem > 48 * 1024)
    TTB2_CUDA_CHECK(cudaFuncSetAttribute(expm_forward_kernel,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid(m.B * m.K, draws);
  expm_forward_kernel<<<grid, EX_THREADS, smem, e.stream>>>(e.qnorm, e.qDraws, e.bl, e.rates,
                                                            e.rateDraws, e.mats, m.B, m.K, m.S);
  ++e.launches;
  TTB2_CUDA_CHECK(cudaGetLastError());
  return TTB2_OK;
}

int small_expm_backward(Engine& e, int draws) {
  const Dims& m = e.dm;
  const size_t smem = expm_smem_bytes(m.S, true);
  if (smem > 48 * 1024)
    TTB2_CUDA_CHECK(cudaFuncSetAttribute(expm_backward_kernel,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid(m.B * m.K, draws);
  expm_backward_kernel<<<grid, EX_THREADS, smem, e.stream>>>(
      e.qnorm, e.qDraws, e.bl, e.rates, e.rateDraws, e.dmat, e.hpart, e.gscal, m.B, m.K, m.S);
  ++e.launches;
  TTB2_CUDA_CHECK(cudaGetLastError());
  return TTB2_OK;
}

}  // namespace ttb2
