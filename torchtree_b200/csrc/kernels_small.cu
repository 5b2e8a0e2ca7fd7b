// Small per-evaluation kernels: transition matrices for every branch x rate
// category x draw in one launch, deterministic second-stage reductions, and the
// contraction of d lnL / d P into branch-length, site-rate and generator
// gradients.  All are O(B K S^3) or O(N) and negligible next to the peeling.
#include <algorithm>

#include <cstdlib>
#include "engine.cuh"

namespace ttb2 {

namespace {

constexpr int RED_THREADS = 256;

__device__ __forceinline__ double block_sum256(double v, double* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  double t = 0.0;
  if (threadIdx.x == 0) {
    const int nw = (blockDim.x + 31) >> 5;
    for (int w = 0; w < nw; ++w) t += red[w];
  }
  return t;
}

// P[d][b][k] = V diag(exp(lambda * t_b * r_k)) V^-1
// (SymmetricSubstitutionModel.p_t, substitution_model/abstract.py:57-76, with
// the eigen-system supplied by the host; tree_likelihood.py:346 for t*r)
__global__ void pmatrix_kernel(const double* __restrict__ bl,
                               const double* __restrict__ rates, int rateDraws,
                               const double* __restrict__ evec,
                               const double* __restrict__ ivec,
                               const double* __restrict__ eval, int eigDraws,
                               double* __restrict__ mats, int B, int K, int S) {
  extern __shared__ double ex[];
  const int bk = blockIdx.x;
  const int b = bk / K, k = bk - b * K;
  const int d = blockIdx.y;
  const int de = eigDraws > 1 ? d : 0;
  const double t = bl[(size_t)d * B + b] * rates[(size_t)(rateDraws > 1 ? d : 0) * K + k];
  const double* lam = eval + (size_t)de * S;
  for (int m = threadIdx.x; m < S; m += blockDim.x) ex[m] = exp(lam[m] * t);
  __syncthreads();
  const double* V = evec + (size_t)de * S * S;
  const double* Vi = ivec + (size_t)de * S * S;
  double* P = mats + (((size_t)d * B + b) * K + k) * S * S;
  for (int idx = threadIdx.x; idx < S * S; idx += blockDim.x) {
    const int i = idx / S, j = idx - i * S;
    double acc = 0.0;
    for (int m = 0; m < S; ++m) acc = fma(V[i * S + m] * ex[m], Vi[m * S + j], acc);
    P[idx] = acc;
  }
}

// Four states: a half-warp per (draw, branch, category) matrix, lane = element (i, j), 16
// matrices per CTA (a batch of draws otherwise launches one 32-thread CTA per matrix: 127 744 at
// config 3).  Same FMA order as pmatrix_kernel.
__global__ void __launch_bounds__(256)
pmatrix4_kernel(const double* __restrict__ bl, const double* __restrict__ rates, int rateDraws,
                const double* __restrict__ evec, const double* __restrict__ ivec,
                const double* __restrict__ eval, int eigDraws, double* __restrict__ mats, int B,
                int K, int draws) {
  const int e = threadIdx.x & 15;
  const long items = (long)draws * B * K;
  const long item = (long)blockIdx.x * 16 + (threadIdx.x >> 4);
  if (item >= items) return;
  const int d = (int)(item / ((long)B * K));
  const int bk = (int)(item - (long)d * B * K);
  const int b = bk / K, k = bk - b * K;
  const int de = eigDraws > 1 ? d : 0;
  const double t = bl[(size_t)d * B + b] * rates[(size_t)(rateDraws > 1 ? d : 0) * K + k];
  const double* lam = eval + (size_t)de * 4;
  const double* V = evec + (size_t)de * 16;
  const double* Vi = ivec + (size_t)de * 16;
  const int i = e >> 2, j = e & 3;
  double acc = 0.0;
#pragma unroll
  for (int m = 0; m < 4; ++m) acc = fma(V[i * 4 + m] * exp(lam[m] * t), Vi[m * 4 + j], acc);
  mats[(size_t)item * 16 + e] = acc;
}

// out[d] = sum_j part[d][j]   (fixed order: deterministic)
__global__ void __launch_bounds__(RED_THREADS)
reduce_rows_kernel(const double* __restrict__ part, double* __restrict__ out, int n,
                   int width, int col) {
  // part is [rows][n][width]; reduces column `col + blockIdx.x` of row blockIdx.y
  __shared__ double red[RED_THREADS / 32];
  const int c = col + blockIdx.x;
  const double* p = part + (size_t)blockIdx.y * n * width + c;
  double acc = 0.0;
  for (int j = threadIdx.x; j < n; j += blockDim.x) acc += p[(size_t)j * width];
  const double t = block_sum256(acc, red);
  if (threadIdx.x == 0) out[(size_t)blockIdx.y * gridDim.x + blockIdx.x] = t;
}

// dmat[d][b][k][e] = sum_chunk gpart[d][chunkBase[b] + k * chunkCount[b] + chunk][e]
// Grid (items = d x b x k, element blocks of GR_THREADS): the few branches near the root have
// hundreds of chunks, so the elements of one item are spread over blocks (a 61-state item has
// 3721) and every thread keeps four loads in flight.  SS <= GR_THREADS: GR_THREADS / SS row
// groups stride over the chunks, then a fixed-order sum over the groups.  The order of the
// additions is fixed by (SS, chunk count) alone: results are reproducible bit for bit.
constexpr int GR_THREADS = 256;

__global__ void __launch_bounds__(GR_THREADS)
gpart_reduce_kernel(const double* __restrict__ gpart, const int* __restrict__ chunkBase,
                    const int* __restrict__ chunkCount, double* __restrict__ dmat,
                    size_t chunkTotal, int B, int K, int SS) {
  extern __shared__ double part[];   // [groups][SS]
  const size_t item = blockIdx.x;
  const int k = (int)(item % K);
  const int b = (int)((item / K) % B);
  const size_t d = item / ((size_t)K * B);
  const int n = chunkCount[b];
  const double* p = gpart + (d * chunkTotal + chunkBase[b] + (size_t)k * n) * SS;
  if (SS >= GR_THREADS) {
    const int e = blockIdx.y * GR_THREADS + threadIdx.x;
    if (e >= SS) return;
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
    int c = 0;
    for (; c + 4 <= n; c += 4) {
      a0 += p[(size_t)c * SS + e];
      a1 += p[(size_t)(c + 1) * SS + e];
      a2 += p[(size_t)(c + 2) * SS + e];
      a3 += p[(size_t)(c + 3) * SS + e];
    }
    for (; c < n; ++c) a0 += p[(size_t)c * SS + e];
    dmat[item * SS + e] = (a0 + a1) + (a2 + a3);
    return;
  }
  const int groups = GR_THREADS / SS;
  const int e = (int)(threadIdx.x % SS);
  const int r = (int)(threadIdx.x / SS);
  double acc = 0.0;
  if (r < groups)
    for (int c = r; c < n; c += groups) acc += p[(size_t)c * SS + e];
  if (r < groups) part[r * SS + e] = acc;
  __syncthreads();
  if (r == 0) {
    double t = 0.0;
    for (int g = 0; g < groups; ++g) t += part[g * SS + e];
    dmat[item * SS + e] = t;
  }
}

// out[d][j] = g[d] * in[d][j]
__global__ void scale_rows_kernel(const double* __restrict__ in, const double* __restrict__ g,
                                  double* __restrict__ out, size_t width, int draws) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= width * draws) return;
  out[idx] = in[idx] * g[idx / width];
}

// d_props (blockIdx.y == 0) and the root term of d_freqs (blockIdx.y == 1) from rootGrad [D][K+S]
__global__ void root_outputs_kernel(const double* __restrict__ in, const double* __restrict__ g,
                                    double* __restrict__ outProps, double* __restrict__ outFreqs,
                                    int K, int S, int inWidth, int draws, int propOut, int freqOut) {
  const bool second = blockIdx.y == 1;
  const int width = second ? S : K, off = second ? K : 0, outDraws = second ? freqOut : propOut;
  double* out = second ? outFreqs : outProps;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (outDraws > 1) {
    if (idx >= width * draws) return;
    const int d = idx / width, j = idx - d * width;
    out[idx] = g[d] * in[(size_t)d * inWidth + off + j];
  } else {
    // one warp per output, lanes over the draws (a parameter shared by a batch of draws)
    const int o = idx >> 5, lane = idx & 31;
    if (o >= width) return;   // whole warps: blockDim is a multiple of 32
    double acc = 0.0;
    for (int d = lane; d < draws; d += 32) acc = fma(g[d], in[(size_t)d * inWidth + off + o], acc);
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, s);
    if (lane == 0) out[o] = acc;
  }
}

// Per branch x category: M = V^T G V^-T (eigenbasis),
//   gscal = sum_i M_ii lambda_i exp(lambda_i tau)        (= tr(G^T Q P), d/d tau)
//   hpart = M o Phi(tau), Phi_ij = (e^{l_i tau} - e^{l_j tau}) / (l_i - l_j),
//           Phi_ii = tau e^{l_i tau}   (divided differences, expm1 form: finite
//           and accurate at repeated eigenvalues -- SURVEY F12, Appendix B)
// With `gpart` != nullptr the per-chunk partial sums of G are reduced here (fixed order:
// groups of threads stride over the chunks, then a serial sum over the groups) and the result
// is also stored to `dmat` -- one launch and one pass over G less than gpart_reduce_kernel +
// this kernel (the small-kernel tail is 10 % of an 8-GPU shard's step).
__global__ void eigen_contract_kernel(double* __restrict__ dmat,
                                      const double* __restrict__ gpart,
                                      const int* __restrict__ chunkBase,
                                      const int* __restrict__ chunkCount, size_t chunkTotal,
                                      const double* __restrict__ bl,
                                      const double* __restrict__ rates, int rateDraws,
                                      const double* __restrict__ evec,
                                      const double* __restrict__ ivec,
                                      const double* __restrict__ eval, int eigDraws,
                                      double* __restrict__ hpart, double* __restrict__ gscal,
                                      int B, int K, int S) {
  extern __shared__ double sm[];
  double* sG = sm;              // [S][S]
  double* sT = sG + S * S;      // [S][S] tmp = V^T G
  double* sV = sT + S * S;      // [S][S]
  double* sVi = sV + S * S;     // [S][S]
  double* sLam = sVi + S * S;   // [S]
  double* sEx = sLam + S;       // [S]
  __shared__ double red[32];
  const int bk = blockIdx.x;
  const int b = bk / K, k = bk - b * K;
  const int d = blockIdx.y;
  const int de = eigDraws > 1 ? d : 0;
  const double tau = bl[(size_t)d * B + b] * rates[(size_t)(rateDraws > 1 ? d : 0) * K + k];
  const size_t item = ((size_t)d * B + b) * K + k;
  double* G = dmat + item * S * S;
  if (gpart != nullptr) {
    const int SS = S * S;
    const int n = chunkCount[b];
    const double* p = gpart + ((size_t)d * chunkTotal + chunkBase[b] + (size_t)k * n) * SS;
    double* part = sEx + S;   // scratch [groups][per] = blockDim.x doubles behind the tables
    const int per = SS < (int)blockDim.x ? SS : (int)blockDim.x;
    const int groups = SS < (int)blockDim.x ? (int)blockDim.x / SS : 1;
    for (int e0 = 0; e0 < SS; e0 += per) {
      const int e = e0 + (int)(threadIdx.x % per);
      const int r = (int)(threadIdx.x / per);
      double acc = 0.0;
      if (r < groups && e < SS)
        for (int c = r; c < n; c += groups) acc += p[(size_t)c * SS + e];
      if (r < groups && e < SS) part[r * per + (e - e0)] = acc;
      __syncthreads();
      if (r == 0 && e < SS) {
        double t = 0.0;
        for (int g = 0; g < groups; ++g) t += part[g * per + (e - e0)];
        sG[e] = t;
        G[e] = t;
      }
      __syncthreads();
    }
  }
  for (int idx = threadIdx.x; idx < S * S; idx += blockDim.x) {
    if (gpart == nullptr) sG[idx] = G[idx];
    sV[idx] = evec[(size_t)de * S * S + idx];
    sVi[idx] = ivec[(size_t)de * S * S + idx];
  }
  for (int m = threadIdx.x; m < S; m += blockDim.x) {
    const double l = eval[(size_t)de * S + m];
    sLam[m] = l;
    sEx[m] = exp(l * tau);
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < S * S; idx += blockDim.x) {
    const int i = idx / S, c = idx - i * S;
    double acc = 0.0;
    for (int a = 0; a < S; ++a) acc = fma(sV[a * S + i], sG[a * S + c], acc);
    sT[idx] = acc;
  }
  __syncthreads();
  double diag = 0.0;
  double* H = hpart + item * S * S;
  for (int idx = threadIdx.x; idx < S * S; idx += blockDim.x) {
    const int i = idx / S, j = idx - i * S;
    double m = 0.0;
    for (int c = 0; c < S; ++c) m = fma(sT[i * S + c], sVi[j * S + c], m);
    double phi;
    if (i == j) {
      phi = tau * sEx[i];
      diag = fma(m, sLam[i] * sEx[i], diag);
    } else {
      const double a = sLam[i] * tau, bb = sLam[j] * tau;
      const double hi = a > bb ? a : bb;
      const double x = -fabs(a - bb);  // <= 0
      const double ratio = (x > -1e-9) ? 1.0 + 0.5 * x : expm1(x) / x;
      phi = tau * exp(hi) * ratio;
    }
    H[idx] = m * phi;
  }
  const double t = block_sum256(diag, red);
  if (threadIdx.x == 0) gscal[item] = t;
}

// ---------------------------------------------------------------------------
// Four states: the contraction above spends a 256-thread CTA on a 4 x 4 problem -- with a batch of
// draws (config 3: 998 branches x 128 draws = 127 744 CTAs) 0.66 ms of launch overhead per
// evaluation.  Here a half-warp owns one (draw, branch, category) item, lane = element (i, j);
// the two 4 x 4 products go through half-warp shuffles, 16 items per CTA.  Same FMA order as
// eigen_contract_kernel; the sum of the per-chunk partials of G is taken serially (fixed order).
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
eigen_contract4_kernel(double* __restrict__ dmat, const double* __restrict__ gpart,
                       const int* __restrict__ chunkBase, const int* __restrict__ chunkCount,
                       size_t chunkTotal, const double* __restrict__ bl,
                       const double* __restrict__ rates, int rateDraws,
                       const double* __restrict__ evec, const double* __restrict__ ivec,
                       const double* __restrict__ eval, int eigDraws, double* __restrict__ hpart,
                       double* __restrict__ gscal, int B, int K, int draws) {
  const int e = threadIdx.x & 15;
  const long items = (long)draws * B * K;
  const long item = (long)blockIdx.x * 16 + (threadIdx.x >> 4);
  const bool live = item < items;
  const long it = live ? item : items - 1;   // dead half-warps shadow the last item (no stores)
  const int d = (int)(it / ((long)B * K));
  const int bk = (int)(it - (long)d * B * K);
  const int b = bk / K, k = bk - b * K;
  const int de = eigDraws > 1 ? d : 0;
  const double tau = bl[(size_t)d * B + b] * rates[(size_t)(rateDraws > 1 ? d : 0) * K + k];
  const int i = e >> 2, j = e & 3;
  double G;
  if (gpart != nullptr) {
    const int n = chunkCount[b];
    const double* p = gpart + ((size_t)d * chunkTotal + chunkBase[b] + (size_t)k * n) * 16 + e;
    G = 0.0;
    for (int c = 0; c < n; ++c) G += p[(size_t)c * 16];
    if (live) dmat[(size_t)it * 16 + e] = G;
  } else {
    G = dmat[(size_t)it * 16 + e];
  }
  const double* V = evec + (size_t)de * 16;
  const double* Vi = ivec + (size_t)de * 16;
  const double* lam = eval + (size_t)de * 4;
  double T = 0.0;   // T[i][j] = sum_a V[a][i] G[a][j]
#pragma unroll
  for (int a = 0; a < 4; ++a) T = fma(V[a * 4 + i], __shfl_sync(0xffffffffu, G, a * 4 + j, 16), T);
  double M = 0.0;   // M[i][j] = sum_c T[i][c] Vi[j][c]
#pragma unroll
  for (int c = 0; c < 4; ++c) M = fma(__shfl_sync(0xffffffffu, T, i * 4 + c, 16), Vi[j * 4 + c], M);
  const double li = lam[i], lj = lam[j];
  const double exi = exp(li * tau);
  double phi, diag = 0.0;
  if (i == j) {
    phi = tau * exi;
    diag = M * (li * exi);
  } else {
    const double a = li * tau, bb = lj * tau;
    const double hi = a > bb ? a : bb;
    const double x = -fabs(a - bb);  // <= 0
    const double ratio = (x > -1e-9) ? 1.0 + 0.5 * x : expm1(x) / x;
    phi = tau * exp(hi) * ratio;
  }
  if (live) hpart[(size_t)it * 16 + e] = M * phi;
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) diag += __shfl_xor_sync(0xffffffffu, diag, o, 16);
  if (live && e == 0) gscal[it] = diag;
}

// ---------------------------------------------------------------------------
// Large alphabets (S > 32: the 61-state codon path).  The two kernels above read both operands
// of every FMA from memory (global for pmatrix, shared for the contraction: two LDS per FMA, so
// the shared-memory pipe bounded them: 0.16 and 0.29 ms per evaluation at config 5).  Here the
// operands are staged once in shared memory as [S][LD] (LD = S rounded up to 4, zero padding
// columns) and a thread owns FOUR adjacent columns of a row: per reduction step one broadcast
// load of A and two 128-bit loads of B feed four FMAs.
// ---------------------------------------------------------------------------
// c[0..3] += sum_m A(i, m) B[m][j0 .. j0 + 3];  A(i, m) = TRANS_A ? A[m][i] : A[i][m]
template <bool TRANS_A>
__device__ __forceinline__ void mm_row4(double (&c)[4], const double* A, const double* B, int i,
                                        int j0, int S, int LD) {
  for (int m = 0; m < S; ++m) {
    const double a = TRANS_A ? A[m * LD + i] : A[i * LD + m];
    const double2 b0 = *reinterpret_cast<const double2*>(B + m * LD + j0);
    const double2 b1 = *reinterpret_cast<const double2*>(B + m * LD + j0 + 2);
    c[0] = fma(a, b0.x, c[0]);
    c[1] = fma(a, b0.y, c[1]);
    c[2] = fma(a, b1.x, c[2]);
    c[3] = fma(a, b1.y, c[3]);
  }
}

// shared: A = V diag(exp(lambda tau)) [S][LD] | B = V^-1 [S][LD] | ex [S]
__global__ void __launch_bounds__(512)
pmatrix_tiled_kernel(const double* __restrict__ bl, const double* __restrict__ rates, int rateDraws,
                     const double* __restrict__ evec, const double* __restrict__ ivec,
                     const double* __restrict__ eval, int eigDraws, double* __restrict__ mats,
                     int B, int K, int S) {
  extern __shared__ __align__(16) double sm[];
  const int LD = (S + 3) & ~3;
  double* sA = sm;
  double* sB = sA + S * LD;
  double* ex = sB + S * LD;
  const int bk = blockIdx.x;
  const int b = bk / K, k = bk - b * K;
  const int d = blockIdx.y;
  const int de = eigDraws > 1 ? d : 0;
  const double t = bl[(size_t)d * B + b] * rates[(size_t)(rateDraws > 1 ? d : 0) * K + k];
  const double* lam = eval + (size_t)de * S;
  for (int m = threadIdx.x; m < S; m += blockDim.x) ex[m] = exp(lam[m] * t);
  __syncthreads();
  const double* V = evec + (size_t)de * S * S;
  const double* Vi = ivec + (size_t)de * S * S;
  for (int idx = threadIdx.x; idx < S * LD; idx += blockDim.x) {
    const int i = idx / LD, j = idx - i * LD;
    sA[idx] = j < S ? V[i * S + j] * ex[j] : 0.0;
    sB[idx] = j < S ? Vi[i * S + j] : 0.0;
  }
  __syncthreads();
  double* P = mats + (((size_t)d * B + b) * K + k) * S * S;
  const int Q4 = LD / 4;
  for (int item = threadIdx.x; item < S * Q4; item += blockDim.x) {
    const int i = item / Q4, j0 = (item - i * Q4) * 4;
    double c[4] = {0.0, 0.0, 0.0, 0.0};
    mm_row4<false>(c, sA, sB, i, j0, S, LD);
#pragma unroll
    for (int q = 0; q < 4; ++q)
      if (j0 + q < S) P[i * S + j0 + q] = c[q];
  }
}

// shared: G [S][LD] | T = V^T G [S][LD] | V [S][LD] | ViT = (V^-1)^T [S][LD] | lambda [S] | ex [S]
__global__ void __launch_bounds__(1024)
eigen_contract_tiled_kernel(const double* __restrict__ dmat, const double* __restrict__ bl,
                            const double* __restrict__ rates, int rateDraws,
                            const double* __restrict__ evec, const double* __restrict__ ivec,
                            const double* __restrict__ eval, int eigDraws,
                            double* __restrict__ hpart, double* __restrict__ gscal, int B, int K,
                            int S) {
  extern __shared__ __align__(16) double sm[];
  __shared__ double red[32];
  const int LD = (S + 3) & ~3;
  double* sG = sm;
  double* sT = sG + S * LD;
  double* sV = sT + S * LD;
  double* sViT = sV + S * LD;
  double* sLam = sViT + S * LD;
  double* sEx = sLam + S;
  const int bk = blockIdx.x;
  const int b = bk / K, k = bk - b * K;
  const int d = blockIdx.y;
  const int de = eigDraws > 1 ? d : 0;
  const double tau = bl[(size_t)d * B + b] * rates[(size_t)(rateDraws > 1 ? d : 0) * K + k];
  const size_t item0 = ((size_t)d * B + b) * K + k;
  const double* G = dmat + item0 * S * S;
  const double* V = evec + (size_t)de * S * S;
  const double* Vi = ivec + (size_t)de * S * S;
  for (int idx = threadIdx.x; idx < S * LD; idx += blockDim.x) {
    const int i = idx / LD, j = idx - i * LD;
    sG[idx] = j < S ? G[i * S + j] : 0.0;
    sV[idx] = j < S ? V[i * S + j] : 0.0;
    sViT[idx] = j < S ? Vi[j * S + i] : 0.0;   // transposed: row c, column j
  }
  for (int m = threadIdx.x; m < S; m += blockDim.x) {
    const double l = eval[(size_t)de * S + m];
    sLam[m] = l;
    sEx[m] = exp(l * tau);
  }
  __syncthreads();
  const int Q4 = LD / 4;
  // T[i][c] = sum_a V[a][i] G[a][c]
  for (int item = threadIdx.x; item < S * Q4; item += blockDim.x) {
    const int i = item / Q4, j0 = (item - i * Q4) * 4;
    double c[4] = {0.0, 0.0, 0.0, 0.0};
    mm_row4<true>(c, sV, sG, i, j0, S, LD);
#pragma unroll
    for (int q = 0; q < 4; ++q) sT[i * LD + j0 + q] = c[q];
  }
  __syncthreads();
  // M[i][j] = sum_c T[i][c] Vi[j][c];  H = M o Phi(tau)
  double diag = 0.0;
  double* H = hpart + item0 * S * S;
  for (int item = threadIdx.x; item < S * Q4; item += blockDim.x) {
    const int i = item / Q4, j0 = (item - i * Q4) * 4;
    double c[4] = {0.0, 0.0, 0.0, 0.0};
    mm_row4<false>(c, sT, sViT, i, j0, S, LD);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int j = j0 + q;
      if (j >= S) continue;
      double phi;
      if (i == j) {
        phi = tau * sEx[i];
        diag = fma(c[q], sLam[i] * sEx[i], diag);
      } else {
        const double a = sLam[i] * tau, bb = sLam[j] * tau;
        const double hi = a > bb ? a : bb;
        const double x = -fabs(a - bb);  // <= 0
        const double ratio = (x > -1e-9) ? 1.0 + 0.5 * x : expm1(x) / x;
        phi = tau * exp(hi) * ratio;
      }
      H[i * S + j] = c[q] * phi;
    }
  }
  const double t = block_sum256(diag, red);
  if (threadIdx.x == 0) gscal[item0] = t;
}

// d_bl[d][b] = g[d] * sum_k r_k gscal[d][b][k]
__global__ void branch_grad_kernel(const double* __restrict__ gscal,
                                   const double* __restrict__ rates, int rateDraws,
                                   const double* __restrict__ g, double* __restrict__ out,
                                   int B, int K, int draws) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * draws) return;
  const int d = idx / B;
  const double* r = rates + (size_t)(rateDraws > 1 ? d : 0) * K;
  const double* gs = gscal + (size_t)idx * K;
  double acc = 0.0;
  for (int k = 0; k < K; ++k) acc = fma(r[k], gs[k], acc);
  out[idx] = g[d] * acc;
}

// d_rate[od][k] = sum_d g[d] sum_b t[d][b] gscal[d][b][k]   (one block per (k, od))
__global__ void __launch_bounds__(1024)
rate_grad_kernel(const double* __restrict__ gscal, const double* __restrict__ bl,
                 const double* __restrict__ g, double* __restrict__ out, int B, int K,
                 int draws, int outDraws) {
  __shared__ double red[32];
  const int k = blockIdx.x;
  const int od = blockIdx.y;
  const int d0 = outDraws > 1 ? od : 0;
  const int d1 = outDraws > 1 ? od + 1 : draws;
  // (draw, branch) pairs flattened over the threads: a rate shared by a batch of draws used to
  // walk the draws one after the other in a single block (0.26 ms at 128 draws)
  double acc = 0.0;
  const long n = (long)(d1 - d0) * B;
  for (long idx = threadIdx.x; idx < n; idx += blockDim.x) {
    const int d = d0 + (int)(idx / B);
    const size_t at = (size_t)d0 * B + idx;   // = d * B + b
    acc = fma(g[d] * bl[at], gscal[at * K + k], acc);
  }
  const double t = block_sum256(acc, red);
  if (threadIdx.x == 0) out[(size_t)od * K + k] = t;
}

// Hs[od][slice][e] = sum_d g[d] sum_{item in slice} hpart[d][item][e]  (one block per
// (e, od, slice); q_grad_kernel adds the H_SLICES slices in fixed order)
constexpr int H_SLICES = 8;

__global__ void __launch_bounds__(RED_THREADS)
h_reduce_kernel(const double* __restrict__ hpart, const double* __restrict__ g,
                double* __restrict__ Hs, int items, int SS, int draws, int outDraws) {
  __shared__ double red[RED_THREADS / 32];
  const int e = blockIdx.x;
  const int od = blockIdx.y;
  const int slice = blockIdx.z;
  const int per = (items + H_SLICES - 1) / H_SLICES;
  const int lo = slice * per, hi = min(items, lo + per);
  const int d0 = outDraws > 1 ? od : 0;
  const int d1 = outDraws > 1 ? od + 1 : draws;
  // (draw, item) pairs flattened over the threads (a generator shared by a batch of draws used to
  // walk the draws one after the other)
  double acc = 0.0;
  const int span = hi - lo;
  const long n = span > 0 ? (long)(d1 - d0) * span : 0;
  for (long idx = threadIdx.x; idx < n; idx += blockDim.x) {
    const int d = d0 + (int)(idx / span);
    const int it = lo + (int)(idx - (long)(d - d0) * span);
    acc = fma(g[d], hpart[((size_t)d * items + it) * SS + e], acc);
  }
  const double t = block_sum256(acc, red);
  if (threadIdx.x == 0) Hs[((size_t)od * H_SLICES + slice) * SS + e] = t;
}

// dQ = V^-T H V^T  (one block per eigen-system)
// (H and dQ may alias: H is fully staged in shared memory before dQ is written)
__global__ void q_grad_kernel(const double* __restrict__ Hs, const double* __restrict__ evec,
                              const double* __restrict__ ivec, double* __restrict__ dQ, int S) {
  extern __shared__ double sm[];
  double* sH = sm;
  double* sT = sH + S * S;
  double* sV = sT + S * S;
  double* sVi = sV + S * S;
  const size_t off = (size_t)blockIdx.x * S * S;
  for (int idx = threadIdx.x; idx < S * S; idx += blockDim.x) {
    double h = 0.0;
#pragma unroll
    for (int sl = 0; sl < H_SLICES; ++sl)
      h += Hs[((size_t)blockIdx.x * H_SLICES + sl) * S * S + idx];
    sH[idx] = h;
    sV[idx] = evec[off + idx];
    sVi[idx] = ivec[off + idx];
  }
  __syncthreads();
  // T[a][j] = sum_i Vi[i][a] H[i][j]
  for (int idx = threadIdx.x; idx < S * S; idx += blockDim.x) {
    const int a = idx / S, j = idx - a * S;
    double acc = 0.0;
    for (int i = 0; i < S; ++i) acc = fma(sVi[i * S + a], sH[i * S + j], acc);
    sT[idx] = acc;
  }
  __syncthreads();
  // dQ[a][b] = sum_j T[a][j] V[b][j]
  for (int idx = threadIdx.x; idx < S * S; idx += blockDim.x) {
    const int a = idx / S, b = idx - a * S;
    double acc = 0.0;
    for (int j = 0; j < S; ++j) acc = fma(sT[a * S + j], sV[b * S + j], acc);
    dQ[off + idx] = acc;
  }
}

struct GatherArgs {
  const double* src[8];
  unsigned long long off[8], n[8];
};

// block y = segment: dst[off + i] = src[i]
__global__ void gather_inputs_kernel(GatherArgs a, double* __restrict__ dst) {
  const int seg = blockIdx.y;
  const double* __restrict__ src = a.src[seg];
  double* out = dst + a.off[seg];
  for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
       i < a.n[seg]; i += (unsigned long long)gridDim.x * blockDim.x)
    out[i] = src[i];
}

// out[od][e] = sum_slice Hs[od][slice][e]  (the expm route needs no change of basis)
__global__ void sum_slices_kernel(const double* __restrict__ Hs, double* __restrict__ out, int SS) {
  const int od = blockIdx.y;
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= SS) return;
  double h = 0.0;
#pragma unroll
  for (int sl = 0; sl < H_SLICES; ++sl) h += Hs[((size_t)od * H_SLICES + sl) * SS + e];
  out[(size_t)od * SS + e] = h;
}

int round_threads(int n, int cap) {
  int t = (n + 31) / 32 * 32;
  if (t < 32) t = 32;
  return t > cap ? cap : t;
}

}  // namespace

int plan_chunks(Engine& e, int draws, int granule, int ctasPerSm, int residentPerSm,
                int residentLevel1) {
  const Dims& m = e.dm;
  if (e.chunkPlanDraws == draws && e.chunkBase) return TTB2_OK;
  const int nLevels = (int)e.levelOff.size() - 1;
  // never less than 2 granules per chunk
  // ctasPerSm: CTAs per SM aimed at in every launch (several waves keep the tail short)
  const long target = (long)e.smCount * ctasPerSm;
  // never less than 2 granules per chunk; 8 for large alphabets, whose CTAs stage two S x S
  // matrices (about the cost of one 32-pattern tile) before their first pattern
  const int maxChunks = std::max(1, m.Npad / ((m.S > 32 ? 8 : 2) * granule));
  e.levelChunks.assign(nLevels, 1);
  e.hostChunkBase.assign(m.B + 1, 0);
  e.hostChunkCount.assign(m.B + 1, 0);
  // levels of a chain run share one chunk count: the run is one launch of (chunks x K x draws)
  // CTAs, each walking all of the run's nodes -- one wave of 5 CTAs per SM
  const ChainRuns runs = chain_runs(e);
  std::vector<char> inRun(nLevels, 0);
  for (int l = 0; l < nLevels; ++l)
    if (runs.endOfStart[l] >= 0)
      for (int j = l; j <= runs.endOfStart[l]; ++j) inRun[j] = 1;
  for (int l = 0; l < nLevels; ++l) {
    const long count = e.levelOff[l + 1] - e.levelOff[l];
    long want = (target + count * m.K * draws - 1) / (count * m.K * draws);
    if (inRun[l]) want = ((long)e.smCount * 5 + (long)m.K * draws - 1) / ((long)m.K * draws);
    else if (residentPerSm > 0)
      want = wave_aware_chunks(count * m.K * draws, want, maxChunks,
                               (long)e.smCount * (l == 0 && residentLevel1 > 0 ? residentLevel1
                                                                                : residentPerSm));
    if (want > maxChunks) want = maxChunks;
    if (want < 1) want = 1;
    e.levelChunks[l] = (int)want;
    for (int j = e.levelOff[l]; j < e.levelOff[l + 1]; ++j) {
      e.hostChunkCount[e.hostOps[j].left] = (int)want;
      e.hostChunkCount[e.hostOps[j].right] = (int)want;
    }
  }
  size_t total = 0;
  for (int b = 0; b < m.B; ++b) {
    e.hostChunkBase[b] = (int)total;
    total += (size_t)m.K * e.hostChunkCount[b];
  }
  e.chunkTotal = total;
  if (!e.chunkBase) {
    TTB2_CUDA_CHECK(cudaMalloc((void**)&e.chunkBase, (m.B + 1) * sizeof(int)));
    TTB2_CUDA_CHECK(cudaMalloc((void**)&e.chunkCount, (m.B + 1) * sizeof(int)));
    e.deviceBytes += 2 * (int64_t)(m.B + 1) * sizeof(int);
  }
  TTB2_CUDA_CHECK(cudaStreamSynchronize(e.stream));
  TTB2_CUDA_CHECK(cudaMemcpy(e.chunkBase, e.hostChunkBase.data(), (m.B + 1) * sizeof(int),
                             cudaMemcpyHostToDevice));
  TTB2_CUDA_CHECK(cudaMemcpy(e.chunkCount, e.hostChunkCount.data(), (m.B + 1) * sizeof(int),
                             cudaMemcpyHostToDevice));
  e.chunkPlanDraws = draws;
  return TTB2_OK;
}

size_t planned_gpart_doubles(const Engine& e, int draws) {
  return (size_t)draws * e.chunkTotal * e.dm.S * e.dm.S;
}

int small_pmatrix(Engine& e, int draws) {
  const Dims& m = e.dm;
  dim3 grid(m.B * m.K, draws);
  if (m.S == 4) {
    const long items = (long)draws * m.B * m.K;
    pmatrix4_kernel<<<(unsigned)((items + 15) / 16), 256, 0, e.stream>>>(
        e.bl, e.rates, e.rateDraws, e.evec, e.ivec, e.eval, e.eigDraws, e.mats, m.B, m.K, draws);
    ++e.launches;
    TTB2_CUDA_CHECK(cudaGetLastError());
    return TTB2_OK;
  }
  if (m.S > 32) {
    const int LD = (m.S + 3) & ~3;
    const size_t smem = (2 * (size_t)m.S * LD + m.S) * sizeof(double);
    if (smem > 48 * 1024)
      TTB2_CUDA_CHECK(cudaFuncSetAttribute(pmatrix_tiled_kernel,
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    pmatrix_tiled_kernel<<<grid, 512, smem, e.stream>>>(
        e.bl, e.rates, e.rateDraws, e.evec, e.ivec, e.eval, e.eigDraws, e.mats, m.B, m.K, m.S);
    ++e.launches;
    TTB2_CUDA_CHECK(cudaGetLastError());
    return TTB2_OK;
  }
  const int threads = round_threads(m.S * m.S, 256);
  pmatrix_kernel<<<grid, threads, m.S * sizeof(double), e.stream>>>(
      e.bl, e.rates, e.rateDraws, e.evec, e.ivec, e.eval, e.eigDraws, e.mats, m.B, m.K, m.S);
  ++e.launches;
  TTB2_CUDA_CHECK(cudaGetLastError());
  return TTB2_OK;
}

int small_gather_inputs(Engine& e, const double* const* src, const size_t* off, const size_t* n,
                        int nseg) {
  GatherArgs a{};
  size_t longest = 1;
  for (int j = 0; j < nseg; ++j) {
    a.src[j] = src[j];
    a.off[j] = off[j];
    a.n[j] = n[j];
    if (n[j] > longest) longest = n[j];
  }
  const unsigned bx = (unsigned)std::min<size_t>((longest + 255) / 256, 64);
  gather_inputs_kernel<<<dim3(bx, nseg), 256, 0, e.stream>>>(a, e.inPacked);
  ++e.launches;
  TTB2_CUDA_CHECK(cudaGetLastError());
  return TTB2_OK;
}

int small_reduce_lnl(Engine& e, int draws, int nblocks) {
  dim3 grid(1, draws);
  reduce_rows_kernel<<<grid, RED_THREADS, 0, e.stream>>>(e.redPart, e.lnl, nblocks, 1, 0);
  ++e.launches;
  TTB2_CUDA_CHECK(cudaGetLastError());
  return TTB2_OK;
}

int small_root_grad_reduce(Engine& e, int draws, int nblocks) {
  const Dims& m = e.dm;
  dim3 grid(m.K + m.S, draws);
  reduce_rows_kernel<<<grid, RED_THREADS, 0, e.stream>>>(e.redPart, e.rootGrad, nblocks,
                                                          m.K + m.S, 0);
  ++e.launches;
  TTB2_CUDA_CHECK(cudaGetLastError());
  return TTB2_OK;
}

int small_gpart_reduce(Engine& e, int draws) {
  const Dims& m = e.dm;
  if (e.deferGpart && m.S * m.S < GR_THREADS) {   // small alphabets: the eigen contraction that follows reduces the chunks itself
    e.gpartPending = true;
    return TTB2_OK;
  }
  const size_t items = (size_t)draws * m.B * m.K;
  const int SS = m.S * m.S;
  const dim3 grid((unsigned)items, SS >= GR_THREADS ? (SS + GR_THREADS - 1) / GR_THREADS : 1);
  gpart_reduce_kernel<<<grid, GR_THREADS, GR_THREADS * sizeof(double), e.stream>>>(
      e.gpart, e.chunkBase, e.chunkCount, e.dmat, e.chunkTotal, m.B, m.K, SS);
  ++e.launches;
  TTB2_CUDA_CHECK(cudaGetLastError());
  return TTB2_OK;
}

int small_scale_dmat(Engine& e, int draws, double* out) {
  const Dims& m = e.dm;
  const size_t width = (size_t)m.B * m.K * m.S * m.S;
  const size_t total = width * draws;
  const int threads = 256;
  const unsigned blocks = (unsigned)((total + threads - 1) / threads);
  scale_rows_kernel<<<blocks, threads, 0, e.stream>>>(e.dmat, e.gradLnl, out, width, draws);
  ++e.launches;
  TTB2_CUDA_CHECK(cudaGetLastError());
  return TTB2_OK;
}

// rootGrad [D][K+S] -> outProps [propDraws][K], outFreqs [freqDraws][S] (one launch: block y = 0 / 1)
int small_root_outputs(Engine& e, int draws) {
  const Dims& m = e.dm;
  const int w = m.K + m.S;
  // per-draw outputs: one thread each; shared outputs: one warp each
  const int nP = e.propDraws > 1 ? m.K * draws : m.K * 32, nF = e.freqDraws > 1 ? m.S * draws : m.S * 32;
  dim3 grid(((nP > nF ? nP : nF) + 127) / 128, 2);
  root_outputs_kernel<<<grid, 128, 0, e.stream>>>(e.rootGrad, e.gradLnl, e.outProps, e.outFreqs,
                                                  m.K, m.S, w, draws, e.propDraws > 1 ? draws : 1,
                                                  e.freqDraws > 1 ? draws : 1);
  ++e.launches;
  TTB2_CUDA_CHECK(cudaGetLastError());
  return TTB2_OK;
}

int small_eigen_contract(Engine& e, int draws) {
  const Dims& m = e.dm;
  const int SS = m.S * m.S;
  {
    dim3 grid(m.B * m.K, draws);
    const bool fused = e.gpartPending;
    e.gpartPending = false;
    const int threads = fused ? 256 : round_threads(SS, 256);
    if (m.S == 4) {
      const long items = (long)draws * m.B * m.K;
      eigen_contract4_kernel<<<(unsigned)((items + 15) / 16), 256, 0, e.stream>>>(
          e.dmat, fused ? e.gpart : nullptr, e.chunkBase, e.chunkCount, e.chunkTotal, e.bl, e.rates,
          e.rateDraws, e.evec, e.ivec, e.eval, e.eigDraws, e.hpart, e.gscal, m.B, m.K, draws);
    } else if (m.S > 32 && !fused) {
      const int LD = (m.S + 3) & ~3;
      const size_t smemT = (4 * (size_t)m.S * LD + 2 * m.S) * sizeof(double);
      if (smemT > 48 * 1024)
        TTB2_CUDA_CHECK(cudaFuncSetAttribute(eigen_contract_tiled_kernel,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)smemT));
      eigen_contract_tiled_kernel<<<grid, 1024, smemT, e.stream>>>(
          e.dmat, e.bl, e.rates, e.rateDraws, e.evec, e.ivec, e.eval, e.eigDraws, e.hpart, e.gscal,
          m.B, m.K, m.S);
    } else {
    const size_t smem = (4 * (size_t)SS + 2 * m.S + (fused ? threads : 0)) * sizeof(double);
    if (smem > 48 * 1024)
      TTB2_CUDA_CHECK(cudaFuncSetAttribute(eigen_contract_kernel,
                                           cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)smem));
    eigen_contract_kernel<<<grid, threads, smem, e.stream>>>(
        e.dmat, fused ? e.gpart : nullptr, e.chunkBase, e.chunkCount, e.chunkTotal, e.bl, e.rates, e.rateDraws, e.evec, e.ivec, e.eval, e.eigDraws, e.hpart,
        e.gscal, m.B, m.K, m.S);
    }
    ++e.launches;
  }
  {
    const int n = m.B * draws;
    branch_grad_kernel<<<(n + 127) / 128, 128, 0, e.stream>>>(
        e.gscal, e.rates, e.rateDraws, e.gradLnl, e.outBl, m.B, m.K, draws);
    ++e.launches;
  }
  {
    const int od = e.rateDraws > 1 ? draws : 1;
    dim3 grid(m.K, od);
    rate_grad_kernel<<<grid, (draws > 8 && e.rateDraws <= 1) ? 1024 : RED_THREADS, 0, e.stream>>>(e.gscal, e.bl, e.gradLnl, e.outRates,
                                                        m.B, m.K, draws, od);
    ++e.launches;
  }
  {
    const int od = e.eigDraws > 1 ? draws : 1;
    dim3 grid(SS, od, H_SLICES);
    h_reduce_kernel<<<grid, RED_THREADS, 0, e.stream>>>(e.hpart, e.gradLnl, e.hred,
                                                       m.B * m.K, SS, draws, od);
    ++e.launches;
    const int threads = round_threads(SS, 256);
    const size_t smem = 4 * (size_t)SS * sizeof(double);
    if (smem > 48 * 1024)
      TTB2_CUDA_CHECK(cudaFuncSetAttribute(q_grad_kernel,
                                           cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)smem));
    q_grad_kernel<<<od, threads, smem, e.stream>>>(e.hred, e.evec, e.ivec, e.outQ, m.S);
    ++e.launches;
  }
  TTB2_CUDA_CHECK(cudaGetLastError());
  return small_root_outputs(e, draws);
}

// Gradient outputs of the matrix-exponential route (expm.cu): the Frechet adjoint per branch x
// category, then the same reductions as the eigen route -- d_bl, d_rates, d lnL / d Q (no change
// of basis), d_props, d_freqs.
int small_expm_contract(Engine& e, int draws) {
  const Dims& m = e.dm;
  const int SS = m.S * m.S;
  int rc = small_expm_backward(e, draws);
  if (rc) return rc;
  {
    const int n = m.B * draws;
    branch_grad_kernel<<<(n + 127) / 128, 128, 0, e.stream>>>(
        e.gscal, e.rates, e.rateDraws, e.gradLnl, e.outBl, m.B, m.K, draws);
    ++e.launches;
  }
  {
    const int od = e.rateDraws > 1 ? draws : 1;
    dim3 grid(m.K, od);
    rate_grad_kernel<<<grid, (draws > 8 && e.rateDraws <= 1) ? 1024 : RED_THREADS, 0, e.stream>>>(e.gscal, e.bl, e.gradLnl, e.outRates,
                                                        m.B, m.K, draws, od);
    ++e.launches;
  }
  {
    const int od = e.eigDraws > 1 ? draws : 1;
    dim3 grid(SS, od, H_SLICES);
    h_reduce_kernel<<<grid, RED_THREADS, 0, e.stream>>>(e.hpart, e.gradLnl, e.hred,
                                                       m.B * m.K, SS, draws, od);
    ++e.launches;
    sum_slices_kernel<<<dim3((SS + 127) / 128, od), 128, 0, e.stream>>>(e.hred, e.outQ, SS);
    ++e.launches;
  }
  TTB2_CUDA_CHECK(cudaGetLastError());
  return small_root_outputs(e, draws);
}

}  // namespace ttb2
