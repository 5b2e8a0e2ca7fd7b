// Batched symmetric eigen-decomposition of reversible generators on the device
// (SURVEY 8(f) row f4).
//
// Replaces the host round trip of SymmetricSubstitutionModel.p_t
// (torchtree/evolution/substitution_model/abstract.py:57-66):
//     sym = sqrt(pi) Q sqrt(pi)^-1;  e, U = eigh(sym);  V = sqrt(pi)^-1 U;  V^-1 = U^T sqrt(pi)
// with one CTA per generator running a parallel cyclic Jacobi iteration in shared memory:
// every step rotates n/2 disjoint index pairs (round-robin tournament ordering, n-1 steps
// per sweep); A <- J^T A J is applied block-wise in one pass (each 2 x 2 block of a row pair x
// column pair is rotated from both sides by one thread), U <- U J alongside: two barriers per step.  Jacobi converges
// quadratically and delivers eigenvalues / vectors at least as accurately as LAPACK's
// syevd (which the reference calls through torch.linalg.eigh); 6-9 sweeps for S = 4..64.
// Like eigh's default, only the lower triangle of `sym` is read.  Eigenvalues come out in
// ascending order (eigh's convention); the sign of an eigenvector is immaterial because
// V and V^-1 are produced as a pair.  S <= 64.
#include "engine.cuh"

namespace ttb2 {

namespace {

constexpr int EIGH_THREADS = 256;
constexpr int EIGH_MAX_SWEEPS = 30;

// pair k (0 <= k < n/2) of step r (0 <= r < n-1) of the round-robin schedule on n (even) indices
__device__ __forceinline__ void rr_pair(int n, int r, int k, int& p, int& q) {
  if (k == 0) {
    p = r % (n - 1);
    q = n - 1;
  } else {
    p = (r + k) % (n - 1);
    q = (r - k + (n - 1)) % (n - 1);
  }
}

__global__ void __launch_bounds__(EIGH_THREADS)
sym_eigh_kernel(const double* __restrict__ q, int qDraws, const double* __restrict__ freqs,
                int freqDraws, double* __restrict__ evec, double* __restrict__ ivec,
                double* __restrict__ eval, int S) {
  extern __shared__ double sm[];
  const int n = S + (S & 1);  // an odd S gets a decoupled dummy index (zero row / column)
  const int ld = n + 1;       // odd leading dimension: conflict-free row and column walks
  double* A = sm;             // [n][ld]
  double* U = A + n * ld;     // [n][ld]
  double* rot = U + n * ld;   // [n/2][2] (c, s) of the current step
  double* root = rot + n;     // [n] sqrt(pi)
  double* red = root + n;     // [2 * 8] block reduction scratch
  __shared__ int rank[64];
  __shared__ int pairs[64];  // (p, q) of every pair of the current step
  __shared__ int done;

  const int d = blockIdx.x;
  const double* qd = q + (size_t)(qDraws == 1 ? 0 : d) * S * S;
  const double* fd = freqs + (size_t)(freqDraws == 1 ? 0 : d) * S;
  const int tid = threadIdx.x;

  for (int i = tid; i < n; i += EIGH_THREADS) root[i] = i < S ? sqrt(fd[i]) : 1.0;
  __syncthreads();
  for (int e = tid; e < n * n; e += EIGH_THREADS) {
    const int i = e / n, j = e % n;
    double a = 0.0;
    if (i < S && j < S) {
      const int hi = max(i, j), lo = min(i, j);  // lower triangle, mirrored
      a = root[hi] * qd[hi * S + lo] / root[lo];
    }
    A[i * ld + j] = a;
    U[i * ld + j] = i == j ? 1.0 : 0.0;
  }
  __syncthreads();

  const int half = n / 2;  // <= 32: one lane per pair
  const int warp = tid >> 5, lane = tid & 31;
  constexpr int NWARP = EIGH_THREADS / 32;
  for (int sweep = 0; sweep < EIGH_MAX_SWEEPS; ++sweep) {
    // off-diagonal and total Frobenius norms, summed directly (no cancellation)
    double off = 0.0, tot = 0.0;
    for (int i = warp; i < n; i += NWARP)
      for (int j = lane; j < n; j += 32) {
        const double a = A[i * ld + j];
        tot += a * a;
        if (i != j) off += a * a;
      }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      off += __shfl_xor_sync(0xffffffffu, off, o);
      tot += __shfl_xor_sync(0xffffffffu, tot, o);
    }
    if (lane == 0) {
      red[warp] = off;
      red[8 + warp] = tot;
    }
    __syncthreads();
    if (tid == 0) {
      double o = 0.0, t = 0.0;
      for (int w = 0; w < NWARP; ++w) {
        o += red[w];
        t += red[8 + w];
      }
      // NaN input: the comparison is false for ever, the sweep cap ends the loop and the
      // NaNs propagate into P (NaN in -> NaN out)
      done = o <= 1e-32 * t;
    }
    __syncthreads();
    if (done) break;

    for (int r = 0; r < n - 1; ++r) {
      // rotation of every pair of this step, from the current A
      if (tid < half) {
        int p, qq;
        rr_pair(n, r, tid, p, qq);
        const double apq = A[p * ld + qq];
        double c = 1.0, s = 0.0;
        if (!(fabs(apq) <= 1e-300)) {  // also taken by NaN, which must propagate
          const double theta = (A[qq * ld + qq] - A[p * ld + p]) / (2.0 * apq);
          double t;
          if (fabs(theta) > 1e150)
            t = 0.5 / theta;
          else
            t = copysign(1.0, theta) / (fabs(theta) + sqrt(theta * theta + 1.0));
          c = 1.0 / sqrt(t * t + 1.0);
          s = t * c;
        }
        rot[2 * tid] = c;
        rot[2 * tid + 1] = s;
        pairs[2 * tid] = p;
        pairs[2 * tid + 1] = qq;
      }
      __syncthreads();
      // A <- J^T A J in one pass: the 2 x 2 block (pair ki, pair kj) is read once, rotated from
      // both sides and written once; every element belongs to exactly one block.
      // Lane = column pair, warps stride over the row pairs.
      if (lane < half) {
        const int pj = pairs[2 * lane], qj = pairs[2 * lane + 1];
        const double cj = rot[2 * lane], sj = rot[2 * lane + 1];
        for (int ki = warp; ki < half; ki += NWARP) {
          const int pi = pairs[2 * ki], qi = pairs[2 * ki + 1];
          const double ci = rot[2 * ki], si = rot[2 * ki + 1];
          if (si != 0.0 || sj != 0.0) {
            const double a = A[pi * ld + pj], b = A[pi * ld + qj];
            const double cc = A[qi * ld + pj], d = A[qi * ld + qj];
            const double a1 = ci * a - si * cc, c1 = si * a + ci * cc;
            const double b1 = ci * b - si * d, d1 = si * b + ci * d;
            const bool diag = ki == lane;  // the annihilated pair is set to exactly zero
            A[pi * ld + pj] = cj * a1 - sj * b1;
            A[pi * ld + qj] = diag ? 0.0 : sj * a1 + cj * b1;
            A[qi * ld + pj] = diag ? 0.0 : cj * c1 - sj * d1;
            A[qi * ld + qj] = sj * c1 + cj * d1;
          }
        }
        // U <- U J: lane = column pair, warps stride over the rows
        if (sj != 0.0) {
          for (int i = warp; i < n; i += NWARP) {
            const double up = U[i * ld + pj], uq = U[i * ld + qj];
            U[i * ld + pj] = cj * up - sj * uq;
            U[i * ld + qj] = sj * up + cj * uq;
          }
        }
      }
      __syncthreads();
    }
  }

  // ascending order (ties by index): rank[j] = position of eigenvalue j
  for (int j = tid; j < S; j += EIGH_THREADS) {
    const double lj = A[j * ld + j];
    int rk = 0;
    for (int i = 0; i < S; ++i) {
      const double li = A[i * ld + i];
      rk += (li < lj) || (li == lj && i < j);
    }
    // NaN eigenvalues compare false everywhere: keep them in place
    rank[j] = (lj == lj) ? rk : j;
    eval[(size_t)d * S + rank[j]] = lj;
  }
  __syncthreads();
  double* ev = evec + (size_t)d * S * S;
  double* iv = ivec + (size_t)d * S * S;
  for (int e = tid; e < S * S; e += EIGH_THREADS) {
    const int i = e / S, j = e % S;
    const double u = U[i * ld + j];
    ev[i * S + rank[j]] = u / root[i];
    iv[rank[j] * S + i] = u * root[i];
  }
}

}  // namespace

// e.qnorm [qDraws][S][S], e.freqs [freqDraws][S] -> e.evec / e.ivec / e.eval [eigDraws]
int small_sym_eigh(Engine& e, int qDraws, int eigDraws) {
  const Dims& m = e.dm;
  if (m.S > 64) {
    set_error("device eigen-decomposition supports at most 64 states");
    return TTB2_E_INVALID;
  }
  const int n = m.S + (m.S & 1);
  const size_t smem = ((size_t)2 * n * (n + 1) + 2 * n + 16) * sizeof(double);
  if (smem > 48 * 1024)
    TTB2_CUDA_CHECK(cudaFuncSetAttribute(sym_eigh_kernel,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  sym_eigh_kernel<<<eigDraws, EIGH_THREADS, smem, e.stream>>>(
      e.qnorm, qDraws, e.freqs, e.freqDraws, e.evec, e.ivec, e.eval, m.S);
  ++e.launches;
  TTB2_CUDA_CHECK(cudaGetLastError());
  return TTB2_OK;
}

}  // namespace ttb2
