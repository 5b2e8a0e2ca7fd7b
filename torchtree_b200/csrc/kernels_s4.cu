// 4-state (nucleotide) level-synchronous peeling kernels for sm_100a.
//
// Layout: conditional-likelihood vectors are stored [draw][inode][k][pattern][4]
// so that one pattern's 4-state vector is one 32-byte fp64x4 access
// (LDG/STG.256) and a warp touches 1 KB of contiguous memory per category.
// Patterns map to threads.  Per-branch transition matrices (or, for tip
// children, the K x C table P . tipvector[code]) are staged in shared memory.
// Rescaling uses exact power-of-two factors: s = 2^e with e taken from the
// exponent field of max_{k,s} partial, stored as int16 per (node, pattern).
//
// Replaces: calculate_treelikelihood_discrete_rescaled and the tip-state
// variants (torchtree/evolution/tree_likelihood.py:186-278) for the post-order
// pass, and the autograd tape by the pre-order pass of
// SURVEY Appendix B.
#include <cstdlib>
#include <vector>

#include <algorithm>

#include "engine.cuh"

namespace ttb2 {

namespace {

constexpr int FWD_THREADS = 128;
constexpr int BWD_THREADS = 256;
constexpr int ROOT_THREADS = 128;

struct __align__(32) V4 {
  double x, y, z, w;
};

__device__ __forceinline__ V4 ldg4(const double* p) {
  V4 v;
  asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];"
               : "=d"(v.x), "=d"(v.y), "=d"(v.z), "=d"(v.w)
               : "l"(p));
  return v;
}

// coherent variant: for vectors this thread wrote earlier in the same kernel (chain mode)
__device__ __forceinline__ V4 ldg4_coherent(const double* p) {
  V4 v;
  asm volatile("ld.global.v4.f64 {%0,%1,%2,%3}, [%4];"
               : "=d"(v.x), "=d"(v.y), "=d"(v.z), "=d"(v.w)
               : "l"(p)
               : "memory");
  return v;
}

__device__ __forceinline__ void stg4(double* p, const V4& v) {
  asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(v.x),
               "d"(v.y), "d"(v.z), "d"(v.w)
               : "memory");
}

__device__ __forceinline__ V4 lds4(const double* p) {
  const double2 a = *reinterpret_cast<const double2*>(p);
  const double2 b = *reinterpret_cast<const double2*>(p + 2);
  return V4{a.x, a.y, b.x, b.y};
}

// u = P v, P row-major 4x4 in shared memory (warp-uniform address: broadcast)
__device__ __forceinline__ V4 matvec(const double* P, const V4& v) {
  V4 u;
  u.x = fma(P[3], v.w, fma(P[2], v.z, fma(P[1], v.y, P[0] * v.x)));
  u.y = fma(P[7], v.w, fma(P[6], v.z, fma(P[5], v.y, P[4] * v.x)));
  u.z = fma(P[11], v.w, fma(P[10], v.z, fma(P[9], v.y, P[8] * v.x)));
  u.w = fma(P[15], v.w, fma(P[14], v.z, fma(P[13], v.y, P[12] * v.x)));
  return u;
}

// u = P^T v
__device__ __forceinline__ V4 matvec_t(const double* P, const V4& v) {
  V4 u;
  u.x = fma(P[12], v.w, fma(P[8], v.z, fma(P[4], v.y, P[0] * v.x)));
  u.y = fma(P[13], v.w, fma(P[9], v.z, fma(P[5], v.y, P[1] * v.x)));
  u.z = fma(P[14], v.w, fma(P[10], v.z, fma(P[6], v.y, P[2] * v.x)));
  u.w = fma(P[15], v.w, fma(P[11], v.z, fma(P[7], v.y, P[3] * v.x)));
  return u;
}

__device__ __forceinline__ V4 mul4(const V4& a, const V4& b) {
  return V4{a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w};
}

__device__ __forceinline__ V4 scale4(const V4& a, double f) {
  return V4{a.x * f, a.y * f, a.z * f, a.w * f};
}

__device__ __forceinline__ double max4(const V4& a) {
  return fmax(fmax(a.x, a.y), fmax(a.z, a.w));
}

// Shared-memory table of one child for all K categories.
//   internal child: tab[k][16] = P_k (row-major)
//   tip child     : tab[k][code][4] = P_k . codeP[code]
template <int K>
__device__ __forceinline__ void build_child_table(double* tab, const double* P,
                                                  bool tip, const double* codeP,
                                                  int C) {
  if (!tip) {
    for (int j = threadIdx.x; j < K * 16; j += blockDim.x) tab[j] = P[j];
  } else {
    for (int j = threadIdx.x; j < K * C * 4; j += blockDim.x) {
      const int s = j & 3;
      const int code = (j >> 2) % C;
      const int k = (j >> 2) / C;
      const double* row = P + k * 16 + s * 4;
      const double* cp = codeP + code * 4;
      tab[j] = fma(row[3], cp[3], fma(row[2], cp[2], fma(row[1], cp[1], row[0] * cp[0])));
    }
  }
}

__host__ __device__ inline int child_table_doubles(int K, int C) {
  return K * (C > 4 ? C : 4) * 4;
}

// ---------------------------------------------------------------------------
// post-order: one launch per level, grid (pattern blocks, nodes of level, draws)
// ---------------------------------------------------------------------------
template <int K>
__global__ void __launch_bounds__(FWD_THREADS, (K <= 4 ? 4 : 2))
fwd4_kernel(const NodeOp* __restrict__ ops, int opBegin,
            const double* __restrict__ mats, const uint8_t* __restrict__ tips,
            const double* __restrict__ codeP, double* __restrict__ partials,
            int16_t* __restrict__ expo, int T, int Npad, int C, int B, int ppt) {
  extern __shared__ double sm[];
  const NodeOp op = ops[opBegin + blockIdx.y];
  const int d = blockIdx.z;
  const int I = T - 1;
  const int tabN = child_table_doubles(K, C);
  double* tabL = sm;
  double* tabR = sm + tabN;
  const double* matsD = mats + (size_t)d * B * K * 16;
  const bool tipL = op.left < T, tipR = op.right < T;
  build_child_table<K>(tabL, matsD + (size_t)op.left * K * 16, tipL, codeP, C);
  build_child_table<K>(tabR, matsD + (size_t)op.right * K * 16, tipR, codeP, C);
  __syncthreads();
  pdl_wait_then_trigger();

  const size_t nodeStride = (size_t)K * Npad * 4;
  double* base = partials + (size_t)d * I * nodeStride;
  const int i0 = blockIdx.x * (FWD_THREADS * ppt) + threadIdx.x;
#pragma unroll 1
  for (int it = 0; it < ppt; ++it) {
    const int i = i0 + it * FWD_THREADS;
    if (i >= Npad) break;
    V4 a[K], b[K];
    int codeL = 0, codeR = 0;
    if (tipL) {
      codeL = tips[(size_t)op.left * Npad + i];
    } else {
      const double* p = base + (size_t)(op.left - T) * nodeStride + (size_t)i * 4;
#pragma unroll
      for (int k = 0; k < K; ++k) a[k] = ldg4(p + (size_t)k * Npad * 4);
    }
    if (tipR) {
      codeR = tips[(size_t)op.right * Npad + i];
    } else {
      const double* p = base + (size_t)(op.right - T) * nodeStride + (size_t)i * 4;
#pragma unroll
      for (int k = 0; k < K; ++k) b[k] = ldg4(p + (size_t)k * Npad * 4);
    }

    V4 out[K];
    double m = 0.0;
#pragma unroll
    for (int k = 0; k < K; ++k) {
      const V4 ul = tipL ? lds4(tabL + (k * C + codeL) * 4) : matvec(tabL + k * 16, a[k]);
      const V4 ur = tipR ? lds4(tabR + (k * C + codeR) * 4) : matvec(tabR + k * 16, b[k]);
      out[k] = mul4(ul, ur);
      m = fmax(m, max4(out[k]));
    }
    // exact power-of-two rescaling: m * f in [0.5, 1)
    int eb = (__double2hiint(m) >> 20) & 0x7ff;
    eb = eb > 2044 ? 2044 : eb;
    const double f = __hiloint2double((2045 - eb) << 20, 0);
    double* q = base + (size_t)(op.node - T) * nodeStride + (size_t)i * 4;
#pragma unroll
    for (int k = 0; k < K; ++k) stg4(q + (size_t)k * Npad * 4, scale4(out[k], f));
    expo[((size_t)d * I + (op.node - T)) * Npad + i] = (int16_t)(eb - 1022);
  }
}

// generic-K fallback (K not instantiated): two sweeps over the categories
__global__ void __launch_bounds__(FWD_THREADS)
fwd4_kernel_anyk(const NodeOp* __restrict__ ops, int opBegin,
                 const double* __restrict__ mats, const uint8_t* __restrict__ tips,
                 const double* __restrict__ codeP, double* __restrict__ partials,
                 int16_t* __restrict__ expo, int T, int Npad, int C, int B, int K) {
  extern __shared__ double sm[];
  const NodeOp op = ops[opBegin + blockIdx.y];
  const int d = blockIdx.z;
  const int I = T - 1;
  const int tabN = child_table_doubles(K, C);
  double* tabL = sm;
  double* tabR = sm + tabN;
  const double* matsD = mats + (size_t)d * B * K * 16;
  const bool tipL = op.left < T, tipR = op.right < T;
  for (int side = 0; side < 2; ++side) {
    double* tab = side ? tabR : tabL;
    const int child = side ? op.right : op.left;
    const bool tip = side ? tipR : tipL;
    const double* P = matsD + (size_t)child * K * 16;
    if (!tip) {
      for (int j = threadIdx.x; j < K * 16; j += blockDim.x) tab[j] = P[j];
    } else {
      for (int j = threadIdx.x; j < K * C * 4; j += blockDim.x) {
        const int s = j & 3, code = (j >> 2) % C, k = (j >> 2) / C;
        const double* row = P + k * 16 + s * 4;
        const double* cp = codeP + code * 4;
        tab[j] = fma(row[3], cp[3], fma(row[2], cp[2], fma(row[1], cp[1], row[0] * cp[0])));
      }
    }
  }
  __syncthreads();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= Npad) return;
  const size_t nodeStride = (size_t)K * Npad * 4;
  double* base = partials + (size_t)d * I * nodeStride;
  const int codeL = tipL ? tips[(size_t)op.left * Npad + i] : 0;
  const int codeR = tipR ? tips[(size_t)op.right * Npad + i] : 0;
  const double* pl = base + (size_t)(tipL ? 0 : op.left - T) * nodeStride + (size_t)i * 4;
  const double* pr = base + (size_t)(tipR ? 0 : op.right - T) * nodeStride + (size_t)i * 4;
  double* q = base + (size_t)(op.node - T) * nodeStride + (size_t)i * 4;
  double m = 0.0;
  for (int k = 0; k < K; ++k) {
    const V4 ul = tipL ? lds4(tabL + (k * C + codeL) * 4)
                       : matvec(tabL + k * 16, ldg4(pl + (size_t)k * Npad * 4));
    const V4 ur = tipR ? lds4(tabR + (k * C + codeR) * 4)
                       : matvec(tabR + k * 16, ldg4(pr + (size_t)k * Npad * 4));
    const V4 o = mul4(ul, ur);
    m = fmax(m, max4(o));
    stg4(q + (size_t)k * Npad * 4, o);
  }
  int eb = (__double2hiint(m) >> 20) & 0x7ff;
  eb = eb > 2044 ? 2044 : eb;
  const double f = __hiloint2double((2045 - eb) << 20, 0);
  for (int k = 0; k < K; ++k) {
    double* qq = q + (size_t)k * Npad * 4;
    V4 o;
    // plain (coherent) loads: this thread wrote these values above
    o.x = qq[0]; o.y = qq[1]; o.z = qq[2]; o.w = qq[3];
    stg4(qq, scale4(o, f));
  }
  expo[((size_t)d * I + (op.node - T)) * Npad + i] = (int16_t)(eb - 1022);
}

// ---------------------------------------------------------------------------
// block-level sum of one double per thread; result valid in thread 0
// ---------------------------------------------------------------------------
__device__ __forceinline__ double block_sum(double v, double* red) {
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

// ---------------------------------------------------------------------------
// root: site log-likelihoods + per-block weighted partial sums
//   lnL_i = log(sum_k rho_k pi . p~_root[k,:,i]) + ln2 * sum_n e_n[i]
// (tree_likelihood.py:215-221)
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(ROOT_THREADS)
root4_kernel(const double* __restrict__ partials, const int16_t* __restrict__ expo,
             const double* __restrict__ freqs, int freqDraws,
             const double* __restrict__ props, int propDraws,
             const double* __restrict__ weights, double* __restrict__ siteLnl,
             double* __restrict__ blockPart, int T, int Npad, int K, int rootInode) {
  __shared__ double red[ROOT_THREADS / 32];
  const int d = blockIdx.y;
  const int I = T - 1;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const double* fr = freqs + (freqDraws > 1 ? (size_t)d * 4 : 0);
  const double* pr = props + (propDraws > 1 ? (size_t)d * K : 0);
  double contrib = 0.0;
  if (i < Npad) {
    const size_t nodeStride = (size_t)K * Npad * 4;
    const double* p = partials + ((size_t)d * I + rootInode) * nodeStride + (size_t)i * 4;
    double L = 0.0;
    for (int k = 0; k < K; ++k) {
      const V4 v = ldg4(p + (size_t)k * Npad * 4);
      const double dot = fma(fr[3], v.w, fma(fr[2], v.z, fma(fr[1], v.y, fr[0] * v.x)));
      L = fma(pr[k], dot, L);
    }
    int esum = 0;
    const int16_t* e = expo + (size_t)d * I * Npad + i;
    for (int n = 0; n < I; ++n) esum += e[(size_t)n * Npad];
    const double site = log(L) + (double)esum * 0.693147180559945309417232121458;
    siteLnl[(size_t)d * Npad + i] = site;
    const double w = weights[i];
    contrib = (w != 0.0) ? w * site : 0.0;
  }
  const double t = block_sum(contrib, red);
  if (threadIdx.x == 0) blockPart[(size_t)d * gridDim.x + blockIdx.x] = t;
}

// The same for small shards.  Block = 32 patterns x ROOT_SLICES node slices: the sum of
// the I scale exponents of a pattern (I dependent-free loads per pattern, the long part)
// is split over the slices so that a few thousand patterns still put enough loads in flight.
constexpr int ROOT_SLICES = 32;

__global__ void __launch_bounds__(32 * ROOT_SLICES)
root4_small_kernel(const double* __restrict__ partials, const int16_t* __restrict__ expo,
             const double* __restrict__ freqs, int freqDraws,
             const double* __restrict__ props, int propDraws,
             const double* __restrict__ weights, double* __restrict__ siteLnl,
             double* __restrict__ blockPart, int T, int Npad, int K, int rootInode) {
  __shared__ int esums[ROOT_SLICES][32];
  const int d = blockIdx.y;
  const int I = T - 1;
  const int lane = threadIdx.x, slice = threadIdx.y;
  const int i = blockIdx.x * 32 + lane;      // Npad is a multiple of 32
  const double* fr = freqs + (freqDraws > 1 ? (size_t)d * 4 : 0);
  const double* pr = props + (propDraws > 1 ? (size_t)d * K : 0);
  int esum = 0;
  const int16_t* e = expo + (size_t)d * I * Npad + i;
#pragma unroll 4
  for (int n = slice; n < I; n += ROOT_SLICES) esum += e[(size_t)n * Npad];
  esums[slice][lane] = esum;
  __syncthreads();
  if (slice != 0) return;
  esum = 0;
#pragma unroll
  for (int j = 0; j < ROOT_SLICES; ++j) esum += esums[j][lane];
  const size_t nodeStride = (size_t)K * Npad * 4;
  const double* p = partials + ((size_t)d * I + rootInode) * nodeStride + (size_t)i * 4;
  double L = 0.0;
  for (int k = 0; k < K; ++k) {
    const V4 v = ldg4(p + (size_t)k * Npad * 4);
    const double dot = fma(fr[3], v.w, fma(fr[2], v.z, fma(fr[1], v.y, fr[0] * v.x)));
    L = fma(pr[k], dot, L);
  }
  const double site = log(L) + (double)esum * 0.693147180559945309417232121458;
  siteLnl[(size_t)d * Npad + i] = site;
  const double w = weights[i];
  double t = (w != 0.0) ? w * site : 0.0;
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) t += __shfl_xor_sync(0xffffffffu, t, off);
  if (lane == 0) blockPart[(size_t)d * gridDim.x + blockIdx.x] = t;
}

// ---------------------------------------------------------------------------
// pre-order root: q^_root[k,s,i] = rho_k pi_s / (L~_i * 2^{e_root});
// per-block partials of d lnL / d rho_k and the root term of d lnL / d pi_s
// 
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(ROOT_THREADS)
root4_bwd_kernel(const double* __restrict__ partials, const int16_t* __restrict__ expo,
                 const double* __restrict__ freqs, int freqDraws,
                 const double* __restrict__ props, int propDraws,
                 const double* __restrict__ weights, double* __restrict__ pre,
                 double* __restrict__ blockPart, int T, int Npad, int K, int rootInode) {
  __shared__ double red[ROOT_THREADS / 32];
  const int d = blockIdx.y;
  const int I = T - 1;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const double* fr = freqs + (freqDraws > 1 ? (size_t)d * 4 : 0);
  const double* pr = props + (propDraws > 1 ? (size_t)d * K : 0);
  const size_t nodeStride = (size_t)K * Npad * 4;
  const bool live = i < Npad;
  double w = 0.0, invL = 0.0;
  const double* p = nullptr;
  if (live) {
    p = partials + ((size_t)d * I + rootInode) * nodeStride + (size_t)i * 4;
    double L = 0.0;
    for (int k = 0; k < K; ++k) {
      const V4 v = ldg4(p + (size_t)k * Npad * 4);
      const double dot = fma(fr[3], v.w, fma(fr[2], v.z, fma(fr[1], v.y, fr[0] * v.x)));
      L = fma(pr[k], dot, L);
    }
    invL = 1.0 / L;
    w = weights[i];
    const int e = expo[((size_t)d * I + rootInode) * Npad + i];
    // zero-weight patterns (padding, or masked by the caller) get q^ = 0 here, so every
    // pre-order quantity below the root is exactly 0 for them without further tests
    const double scale = (w != 0.0) ? invL * __hiloint2double((1023 - e) << 20, 0) : 0.0;
    double* q = pre + ((size_t)d * I + rootInode) * nodeStride + (size_t)i * 4;
    for (int k = 0; k < K; ++k) {
      const double c = pr[k] * scale;
      stg4(q + (size_t)k * Npad * 4, V4{c * fr[0], c * fr[1], c * fr[2], c * fr[3]});
    }
  }
  // d/d rho_k = sum_i w_i (pi . p~[k]) / L~ ; d/d pi_s = sum_i w_i sum_k rho_k p~[k,s] / L~
  const double wl = (w != 0.0) ? w * invL : 0.0;
  double* out = blockPart + ((size_t)d * gridDim.x + blockIdx.x) * (K + 4);
  V4 accF{0.0, 0.0, 0.0, 0.0};
  for (int k = 0; k < K; ++k) {
    V4 v{0.0, 0.0, 0.0, 0.0};
    if (live && wl != 0.0) v = ldg4(p + (size_t)k * Npad * 4);
    const double dot = fma(fr[3], v.w, fma(fr[2], v.z, fma(fr[1], v.y, fr[0] * v.x)));
    const double t = block_sum(wl * dot, red);
    if (threadIdx.x == 0) out[k] = t;
    const double c = wl * pr[k];
    accF.x = fma(c, v.x, accF.x);
    accF.y = fma(c, v.y, accF.y);
    accF.z = fma(c, v.z, accF.z);
    accF.w = fma(c, v.w, accF.w);
  }
  double t;
  t = block_sum(accF.x, red); if (threadIdx.x == 0) out[K + 0] = t;
  t = block_sum(accF.y, red); if (threadIdx.x == 0) out[K + 1] = t;
  t = block_sum(accF.z, red); if (threadIdx.x == 0) out[K + 2] = t;
  t = block_sum(accF.w, red); if (threadIdx.x == 0) out[K + 3] = t;
}

// ---------------------------------------------------------------------------
// pre-order level kernel: grid (pattern chunks, nodes of level x K, draws).
// For parent n with children l, r:
//   u_c = P_c p~_c ; m_l = q^_n o u_r ; m_r = q^_n o u_l
//   G_c += w_i m_c (x) p~_c          (= d lnL / d P_c[k], rows = parent state)
//   q^_c = P_c^T m_c * 2^{-e_c}      (internal children only)
// ---------------------------------------------------------------------------
// Sum 32 per-thread values across the warp with a halving butterfly
// (31 shuffles instead of 160); on return lane L holds the total of v[L].
__device__ __forceinline__ double warp_transpose_sum(double (&v)[32]) {
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int half = 16; half >= 1; half >>= 1) {
    const bool upper = (lane & half) != 0;
#pragma unroll
    for (int j = 0; j < half; ++j) {
      const double send = upper ? v[j] : v[j + half];
      const double keep = upper ? v[j + half] : v[j];
      v[j] = keep + __shfl_xor_sync(0xffffffffu, send, half);
    }
  }
  return v[0];
}

// ---------------------------------------------------------------------------
// Level 1 (both children are tips; a third of all nodes of a random tree).
// The generic kernels spend their time in shared-memory lookups and MMA staging
// there, while the only HBM traffic is one vector per unit -- so level 1 gets
// its own kernels.
// ---------------------------------------------------------------------------
// post-order: the node's vector depends only on the pair of tip codes, so the
// CTA builds the table pair[cL][cR] -> (K scaled vectors, exponent) once and
// every pattern is two byte loads, K table reads and K 32-byte stores.
template <int K>
__global__ void __launch_bounds__(FWD_THREADS, 4)
fwd4_tips_kernel(const NodeOp* __restrict__ ops, int opBegin,
                 const double* __restrict__ mats, const uint8_t* __restrict__ tips,
                 const double* __restrict__ codeP, double* __restrict__ partials,
                 int16_t* __restrict__ expo, int T, int Npad, int C, int B, int ppt) {
  extern __shared__ double sm[];
  const NodeOp op = ops[opBegin + blockIdx.y];
  const int d = blockIdx.z;
  const int I = T - 1;
  const int tabN = child_table_doubles(K, C);
  double* tabL = sm;
  double* tabR = sm + tabN;
  // pair table, transposed so that lanes with different code pairs read
  // consecutive 16-byte words: pair2[(k*2 + half)][pc] (double2)
  double2* pair2 = reinterpret_cast<double2*>(tabR + tabN);
  int* pexp = reinterpret_cast<int*>(pair2 + (size_t)C * C * K * 2);  // [C*C]
  const int CC = C * C;
  const double* matsD = mats + (size_t)d * B * K * 16;
  build_child_table<K>(tabL, matsD + (size_t)op.left * K * 16, true, codeP, C);
  build_child_table<K>(tabR, matsD + (size_t)op.right * K * 16, true, codeP, C);
  __syncthreads();
  for (int pc = threadIdx.x; pc < C * C; pc += blockDim.x) {
    const int cl = pc / C, cr = pc - cl * C;
    V4 out[K];
    double m = 0.0;
#pragma unroll
    for (int k = 0; k < K; ++k) {
      out[k] = mul4(lds4(tabL + (k * C + cl) * 4), lds4(tabR + (k * C + cr) * 4));
      m = fmax(m, max4(out[k]));
    }
    int eb = (__double2hiint(m) >> 20) & 0x7ff;
    eb = eb > 2044 ? 2044 : eb;
    const double f = __hiloint2double((2045 - eb) << 20, 0);
#pragma unroll
    for (int k = 0; k < K; ++k) {
      pair2[(k * 2 + 0) * CC + pc] = make_double2(out[k].x * f, out[k].y * f);
      pair2[(k * 2 + 1) * CC + pc] = make_double2(out[k].z * f, out[k].w * f);
    }
    pexp[pc] = eb - 1022;
  }
  __syncthreads();
  pdl_wait_then_trigger();

  const size_t nodeStride = (size_t)K * Npad * 4;
  double* q = partials + ((size_t)d * I + (op.node - T)) * nodeStride;
  int16_t* eo = expo + ((size_t)d * I + (op.node - T)) * Npad;
  const uint8_t* tl = tips + (size_t)op.left * Npad;
  const uint8_t* tr = tips + (size_t)op.right * Npad;
  const int i0 = blockIdx.x * (FWD_THREADS * ppt) + threadIdx.x;
#pragma unroll 2
  for (int it = 0; it < ppt; ++it) {
    const int i = i0 + it * FWD_THREADS;
    if (i >= Npad) break;
    const int pc = (int)tl[i] * C + (int)tr[i];
#pragma unroll
    for (int k = 0; k < K; ++k) {
      const double2 lo = pair2[(k * 2 + 0) * CC + pc];
      const double2 hi = pair2[(k * 2 + 1) * CC + pc];
      stg4(q + ((size_t)k * Npad + i) * 4, V4{lo.x, lo.y, hi.x, hi.y});
    }
    eo[i] = (int16_t)pexp[pc];
  }
}

// ---------------------------------------------------------------------------
// Cherry fusion.  A "cherry" is an internal node whose two children are tips.
// Its (rescaled) vector is a pure function of the pair of tip codes, so it is
// tabulated once per evaluation -- cherryVec[d][c][k][cL*C+cR][4] and the scale
// exponent cherryExp[d][c][cL*C+cR] -- and never stored per pattern: parents
// treat a cherry child like a tip with C*C symbol codes.  On a random tree a
// third of the internal nodes are cherries, so a third of the post-order
// writes/reads and of the pre-order child reads disappear from HBM.
// ---------------------------------------------------------------------------
struct CherryArgs {
  const int* idx;      // [I] cherry index of internal node (node - T), or -1
  const int* info;     // [n][3] (left tip, right tip, node)
  const double* vec;   // [D][n][K][CC][4]
  const int* exps;     // [D][n][CC]
  const uint8_t* code; // [n][Npad] pair code of every pattern
  int n;               // number of cherries (0 = fusion off)
  int CC;              // C * C
};

__global__ void __launch_bounds__(128)
cherry_table_kernel(const int* __restrict__ info, const double* __restrict__ mats,
                    const double* __restrict__ codeP, double* __restrict__ vec,
                    int* __restrict__ exps, int n, int C, int B, int K) {
  extern __shared__ double sm[];  // tabL[K][C][4] tabR[K][C][4]
  const int c = blockIdx.x, d = blockIdx.y;
  const int CC = C * C;
  double* tabL = sm;
  double* tabR = sm + K * C * 4;
  const int tipL = info[c * 3], tipR = info[c * 3 + 1];
  const double* matsD = mats + (size_t)d * B * K * 16;
  for (int j = threadIdx.x; j < 2 * K * C * 4; j += blockDim.x) {
    const int side = j / (K * C * 4);
    const int r = j - side * (K * C * 4);
    const int s = r & 3, code = (r >> 2) % C, k = (r >> 2) / C;
    const double* row = matsD + ((size_t)(side ? tipR : tipL) * K + k) * 16 + s * 4;
    const double* cp = codeP + code * 4;
    (side ? tabR : tabL)[r] = fma(row[3], cp[3], fma(row[2], cp[2], fma(row[1], cp[1], row[0] * cp[0])));
  }
  __syncthreads();
  for (int pc = threadIdx.x; pc < CC; pc += blockDim.x) {
    const int cl = pc / C, cr = pc - cl * C;
    double m = 0.0;
    for (int k = 0; k < K; ++k)
      m = fmax(m, max4(mul4(lds4(tabL + (k * C + cl) * 4), lds4(tabR + (k * C + cr) * 4))));
    int eb = (__double2hiint(m) >> 20) & 0x7ff;
    eb = eb > 2044 ? 2044 : eb;
    const double f = __hiloint2double((2045 - eb) << 20, 0);
    for (int k = 0; k < K; ++k) {
      const V4 o = scale4(mul4(lds4(tabL + (k * C + cl) * 4), lds4(tabR + (k * C + cr) * 4)), f);
      double* dst = vec + ((((size_t)d * n + c) * K + k) * CC + pc) * 4;
      dst[0] = o.x; dst[1] = o.y; dst[2] = o.z; dst[3] = o.w;
    }
    exps[((size_t)d * n + c) * CC + pc] = eb - 1022;
  }
}

// per-pattern scale exponents of the cherries (2 bytes per pattern; the only
// per-pattern data a cherry keeps -- read by the root sum and the pre-order pass).
// A thread translates 16 consecutive pair codes (one 16-byte load, 32 bytes stored).
__global__ void __launch_bounds__(128)
cherry_expo_kernel(const int* __restrict__ info, const int* __restrict__ exps,
                   const uint8_t* __restrict__ code, int16_t* __restrict__ expo, int n, int CC,
                   int T, int Npad) {
  __shared__ int16_t tab[256];
  const int c = blockIdx.y, d = blockIdx.z;
  for (int j = threadIdx.x; j < CC; j += blockDim.x)
    tab[j] = (int16_t)exps[((size_t)d * n + c) * CC + j];
  __syncthreads();
  const int node = info[c * 3 + 2];
  const int i = (blockIdx.x * blockDim.x + threadIdx.x) * 16;   // Npad is a multiple of 32
  if (i >= Npad) return;
  const uint4 in = *reinterpret_cast<const uint4*>(code + (size_t)c * Npad + i);
  const unsigned words[4] = {in.x, in.y, in.z, in.w};
  __align__(16) int16_t out[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) out[j] = tab[(words[j >> 2] >> ((j & 3) * 8)) & 0xff];
  int16_t* dst = expo + ((size_t)d * (T - 1) + (node - T)) * Npad + i;
  reinterpret_cast<uint4*>(dst)[0] = reinterpret_cast<const uint4*>(out)[0];
  reinterpret_cast<uint4*>(dst)[1] = reinterpret_cast<const uint4*>(out)[1];
}

// pair code of every pattern (static: built once per topology)
__global__ void __launch_bounds__(256)
cherry_code_kernel(const int* __restrict__ info, const uint8_t* __restrict__ tips,
                   uint8_t* __restrict__ code, int C, int Npad) {
  const int c = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= Npad) return;
  const int tipL = info[c * 3], tipR = info[c * 3 + 1];
  code[(size_t)c * Npad + i] =
      (uint8_t)((int)tips[(size_t)tipL * Npad + i] * C + (int)tips[(size_t)tipR * Npad + i]);
}

// child kinds
enum { KIND_STORED = 0, KIND_TIP = 1, KIND_CHERRY = 2 };

__device__ __forceinline__ int child_kind(int child, int T, const CherryArgs& ch, int& cidx) {
  cidx = -1;
  if (child < T) return KIND_TIP;
  if (ch.n > 0) {
    cidx = ch.idx[child - T];
    if (cidx >= 0) return KIND_CHERRY;
  }
  return KIND_STORED;
}

// tab[(k*NC + code)*4 + s] = sum_s' P_k[s][s'] vec_k[code][s']   (tip / cherry child)
// tab[k*16 + ..] = P_k                                             (stored child)
template <int K>
__device__ __forceinline__ void build_child_table_c(double* tab, const double* P, int kind,
                                                    const double* codeP, int C,
                                                    const double* cvec, int CC, int NC) {
  if (kind == KIND_STORED) {
    for (int j = threadIdx.x; j < K * 16; j += blockDim.x) tab[j] = P[j];
    return;
  }
  // coded child: two double2 planes per category, tab2[(k*2 + half)*NC + code], so
  // that lanes holding different codes read different 16-byte words
  const int n = kind == KIND_TIP ? C : CC;
  for (int j = threadIdx.x; j < K * n * 4; j += blockDim.x) {
    const int s = j & 3;
    const int code = (j >> 2) % n;
    const int k = (j >> 2) / n;
    const double* row = P + k * 16 + s * 4;
    const double* v = kind == KIND_TIP ? codeP + code * 4 : cvec + ((size_t)k * CC + code) * 4;
    tab[(((k * 2 + (s >> 1)) * NC + code) << 1) + (s & 1)] =
        fma(row[3], v[3], fma(row[2], v[2], fma(row[1], v[1], row[0] * v[0])));
  }
}

__device__ __forceinline__ V4 lds4_planes(const double* tab, int k, int NC, int code) {
  const double2 lo = *reinterpret_cast<const double2*>(tab + (((k * 2 + 0) * NC + code) << 1));
  const double2 hi = *reinterpret_cast<const double2*>(tab + (((k * 2 + 1) * NC + code) << 1));
  return V4{lo.x, lo.y, hi.x, hi.y};
}

// post-order level kernel, cherry-aware (levels >= 2 when fusion is on).
// CHAIN: one launch walks `chainOps` consecutive ops (a run of levels with very few nodes,
// e.g. the top of the tree, or all of a ladder-like tree) for its own pattern slice:
// a pattern only depends on itself, so a thread re-reads what it wrote one op earlier
// (coherent loads, mostly L2 hits) and no grid-wide barrier is needed between levels.
template <int K, bool CHAIN>
__global__ void __launch_bounds__(FWD_THREADS, (K <= 4 ? 4 : 2))
fwd4c_kernel(const NodeOp* __restrict__ ops, int opBegin, const double* __restrict__ mats,
             const uint8_t* __restrict__ tips, const double* __restrict__ codeP,
             double* __restrict__ partials, int16_t* __restrict__ expo, CherryArgs ch, int T,
             int Npad, int C, int B, int ppt, int chainOps) {
  extern __shared__ double sm[];
  const int d = blockIdx.z;
  const int I = T - 1;
  const int NC = ch.CC > C ? ch.CC : C;
  const int tabN = K * (NC > 4 ? NC : 4) * 4;
  double* tabL = sm;
  double* tabR = sm + tabN;
  const double* matsD = mats + (size_t)d * B * K * 16;
  const size_t nodeStride = (size_t)K * Npad * 4;
  double* base = partials + (size_t)d * I * nodeStride;
  const int i0 = blockIdx.x * (FWD_THREADS * ppt) + threadIdx.x;
  const int nOps = CHAIN ? chainOps : 1;
  for (int j = 0; j < nOps; ++j) {
    const NodeOp op = ops[opBegin + (CHAIN ? j : (int)blockIdx.y)];
    int cidxL, cidxR;
    const int kindL = child_kind(op.left, T, ch, cidxL);
    const int kindR = child_kind(op.right, T, ch, cidxR);
    const double* cvL = kindL == KIND_CHERRY ? ch.vec + ((size_t)d * ch.n + cidxL) * K * ch.CC * 4 : nullptr;
    const double* cvR = kindR == KIND_CHERRY ? ch.vec + ((size_t)d * ch.n + cidxR) * K * ch.CC * 4 : nullptr;
    if (CHAIN && j > 0) __syncthreads();   // the previous op's tables are no longer read
    build_child_table_c<K>(tabL, matsD + (size_t)op.left * K * 16, kindL, codeP, C, cvL, ch.CC, NC);
    build_child_table_c<K>(tabR, matsD + (size_t)op.right * K * 16, kindR, codeP, C, cvR, ch.CC, NC);
    __syncthreads();
    if (j == 0) pdl_wait_then_trigger();
    // byte row holding the child's symbol code (tip code, or pair code of a cherry)
    const uint8_t* l0 = nullptr;
    const uint8_t* r0 = nullptr;
    if (kindL == KIND_TIP) l0 = tips + (size_t)op.left * Npad;
    if (kindL == KIND_CHERRY) l0 = ch.code + (size_t)cidxL * Npad;
    if (kindR == KIND_TIP) r0 = tips + (size_t)op.right * Npad;
    if (kindR == KIND_CHERRY) r0 = ch.code + (size_t)cidxR * Npad;

#pragma unroll 1
    for (int it = 0; it < ppt; ++it) {
      const int i = i0 + it * FWD_THREADS;
      if (i >= Npad) break;
      V4 a[K], b[K];
      int codeL = 0, codeR = 0;
      if (kindL == KIND_STORED) {
        const double* p = base + (size_t)(op.left - T) * nodeStride + (size_t)i * 4;
#pragma unroll
        for (int k = 0; k < K; ++k)
          a[k] = CHAIN ? ldg4_coherent(p + (size_t)k * Npad * 4) : ldg4(p + (size_t)k * Npad * 4);
      } else {
        codeL = l0[i];
      }
      if (kindR == KIND_STORED) {
        const double* p = base + (size_t)(op.right - T) * nodeStride + (size_t)i * 4;
#pragma unroll
        for (int k = 0; k < K; ++k)
          b[k] = CHAIN ? ldg4_coherent(p + (size_t)k * Npad * 4) : ldg4(p + (size_t)k * Npad * 4);
      } else {
        codeR = r0[i];
      }
      V4 out[K];
      double m = 0.0;
#pragma unroll
      for (int k = 0; k < K; ++k) {
        const V4 ul = kindL != KIND_STORED ? lds4_planes(tabL, k, NC, codeL) : matvec(tabL + k * 16, a[k]);
        const V4 ur = kindR != KIND_STORED ? lds4_planes(tabR, k, NC, codeR) : matvec(tabR + k * 16, b[k]);
        out[k] = mul4(ul, ur);
        m = fmax(m, max4(out[k]));
      }
      int eb = (__double2hiint(m) >> 20) & 0x7ff;
      eb = eb > 2044 ? 2044 : eb;
      const double f = __hiloint2double((2045 - eb) << 20, 0);
      double* q = base + (size_t)(op.node - T) * nodeStride + (size_t)i * 4;
#pragma unroll
      for (int k = 0; k < K; ++k) stg4(q + (size_t)k * Npad * 4, scale4(out[k], f));
      expo[((size_t)d * I + (op.node - T)) * Npad + i] = (int16_t)(eb - 1022);
    }
  }
}

// fp64 tensor-core instruction: C[8x8] += A[8x4] . B[4x8]
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

// ---------------------------------------------------------------------------
// Pre-order level kernel in DMMA fragment layout: no shared-memory staging and no
// transition-matrix broadcasts in the pattern loop.
//
// A warp step handles 8 patterns; lane = (p = lane>>2 pattern, c = lane&3).  With
// m8n8k4 (A[8x4]: lane holds A[p][c]; B[4x8]: lane holds B[c][p]; C[8x8]: lane
// holds C[p][2c], C[p][2c+1]) the whole node update is four MMAs whose constant
// operands are one double per lane each:
//   U  = p~_l . [P_l^T | 0] + p~_r . [0 | P_r^T]     -> C cols 0-3 = u_l, 4-7 = u_r
//        (the A operands p~[p][c] are one coalesced 256-byte load per child)
//   M  = q^_n o U (lane-local)                        -> cols 0-3 = m_r, 4-7 = m_l
//   O  = M[:, {0,2,4,6}] . W_even + M[:, {1,3,5,7}] . W_odd,  W = diag(P_r, P_l)
//        (the reduction index is permuted so that every lane feeds the two C values
//         it already holds as A operands: no transpose between the products)
//                                                     -> cols 0-3 = q^_r, 4-7 = q^_l
// Lanes c<2 therefore own the right child's outputs and lanes c>=2 the left
// child's.  G = sum_i w m (x) p~ needs the child's whole p~ vector next to the
// lane's two m values: a quad all-gather (three shuffles) and 8 FMAs per lane.
//
// Inputs: the three vectors of a 32-pattern block (q^_n, p~_l, p~_r: 1 KB each,
// contiguous) are brought in by bulk asynchronous copies (cp.async.bulk -> SASS
// UBLKCP) into a per-warp ring of shared-memory slots, completion tracked by one
// mbarrier per slot; weights, scale exponents and tip codes ride along as 16-byte
// cp.async.  Every warp is its own producer and consumer, so there is no CTA-wide
// synchronisation in the pattern loop, and the loads in flight do not occupy
// registers: STAGES-1 blocks (3 KB each) per warp.
constexpr int BWDF_THREADS = 128;

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes,
                                         uint32_t bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
          "r"(dst), "l"(src), "r"(bytes), "r"(bar)
      : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}\n" ::"r"(bar), "r"(parity)
      : "memory");
}

// slot (bytes): q^_n 1024 | p~_l 1024 | p~_r 1024 | w 256 | e_l 64 | e_r 64 | code_l 32 | code_r 32
constexpr int BWDT_SLOT = 440;      // doubles per slot (3520 bytes)
constexpr int BWDT_W = 384;         // double offset of the weights
constexpr int BWDT_EL = 3328, BWDT_ER = 3392, BWDT_CL = 3456, BWDT_CR = 3488;  // byte offsets

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}

// A child is STORED (its vector is bulk-copied), a TIP (vector = codeP[tip code]) or, with
// cherry tabulation on, a CHERRY (vector = cherryVec[pair code]; it still receives q^ and
// has its exponents in `expo`).
// CHAIN: the launch covers a run of levels with very few nodes (see chain_runs): grid
// (chunks, K, draws), and every CTA walks the run's ops from the top level down for its own
// (pattern chunk, category).  What a child reads (q^ written while its parent was processed)
// was written by the same warp of the same CTA, so levels need no grid-wide barrier -- only
// a generic->async proxy fence, because the reads are bulk copies.
template <int STAGES, int MINBLOCKS, bool CHAIN>
__global__ void __launch_bounds__(BWDF_THREADS, MINBLOCKS)
bwd4_tma_kernel(const NodeOp* __restrict__ ops, int opBegin,
                const double* __restrict__ mats, const uint8_t* __restrict__ tips,
                const double* __restrict__ codeP, const double* __restrict__ partials,
                const int16_t* __restrict__ expo, const double* __restrict__ weights,
                double* __restrict__ pre, double* __restrict__ gpart,
                const int* __restrict__ chunkBase, size_t chunkTotal, CherryArgs ch, int T,
                int Npad, int C, int B, int K, int chunkPatterns, int nChunk, int chainOps) {
  extern __shared__ __align__(128) double sm[];
  // sm: slots[warps][STAGES][440] | cp[C][4] | vecL[CC][4] vecR[CC][4] | mbarriers[warps][STAGES]
  // (5 CTAs/SM need <= 45 KB each; the final reduction reuses the head of each warp's ring)
  constexpr int NW = BWDF_THREADS / 32;
  double* slots = sm;
  double* cp = slots + NW * STAGES * BWDT_SLOT;
  double* vecL = cp + C * 4;
  double* vecR = vecL + ch.CC * 4;
  uint64_t* bars = reinterpret_cast<uint64_t*>(vecR + ch.CC * 4);

  const int d = blockIdx.z;
  const int I = T - 1;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int p = lane >> 2, c = lane & 3;
  const bool rside = c < 2;
  const int s0 = (c & 1) * 2;
  if (lane == 0) {
#pragma unroll
    for (int st = 0; st < STAGES; ++st) mbar_init(smem_u32(bars + warp * STAGES + st), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  const int begin = blockIdx.x * chunkPatterns;
  int end = begin + chunkPatterns;
  end = end < Npad ? end : Npad;
  const int first = begin + warp * 32;
  const int nb = first < end ? (end - first + BWDF_THREADS - 1) / BWDF_THREADS : 0;
  double* mySlots = slots + warp * STAGES * BWDT_SLOT;
  const uint32_t myBars = smem_u32(bars + warp * STAGES);
  const size_t nodeStride = (size_t)K * Npad * 4;
  const size_t drawBase = (size_t)d * I * nodeStride;
  const double* matsD = mats + (size_t)d * B * K * 16;
  int ring = 0;   // blocks this warp has pushed through its ring so far (all ops)
  bool cpLoaded = false;

  const int nOps = CHAIN ? chainOps : 1;
  for (int jop = 0; jop < nOps; ++jop) {
    int k;
    NodeOp op;
    if (CHAIN) {
      k = blockIdx.y;
      op = ops[opBegin + chainOps - 1 - jop];   // ops are sorted by level: walk downwards
    } else {
      const int nodeSlot = blockIdx.y / K;
      k = blockIdx.y - nodeSlot * K;
      op = ops[opBegin + nodeSlot];
    }
    int cidxL, cidxR;
    const int kindL = child_kind(op.left, T, ch, cidxL);
    const int kindR = child_kind(op.right, T, ch, cidxR);
    const bool storedL = kindL == KIND_STORED, storedR = kindR == KIND_STORED;
    const bool tipL = kindL == KIND_TIP, tipR = kindR == KIND_TIP;
    const double* gPl = matsD + ((size_t)op.left * K + k) * 16;
    const double* gPr = matsD + ((size_t)op.right * K + k) * 16;
    if ((tipL || tipR) && !cpLoaded) {
      for (int j = threadIdx.x; j < C * 4; j += blockDim.x) cp[j] = codeP[j];
      cpLoaded = true;   // block-uniform
    }
    if (kindL == KIND_CHERRY) {
      const double* src = ch.vec + (((size_t)d * ch.n + cidxL) * K + k) * ch.CC * 4;
      for (int j = threadIdx.x; j < ch.CC * 4; j += blockDim.x) vecL[j] = src[j];
    }
    if (kindR == KIND_CHERRY) {
      const double* src = ch.vec + (((size_t)d * ch.n + cidxR) * K + k) * ch.CC * 4;
      for (int j = threadIdx.x; j < ch.CC * 4; j += blockDim.x) vecR[j] = src[j];
    }
    __syncthreads();

    const double b1 = p < 4 ? gPl[p * 4 + c] : 0.0;
    const double b2 = p >= 4 ? gPr[(p - 4) * 4 + c] : 0.0;
    double ba, bb;
    if (rside) {
      ba = p < 4 ? gPr[(2 * c) * 4 + p] : 0.0;
      bb = p < 4 ? gPr[(2 * c + 1) * 4 + p] : 0.0;
    } else {
      ba = p >= 4 ? gPl[(2 * c - 4) * 4 + (p - 4)] : 0.0;
      bb = p >= 4 ? gPl[(2 * c - 3) * 4 + (p - 4)] : 0.0;
    }

    const size_t kOff = (size_t)k * Npad * 4;
    const double* qn = pre + drawBase + (size_t)(op.node - T) * nodeStride + kOff;
    const double* pl = storedL ? partials + drawBase + (size_t)(op.left - T) * nodeStride + kOff : nullptr;
    const double* prr = storedR ? partials + drawBase + (size_t)(op.right - T) * nodeStride + kOff : nullptr;
    const int mine = rside ? op.right : op.left;
    const bool store = mine >= T;
    double* qc = store ? pre + drawBase + (size_t)(mine - T) * nodeStride + kOff : nullptr;

    // the one 16-byte piece of per-pattern side data this lane copies per block
    const char* sideSrc = nullptr;   // source at pattern 0
    int sideDst = 0, sideScale = 0;  // byte offset in the slot; source bytes per pattern
    const uint8_t* rowL = tipL ? tips + (size_t)op.left * Npad
                                : (storedL ? nullptr : ch.code + (size_t)cidxL * Npad);
    const uint8_t* rowR = tipR ? tips + (size_t)op.right * Npad
                                : (storedR ? nullptr : ch.code + (size_t)cidxR * Npad);
    const double* tabL = tipL ? cp : vecL;   // coded child: vector of a code
    const double* tabR = tipR ? cp : vecR;
    if (lane < 16) {
      sideSrc = reinterpret_cast<const char*>(weights) + lane * 16;
      sideDst = BWDT_W * 8 + lane * 16; sideScale = 8;
    } else if (lane < 20) {
      if (!tipL) sideSrc = reinterpret_cast<const char*>(expo + ((size_t)d * I + (op.left - T)) * Npad) + (lane - 16) * 16;
      sideDst = BWDT_EL + (lane - 16) * 16; sideScale = 2;
    } else if (lane < 24) {
      if (!tipR) sideSrc = reinterpret_cast<const char*>(expo + ((size_t)d * I + (op.right - T)) * Npad) + (lane - 20) * 16;
      sideDst = BWDT_ER + (lane - 20) * 16; sideScale = 2;
    } else if (lane < 26) {
      if (rowL) sideSrc = reinterpret_cast<const char*>(rowL) + (lane - 24) * 16;
      sideDst = BWDT_CL + (lane - 24) * 16; sideScale = 1;
    } else if (lane < 28) {
      if (rowR) sideSrc = reinterpret_cast<const char*>(rowR) + (lane - 26) * 16;
      sideDst = BWDT_CR + (lane - 26) * 16; sideScale = 1;
    }

    double g[2][4];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) g[i][j] = 0.0;

    const uint32_t txBytes = 1024u + (storedL ? 1024u : 0u) + (storedR ? 1024u : 0u);

    // all lanes; exactly one cp.async group is committed per call (possibly empty)
    auto issue = [&](int blk) {
      if (blk < nb) {
        const int st = (ring + blk) % STAGES;
        const int i0 = first + blk * BWDF_THREADS;
        const uint32_t dst = smem_u32(mySlots + st * BWDT_SLOT);
        if (sideSrc) cp_async16(dst + sideDst, sideSrc + (size_t)i0 * sideScale);
        if (lane == 0) {
          const size_t off = (size_t)i0 * 4;
          const uint32_t bar = myBars + st * 8;
          mbar_expect_tx(bar, txBytes);
          bulk_g2s(dst, qn + off, 1024u, bar);
          if (storedL) bulk_g2s(dst + 1024u, pl + off, 1024u, bar);
          if (storedR) bulk_g2s(dst + 2048u, prr + off, 1024u, bar);
        }
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    };
    if (jop == 0) pdl_wait_then_trigger();   // everything above read only ops / mats / codeP / cherry tables
#pragma unroll
    for (int blk = 0; blk < STAGES; ++blk) issue(blk);

    for (int blk = 0; blk < nb; ++blk) {
      const int base = first + blk * BWDF_THREADS;
      const int st = (ring + blk) % STAGES;
      const double* slot = mySlots + st * BWDT_SLOT;
      const char* slotB = reinterpret_cast<const char*>(slot);
      asm volatile("cp.async.wait_group %0;" ::"n"(STAGES - 1) : "memory");
      __syncwarp();
      mbar_wait(myBars + st * 8, (uint32_t)(((ring + blk) / STAGES) & 1));
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const int j = t * 8 + p;
        const int i = base + j;
        const double2 q = *reinterpret_cast<const double2*>(slot + j * 4 + s0);
        const double al = storedL ? slot[128 + t * 32 + lane]
                                  : tabL[(int)reinterpret_cast<const uint8_t*>(slotB + BWDT_CL)[j] * 4 + c];
        const double ar = storedR ? slot[256 + t * 32 + lane]
                                  : tabR[(int)reinterpret_cast<const uint8_t*>(slotB + BWDT_CR)[j] * 4 + c];
        const double w = slot[BWDT_W + j];
        double u0 = 0.0, u1 = 0.0;
        dmma884(u0, u1, al, b1);
        dmma884(u0, u1, ar, b2);
        const double m0 = q.x * u0, m1 = q.y * u1;
        double o0 = 0.0, o1 = 0.0;
        dmma884(o0, o1, m0, ba);
        dmma884(o0, o1, m1, bb);
        if (store) {
          const int ex = reinterpret_cast<const int16_t*>(slotB + (rside ? BWDT_ER : BWDT_EL))[j];
          const double f = __hiloint2double((1023 - ex) << 20, 0);
          *reinterpret_cast<double2*>(qc + (size_t)i * 4 + s0) = make_double2(o0 * f, o1 * f);
        }
        // (q^_n, hence m, is exactly 0 where w == 0: see root4_bwd_kernel)
        const double own = rside ? ar : al;
        const double oth = rside ? al : ar;
        const double wm0 = w * m0, wm1 = w * m1;
        const double x1 = __shfl_xor_sync(0xffffffffu, own, 1);
        const double x2 = __shfl_xor_sync(0xffffffffu, oth, 2);
        const double x3 = __shfl_xor_sync(0xffffffffu, oth, 3);
        g[0][0] = fma(wm0, own, g[0][0]); g[1][0] = fma(wm1, own, g[1][0]);
        g[0][1] = fma(wm0, x1, g[0][1]);  g[1][1] = fma(wm1, x1, g[1][1]);
        g[0][2] = fma(wm0, x2, g[0][2]);  g[1][2] = fma(wm1, x2, g[1][2]);
        g[0][3] = fma(wm0, x3, g[0][3]);  g[1][3] = fma(wm1, x3, g[1][3]);
      }
      __syncwarp();   // every lane has read the slot: it can be refilled
      issue(blk + STAGES);
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    ring += nb;
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        double v = g[i][j];
        v += __shfl_xor_sync(0xffffffffu, v, 4);
        v += __shfl_xor_sync(0xffffffffu, v, 8);
        v += __shfl_xor_sync(0xffffffffu, v, 16);
        if (p == 0) mySlots[(rside ? 16 : 0) + (s0 + i) * 4 + (c ^ j)] = v;
      }
    if (CHAIN) asm volatile("fence.proxy.async;" ::: "memory");   // q^ stores -> later bulk reads
    __syncthreads();
    if (threadIdx.x < 32) {
      const int child = threadIdx.x >> 4;
      double t = 0.0;
#pragma unroll
      for (int w2 = 0; w2 < NW; ++w2) t += slots[w2 * STAGES * BWDT_SLOT + threadIdx.x];
      const int branch = child ? op.right : op.left;
      gpart[((size_t)d * chunkTotal + chunkBase[branch] + (size_t)k * nChunk + blockIdx.x) * 16 +
            (threadIdx.x & 15)] = t;
    }
    if (CHAIN) __syncthreads();   // ring heads and code tables are reused by the next op
  }
}


// ---------------------------------------------------------------------------
// Pre-order kernel for nodes whose two children are tips (level 1), bulk-copy
// pipelined: same arithmetic as bwd4_tips_kernel (lane = pattern, predicated
// accumulation of G in registers), but q^_n arrives through a per-warp ring of
// shared-memory slots filled by cp.async.bulk, weights and tip codes by 16-byte
// cp.async, so that several blocks per warp are in flight without holding registers.
// slot (bytes): q^_n 1024 | w 256 | code_l 32 | code_r 32 | pad -> 1408
constexpr int BTT_SLOT = 1408;
constexpr int BTT_W = 1024, BTT_CL = 1280, BTT_CR = 1312;

template <int STAGES>
__global__ void __launch_bounds__(BWD_THREADS, 2)
bwd4_tips_tma_kernel(const NodeOp* __restrict__ ops, int opBegin,
                     const double* __restrict__ mats, const uint8_t* __restrict__ tips,
                     const double* __restrict__ codeP, const int* __restrict__ codeMask,
                     const double* __restrict__ weights, const double* __restrict__ pre,
                     double* __restrict__ gpart, const int* __restrict__ chunkBase,
                     size_t chunkTotal, int T, int Npad, int C, int B, int K,
                     int chunkPatterns, int nChunk) {
  extern __shared__ __align__(128) double sm[];
  // sm: slots[warps][STAGES][1408 B] | Pl[16] Pr[16] | tabL[C][4] tabR[C][4] | red[warps][32]
  //     | mbarriers[warps][STAGES] | masks[C] (int)
  constexpr int NW = BWD_THREADS / 32;
  char* slots = reinterpret_cast<char*>(sm);
  double* Pl = sm + NW * STAGES * (BTT_SLOT / 8);
  double* Pr = Pl + 16;
  double* tabL = Pr + 16;
  double* tabR = tabL + C * 4;
  double* red = tabR + C * 4;
  uint64_t* bars = reinterpret_cast<uint64_t*>(red + NW * 32);
  int* masks = reinterpret_cast<int*>(bars + NW * STAGES);

  const int nodeSlot = blockIdx.y / K;
  const int k = blockIdx.y - nodeSlot * K;
  const NodeOp op = ops[opBegin + nodeSlot];
  const int d = blockIdx.z;
  const int I = T - 1;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const double* matsD = mats + (size_t)d * B * K * 16;
  const double* gPl = matsD + ((size_t)op.left * K + k) * 16;
  const double* gPr = matsD + ((size_t)op.right * K + k) * 16;
  if (threadIdx.x < 16) Pl[threadIdx.x] = gPl[threadIdx.x];
  else if (threadIdx.x < 32) Pr[threadIdx.x - 16] = gPr[threadIdx.x - 16];
  for (int j = threadIdx.x; j < C; j += blockDim.x) masks[j] = codeMask[j];
  if (lane == 0) {
#pragma unroll
    for (int st = 0; st < STAGES; ++st) mbar_init(smem_u32(bars + warp * STAGES + st), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();

  const size_t nodeStride = (size_t)K * Npad * 4;
  const double* qn = pre + ((size_t)d * I + (op.node - T)) * nodeStride + (size_t)k * Npad * 4;
  const char* sideSrc = nullptr;
  int sideDst = 0, sideScale = 0;
  if (lane < 16) {
    sideSrc = reinterpret_cast<const char*>(weights) + lane * 16;
    sideDst = BTT_W + lane * 16; sideScale = 8;
  } else if (lane < 18) {
    sideSrc = reinterpret_cast<const char*>(tips + (size_t)op.left * Npad) + (lane - 16) * 16;
    sideDst = BTT_CL + (lane - 16) * 16; sideScale = 1;
  } else if (lane < 20) {
    sideSrc = reinterpret_cast<const char*>(tips + (size_t)op.right * Npad) + (lane - 18) * 16;
    sideDst = BTT_CR + (lane - 18) * 16; sideScale = 1;
  }
  const int begin = blockIdx.x * chunkPatterns;
  int end = begin + chunkPatterns;
  end = end < Npad ? end : Npad;
  const int first = begin + warp * 32;
  const int nb = first < end ? (end - first + BWD_THREADS - 1) / BWD_THREADS : 0;
  char* mySlots = slots + warp * STAGES * BTT_SLOT;
  const uint32_t myBars = smem_u32(bars + warp * STAGES);
  auto issue = [&](int blk) {
    if (blk < nb) {
      const int st = blk % STAGES;
      const int i0 = first + blk * BWD_THREADS;
      const uint32_t dst = smem_u32(mySlots + st * BTT_SLOT);
      if (sideSrc) cp_async16(dst + sideDst, sideSrc + (size_t)i0 * sideScale);
      if (lane == 0) {
        const uint32_t bar = myBars + st * 8;
        mbar_expect_tx(bar, 1024u);
        bulk_g2s(dst, qn + (size_t)i0 * 4, 1024u, bar);
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  // transition tables (depend on mats / codeP only), then wait for the level above
  for (int j = threadIdx.x; j < C * 4; j += blockDim.x) {
    const int s = j & 3, code = j >> 2;
    const double* c = codeP + code * 4;
    const double* rl = Pl + s * 4;
    const double* rr = Pr + s * 4;
    tabL[j] = fma(rl[3], c[3], fma(rl[2], c[2], fma(rl[1], c[1], rl[0] * c[0])));
    tabR[j] = fma(rr[3], c[3], fma(rr[2], c[2], fma(rr[1], c[1], rr[0] * c[0])));
  }
  __syncthreads();
  pdl_wait_then_trigger();
#pragma unroll
  for (int blk = 0; blk < STAGES; ++blk) issue(blk);

  double g[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) g[j] = 0.0;
  for (int blk = 0; blk < nb; ++blk) {
    const int st = blk % STAGES;
    const char* slot = mySlots + st * BTT_SLOT;
    asm volatile("cp.async.wait_group %0;" ::"n"(STAGES - 1) : "memory");
    __syncwarp();
    mbar_wait(myBars + st * 8, (uint32_t)((blk / STAGES) & 1));
    const V4 q = lds4(reinterpret_cast<const double*>(slot) + lane * 4);
    const double w = reinterpret_cast<const double*>(slot + BTT_W)[lane];
    const int cl = reinterpret_cast<const uint8_t*>(slot + BTT_CL)[lane];
    const int cr = reinterpret_cast<const uint8_t*>(slot + BTT_CR)[lane];
    __syncwarp();
    issue(blk + STAGES);
    const V4 ul = lds4(tabL + cl * 4);
    const V4 ur = lds4(tabR + cr * 4);
    const int ml_mask = masks[cl], mr_mask = masks[cr];
    const V4 qw = scale4(q, w);
    const V4 a = mul4(qw, ur);  // w m_l
    const V4 b = mul4(qw, ul);  // w m_r
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const double sl = (ml_mask >> c) & 1 ? 1.0 : 0.0;
      const double sr = (mr_mask >> c) & 1 ? 1.0 : 0.0;
      g[0 + c] = fma(a.x, sl, g[0 + c]);
      g[4 + c] = fma(a.y, sl, g[4 + c]);
      g[8 + c] = fma(a.z, sl, g[8 + c]);
      g[12 + c] = fma(a.w, sl, g[12 + c]);
      g[16 + c] = fma(b.x, sr, g[16 + c]);
      g[20 + c] = fma(b.y, sr, g[20 + c]);
      g[24 + c] = fma(b.z, sr, g[24 + c]);
      g[28 + c] = fma(b.w, sr, g[28 + c]);
    }
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  const double mine = warp_transpose_sum(g);
  red[warp * 32 + lane] = mine;
  __syncthreads();
  if (threadIdx.x < 32) {
    double t = 0.0;
#pragma unroll
    for (int w2 = 0; w2 < NW; ++w2) t += red[w2 * 32 + threadIdx.x];
    const int branch = threadIdx.x < 16 ? op.left : op.right;
    gpart[((size_t)d * chunkTotal + chunkBase[branch] + (size_t)k * nChunk + blockIdx.x) * 16 +
          (threadIdx.x & 15)] = t;
  }
}

// patterns per thread of a post-order launch: long-lived CTAs on large levels
// (amortises the shared-memory table build), ~8 CTAs per SM on small ones
int fwd_patterns_per_thread(const Engine& e, int draws, int levelCount) {
  const Dims& m = e.dm;
  const long target = (long)e.smCount * 8;
  long ppt = ((long)m.Npad * levelCount * draws) / ((long)FWD_THREADS * target);
  if (ppt < 1) ppt = 1;
  if (ppt > 64) ppt = 64;
  return (int)ppt;
}

CherryArgs cherry_args(const Engine& e) {
  CherryArgs ch;
  ch.idx = e.cherryIdx;
  ch.info = e.cherryInfo;
  ch.vec = e.cherryVec;
  ch.exps = e.cherryExp;
  ch.code = e.cherryCode;
  ch.n = e.cherryOn ? e.nCherry : 0;
  ch.CC = e.cherryOn ? e.dm.C * e.dm.C : 0;
  return ch;
}

template <int K>
void launch_fwdc(Engine& e, int draws, int opBegin, int count, int ppt, bool pdl, bool chain) {
  const Dims& m = e.dm;
  const CherryArgs ch = cherry_args(e);
  const int NC = ch.CC > m.C ? ch.CC : m.C;
  const size_t smem = 2 * (size_t)K * (NC > 4 ? NC : 4) * 4 * sizeof(double);
  const int per = FWD_THREADS * ppt;
  if (chain) {   // `count` consecutive ops walked by every CTA for its own patterns
    dim3 grid((m.Npad + per - 1) / per, 1, draws);
    launch_level(fwd4c_kernel<K, true>, grid, FWD_THREADS, smem, e.stream, pdl, e.ops, opBegin,
                 e.mats, e.tips, e.codeP, e.partials, e.expo, ch, m.T, m.Npad, m.C, m.B, ppt, count);
    return;
  }
  dim3 grid((m.Npad + per - 1) / per, count, draws);
  launch_level(fwd4c_kernel<K, false>, grid, FWD_THREADS, smem, e.stream, pdl, e.ops, opBegin,
               e.mats, e.tips, e.codeP, e.partials, e.expo, ch, m.T, m.Npad, m.C, m.B, ppt, 0);
}

template <int K>
void launch_fwd(Engine& e, int draws, int opBegin, int count, size_t smem, int ppt,
                bool tipLevel, bool pdl) {
  const Dims& m = e.dm;
  const int per = FWD_THREADS * ppt;
  dim3 grid((m.Npad + per - 1) / per, count, draws);
  const size_t pairBytes = smem + ((size_t)m.C * m.C * K * 4) * sizeof(double) +
                           (size_t)m.C * m.C * sizeof(int);
  if (tipLevel && pairBytes <= 40 * 1024) {
    fwd4_tips_kernel<K><<<grid, FWD_THREADS, pairBytes, e.stream>>>(
        e.ops, opBegin, e.mats, e.tips, e.codeP, e.partials, e.expo, m.T, m.Npad, m.C, m.B, ppt);
    return;
  }
  launch_level(fwd4_kernel<K>, grid, FWD_THREADS, smem, e.stream, pdl, e.ops, opBegin, e.mats,
               e.tips, e.codeP, e.partials, e.expo, m.T, m.Npad, m.C, m.B, ppt);
}

}  // namespace

// cherries = level-1 nodes (both children tips) other than the root, when the
// pair alphabet is small enough for the shared-memory tables
int s4_build_cherries(Engine& e) {
  const Dims& m = e.dm;
  e.cherryOn = false;
  e.nCherry = 0;
  const bool allowed = e.spec4 && !(e.cfg.flags & TTB2_FLAG_NO_CHERRY) && m.C * m.C <= 64 &&
                       m.T > 2 && (m.K <= 6 || m.K == 8);
  if (!allowed) return TTB2_OK;
  std::vector<int> idx(m.I, -1), info;
  const int root = e.hostOps.back().node;
  for (int j = e.levelOff[0]; j < e.levelOff[1]; ++j) {
    const NodeOp& op = e.hostOps[j];
    if (op.node == root || op.left >= m.T || op.right >= m.T) return TTB2_OK;  // unexpected
    idx[op.node - m.T] = e.nCherry++;
    info.push_back(op.left);
    info.push_back(op.right);
    info.push_back(op.node);
  }
  if (e.nCherry == 0) return TTB2_OK;
  const int D = e.cfg.max_draws;
  const int CC = m.C * m.C;
  if (e.cherryIdx) { cudaFree(e.cherryIdx); e.cherryIdx = nullptr; }
  if (e.cherryInfo) { cudaFree(e.cherryInfo); e.cherryInfo = nullptr; }
  if (e.cherryVec) { cudaFree(e.cherryVec); e.cherryVec = nullptr; }
  if (e.cherryExp) { cudaFree(e.cherryExp); e.cherryExp = nullptr; }
  TTB2_CUDA_CHECK(cudaMalloc((void**)&e.cherryIdx, m.I * sizeof(int)));
  TTB2_CUDA_CHECK(cudaMalloc((void**)&e.cherryInfo, info.size() * sizeof(int)));
  TTB2_CUDA_CHECK(cudaMalloc((void**)&e.cherryVec, (size_t)D * e.nCherry * m.K * CC * 4 * sizeof(double)));
  TTB2_CUDA_CHECK(cudaMalloc((void**)&e.cherryExp, (size_t)D * e.nCherry * CC * sizeof(int)));
  TTB2_CUDA_CHECK(cudaMemcpy(e.cherryIdx, idx.data(), m.I * sizeof(int), cudaMemcpyHostToDevice));
  TTB2_CUDA_CHECK(cudaMemcpy(e.cherryInfo, info.data(), info.size() * sizeof(int),
                             cudaMemcpyHostToDevice));
  if (e.cherryCode) { cudaFree(e.cherryCode); e.cherryCode = nullptr; }
  TTB2_CUDA_CHECK(cudaMalloc((void**)&e.cherryCode, (size_t)e.nCherry * m.Npad));
  {
    dim3 grid((m.Npad + 255) / 256, e.nCherry);
    cherry_code_kernel<<<grid, 256, 0, e.stream>>>(e.cherryInfo, e.tips, e.cherryCode, m.C, m.Npad);
    TTB2_CUDA_CHECK(cudaGetLastError());
    TTB2_CUDA_CHECK(cudaStreamSynchronize(e.stream));
  }
  e.cherryOn = true;
  return TTB2_OK;
}

int s4_forward(Engine& e, int draws) {
  const Dims& m = e.dm;
  const size_t smem = 2 * (size_t)child_table_doubles(m.K, m.C) * sizeof(double);
  const int nLevels = (int)e.levelOff.size() - 1;
  if (e.cherryOn) {
    // tabulate the cherries (level 1) instead of computing them per pattern
    dim3 grid(e.nCherry, draws);
    cherry_table_kernel<<<grid, 128, 2 * (size_t)m.K * m.C * 4 * sizeof(double), e.stream>>>(
        e.cherryInfo, e.mats, e.codeP, e.cherryVec, e.cherryExp, e.nCherry, m.C, m.B, m.K);
    ++e.launches;
    dim3 grid2((m.Npad / 16 + 127) / 128, e.nCherry, draws);
    cherry_expo_kernel<<<grid2, 128, 0, e.stream>>>(e.cherryInfo, e.cherryExp, e.cherryCode, e.expo,
                                                    e.nCherry, m.C * m.C, m.T, m.Npad);
    ++e.launches;
  }
  const ChainRuns runs = chain_runs(e);
  for (int l = e.cherryOn ? 1 : 0; l < nLevels; ++l) {
    int opBegin = e.levelOff[l];
    int count = e.levelOff[l + 1] - opBegin;
    const bool chain = l >= 1 && runs.endOfStart[l] >= 0;
    if (chain) count = e.levelOff[runs.endOfStart[l] + 1] - opBegin;   // all ops of the run
    int ppt = fwd_patterns_per_thread(e, draws, count);
    if (chain) {   // one wave: every CTA is resident from the first op to the last
      const long slots = std::max(1L, (long)e.smCount * (m.K <= 4 ? 4 : 2) / draws);
      ppt = (int)((m.Npad + (long)FWD_THREADS * slots - 1) / ((long)FWD_THREADS * slots));
      if (ppt < 1) ppt = 1;
    }
    if (chain || (e.cherryOn && (m.K <= 6 || m.K == 8))) {
      const bool pdlc = l > 1 && pdl_enabled();   // the first level follows the cherry tables
      for (int done = 0; done < count; done += 65535) {
        const int c = chain ? count : ((count - done) < 65535 ? (count - done) : 65535);
        switch (m.K) {
          case 1: launch_fwdc<1>(e, draws, opBegin + done, c, ppt, pdlc, chain); break;
          case 2: launch_fwdc<2>(e, draws, opBegin + done, c, ppt, pdlc, chain); break;
          case 3: launch_fwdc<3>(e, draws, opBegin + done, c, ppt, pdlc, chain); break;
          case 4: launch_fwdc<4>(e, draws, opBegin + done, c, ppt, pdlc, chain); break;
          case 5: launch_fwdc<5>(e, draws, opBegin + done, c, ppt, pdlc, chain); break;
          case 6: launch_fwdc<6>(e, draws, opBegin + done, c, ppt, pdlc, chain); break;
          default: launch_fwdc<8>(e, draws, opBegin + done, c, ppt, pdlc, chain); break;
        }
        ++e.launches;
        if (chain) break;
      }
      if (chain) l = runs.endOfStart[l];
      continue;
    }
    const bool tipLevel = (l == 0);  // level 1: tip-tip nodes
    const bool pdl = l > 0 && pdl_enabled();   // the first level follows pmatrix: ordinary launch
    // grid.y is limited to 65535
    for (int done = 0; done < count; done += 65535) {
      const int c = (count - done) < 65535 ? (count - done) : 65535;
      switch (m.K) {
        case 1: launch_fwd<1>(e, draws, opBegin + done, c, smem, ppt, tipLevel, pdl); break;
        case 2: launch_fwd<2>(e, draws, opBegin + done, c, smem, ppt, tipLevel, pdl); break;
        case 3: launch_fwd<3>(e, draws, opBegin + done, c, smem, ppt, tipLevel, pdl); break;
        case 4: launch_fwd<4>(e, draws, opBegin + done, c, smem, ppt, tipLevel, pdl); break;
        case 5: launch_fwd<5>(e, draws, opBegin + done, c, smem, ppt, tipLevel, pdl); break;
        case 6: launch_fwd<6>(e, draws, opBegin + done, c, smem, ppt, tipLevel, pdl); break;
        case 8: launch_fwd<8>(e, draws, opBegin + done, c, smem, ppt, tipLevel, pdl); break;
        default: {
          dim3 grid((m.Npad + FWD_THREADS - 1) / FWD_THREADS, c, draws);
          fwd4_kernel_anyk<<<grid, FWD_THREADS, smem, e.stream>>>(
              e.ops, opBegin + done, e.mats, e.tips, e.codeP, e.partials, e.expo,
              m.T, m.Npad, m.C, m.B, m.K);
        }
      }
      ++e.launches;
    }
  }
  TTB2_CUDA_CHECK(cudaGetLastError());
  return TTB2_OK;
}

int s4_root(Engine& e, int draws) {
  const Dims& m = e.dm;
  const int rootInode = e.hostOps.back().node - m.T;
  if ((long)m.Npad * draws < 65536) {
    const int nblocks = m.Npad / 32;   // <= redPartCap: (Npad/128) * (K + 4) entries per draw
    dim3 grid(nblocks, draws);
    root4_small_kernel<<<grid, dim3(32, ROOT_SLICES), 0, e.stream>>>(
        e.partials, e.expo, e.freqs, e.freqDraws, e.props, e.propDraws, e.weights,
        e.siteLnl, e.redPart, m.T, m.Npad, m.K, rootInode);
    ++e.launches;
    TTB2_CUDA_CHECK(cudaGetLastError());
    return small_reduce_lnl(e, draws, nblocks);
  }
  const int nblocks = (m.Npad + ROOT_THREADS - 1) / ROOT_THREADS;
  dim3 grid(nblocks, draws);
  root4_kernel<<<grid, ROOT_THREADS, 0, e.stream>>>(
      e.partials, e.expo, e.freqs, e.freqDraws, e.props, e.propDraws, e.weights,
      e.siteLnl, e.redPart, m.T, m.Npad, m.K, rootInode);
  ++e.launches;
  TTB2_CUDA_CHECK(cudaGetLastError());
  return small_reduce_lnl(e, draws, nblocks);
}

int s4_backward(Engine& e, int draws) {
  const Dims& m = e.dm;
  const int rootInode = e.hostOps.back().node - m.T;
  {
    const int nblocks = (m.Npad + ROOT_THREADS - 1) / ROOT_THREADS;
    dim3 grid(nblocks, draws);
    root4_bwd_kernel<<<grid, ROOT_THREADS, 0, e.stream>>>(
        e.partials, e.expo, e.freqs, e.freqDraws, e.props, e.propDraws, e.weights,
        e.pre, e.redPart, m.T, m.Npad, m.K, rootInode);
    ++e.launches;
    TTB2_CUDA_CHECK(cudaGetLastError());
    int rc = small_root_grad_reduce(e, draws, nblocks);
    if (rc) return rc;
  }
  const int nLevels = (int)e.levelOff.size() - 1;
  const int maxNodes = 65535 / m.K;
  const ChainRuns runs = chain_runs(e);
  for (int l = nLevels - 1; l >= 0; --l) {
    const int opBegin = e.levelOff[l];
    const int count = e.levelOff[l + 1] - opBegin;
    const int nChunk = e.levelChunks[l];
    int chunkPatterns = (m.Npad + nChunk - 1) / nChunk;
    chunkPatterns = (chunkPatterns + 31) / 32 * 32;
    const bool pdl = l < nLevels - 1 && pdl_enabled();   // the top level follows root4_bwd
    // a run of sparse levels ending here (we walk downwards) goes out as one chain launch
    int chainBegin = 0, chainCount = 0;
    if (runs.startOfEnd[l] >= 0) {
      chainBegin = e.levelOff[runs.startOfEnd[l]];
      chainCount = e.levelOff[l + 1] - chainBegin;
    }
    for (int done = 0; done < count; done += maxNodes) {
      const int c = (count - done) < maxNodes ? (count - done) : maxNodes;
      dim3 grid(nChunk, c * m.K, draws);
      if (l == 0 && e.codes01) {   // level 1: both children are tips with 0/1 code vectors
        constexpr int ST = 4;
        auto smemTips = [&](int codes) {
          return (size_t)(BWD_THREADS / 32) * ST * BTT_SLOT +
                 (32 + 2 * (size_t)codes * 4 + (BWD_THREADS / 32) * 32) * sizeof(double) +
                 (size_t)(BWD_THREADS / 32) * ST * sizeof(uint64_t) + (size_t)codes * sizeof(int);
        };
        if (!e.smemAttrTips) {   // sized for the largest code table (uint8 codes)
          TTB2_CUDA_CHECK(cudaFuncSetAttribute(bwd4_tips_tma_kernel<ST>,
                                               cudaFuncAttributeMaxDynamicSharedMemorySize,
                                               (int)smemTips(256)));
          e.smemAttrTips = true;
        }
        launch_level(bwd4_tips_tma_kernel<ST>, grid, BWD_THREADS, smemTips(m.C), e.stream, pdl,
                     e.ops, opBegin + done, e.mats, e.tips, e.codeP, e.codeMask, e.weights, e.pre,
                     e.gpart, e.chunkBase, e.chunkTotal, m.T, m.Npad, m.C, m.B, m.K,
                     chunkPatterns, nChunk);
      } else {
        constexpr int ST = 3, MB = 5, NWF = BWDF_THREADS / 32;
        const CherryArgs ch = cherry_args(e);
        auto smemOf = [&](int codes, int pairCodes) {
          return ((size_t)NWF * ST * BWDT_SLOT + (size_t)codes * 4 +
                  2 * (size_t)pairCodes * 4) * sizeof(double) +
                 (size_t)NWF * ST * sizeof(uint64_t);
        };
        if (!e.smemAttrTma) {   // largest code table (uint8 codes); cherries need C * C <= 64
          TTB2_CUDA_CHECK(cudaFuncSetAttribute(bwd4_tma_kernel<ST, MB, false>,
                                               cudaFuncAttributeMaxDynamicSharedMemorySize,
                                               (int)smemOf(256, 64)));
          TTB2_CUDA_CHECK(cudaFuncSetAttribute(bwd4_tma_kernel<ST, MB, true>,
                                               cudaFuncAttributeMaxDynamicSharedMemorySize,
                                               (int)smemOf(256, 64)));
          e.smemAttrTma = true;
        }
        if (chainCount > 0) {   // the whole run of levels in one launch
          dim3 cgrid(nChunk, m.K, draws);
          launch_level(bwd4_tma_kernel<ST, MB, true>, cgrid, BWDF_THREADS, smemOf(m.C, ch.CC),
                       e.stream, pdl, e.ops, chainBegin, e.mats, e.tips, e.codeP, e.partials,
                       e.expo, e.weights, e.pre, e.gpart, e.chunkBase, e.chunkTotal, ch, m.T,
                       m.Npad, m.C, m.B, m.K, chunkPatterns, nChunk, chainCount);
        } else {
          launch_level(bwd4_tma_kernel<ST, MB, false>, grid, BWDF_THREADS, smemOf(m.C, ch.CC),
                       e.stream, pdl, e.ops, opBegin + done, e.mats, e.tips, e.codeP, e.partials,
                       e.expo, e.weights, e.pre, e.gpart, e.chunkBase, e.chunkTotal, ch, m.T,
                       m.Npad, m.C, m.B, m.K, chunkPatterns, nChunk, 0);
        }
      }
      ++e.launches;
      if (chainCount > 0) break;
    }
    if (chainCount > 0) l = runs.startOfEnd[l];   // the run is done; continue below it
  }
  TTB2_CUDA_CHECK(cudaGetLastError());
  return small_gpart_reduce(e, draws);
}

}  // namespace ttb2
