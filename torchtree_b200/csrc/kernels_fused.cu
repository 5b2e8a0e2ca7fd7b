// Fused-traversal kernels for 4-state models (sm_100a).
//
// Site patterns are conditionally independent given the tree, so nothing
// forces a kernel boundary between tree levels: a warp that owns 32 patterns of
// one rate category can walk the *whole* tree by itself.  One persistent launch
// per sweep replaces the per-level launches of kernels_s4.cu:
//
//   post-order sweep  each warp executes the post-order program node by node;
//                     the conditional-likelihood vectors of the nodes whose
//                     parent has not been reached yet live in a small per-warp
//                     shared-memory stack (slots assigned on the host, at most
//                     Strahler-number(tree) <= log2(T)+1 of them), so a child's
//                     vector is never re-read from HBM.  Every node's vector is
//                     still written once to HBM (the pre-order sweep needs it).
//   pre-order sweep   the same walk top-down with the pre-order vectors q^ on
//                     the stack: per node it reads the two children's vectors
//                     from HBM (prefetched one node ahead) and writes nothing
//                     but per-branch gradient scalars.
//
// HBM traffic per (pattern, node, category) drops from 5 vectors (level-
// synchronous) to 2 (one write, one read).
//
// Rescaling is per (pattern, category) chain with exact power-of-two factors,
// which removes every cross-category dependency inside the sweeps; the chains
// are recombined at the root with their exponent sums (mathematically the same
// likelihood as the reference's max-over-(k,s) scaler,
// tree_likelihood.py:186-221).
//
// The pre-order sweep does not form d lnL / d P per branch (that would need a
// cross-pattern reduction of 16 values per branch and category at every node).
// In the eigenbasis P = V exp(L tau) V^-1 it accumulates, per thread,
//   gs[c,k]  = sum_i w_i sum_j a_j b_j l_j e^{l_j tau}      (-> d/d tau: branch
//              lengths and site rates), one warp reduction per branch, and
//   H       += w_i (a (x) b) o Phi(tau_{c,k})                (-> d lnL / d Q),
//              thread-private over the whole walk, reduced once at the end,
// with a = V^T m_c, b = V^-1 p~_c.
#include <algorithm>
#include <climits>

#include "engine.cuh"

namespace ttb2 {

namespace {

constexpr double LN2 = 0.693147180559945309417232121458;

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

__device__ __forceinline__ void stg4(double* p, const V4& v) {
  asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(v.x), "d"(v.y),
               "d"(v.z), "d"(v.w)
               : "memory");
}

__device__ __forceinline__ void prefetch_l1(const void* p) {
  asm volatile("prefetch.global.L1 [%0];" ::"l"(p));
}

// per-warp stack slot: two planes of 32 double2 (conflict-free 128-bit accesses)
__device__ __forceinline__ V4 slot_load(const double2* stack, int slot, int lane) {
  const double2 a = stack[slot * 64 + lane];
  const double2 b = stack[slot * 64 + 32 + lane];
  return V4{a.x, a.y, b.x, b.y};
}

__device__ __forceinline__ void slot_store(double2* stack, int slot, int lane, const V4& v) {
  stack[slot * 64 + lane] = make_double2(v.x, v.y);
  stack[slot * 64 + 32 + lane] = make_double2(v.z, v.w);
}

__device__ __forceinline__ double dot4(const V4& a, const V4& b) {
  return fma(a.w, b.w, fma(a.z, b.z, fma(a.y, b.y, a.x * b.x)));
}

// 4x4 row-major matrix in global memory, read row by row with warp-uniform
// 256-bit loads (L1-resident: every warp of the SM walks the same program);
// nothing of the matrix stays live in registers.
__device__ __forceinline__ V4 gmv(const double* m, const V4& v) {  // M v
  V4 u;
  u.x = dot4(ldg4(m), v);
  u.y = dot4(ldg4(m + 4), v);
  u.z = dot4(ldg4(m + 8), v);
  u.w = dot4(ldg4(m + 12), v);
  return u;
}

__device__ __forceinline__ V4 gmtv(const double* m, const V4& v) {  // M^T v
  const V4 r0 = ldg4(m);
  V4 u{r0.x * v.x, r0.y * v.x, r0.z * v.x, r0.w * v.x};
  const V4 r1 = ldg4(m + 4);
  u.x = fma(r1.x, v.y, u.x); u.y = fma(r1.y, v.y, u.y);
  u.z = fma(r1.z, v.y, u.z); u.w = fma(r1.w, v.y, u.w);
  const V4 r2 = ldg4(m + 8);
  u.x = fma(r2.x, v.z, u.x); u.y = fma(r2.y, v.z, u.y);
  u.z = fma(r2.z, v.z, u.z); u.w = fma(r2.w, v.z, u.w);
  const V4 r3 = ldg4(m + 12);
  u.x = fma(r3.x, v.w, u.x); u.y = fma(r3.y, v.w, u.y);
  u.z = fma(r3.z, v.w, u.z); u.w = fma(r3.w, v.w, u.w);
  return u;
}

__device__ __forceinline__ V4 mul4(const V4& a, const V4& b) {
  return V4{a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w};
}

__device__ __forceinline__ V4 scale4(const V4& a, double f) {
  return V4{a.x * f, a.y * f, a.z * f, a.w * f};
}

__device__ __forceinline__ double pow2i(int e) {  // 2^e, e in [-1022, 1023]
  return __hiloint2double((1023 + e) << 20, 0);
}

// shared-memory 4x4 (warp-uniform reads -> broadcast)
__device__ __forceinline__ V4 smv(const double* m, const V4& v) {  // M v
  V4 u;
  u.x = fma(m[3], v.w, fma(m[2], v.z, fma(m[1], v.y, m[0] * v.x)));
  u.y = fma(m[7], v.w, fma(m[6], v.z, fma(m[5], v.y, m[4] * v.x)));
  u.z = fma(m[11], v.w, fma(m[10], v.z, fma(m[9], v.y, m[8] * v.x)));
  u.w = fma(m[15], v.w, fma(m[14], v.z, fma(m[13], v.y, m[12] * v.x)));
  return u;
}

__device__ __forceinline__ V4 smtv(const double* m, const V4& v) {  // M^T v
  V4 u;
  u.x = fma(m[12], v.w, fma(m[8], v.z, fma(m[4], v.y, m[0] * v.x)));
  u.y = fma(m[13], v.w, fma(m[9], v.z, fma(m[5], v.y, m[1] * v.x)));
  u.z = fma(m[14], v.w, fma(m[10], v.z, fma(m[6], v.y, m[2] * v.x)));
  u.w = fma(m[15], v.w, fma(m[11], v.z, fma(m[7], v.y, m[3] * v.x)));
  return u;
}

__device__ __forceinline__ int next_item(int* counter, int lane) {
  int it = 0;
  if (lane == 0) it = atomicAdd(counter, 1);
  return __shfl_sync(0xffffffffu, it, 0);
}

// ---------------------------------------------------------------------------
// Program-ordered streams.  Everything a warp needs at step j of its walk is
// addressed by running counters, never by a value loaded in the same step, so
// all global loads can be issued (or L1/L2-prefetched) steps ahead:
//   recs[j]           static 16-byte record: children ids, stack slots, output
//                     position (uniform load, L1-resident)
//   stream[d][k][j]   the step's transition matrices (and gradient tables),
//                     gathered per evaluation in program order
//   tips4[g][i]       tip codes packed 4 per word in order of use
//   partials[pos]     node vectors stored in the order the pre-order sweep
//                     reads them
// ---------------------------------------------------------------------------
__device__ __forceinline__ void prefetch_l2(const void* p) {
  asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}

__device__ __forceinline__ FusedRec load_rec(const FusedRec* p) {
  const int4 r = __ldg(reinterpret_cast<const int4*>(p));
  FusedRec out;
  out.left = r.x;
  out.right = r.y;
  out.aux = r.z;
  out.slots = r.w;
  return out;
}

__device__ __forceinline__ int slot_a(int slots) { return (int)(signed char)(slots & 0xff); }
__device__ __forceinline__ int slot_b(int slots) { return (int)(signed char)((slots >> 8) & 0xff); }
__device__ __forceinline__ int slot_c(int slots) { return (int)(signed char)((slots >> 16) & 0xff); }

// sequential reader of the packed tip-code stream of one pattern
struct TipReader {
  const uint32_t* base;  // tips4 + i
  size_t stride;         // Npad
  int groups;            // rows available
  uint32_t cur, next;
  int seq;
  __device__ __forceinline__ void init(const uint32_t* b, size_t st, int g) {
    base = b; stride = st; groups = g; seq = 0;
    cur = base[0];
    next = base[(size_t)(groups > 1 ? 1 : 0) * stride];
  }
  __device__ __forceinline__ int pop() {
    const int code = (cur >> ((seq & 3) * 8)) & 0xff;
    ++seq;
    if ((seq & 3) == 0) {
      cur = next;
      const int g = (seq >> 2) + 1;
      next = base[(size_t)(g < groups ? g : groups - 1) * stride];
      const int gp = g + 6;
      prefetch_l2(base + (size_t)(gp < groups ? gp : groups - 1) * stride);
    }
    return code;
  }
};

constexpr int BWD_WARPS = 12;   // 168 registers per thread: no spills in the pre-order sweep
constexpr int FWD_STRIDE = 32;  // doubles per step of the post-order stream: P_l, P_r
constexpr int BWD_STRIDE = 80;  // P_l, P_r, aux_l[20], aux_r[20], pad

// ---------------------------------------------------------------------------
// post-order sweep.  item = (draw, pattern block of 32, category)
// shared: codeP[C][4] | per-warp stacks (M slots x 1 KB)
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(512, 1)
fused_fwd_kernel(const FusedRec* __restrict__ recs, const double* __restrict__ stream,
                 const uint32_t* __restrict__ tips4, int tipGroups,
                 const double* __restrict__ codeP, double* __restrict__ partials,
                 int16_t* __restrict__ expoK, int* __restrict__ esum,
                 int* __restrict__ counter, int T, int Npad, int C, int K, int nItems, int M) {
  extern __shared__ __align__(32) unsigned char smraw[];
  double* sCode = reinterpret_cast<double*>(smraw);
  const int codeDoubles = (C * 4 + 3) & ~3;
  double2* stacks = reinterpret_cast<double2*>(sCode + codeDoubles);
  for (int j = threadIdx.x; j < C * 4; j += blockDim.x) sCode[j] = codeP[j];
  __syncthreads();

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double2* stack = stacks + (size_t)warp * M * 64;
  const int I = T - 1;
  const int nBlocks = Npad >> 5;
  const size_t nodeStride = (size_t)K * Npad * 4;

  for (int item = next_item(counter, lane); item < nItems; item = next_item(counter, lane)) {
    const int k = item % K;
    const int rest = item / K;
    const int blk = rest % nBlocks;
    const int d = rest / nBlocks;
    const int i = (blk << 5) + lane;
    const double* strm = stream + ((size_t)d * K + k) * I * FWD_STRIDE;
    double* base = partials + (size_t)d * I * nodeStride + ((size_t)k * Npad + i) * 4;
    int16_t* ebase = expoK + (((size_t)d * I) * K + k) * Npad + i;
    int esumLocal = 0;
    TipReader tr;
    tr.init(tips4 + i, (size_t)Npad, tipGroups);
    FusedRec rec = load_rec(recs);
    prefetch_l1(strm);
    prefetch_l1(strm + 16);
    prefetch_l1(strm + FWD_STRIDE);
    prefetch_l1(strm + FWD_STRIDE + 16);
    for (int j = 0; j < I; ++j) {
      const FusedRec nrec = load_rec(recs + (j + 1 < I ? j + 1 : j));
      if (j + 2 < I) {
        prefetch_l1(strm + (size_t)(j + 2) * FWD_STRIDE);
        prefetch_l1(strm + (size_t)(j + 2) * FWD_STRIDE + 16);
      }
      const double* Pl = strm + (size_t)j * FWD_STRIDE;
      V4 vl, vr;
      if (rec.left < T) vl = *reinterpret_cast<const V4*>(sCode + tr.pop() * 4);
      else vl = slot_load(stack, slot_a(rec.slots), lane);
      if (rec.right < T) vr = *reinterpret_cast<const V4*>(sCode + tr.pop() * 4);
      else vr = slot_load(stack, slot_b(rec.slots), lane);
      V4 out = mul4(gmv(Pl, vl), gmv(Pl + 16, vr));
      const double m = fmax(fmax(out.x, out.y), fmax(out.z, out.w));
      const int eb = (__double2hiint(m) >> 20) & 0x7ff;
      // all-zero vector (impossible data under this category): leave it, exponent 0
      const int e = (m > 0.0) ? (eb > 2044 ? 2044 : eb) - 1022 : 0;
      out = scale4(out, pow2i(-e));
      stg4(base + (size_t)rec.aux * nodeStride, out);
      ebase[(size_t)rec.aux * K * Npad] = (int16_t)e;
      esumLocal += e;
      const int so = slot_c(rec.slots);
      if (so >= 0) slot_store(stack, so, lane, out);
      rec = nrec;
    }
    esum[((size_t)d * K + k) * Npad + i] = esumLocal;
    __syncwarp();
  }
}

// ---------------------------------------------------------------------------
// root: recombine the per-category chains.
//   L~_i = sum_k rho_k 2^{E_k - Emax} pi . p~_root[k];  lnL_i = log L~_i + Emax ln 2
// optionally also q^_root[k] = rho_k pi 2^{E_k - Emax} / (2^{e_root,k} L~_i) and the
// per-block partials of d lnL / d rho_k and the root term of d lnL / d pi.
// ---------------------------------------------------------------------------
constexpr int ROOTF_THREADS = 128;
constexpr int MAXK_ROOT = 16;

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

__device__ __forceinline__ double chain_weight(int diff) {  // 2^diff (diff <= 0 normally)
  return diff < -1000 ? 0.0 : pow2i(diff > 1000 ? 1000 : diff);
}

__global__ void __launch_bounds__(ROOTF_THREADS)
fused_root_kernel(const double* __restrict__ partials, const int* __restrict__ esum,
                  const double* __restrict__ freqs, int freqDraws,
                  const double* __restrict__ props, int propDraws,
                  const double* __restrict__ weights, double* __restrict__ siteLnl,
                  double* __restrict__ blockPart, int T, int Npad, int K, int rootInode) {
  __shared__ double red[ROOTF_THREADS / 32];
  const int d = blockIdx.y;
  const int I = T - 1;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const double* fr = freqs + (freqDraws > 1 ? (size_t)d * 4 : 0);
  const double* pr = props + (propDraws > 1 ? (size_t)d * K : 0);
  double contrib = 0.0;
  if (i < Npad) {
    const size_t nodeStride = (size_t)K * Npad * 4;
    const double* p = partials + ((size_t)d * I + rootInode) * nodeStride + (size_t)i * 4;
    const int* es = esum + (size_t)d * K * Npad + i;
    // chains whose root value is exactly zero (e.g. a rate-0 category at a
    // variable pattern) carry no exponent information: leave them out of Emax
    int emax = INT_MIN;
    for (int k = 0; k < K; ++k) {
      const V4 v = ldg4(p + (size_t)k * Npad * 4);
      const double dot = fma(fr[3], v.w, fma(fr[2], v.z, fma(fr[1], v.y, fr[0] * v.x)));
      if (dot > 0.0 && pr[k] > 0.0) emax = max(emax, es[(size_t)k * Npad]);
    }
    double L = 0.0;
    for (int k = 0; k < K; ++k) {
      const V4 v = ldg4(p + (size_t)k * Npad * 4);
      const double dot = fma(fr[3], v.w, fma(fr[2], v.z, fma(fr[1], v.y, fr[0] * v.x)));
      const double cw = (dot > 0.0) ? chain_weight(es[(size_t)k * Npad] - emax) : 0.0;
      L = fma(pr[k] * cw, dot, L);
    }
    const double site = log(L) + (double)emax * LN2;
    siteLnl[(size_t)d * Npad + i] = site;
    const double w = weights[i];
    contrib = (w != 0.0) ? w * site : 0.0;
  }
  const double t = block_sum(contrib, red);
  if (threadIdx.x == 0) blockPart[(size_t)d * gridDim.x + blockIdx.x] = t;
}

__global__ void __launch_bounds__(ROOTF_THREADS)
fused_root_bwd_kernel(const double* __restrict__ partials, const int* __restrict__ esum,
                      const int16_t* __restrict__ expoK, const double* __restrict__ freqs,
                      int freqDraws, const double* __restrict__ props, int propDraws,
                      const double* __restrict__ weights, double* __restrict__ qroot,
                      double* __restrict__ blockPart, int T, int Npad, int K, int rootInode) {
  __shared__ double red[ROOTF_THREADS / 32];
  const int d = blockIdx.y;
  const int I = T - 1;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const double* fr = freqs + (freqDraws > 1 ? (size_t)d * 4 : 0);
  const double* pr = props + (propDraws > 1 ? (size_t)d * K : 0);
  const size_t nodeStride = (size_t)K * Npad * 4;
  const bool live = i < Npad;
  const double* p = partials + ((size_t)d * I + rootInode) * nodeStride + (size_t)(live ? i : 0) * 4;
  const int* es = esum + (size_t)d * K * Npad + (live ? i : 0);
  double w = 0.0, invL = 0.0;
  int emax = INT_MIN;
  if (live) {
    for (int k = 0; k < K; ++k) {
      const V4 v = ldg4(p + (size_t)k * Npad * 4);
      const double dot = fma(fr[3], v.w, fma(fr[2], v.z, fma(fr[1], v.y, fr[0] * v.x)));
      if (dot > 0.0 && pr[k] > 0.0) emax = max(emax, es[(size_t)k * Npad]);
    }
    double L = 0.0;
    for (int k = 0; k < K; ++k) {
      const V4 v = ldg4(p + (size_t)k * Npad * 4);
      const double dot = fma(fr[3], v.w, fma(fr[2], v.z, fma(fr[1], v.y, fr[0] * v.x)));
      const double cw = (dot > 0.0) ? chain_weight(es[(size_t)k * Npad] - emax) : 0.0;
      L = fma(pr[k] * cw, dot, L);
    }
    invL = 1.0 / L;
    w = weights[i];
    double* q = qroot + (size_t)d * nodeStride + (size_t)i * 4;
    for (int k = 0; k < K; ++k) {
      const V4 v = ldg4(p + (size_t)k * Npad * 4);
      const double dot = fma(fr[3], v.w, fma(fr[2], v.z, fma(fr[1], v.y, fr[0] * v.x)));
      // an exactly-zero chain contributes nothing to lnL; its pre-order vector is
      // set to zero (see DESIGN.md "zero chains")
      const double cw = (dot > 0.0) ? chain_weight(es[(size_t)k * Npad] - emax) : 0.0;
      const int e = expoK[(((size_t)d * I + rootInode) * K + k) * Npad + i];
      const double c = pr[k] * cw * invL * pow2i(-e);
      stg4(q + (size_t)k * Npad * 4, V4{c * fr[0], c * fr[1], c * fr[2], c * fr[3]});
    }
  }
  const double wl = (live && w != 0.0) ? w * invL : 0.0;
  double* out = blockPart + ((size_t)d * gridDim.x + blockIdx.x) * (K + 4);
  V4 accF{0.0, 0.0, 0.0, 0.0};
  for (int k = 0; k < K; ++k) {
    V4 v{0.0, 0.0, 0.0, 0.0};
    double cw = 0.0;
    if (wl != 0.0) v = ldg4(p + (size_t)k * Npad * 4);
    const double dot = fma(fr[3], v.w, fma(fr[2], v.z, fma(fr[1], v.y, fr[0] * v.x)));
    if (wl != 0.0 && dot > 0.0) cw = chain_weight(es[(size_t)k * Npad] - emax);
    const double t = block_sum(wl * cw * dot, red);
    if (threadIdx.x == 0) out[k] = t;
    const double c = wl * cw * pr[k];
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
// per (draw, branch, category) tables for the pre-order sweep:
//   aux[.][0..15]  Phi(tau) (row-major, symmetric),  aux[.][16..19] lambda_j e^{lambda_j tau}
// ---------------------------------------------------------------------------
constexpr int AUX = 20;

__global__ void fused_aux_kernel(const double* __restrict__ bl, const double* __restrict__ rates,
                                 int rateDraws, const double* __restrict__ eval, int eigDraws,
                                 double* __restrict__ aux, int B, int K, int draws) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= draws * B * K) return;
  const int k = idx % K;
  const int b = (idx / K) % B;
  const int d = idx / (K * B);
  const double tau = bl[(size_t)d * B + b] * rates[(size_t)(rateDraws > 1 ? d : 0) * K + k];
  const double* lam = eval + (size_t)(eigDraws > 1 ? d : 0) * 4;
  double a[4], ex[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    a[j] = lam[j] * tau;
    ex[j] = exp(a[j]);
  }
  double* out = aux + (size_t)idx * AUX;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      double phi;
      if (i == j) {
        phi = tau * ex[i];
      } else {
        const double hi = a[i] > a[j] ? a[i] : a[j];
        const double x = -fabs(a[i] - a[j]);
        const double ratio = (x > -1e-9) ? 1.0 + 0.5 * x : expm1(x) / x;
        phi = tau * exp(hi) * ratio;
      }
      out[i * 4 + j] = phi;
    }
    out[16 + i] = lam[i] * ex[i];
  }
}

// ---------------------------------------------------------------------------
// pre-order sweep.  shared: codeP[C][4] | V[16] Vi[16] per draw slot | stacks
// ---------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ double warp_transpose_sum16(double (&v)[16]) {
  // lanes 0..15 end up with the warp totals of v[0..15] (lane L: v[L & 15])
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int j = 0; j < 16; ++j) v[j] += __shfl_xor_sync(0xffffffffu, v[j], 16);
#pragma unroll
  for (int half = 8; half >= 1; half >>= 1) {
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

template <bool NEEDQ>
__global__ void __launch_bounds__(BWD_WARPS * 32, 1)
fused_bwd_kernel(const FusedRec* __restrict__ recs, const double* __restrict__ stream,
                 const double* __restrict__ evec, const double* __restrict__ ivec, int eigDraws,
                 const uint32_t* __restrict__ tips4, int tipGroups,
                 const double* __restrict__ codeP, const double* __restrict__ partials,
                 const int16_t* __restrict__ expoK, const double* __restrict__ qroot,
                 const double* __restrict__ weights, double* __restrict__ gspart,
                 double* __restrict__ hpart, int* __restrict__ counter, int T, int Npad, int C,
                 int B, int K, int nItems, int M) {
  extern __shared__ __align__(32) unsigned char smraw[];
  double* sCode = reinterpret_cast<double*>(smraw);
  const int codeDoubles = (C * 4 + 3) & ~3;
  const int nWarps = blockDim.x >> 5;
  double* sV = sCode + codeDoubles;  // per warp: V[16], Vi[16]
  double2* stacks = reinterpret_cast<double2*>(sV + nWarps * 32);
  for (int j = threadIdx.x; j < C * 4; j += blockDim.x) sCode[j] = codeP[j];
  __syncthreads();

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double2* stack = stacks + (size_t)warp * M * 64;
  double* Vw = sV + warp * 32;
  double* Viw = Vw + 16;
  const int I = T - 1;
  const int lastPos = I - 1;
  const int nBlocks = Npad >> 5;
  const size_t nodeStride = (size_t)K * Npad * 4;
  const size_t expoStride = (size_t)K * Npad;
  int vDraw = -1;

  for (int item = next_item(counter, lane); item < nItems; item = next_item(counter, lane)) {
    const int k = item % K;
    const int rest = item / K;
    const int blk = rest % nBlocks;
    const int d = rest / nBlocks;
    const int i = (blk << 5) + lane;
    const int de = eigDraws > 1 ? d : 0;
    if (NEEDQ && de != vDraw) {
      __syncwarp();
      if (lane < 16) Vw[lane] = evec[(size_t)de * 16 + lane];
      else Viw[lane - 16] = ivec[(size_t)de * 16 + lane - 16];
      vDraw = de;
      __syncwarp();
    }
    const double* strm = stream + ((size_t)d * K + k) * I * BWD_STRIDE;
    const double* pbase = partials + (size_t)d * I * nodeStride + ((size_t)k * Npad + i) * 4;
    const int16_t* ebase = expoK + (((size_t)d * I) * K + k) * Npad + i;
    const double w = weights[i];
    double H[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) H[j] = 0.0;

    // two-deep FIFO over the node vectors in read order (positions 0, 1, ...)
    int rseq = 0;
    V4 f0 = ldg4(pbase);
    int e0 = ebase[0];
    V4 f1 = ldg4(pbase + (size_t)(lastPos > 0 ? 1 : 0) * nodeStride);
    int e1 = ebase[(size_t)(lastPos > 0 ? 1 : 0) * expoStride];
    TipReader tr;
    tr.init(tips4 + i, (size_t)Npad, tipGroups);
    FusedRec rec = load_rec(recs);
    slot_store(stack, slot_a(rec.slots), lane,
               ldg4(qroot + (size_t)d * nodeStride + ((size_t)k * Npad + i) * 4));
#pragma unroll
    for (int c = 0; c < 2 * BWD_STRIDE; c += 16) prefetch_l1(strm + c);

    for (int j = 0; j < I; ++j) {
      const FusedRec nrec = load_rec(recs + (j + 1 < I ? j + 1 : j));
      if (j + 2 < I) {
#pragma unroll
        for (int c = 0; c < BWD_STRIDE; c += 16)
          prefetch_l1(strm + (size_t)(j + 2) * BWD_STRIDE + c);
      }
      const double* st = strm + (size_t)j * BWD_STRIDE;
      V4 vl, vr;
      int el = 0, er = 0;
      const bool tipL = rec.left < T, tipR = rec.right < T;
      if (tipL) {
        vl = *reinterpret_cast<const V4*>(sCode + tr.pop() * 4);
      } else {
        vl = f0; el = e0;
        f0 = f1; e0 = e1;
        ++rseq;
        const int np = rseq + 1 < lastPos ? rseq + 1 : lastPos;
        f1 = ldg4(pbase + (size_t)np * nodeStride);
        e1 = ebase[(size_t)np * expoStride];
        const int pp = rseq + 6 < lastPos ? rseq + 6 : lastPos;
        prefetch_l2(pbase + (size_t)pp * nodeStride);
      }
      if (tipR) {
        vr = *reinterpret_cast<const V4*>(sCode + tr.pop() * 4);
      } else {
        vr = f0; er = e0;
        f0 = f1; e0 = e1;
        ++rseq;
        const int np = rseq + 1 < lastPos ? rseq + 1 : lastPos;
        f1 = ldg4(pbase + (size_t)np * nodeStride);
        e1 = ebase[(size_t)np * expoStride];
        const int pp = rseq + 6 < lastPos ? rseq + 6 : lastPos;
        prefetch_l2(pbase + (size_t)pp * nodeStride);
      }
      const V4 q = slot_load(stack, slot_a(rec.slots), lane);
      const V4 ul = gmv(st, vl);
      const V4 ur = gmv(st + 16, vr);
      const V4 ml = mul4(q, ur);
      const V4 mr = mul4(q, ul);
      if (!tipL) slot_store(stack, slot_b(rec.slots), lane, scale4(gmtv(st, ml), pow2i(-el)));
      if (!tipR) slot_store(stack, slot_c(rec.slots), lane, scale4(gmtv(st + 16, mr), pow2i(-er)));
      double gl, gr;
      if (NEEDQ) {
        // eigenbasis: a = V^T m, b = V^-1 p~
#pragma unroll
        for (int side = 0; side < 2; ++side) {
          const V4 a = smtv(Vw, side ? mr : ml);
          const V4 b = smv(Viw, side ? vr : vl);
          const double* ax = st + 32 + side * 20;
          const V4 le = ldg4(ax + 16);
          const V4 ab = mul4(a, b);
          const double g = dot4(ab, le);
          if (side) gr = g; else gl = g;
          const V4 wa = scale4(a, w);
          const V4 p0 = ldg4(ax), p1 = ldg4(ax + 4), p2 = ldg4(ax + 8), p3 = ldg4(ax + 12);
          H[0] = fma(wa.x, b.x * p0.x, H[0]);   H[1] = fma(wa.x, b.y * p0.y, H[1]);
          H[2] = fma(wa.x, b.z * p0.z, H[2]);   H[3] = fma(wa.x, b.w * p0.w, H[3]);
          H[4] = fma(wa.y, b.x * p1.x, H[4]);   H[5] = fma(wa.y, b.y * p1.y, H[5]);
          H[6] = fma(wa.y, b.z * p1.z, H[6]);   H[7] = fma(wa.y, b.w * p1.w, H[7]);
          H[8] = fma(wa.z, b.x * p2.x, H[8]);   H[9] = fma(wa.z, b.y * p2.y, H[9]);
          H[10] = fma(wa.z, b.z * p2.z, H[10]); H[11] = fma(wa.z, b.w * p2.w, H[11]);
          H[12] = fma(wa.w, b.x * p3.x, H[12]); H[13] = fma(wa.w, b.y * p3.y, H[13]);
          H[14] = fma(wa.w, b.z * p3.z, H[14]); H[15] = fma(wa.w, b.w * p3.w, H[15]);
        }
      } else {
        // d/d tau only: m^T (Q P) p~ with (Q P) supplied in aux[0..15]
        gl = dot4(ml, gmv(st + 32, vl));
        gr = dot4(mr, gmv(st + 52, vr));
      }
      gl = warp_sum(w * gl);
      gr = warp_sum(w * gr);
      if (lane == 0) {
        // gspart [d][branch][k][block]
        gspart[(((size_t)d * B + rec.left) * K + k) * nBlocks + blk] = gl;
        gspart[(((size_t)d * B + rec.right) * K + k) * nBlocks + blk] = gr;
      }
      rec = nrec;
    }
    if (NEEDQ) {
      const double t = warp_transpose_sum16(H);
      // hpart [d][block*K + k][16]
      if (lane < 16) hpart[(((size_t)d * nBlocks + blk) * K + k) * 16 + lane] = t;
    }
    __syncwarp();
  }
}

// gather kernels: program-ordered streams from mats [d][b][k][16] and aux [d][b][k][20]
__global__ void fused_stream_fwd_kernel(const FusedRec* __restrict__ recs,
                                        const double* __restrict__ mats,
                                        double* __restrict__ stream, int I, int B, int K,
                                        int draws) {
  // one thread per (d, k, j, side, row): copies 4 doubles
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t total = (size_t)draws * K * I * 8;
  if (idx >= total) return;
  const int row = idx & 3;
  const int side = (idx >> 2) & 1;
  const size_t r = idx >> 3;
  const int j = (int)(r % I);
  const int k = (int)((r / I) % K);
  const int d = (int)(r / ((size_t)I * K));
  const int child = side ? recs[j].right : recs[j].left;
  const double* src = mats + (((size_t)d * B + child) * K + k) * 16 + row * 4;
  double* dst = stream + (((size_t)d * K + k) * I + j) * FWD_STRIDE + side * 16 + row * 4;
  stg4(dst, ldg4(src));
}

__global__ void fused_stream_bwd_kernel(const FusedRec* __restrict__ recs,
                                        const double* __restrict__ mats,
                                        const double* __restrict__ aux,
                                        double* __restrict__ stream, int I, int B, int K,
                                        int draws) {
  // one thread per (d, k, j, side, part): parts 0..3 = P rows, 4..8 = aux rows
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t total = (size_t)draws * K * I * 2 * 9;
  if (idx >= total) return;
  const int part = (int)(idx % 9);
  const int side = (int)((idx / 9) & 1);
  const size_t r = idx / 18;
  const int j = (int)(r % I);
  const int k = (int)((r / I) % K);
  const int d = (int)(r / ((size_t)I * K));
  const int child = side ? recs[j].right : recs[j].left;
  double* base = stream + (((size_t)d * K + k) * I + j) * BWD_STRIDE;
  if (part < 4) {
    const double* src = mats + (((size_t)d * B + child) * K + k) * 16 + part * 4;
    stg4(base + side * 16 + part * 4, ldg4(src));
  } else {
    const double* src = aux + (((size_t)d * B + child) * K + k) * AUX + (part - 4) * 4;
    stg4(base + 32 + side * 20 + (part - 4) * 4, ldg4(src));
  }
}

// tips4[g][i] = codes of the 4g..4g+3-th tips in order of use, packed little-endian
__global__ void fused_pack_tips_kernel(const uint8_t* __restrict__ tips,
                                       const int* __restrict__ order, int nTips,
                                       uint32_t* __restrict__ tips4, int groups, int Npad) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)groups * Npad) return;
  const int i = (int)(idx % Npad);
  const int g = (int)(idx / Npad);
  uint32_t word = 0;
  for (int c = 0; c < 4; ++c) {
    const int s = g * 4 + c;
    if (s < nTips) word |= (uint32_t)tips[(size_t)order[s] * Npad + i] << (8 * c);
  }
  tips4[idx] = word;
}

// (Q P)[d][b][k] for the NEEDQ=false variant, written into aux[0..15]
__global__ void fused_dp_kernel(const double* __restrict__ bl, const double* __restrict__ rates,
                                int rateDraws, const double* __restrict__ evec,
                                const double* __restrict__ ivec, const double* __restrict__ eval,
                                int eigDraws, double* __restrict__ aux, int B, int K, int draws) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= draws * B * K) return;
  const int k = idx % K;
  const int b = (idx / K) % B;
  const int d = idx / (K * B);
  const int de = eigDraws > 1 ? d : 0;
  const double tau = bl[(size_t)d * B + b] * rates[(size_t)(rateDraws > 1 ? d : 0) * K + k];
  const double* lam = eval + (size_t)de * 4;
  const double* V = evec + (size_t)de * 16;
  const double* Vi = ivec + (size_t)de * 16;
  double le[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) le[j] = lam[j] * exp(lam[j] * tau);
  double* out = aux + (size_t)idx * AUX;
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      double acc = 0.0;
#pragma unroll
      for (int m = 0; m < 4; ++m) acc = fma(V[i * 4 + m] * le[m], Vi[m * 4 + j], acc);
      out[i * 4 + j] = acc;
    }
}

// gscal[d][b][k] = sum_block gspart[d][b][k][block]
__global__ void gs_reduce_kernel(const double* __restrict__ gspart, double* __restrict__ gscal,
                                 size_t items, int nBlocks) {
  // one warp per (d, b, k)
  const size_t wid = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (wid >= items) return;
  const double* p = gspart + wid * nBlocks;
  double acc = 0.0;
  for (int j = lane; j < nBlocks; j += 32) acc += p[j];
  acc = warp_sum(acc);
  if (lane == 0) gscal[wid] = acc;
}

}  // namespace

// ---------------------------------------------------------------------------
// host: traversal programs with stack-slot assignment
// ---------------------------------------------------------------------------
static inline int pack_slots(int a, int b, int c) {
  return (a & 0xff) | ((b & 0xff) << 8) | ((c & 0xff) << 16);
}

int fused_build_programs(Engine& e) {
  const Dims& m = e.dm;
  const int T = m.T, I = m.I;
  std::vector<int> left(2 * T - 1, -1), right(2 * T - 1, -1);
  for (const NodeOp& op : e.hostOps) {
    left[op.node] = op.left;
    right[op.node] = op.right;
  }
  const int root = e.hostOps.back().node;
  // hostOps are level-sorted: children always precede parents
  std::vector<int> needPost(2 * T - 1, 0), needPre(2 * T - 1, 0);
  for (const NodeOp& op : e.hostOps) {
    const int a = std::max(needPost[op.left], needPost[op.right]);
    const int b = std::min(needPost[op.left], needPost[op.right]);
    const bool li = op.left >= T, ri = op.right >= T;
    needPost[op.node] = (li && ri) ? std::max(a, b + 1) : std::max(a, 1);
    const int pa = std::max(needPre[op.left], needPre[op.right]);
    const int pb = std::min(needPre[op.left], needPre[op.right]);
    needPre[op.node] = (li && ri) ? (pa > pb ? pa : pa + 1) : std::max(pa, 1);
  }
  e.fusedSlots = std::max(needPost[root], needPre[root]);
  if (e.fusedSlots > 100) {
    e.fusedSlots = 0;  // would not fit the packed slot bytes; use the per-level kernels
    return TTB2_OK;
  }

  // ---- pre-order program (also fixes the storage order of the node vectors) ----
  std::vector<FusedRec> bwd;
  bwd.reserve(I);
  std::vector<int> pos(2 * T - 1, -1);
  std::vector<int> tipOrderB;
  tipOrderB.reserve(T);
  {
    int nextPos = 0;
    std::vector<int> freeSlots;
    for (int s = e.fusedSlots - 1; s >= 1; --s) freeSlots.push_back(s);
    std::vector<std::pair<int, int>> st;  // (node, slot of its q^)
    st.push_back({root, 0});
    while (!st.empty()) {
      auto [n, sn] = st.back();
      st.pop_back();
      const int l = left[n], r = right[n];
      const bool li = l >= T, ri = r >= T;
      int sl = -1, sr = -1;
      if (li) pos[l] = nextPos++; else tipOrderB.push_back(l);
      if (ri) pos[r] = nextPos++; else tipOrderB.push_back(r);
      if (li && ri) {
        sl = sn;
        if (freeSlots.empty()) {
          set_error("fused program: slot allocation underflow (pre-order)");
          return TTB2_E_INVALID;
        }
        sr = freeSlots.back();
        freeSlots.pop_back();
        // the subtree with the smaller requirement is walked first
        if (needPre[l] <= needPre[r]) {
          st.push_back({r, sr});
          st.push_back({l, sl});
        } else {
          st.push_back({l, sl});
          st.push_back({r, sr});
        }
      } else if (li) {
        sl = sn;
        st.push_back({l, sl});
      } else if (ri) {
        sr = sn;
        st.push_back({r, sr});
      } else {
        freeSlots.push_back(sn);
      }
      bwd.push_back(FusedRec{l, r, n, pack_slots(sn, sl, sr)});
    }
    pos[root] = I - 1;
    if (nextPos != I - 1) {
      set_error("fused program: pre-order walk did not number every internal node");
      return TTB2_E_INVALID;
    }
  }

  // ---- post-order program ----
  std::vector<FusedRec> fwd;
  fwd.reserve(I);
  std::vector<int> tipOrderF;
  tipOrderF.reserve(T);
  {
    std::vector<int> freeSlots;
    for (int s = e.fusedSlots - 1; s >= 0; --s) freeSlots.push_back(s);
    std::vector<int> slotOf(2 * T - 1, -1);
    std::vector<std::pair<int, int>> st;  // state 0 = expand, 1 = emit
    st.push_back({root, 0});
    while (!st.empty()) {
      auto [n, state] = st.back();
      st.pop_back();
      if (n < T) continue;
      if (state == 0) {
        st.push_back({n, 1});
        int first = left[n], second = right[n];
        if (needPost[second] > needPost[first]) std::swap(first, second);
        st.push_back({second, 0});  // pushed in reverse: `first` is visited first
        st.push_back({first, 0});
      } else {
        const int l = left[n], r = right[n];
        const int sl = l >= T ? slotOf[l] : -1;
        const int sr = r >= T ? slotOf[r] : -1;
        if (l < T) tipOrderF.push_back(l);
        if (r < T) tipOrderF.push_back(r);
        int so;
        if (n == root) {
          so = -1;
          if (sl >= 0) freeSlots.push_back(sl);
          if (sr >= 0) freeSlots.push_back(sr);
        } else if (sl >= 0) {
          so = sl;
          if (sr >= 0) freeSlots.push_back(sr);
        } else if (sr >= 0) {
          so = sr;
        } else {
          if (freeSlots.empty()) {
            set_error("fused program: slot allocation underflow (post-order)");
            return TTB2_E_INVALID;
          }
          so = freeSlots.back();
          freeSlots.pop_back();
        }
        slotOf[n] = so;
        fwd.push_back(FusedRec{l, r, pos[n], pack_slots(sl, sr, so)});
      }
    }
  }
  if ((int)fwd.size() != I || (int)bwd.size() != I || (int)tipOrderF.size() != T ||
      (int)tipOrderB.size() != T) {
    set_error("fused program: traversal did not visit every node");
    return TTB2_E_INVALID;
  }
  e.hostFwdProg = fwd;
  e.hostBwdProg = bwd;
  e.hostTipOrderF = tipOrderF;
  e.hostTipOrderB = tipOrderB;
  return TTB2_OK;
}

namespace {

int fused_warps(const Engine& e, size_t extraBytes) {
  // per-warp stack: fusedSlots KB; keep the CTA under ~216 KB of shared memory
  if (e.fusedSlots <= 0) return 0;
  const size_t budget = 216 * 1024 - extraBytes;
  int warps = (int)(budget / ((size_t)e.fusedSlots * 1024 + 256));
  if (warps > 16) warps = 16;
  return warps;
}

size_t code_bytes(const Dims& m) { return (size_t)((m.C * 4 + 3) & ~3) * sizeof(double); }

}  // namespace

bool fused_supported(const Engine& e) {
  return e.spec4 && e.fusedSlots > 0 && fused_warps(e, 16384) >= 4 && e.dm.K <= MAXK_ROOT;
}

int fused_tip_groups(const Engine& e) { return (e.dm.T + 3) / 4 + 1; }

// tips4 streams (device) from the tip order arrays; called after the programs change
int fused_pack_tips(Engine& e) {
  const Dims& m = e.dm;
  const int groups = fused_tip_groups(e);
  const size_t total = (size_t)groups * m.Npad;
  const unsigned blocks = (unsigned)((total + 255) / 256);
  for (int which = 0; which < 2; ++which) {
    const std::vector<int>& order = which ? e.hostTipOrderB : e.hostTipOrderF;
    TTB2_CUDA_CHECK(cudaMemcpy(e.tipOrder, order.data(), m.T * sizeof(int),
                               cudaMemcpyHostToDevice));
    fused_pack_tips_kernel<<<blocks, 256, 0, e.stream>>>(e.tips, e.tipOrder, m.T,
                                                        which ? e.tipsB4 : e.tipsF4, groups,
                                                        m.Npad);
    ++e.launches;
    TTB2_CUDA_CHECK(cudaGetLastError());
    TTB2_CUDA_CHECK(cudaStreamSynchronize(e.stream));
  }
  return TTB2_OK;
}

int fused_forward(Engine& e, int draws) {
  const Dims& m = e.dm;
  {
    const size_t total = (size_t)draws * m.K * m.I * 8;
    fused_stream_fwd_kernel<<<(unsigned)((total + 255) / 256), 256, 0, e.stream>>>(
        e.fwdProg, e.mats, e.streamF, m.I, m.B, m.K, draws);
    ++e.launches;
  }
  const int nItems = (m.Npad / 32) * m.K * draws;
  const size_t codeBytes = code_bytes(m);
  const int warps = fused_warps(e, codeBytes);
  const size_t smem = codeBytes + (size_t)warps * e.fusedSlots * 1024;
  TTB2_CUDA_CHECK(cudaFuncSetAttribute(fused_fwd_kernel,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  TTB2_CUDA_CHECK(cudaMemsetAsync(e.fusedCounter, 0, sizeof(int), e.stream));
  int grid = (nItems + warps - 1) / warps;
  if (grid > e.smCount) grid = e.smCount;
  fused_fwd_kernel<<<grid, warps * 32, smem, e.stream>>>(
      e.fwdProg, e.streamF, e.tipsF4, fused_tip_groups(e), e.codeP, e.partials, e.expoK, e.esum,
      e.fusedCounter, m.T, m.Npad, m.C, m.K, nItems, e.fusedSlots);
  ++e.launches;
  TTB2_CUDA_CHECK(cudaGetLastError());
  return TTB2_OK;
}

int fused_root(Engine& e, int draws) {
  const Dims& m = e.dm;
  const int nblocks = (m.Npad + ROOTF_THREADS - 1) / ROOTF_THREADS;
  dim3 grid(nblocks, draws);
  const int rootPos = m.I - 1;
  fused_root_kernel<<<grid, ROOTF_THREADS, 0, e.stream>>>(
      e.partials, e.esum, e.freqs, e.freqDraws, e.props, e.propDraws, e.weights, e.siteLnl,
      e.redPart, m.T, m.Npad, m.K, rootPos);
  ++e.launches;
  TTB2_CUDA_CHECK(cudaGetLastError());
  return small_reduce_lnl(e, draws, nblocks);
}

size_t fused_gspart_doubles(const Engine& e, int draws) {
  const Dims& m = e.dm;
  return (size_t)draws * m.B * m.K * (m.Npad / 32);
}

int fused_backward(Engine& e, int draws, bool needQ) {
  const Dims& m = e.dm;
  const int rootPos = m.I - 1;
  const int nBlocks = m.Npad / 32;
  {
    const int nblocks = (m.Npad + ROOTF_THREADS - 1) / ROOTF_THREADS;
    dim3 grid(nblocks, draws);
    fused_root_bwd_kernel<<<grid, ROOTF_THREADS, 0, e.stream>>>(
        e.partials, e.esum, e.expoK, e.freqs, e.freqDraws, e.props, e.propDraws, e.weights,
        e.qroot, e.redPart, m.T, m.Npad, m.K, rootPos);
    ++e.launches;
    TTB2_CUDA_CHECK(cudaGetLastError());
    int rc = small_root_grad_reduce(e, draws, nblocks);
    if (rc) return rc;
  }
  {
    const int n = draws * m.B * m.K;
    if (needQ)
      fused_aux_kernel<<<(n + 127) / 128, 128, 0, e.stream>>>(
          e.bl, e.rates, e.rateDraws, e.eval, e.eigDraws, e.aux, m.B, m.K, draws);
    else
      fused_dp_kernel<<<(n + 127) / 128, 128, 0, e.stream>>>(
          e.bl, e.rates, e.rateDraws, e.evec, e.ivec, e.eval, e.eigDraws, e.aux, m.B, m.K, draws);
    ++e.launches;
    const size_t total = (size_t)draws * m.K * m.I * 18;
    fused_stream_bwd_kernel<<<(unsigned)((total + 255) / 256), 256, 0, e.stream>>>(
        e.bwdProg, e.mats, e.aux, e.streamB, m.I, m.B, m.K, draws);
    ++e.launches;
  }
  const int nItems = nBlocks * m.K * draws;
  const size_t codeBytes = code_bytes(m);
  int warps = fused_warps(e, codeBytes + 16 * 32 * sizeof(double));
  if (warps > BWD_WARPS) warps = BWD_WARPS;
  const size_t smem = codeBytes + (size_t)warps * 32 * sizeof(double) +
                      (size_t)warps * e.fusedSlots * 1024;
  TTB2_CUDA_CHECK(cudaMemsetAsync(e.fusedCounter, 0, sizeof(int), e.stream));
  int grid = (nItems + warps - 1) / warps;
  if (grid > e.smCount) grid = e.smCount;
  if (needQ) {
    TTB2_CUDA_CHECK(cudaFuncSetAttribute(fused_bwd_kernel<true>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    fused_bwd_kernel<true><<<grid, warps * 32, smem, e.stream>>>(
        e.bwdProg, e.streamB, e.evec, e.ivec, e.eigDraws, e.tipsB4, fused_tip_groups(e), e.codeP,
        e.partials, e.expoK, e.qroot, e.weights, e.gpart, e.hpart, e.fusedCounter, m.T, m.Npad,
        m.C, m.B, m.K, nItems, e.fusedSlots);
  } else {
    TTB2_CUDA_CHECK(cudaFuncSetAttribute(fused_bwd_kernel<false>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    fused_bwd_kernel<false><<<grid, warps * 32, smem, e.stream>>>(
        e.bwdProg, e.streamB, e.evec, e.ivec, e.eigDraws, e.tipsB4, fused_tip_groups(e), e.codeP,
        e.partials, e.expoK, e.qroot, e.weights, e.gpart, e.hpart, e.fusedCounter, m.T, m.Npad,
        m.C, m.B, m.K, nItems, e.fusedSlots);
  }
  ++e.launches;
  TTB2_CUDA_CHECK(cudaGetLastError());
  {
    const size_t items = (size_t)draws * m.B * m.K;
    const int threads = 256;
    const unsigned blocks = (unsigned)((items * 32 + threads - 1) / threads);
    gs_reduce_kernel<<<blocks, threads, 0, e.stream>>>(e.gpart, e.gscal, items, nBlocks);
    ++e.launches;
  }
  TTB2_CUDA_CHECK(cudaGetLastError());
  return TTB2_OK;
}

}  // namespace ttb2
