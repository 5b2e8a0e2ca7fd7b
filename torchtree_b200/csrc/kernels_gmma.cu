// Tensor-core (fp64 DMMA) level-synchronous peeling kernels for S >= 8 states:
// 20-state amino-acid and 61-state codon models (any 8 <= S <= 64).
//
// Per (node, category, tile of 32 patterns) the work is three small GEMMs:
//   post-order  U_c = P_c [S x S] . V_c [S x 32]            (both children)
//   pre-order   U_c as above,  Q^_c = P_c^T . M_c [S x 32],
//               G_c += (w o M_c) [S x 32] . V_c^T [32 x S]  (= d lnL / d P_c)
// all issued as mma.sync.m8n8k4.f64 (SASS DMMA.8x8x4) with fragments read from
// shared memory.  Leading dimensions are = 4 (mod 16) doubles so that the 16
// lanes of a half-warp (4 rows x 4 columns of a fragment) hit distinct banks.
// Matrices are zero-padded to multiples of 8 rows / 4 columns in shared memory.
//
// Layout in HBM is the generic one ([draw][inode][k][s][pattern], kernels_gen.cu)
// and the root / reduction kernels are shared with it.
#include <algorithm>
#include <climits>
#include <cstdlib>

#include "engine.cuh"

namespace ttb2 {

namespace {

constexpr int GM_WARPS = 8;
constexpr int GM_TP = 32;    // patterns per tile
constexpr int GM_LDT = 36;   // leading dimension of [rows][32 patterns] tiles
constexpr int GM_MAXACC = 16;  // G accumulator tiles per warp: 2 * (64/8)^2 / 8

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  // volatile: the DMMAs issue in program order -- the call sites interleave their independent
  // accumulator chains explicitly (left to itself the compiler hoists loads across whole phases
  // and spills at the 128-register cap of the 16-warp kernels)
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

struct GmShape {
  int S;    // states
  int Sp;   // rows padded to a multiple of 8
  int Kp;   // contraction length padded to a multiple of 4
  int PLD;  // leading dimension of the staged P matrices (= 4 mod 16)
  int R;    // rows of the pattern tiles = max(Sp, Kp)
};

__host__ __device__ inline GmShape gm_shape(int S) {
  GmShape g;
  g.S = S;
  g.Sp = (S + 7) / 8 * 8;
  g.Kp = (S + 3) / 4 * 4;
  // P is read both as [row][k] and transposed, so it is staged Sp x Sp (>= Kp)
  int ld = g.Sp;
  while (ld % 16 != 4) ++ld;
  g.PLD = ld;
  g.R = g.Sp;
  return g;
}

// zero-padded P [Sp][PLD]
__device__ __forceinline__ void gm_stage_matrix(double* dst, const double* src, const GmShape g) {
  for (int j = threadIdx.x; j < g.Sp * g.PLD; j += blockDim.x) {
    const int r = j / g.PLD, c = j - r * g.PLD;
    dst[j] = (r < g.S && c < g.S) ? src[r * g.S + c] : 0.0;
  }
}

// Tip children need no GEMM: u = P . codeP[code] is one of C vectors.  The table UT[code][s]
// (leading dimension Sp + 1: rows of different codes fall into different banks) takes the place
// of the tip side's staged P, which nothing else reads (the Q phase skips tip children, the
// G phase reads the code vectors from the pattern tile).  Codes below S are unit vectors by
// the C ABI's contract, so their rows are columns of P; only gap / ambiguity codes are summed.
__host__ __device__ inline bool gm_utab_fits(const GmShape g, int codeCount) {
  return (long)codeCount * (g.Sp + 1) <= (long)g.Sp * g.PLD;
}

// (CTA-uniform call: contains a __syncthreads.)  P is read row by row, coalesced, and stored
// transposed -- the unit-vector codes' rows are the columns of P; gap / ambiguity rows are then
// summed from the transposed copy in shared memory.  (Reading P[s][code] per table entry was a
// strided 8-byte gather: tens of microseconds per CTA at 61 states.)
__device__ __forceinline__ void gm_stage_utab(double* dst, const double* P, const double* codeP,
                                              int codeCount, const GmShape g) {
  const int uld = g.Sp + 1;
  for (int j = threadIdx.x; j < g.S * g.S; j += blockDim.x) {
    const int s2 = j / g.S, code = j - s2 * g.S;
    dst[code * uld + s2] = P[j];
  }
  for (int j = threadIdx.x; j < g.S * (g.Sp - g.S); j += blockDim.x) {   // padding rows s2 >= S
    const int code = j / (g.Sp - g.S), s2 = g.S + j - code * (g.Sp - g.S);
    dst[code * uld + s2] = 0.0;
  }
  __syncthreads();
  for (int j = threadIdx.x; j < (codeCount - g.S) * g.Sp; j += blockDim.x) {
    const int code = g.S + j / g.Sp, s2 = j - (code - g.S) * g.Sp;
    double v = 0.0;
    if (s2 < g.S) {
      const double* cp = codeP + (size_t)code * g.S;
      for (int t = 0; t < g.S; ++t) v = fma(dst[t * uld + s2], cp[t], v);
    }
    dst[code * uld + s2] = v;
  }
}

// C fragments of a tip child's u for row tile mt, pattern tiles nt0 .. nt0 + NTG
template <int NTG>
__device__ __forceinline__ void gm_u_tip(double (&c)[NTG][2], const double* ut, int uld,
                                         const uint8_t* codes, int mt, int nt0, int lane) {
  const int row = mt * 8 + (lane >> 2);
#pragma unroll
  for (int n = 0; n < NTG; ++n) {
    const int col = (nt0 + n) * 8 + (lane & 3) * 2;
    c[n][0] = ut[(int)codes[col] * uld + row];
    c[n][1] = ut[(int)codes[col + 1] * uld + row];
  }
}

// D[mt][nt0..nt0+NTG) += A[mt rows][.] . B[.][cols]: A row-major with leading dimension
// lda, B a [contraction][32 patterns] tile.  One A fragment feeds NTG pattern tiles
// (NTG independent accumulator chains, 1 + NTG shared-memory loads per NTG MMAs).
template <int NTG>
__device__ __forceinline__ void gm_mma_ab_g(double (&c)[NTG][2], const double* A, int lda,
                                            const double* Bt, int mt, int nt0, int ksteps,
                                            int lane) {
  const double* a = A + (mt * 8 + (lane >> 2)) * lda + (lane & 3);
  const double* b = Bt + (lane & 3) * GM_LDT + nt0 * 8 + (lane >> 2);
  for (int kt = 0; kt < ksteps; ++kt) {
    const double av = a[kt * 4];
#pragma unroll
    for (int n = 0; n < NTG; ++n) dmma884(c[n][0], c[n][1], av, b[kt * 4 * GM_LDT + n * 8]);
  }
}

// both children in one k loop: 2 * NTG independent accumulator chains in flight
template <int NTG>
__device__ __forceinline__ void gm_mma_ab_g2(double (&cl)[NTG][2], double (&cr)[NTG][2],
                                             const double* Al, const double* Ar, int lda,
                                             const double* Bl, const double* Br, int mt, int nt0,
                                             int ksteps, int lane) {
  const int ao = (mt * 8 + (lane >> 2)) * lda + (lane & 3);
  const int bo = (lane & 3) * GM_LDT + nt0 * 8 + (lane >> 2);
#pragma unroll 4
  for (int kt = 0; kt < ksteps; ++kt) {
    const double al = Al[ao + kt * 4], ar = Ar[ao + kt * 4];
#pragma unroll
    for (int n = 0; n < NTG; ++n) {
      dmma884(cl[n][0], cl[n][1], al, Bl[bo + kt * 4 * GM_LDT + n * 8]);
      dmma884(cr[n][0], cr[n][1], ar, Br[bo + kt * 4 * GM_LDT + n * 8]);
    }
  }
}

template <int NTG>
__device__ __forceinline__ void gm_mma_atb_g(double (&c)[NTG][2], const double* A, int lda,
                                             const double* Bt, int mt, int nt0, int ksteps,
                                             int lane) {
  const double* a = A + (lane & 3) * lda + mt * 8 + (lane >> 2);
  const double* b = Bt + (lane & 3) * GM_LDT + nt0 * 8 + (lane >> 2);
  for (int kt = 0; kt < ksteps; ++kt) {
    const double av = a[kt * 4 * lda];
#pragma unroll
    for (int n = 0; n < NTG; ++n) dmma884(c[n][0], c[n][1], av, b[kt * 4 * GM_LDT + n * 8]);
  }
}

// gm_mma_atb_g with the operands of k-step kt + 1 loaded before the MMAs of k-step kt (compile-
// time trip count, fully unrolled): with two warps per scheduler in the tensor cores a warp that
// loads each B fragment right before its MMA issues one MMA per shared-memory latency and the
// pipe starves (profiles/r02_codon.md)
// shared-memory load the compiler may not move (ptxas otherwise sinks every operand load to
// right before its MMA to save a register)
__device__ __forceinline__ double lds_pinned(unsigned addr) {
  double v;
  asm volatile("ld.volatile.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
  return v;
}

template <int NTG, int KSTEPS>
__device__ __forceinline__ void gm_mma_atb_pf(double (&c)[NTG][2], const double* A, int lda,
                                              const double* Bt, int mt, int lane) {
  const unsigned a = (unsigned)__cvta_generic_to_shared(A + (lane & 3) * lda + mt * 8 + (lane >> 2));
  const unsigned b = (unsigned)__cvta_generic_to_shared(Bt + (lane & 3) * GM_LDT + (lane >> 2));
  double av = lds_pinned(a), bv[NTG];
#pragma unroll
  for (int n = 0; n < NTG; ++n) bv[n] = lds_pinned(b + n * 64);
#pragma unroll
  for (int kt = 0; kt < KSTEPS; ++kt) {
    double an = 0.0, bn[NTG];
    if (kt + 1 < KSTEPS) {
      an = lds_pinned(a + (kt + 1) * 4 * lda * 8);
#pragma unroll
      for (int n = 0; n < NTG; ++n) bn[n] = lds_pinned(b + ((kt + 1) * 4 * GM_LDT + n * 8) * 8);
    }
#pragma unroll
    for (int n = 0; n < NTG; ++n) dmma884(c[n][0], c[n][1], av, bv[n]);
    if (kt + 1 < KSTEPS) {
      av = an;
#pragma unroll
      for (int n = 0; n < NTG; ++n) bv[n] = bn[n];
    }
  }
}

// ===========================================================================
// Per-(pattern, category) rescaling + software-pipelined staging.
// Every (node, category) is independent of the other categories (their chains
// are recombined at the root with their exponent sums, like the fused 4-state
// path), so a CTA owns (node, k, chunk of pattern tiles): the two transition
// matrices are staged once and the child tiles stream through a double buffer
// filled with cp.async while the previous tile is in the tensor cores.
// Exponents: expoK [draw][inode][k][pattern] (int16).
// ===========================================================================
__device__ __forceinline__ void cp_async8(double* smem_dst, const double* gmem_src) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(s), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async16(double* smem_dst, const double* gmem_src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(
                   (unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gmem_src) {
  const unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(sa), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// asynchronous version of gm_stage_tile (tips and padding rows are written directly)
template <int NW>
__device__ __forceinline__ void gm_stage_tile_async(double* tile, bool tip, const uint8_t* tipRow,
                                                    const double* codeP, const double* plane,
                                                    int i0, int Npad, const GmShape g) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int i = i0 + lane;
  if (tip) {
    const double* cp = codeP + (size_t)tipRow[i] * g.S;
    for (int s = warp; s < g.R; s += NW) tile[s * GM_LDT + lane] = s < g.S ? cp[s] : 0.0;
  } else {
    for (int s = warp; s < g.R; s += NW) {
      if (s < g.S) cp_async8(tile + s * GM_LDT + lane, plane + (size_t)s * Npad + i);
      else tile[s * GM_LDT + lane] = 0.0;
    }
  }
}

// post-order v2.  grid (pattern chunks, nodes of level x K, draws)
// shared: Pl Pr [Sp*PLD] | cl[2] cr[2] [R*LDT] | out [Sp*LDT] | wmax [8*32]
template <int NTG, int NW, int CS>
__global__ void __launch_bounds__(NW * 32)
gm_fwd2_kernel(const NodeOp* __restrict__ ops, int opBegin, const double* __restrict__ mats,
               const uint8_t* __restrict__ tips, const double* __restrict__ codeP,
               double* __restrict__ partials, int16_t* __restrict__ expoK, int T, int Npad, int B,
               int K, int Srt, int chunkPatterns, int codeCount) {
  extern __shared__ double sm[];
  const int S = CS ? CS : Srt;  // compile-time state count for the common alphabets
  const GmShape g = gm_shape(S);
  const int SS = S * S;
  const int tileN = g.R * GM_LDT;
  double* Pl = sm;
  double* Pr = Pl + g.Sp * g.PLD;
  double* cl = Pr + g.Sp * g.PLD;   // two buffers
  double* cr = cl + 2 * tileN;      // two buffers
  double* out = cr + 2 * tileN;
  double* wmax = out + tileN;
  uint8_t* codesS = reinterpret_cast<uint8_t*>(wmax + NW * 32);  // [2 buffers][2 sides][32]
  const bool utab = gm_utab_fits(g, codeCount);
  const int uld = g.Sp + 1;

  const int nodeSlot = blockIdx.y / K;
  const int k = blockIdx.y - nodeSlot * K;
  const NodeOp op = ops[opBegin + nodeSlot];
  const int d = blockIdx.z;
  const int I = T - 1;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const bool tipL = op.left < T, tipR = op.right < T;
  const size_t plane = (size_t)S * Npad;
  const size_t nodeStride = (size_t)K * plane;
  double* base = partials + (size_t)d * I * nodeStride + k * plane;
  const double* matsD = mats + (size_t)d * B * K * SS;
  const int MT = g.Sp / 8, KT = g.Kp / 4;
  const uint8_t* tl = tips + (size_t)(tipL ? op.left : 0) * Npad;
  const uint8_t* tr = tips + (size_t)(tipR ? op.right : 0) * Npad;
  const double* pl = base + (size_t)(tipL ? 0 : op.left - T) * nodeStride;
  const double* pr = base + (size_t)(tipR ? 0 : op.right - T) * nodeStride;
  double* qn = base + (size_t)(op.node - T) * nodeStride;
  int16_t* en = expoK + (((size_t)d * I + (op.node - T)) * K + k) * Npad;

  const bool tabL = tipL && utab, tabR = tipR && utab;
  if (tabL) gm_stage_utab(Pl, matsD + ((size_t)op.left * K + k) * SS, codeP, codeCount, g);
  else gm_stage_matrix(Pl, matsD + ((size_t)op.left * K + k) * SS, g);
  if (tabR) gm_stage_utab(Pr, matsD + ((size_t)op.right * K + k) * SS, codeP, codeCount, g);
  else gm_stage_matrix(Pr, matsD + ((size_t)op.right * K + k) * SS, g);
  pdl_wait_then_trigger();   // the matrices come from pmatrix; the tiles below from the previous level

  const int begin = blockIdx.x * chunkPatterns;
  int end = begin + chunkPatterns;
  end = end < Npad ? end : Npad;
  // a tabulated tip child stages its 32 codes instead of a tile of code vectors
  auto stage = [&](int b, int i0) {
    // (the small per-tile arrays are copied asynchronously as well: a plain load + shared store
    // would park the staging warp on the load while the other warps wait for it at the barrier)
    if (tabL) { if (warp == 0 && lane < 8) cp_async4(codesS + b * 64 + 4 * lane, tl + i0 + 4 * lane); }
    else gm_stage_tile_async<NW>(cl + b * tileN, tipL, tl, codeP, pl, i0, Npad, g);
    if (tabR) { if (warp == 1 && lane < 8) cp_async4(codesS + b * 64 + 32 + 4 * lane, tr + i0 + 4 * lane); }
    else gm_stage_tile_async<NW>(cr + b * tileN, tipR, tr, codeP, pr, i0, Npad, g);
  };
  if (begin < end) stage(0, begin);
  cp_async_commit();
  int buf = 0;
  for (int i0 = begin; i0 < end; i0 += GM_TP, buf ^= 1) {
    cp_async_wait_all();
    __syncthreads();  // tile `buf` (and P on the first trip) visible; `out` free
    if (i0 + GM_TP < end) stage(buf ^ 1, i0 + GM_TP);
    cp_async_commit();
    const double* tcl = cl + buf * tileN;
    const double* tcr = cr + buf * tileN;
    const uint8_t* cds = codesS + buf * 64;
    constexpr int NG = 4 / NTG;
    for (int item = warp; item < MT * NG; item += NW) {
      const int mt = item / NG, nt0 = (item - mt * NG) * NTG;
      double accL[NTG][2], accR[NTG][2];
#pragma unroll
      for (int n = 0; n < NTG; ++n) accL[n][0] = accL[n][1] = accR[n][0] = accR[n][1] = 0.0;
      if (tabL) gm_u_tip<NTG>(accL, Pl, uld, cds, mt, nt0, lane);
      else gm_mma_ab_g<NTG>(accL, Pl, g.PLD, tcl, mt, nt0, KT, lane);
      if (tabR) gm_u_tip<NTG>(accR, Pr, uld, cds + 32, mt, nt0, lane);
      else gm_mma_ab_g<NTG>(accR, Pr, g.PLD, tcr, mt, nt0, KT, lane);
      double* o = out + (mt * 8 + (lane >> 2)) * GM_LDT + nt0 * 8 + (lane & 3) * 2;
#pragma unroll
      for (int n = 0; n < NTG; ++n) {
        o[n * 8] = accL[n][0] * accR[n][0];
        o[n * 8 + 1] = accL[n][1] * accR[n][1];
      }
    }
    __syncthreads();
    double m = 0.0;
    for (int s2 = warp; s2 < S; s2 += NW) m = fmax(m, out[s2 * GM_LDT + lane]);
    wmax[warp * 32 + lane] = m;
    __syncthreads();
    double mm = 0.0;
#pragma unroll
    for (int w = 0; w < NW; ++w) mm = fmax(mm, wmax[w * 32 + lane]);
    const int eb = (__double2hiint(mm) >> 20) & 0x7ff;
    const int e = (mm > 0.0) ? (eb > 2044 ? 2044 : eb) - 1022 : 0;
    const double f = __hiloint2double((1023 - e) << 20, 0);
    for (int s2 = warp; s2 < S; s2 += NW)
      qn[(size_t)s2 * Npad + i0 + lane] = out[s2 * GM_LDT + lane] * f;
    if (warp == 0) en[i0 + lane] = (int16_t)e;
  }
}

// ---------------------------------------------------------------------------
// post-order v3 (61-state codon path): warp-specialised.  gm_fwd2_kernel runs each 32-pattern
// tile as MMA phase -> barrier -> column maximum -> barrier -> rescale + store, all 8 warps in
// lock step, and the fp64 tensor pipe idles through every epilogue and every tile load (35 %
// busy, profiles/r02_codon_ncu.md).  Here two groups of 8 warps take alternate tiles and only
// stage + multiply (while one group waits for its cp.async tile the other is in the tensor
// cores), and 8 more warps take the per-pattern maximum, rescale and store the finished tiles;
// the roles meet through named barriers (producer bar.arrive / consumer bar.sync on FULL[group]
// and EMPTY[group]) instead of CTA-wide __syncthreads.
// shared: Pl Pr [Sp*PLD] | cl[2] cr[2] [R*LDT] | out[2] [Sp*LDT] | wmax [2][8*32] | codes
// ---------------------------------------------------------------------------
// bar.arrive / bar.sync order the shared-memory writes of the arriving threads before the reads
// of the threads the barrier releases, so no fence is needed in front of an arrive (a
// __threadfence_block there would also wait for the thread's GLOBAL stores to drain)
__device__ __forceinline__ void named_sync(int id, int count) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
}
__device__ __forceinline__ void named_arrive(int id, int count) {
  asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory");
}

template <int CS>
__global__ void __launch_bounds__(768)
gm_fwd3_kernel(const NodeOp* __restrict__ ops, int opBegin, const double* __restrict__ mats,
               const uint8_t* __restrict__ tips, const double* __restrict__ codeP,
               double* __restrict__ partials, int16_t* __restrict__ expoK,
               double* __restrict__ ustore, int T, int Npad, int B, int K, int Srt,
               int chunkPatterns, int codeCount, int tokenAt) {
  extern __shared__ double sm[];
  // 24 warps: two DMMA groups of 8 (group g takes the tiles t = g, g + 2, ...: while one group
  // waits for its tile to land the other one is in the tensor cores) and 8 epilogue warps
  constexpr int NTG = 4, NWG = 8;
  constexpr int BAR_GROUP = 1, BAR_EPI = 3, BAR_FULL = 4, BAR_EMPTY = 6, BAR_GO = 8;
  const int S = CS ? CS : Srt;
  const GmShape g = gm_shape(S);
  const int SS = S * S;
  const int tileN = g.R * GM_LDT;
  double* Pl = sm;
  double* Pr = Pl + g.Sp * g.PLD;
  double* cl = Pr + g.Sp * g.PLD;   // one buffer per group
  double* cr = cl + 2 * tileN;      // one buffer per group
  double* outB = cr + 2 * tileN;    // one buffer per group
  double* wmaxB = outB + 2 * tileN; // [2][8][32]
  uint8_t* codesS = reinterpret_cast<uint8_t*>(wmaxB + 2 * NWG * 32);  // [2 groups][2 sides][32]
  const bool utab = gm_utab_fits(g, codeCount);
  const int uld = g.Sp + 1;

  const int nodeSlot = blockIdx.y / K;
  const int k = blockIdx.y - nodeSlot * K;
  const NodeOp op = ops[opBegin + nodeSlot];
  const int d = blockIdx.z;
  const int I = T - 1;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const bool tipL = op.left < T, tipR = op.right < T;
  const size_t plane = (size_t)S * Npad;
  const size_t nodeStride = (size_t)K * plane;
  double* base = partials + (size_t)d * I * nodeStride + k * plane;
  const double* matsD = mats + (size_t)d * B * K * SS;
  const int MT = g.Sp / 8, KT = g.Kp / 4;
  const uint8_t* tl = tips + (size_t)(tipL ? op.left : 0) * Npad;
  const uint8_t* tr = tips + (size_t)(tipR ? op.right : 0) * Npad;
  const double* pl = base + (size_t)(tipL ? 0 : op.left - T) * nodeStride;
  const double* pr = base + (size_t)(tipR ? 0 : op.right - T) * nodeStride;
  double* qn = base + (size_t)(op.node - T) * nodeStride;
  int16_t* en = expoK + (((size_t)d * I + (op.node - T)) * K + k) * Npad;
  // u_c = P_c p_c of the internal children, kept for the pre-order sweep (gm_bwd3_kernel<., true>
  // reads them instead of repeating these two products); null: not kept
  double* ul = (ustore && !tipL)
      ? ustore + (size_t)d * I * nodeStride + k * plane + (size_t)(op.left - T) * nodeStride : nullptr;
  double* ur = (ustore && !tipR)
      ? ustore + (size_t)d * I * nodeStride + k * plane + (size_t)(op.right - T) * nodeStride : nullptr;

  const bool tabL = tipL && utab, tabR = tipR && utab;
  if (tabL) gm_stage_utab(Pl, matsD + ((size_t)op.left * K + k) * SS, codeP, codeCount, g);
  else gm_stage_matrix(Pl, matsD + ((size_t)op.left * K + k) * SS, g);
  if (tabR) gm_stage_utab(Pr, matsD + ((size_t)op.right * K + k) * SS, codeP, codeCount, g);
  else gm_stage_matrix(Pr, matsD + ((size_t)op.right * K + k) * SS, g);
  pdl_wait_then_trigger();
  __syncthreads();   // the matrices are staged by all warps

  const int begin = blockIdx.x * chunkPatterns;
  int end = begin + chunkPatterns;
  end = end < Npad ? end : Npad;

  if (warp < 2 * NWG) {
    // ---- DMMA groups: stage own tile, multiply, write u_l o u_r into the group's buffer ----
    const int grp = warp / NWG, gw = warp - grp * NWG;
    double* tcl = cl + grp * tileN;
    double* tcr = cr + grp * tileN;
    uint8_t* cds = codesS + grp * 64;
    double* out = outB + grp * tileN;
    int round = 0;
    // 16 bytes per lane: one instruction moves two rows of the 32-pattern tile; row offsets fixed
    // up front (S <= 64: at most four row pairs per thread); the padding rows are zeroed in the
    // first trip only (nothing overwrites them)
    const int half = lane >> 4, l2 = (lane & 15) * 2;
    const int row0 = gw * 2 + half;
    const size_t srcOff = (size_t)row0 * Npad + l2, srcStep = (size_t)(2 * NWG) * Npad;
    const int dstOff = row0 * GM_LDT + l2;
    // the two groups take turns in the tensor cores (token through BAR_GO + group, see
    // gm_bwd3_kernel): left to themselves they fall into lock step, both waiting for tiles and
    // then sharing the pipe
    for (int i0 = begin + grp * GM_TP; i0 < end; i0 += 2 * GM_TP, ++round) {
      // (the group's previous tile was fully read before the barrier at the end of the last trip)
      const int i = i0 + lane;
      if (tabL) { if (gw == 0 && lane < 8) cp_async4(cds + 4 * lane, tl + i0 + 4 * lane); }
      else if (tipL) {
        const double* cp = codeP + (size_t)tl[i] * g.S;
        for (int s2 = gw; s2 < g.R; s2 += NWG) tcl[s2 * GM_LDT + lane] = s2 < g.S ? cp[s2] : 0.0;
      } else {
        const double* sp = pl + srcOff + i0;
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (row0 + j * 2 * NWG < g.R) {
            if (row0 + j * 2 * NWG < g.S) cp_async16(tcl + dstOff + j * 2 * NWG * GM_LDT, sp + j * srcStep);
            else if (round == 0) tcl[dstOff + j * 2 * NWG * GM_LDT] = tcl[dstOff + j * 2 * NWG * GM_LDT + 1] = 0.0;
          }
      }
      if (tabR) { if (gw == 1 && lane < 8) cp_async4(cds + 32 + 4 * lane, tr + i0 + 4 * lane); }
      else if (tipR) {
        const double* cp = codeP + (size_t)tr[i] * g.S;
        for (int s2 = gw; s2 < g.R; s2 += NWG) tcr[s2 * GM_LDT + lane] = s2 < g.S ? cp[s2] : 0.0;
      } else {
        const double* sp = pr + srcOff + i0;
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (row0 + j * 2 * NWG < g.R) {
            if (row0 + j * 2 * NWG < g.S) cp_async16(tcr + dstOff + j * 2 * NWG * GM_LDT, sp + j * srcStep);
            else if (round == 0) tcr[dstOff + j * 2 * NWG * GM_LDT] = tcr[dstOff + j * 2 * NWG * GM_LDT + 1] = 0.0;
          }
      }
      cp_async_commit();
      cp_async_wait_all();
      named_sync(BAR_GROUP + grp, NWG * 32);   // the tile is visible to the whole group
      if (round >= 1) named_sync(BAR_EMPTY + grp, 512);   // the epilogue released out[grp]
      if (tokenAt && (grp == 1 || round > 0)) named_sync(BAR_GO + grp, 512);
      for (int mt = gw; mt < MT; mt += NWG) {
        double accL[NTG][2], accR[NTG][2];
#pragma unroll
        for (int n = 0; n < NTG; ++n) accL[n][0] = accL[n][1] = accR[n][0] = accR[n][1] = 0.0;
        if (!tabL && !tabR) {
          gm_mma_ab_g2<NTG>(accL, accR, Pl, Pr, g.PLD, tcl, tcr, mt, 0, KT, lane);
        } else {
          if (tabL) gm_u_tip<NTG>(accL, Pl, uld, cds, mt, 0, lane);
          else gm_mma_ab_g<NTG>(accL, Pl, g.PLD, tcl, mt, 0, KT, lane);
          if (tabR) gm_u_tip<NTG>(accR, Pr, uld, cds + 32, mt, 0, lane);
          else gm_mma_ab_g<NTG>(accR, Pr, g.PLD, tcr, mt, 0, KT, lane);
        }
        double* o = out + (mt * 8 + (lane >> 2)) * GM_LDT + (lane & 3) * 2;
#pragma unroll
        for (int n = 0; n < NTG; ++n) {
          o[n * 8] = accL[n][0] * accR[n][0];
          o[n * 8 + 1] = accL[n][1] * accR[n][1];
        }
        // straight from the accumulator fragments: 64-byte runs per row
        const int row = mt * 8 + (lane >> 2);
        if (row < S) {
          const size_t at = (size_t)row * Npad + i0 + (lane & 3) * 2;
          if (ul) {
#pragma unroll
            for (int n = 0; n < NTG; ++n)
              *reinterpret_cast<double2*>(ul + at + n * 8) = make_double2(accL[n][0], accL[n][1]);
          }
          if (ur) {
#pragma unroll
            for (int n = 0; n < NTG; ++n)
              *reinterpret_cast<double2*>(ur + at + n * 8) = make_double2(accR[n][0], accR[n][1]);
          }
        }
      }
      if (tokenAt && i0 + GM_TP < end) named_arrive(BAR_GO + (grp ^ 1), 512);
      named_arrive(BAR_FULL + grp, 512);
      named_sync(BAR_GROUP + grp, NWG * 32);   // every warp is done reading the tile
    }
  } else {
    // ---- epilogue warps: per-pattern maximum, power-of-two rescaling, stores ----
    // The fp64 pipe belongs to the DMMAs: an fmax (DSETP + selects) or DMUL issued from here
    // queues behind them ("math pipe throttle" was the top stall of these warps,
    // profiles/r02_codon.md).  The maximum is therefore taken on the integer pipe: for
    // non-negative doubles the order of the high words is the order of the values and only the
    // exponent of the maximum is needed.  mh: largest high word; lz: OR of the low words of the
    // values whose high word is zero (max > 0 <=> mh > 0 or lz != 0, as in the floating-point
    // test); each value is read from shared memory once and kept for the rescaled store.
    const int ew = warp - 2 * NWG;
    constexpr int RPW = CS ? (CS + NWG - 1) / NWG : 8;   // rows per warp (S <= 64)
    int2* wmax2 = reinterpret_cast<int2*>(wmaxB);        // [2][NWG][32]
    double* qrow = qn + (size_t)ew * Npad + lane;
    const size_t rowStep = (size_t)NWG * Npad;
    int t = 0;
    for (int i0 = begin; i0 < end; i0 += GM_TP, ++t) {
      const int buf = t & 1;
      named_sync(BAR_FULL + buf, 512);
      const double* out = outB + buf * tileN + ew * GM_LDT + lane;
      double v[RPW];
      int mh = 0, lz = 0;
#pragma unroll
      for (int j = 0; j < RPW; ++j) {
        v[j] = (ew + j * NWG < S) ? out[j * NWG * GM_LDT] : 0.0;
        const int hi = __double2hiint(v[j]), lo = __double2loint(v[j]);
        mh = max(mh, hi);
        lz |= (hi == 0) ? lo : 0;
      }
      wmax2[(buf * NWG + ew) * 32 + lane] = make_int2(mh, lz);
      named_sync(BAR_EPI, NWG * 32);
#pragma unroll
      for (int w = 0; w < NWG; ++w) {
        const int2 o = wmax2[(buf * NWG + w) * 32 + lane];
        mh = max(mh, o.x);
        lz |= o.y;
      }
      const int eb = (mh >> 20) & 0x7ff;
      const int e = (mh > 0 || lz != 0) ? (eb > 2044 ? 2044 : eb) - 1022 : 0;
      const double f = __hiloint2double((1023 - e) << 20, 0);
#pragma unroll
      for (int j = 0; j < RPW; ++j)
        if (ew + j * NWG < S) qrow[j * rowStep + i0] = v[j] * f;
      if (ew == 0) en[i0 + lane] = (int16_t)e;
      named_arrive(BAR_EMPTY + buf, 512);
    }
  }
}

// ---------------------------------------------------------------------------
// post-order, level 1 (both children are tips; a third of the nodes of a random tree): no GEMM
// at all -- u_c = P_c . codeP[code] is a table row, so the node's vector is the product of two
// table rows.  The tile kernels spend 1.7 us per 32-pattern tile on this (barriers, 1 CTA per
// SM: 1.4 TB/s of stores, profiles/r02_codon_ncu.md); this kernel is a plain streaming kernel:
// thread = pattern, the two C x S tables in shared memory, coalesced stores.
// grid (pattern blocks, nodes x K, draws), 128 threads
// ---------------------------------------------------------------------------
constexpr int GMC_THREADS = 256;

template <int CS>
__global__ void __launch_bounds__(GMC_THREADS)
gm_cherry_fwd_kernel(const NodeOp* __restrict__ ops, int opBegin, const double* __restrict__ mats,
                     const uint8_t* __restrict__ tips, const double* __restrict__ codeP,
                     double* __restrict__ partials, int16_t* __restrict__ expoK, int T, int Npad,
                     int B, int K, int Srt, int patternsPerBlock, int codeCount) {
  extern __shared__ double sm[];
  const int S = CS ? CS : Srt;
  const int LD = S | 1;   // odd leading dimension: rows of different codes in different banks
  double* utL = sm;
  double* utR = sm + (size_t)codeCount * LD;
  const int nodeSlot = blockIdx.y / K;
  const int k = blockIdx.y - nodeSlot * K;
  const NodeOp op = ops[opBegin + nodeSlot];
  const int d = blockIdx.z;
  const int I = T - 1;
  const int SS = S * S;
  const double* matsD = mats + (size_t)d * B * K * SS;
  for (int j = threadIdx.x; j < 2 * SS; j += blockDim.x) {   // transposed, coalesced reads of P
    const int side = j >= SS, jj = j - side * SS;
    const int s2 = jj / S, code = jj - s2 * S;
    const double* P = matsD + ((size_t)(side ? op.right : op.left) * K + k) * SS;
    (side ? utR : utL)[code * LD + s2] = P[jj];
  }
  __syncthreads();
  for (int j = threadIdx.x; j < 2 * (codeCount - S) * S; j += blockDim.x) {   // gap / ambiguity rows
    const int side = j >= (codeCount - S) * S, jj = j - side * (codeCount - S) * S;
    const int code = S + jj / S, s2 = jj - (code - S) * S;
    double* ut = side ? utR : utL;
    const double* cp = codeP + (size_t)code * S;
    double v = 0.0;
    for (int t = 0; t < S; ++t) v = fma(ut[t * LD + s2], cp[t], v);
    ut[code * LD + s2] = v;
  }
  pdl_wait_then_trigger();
  __syncthreads();
  const size_t plane = (size_t)S * Npad;
  double* qn = partials + ((size_t)d * I + (op.node - T)) * K * plane + k * plane;
  int16_t* en = expoK + (((size_t)d * I + (op.node - T)) * K + k) * Npad;
  const uint8_t* tl = tips + (size_t)op.left * Npad;
  const uint8_t* tr = tips + (size_t)op.right * Npad;
  const int begin = blockIdx.x * patternsPerBlock;
  int end = begin + patternsPerBlock;
  end = end < Npad ? end : Npad;
  for (int i = begin + threadIdx.x; i < end; i += blockDim.x) {
    const double* a = utL + (int)tl[i] * LD;
    const double* b = utR + (int)tr[i] * LD;
    double m = 0.0;
    if (CS > 0) {   // compile-time state count: the products stay in registers
      double v[CS > 0 ? CS : 1];
#pragma unroll
      for (int s2 = 0; s2 < CS; ++s2) {
        v[s2] = a[s2] * b[s2];
        m = fmax(m, v[s2]);
      }
      const int eb = (__double2hiint(m) >> 20) & 0x7ff;
      const int e = (m > 0.0) ? (eb > 2044 ? 2044 : eb) - 1022 : 0;
      const double f = __hiloint2double((1023 - e) << 20, 0);
#pragma unroll
      for (int s2 = 0; s2 < CS; ++s2) qn[(size_t)s2 * Npad + i] = v[s2] * f;
      en[i] = (int16_t)e;
    } else {
#pragma unroll 4
      for (int s2 = 0; s2 < S; ++s2) m = fmax(m, a[s2] * b[s2]);
      const int eb = (__double2hiint(m) >> 20) & 0x7ff;
      const int e = (m > 0.0) ? (eb > 2044 ? 2044 : eb) - 1022 : 0;
      const double f = __hiloint2double((1023 - e) << 20, 0);
#pragma unroll 4
      for (int s2 = 0; s2 < S; ++s2) qn[(size_t)s2 * Npad + i] = a[s2] * b[s2] * f;
      en[i] = (int16_t)e;
    }
  }
}

// root v2: recombine the per-category chains (cf. fused_root_kernel)
__device__ __forceinline__ double gm_chain_weight(int diff) {
  return diff < -1000 ? 0.0 : __hiloint2double((1023 + (diff > 1000 ? 1000 : diff)) << 20, 0);
}

constexpr int GM_MAXK = 16;

// Root of the DMMA paths: per pattern, the K category chains are recombined with their exponent
// sums (lnL = log sum_k rho_k 2^(e_k - emax) pi . p_k + emax ln 2).
// WITH_PRE: also writes q^_root and the per-block partials of d/d rho, root d/d pi.
// Block = 32 patterns x K categories (thread (lane, k) owns one chain: its S-term dot product and
// its sum of I scale exponents), so that the kernel has K times the threads of a pattern-per-
// thread layout and the d/d rho, d/d pi partial sums are warp reductions plus ONE barrier (the
// first version took a block-wide sum with two barriers for each of the K + S outputs: 0.19 ms
// at 20k codon patterns, one CTA of 4 warps per SM).
template <bool WITH_PRE>
__global__ void __launch_bounds__(32 * GM_MAXK)
gm_root2_kernel(const double* __restrict__ partials, const int16_t* __restrict__ expoK,
                const double* __restrict__ freqs, int freqDraws,
                const double* __restrict__ props, int propDraws,
                const double* __restrict__ weights, double* __restrict__ siteLnl,
                double* __restrict__ pre, double* __restrict__ blockPart, int T, int Npad, int K,
                int S, int rootInode) {
  __shared__ double sDot[GM_MAXK][32];
  __shared__ int sE[GM_MAXK][32];
  __shared__ double sRed[GM_MAXK][64 + GM_MAXK];
  const int d = blockIdx.y;
  const int I = T - 1;
  const int lane = threadIdx.x, k = threadIdx.y;
  const int i = blockIdx.x * 32 + lane;   // Npad is a multiple of 32: always a valid pattern
  const double* fr = freqs + (freqDraws > 1 ? (size_t)d * S : 0);
  const double* pr = props + (propDraws > 1 ? (size_t)d * K : 0);
  const size_t plane = (size_t)S * Npad;
  const double* p = partials + (((size_t)d * I + rootInode) * K + k) * plane + i;
  {
    double dot = 0.0;
    for (int s = 0; s < S; ++s) dot = fma(fr[s], p[(size_t)s * Npad], dot);
    int e = 0;
    const int16_t* ep = expoK + ((size_t)d * I * K + k) * Npad + i;
    for (int n = 0; n < I; ++n) e += ep[(size_t)n * K * Npad];
    sDot[k][lane] = dot;
    sE[k][lane] = e;
  }
  __syncthreads();
  int emax = INT_MIN;
  for (int kk = 0; kk < K; ++kk)
    if (sDot[kk][lane] > 0.0 && pr[kk] > 0.0) emax = max(emax, sE[kk][lane]);
  double L = 0.0;
  for (int kk = 0; kk < K; ++kk) {
    const double cw = sDot[kk][lane] > 0.0 ? gm_chain_weight(sE[kk][lane] - emax) : 0.0;
    L = fma(pr[kk] * cw, sDot[kk][lane], L);
  }
  const double site = log(L) + (double)emax * 0.693147180559945309417232121458;
  const double invL = 1.0 / L;
  const double w = weights[i];
  if (!WITH_PRE) {
    if (k == 0) {
      siteLnl[(size_t)d * Npad + i] = site;
      double v = w != 0.0 ? w * site : 0.0;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (lane == 0) blockPart[(size_t)d * gridDim.x + blockIdx.x] = v;
    }
    return;
  }
  const double dotk = sDot[k][lane];
  const double cwk = dotk > 0.0 ? gm_chain_weight(sE[k][lane] - emax) : 0.0;
  {
    double* q = pre + (((size_t)d * I + rootInode) * K + k) * plane + i;
    const int e = expoK[(((size_t)d * I + rootInode) * K + k) * Npad + i];
    const double c = pr[k] * cwk * invL * __hiloint2double((1023 - e) << 20, 0);
    for (int s = 0; s < S; ++s) q[(size_t)s * Npad] = c * fr[s];
  }
  const double wl = w != 0.0 ? w * invL : 0.0;
  {
    double v = (wl != 0.0 && dotk > 0.0) ? wl * cwk * dotk : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) sRed[k][S] = v;
  }
  const double ck = (wl != 0.0 && dotk > 0.0) ? wl * pr[k] * cwk : 0.0;
  for (int s = 0; s < S; ++s) {
    double v = ck != 0.0 ? ck * p[(size_t)s * Npad] : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) sRed[k][s] = v;
  }
  __syncthreads();
  double* out = blockPart + ((size_t)d * gridDim.x + blockIdx.x) * (K + S);
  const int t0 = k * 32 + lane;
  if (t0 < K) out[t0] = sRed[t0][S];
  for (int t = t0; t < S; t += 32 * K) {
    double acc = 0.0;
    for (int kk = 0; kk < K; ++kk) acc += sRed[kk][t];
    out[K + t] = acc;
  }
}

// ---------------------------------------------------------------------------
// pre-order v2 (per-category exponents, double-buffered cp.async staging).
// shared: Pl Pr [Sp*PLD] | tq[2] vl[2] vr[2] ml mr [R*LDT each] | ws[2][32] | el er [2][32] (int)
// ---------------------------------------------------------------------------
template <int NTG, int NW, int CS>
__global__ void __launch_bounds__(NW * 32)
gm_bwd2_kernel(const NodeOp* __restrict__ ops, int opBegin, const double* __restrict__ mats,
              const uint8_t* __restrict__ tips, const double* __restrict__ codeP,
              const double* __restrict__ partials, const int16_t* __restrict__ expoK,
              const double* __restrict__ weights, double* __restrict__ pre,
              double* __restrict__ gpart, const int* __restrict__ chunkBase, size_t chunkTotal,
              int T, int Npad, int B, int K, int Srt, int chunkPatterns, int nChunk,
              int codeCount) {
  extern __shared__ double sm[];
  const int S = CS ? CS : Srt;
  const GmShape g = gm_shape(S);
  const int SS = S * S;
  const bool utab = gm_utab_fits(g, codeCount);
  const int uld = g.Sp + 1;
  double* Pl = sm;
  double* Pr = Pl + g.Sp * g.PLD;
  const int tileN = g.R * GM_LDT;
  double* tqB = Pr + g.Sp * g.PLD;   // 2 buffers each
  double* vlB = tqB + 2 * tileN;
  double* vrB = vlB + 2 * tileN;
  double* ml = vrB + 2 * tileN;
  double* mr = ml + tileN;
  double* wsB = mr + tileN;          // [2][32]
  int16_t* seB = reinterpret_cast<int16_t*>(wsB + 64);  // [2][el[32], er[32]] (512 bytes reserved)
  uint8_t* codesS = reinterpret_cast<uint8_t*>(wsB + 64) + 512;  // [2 buffers][2 sides][32]

  const int nodeSlot = blockIdx.y / K;
  const int k = blockIdx.y - nodeSlot * K;
  const NodeOp op = ops[opBegin + nodeSlot];
  const int d = blockIdx.z;
  const int I = T - 1;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const bool tipL = op.left < T, tipR = op.right < T;
  const size_t plane = (size_t)S * Npad;
  const size_t nodeStride = (size_t)K * plane;
  const size_t drawBase = (size_t)d * I * nodeStride;
  const double* matsD = mats + (size_t)d * B * K * SS;
  const int MT = g.Sp / 8, KT = g.Kp / 4, KTr = g.Sp / 4;
  // tabulated tip children: UT takes the place of P (the Q phase skips tips, G reads the tiles)
  const bool tabL = tipL && utab, tabR = tipR && utab;
  if (tabL) gm_stage_utab(Pl, matsD + ((size_t)op.left * K + k) * SS, codeP, codeCount, g);
  else gm_stage_matrix(Pl, matsD + ((size_t)op.left * K + k) * SS, g);
  if (tabR) gm_stage_utab(Pr, matsD + ((size_t)op.right * K + k) * SS, codeP, codeCount, g);
  else gm_stage_matrix(Pr, matsD + ((size_t)op.right * K + k) * SS, g);
  pdl_wait_then_trigger();

  // persistent G accumulators: (child, row tile) combos dealt round-robin to the warps, NCJ per
  // warp (16 combos at most: 2 children x 8 row tiles), 8 column tiles each
  constexpr int NCJ = (16 + NW - 1) / NW;
  double acc[NCJ * 8][2];
#pragma unroll
  for (int j = 0; j < NCJ * 8; ++j) acc[j][0] = acc[j][1] = 0.0;
  // Tip children whose codes are unit vectors or the all-ones gap (no ambiguity rows in the code
  // table): the B operand of the G product, v_c[t][p] = [code_p == t or gap], is made in
  // registers from the tile's 32 codes -- no tile of code vectors is staged or read for them.
  const bool hotL = tabL && codeCount == S + 1, hotR = tabR && codeCount == S + 1;

  const int begin = blockIdx.x * chunkPatterns;
  int end = begin + chunkPatterns;
  end = end < Npad ? end : Npad;
  const double* qsrc = pre + drawBase + (size_t)(op.node - T) * nodeStride + k * plane;
  const double* lsrc = partials + drawBase + (size_t)(tipL ? 0 : op.left - T) * nodeStride + k * plane;
  const double* rsrc = partials + drawBase + (size_t)(tipR ? 0 : op.right - T) * nodeStride + k * plane;
  const uint8_t* tl = tips + (size_t)(tipL ? op.left : 0) * Npad;
  const uint8_t* tr = tips + (size_t)(tipR ? op.right : 0) * Npad;
  const int16_t* elp = tipL ? nullptr : expoK + (((size_t)d * I + (op.left - T)) * K + k) * Npad;
  const int16_t* erp = tipR ? nullptr : expoK + (((size_t)d * I + (op.right - T)) * K + k) * Npad;
  auto stage = [&](int b, int i0) {
    gm_stage_tile_async<NW>(tqB + b * tileN, false, nullptr, codeP, qsrc, i0, Npad, g);
    if (!hotL) gm_stage_tile_async<NW>(vlB + b * tileN, tipL, tl, codeP, lsrc, i0, Npad, g);
    if (!hotR) gm_stage_tile_async<NW>(vrB + b * tileN, tipR, tr, codeP, rsrc, i0, Npad, g);
    // weights, child exponents and tip codes: asynchronous copies too -- a plain load + shared
    // store would park warp 0 on the load while the other warps wait for it at the barrier
    if (warp == 0) {
      cp_async8(wsB + b * 32 + lane, weights + i0 + lane);
      if (lane < 16) {
        if (!tipL) cp_async4(seB + b * 64 + 2 * lane, elp + i0 + 2 * lane);
        if (!tipR) cp_async4(seB + b * 64 + 32 + 2 * lane, erp + i0 + 2 * lane);
      } else if (lane < 24) {
        if (tabL) cp_async4(codesS + b * 64 + 4 * (lane - 16), tl + i0 + 4 * (lane - 16));
      } else {
        if (tabR) cp_async4(codesS + b * 64 + 32 + 4 * (lane - 24), tr + i0 + 4 * (lane - 24));
      }
    }
  };
  if (begin < end) stage(0, begin);
  cp_async_commit();
  int buf = 0;
  for (int i0 = begin; i0 < end; i0 += GM_TP, buf ^= 1) {
    cp_async_wait_all();
    __syncthreads();  // buffers `buf` visible; ml/mr and buffers `buf^1` free
    if (i0 + GM_TP < end) stage(buf ^ 1, i0 + GM_TP);
    cp_async_commit();
    const double* tq = tqB + buf * tileN;
    const double* vl = vlB + buf * tileN;
    const double* vr = vrB + buf * tileN;
    const double* ws = wsB + buf * 32;
    const int16_t* se = seB + buf * 64;
    // U phase: u_l = P_l v_l, u_r = P_r v_r;  m_l = q^ o u_r, m_r = q^ o u_l
    constexpr int NG = 4 / NTG;
    for (int item = warp; item < MT * NG; item += NW) {
      const int mt = item / NG, nt0 = (item - mt * NG) * NTG;
      double accL[NTG][2], accR[NTG][2];
#pragma unroll
      for (int n = 0; n < NTG; ++n) accL[n][0] = accL[n][1] = accR[n][0] = accR[n][1] = 0.0;
      if (tabL) gm_u_tip<NTG>(accL, Pl, uld, codesS + buf * 64, mt, nt0, lane);
      else gm_mma_ab_g<NTG>(accL, Pl, g.PLD, vl, mt, nt0, KT, lane);
      if (tabR) gm_u_tip<NTG>(accR, Pr, uld, codesS + buf * 64 + 32, mt, nt0, lane);
      else gm_mma_ab_g<NTG>(accR, Pr, g.PLD, vr, mt, nt0, KT, lane);
#pragma unroll
      for (int n = 0; n < NTG; ++n) {
        const int off = (mt * 8 + (lane >> 2)) * GM_LDT + (nt0 + n) * 8 + (lane & 3) * 2;
        const double q0 = tq[off], q1 = tq[off + 1];
        ml[off] = q0 * accR[n][0];
        ml[off + 1] = q1 * accR[n][1];
        mr[off] = q0 * accL[n][0];
        mr[off + 1] = q1 * accL[n][1];
      }
    }
    __syncthreads();
    // Q phase: q^_c = P_c^T m_c * 2^{-e_c}  (internal children)
    for (int side = 0; side < 2; ++side) {
      if (side ? tipR : tipL) continue;
      const double* P = side ? Pr : Pl;
      const double* mm = side ? mr : ml;
      const int child = side ? op.right : op.left;
      double* qout = pre + drawBase + (size_t)(child - T) * nodeStride + k * plane + i0;
      for (int item = warp; item < MT * NG; item += NW) {
        const int mt = item / NG, nt0 = (item - mt * NG) * NTG;
        double c[NTG][2];
#pragma unroll
        for (int n = 0; n < NTG; ++n) c[n][0] = c[n][1] = 0.0;
        gm_mma_atb_g<NTG>(c, P, g.PLD, mm, mt, nt0, KTr, lane);
        const int row = mt * 8 + (lane >> 2);
        if (row < S) {
#pragma unroll
          for (int n = 0; n < NTG; ++n) {
            const int col = (nt0 + n) * 8 + (lane & 3) * 2;
            const double f0 = __hiloint2double((1023 - (int)se[side * 32 + col]) << 20, 0);
            const double f1 = __hiloint2double((1023 - (int)se[side * 32 + col + 1]) << 20, 0);
            *reinterpret_cast<double2*>(qout + (size_t)row * Npad + col) =
                make_double2(c[n][0] * f0, c[n][1] * f1);
          }
        }
      }
    }
    // G phase: G_c[s][t] += sum_p (w_p m_c[s][p]) v_c[t][p].  A warp owns up to two
    // (child, row tile) combos and all their column tiles: one A fragment (w o m)
    // feeds MT independent accumulator chains.
#pragma unroll
    for (int cj = 0; cj < NCJ; ++cj) {
      const int combo = warp + cj * NW;
      if (combo < 2 * MT) {
        const int side = combo / MT;
        const int mt = combo - side * MT;
        const double* mm = (side ? mr : ml) + (mt * 8 + (lane >> 2)) * GM_LDT + (lane & 3);
        const double* wp = ws + (lane & 3);
        if (side ? hotR : hotL) {
          // (skipping the column tiles that hold none of a k-step's 4 codes was tried: fewer
          // DMMAs, but the warp-uniform branches cost more than they save -- 12.8 vs 12.4 ms)
          const uint8_t* cds = codesS + buf * 64 + side * 32 + (lane & 3);
          const int colBase = lane >> 2;
#pragma unroll
          for (int kt = 0; kt < GM_TP / 4; ++kt) {
            const double av = mm[kt * 4] * wp[kt * 4];
            const int code = cds[kt * 4];
#pragma unroll
            for (int mt2 = 0; mt2 < 8; ++mt2)
              if (mt2 < MT) {
                const int col = mt2 * 8 + colBase;
                const double bv = (col < S && (code == col || code == S)) ? 1.0 : 0.0;
                dmma884(acc[cj * 8 + mt2][0], acc[cj * 8 + mt2][1], av, bv);
              }
          }
        } else {
          const double* vv = (side ? vr : vl) + (lane >> 2) * GM_LDT + (lane & 3);
#pragma unroll
          for (int kt = 0; kt < GM_TP / 4; ++kt) {
            const double av = mm[kt * 4] * wp[kt * 4];
#pragma unroll
            for (int mt2 = 0; mt2 < 8; ++mt2)
              if (mt2 < MT)
                dmma884(acc[cj * 8 + mt2][0], acc[cj * 8 + mt2][1], av,
                        vv[mt2 * 8 * GM_LDT + kt * 4]);
          }
        }
      }
    }
  }
#pragma unroll
  for (int cj = 0; cj < NCJ; ++cj) {
    const int combo = warp + cj * NW;
    if (combo < 2 * MT) {
      const int side = combo / MT;
      const int mt = combo - side * MT;
      const int branch = side ? op.right : op.left;
      double* o = gpart + ((size_t)d * chunkTotal + chunkBase[branch] + (size_t)k * nChunk +
                           blockIdx.x) * SS;
      const int row = mt * 8 + (lane >> 2);
#pragma unroll
      for (int mt2 = 0; mt2 < 8; ++mt2) {
        if (mt2 < MT && row < S) {
          const int col = mt2 * 8 + (lane & 3) * 2;
          if (col < S) o[row * S + col] = acc[cj * 8 + mt2][0];
          if (col + 1 < S) o[row * S + col + 1] = acc[cj * 8 + mt2][1];
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------
// pre-order v3 (61-state codon path): two groups of 8 warps take alternate 32-pattern tiles.
// gm_bwd2_kernel has all its warps in lock step (tile wait -> U phase -> barrier -> Q + G phase),
// so the fp64 tensor pipe idles while a tile lands and around every barrier (66 % busy on the
// large levels, profiles/r02_codon_ncu.md).  Here each group owns single-buffered tiles and
// synchronises only with itself (named barriers): while one group waits for its cp.async tile
// or sits at its barrier the other is in the tensor cores.  m_l overwrites the q^ tile in place
// (every element is read and written by the same thread), so a group needs four tiles
// (q^ -> m_l, v_l, v_r, m_r) and both groups fit beside the two staged matrices.  Each group
// accumulates its own partial G in registers; they are added in fixed order (group 0 + group 1)
// through shared memory at the end.
// shared: Pl Pr [Sp*PLD] | per group: tq vl vr mr [R*LDT] , ws [32], se [64 int16], codes [64]
// ---------------------------------------------------------------------------
#ifdef TTB2_GM_TRACE
// phase timeline of one CTA (tools/profile_eval.py --trace): [group][trip][6] SM clock stamps
__device__ long long gm_trace[2 * 64 * 6];
#define GM_TRACE(slot)                                                                    \
  if (blockIdx.x == 1 && blockIdx.y == 0 && gw == 0 && lane == 0 && trip < 64)            \
    gm_trace[(grp * 64 + trip) * 6 + (slot)] = clock64();
#else
#define GM_TRACE(slot)
#endif

// US: the post-order sweep kept u_c = P_c p_c of every internal child (Engine::ustore), so the U
// phase is two element-wise products instead of two of the six matrix products per tile: the
// u_r tile lands in the m_l buffer and the u_l tile in the m_r buffer (m_l = q^ o u_r, m_r =
// q^ o u_l are formed in place, every element by the thread that read it), and q^ is read
// straight into the accumulator-fragment layout.
template <int CS, bool US>
__global__ void __launch_bounds__(512)
gm_bwd3_kernel(const NodeOp* __restrict__ ops, int opBegin, const double* __restrict__ mats,
               const uint8_t* __restrict__ tips, const double* __restrict__ codeP,
               const double* __restrict__ partials, const int16_t* __restrict__ expoK,
               const double* __restrict__ ustore,
               const double* __restrict__ weights, double* __restrict__ pre,
               double* __restrict__ gpart, const int* __restrict__ chunkBase, size_t chunkTotal,
               int T, int Npad, int B, int K, int chunkPatterns, int nChunk, int codeCount,
               int tokenAt) {
  extern __shared__ double sm[];
  constexpr int S = CS, NTG = 4, NWG = 8, BAR_GROUP = 1, BAR_GO = 3;
  const GmShape g = gm_shape(S);
  constexpr int SS = S * S;
  const bool utab = gm_utab_fits(g, codeCount);
  const int uld = g.Sp + 1;
  double* Pl = sm;
  double* Pr = Pl + g.Sp * g.PLD;
  const int tileN = g.R * GM_LDT;
  const int groupN = 4 * tileN + 32 + 16 + 8;   // tiles | weights | exponents | codes
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int grp = warp / NWG, gw = warp - grp * NWG;
  double* mine = Pr + g.Sp * g.PLD + grp * groupN;
  double* tq = mine;                 // becomes m_l after the U phase
  double* vl = tq + tileN;
  double* vr = vl + tileN;
  double* mr = vr + tileN;
  double* ws = mr + tileN;
  int16_t* se = reinterpret_cast<int16_t*>(ws + 32);            // [el[32], er[32]]
  uint8_t* codesS = reinterpret_cast<uint8_t*>(ws + 32 + 16);   // [2 sides][32]

  const int nodeSlot = blockIdx.y / K;
  const int k = blockIdx.y - nodeSlot * K;
  const NodeOp op = ops[opBegin + nodeSlot];
  const int d = blockIdx.z;
  const int I = T - 1;
  const bool tipL = op.left < T, tipR = op.right < T;
  const size_t plane = (size_t)S * Npad;
  const size_t nodeStride = (size_t)K * plane;
  const size_t drawBase = (size_t)d * I * nodeStride;
  const double* matsD = mats + (size_t)d * B * K * SS;
  const int MT = g.Sp / 8, KT = g.Kp / 4, KTr = g.Sp / 4;
  const bool tabL = tipL && utab, tabR = tipR && utab;
  if (tabL) gm_stage_utab(Pl, matsD + ((size_t)op.left * K + k) * SS, codeP, codeCount, g);
  else gm_stage_matrix(Pl, matsD + ((size_t)op.left * K + k) * SS, g);
  if (tabR) gm_stage_utab(Pr, matsD + ((size_t)op.right * K + k) * SS, codeP, codeCount, g);
  else gm_stage_matrix(Pr, matsD + ((size_t)op.right * K + k) * SS, g);
  pdl_wait_then_trigger();
  __syncthreads();

  constexpr int NCJ = 2;   // 16 (child, row tile) combos over the 8 warps of a group
  double acc[NCJ * 8][2];
#pragma unroll
  for (int j = 0; j < NCJ * 8; ++j) acc[j][0] = acc[j][1] = 0.0;
  const bool hotL = tabL && codeCount == S + 1, hotR = tabR && codeCount == S + 1;

  const int begin = blockIdx.x * chunkPatterns;
  int end = begin + chunkPatterns;
  end = end < Npad ? end : Npad;
  const double* qsrc = pre + drawBase + (size_t)(op.node - T) * nodeStride + k * plane;
  const double* lsrc = partials + drawBase + (size_t)(tipL ? 0 : op.left - T) * nodeStride + k * plane;
  const double* rsrc = partials + drawBase + (size_t)(tipR ? 0 : op.right - T) * nodeStride + k * plane;
  const double* ulsrc = US ? ustore + drawBase + (size_t)(tipL ? 0 : op.left - T) * nodeStride + k * plane : nullptr;
  const double* ursrc = US ? ustore + drawBase + (size_t)(tipR ? 0 : op.right - T) * nodeStride + k * plane : nullptr;
  static_assert(!US || (CS + 7) / 8 == NWG, "US: one row tile per warp of a group");
  const uint8_t* tl = tips + (size_t)(tipL ? op.left : 0) * Npad;
  const uint8_t* tr = tips + (size_t)(tipR ? op.right : 0) * Npad;
  const int16_t* elp = tipL ? nullptr : expoK + (((size_t)d * I + (op.left - T)) * K + k) * Npad;
  const int16_t* erp = tipR ? nullptr : expoK + (((size_t)d * I + (op.right - T)) * K + k) * Npad;
  // one [R][LDT] tile of this group <- rows of `plane`, or the code vectors of a tip.
  // 16 bytes per lane: one instruction moves two rows of the 32-pattern tile, four per tile and
  // thread with the row offsets fixed up front (the staging code used to be the largest block of
  // executed instructions of the kernel, profiles/r02_codon.md); rows S .. R - 1 are zeroed once
  constexpr int RR = (CS + 7) / 8 * 8;
  static_assert(RR % (2 * NWG) == 0, "row pairs per warp");
  const int half = lane >> 4, l2 = (lane & 15) * 2;
  const int row0 = gw * 2 + half;
  const size_t srcOff = (size_t)row0 * Npad + l2, srcStep = (size_t)(2 * NWG) * Npad;
  const int dstOff = row0 * GM_LDT + l2;
  for (int idx = gw * 32 + lane; idx < 4 * (RR - S) * GM_LDT; idx += NWG * 32) {
    const int t4 = idx / ((RR - S) * GM_LDT), rem = idx - t4 * ((RR - S) * GM_LDT);
    mine[t4 * tileN + S * GM_LDT + rem] = 0.0;
  }
  auto stage_tile = [&](double* tile, bool tip, const uint8_t* tipRow, const double* src, int i0) {
    if (tip) {
      const double* cp = codeP + (size_t)tipRow[i0 + lane] * S;
      for (int s2 = gw; s2 < g.R; s2 += NWG) tile[s2 * GM_LDT + lane] = s2 < S ? cp[s2] : 0.0;
    } else {
      const double* sp = src + srcOff + i0;
#pragma unroll
      for (int j = 0; j < RR / (2 * NWG); ++j)
        if (row0 + j * 2 * NWG < S)
          cp_async16(tile + dstOff + j * 2 * NWG * GM_LDT, sp + j * srcStep);
    }
  };
  // The two groups take turns in the tensor cores: a group enters its Q + G phases when the other
  // one has left its own (token through the named barriers BAR_GO + group), so that one group's
  // tile wait and element-wise U phase always fall into the other's DMMA phases.  Left to
  // themselves -- even started half a period apart -- the groups drift into lock step within two
  // trips (both wait for tiles, then both share the pipe: 56 % busy, profiles/r02_codon.md).
  int trip = 0;
  for (int i0 = begin + grp * GM_TP; i0 < end; i0 += 2 * GM_TP, ++trip) {
    GM_TRACE(0)
    // (the group finished reading its buffers before the barrier at the end of the last trip)
    double2 qf[NTG];
    if (US) {
      const int row = gw * 8 + (lane >> 2);
      const double* qp = qsrc + (size_t)row * Npad + i0 + (lane & 3) * 2;
#pragma unroll
      for (int n = 0; n < NTG; ++n)
        qf[n] = row < S ? __ldg(reinterpret_cast<const double2*>(qp + n * 8)) : make_double2(0.0, 0.0);
      if (!tipR) stage_tile(tq, false, nullptr, ursrc, i0);   // u_r, becomes m_l
      if (!tipL) stage_tile(mr, false, nullptr, ulsrc, i0);   // u_l, becomes m_r
    } else {
      stage_tile(tq, false, nullptr, qsrc, i0);
    }
    if (!hotL) stage_tile(vl, tipL, tl, lsrc, i0);
    if (!hotR) stage_tile(vr, tipR, tr, rsrc, i0);
    if (gw == 0) {
      cp_async8(ws + lane, weights + i0 + lane);
      if (lane < 16) {
        if (!tipL) cp_async4(se + 2 * lane, elp + i0 + 2 * lane);
        if (!tipR) cp_async4(se + 32 + 2 * lane, erp + i0 + 2 * lane);
      } else if (lane < 24) {
        if (tabL) cp_async4(codesS + 4 * (lane - 16), tl + i0 + 4 * (lane - 16));
      } else {
        if (tabR) cp_async4(codesS + 32 + 4 * (lane - 24), tr + i0 + 4 * (lane - 24));
      }
    }
    cp_async_commit();
    cp_async_wait_all();
    named_sync(BAR_GROUP + grp, NWG * 32);   // the group's tiles are visible
    GM_TRACE(1)
    // U phase: u_l = P_l v_l, u_r = P_r v_r;  m_l = q^ o u_r (in place of q^), m_r = q^ o u_l
    for (int mt = gw; mt < MT; mt += NWG) {
      double accL[NTG][2], accR[NTG][2];
#pragma unroll
      for (int n = 0; n < NTG; ++n) accL[n][0] = accL[n][1] = accR[n][0] = accR[n][1] = 0.0;
      if (US) {
        const int off0 = (mt * 8 + (lane >> 2)) * GM_LDT + (lane & 3) * 2;
        if (tabL) gm_u_tip<NTG>(accL, Pl, uld, codesS, mt, 0, lane);
        else if (tipL) gm_mma_ab_g<NTG>(accL, Pl, g.PLD, vl, mt, 0, KT, lane);
        else {
#pragma unroll
          for (int n = 0; n < NTG; ++n) {
            const double2 u = *reinterpret_cast<const double2*>(mr + off0 + n * 8);
            accL[n][0] = u.x; accL[n][1] = u.y;
          }
        }
        if (tabR) gm_u_tip<NTG>(accR, Pr, uld, codesS + 32, mt, 0, lane);
        else if (tipR) gm_mma_ab_g<NTG>(accR, Pr, g.PLD, vr, mt, 0, KT, lane);
        else {
#pragma unroll
          for (int n = 0; n < NTG; ++n) {
            const double2 u = *reinterpret_cast<const double2*>(tq + off0 + n * 8);
            accR[n][0] = u.x; accR[n][1] = u.y;
          }
        }
      } else if (!tabL && !tabR) {
        gm_mma_ab_g2<NTG>(accL, accR, Pl, Pr, g.PLD, vl, vr, mt, 0, KT, lane);
      } else {
        if (tabL) gm_u_tip<NTG>(accL, Pl, uld, codesS, mt, 0, lane);
        else gm_mma_ab_g<NTG>(accL, Pl, g.PLD, vl, mt, 0, KT, lane);
        if (tabR) gm_u_tip<NTG>(accR, Pr, uld, codesS + 32, mt, 0, lane);
        else gm_mma_ab_g<NTG>(accR, Pr, g.PLD, vr, mt, 0, KT, lane);
      }
#pragma unroll
      for (int n = 0; n < NTG; ++n) {
        const int off = (mt * 8 + (lane >> 2)) * GM_LDT + n * 8 + (lane & 3) * 2;
        const double q0 = US ? qf[n].x : tq[off], q1 = US ? qf[n].y : tq[off + 1];
        tq[off] = q0 * accR[n][0];
        tq[off + 1] = q1 * accR[n][1];
        mr[off] = q0 * accL[n][0];
        mr[off + 1] = q1 * accL[n][1];
      }
    }
    named_sync(BAR_GROUP + grp, NWG * 32);
    GM_TRACE(2)
    if (grp == 1 || trip > 0) named_sync(BAR_GO + grp, 512);   // the other group left the pipe
    GM_TRACE(3)
    const double* ml = tq;
    // Q phase: q^_c = P_c^T m_c * 2^{-e_c}  (internal children)
    for (int side = 0; side < 2; ++side) {
      if (side ? tipR : tipL) continue;
      const double* P = side ? Pr : Pl;
      const double* mm = side ? mr : ml;
      const int child = side ? op.right : op.left;
      double* qout = pre + drawBase + (size_t)(child - T) * nodeStride + k * plane + i0;
      for (int mt = gw; mt < MT; mt += NWG) {
        double c[NTG][2];
#pragma unroll
        for (int n = 0; n < NTG; ++n) c[n][0] = c[n][1] = 0.0;
        gm_mma_atb_pf<NTG, (CS + 7) / 8 * 2>(c, P, g.PLD, mm, mt, lane);
        const int row = mt * 8 + (lane >> 2);
        if (row < S) {
#pragma unroll
          for (int n = 0; n < NTG; ++n) {
            const int col = n * 8 + (lane & 3) * 2;
            const double f0 = __hiloint2double((1023 - (int)se[side * 32 + col]) << 20, 0);
            const double f1 = __hiloint2double((1023 - (int)se[side * 32 + col + 1]) << 20, 0);
            *reinterpret_cast<double2*>(qout + (size_t)row * Npad + col) =
                make_double2(c[n][0] * f0, c[n][1] * f1);
          }
        }
      }
    }
    // pass the token if the other group has a trip left that waits for it (group 1's trip
    // `trip` after group 0's, group 0's trip `trip + 1` after group 1's: both start at i0 + 32)
    if (tokenAt == 0 && i0 + GM_TP < end) named_arrive(BAR_GO + (grp ^ 1), 512);
    GM_TRACE(4)
    // G phase: G_c[s][t] += sum_p (w_p m_c[s][p]) v_c[t][p]
#pragma unroll
    for (int cj = 0; cj < NCJ; ++cj) {
      if (cj == 1 && tokenAt == 1 && i0 + GM_TP < end) named_arrive(BAR_GO + (grp ^ 1), 512);
      const int combo = gw + cj * NWG;
      if (combo < 2 * MT) {
        const int side = combo / MT;
        const int mt = combo - side * MT;
        const double* mm = (side ? mr : ml) + (mt * 8 + (lane >> 2)) * GM_LDT + (lane & 3);
        const double* wp = ws + (lane & 3);
        if (side ? hotR : hotL) {
          const uint8_t* cds = codesS + side * 32 + (lane & 3);
          const int colBase = lane >> 2;
#pragma unroll
          for (int kt = 0; kt < GM_TP / 4; ++kt) {
            const double av = mm[kt * 4] * wp[kt * 4];
            const int code = cds[kt * 4];
#pragma unroll
            for (int mt2 = 0; mt2 < 8; ++mt2)
              if (mt2 < MT) {
                const int col = mt2 * 8 + colBase;
                const double bv = (col < S && (code == col || code == S)) ? 1.0 : 0.0;
                dmma884(acc[cj * 8 + mt2][0], acc[cj * 8 + mt2][1], av, bv);
              }
          }
        } else {
          const double* vv = (side ? vr : vl) + (lane >> 2) * GM_LDT + (lane & 3);
          // operands four MMAs ahead (see gm_mma_atb_pf)
          constexpr int NJ = (GM_TP / 4) * 8, AHEAD = 4;
          double bq[AHEAD], av = mm[0] * wp[0];
#pragma unroll
          for (int j = 0; j < AHEAD; ++j) bq[j] = vv[(j & 7) * 8 * GM_LDT + (j >> 3) * 4];
#pragma unroll
          for (int kt = 0; kt < GM_TP / 4; ++kt) {
            double avn = 0.0;
            if (kt + 1 < GM_TP / 4) avn = mm[(kt + 1) * 4] * wp[(kt + 1) * 4];
#pragma unroll
            for (int mt2 = 0; mt2 < 8; ++mt2) {
              const int j = kt * 8 + mt2;
              const double bcur = bq[j % AHEAD];
              if (j + AHEAD < NJ) {
                const int j2 = j + AHEAD;
                bq[j % AHEAD] = vv[(j2 & 7) * 8 * GM_LDT + (j2 >> 3) * 4];
              }
              if (mt2 < MT) dmma884(acc[cj * 8 + mt2][0], acc[cj * 8 + mt2][1], av, bcur);
            }
            av = avn;
          }
        }
      }
    }
    if (tokenAt == 2 && i0 + GM_TP < end) named_arrive(BAR_GO + (grp ^ 1), 512);
    named_sync(BAR_GROUP + grp, NWG * 32);   // every warp of the group is done with the tiles
    GM_TRACE(5)
  }
  // partial G of group 1 -> shared memory (its own tile area: 4 tiles hold 2 * 64 * 64 doubles
  // only if LDT >= 32: 4 * 64 * 36 = 9216 >= 8192), then group 0 adds and stores
  __syncthreads();
  double* stash = Pr + g.Sp * g.PLD + groupN;   // group 1's area
  if (grp == 1) {
#pragma unroll
    for (int j = 0; j < NCJ * 8; ++j) {
      stash[((gw * NCJ * 8 + j) * 32 + lane) * 2] = acc[j][0];
      stash[((gw * NCJ * 8 + j) * 32 + lane) * 2 + 1] = acc[j][1];
    }
  }
  __syncthreads();
  if (grp == 0) {
#pragma unroll
    for (int cj = 0; cj < NCJ; ++cj) {
      const int combo = gw + cj * NWG;
      if (combo < 2 * MT) {
        const int side = combo / MT;
        const int mt = combo - side * MT;
        const int branch = side ? op.right : op.left;
        double* o = gpart + ((size_t)d * chunkTotal + chunkBase[branch] + (size_t)k * nChunk +
                             blockIdx.x) * SS;
        const int row = mt * 8 + (lane >> 2);
#pragma unroll
        for (int mt2 = 0; mt2 < 8; ++mt2) {
          if (mt2 < MT && row < S) {
            const int j = cj * 8 + mt2;
            const double g0 = acc[j][0] + stash[((gw * NCJ * 8 + j) * 32 + lane) * 2];
            const double g1 = acc[j][1] + stash[((gw * NCJ * 8 + j) * 32 + lane) * 2 + 1];
            const int col = mt2 * 8 + (lane & 3) * 2;
            if (col < S) o[row * S + col] = g0;
            if (col + 1 < S) o[row * S + col + 1] = g1;
          }
        }
      }
    }
  }
}

size_t gm_bwd3_smem(int S) {
  const GmShape g = gm_shape(S);
  return (2 * (size_t)g.Sp * g.PLD + 2 * (4 * (size_t)g.R * GM_LDT + 56)) * sizeof(double);
}

// ---------------------------------------------------------------------------
// pre-order, level 1 (both children are tips): the children receive no q^ and their u vectors
// are table rows, so the only GEMM left is G_c += (w o m_c) . onehot(code_c)^T.  The general
// kernel still stages two S x S matrices and runs its U phase, barrier, ml / mr round trip
// through shared memory (one CTA per SM, DMMA pipe 54 % busy on this level,
// profiles/r02_codon_ncu.md).  Here the A operand w_p q^[s][p] u_sib[code_sib(p)][s] is formed
// in registers straight from the q^ tile and the sibling's table, the one-hot B operand from the
// child's own code: one phase per tile, no ml / mr tiles, 102 KB of shared memory -> two CTAs
// per SM whose barriers and tile waits overlap.
// Needs unit-vector / gap codes only (codeCount == S + 1).
// shared: utL utR [S+1][Sp+1] | ring[2]: tq [Sp][LDT], w [32], codes [2][32]
// ---------------------------------------------------------------------------
template <int CS, int NW>
__global__ void __launch_bounds__(NW * 32, 2)
gm_cherry_bwd_kernel(const NodeOp* __restrict__ ops, int opBegin, const double* __restrict__ mats,
                     const uint8_t* __restrict__ tips, const double* __restrict__ codeP,
                     const double* __restrict__ weights, const double* __restrict__ pre,
                     double* __restrict__ gpart, const int* __restrict__ chunkBase,
                     size_t chunkTotal, int T, int Npad, int B, int K, int chunkPatterns,
                     int nChunk) {
  extern __shared__ double sm[];
  constexpr int S = CS;
  const GmShape g = gm_shape(S);
  constexpr int SS = S * S;
  const int uld = g.Sp + 1;
  const int tileN = g.Sp * GM_LDT;
  const int slotN = tileN + 32 + 8;          // tile | weights | 64 code bytes
  double* utL = sm;
  double* utR = utL + (S + 1) * uld;
  double* ring = utR + (S + 1) * uld;
  const int nodeSlot = blockIdx.y / K;
  const int k = blockIdx.y - nodeSlot * K;
  const NodeOp op = ops[opBegin + nodeSlot];
  const int d = blockIdx.z;
  const int I = T - 1;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const size_t plane = (size_t)S * Npad;
  const size_t nodeStride = (size_t)K * plane;
  const double* matsD = mats + (size_t)d * B * K * SS;
  gm_stage_utab(utL, matsD + ((size_t)op.left * K + k) * SS, codeP, S + 1, g);
  gm_stage_utab(utR, matsD + ((size_t)op.right * K + k) * SS, codeP, S + 1, g);
  // rows S .. Sp-1 of the q^ tiles are never copied: keep them zero
  for (int j = threadIdx.x; j < 2 * (g.Sp - S) * GM_LDT; j += blockDim.x) {
    const int b = j / ((g.Sp - S) * GM_LDT), r = j - b * (g.Sp - S) * GM_LDT;
    ring[b * slotN + S * GM_LDT + r] = 0.0;
  }
  pdl_wait_then_trigger();

  constexpr int MT = (S + 7) / 8;
  constexpr int NCJ = (2 * MT + NW - 1) / NW;
  double acc[NCJ * 8][2];
#pragma unroll
  for (int j = 0; j < NCJ * 8; ++j) acc[j][0] = acc[j][1] = 0.0;

  const int begin = blockIdx.x * chunkPatterns;
  int end = begin + chunkPatterns;
  end = end < Npad ? end : Npad;
  const double* qsrc = pre + (size_t)d * I * nodeStride + (size_t)(op.node - T) * nodeStride + k * plane;
  const uint8_t* tl = tips + (size_t)op.left * Npad;
  const uint8_t* tr = tips + (size_t)op.right * Npad;
  auto stage = [&](int b, int i0) {
    double* slot = ring + b * slotN;
    for (int s2 = warp; s2 < S; s2 += NW)
      cp_async8(slot + s2 * GM_LDT + lane, qsrc + (size_t)s2 * Npad + i0 + lane);
    if (warp == 0) cp_async8(slot + tileN + lane, weights + i0 + lane);
    if (warp == 1 && lane < 16) {
      uint8_t* cd = reinterpret_cast<uint8_t*>(slot + tileN + 32);
      if (lane < 8) cp_async4(cd + 4 * lane, tl + i0 + 4 * lane);
      else cp_async4(cd + 32 + 4 * (lane - 8), tr + i0 + 4 * (lane - 8));
    }
  };
  if (begin < end) stage(0, begin);
  cp_async_commit();
  int buf = 0;
  for (int i0 = begin; i0 < end; i0 += GM_TP, buf ^= 1) {
    cp_async_wait_all();
    __syncthreads();   // slot `buf` (and the tables on the first trip) visible; slot `buf^1` free
    if (i0 + GM_TP < end) stage(buf ^ 1, i0 + GM_TP);
    cp_async_commit();
    const double* slot = ring + buf * slotN;
    const double* ws = slot + tileN + (lane & 3);
    const uint8_t* cd = reinterpret_cast<const uint8_t*>(slot + tileN + 32) + (lane & 3);
#pragma unroll
    for (int cj = 0; cj < NCJ; ++cj) {
      const int combo = warp + cj * NW;
      if (combo < 2 * MT) {
        const int side = combo / MT;
        const int mt = combo - side * MT;
        const int row = mt * 8 + (lane >> 2);
        const double* tq = slot + row * GM_LDT + (lane & 3);
        const double* utSib = (side ? utL : utR) + row;   // m_side = q^ o u_sibling
        const uint8_t* own = cd + side * 32;
        const uint8_t* sib = cd + (1 - side) * 32;
        const int colBase = lane >> 2;
#pragma unroll
        for (int kt = 0; kt < GM_TP / 4; ++kt) {
          const double av = ws[kt * 4] * tq[kt * 4] * utSib[(int)sib[kt * 4] * uld];
          const int code = own[kt * 4];
#pragma unroll
          for (int mt2 = 0; mt2 < MT; ++mt2) {
            const int col = mt2 * 8 + colBase;
            const double bv = (col < S && (code == col || code == S)) ? 1.0 : 0.0;
            dmma884(acc[cj * 8 + mt2][0], acc[cj * 8 + mt2][1], av, bv);
          }
        }
      }
    }
  }
#pragma unroll
  for (int cj = 0; cj < NCJ; ++cj) {
    const int combo = warp + cj * NW;
    if (combo < 2 * MT) {
      const int side = combo / MT;
      const int mt = combo - side * MT;
      const int branch = side ? op.right : op.left;
      double* o = gpart + ((size_t)d * chunkTotal + chunkBase[branch] + (size_t)k * nChunk +
                           blockIdx.x) * SS;
      const int row = mt * 8 + (lane >> 2);
#pragma unroll
      for (int mt2 = 0; mt2 < MT; ++mt2) {
        if (row < S) {
          const int col = mt2 * 8 + (lane & 3) * 2;
          if (col < S) o[row * S + col] = acc[cj * 8 + mt2][0];
          if (col + 1 < S) o[row * S + col + 1] = acc[cj * 8 + mt2][1];
        }
      }
    }
  }
}

size_t gm_cherry_bwd_smem(int S) {
  const GmShape g = gm_shape(S);
  return (2 * (size_t)(S + 1) * (g.Sp + 1) + 2 * ((size_t)g.Sp * GM_LDT + 40)) * sizeof(double);
}

size_t gm_fwd2_smem(const Dims& m) {
  const GmShape g = gm_shape(m.S);
  return (2 * (size_t)g.Sp * g.PLD + 5 * (size_t)g.R * GM_LDT + 16 * 32) * sizeof(double) + 128;
}

size_t gm_fwd3_smem(const Dims& m) {
  const GmShape g = gm_shape(m.S);
  return (2 * (size_t)g.Sp * g.PLD + 6 * (size_t)g.R * GM_LDT + 2 * 8 * 32) * sizeof(double) + 128;
}

size_t gm_bwd2_smem(const Dims& m) {
  const GmShape g = gm_shape(m.S);
  return (2 * (size_t)g.Sp * g.PLD + 8 * (size_t)g.R * GM_LDT + 64) * sizeof(double) +
         128 * sizeof(int) + 128;
}

}  // namespace

bool gmma_supported(const Engine& e) {
  const Dims& m = e.dm;
  return !e.spec4 && m.S >= 8 && m.S <= 64 && !(e.cfg.flags & TTB2_FLAG_NO_MMA) &&
         !(e.cfg.flags & TTB2_FLAG_FORCE_GENERIC) && m.K <= GM_MAXK &&
         gm_fwd2_smem(m) <= 227 * 1024 && gm_bwd2_smem(m) <= 227 * 1024;
}

// 61 states through gm_fwd3_kernel / gm_bwd3_kernel (TTB2_GM_NO_USTORE=1: always recompute)
bool gmma_keeps_u(const Engine& e) {
  return gmma_supported(e) && e.dm.S == 61 && gm_fwd3_smem(e.dm) <= 227 * 1024 &&
         !getenv("TTB2_GM_NO_USTORE");
}

size_t gmma_expo_elems(const Engine& e) {
  const Dims& m = e.dm;
  return (size_t)e.cfg.max_draws * m.I * m.K * m.Npad;
}

bool gmma_cherry_level_supported(const Engine& e) {
  const size_t smem = 2 * (size_t)e.cfg.code_count * (e.dm.S | 1) * sizeof(double);
  return smem <= 200 * 1024 && !getenv("TTB2_GM_NO_CHERRY");
}

// post-order level 1 (nodes whose two children are tips) for any 8 <= S <= 64
int gmma_cherry_forward_level(Engine& e, int draws) {
  const Dims& m = e.dm;
  const size_t cherrySmem = 2 * (size_t)e.cfg.code_count * (m.S | 1) * sizeof(double);
  auto ck = m.S == 61 ? gm_cherry_fwd_kernel<61>
            : m.S == 20 ? gm_cherry_fwd_kernel<20> : gm_cherry_fwd_kernel<0>;
  if (cherrySmem > 48 * 1024)
    TTB2_CUDA_CHECK(cudaFuncSetAttribute(ck, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)cherrySmem));
  const int opBegin = e.levelOff[0];
  const int count = e.levelOff[1] - opBegin;
  // ~8 blocks per SM: long enough to amortise the table build
  long blocks = ((long)e.smCount * 8 + (long)count * m.K * draws - 1) / ((long)count * m.K * draws);
  if (blocks < 1) blocks = 1;
  int per = (int)((m.Npad + blocks - 1) / blocks);
  per = (per + GMC_THREADS - 1) / GMC_THREADS * GMC_THREADS;
  const int nBlock = (m.Npad + per - 1) / per;
  const int maxNodes = 65535 / m.K;
  for (int done = 0; done < count; done += maxNodes) {
    const int c = (count - done) < maxNodes ? (count - done) : maxNodes;
    dim3 grid(nBlock, c * m.K, draws);
    launch_level(ck, grid, GMC_THREADS, cherrySmem, e.stream, false, e.ops, opBegin + done,
                 e.mats, e.tips, e.codeP, e.partials, e.expoK, m.T, m.Npad, m.B, m.K, m.S, per,
                 e.cfg.code_count);
    ++e.launches;
  }
  TTB2_CUDA_CHECK(cudaGetLastError());
  return TTB2_OK;
}

// pre-order level 1 (both children tips, unit / gap codes) for the compiled alphabets
bool gmma_cherry_backward_supported(const Engine& e) {
  return (e.dm.S == 61 || e.dm.S == 20) && e.cfg.code_count == e.dm.S + 1 &&
         !getenv("TTB2_GM_NO_CHERRY");
}

int gmma_cherry_backward_level(Engine& e, int draws, bool pdl) {
  const Dims& m = e.dm;
  const int opBegin = e.levelOff[0];
  const int count = e.levelOff[1] - opBegin;
  const int nChunk = e.levelChunks[0];
  int chunkPatterns = (m.Npad + nChunk - 1) / nChunk;
  chunkPatterns = (chunkPatterns + GM_TP - 1) / GM_TP * GM_TP;
  const int maxNodes = 65535 / m.K;
  const size_t csm = gm_cherry_bwd_smem(m.S);
  auto kern = m.S == 61 ? gm_cherry_bwd_kernel<61, 8> : gm_cherry_bwd_kernel<20, 6>;
  const int threads = m.S == 61 ? 256 : 192;
  if (csm > 48 * 1024)
    TTB2_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)csm));
  for (int done = 0; done < count; done += maxNodes) {
    const int c = (count - done) < maxNodes ? (count - done) : maxNodes;
    dim3 grid(nChunk, c * m.K, draws);
    launch_level(kern, grid, threads, csm, e.stream, pdl, e.ops, opBegin + done, e.mats, e.tips,
                 e.codeP, e.weights, e.pre, e.gpart, e.chunkBase, e.chunkTotal, m.T, m.Npad, m.B,
                 m.K, chunkPatterns, nChunk);
    ++e.launches;
  }
  TTB2_CUDA_CHECK(cudaGetLastError());
  return TTB2_OK;
}

// chunk of patterns per post-order CTA (enough CTAs to fill the GPU, long-lived otherwise)
static int gm_fwd_chunk(const Engine& e, int draws, int count) {
  const Dims& m = e.dm;
  // staging the two S x S matrices is the per-CTA fixed cost: many short CTAs for small
  // alphabets (config 4, S=20: 32/SM best), few long ones for large (config 5, S=61: 8/SM)
  const long target = (long)e.smCount * (m.S <= 32 ? 32 : 8);
  long chunks = (target + (long)count * m.K * draws - 1) / ((long)count * m.K * draws);
  // large alphabets: at least 8 tiles per CTA (staging the two matrices costs about one tile)
  const long maxChunks = m.S > 32 ? std::max(1, m.Npad / (8 * GM_TP)) : m.Npad / GM_TP;
  // large alphabets run one CTA per SM: launch whole waves (engine.cuh wave_aware_chunks)
  if (m.S > 32)
    chunks = wave_aware_chunks((long)count * m.K * draws, chunks, maxChunks, e.smCount);
  if (chunks > maxChunks) chunks = maxChunks;
  if (chunks < 1) chunks = 1;
  int chunkPatterns = (int)((m.Npad + chunks - 1) / chunks);
  return (chunkPatterns + GM_TP - 1) / GM_TP * GM_TP;
}

int gmma_forward2(Engine& e, int draws) {
  e.uValid = false;
  if (gwarp_supported(e, false)) return gwarp_forward(e, draws);
  const Dims& m = e.dm;
  const size_t smem = gm_fwd2_smem(m);
  const int MT = gm_shape(m.S).Sp / 8;
  const int ntg = MT >= 8 ? 4 : (MT >= 4 ? 2 : 1);
  int nw = MT >= 5 ? 8 : 4;
  // 20 states: 8 warps x 2 pattern tiles per A fragment measured best on config 4
  // (<1,4>: 26.6 ms, <1,8>: 22.5, <2,4>: 26.8, <2,8>: 21.6 per evaluation)
  if (m.S == 20) nw = 8;
  // 61 states: the warp-specialised kernel (two groups of 8 DMMA warps + 8 epilogue warps)
  const bool spec61 = m.S == 61 && gm_fwd3_smem(m) <= 227 * 1024;
  // level 1 (tip-tip nodes) as a streaming kernel when its two code tables fit in shared memory
  const bool cherryLevel = gmma_cherry_level_supported(e);
  auto launch_cherry_level = [&]() -> int { return gmma_cherry_forward_level(e, draws); };
  if (spec61) {
    static const int tokenF = getenv("TTB2_GM_TOKENF") ? atoi(getenv("TTB2_GM_TOKENF")) : 1;
    const size_t smem3 = gm_fwd3_smem(m);
    TTB2_CUDA_CHECK(cudaFuncSetAttribute(gm_fwd3_kernel<61>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem3));
    const int nLevels = (int)e.levelOff.size() - 1;
    const int maxNodes = 65535 / m.K;
    if (cherryLevel) {
      const int rc = launch_cherry_level();
      if (rc) return rc;
    }
    for (int l = cherryLevel ? 1 : 0; l < nLevels; ++l) {
      const int opBegin = e.levelOff[l];
      const int count = e.levelOff[l + 1] - opBegin;
      const int chunkPatterns = gm_fwd_chunk(e, draws, count);
      const int nChunk = (m.Npad + chunkPatterns - 1) / chunkPatterns;
      for (int done = 0; done < count; done += maxNodes) {
        const int c = (count - done) < maxNodes ? (count - done) : maxNodes;
        dim3 grid(nChunk, c * m.K, draws);
        launch_level(gm_fwd3_kernel<61>, grid, 768, smem3, e.stream, l > 0 && pdl_enabled(), e.ops,
                     opBegin + done, e.mats, e.tips, e.codeP, e.partials, e.expoK, e.ustore, m.T,
                     m.Npad, m.B, m.K, m.S, chunkPatterns, e.cfg.code_count, tokenF);
        ++e.launches;
      }
    }
    TTB2_CUDA_CHECK(cudaGetLastError());
    e.uValid = e.ustore != nullptr;
    return TTB2_OK;
  }
  auto kern = m.S == 20 ? gm_fwd2_kernel<2, 8, 20>
              : ntg == 4  ? gm_fwd2_kernel<4, 8, 0>
              : nw == 8   ? gm_fwd2_kernel<2, 8, 0>
              : ntg == 2  ? gm_fwd2_kernel<2, 4, 0>
                          : gm_fwd2_kernel<1, 4, 0>;
  if (smem > 48 * 1024)
    TTB2_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)smem));
  const int nLevels = (int)e.levelOff.size() - 1;
  const int maxNodes = 65535 / m.K;
  if (cherryLevel) {
    const int rc = launch_cherry_level();
    if (rc) return rc;
  }
  for (int l = cherryLevel ? 1 : 0; l < nLevels; ++l) {
    const int opBegin = e.levelOff[l];
    const int count = e.levelOff[l + 1] - opBegin;
    const int chunkPatterns = gm_fwd_chunk(e, draws, count);
    const int nChunk = (m.Npad + chunkPatterns - 1) / chunkPatterns;
    for (int done = 0; done < count; done += maxNodes) {
      const int c = (count - done) < maxNodes ? (count - done) : maxNodes;
      dim3 grid(nChunk, c * m.K, draws);
      launch_level(kern, grid, nw * 32, smem, e.stream, l > 0 && pdl_enabled(), e.ops,
                   opBegin + done, e.mats, e.tips, e.codeP, e.partials, e.expoK, m.T, m.Npad, m.B,
                   m.K, m.S, chunkPatterns, e.cfg.code_count);
      ++e.launches;
    }
  }
  TTB2_CUDA_CHECK(cudaGetLastError());
  return TTB2_OK;
}

int gmma_root2(Engine& e, int draws) {
  const Dims& m = e.dm;
  const int nblocks = m.Npad / 32;
  dim3 grid(nblocks, draws);
  const int rootInode = e.hostOps.back().node - m.T;
  gm_root2_kernel<false><<<grid, dim3(32, m.K), 0, e.stream>>>(
      e.partials, e.expoK, e.freqs, e.freqDraws, e.props, e.propDraws, e.weights, e.siteLnl,
      nullptr, e.redPart, m.T, m.Npad, m.K, m.S, rootInode);
  ++e.launches;
  TTB2_CUDA_CHECK(cudaGetLastError());
  return small_reduce_lnl(e, draws, nblocks);
}

int gmma_backward2(Engine& e, int draws) {
  const Dims& m = e.dm;
  const int rootInode = e.hostOps.back().node - m.T;
  {
    const int nblocks = m.Npad / 32;
    dim3 grid(nblocks, draws);
    gm_root2_kernel<true><<<grid, dim3(32, m.K), 0, e.stream>>>(
        e.partials, e.expoK, e.freqs, e.freqDraws, e.props, e.propDraws, e.weights, e.siteLnl,
        e.pre, e.redPart, m.T, m.Npad, m.K, m.S, rootInode);
    ++e.launches;
    TTB2_CUDA_CHECK(cudaGetLastError());
    int rc = small_root_grad_reduce(e, draws, nblocks);
    if (rc) return rc;
  }
  if (gwarp_supported(e, true)) {
    const int rc = gwarp_backward_levels(e, draws);
    return rc ? rc : small_gpart_reduce(e, draws);
  }
  const size_t smem = gm_bwd2_smem(m);
  const int MT = gm_shape(m.S).Sp / 8;
  const int ntg = MT >= 8 ? 4 : (MT >= 4 ? 2 : 1);
  int nw = MT >= 5 ? 8 : 4;
  // 20 states: 8 warps x 2 pattern tiles per A fragment measured best on config 4
  // (<1,4>: 26.6 ms, <1,8>: 22.5, <2,4>: 26.8, <2,8>: 21.6 per evaluation)
  if (m.S == 20) nw = 8;
  // (61 states run gm_bwd3_kernel / gm_cherry_bwd_kernel below; this instance only serves the
  // other alphabets)
  auto kern = m.S == 20 ? gm_bwd2_kernel<2, 8, 20>
              : ntg == 4  ? gm_bwd2_kernel<4, 8, 0>
              : nw == 8   ? gm_bwd2_kernel<2, 8, 0>
              : ntg == 2  ? gm_bwd2_kernel<2, 4, 0>
                          : gm_bwd2_kernel<1, 4, 0>;
  if (smem > 48 * 1024)
    TTB2_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)smem));
  const int nLevels = (int)e.levelOff.size() - 1;
  const int maxNodes = 65535 / m.K;
  for (int l = nLevels - 1; l >= 0; --l) {
    const int opBegin = e.levelOff[l];
    const int count = e.levelOff[l + 1] - opBegin;
    const int nChunk = e.levelChunks[l];
    int chunkPatterns = (m.Npad + nChunk - 1) / nChunk;
    chunkPatterns = (chunkPatterns + GM_TP - 1) / GM_TP * GM_TP;
    // level 1 of the codon path: tip-tip nodes through the dedicated kernel (unit / gap codes)
    if (l == 0 && gmma_cherry_backward_supported(e)) {
      const int rc = gmma_cherry_backward_level(e, draws, l < nLevels - 1 && pdl_enabled());
      if (rc) return rc;
      continue;
    }
    if (m.S == 61) {   // two-group kernel
      const size_t sm3 = gm_bwd3_smem(m.S);
      // with the u vectors of this evaluation's post-order sweep when it kept them
      auto k3 = e.uValid ? gm_bwd3_kernel<61, true> : gm_bwd3_kernel<61, false>;
      static const int tokenAt = getenv("TTB2_GM_TOKEN") ? atoi(getenv("TTB2_GM_TOKEN")) : 1;
      TTB2_CUDA_CHECK(cudaFuncSetAttribute(k3, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)sm3));
      for (int done = 0; done < count; done += maxNodes) {
        const int c = (count - done) < maxNodes ? (count - done) : maxNodes;
        dim3 grid(nChunk, c * m.K, draws);
        launch_level(k3, grid, 512, sm3, e.stream,
                     l < nLevels - 1 && pdl_enabled(), e.ops, opBegin + done, e.mats, e.tips,
                     e.codeP, e.partials, e.expoK, e.ustore, e.weights, e.pre, e.gpart, e.chunkBase,
                     e.chunkTotal, m.T, m.Npad, m.B, m.K, chunkPatterns, nChunk, e.cfg.code_count,
                     tokenAt);
        ++e.launches;
      }
      continue;
    }
    for (int done = 0; done < count; done += maxNodes) {
      const int c = (count - done) < maxNodes ? (count - done) : maxNodes;
      dim3 grid(nChunk, c * m.K, draws);
      launch_level(kern, grid, nw * 32, smem, e.stream, l < nLevels - 1 && pdl_enabled(), e.ops,
                   opBegin + done, e.mats, e.tips, e.codeP, e.partials, e.expoK, e.weights, e.pre,
                   e.gpart, e.chunkBase, e.chunkTotal, m.T, m.Npad, m.B, m.K, m.S, chunkPatterns,
                   nChunk, e.cfg.code_count);
      ++e.launches;
    }
  }
  TTB2_CUDA_CHECK(cudaGetLastError());
  return small_gpart_reduce(e, draws);
}

}  // namespace ttb2

#ifdef TTB2_GM_TRACE
extern "C" int ttb2_debug_gm_trace(long long* out) {
  return (int)cudaMemcpyFromSymbol(out, ttb2::gm_trace, sizeof(long long) * 2 * 64 * 6);
}
#endif
