// Warp-autonomous fp64 tensor-core (DMMA) level kernels for the 20-state amino-acid path.
//
// kernels_gmma.cu stages 32-pattern tiles in shared memory and synchronises the whole CTA
// three times per tile (config 4 ran at 0.25 of its HBM roofline with the DMMA pipe 27-39 %
// busy, profiles/r01_generic_states_ncu.md).  Here a warp owns 8 patterns at a time and never
// meets another warp inside the pattern loop.  The GEMMs are issued TRANSPOSED, patterns on
// the M axis of mma.sync.m8n8k4.f64:
//
//   U_c^T [8 x S]  = V_c^T [8 x S] . P_c^T [S x S]       A = child vectors from the warp's own
//                                                         cp.async ring (lane = (pattern r,
//                                                         state c)), B = constant fragments
//                                                         of P_c^T
//   m_c^T          = q^_n^T o U_sibling^T                lane-local: C fragments line up
//   q^_c^T [8 x S] = m_c^T [8 x S] . P_c [S x S]         the C fragment of m^T IS the A operand:
//                                                         the reduction index of a k-step is
//                                                         permuted to (8 nt + 2 c + e), the two
//                                                         values lane (r, c) already holds, and
//                                                         the constant B fragments of P_c are
//                                                         permuted to match (the trick of the
//                                                         4-state bwd4_tma_kernel)
//   G_c [S x S]   += (w o m_c) [S x 8] . V_c^T [8 x S]   the one product that contracts over
//                                                         patterns: m^T is transposed inside
//                                                         the warp with shuffles, V_c^T is
//                                                         read from the ring in B-fragment
//                                                         layout
//
// The constant fragments live in shared memory as [fragment][lane] (one conflict-free
// LDS.64 per DMMA), written once per CTA before the only __syncthreads of the kernel.
// HBM layout, exponents (per pattern and category, expoK) and the root kernel are those of
// kernels_gmma.cu.  Outputs are identical in meaning; bit patterns differ from the tile
// kernels only through the order of the fp64 sums.
#include "engine.cuh"

namespace ttb2 {

namespace {

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  // not volatile: a pure function of its operands, so the compiler may interleave the
  // independent accumulator chains of different phases
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

template <int S>
struct GwShape {
  static constexpr int KT = (S + 3) / 4;   // k-steps over child states
  static constexpr int NT = (S + 7) / 8;   // 8-wide tiles over states
  static constexpr int UF = 2 * KT * NT;   // U fragments (both children)
  static constexpr int QF = 2 * NT * 2 * NT;  // Q fragments (both children)
};

// fragments of P_c^T for U^T = V^T P^T:  B[k = child state 4 kt + c][n = parent state 8 nt + r]
template <int S>
__device__ __forceinline__ void gw_stage_u(double* dst, const double* Pl, const double* Pr,
                                           int nthreads) {
  using G = GwShape<S>;
  for (int idx = threadIdx.x; idx < G::UF * 32; idx += nthreads) {
    const int lane = idx & 31, f = idx >> 5;
    const int nt = f % G::NT, kt = (f / G::NT) % G::KT, side = f / (G::NT * G::KT);
    const int child = 4 * kt + (lane & 3), parent = 8 * nt + (lane >> 2);
    const double* P = side ? Pr : Pl;
    dst[idx] = (child < S && parent < S) ? P[parent * S + child] : 0.0;
  }
}

// fragments of P_c for q^_c^T = m_c^T P_c with the permuted reduction index:
// B[k-lane c][n = child state 8 jt + r] = P_c[parent 8 nt + 2 c + e][8 jt + r]
template <int S>
__device__ __forceinline__ void gw_stage_q(double* dst, const double* Pl, const double* Pr,
                                           int nthreads) {
  using G = GwShape<S>;
  for (int idx = threadIdx.x; idx < G::QF * 32; idx += nthreads) {
    const int lane = idx & 31, f = idx >> 5;
    const int jt = f % G::NT, e = (f / G::NT) & 1, nt = (f / (2 * G::NT)) % G::NT,
              side = f / (2 * G::NT * G::NT);
    const int parent = 8 * nt + 2 * (lane & 3) + e, child = 8 * jt + (lane >> 2);
    const double* P = side ? Pr : Pl;
    dst[idx] = (child < S && parent < S) ? P[parent * S + child] : 0.0;
  }
}

// ---- per-warp cp.async ring --------------------------------------------------------------
// A stage holds the vectors one 8-pattern group needs, as [state][GW_LD] tiles (8 patterns of a
// state row = 64 contiguous bytes in HBM = four 16-byte cp.async; GW_LD = 12 doubles makes the
// A- and B-fragment reads below bank-conflict-free).  Every warp is its own producer and
// consumer: STAGES - 1 groups are in flight per warp and cost no registers.
constexpr int GW_LD = 12;

__device__ __forceinline__ void cp_async16(double* smem_dst, const double* gmem_src) {
  const unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async8(void* smem_dst, const void* gmem_src) {
  const unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(sa), "l"(gmem_src) : "memory");
}
template <int N>
__device__ __forceinline__ void cp_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// one [S][GW_LD] tile <- rows of `plane` at patterns i0 .. i0 + 7
template <int S>
__device__ __forceinline__ void gw_issue_tile(double* tile, const double* plane, int Npad, int i0,
                                              int lane) {
#pragma unroll
  for (int id = lane; id < S * 4; id += 32) {
    const int row = id >> 2, part = id & 3;
    cp_async16(tile + row * GW_LD + part * 2, plane + (size_t)row * Npad + i0 + part * 2);
  }
}

// A fragments of one child from its staged tile / from the tip code table (`code` = this
// lane's tip code for pattern i0 + r, prefetched one group ahead)
template <int S>
__device__ __forceinline__ void gw_frag_child(double (&a)[GwShape<S>::KT], bool tip, int code,
                                              const double* table, const double* tile, int lane) {
  using G = GwShape<S>;
  static_assert(S % 4 == 0, "state rows beyond S are not staged");
  const int r = lane >> 2, c = lane & 3;
  if (tip) {
    const double* cp = table + code * S + c;
#pragma unroll
    for (int kt = 0; kt < G::KT; ++kt) a[kt] = cp[4 * kt];
  } else {
#pragma unroll
    for (int kt = 0; kt < G::KT; ++kt) a[kt] = tile[(4 * kt + c) * GW_LD + r];
  }
}

// the tip code table in shared memory when it is small (it always is for real alphabets)
constexpr int GW_MAX_CODES = 24;
constexpr int GW_CODE_AHEAD = 5;  // post-order: tip codes are copied this many groups ahead of the tiles

// Tip children need no GEMM: u = P_c . codeP[code] is one of C vectors, tabulated per CTA
// (UT[side][code][parent state]) and read in C-fragment layout with one LDS.128 per tile.
// On a random tree half of all child slots are tips: half of the U-phase DMMAs disappear.
template <int S>
__device__ __forceinline__ void gw_stage_utab(double* ut, const double* Pl, const double* Pr,
                                              bool tipL, bool tipR, const double* codeP,
                                              int codeCount, int nthreads) {
  for (int idx = threadIdx.x; idx < 2 * codeCount * S; idx += nthreads) {
    const int side = idx / (codeCount * S), rem = idx - side * codeCount * S;
    if (!(side ? tipR : tipL)) continue;
    const int code = rem / S, s = rem - code * S;
    const double* P = (side ? Pr : Pl) + s * S;
    const double* v = codeP + code * S;
    double acc = 0.0;
    for (int j = 0; j < S; ++j) acc = fma(P[j], v[j], acc);
    ut[(side * GW_MAX_CODES + code) * S + s] = acc;
  }
}

// u^T of a tip child in C-fragment layout: pattern r, states 8 nt + 2 c + {0, 1}
template <int S>
__device__ __forceinline__ void gw_u_tip(double (&acc)[GwShape<S>::NT][2], const double* ut,
                                         int side, int code, int lane) {
  static_assert(S % 2 == 0, "pairs of states are read as double2");
  const int c = lane & 3;
  const double* row = ut + (side * GW_MAX_CODES + code) * S + 2 * c;
#pragma unroll
  for (int nt = 0; nt < GwShape<S>::NT; ++nt) {
    if (8 * nt + 2 * c < S) {
      const double2 v = *reinterpret_cast<const double2*>(row + 8 * nt);
      acc[nt][0] = v.x;
      acc[nt][1] = v.y;
    } else {
      acc[nt][0] = acc[nt][1] = 0.0;
    }
  }
}
template <int S>
__device__ __forceinline__ const double* gw_stage_codes(double* dst, const double* codeP,
                                                        int codeCount, int nthreads) {
  if (codeCount > GW_MAX_CODES) return codeP;
  for (int idx = threadIdx.x; idx < codeCount * S; idx += nthreads) dst[idx] = codeP[idx];
  return dst;
}

// acc[nt] += A . B over all k-steps (C fragment: pattern r, states 8 nt + 2 c + {0, 1})
template <int S>
__device__ __forceinline__ void gw_u(double (&acc)[GwShape<S>::NT][2],
                                     const double (&a)[GwShape<S>::KT], const double* frag,
                                     int lane) {
  using G = GwShape<S>;
#pragma unroll
  for (int kt = 0; kt < G::KT; ++kt)
#pragma unroll
    for (int nt = 0; nt < G::NT; ++nt)
      dmma(acc[nt][0], acc[nt][1], a[kt], frag[(kt * G::NT + nt) * 32 + lane]);
}

// ---------------------------------------------------------------------------------------
// post-order: grid (pattern chunks, nodes of the level x K, draws)
// shared: fragU [UF][32] | ring [NW][STAGES][2][S][GW_LD]
// ---------------------------------------------------------------------------------------
template <int S, int NW, int STAGES>
__global__ void __launch_bounds__(NW * 32)
gw_fwd_kernel(const NodeOp* __restrict__ ops, int opBegin, const double* __restrict__ mats,
              const uint8_t* __restrict__ tips, const double* __restrict__ codeP,
              double* __restrict__ partials, int16_t* __restrict__ expoK, int T, int Npad, int B,
              int K, int chunkPatterns, int codeCount) {
  using G = GwShape<S>;
  extern __shared__ double sm[];
  double* fragU = sm;  // [2][KT][NT][32]
  constexpr int SS = S * S;
  constexpr int TILE = S * GW_LD;
  double* smTable = sm + G::UF * 32 + NW * STAGES * 2 * TILE + NW * (GW_CODE_AHEAD + STAGES) * 2;
  const double* table = gw_stage_codes<S>(smTable, codeP, codeCount, NW * 32);
  const double* utab = smTable + GW_MAX_CODES * S;
  const bool useUtab = codeCount <= GW_MAX_CODES;
  const int nodeSlot = blockIdx.y / K;
  const int k = blockIdx.y - nodeSlot * K;
  const NodeOp op = ops[opBegin + nodeSlot];
  const int d = blockIdx.z;
  const int I = T - 1;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int r = lane >> 2, c = lane & 3;
  constexpr int SLOT = 2 * TILE;  // a stage: the two child tiles of a group
  double* ring = fragU + G::UF * 32 + (size_t)warp * STAGES * SLOT;
  // Tip codes travel in their own small ring (8 + 8 bytes per group), CODE_AHEAD groups ahead of
  // the tiles: a group of two tip children has nothing else to wait for, and its trip is far
  // shorter than an L2 round trip (the dependent code -> table address chain was 13 % of all
  // stall samples with codes prefetched two trips ahead in registers).
  // (a code slot is rewritten CODE_AHEAD + STAGES groups later: after its group has started)
  constexpr int CODE_AHEAD = GW_CODE_AHEAD, CSTAGES = GW_CODE_AHEAD + STAGES;
  double* cring = fragU + G::UF * 32 + (size_t)NW * STAGES * SLOT + (size_t)warp * CSTAGES * 2;
  const bool tipL = op.left < T, tipR = op.right < T;
  const size_t plane = (size_t)S * Npad;
  const size_t nodeStride = (size_t)K * plane;
  double* base = partials + (size_t)d * I * nodeStride + k * plane;
  const double* matsD = mats + (size_t)d * B * K * SS;
  gw_stage_u<S>(fragU, matsD + ((size_t)op.left * K + k) * SS,
                matsD + ((size_t)op.right * K + k) * SS, NW * 32);
  if (useUtab)
    gw_stage_utab<S>(smTable + GW_MAX_CODES * S, matsD + ((size_t)op.left * K + k) * SS,
                     matsD + ((size_t)op.right * K + k) * SS, tipL, tipR, codeP, codeCount, NW * 32);
  pdl_wait_then_trigger();  // the matrices come from pmatrix; the child vectors from the previous level
  __syncthreads();
  const uint8_t* tl = tips + (size_t)(tipL ? op.left : 0) * Npad;
  const uint8_t* tr = tips + (size_t)(tipR ? op.right : 0) * Npad;
  const double* pl = base + (size_t)(tipL ? 0 : op.left - T) * nodeStride;
  const double* pr = base + (size_t)(tipR ? 0 : op.right - T) * nodeStride;
  double* qn = base + (size_t)(op.node - T) * nodeStride;
  int16_t* en = expoK + (((size_t)d * I + (op.node - T)) * K + k) * Npad;

  const int begin = blockIdx.x * chunkPatterns;
  int end = begin + chunkPatterns;
  end = end < Npad ? end : Npad;
  const int first = begin + warp * 8;
  auto issue_codes = [&](int g) {
    const int i0 = first + g * (NW * 8);
    if (i0 < end) {
      if (tipL && lane == 0) cp_async8(cring + (g % CSTAGES) * 2, tl + i0);
      if (tipR && lane == 1) cp_async8(cring + (g % CSTAGES) * 2 + 1, tr + i0);
    }
  };
  auto issue = [&](int j) {
    const int i0 = first + j * (NW * 8);
    if (i0 < end) {
      double* slot = ring + (j % STAGES) * SLOT;
      if (!tipL) gw_issue_tile<S>(slot, pl, Npad, i0, lane);
      if (!tipR) gw_issue_tile<S>(slot + TILE, pr, Npad, i0, lane);
    }
    issue_codes(j + CODE_AHEAD);
    cp_commit();
  };
  static_assert(STAGES >= 3, "two groups are being consumed while the next ones travel");
#pragma unroll
  for (int g = 0; g < CODE_AHEAD; ++g) issue_codes(g);  // joins the first commit group below
#pragma unroll
  for (int j = 0; j < STAGES - 1; ++j) issue(j);
  if (first >= end) {
    cp_wait<0>();
    return;
  }
  // Software pipeline over the warp's groups: the DMMAs of group j + 1 are issued BEFORE the
  // epilogue (product, rescaling, stores) of group j, so the tensor pipe never drains while a
  // warp finishes a group.  Two accumulator sets alternate (the loop is unrolled by two).
  auto start_group = [&](double (&accL)[G::NT][2], double (&accR)[G::NT][2], int jj) {
    const double* slot = ring + (jj % STAGES) * SLOT;
    const uint8_t* codes = reinterpret_cast<const uint8_t*>(cring + (jj % CSTAGES) * 2);
    const int codeL = tipL ? codes[r] : 0, codeR = tipR ? codes[8 + r] : 0;
    if (tipL && useUtab) {
      gw_u_tip<S>(accL, utab, 0, codeL, lane);
    } else {
      double aL[G::KT];
      gw_frag_child<S>(aL, tipL, codeL, table, slot, lane);
#pragma unroll
      for (int nt = 0; nt < G::NT; ++nt) accL[nt][0] = accL[nt][1] = 0.0;
      gw_u<S>(accL, aL, fragU, lane);
    }
    if (tipR && useUtab) {
      gw_u_tip<S>(accR, utab, 1, codeR, lane);
    } else {
      double aR[G::KT];
      gw_frag_child<S>(aR, tipR, codeR, table, slot + TILE, lane);
#pragma unroll
      for (int nt = 0; nt < G::NT; ++nt) accR[nt][0] = accR[nt][1] = 0.0;
      gw_u<S>(accR, aR, fragU + G::KT * G::NT * 32, lane);
    }
  };
  auto finish_group = [&](double (&accL)[G::NT][2], double (&accR)[G::NT][2], int i0) {
    // maximum over the states on the integer pipe (the fp64 pipe is the DMMAs'): for non-negative
    // doubles the high words order like the values; lz = OR of the low words of values whose
    // high word is zero, so that (mh > 0 || lz != 0) <=> max > 0 (kernels_gmma.cu, gm_fwd3_kernel)
    int mh = 0, lz = 0;
#pragma unroll
    for (int nt = 0; nt < G::NT; ++nt) {
      accL[nt][0] *= accR[nt][0];
      accL[nt][1] *= accR[nt][1];
      const int h0 = __double2hiint(accL[nt][0]), h1 = __double2hiint(accL[nt][1]);
      mh = max(mh, max(h0, h1));
      lz |= (h0 == 0 ? __double2loint(accL[nt][0]) : 0) | (h1 == 0 ? __double2loint(accL[nt][1]) : 0);
    }
    mh = max(mh, __shfl_xor_sync(0xffffffffu, mh, 1));
    lz |= __shfl_xor_sync(0xffffffffu, lz, 1);
    mh = max(mh, __shfl_xor_sync(0xffffffffu, mh, 2));
    lz |= __shfl_xor_sync(0xffffffffu, lz, 2);
    const int eb = (mh >> 20) & 0x7ff;
    const int e = (mh > 0 || lz != 0) ? (eb > 2044 ? 2044 : eb) - 1022 : 0;
    const double f = __hiloint2double((1023 - e) << 20, 0);
    double* o = qn + i0 + r;
#pragma unroll
    for (int nt = 0; nt < G::NT; ++nt) {
      const int s0 = 8 * nt + 2 * c;
      if (s0 < S) o[(size_t)s0 * Npad] = accL[nt][0] * f;
      if (s0 + 1 < S) o[(size_t)(s0 + 1) * Npad] = accL[nt][1] * f;
    }
    if (c == 0) en[i0 + r] = (int16_t)e;
  };
  // one trip: start group j + 1 (if any) with `nxt`, finish group j held in `cur`
  auto trip = [&](double (&curL)[G::NT][2], double (&curR)[G::NT][2], double (&nxtL)[G::NT][2],
                  double (&nxtR)[G::NT][2], int jj, int i0) {
    issue(jj + STAGES - 1);  // into the slot of group j - 1, read one trip ago
    const int i1 = i0 + NW * 8;
    if (i1 < end) {
      cp_wait<STAGES - 2>();  // group j + 1 has landed
      __syncwarp();
      start_group(nxtL, nxtR, jj + 1);
    }
    finish_group(curL, curR, i0);
    __syncwarp();  // every lane has read slot j + 1 before the next trip overwrites slot j
  };
  double a0L[G::NT][2], a0R[G::NT][2], a1L[G::NT][2], a1R[G::NT][2];
  {
    cp_wait<STAGES - 2>();  // group 0
    __syncwarp();
    start_group(a0L, a0R, 0);
  }
  int j = 0;
  for (int i0 = first; i0 < end; i0 += 2 * NW * 8, j += 2) {
    trip(a0L, a0R, a1L, a1R, j, i0);
    if (i0 + NW * 8 < end) trip(a1L, a1R, a0L, a0R, j + 1, i0 + NW * 8);
  }
  cp_wait<0>();
}

// ---------------------------------------------------------------------------------------
// pre-order: grid (pattern chunks, nodes of the level x K, draws)
// shared: fragU [UF][32] | fragQ [QF][32] | ring [NW][STAGES][3][S][GW_LD] (q^_n, v_l, v_r);
// the ring is reused at the end to sum the warps' partial G in fixed order
// ---------------------------------------------------------------------------------------
template <int S, int NW, int STAGES>
__global__ void __launch_bounds__(NW * 32)
gw_bwd_kernel(const NodeOp* __restrict__ ops, int opBegin, const double* __restrict__ mats,
              const uint8_t* __restrict__ tips, const double* __restrict__ codeP,
              const double* __restrict__ partials, const int16_t* __restrict__ expoK,
              const double* __restrict__ weights, double* __restrict__ pre,
              double* __restrict__ gpart, const int* __restrict__ chunkBase, size_t chunkTotal,
              int T, int Npad, int B, int K, int chunkPatterns, int nChunk, int codeCount) {
  using G = GwShape<S>;
  extern __shared__ double sm[];
  double* fragU = sm;
  double* fragQ = fragU + G::UF * 32;
  double* ringAll = fragQ + G::QF * 32;
  constexpr int SS = S * S;
  constexpr int NT = G::NT;
  constexpr int TILE = S * GW_LD;
  // a stage: q^_n, v_l, v_r tiles | 8 + 8 tip codes (2 doubles) | 8 + 8 int16 scale exponents
  // of the children (4 doubles) | 8 pattern weights: everything a group reads rides the ring --
  // a plain load of the exponents at the top of the trip was 20 % of all stall samples
  constexpr int SLOT = 3 * TILE + 14;
  static_assert(STAGES * SLOT >= 2 * SS, "the ring doubles as the G staging area");
  double* smTable = ringAll + NW * STAGES * SLOT;
  const double* table = gw_stage_codes<S>(smTable, codeP, codeCount, NW * 32);
  const double* utab = smTable + GW_MAX_CODES * S;
  const bool useUtab = codeCount <= GW_MAX_CODES;
  const int nodeSlot = blockIdx.y / K;
  const int k = blockIdx.y - nodeSlot * K;
  const NodeOp op = ops[opBegin + nodeSlot];
  const int d = blockIdx.z;
  const int I = T - 1;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int r = lane >> 2, c = lane & 3;
  double* ring = ringAll + (size_t)warp * STAGES * SLOT;
  const bool tipL = op.left < T, tipR = op.right < T;
  const size_t plane = (size_t)S * Npad;
  const size_t nodeStride = (size_t)K * plane;
  const size_t drawBase = (size_t)d * I * nodeStride;
  const double* matsD = mats + (size_t)d * B * K * SS;
  {
    const double* Pl = matsD + ((size_t)op.left * K + k) * SS;
    const double* Pr = matsD + ((size_t)op.right * K + k) * SS;
    gw_stage_u<S>(fragU, Pl, Pr, NW * 32);
    gw_stage_q<S>(fragQ, Pl, Pr, NW * 32);
    if (useUtab)
      gw_stage_utab<S>(smTable + GW_MAX_CODES * S, Pl, Pr, tipL, tipR, codeP, codeCount, NW * 32);
  }
  pdl_wait_then_trigger();
  __syncthreads();

  const double* qsrc = pre + drawBase + (size_t)(op.node - T) * nodeStride + k * plane;
  const double* lsrc = partials + drawBase + (size_t)(tipL ? 0 : op.left - T) * nodeStride + k * plane;
  const double* rsrc = partials + drawBase + (size_t)(tipR ? 0 : op.right - T) * nodeStride + k * plane;
  double* lout = pre + drawBase + (size_t)(tipL ? 0 : op.left - T) * nodeStride + k * plane;
  double* rout = pre + drawBase + (size_t)(tipR ? 0 : op.right - T) * nodeStride + k * plane;
  const uint8_t* tl = tips + (size_t)(tipL ? op.left : 0) * Npad;
  const uint8_t* tr = tips + (size_t)(tipR ? op.right : 0) * Npad;
  const int16_t* elp = tipL ? nullptr : expoK + (((size_t)d * I + (op.left - T)) * K + k) * Npad;
  const int16_t* erp = tipR ? nullptr : expoK + (((size_t)d * I + (op.right - T)) * K + k) * Npad;

  // persistent d lnL / d P accumulators: C fragments [mt = parent tile][nt = child tile]
  double gL[NT][NT][2], gR[NT][NT][2];
#pragma unroll
  for (int a = 0; a < NT; ++a)
#pragma unroll
    for (int b = 0; b < NT; ++b) gL[a][b][0] = gL[a][b][1] = gR[a][b][0] = gR[a][b][1] = 0.0;

  const int begin = blockIdx.x * chunkPatterns;
  int end = begin + chunkPatterns;
  end = end < Npad ? end : Npad;
  const int first = begin + warp * 8;
  auto issue = [&](int j) {
    const int i0 = first + j * (NW * 8);
    if (i0 < end) {
      double* slot = ring + (j % STAGES) * SLOT;
      gw_issue_tile<S>(slot, qsrc, Npad, i0, lane);
      if (!tipL) gw_issue_tile<S>(slot + TILE, lsrc, Npad, i0, lane);
      else if (lane == 0) cp_async8(slot + 3 * TILE, tl + i0);
      if (!tipR) gw_issue_tile<S>(slot + 2 * TILE, rsrc, Npad, i0, lane);
      else if (lane == 1) cp_async8(slot + 3 * TILE + 1, tr + i0);
      if (!tipL && lane == 2) cp_async16(slot + 3 * TILE + 2, reinterpret_cast<const double*>(elp + i0));
      if (!tipR && lane == 3) cp_async16(slot + 3 * TILE + 4, reinterpret_cast<const double*>(erp + i0));
      if (lane >= 4 && lane < 8)
        cp_async16(slot + 3 * TILE + 6 + 2 * (lane - 4), weights + i0 + 2 * (lane - 4));
    }
    cp_commit();
  };
#pragma unroll
  for (int j = 0; j < STAGES - 1; ++j) issue(j);
  int j = 0;
  for (int i0 = first; i0 < end; i0 += NW * 8, ++j) {
    issue(j + STAGES - 1);
    cp_wait<STAGES - 1>();
    __syncwarp();
    const double* slot = ring + (j % STAGES) * SLOT;
    const uint8_t* codes = reinterpret_cast<const uint8_t*>(slot + 3 * TILE);
    const int16_t* expo = reinterpret_cast<const int16_t*>(slot + 3 * TILE + 2);
    const double w = slot[3 * TILE + 6 + r];
    const int el = tipL ? 0 : (int)expo[r];
    const int er = tipR ? 0 : (int)expo[8 + r];
    const int codeL = tipL ? codes[r] : 0, codeR = tipR ? codes[8 + r] : 0;
    double uL[NT][2], uR[NT][2];
    if (tipL && useUtab) {
      gw_u_tip<S>(uL, utab, 0, codeL, lane);
    } else {
      double aL[G::KT];
      gw_frag_child<S>(aL, tipL, codeL, table, slot + TILE, lane);
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) uL[nt][0] = uL[nt][1] = 0.0;
      gw_u<S>(uL, aL, fragU, lane);
    }
    if (tipR && useUtab) {
      gw_u_tip<S>(uR, utab, 1, codeR, lane);
    } else {
      double aR[G::KT];
      gw_frag_child<S>(aR, tipR, codeR, table, slot + 2 * TILE, lane);
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) uR[nt][0] = uR[nt][1] = 0.0;
      gw_u<S>(uR, aR, fragU + G::KT * NT * 32, lane);
    }
    // m_l = q^ o u_r, m_r = q^ o u_l with q^_n^T read in C-fragment layout
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      const int s0 = 8 * nt + 2 * c;
      const double q0 = s0 < S ? slot[s0 * GW_LD + r] : 0.0;
      const double q1 = s0 + 1 < S ? slot[(s0 + 1) * GW_LD + r] : 0.0;
      const double ml0 = q0 * uR[nt][0], ml1 = q1 * uR[nt][1];
      uR[nt][0] = q0 * uL[nt][0];
      uR[nt][1] = q1 * uL[nt][1];
      uL[nt][0] = ml0;
      uL[nt][1] = ml1;
    }
    // from here: uL = m_l, uR = m_r
#pragma unroll
    for (int side = 0; side < 2; ++side) {
      const bool tip = side ? tipR : tipL;
      const double(&m)[NT][2] = side ? uR : uL;
      // Q phase: q^_c^T = m_c^T P_c, scaled by 2^-e_c (internal children only)
      if (!tip) {
        double qo[NT][2];
#pragma unroll
        for (int jt = 0; jt < NT; ++jt) qo[jt][0] = qo[jt][1] = 0.0;
        const double* fq = fragQ + side * (2 * NT * NT) * 32 + lane;
#pragma unroll
        for (int nt = 0; nt < NT; ++nt)
#pragma unroll
          for (int e = 0; e < 2; ++e)
#pragma unroll
            for (int jt = 0; jt < NT; ++jt)
              dmma(qo[jt][0], qo[jt][1], m[nt][e], fq[((nt * 2 + e) * NT + jt) * 32]);
        const double f = __hiloint2double((1023 - (side ? er : el)) << 20, 0);
        double* o = (side ? rout : lout) + i0 + r;
#pragma unroll
        for (int jt = 0; jt < NT; ++jt) {
          const int s0 = 8 * jt + 2 * c;
          if (s0 < S) o[(size_t)s0 * Npad] = qo[jt][0] * f;
          if (s0 + 1 < S) o[(size_t)(s0 + 1) * Npad] = qo[jt][1] * f;
        }
      }
      // G phase: G_c[s][t] += sum_p (w_p m_c[s][p]) v_c[t][p], contraction over the 8 patterns
      const double* vt = slot + (1 + side) * TILE;
#pragma unroll
      for (int kp = 0; kp < 2; ++kp) {
        // B[k = pattern 4 kp + c][n = child state 8 nt + r]
        double bv[NT];
        if (tip) {
          const double* cp = table + (int)codes[side * 8 + 4 * kp + c] * S + r;
#pragma unroll
          for (int nt = 0; nt < NT; ++nt) bv[nt] = (8 * nt + r < S) ? cp[8 * nt] : 0.0;
        } else {
#pragma unroll
          for (int nt = 0; nt < NT; ++nt)
            bv[nt] = (8 * nt + r < S) ? vt[(8 * nt + r) * GW_LD + 4 * kp + c] : 0.0;
        }
        // A[row = parent state 8 mt + r][k = pattern 4 kp + c]: held by lane
        // (pattern 4 kp + c, c' = r >> 1) in slot r & 1
        const int src = ((4 * kp + c) << 2) | (r >> 1);
#pragma unroll
        for (int mt = 0; mt < NT; ++mt) {
          const double x0 = __shfl_sync(0xffffffffu, w * m[mt][0], src);
          const double x1 = __shfl_sync(0xffffffffu, w * m[mt][1], src);
          const double av = (r & 1) ? x1 : x0;
#pragma unroll
          for (int nt = 0; nt < NT; ++nt) {
            if (side) dmma(gR[mt][nt][0], gR[mt][nt][1], av, bv[nt]);
            else dmma(gL[mt][nt][0], gL[mt][nt][1], av, bv[nt]);
          }
        }
      }
    }
    __syncwarp();  // the slot is free for the copy issued in the next trip
  }
  cp_wait<0>();
  __syncwarp();

  // per-warp partial G -> the warp's own ring area, summed over warps in fixed order
  {
    double* mine = ring;
#pragma unroll
    for (int mt = 0; mt < NT; ++mt)
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        const int row = 8 * mt + r, col = 8 * nt + 2 * c;
        if (row < S && col < S) {
          mine[row * S + col] = gL[mt][nt][0];
          mine[SS + row * S + col] = gR[mt][nt][0];
        }
        if (row < S && col + 1 < S) {
          mine[row * S + col + 1] = gL[mt][nt][1];
          mine[SS + row * S + col + 1] = gR[mt][nt][1];
        }
      }
  }
  __syncthreads();
  double* oL = gpart + ((size_t)d * chunkTotal + chunkBase[op.left] + (size_t)k * nChunk + blockIdx.x) * SS;
  double* oR = gpart + ((size_t)d * chunkTotal + chunkBase[op.right] + (size_t)k * nChunk + blockIdx.x) * SS;
  for (int idx = threadIdx.x; idx < 2 * SS; idx += NW * 32) {
    double t = 0.0;
#pragma unroll
    for (int wv = 0; wv < NW; ++wv) t += ringAll[(size_t)wv * STAGES * SLOT + idx];
    if (idx < SS) oL[idx] = t;
    else oR[idx - SS] = t;
  }
}

template <int S, int NW, int STAGES>
size_t gw_fwd_smem() {
  return ((size_t)GwShape<S>::UF * 32 + (size_t)NW * STAGES * 2 * S * GW_LD +
          (size_t)NW * (GW_CODE_AHEAD + STAGES) * 2 + (size_t)3 * GW_MAX_CODES * S) * sizeof(double);
}
template <int S, int NW, int STAGES>
size_t gw_bwd_smem() {
  return ((size_t)(GwShape<S>::UF + GwShape<S>::QF) * 32 + (size_t)NW * STAGES * (3 * S * GW_LD + 14) +
          (size_t)3 * GW_MAX_CODES * S) * sizeof(double);
}

constexpr int GW_GRANULE = 64;  // chunk sizes are multiples of 8 patterns x 8 warps

template <int NW, int STAGES>
int gw_launch_fwd(Engine& e, int draws, int ctas) {
  const Dims& m = e.dm;
  auto kern = gw_fwd_kernel<20, NW, STAGES>;
  const size_t smem = gw_fwd_smem<20, NW, STAGES>();
  if (smem > 48 * 1024)
    TTB2_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)smem));
  const int nLevels = (int)e.levelOff.size() - 1;
  const int maxNodes = 65535 / m.K;
  // level 1 (tip-tip nodes) needs no tensor cores: the streaming kernel of kernels_gmma.cu
  const bool cherryLevel = gmma_cherry_level_supported(e);
  if (cherryLevel) {
    const int rc = gmma_cherry_forward_level(e, draws);
    if (rc) return rc;
  }
  for (int l = cherryLevel ? 1 : 0; l < nLevels; ++l) {
    const int opBegin = e.levelOff[l];
    const int count = e.levelOff[l + 1] - opBegin;
    // patterns per CTA: `ctas` CTAs per SM per launch
    const long target = (long)e.smCount * ctas;
    long chunks = (target + (long)count * m.K * draws - 1) / ((long)count * m.K * draws);
    const long maxChunks = (m.Npad + GW_GRANULE - 1) / GW_GRANULE;
    if (chunks > maxChunks) chunks = maxChunks;
    if (chunks < 1) chunks = 1;
    int chunkPatterns = (int)((m.Npad + chunks - 1) / chunks);
    chunkPatterns = (chunkPatterns + GW_GRANULE - 1) / GW_GRANULE * GW_GRANULE;
    const int nChunk = (m.Npad + chunkPatterns - 1) / chunkPatterns;
    for (int done = 0; done < count; done += maxNodes) {
      const int cnt = (count - done) < maxNodes ? (count - done) : maxNodes;
      dim3 grid(nChunk, cnt * m.K, draws);
      launch_level(kern, grid, NW * 32, smem, e.stream, l > 0 && pdl_enabled(), e.ops,
                   opBegin + done, e.mats, e.tips, e.codeP, e.partials, e.expoK, m.T, m.Npad, m.B,
                   m.K, chunkPatterns, e.cfg.code_count);
      ++e.launches;
    }
  }
  TTB2_CUDA_CHECK(cudaGetLastError());
  return TTB2_OK;
}

template <int NW, int STAGES>
int gw_launch_bwd(Engine& e, int draws) {
  const Dims& m = e.dm;
  auto kern = gw_bwd_kernel<20, NW, STAGES>;
  const size_t smem = gw_bwd_smem<20, NW, STAGES>();
  if (smem > 48 * 1024)
    TTB2_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)smem));
  const int nLevels = (int)e.levelOff.size() - 1;
  const int maxNodes = 65535 / m.K;
  for (int l = nLevels - 1; l >= 0; --l) {
    if (l == 0 && gmma_cherry_backward_supported(e)) {   // tip-tip level: G-only kernel
      const int rc = gmma_cherry_backward_level(e, draws, l < nLevels - 1 && pdl_enabled());
      if (rc) return rc;
      continue;
    }
    const int opBegin = e.levelOff[l];
    const int count = e.levelOff[l + 1] - opBegin;
    const int nChunk = e.levelChunks[l];
    int chunkPatterns = (m.Npad + nChunk - 1) / nChunk;
    chunkPatterns = (chunkPatterns + 31) / 32 * 32;   // as planned by plan_chunks (granule 32)
    for (int done = 0; done < count; done += maxNodes) {
      const int cnt = (count - done) < maxNodes ? (count - done) : maxNodes;
      dim3 grid(nChunk, cnt * m.K, draws);
      launch_level(kern, grid, NW * 32, smem, e.stream, l < nLevels - 1 && pdl_enabled(), e.ops,
                   opBegin + done, e.mats, e.tips, e.codeP, e.partials, e.expoK, e.weights, e.pre,
                   e.gpart, e.chunkBase, e.chunkTotal, m.T, m.Npad, m.B, m.K, chunkPatterns,
                   nChunk, e.cfg.code_count);
      ++e.launches;
    }
  }
  TTB2_CUDA_CHECK(cudaGetLastError());
  return TTB2_OK;
}

}  // namespace

// TTB2_GM_LEGACY=1: shared-memory tile kernels for both sweeps; =fwd / =bwd: for that sweep only
bool gwarp_supported(const Engine& e, bool backward) {
  const char* legacy = getenv("TTB2_GM_LEGACY");  // read per call: the tests switch it
  if (e.dm.S != 20) return false;
  if (!legacy) return true;
  if (legacy[0] == 'f') return backward;
  if (legacy[0] == 'b') return !backward;
  return false;
}

int gwarp_forward(Engine& e, int draws) {
  // 8 warps x 3 stages, 16 CTAs per SM aimed at per launch: measured best on config 4 against
  // <4,3>, <4,4>, <4,5> and 8 / 32 CTAs per SM (profiles/r01_config4_gwarp_tuning.log)
  return gw_launch_fwd<8, 3>(e, draws, 16);
}

// the pre-order level sweep (the root kernel and the gpart reduction stay with the caller)
int gwarp_backward_levels(Engine& e, int draws) {
  // 4 warps x 2 stages: 68 KB of shared memory per CTA -> 3 CTAs (12 warps) per SM; measured on
  // config 4: <4,2> 10.0 ms, <4,3> 12.5, <4,4> 12.6, <8,3> 13.6 (profiles/r01_config4_gwarp_tuning.log)
  return gw_launch_bwd<4, 2>(e, draws);
}

}  // namespace ttb2
