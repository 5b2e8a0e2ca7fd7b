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
#include "engine.cuh"

namespace ttb2 {

namespace {

constexpr int GM_WARPS = 8;
constexpr int GM_THREADS = GM_WARPS * 32;
constexpr int GM_TP = 32;    // patterns per tile
constexpr int GM_LDT = 36;   // leading dimension of [rows][32 patterns] tiles
constexpr int GM_MAXACC = 16;  // G accumulator tiles per warp: 2 * (64/8)^2 / 8

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
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

// zero-padded child tile [R][LDT]
__device__ __forceinline__ void gm_stage_tile(double* tile, bool tip, const uint8_t* tipRow,
                                              const double* codeP, const double* plane, int i0,
                                              int Npad, const GmShape g) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int i = i0 + lane;
  if (tip) {
    const double* cp = codeP + (size_t)tipRow[i] * g.S;
    for (int s = warp; s < g.R; s += GM_WARPS) tile[s * GM_LDT + lane] = s < g.S ? cp[s] : 0.0;
  } else {
    for (int s = warp; s < g.R; s += GM_WARPS)
      tile[s * GM_LDT + lane] = s < g.S ? plane[(size_t)s * Npad + i] : 0.0;
  }
}

// D[mt][nt] (+)= A[mt rows][.] . B[.][nt cols], A row-major with leading dimension lda
// (element (r, c) at A[r*lda + c]), B as [contraction][32 patterns] tile.
__device__ __forceinline__ void gm_mma_ab(double& c0, double& c1, const double* A, int lda,
                                          const double* Bt, int mt, int nt, int ksteps,
                                          int lane) {
  const double* a = A + (mt * 8 + (lane >> 2)) * lda + (lane & 3);
  const double* b = Bt + (lane & 3) * GM_LDT + nt * 8 + (lane >> 2);
  for (int kt = 0; kt < ksteps; ++kt) dmma884(c0, c1, a[kt * 4], b[kt * 4 * GM_LDT]);
}

// same with A transposed: element (r, c) of the operand at A[c*lda + r]
__device__ __forceinline__ void gm_mma_atb(double& c0, double& c1, const double* A, int lda,
                                           const double* Bt, int mt, int nt, int ksteps,
                                           int lane) {
  const double* a = A + (lane & 3) * lda + mt * 8 + (lane >> 2);
  const double* b = Bt + (lane & 3) * GM_LDT + nt * 8 + (lane >> 2);
  for (int kt = 0; kt < ksteps; ++kt) dmma884(c0, c1, a[kt * 4 * lda], b[kt * 4 * GM_LDT]);
}

// grouped variants: one A fragment feeds NTG pattern tiles (NTG independent
// accumulator chains, 1 + NTG shared-memory loads per NTG MMAs)
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

// ---------------------------------------------------------------------------
// post-order.  grid (pattern tiles, nodes of level, draws)
// shared: Pl Pr [Sp*PLD] | cl cr [R*LDT] | out [K][Sp][LDT] | wmax [8*32]
// ---------------------------------------------------------------------------
template <int NTG>
__global__ void __launch_bounds__(GM_THREADS)
gm_fwd_kernel(const NodeOp* __restrict__ ops, int opBegin, const double* __restrict__ mats,
              const uint8_t* __restrict__ tips, const double* __restrict__ codeP,
              double* __restrict__ partials, int16_t* __restrict__ expo, int T, int Npad, int B,
              int K, int S) {
  extern __shared__ double sm[];
  const GmShape g = gm_shape(S);
  const int SS = S * S;
  double* Pl = sm;
  double* Pr = Pl + g.Sp * g.PLD;
  double* cl = Pr + g.Sp * g.PLD;
  double* cr = cl + g.R * GM_LDT;
  double* out = cr + g.R * GM_LDT;
  double* wmax = out + (size_t)K * g.Sp * GM_LDT;

  const NodeOp op = ops[opBegin + blockIdx.y];
  const int d = blockIdx.z;
  const int I = T - 1;
  const int i0 = blockIdx.x * GM_TP;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const bool tipL = op.left < T, tipR = op.right < T;
  const size_t plane = (size_t)S * Npad;
  const size_t nodeStride = (size_t)K * plane;
  double* base = partials + (size_t)d * I * nodeStride;
  const double* matsD = mats + (size_t)d * B * K * SS;
  const int MT = g.Sp / 8, KT = g.Kp / 4;

  for (int k = 0; k < K; ++k) {
    __syncthreads();
    gm_stage_matrix(Pl, matsD + ((size_t)op.left * K + k) * SS, g);
    gm_stage_matrix(Pr, matsD + ((size_t)op.right * K + k) * SS, g);
    gm_stage_tile(cl, tipL, tips + (size_t)(tipL ? op.left : 0) * Npad, codeP,
                  base + (size_t)(tipL ? 0 : op.left - T) * nodeStride + k * plane, i0, Npad, g);
    gm_stage_tile(cr, tipR, tips + (size_t)(tipR ? op.right : 0) * Npad, codeP,
                  base + (size_t)(tipR ? 0 : op.right - T) * nodeStride + k * plane, i0, Npad, g);
    __syncthreads();
    constexpr int NG = 4 / NTG;  // groups of pattern tiles
    for (int item = warp; item < MT * NG; item += GM_WARPS) {
      const int mt = item / NG, nt0 = (item - mt * NG) * NTG;
      double accL[NTG][2], accR[NTG][2];
#pragma unroll
      for (int n = 0; n < NTG; ++n) accL[n][0] = accL[n][1] = accR[n][0] = accR[n][1] = 0.0;
      gm_mma_ab_g<NTG>(accL, Pl, g.PLD, cl, mt, nt0, KT, lane);
      gm_mma_ab_g<NTG>(accR, Pr, g.PLD, cr, mt, nt0, KT, lane);
      double* o = out + ((size_t)k * g.Sp + mt * 8 + (lane >> 2)) * GM_LDT + nt0 * 8 + (lane & 3) * 2;
#pragma unroll
      for (int n = 0; n < NTG; ++n) {
        o[n * 8] = accL[n][0] * accR[n][0];
        o[n * 8 + 1] = accL[n][1] * accR[n][1];
      }
    }
  }
  __syncthreads();
  double m = 0.0;
  for (int ks = warp; ks < K * g.Sp; ks += GM_WARPS) {
    const int s = ks % g.Sp;
    if (s < S) m = fmax(m, out[(size_t)ks * GM_LDT + lane]);
  }
  wmax[warp * 32 + lane] = m;
  __syncthreads();
  double mm = 0.0;
#pragma unroll
  for (int w = 0; w < GM_WARPS; ++w) mm = fmax(mm, wmax[w * 32 + lane]);
  int eb = (__double2hiint(mm) >> 20) & 0x7ff;
  eb = eb > 2044 ? 2044 : eb;
  const double f = __hiloint2double((2045 - eb) << 20, 0);
  double* q = base + (size_t)(op.node - T) * nodeStride + i0 + lane;
  for (int ks = warp; ks < K * S; ks += GM_WARPS) {
    const int k = ks / S, s = ks - k * S;
    q[(size_t)ks * Npad] = out[((size_t)k * g.Sp + s) * GM_LDT + lane] * f;
  }
  if (warp == 0) expo[((size_t)d * I + (op.node - T)) * Npad + i0 + lane] = (int16_t)(eb - 1022);
}

// ---------------------------------------------------------------------------
// pre-order.  grid (pattern chunks, nodes of level x K, draws)
// shared: Pl Pr [Sp*PLD] | tq vl vr ml mr [R*LDT each] | ws[32] | el er [32] (int)
// ---------------------------------------------------------------------------
template <int NTG>
__global__ void __launch_bounds__(GM_THREADS)
gm_bwd_kernel(const NodeOp* __restrict__ ops, int opBegin, const double* __restrict__ mats,
              const uint8_t* __restrict__ tips, const double* __restrict__ codeP,
              const double* __restrict__ partials, const int16_t* __restrict__ expo,
              const double* __restrict__ weights, double* __restrict__ pre,
              double* __restrict__ gpart, const int* __restrict__ chunkBase, size_t chunkTotal,
              int T, int Npad, int B, int K, int S, int chunkPatterns, int nChunk) {
  extern __shared__ double sm[];
  const GmShape g = gm_shape(S);
  const int SS = S * S;
  double* Pl = sm;
  double* Pr = Pl + g.Sp * g.PLD;
  double* tq = Pr + g.Sp * g.PLD;
  double* vl = tq + g.R * GM_LDT;
  double* vr = vl + g.R * GM_LDT;
  double* ml = vr + g.R * GM_LDT;
  double* mr = ml + g.R * GM_LDT;
  double* ws = mr + g.R * GM_LDT;
  int* se = reinterpret_cast<int*>(ws + 32);  // el[32], er[32]

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
  gm_stage_matrix(Pl, matsD + ((size_t)op.left * K + k) * SS, g);
  gm_stage_matrix(Pr, matsD + ((size_t)op.right * K + k) * SS, g);

  // persistent G accumulators: items (child, mt, mt2) dealt round-robin to the warps
  double acc[GM_MAXACC][2];
#pragma unroll
  for (int j = 0; j < GM_MAXACC; ++j) acc[j][0] = acc[j][1] = 0.0;

  const int begin = blockIdx.x * chunkPatterns;
  int end = begin + chunkPatterns;
  end = end < Npad ? end : Npad;
  for (int i0 = begin; i0 < end; i0 += GM_TP) {
    __syncthreads();
    gm_stage_tile(tq, false, nullptr, codeP,
                  pre + drawBase + (size_t)(op.node - T) * nodeStride + k * plane, i0, Npad, g);
    gm_stage_tile(vl, tipL, tips + (size_t)(tipL ? op.left : 0) * Npad, codeP,
                  partials + drawBase + (size_t)(tipL ? 0 : op.left - T) * nodeStride + k * plane,
                  i0, Npad, g);
    gm_stage_tile(vr, tipR, tips + (size_t)(tipR ? op.right : 0) * Npad, codeP,
                  partials + drawBase + (size_t)(tipR ? 0 : op.right - T) * nodeStride + k * plane,
                  i0, Npad, g);
    if (warp == 0) {
      const int i = i0 + lane;
      ws[lane] = weights[i];
      se[lane] = tipL ? 0 : (int)expo[((size_t)d * I + (op.left - T)) * Npad + i];
      se[32 + lane] = tipR ? 0 : (int)expo[((size_t)d * I + (op.right - T)) * Npad + i];
    }
    __syncthreads();
    // U phase: u_l = P_l v_l, u_r = P_r v_r;  m_l = q^ o u_r, m_r = q^ o u_l
    constexpr int NG = 4 / NTG;
    for (int item = warp; item < MT * NG; item += GM_WARPS) {
      const int mt = item / NG, nt0 = (item - mt * NG) * NTG;
      double accL[NTG][2], accR[NTG][2];
#pragma unroll
      for (int n = 0; n < NTG; ++n) accL[n][0] = accL[n][1] = accR[n][0] = accR[n][1] = 0.0;
      gm_mma_ab_g<NTG>(accL, Pl, g.PLD, vl, mt, nt0, KT, lane);
      gm_mma_ab_g<NTG>(accR, Pr, g.PLD, vr, mt, nt0, KT, lane);
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
      for (int item = warp; item < MT * NG; item += GM_WARPS) {
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
            const double f0 = __hiloint2double((1023 - se[side * 32 + col]) << 20, 0);
            const double f1 = __hiloint2double((1023 - se[side * 32 + col + 1]) << 20, 0);
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
    for (int cj = 0; cj < 2; ++cj) {
      const int combo = warp + cj * GM_WARPS;
      if (combo < 2 * MT) {
        const int side = combo / MT;
        const int mt = combo - side * MT;
        const double* mm = (side ? mr : ml) + (mt * 8 + (lane >> 2)) * GM_LDT + (lane & 3);
        const double* vv = (side ? vr : vl) + (lane >> 2) * GM_LDT + (lane & 3);
        const double* wp = ws + (lane & 3);
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
#pragma unroll
  for (int cj = 0; cj < 2; ++cj) {
    const int combo = warp + cj * GM_WARPS;
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

size_t gm_fwd_smem(const Dims& m) {
  const GmShape g = gm_shape(m.S);
  return (2 * (size_t)g.Sp * g.PLD + 2 * (size_t)g.R * GM_LDT + (size_t)m.K * g.Sp * GM_LDT +
          GM_WARPS * 32) * sizeof(double);
}

size_t gm_bwd_smem(const Dims& m) {
  const GmShape g = gm_shape(m.S);
  return (2 * (size_t)g.Sp * g.PLD + 5 * (size_t)g.R * GM_LDT + 32) * sizeof(double) +
         64 * sizeof(int);
}

}  // namespace

bool gmma_supported(const Engine& e) {
  const Dims& m = e.dm;
  return !e.spec4 && m.S >= 8 && m.S <= 64 && !(e.cfg.flags & TTB2_FLAG_NO_MMA) &&
         !(e.cfg.flags & TTB2_FLAG_FORCE_GENERIC) && gm_fwd_smem(m) <= 227 * 1024 &&
         gm_bwd_smem(m) <= 227 * 1024;
}

int gmma_forward(Engine& e, int draws) {
  const Dims& m = e.dm;
  const size_t smem = gm_fwd_smem(m);
  const int MT = gm_shape(m.S).Sp / 8;
  const int ntg = MT >= 8 ? 4 : (MT >= 4 ? 2 : 1);
  auto kern = ntg == 4 ? gm_fwd_kernel<4> : (ntg == 2 ? gm_fwd_kernel<2> : gm_fwd_kernel<1>);
  if (smem > 48 * 1024)
    TTB2_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)smem));
  const int nLevels = (int)e.levelOff.size() - 1;
  for (int l = 0; l < nLevels; ++l) {
    const int opBegin = e.levelOff[l];
    const int count = e.levelOff[l + 1] - opBegin;
    for (int done = 0; done < count; done += 65535) {
      const int c = (count - done) < 65535 ? (count - done) : 65535;
      dim3 grid(m.Npad / GM_TP, c, draws);
      kern<<<grid, GM_THREADS, smem, e.stream>>>(
          e.ops, opBegin + done, e.mats, e.tips, e.codeP, e.partials, e.expo, m.T, m.Npad, m.B,
          m.K, m.S);
      ++e.launches;
    }
  }
  TTB2_CUDA_CHECK(cudaGetLastError());
  return TTB2_OK;
}

// pre-order level launches only (the root kernels are the generic ones)
int gmma_backward_levels(Engine& e, int draws) {
  const Dims& m = e.dm;
  const size_t smem = gm_bwd_smem(m);
  const int MT = gm_shape(m.S).Sp / 8;
  const int ntg = MT >= 8 ? 4 : (MT >= 4 ? 2 : 1);
  auto kern = ntg == 4 ? gm_bwd_kernel<4> : (ntg == 2 ? gm_bwd_kernel<2> : gm_bwd_kernel<1>);
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
    for (int done = 0; done < count; done += maxNodes) {
      const int c = (count - done) < maxNodes ? (count - done) : maxNodes;
      dim3 grid(nChunk, c * m.K, draws);
      kern<<<grid, GM_THREADS, smem, e.stream>>>(
          e.ops, opBegin + done, e.mats, e.tips, e.codeP, e.partials, e.expo, e.weights, e.pre,
          e.gpart, e.chunkBase, e.chunkTotal, m.T, m.Npad, m.B, m.K, m.S, chunkPatterns, nChunk);
      ++e.launches;
    }
  }
  TTB2_CUDA_CHECK(cudaGetLastError());
  return TTB2_OK;
}

}  // namespace ttb2
