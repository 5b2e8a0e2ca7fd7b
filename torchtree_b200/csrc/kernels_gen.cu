// Generic-S level-synchronous peeling kernels (any state count: 20-state amino
// acid, 61-state codon, general discrete traits).
//
// Layout: [draw][inode][k][s][pattern] (state-major planes) so that a warp
// reading one state of 32 consecutive patterns is one coalesced 256-byte
// access.  A CTA owns a tile of 32 patterns of one node; the 32 lanes of each
// warp are the patterns and the warps split the state rows.  Child vectors and
// the two S x S transition matrices are staged in shared memory.
//
// Same algorithm and scaling convention as kernels_s4.cu (power-of-two
// rescaling, int16 exponents per (node, pattern)).
#include "engine.cuh"

namespace ttb2 {

namespace {

constexpr int GEN_WARPS = 8;
constexpr int GEN_THREADS = GEN_WARPS * 32;
constexpr int TP = 32;    // patterns per tile
constexpr int LDT = 33;   // padded leading dimension of [S][TP] tiles
constexpr int ROOT_THREADS = 128;

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

// Stage one child's conditional-likelihood tile [S][LDT] for category k.
__device__ __forceinline__ void stage_child(double* tile, bool tip, const uint8_t* tipRow,
                                            const double* codeP, const double* plane,
                                            int i0, int Npad, int S) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int i = i0 + lane;
  if (tip) {
    const int code = tipRow[i];
    const double* cp = codeP + (size_t)code * S;
    for (int s = warp; s < S; s += GEN_WARPS) tile[s * LDT + lane] = cp[s];
  } else {
    for (int s = warp; s < S; s += GEN_WARPS) tile[s * LDT + lane] = plane[(size_t)s * Npad + i];
  }
}

__device__ __forceinline__ void stage_matrix(double* dst, const double* src, int SS) {
  for (int j = threadIdx.x; j < SS; j += blockDim.x) dst[j] = src[j];
}

// shared memory (doubles): P_l[S*S] P_r[S*S] cl[S*LDT] cr[S*LDT] out[K*S*LDT] wmax[GEN_WARPS*32]
__global__ void __launch_bounds__(GEN_THREADS)
gen_fwd_kernel(const NodeOp* __restrict__ ops, int opBegin, const double* __restrict__ mats,
               const uint8_t* __restrict__ tips, const double* __restrict__ codeP,
               double* __restrict__ partials, int16_t* __restrict__ expo, int T, int Npad,
               int B, int K, int S) {
  extern __shared__ double sm[];
  const int SS = S * S;
  double* Pl = sm;
  double* Pr = Pl + SS;
  double* cl = Pr + SS;
  double* cr = cl + S * LDT;
  double* out = cr + S * LDT;
  double* wmax = out + (size_t)K * S * LDT;

  const NodeOp op = ops[opBegin + blockIdx.y];
  const int d = blockIdx.z;
  const int I = T - 1;
  const int i0 = blockIdx.x * TP;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const bool tipL = op.left < T, tipR = op.right < T;
  const size_t plane = (size_t)S * Npad;          // one category of one node
  const size_t nodeStride = (size_t)K * plane;
  double* base = partials + (size_t)d * I * nodeStride;
  const double* matsD = mats + (size_t)d * B * K * SS;

  double m = 0.0;
  for (int k = 0; k < K; ++k) {
    __syncthreads();  // previous category's tiles are no longer read
    stage_matrix(Pl, matsD + ((size_t)op.left * K + k) * SS, SS);
    stage_matrix(Pr, matsD + ((size_t)op.right * K + k) * SS, SS);
    stage_child(cl, tipL, tips + (size_t)(tipL ? op.left : 0) * Npad, codeP,
                base + (size_t)(tipL ? 0 : op.left - T) * nodeStride + k * plane, i0, Npad, S);
    stage_child(cr, tipR, tips + (size_t)(tipR ? op.right : 0) * Npad, codeP,
                base + (size_t)(tipR ? 0 : op.right - T) * nodeStride + k * plane, i0, Npad, S);
    __syncthreads();
    for (int s = warp; s < S; s += GEN_WARPS) {
      double ul = 0.0, ur = 0.0;
      const double* rl = Pl + s * S;
      const double* rr = Pr + s * S;
      for (int t = 0; t < S; ++t) {
        ul = fma(rl[t], cl[t * LDT + lane], ul);
        ur = fma(rr[t], cr[t * LDT + lane], ur);
      }
      const double o = ul * ur;
      out[((size_t)k * S + s) * LDT + lane] = o;
      m = fmax(m, o);
    }
  }
  wmax[warp * 32 + lane] = m;
  __syncthreads();
  double mm = 0.0;
#pragma unroll
  for (int w = 0; w < GEN_WARPS; ++w) mm = fmax(mm, wmax[w * 32 + lane]);
  int eb = (__double2hiint(mm) >> 20) & 0x7ff;
  eb = eb > 2044 ? 2044 : eb;
  const double f = __hiloint2double((2045 - eb) << 20, 0);
  double* q = base + (size_t)(op.node - T) * nodeStride + i0 + lane;
  for (int ks = warp; ks < K * S; ks += GEN_WARPS)
    q[(size_t)ks * Npad] = out[(size_t)ks * LDT + lane] * f;
  if (warp == 0) expo[((size_t)d * I + (op.node - T)) * Npad + i0 + lane] = (int16_t)(eb - 1022);
}

__global__ void __launch_bounds__(ROOT_THREADS)
gen_root_kernel(const double* __restrict__ partials, const int16_t* __restrict__ expo,
                const double* __restrict__ freqs, int freqDraws,
                const double* __restrict__ props, int propDraws,
                const double* __restrict__ weights, double* __restrict__ siteLnl,
                double* __restrict__ blockPart, int T, int Npad, int K, int S, int rootInode) {
  __shared__ double red[ROOT_THREADS / 32];
  const int d = blockIdx.y;
  const int I = T - 1;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const double* fr = freqs + (freqDraws > 1 ? (size_t)d * S : 0);
  const double* pr = props + (propDraws > 1 ? (size_t)d * K : 0);
  double contrib = 0.0;
  if (i < Npad) {
    const size_t plane = (size_t)S * Npad;
    const double* p = partials + ((size_t)d * I + rootInode) * K * plane + i;
    double L = 0.0;
    for (int k = 0; k < K; ++k) {
      double dot = 0.0;
      for (int s = 0; s < S; ++s) dot = fma(fr[s], p[k * plane + (size_t)s * Npad], dot);
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

__global__ void __launch_bounds__(ROOT_THREADS)
gen_root_bwd_kernel(const double* __restrict__ partials, const int16_t* __restrict__ expo,
                    const double* __restrict__ freqs, int freqDraws,
                    const double* __restrict__ props, int propDraws,
                    const double* __restrict__ weights, double* __restrict__ pre,
                    double* __restrict__ blockPart, int T, int Npad, int K, int S,
                    int rootInode) {
  __shared__ double red[ROOT_THREADS / 32];
  const int d = blockIdx.y;
  const int I = T - 1;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const double* fr = freqs + (freqDraws > 1 ? (size_t)d * S : 0);
  const double* pr = props + (propDraws > 1 ? (size_t)d * K : 0);
  const size_t plane = (size_t)S * Npad;
  const bool live = i < Npad;
  const double* p = partials + ((size_t)d * I + rootInode) * K * plane + (live ? i : 0);
  double w = 0.0, invL = 0.0;
  if (live) {
    double L = 0.0;
    for (int k = 0; k < K; ++k) {
      double dot = 0.0;
      for (int s = 0; s < S; ++s) dot = fma(fr[s], p[k * plane + (size_t)s * Npad], dot);
      L = fma(pr[k], dot, L);
    }
    invL = 1.0 / L;
    w = weights[i];
    const int e = expo[((size_t)d * I + rootInode) * Npad + i];
    const double scale = invL * __hiloint2double((1023 - e) << 20, 0);
    double* q = pre + ((size_t)d * I + rootInode) * K * plane + i;
    for (int k = 0; k < K; ++k) {
      const double c = pr[k] * scale;
      for (int s = 0; s < S; ++s) q[k * plane + (size_t)s * Npad] = c * fr[s];
    }
  }
  const double wl = (live && w != 0.0) ? w * invL : 0.0;
  double* out = blockPart + ((size_t)d * gridDim.x + blockIdx.x) * (K + S);
  for (int k = 0; k < K; ++k) {
    double dot = 0.0;
    if (wl != 0.0)
      for (int s = 0; s < S; ++s) dot = fma(fr[s], p[k * plane + (size_t)s * Npad], dot);
    const double t = block_sum(wl * dot, red);
    if (threadIdx.x == 0) out[k] = t;
  }
  for (int s = 0; s < S; ++s) {
    double acc = 0.0;
    if (wl != 0.0)
      for (int k = 0; k < K; ++k) acc = fma(pr[k], p[k * plane + (size_t)s * Npad], acc);
    const double t = block_sum(wl * acc, red);
    if (threadIdx.x == 0) out[K + s] = t;
  }
}

// pre-order: grid (pattern chunks, nodes x K, draws); a CTA walks the 32-pattern
// tiles of its chunk.  Every thread owns S*S/256 entries of G_l and G_r in
// registers, so no cross-thread reduction is needed.
// shared (doubles): P_l P_r [SS each] | q vl vr ul ur ml mr [S*LDT each]
constexpr int GEN_MAX_OWN = 16;  // ceil(64*64/256)

__global__ void __launch_bounds__(GEN_THREADS)
gen_bwd_kernel(const NodeOp* __restrict__ ops, int opBegin, const double* __restrict__ mats,
               const uint8_t* __restrict__ tips, const double* __restrict__ codeP,
               const double* __restrict__ partials, const int16_t* __restrict__ expo,
               const double* __restrict__ weights, double* __restrict__ pre,
               double* __restrict__ gpart, const int* __restrict__ chunkBase,
               size_t chunkTotal, int T, int Npad, int B, int K, int S, int chunkPatterns,
               int nChunk) {
  extern __shared__ double sm[];
  const int SS = S * S;
  double* Pl = sm;
  double* Pr = Pl + SS;
  double* tq = Pr + SS;
  double* vl = tq + S * LDT;
  double* vr = vl + S * LDT;
  double* ul = vr + S * LDT;
  double* ur = ul + S * LDT;
  double* ml = ur + S * LDT;   // w * m_l after the q^ update
  double* mr = ml + S * LDT;

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
  stage_matrix(Pl, matsD + ((size_t)op.left * K + k) * SS, SS);
  stage_matrix(Pr, matsD + ((size_t)op.right * K + k) * SS, SS);

  double gl[GEN_MAX_OWN], gr[GEN_MAX_OWN];
#pragma unroll
  for (int j = 0; j < GEN_MAX_OWN; ++j) gl[j] = gr[j] = 0.0;

  const int begin = blockIdx.x * chunkPatterns;
  int end = begin + chunkPatterns;
  end = end < Npad ? end : Npad;
  for (int i0 = begin; i0 < end; i0 += TP) {
    __syncthreads();
    stage_child(tq, false, nullptr, codeP,
                pre + drawBase + (size_t)(op.node - T) * nodeStride + k * plane, i0, Npad, S);
    stage_child(vl, tipL, tips + (size_t)(tipL ? op.left : 0) * Npad, codeP,
                partials + drawBase + (size_t)(tipL ? 0 : op.left - T) * nodeStride + k * plane,
                i0, Npad, S);
    stage_child(vr, tipR, tips + (size_t)(tipR ? op.right : 0) * Npad, codeP,
                partials + drawBase + (size_t)(tipR ? 0 : op.right - T) * nodeStride + k * plane,
                i0, Npad, S);
    __syncthreads();
    for (int s = warp; s < S; s += GEN_WARPS) {
      double a = 0.0, b = 0.0;
      const double* rl = Pl + s * S;
      const double* rr = Pr + s * S;
      for (int t = 0; t < S; ++t) {
        a = fma(rl[t], vl[t * LDT + lane], a);
        b = fma(rr[t], vr[t * LDT + lane], b);
      }
      ul[s * LDT + lane] = a;
      ur[s * LDT + lane] = b;
    }
    __syncthreads();
    for (int s = warp; s < S; s += GEN_WARPS) {
      const double q = tq[s * LDT + lane];
      ml[s * LDT + lane] = q * ur[s * LDT + lane];
      mr[s * LDT + lane] = q * ul[s * LDT + lane];
    }
    __syncthreads();
    const int i = i0 + lane;
    if (!tipL) {
      const int e = expo[((size_t)d * I + (op.left - T)) * Npad + i];
      const double f = __hiloint2double((1023 - e) << 20, 0);
      double* q = pre + drawBase + (size_t)(op.left - T) * nodeStride + k * plane + i;
      for (int t = warp; t < S; t += GEN_WARPS) {
        double acc = 0.0;
        for (int s = 0; s < S; ++s) acc = fma(Pl[s * S + t], ml[s * LDT + lane], acc);
        q[(size_t)t * Npad] = acc * f;
      }
    }
    if (!tipR) {
      const int e = expo[((size_t)d * I + (op.right - T)) * Npad + i];
      const double f = __hiloint2double((1023 - e) << 20, 0);
      double* q = pre + drawBase + (size_t)(op.right - T) * nodeStride + k * plane + i;
      for (int t = warp; t < S; t += GEN_WARPS) {
        double acc = 0.0;
        for (int s = 0; s < S; ++s) acc = fma(Pr[s * S + t], mr[s * LDT + lane], acc);
        q[(size_t)t * Npad] = acc * f;
      }
    }
    __syncthreads();
    // fold the pattern weight into m (each thread scales its own rows)
    {
      const double w = weights[i];
      for (int s = warp; s < S; s += GEN_WARPS) {
        ml[s * LDT + lane] *= w;
        mr[s * LDT + lane] *= w;
      }
    }
    __syncthreads();
    // G_c[s][t] += sum_lane (w m_c[s]) * p~_c[t]
#pragma unroll
    for (int j = 0; j < GEN_MAX_OWN; ++j) {
      const int idx = threadIdx.x + j * GEN_THREADS;
      if (idx < SS) {
        const int s = idx / S, t = idx - s * S;
        const double* a = ml + s * LDT;
        const double* b = vl + t * LDT;
        const double* c = mr + s * LDT;
        const double* e2 = vr + t * LDT;
        double accl = gl[j], accr = gr[j];
#pragma unroll 8
        for (int p = 0; p < TP; ++p) {
          accl = fma(a[p], b[p], accl);
          accr = fma(c[p], e2[p], accr);
        }
        gl[j] = accl;
        gr[j] = accr;
      }
    }
  }
  double* outL = gpart + ((size_t)d * chunkTotal + chunkBase[op.left] + (size_t)k * nChunk +
                          blockIdx.x) * SS;
  double* outR = gpart + ((size_t)d * chunkTotal + chunkBase[op.right] + (size_t)k * nChunk +
                          blockIdx.x) * SS;
#pragma unroll
  for (int j = 0; j < GEN_MAX_OWN; ++j) {
    const int idx = threadIdx.x + j * GEN_THREADS;
    if (idx < SS) {
      outL[idx] = gl[j];
      outR[idx] = gr[j];
    }
  }
}

size_t fwd_smem(const Dims& m) {
  return (2 * (size_t)m.S * m.S + 2 * (size_t)m.S * LDT + (size_t)m.K * m.S * LDT +
          GEN_WARPS * 32) * sizeof(double);
}

size_t bwd_smem(const Dims& m) {
  return (2 * (size_t)m.S * m.S + 7 * (size_t)m.S * LDT) * sizeof(double);
}

}  // namespace

int gen_forward(Engine& e, int draws) {
  const Dims& m = e.dm;
  if (m.S > 64) {
    set_error("generic kernels support at most 64 states");
    return TTB2_E_INVALID;
  }
  const size_t smem = fwd_smem(m);
  if (smem > 227 * 1024) {
    set_error("K*S too large for the generic forward kernel's shared-memory tile");
    return TTB2_E_INVALID;
  }
  if (smem > 48 * 1024)
    TTB2_CUDA_CHECK(cudaFuncSetAttribute(gen_fwd_kernel,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int nLevels = (int)e.levelOff.size() - 1;
  for (int l = 0; l < nLevels; ++l) {
    const int opBegin = e.levelOff[l];
    const int count = e.levelOff[l + 1] - opBegin;
    for (int done = 0; done < count; done += 65535) {
      const int c = (count - done) < 65535 ? (count - done) : 65535;
      dim3 grid(m.Npad / TP, c, draws);
      gen_fwd_kernel<<<grid, GEN_THREADS, smem, e.stream>>>(
          e.ops, opBegin + done, e.mats, e.tips, e.codeP, e.partials, e.expo, m.T, m.Npad,
          m.B, m.K, m.S);
      ++e.launches;
    }
  }
  TTB2_CUDA_CHECK(cudaGetLastError());
  return TTB2_OK;
}

int gen_root(Engine& e, int draws) {
  const Dims& m = e.dm;
  const int nblocks = (m.Npad + ROOT_THREADS - 1) / ROOT_THREADS;
  dim3 grid(nblocks, draws);
  const int rootInode = e.hostOps.back().node - m.T;
  gen_root_kernel<<<grid, ROOT_THREADS, 0, e.stream>>>(
      e.partials, e.expo, e.freqs, e.freqDraws, e.props, e.propDraws, e.weights, e.siteLnl,
      e.redPart, m.T, m.Npad, m.K, m.S, rootInode);
  ++e.launches;
  TTB2_CUDA_CHECK(cudaGetLastError());
  return small_reduce_lnl(e, draws, nblocks);
}

int gen_backward(Engine& e, int draws) {
  const Dims& m = e.dm;
  const int rootInode = e.hostOps.back().node - m.T;
  {
    const int nblocks = (m.Npad + ROOT_THREADS - 1) / ROOT_THREADS;
    dim3 grid(nblocks, draws);
    gen_root_bwd_kernel<<<grid, ROOT_THREADS, 0, e.stream>>>(
        e.partials, e.expo, e.freqs, e.freqDraws, e.props, e.propDraws, e.weights, e.pre,
        e.redPart, m.T, m.Npad, m.K, m.S, rootInode);
    ++e.launches;
    TTB2_CUDA_CHECK(cudaGetLastError());
    int rc = small_root_grad_reduce(e, draws, nblocks);
    if (rc) return rc;
  }
  const size_t smem = bwd_smem(m);
  if (smem > 48 * 1024)
    TTB2_CUDA_CHECK(cudaFuncSetAttribute(gen_bwd_kernel,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int nLevels = (int)e.levelOff.size() - 1;
  const int maxNodes = 65535 / m.K;
  for (int l = nLevels - 1; l >= 0; --l) {
    const int opBegin = e.levelOff[l];
    const int count = e.levelOff[l + 1] - opBegin;
    const int nChunk = e.levelChunks[l];
    int chunkPatterns = (m.Npad + nChunk - 1) / nChunk;
    chunkPatterns = (chunkPatterns + TP - 1) / TP * TP;
    for (int done = 0; done < count; done += maxNodes) {
      const int c = (count - done) < maxNodes ? (count - done) : maxNodes;
      dim3 grid(nChunk, c * m.K, draws);
      gen_bwd_kernel<<<grid, GEN_THREADS, smem, e.stream>>>(
          e.ops, opBegin + done, e.mats, e.tips, e.codeP, e.partials, e.expo, e.weights,
          e.pre, e.gpart, e.chunkBase, e.chunkTotal, m.T, m.Npad, m.B, m.K, m.S, chunkPatterns,
          nChunk);
      ++e.launches;
    }
  }
  TTB2_CUDA_CHECK(cudaGetLastError());
  return small_gpart_reduce(e, draws);
}

}  // namespace ttb2
