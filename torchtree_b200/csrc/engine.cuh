// Internal declarations shared by the ttb200 translation units.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdlib>
#include <string>
#include <vector>

#include "../../include/ttb200.h"

namespace ttb2 {

// One internal node of the traversal; ops are stored level by level
// (level = height above the tips; nodes of one level are independent).
struct NodeOp {
  int32_t node, left, right, level;
};

struct Dims {
  int T;     // tips
  int I;     // internal nodes T-1
  int B;     // branches 2T-2
  int N;     // patterns
  int Npad;  // patterns padded to a multiple of 32 (padding: all-gap, weight 0)
  int S;     // states
  int K;     // rate categories
  int C;     // tip symbol codes
};

enum Mode { MODE_NONE = 0, MODE_MATS = 1, MODE_EIGEN = 2, MODE_EXPM = 3 };

struct Engine {
  ttb2_config cfg{};
  Dims dm{};
  int device = 0;
  cudaStream_t stream = nullptr;
  bool spec4 = false;  // S==4 specialised kernels (AoS [k][N][4] layout)

  // static data
  uint8_t* tips = nullptr;      // [T][Npad]
  double* weights = nullptr;    // [Npad]
  double* codeP = nullptr;      // [C][S]
  int* codeMask = nullptr;      // [C] bit s set iff codeP[c][s] != 0 (S <= 32)
  bool codes01 = false;         // every entry of codeP is exactly 0 or 1
  NodeOp* ops = nullptr;        // [I] level-sorted
  std::vector<NodeOp> hostOps;
  std::vector<int> levelOff;    // ops of level l (1-based height): [levelOff[l-1], levelOff[l])

  // per-evaluation state (sized for max_draws)
  double* partials = nullptr;   // spec4: [D][I][K][Npad][4]; generic: [D][I][K][S][Npad]
  int16_t* expo = nullptr;      // [D][I][Npad] power-of-two scale exponents
  double* pre = nullptr;        // pre-order vectors, same layout as partials
  // codon path: u_c = P_c p_c of every internal child, written by the post-order sweep and read by
  // the pre-order sweep (same layout as partials, indexed by the child).  Allocated with the
  // gradient buffers when it fits; uValid: the last post-order sweep filled it.
  double* ustore = nullptr;
  bool uValid = false, uTried = false;
  double* mats = nullptr;       // [D][B][K][S][S]
  double* dmat = nullptr;       // d lnL / d mats, same shape
  double* gpart = nullptr;      // per-chunk partial sums of dmat
  size_t gpartCap = 0;          // doubles
  // pattern-chunk plan of the per-level pre-order kernels (depends on draws)
  std::vector<int> levelChunks;     // chunks per level
  std::vector<int> hostChunkBase;   // [B] offset (in chunk slots) of branch b, category 0
  std::vector<int> hostChunkCount;  // [B] chunks of branch b
  int* chunkBase = nullptr;
  int* chunkCount = nullptr;
  int chunkPlanDraws = 0;           // draws the uploaded plan was made for (0 = none)
  size_t chunkTotal = 0;            // chunk slots per draw (sum over branches of K * chunks)
  double* siteLnl = nullptr;    // [D][Npad]
  double* redPart = nullptr;    // per-block partials of pattern reductions
  size_t redPartCap = 0;
  double* lnl = nullptr;        // [D]
  double* rootGrad = nullptr;   // [D][K+S] (d_props | d_freqs) per draw
  double* hpart = nullptr;      // [D][B*K][S][S] (M o Phi) per branch x category
  double* gscal = nullptr;      // [D][B*K] d lnL / d (r t)
  double* hred = nullptr;       // [D][8][S][S] slice sums of hpart (h_reduce_kernel -> q_grad_kernel)
  bool deferGpart = false;      // the caller's eigen contraction reduces the G chunks itself
  bool gpartPending = false;    // ... and this sweep left them unreduced
  // staged inputs: slices of one staging buffer, laid out per call (api.cu stage_inputs)
  double* inPacked = nullptr;   // [bl | rates | props | freqs | q_norm]
  double* hostIn = nullptr;     // pinned host mirror (one H2D copy for host inputs)
  size_t inCap = 0;             // doubles
  double* freqs = nullptr;      // [freqDraws][S]
  double* props = nullptr;      // [propDraws][K]
  double* bl = nullptr;         // [draws][B]
  double* rates = nullptr;      // [rateDraws][K]
  double* qnorm = nullptr;      // [qDraws][S][S] generator for the device eigen-decomposition
  double* evec = nullptr;       // [Dmax][S][S] (supplied, or computed by sym_eigh_kernel)
  double* ivec = nullptr;
  double* eval = nullptr;       // [Dmax][S]
  double* gradLnl = nullptr;    // [Dmax]
  double* ones = nullptr;       // [Dmax] default grad_lnl
  // staged outputs (small): slices of one buffer, laid out per call (api.cu layout_outputs)
  double* outPacked = nullptr;  // [lnL | d_bl | d_rates | d_props | d_q | d_freqs]
  int64_t packedCount = 0;      // doubles in use for the latest call's draw counts
  double* outBl = nullptr;      // [draws][B]
  double* outRates = nullptr;   // [rateDraws][K]
  double* outProps = nullptr;   // [propDraws][K]
  double* outFreqs = nullptr;   // [freqDraws][S]
  double* outQ = nullptr;       // [eigDraws][S][S]

  // cherry fusion (kernels_s4.cu): level-1 nodes tabulated by tip-code pair
  bool cherryOn = false;
  int nCherry = 0;
  int* cherryIdx = nullptr;     // [I]
  int* cherryInfo = nullptr;    // [nCherry][3] (left tip, right tip, node)
  double* cherryVec = nullptr;  // [D][nCherry][K][C*C][4]
  int* cherryExp = nullptr;     // [D][nCherry][C*C]
  bool smemAttrTma = false, smemAttrTips = false;  // opt-in shared-memory sizes set on this device
  uint8_t* cherryCode = nullptr;  // [nCherry][Npad] pair code = code(left tip) * C + code(right tip)

  int smCount = 148;
  int16_t* expoK = nullptr;     // [D][I][K][Npad] per-(pattern, category) exponents (DMMA paths)
  size_t hpartCap = 0;

  int draws = 0, freqDraws = 0, propDraws = 0, rateDraws = 0, eigDraws = 0;
  Mode mode = MODE_NONE;
  bool preValid = false;

  int64_t launches = 0;
  int qDraws = 0;  // generator draws of the latest ttb2_loglik_q call (0: eigen-system supplied)
  int64_t evalSerial = 0;  // loglik calls so far (ttb2_eval_serial)
  int64_t deviceBytes = 0;

  // CUDA-graph replay of the eigen-mode kernel sequences (api.cu)
  struct GraphSlot {
    cudaGraphExec_t exec = nullptr;
    int draws = -1, fd = -1, pd = -1, rd = -1, ed = -1, qd = -1;
    int64_t kernels = 0;
    bool us = false;   // captured with the kept u vectors (Engine::uValid)
  };
  GraphSlot gFwd, gBwd;
  cudaStream_t ownStream = nullptr;   // capture / replay stream
  cudaEvent_t evIn = nullptr, evOut = nullptr;

  // optional phase timing (ttb2_enable_timing)
  bool timing = false;
  cudaEvent_t ev[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  bool evSet[8] = {false, false, false, false, false, false, false, false};
  int fwdLevelLaunches = 0, bwdLevelLaunches = 0;
};

// event slots: 0 start-fwd, 1 after pmatrix, 2 after post-order levels, 3 after root,
//              4 start-bwd, 5 after pre-order levels, 6 after contraction
inline void mark(Engine& e, int slot) {
  if (e.timing && e.ev[slot]) {
    cudaEventRecord(e.ev[slot], e.stream);
    e.evSet[slot] = true;
  }
}

void set_error(const std::string& msg);

#define TTB2_CUDA_CHECK(expr)                                                  \
  do {                                                                         \
    cudaError_t err__ = (expr);                                                \
    if (err__ != cudaSuccess) {                                                \
      ttb2::set_error(std::string(#expr) + ": " + cudaGetErrorString(err__) + \
                      " (" + __FILE__ + ":" + std::to_string(__LINE__) + ")"); \
      return TTB2_E_CUDA;                                                      \
    }                                                                          \
  } while (0)

// ---- launchers (each returns a TTB2_* status) ------------------------------
// S = 4 specialised path (kernels_s4.cu)
int s4_build_cherries(Engine& e);
int s4_forward(Engine& e, int draws);
int s4_root(Engine& e, int draws);
int s4_backward(Engine& e, int draws);

// generic-S path (kernels_gen.cu)
int gen_forward(Engine& e, int draws);
int gen_root(Engine& e, int draws);
int gen_backward(Engine& e, int draws);

// fp64 tensor-core path for 8 <= S <= 64 (kernels_gmma.cu); shares the generic layout
bool gmma_supported(const Engine& e);
bool gmma_keeps_u(const Engine& e);   // the post-order sweep can keep u_c = P_c p_c for the pre-order sweep
// per-(pattern, category) rescaling, pipelined staging (own root kernels, expoK layout)
size_t gmma_expo_elems(const Engine& e);
int gmma_forward2(Engine& e, int draws);
// post-order level 1 (both children tips) as a streaming kernel, any 8 <= S <= 64
bool gmma_cherry_level_supported(const Engine& e);
int gmma_cherry_forward_level(Engine& e, int draws);
// pre-order level 1 (tip-tip nodes, unit / gap codes): G-only kernel for S = 20 / 61
bool gmma_cherry_backward_supported(const Engine& e);
int gmma_cherry_backward_level(Engine& e, int draws, bool pdl);
int gmma_root2(Engine& e, int draws);
int gmma_backward2(Engine& e, int draws);
// warp-autonomous DMMA level kernels for 20 states (kernels_gwarp.cu); TTB2_GM_LEGACY=1 keeps
// the shared-memory tile kernels of kernels_gmma.cu
bool gwarp_supported(const Engine& e, bool backward);
int gwarp_forward(Engine& e, int draws);
int gwarp_backward_levels(Engine& e, int draws);

// small kernels (kernels_small.cu)
int small_pmatrix(Engine& e, int draws);
// copies up to 8 device arrays into e.inPacked at the given offsets (one launch)
int small_gather_inputs(Engine& e, const double* const* src, const size_t* off, const size_t* n,
                        int nseg);
// batched Jacobi eigen-decomposition of the staged generators (eigen.cu)
int small_sym_eigh(Engine& e, int qDraws, int eigDraws);
int small_reduce_lnl(Engine& e, int draws, int nblocks);
int small_root_grad_reduce(Engine& e, int draws, int nblocks);
int small_gpart_reduce(Engine& e, int draws);
int small_scale_dmat(Engine& e, int draws, double* out);
int small_eigen_contract(Engine& e, int draws);
int small_root_outputs(Engine& e, int draws);
// matrix-exponential route for general generators (expm.cu)
int small_expm_forward(Engine& e, int draws);
int small_expm_backward(Engine& e, int draws);
int small_expm_contract(Engine& e, int draws);

// Plans how many pattern chunks every level's pre-order launch uses (enough CTAs
// to fill the GPU on small levels, long-lived CTAs on large ones) and uploads the
// per-branch offsets into the partial-sum buffer.  `granule` = patterns a chunk
// must be a multiple of.
int plan_chunks(Engine& e, int draws, int granule, int ctasPerSm, int residentPerSm = 0,
                int residentLevel1 = 0);

// Chunk count near `want` whose launch of items * chunks CTAs fills whole waves of `slots`
// co-resident CTAs: with one or a few CTAs resident per SM (the DMMA kernels) a launch of
// 8.03 waves costs nine -- e.g. 132 (node, category) items x 9 chunks = 1188 CTAs on 148
// single-CTA SMs; x 10 chunks = 1320 = 8.92 waves fills the ninth.
inline long wave_aware_chunks(long items, long want, long maxChunks, long slots) {
  if (slots <= 0 || items <= 0) return want;
  long lo = want * 2 / 3, hi = want * 3 / 2 + 1;
  if (lo < 1) lo = 1;
  if (hi > maxChunks) hi = maxChunks;
  long best = want < 1 ? 1 : (want > maxChunks ? maxChunks : want);
  double bestScore = -1.0;
  for (long c = lo; c <= hi; ++c) {
    const long n = items * c;
    const long waves = (n + slots - 1) / slots;
    const double eff = (double)n / (double)(waves * slots);
    const double dist = want > 0 ? (double)(c > want ? c - want : want - c) / (double)want : 0.0;
    const double score = eff - 0.02 * dist;
    if (score > bestScore) {
      bestScore = score;
      best = c;
    }
  }
  return best;
}
size_t planned_gpart_doubles(const Engine& e, int draws);

// Programmatic dependent launch: a level kernel launched with the
// programmaticStreamSerialization attribute may start (and run its prologue: tables,
// matrix fragments, barrier setup -- nothing the previous level wrote) while the
// previous level's last CTAs drain; pdl_wait() returns once that grid has completed
// and its writes are visible.  The trigger follows the wait, so a kernel can only
// overlap its immediate predecessor.
__device__ __forceinline__ void pdl_wait_then_trigger() {
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

template <typename... KArgs, typename... Args>
inline cudaError_t launch_level(void (*kernel)(KArgs...), dim3 grid, int threads, size_t smem,
                         cudaStream_t stream, bool pdl, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = dim3(threads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = pdl ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}


// Chain runs: maximal runs of >= 2 consecutive levels (above level 1) that hold at most
// TTB2_CHAIN_MAX (default 4) nodes each -- the top of any tree, all of a ladder-like one.
// The 4-state path walks such a run in ONE launch per sweep (every CTA walks the run's
// nodes for its own pattern slice) instead of one small launch per level.
// first[l] = last level of the run starting at l (else -1); last[l] = its first level.
struct ChainRuns {
  std::vector<int> endOfStart, startOfEnd;
};
inline ChainRuns chain_runs(const Engine& e) {
  // (read at every call: the tests switch chain mode on for small problems)
  const int mx = getenv("TTB2_CHAIN_MAX") ? atoi(getenv("TTB2_CHAIN_MAX")) : 4;
  const long minPatterns =
      getenv("TTB2_CHAIN_MIN_PATTERNS") ? atol(getenv("TTB2_CHAIN_MIN_PATTERNS")) : 40000;
  const int nLevels = (int)e.levelOff.size() - 1;
  ChainRuns r;
  r.endOfStart.assign(nLevels, -1);
  r.startOfEnd.assign(nLevels, -1);
  const bool kOk = e.dm.K <= 6 || e.dm.K == 8;   // template instances of the chain kernels
  // a chain trades node-level parallelism for fewer launches: only when the pattern axis alone
  // fills the GPU (measured: a gain from ~40k patterns per GPU, a small loss at 12.5k)
  const bool wide = (long)e.dm.Npad * e.cfg.max_draws >= minPatterns;
  if (mx <= 0 || !e.spec4 || !kOk || !wide) return r;
  int l = 1;
  while (l < nLevels) {
    int j = l;
    while (j < nLevels && e.levelOff[j + 1] - e.levelOff[j] <= mx) ++j;
    if (j - l >= 2) {
      r.endOfStart[l] = j - 1;
      r.startOfEnd[j - 1] = l;
      l = j;
    } else {
      l = j > l ? j : l + 1;
    }
  }
  return r;
}

// TTB2_NO_PDL=1 launches every level kernel with ordinary stream ordering (A/B runs)
inline bool pdl_enabled() {
  static const bool on = getenv("TTB2_NO_PDL") == nullptr;
  return on;
}

}  // namespace ttb2
