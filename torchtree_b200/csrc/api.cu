// C ABI of the ttb200 engine (see include/ttb200.h): engine life-cycle, level
// schedule, input/output staging and the per-evaluation launch sequences.
#include <algorithm>
#include <cstring>
#include <new>

#include <nvtx3/nvToolsExt.h>

#include "engine.cuh"

namespace ttb2 {

// NVTX ranges around the phases of an evaluation (visible in Nsight Systems / Compute; a no-op
// without an attached tool): ttb2:loglik {pmatrix, postorder, root}, ttb2:grad {preorder, contract}
struct NvtxRange {
  explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
  ~NvtxRange() { nvtxRangePop(); }
};

static thread_local std::string g_last_error;

void set_error(const std::string& msg) { g_last_error = msg; }

namespace {

template <typename T>
int dev_alloc(Engine& e, T** ptr, size_t count) {
  *ptr = nullptr;
  if (count == 0) count = 1;
  cudaError_t err = cudaMalloc((void**)ptr, count * sizeof(T));
  if (err != cudaSuccess) {
    set_error(std::string("cudaMalloc of ") + std::to_string(count * sizeof(T)) +
              " bytes failed: " + cudaGetErrorString(err));
    return TTB2_E_NOMEM;
  }
  e.deviceBytes += (int64_t)(count * sizeof(T));
  return TTB2_OK;
}

template <typename T>
void dev_free(T*& ptr) {
  if (ptr) cudaFree(ptr);
  ptr = nullptr;
}

int build_schedule(Engine& e, const int32_t* post) {
  const Dims& m = e.dm;
  std::vector<int> level(2 * m.T - 1, -1);
  for (int t = 0; t < m.T; ++t) level[t] = 0;
  std::vector<NodeOp> opsv(m.I);
  std::vector<char> isChild(2 * m.T - 1, 0);
  for (int j = 0; j < m.I; ++j) {
    const int node = post[3 * j], l = post[3 * j + 1], r = post[3 * j + 2];
    if (node < m.T || node > 2 * m.T - 2 || l < 0 || r < 0 || l > 2 * m.T - 2 ||
        r > 2 * m.T - 2 || l == r) {
      set_error("postorder: node index out of range at triple " + std::to_string(j));
      return TTB2_E_INVALID;
    }
    if (level[node] != -1) {
      set_error("postorder: node " + std::to_string(node) + " appears twice");
      return TTB2_E_INVALID;
    }
    if (level[l] < 0 || level[r] < 0) {
      set_error("postorder: child visited after its parent at triple " + std::to_string(j));
      return TTB2_E_INVALID;
    }
    if (isChild[l] || isChild[r]) {
      set_error("postorder: a node has two parents at triple " + std::to_string(j));
      return TTB2_E_INVALID;
    }
    isChild[l] = isChild[r] = 1;
    level[node] = 1 + std::max(level[l], level[r]);
    opsv[j] = NodeOp{node, l, r, level[node]};
  }
  // the last triple must be the root: the only node that is nobody's child
  for (int n = 0; n < 2 * m.T - 1; ++n) {
    const bool isRoot = (n == post[3 * (m.I - 1)]);
    if (!isChild[n] && !isRoot) {
      set_error("postorder: node " + std::to_string(n) + " is not connected to the root");
      return TTB2_E_INVALID;
    }
  }
  const NodeOp rootOp = opsv.back();
  std::stable_sort(opsv.begin(), opsv.end(),
                   [](const NodeOp& a, const NodeOp& b) { return a.level < b.level; });
  if (opsv.back().node != rootOp.node) {
    set_error("postorder: last triple is not the root");
    return TTB2_E_INVALID;
  }
  e.hostOps = opsv;
  e.levelOff.clear();
  int cur = 0;
  for (int j = 0; j < m.I; ++j) {
    while (cur < opsv[j].level) {
      e.levelOff.push_back(j);
      ++cur;
    }
  }
  e.levelOff.push_back(m.I);
  // levelOff[l-1]..levelOff[l] = ops of level l; drop the leading dummy for level 0
  // (cur started at 0, the first push is for level 1)
  return TTB2_OK;
}

int copy_in(Engine& e, void* dst, const void* src, size_t bytes, int where) {
  if (bytes == 0) return TTB2_OK;
  TTB2_CUDA_CHECK(cudaMemcpyAsync(dst, src, bytes,
                                  where == TTB2_HOST ? cudaMemcpyHostToDevice
                                                     : cudaMemcpyDeviceToDevice,
                                  e.stream));
  return TTB2_OK;
}

int copy_out(Engine& e, void* dst, const void* src, size_t bytes, int where) {
  if (bytes == 0 || dst == nullptr) return TTB2_OK;
  TTB2_CUDA_CHECK(cudaMemcpyAsync(dst, src, bytes,
                                  where == TTB2_HOST ? cudaMemcpyDeviceToHost
                                                     : cudaMemcpyDeviceToDevice,
                                  e.stream));
  return TTB2_OK;
}

int finish(Engine& e, int where) {
  if (where == TTB2_HOST) TTB2_CUDA_CHECK(cudaStreamSynchronize(e.stream));
  return TTB2_OK;
}

bool bad_draws(int d, int draws) { return !(d == 1 || d == draws); }

// The small inputs of an evaluation are staged back to back in one device buffer,
//   [ bl[D][B] | rates[rd][K] | props[pd][K] | freqs[fd][S] | q_norm[qd][S][S] ],
// so that host inputs cost ONE host-to-device copy (through a pinned mirror; a pageable
// cudaMemcpyAsync per array is a driver-staged, synchronous copy each) and device inputs ONE
// gather kernel instead of five copy launches -- at the 8-GPU shard of the headline problem the
// per-call fixed costs are what limits scaling.
struct InSeg {
  const double* src;
  double** slot;
  size_t n;
};

int stage_inputs(Engine& e, const InSeg* segs, int nseg, int where) {
  size_t off[8], total = 0;
  for (int j = 0; j < nseg; ++j) {
    off[j] = total;
    *segs[j].slot = e.inPacked + total;
    total += segs[j].n;
  }
  if (total > e.inCap) {
    set_error("internal: staged inputs exceed the staging buffer");
    return TTB2_E_INVALID;
  }
  if (where == TTB2_HOST) {
    for (int j = 0; j < nseg; ++j)
      std::memcpy(e.hostIn + off[j], segs[j].src, segs[j].n * sizeof(double));
    TTB2_CUDA_CHECK(cudaMemcpyAsync(e.inPacked, e.hostIn, total * sizeof(double),
                                    cudaMemcpyHostToDevice, e.stream));
    return TTB2_OK;
  }
  const double* src[8];
  size_t n[8];
  for (int j = 0; j < nseg; ++j) {
    src[j] = segs[j].src;
    n[j] = segs[j].n;
  }
  return small_gather_inputs(e, src, off, n, nseg);
}

// The small outputs of a gradient call live back to back in one buffer,
//   [ lnL[D] | d_bl[D][B] | d_rates[rd][K] | d_props[pd][K] | d_q[ed][S][S] | d_freqs[fd][S] ],
// so that a caller that reduces them across shards (NCCL all-reduce over NVLink) or copies them
// to the host moves ONE contiguous vector (ttb2_grad_eigen_packed).  The layout is a function
// of the draw counts of the latest loglik call (the key of the captured CUDA graphs as well).
void layout_outputs(Engine& e, int draws) {
  const Dims& m = e.dm;
  size_t off = (size_t)draws;
  e.outBl = e.outPacked + off;     off += (size_t)draws * m.B;
  e.outRates = e.outPacked + off;  off += (size_t)e.rateDraws * m.K;
  e.outProps = e.outPacked + off;  off += (size_t)e.propDraws * m.K;
  e.outQ = e.outPacked + off;      off += (size_t)e.eigDraws * m.S * m.S;
  e.outFreqs = e.outPacked + off;  off += (size_t)e.freqDraws * m.S;
  e.packedCount = (int64_t)off;
}

void drop_graphs(Engine& e);

int ensure_grad_buffers(Engine& e, int draws) {
  const Dims& m = e.dm;
  int rc;
  if (!e.pre) {
    const size_t n = (size_t)e.cfg.max_draws * m.I * m.K * m.Npad * m.S;
    if ((rc = dev_alloc(e, &e.pre, n))) return rc;
  }
  // codon path: keep the u vectors of the post-order sweep from the next evaluation on (a third
  // buffer of the size of the partials: only when it leaves half of the free memory alone)
  if (!e.uTried && gmma_keeps_u(e)) {
    e.uTried = true;
    const size_t n = (size_t)e.cfg.max_draws * m.I * m.K * m.Npad * m.S;
    size_t freeB = 0, totalB = 0;
    if (cudaMemGetInfo(&freeB, &totalB) == cudaSuccess && n * sizeof(double) <= freeB / 2) {
      if ((rc = dev_alloc(e, &e.ustore, n))) return rc;
      drop_graphs(e);
    }
  }
  const size_t matN = (size_t)e.cfg.max_draws * m.B * m.K * m.S * m.S;
  if (!e.dmat && (rc = dev_alloc(e, &e.dmat, matN))) return rc;
  if (!e.rootGrad && (rc = dev_alloc(e, &e.rootGrad, (size_t)e.cfg.max_draws * (m.K + m.S))))
    return rc;
  const int planBefore = e.chunkPlanDraws;
  // CTAs per SM aimed at per pre-order launch: many short CTAs shorten the last wave of the
  // big levels, but a CTA needs a few hundred patterns to amortise its prologue -- measured
  // optimum on the 1000-taxon problem: 8 at 12.5k patterns, 16 at 25k, 32 from 50k up
  // (large alphabets run one CTA per SM and pay ~4 us per CTA for staging two S x S matrices:
  // 8 per SM measured best on the 61-state config, 18.3 ms vs 18.6 at 16 and 19.2 at 32)
  int perSm = e.dm.S <= 32 ? 32 : 8;
  if (gwarp_supported(e, true)) perSm = 16;   // warp-autonomous 20-state kernels: 9.8 ms vs 10.0 at 32
  if (e.spec4) perSm = e.dm.Npad < 20000 ? 8 : (e.dm.Npad < 40000 ? 16 : 32);
  // CTAs co-resident per SM of the pre-order kernel, for whole-wave launches of the compute-bound
  // DMMA kernels (0: memory-bound 4-state path and SIMT fallback, where a partial last wave simply
  // gets more bandwidth per CTA)
  int resident = 0;
  if (!e.spec4 && gmma_supported(e)) resident = gwarp_supported(e, true) ? 3 : 1;
  // (the codon path's tip-tip level has its own kernel, two CTAs per SM: kernels_gmma.cu)
  const int residentLevel1 = (resident == 1 && e.dm.S == 61 && e.cfg.code_count == e.dm.S + 1) ? 2 : 0;
  if ((rc = plan_chunks(e, draws, e.spec4 ? 128 : 32, perSm, resident, residentLevel1))) return rc;
  if (planBefore != e.chunkPlanDraws) drop_graphs(e);
  const size_t need = planned_gpart_doubles(e, draws);
  if (need > e.gpartCap) {
    drop_graphs(e);
    if (e.gpart) {
      e.deviceBytes -= (int64_t)(e.gpartCap * sizeof(double));
      dev_free(e.gpart);
    }
    if ((rc = dev_alloc(e, &e.gpart, need))) return rc;
    e.gpartCap = need;
  }
  return TTB2_OK;
}

int ensure_eigen_grad_buffers(Engine& e) {
  const Dims& m = e.dm;
  int rc;
  const size_t matN = (size_t)e.cfg.max_draws * m.B * m.K * m.S * m.S;
  if (!e.hpart) {
    if ((rc = dev_alloc(e, &e.hpart, matN))) return rc;
    e.hpartCap = matN;
  }
  if (!e.gscal && (rc = dev_alloc(e, &e.gscal, (size_t)e.cfg.max_draws * m.B * m.K))) return rc;
  if (!e.hred && (rc = dev_alloc(e, &e.hred, (size_t)e.cfg.max_draws * 8 * m.S * m.S))) return rc;
  return TTB2_OK;
}

int refresh_topology_tables(Engine& e) { return s4_build_cherries(e); }

int run_forward(Engine& e, int draws, double* lnl, int where) {
  int rc;
  mark(e, 1);
  const int64_t before = e.launches;
  if (e.spec4) {
    if ((rc = s4_forward(e, draws))) return rc;
    e.fwdLevelLaunches = (int)(e.launches - before);
    mark(e, 2);
    if ((rc = s4_root(e, draws))) return rc;
  } else if (gmma_supported(e)) {
    if (!e.expoK && (rc = dev_alloc(e, &e.expoK, gmma_expo_elems(e)))) return rc;
    if ((rc = gmma_forward2(e, draws))) return rc;
    e.fwdLevelLaunches = (int)(e.launches - before);
    mark(e, 2);
    if ((rc = gmma_root2(e, draws))) return rc;
  } else {
    if ((rc = gen_forward(e, draws))) return rc;
    e.fwdLevelLaunches = (int)(e.launches - before);
    mark(e, 2);
    if ((rc = gen_root(e, draws))) return rc;
  }
  mark(e, 3);
  e.draws = draws;
  e.preValid = false;
  if ((rc = copy_out(e, lnl, e.lnl, (size_t)draws * sizeof(double), where))) return rc;
  return finish(e, where);
}

// ---- CUDA-graph replay ----------------------------------------------------
void drop_graphs(Engine& e) {
  if (e.gFwd.exec) cudaGraphExecDestroy(e.gFwd.exec);
  if (e.gBwd.exec) cudaGraphExecDestroy(e.gBwd.exec);
  e.gFwd = Engine::GraphSlot();
  e.gBwd = Engine::GraphSlot();
}

// (bench.py --patterns 12500 / 25000, profiles/r01_small_shard.log: replay gains 3 % / 1.6 % there,
// i.e. on the 8- / 4-GPU shards of the headline problem, so the limit covers those shards:
// 1.2e8 unit-equivalents; TTB2_GRAPH_MAX_UNITS overrides it)
// Graph replay pays when an evaluation is launch-bound (fluA: 79 launches of a few
// microseconds each, 0.53 -> 0.37 ms).  When the level kernels run for milliseconds the
// host is far ahead of the device anyway and replay only adds its own launch and
// event-ordering cost (measured +0.1 ms on 1000 taxa x 100k patterns), so large
// problems use ordinary stream launches.
bool graphs_enabled(const Engine& e, int draws) {
  if ((e.cfg.flags & TTB2_FLAG_NO_GRAPH) || e.timing || e.ownStream == nullptr) return false;
  const double units = (double)e.dm.Npad * e.dm.I * e.dm.K * draws * (e.dm.S / 4.0);
  static const double maxUnits =
      getenv("TTB2_GRAPH_MAX_UNITS") ? atof(getenv("TTB2_GRAPH_MAX_UNITS")) : 1.2e8;
  return units <= maxUnits;
}

bool slot_matches(const Engine::GraphSlot& g, const Engine& e, int draws) {
  return g.exec && g.draws == draws && g.fd == e.freqDraws && g.pd == e.propDraws &&
         g.rd == e.rateDraws && g.ed == e.eigDraws && g.qd == e.qDraws && g.us == e.uValid;
}

// Runs `body` (a sequence of kernel launches on e.stream) through a cached CUDA
// graph: captured on the engine's own stream the first time a shape is seen, then
// replayed.  The user's stream is ordered before and after with events.
template <typename Body>
int run_graphed(Engine& e, Engine::GraphSlot& slot, int draws, Body body) {
  if (!graphs_enabled(e, draws)) return body();
  cudaStream_t user = e.stream;
  if (!slot_matches(slot, e, draws)) {
    if (slot.exec) cudaGraphExecDestroy(slot.exec);
    slot = Engine::GraphSlot();
    const int64_t before = e.launches;
    TTB2_CUDA_CHECK(cudaStreamBeginCapture(e.ownStream, cudaStreamCaptureModeThreadLocal));
    e.stream = e.ownStream;
    const int rc = body();
    e.stream = user;
    cudaGraph_t graph = nullptr;
    const cudaError_t err = cudaStreamEndCapture(e.ownStream, &graph);
    if (rc) {
      if (graph) cudaGraphDestroy(graph);
      return rc;
    }
    if (err != cudaSuccess || !graph) {
      set_error(std::string("CUDA graph capture failed: ") + cudaGetErrorString(err));
      return TTB2_E_CUDA;
    }
    const cudaError_t ierr = cudaGraphInstantiate(&slot.exec, graph, 0);
    cudaGraphDestroy(graph);
    if (ierr != cudaSuccess) {
      slot.exec = nullptr;
      set_error(std::string("cudaGraphInstantiate: ") + cudaGetErrorString(ierr));
      return TTB2_E_CUDA;
    }
    slot.kernels = e.launches - before;
    e.launches = before;  // counted per replay below
    slot.draws = draws;
    slot.fd = e.freqDraws; slot.pd = e.propDraws; slot.rd = e.rateDraws; slot.ed = e.eigDraws;
    slot.qd = e.qDraws;
    slot.us = e.uValid;
  }
  TTB2_CUDA_CHECK(cudaEventRecord(e.evIn, user));
  TTB2_CUDA_CHECK(cudaStreamWaitEvent(e.ownStream, e.evIn, 0));
  TTB2_CUDA_CHECK(cudaGraphLaunch(slot.exec, e.ownStream));
  TTB2_CUDA_CHECK(cudaEventRecord(e.evOut, e.ownStream));
  TTB2_CUDA_CHECK(cudaStreamWaitEvent(user, e.evOut, 0));
  e.launches += slot.kernels;
  return TTB2_OK;
}

int stage_grad_lnl(Engine& e, const double* grad_lnl, int where) {
  if (grad_lnl)
    return copy_in(e, e.gradLnl, grad_lnl, (size_t)e.draws * sizeof(double), where);
  TTB2_CUDA_CHECK(cudaMemcpyAsync(e.gradLnl, e.ones, e.draws * sizeof(double),
                                  cudaMemcpyDeviceToDevice, e.stream));
  return TTB2_OK;
}

int run_backward(Engine& e, const double* grad_lnl, int where) {
  int rc;
  const int draws = e.draws;
  if ((rc = ensure_grad_buffers(e, draws))) return rc;
  if (grad_lnl) {
    if ((rc = copy_in(e, e.gradLnl, grad_lnl, (size_t)draws * sizeof(double), where))) return rc;
  } else {
    TTB2_CUDA_CHECK(cudaMemcpyAsync(e.gradLnl, e.ones, draws * sizeof(double),
                                    cudaMemcpyDeviceToDevice, e.stream));
  }
  mark(e, 4);
  if (!e.preValid) {
    const int64_t before = e.launches;
    rc = e.spec4 ? s4_backward(e, draws)
                 : (gmma_supported(e) ? gmma_backward2(e, draws) : gen_backward(e, draws));
    if (rc) return rc;
    e.bwdLevelLaunches = (int)(e.launches - before) - 3;  // minus root, root reduce, gpart reduce
    e.preValid = true;
  }
  mark(e, 5);
  return TTB2_OK;
}

}  // namespace
}  // namespace ttb2

using namespace ttb2;

extern "C" {

int ttb2_version(void) { return TTB2_VERSION; }

const char* ttb2_last_error(void) { return g_last_error.c_str(); }

int ttb2_create(const ttb2_config* config, const uint8_t* tip_codes,
                const double* code_partials, const double* weights, const int32_t* postorder,
                ttb2_engine** out) {
  if (!config || !tip_codes || !code_partials || !weights || !postorder || !out) {
    set_error("ttb2_create: null argument");
    return TTB2_E_INVALID;
  }
  *out = nullptr;
  const ttb2_config& c = *config;
  if (c.tip_count < 2 || c.pattern_count < 1 || c.state_count < 2 || c.category_count < 1 ||
      c.max_draws < 1 || c.code_count < c.state_count + 1 || c.code_count > 255) {
    set_error("ttb2_create: invalid configuration (need T>=2, N>=1, S>=2, K>=1, D>=1, "
              "S+1<=C<=255)");
    return TTB2_E_INVALID;
  }
  if (c.state_count > 64) {
    set_error("ttb2_create: state_count > 64 is not supported");
    return TTB2_E_INVALID;
  }
  // code table contract: unit vectors then an all-ones row
  for (int r = 0; r <= c.state_count; ++r)
    for (int s = 0; s < c.state_count; ++s) {
      const double want = (r == c.state_count || r == s) ? 1.0 : 0.0;
      if (code_partials[(size_t)r * c.state_count + s] != want) {
        set_error("ttb2_create: code_partials rows 0..S-1 must be unit vectors and row S all ones");
        return TTB2_E_INVALID;
      }
    }
  const size_t TN = (size_t)c.tip_count * c.pattern_count;
  for (size_t j = 0; j < TN; ++j)
    if (tip_codes[j] >= c.code_count) {
      set_error("ttb2_create: tip code out of range of code_partials");
      return TTB2_E_INVALID;
    }
  int ndev = 0;
  cudaError_t err = cudaGetDeviceCount(&ndev);
  if (err != cudaSuccess || ndev == 0) {
    set_error(std::string("ttb2_create: no CUDA device available (") +
              cudaGetErrorString(err) + "); this engine has no CPU fallback");
    return TTB2_E_CUDA;
  }
  if (c.device < 0 || c.device >= ndev) {
    set_error("ttb2_create: device ordinal out of range");
    return TTB2_E_INVALID;
  }
  TTB2_CUDA_CHECK(cudaSetDevice(c.device));

  Engine* ep = new (std::nothrow) Engine();
  if (!ep) {
    set_error("out of host memory");
    return TTB2_E_NOMEM;
  }
  Engine& e = *ep;
  e.cfg = c;
  e.device = c.device;
  Dims& m = e.dm;
  m.T = c.tip_count;
  m.I = m.T - 1;
  m.B = 2 * m.T - 2;
  m.N = c.pattern_count;
  m.Npad = (m.N + 31) / 32 * 32;
  m.S = c.state_count;
  m.K = c.category_count;
  m.C = c.code_count;
  e.spec4 = (m.S == 4) && !(c.flags & TTB2_FLAG_FORCE_GENERIC);

  int rc = build_schedule(e, postorder);
  if (rc) {
    delete ep;
    return rc;
  }

#define TRY(x)            \
  do {                    \
    rc = (x);             \
    if (rc) {             \
      ttb2_destroy(reinterpret_cast<ttb2_engine*>(ep)); \
      return rc;          \
    }                     \
  } while (0)
#define TRY_CUDA(x)                                                     \
  do {                                                                  \
    cudaError_t e__ = (x);                                              \
    if (e__ != cudaSuccess) {                                           \
      set_error(std::string(#x) + ": " + cudaGetErrorString(e__));      \
      ttb2_destroy(reinterpret_cast<ttb2_engine*>(ep));                 \
      return TTB2_E_CUDA;                                               \
    }                                                                   \
  } while (0)

  const int D = c.max_draws;
  TRY(dev_alloc(e, &e.tips, (size_t)m.T * m.Npad));
  TRY(dev_alloc(e, &e.weights, (size_t)m.Npad));
  TRY(dev_alloc(e, &e.codeP, (size_t)m.C * m.S));
  TRY(dev_alloc(e, &e.ops, (size_t)m.I));
  TRY(dev_alloc(e, &e.partials, (size_t)D * m.I * m.K * m.Npad * m.S));
  TRY(dev_alloc(e, &e.expo, (size_t)D * m.I * m.Npad));
  TRY(dev_alloc(e, &e.mats, (size_t)D * m.B * m.K * m.S * m.S));
  TRY(dev_alloc(e, &e.siteLnl, (size_t)D * m.Npad));
  const int nblocks = m.Npad / 32;   // the root kernels write one row of partial sums per 32 patterns
  e.redPartCap = (size_t)D * nblocks * (m.K + m.S);
  TRY(dev_alloc(e, &e.redPart, e.redPartCap));
  TRY(dev_alloc(e, &e.lnl, (size_t)D));
  e.inCap = (size_t)D * (m.B + 2 * m.K + m.S + m.S * m.S);
  TRY(dev_alloc(e, &e.inPacked, e.inCap));
  TRY_CUDA(cudaHostAlloc((void**)&e.hostIn, e.inCap * sizeof(double), cudaHostAllocDefault));
  TRY(dev_alloc(e, &e.evec, (size_t)D * m.S * m.S));
  TRY(dev_alloc(e, &e.ivec, (size_t)D * m.S * m.S));
  TRY(dev_alloc(e, &e.eval, (size_t)D * m.S));
  TRY(dev_alloc(e, &e.gradLnl, (size_t)D));
  TRY(dev_alloc(e, &e.ones, (size_t)D));
  {
    std::vector<double> ones(D, 1.0);
    TRY_CUDA(cudaMemcpy(e.ones, ones.data(), D * sizeof(double), cudaMemcpyHostToDevice));
  }
  TRY(dev_alloc(e, &e.outPacked, (size_t)D * (1 + m.B + 2 * m.K + m.S * m.S + m.S)));
  e.rateDraws = e.propDraws = e.eigDraws = e.freqDraws = 1;
  layout_outputs(e, 1);

  // tips, padded with the all-ones code S; weights padded with 0
  {
    std::vector<uint8_t> padded((size_t)m.T * m.Npad, (uint8_t)m.S);
    for (int t = 0; t < m.T; ++t)
      std::memcpy(&padded[(size_t)t * m.Npad], tip_codes + (size_t)t * m.N, m.N);
    TRY_CUDA(cudaMemcpy(e.tips, padded.data(), padded.size(), cudaMemcpyHostToDevice));
    std::vector<double> w(m.Npad, 0.0);
    std::memcpy(w.data(), weights, m.N * sizeof(double));
    TRY_CUDA(cudaMemcpy(e.weights, w.data(), w.size() * sizeof(double), cudaMemcpyHostToDevice));
  }
  TRY_CUDA(cudaMemcpy(e.codeP, code_partials, (size_t)m.C * m.S * sizeof(double),
                      cudaMemcpyHostToDevice));
  {
    std::vector<int> masks(m.C, 0);
    e.codes01 = m.S <= 32;
    for (int cc = 0; cc < m.C; ++cc)
      for (int s2 = 0; s2 < m.S; ++s2) {
        const double v = code_partials[(size_t)cc * m.S + s2];
        if (v != 0.0 && v != 1.0) e.codes01 = false;
        if (v != 0.0 && s2 < 32) masks[cc] |= 1 << s2;
      }
    TRY(dev_alloc(e, &e.codeMask, (size_t)m.C));
    TRY_CUDA(cudaMemcpy(e.codeMask, masks.data(), m.C * sizeof(int), cudaMemcpyHostToDevice));
  }
  TRY_CUDA(cudaMemcpy(e.ops, e.hostOps.data(), m.I * sizeof(NodeOp), cudaMemcpyHostToDevice));
  {
    int sms = 0;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c.device) == cudaSuccess &&
        sms > 0)
      e.smCount = sms;
  }
  TRY(refresh_topology_tables(e));
  TRY_CUDA(cudaStreamCreateWithFlags(&e.ownStream, cudaStreamNonBlocking));
  TRY_CUDA(cudaEventCreateWithFlags(&e.evIn, cudaEventDisableTiming));
  TRY_CUDA(cudaEventCreateWithFlags(&e.evOut, cudaEventDisableTiming));
  if (c.flags & TTB2_FLAG_PREALLOC_GRAD) {
    TRY(ensure_grad_buffers(e, D));
    TRY(ensure_eigen_grad_buffers(e));
  }
#undef TRY
#undef TRY_CUDA
  *out = reinterpret_cast<ttb2_engine*>(ep);
  return TTB2_OK;
}

int ttb2_set_postorder(ttb2_engine* engine, const int32_t* postorder) {
  if (!engine || !postorder) {
    set_error("ttb2_set_postorder: null argument");
    return TTB2_E_INVALID;
  }
  Engine& e = *reinterpret_cast<Engine*>(engine);
  TTB2_CUDA_CHECK(cudaSetDevice(e.device));
  std::vector<NodeOp> oldOps = e.hostOps;
  std::vector<int> oldOff = e.levelOff;
  int rc = build_schedule(e, postorder);
  if (rc) {
    e.hostOps = oldOps;
    e.levelOff = oldOff;
    return rc;
  }
  TTB2_CUDA_CHECK(cudaStreamSynchronize(e.stream));
  TTB2_CUDA_CHECK(cudaMemcpy(e.ops, e.hostOps.data(), e.dm.I * sizeof(NodeOp),
                             cudaMemcpyHostToDevice));
  e.mode = MODE_NONE;
  e.preValid = false;
  e.chunkPlanDraws = 0;
  ++e.evalSerial;   // a deferred gradient must not be served from the old topology's buffers
  drop_graphs(e);
  return refresh_topology_tables(e);
}

void ttb2_destroy(ttb2_engine* engine) {
  if (!engine) return;
  Engine* ep = reinterpret_cast<Engine*>(engine);
  Engine& e = *ep;
  cudaSetDevice(e.device);
  cudaStreamSynchronize(e.stream);
  dev_free(e.tips); dev_free(e.weights); dev_free(e.codeP); dev_free(e.codeMask); dev_free(e.ops);
  dev_free(e.partials); dev_free(e.expo); dev_free(e.pre); dev_free(e.ustore); dev_free(e.mats);
  dev_free(e.dmat); dev_free(e.gpart); dev_free(e.siteLnl); dev_free(e.redPart);
  dev_free(e.lnl); dev_free(e.rootGrad); dev_free(e.hpart); dev_free(e.gscal); dev_free(e.hred);
  dev_free(e.inPacked);
  if (e.hostIn) cudaFreeHost(e.hostIn);
  e.hostIn = nullptr;
  dev_free(e.evec); dev_free(e.ivec); dev_free(e.eval); dev_free(e.gradLnl); dev_free(e.ones);
  dev_free(e.outPacked);
  dev_free(e.expoK);
  dev_free(e.chunkBase); dev_free(e.chunkCount);
  dev_free(e.cherryIdx); dev_free(e.cherryInfo); dev_free(e.cherryVec); dev_free(e.cherryExp);
  dev_free(e.cherryCode);
  for (int j = 0; j < 8; ++j)
    if (e.ev[j]) cudaEventDestroy(e.ev[j]);
  drop_graphs(e);
  if (e.evIn) cudaEventDestroy(e.evIn);
  if (e.evOut) cudaEventDestroy(e.evOut);
  if (e.ownStream) cudaStreamDestroy(e.ownStream);
  delete ep;
}

int ttb2_set_stream(ttb2_engine* engine, void* cuda_stream) {
  if (!engine) {
    set_error("ttb2_set_stream: null engine");
    return TTB2_E_INVALID;
  }
  Engine& e = *reinterpret_cast<Engine*>(engine);
  cudaStream_t next = reinterpret_cast<cudaStream_t>(cuda_stream);
  if (next == e.stream) return TTB2_OK;
  TTB2_CUDA_CHECK(cudaSetDevice(e.device));
  // order the new stream behind the work already queued on the old one (no host sync)
  TTB2_CUDA_CHECK(cudaEventRecord(e.evIn, e.stream));
  TTB2_CUDA_CHECK(cudaStreamWaitEvent(next, e.evIn, 0));
  e.stream = next;
  return TTB2_OK;
}

int ttb2_synchronize(ttb2_engine* engine) {
  if (!engine) {
    set_error("ttb2_synchronize: null engine");
    return TTB2_E_INVALID;
  }
  Engine& e = *reinterpret_cast<Engine*>(engine);
  TTB2_CUDA_CHECK(cudaSetDevice(e.device));
  TTB2_CUDA_CHECK(cudaStreamSynchronize(e.stream));
  return TTB2_OK;
}

int ttb2_loglik_mats(ttb2_engine* engine, int32_t draws, const double* mats,
                     const double* freqs, int32_t freq_draws, const double* props,
                     int32_t prop_draws, double* lnl, int32_t where) {
  if (!engine || !mats || !freqs || !props) {
    set_error("ttb2_loglik_mats: null argument");
    return TTB2_E_INVALID;
  }
  Engine& e = *reinterpret_cast<Engine*>(engine);
  const Dims& m = e.dm;
  if (draws < 1 || draws > e.cfg.max_draws || bad_draws(freq_draws, draws) ||
      bad_draws(prop_draws, draws)) {
    set_error("ttb2_loglik_mats: draws out of range (or a *_draws that is neither 1 nor draws)");
    return TTB2_E_INVALID;
  }
  TTB2_CUDA_CHECK(cudaSetDevice(e.device));
  int rc;
  mark(e, 0);
  if ((rc = copy_in(e, e.mats, mats, (size_t)draws * m.B * m.K * m.S * m.S * sizeof(double), where)))
    return rc;
  {
    const InSeg segs[] = {{props, &e.props, (size_t)prop_draws * m.K},
                          {freqs, &e.freqs, (size_t)freq_draws * m.S}};
    if ((rc = stage_inputs(e, segs, 2, where))) return rc;
  }
  e.freqDraws = freq_draws;
  e.propDraws = prop_draws;
  e.rateDraws = e.eigDraws = 1;
  e.mode = MODE_MATS;
  ++e.evalSerial;
  layout_outputs(e, draws);
  NvtxRange range("ttb2:loglik_mats");
  return run_forward(e, draws, lnl, where);
}

// Shared tail of the two eigen-mode entry points: inputs are staged; `qDraws` > 0 asks for
// the eigen-system to be computed on the device from e.qnorm / e.freqs first.
static int run_eigen_mode(Engine& e, int draws, int qDraws, double* lnl, int where) {
  int rc;
  e.mode = MODE_EIGEN;
  ++e.evalSerial;
  e.draws = draws;
  e.qDraws = qDraws;
  layout_outputs(e, draws);
  NvtxRange range("ttb2:loglik_eigen");
  if (!e.spec4 && !gmma_supported(e)) {
    if (qDraws && (rc = small_sym_eigh(e, qDraws, e.eigDraws))) return rc;
    if ((rc = small_pmatrix(e, draws))) return rc;
    return run_forward(e, draws, lnl, where);
  }
  if (!e.spec4 && !e.expoK && (rc = dev_alloc(e, &e.expoK, gmma_expo_elems(e)))) return rc;
  rc = run_graphed(e, e.gFwd, draws, [&]() -> int {
    int r;
    if (qDraws && (r = small_sym_eigh(e, qDraws, e.eigDraws))) return r;
    if ((r = small_pmatrix(e, draws))) return r;
    return run_forward(e, draws, nullptr, TTB2_DEVICE);
  });
  if (rc) return rc;
  e.draws = draws;
  e.preValid = false;
  if ((rc = copy_out(e, lnl, e.lnl, (size_t)draws * sizeof(double), where))) return rc;
  return finish(e, where);
}

int ttb2_loglik_eigen(ttb2_engine* engine, int32_t draws, const double* branch_lengths,
                      const double* site_rates, int32_t rate_draws, const double* props,
                      int32_t prop_draws, const double* evec, const double* ivec,
                      const double* eval, int32_t eig_draws, const double* freqs,
                      int32_t freq_draws, double* lnl, int32_t where) {
  if (!engine || !branch_lengths || !site_rates || !props || !evec || !ivec || !eval || !freqs) {
    set_error("ttb2_loglik_eigen: null argument");
    return TTB2_E_INVALID;
  }
  Engine& e = *reinterpret_cast<Engine*>(engine);
  const Dims& m = e.dm;
  if (draws < 1 || draws > e.cfg.max_draws || bad_draws(freq_draws, draws) ||
      bad_draws(prop_draws, draws) || bad_draws(rate_draws, draws) ||
      bad_draws(eig_draws, draws)) {
    set_error("ttb2_loglik_eigen: draws out of range (or a *_draws that is neither 1 nor draws)");
    return TTB2_E_INVALID;
  }
  TTB2_CUDA_CHECK(cudaSetDevice(e.device));
  int rc;
  mark(e, 0);
  const size_t SS = (size_t)m.S * m.S;
  {
    const InSeg segs[] = {{branch_lengths, &e.bl, (size_t)draws * m.B},
                          {site_rates, &e.rates, (size_t)rate_draws * m.K},
                          {props, &e.props, (size_t)prop_draws * m.K},
                          {freqs, &e.freqs, (size_t)freq_draws * m.S}};
    if ((rc = stage_inputs(e, segs, 4, where))) return rc;
  }
  if ((rc = copy_in(e, e.evec, evec, eig_draws * SS * sizeof(double), where))) return rc;
  if ((rc = copy_in(e, e.ivec, ivec, eig_draws * SS * sizeof(double), where))) return rc;
  if ((rc = copy_in(e, e.eval, eval, (size_t)eig_draws * m.S * sizeof(double), where))) return rc;
  e.freqDraws = freq_draws;
  e.propDraws = prop_draws;
  e.rateDraws = rate_draws;
  e.eigDraws = eig_draws;
  return run_eigen_mode(e, draws, 0, lnl, where);
}

int ttb2_loglik_q(ttb2_engine* engine, int32_t draws, const double* branch_lengths,
                  const double* site_rates, int32_t rate_draws, const double* props,
                  int32_t prop_draws, const double* q_norm, int32_t q_draws, const double* freqs,
                  int32_t freq_draws, double* lnl, int32_t where) {
  if (!engine || !branch_lengths || !site_rates || !props || !q_norm || !freqs) {
    set_error("ttb2_loglik_q: null argument");
    return TTB2_E_INVALID;
  }
  Engine& e = *reinterpret_cast<Engine*>(engine);
  const Dims& m = e.dm;
  if (draws < 1 || draws > e.cfg.max_draws || bad_draws(freq_draws, draws) ||
      bad_draws(prop_draws, draws) || bad_draws(rate_draws, draws) || bad_draws(q_draws, draws)) {
    set_error("ttb2_loglik_q: draws out of range (or a *_draws that is neither 1 nor draws)");
    return TTB2_E_INVALID;
  }
  if (m.S > 64) {
    set_error("ttb2_loglik_q: the device eigen-decomposition supports at most 64 states; "
              "decompose on the host and call ttb2_loglik_eigen");
    return TTB2_E_INVALID;
  }
  TTB2_CUDA_CHECK(cudaSetDevice(e.device));
  int rc;
  mark(e, 0);
  const size_t SS = (size_t)m.S * m.S;
  {
    const InSeg segs[] = {{branch_lengths, &e.bl, (size_t)draws * m.B},
                          {site_rates, &e.rates, (size_t)rate_draws * m.K},
                          {props, &e.props, (size_t)prop_draws * m.K},
                          {freqs, &e.freqs, (size_t)freq_draws * m.S},
                          {q_norm, &e.qnorm, (size_t)q_draws * SS}};
    if ((rc = stage_inputs(e, segs, 5, where))) return rc;
  }
  e.freqDraws = freq_draws;
  e.propDraws = prop_draws;
  e.rateDraws = rate_draws;
  // one eigen-system per generator draw or per frequency draw, whichever varies
  e.eigDraws = q_draws > freq_draws ? q_draws : freq_draws;
  return run_eigen_mode(e, draws, q_draws, lnl, where);
}

int ttb2_loglik_expm(ttb2_engine* engine, int32_t draws, const double* branch_lengths,
                     const double* site_rates, int32_t rate_draws, const double* props,
                     int32_t prop_draws, const double* q, int32_t q_draws, const double* freqs,
                     int32_t freq_draws, double* lnl, int32_t where) {
  if (!engine || !branch_lengths || !site_rates || !props || !q || !freqs) {
    set_error("ttb2_loglik_expm: null argument");
    return TTB2_E_INVALID;
  }
  Engine& e = *reinterpret_cast<Engine*>(engine);
  const Dims& m = e.dm;
  if (draws < 1 || draws > e.cfg.max_draws || bad_draws(freq_draws, draws) ||
      bad_draws(prop_draws, draws) || bad_draws(rate_draws, draws) || bad_draws(q_draws, draws)) {
    set_error("ttb2_loglik_expm: draws out of range (or a *_draws that is neither 1 nor draws)");
    return TTB2_E_INVALID;
  }
  TTB2_CUDA_CHECK(cudaSetDevice(e.device));
  int rc;
  mark(e, 0);
  const size_t SS = (size_t)m.S * m.S;
  {
    const InSeg segs[] = {{branch_lengths, &e.bl, (size_t)draws * m.B},
                          {site_rates, &e.rates, (size_t)rate_draws * m.K},
                          {props, &e.props, (size_t)prop_draws * m.K},
                          {freqs, &e.freqs, (size_t)freq_draws * m.S},
                          {q, &e.qnorm, (size_t)q_draws * SS}};
    if ((rc = stage_inputs(e, segs, 5, where))) return rc;
  }
  e.freqDraws = freq_draws;
  e.propDraws = prop_draws;
  e.rateDraws = rate_draws;
  e.eigDraws = q_draws;   // leading extent of d_q
  e.qDraws = q_draws;
  e.mode = MODE_EXPM;
  ++e.evalSerial;
  e.draws = draws;
  layout_outputs(e, draws);
  NvtxRange range("ttb2:loglik_expm");
  if ((rc = small_expm_forward(e, draws))) return rc;
  return run_forward(e, draws, lnl, where);
}

int ttb2_get_eigen(ttb2_engine* engine, double* evec, double* ivec, double* eval, int32_t where) {
  if (!engine) {
    set_error("ttb2_get_eigen: null engine");
    return TTB2_E_INVALID;
  }
  Engine& e = *reinterpret_cast<Engine*>(engine);
  if (e.mode != MODE_EIGEN) {
    set_error("ttb2_get_eigen: the latest evaluation was not an eigen-mode one");
    return TTB2_E_STATE;
  }
  const Dims& m = e.dm;
  TTB2_CUDA_CHECK(cudaSetDevice(e.device));
  int rc;
  const size_t SS = (size_t)m.S * m.S;
  if (evec && (rc = copy_out(e, evec, e.evec, e.eigDraws * SS * sizeof(double), where))) return rc;
  if (ivec && (rc = copy_out(e, ivec, e.ivec, e.eigDraws * SS * sizeof(double), where))) return rc;
  if (eval && (rc = copy_out(e, eval, e.eval, (size_t)e.eigDraws * m.S * sizeof(double), where)))
    return rc;
  return finish(e, where);
}

int ttb2_grad_mats(ttb2_engine* engine, const double* grad_lnl, double* d_mats,
                   double* d_freqs, double* d_props, int32_t where) {
  if (!engine) {
    set_error("ttb2_grad_mats: null engine");
    return TTB2_E_INVALID;
  }
  Engine& e = *reinterpret_cast<Engine*>(engine);
  const Dims& m = e.dm;
  if (e.mode == MODE_NONE) {
    set_error("ttb2_grad_mats: no log-likelihood has been evaluated yet");
    return TTB2_E_STATE;
  }
  TTB2_CUDA_CHECK(cudaSetDevice(e.device));
  int rc;
  if ((rc = run_backward(e, grad_lnl, where))) return rc;
  const int draws = e.draws;
  const size_t matBytes = (size_t)draws * m.B * m.K * m.S * m.S * sizeof(double);
  if (d_mats) {
    if (where == TTB2_DEVICE) {
      if ((rc = small_scale_dmat(e, draws, d_mats))) return rc;
    } else {
      // scale into the hpart scratch (same shape as dmat), then copy out
      if ((rc = ensure_eigen_grad_buffers(e))) return rc;
      if ((rc = small_scale_dmat(e, draws, e.hpart))) return rc;
      if ((rc = copy_out(e, d_mats, e.hpart, matBytes, where))) return rc;
    }
  }
  if (d_freqs || d_props) {
    if ((rc = small_root_outputs(e, draws))) return rc;
    if ((rc = copy_out(e, d_props, e.outProps, (size_t)e.propDraws * m.K * sizeof(double), where)))
      return rc;
    if ((rc = copy_out(e, d_freqs, e.outFreqs, (size_t)e.freqDraws * m.S * sizeof(double), where)))
      return rc;
  }
  return finish(e, where);
}

// pre-order sweep + contraction of the latest eigen-mode evaluation: fills the packed outputs
static int grad_eigen_compute(Engine& e, const double* grad_lnl, int where) {
  int rc;
  const int draws = e.draws;
  NvtxRange range("ttb2:grad_eigen");
  if (e.mode == MODE_EXPM) {
    // general generator: pre-order sweep -> d lnL / d P, then the Frechet adjoint of exp
    if ((rc = ensure_eigen_grad_buffers(e))) return rc;
    if ((rc = run_backward(e, grad_lnl, where))) return rc;
    return small_expm_contract(e, draws);
  }
  // the contraction kernel reduces the per-chunk sums of G itself (one launch less)
  struct Defer {
    Engine& e;
    explicit Defer(Engine& e_) : e(e_) { e.deferGpart = true; }
    ~Defer() { e.deferGpart = false; }
  } defer(e);
  if (e.preValid || !graphs_enabled(e, draws) || !(e.spec4 || gmma_supported(e))) {
    if ((rc = ensure_eigen_grad_buffers(e))) return rc;
    if ((rc = run_backward(e, grad_lnl, where))) return rc;
    if ((rc = small_eigen_contract(e, draws))) return rc;
  } else {
    if ((rc = ensure_eigen_grad_buffers(e))) return rc;
    if ((rc = ensure_grad_buffers(e, draws))) return rc;  // allocations + chunk plan: not capturable
    if ((rc = stage_grad_lnl(e, grad_lnl, where))) return rc;
    rc = run_graphed(e, e.gBwd, draws, [&]() -> int {
      int r = e.spec4 ? s4_backward(e, draws) : gmma_backward2(e, draws);
      if (r) return r;
      return small_eigen_contract(e, draws);
    });
    if (rc) return rc;
    e.preValid = true;
  }
  return TTB2_OK;
}

int ttb2_grad_eigen(ttb2_engine* engine, const double* grad_lnl, double* d_branch_lengths,
                    double* d_site_rates, double* d_props, double* d_q, double* d_freqs,
                    int32_t where) {
  if (!engine) {
    set_error("ttb2_grad_eigen: null engine");
    return TTB2_E_INVALID;
  }
  Engine& e = *reinterpret_cast<Engine*>(engine);
  const Dims& m = e.dm;
  if (e.mode != MODE_EIGEN && e.mode != MODE_EXPM) {
    set_error("ttb2_grad_eigen: the latest evaluation was not ttb2_loglik_eigen / _q / _expm");
    return TTB2_E_STATE;
  }
  TTB2_CUDA_CHECK(cudaSetDevice(e.device));
  int rc;
  const int draws = e.draws;
  if ((rc = grad_eigen_compute(e, grad_lnl, where))) return rc;
  if ((rc = copy_out(e, d_branch_lengths, e.outBl, (size_t)draws * m.B * sizeof(double), where)))
    return rc;
  if ((rc = copy_out(e, d_site_rates, e.outRates, (size_t)e.rateDraws * m.K * sizeof(double), where)))
    return rc;
  if ((rc = copy_out(e, d_props, e.outProps, (size_t)e.propDraws * m.K * sizeof(double), where)))
    return rc;
  if ((rc = copy_out(e, d_q, e.outQ, (size_t)e.eigDraws * m.S * m.S * sizeof(double), where)))
    return rc;
  if ((rc = copy_out(e, d_freqs, e.outFreqs, (size_t)e.freqDraws * m.S * sizeof(double), where)))
    return rc;
  mark(e, 6);
  return finish(e, where);
}

int64_t ttb2_packed_count(const ttb2_engine* engine) {
  return engine ? reinterpret_cast<const Engine*>(engine)->packedCount : 0;
}

int ttb2_grad_eigen_packed(ttb2_engine* engine, const double* grad_lnl, double* packed,
                           int64_t capacity, int32_t where) {
  if (!engine || !packed) {
    set_error("ttb2_grad_eigen_packed: null argument");
    return TTB2_E_INVALID;
  }
  Engine& e = *reinterpret_cast<Engine*>(engine);
  if (e.mode != MODE_EIGEN && e.mode != MODE_EXPM) {
    set_error("ttb2_grad_eigen_packed: the latest evaluation was not ttb2_loglik_eigen / _q / _expm");
    return TTB2_E_STATE;
  }
  if (capacity < e.packedCount) {
    set_error("ttb2_grad_eigen_packed: output buffer holds " + std::to_string(capacity) +
              " doubles, " + std::to_string(e.packedCount) + " needed (ttb2_packed_count)");
    return TTB2_E_INVALID;
  }
  TTB2_CUDA_CHECK(cudaSetDevice(e.device));
  int rc;
  if ((rc = grad_eigen_compute(e, grad_lnl, where))) return rc;
  TTB2_CUDA_CHECK(cudaMemcpyAsync(e.outPacked, e.lnl, (size_t)e.draws * sizeof(double),
                                  cudaMemcpyDeviceToDevice, e.stream));
  if ((rc = copy_out(e, packed, e.outPacked, (size_t)e.packedCount * sizeof(double), where)))
    return rc;
  mark(e, 6);
  return finish(e, where);
}

int ttb2_enable_timing(ttb2_engine* engine, int32_t on) {
  if (!engine) {
    set_error("ttb2_enable_timing: null engine");
    return TTB2_E_INVALID;
  }
  Engine& e = *reinterpret_cast<Engine*>(engine);
  TTB2_CUDA_CHECK(cudaSetDevice(e.device));
  if (on) {
    for (int j = 0; j < 8; ++j)
      if (!e.ev[j]) TTB2_CUDA_CHECK(cudaEventCreate(&e.ev[j]));
  }
  for (int j = 0; j < 8; ++j) e.evSet[j] = false;
  e.timing = on != 0;
  return TTB2_OK;
}

int ttb2_phase_ms(ttb2_engine* engine, double* out) {
  if (!engine || !out) {
    set_error("ttb2_phase_ms: null argument");
    return TTB2_E_INVALID;
  }
  Engine& e = *reinterpret_cast<Engine*>(engine);
  TTB2_CUDA_CHECK(cudaSetDevice(e.device));
  TTB2_CUDA_CHECK(cudaStreamSynchronize(e.stream));
  const int pairs[5][2] = {{0, 1}, {1, 2}, {2, 3}, {4, 5}, {5, 6}};
  for (int j = 0; j < 5; ++j) {
    out[j] = 0.0;
    const int a = pairs[j][0], b = pairs[j][1];
    if (e.ev[a] && e.ev[b] && e.evSet[a] && e.evSet[b]) {
      float ms = 0.f;
      if (cudaEventElapsedTime(&ms, e.ev[a], e.ev[b]) == cudaSuccess) out[j] = ms;
    }
  }
  out[5] = e.fwdLevelLaunches;
  out[6] = e.bwdLevelLaunches;
  return TTB2_OK;
}

int ttb2_site_loglik(ttb2_engine* engine, double* out, int32_t where) {
  if (!engine || !out) {
    set_error("ttb2_site_loglik: null argument");
    return TTB2_E_INVALID;
  }
  Engine& e = *reinterpret_cast<Engine*>(engine);
  if (e.mode == MODE_NONE) {
    set_error("ttb2_site_loglik: no log-likelihood has been evaluated yet");
    return TTB2_E_STATE;
  }
  TTB2_CUDA_CHECK(cudaSetDevice(e.device));
  const Dims& m = e.dm;
  TTB2_CUDA_CHECK(cudaMemcpy2DAsync(
      out, (size_t)m.N * sizeof(double), e.siteLnl, (size_t)m.Npad * sizeof(double),
      (size_t)m.N * sizeof(double), e.draws,
      where == TTB2_HOST ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice, e.stream));
  return finish(e, where);
}

int ttb2_get_mats(ttb2_engine* engine, double* out, int32_t where) {
  if (!engine || !out) {
    set_error("ttb2_get_mats: null argument");
    return TTB2_E_INVALID;
  }
  Engine& e = *reinterpret_cast<Engine*>(engine);
  if (e.mode == MODE_NONE) {
    set_error("ttb2_get_mats: no log-likelihood has been evaluated yet");
    return TTB2_E_STATE;
  }
  TTB2_CUDA_CHECK(cudaSetDevice(e.device));
  const Dims& m = e.dm;
  int rc = copy_out(e, out, e.mats, (size_t)e.draws * m.B * m.K * m.S * m.S * sizeof(double), where);
  if (rc) return rc;
  return finish(e, where);
}

int64_t ttb2_launch_count(const ttb2_engine* engine) {
  return engine ? reinterpret_cast<const Engine*>(engine)->launches : 0;
}

int64_t ttb2_device_bytes(const ttb2_engine* engine) {
  return engine ? reinterpret_cast<const Engine*>(engine)->deviceBytes : 0;
}

int64_t ttb2_eval_serial(const ttb2_engine* engine) {
  return engine ? reinterpret_cast<const Engine*>(engine)->evalSerial : 0;
}

int ttb2_get_config(const ttb2_engine* engine, ttb2_config* out) {
  if (!engine || !out) {
    set_error("ttb2_get_config: null argument");
    return TTB2_E_INVALID;
  }
  *out = reinterpret_cast<const Engine*>(engine)->cfg;
  return TTB2_OK;
}

}  // extern "C"
