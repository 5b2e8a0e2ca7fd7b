/*
 * ttb200 -- B200-native phylogenetic tree-likelihood engine: C ABI.
 *
 * This is the drop-in boundary for the tree-likelihood hot path of
 * 4ment/torchtree.  The reference has no native code and no FFI,
 * so there is no existing binding to mirror symbol-for-symbol; each entry point
 * below names the reference function(s) whose work it takes over
 * (paths relative to the reference checkout):
 *
 *   ttb2_create            TreeLikelihoodModel.__init__
 *                          torchtree/evolution/tree_likelihood.py:283-311
 *                          (tip partials / tip states + pattern weights are
 *                          captured once; SitePattern.compute_tips_* results)
 *   ttb2_set_postorder     TreeModel.postorder / update_traversals
 *                          torchtree/evolution/tree_model.py:187-195
 *   ttb2_loglik_mats       calculate_treelikelihood_discrete(_rescaled) and the
 *                          tip-state variants, tree_likelihood.py:40-75,
 *                          :78-131, :186-221, :224-278 (matrices supplied by the
 *                          caller, e.g. from SubstitutionModel.p_t)
 *   ttb2_loglik_eigen      TreeLikelihoodModel._call, tree_likelihood.py:313-356:
 *                          bls x site rates -> SymmetricSubstitutionModel.p_t
 *                          (substitution_model/abstract.py:57-76) -> peeling
 *   ttb2_grad_mats /       the autograd backward of the above,
 *   ttb2_grad_eigen        replaced by an analytic pre-order pass
 *   ttb2_site_loglik       per-pattern log-likelihoods (the tensor the
 *                          reference reduces at tree_likelihood.py:71-75)
 *
 * Conventions
 *   - all floating point is IEEE fp64; integers are int32 unless stated;
 *   - every function returns 0 on success, a negative TTB2_E_* code otherwise;
 *     ttb2_last_error() returns a message for the calling thread's last error;
 *   - `where` says whether the *data* pointers of that call are host
 *     (TTB2_HOST: pageable or pinned) or device (TTB2_DEVICE) pointers.  Host
 *     outputs are complete when the call returns; device outputs are ordered
 *     on the engine's stream (ttb2_set_stream / ttb2_synchronize);
 *   - the engine never frees caller memory; the caller never frees engine
 *     memory; one engine is used from one host thread at a time;
 *   - shapes use T tips, I=T-1 internal nodes, B=2T-2 branches (branch b is the
 *     edge above node b), N site patterns, K rate categories, S states,
 *     D draws (batch of parameter samples);
 *   - an input that is shared by all draws is passed with its `*_draws`
 *     argument equal to 1; its gradient is then the sum over draws.
 *
 * There is no CPU fallback: every entry point requires a CUDA device.
 */
#ifndef TTB200_H
#define TTB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TTB2_VERSION 200

#define TTB2_HOST 0
#define TTB2_DEVICE 1

#define TTB2_OK 0
#define TTB2_E_INVALID (-1)   /* bad argument / unsupported configuration */
#define TTB2_E_CUDA (-2)      /* CUDA runtime error (see ttb2_last_error) */
#define TTB2_E_STATE (-3)     /* call order violated (e.g. grad before loglik) */
#define TTB2_E_NOMEM (-4)

typedef struct ttb2_engine ttb2_engine;

typedef struct ttb2_config {
  int32_t tip_count;      /* T  >= 2 */
  int32_t pattern_count;  /* N  >= 1 (the patterns owned by this engine/shard) */
  int32_t state_count;    /* S  >= 2 */
  int32_t category_count; /* K  >= 1 */
  int32_t max_draws;      /* D_max >= 1 */
  int32_t code_count;     /* C: rows of code_partials, S+1 <= C <= 255 */
  int32_t device;         /* CUDA device ordinal */
  int32_t flags;          /* TTB2_FLAG_* */
} ttb2_config;

#define TTB2_FLAG_NONE 0
/* keep per-(node,pattern) data for the gradient pass resident from create()
 * (otherwise allocated at the first ttb2_grad_* call) */
#define TTB2_FLAG_PREALLOC_GRAD 1
/* force the generic-S SIMT kernels even when a specialised path exists (testing: an
 * independent implementation of the same sweeps) */
#define TTB2_FLAG_FORCE_GENERIC 2
/* accepted and ignored (round 1 had an experimental fused whole-tree traversal behind it) */
#define TTB2_FLAG_FUSED 4
/* 8..64 states: plain fp64 FMA kernels instead of the fp64 tensor-core (DMMA) kernels
 * (testing / comparison); ignored by the 4-state path */
#define TTB2_FLAG_NO_MMA 8
/* accepted and ignored: cherry tabulation (below) used to be opt-in */
#define TTB2_FLAG_CHERRY 16
/* do not replay the per-evaluation kernel sequence from a CUDA graph (the default
 * captures the eigen-mode forward and backward sequences once per shape and
 * replays them: one graph launch instead of ~25-80 kernel launches) */
#define TTB2_FLAG_NO_GRAPH 32
/* 4-state models tabulate "cherries" (nodes whose two children are tips) by tip-code
 * pair instead of storing their vectors per pattern whenever the pair alphabet is
 * small (C*C <= 64): on a random tree a third of the internal nodes are cherries, so a
 * third of the post-order traffic and a ninth of the pre-order reads never touch HBM
 * (config 2: 10.8 -> 9.8 ms per evaluation).  This flag stores them like any other node. */
#define TTB2_FLAG_NO_CHERRY 64

/*
 * tip_codes      uint8 [T][N]: symbol code of tip t at pattern i
 * code_partials  double [C][S]: tip conditional-likelihood vector of each code.
 *                Rows 0..S-1 must be the unit vectors (unambiguous states) and
 *                row S all ones (gap / unknown: datatype.py:105-111); further
 *                rows are ambiguity masks (use_ambiguities).
 * weights        double [N] pattern multiplicities (site_pattern.py:89-95)
 * postorder      int32 [T-1][3] (node, left, right) triples in post-order;
 *                tips are 0..T-1, the last triple is the root
 *                (tree_model.py:37-53, :187-195)
 * These are host pointers; the data is copied to the device.
 */
int ttb2_create(const ttb2_config* config, const uint8_t* tip_codes,
                const double* code_partials, const double* weights,
                const int32_t* postorder, ttb2_engine** out);

/* New topology (same tips): rebuilds the level schedule. Host pointer. */
int ttb2_set_postorder(ttb2_engine* engine, const int32_t* postorder);

void ttb2_destroy(ttb2_engine* engine);

/* Use `cuda_stream` (a cudaStream_t) for all subsequent work; NULL = the
 * legacy default stream.  The new stream is ordered behind the work already queued on
 * the previous one (event wait, no host synchronisation); a no-op if unchanged. */
int ttb2_set_stream(ttb2_engine* engine, void* cuda_stream);
int ttb2_synchronize(ttb2_engine* engine);

/*
 * Log-likelihood with caller-supplied transition matrices.
 *   mats   [D][B][K][S][S]  rows = parent state (applied as P @ partial)
 *   freqs  [freq_draws][S]   props [prop_draws][K]
 *   lnl    [D] (output)
 */
int ttb2_loglik_mats(ttb2_engine* engine, int32_t draws, const double* mats,
                     const double* freqs, int32_t freq_draws,
                     const double* props, int32_t prop_draws, double* lnl,
                     int32_t where);

/*
 * Gradient of sum_d grad_lnl[d] * lnL[d] for the latest ttb2_loglik_mats call.
 *   grad_lnl [D] or NULL (= ones)
 *   d_mats   [D][B][K][S][S]; d_freqs [freq_draws][S]; d_props [prop_draws][K]
 *   any output pointer may be NULL (skipped).
 */
int ttb2_grad_mats(ttb2_engine* engine, const double* grad_lnl, double* d_mats,
                   double* d_freqs, double* d_props, int32_t where);

/*
 * Log-likelihood from an eigen-decomposed generator, P = V exp(L r t) V^-1
 * evaluated on the device for every branch x category x draw.
 *   branch_lengths [D][B]   (already multiplied by clock rates; unrooted trees
 *                            carry the zero pad for node 2T-3)
 *   site_rates [rate_draws][K], props [prop_draws][K]
 *   evec [eig_draws][S][S] = V, ivec [eig_draws][S][S] = V^-1,
 *   eval [eig_draws][S]; freqs [freq_draws][S]
 */
int ttb2_loglik_eigen(ttb2_engine* engine, int32_t draws,
                      const double* branch_lengths, const double* site_rates,
                      int32_t rate_draws, const double* props,
                      int32_t prop_draws, const double* evec,
                      const double* ivec, const double* eval,
                      int32_t eig_draws, const double* freqs,
                      int32_t freq_draws, double* lnl, int32_t where);

/*
 * The same evaluation from the normalised generator itself: the eigen-system of every draw is
 * computed on the device (one CTA per generator, parallel cyclic Jacobi on the sqrt(pi)
 * symmetrisation, csrc/eigen.cu) -- replaces the host torch.linalg.eigh round trip of
 * SymmetricSubstitutionModel.p_t, substitution_model/abstract.py:57-66.
 *   q_norm [q_draws][S][S]  reversible generator, already normalised (abstract.py:49-50);
 *                           only the lower triangle of its symmetrisation is read, as eigh does
 *   freqs  [freq_draws][S]  its stationary frequencies (also the root frequencies)
 * One eigen-system per max(q_draws, freq_draws).  S <= 64.  ttb2_grad_eigen follows it exactly
 * as it follows ttb2_loglik_eigen (d_q then has max(q_draws, freq_draws) leading entries).
 */
int ttb2_loglik_q(ttb2_engine* engine, int32_t draws, const double* branch_lengths,
                  const double* site_rates, int32_t rate_draws, const double* props,
                  int32_t prop_draws, const double* q_norm, int32_t q_draws,
                  const double* freqs, int32_t freq_draws, double* lnl, int32_t where);

/*
 * The same evaluation for a GENERAL generator (not necessarily reversible: complex spectrum, no
 * real eigen route): P = exp(Q r t) by scaling and squaring on the device, one CTA per branch x
 * category x draw (csrc/expm.cu) -- replaces NonSymmetricSubstitutionModel.p_t =
 * torch.matrix_exp(Q t), substitution_model/abstract.py:89-94, the route of
 * GeneralNonSymmetricSubstitutionModel and the discrete-trait likelihoods of
 * cli/evolution.py:540-611.
 *   q [q_draws][S][S]  generator as the model normalises it (abstract.py:90-91)
 * ttb2_grad_eigen / ttb2_grad_eigen_packed follow it exactly as they follow ttb2_loglik_q; d_q
 * ([q_draws][S][S], all S*S entries independent) comes from the exact adjoint of the matrix
 * exponential (Frechet derivative in dual arithmetic, no stored intermediates).
 */
int ttb2_loglik_expm(ttb2_engine* engine, int32_t draws, const double* branch_lengths,
                     const double* site_rates, int32_t rate_draws, const double* props,
                     int32_t prop_draws, const double* q, int32_t q_draws, const double* freqs,
                     int32_t freq_draws, double* lnl, int32_t where);

/* The eigen-system used by the latest eigen-mode call: evec / ivec [eig_draws][S][S],
 * eval [eig_draws][S] (ascending); any pointer may be NULL. */
int ttb2_get_eigen(ttb2_engine* engine, double* evec, double* ivec, double* eval,
                   int32_t where);

/*
 * Gradient for the latest ttb2_loglik_eigen / ttb2_loglik_q / ttb2_loglik_expm call.
 *   d_branch_lengths [D][B]; d_site_rates [rate_draws][K];
 *   d_props [prop_draws][K]; d_freqs [freq_draws][S] (root term only);
 *   d_q [eig_draws][S][S] = d lnL / d Q for the generator Q = V L V^-1 with all
 *   S*S entries independent (chain it through the model's Q builder).
 *   Finite at repeated eigenvalues (divided differences), unlike the
 *   reference's eigh backward.
 */
int ttb2_grad_eigen(ttb2_engine* engine, const double* grad_lnl,
                    double* d_branch_lengths, double* d_site_rates,
                    double* d_props, double* d_q, double* d_freqs,
                    int32_t where);

/*
 * The same gradient as ONE contiguous vector, for callers that reduce it across pattern shards
 * (one NCCL all-reduce over NVLink) or move it to the host in one copy:
 *   packed = [ lnL[D] | d_branch_lengths[D][B] | d_site_rates[rate_draws][K] |
 *              d_props[prop_draws][K] | d_q[eig_draws][S][S] | d_freqs[freq_draws][S] ]
 * with the draw counts of the latest ttb2_loglik_eigen / ttb2_loglik_q call;
 * ttb2_packed_count returns its length in doubles, `capacity` is the length of `packed`.
 * The output kernels write this layout directly (no gather step).
 */
int ttb2_grad_eigen_packed(ttb2_engine* engine, const double* grad_lnl, double* packed,
                           int64_t capacity, int32_t where);
int64_t ttb2_packed_count(const ttb2_engine* engine);

/* Per-pattern log-likelihoods of the latest loglik call: out [D][N]. */
int ttb2_site_loglik(ttb2_engine* engine, double* out, int32_t where);

/* Transition matrices used by the latest loglik call: out [D][B][K][S][S]. */
int ttb2_get_mats(ttb2_engine* engine, double* out, int32_t where);

/*
 * Optional phase timing with CUDA events on the engine's stream (for bench.py's
 * live roofline figure).  When enabled, every loglik/grad call records events
 * around its kernel groups; ttb2_phase_ms synchronises and returns the
 * durations of the latest call of each kind, in milliseconds:
 *   out[0] transition matrices   out[1] post-order level kernels
 *   out[2] root + lnL reduction  out[3] pre-order level kernels (incl. root)
 *   out[4] gradient contraction / output assembly
 * and the number of level-kernel launches in out[5] (post-order) and out[6]
 * (pre-order).
 */
#define TTB2_PHASE_COUNT 7
int ttb2_enable_timing(ttb2_engine* engine, int32_t on);
int ttb2_phase_ms(ttb2_engine* engine, double* out /*[TTB2_PHASE_COUNT]*/);

/*
 * Site-pattern compression (host code, no GPU needed): the unique columns of an
 * alignment in the reference's order with their multiplicities -- replaces
 * `compress`, torchtree/evolution/site_pattern.py:69-97.
 *   sequences [taxa][length] bytes (taxa in Taxa order); a site is `group`
 *   consecutive characters (3 for codons, site_pattern.py:79-82)
 *   patterns  [taxa][n][group] bytes, out (caller allocates taxa*length bytes)
 *   weights   [n] out (caller allocates length/group doubles)
 */
int ttb2_compress_patterns(const uint8_t* sequences, int32_t taxa, int64_t length,
                           int32_t group, uint8_t* patterns, double* weights,
                           int64_t* pattern_count);

/*
 * Node-height reparameterisation of time trees on the device -- replaces the Python
 * loop (and its autograd tape) of GeneralNodeHeightTransform._call,
 * torchtree/evolution/tree_height_transform.py:58-66:
 *     h[root] = x[root],  h[c] = b[c] + x[c] * (h[parent(c)] - b[c])
 * for the T-1 internal nodes (index = node - T); b = the transform's lower bounds
 * (`_bounds[taxa_count:]`, update_bounds :36-56).  One plan per topology.
 *   postorder [T-1][3] (node, left, right), root last;  bounds [T-1]
 *   x, heights, grad_heights, grad_x: [draws][T-1]; all host or all device (`where`)
 * The calls return after the result is complete (they synchronise their own stream).
 */
typedef struct ttb2_heights ttb2_heights;
int ttb2_heights_create(int32_t tip_count, const int32_t* postorder, const double* bounds,
                        int32_t device, ttb2_heights** out);
int ttb2_heights_destroy(ttb2_heights* plan);
int ttb2_heights_forward(ttb2_heights* plan, int32_t draws, const double* x, double* heights,
                         int32_t where);
/* grad_x = (d heights / d x)^T grad_heights: the backward of the call above. */
int ttb2_heights_backward(ttb2_heights* plan, int32_t draws, const double* x,
                          const double* heights, const double* grad_heights, double* grad_x,
                          int32_t where);

/*
 * Constant-population coalescent log-density of a batch of time trees on the device -- replaces
 * ConstantCoalescent.log_prob, torchtree/evolution/coalescent.py:112-134 (argsort of the 2T-1 node
 * heights, lineage counts by cumulative sum, sum of C(k,2) x interval / theta) and its autograd
 * backward (closed form once the order is known).  One CTA per draw; T >= 2 (up to
 * 4096 tips the sort runs in shared memory, beyond that on a global-memory scratch area).
 *   node_heights [draws][2T-1]  tips first (sampling times), then the T-1 internal nodes
 *   theta        [theta_draws]  population size, theta_draws = 1 (shared) or draws
 *   log_prob     [draws]
 *   d_heights    [draws][2T-1]  d log_prob[d] / d node_heights[d][.]   (may be NULL)
 *   d_theta      [draws]        d log_prob[d] / d theta (per draw; sum them for a shared theta; may be NULL)
 * Stateless; all pointers host or all device (`where`); returns after the result is complete.
 */
int ttb2_coalescent_constant(int32_t device, int32_t draws, int32_t tip_count,
                             const double* node_heights, const double* theta,
                             int32_t theta_draws, double* log_prob, double* d_heights,
                             double* d_theta, int32_t where);

/*
 * Piecewise-constant coalescents of a batch of time trees on the device: the skyride (one
 * population size per inter-coalescent interval; grid_count = 0, theta_count = T - 1) and the
 * skygrid (one per segment of a fixed time grid; theta_count = grid_count + 1) -- replace
 * PiecewiseConstantCoalescent.log_prob, torchtree/evolution/coalescent.py:311-396, and
 * PiecewiseConstantCoalescentGrid.log_prob, :459-549, with their autograd backward.  Same
 * structure as ttb2_coalescent_constant (sort of the events -- node heights and grid points --,
 * scans for the lineage count and the population-size index of every interval, closed-form
 * derivatives).
 *   node_heights [draws][2T-1]; theta [theta_draws][theta_count], theta_draws = 1 or draws;
 *   grid [grid_count] ascending (NULL for the skyride)
 *   log_prob [draws]; d_heights [draws][2T-1]; d_theta [draws][theta_count] (per draw; sum them
 *   for a shared theta); either gradient pointer may be NULL.
 */
int ttb2_coalescent_piecewise(int32_t device, int32_t draws, int32_t tip_count,
                              const double* node_heights, const double* theta,
                              int32_t theta_draws, int32_t theta_count, const double* grid,
                              int32_t grid_count, double* log_prob, double* d_heights,
                              double* d_theta, int32_t where);

/* Kernels launched by this engine since creation (bench `gpu_launches`). */
int64_t ttb2_launch_count(const ttb2_engine* engine);
/* Device bytes currently held by this engine. */
int64_t ttb2_device_bytes(const ttb2_engine* engine);
/* Number of ttb2_loglik_* calls this engine has served.  A gradient call always refers to
 * the latest one: a binding that defers the gradient (autograd) compares this serial with
 * the one it saw after its own forward and re-runs the forward if another one intervened
 * (SURVEY 8(b) autograd contract). */
int64_t ttb2_eval_serial(const ttb2_engine* engine);
/* The configuration the engine was created with (shapes for argument checks in bindings). */
int ttb2_get_config(const ttb2_engine* engine, ttb2_config* out);

const char* ttb2_last_error(void);
int ttb2_version(void);

#ifdef __cplusplus
}
#endif
#endif /* TTB200_H */
