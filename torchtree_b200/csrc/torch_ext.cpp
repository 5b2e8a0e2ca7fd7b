// torch C++ extension: custom autograd Functions over the ttb200 C ABI.
//
// This is the layer BASELINE.json's north_star asks for between the drop-in Python
// class (torchtree_b200/tree_likelihood.py) and the CUDA engine: the reference
// differentiates its peeling loop with the autograd tape
// (torchtree/evolution/tree_likelihood.py:40-278); here
// `forward` calls ttb2_loglik_* and `backward` calls the engine's analytic pre-order
// pass (ttb2_grad_*), so no tape over the tree exists.  Everything below only marshals
// tensors into the plain-pointer C ABI of include/ttb200.h -- no arithmetic on the path
// happens here (the eigen-decomposition of the generator runs on the device inside
// ttb2_loglik_q; only state spaces beyond 64 states use at::linalg_eigh on the host, as
// SymmetricSubstitutionModel.p_t does at substitution_model/abstract.py:57-66).
//
// Built by torchtree_b200/build.py into torchtree_b200/_ttb200_torch.so, linked against
// lib/libttb200.so (rpath $ORIGIN/lib).  There is no fallback: without a CUDA device the
// C ABI calls fail and the error is raised.
#include <c10/cuda/CUDAStream.h>
#include <torch/extension.h>

#include <algorithm>
#include <cstdint>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "ttb200.h"

namespace {

using torch::Tensor;
using torch::autograd::AutogradContext;
using torch::autograd::variable_list;

void check(int status, const char* what) {
  if (status != 0) {
    const char* msg = ttb2_last_error();
    // a plain std::runtime_error (pybind11 raises it as RuntimeError): building a c10::Error here,
    // right after a failed CUDA runtime call inside libttb200's static runtime, crashed while it
    // captured its backtrace
    throw std::runtime_error(std::string(what) + " failed (status " + std::to_string(status) +
                             "): " + (msg ? msg : "?"));
  }
}

// Owner of one ttb2_engine.  The Python `Engine` holds one reference; every autograd node
// made from a forward on that engine holds another (through a keep-alive tensor in
// ctx->saved_data), so a `backward` that runs after the Python object went away -- the model was
// garbage-collected, or TreeLikelihoodModel replaced its engine by a larger one for more draws
// -- still finds the buffers of ITS forward.  An explicit close() destroys the engine at once;
// a later backward then raises instead of touching freed memory.
struct EngineRef {
  ttb2_engine* h = nullptr;
  // host eigen-system of the latest generator (large single generators, see device_eigh)
  Tensor cacheQ, cacheF, evec, ivec, evals;
  explicit EngineRef(int64_t handle)
      : h(reinterpret_cast<ttb2_engine*>(static_cast<intptr_t>(handle))) {}
  EngineRef(const EngineRef&) = delete;
  EngineRef& operator=(const EngineRef&) = delete;
  ~EngineRef() { close(); }
  void close() {
    if (h) ttb2_destroy(h);
    h = nullptr;
  }
  ttb2_engine* get() const {
    if (!h)
      throw std::runtime_error(
          "ttb200: the engine of this evaluation has been closed (Engine.close()) before its "
          "backward pass ran");
    return h;
  }
};
using EnginePtr = std::shared_ptr<EngineRef>;

// a 1-element tensor whose storage deleter owns a reference to the engine
Tensor keepalive(const EnginePtr& p) {
  auto* box = new EnginePtr(p);
  return at::from_blob(
      box, {1}, [](void* b) { delete static_cast<EnginePtr*>(b); },
      at::TensorOptions().dtype(at::kByte));
}

const EnginePtr& engine_of(const Tensor& keep) { return *static_cast<const EnginePtr*>(keep.data_ptr()); }

ttb2_config config_of(ttb2_engine* eng) {
  ttb2_config cfg;
  check(ttb2_get_config(eng, &cfg), "ttb2_get_config");
  return cfg;
}

// contiguous fp64 [d, tail...]
Tensor prep(const Tensor& x, at::IntArrayRef tail, const char* name) {
  Tensor t = x.detach();
  if (t.scalar_type() != at::kDouble) t = t.to(at::kDouble);
  TORCH_CHECK(t.dim() == (int64_t)tail.size() + 1, "ttb200: ", name, " must have ",
              tail.size() + 1, " dimensions (draws first), got ", t.dim());
  for (size_t i = 0; i < tail.size(); ++i)
    TORCH_CHECK(t.size(i + 1) == tail[i], "ttb200: ", name, " has extent ", t.size(i + 1),
                " in dimension ", i + 1, ", expected ", tail[i]);
  return t.contiguous();
}

// TTB2_HOST / TTB2_DEVICE for one call; device tensors must live on the engine's device.
// With device tensors the engine is bound to torch's CURRENT stream of that device
// (ttb2_set_stream: event-ordered, no host synchronisation), so its reads follow the producers
// of the inputs and its outputs are ordered for whoever consumes them on that stream.
int where_of(std::initializer_list<Tensor> tensors, int device, ttb2_engine* eng = nullptr) {
  int cuda = 0, total = 0;
  for (const Tensor& t : tensors) {
    if (!t.defined()) continue;
    ++total;
    if (t.is_cuda()) {
      ++cuda;
      TORCH_CHECK(t.get_device() == device, "ttb200: input tensor is on CUDA device ",
                  t.get_device(), " but the engine lives on device ", device);
    }
  }
  TORCH_CHECK(cuda == 0 || cuda == total, "ttb200: inputs must be all host or all device tensors");
  if (cuda) {
    auto stream = c10::cuda::getCurrentCUDAStream((c10::DeviceIndex)device);
    if (eng)
      check(ttb2_set_stream(eng, stream.stream()), "ttb2_set_stream");
    else
      stream.synchronize();   // stateless entry points run on their own stream
    return TTB2_DEVICE;
  }
  return TTB2_HOST;
}

double* dptr(const Tensor& t) { return t.defined() ? t.data_ptr<double>() : nullptr; }

// dtype policy (SURVEY 8(b) variant 5, `--dtype float32` runs): the engine computes in fp64
// whatever the inputs are; inputs of another floating type are converted on the way in, and
// lnL and every gradient are handed back in the dtype of the tensor they belong to.
Tensor like_input(const Tensor& value, at::ScalarType dtype) {
  return value.defined() && value.scalar_type() != dtype ? value.to(dtype) : value;
}

// ---------------------------------------------------------------------------------------
// reversible models: P = V exp(L r t) V^-1 on the device (ttb2_loglik_eigen / ttb2_grad_eigen)
struct EigenLikelihood : public torch::autograd::Function<EigenLikelihood> {
  // Where the S x S eigen-system is computed.  The device Jacobi kernel (csrc/eigen.cu, one
  // CTA per generator) takes a flat ~0.02 / 0.13 / 0.85 ms for S = 4 / 20 / 61 however many
  // draws there are (up to one per SM); LAPACK on the host takes 0.02 / 0.03 / 0.18 ms PER
  // generator (profiles/r01_eigen_device.jsonl).  So: small alphabets and batches of draws
  // decompose on the device (ttb2_loglik_q), a single large generator on the host
  // (at::linalg_eigh + ttb2_loglik_eigen), as does anything beyond the kernel's 64 states.
  static bool device_eigh(int64_t S, int64_t eig_draws) {
    return S <= 64 && (S <= 8 || eig_draws >= 6);
  }

  // `general`: the generator is not reversible -- P = exp(Q r t) by scaling and squaring on
  // the device (ttb2_loglik_expm, csrc/expm.cu) instead of the eigen route
  static void run_forward(EngineRef& ref, const ttb2_config& cfg, const Tensor& bls,
                          const Tensor& rates, const Tensor& props, const Tensor& q,
                          const Tensor& freqs, Tensor& lnl, bool general) {
    ttb2_engine* eng = ref.get();
    const int where = where_of({bls, rates, props, q, freqs}, cfg.device, eng);
    if (general) {
      check(ttb2_loglik_expm(eng, (int32_t)bls.size(0), dptr(bls), dptr(rates),
                             (int32_t)rates.size(0), dptr(props), (int32_t)props.size(0), dptr(q),
                             (int32_t)q.size(0), dptr(freqs), (int32_t)freqs.size(0), dptr(lnl),
                             where),
            "ttb2_loglik_expm");
      return;
    }
    if (device_eigh(cfg.state_count, std::max(q.size(0), freqs.size(0)))) {
      check(ttb2_loglik_q(eng, (int32_t)bls.size(0), dptr(bls), dptr(rates),
                          (int32_t)rates.size(0), dptr(props), (int32_t)props.size(0), dptr(q),
                          (int32_t)q.size(0), dptr(freqs), (int32_t)freqs.size(0), dptr(lnl),
                          where),
            "ttb2_loglik_q");
      return;
    }
    // eigen-system through the sqrt(pi) symmetrisation (abstract.py:57-66), no graph.  It is
    // kept with the engine and reused while the generator does not change -- the reference
    // decomposes empirical models (LG, WAG) once, general.py:300-306.
    if (!(ref.cacheQ.defined() && ref.cacheQ.sizes() == q.sizes() &&
          ref.cacheF.sizes() == freqs.sizes() && ref.cacheQ.device() == q.device() &&
          at::equal(ref.cacheQ, q) && at::equal(ref.cacheF, freqs))) {
      Tensor root = freqs.sqrt();
      Tensor sym = root.unsqueeze(-1) * q / root.unsqueeze(-2);
      auto eig = at::linalg_eigh(sym, "L");
      Tensor u = std::get<1>(eig);
      ref.evals = std::get<0>(eig).contiguous();
      ref.evec = (u / root.unsqueeze(-1)).contiguous();
      ref.ivec = (u.transpose(-1, -2) * root.unsqueeze(-2)).contiguous();
      ref.cacheQ = q.clone();
      ref.cacheF = freqs.clone();
    }
    check(ttb2_loglik_eigen(eng, (int32_t)bls.size(0), dptr(bls), dptr(rates),
                            (int32_t)rates.size(0), dptr(props), (int32_t)props.size(0),
                            dptr(ref.evec), dptr(ref.ivec), dptr(ref.evals),
                            (int32_t)ref.evec.size(0), dptr(freqs), (int32_t)freqs.size(0),
                            dptr(lnl), where),
          "ttb2_loglik_eigen");
  }

  static Tensor forward(AutogradContext* ctx, const Tensor& engine_keepalive,
                        const Tensor& branch_lengths, const Tensor& site_rates,
                        const Tensor& site_props, const Tensor& q_norm,
                        const Tensor& frequencies, bool general) {
    const EnginePtr ref = engine_of(engine_keepalive);
    const ttb2_config cfg = config_of(ref->get());
    const int64_t S = cfg.state_count, K = cfg.category_count, B = 2 * (int64_t)cfg.tip_count - 2;
    Tensor bls = prep(branch_lengths, {B}, "branch_lengths");
    Tensor rates = prep(site_rates, {K}, "site_rates");
    Tensor props = prep(site_props, {K}, "site_props");
    Tensor q = prep(q_norm, {S, S}, "q_norm");
    Tensor freqs = prep(frequencies, {S}, "freqs");
    Tensor lnl = at::empty({bls.size(0)}, bls.options());
    {
      pybind11::gil_scoped_release nogil;
      run_forward(*ref, cfg, bls, rates, props, q, freqs, lnl, general);
    }
    ctx->saved_data["engine"] = engine_keepalive;
    ctx->saved_data["general"] = general;
    ctx->saved_data["serial"] = ttb2_eval_serial(ref->get());
    ctx->saved_data["dtypes"] = std::vector<int64_t>{
        (int64_t)branch_lengths.scalar_type(), (int64_t)site_rates.scalar_type(),
        (int64_t)site_props.scalar_type(), (int64_t)q_norm.scalar_type(),
        (int64_t)frequencies.scalar_type()};
    ctx->save_for_backward({bls, rates, props, q, freqs});
    return like_input(lnl, branch_lengths.scalar_type());
  }

  static variable_list backward(AutogradContext* ctx, variable_list grad_out) {
    const EnginePtr ref = engine_of(ctx->saved_data["engine"].toTensor());
    ttb2_engine* eng = ref->get();
    const ttb2_config cfg = config_of(eng);
    auto saved = ctx->get_saved_variables();
    const Tensor &bls = saved[0], &rates = saved[1], &props = saved[2], &q = saved[3],
                 &freqs = saved[4];
    const bool general = ctx->saved_data["general"].toBool();
    if (ttb2_eval_serial(eng) != ctx->saved_data["serial"].toInt()) {
      // another forward ran on this engine since ours and overwrote its buffers:
      // recompute (SURVEY 8(b) autograd contract)
      Tensor lnl = at::empty({bls.size(0)}, bls.options());
      run_forward(*ref, cfg, bls, rates, props, q, freqs, lnl, general);
      ctx->saved_data["serial"] = ttb2_eval_serial(eng);
    }
    const int64_t S = cfg.state_count, K = cfg.category_count, D = bls.size(0), B = bls.size(1);
    Tensor g = grad_out[0].defined()
                   ? grad_out[0].detach().to(bls.device(), at::kDouble).reshape({-1}).contiguous()
                   : Tensor();
    TORCH_CHECK(!g.defined() || g.numel() == D, "ttb200: grad_lnl must have one entry per draw");
    // one eigen-system per generator draw or frequency draw, whichever varies; d_q has that
    // leading extent and is summed back onto a shared generator below
    const int64_t eig_draws = general ? q.size(0) : std::max(q.size(0), freqs.size(0));
    // the engine writes all small outputs as one vector
    // [lnL | d_bl | d_rates | d_props | d_q | d_freqs]: one copy, then views
    const int64_t count = ttb2_packed_count(eng);
    TORCH_CHECK(count == D + D * B + (rates.size(0) + props.size(0)) * K + eig_draws * S * S +
                             freqs.size(0) * S,
                "ttb200: packed gradient layout does not match the saved inputs");
    Tensor packed = at::empty({count}, bls.options());
    const int where = where_of({bls, g}, cfg.device, eng);
    check(ttb2_grad_eigen_packed(eng, dptr(g), dptr(packed), count, where),
          "ttb2_grad_eigen_packed");
    int64_t off = D;
    auto take = [&](int64_t n, at::IntArrayRef shape) {
      Tensor v = packed.narrow(0, off, n).view(shape);
      off += n;
      return v;
    };
    Tensor d_bl = take(D * B, {D, B});
    Tensor d_rates = take(rates.size(0) * K, {rates.size(0), K});
    Tensor d_props = take(props.size(0) * K, {props.size(0), K});
    Tensor d_q = take(eig_draws * S * S, {eig_draws, S, S});
    Tensor d_freqs = take(freqs.size(0) * S, {freqs.size(0), S});
    if (d_q.size(0) != q.size(0)) d_q = d_q.sum(0, /*keepdim=*/true);
    const auto dt = ctx->saved_data["dtypes"].toIntVector();
    return {Tensor(),
            like_input(d_bl, (at::ScalarType)dt[0]),
            like_input(d_rates, (at::ScalarType)dt[1]),
            like_input(d_props, (at::ScalarType)dt[2]),
            ctx->needs_input_grad(4) ? like_input(d_q, (at::ScalarType)dt[3]) : Tensor(),
            like_input(d_freqs, (at::ScalarType)dt[4]), Tensor()};
  }
};

// ---------------------------------------------------------------------------------------
// caller-supplied transition matrices (ttb2_loglik_mats / ttb2_grad_mats): any
// SubstitutionModel.p_t, e.g. NonSymmetricSubstitutionModel (abstract.py:89-94)
struct MatsLikelihood : public torch::autograd::Function<MatsLikelihood> {
  static void run_forward(EngineRef& ref, const ttb2_config& cfg, const Tensor& mats,
                          const Tensor& freqs, const Tensor& props, Tensor& lnl) {
    ttb2_engine* eng = ref.get();
    const int where = where_of({mats, freqs, props}, cfg.device, eng);
    check(ttb2_loglik_mats(eng, (int32_t)mats.size(0), dptr(mats), dptr(freqs),
                           (int32_t)freqs.size(0), dptr(props), (int32_t)props.size(0), dptr(lnl),
                           where),
          "ttb2_loglik_mats");
  }

  static Tensor forward(AutogradContext* ctx, const Tensor& engine_keepalive,
                        const Tensor& matrices, const Tensor& frequencies,
                        const Tensor& site_props) {
    const EnginePtr ref = engine_of(engine_keepalive);
    const ttb2_config cfg = config_of(ref->get());
    const int64_t S = cfg.state_count, K = cfg.category_count, B = 2 * (int64_t)cfg.tip_count - 2;
    Tensor mats = prep(matrices, {B, K, S, S}, "mats");
    Tensor freqs = prep(frequencies, {S}, "freqs");
    Tensor props = prep(site_props, {K}, "site_props");
    Tensor lnl = at::empty({mats.size(0)}, mats.options());
    {
      pybind11::gil_scoped_release nogil;
      run_forward(*ref, cfg, mats, freqs, props, lnl);
    }
    ctx->saved_data["engine"] = engine_keepalive;
    ctx->saved_data["serial"] = ttb2_eval_serial(ref->get());
    ctx->saved_data["dtypes"] = std::vector<int64_t>{(int64_t)matrices.scalar_type(),
                                                      (int64_t)frequencies.scalar_type(),
                                                      (int64_t)site_props.scalar_type()};
    ctx->save_for_backward({mats, freqs, props});
    return like_input(lnl, matrices.scalar_type());
  }

  static variable_list backward(AutogradContext* ctx, variable_list grad_out) {
    const EnginePtr ref = engine_of(ctx->saved_data["engine"].toTensor());
    ttb2_engine* eng = ref->get();
    const ttb2_config cfg = config_of(eng);
    auto saved = ctx->get_saved_variables();
    const Tensor &mats = saved[0], &freqs = saved[1], &props = saved[2];
    if (ttb2_eval_serial(eng) != ctx->saved_data["serial"].toInt()) {
      Tensor lnl = at::empty({mats.size(0)}, mats.options());
      run_forward(*ref, cfg, mats, freqs, props, lnl);
      ctx->saved_data["serial"] = ttb2_eval_serial(eng);
    }
    Tensor g = grad_out[0].defined()
                   ? grad_out[0].detach().to(mats.device(), at::kDouble).reshape({-1}).contiguous()
                   : Tensor();
    TORCH_CHECK(!g.defined() || g.numel() == mats.size(0), "ttb200: grad_lnl must have one entry per draw");
    Tensor d_mats = ctx->needs_input_grad(1) ? at::empty_like(mats) : Tensor();
    Tensor d_freqs = at::empty_like(freqs), d_props = at::empty_like(props);
    const int where = where_of({mats, g}, cfg.device, eng);
    check(ttb2_grad_mats(eng, dptr(g), dptr(d_mats), dptr(d_freqs), dptr(d_props), where),
          "ttb2_grad_mats");
    const auto dt = ctx->saved_data["dtypes"].toIntVector();
    return {Tensor(), like_input(d_mats, (at::ScalarType)dt[0]),
            like_input(d_freqs, (at::ScalarType)dt[1]), like_input(d_props, (at::ScalarType)dt[2])};
  }
};

// ---------------------------------------------------------------------------------------
// ratios + root height -> internal node heights (ttb2_heights_*), replacing the taped
// Python loop of GeneralNodeHeightTransform._call (tree_height_transform.py:58-66)
struct NodeHeights : public torch::autograd::Function<NodeHeights> {
  static Tensor forward(AutogradContext* ctx, int64_t plan, int64_t device, const Tensor& x) {
    TORCH_CHECK(plan != 0, "ttb200: null node-height plan");
    TORCH_CHECK(x.dim() >= 1, "ttb200: node_heights needs at least one dimension");
    const int64_t I = x.size(-1);
    Tensor xf = x.detach();
    if (xf.scalar_type() != at::kDouble) xf = xf.to(at::kDouble);
    xf = xf.reshape({-1, I}).contiguous();
    Tensor out = at::empty_like(xf);
    const int where = where_of({xf}, (int)device);
    check(ttb2_heights_forward(reinterpret_cast<ttb2_heights*>(static_cast<intptr_t>(plan)),
                               (int32_t)xf.size(0), dptr(xf), dptr(out), where),
          "ttb2_heights_forward");
    ctx->saved_data["plan"] = plan;
    ctx->saved_data["device"] = device;
    ctx->save_for_backward({xf, out});
    return out.reshape(x.sizes()).to(x.scalar_type());
  }

  static variable_list backward(AutogradContext* ctx, variable_list grad_out) {
    auto saved = ctx->get_saved_variables();
    const Tensor &xf = saved[0], &heights = saved[1];
    const int64_t I = xf.size(-1);
    Tensor grad = grad_out[0];
    Tensor gf = grad.detach().to(xf.device(), at::kDouble).reshape({-1, I}).contiguous();
    Tensor out = at::empty_like(xf);
    const int where = where_of({xf, gf}, (int)ctx->saved_data["device"].toInt());
    check(ttb2_heights_backward(
              reinterpret_cast<ttb2_heights*>(static_cast<intptr_t>(ctx->saved_data["plan"].toInt())),
              (int32_t)xf.size(0), dptr(xf), dptr(heights), dptr(gf), dptr(out), where),
          "ttb2_heights_backward");
    return {Tensor(), Tensor(), out.reshape(grad.sizes()).to(grad.scalar_type())};
  }
};

// ---------------------------------------------------------------------------------------
// constant-population coalescent (ttb2_coalescent_constant), replacing the argsort / gather /
// cumsum graph of ConstantCoalescent.log_prob (coalescent.py:112-134).  The kernel returns the
// partial derivatives with the value; backward only scales them.
struct ConstantCoalescent : public torch::autograd::Function<ConstantCoalescent> {
  static Tensor forward(AutogradContext* ctx, int64_t device, const Tensor& node_heights,
                        const Tensor& theta) {
    TORCH_CHECK(node_heights.dim() == 2, "ttb200: node_heights must be [draws, 2T-1]");
    TORCH_CHECK(theta.dim() == 1, "ttb200: theta must be [1] or [draws]");
    Tensor h = node_heights.detach();
    if (h.scalar_type() != at::kDouble) h = h.to(at::kDouble);
    h = h.contiguous();
    Tensor th = theta.detach();
    if (th.scalar_type() != at::kDouble) th = th.to(at::kDouble);
    th = th.contiguous();
    const int64_t D = h.size(0), n = h.size(1);
    TORCH_CHECK(n % 2 == 1 && n >= 3, "ttb200: node_heights needs 2T-1 columns");
    TORCH_CHECK(th.size(0) == 1 || th.size(0) == D, "ttb200: theta must have 1 or `draws` entries");
    Tensor lp = at::empty({D}, h.options()), dh = at::empty_like(h), dth = at::empty({D}, h.options());
    const int where = where_of({h, th}, (int)device);
    check(ttb2_coalescent_constant((int32_t)device, (int32_t)D, (int32_t)((n + 1) / 2), dptr(h),
                                   dptr(th), (int32_t)th.size(0), dptr(lp), dptr(dh), dptr(dth),
                                   where),
          "ttb2_coalescent_constant");
    ctx->saved_data["theta_draws"] = th.size(0);
    ctx->save_for_backward({dh, dth});
    return lp;
  }

  static variable_list backward(AutogradContext* ctx, variable_list grad_out) {
    auto saved = ctx->get_saved_variables();
    const Tensor &dh = saved[0], &dth = saved[1];
    Tensor g = grad_out[0].to(dh.device(), at::kDouble).reshape({-1});
    Tensor gh = dh * g.unsqueeze(1);
    Tensor gt = dth * g;
    if (ctx->saved_data["theta_draws"].toInt() == 1) gt = gt.sum(0, /*keepdim=*/true);
    return {Tensor(), gh, gt};
  }
};

// piecewise-constant coalescents: skyride (grid empty) and skygrid (ttb2_coalescent_piecewise),
// replacing PiecewiseConstantCoalescent(.Grid).log_prob (coalescent.py:311-396, :459-549)
struct PiecewiseCoalescent : public torch::autograd::Function<PiecewiseCoalescent> {
  static Tensor forward(AutogradContext* ctx, int64_t device, const Tensor& node_heights,
                        const Tensor& theta, const Tensor& grid) {
    TORCH_CHECK(node_heights.dim() == 2, "ttb200: node_heights must be [draws, 2T-1]");
    TORCH_CHECK(theta.dim() == 2, "ttb200: theta must be [1 or draws, M]");
    Tensor h = node_heights.detach().to(at::kDouble).contiguous();
    Tensor th = theta.detach().to(at::kDouble).contiguous();
    Tensor gr = grid.detach().to(h.device(), at::kDouble).contiguous();
    const int64_t D = h.size(0), n = h.size(1), M = th.size(1), G = gr.numel();
    TORCH_CHECK(n % 2 == 1 && n >= 3, "ttb200: node_heights needs 2T-1 columns");
    TORCH_CHECK(th.size(0) == 1 || th.size(0) == D, "ttb200: theta must have 1 or `draws` rows");
    Tensor lp = at::empty({D}, h.options()), dh = at::empty_like(h),
           dth = at::empty({D, M}, h.options());
    const int where = where_of({h, th, gr}, (int)device);
    check(ttb2_coalescent_piecewise((int32_t)device, (int32_t)D, (int32_t)((n + 1) / 2), dptr(h),
                                    dptr(th), (int32_t)th.size(0), (int32_t)M,
                                    G ? dptr(gr) : nullptr, (int32_t)G, dptr(lp), dptr(dh),
                                    dptr(dth), where),
          "ttb2_coalescent_piecewise");
    ctx->saved_data["theta_draws"] = th.size(0);
    ctx->save_for_backward({dh, dth});
    return lp.to(node_heights.scalar_type());
  }

  static variable_list backward(AutogradContext* ctx, variable_list grad_out) {
    auto saved = ctx->get_saved_variables();
    const Tensor &dh = saved[0], &dth = saved[1];
    Tensor g = grad_out[0].to(dh.device(), at::kDouble).reshape({-1, 1});
    Tensor gh = dh * g;
    Tensor gt = dth * g;
    if (ctx->saved_data["theta_draws"].toInt() == 1) gt = gt.sum(0, /*keepdim=*/true);
    return {Tensor(), gh, gt, Tensor()};
  }
};

Tensor piecewise_coalescent(int64_t device, const Tensor& node_heights, const Tensor& theta,
                            const Tensor& grid) {
  return PiecewiseCoalescent::apply(device, node_heights, theta, grid);
}

Tensor constant_coalescent(int64_t device, const Tensor& node_heights, const Tensor& theta) {
  return ConstantCoalescent::apply(device, node_heights, theta);
}

Tensor log_likelihood_eigen(const EnginePtr& engine, const Tensor& branch_lengths,
                            const Tensor& site_rates, const Tensor& site_props,
                            const Tensor& q_norm, const Tensor& freqs) {
  TORCH_CHECK(engine, "ttb200: null engine");
  return EigenLikelihood::apply(keepalive(engine), branch_lengths, site_rates, site_props, q_norm,
                                freqs, false);
}

Tensor log_likelihood_expm(const EnginePtr& engine, const Tensor& branch_lengths,
                           const Tensor& site_rates, const Tensor& site_props, const Tensor& q,
                           const Tensor& freqs) {
  TORCH_CHECK(engine, "ttb200: null engine");
  return EigenLikelihood::apply(keepalive(engine), branch_lengths, site_rates, site_props, q,
                                freqs, true);
}

Tensor log_likelihood_mats(const EnginePtr& engine, const Tensor& mats, const Tensor& freqs,
                           const Tensor& site_props) {
  TORCH_CHECK(engine, "ttb200: null engine");
  return MatsLikelihood::apply(keepalive(engine), mats, freqs, site_props);
}

Tensor node_heights(int64_t plan, int64_t device, const Tensor& x) {
  return NodeHeights::apply(plan, device, x);
}

}  // namespace

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
  m.doc() = "torchtree_b200: torch autograd Functions over the ttb200 C ABI (include/ttb200.h)";
  py::class_<EngineRef, EnginePtr>(m, "EngineRef",
                                   "Owner of one ttb2_engine handle (shared with the autograd "
                                   "nodes of its evaluations)")
      .def(py::init<int64_t>(), py::arg("handle"))
      .def("close", &EngineRef::close, "destroy the engine now")
      .def_property_readonly("closed", [](const EngineRef& r) { return r.h == nullptr; })
      .def_property_readonly("handle", [](const EngineRef& r) {
        return (int64_t) reinterpret_cast<intptr_t>(r.h);
      });
  m.def("log_likelihood_eigen", &log_likelihood_eigen,
        "lnL [D] of a reversible model; backward = analytic pre-order gradient",
        py::arg("engine"), py::arg("branch_lengths"), py::arg("site_rates"),
        py::arg("site_props"), py::arg("q_norm"), py::arg("freqs"));
  m.def("log_likelihood_expm", &log_likelihood_expm,
        "lnL [D] of a general (non-reversible) generator: P = exp(Q r t) on the device; backward = "
        "pre-order gradient + Frechet adjoint of the matrix exponential",
        py::arg("engine"), py::arg("branch_lengths"), py::arg("site_rates"), py::arg("site_props"),
        py::arg("q"), py::arg("freqs"));
  m.def("log_likelihood_mats", &log_likelihood_mats,
        "lnL [D] from caller-supplied transition matrices [D,B,K,S,S]", py::arg("engine"),
        py::arg("mats"), py::arg("freqs"), py::arg("site_props"));
  m.def("node_heights", &node_heights, "ratios / root height -> internal node heights",
        py::arg("plan_handle"), py::arg("device"), py::arg("x"));
  m.def("constant_coalescent", &constant_coalescent,
        "log-density [D] of the constant-population coalescent for node heights [D, 2T-1]",
        py::arg("device"), py::arg("node_heights"), py::arg("theta"));
  m.def("piecewise_coalescent", &piecewise_coalescent,
        "log-density [D] of the skyride (empty grid) / skygrid coalescent for node heights [D, 2T-1]",
        py::arg("device"), py::arg("node_heights"), py::arg("theta"), py::arg("grid"));
  m.def("abi_version", []() { return ttb2_version(); });
}
