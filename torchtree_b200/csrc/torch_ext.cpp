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
#include <torch/extension.h>

#include <algorithm>
#include <cstdint>
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

ttb2_engine* as_engine(int64_t handle) {
  TORCH_CHECK(handle != 0, "ttb200: null engine handle");
  return reinterpret_cast<ttb2_engine*>(static_cast<intptr_t>(handle));
}

ttb2_config config_of(int64_t handle) {
  ttb2_config cfg;
  check(ttb2_get_config(as_engine(handle), &cfg), "ttb2_get_config");
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

// TTB2_HOST / TTB2_DEVICE for one call; device tensors must live on the engine's device
// and are made visible to the engine's stream by synchronising the producer stream.
int where_of(std::initializer_list<Tensor> tensors, int device) {
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
    const c10::impl::VirtualGuardImpl impl(c10::DeviceType::CUDA);
    impl.getStream(c10::Device(c10::DeviceType::CUDA, (c10::DeviceIndex)device)).synchronize();
    return TTB2_DEVICE;
  }
  return TTB2_HOST;
}

double* dptr(const Tensor& t) { return t.defined() ? t.data_ptr<double>() : nullptr; }

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

  static void run_forward(int64_t handle, const ttb2_config& cfg, const Tensor& bls,
                          const Tensor& rates, const Tensor& props, const Tensor& q,
                          const Tensor& freqs, Tensor& lnl) {
    const int where = where_of({bls, rates, props, q, freqs}, cfg.device);
    if (device_eigh(cfg.state_count, std::max(q.size(0), freqs.size(0)))) {
      check(ttb2_loglik_q(as_engine(handle), (int32_t)bls.size(0), dptr(bls), dptr(rates),
                          (int32_t)rates.size(0), dptr(props), (int32_t)props.size(0), dptr(q),
                          (int32_t)q.size(0), dptr(freqs), (int32_t)freqs.size(0), dptr(lnl),
                          where),
            "ttb2_loglik_q");
      return;
    }
    // eigen-system through the sqrt(pi) symmetrisation (abstract.py:57-66), no graph
    Tensor root = freqs.sqrt();
    Tensor sym = root.unsqueeze(-1) * q / root.unsqueeze(-2);
    auto eig = at::linalg_eigh(sym, "L");
    Tensor evals = std::get<0>(eig).contiguous();
    Tensor u = std::get<1>(eig);
    Tensor evec = (u / root.unsqueeze(-1)).contiguous();
    Tensor ivec = (u.transpose(-1, -2) * root.unsqueeze(-2)).contiguous();
    check(ttb2_loglik_eigen(as_engine(handle), (int32_t)bls.size(0), dptr(bls), dptr(rates),
                            (int32_t)rates.size(0), dptr(props), (int32_t)props.size(0),
                            dptr(evec), dptr(ivec), dptr(evals), (int32_t)evec.size(0),
                            dptr(freqs), (int32_t)freqs.size(0), dptr(lnl), where),
          "ttb2_loglik_eigen");
  }

  static Tensor forward(AutogradContext* ctx, int64_t handle, const Tensor& branch_lengths,
                        const Tensor& site_rates, const Tensor& site_props,
                        const Tensor& q_norm, const Tensor& frequencies) {
    const ttb2_config cfg = config_of(handle);
    const int64_t S = cfg.state_count, K = cfg.category_count, B = 2 * (int64_t)cfg.tip_count - 2;
    Tensor bls = prep(branch_lengths, {B}, "branch_lengths");
    Tensor rates = prep(site_rates, {K}, "site_rates");
    Tensor props = prep(site_props, {K}, "site_props");
    Tensor q = prep(q_norm, {S, S}, "q_norm");
    Tensor freqs = prep(frequencies, {S}, "freqs");
    Tensor lnl = at::empty({bls.size(0)}, bls.options());
    run_forward(handle, cfg, bls, rates, props, q, freqs, lnl);
    ctx->saved_data["handle"] = handle;
    ctx->saved_data["serial"] = ttb2_eval_serial(as_engine(handle));
    ctx->save_for_backward({bls, rates, props, q, freqs});
    return lnl;
  }

  static variable_list backward(AutogradContext* ctx, variable_list grad_out) {
    const int64_t handle = ctx->saved_data["handle"].toInt();
    const ttb2_config cfg = config_of(handle);
    auto saved = ctx->get_saved_variables();
    const Tensor &bls = saved[0], &rates = saved[1], &props = saved[2], &q = saved[3],
                 &freqs = saved[4];
    if (ttb2_eval_serial(as_engine(handle)) != ctx->saved_data["serial"].toInt()) {
      // another forward ran on this engine since ours and overwrote its buffers:
      // recompute (SURVEY 8(b) autograd contract)
      Tensor lnl = at::empty({bls.size(0)}, bls.options());
      run_forward(handle, cfg, bls, rates, props, q, freqs, lnl);
      ctx->saved_data["serial"] = ttb2_eval_serial(as_engine(handle));
    }
    const int64_t S = cfg.state_count;
    Tensor g = grad_out[0].defined()
                   ? grad_out[0].detach().to(bls.device(), at::kDouble).reshape({-1}).contiguous()
                   : Tensor();
    TORCH_CHECK(!g.defined() || g.numel() == bls.size(0), "ttb200: grad_lnl must have one entry per draw");
    Tensor d_bl = at::empty_like(bls), d_rates = at::empty_like(rates),
           d_props = at::empty_like(props), d_freqs = at::empty_like(freqs);
    // one eigen-system per generator draw or frequency draw, whichever varies; d_q has that
    // leading extent and is summed back onto a shared generator below
    // (needs_input_grad indexes the tensor arguments only: bls 0, rates 1, props 2, q 3, freqs 4)
    const int64_t eig_draws = std::max(q.size(0), freqs.size(0));
    Tensor d_q = ctx->needs_input_grad(3) ? at::empty({eig_draws, S, S}, bls.options()) : Tensor();
    const int where = where_of({bls, g}, cfg.device);
    check(ttb2_grad_eigen(as_engine(handle), dptr(g), dptr(d_bl), dptr(d_rates), dptr(d_props),
                          dptr(d_q), dptr(d_freqs), where),
          "ttb2_grad_eigen");
    if (d_q.defined() && d_q.size(0) != q.size(0)) d_q = d_q.sum(0, /*keepdim=*/true);
    return {Tensor(), d_bl, d_rates, d_props, d_q, d_freqs};
  }
};

// ---------------------------------------------------------------------------------------
// caller-supplied transition matrices (ttb2_loglik_mats / ttb2_grad_mats): any
// SubstitutionModel.p_t, e.g. NonSymmetricSubstitutionModel (abstract.py:89-94)
struct MatsLikelihood : public torch::autograd::Function<MatsLikelihood> {
  static void run_forward(int64_t handle, const ttb2_config& cfg, const Tensor& mats,
                          const Tensor& freqs, const Tensor& props, Tensor& lnl) {
    const int where = where_of({mats, freqs, props}, cfg.device);
    check(ttb2_loglik_mats(as_engine(handle), (int32_t)mats.size(0), dptr(mats), dptr(freqs),
                           (int32_t)freqs.size(0), dptr(props), (int32_t)props.size(0), dptr(lnl),
                           where),
          "ttb2_loglik_mats");
  }

  static Tensor forward(AutogradContext* ctx, int64_t handle, const Tensor& matrices,
                        const Tensor& frequencies, const Tensor& site_props) {
    const ttb2_config cfg = config_of(handle);
    const int64_t S = cfg.state_count, K = cfg.category_count, B = 2 * (int64_t)cfg.tip_count - 2;
    Tensor mats = prep(matrices, {B, K, S, S}, "mats");
    Tensor freqs = prep(frequencies, {S}, "freqs");
    Tensor props = prep(site_props, {K}, "site_props");
    Tensor lnl = at::empty({mats.size(0)}, mats.options());
    run_forward(handle, cfg, mats, freqs, props, lnl);
    ctx->saved_data["handle"] = handle;
    ctx->saved_data["serial"] = ttb2_eval_serial(as_engine(handle));
    ctx->save_for_backward({mats, freqs, props});
    return lnl;
  }

  static variable_list backward(AutogradContext* ctx, variable_list grad_out) {
    const int64_t handle = ctx->saved_data["handle"].toInt();
    const ttb2_config cfg = config_of(handle);
    auto saved = ctx->get_saved_variables();
    const Tensor &mats = saved[0], &freqs = saved[1], &props = saved[2];
    if (ttb2_eval_serial(as_engine(handle)) != ctx->saved_data["serial"].toInt()) {
      Tensor lnl = at::empty({mats.size(0)}, mats.options());
      run_forward(handle, cfg, mats, freqs, props, lnl);
      ctx->saved_data["serial"] = ttb2_eval_serial(as_engine(handle));
    }
    Tensor g = grad_out[0].defined()
                   ? grad_out[0].detach().to(mats.device(), at::kDouble).reshape({-1}).contiguous()
                   : Tensor();
    TORCH_CHECK(!g.defined() || g.numel() == mats.size(0), "ttb200: grad_lnl must have one entry per draw");
    Tensor d_mats = ctx->needs_input_grad(0) ? at::empty_like(mats) : Tensor();
    Tensor d_freqs = at::empty_like(freqs), d_props = at::empty_like(props);
    const int where = where_of({mats, g}, cfg.device);
    check(ttb2_grad_mats(as_engine(handle), dptr(g), dptr(d_mats), dptr(d_freqs), dptr(d_props),
                         where),
          "ttb2_grad_mats");
    return {Tensor(), d_mats, d_freqs, d_props};
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

Tensor constant_coalescent(int64_t device, const Tensor& node_heights, const Tensor& theta) {
  return ConstantCoalescent::apply(device, node_heights, theta);
}

Tensor log_likelihood_eigen(int64_t handle, const Tensor& branch_lengths, const Tensor& site_rates,
                            const Tensor& site_props, const Tensor& q_norm, const Tensor& freqs) {
  return EigenLikelihood::apply(handle, branch_lengths, site_rates, site_props, q_norm, freqs);
}

Tensor log_likelihood_mats(int64_t handle, const Tensor& mats, const Tensor& freqs,
                           const Tensor& site_props) {
  return MatsLikelihood::apply(handle, mats, freqs, site_props);
}

Tensor node_heights(int64_t plan, int64_t device, const Tensor& x) {
  return NodeHeights::apply(plan, device, x);
}

}  // namespace

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
  m.doc() = "torchtree_b200: torch autograd Functions over the ttb200 C ABI (include/ttb200.h)";
  m.def("log_likelihood_eigen", &log_likelihood_eigen,
        "lnL [D] of a reversible model; backward = analytic pre-order gradient",
        py::arg("engine_handle"), py::arg("branch_lengths"), py::arg("site_rates"),
        py::arg("site_props"), py::arg("q_norm"), py::arg("freqs"));
  m.def("log_likelihood_mats", &log_likelihood_mats,
        "lnL [D] from caller-supplied transition matrices [D,B,K,S,S]", py::arg("engine_handle"),
        py::arg("mats"), py::arg("freqs"), py::arg("site_props"));
  m.def("node_heights", &node_heights, "ratios / root height -> internal node heights",
        py::arg("plan_handle"), py::arg("device"), py::arg("x"));
  m.def("constant_coalescent", &constant_coalescent,
        "log-density [D] of the constant-population coalescent for node heights [D, 2T-1]",
        py::arg("device"), py::arg("node_heights"), py::arg("theta"));
  m.def("abi_version", []() { return ttb2_version(); });
}
