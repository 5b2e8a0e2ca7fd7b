"""Python handle on one ttb200 engine (one tree-likelihood instance on one GPU).

Takes the place of the data the reference keeps on `TreeLikelihoodModel`
(tip partials / states, weights: torchtree/evolution/tree_likelihood.py:305-311)
and of the peeling functions it calls (:40-278).  All arithmetic happens in
libttb200.so; this file only marshals pointers.
"""
from __future__ import annotations

import ctypes
from typing import Optional, Sequence

import numpy as np
import torch

from . import _lib
from ._lib import TTB2_DEVICE, TTB2_HOST, EngineError, Ttb2Config


def default_code_partials(state_count: int) -> np.ndarray:
    """Unit vectors for the S states plus the all-ones row for gaps/unknowns
    (torchtree/evolution/datatype.py:105-111)."""
    return np.concatenate([np.eye(state_count), np.ones((1, state_count))], 0)


def codes_from_tip_partials(partials: Sequence, state_count: int):
    """Reference-style tip partials (T tensors [S,N], as produced by
    site_pattern.compress_alignment, site_pattern.py:100-124) -> uint8 codes
    [T,N] and the code table [C,S].  Distinct non-standard columns (ambiguity
    masks under use_ambiguities) get codes S+1, S+2, ..."""
    S = state_count
    table = [tuple(row) for row in default_code_partials(S)]
    index = {v: i for i, v in enumerate(table)}
    T = len(partials)
    N = partials[0].shape[-1]
    codes = np.empty((T, N), dtype=np.uint8)
    for t, p in enumerate(partials):
        cols = np.asarray(p, dtype=np.float64).T  # [N,S]
        uniq, inverse = np.unique(cols, axis=0, return_inverse=True)
        lut = np.empty(len(uniq), dtype=np.int64)
        for u, row in enumerate(uniq):
            key = tuple(row)
            if key not in index:
                if len(table) >= 255:
                    raise EngineError("more than 255 distinct tip partial vectors")
                index[key] = len(table)
                table.append(key)
            lut[u] = index[key]
        codes[t] = lut[inverse.reshape(-1)]
    return codes, np.array(table, dtype=np.float64)


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


class Engine:
    """One engine = one (tree, alignment shard) on one CUDA device."""

    def __init__(
        self,
        tip_codes,
        weights,
        postorder,
        state_count: int,
        category_count: int,
        code_partials=None,
        max_draws: int = 1,
        device: int = 0,
        flags: int = 0,
    ):
        self._h = None
        self._lib = _lib.load()
        tip_codes = np.ascontiguousarray(tip_codes, dtype=np.uint8)
        weights = np.ascontiguousarray(weights, dtype=np.float64)
        postorder = np.ascontiguousarray(postorder, dtype=np.int32)
        if code_partials is None:
            code_partials = default_code_partials(state_count)
            tip_codes = np.minimum(tip_codes, state_count).astype(np.uint8)
        code_partials = np.ascontiguousarray(code_partials, dtype=np.float64)
        T, N = tip_codes.shape
        if postorder.shape != (T - 1, 3):
            raise EngineError("postorder must have shape [T-1,3]")
        if weights.shape != (N,):
            raise EngineError("weights must have shape [N]")
        self.T, self.N, self.S, self.K = T, N, int(state_count), int(category_count)
        self.B = 2 * T - 2
        self.max_draws = int(max_draws)
        self.device = int(device)
        cfg = Ttb2Config(T, N, self.S, self.K, self.max_draws, code_partials.shape[0],
                         self.device, int(flags))
        handle = ctypes.c_void_p()
        _lib.check(
            self._lib.ttb2_create(
                ctypes.byref(cfg),
                tip_codes.ctypes.data_as(ctypes.c_void_p),
                code_partials.ctypes.data_as(ctypes.c_void_p),
                weights.ctypes.data_as(ctypes.c_void_p),
                postorder.ctypes.data_as(ctypes.c_void_p),
                ctypes.byref(handle),
            ),
            "ttb2_create",
        )
        self._h = handle
        self._ref = None
        self._draws = 0
        self._shapes = None

    # -- life-cycle ---------------------------------------------------------
    def share(self, ext):
        """Hand the ownership of the native engine to the torch extension's `EngineRef`
        (shared with the autograd nodes of its evaluations, csrc/torch_ext.cpp); from then on
        dropping this object only drops one reference."""
        if self._ref is None:
            self._ref = ext.EngineRef(int(self._h.value))
        return self._ref

    def close(self):
        """Destroy the native engine now.  A backward pass of an earlier evaluation that runs
        after this raises an error."""
        if self._h is not None:
            if self._ref is not None:
                self._ref.close()
                self._ref = None
            else:
                self._lib.ttb2_destroy(self._h)
            self._h = None

    def release(self):
        """Drop this object's reference without destroying an engine that pending autograd
        nodes still hold (TreeLikelihoodModel replaces its engine this way)."""
        if self._h is not None and self._ref is not None:
            self._ref = None
            self._h = None
        else:
            self.close()

    def __del__(self):
        try:
            self.release()
        except Exception:
            pass

    def set_postorder(self, postorder):
        postorder = np.ascontiguousarray(postorder, dtype=np.int32)
        _lib.check(self._lib.ttb2_set_postorder(
            self._h, postorder.ctypes.data_as(ctypes.c_void_p)), "ttb2_set_postorder")

    def set_stream(self, cuda_stream: int):
        _lib.check(self._lib.ttb2_set_stream(self._h, ctypes.c_void_p(cuda_stream)),
                   "ttb2_set_stream")

    def synchronize(self):
        _lib.check(self._lib.ttb2_synchronize(self._h), "ttb2_synchronize")

    def enable_timing(self, on: bool = True):
        _lib.check(self._lib.ttb2_enable_timing(self._h, int(on)), "ttb2_enable_timing")

    def phase_ms(self) -> dict:
        """Durations (ms) of the kernel groups of the latest loglik / grad calls."""
        buf = (ctypes.c_double * 7)()
        _lib.check(self._lib.ttb2_phase_ms(self._h, ctypes.cast(buf, ctypes.c_void_p)),
                   "ttb2_phase_ms")
        keys = ("pmatrix", "postorder", "root", "preorder", "contract")
        out = {k: buf[i] for i, k in enumerate(keys)}
        out["postorder_launches"] = int(buf[5])
        out["preorder_launches"] = int(buf[6])
        return out

    @property
    def eval_serial(self) -> int:
        """Number of loglik calls served so far (ttb2_eval_serial): a deferred gradient compares
        it with the value it saw after its own forward."""
        return int(self._lib.ttb2_eval_serial(self._h))

    @property
    def launch_count(self) -> int:
        return int(self._lib.ttb2_launch_count(self._h))

    @property
    def device_bytes(self) -> int:
        return int(self._lib.ttb2_device_bytes(self._h))

    # -- helpers ------------------------------------------------------------
    def _prep(self, x, shape_tail, name):
        """-> contiguous fp64 tensor [d, *shape_tail] (d = leading draws dim)."""
        if not isinstance(x, torch.Tensor):
            x = torch.from_numpy(np.ascontiguousarray(x, dtype=np.float64))
        x = x.detach()
        if x.dtype != torch.float64:
            x = x.to(torch.float64)
        nt = len(shape_tail)
        if x.dim() == nt:
            x = x.unsqueeze(0)
        if x.dim() != nt + 1 or tuple(x.shape[1:]) != tuple(shape_tail):
            raise EngineError("%s: expected shape [draws,%s], got %s" % (
                name, ",".join(map(str, shape_tail)), tuple(x.shape)))
        return x.contiguous()

    def _where(self, tensors):
        cuda = [t.is_cuda for t in tensors]
        if all(cuda):
            for t in tensors:
                if t.device.index != self.device:
                    raise EngineError("input tensor is on a different CUDA device")
            return TTB2_DEVICE, torch.device("cuda", self.device)
        if any(cuda):
            raise EngineError("inputs must be all host or all device tensors")
        return TTB2_HOST, torch.device("cpu")

    # -- P-mode -------------------------------------------------------------
    def loglik_mats(self, mats, freqs, props, out=None) -> torch.Tensor:
        """lnL [D] from caller-supplied matrices [D,B,K,S,S] (ttb2_loglik_mats)."""
        S, K, B = self.S, self.K, self.B
        mats = self._prep(mats, (B, K, S, S), "mats")
        freqs = self._prep(freqs, (S,), "freqs")
        props = self._prep(props, (K,), "props")
        D = mats.shape[0]
        where, dev = self._where([mats, freqs, props])
        lnl = out if out is not None else torch.empty(D, dtype=torch.float64, device=dev)
        _lib.check(self._lib.ttb2_loglik_mats(
            self._h, D, _ptr(mats), _ptr(freqs), freqs.shape[0], _ptr(props), props.shape[0],
            _ptr(lnl), where), "ttb2_loglik_mats")
        self._draws = D
        self._shapes = dict(where=where, dev=dev, D=D, fd=freqs.shape[0], pd=props.shape[0])
        return lnl

    def grad_mats(self, grad_lnl=None, want_mats=True):
        sh = self._shapes
        if sh is None:
            raise EngineError("grad_mats before loglik")
        S, K, B, D, dev = self.S, self.K, self.B, sh["D"], sh["dev"]
        g = None
        if grad_lnl is not None:
            g = self._prep(grad_lnl.reshape(-1), (), "grad_lnl").reshape(-1).to(dev)
        d_mats = torch.empty((D, B, K, S, S), dtype=torch.float64, device=dev) if want_mats else None
        d_freqs = torch.empty((sh["fd"], S), dtype=torch.float64, device=dev)
        d_props = torch.empty((sh["pd"], K), dtype=torch.float64, device=dev)
        _lib.check(self._lib.ttb2_grad_mats(
            self._h, _ptr(g), _ptr(d_mats), _ptr(d_freqs), _ptr(d_props), sh["where"]),
            "ttb2_grad_mats")
        return d_mats, d_freqs, d_props

    # -- eigen mode -----------------------------------------------------------
    def loglik_eigen(self, branch_lengths, site_rates, props, evec, ivec, evals, freqs,
                     out=None) -> torch.Tensor:
        """lnL [D]; P = V exp(L r t) V^-1 computed on the device (ttb2_loglik_eigen)."""
        S, K, B = self.S, self.K, self.B
        bl = self._prep(branch_lengths, (B,), "branch_lengths")
        rates = self._prep(site_rates, (K,), "site_rates")
        props = self._prep(props, (K,), "props")
        evec = self._prep(evec, (S, S), "evec")
        ivec = self._prep(ivec, (S, S), "ivec")
        evals = self._prep(evals, (S,), "evals")
        freqs = self._prep(freqs, (S,), "freqs")
        if not (evec.shape[0] == ivec.shape[0] == evals.shape[0]):
            raise EngineError("evec/ivec/evals must share their draws dimension")
        D = bl.shape[0]
        where, dev = self._where([bl, rates, props, evec, ivec, evals, freqs])
        lnl = out if out is not None else torch.empty(D, dtype=torch.float64, device=dev)
        _lib.check(self._lib.ttb2_loglik_eigen(
            self._h, D, _ptr(bl), _ptr(rates), rates.shape[0], _ptr(props), props.shape[0],
            _ptr(evec), _ptr(ivec), _ptr(evals), evec.shape[0], _ptr(freqs), freqs.shape[0],
            _ptr(lnl), where), "ttb2_loglik_eigen")
        self._draws = D
        self._shapes = dict(where=where, dev=dev, D=D, fd=freqs.shape[0], pd=props.shape[0],
                            rd=rates.shape[0], ed=evec.shape[0])
        return lnl

    def loglik_q(self, branch_lengths, site_rates, props, q_norm, freqs, out=None) -> torch.Tensor:
        """lnL [D] from the normalised reversible generator; its eigen-system is computed on
        the device (ttb2_loglik_q, csrc/eigen.cu).  `grad_eigen` follows as usual."""
        S, K, B = self.S, self.K, self.B
        bl = self._prep(branch_lengths, (B,), "branch_lengths")
        rates = self._prep(site_rates, (K,), "site_rates")
        props = self._prep(props, (K,), "props")
        q = self._prep(q_norm, (S, S), "q_norm")
        freqs = self._prep(freqs, (S,), "freqs")
        D = bl.shape[0]
        where, dev = self._where([bl, rates, props, q, freqs])
        lnl = out if out is not None else torch.empty(D, dtype=torch.float64, device=dev)
        _lib.check(self._lib.ttb2_loglik_q(
            self._h, D, _ptr(bl), _ptr(rates), rates.shape[0], _ptr(props), props.shape[0],
            _ptr(q), q.shape[0], _ptr(freqs), freqs.shape[0], _ptr(lnl), where), "ttb2_loglik_q")
        self._draws = D
        self._shapes = dict(where=where, dev=dev, D=D, fd=freqs.shape[0], pd=props.shape[0],
                            rd=rates.shape[0], ed=max(q.shape[0], freqs.shape[0]))
        return lnl

    def loglik_expm(self, branch_lengths, site_rates, props, q, freqs, out=None) -> torch.Tensor:
        """lnL [D] for a general (non-reversible) generator: P = exp(Q r t) on the device
        (ttb2_loglik_expm, csrc/expm.cu).  `grad_eigen` follows as usual."""
        S, K, B = self.S, self.K, self.B
        bl = self._prep(branch_lengths, (B,), "branch_lengths")
        rates = self._prep(site_rates, (K,), "site_rates")
        props = self._prep(props, (K,), "props")
        q = self._prep(q, (S, S), "q")
        freqs = self._prep(freqs, (S,), "freqs")
        D = bl.shape[0]
        where, dev = self._where([bl, rates, props, q, freqs])
        lnl = out if out is not None else torch.empty(D, dtype=torch.float64, device=dev)
        _lib.check(self._lib.ttb2_loglik_expm(
            self._h, D, _ptr(bl), _ptr(rates), rates.shape[0], _ptr(props), props.shape[0],
            _ptr(q), q.shape[0], _ptr(freqs), freqs.shape[0], _ptr(lnl), where),
            "ttb2_loglik_expm")
        self._draws = D
        self._shapes = dict(where=where, dev=dev, D=D, fd=freqs.shape[0], pd=props.shape[0],
                            rd=rates.shape[0], ed=q.shape[0])
        return lnl

    def get_eigen(self):
        """(evec, ivec, evals) of the latest eigen-mode call, as the engine holds them."""
        sh = self._shapes
        if sh is None or "ed" not in sh:
            raise EngineError("get_eigen before an eigen-mode loglik call")
        S, ed, dev = self.S, sh["ed"], sh["dev"]
        evec = torch.empty((ed, S, S), dtype=torch.float64, device=dev)
        ivec = torch.empty((ed, S, S), dtype=torch.float64, device=dev)
        evals = torch.empty((ed, S), dtype=torch.float64, device=dev)
        _lib.check(self._lib.ttb2_get_eigen(self._h, _ptr(evec), _ptr(ivec), _ptr(evals),
                                            sh["where"]), "ttb2_get_eigen")
        return evec, ivec, evals

    def grad_eigen(self, grad_lnl=None, out=None):
        sh = self._shapes
        if sh is None or "ed" not in sh:
            raise EngineError("grad_eigen before loglik_eigen")
        S, K, B, D, dev = self.S, self.K, self.B, sh["D"], sh["dev"]
        g = None
        if grad_lnl is not None:
            g = self._prep(grad_lnl.reshape(-1), (), "grad_lnl").reshape(-1).to(dev)
        if out is None:
            out = dict(
                branch_lengths=torch.empty((D, B), dtype=torch.float64, device=dev),
                site_rates=torch.empty((sh["rd"], K), dtype=torch.float64, device=dev),
                props=torch.empty((sh["pd"], K), dtype=torch.float64, device=dev),
                q=torch.empty((sh["ed"], S, S), dtype=torch.float64, device=dev),
                freqs=torch.empty((sh["fd"], S), dtype=torch.float64, device=dev),
            )
        _lib.check(self._lib.ttb2_grad_eigen(
            self._h, _ptr(g), _ptr(out["branch_lengths"]), _ptr(out["site_rates"]),
            _ptr(out["props"]), _ptr(out.get("q")), _ptr(out["freqs"]), sh["where"]),
            "ttb2_grad_eigen")
        return out

    def grad_eigen_packed(self, grad_lnl=None, out=None) -> torch.Tensor:
        """The gradient of the latest eigen-mode call as one vector
        [lnL | d_bl | d_rates | d_props | d_q | d_freqs] (ttb2_grad_eigen_packed): the payload
        of the all-reduce of a pattern-sharded run.  `unpack` gives the views."""
        sh = self._shapes
        if sh is None or "ed" not in sh:
            raise EngineError("grad_eigen_packed before loglik_eigen")
        dev = sh["dev"]
        g = None
        if grad_lnl is not None:
            g = self._prep(grad_lnl.reshape(-1), (), "grad_lnl").reshape(-1).to(dev)
        count = int(self._lib.ttb2_packed_count(self._h))
        if out is None:
            out = torch.empty(count, dtype=torch.float64, device=dev)
        _lib.check(self._lib.ttb2_grad_eigen_packed(self._h, _ptr(g), _ptr(out), out.numel(),
                                                    sh["where"]), "ttb2_grad_eigen_packed")
        return out

    def unpack(self, packed: torch.Tensor) -> dict:
        """Views of a packed gradient vector, keyed like `grad_eigen`'s result (+ "lnL")."""
        sh, S, K, B = self._shapes, self.S, self.K, self.B
        D = sh["D"]
        out, off = {}, 0
        for key, shape in (("lnL", (D,)), ("branch_lengths", (D, B)), ("site_rates", (sh["rd"], K)),
                           ("props", (sh["pd"], K)), ("q", (sh["ed"], S, S)),
                           ("freqs", (sh["fd"], S))):
            n = 1
            for x in shape:
                n *= x
            out[key] = packed[off:off + n].view(shape)
            off += n
        return out

    # -- inspection -----------------------------------------------------------
    def site_loglik(self) -> torch.Tensor:
        sh = self._shapes
        out = torch.empty((sh["D"], self.N), dtype=torch.float64, device=sh["dev"])
        _lib.check(self._lib.ttb2_site_loglik(self._h, _ptr(out), sh["where"]), "ttb2_site_loglik")
        return out

    def get_mats(self) -> torch.Tensor:
        sh = self._shapes
        out = torch.empty((sh["D"], self.B, self.K, self.S, self.S), dtype=torch.float64,
                          device=sh["dev"])
        _lib.check(self._lib.ttb2_get_mats(self._h, _ptr(out), sh["where"]), "ttb2_get_mats")
        return out
