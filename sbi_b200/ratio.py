"""Ratio estimator (NRE) at sbi's estimator boundary, backed by the sm_100a kernels.

`RatioEstimator` mirrors /root/reference/sbi/neural_nets/ratio_estimators.py:11-157 for the
`resnet` classifier of /root/reference/sbi/neural_nets/net_builders/classifier.py:172-235:
`forward(theta, x)` / `unnormalized_log_ratio` return logits of shape `(*batch_shape)` for
equally-prefixed `theta` and `x` (no broadcasting, same error), `state_dict()` uses the
reference's keys.  `classifier_nn` / `build_resnet_classifier` mirror factory.py:174-241.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Any, Callable, Optional

import torch
from torch import Tensor, nn
from torch.nn import init

from . import _lib as L
from .estimators import Standardize
from .neural_nets import _linear_init, check_data_device, standardizing_stats, z_score_parser
from .pack import RatioLayout


class _RatioNet(nn.Module):
    """Sits at `estimator.net` (the reference's ResidualNet); owns the flat parameter buffer."""

    def __init__(self, layout: RatioLayout):
        super().__init__()
        self.layout = layout
        self.flat = nn.Parameter(torch.zeros(layout.n_params, dtype=torch.float32))
        self.register_buffer("_tab", torch.from_numpy(layout.tab.copy()), persistent=False)
        self.register_buffer("_mask", layout.trainable_mask(), persistent=False)
        self.hidden_features = layout.H

    def _save_to_state_dict(self, destination, prefix, keep_vars):
        for k, t in self.layout.unpack(self.flat).items():
            destination[prefix + k[len("net."):]] = t

    def _load_from_state_dict(self, state_dict, prefix, local_metadata, strict, missing_keys,
                              unexpected_keys, error_msgs):
        lay = self.layout
        if prefix + "flat" in state_dict:
            with torch.no_grad():
                self.flat.copy_(state_dict.pop(prefix + "flat"))
            return
        src = {}
        for k in lay.index:
            kk = prefix + k[len("net."):]
            if kk in state_dict:
                src[k] = state_dict.pop(kk)
            elif strict:
                missing_keys.append(kk)
        if len(src) == len(lay.index):
            with torch.no_grad():
                lay.pack(src, out=self.flat.data)


class RatioEstimator(nn.Module):
    r"""log r(theta, x) = classifier logit; trained by NRE (ratio_estimators.py:11-157)."""

    def __init__(self, layout: RatioLayout, theta_shape, x_shape, theta_stats, x_stats,
                 embedding_net_theta: nn.Module = None, embedding_net_x: nn.Module = None):
        super().__init__()
        self._input_shape = torch.Size(theta_shape)
        self._condition_shape = torch.Size(x_shape)
        self.theta_shape, self.x_shape = self._input_shape, self._condition_shape
        et = embedding_net_theta if embedding_net_theta is not None else nn.Identity()
        ex = embedding_net_x if embedding_net_x is not None else nn.Identity()
        if not isinstance(et, nn.Identity) or not isinstance(ex, nn.Identity):
            raise NotImplementedError("the sm_100a ratio kernels take nn.Identity() embedding nets")
        self.embedding_net_theta = nn.Sequential(Standardize(*theta_stats), et) if theta_stats else et
        self.embedding_net_x = nn.Sequential(Standardize(*x_stats), ex) if x_stats else ex
        self.net = _RatioNet(layout)
        self._cache = {}

    input_shape = property(lambda self: self._input_shape)
    condition_shape = property(lambda self: self._condition_shape)
    layout = property(lambda self: self.net.layout)
    flat = property(lambda self: self.net.flat)

    def __deepcopy__(self, memo):
        import copy
        new = self.__class__.__new__(self.__class__)
        memo[id(self)] = new
        for k, v in self.__dict__.items():
            new.__dict__[k] = {} if k == "_cache" else copy.deepcopy(v, memo)
        return new

    def __getstate__(self):
        d = dict(self.__dict__)
        d["_cache"] = {}
        return d

    # ---- kernel views
    def _stats(self) -> Tensor:
        lay = self.layout
        srcs = []
        for emb in (self.embedding_net_theta, self.embedding_net_x):
            if isinstance(emb, nn.Sequential) and isinstance(emb[0], Standardize):
                srcs += [emb[0]._mean, emb[0]._std]
        key = tuple((t.data_ptr(), t._version) for t in srcs) + (str(self.net.flat.device),)
        hit = self._cache.get("stats")
        if hit is not None and hit[0] == key:
            return hit[1]
        dev = self.net.flat.device
        st = torch.zeros(2 * lay.Dtp + 2 * lay.Dxp, dtype=torch.float32, device=dev)
        st[lay.Dtp:2 * lay.Dtp] = 1.0
        st[2 * lay.Dtp + lay.Dxp:] = 1.0
        emb = self.embedding_net_theta
        if isinstance(emb, nn.Sequential):
            st[:lay.Dt] = emb[0]._mean.reshape(-1).expand(lay.Dt)
            st[lay.Dtp:lay.Dtp + lay.Dt] = emb[0]._std.reshape(-1).expand(lay.Dt)
        emb = self.embedding_net_x
        if isinstance(emb, nn.Sequential):
            st[2 * lay.Dtp:2 * lay.Dtp + lay.Dx] = emb[0]._mean.reshape(-1).expand(lay.Dx)
            st[2 * lay.Dtp + lay.Dxp:2 * lay.Dtp + lay.Dxp + lay.Dx] = emb[0]._std.reshape(-1).expand(lay.Dx)
        self._cache["stats"] = (key, st)
        return st

    def _model(self, nbuf: int) -> L.RatioModel:
        L.require_cuda(self.net.flat, "estimator parameters")
        st = self._stats()
        s = L.RatioModel()
        self.layout.fill_struct(s, nbuf)
        s.d_params = self.net.flat.data_ptr()
        s.d_tab = self.net._tab.data_ptr()
        s.d_stats = st.data_ptr()
        s._keep = (st,)
        return s

    def _gpart(self, n_part: int) -> Tensor:
        buf = self._cache.get("gpart")
        if buf is None or buf.shape[0] < n_part or buf.device != self.net.flat.device:
            buf = torch.zeros(max(n_part, 1), self.layout.n_params, dtype=torch.float32,
                              device=self.net.flat.device)
            self._cache["gpart"] = buf
        return buf

    # ---- shape checks: ratio_estimators.py:53-112
    def _check(self, theta: Tensor, x: Tensor):
        if theta.shape[-len(self.theta_shape):] != self.theta_shape:
            raise ValueError(f"The trailing dimensions of `theta` do not match the `theta_shape`: "
                             f"{theta.shape[-len(self.theta_shape):]} != {self.theta_shape}.")
        if x.shape[-len(self.x_shape):] != self.x_shape:
            raise ValueError(f"The trailing dimensions of `x` do not match the `x_shape`: "
                             f"{x.shape[-len(self.x_shape):]} != {self.x_shape}.")
        tp, xp = theta.shape[:-len(self.theta_shape)], x.shape[:-len(self.x_shape)]
        if tp != xp:
            raise ValueError(f"The shape prefixes of `theta` and `x` must match: {tuple(tp)=} != "
                             f"{tuple(xp)=}. Make them agree, since we do not broadcast for you.")
        return tp

    def unnormalized_log_ratio(self, theta: Tensor, x: Tensor) -> Tensor:
        prefix = self._check(theta, x)
        th = theta.reshape(-1, self.layout.Dt).contiguous().float()
        xx = x.reshape(-1, self.layout.Dx).contiguous().float()
        return _RatioFn.apply(self.net.flat, th, xx, self, None, None, False).reshape(*prefix)

    def forward(self, *args, **kwargs) -> Tensor:
        return self.unnormalized_log_ratio(*args, **kwargs)

    def loss(self, input: Tensor, condition: Tensor, **kwargs) -> Tensor:
        raise NotImplementedError()

    # ---- raw entry (no autograd): pairs given by optional index arrays / shared x
    def logits_raw(self, theta: Tensor, x: Tensor, ti: Optional[Tensor] = None,
                   xi: Optional[Tensor] = None, x_shared: bool = False, R: Optional[int] = None) -> Tensor:
        lib = L.load()
        L.require_cuda(theta, "theta")
        L.require_cuda(x, "x")
        R = (ti.shape[0] if ti is not None else theta.shape[0]) if R is None else R
        out = torch.empty(R, dtype=torch.float32, device=theta.device)
        m = self._model(nbuf=2)
        pr = L.Pairs(theta.data_ptr(), x.data_ptr(), None if ti is None else ti.data_ptr(),
                     None if xi is None else xi.data_ptr(), R, 1 if x_shared else 0)
        tc = self._tc_state(m) if (R >= self.TC_MIN_ROWS or os.environ.get("SBI_B200_TC", "") == "1") else None
        if tc is not None:
            L.check(lib.sbi_b200_ratio_forward_tc(C.byref(m), C.byref(tc), C.byref(pr), L.ptr(out),
                                                  L.stream_ptr()), "ratio_forward_tc")
            return out
        L.check(lib.sbi_b200_ratio_forward(C.byref(m), C.byref(pr), L.ptr(out), L.stream_ptr()), "ratio_forward")
        return out

    #: pairs from which the logits go through the tcgen05 kernel (csrc/ratio_tc.cu); measured
    #: (profiles/tc_ratio_time.py): 10 k pairs 29 us SIMT vs 35 us, 131 k pairs 140 vs 58 us
    TC_MIN_ROWS = int(os.environ.get("SBI_B200_RATIO_TC_MIN_ROWS", 32768))

    def _tc_state(self, m):
        """`NsfTc` descriptor with freshly packed operands, or None if the model is outside what
        the tensor-core kernel instantiates (see FlowEstimator._tc_state)."""
        if os.environ.get("SBI_B200_TC", "") == "0":
            return None
        flat = self.net.flat
        st = self._cache.get("tc")
        if st is None or st["dev"] != flat.device:
            plan = self.layout.tc_plan()
            st = {"dev": flat.device, "plan": plan}
            if plan is not None:
                st.update(src=torch.as_tensor(plan["src"], device=flat.device),
                          tab=torch.as_tensor(plan["tab"], device=flat.device),
                          tcw=torch.empty(plan["n_words"], dtype=torch.float32, device=flat.device))
            self._cache["tc"] = st
        if st["plan"] is None:
            return None
        tc = L.NsfTc(st["plan"]["n_words"], st["plan"]["stage_cap"], st["src"].data_ptr(),
                     st["tab"].data_ptr(), st["tcw"].data_ptr())
        lib = L.load()
        if not lib.sbi_b200_ratio_tc_supported(C.byref(m), C.byref(tc)):
            return None
        L.check(lib.sbi_b200_ratio_tc_pack(C.byref(m), C.byref(tc), L.stream_ptr()), "ratio_tc_pack")
        return tc


class _RatioFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, flat, theta, x, est: RatioEstimator, ti, xi, x_shared):
        out = est.logits_raw(theta, x, ti, xi, x_shared)
        ctx.save_for_backward(theta, x)
        ctx.est, ctx.ti, ctx.xi, ctx.x_shared = est, ti, xi, x_shared
        return out

    @staticmethod
    def backward(ctx, g):
        theta, x = ctx.saved_tensors
        est = ctx.est
        lib = L.load()
        R = g.shape[0]
        n_part = lib.sbi_b200_ratio_vjp_parts(R)
        gpart = est._gpart(n_part)
        need_flat, need_th = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        gth = torch.empty(R, est.layout.Dt, dtype=torch.float32, device=theta.device) if need_th else None
        m = est._model(nbuf=3)
        pr = L.Pairs(theta.data_ptr(), x.data_ptr(), None if ctx.ti is None else ctx.ti.data_ptr(),
                     None if ctx.xi is None else ctx.xi.data_ptr(), R, 1 if ctx.x_shared else 0)
        g = g.contiguous().float()
        L.check(lib.sbi_b200_ratio_vjp(C.byref(m), C.byref(pr), L.ptr(g), None, L.ptr(gpart), L.ptr(gth),
                                       L.stream_ptr()), "ratio_vjp")
        gflat = None
        if need_flat:
            gflat = torch.empty(est.layout.n_params, dtype=torch.float32, device=theta.device)
            L.check(lib.sbi_b200_reduce_partials(L.ptr(gpart), n_part, est.layout.n_params, L.ptr(gflat),
                                                 L.stream_ptr()), "reduce_partials")
        if need_th and ctx.ti is not None:
            raise NotImplementedError("theta gradients with an index gather are not needed by any caller")
        return gflat, gth, None, None, None, None, None


def build_resnet_classifier(
    batch_x: Tensor, batch_y: Tensor, z_score_x: Optional[str] = "independent",
    z_score_y: Optional[str] = "independent", hidden_features: int = 50,
    embedding_net_x: nn.Module = nn.Identity(), embedding_net_y: nn.Module = nn.Identity(),
    num_blocks: int = 2, dropout_probability: float = 0.0, use_batch_norm: bool = False,
) -> RatioEstimator:
    """classifier.py:172-235 (in the classifier's view x = theta, y = x).  Parameters are drawn in
    nflows' construction order (initial layer, per block two linears with the second re-drawn
    U(-1e-3, 1e-3), final layer), so a seed reproduces the reference's initial weights."""
    check_data_device(batch_x, batch_y)
    if z_score_x == "transform_to_unconstrained":
        raise ValueError("Ratio-based classifiers (NRE) do not implement `transform_to_unconstrained`.")
    if dropout_probability != 0.0 or use_batch_norm:
        raise NotImplementedError("dropout / batch norm are not implemented in the sm_100a ratio kernels")
    Dt, Dx, H = batch_x[0].numel(), batch_y[0].numel(), hidden_features
    lay = RatioLayout(Dt=Dt, Dx=Dx, H=H, NB=num_blocks)
    state = {}
    state["net.initial_layer.weight"], state["net.initial_layer.bias"] = _linear_init(H, Dt + Dx)
    for b in range(num_blocks):
        state[f"net.blocks.{b}.linear_layers.0.weight"], state[f"net.blocks.{b}.linear_layers.0.bias"] = _linear_init(H, H)
        w, bb = _linear_init(H, H)
        init.uniform_(w, -1e-3, 1e-3)
        init.uniform_(bb, -1e-3, 1e-3)
        state[f"net.blocks.{b}.linear_layers.1.weight"], state[f"net.blocks.{b}.linear_layers.1.bias"] = w, bb
    state["net.final_layer.weight"], state["net.final_layer.bias"] = _linear_init(1, H)
    zx, sx = z_score_parser(z_score_x)
    zy, sy = z_score_parser(z_score_y)
    t_stats = standardizing_stats(batch_x, sx) if zx else None
    x_stats = standardizing_stats(batch_y, sy) if zy else None
    est = RatioEstimator(lay, batch_x[0].shape, batch_y[0].shape, t_stats, x_stats,
                         embedding_net_x, embedding_net_y)
    with torch.no_grad():
        lay.pack(state, out=est.net.flat.data)
    return est


def classifier_nn(
    model: str, z_score_theta: Optional[str] = "independent", z_score_x: Optional[str] = "independent",
    hidden_features: int = 50, embedding_net_theta: nn.Module = nn.Identity(),
    embedding_net_x: nn.Module = nn.Identity(), **kwargs: Any,
) -> Callable:
    """factory.py:174-241: build function for the NRE classifier."""
    if model != "resnet":
        raise NotImplementedError(f"sbi_b200 implements the 'resnet' classifier on sm_100a; got {model!r}.")

    def build_fn(batch_theta, batch_x):
        from ._refabc import register_with_reference
        register_with_reference()
        return build_resnet_classifier(
            batch_x=batch_theta, batch_y=batch_x, z_score_x=z_score_theta, z_score_y=z_score_x,
            hidden_features=hidden_features, embedding_net_x=embedding_net_theta,
            embedding_net_y=embedding_net_x, **kwargs)

    return build_fn
