"""Estimator modules at sbi's estimator boundary, backed by the sm_100a kernels.

`NSFEstimator` mirrors `sbi.neural_nets.estimators.NFlowsFlow`
(/root/reference/sbi/neural_nets/estimators/nflows_flow.py:14-151) on the shape rules of
`ConditionalDensityEstimator` (/root/reference/sbi/neural_nets/estimators/base.py:35-306):
same method names, argument meaning, output shapes and error behaviour, and a
`state_dict()` with the reference's own keys, so reference checkpoints load verbatim.
All parameters live in ONE flat `nn.Parameter` (the packed layout of `pack.NsfLayout`);
`log_prob` is a `torch.autograd.Function` whose forward is the fused log-prob kernel and
whose backward is the fused forward+backward (VJP) kernel.  No CPU path exists.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Tuple

import torch
from torch import Tensor, nn

from . import _lib as L
from .pack import MafLayout, NsfLayout


class _Family:
    """C-ABI entry points + model struct of one flow family (include/sbi_b200.h)."""

    def __init__(self, name, struct, tab_fields, c_prefix=None):
        self.name, self.struct, self.tab_fields = name, struct, tab_fields
        self.c_prefix = c_prefix or name

    def fn(self, what):
        return getattr(L.load(), f"sbi_b200_{self.c_prefix}_{what}")


FAMILIES = {
    "nsf": _Family("nsf", L.NsfModel, ("d_layer_tab", "d_feat_tab")),
    "maf": _Family("maf", L.MafModel, ("d_layer_tab", "d_perm_tab")),
    # `made`: the masked residual conditioner + mixture head run on the NSF kernels (head = SBI_NSF_MOG)
    "made": _Family("made", L.NsfModel, ("d_layer_tab", "d_feat_tab"), c_prefix="nsf"),
}


class Standardize(nn.Module):
    """(t - mean) / std -- mirrors sbi.utils.sbiutils.Standardize (sbiutils.py:418-428)."""

    def __init__(self, mean, std):
        super().__init__()
        mean, std = map(torch.as_tensor, (mean, std))
        self.register_buffer("_mean", mean.clone().float())
        self.register_buffer("_std", std.clone().float())

    def forward(self, tensor):
        return (tensor - self._mean) / self._std


class _FlowNet(nn.Module):
    """Plays the role of the nflows `Flow` object that sits at `estimator.net`."""

    def __init__(self, layout, shift: Tensor, scale: Tensor, embedding_net: nn.Module):
        super().__init__()
        self.layout = layout
        self.flat = nn.Parameter(torch.zeros(layout.n_params, dtype=torch.float32))
        self.register_buffer("_shift", shift.clone().float(), persistent=False)
        self.register_buffer("_scale", scale.clone().float(), persistent=False)
        tab_a, tab_b = layout.tables()
        self.register_buffer("_layer_tab", torch.from_numpy(tab_a.copy()), persistent=False)
        self.register_buffer("_feat_tab", torch.from_numpy(tab_b.copy()), persistent=False)
        self.register_buffer("_mask", layout.trainable_mask(), persistent=False)
        # masked-out raw MADE weights, kept only so that state_dict() round-trips exactly
        self.register_buffer("_raw", torch.zeros(layout.n_params if layout._wm() else 1), persistent=False)
        self._embedding_net = embedding_net

    # -- reference-compatible (de)serialisation ---------------------------------------------
    def _save_to_state_dict(self, destination, prefix, keep_vars):
        lay = self.layout
        if lay.zscore_input:
            destination[prefix + "_transform._transforms.0._shift"] = self._shift.detach().clone()
            destination[prefix + "_transform._transforms.0._scale"] = self._scale.detach().clone()
        for k, t in lay.unpack(self.flat, self._raw if lay._wm() else None).items():
            destination[prefix + k[len("net."):]] = t
        for k, t in lay.buffers.items():
            destination[prefix + k[len("net."):]] = t.to(self.flat.device)

    def _load_from_state_dict(self, state_dict, prefix, local_metadata, strict, missing_keys,
                              unexpected_keys, error_msgs):
        lay = self.layout
        if prefix + "flat" in state_dict:   # native format
            with torch.no_grad():
                self.flat.copy_(state_dict.pop(prefix + "flat"))
        else:
            src = {}
            for k in lay.index:
                kk = prefix + k[len("net."):]
                if kk in state_dict:
                    src[k] = state_dict.pop(kk)
                elif strict:
                    missing_keys.append(kk)
            if len(src) == len(lay.index):
                with torch.no_grad():
                    lay.pack(src, out=self.flat.data, raw_out=self._raw if lay._wm() else None)
        for name, buf in (("_shift", self._shift), ("_scale", self._scale)):
            kk = prefix + "_transform._transforms.0." + name
            if kk in state_dict:
                with torch.no_grad():
                    buf.copy_(state_dict.pop(kk))
        incoming = {}
        for k in lay.buffers:
            kk = prefix + k[len("net."):]
            if kk in state_dict:
                incoming[k] = state_dict.pop(kk)
        if incoming and hasattr(lay, "load_buffers"):
            # structural buffers that are data (MAF permutations): adopt them
            new_tab_b = lay.load_buffers(incoming)
            if new_tab_b is not None:
                with torch.no_grad():
                    self._feat_tab.copy_(torch.from_numpy(new_tab_b).to(self._feat_tab.device))


def _is_identity(m: nn.Module) -> bool:
    return isinstance(m, nn.Identity)


class FlowEstimator(nn.Module):
    r"""Normalizing flow q(input | condition) evaluated by hand-written sm_100a kernels
    (families: neural spline flow `nsf`, masked autoregressive flow `maf`)."""

    def __init__(self, layout, input_shape, condition_shape, shift: Tensor,
                 scale: Tensor, cond_mean: Optional[Tensor], cond_std: Optional[Tensor],
                 embedding_net: Optional[nn.Module] = None):
        super().__init__()
        self._input_shape = torch.Size(input_shape)
        self._condition_shape = torch.Size(condition_shape)
        user_net = embedding_net if embedding_net is not None else nn.Identity()
        self._embed_identity = _is_identity(user_net)
        if cond_mean is not None:
            emb = nn.Sequential(Standardize(cond_mean, cond_std), user_net)
        else:
            emb = user_net
        self.net = _FlowNet(layout, shift, scale, emb)
        self._cache = {}
        self.fam = FAMILIES[layout.family]

    # ---- properties of the reference interface ---------------------------------------------
    @property
    def layout(self):
        return self.net.layout

    @property
    def input_shape(self) -> torch.Size:
        return self._input_shape

    @property
    def condition_shape(self) -> torch.Size:
        return self._condition_shape

    @property
    def embedding_net(self) -> nn.Module:
        return self.net._embedding_net

    @property
    def flat(self) -> nn.Parameter:
        return self.net.flat

    def __deepcopy__(self, memo):
        import copy
        cls = self.__class__
        new = cls.__new__(cls)
        memo[id(self)] = new
        for k, v in self.__dict__.items():
            if k == "_cache":
                new.__dict__[k] = {}
            else:
                new.__dict__[k] = copy.deepcopy(v, memo)
        return new

    def __getstate__(self):
        d = dict(self.__dict__)
        d["_cache"] = {}
        return d

    # ---- kernel-side views --------------------------------------------------------------------
    def _kernel_stats(self, raw_condition: bool = False) -> Tuple[Tensor, float]:
        """[shift(Dp) | scale(Dp) | ctx_mean(Cp) | ctx_std(Cp)] on the parameter device.
        raw_condition: identity statistics for the condition (see inverse_transform)."""
        lay = self.layout
        net = self.net
        emb = net._embedding_net
        std_mod = emb[0] if isinstance(emb, nn.Sequential) and isinstance(emb[0], Standardize) else None
        srcs = [net._shift, net._scale] + ([std_mod._mean, std_mod._std] if std_mod is not None else [])
        key = tuple((t.data_ptr(), t._version) for t in srcs) + (str(net.flat.device), raw_condition)
        ck = "stats_raw" if raw_condition else "stats"
        hit = self._cache.get(ck)
        if hit is not None and hit[0] == key:
            return hit[1], hit[2]
        dev = net.flat.device
        st = torch.zeros(2 * lay.Dp + 2 * lay.Cp, dtype=torch.float32, device=dev)
        st[lay.Dp:2 * lay.Dp] = 1.0
        st[2 * lay.Dp + lay.Cp:] = 1.0
        st[:lay.D] = net._shift.expand(lay.D)
        st[lay.Dp:lay.Dp + lay.D] = net._scale.expand(lay.D)
        if std_mod is not None and self._embed_identity and not raw_condition:
            st[2 * lay.Dp:2 * lay.Dp + lay.C] = std_mod._mean.reshape(-1).expand(lay.C)
            st[2 * lay.Dp + lay.Cp:2 * lay.Dp + lay.Cp + lay.C] = std_mod._std.reshape(-1).expand(lay.C)
        ld = float(torch.log(torch.abs(net._scale.double())).expand(lay.D).sum())
        self._cache[ck] = (key, st, ld)
        return st, ld

    def _model(self, nbuf: int, raw_condition: bool = False):
        net = self.net
        L.require_cuda(net.flat, "estimator parameters")
        st, ld = self._kernel_stats(raw_condition)
        s = self.fam.struct()
        self.layout.fill_struct(s, nbuf)
        s.ld_zscore = ld
        s.d_params = net.flat.data_ptr()
        setattr(s, self.fam.tab_fields[0], net._layer_tab.data_ptr())
        setattr(s, self.fam.tab_fields[1], net._feat_tab.data_ptr())
        s.d_stats = st.data_ptr()
        s._keep = (st,)
        return s

    def _embed(self, condition: Tensor) -> Tensor:
        """Context fed to the kernels.  Identity embedding: raw condition (standardised
        in-kernel); otherwise the torch embedding net runs first."""
        if self._embed_identity:
            return condition.reshape(condition.shape[0], -1)
        return self.net._embedding_net(condition).reshape(condition.shape[0], -1)

    def _gpart(self, n_part: int) -> Tensor:
        buf = self._cache.get("gpart")
        n = self.layout.n_params
        if buf is None or buf.shape[0] < n_part or buf.device != self.net.flat.device:
            buf = torch.zeros(max(n_part, 1), n, dtype=torch.float32, device=self.net.flat.device)
            self._cache["gpart"] = buf
        return buf

    # ---- shape handling: base.py:84-198 --------------------------------------------------------
    def _check_condition_shape(self, condition: Tensor):
        exp = self.condition_shape
        if len(condition.shape) < len(exp):
            raise ValueError(
                "Dimensionality of condition is too small and does not match the "
                f"expected dimensionality {len(exp)}. It should "
                f"be compatible with condition_shape {exp}.")
        if condition.shape[-len(exp):] != exp:
            raise ValueError(
                f"Shape of condition {condition.shape[-len(exp):]} does not match the "
                f"expected input dimensionality {exp}, as "
                "provided by condition_shape. Please reshape it accordingly.")

    def _check_input_shape(self, input: Tensor):
        exp = self.input_shape
        if len(input.shape) < len(exp):
            raise ValueError(
                "Dimensionality of input is too small and does not match the "
                f"expected dimensionality {len(exp)}. It should "
                f"be compatible with the provided input_shape {exp}.")
        if input.shape[-len(exp):] != exp:
            raise ValueError(
                f"Shape of input {input.shape[-len(exp):]} does not match the "
                f"expected input dimensionality {exp}, as "
                "provided by input_shape. Please reshape it accordingly.")

    def _align(self, input: Tensor, condition: Tensor):
        """_broadcast_and_align (base.py:142-198) without materialising a broadcast
        condition: returns input (S*B, D), condition rows and a `shared` flag."""
        in_ev, c_ev = len(self.input_shape), len(self.condition_shape)
        if input.dim() <= in_ev + 1:
            input = input.unsqueeze(0)
        S, Bi = input.shape[0], input.shape[1]
        cond_has_sample = condition.dim() > c_ev + 1
        Bc = condition.shape[1] if cond_has_sample else condition.shape[0]
        try:
            B = torch.broadcast_shapes((Bi,), (Bc,))[0]
        except RuntimeError as err:
            raise RuntimeError(
                "Expected `input` and `condition` to have broadcastable batch "
                "dimensions: their batch sizes must match, or one of them must be 1. "
                f"Got input={Bi} and condition={Bc}.") from err
        input = input.expand(S, B, *self.input_shape).reshape(S * B, -1)
        if not cond_has_sample and Bc == 1:
            return input, condition.reshape(1, *self.condition_shape), True, S, B
        if cond_has_sample:
            condition = condition.expand(S, B, *self.condition_shape)
        else:
            condition = condition.expand(B, *self.condition_shape).unsqueeze(0).expand(
                S, B, *self.condition_shape)
        return input, condition.reshape(S * B, *self.condition_shape), False, S, B

    # ---- the reference API ------------------------------------------------------------------------
    def log_prob(self, input: Tensor, condition: Tensor) -> Tensor:
        """(sample_dim, batch_dim) log-probabilities; nflows_flow.py:77-97."""
        self._check_input_shape(input)
        self._check_condition_shape(condition)
        if not self.net.flat.is_cuda:
            # the reference probes a freshly built (CPU-resident) net with two CPU rows before it
            # moves it to the training device (user_input_checks.py:767-795): device hop, no CPU math
            from ._refabc import hop_to_device
            return hop_to_device(self, "log_prob", input, condition)
        inp, cond, shared, S, B = self._align(input, condition)
        ctx = self._embed(cond)
        lp = _NsfLogProb.apply(self.net.flat, inp.contiguous().float(), ctx.contiguous().float(),
                               self, shared)
        return lp.reshape(S, B)

    def loss(self, input: Tensor, condition: Tensor) -> Tensor:
        """(batch_dim,) negative log-probabilities; nflows_flow.py:99-109."""
        return -self.log_prob(input.unsqueeze(0), condition)[0]

    def inverse_transform(self, input: Tensor, condition: Tensor) -> Tensor:
        """Base-space noise of the inputs; nflows_flow.py:42-75.  Like the reference (:73), the
        RAW condition feeds the transform here: neither the condition z-scoring nor the
        embedding net is applied on this code path."""
        self._check_condition_shape(condition)
        cdims = len(self.condition_shape)
        bshape = torch.broadcast_shapes(input.shape[:-1], condition.shape[:-cdims])
        inp = input.expand(bshape + (input.shape[-1],)).reshape(-1, input.shape[-1])
        cond = condition.expand(bshape + self.condition_shape).reshape(-1, *self.condition_shape)
        ctx = cond.reshape(cond.shape[0], -1)
        if ctx.shape[1] != self.layout.C:
            raise ValueError("inverse_transform: the raw condition does not have the embedded "
                             "context size (reference behaviour: no embedding on this path)")
        _, noise = self._logprob_raw(inp.contiguous().float(), ctx.contiguous().float(), False,
                                     want_noise=True, raw_condition=True)
        return noise.reshape(bshape + (noise.shape[-1],))

    @torch.no_grad()
    def sample(self, sample_shape, condition: Tensor) -> Tensor:
        """(*sample_shape, batch_dim, *input_shape); nflows_flow.py:111-128.  Noise is drawn
        with torch.randn on the parameter device in the order nflows draws it
        (B*n rows, condition-major), then pushed through the inverse-flow kernel."""
        self._check_condition_shape(condition)
        Bc = condition.shape[0]
        n = torch.Size(sample_shape).numel()
        D = self.layout.D
        noise = torch.randn(Bc * n, D, device=self.net.flat.device)
        x, _ = self.inverse_flow(noise, condition, n)
        x = x.reshape(Bc, n, D).transpose(0, 1)
        return x.reshape((*sample_shape, Bc, *self.input_shape))

    @torch.no_grad()
    def sample_and_log_prob(self, sample_shape, condition: Tensor, **kwargs):
        """nflows_flow.py:130-151 (via nflows Flow.sample_and_log_prob)."""
        Bc = condition.shape[0]
        n = torch.Size(sample_shape).numel()
        D = self.layout.D
        noise = torch.randn(Bc * n, D, device=self.net.flat.device)
        base_lp = -0.5 * (noise ** 2).sum(1) - 0.5 * D * 1.8378770664093453
        x, lad = self.inverse_flow(noise, condition, n)
        samples = x.reshape(Bc, n, D).reshape((*sample_shape, Bc, -1))
        log_probs = (base_lp - lad).reshape(Bc, n).reshape((*sample_shape, -1))
        return samples, log_probs

    @torch.no_grad()
    def inverse_flow(self, noise: Tensor, condition: Tensor, reps: int = 1):
        """x = T^{-1}(noise | condition); `noise` (B*reps, D) condition-major,
        `condition` (B, *condition_shape).  Returns (x, log|det dx/dnoise|)."""
        lib = L.load()
        L.require_cuda(noise, "noise")
        Bc = condition.shape[0]
        ctx = self._embed(condition.to(noise.device)).contiguous().float()
        shared = Bc == 1
        if not shared:
            ctx = ctx.repeat_interleave(reps, dim=0).contiguous()
        noise = noise.contiguous().float()
        R = noise.shape[0]
        out = torch.empty_like(noise)
        lad = torch.empty(R, dtype=torch.float32, device=noise.device)
        m = self._model(nbuf=2)
        rows = L.Rows(noise.data_ptr(), ctx.data_ptr(), None, R, 1 if shared else 0)
        force = os.environ.get("SBI_B200_TC", "") == "1"
        if self.fam.name == "nsf" and (R >= self.TC_MIN_ROWS or force):
            tc = self._tc_state(m)
            if tc is not None:
                L.check(L.load().sbi_b200_nsf_inverse_tc(C.byref(m), C.byref(tc), C.byref(rows), L.ptr(out),
                                                         L.ptr(lad), L.stream_ptr()), "nsf_inverse_tc")
                return out, lad
        L.check(self.fam.fn("inverse")(C.byref(m), C.byref(rows), L.ptr(out), L.ptr(lad),
                                       L.stream_ptr()), f"{self.fam.name}_inverse")
        return out, lad

    # ---- tensor-core bulk path (nsf only) --------------------------------------------------------
    #: rows from which log_prob / sampling go through the tcgen05 kernel (csrc/nsf_tc.cu).
    #: Measured crossover (profiles/tc_cross.py): one 128-row tile takes ~90 us end to end
    #: including the operand re-pack, the SIMT kernel 116 us at 2048 rows and 250 us at 10 000.
    #: SBI_B200_TC=0 disables, =1 forces.
    TC_MIN_ROWS = int(os.environ.get("SBI_B200_TC_MIN_ROWS", 1024))

    def _tc_state(self, m):
        """NsfTc struct for the current parameters, or None if the model is outside what the
        tensor-core kernel instantiates.  The operands are re-packed from the flat parameter
        buffer on every call (a ~1 MB elementwise kernel): the optimizer kernels update the
        parameters through raw pointers, so no version counter could be trusted."""
        if self.fam.name != "nsf" or os.environ.get("SBI_B200_TC", "") == "0":
            return None
        flat = self.net.flat
        st = self._cache.get("tc")
        if st is None or st["dev"] != flat.device:
            plan = self.layout.tc_plan()
            if plan is None:
                self._cache["tc"] = {"dev": flat.device, "plan": None}
                return None
            dev = flat.device
            st = {"dev": dev, "plan": plan,
                  "src": torch.as_tensor(plan["src"], device=dev),
                  "tab": torch.as_tensor(plan["tab"], device=dev),
                  "tcw": torch.empty(plan["n_words"], dtype=torch.float32, device=dev)}
            self._cache["tc"] = st
        if st["plan"] is None:
            return None
        tc = L.NsfTc(st["plan"]["n_words"], st["plan"]["stage_cap"], st["src"].data_ptr(),
                     st["tab"].data_ptr(), st["tcw"].data_ptr())
        lib = L.load()
        if not lib.sbi_b200_nsf_tc_supported(C.byref(m), C.byref(tc)):
            return None
        L.check(lib.sbi_b200_nsf_tc_pack(C.byref(m), C.byref(tc), L.stream_ptr()), "nsf_tc_pack")
        return tc

    # ---- fused forward+backward of a batch (parameter / input / condition gradients) --------------
    #: rows from which the training VJP runs on the tensor cores (csrc/nsf_vjp_tc.cu) when only parameter
    #: gradients are wanted; SBI_B200_VJP_TC=0 disables, =1 forces
    VJP_TC_MIN_ROWS = int(os.environ.get("SBI_B200_VJP_TC_MIN_ROWS", 256))

    def _tc_train_state(self, m, pack: bool = True):
        """(tc_fwd, tc_bwd, tc_both) operand descriptors for the tensor-core training step, freshly
        packed from the current parameters by ONE pack launch over both plans (`tc_both`), or None."""
        if self.fam.name != "nsf" or os.environ.get("SBI_B200_VJP_TC", "") == "0":
            return None
        flat = self.net.flat
        st = self._cache.get("tc_train")
        if st is None or st["dev"] != flat.device:
            pf, pb = self.layout.tc_plan(), (self.layout.tc_bwd_plan() if hasattr(self.layout, "tc_bwd_plan") else None)
            st = {"dev": flat.device, "ok": pf is not None and pb is not None}
            if st["ok"]:
                import numpy as np
                dev = flat.device
                st.update(nf=pf["n_words"], nb=pb["n_words"], cap_f=pf["stage_cap"], cap_b=pb["stage_cap"],
                          src=torch.as_tensor(np.concatenate([pf["src"], pb["src"]]), device=dev),
                          tab_f=torch.as_tensor(pf["tab"], device=dev), tab_b=torch.as_tensor(pb["tab"], device=dev),
                          tcw=torch.empty(pf["n_words"] + pb["n_words"], dtype=torch.float32, device=dev))
            self._cache["tc_train"] = st
        if not st["ok"]:
            return None
        lib = L.load()
        both = L.NsfTc(st["nf"] + st["nb"], st["cap_f"], st["src"].data_ptr(), st["tab_f"].data_ptr(),
                       st["tcw"].data_ptr())
        tcf = L.NsfTc(st["nf"], st["cap_f"], st["src"].data_ptr(), st["tab_f"].data_ptr(), st["tcw"].data_ptr())
        tcb = L.NsfTc(st["nb"], st["cap_b"], st["src"].data_ptr() + 4 * st["nf"], st["tab_b"].data_ptr(),
                      st["tcw"].data_ptr() + 4 * st["nf"])
        if not lib.sbi_b200_nsf_vjp_tc_supported(C.byref(m), C.byref(tcf), C.byref(tcb)):
            st["ok"] = False
            return None
        if pack:
            L.check(lib.sbi_b200_nsf_tc_pack(C.byref(m), C.byref(both), L.stream_ptr()), "nsf_tc_pack")
        return tcf, tcb, both

    def vjp_parts(self, R: int, param_grads_only: bool = True) -> int:
        """Number of partial-gradient slabs `vjp` writes for R rows."""
        if self._vjp_uses_tc(R, param_grads_only):
            return L.load().sbi_b200_nsf_vjp_tc_parts(R)
        return self.fam.fn("vjp_parts")(R)

    def _vjp_uses_tc(self, R: int, param_grads_only: bool) -> bool:
        env = os.environ.get("SBI_B200_VJP_TC", "")
        if self.fam.name != "nsf" or env == "0" or not param_grads_only:
            return False
        if R < self.VJP_TC_MIN_ROWS and env != "1":
            return False
        st = self._cache.get("tc_train")
        if st is not None and st["dev"] == self.net.flat.device:
            return bool(st["ok"])
        return self._tc_train_state(self._model(nbuf=3)) is not None

    def vjp(self, m, rows, R: int, gout, g_const: float, logp, gpart, ginput=None, gcond=None, loss_acc=None):
        """One launch (pair) of the fused forward+backward of `R` rows: partial parameter gradients of
        sum_r g_r log q_r into `gpart` ((vjp_parts(R, ...), n_params)), optionally the gradients
        w.r.t. the inputs / conditions, the log-probs and the loss statistics.  Tensor-core kernels when
        only parameter gradients are wanted (the trainer's case), else the SIMT kernel."""
        lib = L.load()
        if ginput is None and gcond is None and self._vjp_uses_tc(R, True):
            tcs = self._tc_train_state(m)
            if tcs is not None:
                nbytes = int(lib.sbi_b200_nsf_vjp_tc_save_bytes(C.byref(m), R))
                save = self._cache.get("vjp_save")
                if save is None or save.numel() * 4 < nbytes or save.device != self.net.flat.device:
                    save = torch.empty((nbytes + 3) // 4, dtype=torch.float32, device=self.net.flat.device)
                    self._cache["vjp_save"] = save
                L.check(lib.sbi_b200_nsf_vjp_tc(C.byref(m), C.byref(tcs[0]), C.byref(tcs[1]), C.byref(rows), L.ptr(gout),
                                                g_const, L.ptr(logp), L.ptr(gpart), L.ptr(loss_acc), L.ptr(save),
                                                save.numel() * 4, L.stream_ptr()), "nsf_vjp_tc")
                return
        L.check(self.fam.fn("vjp")(C.byref(m), C.byref(rows), L.ptr(gout), g_const, L.ptr(logp), L.ptr(gpart),
                                   L.ptr(ginput), L.ptr(gcond), L.ptr(loss_acc), L.stream_ptr()),
                f"{self.fam.name}_vjp")

    # ---- raw kernel entry (no autograd) --------------------------------------------------------------
    def _logprob_raw(self, inp: Tensor, ctx: Tensor, shared: bool, want_noise=False,
                     index: Optional[Tensor] = None, n_rows: Optional[int] = None,
                     raw_condition: bool = False):
        lib = L.load()
        L.require_cuda(inp, "input")
        L.require_cuda(ctx, "condition")
        R = inp.shape[0] if n_rows is None else n_rows
        lp = torch.empty(R, dtype=torch.float32, device=inp.device)
        noise = torch.empty(R, self.layout.D, dtype=torch.float32, device=inp.device) if want_noise else None
        m = self._model(nbuf=2, raw_condition=raw_condition)
        rows = L.Rows(inp.data_ptr(), ctx.data_ptr(),
                      None if index is None else index.data_ptr(), R, 1 if shared else 0)
        force = os.environ.get("SBI_B200_TC", "") == "1"
        if self.fam.name == "nsf" and (R >= self.TC_MIN_ROWS or force):
            tc = self._tc_state(m)
            if tc is not None:
                L.check(lib.sbi_b200_nsf_logprob_tc(C.byref(m), C.byref(tc), C.byref(rows), L.ptr(lp),
                                                    L.ptr(noise), L.stream_ptr()), "nsf_logprob_tc")
                return lp, noise
        L.check(self.fam.fn("logprob")(C.byref(m), C.byref(rows), L.ptr(lp), L.ptr(noise),
                                       L.stream_ptr()), f"{self.fam.name}_logprob")
        return lp, noise


class _NsfLogProb(torch.autograd.Function):
    @staticmethod
    def forward(ctx, flat, inp, cond, est, shared: bool):
        lp, _ = est._logprob_raw(inp, cond, shared)
        ctx.save_for_backward(inp, cond)
        ctx.est, ctx.shared = est, shared
        return lp

    @staticmethod
    def backward(ctx, g):
        inp, cond = ctx.saved_tensors
        est, shared = ctx.est, ctx.shared
        lib = L.load()
        need_flat, need_inp, need_cond = ctx.needs_input_grad[0], ctx.needs_input_grad[1], ctx.needs_input_grad[2]
        R = inp.shape[0]
        n_part = est.vjp_parts(R, not (need_inp or need_cond))
        gpart = est._gpart(n_part)
        ginp = torch.empty_like(inp) if need_inp else None
        gcond = torch.empty(R, cond.shape[1], dtype=torch.float32, device=inp.device) if need_cond else None
        m = est._model(nbuf=3)
        rows = L.Rows(inp.data_ptr(), cond.data_ptr(), None, R, 1 if shared else 0)
        g = g.contiguous().float()
        est.vjp(m, rows, R, g, 0.0, None, gpart, ginp, gcond, None)
        gflat = None
        if need_flat:
            gflat = torch.empty(est.layout.n_params, dtype=torch.float32, device=inp.device)
            L.check(lib.sbi_b200_reduce_partials(L.ptr(gpart), n_part, est.layout.n_params,
                                                 L.ptr(gflat), L.stream_ptr()), "reduce_partials")
        if need_cond and shared:
            gcond = gcond.sum(0, keepdim=True)
        return gflat, ginp, gcond, None, None


NSFEstimator = FlowEstimator
MAFEstimator = FlowEstimator


class MadeEstimator(FlowEstimator):
    r"""sbi's `made` density estimator (flow.py:37-112): z-scoring followed by a conditional MADE with a
    mixture-of-Gaussians head (nflows MADEMoG behind sbi's MADEMoGWrapper, nn_utils.py:133-201), evaluated by
    the NSF kernels with head = SBI_NSF_MOG.  The wrapper's dummy first feature is part of the network: the
    kernels see `input dim + 1` features, feature 0 is fed 0 by `log_prob` and -- like the reference's
    `_sample` -- drawn from its own mixture in `sample` and dropped from the result."""

    def _kernel_stats(self, raw_condition: bool = False):
        lay, net = self.layout, self.net
        emb = net._embedding_net
        std_mod = emb[0] if isinstance(emb, nn.Sequential) and isinstance(emb[0], Standardize) else None
        srcs = [net._shift, net._scale] + ([std_mod._mean, std_mod._std] if std_mod is not None else [])
        key = tuple((t.data_ptr(), t._version) for t in srcs) + (str(net.flat.device), raw_condition)
        ck = "stats_raw" if raw_condition else "stats"
        hit = self._cache.get(ck)
        if hit is not None and hit[0] == key:
            return hit[1], hit[2]
        dev = net.flat.device
        Din = lay.D - 1
        st = torch.zeros(2 * lay.Dp + 2 * lay.Cp, dtype=torch.float32, device=dev)
        st[lay.Dp:2 * lay.Dp] = 1.0
        st[2 * lay.Dp + lay.Cp:] = 1.0
        st[1:lay.D] = net._shift.expand(Din)                      # feature 0 (dummy): shift 0, scale 1
        st[lay.Dp + 1:lay.Dp + lay.D] = net._scale.expand(Din)
        if std_mod is not None and self._embed_identity and not raw_condition:
            st[2 * lay.Dp:2 * lay.Dp + lay.C] = std_mod._mean.reshape(-1).expand(lay.C)
            st[2 * lay.Dp + lay.Cp:2 * lay.Dp + lay.Cp + lay.C] = std_mod._std.reshape(-1).expand(lay.C)
        ld = float(torch.log(torch.abs(net._scale.double())).expand(Din).sum())
        self._cache[ck] = (key, st, ld)
        return st, ld

    @staticmethod
    def with_dummy(inp: Tensor) -> Tensor:
        """(R, D) -> (R, D + 1) with the wrapper's zero first feature (nn_utils.py:166-167)."""
        return torch.cat([torch.zeros(inp.shape[0], 1, dtype=inp.dtype, device=inp.device), inp], dim=1)

    def log_prob(self, input: Tensor, condition: Tensor) -> Tensor:
        self._check_input_shape(input)
        self._check_condition_shape(condition)
        if not self.net.flat.is_cuda:
            from ._refabc import hop_to_device
            return hop_to_device(self, "log_prob", input, condition)
        inp, cond, shared, S, B = self._align(input, condition)
        ctx = self._embed(cond)
        lp = _NsfLogProb.apply(self.net.flat, self.with_dummy(inp.float()).contiguous(), ctx.contiguous().float(),
                               self, shared)
        return lp.reshape(S, B)

    def inverse_transform(self, input: Tensor, condition: Tensor) -> Tensor:
        """The flow's transform is the z-scoring alone (CompositeTransform([standardize, identity]))."""
        return input * self.net._scale + self.net._shift

    @torch.no_grad()
    def sample(self, sample_shape, condition: Tensor) -> Tensor:
        """(*sample_shape, batch_dim, *input_shape): D + 1 sequential conditioner passes in one kernel
        (csrc/nsf.cu `made_sample_kernel`); the normal draws and the component-selecting uniforms come from
        torch on the parameter device, condition-major like nflows (`repeat_interleave(context, n)`)."""
        self._check_condition_shape(condition)
        lib = L.load()
        dev = self.net.flat.device
        Bc = condition.shape[0]
        n = torch.Size(sample_shape).numel()
        Dn = self.layout.D
        R = Bc * n
        noise = torch.randn(R, Dn, device=dev)
        unif = torch.rand(R, Dn, device=dev)
        ctx = self._embed(condition.to(dev)).contiguous().float()
        shared = Bc == 1
        if not shared:
            ctx = ctx.repeat_interleave(n, dim=0).contiguous()
        out = torch.empty(R, Dn, dtype=torch.float32, device=dev)
        m = self._model(nbuf=2)
        rows = L.Rows(noise.data_ptr(), ctx.data_ptr(), None, R, 1 if shared else 0)
        L.check(lib.sbi_b200_made_sample(C.byref(m), C.byref(rows), unif.data_ptr(), out.data_ptr(), L.stream_ptr()),
                "made_sample")
        x = out[:, 1:].reshape(Bc, n, Dn - 1).transpose(0, 1)
        return x.reshape((*sample_shape, Bc, *self.input_shape))

    @torch.no_grad()
    def sample_and_log_prob(self, sample_shape, condition: Tensor, **kwargs):
        samples = self.sample(sample_shape, condition)
        n = torch.Size(sample_shape).numel()
        flat = samples.reshape(n, condition.shape[0], -1)
        return samples, self.log_prob(flat, condition).reshape((*sample_shape, -1))

    def inverse_flow(self, noise: Tensor, condition: Tensor, reps: int = 1):
        raise NotImplementedError("`made` is a conditional distribution, not an invertible flow of base noise")
