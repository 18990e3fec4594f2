"""Packed parameter layout of the NSF kernels and its mapping to nflows' state_dict.

The kernels read ONE flat fp32 buffer.  Every matrix keeps PyTorch's native [out][in]
layout with the input dimension zero-padded to a multiple of 4 floats (16-byte rows for
cp.async.bulk and float4 shared-memory reads) and the output dimension padded to a multiple
of 4 with zero rows.  `NsfLayout` computes offsets, the per-layer descriptor table the
kernels index (include/sbi_b200.h, SBI_L_*), and, for every tensor of the reference
module (`net._transform._transforms.{i}...`, names as produced by the reference builder
/root/reference/sbi/neural_nets/net_builders/flow.py:333-460 on nflows 0.14), an index map
into the flat buffer, so a reference state_dict can be loaded / exported verbatim.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List

import numpy as np
import torch

from . import _lib as L


def round4(x: int) -> int:
    return (x + 3) & ~3


class _LayoutOps:
    """pack / unpack between reference-named tensors and the flat kernel buffer.  `weight_masks`
    holds, for masked (MADE) weights, the 0/1 mask: the flat buffer stores W*M, the masked-out raw
    values are kept aside (`raw`) only so that state_dict() round-trips exactly."""

    def _wm(self):
        return getattr(self, "_weight_masks", None) or {}

    def trainable_mask(self) -> torch.Tensor:
        m = torch.zeros(self.n_params, dtype=torch.uint8)
        for k, v in self.index.items():
            ix = torch.as_tensor(v.reshape(-1))
            if k in self._wm():
                m[ix] = torch.as_tensor(self._wm()[k].reshape(-1)).to(torch.uint8)
            else:
                m[ix] = 1
        return m

    def pack(self, state: Dict[str, torch.Tensor], out: torch.Tensor = None,
             raw_out: torch.Tensor = None) -> torch.Tensor:
        """Reference-named tensors -> flat buffer (on out's device if given)."""
        flat = torch.zeros(self.n_params, dtype=torch.float32) if out is None else out
        for k, ix in self.index.items():
            src = state[k].detach().to(dtype=torch.float32, device=flat.device).reshape(-1)
            pos = torch.as_tensor(ix.reshape(-1), device=flat.device)
            if k in self._wm():
                mk = torch.as_tensor(self._wm()[k].reshape(-1), dtype=torch.float32, device=flat.device)
                if raw_out is not None:
                    raw_out[pos] = src * (1 - mk)
                src = src * mk
            flat[pos] = src
        return flat

    def unpack(self, flat: torch.Tensor, raw: torch.Tensor = None) -> Dict[str, torch.Tensor]:
        out = {}
        for k, ix in self.index.items():
            pos = torch.as_tensor(ix.reshape(-1), device=flat.device)
            t = flat.detach()[pos]
            if raw is not None and k in self._wm():
                t = t + raw[pos]
            out[k] = t.reshape(ix.shape).clone()
        return out

    def num_real_params(self) -> int:
        return int(sum(v.size for v in self.index.values()))


@dataclass
class NsfLayout(_LayoutOps):
    D: int
    C: int
    H: int = 50
    NB: int = 2
    KB: int = 10
    T: int = 5
    tail_bound: float = 3.0
    zscore_input: bool = True
    zscore_cond: bool = True
    embed_is_identity: bool = True
    wcap_target: int = int(__import__('os').environ.get('SBI_B200_WCAP', 4096))

    # derived
    n_params: int = 0
    index: Dict[str, np.ndarray] = field(default_factory=dict, repr=False)

    def __post_init__(self):
        D, C, H, NB, KB, T = self.D, self.C, self.H, self.NB, self.KB, self.T
        if D < 2:
            raise NotImplementedError("NSF kernels need input dim >= 2")
        if NB > L.SBI_NSF_MAX_BLOCKS:
            raise ValueError(f"num_blocks <= {L.SBI_NSF_MAX_BLOCKS}")
        self.Dp, self.Cp, self.Hp = round4(D), round4(C), round4(H)
        self.NPAR = 3 * KB - 1
        self.PR = round4(self.NPAR)
        # alternating masks, reference: torchutils.py:396-410 / flow.py:396-397
        self.id_feats: List[np.ndarray] = []
        self.tr_feats: List[np.ndarray] = []
        for i in range(T):
            mask = np.zeros(D, np.int64)
            mask[(0 if i % 2 == 0 else 1)::2] = 1
            self.tr_feats.append(np.nonzero(mask > 0)[0])
            self.id_feats.append(np.nonzero(mask <= 0)[0])
        self.IDp = round4(max(len(f) for f in self.id_feats))
        self.TRmax = max(len(f) for f in self.tr_feats)
        self.K0p = self.Cp + self.IDp

        # weight-ring chunking
        Hp, Cp = self.Hp, self.Cp
        cap = max(self.wcap_target, 4 * self.K0p, 4 * (Hp + Cp), self.PR * Hp)

        def rows(rowlen):
            return max(4, min(Hp, (cap // rowlen) & ~3))

        self.rpc0, self.rpc1, self.rpc2 = rows(self.K0p), rows(Hp), rows(Hp + Cp)
        self.nf_chunk = max(1, min(self.TRmax, cap // (self.PR * Hp)))
        used = max(self.rpc0 * self.K0p, self.rpc1 * Hp, self.rpc2 * (Hp + Cp),
                   self.nf_chunk * self.PR * Hp)
        self.wcap = (used + 31) & ~31

        # ---- offsets -------------------------------------------------------------------
        off = 0

        def take(n):
            nonlocal off
            o = off
            off += round4(n)
            return o

        tab = np.zeros((T, L.SBI_NSF_LAYER_STRIDE), np.int32)
        feat = []
        ntri = D * (D - 1) // 2
        tri_lo = np.tril_indices(D, -1)
        tri_up = np.triu_indices(D, 1)
        base = 1 if self.zscore_input else 0
        idx: Dict[str, np.ndarray] = {}
        self.buffers: Dict[str, torch.Tensor] = {}
        for l in range(T):
            idf, trf = self.id_feats[l], self.tr_feats[l]
            n_id, n_tr = len(idf), len(trf)
            ci, li = base + 2 * l, base + 2 * l + 1
            pc = f"net._transform._transforms.{ci}."
            pl = f"net._transform._transforms.{li}."
            tab[l, L.L_NID], tab[l, L.L_NTR] = n_id, n_tr
            tab[l, L.L_FEAT] = len(feat)
            feat += list(idf) + list(trf)
            self.buffers[pc + "identity_features"] = torch.as_tensor(idf)
            self.buffers[pc + "transform_features"] = torch.as_tensor(trf)
            # initial layer: nflows columns [id | ctx] -> packed columns [ctx | pad | id | pad]
            o = take(Hp * self.K0p)
            tab[l, L.L_W0] = o
            cols = np.concatenate([Cp + np.arange(n_id), np.arange(C)])
            idx[pc + "transform_net.initial_layer.weight"] = (
                o + np.arange(H)[:, None] * self.K0p + cols[None, :])
            o = take(Hp)
            tab[l, L.L_B0] = o
            idx[pc + "transform_net.initial_layer.bias"] = o + np.arange(H)
            for b in range(NB):
                pb = pc + f"transform_net.blocks.{b}."
                t = L.L_BLK0 + 6 * b
                for slot, (name, K, Kp) in enumerate(
                        [("linear_layers.0", H, Hp), ("linear_layers.1", H, Hp),
                         ("context_layer", C, Cp)]):
                    o = take(Hp * Kp)
                    tab[l, t + 2 * slot] = o
                    idx[pb + name + ".weight"] = o + np.arange(H)[:, None] * Kp + np.arange(K)[None, :]
                    o = take(Hp)
                    tab[l, t + 2 * slot + 1] = o
                    idx[pb + name + ".bias"] = o + np.arange(H)
            # final layer: feature f owns packed rows f*PR .. f*PR+NPAR-1
            o = take(n_tr * self.PR * Hp)
            tab[l, L.L_WF] = o
            prow = (np.arange(n_tr)[:, None] * self.PR + np.arange(self.NPAR)[None, :]).reshape(-1)
            idx[pc + "transform_net.final_layer.weight"] = o + prow[:, None] * Hp + np.arange(H)[None, :]
            o = take(n_tr * self.PR)
            tab[l, L.L_BF] = o
            idx[pc + "transform_net.final_layer.bias"] = o + prow
            # LULinear
            tab[l, L.L_HAS_LU] = 1
            o = take(ntri)
            tab[l, L.L_LU_LOWER] = o
            idx[pl + "lower_entries"] = o + np.arange(ntri)
            o = take(ntri)
            tab[l, L.L_LU_UPPER] = o
            idx[pl + "upper_entries"] = o + np.arange(ntri)
            o = take(D)
            tab[l, L.L_LU_DIAG] = o
            idx[pl + "unconstrained_upper_diag"] = o + np.arange(D)
            o = take(D)
            tab[l, L.L_LU_BIAS] = o
            idx[pl + "bias"] = o + np.arange(D)
        del tri_lo, tri_up
        self.n_params = off
        self.index = idx
        self.layer_tab = tab
        self.feat_tab = np.asarray(feat, np.int32)
        self.edge_raw = float(np.log(np.exp(1 - 1e-3) - 1))

    # ------------------------------------------------------------------------------ helpers
    def tables(self):
        return self.layer_tab.reshape(-1).astype(np.int32), self.feat_tab.astype(np.int32)

    # ------------------------------------------------------------- tensor-core operand plan
    def tc_plan(self):
        """Gather map + stage table for the tcgen05 evaluation path (include/sbi_b200.h,
        `sbi_nsf_tc`; kernel sbi_b200/csrc/nsf_tc.cu), or None when the model is outside what
        that kernel instantiates.

        Every linear of the conditioner (nflows ResidualNet) becomes one or two K-major
        no-swizzle UMMA operand blocks [K/4 slabs][N rows][4 floats].  The hidden operand's
        columns are [hidden (H) | context (C) | 0] so that the context never needs its own
        staging: the GLU gate reads the K-steps that cover columns H..H+C-1.
        Returns dict(src=int32 (n_words,), tab=int32 (T*STRIDE,), stage_cap=int, n_words=int).
        """
        D, C, H, NB, T = self.D, self.C, self.H, self.NB, self.T
        if H != 50 or self.KB != 10 or H + C > 64 or self.IDp > 48 or self.PR > 32 or D > 16:
            return None
        Hp, Cp, K0p, PR, NPAR = self.Hp, self.Cp, self.K0p, self.PR, self.NPAR
        HP8 = (H + 7) & ~7
        KC0 = H // 8
        nkc = (H + C + 7) // 8 - KC0
        tab = np.zeros((T, L.SBI_NSF_TC_STRIDE), np.int32)
        chunks = []          # per stage: int64 array of the hi half (param index or -1)
        off = 0
        stage_cap = 0

        def block(N, K, fill):
            """fill(n, k) -> param index or -1 ; returns the flattened [K/4][N][4] block"""
            blk = np.full((K // 4, N, 4), -1, np.int64)
            n = np.arange(N)[:, None]
            k = np.arange(K)[None, :]
            vals = fill(n + 0 * k, k + 0 * n)
            blk[(k // 4) + 0 * n, n + 0 * k, (k % 4) + 0 * n] = vals
            return blk.reshape(-1)

        def ctx_block(woff, rowlen):
            def fill(n, k):
                c = 8 * KC0 + k - H
                ok = (n < H) & (c >= 0) & (c < C)
                return np.where(ok, woff + n * rowlen + np.clip(c, 0, max(C - 1, 0)), -1)
            return block(64, 8 * nkc, fill)

        def hidden_block(woff, N, rowmap):
            """rowmap(n) -> (valid, packed row index) of the [.,Hp] weight matrix"""
            def fill(n, k):
                valid, row = rowmap(n)
                ok = valid & (k < H)
                return np.where(ok, woff + row * Hp + np.minimum(k, H - 1), -1)
            return block(N, HP8, fill)

        for l in range(T):
            lt = self.layer_tab[l]
            n_id, n_tr = int(lt[L.L_NID]), int(lt[L.L_NTR])
            kid8 = (n_id + 7) & ~7
            stages = []      # (hi half, N, aux)
            w0 = int(lt[L.L_W0])

            def fill_id(n, k, w0=w0, n_id=n_id):
                ok = (n < H) & (k < n_id)
                return np.where(ok, w0 + n * K0p + Cp + np.minimum(k, max(n_id - 1, 0)), -1)

            stages.append((np.concatenate([block(64, kid8, fill_id), ctx_block(w0, K0p)]), 64, 0))
            ident = lambda n: (n < H, np.minimum(n, H - 1))
            for b in range(NB):
                t = L.L_BLK0 + 6 * b
                w1, w2, wc = int(lt[t + 0]), int(lt[t + 2]), int(lt[t + 4])
                stages.append((ctx_block(wc, Cp), 64, 0))
                stages.append((hidden_block(w1, 64, ident), 64, 0))
                stages.append((hidden_block(w2, 64, ident), 64, 0))
            wf = int(lt[L.L_WF])
            f0 = 0
            while f0 < n_tr:
                nf = min(2, n_tr - f0)

                def rowmap(n, f0=f0, nf=nf):
                    f, i = n // 32, n % 32
                    valid = (f < nf) & (i < NPAR)
                    return valid, np.where(valid, (f0 + f) * PR + i, 0)

                stages.append((hidden_block(wf, 32 * nf, rowmap), 32 * nf, f0 | (nf << 16)))
                f0 += nf
            if len(stages) > L.SBI_NSF_TC_MAX_STAGES:
                return None
            tab[l, 0], tab[l, 1] = len(stages), kid8
            for s, (hi, N, aux) in enumerate(stages):
                nfl = 2 * hi.size
                tab[l, 4 + 4 * s: 8 + 4 * s] = (off, nfl, N, aux)
                chunks.append(hi)
                chunks.append(np.where(hi >= 0, -2 - hi, -1))
                off += nfl
                stage_cap = max(stage_cap, nfl)
        src = np.concatenate(chunks).astype(np.int32)
        assert src.size == off
        return dict(src=src, tab=tab.reshape(-1).astype(np.int32),
                    stage_cap=int((stage_cap + 31) & ~31), n_words=int(off))

    def tc_bwd_plan(self):
        """Gather map + stage table of the TRANSPOSED linears for the tcgen05 training kernel's
        input-gradient chain (kernel sbi_b200/csrc/nsf_vjp_tc.cu): dX = dY W needs, as the B operand
        [N = in-features][K = out-features] in the same K-major no-swizzle layout, B[n][k] = W[k][n].
        Stages of a layer in the order the backward sweep uses them (aux = number of K-steps):
            final layer, one pass per <= 2 spline features:  N = 64 (hidden), K = 32 * nf
            per block b = NB-1 .. 0:  W2^T (N = 64, K = 56),  W1^T (N = 64, K = 56)
            initial layer, identity-feature columns only:  N = 16, K = 56
        Same return format as tc_plan (the context columns are not needed: training never asks for
        the condition's gradient on this path)."""
        D, C, H, NB, T = self.D, self.C, self.H, self.NB, self.T
        if self.tc_plan() is None or self.IDp > 16:
            return None
        Hp, Cp, K0p, PR, NPAR = self.Hp, self.Cp, self.K0p, self.PR, self.NPAR
        HP8 = (H + 7) & ~7
        tab = np.zeros((T, L.SBI_NSF_TC_STRIDE), np.int32)
        chunks, off, stage_cap = [], 0, 0

        def block(N, K, fill):
            blk = np.full((K // 4, N, 4), -1, np.int64)
            n = np.arange(N)[:, None]
            k = np.arange(K)[None, :]
            vals = fill(n + 0 * k, k + 0 * n)
            blk[(k // 4) + 0 * n, n + 0 * k, (k % 4) + 0 * n] = vals
            return blk.reshape(-1)

        for l in range(T):
            lt = self.layer_tab[l]
            n_id, n_tr = int(lt[L.L_NID]), int(lt[L.L_NTR])
            stages = []
            wf = int(lt[L.L_WF])
            f0 = 0
            while f0 < n_tr:
                nf = min(2, n_tr - f0)

                def fill_f(n, k, f0=f0, nf=nf):
                    f, i = k // 32, k % 32
                    ok = (n < H) & (f < nf) & (i < NPAR)
                    return np.where(ok, wf + ((f0 + np.minimum(f, nf - 1)) * PR + np.minimum(i, NPAR - 1)) * Hp
                                    + np.minimum(n, H - 1), -1)

                stages.append((block(64, 32 * nf, fill_f), 64, 4 * nf))
                f0 += nf
            for b in range(NB - 1, -1, -1):
                t = L.L_BLK0 + 6 * b
                for w in (int(lt[t + 2]), int(lt[t + 0])):          # W2 then W1

                    def fill_h(n, k, w=w):
                        ok = (n < H) & (k < H)
                        return np.where(ok, w + np.minimum(k, H - 1) * Hp + np.minimum(n, H - 1), -1)

                    stages.append((block(64, HP8, fill_h), 64, HP8 // 8))
            w0 = int(lt[L.L_W0])

            def fill_0(n, k, w0=w0, n_id=n_id):
                ok = (n < n_id) & (k < H)
                return np.where(ok, w0 + np.minimum(k, H - 1) * K0p + Cp + np.minimum(n, max(n_id - 1, 0)), -1)

            stages.append((block(16, HP8, fill_0), 16, HP8 // 8))
            tab[l, 0], tab[l, 1] = len(stages), (n_tr + 1) // 2
            for s, (hi, N, aux) in enumerate(stages):
                nfl = 2 * hi.size
                tab[l, 4 + 4 * s: 8 + 4 * s] = (off, nfl, N, aux)
                chunks.append(hi)
                chunks.append(np.where(hi >= 0, -2 - hi, -1))
                off += nfl
                stage_cap = max(stage_cap, nfl)
        src = np.concatenate(chunks).astype(np.int32)
        return dict(src=src, tab=tab.reshape(-1).astype(np.int32),
                    stage_cap=int((stage_cap + 31) & ~31), n_words=int(off))

    def fill_struct(self, s: "L.NsfModel", nbuf: int):
        s.D, s.C, s.H, s.NB, s.KB, s.T = self.D, self.C, self.H, self.NB, self.KB, self.T
        s.Dp, s.Cp, s.IDp, s.Hp, s.PR = self.Dp, self.Cp, self.IDp, self.Hp, self.PR
        s.TRmax, s.nf_chunk = self.TRmax, self.nf_chunk
        s.rpc0, s.rpc1, s.rpc2 = self.rpc0, self.rpc1, self.rpc2
        s.wcap, s.nbuf, s.n_params = self.wcap, nbuf, self.n_params
        s.tail_bound = self.tail_bound
        s.inv_sqrt_h = 1.0 / math.sqrt(self.H)
        s.min_bw = s.min_bh = s.min_d = 1e-3
        s.edge_raw = self.edge_raw
        s.head, s.M, s.mog_eps, s.cond_mlp = 0, 0, 0.0, 0
        return s


NsfLayout.family = "nsf"


@dataclass
class Nsf1dLayout(_LayoutOps):
    """Packed layout of the ONE-dimensional neural spline flow (flow.py:401-432 with x_numel == 1): T spline
    transforms of the single feature whose 3K-1 parameters come from a context-only MLP (`ContextSplineMap`,
    flow.py:1419-1478: Linear -> ReLU -> hidden_layers x [one shared Linear -> ReLU] -> Linear), no LULinear.
    Runs on the NSF kernels (`sbi_nsf_model` with cond_mlp = 1, NB = hidden_layers, no identity features)."""
    C: int
    H: int = 50
    NB: int = 1                  # hidden_layers_spline_context
    KB: int = 10
    T: int = 5
    tail_bound: float = 3.0
    zscore_input: bool = True
    zscore_cond: bool = True
    embed_is_identity: bool = True
    wcap_target: int = 4096
    n_params: int = 0
    index: Dict[str, np.ndarray] = field(default_factory=dict, repr=False)
    D: int = 1

    def __post_init__(self):
        C, H, NB, KB, T = self.C, self.H, self.NB, self.KB, self.T
        if NB < 0 or NB > L.SBI_NSF_MAX_BLOCKS:
            raise ValueError(f"0 <= hidden_layers_spline_context <= {L.SBI_NSF_MAX_BLOCKS}")
        self.Dp, self.Cp, self.Hp = 4, round4(C), round4(H)
        self.NPAR = 3 * KB - 1
        self.PR = round4(self.NPAR)
        if self.PR > self.Hp:
            raise ValueError("hidden_features must be >= the padded spline parameter count")
        self.IDp, self.TRmax, self.K0p = 0, 1, self.Cp
        Hp, Cp = self.Hp, self.Cp
        cap = max(self.wcap_target, 4 * self.K0p, 4 * (Hp + Cp), self.PR * Hp)

        def rows(rowlen):
            return max(4, min(Hp, (cap // rowlen) & ~3))

        self.rpc0, self.rpc1, self.rpc2 = rows(self.K0p), rows(Hp), rows(Hp + Cp)
        self.nf_chunk = 1
        used = max(self.rpc0 * self.K0p, self.rpc1 * Hp, self.rpc2 * (Hp + Cp), self.PR * Hp)
        self.wcap = (used + 31) & ~31
        off = 0

        def take(n):
            nonlocal off
            o = off
            off += round4(n)
            return o

        tab = np.zeros((T, L.SBI_NSF_LAYER_STRIDE), np.int32)
        idx: Dict[str, np.ndarray] = {}
        self.buffers: Dict[str, torch.Tensor] = {}
        base = 1 if self.zscore_input else 0
        for l in range(T):
            pc = f"net._transform._transforms.{base + l}."
            pn = pc + "transform_net.spline_predictor."
            tab[l, L.L_NID], tab[l, L.L_NTR], tab[l, L.L_FEAT] = 0, 1, l
            self.buffers[pc + "identity_features"] = torch.zeros(0, dtype=torch.int64)
            self.buffers[pc + "transform_features"] = torch.zeros(1, dtype=torch.int64)
            o = take(Hp * Cp)
            tab[l, L.L_W0] = o
            idx[pn + "0.weight"] = o + np.arange(H)[:, None] * Cp + np.arange(C)[None, :]
            o = take(Hp)
            tab[l, L.L_B0] = o
            idx[pn + "0.bias"] = o + np.arange(H)
            ow, ob = take(Hp * Hp), take(Hp)
            tab[l, L.L_BLK0], tab[l, L.L_BLK0 + 1] = ow, ob
            for k in range(NB):      # nn.Sequential lists the shared module once per position
                idx[pn + f"{2 + 2 * k}.weight"] = ow + np.arange(H)[:, None] * Hp + np.arange(H)[None, :]
                idx[pn + f"{2 + 2 * k}.bias"] = ob + np.arange(H)
            o = take(self.PR * Hp)
            tab[l, L.L_WF] = o
            idx[pn + f"{2 + 2 * NB}.weight"] = o + np.arange(self.NPAR)[:, None] * Hp + np.arange(H)[None, :]
            o = take(self.PR)
            tab[l, L.L_BF] = o
            idx[pn + f"{2 + 2 * NB}.bias"] = o + np.arange(self.NPAR)
            tab[l, L.L_HAS_LU] = 0
        self.n_params = off
        self.index = idx
        self.layer_tab = tab
        self.feat_tab = np.zeros(T, np.int32)            # layer l: no identity features, transformed feature 0
        self.edge_raw = float(np.log(np.exp(1 - 1e-3) - 1))

    def tables(self):
        return self.layer_tab.reshape(-1).astype(np.int32), self.feat_tab.astype(np.int32)

    def tc_plan(self):
        return None

    def num_real_params(self) -> int:
        return int(len({int(i) for v in self.index.values() for i in v.reshape(-1)}))

    def fill_struct(self, s: "L.NsfModel", nbuf: int):
        s.D, s.C, s.H, s.NB, s.KB, s.T = 1, self.C, self.H, self.NB, self.KB, self.T
        s.Dp, s.Cp, s.IDp, s.Hp, s.PR = self.Dp, self.Cp, 0, self.Hp, self.PR
        s.TRmax, s.nf_chunk = 1, 1
        s.rpc0, s.rpc1, s.rpc2 = self.rpc0, self.rpc1, self.rpc2
        s.wcap, s.nbuf, s.n_params = self.wcap, nbuf, self.n_params
        s.tail_bound = self.tail_bound
        s.inv_sqrt_h = 1.0 / math.sqrt(self.H)
        s.min_bw = s.min_bh = s.min_d = 1e-3
        s.edge_raw = self.edge_raw
        s.head, s.M, s.mog_eps, s.cond_mlp = 0, 0, 0.0, 1
        return s


Nsf1dLayout.family = "nsf"


@dataclass
class MadeLayout(_LayoutOps):
    """Packed layout of sbi's `made` density estimator (flow.py:37-112): ONE masked residual network
    (nflows MixtureOfGaussiansMADE behind sbi's MADEMoGWrapper, nn_utils.py:133-201: `features + 1`
    inputs with a dummy first feature) emitting, per feature, num_mixture_components x (logit, mean,
    unconstrained std).  The network is structurally the NSF conditioner (initial linear on
    [context | inputs], residual blocks with GLU context gates, final linear), so it runs on the NSF
    kernels (include/sbi_b200.h `sbi_nsf_model` with head = SBI_NSF_MOG): masks are folded into the
    packed weights, every feature is both a conditioner input and an output, T = 1, no LU.
    `D` here is the NETWORK's feature count (the estimator's input dim + 1)."""
    D: int
    C: int
    H: int = 50
    NB: int = 5
    M: int = 10
    epsilon: float = 1e-2
    zscore_input: bool = True
    zscore_cond: bool = True
    embed_is_identity: bool = True
    wcap_target: int = 4096
    n_params: int = 0
    index: Dict[str, np.ndarray] = field(default_factory=dict, repr=False)

    def __post_init__(self):
        D, C, H, NB, M = self.D, self.C, self.H, self.NB, self.M
        if NB > L.SBI_NSF_MAX_BLOCKS:
            raise ValueError(f"num_blocks <= {L.SBI_NSF_MAX_BLOCKS}")
        if not (1 <= M <= 16):
            raise ValueError("the mixture code keeps <= 16 components in registers")
        self.T, self.KB = 1, 0
        self.Dp, self.Cp, self.Hp = round4(D), round4(C), round4(H)
        self.NPAR = 3 * M
        self.PR = round4(self.NPAR)
        self.IDp = round4(D)
        self.TRmax = D
        self.K0p = self.Cp + self.IDp
        Hp, Cp = self.Hp, self.Cp
        cap = max(self.wcap_target, 4 * self.K0p, 4 * (Hp + Cp), self.PR * Hp)

        def rows(rowlen):
            return max(4, min(Hp, (cap // rowlen) & ~3))

        self.rpc0, self.rpc1, self.rpc2 = rows(self.K0p), rows(Hp), rows(Hp + Cp)
        self.nf_chunk = max(1, min(self.TRmax, cap // (self.PR * Hp)))
        used = max(self.rpc0 * self.K0p, self.rpc1 * Hp, self.rpc2 * (Hp + Cp), self.nf_chunk * self.PR * Hp)
        self.wcap = (used + 31) & ~31

        # MADE degrees / masks (nflows transforms/made.py, sequential degrees; residual blocks)
        in_deg = np.arange(1, D + 1)
        max_, min_ = max(1, D - 1), min(1, D - 1)
        hid_deg = np.arange(H) % max_ + min_
        m_init = (hid_deg[:, None] >= in_deg[None, :]).astype(np.float32)      # (H, D)
        m_hid = (hid_deg[:, None] >= hid_deg[None, :]).astype(np.float32)      # (H, H)
        out_deg = np.repeat(in_deg, 3 * M)
        m_out = (out_deg[:, None] > hid_deg[None, :]).astype(np.float32)       # (3M*D, H)

        off = 0

        def take(n):
            nonlocal off
            o = off
            off += round4(n)
            return o

        tab = np.zeros((1, L.SBI_NSF_LAYER_STRIDE), np.int32)
        idx: Dict[str, np.ndarray] = {}
        wm: Dict[str, np.ndarray] = {}
        self.buffers: Dict[str, torch.Tensor] = {}
        pm = "net._distribution._made."
        feats = list(range(D))
        tab[0, L.L_NID], tab[0, L.L_NTR], tab[0, L.L_FEAT] = D, D, 0
        # initial layer: packed columns [ctx | pad | inputs | pad]; MADE's context_layer supplies the ctx columns
        o = take(Hp * self.K0p)
        tab[0, L.L_W0] = o
        idx[pm + "initial_layer.weight"] = o + np.arange(H)[:, None] * self.K0p + (Cp + np.arange(D))[None, :]
        wm[pm + "initial_layer.weight"] = m_init
        self.buffers[pm + "initial_layer.mask"] = torch.as_tensor(m_init)
        self.buffers[pm + "initial_layer.degrees"] = torch.as_tensor(hid_deg)
        idx[pm + "context_layer.weight"] = o + np.arange(H)[:, None] * self.K0p + np.arange(C)[None, :]
        o = take(Hp)
        tab[0, L.L_B0] = o
        idx[pm + "initial_layer.bias"] = o + np.arange(H)
        o = take(Hp)
        tab[0, L.L_BC0] = o
        idx[pm + "context_layer.bias"] = o + np.arange(H)
        for b in range(NB):
            pb = pm + f"blocks.{b}."
            t = L.L_BLK0 + 6 * b
            for slot, (name, K, Kp, masked) in enumerate(
                    [("linear_layers.0", H, Hp, True), ("linear_layers.1", H, Hp, True),
                     ("context_layer", C, Cp, False)]):
                o = take(Hp * Kp)
                tab[0, t + 2 * slot] = o
                idx[pb + name + ".weight"] = o + np.arange(H)[:, None] * Kp + np.arange(K)[None, :]
                if masked:
                    wm[pb + name + ".weight"] = m_hid
                    self.buffers[pb + name + ".mask"] = torch.as_tensor(m_hid)
                    self.buffers[pb + name + ".degrees"] = torch.as_tensor(hid_deg)
                o = take(Hp)
                tab[0, t + 2 * slot + 1] = o
                idx[pb + name + ".bias"] = o + np.arange(H)
        # final layer: feature f owns packed rows f*PR .. f*PR + 3M - 1 (nflows row f*3M + 3m + k)
        o = take(D * self.PR * Hp)
        tab[0, L.L_WF] = o
        prow = (np.arange(D)[:, None] * self.PR + np.arange(self.NPAR)[None, :]).reshape(-1)
        idx[pm + "final_layer.weight"] = o + prow[:, None] * Hp + np.arange(H)[None, :]
        wm[pm + "final_layer.weight"] = m_out
        self.buffers[pm + "final_layer.mask"] = torch.as_tensor(m_out)
        self.buffers[pm + "final_layer.degrees"] = torch.as_tensor(out_deg)
        o = take(D * self.PR)
        tab[0, L.L_BF] = o
        idx[pm + "final_layer.bias"] = o + prow
        tab[0, L.L_HAS_LU] = 0
        self.n_params = off
        self.index = idx
        self._weight_masks = wm
        self.layer_tab = tab
        self.feat_tab = np.asarray(feats + feats, np.int32)

    def tables(self):
        return self.layer_tab.reshape(-1).astype(np.int32), self.feat_tab.astype(np.int32)

    def tc_plan(self):
        return None

    def fill_struct(self, s: "L.NsfModel", nbuf: int):
        s.D, s.C, s.H, s.NB, s.KB, s.T = self.D, self.C, self.H, self.NB, 2, 1
        s.Dp, s.Cp, s.IDp, s.Hp, s.PR = self.Dp, self.Cp, self.IDp, self.Hp, self.PR
        s.TRmax, s.nf_chunk = self.TRmax, self.nf_chunk
        s.rpc0, s.rpc1, s.rpc2 = self.rpc0, self.rpc1, self.rpc2
        s.wcap, s.nbuf, s.n_params = self.wcap, nbuf, self.n_params
        s.tail_bound, s.inv_sqrt_h, s.edge_raw = 1.0, 1.0, 0.0
        s.min_bw = s.min_bh = s.min_d = 1e-3
        s.head, s.M, s.mog_eps, s.cond_mlp = 1, self.M, self.epsilon, 0
        return s


MadeLayout.family = "made"


@dataclass
class MafLayout(_LayoutOps):
    """Packed layout of the MAF kernels (include/sbi_b200.h `sbi_maf_model`), mapping the tensors of
    the reference module built by /root/reference/sbi/neural_nets/net_builders/flow.py:115-209 on
    nflows 0.14 (MaskedAffineAutoregressiveTransform(MADE) + RandomPermutation per layer)."""
    D: int
    C: int
    H: int = 50
    NB: int = 2
    T: int = 5
    perms: List[np.ndarray] = None       # permutation of each layer (RandomPermutation buffer)
    zscore_input: bool = True
    zscore_cond: bool = True
    embed_is_identity: bool = True
    scale_softplus: bool = True          # softplus(s)+1e-3 (see oracle/nflows_port/transforms/autoregressive.py)
    wcap_target: int = 4096
    n_params: int = 0
    index: Dict[str, np.ndarray] = field(default_factory=dict, repr=False)
    # element-wise transform: "affine" (maf) or "rqs" (maf_rqs, flow.py:212-330; linear tails)
    head: str = "affine"
    KB: int = 10
    tail_bound: float = 3.0
    min_bin_width: float = 1e-3
    min_bin_height: float = 1e-3
    min_derivative: float = 1e-3

    def __post_init__(self):
        D, C, H, NB, T = self.D, self.C, self.H, self.NB, self.T
        if NB > 8:
            raise ValueError("num_blocks <= 8")
        if self.head not in ("affine", "rqs"):
            raise ValueError(self.head)
        if self.head == "rqs" and not (2 <= self.KB <= 16):
            raise ValueError("the spline code keeps <= 16 bins in registers")
        self.OUTM = 2 if self.head == "affine" else 3 * self.KB - 1
        self.Dp, self.Cp, self.Hp = round4(D), round4(C), round4(H)
        self.OUTp = round4(self.OUTM * D)
        Hp, Dp, Cp = self.Hp, self.Dp, self.Cp
        cap = max(self.wcap_target, 4 * (Dp + Cp), 4 * Hp)

        def rows(rowlen, nmax):
            return max(4, min(nmax, (cap // rowlen) & ~3))

        self.rpc0, self.rpc1, self.rpcf = rows(Dp + Cp, Hp), rows(Hp, Hp), rows(Hp, self.OUTp)
        used = max(self.rpc0 * (Dp + Cp), self.rpc1 * Hp, self.rpcf * Hp, 4 * Cp, 4 * Dp)
        self.wcap = (used + 31) & ~31

        # MADE degrees / masks (nflows transforms/made.py, sequential degrees)
        in_deg = np.arange(1, D + 1)
        max_, min_ = max(1, D - 1), min(1, D - 1)
        hid_deg = np.arange(H) % max_ + min_
        m_init = (hid_deg[:, None] >= in_deg[None, :]).astype(np.float32)      # (H, D)
        m_hid = (hid_deg[:, None] >= hid_deg[None, :]).astype(np.float32)      # (H, H)
        out_deg = np.repeat(in_deg, self.OUTM)                                 # tile(.., OUTM)
        m_out = (out_deg[:, None] > hid_deg[None, :]).astype(np.float32)       # (OUTM*D, H)
        self.degrees = dict(input=in_deg, hidden=hid_deg, output=out_deg)

        off = 0

        def take(n):
            nonlocal off
            o = off
            off += round4(n)
            return o

        if self.perms is None:
            self.perms = [np.arange(D) for _ in range(T)]
        tab = np.zeros((T, L.SBI_MAF_LAYER_STRIDE), np.int32)
        ptab = []
        idx: Dict[str, np.ndarray] = {}
        wm: Dict[str, np.ndarray] = {}
        self.buffers: Dict[str, torch.Tensor] = {}
        base = 1 if self.zscore_input else 0
        for l in range(T):
            pa = f"net._transform._transforms.{base + 2 * l}.autoregressive_net."
            pp = f"net._transform._transforms.{base + 2 * l + 1}."
            perm = np.asarray(self.perms[l], np.int64)
            tab[l, L.M_PERM] = len(ptab)
            ptab += list(perm) + list(np.argsort(perm))
            self.buffers[pp + "_permutation"] = torch.as_tensor(perm)
            o = take(Hp * Dp)
            tab[l, L.M_W0] = o
            idx[pa + "initial_layer.weight"] = o + np.arange(H)[:, None] * Dp + np.arange(D)[None, :]
            wm[pa + "initial_layer.weight"] = m_init
            self.buffers[pa + "initial_layer.mask"] = torch.as_tensor(m_init)
            self.buffers[pa + "initial_layer.degrees"] = torch.as_tensor(hid_deg)
            o = take(Hp)
            tab[l, L.M_B0] = o
            idx[pa + "initial_layer.bias"] = o + np.arange(H)
            o = take(Hp * Cp)
            tab[l, L.M_WC] = o
            idx[pa + "context_layer.weight"] = o + np.arange(H)[:, None] * Cp + np.arange(C)[None, :]
            o = take(Hp)
            tab[l, L.M_BC] = o
            idx[pa + "context_layer.bias"] = o + np.arange(H)
            for b in range(NB):
                pb = pa + f"blocks.{b}.linear."
                o = take(Hp * Hp)
                tab[l, L.M_BLK0 + 2 * b] = o
                idx[pb + "weight"] = o + np.arange(H)[:, None] * Hp + np.arange(H)[None, :]
                wm[pb + "weight"] = m_hid
                self.buffers[pb + "mask"] = torch.as_tensor(m_hid)
                self.buffers[pb + "degrees"] = torch.as_tensor(hid_deg)
                o = take(Hp)
                tab[l, L.M_BLK0 + 2 * b + 1] = o
                idx[pb + "bias"] = o + np.arange(H)
            o = take(self.OUTp * Hp)
            tab[l, L.M_WF] = o
            idx[pa + "final_layer.weight"] = o + np.arange(self.OUTM * D)[:, None] * Hp + np.arange(H)[None, :]
            wm[pa + "final_layer.weight"] = m_out
            self.buffers[pa + "final_layer.mask"] = torch.as_tensor(m_out)
            self.buffers[pa + "final_layer.degrees"] = torch.as_tensor(out_deg)
            o = take(self.OUTp)
            tab[l, L.M_BF] = o
            idx[pa + "final_layer.bias"] = o + np.arange(self.OUTM * D)
        self.n_params = off
        self.index = idx
        self._weight_masks = wm
        self.layer_tab = tab
        self.perm_tab = np.asarray(ptab, np.int32)

    def tables(self):
        return self.layer_tab.reshape(-1).astype(np.int32), self.perm_tab.astype(np.int32)

    def load_buffers(self, incoming: Dict[str, torch.Tensor]):
        """Adopt the permutations of a loaded reference state_dict (RandomPermutation buffers are
        drawn at construction, so they are data, not architecture).  Returns the new perm table."""
        base = 1 if self.zscore_input else 0
        changed = False
        ptab = []
        for l in range(self.T):
            key = f"net._transform._transforms.{base + 2 * l + 1}._permutation"
            if key in incoming:
                perm = incoming[key].detach().cpu().numpy().astype(np.int64)
                if perm.shape != (self.D,) or sorted(perm.tolist()) != list(range(self.D)):
                    raise ValueError(f"{key}: not a permutation of {self.D} features")
                if not np.array_equal(perm, self.perms[l]):
                    changed = True
                self.perms[l] = perm
                self.buffers[key] = torch.as_tensor(perm)
            ptab += list(self.perms[l]) + list(np.argsort(self.perms[l]))
        self.perm_tab = np.asarray(ptab, np.int32)
        return self.perm_tab if changed else None

    def fill_struct(self, s: "L.MafModel", nbuf: int):
        s.D, s.C, s.H, s.NB, s.T = self.D, self.C, self.H, self.NB, self.T
        s.Dp, s.Cp, s.Hp, s.OUTp = self.Dp, self.Cp, self.Hp, self.OUTp
        s.rpc0, s.rpc1, s.rpcf = self.rpc0, self.rpc1, self.rpcf
        s.wcap, s.nbuf, s.n_params = self.wcap, nbuf, self.n_params
        s.scale_softplus = 1 if self.scale_softplus else 0
        s.head, s.KB, s.OUTM = (0 if self.head == "affine" else 1), self.KB, self.OUTM
        s.tail_bound, s.min_w, s.min_h, s.min_d = (self.tail_bound, self.min_bin_width, self.min_bin_height,
                                                   self.min_derivative)
        s.isq = 1.0     # nflows' MADE has no `hidden_features` attribute: no 1/sqrt(H) logit scaling here
        return s


MafLayout.family = "maf"


@dataclass
class RatioLayout(_LayoutOps):
    """Packed layout of the NRE `resnet` classifier (include/sbi_b200.h `sbi_ratio_model`): nflows
    ResidualNet(in=Dt+Dx, out=1, hidden, context=None, num_blocks) as built by
    /root/reference/sbi/neural_nets/net_builders/classifier.py:172-235."""
    Dt: int
    Dx: int
    H: int = 50
    NB: int = 2
    wcap_target: int = 4096
    n_params: int = 0
    index: Dict[str, np.ndarray] = field(default_factory=dict, repr=False)

    def __post_init__(self):
        Dt, Dx, H, NB = self.Dt, self.Dx, self.H, self.NB
        self.Dtp, self.Dxp, self.Hp = round4(Dt), round4(Dx), round4(H)
        K0p, Hp = self.Dtp + self.Dxp, self.Hp
        cap = max(self.wcap_target, 4 * K0p, 4 * Hp)
        self.rpc0 = max(4, min(Hp, (cap // K0p) & ~3))
        self.rpc1 = max(4, min(Hp, (cap // Hp) & ~3))
        self.wcap = (max(self.rpc0 * K0p, self.rpc1 * Hp, 4 * Hp) + 31) & ~31
        off = 0

        def take(n):
            nonlocal off
            o = off
            off += round4(n)
            return o

        tab = np.zeros(4 + 4 * 8, np.int32)
        idx: Dict[str, np.ndarray] = {}
        o = take(Hp * K0p)
        tab[L.R_W0] = o
        cols = np.concatenate([np.arange(Dt), self.Dtp + np.arange(Dx)])   # [theta | pad | x | pad]
        idx["net.initial_layer.weight"] = o + np.arange(H)[:, None] * K0p + cols[None, :]
        o = take(Hp)
        tab[L.R_B0] = o
        idx["net.initial_layer.bias"] = o + np.arange(H)
        for b in range(NB):
            for j in range(2):
                o = take(Hp * Hp)
                tab[L.R_BLK0 + 4 * b + 2 * j] = o
                idx[f"net.blocks.{b}.linear_layers.{j}.weight"] = o + np.arange(H)[:, None] * Hp + np.arange(H)[None, :]
                o = take(Hp)
                tab[L.R_BLK0 + 4 * b + 2 * j + 1] = o
                idx[f"net.blocks.{b}.linear_layers.{j}.bias"] = o + np.arange(H)
        o = take(4 * Hp)
        tab[L.R_WF] = o
        idx["net.final_layer.weight"] = o + np.arange(H)[None, :]
        o = take(4)
        tab[L.R_BF] = o
        idx["net.final_layer.bias"] = o + np.arange(1)
        self.n_params = off
        self.index = idx
        self.tab = tab
        self.buffers = {}

    def tc_plan(self):
        """Gather map + stage table for the tcgen05 evaluation path (csrc/ratio_tc.cu), or None."""
        Dt, Dx, H, NB = self.Dt, self.Dx, self.H, self.NB
        if H != 50 or Dt + Dx > 56:
            return None
        Hp, K0p, Dtp = self.Hp, self.Dtp + self.Dxp, self.Dtp
        HP8 = (H + 7) & ~7
        K0 = Dt + Dx
        k0p8 = (K0 + 7) & ~7
        tab = np.zeros(L.SBI_NSF_TC_STRIDE, np.int32)

        def block(N, K, fill):
            blk = np.full((K // 4, N, 4), -1, np.int64)
            n = np.arange(N)[:, None] + 0 * np.arange(K)[None, :]
            k = np.arange(K)[None, :] + 0 * np.arange(N)[:, None]
            blk[k // 4, n, k % 4] = fill(n, k)
            return blk.reshape(-1)

        t = self.tab
        w0 = int(t[L.R_W0])

        def fill0(n, k):
            col = np.where(k < Dt, k, Dtp + (k - Dt))       # packed columns [theta | pad | x | pad]
            ok = (n < H) & (k < K0)
            return np.where(ok, w0 + n * K0p + np.clip(col, 0, K0p - 1), -1)

        def hidden(woff, N, rows_valid):
            def fill(n, k):
                ok = (n < rows_valid) & (k < H)
                return np.where(ok, woff + np.minimum(n, rows_valid - 1) * Hp + np.minimum(k, H - 1), -1)
            return block(N, HP8, fill)

        stages = [(block(64, k0p8, fill0), 64)]
        for b in range(NB):
            stages.append((hidden(int(t[L.R_BLK0 + 4 * b + 0]), 64, H), 64))
            stages.append((hidden(int(t[L.R_BLK0 + 4 * b + 2]), 64, H), 64))
        stages.append((hidden(int(t[L.R_WF]), 16, 1), 16))
        chunks, off, cap = [], 0, 0
        tab[0], tab[1] = len(stages), k0p8
        for s_, (hi, N) in enumerate(stages):
            nfl = 2 * hi.size
            tab[4 + 4 * s_: 8 + 4 * s_] = (off, nfl, N, 0)
            chunks += [hi, np.where(hi >= 0, -2 - hi, -1)]
            off += nfl
            cap = max(cap, nfl)
        return dict(src=np.concatenate(chunks).astype(np.int32), tab=tab.astype(np.int32),
                    stage_cap=int((cap + 31) & ~31), n_words=int(off))

    def fill_struct(self, s: "L.RatioModel", nbuf: int):
        s.Dt, s.Dx, s.H, s.NB = self.Dt, self.Dx, self.H, self.NB
        s.Dtp, s.Dxp, s.Hp = self.Dtp, self.Dxp, self.Hp
        s.rpc0, s.rpc1 = self.rpc0, self.rpc1
        s.wcap, s.nbuf, s.n_params = self.wcap, nbuf, self.n_params
        return s


RatioLayout.family = "ratio"


@dataclass
class FmLayout(_LayoutOps):
    """Packed layout of the flow-matching VectorFieldMLP (include/sbi_b200.h `sbi_fm_model`), tensor
    names as in /root/reference/sbi/neural_nets/net_builders/vector_field_nets.py:610-719 under the
    estimator attribute `net` (FlowMatchingEstimator.net)."""
    D: int
    C: int
    H: int = 100
    NL: int = 5
    TE: int = 32
    wcap_target: int = 3200
    n_params: int = 0
    index: Dict[str, np.ndarray] = field(default_factory=dict, repr=False)

    def __post_init__(self):
        D, C, H, NL, TE = self.D, self.C, self.H, self.NL, self.TE
        if NL < 2 or NL > 12:
            raise ValueError("2 <= num_layers <= 12")
        if TE % 2:
            raise ValueError("embedding dimension must be even")
        self.Dp, self.Cp, self.Hp, self.TEp = round4(D), round4(C), round4(H), round4(TE)
        Hp = self.Hp
        cap = max(self.wcap_target, 4 * 2 * Hp)

        def rows(rowlen, nmax):
            return max(4, min(nmax, (cap // rowlen) & ~3))

        self.rpc_i, self.rpc_c, self.rpc_m = rows(self.Dp, Hp), rows(self.Cp, Hp), rows(2 * Hp, Hp)
        self.rpc_t, self.rpc_h, self.rpc_o = rows(self.TEp, Hp), rows(Hp, Hp), rows(Hp, self.Dp)
        used = max(self.rpc_i * self.Dp, self.rpc_c * self.Cp, self.rpc_m * 2 * Hp, self.rpc_t * self.TEp,
                   self.rpc_h * Hp, self.rpc_o * Hp)
        self.wcap = (used + 31) & ~31
        off = 0

        def take(n):
            nonlocal off
            o = off
            off += round4(n)
            return o

        tab = np.zeros(L.F_LAYER0 + 4 * 12, np.int32)
        idx: Dict[str, np.ndarray] = {}

        def lin(name, slot_w, slot_b, N, K, Kp, cols=None, Np=None):
            Np = round4(N) if Np is None else Np
            o = take(Np * Kp)
            tab[slot_w] = o
            c = np.arange(K) if cols is None else cols
            idx[name + ".weight"] = o + np.arange(N)[:, None] * Kp + c[None, :]
            o = take(Np)
            tab[slot_b] = o
            idx[name + ".bias"] = o + np.arange(N)

        lin("net.input_layer", L.F_WI, L.F_BI, H, D, self.Dp)
        lin("net.condition_layer", L.F_WC, L.F_BC, H, C, self.Cp)
        lin("net.input_merge_layer", L.F_WM, L.F_BM, H, 2 * H, 2 * Hp,
            cols=np.concatenate([np.arange(H), Hp + np.arange(H)]))
        lin("net.time_linear_layer", L.F_WT, L.F_BT, H, TE, self.TEp)
        lin("net.output_layer", L.F_WO, L.F_BO, D, H, Hp)
        for i in range(NL):
            lin(f"net.layers.{i}", L.F_LAYER0 + 4 * i, L.F_LAYER0 + 4 * i + 1, H, H, Hp)
            o = take(Hp)
            tab[L.F_LAYER0 + 4 * i + 2] = o
            idx[f"net.layers_norm.{i}.weight"] = o + np.arange(H)
            o = take(Hp)
            tab[L.F_LAYER0 + 4 * i + 3] = o
            idx[f"net.layers_norm.{i}.bias"] = o + np.arange(H)
        self.n_params = off
        self.index = idx
        self.tab = tab
        self.buffers = {}

    def fill_struct(self, s: "L.FmModel", nbuf: int):
        s.D, s.C, s.H, s.NL, s.TE = self.D, self.C, self.H, self.NL, self.TE
        s.Dp, s.Cp, s.Hp, s.TEp = self.Dp, self.Cp, self.Hp, self.TEp
        s.rpc_i, s.rpc_c, s.rpc_m = self.rpc_i, self.rpc_c, self.rpc_m
        s.rpc_t, s.rpc_h, s.rpc_o = self.rpc_t, self.rpc_h, self.rpc_o
        s.wcap, s.nbuf, s.n_params = self.wcap, nbuf, self.n_params
        s.noise_scale, s.ln_eps = 1e-3, 1e-5
        return s


FmLayout.family = "fm"
