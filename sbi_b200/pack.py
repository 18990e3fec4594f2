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


@dataclass
class NsfLayout:
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
    wcap_target: int = 4096

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
    def trainable_mask(self) -> torch.Tensor:
        m = torch.zeros(self.n_params, dtype=torch.uint8)
        for v in self.index.values():
            m[torch.as_tensor(v.reshape(-1))] = 1
        return m

    def pack(self, state: Dict[str, torch.Tensor], out: torch.Tensor = None) -> torch.Tensor:
        """Reference-named tensors -> flat buffer (on out's device if given)."""
        flat = torch.zeros(self.n_params, dtype=torch.float32) if out is None else out
        for k, ix in self.index.items():
            src = state[k].detach().to(dtype=torch.float32, device=flat.device).reshape(-1)
            flat[torch.as_tensor(ix.reshape(-1), device=flat.device)] = src
        return flat

    def unpack(self, flat: torch.Tensor) -> Dict[str, torch.Tensor]:
        out = {}
        for k, ix in self.index.items():
            t = flat.detach()[torch.as_tensor(ix.reshape(-1), device=flat.device)]
            out[k] = t.reshape(ix.shape).clone()
        return out

    def num_real_params(self) -> int:
        return int(sum(v.size for v in self.index.values()))

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
        return s
