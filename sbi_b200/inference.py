"""Device-resident trainers and posteriors mirroring the reference's public API for the path.

`NPE` / `NLE` mirror `sbi.inference.NPE` (NPE-C, first round) and `sbi.inference.NLE`
(/root/reference/sbi/inference/trainers/npe/npe_base.py:81-418,
 /root/reference/sbi/inference/trainers/nle/nle_base.py:190-272) and the shared loop of
/root/reference/sbi/inference/trainers/base.py:499-563 (get_dataloaders), :1060-1148
(_run_training_loop), :1150-1225 (_train_epoch/_validate_epoch), :1254-1284 (_converged):

* same `append_simulations(theta, x).train(...)` / `build_posterior()` call sequence,
  argument names, defaults and stopping rule (validation loss, `stop_after_epochs`);
* same loss, optimiser (Adam, lr 5e-4), gradient clipping (5.0), 90/10 split, per-epoch
  reshuffle with `drop_last`, epoch-granular early stopping restoring the best weights.

What is different is where the work happens: the (theta, x) set lives in HBM, an epoch is ONE
CUDA-graph launch (steps_per_epoch x [fused fwd+bwd kernel -> partial-gradient reduce ->
clip+Adam kernel] + the validation pass), and the host reads three scalars per epoch instead
of syncing twice per step.
"""
from __future__ import annotations

import ctypes as C
import math
import os
import time
import warnings
from copy import deepcopy
from typing import Any, Callable, Dict, Optional, Union

import torch
from torch import Tensor, nn

from . import _lib as L
from .estimators import FlowEstimator, NSFEstimator
from .neural_nets import likelihood_nn, posterior_nn


def _process_device(device: str) -> str:
    """torchutils.py:54-103 (subset): only CUDA devices are valid here."""
    if device in ("gpu", "cuda"):
        device = "cuda:0"
    if not str(device).startswith("cuda"):
        raise RuntimeError(
            f"sbi_b200 trains on a CUDA (sm_100a) device only; got device={device!r}. "
            "There is no CPU fallback.")
    if not torch.cuda.is_available():
        raise RuntimeError("sbi_b200: no CUDA device available (no CPU fallback)")
    return str(device)


class _FlowTrainer:
    """Shared first-round trainer for density estimators (NPE: q(theta|x); NLE: q(x|theta))."""

    _swap = False   # NLE: estimator input = x, condition = theta

    def __init__(self, prior=None, density_estimator: Union[str, Callable] = "nsf",
                 device: str = "cuda", logging_level: Union[int, str] = "WARNING",
                 summary_writer=None, tracker=None, show_progress_bars: bool = False):
        self._prior = prior
        self._device = _process_device(device)
        if isinstance(density_estimator, str):
            factory = likelihood_nn if self._swap else posterior_nn
            self._build_neural_net = factory(model=density_estimator)
        else:
            self._build_neural_net = density_estimator
        self._neural_net: Optional[NSFEstimator] = None
        self._theta: Optional[Tensor] = None
        self._x: Optional[Tensor] = None
        self._show_progress_bars = show_progress_bars
        self._round = 0
        self.epoch = 0
        self._val_loss = float("Inf")
        self._summary: Dict[str, list] = dict(
            epochs_trained=[], best_validation_loss=[], validation_loss=[], training_loss=[],
            epoch_durations_sec=[])
        self._graphs = {}
        self._dist = None   # (rank, world) when data-parallel
        # multi-round bookkeeping (npe_base.py:188-299): round of every appended block and its proposal
        self._data_round_index: list = []
        self._proposal_roundwise: list = []
        self._round_rows: list = []          # rows of every appended block

    # ------------------------------------------------------------------ data
    def data_parallel(self, partition: str = "global"):
        """Train data-parallel over the initialised `torch.distributed` group (one process per
        GPU).  Replicas are identical: rank 0's initial network and statistics are
        broadcast when `train()` starts, each step's flat gradients are summed over the ranks
        (NVLink peer-memory kernel on one node, NCCL otherwise) and every rank applies the same
        deterministic clip + Adam, so the replicas stay bit-identical.

        partition="global" (SURVEY 8e): every rank holds the SAME simulations; the split and every
            epoch permutation come from rank 0, and rank r differentiates rows
            [r*B/G, (r+1)*B/G) of each global batch of `training_batch_size` rows, so batch
            composition, loss and update equal the single-GPU run (strong scaling).
        partition="local": every rank appends ITS OWN simulations and draws its own batches of
            `training_batch_size` rows (global batch = G x that; weak scaling)."""
        from . import parallel
        if partition not in ("global", "local"):
            raise ValueError("partition must be 'global' or 'local'")
        self._dist = parallel.world()
        self._partition = partition
        return self

    # ---- data-parallel helpers (no-ops on one process) ---------------------------------------------
    def _dp(self):
        rank, world = self._dist if getattr(self, "_dist", None) is not None else (0, 1)
        return rank, world, (getattr(self, "_partition", "global") if world > 1 else "local")

    def _dp_agree(self, value: int, what: str):
        """All ranks must see the same `value` (step counts, set sizes): a mismatch would leave a
        rank waiting in a collective forever."""
        rank, world, _ = self._dp()
        if world == 1:
            return
        t = torch.tensor([value, -value], dtype=torch.int64, device=self._device)
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        if int(t[0]) != value or int(-t[1]) != value:
            raise RuntimeError(f"data-parallel training needs the same {what} on every rank "
                               f"(this rank: {value}, max {int(t[0])}, min {int(-t[1])})")

    def _dp_split(self, N: int, n_train: int):
        """90/10 split indices (base.py:525-539); with partition='global' rank 0's split is used."""
        perm = torch.randperm(N)
        rank, world, part = self._dp()
        if world > 1 and part == "global":
            p = perm.to(self._device)
            torch.distributed.broadcast(p, 0)
            perm = p.cpu()
        return perm[:n_train], perm[n_train:]

    def _dp_sync_net(self, net):
        """Replicas start identical: parameters AND standardisation statistics of rank 0."""
        _, world, _ = self._dp()
        if world > 1:
            for t in list(net.parameters()) + list(net.buffers()):
                torch.distributed.broadcast(t.data, 0)
            net._cache.clear()

    def _dp_sum(self, values):
        """Sum of per-rank epoch statistics (list of floats), identical on every rank."""
        _, world, _ = self._dp()
        if world == 1:
            return values
        t = torch.tensor(values, dtype=torch.float64, device=self._device)
        torch.distributed.all_reduce(t)
        return t.tolist()

    def append_simulations(self, theta: Tensor, x: Tensor, proposal=None,
                           exclude_invalid_x: Optional[bool] = None, data_device: Optional[str] = None):
        """Store simulations (npe_base.py:188-299): float32 only, rows with NaN/Inf in x are
        dropped (user_input_checks.py:708-765, sbiutils.py:491-525)."""
        if not hasattr(self, "_data_round_index"):      # trainers with their own __init__ (NRE, FMPE)
            self._data_round_index, self._proposal_roundwise, self._round_rows = [], [], []
        # round of this block (npe_base.py:224-241): prior samples are round 0, anything else opens a new round
        if proposal is None or proposal is self._prior:
            current_round = 0
        elif not self._data_round_index:
            current_round = 1
        else:
            current_round = max(self._data_round_index) + 1
        if current_round > 0:
            if self._prior is None:
                raise ValueError("You did not pass a prior at initialization, but now you passed a proposal. "
                                 "Multi-round inference needs a prior.")
            if getattr(proposal, "default_x", "unset") is None:     # check_if_proposal_has_default_x
                raise ValueError("`proposal.default_x` is None: call `proposal.set_default_x(x_o)` first.")
            if getattr(proposal, "posterior_estimator", None) is not None \
                    and proposal.posterior_estimator is self._neural_net:
                raise ValueError("The proposal's posterior_estimator is the same object as the trainer's "
                                 "neural network; use trainer.build_posterior() or a deepcopy.")
        if exclude_invalid_x is None:
            exclude_invalid_x = current_round == 0
        if theta.dtype != torch.float32 or x.dtype != torch.float32:
            raise AssertionError("theta and x must be float32")
        if theta.shape[0] != x.shape[0]:
            raise AssertionError("Number of parameter sets must equal number of simulation outputs")
        xf = x.reshape(x.shape[0], -1)
        ok = ~torch.isnan(xf).any(1) & ~torch.isinf(xf).any(1)
        if not bool(ok.all()):
            if exclude_invalid_x:
                warnings.warn(f"Found {int((~ok).sum())} invalid simulations; they are excluded.",
                              stacklevel=2)
                theta, x = theta[ok], x[ok]
            else:
                warnings.warn(f"Found {int((~ok).sum())} simulations with NaN/Inf; they are kept "
                              "(multi-round losses normalise across the batch).", stacklevel=2)
        theta = theta.reshape(theta.shape[0], -1)
        self._data_round_index.append(current_round)
        self._proposal_roundwise.append(proposal)
        self._round_rows.append(int(theta.shape[0]))
        th = theta.to(self._device).contiguous()
        xx = x.to(self._device).contiguous()
        if self._theta is None:
            self._theta, self._x = th, xx
        else:
            self._theta = torch.cat([self._theta, th])
            self._x = torch.cat([self._x, xx])
        return self

    def get_simulations(self):
        return self._theta, self._x

    # ------------------------------------------------------------------ training
    def _inp_cond(self):
        """(estimator input set, condition set) as (N, .) device tensors."""
        th, xx = self._theta, self._x.reshape(self._x.shape[0], -1)
        return (xx, th) if self._swap else (th, xx)

    def train(self, training_batch_size: int = 200, learning_rate: float = 5e-4,
              validation_fraction: float = 0.1, stop_after_epochs: int = 20,
              max_num_epochs: int = 2 ** 31 - 1, clip_max_norm: Optional[float] = 5.0,
              calibration_kernel: Optional[Callable] = None, resume_training: bool = False,
              force_first_round_loss: bool = False, discard_prior_samples: bool = False,
              retrain_from_scratch: bool = False, show_train_summary: bool = False,
              dataloader_kwargs: Optional[dict] = None) -> NSFEstimator:
        if calibration_kernel is not None and self._swap:
            raise ValueError("calibration_kernel is an argument of the posterior-estimator trainers (NPE)")
        if self._theta is None:
            raise RuntimeError("call append_simulations() first")
        self._round = max(self._data_round_index) if getattr(self, "_data_round_index", None) else 0
        if self._round > 0 and not self._swap and not force_first_round_loss:
            # later rounds of NPE: atomic proposal correction (npe_c.py), eager steps
            return self._train_multiround(
                training_batch_size=training_batch_size, learning_rate=learning_rate,
                validation_fraction=validation_fraction, stop_after_epochs=stop_after_epochs,
                max_num_epochs=max_num_epochs, clip_max_norm=clip_max_norm, calibration_kernel=calibration_kernel,
                resume_training=resume_training, discard_prior_samples=discard_prior_samples,
                retrain_from_scratch=retrain_from_scratch)
        if self._round > 0 and discard_prior_samples:
            raise NotImplementedError("discard_prior_samples with the first-round loss is not implemented "
                                      "(append only the rounds to train on)")
        lib = L.load()
        dev = self._device
        N = self._theta.shape[0]
        rank, world, part = self._dp()
        glob = world > 1 and part == "global"
        self._dp_agree(N, "number of simulations") if glob else None
        # --- split (base.py:525-539): CPU global generator, like the reference
        n_train = int((1 - validation_fraction) * N)
        n_val = N - n_train
        if not resume_training or not hasattr(self, "train_indices"):
            self.train_indices, self.val_indices = self._dp_split(N, n_train)
        # --- network (npe_base.py:674-708): built from the CPU training split
        if self._neural_net is None or retrain_from_scratch:
            th_cpu = self._theta[self.train_indices.to(dev)].cpu()
            x_cpu = self._x[self.train_indices.to(dev)].cpu()
            self._neural_net = self._build_neural_net(th_cpu, x_cpu)
            del th_cpu, x_cpu
        net = self._neural_net.to(dev)
        self._neural_net = net
        if not resume_training:
            self._dp_sync_net(net)
        if not isinstance(net, FlowEstimator):
            raise TypeError(f"{type(self).__name__} needs an sbi_b200 flow estimator, "
                            f"got {type(net).__name__}")
        lay = net.layout
        if not net._embed_identity:
            raise NotImplementedError(
                "the fused trainer supports nn.Identity() embedding nets; train estimators with "
                "torch embedding nets through estimator.loss(...).backward()")
        P = lay.n_params
        B = min(training_batch_size, n_train)        # rows of one optimisation step (global batch if `glob`)
        Bv = min(training_batch_size, n_val)
        steps = n_train // B
        vsteps = n_val // Bv if Bv > 0 else 0
        if glob and B % world:
            raise ValueError(f"partition='global' needs training_batch_size ({B}) divisible by the "
                             f"number of ranks ({world})")
        Bl = B // world if glob else B                 # rows this rank differentiates per step
        Btot = B if glob else B * world                # rows behind one update
        self._dp_agree(steps, "number of steps per epoch")
        # validation rows: every rank evaluates its contiguous share of the epoch's validation order
        from .parallel import shard_range
        v_lo, v_hi = shard_range(vsteps * Bv, rank, world) if glob else (0, vsteps * Bv)

        if not resume_training or not hasattr(self, "_opt_state"):
            self._opt_state = torch.zeros(2 * P, dtype=torch.float32, device=dev)
            self._opt_step = torch.zeros(2, dtype=torch.int32, device=dev)
            self.epoch, self._val_loss = 0, float("Inf")
            self._best_val_loss = float("Inf")
            self._best_flat = None
            self._epochs_since_last_improvement = 0

        inp_all, cond_all = self._inp_cond()
        if hasattr(net, "with_dummy"):     # `made`: the network's dummy first feature (nn_utils.py:166-167)
            inp_all = net.with_dummy(inp_all).contiguous()
        train_idx = self.train_indices.to(dev)
        val_idx = self.val_indices.to(dev)
        # static buffers the epoch graph reads
        perm_buf = torch.empty(steps * B, dtype=torch.int64, device=dev)
        vperm_buf = torch.empty(max(vsteps * Bv, 1), dtype=torch.int64, device=dev)
        grad = torch.zeros(P, dtype=torch.float32, device=dev)
        sumsq = torch.zeros(lib.sbi_b200_sumsq_blocks(P), dtype=torch.float32, device=dev)
        n_part = net.vjp_parts(Bl)
        gpart = net._gpart(n_part)
        loss_acc = torch.zeros(2, dtype=torch.float32, device=dev)
        val_lp = torch.empty(max(vsteps * Bv, 1), dtype=torch.float32, device=dev)
        stats = torch.zeros(4, dtype=torch.float32, device=dev)   # train nll sum, bad, val nll sum, val bad
        mask = net.net._mask
        max_norm = float(clip_max_norm) if clip_max_norm is not None else 0.0
        # data-parallel on one node: sum the gradients with our peer-memory kernel (graph-capturable);
        # SBI_B200_NCCL=1 keeps the NCCL all-reduce (eager launches)
        peer, grad_local = None, grad
        if world > 1:
            from .parallel import make_gradient_exchange
            peer = make_gradient_exchange(P)      # None -> NCCL all-reduce, eager launches
            if peer is not None:
                grad_local = torch.zeros(P, dtype=torch.float32, device=dev)

        # calibration kernel (npe_base.py:373-378, :563-575): loss_r = K(x_r) * (-log q_r); the weights of
        # all simulations are evaluated once, a step's upstream gradient is -K(x_r) / B per row
        w_all = None
        if calibration_kernel is not None:
            w_all = torch.as_tensor(calibration_kernel(self._x), dtype=torch.float32).reshape(-1).to(dev).contiguous()
            if w_all.shape[0] != N:
                raise ValueError("calibration_kernel(x) must return one weight per simulation")
            g_rows = torch.empty(Bl, dtype=torch.float32, device=dev)
            lp_rows = torch.empty(Bl, dtype=torch.float32, device=dev)

        def run_epoch():
            """All kernels of one epoch on the current stream (graph-capturable)."""
            m_tr = net._model(nbuf=3)
            loss_acc.zero_()
            for s in range(steps):
                o = s * B + (rank * Bl if glob else 0)
                idx = perm_buf[o:o + Bl]
                rows = L.Rows(inp_all.data_ptr(), cond_all.data_ptr(), idx.data_ptr(), Bl, 0)
                if w_all is None:
                    net.vjp(m_tr, rows, Bl, None, -1.0 / Btot, None, gpart, None, None, loss_acc)
                else:
                    w = w_all[idx]
                    torch.mul(w, -1.0 / Btot, out=g_rows)
                    net.vjp(m_tr, rows, Bl, g_rows, 0.0, lp_rows, gpart, None, None, None)
                    fin = torch.isfinite(lp_rows)
                    loss_acc[0] -= (torch.where(fin, lp_rows, torch.zeros_like(lp_rows)) * w).sum()
                    loss_acc[1] += (~fin).sum()
                if peer is not None:   # gradients summed over NVLink peer memory (csrc/peer.cu)
                    L.check(lib.sbi_b200_reduce_partials(L.ptr(gpart), n_part, P, L.ptr(grad_local),
                                                         L.stream_ptr()), "reduce_partials")
                    peer.sum(grad_local, grad, mask, sumsq)
                    L.check(lib.sbi_b200_adam_clip_step_norm(
                        L.ptr(net.flat.data), L.ptr(grad), L.ptr(self._opt_state), L.ptr(self._opt_step),
                        L.ptr(mask), P, learning_rate, 0.9, 0.999, 1e-8, max_norm, 1.0, L.ptr(sumsq),
                        peer.n_sumsq, L.stream_ptr()), "adam_clip_step")
                elif world > 1:   # the clip norm must be taken on the all-reduced gradient
                    L.check(lib.sbi_b200_reduce_partials(L.ptr(gpart), n_part, P, L.ptr(grad),
                                                         L.stream_ptr()), "reduce_partials")
                    torch.distributed.all_reduce(grad)
                    L.check(lib.sbi_b200_adam_clip_step(
                        L.ptr(net.flat.data), L.ptr(grad), L.ptr(self._opt_state), L.ptr(self._opt_step),
                        L.ptr(mask), P, learning_rate, 0.9, 0.999, 1e-8, max_norm, 1.0,
                        L.stream_ptr()), "adam_clip_step")
                else:           # single GPU: the reduction kernel also emits sum(g^2) partials
                    L.check(lib.sbi_b200_reduce_partials_norm(L.ptr(gpart), n_part, P, L.ptr(grad), L.ptr(mask),
                                                              L.ptr(sumsq), L.stream_ptr()), "reduce_partials")
                    L.check(lib.sbi_b200_adam_clip_step_norm(
                        L.ptr(net.flat.data), L.ptr(grad), L.ptr(self._opt_state), L.ptr(self._opt_step),
                        L.ptr(mask), P, learning_rate, 0.9, 0.999, 1e-8, max_norm, 1.0, L.ptr(sumsq),
                        sumsq.shape[0], L.stream_ptr()), "adam_clip_step")
            stats[0:2].copy_(loss_acc)
            stats[2:4].zero_()
            if v_hi > v_lo:
                m_ev = net._model(nbuf=2)
                vrows = vperm_buf[v_lo:v_hi]
                vlp = val_lp[:v_hi - v_lo]
                rows = L.Rows(inp_all.data_ptr(), cond_all.data_ptr(), vrows.data_ptr(), v_hi - v_lo, 0)
                # validation rows go through the tensor-core kernel when the model fits it (its
                # operands are re-packed from the just-updated parameters inside _tc_state)
                tc_ev = net._tc_state(m_ev) if v_hi - v_lo >= net.TC_MIN_ROWS else None
                if tc_ev is not None:
                    L.check(lib.sbi_b200_nsf_logprob_tc(C.byref(m_ev), C.byref(tc_ev), C.byref(rows),
                                                        L.ptr(vlp), None, L.stream_ptr()), "nsf_logprob_tc")
                else:
                    L.check(net.fam.fn("logprob")(C.byref(m_ev), C.byref(rows), L.ptr(vlp), None,
                                                  L.stream_ptr()), "flow_logprob")
                if w_all is None:
                    # -sum of the finite log-probs and the count of non-finite ones, one fused launch
                    L.check(lib.sbi_b200_nll_stats(L.ptr(vlp), v_hi - v_lo, L.ptr(stats[2:]), L.stream_ptr()),
                            "nll_stats")
                else:
                    fin = torch.isfinite(vlp)
                    stats[2] = -(torch.where(fin, vlp, torch.zeros_like(vlp)) * w_all[vrows]).sum()
                    stats[3] = (~fin).sum().float()

        def fill_perms():
            perm_buf.copy_(train_idx[torch.randperm(n_train, device=dev)[:steps * B]])
            if vsteps > 0:
                vperm_buf.copy_(val_idx[torch.randperm(n_val, device=dev)[:vsteps * Bv]])
            if glob:      # one epoch order for all ranks: rank 0's
                torch.distributed.broadcast(perm_buf, 0)
                if vsteps > 0:
                    torch.distributed.broadcast(vperm_buf, 0)

        # warm-up (also sets kernel attributes) on a throw-away copy of the state, then capture
        graph = None
        if world == 1 or peer is not None:
            snap = (net.flat.data.clone(), self._opt_state.clone(), self._opt_step.clone())
            rng = torch.cuda.get_rng_state(dev)      # the warm-up's permutation draw leaves no trace:
            fill_perms()                             # a seed gives the same run with and without graphs
            torch.cuda.set_rng_state(rng, dev)
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                run_epoch()
            torch.cuda.current_stream().wait_stream(side)
            net.flat.data.copy_(snap[0]); self._opt_state.copy_(snap[1]); self._opt_step.copy_(snap[2])
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                run_epoch()
            net.flat.data.copy_(snap[0]); self._opt_state.copy_(snap[1]); self._opt_step.copy_(snap[2])

        def converged() -> bool:
            """base.py:1254-1284."""
            if self.epoch == 0 or self._val_loss < self._best_val_loss:
                self._best_val_loss = self._val_loss
                self._epochs_since_last_improvement = 0
                self._best_flat = net.flat.data.clone()
            else:
                self._epochs_since_last_improvement += 1
            if self._epochs_since_last_improvement > stop_after_epochs - 1:
                net.flat.data.copy_(self._best_flat)
                return True
            return False

        while self.epoch <= max_num_epochs and not converged():
            t0 = time.time()
            fill_perms()
            if graph is not None:
                graph.replay()
            else:
                run_epoch()
            s = self._dp_sum(stats.tolist())   # the one host sync of the epoch (+ one tiny all-reduce)
            if s[1] > 0 or s[3] > 0:
                raise AssertionError("NaN/Inf present in NPE loss.")
            train_loss = s[0] / (steps * Btot)
            self._val_loss = s[2] / (vsteps * Bv * (1 if glob else world)) if vsteps > 0 else float("nan")
            self._summary["training_loss"].append(train_loss)
            self._summary["validation_loss"].append(self._val_loss)
            self._summary["epoch_durations_sec"].append(time.time() - t0)
            self.epoch += 1

        if self.epoch > max_num_epochs:   # base.py:1122-1129
            if self._val_loss < self._best_val_loss:
                self._best_val_loss = self._val_loss
                self._best_flat = net.flat.data.clone()
            elif self._best_flat is not None:
                net.flat.data.copy_(self._best_flat)
            warnings.warn(f"Maximum number of epochs `max_num_epochs={max_num_epochs}` reached, "
                          "but network has not yet fully converged.", stacklevel=2)
        self._summary["epochs_trained"].append(self.epoch)
        self._summary["best_validation_loss"].append(self._best_val_loss)
        net.zero_grad(set_to_none=True)
        if peer is not None:
            timed_out = peer.error()
            peer.close()
            if timed_out:
                raise RuntimeError("peer-memory gradient exchange timed out (a rank fell behind or died)")
        return deepcopy(net)

    def _train_multiround(self, training_batch_size, learning_rate, validation_fraction, stop_after_epochs,
                          max_num_epochs, clip_max_norm, calibration_kernel, resume_training,
                          discard_prior_samples, retrain_from_scratch):
        """Rounds > 0 of NPE-C (npe_c.py:126-231 + the shared loop of trainers/base.py:1060-1284): the
        network of the previous round keeps training on the simulations of all rounds with the atomic
        proposal-posterior loss.  Steps are eager (log-prob kernel -> soft-max head in torch -> fused
        forward+backward kernel through autograd -> clip_grad_norm_ -> Adam), exactly the reference's
        operation sequence; the B x num_atoms evaluations per step are what the kernels are for."""
        from .multiround import atomic_log_prob_proposal_posterior, clamp_num_atoms
        from .posteriors import prior_to_device
        if getattr(self, "_dist", None) is not None and self._dp()[1] > 1:
            raise NotImplementedError("multi-round training is single-process")
        dev = self._device
        num_atoms = int(getattr(self, "_num_atoms", 10))
        combined = bool(getattr(self, "_use_combined_loss", False))
        start_round = int(bool(discard_prior_samples) and self._round > 0)        # base.py _get_start_index
        rounds = torch.repeat_interleave(torch.as_tensor(self._data_round_index),
                                         torch.as_tensor(self._round_rows)).to(dev)
        sel = torch.nonzero(rounds >= start_round).reshape(-1)
        theta_all, x_all = self._theta[sel], self._x[sel]
        masks_all = (rounds[sel] == 0).to(torch.float32)                         # mask_sims_from_prior
        N = theta_all.shape[0]
        n_train = int((1 - validation_fraction) * N)
        n_val = N - n_train
        if not resume_training or getattr(self, "_mr_split_n", None) != N:
            perm = torch.randperm(N)
            self.train_indices, self.val_indices = perm[:n_train], perm[n_train:]
            self._mr_split_n = N
        if self._neural_net is None or retrain_from_scratch:
            tr = self.train_indices.to(dev)
            self._neural_net = self._build_neural_net(theta_all[tr].cpu(), x_all[tr].cpu())
        net = self._neural_net.to(dev)
        self._neural_net = net
        prior = prior_to_device(self._prior, dev)
        B, Bv = min(training_batch_size, n_train), min(training_batch_size, n_val)
        steps, vsteps = n_train // B, (n_val // Bv if Bv > 0 else 0)
        if not resume_training or not hasattr(self, "_mr_opt"):
            self._mr_opt = torch.optim.Adam(list(net.parameters()), lr=learning_rate)
            self.epoch, self._val_loss = 0, float("Inf")
            self._best_val_loss, self._best_flat, self._epochs_since_last_improvement = float("Inf"), None, 0
        opt = self._mr_opt
        train_idx, val_idx = self.train_indices.to(dev), self.val_indices.to(dev)
        w_all = None
        if calibration_kernel is not None:
            w_all = torch.as_tensor(calibration_kernel(x_all), dtype=torch.float32).reshape(-1).to(dev)

        def losses_of(idx):
            th, xx, mk = theta_all[idx], x_all[idx], masks_all[idx]
            lp = atomic_log_prob_proposal_posterior(net, prior, th, xx, mk, num_atoms, combined)
            if not bool(torch.isfinite(lp).all()):
                raise AssertionError("NaN/Inf present in NPE loss.")
            return -lp if w_all is None else -lp * w_all[idx]

        def converged() -> bool:
            if self.epoch == 0 or self._val_loss < self._best_val_loss:
                self._best_val_loss, self._epochs_since_last_improvement = self._val_loss, 0
                self._best_flat = net.flat.data.clone()
            else:
                self._epochs_since_last_improvement += 1
            if self._epochs_since_last_improvement > stop_after_epochs - 1:
                net.flat.data.copy_(self._best_flat)
                return True
            return False

        clamp_num_atoms(num_atoms, min(B, Bv) if vsteps > 0 else B)      # warn once, like the reference does per call
        with warnings.catch_warnings():
            warnings.filterwarnings("ignore", message="num_atoms=")
            while self.epoch <= max_num_epochs and not converged():
                t0 = time.time()
                net.train()
                perm = train_idx[torch.randperm(n_train, device=dev)]
                tsum = torch.zeros((), device=dev)
                for s_ in range(steps):
                    opt.zero_grad()
                    losses = losses_of(perm[s_ * B:(s_ + 1) * B])
                    losses.mean().backward()
                    tsum += losses.detach().sum()
                    if clip_max_norm is not None:
                        torch.nn.utils.clip_grad_norm_(net.parameters(), max_norm=clip_max_norm)
                    opt.step()
                net.eval()
                vsum = torch.zeros((), device=dev)
                with torch.no_grad():
                    vperm = val_idx[torch.randperm(n_val, device=dev)] if vsteps > 0 else None
                    for s_ in range(vsteps):
                        vsum += losses_of(vperm[s_ * Bv:(s_ + 1) * Bv]).sum()
                self._summary["training_loss"].append(float(tsum.item()) / (steps * B))
                self._val_loss = float(vsum.item()) / (vsteps * Bv) if vsteps > 0 else float("nan")
                self._summary["validation_loss"].append(self._val_loss)
                self._summary["epoch_durations_sec"].append(time.time() - t0)
                self.epoch += 1
        if self.epoch > max_num_epochs:
            if self._val_loss < self._best_val_loss:
                self._best_val_loss, self._best_flat = self._val_loss, net.flat.data.clone()
            elif self._best_flat is not None:
                net.flat.data.copy_(self._best_flat)
            warnings.warn(f"Maximum number of epochs `max_num_epochs={max_num_epochs}` reached, "
                          "but network has not yet fully converged.", stacklevel=3)
        self._summary["epochs_trained"].append(self.epoch)
        self._summary["best_validation_loss"].append(self._best_val_loss)
        net.zero_grad(set_to_none=True)
        net._cache.clear()
        return deepcopy(net)

    @property
    def summary(self):
        return self._summary


class NPE(_FlowTrainer):
    """Neural posterior estimation, first round (reference: NPE_C, npe_c.py:91 / npe_base.py)."""
    _swap = False

    def train(self, num_atoms: int = 10, training_batch_size: int = 200, learning_rate: float = 5e-4,
              validation_fraction: float = 0.1, stop_after_epochs: int = 20, max_num_epochs: int = 2 ** 31 - 1,
              clip_max_norm: Optional[float] = 5.0, calibration_kernel: Optional[Callable] = None,
              resume_training: bool = False, force_first_round_loss: bool = False,
              discard_prior_samples: bool = False, use_combined_loss: bool = False,
              retrain_from_scratch: bool = False, show_train_summary: bool = False,
              dataloader_kwargs: Optional[dict] = None):
        """npe_c.py:126-231 (same argument order): `num_atoms` / `use_combined_loss` only matter from the
        second round on (atomic proposal correction, `_train_multiround`)."""
        self._num_atoms, self._use_combined_loss = num_atoms, use_combined_loss
        return super().train(training_batch_size=training_batch_size, learning_rate=learning_rate,
                             validation_fraction=validation_fraction, stop_after_epochs=stop_after_epochs,
                             max_num_epochs=max_num_epochs, clip_max_norm=clip_max_norm,
                             calibration_kernel=calibration_kernel, resume_training=resume_training,
                             force_first_round_loss=force_first_round_loss,
                             discard_prior_samples=discard_prior_samples,
                             retrain_from_scratch=retrain_from_scratch, show_train_summary=show_train_summary,
                             dataloader_kwargs=dataloader_kwargs)

    def build_posterior(self, density_estimator: Optional[nn.Module] = None, prior=None,
                        sample_with: str = "direct", **kwargs):
        from .posteriors import DirectPosterior
        if sample_with != "direct":
            raise NotImplementedError("NPE.build_posterior supports sample_with='direct'")
        est = density_estimator if density_estimator is not None else self._neural_net
        prior = prior if prior is not None else self._prior
        return DirectPosterior(deepcopy(est), prior, device=self._device)


NPE_C = NPE
SNPE = NPE


class NLE(_FlowTrainer):
    """Neural likelihood estimation (reference: NLE_A, nle_base.py:190-272, _loss :380-392)."""
    _swap = True


NLE_A = NLE
SNLE = NLE


# =================================================================================================
def _nle_build_posterior(self, density_estimator=None, prior=None, sample_with: str = "mcmc",
                         mcmc_method: str = "slice_np_vectorized", mcmc_parameters: Optional[dict] = None,
                         rejection_sampling_parameters: Optional[dict] = None, **kwargs):
    """nle_base.py:274-378 (`sample_with` in {"mcmc", "rejection"})."""
    from .posteriors import MCMCPosterior, RejectionPosterior
    from .potentials import likelihood_estimator_based_potential
    est = deepcopy(density_estimator if density_estimator is not None else self._neural_net)
    prior = prior if prior is not None else self._prior
    potential_fn, theta_transform = likelihood_estimator_based_potential(est, prior, x_o=None)
    if sample_with == "mcmc":
        return MCMCPosterior(potential_fn, proposal=prior, theta_transform=theta_transform, method=mcmc_method,
                             device=self._device, **(mcmc_parameters or {}))
    if sample_with == "rejection":
        return RejectionPosterior(potential_fn, proposal=prior, device=self._device,
                                  **(rejection_sampling_parameters or {}))
    raise NotImplementedError(sample_with)


NLE.build_posterior = _nle_build_posterior


class NRE_B(_FlowTrainer):
    """Neural ratio estimation, NRE-B / SRE (reference: trainers/nre/nre_base.py:184-309 train,
    :396-415 `_classifier_logits`; trainers/nre/nre_b.py:157-182 `_loss`): 1-out-of-`num_atoms`
    classification of the jointly drawn (theta, x) pair against `num_atoms - 1` contrastive thetas
    from the same batch.  The classifier forward / backward run in the ratio kernels; the contrastive
    index draw and the softmax head are a handful of torch device ops."""

    def __init__(self, prior=None, classifier: Union[str, Callable] = "resnet", device: str = "cuda",
                 logging_level: Union[int, str] = "warning", summary_writer=None, tracker=None,
                 show_progress_bars: bool = False):
        from .ratio import classifier_nn
        self._prior = prior
        self._device = _process_device(device)
        self._build_neural_net = classifier_nn(classifier) if isinstance(classifier, str) else classifier
        self._neural_net = None
        self._theta = self._x = None
        self.epoch, self._val_loss = 0, float("Inf")
        self._summary = dict(epochs_trained=[], best_validation_loss=[], validation_loss=[],
                             training_loss=[], epoch_durations_sec=[])
        self._dist = None

    @staticmethod
    def _contrastive_choices(B: int, k: int, device, rows: Optional[tuple] = None) -> Tensor:
        """(n, k) indices j != i into a batch of B, distinct per row, uniform, for the batch rows
        i in [rows[0], rows[1]) (default: all B): same law as
        `torch.multinomial((1 - eye) / (B - 1), k, replacement=False)` (nre_base.py:406-408) without the
        O(B^2) probability matrix: k draws without replacement from range(B-1), shifted past i."""
        lo, hi = rows if rows is not None else (0, B)
        n = hi - lo
        if B - 1 <= 4096:
            draws = torch.multinomial(torch.ones(n, B - 1, device=device), k, replacement=False)
        else:
            draws = torch.randint(0, B - 1, (n, k), device=device)
            while True:
                srt = draws.sort(dim=1).values
                dup = (srt[:, 1:] == srt[:, :-1]).any(dim=1)
                nd = int(dup.sum().item())
                if nd == 0:
                    break
                draws[dup] = torch.randint(0, B - 1, (nd, k), device=device)
        own = torch.arange(lo, hi, device=device).unsqueeze(1)
        return draws + (draws >= own).long()

    def _logits_on(self, net, idx: Tensor, num_atoms: int, choices: Optional[Tensor] = None,
                   rows: Optional[tuple] = None) -> Tensor:
        """`_classifier_logits` (nre_base.py:396-415) of the batch rows [rows[0], rows[1]) (default: all)
        of the batch `idx`: (n, num_atoms) logits, column 0 the jointly drawn pair; the contrastive thetas
        of a row come from the WHOLE batch (with the global batch on every rank the
        data-parallel loss keeps the single-GPU semantics)."""
        from .ratio import _RatioFn
        B = idx.shape[0]
        lo, hi = rows if rows is not None else (0, B)
        if choices is None:
            choices = self._contrastive_choices(B, num_atoms - 1, idx.device, (lo, hi))
        local = torch.cat([torch.arange(lo, hi, device=idx.device).unsqueeze(1), choices], dim=1)   # (n, A)
        ti = idx[local].reshape(-1).contiguous()
        xi = idx[lo:hi].repeat_interleave(num_atoms).contiguous()
        return _RatioFn.apply(net.net.flat, self._theta, self._x2d, net, ti, xi, False).reshape(hi - lo, num_atoms)

    def _loss_on(self, net, idx: Tensor, num_atoms: int, choices: Optional[Tensor] = None,
                 rows: Optional[tuple] = None) -> Tensor:
        """NRE-B loss (nre_b.py:157-182): 1-out-of-`num_atoms` cross-entropy."""
        logits = self._logits_on(net, idx, num_atoms, choices, rows)
        log_prob = logits[:, 0] - torch.logsumexp(logits, dim=-1)
        return -torch.mean(log_prob)

    def train(self, num_atoms: int = 10, training_batch_size: int = 200, learning_rate: float = 5e-4,
              validation_fraction: float = 0.1, stop_after_epochs: int = 20, max_num_epochs: int = 2 ** 31 - 1,
              clip_max_norm: Optional[float] = 5.0, resume_training: bool = False,
              discard_prior_samples: bool = False, retrain_from_scratch: bool = False,
              show_train_summary: bool = False, dataloader_kwargs: Optional[dict] = None):
        if self._theta is None:
            raise RuntimeError("call append_simulations() first")
        lib = L.load()
        dev = self._device
        N = self._theta.shape[0]
        self._x2d = self._x.reshape(N, -1).contiguous()
        rank, world, part = self._dp()
        glob = world > 1 and part == "global"
        if glob:
            self._dp_agree(N, "number of simulations")
        n_train = int((1 - validation_fraction) * N)
        n_val = N - n_train
        if not resume_training or not hasattr(self, "train_indices"):
            self.train_indices, self.val_indices = self._dp_split(N, n_train)
        B = min(training_batch_size, n_train)
        Bv = min(training_batch_size, n_val)
        clipped = min(B, Bv)
        num_atoms = int(min(max(num_atoms, 2), clipped))     # nre_base.py:236-238 (clamp to batch size)
        if self._neural_net is None or retrain_from_scratch:
            tr = self.train_indices.to(dev)
            self._neural_net = self._build_neural_net(self._theta[tr].cpu(), self._x[tr].cpu())
        net = self._neural_net.to(dev)
        self._neural_net = net
        if not resume_training:
            self._dp_sync_net(net)
        P = net.layout.n_params
        if not resume_training or not hasattr(self, "_opt_state"):
            self._opt_state = torch.zeros(2 * P, dtype=torch.float32, device=dev)
            self._opt_step = torch.zeros(2, dtype=torch.int32, device=dev)
            self.epoch, self._val_loss = 0, float("Inf")
            self._best_val_loss, self._best_flat, self._epochs_since_last_improvement = float("Inf"), None, 0
        train_idx, val_idx = self.train_indices.to(dev), self.val_indices.to(dev)
        steps, vsteps = n_train // B, (n_val // Bv if Bv > 0 else 0)
        self._dp_agree(steps, "number of steps per epoch")
        self._dp_agree(vsteps, "number of validation steps per epoch")
        if glob and (B % world or Bv % world):
            raise ValueError(f"partition='global' needs the batch sizes ({B}, {Bv}) divisible by the "
                             f"number of ranks ({world})")
        # rows of each (global) batch whose loss this rank differentiates / evaluates
        t_rows = (rank * (B // world), (rank + 1) * (B // world)) if glob else (0, B)
        v_rows = (rank * (Bv // world), (rank + 1) * (Bv // world)) if glob else (0, Bv)
        max_norm = float(clip_max_norm) if clip_max_norm is not None else 0.0
        peer = None
        if world > 1:
            from .parallel import make_gradient_exchange
            peer = make_gradient_exchange(P)          # None -> NCCL all-reduce, eager launches
        grad_sum = torch.zeros(P, dtype=torch.float32, device=dev) if peer is not None else None
        sumsq = torch.zeros(peer.n_sumsq, dtype=torch.float32, device=dev) if peer is not None else None

        def converged() -> bool:
            if self.epoch == 0 or self._val_loss < self._best_val_loss:
                self._best_val_loss, self._epochs_since_last_improvement = self._val_loss, 0
                self._best_flat = net.flat.data.clone()
            else:
                self._epochs_since_last_improvement += 1
            if self._epochs_since_last_improvement > stop_after_epochs - 1:
                net.flat.data.copy_(self._best_flat)
                return True
            return False

        # One CUDA graph per optimisation step and one per validation step: a step is ~15 small
        # launches (contrastive draws, index gathers, logits kernel, logsumexp, VJP kernel, reduce,
        # clip+Adam) that the host cannot issue as fast as the GPU runs them at the default batch
        # of 200.  The batch indices come from a static buffer filled before each replay; random
        # draws inside the graph use torch's graph-safe Philox offsets.
        idx_buf = torch.zeros(B, dtype=torch.int64, device=dev)
        vidx_buf = torch.zeros(max(Bv, 1), dtype=torch.int64, device=dev)
        train_sum = torch.zeros((), device=dev)
        val_sum = torch.zeros((), device=dev)

        def train_step():
            net.net.flat.grad = None
            # every rank's rows weigh 1/world of the update's batch mean; gradients are summed
            loss = self._loss_on(net, idx_buf, num_atoms, rows=t_rows) / world
            loss.backward()
            train_sum.add_(loss.detach())
            if peer is not None:       # gradient sum over NVLink peer memory + sum(g^2) partials
                peer.sum(net.flat.grad, grad_sum, net.net._mask, sumsq)
                L.check(lib.sbi_b200_adam_clip_step_norm(
                    L.ptr(net.flat.data), L.ptr(grad_sum), L.ptr(self._opt_state), L.ptr(self._opt_step),
                    L.ptr(net.net._mask), P, learning_rate, 0.9, 0.999, 1e-8, max_norm, 1.0, L.ptr(sumsq),
                    peer.n_sumsq, L.stream_ptr()), "adam_clip_step")
                return
            if world > 1:
                torch.distributed.all_reduce(net.flat.grad)
            L.check(lib.sbi_b200_adam_clip_step(
                L.ptr(net.flat.data), L.ptr(net.flat.grad), L.ptr(self._opt_state), L.ptr(self._opt_step),
                L.ptr(net.net._mask), P, learning_rate, 0.9, 0.999, 1e-8, max_norm, 1.0, L.stream_ptr()),
                "adam_clip_step")

        def val_step():
            with torch.no_grad():
                val_sum.add_(self._loss_on(net, vidx_buf, num_atoms, rows=v_rows) / world)

        g_train = g_val = None
        if (os.environ.get("SBI_B200_NRE_GRAPH", "1") != "0" and B - 1 <= 4096 and steps > 0
                and (world == 1 or peer is not None)):
            snap = (net.flat.data.clone(), self._opt_state.clone(), self._opt_step.clone())
            idx_buf.copy_(train_idx[:B])
            if vsteps > 0:
                vidx_buf.copy_(val_idx[:Bv])
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):          # warm-up outside capture (allocations, kernel attributes)
                for _ in range(3):
                    train_step()
                if vsteps > 0:
                    val_step()
            torch.cuda.current_stream().wait_stream(side)
            g_train = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g_train):
                train_step()
            if vsteps > 0:
                g_val = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g_val):
                    val_step()
            net.flat.data.copy_(snap[0]); self._opt_state.copy_(snap[1]); self._opt_step.copy_(snap[2])

        while self.epoch <= max_num_epochs and not converged():
            t0 = time.time()
            perm = train_idx[torch.randperm(n_train, device=dev)]
            vperm = val_idx[torch.randperm(n_val, device=dev)] if vsteps > 0 else None
            if glob:      # one epoch order for all ranks: rank 0's
                torch.distributed.broadcast(perm, 0)
                if vperm is not None:
                    torch.distributed.broadcast(vperm, 0)
            train_sum.zero_()
            for s in range(steps):
                idx_buf.copy_(perm[s * B:(s + 1) * B])
                if g_train is not None:
                    g_train.replay()
                else:
                    train_step()
            val_sum.zero_()
            for s in range(vsteps):
                vidx_buf.copy_(vperm[s * Bv:(s + 1) * Bv])
                if g_val is not None:
                    g_val.replay()
                else:
                    val_step()
            tl, vl = self._dp_sum([float(train_sum.item()), float(val_sum.item())])
            if not (math.isfinite(tl) and math.isfinite(vl)):
                raise AssertionError("NaN/Inf present in NRE-B loss.")
            # the reference divides the sum of per-batch MEAN losses by steps * batch_size
            self._summary["training_loss"].append(tl / (steps * B))
            self._val_loss = vl / (vsteps * Bv) if vsteps > 0 else float("nan")
            self._summary["validation_loss"].append(self._val_loss)
            self._summary["epoch_durations_sec"].append(time.time() - t0)
            self.epoch += 1
        if self.epoch > max_num_epochs:
            if self._val_loss < self._best_val_loss:
                self._best_val_loss, self._best_flat = self._val_loss, net.flat.data.clone()
            elif self._best_flat is not None:
                net.flat.data.copy_(self._best_flat)
            warnings.warn(f"Maximum number of epochs `max_num_epochs={max_num_epochs}` reached, "
                          "but network has not yet fully converged.", stacklevel=2)
        self._summary["epochs_trained"].append(self.epoch)
        self._summary["best_validation_loss"].append(self._best_val_loss)
        net.zero_grad(set_to_none=True)
        if peer is not None:
            timed_out = peer.error()
            peer.close()
            if timed_out:
                raise RuntimeError("peer-memory gradient exchange timed out (a rank fell behind or died)")
        return deepcopy(net)

    def build_posterior(self, density_estimator=None, prior=None, sample_with: str = "mcmc",
                        mcmc_method: str = "slice_np_vectorized", mcmc_parameters: Optional[dict] = None,
                        rejection_sampling_parameters: Optional[dict] = None, **kwargs):
        """nre_base.py:311-394."""
        from .posteriors import MCMCPosterior, RejectionPosterior
        from .potentials import ratio_estimator_based_potential
        est = deepcopy(density_estimator if density_estimator is not None else self._neural_net)
        prior = prior if prior is not None else self._prior
        potential_fn, theta_transform = ratio_estimator_based_potential(est, prior, x_o=None)
        if sample_with == "mcmc":
            return MCMCPosterior(potential_fn, proposal=prior, theta_transform=theta_transform, method=mcmc_method,
                                 device=self._device, **(mcmc_parameters or {}))
        if sample_with == "rejection":
            return RejectionPosterior(potential_fn, proposal=prior, device=self._device,
                                      **(rejection_sampling_parameters or {}))
        raise NotImplementedError(sample_with)


SNRE_B = NRE_B
SNRE = NRE_B


class NRE_A(NRE_B):
    """AALR / NRE-A (reference: trainers/nre/nre_a.py:103-190): binary classification of the jointly drawn
    pair against ONE contrastive pair; same trainer, two atoms, BCE head."""

    def train(self, training_batch_size: int = 200, learning_rate: float = 5e-4, validation_fraction: float = 0.1,
              stop_after_epochs: int = 20, max_num_epochs: int = 2 ** 31 - 1, clip_max_norm: Optional[float] = 5.0,
              resume_training: bool = False, discard_prior_samples: bool = False, retrain_from_scratch: bool = False,
              show_train_summary: bool = False, dataloader_kwargs: Optional[dict] = None):
        return NRE_B.train(self, num_atoms=2, training_batch_size=training_batch_size, learning_rate=learning_rate,
                           validation_fraction=validation_fraction, stop_after_epochs=stop_after_epochs,
                           max_num_epochs=max_num_epochs, clip_max_norm=clip_max_norm,
                           resume_training=resume_training, discard_prior_samples=discard_prior_samples,
                           retrain_from_scratch=retrain_from_scratch, show_train_summary=show_train_summary,
                           dataloader_kwargs=dataloader_kwargs)

    def _loss_on(self, net, idx, num_atoms, choices=None, rows=None):
        from .multiround import nre_a_loss
        return nre_a_loss(self._logits_on(net, idx, 2, choices, rows))


SNRE_A = NRE_A
AALR = NRE_A


class BNRE(NRE_A):
    """Balanced NRE (reference: trainers/nre/bnre.py:103-200): NRE-A plus the balancing regulariser
    `regularization_strength * (E[sigmoid(l_joint) + sigmoid(l_marginal) - 1])^2`."""

    def train(self, regularization_strength: float = 100.0, training_batch_size: int = 200, **kwargs):
        self._regularization_strength = float(regularization_strength)
        if getattr(self, "_dist", None) is not None and self._dp()[1] > 1:
            raise NotImplementedError("the balancing regulariser is a function of the batch mean: BNRE trains "
                                      "on one process")
        return NRE_A.train(self, training_batch_size=training_batch_size, **kwargs)

    def _loss_on(self, net, idx, num_atoms, choices=None, rows=None):
        from .multiround import bnre_loss
        return bnre_loss(self._logits_on(net, idx, 2, choices, rows), self._regularization_strength)


class NRE_C(NRE_B):
    """Contrastive NRE (reference: trainers/nre/nre_c.py:103-259): `num_classes` = K contrastive classes
    and the odds `gamma` of a jointly drawn pair; two independent contrastive draws per step (K + 1 and K
    atoms)."""

    def train(self, num_classes: int = 5, gamma: float = 1.0, training_batch_size: int = 200, **kwargs):
        self._gamma = float(gamma)
        return NRE_B.train(self, num_atoms=num_classes + 1, training_batch_size=training_batch_size, **kwargs)

    def _loss_on(self, net, idx, num_atoms, choices=None, rows=None):
        from .multiround import nre_c_loss
        K = num_atoms - 1
        if K < 1:
            raise AssertionError(f"num_classes = {K} must be greater than 1.")
        cm, cj = (choices if choices is not None else (None, None))
        logits_marginal = self._logits_on(net, idx, K + 1, cm, rows)
        logits_joint = self._logits_on(net, idx, K, cj, rows)
        return nre_c_loss(logits_marginal, logits_joint, self._gamma)


SNRE_C = NRE_C


# =================================================================================================
class FMPE(_FlowTrainer):
    """Flow-matching posterior estimation (reference: trainers/vfpe/fmpe.py, base_vf_inference.py:
    train :206-350, validation at fixed times :524-543, EMA-smoothed summaries :589-636, z-score of
    the loss as stopping rule :352-420)."""

    def __init__(self, prior=None, density_estimator: Union[str, Callable] = "mlp", device: str = "cuda",
                 logging_level: Union[int, str] = "WARNING", summary_writer=None, tracker=None,
                 show_progress_bars: bool = False):
        from .flowmatching import posterior_flow_nn
        self._prior = prior
        self._device = _process_device(device)
        self._build_neural_net = (posterior_flow_nn(model=density_estimator)
                                  if isinstance(density_estimator, str) else density_estimator)
        self._neural_net = None
        self._theta = self._x = None
        self.epoch, self._val_loss = 0, float("Inf")
        self._summary = dict(epochs_trained=[], best_validation_loss=[], validation_loss=[], training_loss=[],
                             epoch_durations_sec=[])
        self._dist = None

    def train(self, training_batch_size: int = 200, learning_rate: float = 5e-4, validation_fraction: float = 0.1,
              stop_after_epochs: int = 20, max_num_epochs: int = 2 ** 31 - 1, clip_max_norm: Optional[float] = 5.0,
              calibration_kernel=None, ema_loss_decay: float = 0.1, validation_times: Union[Tensor, int] = 10,
              validation_times_nugget: float = 0.05, resume_training: bool = False, **kwargs):
        if self._theta is None:
            raise RuntimeError("call append_simulations() first")
        self._vf_check_rounds(kwargs)
        lib = L.load()
        dev = self._device
        N = self._theta.shape[0]
        x2d = self._x.reshape(N, -1).contiguous()
        rank, world, part = self._dp()
        glob = world > 1 and part == "global"
        if glob:
            self._dp_agree(N, "number of simulations")
        n_train = int((1 - validation_fraction) * N)
        n_val = N - n_train
        if not resume_training or not hasattr(self, "train_indices"):
            self.train_indices, self.val_indices = self._dp_split(N, n_train)
        if self._neural_net is None:
            tr = self.train_indices.to(dev)
            self._neural_net = self._build_neural_net(self._theta[tr].cpu(), self._x[tr].cpu())
        net = self._neural_net.to(dev)
        self._neural_net = net
        if not resume_training:
            self._dp_sync_net(net)
        P, D = net.layout.n_params, net.layout.D
        B, Bv = min(training_batch_size, n_train), min(training_batch_size, n_val)
        steps, vsteps = n_train // B, (n_val // Bv if Bv > 0 else 0)
        self._dp_agree(steps, "number of steps per epoch")
        self._dp_agree(vsteps, "number of validation steps per epoch")
        if glob and (B % world or Bv % world):
            raise ValueError(f"partition='global' needs the batch sizes ({B}, {Bv}) divisible by the "
                             f"number of ranks ({world})")
        Bl, Bvl = (B // world, Bv // world) if glob else (B, Bv)    # rows of a batch this rank handles
        o_t, o_v = (rank * Bl, rank * Bvl) if glob else (0, 0)
        Btot = B if glob else B * world                              # rows behind one update
        if isinstance(validation_times, int):
            validation_times = torch.linspace(net.t_min + validation_times_nugget,
                                              net.t_max - validation_times_nugget, validation_times)
        vt = validation_times.to(dev).float()
        if not resume_training or not hasattr(self, "_opt_state"):
            self._opt_state = torch.zeros(2 * P, dtype=torch.float32, device=dev)
            self._opt_step = torch.zeros(2, dtype=torch.int32, device=dev)
            self.epoch, self._val_loss = 0, float("Inf")
        train_idx, val_idx = self.train_indices.to(dev), self.val_indices.to(dev)
        grad = torch.zeros(P, dtype=torch.float32, device=dev)
        loss_acc = torch.zeros(2, dtype=torch.float32, device=dev)
        max_norm = float(clip_max_norm) if clip_max_norm is not None else 0.0
        # data-parallel on one node: the gradient sum is our peer-memory kernel inside the epoch graph
        peer, grad_local = None, grad
        sumsq = None
        if world > 1:
            from .parallel import make_gradient_exchange
            peer = make_gradient_exchange(P)      # None -> NCCL all-reduce, eager launches
            if peer is not None:
                grad_local = torch.zeros(P, dtype=torch.float32, device=dev)
                sumsq = torch.zeros(peer.n_sumsq, dtype=torch.float32, device=dev)

        converged = lambda: self._vf_converged(net, stop_after_epochs)

        # One CUDA graph per epoch (single GPU): every step is [t ~ U, theta_1 ~ N draws, fused loss
        # fwd+bwd kernel, reduce, clip+Adam]; the epoch's row permutations live in static buffers.
        perm_buf = torch.zeros(max(steps * B, 1), dtype=torch.int64, device=dev)
        vperm_buf = torch.zeros(max(vsteps * Bv, 1), dtype=torch.int64, device=dev)
        stats = torch.zeros(4, dtype=torch.float32, device=dev)     # train loss sum, bad, val loss sum, bad

        def run_epoch():
            loss_acc.zero_()
            for s in range(steps):
                idx = perm_buf[s * B + o_t:s * B + o_t + Bl]
                tms = torch.rand(Bl, device=dev)
                eps = torch.randn(Bl, D, device=dev)
                _, gpart, n_part = net.loss_raw(self._theta, x2d, tms, eps, index=idx, g_const=1.0 / Btot,
                                                loss_acc=loss_acc, want_loss=False)
                L.check(lib.sbi_b200_reduce_partials(L.ptr(gpart), n_part, P, L.ptr(grad_local), L.stream_ptr()),
                        "reduce")
                if peer is not None:
                    peer.sum(grad_local, grad, net.net._mask, sumsq)
                    L.check(lib.sbi_b200_adam_clip_step_norm(
                        L.ptr(net.flat.data), L.ptr(grad), L.ptr(self._opt_state), L.ptr(self._opt_step),
                        L.ptr(net.net._mask), P, learning_rate, 0.9, 0.999, 1e-8, max_norm, 1.0, L.ptr(sumsq),
                        peer.n_sumsq, L.stream_ptr()), "adam")
                    continue
                if world > 1:
                    torch.distributed.all_reduce(grad)
                L.check(lib.sbi_b200_adam_clip_step(L.ptr(net.flat.data), L.ptr(grad), L.ptr(self._opt_state),
                                                    L.ptr(self._opt_step), L.ptr(net.net._mask), P, learning_rate,
                                                    0.9, 0.999, 1e-8, max_norm, 1.0, L.stream_ptr()), "adam")
            stats[0:2].copy_(loss_acc)
            # validation: every batch evaluated at all validation times (:524-543); g = 0 -> loss only
            loss_acc.zero_()
            if vsteps > 0:
                nt = vt.shape[0]
                for s in range(vsteps):
                    idx = vperm_buf[s * Bv + o_v:s * Bv + o_v + Bvl].repeat(nt).contiguous()
                    tms = vt.repeat_interleave(Bvl).contiguous()
                    eps = torch.randn(Bvl * nt, D, device=dev)
                    net.loss_raw(self._theta, x2d, tms, eps, index=idx, g_const=0.0, loss_acc=loss_acc, want_loss=False)
            stats[2:4].copy_(loss_acc)

        def fill_perms():
            perm_buf[:steps * B].copy_(train_idx[torch.randperm(n_train, device=dev)[:steps * B]])
            if vsteps > 0:
                vperm_buf[:vsteps * Bv].copy_(val_idx[torch.randperm(n_val, device=dev)[:vsteps * Bv]])
            if glob:      # one epoch order for all ranks: rank 0's
                torch.distributed.broadcast(perm_buf, 0)
                torch.distributed.broadcast(vperm_buf, 0)

        graph = None
        if (world == 1 or peer is not None) and os.environ.get("SBI_B200_FMPE_GRAPH", "1") != "0" and steps > 0:
            snap = (net.flat.data.clone(), self._opt_state.clone(), self._opt_step.clone())
            fill_perms()
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                run_epoch()
            torch.cuda.current_stream().wait_stream(side)
            net.flat.data.copy_(snap[0]); self._opt_state.copy_(snap[1]); self._opt_step.copy_(snap[2])
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                run_epoch()
            net.flat.data.copy_(snap[0]); self._opt_state.copy_(snap[1]); self._opt_step.copy_(snap[2])

        while self.epoch <= max_num_epochs and not converged():
            t0 = time.time()
            fill_perms()
            if graph is not None:
                graph.replay()
            else:
                run_epoch()
            tl, tb, vl, vb = self._dp_sum(stats.tolist())      # the one host sync of the epoch
            if tb > 0 or vb > 0:
                raise AssertionError("NaN/Inf present in FMPE loss.")
            train_loss = tl / (steps * Btot)
            val_loss = vl / (vsteps * Bv * (1 if glob else world) * vt.shape[0]) if vsteps > 0 else float("nan")
            # the reference normalises by len(loader) * loader.batch_size, i.e. WITHOUT the repeat over times
            val_loss *= vt.shape[0] if vsteps > 0 else 1.0
            self._vf_record_epoch(train_loss, val_loss, ema_loss_decay, t0)
        self._vf_finish(net, max_num_epochs)
        if peer is not None:
            timed_out = peer.error()
            peer.close()
            if timed_out:
                raise RuntimeError("peer-memory gradient exchange timed out (a rank fell behind or died)")
        return deepcopy(net)

    def _vf_check_rounds(self, kwargs):
        """base_vf_inference.py:451-496: only the first-round loss exists for vector-field trainers."""
        rounds = getattr(self, "_data_round_index", None) or [0]
        if max(rounds) > 0 and not kwargs.get("force_first_round_loss", False):
            raise NotImplementedError(
                f"Multi-round {self.__class__.__name__} with arbitrary proposals is not implemented")

    def _vf_converged(self, net, stop_after_epochs: int) -> bool:
        """base_vf_inference.py:352-420: an epoch counts as "no improvement" only if the validation loss sits more
        than two standard deviations (of the recent EMA-smoothed losses) above the best one."""
        if self.epoch == 0:
            self._best_val_loss, self._epochs_since_last_improvement, self._best_flat = float("inf"), 0, None
        if self._val_loss < self._best_val_loss:
            self._best_val_loss, self._epochs_since_last_improvement = self._val_loss, 0
            self._best_flat = net.flat.data.clone()
        else:
            if len(self._summary["validation_loss"]) >= stop_after_epochs:
                recent = torch.tensor(self._summary["validation_loss"][-stop_after_epochs * 2:])
                z = (self._val_loss - self._best_val_loss) / recent.std().item()
                self._epochs_since_last_improvement = self._epochs_since_last_improvement + 1 if z > 2.0 else 0
            else:
                return False
        if self._epochs_since_last_improvement > stop_after_epochs - 1:
            if self._best_flat is not None:
                net.flat.data.copy_(self._best_flat)
            return True
        return False

    def _vf_record_epoch(self, train_loss: float, val_loss: float, ema_loss_decay: float, t0: float):
        # base.py:1110 keeps the RAW validation loss in self._val_loss (what _converged compares
        # with the best loss); only the summaries hold the exponential moving averages
        # (base_vf_inference.py:589-636), whose spread normalises the stopping rule
        self._val_loss = val_loss
        if self._summary["training_loss"]:
            train_loss = (1 - ema_loss_decay) * self._summary["training_loss"][-1] + ema_loss_decay * train_loss
            val_loss = (1 - ema_loss_decay) * self._summary["validation_loss"][-1] + ema_loss_decay * val_loss
        self._summary["training_loss"].append(train_loss)
        self._summary["validation_loss"].append(val_loss)
        self._summary["epoch_durations_sec"].append(time.time() - t0)
        self.epoch += 1

    def _vf_finish(self, net, max_num_epochs: int):
        if self.epoch > max_num_epochs:
            if self._val_loss < self._best_val_loss:
                self._best_val_loss, self._best_flat = self._val_loss, net.flat.data.clone()
            elif self._best_flat is not None:
                net.flat.data.copy_(self._best_flat)
        self._summary["epochs_trained"].append(self.epoch)
        self._summary["best_validation_loss"].append(self._best_val_loss)

    def build_posterior(self, density_estimator=None, prior=None, sample_with: str = "ode", **kwargs):
        from .posteriors import VectorFieldPosterior
        if sample_with not in ("ode", "sde"):
            raise ValueError(f"sample_with must be 'ode' or 'sde', but is {sample_with}.")
        est = deepcopy(density_estimator if density_estimator is not None else self._neural_net)
        return VectorFieldPosterior(est, prior if prior is not None else self._prior, device=self._device,
                                    sample_with=sample_with)



class NPSE(FMPE):
    """Neural posterior score estimation (reference: trainers/vfpe/npse.py:69-268 on the shared loop of
    base_vf_inference.py): denoising score matching of a VE / VP / sub-VP score network (score.py).

    A step is [times + noise draws, noising and target arithmetic in torch, ONE launch of the network kernel over
    the noised inputs and the control-variate means, loss head in torch, ONE launch of the network's
    parameter-gradient kernel, clip + Adam kernel], captured once as a CUDA graph and replayed per batch; the
    validation step (all validation times at once, base_vf_inference.py:524-543) is a second graph."""

    def __init__(self, prior=None, vf_estimator: Union[str, Callable, None] = None,
                 score_estimator: Union[str, Callable, None] = None, density_estimator: Optional[Callable] = None,
                 sde_type: Optional[str] = None, device: str = "cuda", logging_level: Union[int, str] = "WARNING",
                 summary_writer=None, tracker=None, show_progress_bars: bool = False):
        from .score import posterior_score_nn
        given = [e for e in (vf_estimator, score_estimator, density_estimator) if e is not None]
        if len(given) > 1:
            raise ValueError("pass only one of vf_estimator / score_estimator / density_estimator")
        est = given[0] if given else "mlp"
        super().__init__(prior=prior, density_estimator=(lambda *a: None), device=device)
        if isinstance(est, str):
            self._build_neural_net = posterior_score_nn(model=est, sde_type=sde_type or "ve")
        else:
            if sde_type is not None:
                warnings.warn("sde_type is ignored when a build function is passed", stacklevel=2)
            self._build_neural_net = est

    def train(self, training_batch_size: int = 200, learning_rate: float = 5e-4, validation_fraction: float = 0.1,
              stop_after_epochs: int = 20, max_num_epochs: int = 2 ** 31 - 1, clip_max_norm: Optional[float] = 5.0,
              calibration_kernel=None, ema_loss_decay: float = 0.1, validation_times: Union[Tensor, int] = 10,
              validation_times_nugget: float = 0.05, resume_training: bool = False, **kwargs):
        if self._theta is None:
            raise RuntimeError("call append_simulations() first")
        self._vf_check_rounds(kwargs)
        if self._dp()[1] > 1:
            raise NotImplementedError("NPSE training is single-process")
        lib = L.load()
        dev = self._device
        N = self._theta.shape[0]
        x2d = self._x.reshape(N, -1).contiguous()
        n_train = int((1 - validation_fraction) * N)
        n_val = N - n_train
        if not resume_training or not hasattr(self, "train_indices"):
            self.train_indices, self.val_indices = self._dp_split(N, n_train)
        if self._neural_net is None:
            tr = self.train_indices.to(dev)
            self._neural_net = self._build_neural_net(self._theta[tr].cpu(), self._x[tr].cpu())
        net = self._neural_net.to(dev)
        self._neural_net = net
        P = net.layout.n_params
        B, Bv = min(training_batch_size, n_train), min(training_batch_size, n_val)
        steps, vsteps = n_train // B, (n_val // Bv if Bv > 0 else 0)
        if isinstance(validation_times, int):
            validation_times = torch.linspace(net.t_min + validation_times_nugget,
                                              net.t_max - validation_times_nugget, validation_times)
        vt = validation_times.to(dev).float()
        nt = vt.shape[0]
        if not resume_training or not hasattr(self, "_opt_state"):
            self._opt_state = torch.zeros(2 * P, dtype=torch.float32, device=dev)
            self._opt_step = torch.zeros(2, dtype=torch.int32, device=dev)
            self.epoch, self._val_loss = 0, float("Inf")
        train_idx, val_idx = self.train_indices.to(dev), self.val_indices.to(dev)
        max_norm = float(clip_max_norm) if clip_max_norm is not None else 0.0
        w_all = None
        if calibration_kernel is not None:
            w_all = torch.as_tensor(calibration_kernel(self._x), dtype=torch.float32).reshape(-1).to(dev)
        idx_buf = torch.zeros(B, dtype=torch.int64, device=dev)
        vidx_buf = torch.zeros(max(Bv, 1), dtype=torch.int64, device=dev)
        stats = torch.zeros(4, dtype=torch.float32, device=dev)     # train loss sum, bad, val loss sum, bad

        def losses_on(idx, times):
            losses = net.loss(self._theta[idx], x2d[idx], times=times)
            return losses if w_all is None else w_all[idx] * losses

        def train_step():
            net.net.flat.grad = None
            losses = losses_on(idx_buf, None)
            losses.mean().backward()
            ld = losses.detach()
            stats[0] += ld.sum()
            stats[1] += (~torch.isfinite(ld)).sum()
            L.check(lib.sbi_b200_adam_clip_step(
                L.ptr(net.flat.data), L.ptr(net.flat.grad), L.ptr(self._opt_state), L.ptr(self._opt_step),
                L.ptr(net.net._mask), P, learning_rate, 0.9, 0.999, 1e-8, max_norm, 1.0, L.stream_ptr()),
                "adam_clip_step")

        def val_step():      # the batch repeated over all validation times (base_vf_inference.py:524-543)
            with torch.no_grad():
                ld = losses_on(vidx_buf.repeat(nt), vt.repeat_interleave(Bv))
                stats[2] += ld.sum()
                stats[3] += (~torch.isfinite(ld)).sum()

        g_train = g_val = None
        if os.environ.get("SBI_B200_NPSE_GRAPH", "1") != "0" and steps > 0:
            snap = (net.flat.data.clone(), self._opt_state.clone(), self._opt_step.clone())
            idx_buf.copy_(train_idx[:B])
            if vsteps > 0:
                vidx_buf.copy_(val_idx[:Bv])
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):          # warm-up outside capture (allocations, kernel attributes)
                for _ in range(3):
                    train_step()
                if vsteps > 0:
                    val_step()
            torch.cuda.current_stream().wait_stream(side)
            g_train = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g_train):
                train_step()
            if vsteps > 0:
                g_val = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g_val):
                    val_step()
            net.flat.data.copy_(snap[0]); self._opt_state.copy_(snap[1]); self._opt_step.copy_(snap[2])

        while self.epoch <= max_num_epochs and not self._vf_converged(net, stop_after_epochs):
            t0 = time.time()
            perm = train_idx[torch.randperm(n_train, device=dev)]
            vperm = val_idx[torch.randperm(n_val, device=dev)] if vsteps > 0 else None
            stats.zero_()
            for s_ in range(steps):
                idx_buf.copy_(perm[s_ * B:(s_ + 1) * B])
                g_train.replay() if g_train is not None else train_step()
            for s_ in range(vsteps):
                vidx_buf.copy_(vperm[s_ * Bv:(s_ + 1) * Bv])
                g_val.replay() if g_val is not None else val_step()
            tl, tb, vl, vb = stats.tolist()                    # the one host sync of the epoch
            if tb > 0 or vb > 0:
                raise AssertionError("NaN/Inf present in NPSE loss.")
            # the reference normalises the validation sum by len(loader) * batch_size, i.e. WITHOUT the repeat over times
            self._vf_record_epoch(tl / (steps * B), vl / (vsteps * Bv) if vsteps > 0 else float("nan"),
                                  ema_loss_decay, t0)
        self._vf_finish(net, max_num_epochs)
        net.zero_grad(set_to_none=True)
        net._cache.clear()
        return deepcopy(net)

    def build_posterior(self, vector_field_estimator=None, prior=None, sample_with: str = "sde", **kwargs):
        """npse.py:219-261: same posterior as FMPE's, reverse-SDE sampling by default."""
        return super().build_posterior(density_estimator=vector_field_estimator, prior=prior, sample_with=sample_with,
                                       **kwargs)
