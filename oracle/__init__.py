"""oracle/ -- TEST INFRASTRUCTURE ONLY (CPU restatement of the reference hot path).

Nothing under ``oracle/`` is product code.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may import it, and
only as the checker / the timed CPU baseline -- never as the thing shipped.

Contents
--------
``nflows_port``  plain-PyTorch (CPU, fp32/fp64) restatement of the pieces of
                 ``nflows==0.14`` that sbi's hot path calls.
                 nflows is a third-party dependency that is NOT vendored in
                 /root/reference and is not installable offline, so its published
                 algorithm is restated here from the paper (Durkan et al. 2019,
                 "Neural Spline Flows"; Papamakarios et al. 2017, "MAF";
                 Germain et al. 2015, "MADE") and from the library's documented
                 behaviour.  **Parity status: unpinned at the level of nflows
                 numerics** -- no golden vectors for spline / MADE / ResidualNet exist
                 in the reference's tests; the port is instead pinned by mathematical
                 self-checks (invertibility, log|det| vs autograd Jacobian, density
                 normalisation by quadrature) in ``tests/test_oracle_*.py`` and by the
                 one known-answer test the reference holds for this path
                 (``tests/torchutils_test.py:135-157``, bin search).
``ref_shim``     registers ``nflows_port`` under the module name ``nflows`` plus inert
                 stubs for zuko/pyro/matplotlib/skorch/pymc so that the UNMODIFIED
                 reference ``sbi`` in /root/reference imports in the build container;
                 used only by ``tests/golden/make_golden.py`` to generate fixtures.
``sbi_port``     CPU restatement of the sbi-owned part of the path (builders,
                 z-scoring, estimator wrapper, training loop, samplers) used where
                 /root/reference is not available (the GPU box): the parity checker in
                 ``-m gpu`` tests and the timed ``cpu_baseline`` / ``--impl reference``.
"""
