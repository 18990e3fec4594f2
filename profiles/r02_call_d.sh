#!/bin/bash
# flow-matching kernels: one output row per thread (rn1) and per-kernel re-chunking of the weight stream (tuned)
cd "$GRAFT_REPO_ROOT"
for v in rn1 tuned; do
  export SBI_B200_LIB=sbi_b200/lib/libsbi_b200_$v.so
  timeout 200 python -m pytest tests/test_fm_gpu.py tests/test_score_gpu.py tests/test_ode_gpu.py tests/test_dropin_gpu.py -q -m gpu > gpurun_out/r02_pytest_fm_$v.log 2>&1; echo "$v pytest rc=$?"; tail -2 gpurun_out/r02_pytest_fm_$v.log
  timeout 150 python bench.py --workload cfg4 --steps 162 --no-cpu-baseline > gpurun_out/r02_bench_cfg4_$v.json 2> gpurun_out/cfg4_$v.err; cut -c1-200 gpurun_out/r02_bench_cfg4_$v.json
done
timeout 100 python profiles/npse_time.py > gpurun_out/r02_npse_time_tuned.log 2>&1; tail -4 gpurun_out/r02_npse_time_tuned.log
