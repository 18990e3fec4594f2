#!/bin/bash
# deeper weight ring in the flow-matching kernels: parity of everything that runs on them + cfg4 A/B
cd "$GRAFT_REPO_ROOT"
timeout 300 python -m pytest tests/test_fm_gpu.py tests/test_score_gpu.py tests/test_ode_gpu.py tests/test_dropin_gpu.py -q -m gpu > gpurun_out/r02_pytest_fm_c.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02_pytest_fm_c.log
timeout 200 python bench.py --workload cfg4 --steps 162 > gpurun_out/r02_bench_cfg4_c.json 2> gpurun_out/cfg4_c.err; cut -c1-250 gpurun_out/r02_bench_cfg4_c.json
SBI_B200_LIB=sbi_b200/lib/libsbi_b200_noring.so timeout 200 python bench.py --workload cfg4 --steps 162 --no-cpu-baseline > gpurun_out/r02_bench_cfg4_noring.json 2> gpurun_out/cfg4_noring.err; cut -c1-250 gpurun_out/r02_bench_cfg4_noring.json
timeout 120 python profiles/npse_time.py > gpurun_out/r02_npse_time_c.log 2>&1; tail -4 gpurun_out/r02_npse_time_c.log
