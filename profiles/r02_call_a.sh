#!/bin/bash
# round 2, late: full GPU suite on the RED gradient write-out + loss-only FMPE validation, cfg4 A/B, NPSE timing
cd "$GRAFT_REPO_ROOT"
timeout 420 python -m pytest tests -q -m gpu -x > gpurun_out/r02_pytest_gpu_b.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r02_pytest_gpu_b.log
timeout 200 python bench.py --workload cfg4 --steps 162 > gpurun_out/r02_bench_cfg4_b.json 2> gpurun_out/cfg4_b.err; cat gpurun_out/r02_bench_cfg4_b.json | cut -c1-260
SBI_B200_LIB=sbi_b200/lib/libsbi_b200_rmw.so timeout 200 python bench.py --workload cfg4 --steps 162 --no-cpu-baseline > gpurun_out/r02_bench_cfg4_rmw.json 2> gpurun_out/cfg4_rmw.err; cat gpurun_out/r02_bench_cfg4_rmw.json | cut -c1-260
timeout 120 python profiles/npse_time.py > gpurun_out/r02_npse_time.log 2>&1; cat gpurun_out/r02_npse_time.log | tail -5
