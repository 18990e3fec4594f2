#!/bin/bash
# Round-2 measurement run on one B200 (gpurun): GPU test suite, the bench lines of every BASELINE config,
# the reference arm, ncu captures of the training pair.  Outputs land in gpurun_out/r02_*; the ones worth
# keeping are copied to profiles/ afterwards.
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
nvidia-smi -L > $O/r02_smi.log
nproc > $O/r02_nproc.log
timeout 900 python -m pytest tests -m gpu -q > $O/r02_pytest.log 2>&1; echo "rc=$?" >> $O/r02_pytest.log
timeout 600 python bench.py > $O/r02_bench_n1.json 2> $O/r02_bench_n1.err
timeout 400 python bench.py --impl reference --steps 42 --warmup 21 > $O/r02_bench_n1_reference.json 2> $O/r02_bench_ref.err
for w in cfg3 cfg4 cfg5; do
  timeout 400 python bench.py --workload $w --steps 60 > $O/r02_bench_$w.json 2> $O/r02_bench_$w.err
done
# ncu: launch list of smoke(), then full captures of the tensor-core training pair
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r02_smoke_launches.csv \
    python -c "import __graft_entry__ as g; g.smoke()" > $O/r02_smoke.log 2>&1
ncu --set full --clock-control none --import-source on --profile-from-start off \
    -k regex:"nsf_vjp_tc|nsf_logprob_tc" -o $O/r02_vjp_tc python profiles/vjp_tc_prof.py 4096 > $O/r02_ncu_pair.log 2>&1
ncu -i $O/r02_vjp_tc.ncu-rep --page raw --csv > $O/r02_vjp_tc_raw.csv
ncu -i $O/r02_vjp_tc.ncu-rep --page source --print-source cuda,sass --csv > $O/r02_vjp_tc_source.csv 2>/dev/null
python profiles/source_hot.py $O/r02_vjp_tc_source.csv 40 > $O/r02_vjp_tc_hot.md 2>&1
rm -f $O/r02_vjp_tc_source.csv
tail -3 $O/r02_pytest.log
python - <<PY
import json
for f in ("r02_bench_n1","r02_bench_n1_reference","r02_bench_cfg3","r02_bench_cfg4","r02_bench_cfg5"):
    try:
        l=json.loads(open("$O/%s.json"%f).read().strip().splitlines()[-1])
        print(f, l.get("value"), l.get("unit"), l.get("ms_per_step"), (l.get("e2e") or {}).get("value"))
    except Exception as e:
        print(f, "FAILED", e)
PY
