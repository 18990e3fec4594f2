#!/bin/bash
# ncu capture of the flow-matching training kernel after the RED gradient write-out (cfg4 shape: D = C = 20, 16384 rows)
cd "$GRAFT_REPO_ROOT"
timeout 300 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:fm_vjp \
  -o gpurun_out/r02_fm_vjp python profiles/prof_all.py fm > gpurun_out/r02_fm_vjp_ncu.log 2>&1
tail -3 gpurun_out/r02_fm_vjp_ncu.log
ncu -i gpurun_out/r02_fm_vjp.ncu-rep --page raw --csv > gpurun_out/r02_fm_vjp_raw.csv
ncu -i gpurun_out/r02_fm_vjp.ncu-rep --page source --print-source cuda,sass --csv > gpurun_out/r02_fm_vjp_source.csv
ls -la gpurun_out/
rm -f gpurun_out/r02_fm_vjp.ncu-rep
find gpurun_out -size +30M -delete
