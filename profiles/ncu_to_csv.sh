#!/bin/bash
# On the GPU box: turn an .ncu-rep into the small CSV pages that travel back (the reports themselves
# exceed gpurun's 64 MiB return limit):  profiles/ncu_to_csv.sh gpurun_out/x.ncu-rep
set -e
rep="$1"; base="${rep%.ncu-rep}"
ncu -i "$rep" --page raw --csv > "${base}_raw.csv"
if [ "${2:-}" = "source" ]; then ncu -i "$rep" --page source --csv > "${base}_source.csv" || true; fi
sz=$(stat -c %s "$rep")
if [ "$sz" -gt 20000000 ]; then rm -f "$rep"; fi
