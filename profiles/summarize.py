"""Turn an ncu report into the markdown summary committed under profiles/.

    python profiles/summarize.py gpurun_out/prof_vjp.ncu-rep nsf_vjp_kernel > profiles/r01_nsf_vjp.md
"""
import collections
import csv
import os
import re
import subprocess
import sys
import tempfile

rep, kern = sys.argv[1], sys.argv[2]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "sbi_b200", "lib", "libsbi_b200.so")


def run(cmd):
    return subprocess.run(cmd, capture_output=True, text=True).stdout


raw = list(csv.reader(run(["ncu", "-i", rep, "--page", "raw", "--csv"]).splitlines()))
hdr, units, vals = raw[0], raw[1], raw[2]
want = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed.sum",
]
print(f"# ncu summary: `{kern}`\n\nreport: `{os.path.basename(rep)}` (`ncu --set full --clock-control none "
      f"--import-source on`)\n\n| metric | unit | value |\n|---|---|---|")
for w in want:
    if w in hdr:
        i = hdr.index(w)
        print(f"| {w} | {units[i]} | {vals[i]} |")

sass = list(csv.reader(run(["ncu", "-i", rep, "--page", "source", "--csv"]).splitlines()))
h2, data = sass[1], sass[2:]
ix = {h: i for i, h in enumerate(h2)}
stalls = [h for h in h2 if h.startswith("stall_") and "Not Issued" not in h]
tot = collections.Counter()
for r in data:
    for s in stalls:
        try:
            tot[s] += int(r[ix[s]])
        except Exception:
            pass
T = sum(tot.values())
print(f"\n## warp stall sampling ({T} samples)\n\n| reason | share |\n|---|---|")
for s, v in tot.most_common(8):
    print(f"| {s} | {100 * v / T:.1f}% |")

# map SASS offsets to source lines with nvdisasm -g
with tempfile.TemporaryDirectory() as td:
    subprocess.run(["cuobjdump", "-xelf", "all", LIB], cwd=td, capture_output=True)
    off2line = {}
    for f in os.listdir(td):
        if not f.endswith(".cubin"):
            continue
        asm = run(["nvdisasm", "-g", "-c", os.path.join(td, f)])
        cur_fn = cur = None
        for ln in asm.splitlines():
            m = re.match(r"\s*\.text\.(\S+):", ln)
            if m:
                cur_fn = m.group(1)
                continue
            m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
            if m:
                cur = (m.group(1).split("/")[-1], int(m.group(2)))
                continue
            m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
            if m and cur_fn and kern in cur_fn and cur_fn == (sys.argv[3] if len(sys.argv) > 3 else cur_fn):
                off2line.setdefault(cur_fn, {})[int(m.group(1), 16)] = cur
    # pick the function whose instruction count matches the report best
    fn = min(off2line, key=lambda k: abs(len(off2line[k]) - len(data))) if off2line else None
    base = int(data[0][ix["Address"]], 16)
    agg, aggb = collections.Counter(), collections.Counter()
    for r in data:
        key = off2line.get(fn, {}).get(int(r[ix["Address"]], 16) - base, ("?", 0))
        agg[key] += int(r[ix["# Samples"]] or 0)
        aggb[key] += int(r[ix["stall_barrier"]] or 0)
    print(f"\n## hottest source lines (function `{fn}`)\n\n| samples | share | of which barrier | line |\n|---|---|---|---|")
    for k, v in agg.most_common(14):
        print(f"| {v} | {100 * v / max(T, 1):.1f}% | {aggb[k]} | {k[0]}:{k[1]} |")
