"""ncu `--page raw --csv` (one row per profiled launch) -> markdown table, one row per distinct kernel.

    python profiles/summarize_raw.py gpurun_out/r2_prof_all_raw.csv "title" > profiles/r02_kernels.md
"""
import csv
import re
import sys

src, title = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "ncu summary")
rows = list(csv.reader(open(src)))
hdr, units = rows[0], rows[1]
ix = {h: i for i, h in enumerate(hdr)}
COLS = [
    ("gpu__time_duration.sum", "time"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
    ("launch__registers_per_thread", "regs"), ("launch__shared_mem_per_block_dynamic", "dyn smem"),
    ("dram__bytes_read.sum", "DRAM rd"), ("dram__bytes_write.sum", "DRAM wr"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps act %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "FMA pipe %"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe %"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts"),
]
STALLS = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
seen = {}
for r in rows[2:]:
    name = r[ix["Kernel Name"]]
    if not re.search(r"sbi::|tc::|nsf_|maf_|fm_|ratio_|slice_|ode|sde|adam|reduce_partials|peer_", name):
        continue
    short = re.sub(r"\(.*", "", name).replace("void ", "")
    key = (short, r[ix["launch__grid_size"]])
    seen.setdefault(key, r)
print(f"# {title}\n\nsource: `{src.split('/')[-1]}` (`ncu --set full --clock-control none`, one launch per kernel inside a "
      "cudaProfiler range after warm-up; times are cold-cache single launches)\n")
print("| kernel | " + " | ".join(c[1] for c in COLS) + " | top stalls (warps per issue) |")
print("|---|" + "---|" * (len(COLS) + 1))
for (short, _), r in seen.items():
    cells = []
    for h, _lab in COLS:
        if h in ix:
            v, u = r[ix[h]], units[ix[h]]
            try:
                f = float(v.replace(",", ""))
                v = f"{f:.3g}" if abs(f) < 1000 else f"{f:,.0f}"
            except ValueError:
                pass
            cells.append(f"{v} {u}".strip().replace("register/thread", "").replace("Kbyte/block", "KB").replace("Mbyte", "MB"))
        else:
            cells.append("-")
    st = []
    for h in STALLS:
        try:
            st.append((float(r[ix[h]]), h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]))
        except ValueError:
            pass
    st.sort(reverse=True)
    print(f"| `{short}` | " + " | ".join(cells) + " | " + ", ".join(f"{n} {v:.2f}" for v, n in st[:3]) + " |")
