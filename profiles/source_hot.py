"""Top source lines by warp-stall samples from `ncu -i X.ncu-rep --page source --print-source cuda,sass --csv`.
    python profiles/source_hot.py file.csv [top_n]"""
import csv, sys, collections
csv.field_size_limit(10 ** 9)
path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
kernel, fpath, hdr = None, None, None
per = collections.OrderedDict()
for r in csv.reader(open(path)):
    if not r:
        continue
    if r[0] in ("Kernel Name", "Function Name"):
        kernel = r[1].split("(")[0][-40:]
        per.setdefault(kernel, [])
        continue
    if r[0] == "File Path":
        fpath = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = r
        continue
    if hdr is None or kernel is None or not r[0].strip().isdigit():
        continue
    d = dict(zip(hdr[4:], r[4:]))
    try:
        n = int(d.get("# Samples") or 0)
    except ValueError:
        continue
    if n:
        stalls = {k[6:]: int(v or 0) for k, v in d.items() if k.startswith("stall_") and "Not Issued" not in k and v not in ("", "0")}
        per[kernel].append((n, fpath, int(r[0]), r[1].strip()[:110], stalls, int(d.get("Instructions Executed") or 0)))
for k, rows in per.items():
    tot = sum(x[0] for x in rows)
    print(f"\n## {k}: {tot} samples")
    for n, f, ln, src, st, ie in sorted(rows, key=lambda t: -t[0])[:top]:
        s = ", ".join(f"{a} {b}" for a, b in sorted(st.items(), key=lambda t: -t[1])[:3])
        print(f"{100 * n / tot:5.1f}%  {f}:{ln:<4d} inst {ie:<7d} [{s}]  {src}")
