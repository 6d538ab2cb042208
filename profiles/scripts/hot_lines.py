#!/usr/bin/env python
"""Hot CUDA source lines of one kernel from an ncu report captured with --import-source on.
  python profiles/scripts/hot_lines.py gpurun_out/prof.ncu-rep seq_post [top]"""
import csv
import io
import subprocess
import sys

rep, pat = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "-k", f"regex:{pat}"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
fname, hdr, data, seen_kernel = "", None, {}, 0
first_file = None
for r in rows:
    if len(r) >= 2 and r[0] == "Kernel Name":
        seen_kernel += 1
        if seen_kernel > 1:
            break
        continue
    if len(r) >= 2 and r[0] in ("File Name", "File Path"):
        fname = r[1].split("/")[-1]
        if first_file is None:
            first_file = fname
        elif fname == first_file:
            break   # the next captured launch starts over with the first file
        continue
    if r and r[0] == "Line No":
        hdr = r
        continue
    if hdr is None or len(r) < len(hdr) - 2 or not r[0]:
        continue   # SASS rows have an empty first column
    try:
        w = int(r[hdr.index("# Samples")]) if "# Samples" in hdr else 0
        n = int(r[hdr.index("Instructions Executed")])
    except ValueError:
        continue
    stalls = {h: int(v) for h, v in zip(hdr, r) if h.startswith("stall_") and "Not Issued" not in h and v.isdigit() and int(v) > 0}
    key = (fname, r[0])
    data[key] = (w, n, r[1].strip()[:100], stalls)
tot = sum(v[0] for v in data.values()) or 1
print(f"# {pat}: {tot} samples")
for (f, ln), (w, n, src, st) in sorted(data.items(), key=lambda kv: -kv[1][0])[:top]:
    top_st = ",".join(f"{k[6:]}:{v}" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:3])
    print(f"{100 * w / tot:5.1f}% {n:9d} inst  {f}:{ln:<5s} {src}   [{top_st}]")
