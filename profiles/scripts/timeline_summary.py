"""Summary of a torch.profiler chrome trace written by timeline.py: per-stream busy fraction, host turnaround gaps
between a group's steps, kernel duration statistics.   python profiles/scripts/timeline_summary.py gpurun_out/timeline_8.json"""
import json
import sys
from collections import defaultdict

ev = [e for e in json.load(open(sys.argv[1]))["traceEvents"] if e.get("cat") == "kernel"]
ev.sort(key=lambda e: e["ts"])
t0 = ev[0]["ts"]
by = defaultdict(list)
for e in ev:
    by[e["args"]["stream"]].append(e)
span = max(e["ts"] + e["dur"] for e in ev) - t0
short = lambda n: n.split("::")[-1].split("(")[0]
print(f"span {span:.0f} us, {len(by)} streams, {len(ev)} kernels")
gaps, chains = [], []
for s, L in sorted(by.items()):
    names = sorted(set(short(x["name"]) for x in L))
    busy = sum(x["dur"] for x in L)
    kind = "track" if any("seq_post" in n for n in names) else "build"
    print(f"  stream {s} ({kind}): busy {100 * busy / span:.0f}%")
    if kind == "track":
        start = None
        for a, b in zip(L, L[1:] + [None]):
            if "seq_align" in a["name"] or "seq_prep" in a["name"]:
                start = a["ts"]
            if "seq_post" in a["name"]:
                if start is not None:
                    chains.append(a["ts"] + a["dur"] - start)
                if b is not None:
                    gaps.append(b["ts"] - (a["ts"] + a["dur"]))
med = lambda v: sorted(v)[len(v) // 2] if v else float("nan")
print(f"tracking chain align..post: median {med(chains):.0f} us; gap (post end -> next chain start; negative = overlapped by programmatic dependent launch): median {med(gaps):.0f} us, max {max(gaps):.0f} us")
d = defaultdict(list)
for e in ev:
    d[short(e["name"])].append(e["dur"])
for k, v in d.items():
    print(f"  {k:22s} n={len(v):4d} avg {sum(v) / len(v):7.1f} us  total {sum(v):8.0f}")

# per tracking stream: period between consecutive chain starts, and what the stream waited for before each chain
periods = []
for s_, L in sorted(by.items()):
    starts = [x["ts"] for x in L if "seq_align" in x["name"]]
    periods += [b - a for a, b in zip(starts, starts[1:])]
if periods:
    periods.sort()
    print(f"chain period (align start -> next align start): median {med(periods):.0f} us, p10 {periods[len(periods)//10]:.0f}, p90 {periods[len(periods)*9//10]:.0f}, n={len(periods)}")
# build chain: upload start -> select end per build stream
bch = []
for s_, L in sorted(by.items()):
    st = None
    for x in L:
        n = short(x["name"])
        if n.startswith("upload"):
            st = x["ts"]
        if n.startswith("fast_select") and st is not None:
            bch.append(x["ts"] + x["dur"] - st)
            st = None
if bch:
    bch.sort()
    print(f"build chain (upload start -> select end): median {med(bch):.0f} us, p90 {bch[len(bch)*9//10]:.0f}, n={len(bch)}")
# GPU-wide: time with at least one kernel running
def union(evs):
    tot, end = 0.0, None
    for e in sorted(evs, key=lambda e: e["ts"]):
        a, b = e["ts"], e["ts"] + e["dur"]
        if end is None or a > end:
            tot += b - a; end = b
        elif b > end:
            tot += b - end; end = b
    return tot
print(f"any kernel running {100*union(ev)/span:.0f}% of the span; sum of durations / span = {sum(e['dur'] for e in ev)/span:.1f} kernels in flight")
