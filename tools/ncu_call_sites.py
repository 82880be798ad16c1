"""Per CALL SITE in the kernel's own source file: share of executed warp instructions / stall samples and mean active
threads, with everything inlined below a call site folded into it (ncu source CSV joined with `nvdisasm -gi`).

    python tools/ncu_call_sites.py <report.ncu-rep> <kernel-substring[#k]> <cubin> <mangled-substring> <outer-file> [top]
"""
import csv, os, re, subprocess, sys
from collections import defaultdict

rep, kern, cubin, mangled_sub, outer = sys.argv[1:6]
top = int(sys.argv[6]) if len(sys.argv) > 6 else 30
kern, _, nth = kern.partition("#")
nth = int(nth) if nth else 0
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
start = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name" and kern in r[1]][nth]
head = rows[start + 1]
body = []
for r in rows[start + 2:]:
    if not r or r[0] == "Kernel Name":
        break
    body.append(r)
ia, isamp, iinst, ithr = (head.index(k) for k in ("Address", "# Samples", "Instructions Executed", "Thread Instructions Executed"))
base = int(body[0][ia], 16)
dis = subprocess.run(["nvdisasm", "-gi", "-c", cubin], capture_output=True, text=True).stdout
site_of, cur_fn, chain, want = {}, None, [], False
pending = []
for l in dis.splitlines():
    m = re.match(r"\s*\.text\.(\S+):", l)
    if m:
        cur_fn = m.group(1)
        want = mangled_sub in cur_fn
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?', l)
    if m:
        pending.append((os.path.basename(m.group(1)), int(m.group(2)), os.path.basename(m.group(3)) if m.group(3) else None,
                        int(m.group(4)) if m.group(4) else None))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+\S", l)
    if m and want:
        if pending:
            chain = pending
        pending = []
        # outermost frame in `outer`: the last "inlined at" whose file is outer, else the line itself
        site = None
        for f, ln, pf, pl in chain:
            if pf == outer:
                site = (pf, pl)
            elif pf is None and f == outer:
                site = (f, ln)
        if site is None and chain:
            site = (chain[-1][2] or chain[-1][0], chain[-1][3] or chain[-1][1])
        site_of[int(m.group(1), 16)] = site
agg = defaultdict(lambda: [0.0, 0.0, 0.0])
for r in body:
    site = site_of.get(int(r[ia], 16) - base)
    a = agg[site]
    a[0] += float(r[isamp] or 0); a[1] += float(r[iinst] or 0); a[2] += float(r[ithr] or 0)
ts, ti = sum(a[0] for a in agg.values()), sum(a[1] for a in agg.values())
print(f"{'call site':28s} samples%   inst% thr/inst")
for site, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    name = f"{site[0]}:{site[1]}" if site else "?"
    print(f"{name:28s} {100*a[0]/ts:8.2f} {100*a[1]/ti:7.2f} {a[2]/max(a[1],1):8.1f}")
