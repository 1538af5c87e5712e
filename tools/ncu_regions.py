#!/usr/bin/env python3
"""Per-region cost of a kernel from an .ncu-rep source page: SASS lines are cut into runs of (nearly) equal execution
count; prints instructions / stall samples per run with the dominant opcodes and stall reasons."""
import csv, io, subprocess, re, sys
rep = sys.argv[1]
thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.004
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]; data = rows[2:]
iS = hdr.index('Source'); iE = hdr.index('Instructions Executed'); iN = hdr.index('# Samples')
stall_cols = [i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
E = [int(r[iE]) for r in data]; S = [int(r[iN]) for r in data]; T = [r[iS].strip() for r in data]
tot = sum(E); totS = sum(S)
print("total warp-instr %.4g, samples %d, SASS lines %d" % (tot, totS, len(data)))
seg = []; start = 0
for k in range(1, len(E) + 1):
    if k == len(E) or abs(E[k] - E[start]) > 0.02 * max(E[start], 1) + 2000:
        seg.append((start, k)); start = k
for a, b in seg:
    e = sum(E[a:b]); s = sum(S[a:b])
    if e > tot * thr or s > totS * thr:
        ops = {}
        for t in T[a:b]:
            m = re.match(r'(@!?U?P\d+\s+)?([A-Z0-9_.]+)', t)
            if m:
                o = m.group(2).split('.')[0]; ops[o] = ops.get(o, 0) + 1
        st = {}
        for k in range(a, b):
            for i in stall_cols:
                v = data[k][i]
                if v not in ('', '0'):
                    st[hdr[i][6:]] = st.get(hdr[i][6:], 0) + int(v)
        top = " ".join("%s:%d" % x for x in sorted(ops.items(), key=lambda x: -x[1])[:5])
        sts = " ".join("%s:%.1f%%" % (n, 100 * v / totS) for n, v in sorted(st.items(), key=lambda x: -x[1])[:4])
        print("%4d-%4d n=%4d exec %.3fM  instr %5.1f%%  samples %5.1f%%  | %s | %s" % (a, b, b - a, E[a] / 1e6, 100 * e / tot, 100 * s / totS, top, sts))
