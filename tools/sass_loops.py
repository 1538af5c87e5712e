#!/usr/bin/env python3
"""Opcode histogram of a kernel's SASS and of its innermost loops (backward branches): tools/sass_loops.py <lib.so> <kernel-substring>"""
import re, subprocess, sys, collections
lib, pat = sys.argv[1], sys.argv[2]
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
blocks = out.split("Function : ")
for b in blocks[1:]:
    name = b.split("\n", 1)[0]
    if pat not in name:
        continue
    ins = []
    for l in b.split("\n"):
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
        if m:
            ins.append((int(m.group(1), 16), m.group(2)))
    print("==", name, len(ins), "instructions")
    addr_index = {a: i for i, (a, _) in enumerate(ins)}
    hist = collections.Counter(re.sub(r"^@!?U?P\d+\s+", "", t).split()[0].split(".")[0] for _, t in ins)
    print("  all:", " ".join("%s:%d" % x for x in hist.most_common(14)))
    for i, (a, t) in enumerate(ins):
        m = re.search(r"BRA\S*\s+(?:.*?,\s*)?`?\(?\.?L_x_\d+\)?|BRA\S*\s+0x([0-9a-f]+)", t)
        m2 = re.search(r"BRA.*?0x([0-9a-f]+)", t)
        if m2:
            tgt = int(m2.group(1), 16)
            if tgt <= a and tgt in addr_index:
                j = addr_index[tgt]
                n = i - j + 1
                if n >= 40:
                    h = collections.Counter(re.sub(r"^@!?U?P\d+\s+", "", x).split()[0].split(".")[0] for _, x in ins[j:i + 1])
                    print("  loop %5d..%5d n=%4d : %s" % (j, i, n, " ".join("%s:%d" % x for x in h.most_common(16))))
