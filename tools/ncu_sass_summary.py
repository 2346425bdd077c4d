#!/usr/bin/env python
"""Summarise `ncu -i X.ncu-rep --page source --csv --print-source sass` output: instruction mix by opcode and
stall samples by code region (regions split at the big role branches are found by sample clustering)."""
import csv, sys, collections, re
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
ops = collections.Counter(); samp = collections.Counter(); tot = 0; tots = 0
data = []
for r in rows[2:]:
    if len(r) < len(hdr) - 5: continue
    src = r[ix["Source"]].strip()
    m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_.]+)", src)
    op = m.group(2) if m else src
    base = op.split(".")[0]
    n = int(r[ix["Instructions Executed"]] or 0); s = int(r[ix["# Samples"]] or 0)
    ops[base] += n; samp[base] += s; tot += n; tots += s
    data.append((src, n, s, r))
print("total warp instructions %d, samples %d" % (tot, tots))
for op, n in ops.most_common(28):
    print("  %-10s %12d  %5.1f%%   samples %6d %5.1f%%" % (op, n, 100.0 * n / tot, samp[op], 100.0 * samp[op] / max(1, tots)))
# regions: cut where executed count changes by role -> print coarse histogram over address order in 40 bins
nb = int(sys.argv[2]) if len(sys.argv) > 2 else 40
L = len(data); step = (L + nb - 1) // nb
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
print("bins over SASS order: [first line] instr-executed, samples, top stalls")
for b in range(nb):
    seg = data[b * step:(b + 1) * step]
    if not seg: break
    n = sum(x[1] for x in seg); s = sum(x[2] for x in seg)
    st = collections.Counter()
    for x in seg:
        for h in stall_cols:
            v = x[3][ix[h]]
            if v and v != "0": st[h[6:]] += int(v)
    dfma = sum(x[1] for x in seg if x[0].lstrip().startswith("DFMA") or " DFMA" in x[0][:14])
    print("  %4d %-40s n=%11d dfma=%10d s=%6d  %s" % (b * step, seg[0][0][:40], n, dfma, s, ", ".join("%s %d" % kv for kv in st.most_common(4))))
