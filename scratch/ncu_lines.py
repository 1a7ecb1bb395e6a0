"""Join an ncu --page source CSV (SASS view, possibly several kernels) with nvdisasm --print-line-info
output.  usage: ncu_lines.py src.csv lines.sass [topN] [kernel substring] [sort: samples|inst]"""
import csv, re, sys, collections
src_csv, lines_sass = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
want = sys.argv[4] if len(sys.argv) > 4 else ""
sortby = sys.argv[5] if len(sys.argv) > 5 else "samples"
addr2line = {}
cur = None
for ln in open(lines_sass):
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (m.group(1).split('/')[-1], int(m.group(2))); continue
    m = re.match(r'\s+/\*([0-9a-f]+)\*/', ln)
    if m and cur: addr2line[int(m.group(1), 16)] = cur
rows = list(csv.reader(open(src_csv)))
# sections start with a "Kernel Name" row followed by a header row
sections = []
for i, r in enumerate(rows):
    if r and r[0] == "Kernel Name": sections.append(i)
sections.append(len(rows))
for si in range(len(sections) - 1):
    name = rows[sections[si]][1]
    if want not in name: continue
    hdr = rows[sections[si] + 1]
    body = rows[sections[si] + 2:sections[si + 1]]
    ia, ii, isamp = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("# Samples")
    stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    base = None
    agg = collections.defaultdict(lambda: [0, 0, 0, collections.Counter()])
    tot_i = tot_s = 0
    for r in body:
        if len(r) <= isamp: continue
        a = int(r[ia], 16)
        if base is None: base = a
        key = addr2line.get(a - base, ("?", 0))
        n = int(r[ii] or 0); s = int(r[isamp] or 0)
        e = agg[key]; e[0] += n; e[1] += s; e[2] += 1
        for c in stall_cols:
            v = int(r[c] or 0)
            if v: e[3][hdr[c]] += v
        tot_i += n; tot_s += s
    print("== %s\ntotal inst executed %d, samples %d, static instrs %d" % (name[:60], tot_i, tot_s, sum(e[2] for e in agg.values())))
    idx = 1 if sortby == "samples" else 0
    for key, e in sorted(agg.items(), key=lambda kv: -kv[1][idx])[:top]:
        st = ", ".join("%s %d" % (k.replace("stall_", ""), v) for k, v in e[3].most_common(3))
        print("%-22s:%4d  inst %5.1f%%  samples %5.1f%%  static %5d   %s" % (key[0], key[1], 100 * e[0] / max(tot_i, 1), 100 * e[1] / max(tot_s, 1), e[2], st))
