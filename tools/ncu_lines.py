"""Attribute executed instructions / stall samples of one kernel to CUDA source lines.
usage: python tools/ncu_lines.py <lib.so> <mangled kernel name> <ncu source-page csv> [metric column]
Joins `nvdisasm --print-line-info` of the cubin in <lib.so> with the SASS rows of the ncu source page."""
import re, csv, collections, subprocess, sys, tempfile, os, glob
so, kern, src = sys.argv[1:4]
col = sys.argv[4] if len(sys.argv) > 4 else 'Instructions Executed'
d = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=d, check=True, capture_output=True)
L = subprocess.run(["nvdisasm", "--print-line-info", glob.glob(d + "/*.cubin")[0]], capture_output=True, text=True).stdout.split('\n')
start = next(i for i, l in enumerate(L) if l.startswith('.text.' + kern + ':'))
end = next(i for i in range(start + 1, len(L)) if L[i].startswith('\t.section'))
cur = None; ins = []
for l in L[start:end]:
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: cur = (m.group(1).split('/')[-1], int(m.group(2))); continue
    m = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(.*?);', l)
    if m: ins.append((cur, m.group(2)))
rows = list(csv.reader(open(src))); h = rows[1]; data = rows[2:]
assert len(ins) == len(data), (len(ins), len(data), "library and capture are different builds")
c = h.index(col); by = collections.Counter(); tot = 0
for (tag, txt), r in zip(ins, data):
    n = int(r[c] or 0); tot += n; by[tag] += n
srcs = {}
for k, v in by.most_common(40):
    if k and k[0] not in srcs:
        cand = glob.glob(os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'voxelyze_b200', 'csrc', k[0]))
        srcs[k[0]] = open(cand[0]).read().split('\n') if cand else None
    text = srcs[k[0]][k[1] - 1].strip()[:90] if k and srcs.get(k[0]) else ''
    print("%6.2f%%  %-28s %s" % (100 * v / tot, "%s:%d" % k if k else None, text))
