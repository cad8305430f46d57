"""Join an ncu SASS source page with nvdisasm line info: executed instructions per CUDA source line.
usage: python tools/ncu_lines.py <report.ncu-rep> <mangled-kernel-substring> [top] [ncu kernel-name regex if the report holds several kernels]"""
import collections, csv, glob, os, re, subprocess, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(ROOT, "montecarlolocalisation_b200", "libmcl_b200.so")], cwd=tmp, capture_output=True)
m = {}
for cub in glob.glob(tmp + "/*.cubin"):
    txt = subprocess.run(["nvdisasm", "-g", "-c", cub], capture_output=True, text=True).stdout
    heads = [mm.start() for mm in re.finditer(r"//-+ \.text\.", txt)]
    for i, st in enumerate(heads):
        en = heads[i + 1] if i + 1 < len(heads) else len(txt)
        body = txt[st:en]
        if kern not in body[:300]:
            continue
        line = None
        for l in body.split("\n"):
            mm = re.search(r'//## File "([^"]+)", line (\d+)', l)
            if mm:
                line = (mm.group(1).split("/")[-1], int(mm.group(2)))
                continue
            mm = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
            if mm:
                m[int(mm.group(1), 16)] = line
sel = ["-k", "regex:" + sys.argv[4], "-c", "1"] if len(sys.argv) > 4 else []
csvtxt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"] + sel, capture_output=True, text=True).stdout
rows = list(csv.reader(csvtxt.splitlines()))
hi = [i for i, r in enumerate(rows) if "Source" in r][0]
h = rows[hi]
iA, iE, iT, iSm = h.index("Address"), h.index("Instructions Executed"), h.index("Thread Instructions Executed"), h.index("# Samples")
base = None
agg, aggs, aggt = collections.Counter(), collections.Counter(), collections.Counter()
tot = 0
for r in rows[hi + 1:]:
    if "Source" in r:
        break
    try:
        a = int(r[iA], 16); e = int(r[iE])
    except Exception:
        continue
    if base is None:
        base = a
    ln = m.get(a - base)
    agg[ln] += e; aggs[ln] += int(r[iSm] or 0); aggt[ln] += int(r[iT]); tot += e
src = {}
print("mapped", len(m), "instructions; total executed warp-instructions", tot)
for ln, e in agg.most_common(top):
    t = ""
    if ln:
        if ln[0] not in src:
            p = os.path.join(ROOT, "montecarlolocalisation_b200", "csrc", ln[0])
            src[ln[0]] = open(p).read().split("\n") if os.path.exists(p) else []
        if ln[1] - 1 < len(src[ln[0]]):
            t = src[ln[0]][ln[1] - 1].strip()[:100]
    print("%10d %5.1f%% lanes %5.1f smp %5d %s | %s" % (e, 100 * e / tot, aggt[ln] / max(e, 1), aggs[ln], ln, t))
