"""Join an ncu source-page CSV (per SASS instruction) with nvdisasm --print-line-info of the
same cubin to attribute executed instructions and stall samples to source lines."""
import collections
import csv
import re
import sys

sass, srccsv, kernel_pat, csrc = sys.argv[1:5]
units = float(sys.argv[5]) if len(sys.argv) > 5 else None
lines = open(sass).read().split("\n")
start = next(i for i, l in enumerate(lines) if l.startswith(".text.") and kernel_pat in l)
end = next((i for i, l in enumerate(lines[start + 1:], start + 1) if l.startswith("//--------------------- .text.")), len(lines))
cur = None
insts = []
for l in lines[start:end]:
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m:
        insts.append((cur, m.group(2)))
rows = list(csv.reader(open(srccsv)))
hdr, data = rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
assert len(insts) == len(data), (len(insts), len(data))
by = collections.Counter()
sb = collections.Counter()
opsby = collections.defaultdict(collections.Counter)
for (loc, txt), r in zip(insts, data):
    n = int(r[ix["Instructions Executed"]])
    by[loc] += n
    sb[loc] += int(r[ix["# Samples"]])
    t = txt.split()
    op = t[1] if t[0].startswith("@") else t[0]
    opsby[loc][op.split(".")[0] if not op.startswith("MUFU") else op] += n
tot = sum(by.values())
stot = sum(sb.values())
src = {}
for fn in ("h9_physics.h", "h9_physics_fast.cuh", "h9_physics_pair.cuh", "h9_kernels.cuh", "h9_kernels_fast.cu",
           "h9_kernels_pair.cu"):
    try:
        src[fn] = open(csrc + "/" + fn).read().split("\n")
    except OSError:
        pass
print(f"total warp-level instructions {tot:.4e}")
order = sb.most_common if (len(sys.argv) > 7 and sys.argv[7] == "samples") else by.most_common
for loc, _ in order(int(sys.argv[6]) if len(sys.argv) > 6 else 60):
    n = by[loc]
    fn, ln = loc if loc else ("?", 0)
    text = src[fn][ln - 1].strip()[:80] if fn in src and 0 < ln <= len(src[fn]) else ""
    per = f" {n / (units / 32):6.1f}/step" if units else ""
    top = ",".join(f"{k}:{v / (units / 32):.0f}" for k, v in opsby[loc].most_common(4)) if units else ""
    print(f"{fn}:{ln:4d} inst {100 * n / tot:5.2f}%{per} samp {100 * sb[loc] / stot:5.2f}%  [{top}]  {text}")
