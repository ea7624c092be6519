"""Per-source-line instruction / stall-sample totals of lmpc_qp_kernel from an ncu report.

usage: python scripts/ncu_lines.py <report.ncu-rep> <liblmpc_b200.so> [mangled-kernel-substring] [top]

ncu's CSV source page is SASS only; the file:line of every SASS instruction comes from `nvdisasm -g` on the cubin
inside the .so (compiled with -lineinfo).  The two are joined on the instruction offset.
"""
import csv, io, os, re, subprocess, sys, tempfile, collections

rep, so = sys.argv[1], sys.argv[2]
kern = sys.argv[3] if len(sys.argv) > 3 else "lmpc_qp_kernelILi1ELi3ELi20ELi16"
top = int(sys.argv[4]) if len(sys.argv) > 4 else 50
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, check=True, capture_output=True)
dis = ""   # one cubin per translation unit: the kernel's section is in one of them
for cubin in sorted(f for f in os.listdir(tmp) if f.endswith(".cubin")):
    dis += subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
line_of = {}
infn = False; cur = None
for l in dis.splitlines():
    if l.startswith(".text."):
        infn = kern in l
        continue
    if not infn:
        continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m:
        line_of[int(m.group(1), 16)] = (cur, m.group(2).strip())
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hi = next(i for i, r in enumerate(rows) if "Source" in r and "Address" in r)
hdr = rows[hi]
iA, iI, iS = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("# Samples")
base = None
agg = collections.defaultdict(lambda: [0, 0])
tot = [0, 0]
ops = collections.Counter()
for r in rows[hi + 1:]:
    if len(r) < len(hdr): continue
    a = int(r[iA], 16)
    if base is None: base = a
    ln, text = line_of.get(a - base, (None, ""))
    n, s = int(r[iI]), int(r[iS])
    agg[ln][0] += n; agg[ln][1] += s; tot[0] += n; tot[1] += s
    ops[text.split()[0].split(".")[0] if text and not text.startswith("@") else (text.split()[1].split(".")[0] if text else "?")] += n
print(f"total warp instructions {tot[0]}, samples {tot[1]}")
print("top opcodes by executed instructions:", ", ".join(f"{k} {100*v/tot[0]:.1f}%" for k, v in ops.most_common(18)))
src_cache = {}
def src(ln):
    if ln is None: return ""
    f, n = ln
    for root in ("racing-lmpc-ros2_b200/csrc", "."):
        p = os.path.join(root, f)
        if os.path.exists(p):
            if p not in src_cache: src_cache[p] = open(p).read().splitlines()
            return src_cache[p][n - 1].strip()[:110]
    return ""
for ln, (n, s) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print(f"{100*n/tot[0]:5.2f}% inst {100*s/max(tot[1],1):5.2f}% samples  {ln}  {src(ln)}")

# ---- section totals (line ranges of lmpc_qp_core.cuh; everything from other files = collectives / intrinsics)
# section boundaries are found from the marker comments of lmpc_qp_core.cuh (so the table follows the source as it moves)
MARKERS = [("load + initial point", "---- load"), ("iteration control / classify", "==== interior-point iterations"),
           ("row assembly (Hessian/gradient)", "---------- rows -> per-stage Hessian"), ("terminal block (safe-set simplex)", "---------- terminal value"),
           ("stage control pieces", "---------- per-stage control pieces"), ("backward sweep, factor pass", "---------- backward Riccati sweep"),
           ("backward sweep, rhs-only pass", "// pass 1: right-hand side"), ("forward sweep", "---------- sigma_b step, forward sweep"),
           ("lambda directions", "---------- terminal directions (lambda)"), ("row directions / step length", "---------- row directions, step length"),
           ("polish update", "if (restart) { it--; continue; }"), ("iterate update", "---------- update the iterate"), ("outputs", "---- outputs")]
core = open(os.path.join("racing-lmpc-ros2_b200", "csrc", "lmpc_qp_core.cuh")).read().splitlines()
starts = []
for name, mk in MARKERS:
    ln = next(i + 1 for i, l in enumerate(core) if mk in l)
    starts.append((name, ln))
SECTIONS = [(name, a, (starts[k + 1][1] - 1) if k + 1 < len(starts) else len(core)) for k, (name, a) in enumerate(starts)]
FIRST = starts[0][1]
sec = collections.OrderedDict((name, [0, 0]) for name, _, _ in SECTIONS)
sec["row helpers (row_val/row_bound/FOR_ROWS)"] = [0, 0]
sec["collectives / shuffles (other files)"] = [0, 0]
for ln, (n, s) in agg.items():
    key = "collectives / shuffles (other files)"
    if ln is not None and ln[0] == "lmpc_qp_core.cuh":
        if ln[1] < FIRST: key = "row helpers (row_val/row_bound/FOR_ROWS)"
        else:
            for name, a, b in SECTIONS:
                if a <= ln[1] <= b: key = name; break
    sec[key][0] += n; sec[key][1] += s
print("\nsections (share of executed warp instructions / of stall samples):")
for name, (n, s) in sec.items():
    print(f"{100*n/tot[0]:6.2f}% inst {100*s/max(tot[1],1):6.2f}% samples  {name}")
