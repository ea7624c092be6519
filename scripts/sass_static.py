"""Static SASS instruction count per source line / section of one kernel (code-size view: the QP kernel is hundreds of
KB and instruction-cache misses cost real time).  usage: python scripts/sass_static.py <obj-or-so> <mangled-substring> [top]"""
import collections, os, re, subprocess, sys, tempfile
obj, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, check=True, capture_output=True)
cnt = collections.Counter(); total = 0
for cubin in sorted(f for f in os.listdir(tmp) if f.endswith(".cubin")):
    dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
    infn = False; cur = None
    for l in dis.splitlines():
        if l.startswith(".text."):
            infn = kern in l; continue
        if not infn: continue
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
        if m: cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
        if re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+\S", l): cnt[cur] += 1; total += 1
print("static instructions", total, "=", total * 16 // 1024, "KB")
core = open(os.path.join("racing-lmpc-ros2_b200", "csrc", "lmpc_qp_core.cuh")).read().splitlines()
for ln, n in cnt.most_common(top):
    src = core[ln[1] - 1].strip()[:100] if ln and ln[0] == "lmpc_qp_core.cuh" else ""
    print(f"{n:6d} {100*n/total:5.1f}%  {ln}  {src}")
# by 50-line buckets of the core file
b = collections.Counter()
for ln, n in cnt.items():
    if ln and ln[0] == "lmpc_qp_core.cuh": b[ln[1] // 50 * 50] += n
    else: b[str(ln[0]) if ln else "?"] += n
print("\nby 50-line bucket:")
for k in sorted(b, key=lambda x: (isinstance(x, str), x)): print(f"  {k}: {b[k]} ({100*b[k]/total:.1f}%)")
