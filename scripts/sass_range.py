"""Prints the SASS of one kernel between the first and last instruction attributed to a source-line range
(nvdisasm -g).  usage: sass_range.py <mangled-kernel-substr> <lo> <hi> [which-occurrence]"""
import os, re, subprocess, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.environ.get("CVO_B200_LIB") or os.path.join(ROOT, "cvo_slam_b200", "libcvo_b200.so")
kern, lo, hi = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", so], cwd=tmp, capture_output=True)
recs = []
for f in os.listdir(tmp):
    if "align" not in f or not f.endswith(".cubin"): continue
    out = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, f)], capture_output=True, text=True).stdout
    cur_k, cur_l = None, None
    for ln in out.splitlines():
        m = re.match(r"\s*\.section\s+\.text\.(\S+?),", ln)
        if m: cur_k = m.group(1); continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m: cur_l = (os.path.basename(m.group(1)), int(m.group(2))); continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(\S.*);", ln)
        if m and cur_k and kern in cur_k: recs.append((int(m.group(1), 16), cur_l, m.group(2)))
        elif cur_k and kern in cur_k and re.match(r"\s*\.L_x_\d+:", ln): recs.append((None, None, ln.strip()))
idx = [k for k, (a, l, t) in enumerate(recs) if l and l[0] == "align.cu" and lo <= l[1] <= hi]
# contiguous groups
groups = []; cur = [idx[0]]
for k in idx[1:]:
    if k - cur[-1] > 40: groups.append(cur); cur = [k]
    else: cur.append(k)
groups.append(cur)
which = int(sys.argv[4]) if len(sys.argv) > 4 else 0
g = groups[which]
print(f"{len(groups)} group(s); group {which}: {g[-1]-g[0]+1} instructions")
for a, l, t in recs[g[0]:g[-1] + 1]:
    if a is None: print("        ", t)
    else: print(f"{a:6x} {l[1] if l else '?':>5} {t}")
