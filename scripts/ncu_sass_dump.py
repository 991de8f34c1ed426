"""Dumps the SASS of an ncu source page with executed counts and samples, annotated with the CUDA line, between two
source lines of align.cu.  usage: ncu_sass_dump.py <rep> <mangled-kernel-substr> <lo> <hi>"""
import csv, os, re, subprocess, sys, tempfile
rep, kern, lo, hi = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4])
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.environ.get("CVO_B200_LIB") or os.path.join(ROOT, "cvo_slam_b200", "libcvo_b200.so")
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", so], cwd=tmp, capture_output=True)
addr2line = {}
for f in os.listdir(tmp):
    if not f.endswith(".cubin"): continue
    out = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, f)], capture_output=True, text=True).stdout
    cur_kernel, cur_line = None, None
    for ln in out.splitlines():
        m = re.match(r"\s*\.section\s+\.text\.(\S+?),", ln)
        if m: cur_kernel = m.group(1); continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m: cur_line = (os.path.basename(m.group(1)), int(m.group(2))); continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(\S.*);", ln)
        if m and cur_kernel and kern in cur_kernel: addr2line[int(m.group(1), 16)] = cur_line
csvtxt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(csvtxt.splitlines()))
hdrs = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
i0 = hdrs[0]; hdr = rows[i0]; data = rows[i0 + 1:]
ca, cs, ci, csrc = hdr.index("Address"), hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Source")
base = None; on = False; first = last = None
recs = []
for r in data:
    if len(r) <= ci or not r[ci].isdigit(): continue
    a = int(r[ca], 16) if r[ca].startswith("0x") else int(r[ca])
    if base is None: base = a
    recs.append((a - base, addr2line.get(a - base), int(r[cs] or 0), int(r[ci]), r[csrc]))
idx = [k for k, (a, l, s, n, t) in enumerate(recs) if l and l[0] == "align.cu" and lo <= l[1] <= hi]
for a, l, s, n, t in recs[idx[0]:idx[-1] + 1]:
    print(f"{a:6x} {str(l[1]) if l else '?':>5} {l[0][:12] if l else '':12s} {n:11d} {s:6d}  {t}")
