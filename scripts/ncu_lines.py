"""Joins an ncu SASS-level source csv with nvdisasm -g line info and aggregates stall samples and
executed instructions per CUDA source line.  usage: ncu_lines.py <ncu-rep> <kernel-substr> [top]"""
import csv, os, re, subprocess, sys, tempfile
rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.environ.get("CVO_B200_LIB") or os.path.join(ROOT, "cvo_slam_b200", "libcvo_b200.so")   # the build the report was taken on
CSRC = os.environ.get("CVO_B200_CSRC") or os.path.join(ROOT, "cvo_slam_b200", "csrc")
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
        if m and cur_kernel and kern in cur_kernel:
            addr2line[int(m.group(1), 16)] = cur_line
csvtxt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(csvtxt.splitlines()))
hdrs = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
names = [rows[i - 1][1] if i > 0 and len(rows[i - 1]) > 1 else "" for i in hdrs]
sel = [k for k, n in enumerate(names) if kern in n.replace("(bool)1", "ILb1").replace("(bool)0", "ILb0") or kern in n]
k = sel[0] if sel else 0
i0 = hdrs[k]; i1 = hdrs[k + 1] - 1 if k + 1 < len(hdrs) else len(rows)
hdr = rows[i0]; data = rows[i0 + 1:i1]
ca, cs, ci, ct = hdr.index("Address"), hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed")
base = None
agg = {}
tot_s = tot_i = 0
for r in data:
    if len(r) <= ct or not r[ci].isdigit(): continue
    a = int(r[ca], 16) if r[ca].startswith("0x") else int(r[ca])
    if base is None: base = a
    line = addr2line.get(a - base)
    s, n, tn = int(r[cs] or 0), int(r[ci]), int(r[ct])
    d = agg.setdefault(line, [0, 0, 0])
    d[0] += s; d[1] += n; d[2] += tn
    tot_s += s; tot_i += n
print("kernel:", names[k][:100]); print("total samples", tot_s, "warp instr", tot_i)
src = {}
def text(line):
    if not line: return ""
    f = os.path.join(CSRC, line[0])
    if f not in src:
        try: src[f] = open(f).read().splitlines()
        except Exception: src[f] = []
    L = src[f]
    return L[line[1] - 1].strip()[:90] if 0 < line[1] <= len(L) else ""
for line, (s, n, tn) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{100*s/max(tot_s,1):5.1f}% smp {100*n/max(tot_i,1):5.1f}% ins thr/ins {tn/max(n,1):4.1f} {line} | {text(line)}")
