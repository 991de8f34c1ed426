"""Aggregates an ncu SASS-level source csv of an align kernel by PHASE of align_one: every SASS instruction is
attributed to the phase of the nearest preceding instruction whose line lies in align_one's body (inlined helpers
inherit it).  usage: ncu_phases.py <ncu-rep> <mangled-kernel-substr>   (phase boundaries are found by marker comments)"""
import csv, os, re, subprocess, sys, tempfile
rep, kern = sys.argv[1], sys.argv[2]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.environ.get("CVO_B200_LIB") or os.path.join(ROOT, "cvo_slam_b200", "libcvo_b200.so")
src = open(os.path.join(ROOT, "cvo_slam_b200", "csrc", "align.cu")).read().splitlines()
def find(pat, start=0):
    for i in range(start, len(src)):
        if pat in src[i]: return i + 1
    raise SystemExit("marker not found: " + pat)
body0 = find("__device__ void align_one(")
marks = [("setup", body0), ("grid", find("if (sh.do_grid) {", body0)), ("P1a filter", find("---- filter: the list of a smaller length scale", body0)), ("P1a search", find("const int tile = q * csize + crank;", body0)),
         ("P1a ck+prune", find("colour kernel of the tile's raw hits", body0)), ("sort rows", find("---- columns sorted by entry count", body0)),
         ("tile widths", find("tile widths (steps of 32 entries)", body0)), ("pads+scatter", find("pads of the tiles, then the kept entries", body0)),
         ("P1b", find("---------------- P1b", body0)), ("P1b tiles", find("while (Tcur < nT) {", body0)), ("reduce1+omega", find("wg_reduce_i64<kMode>(sh);", body0)),
         ("P2 rows", find("---------------- P2", body0)), ("P2 entries", find("float fB = 0.f, fC = 0.f", body0)), ("P2 tiles", find("the warp's next tile: its row data into registers, its verdicts into L2", body0)),
         ("reduce2", find("wg_reduce_dd4<kMode>(bc, sh);", body0)), ("P3", find("---------------- P3", body0)),
         ("epilogue", find("if (kMode == 2) {   // gather the overflow flags", body0))]
body1 = find("struct ScratchBase", body0)
def phase_of(line):
    if line is None or line[0] != "align.cu" or not (body0 <= line[1] < body1): return None
    p = None
    for name, l0 in marks:
        if line[1] >= l0: p = name
    return p
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", so], cwd=tmp, capture_output=True)
addr2line, addr2op = {}, {}
for f in os.listdir(tmp):
    if not f.endswith(".cubin"): continue
    out = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, f)], capture_output=True, text=True).stdout
    cur_kernel, cur_line = None, None
    for ln in out.splitlines():
        m = re.match(r"\s*\.section\s+\.text\.(\S+?),", ln)
        if m: cur_kernel = m.group(1); continue
        m = re.search(r'//## File "([^"]+)", line (\d+)( inlined at "([^"]+)", line (\d+))?', ln)
        if m:
            cur_line = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(\S.*);", ln)
        if m and cur_kernel and kern in cur_kernel:
            addr2line[int(m.group(1), 16)] = cur_line
            addr2op[int(m.group(1), 16)] = m.group(2)
csvtxt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(csvtxt.splitlines()))
hdrs = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
i0 = hdrs[0]; hdr = rows[i0]; data = rows[i0 + 1:]
ca, cs, ci, ct = hdr.index("Address"), hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed")
STALLS = ["stall_barrier", "stall_branch_resolving", "stall_dispatch", "stall_lg", "stall_long_sb", "stall_math", "stall_mio", "stall_no_inst",
          "stall_not_selected", "stall_selected", "stall_short_sb", "stall_wait", "stall_sleep", "stall_membar", "stall_drain", "stall_misc", "stall_tex"]
cst = [hdr.index(n) for n in STALLS]
stl = {}
base = None; agg = {}; cur = "setup"; tot = [0, 0, 0]; ops = {}
for r in data:
    if len(r) <= ct or not r[ci].isdigit(): continue
    a = int(r[ca], 16) if r[ca].startswith("0x") else int(r[ca])
    if base is None: base = a
    p = phase_of(addr2line.get(a - base))
    if p: cur = p
    s, n, tn = int(r[cs] or 0), int(r[ci]), int(r[ct])
    d = agg.setdefault(cur, [0, 0, 0]); d[0] += s; d[1] += n; d[2] += tn
    sv = stl.setdefault(cur, [0] * len(STALLS))
    for q, c in enumerate(cst): sv[q] += int(r[c] or 0)
    tot[0] += s; tot[1] += n; tot[2] += tn
    op = addr2op.get(a - base, "?").split()[0]
    if op.startswith("@"): op = addr2op[a - base].split()[1]
    op = op.split(".")[0]
    o = ops.setdefault((cur, op), [0, 0]); o[0] += n; o[1] += s
print(f"total samples {tot[0]} warp-instr {tot[1]} thread-instr {tot[2]}")
for name, _ in marks:
    if name in agg:
        s, n, tn = agg[name]
        sv = stl[name]; top = sorted(range(len(STALLS)), key=lambda q: -sv[q])[:5]
        print(f"{name:14s} {100*s/tot[0]:5.1f}% smp {100*n/tot[1]:5.1f}% ins  thr/ins {tn/max(n,1):4.1f}  smp/ins {s/max(n,1)*tot[1]/tot[0]:4.2f}  " +
              " ".join(f"{STALLS[q][6:]}:{100*sv[q]/max(sum(sv),1):.0f}%" for q in top))
if len(sys.argv) > 3:
    for ph in sys.argv[3:]:
        print("--", ph)
        for (p, op), (n, s) in sorted(ops.items(), key=lambda kv: -kv[1][0]):
            if p == ph and n > 0.002 * agg[ph][1]: print(f"   {op:10s} {100*n/agg[ph][1]:5.1f}% ins {100*s/max(agg[ph][0],1):5.1f}% smp")
