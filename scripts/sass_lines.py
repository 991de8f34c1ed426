"""Static SASS instruction count per CUDA source line of one kernel (nvdisasm -g line info).
usage: sass_lines.py <kernel-substr> [lo-line hi-line]   — run after building the library."""
import collections, os, re, subprocess, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.environ.get("CVO_B200_LIB") or os.path.join(ROOT, "cvo_slam_b200", "libcvo_b200.so")
kern = sys.argv[1]
lo, hi = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (0, 1 << 30)
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", so], cwd=tmp, capture_output=True)
cnt = collections.Counter(); ops = collections.defaultdict(collections.Counter)
for f in os.listdir(tmp):
    if "align" not in f or not f.endswith(".cubin"): continue
    out = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, f)], capture_output=True, text=True).stdout
    cur_k, cur_l = None, None
    for ln in out.splitlines():
        m = re.match(r"\s*\.section\s+\.text\.(\S+?),", ln)
        if m: cur_k = m.group(1); continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m: cur_l = (os.path.basename(m.group(1)), int(m.group(2))); continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d\s+)?(\S+)", ln)
        if m and cur_k and kern in cur_k and cur_l:
            cnt[cur_l] += 1; ops[cur_l][m.group(1).split(".")[0]] += 1
src = open(os.path.join(ROOT, "cvo_slam_b200", "csrc", "align.cu")).read().splitlines()
tot = 0
for (f, l), n in sorted(cnt.items(), key=lambda kv: (kv[0][0], kv[0][1])):
    if f != "align.cu" or not (lo <= l <= hi): continue
    tot += n
    top = " ".join(f"{k}:{v}" for k, v in ops[(f, l)].most_common(6))
    print(f"{l:5d} {n:4d}  {top:60s} | {src[l-1].strip()[:70]}")
print("total", tot)
