"""Like ncu_lines.py but sorted by executed warp-instructions: ncu_lines_ins.py <rep> <kernel-substr> [top]"""
import subprocess, sys
out = subprocess.run([sys.executable, __file__.replace("ncu_lines_ins.py", "ncu_lines.py"), sys.argv[1], sys.argv[2], "2000"], capture_output=True, text=True).stdout
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
rows = []
for l in out.splitlines()[2:]:
    try:
        smp = float(l.split("% smp")[0]); ins = float(l.split("% smp")[1].split("% ins")[0])
        rows.append((ins, smp, l.split("thr/ins")[1]))
    except Exception: pass
rows.sort(reverse=True)
print("\n".join(out.splitlines()[:2]))
for ins, smp, rest in rows[:top]: print(f"{ins:5.1f}% ins {smp:5.1f}% smp {rest[:140]}")
