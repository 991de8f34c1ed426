"""Prints the key numbers of bench.py JSON lines: show_bench.py file..."""
import json, sys
for f in sys.argv[1:]:
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, "ERR", e, open(f).read()[-300:]); continue
    r = d.get("roofline", {})
    print(f, "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ms/step", round(d["ms_per_step"], 1),
          "kernel ms", round(r.get("kernel_ms_per_launch", 0), 1), "frac", round(r.get("frac", 0), 4),
          "iters/pair", round(r.get("iterations_per_pair", 0), 1), r.get("phase_share"), d.get("check"))
