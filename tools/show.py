"""print the key numbers of bench.py JSON lines read from stdin (or files)"""
import json, sys
for f in (sys.argv[1:] or ["-"]):
    for l in (sys.stdin if f == "-" else open(f)):
        if not l.startswith('{"metric'):
            continue
        d = json.loads(l)
        p = d.get("phases_ms_rank0", {})
        print(d["config"].get("name"), "value", round(d["value"]), "ms", round(d["ms_per_step"], 1), "e2e", round(d["e2e"]["value"]),
              "py", d.get("e2e_python", {}).get("ms_per_step") and round(d["e2e_python"]["ms_per_step"], 1),
              "frac", round(d["roofline"]["frac"], 3), {k[3:]: round(v, 1) for k, v in p.items() if k.startswith("ms_")})
