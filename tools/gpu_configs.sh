# every BASELINE config through bench.py (one JSON line each into gpurun_out/cfg_<name>.json); usage: bash tools/gpu_configs.sh [names...]
names="${@:-cfg3 cfg2 cfg5 cfg5-tt4}"
for c in $names; do
  steps=10; [ "$c" = cfg5 ] || [ "$c" = cfg5-tt4 ] || [ "$c" = cfg4-full ] && steps=3
  timeout 1500 python bench.py --config $c --steps $steps --warmup 3 > gpurun_out/cfg_$c.log 2>&1
  grep '^{"metric' gpurun_out/cfg_$c.log > gpurun_out/cfg_$c.json
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/cfg_$c.json").read())
    print("$c", "value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "py", d.get("e2e_python"), "roof", round(d["roofline"]["frac"],4), round(d["roofline"]["kernel_ms"],2), "cpu", d["cpu_baseline"] and round(d["cpu_baseline"]["value"],2), "parity", d.get("parity_checked"), d.get("parity_mismatches"), {k:round(v,1) for k,v in d["phases_ms_rank0"].items() if k!="host_issue_ms"})
except Exception as e:
    print("$c FAILED", e); print(open("gpurun_out/cfg_$c.log").read()[-1500:])
PY
done
