# A/B helper for gpurun: parity tests, then bench.py under alternative kernel selections (env switches of the library).
# Switches that are read once per process (PGPU_SCORE_BY_CLASS, PGPU_CODING_MINB) can also be verified by running the
# whole GPU suite under them: `PGPU_SCORE_BY_CLASS=1 python -m pytest tests -m gpu -x -q`.
python -m pytest tests -m gpu -x -q > gpurun_out/ab_gpu_tests.log 2>&1; tail -3 gpurun_out/ab_gpu_tests.log
run() { name=$1; shift; env "$@" python bench.py --no-cpu --steps 3 --warmup 3 > gpurun_out/ab_bench_$name.log 2>&1; python - <<PY
import json
for l in open("gpurun_out/ab_bench_$name.log"):
    if l.startswith('{"metric'):
        d=json.loads(l); p=d['phases_ms_rank0']; print("$name", round(d['value']), round(d['e2e']['value']), {k:round(v,1) for k,v in p.items() if k!='host_issue_ms'})
PY
}
run default X=1
for v in "$@"; do run "$(echo $v | tr '=' '_')" $v; done
