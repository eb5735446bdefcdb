# A/B helper for one gpurun call: parity tests (TESTS = pytest -k filter, default all GPU tests), then bench.py under
# alternative kernel selections (env switches of the library: arguments NAME=VALUE), then optionally one ncu --set full
# capture (NCU = kernel-name regex, TAG = output name).  Example:
#   TESTS="shared_memory or two_lane" NCU="k_coding_flat|k_start_score_lean" TAG=r2h bash tools/gpu_ab.sh PGPU_CODING_SMEM=0
python -m pytest tests -m gpu -x -q ${TESTS:+-k "$TESTS"} > gpurun_out/ab_gpu_tests.log 2>&1; tail -3 gpurun_out/ab_gpu_tests.log
run() { name=$1; shift; env "$@" python bench.py --no-cpu --steps 3 --warmup 3 > gpurun_out/ab_bench_$name.log 2>&1; python - <<PY
import json
for l in open("gpurun_out/ab_bench_$name.log"):
    if l.startswith('{"metric'):
        d=json.loads(l); p=d['phases_ms_rank0']; print("$name", round(d['value']), round(d['e2e']['value']), {k:round(v,1) for k,v in p.items() if k!='host_issue_ms'})
PY
}
run default X=1
for v in "$@"; do run "$(echo $v | tr '=' '_')" $v; done
if [ -n "$NCU" ]; then
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:"$NCU" -c ${NCU_COUNT:-3} -f -o gpurun_out/${TAG:-ab}_ncu python bench.py --steps 1 --warmup 0 --no-cpu > gpurun_out/${TAG:-ab}_ncu.log 2>&1; ls -la gpurun_out/${TAG:-ab}_ncu.ncu-rep
fi
