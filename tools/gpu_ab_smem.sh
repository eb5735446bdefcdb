# A/B of k_coding_smem and the resident lanes (one gpurun call): the new parity tests, three bench runs, one ncu capture
python -m pytest tests -m gpu -x -q -k "shared_memory or two_lane or resident or lane_groups" > gpurun_out/ab_gpu_tests.log 2>&1; tail -3 gpurun_out/ab_gpu_tests.log
run() { name=$1; shift; env "$@" python bench.py --no-cpu --steps 3 --warmup 3 > gpurun_out/ab_bench_$name.log 2>&1; python - <<PY
import json
for l in open("gpurun_out/ab_bench_$name.log"):
    if l.startswith('{"metric'):
        d=json.loads(l); p=d['phases_ms_rank0']; print("$name", round(d['value']), round(d['e2e']['value']), {k:round(v,1) for k,v in p.items() if k!='host_issue_ms'})
PY
}
run default X=1
for v in "$@"; do run "$(echo $v | tr '=' '_')" $v; done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_coding_flat|k_cq_plan|k_cq_owner|k_orf_links" -c 4 -f -o gpurun_out/r2g_coding_flat python bench.py --steps 1 --warmup 0 --no-cpu > gpurun_out/r2g_ncu.log 2>&1; ls -la gpurun_out/r2g_coding_flat.ncu-rep
