python -m pytest tests -m gpu -x -q > gpurun_out/r1g_gpu_tests.log 2>&1; tail -2 gpurun_out/r1g_gpu_tests.log
python bench.py > gpurun_out/r1g_bench.log 2>&1; tail -c 600 gpurun_out/r1g_bench.log
python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/r1g_bench_reference.log 2>&1; tail -c 400 gpurun_out/r1g_bench_reference.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1g_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/r1g_launch_bench.log 2>&1; wc -l gpurun_out/r1g_launches.csv
timeout 240 ncu --set full --clock-control none --import-source on -k regex:"k_dp_ml|k_coding_orf|k_start_score|k_overlap|k_codon_bits|k_extract_b|k_encode" -c 8 -f -o gpurun_out/r1g_top python bench.py --steps 1 --warmup 0 --no-cpu > gpurun_out/r1g_ncu.log 2>&1; ls -la gpurun_out/r1g_top.ncu-rep
