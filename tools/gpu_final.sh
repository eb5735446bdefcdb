# round-end evidence in one gpurun call: parity tests, the bench line, the reference arm, the ncu launch list of the same
# command and one --set full capture of the top kernels (outputs under gpurun_out/, tag = $1)
T=${1:-r2}
python -m pytest tests -x -q -m gpu > gpurun_out/${T}_gpu_tests.log 2>&1; tail -2 gpurun_out/${T}_gpu_tests.log
python bench.py > gpurun_out/${T}_bench.log 2>&1; tail -c 600 gpurun_out/${T}_bench.log
python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/${T}_bench_reference.log 2>&1; tail -c 400 gpurun_out/${T}_bench_reference.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/${T}_launch_bench.log 2>&1; wc -l gpurun_out/${T}_launches.csv
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"k_dp_ml|k_coding_flat|k_orf_links|k_start_score_lean|k_overlap_lanes|k_codon_bits|k_extract_b|k_encode|k_node_prep|k_trace" -s 33 -c 11 -f -o gpurun_out/${T}_top python bench.py --steps 1 --warmup 0 --no-cpu > gpurun_out/${T}_ncu.log 2>&1; ls -la gpurun_out/${T}_top.ncu-rep
