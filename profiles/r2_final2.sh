# final 1-GPU pass of the round: the whole gpu suite (incl. full-size parity and the normalize check), smoke, bench lines
# of C1..C4 + the reference arm, then ncu evidence of the final kernels (never bench values)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_final_gpu_tests.log 2>&1
echo "gpu tests exit code $?"; tail -4 gpurun_out/r2_final_gpu_tests.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
bash profiles/r2_bench_all.sh
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2_bench_reference_c3.json 2> gpurun_out/r2_bench_reference_c3.err; tail -c 400 gpurun_out/r2_bench_reference_c3.json
K='regex:k[0-9]*_|k_'
timeout 400 ncu --set full --clock-control none --import-source on -k "$K" -s 18 -c 9 -f -o gpurun_out/r2_final_c3_early \
    python profiles/ncu_c3_chain.py > gpurun_out/r2_cap1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k "$K" -s 1953 -c 9 -f -o gpurun_out/r2_final_c3_late \
    python profiles/ncu_c3_chain.py 27 1 > gpurun_out/r2_cap2.log 2>&1
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 180 --csv --log-file gpurun_out/r2_launches_bench_steps2.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-scaling-ref --no-clock-window > gpurun_out/r2_cap3.log 2>&1
ls -la gpurun_out/*.ncu-rep gpurun_out/r2_launches_*.csv
