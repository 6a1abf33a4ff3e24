# Round-2 evidence capture on one B200 (final kernels): ncu --set full of EVERY kernel of the default path on C3
# (early = free fall, late = piled up), launch lists (device time of every launch, serialised / cold cache: compare
# shares, not absolutes) of the benchmark command, of the late state and of one strip rank of the 8-GPU run.
# Numbers printed under ncu are never bench values.
mkdir -p gpurun_out
K='regex:k[0-9]*_|k_'
timeout 400 ncu --set full --clock-control none --import-source on -k "$K" -s 18 -c 9 -o gpurun_out/r2_final_c3_early \
    python profiles/ncu_c3_chain.py > gpurun_out/r2_cap1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k "$K" -s 1953 -c 9 -o gpurun_out/r2_final_c3_late \
    python profiles/ncu_c3_chain.py 27 1 > gpurun_out/r2_cap2.log 2>&1
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 180 --csv --log-file gpurun_out/r2_launches_bench_steps2.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-scaling-ref --no-clock-window > gpurun_out/r2_cap3.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -s 1944 -c 72 --csv --log-file gpurun_out/r2_launches_c3_late.csv \
    python profiles/ncu_c3_chain.py 27 1 > gpurun_out/r2_cap4.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -s 200 -c 60 --csv --log-file gpurun_out/r2_launches_strip_rank.csv \
    python profiles/strip_rank_kernels.py ncu 8 3 > gpurun_out/r2_cap5.log 2>&1
tail -2 gpurun_out/r2_cap*.log
ls -la gpurun_out/*.ncu-rep gpurun_out/r2_launches_*.csv
