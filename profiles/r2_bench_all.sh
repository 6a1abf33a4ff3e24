# bench lines of every named configuration on one GPU (BASELINE.json configs C1..C4; C5 is the multi-GPU series)
mkdir -p gpurun_out
for w in c1 c2 c4 c3; do
  timeout 900 python bench.py --workload $w --steps 20 --warmup 5 $( [ $w != c3 ] && echo --no-scaling-ref ) > gpurun_out/r2_bench_$w.json 2> gpurun_out/r2_bench_$w.err
  echo "== $w rc=$?"; python profiles/show_bench.py gpurun_out/r2_bench_$w.json | head -4
done
