# final 1-GPU pass: the whole gpu suite (incl. full-size parity), smoke, bench lines of C1..C4 with the final build
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_final_gpu_tests.log 2>&1
echo "gpu tests exit code $?"; tail -4 gpurun_out/r2_final_gpu_tests.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
bash profiles/r2_bench_all.sh
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2_bench_reference_c3.json 2> gpurun_out/r2_bench_reference_c3.err; tail -c 600 gpurun_out/r2_bench_reference_c3.json
