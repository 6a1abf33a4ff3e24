# device check of the current build: fast gpu tests, C3 early / mid / late, per-kernel event times
mkdir -p gpurun_out
OUT=gpurun_out/r2_check.txt
: > $OUT
timeout 900 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_full_size.py > gpurun_out/r2_check_tests.log 2>&1
echo "gpu tests exit code $?" | tee -a $OUT
tail -4 gpurun_out/r2_check_tests.log | tee -a $OUT
timeout 200 python profiles/quick_c3.py "C3 ${1:-build}" | tee -a $OUT
timeout 200 python profiles/late_state_kernels.py 2>&1 | grep after | tee -a $OUT
