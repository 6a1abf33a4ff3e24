# Round-2 first GPU call (1 GPU, ~5 min): baseline of the round-1 path on today's box, the lane-dense
# narrowphase variant on the device (parity + early/late timing), per-kernel times along the C3 evolution.
mkdir -p gpurun_out
OUT=gpurun_out/r2_first.txt
: > $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv | tee -a $OUT
BENDY_TEST_UNPROVEN=1 timeout 300 python -m pytest tests/test_z_gpu_variants.py -m gpu -x -q -k "dense" > gpurun_out/r2_dense_tests.log 2>&1
echo "dense variant tests exit code $?" | tee -a $OUT
tail -3 gpurun_out/r2_dense_tests.log | tee -a $OUT
for cfg in "" "BENDY_NARROW_DENSE=1"; do
  env $cfg timeout 200 python profiles/quick_c3.py "C3 ${cfg:-default}" | tee -a $OUT
done
timeout 200 python profiles/late_state_kernels.py 2>&1 | tee -a $OUT
BENDY_NARROW_DENSE=1 timeout 200 python profiles/late_state_kernels.py 2>&1 | tee -a $OUT
