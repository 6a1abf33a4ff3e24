# A/B of the opt-in variants of round 1 on the device (C3 early / mid / late)
mkdir -p gpurun_out
OUT=gpurun_out/r2_ab.txt
: > $OUT
BENDY_TEST_UNPROVEN=1 timeout 600 python -m pytest tests/test_z_gpu_variants.py -m gpu -q > gpurun_out/r2_variant_tests.log 2>&1
echo "variant tests exit code $?" | tee -a $OUT
tail -4 gpurun_out/r2_variant_tests.log | tee -a $OUT
for cfg in "" "BENDY_SCATTER_ILP=1" "BENDY_SORT_FUSED=1" "BENDY_SCATTER_AGG=1" "BENDY_SCATTER_ILP=1 BENDY_K3_THREADS=192" $EXTRA_CFGS; do
  env $cfg timeout 200 python profiles/quick_c3.py "C3 ${cfg:-default}" | tee -a $OUT
done
