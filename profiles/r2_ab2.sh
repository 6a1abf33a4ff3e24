# A/B of narrowphase scan-loop variants built into ab_libs/<name>/ (same box, same run): C3 early / mid / late
mkdir -p gpurun_out
OUT=gpurun_out/r2_ab2.txt
: > $OUT
for v in ${VARIANTS:-head new new_u8 new_mb10 new_mb12 new_u8_mb6 new_u2 head new}; do
  BENDY2D_B200_LIB=$PWD/ab_libs/$v/libbendy2d_b200.so timeout 200 python profiles/quick_c3.py "$v" | tee -a $OUT
done
if [ -z "$SKIP_TESTS" ]; then timeout 900 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_full_size.py > gpurun_out/r2_ab2_tests.log 2>&1; fi
echo "gpu tests (default build) exit code $?" | tee -a $OUT
tail -3 gpurun_out/r2_ab2_tests.log | tee -a $OUT
