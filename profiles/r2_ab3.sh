# A/B of the link-kernel instruction diet (ab_libs/<name>/): C3 graph substep early / mid / late + the device check of
# the shared-reciprocal normalize against IEEE division
mkdir -p gpurun_out
OUT=gpurun_out/r2_ab3.txt
: > $OUT
[ -n "$SKIP_TESTS" ] || timeout 300 python -m pytest tests/test_gpu_normalize.py -m gpu -q 2>&1 | tail -3 | tee -a $OUT
for v in ${VARIANTS:-head norm k3 head k3}; do
  BENDY2D_B200_LIB=$PWD/ab_libs/$v/libbendy2d_b200.so timeout 200 python profiles/quick_c3.py "$v" 2>&1 | tail -1 | tee -a $OUT
done
