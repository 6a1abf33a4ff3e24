# Round-2 first GPU call: prove and measure the opt-in variants that round 1 could only validate on the
# CPU emulator (tests/cuemu) after its GPU budget was spent.  Everything runs under a timeout so that a
# variant that hangs on the device cannot hold the box.
#   gpurun --timeout 1500 -- 'bash profiles/r2_variants_ab.sh'              (1 GPU part)
#   gpurun --gpus 8 --timeout 900 -- 'bash profiles/r2_variants_ab.sh strips 8'   (strips part)
# Output: gpurun_out/r2_ab.txt
mkdir -p gpurun_out
OUT=gpurun_out/r2_ab.txt
if [ "$1" = "strips" ]; then
  N=${2:-8}
  for cfg in "" "BENDY_SCAN_MT=1" "BENDY_HALO_FUSED=1" "BENDY_PDL_NCCL=1" "BENDY_NARROW_DENSE=1" "BENDY_SCATTER_ILP=1" \
             "BENDY_SCAN_MT=1 BENDY_HALO_FUSED=1 BENDY_PDL_NCCL=1" \
             "BENDY_SCAN_MT=1 BENDY_HALO_FUSED=1 BENDY_PDL_NCCL=1 BENDY_NARROW_DENSE=1"; do
    label=$(echo "strips${N}_${cfg:-default}" | tr ' =' '__')
    env $cfg timeout 280 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
        --master-port 29517 bench.py --gpus $N --steps 12 --warmup 3 --no-e2e > gpurun_out/$label.json 2> gpurun_out/$label.err
    python - "$label" >> $OUT <<'PY'
import json, sys
try:
    d = json.loads(open(f'gpurun_out/{sys.argv[1]}.json').read().strip().splitlines()[-1])
    print(f"{sys.argv[1]:90s} {d['value']:.4e} particle-substeps/s  {d['ms_per_step'] / 8 * 1000:7.1f} us/substep")
except Exception as e:
    print(f"{sys.argv[1]:90s} FAILED ({e})")
PY
  done
  cat $OUT
  exit 0
fi
# 1. the variants' own parity tests on the device (default gpu run skips them until this has passed once)
BENDY_TEST_UNPROVEN=1 timeout 600 python -m pytest tests/test_z_gpu_variants.py -m gpu -x -q > gpurun_out/r2_variant_tests.log 2>&1
echo "variant tests exit code $?" | tee -a $OUT
tail -3 gpurun_out/r2_variant_tests.log | tee -a $OUT
# 2. C3 graph-mode substep time early / mid / late per switch
for cfg in "" "BENDY_NARROW_DENSE=1" "BENDY_SCATTER_ILP=1" "BENDY_SORT_FUSED=1" "BENDY_SORT_FUSED=1 BENDY_NARROW_DENSE=1"; do
  env $cfg timeout 200 python profiles/quick_c3.py "C3 ${cfg:-default}" | tee -a $OUT
done
# 3. the 2M-disc-per-rank strip problem of the 8-GPU run on ONE GPU (same kernels, no exchange partner): grid build
for cfg in "" "BENDY_SCAN_MT=1" "BENDY_SCAN_MT=1 BENDY_SCATTER_ILP=1" "BENDY_SCAN_MT=1 BENDY_NARROW_DENSE=1"; do
  env $cfg timeout 300 python profiles/strip_rank_kernels.py "rank-of-8 ${cfg:-default}" | tee -a $OUT
done
# 4. compute-sanitizer on the variants (small scenes)
for tool in memcheck racecheck; do
  BENDY_TEST_UNPROVEN=1 timeout 900 compute-sanitizer --tool $tool python -m pytest tests/test_z_gpu_variants.py -m gpu -x -q \
      -k "pool_flushes or 768 or strip_variants" > gpurun_out/r2_sanitizer_$tool.log 2>&1
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed" gpurun_out/r2_sanitizer_$tool.log | tee -a $OUT
done
# 5. refresh the launch list of the default path (profiles/r1_launches_* predate the fused polygon prepare kernel) and
#    one full capture of the main-chain kernels; numbers printed under ncu are never bench values
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 208 --csv --log-file gpurun_out/r2_launches_bench_steps2.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-scaling-ref > gpurun_out/r2_ncu_bench.log 2>&1
python profiles/summarize_launches.py gpurun_out/r2_launches_bench_steps2.csv > gpurun_out/r2_launches_summary.txt 2>&1
tail -20 gpurun_out/r2_launches_summary.txt | tee -a $OUT
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k2_narrow|k3_links_local|k2_scan|k2_scatter" -s 40 -c 12 \
    -o gpurun_out/r2_full python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-scaling-ref > gpurun_out/r2_ncu_full.log 2>&1
