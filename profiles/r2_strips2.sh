mkdir -p gpurun_out
OUT=gpurun_out/r2_strips2.txt
: > $OUT
for cfg in "" "BENDY_PDL_NCCL=1" $EXTRA; do
  env $cfg timeout 280 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 \
      profiles/strips_timing.py "${cfg:-default}" 2> gpurun_out/r2_strips2.err | grep "rank" | tee -a $OUT
done
timeout 600 python -m pytest tests/test_gpu_strips.py tests/test_z_gpu_strips_replicated.py -m gpu -q -x 2>&1 | tail -3 | tee -a $OUT
