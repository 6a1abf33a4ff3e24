# C5 strong scaling with the final build: bench.py --gpus N (parity_check included), N given as $1
mkdir -p gpurun_out
N=$1
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2_bench_c5_n$N.json 2> gpurun_out/r2_bench_c5_n$N.err
echo "rc=$?"; grep parity_check gpurun_out/r2_bench_c5_n$N.err | head -2; python profiles/show_bench.py gpurun_out/r2_bench_c5_n$N.json 2>/dev/null | head -8
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus $N --steps 20 --warmup 5 > gpurun_out/r2_bench_reference_c5_n$N.json 2> /dev/null; tail -c 400 gpurun_out/r2_bench_reference_c5_n$N.json
