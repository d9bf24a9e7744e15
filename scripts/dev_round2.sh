set -x
mkdir -p gpurun_out
timeout -s KILL 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_scale_8gpu.json 2> gpurun_out/bench_scale_8gpu.err
cat gpurun_out/bench_scale_8gpu.json | cut -c1-700; tail -3 gpurun_out/bench_scale_8gpu.err
