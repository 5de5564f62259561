set -x
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511"
$TR bench.py --gpus 8 --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r2_8gpu_default.json 2> gpurun_out/bench_r2_8gpu_default.err
cut -c1-330 gpurun_out/bench_r2_8gpu_default.json
$TR bench.py --gpus 8 --steps 4 --warmup 3 --no-cpu-baseline --ddp-bucket-mb 100 --ddp-bf16 1 > gpurun_out/bench_r2_8gpu_b100_bf16.json 2> gpurun_out/bench_r2_8gpu_b100_bf16.err
cut -c1-330 gpurun_out/bench_r2_8gpu_b100_bf16.json
$TR bench.py --gpus 8 --steps 4 --warmup 3 --no-cpu-baseline --ddp-no-sync > gpurun_out/bench_r2_8gpu_nosync.json 2> gpurun_out/bench_r2_8gpu_nosync.err
cut -c1-330 gpurun_out/bench_r2_8gpu_nosync.json
$TR scripts/bench_configs.py --which 5 --samples 50 --iters 1 > gpurun_out/config5_8gpu_S50.json 2> gpurun_out/config5_8gpu_S50.err
cat gpurun_out/config5_8gpu_S50.json; tail -2 gpurun_out/config5_8gpu_S50.err
