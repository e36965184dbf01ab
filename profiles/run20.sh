set -x
nvidia-smi -L | wc -l
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29520 bench.py --gpus 8 --steps 500 --warmup 20 --no-e2e > gpurun_out/bench_n8.json 2> gpurun_out/bench_n8.err; tail -3 gpurun_out/bench_n8.err; cat gpurun_out/bench_n8.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 500 --warmup 20 --no-e2e --no-parity --comm nccl > gpurun_out/bench_n8_nccl.json 2>> gpurun_out/bench_n8.err; cat gpurun_out/bench_n8_nccl.json | cut -c1-400
