set -x
nvidia-smi -L
time python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 50 --warmup 5 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
tail -5 gpurun_out/bench_n2.err
cat gpurun_out/bench_n2.json
time python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 200 --warmup 10 --no-e2e --no-parity > gpurun_out/bench_n2b.json 2>> gpurun_out/bench_n2.err
cat gpurun_out/bench_n2b.json
timeout 600 python -m pytest tests -m gpu -x -q -k "spec or diag or ragged" 2>&1 | tail -5
