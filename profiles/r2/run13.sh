# round 2, call 13 (2 GPUs): multi-GPU tests (peer mailboxes / NCCL, one and two surface types), C4 at 2 ranks
set -x
export COLUMNS=200
timeout 500 python -m pytest tests/test_gpu_multi.py -m gpu -q -rf --tb=short --timeout 240 -p no:cacheprovider 2>&1 | tail -12
timeout 300 python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --nproc-per-node 2 --master-port 29521 bench.py --gpus 2 --steps 1000 --warmup 20 > gpurun_out/r2_13_c4_n2.json 2> gpurun_out/r2_13_c4_n2.err; cut -c1-300 gpurun_out/r2_13_c4_n2.json; tail -2 gpurun_out/r2_13_c4_n2.err
