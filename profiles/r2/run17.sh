# round 2, call 17 (2 GPUs): the two-rank tests (peer mailboxes / NCCL, one and two surface types) and the scaling job's
# 2-rank line (20 steps after 5 warm-up steps) on the library and bench.py as shipped
set -x
export COLUMNS=200
timeout 150 python -m pytest tests/test_gpu_multi.py -m gpu -q -rf --tb=short --timeout 120 -p no:cacheprovider -x 2>&1 | tail -6
timeout 60 python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --nproc-per-node 2 --master-port 29531 bench.py --gpus 2 --steps 20 --warmup 5 --no-e2e > gpurun_out/r2_17_c4_n2_short.json 2> gpurun_out/r2_17_c4_n2_short.err; cut -c1-300 gpurun_out/r2_17_c4_n2_short.json; tail -2 gpurun_out/r2_17_c4_n2_short.err
