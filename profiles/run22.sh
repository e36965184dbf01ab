export FC_BENCH_PER_RANK=1
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -3
for v in "--comm p2p" "--comm nccl"; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29530 bench.py --gpus 2 --cells 2500000 --steps 1000 --warmup 20 --no-e2e --no-parity $v 2>gpurun_out/r22.err | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']; print('$v', 'ms/step', round(d['ms_per_step'],4), 'kernel_ms', round(r['kernel_ms'],4), d['diagnostics_sample'])
"
  grep "^rank" gpurun_out/r22.err
done
