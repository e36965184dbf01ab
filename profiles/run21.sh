export FC_BENCH_PER_RANK=1
for v in "FC_BENCH_NO_SAMPLER=0 --comm p2p" "FC_BENCH_NO_SAMPLER=1 --comm p2p" "FC_BENCH_NO_SAMPLER=1 --comm p2p --diag 0"; do
  env ${v%% *} python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29530 bench.py --gpus 2 --cells 2500000 --steps 1000 --warmup 20 --no-e2e --no-parity ${v#* } 2>gpurun_out/r21.err | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']; print('$v', 'ms/step', round(d['ms_per_step'],4), 'kernel_ms', round(r['kernel_ms'],4))
"
  grep "^rank" gpurun_out/r21.err
done
FC_BENCH_NO_SAMPLER=1 python bench.py --cells 1250000 --steps 1000 --warmup 20 --no-e2e --no-parity --no-cpu-baseline 2>&1 | grep "^rank"
python bench.py --cells 1250000 --steps 1000 --warmup 20 --no-e2e --no-parity --no-cpu-baseline 2>&1 | grep "^rank"
