set -x
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -15
for comm in p2p nccl; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 300 --warmup 10 --no-e2e --no-parity --comm $comm 2>gpurun_out/n2_$comm.err | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']; print('N2 $comm', 'ms/step', round(d['ms_per_step'],4), 'kernel_ms', round(r['kernel_ms'],4), 'frac', round(r['frac'],3), 'launches', d['gpu_launches'], d['config']['diagnostics_exchange'], d['diagnostics_sample'])
    else: print(l.rstrip()[:300])
"
tail -3 gpurun_out/n2_$comm.err
done
# emulate the 8-GPU shard size on 2 GPUs: 2.5e6 cells over 2 ranks = 1.25e6 per rank
for comm in p2p nccl; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 2 --cells 2500000 --steps 500 --warmup 10 --no-e2e --no-parity --comm $comm 2>>gpurun_out/n2_$comm.err | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']; print('N2 small $comm', 'ms/step', round(d['ms_per_step'],4), 'kernel_ms', round(r['kernel_ms'],4), 'frac', round(r['frac'],3), 'launches', d['gpu_launches'])
    else: print(l.rstrip()[:300])
"
done
