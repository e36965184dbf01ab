# round 2, call 11 (8 GPUs): C4 / C5 / C3 sharded over 8 ranks (strong scaling), 8 concurrent PCIe probes, topology
set -x
nvidia-smi topo -m > gpurun_out/r2_11_topo.txt 2>&1
T="timeout 400 python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --nproc-per-node 8"
$T --master-port 29511 bench.py --gpus 8 --steps 1000 --warmup 20 > gpurun_out/r2_11_c4_n8.json 2> gpurun_out/r2_11_c4_n8.err; cut -c1-300 gpurun_out/r2_11_c4_n8.json; tail -3 gpurun_out/r2_11_c4_n8.err
$T --master-port 29512 bench.py --gpus 8 --steps 1000 --warmup 20 --no-e2e --opt chain=0 > gpurun_out/r2_11_c4_n8_nochain.json 2> gpurun_out/r2_11_c4_n8_nochain.err; cut -c1-300 gpurun_out/r2_11_c4_n8_nochain.json
$T --master-port 29513 bench.py --gpus 8 --workload C5 --no-e2e > gpurun_out/r2_11_c5_n8.json 2> gpurun_out/r2_11_c5_n8.err; cut -c1-300 gpurun_out/r2_11_c5_n8.json; tail -3 gpurun_out/r2_11_c5_n8.err
$T --master-port 29514 bench.py --gpus 8 --workload C3 --steps 1000 --warmup 20 --no-e2e > gpurun_out/r2_11_c3_n8.json 2> gpurun_out/r2_11_c3_n8.err; cut -c1-300 gpurun_out/r2_11_c3_n8.json
for i in 0 1 2 3 4 5 6 7; do CUDA_VISIBLE_DEVICES=$i timeout 120 python profiles/pcie_probe.py > gpurun_out/r2_11_pcie_$i.txt 2>&1 & done; wait
cat gpurun_out/r2_11_pcie_*.txt | tail -24
