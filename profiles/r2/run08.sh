# round 2, call 8: per-CTA chaining of consecutive static steps, DYN as its own instantiation
set -x
export COLUMNS=220
timeout 1700 python -m pytest tests -m gpu -q -rf --tb=short 2>&1 | tail -30
B="timeout 300 python bench.py --no-e2e --no-cpu-baseline --no-parity"
run() { name=$1; shift; "$@" > gpurun_out/r2_08_$name.json 2>>gpurun_out/r2_08.err; cut -c1-160 gpurun_out/r2_08_$name.json; }
run shard $B --workload C4 --cells 1250000 --steps 1000 --warmup 50
run shard_nochain $B --workload C4 --cells 1250000 --steps 1000 --warmup 50 --opt chain=0
run shard_nodiag $B --workload C4 --cells 1250000 --diag 0 --steps 1000 --warmup 50
run shard_nodiag_nochain $B --workload C4 --cells 1250000 --diag 0 --steps 1000 --warmup 50 --opt chain=0
run c4 $B --workload C4
run c4_nochain $B --workload C4 --opt chain=0
run c4_nodiag $B --workload C4 --diag 0
run c3 $B --workload C3 --steps 1000 --warmup 50
run c3_nochain $B --workload C3 --steps 1000 --warmup 50 --opt chain=0
run c5 $B --workload C5
run c2 $B --workload C2 --steps 2000 --warmup 50
tail -5 gpurun_out/r2_08.err
