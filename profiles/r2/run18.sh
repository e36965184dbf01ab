# round 2, call 18 (1 GPU): ncu capture of the generic fused kernel (three surface types, 4*10^6 cells per grid) after its
# move to the flag-and-recompute scheme
set -x
O=gpurun_out
B="timeout 100 python bench.py --no-e2e --no-cpu-baseline --no-parity --workload S3 --cells 4000000 --steps 3 --warmup 3 --profile-steps 0"
$B > $O/r2_18_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:fused_step_kernel -s 3 -c 1 -f -o $O/r2_18_prof_s3 $B > $O/r2_18_ncu.log 2>&1
cut -c1-200 $O/r2_18_plain.log; tail -3 $O/r2_18_ncu.log
