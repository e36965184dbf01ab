# round 2, call 9: which GPU test is slow / hangs?  per-test timeout, durations
set -x
export COLUMNS=200
timeout 900 python -m pytest tests -m gpu -q -rf --tb=short --timeout 150 --durations=25 -p no:cacheprovider > gpurun_out/r2_09_pytest.log 2>&1
tail -70 gpurun_out/r2_09_pytest.log
