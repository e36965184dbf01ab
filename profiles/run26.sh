set -x
timeout 1200 python -m pytest tests/test_gpu_step_parity.py -m gpu -x -q -k "diag or two_surface or cold or surface_types or schedules" 2>&1 | tail -12
