"""do_regridding on the device (fc_set_regrid_matrix / fc_regrid, regrid_csr_kernel): throughput on a 10^7-row matrix.
Run on the GPU box:  python profiles/regrid_bench.py > gpurun_out/regrid_bench.json
The matrix mimics a t -> u regridding of the exchange grid: 2 sources per destination cell (neighbouring cells), weights
summing to one, elements shuffled in blocks so that the stable grouping by destination has something to do.
Algorithmic bytes per launch: 8 B (row pointer) + nnz/row * (4 B index + 8 B weight + 8 B gathered source) + 8 B result."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import components.flux_calculator_b200 as m  # noqa: E402
from components.flux_calculator_b200 import DeviceArray  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
rng = np.random.default_rng(1)
dst = np.repeat(np.arange(1, N + 1, dtype=np.int32), 2)
src = np.empty(2 * N, dtype=np.int32)
src[0::2] = np.arange(1, N + 1)
src[1::2] = np.minimum(np.arange(2, N + 2), N)
w = np.empty(2 * N)
w[0::2] = rng.uniform(0.3, 0.7, N)
w[1::2] = 1.0 - w[0::2]
perm = (np.arange(2 * N).reshape(-1, 1024)[rng.permutation(2 * N // 1024)]).ravel() if (2 * N) % 1024 == 0 else np.arange(2 * N)
src, dst, w = src[perm], dst[perm], w[perm]
x = rng.standard_normal(N)
fc = m.FluxCalculator((N, N, N), 1)
t0 = time.perf_counter()
fc.set_regrid_matrix(2, src, dst, w)
setup_s = time.perf_counter() - t0
dx, dy = DeviceArray.from_numpy(x), DeviceArray(N)
for _ in range(3):
    fc.regrid(2, dy, dx)
fc.synchronize()
reps = 50
fc.event_record(0)
for _ in range(reps):
    fc.regrid(2, dy, dx)
fc.event_record(1)
ms = fc.event_elapsed_ms() / reps
y = dy.download()
ref = np.zeros(N)
np.add.at(ref, dst - 1, x[src - 1] * w)      # (numpy's order of accumulation per destination is the element order as well)
bytes_per_launch = N * 8 + 2 * N * (4 + 8 + 8) + N * 8
peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
print(json.dumps({"kernel": "regrid_csr_kernel", "rows": N, "nnz": 2 * N, "ms_per_launch": ms, "algorithmic_bytes": bytes_per_launch,
                  "achieved_gbs": bytes_per_launch / ms / 1e6, "frac_of_measured_hbm_peak": bytes_per_launch / ms / 1e6 / peak,
                  "matrix_setup_s": setup_s, "max_abs_dev_from_numpy": float(np.max(np.abs(y - ref))), "bit_identical_to_numpy": bool(np.array_equal(y, ref))}))
