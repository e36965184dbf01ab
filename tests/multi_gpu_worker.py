"""One rank of the multi-GPU diagnostics check (launched by tests/test_gpu_multi.py or by hand under torchrun):
shard the synthetic grid (fc_shard_range), run fused steps with diagnostics level 2 on this rank's GPU, exchange the
diagnostics (peer mailboxes or NCCL) and compare the global values with numpy over the gathered outputs."""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--comm", default="p2p", choices=["p2p", "nccl"])
    ap.add_argument("--cells", type=int, default=700_001)
    ap.add_argument("--fset", default="CCLM")
    ap.add_argument("--S", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    args = ap.parse_args()
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    import torch.distributed as dist
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group(backend="gloo", rank=rank, world_size=world)
    import components.flux_calculator_b200 as m
    from components.flux_calculator_b200 import DeviceArray
    from synthetic import Scenario

    off, size = m.shard_range(args.cells, rank, world, 512)
    sc = Scenario(args.fset, n=(size, size, size), S=args.S, bias=True, averaging=True, offset=(off, off, off))
    g_in, g_out = sc.clone()
    fc = m.FluxCalculator(sc.n, sc.S, device=rank)
    wrapped = sc.apply(fc, g_in, g_out, wrap=lambda a: DeviceArray.from_numpy(a, rank))
    for g in (1, 2, 3):
        fc.set_area(g, sc.area[g])
    fc.set_option("diagnostics", 2)
    fc.set_option("staged", 2)
    fc.prepare()
    if args.comm == "p2p":
        hs = [None] * world
        dist.all_gather_object(hs, fc.comm_p2p_handle())
        fc.comm_p2p_connect(hs, rank, world)
    else:
        uid = [m.comm_get_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        fc.comm_init(uid[0], rank, world)
    for k in range(args.steps):
        fc.step_all(600 * k)
        fc.allreduce_diagnostics()
        if k % 2 == 0:      # read some steps, skip others (double-buffered records)
            fc.diagnostics(1, 1, "HSEN")
    fc.synchronize()
    worst = 0.0
    for key in sorted(g_out):
        arr = wrapped[id(g_out[key])].download()
        if arr.size == 0:
            continue
        loc = np.array([float(np.sum(sc.area[key[1]] * arr)), float(np.sum(np.abs(sc.area[key[1]] * arr))), arr.min(), arr.max()])
        allv = [None] * world
        dist.all_gather_object(allv, loc)
        s, mn, mx = fc.diagnostics(*key)
        ref_s = sum(v[0] for v in allv)
        ref_abs = sum(v[1] for v in allv)
        assert mn == min(v[2] for v in allv) and mx == max(v[3] for v in allv), (key, mn, mx)
        assert abs(s - ref_s) <= 1e-11 * ref_abs, (key, s, ref_s)
        worst = max(worst, abs(s - ref_s) / max(ref_abs, 1e-300))
        # every rank must hold the same bits
        same = [None] * world
        dist.all_gather_object(same, (s, mn, mx))
        if args.comm == "p2p":
            assert all(x == same[0] for x in same), (key, same)
    if rank == 0:
        print("multi-gpu diagnostics ok: comm=%s ranks=%d fields=%d worst rel err %.3g" % (args.comm, world, len(g_out), worst))
    dist.barrier()
    fc.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
