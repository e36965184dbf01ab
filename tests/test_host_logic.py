"""not-gpu: host-side logic around the path -- synthetic grid shard invariance, the multi-rank shard rule and the
diagnostics reduction semantics exercised with a world_size-2 gloo process group on CPU (the arithmetic in these
tests is the ORACLE's; the CUDA path is covered by the -m gpu tests)."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_synthetic_fields_are_shard_invariant():
    from synthetic import Scenario
    full = Scenario("CCLM", n=(4096, 4096, 4096), S=2, bias=True, averaging=True)
    part = Scenario("CCLM", n=(1024, 1024, 1024), S=2, bias=True, averaging=True, offset=(2048, 2048, 2048))
    for k, a in part.inputs.items():
        assert np.array_equal(a, full.inputs[k][2048:3072]), k
    assert np.array_equal(part.corrections, full.corrections[2048:3072])
    assert np.array_equal(part.area[1], full.area[1][2048:3072])


def test_scenario_aliasing_mirrors_distribute_input_field():
    from synthetic import Scenario
    sc = Scenario("CCLM", n=(64, 64, 64), S=3)
    for g in (1, 2, 3):
        assert sc.inputs[(0, g, "PSUR")] is sc.inputs[(1, g, "PSUR")] is sc.inputs[(3, g, "PSUR")]   # basic.F90:349
    assert sc.inputs[(1, 1, "TSUR")] is not sc.inputs[(2, 1, "TSUR")]
    ins, outs = sc.clone()
    assert ins[(0, 1, "PSUR")] is ins[(2, 1, "PSUR")] and ins[(0, 1, "PSUR")] is not sc.inputs[(0, 1, "PSUR")]


def test_oracle_ranks_equal_single_rank():
    """P independent ranks over contiguous ranges == one rank (the reference's only parallelism)"""
    from synthetic import Scenario
    from oracle_py import Oracle
    sc = Scenario("MOM5", n=(5001, 4999, 5003), S=2, bias=True, averaging=True)
    res = []
    for P in (1, 3, 8):
        ins, outs = sc.clone()
        o = Oracle(sc.n, sc.S)
        sc.apply(o, ins, outs)
        o.run_ranks(P, 1, 600, 40 * 86400)
        res.append(outs)
    for k in res[0]:
        assert np.array_equal(res[0][k], res[1][k], equal_nan=True) and np.array_equal(res[0][k], res[2][k], equal_nan=True), k


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _rank_main(rank, world, port, n_total, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist
    import components.flux_calculator_b200 as m
    from synthetic import Scenario
    from oracle_py import Oracle
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    off, size = m.shard_range(n_total, rank, world, 512)
    sc = Scenario("CCLM", n=(size, size, size), S=1, bias=True, offset=(off, off, off))
    ins, outs = sc.clone()
    o = Oracle(sc.n, sc.S)
    sc.apply(o, ins, outs)
    o.step_all(0)
    # diagnostics: local (sum area*x, min, max), then the same reduction the library does with NCCL
    x = outs[(1, 1, "HSEN")]
    loc = torch.tensor([float(np.sum(sc.area[1] * x))], dtype=torch.float64)
    mn = torch.tensor([x.min()], dtype=torch.float64)
    mx = torch.tensor([x.max()], dtype=torch.float64)
    dist.all_reduce(loc, op=dist.ReduceOp.SUM)
    dist.all_reduce(mn, op=dist.ReduceOp.MIN)
    dist.all_reduce(mx, op=dist.ReduceOp.MAX)
    gathered = [None] * world
    dist.all_gather_object(gathered, (off, size, x))
    if rank == 0:
        q.put((float(loc), float(mn), float(mx), gathered))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_shards_reproduce_the_unsharded_grid():
    import torch.multiprocessing as mp
    from synthetic import Scenario
    from oracle_py import Oracle
    n_total, world = 20_000, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_rank_main, args=(r, world, port, n_total, q)) for r in range(world)]
    for p in procs:
        p.start()
    s, mn, mx, gathered = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    full = Scenario("CCLM", n=(n_total,) * 3, S=1, bias=True)
    ins, outs = full.clone()
    o = Oracle(full.n, 1)
    full.apply(o, ins, outs)
    o.step_all(0)
    ref = outs[(1, 1, "HSEN")]
    cat = np.concatenate([g[2] for g in sorted(gathered, key=lambda t: t[0])])
    assert np.array_equal(cat, ref)                       # shard invariance, bit for bit
    assert mn == ref.min() and mx == ref.max()
    tot = float(np.sum(full.area[1] * ref))
    assert abs(s - tot) <= 1e-12 * float(np.sum(np.abs(full.area[1] * ref)))
