"""-m gpu: BASELINE.json's full size (10^7 cells per grid, configs[3]) through size-independent properties and a
sampled oracle comparison (cells are independent, so the oracle on a subset of the inputs must reproduce the same
subset of the outputs)."""
import numpy as np
import pytest

from oracle_py import Oracle
from tolerances import check_field, check_scenario

pytestmark = pytest.mark.gpu
N = 10_000_000


@pytest.fixture(scope="module")
def full(fcmod):
    from components.flux_calculator_b200 import DeviceArray
    from synthetic import Scenario
    sc = Scenario("CCLM", n=(N, N, N), S=1, bias=True)
    g_in, g_out = sc.clone()
    fc = fcmod.FluxCalculator(sc.n, sc.S)
    wrapped = sc.apply(fc, g_in, g_out, wrap=lambda a: DeviceArray.from_numpy(a))
    for g in (1, 2, 3):
        fc.set_area(g, sc.area[g])
    fc.set_option("diagnostics", 2)
    fc.prepare()
    assert fc.info("spec_kernel") == 1
    fc.exact_calls_before = fc.info("exact_path_calls")      # (a process-wide counter: other tests force that path on purpose)
    fc.step_all(0)
    fc.synchronize()
    fc.exact_calls_after = fc.info("exact_path_calls")
    for k, a in g_out.items():
        wrapped[id(a)].download(a)
    yield fc, sc, g_in, g_out
    fc.close()
    for w in wrapped.values():
        w.free()


def test_sampled_cells_match_the_oracle(full):
    from synthetic import Scenario
    fc, sc, g_in, g_out = full
    idx = np.unique(np.concatenate([np.arange(4096), np.arange(N - 4096, N), np.arange(0, N, 997),
                                    np.arange(511, N, 512 * 296)]))      # tile edges of the persistent schedule included
    small = Scenario("CCLM", n=(idx.size,) * 3, S=1, bias=True)
    o_in, o_out = small.clone()
    done = set()
    for key, arr in o_in.items():
        if id(arr) not in done:
            arr[:] = g_in[key][idx]
            done.add(id(arr))
    small.corrections = np.ascontiguousarray(sc.corrections[idx])
    orc = Oracle(small.n, small.S)
    small.apply(orc, o_in, o_out)
    orc.step_all(0)
    small.inputs = o_in      # the sampled cells of the full grid
    check_scenario(small, {k: g_out[k][idx] for k in o_out}, o_out)


def test_exact_identities_over_all_cells(full):
    fc, sc, g_in, g_out = full
    T = g_in[(1, 1, "TSUR")]
    assert np.array_equal(g_out[(1, 1, "RBBR")], 5.67e-8 * ((T * T) * (T * T)))          # flux_radiation_blackbody.F90:40
    assert np.array_equal(g_out[(1, 1, "RSDR")], g_in[(0, 1, "RSDD")])                    # distribute_radiation_flux.F90:24
    assert np.array_equal(g_out[(1, 1, "HLAT")], g_out[(1, 1, "MEVA")] * 2.501e6)         # heat_latent.F90:65, corrected MEVA
    for g, comp, out in ((2, "UATM", "UMOM"), (3, "VATM", "VMOM")):                        # tau = -flux_air * wind, flux_air > 0
        w, tau = g_in[(0, g, comp)], g_out[(1, g, out)]
        assert np.all(np.sign(tau) == -np.sign(w))
    for key, arr in g_out.items():
        assert np.isfinite(arr).all(), key
    assert fc.exact_calls_after == fc.exact_calls_before          # physical data never leaves the lock-step path


def test_diagnostics_over_all_cells(full):
    fc, sc, g_in, g_out = full
    for (i, g, name), arr in g_out.items():
        s, mn, mx = fc.diagnostics(i, g, name)
        assert mn == arr.min() and mx == arr.max(), (i, g, name)
        ref = float(np.sum(sc.area[g] * arr))
        assert abs(s - ref) <= 1e-11 * float(np.sum(np.abs(sc.area[g] * arr))), (i, g, name)


def test_two_halves_equal_the_whole(fcmod, full):
    """shard invariance at full size: the second half of the grid computed as its own context (what rank 1 of 2 does)
    gives the bits of the whole-grid run"""
    from components.flux_calculator_b200 import DeviceArray
    from synthetic import Scenario
    _, sc, g_in, g_out = full
    off, size = fcmod.shard_range(N, 1, 2, 512)
    half = Scenario("CCLM", n=(size,) * 3, S=1, bias=True, offset=(off,) * 3)
    h_in, h_out = half.clone()
    for key in h_in:
        assert np.array_equal(h_in[key], g_in[key][off:off + size]), key      # counter-based RNG keyed by global cell index
    fc = fcmod.FluxCalculator(half.n, half.S)
    wrapped = half.apply(fc, h_in, h_out, wrap=lambda a: DeviceArray.from_numpy(a))
    fc.prepare()
    fc.step_all(0)
    fc.synchronize()
    for key, a in h_out.items():
        wrapped[id(a)].download(a)
        assert np.array_equal(a, g_out[key][off:off + size]), key
    fc.close()
    for w in wrapped.values():
        w.free()


def test_two_surface_types_full_size(fcmod):
    """BASELINE.json configs[4] at its full size: 10^7 cells per grid, open water + ice with area-fraction averaging, bias;
    1000 consecutive device-resident steps through fc_run_steps.  Oracle on sampled cells (tile edges of the schedule
    included), exact averaging identity over all cells."""
    from components.flux_calculator_b200 import DeviceArray
    from synthetic import Scenario
    sc = Scenario("CCLM", n=(N, N, N), S=2, bias=True, averaging=True)
    g_in, g_out = sc.clone()
    fc = fcmod.FluxCalculator(sc.n, sc.S)
    wrapped = sc.apply(fc, g_in, g_out, wrap=lambda a: DeviceArray.from_numpy(a))
    fc.prepare()
    assert fc.info("spec_kernel") == 1
    nsteps, dt = 1000, 600
    fc.run_steps(0, dt, nsteps)
    fc.synchronize()
    for k, a in g_out.items():
        wrapped[id(a)].download(a)
    fc.close()
    for w in wrapped.values():
        w.free()
    idx = np.unique(np.concatenate([np.arange(4096), np.arange(N - 4096, N), np.arange(0, N, 1999), np.arange(511, N, 512 * 148)]))
    small = Scenario("CCLM", n=(idx.size,) * 3, S=2, bias=True, averaging=True)
    o_in, o_out = small.clone()
    done = set()
    for key, a in o_in.items():
        if id(a) not in done:
            a[:] = g_in[key][idx]
            done.add(id(a))
    small.corrections = np.ascontiguousarray(sc.corrections[idx])
    orc = Oracle(small.n, small.S)
    small.apply(orc, o_in, o_out)
    orc.step_all((nsteps - 1) * dt)
    small.inputs = o_in
    check_scenario(small, {k: g_out[k][idx] for k in o_out}, o_out)
    for (i, g, name) in sc.send:      # average_across_surface_types (calculate.F90:368-385) is exact arithmetic on the per-type values
        acc = np.zeros(N)
        for t in (1, 2):
            acc = acc + g_out[(t, g, name)] * g_in[(t, g, "FARE")]
        assert np.array_equal(acc, g_out[(0, g, name)]), name
