"""ctypes wrapper of the CPU oracle (oracle/liboracle.so).  TEST INFRASTRUCTURE ONLY: the checker for the
CUDA path, never part of it.  Mirrors the FluxCalculator interface so parity tests bind one Scenario to both."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")

IDX = {n: i + 1 for i, n in enumerate(
    "ALBE ALBA AMOI AMOM FARE FICE PATM PSUR QATM TATM TSUR UATM VATM U10M V10M CMOM CMOI CHEA QSUR HLAT HSEN "
    "MEVA MPRE MRAI MSNO RBBR RLWD RLWU RSID RSIU RSIN RSDD RSDR UMOM VMOM".split())}


def load(fast=False, path=None):
    name = "liboracle_fast.so" if fast else "liboracle.so"
    path = path or os.path.join(ORACLE_DIR, name)
    if not os.path.exists(path):
        subprocess.check_call(["make", "-C", ORACLE_DIR], stdout=subprocess.DEVNULL)
    lib = C.CDLL(path)
    lib.orc_state_new.restype = C.c_void_p
    lib.orc_state_new.argtypes = [C.c_int, C.POINTER(C.c_int64)]
    lib.orc_state_free.argtypes = [C.c_void_p]
    lib.orc_bind.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int]
    lib.orc_set_method.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.c_int, C.c_char_p]
    lib.orc_set_corrections.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int]
    lib.orc_set_time.argtypes = [C.c_void_p, C.c_int64]
    lib.orc_set_distribute_shortwave.argtypes = [C.c_void_p, C.c_int]
    lib.orc_add_output.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
    lib.orc_run_ranks.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int64]
    lib.orc_current_month.argtypes = [C.c_int, C.c_int64]
    lib.orc_decomp_apple.argtypes = [C.c_int64, C.c_int, C.c_int, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
    for f in ("orc_calc_spec_vapor_surface", "orc_calc_flux_momentum_east", "orc_calc_flux_momentum_north"):
        getattr(lib, f).argtypes = [C.c_void_p, C.c_int]
    for f in ("orc_calc_flux_mass_evap", "orc_calc_flux_heat_latent", "orc_calc_flux_heat_sensible",
              "orc_calc_flux_radiation_blackbody", "orc_distribute_shortwave_radiation_flux", "orc_step_early",
              "orc_step_normal"):
        getattr(lib, f).argtypes = [C.c_void_p]
    lib.orc_average_across_surface_types.argtypes = [C.c_void_p, C.c_int, C.c_int]
    return lib


_WHICH = {"which_spec_vapor_surface_t": ("QSUR", 1), "which_spec_vapor_surface_u": ("QSUR", 2),
          "which_spec_vapor_surface_v": ("QSUR", 3), "which_flux_mass_evap": ("MEVA", 0),
          "which_flux_heat_latent": ("HLAT", 0), "which_flux_heat_sensible": ("HSEN", 0),
          "which_flux_momentum": ("MOM", 0), "which_flux_radiation_blackbody": ("RBBR", 0)}


class Oracle:
    def __init__(self, grid_size, num_surface_types=1, fast=False, lib_path=None):
        self.lib = load(fast, lib_path)
        gs = (C.c_int64 * 3)(*[int(x) for x in grid_size])
        self.s = self.lib.orc_state_new(int(num_surface_types), gs)
        self.grid_size = tuple(int(x) for x in grid_size)
        self._keep = []
        self._bound = {}
        self._alloc = {}      # explicit %allocated flags (set_allocated)

    def __del__(self):
        try:
            self.lib.orc_state_free(self.s)
        except Exception:
            pass

    def bind_field(self, surface_type, grid, var, array):
        idx = IDX[var] if isinstance(var, str) else var
        if array is None:
            self.lib.orc_bind(self.s, surface_type, grid, idx, None, 0)
            return
        assert array.dtype == np.float64 and array.flags.c_contiguous
        # allocated <=> a type-0 slot whose array no surface type >= 1 shares (same rule as the C ABI)
        self._bound[(surface_type, grid, idx)] = array
        self._keep.append(array)
        self.lib.orc_bind(self.s, surface_type, grid, idx, array.ctypes.data, 1)
        self._fix_allocated()

    def set_allocated(self, surface_type, grid, var, allocated):
        idx = IDX[var] if isinstance(var, str) else var
        if allocated is None:
            self._alloc.pop((surface_type, grid, idx), None)
        else:
            self._alloc[(surface_type, grid, idx)] = int(bool(allocated))
        self._fix_allocated()

    def _fix_allocated(self):
        for (i, g, idx), a in self._bound.items():
            own = 1
            if (i, g, idx) in self._alloc:
                own = self._alloc[(i, g, idx)]
            elif i == 0:
                for (ii, gg, vv), b in self._bound.items():
                    if ii >= 1 and b is a:
                        own = 0
            self.lib.orc_bind(self.s, i, g, idx, a.ctypes.data, own)

    def set_method(self, which, surface_type, method):
        q, g = _WHICH[which]
        assert self.lib.orc_set_method(self.s, q.encode(), g, surface_type, method.encode()) == 0

    def set_distribute_shortwave(self, on):
        self.lib.orc_set_distribute_shortwave(self.s, int(bool(on)))

    def set_corrections(self, corrections, enabled=True, init_date=19610101):
        corrections = np.ascontiguousarray(corrections, dtype=np.float64)
        self._keep.append(corrections)
        self.lib.orc_set_corrections(self.s, corrections.ctypes.data, int(bool(enabled)), int(init_date))

    def add_output_field(self, surface_type, grid, var):
        self.lib.orc_add_output(self.s, surface_type, grid, IDX[var] if isinstance(var, str) else var)

    def set_time(self, t):
        self.lib.orc_set_time(self.s, int(t))

    def calc_spec_vapor_surface(self, g): self.lib.orc_calc_spec_vapor_surface(self.s, g)
    def calc_flux_mass_evap(self): self.lib.orc_calc_flux_mass_evap(self.s)
    def calc_flux_heat_latent(self): self.lib.orc_calc_flux_heat_latent(self.s)
    def calc_flux_heat_sensible(self): self.lib.orc_calc_flux_heat_sensible(self.s)
    def calc_flux_momentum_east(self, g=2): self.lib.orc_calc_flux_momentum_east(self.s, g)
    def calc_flux_momentum_north(self, g=3): self.lib.orc_calc_flux_momentum_north(self.s, g)
    def calc_flux_radiation_blackbody(self): self.lib.orc_calc_flux_radiation_blackbody(self.s)
    def distribute_shortwave_radiation_flux(self): self.lib.orc_distribute_shortwave_radiation_flux(self.s)
    def average_across_surface_types(self, g, var):
        self.lib.orc_average_across_surface_types(self.s, g, IDX[var] if isinstance(var, str) else var)

    # direction: 0 = u->t, 1 = v->t, 2 = t->u, 3 = t->v (same numbering as fc_set_regrid_matrix)
    def set_regrid_matrix(self, direction, src_index, dst_index, weight):
        s = np.ascontiguousarray(src_index, dtype=np.int32)
        d = np.ascontiguousarray(dst_index, dtype=np.int32)
        w = np.ascontiguousarray(weight, dtype=np.float64)
        if not hasattr(self, "_mat"):
            self._mat = {}
        self._mat[direction] = (s, d, w)

    def regrid(self, direction, dst, src):
        class M(C.Structure):
            _fields_ = [("num_elements", C.c_int64), ("src_index", C.c_void_p), ("dst_index", C.c_void_p), ("weight", C.c_void_p)]
        s, d, w = self._mat[direction]
        m = M(s.size, s.ctypes.data, d.ctypes.data, w.ctypes.data)
        self.lib.orc_regrid.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]
        self.lib.orc_regrid.restype = None
        self.lib.orc_regrid(dst.ctypes.data, dst.size, src.ctypes.data, C.byref(m))

    def step_early(self, t=0):
        self.set_time(t)
        self.lib.orc_step_early(self.s)

    def step_normal(self, t=0):
        self.set_time(t)
        self.lib.orc_step_normal(self.s)

    def step_all(self, t=0):
        self.step_early(t)
        self.step_normal(t)

    def run_ranks(self, npes, nsteps=1, timestep=600, t0=0):
        self.set_time(t0)
        assert self.lib.orc_run_ranks(self.s, int(npes), int(nsteps), int(timestep)) == 0


def ulp_diff(a, b):
    """element-wise distance in units in the last place between float64 arrays (0 only for bit-equal values, with
    +0 == -0; both NaN -> 0).  Exact integer arithmetic on the ordered bit patterns (a float64 detour would lose the
    low 9 bits of the 63-bit patterns and hide differences of up to 255 ulp)."""
    a = np.ascontiguousarray(a, dtype=np.float64)
    b = np.ascontiguousarray(b, dtype=np.float64)
    ia = a.view(np.int64).copy()
    ib = b.view(np.int64).copy()
    ia = np.where(ia < 0, np.int64(-2**63) - ia, ia)      # negative floats: mirror so that the integers are ordered like the reals
    ib = np.where(ib < 0, np.int64(-2**63) - ib, ib)
    hi, lo = np.maximum(ia, ib), np.minimum(ia, ib)
    d = (hi.astype(np.uint64) - lo.astype(np.uint64)).astype(np.float64)      # exact below 2^53, which is all we care about
    both_nan = np.isnan(a) & np.isnan(b)
    return np.where(both_nan, 0.0, d)