"""Host-side mirror of the reference's `module flux_library` (flux_lib/flux_library.F90:32-45).

Same 14 public names, same argument order as the Fortran dummies; arguments are arrays instead of
scalars (NumPy float64 arrays on the host or `DeviceArray`s); the trailing keyword arguments are the
Fortran OPTIONAL constants (None == not PRESENT).  Every call runs the CUDA kernel behind the C ABI
(include/fluxcalc.h); nothing is computed in Python.
"""
import ctypes as C

import numpy as np

from ._lib import lib, check
from .memory import DeviceArray, addr_of

__all__ = [
    "flux_heat_latent_ice", "flux_heat_latent_water", "flux_heat_sensible_cclm", "flux_heat_sensible_mom5",
    "flux_heat_sensible_rco", "flux_mass_evap_cclm", "flux_mass_evap_mom5", "flux_mass_evap_rco",
    "flux_momentum_cclm", "flux_momentum_mom5", "flux_momentum_rco", "flux_radiation_blackbody_StBo",
    "distribute_radiation_flux", "spec_vapor_surface_cclm",
]


def _n(*arrays):
    n = None
    for a in arrays:
        if a is None:
            continue
        m = a.n if isinstance(a, DeviceArray) else int(np.asarray(a).size)
        if n is None:
            n = m
        elif n != m:
            raise ValueError("array length mismatch: %d vs %d" % (n, m))
    return n


def _opt(x):
    return None if x is None else C.byref(C.c_double(float(x)))


def _in(a):
    if isinstance(a, DeviceArray):
        return a, a.ptr
    arr = np.ascontiguousarray(a, dtype=np.float64)
    return arr, arr.ctypes.data


def _out(a):
    if a is None:
        return None, None
    if isinstance(a, DeviceArray):
        return a, a.ptr
    if not (isinstance(a, np.ndarray) and a.dtype == np.float64 and a.flags.c_contiguous and a.flags.writeable):
        raise TypeError("result arrays must be writable contiguous float64 NumPy arrays or DeviceArrays")
    return a, a.ctypes.data


def _call(fn, outs, ins, opts=(), stream=None):
    keep = []
    ptrs = []
    for o in outs:
        k, p = _out(o)
        keep.append(k)
        ptrs.append(p)
    for i in ins:
        k, p = _in(i) if i is not None else (None, None)
        keep.append(k)
        ptrs.append(p)
    n = _n(*[k for k in keep if k is not None])
    check(fn(*ptrs, n, *[_opt(o) for o in opts], stream))


def spec_vapor_surface_cclm(specific_vapor_content_surface, fraction_ice, pressure_surface, temperature_surface,
                            gas_constant_air_new=None, gas_constant_vapor_new=None, stream=None):
    """flux_lib/auxiliaries/flux_aux_vapor.F90:20-70"""
    _call(lib.fc_spec_vapor_surface_cclm, [specific_vapor_content_surface],
          [fraction_ice, pressure_surface, temperature_surface], [gas_constant_air_new, gas_constant_vapor_new], stream)


def flux_mass_evap_cclm(flux_mass_evap, diffusion_coefficient_moisture, pressure_surface, specific_vapor_content_atmos,
                        specific_vapor_content_surface, temperature_surface, u_atmos, v_atmos, u_min_evap_new=None,
                        gas_constant_air_new=None, gas_constant_vapor_new=None, stream=None):
    """flux_lib/mass/flux_mass_evap.F90:22-85"""
    _call(lib.fc_flux_mass_evap_cclm, [flux_mass_evap],
          [diffusion_coefficient_moisture, pressure_surface, specific_vapor_content_atmos,
           specific_vapor_content_surface, temperature_surface, u_atmos, v_atmos],
          [u_min_evap_new, gas_constant_air_new, gas_constant_vapor_new], stream)


def flux_mass_evap_mom5(flux_mass_evap, diffusion_coefficient_moisture, pressure_surface, specific_vapor_content_atmos,
                        specific_vapor_content_surface, temperature_surface, u_atmos, v_atmos, stream=None):
    """flux_lib/mass/flux_mass_evap.F90:87-118"""
    _call(lib.fc_flux_mass_evap_mom5, [flux_mass_evap],
          [diffusion_coefficient_moisture, pressure_surface, specific_vapor_content_atmos,
           specific_vapor_content_surface, temperature_surface, u_atmos, v_atmos], [], stream)


def flux_mass_evap_rco(flux_mass_evap, specific_vapor_content_atmos, temperature_surface, u_atmos, v_atmos, stream=None):
    """flux_lib/mass/flux_mass_evap.F90:120-158"""
    _call(lib.fc_flux_mass_evap_rco, [flux_mass_evap],
          [specific_vapor_content_atmos, temperature_surface, u_atmos, v_atmos], [], stream)


def flux_heat_latent_ice(flux_heat_latent, flux_mass_evap, latent_heat_sublimation_new=None, stream=None):
    """flux_lib/heat/flux_heat_latent.F90:23-43"""
    _call(lib.fc_flux_heat_latent_ice, [flux_heat_latent], [flux_mass_evap], [latent_heat_sublimation_new], stream)


def flux_heat_latent_water(flux_heat_latent, flux_mass_evap, latent_heat_vaporization_new=None, stream=None):
    """flux_lib/heat/flux_heat_latent.F90:47-67"""
    _call(lib.fc_flux_heat_latent_water, [flux_heat_latent], [flux_mass_evap], [latent_heat_vaporization_new], stream)


def flux_heat_sensible_cclm(flux_heat_sensible, diffusion_coefficient_moisture, pressure_atmos, pressure_surface,
                            specific_vapor_content_surface, temperature_atmos, temperature_surface, u_atmos, v_atmos,
                            heat_capacity_air_new=None, u_min_evap_new=None, gas_constant_air_new=None,
                            gas_constant_vapor_new=None, stream=None):
    """flux_lib/heat/flux_heat_sensible.F90:24-99"""
    _call(lib.fc_flux_heat_sensible_cclm, [flux_heat_sensible],
          [diffusion_coefficient_moisture, pressure_atmos, pressure_surface, specific_vapor_content_surface,
           temperature_atmos, temperature_surface, u_atmos, v_atmos],
          [heat_capacity_air_new, u_min_evap_new, gas_constant_air_new, gas_constant_vapor_new], stream)


def flux_heat_sensible_mom5(flux_heat_sensible, diffusion_coefficient_moisture, pressure_atmos, pressure_surface,
                            specific_vapor_content_surface, temperature_atmos, temperature_surface, u_atmos, v_atmos,
                            stream=None):
    """flux_lib/heat/flux_heat_sensible.F90:101-135"""
    _call(lib.fc_flux_heat_sensible_mom5, [flux_heat_sensible],
          [diffusion_coefficient_moisture, pressure_atmos, pressure_surface, specific_vapor_content_surface,
           temperature_atmos, temperature_surface, u_atmos, v_atmos], [], stream)


def flux_heat_sensible_rco(flux_heat_sensible, temperature_atmos, temperature_surface, u_atmos, v_atmos, stream=None):
    """flux_lib/heat/flux_heat_sensible.F90:137-167"""
    _call(lib.fc_flux_heat_sensible_rco, [flux_heat_sensible],
          [temperature_atmos, temperature_surface, u_atmos, v_atmos], [], stream)


def flux_momentum_cclm(flux_momentum_east, flux_momentum_north, diffusion_coefficient_momentum, pressure_surface,
                       specific_vapor_content_surface, temperature_surface, u_atmos, v_atmos,
                       gas_constant_air_new=None, gas_constant_vapor_new=None, stream=None):
    """flux_lib/momentum/flux_momentum.F90:22-76; either result may be None (the reference's `dummy`)"""
    _call(lib.fc_flux_momentum_cclm, [flux_momentum_east, flux_momentum_north],
          [diffusion_coefficient_momentum, pressure_surface, specific_vapor_content_surface, temperature_surface,
           u_atmos, v_atmos], [gas_constant_air_new, gas_constant_vapor_new], stream)


def flux_momentum_mom5(flux_momentum_east, flux_momentum_north, diffusion_coefficient_momentum, pressure_surface,
                       specific_vapor_content_surface, temperature_surface, u_atmos, v_atmos, stream=None):
    """flux_lib/momentum/flux_momentum.F90:78-108"""
    _call(lib.fc_flux_momentum_mom5, [flux_momentum_east, flux_momentum_north],
          [diffusion_coefficient_momentum, pressure_surface, specific_vapor_content_surface, temperature_surface,
           u_atmos, v_atmos], [], stream)


def flux_momentum_rco(flux_momentum_east, flux_momentum_north, u_atmos, v_atmos, stream=None):
    """flux_lib/momentum/flux_momentum.F90:110-138"""
    _call(lib.fc_flux_momentum_rco, [flux_momentum_east, flux_momentum_north], [u_atmos, v_atmos], [], stream)


def flux_radiation_blackbody_StBo(flux_radiation_blackbody, temperature_surface, stefan_boltzmann_constant_new=None,
                                  stream=None):
    """flux_lib/radiation/flux_radiation_blackbody.F90:22-42"""
    _call(lib.fc_flux_radiation_blackbody_StBo, [flux_radiation_blackbody], [temperature_surface],
          [stefan_boltzmann_constant_new], stream)


def distribute_radiation_flux(flux_radiation_surface_type, flux_radiation_averaged, albedo_averaged=None,
                              albedo_surface_type=None, stream=None):
    """flux_lib/radiation/distribute_radiation_flux.F90:12-26 (albedos accepted and unused, as in the reference)"""
    _call(lib.fc_distribute_radiation_flux, [flux_radiation_surface_type],
          [flux_radiation_averaged, albedo_averaged, albedo_surface_type], [], stream)
