"""Variable table, grids and method vocabulary of the reference (flux_calculator_basic.F90:42-63,
flux_calculator.F90:99-107).  Pure data, no arithmetic."""

VARNAMES = [
    "ALBE", "ALBA", "AMOI", "AMOM", "FARE", "FICE", "PATM", "PSUR",
    "QATM", "TATM", "TSUR", "UATM", "VATM", "U10M", "V10M",
    "CMOM", "CMOI", "CHEA",
    "QSUR",
    "HLAT", "HSEN",
    "MEVA", "MPRE", "MRAI", "MSNO",
    "RBBR", "RLWD", "RLWU", "RSID", "RSIU", "RSIN", "RSDD", "RSDR",
    "UMOM", "VMOM",
]
IDX = {name: i + 1 for i, name in enumerate(VARNAMES)}      # idx_* are 1-based
T_GRID, U_GRID, V_GRID = 1, 2, 3
GRID_NAMES = {1: "t_grid", 2: "u_grid", 3: "v_grid"}
MAX_SURFACE_TYPES = 10

METHODS = {
    "which_spec_vapor_surface_t": ("none", "copy", "CCLM"),
    "which_spec_vapor_surface_u": ("none", "copy", "CCLM"),
    "which_spec_vapor_surface_v": ("none", "copy", "CCLM"),
    "which_flux_mass_evap": ("none", "zero", "copy", "CCLM", "MOM5", "RCO"),
    "which_flux_heat_latent": ("none", "zero", "copy", "water", "ice"),
    "which_flux_heat_sensible": ("none", "zero", "copy", "CCLM", "MOM5", "RCO"),
    "which_flux_momentum": ("none", "zero", "copy", "CCLM", "MOM5", "RCO"),
    "which_flux_radiation_blackbody": ("none", "zero", "copy", "StBo"),
}
EARLY_OUTPUTS = ("RBBR", "TSUR", "FICE", "ALBE")            # basic.F90:271-273


def var_index(v):
    return IDX[v] if isinstance(v, str) else int(v)
