!> Drop-in replacement of MODULE flux_calculator_calculate (src/flux_calculator_calculate.F90): the same nine public
!! procedure names and argument lists, forwarding to libfluxcalc_b200.so.  Source only (no Fortran compiler in this
!! image).  The host program keeps owning every array (flux_calculator_basic.F90:298, prepare.F90:40); this module
!! registers them once (gpu_register_fields, called after the prepare_* / add_output_field block,
!! flux_calculator.F90:594-761) and the calculators then act on the registered context.
MODULE flux_calculator_calculate

    use flux_calculator_basic
    use bias_corrections, only: lcorrections, corrections, E_MASS_EVAP_CORRECTION, init_date
    use fluxcalc_c_api
    use, intrinsic :: iso_c_binding

    IMPLICIT NONE

    PUBLIC calc_spec_vapor_surface, calc_flux_mass_evap, calc_flux_heat_latent, calc_flux_heat_sensible
    PUBLIC calc_flux_momentum_east, calc_flux_momentum_north, calc_flux_radiation_blackbody
    PUBLIC distribute_shortwave_radiation_flux, average_across_surface_types
    PUBLIC gpu_register_fields, gpu_step_early, gpu_step_normal

    TYPE(c_ptr), SAVE :: ctx = c_null_ptr

    CONTAINS

    !> one-time registration: local_field pointers, method strings, corrections, send list
    SUBROUTINE gpu_register_fields(my_bottom_model, num_surface_types, grid_size, local_field,               &
                                   which_spec_vapor_surface_t, which_spec_vapor_surface_u,                  &
                                   which_spec_vapor_surface_v, which_flux_mass_evap, which_flux_heat_latent, &
                                   which_flux_heat_sensible, which_flux_momentum,                            &
                                   which_flux_radiation_blackbody, num_output_fields, output_field, device)
        INTEGER,                                  INTENT(IN) :: my_bottom_model, num_surface_types, device
        INTEGER,                 DIMENSION(:),    INTENT(IN) :: grid_size
        TYPE(local_fields_type), DIMENSION(0:,:), INTENT(IN), TARGET :: local_field
        CHARACTER(len=20),       DIMENSION(:,:),  INTENT(IN) :: which_spec_vapor_surface_t, which_spec_vapor_surface_u
        CHARACTER(len=20),       DIMENSION(:,:),  INTENT(IN) :: which_spec_vapor_surface_v, which_flux_mass_evap
        CHARACTER(len=20),       DIMENSION(:,:),  INTENT(IN) :: which_flux_heat_latent, which_flux_heat_sensible
        CHARACTER(len=20),       DIMENSION(:,:),  INTENT(IN) :: which_flux_momentum, which_flux_radiation_blackbody
        INTEGER,                                  INTENT(IN) :: num_output_fields
        TYPE(io_fields_type),    DIMENSION(:),    INTENT(IN) :: output_field
        INTEGER(c_int64_t) :: gs(3)
        INTEGER :: i, g, k
        gs = INT(grid_size(1:3), c_int64_t)
        CALL fc_check(ctx, fc_create(ctx, gs, INT(num_surface_types, c_int), INT(device, c_int)), 'fc_create')
        DO i = 0, num_surface_types
            DO g = 1, 3
                DO k = 1, MAX_VARNAMES
                    IF (ASSOCIATED(local_field(i,g)%var(k)%field)) THEN
                        CALL fc_check(ctx, fc_bind_field(ctx, INT(i, c_int), INT(g, c_int), INT(k, c_int),         &
                                      c_loc(local_field(i,g)%var(k)%field), INT(grid_size(g), c_int64_t)), 'fc_bind_field')
                        ! the flag average_across_surface_types tests (flux_calculator_calculate.F90:376), as the registry has it
                        CALL fc_check(ctx, fc_set_allocated(ctx, INT(i, c_int), INT(g, c_int), INT(k, c_int),      &
                                      MERGE(1_c_int, 0_c_int, local_field(i,g)%var(k)%allocated)), 'fc_set_allocated')
                    ENDIF
                ENDDO
            ENDDO
        ENDDO
        DO i = 1, num_surface_types
            CALL set_m('which_spec_vapor_surface_t', i, which_spec_vapor_surface_t(my_bottom_model, i))
            CALL set_m('which_spec_vapor_surface_u', i, which_spec_vapor_surface_u(my_bottom_model, i))
            CALL set_m('which_spec_vapor_surface_v', i, which_spec_vapor_surface_v(my_bottom_model, i))
            CALL set_m('which_flux_mass_evap', i, which_flux_mass_evap(my_bottom_model, i))
            CALL set_m('which_flux_heat_latent', i, which_flux_heat_latent(my_bottom_model, i))
            CALL set_m('which_flux_heat_sensible', i, which_flux_heat_sensible(my_bottom_model, i))
            CALL set_m('which_flux_momentum', i, which_flux_momentum(my_bottom_model, i))
            CALL set_m('which_flux_radiation_blackbody', i, which_flux_radiation_blackbody(my_bottom_model, i))
        ENDDO
        IF (lcorrections(E_MASS_EVAP_CORRECTION)) THEN
            CALL fc_check(ctx, fc_set_corrections(ctx, 1_c_int, c_loc(corrections), INT(grid_size(1), c_int64_t),  &
                                                  1_c_int, INT(init_date, c_int)), 'fc_set_corrections')
        ENDIF
        DO k = 1, num_output_fields
            CALL fc_check(ctx, fc_add_output_field(ctx, INT(output_field(k)%surface_type, c_int),                  &
                          INT(output_field(k)%which_grid, c_int), INT(output_field(k)%idx, c_int)), 'fc_add_output_field')
        ENDDO
        CALL fc_check(ctx, fc_prepare(ctx, 0_c_int), 'fc_prepare')
    CONTAINS
        SUBROUTINE set_m(which, st, method)
            CHARACTER(len=*), INTENT(IN) :: which
            INTEGER, INTENT(IN) :: st
            CHARACTER(len=20), INTENT(IN) :: method
            CALL fc_check(ctx, fc_set_method(ctx, which//c_null_char, INT(st, c_int), trim(method)//c_null_char), which)
        END SUBROUTINE set_m
    END SUBROUTINE gpu_register_fields

    !!!!!!!!!! the nine calculators: identical interfaces to the reference, arguments other than which_grid are implied
    !!!!!!!!!! by the registered context

    SUBROUTINE calc_spec_vapor_surface(my_bottom_model, num_surface_types, which_grid, methods, grid_size, local_field)
        INTEGER, INTENT(IN) :: my_bottom_model, num_surface_types, which_grid
        CHARACTER(len=20), DIMENSION(:,:), INTENT(IN) :: methods
        INTEGER, DIMENSION(:), INTENT(IN) :: grid_size
        TYPE(local_fields_type), DIMENSION(0:,:), INTENT(INOUT) :: local_field
        CALL fc_check(ctx, fc_set_time(ctx, INT(current_step_time, c_int64_t)), 'fc_set_time')
        CALL fc_check(ctx, fc_calc_spec_vapor_surface(ctx, INT(which_grid, c_int)), 'calc_spec_vapor_surface')
    END SUBROUTINE calc_spec_vapor_surface

    SUBROUTINE calc_flux_mass_evap(my_bottom_model, num_surface_types, methods, grid_size, local_field)
        INTEGER, INTENT(IN) :: my_bottom_model, num_surface_types
        CHARACTER(len=20), DIMENSION(:,:), INTENT(IN) :: methods
        INTEGER, DIMENSION(:), INTENT(IN) :: grid_size
        TYPE(local_fields_type), DIMENSION(0:,:), INTENT(INOUT) :: local_field
        CALL fc_check(ctx, fc_set_time(ctx, INT(current_step_time, c_int64_t)), 'fc_set_time')   ! month for the bias
        CALL fc_check(ctx, fc_calc_flux_mass_evap(ctx), 'calc_flux_mass_evap')
    END SUBROUTINE calc_flux_mass_evap

    SUBROUTINE calc_flux_heat_latent(my_bottom_model, num_surface_types, methods, grid_size, local_field)
        INTEGER, INTENT(IN) :: my_bottom_model, num_surface_types
        CHARACTER(len=20), DIMENSION(:,:), INTENT(IN) :: methods
        INTEGER, DIMENSION(:), INTENT(IN) :: grid_size
        TYPE(local_fields_type), DIMENSION(0:,:), INTENT(INOUT) :: local_field
        CALL fc_check(ctx, fc_calc_flux_heat_latent(ctx), 'calc_flux_heat_latent')
    END SUBROUTINE calc_flux_heat_latent

    SUBROUTINE calc_flux_heat_sensible(my_bottom_model, num_surface_types, methods, grid_size, local_field)
        INTEGER, INTENT(IN) :: my_bottom_model, num_surface_types
        CHARACTER(len=20), DIMENSION(:,:), INTENT(IN) :: methods
        INTEGER, DIMENSION(:), INTENT(IN) :: grid_size
        TYPE(local_fields_type), DIMENSION(0:,:), INTENT(INOUT) :: local_field
        CALL fc_check(ctx, fc_calc_flux_heat_sensible(ctx), 'calc_flux_heat_sensible')
    END SUBROUTINE calc_flux_heat_sensible

    SUBROUTINE calc_flux_momentum_east(my_bottom_model, num_surface_types, which_grid, methods, grid_size, local_field)
        INTEGER, INTENT(IN) :: my_bottom_model, num_surface_types, which_grid
        CHARACTER(len=20), DIMENSION(:,:), INTENT(IN) :: methods
        INTEGER, DIMENSION(:), INTENT(IN) :: grid_size
        TYPE(local_fields_type), DIMENSION(0:,:), INTENT(INOUT) :: local_field
        CALL fc_check(ctx, fc_calc_flux_momentum_east(ctx, INT(which_grid, c_int)), 'calc_flux_momentum_east')
    END SUBROUTINE calc_flux_momentum_east

    SUBROUTINE calc_flux_momentum_north(my_bottom_model, num_surface_types, which_grid, methods, grid_size, local_field)
        INTEGER, INTENT(IN) :: my_bottom_model, num_surface_types, which_grid
        CHARACTER(len=20), DIMENSION(:,:), INTENT(IN) :: methods
        INTEGER, DIMENSION(:), INTENT(IN) :: grid_size
        TYPE(local_fields_type), DIMENSION(0:,:), INTENT(INOUT) :: local_field
        CALL fc_check(ctx, fc_calc_flux_momentum_north(ctx, INT(which_grid, c_int)), 'calc_flux_momentum_north')
    END SUBROUTINE calc_flux_momentum_north

    SUBROUTINE calc_flux_radiation_blackbody(my_bottom_model, num_surface_types, methods, grid_size, local_field)
        INTEGER, INTENT(IN) :: my_bottom_model, num_surface_types
        CHARACTER(len=20), DIMENSION(:,:), INTENT(IN) :: methods
        INTEGER, DIMENSION(:), INTENT(IN) :: grid_size
        TYPE(local_fields_type), DIMENSION(0:,:), INTENT(INOUT) :: local_field
        CALL fc_check(ctx, fc_calc_flux_radiation_blackbody(ctx), 'calc_flux_radiation_blackbody')
    END SUBROUTINE calc_flux_radiation_blackbody

    SUBROUTINE distribute_shortwave_radiation_flux(my_bottom_model, num_surface_types, grid_size, local_field)
        INTEGER, INTENT(IN) :: my_bottom_model, num_surface_types
        INTEGER, DIMENSION(:), INTENT(IN) :: grid_size
        TYPE(local_fields_type), DIMENSION(0:,:), INTENT(INOUT) :: local_field
        CALL fc_check(ctx, fc_distribute_shortwave_radiation_flux(ctx), 'distribute_shortwave_radiation_flux')
    END SUBROUTINE distribute_shortwave_radiation_flux

    SUBROUTINE average_across_surface_types(which_grid, my_idx, num_surface_types, grid_size, local_field)
        INTEGER, INTENT(IN) :: which_grid, my_idx, num_surface_types
        INTEGER, DIMENSION(:), INTENT(IN) :: grid_size
        TYPE(local_fields_type), DIMENSION(0:,:), INTENT(INOUT) :: local_field
        CALL fc_check(ctx, fc_average_across_surface_types(ctx, INT(which_grid, c_int), INT(my_idx, c_int)), &
                      'average_across_surface_types')
    END SUBROUTINE average_across_surface_types

    !> fused replacements of the two calculation blocks of the time loop; with these the host drops the individual
    !! calc_* calls (flux_calculator.F90:902 and :972-991) and the averaging calls of the send loops (:912-919, :1002-1009)
    SUBROUTINE gpu_step_early()
        CALL fc_check(ctx, fc_step_early(ctx, INT(current_step_time, c_int64_t)), 'fc_step_early')
    END SUBROUTINE gpu_step_early

    SUBROUTINE gpu_step_normal()
        CALL fc_check(ctx, fc_step_normal(ctx, INT(current_step_time, c_int64_t)), 'fc_step_normal')
    END SUBROUTINE gpu_step_normal

END MODULE flux_calculator_calculate
