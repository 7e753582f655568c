! decipher_c_api.f90 -- ISO_C_BINDING interface to libdecipher_b200.so (include/decipher_b200.h).
!
! NOT COMPILED IN THIS REPOSITORY'S CI: the build image has no Fortran compiler (gfortran, flang,
! nvfortran and ifort are all absent), so this module is shipped as source, written 1:1 against the C
! header.  Every derived-type component below has the same order, kind and meaning as the field of
! the same name in `dcp_catchment_desc` / `dcp_run_opts` / `dcp_run_stats` / `dcp_multi_stats`; tests/test_shim_consistency_cpu.py
! checks that mechanically against include/decipher_b200.h, together with the argument lists of the interfaces.
!
! Use: add this file and shim/dyna_main_gpu.f90 to RRMODEL's makefile sources, drop dyna_main.f90,
! link with -ldecipher_b200 -lcudart.  All the reference's setup routines (project, Tread_dyna,
! Inputs, mc_setup, modelstruct_setup, read_routingdata, genpar, write_output's formats) stay Fortran.
module decipher_c_api
    use, intrinsic :: iso_c_binding
    implicit none

    integer(c_int), parameter :: DCP_OK = 0
    integer(c_int32_t), parameter :: DCP_OPT_CHECK_BALANCE = 1, DCP_OPT_DEVICE_BUFFERS = 2, DCP_OPT_FAST_MATH = 4
    integer(c_int32_t), parameter :: DCP_ST_SOLFLAG_MASK = 3, DCP_ST_RZ_BALANCE = 4, DCP_ST_TS_OUTSUM = 8, &
                                     DCP_ST_TS_WBAL = 16, DCP_ST_NONFINITE = 32, DCP_ST_INIT_ITERCAP = 64, &
                                     DCP_ST_HIST_DROPPED = 128
    integer(c_int), parameter :: DCP_NMEASURE = 5   ! NSE, log-NSE, KGE, PBIAS, RMSE
    integer(c_int32_t), parameter :: DCP_PERF_NSE = 0, DCP_PERF_LOGNSE = 1, DCP_PERF_KGE = 2, DCP_PERF_PBIAS = 3, DCP_PERF_RMSE = 4
    integer(c_int32_t), parameter :: DCP_OPT_SIGNATURES = 8
    integer(c_int), parameter :: DCP_NSIG = 4, DCP_ABI_VERSION = 4

    type, bind(C) :: dcp_catchment_desc
        integer(c_int32_t) :: struct_size, nac, nriv, npt, nstep, ngp, nge, ntt
        real(c_double)     :: dt
        type(c_ptr) :: ac, st, mslp                          ! real(c_double) (nac)
        type(c_ptr) :: ippt, ipet, ipar, ipriv, ims_rz, ims_uz   ! integer(c_int32_t) (nac), 1-based values
        type(c_ptr) :: ntrans, itrans, wtrans                ! (nac), (sum ntrans), (sum ntrans+1)
        type(c_ptr) :: rivers, sum_ac_riv, qobs_start        ! (nac,nriv), (nriv), (nriv)
        integer(c_int32_t) :: nnode, ncell
        type(c_ptr) :: node_gauge_id, node_type, uptree_count, uptree_index
        type(c_ptr) :: frac_sub_area, river_data, node_to_flow_mapping
        type(c_ptr) :: rain, pet                             ! r_gau_step(ngp,nstep), pe_step(nge,nstep)
        type(c_ptr) :: qobs                                  ! qobs_riv_step(nriv,nstep_obs) or c_null_ptr
        integer(c_int32_t) :: nstep_obs, t_warm
        real(c_double)     :: flow_err
    end type

    type, bind(C) :: dcp_run_opts
        integer(c_int32_t) :: struct_size, flags, kernel, members_per_batch, steps_per_tile, reserved
        type(c_ptr) :: sig_out          ! DCP_OPT_SIGNATURES: (DCP_NSIG, nriv, n_member) flow-duration signatures, else c_null_ptr
    end type

    integer(c_int), parameter :: DCP_MAX_RANKS = 16

    ! timing of the last dcp_run_ensemble call on a handle (CUDA events, milliseconds)
    type, bind(C) :: dcp_run_stats
        real(c_double) :: ms_total, ms_init, ms_physics, ms_route, ms_h2d, ms_d2h
        integer(c_int64_t) :: n_physics_launches, n_launches
        integer(c_int32_t) :: kernel_used, hist_len
    end type

    ! what dcp_run_ensemble_multi reports about a run over several GPUs
    type, bind(C) :: dcp_multi_stats
        integer(c_int32_t) :: n_ranks, used_nccl, kernel_used, reserved
        real(c_double) :: ms_wall, ms_gather, ms_busy_max, ms_busy_sum
        real(c_double) :: ms_rank_device(DCP_MAX_RANKS)
        real(c_double) :: ms_rank_wall(DCP_MAX_RANKS)
    end type

    interface
        integer(c_int) function dcp_device_count() bind(C, name="dcp_device_count")
            import :: c_int
        end function
        integer(c_int) function dcp_set_device(device) bind(C, name="dcp_set_device")
            import :: c_int
            integer(c_int), value :: device
        end function
        ! compare with DCP_ABI_VERSION before the first call
        integer(c_int) function dcp_abi_version() bind(C, name="dcp_abi_version")
            import :: c_int
        end function
        type(c_ptr) function dcp_last_error() bind(C, name="dcp_last_error")
            import :: c_ptr
        end function
        integer(c_int) function dcp_catchment_create(desc, handle) bind(C, name="dcp_catchment_create")
            import :: c_int, c_ptr, dcp_catchment_desc
            type(dcp_catchment_desc), intent(in) :: desc
            type(c_ptr), intent(out) :: handle
        end function
        integer(c_int) function dcp_catchment_destroy(handle) bind(C, name="dcp_catchment_destroy")
            import :: c_int, c_ptr
            type(c_ptr), value :: handle
        end function
        ! mcpar(npt,7,n_member) in/out; q_out(nriv,nstep,n_member); perf_out(DCP_NMEASURE,nriv,n_member);
        ! reach_totals(3,nriv,n_member); init_diag(3,n_member); status(n_member).  Pass c_null_ptr to skip an output.
        integer(c_int) function dcp_run_ensemble(handle, first_member, n_member, mcpar, opts, q_out, perf_out, &
                                                 reach_totals, init_diag, status) bind(C, name="dcp_run_ensemble")
            import :: c_int, c_int64_t, c_ptr, dcp_run_opts
            type(c_ptr), value :: handle
            integer(c_int64_t), value :: first_member, n_member
            type(c_ptr), value :: mcpar
            type(dcp_run_opts), intent(in) :: opts
            type(c_ptr), value :: q_out, perf_out, reach_totals, init_diag, status
        end function
        ! standalone router (DTA/route_run_standalone.f90): velocity(n_series) in metres per timestep
        ! (tdh%v_param(1)); inflow(nriv,nstep,n_series) -> routed(nriv,nstep,n_series); status(n_series) or c_null_ptr
        integer(c_int) function dcp_route_timeseries(handle, n_series, velocity, inflow, routed, status, flags) &
                                                     bind(C, name="dcp_route_timeseries")
            import :: c_int, c_int32_t, c_int64_t, c_ptr
            type(c_ptr), value :: handle
            integer(c_int64_t), value :: n_series
            type(c_ptr), value :: velocity, inflow, routed, status
            integer(c_int32_t), value :: flags
        end function
        integer(c_int) function dcp_get_run_stats(handle, stats) bind(C, name="dcp_get_run_stats")
            import :: c_int, c_ptr, dcp_run_stats
            type(c_ptr), value :: handle
            type(dcp_run_stats), intent(out) :: stats
        end function
        ! GLUE-style behavioural filter: perf(DCP_NMEASURE,nriv,n_member) -> index_out(n_member) (0-based members, ascending),
        ! weight_out(n_member) or c_null_ptr, n_selected; gauge is 0-based, -1 = every gauge
        integer(c_int) function dcp_select_behavioural(perf, n_member, nriv, measure, gauge, threshold, higher_is_better, &
                                                       index_out, weight_out, n_selected, flags) &
                                                       bind(C, name="dcp_select_behavioural")
            import :: c_int, c_int32_t, c_int64_t, c_double, c_ptr
            type(c_ptr), value :: perf
            integer(c_int64_t), value :: n_member
            integer(c_int32_t), value :: nriv, measure, gauge
            real(c_double), value :: threshold
            integer(c_int32_t), value :: higher_is_better
            type(c_ptr), value :: index_out, weight_out
            integer(c_int64_t), intent(out) :: n_selected
            integer(c_int32_t), value :: flags
        end function
        ! ---- every GPU of the box: the MC loop of dyna_main.f90:166-206 over all ranks of a communicator ----
        ! devices(n_ranks): CUDA device of each rank, or c_null_ptr for rank r -> device r
        integer(c_int) function dcp_comm_init(n_ranks, devices, comm) bind(C, name="dcp_comm_init")
            import :: c_int, c_ptr
            integer(c_int), value :: n_ranks
            type(c_ptr), value :: devices
            type(c_ptr), intent(out) :: comm
        end function
        integer(c_int) function dcp_comm_destroy(comm) bind(C, name="dcp_comm_destroy")
            import :: c_int, c_ptr
            type(c_ptr), value :: comm
        end function
        integer(c_int) function dcp_comm_size(comm) bind(C, name="dcp_comm_size")
            import :: c_int, c_ptr
            type(c_ptr), value :: comm
        end function
        integer(c_int) function dcp_comm_uses_nccl(comm) bind(C, name="dcp_comm_uses_nccl")
            import :: c_int, c_ptr
            type(c_ptr), value :: comm
        end function
        ! arguments as dcp_run_ensemble, over ALL members; stats: c_null_ptr or c_loc of a type(dcp_multi_stats)
        integer(c_int) function dcp_run_ensemble_multi(comm, desc, n_member, mcpar, opts, q_out, perf_out, reach_totals, &
                                                       init_diag, status, stats) bind(C, name="dcp_run_ensemble_multi")
            import :: c_int, c_int64_t, c_ptr, dcp_catchment_desc, dcp_run_opts
            type(c_ptr), value :: comm
            type(dcp_catchment_desc), intent(in) :: desc
            integer(c_int64_t), value :: n_member
            type(c_ptr), value :: mcpar
            type(dcp_run_opts), intent(in) :: opts
            type(c_ptr), value :: q_out, perf_out, reach_totals, init_diag, status, stats
        end function
    end interface
end module decipher_c_api
