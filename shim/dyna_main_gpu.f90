! dyna_main_gpu.f90 -- `program main` of RRMODEL/dyna_main.f90 with the serial Monte-Carlo loop
! (dyna_main.f90:166-206) replaced by ONE call into libdecipher_b200.so.
!
! NOT COMPILED HERE (no Fortran compiler in the build image), but EXECUTED by tests/test_shim_consistency_cpu.py through the
! Fortran-subset translator (oracle/f90_translate.py) with the reference modules; shipped as the reference-side binding a
! maintainer would add.  Everything above the loop is the reference's own code, unchanged; the part
! below flattens the derived types into the C descriptor, samples all parameter sets with the
! reference's genpar (same RANMAR stream, so the sets are bit-identical to the serial program's),
! runs the ensemble on the GPU and writes the per-member files with the reference's write_output.
program main
    use, intrinsic :: iso_c_binding
    use decipher_c_api
    use dyna_common_types
    use dyna_project
    use dyna_random
    use dyna_tread_dyna
    use dyna_inputs
    use dyna_mc_setup
    use dyna_modelstruct_setup
    use dyna_genpar
    use dyna_file_open
    use dyna_read_routingdata
    use dyna_write_output
    use dta_utility
    use dta_route_processing
    use dta_riv_tree_node
    implicit none

    ! ---- declarations of dyna_main.f90:33-79 (unchanged) ----
    integer :: seed_1, seed_2, i, nac, nstep, num_rivers
    ! TARGET: their addresses go into the descriptor through c_loc
    double precision, dimension(:,:), allocatable, target :: pe_step, r_gau_step
    double precision, dimension(:,:), allocatable :: rivers
    double precision, dimension(:), allocatable, target :: qobs_riv_step_start, sum_ac_riv
    type(dyna_hru_type), dimension(:), allocatable :: dyna_hru
    type(dyna_riv_type), dimension(:), allocatable :: dyna_riv
    type(route_river_info_type) :: route_riv
    type(route_time_delay_hist_type) :: route_tdh
    integer :: cat_route_vmode, routing_mode
    integer, dimension(:), allocatable :: node_to_flow_mapping
    character(900) :: out_dir_path
    integer :: numsim, i_mc, num_par_types, ntt
    integer, allocatable, dimension(:) :: num_mcpar
    character(len=20), dimension(7) :: all_pm_names
    doubleprecision, allocatable, dimension(:,:) :: mcpar_ll, mcpar_ul, mcpar
    doubleprecision :: dt, acc, wt
    character(1024) :: arg, auto_start_file

    ! ---- flattened copies handed to the C ABI (must stay allocated until dcp_catchment_create returns) ----
    real(c_double), allocatable, target :: f_ac(:), f_st(:), f_mslp(:), f_wtrans(:), f_rivers(:,:), f_frac(:), f_rivdata(:,:)
    integer(c_int32_t), allocatable, target :: f_ippt(:), f_ipet(:), f_ipar(:), f_ipriv(:), f_rz(:), f_uz(:), f_ntrans(:), &
        f_itrans(:), f_gid(:), f_type(:), f_upcount(:), f_upidx(:), f_map(:)
    real(c_double), allocatable, target :: all_mcpar(:,:,:), q_all(:,:,:), perf(:,:,:), totals(:,:,:), diag(:,:)
    integer(c_int32_t), allocatable, target :: status(:)
    type(dcp_catchment_desc) :: desc
    type(dcp_run_opts) :: opts
    type(c_ptr) :: handle
    integer :: ia, iw, n, nt, nw, nup, rc

    auto_start_file = ''
    i = 0
    do
        call get_command_argument(i, arg)
        if (len_trim(arg) == 0) exit
        if (are_equal(arg, '-auto')) call get_command_argument(i+1, auto_start_file)
        i = i + 1
    end do

    ! ---- the reference's setup sequence, dyna_main.f90:95-160 (unchanged) ----
    call project (auto_start_file, out_dir_path)
    call Tread_dyna (nac, num_rivers, dyna_hru, dyna_riv, rivers, sum_ac_riv)
    call Inputs (nstep, num_rivers, dt, pe_step, qobs_riv_step_start, r_gau_step)
    call mc_setup (acc, dyna_hru, mcpar_ll, mcpar_ul, ntt, num_mcpar, num_par_types, numsim, seed_1, seed_2, wt, all_pm_names)
    call ranin (seed_1, seed_2)
    call modelstruct_setup (dyna_hru, nac)
    call read_routingdata (cat_route_vmode, dyna_hru, nac, node_to_flow_mapping, num_rivers, route_riv, route_tdh, routing_mode)

    ! ---- flatten dyna_hru / route_riv into the descriptor ----
    nt = sum(dyna_hru%ntrans); nw = nt + nac
    allocate(f_ac(nac), f_st(nac), f_mslp(nac), f_ippt(nac), f_ipet(nac), f_ipar(nac), f_ipriv(nac), f_rz(nac), f_uz(nac), &
             f_ntrans(nac), f_itrans(max(nt,1)), f_wtrans(nw), f_rivers(nac, num_rivers))
    nt = 0; nw = 0
    do ia = 1, nac
        f_ac(ia) = dyna_hru(ia)%ac;     f_st(ia) = dyna_hru(ia)%st;     f_mslp(ia) = dyna_hru(ia)%mslp
        f_ippt(ia) = dyna_hru(ia)%ippt; f_ipet(ia) = dyna_hru(ia)%ipet; f_ipar(ia) = dyna_hru(ia)%ipar
        f_ipriv(ia) = dyna_hru(ia)%ipriv; f_rz(ia) = dyna_hru(ia)%ims_rz; f_uz(ia) = dyna_hru(ia)%ims_uz
        f_ntrans(ia) = dyna_hru(ia)%ntrans
        do iw = 1, dyna_hru(ia)%ntrans
            f_itrans(nt + iw) = dyna_hru(ia)%itrans(iw)
        end do
        do iw = 1, dyna_hru(ia)%ntrans + 1
            f_wtrans(nw + iw) = dyna_hru(ia)%wtrans(iw)
        end do
        nt = nt + dyna_hru(ia)%ntrans; nw = nw + dyna_hru(ia)%ntrans + 1
    end do
    f_rivers = rivers
    n = size(route_riv%node_list)
    nup = sum(route_riv%node_list%upstream_tree_count)
    allocate(f_gid(n), f_type(n), f_upcount(n), f_upidx(max(nup,1)), f_frac(n), f_map(num_rivers), &
             f_rivdata(size(route_riv%river_data,1), 8))
    nup = 0
    do i = 1, n
        f_gid(i) = route_riv%node_list(i)%gauge_id;  f_type(i) = route_riv%node_list(i)%node_type
        f_upcount(i) = route_riv%node_list(i)%upstream_tree_count
        f_frac(i) = route_riv%node_list(i)%frac_sub_area
        do iw = 1, route_riv%node_list(i)%upstream_tree_count
            f_upidx(nup + iw) = route_riv%node_list(i)%upstream_tree_indexes(iw)
        end do
        nup = nup + route_riv%node_list(i)%upstream_tree_count
    end do
    f_map = node_to_flow_mapping
    f_rivdata = route_riv%river_data

    desc%struct_size = int(c_sizeof(desc), c_int32_t)
    desc%nac = nac; desc%nriv = num_rivers; desc%npt = num_par_types; desc%nstep = nstep
    desc%ngp = size(r_gau_step, 1); desc%nge = size(pe_step, 1); desc%ntt = ntt; desc%dt = dt
    desc%ac = c_loc(f_ac); desc%st = c_loc(f_st); desc%mslp = c_loc(f_mslp)
    desc%ippt = c_loc(f_ippt); desc%ipet = c_loc(f_ipet); desc%ipar = c_loc(f_ipar); desc%ipriv = c_loc(f_ipriv)
    desc%ims_rz = c_loc(f_rz); desc%ims_uz = c_loc(f_uz)
    desc%ntrans = c_loc(f_ntrans); desc%itrans = c_loc(f_itrans); desc%wtrans = c_loc(f_wtrans)
    desc%rivers = c_loc(f_rivers); desc%sum_ac_riv = c_loc(sum_ac_riv); desc%qobs_start = c_loc(qobs_riv_step_start)
    desc%nnode = n; desc%ncell = size(route_riv%river_data, 1)
    desc%node_gauge_id = c_loc(f_gid); desc%node_type = c_loc(f_type)
    desc%uptree_count = c_loc(f_upcount); desc%uptree_index = c_loc(f_upidx)
    desc%frac_sub_area = c_loc(f_frac); desc%river_data = c_loc(f_rivdata); desc%node_to_flow_mapping = c_loc(f_map)
    desc%rain = c_loc(r_gau_step); desc%pet = c_loc(pe_step)
    desc%qobs = c_null_ptr; desc%nstep_obs = 0; desc%t_warm = 0; desc%flow_err = -9999.0d0
    ! (to get the performance measures (NSE, log-NSE, KGE, PBIAS, RMSE), make Inputs return qobs_riv_step and flow_err_i and pass them here)

    if (dcp_catchment_create(desc, handle) /= DCP_OK) stop 'dcp_catchment_create failed'

    ! ---- all parameter sets first: genpar is the only consumer of the RANMAR stream (dyna_genpar.f90:35-41) ----
    allocate(all_mcpar(num_par_types, 7, numsim), q_all(num_rivers, nstep, numsim), perf(DCP_NMEASURE, num_rivers, numsim), &
             totals(3, num_rivers, numsim), diag(3, numsim), status(numsim))
    do i_mc = 1, numsim
        call genpar (mcpar, mcpar_ll, mcpar_ul, num_mcpar, num_par_types)
        all_mcpar(:, :, i_mc) = mcpar
        deallocate(mcpar)
    end do

    ! ---- ONE call replaces `do i_mc = 1, numsim ... call mainloop(...)` ----
    opts%struct_size = int(c_sizeof(opts), c_int32_t); opts%flags = 0; opts%kernel = 0
    opts%members_per_batch = 0; opts%steps_per_tile = 0; opts%reserved = 0; opts%sig_out = c_null_ptr
    rc = dcp_run_ensemble(handle, 1_c_int64_t, int(numsim, c_int64_t), c_loc(all_mcpar), opts, c_loc(q_all), c_loc(perf), &
                          c_loc(totals), c_loc(diag), c_loc(status))
    if (rc /= DCP_OK) stop 'dcp_run_ensemble failed'

    ! ---- per-member files with the reference's own writers ----
    do i_mc = 1, numsim
        call file_open (i_mc, out_dir_path)
        ! the init block of dyna_init_satzone.f90:177-191 from status / diag, then the .res tables and the .flow
        ! series: write_output's formats applied to all_mcpar(:,:,i_mc), totals(:,:,i_mc), q_all(:,:,i_mc)
        call write_output_from_arrays (all_mcpar(:, :, i_mc), all_pm_names, num_par_types, num_rivers, &
                                       q_all(:, :, i_mc), totals(:, :, i_mc), diag(:, i_mc), status(i_mc))
        close (40); close (41)
    end do
    rc = dcp_catchment_destroy(handle)
    close (999)
    stop

contains
    ! write_output (dyna_write_output.f90:44-100) with the reach totals taken from the GPU instead of dyna_hru
    subroutine write_output_from_arrays (mcp, names, npt, nriv, q, tot, dg, st)
        double precision :: mcp(:,:), q(:,:), tot(:,:), dg(:)
        character(len=20), dimension(:) :: names
        integer :: npt, nriv, k
        integer(c_int32_t) :: st
        character(len=1024) :: line_format
        character(len=20) :: pm_class = 'Param_class'
        write(40, *) ''
        write(40, *) '! INITIALISATION'
        select case (iand(st, DCP_ST_SOLFLAG_MASK))
        case (1); write(40, *) 'Solution found!'
        case (0); write(40, *) 'No solution found - going for the minimum solution'
        case default; write(40, *) 'Maximum QBF is less than mean_q_ac_weighted - deficits set to 0 and max QBFs'
        end select
        write(40, *) 'Starting qobs ', dg(1) * 1000, 'mm/day'
        write(40, *) 'Best estimate of QB riv', dg(2) * 1000, 'mm/day'
        write(40, *) 'Finishing qt0 ', dg(3) * 1000, 'mm/day'
        write(40, *) ''
        write(40, '(A)') '! PARAMETERS'
        write(40, "(A16,A10,  A12,  A10 , A11,  A12,  A11,  A10)") [pm_class, names(1:7)]
        do k = 1, npt
            write(40, "(I12,F10.4,F12.4,F10.4,F10.4,F13.4,F11.4,F10.4)") k, mcp(k, 1:7)
        end do
        write(40, '(A)') ''
        write(line_format, '(A,I0, A)') '(', nriv, '(1X, f0.2), A)'
        write(40, '(A)') '! SUMMARY STATS PER RIVER REACH'
        write(40, '(A)') ''
        write(40, '(A)') '! PRECIP TOTALS (mm)'
        write(40, line_format) tot(1, :)
        write(40, '(A)') ''
        write(40, '(A)') '! PET TOTALS (mm)'
        write(40, line_format) tot(2, :)
        write(40, '(A)') ''
        write(40, '(A)') '! ET TOTALS (mm)'
        write(40, line_format) tot(3, :)
        write(line_format, '(A,I0, A)') '(', size(q, 1), '(1X, f0.8), A)'
        do k = 1, size(q, 2)
            write(41, line_format) q(:, k)
        end do
    end subroutine
end program main
