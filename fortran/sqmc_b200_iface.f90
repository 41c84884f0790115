!==============================================================================
! sqmc_b200_iface.f90 -- ISO_C_BINDING interfaces to libsqmc_b200.so
! (include/sqmc_b200.h).  Drop this file into the reference's src/ directory, add it
! to the Makefile object list before chemistry.f90, and link with -lsqmc_b200.
! The call-site patches are listed in INTEGRATION.md.
!
! NOTE: the authoring container has no Fortran compiler; this shim is written
! against the Fortran 2003 standard and has not been compiled here.
!==============================================================================
module sqmc_b200_iface
  use, intrinsic :: iso_c_binding
  implicit none
  private
  public :: b200_init, b200_system_chem, b200_system_heg, b200_system_hubbardk, b200_build_h, &
            b200_export_upper, b200_matvec, b200_projector, b200_scale_values, b200_davidson, b200_lanczos, b200_free, b200_check
  public :: sqmc_b200_get_unique_id

  interface
    integer(c_int) function sqmc_b200_get_unique_id(id) bind(C, name="sqmc_b200_get_unique_id")
      import :: c_int, c_char
      character(kind=c_char) :: id(128)
    end function
    integer(c_int) function sqmc_b200_init(device, rank, nranks, id) bind(C, name="sqmc_b200_init")
      import :: c_int, c_char
      integer(c_int), value :: device, rank, nranks
      character(kind=c_char) :: id(128)
    end function
    function sqmc_b200_last_error() result(p) bind(C, name="sqmc_b200_last_error")
      import :: c_ptr
      type(c_ptr) :: p
    end function
    integer(c_int) function sqmc_b200_system_chem(h, norb, nup, ndn, integrals, nint, combine_2, time_sym, z) bind(C, name="sqmc_b200_system_chem")
      import :: c_int, c_int64_t, c_double, c_ptr, c_int32_t
      type(c_ptr) :: h
      integer(c_int), value :: norb, nup, ndn, time_sym, z
      real(c_double) :: integrals(*)
      integer(c_int64_t), value :: nint
      integer(c_int32_t) :: combine_2(*)
    end function
    integer(c_int) function sqmc_b200_system_heg(h, norb, n_dim, k_vectors, length_cell, nup, ndn) bind(C, name="sqmc_b200_system_heg")
      import :: c_int, c_double, c_ptr
      type(c_ptr) :: h
      integer(c_int), value :: norb, n_dim, nup, ndn
      real(c_double) :: k_vectors(*)
      real(c_double), value :: length_cell
    end function
    integer(c_int) function sqmc_b200_system_hubbardk(h, l_x, l_y, k_vectors, k_energies, ubyn, nup, ndn) bind(C, name="sqmc_b200_system_hubbardk")
      import :: c_int, c_double, c_ptr, c_int32_t
      type(c_ptr) :: h
      integer(c_int), value :: l_x, l_y, nup, ndn
      integer(c_int32_t) :: k_vectors(*)
      real(c_double) :: k_energies(*)
      real(c_double), value :: ubyn
    end function
    integer(c_int) function sqmc_b200_free(h) bind(C, name="sqmc_b200_free")
      import :: c_int, c_ptr
      type(c_ptr), value :: h
    end function
    integer(c_int) function sqmc_b200_build_h(h, n, dets_up, dets_dn, ndet_old, nnz_upper) bind(C, name="sqmc_b200_build_h")
      import :: c_int, c_int64_t, c_ptr
      type(c_ptr), value :: h, dets_up, dets_dn      ! c_loc of integer(ik) arrays (16 B per det)
      integer(c_int64_t), value :: n, ndet_old
      integer(c_int64_t) :: nnz_upper
    end function
    integer(c_int) function sqmc_b200_export_upper(h, counts, indices, values) bind(C, name="sqmc_b200_export_upper")
      import :: c_int, c_int64_t, c_double, c_ptr
      type(c_ptr), value :: h
      integer(c_int64_t) :: counts(*), indices(*)
      real(c_double) :: values(*)
    end function
    integer(c_int) function sqmc_b200_matvec(h, x, y, nvec, ldx) bind(C, name="sqmc_b200_matvec")
      import :: c_int, c_int64_t, c_double, c_ptr
      type(c_ptr), value :: h
      real(c_double) :: x(*), y(*)
      integer(c_int), value :: nvec
      integer(c_int64_t), value :: ldx
    end function
    integer(c_int) function sqmc_b200_projector(h, tau, e_trial, w, deltaw) bind(C, name="sqmc_b200_projector")
      import :: c_int, c_double, c_ptr
      type(c_ptr), value :: h
      real(c_double), value :: tau, e_trial
      real(c_double) :: w(*), deltaw(*)
    end function
    integer(c_int) function sqmc_b200_scale_values(h, ratio) bind(C, name="sqmc_b200_scale_values")
      import :: c_int, c_double, c_ptr
      type(c_ptr), value :: h
      real(c_double), value :: ratio
    end function
    integer(c_int) function sqmc_b200_davidson(h, n_states, v0, evecs, evals, tol, max_vec, n_matvec, ritz_log, ritz_cap, n_ritz) bind(C, name="sqmc_b200_davidson")
      import :: c_int, c_double, c_ptr
      type(c_ptr), value :: h, v0, ritz_log          ! v0 / ritz_log may be c_null_ptr
      integer(c_int), value :: n_states, max_vec, ritz_cap
      real(c_double) :: evecs(*), evals(*)
      real(c_double), value :: tol
      integer(c_int) :: n_matvec, n_ritz
    end function
    integer(c_int) function sqmc_b200_lanczos(h, v0, evec, eig3, tol, max_iter, n_iter, ritz_log, ritz_cap, n_ritz) bind(C, name="sqmc_b200_lanczos")
      import :: c_int, c_double, c_ptr
      type(c_ptr), value :: h, v0, ritz_log          ! v0 / ritz_log may be c_null_ptr
      real(c_double) :: evec(*), eig3(3)
      real(c_double), value :: tol
      integer(c_int), value :: max_iter, ritz_cap
      integer(c_int) :: n_iter, n_ritz
    end function
    integer(c_int) function sqmc_b200_set_ownership(h, owner_of_row, n_owned) bind(C, name="sqmc_b200_set_ownership")
      import :: c_int, c_int32_t, c_int64_t, c_ptr
      type(c_ptr), value :: h
      integer(c_int32_t) :: owner_of_row(*)          ! rank (0-based) owning each determinant of the list (get_det_owner)
      integer(c_int64_t) :: n_owned
    end function
    integer(c_int) function sqmc_b200_matvec_local(h, x_local, y_local, nvec, ld_local) bind(C, name="sqmc_b200_matvec_local")
      import :: c_int, c_int64_t, c_double, c_ptr
      type(c_ptr), value :: h
      real(c_double) :: x_local(*), y_local(*)
      integer(c_int), value :: nvec
      integer(c_int64_t), value :: ld_local
    end function
    integer(c_int) function sqmc_b200_projector_local(h, tau, e_trial, w_local, deltaw_local) bind(C, name="sqmc_b200_projector_local")
      import :: c_int, c_double, c_ptr
      type(c_ptr), value :: h
      real(c_double), value :: tau, e_trial
      real(c_double) :: w_local(*), deltaw_local(*)
    end function
    integer(c_int) function sqmc_b200_davidson_local(h, n_states, v0, evecs, evals, tol, max_vec, n_matvec, ritz_log, ritz_cap, n_ritz) &
        bind(C, name="sqmc_b200_davidson_local")
      import :: c_int, c_double, c_ptr
      type(c_ptr), value :: h, v0, ritz_log          ! v0 / ritz_log may be c_null_ptr
      integer(c_int), value :: n_states, max_vec, ritz_cap
      real(c_double) :: evecs(*), evals(*)
      real(c_double), value :: tol
      integer(c_int) :: n_matvec, n_ritz
    end function
    integer(c_int) function sqmc_b200_pt2_alias(h, n, dets_up, dets_dn, wts, var_energy, eps_pt, eps_pt_big, n_mc, target_error, rannyu_state, &
        max_samples, pt_energy, pt_energy_std_dev, n_samples, e_2pt_samples, n_connected) bind(C, name="sqmc_b200_pt2_alias")
      import :: c_int, c_int32_t, c_int64_t, c_double, c_ptr
      type(c_ptr), value :: h, dets_up, dets_dn, e_2pt_samples   ! c_loc of the integer(ik) arrays; e_2pt_samples may be c_null_ptr
      integer(c_int64_t), value :: n
      real(c_double) :: wts(*)
      real(c_double), value :: var_energy, eps_pt, eps_pt_big, target_error
      integer(c_int), value :: n_mc, max_samples
      integer(c_int32_t) :: rannyu_state(4)          ! savern(rannyu_state) before, setrn(rannyu_state) after
      real(c_double) :: pt_energy, pt_energy_std_dev
      integer(c_int) :: n_samples
      integer(c_int64_t) :: n_connected
    end function
    integer(c_int) function sqmc_b200_pt2_sample(h, n, dets_up, dets_dn, n_sampled, sampled_up, sampled_dn, sampled_coeffs, w_over_p, n_mc, &
        var_energy, eps_pt, eps_pt_big, e_2pt_this_sample, n_connected) bind(C, name="sqmc_b200_pt2_sample")
      import :: c_int, c_int64_t, c_double, c_ptr
      type(c_ptr), value :: h, dets_up, dets_dn, sampled_up, sampled_dn
      integer(c_int64_t), value :: n, n_sampled
      real(c_double) :: sampled_coeffs(*), w_over_p(*)
      integer(c_int), value :: n_mc
      real(c_double), value :: var_energy, eps_pt, eps_pt_big
      real(c_double) :: e_2pt_this_sample
      integer(c_int64_t) :: n_connected
    end function
    integer(c_int) function sqmc_b200_register_host(ptr, bytes) bind(C, name="sqmc_b200_register_host")
      import :: c_int, c_int64_t, c_ptr
      type(c_ptr), value :: ptr
      integer(c_int64_t), value :: bytes
    end function
  end interface

contains

  subroutine b200_check(ierr)
    ! maps a non-zero status to the reference's error style (mpi_stop, mpi_routines.f90:5137)
    use mpi_routines, only : mpi_stop
    integer(c_int), intent(in) :: ierr
    character(kind=c_char), pointer :: msg(:)
    character(len=512) :: text
    integer :: i
    if (ierr == 0) return
    call c_f_pointer(sqmc_b200_last_error(), msg, [512])
    text = ' '
    do i = 1, 512
      if (msg(i) == c_null_char) exit
      text(i:i) = msg(i)
    enddo
    call mpi_stop('sqmc_b200: '//trim(text))
  end subroutine

  subroutine b200_init(device, rank, nranks, id)
    integer, intent(in) :: device, rank, nranks
    character(kind=c_char), intent(inout) :: id(128)   ! from sqmc_b200_get_unique_id on rank 0 + MPI_Bcast
    call b200_check(sqmc_b200_init(int(device, c_int), int(rank, c_int), int(nranks, c_int), id))
  end subroutine

  subroutine b200_system_chem(h, norb, nup, ndn, integrals, combine_2, time_sym, z)
    type(c_ptr), intent(out) :: h
    integer, intent(in) :: norb, nup, ndn, z
    real(c_double), intent(in) :: integrals(:)
    integer, intent(in) :: combine_2(:, :)            ! (norb+1, norb+1), default integer = 32 bit
    logical, intent(in) :: time_sym
    integer(c_int32_t), allocatable :: c2(:)
    allocate(c2(size(combine_2)))
    c2 = reshape(combine_2, [size(combine_2)])
    call b200_check(sqmc_b200_system_chem(h, int(norb, c_int), int(nup, c_int), int(ndn, c_int), integrals, &
                    int(size(integrals), c_int64_t), c2, merge(1_c_int, 0_c_int, time_sym), int(z, c_int)))
  end subroutine

  subroutine b200_system_heg(h, norb, n_dim, k_vectors, length_cell, nup, ndn)
    type(c_ptr), intent(out) :: h
    integer, intent(in) :: norb, n_dim, nup, ndn
    real(c_double), intent(in) :: k_vectors(:, :), length_cell
    call b200_check(sqmc_b200_system_heg(h, int(norb, c_int), int(n_dim, c_int), k_vectors, length_cell, int(nup, c_int), int(ndn, c_int)))
  end subroutine

  subroutine b200_system_hubbardk(h, l_x, l_y, k_vectors, k_energies, ubyn, nup, ndn)
    type(c_ptr), intent(out) :: h
    integer, intent(in) :: l_x, l_y, nup, ndn
    integer, intent(in) :: k_vectors(:, :)
    real(c_double), intent(in) :: k_energies(:), ubyn
    integer(c_int32_t), allocatable :: kv(:)
    allocate(kv(size(k_vectors)))
    kv = reshape(k_vectors, [size(k_vectors)])
    call b200_check(sqmc_b200_system_hubbardk(h, int(l_x, c_int), int(l_y, c_int), kv, k_energies, ubyn, int(nup, c_int), int(ndn, c_int)))
  end subroutine

  subroutine b200_build_h(h, dets_up, dets_dn, ndet_old, nnz_upper)
    ! dets_up/dn: integer(ik) arrays, passed by address (16 B per determinant)
    use types, only : ik
    type(c_ptr), intent(in) :: h
    integer(ik), intent(in), target :: dets_up(:), dets_dn(:)
    integer(c_int64_t), intent(in) :: ndet_old
    integer(c_int64_t), intent(out) :: nnz_upper
    call b200_check(sqmc_b200_build_h(h, int(size(dets_up), c_int64_t), c_loc(dets_up), c_loc(dets_dn), ndet_old, nnz_upper))
  end subroutine

  subroutine b200_export_upper(h, H_nonzero_elements, H_indices, H_values)
    type(c_ptr), intent(in) :: h
    integer(c_int64_t), intent(out) :: H_nonzero_elements(:), H_indices(:)
    real(c_double), intent(out) :: H_values(:)
    call b200_check(sqmc_b200_export_upper(h, H_nonzero_elements, H_indices, H_values))
  end subroutine

  subroutine b200_matvec(h, n, vector, answer)
    type(c_ptr), intent(in) :: h
    integer, intent(in) :: n
    real(c_double), intent(in) :: vector(:)
    real(c_double), intent(out) :: answer(:)
    call b200_check(sqmc_b200_matvec(h, vector, answer, 1_c_int, int(n, c_int64_t)))
  end subroutine

  subroutine b200_projector(h, tau, e_trial, w, deltaw)
    type(c_ptr), intent(in) :: h
    real(c_double), intent(in) :: tau, e_trial, w(:)
    real(c_double), intent(out) :: deltaw(:)
    call b200_check(sqmc_b200_projector(h, tau, e_trial, w, deltaw))
  end subroutine

  subroutine b200_scale_values(h, ratio)
    type(c_ptr), intent(in) :: h
    real(c_double), intent(in) :: ratio
    call b200_check(sqmc_b200_scale_values(h, ratio))
  end subroutine

  subroutine b200_davidson(h, n_states, final_vector, lowest_eigenvalues, initial_vector)
    type(c_ptr), intent(in) :: h
    integer, intent(in) :: n_states
    real(c_double), intent(out) :: final_vector(:, :), lowest_eigenvalues(:)
    real(c_double), intent(in), optional, target :: initial_vector(:, :)
    integer(c_int) :: nmv, nlog
    type(c_ptr) :: v0
    v0 = c_null_ptr
    if (present(initial_vector)) v0 = c_loc(initial_vector)
    call b200_check(sqmc_b200_davidson(h, int(n_states, c_int), v0, final_vector, lowest_eigenvalues, 1.e-10_c_double, &
                    50_c_int, nmv, c_null_ptr, 0_c_int, nlog))
  end subroutine

  ! matrix_lanczos_sparse(n,lowest_eigenvector,lowest_eigenvalue,...,highest_eigenvalue,second_lowest_eigenvalue,initial_vector)
  ! (more_tools.f90:1742) on the handle's resident matrix
  subroutine b200_lanczos(h, lowest_eigenvector, lowest_eigenvalue, highest_eigenvalue, second_lowest_eigenvalue, initial_vector)
    type(c_ptr), intent(in) :: h
    real(c_double), intent(out) :: lowest_eigenvector(:), lowest_eigenvalue
    real(c_double), intent(out), optional :: highest_eigenvalue, second_lowest_eigenvalue
    real(c_double), intent(in), optional, target :: initial_vector(:)
    real(c_double) :: eig3(3)
    integer(c_int) :: nit, nlog
    type(c_ptr) :: v0
    v0 = c_null_ptr
    if (present(initial_vector)) v0 = c_loc(initial_vector)
    call b200_check(sqmc_b200_lanczos(h, v0, lowest_eigenvector, eig3, 1.e-10_c_double, 50_c_int, nit, c_null_ptr, 0_c_int, nlog))
    lowest_eigenvalue = eig3(1)
    if (present(highest_eigenvalue)) highest_eigenvalue = eig3(2)
    if (present(second_lowest_eigenvalue)) second_lowest_eigenvalue = eig3(3)
  end subroutine

  ! ---- the MPI data distribution of the reference: slices in, slices out --------------------------------------------
  ! owner(i) = get_det_owner(dets_up(i), dets_dn(i)) (mpi_routines.f90:419) for the list last passed to b200_build_h
  subroutine b200_set_ownership(h, owner, my_n)
    type(c_ptr), intent(in) :: h
    integer, intent(in) :: owner(:)
    integer(c_int64_t), intent(out) :: my_n
    integer(c_int32_t), allocatable :: o(:)
    allocate(o(size(owner)))
    o = int(owner, c_int32_t)
    call b200_check(sqmc_b200_set_ownership(h, o, my_n))
  end subroutine

  ! fast_sparse_matrix_multiply_local_band + mpi_redscatt_real_dparray (do_walk.f90:2259-2260): deltaw(1:my_nimp) on return
  subroutine b200_matvec_local(h, vector_local, answer_local)
    type(c_ptr), intent(in) :: h
    real(c_double), intent(in) :: vector_local(:)
    real(c_double), intent(out) :: answer_local(:)
    call b200_check(sqmc_b200_matvec_local(h, vector_local, answer_local, 1_c_int, int(max(size(vector_local), 1), c_int64_t)))
  end subroutine

  ! the same + e_trial*tau*my_imp_wt (do_walk.f90:2290)
  subroutine b200_projector_local(h, tau, e_trial, my_imp_wt, deltaw_local)
    type(c_ptr), intent(in) :: h
    real(c_double), intent(in) :: tau, e_trial, my_imp_wt(:)
    real(c_double), intent(out) :: deltaw_local(:)
    call b200_check(sqmc_b200_projector_local(h, tau, e_trial, my_imp_wt, deltaw_local))
  end subroutine

  ! davidson_sparse_mpi2 (more_tools.f90:2525): final_vector(local_det_map%ndets, n_states)
  subroutine b200_davidson_local(h, n_states, final_vector, lowest_eigenvalues, initial_vector)
    type(c_ptr), intent(in) :: h
    integer, intent(in) :: n_states
    real(c_double), intent(out) :: final_vector(:, :), lowest_eigenvalues(:)
    real(c_double), intent(in), optional, target :: initial_vector(:, :)
    integer(c_int) :: nmv, nlog
    type(c_ptr) :: v0
    v0 = c_null_ptr
    if (present(initial_vector)) v0 = c_loc(initial_vector)
    call b200_check(sqmc_b200_davidson_local(h, int(n_states, c_int), v0, final_vector, lowest_eigenvalues, 1.e-10_c_double, &
                    50_c_int, nmv, c_null_ptr, 0_c_int, nlog))
  end subroutine

  ! page-lock walk_wt / deltaw once: the projector moves them across PCIe every Monte Carlo step
  subroutine b200_register_host(array)
    real(c_double), intent(in), target :: array(:)
    call b200_check(sqmc_b200_register_host(c_loc(array), int(size(array), c_int64_t) * 8_c_int64_t))
  end subroutine

  subroutine b200_free(h)
    type(c_ptr), intent(inout) :: h
    call b200_check(sqmc_b200_free(h))
    h = c_null_ptr
  end subroutine

end module sqmc_b200_iface
