!  sosgpu_iso_c.f90 -- ISO_C_BINDING interface of libsosgpu.so (include/sosgpu.h) for the Fortran host side.
!
!  COMPILE-UNTESTED: no Fortran compiler exists in the build image (SURVEY.md F2).  The C ABI it binds is
!  exercised by the C++ shims (csrc/sosgpu_shims.cu) and the ctypes layer (api.py).
!
!  Two ways to drop the GPU path into SOS_PROC.F:
!   (1) link libsosgpu.so *instead of* SOS_OS.o / SOS_AGGREGATE.o: it exports the F77 symbols sos_os_ and
!       sos_aggregate_ with the reference's argument lists (SOS_OS.F:303-308, SOS_AGGREGATE.F:172-178);
!       no source change in SOS.F / SOS_PROC.F.
!   (2) replace the CKD loop SOS_PROC.F:3459-3594 by one call of sosgpu_solve_batch (all terms at once) using the
!       interfaces below.
module sosgpu_iso_c
  use, intrinsic :: iso_c_binding
  implicit none

  type, bind(c) :: sosgpu_optics            ! SOS_PREPA_OS outputs + SOS scalars (include/sosgpu.h)
    integer(c_int)    :: nbmu
    type(c_ptr)       :: rmu, ga
    integer(c_int)    :: n0
    real(c_double)    :: tetas
    integer(c_int)    :: os_nb
    type(c_ptr)       :: alpha, beta, gamma, zeta
    real(c_double)    :: a_trunc, piz, piztr, ron, rho
    integer(c_int)    :: imat_surf, ifresnel
    real(c_double)    :: ind_surf
    type(c_ptr)       :: surf
    integer(c_int)    :: n_surf_rec, igmax, ipolar
    real(c_double)    :: zout
  end type

  type, bind(c) :: sosgpu_term              ! PROFIL_TMP content of one CKD term + its weight AIK
    integer(c_int)    :: optics, group
    real(c_double)    :: aik
    integer(c_int)    :: nt
    type(c_ptr)       :: zprof, h, pcaer, pcmol
  end type

  type, bind(c) :: sosgpu_ckd               ! tables of READ_CKD_COEFF, Fortran storage (inc/SOS.h extents)
    integer(c_int)    :: nb_temp, nb_pres, nb_conc_h2o
    type(c_ptr)       :: tab_temp, tab_pres, tab_conc_h2o, nexp, kdis_ki, kdis_ki_h2o
  end type

  type, bind(c) :: sosgpu_gas_profile       ! ABS_USERPROFIL(50,13), ALTABS(50), RO(8,50)
    type(c_ptr)       :: userprofil, altabs, ro
  end type

  type, bind(c) :: sosgpu_profile_term      ! one (wavelength, CKD term) of the profile chain
    integer(c_int)    :: lamb1
    integer(c_int)    :: ik(8)
    integer(c_int)    :: absprofil, iprofil
    real(c_double)    :: tr, hr, ta, ha, zmin, zmax
  end type

  type, bind(c) :: sosgpu_aer_component     ! one aerosol mode at one wavelength (SOS_MIE / SOS_GRANU arguments)
    real(c_double)    :: rn, in               ! refractive index, in <= 0
    real(c_double)    :: alpha0, alphaf       ! size-parameter range of its Mie table
    integer(c_int)    :: igranu               ! 1 log-normal (v1 modal radius, v2 sigma), 2 Junge (v1 r0, v2 slope, v3 rmax)
    real(c_double)    :: v1, v2, v3, wa
  end type
  type, bind(c) :: sosgpu_aer_model         ! one wavelength: mixture of components + truncation option
    integer(c_int)    :: ncomp                ! 0: comp(1) as it is; 1..4: mixture
    integer(c_int)    :: comp(4)              ! 0-based indices into the component list
    real(c_double)    :: weight(4)            ! number fractions (N(I)/NTOT, normalised CVI)
    integer(c_int)    :: itronc
  end type
  type, bind(c) :: sosgpu_term_out
    type(c_ptr) :: rec, n_fourier, n_scatter, stop_reason, emoins, eplus, ttot_tronc, ttot_vrai, tauout, ier
  end type

  type, bind(c) :: sosgpu_group_out
    type(c_ptr) :: rec, n_rec, emoins, eplus, ttot_tronc, ttot_vrai, tauout
  end type

  interface
    integer(c_int) function sosgpu_create(ctx, device) bind(c, name="sosgpu_create")
      import :: c_ptr, c_int
      type(c_ptr), intent(out) :: ctx
      integer(c_int), value :: device
    end function
    subroutine sosgpu_destroy(ctx) bind(c, name="sosgpu_destroy")
      import :: c_ptr
      type(c_ptr), value :: ctx
    end subroutine
    integer(c_int) function sosgpu_solve_batch(ctx, optics, noptics, terms, nterm, ngroup, rec_stride, wmax, &
                                               part_only, term_out, group_out) bind(c, name="sosgpu_solve_batch")
      import :: c_ptr, c_int, sosgpu_optics, sosgpu_term, sosgpu_term_out, sosgpu_group_out
      type(c_ptr), value :: ctx
      type(sosgpu_optics), intent(in) :: optics(*)
      integer(c_int), value :: noptics
      type(sosgpu_term), intent(in) :: terms(*)
      integer(c_int), value :: nterm, ngroup, rec_stride, wmax, part_only
      type(sosgpu_term_out), intent(in) :: term_out
      type(sosgpu_group_out), intent(in) :: group_out
    end function
    integer(c_int) function sosgpu_trphi_option(ctx, rec, nrec, nbmu, rmu, tau, tauout, igli, n0, wind, ind_surf, &
                                                ifresnel, itrphi, phios, pas_phi, ipolar, phi_fin, theta_fin, &
                                                up, down, nphi_cap) bind(c, name="sosgpu_trphi_option")
      import :: c_ptr, c_int, c_double
      type(c_ptr), value :: ctx
      real(c_double), intent(in) :: rec(*), rmu(*)
      integer(c_int), value :: nrec, nbmu, igli, n0, ifresnel, itrphi, pas_phi, ipolar, nphi_cap
      real(c_double), value :: tau, tauout, wind, ind_surf, phios
      real(c_double), intent(out) :: phi_fin(*), theta_fin(*), up(*), down(*)
    end function
    ! ---- multi-GPU: one process (MPI rank) per GPU, the library owns the NCCL communicator --------------------------
    integer(c_int) function sosgpu_comm_unique_id(id) bind(c, name="sosgpu_comm_unique_id")
      import :: c_int, c_char
      character(kind=c_char), intent(out) :: id(128)           ! rank 0 creates it; MPI_Bcast the 128 bytes
    end function
    integer(c_int) function sosgpu_comm_init(ctx, nranks, rank, id) bind(c, name="sosgpu_comm_init")
      import :: c_ptr, c_int, c_char
      type(c_ptr), value :: ctx
      integer(c_int), value :: nranks, rank
      character(kind=c_char), intent(in) :: id(128)
    end function
    integer(c_int) function sosgpu_batch_upload(ctx, optics, noptics, terms, nterm, ngroup, batch) &
                                                bind(c, name="sosgpu_batch_upload")
      import :: c_ptr, c_int, sosgpu_optics, sosgpu_term
      type(c_ptr), value :: ctx
      type(sosgpu_optics), intent(in) :: optics(*)
      integer(c_int), value :: noptics
      type(sosgpu_term), intent(in) :: terms(*)
      integer(c_int), value :: nterm, ngroup
      type(c_ptr), intent(out) :: batch
    end function
    integer(c_int) function sosgpu_batch_run(ctx, batch, rec_stride, wmax, part_only, term_out, group_out) &
                                             bind(c, name="sosgpu_batch_run")
      import :: c_ptr, c_int
      type(c_ptr), value :: ctx, batch, term_out, group_out     ! c_loc of the out structures, or c_null_ptr
      integer(c_int), value :: rec_stride, wmax, part_only
    end function
    integer(c_int) function sosgpu_batch_reduce_groups(ctx, batch, root) bind(c, name="sosgpu_batch_reduce_groups")
      import :: c_ptr, c_int
      type(c_ptr), value :: ctx, batch
      integer(c_int), value :: root                             ! ONE ncclReduce of the partial CKD sums + metadata
    end function
    integer(c_int) function sosgpu_batch_groups(ctx, batch, rec_stride, wmax, group_out) bind(c, name="sosgpu_batch_groups")
      import :: c_ptr, c_int, sosgpu_group_out
      type(c_ptr), value :: ctx, batch
      integer(c_int), value :: rec_stride, wmax
      type(sosgpu_group_out), intent(in) :: group_out
    end function
    integer(c_int) function sosgpu_batch_trphi(ctx, batch, igli, wind, ind_surf, ifresnel, itrphi, phios, pas_phi, ipolar, &
                                               nphi_cap, up, down) bind(c, name="sosgpu_batch_trphi")
      import :: c_ptr, c_int, c_double
      type(c_ptr), value :: ctx, batch, up, down
      integer(c_int), value :: igli, ifresnel, itrphi, pas_phi, ipolar, nphi_cap
      real(c_double), value :: wind, ind_surf, phios
    end function
    integer(c_int) function sosgpu_batch_gather_tables(ctx, batch, root, groups_of_rank, nphi, nmax, up, down) &
                                                       bind(c, name="sosgpu_batch_gather_tables")
      import :: c_ptr, c_int
      type(c_ptr), value :: ctx, batch, up, down
      integer(c_int), value :: root, nphi, nmax
      integer(c_int), intent(in) :: groups_of_rank(*)
    end function
    ! SOS_ABSPROFILE -> SOS_PROFILE -> PROFIL_TMP round trip for nterm terms (SOS_PROC.F:3494-3537)
    integer(c_int) function sosgpu_profile_chain(ctx, ckd, atm, terms, nterm, text_hop, tauabs, nt, zprof, h, pcaer, pcmol, ier) &
        bind(c, name="sosgpu_profile_chain")
      import :: c_ptr, c_int, c_double, sosgpu_ckd, sosgpu_gas_profile, sosgpu_profile_term
      type(c_ptr), value :: ctx
      type(sosgpu_ckd), intent(in) :: ckd
      type(sosgpu_gas_profile), intent(in) :: atm
      type(sosgpu_profile_term), intent(in) :: terms(*)
      integer(c_int), value :: nterm, text_hop
      type(c_ptr), value :: tauabs                     ! [50, nterm] or c_null_ptr
      integer(c_int), intent(out) :: nt(*), ier(*)
      real(c_double), intent(out) :: zprof(0:600,*), h(0:600,*), pcaer(0:600,*), pcmol(0:600,*)
    end function
    subroutine sosgpu_batch_free(ctx, batch) bind(c, name="sosgpu_batch_free")
      import :: c_ptr
      type(c_ptr), value :: ctx, batch
    end subroutine
    ! ---- SOS_Up.txt / SOS_Down.txt (SOS_ABS_MAIN.F:2250-2519), byte-compatible ---------------------------------------
    ! aerosol optics of a wavelength list (SOS_MIE -> SOS_GRANU -> mixture -> SOS_DECOMPO_LEGENDRE); outputs may be c_null_ptr
    integer(c_int) function sosgpu_aerosols(ctx, nbmu, xmu, xhr, ncomp, comp, nmodel, models, os_nb, comp_k, comp_phase, &
        comp_ier, scal, coef, phase, model_ier) bind(c, name="sosgpu_aerosols")
      import :: c_ptr, c_int, sosgpu_aer_component, sosgpu_aer_model
      type(c_ptr), value :: ctx
      integer(c_int), value :: nbmu, ncomp, nmodel, os_nb
      type(c_ptr), value :: xmu, xhr                   ! XMU(-nbmu:nbmu), XHR(-nbmu:nbmu)
      type(sosgpu_aer_component), intent(in) :: comp(*)
      type(sosgpu_aer_model), intent(in) :: models(*)
      type(c_ptr), value :: comp_k, comp_phase, comp_ier   ! [3,ncomp], [nang,3,ncomp], [ncomp]
      type(c_ptr), value :: scal, coef, phase, model_ier   ! [8,nmodel], [0:os_nb,6,nmodel], [nang,4,nmodel], [nmodel]
    end function
    integer(c_int) function sosgpu_write_aerosols(path, os_nb, kmat1, kmat2, asym, coef_tronca, piztr, alp, beta11, gamma12, zeta) &
        bind(c, name="sosgpu_write_aerosols")
      import :: c_char, c_int, c_double
      character(kind=c_char), intent(in) :: path(*)    ! NUL-terminated
      integer(c_int), value :: os_nb
      real(c_double), value :: kmat1, kmat2, asym, coef_tronca, piztr
      real(c_double), intent(in) :: alp(0:*), beta11(0:*), gamma12(0:*), zeta(0:*)
    end function
    integer(c_int) function sosgpu_write_updown(fic_up, fic_down, nbmu, itrphi, phios, pas_phi, zout, phi_fin, theta_fin, &
                                                up, down, nphi_cap, fix_sca_index) bind(c, name="sosgpu_write_updown")
      import :: c_int, c_double, c_char
      character(kind=c_char), intent(in) :: fic_up(*), fic_down(*)   ! NUL-terminated
      integer(c_int), value :: nbmu, itrphi, pas_phi, nphi_cap, fix_sca_index
      real(c_double), value :: phios, zout
      real(c_double), intent(in) :: phi_fin(*), theta_fin(*), up(*), down(*)
    end function
  end interface
end module sosgpu_iso_c
