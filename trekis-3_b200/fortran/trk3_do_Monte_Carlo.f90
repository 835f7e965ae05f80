!-----------------------------------------------------------------------------------------------
! trk3_do_Monte_Carlo.f90 -- ISO_C_BINDING shim that lets the UNCHANGED Fortran host of TREKIS-3
! (Universal_MC_for_SHI_MAIN.f90:271-276) run its Monte-Carlo section on the CUDA engine.
!
! It provides MODULE Monte_Carlo with the one public procedure of the reference's module,
!     subroutine do_Monte_Carlo(NMC, SHI, SHI_MFP, ...)            (Monte_Carlo.f90:33, :39-44)
! with the same dummy arguments.  Build the host with this file in place of Monte_Carlo.f90
! (the main program pulls its modules in by #include, Universal_MC_for_SHI_MAIN.f90:48-59) and
! link libtrekis3_gpu.so.
!
! What it does: flattens the derived types (Objects.f90:184-275) into the structs of
! include/trekis3_gpu.h -- shells in (atom, shell) order, matrices [shell][energy], differential
! tables as CSR -- runs this rank's share of the iterations, and adds the packed tally buffer back
! into the caller's Out_* arrays (same element order: the buffer keeps every array in Fortran
! order at the offsets of trk3_tally_layout).  With MPI the existing reductions of the host
! stay valid because every rank returns only its own contribution, as the reference does.
!
! NOTE: this image has no Fortran compiler (SURVEY.md F1): the file is written against the
! reference's type definitions but has not been compiled here.  The C++/Python host in
! trekis-3_b200/csrc/host does the same job for drivers without Fortran and is what the tests use.
!-----------------------------------------------------------------------------------------------
module trk3_gpu_binding
  use iso_c_binding
  implicit none
  integer, parameter :: TRK3_MAX_ATOMS = 8, TRK3_MAX_SHELLS = 32, TRK3_MAX_NT = 256, TRK3_N_TALLIES = 26

  type, bind(C) :: trk3_config                      ! include/trekis3_gpu.h: trk3_config
     real(c_double) :: shi_E, shi_mass, shi_fixed_Zeff
     integer(c_int32_t) :: shi_Z, shi_kind_Zeff
     real(c_double) :: Tim, dt
     integer(c_int32_t) :: dt_flag, include_photons
     real(c_double) :: cut_off, layer, hole_mass, work_function, bar_length, bar_height
     integer(c_int32_t) :: kind_of_EMFP, reserved0
     integer(c_int64_t) :: seed
  end type

  type, bind(C) :: trk3_tables                      ! include/trekis3_gpu.h: trk3_tables
     integer(c_int32_t) :: n_atoms, n_shells, vb_shell, nshl_atom1
     integer(c_int32_t) :: atom_Z(TRK3_MAX_ATOMS), atom_nshl(TRK3_MAX_ATOMS), atom_first(TRK3_MAX_ATOMS)
     real(c_double) :: atom_mass(TRK3_MAX_ATOMS), atom_pers(TRK3_MAX_ATOMS)
     integer(c_int32_t) :: shell_atom(TRK3_MAX_SHELLS), shell_num(TRK3_MAX_SHELLS)
     real(c_double) :: shell_Ip(TRK3_MAX_SHELLS), shell_Nel(TRK3_MAX_SHELLS), shell_auger(TRK3_MAX_SHELLS), shell_radiat(TRK3_MAX_SHELLS)
     integer(c_int32_t) :: n_ei;  type(c_ptr) :: ei_E, ei_L
     integer(c_int32_t) :: n_ee;  type(c_ptr) :: ee_E, ee_L
     integer(c_int32_t) :: n_hi;  type(c_ptr) :: hi_E, hi_L
     integer(c_int32_t) :: n_he;  type(c_ptr) :: he_E, he_L
     integer(c_int32_t) :: n_ph;  type(c_ptr) :: ph_E, ph_L
     integer(c_int32_t) :: n_shi; type(c_ptr) :: shi_E, shi_L, shi_dEdx
     type(c_ptr) :: dshi_off, dshi_E, dshi_L
     type(c_ptr) :: eid_off, eid_hw, eid_L, eed_off, eed_hw, eed_L, hid_off, hid_hw, hid_L, hed_off, hed_hw, hed_L
     integer(c_int32_t) :: n_dos; type(c_ptr) :: dos_E, dos_DOS, dos_int, dos_effm
     integer(c_int32_t) :: n_r;   type(c_ptr) :: out_R, out_V
     integer(c_int32_t) :: shell_kocs(TRK3_MAX_SHELLS)                      ! 1 CDF shell, 2 BEB shell (Target_atoms%KOCS)
     real(c_double) :: shell_Ek(TRK3_MAX_SHELLS), at_dens                   ! Target_atoms%Ek, Matter%At_Dens
     integer(c_int32_t) :: delta_cdf, osc_off(TRK3_MAX_SHELLS + 1)          ! kind_of_DR = 4: oscillators of flat shell s = [osc_off(s), osc_off(s+1))
     type(c_ptr) :: osc_E0, osc_alpha                                       ! Target_atoms%Ritchi%E0 / %alpha, flattened
     integer(c_int32_t) :: n_dsf_e; type(c_ptr) :: dsf_e_dE, dsf_e_emit, dsf_e_absorb, ee_emit, ee_absorb   ! DSF_DEMFP (kind_of_EMFP = 2)
     integer(c_int32_t) :: n_dsf_h; type(c_ptr) :: dsf_h_dE, dsf_h_emit, dsf_h_absorb, he_emit, he_absorb   ! DSF_DEMFP_H
  end type

  type, bind(C) :: trk3_tally_layout                ! include/trekis3_gpu.h: trk3_tally_layout
     integer(c_int32_t) :: Nt, n_r, n_atoms, nshl1, n_dos, reserved
     integer(c_int64_t) :: off(TRK3_N_TALLIES), len(TRK3_N_TALLIES), total
     real(c_double) :: time_grid(TRK3_MAX_NT + 1)
  end type

  interface
     integer(c_int) function trk3_tally_layout_init(cfg, tab, lay) bind(C, name="trk3_tally_layout_init")
       import
       type(trk3_config), intent(in) :: cfg
       type(trk3_tables), intent(in) :: tab
       type(trk3_tally_layout), intent(out) :: lay
     end function
     integer(c_int) function trk3_mc_create(cfg, tab, device, eng) bind(C, name="trk3_mc_create")
       import
       type(trk3_config), intent(in) :: cfg
       type(trk3_tables), intent(in) :: tab
       integer(c_int), value :: device
       type(c_ptr), intent(out) :: eng
     end function
     integer(c_int) function trk3_mc_run(eng, it_begin, it_end, tallies, stats) bind(C, name="trk3_mc_run")
       import
       type(c_ptr), value :: eng
       integer(c_int64_t), value :: it_begin, it_end
       real(c_double), intent(inout) :: tallies(*)
       type(c_ptr), value :: stats
     end function
     function trk3_mc_last_error(eng) bind(C, name="trk3_mc_last_error") result(msg)
       import
       type(c_ptr), value :: eng
       type(c_ptr) :: msg
     end function
     subroutine trk3_mc_destroy(eng) bind(C, name="trk3_mc_destroy")
       import
       type(c_ptr), value :: eng
     end subroutine
  end interface
end module trk3_gpu_binding


MODULE Monte_Carlo
  use Universal_Constants
  use Objects
  use MPI_subroutines, only : Save_error_details
  use trk3_gpu_binding
  implicit none
  private
  public :: do_Monte_Carlo

contains

subroutine do_Monte_Carlo(NMC, SHI, SHI_MFP, diff_SHI_MFP, Target_atoms, Lowest_Ip_At, Lowest_Ip_Shl, CDF_Phonon, &
     Total_el_MFPs, Elastic_MFP, Total_Hole_MFPs, Elastic_Hole_MFP, Total_Photon_MFPs, Mat_DOS, Tim, dt, Matter, NumPar, aidCS, &
     Out_R, Out_V, Out_ne, Out_Ee, Out_nphot, Out_Ephot, Out_Ee_vs_E, Out_Eh_vs_E, Out_Elat, &
     Out_nh, Out_Eh, Out_Ehkin, Out_tot_Ne, Out_tot_Nphot, Out_tot_E, &
     Out_E_e, Out_E_phot, Out_E_at, Out_E_h, Out_Eat_dens, Out_theta, Out_theta_h, Out_theta1, Out_Ne_Em, Out_E_Em, Out_Ee_vs_E_Em, &
     Error_message, DSF_DEMFP, DSF_DEMFP_H, Out_field_all, Out_E_field, Out_diff_coeff, MPI_param)
    integer, intent(in) :: NMC
    type(Ion), intent(in) :: SHI
    type(All_MFP), dimension(:), allocatable, intent(in) :: SHI_MFP, diff_SHI_MFP
    type(Atom), dimension(:), intent(in) :: Target_atoms
    integer, intent(in) :: Lowest_Ip_At, Lowest_Ip_Shl
    type(CDF), intent(in) :: CDF_Phonon
    real(8), intent(in) :: Tim, dt
    type(All_MFP), dimension(:), intent(in), target :: Total_el_MFPs, Total_Hole_MFPs, Total_Photon_MFPs
    type(MFP_elastic), intent(in), target :: Elastic_MFP, Elastic_Hole_MFP
    type(Solid), intent(in) :: Matter
    type(Density_of_states), intent(in) :: Mat_DOS
    type(Flag), intent(inout) :: NumPar
    real(8), dimension(:), intent(inout) :: Out_R, Out_V
    real(8), dimension(:,:), intent(inout) :: Out_ne, Out_Ee, Out_nphot, Out_Ephot, Out_Ee_vs_E, Out_Eh_vs_E, Out_Elat, Out_Eat_dens
    real(8), dimension(:,:,:,:), intent(inout) :: Out_nh, Out_Eh, Out_Ehkin
    real(8), dimension(:), intent(inout) :: Out_tot_Ne, Out_tot_Nphot, Out_tot_E, Out_E_e, Out_E_phot, Out_E_at
    real(8), dimension(:,:,:), intent(inout) :: Out_E_h
    real(8), dimension(:,:), intent(inout) :: Out_theta, Out_theta_h
    real(8), dimension(:), intent(inout) :: Out_theta1
    real(8), dimension(:,:), intent(inout) :: Out_field_all, Out_Ee_vs_E_Em
    real(8), dimension(:), intent(inout) :: Out_Ne_Em, Out_E_Em, Out_E_field
    type(Error_handling), intent(inout) :: Error_message
    type(Differential_MFP), dimension(:), intent(in) :: DSF_DEMFP, DSF_DEMFP_H
    type(All_diff_CS), intent(in) :: aidCS
    real(8), dimension(:), intent(inout) :: Out_diff_coeff
    type(Used_MPI_parameters), intent(inout) :: MPI_param
    !------------------------------------------
    type(trk3_config) :: cfg
    type(trk3_tables) :: tab
    type(trk3_tally_layout) :: lay
    type(c_ptr) :: eng
    integer(c_int) :: ierr
    integer(c_int64_t) :: it0, it1
    integer :: Nat, NS, a, k, q, i, n
    real(c_double), allocatable, target :: ei_E(:), ei_L(:), ee_E(:), ee_L(:), hi_E(:), hi_L(:), he_E(:), he_L(:), ph_E(:), ph_L(:)
    real(c_double), allocatable, target :: shi_E(:), shi_L(:), shi_dEdx(:), dshi_E(:), dshi_L(:)
    real(c_double), allocatable, target :: eid_hw(:), eid_L(:), eed_hw(:), eed_L(:), hid_hw(:), hid_L(:), hed_hw(:), hed_L(:)
    integer(c_int64_t), allocatable, target :: dshi_off(:), eid_off(:), eed_off(:), hid_off(:), hed_off(:)
    real(c_double), allocatable, target :: dos_E(:), dos_DOS(:), dos_int(:), dos_effm(:), R_c(:), V_c(:)
    real(c_double), allocatable, target :: buf(:), osc_E0(:), osc_alpha(:)
    real(c_double), allocatable, target :: dsf_e_dE(:), dsf_e_emit(:), dsf_e_absorb(:), ee_emit(:), ee_absorb(:)
    real(c_double), allocatable, target :: dsf_h_dE(:), dsf_h_emit(:), dsf_h_absorb(:), he_emit(:), he_absorb(:)
    integer :: n_osc, l

    Nat = size(Target_atoms)
    NS = 0
    do a = 1, Nat
       NS = NS + size(Target_atoms(a)%Ip)
    enddo
    if (Nat > TRK3_MAX_ATOMS .or. NS > TRK3_MAX_SHELLS) then
       call Save_error_details(Error_message, 60, 'trk3 GPU engine: too many atoms or shells for the C ABI', MPI_param)
       return
    endif

    ! ---- scalars (trk3_config)
    cfg%shi_E = SHI%E; cfg%shi_mass = SHI%Mass; cfg%shi_fixed_Zeff = SHI%fixed_Zeff
    cfg%shi_Z = SHI%Zat; cfg%shi_kind_Zeff = SHI%Kind_Zeff
    cfg%Tim = Tim; cfg%dt = dt; cfg%dt_flag = NumPar%dt_flag
    cfg%include_photons = merge(1, 0, NumPar%include_photons)
    cfg%cut_off = Matter%cut_off; cfg%layer = Matter%Layer; cfg%hole_mass = Matter%hole_mass
    cfg%work_function = Matter%work_function; cfg%bar_length = Matter%bar_length; cfg%bar_height = Matter%bar_height
    cfg%kind_of_EMFP = NumPar%kind_of_EMFP; cfg%reserved0 = 0
    cfg%seed = 20260101_c_int64_t

    ! ---- target: shells flattened in (atom, shell) order, 0-based indices in the C structs
    tab%n_atoms = Nat; tab%n_shells = NS; tab%nshl_atom1 = size(Target_atoms(1)%Ip)
    q = 0
    do a = 1, Nat
       tab%atom_Z(a) = Target_atoms(a)%Zat; tab%atom_nshl(a) = size(Target_atoms(a)%Ip); tab%atom_first(a) = q
       tab%atom_mass(a) = Target_atoms(a)%Mass; tab%atom_pers(a) = Target_atoms(a)%Pers
       do k = 1, size(Target_atoms(a)%Ip)
          q = q + 1
          tab%shell_atom(q) = a - 1; tab%shell_num(q) = k - 1
          tab%shell_Ip(q) = Target_atoms(a)%Ip(k); tab%shell_Nel(q) = Target_atoms(a)%Nel(k)
          tab%shell_auger(q) = Target_atoms(a)%Auger(k); tab%shell_radiat(q) = Target_atoms(a)%Radiat(k)
          if (a == Lowest_Ip_At .and. k == Lowest_Ip_Shl) tab%vb_shell = q - 1
          tab%shell_kocs(q) = Target_atoms(a)%KOCS(k)           ! 2: BEB shell (negative designator in the .cdf)
          tab%shell_Ek(q) = Target_atoms(a)%Ek(k)
       enddo
    enddo
    tab%at_dens = Matter%At_Dens
    ! ---- delta-function CDF (kind_of_DR = 4): positions and weights of the oscillators, flattened in shell order
    tab%delta_cdf = merge(1, 0, NumPar%kind_of_DR == 4)
    n_osc = 0
    if (tab%delta_cdf == 1) then
       do a = 1, Nat
          do k = 1, size(Target_atoms(a)%Ip)
             n_osc = n_osc + size(Target_atoms(a)%Ritchi(k)%E0)
          enddo
       enddo
    endif
    allocate(osc_E0(max(n_osc, 1)), osc_alpha(max(n_osc, 1)))
    osc_E0 = 0.0d0; osc_alpha = 0.0d0
    tab%osc_off(:) = 0
    q = 0; n_osc = 0
    do a = 1, Nat
       do k = 1, size(Target_atoms(a)%Ip)
          q = q + 1
          tab%osc_off(q) = n_osc
          if (tab%delta_cdf == 1) then
             do l = 1, size(Target_atoms(a)%Ritchi(k)%E0)
                n_osc = n_osc + 1
                osc_E0(n_osc) = Target_atoms(a)%Ritchi(k)%E0(l); osc_alpha(n_osc) = Target_atoms(a)%Ritchi(k)%alpha(l)
             enddo
          endif
       enddo
    enddo
    tab%osc_off(q + 1:) = n_osc
    tab%osc_E0 = c_loc(osc_E0); tab%osc_alpha = c_loc(osc_alpha)
    ! ---- DSF elastic scattering (kind_of_EMFP = 2): the rows of DSF_DEMFP / DSF_DEMFP_H, [particle energy][transferred energy]
    call flatten_dsf(DSF_DEMFP, Elastic_MFP%Emit%L, Elastic_MFP%Absorb%L, tab%n_dsf_e, dsf_e_dE, dsf_e_emit, dsf_e_absorb, ee_emit, ee_absorb)
    call flatten_dsf(DSF_DEMFP_H, Elastic_Hole_MFP%Emit%L, Elastic_Hole_MFP%Absorb%L, tab%n_dsf_h, dsf_h_dE, dsf_h_emit, dsf_h_absorb, he_emit, he_absorb)
    tab%dsf_e_dE = c_loc(dsf_e_dE); tab%dsf_e_emit = c_loc(dsf_e_emit); tab%dsf_e_absorb = c_loc(dsf_e_absorb)
    tab%ee_emit = c_loc(ee_emit); tab%ee_absorb = c_loc(ee_absorb)
    tab%dsf_h_dE = c_loc(dsf_h_dE); tab%dsf_h_emit = c_loc(dsf_h_emit); tab%dsf_h_absorb = c_loc(dsf_h_absorb)
    tab%he_emit = c_loc(he_emit); tab%he_absorb = c_loc(he_absorb)

    ! ---- mean free paths: one energy grid per family, rows [shell][energy]
    call flatten_mfp(Total_el_MFPs, ei_E, ei_L)
    call flatten_mfp(Total_Hole_MFPs, hi_E, hi_L)
    call flatten_mfp(SHI_MFP, shi_E, shi_L, shi_dEdx)
    if (NumPar%include_photons) then
       call flatten_mfp(Total_Photon_MFPs, ph_E, ph_L)
    else
       allocate(ph_E(0), ph_L(0))
    endif
    ee_E = Elastic_MFP%Total%E;      ee_L = Elastic_MFP%Total%L
    he_E = Elastic_Hole_MFP%Total%E; he_L = Elastic_Hole_MFP%Total%L

    ! ---- differential tables as CSR: row r occupies [off(r), off(r+1)) (0-based offsets)
    n = 0
    do a = 1, Nat
       do k = 1, size(Target_atoms(a)%Ip)
          n = n + size(diff_SHI_MFP(a)%ELMFP(k)%E)
       enddo
    enddo
    allocate(dshi_off(NS + 1), dshi_E(n), dshi_L(n))
    dshi_off(1) = 0; q = 0; n = 0
    do a = 1, Nat
       do k = 1, size(Target_atoms(a)%Ip)
          q = q + 1
          i = size(diff_SHI_MFP(a)%ELMFP(k)%E)
          dshi_E(n + 1 : n + i) = diff_SHI_MFP(a)%ELMFP(k)%E; dshi_L(n + 1 : n + i) = diff_SHI_MFP(a)%ELMFP(k)%L
          n = n + i
          dshi_off(q + 1) = n
       enddo
    enddo
    call flatten_dcs_all(aidCS%EIdCS, eid_off, eid_hw, eid_L)
    call flatten_dcs(aidCS%EEdCS, eed_off, eed_hw, eed_L)
    call flatten_dcs(aidCS%HIdCS, hid_off, hid_hw, hid_L)
    call flatten_dcs(aidCS%HEdCS, hed_off, hed_hw, hed_L)

    dos_E = Mat_DOS%E; dos_DOS = Mat_DOS%DOS; dos_int = Mat_DOS%int_DOS; dos_effm = Mat_DOS%Eff_m
    R_c = Out_R; V_c = Out_V

    tab%n_ei = size(ei_E);   tab%ei_E = c_loc(ei_E);   tab%ei_L = c_loc(ei_L)
    tab%n_ee = size(ee_E);   tab%ee_E = c_loc(ee_E);   tab%ee_L = c_loc(ee_L)
    tab%n_hi = size(hi_E);   tab%hi_E = c_loc(hi_E);   tab%hi_L = c_loc(hi_L)
    tab%n_he = size(he_E);   tab%he_E = c_loc(he_E);   tab%he_L = c_loc(he_L)
    tab%n_ph = size(ph_E);   tab%ph_E = c_null_ptr;    tab%ph_L = c_null_ptr
    if (size(ph_E) > 0) then
       tab%ph_E = c_loc(ph_E); tab%ph_L = c_loc(ph_L)
    endif
    tab%n_shi = size(shi_E); tab%shi_E = c_loc(shi_E); tab%shi_L = c_loc(shi_L); tab%shi_dEdx = c_loc(shi_dEdx)
    tab%dshi_off = c_loc(dshi_off); tab%dshi_E = c_loc(dshi_E); tab%dshi_L = c_loc(dshi_L)
    tab%eid_off = c_loc(eid_off); tab%eid_hw = c_loc(eid_hw); tab%eid_L = c_loc(eid_L)
    tab%eed_off = c_loc(eed_off); tab%eed_hw = c_loc(eed_hw); tab%eed_L = c_loc(eed_L)
    tab%hid_off = c_loc(hid_off); tab%hid_hw = c_loc(hid_hw); tab%hid_L = c_loc(hid_L)
    tab%hed_off = c_loc(hed_off); tab%hed_hw = c_loc(hed_hw); tab%hed_L = c_loc(hed_L)
    tab%n_dos = size(dos_E); tab%dos_E = c_loc(dos_E); tab%dos_DOS = c_loc(dos_DOS); tab%dos_int = c_loc(dos_int); tab%dos_effm = c_loc(dos_effm)
    tab%n_r = size(R_c); tab%out_R = c_loc(R_c); tab%out_V = c_loc(V_c)

    ! ---- run this rank's share of the GLOBAL iteration indices (the random streams are keyed by the global index,
    !      so the union over ranks does not depend on the split; replaces the cyclic split of Monte_Carlo.f90:114-120)
    ierr = trk3_tally_layout_init(cfg, tab, lay)
    if (ierr /= 0) then
       call Save_error_details(Error_message, 61, 'trk3 GPU engine: invalid time grid', MPI_param)
       return
    endif
    allocate(buf(lay%total)); buf = 0.0d0
    ierr = trk3_mc_create(cfg, tab, -1_c_int, eng)
    if (ierr == 0) then
       it0 = int(MPI_param%process_rank, c_int64_t) * NMC / max(1, MPI_param%size_of_cluster)
       it1 = int(MPI_param%process_rank + 1, c_int64_t) * NMC / max(1, MPI_param%size_of_cluster)
       ierr = trk3_mc_run(eng, it0, it1, buf, c_null_ptr)
    endif
    if (ierr /= 0) call Save_error_details(Error_message, 62, 'trk3 GPU engine failed (see trk3_mc_last_error)', MPI_param)
    call trk3_mc_destroy(eng)
    if (ierr /= 0) return

    ! ---- add the contributions to the caller's arrays: identical shapes and element order (enum trk3_tally_id, 0-based)
    Out_ne        = Out_ne        + reshape(slice(0),  shape(Out_ne))
    Out_Ee        = Out_Ee        + reshape(slice(1),  shape(Out_Ee))
    Out_nphot     = Out_nphot     + reshape(slice(2),  shape(Out_nphot))
    Out_Ephot     = Out_Ephot     + reshape(slice(3),  shape(Out_Ephot))
    Out_Ee_vs_E   = Out_Ee_vs_E   + reshape(slice(4),  shape(Out_Ee_vs_E))
    Out_Eh_vs_E   = Out_Eh_vs_E   + reshape(slice(5),  shape(Out_Eh_vs_E))
    Out_Elat      = Out_Elat      + reshape(slice(6),  shape(Out_Elat))
    Out_nh        = Out_nh        + reshape(slice(7),  shape(Out_nh))
    Out_Eh        = Out_Eh        + reshape(slice(8),  shape(Out_Eh))
    Out_Ehkin     = Out_Ehkin     + reshape(slice(9),  shape(Out_Ehkin))
    Out_tot_Ne    = Out_tot_Ne    + slice(10)
    Out_tot_Nphot = Out_tot_Nphot + slice(11)
    Out_tot_E     = Out_tot_E     + slice(12)
    Out_E_e       = Out_E_e       + slice(13)
    Out_E_phot    = Out_E_phot    + slice(14)
    Out_E_at      = Out_E_at      + slice(15)
    Out_E_h       = Out_E_h       + reshape(slice(16), shape(Out_E_h))
    Out_Eat_dens  = Out_Eat_dens  + reshape(slice(17), shape(Out_Eat_dens))
    Out_theta     = Out_theta     + reshape(slice(18), shape(Out_theta))
    Out_theta_h   = Out_theta_h   + reshape(slice(19), shape(Out_theta_h))
    Out_Ne_Em     = Out_Ne_Em     + slice(20)
    Out_E_Em      = Out_E_Em      + slice(21)
    Out_Ee_vs_E_Em = Out_Ee_vs_E_Em + reshape(slice(22), shape(Out_Ee_vs_E_Em))
    Out_field_all = Out_field_all + reshape(slice(23), shape(Out_field_all))
    Out_E_field   = Out_E_field   + slice(24)
    Out_diff_coeff = Out_diff_coeff + slice(25)

contains

    function slice(id) result(v)             ! array `id` (0-based, enum trk3_tally_id) of the packed buffer
       integer, intent(in) :: id
       real(8), allocatable :: v(:)
       v = buf(lay%off(id + 1) + 1 : lay%off(id + 1) + lay%len(id + 1))
    end function slice

    subroutine flatten_mfp(T, E, L, dEdx)    ! All_MFP(atom)%ELMFP(shell)%{E,L,dEdx} -> shared grid + rows [shell][energy]
       type(All_MFP), dimension(:), intent(in) :: T
       real(c_double), allocatable, intent(out) :: E(:), L(:)
       real(c_double), allocatable, intent(out), optional :: dEdx(:)
       integer :: a1, k1, q1, N1
       E = T(1)%ELMFP(1)%E
       N1 = size(E)
       allocate(L(NS * N1))
       if (present(dEdx)) allocate(dEdx(NS * N1))
       q1 = 0
       do a1 = 1, size(T)
          do k1 = 1, size(T(a1)%ELMFP)
             L(q1 * N1 + 1 : (q1 + 1) * N1) = T(a1)%ELMFP(k1)%L
             if (present(dEdx)) dEdx(q1 * N1 + 1 : (q1 + 1) * N1) = T(a1)%ELMFP(k1)%dEdx
             q1 = q1 + 1
          enddo
       enddo
    end subroutine flatten_mfp

    subroutine flatten_dsf(D, Lem, Lab, nW, dE, em, ab, Lem_c, Lab_c)    ! Differential_MFP(iE)%{dE,dL_emit,dL_absorb} -> rows [iE][transfer]
       type(Differential_MFP), dimension(:), intent(in) :: D
       real(8), dimension(:), allocatable, intent(in) :: Lem, Lab           ! Elastic_MFP%Emit%L, %Absorb%L (allocated for kind_of_EMFP = 2 only)
       integer(c_int32_t), intent(out) :: nW
       real(c_double), allocatable, intent(out) :: dE(:), em(:), ab(:), Lem_c(:), Lab_c(:)
       integer :: i1
       nW = 0
       if (size(D) > 0 .and. NumPar%kind_of_EMFP == 2) nW = size(D(1)%dE)
       if (nW == 0) then
          allocate(dE(1), em(1), ab(1), Lem_c(1), Lab_c(1))
          dE = 0.0d0; em = 0.0d0; ab = 0.0d0; Lem_c = 0.0d0; Lab_c = 0.0d0
          return
       endif
       allocate(dE(size(D) * nW), em(size(D) * nW), ab(size(D) * nW), Lem_c(size(D)), Lab_c(size(D)))
       do i1 = 1, size(D)
          dE((i1 - 1) * nW + 1 : i1 * nW) = D(i1)%dE
          em((i1 - 1) * nW + 1 : i1 * nW) = D(i1)%dL_emit
          ab((i1 - 1) * nW + 1 : i1 * nW) = D(i1)%dL_absorb
       enddo
       Lem_c = Lem; Lab_c = Lab
    end subroutine flatten_dsf

    subroutine flatten_dcs(D, off, hw, L)    ! diff_CS%diffCS(iE)%{hw,dsdhw} -> CSR
       type(diff_CS), intent(in) :: D
       integer(c_int64_t), allocatable, intent(out) :: off(:)
       real(c_double), allocatable, intent(out) :: hw(:), L(:)
       integer :: i1, n1, m1
       n1 = 0
       do i1 = 1, size(D%diffCS)
          n1 = n1 + size(D%diffCS(i1)%hw)
       enddo
       allocate(off(size(D%diffCS) + 1), hw(n1), L(n1))
       off(1) = 0; n1 = 0
       do i1 = 1, size(D%diffCS)
          m1 = size(D%diffCS(i1)%hw)
          hw(n1 + 1 : n1 + m1) = D%diffCS(i1)%hw; L(n1 + 1 : n1 + m1) = D%diffCS(i1)%dsdhw
          n1 = n1 + m1
          off(i1 + 1) = n1
       enddo
    end subroutine flatten_dcs

    subroutine flatten_dcs_all(A, off, hw, L)    ! EIdCS(atom)%Int_diff_CS(shell)%diffCS(iE): rows ordered (shell, iE)
       type(Array_dCS), dimension(:), intent(in) :: A
       integer(c_int64_t), allocatable, intent(out) :: off(:)
       real(c_double), allocatable, intent(out) :: hw(:), L(:)
       integer :: a1, k1, i1, n1, m1, r1, nrows
       n1 = 0; nrows = 0
       do a1 = 1, size(A)
          do k1 = 1, size(A(a1)%Int_diff_CS)
             do i1 = 1, size(A(a1)%Int_diff_CS(k1)%diffCS)
                n1 = n1 + size(A(a1)%Int_diff_CS(k1)%diffCS(i1)%hw); nrows = nrows + 1
             enddo
          enddo
       enddo
       allocate(off(nrows + 1), hw(n1), L(n1))
       off(1) = 0; n1 = 0; r1 = 0
       do a1 = 1, size(A)
          do k1 = 1, size(A(a1)%Int_diff_CS)
             do i1 = 1, size(A(a1)%Int_diff_CS(k1)%diffCS)
                m1 = size(A(a1)%Int_diff_CS(k1)%diffCS(i1)%hw)
                hw(n1 + 1 : n1 + m1) = A(a1)%Int_diff_CS(k1)%diffCS(i1)%hw
                L(n1 + 1 : n1 + m1) = A(a1)%Int_diff_CS(k1)%diffCS(i1)%dsdhw
                n1 = n1 + m1; r1 = r1 + 1
                off(r1 + 1) = n1
             enddo
          enddo
       enddo
    end subroutine flatten_dcs_all

end subroutine do_Monte_Carlo

END MODULE Monte_Carlo
