!> Drop-in replacement for module qcxms_tblite (reference src/tblite.f90): same public names and the same
!> get_xtb_egrad signature, forwarding to the C ABI of libqcxms_b200 (include/qcxms_b200.h) via iso_c_binding.
!> To use it, replace src/tblite.f90 in src/meson.build by this file and link libqcxms_b200.so
!> (see INTEGRATION.md).  Not compiled in the build container (no Fortran compiler there, SURVEY.md F4).
module qcxms_tblite
   use, intrinsic :: iso_c_binding
   implicit none
   private

   public :: get_xtb_egrad
   public :: gfn1_xtb, gfn2_xtb, ipea1_xtb
   public :: md_config, md_result
   public :: ensemble_create, ensemble_destroy, ensemble_set_trajectory, ensemble_run_md, ensemble_get_result

   integer, parameter :: wp = selected_real_kind(15)

   !> same enumerated method selector as the reference (src/tblite.f90:29-40)
   type :: method_selector
      integer :: id
   end type method_selector
   type(method_selector), parameter :: gfn2_xtb = method_selector(2)
   type(method_selector), parameter :: gfn1_xtb = method_selector(1)
   type(method_selector), parameter :: ipea1_xtb = method_selector(11)

   !> qcxms_b200_md_config_t
   type, bind(c) :: md_config
      integer(c_int32_t) :: method_id, mchrg, nfragexit, exit_rules, nmax, isec
      real(c_double) :: tstep, etemp_in, ieetemp, ax
   end type md_config

   !> qcxms_b200_md_result_t
   type, bind(c) :: md_result
      integer(c_int32_t) :: mdok, fragstate, nstep, nfrag, status, scc_iter_total
      real(c_double) :: Tav, Epav, Ekav, aTlast, dtime, ttime, Epot, Ekin
   end type md_result

   ! qcxms_b200_cid_config_t / qcxms_b200_cid_result_t (include/qcxms_b200.h)
   type, bind(c) :: cid_config
      integer(c_int32_t) :: method_id, mchrg, gas_z, eexact, manual_dist, ntot
      real(c_double) :: gas_mass, tstep, etemp, elab, ecom
   end type cid_config
   type, bind(c) :: cid_result
      integer(c_int32_t) :: stopcid, nstep, nfrag, collided, status, scc_iter_total
      real(c_double) :: velo_cm, aTlast, ttime, epot, direc(3)
   end type cid_result

   interface
      integer(c_int) function qcxms_b200_egrad(nat, num, xyz, charge, multiplicity, method_id, etemp, &
            & qat, energy, gradient, stat) bind(c, name="qcxms_b200_egrad")
         import :: c_int, c_int32_t, c_double
         integer(c_int), value :: nat, charge, multiplicity, method_id
         integer(c_int32_t), intent(in) :: num(*)
         real(c_double), intent(in) :: xyz(3, *)
         real(c_double), value :: etemp
         real(c_double), intent(out) :: qat(*), energy, gradient(3, *)
         integer(c_int32_t), intent(out) :: stat
      end function qcxms_b200_egrad

      integer(c_int) function qcxms_b200_basis_size(nat, num, method_id, nao) bind(c, name="qcxms_b200_basis_size")
         import :: c_int, c_int32_t
         integer(c_int), value :: nat, method_id
         integer(c_int32_t), intent(in) :: num(*)
         integer(c_int32_t), intent(out) :: nao
      end function qcxms_b200_basis_size

      integer(c_int) function qcxms_b200_egrad_spec(nat, num, xyz, charge, multiplicity, method_id, etemp, qat, energy, gradient, &
            & stat, nao, ihomo, emo, focc, qmo) bind(c, name="qcxms_b200_egrad_spec")
         import :: c_int, c_int32_t, c_double
         integer(c_int), value :: nat, charge, multiplicity, method_id
         integer(c_int32_t), intent(in) :: num(*)
         real(c_double), intent(in) :: xyz(3, *)
         real(c_double), value :: etemp
         real(c_double), intent(out) :: qat(*), energy, gradient(3, *), emo(*), focc(*), qmo(nat, *)
         integer(c_int32_t), intent(out) :: stat, nao, ihomo
      end function qcxms_b200_egrad_spec

      integer(c_int) function ensemble_create(cfg, ntraj, nat, num, mass, device, handle) &
            & bind(c, name="qcxms_b200_ensemble_create")
         import :: c_int, c_int32_t, c_double, c_ptr, md_config
         type(md_config), intent(in) :: cfg
         integer(c_int), value :: ntraj, nat, device
         integer(c_int32_t), intent(in) :: num(*)
         real(c_double), intent(in) :: mass(*)
         type(c_ptr), intent(out) :: handle
      end function ensemble_create

      integer(c_int) function ensemble_destroy(handle) bind(c, name="qcxms_b200_ensemble_destroy")
         import :: c_int, c_ptr
         type(c_ptr), value :: handle
      end function ensemble_destroy

      integer(c_int) function ensemble_set_trajectory(handle, itrj, xyz, velo, velof, eimp, tadd) &
            & bind(c, name="qcxms_b200_ensemble_set_trajectory")
         import :: c_int, c_double, c_ptr
         type(c_ptr), value :: handle
         integer(c_int), value :: itrj
         real(c_double), intent(in) :: xyz(3, *), velo(3, *), velof(*)
         real(c_double), value :: eimp, tadd
      end function ensemble_set_trajectory

      integer(c_int) function ensemble_run_md(handle, max_steps, steps_done) bind(c, name="qcxms_b200_ensemble_run_md")
         import :: c_int, c_int64_t, c_ptr
         type(c_ptr), value :: handle
         integer(c_int), value :: max_steps
         integer(c_int64_t), intent(out) :: steps_done
      end function ensemble_run_md

      integer(c_int) function ensemble_get_result(handle, itrj, xyz, velo, grad, list, achrg, axyz, res) &
            & bind(c, name="qcxms_b200_ensemble_get_result")
         import :: c_int, c_int32_t, c_double, c_ptr, md_result
         type(c_ptr), value :: handle
         integer(c_int), value :: itrj
         real(c_double), intent(out) :: xyz(3, *), velo(3, *), grad(3, *), achrg(*), axyz(3, *)
         integer(c_int32_t), intent(out) :: list(*)
         type(md_result), intent(out) :: res
      end function ensemble_get_result
      ! one collision of cid() for a batch of ions (reference src/cid.f90:24-28); rnd(9,ntraj): the uniform random
      ! numbers the reference draws inside the call (euler_rotation a,b,c; vary_energies dum,dum2; placement f,g,lmin,lpos)
      integer(c_int) function cid_batch(cfg, ntraj, nuc, num, mass, icoll, xyz, velo, rnd, velo_cm, direc, collided, &
            & grad, achrg, axyz, list, res, device) bind(c, name="qcxms_b200_cid_batch")
         import :: c_int, c_int32_t, c_double, cid_config, cid_result
         type(cid_config), intent(in) :: cfg
         integer(c_int), value :: ntraj, nuc, icoll, device
         integer(c_int32_t), intent(in) :: num(*)
         real(c_double), intent(in) :: mass(*), rnd(9, *), velo_cm(*)
         real(c_double), intent(inout) :: xyz(3, nuc, *), velo(3, nuc, *), direc(3, *)
         integer(c_int32_t), intent(inout) :: collided(*)
         real(c_double), intent(out) :: grad(3, nuc, *), achrg(nuc, *), axyz(3, nuc, *)
         integer(c_int32_t), intent(out) :: list(nuc, *)
         type(cid_result), intent(out) :: res(*)
      end function cid_batch
      ! ---- round 2: ground-state md() modes, ESI heating MD, the C-ABI collective (include/qcxms_b200.h)
      integer(c_int) function ensemble_set_gs_mode(handle, it, tsoll) bind(c, name="qcxms_b200_ensemble_set_gs_mode")
         import :: c_int, c_double, c_ptr
         type(c_ptr), value :: handle
         integer(c_int), value :: it               ! -1 equilibration, 0 sampling (src/main.F90:545-567), 1 production
         real(c_double), value :: tsoll
      end function ensemble_set_gs_mode

      integer(c_int) function ensemble_get_gs(handle, itrj, first, count, xyzvelo) bind(c, name="qcxms_b200_ensemble_get_gs")
         import :: c_int, c_double, c_ptr
         type(c_ptr), value :: handle
         integer(c_int), value :: itrj, first, count
         real(c_double), intent(out) :: xyzvelo(6, *)      ! (x,y,z,vx,vy,vz) per atom and step: the records of qcxms.gs
      end function ensemble_get_gs

      integer(c_int) function ensemble_set_esi(handle, tscale) bind(c, name="qcxms_b200_ensemble_set_esi")
         import :: c_int, c_double, c_ptr
         type(c_ptr), value :: handle
         real(c_double), value :: tscale           ! Tsoll of md(..., starting_md = .true.) (src/main.F90:1357-1362)
      end function ensemble_set_esi

      integer(c_int) function comm_unique_id(id128) bind(c, name="qcxms_b200_comm_unique_id")
         import :: c_int, c_char
         character(kind=c_char), intent(out) :: id128(128)
      end function comm_unique_id

      integer(c_int) function comm_create(id128, nranks, rank, device, comm) bind(c, name="qcxms_b200_comm_create")
         import :: c_int, c_char, c_ptr
         character(kind=c_char), intent(in) :: id128(128)
         integer(c_int), value :: nranks, rank, device
         type(c_ptr), intent(out) :: comm
      end function comm_create

      integer(c_int) function comm_allreduce_sum(comm, inout, n) bind(c, name="qcxms_b200_comm_allreduce_sum")
         import :: c_int, c_double, c_ptr
         type(c_ptr), value :: comm
         real(c_double), intent(inout) :: inout(*)
         integer(c_int), value :: n
      end function comm_allreduce_sum

      integer(c_int) function comm_destroy(comm) bind(c, name="qcxms_b200_comm_destroy")
         import :: c_int, c_ptr
         type(c_ptr), value :: comm
      end function comm_destroy

   end interface

contains

!> Entry point for QCxMS to request calculations (signature of reference src/tblite.f90:65-66)
subroutine get_xtb_egrad(num, xyz, charge, multiplicity, method, etemp, &
      & output_file, qat, energy, gradient, stat, spec_calc)
   integer, intent(in) :: num(:)
   real(wp), intent(in) :: xyz(:, :)
   integer, intent(in) :: charge
   integer, intent(in) :: multiplicity
   type(method_selector), intent(in) :: method
   real(wp), intent(in) :: etemp
   character(len=*), intent(in) :: output_file
   real(wp), intent(out) :: qat(:)
   real(wp), intent(out) :: energy
   real(wp), intent(out) :: gradient(:, :)
   integer, intent(out) :: stat
   logical :: spec_calc

   integer(c_int32_t) :: cstat, cnao, cihomo
   integer(c_int) :: rc
   integer :: unit, nat, k, j, io_tmp, io_mspec
   real(c_double), allocatable :: emo(:), focc(:), qmo(:, :)
   real(wp), parameter :: autoev = 27.21138505_wp

   ! the reference redirects tblite's printout to output_file on every call (src/tblite.f90:108,173);
   ! keep the file so that downstream scripts find it, but write a one-line stub only
   open(newunit=unit, file=output_file)
   write(unit, '(a)') "[Info] qcxms_b200: GFN-xTB calculation on the GPU"
   close(unit)

   nat = size(num)
   if (.not. spec_calc) then
      rc = qcxms_b200_egrad(int(nat, c_int), int(num, c_int32_t), xyz, int(charge, c_int), &
         & int(multiplicity, c_int), int(method%id, c_int), real(etemp, c_double), qat, energy, gradient, cstat)
      stat = int(cstat)
      if (rc /= 0) stat = -1   ! stat_fatal (src/tblite.f90:55)
      return
   end if

   ! spec_calc: MO energies, occupations and atomic populations for getspec; the files are written with the statements of
   ! write_qmo (reference src/mo_energ.f90:56-74) so that their layout is the compiler's own list-directed one
   rc = qcxms_b200_basis_size(int(nat, c_int), int(num, c_int32_t), int(method%id, c_int), cnao)
   if (rc /= 0) then
      stat = -1
      return
   end if
   allocate(emo(cnao), focc(cnao), qmo(nat, cnao))      ! C layout [nao][nat]
   rc = qcxms_b200_egrad_spec(int(nat, c_int), int(num, c_int32_t), xyz, int(charge, c_int), int(multiplicity, c_int), &
      & int(method%id, c_int), real(etemp, c_double), qat, energy, gradient, cstat, cnao, cihomo, emo, focc, qmo)
   stat = int(cstat)
   if (rc /= 0) stat = -1
   if (stat /= 0) return
   open(file='tmp.mspec', newunit=io_tmp, status='replace')
   open(file='qcxms.Mspec.tbxtb', newunit=io_mspec, status='replace')
   write(io_mspec,*) int(cnao), int(cihomo)
   do k = 1, cnao
      write (io_mspec, *)
      write (io_mspec,'(1(i3 ,1x, f10.3))') k, emo(k)*autoev
      write (io_mspec,'(1(1x, f6.2))')     focc(k)
      write (io_mspec,'(10(1x, f6.2))')   (qmo(j,k)*100.0d0, j=1,nat)
   enddo
   write(io_tmp,*) int(cnao), int(cihomo)
   do k = 1, cnao
      write(io_tmp,*) emo(k)*autoev
      write(io_tmp,*) focc(k)
      do j = 1, nat
         write (io_tmp, *) qmo(j,k)
      end do
   enddo
   close(io_tmp)
   close(io_mspec)
end subroutine get_xtb_egrad

end module qcxms_tblite
