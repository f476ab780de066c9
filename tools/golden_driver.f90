!> Golden-vector generator to be run wherever tblite v0.2.1 can be built (SURVEY.md 8c): it repeats the call
!> protocol of get_xtb_egrad (reference src/tblite.f90:95-151: new structure, new calculator, NEW ZEROED
!> wavefunction, accuracy 1.0, kt = etemp*ktoau) and prints energy / gradient / charges with full precision
!> in the JSON layout of tests/golden/egrad_golden.json.
!> Usage: golden_driver <xyz-file in bohr> <charge> <etemp> [method id: 2 (GFN2, default) | 1 (GFN1) | 11 (IPEA1)]
program golden_driver
   use mctc_env, only : wp, error_type
   use mctc_io, only : structure_type, read_structure
   use tblite_context_type, only : context_type
   use tblite_wavefunction_type, only : wavefunction_type, new_wavefunction
   use tblite_xtb_calculator, only : xtb_calculator
   use tblite_xtb_gfn2, only : new_gfn2_calculator
   use tblite_xtb_gfn1, only : new_gfn1_calculator
   use tblite_xtb_ipea1, only : new_ipea1_calculator
   use tblite_xtb_singlepoint, only : xtb_singlepoint
   implicit none
   type(context_type) :: ctx
   type(structure_type) :: mol
   type(error_type), allocatable :: error
   type(xtb_calculator) :: calc
   type(wavefunction_type) :: wfn
   real(wp), parameter :: ktoau = 3.166808578545117e-06_wp
   real(wp) :: energy, sigma(3, 3), etemp
   real(wp), allocatable :: gradient(:, :)
   character(len=256) :: arg
   integer :: charge, i, method

   call get_command_argument(1, arg)
   call read_structure(mol, trim(arg), error)
   if (allocated(error)) error stop error%message
   call get_command_argument(2, arg); read(arg, *) charge
   call get_command_argument(3, arg); read(arg, *) etemp
   method = 2
   if (command_argument_count() > 3) then
      call get_command_argument(4, arg); read(arg, *) method
   end if
   mol%charge = real(charge, wp)
   mol%uhf = 0                        ! min(multiplicity-1, 0) is never positive (src/tblite.f90:111)
   allocate(gradient(3, mol%nat))
   select case (method)                 ! selectors of src/tblite.f90:34-40, calculators :123-129
   case (1); call new_gfn1_calculator(calc, mol)
   case (11); call new_ipea1_calculator(calc, mol)
   case default; call new_gfn2_calculator(calc, mol)
   end select
   call new_wavefunction(wfn, mol%nat, calc%bas%nsh, calc%bas%nao, 1, etemp * ktoau)
   call xtb_singlepoint(ctx, mol, calc, wfn, 1.0_wp, energy, gradient, sigma, 1)
   print '(a,es24.16,a)', '{"energy": ', energy, ','
   print '(a)', ' "gradient": ['
   do i = 1, mol%nat
      print '(a,es24.16,a,es24.16,a,es24.16,a)', '  [', gradient(1, i), ',', gradient(2, i), ',', gradient(3, i), merge('] ', '],', i == mol%nat)
   end do
   print '(a)', ' ], "qat": ['
   do i = 1, mol%nat
      print '(es24.16,a)', wfn%qat(i, 1), merge(' ', ',', i == mol%nat)
   end do
   print '(a)', ' ]}'
end program golden_driver
