!===============================================================================
! evr_sg4_shim.f90 -- ISO_C_BINDING shim that routes ElVibRot's SG4 operator
! action to the B200 library (include/evr_sg4.h).
!
! It provides a drop-in body for
!     SUBROUTINE sub_TabOpPsi_FOR_SGtype4(Psi,OpPsi,para_Op)
!         (Source_ElVibRot/sub_Operator/sub_OpPsi_SG4.f90:678-979)
! with the same dummy arguments, so sub_OpPsi / sub_TabOpPsi
! (sub_Operator/sub_OpPsi.f90:399,413,878) and every driver above them
! (Davidson, Chebyshev/SIL/RK propagators) stay untouched.
!
! First call : flatten para_Op%BasisnD (param_SGType2, WeightSG,
!              tab_basisPrimSG(L,k)%dnRGB/dnRBGwrho/dnRGG) and the cached
!              operator grids para_Op%OpGrid(:) into contiguous arrays and
!              create the device plan.
! Every call : pack Psi(:)%RvecB -> evr_sg4_apply -> unpack into OpPsi(:)%RvecB.
! Errors     : non-zero status -> message + STOP (the reference's behaviour).
!
! NOT COMPILED IN THIS REPOSITORY'S CI: the build image has no Fortran
! compiler (SURVEY.md F2).  The flattening below follows the reference's type
! definitions (file:line cited inline); INTEGRATION.md lists the makefile lines
! to add.
!===============================================================================
MODULE mod_evr_sg4_shim
  USE, INTRINSIC :: ISO_C_BINDING
  IMPLICIT NONE
  PRIVATE
  PUBLIC :: sub_TabOpPsi_FOR_SGtype4_GPU, evr_sg4_shim_release

  TYPE(C_PTR), SAVE :: plan = C_NULL_PTR     ! one cached plan (one para_Op: the Hamiltonian)
  integer,     SAVE :: plan_n_Op = -huge(1)

  INTERFACE
    FUNCTION evr_sg4_plan_create(plan, device, D, nb_SG, nb0, nb, LG, tab_l, WeightSG,          &
                                 tab_nq, tab_nb, tab_iB, nq_of, nb_of, B, BTw, D1, D2,          &
                                 iG_begin, iG_end) BIND(C, name='evr_sg4_plan_create') RESULT(ierr)
      IMPORT :: C_PTR, C_INT, C_INT32_T, C_INT64_T, C_DOUBLE
      TYPE(C_PTR),            intent(inout) :: plan
      integer(C_INT),  VALUE                :: device, D, nb_SG, nb0, LG, iG_begin, iG_end
      integer(C_INT64_T), VALUE             :: nb
      integer(C_INT32_T),     intent(in)    :: tab_l(*), tab_nq(*), tab_nb(*), tab_iB(*), nq_of(*), nb_of(*)
      real(C_DOUBLE),         intent(in)    :: WeightSG(*), B(*), BTw(*), D1(*), D2(*)
      integer(C_INT)                        :: ierr
    END FUNCTION
    FUNCTION evr_sg4_plan_set_op(plan, type_Op, nb_Term, term_mode, grid_zero, grid_cte,        &
                                 Mat_cte, grids) BIND(C, name='evr_sg4_plan_set_op') RESULT(ierr)
      IMPORT :: C_PTR, C_INT, C_INT32_T, C_INT8_T, C_DOUBLE
      TYPE(C_PTR),     VALUE                :: plan
      integer(C_INT),  VALUE                :: type_Op, nb_Term
      integer(C_INT32_T),     intent(in)    :: term_mode(*)
      integer(C_INT8_T),      intent(in)    :: grid_zero(*), grid_cte(*)
      real(C_DOUBLE),         intent(in)    :: Mat_cte(*)
      TYPE(C_PTR),            intent(in)    :: grids(*)
      integer(C_INT)                        :: ierr
    END FUNCTION
    FUNCTION evr_sg4_apply(plan, npsi, psi, Hpsi) BIND(C, name='evr_sg4_apply') RESULT(ierr)
      IMPORT :: C_PTR, C_INT, C_DOUBLE
      TYPE(C_PTR),     VALUE                :: plan
      integer(C_INT),  VALUE                :: npsi
      real(C_DOUBLE),         intent(in)    :: psi(*)
      real(C_DOUBLE),         intent(inout) :: Hpsi(*)
      integer(C_INT)                        :: ierr
    END FUNCTION
    FUNCTION evr_sg4_plan_destroy(plan) BIND(C, name='evr_sg4_plan_destroy') RESULT(ierr)
      IMPORT :: C_PTR, C_INT
      TYPE(C_PTR),            intent(inout) :: plan
      integer(C_INT)                        :: ierr
    END FUNCTION
    FUNCTION evr_sg4_ini_iGs(nb_SG, np, rank, iG_begin, iG_end) BIND(C, name='evr_sg4_ini_iGs') RESULT(ierr)
      IMPORT :: C_INT
      integer(C_INT),  VALUE                :: nb_SG, np, rank
      integer(C_INT),         intent(out)   :: iG_begin, iG_end
      integer(C_INT)                        :: ierr
    END FUNCTION
    FUNCTION evr_sg4_last_error() BIND(C, name='evr_sg4_last_error') RESULT(msg)
      IMPORT :: C_PTR
      TYPE(C_PTR) :: msg
    END FUNCTION
  END INTERFACE

CONTAINS

  SUBROUTINE shim_stop(where)
    character(len=*), intent(in) :: where
    character(kind=C_CHAR), pointer :: cmsg(:)
    integer :: i
    write(6,*) ' ERROR in ', where, ' (evr_sg4 GPU library):'
    CALL C_F_POINTER(evr_sg4_last_error(), cmsg, [512])
    DO i = 1, 512
      IF (cmsg(i) == C_NULL_CHAR) EXIT
      write(6,'(a)',ADVANCE='no') cmsg(i)
    END DO
    write(6,*)
    STOP 'ERROR in the evr_sg4 GPU library'
  END SUBROUTINE shim_stop

  !-----------------------------------------------------------------------------
  ! Build the device plan from para_Op (first call only).
  !-----------------------------------------------------------------------------
  SUBROUTINE shim_build_plan(para_Op)
    USE mod_system                                   ! Rkind, MPI_id, MPI_np, openmpi
    USE mod_basis_set_alloc, ONLY : basis, get_nq_FROM_basis, get_nb_FROM_basis
    USE mod_SetOp,           ONLY : param_Op
    TYPE (param_Op), intent(inout), target :: para_Op

    TYPE (basis), pointer :: BasisnD
    integer :: D, LG, nb_SG, nb0, k, L, iterm, i, j, nq, nb, iG_begin, iG_end, ierr, iq
    integer(C_INT32_T), allocatable :: nq_of(:), nb_of(:), tab_l(:), term_mode(:)
    integer(C_INT8_T),  allocatable :: gzero(:), gcte(:)
    real(C_DOUBLE),     allocatable :: B(:), BTw(:), D1(:), D2(:), Mat_cte(:)
    TYPE(C_PTR),        allocatable :: grids(:)
    integer(C_INT64_T) :: oB, oG

    BasisnD => para_Op%BasisnD
    D     = BasisnD%nb_basis
    LG    = BasisnD%L_SparseGrid
    nb_SG = BasisnD%para_SGType2%nb_SG                 ! sub_module_param_SGType2.f90:54-102
    nb0   = BasisnD%para_SGType2%nb0

    ! level sizes and concatenated 1-D tables: mode k outer, level L inner (include/evr_sg4.h)
    allocate(nq_of(D*(LG+1)), nb_of(D*(LG+1)))
    oB = 0 ; oG = 0
    DO k = 1, D
    DO L = 0, LG
      nq = get_nq_FROM_basis(BasisnD%tab_basisPrimSG(L,k))
      nb = get_nb_FROM_basis(BasisnD%tab_basisPrimSG(L,k))
      IF (BasisnD%tab_basisPrimSG(L,k)%ndim /= 1) STOP 'evr_sg4 shim: only 1-D primitive bases are supported'
      nq_of((k-1)*(LG+1)+L+1) = nq ; nb_of((k-1)*(LG+1)+L+1) = nb
      oB = oB + nq*nb ; oG = oG + nq*nq
    END DO
    END DO
    allocate(B(oB), BTw(oB), D1(oG), D2(oG))
    oB = 0 ; oG = 0
    DO k = 1, D
    DO L = 0, LG
      nq = nq_of((k-1)*(LG+1)+L+1) ; nb = nb_of((k-1)*(LG+1)+L+1)
      B  (oB+1:oB+nq*nb) = reshape(BasisnD%tab_basisPrimSG(L,k)%dnRGB%d0,          [nq*nb])  ! (nq,nb)
      BTw(oB+1:oB+nq*nb) = reshape(BasisnD%tab_basisPrimSG(L,k)%dnRBGwrho%d0,      [nq*nb])  ! (nb,nq)
      D1 (oG+1:oG+nq*nq) = reshape(BasisnD%tab_basisPrimSG(L,k)%dnRGG%d1(:,:,1),   [nq*nq])
      D2 (oG+1:oG+nq*nq) = reshape(BasisnD%tab_basisPrimSG(L,k)%dnRGG%d2(:,:,1,1), [nq*nq])
      oB = oB + nq*nb ; oG = oG + nq*nq
    END DO
    END DO

    ! term table (packed nDind_SmolyakRep is allocated in every shipped SG4 input, sub_OpPsi_SG4.f90:812)
    allocate(tab_l(D*nb_SG))
    tab_l(:) = reshape(BasisnD%para_SGType2%nDind_SmolyakRep%Tab_nDval(:,1:nb_SG), [D*nb_SG])

    ! term range of this MPI rank (scheme 1); all terms without MPI
    iG_begin = 0 ; iG_end = nb_SG
    IF (openmpi) THEN
      ierr = evr_sg4_ini_iGs(nb_SG, int(MPI_np), int(MPI_id), iG_begin, iG_end)
      IF (ierr /= 0) CALL shim_stop('evr_sg4_ini_iGs')
    END IF

    ierr = evr_sg4_plan_create(plan, -1, D, nb_SG, nb0, int(BasisnD%nb, C_INT64_T), LG,              &
             tab_l, BasisnD%WeightSG, BasisnD%para_SGType2%tab_nq_OF_SRep,                           &
             BasisnD%para_SGType2%tab_nb_OF_SRep, BasisnD%para_SGType2%tab_iB_OF_SRep_TO_iB,         &
             nq_of, nb_of, B, BTw, D1, D2, iG_begin, iG_end)
    IF (ierr /= 0) CALL shim_stop('evr_sg4_plan_create')

    ! operator terms (type_Op = 0 or 1), sub_OpPsi_SG4.f90:1447-1546
    allocate(term_mode(2*para_Op%nb_Term), gzero(para_Op%nb_Term), gcte(para_Op%nb_Term))
    allocate(Mat_cte(nb0*nb0*para_Op%nb_Term), grids(para_Op%nb_Term))
    term_mode(:) = 0 ; Mat_cte(:) = 0._C_DOUBLE
    DO iterm = 1, para_Op%nb_Term
      ! which SG4 mode owns each derivative index: Tabder_Qdyn_TO_Qbasis of the level-0 primitive
      DO i = 1, 2
        iq = para_Op%derive_termQdyn(i,iterm)
        IF (iq > 0) THEN
          DO k = 1, D
            IF (BasisnD%tab_basisPrimSG(0,k)%Tabder_Qdyn_TO_Qbasis(iq) /= 0) term_mode(2*(iterm-1)+i) = k
          END DO
        END IF
      END DO
      gzero(iterm) = merge(1_C_INT8_T, 0_C_INT8_T, para_Op%OpGrid(iterm)%grid_zero)
      gcte(iterm)  = merge(1_C_INT8_T, 0_C_INT8_T, para_Op%OpGrid(iterm)%grid_cte)
      grids(iterm) = C_NULL_PTR
      IF (para_Op%OpGrid(iterm)%grid_zero) CYCLE
      IF (para_Op%OpGrid(iterm)%grid_cte) THEN
        DO j = 1, nb0
        DO i = 1, nb0
          Mat_cte((iterm-1)*nb0*nb0 + i + nb0*(j-1)) = para_Op%OpGrid(iterm)%Mat_cte(i,j)
        END DO
        END DO
      ELSE
        ! whole-grid storage OpGrid(iterm)%Grid(nqq,nb0,nb0) (Type_FileGrid=4 "direct=4" inputs use the
        ! ragged %SRep instead: flatten it first with tabR2bis_TO_SmolyakRep1's inverse, sub_OpPsi_SG4.f90:2943-2971)
        IF (.NOT. associated(para_Op%OpGrid(iterm)%Grid)) STOP 'evr_sg4 shim: flatten OpGrid%SRep first'
        grids(iterm) = C_LOC(para_Op%OpGrid(iterm)%Grid(1,1,1))
      END IF
    END DO
    ierr = evr_sg4_plan_set_op(plan, para_Op%type_Op, para_Op%nb_Term, term_mode, gzero, gcte, Mat_cte, grids)
    IF (ierr /= 0) CALL shim_stop('evr_sg4_plan_set_op')
    plan_n_Op = para_Op%n_Op
  END SUBROUTINE shim_build_plan

  !-----------------------------------------------------------------------------
  ! Same interface as sub_TabOpPsi_FOR_SGtype4 (sub_OpPsi_SG4.f90:678).
  !-----------------------------------------------------------------------------
  SUBROUTINE sub_TabOpPsi_FOR_SGtype4_GPU(Psi,OpPsi,para_Op)
    USE mod_system
    USE mod_psi,    ONLY : param_psi, Set_symab_OF_psiBasisRep
    USE mod_SetOp,  ONLY : param_Op
    USE mod_SymAbelian, ONLY : Calc_symab1_EOR_symab2
    USE mod_MPI_aux
    TYPE (param_psi), intent(in)      :: Psi(:)
    TYPE (param_psi), intent(inout)   :: OpPsi(:)
    TYPE (param_Op),  intent(inout)   :: para_Op

    real(C_DOUBLE), allocatable :: x(:), y(:)
    integer :: itab, n, npsi, ierr, iterm00, OpPsi_symab

    IF (size(Psi) == 0) STOP ' ERROR in sub_TabOpPsi_FOR_SGtype4: size(Psi) = 0'      ! :738-743
    IF (Psi(1)%cplx)    STOP ' ERROR in sub_TabOpPsi_FOR_SGtype4: Psi(1) is complex'  ! :744-749

    ! the operator grids must exist: the reference fills them during its first H|psi>
    ! (get_OpGrid_type1_OF_ONEDP_FOR_SG4, :2982-3006).  Call the Fortran path once, then switch.
    IF (.NOT. C_ASSOCIATED(plan) .OR. plan_n_Op /= para_Op%n_Op) THEN
      IF (C_ASSOCIATED(plan)) ierr = evr_sg4_plan_destroy(plan)
      CALL shim_build_plan(para_Op)
    END IF

    npsi = size(Psi)
    n    = size(Psi(1)%RvecB)
    allocate(x(n*npsi), y(n*npsi))
    DO itab = 1, npsi
      x((itab-1)*n+1:itab*n) = Psi(itab)%RvecB(:)
    END DO
    ierr = evr_sg4_apply(plan, npsi, x, y)
    IF (ierr /= 0) CALL shim_stop('evr_sg4_apply')
    IF (openmpi .AND. MPI_np > 1) CALL MPI_Reduce_sum_Bcast(y, n*npsi)               ! scheme 1, sub_OpPsi_SG4_MPI.f90:553-557
    DO itab = 1, npsi
      OpPsi(itab) = Psi(itab)                                                         ! allocation, as :762-766
      OpPsi(itab)%RvecB(:) = y((itab-1)*n+1:itab*n)
      OpPsi_symab = Calc_symab1_EOR_symab2(para_Op%symab,Psi(itab)%symab)             ! :958-966
      CALL Set_symab_OF_psiBasisRep(OpPsi(itab),OpPsi_symab)
    END DO
    iterm00 = para_Op%derive_term_TO_iterm(0,0)                                        ! :949-956
    IF (associated(para_Op%OpGrid)) para_Op%OpGrid(iterm00)%para_FileGrid%Save_MemGrid_done = .TRUE.
    deallocate(x, y)
  END SUBROUTINE sub_TabOpPsi_FOR_SGtype4_GPU

  SUBROUTINE evr_sg4_shim_release()
    integer :: ierr
    IF (C_ASSOCIATED(plan)) ierr = evr_sg4_plan_destroy(plan)
    plan = C_NULL_PTR
  END SUBROUTINE evr_sg4_shim_release

END MODULE mod_evr_sg4_shim
