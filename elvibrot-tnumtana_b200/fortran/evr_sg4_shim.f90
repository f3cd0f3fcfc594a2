!===============================================================================
! evr_sg4_shim.f90 -- ISO_C_BINDING shim that routes ElVibRot's SG4 operator
! action to the B200 library (include/evr_sg4.h).
!
! It provides a drop-in body for
!     SUBROUTINE sub_TabOpPsi_FOR_SGtype4(Psi,OpPsi,para_Op)
!         (Source_ElVibRot/sub_Operator/sub_OpPsi_SG4.f90:678-979)
! and, in MPI builds, for Action_MPI_S1 (sub_OpPsi_SG4_MPI.f90:454-571), with the
! same dummy arguments, so sub_OpPsi / sub_TabOpPsi (sub_Operator/sub_OpPsi.f90:
! 399,413,878) and every driver above them (Davidson, Chebyshev/SIL/RK
! propagators) stay untouched.
!
! First call with filled operator grids:
!              flatten para_Op%BasisnD (param_SGType2, WeightSG,
!              tab_basisPrimSG(L,k)%dnRGB/dnRBGwrho/dnRGG) and the cached
!              operator grids para_Op%OpGrid(:) -- whole-grid %Grid(:,:,:) or the
!              ragged %SRep%SmolyakRep(iG)%V of "direct=4" inputs -- into
!              contiguous arrays and create the device plan.
! While the grids are not filled yet (Save_MemGrid_done = F: the reference
!              computes them with Tnum during its first H|psi>, :2982-3006) the
!              call is forwarded to the reference's own routine, renamed
!              sub_TabOpPsi_FOR_SGtype4_ref (INTEGRATION.md).
! Every later call: pack Psi(:)%RvecB into page-locked buffers -> evr_sg4_apply ->
!              unpack into OpPsi(:)%RvecB.
! type_Op=10 : G, Jac and sqrt(rho/Jac) are evaluated ONCE per grid point with the
!              reference's get_OpGrid_type10_OF_ONEDP_FOR_SG4 (:2538) and cached
!              on the device (the reference recomputes them at every call).
! Several GPUs: EVR_SG4_NDEV=n in the environment (no MPI) makes the plan span n
!              GPUs of the node (evr_sg4_set_devices); with MPI scheme 1 every rank
!              drives GPU mod(MPI_id, #GPUs) with its own term range iGs_MPI and
!              its own slice of tab_iB_OF_SRep_TO_iB (evr_sg4_plan_create_ex).
! Errors     : non-zero status -> message + STOP (the reference's behaviour).
!
! Kinds      : the reference's default build (makefile INT = 4, Rkind = real64): its integer tables and Rkind arrays are
!              passed to the C-ABI as int32_t / double without a copy; an INT=8 build fails at the explicit interfaces.
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
  PUBLIC :: sub_TabOpPsi_FOR_SGtype4_GPU, sub_scaledTabOpPsi_FOR_SGtype4_GPU, evr_sg4_shim_release

  TYPE(C_PTR), SAVE :: plan = C_NULL_PTR     ! one cached plan (one para_Op: the Hamiltonian)
  integer,     SAVE :: plan_n_Op = -huge(1)
  logical,     SAVE :: devices_set = .FALSE.
  ! packed psi / H psi, page-locked once (evr_sg4_host_register) and reused by every call
  real(C_DOUBLE), allocatable, target, SAVE :: x(:), y(:)

  INTERFACE
    FUNCTION evr_sg4_plan_create_ex(plan, device, D, nb_SG, nb0, nb, LG, tab_l, WeightSG,       &
                                 tab_nq, tab_nb, tab_iB, tab_iB_first, tab_iB_len,              &
                                 nq_of, nb_of, B, BTw, D1, D2,                                  &
                                 iG_begin, iG_end) BIND(C, name='evr_sg4_plan_create_ex') RESULT(ierr)
      IMPORT :: C_PTR, C_INT, C_INT32_T, C_INT64_T, C_DOUBLE
      TYPE(C_PTR),            intent(inout) :: plan
      integer(C_INT),  VALUE                :: device, D, nb_SG, nb0, LG, iG_begin, iG_end
      integer(C_INT64_T), VALUE             :: nb, tab_iB_first, tab_iB_len
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
    FUNCTION evr_sg4_plan_set_op10(plan, n_act, act_mode, V, GG, Jac, sq)                       &
                                 BIND(C, name='evr_sg4_plan_set_op10') RESULT(ierr)
      IMPORT :: C_PTR, C_INT, C_INT32_T, C_DOUBLE
      TYPE(C_PTR),     VALUE                :: plan
      integer(C_INT),  VALUE                :: n_act
      integer(C_INT32_T),     intent(in)    :: act_mode(*)
      TYPE(C_PTR),     VALUE                :: V                 ! Grid(1:NQ,1:nb0,1:nb0) of the (0,0) term or C_NULL_PTR
      real(C_DOUBLE),         intent(in)    :: GG(*), Jac(*), sq(*)
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
    ! device-resident entry points (psi / Hpsi are device pointers, e.g. from the driver-algebra calls of
    ! include/evr_sg4_vec.h; cuda_stream = C_NULL_PTR is the default stream)
    FUNCTION evr_sg4_apply_device(plan, npsi, d_psi, d_Hpsi, cuda_stream)                        &
                                 BIND(C, name='evr_sg4_apply_device') RESULT(ierr)
      IMPORT :: C_PTR, C_INT
      TYPE(C_PTR),     VALUE                :: plan, d_psi, d_Hpsi, cuda_stream
      integer(C_INT),  VALUE                :: npsi
      integer(C_INT)                        :: ierr
    END FUNCTION
    FUNCTION evr_sg4_apply_device_scaled(plan, npsi, d_psi, d_Hpsi, E0, Esc, cuda_stream)        &
                                 BIND(C, name='evr_sg4_apply_device_scaled') RESULT(ierr)
      IMPORT :: C_PTR, C_INT, C_DOUBLE
      TYPE(C_PTR),     VALUE                :: plan, d_psi, d_Hpsi, cuda_stream
      integer(C_INT),  VALUE                :: npsi
      real(C_DOUBLE),  VALUE                :: E0, Esc
      integer(C_INT)                        :: ierr
    END FUNCTION
    FUNCTION evr_sg4_plan_destroy(plan) BIND(C, name='evr_sg4_plan_destroy') RESULT(ierr)
      IMPORT :: C_PTR, C_INT
      TYPE(C_PTR),            intent(inout) :: plan
      integer(C_INT)                        :: ierr
    END FUNCTION
    FUNCTION evr_sg4_set_devices(ndev) BIND(C, name='evr_sg4_set_devices') RESULT(ierr)
      IMPORT :: C_INT
      integer(C_INT),  VALUE                :: ndev
      integer(C_INT)                        :: ierr
    END FUNCTION
    FUNCTION evr_sg4_device_count() BIND(C, name='evr_sg4_device_count') RESULT(n)
      IMPORT :: C_INT
      integer(C_INT)                        :: n
    END FUNCTION
    FUNCTION evr_sg4_host_register(ptr, bytes) BIND(C, name='evr_sg4_host_register') RESULT(ierr)
      IMPORT :: C_PTR, C_INT, C_INT64_T
      TYPE(C_PTR),     VALUE                :: ptr
      integer(C_INT64_T), VALUE             :: bytes
      integer(C_INT)                        :: ierr
    END FUNCTION
    FUNCTION evr_sg4_host_unregister(ptr) BIND(C, name='evr_sg4_host_unregister') RESULT(ierr)
      IMPORT :: C_PTR, C_INT
      TYPE(C_PTR),     VALUE                :: ptr
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
  ! Are the operator grids of para_Op in memory?  The reference fills them during
  ! its first H|psi> (get_OpGrid_type1_OF_ONEDP_FOR_SG4, sub_OpPsi_SG4.f90:2982-3031)
  ! and then sets Save_MemGrid_done (:949-956).  Without Save_MemGrid (grids
  ! recomputed or read from file at every call) the GPU path cannot cache them.
  !-----------------------------------------------------------------------------
  FUNCTION shim_grids_ready(para_Op) RESULT(ready)
    USE mod_SetOp, ONLY : param_Op
    TYPE (param_Op), intent(in) :: para_Op
    logical :: ready
    integer :: iterm00
    ready = .FALSE.
    IF (.NOT. associated(para_Op%OpGrid)) RETURN
    iterm00 = para_Op%derive_term_TO_iterm(0,0)
    ready = para_Op%OpGrid(iterm00)%para_FileGrid%Save_MemGrid .AND.                          &
            para_Op%OpGrid(iterm00)%para_FileGrid%Save_MemGrid_done
  END FUNCTION shim_grids_ready

  !-----------------------------------------------------------------------------
  ! Build the device plan from para_Op (first call with filled grids).
  !-----------------------------------------------------------------------------
  SUBROUTINE shim_build_plan(para_Op)
    USE mod_system                                   ! Rkind, MPI_id, MPI_np, openmpi, MPI_scheme, iGs_MPI
    USE mod_basis_set_alloc, ONLY : basis, get_nq_FROM_basis, get_nb_FROM_basis
    USE mod_SetOp,           ONLY : param_Op
    USE mod_OpPsi_SG4,       ONLY : get_OpGrid_type10_OF_ONEDP_FOR_SG4      ! PUBLIC, sub_OpPsi_SG4.f90:51
    TYPE (param_Op), intent(inout), target :: para_Op

    TYPE (basis), pointer :: BasisnD
    integer :: D, LG, nb_SG, nb0, k, L, iterm, i, j, nq, nb, iG, iG_begin, iG_end, ierr, iq, ndev, device, n_act
    integer :: env_len, env_stat
    character(len=16) :: env_val
    integer(C_INT32_T), allocatable :: nq_of(:), nb_of(:), tab_l(:), term_mode(:), act_mode(:)
    integer(C_INT8_T),  allocatable :: gzero(:), gcte(:)
    real(C_DOUBLE),     allocatable :: B(:), BTw(:), D1(:), D2(:), Mat_cte(:)
    real(C_DOUBLE),     allocatable, target :: flat(:,:,:,:)        ! (nqq,nb0,nb0,slot): flattened %SRep grids
    real(C_DOUBLE),     allocatable, target :: GGall(:,:,:), Jacall(:), sqall(:), Vall(:,:,:)
    real (kind=Rkind),  allocatable :: V1(:,:,:), GG1(:,:,:), sq1(:), Jac1(:)
    TYPE(C_PTR),        allocatable :: grids(:)
    integer,            allocatable :: slot_of(:)
    integer(C_INT64_T) :: oB, oG, first, tlen
    integer :: nslot, off, nqq

    BasisnD => para_Op%BasisnD
    D     = BasisnD%nb_basis
    LG    = BasisnD%L_SparseGrid
    nb_SG = BasisnD%para_SGType2%nb_SG                 ! sub_module_param_SGType2.f90:54-102
    nb0   = BasisnD%para_SGType2%nb0
    nqq   = BasisnD%para_SGType2%tab_Sum_nq_OF_SRep(nb_SG)

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

    ! term range, device and mapping-table slice
    !  * no MPI: all terms; EVR_SG4_NDEV=n spreads the plan over n GPUs (device = -1: evr_sg4_set_devices decides)
    !  * MPI scheme 1: the rank's range iGs_MPI(1:2,MPI_id) (1-based inclusive; ini_iGs_MPI / auto_iGs_MPI,
    !    sub_module_basis_BtoG_GtoB_SG4_MPI.f90:639-669) on GPU mod(MPI_id, #GPUs), and the rank's slice of the mapping
    !    table: tab_iB_OF_SRep_TO_iB(bounds_MPI(1,id):bounds_MPI(2,id)) (Mapping_table_allocate_MPI, :62-77)
    iG_begin = 0 ; iG_end = nb_SG ; device = -1
    first = int(lbound(BasisnD%para_SGType2%tab_iB_OF_SRep_TO_iB,1) - 1, C_INT64_T)
    tlen  = int(size(BasisnD%para_SGType2%tab_iB_OF_SRep_TO_iB), C_INT64_T)
    IF (openmpi) THEN
      IF (MPI_scheme /= 1) STOP 'evr_sg4 shim: only MPI scheme 1 (replicated psi, terms split over the ranks) is supported'
      iG_begin = iGs_MPI(1,MPI_id) - 1
      iG_end   = iGs_MPI(2,MPI_id)
      ndev = evr_sg4_device_count()
      IF (ndev < 1) STOP 'evr_sg4 shim: no CUDA device (the library has no CPU fallback)'
      device = mod(int(MPI_id), ndev)
    ELSE IF (.NOT. devices_set) THEN
      CALL get_environment_variable('EVR_SG4_NDEV', env_val, env_len, env_stat)
      IF (env_stat == 0 .AND. env_len > 0) THEN
        read(env_val(1:env_len),*) ndev
        ierr = evr_sg4_set_devices(ndev)
        IF (ierr /= 0) CALL shim_stop('evr_sg4_set_devices')
      END IF
      devices_set = .TRUE.
    END IF

    ierr = evr_sg4_plan_create_ex(plan, device, D, nb_SG, nb0, int(BasisnD%nb, C_INT64_T), LG,       &
             tab_l, BasisnD%WeightSG, BasisnD%para_SGType2%tab_nq_OF_SRep,                           &
             BasisnD%para_SGType2%tab_nb_OF_SRep, BasisnD%para_SGType2%tab_iB_OF_SRep_TO_iB,         &
             first, tlen, nq_of, nb_of, B, BTw, D1, D2, iG_begin, iG_end)
    IF (ierr /= 0) CALL shim_stop('evr_sg4_plan_create_ex')

    IF (para_Op%type_Op == 10) THEN
      !-------------------------------------------------------------------------
      ! type_Op = 10 (sub_OpPsi_SG4.f90:1548-1650): the metric tensor, the Jacobian and sqrt(rho/Jac) of every grid point
      ! of this rank's terms, computed once with the reference's own routine and kept on the device.
      !-------------------------------------------------------------------------
      n_act = para_Op%mole%nb_act1
      allocate(act_mode(n_act))
      act_mode(:) = 0
      DO j = 1, n_act
        iq = para_Op%mole%liste_QactTOQdyn(j)
        DO k = 1, D
          IF (BasisnD%tab_basisPrimSG(0,k)%Tabder_Qdyn_TO_Qbasis(iq) /= 0) act_mode(j) = k
        END DO
      END DO
      allocate(GGall(nqq,n_act,n_act), Jacall(nqq), sqall(nqq), Vall(nqq,nb0,nb0))
      GGall = 0._C_DOUBLE ; Jacall = 1._C_DOUBLE ; sqall = 1._C_DOUBLE ; Vall = 0._C_DOUBLE
      DO iG = iG_begin+1, iG_end
        nq  = BasisnD%para_SGType2%tab_nq_OF_SRep(iG)
        off = BasisnD%para_SGType2%tab_Sum_nq_OF_SRep(iG) - nq
        CALL get_OpGrid_type10_OF_ONEDP_FOR_SG4(iG, BasisnD%para_SGType2%nDind_SmolyakRep%Tab_nDval(:,iG),  &
                                                para_Op, V1, GG1, sq1, Jac1)
        GGall(off+1:off+nq,:,:) = GG1(:,:,:)
        Jacall(off+1:off+nq)    = Jac1(:)
        sqall(off+1:off+nq)     = sq1(:)
        Vall(off+1:off+nq,:,:)  = V1(:,:,:)
        deallocate(V1, GG1, sq1, Jac1)
      END DO
      ierr = evr_sg4_plan_set_op10(plan, n_act, act_mode, C_LOC(Vall(1,1,1)), GGall, Jacall, sqall)
      IF (ierr /= 0) CALL shim_stop('evr_sg4_plan_set_op10')
      plan_n_Op = para_Op%n_Op
      RETURN
    END IF

    ! operator terms (type_Op = 0 or 1), sub_OpPsi_SG4.f90:1447-1546
    allocate(term_mode(2*para_Op%nb_Term), gzero(para_Op%nb_Term), gcte(para_Op%nb_Term))
    allocate(Mat_cte(nb0*nb0*para_Op%nb_Term), grids(para_Op%nb_Term), slot_of(para_Op%nb_Term))
    term_mode(:) = 0 ; Mat_cte(:) = 0._C_DOUBLE ; slot_of(:) = 0
    ! "direct=4" inputs keep the grids ragged in OpGrid(iterm)%SRep%SmolyakRep(iG)%V(nq*nb0*nb0) (:2966-2968, :3026-3031):
    ! they are flattened here into whole-grid arrays (only this rank's terms are filled and read)
    nslot = 0
    DO iterm = 1, para_Op%nb_Term
      IF (para_Op%OpGrid(iterm)%grid_zero .OR. para_Op%OpGrid(iterm)%grid_cte) CYCLE
      IF (.NOT. associated(para_Op%OpGrid(iterm)%Grid)) THEN
        IF (.NOT. allocated(para_Op%OpGrid(iterm)%SRep%SmolyakRep))                                     &
          STOP 'evr_sg4 shim: operator term without %Grid and without %SRep (Save_MemGrid=f ?)'
        nslot = nslot + 1 ; slot_of(iterm) = nslot
      END IF
    END DO
    IF (nslot > 0) THEN
      allocate(flat(nqq,nb0,nb0,nslot))
      DO iterm = 1, para_Op%nb_Term
        IF (slot_of(iterm) == 0) CYCLE
        DO iG = iG_begin+1, iG_end
          nq  = BasisnD%para_SGType2%tab_nq_OF_SRep(iG)
          off = BasisnD%para_SGType2%tab_Sum_nq_OF_SRep(iG) - nq
          flat(off+1:off+nq,:,:,slot_of(iterm)) =                                                       &
               reshape(para_Op%OpGrid(iterm)%SRep%SmolyakRep(iG)%V, shape=[nq,nb0,nb0])
        END DO
      END DO
    END IF
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
      ELSE IF (slot_of(iterm) > 0) THEN
        grids(iterm) = C_LOC(flat(1,1,1,slot_of(iterm)))
      ELSE
        grids(iterm) = C_LOC(para_Op%OpGrid(iterm)%Grid(1,1,1))      ! whole-grid storage Grid(nqq,nb0,nb0)
      END IF
    END DO
    ierr = evr_sg4_plan_set_op(plan, para_Op%type_Op, para_Op%nb_Term, term_mode, gzero, gcte, Mat_cte, grids)
    IF (ierr /= 0) CALL shim_stop('evr_sg4_plan_set_op')
    plan_n_Op = para_Op%n_Op
    ! (flat, B, ... are released here: the plan owns device copies)
  END SUBROUTINE shim_build_plan

  ! packed, page-locked host buffers of n doubles (grown on demand, registered once per allocation)
  SUBROUTINE shim_buffers(n)
    integer, intent(in) :: n
    integer :: ierr
    IF (allocated(x)) THEN
      IF (size(x) >= n) RETURN
      ierr = evr_sg4_host_unregister(C_LOC(x(1))) ; ierr = evr_sg4_host_unregister(C_LOC(y(1)))
      deallocate(x, y)
    END IF
    allocate(x(n), y(n))
    ierr = evr_sg4_host_register(C_LOC(x(1)), int(n,C_INT64_T)*8_C_INT64_T)
    IF (ierr == 0) ierr = evr_sg4_host_register(C_LOC(y(1)), int(n,C_INT64_T)*8_C_INT64_T)
    ! (a failed registration only costs speed: pageable copies are staged by the CUDA driver)
  END SUBROUTINE shim_buffers

  !-----------------------------------------------------------------------------
  ! Same interface as sub_TabOpPsi_FOR_SGtype4 (sub_OpPsi_SG4.f90:678).
  !-----------------------------------------------------------------------------
  SUBROUTINE sub_TabOpPsi_FOR_SGtype4_GPU(Psi,OpPsi,para_Op)
    USE mod_system
    USE mod_psi,    ONLY : param_psi, Set_symab_OF_psiBasisRep
    USE mod_SetOp,  ONLY : param_Op
    USE mod_SymAbelian, ONLY : Calc_symab1_EOR_symab2
    USE mod_OpPsi_SG4,  ONLY : sub_TabOpPsi_FOR_SGtype4_ref      ! the reference's routine, renamed (INTEGRATION.md)
    USE mod_MPI_aux
    TYPE (param_psi), intent(in)      :: Psi(:)
    TYPE (param_psi), intent(inout)   :: OpPsi(:)
    TYPE (param_Op),  intent(inout)   :: para_Op

    integer :: itab, n, npsi, ierr, OpPsi_symab

    IF (size(Psi) == 0) STOP ' ERROR in sub_TabOpPsi_FOR_SGtype4: size(Psi) = 0'      ! :738-743
    IF (Psi(1)%cplx)    STOP ' ERROR in sub_TabOpPsi_FOR_SGtype4: Psi(1) is complex'  ! :744-749

    ! The operator grids must exist: the reference fills them during its first H|psi>
    ! (get_OpGrid_type1_OF_ONEDP_FOR_SG4, :2982-3031).  Until then the call goes to the Fortran path, which
    ! computes the grids with Tnum, stores them and sets Save_MemGrid_done (:949-956); type_Op = 10 has no such
    ! first pass (its G/Jac/rho are never stored by the reference), the plan builder evaluates them itself.
    IF (para_Op%type_Op /= 10 .AND. .NOT. shim_grids_ready(para_Op)) THEN
      CALL sub_TabOpPsi_FOR_SGtype4_ref(Psi,OpPsi,para_Op)
      RETURN
    END IF
    IF (.NOT. C_ASSOCIATED(plan) .OR. plan_n_Op /= para_Op%n_Op) THEN
      IF (C_ASSOCIATED(plan)) ierr = evr_sg4_plan_destroy(plan)
      CALL shim_build_plan(para_Op)
    END IF

    npsi = size(Psi)
    n    = size(Psi(1)%RvecB)
    CALL shim_buffers(n*npsi)
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
  END SUBROUTINE sub_TabOpPsi_FOR_SGtype4_GPU

  !-----------------------------------------------------------------------------
  ! H|psi> followed by sub_scaledOpPsi (sub_OpPsi.f90:2823-2866) for the Chebyshev / SIL recursions
  ! (sub_module_propa_march.f90:4294-4345): OpPsi <- (H Psi - E0 Psi)/Esc.  Host vectors; the device-resident variant is
  ! evr_sg4_apply_device_scaled (interface above) for drivers that keep their vectors on the GPU.
  !-----------------------------------------------------------------------------
  SUBROUTINE sub_scaledTabOpPsi_FOR_SGtype4_GPU(Psi,OpPsi,para_Op,E0,Esc)
    USE mod_system
    USE mod_psi,    ONLY : param_psi
    USE mod_SetOp,  ONLY : param_Op
    TYPE (param_psi), intent(in)      :: Psi(:)
    TYPE (param_psi), intent(inout)   :: OpPsi(:)
    TYPE (param_Op),  intent(inout)   :: para_Op
    real (kind=Rkind), intent(in)     :: E0, Esc
    integer :: itab
    CALL sub_TabOpPsi_FOR_SGtype4_GPU(Psi,OpPsi,para_Op)
    DO itab = 1, size(Psi)
      OpPsi(itab)%RvecB(:) = (OpPsi(itab)%RvecB(:) - E0*Psi(itab)%RvecB(:)) / Esc
    END DO
  END SUBROUTINE sub_scaledTabOpPsi_FOR_SGtype4_GPU

  SUBROUTINE evr_sg4_shim_release()
    integer :: ierr
    IF (C_ASSOCIATED(plan)) ierr = evr_sg4_plan_destroy(plan)
    plan = C_NULL_PTR
    IF (allocated(x)) THEN
      ierr = evr_sg4_host_unregister(C_LOC(x(1))) ; ierr = evr_sg4_host_unregister(C_LOC(y(1)))
      deallocate(x, y)
    END IF
  END SUBROUTINE evr_sg4_shim_release

END MODULE mod_evr_sg4_shim
