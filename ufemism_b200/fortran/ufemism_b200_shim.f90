MODULE ufemism_b200_shim
  ! ISO_C_BINDING shim between the UFEMISM Fortran host and libufemism_b200.so (include/ufemism_b200.h).
  !
  ! Drop-in: a maintainer adds this file to src/, lists it in src/Makefile_include_local.txt before
  ! ice_dynamics_module, links with -lufemism_b200 -lcudart, and replaces four call sites in
  ! run_model (src/UFEMISM_main_model.f90:90,115,124,132) by the *_b200 wrappers below; the names,
  ! argument lists and error behaviour of the four entry points are those of the reference:
  !
  !   calculate_ice_thickness_change( mesh, ice, SMB, BMB, dt, mask_noice)   ice_dynamics_module.f90:31
  !   update_general_ice_model_data(  mesh, ice, time)                       general_ice_model_data_module.f90:23
  !   solve_SIA(                      mesh, ice)                             ice_dynamics_module.f90:240
  !   solve_SSA(                      mesh, ice)                             ice_dynamics_module.f90:408
  !
  ! Process model: the host stays SPMD (all ranks call every routine).  Only par%master talks to the
  ! GPU; every wrapper ends in CALL sync, and because all arrays are MPI shared-memory windows
  ! (src/parallel_module.f90:136-160) what the master downloads is visible to every rank.
  !
  ! NOTE: this image has no Fortran compiler, so this file is delivered as source; the same C entry
  ! points are exercised from Python/ctypes with Fortran-ordered arrays (ufemism_b200/capi.py) and the
  ! struct layouts below are checked against the header by tests/test_abi.py.

  USE, INTRINSIC :: ISO_C_BINDING
  USE mpi
  USE configuration_module,          ONLY: dp, C
  USE parallel_module,               ONLY: par, sync, ierr, cerr
  USE data_types_module,             ONLY: type_mesh, type_ice_model, type_SMB_model, type_BMB_model, type_climate_model, type_model_region
  USE data_types_netcdf_module,      ONLY: type_netcdf_restart
  USE mesh_memory_module,            ONLY: allocate_mesh_primary

  IMPLICIT NONE

  INTEGER, PARAMETER :: UFM_MAX_NZ = 32

  ! enum ufm_benchmark
  INTEGER(C_INT), PARAMETER :: UFM_BM_NONE = 0, UFM_BM_EISMINT_1 = 1, UFM_BM_HALFAR = 7, UFM_BM_BUELER = 8, &
                               UFM_BM_MISMIP_MOD = 9, UFM_BM_MESH_GENERATION_TEST = 10, UFM_BM_SSA_ICESTREAM = 11

  ! enum ufm_field (only the ids the shim moves; full list in the header)
  INTEGER(C_INT), PARAMETER :: UFM_F_HI = 0, UFM_F_HB = 1, UFM_F_SL = 2, UFM_F_DHB_DT = 3, UFM_F_SMB_YEAR = 4, UFM_F_BMB = 5, &
                               UFM_F_MASK_NOICE = 6, UFM_F_HS = 7, UFM_F_DHI_DT = 8, UFM_F_DHS_DT = 9, UFM_F_HI_PREV = 10, &
                               UFM_F_DHI_DX = 11, UFM_F_DHI_DY = 12, UFM_F_DHS_DX = 13, UFM_F_DHS_DY = 14, &
                               UFM_F_U_SIA = 18, UFM_F_V_SIA = 19, UFM_F_D_SIA = 20, UFM_F_U_SSA = 21, UFM_F_V_SSA = 22, &
                               UFM_F_MASK_LAND = 23, UFM_F_MASK_OCEAN = 24, UFM_F_MASK_LAKE = 25, UFM_F_MASK_ICE = 26, &
                               UFM_F_MASK_SHEET = 27, UFM_F_MASK_SHELF = 28, UFM_F_MASK_COAST = 29, UFM_F_MASK_MARGIN = 30, &
                               UFM_F_MASK_GL = 31, UFM_F_MASK_CF = 32, UFM_F_MASK = 33, &
                               UFM_F_U_3D = 94, UFM_F_V_3D = 95, UFM_F_TI = 96, UFM_F_W_3D = 97, UFM_F_GHF = 98, &
                               UFM_F_T2M = 99, UFM_F_FRICTIONAL_HEATING = 100

  TYPE, BIND(C) :: ufm_params
    INTEGER(C_INT)  :: nZ
    REAL(C_DOUBLE)  :: zeta( UFM_MAX_NZ)
    REAL(C_DOUBLE)  :: m_enh_sia, m_enh_ssa
    INTEGER(C_INT)  :: use_analytical_GL_flux
    REAL(C_DOUBLE)  :: SSA_RN_tol
    INTEGER(C_INT)  :: SSA_max_outer_loops
    REAL(C_DOUBLE)  :: SSA_max_residual_UV, SSA_SOR_omega
    INTEGER(C_INT)  :: SSA_max_inner_loops
    REAL(C_DOUBLE)  :: dt_max
    INTEGER(C_INT)  :: benchmark
    INTEGER(C_INT)  :: exact_xy
    REAL(C_DOUBLE)  :: dt_thermo
  END TYPE ufm_params

  TYPE, BIND(C) :: ufm_mesh_desc
    INTEGER(C_INT)  :: nV, nAc, nC_mem
    INTEGER(C_INT)  :: ldV, ldAc, ldAaAc
    TYPE(C_PTR)     :: V, A, nC, C, Cw, edge_index, Nx, Ny
    TYPE(C_PTR)     :: Aci, iAci, edge_index_Ac, Nx_Ac, Ny_Ac, No_Ac, Np_Ac
    TYPE(C_PTR)     :: nCAaAc, CAaAc, Nx_AaAc, Ny_AaAc, Nxx_AaAc, Nxy_AaAc, Nyy_AaAc
    TYPE(C_PTR)     :: colour_vi, colour_nV
    ! read only by update_ice_temperature_b200 (C_NULL_PTR: thermodynamics stays unavailable on the device)
    INTEGER(C_INT)  :: nTri, ldTri
    TYPE(C_PTR)     :: Tri, niTri, iTri, R, NxTri, NyTri
  END TYPE ufm_mesh_desc

  ! primary mesh data only (row N3): the library derives A, Cw, the Ac / AaAc meshes, the colouring and all neighbour functions
  TYPE, BIND(C) :: ufm_mesh_primary
    INTEGER(C_INT)  :: nV, nTri, nC_mem
    INTEGER(C_INT)  :: ldV, ldTri
    REAL(C_DOUBLE)  :: xmin, xmax, ymin, ymax
    TYPE(C_PTR)     :: V, nC, C, niTri, iTri, edge_index, Tri
    INTEGER(C_INT)  :: thermo
  END TYPE ufm_mesh_primary

  ! restart / help_fields files (row N4)
  TYPE, BIND(C) :: ufm_nc_mesh
    INTEGER(C_INT)  :: nV, nTri, nC_mem, nAc, nV_transect, nVAaAc, nTriAaAc
    TYPE(C_PTR)     :: V, Tri, nC, C, niTri, iTri, edge_index, Tricc, TriC, Tri_edge_index, VAc, Aci, iAci, VAaAc, TriAaAc, A, R
    TYPE(C_PTR)     :: vi_transect, w_transect
  END TYPE ufm_nc_mesh

  TYPE, BIND(C) :: ufm_thermo_stats
    INTEGER(C_INT)  :: n_unstable, rc
  END TYPE ufm_thermo_stats

  TYPE, BIND(C) :: ufm_ssa_stats
    INTEGER(C_INT)  :: n_outer, n_inner_total, n_inner_last, did_reset, rc
    REAL(C_DOUBLE)  :: last_max_residual, last_RN
  END TYPE ufm_ssa_stats

  INTERFACE
    FUNCTION ufm_create( device, params, handle) BIND(C, NAME='ufm_create') RESULT( rc)
      IMPORT :: C_INT, C_PTR, ufm_params
      INTEGER(C_INT), VALUE       :: device
      TYPE(ufm_params), INTENT(IN):: params
      TYPE(C_PTR), INTENT(OUT)    :: handle
      INTEGER(C_INT)              :: rc
    END FUNCTION ufm_create
    FUNCTION ufm_destroy( handle) BIND(C, NAME='ufm_destroy') RESULT( rc)
      IMPORT :: C_INT, C_PTR
      TYPE(C_PTR), VALUE          :: handle
      INTEGER(C_INT)              :: rc
    END FUNCTION ufm_destroy
    FUNCTION ufm_mesh_upload( handle, mesh) BIND(C, NAME='ufm_mesh_upload') RESULT( rc)
      IMPORT :: C_INT, C_PTR, ufm_mesh_desc
      TYPE(C_PTR), VALUE          :: handle
      TYPE(ufm_mesh_desc), INTENT(IN) :: mesh
      INTEGER(C_INT)              :: rc
    END FUNCTION ufm_mesh_upload
    FUNCTION ufm_mesh_free( handle) BIND(C, NAME='ufm_mesh_free') RESULT( rc)
      IMPORT :: C_INT, C_PTR
      TYPE(C_PTR), VALUE          :: handle
      INTEGER(C_INT)              :: rc
    END FUNCTION ufm_mesh_free
    FUNCTION ufm_state_upload( handle, field, host) BIND(C, NAME='ufm_state_upload') RESULT( rc)
      IMPORT :: C_INT, C_PTR
      TYPE(C_PTR), VALUE          :: handle
      INTEGER(C_INT), VALUE       :: field
      TYPE(C_PTR), VALUE          :: host
      INTEGER(C_INT)              :: rc
    END FUNCTION ufm_state_upload
    FUNCTION ufm_state_download( handle, field, host) BIND(C, NAME='ufm_state_download') RESULT( rc)
      IMPORT :: C_INT, C_PTR
      TYPE(C_PTR), VALUE          :: handle
      INTEGER(C_INT), VALUE       :: field
      TYPE(C_PTR), VALUE          :: host
      INTEGER(C_INT)              :: rc
    END FUNCTION ufm_state_download
    FUNCTION ufm_thickness_update( handle, dt) BIND(C, NAME='ufm_thickness_update') RESULT( rc)
      IMPORT :: C_INT, C_PTR, C_DOUBLE
      TYPE(C_PTR), VALUE          :: handle
      REAL(C_DOUBLE), VALUE       :: dt
      INTEGER(C_INT)              :: rc
    END FUNCTION ufm_thickness_update
    FUNCTION ufm_update_general( handle, time) BIND(C, NAME='ufm_update_general') RESULT( rc)
      IMPORT :: C_INT, C_PTR, C_DOUBLE
      TYPE(C_PTR), VALUE          :: handle
      REAL(C_DOUBLE), VALUE       :: time
      INTEGER(C_INT)              :: rc
    END FUNCTION ufm_update_general
    FUNCTION ufm_solve_SIA( handle) BIND(C, NAME='ufm_solve_SIA') RESULT( rc)
      IMPORT :: C_INT, C_PTR
      TYPE(C_PTR), VALUE          :: handle
      INTEGER(C_INT)              :: rc
    END FUNCTION ufm_solve_SIA
    FUNCTION ufm_solve_SSA( handle, stats) BIND(C, NAME='ufm_solve_SSA') RESULT( rc)
      IMPORT :: C_INT, C_PTR, ufm_ssa_stats
      TYPE(C_PTR), VALUE          :: handle
      TYPE(ufm_ssa_stats), INTENT(OUT) :: stats
      INTEGER(C_INT)              :: rc
    END FUNCTION ufm_solve_SSA
    FUNCTION ufm_cfl( handle, out3) BIND(C, NAME='ufm_cfl') RESULT( rc)
      IMPORT :: C_INT, C_PTR, C_DOUBLE
      TYPE(C_PTR), VALUE          :: handle
      REAL(C_DOUBLE), INTENT(OUT) :: out3( 3)
      INTEGER(C_INT)              :: rc
    END FUNCTION ufm_cfl
    FUNCTION ufm_solve_SIA_3D( handle) BIND(C, NAME='ufm_solve_SIA_3D') RESULT( rc)
      IMPORT :: C_INT, C_PTR
      TYPE(C_PTR), VALUE          :: handle
      INTEGER(C_INT)              :: rc
    END FUNCTION ufm_solve_SIA_3D
    FUNCTION ufm_update_ice_temperature( handle, stats) BIND(C, NAME='ufm_update_ice_temperature') RESULT( rc)
      IMPORT :: C_PTR, C_INT, ufm_thermo_stats
      TYPE(C_PTR), VALUE        :: handle
      TYPE(ufm_thermo_stats)    :: stats
      INTEGER(C_INT)            :: rc
    END FUNCTION ufm_update_ice_temperature
    FUNCTION ufm_host_register( handle, host, bytes) BIND(C, NAME='ufm_host_register') RESULT( rc)
      IMPORT :: C_INT, C_PTR, C_LONG_LONG
      TYPE(C_PTR), VALUE          :: handle
      TYPE(C_PTR), VALUE          :: host
      INTEGER(C_LONG_LONG), VALUE :: bytes
      INTEGER(C_INT)              :: rc
    END FUNCTION ufm_host_register
    FUNCTION ufm_partition_set( handle, rank, nranks) BIND(C, NAME='ufm_partition_set') RESULT( rc)
      IMPORT :: C_INT, C_PTR
      TYPE(C_PTR), VALUE          :: handle
      INTEGER(C_INT), VALUE       :: rank, nranks
      INTEGER(C_INT)              :: rc
    END FUNCTION ufm_partition_set
    FUNCTION ufm_comm_export( handle, blob) BIND(C, NAME='ufm_comm_export') RESULT( rc)
      IMPORT :: C_INT, C_PTR, C_CHAR
      TYPE(C_PTR), VALUE          :: handle
      CHARACTER(KIND=C_CHAR), INTENT(OUT) :: blob( 256)
      INTEGER(C_INT)              :: rc
    END FUNCTION ufm_comm_export
    FUNCTION ufm_comm_connect( handle, blobs) BIND(C, NAME='ufm_comm_connect') RESULT( rc)
      IMPORT :: C_INT, C_PTR, C_CHAR
      TYPE(C_PTR), VALUE          :: handle
      CHARACTER(KIND=C_CHAR), INTENT(IN) :: blobs( *)
      INTEGER(C_INT)              :: rc
    END FUNCTION ufm_comm_connect
    FUNCTION ufm_mesh_upload_primary( handle, mesh) BIND(C, NAME='ufm_mesh_upload_primary') RESULT( rc)
      IMPORT :: C_PTR, C_INT, ufm_mesh_primary
      TYPE(C_PTR), VALUE         :: handle
      TYPE(ufm_mesh_primary), INTENT(IN) :: mesh
      INTEGER(C_INT)             :: rc
    END FUNCTION
    FUNCTION ufm_mesh_secondary_get( handle, desc, Tricc, Tri_edge_index, VAc, VAaAc, colour) BIND(C, NAME='ufm_mesh_secondary_get') RESULT( rc)
      IMPORT :: C_PTR, C_INT, ufm_mesh_desc
      TYPE(C_PTR), VALUE         :: handle
      TYPE(ufm_mesh_desc), INTENT(OUT) :: desc
      TYPE(C_PTR), INTENT(OUT)   :: Tricc, Tri_edge_index, VAc, VAaAc, colour
      INTEGER(C_INT)             :: rc
    END FUNCTION
    FUNCTION ufm_restart_create( filename, mesh, nZ, zeta) BIND(C, NAME='ufm_restart_create') RESULT( rc)
      IMPORT :: C_CHAR, C_INT, C_DOUBLE, ufm_nc_mesh
      CHARACTER(KIND=C_CHAR), INTENT(IN) :: filename( *)
      TYPE(ufm_nc_mesh), INTENT(IN) :: mesh
      INTEGER(C_INT), VALUE      :: nZ
      REAL(C_DOUBLE), INTENT(IN) :: zeta( *)
      INTEGER(C_INT)             :: rc
    END FUNCTION
    FUNCTION ufm_restart_write( handle, filename, time, FirnDepth, MeltPreviousYear) BIND(C, NAME='ufm_restart_write') RESULT( rc)
      IMPORT :: C_PTR, C_CHAR, C_INT, C_DOUBLE
      TYPE(C_PTR), VALUE         :: handle
      CHARACTER(KIND=C_CHAR), INTENT(IN) :: filename( *)
      REAL(C_DOUBLE), VALUE      :: time
      TYPE(C_PTR), VALUE         :: FirnDepth, MeltPreviousYear
      INTEGER(C_INT)             :: rc
    END FUNCTION
    FUNCTION ufm_restart_inquire_mesh( filename, nV, nTri, nC_mem) BIND(C, NAME='ufm_restart_inquire_mesh') RESULT( rc)
      IMPORT :: C_CHAR, C_INT
      CHARACTER(KIND=C_CHAR), INTENT(IN) :: filename( *)
      INTEGER(C_INT), INTENT(OUT) :: nV, nTri, nC_mem
      INTEGER(C_INT)             :: rc
    END FUNCTION
    FUNCTION ufm_restart_read_mesh( filename, V, nC, C, niTri, iTri, edge_index, Tri, Tricc, TriC, Tri_edge_index) &
        BIND(C, NAME='ufm_restart_read_mesh') RESULT( rc)
      IMPORT :: C_PTR, C_CHAR, C_INT
      CHARACTER(KIND=C_CHAR), INTENT(IN) :: filename( *)
      TYPE(C_PTR), VALUE         :: V, nC, C, niTri, iTri, edge_index, Tri, Tricc, TriC, Tri_edge_index
      INTEGER(C_INT)             :: rc
    END FUNCTION
    FUNCTION ufm_restart_inquire_init( filename, nZ, zeta, nt) BIND(C, NAME='ufm_restart_inquire_init') RESULT( rc)
      IMPORT :: C_CHAR, C_INT, C_DOUBLE
      CHARACTER(KIND=C_CHAR), INTENT(IN) :: filename( *)
      INTEGER(C_INT), VALUE      :: nZ
      REAL(C_DOUBLE), INTENT(IN) :: zeta( *)
      INTEGER(C_INT), INTENT(OUT) :: nt
      INTEGER(C_INT)             :: rc
    END FUNCTION
    FUNCTION ufm_restart_load( handle, filename, time_to_restart_from, FirnDepth, MeltPreviousYear) BIND(C, NAME='ufm_restart_load') RESULT( rc)
      IMPORT :: C_PTR, C_CHAR, C_INT, C_DOUBLE
      TYPE(C_PTR), VALUE         :: handle
      CHARACTER(KIND=C_CHAR), INTENT(IN) :: filename( *)
      REAL(C_DOUBLE), VALUE      :: time_to_restart_from
      TYPE(C_PTR), VALUE         :: FirnDepth, MeltPreviousYear
      INTEGER(C_INT)             :: rc
    END FUNCTION
    FUNCTION ufm_help_fields_create( filename, mesh, nZ, zeta, n_fields, names) BIND(C, NAME='ufm_help_fields_create') RESULT( rc)
      IMPORT :: C_PTR, C_CHAR, C_INT, C_DOUBLE, ufm_nc_mesh
      CHARACTER(KIND=C_CHAR), INTENT(IN) :: filename( *)
      TYPE(ufm_nc_mesh), INTENT(IN) :: mesh
      INTEGER(C_INT), VALUE      :: nZ, n_fields
      REAL(C_DOUBLE), INTENT(IN) :: zeta( *)
      TYPE(C_PTR), INTENT(IN)    :: names( *)
      INTEGER(C_INT)             :: rc
    END FUNCTION
    FUNCTION ufm_help_fields_write( handle, filename, time, n_fields, names, host_data) BIND(C, NAME='ufm_help_fields_write') RESULT( rc)
      IMPORT :: C_PTR, C_CHAR, C_INT, C_DOUBLE
      TYPE(C_PTR), VALUE         :: handle
      CHARACTER(KIND=C_CHAR), INTENT(IN) :: filename( *)
      REAL(C_DOUBLE), VALUE      :: time
      INTEGER(C_INT), VALUE      :: n_fields
      TYPE(C_PTR), INTENT(IN)    :: names( *), host_data( *)
      INTEGER(C_INT)             :: rc
    END FUNCTION
    FUNCTION ufm_last_error() BIND(C, NAME='ufm_last_error') RESULT( msg)
      IMPORT :: C_PTR
      TYPE(C_PTR)                 :: msg
    END FUNCTION ufm_last_error
  END INTERFACE

  ! One handle per model region (NAM, EAS, GRL, ANT), selected by the caller
  TYPE(C_PTR), SAVE :: b200_handle = C_NULL_PTR

CONTAINS

  SUBROUTINE b200_check( rc, where)
    ! Error convention of the reference: message on unit 0, then MPI_ABORT (107 call sites in src/)
    INTEGER(C_INT),   INTENT(IN) :: rc
    CHARACTER(LEN=*), INTENT(IN) :: where
    CHARACTER(KIND=C_CHAR), POINTER :: msg(:)
    INTEGER :: n
    IF (rc == 0) RETURN
    CALL C_F_POINTER( ufm_last_error(), msg, [512])
    n = 1
    DO WHILE (n < 512 .AND. msg( n) /= C_NULL_CHAR)
      n = n + 1
    END DO
    IF (rc > 0) THEN
      WRITE(0,*) msg( 1:n-1)                    ! warnings: the reference prints and carries on
    ELSE
      WRITE(0,*) '  ERROR in ', where, ': ', msg( 1:n-1)
      CALL MPI_ABORT( MPI_COMM_WORLD, cerr, ierr)
    END IF
  END SUBROUTINE b200_check

  FUNCTION b200_benchmark_id() RESULT( id)
    INTEGER(C_INT) :: id
    id = UFM_BM_NONE
    IF (.NOT. C%do_benchmark_experiment) RETURN
    SELECT CASE (TRIM( C%choice_benchmark_experiment))
      CASE ('EISMINT_1'); id = 1
      CASE ('EISMINT_2'); id = 2
      CASE ('EISMINT_3'); id = 3
      CASE ('EISMINT_4'); id = 4
      CASE ('EISMINT_5'); id = 5
      CASE ('EISMINT_6'); id = 6
      CASE ('Halfar');    id = UFM_BM_HALFAR
      CASE ('Bueler');    id = UFM_BM_BUELER
      CASE ('MISMIP_mod');id = UFM_BM_MISMIP_MOD
      CASE ('mesh_generation_test'); id = UFM_BM_MESH_GENERATION_TEST
      CASE ('SSA_icestream');        id = UFM_BM_SSA_ICESTREAM
      CASE DEFAULT
        WRITE(0,*) '  ERROR: benchmark experiment "', TRIM( C%choice_benchmark_experiment), '" unknown to ufemism_b200_shim!'
        CALL MPI_ABORT( MPI_COMM_WORLD, cerr, ierr)
    END SELECT
  END FUNCTION b200_benchmark_id

  SUBROUTINE b200_initialise( device)
    ! Call once after initialize_main_constants (src/UFEMISM_program.f90:103)
    INTEGER, INTENT(IN) :: device
    TYPE(ufm_params)    :: p
    IF (par%master) THEN
      p%nZ = C%nZ
      p%zeta = 0._dp
      p%zeta( 1:C%nZ)           = C%zeta( 1:C%nZ)
      p%m_enh_sia               = C%m_enh_sia
      p%m_enh_ssa               = C%m_enh_ssa
      p%use_analytical_GL_flux  = MERGE( 1, 0, C%use_analytical_GL_flux)
      p%SSA_RN_tol              = C%SSA_RN_tol
      p%SSA_max_outer_loops     = C%SSA_max_outer_loops
      p%SSA_max_residual_UV     = C%SSA_max_residual_UV
      p%SSA_SOR_omega           = C%SSA_SOR_omega
      p%SSA_max_inner_loops     = C%SSA_max_inner_loops
      p%dt_max                  = C%dt_max
      p%benchmark               = b200_benchmark_id()
      p%exact_xy                = 1
      p%dt_thermo               = C%dt_thermo
      CALL b200_check( ufm_create( INT( device, C_INT), p, b200_handle), 'ufm_create')
    END IF
    CALL sync
  END SUBROUTINE b200_initialise

  SUBROUTINE b200_upload_mesh( mesh)
    ! Call at the end of create_final_mesh_from_merged_submesh (src/mesh_creation_module.f90:1737),
    ! read_mesh_from_restart_file (src/restart_module.f90:102) and after the mesh swap
    ! (src/UFEMISM_main_model.f90:294).  All state on the device is reallocated and zero afterwards,
    ! so the host must upload the remapped Hi etc. (the reference remaps only Hi, Hi_prev, Ti, dHb:
    ! src/ice_dynamics_module.f90:1208-1218; U_SSA restarts from zero).
    TYPE(type_mesh), TARGET, INTENT(IN) :: mesh
    TYPE(ufm_mesh_desc) :: d
    IF (par%master) THEN
      d%nV = mesh%nV;  d%nAc = mesh%nAc;  d%nC_mem = mesh%nC_mem
      d%ldV = SIZE( mesh%V, 1);  d%ldAc = SIZE( mesh%Aci, 1);  d%ldAaAc = SIZE( mesh%CAaAc, 1)
      d%V = C_LOC( mesh%V);  d%A = C_LOC( mesh%A);  d%nC = C_LOC( mesh%nC);  d%C = C_LOC( mesh%C);  d%Cw = C_LOC( mesh%Cw)
      d%edge_index = C_LOC( mesh%edge_index);  d%Nx = C_LOC( mesh%Nx);  d%Ny = C_LOC( mesh%Ny)
      d%Aci = C_LOC( mesh%Aci);  d%iAci = C_LOC( mesh%iAci);  d%edge_index_Ac = C_LOC( mesh%edge_index_Ac)
      d%Nx_Ac = C_LOC( mesh%Nx_Ac);  d%Ny_Ac = C_LOC( mesh%Ny_Ac);  d%No_Ac = C_LOC( mesh%No_Ac);  d%Np_Ac = C_LOC( mesh%Np_Ac)
      d%nCAaAc = C_LOC( mesh%nCAaAc);  d%CAaAc = C_LOC( mesh%CAaAc)
      d%Nx_AaAc = C_LOC( mesh%Nx_AaAc);  d%Ny_AaAc = C_LOC( mesh%Ny_AaAc);  d%Nxx_AaAc = C_LOC( mesh%Nxx_AaAc)
      d%Nxy_AaAc = C_LOC( mesh%Nxy_AaAc);  d%Nyy_AaAc = C_LOC( mesh%Nyy_AaAc)
      d%colour_vi = C_LOC( mesh%colour_vi);  d%colour_nV = C_LOC( mesh%colour_nV)
      ! triangle data of the upwind temperature advection (src/mesh_derivatives_module.f90:435-483)
      d%nTri = mesh%nTri;  d%ldTri = SIZE( mesh%Tri, 1)
      d%Tri = C_LOC( mesh%Tri);  d%niTri = C_LOC( mesh%niTri);  d%iTri = C_LOC( mesh%iTri);  d%R = C_LOC( mesh%R)
      d%NxTri = C_LOC( mesh%NxTri);  d%NyTri = C_LOC( mesh%NyTri)
      CALL b200_check( ufm_mesh_upload( b200_handle, d), 'ufm_mesh_upload')
    END IF
    CALL sync
  END SUBROUTINE b200_upload_mesh

  ! ---- the four drop-in entry points -----------------------------------------------------------

  SUBROUTINE calculate_ice_thickness_change_b200( mesh, ice, SMB, BMB, dt, mask_noice)
    TYPE(type_mesh),                     INTENT(IN)    :: mesh
    TYPE(type_ice_model), TARGET,        INTENT(INOUT) :: ice
    TYPE(type_SMB_model), TARGET,        INTENT(IN)    :: SMB
    TYPE(type_BMB_model), TARGET,        INTENT(IN)    :: BMB
    REAL(dp),                            INTENT(IN)    :: dt
    INTEGER,  DIMENSION(:    ), TARGET,  INTENT(IN)    :: mask_noice
    IF (par%master) THEN
      ! written by CPU components since the last call: SMB/BMB (every dt_SMB), ELRA (Hb), remapping (Hi)
      CALL b200_check( ufm_state_upload( b200_handle, UFM_F_SMB_YEAR,   C_LOC( SMB%SMB_year)), 'upload SMB_year')
      CALL b200_check( ufm_state_upload( b200_handle, UFM_F_BMB,        C_LOC( BMB%BMB     )), 'upload BMB')
      CALL b200_check( ufm_state_upload( b200_handle, UFM_F_MASK_NOICE, C_LOC( mask_noice  )), 'upload mask_noice')
      CALL b200_check( ufm_state_upload( b200_handle, UFM_F_HI,         C_LOC( ice%Hi      )), 'upload Hi')
      CALL b200_check( ufm_thickness_update( b200_handle, dt), 'calculate_ice_thickness_change')
      CALL b200_check( ufm_state_download( b200_handle, UFM_F_HI,      C_LOC( ice%Hi     )), 'download Hi')
      CALL b200_check( ufm_state_download( b200_handle, UFM_F_HI_PREV, C_LOC( ice%Hi_prev)), 'download Hi_prev')
      CALL b200_check( ufm_state_download( b200_handle, UFM_F_DHI_DT,  C_LOC( ice%dHi_dt )), 'download dHi_dt')
    END IF
    CALL sync
  END SUBROUTINE calculate_ice_thickness_change_b200

  SUBROUTINE update_general_ice_model_data_b200( mesh, ice, time)
    TYPE(type_mesh),                     INTENT(IN)    :: mesh
    TYPE(type_ice_model), TARGET,        INTENT(INOUT) :: ice
    REAL(dp),                            INTENT(IN)    :: time
    IF (par%master) THEN
      CALL b200_check( ufm_state_upload( b200_handle, UFM_F_HB,     C_LOC( ice%Hb    )), 'upload Hb')
      CALL b200_check( ufm_state_upload( b200_handle, UFM_F_SL,     C_LOC( ice%SL    )), 'upload SL')
      CALL b200_check( ufm_state_upload( b200_handle, UFM_F_DHB_DT, C_LOC( ice%dHb_dt)), 'upload dHb_dt')
      CALL b200_check( ufm_update_general( b200_handle, time), 'update_general_ice_model_data')
      ! what CPU components read: surface elevation, slopes (climate/SMB), masks (BMB, mesh fitness, output)
      CALL b200_check( ufm_state_download( b200_handle, UFM_F_HS,          C_LOC( ice%Hs         )), 'download Hs')
      CALL b200_check( ufm_state_download( b200_handle, UFM_F_DHS_DT,      C_LOC( ice%dHs_dt     )), 'download dHs_dt')
      CALL b200_check( ufm_state_download( b200_handle, UFM_F_DHS_DX,      C_LOC( ice%dHs_dx     )), 'download dHs_dx')
      CALL b200_check( ufm_state_download( b200_handle, UFM_F_DHS_DY,      C_LOC( ice%dHs_dy     )), 'download dHs_dy')
      CALL b200_check( ufm_state_download( b200_handle, UFM_F_DHI_DX,      C_LOC( ice%dHi_dx     )), 'download dHi_dx')
      CALL b200_check( ufm_state_download( b200_handle, UFM_F_DHI_DY,      C_LOC( ice%dHi_dy     )), 'download dHi_dy')
      CALL b200_check( ufm_state_download( b200_handle, UFM_F_MASK,        C_LOC( ice%mask       )), 'download mask')
      CALL b200_check( ufm_state_download( b200_handle, UFM_F_MASK_LAND,   C_LOC( ice%mask_land  )), 'download mask_land')
      CALL b200_check( ufm_state_download( b200_handle, UFM_F_MASK_OCEAN,  C_LOC( ice%mask_ocean )), 'download mask_ocean')
      CALL b200_check( ufm_state_download( b200_handle, UFM_F_MASK_ICE,    C_LOC( ice%mask_ice   )), 'download mask_ice')
      CALL b200_check( ufm_state_download( b200_handle, UFM_F_MASK_SHEET,  C_LOC( ice%mask_sheet )), 'download mask_sheet')
      CALL b200_check( ufm_state_download( b200_handle, UFM_F_MASK_SHELF,  C_LOC( ice%mask_shelf )), 'download mask_shelf')
      CALL b200_check( ufm_state_download( b200_handle, UFM_F_MASK_COAST,  C_LOC( ice%mask_coast )), 'download mask_coast')
      CALL b200_check( ufm_state_download( b200_handle, UFM_F_MASK_MARGIN, C_LOC( ice%mask_margin)), 'download mask_margin')
      CALL b200_check( ufm_state_download( b200_handle, UFM_F_MASK_GL,     C_LOC( ice%mask_gl    )), 'download mask_gl')
      CALL b200_check( ufm_state_download( b200_handle, UFM_F_MASK_CF,     C_LOC( ice%mask_cf    )), 'download mask_cf')
    END IF
    CALL sync
  END SUBROUTINE update_general_ice_model_data_b200

  SUBROUTINE solve_SIA_b200( mesh, ice)
    TYPE(type_mesh),                     INTENT(IN)    :: mesh
    TYPE(type_ice_model), TARGET,        INTENT(INOUT) :: ice
    IF (par%master) THEN
      CALL b200_check( ufm_solve_SIA( b200_handle), 'solve_SIA')
      ! diagnostic Aa fields (output, U_vav: src/UFEMISM_main_model.f90:178-181)
      CALL b200_check( ufm_state_download( b200_handle, UFM_F_U_SIA, C_LOC( ice%U_SIA)), 'download U_SIA')
      CALL b200_check( ufm_state_download( b200_handle, UFM_F_V_SIA, C_LOC( ice%V_SIA)), 'download V_SIA')
      CALL b200_check( ufm_state_download( b200_handle, UFM_F_D_SIA, C_LOC( ice%D_SIA)), 'download D_SIA')
    END IF
    CALL sync
  END SUBROUTINE solve_SIA_b200

  SUBROUTINE solve_SSA_b200( mesh, ice)
    TYPE(type_mesh),                     INTENT(IN)    :: mesh
    TYPE(type_ice_model), TARGET,        INTENT(INOUT) :: ice
    TYPE(ufm_ssa_stats) :: stats
    IF (par%master) THEN
      ! rc = 1 -> ' WARNING - SSA SOR solver doesnt converge!' (printed, run continues: ice_dynamics_module.f90:686)
      ! rc = -1 -> 'solve_SSA - ERROR: SSA remains unstable after resetting velocities to zero!' + MPI_ABORT (:537-538)
      CALL b200_check( ufm_solve_SSA( b200_handle, stats), 'solve_SSA')
      CALL b200_check( ufm_state_download( b200_handle, UFM_F_U_SSA, C_LOC( ice%U_SSA)), 'download U_SSA')
      CALL b200_check( ufm_state_download( b200_handle, UFM_F_V_SSA, C_LOC( ice%V_SSA)), 'download V_SSA')
    END IF
    CALL sync
  END SUBROUTINE solve_SSA_b200

  SUBROUTINE b200_connect_gpus()
    ! Vertex-partitioned run, one MPI rank per GPU (instead of "master only"): call after b200_upload_mesh on every rank.
    ! MPI_ALLGATHER stands where the Python drivers use torch.distributed.
    CHARACTER(KIND=C_CHAR) :: blob( 256)
    CHARACTER(KIND=C_CHAR), ALLOCATABLE :: blobs(:)
    ALLOCATE( blobs( 256 * par%n))
    CALL b200_check( ufm_comm_export( b200_handle, blob), 'ufm_comm_export')
    CALL MPI_ALLGATHER( blob, 256, MPI_CHARACTER, blobs, 256, MPI_CHARACTER, MPI_COMM_WORLD, ierr)
    CALL b200_check( ufm_comm_connect( b200_handle, blobs), 'ufm_comm_connect')
    CALL sync
    DEALLOCATE( blobs)
  END SUBROUTINE b200_connect_gpus

  SUBROUTINE solve_SIA_3D_b200( mesh, ice)
    ! U_3D / V_3D half of solve_SIA_3D (src/ice_dynamics_module.f90:317-367), what the critical time step reads; the
    ! vertical velocity belongs to update_ice_temperature_b200
    TYPE(type_mesh),                     INTENT(IN)    :: mesh
    TYPE(type_ice_model), TARGET,        INTENT(INOUT) :: ice
    IF (par%master) CALL b200_check( ufm_solve_SIA_3D( b200_handle), 'solve_SIA_3D')
    CALL sync
  END SUBROUTINE solve_SIA_3D_b200

  SUBROUTINE update_ice_temperature_b200( mesh, ice, climate, SMB)
    ! Drop-in for update_ice_temperature (src/thermodynamics_module.f90:23-202): the CPU climate / SMB components own T2m and
    ! SMB_year, the geothermal heat flux is static; Ti lives on the device between calls and is downloaded for output / restart.
    TYPE(type_mesh),                     INTENT(IN)    :: mesh
    TYPE(type_ice_model), TARGET,        INTENT(INOUT) :: ice
    TYPE(type_climate_model), TARGET,    INTENT(IN)    :: climate
    TYPE(type_SMB_model), TARGET,        INTENT(IN)    :: SMB
    TYPE(ufm_thermo_stats) :: st
    IF (par%master) THEN
      CALL b200_check( ufm_state_upload( b200_handle, UFM_F_T2M,      C_LOC( climate%applied%T2m)), 'upload T2m')
      CALL b200_check( ufm_state_upload( b200_handle, UFM_F_SMB_YEAR, C_LOC( SMB%SMB_year)),        'upload SMB_year')
      CALL b200_check( ufm_state_upload( b200_handle, UFM_F_GHF,      C_LOC( ice%GHF)),             'upload GHF')
      ! rc -8 = "heat equation solver unstable for more than 1% of vertices" (STOP at :195-199), -9 = DGTSV info /= 0 (STOP at :353)
      CALL b200_check( ufm_update_ice_temperature( b200_handle, st), 'update_ice_temperature')
      CALL b200_check( ufm_state_download( b200_handle, UFM_F_TI,   C_LOC( ice%Ti)),   'download Ti')
      CALL b200_check( ufm_state_download( b200_handle, UFM_F_U_3D, C_LOC( ice%U_3D)), 'download U_3D')
      CALL b200_check( ufm_state_download( b200_handle, UFM_F_V_3D, C_LOC( ice%V_3D)), 'download V_3D')
      CALL b200_check( ufm_state_download( b200_handle, UFM_F_W_3D, C_LOC( ice%W_3D)), 'download W_3D')
    END IF
    CALL sync
  END SUBROUTINE update_ice_temperature_b200

  SUBROUTINE critical_timesteps_b200( dt_D_2D_min, dt_V_2D_SSA_min, dt_V_3D_SIA_min)
    ! Replaces the three loops + MPI_ALLREDUCE MIN of determine_timesteps_and_actions
    ! (src/UFEMISM_main_model.f90:747-778); the values come back already multiplied by 0.9.
    REAL(dp), INTENT(OUT) :: dt_D_2D_min, dt_V_2D_SSA_min, dt_V_3D_SIA_min
    REAL(C_DOUBLE) :: out3( 3)
    IF (par%master) CALL b200_check( ufm_cfl( b200_handle, out3), 'determine_timesteps_and_actions')
    CALL MPI_BCAST( out3, 3, MPI_DOUBLE_PRECISION, 0, MPI_COMM_WORLD, ierr)
    dt_D_2D_min = out3( 1);  dt_V_2D_SSA_min = out3( 2);  dt_V_3D_SIA_min = out3( 3)
  END SUBROUTINE critical_timesteps_b200

  ! ---- row N3: mesh update / restart without the CPU rebuild of the secondary mesh data --------------------------------

  SUBROUTINE b200_upload_mesh_primary( mesh)
    ! Call INSTEAD of the block find_Voronoi_cell_areas .. calculate_five_colouring_AaAc of read_mesh_from_restart_file
    ! (src/restart_module.f90:88-103) / create_final_mesh_from_merged_submesh (src/mesh_creation_module.f90:1724-1737) when no
    ! CPU component needs those arrays: only the primary mesh data must be valid.  The few derived arrays the host still reads
    ! (A, Cw, Aci, iAci, VAc, colour lists) can be copied back with ufm_mesh_secondary_get.
    TYPE(type_mesh), TARGET, INTENT(IN) :: mesh
    TYPE(ufm_mesh_primary) :: p
    IF (par%master) THEN
      p%nV = mesh%nV;  p%nTri = mesh%nTri;  p%nC_mem = mesh%nC_mem
      p%ldV = SIZE( mesh%V, 1);  p%ldTri = SIZE( mesh%Tri, 1)
      p%xmin = mesh%xmin;  p%xmax = mesh%xmax;  p%ymin = mesh%ymin;  p%ymax = mesh%ymax
      p%V = C_LOC( mesh%V);  p%nC = C_LOC( mesh%nC);  p%C = C_LOC( mesh%C);  p%niTri = C_LOC( mesh%niTri);  p%iTri = C_LOC( mesh%iTri)
      p%edge_index = C_LOC( mesh%edge_index);  p%Tri = C_LOC( mesh%Tri)
      p%thermo = 1
      CALL b200_check( ufm_mesh_upload_primary( b200_handle, p), 'ufm_mesh_upload_primary')
    END IF
    CALL sync
  END SUBROUTINE b200_upload_mesh_primary

  ! ---- row N4: restart / help_fields files written from and read onto the device --------------------------------------

  FUNCTION b200_cstring( f) RESULT( c)
    CHARACTER(LEN=*), INTENT(IN) :: f
    CHARACTER(KIND=C_CHAR)       :: c( LEN_TRIM( f) + 1)
    INTEGER :: n
    DO n = 1, LEN_TRIM( f)
      c( n) = f( n:n)
    END DO
    c( LEN_TRIM( f) + 1) = C_NULL_CHAR
  END FUNCTION b200_cstring

  SUBROUTINE b200_nc_mesh( mesh, m)
    TYPE(type_mesh), TARGET, INTENT(IN)  :: mesh
    TYPE(ufm_nc_mesh),       INTENT(OUT) :: m
    m%nV = mesh%nV;  m%nTri = mesh%nTri;  m%nC_mem = mesh%nC_mem;  m%nAc = mesh%nAc;  m%nV_transect = mesh%nV_transect
    m%nVAaAc = mesh%nVAaAc;  m%nTriAaAc = mesh%nTriAaAc
    m%V = C_LOC( mesh%V);  m%Tri = C_LOC( mesh%Tri);  m%nC = C_LOC( mesh%nC);  m%C = C_LOC( mesh%C);  m%niTri = C_LOC( mesh%niTri)
    m%iTri = C_LOC( mesh%iTri);  m%edge_index = C_LOC( mesh%edge_index);  m%Tricc = C_LOC( mesh%Tricc);  m%TriC = C_LOC( mesh%TriC)
    m%Tri_edge_index = C_LOC( mesh%Tri_edge_index);  m%VAc = C_LOC( mesh%VAc);  m%Aci = C_LOC( mesh%Aci);  m%iAci = C_LOC( mesh%iAci)
    m%VAaAc = C_LOC( mesh%VAaAc);  m%TriAaAc = C_LOC( mesh%TriAaAc);  m%A = C_LOC( mesh%A);  m%R = C_LOC( mesh%R)
    m%vi_transect = C_LOC( mesh%vi_transect);  m%w_transect = C_LOC( mesh%w_transect)
  END SUBROUTINE b200_nc_mesh

  SUBROUTINE create_restart_file_mesh_b200( region, netcdf)
    ! Drop-in for create_restart_file_mesh (src/netcdf_module.f90:489-633): same file, no NetCDF library
    TYPE(type_model_region),   INTENT(INOUT) :: region
    TYPE(type_netcdf_restart), INTENT(INOUT) :: netcdf
    TYPE(ufm_nc_mesh) :: m
    IF (.NOT. par%master) RETURN
    netcdf%ti = 1
    CALL b200_nc_mesh( region%mesh, m)
    CALL b200_check( ufm_restart_create( b200_cstring( netcdf%filename), m, INT( C%nZ, C_INT), C%zeta), 'create_restart_file_mesh')
  END SUBROUTINE create_restart_file_mesh_b200

  SUBROUTINE write_to_restart_file_mesh_b200( region, netcdf)
    ! Drop-in for write_to_restart_file_mesh (src/netcdf_module.f90:180-214): Hi, Hb, Hs, U/V_SIA, U/V_SSA, Ti come straight
    ! from the device; FirnDepth and MeltPreviousYear belong to the CPU SMB model
    TYPE(type_model_region), TARGET, INTENT(INOUT) :: region
    TYPE(type_netcdf_restart),       INTENT(INOUT) :: netcdf
    INTEGER(C_INT) :: ti
    IF (.NOT. par%master) RETURN
    ti = ufm_restart_write( b200_handle, b200_cstring( netcdf%filename), region%time, C_LOC( region%SMB%FirnDepth), C_LOC( region%SMB%MeltPreviousYear))
    IF (ti < 0) CALL b200_check( ti, 'write_to_restart_file_mesh')
    netcdf%ti = ti + 1
  END SUBROUTINE write_to_restart_file_mesh_b200

  SUBROUTINE read_mesh_from_restart_file_b200( region)
    ! Drop-in for read_mesh_from_restart_file (src/restart_module.f90:31-116): primary mesh data from the file, everything
    ! else derived inside the library and resident on the device when this returns
    TYPE(type_model_region), TARGET, INTENT(INOUT) :: region
    INTEGER(C_INT) :: nV, nTri, nC_mem
    IF (par%master) CALL b200_check( ufm_restart_inquire_mesh( b200_cstring( region%init%netcdf_restart%filename), nV, nTri, nC_mem), 'inquire_restart_file_mesh')
    CALL MPI_BCAST( nV,     1, MPI_INTEGER, 0, MPI_COMM_WORLD, ierr)
    CALL MPI_BCAST( nTri,   1, MPI_INTEGER, 0, MPI_COMM_WORLD, ierr)
    CALL MPI_BCAST( nC_mem, 1, MPI_INTEGER, 0, MPI_COMM_WORLD, ierr)
    CALL allocate_mesh_primary( region%mesh, region%name, nV, nTri, nC_mem)
    IF (par%master) THEN
      region%mesh%nV = nV;  region%mesh%nTri = nTri
      CALL b200_check( ufm_restart_read_mesh( b200_cstring( region%init%netcdf_restart%filename), C_LOC( region%mesh%V), C_LOC( region%mesh%nC), &
        C_LOC( region%mesh%C), C_LOC( region%mesh%niTri), C_LOC( region%mesh%iTri), C_LOC( region%mesh%edge_index), C_LOC( region%mesh%Tri), &
        C_LOC( region%mesh%Tricc), C_LOC( region%mesh%TriC), C_LOC( region%mesh%Tri_edge_index)), 'read_restart_file_mesh')
      region%mesh%xmin = MINVAL( region%mesh%V( 1:nV,1));  region%mesh%xmax = MAXVAL( region%mesh%V( 1:nV,1))
      region%mesh%ymin = MINVAL( region%mesh%V( 1:nV,2));  region%mesh%ymax = MAXVAL( region%mesh%V( 1:nV,2))
    END IF
    CALL sync
    CALL b200_upload_mesh_primary( region%mesh)
  END SUBROUTINE read_mesh_from_restart_file_b200

  SUBROUTINE read_init_data_from_restart_file_b200( region)
    ! Drop-in for read_init_data_from_restart_file (src/restart_module.f90:118-142): the time frame closest to
    ! C%time_to_restart_from goes straight to the device (Hi, Hb, Ti, U_SSA, V_SSA); the SMB model's fields to the host
    TYPE(type_model_region), TARGET, INTENT(INOUT) :: region
    INTEGER(C_INT) :: rc, nt
    IF (par%master) THEN
      rc = ufm_restart_inquire_init( b200_cstring( region%init%netcdf_restart%filename), INT( C%nZ, C_INT), C%zeta, nt)
      CALL b200_check( rc, 'inquire_restart_file_init')        ! rc 1: zeta differs, the reference only warns
      rc = ufm_restart_load( b200_handle, b200_cstring( region%init%netcdf_restart%filename), C%time_to_restart_from, &
                             C_LOC( region%init%FirnDepth), C_LOC( region%init%MeltPreviousYear))
      IF (rc < 0) CALL b200_check( rc, 'read_restart_file_init')
    END IF
    CALL sync
  END SUBROUTINE read_init_data_from_restart_file_b200

END MODULE ufemism_b200_shim
