!  m_gritas_b200.F90 -- ISO_C_BINDING shim between the unchanged GrITAS Fortran driver and
!  libgritas_b200.so (include/gritas_b200.h).
!
!  NOT COMPILED IN THIS REPOSITORY: the build image has no Fortran compiler.  tools/check_fortran_shim.py
!  (run by tests/test_fortran_shim.py) checks what can be checked without one: every bind(C) interface
!  below matches its prototype in include/gritas_b200.h argument for argument, and the module
!  dependency graph  m_gritas_grids -> m_gritas_b200 -> { iso_c_binding, m_odsmeta }  has no cycle.
!
!  Module dependencies: this module uses NOTHING of m_gritas_grids.  The grid dimensions im, jm, km,
!  nlevs (module variables of m_gritas_grids, m_gritas_grids.f:24-27) come in as arguments.  The
!  reference-side change is fortran/m_gritas_grids_b200.inc: the bodies of GrITAS_Accum_ /
!  GrITAS_Norm_ / GrITAS_MinMax_ / GrITAS_3CH_Norm_ inside m_gritas_grids become one-line forwards
!  to the procedures here (m_gritas_grids uses m_gritas_b200, never the other way round).
!  gritas.f, grisas.f, the rc/namelist handling and the GrADS/NetCDF writers stay untouched.
!
!  Semantics that differ from the host implementation (see INTEGRATION.md):
!    * the accumulators live on the GPU; the host grids passed to B200_Accum are not modified
!      until B200_Norm (or B200_GetRaw) fills them.  Code that reads the raw sums between the
!      two calls (gritas.f:812-818 -nonorm, the scorecard) must call B200_GetRaw first;
!    * grisas.f:285-296 zeroes the host grids at every synoptic time.  B200_Accum notices that the
!      host count grids it filled at the last B200_Norm are zero again and resets the device
!      accumulators (B200_Reset does it explicitly);
!    * B200_SetMask may be called before the first B200_Accum (the mask is kept and applied when the
!      context is created, so that the first file is masked like every other one: gritas.f:745);
!    * B200_Norm without any B200_Accum (a run without observations) returns what GrITAS_Norm_ makes
!      of all-zero grids: AMISS statistics, zero counts;
!    * all reals are 64-bit (the library is built for the FREAL8 configuration of
!      src/Components/gritas/CMakeLists.txt:3-4).

      module m_gritas_b200

      use, intrinsic :: iso_c_binding
      implicit none
      private

      public :: B200_SetMode, B200_SetFlags
      public :: B200_Accum, B200_Accum_ODS, B200_Norm, B200_3CH_Norm, B200_MinMax, B200_SetMask
      public :: B200_GetRaw, B200_Reset, B200_Wait, B200_Clean, B200_PreSort
      public :: B200_UniqueId, B200_CommInit, B200_PeerExport, B200_PeerAttach, B200_Slab
      public :: B200_Norm_Sharded, B200_Norm_Reduce
      public :: B200_XCov_Accum, B200_XCov_Prs_Accum, B200_XCov_Norm, B200_XCov_Clean
      public :: B200_PEER_BLOB_BYTES

      integer(c_int), parameter :: MODE_FAST = 0, MODE_DETERMINISTIC = 1
      integer(c_int), parameter :: FLAG_ASYNC = 256, FLAG_MINMAX = 8192, FLAG_NO_WRITEBACK = 16384
      integer(c_int), parameter :: ERR_LEVEL = 7
      integer, parameter        :: B200_PEER_BLOB_BYTES = 512
      real(c_double), parameter :: AMISS = 1.0D15                      ! m_gritas_grids.f:72

      type(c_ptr), save      :: handle = c_null_ptr
      type(c_ptr), save      :: xhandle = c_null_ptr            ! cross-covariance context (GrITAS_XCov_*)
      integer(c_int), save   :: mode   = MODE_FAST
      integer(c_int), save   :: flags  = 0
!     host count grids filled by the last B200_Norm: position of one non-zero count (0: none)
      logical, save          :: normed = .false.
      integer(c_size_t), save :: sent2 = 0, sent3 = 0
!     mask handed over before the context exists
      real(c_double), allocatable, save :: pending_mask(:)

      interface
         integer(c_int) function gritas_b200_create ( h, im, jm, km, nlevs, l2d, l3d, nr,       &
                                  p1, p2, kxkt_table, kxmax, ktmax, mode, device )              &
                                  bind(C, name='gritas_b200_create')
            import :: c_ptr, c_int, c_double
            type(c_ptr),    intent(out)       :: h
            integer(c_int), value             :: im, jm, km, nlevs, l2d, l3d, nr
            real(c_double), intent(in)        :: p1(*), p2(*)
            integer(c_int), intent(in)        :: kxkt_table(*)
            integer(c_int), value             :: kxmax, ktmax, mode, device
         end function
         subroutine gritas_b200_destroy ( h ) bind(C, name='gritas_b200_destroy')
            import :: c_ptr
            type(c_ptr), value :: h
         end subroutine
         integer(c_int) function gritas_b200_accum ( h, nobs, lat, lon, lev, del, kx, kt,       &
                                  ob_flag, bad_lev ) bind(C, name='gritas_b200_accum')
            import :: c_ptr, c_int, c_double
            type(c_ptr),    value             :: h
            integer(c_int), value             :: nobs
            real(c_double), intent(in)        :: lat(*), lon(*), lev(*), del(*)
            integer(c_int), intent(in)        :: kx(*), kt(*)
            integer(c_int), intent(inout)     :: ob_flag(*)
            real(c_double), intent(out)       :: bad_lev
         end function
         integer(c_int) function gritas_b200_accum_ods ( h, nobs, lat, lon, lev, time, kt, kx,  &
                                  qcexcl, omf, oma, obs, xvec, iopt, scale_res, ioptn, t1, t2,  &
                                  lona, lonb, x_passive, x_pre_bad, ob_flag_out, bad_lev )      &
                                  bind(C, name='gritas_b200_accum_ods')
            import :: c_ptr, c_int, c_double
            type(c_ptr),    value             :: h
            integer(c_int), value             :: nobs
            real(c_double), intent(in)        :: lat(*), lon(*), lev(*)
            integer(c_int), intent(in)        :: time(*), kt(*), kx(*), qcexcl(*)
            real(c_double), intent(in)        :: omf(*), oma(*), obs(*), xvec(*)
            integer(c_int), value             :: iopt, scale_res, ioptn, t1, t2
            real(c_double), value             :: lona, lonb
            integer(c_int), value             :: x_passive, x_pre_bad
            integer(c_int), intent(inout)     :: ob_flag_out(*)
            real(c_double), intent(out)       :: bad_lev
         end function
         integer(c_int) function gritas_b200_norm ( h, nrmlz, ntresh,                           &
                                  bias_2d, rtms_2d, stdv_2d, xcov_2d, nobs_2d,                  &
                                  bias_3d, rtms_3d, stdv_3d, xcov_3d, nobs_3d )                 &
                                  bind(C, name='gritas_b200_norm')
            import :: c_ptr, c_int, c_double
            type(c_ptr),    value             :: h
            integer(c_int), value             :: nrmlz, ntresh
            real(c_double), intent(inout)     :: bias_2d(*), rtms_2d(*), stdv_2d(*), xcov_2d(*), nobs_2d(*)
            real(c_double), intent(inout)     :: bias_3d(*), rtms_3d(*), stdv_3d(*), xcov_3d(*), nobs_3d(*)
         end function
         integer(c_int) function gritas_b200_norm_sharded ( h, nrmlz, ntresh,                   &
                                  bias_2d, rtms_2d, stdv_2d, xcov_2d, nobs_2d,                  &
                                  bias_3d, rtms_3d, stdv_3d, xcov_3d, nobs_3d )                 &
                                  bind(C, name='gritas_b200_norm_sharded')
            import :: c_ptr, c_int, c_double
            type(c_ptr),    value             :: h
            integer(c_int), value             :: nrmlz, ntresh
            real(c_double), intent(inout)     :: bias_2d(*), rtms_2d(*), stdv_2d(*), xcov_2d(*), nobs_2d(*)
            real(c_double), intent(inout)     :: bias_3d(*), rtms_3d(*), stdv_3d(*), xcov_3d(*), nobs_3d(*)
         end function
         integer(c_int) function gritas_b200_norm_reduce ( h, root, nrmlz, ntresh,              &
                                  bias_2d, rtms_2d, stdv_2d, xcov_2d, nobs_2d,                  &
                                  bias_3d, rtms_3d, stdv_3d, xcov_3d, nobs_3d )                 &
                                  bind(C, name='gritas_b200_norm_reduce')
            import :: c_ptr, c_int, c_double
            type(c_ptr),    value             :: h
            integer(c_int), value             :: root, nrmlz, ntresh
            real(c_double), intent(inout)     :: bias_2d(*), rtms_2d(*), stdv_2d(*), xcov_2d(*), nobs_2d(*)
            real(c_double), intent(inout)     :: bias_3d(*), rtms_3d(*), stdv_3d(*), xcov_3d(*), nobs_3d(*)
         end function
         integer(c_int) function gritas_b200_norm_3ch ( h, nrmlz, ntresh,                       &
                                  bias_2d, rtms_2d, var_2d, xcov_2d, nobs_2d,                   &
                                  bias_3d, rtms_3d, var_3d, xcov_3d, nobs_3d )                  &
                                  bind(C, name='gritas_b200_norm_3ch')
            import :: c_ptr, c_int, c_double
            type(c_ptr),    value             :: h
            integer(c_int), value             :: nrmlz, ntresh
            real(c_double), intent(inout)     :: bias_2d(*), rtms_2d(*), var_2d(*), xcov_2d(*), nobs_2d(*)
            real(c_double), intent(inout)     :: bias_3d(*), rtms_3d(*), var_3d(*), xcov_3d(*), nobs_3d(*)
         end function
         integer(c_int) function gritas_b200_presort ( h, nobs, qcexcl, lat, lon, lev, time,    &
                                  kt, ks, kx, index_base, is )                                  &
                                  bind(C, name='gritas_b200_presort')
            import :: c_ptr, c_int, c_double
            type(c_ptr),    value             :: h
            integer(c_int), value             :: nobs
            integer(c_int), intent(in)        :: qcexcl(*)
            real(c_double), intent(in)        :: lat(*), lon(*), lev(*)
            integer(c_int), intent(in)        :: time(*), kt(*), ks(*), kx(*)
            integer(c_int), value             :: index_base
            integer(c_int), intent(out)       :: is(*)
         end function
         integer(c_int) function gritas_b200_get_raw ( h, bias_2d, rtms_2d, xcov_2d, nobs_2d,   &
                                  bias_3d, rtms_3d, xcov_3d, nobs_3d )                          &
                                  bind(C, name='gritas_b200_get_raw')
            import :: c_ptr, c_int, c_double
            type(c_ptr),    value             :: h
            real(c_double), intent(inout)     :: bias_2d(*), rtms_2d(*), xcov_2d(*), nobs_2d(*)
            real(c_double), intent(inout)     :: bias_3d(*), rtms_3d(*), xcov_3d(*), nobs_3d(*)
         end function
         integer(c_int) function gritas_b200_get_minmax ( h, minv_2d, maxv_2d, nobs_2d,         &
                                  minv_3d, maxv_3d, nobs_3d ) bind(C, name='gritas_b200_get_minmax')
            import :: c_ptr, c_int, c_double
            type(c_ptr),    value             :: h
            real(c_double), intent(inout)     :: minv_2d(*), maxv_2d(*), nobs_2d(*)
            real(c_double), intent(inout)     :: minv_3d(*), maxv_3d(*), nobs_3d(*)
         end function
         integer(c_int) function gritas_b200_set_mask ( h, maskfld ) bind(C, name='gritas_b200_set_mask')
            import :: c_ptr, c_int, c_double
            type(c_ptr),    value             :: h
            real(c_double), intent(in)        :: maskfld(*)
         end function
         integer(c_int) function gritas_b200_reset ( h ) bind(C, name='gritas_b200_reset')
            import :: c_ptr, c_int
            type(c_ptr), value :: h
         end function
         integer(c_int) function gritas_b200_wait ( h, bad_lev ) bind(C, name='gritas_b200_wait')
            import :: c_ptr, c_int, c_double
            type(c_ptr),    value             :: h
            real(c_double), intent(out)       :: bad_lev
         end function
         integer(c_int) function gritas_b200_nccl_unique_id ( id128 ) bind(C, name='gritas_b200_nccl_unique_id')
            import :: c_int, c_char
            character(kind=c_char), intent(out) :: id128(*)
         end function
         integer(c_int) function gritas_b200_comm_init ( h, nranks, rank, id128 ) bind(C, name='gritas_b200_comm_init')
            import :: c_ptr, c_int, c_char
            type(c_ptr),    value             :: h
            integer(c_int), value             :: nranks, rank
            character(kind=c_char), intent(in) :: id128(*)
         end function
         integer(c_int) function gritas_b200_peer_export ( h, blob ) bind(C, name='gritas_b200_peer_export')
            import :: c_ptr, c_int, c_char
            type(c_ptr),    value             :: h
            character(kind=c_char), intent(out) :: blob(*)
         end function
         integer(c_int) function gritas_b200_peer_attach ( h, nranks, rank, blobs ) bind(C, name='gritas_b200_peer_attach')
            import :: c_ptr, c_int, c_char
            type(c_ptr),    value             :: h
            integer(c_int), value             :: nranks, rank
            character(kind=c_char), intent(in) :: blobs(*)
         end function
         integer(c_int) function gritas_b200_slab ( h, rank, nranks, jl3_0, jl3_1, jl2_0, jl2_1 ) bind(C, name='gritas_b200_slab')
            import :: c_ptr, c_int
            type(c_ptr),    value             :: h
            integer(c_int), value             :: rank, nranks
            integer(c_int), intent(out)       :: jl3_0, jl3_1, jl2_0, jl2_1
         end function
         integer(c_int) function gritas_b200_xcov_create ( h, km, nchan, achan, l3d, nr, tch, kxkt_table, kxmax, ktmax, device ) &
                                  bind(C, name='gritas_b200_xcov_create')
            import :: c_ptr, c_int
            type(c_ptr),    intent(out)       :: h
            integer(c_int), value             :: km, nchan
            integer(c_int), intent(in)        :: achan(*)
            integer(c_int), value             :: l3d, nr, tch
            integer(c_int), intent(in)        :: kxkt_table(*)
            integer(c_int), value             :: kxmax, ktmax, device
         end function
         integer(c_int) function gritas_b200_xcov_prs_create ( h, nlevs, p1, p2, l3d, nr, tch, device )                         &
                                  bind(C, name='gritas_b200_xcov_prs_create')
            import :: c_ptr, c_int, c_double
            type(c_ptr),    intent(out)       :: h
            integer(c_int), value             :: nlevs
            real(c_double), intent(in)        :: p1(*), p2(*)
            integer(c_int), value             :: l3d, nr, tch, device
         end function
         integer(c_int) function gritas_b200_xcov_accum ( h, nobs, lev, kx, kt, del, ob_flag, nprof )                           &
                                  bind(C, name='gritas_b200_xcov_accum')
            import :: c_ptr, c_int, c_double
            type(c_ptr),    value             :: h
            integer(c_int), value             :: nobs
            real(c_double), intent(in)        :: lev(*)
            integer(c_int), intent(in)        :: kx(*), kt(*)
            real(c_double), intent(in)        :: del(*)
            integer(c_int), intent(inout)     :: ob_flag(*)
            integer(c_int), intent(out)       :: nprof
         end function
         integer(c_int) function gritas_b200_xcov_prs_accum ( h, nobs, lev, ks, del, nprof, bad_lev )                           &
                                  bind(C, name='gritas_b200_xcov_prs_accum')
            import :: c_ptr, c_int, c_double
            type(c_ptr),    value             :: h
            integer(c_int), value             :: nobs
            real(c_double), intent(in)        :: lev(*)
            integer(c_int), intent(in)        :: ks(*)
            real(c_double), intent(in)        :: del(*)
            integer(c_int), intent(out)       :: nprof
            real(c_double), intent(out)       :: bad_lev
         end function
         integer(c_int) function gritas_b200_xcov_norm ( h, nrmlz, self, ntresh, bias_lv, xcov_lv, nobs_lv )                    &
                                  bind(C, name='gritas_b200_xcov_norm')
            import :: c_ptr, c_int, c_double
            type(c_ptr),    value             :: h
            integer(c_int), value             :: nrmlz, self, ntresh
            real(c_double), intent(inout)     :: bias_lv(*), xcov_lv(*), nobs_lv(*)
         end function
         subroutine gritas_b200_xcov_destroy ( h ) bind(C, name='gritas_b200_xcov_destroy')
            import :: c_ptr
            type(c_ptr), value :: h
         end subroutine
      end interface

      contains

!     Select the accumulation mode before the first B200_Accum (0 fast, 1 deterministic).
      subroutine B200_SetMode ( m )
      integer, intent(in) :: m
      mode = int(m, c_int)
      end subroutine B200_SetMode

!     Before the first B200_Accum.  async: B200_Accum returns once the columns are staged (a level miss is
!     reported by the next call / B200_Wait / B200_Norm instead).  writeback: ob_flag is copied back on every
!     call (m_gritas_grids.f:413-416); only GrITAS_LBTally (-lb, gritas.f:791-796) reads it, so a driver
!     without -lb passes .false. and saves the download and the synchronisation per synoptic time.
      subroutine B200_SetFlags ( async, writeback )
      logical, intent(in) :: async, writeback
      flags = 0
      if ( async )           flags = flags + FLAG_ASYNC
      if ( .not. writeback ) flags = flags + FLAG_NO_WRITEBACK
      end subroutine B200_SetFlags

      subroutine ensure_ ( im, jm, km, nlevs, l2d, l3d, nr, p1, p2, kxkt_table, extra )
      use m_odsmeta, only : KTMAX, KXMAX
      integer, intent(in)        :: im, jm, km, nlevs, l2d, l3d, nr
      real(c_double), intent(in) :: p1(*), p2(*)
      integer(c_int), intent(in) :: kxkt_table(*)
      integer(c_int), intent(in) :: extra
      integer(c_int) :: rc
      if ( c_associated(handle) ) return                                ! first call: gritas.f:336-363
      rc = gritas_b200_create ( handle, im, jm, km, nlevs, l2d, l3d, nr, p1, p2,                &
                                kxkt_table, KXMAX, KTMAX, mode + flags + extra, -1_c_int )
      if ( rc /= 0 ) call gritas_perror ('Accum: cannot create the B200 context')
      if ( allocated(pending_mask) ) then                               ! B200_SetMask came first (gritas.f:745)
         rc = gritas_b200_set_mask ( handle, pending_mask )
         if ( rc /= 0 ) call gritas_perror ('SetMask: B200 upload failed')
         deallocate ( pending_mask )
      end if
      end subroutine ensure_

!     grisas.f:285-296: the caller zeroed its grids since the last Norm -> zero the device accumulators too.
      subroutine auto_reset_ ( nobs_2d, nobs_3d )
      real(c_double), intent(in) :: nobs_2d(*), nobs_3d(*)
      integer(c_int) :: rc
      logical :: zeroed
      if ( .not. normed ) return
      normed = .false.
      zeroed = .false.
      if ( sent3 > 0 ) then
         zeroed = ( nobs_3d(sent3) == 0.0_c_double )
      else if ( sent2 > 0 ) then
         zeroed = ( nobs_2d(sent2) == 0.0_c_double )
      end if
      if ( zeroed .and. c_associated(handle) ) rc = gritas_b200_reset ( handle )
      end subroutine auto_reset_

!     GrITAS_Accum_ (m_gritas_grids.f:307-311) plus the four module variables of m_gritas_grids.  Only the count
!     grids are looked at (auto-reset); no host grid is written before B200_Norm.
      subroutine B200_Accum ( im, jm, km, nlevs, lat, lon, lev, kx, kt, del, nobs, nr,          &
                              kxkt_table, p1, p2, ob_flag, l2d, nobs_2d, l3d, nobs_3d )
      integer, intent(in)           :: im, jm, km, nlevs, nobs, nr, l2d, l3d
      real(c_double), intent(in)    :: lat(*), lon(*), lev(*), del(*)
      integer(c_int), intent(in)    :: kx(*), kt(*), kxkt_table(*)
      real(c_double), intent(in)    :: p1(*), p2(*)
      integer(c_int), intent(inout) :: ob_flag(*)
      real(c_double), intent(in)    :: nobs_2d(*), nobs_3d(*)

      integer(c_int) :: rc
      real(c_double) :: bad_lev

      if ( nr > 3 ) call gritas_perror ('Accum: could not handle '                             &
                                     // 'more then 2 residuals')        ! m_gritas_grids.f:395
      call ensure_ ( im, jm, km, nlevs, l2d, l3d, nr, p1, p2, kxkt_table, 0_c_int )
      call auto_reset_ ( nobs_2d, nobs_3d )
      rc = gritas_b200_accum ( handle, nobs, lat, lon, lev, del, kx, kt, ob_flag, bad_lev )
      call check_accum_ ( rc, bad_lev )
      end subroutine B200_Accum

      subroutine check_accum_ ( rc, bad_lev )
      integer(c_int), intent(in) :: rc
      real(c_double), intent(in) :: bad_lev
      if ( rc == ERR_LEVEL ) then                                       ! m_gritas_grids.f:458-459
         print *, 'Accum: obs level is ', bad_lev, ' hPa'
         call gritas_perror ('Accum: could not find level')
      else if ( rc /= 0 ) then
         call gritas_perror ('Accum: B200 accumulate failed')
      end if
      end subroutine check_accum_

!     The fused entry: residual selection (gritas.f:562-587), QC / time / longitude filter (gritas.f:710-739) and
!     GrITAS_Accum_ in one kernel, straight from the ODS columns.  Replaces gritas.f:562-587 + 710-739 + 777-781 for
!     iopt 0..3 (and 101: omf and oma in one pass, nr = 2).  X_PASSIVE / X_PRE_BAD come from m_odsmeta.
      subroutine B200_Accum_ODS ( im, jm, km, nlevs, nobs, lat, lon, lev, time, kt, kx, qcexcl, &
                                  omf, oma, obs, xvec, iopt, scale_res, ioptn, t1, t2,          &
                                  lona, lonb, nr, kxkt_table, p1, p2, ob_flag, l2d, l3d )
      use m_odsmeta, only : X_PASSIVE, X_PRE_BAD
      integer, intent(in)           :: im, jm, km, nlevs, nobs, nr, l2d, l3d
      real(c_double), intent(in)    :: lat(*), lon(*), lev(*), omf(*), oma(*), obs(*), xvec(*)
      integer(c_int), intent(in)    :: time(*), kt(*), kx(*), qcexcl(*), kxkt_table(*)
      integer, intent(in)           :: iopt, scale_res, ioptn, t1, t2
      real(c_double), intent(in)    :: lona, lonb, p1(*), p2(*)
      integer(c_int), intent(inout) :: ob_flag(*)
      integer(c_int) :: rc
      real(c_double) :: bad_lev
      call ensure_ ( im, jm, km, nlevs, l2d, l3d, nr, p1, p2, kxkt_table, 0_c_int )
      normed = .false.
      rc = gritas_b200_accum_ods ( handle, nobs, lat, lon, lev, time, kt, kx, qcexcl, omf, oma, obs, xvec,   &
                                   iopt, scale_res, ioptn, t1, t2, lona, lonb, X_PASSIVE, X_PRE_BAD, ob_flag, bad_lev )
      call check_accum_ ( rc, bad_lev )
      end subroutine B200_Accum_ODS

!     what GrITAS_Norm_ makes of grids nothing was ever added to (m_gritas_grids.f:584-588, 597-600, 646-647)
      subroutine norm_empty_ ( im, jm, km, nlevs, nr, l2d, bias_2d, rtms_2d, stdv_2d, xcov_2d, nobs_2d,      &
                               l3d, bias_3d, rtms_3d, stdv_3d, xcov_3d, nobs_3d )
      integer, intent(in) :: im, jm, km, nlevs, nr, l2d, l3d
      real(c_double), intent(inout) :: bias_2d(*), rtms_2d(*), stdv_2d(*), xcov_2d(*), nobs_2d(*)
      real(c_double), intent(inout) :: bias_3d(*), rtms_3d(*), stdv_3d(*), xcov_3d(*), nobs_3d(*)
      integer(c_size_t) :: n2, n3, i
      n2 = int(im, c_size_t) * jm * l2d
      n3 = int(im, c_size_t) * jm * km * l3d
      do i = 1, n2 * nr
         bias_2d(i) = AMISS ; rtms_2d(i) = AMISS ; stdv_2d(i) = AMISS ; xcov_2d(i) = AMISS
      end do
      do i = 1, n2
         nobs_2d(i) = 0.0_c_double
      end do
      if ( nlevs < 2 ) return                                           ! m_gritas_grids.f:607
      do i = 1, n3 * nr
         bias_3d(i) = AMISS ; rtms_3d(i) = AMISS ; stdv_3d(i) = AMISS ; xcov_3d(i) = AMISS
      end do
      do i = 1, n3
         nobs_3d(i) = 0.0_c_double
      end do
      end subroutine norm_empty_

!     remember one non-zero count of the grids just written (auto-reset, see B200_Accum)
      subroutine remember_ ( n2, nobs_2d, n3, nobs_3d )
      integer(c_size_t), intent(in) :: n2, n3
      real(c_double), intent(in)    :: nobs_2d(*), nobs_3d(*)
      integer(c_size_t) :: i
      sent2 = 0
      sent3 = 0
      do i = 1, n3
         if ( nobs_3d(i) /= 0.0_c_double ) then
            sent3 = i
            exit
         end if
      end do
      if ( sent3 == 0 ) then
         do i = 1, n2
            if ( nobs_2d(i) /= 0.0_c_double ) then
               sent2 = i
               exit
            end if
         end do
      end if
      normed = .true.
      end subroutine remember_

!     GrITAS_Norm_ (m_gritas_grids.f:490-493) plus im, jm, km, nlevs.
      subroutine B200_Norm ( im, jm, km, nlevs, nr, nrmlz,                                      &
                             l2d, bias_2d, rtms_2d, stdv_2d, xcov_2d, nobs_2d,                  &
                             l3d, bias_3d, rtms_3d, stdv_3d, xcov_3d, nobs_3d,                  &
                             ntresh )
      integer, intent(in)           :: im, jm, km, nlevs, nr, l2d, l3d
      logical, intent(in)           :: nrmlz
      integer, OPTIONAL, intent(in) :: ntresh
      real(c_double), intent(inout) :: bias_2d(*), rtms_2d(*), stdv_2d(*), xcov_2d(*), nobs_2d(*)
      real(c_double), intent(inout) :: bias_3d(*), rtms_3d(*), stdv_3d(*), xcov_3d(*), nobs_3d(*)

      integer(c_int) :: rc, inrm, intr

      if ( nr > 2 ) call gritas_perror ('Accum: could not handle more then 2 residuals')   ! :555
      if ( .not. c_associated(handle) ) then                            ! no observation was ever accumulated
         call norm_empty_ ( im, jm, km, nlevs, nr, l2d, bias_2d, rtms_2d, stdv_2d, xcov_2d, nobs_2d,         &
                            l3d, bias_3d, rtms_3d, stdv_3d, xcov_3d, nobs_3d )
         return
      end if
      inrm = 0
      if ( nrmlz ) inrm = 1
      intr = -1                                                         ! "absent"
      if ( present(ntresh) ) intr = ntresh
      rc = gritas_b200_norm ( handle, inrm, intr, bias_2d, rtms_2d, stdv_2d, xcov_2d, nobs_2d,  &
                              bias_3d, rtms_3d, stdv_3d, xcov_3d, nobs_3d )
      if ( rc == ERR_LEVEL ) call gritas_perror ('Accum: could not find level')            ! deferred (async)
      if ( rc /= 0 ) call gritas_perror ('Norm: B200 finalize failed')
      call remember_ ( int(im, c_size_t) * jm * l2d, nobs_2d, int(im, c_size_t) * jm * km * l3d, nobs_3d )
      end subroutine B200_Norm

!     Multi-rank GrITAS_Norm, one process per GPU on one node (the file loop gritas.f:389-393 runs on files
!     f mod nranks).  Every rank finalizes and writes its slab of rows (B200_Slab) of the arrays; pass arrays that
!     live in a segment shared by the ranks and the writer sees the whole grids.  Collective.
      subroutine B200_Norm_Sharded ( nr, nrmlz, bias_2d, rtms_2d, stdv_2d, xcov_2d, nobs_2d,    &
                                     bias_3d, rtms_3d, stdv_3d, xcov_3d, nobs_3d, ntresh )
      integer, intent(in)           :: nr
      logical, intent(in)           :: nrmlz
      integer, OPTIONAL, intent(in) :: ntresh
      real(c_double), intent(inout) :: bias_2d(*), rtms_2d(*), stdv_2d(*), xcov_2d(*), nobs_2d(*)
      real(c_double), intent(inout) :: bias_3d(*), rtms_3d(*), stdv_3d(*), xcov_3d(*), nobs_3d(*)
      integer(c_int) :: rc, inrm, intr
      if ( nr > 2 ) call gritas_perror ('Accum: could not handle more then 2 residuals')   ! :555
      inrm = 0
      if ( nrmlz ) inrm = 1
      intr = -1
      if ( present(ntresh) ) intr = ntresh
      rc = gritas_b200_norm_sharded ( handle, inrm, intr, bias_2d, rtms_2d, stdv_2d, xcov_2d, nobs_2d,       &
                                      bias_3d, rtms_3d, stdv_3d, xcov_3d, nobs_3d )
      if ( rc /= 0 ) call gritas_perror ('Norm: B200 sharded finalize failed')
      end subroutine B200_Norm_Sharded

!     Multi-rank GrITAS_Norm through NCCL: partial sums reduced to `root`, which alone gets the grids.  Collective.
      subroutine B200_Norm_Reduce ( root, nr, nrmlz, bias_2d, rtms_2d, stdv_2d, xcov_2d, nobs_2d, &
                                    bias_3d, rtms_3d, stdv_3d, xcov_3d, nobs_3d, ntresh )
      integer, intent(in)           :: root, nr
      logical, intent(in)           :: nrmlz
      integer, OPTIONAL, intent(in) :: ntresh
      real(c_double), intent(inout) :: bias_2d(*), rtms_2d(*), stdv_2d(*), xcov_2d(*), nobs_2d(*)
      real(c_double), intent(inout) :: bias_3d(*), rtms_3d(*), stdv_3d(*), xcov_3d(*), nobs_3d(*)
      integer(c_int) :: rc, inrm, intr
      if ( nr > 2 ) call gritas_perror ('Accum: could not handle more then 2 residuals')   ! :555
      inrm = 0
      if ( nrmlz ) inrm = 1
      intr = -1
      if ( present(ntresh) ) intr = ntresh
      rc = gritas_b200_norm_reduce ( handle, root, inrm, intr, bias_2d, rtms_2d, stdv_2d, xcov_2d, nobs_2d,  &
                                     bias_3d, rtms_3d, stdv_3d, xcov_3d, nobs_3d )
      if ( rc /= 0 ) call gritas_perror ('Norm: B200 reduce + finalize failed')
      end subroutine B200_Norm_Reduce

!     rank 0 fills the 128-byte NCCL id; broadcast it (MPI_Bcast) and hand it to B200_CommInit on every rank
      subroutine B200_UniqueId ( id128 )
      character(kind=c_char), intent(out) :: id128(128)
      if ( gritas_b200_nccl_unique_id ( id128 ) /= 0 ) call gritas_perror ('UniqueId: NCCL not available')
      end subroutine B200_UniqueId

      subroutine B200_CommInit ( nranks, rank, id128 )
      integer, intent(in) :: nranks, rank
      character(kind=c_char), intent(in) :: id128(128)
      if ( .not. c_associated(handle) ) call gritas_perror ('CommInit: call after the context exists (first Accum)')
      if ( gritas_b200_comm_init ( handle, nranks, rank, id128 ) /= 0 ) call gritas_perror ('CommInit: NCCL init failed')
      end subroutine B200_CommInit

!     this rank's blob; MPI_Allgather the blobs of all ranks (rank order) and hand them to B200_PeerAttach
      subroutine B200_PeerExport ( blob )
      character(kind=c_char), intent(out) :: blob(B200_PEER_BLOB_BYTES)
      if ( .not. c_associated(handle) ) call gritas_perror ('PeerExport: call after the context exists (first Accum)')
      if ( gritas_b200_peer_export ( handle, blob ) /= 0 ) call gritas_perror ('PeerExport: failed')
      end subroutine B200_PeerExport

      subroutine B200_PeerAttach ( nranks, rank, blobs )
      integer, intent(in) :: nranks, rank
      character(kind=c_char), intent(in) :: blobs(*)
      if ( gritas_b200_peer_attach ( handle, nranks, rank, blobs ) /= 0 )                      &
         call gritas_perror ('PeerAttach: ranks differ in their settings, or no peer access')
      end subroutine B200_PeerAttach

!     rows jl = j-1 + jm*(l-1) (0-based, half-open) of the 3-D / 2-D arrays that `rank` writes in B200_Norm_Sharded
      subroutine B200_Slab ( rank, nranks, jl3_0, jl3_1, jl2_0, jl2_1 )
      integer, intent(in)  :: rank, nranks
      integer, intent(out) :: jl3_0, jl3_1, jl2_0, jl2_1
      integer(c_int) :: a, b, c, d
      if ( gritas_b200_slab ( handle, rank, nranks, a, b, c, d ) /= 0 ) call gritas_perror ('Slab: bad arguments')
      jl3_0 = a ; jl3_1 = b ; jl2_0 = c ; jl2_1 = d
      end subroutine B200_Slab

!     Replaces IndexSet + the eight IndexSort calls of gritas.f:500-508:
!         call B200_PreSort ( nobs, is, ods%data%qcexcl, ods%data%lat, ods%data%lon, ods%data%lev,
!                             ods%data%time, ods%data%kt, ods%data%ks, ods%data%kx )
!     `is` comes back 1-based, ready for the gathers of gritas.f:510-523.  Works before the first B200_Accum.
      subroutine B200_PreSort ( nobs, is, qcexcl, lat, lon, lev, time, kt, ks, kx )
      integer, intent(in)         :: nobs
      integer(c_int), intent(out) :: is(*)
      integer(c_int), intent(in)  :: qcexcl(*), time(*), kt(*), ks(*), kx(*)
      real(c_double), intent(in)  :: lat(*), lon(*), lev(*)
      integer(c_int) :: rc
      rc = gritas_b200_presort ( handle, nobs, qcexcl, lat, lon, lev, time, kt, ks, kx, 1_c_int, is )
      if ( rc /= 0 ) call gritas_perror ('PreSort: B200 sort failed')
      end subroutine B200_PreSort

!     GrITAS_3CH_Norm_ (m_gritas_grids.f:1178-1181; gritas.f:813-818, -tch).
      subroutine B200_3CH_Norm ( nr, nrmlz,                                                     &
                                 bias_2d, rtms_2d, var_2d, xcov_2d, nobs_2d,                    &
                                 bias_3d, rtms_3d, var_3d, xcov_3d, nobs_3d,                    &
                                 ntresh )
      integer, intent(in)           :: nr
      logical, intent(in)           :: nrmlz
      integer, OPTIONAL, intent(in) :: ntresh
      real(c_double), intent(inout) :: bias_2d(*), rtms_2d(*), var_2d(*), xcov_2d(*), nobs_2d(*)
      real(c_double), intent(inout) :: bias_3d(*), rtms_3d(*), var_3d(*), xcov_3d(*), nobs_3d(*)

      integer(c_int) :: rc, inrm, intr

      if ( nr > 3 ) call gritas_perror ('GrITAS_3CH_Norm_ could not handle more then 2 residuals')   ! :1240
      if ( .not. c_associated(handle) ) call gritas_perror ('3CH_Norm: nothing was accumulated')
      inrm = 0
      if ( nrmlz ) inrm = 1
      intr = -1                                                         ! "absent"
      if ( present(ntresh) ) intr = ntresh
      rc = gritas_b200_norm_3ch ( handle, inrm, intr, bias_2d, rtms_2d, var_2d, xcov_2d, nobs_2d, &
                                  bias_3d, rtms_3d, var_3d, xcov_3d, nobs_3d )
      if ( rc /= 0 ) call gritas_perror ('3CH_Norm: B200 finalize failed')
      end subroutine B200_3CH_Norm

!     GrITAS_MinMax_ (m_gritas_grids.f:1019-1023) plus im, jm, km, nlevs.  Unlike Accum/Norm there is no
!     separate finalize call in the driver (gritas.f:884 skips Norm with -minmax), so the min / max / count
!     grids are downloaded after every call; drivers that process many files may call the C entry
!     gritas_b200_get_minmax once after the file loop instead.
      subroutine B200_MinMax ( im, jm, km, nlevs, lat, lon, lev, kx, kt, del, nobs, nr,         &
                               kxkt_table, p1, p2, ob_flag,                                     &
                               l2d, minv_2d, maxv_2d, nobs_2d,                                  &
                               l3d, minv_3d, maxv_3d, nobs_3d )
      integer, intent(in)           :: im, jm, km, nlevs, nobs, nr, l2d, l3d
      real(c_double), intent(in)    :: lat(*), lon(*), lev(*), del(*)
      integer(c_int), intent(in)    :: kx(*), kt(*), kxkt_table(*)
      real(c_double), intent(in)    :: p1(*), p2(*)
      integer(c_int), intent(inout) :: ob_flag(*)
      real(c_double), intent(inout) :: minv_2d(*), maxv_2d(*), nobs_2d(*)
      real(c_double), intent(inout) :: minv_3d(*), maxv_3d(*), nobs_3d(*)
      integer(c_int) :: rc
      real(c_double) :: bad_lev
      if ( nr > 1 ) call gritas_perror ('MinMax: could not handle more then 1 field')      ! :1071-1072
      call ensure_ ( im, jm, km, nlevs, l2d, l3d, nr, p1, p2, kxkt_table, FLAG_MINMAX )
      rc = gritas_b200_accum ( handle, nobs, lat, lon, lev, del, kx, kt, ob_flag, bad_lev )
      call check_accum_ ( rc, bad_lev )                                  ! m_gritas_grids.f:1132-1135
      rc = gritas_b200_get_minmax ( handle, minv_2d, maxv_2d, nobs_2d, minv_3d, maxv_3d, nobs_3d )
      if ( rc /= 0 ) call gritas_perror ('MinMax: B200 download failed')
      end subroutine B200_MinMax

!     gritas_maskout (m_gritas_masks.F90:137-186): hand the (im,jm) fraction field of ini_masks_ to the library;
!     every later Accum flags the observations whose box holds less than 0.99 (replaces gritas.f:745).  May be
!     called before the first B200_Accum: the field is kept until the context exists.
      subroutine B200_SetMask ( im, jm, maskfld )
      integer, intent(in)        :: im, jm
      real(c_double), intent(in) :: maskfld(*)
      integer(c_int) :: rc
      if ( .not. c_associated(handle) ) then
         if ( allocated(pending_mask) ) deallocate ( pending_mask )
         allocate ( pending_mask(im*jm) )
         pending_mask(1:im*jm) = maskfld(1:im*jm)
         return
      end if
      rc = gritas_b200_set_mask ( handle, maskfld )
      if ( rc /= 0 ) call gritas_perror ('SetMask: B200 upload failed')
      end subroutine B200_SetMask

!     Un-normalised sums into the host grids (what they would hold between Accum and Norm).
      subroutine B200_GetRaw ( bias_2d, rtms_2d, xcov_2d, nobs_2d,                              &
                               bias_3d, rtms_3d, xcov_3d, nobs_3d )
      real(c_double), intent(inout) :: bias_2d(*), rtms_2d(*), xcov_2d(*), nobs_2d(*)
      real(c_double), intent(inout) :: bias_3d(*), rtms_3d(*), xcov_3d(*), nobs_3d(*)
      integer(c_int) :: rc
      if ( .not. c_associated(handle) ) return                          ! nothing accumulated: the grids are as the caller left them
      rc = gritas_b200_get_raw ( handle, bias_2d, rtms_2d, xcov_2d, nobs_2d,                    &
                                 bias_3d, rtms_3d, xcov_3d, nobs_3d )
      if ( rc /= 0 ) call gritas_perror ('GetRaw: B200 download failed')
      end subroutine B200_GetRaw

!     grisas.f:285-296 zeroes the grids at every synoptic time.
      subroutine B200_Reset ( )
      integer(c_int) :: rc
      normed = .false.
      if ( c_associated(handle) ) rc = gritas_b200_reset ( handle )
      end subroutine B200_Reset

!     With B200_SetFlags(async=.true.): wait for the queued work and pick up a deferred level miss.
      subroutine B200_Wait ( )
      integer(c_int) :: rc
      real(c_double) :: bad_lev
      if ( .not. c_associated(handle) ) return
      rc = gritas_b200_wait ( handle, bad_lev )
      call check_accum_ ( rc, bad_lev )
      end subroutine B200_Wait

!     GrITAS_XCov_Accum_ (m_gritas_grids.f:667-670) plus the module variables km, nchan and the achan(:,ichan_version) column
!     (m_gritas_grids.f:26, 28, 69-71).  lat / lon are arguments of the reference routine that it never reads.
      subroutine B200_XCov_Accum ( km, nchan, achan, lev, kx, kt, del, nobs, nr, kxkt_table, ob_flag, l3d, nprof, tch )
      use m_odsmeta, only : KTMAX, KXMAX
      integer, intent(in)           :: km, nchan, nobs, nr, l3d
      integer(c_int), intent(in)    :: achan(*), kx(*), kt(*), kxkt_table(*)
      real(c_double), intent(in)    :: lev(*), del(*)
      integer(c_int), intent(inout) :: ob_flag(*)
      integer, intent(out)          :: nprof
      logical, intent(in)           :: tch
      integer(c_int) :: rc, itch, np
      itch = 0
      if ( tch ) itch = 1
      if ( .not. c_associated(xhandle) ) then
         rc = gritas_b200_xcov_create ( xhandle, km, nchan, achan, l3d, nr, itch, kxkt_table, KXMAX, KTMAX, -1_c_int )
         if ( rc /= 0 ) call gritas_perror ('XCov_Accum: improper dim settings / cannot create the B200 context')   ! :734-745
      end if
      rc = gritas_b200_xcov_accum ( xhandle, nobs, lev, kx, kt, del, ob_flag, np )
      if ( rc /= 0 ) call gritas_perror ('XCov_Accum: not all profiles complete, aborting ...')                   ! :754-757
      nprof = np
      print *, 'GrITAS_XCov_Accum, total profiles processed = ', nprof                                              ! :840
      end subroutine B200_XCov_Accum

!     GrITAS_XCov_Prs_Accum_ (m_gritas_grids.f:1862-1865) plus nlevs; kx / kt / kxkt_table / ob_flag are unused by the reference.
      subroutine B200_XCov_Prs_Accum ( nlevs, lev, ks, del, nobs, nr, p1, p2, l3d, nprof, tch )
      integer, intent(in)        :: nlevs, nobs, nr, l3d
      real(c_double), intent(in) :: lev(*), del(*), p1(*), p2(*)
      integer(c_int), intent(in) :: ks(*)
      integer, intent(out)       :: nprof
      logical, intent(in)        :: tch
      integer(c_int) :: rc, itch, np
      real(c_double) :: bad_lev
      itch = 0
      if ( tch ) itch = 1
      if ( .not. c_associated(xhandle) ) then
         rc = gritas_b200_xcov_prs_create ( xhandle, nlevs, p1, p2, l3d, nr, itch, -1_c_int )
         if ( rc /= 0 ) call gritas_perror ('XCov_Prs_Accum: improper dim settings / cannot create the B200 context')
      end if
      rc = gritas_b200_xcov_prs_accum ( xhandle, nobs, lev, ks, del, np, bad_lev )
      if ( rc == ERR_LEVEL ) call gritas_perror ('Accum: could not find level(1)')                                 ! :1990
      if ( rc /= 0 ) call gritas_perror ('XCov_Prs_Accum: B200 accumulate failed')
      nprof = np
      end subroutine B200_XCov_Prs_Accum

!     GrITAS_XCov_Norm_ (m_gritas_grids.f:859-861).
      subroutine B200_XCov_Norm ( nr, nrmlz, tch, self, bias_3d, xcov_3d, nobs_3d, ntresh )
      integer, intent(in)           :: nr
      logical, intent(in)           :: nrmlz, tch, self
      integer, OPTIONAL, intent(in) :: ntresh
      real(c_double), intent(inout) :: bias_3d(*), xcov_3d(*), nobs_3d(*)
      integer(c_int) :: rc, inrm, iself, intr
      if ( nr > 3 ) call gritas_perror ('GrITAS_XCov_Norm_: improper dim settings')                               ! :907-909
      if ( .not. c_associated(xhandle) ) return                         ! nothing accumulated: the arrays stay as they are
      inrm = 0
      if ( nrmlz ) inrm = 1
      iself = 0
      if ( self ) iself = 1
      intr = -1
      if ( present(ntresh) ) intr = ntresh
      rc = gritas_b200_xcov_norm ( xhandle, inrm, iself, intr, bias_3d, xcov_3d, nobs_3d )
      if ( rc /= 0 ) call gritas_perror ('XCov_Norm: B200 finalize failed')
      end subroutine B200_XCov_Norm

      subroutine B200_XCov_Clean ( )
      if ( c_associated(xhandle) ) call gritas_b200_xcov_destroy ( xhandle )
      xhandle = c_null_ptr
      end subroutine B200_XCov_Clean

!     GrITAS_Clean (gritas.f:2017-2026).
      subroutine B200_Clean ( )
      if ( c_associated(handle) ) call gritas_b200_destroy ( handle )
      handle = c_null_ptr
      normed = .false.
      if ( allocated(pending_mask) ) deallocate ( pending_mask )
      end subroutine B200_Clean

      end module m_gritas_b200
