!  m_gritas_b200.F90 -- ISO_C_BINDING shim between the unchanged GrITAS Fortran driver and
!  libgritas_b200.so (include/gritas_b200.h).
!
!  NOT COMPILED IN THIS REPOSITORY'S CI: the build image has no Fortran compiler.  The module is
!  deliberately mechanical -- every procedure forwards its arguments one-to-one to the C entry
!  point that tests/ exercise through ctypes (gritas_b200/cabi.py binds the same symbols with the
!  same argument order).
!
!  Drop-in use: the procedures below have exactly the argument lists of GrITAS_Accum_ and
!  GrITAS_Norm_ in m_gritas_grids.f (lines 307-311 and 490-493).  In m_gritas_grids.f replace the
!  bodies of those two routines by a call to the routines here (or point the generics at them):
!
!      Interface GrITAS_Accum
!         module procedure B200_Accum_      ! was GrITAS_Accum_   (m_gritas_grids.f:88)
!      end Interface
!      Interface GrITAS_Norm
!         module procedure B200_Norm_       ! was GrITAS_Norm_    (m_gritas_grids.f:89)
!      end Interface
!
!  gritas.f, grisas.f, the rc/namelist handling and the GrADS/NetCDF writers stay untouched.
!  Semantics that differ from the host implementation (see INTEGRATION.md):
!    * the accumulators live on the GPU; the host grids passed to B200_Accum_ are not modified
!      until B200_Norm_ (or B200_GetRaw) fills them.  Code that reads the raw sums between the
!      two calls (gritas.f:812-818 -nonorm, the scorecard) must call B200_GetRaw first;
!    * all reals are 64-bit (the library is built for the FREAL8 configuration of
!      src/Components/gritas/CMakeLists.txt:3-4).

      module m_gritas_b200

      use, intrinsic :: iso_c_binding
      implicit none
      private

      public :: B200_Accum_, B200_Norm_, B200_3CH_Norm_, B200_GetRaw, B200_Reset, B200_Clean, B200_PreSort
      public :: B200_MinMax_, B200_SetMask
      public :: B200_SetMode

      integer(c_int), parameter :: MODE_FAST = 0, MODE_DETERMINISTIC = 1, FLAG_MINMAX = 8192
      integer(c_int), parameter :: ERR_LEVEL = 7

      type(c_ptr), save      :: handle = c_null_ptr
      integer(c_int), save   :: mode   = MODE_FAST

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
            integer(c_int), value             :: nobs, index_base
            integer(c_int), intent(in)        :: qcexcl(*), time(*), kt(*), ks(*), kx(*)
            real(c_double), intent(in)        :: lat(*), lon(*), lev(*)
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
      end interface

      contains

!     Select the accumulation mode before the first B200_Accum_ (0 fast, 1 deterministic).
      subroutine B200_SetMode ( m )
      integer, intent(in) :: m
      mode = int(m, c_int)
      end subroutine B200_SetMode

!     Same argument list as GrITAS_Accum_ (m_gritas_grids.f:307-311).  im, jm, km, nlevs are the
!     module variables of m_gritas_grids (m_gritas_grids.f:24-28); KXMAX, KTMAX come from m_odsmeta.
      subroutine B200_Accum_ ( lat, lon, lev, kx, kt, del, nobs, nr,                            &
                               kxkt_table, p1, p2,                                              &
                               ob_flag,                                                         &
                               l2d, bias_2d, rtms_2d, xcov_2d, nobs_2d,                         &
                               l3d, bias_3d, rtms_3d, xcov_3d, nobs_3d )
      use m_gritas_grids, only : im, jm, km, nlevs
      use m_odsmeta,      only : KTMAX, KXMAX
      integer, intent(in)    :: nobs, nr
      real,    intent(in)    :: lat(nobs), lon(nobs), lev(nobs), del(nobs,nr)
      integer, intent(in)    :: kx(nobs), kt(nobs)
      integer, intent(in)    :: kxkt_table(KXMAX,KTMAX)
      real,    intent(in)    :: p1(nlevs), p2(nlevs)
      integer, intent(in)    :: l2d, l3d
      integer, intent(inout) :: ob_flag(nobs)
      real,    intent(inout) :: bias_2d(im,jm,l2d,nr), rtms_2d(im,jm,l2d,nr)
      real,    intent(inout) :: xcov_2d(im,jm,l2d),    nobs_2d(im,jm,l2d)
      real,    intent(inout) :: bias_3d(im,jm,km,l3d,nr), rtms_3d(im,jm,km,l3d,nr)
      real,    intent(inout) :: xcov_3d(im,jm,km,l3d),    nobs_3d(im,jm,km,l3d)

      integer(c_int) :: rc
      real(c_double) :: bad_lev

      if ( nr > 3 ) call gritas_perror ('Accum: could not handle '                             &
                                     // 'more then 2 residuals')        ! m_gritas_grids.f:395
      if ( .not. c_associated(handle) ) then                            ! first call: gritas.f:336-363
         rc = gritas_b200_create ( handle, im, jm, km, nlevs, l2d, l3d, nr, p1, p2,             &
                                   kxkt_table, KXMAX, KTMAX, mode, -1_c_int )
         if ( rc /= 0 ) call gritas_perror ('Accum: cannot create the B200 context')
      end if

      rc = gritas_b200_accum ( handle, nobs, lat, lon, lev, del, kx, kt, ob_flag, bad_lev )
      if ( rc == ERR_LEVEL ) then                                       ! m_gritas_grids.f:458-459
         print *, 'Accum: obs level is ', bad_lev, ' hPa'
         call gritas_perror ('Accum: could not find level')
      else if ( rc /= 0 ) then
         call gritas_perror ('Accum: B200 accumulate failed')
      end if
      end subroutine B200_Accum_

!     Same argument list as GrITAS_Norm_ (m_gritas_grids.f:490-493).
      subroutine B200_Norm_ ( nr, nrmlz,                                                        &
                              l2d, bias_2d, rtms_2d, stdv_2d, xcov_2d, nobs_2d,                 &
                              l3d, bias_3d, rtms_3d, stdv_3d, xcov_3d, nobs_3d,                 &
                              ntresh )
      use m_gritas_grids, only : im, jm, km
      integer, intent(in)           :: nr, l2d, l3d
      logical, intent(in)           :: nrmlz
      integer, OPTIONAL, intent(in) :: ntresh
      real, intent(inout) :: bias_2d(im,jm,l2d,nr), rtms_2d(im,jm,l2d,nr), xcov_2d(im,jm,l2d,nr)
      real, intent(inout) :: stdv_2d(im,jm,l2d,nr), nobs_2d(im,jm,l2d)
      real, intent(inout) :: bias_3d(im,jm,km,l3d,nr), rtms_3d(im,jm,km,l3d,nr), xcov_3d(im,jm,km,l3d,nr)
      real, intent(inout) :: stdv_3d(im,jm,km,l3d,nr), nobs_3d(im,jm,km,l3d)

      integer(c_int) :: rc, inrm, intr

      if ( nr > 2 ) call gritas_perror ('Accum: could not handle more then 2 residuals')   ! :555
      inrm = 0
      if ( nrmlz ) inrm = 1
      intr = -1                                                         ! "absent"
      if ( present(ntresh) ) intr = ntresh
      rc = gritas_b200_norm ( handle, inrm, intr, bias_2d, rtms_2d, stdv_2d, xcov_2d, nobs_2d,  &
                              bias_3d, rtms_3d, stdv_3d, xcov_3d, nobs_3d )
      if ( rc /= 0 ) call gritas_perror ('Norm: B200 finalize failed')
      end subroutine B200_Norm_

!     Replaces IndexSet + the eight IndexSort calls of gritas.f:500-508:
!         call B200_PreSort ( nobs, is, ods%data%qcexcl, ods%data%lat, ods%data%lon, ods%data%lev,
!                             ods%data%time, ods%data%kt, ods%data%ks, ods%data%kx )
!     `is` comes back 1-based, ready for the gathers of gritas.f:510-523.  Works before the first B200_Accum_.
      subroutine B200_PreSort ( nobs, is, qcexcl, lat, lon, lev, time, kt, ks, kx )
      integer, intent(in)  :: nobs
      integer, intent(out) :: is(nobs)
      integer, intent(in)  :: qcexcl(nobs), time(nobs), kt(nobs), ks(nobs), kx(nobs)
      real,    intent(in)  :: lat(nobs), lon(nobs), lev(nobs)
      integer(c_int) :: rc
      rc = gritas_b200_presort ( handle, nobs, qcexcl, lat, lon, lev, time, kt, ks, kx, 1, is )
      if ( rc /= 0 ) call gritas_perror ('PreSort: B200 sort failed')
      end subroutine B200_PreSort

!     Same argument list as GrITAS_3CH_Norm_ (m_gritas_grids.f:1178-1181; gritas.f:813-818, -tch).
      subroutine B200_3CH_Norm_ ( nr, nrmlz,                                                    &
                                  l2d, bias_2d, rtms_2d, var_2d, xcov_2d, nobs_2d,              &
                                  l3d, bias_3d, rtms_3d, var_3d, xcov_3d, nobs_3d,              &
                                  ntresh )
      use m_gritas_grids, only : im, jm, km
      integer, intent(in)           :: nr, l2d, l3d
      logical, intent(in)           :: nrmlz
      integer, OPTIONAL, intent(in) :: ntresh
      real, intent(inout) :: bias_2d(im,jm,l2d,nr), rtms_2d(im,jm,l2d,nr), xcov_2d(im,jm,l2d,nr)
      real, intent(inout) :: var_2d(im,jm,l2d,nr), nobs_2d(im,jm,l2d)
      real, intent(inout) :: bias_3d(im,jm,km,l3d,nr), rtms_3d(im,jm,km,l3d,nr), xcov_3d(im,jm,km,l3d,nr)
      real, intent(inout) :: var_3d(im,jm,km,l3d,nr), nobs_3d(im,jm,km,l3d)

      integer(c_int) :: rc, inrm, intr

      if ( nr > 3 ) call gritas_perror ('GrITAS_3CH_Norm_ could not handle more then 2 residuals')   ! :1240
      inrm = 0
      if ( nrmlz ) inrm = 1
      intr = -1                                                         ! "absent"
      if ( present(ntresh) ) intr = ntresh
      rc = gritas_b200_norm_3ch ( handle, inrm, intr, bias_2d, rtms_2d, var_2d, xcov_2d, nobs_2d, &
                                  bias_3d, rtms_3d, var_3d, xcov_3d, nobs_3d )
      if ( rc /= 0 ) call gritas_perror ('3CH_Norm: B200 finalize failed')
      end subroutine B200_3CH_Norm_

!     Same argument list as GrITAS_MinMax_ (m_gritas_grids.f:1019-1023).  Unlike Accum/Norm there is no
!     separate finalize call in the driver (gritas.f:884 skips Norm with -minmax), so the min / max / count
!     grids are downloaded after every call; drivers that process many files may call the C entry
!     gritas_b200_get_minmax once after the file loop instead.
      subroutine B200_MinMax_ ( lat, lon, lev, kx, kt, del, nobs, nr,                           &
                                kxkt_table, p1, p2, ob_flag,                                    &
                                l2d, minv_2d, maxv_2d, nobs_2d,                                 &
                                l3d, minv_3d, maxv_3d, nobs_3d )
      use m_gritas_grids, only : im, jm, km, nlevs
      use m_odsmeta,      only : KTMAX, KXMAX
      integer, intent(in)    :: nobs, nr
      real,    intent(in)    :: lat(nobs), lon(nobs), lev(nobs), del(nobs,nr)
      integer, intent(in)    :: kx(nobs), kt(nobs)
      integer, intent(in)    :: kxkt_table(KXMAX,KTMAX)
      real,    intent(in)    :: p1(nlevs), p2(nlevs)
      integer, intent(in)    :: l2d, l3d
      integer, intent(inout) :: ob_flag(nobs)
      real,    intent(inout) :: minv_2d(im,jm,l2d,nr), maxv_2d(im,jm,l2d,nr), nobs_2d(im,jm,l2d)
      real,    intent(inout) :: minv_3d(im,jm,km,l3d,nr), maxv_3d(im,jm,km,l3d,nr), nobs_3d(im,jm,km,l3d)
      integer(c_int) :: rc
      real(c_double) :: bad_lev
      if ( nr > 1 ) call gritas_perror ('MinMax: could not handle more then 1 field')      ! :1071-1072
      if ( .not. c_associated(handle) ) then
         rc = gritas_b200_create ( handle, im, jm, km, nlevs, l2d, l3d, nr, p1, p2,             &
                                   kxkt_table, KXMAX, KTMAX, mode + FLAG_MINMAX, -1_c_int )
         if ( rc /= 0 ) call gritas_perror ('MinMax: cannot create the B200 context')
      end if
      rc = gritas_b200_accum ( handle, nobs, lat, lon, lev, del, kx, kt, ob_flag, bad_lev )
      if ( rc == ERR_LEVEL ) then                                        ! m_gritas_grids.f:1132-1135
         print *, 'Accum: obs level is ', bad_lev, ' hPa'
         call gritas_perror ('Accum: could not find level')
      else if ( rc /= 0 ) then
         call gritas_perror ('MinMax: B200 accumulate failed')
      end if
      rc = gritas_b200_get_minmax ( handle, minv_2d, maxv_2d, nobs_2d, minv_3d, maxv_3d, nobs_3d )
      if ( rc /= 0 ) call gritas_perror ('MinMax: B200 download failed')
      end subroutine B200_MinMax_

!     gritas_maskout (m_gritas_masks.F90:137-186): hand the (im,jm) fraction field of ini_masks_ to the library
!     once; every later Accum flags the observations whose box holds less than 0.99 (replaces gritas.f:745).
      subroutine B200_SetMask ( maskfld )
      real, intent(in) :: maskfld(*)
      integer(c_int) :: rc
      if ( .not. c_associated(handle) ) call gritas_perror ('SetMask: call after the first Accum')
      rc = gritas_b200_set_mask ( handle, maskfld )
      if ( rc /= 0 ) call gritas_perror ('SetMask: B200 upload failed')
      end subroutine B200_SetMask

!     Un-normalised sums into the host grids (what they would hold between Accum and Norm).
      subroutine B200_GetRaw ( bias_2d, rtms_2d, xcov_2d, nobs_2d,                              &
                               bias_3d, rtms_3d, xcov_3d, nobs_3d )
      real, intent(inout) :: bias_2d(*), rtms_2d(*), xcov_2d(*), nobs_2d(*)
      real, intent(inout) :: bias_3d(*), rtms_3d(*), xcov_3d(*), nobs_3d(*)
      integer(c_int) :: rc
      rc = gritas_b200_get_raw ( handle, bias_2d, rtms_2d, xcov_2d, nobs_2d,                    &
                                 bias_3d, rtms_3d, xcov_3d, nobs_3d )
      if ( rc /= 0 ) call gritas_perror ('GetRaw: B200 download failed')
      end subroutine B200_GetRaw

!     grisas.f:285-296 zeroes the grids at every synoptic time.
      subroutine B200_Reset ( )
      integer(c_int) :: rc
      if ( c_associated(handle) ) rc = gritas_b200_reset ( handle )
      end subroutine B200_Reset

!     GrITAS_Clean (gritas.f:2017-2026).
      subroutine B200_Clean ( )
      if ( c_associated(handle) ) call gritas_b200_destroy ( handle )
      handle = c_null_ptr
      end subroutine B200_Clean

      end module m_gritas_b200
