! specfab_b200.f90 -- iso_c_binding shim: Fortran / f2py / Elmer callers -> libspecfab_b200.so
!
! Shipped as SOURCE (no Fortran compiler exists in the build image, DESIGN.md section 6).  A maintainer adds
! this file to src/Makefile next to specfab.f90 and links with -lspecfab_b200.  The batched procedures keep
! the reference's own *_arr convention (leading node dimension, cf. Eij_tranisotropic_arr,
! src/specfabpy.f90:474-486); Fortran's column-major order IS the library layout, so arrays are passed as they are.
! Non-zero return codes are turned into `stop`, which is what the reference does on error
! (src/homogenizations.f90:183, src/frames.f90:48).
module specfab_b200
    use iso_c_binding
    implicit none
    integer, parameter, private :: dp = 8

    integer(c_int), parameter :: SFB_LROT = 1, SFB_DDRX = 2, SFB_CDRX = 4, SFB_REG = 8
    integer(c_int), parameter :: SFB_EULER = 1, SFB_RK4 = 4

    type, bind(c) :: sfb_step_opts
        real(c_double) :: dt, iota, zeta, nu_mult, gamma0, lambda
        type(c_ptr)    :: gamma0_arr, lambda_arr
        integer(c_int32_t) :: terms, scheme, nsteps, reserved
    end type

    interface
        integer(c_int) function sfb_init(L) bind(c, name='sfb_init')                     ! src/specfabpy.f90:150
            import; integer(c_int), value :: L
        end function
        integer(c_int) function sfb_nlm_len() bind(c, name='sfb_nlm_len')
            import
        end function
        integer(c_int) function sfb_step_arr(nlm_in, nlm_out, N, ld, ugrad, tau, opts) bind(c, name='sfb_step_arr')
            import; type(c_ptr), value :: nlm_in, nlm_out, ugrad, tau
            integer(c_int64_t), value :: N, ld
            type(sfb_step_opts), intent(in) :: opts
        end function
        ! one host array sharded over several GPUs (contiguous node ranges, one host thread + staging ring per device)
        integer(c_int) function sfb_step_arr_multi(nlm_in, nlm_out, N, ld, ugrad, tau, opts, devices, ndev) bind(c, name='sfb_step_arr_multi')
            import; type(c_ptr), value :: nlm_in, nlm_out, ugrad, tau
            integer(c_int64_t), value :: N, ld
            type(sfb_step_opts), intent(in) :: opts
            integer(c_int), intent(in) :: devices(*)
            integer(c_int), value :: ndev
        end function
        ! page-lock an existing array once (4x faster host-pointer calls), release before deallocation
        integer(c_int) function sfb_host_register(p, bytes) bind(c, name='sfb_host_register')
            import; type(c_ptr), value :: p; integer(c_int64_t), value :: bytes
        end function
        integer(c_int) function sfb_host_unregister(p) bind(c, name='sfb_host_unregister')
            import; type(c_ptr), value :: p
        end function
        ! the same step on reduced-form states rnlm(N, rnlm_len) (src/reducedform.f90:160-187)
        integer(c_int) function sfb_step_rnlm_arr(rnlm_in, rnlm_out, N, ld, ugrad, tau, opts) bind(c, name='sfb_step_rnlm_arr')
            import; type(c_ptr), value :: rnlm_in, rnlm_out, ugrad, tau
            integer(c_int64_t), value :: N, ld
            type(sfb_step_opts), intent(in) :: opts
        end function
        integer(c_int) function sfb_a2_arr(nlm, N, ld, a2) bind(c, name='sfb_a2_arr')      ! src/specfabpy.f90:583
            import; type(c_ptr), value :: nlm, a2; integer(c_int64_t), value :: N, ld
        end function
        integer(c_int) function sfb_a4_arr(nlm, N, ld, a4) bind(c, name='sfb_a4_arr')      ! src/specfabpy.f90:592
            import; type(c_ptr), value :: nlm, a4; integer(c_int64_t), value :: N, ld
        end function
        integer(c_int) function sfb_eig_arr(nlm, N, ld, ei, lami) bind(c, name='sfb_eig_arr')   ! src/specfabpy.f90:312
            import; type(c_ptr), value :: nlm, ei, lami; integer(c_int64_t), value :: N, ld
        end function
        integer(c_int) function sfb_eigframe_arr(M, N, plane, ei, lami) bind(c, name='sfb_eigframe_arr')  ! src/specfabpy.f90:333
            import; type(c_ptr), value :: M, ei, lami; integer(c_int64_t), value :: N
            character(kind=c_char), intent(in) :: plane(*)
        end function
        integer(c_int) function sfb_Eij_tranisotropic_arr(nlm, N, ld, e1, e2, e3, Eij_grain, alpha, n_grain, Eij, status) &
                bind(c, name='sfb_Eij_tranisotropic_arr')                                 ! src/specfabpy.f90:474
            import; type(c_ptr), value :: nlm, e1, e2, e3, Eij, status
            integer(c_int64_t), value :: N, ld
            real(c_double), intent(in) :: Eij_grain(2)
            real(c_double), value :: alpha
            integer(c_int), value :: n_grain
        end function
        integer(c_int) function sfb_Eij_orthotropic_arr(nlm_1, nlm_2, nlm_3, N, ld, e1, e2, e3, Eij_grain, alpha, n_grain, Eij) &
                bind(c, name='sfb_Eij_orthotropic_arr')                                   ! src/specfabpy.f90:488
            import; type(c_ptr), value :: nlm_1, nlm_2, nlm_3, e1, e2, e3, Eij
            integer(c_int64_t), value :: N, ld
            real(c_double), intent(in) :: Eij_grain(6)
            real(c_double), value :: alpha
            integer(c_int), value :: n_grain
        end function
        integer(c_int) function sfb_E_CAFFE_arr(nlm, N, ld, eps, Emin, Emax, n_grain, E) bind(c, name='sfb_E_CAFFE_arr')  ! src/specfabpy.f90:543
            import; type(c_ptr), value :: nlm, eps, E; integer(c_int64_t), value :: N, ld
            real(c_double), value :: Emin, Emax; integer(c_int), value :: n_grain
        end function
        integer(c_int) function sfb_M_LROT_reduced_arr(eps, omg, N, iota, zeta, Mrr, Mri, Mir, Mii) &
                bind(c, name='sfb_M_LROT_reduced_arr')                                    ! src/reducedform.f90:76 of src/dynamics.f90:52
            import; type(c_ptr), value :: eps, omg, Mrr, Mri, Mir, Mii; integer(c_int64_t), value :: N
            real(c_double), value :: iota, zeta
        end function
        integer(c_int) function sfb_M_DDRX_reduced_arr(nlm, ld_nlm, tau, N, src_only, Mrr, Mri, Mir, Mii) &
                bind(c, name='sfb_M_DDRX_reduced_arr')                                    ! src/reducedform.f90:76 of src/dynamics.f90:251
            import; type(c_ptr), value :: nlm, tau, Mrr, Mri, Mir, Mii; integer(c_int64_t), value :: ld_nlm, N
            integer(c_int), value :: src_only
        end function
        integer(c_int) function sfb_apply_bounds_arr(nlm_in, nlm_out, N, ld) bind(c, name='sfb_apply_bounds_arr')   ! src/dynamics.f90:530
            import; type(c_ptr), value :: nlm_in, nlm_out; integer(c_int64_t), value :: N, ld
        end function
        integer(c_int) function sfb_ai_to_nlm_arr(rank, a, N, nlm) bind(c, name='sfb_ai_to_nlm_arr')               ! src/moments.f90:68-92
            import; integer(c_int), value :: rank; type(c_ptr), value :: a, nlm; integer(c_int64_t), value :: N
        end function
    end interface

contains

    subroutine check(rc, what)
        integer(c_int), intent(in) :: rc
        character(*), intent(in)   :: what
        if (rc /= 0) then
            print *, 'specfab_b200 error in ', what, ' code ', rc
            stop 'specfab error'
        end if
    end subroutine

    ! nlm(N,nlm_len) <- one fused step of every node: nlm + dt*matmul(M_LROT+Gamma0*M_DDRX+Lambda*M_CDRX+M_REG, nlm)
    ! batches the loop of src/dynamics.f90:99-110 / src/specfabpy/integrator.py:73-77
    subroutine step_arr(nlm, ugrad, tau, dt, iota, zeta, Gamma0, Lambda, terms, scheme)
        complex(kind=dp), intent(inout), target, contiguous :: nlm(:,:)          ! (N, nlm_len); c_loc needs contiguous storage
        real(kind=dp), intent(in), target, contiguous       :: ugrad(:,:,:), tau(:,:,:)   ! (N,3,3)
        real(kind=dp), intent(in)               :: dt, iota, zeta, Gamma0, Lambda
        integer, intent(in)                     :: terms, scheme
        type(sfb_step_opts) :: o
        o = sfb_step_opts(dt, iota, zeta, 1.0d0, Gamma0, Lambda, c_null_ptr, c_null_ptr, terms, scheme, 1, 0)
        call check(sfb_step_arr(c_loc(nlm), c_loc(nlm), int(size(nlm,1),c_int64_t), int(size(nlm,1),c_int64_t), &
                                c_loc(ugrad), c_loc(tau), o), 'step_arr')
    end subroutine

    ! the same step with the batch sharded over the GPUs listed in devices (0-based CUDA ordinals)
    subroutine step_arr_multi(nlm, ugrad, tau, dt, iota, zeta, Gamma0, Lambda, terms, scheme, devices)
        complex(kind=dp), intent(inout), target, contiguous :: nlm(:,:)
        real(kind=dp), intent(in), target, contiguous       :: ugrad(:,:,:), tau(:,:,:)
        real(kind=dp), intent(in)               :: dt, iota, zeta, Gamma0, Lambda
        integer, intent(in)                     :: terms, scheme
        integer(c_int), intent(in)              :: devices(:)
        type(sfb_step_opts) :: o
        o = sfb_step_opts(dt, iota, zeta, 1.0d0, Gamma0, Lambda, c_null_ptr, c_null_ptr, terms, scheme, 1, 0)
        call check(sfb_step_arr_multi(c_loc(nlm), c_loc(nlm), int(size(nlm,1),c_int64_t), int(size(nlm,1),c_int64_t), &
                                      c_loc(ugrad), c_loc(tau), o, devices, int(size(devices),c_int)), 'step_arr_multi')
    end subroutine

    ! the same for states kept in reduced form (what src/specfabpy/fenics/CPO.py holds): rnlm(N, rnlm_len), m >= 0 only
    subroutine step_rnlm_arr(rnlm, ugrad, tau, dt, iota, zeta, Gamma0, Lambda, terms, scheme)
        complex(kind=dp), intent(inout), target, contiguous :: rnlm(:,:)         ! (N, rnlm_len)
        real(kind=dp), intent(in), target, contiguous       :: ugrad(:,:,:), tau(:,:,:)   ! (N,3,3)
        real(kind=dp), intent(in)               :: dt, iota, zeta, Gamma0, Lambda
        integer, intent(in)                     :: terms, scheme
        type(sfb_step_opts) :: o
        o = sfb_step_opts(dt, iota, zeta, 1.0d0, Gamma0, Lambda, c_null_ptr, c_null_ptr, terms, scheme, 1, 0)
        call check(sfb_step_rnlm_arr(c_loc(rnlm), c_loc(rnlm), int(size(rnlm,1),c_int64_t), int(size(rnlm,1),c_int64_t), &
                                     c_loc(ugrad), c_loc(tau), o), 'step_rnlm_arr')
    end subroutine

    ! drop-in for Eij_tranisotropic_arr (src/specfabpy.f90:474-486)
    function Eij_tranisotropic_arr(nlm, e1,e2,e3, Eij_grain,alpha,n_grain) result(Eij)
        complex(kind=dp), intent(in), target, contiguous :: nlm(:,:)
        real(kind=dp), intent(in), target    :: e1(size(nlm,1),3), e2(size(nlm,1),3), e3(size(nlm,1),3)
        real(kind=dp), intent(in)            :: Eij_grain(2), alpha
        integer, intent(in)                  :: n_grain
        real(kind=dp), target                :: Eij(size(nlm,1),6)
        call check(sfb_Eij_tranisotropic_arr(c_loc(nlm), int(size(nlm,1),c_int64_t), int(size(nlm,1),c_int64_t), &
                   c_loc(e1), c_loc(e2), c_loc(e3), Eij_grain, alpha, int(n_grain,c_int), c_loc(Eij), c_null_ptr), &
                   'Eij_tranisotropic_arr')
    end function

    ! drop-in for Eij_orthotropic_arr (src/specfabpy.f90:488-500)
    function Eij_orthotropic_arr(nlm_1, nlm_2, nlm_3, e1,e2,e3, Eij_grain,alpha,n_grain) result(Eij)
        complex(kind=dp), intent(in), target, contiguous :: nlm_1(:,:), nlm_2(:,:), nlm_3(:,:)
        real(kind=dp), intent(in), target    :: e1(size(nlm_1,1),3), e2(size(nlm_1,1),3), e3(size(nlm_1,1),3)
        real(kind=dp), intent(in)            :: Eij_grain(6), alpha
        integer, intent(in)                  :: n_grain
        real(kind=dp), target                :: Eij(size(nlm_1,1),6)
        call check(sfb_Eij_orthotropic_arr(c_loc(nlm_1), c_loc(nlm_2), c_loc(nlm_3), int(size(nlm_1,1),c_int64_t), &
                   int(size(nlm_1,1),c_int64_t), c_loc(e1), c_loc(e2), c_loc(e3), Eij_grain, alpha, int(n_grain,c_int), c_loc(Eij)), &
                   'Eij_orthotropic_arr')
    end function

    ! drop-in for E_CAFFE_arr (src/specfabpy.f90:543-554)
    function E_CAFFE_arr(nlm, eps, Emin, Emax, n_grain) result(E)
        complex(kind=dp), intent(in), target, contiguous :: nlm(:,:)
        real(kind=dp), intent(in), target    :: eps(size(nlm,1),3,3)
        real(kind=dp), intent(in)            :: Emin, Emax
        integer, intent(in)                  :: n_grain
        real(kind=dp), target                :: E(size(nlm,1))
        call check(sfb_E_CAFFE_arr(c_loc(nlm), int(size(nlm,1),c_int64_t), int(size(nlm,1),c_int64_t), c_loc(eps), Emin, Emax, &
                   int(n_grain,c_int), c_loc(E)), 'E_CAFFE_arr')
    end function

    ! scalar forms keep the reference signatures (N = 1 batches)
    function a2(nlm) result(res)                                   ! src/moments.f90:37
        complex(kind=dp), intent(in), target, contiguous :: nlm(:)
        real(kind=dp), target :: res(3,3)
        call check(sfb_a2_arr(c_loc(nlm), 1_c_int64_t, 1_c_int64_t, c_loc(res)), 'a2')
    end function

    function a4(nlm) result(res)                                   ! src/moments.f90:46
        complex(kind=dp), intent(in), target, contiguous :: nlm(:)
        real(kind=dp), target :: res(3,3,3,3)
        call check(sfb_a4_arr(c_loc(nlm), 1_c_int64_t, 1_c_int64_t, c_loc(res)), 'a4')
    end function

    subroutine eig(nlm, ei, lami)                                  ! src/frames.f90:14
        complex(kind=dp), intent(in), target, contiguous :: nlm(:)
        real(kind=dp), intent(out), target   :: ei(3,3), lami(3)
        call check(sfb_eig_arr(c_loc(nlm), 1_c_int64_t, 1_c_int64_t, c_loc(ei), c_loc(lami)), 'eig')
    end subroutine

end module specfab_b200
