!> krylov_b200.f90 -- ISO_C_BINDING shim that puts libkrylov_b200.so (B200 / sm_100a) behind the
!> reference's own module and procedure names, so that the reference's driver programs
!> (tests/test_poisson_mf.f90, tests/test_cg.f90, tests/test_bicgstab.f90, tests/strong_scaling.f90,
!> tests/weak_scaling.f90) compile UNCHANGED and run on the GPU.
!>
!> Build (on a machine that has gfortran -- the build image of this repository does not, so this
!> file is shipped uncompiled; see INTEGRATION.md):
!>
!>   gfortran -O2 -c src/interfaces.f90 src/problems/poisson.f90 src/preconds/chebyshev.f90 \
!>            src/utils/utils.f90                      # reference modules that stay (plug-ins, printing)
!>   gfortran -O2 -c fortran/krylov_b200.f90           # replaces src/gmres_mgsr.f90, src/gmres_hh.f90,
!>                                                      #          src/cg.f90, src/bicgstab.f90
!>   gfortran -fopenmp tests/test_poisson_mf.f90 *.o -L gmres_b200 -lkrylov_b200 -o test_mfp
!>
!> The modules below have the reference's names (GMRES_MGSR_MOD, gmres_hh_mod,
!> conjugate_gradient, bicgstab_mod) and export procedures with the reference's names and dummy
!> argument lists (src/gmres_mgsr.f90:98,277; src/gmres_hh.f90:211,388; src/cg.f90:11,44,83,154;
!> src/bicgstab.f90:12,49,91).  The operator / preconditioner dummy procedures are recognised by
!> address: poisson::stvec, poisson::stv_poisson and chebyshev_precond::cbpr2 map to the library's
!> fused built-ins.  Any other procedure is a host routine that cannot run on device-resident
!> vectors: the shim stops with a message instead of silently falling back to the CPU.
module krylov_b200_c
    use iso_c_binding
    implicit none
    integer(c_int), parameter :: KL_OK = 0, KL_NOT_CONVERGED = 1, KL_BREAKDOWN = 2
    integer(c_int), parameter :: KL_OP_POISSON5 = 0, KL_OP_POISSON5_BRANCHY = 1, KL_OP_ANISO5 = 2, KL_OP_USER = 100
    integer(c_int), parameter :: KL_PC_NONE = 0, KL_PC_CBPR2 = 1, KL_PC_CHEB = 2, KL_PC_USER = 100

    type, bind(C) :: kl_operator_t      ! include/krylov_b200.h  kl_operator_t
        integer(c_int) :: kind
        real(c_double) :: eps_x, eps_y
        type(c_funptr) :: fn
        type(c_ptr)    :: user
    end type
    type, bind(C) :: kl_precond_t       ! include/krylov_b200.h  kl_precond_t
        integer(c_int) :: kind
        integer(c_int) :: degree
        type(c_funptr) :: fn
        type(c_ptr)    :: user
    end type

    type(c_ptr), save :: kl_handle = c_null_ptr   ! one process-wide handle on device 0

    interface
        function kl_create(h, device) bind(C, name="kl_create") result(rc)
            import :: c_ptr, c_int
            type(c_ptr), intent(out) :: h
            integer(c_int), value :: device
            integer(c_int) :: rc
        end function
        function kl_gmres_mgsr_omp(h, A, b, x, nx, ny, m, tol, final_err, v_err, n_out, restart_out, M, params, np) &
                bind(C, name="kl_gmres_mgsr_omp") result(rc)
            import :: c_ptr, c_int, c_double, kl_operator_t, kl_precond_t
            type(c_ptr), value :: h
            type(kl_operator_t), intent(in) :: A
            real(c_double), intent(in) :: b(*)
            real(c_double), intent(out) :: x(*), final_err(*), v_err(*)
            integer(c_int), value :: nx, ny, m
            real(c_double), value :: tol
            integer(c_int), intent(out) :: n_out, restart_out
            type(kl_precond_t), intent(in) :: M
            real(c_double), intent(in) :: params(*)
            integer(c_int), value :: np
            integer(c_int) :: rc
        end function
        function kl_gmres_mgsr_mf(h, A, b, x, nx, ny, m, tol, final_err, v_err, n_out, restart_out, M, params, np) &
                bind(C, name="kl_gmres_mgsr_mf") result(rc)
            import :: c_ptr, c_int, c_double, kl_operator_t, kl_precond_t
            type(c_ptr), value :: h
            type(kl_operator_t), intent(in) :: A
            real(c_double), intent(in) :: b(*)
            real(c_double), intent(out) :: x(*), final_err(*), v_err(*)
            integer(c_int), value :: nx, ny, m
            real(c_double), value :: tol
            integer(c_int), intent(out) :: n_out, restart_out
            type(kl_precond_t), intent(in) :: M
            real(c_double), intent(in) :: params(*)
            integer(c_int), value :: np
            integer(c_int) :: rc
        end function
        function kl_gmres_hh_omp(h, A, b, x, nx, ny, m, tol, final_err, v_err, n_out, stages_out) &
                bind(C, name="kl_gmres_hh_omp") result(rc)
            import :: c_ptr, c_int, c_double, kl_operator_t
            type(c_ptr), value :: h
            type(kl_operator_t), intent(in) :: A
            real(c_double), intent(in) :: b(*)
            real(c_double), intent(out) :: x(*), final_err(*), v_err(*)
            integer(c_int), value :: nx, ny, m
            real(c_double), value :: tol
            integer(c_int), intent(out) :: n_out, stages_out
            integer(c_int) :: rc
        end function
        function kl_gmres_hh_prec_omp(h, A, b, x, nx, ny, m, tol, final_err, v_err, n_out, stages_out, M, params, np) &
                bind(C, name="kl_gmres_hh_prec_omp") result(rc)
            import :: c_ptr, c_int, c_double, kl_operator_t, kl_precond_t
            type(c_ptr), value :: h
            type(kl_operator_t), intent(in) :: A
            real(c_double), intent(in) :: b(*)
            real(c_double), intent(out) :: x(*), final_err(*), v_err(*)
            integer(c_int), value :: nx, ny, m
            real(c_double), value :: tol
            integer(c_int), intent(out) :: n_out, stages_out
            type(kl_precond_t), intent(in) :: M
            real(c_double), intent(in) :: params(*)
            integer(c_int), value :: np
            integer(c_int) :: rc
        end function
        function kl_cg_omp(h, A, b, x, nx, ny, tol, iter, res) bind(C, name="kl_cg_omp") result(rc)
            import :: c_ptr, c_int, c_double, kl_operator_t
            type(c_ptr), value :: h
            type(kl_operator_t), intent(in) :: A
            real(c_double), intent(in) :: b(*)
            real(c_double), intent(out) :: x(*)
            integer(c_int), value :: nx, ny
            real(c_double), value :: tol
            integer(c_int), intent(inout) :: iter
            real(c_double), intent(out) :: res
            integer(c_int) :: rc
        end function
        function kl_pcg_omp(h, A, b, x, nx, ny, tol, iter, res, M, params, np) bind(C, name="kl_pcg_omp") result(rc)
            import :: c_ptr, c_int, c_double, kl_operator_t, kl_precond_t
            type(c_ptr), value :: h
            type(kl_operator_t), intent(in) :: A
            real(c_double), intent(in) :: b(*)
            real(c_double), intent(out) :: x(*)
            integer(c_int), value :: nx, ny
            real(c_double), value :: tol
            integer(c_int), intent(inout) :: iter
            real(c_double), intent(out) :: res
            type(kl_precond_t), intent(in) :: M
            real(c_double), intent(in) :: params(*)
            integer(c_int), value :: np
            integer(c_int) :: rc
        end function
        function kl_bicgstab(h, A, b, x, nx, ny, tol, iter, res) bind(C, name="kl_bicgstab") result(rc)
            import :: c_ptr, c_int, c_double, kl_operator_t
            type(c_ptr), value :: h
            type(kl_operator_t), intent(in) :: A
            real(c_double), intent(in) :: b(*)
            real(c_double), intent(out) :: x(*)
            integer(c_int), value :: nx, ny
            real(c_double), value :: tol
            integer(c_int), intent(inout) :: iter
            real(c_double), intent(out) :: res
            integer(c_int) :: rc
        end function
        function kl_pbicgstab_omp(h, A, b, x, nx, ny, tol, iter, res, M, params, np) &
                bind(C, name="kl_pbicgstab_omp") result(rc)
            import :: c_ptr, c_int, c_double, kl_operator_t, kl_precond_t
            type(c_ptr), value :: h
            type(kl_operator_t), intent(in) :: A
            real(c_double), intent(in) :: b(*)
            real(c_double), intent(out) :: x(*)
            integer(c_int), value :: nx, ny
            real(c_double), value :: tol
            integer(c_int), intent(inout) :: iter
            real(c_double), intent(out) :: res
            type(kl_precond_t), intent(in) :: M
            real(c_double), intent(in) :: params(*)
            integer(c_int), value :: np
            integer(c_int) :: rc
        end function
        function kl_gmres_mgsr_dense(h, A, n, b, x, m, tol, final_err, v_err, n_out, restart_out) &
                bind(C, name="kl_gmres_mgsr_dense") result(rc)
            import :: c_ptr, c_int, c_double
            type(c_ptr), value :: h
            real(c_double), intent(in) :: A(*), b(*)
            integer(c_int), value :: n, m
            real(c_double), intent(out) :: x(*), final_err(*), v_err(*)
            real(c_double), value :: tol
            integer(c_int), intent(out) :: n_out, restart_out
            integer(c_int) :: rc
        end function
        function kl_gmres_hh_dense(h, A, n, b, x, m, tol, final_err, v_err, n_out, stages_out) &
                bind(C, name="kl_gmres_hh_dense") result(rc)
            import :: c_ptr, c_int, c_double
            type(c_ptr), value :: h
            real(c_double), intent(in) :: A(*), b(*)
            integer(c_int), value :: n, m
            real(c_double), intent(out) :: x(*), final_err(*), v_err(*)
            real(c_double), value :: tol
            integer(c_int), intent(out) :: n_out, stages_out
            integer(c_int) :: rc
        end function
        function kl_generate_matrix(h, Hm, n) bind(C, name="kl_generate_matrix") result(rc)
            import :: c_ptr, c_int, c_double
            type(c_ptr), value :: h
            real(c_double), intent(out) :: Hm(*)
            integer(c_int), value :: n
            integer(c_int) :: rc
        end function
    end interface

contains

    !> lazily created process-wide handle
    function the_handle() result(h)
        type(c_ptr) :: h
        integer(c_int) :: rc
        if (.not. c_associated(kl_handle)) then
            rc = kl_create(kl_handle, 0_c_int)
            if (rc /= KL_OK) error stop "krylov_b200: kl_create failed (no CUDA device; there is no CPU fallback)"
        end if
        h = kl_handle
    end function

    !> procedure(stencil_vector) -> descriptor (src/interfaces.f90:12-18)
    function operator_of(Ax_vec) result(op)
        use interfaces
        use poisson, only: stvec, stv_poisson
        procedure(stencil_vector) :: Ax_vec
        type(kl_operator_t) :: op
        op%eps_x = 1.0d0; op%eps_y = 1.0d0; op%fn = c_null_funptr; op%user = c_null_ptr
        if (c_associated(c_funloc(Ax_vec), c_funloc(stvec))) then
            op%kind = KL_OP_POISSON5
        else if (c_associated(c_funloc(Ax_vec), c_funloc(stv_poisson))) then
            op%kind = KL_OP_POISSON5_BRANCHY
        else
            error stop "krylov_b200: operator is a host procedure; pass poisson::stvec / stv_poisson or a kl_operator_t KL_OP_USER"
        end if
    end function

    !> procedure(precond) -> descriptor (src/interfaces.f90:19-28)
    function precond_of(M_inv) result(pc)
        use interfaces
        use chebyshev_precond, only: cbpr2
        procedure(precond) :: M_inv
        type(kl_precond_t) :: pc
        pc%degree = 0; pc%fn = c_null_funptr; pc%user = c_null_ptr
        if (c_associated(c_funloc(M_inv), c_funloc(cbpr2))) then
            pc%kind = KL_PC_CBPR2
        else
            error stop "krylov_b200: preconditioner is a host procedure; pass chebyshev_precond::cbpr2"
        end if
    end function

    integer function grid_side(n)          ! nsize = int(sqrt(real(n)))  (gmres_mgsr.f90:298)
        integer, intent(in) :: n
        grid_side = int(sqrt(real(n)))
    end function
end module krylov_b200_c


MODULE GMRES_MGSR_MOD                      ! replaces src/gmres_mgsr.f90
    use interfaces
    use krylov_b200_c
    implicit none
    private
    public :: gmres_mgsr_omp, gmres_mgsr_mf, gmres_mgsr_dense
CONTAINS
    subroutine gmres_mgsr_omp(Ax_vec, b, x, m, tol, final_err, v_err, n_out, restart_out, M_inv, params)
        procedure(stencil_vector) :: Ax_vec
        real(8), intent(in) :: b(:)
        real(8), allocatable, intent(out) :: x(:)
        integer, intent(in) :: m
        real(8), intent(in) :: tol
        real(8), allocatable, intent(out) :: final_err(:), v_err(:)
        integer, intent(out) :: n_out, restart_out
        procedure(precond) :: M_inv
        real(8), intent(in) :: params(:)
        integer(c_int) :: rc, ns
        type(kl_operator_t) :: op
        type(kl_precond_t) :: pc
        ns = grid_side(size(b))
        allocate(x(size(b)), final_err(m), v_err(m + 1))      ! callee allocates (gmres_mgsr.f90:302)
        op = operator_of(Ax_vec); pc = precond_of(M_inv)
        rc = kl_gmres_mgsr_omp(the_handle(), op, b, x, ns, ns, int(m, c_int), tol, final_err, v_err, n_out, &
                               restart_out, pc, params, int(size(params), c_int))
        if (rc < 0) error stop "krylov_b200: kl_gmres_mgsr_omp failed"
    end subroutine

    subroutine gmres_mgsr_mf(Ax_vec, b, x, m, tol, final_err, v_err, n_out, restart_out, M_inv, params)
        procedure(stencil_vector) :: Ax_vec
        real(8), intent(in) :: b(:)
        real(8), allocatable, intent(out) :: x(:)
        integer, intent(in) :: m
        real(8), intent(in) :: tol
        real(8), allocatable, intent(out) :: final_err(:), v_err(:)
        integer, intent(out) :: n_out, restart_out
        procedure(precond) :: M_inv
        real(8), intent(in) :: params(:)
        integer(c_int) :: rc, ns
        type(kl_operator_t) :: op
        type(kl_precond_t) :: pc
        ns = grid_side(size(b))
        allocate(x(size(b)), final_err(m), v_err(m + 1))
        op = operator_of(Ax_vec); pc = precond_of(M_inv)
        rc = kl_gmres_mgsr_mf(the_handle(), op, b, x, ns, ns, int(m, c_int), tol, final_err, v_err, n_out, &
                              restart_out, pc, params, int(size(params), c_int))
        if (rc < 0) error stop "krylov_b200: kl_gmres_mgsr_mf failed"
    end subroutine

    subroutine gmres_mgsr_dense(A, b, x, m, tol, final_err, v_err, n_out, restart_out)   ! gmres_mgsr.f90:11
        real(8), intent(in) :: A(:,:), b(:)
        real(8), allocatable, intent(out) :: x(:)
        integer, intent(in) :: m
        real(8), intent(in) :: tol
        real(8), allocatable, intent(out) :: final_err(:), v_err(:)
        integer, intent(out) :: n_out, restart_out
        integer(c_int) :: rc
        allocate(x(size(b)), final_err(m), v_err(m + 1))      ! gmres_mgsr.f90:27
        rc = kl_gmres_mgsr_dense(the_handle(), A, int(size(A, 1), c_int), b, x, int(m, c_int), tol, final_err, v_err, &
                                 n_out, restart_out)
        if (rc < 0) error stop "krylov_b200: kl_gmres_mgsr_dense failed"
    end subroutine
END MODULE GMRES_MGSR_MOD


MODULE gmres_hh_mod                        ! replaces src/gmres_hh.f90
    use interfaces
    use krylov_b200_c
    implicit none
    private
    public :: gmres_hh_omp, gmres_hh_prec_omp, gmres_hh_dense
CONTAINS
    subroutine gmres_hh_omp(Ax_vec, b, x, m, tol, final_err, v_err, n_out, stages_out)
        procedure(stencil_vector) :: Ax_vec
        real(8), intent(in) :: b(:)
        real(8), allocatable, intent(out) :: x(:)
        integer, intent(in) :: m
        real(8), intent(in) :: tol
        real(8), allocatable, intent(out) :: final_err(:), v_err(:)
        integer, intent(out) :: n_out, stages_out
        integer(c_int) :: rc, ns
        type(kl_operator_t) :: op
        ns = grid_side(size(b))
        allocate(x(size(b)), final_err(m), v_err(m + 1))      ! gmres_hh.f90:232-234
        op = operator_of(Ax_vec)
        rc = kl_gmres_hh_omp(the_handle(), op, b, x, ns, ns, int(m, c_int), tol, final_err, v_err, n_out, stages_out)
        if (rc < 0) error stop "krylov_b200: kl_gmres_hh_omp failed"
    end subroutine

    subroutine gmres_hh_prec_omp(Ax_vec, b, x, m, tol, final_err, v_err, n_out, stages_out, m_inv, params)
        procedure(stencil_vector) :: Ax_vec
        real(8), intent(in) :: b(:)
        real(8), allocatable, intent(out) :: x(:)
        integer, intent(in) :: m
        real(8), intent(in) :: tol
        real(8), allocatable, intent(out) :: final_err(:), v_err(:)
        integer, intent(out) :: n_out, stages_out
        procedure(precond) :: m_inv
        real(8), intent(in) :: params(:)
        integer(c_int) :: rc, ns
        type(kl_operator_t) :: op
        type(kl_precond_t) :: pc
        ns = grid_side(size(b))
        allocate(x(size(b)), final_err(m), v_err(m + 1))
        op = operator_of(Ax_vec); pc = precond_of(m_inv)
        rc = kl_gmres_hh_prec_omp(the_handle(), op, b, x, ns, ns, int(m, c_int), tol, final_err, v_err, n_out, &
                                  stages_out, pc, params, int(size(params), c_int))
        if (rc < 0) error stop "krylov_b200: kl_gmres_hh_prec_omp failed"
    end subroutine

    subroutine gmres_hh_dense(A, b, x, m, tol, final_err, v_err, n_out, stages_out)      ! gmres_hh.f90:10
        real(8), intent(in) :: A(:,:), b(:)
        real(8), allocatable, intent(out) :: x(:)
        integer, intent(in) :: m
        real(8), intent(in) :: tol
        real(8), allocatable, intent(out) :: final_err(:), v_err(:)
        integer, intent(out) :: n_out, stages_out
        integer(c_int) :: rc
        allocate(x(size(b)), final_err(m), v_err(m + 1))      ! gmres_hh.f90:28-30
        rc = kl_gmres_hh_dense(the_handle(), A, int(size(A, 1), c_int), b, x, int(m, c_int), tol, final_err, v_err, &
                               n_out, stages_out)
        if (rc < 0) error stop "krylov_b200: kl_gmres_hh_dense failed"
    end subroutine
END MODULE gmres_hh_mod


MODULE hilbert                             ! replaces src/problems/hilbert.f90
    use krylov_b200_c
    implicit none
    private
    public :: generate_matrix
CONTAINS
    subroutine generate_matrix(H, n)
        real(8), allocatable, intent(out) :: H(:,:)
        integer, intent(in) :: n
        integer(c_int) :: rc
        allocate(H(n, n))                                     ! hilbert.f90:11
        rc = kl_generate_matrix(the_handle(), H, int(n, c_int))
        if (rc < 0) error stop "krylov_b200: kl_generate_matrix failed"
    end subroutine
END MODULE hilbert


MODULE conjugate_gradient                  ! replaces src/cg.f90
    use interfaces
    use krylov_b200_c
    implicit none
    private
    public :: cg, pcg, cg_omp, pcg_omp
CONTAINS
    subroutine cg_omp(Ax_op, b, x, tol, iter, res)
        procedure(stencil_vector) :: Ax_op
        real(8), intent(in) :: b(:)
        real(8), allocatable, intent(out) :: x(:)
        real(8), intent(in) :: tol
        integer, intent(inout) :: iter           ! maximum on entry, count on exit (cg.f90:88)
        real(8), intent(out) :: res
        integer(c_int) :: rc, ns
        type(kl_operator_t) :: op
        ns = grid_side(size(b))
        allocate(x(size(b)))
        op = operator_of(Ax_op)
        rc = kl_cg_omp(the_handle(), op, b, x, ns, ns, tol, iter, res)
        if (rc < 0) error stop "krylov_b200: kl_cg_omp failed"
    end subroutine
    subroutine cg(Ax_op, b, x, tol, iter, res)     ! serial twin: same device path (cg.f90:11)
        procedure(stencil_vector) :: Ax_op
        real(8), intent(in) :: b(:)
        real(8), allocatable, intent(out) :: x(:)
        real(8), intent(in) :: tol
        integer, intent(inout) :: iter
        real(8), intent(out) :: res
        call cg_omp(Ax_op, b, x, tol, iter, res)
    end subroutine
    subroutine pcg_omp(Ax_op, b, x, tol, iter, res, M_inv, params)
        procedure(stencil_vector) :: Ax_op
        real(8), intent(in) :: b(:)
        real(8), allocatable, intent(out) :: x(:)
        real(8), intent(in) :: tol
        integer, intent(inout) :: iter
        real(8), intent(out) :: res
        procedure(precond) :: M_inv
        real(8), intent(in) :: params(:)
        integer(c_int) :: rc, ns
        type(kl_operator_t) :: op
        type(kl_precond_t) :: pc
        ns = grid_side(size(b))
        allocate(x(size(b)))
        op = operator_of(Ax_op); pc = precond_of(M_inv)
        rc = kl_pcg_omp(the_handle(), op, b, x, ns, ns, tol, iter, res, pc, params, int(size(params), c_int))
        if (rc < 0) error stop "krylov_b200: kl_pcg_omp failed"
    end subroutine
    subroutine pcg(Ax_op, b, x, tol, iter, res, M_inv, params)   ! cg.f90:44
        procedure(stencil_vector) :: Ax_op
        real(8), intent(in) :: b(:)
        real(8), allocatable, intent(out) :: x(:)
        real(8), intent(in) :: tol
        integer, intent(inout) :: iter
        real(8), intent(out) :: res
        procedure(precond) :: M_inv
        real(8), intent(in) :: params(:)
        call pcg_omp(Ax_op, b, x, tol, iter, res, M_inv, params)
    end subroutine
END MODULE conjugate_gradient


module bicgstab_mod                        ! replaces src/bicgstab.f90
    use interfaces
    use krylov_b200_c
    implicit none
    private
    public :: bicgstab, pbicgstab, pbicgstab_omp
CONTAINS
    subroutine bicgstab(ax_op, b, x, tol, iter, res)
        procedure(stencil_vector) :: ax_op
        real(8), intent(in) :: b(:)
        real(8), allocatable, intent(out) :: x(:)
        real(8), intent(in) :: tol
        integer, intent(inout) :: iter
        real(8), intent(out) :: res
        integer(c_int) :: rc, ns
        type(kl_operator_t) :: op
        ns = grid_side(size(b))
        allocate(x(size(b)))
        op = operator_of(ax_op)
        rc = kl_bicgstab(the_handle(), op, b, x, ns, ns, tol, iter, res)
        if (rc < 0) error stop "krylov_b200: kl_bicgstab failed"
    end subroutine
    subroutine pbicgstab_omp(ax_op, b, x, tol, max_iter, res, m_inv, params)
        procedure(stencil_vector) :: ax_op
        real(8), intent(in) :: b(:)
        real(8), allocatable, intent(out) :: x(:)
        real(8), intent(in) :: tol
        integer, intent(inout) :: max_iter       ! bicgstab.f90:96; defined on exit even without convergence
        real(8), intent(out) :: res
        procedure(precond) :: m_inv
        real(8), intent(in) :: params(:)
        integer(c_int) :: rc, ns
        type(kl_operator_t) :: op
        type(kl_precond_t) :: pc
        ns = grid_side(size(b))
        allocate(x(size(b)))
        op = operator_of(ax_op); pc = precond_of(m_inv)
        rc = kl_pbicgstab_omp(the_handle(), op, b, x, ns, ns, tol, max_iter, res, pc, params, int(size(params), c_int))
        if (rc < 0) error stop "krylov_b200: kl_pbicgstab_omp failed"
    end subroutine
    subroutine pbicgstab(ax_op, b, x, tol, iter, res, m_inv, params)   ! bicgstab.f90:49
        procedure(stencil_vector) :: ax_op
        real(8), intent(in) :: b(:)
        real(8), allocatable, intent(out) :: x(:)
        real(8), intent(in) :: tol
        integer, intent(inout) :: iter
        real(8), intent(out) :: res
        procedure(precond) :: m_inv
        real(8), intent(in) :: params(:)
        call pbicgstab_omp(ax_op, b, x, tol, iter, res, m_inv, params)
    end subroutine
end module bicgstab_mod
