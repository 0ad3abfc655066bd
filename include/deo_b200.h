/*
 * deo_b200.h -- C ABI of the B200-native operator-application library
 * (libdeo_b200.so), the drop-in for the `mul!` / `*` hot path of
 * SciML/DiffEqOperators.jl v4.45.0.
 *
 * The reference has no FFI: the path sits behind Julia multiple dispatch
 * (methods of LinearAlgebra.mul! and Base.:* ).  Each entry point below names
 * the reference method(s) it replaces (paths relative to the reference's
 * src/).  The Julia side keeps its constructors and operator objects and hands
 * their *fields* across this boundary with `ccall`
 * (diffeqoperators.jl_b200/julia/DiffEqOperatorsB200.jl).
 *
 * Conventions
 *  - every function returns an int32 status (DEO_OK == 0); no C++ exception or
 *    abort crosses the boundary; deo_last_error() gives the thread-local text.
 *  - arrays are column-major (Julia): dims[0] is the fastest axis.
 *  - axes are 0-based here (Julia's {N} parameter minus one).
 *  - device buffers are owned by the library (deo_buffer); a plan or buffer may
 *    be used from any host thread, one thread at a time.
 *  - deo_plan_apply() returns after enqueue on the library stream of the
 *    current device; deo_buffer_download() and deo_sync() synchronise.
 *  - there is NO CPU fallback: without a CUDA device every compute entry point
 *    fails with DEO_ERR_CUDA.
 */
#ifndef DEO_B200_H
#define DEO_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DEO_ABI_VERSION 1

enum { DEO_OK = 0, DEO_ERR_INVALID = 1, DEO_ERR_CUDA = 2, DEO_ERR_UNSUPPORTED = 3, DEO_ERR_NCCL = 4, DEO_ERR_NOMEM = 5 };
enum { DEO_F32 = 0, DEO_F64 = 1 };
enum { DEO_OP_CENTERED = 0, DEO_OP_UPWIND = 1 };              /* Wind type parameter, derivative_operator.jl:14 */
enum { DEO_BC_NONE = 0,      /* input already carries its ghost layer along this axis (plain padded array) */
       DEO_BC_AFFINE = 1,    /* Robin/General/Dirichlet/Neumann: ghost = a . u[edge] + b, bc_operators.jl:188-191 */
       DEO_BC_PERIODIC = 2 };/* bc_operators.jl:192 (1-D), multi_dim_bc_operators.jl:221-228 (N-D) */
enum { DEO_FLAG_FORCE_GENERIC = 1 };   /* plan flags: always use the per-point kernel (debug / A-B runs) */
#define DEO_MAX_DIMS 3
#define DEO_MAX_OPS 16
#define DEO_MAX_TAPS 17      /* stencil_length and boundary_stencil_length limit (d + a <= 17) */

/* The operand bundle of one DerivativeOperator, field for field
 * (derivative_operators/derivative_operator.jl:14-29).  Host pointers, element
 * type = the plan dtype; copied during deo_plan_create.
 *   stencil_coefs        uniform: [sl]
 *                        non-uniform centered: [len-2*bpc][sl]            (:166-170)
 *                        non-uniform upwind:   [2][len-2*bpc][sl]         (:565-579; set 0 upwind, 1 downwind)
 *   low_boundary_coefs   [bpc][bsl]; non-uniform upwind [2][bpc][bsl]     (:531-562)
 *   high_boundary_coefs  centered [bpc][bsl]; uniform upwind [bpc+offside][bsl];
 *                        non-uniform upwind [2][bpc+offside][bsl]         (:582-623)
 *   coefficients         [len]                                             (coefficient_functions.jl:7-15) */
typedef struct deo_op_desc {
    int32_t axis;                       /* diff_axis(A) - 1 */
    int32_t kind;                       /* DEO_OP_* (use_winding(A)) */
    int32_t nonuniform;                 /* !(A.dx isa Number) */
    int32_t derivative_order;
    int32_t len;
    int32_t stencil_length;
    int32_t boundary_stencil_length;
    int32_t boundary_point_count;
    int32_t offside;
    int32_t reserved;
    const void *stencil_coefs;
    const void *low_boundary_coefs;
    const void *high_boundary_coefs;
    const void *coefficients;
} deo_op_desc;

/* One boundary-condition operator for one axis: the 4-tuple (a_l, b_l, a_r,
 * b_r) of an AffineBC (bc_operators.jl:21-25, :85-89).  per_face = 0: one BC
 * for every boundary pencil (MultiDimBC{dim}(BC, size), multi_dim_bc_operators.jl:97-100);
 * per_face = 1: tables with one row per boundary pencil, pencils enumerated
 * column-major over the remaining axes (MultiDimDirectionalBC.BCs, :54-57):
 * a_l[face][K_l], b_l[face], a_r[face][K_r], b_r[face]. */
typedef struct deo_bc_desc {
    int32_t kind;                       /* DEO_BC_* */
    int32_t per_face;
    int32_t K_l, K_r;                   /* length(a_l), length(a_r) */
    const void *a_l, *b_l, *a_r, *b_r;
} deo_bc_desc;

/* A fused application  du = sum_k (L_k o Q)(u)   (ghost_derivative_operator.jl:11-24,
 * composite_operators.jl:64-65), or du = sum_k L_k * M for a pre-padded M
 * (derivative_operator_functions.jl:27-69, :203, :466). */
typedef struct deo_plan_desc {
    int32_t dtype;                      /* DEO_F32 / DEO_F64 */
    int32_t ndims;                      /* 1..3 (collapse other dims: an axis-k op on an N-D array is (pre, n, post)) */
    int64_t dims[DEO_MAX_DIMS];         /* size(du) */
    int32_t padded[DEO_MAX_DIMS];       /* 1: size(u, axis) == dims[axis] + 2 (ghost layer present in the input) */
    int32_t nops;
    int32_t accumulate;                 /* overwrite = false (convolutions.jl:17-22): du += result */
    const deo_op_desc *ops;             /* in A.ops order: the sum is folded left in this order */
    deo_bc_desc bc[DEO_MAX_DIMS];       /* per axis; DEO_BC_NONE where padded[axis] or no operator acts */
    int32_t flags;
    int32_t reserved;
} deo_plan_desc;

typedef struct deo_plan deo_plan;
typedef struct deo_buffer deo_buffer;

/* ---- runtime ---------------------------------------------------------------------------- */
int32_t deo_abi_version(void);
int32_t deo_device_count(int32_t *count);
int32_t deo_init(int32_t device);                 /* select device, create the library stream */
int32_t deo_sync(void);                           /* wait for the library stream */
int32_t deo_last_error(char *buf, size_t len);    /* copy the calling thread's last message */
int32_t deo_launch_count(int64_t *count);         /* kernels launched by this library so far (process-wide) */

/* ---- device buffers (library-owned; the Julia DeviceArray handle wraps one) ---------------- */
int32_t deo_buffer_create(size_t bytes, deo_buffer **out);
int32_t deo_buffer_free(deo_buffer *buf);
int32_t deo_buffer_size(const deo_buffer *buf, size_t *bytes);
int32_t deo_buffer_upload(deo_buffer *dst, const void *host, size_t bytes);     /* copyto!(::DeviceArray, ::Array) */
int32_t deo_buffer_download(void *host, const deo_buffer *src, size_t bytes);   /* Array(::DeviceArray), synchronises */
int32_t deo_buffer_devptr(const deo_buffer *buf, void **devptr);                /* raw pointer, e.g. to alias from torch */
int32_t deo_buffer_wrap(void *devptr, size_t bytes, deo_buffer **out);          /* non-owning view of foreign device memory */
/* dst[i] = scale * a[i] * b[i] (+ dst[i] when accumulate): the products of derivative results that
 * nonlinear_diffusion! (derivative_operators/derivative_operator.jl:31-66) sums; first n elements of type dtype. */
int32_t deo_buffer_muladd(deo_buffer *dst, const deo_buffer *a, const deo_buffer *b, double scale, int32_t accumulate, int64_t n, int32_t dtype);
int32_t deo_host_alloc(size_t bytes, void **host);                              /* pinned host memory for the host-buffer path */
int32_t deo_host_free(void *host);

/* ---- plans ---------------------------------------------------------------------------------- */
/* Built lazily by the Julia glue from the operator object and cached there. */
int32_t deo_plan_create(const deo_plan_desc *desc, deo_plan **out);
int32_t deo_plan_destroy(deo_plan *plan);
/* DiffEqBase.update_coefficients!(A,u,p,t) (abstract_operator_functions.jl:190-194,
 * ghost_derivative_operator.jl:61-63): replace A.ops[op].coefficients. */
int32_t deo_plan_update_coefficients(deo_plan *plan, int32_t op, const void *coefficients);
/* mul!(du, A, u): convolutions.jl:17-22 (1-D), derivative_operator_functions.jl:18-69 (N-D),
 * ghost_derivative_operator.jl:15-24 (L*Q), composite_operators.jl:64-65,:76-83 (sums). */
int32_t deo_plan_apply(deo_plan *plan, deo_buffer *du, const deo_buffer *u);
/* out = u + dt * (A u): the update every explicit stepper performs right after mul! (test/DerivativeOperators/
 * 3D_laplacian.jl:20-24, heat_equation.jl:30-33), fused into the store of the tiled kernels: one read of u and one
 * write of out per point instead of mul! followed by an AXPY pass.  u must have the shape of the result. */
int32_t deo_plan_apply_axpy(deo_plan *plan, deo_buffer *out, const deo_buffer *u, double dt);
/* `reps` back-to-back applications replayed from one CUDA graph ("repeated mul!"). */
int32_t deo_plan_apply_n(deo_plan *plan, deo_buffer *du, const deo_buffer *u, int32_t reps);
/* Host-buffer form of mul!: H2D copy of u, apply, D2H copy of du, synchronous. */
int32_t deo_plan_apply_host(deo_plan *plan, void *du_host, const void *u_host);
/* Which kernel the plan dispatches to ("generic", "star", ...) and how many kernel launches one apply costs. */
int32_t deo_plan_info(const deo_plan *plan, char *kernel_name, size_t len, int32_t *launches_per_apply);
/* Timing helper for benchmarks: average milliseconds per apply over `reps` graph-replayed
 * applications, measured with CUDA events on the library stream. */
int32_t deo_plan_time(deo_plan *plan, deo_buffer *du, const deo_buffer *u, int32_t reps, float *ms_per_apply);

/* ---- slab-decomposed plans: one process per GPU, slabs along the last axis ----------------------- */
/* The host runtime (torch.distributed / MPI / Distributed.jl) only has to move the 128-byte
 * NCCL unique id from rank 0 to the other ranks. */
#define DEO_DIST_ID_BYTES 128
typedef struct deo_dist deo_dist;
int32_t deo_dist_unique_id(void *id_bytes);                                           /* rank 0 */
int32_t deo_dist_init(const void *id_bytes, int32_t rank, int32_t nranks, deo_dist **out);
int32_t deo_dist_destroy(deo_dist *ctx);
/* desc describes the GLOBAL problem; the last axis (ndims-1) is split into `nranks` contiguous slabs
 * (deo_dist_slab gives [start, start+count) for a rank).  The local field buffer holds
 * count + 2*halo planes (halo = deo_dist_plan_halo): [halo | own planes | halo]. */
int32_t deo_dist_slab(int64_t n_last, int32_t nranks, int32_t rank, int64_t *start, int64_t *count);
int32_t deo_dist_plan_create(deo_dist *ctx, const deo_plan_desc *global_desc, deo_plan **out);
int32_t deo_dist_plan_halo(const deo_plan *plan, int32_t *halo);
/* Halo exchange (ncclSend/ncclRecv to both slab neighbours on a communication stream) overlapped
 * with the interior planes; boundary planes follow.  u: extended local buffer, du: count planes. */
int32_t deo_dist_plan_apply(deo_plan *plan, deo_buffer *du, deo_buffer *u);
/* Host-buffer form of mul! on a slab (the multi-GPU counterpart of deo_plan_apply_host): u_own_host and du_host
 * hold this rank's `count` planes (no halo).  Chunks of planes are uploaded, computed and downloaded on three
 * streams; the edge chunks go up first so that the halo pushes to the neighbours overlap the rest.  Synchronous. */
int32_t deo_dist_plan_apply_host(deo_plan *plan, void *du_host, const void *u_own_host);
int32_t deo_dist_plan_time(deo_plan *plan, deo_buffer *du, deo_buffer *u, int32_t reps, float *ms_per_apply);
/* Single-process emulation used by the tests: build rank `rank` of `nranks`' local plan without NCCL;
 * the caller fills the halo planes itself. */
int32_t deo_dist_plan_create_local(const deo_plan_desc *global_desc, int32_t rank, int32_t nranks, deo_plan **out);

#ifdef __cplusplus
}
#endif
#endif /* DEO_B200_H */
