// Plan objects of libdeo_b200: validation and deep copy of the caller's operand bundles, kernel
// dispatch, CUDA-graph replay for repeated applications, the host-buffer path.
#include <cstdlib>

#include "common.hpp"

using namespace deo;

namespace deo {

static size_t esize(int dtype) { return dtype == DEO_F64 ? 8 : 4; }

static void copy_bytes(std::vector<unsigned char>& dst, const void* src, size_t bytes) {
    dst.resize(bytes);
    if (bytes) memcpy(dst.data(), src, bytes);
}

// Validates `desc` and deep-copies everything it points to into `plan` (global problem description).
int32_t plan_from_desc(const deo_plan_desc* desc, deo_plan* plan) {
    DEO_REQUIRE(desc != nullptr, "plan: null descriptor");
    DEO_REQUIRE(desc->dtype == DEO_F32 || desc->dtype == DEO_F64, "plan: dtype must be DEO_F32 or DEO_F64");
    DEO_REQUIRE(desc->ndims >= 1 && desc->ndims <= kMaxDims, "plan: ndims must be 1..%d (collapse other dims)", kMaxDims);
    DEO_REQUIRE(desc->nops >= 1 && desc->nops <= kMaxOps, "plan: nops must be 1..%d", kMaxOps);
    DEO_REQUIRE(desc->ops != nullptr, "plan: null ops");
    plan->dtype = desc->dtype;
    plan->ndims = desc->ndims;
    plan->accumulate = desc->accumulate ? 1 : 0;
    plan->flags = desc->flags;
    const size_t es = esize(desc->dtype);
    for (int a = 0; a < kMaxDims; ++a) {
        plan->dims[a] = a < desc->ndims ? desc->dims[a] : 1;
        plan->padded[a] = a < desc->ndims ? (desc->padded[a] ? 1 : 0) : 0;
        if (a < desc->ndims) {
            DEO_REQUIRE(desc->dims[a] >= 1 && desc->dims[a] < (1LL << 31) - 4, "plan: dims[%d] = %lld out of range", a, (long long)desc->dims[a]);
        }
    }
    bool axis_has_op[kMaxDims] = {false, false, false};
    plan->ops.clear();
    for (int k = 0; k < desc->nops; ++k) {
        const deo_op_desc& o = desc->ops[k];
        DEO_REQUIRE(o.axis >= 0 && o.axis < desc->ndims, "op %d: axis %d outside the array (ndims %d)", k, o.axis, desc->ndims);
        DEO_REQUIRE(o.kind == DEO_OP_CENTERED || o.kind == DEO_OP_UPWIND, "op %d: unknown kind %d", k, o.kind);
        // derivative_operator_functions.jl:40  @assert size(x_temp, N) + 2 == size(M, N)
        DEO_REQUIRE(o.len == desc->dims[o.axis], "op %d: len %d does not match size(du, %d) = %lld", k, o.len, o.axis + 1, (long long)desc->dims[o.axis]);
        DEO_REQUIRE(o.stencil_length >= 1 && o.stencil_length <= kMaxTaps, "op %d: stencil_length %d unsupported (max %d)", k, o.stencil_length, kMaxTaps);
        DEO_REQUIRE(o.boundary_stencil_length >= 1 && o.boundary_stencil_length <= kMaxBTaps, "op %d: boundary_stencil_length %d unsupported (max %d)", k, o.boundary_stencil_length, kMaxBTaps);
        DEO_REQUIRE(o.boundary_point_count >= 0 && o.offside >= 0, "op %d: negative boundary_point_count/offside", k);
        DEO_REQUIRE(o.stencil_coefs && o.coefficients, "op %d: null stencil_coefs/coefficients", k);
        const int nhigh = o.kind == DEO_OP_UPWIND ? o.boundary_point_count + o.offside : o.boundary_point_count;
        DEO_REQUIRE((o.boundary_point_count == 0 || o.low_boundary_coefs) && (nhigh == 0 || o.high_boundary_coefs), "op %d: null boundary coefficient arrays", k);
        if (o.kind == DEO_OP_CENTERED) {
            DEO_REQUIRE(o.stencil_length % 2 == 1 && o.boundary_point_count == o.stencil_length / 2 - 1,
                        "op %d: centered operator with inconsistent stencil_length/boundary_point_count", k);
        } else {
            DEO_REQUIRE(o.stencil_length == o.boundary_stencil_length && o.boundary_point_count == o.boundary_stencil_length - 2 - o.offside,
                        "op %d: upwind operator with inconsistent stencil geometry", k);
        }
        const int sets = (o.kind == DEO_OP_UPWIND && o.nonuniform) ? 2 : 1;
        const long long nint = o.nonuniform ? (long long)o.len - 2LL * o.boundary_point_count : 1;
        DEO_REQUIRE(nint >= 0, "op %d: len %d smaller than 2*boundary_point_count", k, o.len);
        HostOp h;
        h.d = o;
        copy_bytes(h.stencil, o.stencil_coefs, (size_t)sets * (size_t)nint * o.stencil_length * es);
        copy_bytes(h.low, o.low_boundary_coefs, (size_t)sets * o.boundary_point_count * o.boundary_stencil_length * es);
        copy_bytes(h.high, o.high_boundary_coefs, (size_t)sets * nhigh * o.boundary_stencil_length * es);
        copy_bytes(h.coeff, o.coefficients, (size_t)o.len * es);
        h.d.stencil_coefs = h.d.low_boundary_coefs = h.d.high_boundary_coefs = h.d.coefficients = nullptr;
        plan->ops.push_back(std::move(h));
        axis_has_op[o.axis] = true;
    }
    for (int a = 0; a < desc->ndims; ++a) {
        const deo_bc_desc& b = desc->bc[a];
        HostBC& H = plan->bc[a];
        H.d = b;
        DEO_REQUIRE(b.kind == DEO_BC_NONE || b.kind == DEO_BC_AFFINE || b.kind == DEO_BC_PERIODIC, "bc[%d]: unknown kind %d", a, b.kind);
        if (plan->padded[a])
            DEO_REQUIRE(b.kind == DEO_BC_NONE, "bc[%d]: the input already carries a ghost layer along this axis", a);
        if (axis_has_op[a])
            DEO_REQUIRE(plan->padded[a] || b.kind != DEO_BC_NONE,
                        "axis %d: the differentiated dimension must be padded (derivative_operator_functions.jl:40) or have a BC", a);
        if (b.kind == DEO_BC_AFFINE) {
            DEO_REQUIRE(b.K_l >= 0 && b.K_r >= 0 && b.K_l <= desc->dims[a] && b.K_r <= desc->dims[a], "bc[%d]: K_l/K_r out of range", a);
            DEO_REQUIRE(b.b_l && b.b_r && (b.K_l == 0 || b.a_l) && (b.K_r == 0 || b.a_r), "bc[%d]: null coefficient arrays", a);
            size_t nface = 1;
            if (b.per_face) for (int c = 0; c < desc->ndims; ++c) if (c != a) nface *= (size_t)desc->dims[c];
            copy_bytes(H.a_l, b.a_l, nface * b.K_l * es);
            copy_bytes(H.a_r, b.a_r, nface * b.K_r * es);
            copy_bytes(H.b_l, b.b_l, nface * es);
            copy_bytes(H.b_r, b.b_r, nface * es);
        }
        H.d.a_l = H.d.b_l = H.d.a_r = H.d.b_r = nullptr;
    }
    plan->slab_axis = -1;
    plan->slab_start = 0;
    plan->slab_count = 0;
    plan->halo = 0;
    return DEO_OK;
}

int32_t finalize_plan(deo_plan* plan) {
    // staging arena: wait until the previous build's copies have left it (normally long done), grow it to what that
    // build needed, then restart the allocation cursor so that this build reuses the same device blocks in the same order
    if (plan->stage_ev) DEO_CUDA(cudaEventSynchronize(plan->stage_ev));
    if (plan->stage_want > plan->stage_cap) {
        if (plan->stage) cudaFreeHost(plan->stage);
        plan->stage = nullptr; plan->stage_cap = 0;
        if (cudaHostAlloc((void**)&plan->stage, plan->stage_want, cudaHostAllocDefault) == cudaSuccess) plan->stage_cap = plan->stage_want;
        else { cudaGetLastError(); plan->stage = nullptr; }
    }
    plan->stage_used = 0; plan->stage_want = 0; plan->blob_cursor = 0;
    int32_t rc = build_device_plan(plan);
    if (rc) return rc;
    plan->kernel = "generic";
    plan->launches_per_apply = 1;
    plan->star.reset();
    plan->line.reset();
    if (!(plan->flags & DEO_FLAG_FORCE_GENERIC)) {
        rc = star_configure(plan);   // sets plan->kernel = "star" / "star-table" when eligible
        if (rc) return rc;
        rc = line_configure(plan);   // 1-D plans: "line" / "line-table"
        if (rc) return rc;
    }
    if (plan->graph_exec) { cudaGraphExecDestroy(plan->graph_exec); plan->graph_exec = nullptr; }
    if (plan->blobs.size() > plan->blob_cursor) plan->blobs.resize(plan->blob_cursor);   // blocks this build no longer uses
    if (!plan->stage_ev) DEO_CUDA(cudaEventCreateWithFlags(&plan->stage_ev, cudaEventDisableTiming));
    DEO_CUDA(cudaEventRecord(plan->stage_ev, rt().stream));
    return DEO_OK;
}

int32_t launch_plan(const deo_plan* plan, void* du, const void* u, long long z0, long long z1, cudaStream_t s) {
    if (z1 <= z0) return DEO_OK;
    if (plan->star) return launch_star(plan, du, u, z0, z1, s);
    if (plan->line) return launch_line(plan, du, u, s);
    return launch_generic(plan, du, u, z0, z1, s);
}

static int32_t check_buffers(const deo_plan* plan, const deo_buffer* du, const deo_buffer* u) {
    DEO_REQUIRE(plan && du && u, "apply: null argument");
    DEO_REQUIRE(u->bytes >= plan->in_elems() * plan->elem(), "apply: input buffer holds %zu bytes, plan needs %zu",
                u->bytes, plan->in_elems() * plan->elem());
    DEO_REQUIRE(du->bytes >= plan->out_elems() * plan->elem(), "apply: output buffer holds %zu bytes, plan needs %zu",
                du->bytes, plan->out_elems() * plan->elem());
    DEO_REQUIRE(du->ptr != u->ptr, "apply: du and u must not alias");
    return DEO_OK;
}

static long long last_extent(const deo_plan* plan) { return plan->ndims == 3 ? plan->local_dim(2) : 1; }

static int32_t ensure_graph(deo_plan* plan, deo_buffer* du, const deo_buffer* u, int reps) {
    if (plan->graph_exec && plan->graph_u == u->ptr && plan->graph_du == du->ptr && plan->graph_reps == reps) return DEO_OK;
    if (plan->graph_exec) { cudaGraphExecDestroy(plan->graph_exec); plan->graph_exec = nullptr; }
    cudaStream_t s = rt().stream;
    cudaGraph_t graph = nullptr;
    DEO_CUDA(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
    int32_t rc = DEO_OK;
    for (int i = 0; i < reps && rc == DEO_OK; ++i) rc = launch_plan(plan, du->ptr, u->ptr, 0, last_extent(plan), s);
    cudaError_t e = cudaStreamEndCapture(s, &graph);
    if (rc) { if (graph) cudaGraphDestroy(graph); return rc; }
    if (e != cudaSuccess) return cuda_fail(e, "cudaStreamEndCapture", __FILE__, __LINE__);
    e = cudaGraphInstantiate(&plan->graph_exec, graph, 0);
    cudaGraphDestroy(graph);
    if (e != cudaSuccess) { plan->graph_exec = nullptr; return cuda_fail(e, "cudaGraphInstantiate", __FILE__, __LINE__); }
    plan->graph_u = u->ptr; plan->graph_du = du->ptr; plan->graph_reps = reps;
    return DEO_OK;
}

template <typename T>
__global__ void k_axpy_inplace(T* __restrict__ out, const T* __restrict__ u, T dt, long long n) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        out[i] = fma(dt, out[i], u[i]);
}

// dst[i] = scale * a[i] * b[i] (+ dst[i]): the products of derivative results nonlinear_diffusion! forms
// (derivative_operator.jl:31-66)
template <typename T>
__global__ void k_muladd(T* __restrict__ dst, const T* __restrict__ a, const T* __restrict__ b, T scale, int accumulate, long long n) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const T v = scale * (a[i] * b[i]);
        dst[i] = accumulate ? dst[i] + v : v;
    }
}

}  // namespace deo

extern "C" {

int32_t deo_buffer_muladd(deo_buffer* dst, const deo_buffer* a, const deo_buffer* b, double scale, int32_t accumulate, int64_t n, int32_t dtype) {
    DEO_REQUIRE(dst && a && b && n >= 0, "deo_buffer_muladd: bad arguments");
    DEO_REQUIRE(dtype == DEO_F32 || dtype == DEO_F64, "deo_buffer_muladd: dtype must be DEO_F32 or DEO_F64");
    const size_t es = dtype == DEO_F64 ? 8 : 4;
    DEO_REQUIRE(dst->bytes >= (size_t)n * es && a->bytes >= (size_t)n * es && b->bytes >= (size_t)n * es, "deo_buffer_muladd: buffers shorter than %lld elements", (long long)n);
    int32_t rc = ensure_init();
    if (rc) return rc;
    if (n == 0) return DEO_OK;
    const unsigned grid = (unsigned)((n + 255) / 256 < 148 * 16 ? (n + 255) / 256 : 148 * 16);
    if (dtype == DEO_F64) k_muladd<double><<<grid, 256, 0, rt().stream>>>((double*)dst->ptr, (const double*)a->ptr, (const double*)b->ptr, scale, accumulate, n);
    else k_muladd<float><<<grid, 256, 0, rt().stream>>>((float*)dst->ptr, (const float*)a->ptr, (const float*)b->ptr, (float)scale, accumulate, n);
    DEO_CUDA(cudaGetLastError());
    g_launches += 1;
    return DEO_OK;
}

int32_t deo_plan_create(const deo_plan_desc* desc, deo_plan** out) {
    DEO_REQUIRE(out != nullptr, "deo_plan_create: null argument");
    *out = nullptr;
    int32_t rc = ensure_init();
    if (rc) return rc;
    std::unique_ptr<deo_plan> plan(new (std::nothrow) deo_plan());
    if (!plan) { set_error("out of host memory"); return DEO_ERR_NOMEM; }
    rc = plan_from_desc(desc, plan.get());
    if (rc) return rc;
    rc = finalize_plan(plan.get());
    if (rc) return rc;
    *out = plan.release();
    return DEO_OK;
}

int32_t deo_plan_destroy(deo_plan* plan) {
    if (!plan) return DEO_OK;
    if (rt().ready) { cudaStreamSynchronize(rt().stream); cudaStreamSynchronize(rt().comm_stream); }
    if (plan->graph_exec) cudaGraphExecDestroy(plan->graph_exec);
    if (plan->host_u) deo_buffer_free(plan->host_u);
    if (plan->host_du) deo_buffer_free(plan->host_du);
    for (cudaEvent_t e : plan->host_ev) cudaEventDestroy(e);
    if (plan->stage_ev) cudaEventDestroy(plan->stage_ev);
    if (plan->stage) cudaFreeHost(plan->stage);
    delete plan;
    return DEO_OK;
}

int32_t deo_plan_update_coefficients(deo_plan* plan, int32_t op, const void* coefficients) {
    DEO_REQUIRE(plan && coefficients, "deo_plan_update_coefficients: null argument");
    DEO_REQUIRE(op >= 0 && op < (int)plan->ops.size(), "deo_plan_update_coefficients: op %d out of range", op);
    HostOp& h = plan->ops[op];
    memcpy(h.coeff.data(), coefficients, h.coeff.size());
    // No synchronisation and no allocation: the row tables are recomputed on the host (the reference's own row / branch
    // logic, including the per-row wind direction of upwind operators from sign(c)), written into the pinned arena and sent
    // to the SAME device blocks with copies ordered on the library stream behind every application already enqueued; the
    // kernels' parameter blocks are host-side values taken by each launch.
    return finalize_plan(plan);
}

int32_t deo_plan_apply(deo_plan* plan, deo_buffer* du, const deo_buffer* u) {
    int32_t rc = check_buffers(plan, du, u);
    if (rc) return rc;
    DEO_REQUIRE(plan->dist == nullptr && plan->nranks == 1, "deo_plan_apply: slab plans go through deo_dist_plan_apply");
    g_launches += plan->launches_per_apply;
    return launch_plan(plan, du->ptr, u->ptr, 0, last_extent(plan), rt().stream);
}

// out = u + dt * (A u).  Tiled 2-D / 3-D plans fuse the update into the kernel's store (16 B/point instead of the 40 B/point
// of apply + a separate AXPY pass); every other plan applies and then runs the small update kernel below.
int32_t deo_plan_apply_axpy(deo_plan* plan, deo_buffer* out, const deo_buffer* u, double dt) {
    int32_t rc = check_buffers(plan, out, u);
    if (rc) return rc;
    DEO_REQUIRE(plan->dist == nullptr && plan->nranks == 1, "deo_plan_apply_axpy: slab plans are not supported");
    DEO_REQUIRE(!plan->accumulate, "deo_plan_apply_axpy: the plan accumulates (overwrite = false)");
    for (int a = 0; a < plan->ndims; ++a) DEO_REQUIRE(!plan->padded[a], "deo_plan_apply_axpy: u must have the shape of the result (no pre-padded axes)");
    cudaStream_t s = rt().stream;
    g_launches += plan->launches_per_apply;
    if (star_can_axpy(plan)) {
        star_set_axpy(plan, true, dt);
        rc = launch_plan(plan, out->ptr, u->ptr, 0, last_extent(plan), s);
        star_set_axpy(plan, false, 0.0);
        return rc;
    }
    rc = launch_plan(plan, out->ptr, u->ptr, 0, last_extent(plan), s);
    if (rc) return rc;
    const long long n = (long long)plan->out_elems();
    const unsigned grid = (unsigned)((n + 255) / 256 < 148 * 16 ? (n + 255) / 256 : 148 * 16);
    if (plan->dtype == DEO_F64) k_axpy_inplace<double><<<grid, 256, 0, s>>>((double*)out->ptr, (const double*)u->ptr, dt, n);
    else k_axpy_inplace<float><<<grid, 256, 0, s>>>((float*)out->ptr, (const float*)u->ptr, (float)dt, n);
    DEO_CUDA(cudaGetLastError());
    g_launches += 1;
    return DEO_OK;
}

int32_t deo_plan_apply_n(deo_plan* plan, deo_buffer* du, const deo_buffer* u, int32_t reps) {
    int32_t rc = check_buffers(plan, du, u);
    if (rc) return rc;
    DEO_REQUIRE(reps >= 1 && reps <= 100000, "deo_plan_apply_n: reps out of range");
    rc = ensure_graph(plan, du, u, reps);
    if (rc) return rc;
    DEO_CUDA(cudaGraphLaunch(plan->graph_exec, rt().stream));
    g_launches += (long long)reps * plan->launches_per_apply;
    return DEO_OK;
}

// Host-buffer form of mul!.  The field is cut into chunks of planes along the last axis and pushed through a
// three-stage pipeline on three streams -- upload chunk k+1 | fused kernel on chunk k | download chunk k-1 -- so the two
// PCIe directions and the kernel overlap; the kernel of chunk k needs the first `reach` planes of chunk k+1, so it
// waits for that upload.  The device staging buffers are allocated on the first call and kept by the plan.
int32_t deo_plan_apply_host(deo_plan* plan, void* du_host, const void* u_host) {
    DEO_REQUIRE(plan && du_host && u_host, "deo_plan_apply_host: null argument");
    DEO_REQUIRE(plan->dist == nullptr && plan->nranks == 1, "deo_plan_apply_host: slab plans go through deo_dist_plan_apply");
    const size_t in_b = plan->in_elems() * plan->elem(), out_b = plan->out_elems() * plan->elem();
    int32_t rc = DEO_OK;
    if (!plan->host_u) { rc = deo_buffer_create(in_b, &plan->host_u); if (rc) return rc; }
    if (!plan->host_du) { rc = deo_buffer_create(out_b, &plan->host_du); if (rc) return rc; }
    deo_buffer *u = plan->host_u, *du = plan->host_du;
    Runtime& R = rt();
    cudaStream_t s = R.stream;
    const int last = plan->ndims - 1;
    const long long nlast = plan->local_dim(last);
    const size_t in_plane = in_b / (size_t)plan->in_dim(last), out_plane = out_b / (size_t)nlast;
    // planes the operators along the last axis reach across (one-sided boundary rows included), and chunk length
    int reach = 0;
    for (const HostOp& h : plan->ops)
        if (h.d.axis == last) reach = reach > h.d.boundary_stencil_length ? reach : h.d.boundary_stencil_length;
    const size_t chunk_bytes = getenv("DEO_HOST_CHUNK_BYTES") ? (size_t)atoll(getenv("DEO_HOST_CHUNK_BYTES")) : ((size_t)96 << 20);
    long long chunk = (long long)(chunk_bytes / (out_plane ? out_plane : 1));
    if (chunk < 2 * reach + 8) chunk = 2 * reach + 8;
    const bool rangeable = plan->ndims == 3 || (plan->ndims == 2 && plan->star);   // kernels that take a range of the last axis
    // chunk boundaries: equal chunks, a short remainder is merged into the last one (a chunk must hold the one-sided rows)
    std::vector<long long> zb;
    if (rangeable && !plan->padded[last] && plan->bc[last].d.kind != DEO_BC_PERIODIC) {
        for (long long z = 0; z < nlast; z += chunk) zb.push_back(z);
        if (zb.size() > 1 && nlast - zb.back() < 2 * reach + 2) zb.pop_back();
        zb.push_back(nlast);
    }
    const long long nchunks = zb.empty() ? 1 : (long long)zb.size() - 1;
    const bool pipelined = nchunks >= 3;
    if (!pipelined) {
        DEO_CUDA(cudaMemcpyAsync(u->ptr, u_host, in_b, cudaMemcpyHostToDevice, s));
        if (plan->accumulate) DEO_CUDA(cudaMemcpyAsync(du->ptr, du_host, out_b, cudaMemcpyHostToDevice, s));
        rc = launch_plan(plan, du->ptr, u->ptr, 0, last_extent(plan), s);
        g_launches += plan->launches_per_apply;
        if (rc) return rc;
        DEO_CUDA(cudaMemcpyAsync(du_host, du->ptr, out_b, cudaMemcpyDeviceToHost, s));
        DEO_CUDA(cudaStreamSynchronize(s));
        return DEO_OK;
    }
    while ((long long)plan->host_ev.size() < 2 * nchunks) {
        cudaEvent_t e;
        DEO_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        plan->host_ev.push_back(e);
    }
    DEO_CUDA(cudaStreamSynchronize(s));
    // stage 1: uploads, in order, on the upload stream
    for (long long k = 0; k < nchunks; ++k) {
        const long long z0 = zb[(size_t)k], z1 = zb[(size_t)k + 1];
        DEO_CUDA(cudaMemcpyAsync((char*)u->ptr + (size_t)z0 * in_plane, (const char*)u_host + (size_t)z0 * in_plane, (size_t)(z1 - z0) * in_plane,
                                 cudaMemcpyHostToDevice, R.h2d_stream));
        if (plan->accumulate)
            DEO_CUDA(cudaMemcpyAsync((char*)du->ptr + (size_t)z0 * out_plane, (const char*)du_host + (size_t)z0 * out_plane,
                                     (size_t)(z1 - z0) * out_plane, cudaMemcpyHostToDevice, R.h2d_stream));
        DEO_CUDA(cudaEventRecord(plan->host_ev[(size_t)k], R.h2d_stream));
    }
    // stages 2 and 3: kernel on chunk k once chunk k+1 has landed, download behind it
    for (long long k = 0; k < nchunks; ++k) {
        const long long z0 = zb[(size_t)k], z1 = zb[(size_t)k + 1];
        DEO_CUDA(cudaStreamWaitEvent(s, plan->host_ev[(size_t)(k + 1 < nchunks ? k + 1 : k)], 0));
        if (plan->ndims == 3) rc = launch_plan(plan, du->ptr, u->ptr, z0, z1, s);
        else rc = launch_star(plan, du->ptr, u->ptr, z0, z1, s, true);     // 2-D: a range of rows of the last axis
        if (rc) return rc;
        g_launches += plan->launches_per_apply;
        DEO_CUDA(cudaEventRecord(plan->host_ev[(size_t)(nchunks + k)], s));
        DEO_CUDA(cudaStreamWaitEvent(R.d2h_stream, plan->host_ev[(size_t)(nchunks + k)], 0));
        DEO_CUDA(cudaMemcpyAsync((char*)du_host + (size_t)z0 * out_plane, (const char*)du->ptr + (size_t)z0 * out_plane, (size_t)(z1 - z0) * out_plane,
                                 cudaMemcpyDeviceToHost, R.d2h_stream));
    }
    DEO_CUDA(cudaStreamSynchronize(R.d2h_stream));
    DEO_CUDA(cudaStreamSynchronize(s));
    return DEO_OK;
}

int32_t deo_plan_info(const deo_plan* plan, char* kernel_name, size_t len, int32_t* launches_per_apply) {
    DEO_REQUIRE(plan != nullptr, "deo_plan_info: null plan");
    if (kernel_name && len) {
        size_t n = plan->kernel.size() < len - 1 ? plan->kernel.size() : len - 1;
        memcpy(kernel_name, plan->kernel.data(), n);
        kernel_name[n] = 0;
    }
    if (launches_per_apply) *launches_per_apply = plan->launches_per_apply;
    return DEO_OK;
}

int32_t deo_plan_time(deo_plan* plan, deo_buffer* du, const deo_buffer* u, int32_t reps, float* ms_per_apply) {
    int32_t rc = check_buffers(plan, du, u);
    if (rc) return rc;
    DEO_REQUIRE(reps >= 1 && ms_per_apply, "deo_plan_time: bad arguments");
    rc = ensure_graph(plan, du, u, reps);
    if (rc) return rc;
    cudaStream_t s = rt().stream;
    cudaEvent_t a, b;
    DEO_CUDA(cudaEventCreate(&a));
    DEO_CUDA(cudaEventCreate(&b));
    DEO_CUDA(cudaStreamSynchronize(s));
    DEO_CUDA(cudaEventRecord(a, s));
    DEO_CUDA(cudaGraphLaunch(plan->graph_exec, s));
    DEO_CUDA(cudaEventRecord(b, s));
    DEO_CUDA(cudaEventSynchronize(b));
    float ms = 0;
    DEO_CUDA(cudaEventElapsedTime(&ms, a, b));
    cudaEventDestroy(a);
    cudaEventDestroy(b);
    g_launches += (long long)reps * plan->launches_per_apply;
    *ms_per_apply = ms / reps;
    return DEO_OK;
}

}  // extern "C"
