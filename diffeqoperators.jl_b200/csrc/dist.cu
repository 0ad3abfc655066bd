// Slab decomposition along the last axis, one process per GPU.
//
// The reference has no distributed code at all; this is the one parallelism
// strategy the path needs: output points are independent given u, so the only coupling between
// slabs is the stencil reach along the decomposed axis.  In column-major storage a plane of the last
// axis is one contiguous block, so halo planes are sent/received in place (no packing):
//
//   local field buffer:  [ halo planes | own planes (count) | halo planes ]
//
// apply = { ncclSend/ncclRecv of `halo` planes with both neighbours on the communication stream }
//         overlapped with { the kernel on the planes that need no halo } on the compute stream,
//         then the kernel on the first/last `halo` planes once the exchange has landed.
//
// NCCL is bound at run time with dlopen so that a process that already carries an NCCL (torch's
// bundled one) shares it; nothing here links against libnccl.
#include <dlfcn.h>

#include <cstdlib>

#include "common.hpp"

namespace deo {
int32_t plan_from_desc(const deo_plan_desc* desc, deo_plan* plan);
int32_t finalize_plan(deo_plan* plan);
}  // namespace deo

using namespace deo;

// ---- minimal NCCL ABI (stable since NCCL 2.x; see nccl.h) -------------------------------------------
extern "C" {
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef enum { ncclSuccess = 0 } ncclResult_t;
typedef enum { ncclInt8 = 0, ncclChar = 0, ncclUint8 = 1, ncclInt32 = 2, ncclFloat32 = 7, ncclFloat64 = 8 } ncclDataType_t;
}
static_assert(sizeof(ncclUniqueId) == DEO_DIST_ID_BYTES, "unique id size");

namespace {
struct NcclApi {
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    bool ok = false;
};

NcclApi& nccl() {
    static NcclApi api;
    if (api.ok || api.handle) return api;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
        api.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (api.handle) break;
    }
    if (!api.handle) return api;
    auto sym = [&](const char* s) { return dlsym(api.handle, s); };
    api.GetUniqueId = (decltype(api.GetUniqueId))sym("ncclGetUniqueId");
    api.CommInitRank = (decltype(api.CommInitRank))sym("ncclCommInitRank");
    api.CommDestroy = (decltype(api.CommDestroy))sym("ncclCommDestroy");
    api.Send = (decltype(api.Send))sym("ncclSend");
    api.Recv = (decltype(api.Recv))sym("ncclRecv");
    api.GroupStart = (decltype(api.GroupStart))sym("ncclGroupStart");
    api.GroupEnd = (decltype(api.GroupEnd))sym("ncclGroupEnd");
    api.GetErrorString = (decltype(api.GetErrorString))sym("ncclGetErrorString");
    api.ok = api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.Send && api.Recv && api.GroupStart && api.GroupEnd;
    return api;
}

int32_t nccl_fail(ncclResult_t r, const char* what) {
    NcclApi& a = nccl();
    set_error("NCCL error %d (%s) in %s", (int)r, a.GetErrorString ? a.GetErrorString(r) : "?", what);
    return DEO_ERR_NCCL;
}
#define DEO_NCCL(call)                                                   \
    do {                                                                 \
        ncclResult_t r_ = (call);                                        \
        if (r_ != ncclSuccess) return nccl_fail(r_, #call);              \
    } while (0)
}  // namespace

struct deo_dist {
    ncclComm_t comm = nullptr;
    int rank = 0, nranks = 1;
    cudaEvent_t ev_ready = nullptr, ev_halo = nullptr;
    int* halo_flag = nullptr;      // device word: index of the last application whose halo planes have landed
    int halo_step = 0;
};

__global__ void k_publish_halo(int* flag, int step) {
    __threadfence();
    *reinterpret_cast<volatile int*>(flag) = step;
}

namespace {

// Reach of the operators acting on `axis`, in planes, below and above the output row.
int axis_reach(const deo_plan* plan, int axis) {
    int reach = 0;
    for (const HostOp& h : plan->ops) {
        if (h.d.axis != axis) continue;
        const int sl = h.d.stencil_length;
        int lo, hi;
        if (h.d.kind == DEO_OP_CENTERED) { lo = hi = sl / 2; }
        else { lo = hi = sl - 1 - h.d.offside; }   // either wind direction: max(offside, sl-1-offside)
        reach = reach > lo ? reach : lo;
        reach = reach > hi ? reach : hi;
    }
    return reach;
}

int32_t make_slab_plan(const deo_plan_desc* desc, int rank, int nranks, deo_dist* ctx, deo_plan** out) {
    DEO_REQUIRE(out != nullptr, "dist plan: null argument");
    *out = nullptr;
    int32_t rc = ensure_init();
    if (rc) return rc;
    DEO_REQUIRE(nranks >= 1 && rank >= 0 && rank < nranks, "dist plan: rank %d of %d", rank, nranks);
    std::unique_ptr<deo_plan> plan(new (std::nothrow) deo_plan());
    if (!plan) { set_error("out of host memory"); return DEO_ERR_NOMEM; }
    rc = plan_from_desc(desc, plan.get());
    if (rc) return rc;
    DEO_REQUIRE(plan->ndims == 3, "dist plan: slab decomposition is defined for 3-D arrays (last axis is split)");
    const int ax = plan->ndims - 1;
    DEO_REQUIRE(!plan->padded[ax], "dist plan: the decomposed axis must carry a boundary condition, not a pre-padded ghost layer");
    DEO_REQUIRE(plan->bc[ax].d.kind != DEO_BC_PERIODIC, "dist plan: PeriodicBC along the decomposed axis is not supported");
    int64_t start = 0, count = 0;
    rc = deo_dist_slab(plan->dims[ax], nranks, rank, &start, &count);
    if (rc) return rc;
    const int halo = axis_reach(plan.get(), ax);
    // every rank must own enough planes for its neighbours' halos, the one-sided boundary stencils and the BC stencils
    int need = halo > 1 ? halo : 1;
    for (const HostOp& h : plan->ops)
        if (h.d.axis == ax) need = need > h.d.boundary_stencil_length ? need : h.d.boundary_stencil_length;
    need = need > plan->bc[ax].d.K_l ? need : plan->bc[ax].d.K_l;
    need = need > plan->bc[ax].d.K_r ? need : plan->bc[ax].d.K_r;
    DEO_REQUIRE(count >= need, "dist plan: slab of %lld planes is thinner than the %d planes the stencils need", (long long)count, need);
    DEO_REQUIRE(plan->dims[ax] > 2 * (DEO_MAX_TAPS + 1), "dist plan: decomposed axis too short");
    plan->slab_axis = ax;
    plan->slab_start = start;
    plan->slab_count = count;
    plan->halo = halo;
    plan->rank = rank;
    plan->nranks = nranks;
    plan->dist = ctx;
    rc = finalize_plan(plan.get());
    if (rc) return rc;
    plan->launches_per_apply *= (count > 2LL * halo && nranks > 1) ? 3 : 1;
    *out = plan.release();
    return DEO_OK;
}

int32_t dist_apply(deo_plan* plan, deo_buffer* du, deo_buffer* u) {
    DEO_REQUIRE(plan && du && u, "deo_dist_plan_apply: null argument");
    DEO_REQUIRE(plan->slab_axis >= 0, "deo_dist_plan_apply: not a slab plan");
    DEO_REQUIRE(u->bytes >= plan->in_elems() * plan->elem(), "deo_dist_plan_apply: field buffer holds %zu bytes, needs %zu (own planes + 2*halo)",
                u->bytes, plan->in_elems() * plan->elem());
    DEO_REQUIRE(du->bytes >= plan->out_elems() * plan->elem(), "deo_dist_plan_apply: output buffer too small");
    Runtime& R = rt();
    const long long cnt = plan->slab_count;
    const int H = plan->halo;
    deo_dist* ctx = plan->dist;
    // The halo planes outside a physical face are never part of a result, but the tiled kernel streams them
    // through its pipeline (multiplied by zero weights at most): keep them finite.
    if (H > 0) {
        const size_t plane_b = (size_t)plan->in_dim(0) * (size_t)plan->in_dim(1) * plan->elem();
        if (plan->rank == 0) DEO_CUDA(cudaMemsetAsync(u->ptr, 0, (size_t)H * plane_b, R.stream));
        if (plan->rank == plan->nranks - 1)
            DEO_CUDA(cudaMemsetAsync((char*)u->ptr + (size_t)(cnt + H) * plane_b, 0, (size_t)H * plane_b, R.stream));
    }
    if (!ctx || plan->nranks == 1 || H == 0) {   // local emulation or single rank: halos are the caller's business
        g_launches += 1;
        return launch_plan(plan, du->ptr, u->ptr, 0, cnt, R.stream);
    }
    NcclApi& N = nccl();
    const size_t plane = (size_t)plan->in_dim(0) * (size_t)plan->in_dim(1);
    const size_t es = plan->elem();
    const ncclDataType_t dt = plan->dtype == DEO_F64 ? ncclFloat64 : ncclFloat32;
    char* base = (char*)u->ptr;
    const int lo = plan->rank - 1, hi = plan->rank + 1;
    const bool has_lo = lo >= 0, has_hi = hi < plan->nranks;
    // exchange on the communication stream, after everything already queued on the compute stream (u may be its output)
    DEO_CUDA(cudaEventRecord(ctx->ev_ready, R.stream));
    DEO_CUDA(cudaStreamWaitEvent(R.comm_stream, ctx->ev_ready, 0));
    DEO_NCCL(N.GroupStart());
    if (has_lo) {
        DEO_NCCL(N.Send(base + (size_t)H * plane * es, (size_t)H * plane, dt, lo, ctx->comm, R.comm_stream));          // my first H own planes
        DEO_NCCL(N.Recv(base, (size_t)H * plane, dt, lo, ctx->comm, R.comm_stream));                                   // low halo
    }
    if (has_hi) {
        DEO_NCCL(N.Send(base + (size_t)cnt * plane * es, (size_t)H * plane, dt, hi, ctx->comm, R.comm_stream));        // my last H own planes
        DEO_NCCL(N.Recv(base + (size_t)(cnt + H) * plane * es, (size_t)H * plane, dt, hi, ctx->comm, R.comm_stream));  // high halo
    }
    DEO_NCCL(N.GroupEnd());
    // One launch for the whole slab when the tiled kernel can take it: the chunks that read halo planes are scheduled
    // last and wait in the kernel for the flag the communication stream publishes here.
    if (plan->star && !getenv("DEO_DIST_NO_FUSED")) {
        const int step = ++ctx->halo_step;
        k_publish_halo<<<1, 1, 0, R.comm_stream>>>(ctx->halo_flag, step);
        DEO_CUDA(cudaGetLastError());
        const int32_t frc = launch_star_fused(plan, du->ptr, u->ptr, cnt, R.stream, ctx->halo_flag, step, (has_lo ? 1 : 0) | (has_hi ? 2 : 0));
        if (frc == DEO_OK) {
            g_launches += 1;
            // the next exchange may overwrite the halo planes only after this launch; nothing else to order
            return DEO_OK;
        }
        if (frc != DEO_ERR_UNSUPPORTED) return frc;
    }
    DEO_CUDA(cudaEventRecord(ctx->ev_halo, R.comm_stream));
    // planes that read no halo run concurrently with the exchange
    const long long z_lo = has_lo ? (H < cnt ? H : cnt) : 0;
    const long long z_hi = has_hi ? (cnt - H > z_lo ? cnt - H : z_lo) : cnt;
    int32_t rc = launch_plan(plan, du->ptr, u->ptr, z_lo, z_hi, R.stream);
    if (rc) return rc;
    DEO_CUDA(cudaStreamWaitEvent(R.stream, ctx->ev_halo, 0));
    rc = launch_plan(plan, du->ptr, u->ptr, 0, z_lo, R.stream);
    if (rc) return rc;
    rc = launch_plan(plan, du->ptr, u->ptr, z_hi, cnt, R.stream);
    g_launches += (z_hi > z_lo) + (z_lo > 0) + (cnt > z_hi);
    return rc;
}

}  // namespace

extern "C" {

int32_t deo_dist_slab(int64_t n_last, int32_t nranks, int32_t rank, int64_t* start, int64_t* count) {
    DEO_REQUIRE(start && count && nranks >= 1 && rank >= 0 && rank < nranks && n_last >= nranks, "deo_dist_slab: bad arguments");
    const int64_t base = n_last / nranks, rem = n_last % nranks;
    *start = rank * base + (rank < rem ? rank : rem);
    *count = base + (rank < rem ? 1 : 0);
    return DEO_OK;
}

int32_t deo_dist_unique_id(void* id_bytes) {
    DEO_REQUIRE(id_bytes != nullptr, "deo_dist_unique_id: null argument");
    NcclApi& N = nccl();
    if (!N.ok) { set_error("NCCL (libnccl.so.2) could not be loaded: %s", dlerror()); return DEO_ERR_NCCL; }
    ncclUniqueId id;
    DEO_NCCL(N.GetUniqueId(&id));
    memcpy(id_bytes, &id, sizeof id);
    return DEO_OK;
}

int32_t deo_dist_init(const void* id_bytes, int32_t rank, int32_t nranks, deo_dist** out) {
    DEO_REQUIRE(id_bytes && out, "deo_dist_init: null argument");
    *out = nullptr;
    int32_t rc = ensure_init();
    if (rc) return rc;
    NcclApi& N = nccl();
    if (!N.ok) { set_error("NCCL (libnccl.so.2) could not be loaded"); return DEO_ERR_NCCL; }
    std::unique_ptr<deo_dist> ctx(new (std::nothrow) deo_dist());
    if (!ctx) { set_error("out of host memory"); return DEO_ERR_NOMEM; }
    ncclUniqueId id;
    memcpy(&id, id_bytes, sizeof id);
    DEO_NCCL(N.CommInitRank(&ctx->comm, nranks, id, rank));
    ctx->rank = rank;
    ctx->nranks = nranks;
    DEO_CUDA(cudaEventCreateWithFlags(&ctx->ev_ready, cudaEventDisableTiming));
    DEO_CUDA(cudaEventCreateWithFlags(&ctx->ev_halo, cudaEventDisableTiming));
    DEO_CUDA(cudaMalloc(&ctx->halo_flag, sizeof(int)));
    DEO_CUDA(cudaMemset(ctx->halo_flag, 0, sizeof(int)));
    *out = ctx.release();
    return DEO_OK;
}

int32_t deo_dist_destroy(deo_dist* ctx) {
    if (!ctx) return DEO_OK;
    deo_sync();
    if (ctx->comm) nccl().CommDestroy(ctx->comm);
    if (ctx->ev_ready) cudaEventDestroy(ctx->ev_ready);
    if (ctx->ev_halo) cudaEventDestroy(ctx->ev_halo);
    if (ctx->halo_flag) cudaFree(ctx->halo_flag);
    delete ctx;
    return DEO_OK;
}

int32_t deo_dist_plan_create(deo_dist* ctx, const deo_plan_desc* global_desc, deo_plan** out) {
    DEO_REQUIRE(ctx != nullptr, "deo_dist_plan_create: null context");
    return make_slab_plan(global_desc, ctx->rank, ctx->nranks, ctx, out);
}

int32_t deo_dist_plan_create_local(const deo_plan_desc* global_desc, int32_t rank, int32_t nranks, deo_plan** out) {
    return make_slab_plan(global_desc, rank, nranks, nullptr, out);
}

int32_t deo_dist_plan_halo(const deo_plan* plan, int32_t* halo) {
    DEO_REQUIRE(plan && halo, "deo_dist_plan_halo: null argument");
    *halo = plan->halo;
    return DEO_OK;
}

int32_t deo_dist_plan_apply(deo_plan* plan, deo_buffer* du, deo_buffer* u) { return dist_apply(plan, du, u); }

int32_t deo_dist_plan_time(deo_plan* plan, deo_buffer* du, deo_buffer* u, int32_t reps, float* ms_per_apply) {
    DEO_REQUIRE(reps >= 1 && ms_per_apply, "deo_dist_plan_time: bad arguments");
    cudaStream_t s = rt().stream;
    cudaEvent_t a, b;
    DEO_CUDA(cudaEventCreate(&a));
    DEO_CUDA(cudaEventCreate(&b));
    DEO_CUDA(cudaStreamSynchronize(s));
    DEO_CUDA(cudaStreamSynchronize(rt().comm_stream));
    DEO_CUDA(cudaEventRecord(a, s));
    int32_t rc = DEO_OK;
    for (int i = 0; i < reps && rc == DEO_OK; ++i) rc = dist_apply(plan, du, u);
    if (rc) return rc;
    DEO_CUDA(cudaEventRecord(b, s));
    DEO_CUDA(cudaEventSynchronize(b));
    DEO_CUDA(cudaStreamSynchronize(rt().comm_stream));
    float ms = 0;
    DEO_CUDA(cudaEventElapsedTime(&ms, a, b));
    cudaEventDestroy(a);
    cudaEventDestroy(b);
    *ms_per_apply = ms / reps;
    return DEO_OK;
}

}  // extern "C"
