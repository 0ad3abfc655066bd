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
#include <cuda.h>
#include <dlfcn.h>

#include <cstdlib>
#include <map>

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
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
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
    api.AllReduce = (decltype(api.AllReduce))sym("ncclAllReduce");
    api.GetErrorString = (decltype(api.GetErrorString))sym("ncclGetErrorString");
    api.ok = api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.Send && api.Recv && api.GroupStart && api.GroupEnd && api.AllReduce;
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

// Flag words of one rank (device memory, exported to both slab neighbours through CUDA IPC).  All values are
// application indices (monotonic per context).
enum { F_HALO_FROM_LOW = 0,   // the low neighbour's planes for application i have landed in my low halo
       F_HALO_FROM_HIGH = 1,  // same from the high neighbour
       F_LOW_DONE = 2,        // the low neighbour has finished application i (its halo planes may be overwritten)
       F_HIGH_DONE = 3,
       F_STEPWORD = 4,        // staging words: values copied into the neighbours' flags by the copy engines
       F_ACKWORD = 5,
       F_COUNT = 16 };

struct PeerField { void* lo = nullptr; void* hi = nullptr; bool ok = false; };

struct deo_dist {
    ncclComm_t comm = nullptr;
    int rank = 0, nranks = 1;
    cudaEvent_t ev_ready = nullptr, ev_halo = nullptr;
    // peer-to-peer halo exchange over NVLink (copy engines, no kernels): see dist_apply
    bool p2p = false;
    int* flags = nullptr;                  // F_COUNT ints
    int *lo_flags = nullptr, *hi_flags = nullptr;   // the neighbours' flag arrays, mapped
    unsigned char* xbuf = nullptr;         // 3 x 64 bytes device staging for the handle exchange
    int* abuf = nullptr;                   // 2 ints: agreement all-reduce
    std::map<void*, PeerField> fields;     // my field buffer -> the neighbours' mappings of theirs
    int step = 0;
};

namespace deo {
// live contexts, so that freeing a field buffer drops the neighbours' mappings registered under its address
static std::vector<deo_dist*> g_contexts;
void dist_forget_buffer(void* ptr) {
    for (deo_dist* c : g_contexts) {
        auto it = c->fields.find(ptr);
        if (it == c->fields.end()) continue;
        cudaStreamSynchronize(rt().comm_stream);
        if (it->second.lo) cudaIpcCloseMemHandle(it->second.lo);
        if (it->second.hi) cudaIpcCloseMemHandle(it->second.hi);
        c->fields.erase(it);
    }
}
}  // namespace deo

namespace {

typedef CUresult (*PFN_streamValue32)(CUstream, CUdeviceptr, cuuint32_t, unsigned int);
PFN_streamValue32 stream_op(const char* name) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess) return (PFN_streamValue32)p;
    return nullptr;
}
PFN_streamValue32 wait_value32() { static PFN_streamValue32 f = stream_op("cuStreamWaitValue32"); return f; }
PFN_streamValue32 write_value32() { static PFN_streamValue32 f = stream_op("cuStreamWriteValue32"); return f; }

// Collective over ALL ranks: true only where `mine` is true on every rank.  Every decision that selects a schedule
// (peer-to-peer pushes + flags vs. NCCL send/recv) goes through this, so two neighbours can never run different
// protocols against each other.
int32_t agree(deo_dist* ctx, bool mine, bool* all) {
    NcclApi& N = nccl();
    cudaStream_t cs = rt().comm_stream;
    int v = mine ? 1 : 0, out = 0;
    DEO_CUDA(cudaMemcpyAsync(ctx->abuf, &v, sizeof(int), cudaMemcpyHostToDevice, cs));
    DEO_NCCL(N.AllReduce(ctx->abuf, ctx->abuf + 1, 1, ncclInt32, 3 /* ncclMin */, ctx->comm, cs));
    DEO_CUDA(cudaMemcpyAsync(&out, ctx->abuf + 1, sizeof(int), cudaMemcpyDeviceToHost, cs));
    DEO_CUDA(cudaStreamSynchronize(cs));
    *all = out == 1;
    return DEO_OK;
}

// Neighbours swap one 64-byte CUDA IPC handle each way (through NCCL, the only channel the library has); collective
// over the slab neighbours, synchronous.  lo/hi are left untouched where there is no neighbour.
int32_t swap_handles(deo_dist* ctx, const cudaIpcMemHandle_t& mine, cudaIpcMemHandle_t* lo, cudaIpcMemHandle_t* hi) {
    NcclApi& N = nccl();
    cudaStream_t cs = rt().comm_stream;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "ipc handle size");
    DEO_CUDA(cudaMemcpyAsync(ctx->xbuf, &mine, 64, cudaMemcpyHostToDevice, cs));
    const bool has_lo = ctx->rank > 0, has_hi = ctx->rank + 1 < ctx->nranks;
    DEO_NCCL(N.GroupStart());
    if (has_lo) { DEO_NCCL(N.Send(ctx->xbuf, 64, ncclUint8, ctx->rank - 1, ctx->comm, cs)); DEO_NCCL(N.Recv(ctx->xbuf + 64, 64, ncclUint8, ctx->rank - 1, ctx->comm, cs)); }
    if (has_hi) { DEO_NCCL(N.Send(ctx->xbuf, 64, ncclUint8, ctx->rank + 1, ctx->comm, cs)); DEO_NCCL(N.Recv(ctx->xbuf + 128, 64, ncclUint8, ctx->rank + 1, ctx->comm, cs)); }
    DEO_NCCL(N.GroupEnd());
    unsigned char host[192];
    DEO_CUDA(cudaMemcpyAsync(host, ctx->xbuf, 192, cudaMemcpyDeviceToHost, cs));
    DEO_CUDA(cudaStreamSynchronize(cs));
    if (has_lo) memcpy(lo, host + 64, 64);
    if (has_hi) memcpy(hi, host + 128, 64);
    return DEO_OK;
}

// Maps the neighbours' copies of one exported allocation.  Collective over ALL ranks: the handle swap always takes
// place (a rank whose export failed sends zeros) and the outcome is agreed on, so *mapped is the same everywhere.
// A non-zero return is a broken communicator (fatal), not a failed mapping.
int32_t map_neighbours(deo_dist* ctx, void* mine, void** lo, void** hi, bool* mapped) {
    cudaIpcMemHandle_t hm, hl, hh;
    bool good = mine != nullptr && cudaIpcGetMemHandle(&hm, mine) == cudaSuccess;
    if (!good) { memset(&hm, 0, sizeof hm); cudaGetLastError(); }
    int32_t rc = swap_handles(ctx, hm, &hl, &hh);
    if (rc) return rc;
    *lo = *hi = nullptr;
    bool peers = false;
    rc = agree(ctx, good, &peers);            // opening a zeroed handle is pointless: first agree that every export worked
    if (rc) return rc;
    if (peers) {
        if (ctx->rank > 0 && cudaIpcOpenMemHandle(lo, hl, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { good = false; *lo = nullptr; }
        if (ctx->rank + 1 < ctx->nranks && cudaIpcOpenMemHandle(hi, hh, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { good = false; *hi = nullptr; }
        if (!good) cudaGetLastError();
        rc = agree(ctx, good, &peers);
        if (rc) return rc;
    }
    if (!peers) {
        if (*lo) cudaIpcCloseMemHandle(*lo);
        if (*hi) cudaIpcCloseMemHandle(*hi);
        *lo = *hi = nullptr;
    }
    *mapped = peers;
    return DEO_OK;
}

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

#define DEO_DRV(call, what)                                                                  \
    do {                                                                                     \
        if ((call) != CUDA_SUCCESS) { set_error("%s failed", what); return DEO_ERR_CUDA; }  \
    } while (0)

// The neighbours' mappings of field buffer `ptr` (collective registration on first use).  pf->ok is the same on every
// rank: false means "use the NCCL schedule for this buffer".
int32_t peer_field(deo_dist* ctx, void* ptr, const PeerField** pf) {
    auto it = ctx->fields.find(ptr);
    if (it == ctx->fields.end()) {
        DEO_CUDA(cudaStreamSynchronize(rt().stream));
        PeerField f;
        int32_t rc = map_neighbours(ctx, ptr, &f.lo, &f.hi, &f.ok);
        if (rc) return rc;
        it = ctx->fields.emplace(ptr, f).first;
    }
    *pf = &it->second;
    return DEO_OK;
}

// Peer-to-peer halo exchange of application `step`, enqueued on the communication stream: my first / last H own planes
// go straight into the neighbours' field buffers (mapped through CUDA IPC) over NVLink by the copy engines, each followed
// by a 4-byte flag.  `ready_lo` / `ready_hi`: events after which my first / last H own planes are valid.  Before
// overwriting a neighbour's halo planes the stream waits (cuStreamWaitValue32) until that neighbour has reported the
// previous application finished.  No kernel takes part, so the exchange cannot compete with the stencil kernel for SMs.
int32_t push_halos(deo_dist* ctx, const deo_plan* plan, const PeerField& pf, char* base, int step, cudaEvent_t ready_lo, cudaEvent_t ready_hi) {
    Runtime& R = rt();
    const int H = plan->halo;
    const long long cnt = plan->slab_count;
    const size_t plane_b = (size_t)plan->in_dim(0) * (size_t)plan->in_dim(1) * plan->elem();
    const int lo = plan->rank - 1, hi = plan->rank + 1;
    const bool has_lo = lo >= 0, has_hi = hi < plan->nranks;
    CUstream cs = (CUstream)R.comm_stream;
    DEO_DRV(write_value32()(cs, (CUdeviceptr)(ctx->flags + F_STEPWORD), (cuuint32_t)step, CU_STREAM_WRITE_VALUE_DEFAULT), "cuStreamWriteValue32");
    if (has_lo) {
        int64_t s_lo = 0, c_lo = 0;
        deo_dist_slab(plan->dims[plan->slab_axis], plan->nranks, lo, &s_lo, &c_lo);
        DEO_CUDA(cudaStreamWaitEvent(R.comm_stream, ready_lo, 0));
        DEO_DRV(wait_value32()(cs, (CUdeviceptr)(ctx->flags + F_LOW_DONE), (cuuint32_t)(step - 1), CU_STREAM_WAIT_VALUE_GEQ), "cuStreamWaitValue32");
        DEO_CUDA(cudaMemcpyAsync((char*)pf.lo + (size_t)(H + c_lo) * plane_b, base + (size_t)H * plane_b, (size_t)H * plane_b, cudaMemcpyDefault, R.comm_stream));
        DEO_CUDA(cudaMemcpyAsync(ctx->lo_flags + F_HALO_FROM_HIGH, ctx->flags + F_STEPWORD, sizeof(int), cudaMemcpyDefault, R.comm_stream));
    }
    if (has_hi) {
        DEO_CUDA(cudaStreamWaitEvent(R.comm_stream, ready_hi, 0));
        DEO_DRV(wait_value32()(cs, (CUdeviceptr)(ctx->flags + F_HIGH_DONE), (cuuint32_t)(step - 1), CU_STREAM_WAIT_VALUE_GEQ), "cuStreamWaitValue32");
        DEO_CUDA(cudaMemcpyAsync((char*)pf.hi, base + (size_t)cnt * plane_b, (size_t)H * plane_b, cudaMemcpyDefault, R.comm_stream));
        DEO_CUDA(cudaMemcpyAsync(ctx->hi_flags + F_HALO_FROM_LOW, ctx->flags + F_STEPWORD, sizeof(int), cudaMemcpyDefault, R.comm_stream));
    }
    return DEO_OK;
}

// Tells the neighbours (after everything queued on stream `s`) that application `step` is finished here: they may
// overwrite my halo planes.
int32_t ack_done(deo_dist* ctx, const deo_plan* plan, int step, cudaStream_t s) {
    const bool has_lo = plan->rank > 0, has_hi = plan->rank + 1 < plan->nranks;
    DEO_DRV(write_value32()((CUstream)s, (CUdeviceptr)(ctx->flags + F_ACKWORD), (cuuint32_t)step, CU_STREAM_WRITE_VALUE_DEFAULT), "cuStreamWriteValue32");
    if (has_lo) DEO_CUDA(cudaMemcpyAsync(ctx->lo_flags + F_HIGH_DONE, ctx->flags + F_ACKWORD, sizeof(int), cudaMemcpyDefault, s));
    if (has_hi) DEO_CUDA(cudaMemcpyAsync(ctx->hi_flags + F_LOW_DONE, ctx->flags + F_ACKWORD, sizeof(int), cudaMemcpyDefault, s));
    return DEO_OK;
}

// The halo planes outside a physical face are never part of a result, but the tiled kernel streams them through its
// pipeline (multiplied by zero weights at most): keep them finite.
int32_t clear_outer_halos(const deo_plan* plan, void* u, cudaStream_t s) {
    const int H = plan->halo;
    if (H <= 0) return DEO_OK;
    const size_t plane_b = (size_t)plan->in_dim(0) * (size_t)plan->in_dim(1) * plan->elem();
    if (plan->rank == 0) DEO_CUDA(cudaMemsetAsync(u, 0, (size_t)H * plane_b, s));
    if (plan->rank == plan->nranks - 1)
        DEO_CUDA(cudaMemsetAsync((char*)u + (size_t)(plan->slab_count + H) * plane_b, 0, (size_t)H * plane_b, s));
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
    int32_t rc = clear_outer_halos(plan, u->ptr, R.stream);
    if (rc) return rc;
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
    // ---- fused schedule: peer-to-peer halo pushes by the copy engines + ONE kernel launch for the whole slab ----------
    // The stencil kernel's first / last march-axis chunks are scheduled last and wait in the kernel for the flag
    // (kernel_star.cuh).  Every rank takes the same decision: global extents, agreed p2p state, agreed buffer mapping.
    {
        const StarLimits lim = star_limits(plan);
        const long long min_cnt = plan->dims[plan->slab_axis] / plan->nranks;
        if (ctx->p2p && plan->star && u->owned && lim.fusable && min_cnt >= lim.min_fused_planes && !getenv("DEO_DIST_NO_FUSED")) {
            const PeerField* pf = nullptr;
            rc = peer_field(ctx, u->ptr, &pf);
            if (rc) return rc;
            if (pf->ok) {
                const int step = ++ctx->step;
                DEO_CUDA(cudaEventRecord(ctx->ev_ready, R.stream));
                rc = push_halos(ctx, plan, *pf, base, step, ctx->ev_ready, ctx->ev_ready);
                if (rc) return rc;
                rc = launch_star_fused(plan, du->ptr, u->ptr, cnt, R.stream, ctx->flags, step, (has_lo ? 1 : 0) | (has_hi ? 2 : 0));
                if (rc) return rc;
                g_launches += 1;
                return ack_done(ctx, plan, step, R.stream);
            }
        }
    }
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
    DEO_CUDA(cudaEventRecord(ctx->ev_halo, R.comm_stream));
    // planes that read no halo run concurrently with the exchange
    const long long z_lo = has_lo ? (H < cnt ? H : cnt) : 0;
    const long long z_hi = has_hi ? (cnt - H > z_lo ? cnt - H : z_lo) : cnt;
    rc = launch_plan(plan, du->ptr, u->ptr, z_lo, z_hi, R.stream);
    if (rc) return rc;
    DEO_CUDA(cudaStreamWaitEvent(R.stream, ctx->ev_halo, 0));
    rc = launch_plan(plan, du->ptr, u->ptr, 0, z_lo, R.stream);
    if (rc) return rc;
    rc = launch_plan(plan, du->ptr, u->ptr, z_hi, cnt, R.stream);
    g_launches += (z_hi > z_lo) + (z_lo > 0) + (cnt > z_hi);
    return rc;
}

// Host-buffer form of mul! on a slab: u_own / du_host hold this rank's `count` planes.  Chunks of planes go through
// upload | kernel | download on three streams like deo_plan_apply_host; the two edge chunks are uploaded first so that
// the halo pushes to the neighbours start at once, and computed last, behind a stream wait on the neighbours' flags.
int32_t dist_apply_host(deo_plan* plan, void* du_host, const void* u_own) {
    DEO_REQUIRE(plan && du_host && u_own, "deo_dist_plan_apply_host: null argument");
    DEO_REQUIRE(plan->slab_axis >= 0, "deo_dist_plan_apply_host: not a slab plan");
    DEO_REQUIRE(plan->dist != nullptr || plan->nranks == 1, "deo_dist_plan_apply_host: the plan has no communicator (deo_dist_plan_create_local plans take device buffers)");
    Runtime& R = rt();
    deo_dist* ctx = plan->dist;
    const size_t in_b = plan->in_elems() * plan->elem(), out_b = plan->out_elems() * plan->elem();
    int32_t rc = DEO_OK;
    if (!plan->host_u) { rc = deo_buffer_create(in_b, &plan->host_u); if (rc) return rc; }
    if (!plan->host_du) { rc = deo_buffer_create(out_b, &plan->host_du); if (rc) return rc; }
    deo_buffer *u = plan->host_u, *du = plan->host_du;
    const int H = plan->halo;
    const long long cnt = plan->slab_count;
    const size_t plane_b = (size_t)plan->in_dim(0) * (size_t)plan->in_dim(1) * plan->elem();   // in and out planes are equal (no x/y padding)
    char* ub = (char*)u->ptr + (size_t)H * plane_b;                                           // own planes start here
    const bool has_lo = plan->rank > 0, has_hi = plan->rank + 1 < plan->nranks;
    // chunking (same rule on every rank would not be needed: the exchange only involves the edge planes)
    int reach = 0;
    for (const HostOp& h : plan->ops)
        if (h.d.axis == plan->slab_axis) reach = reach > h.d.boundary_stencil_length ? reach : h.d.boundary_stencil_length;
    const size_t chunk_bytes = getenv("DEO_HOST_CHUNK_BYTES") ? (size_t)atoll(getenv("DEO_HOST_CHUNK_BYTES")) : ((size_t)96 << 20);
    long long chunk = (long long)(chunk_bytes / (plane_b ? plane_b : 1));
    if (chunk < 2 * reach + 8) chunk = 2 * reach + 8;
    if (chunk < 2 * H) chunk = 2 * H;
    std::vector<long long> zb;
    for (long long z = 0; z < cnt; z += chunk) zb.push_back(z);
    if (zb.size() > 1 && cnt - zb.back() < 2 * reach + 2) zb.pop_back();
    zb.push_back(cnt);
    const long long nchunks = (long long)zb.size() - 1;
    const PeerField* pf = nullptr;
    bool p2p = ctx && plan->nranks > 1 && H > 0 && ctx->p2p;
    if (p2p) {
        rc = peer_field(ctx, u->ptr, &pf);
        if (rc) return rc;
        p2p = pf->ok;
    }
    const bool single = !ctx || plan->nranks == 1 || H == 0;
    if (nchunks < 3 || !(p2p || single)) {
        // serial: upload, the device-resident application (whatever exchange schedule it picks), download
        DEO_CUDA(cudaMemcpyAsync(ub, u_own, out_b, cudaMemcpyHostToDevice, R.stream));
        rc = dist_apply(plan, du, u);
        if (rc) return rc;
        DEO_CUDA(cudaMemcpyAsync(du_host, du->ptr, out_b, cudaMemcpyDeviceToHost, R.stream));
        DEO_CUDA(cudaStreamSynchronize(R.stream));
        DEO_CUDA(cudaStreamSynchronize(R.comm_stream));
        return DEO_OK;
    }
    while ((long long)plan->host_ev.size() < 2 * nchunks) {
        cudaEvent_t e;
        DEO_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        plan->host_ev.push_back(e);
    }
    DEO_CUDA(cudaStreamSynchronize(R.stream));
    rc = clear_outer_halos(plan, u->ptr, R.h2d_stream);
    if (rc) return rc;
    // upload order: first chunk, last chunk, then the middle ones
    std::vector<long long> order;
    order.push_back(0);
    order.push_back(nchunks - 1);
    for (long long k = 1; k + 1 < nchunks; ++k) order.push_back(k);
    for (long long k : order) {
        const long long z0 = zb[(size_t)k], z1 = zb[(size_t)k + 1];
        DEO_CUDA(cudaMemcpyAsync(ub + (size_t)z0 * plane_b, (const char*)u_own + (size_t)z0 * plane_b, (size_t)(z1 - z0) * plane_b, cudaMemcpyHostToDevice, R.h2d_stream));
        DEO_CUDA(cudaEventRecord(plan->host_ev[(size_t)k], R.h2d_stream));
    }
    int step = 0;
    if (!single) {
        step = ++ctx->step;
        rc = push_halos(ctx, plan, *pf, (char*)u->ptr, step, plan->host_ev[0], plan->host_ev[(size_t)nchunks - 1]);
        if (rc) return rc;
    }
    // kernels: middle chunks in order (chunk k reads into chunks k-1 and k+1), then the two edge chunks
    std::vector<long long> korder;
    for (long long k = 1; k + 1 < nchunks; ++k) korder.push_back(k);
    korder.push_back(0);
    korder.push_back(nchunks - 1);
    for (long long k : korder) {
        const long long z0 = zb[(size_t)k], z1 = zb[(size_t)k + 1];
        const long long dep = k + 1 < nchunks - 1 ? k + 1 : k;               // uploads are in order: the latest middle chunk read covers the earlier ones
        DEO_CUDA(cudaStreamWaitEvent(R.stream, plan->host_ev[(size_t)(k == 0 ? (nchunks > 2 ? 1 : 0) : (k == nchunks - 1 ? nchunks - 2 : dep))], 0));
        if (k == nchunks - 1) DEO_CUDA(cudaStreamWaitEvent(R.stream, plan->host_ev[(size_t)k], 0));
        if (!single && k == 0 && has_lo)
            DEO_DRV(wait_value32()((CUstream)R.stream, (CUdeviceptr)(ctx->flags + F_HALO_FROM_LOW), (cuuint32_t)step, CU_STREAM_WAIT_VALUE_GEQ), "cuStreamWaitValue32");
        if (!single && k == nchunks - 1 && has_hi)
            DEO_DRV(wait_value32()((CUstream)R.stream, (CUdeviceptr)(ctx->flags + F_HALO_FROM_HIGH), (cuuint32_t)step, CU_STREAM_WAIT_VALUE_GEQ), "cuStreamWaitValue32");
        rc = launch_plan(plan, du->ptr, u->ptr, z0, z1, R.stream);
        if (rc) return rc;
        g_launches += 1;
        DEO_CUDA(cudaEventRecord(plan->host_ev[(size_t)(nchunks + k)], R.stream));
        DEO_CUDA(cudaStreamWaitEvent(R.d2h_stream, plan->host_ev[(size_t)(nchunks + k)], 0));
        DEO_CUDA(cudaMemcpyAsync((char*)du_host + (size_t)z0 * plane_b, (const char*)du->ptr + (size_t)z0 * plane_b, (size_t)(z1 - z0) * plane_b,
                                 cudaMemcpyDeviceToHost, R.d2h_stream));
    }
    if (!single) {
        rc = ack_done(ctx, plan, step, R.stream);
        if (rc) return rc;
    }
    DEO_CUDA(cudaStreamSynchronize(R.d2h_stream));
    DEO_CUDA(cudaStreamSynchronize(R.stream));
    DEO_CUDA(cudaStreamSynchronize(R.comm_stream));
    return DEO_OK;
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
    DEO_CUDA(cudaMalloc(&ctx->abuf, 2 * sizeof(int)));
    // peer-to-peer exchange state.  Every step that can fail on one rank only (allocation, IPC export / import, missing
    // stream memory operations) is followed by an agreement over all ranks, so either every rank ends with p2p == true
    // or every rank leaves the NCCL send/recv schedule in charge.
    if (nranks > 1) {
        bool ok = !getenv("DEO_DIST_NO_P2P") && wait_value32() && write_value32();
        ok = ok && cudaMalloc(&ctx->flags, F_COUNT * sizeof(int)) == cudaSuccess && cudaMemset(ctx->flags, 0, F_COUNT * sizeof(int)) == cudaSuccess &&
             cudaMalloc(&ctx->xbuf, 192) == cudaSuccess && cudaDeviceSynchronize() == cudaSuccess;
        if (!ok) cudaGetLastError();
        bool all = false;
        rc = agree(ctx.get(), ok, &all);
        if (rc) return rc;
        if (all) {
            void *lf = nullptr, *hf = nullptr;
            bool mapped = false;
            rc = map_neighbours(ctx.get(), ctx->flags, &lf, &hf, &mapped);
            if (rc) return rc;
            ctx->lo_flags = (int*)lf;
            ctx->hi_flags = (int*)hf;
            ctx->p2p = mapped;
        }
    }
    *out = ctx.release();
    g_contexts.push_back(*out);
    return DEO_OK;
}

int32_t deo_dist_destroy(deo_dist* ctx) {
    if (!ctx) return DEO_OK;
    deo_sync();
    for (size_t i = 0; i < g_contexts.size(); ++i) if (g_contexts[i] == ctx) { g_contexts.erase(g_contexts.begin() + (long)i); break; }
    if (ctx->comm) nccl().CommDestroy(ctx->comm);
    if (ctx->ev_ready) cudaEventDestroy(ctx->ev_ready);
    if (ctx->ev_halo) cudaEventDestroy(ctx->ev_halo);
    for (auto& kv : ctx->fields) { if (kv.second.lo) cudaIpcCloseMemHandle(kv.second.lo); if (kv.second.hi) cudaIpcCloseMemHandle(kv.second.hi); }
    if (ctx->lo_flags) cudaIpcCloseMemHandle(ctx->lo_flags);
    if (ctx->hi_flags) cudaIpcCloseMemHandle(ctx->hi_flags);
    if (ctx->flags) cudaFree(ctx->flags);
    if (ctx->xbuf) cudaFree(ctx->xbuf);
    if (ctx->abuf) cudaFree(ctx->abuf);
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

int32_t deo_dist_plan_apply_host(deo_plan* plan, void* du_host, const void* u_own_host) { return dist_apply_host(plan, du_host, u_own_host); }

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
