// Internal declarations of libdeo_b200 (not part of the C ABI).
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include "../../include/deo_b200.h"

namespace deo {

constexpr int kMaxDims = DEO_MAX_DIMS;
constexpr int kMaxOps = DEO_MAX_OPS;
constexpr int kMaxTaps = DEO_MAX_TAPS;        // interior stencil taps
constexpr int kMaxBTaps = DEO_MAX_TAPS + 1;   // boundary-row taps

// ---- error plumbing -------------------------------------------------------------------------
void set_error(const char* fmt, ...);
const std::string& last_error();
int32_t cuda_fail(cudaError_t e, const char* what, const char* file, int line);

#define DEO_CUDA(call)                                                         \
    do {                                                                       \
        cudaError_t e_ = (call);                                               \
        if (e_ != cudaSuccess) return ::deo::cuda_fail(e_, #call, __FILE__, __LINE__); \
    } while (0)

#define DEO_REQUIRE(cond, ...)                                                 \
    do {                                                                       \
        if (!(cond)) { ::deo::set_error(__VA_ARGS__); return DEO_ERR_INVALID; } \
    } while (0)

// ---- runtime state --------------------------------------------------------------------------
struct Runtime {
    int device = -1;
    cudaStream_t stream = nullptr;        // compute stream of the library
    cudaStream_t comm_stream = nullptr;   // halo-exchange stream
    cudaStream_t h2d_stream = nullptr;    // host-buffer path: upload / download streams of the chunk pipeline
    cudaStream_t d2h_stream = nullptr;
    int sm_count = 0;
    bool ready = false;
};
Runtime& rt();
int32_t ensure_init();
extern std::atomic<long long> g_launches;

}  // namespace deo

struct deo_buffer {
    void* ptr = nullptr;
    size_t bytes = 0;
    bool owned = true;
};

namespace deo {

// ---- device-side plan description (passed by value as a __grid_constant__ kernel parameter,
//      i.e. it lives in the constant bank: stencil weights are constant-memory operands) ----------
enum OpMode : int { MODE_CONST = 0, MODE_SIGNSEL = 1, MODE_TABLE = 2 };

// One explicitly enumerated output row (one-sided boundary stencils and every other row that does
// not follow the interior pattern).  Weights are pre-multiplied by the row's coefficient,
// (c*w) exactly as the reference forms `cur_coeff * cur_stencil[idx]` before touching x.
template <typename T>
struct BRow {
    int start;   // first tap, index into the padded pencil q[0..n+1]
    int ntaps;
    int acc64;   // uniform-upwind BC rows accumulate in Float64 (convolutions.jl:154,:180 `xtempi = 0.0`)
    int pad_;
    T w[kMaxBTaps];
};

template <typename T>
struct DevOp {
    int axis;
    int n;          // global A.len
    int mode;       // OpMode
    int ntaps;      // taps of the interior pattern
    int soff[2];    // first tap relative to the centre q[r+1]; [0]: c>=0 set, [1]: c<0 set
    int nlow;       // rows [0,nlow) and [n-nhigh,n) are explicit BRows
    int nhigh;
    int cshift;     // coefficient index shift of the BoundaryPaddedVector centered methods (convolutions.jl:384)
    int pad_;
    T w[2][kMaxTaps];       // MODE_CONST: [0] = c*w ; MODE_SIGNSEL: raw stencils, multiplied by c[r] per tap
    const T* coeff;         // MODE_SIGNSEL: device c[n]
    const T* table;         // MODE_TABLE: device [n][ntaps] (c*w), rows of the explicit zones unused
    const int* table_soff;  // MODE_TABLE: device [n]
    const BRow<T>* brows;   // device [nlow + nhigh]
};

template <typename T>
struct DevBC {
    int kind;       // DEO_BC_*
    int per_face;
    int K_l, K_r;
    int per_lo, per_hi;     // periodic: global row whose value is the low / high ghost
    const T *a_l, *b_l, *a_r, *b_r;   // device
};

template <typename T>
struct DevPlan {
    int ndims, nops, accumulate, pad_;
    int n_out[kMaxDims];        // local output extent
    int n_glob[kMaxDims];       // global extent (== n_out except along a slab-decomposed axis)
    int row0[kMaxDims];         // global row index of local output row 0
    int in_off[kMaxDims];       // input index = local output index + in_off (1: ghost layer in the input; slab: halo)
    int padded[kMaxDims];       // ghosts come from the input array instead of a BC
    long long in_stride[kMaxDims];
    long long out_stride[kMaxDims];
    DevBC<T> bc[kMaxDims];
    DevOp<T> ops[kMaxOps];
};

// ---- host-side plan ---------------------------------------------------------------------------
struct HostOp {   // deep copy of a deo_op_desc (element type erased into bytes)
    deo_op_desc d;
    std::vector<unsigned char> stencil, low, high, coeff;
};
struct HostBC {
    deo_bc_desc d;
    std::vector<unsigned char> a_l, b_l, a_r, b_r;
};

struct DeviceBlob {   // RAII device allocation
    void* p = nullptr;
    size_t bytes = 0;
    DeviceBlob() = default;
    DeviceBlob(const DeviceBlob&) = delete;
    DeviceBlob& operator=(const DeviceBlob&) = delete;
    ~DeviceBlob() { if (p) cudaFree(p); }
};

struct StarConfig;   // kernel_star.cu

}  // namespace deo

struct deo_dist;

struct deo_plan {
    int dtype = DEO_F64;
    int ndims = 0;
    long long dims[deo::kMaxDims] = {1, 1, 1};      // global output dims
    int padded[deo::kMaxDims] = {0, 0, 0};
    int accumulate = 0;
    int flags = 0;
    std::vector<deo::HostOp> ops;
    deo::HostBC bc[deo::kMaxDims];

    // slab decomposition along the last axis (single GPU: slab == whole axis, halo == 0)
    int slab_axis = -1;
    long long slab_start = 0, slab_count = 0;
    int halo = 0;
    int rank = 0, nranks = 1;
    deo_dist* dist = nullptr;

    // device-side
    std::vector<std::unique_ptr<deo::DeviceBlob>> blobs;
    // (Re)builds reuse the device allocations of the previous build in order (plan_upload) and send their contents with
    // asynchronous copies out of a pinned staging arena, so deo_plan_update_coefficients neither allocates nor synchronises.
    size_t blob_cursor = 0;
    unsigned char* stage = nullptr;
    size_t stage_cap = 0, stage_used = 0, stage_want = 0;
    cudaEvent_t stage_ev = nullptr;
    std::vector<std::vector<unsigned char>> host_brows;   // per operator: host copy of its explicit boundary rows (BRow<T>[])
    std::vector<unsigned char> devplan;     // DevPlan<T> bytes
    std::string kernel = "generic";
    int launches_per_apply = 1;
    std::shared_ptr<void> star;             // StarConfig when the tiled 2-D/3-D kernel is eligible
    std::shared_ptr<void> line;             // LineConfig when the 1-D kernel is eligible

    // host-buffer path (deo_plan_apply_host): device staging buffers kept across calls, pipeline events
    deo_buffer* host_u = nullptr;
    deo_buffer* host_du = nullptr;
    std::vector<cudaEvent_t> host_ev;

    // graph cache for apply_n
    cudaGraphExec_t graph_exec = nullptr;
    const void* graph_u = nullptr;
    void* graph_du = nullptr;
    int graph_reps = 0;

    size_t elem() const { return dtype == DEO_F64 ? 8 : 4; }
    long long local_dim(int a) const { return a == slab_axis ? slab_count : dims[a]; }
    long long in_dim(int a) const {
        return local_dim(a) + (padded[a] ? 2 : 0) + (a == slab_axis ? 2LL * halo : 0);
    }
    size_t in_elems() const { size_t n = 1; for (int a = 0; a < ndims; ++a) n *= (size_t)in_dim(a); return n; }
    size_t out_elems() const { size_t n = 1; for (int a = 0; a < ndims; ++a) n *= (size_t)local_dim(a); return n; }
};

namespace deo {
// plan_build.cu
int32_t build_device_plan(deo_plan* plan);
// Device copy of `bytes` host bytes owned by the plan: reuses the allocation the previous build made at the same position
// when the size matches; the copy is stream-ordered on the library stream.  nullptr + *err on failure.
void* plan_upload(deo_plan* plan, const void* host, size_t bytes, cudaError_t* err);
struct HostRow { int start = 0, ntaps = 0; double w[kMaxBTaps]; };   // row r = sum_k w[k] * q[start + k]
struct RowGenerator {                      // lazily enumerates the rows of one operator of a plan
    int n = 0;
    bool interior_uniform = false;         // uniform grid and constant coefficient: every interior row has the same weights
    virtual ~RowGenerator() {}
    virtual bool ok() const = 0;
    virtual bool row(int r0, HostRow& out) = 0;   // 0-based row
};
std::unique_ptr<RowGenerator> make_row_generator(const deo_plan* plan, int op_index);
// kernel_generic.cu : computes local output planes [z0, z1) of the last axis (whole array when ndims<3 uses z in [0,1))
int32_t launch_generic(const deo_plan* plan, void* du, const void* u, long long z0, long long z1, cudaStream_t s);
// kernel_star.cu
int32_t star_configure(deo_plan* plan);
int32_t launch_star(const deo_plan* plan, void* du, const void* u, long long z0, long long z1, cudaStream_t s, bool explicit_range = false);
struct StarLimits { bool fusable = false; long long min_fused_planes = 0; };
StarLimits star_limits(const deo_plan* plan);
bool star_take_halo_timeout();
bool star_can_axpy(const deo_plan* plan);
void star_set_axpy(const deo_plan* plan, bool on, double dt);
int32_t launch_star_fused(const deo_plan* plan, void* du, const void* u, long long cnt, cudaStream_t s, const int* halo_flag, int expect, int sides);
// dist.cu
void dist_forget_buffer(void* ptr);
// kernel_line.cu
int32_t line_configure(deo_plan* plan);
int32_t launch_line(const deo_plan* plan, void* du, const void* u, cudaStream_t s);
// plan.cu
int32_t launch_plan(const deo_plan* plan, void* du, const void* u, long long z0, long long z1, cudaStream_t s);
}  // namespace deo
