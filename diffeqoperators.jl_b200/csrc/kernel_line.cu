// 1-D kernel (BASELINE config 1: repeated mul! of CenteredDifference * BC on a long vector).
//
// A 1-D application moves 2*sizeof(T) bytes per point and has no reuse to organise, so the kernel is a flat
// vectorised sweep: every thread owns VEC = 16 B / sizeof(T) consecutive rows, loads its 16-byte vector of u plus
// the R halo values on each side through the read-only path, and stores one 16-byte vector of du.  All operators of
// the plan are folded into one per-row stencil exactly like the tiled kernel's TABLE variants (kernel_star.cu):
//   * CONST: every interior row has the same weights (uniform grid, constant coefficient) -> weights in the kernel
//     parameter block (constant bank); with a single operator the arithmetic order equals the per-point kernel's;
//   * TABLE: per-row weights [n][2R+1] streamed from global memory (non-uniform grids, coefficient vectors, upwind
//     with per-row wind direction, sums of operators); these bytes are part of the algorithmic traffic.
// The R rows at each end (one-sided boundary stencils and every row whose window touches a ghost) are evaluated by
// the threads that own them from dense rows of 2R+2 taps; the ghost value b + a.u[edge] (bc_operators.jl:188-191)
// is computed in the kernel, or read from the input when it already carries its ghost layer
// (convolutions.jl:17-22 on a plain padded vector).  The launch is small (a few hundred bytes of parameters), which
// matters here: at N = 1e6 the whole application is ~3 us of L2-resident traffic.
#include <cstdlib>

#include "generic_device.cuh"

namespace deo {

constexpr int kLineMaxK = 8;

template <typename T, int R>
struct LineParams {
    static constexpr int NQ = 2 * R + 1, TB = 2 * R + 2;
    int n, padded, accumulate, table;
    int K_l, K_r, periodic, pad1_;   // periodic: 1-D PeriodicBC, ghosts l = u[end], r = u[1] (bc_operators.jl:192)
    T a_l[kLineMaxK], a_r[kLineMaxK];
    T b_l, b_r;
    T w[NQ];
    const T* tab;            // TABLE: [n][NQ]
    T bw[2][R][TB];          // low rows: tap k <-> q[k]; high rows: tap k <-> q[n+2-TB+k]
};

struct LineConfig {
    int R = 0;
    std::vector<unsigned char> params;
};

template <typename T> struct LVec;
template <> struct LVec<double> { using type = double2; static constexpr int N = 2; };
template <> struct LVec<float> { using type = float4; static constexpr int N = 4; };

template <typename T, int R, bool TABLE>
__device__ __forceinline__ void line_vector(const LineParams<T, R>& S, const T* __restrict__ in, T* __restrict__ du, long long x0) {
    constexpr int VEC = LVec<T>::N, NQ = 2 * R + 1, TB = 2 * R + 2;
    using V = typename LVec<T>::type;
    const int n = S.n;
    const T* u = in + (S.padded ? 1 : 0);                 // u[j] = q[j+1]
    // window u[x0-R .. x0+VEC-1+R], clamped loads (clamped values are only used by rows that are recomputed below)
    T xw[VEC + 2 * R];
    const bool inner = x0 >= R && x0 + VEC + R <= n;
    if (inner && !S.padded) {
        const V v = __ldg(reinterpret_cast<const V*>(u + x0));
        if constexpr (VEC == 2) { xw[R] = v.x; xw[R + 1] = v.y; }
        else { xw[R] = v.x; xw[R + 1] = v.y; xw[R + 2] = v.z; xw[R + 3] = v.w; }
#pragma unroll
        for (int i = 0; i < R; ++i) { xw[i] = __ldg(u + x0 - R + i); xw[R + VEC + i] = __ldg(u + x0 + VEC + i); }
    } else {
#pragma unroll
        for (int i = 0; i < VEC + 2 * R; ++i) {
            long long j = x0 - R + i;
            j = j < 0 ? 0 : (j > n - 1 ? n - 1 : j);
            xw[i] = __ldg(u + j);
        }
    }
    T out[VEC];
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
        T a = T(0);
        if constexpr (TABLE) {
            const long long r = x0 + v < n ? x0 + v : n - 1;
            const T* wr = S.tab + r * NQ;
#pragma unroll
            for (int t = 0; t < NQ; ++t) a = fma_t(__ldg(wr + t), xw[v + t], a);
        } else {
#pragma unroll
            for (int t = 0; t < NQ; ++t) a = fma_t(S.w[t], xw[v + t], a);
        }
        out[v] = a;
    }
    // rows whose window touches a ghost: x < R or x >= n - R
    if (x0 < R || x0 + VEC > n - R) {
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
            const long long x = x0 + v;
            if (x >= n) break;
            const bool low = x < R, high = x >= n - R;
            if (!(low || high)) continue;
            T g;
            if (S.padded) g = __ldg(in + (high ? n + 1 : 0));
            else if (S.periodic) g = __ldg(u + (high ? 0 : n - 1));      // the wrap-around read
            else {
                g = T(0);
                const int K = high ? S.K_r : S.K_l;
                const T* a = high ? S.a_r : S.a_l;
                for (int k = 0; k < K; ++k) g = fma_t(a[k], __ldg(u + (high ? n - K : 0) + k), g);
                g += high ? S.b_r : S.b_l;
            }
            const T* w = S.bw[high ? 1 : 0][high ? (int)(x - (n - R)) : (int)x];
            T res = T(0);
            for (int k = 0; k < TB; ++k) {
                const long long q = high ? (long long)n + 2 - TB + k : k;      // padded-pencil index
                const T val = (q == 0 || q == n + 1) ? g : __ldg(u + q - 1);
                res = fma_t(w[k], val, res);
            }
            out[v] = res;
        }
    }
    if (x0 + VEC <= n) {
        V o;
        if (S.accumulate) {
            const V old = *reinterpret_cast<const V*>(du + x0);
            if constexpr (VEC == 2) { out[0] += old.x; out[1] += old.y; }
            else { out[0] += old.x; out[1] += old.y; out[2] += old.z; out[3] += old.w; }
        }
        if constexpr (VEC == 2) { o.x = out[0]; o.y = out[1]; }
        else { o.x = out[0]; o.y = out[1]; o.z = out[2]; o.w = out[3]; }
        *reinterpret_cast<V*>(du + x0) = o;
    } else {
        for (int v = 0; v < VEC && x0 + v < n; ++v) du[x0 + v] = S.accumulate ? du[x0 + v] + out[v] : out[v];
    }
}

// VPT vectors per thread, a whole grid apart (coalesced): their loads are independent, so a thread has VPT times the bytes in
// flight -- the 1-D application is ~3 us of L2-resident traffic and latency-bound, not bandwidth-bound.
template <typename T, int R, bool TABLE, int VPT>
__global__ void __launch_bounds__(256)
k_line(const __grid_constant__ LineParams<T, R> S, const T* __restrict__ in, T* __restrict__ du) {
    constexpr int VEC = LVec<T>::N;
    const long long stride = (long long)gridDim.x * blockDim.x * VEC;
    const long long x0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * VEC;
    // Programmatic dependent launch: let the next application's CTAs be scheduled while this one runs, and wait -- before
    // the first global access -- until the previous one has completed and flushed.  No-ops without the launch attribute.
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
#pragma unroll
    for (int k = 0; k < VPT; ++k) {
        const long long x = x0 + k * stride;
        if (x < S.n) line_vector<T, R, TABLE>(S, in, du, x);
    }
}

namespace {

template <typename T, int R>
bool fill_line(deo_plan* plan, std::vector<std::unique_ptr<RowGenerator>>& gens, LineConfig& C) {
    using LP = LineParams<T, R>;
    constexpr int NQ = LP::NQ, TB = LP::TB;
    C.params.assign(sizeof(LP), 0);
    LP& S = *reinterpret_cast<LP*>(C.params.data());
    const int n = (int)plan->dims[0];
    S.n = n;
    S.padded = plan->padded[0];
    S.accumulate = plan->accumulate;
    bool uniform = true;
    for (auto& g : gens) uniform = uniform && g->interior_uniform;
    S.table = uniform ? 0 : 1;
    std::vector<double> low((size_t)R * TB, 0.0), high((size_t)R * TB, 0.0), wsum((size_t)NQ, 0.0), tab;
    if (!uniform) tab.assign((size_t)n * NQ, 0.0);
    HostRow h;
    for (auto& g : gens) {
        auto add = [&](int r) -> bool {
            if (!g->row(r, h)) return false;
            for (int k = 0; k < h.ntaps; ++k) {
                const int q = h.start + k;
                if (r < R) { if (q > TB - 1) return false; low[(size_t)r * TB + q] += h.w[k]; }
                else if (r >= n - R) { if (q < n + 2 - TB) return false; high[(size_t)(r - (n - R)) * TB + (q - (n + 2 - TB))] += h.w[k]; }
                else {
                    const int t = q - (r + 1 - R);
                    if (t < 0 || t >= NQ) return false;
                    if (uniform) { if (r == n / 2) wsum[(size_t)t] += h.w[k]; }
                    else tab[(size_t)r * NQ + t] += h.w[k];
                }
            }
            return true;
        };
        if (uniform) {
            // interior rows share one pattern, but rows next to the explicit boundary rows must still fit the window:
            // check a band at each end and the sample row
            for (int r = 0; r < 2 * kMaxBTaps && r < n; ++r) if (!add(r)) return false;
            for (int r = n - 2 * kMaxBTaps > 2 * kMaxBTaps ? n - 2 * kMaxBTaps : 2 * kMaxBTaps; r < n; ++r) if (!add(r)) return false;
            if (n / 2 >= 2 * kMaxBTaps && n / 2 < n - 2 * kMaxBTaps && !add(n / 2)) return false;
        } else {
            for (int r = 0; r < n; ++r) if (!add(r)) return false;
        }
    }
    for (int t = 0; t < NQ; ++t) S.w[t] = (T)wsum[(size_t)t];
    for (int r = 0; r < R; ++r)
        for (int k = 0; k < TB; ++k) { S.bw[0][r][k] = (T)low[(size_t)r * TB + k]; S.bw[1][r][k] = (T)high[(size_t)r * TB + k]; }
    if (!uniform) {
        std::vector<T> tabT(tab.size());
        for (size_t i = 0; i < tab.size(); ++i) tabT[i] = (T)tab[i];
        cudaError_t e;
        S.tab = (const T*)plan_upload(plan, tabT.data(), tabT.size() * sizeof(T), &e);
        if (!S.tab) { cudaGetLastError(); return false; }
    }
    if (!S.padded && plan->bc[0].d.kind == DEO_BC_PERIODIC) {
        S.periodic = 1;
    } else if (!S.padded) {
        const HostBC& H = plan->bc[0];
        if (H.d.kind != DEO_BC_AFFINE || H.d.per_face || H.d.K_l > kLineMaxK || H.d.K_r > kLineMaxK) return false;
        S.K_l = H.d.K_l; S.K_r = H.d.K_r;
        for (int t = 0; t < H.d.K_l; ++t) S.a_l[t] = ((const T*)H.a_l.data())[t];
        for (int t = 0; t < H.d.K_r; ++t) S.a_r[t] = ((const T*)H.a_r.data())[t];
        S.b_l = *(const T*)H.b_l.data();
        S.b_r = *(const T*)H.b_r.data();
    }
    return true;
}

// smallest R in 1..4 for which fill succeeds
template <typename T>
bool fill_line_any(deo_plan* plan, std::vector<std::unique_ptr<RowGenerator>>& gens, LineConfig& C) {
    const int n = (int)plan->dims[0];
    const size_t cursor0 = plan->blob_cursor;
    for (int R = 1; R <= 4; ++R) {
        if (n < 4 * R + 4) return false;
        bool ok = false;
        switch (R) {
            case 1: ok = fill_line<T, 1>(plan, gens, C); break;
            case 2: ok = fill_line<T, 2>(plan, gens, C); break;
            case 3: ok = fill_line<T, 3>(plan, gens, C); break;
            case 4: ok = fill_line<T, 4>(plan, gens, C); break;
        }
        if (ok) { C.R = R; return true; }
        plan->blob_cursor = cursor0;
    }
    return false;
}

template <typename T, int R>
int32_t launch_line_R(const LineConfig& C, const void* u, void* du, cudaStream_t s) {
    const LineParams<T, R>& S = *reinterpret_cast<const LineParams<T, R>*>(C.params.data());
    constexpr int VEC = LVec<T>::N;
    const long long threads = ((long long)S.n + VEC - 1) / VEC;
    constexpr int block = 256;
    static const int vpt = [] {                           // vectors per thread (DEO_LINE_VPT: 1, 2 or 4)
        const char* e = getenv("DEO_LINE_VPT");
        const int v = e ? atoi(e) : 1;                    // measured on B200, N = 1e6: 1 -> 3.01 us, 2 -> 3.45 us, 4 -> 3.26 us per application
        return (v == 1 || v == 2 || v == 4) ? v : 1;
    }();
    const unsigned grid = (unsigned)((threads + (long long)block * vpt - 1) / ((long long)block * vpt));
    static const bool pdl = getenv("DEO_LINE_PDL") && atoi(getenv("DEO_LINE_PDL")) == 1;   // measured: no gain (3.03 vs 3.01 us), off by default
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(block); cfg.dynamicSmemBytes = 0; cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = pdl ? 1 : 0;
    const T* up = (const T*)u;
    T* dup = (T*)du;
#define DEO_LINE_LAUNCH(TAB, V) DEO_CUDA(cudaLaunchKernelEx(&cfg, k_line<T, R, TAB, V>, S, up, dup))
    if (S.table) { if (vpt == 1) DEO_LINE_LAUNCH(true, 1); else if (vpt == 2) DEO_LINE_LAUNCH(true, 2); else DEO_LINE_LAUNCH(true, 4); }
    else { if (vpt == 1) DEO_LINE_LAUNCH(false, 1); else if (vpt == 2) DEO_LINE_LAUNCH(false, 2); else DEO_LINE_LAUNCH(false, 4); }
#undef DEO_LINE_LAUNCH
    return DEO_OK;
}

template <typename T>
int32_t launch_line_T(const LineConfig& C, const void* u, void* du, cudaStream_t s) {
    switch (C.R) {
        case 1: return launch_line_R<T, 1>(C, u, du, s);
        case 2: return launch_line_R<T, 2>(C, u, du, s);
        case 3: return launch_line_R<T, 3>(C, u, du, s);
        case 4: return launch_line_R<T, 4>(C, u, du, s);
    }
    set_error("line kernel: unsupported radius %d", C.R);
    return DEO_ERR_UNSUPPORTED;
}

}  // namespace

// Attaches a LineConfig when the plan is a 1-D sum of stencils of reach <= 4 with an affine BC (or a pre-padded input).
int32_t line_configure(deo_plan* plan) {
    if (plan->ndims != 1 || getenv("DEO_NO_LINE")) return DEO_OK;
    if (!plan->padded[0] && plan->bc[0].d.kind != DEO_BC_PERIODIC && (plan->bc[0].d.kind != DEO_BC_AFFINE || plan->bc[0].d.per_face)) return DEO_OK;
    if (plan->dims[0] > (1LL << 30)) return DEO_OK;
    std::vector<std::unique_ptr<RowGenerator>> gens;
    for (size_t k = 0; k < plan->ops.size(); ++k) {
        gens.push_back(make_row_generator(plan, (int)k));
        if (!gens.back()->ok()) return DEO_OK;
    }
    auto cfg = std::make_shared<LineConfig>();
    const bool ok = plan->dtype == DEO_F64 ? fill_line_any<double>(plan, gens, *cfg) : fill_line_any<float>(plan, gens, *cfg);
    if (!ok) return DEO_OK;
    plan->line = cfg;
    plan->kernel = reinterpret_cast<const int*>(cfg->params.data())[3] ? "line-table" : "line";
    return DEO_OK;
}

int32_t launch_line(const deo_plan* plan, void* du, const void* u, cudaStream_t s) {
    const LineConfig& C = *static_cast<const LineConfig*>(plan->line.get());
    DEO_REQUIRE((reinterpret_cast<uintptr_t>(u) & 15) == 0 && (reinterpret_cast<uintptr_t>(du) & 15) == 0,
                "line kernel: buffers must be 16-byte aligned");
    return plan->dtype == DEO_F64 ? launch_line_T<double>(C, u, du, s) : launch_line_T<float>(C, u, du, s);
}

}  // namespace deo
