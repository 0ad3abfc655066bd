// Host side of the tiled streaming kernel (kernel_star.cuh): eligibility, parameter packing, tensor map, dispatch.
#include <cstdlib>

#include "kernel_star2.cuh"

namespace deo {

// ===================================== host side ==========================================================
namespace {

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

PFN_encodeTiled get_encode() {
    static PFN_encodeTiled fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = (PFN_encodeTiled)p;
    }
    return fn;
}


template <typename T>
int32_t dispatch2_T(const StarConfig& C, const void* u, void* du, long long z0, long long z1, cudaStream_t s) {
    switch (C.R) {
        case 1: return star2_launch_R<T, 1>(C, u, du, z0, z1, s);
        case 2: return star2_launch_R<T, 2>(C, u, du, z0, z1, s);
        case 3: return star2_launch_R<T, 3>(C, u, du, z0, z1, s);
        case 4: return star2_launch_R<T, 4>(C, u, du, z0, z1, s);
    }
    set_error("star kernel: unsupported radius %d", C.R);
    return DEO_ERR_UNSUPPORTED;
}

template <typename T>
int32_t dispatch_T(const StarConfig& C, const void* u, void* du, long long z0, long long z1, cudaStream_t s) {
    if (C.v2) return dispatch2_T<T>(C, u, du, z0, z1, s);
    switch (C.R) {
        case 1: return star_launch_R<T, 1>(C, u, du, z0, z1, s);
        case 2: return star_launch_R<T, 2>(C, u, du, z0, z1, s);
        case 3: return star_launch_R<T, 3>(C, u, du, z0, z1, s);
        case 4: return star_launch_R<T, 4>(C, u, du, z0, z1, s);
    }
    set_error("star kernel: unsupported radius %d", C.R);
    return DEO_ERR_UNSUPPORTED;
}

// Boundary data of kernel axis `ka` (plan axis `pa`): one affine BC for the whole face (constant bank), and -- persistent
// kernel only -- per-pencil BC tables (multi_dim_bc_operators.jl:54-57) or a pre-padded input whose ghost layer is read
// instead of computed (derivative_operator_functions.jl:27-69).
template <typename T, int R>
bool fill_bc(const deo_plan* plan, int pa, int ka, bool v2, StarParams<T, R>& S) {
    constexpr int NQ = StarParams<T, R>::NQ;
    const DevPlan<T>& P = *reinterpret_cast<const DevPlan<T>*>(plan->devplan.data());
    S.padded[ka] = plan->padded[pa];
    S.per_face[ka] = 0;
    if (plan->padded[pa]) return v2;
    const HostBC& H = plan->bc[pa];
    if (H.d.kind != DEO_BC_AFFINE) return false;
    const int Kmax = ka == 2 ? NQ : kStarMaxK;
    if (H.d.K_l > Kmax || H.d.K_r > Kmax || H.d.K_l > kStarMaxK || H.d.K_r > kStarMaxK) return false;
    S.K_l[ka] = H.d.K_l;
    S.K_r[ka] = H.d.K_r;
    if (H.d.per_face) {
        if (!v2) return false;
        S.per_face[ka] = 1;
        S.pf_a_l[ka] = P.bc[pa].a_l; S.pf_b_l[ka] = P.bc[pa].b_l;
        S.pf_a_r[ka] = P.bc[pa].a_r; S.pf_b_r[ka] = P.bc[pa].b_r;
        return true;
    }
    const T* al = (const T*)H.a_l.data();
    const T* ar = (const T*)H.a_r.data();
    for (int t = 0; t < H.d.K_l; ++t) S.a_l[ka][t] = al[t];
    for (int t = 0; t < H.d.K_r; ++t) S.a_r[ka][t] = ar[t];
    S.b_l[ka] = *(const T*)H.b_l.data();
    S.b_r[ka] = *(const T*)H.b_r.data();
    if (ka == 2) {
        for (int t = 0; t < H.d.K_l; ++t) S.azl_pad[t] = al[t];
        for (int t = 0; t < H.d.K_r; ++t) S.azr_pad[NQ - H.d.K_r + t] = ar[t];
    }
    return true;
}

template <typename T, int R>
bool fill_params(const deo_plan* plan, const int kaxis_of_plan_axis[3], bool mid, StarConfig& C) {
    using SP = StarParams<T, R>;
    constexpr int NQ = SP::NQ, TB = SP::TB;
    const DevPlan<T>& P = *reinterpret_cast<const DevPlan<T>*>(plan->devplan.data());
    C.params.assign(sizeof(SP), 0);
    SP& S = *reinterpret_cast<SP*>(C.params.data());
    const int nd = plan->ndims;
    const int march_plan_axis = nd - 1;
    S.nx = P.n_out[0];
    S.ny = mid ? P.n_out[1] : 1;
    S.nz = P.n_out[march_plan_axis];
    S.in_off_z = P.in_off[march_plan_axis];
    S.in_off_x = P.in_off[0];
    S.in_off_y = mid ? P.in_off[1] : 0;
    S.row0_z = P.row0[march_plan_axis];
    S.nglob_z = P.n_glob[march_plan_axis];
    S.isy = mid ? P.in_stride[1] : 0;
    S.osy = mid ? P.out_stride[1] : 0;
    S.isz = P.in_stride[march_plan_axis];
    S.osz = P.out_stride[march_plan_axis];
    for (int a = 0; a < 3; ++a) { S.has[a] = 0; S.opidx[a] = -1; S.nedge[a] = 0; S.K_l[a] = S.K_r[a] = 0; S.padded[a] = 0; S.per_face[a] = 0; }
    // the explicit rows live in device memory: fetch them back through the host copies kept in the plan
    for (int k = 0; k < P.nops; ++k) {
        const DevOp<T>& op = P.ops[k];
        const int ka = kaxis_of_plan_axis[op.axis];
        if (ka < 0 || S.has[ka]) return false;
        S.has[ka] = 1;
        C.mask |= 1 << ka;
        S.opidx[ka] = k;
        const int r = -op.soff[0];
        if (op.mode != MODE_CONST || op.ntaps != 2 * r + 1 || r < 1 || r > R) return false;
        for (int t = 0; t < op.ntaps; ++t) S.w[ka][R - r + t] = op.w[0][t];
        if (op.nlow >= r + 1 || op.nhigh >= r + 1) return false;
        const int n = op.n;
        if (n < 4 * R + 4) return false;
        S.nedge[ka] = r;                                   // rows 0..r-1 and n-r..n-1 read a ghost
        std::vector<BRow<T>> br((size_t)op.nlow + op.nhigh);
        if (!br.empty()) {                                    // the host copy kept by the plan: no device read-back
            if ((size_t)k >= plan->host_brows.size() || plan->host_brows[(size_t)k].size() != br.size() * sizeof(BRow<T>)) return false;
            memcpy(br.data(), plan->host_brows[(size_t)k].data(), br.size() * sizeof(BRow<T>));
        }
        for (int i = 0; i < r; ++i) {                      // low face, global row i; tap k <-> q[k]
            if (i < op.nlow) {
                const BRow<T>& b = br[i];
                if (b.start != 0 || b.ntaps > TB || b.acc64) return false;
                for (int t = 0; t < b.ntaps; ++t) S.bw[ka][0][i][t] = b.w[t];
            } else {                                       // interior stencil whose window q[i+1-r .. i+1+r] starts at or after q[0]
                for (int t = 0; t < op.ntaps; ++t) S.bw[ka][0][i][i + 1 - r + t] = op.w[0][t];
            }
        }
        for (int i = 0; i < r; ++i) {                      // high face, global row n-r+i; tap k <-> q[n+2-TB+k]
            const int row = n - r + i;
            if (row >= n - op.nhigh) {
                const BRow<T>& b = br[(size_t)op.nlow + (row - (n - op.nhigh))];
                if (b.start + b.ntaps != n + 2 || b.ntaps > TB || b.acc64) return false;
                for (int t = 0; t < b.ntaps; ++t) S.bw[ka][1][i][TB - b.ntaps + t] = b.w[t];
            } else {
                const int k0 = (row + 1 - r) - (n + 2 - TB);
                if (k0 < 0) return false;
                for (int t = 0; t < op.ntaps; ++t) S.bw[ka][1][i][k0 + t] = op.w[0][t];
            }
        }
        // boundary condition of this axis
        if (!fill_bc<T, R>(plan, op.axis, ka, C.v2, S)) return false;
    }
    return true;
}


// ---- merged per-row tables (TABLE variants) ------------------------------------------------------------------
// Every operator is linear in the same ghost-padded pencil, so all operators acting along one axis fold into ONE
// per-row stencil: W_a[r][t] = sum_ops (c*w)_op,r placed on the common window q[r+1-R .. r+1+R].  That covers
// non-uniform grids (per-row Fornberg weights), coefficient vectors, upwind operators (the wind direction of row r
// is known from sign(c[r]) when the plan is built or its coefficients are updated) and several operators per axis
// in one pass.  The sum over operators is formed here in Float64 and rounded once to T, so the result differs from
// the reference's operator-by-operator sum by reassociation only (within the parity tolerance, not bitwise).
struct AxisRows {
    int n = 0;
    std::vector<std::vector<HostRow>> ops;    // per operator on this axis: n rows
};

// smallest template radius R such that rows [R, n-R) fit q[r+1-R .. r+1+R] and the R rows at each face fit the
// TB = 2R+2 taps nearest to it
int table_radius(const AxisRows& A) {
    for (int R = 1; R <= 4; ++R) {
        const int TB = 2 * R + 2, n = A.n;
        if (n < 4 * R + 4) return 0;
        bool ok = true;
        for (const auto& rows : A.ops) {
            for (int r = 0; r < n && ok; ++r) {
                const HostRow& h = rows[(size_t)r];
                const int lo = h.start, hi = h.start + h.ntaps - 1;   // q indices
                if (r < R) ok = hi <= TB - 1;
                else if (r >= n - R) ok = lo >= n + 2 - TB;
                else ok = lo >= r + 1 - R && hi <= r + 1 + R;
            }
            if (!ok) break;
        }
        if (ok) return R;
    }
    return 0;
}

template <typename T, int R>
bool fill_params_table(deo_plan* plan, const AxisRows (&axes)[3], const int plan_axis_of_kaxis[3], bool mid, StarConfig& C) {
    using SP = StarParams<T, R>;
    constexpr int NQ = SP::NQ, TB = SP::TB;
    const DevPlan<T>& P = *reinterpret_cast<const DevPlan<T>*>(plan->devplan.data());
    C.params.assign(sizeof(SP), 0);
    SP& S = *reinterpret_cast<SP*>(C.params.data());
    const int march_plan_axis = plan->ndims - 1;
    S.nx = P.n_out[0];
    S.ny = mid ? P.n_out[1] : 1;
    S.nz = P.n_out[march_plan_axis];
    S.in_off_z = P.in_off[march_plan_axis];
    S.in_off_x = P.in_off[0];
    S.in_off_y = mid ? P.in_off[1] : 0;
    S.row0_z = P.row0[march_plan_axis];
    S.nglob_z = P.n_glob[march_plan_axis];
    S.isy = mid ? P.in_stride[1] : 0;
    S.osy = mid ? P.out_stride[1] : 0;
    S.isz = P.in_stride[march_plan_axis];
    S.osz = P.out_stride[march_plan_axis];
    for (int ka = 0; ka < 3; ++ka) {
        S.has[ka] = 0; S.opidx[ka] = -1; S.nedge[ka] = 0; S.K_l[ka] = S.K_r[ka] = 0; S.tab[ka] = nullptr; S.padded[ka] = 0; S.per_face[ka] = 0;
        const AxisRows& A = axes[ka];
        if (A.ops.empty()) continue;
        S.has[ka] = 1;
        C.mask |= 1 << ka;
        S.nedge[ka] = R;
        const int n = A.n;
        std::vector<double> tab((size_t)n * NQ, 0.0), lowrows((size_t)R * TB, 0.0), highrows((size_t)R * TB, 0.0);
        for (const auto& rows : A.ops) {
            for (int r = 0; r < n; ++r) {
                const HostRow& h = rows[(size_t)r];
                for (int k = 0; k < h.ntaps; ++k) {
                    const int q = h.start + k;
                    if (r < R) lowrows[(size_t)r * TB + q] += h.w[k];
                    else if (r >= n - R) highrows[(size_t)(r - (n - R)) * TB + (q - (n + 2 - TB))] += h.w[k];
                    else tab[(size_t)r * NQ + (q - (r + 1 - R))] += h.w[k];
                }
            }
        }
        for (int r = 0; r < R; ++r)
            for (int k = 0; k < TB; ++k) {
                S.bw[ka][0][r][k] = (T)lowrows[(size_t)r * TB + k];
                S.bw[ka][1][r][k] = (T)highrows[(size_t)r * TB + k];
            }
        std::vector<T> tabT(tab.size());
        for (size_t i = 0; i < tab.size(); ++i) tabT[i] = (T)tab[i];
        cudaError_t e;
        S.tab[ka] = (const T*)plan_upload(plan, tabT.data(), tabT.size() * sizeof(T), &e);
        if (!S.tab[ka]) { cudaGetLastError(); return false; }
        // boundary condition of this axis
        if (!fill_bc<T, R>(plan, plan_axis_of_kaxis[ka], ka, C.v2, S)) return false;
    }
    return true;
}

template <typename T>
bool fill_params_table_R(deo_plan* plan, const AxisRows (&axes)[3], const int pax[3], bool mid, StarConfig& C) {
    switch (C.R) {
        case 1: return fill_params_table<T, 1>(plan, axes, pax, mid, C);
        case 2: return fill_params_table<T, 2>(plan, axes, pax, mid, C);
        case 3: return fill_params_table<T, 3>(plan, axes, pax, mid, C);
        case 4: return fill_params_table<T, 4>(plan, axes, pax, mid, C);
    }
    return false;
}

template <typename T>
bool fill_params_R(const deo_plan* plan, const int kaxis[3], bool mid, StarConfig& C) {
    switch (C.R) {
        case 1: return fill_params<T, 1>(plan, kaxis, mid, C);
        case 2: return fill_params<T, 2>(plan, kaxis, mid, C);
        case 3: return fill_params<T, 3>(plan, kaxis, mid, C);
        case 4: return fill_params<T, 4>(plan, kaxis, mid, C);
    }
    return false;
}

}  // namespace

Star2Runtime& star2_rt() {
    // one set of words per device (deo_init may select another device later in the same process)
    static Star2Runtime per_device[64];
    int dev = 0;
    cudaGetDevice(&dev);
    Star2Runtime& r = per_device[dev >= 0 && dev < 64 ? dev : 0];
    if (!r.err_host) {
        if (cudaHostAlloc((void**)&r.err_host, sizeof(int), cudaHostAllocMapped) == cudaSuccess &&
            cudaHostGetDevicePointer((void**)&r.err_dev, r.err_host, 0) == cudaSuccess) {
            *r.err_host = 0;
            if (cudaMalloc((void**)&r.sched, 2 * sizeof(unsigned int)) != cudaSuccess || cudaMemset(r.sched, 0, 2 * sizeof(unsigned int)) != cudaSuccess) {
                cudaGetLastError();
                r.sched = nullptr;
            }
        } else {
            cudaGetLastError();
            r.err_host = nullptr; r.err_dev = nullptr;
        }
    }
    return r;
}

// Set (and cleared here) when a slab launch gave up waiting for its neighbours' halo planes.
bool star_take_halo_timeout() {
    Star2Runtime& r = star2_rt();
    if (r.err_host && *(volatile int*)r.err_host) { *(volatile int*)r.err_host = 0; return true; }
    return false;
}

// Tile origins (persistent kernel): (tile_x * TX - xshift, tile_y * TY - yshift), see tiling_host.hpp.  The first-generation
// kernel lays its tiles out from the origin only.  Returns false when no shift fits (the plan then runs on the per-point kernel).
static bool choose_shifts(const deo_plan* plan, StarConfig& cfg, bool mid) {
    const size_t es = plan->elem();
    const int R = cfg.R, VEC = (int)(16 / es), HXh = ((R + VEC - 1) / VEC) * VEC;
    const int nwy = cfg.v2 ? 16 : cfg.nwy, py = cfg.v2 ? 2 : cfg.py;
    const long long TX = mid ? 32 * VEC : 32 * VEC * nwy * py, TY = nwy * py;
    const bool may_shift = cfg.v2;
    cfg.xshift = cfg.yshift = 0;
    if ((cfg.mask & 1) != 0 || cfg.xres != 0) {                          // an axis without operator has no face logic: any tiling
        const long long sh = tiling::pick_shift(plan->dims[0], TX, HXh, R, plan->bc[0].d.K_l, plan->bc[0].d.K_r, cfg.xres, VEC, may_shift);
        if (sh < 0) return false;
        cfg.xshift = (int)sh;
    }
    if (mid && (cfg.mask & 2) != 0) {
        const long long sh = tiling::pick_shift(plan->dims[1], TY, R, R, plan->bc[1].d.K_l, plan->bc[1].d.K_r, 0, 1, may_shift);
        if (sh < 0) return false;
        cfg.yshift = (int)sh;
    }
    if (cfg.xshift % VEC != 0) cfg.scalar_io = true;                     // vectors of du straddle 16-byte boundaries
    return true;
}

// Decides whether the plan can run on the tiled kernel; if so attaches a StarConfig to it.
//   CONST variants: one uniform centered constant-coefficient operator per axis, in axis order -- weights in the constant
//                   bank, arithmetic order identical to the per-point kernel (bitwise equal results);
//   TABLE variants: anything else that is a sum of 1-D stencils of reach <= 4 with affine BCs (non-uniform grids,
//                   coefficient vectors, upwind, several operators per axis) -- merged per-row weight tables.
int32_t star_configure(deo_plan* plan) {
    const int nd = plan->ndims;
    if (nd < 2 || nd > 3) return DEO_OK;
    const bool want_v2 = !(getenv("DEO_STAR_V") && atoi(getenv("DEO_STAR_V")) == 1);
    if (plan->accumulate && !want_v2) return DEO_OK;                    // overwrite = false: persistent kernel only
    const size_t es = plan->elem();
    // The first-generation kernel needs a tensor map (row pitch a multiple of 16 B) and takes no pre-padded input.  The
    // persistent kernel falls back to cp.async element copies where no tensor map exists (odd row pitch).  An input padded
    // along the contiguous axis has its rows ONE element behind the rows of du, and a TMA box must start on a 16-byte
    // boundary (an odd Float64 start index is an illegal instruction): the tile origins are shifted by one element instead
    // (tiles start at x = -1), which puts every box of u on a boundary and leaves the vectors of du misaligned -- du is
    // then read / written element-wise.  Slab plans take every path but the pre-padded ones (odd row lengths: loader, three-launch
    // schedule).
    const bool pitch_ok = ((size_t)plan->in_dim(0) * es) % 16 == 0;
    const bool need_loader = !pitch_ok || (plan->padded[0] && getenv("DEO_STAR2_NO_XSHIFT") != nullptr);
    const int xres = (plan->padded[0] && !need_loader) ? 1 : 0;
    bool any_padded = false;
    for (int a = 0; a < nd; ++a) any_padded = any_padded || plan->padded[a];
    if (need_loader && !want_v2) return DEO_OK;
    if (any_padded && (!want_v2 || plan->slab_axis >= 0)) return DEO_OK;   // a slab's input carries its halo planes, never a ghost layer
    const bool mid = nd == 3;
    const int kaxis[3] = {0, mid ? 1 : 2, mid ? 2 : -1};                 // plan axis -> kernel axis (x, mid, march)
    const int paxis[3] = {0, mid ? 1 : -1, mid ? 2 : 1};                 // kernel axis -> plan axis
    auto cfg = std::make_shared<StarConfig>();
    cfg->mid = mid;
    cfg->accumulate = plan->accumulate != 0;
    cfg->loader = need_loader;
    cfg->xres = xres;
    cfg->scalar_io = ((size_t)plan->dims[0] * es) % 16 != 0;              // rows of du not 16-byte aligned (choose_shifts adds: vectors straddling)
    cfg->in_dims[0] = (int)plan->in_dim(0);
    cfg->in_dims[1] = mid ? (int)plan->in_dim(1) : 1;
    cfg->in_dims[2] = (int)plan->in_dim(nd - 1);
    cfg->sm_count = rt().sm_count;
    const char* env_py = getenv("DEO_STAR_PY");
    cfg->py = env_py ? atoi(env_py) : 2;                    // 2 rows per thread, two CTAs per SM measured fastest on B200
    if (cfg->py != 2 && cfg->py != 4) cfg->py = 2;
    const char* env_nwy = getenv("DEO_STAR_NWY");
    cfg->nwy = (env_nwy && atoi(env_nwy) == 16 && cfg->py == 2) ? 16 : 8;
    cfg->zchunk_pref = 64;
    cfg->v2 = !(getenv("DEO_STAR_V") && atoi(getenv("DEO_STAR_V")) == 1);            // A/B: DEO_STAR_V=1 selects the first-generation kernel
    // march-axis chunk bound: short items re-synchronise the CTAs often (L2 reuse of the shared halo rows) at the price of
    // 2R priming planes each; measured best: 24 planes (3-D) / 16 rows (2-D strips) for the persistent kernel, 32 for the first one
    const int zchunk_default = cfg->v2 ? (mid ? 24 : 16) : 32;
    cfg->zchunk_max = getenv("DEO_STAR_ZCHUNK") ? atoi(getenv("DEO_STAR_ZCHUNK")) : zchunk_default;
    if (cfg->zchunk_max < 4 || cfg->zchunk_max > (1 << 20)) cfg->zchunk_max = zchunk_default;   // tuning knobs are clamped, never trusted
    cfg->l2promo = getenv("DEO_TMA_L2PROMO") ? atoi(getenv("DEO_TMA_L2PROMO")) : 3;
    if (cfg->l2promo < 0 || cfg->l2promo > 3) cfg->l2promo = 3;
    cfg->group = getenv("DEO_STAR_GROUP") ? atoi(getenv("DEO_STAR_GROUP")) : -1;   // measured: the plain order is fastest
    if (getenv("DEO_HALO_TIMEOUT_S") && atof(getenv("DEO_HALO_TIMEOUT_S")) > 0)
        cfg->halo_timeout_ns = (unsigned long long)(atof(getenv("DEO_HALO_TIMEOUT_S")) * 1e9);

    // ---- CONST eligibility ----
    bool const_ok = plan->ops.size() <= 3 && !getenv("DEO_STAR_FORCE_TABLE");
    int R = 0, prev_axis = -1;
    for (const HostOp& h : plan->ops) {
        if (h.d.kind != DEO_OP_CENTERED || h.d.nonuniform) const_ok = false;
        if (h.d.axis <= prev_axis) const_ok = false;                     // one operator per axis, in axis order (sum association)
        prev_axis = h.d.axis;
        R = R > h.d.stencil_length / 2 ? R : h.d.stencil_length / 2;
    }
    if (R < 1 || R > 4) const_ok = false;
    if (getenv("DEO_STAR_DEBUG")) fprintf(stderr, "[star_configure] const_ok=%d nops=%zu R=%d\n", (int)const_ok, plan->ops.size(), R);
    if (const_ok) {
        if (plan->slab_axis >= 0 && plan->slab_count < 3 * R + 3) return DEO_OK;
        cfg->R = R;
        cfg->mask = 0;
        const bool ok = plan->dtype == DEO_F64 ? fill_params_R<double>(plan, kaxis, mid, *cfg) : fill_params_R<float>(plan, kaxis, mid, *cfg);
        const bool fits = ok && choose_shifts(plan, *cfg, mid);
        if (getenv("DEO_STAR_DEBUG"))
            fprintf(stderr, "[star_configure] CONST: filled=%d fits=%d R=%d mask=%d shifts=(%d, %d) dims=%lld x %lld\n", (int)ok, (int)fits, R, cfg->mask, cfg->xshift,
                    cfg->yshift, (long long)plan->dims[0], (long long)plan->dims[1]);
        if (fits) {
            for (const HostOp& h : plan->ops) if (h.d.axis == nd - 1) cfg->nedge_march = h.d.stencil_length / 2;
            plan->star = cfg;
            plan->kernel = "star";
            return DEO_OK;
        }
    }
    // ---- TABLE eligibility ----
    if (getenv("DEO_STAR_NO_TABLE")) return DEO_OK;
    AxisRows axes[3];
    for (size_t k = 0; k < plan->ops.size(); ++k) {
        const HostOp& h = plan->ops[k];
        const int ka = kaxis[h.d.axis];
        if (ka < 0 || h.d.len > (1 << 20)) return DEO_OK;
        axes[ka].n = h.d.len;
        axes[ka].ops.emplace_back();
        auto gen = make_row_generator(plan, (int)k);
        if (!gen->ok()) return DEO_OK;
        std::vector<HostRow>& rows = axes[ka].ops.back();
        rows.resize((size_t)h.d.len);
        for (int r = 0; r < h.d.len; ++r) if (!gen->row(r, rows[(size_t)r])) return DEO_OK;
    }
    R = 0;
    for (int ka = 0; ka < 3; ++ka) {
        if (axes[ka].ops.empty()) continue;
        const int Ra = table_radius(axes[ka]);
        if (Ra == 0) return DEO_OK;
        R = R > Ra ? R : Ra;
    }
    if (R < 1) return DEO_OK;
    for (int ka = 0; ka < 3; ++ka) if (!axes[ka].ops.empty() && axes[ka].n < 4 * R + 4) return DEO_OK;
    if (plan->slab_axis >= 0 && plan->slab_count < 3 * R + 3) return DEO_OK;
    cfg->R = R;
    cfg->mask = 0;
    cfg->table = true;
    cfg->py = 2;
    cfg->nwy = 8;
    const size_t cursor0 = plan->blob_cursor;
    const bool ok = plan->dtype == DEO_F64 ? fill_params_table_R<double>(plan, axes, paxis, mid, *cfg)
                                           : fill_params_table_R<float>(plan, axes, paxis, mid, *cfg);
    if (!ok || !choose_shifts(plan, *cfg, mid)) { plan->blob_cursor = cursor0; return DEO_OK; }
    cfg->nedge_march = (cfg->mask & 4) ? R : 0;
    plan->star = cfg;
    plan->kernel = "star-table";
    return DEO_OK;
}

int32_t launch_star(const deo_plan* plan, void* du, const void* u, long long z0, long long z1, cudaStream_t s, bool explicit_range) {
    StarConfig& C = *static_cast<StarConfig*>(plan->star.get());
    PFN_encodeTiled enc = get_encode();
    if (!enc) { set_error("cuTensorMapEncodeTiled is not available from this driver"); return DEO_ERR_CUDA; }
    DEO_REQUIRE((C.loader || (reinterpret_cast<uintptr_t>(u) & 15) == 0) && (C.scalar_io || (reinterpret_cast<uintptr_t>(du) & 15) == 0),
                "star kernel: buffers must be 16-byte aligned");
    DEO_REQUIRE((reinterpret_cast<uintptr_t>(u) % plan->elem()) == 0 && (reinterpret_cast<uintptr_t>(du) % plan->elem()) == 0,
                "star kernel: buffers must be aligned to their element type");
    // tensor map over the input field: (nx, ny, nz_in) for 3-D, (nx, 1, ny_in) for 2-D
    const size_t es = plan->elem();
    const bool mid = C.mid;
    const cuuint64_t nx = (cuuint64_t)plan->in_dim(0);
    const cuuint64_t d1 = mid ? (cuuint64_t)plan->in_dim(1) : 1;
    const cuuint64_t d2 = (cuuint64_t)plan->in_dim(plan->ndims - 1);
    cuuint64_t dims[3] = {nx, d1, d2};
    cuuint64_t strides[2] = {nx * es, nx * d1 * es};
    cuuint32_t estr[3] = {1, 1, 1};
    const int VEC = (int)(16 / es);
    const int HX = ((C.R + VEC - 1) / VEC) * VEC;
    cuuint32_t box[3];
    const int tile_rows = C.v2 ? 32 : C.nwy * C.py;
    if (mid) { box[0] = (cuuint32_t)(32 * VEC + 2 * HX); box[1] = (cuuint32_t)(tile_rows + 2 * C.R); box[2] = 1; }
    else { box[0] = 256; box[1] = 1; box[2] = 1; }
    const int promo_env = C.l2promo;
    const CUtensorMapL2promotion promo = promo_env == 0 ? CU_TENSOR_MAP_L2_PROMOTION_NONE : promo_env == 1 ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B
                                         : promo_env == 2 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B : CU_TENSOR_MAP_L2_PROMOTION_L2_256B;
    CUresult r = C.loader ? CUDA_SUCCESS : enc(&C.tmap, es == 8 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT64 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<void*>(u), dims,
                     strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, promo,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed with CUresult %d", (int)r); return DEO_ERR_CUDA; }
    if (!mid && !explicit_range) { z0 = 0; z1 = plan->local_dim(plan->ndims - 1); }   // 2-D arrays stream along their last axis
    // The one-sided rows of the march axis take their term from the register queue at the steps whose centre is
    // global plane R (low face) / n-1-R (high face); a launch range that contains such rows but not that step
    // (never produced by this library's own callers) runs on the per-point kernel instead.
    {
        const int ax = plan->ndims - 1;
        const long long row0 = ax == plan->slab_axis ? plan->slab_start : 0;
        const long long n = plan->dims[ax];
        const long long g0 = z0 + row0, g1 = z1 + row0;
        const int nlow = C.nedge_march, nhigh = C.nedge_march;
        const bool low_ok = !(nlow > 0 && g0 < nlow) || (g0 == 0 && g1 > C.R);
        const bool high_ok = !(nhigh > 0 && g1 > n - nhigh) || (g1 == n && g0 <= n - 1 - C.R);
        if (!low_ok || !high_ok) {
            if (!mid) { set_error("star kernel: 2-D row range [%lld, %lld) cuts a face's one-sided rows", z0, z1); return DEO_ERR_UNSUPPORTED; }
            return launch_generic(plan, du, u, z0, z1, s);
        }
    }
    return plan->dtype == DEO_F64 ? dispatch_T<double>(C, u, du, z0, z1, s) : dispatch_T<float>(C, u, du, z0, z1, s);
}

// Slab plans: ONE launch over all own planes; the CTAs of the first / last march-axis chunk (scheduled last) wait in the
// kernel until *halo_flag >= expect, which the communication stream publishes after the neighbour exchange has landed.
// Returns DEO_ERR_UNSUPPORTED when the slab is too thin to chunk (the caller then uses the three-launch schedule).
StarLimits star_limits(const deo_plan* plan) {
    StarLimits L;
    if (!plan->star) return L;
    const StarConfig& C = *static_cast<const StarConfig*>(plan->star.get());
    L.fusable = C.mid && C.zchunk_max > 0 && !C.loader;
    L.min_fused_planes = 3LL * C.zchunk_max + 4 * C.R + 4;                 // at least 3 chunks whatever the chunk search picks (it only shortens chunks)
    return L;
}

// du = u + dt * (A u): the explicit-stepper update fused into the store of the persistent kernel.
bool star_can_axpy(const deo_plan* plan) {
    if (!plan->star) return false;
    const StarConfig& C = *static_cast<const StarConfig*>(plan->star.get());
    return C.v2 && !C.accumulate;
}
void star_set_axpy(const deo_plan* plan, bool on, double dt) {
    StarConfig& C = *static_cast<StarConfig*>(plan->star.get());
    C.axpy = on; C.dt = dt;
}

int32_t launch_star_fused(const deo_plan* plan, void* du, const void* u, long long cnt, cudaStream_t s, const int* halo_flag, int expect, int sides) {
    StarConfig& C = *static_cast<StarConfig*>(plan->star.get());
    const StarLimits lim = star_limits(plan);
    if (!lim.fusable || cnt < lim.min_fused_planes) return DEO_ERR_UNSUPPORTED;
    C.halo_flag = halo_flag; C.halo_expect = expect; C.halo_sides = sides;
    const int32_t rc = launch_star(plan, du, u, 0, cnt, s, false);
    C.halo_flag = nullptr;
    return rc;
}

}  // namespace deo
