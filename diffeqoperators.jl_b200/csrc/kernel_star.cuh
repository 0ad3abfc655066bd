// Tiled streaming kernel for star-shaped composites on uniform grids: Dxx [+ Dyy] [+ Dzz] (any centered
// derivative / approximation order up to radius 4, constant coefficient) fused with affine boundary
// conditions, ONE pass over u:
//
//   * x-y tiles with halo are staged into shared memory by TMA (cp.async.bulk.tensor over one 3-D tensor
//     map of u; out-of-bounds elements are zero-filled by the hardware) through a ring of planes guarded
//     by full/empty mbarriers; one elected thread issues the copies two planes ahead, all NWY warps compute;
//   * 2.5-D register streaming along the slowest axis: every thread keeps the 2R+1 planes of its own
//     columns in a register queue, so the centre plane's own values and all march-axis taps cost no load;
//   * every thread owns PY rows x VEC contiguous points (VEC = 16 B / sizeof(T)): all shared-memory loads
//     and global stores are 128-bit and conflict-free, the along-y taps reuse every loaded row PY times;
//   * affine BC ghost values (ghost = b + a . u[edge], bc_operators.jl:188-191) are computed in the kernel
//     and patched into the register windows -- no BoundaryPaddedArray is materialised;
//   * the one-sided boundary rows (convolve_BC_left!/right!, convolutions.jl:76-118) are predicated edge
//     paths: along x/y from the tile, along the streaming axis from the register queue (their term is
//     added to du when the queue holds the planes they need);
//   * 2-D arrays run the same code with no middle axis (the tile is a long x strip, streaming along y).
//
// Arithmetic: acc = fma(w[t], q[t], acc) over the taps in the reference's order (idx = 1..sl), operators
// summed in A.ops order, w = (c*w) pre-multiplied as the reference forms it (convolutions.jl:47).
#pragma once
#include <cuda.h>

#include <cstdlib>
#include <type_traits>

#include "generic_device.cuh"

namespace deo {

constexpr int kStarMaxK = 8;   // BC stencil length (a_l / a_r entries) along x / y

template <typename T, int R>
struct StarParams {
    static constexpr int NQ = 2 * R + 1, TB = 2 * R + 2;
    int nx, ny, nz;              // local output extents along the kernel axes (x, mid, march); ny == 1 without a mid axis
    int has[3];                  // an operator acts along kernel axis a
    int opidx[3];                // its index in the plan
    int in_off_z, row0_z, nglob_z, in_off_x;
    long long isy, isz, osy, osz;
    int nedge[3];                // rows per face whose stencil touches the ghost (= operator radius)
    int in_off_y;                // input index = output index + in_off (1 where the input carries a ghost layer; slab halo along the march axis)
    int padded[3];               // persistent kernel: the ghosts of this axis come from the input array (pre-padded input), not from a BC
    int per_face[3];             // persistent kernel: one affine BC per boundary pencil (MultiDimDirectionalBC.BCs), tables below
    int pad2_;
    const T *pf_a_l[3], *pf_b_l[3], *pf_a_r[3], *pf_b_r[3];   // device tables [face][K] / [face], faces column-major over the other axes
    int K_l[3], K_r[3];
    T a_l[3][kStarMaxK], a_r[3][kStarMaxK];
    T b_l[3], b_r[3];
    T azl_pad[NQ], azr_pad[NQ];  // march-axis BC stencils: a_l left-aligned, a_r right-aligned, zero padded to NQ
    T w[3][NQ];                  // interior stencil, zero padded to the template radius
    const T* tab[3];             // TABLE variants: merged per-row interior weights [n_axis][NQ] (device), indexed by the global row
    T bw[3][2][R][TB];           // ghost-touching rows [axis][low/high][row][tap] (one-sided boundary rows and the interior
                                 // rows next to them): low rows left-aligned (tap k <-> q[k]), high rows right-aligned
                                 // (tap k <-> q[n+2-TB+k]; row i is global row n-nedge+i), zero padded
};

struct StarConfig {
    CUtensorMap tmap;
    std::vector<unsigned char> params;
    int R = 0;
    bool mid = false;
    int py = 2;              // rows (or x segments) per thread
    int nwy = 8;             // warps per CTA (tile rows = nwy * py)
    int mask = 0;            // bit a: an operator acts along kernel axis a (x, mid, march)
    bool table = false;      // per-row weight tables (non-uniform grids, upwind, several operators per axis)
    int nedge_march = 0;     // rows per march-axis face that touch a ghost
    int zchunk_pref = 0;
    int zchunk_max = 0;      // experiments (DEO_STAR_ZCHUNK): upper bound on the planes per CTA along the march axis
    int l2promo = 3;         // CUtensorMapL2promotion of the tensor map (DEO_TMA_L2PROMO)
    const int* halo_flag = nullptr;   // per launch (slab plans): device word the communication stream sets to halo_expect
    int halo_expect = 0, halo_sides = 0;
    int group = -1;          // tiles per launch-order group (0: one wave; < 0: plain order) (DEO_STAR_GROUP)
    int sm_count = 0;
    int xshift = 0, yshift = 0;   // persistent kernel: tile origins shifted by this many elements (see Star2Launch)
    int xres = 0;                 // xshift must equal this modulo the vector length (1: input padded along the contiguous axis)
    bool loader = false;     // persistent kernel: element copies by the helper warps instead of the tensor map (odd row pitch, pre-padded contiguous axis)
    bool scalar_io = false;  // persistent kernel: du rows not 16-byte aligned
    int in_dims[3] = {1, 1, 1};   // input extents in kernel-axis order (x, mid, march)
    bool accumulate = false; // du += result (persistent kernel only)
    bool axpy = false;       // du = u + dt * result, per launch (persistent kernel only)
    double dt = 0.0;
    bool v2 = true;          // persistent warp-specialised kernel (kernel_star2.cuh); false: the first-generation kernel (DEO_STAR_V=1)
    unsigned long long halo_timeout_ns = 30ull * 1000000000ull;   // slab launches: bound of the in-kernel wait for the neighbours' halo planes
};

// ---- PTX wrappers -----------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void mbar_arrive_u32(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait_u32(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}" ::"r"(bar), "r"(parity) : "memory");
}
// calls f(false_type, Face, integral_constant<int, 0>) ... f(false_type, Face, integral_constant<int, N-1>)
template <int N, class Face, int U = 0, class F>
__device__ __forceinline__ void unroll_steps(F& f) {
    if constexpr (U < N) {
        f(std::false_type{}, Face{}, std::integral_constant<int, U>{});
        unroll_steps<N, Face, U + 1>(f);
    }
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(smem_u32(dst)),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}

template <typename T> struct Vec;
template <> struct Vec<double> { using type = double2; static constexpr int N = 2; };
template <> struct Vec<float> { using type = float4; static constexpr int N = 4; };

template <typename T, int N>
__device__ __forceinline__ void ld_vec(const T* p, T (&out)[N]) {
    using V = typename Vec<T>::type;
    const V v = *reinterpret_cast<const V*>(p);
    if constexpr (N == 2) { out[0] = v.x; out[1] = v.y; }
    else { out[0] = v.x; out[1] = v.y; out[2] = v.z; out[3] = v.w; }
}
template <typename T, int N>
__device__ __forceinline__ void st_vec(T* p, const T (&in)[N]) {
    using V = typename Vec<T>::type;
    V v;
    if constexpr (N == 2) { v.x = in[0]; v.y = in[1]; }
    else { v.x = in[0]; v.y = in[1]; v.z = in[2]; v.w = in[3]; }
    *reinterpret_cast<V*>(p) = v;
}

// Geometry shared by host and device.
template <typename T, int R, int PY, int NWY, bool MID>
struct StarGeom {
    static constexpr int VEC = Vec<T>::N;
    static constexpr int HX = ((R + VEC - 1) / VEC) * VEC;        // x halo rounded up: every window load is one 16 B vector
    static constexpr int TX = MID ? 32 * VEC : 32 * VEC * NWY * PY;
    static constexpr int TY = MID ? NWY * PY : 1;
    static constexpr int PITCH = TX + 2 * HX;
    static constexpr int ROWS = MID ? TY + 2 * R : 1;
    static constexpr int BOXW = 256;                              // TMA box limit per dimension (elements)
    static constexpr int NBOX = MID ? 1 : (PITCH + BOXW - 1) / BOXW;
    static constexpr int PLANE = MID ? PITCH * ROWS : NBOX * BOXW;   // elements written per plane
    static constexpr int PLANE_BYTES = ((PLANE * (int)sizeof(T) + 127) / 128) * 128;
#ifndef DEO_STAR_AHEAD
#define DEO_STAR_AHEAD 4
#endif
    static constexpr int NS_WANT = R + 2 + DEO_STAR_AHEAD;        // ring: planes z..z+R live, AHEAD in flight, 1 being drained
    static constexpr int CTAS_PER_SM = (PY <= 2 && NWY <= 8) ? 2 : 1;
    static constexpr int NS_FIT = ((CTAS_PER_SM == 2 ? 110 : 220) * 1024) / PLANE_BYTES;
    static constexpr int NS = NS_WANT < NS_FIT ? NS_WANT : NS_FIT;
    static constexpr int NQ = 2 * R + 1;
    static constexpr int THREADS = NWY * 32;
    static constexpr size_t SMEM = (size_t)NS * PLANE_BYTES + 2 * NS * sizeof(uint64_t);
    static constexpr int TAB_ZMAX = 64;                            // TABLE variants: march-axis rows staged per CTA (chunk bound)
    static constexpr size_t SMEM_TABLE = SMEM + (size_t)(TY + TAB_ZMAX) * (2 * R + 1) * sizeof(T);
    static_assert(NS >= R + 3, "ring too small");
};


// x halo of one vector: exactly R values on each side of the VEC own values, as 16 B vector loads where the
// address is 16 B aligned and one scalar load for the odd element (Float64, odd R).
template <typename T, int R>
__device__ __forceinline__ void load_x_halo(const T* own, T (&xw)[Vec<T>::N + 2 * R]) {
    constexpr int VEC = Vec<T>::N;
    if constexpr (VEC == 2) {
        constexpr int NV = R / 2;
        if constexpr (R % 2 == 1) { xw[0] = own[-R]; xw[R + VEC + R - 1] = own[VEC + R - 1]; }
#pragma unroll
        for (int c = 0; c < NV; ++c) {
            T t[2];
            ld_vec<T, 2>(own - 2 * NV + 2 * c, t);
            xw[(R % 2) + 2 * c] = t[0]; xw[(R % 2) + 2 * c + 1] = t[1];
            ld_vec<T, 2>(own + VEC + 2 * c, t);
            xw[R + VEC + 2 * c] = t[0]; xw[R + VEC + 2 * c + 1] = t[1];
        }
    } else {
        T l[4], r[4];
        ld_vec<T, 4>(own - 4, l);
        ld_vec<T, 4>(own + 4, r);
#pragma unroll
        for (int i = 0; i < R; ++i) { xw[i] = l[4 - R + i]; xw[R + VEC + i] = r[i]; }
    }
}

template <typename T, int R, int PY, int NWY, bool MID, int MASK, bool TABLE>
__global__ void __launch_bounds__(NWY * 32, ((PY <= 2 && NWY <= 8) ? 2 : 1))
k_star(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ StarParams<T, R> S,
       const T* __restrict__ u, T* __restrict__ du, int z_begin, int z_end, int zchunk, int tiles_x, int tiles_xy, int group,
       const int* __restrict__ halo_flag, int halo_expect, int halo_sides) {
    using G = StarGeom<T, R, PY, NWY, MID>;
    constexpr int VEC = G::VEC, HX = G::HX, PITCH = G::PITCH, NS = G::NS, NQ = G::NQ, TB = 2 * R + 2;
    constexpr int XW = VEC + 2 * R;                        // x window of one vector: coordinates gx-R .. gx+VEC-1+R
    constexpr int PLANE_ELEMS = G::PLANE_BYTES / (int)sizeof(T);
    constexpr unsigned FULL = 0xffffffffu;
    // which kernel axes carry an operator is a compile-time property of the variant (MASK bit a = kernel axis a)
    constexpr bool has_x = (MASK & 1) != 0, has_y = MID && (MASK & 2) != 0, has_z = (MASK & 4) != 0;

    extern __shared__ __align__(1024) unsigned char smem_raw[];
    T* const planes = reinterpret_cast<T*>(smem_raw);      // ring of NS planes
    uint64_t* const full = reinterpret_cast<uint64_t*>(smem_raw + (size_t)NS * G::PLANE_BYTES);
    uint64_t* const empty = full + NS;
    T* const sWy = reinterpret_cast<T*>(empty + NS);       // TABLE: [TY][NQ] mid-axis rows of this tile
    T* const sWz = sWy + G::TY * NQ;                       // TABLE: [zc1 - zc0][NQ] march-axis rows of this chunk

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // Launch order: tiles are taken in groups of `group` consecutive tiles; inside a group the march-axis chunk is the
    // slow index.  group == all tiles is the plain order (every tile of chunk 0, then chunk 1, ...), which measured
    // fastest on B200; smaller groups (a tile's chunks about one wave apart, so that the 2R planes two consecutive
    // chunks share could hit in L2) were 8 % slower on the 1024^3 case and are kept as an experiment knob only.
    int tile, chunk;
    {
        const int lin = blockIdx.x, nchunks = (z_end - z_begin + zchunk - 1) / zchunk;
        const int per_group = group * nchunks;
        const int g = lin / per_group, rem = lin - g * per_group;
        const int gsz = min(group, tiles_xy - g * group);
        chunk = rem / gsz;
        tile = g * group + (rem - chunk * gsz);
        // slab plans (one launch for the whole slab): the first and last chunk read halo planes that arrive over NVLink
        // while the launch is running -- they are scheduled last (plain order only; the host guarantees nchunks >= 3)
        if (halo_flag != nullptr) chunk = chunk < nchunks - 2 ? chunk + 1 : (chunk == nchunks - 2 ? 0 : nchunks - 1);
    }
    const int tx0 = (tile % tiles_x) * G::TX;
    const int ty0 = MID ? (tile / tiles_x) * G::TY : 0;
    const int zc0 = z_begin + chunk * zchunk;
    const int zc1 = min(zc0 + zchunk, z_end);
    if (zc0 >= zc1) return;
    const int p_first = zc0 - R;                           // first plane streamed (local index)
    const int n_planes = (zc1 - 1 + R) - p_first + 1;

    // TMA producer = thread 0 (inline): plane ring index kk -> slot kk % NS
    auto issue_plane = [&](int kk) {
        const int slot = kk % NS;
        mbar_expect_tx(&full[slot], (uint32_t)(G::PLANE * sizeof(T)));
        const int pz = p_first + kk + S.in_off_z;
        T* dst = planes + (size_t)slot * PLANE_ELEMS;
        if constexpr (MID) {
            tma_load_3d(dst, &tmap, &full[slot], tx0 - HX, ty0 - R, pz);
        } else {
#pragma unroll
            for (int b = 0; b < G::NBOX; ++b)              // the strip is wider than one TMA box: fixed-width pieces
                tma_load_3d(dst + b * G::BOXW, &tmap, &full[slot], tx0 - HX + b * G::BOXW, 0, pz);
        }
    };
    if (threadIdx.x == 0) {
        for (int s = 0; s < NS; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], NWY); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        if (halo_flag != nullptr) {
            // wait until the neighbours' halo planes of this application have landed: halo_flag[0] (low side) and
            // halo_flag[1] (high side) are written over NVLink after the planes (bounded: a lost exchange must fail
            // loudly, never hang the GPU)
#pragma unroll 1
            for (int side = 0; side < 2; ++side) {
                const bool need = side == 0 ? ((halo_sides & 1) && zc0 - R < z_begin) : ((halo_sides & 2) && zc1 + R > z_end);
                if (!need) continue;
                int seen, spins = 0;
                do {
                    asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(seen) : "l"(halo_flag + side) : "memory");
                    if (seen - halo_expect >= 0) break;
                    __nanosleep(200);
                } while (++spins < (1 << 24));
                if (seen - halo_expect < 0) __trap();
            }
            asm volatile("fence.proxy.async;" ::: "memory");
        }
        for (int k0 = 0; k0 < NS && k0 < n_planes; ++k0) issue_plane(k0);
    }
    if constexpr (TABLE) {
        if constexpr (has_y) {
            for (int i = threadIdx.x; i < G::TY * NQ; i += G::THREADS) {
                const int gyr = min(ty0 + i / NQ, S.ny - 1);
                sWy[i] = __ldg(S.tab[1] + (long long)gyr * NQ + i % NQ);
            }
        }
        if constexpr (has_z) {
            for (int i = threadIdx.x; i < (zc1 - zc0) * NQ; i += G::THREADS)
                sWz[i] = __ldg(S.tab[2] + (long long)(zc0 + S.row0_z + i / NQ) * NQ + i % NQ);
        }
    }
    __syncthreads();

    const int wy = warp;
    const int nx = S.nx, ny = S.ny;
    // this thread's PY vectors: global x, y, and the offset of the vector inside a shared-memory plane
    int gx[PY], gy[PY], soff[PY];
    bool live[PY];
    T* optr[PY];                                           // output pointer of each vector at plane zc0
#pragma unroll
    for (int j = 0; j < PY; ++j) {
        if constexpr (MID) {
            gx[j] = tx0 + lane * VEC;
            gy[j] = ty0 + wy * PY + j;
            soff[j] = (R + wy * PY + j) * PITCH + HX + lane * VEC;
        } else {
            const int seg = (j * NWY + wy) * 32 + lane;
            gx[j] = tx0 + seg * VEC;
            gy[j] = 0;
            soff[j] = HX + seg * VEC;
        }
        live[j] = gx[j] < nx && gy[j] < ny;
        optr[j] = du + (long long)gx[j] + (long long)gy[j] * S.osy + (long long)zc0 * S.osz;
    }
    // TABLE: the x-axis weights of this thread's own points stay in registers for the whole chunk
    T wx[(TABLE && has_x) ? (MID ? 1 : PY) : 1][(TABLE && has_x) ? VEC : 1][(TABLE && has_x) ? NQ : 1];
    if constexpr (TABLE && has_x) {
#pragma unroll
        for (int j = 0; j < (MID ? 1 : PY); ++j)
#pragma unroll
            for (int v = 0; v < VEC; ++v)
#pragma unroll
                for (int t = 0; t < NQ; ++t) wx[j][v][t] = __ldg(S.tab[0] + (long long)min(gx[j] + v, nx - 1) * NQ + t);
    }
    // CTA-uniform face flags: only tiles on a face execute any edge code
    const bool xlo_tile = has_x && tx0 == 0, xhi_tile = has_x && tx0 + G::TX >= nx;
    const bool ylo_tile = has_y && ty0 == 0, yhi_tile = has_y && ty0 + G::TY >= ny;
    const int ex = S.nedge[0], ey = S.nedge[1], ez = S.nedge[2];   // rows per face whose stencil touches the ghost

    // register queue: zq[j][v][t] = plane z-R+t of this thread's columns (t = R is the centre plane)
    T zq[PY][VEC][NQ];
#pragma unroll
    for (int j = 0; j < PY; ++j)
#pragma unroll
        for (int v = 0; v < VEC; ++v)
#pragma unroll
            for (int t = 0; t < NQ; ++t) zq[j][v][t] = T(0);

    bool all_live = true;
#pragma unroll
    for (int j = 0; j < PY; ++j) all_live = all_live && live[j];
    // CTA-uniform: a tile that touches no x/y face and lies fully inside the array runs the plain step
    const bool cta_plain = !(xlo_tile || xhi_tile || ylo_tile || yhi_tile) && __syncthreads_and(all_live ? 1 : 0);

    // One step = acquire plane z+R, compute and store centre plane z.  EDGE = false is the steady-state body with
    // no boundary logic at all; EDGE = true adds priming, face tiles, partial tiles and the march-axis faces.
    // ring state, advanced once per step: slot/parity of the plane being acquired, slot of the centre plane
    int slot_a = 0, par_a = 0, slot_c = NS - R;           // slot_c trails slot_a by R (mod NS); meaningful once z >= zc0
    int z = zc0 - 2 * R, ka = 0;                           // centre plane of the current step (local index), its acquire index
    T* ocur[PY];                                           // output pointer of each vector at the centre plane of the current step
#pragma unroll
    for (int j = 0; j < PY; ++j) ocur[j] = optr[j] - (long long)(2 * R) * S.osz;
    const uint32_t full_u32 = smem_u32(full), empty_u32 = smem_u32(empty);

    // ROT < 0: the queue is shifted every step (zq[t] = plane z-R+t).  ROT = u >= 0 (plain steps, unrolled NQ times):
    // the queue rotates by renaming instead: the new plane overwrites physical slot u, logical tap t lives in
    // physical slot (u + 1 + t) % NQ; after NQ such steps the layout is the shifted one again.
    // ZEDGE: priming / chunk-end / march-axis-face logic (shift mode only).  FACE: the tile touches an x or y face or is
    // partial: x/y fix-up blocks and predicated stores.
    auto step = [&](auto zedge_tag, auto face_tag, auto rot_tag) {
        constexpr bool EDGE = decltype(zedge_tag)::value;
        constexpr bool FACE = decltype(face_tag)::value;
        constexpr int ROT = decltype(rot_tag)::value;
        auto P = [](int t) constexpr { return ROT < 0 ? t : (ROT + 1 + t) % NQ; };
        // --- acquire plane z+R ------------------------------------------------------------------------
        {
            mbar_wait_u32(full_u32 + 8u * slot_a, par_a);
            const T* pn = planes + slot_a * PLANE_ELEMS;
#pragma unroll
            for (int j = 0; j < PY; ++j) {
                T val[VEC];
                ld_vec<T, VEC>(pn + soff[j], val);
#pragma unroll
                for (int v = 0; v < VEC; ++v) {
                    if constexpr (ROT < 0) {
#pragma unroll
                        for (int t = 0; t < NQ - 1; ++t) zq[j][v][t] = zq[j][v][t + 1];
                        zq[j][v][NQ - 1] = val[v];
                    } else {
                        zq[j][v][ROT] = val[v];
                    }
                }
            }
            if (EDGE && (ka < R || ka >= n_planes - R)) {  // planes outside the chunk only feed the queue: free the slot now
                __syncwarp();
                if (lane == 0) mbar_arrive_u32(empty_u32 + 8u * slot_a);
            }
        }
        if (!EDGE || z >= zc0) {
            // --- centre plane z ------------------------------------------------------------------------
            const T* pl = planes + slot_c * PLANE_ELEMS;
            const int gz = z + S.row0_z;
            T tot[PY][VEC];
            // ================= x operator: window = [R halo | VEC own (already in the queue) | R halo] =================
            if constexpr (has_x) {
#pragma unroll
                for (int j = 0; j < PY; ++j) {
                    T xw[XW];
                    load_x_halo<T, R>(pl + soff[j], xw);
#pragma unroll
                    for (int v = 0; v < VEC; ++v) xw[R + v] = zq[j][v][P(R)];
#pragma unroll
                    for (int v = 0; v < VEC; ++v) {
                        T a = T(0);
#pragma unroll
                        for (int t = 0; t < NQ; ++t) {
                            if constexpr (TABLE) a = fma_t(wx[MID ? 0 : j][v][t], xw[v + t], a);
                            else a = fma_t(S.w[0][t], xw[v + t], a);
                        }
                        tot[j][v] = a;
                    }
                }
                // rows whose stencil touches the x ghost (x < ex, x >= nx - ex): full row sums from the tile, in q order
                if (FACE && (xlo_tile || xhi_tile)) {
                    if constexpr (MID) {
                        // cooperative: lane e computes one whole row sum, the owner lane picks it up by shuffle
                        constexpr int NE = PY * R;         // items per face: (tile row jj, edge row r)
                        const int side = lane / NE, jj = (lane % NE) / R, r = lane % R;
                        T res = T(0);
                        if (side < 2 && r < ex && (side ? xhi_tile : xlo_tile)) {
                            const T* row = pl + (R + wy * PY + jj) * PITCH + HX - tx0;     // row[x] = value at global x
                            const T* w = S.bw[0][side][r];
                            const int K = side ? S.K_r[0] : S.K_l[0];
                            const T* a = side ? S.a_r[0] : S.a_l[0];
                            const T* arow = row + (side ? nx - K : 0);
                            T g = T(0);
#pragma unroll 1
                            for (int k = 0; k < K; ++k) g = fma_t(a[k], arow[k], g);
                            g += side ? S.b_r[0] : S.b_l[0];
                            // all taps are loaded first (independent shared-memory loads), then summed in q order
                            T q[TB];
                            const T* qrow = side ? row + (nx + 1 - TB) : row - 1;   // qrow[k] = q[k] (low) / q[n+2-TB+k] (high)
#pragma unroll
                            for (int k = 0; k < TB; ++k) q[k] = ((!side && k == 0) || (side && k == TB - 1)) ? g : qrow[k];
#pragma unroll
                            for (int k = 0; k < TB; ++k) res = fma_t(w[k], q[k], res);
                        }
                        // owners: global x = i (low face) / nx-ex+i (high face), i < ex <= R; row j of this warp
#pragma unroll
                        for (int j = 0; j < PY; ++j) {
#pragma unroll
                            for (int i = 0; i < R; ++i) {
                                if (xlo_tile) {            // CTA-uniform
                                    const T val = __shfl_sync(FULL, res, j * R + i);
#pragma unroll
                                    for (int v = 0; v < VEC; ++v) if (i < ex && gx[j] + v == i) tot[j][v] = val;
                                }
                                if (xhi_tile) {
                                    const T val = __shfl_sync(FULL, res, NE + j * R + i);
#pragma unroll
                                    for (int v = 0; v < VEC; ++v) if (i < ex && gx[j] + v == nx - ex + i) tot[j][v] = val;
                                }
                            }
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < PY; ++j) {
                            if (!live[j] || !((xlo_tile && gx[j] < ex) || (xhi_tile && gx[j] + VEC - 1 >= nx - ex))) continue;
#pragma unroll
                            for (int v = 0; v < VEC; ++v) {
                                const int xg = gx[j] + v;
                                const bool high = xg >= nx - ex;
                                if (!(high || xg < ex)) continue;
                                const T* row = pl + HX - tx0;
                                const T* w = S.bw[0][high ? 1 : 0][high ? xg - (nx - ex) : xg];
                                const int K = high ? S.K_r[0] : S.K_l[0];
                                const T* a = high ? S.a_r[0] : S.a_l[0];
                                T g = T(0);
#pragma unroll 1
                                for (int k = 0; k < K; ++k) g = fma_t(a[k], row[(high ? nx - K : 0) + k], g);
                                g += high ? S.b_r[0] : S.b_l[0];
                                T res = T(0);
#pragma unroll 1
                                for (int k = 0; k < TB; ++k) {
                                    const int c = high ? nx + 1 - TB + k : k - 1;
                                    res = fma_t(w[k], (c == -1 || c == nx) ? g : row[c], res);
                                }
                                tot[j][v] = res;
                            }
                        }
                    }
                }
            }
            // ================= y operator: 2R halo rows loaded once, own rows from the queue =================
            if constexpr (has_y) {
                T acc[PY][VEC];
#pragma unroll
                for (int j = 0; j < PY; ++j)
#pragma unroll
                    for (int v = 0; v < VEC; ++v) acc[j][v] = T(0);
#pragma unroll
                for (int r = 0; r < PY + 2 * R; ++r) {
                    T row[VEC];
                    if (r >= R && r < R + PY) {
#pragma unroll
                        for (int v = 0; v < VEC; ++v) row[v] = zq[r - R][v][P(R)];
                    } else {
                        ld_vec<T, VEC>(pl + soff[0] + (r - R) * PITCH, row);
                    }
#pragma unroll
                    for (int j = 0; j < PY; ++j) {
                        const int t = r - j;
                        if (t >= 0 && t < NQ) {
                            const T wyt = TABLE ? sWy[(wy * PY + j) * NQ + t] : S.w[1][t];
#pragma unroll
                            for (int v = 0; v < VEC; ++v) acc[j][v] = fma_t(wyt, row[v], acc[j][v]);
                        }
                    }
                }
                // rows whose stencil touches the y ghost: warp-uniform (a tile row belongs to one warp)
                if (FACE && (ylo_tile || yhi_tile)) {
                    const T* colbase = pl + (R - ty0) * PITCH + HX + lane * VEC;   // colbase[y*PITCH + v] = value at global row y
#pragma unroll
                    for (int j = 0; j < PY; ++j) {
                        const bool lo = ylo_tile && gy[j] < ey;
                        const bool hi = yhi_tile && gy[j] >= ny - ey && gy[j] < ny;
                        if (!(lo || hi)) continue;
                        const T* w = S.bw[1][hi ? 1 : 0][hi ? gy[j] - (ny - ey) : gy[j]];
                        const int K = hi ? S.K_r[1] : S.K_l[1];
                        const T* a = hi ? S.a_r[1] : S.a_l[1];
                        T g[VEC], res[VEC];
#pragma unroll
                        for (int v = 0; v < VEC; ++v) { g[v] = T(0); res[v] = T(0); }
#pragma unroll 1
                        for (int k = 0; k < K; ++k) {
                            T val[VEC];
                            ld_vec<T, VEC>(colbase + ((hi ? ny - K : 0) + k) * PITCH, val);
#pragma unroll
                            for (int v = 0; v < VEC; ++v) g[v] = fma_t(a[k], val[v], g[v]);
                        }
#pragma unroll
                        for (int v = 0; v < VEC; ++v) g[v] += hi ? S.b_r[1] : S.b_l[1];
                        {
                            const T* qcol = colbase + (hi ? ny + 1 - TB : -1) * PITCH;   // qcol[k*PITCH] = q[k] / q[n+2-TB+k]
                            T q[TB][VEC];
#pragma unroll
                            for (int k = 0; k < TB; ++k) {
                                if ((!hi && k == 0) || (hi && k == TB - 1)) {
#pragma unroll
                                    for (int v = 0; v < VEC; ++v) q[k][v] = g[v];
                                } else {
                                    ld_vec<T, VEC>(qcol + k * PITCH, q[k]);
                                }
                            }
#pragma unroll
                            for (int k = 0; k < TB; ++k)
#pragma unroll
                                for (int v = 0; v < VEC; ++v) res[v] = fma_t(w[k], q[k][v], res[v]);
                        }
#pragma unroll
                        for (int v = 0; v < VEC; ++v) acc[j][v] = res[v];
                    }
                }
#pragma unroll
                for (int j = 0; j < PY; ++j)
#pragma unroll
                    for (int v = 0; v < VEC; ++v) tot[j][v] = has_x ? tot[j][v] + acc[j][v] : acc[j][v];
            }
            // ================= march-axis operator from the register queue =================
            // rows whose stencil touches a march-axis ghost take their term from du (see below)
            if constexpr (has_z) {
                const bool z_low_edge = EDGE && gz < ez, z_high_edge = EDGE && gz >= S.nglob_z - ez;
                if (!(z_low_edge || z_high_edge)) {
                    T wz[NQ];
#pragma unroll
                    for (int t = 0; t < NQ; ++t) wz[t] = TABLE ? sWz[(z - zc0) * NQ + t] : S.w[2][t];
#pragma unroll
                    for (int j = 0; j < PY; ++j) {
#pragma unroll
                        for (int v = 0; v < VEC; ++v) {
                            T s = T(0);
#pragma unroll
                            for (int t = 0; t < NQ; ++t) s = fma_t(wz[t], zq[j][v][P(t)], s);
                            tot[j][v] = (has_x || has_y) ? tot[j][v] + s : s;
                        }
                    }
                } else if (z_high_edge) {                  // the term was parked in du when the queue held its planes
#pragma unroll
                    for (int j = 0; j < PY; ++j) {
                        T a[VEC];
#pragma unroll
                        for (int v = 0; v < VEC; ++v) a[v] = T(0);
                        if (live[j]) ld_vec<T, VEC>(ocur[j], a);
#pragma unroll
                        for (int v = 0; v < VEC; ++v) tot[j][v] = (has_x || has_y) ? tot[j][v] + a[v] : a[v];
                    }
                } else if (!(has_x || has_y)) {            // low edge row of a march-axis-only plan: its term arrives later
#pragma unroll
                    for (int j = 0; j < PY; ++j)
#pragma unroll
                        for (int v = 0; v < VEC; ++v) tot[j][v] = T(0);
                }
            }
            // release the centre plane's slot, then store
            __syncwarp();
            if (lane == 0) mbar_arrive_u32(empty_u32 + 8u * slot_c);
#pragma unroll
            for (int j = 0; j < PY; ++j)
                if (!FACE || live[j]) st_vec<T, VEC>(ocur[j], tot[j]);

            if constexpr (has_z) {
                // --- march-axis rows that touch a ghost, from the register queue -------------------------------
                // low rows r < ez need q[0..TB-1] = ghost, planes 0..2R: exactly the queue when the centre is global plane R.
                // Their x/y part is already in du (stored above at the steps gz = r); add the march-axis term now (it is
                // the last operator, so the association matches the reference's sum).
                if (EDGE && gz == R && ez > 0) {
#pragma unroll
                    for (int j = 0; j < PY; ++j) {
                        if (!live[j]) continue;
                        T gl[VEC];
#pragma unroll
                        for (int v = 0; v < VEC; ++v) {
                            T s = T(0);
#pragma unroll
                            for (int m = 0; m < NQ; ++m) s = fma_t(S.azl_pad[m], zq[j][v][m], s);
                            gl[v] = s + S.b_l[2];
                        }
#pragma unroll 1
                        for (int r = 0; r < ez; ++r) {
                            T* dst = ocur[j] + (long long)(r - S.row0_z - z) * S.osz;
                            T old[VEC];
                            ld_vec<T, VEC>(dst, old);
#pragma unroll
                            for (int v = 0; v < VEC; ++v) {
                                T s = fma_t(S.bw[2][0][r][0], gl[v], T(0));
#pragma unroll
                                for (int kk = 1; kk < TB; ++kk) s = fma_t(S.bw[2][0][r][kk], zq[j][v][kk - 1], s);
                                old[v] = old[v] + s;
                            }
                            st_vec<T, VEC>(dst, old);
                        }
                    }
                }
                // high rows need planes n-1-2R..n-1 and the high ghost: the queue when the centre is global plane n-1-R.
                // Their term is parked in du now and picked up (tot + du) when those rows are computed a few steps later.
                if (EDGE && gz == S.nglob_z - 1 - R && ez > 0) {
#pragma unroll
                    for (int j = 0; j < PY; ++j) {
                        if (!live[j]) continue;
                        T gh[VEC];
#pragma unroll
                        for (int v = 0; v < VEC; ++v) {
                            T s = T(0);
#pragma unroll
                            for (int m = 0; m < NQ; ++m) s = fma_t(S.azr_pad[m], zq[j][v][m], s);
                            gh[v] = s + S.b_r[2];
                        }
#pragma unroll 1
                        for (int r = 0; r < ez; ++r) {
                            T* dst = ocur[j] + (long long)(S.nglob_z - ez + r - S.row0_z - z) * S.osz;
                            T out[VEC];
#pragma unroll
                            for (int v = 0; v < VEC; ++v) {
                                T s = T(0);
#pragma unroll
                                for (int kk = 0; kk < TB - 1; ++kk) s = fma_t(S.bw[2][1][r][kk], zq[j][v][kk], s);
                                out[v] = fma_t(S.bw[2][1][r][TB - 1], gh[v], s);
                            }
                            st_vec<T, VEC>(dst, out);
                        }
                    }
                }
            }
        }
        // --- producer (thread 0): after step ka, plane ka-R-1+NS can reuse the slot of the previous centre (ring index
        // ka-R-1), which every warp released when it finished step ka-1
        {
            const int p = ka - R - 1 + NS;
            if (threadIdx.x == 0 && ka >= R + 1 && p < n_planes) {
                mbar_wait(&empty[p % NS], ((p / NS) - 1) & 1);
                issue_plane(p);
            }
        }
        // --- advance the ring and the output pointers ---------------------------------------------------
        ++z; ++ka;
        if (++slot_a == NS) { slot_a = 0; par_a ^= 1; }
        if (++slot_c == NS) slot_c = 0;
#pragma unroll
        for (int j = 0; j < PY; ++j) ocur[j] += S.osz;
    };

    // steps whose centre is a plain plane: computed (z >= zc0), its acquired plane still inside the chunk, and, when a
    // march-axis operator exists, strictly inside (R, n-1-R) in global planes so that no ghost term is read, parked
    // or added in that step
    int zp0 = zc0, zp1 = zc1 - R;
    if constexpr (has_z) {
        zp0 = max(zp0, R + 1 - S.row0_z);
        zp1 = min(zp1, S.nglob_z - 1 - R - S.row0_z);
    }
    using Shift = std::integral_constant<int, -1>;
#pragma unroll 1
    while (z < zc1 && z < zp0) step(std::true_type{}, std::true_type{}, Shift{});
    if (cta_plain) {
#pragma unroll 1
        while (z + NQ <= zp1) unroll_steps<NQ, std::false_type>(step);   // NQ steps per trip: the queue rotates by renaming
    } else {
#pragma unroll 1
        while (z < zp1) step(std::false_type{}, std::true_type{}, Shift{});   // face / partial tiles: x/y fix-ups, no z-edge logic
    }
#pragma unroll 1
    while (z < zc1) step(std::true_type{}, std::true_type{}, Shift{});
}

template <typename T, int R, int PY, int NWY, bool MID, int MASK, bool TABLE>
int32_t launch_variant(const StarConfig& C, const void* u, void* du, long long z0, long long z1, cudaStream_t s) {
    using G = StarGeom<T, R, PY, NWY, MID>;
    const StarParams<T, R>& S = *reinterpret_cast<const StarParams<T, R>*>(C.params.data());
    constexpr size_t SMEM = TABLE ? G::SMEM_TABLE : G::SMEM;
    static bool attr_set = false;
    auto kern = k_star<T, R, PY, NWY, MID, MASK, TABLE>;
    if (!attr_set) {
        DEO_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM));
        attr_set = true;
    }
    // the tensor map addresses the buffer it was encoded for; re-encode when the caller passes another one
    CUtensorMap map = C.tmap;
    const long long len = z1 - z0;
    const long long tiles = (long long)((S.nx + G::TX - 1) / G::TX) * (MID ? (S.ny + G::TY - 1) / G::TY : 1);
    // Chunking of the march axis.  Short chunks keep the CTAs of a wave in step, so that the halo rows/columns a tile
    // shares with its neighbours are still in L2 when the neighbour asks for them (long chunks let face tiles, which
    // do extra work, drift away: measured on B200, DRAM reads grow from 1.1x to 1.6x the field).  Among the chunk
    // lengths <= zchunk_max pick the one with the shortest makespan in plane-steps (each CTA also primes 2R planes);
    // never cut a face's one-sided rows.
    long long zmax = C.zchunk_max > 0 ? C.zchunk_max : len;
    // short march ranges (thin slabs of a multi-GPU run): at least ~6 chunks, so that the per-CTA pipeline fill / drain
    // bubbles of the few waves do not line up (1024x1024x128: 32-plane chunks 0.415 ms, 22-plane chunks 0.391 ms)
    if (C.zchunk_max > 0 && (len + 5) / 6 < zmax) zmax = (len + 5) / 6;
    if (TABLE && zmax > G::TAB_ZMAX) zmax = G::TAB_ZMAX;
    if (zmax < 4 * R + 4) zmax = 4 * R + 4;
    long long zc = len;
    {
        double best = 1e30;
        const long long slots = (long long)C.sm_count * G::CTAS_PER_SM;
        for (long long nch = (len + zmax - 1) / zmax; nch <= len; ++nch) {
            const long long c = (len + nch - 1) / nch;
            if (c < 4 * R + 4 && nch > 1) break;
            const long long nchunks = (len + c - 1) / c;
            const long long last = len - (nchunks - 1) * c;
            if (nchunks > 1 && last < R + 1) continue;
            const long long waves = (tiles * nchunks + slots - 1) / slots;
            const double cost = (double)waves * (double)(c + 2 * R);
            if (cost < best - 1e-12) { best = cost; zc = c; }
            if (c * 2 < zmax) break;                       // do not go below half the bound
        }
        if (TABLE && zc > G::TAB_ZMAX) { set_error("star kernel: march-axis range too short to chunk"); return DEO_ERR_UNSUPPORTED; }
    }
    const long long tiles_x = (S.nx + G::TX - 1) / G::TX, nchunks = (len + zc - 1) / zc;
    const long long slots_all = (long long)C.sm_count * G::CTAS_PER_SM;
    long long group = C.group > 0 ? C.group : (slots_all >= tiles_x ? slots_all / tiles_x * tiles_x : slots_all);
    if (C.group < 0 || group > tiles) group = tiles;                       // legacy order: all tiles of a chunk, then the next chunk
    const bool fused = C.halo_flag != nullptr;                             // slab launch that waits for its halo planes in the kernel
    if (fused) {
        group = tiles;
        if (nchunks < 3) { set_error("star kernel: fused halo launch needs at least 3 chunks"); return DEO_ERR_UNSUPPORTED; }
    }
    DEO_REQUIRE(tiles * nchunks < (1LL << 31), "star kernel: grid too large");
    kern<<<(unsigned)(tiles * nchunks), G::THREADS, SMEM, s>>>(map, S, (const T*)u, (T*)du, (int)z0, (int)z1, (int)zc, (int)tiles_x, (int)tiles,
                                                               (int)group, fused ? C.halo_flag : nullptr, C.halo_expect, C.halo_sides);
    DEO_CUDA(cudaGetLastError());
    return DEO_OK;
}


// One translation unit per radius instantiates these (star_inst_R*.cu), so the variants compile in parallel.
template <typename T, int R>
int32_t star_launch_R(const StarConfig& C, const void* u, void* du, long long z0, long long z1, cudaStream_t s);

template <typename T, int R, int PY, int NWY, bool MID, bool TABLE>
int32_t star_launch_mask(const StarConfig& C, const void* u, void* du, long long z0, long long z1, cudaStream_t s) {
    switch (C.mask) {
        case 1: return launch_variant<T, R, PY, NWY, MID, 1, TABLE>(C, u, du, z0, z1, s);
        case 4: return launch_variant<T, R, PY, NWY, MID, 4, TABLE>(C, u, du, z0, z1, s);
        case 5: return launch_variant<T, R, PY, NWY, MID, 5, TABLE>(C, u, du, z0, z1, s);
    }
    if constexpr (MID) {
        switch (C.mask) {
            case 2: return launch_variant<T, R, PY, NWY, MID, 2, TABLE>(C, u, du, z0, z1, s);
            case 3: return launch_variant<T, R, PY, NWY, MID, 3, TABLE>(C, u, du, z0, z1, s);
            case 6: return launch_variant<T, R, PY, NWY, MID, 6, TABLE>(C, u, du, z0, z1, s);
            case 7: return launch_variant<T, R, PY, NWY, MID, 7, TABLE>(C, u, du, z0, z1, s);
        }
    }
    set_error("star kernel: unsupported operator mask %d", C.mask);
    return DEO_ERR_UNSUPPORTED;
}

template <typename T, int R, int PY, int NWY, bool TABLE>
int32_t star_launch_mid(const StarConfig& C, const void* u, void* du, long long z0, long long z1, cudaStream_t s) {
    return C.mid ? star_launch_mask<T, R, PY, NWY, true, TABLE>(C, u, du, z0, z1, s)
                 : star_launch_mask<T, R, PY, NWY, false, TABLE>(C, u, du, z0, z1, s);
}

#define DEO_STAR_INSTANTIATE(R_)                                                                                              \
    template <typename T, int R>                                                                                              \
    int32_t star_launch_R(const StarConfig& C, const void* u, void* du, long long z0, long long z1, cudaStream_t s) {        \
        if (C.table) return star_launch_mid<T, R, 2, 8, true>(C, u, du, z0, z1, s);                                           \
        if (C.py == 2 && C.nwy == 16) return star_launch_mid<T, R, 2, 16, false>(C, u, du, z0, z1, s);                       \
        if (C.py == 2) return star_launch_mid<T, R, 2, 8, false>(C, u, du, z0, z1, s);                                        \
        return star_launch_mid<T, R, 4, 8, false>(C, u, du, z0, z1, s);                                                       \
    }                                                                                                                          \
    template int32_t star_launch_R<double, R_>(const StarConfig&, const void*, void*, long long, long long, cudaStream_t);    \
    template int32_t star_launch_R<float, R_>(const StarConfig&, const void*, void*, long long, long long, cudaStream_t);

}  // namespace deo
