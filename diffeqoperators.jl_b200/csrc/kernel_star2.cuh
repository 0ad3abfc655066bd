// Persistent, warp-specialised tiled streaming kernel (second generation of kernel_star.cuh; same operand structures,
// same arithmetic, same results bit for bit).
//
//   * ONE CTA per SM, persistent: 16 compute warps + a helper warpgroup (one TMA producer warp, three evaluator warps); the
//     CTA takes work items (x-y tile, march-axis chunk) from a global queue and the producer keeps the TMA ring full ACROSS
//     item boundaries, so there is no per-chunk pipeline fill / drain bubble and no co-resident CTA is needed to hide one.
//   * compute warps run ONE code path for every tile: the 2.5-D register queue always rotates by renaming (the loop
//     is unrolled 2R+1 times), there is no face code in it.  Rows whose stencil touches an x / y ghost (one-sided
//     boundary rows, convolutions.jl:76-118, and the interior rows next to them) are evaluated by the EVALUATOR warps
//     straight from the shared-memory plane as soon as it has landed -- the affine ghost b + a.u[edge]
//     (bc_operators.jl:188-191) included -- and parked in the plane's own halo cells (which hold nothing but the
//     TMA's out-of-bounds zeros on a face tile); the owning compute lane picks the value up with one predicated
//     load.  A `fixed` mbarrier per ring slot orders the two.
//   * tile = 32*VEC x 32 points (Float64 64 x 32, Float32 128 x 32): 34 % halo instead of the 55 % of 64 x 16 tiles.
//   * the march-axis faces stay with the compute warps (the register queue holds exactly the planes those rows
//     need); only the few steps next to a march-axis face run the shifted-queue edge step.
//   * tile origins may be shifted (Star2Launch.xshift / yshift; host side: tiling_host.hpp): extents just above a tile
//     multiple keep a last tile wide enough for its face, and an input padded along the contiguous axis gets 16-byte
//     aligned TMA boxes (du is then accessed element-wise).  Without a tensor map (row pitch not a multiple of 16 bytes)
//     the four helper warps copy the planes with element-wise cp.async instead (Star2Launch.loader).
//
// Arithmetic: acc = fma(w[t], q[t], acc) over the taps in the reference's order (idx = 1..sl), operators summed in
// A.ops order, w = (c*w) pre-multiplied as the reference forms it (convolutions.jl:47).
#pragma once
#include "kernel_star.cuh"
#include "tiling_host.hpp"

// TABLE variants on 3-D tiles: x-axis weights of a thread's own points in registers up to this radius, in shared memory
// above it (the register queue of the larger radii leaves no room)
#ifndef STAR2_WXREG_MID_R
#define STAR2_WXREG_MID_R 0
#endif

namespace deo {

struct Star2Launch {
    int z_begin, z_end, zchunk, nchunks;
    int tiles_x, tiles_xy, n_items, fused;
    int band;                        // tiles per band (a multiple of tiles_x; tiles_xy when the whole layer is one band)
    const int* halo_flag;            // slab plans: [0] low side, [1] high side, written over NVLink after the halo planes
    int halo_expect, halo_sides;
    int* err_word;                   // mapped host word: set when the halo wait expires (the host turns it into DEO_ERR_CUDA)
    unsigned long long timeout_ns;
    int stagger_ns;                  // experiment: CTA b starts b * stagger_ns late
    int pace_cycles;                 // experiment: minimum SM cycles between two plane issues of a CTA
    int xshift, yshift;              // tile origins are (tile_x * TX - xshift, tile_y * TY - yshift).  A shift widens a first / last tile that
                                     // would be too narrow to hold a face's boundary stencil (extents just above a multiple of the tile);
                                     // an odd xshift also puts the TMA boxes of an input padded along the contiguous axis (rows one element
                                     // behind the rows of du) on 16-byte boundaries -- du is then accessed element-wise
    int loader;                      // 0: TMA tensor map; 1: cp.async element copies by the helper warps (any row pitch / element offset)
    int scalar_io;                   // du rows are not 16-byte aligned: element-wise global loads / stores of du
    int in_nx, in_ny, in_nz;         // input extents (loader == 1: bounds of the element copies)
    int ns;                          // ring slots in use
    int st_cs, ld_policy;            // cache hints: streaming stores of du; TMA loads of u with L2 evict_last (1) / evict_first (2)
    unsigned long long* trace;       // debug (DEO_STAR2_TRACE): per item {first plane issued (ns), last plane issued (ns), SM id}
    unsigned int* sched;             // [0] next work item (dynamic scheduling), [1] CTAs finished; both return to 0 at the end of a launch
    int accumulate, axpy;            // du += result (overwrite = false);  du = u + dt * result
    double dt;
};

template <typename T, int R, bool MID>
struct Star2Geom {
    static constexpr int VEC = Vec<T>::N;
    static constexpr int NW = 16, PY = 2;
    static constexpr int HX = ((R + VEC - 1) / VEC) * VEC;        // x halo rounded up: every window load is one 16 B vector
    static constexpr int TX = MID ? 32 * VEC : 32 * VEC * NW * PY;
    static constexpr int TY = MID ? NW * PY : 1;
    static constexpr int PITCH = TX + 2 * HX;
    static constexpr int ROWS = MID ? TY + 2 * R : 1;
    static constexpr int BOXW = 256;                              // TMA box limit per dimension (elements)
    static constexpr int NBOX = MID ? 1 : (PITCH + BOXW - 1) / BOXW;
    static constexpr int PLANE = MID ? PITCH * ROWS : NBOX * BOXW;   // elements written per plane
    static constexpr int PLANE_BYTES = ((PLANE * (int)sizeof(T) + 127) / 128) * 128;
    static constexpr int TAB_ZMAX = 64;                           // TABLE variants: march-axis rows staged per item (chunk bound)
    // one staging buffer: [TY][NQ] mid-axis rows + [TAB_ZMAX][NQ] march-axis rows + (3-D) [NQ][TX] x-axis rows
    static constexpr int TAB_ELEMS = (TY + TAB_ZMAX + (MID ? TX : 0)) * (2 * R + 1);
    static constexpr int NS_FIT = (216 * 1024 - 2 * TAB_ELEMS * (int)sizeof(T)) / PLANE_BYTES;
    static constexpr int NS_CAP = R + 8;
    static constexpr int NS = NS_FIT < NS_CAP ? NS_FIT : NS_CAP;
    static constexpr int NQ = 2 * R + 1;
    static constexpr int THREADS = (NW + 4) * 32;                 // 4 compute warpgroups + 1 helper warpgroup (producer + 3 evaluators)
    // Register split (setmaxnreg).  The launch allocates 640 * 96 registers to the CTA; setmaxnreg.inc can only draw on
    // what setmaxnreg.dec returned to that per-CTA pool: the helper warpgroup gives back 128 * (96 - 56) = 5120, the four
    // compute warpgroups take 512 * (104 - 96) = 4096.
    static constexpr int REGS_COMPUTE = 104, REGS_HELPER = 56;
    static_assert(512 * (REGS_COMPUTE - 96) <= 128 * (96 - REGS_HELPER), "setmaxnreg: the per-CTA register pool would run dry (deadlock)");
    static constexpr int NIQ = 8;                                 // item queue entries (the producer is never more than 3 items ahead)
    static constexpr size_t TAB_OFF = (((size_t)NS * PLANE_BYTES + (3 * NS + NIQ) * sizeof(uint64_t) + NIQ * sizeof(int) + 15) / 16) * 16;
    static constexpr size_t SMEM = TAB_OFF + 16;
    static constexpr size_t SMEM_TABLE = TAB_OFF + 2 * (size_t)TAB_ELEMS * sizeof(T) + 16;   // + double-buffered weight rows of the current item
    static_assert(NS >= R + 4, "ring too small");
};

__device__ __forceinline__ bool mbar_test_u32(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "mbarrier.test_wait.parity.shared::cta.b64 P1, [%1], %2;\n"
        "selp.u32 %0, 1, 0, P1;\n"
        "}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

// 16-byte store with the streaming (evict-first) hint: du is written once and never read back by this kernel, so it should
// not displace the u lines the neighbouring tiles still have to read from L2
template <typename T, int N>
__device__ __forceinline__ void st_vec_cs(T* p, const T (&in)[N]) {
    if constexpr (N == 2) asm volatile("st.global.cs.v2.f64 [%0], {%1, %2};" ::"l"(p), "d"(in[0]), "d"(in[1]) : "memory");
    else asm volatile("st.global.cs.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(in[0]), "f"(in[1]), "f"(in[2]), "f"(in[3]) : "memory");
}
__device__ __forceinline__ void tma_load_3d_hint(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, uint64_t policy) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4, %5}], [%2], %6;" ::"r"(smem_u32(dst)),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "l"(policy)
        : "memory");
}

// Global loads / stores of one vector of du: 16-byte accesses, or element by element (the first `nvalid` ones) when the
// rows of du are not 16-byte aligned (row length not a multiple of the vector) or the tile origins are shifted (elements
// vfirst .. nvalid-1 of the vector lie inside the array).
template <typename T, int N>
__device__ __forceinline__ void gld(const T* p, T (&out)[N], int vfirst, int nvalid, bool scalar) {
    if (!scalar) { ld_vec<T, N>(p, out); return; }
#pragma unroll
    for (int v = 0; v < N; ++v) out[v] = (v >= vfirst && v < nvalid) ? p[v] : T(0);
}
template <typename T, int N>
__device__ __forceinline__ void gst(T* p, const T (&in)[N], int vfirst, int nvalid, bool scalar) {
    if (!scalar) { st_vec<T, N>(p, in); return; }
#pragma unroll
    for (int v = 0; v < N; ++v) if (v >= vfirst && v < nvalid) p[v] = in[v];
}
template <typename T>
__device__ __forceinline__ void cp_async_elem(T* dst, const T* src, bool inb) {
    const uint32_t d = smem_u32(dst);
    const int sz = inb ? (int)sizeof(T) : 0;             // src-size 0: the destination is zero-filled, the source is not read
    if constexpr (sizeof(T) == 8) asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(d), "l"(src), "r"(sz) : "memory");
    else asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(d), "l"(src), "r"(sz) : "memory");
}

// acc[v] = fma(w[v] or w, x[v], acc[v]) over the VEC points of a thread.  Float32 on sm_100 has a packed two-lane FMA
// (fma.rn.f32x2, FFMA2 in SASS: two IEEE FMAs per issue slot, same rounding as two scalar ones); the Float32 variants of this
// kernel are issue-bound (65 % issue active, 38 % FMA pipe on C3), so halving the FMA instruction count is what moves them.
#ifndef STAR2_FFMA2
#define STAR2_FFMA2 1
#endif
__device__ __forceinline__ void fma2_f32(float& a0, float& a1, float w0, float w1, float x0, float x1) {
    asm("{\n"
        ".reg .b64 ra, rw, rx;\n"
        "mov.b64 ra, {%0, %1};\n"
        "mov.b64 rw, {%2, %3};\n"
        "mov.b64 rx, {%4, %5};\n"
        "fma.rn.f32x2 ra, rw, rx, ra;\n"
        "mov.b64 {%0, %1}, ra;\n"
        "}" : "+f"(a0), "+f"(a1) : "f"(w0), "f"(w1), "f"(x0), "f"(x1));
}
template <typename T, int N>
__device__ __forceinline__ void fma_row(T (&acc)[N], T w, const T (&x)[N]) {
    if constexpr (STAR2_FFMA2 && std::is_same<T, float>::value && N == 4) {
        fma2_f32(acc[0], acc[1], w, w, x[0], x[1]);
        fma2_f32(acc[2], acc[3], w, w, x[2], x[3]);
    } else {
#pragma unroll
        for (int v = 0; v < N; ++v) acc[v] = fma_t(w, x[v], acc[v]);
    }
}
template <typename T, int N>
__device__ __forceinline__ void fma_row_w(T (&acc)[N], const T (&w)[N], const T (&x)[N]) {
    if constexpr (STAR2_FFMA2 && std::is_same<T, float>::value && N == 4) {
        fma2_f32(acc[0], acc[1], w[0], w[1], x[0], x[1]);
        fma2_f32(acc[2], acc[3], w[2], w[3], x[2], x[3]);
    } else {
#pragma unroll
        for (int v = 0; v < N; ++v) acc[v] = fma_t(w[v], x[v], acc[v]);
    }
}

// f(false_type, integral_constant<int, U>) for U = 0 .. N-1, leaving early (returns true) as soon as stop() holds
template <int N, int U = 0, class F, class Stop>
__device__ __forceinline__ bool rot_steps(F& f, Stop& stop) {
    f(std::false_type{}, std::integral_constant<int, U>{});
    if (stop()) return true;
    if constexpr (U + 1 < N) return rot_steps<N, U + 1>(f, stop);
    else return false;
}

// One work item: tile origin and march-axis range.
struct Star2Item { int tx0, ty0, zc0, zc1; };
// Item order.  The x-y tiles are grouped into BANDS of about one grid's worth of tiles (whole tile rows); inside a band the
// list runs chunk-major: every tile of the band for chunk 0, the same tiles for chunk 1, ...  Two kinds of reuse follow:
// the CTAs of the grid work on neighbouring tiles of the same chunk at the same time (shared halo rows / columns hit in
// L2), and a tile's next chunk is taken up right when its previous chunk ends, so its 2R priming planes -- the planes the
// previous chunk read last -- are still in L2 too.  Plain chunk-major order over a layer larger than the grid loses the
// second kind (the next chunk of a tile starts several item-times later).
// Slab launches that wait for their halo planes in the kernel (fused): the same order over the chunks 1 .. nchunks-2, then
// the first and the last chunk of every tile -- the only items that read planes arriving over NVLink -- at the very end.
template <int TX, int TY, bool MID>
__device__ __forceinline__ Star2Item star2_item(const Star2Launch& L, int it) {
    const int nch = L.fused ? L.nchunks - 2 : L.nchunks;           // chunks in the banded part of the list
    int chunk, tile;
    if (L.fused && it >= L.tiles_xy * nch) {
        const int r = it - L.tiles_xy * nch;
        chunk = r < L.tiles_xy ? 0 : L.nchunks - 1;
        tile = r < L.tiles_xy ? r : r - L.tiles_xy;
    } else {
        const int per_band = L.band * nch;
        const int b = it / per_band;
        const int first = b * L.band;                              // first tile of the band
        const int bt = min(L.band, L.tiles_xy - first);             // tiles in this band (the last one may be short)
        const int r = it - b * per_band;
        chunk = r / bt;
        tile = first + (r - chunk * bt);
        if (L.fused) chunk += 1;
    }
    Star2Item I;
    I.tx0 = (tile % L.tiles_x) * TX - L.xshift;
    I.ty0 = MID ? (tile / L.tiles_x) * TY - L.yshift : 0;
    I.zc0 = L.z_begin + chunk * L.zchunk;
    I.zc1 = min(I.zc0 + L.zchunk, L.z_end);
    return I;
}

// ---- evaluation of the x / y rows whose stencil touches a ghost, straight from a landed shared-memory plane --------------
// (helper warps).  SIDE 0 = low face, 1 = high face (compile time: the weights are direct constant-bank operands).  Sums run
// over the taps in ascending q order, exactly like the per-point kernel (bitwise equal results).
//
// x: lane <-> tile row.  The value of x = i (low) is parked HX columns to the left of its owner, i.e. in local column i;
//    (column i - tx0 when the tile origins are shifted); the value of x = nx-ex+i (high) HX columns to the right of its owner.  Those cells hold nothing but the TMA's
//    out-of-bounds zeros on a face tile.
template <typename T, int R, bool MID, int SIDE>
__device__ __forceinline__ void star2_fix_x(const StarParams<T, R>& S, T* pl, int tx0, int ty0, int nx, int ny, int ex, int lane, int gzp) {
    using G = Star2Geom<T, R, MID>;
    constexpr int TB = 2 * R + 2, HX = G::HX, PITCH = G::PITCH;
#pragma unroll 1
    for (int rr = lane; rr < G::TY; rr += 32) {
        if (MID && (ty0 + rr >= ny || ty0 + rr < 0)) continue;
        T* rowl = pl + (MID ? (R + rr) * PITCH : 0);      // local row
        const T* row = rowl + HX - tx0;                  // row[x] = value at global x
        const int K = SIDE ? S.K_r[0] : S.K_l[0];
        const T* arow = row + (SIDE ? nx - K : 0);
        T gh = T(0);
        if (S.padded[0]) {                               // pre-padded input: the ghost is the array's own outer layer
            gh = SIDE ? row[nx] : row[-1];
        } else if (S.per_face[0]) {                      // one BC per boundary pencil: faces column-major over (mid, march)
            const long long face = (long long)(ty0 + rr) + (long long)ny * gzp;
            const T* a = (SIDE ? S.pf_a_r[0] : S.pf_a_l[0]) + face * K;
#pragma unroll 1
            for (int kk = 0; kk < K; ++kk) gh = fma_t(__ldg(a + kk), arow[kk], gh);
            gh += __ldg((SIDE ? S.pf_b_r[0] : S.pf_b_l[0]) + face);
        } else {
            const T* a = SIDE ? S.a_r[0] : S.a_l[0];
#pragma unroll 1
            for (int kk = 0; kk < K; ++kk) gh = fma_t(a[kk], arow[kk], gh);
            gh += SIDE ? S.b_r[0] : S.b_l[0];
        }
        const T* qrow = SIDE ? row + (nx + 1 - TB) : row - 1;   // qrow[k] = q[k] (low) / q[n+2-TB+k] (high)
        T q[TB];
#pragma unroll
        for (int kk = 0; kk < TB; ++kk) q[kk] = (kk == (SIDE ? TB - 1 : 0)) ? gh : qrow[kk];
#pragma unroll
        for (int i = 0; i < R; ++i) {
            if (i < ex) {
                T res = T(0);
#pragma unroll
                for (int kk = 0; kk < TB; ++kk) res = fma_t(S.bw[0][SIDE][i][kk], q[kk], res);
                if (!SIDE) rowl[i - tx0] = res;
                else rowl[(nx - ex + i) - tx0 + 2 * HX] = res;
            }
        }
    }
}
// y: lane <-> one 16-byte vector of columns (conflict-free vector loads).  Row y = i (low) is parked R rows above its owner,
//    i.e. in local row i; row ny-ey+i (high) R rows below its owner.
template <typename T, int R, int SIDE>
__device__ __forceinline__ void star2_fix_y(const StarParams<T, R>& S, T* pl, int tx0, int ty0, int nx, int ny, int ey, int lane, int gzp) {
    using G = Star2Geom<T, R, true>;
    constexpr int TB = 2 * R + 2, HX = G::HX, PITCH = G::PITCH, VEC = G::VEC;
    T* colbase = pl + (R - ty0) * PITCH + HX + lane * VEC;      // colbase[y*PITCH + v] = value at global row y
    const int K = SIDE ? S.K_r[1] : S.K_l[1];
    T gh[VEC];
#pragma unroll
    for (int v = 0; v < VEC; ++v) gh[v] = T(0);
    if (S.padded[1]) {                                   // pre-padded input: the ghost row is part of the plane
        ld_vec<T, VEC>(colbase + (SIDE ? ny : -1) * PITCH, gh);
    } else if (S.per_face[1]) {                          // one BC per boundary pencil: faces column-major over (x, march)
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
            const long long face = (long long)min(max(tx0 + lane * VEC + v, 0), nx - 1) + (long long)nx * gzp;
            const T* a = (SIDE ? S.pf_a_r[1] : S.pf_a_l[1]) + face * K;
#pragma unroll 1
            for (int kk = 0; kk < K; ++kk) gh[v] = fma_t(__ldg(a + kk), colbase[((SIDE ? ny - K : 0) + kk) * PITCH + v], gh[v]);
            gh[v] += __ldg((SIDE ? S.pf_b_r[1] : S.pf_b_l[1]) + face);
        }
    } else {
        const T* a = SIDE ? S.a_r[1] : S.a_l[1];
#pragma unroll 1
        for (int kk = 0; kk < K; ++kk) {
            T val[VEC];
            ld_vec<T, VEC>(colbase + ((SIDE ? ny - K : 0) + kk) * PITCH, val);
#pragma unroll
            for (int v = 0; v < VEC; ++v) gh[v] = fma_t(a[kk], val[v], gh[v]);
        }
#pragma unroll
        for (int v = 0; v < VEC; ++v) gh[v] += SIDE ? S.b_r[1] : S.b_l[1];
    }
    const T* qcol = colbase + (SIDE ? ny + 1 - TB : -1) * PITCH;   // qcol[k*PITCH] = q[k] / q[n+2-TB+k]
#pragma unroll
    for (int i = 0; i < R; ++i) {
        if (i < ey) {
            T res[VEC];
#pragma unroll
            for (int v = 0; v < VEC; ++v) res[v] = T(0);
#pragma unroll
            for (int kk = 0; kk < TB; ++kk) {
                T val[VEC];
                if (kk == (SIDE ? TB - 1 : 0)) {
#pragma unroll
                    for (int v = 0; v < VEC; ++v) val[v] = gh[v];
                } else {
                    ld_vec<T, VEC>(qcol + kk * PITCH, val);
                }
#pragma unroll
                for (int v = 0; v < VEC; ++v) res[v] = fma_t(S.bw[1][SIDE][i][kk], val[v], res[v]);
            }
            st_vec<T, VEC>(colbase + (SIDE ? (ny - ey + i) + R : i - R) * PITCH, res);
        }
    }
}

template <typename T, int R, bool MID, int MASK, bool TABLE>
__global__ void __launch_bounds__(Star2Geom<T, R, MID>::THREADS, 1)
k_star2(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ StarParams<T, R> S, const __grid_constant__ Star2Launch L,
        const T* __restrict__ u, T* __restrict__ du) {
    using G = Star2Geom<T, R, MID>;
    constexpr int VEC = G::VEC, HX = G::HX, PITCH = G::PITCH, NS = G::NS, NQ = G::NQ, TB = 2 * R + 2, PY = G::PY, NW = G::NW;
    constexpr int XW = VEC + 2 * R;
    constexpr int PLANE_ELEMS = G::PLANE_BYTES / (int)sizeof(T);
    constexpr unsigned FULL = 0xffffffffu;
    constexpr bool has_x = (MASK & 1) != 0, has_y = MID && (MASK & 2) != 0, has_z = (MASK & 4) != 0;

    extern __shared__ __align__(1024) unsigned char smem_raw[];
    T* const planes = reinterpret_cast<T*>(smem_raw);
    uint64_t* const full = reinterpret_cast<uint64_t*>(smem_raw + (size_t)NS * G::PLANE_BYTES);
    uint64_t* const empty = full + NS;
    uint64_t* const fixedb = empty + NS;
    uint64_t* const itemb = fixedb + NS;                       // item queue: entry q % NIQ is valid once phase q / NIQ of its barrier completes
    volatile int* const itemq = reinterpret_cast<volatile int*>(itemb + G::NIQ);
    T* const tabbuf = reinterpret_cast<T*>(smem_raw + G::TAB_OFF);   // TABLE only (inside SMEM_TABLE)
    const uint32_t full_u32 = smem_u32(full), empty_u32 = smem_u32(empty), fixed_u32 = smem_u32(fixedb), item_u32 = smem_u32(itemb);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < NS; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], NW); mbar_init(&fixedb[s], 1); }
        for (int s = 0; s < G::NIQ; ++s) mbar_init(&itemb[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int ns = L.ns;                                      // ring slots in use (<= NS): the TMA prefetch depth is ns - R - 1 planes
    const int nx = S.nx, ny = S.ny;
    const int ex = S.nedge[0], ey = S.nedge[1], ez = S.nedge[2];

    if (warp >= NW) {
        // ============================== helper warpgroup (warps NW .. NW+3) ==============================
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(G::REGS_HELPER));
        if (L.loader) {
            // ---- generic-alignment mode: no tensor map (row pitch not a multiple of 16 B, or an input shifted by one element
            // against the output: pre-padded contiguous axis).  The four helper warps are equal workers: a ring slot belongs to
            // worker slot % 4, which waits for the slot to be released, copies the plane row by row with element-wise cp.async
            // (zero-fill outside the array), evaluates its ghost-touching rows and arrives on `fixed`.  Warp NW's lane 0 also
            // hands out the work items, two ahead of its own progress.
            const int me = warp - NW;
            int published = 0, next_item = blockIdx.x;
            bool sentinel = false;
            auto publish = [&]() {
                if (sentinel) return;
                const bool last = next_item >= L.n_items;
                itemq[published % G::NIQ] = last ? -1 : next_item;
                mbar_arrive_u32(item_u32 + 8u * (published % G::NIQ));
                ++published;
                if (last) sentinel = true;
                else next_item = (int)gridDim.x + (int)atomicAdd(L.sched, 1u);
            };
            if (me == 0 && lane == 0) { publish(); publish(); }
            int g = 0;
#pragma unroll 1
            for (int q = 0;; ++q) {
                mbar_wait_u32(item_u32 + 8u * (q % G::NIQ), (q / G::NIQ) & 1);
                const int item = itemq[q % G::NIQ];
                if (item < 0) break;
                if (me == 0 && lane == 0) publish();
                const Star2Item I = star2_item<G::TX, G::TY, MID>(L, item);
                const int n = I.zc1 - I.zc0 + 2 * R;
                const bool f_xlo = has_x && I.tx0 <= 0, f_xhi = has_x && I.tx0 + G::TX >= nx;
                const bool f_ylo = has_y && I.ty0 <= 0, f_yhi = has_y && I.ty0 + G::TY >= ny;
                const bool face = f_xlo || f_xhi || f_ylo || f_yhi;
                // per item: this lane's clamped source column of every 32-column chunk, and which of them lie inside the array
                constexpr int NCH = (PITCH + 31) / 32;
                static_assert(NCH <= 32 || !MID, "column mask");
                const int x0 = I.tx0 - HX + S.in_off_x, y0 = MID ? I.ty0 - R + S.in_off_y : 0;
                int coff[MID ? NCH : 1];
                unsigned xin = 0;
                if constexpr (MID) {
#pragma unroll
                    for (int w = 0; w < NCH; ++w) {
                        const int xi = x0 + w * 32 + lane;
                        coff[w] = min(max(xi, 0), L.in_nx - 1);
                        if (xi >= 0 && xi < L.in_nx) xin |= 1u << w;
                    }
                }
#pragma unroll 1
                for (int k = 0; k < n; ++k, ++g) {
                    const int slot = g % ns;
                    if (slot % 4 != me) continue;
                    if (g >= ns) mbar_wait_u32(empty_u32 + 8u * slot, ((g / ns) - 1) & 1);
                    T* pl = planes + (size_t)slot * PLANE_ELEMS;
                    const int pz = I.zc0 - R + k + S.in_off_z;
                    const bool z_in = pz >= 0 && pz < L.in_nz;
                    // Row by row, 32 columns per instruction; every copy of the plane is in flight before the single wait.  Source
                    // addresses are clamped into the array (column offsets once per item, row / plane indices per row), so the
                    // inner loop is one address add and the copy; copies outside the array zero-fill (src-size 0).
                    const T* gplane = u + (long long)min(max(pz, 0), L.in_nz - 1) * S.isz;
                    if constexpr (MID) {
                        const T* grow = gplane + (long long)min(max(y0, 0), L.in_ny - 1) * S.isy;
                        T* srow = pl + lane;
#pragma unroll 1
                        for (int r = 0; r < G::ROWS; ++r) {
                            const int yi = y0 + r;
                            const bool row_in = z_in && yi >= 0 && yi < L.in_ny;
#pragma unroll
                            for (int w = 0; w < NCH; ++w) {
                                if (NCH * 32 == PITCH || w * 32 + lane < PITCH)
                                    cp_async_elem<T>(srow + w * 32, grow + coff[w], row_in && ((xin >> w) & 1u));
                            }
                            srow += PITCH;
                            if (yi >= 0 && yi < L.in_ny - 1) grow += S.isy;
                        }
                    } else {                                   // 2-D strip: one long row per plane
#pragma unroll 1
                        for (int c = lane; c < PITCH; c += 32) {
                            const int xi = x0 + c;
                            const bool inb = z_in && xi >= 0 && xi < L.in_nx;
                            cp_async_elem<T>(pl + c, inb ? gplane + xi : u, inb);
                        }
                    }
                    asm volatile("cp.async.wait_all;" ::: "memory");
                    __syncwarp();
                    if (face && k >= R && k < n - R) {
                        const int gzp = I.zc0 - R + k + S.row0_z;
                        if constexpr (has_x) {
                            if (f_xlo) star2_fix_x<T, R, MID, 0>(S, pl, I.tx0, I.ty0, nx, ny, ex, lane, gzp);
                            if (f_xhi) star2_fix_x<T, R, MID, 1>(S, pl, I.tx0, I.ty0, nx, ny, ex, lane, gzp);
                        }
                        if constexpr (has_y) {
                            if (f_ylo) star2_fix_y<T, R, 0>(S, pl, I.tx0, I.ty0, nx, ny, ey, lane, gzp);
                            if (f_yhi) star2_fix_y<T, R, 1>(S, pl, I.tx0, I.ty0, nx, ny, ey, lane, gzp);
                        }
                    }
                    __syncwarp();
                    if (lane == 0) mbar_arrive_u32(fixed_u32 + 8u * slot);
                }
            }
            if (me == 0 && lane == 0) {
                __threadfence();
                if (atomicAdd(L.sched + 1, 1u) == gridDim.x - 1) { L.sched[0] = 0u; L.sched[1] = 0u; __threadfence(); }
            }
            return;
        }
        if (warp == NW) {
            // ---- producer: one elected lane keeps the TMA ring full, across item boundaries ----------------
            if (lane != 0) return;
            // Work items are handed out dynamically, in list order: whichever CTA is ready takes the next one, so the items
            // in flight always form one contiguous window of the list -- spatial neighbours stay in step and find each
            // other's halo rows / columns in L2 (a static assignment lets slow CTAs fall behind for good: DRAM reads x1.6).
            // The producer publishes each item to the other warps through the item queue.
            int g = 0, q = 0;
            int item = blockIdx.x;
            uint64_t policy = 0;
            long long t_issue = clock64();
            if (L.ld_policy == 1) asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(policy));
            if (L.ld_policy == 2) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
            if (L.stagger_ns > 0) {
                const unsigned long long t0 = global_ns(), dt = (unsigned long long)L.stagger_ns * blockIdx.x;
                while (global_ns() - t0 < dt) __nanosleep(100);
            }
#pragma unroll 1
            for (;;) {
                const bool last = item >= L.n_items;
                itemq[q % G::NIQ] = last ? -1 : item;
                mbar_arrive_u32(item_u32 + 8u * (q % G::NIQ));
                ++q;
                if (last) break;
                const int next = (int)gridDim.x + (int)atomicAdd(L.sched, 1u);   // fetched one item ahead: its latency hides behind this item's planes
                if (L.trace) {
                    unsigned smid;
                    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
                    L.trace[3 * item] = global_ns();
                    L.trace[3 * item + 2] = smid;
                }
                const Star2Item I = star2_item<G::TX, G::TY, MID>(L, item);
                const int n = I.zc1 - I.zc0 + 2 * R;
                if (L.halo_flag != nullptr) {
                    // slab launch: the first / last chunk read halo planes that arrive over NVLink while the kernel runs
                    // (these items are scheduled last).  A lost exchange raises the error word, it never hangs the GPU.
#pragma unroll 1
                    for (int side = 0; side < 2; ++side) {
                        const bool need = side == 0 ? ((L.halo_sides & 1) && I.zc0 - R < L.z_begin) : ((L.halo_sides & 2) && I.zc1 + R > L.z_end);
                        if (!need) continue;
                        const unsigned long long t0 = global_ns();
                        for (;;) {
                            int seen;
                            asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(seen) : "l"(L.halo_flag + side) : "memory");
                            if (seen - L.halo_expect >= 0) break;
                            if (global_ns() - t0 > L.timeout_ns) {
                                if (L.err_word) asm volatile("st.relaxed.sys.global.s32 [%0], %1;" ::"l"(L.err_word), "r"(1) : "memory");
                                break;                 // give up waiting: this application's results are flagged invalid on the host
                            }
                            __nanosleep(200);
                        }
                    }
                    asm volatile("fence.proxy.async;" ::: "memory");
                }
#pragma unroll 1
                for (int k = 0; k < n; ++k, ++g) {
                    const int slot = g % ns;
                    if (g >= ns) mbar_wait_u32(empty_u32 + 8u * slot, ((g / ns) - 1) & 1);
                    if (L.pace_cycles > 0) {
                        while (clock64() - t_issue < (long long)L.pace_cycles) {}
                        t_issue = clock64();
                    }
                    mbar_expect_tx(&full[slot], (uint32_t)(G::PLANE * sizeof(T)));
                    const int pz = I.zc0 - R + k + S.in_off_z;
                    T* dst = planes + (size_t)slot * PLANE_ELEMS;
                    if constexpr (MID) {
                        if (L.ld_policy) tma_load_3d_hint(dst, &tmap, &full[slot], I.tx0 - HX + S.in_off_x, I.ty0 - R + S.in_off_y, pz, policy);
                        else tma_load_3d(dst, &tmap, &full[slot], I.tx0 - HX + S.in_off_x, I.ty0 - R + S.in_off_y, pz);
                    } else {
#pragma unroll
                        for (int b = 0; b < G::NBOX; ++b)
                            tma_load_3d(dst + b * G::BOXW, &tmap, &full[slot], I.tx0 - HX + S.in_off_x + b * G::BOXW, 0, pz);
                    }
                }
                if (L.trace) L.trace[3 * item + 1] = global_ns();
                item = next;
            }
            // the last CTA to get here puts the scheduler words back to zero for the next launch (launches on a stream do
            // not overlap, and every other CTA has made its final fetch before it counted itself as finished)
            __threadfence();
            if (atomicAdd(L.sched + 1, 1u) == gridDim.x - 1) { L.sched[0] = 0u; L.sched[1] = 0u; __threadfence(); }
            return;
        }
        // ---- evaluators (3 warps, ring slots dealt round-robin): the x / y rows of a landed plane whose stencil touches a ghost ----
        const int me = warp - NW - 1;
        int g = 0;
#pragma unroll 1
        for (int q = 0;; ++q) {
            mbar_wait_u32(item_u32 + 8u * (q % G::NIQ), (q / G::NIQ) & 1);
            const int item = itemq[q % G::NIQ];
            if (item < 0) break;
            const Star2Item I = star2_item<G::TX, G::TY, MID>(L, item);
            const int n = I.zc1 - I.zc0 + 2 * R;
            const bool f_xlo = has_x && I.tx0 <= 0, f_xhi = has_x && I.tx0 + G::TX >= nx;
            const bool f_ylo = has_y && I.ty0 <= 0, f_yhi = has_y && I.ty0 + G::TY >= ny;
            const bool face = f_xlo || f_xhi || f_ylo || f_yhi;
#pragma unroll 1
            for (int k = 0; k < n; ++k, ++g) {
                const int slot = g % ns;
                // a ring slot always belongs to the same evaluator: a waiter must see EVERY phase of a barrier, in order
                // (TMA planes can land out of order; waiting for phase m while m-1 is still open returns at once)
                if (slot % 3 != me) continue;
                // Chain per slot use: TMA lands (`full`) -> this warp parks the ghost-touching rows and arrives on `fixed` ->
                // the compute warps acquire (they wait on `fixed` only) ... release (`empty`) -> the producer refills.
                // Every use gets exactly one arrival on each barrier, so no waiter can be lapped by two phases.
                mbar_wait_u32(full_u32 + 8u * slot, (g / ns) & 1);
                if (face && k >= R && k < n - R) {             // only planes that become a centre plane are read by the x / y parts
                    T* pl = planes + (size_t)slot * PLANE_ELEMS;
                    const int gzp = I.zc0 - R + k + S.row0_z;        // global march-axis index of this plane
                    if constexpr (has_x) {
                        if (f_xlo) star2_fix_x<T, R, MID, 0>(S, pl, I.tx0, I.ty0, nx, ny, ex, lane, gzp);
                        if (f_xhi) star2_fix_x<T, R, MID, 1>(S, pl, I.tx0, I.ty0, nx, ny, ex, lane, gzp);
                    }
                    if constexpr (has_y) {
                        if (f_ylo) star2_fix_y<T, R, 0>(S, pl, I.tx0, I.ty0, nx, ny, ey, lane, gzp);
                        if (f_yhi) star2_fix_y<T, R, 1>(S, pl, I.tx0, I.ty0, nx, ny, ey, lane, gzp);
                    }
                    // the parked values are generic-proxy writes into a slot the TMA (async proxy) overwrites later
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                }
                __syncwarp();
                if (lane == 0) mbar_arrive_u32(fixed_u32 + 8u * slot);
            }
        }
        return;
    }
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(G::REGS_COMPUTE));

    // ============================================ compute warps ============================================
    const int wy = warp;
    int slot_a = 0, par_a = 0;                                // ring state of the next plane to acquire (continuous across items)
    T zq[PY][VEC][NQ];
#pragma unroll
    for (int j = 0; j < PY; ++j)
#pragma unroll
        for (int v = 0; v < VEC; ++v)
#pragma unroll
            for (int t = 0; t < NQ; ++t) zq[j][v][t] = T(0);

#pragma unroll 1
    for (int q = 0;; ++q) {
        mbar_wait_u32(item_u32 + 8u * (q % G::NIQ), (q / G::NIQ) & 1);
        const int item = itemq[q % G::NIQ];
        if (item < 0) break;
        const Star2Item I = star2_item<G::TX, G::TY, MID>(L, item);
        const int tx0 = I.tx0, ty0 = I.ty0, zc0 = I.zc0, zc1 = I.zc1;
        // Per-item state is kept small (the register queue needs the registers): the PY vectors of a thread sit at a constant
        // distance from the first one, in shared memory (SOFF_J) and in du (row stride in 3-D, a constant in 2-D strips).
        constexpr int SOFF_J = MID ? PITCH : NW * 32 * VEC;
        int gx[PY], gy[PY], nv[MID ? 1 : PY];                  // nv: elements of the vector inside the array (3-D: the rows share x)
        int vf[MID ? 1 : PY];                                  // first element of the vector inside the array (> 0 only in a shifted first tile)
        bool live[PY];
        int soff0;
        T* ocur0;                                             // output pointer of the first vector at the centre plane of the current step
        auto soff_of = [&](int j) { return soff0 + j * SOFF_J; };
        auto ocur_of = [&](int j) { return MID ? ocur0 + (long long)j * S.osy : ocur0 + j * SOFF_J; };
#pragma unroll
        for (int j = 0; j < PY; ++j) {
            if constexpr (MID) {
                gx[j] = tx0 + lane * VEC;
                gy[j] = ty0 + wy * PY + j;
                if (j == 0) soff0 = (R + wy * PY) * PITCH + HX + lane * VEC;
            } else {
                const int seg = (j * NW + wy) * 32 + lane;
                gx[j] = tx0 + seg * VEC;
                gy[j] = 0;
                if (j == 0) soff0 = HX + seg * VEC;
            }
            nv[MID ? 0 : j] = min(VEC, nx - gx[j]);
            vf[MID ? 0 : j] = max(0, -gx[j]);
            live[j] = gx[j] < nx && gx[j] + VEC > 0 && gy[j] < ny && gy[j] >= 0;
            if (j == 0) ocur0 = du + (long long)gx[0] + (long long)gy[0] * S.osy + (long long)zc0 * S.osz;
        }
        // which of this thread's values come from the helper's evaluation (bit v: element v of the vector)
        const bool xlo_tile = has_x && tx0 <= 0, xhi_tile = has_x && tx0 + G::TX >= nx;
        const bool ylo_tile = has_y && ty0 <= 0, yhi_tile = has_y && ty0 + G::TY >= ny;
        const bool xface = xlo_tile || xhi_tile, yface = ylo_tile || yhi_tile;
        int xsel[MID ? 1 : PY], xoff[MID ? 1 : PY], ysel[PY];  // 3-D: the PY rows of a thread share x
#pragma unroll
        for (int j = 0; j < PY; ++j) {
            if (!MID || j == 0) { xsel[MID ? 0 : j] = 0; xoff[MID ? 0 : j] = 0; }
            ysel[j] = 0;
#pragma unroll
            for (int v = 0; v < VEC; ++v) {
                if (xlo_tile && gx[j] + v < ex) { xsel[MID ? 0 : j] |= 1 << v; xoff[MID ? 0 : j] = -HX; }
                if (xhi_tile && gx[j] + v >= nx - ex && gx[j] + v < nx) { xsel[MID ? 0 : j] |= 1 << v; xoff[MID ? 0 : j] = HX; }
            }
            if (ylo_tile && gy[j] < ey) ysel[j] = -R * PITCH;
            if (yhi_tile && gy[j] >= ny - ey && gy[j] < ny) ysel[j] = R * PITCH;
        }
        // TABLE, 2-D strips: the x-axis weights of this thread's own points stay in registers for the whole item
        // (3-D tiles stage them in shared memory, see below: registers are what the 2.5-D queue needs)
        constexpr bool WXREG = TABLE && has_x && (!MID || STAR2_WXREG_MID_R >= R);
        constexpr int WXJ = MID ? 1 : PY;                         // 3-D: the PY rows of a thread share x
        T wx[WXREG ? WXJ : 1][WXREG ? VEC : 1][WXREG ? NQ : 1];
        if constexpr (WXREG) {
#pragma unroll
            for (int j = 0; j < WXJ; ++j)
#pragma unroll
                for (int v = 0; v < VEC; ++v)
#pragma unroll
                    for (int t = 0; t < NQ; ++t) wx[j][v][t] = __ldg(S.tab[0] + (long long)min(max(gx[j] + v, 0), nx - 1) * NQ + t);
        }
        // TABLE: the mid-axis rows of this tile and the march-axis rows of this chunk are staged in shared memory by the
        // compute warps (double-buffered per item: a warp can only be one item ahead of the slowest one, see the barrier)
        const T* sWy = tabbuf + (q & 1) * G::TAB_ELEMS;
        const T* sWz = sWy + G::TY * NQ;
        const T* sWx = sWz + G::TAB_ZMAX * NQ;                // [NQ][TX]: tap-major, so that a lane's VEC weights of one tap are one 16-byte load
        if constexpr (TABLE) {
            T* wbuf = tabbuf + (q & 1) * G::TAB_ELEMS;
            if constexpr (has_x && MID && !WXREG) {
                for (int i = threadIdx.x; i < G::TX * NQ; i += NW * 32)
                    wbuf[(G::TY + G::TAB_ZMAX) * NQ + (i % NQ) * G::TX + i / NQ] = __ldg(S.tab[0] + (long long)min(max(tx0 + i / NQ, 0), nx - 1) * NQ + i % NQ);
            }
            if constexpr (has_y) {
                for (int i = threadIdx.x; i < G::TY * NQ; i += NW * 32)
                    wbuf[i] = __ldg(S.tab[1] + (long long)min(max(ty0 + i / NQ, 0), ny - 1) * NQ + i % NQ);
            }
            if constexpr (has_z) {
                for (int i = threadIdx.x; i < (zc1 - zc0) * NQ; i += NW * 32)
                    wbuf[G::TY * NQ + i] = __ldg(S.tab[2] + (long long)(zc0 + S.row0_z + i / NQ) * NQ + i % NQ);
            }
            asm volatile("bar.sync 1, %0;" ::"n"(NW * 32) : "memory");
        }

        // ---- priming: the first 2R planes only feed the queue (physical slots 1..2R; the first step writes slot 0) ----
#pragma unroll
        for (int k = 0; k < 2 * R; ++k) {
            mbar_wait_u32(fixed_u32 + 8u * slot_a, par_a);   // landed AND its ghost-touching rows parked (implies `full`)
            const T* pn = planes + slot_a * PLANE_ELEMS;
#pragma unroll
            for (int j = 0; j < PY; ++j) {
                T val[VEC];
                ld_vec<T, VEC>(pn + soff_of(j), val);
#pragma unroll
                for (int v = 0; v < VEC; ++v) zq[j][v][k + 1] = val[v];
            }
            if (k < R) {                                       // never a centre plane: free the slot now
                __syncwarp();
                if (lane == 0) mbar_arrive_u32(empty_u32 + 8u * slot_a);
            }
            if (++slot_a == ns) { slot_a = 0; par_a ^= 1; }
        }
        int slot_c = slot_a - R;                               // centre plane of the first step = R planes behind
        if (slot_c < 0) slot_c += ns;
        int z = zc0;

        // One step = acquire plane z+R, compute and store centre plane z.  ROT = u >= 0: the new plane overwrites
        // physical queue slot u, logical tap t lives in physical slot (u + 1 + t) % NQ (NQ consecutive steps u = 0..NQ-1
        // return to the identity layout).  ROT < 0 (EDGE): the queue is shifted instead (zq[t] = plane z-R+t) and the
        // step carries the march-axis face logic.
        auto step = [&](auto edge_tag, auto rot_tag) {
            constexpr bool EDGE = decltype(edge_tag)::value;
            constexpr int ROT = decltype(rot_tag)::value;
            auto P = [](int t) constexpr { return ROT < 0 ? t : (ROT + 1 + t) % NQ; };
            // --- acquire plane z+R ------------------------------------------------------------------------
            {
                mbar_wait_u32(fixed_u32 + 8u * slot_a, par_a);   // landed AND its ghost-touching rows parked (implies `full`)
                const T* pn = planes + slot_a * PLANE_ELEMS;
#pragma unroll
                for (int j = 0; j < PY; ++j) {
                    T val[VEC];
                    ld_vec<T, VEC>(pn + soff_of(j), val);
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
                if (z >= zc1 - R) {                            // planes past the chunk only feed the queue: free the slot now
                    __syncwarp();
                    if (lane == 0) mbar_arrive_u32(empty_u32 + 8u * slot_a);
                }
            }
            // --- centre plane z ------------------------------------------------------------------------
            const T* pl = planes + slot_c * PLANE_ELEMS;
            const int gz = z + S.row0_z;
            T tot[PY][VEC];
            T pk[EDGE ? PY : 1][EDGE ? VEC : 1];               // EDGE: march-axis term of a high-face row, parked in du earlier
            bool parked = false;
            // ================= x operator: window = [R halo | VEC own (already in the queue) | R halo] =================
            if constexpr (has_x && TABLE && MID && !WXREG) {
                // per-point weights from shared memory, one 16-byte load per tap shared by the PY rows (same x)
                T xw[PY][XW];
#pragma unroll
                for (int j = 0; j < PY; ++j) {
                    load_x_halo<T, R>(pl + soff_of(j), xw[j]);
#pragma unroll
                    for (int v = 0; v < VEC; ++v) { xw[j][R + v] = zq[j][v][P(R)]; tot[j][v] = T(0); }
                }
#pragma unroll
                for (int t = 0; t < NQ; ++t) {
                    T wv[VEC];
                    ld_vec<T, VEC>(sWx + t * G::TX + lane * VEC, wv);
#pragma unroll
                    for (int j = 0; j < PY; ++j) {
                        T xs[VEC];
#pragma unroll
                        for (int v = 0; v < VEC; ++v) xs[v] = xw[j][v + t];
                        fma_row_w<T, VEC>(tot[j], wv, xs);
                    }
                }
            } else if constexpr (has_x) {
#pragma unroll
                for (int j = 0; j < PY; ++j) {
                    T xw[XW];
                    load_x_halo<T, R>(pl + soff_of(j), xw);
#pragma unroll
                    for (int v = 0; v < VEC; ++v) xw[R + v] = zq[j][v][P(R)];
#pragma unroll
                    for (int v = 0; v < VEC; ++v) tot[j][v] = T(0);
#pragma unroll
                    for (int t = 0; t < NQ; ++t) {                    // per point still the taps in ascending order
                        T xs[VEC];
#pragma unroll
                        for (int v = 0; v < VEC; ++v) xs[v] = xw[v + t];
                        if constexpr (TABLE) {
                            T ws[VEC];
#pragma unroll
                            for (int v = 0; v < VEC; ++v) ws[v] = wx[MID ? 0 : j][v][t];
                            fma_row_w<T, VEC>(tot[j], ws, xs);
                        } else {
                            fma_row<T, VEC>(tot[j], S.w[0][t], xs);
                        }
                    }
                }
            }
            if constexpr (has_x) {
                if (xface) {
#pragma unroll
                    for (int j = 0; j < PY; ++j) {
                        if (xsel[MID ? 0 : j] != 0) {
                            T f[VEC];
                            ld_vec<T, VEC>(pl + soff_of(j) + xoff[MID ? 0 : j], f);
#pragma unroll
                            for (int v = 0; v < VEC; ++v) if ((xsel[MID ? 0 : j] >> v) & 1) tot[j][v] = f[v];
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
                        ld_vec<T, VEC>(pl + soff0 + (r - R) * PITCH, row);
                    }
#pragma unroll
                    for (int j = 0; j < PY; ++j) {
                        const int t = r - j;
                        if (t >= 0 && t < NQ) {
                            T wyt;
                            if constexpr (TABLE) wyt = sWy[(wy * PY + j) * NQ + t]; else wyt = S.w[1][t];
                            fma_row<T, VEC>(acc[j], wyt, row);
                        }
                    }
                }
                if (yface) {
#pragma unroll
                    for (int j = 0; j < PY; ++j)
                        if (ysel[j] != 0) ld_vec<T, VEC>(pl + soff_of(j) + ysel[j], acc[j]);   // warp-uniform: a tile row belongs to one warp
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
                    for (int t = 0; t < NQ; ++t) {
                        if constexpr (TABLE) wz[t] = sWz[(z - zc0) * NQ + t]; else wz[t] = S.w[2][t];
                    }
#pragma unroll
                    for (int j = 0; j < PY; ++j) {
                        T sz[VEC];
#pragma unroll
                        for (int v = 0; v < VEC; ++v) sz[v] = T(0);
#pragma unroll
                        for (int t = 0; t < NQ; ++t) {                // per point still the taps in ascending order
                            T xs[VEC];
#pragma unroll
                            for (int v = 0; v < VEC; ++v) xs[v] = zq[j][v][P(t)];
                            fma_row<T, VEC>(sz, wz[t], xs);
                        }
#pragma unroll
                        for (int v = 0; v < VEC; ++v) tot[j][v] = (has_x || has_y) ? tot[j][v] + sz[v] : sz[v];
                    }
                } else if (z_high_edge) {                  // the term was parked in du when the queue held its planes
                    if constexpr (EDGE) {
                        parked = true;
#pragma unroll
                        for (int j = 0; j < PY; ++j) {
#pragma unroll
                            for (int v = 0; v < VEC; ++v) { pk[j][v] = T(0); if (!(has_x || has_y)) tot[j][v] = T(0); }
                            if (live[j]) gld<T, VEC>(ocur_of(j), pk[j], vf[MID ? 0 : j], nv[MID ? 0 : j], L.scalar_io);
                        }
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
            // epilogue: du = result | du += result (overwrite = false, convolutions.jl:17-22) | du = u + dt * result (the
            // explicit-stepper update fused into the store, cf. test/DerivativeOperators/3D_laplacian.jl:20-24)
            if (L.accumulate | L.axpy) {                       // launch-uniform
                const T sc = L.axpy ? (T)L.dt : T(1);
#pragma unroll
                for (int j = 0; j < PY; ++j) {
                    if (!live[j]) continue;
                    T base[VEC];
#pragma unroll
                    for (int v = 0; v < VEC; ++v) base[v] = T(0);
                    if (L.accumulate && !parked) gld<T, VEC>(ocur_of(j), base, vf[MID ? 0 : j], nv[MID ? 0 : j], L.scalar_io);   // a parked term already contains the old du
#pragma unroll
                    for (int v = 0; v < VEC; ++v) {
                        if (L.axpy) base[v] += zq[j][v][P(R)];
                        tot[j][v] = fma_t(sc, tot[j][v], base[v]);
                        if constexpr (EDGE) { if (parked) tot[j][v] += pk[j][v]; }
                    }
                }
            } else if constexpr (EDGE) {
                if (parked) {
#pragma unroll
                    for (int j = 0; j < PY; ++j)
#pragma unroll
                        for (int v = 0; v < VEC; ++v) tot[j][v] = (has_x || has_y) ? tot[j][v] + pk[j][v] : pk[j][v];
                }
            }
#pragma unroll
            for (int j = 0; j < PY; ++j)
                if (live[j]) { if (L.st_cs && !L.scalar_io) st_vec_cs<T, VEC>(ocur_of(j), tot[j]); else gst<T, VEC>(ocur_of(j), tot[j], vf[MID ? 0 : j], nv[MID ? 0 : j], L.scalar_io); }

            if constexpr (has_z && EDGE) {
                // --- march-axis rows that touch a ghost, from the register queue -------------------------------
                // low rows r < ez need q[0..TB-1] = ghost, planes 0..2R: exactly the queue when the centre is global plane R.
                // Their x/y part is already in du (stored above at the steps gz = r); add the march-axis term now (it is
                // the last operator, so the association matches the reference's sum).
                if (gz == R && ez > 0) {
#pragma unroll
                    for (int j = 0; j < PY; ++j) {
                        if (!live[j]) continue;
                        T gl[VEC];
#pragma unroll
                        for (int v = 0; v < VEC; ++v) {
                            T s = T(0);
                            if (S.padded[2]) {              // pre-padded input: the ghost plane is the array's first plane
                                s = __ldg(u + (long long)(max(gx[j] + v, 0) + S.in_off_x) + (long long)(gy[j] + S.in_off_y) * S.isy);
                            } else if (S.per_face[2]) {     // one BC per boundary pencil: faces column-major over (x, mid)
                                const long long face = (long long)min(max(gx[j] + v, 0), nx - 1) + (long long)nx * gy[j];
                                const T* a = S.pf_a_l[2] + face * S.K_l[2];
#pragma unroll
                                for (int m = 0; m < NQ; ++m) if (m < S.K_l[2]) s = fma_t(__ldg(a + m), zq[j][v][m], s);
                                s += __ldg(S.pf_b_l[2] + face);
                            } else {
#pragma unroll
                                for (int m = 0; m < NQ; ++m) s = fma_t(S.azl_pad[m], zq[j][v][m], s);
                                s += S.b_l[2];
                            }
                            gl[v] = s;
                        }
#pragma unroll 1
                        for (int r = 0; r < ez; ++r) {
                            T* dst = ocur_of(j) + (long long)(r - S.row0_z - z) * S.osz;
                            T old[VEC];
                            gld<T, VEC>(dst, old, vf[MID ? 0 : j], nv[MID ? 0 : j], L.scalar_io);
#pragma unroll
                            for (int v = 0; v < VEC; ++v) {
                                T s = fma_t(S.bw[2][0][r][0], gl[v], T(0));
#pragma unroll
                                for (int kk = 1; kk < TB; ++kk) s = fma_t(S.bw[2][0][r][kk], zq[j][v][kk - 1], s);
                                old[v] = L.axpy ? fma_t((T)L.dt, s, old[v]) : old[v] + s;
                            }
                            gst<T, VEC>(dst, old, vf[MID ? 0 : j], nv[MID ? 0 : j], L.scalar_io);
                        }
                    }
                }
                // high rows need planes n-1-2R..n-1 and the high ghost: the queue when the centre is global plane n-1-R.
                // Their term is parked in du now and picked up (tot + du) when those rows are computed a few steps later.
                if (gz == S.nglob_z - 1 - R && ez > 0) {
#pragma unroll
                    for (int j = 0; j < PY; ++j) {
                        if (!live[j]) continue;
                        T gh[VEC];
#pragma unroll
                        for (int v = 0; v < VEC; ++v) {
                            T s = T(0);
                            if (S.padded[2]) {              // pre-padded input: the ghost plane is the array's last plane
                                s = __ldg(u + (long long)(max(gx[j] + v, 0) + S.in_off_x) + (long long)(gy[j] + S.in_off_y) * S.isy +
                                          (long long)(S.nglob_z + 1) * S.isz);
                            } else if (S.per_face[2]) {
                                const long long face = (long long)min(max(gx[j] + v, 0), nx - 1) + (long long)nx * gy[j];
                                const int K = S.K_r[2];
                                const T* a = S.pf_a_r[2] + face * K;
#pragma unroll
                                for (int m = 0; m < NQ; ++m) if (m >= NQ - K) s = fma_t(__ldg(a + (m - (NQ - K))), zq[j][v][m], s);
                                s += __ldg(S.pf_b_r[2] + face);
                            } else {
#pragma unroll
                                for (int m = 0; m < NQ; ++m) s = fma_t(S.azr_pad[m], zq[j][v][m], s);
                                s += S.b_r[2];
                            }
                            gh[v] = s;
                        }
#pragma unroll 1
                        for (int r = 0; r < ez; ++r) {
                            T* dst = ocur_of(j) + (long long)(S.nglob_z - ez + r - S.row0_z - z) * S.osz;
                            T out[VEC];
#pragma unroll
                            for (int v = 0; v < VEC; ++v) {
                                T s = T(0);
#pragma unroll
                                for (int kk = 0; kk < TB - 1; ++kk) s = fma_t(S.bw[2][1][r][kk], zq[j][v][kk], s);
                                out[v] = fma_t(S.bw[2][1][r][TB - 1], gh[v], s);
                            }
                            if (L.accumulate | L.axpy) {           // parked term = [old du +] dt * term
                                T prev[VEC];
#pragma unroll
                                for (int v = 0; v < VEC; ++v) prev[v] = T(0);
                                if (L.accumulate) gld<T, VEC>(dst, prev, vf[MID ? 0 : j], nv[MID ? 0 : j], L.scalar_io);
#pragma unroll
                                for (int v = 0; v < VEC; ++v) out[v] = fma_t(L.axpy ? (T)L.dt : T(1), out[v], prev[v]);
                            }
                            gst<T, VEC>(dst, out, vf[MID ? 0 : j], nv[MID ? 0 : j], L.scalar_io);
                        }
                    }
                }
            }
            // --- advance the ring and the output pointers ---------------------------------------------------
            ++z;
            if (++slot_a == ns) { slot_a = 0; par_a ^= 1; }
            if (++slot_c == ns) slot_c = 0;
            ocur0 += S.osz;
        };

        // Step ranges of this item: [zc0, z_lo) edge steps next to the low march-axis face (global planes 0..R),
        // [z_lo, z_rot) rotated steps, [z_rot, zc1) edge steps up to the high march-axis face.  The rotated run before
        // edge steps is a whole number of rotations, so that the queue is back in the identity layout.
        int z_lo = zc0, z_rot = zc1;
        if constexpr (has_z) {
            z_lo = max(zc0, min(zc1, R + 1 - S.row0_z));
            const int z_hi = min(zc1, max(z_lo, S.nglob_z - 1 - R - S.row0_z));
            z_rot = z_hi == zc1 ? zc1 : z_lo + ((z_hi - z_lo) / NQ) * NQ;
        }
        using Shift = std::integral_constant<int, -1>;
#pragma unroll 1
        while (z < z_lo) step(std::true_type{}, Shift{});
        if (z < z_rot) {
            auto at_end = [&]() { return z == z_rot; };
#pragma unroll 1
            while (!rot_steps<NQ>(step, at_end)) {}
        }
#pragma unroll 1
        while (z < zc1) step(std::true_type{}, Shift{});
    }
}

struct Star2Runtime {            // per process and device: the mapped error word of the in-kernel halo wait, the scheduler words
    int* err_host = nullptr;
    int* err_dev = nullptr;
    unsigned int* sched = nullptr;   // device memory, 2 words, zero between launches
};
Star2Runtime& star2_rt();

template <typename T, int R, bool MID, int MASK, bool TABLE>
int32_t launch_variant2(const StarConfig& C, const void* u, void* du, long long z0, long long z1, cudaStream_t s) {
    using G = Star2Geom<T, R, MID>;
    const StarParams<T, R>& S = *reinterpret_cast<const StarParams<T, R>*>(C.params.data());
    auto kern = k_star2<T, R, MID, MASK, TABLE>;
    static int attr_device = -1;
    int dev = 0;
    DEO_CUDA(cudaGetDevice(&dev));
    if (attr_device != dev) {
        DEO_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(TABLE ? G::SMEM_TABLE : G::SMEM)));
        attr_device = dev;
    }
    const long long len = z1 - z0;
    const long long tiles_x = (S.nx + C.xshift + G::TX - 1) / G::TX;
    const long long tiles = tiles_x * (MID ? (S.ny + C.yshift + G::TY - 1) / G::TY : 1);
    // Chunking of the march axis (tiling_host.hpp): shortest makespan among the chunk lengths <= zchunk_max; the TABLE variants
    // stage at most TAB_ZMAX rows of march-axis weights per item.
    const long long zc = tiling::pick_chunk(len, C.zchunk_max, R, tiles, C.sm_count, TABLE ? (long long)G::TAB_ZMAX : 0);
    Star2Launch Lp{};
    Lp.z_begin = (int)z0; Lp.z_end = (int)z1; Lp.zchunk = (int)zc;
    Lp.nchunks = (int)((len + zc - 1) / zc);
    Lp.tiles_x = (int)tiles_x; Lp.tiles_xy = (int)tiles;
    DEO_REQUIRE(tiles * Lp.nchunks < (1LL << 31), "star kernel: too many work items");
    Lp.n_items = (int)(tiles * Lp.nchunks);
    const bool fused = C.halo_flag != nullptr;
    if (fused && Lp.nchunks < 3) { set_error("star kernel: fused halo launch needs at least 3 chunks"); return DEO_ERR_UNSUPPORTED; }
    Lp.fused = fused ? 1 : 0;
    {
        // band = the largest whole number of tile rows that fits the grid (at least one row); one band when the layer fits
        long long band = tiles;
        const long long env_band = getenv("DEO_STAR2_BAND") ? atoll(getenv("DEO_STAR2_BAND")) : 0;
        if (tiles > C.sm_count) band = (C.sm_count / tiles_x > 0 ? C.sm_count / tiles_x : 1) * tiles_x;
        if (env_band > 0) band = ((env_band + tiles_x - 1) / tiles_x) * tiles_x;
        if (band > tiles) band = tiles;
        Lp.band = (int)band;
    }
    Lp.halo_flag = fused ? C.halo_flag : nullptr;
    Lp.halo_expect = C.halo_expect; Lp.halo_sides = C.halo_sides;
    Lp.err_word = star2_rt().err_dev;
    Lp.sched = star2_rt().sched;
    Lp.stagger_ns = getenv("DEO_STAR2_STAGGER_NS") ? atoi(getenv("DEO_STAR2_STAGGER_NS")) : 0;
    Lp.pace_cycles = getenv("DEO_STAR2_PACE") ? atoi(getenv("DEO_STAR2_PACE")) : 0;
    Lp.st_cs = getenv("DEO_STAR2_ST_CS") ? atoi(getenv("DEO_STAR2_ST_CS")) : 0;
    Lp.ld_policy = getenv("DEO_STAR2_LD_POLICY") ? atoi(getenv("DEO_STAR2_LD_POLICY")) : 0;
    // Prefetch depth: R + 6 ring slots (R + 1 live planes, 5 in flight).  Deeper rings only lengthen the queues in the
    // memory system: the CTAs drift further apart and the halo rows / columns neighbouring tiles share fall out of L2
    // before the second reader arrives (measured on 1024^3: 11 slots 330, 8 slots 354 Gpoints/s).
    Lp.loader = C.loader ? 1 : 0;
    Lp.xshift = C.xshift; Lp.yshift = MID ? C.yshift : 0;
    Lp.scalar_io = C.scalar_io ? 1 : 0;
    Lp.in_nx = C.in_dims[0]; Lp.in_ny = C.in_dims[1]; Lp.in_nz = C.in_dims[2];
    Lp.ns = G::NS < R + 6 ? G::NS : R + 6;
    if (getenv("DEO_STAR2_NS")) { const int v = atoi(getenv("DEO_STAR2_NS")); if (v >= R + 4 && v <= G::NS) Lp.ns = v; }
    Lp.trace = nullptr;
    if (getenv("DEO_STAR2_TRACE")) {                      // debug: dumps the item schedule of this launch to the named file (synchronous)
        static unsigned long long* buf = nullptr;
        static size_t cap = 0;
        if (cap < (size_t)Lp.n_items * 3) { if (buf) cudaFree(buf); cap = (size_t)Lp.n_items * 3; if (cudaMalloc(&buf, cap * 8) != cudaSuccess) { buf = nullptr; cap = 0; } }
        Lp.trace = buf;
    }
    DEO_REQUIRE(Lp.sched != nullptr, "star kernel: scheduler words could not be allocated");
    Lp.timeout_ns = C.halo_timeout_ns;
    Lp.accumulate = C.accumulate ? 1 : 0; Lp.axpy = C.axpy ? 1 : 0; Lp.dt = C.dt;
    const unsigned grid = (unsigned)(Lp.n_items < C.sm_count ? Lp.n_items : C.sm_count);
    if (TABLE && zc > G::TAB_ZMAX) { set_error("star kernel: march-axis range too short to chunk"); return DEO_ERR_UNSUPPORTED; }
    kern<<<grid, G::THREADS, TABLE ? G::SMEM_TABLE : G::SMEM, s>>>(C.tmap, S, Lp, (const T*)u, (T*)du);
    DEO_CUDA(cudaGetLastError());
    if (Lp.trace) {
        std::vector<unsigned long long> h((size_t)Lp.n_items * 3);
        DEO_CUDA(cudaStreamSynchronize(s));
        DEO_CUDA(cudaMemcpy(h.data(), Lp.trace, h.size() * 8, cudaMemcpyDeviceToHost));
        if (FILE* f = fopen(getenv("DEO_STAR2_TRACE"), "w")) {
            fprintf(f, "# tiles_x %d tiles_xy %d nchunks %d zchunk %d\n", Lp.tiles_x, Lp.tiles_xy, Lp.nchunks, Lp.zchunk);
            for (int i = 0; i < Lp.n_items; ++i) fprintf(f, "%d %llu %llu %llu\n", i, h[3 * (size_t)i], h[3 * (size_t)i + 1], h[3 * (size_t)i + 2]);
            fclose(f);
        }
    }
    return DEO_OK;
}

template <typename T, int R>
int32_t star2_launch_R(const StarConfig& C, const void* u, void* du, long long z0, long long z1, cudaStream_t s);

template <typename T, int R, bool MID, bool TABLE>
int32_t star2_launch_mask(const StarConfig& C, const void* u, void* du, long long z0, long long z1, cudaStream_t s) {
    switch (C.mask) {
        case 1: return launch_variant2<T, R, MID, 1, TABLE>(C, u, du, z0, z1, s);
        case 4: return launch_variant2<T, R, MID, 4, TABLE>(C, u, du, z0, z1, s);
        case 5: return launch_variant2<T, R, MID, 5, TABLE>(C, u, du, z0, z1, s);
    }
    if constexpr (MID) {
        switch (C.mask) {
            case 2: return launch_variant2<T, R, MID, 2, TABLE>(C, u, du, z0, z1, s);
            case 3: return launch_variant2<T, R, MID, 3, TABLE>(C, u, du, z0, z1, s);
            case 6: return launch_variant2<T, R, MID, 6, TABLE>(C, u, du, z0, z1, s);
            case 7: return launch_variant2<T, R, MID, 7, TABLE>(C, u, du, z0, z1, s);
        }
    }
    set_error("star kernel: unsupported operator mask %d", C.mask);
    return DEO_ERR_UNSUPPORTED;
}

#define DEO_STAR2_INSTANTIATE(R_)                                                                                             \
    template <typename T, int R>                                                                                              \
    int32_t star2_launch_R(const StarConfig& C, const void* u, void* du, long long z0, long long z1, cudaStream_t s) {       \
        if (C.table) return C.mid ? star2_launch_mask<T, R, true, true>(C, u, du, z0, z1, s)                                 \
                                  : star2_launch_mask<T, R, false, true>(C, u, du, z0, z1, s);                               \
        return C.mid ? star2_launch_mask<T, R, true, false>(C, u, du, z0, z1, s)                                             \
                     : star2_launch_mask<T, R, false, false>(C, u, du, z0, z1, s);                                           \
    }                                                                                                                          \
    template int32_t star2_launch_R<double, R_>(const StarConfig&, const void*, void*, long long, long long, cudaStream_t);   \
    template int32_t star2_launch_R<float, R_>(const StarConfig&, const void*, void*, long long, long long, cudaStream_t);

}  // namespace deo
