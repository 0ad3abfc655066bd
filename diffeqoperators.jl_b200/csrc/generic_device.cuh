// Per-point evaluation of one operator of a plan at one output point (device code shared by the
// per-point kernel and by the edge paths of the tiled kernel).
//
// du[c] = sum_k  row_k(c[axis_k]) . q_k      with q_k the ghost-padded pencil through c along axis_k:
//   q[0]   = ghost_lo = b_l + a_l . u[0:K_l]          (bc_operators.jl:188-191; computed here, never stored)
//   q[j+1] = u[j]
//   q[n+1] = ghost_hi = b_r + a_r . u[n-K_r:n]
// or q taken straight from a pre-padded input (derivative_operator_functions.jl:50-57).
#pragma once
#include "common.hpp"

namespace deo {

template <typename T> __device__ __forceinline__ T fma_t(T a, T b, T c);
template <> __device__ __forceinline__ double fma_t<double>(double a, double b, double c) { return fma(a, b, c); }
template <> __device__ __forceinline__ float fma_t<float>(float a, float b, float c) { return fmaf(a, b, c); }

template <typename T>
struct Pencil {
    const DevPlan<T>& P;
    const T* __restrict__ u;
    int axis;
    long long base;      // input offset of the pencil's row "global row 0" position minus nothing: see at()
    long long stride;
    int n;               // global length
    int lo_shift;        // input index of global row g is (g - row0 + in_off)
    long long face;      // per-face BC table row
    __device__ __forceinline__ T row(int g) const { return __ldg(u + base + (long long)(g + lo_shift) * stride); }
    __device__ T ghost_lo() const {
        const DevBC<T>& B = P.bc[axis];
        if (B.kind == DEO_BC_PERIODIC) return row(B.per_lo);
        const T* a = B.a_l + (B.per_face ? face * B.K_l : 0);
        T acc = T(0);
        for (int k = 0; k < B.K_l; ++k) acc = fma_t(__ldg(a + k), row(k), acc);
        return acc + __ldg(B.b_l + (B.per_face ? face : 0));
    }
    __device__ T ghost_hi() const {
        const DevBC<T>& B = P.bc[axis];
        if (B.kind == DEO_BC_PERIODIC) return row(B.per_hi);
        const T* a = B.a_r + (B.per_face ? face * B.K_r : 0);
        T acc = T(0);
        for (int k = 0; k < B.K_r; ++k) acc = fma_t(__ldg(a + k), row(n - B.K_r + k), acc);
        return acc + __ldg(B.b_r + (B.per_face ? face : 0));
    }
    // padded pencil entry q[j], j in [0, n+1]
    __device__ __forceinline__ T q(int j) const {
        if (!P.padded[axis]) {
            if (j == 0) return ghost_lo();
            if (j == n + 1) return ghost_hi();
        }
        return row(j - 1);
    }
};

template <typename T>
__device__ __forceinline__ T apply_op(const DevPlan<T>& P, const DevOp<T>& op, const T* __restrict__ u,
                                      int c0, int c1, int c2) {
    const int axis = op.axis;
    const int cl = axis == 0 ? c0 : (axis == 1 ? c1 : c2);   // local output row
    const int r = cl + P.row0[axis];                          // global row
    Pencil<T> pen{P, u, axis, 0, P.in_stride[axis], op.n, P.in_off[axis] - P.row0[axis], 0};
    // offset of the other two coordinates
    long long base = 0;
    if (axis != 0) base += (long long)(c0 + P.in_off[0]) * P.in_stride[0];
    if (axis != 1) base += (long long)(c1 + P.in_off[1]) * P.in_stride[1];
    if (axis != 2) base += (long long)(c2 + P.in_off[2]) * P.in_stride[2];
    pen.base = base;
    if (P.bc[axis].per_face) {
        const int g0 = c0 + P.row0[0], g1 = c1 + P.row0[1], g2 = c2 + P.row0[2];
        pen.face = axis == 0 ? (long long)g1 + (long long)P.n_glob[1] * g2
                 : axis == 1 ? (long long)g0 + (long long)P.n_glob[0] * g2
                             : (long long)g0 + (long long)P.n_glob[0] * g1;
    }
    const int n = op.n;
    if (r < op.nlow || r >= n - op.nhigh) {
        const BRow<T>& b = op.brows[r < op.nlow ? r : op.nlow + (r - (n - op.nhigh))];
        const int start = b.start, nt = b.ntaps;
        if (sizeof(T) == 4 && b.acc64) {
            double acc = 0.0;   // `xtempi = 0.0`: Float32 products summed in Float64 (convolutions.jl:154,:180)
            for (int k = 0; k < nt; ++k) acc += (double)__fmul_rn((float)b.w[k], (float)pen.q(start + k));
            return (T)acc;
        }
        T acc = T(0);
        for (int k = 0; k < nt; ++k) acc = fma_t(b.w[k], pen.q(start + k), acc);
        return acc;
    }
    // interior pattern
    T acc = T(0);
    const int nt = op.ntaps;
    if (op.mode == MODE_CONST) {
        const int start = r + 1 + op.soff[0];
        if (start >= 1 && start + nt - 1 <= n) {   // no ghost touched
            const T* p = u + base + (long long)(start - 1 + pen.lo_shift) * pen.stride;
#pragma unroll 1
            for (int k = 0; k < nt; ++k) acc = fma_t(op.w[0][k], __ldg(p + (long long)k * pen.stride), acc);
        } else {
            for (int k = 0; k < nt; ++k) acc = fma_t(op.w[0][k], pen.q(start + k), acc);
        }
    } else if (op.mode == MODE_SIGNSEL) {
        const T c = __ldg(op.coeff + (r - op.cshift));
        const int set = c >= T(0) ? 0 : 1;
        const int start = r + 1 + op.soff[set];
        for (int k = 0; k < nt; ++k) acc = fma_t(c * op.w[set][k], pen.q(start + k), acc);
    } else {
        const T* w = op.table + (long long)r * nt;
        const int start = r + 1 + __ldg(op.table_soff + r);
        for (int k = 0; k < nt; ++k) acc = fma_t(__ldg(w + k), pen.q(start + k), acc);
    }
    return acc;
}

}  // namespace deo
