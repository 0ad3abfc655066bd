// Host-side plan construction: turns the reference's operand bundles (deo_op_desc, field for field
// a DerivativeOperator) into per-row tap tables, following the row/branch logic of the reference's
// 1-D convolution methods (src/derivative_operators/convolutions.jl).  Everything that branches in
// the reference (boundary rows, upwind sign selection at the edges, coefficient-index quirks) is
// decided here once; the kernels only see "row r = sum_k w[k] * q[start + k]".
//
// Padded-pencil convention: q[0] = low ghost, q[j+1] = u[j], q[n+1] = high ghost (0-based; the
// reference's x[j] is q[j-1]).
#include <map>

#include "common.hpp"

namespace deo {
namespace {

template <typename T>
struct RowSpec {
    int start = 0, ntaps = 0;
    int acc64 = 0;
    T w[kMaxBTaps];   // already multiplied by the row coefficient
};

template <typename T>
struct OpView {
    const deo_op_desc& d;
    const T *stencil, *low, *high, *coeff;
    int n, sl, bsl, bpc, off;
    explicit OpView(const HostOp& h)
        : d(h.d), stencil((const T*)h.stencil.data()), low((const T*)h.low.data()),
          high((const T*)h.high.data()), coeff((const T*)h.coeff.data()),
          n(h.d.len), sl(h.d.stencil_length), bsl(h.d.boundary_stencil_length),
          bpc(h.d.boundary_point_count), off(h.d.offside) {}
    int n_interior() const { return n - 2 * bpc; }
    int n_high() const { return d.kind == DEO_OP_UPWIND ? bpc + off : bpc; }
};

// (-1)^d * reverse(w)  (convolutions.jl:134-136, :184, :193): exact sign flip, no rounding.
template <typename T>
void mirrored(const T* w, int len, int d, T* out) {
    for (int k = 0; k < len; ++k) out[k] = (d % 2 == 0) ? w[len - 1 - k] : -w[len - 1 - k];
}

template <typename T>
struct RowSink {
    std::map<int, RowSpec<T>> rows;   // 1-based row -> spec; later writers replace earlier ones,
                                      // like the reference's left / interior / right sequence
    int n;
    bool ok = true;
    std::string why;
    void emit(int row, int start, int ntaps, const T* w, T c, bool acc64) {
        if (!ok) return;
        if (row < 1 || row > n) { ok = false; why = "an output row falls outside the grid"; return; }
        if (start < 0 || start + ntaps > n + 2 || ntaps > kMaxBTaps) {
            ok = false; why = "a stencil window falls outside the padded pencil"; return;
        }
        RowSpec<T> r;
        r.start = start; r.ntaps = ntaps; r.acc64 = acc64 ? 1 : 0;
        for (int k = 0; k < ntaps; ++k) r.w[k] = c * w[k];   // (cur_coeff * cur_stencil[idx])
        for (int k = ntaps; k < kMaxBTaps; ++k) r.w[k] = T(0);
        rows[row] = r;
    }
};

// ---- centered: convolutions.jl:27-118 (plain), :367-469 (BoundaryPaddedVector) -------------------
template <typename T>
void centered_left(const OpView<T>& A, RowSink<T>& S) {
    for (int i = 1; i <= A.bpc; ++i)                                   // :86-94, :412-420
        S.emit(i, 0, A.bsl, A.low + (size_t)(i - 1) * A.bsl, A.coeff[i - 1], false);
}
template <typename T>
void centered_interior_row(const OpView<T>& A, RowSink<T>& S, int i, bool bpv) {
    const int R = A.sl / 2;
    const T* w = A.d.nonuniform ? A.stencil + (size_t)(i - A.bpc - 1) * A.sl : A.stencil;   // :44, :383
    const T c = bpv ? A.coeff[i - A.bpc - 1] : A.coeff[i - 1];         // :384,:393,:428,:454 vs :45,:54
    S.emit(i, i - R, A.sl, w, c, false);                               // x[i-mid+idx] -> q[i-R+k]
}
template <typename T>
void centered_right(const OpView<T>& A, RowSink<T>& S, bool bpv) {
    for (int i = 1; i <= A.bpc; ++i) {                                 // :107-117, :460-468
        const int row = A.n - A.bpc + i;
        const T c = bpv ? A.coeff[row - 1] : A.coeff[i - 1];           // :462 vs :109 (coeff[i], i in 1:bpc)
        S.emit(row, A.n + 2 - A.bsl, A.bsl, A.high + (size_t)(i - 1) * A.bsl, c, false);
    }
}

// ---- uniform upwind: convolutions.jl:123-215 ------------------------------------------------------
template <typename T>
void upwind_u_left(const OpView<T>& A, RowSink<T>& S) {
    const int xlen = A.n + 2;
    for (int i = 1; i <= A.bpc; ++i) {                                 // :152-167
        const T c = A.coeff[i - 1];
        if (c >= 0 && i + A.sl <= xlen && i >= A.off)
            S.emit(i, i - A.off, A.sl, A.stencil, c, true);            // x[i+idx-off]
        else
            S.emit(i, 0, A.bsl, A.low + (size_t)(i - 1) * A.bsl, c, true);
    }
}
template <typename T>
void upwind_u_interior_row(const OpView<T>& A, RowSink<T>& S, int i) {
    const T c = A.coeff[i - 1];                                        // :133
    if (c >= 0) S.emit(i, i - A.off, A.sl, A.stencil, c, false);       // :139 x[i+idx-off]
    else {
        T rev[kMaxTaps];
        mirrored(A.stencil, A.sl, A.d.derivative_order, rev);          // :134-136
        S.emit(i, i - A.sl + 1 + A.off, A.sl, rev, c, false);          // :138 x[i-sl+1+idx+off]
    }
}
template <typename T>
void upwind_u_right(const OpView<T>& A, RowSink<T>& S) {
    const int xlen = A.n + 2, d = A.d.derivative_order;
    T rev[kMaxBTaps];
    for (int i = 1; i <= A.bpc + A.off; ++i) {                         // :178-214
        const int row = A.n - A.bpc + i - A.off;
        if (row < 1 || row > A.n) { S.ok = false; S.why = "an output row falls outside the grid"; return; }
        const T c = A.coeff[row - 1];
        const bool fit = xlen - A.sl - A.bpc + i >= 1;
        if (c < 0 && fit && i <= A.bpc + 1) {                          // :181-188
            mirrored(A.stencil, A.sl, d, rev);
            S.emit(row, xlen - A.sl - A.bpc + i - 1, A.sl, rev, c, true);   // x[xlen-sl+idx-bpc+i-1]
        } else if (c < 0 && fit && i > A.bpc + 1) {                    // :189-197
            mirrored(A.high + (size_t)(A.bpc + A.off + 1 - i - 1) * A.bsl, A.bsl, d, rev);
            S.emit(row, xlen - A.bsl, A.bsl, rev, c, true);            // x[xlen-bsl+idx]
        } else if (c >= 0 && i < A.off + 1) {                          // :198-203
            S.emit(row, xlen - A.sl + i - A.off, A.sl, A.stencil, c, true);   // x[xlen-sl+i+idx-off]
        } else {                                                       // :204-209
            S.emit(row, xlen - A.bsl, A.bsl, A.high + (size_t)(i - 1) * A.bsl, c, true);
        }
    }
}

// ---- non-uniform upwind: convolutions.jl:221-362 ---------------------------------------------------
template <typename T>
void upwind_n_left(const OpView<T>& A, RowSink<T>& S) {
    const T* LB1 = A.low;
    const T* LB2 = A.low + (size_t)A.bpc * A.bsl;
    for (int i = 1; i <= A.bpc; ++i) {                                 // :285-316
        const T c = A.coeff[i - 1];
        if (c >= 0 && A.off == 0)      S.emit(i, i, A.sl, LB1 + (size_t)(i - 1) * A.bsl, c, false);          // x[i+idx]
        else if (c >= 0 && i < A.off)  S.emit(i, 0, A.sl, LB1 + (size_t)(i - 1) * A.bsl, c, false);          // x[idx]
        else if (c >= 0)               S.emit(i, i - A.off, A.sl, LB1 + (size_t)(i - 1) * A.bsl, c, false);  // x[i+idx-off]
        else                           S.emit(i, 0, A.bsl, LB2 + (size_t)(i - 1) * A.bsl, c, false);
    }
}
template <typename T>
void upwind_n_interior_row(const OpView<T>& A, RowSink<T>& S, int i) {
    const T* SC1 = A.stencil;
    const T* SC2 = A.stencil + (size_t)A.n_interior() * A.sl;
    const T c = A.coeff[i - 1];
    if (c >= 0) S.emit(i, i - A.off, A.sl, SC1 + (size_t)(i - A.bpc - 1) * A.sl, c, false);            // :259-263
    else        S.emit(i, i - A.sl + 1 + A.off, A.sl, SC2 + (size_t)(i - A.bpc - 1) * A.sl, c, false); // :266-270
}
template <typename T>
void upwind_n_right(const OpView<T>& A, RowSink<T>& S) {
    const int n = A.n, nh = A.bpc + A.off;
    const T* SC1 = A.stencil;
    const T* HB1 = A.high;
    const T* HB2 = A.high + (size_t)nh * A.bsl;
    for (int i = n - A.bpc + 1 - A.off; i <= n; ++i) {                 // :329-361
        if (i < 1) { S.ok = false; S.why = "an output row falls outside the grid"; return; }
        const T c = A.coeff[i - 1];
        const int hb = i - n + A.bpc + A.off - 1;
        if (c < 0) {
            if (i <= n - A.off) S.emit(i, i - A.sl + 1 + A.off, A.sl, HB2 + (size_t)hb * A.bsl, c, false);   // :335-337
            else                S.emit(i, n - A.sl + 2, A.sl, HB2 + (size_t)hb * A.bsl, c, false);           // :340-342
        } else {
            if (i <= n - A.bpc) S.emit(i, i - A.sl + 1 + A.off, A.sl, SC1 + (size_t)(i - A.bpc - 1) * A.sl, c, false);  // :348-351
            else                S.emit(i, n - A.bsl + 2, A.bsl, HB1 + (size_t)hb * A.bsl, c, false);                     // :354-357
        }
    }
}

template <typename T>
void* upload(deo_plan* plan, const void* host, size_t bytes, cudaError_t* err) { return plan_upload(plan, host, bytes, err); }

template <typename T>
int32_t build_op(deo_plan* plan, const HostOp& H, bool bpv, DevOp<T>& D, size_t op_index) {
    OpView<T> A(H);
    const bool upwind = H.d.kind == DEO_OP_UPWIND;
    const bool nonuni = H.d.nonuniform != 0;
    const int n = A.n;
    const int nlow_nom = A.bpc, nhigh_nom = A.n_high();
    const int int_first = A.bpc + 1;                                // 1-based interior range (all variants)
    const int int_last = upwind ? n - A.bpc - A.off : n - A.bpc;
    if (bpv && !upwind)
        DEO_REQUIRE(n >= 2 * A.bpc + 2, "op on axis %d: len %d too small for the BoundaryPaddedVector centered method (needs >= %d)",
                    H.d.axis, n, 2 * A.bpc + 2);

    // Small grids: replay left / interior / right over every row (later writers win, as in the reference).
    const bool all_explicit = n < nlow_nom + nhigh_nom + 1 || n <= 2 * kMaxBTaps;
    RowSink<T> S;
    S.n = n;
    auto left = [&] { if (!upwind) centered_left(A, S); else if (!nonuni) upwind_u_left(A, S); else upwind_n_left(A, S); };
    auto right = [&] { if (!upwind) centered_right(A, S, bpv); else if (!nonuni) upwind_u_right(A, S); else upwind_n_right(A, S); };
    auto interior_row = [&](int i) {
        if (!upwind) centered_interior_row(A, S, i, bpv);
        else if (!nonuni) upwind_u_interior_row(A, S, i);
        else upwind_n_interior_row(A, S, i);
    };
    left();
    if (all_explicit) for (int i = int_first; i <= int_last; ++i) interior_row(i);
    right();
    DEO_REQUIRE(S.ok, "op on axis %d (len %d): %s (the reference would throw a BoundsError)", H.d.axis, n, S.why.c_str());

    D.axis = H.d.axis;
    D.n = n;
    D.cshift = (bpv && !upwind) ? A.bpc : 0;
    D.coeff = nullptr; D.table = nullptr; D.table_soff = nullptr; D.brows = nullptr;
    for (int s = 0; s < 2; ++s) for (int k = 0; k < kMaxTaps; ++k) D.w[s][k] = T(0);

    if (all_explicit) {
        for (int i = 1; i <= n; ++i)
            DEO_REQUIRE(S.rows.count(i), "op on axis %d (len %d): row %d is written by no convolution (grid too small)", H.d.axis, n, i);
        D.nlow = n; D.nhigh = 0;
        D.mode = MODE_CONST; D.ntaps = 0; D.soff[0] = D.soff[1] = 0;
    } else {
        D.nlow = nlow_nom; D.nhigh = nhigh_nom;
        D.ntaps = A.sl;
        const int R = A.sl / 2;
        if (!upwind) { D.soff[0] = D.soff[1] = -R; }
        else { D.soff[0] = -A.off; D.soff[1] = 1 - A.sl + A.off; }
        bool constc = true;
        for (int i = 1; i < n; ++i) constc = constc && (memcmp(&A.coeff[i], &A.coeff[0], sizeof(T)) == 0);
        if (nonuni) {
            // per-row (c*w) table for the interior pattern rows; the kernel indexes it with the row
            D.mode = MODE_TABLE;
            std::vector<T> tab((size_t)n * A.sl, T(0));
            std::vector<int> soff((size_t)n, 0);
            for (int i = int_first; i <= int_last; ++i) {
                RowSink<T> one; one.n = n;
                if (!upwind) centered_interior_row(A, one, i, bpv); else upwind_n_interior_row(A, one, i);
                DEO_REQUIRE(one.ok, "op on axis %d: interior row %d: %s", H.d.axis, i, one.why.c_str());
                const RowSpec<T>& r = one.rows[i];
                for (int k = 0; k < A.sl; ++k) tab[(size_t)(i - 1) * A.sl + k] = r.w[k];
                soff[i - 1] = r.start - i;   // relative to the centre q[r+1] = q[i]
            }
            cudaError_t e;
            D.table = (const T*)upload<T>(plan, tab.data(), tab.size() * sizeof(T), &e);
            if (!D.table) return cuda_fail(e, "upload(table)", __FILE__, __LINE__);
            D.table_soff = (const int*)upload<T>(plan, soff.data(), soff.size() * sizeof(int), &e);
            if (!D.table_soff) return cuda_fail(e, "upload(table_soff)", __FILE__, __LINE__);
        } else if (constc) {
            D.mode = MODE_CONST;
            const T c = A.coeff[0];
            if (!upwind || c >= 0) {
                for (int k = 0; k < A.sl; ++k) D.w[0][k] = c * A.stencil[k];
            } else {
                T rev[kMaxTaps];
                mirrored(A.stencil, A.sl, H.d.derivative_order, rev);
                for (int k = 0; k < A.sl; ++k) D.w[0][k] = c * rev[k];
                D.soff[0] = D.soff[1];
            }
            for (int k = 0; k < A.sl; ++k) D.w[1][k] = D.w[0][k];
            D.soff[1] = D.soff[0];
        } else {
            D.mode = MODE_SIGNSEL;
            for (int k = 0; k < A.sl; ++k) D.w[0][k] = A.stencil[k];
            if (upwind) mirrored(A.stencil, A.sl, H.d.derivative_order, D.w[1]);
            else for (int k = 0; k < A.sl; ++k) D.w[1][k] = A.stencil[k];
            cudaError_t e;
            D.coeff = (const T*)upload<T>(plan, A.coeff, (size_t)n * sizeof(T), &e);
            if (!D.coeff) return cuda_fail(e, "upload(coeff)", __FILE__, __LINE__);
        }
    }

    // explicit rows -> device
    std::vector<BRow<T>> br((size_t)D.nlow + D.nhigh);
    auto put = [&](int slot, int row1) -> bool {
        auto it = S.rows.find(row1);
        if (it == S.rows.end()) return false;
        BRow<T>& b = br[slot];
        b.start = it->second.start; b.ntaps = it->second.ntaps; b.acc64 = it->second.acc64; b.pad_ = 0;
        for (int k = 0; k < kMaxBTaps; ++k) b.w[k] = it->second.w[k];
        return true;
    };
    for (int i = 1; i <= D.nlow; ++i)
        DEO_REQUIRE(put(i - 1, i), "op on axis %d: low boundary row %d has no stencil", H.d.axis, i);
    for (int i = 1; i <= D.nhigh; ++i)
        DEO_REQUIRE(put(D.nlow + i - 1, n - D.nhigh + i), "op on axis %d: high boundary row %d has no stencil", H.d.axis, n - D.nhigh + i);
    if (plan->host_brows.size() <= op_index) plan->host_brows.resize(op_index + 1);
    plan->host_brows[op_index].assign((const unsigned char*)br.data(), (const unsigned char*)br.data() + br.size() * sizeof(BRow<T>));
    cudaError_t e;
    D.brows = (const BRow<T>*)upload<T>(plan, br.data(), br.size() * sizeof(BRow<T>), &e);
    if (!D.brows) return cuda_fail(e, "upload(brows)", __FILE__, __LINE__);
    return DEO_OK;
}

template <typename T>
int32_t build_typed(deo_plan* plan) {
    plan->devplan.assign(sizeof(DevPlan<T>), 0);
    DevPlan<T>& P = *reinterpret_cast<DevPlan<T>*>(plan->devplan.data());
    P.ndims = plan->ndims;
    P.nops = (int)plan->ops.size();
    P.accumulate = plan->accumulate;
    long long is = 1, os = 1;
    for (int a = 0; a < kMaxDims; ++a) {
        const bool live = a < plan->ndims;
        const long long nloc = live ? plan->local_dim(a) : 1;
        P.n_out[a] = (int)nloc;
        P.n_glob[a] = live ? (int)plan->dims[a] : 1;
        P.row0[a] = (live && a == plan->slab_axis) ? (int)plan->slab_start : 0;
        P.padded[a] = live ? plan->padded[a] : 0;
        P.in_off[a] = live ? (plan->padded[a] ? 1 : 0) + (a == plan->slab_axis ? plan->halo : 0) : 0;
        P.in_stride[a] = is;
        P.out_stride[a] = os;
        is *= live ? plan->in_dim(a) : 1;
        os *= nloc;
    }
    for (int a = 0; a < kMaxDims; ++a) {
        DevBC<T>& B = P.bc[a];
        memset(&B, 0, sizeof B);
        if (a >= plan->ndims) continue;
        const HostBC& H = plan->bc[a];
        B.kind = H.d.kind;
        B.per_face = H.d.per_face;
        B.K_l = H.d.K_l; B.K_r = H.d.K_r;
        if (B.kind == DEO_BC_PERIODIC) {
            // 1-D: l = u[end], r = u[1] (bc_operators.jl:192); N-D: lower = u[1,...], upper = u[end,...]
            // (multi_dim_bc_operators.jl:221-228).
            const int n = (int)plan->dims[a];
            if (plan->ndims == 1) { B.per_lo = n - 1; B.per_hi = 0; } else { B.per_lo = 0; B.per_hi = n - 1; }
        }
        if (B.kind == DEO_BC_AFFINE) {
            cudaError_t e;
            B.a_l = (const T*)upload<T>(plan, H.a_l.data(), H.a_l.size(), &e); if (!B.a_l) return cuda_fail(e, "upload(a_l)", __FILE__, __LINE__);
            B.b_l = (const T*)upload<T>(plan, H.b_l.data(), H.b_l.size(), &e); if (!B.b_l) return cuda_fail(e, "upload(b_l)", __FILE__, __LINE__);
            B.a_r = (const T*)upload<T>(plan, H.a_r.data(), H.a_r.size(), &e); if (!B.a_r) return cuda_fail(e, "upload(a_r)", __FILE__, __LINE__);
            B.b_r = (const T*)upload<T>(plan, H.b_r.data(), H.b_r.size(), &e); if (!B.b_r) return cuda_fail(e, "upload(b_r)", __FILE__, __LINE__);
        }
    }
    for (size_t k = 0; k < plan->ops.size(); ++k) {
        const HostOp& H = plan->ops[k];
        // The 1-D L*Q*u path goes through the BoundaryPaddedVector methods (convolutions.jl:367-469);
        // N-D arrays and plain padded vectors go through the AbstractVector methods (:27-118).
        const bool bpv = plan->ndims == 1 && plan->bc[H.d.axis].d.kind != DEO_BC_NONE;
        int32_t rc = build_op<T>(plan, H, bpv, P.ops[k], k);
        if (rc) return rc;
    }
    return DEO_OK;
}

// Every output row of operator H as "row r = sum_k w[k] * q[start + k]" (0-based rows; w = c*w as the reference forms it),
// produced lazily with the same left / interior / right logic as build_op.  Used by the merged-table builders of the
// tiled kernels (kernel_star.cu, kernel_line.cu).
template <typename T>
struct RowGenT final : RowGenerator {
    OpView<T> A;
    bool upwind, nonuni, bpv, good = true;
    int int_first, int_last;
    RowSink<T> edges;     // left and right rows (a handful)
    RowGenT(const HostOp& H, bool bpv_) : A(H), upwind(H.d.kind == DEO_OP_UPWIND), nonuni(H.d.nonuniform != 0), bpv(bpv_) {
        n = A.n;
        int_first = A.bpc + 1;
        int_last = upwind ? n - A.bpc - A.off : n - A.bpc;
        edges.n = n;
        // the three parts must not overlap (tiny grids replay "later writers win" in build_op instead)
        if (int_last < int_first || A.bpc + A.n_high() + 1 > n) { good = false; return; }
        if (!upwind) centered_left(A, edges); else if (!nonuni) upwind_u_left(A, edges); else upwind_n_left(A, edges);
        if (!upwind) centered_right(A, edges, bpv); else if (!nonuni) upwind_u_right(A, edges); else upwind_n_right(A, edges);
        good = edges.ok;
        bool constc = true;
        for (int i = 1; i < n; ++i) constc = constc && (memcmp(&A.coeff[i], &A.coeff[0], sizeof(T)) == 0);
        interior_uniform = !nonuni && constc;
    }
    bool ok() const override { return good; }
    bool row(int r0, HostRow& out) override {
        const int i = r0 + 1;
        const RowSpec<T>* spec = nullptr;
        RowSink<T> one;
        if (i >= int_first && i <= int_last) {
            one.n = n;
            if (!upwind) centered_interior_row(A, one, i, bpv);
            else if (!nonuni) upwind_u_interior_row(A, one, i);
            else upwind_n_interior_row(A, one, i);
            if (!one.ok) return false;
            spec = &one.rows[i];
        } else {
            auto it = edges.rows.find(i);
            if (it == edges.rows.end()) return false;
            spec = &it->second;
        }
        out.start = spec->start;
        out.ntaps = spec->ntaps;
        for (int k = 0; k < kMaxBTaps; ++k) out.w[k] = k < spec->ntaps ? (double)spec->w[k] : 0.0;
        return true;
    }
};

}  // namespace

void* plan_upload(deo_plan* plan, const void* host, size_t bytes, cudaError_t* err) {
    *err = cudaSuccess;
    DeviceBlob* b = nullptr;
    if (plan->blob_cursor < plan->blobs.size() && plan->blobs[plan->blob_cursor]->bytes == bytes) {
        b = plan->blobs[plan->blob_cursor].get();                  // same position, same size as in the previous build
    } else {
        auto blob = std::make_unique<DeviceBlob>();
        *err = cudaMalloc(&blob->p, bytes ? bytes : 1);
        if (*err != cudaSuccess) return nullptr;
        blob->bytes = bytes;
        b = blob.get();
        if (plan->blob_cursor < plan->blobs.size()) plan->blobs[plan->blob_cursor] = std::move(blob);   // frees the old one (cudaFree synchronises)
        else plan->blobs.push_back(std::move(blob));
    }
    ++plan->blob_cursor;
    if (bytes) {
        const size_t padded = (bytes + 15) / 16 * 16;
        plan->stage_want += padded;
        if (plan->stage && plan->stage_used + padded <= plan->stage_cap) {
            unsigned char* src = plan->stage + plan->stage_used;
            memcpy(src, host, bytes);
            plan->stage_used += padded;
            *err = cudaMemcpyAsync(b->p, src, bytes, cudaMemcpyHostToDevice, rt().stream);   // pinned source: truly asynchronous
        } else {
            // no arena yet (first build) or it is too small: pageable source, the runtime stages it itself (and waits for
            // the stream first); the arena grows before the next rebuild
            *err = cudaMemcpyAsync(b->p, host, bytes, cudaMemcpyHostToDevice, rt().stream);
        }
        if (*err != cudaSuccess) return nullptr;
    }
    return b->p;
}

std::unique_ptr<RowGenerator> make_row_generator(const deo_plan* plan, int k) {
    const HostOp& H = plan->ops[(size_t)k];
    const bool bpv = plan->ndims == 1 && plan->bc[H.d.axis].d.kind != DEO_BC_NONE;
    if (plan->dtype == DEO_F64) return std::unique_ptr<RowGenerator>(new RowGenT<double>(H, bpv));
    return std::unique_ptr<RowGenerator>(new RowGenT<float>(H, bpv));
}

int32_t build_device_plan(deo_plan* plan) {
    return plan->dtype == DEO_F64 ? build_typed<double>(plan) : build_typed<float>(plan);
}

}  // namespace deo
