// Host-side tiling decisions of the persistent tiled kernel: tile-origin shifts and the chunk length of the march axis.
// Plain C++ without any CUDA dependency, so that tests/tiling_driver.cpp can compile and sweep it on a machine without a GPU.
#pragma once

namespace deo {
namespace tiling {

// What a face tile must hold: the rows next to a face are evaluated from ONE tile's shared-memory plane, so the first / last
// tile along x (y) needs the boundary stencil's 2R+1 points and the BC's stencil (K points) inside the tile plus its halo, and
// must own the R rows next to the face.  w = columns (rows) of the tile inside the array, halo = HX (R).
inline bool face_tile_ok(long long w, int halo, int R, int K) {
    return w + halo >= 2 * R + 1 && w + halo >= K && w >= R;
}

// Tile origins along one axis: tile * T - shift.  Shift 0 wherever it works.  An extent just above a multiple of the tile
// (the 2^k + 1 grids) leaves a last tile too narrow for its face: a shift widens it at the price of the first one.  `res`,
// `step`: the shift must equal res modulo step (x axis: step = vector length, res = 1 for an input padded along the contiguous
// axis, whose TMA boxes must start on 16-byte boundaries; y axis: any shift).  Among the admissible shifts: fewest tiles,
// then the smallest shift.  Returns -1 when no shift fits (the plan then runs on the per-point kernel).
inline long long pick_shift(long long n, long long T, int halo, int R, int Kl, int Kr, int res, int step, bool may_shift) {
    long long best = -1, best_tiles = 0;
    for (long long sh = res; sh < T; sh += step) {
        if (sh > 0 && !may_shift) break;
        const long long tiles = (n + sh + T - 1) / T, wlast = (n - 1 + sh) % T + 1, wfirst = tiles > 1 ? T - sh : n;
        const bool ok = n >= 2 * R + 2 && face_tile_ok(wlast, halo, R, Kr) &&
                        (tiles == 1 ? face_tile_ok(n, halo, R, Kl) : face_tile_ok(wfirst, halo, R, Kl));
        if (ok && (best < 0 || tiles < best_tiles)) { best = sh; best_tiles = tiles; }
    }
    return best;
}

// Chunk length of the march axis for `len` planes: among the chunk lengths <= zmax the one with the shortest makespan in
// plane-steps (one CTA per SM works through ceil(items / slots) items, each costing its planes plus 2R priming planes).
// Every chunk length from the bound down to 4R+4 is a candidate (below half the bound only while nothing fits); a chunk list
// is admissible when its last chunk keeps at least R+1 planes (a face's one-sided rows are never cut) or there is a single
// chunk.  `cap` > 0 bounds the chunk from above for good (TABLE variants stage `cap` rows of weights per item): when nothing
// fits at or below zmax, the shortest admissible chunk above it is taken.
inline long long pick_chunk(long long len, long long zmax, int R, long long tiles, long long slots, long long cap) {
    if (zmax <= 0) zmax = len;
    if (cap > 0 && zmax > cap) zmax = cap;
    if (zmax < 4 * R + 4) zmax = 4 * R + 4;
    long long zc = len;
    double best = 1e30;
    bool found = false;
    const long long clo = 4 * R + 4;
    for (long long c = zmax < len ? zmax : len; c >= clo || c == len; --c) {
        const long long nchunks = (len + c - 1) / c;
        const long long last = len - (nchunks - 1) * c;
        if (!(nchunks > 1 && last < R + 1)) {
            if (found && c * 2 < zmax) break;                      // do not go below half the bound
            const long long rounds = (tiles * nchunks + slots - 1) / slots;
            const double cost = (double)rounds * (double)(c + 2 * R);
            if (cost < best - 1e-12) { best = cost; zc = c; found = true; }
        }
        if (c <= clo) break;
    }
    if (!found) {
        const long long top = cap > 0 ? cap : len;
        for (long long c = zmax + 1; c <= top && c < len; ++c) {
            const long long nchunks = (len + c - 1) / c;
            if (len - (nchunks - 1) * c >= R + 1) { zc = c; break; }
        }
    }
    return zc;
}

}  // namespace tiling
}  // namespace deo
