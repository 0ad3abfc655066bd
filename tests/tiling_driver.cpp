// CPU check of the tiled kernel's host-side tiling decisions (csrc/tiling_host.hpp), swept over extents, radii and vector lengths.
// Compiled and run by tests/test_tiling_cpu.py; prints "OK <cases>" or the first violated invariant.
#include <cstdio>
#include <cstdlib>
#include <initializer_list>
#include "../diffeqoperators.jl_b200/csrc/tiling_host.hpp"

using namespace deo::tiling;

static long long fails = 0;
#define CHECK(cond, ...) do { if (!(cond)) { if (fails < 10) { std::printf("FAIL %s: ", #cond); std::printf(__VA_ARGS__); std::printf("\n"); } ++fails; } } while (0)

int main() {
    long long cases = 0;
    // ---- tile-origin shifts ------------------------------------------------------------------------------------
    for (int VEC : {2, 4})
        for (int R = 1; R <= 4; ++R)
            for (int axis = 0; axis < 3; ++axis) {                 // 0: x of a 3-D tile, 1: x of a 2-D strip, 2: y
                const long long T = axis == 0 ? 32 * VEC : axis == 1 ? 32 * VEC * 32 : 32;
                const int halo = axis == 2 ? R : ((R + VEC - 1) / VEC) * VEC;
                const int step = axis == 2 ? 1 : VEC;
                for (int res : {0, 1}) {
                    if (axis == 2 && res) continue;
                    for (int K : {0, 2, 3})
                        for (long long n = 4 * R + 4; n <= (axis == 1 ? 3 * T + 40 : 6 * T + 40); n += (axis == 1 && n > 200 && (n % T) > 40 && (n % T) < T - 40) ? 37 : 1) {
                            ++cases;
                            const long long sh = pick_shift(n, T, halo, R, K, K, res, step, true);
                            // brute force: is there any admissible shift at all, and what is the fewest number of tiles?
                            long long best_tiles = -1, first = -1;
                            for (long long s = res; s < T; s += step) {
                                const long long tiles = (n + s + T - 1) / T, wlast = (n - 1 + s) % T + 1, wfirst = tiles > 1 ? T - s : n;
                                const bool ok = face_tile_ok(wlast, halo, R, K) && face_tile_ok(tiles == 1 ? n : wfirst, halo, R, K);
                                if (ok && (best_tiles < 0 || tiles < best_tiles)) { best_tiles = tiles; first = s; }
                            }
                            CHECK((sh < 0) == (first < 0), "n=%lld T=%lld R=%d res=%d K=%d: shift %lld, brute force %lld", n, T, R, res, K, sh, first);
                            if (sh < 0) continue;
                            CHECK(sh == first, "n=%lld T=%lld R=%d res=%d: shift %lld is not the smallest with the fewest tiles (%lld)", n, T, R, res, sh, first);
                            CHECK(sh % step == res % step && sh < T, "n=%lld: shift %lld breaks the residue / range", n, sh);
                            const long long tiles = (n + sh + T - 1) / T, wlast = (n - 1 + sh) % T + 1, wfirst = tiles > 1 ? T - sh : n;
                            CHECK(wlast >= R && wlast + halo >= 2 * R + 1 && wlast + halo >= K, "n=%lld sh=%lld: last tile %lld too narrow", n, sh, wlast);
                            CHECK(wfirst >= R && wfirst + halo >= 2 * R + 1 && wfirst + halo >= K, "n=%lld sh=%lld: first tile %lld too narrow", n, sh, wfirst);
                            CHECK((tiles - 1) * T - sh <= n - 1 && tiles * T - sh >= n, "n=%lld sh=%lld: tiles do not cover the array", n, sh);
                            if (res == 0 && face_tile_ok((n - 1) % T + 1, halo, R, K) && face_tile_ok(n > T ? T : n, halo, R, K))
                                CHECK(sh == 0, "n=%lld T=%lld R=%d: shift %lld although the unshifted tiling fits", n, T, R, sh);
                            // without permission to shift only the residue itself is considered
                            const long long sh0 = pick_shift(n, T, halo, R, K, K, res, step, false);
                            CHECK(sh0 == -1 || sh0 == res, "n=%lld: shift %lld without permission", n, sh0);
                        }
                }
            }
    // ---- march-axis chunks -------------------------------------------------------------------------------------
    for (int R = 1; R <= 4; ++R)
        for (long long zmax : {0LL, 16LL, 24LL, 32LL, 64LL})
            for (long long cap : {0LL, 64LL})
                for (long long tiles : {1LL, 7LL, 64LL, 256LL, 1000LL})
                    for (long long len = 1; len <= 1300; ++len) {
                        ++cases;
                        const long long zc = pick_chunk(len, zmax, R, tiles, 148, cap);
                        CHECK(zc >= 1 && zc <= len, "len=%lld zmax=%lld R=%d: chunk %lld out of range", len, zmax, R, zc);
                        const long long nchunks = (len + zc - 1) / zc, last = len - (nchunks - 1) * zc;
                        if (nchunks > 1) {
                            CHECK(last >= R + 1, "len=%lld zmax=%lld R=%d: last chunk of %lld planes cuts a face's rows (chunk %lld)", len, zmax, R, last, zc);
                            CHECK(zc >= 4 * R + 4, "len=%lld zmax=%lld R=%d: chunk %lld shorter than 4R+4", len, zmax, R, zc);
                        }
                        if (cap > 0 && len >= 4 * R + 4)
                            CHECK(zc <= cap, "len=%lld zmax=%lld R=%d: chunk %lld above the cap %lld", len, zmax, R, zc, cap);
                        const long long bound = zmax <= 0 ? len : (zmax < 4 * R + 4 ? 4 * R + 4 : zmax);
                        if (len <= bound) CHECK(nchunks == 1 || zc * 2 >= (cap > 0 && bound > cap ? cap : bound), "len=%lld: needlessly short chunk %lld", len, zc);
                    }
    if (fails) { std::printf("%lld violations in %lld cases\n", fails, cases); return 1; }
    std::printf("OK %lld\n", cases);
    return 0;
}
