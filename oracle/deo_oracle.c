/*
 * TEST INFRASTRUCTURE -- NOT PRODUCT CODE.  See deo_oracle_impl.inc.
 *
 * CPU oracle for the DiffEqOperators.jl operator-application hot path,
 * instantiated for Float64 (_f64) and Float32 (_f32).
 *
 * Parity status: PINNED against the reference's own golden vectors
 * (tests/golden/reference_kats.json, transcribed from the reference test
 * files cited there) by tests/test_oracle_golden.py.  The Julia reference
 * itself cannot be executed in this environment (no julia binary), so there
 * is no oracle/_ref build.  Beyond the golden vectors the oracle passes the
 * reference's analytic tests transcribed in the same file (regular / generic
 * operator validation, operations on matrices, heat equation with Dirichlet /
 * Neumann BCs, the KdV single-soliton test: negative upwind coefficient, offside
 * points, a five-coefficient GeneralBC).  Not pinned by any reference test:
 * coefficient VECTORS of mixed sign through the upwind mul!, Float32 results,
 * non-separable N-D fields.
 *
 * Build: make -C oracle      (gcc -O2 -ffp-contract=off -fopenmp)
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define FN(name) name##_f64
#define REAL double
#include "deo_oracle_impl.inc"
#undef FN
#undef REAL

#define FN(name) name##_f32
#define REAL float
#define REAL_IS_FLOAT 1
#include "deo_oracle_impl.inc"
#undef FN
#undef REAL
#undef REAL_IS_FLOAT

int deo_oracle_abi_version(void) { return 1; }
