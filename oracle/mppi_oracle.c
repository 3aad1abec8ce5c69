/*
 * TEST INFRASTRUCTURE — NOT PRODUCT CODE.
 *
 * mppi_oracle.c: CPU restatement (plain C) of the reference's MPPI update path.
 * See mppi_oracle_impl.h for the per-function reference citations.  Built by
 * oracle/Makefile into oracle/_build/libmppi_oracle.so and loaded with ctypes by
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg
 * ONLY.  The product library (mppi_tf_b200/csrc) never links or calls it.
 *
 * Pinning: every known-answer vector the reference's own tests hold for this path
 * (test/test_model.cpp, test/test_cost.cpp, test/test_controller.cpp,
 * test/test_utile.cpp, scripts/test.py point-mass/static-cost/controller cases) is
 * replayed against these functions by tests/test_oracle_kats.py.  The composed
 * next() has no golden vector in the reference (testAll is empty,
 * test/test_controller.cpp:224-226): it is pinned by composition of the unit KATs
 * and by the fixtures generated from the reference's Python twin
 * (tests/golden/gen_python_twin_fixtures.py).  The MLP step is PARITY UNPINNED.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_MAX_S 32
#define ORC_MAX_A 16
#define ORC_MAX_H 512

#define REAL float
#define SUFFIX _f32
#define REAL_EXP(x) expf(x)
#include "mppi_oracle_impl.h"
#undef REAL
#undef SUFFIX
#undef REAL_EXP

#define REAL double
#define SUFFIX _f64
#define REAL_EXP(x) exp(x)
#include "mppi_oracle_impl.h"
#undef REAL
#undef SUFFIX
#undef REAL_EXP

int orc_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* ------------------------------------------------------------------------------------------
 * Noise stream specification (new design: the reference draws from TF's stateful
 * RandomNormal, src/controller_base.cpp:196-199, whose values cannot be restated; parity runs
 * inject eps).  The stream below is the bit-exact integer contract for the CUDA generator:
 *
 *   Philox4x32-10 (Salmon et al., SC'11; Random123 v1.x constants)
 *     key     = (seed_lo, seed_hi)
 *     counter = (call, sample, update, stream)
 *   where `sample` is the GLOBAL sample index (so results do not depend on the rank count),
 *   `call` c covers the standard normals z[4c..4c+3] of that sample's flattened [T][a] row,
 *   `update` is the controller's update counter and `stream` the controller index of a batch.
 *   Box-Muller on (x0,x1) -> z[4c], z[4c+1] and (x2,x3) -> z[4c+2], z[4c+3]:
 *     f(x) = as_float((x >> 9) | 0x3f800000) in [1,2)
 *     r = sqrt(-2 ln(2 - f(xa))),  th = 2 pi (f(xb) - 1),  z_even = r cos th, z_odd = r sin th
 * ------------------------------------------------------------------------------------------ */
#define PHILOX_M0 0xD2511F53u
#define PHILOX_M1 0xCD9E8D57u
#define PHILOX_W0 0x9E3779B9u
#define PHILOX_W1 0xBB67AE85u

/* Philox4x32-R: R = 10 is Random123's default (and TensorFlow's); R = 7 is the shortest variant Salmon et al. report as
 * passing BigCrush (mppi_config.philox_rounds). */
void orc_philox4x32_r(const uint32_t ctr[4], const uint32_t key[2], int rounds, uint32_t out[4])
{
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
    for (int r = 0; r < rounds; r++) {
        uint64_t p0 = (uint64_t)PHILOX_M0 * c0, p1 = (uint64_t)PHILOX_M1 * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += PHILOX_W0; k1 += PHILOX_W1;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) { orc_philox4x32_r(ctr, key, 10, out); }

static float u01_from_bits(uint32_t x)
{
    uint32_t b = (x >> 9) | 0x3f800000u;
    float f;
    memcpy(&f, &b, 4);
    return f; /* in [1,2) */
}

/* Standard normals z[n_per_sample] for samples [k0,k1) (row-major [k1-k0][n_per_sample]). */
void orc_philox_normals_r(uint64_t seed, uint32_t update, uint32_t stream, uint32_t k0, uint32_t k1,
                          int n_per_sample, int rounds, float *z)
{
    uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
    int ncall = (n_per_sample + 3) / 4;
    for (uint32_t k = k0; k < k1; k++)
        for (int c = 0; c < ncall; c++) {
            uint32_t ctr[4] = {(uint32_t)c, k, update, stream}, x[4];
            float n[4];
            orc_philox4x32_r(ctr, key, rounds, x);
            for (int h = 0; h < 2; h++) {
                double u1 = 2.0 - (double)u01_from_bits(x[2 * h]);
                double th = 6.283185307179586 * ((double)u01_from_bits(x[2 * h + 1]) - 1.0);
                double r = sqrt(-2.0 * log(u1));
                n[2 * h] = (float)(r * cos(th));
                n[2 * h + 1] = (float)(r * sin(th));
            }
            for (int j = 0; j < 4; j++)
                if (4 * c + j < n_per_sample) z[(size_t)(k - k0) * n_per_sample + 4 * c + j] = n[j];
        }
}
void orc_philox_normals(uint64_t seed, uint32_t update, uint32_t stream, uint32_t k0, uint32_t k1,
                        int n_per_sample, float *z)
{
    orc_philox_normals_r(seed, update, stream, k0, k1, n_per_sample, 10, z);
}
