/*
 * numerov_quad.c -- INDEPENDENT HIGH-PRECISION REFERENCE.  TEST INFRASTRUCTURE ONLY.
 *
 * Solves the DISCRETE Numerov eigenproblem in IEEE binary128 (__float128, 113-bit significand,
 * libquadmath) with the TEXTBOOK recurrence -- a division per step, no rescaled variables --
 *
 *     u_{k+1} = (1/fp_k - 10) u_k - u_{k-1},   fp_k = F_k + e/12,   u = (f/12) psi,
 *     u_{-1} = 0, u_0 = 1,      eigenvalue  <=>  u_n(e) = 0,
 *
 * i.e. f_{k+1} psi_{k+1} = (12 - 10 f_k) psi_k - f_{k-1} psi_{k-1} with psi = 0 one grid point left
 * of the window and at its right end.  It shares with the oracle and the
 * kernels only the coefficient table F_k (the DATA of the discrete problem, taken as exact
 * doubles); the recurrence form, the precision, and the root finder are different.  It therefore
 * pins what the 60-digit replay of the X form (tests/golden/make_golden.py) cannot: how far the
 * FP64 eigenvalues of the product recurrences are from the eigenvalues of the discrete problem
 * they discretise -- the rounding-noise floor per grid size.
 *
 * Build: gcc -O2 -shared -fPIC numerov_quad.c -lquadmath [-fopenmp]   (oracle/Makefile)
 */
#include <math.h>
#include <quadmath.h>
#include <stdint.h>

typedef __float128 q128;

/* The discrete problem is defined by its coefficient table, taken as exact doubles.  Two kinds,
 * one per product recurrence:
 *   kind 0 (X form):  F_k = (1 - q_k)/12,   f_k/12 = F_k + e/12,   e = s E;
 *   kind 1 (D form):  A_k = 12 q_k,         f_k    = 1 - (A_k - 12 e)/12.
 * `en` is e/12 (kind 0) or 12 e (kind 1), in binary128. */
static int g_kind = 0;
void quad_set_table_kind(int kind) { g_kind = kind; }

/* Sign changes of u_k, k = 0..n, and the sign of u_n (returned through *tail_sign: -1, 0, +1). */
static uint32_t quad_nodes(const double* F, uint32_t n, q128 en, int* tail_sign) {
    q128      um = 0, u = 1;
    uint32_t  nodes = 0;
    const int kind  = g_kind;
    for (uint32_t k = 0; k < n; k++) {
        const q128 fp = kind == 0 ? (q128)F[k] + en : (1 - ((q128)F[k] - en) / 12) / 12;
        const q128 un = (1 / fp - 10) * u - um;
        if ((un < 0) != (u < 0)) nodes++;
        um = u;
        u  = un;
        if ((k & 127u) == 127u) { /* exact power-of-two rescale, keeps |u| far from the exponent limits */
            int ex;
            (void)frexpq(u, &ex);
            u  = scalbnq(u, -ex);
            um = scalbnq(um, -ex);
        }
    }
    if (tail_sign) *tail_sign = (u > 0) - (u < 0);
    return nodes;
}

/* Node count of the discrete problem at energy E (cm^-1): the number of eigenvalues below E. */
uint32_t quad_node_count(const double* F, uint32_t n, double s, double E) {
    return quad_nodes(F, n, g_kind == 0 ? ((q128)s * (q128)E) / 12 : 12 * ((q128)s * (q128)E), 0);
}

/* Eigenvalue number v of the discrete problem, bracketed by [E_lo, E_hi] (must satisfy
 * nodes(E_lo) <= v < nodes(E_hi)), bisected on the node count in binary128 until the bracket is
 * narrower than 2^-100 of its magnitude.  Returned rounded to double (hi) plus the remainder (lo):
 * E = hi + lo to ~1e-30.  Returns NaN when the bracket does not hold the level. */
double quad_level(const double* F, uint32_t n, double s, uint32_t v, double E_lo, double E_hi, double* lo_part) {
    const q128 c = g_kind == 0 ? (q128)s / 12 : 12 * (q128)s;
    q128       a = (q128)E_lo, b = (q128)E_hi;
    if (!(quad_nodes(F, n, c * a, 0) <= v && quad_nodes(F, n, c * b, 0) > v)) return NAN;
    for (int it = 0; it < 140; it++) {
        const q128 m = (a + b) / 2;
        if (quad_nodes(F, n, c * m, 0) > v) b = m;
        else a = m;
        if (b - a <= scalbnq(fabsq(b), -100)) break;
    }
    const q128   m  = (a + b) / 2;
    const double hi = (double)m;
    if (lo_part) *lo_part = (double)(m - (q128)hi);
    return hi;
}

/* All levels v_min..v_max at once (OpenMP over levels): brackets[2*l], brackets[2*l+1] per level. */
void quad_levels(const double* F, uint32_t n, double s, uint32_t v_min, uint32_t v_max, const double* brackets,
                 double* out_hi, double* out_lo) {
#pragma omp parallel for schedule(dynamic, 1)
    for (int64_t l = 0; l <= (int64_t)(v_max - v_min); l++)
        out_hi[l] = quad_level(F, n, s, v_min + (uint32_t)l, brackets[2 * l], brackets[2 * l + 1], out_lo ? out_lo + l : 0);
}
