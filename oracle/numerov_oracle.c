/*
 * numerov_oracle.c -- CPU ORACLE.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * `--impl reference` legs may build, load or call this file.  Nothing under
 * epseon_backend_b200/ links it; the product path fails loudly without its
 * CUDA library.
 *
 * PARITY UNPINNED against the reference: the reference repository
 * (UniversityOfGdanskTeamPython/epseon_backend @ f2996b0) contains NO
 * implementation of this path -- VibwaAlgorithm<FP>::run only allocates
 * buffers (cpp/gpu/include/epseon/gpu/algorithms/vibwa.hpp:605-637),
 * MorsePotentialGenerator::get_potential_data returns {}
 * (cpp/gpu/include/epseon/gpu/task_configurator/potential_source.hpp:223-225)
 * and _libepseon_cpu is a greet() stub (cpp/cpu/source/libcpu.cpp:3-36).
 * There is no reference arithmetic, golden vector or known-answer test to
 * follow.  This oracle is therefore the BUILD'S OWN plain-C statement of the
 * numerical specification in DESIGN.md section 3; it is pinned instead by
 * independent known answers (tests/test_oracle_pins.py): the analytic Morse
 * spectrum, the harmonic oscillator, a 50-digit mpmath replay of the same
 * recurrence (tests/golden/), and scipy's tridiagonal eigen-solver.
 *
 * What the reference DOES fix, and what is followed here:
 *   - the parameter set: MorsePotentialConfig{dissociation_energy,
 *     equilibrium_bond_distance, well_width, min_r, max_r, point_count}
 *     (potential_source.hpp:103-183) and VibwaAlgorithmConfig{mass_atom_0,
 *     mass_atom_1, integration_step, min_distance_to_asymptote, min_level,
 *     max_level} (algorithm_config.hpp:76-199);
 *   - the output shape: level_count = max_level - min_level + 1 energies per
 *     potential curve (algorithm_config.hpp:177,183-190).
 *
 * Build (see oracle/Makefile):  gcc -O3 -march=native -ffp-contract=off
 * [-fopenmp].  Every floating-point operation below is a single IEEE-754
 * binary64 operation (+, *, or an explicit fma()); -ffp-contract=off keeps
 * the compiler from fusing anything else, so results are bit-identical to
 * the CUDA kernels, which use __dadd_rn/__dmul_rn/__fma_rn in the same order.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_HBAR2_OVER_2 16.857629206 /* amu * Angstrom^2 * cm^-1 */
#define ORC_T_MAX 0.5                 /* window rule: q_i - q_min <= T_MAX */
#define ORC_RENORM_PERIOD 128u        /* exponent renormalisation period   */
#define ORC_EB 16                     /* energies marched together (SIMD)  */

static inline uint64_t d2u(double x) {
    uint64_t u;
    memcpy(&u, &x, 8);
    return u;
}
static inline double u2d(uint64_t u) {
    double x;
    memcpy(&x, &u, 8);
    return x;
}

int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* torchrun exports OMP_NUM_THREADS=1 to every rank; the timed CPU baseline (rank 0 only) asks
 * for the host's cores explicitly. */
void orc_set_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

/* N1. V_i = De * (1 - exp(-a (r_i - re)))^2 on r_i = rmin + i*h,
 * h = (rmax - rmin)/(N-1).  Fills the slot left empty at
 * potential_source.hpp:223-225. */
void orc_morse_tabulate(double De, double re, double a, double rmin, double rmax, uint32_t N,
                        double* V) {
    const double h = (rmax - rmin) / (double)(N - 1);
    for (uint32_t i = 0; i < N; i++) {
        const double r = rmin + (double)i * h;
        const double t = 1.0 - exp(-a * (r - re));
        V[i] = (De * t) * t;
    }
}

/* Lennard-Jones 12-6 shifted so that the minimum is 0 and the asymptote De. */
void orc_lj_tabulate(double De, double re, double rmin, double rmax, uint32_t N, double* V) {
    const double h = (rmax - rmin) / (double)(N - 1);
    for (uint32_t i = 0; i < N; i++) {
        const double r  = rmin + (double)i * h;
        const double x  = re / r;
        const double x2 = x * x;
        const double x6 = (x2 * x2) * x2;
        V[i] = De * ((x6 * x6 - 2.0 * x6) + 1.0);
    }
}

/* Rotational states (spec DESIGN.md section 3.7): effective potential
 *   V_J(r_i) = V(r_i) + J(J+1) * (hbar^2 / 2 mu) / r_i^2,   r_i = rmin + i*h,
 * with hbar^2/2mu = h^2/(12 s) taken from the curve's own scale s, so no mass enters twice.
 * J = 0 copies the table bit for bit.  A point at r = 0 (the reference fixture's min_r,
 * python/test/test_libepseon_gpu.py:183-190) gives an infinite barrier: values above 1e300 are
 * clamped to 1e300 (a wall; the window rule of orc_prep excludes it).  NaN stays NaN. */
void orc_centrifugal(const double* V, uint32_t N, double s, double rmin, double h, uint32_t J,
                     double* out) {
    if (J == 0) {
        memcpy(out, V, sizeof(double) * N);
        return;
    }
    const double jj = (double)((uint64_t)J * ((uint64_t)J + 1u));
    const double cj = (jj * (h * h)) / (12.0 * s);
    for (uint32_t i = 0; i < N; i++) {
        const double r = rmin + (double)i * h;
        double       v = V[i] + cj / (r * r);
        if (v > 1e300) v = 1e300;
        out[i] = v;
    }
}

/* s = h^2 * (2 mu / hbar^2) / 12 with mu = m0 m1/(m0+m1): maps an energy in
 * cm^-1 to the dimensionless Numerov variable.  Masses are
 * VibwaAlgorithmConfig::mass_atom_{0,1} (algorithm_config.hpp:78-79). */
double orc_scale(double m0, double m1, double h) {
    const double mu = (m0 * m1) / (m0 + m1);
    const double c  = mu / ORC_HBAR2_OVER_2;
    return ((h * h) * c) / 12.0;
}

/* Preparation: q_i = s V_i; integration window [i0, iend] around the minimum
 * on which q_i - q_min <= T_MAX (so f_i = 1 - (q_i - e) > 0 for every
 * admissible energy: the Sturm-sequence property needs it); coefficient table
 * F_k = (1 - q_{i0+k}) / 12,  k = 0 .. n_steps-1  (so f_k/12 = F_k + e/12).
 * F must hold N doubles.  Returns 0, or -1 when the table is unusable
 * (non-finite values, or fewer than 2 steps). */
int orc_prep_form(const double* V, uint32_t N, double s, int form, double* F, uint32_t* i0_out,
                  uint32_t* n_steps_out, double* vmin_out);
int orc_prep(const double* V, uint32_t N, double s, double* F, uint32_t* i0_out,
             uint32_t* n_steps_out, double* vmin_out) {
    return orc_prep_form(V, N, s, 0, F, i0_out, n_steps_out, vmin_out);
}

/* form 0: F_k = (1 - q_k)/12 (X form).  form 1: A_k = 12 q_k (D form). */
int orc_prep_form(const double* V, uint32_t N, double s, int form, double* F, uint32_t* i0_out,
                  uint32_t* n_steps_out, double* vmin_out) {
    if (N < 3) return -1;
    uint32_t m = 0;
    for (uint32_t i = 0; i < N; i++) {
        if (!isfinite(V[i])) return -1;
        if (s * V[i] < s * V[m]) m = i;
    }
    const double qmin = s * V[m];
    const double thr  = qmin + ORC_T_MAX;
    uint32_t     ilo  = 0;
    for (uint32_t j = 0; j < m; j++)
        if (s * V[j] > thr) ilo = j + 1;
    uint32_t ihi = N - 1;
    for (uint32_t j = N - 1; j > m; j--)
        if (s * V[j] > thr) ihi = j - 1;
    const uint32_t i0   = ilo < 1 ? 1 : ilo;
    const uint32_t iend = (ihi + 1 < N - 1) ? ihi + 1 : N - 1;
    if (iend < i0 + 2) return -1;
    const uint32_t n = iend - i0;
    for (uint32_t k = 0; k < n; k++) {
        const double q = s * V[i0 + k];
        F[k]           = form == 0 ? (1.0 - q) / 12.0 : 12.0 * q;
    }
    *i0_out      = i0;
    *n_steps_out = n;
    *vmin_out    = V[m];
    return 0;
}

/* Exponent renormalisation: scale (X, S) by the power of two that brings
 * |X| into [1,2); exact, so signs and all later bits are those of the
 * unscaled recurrence.  Skipped when X is zero or subnormal. */
static inline void renorm(double* X, double* S, int32_t* expo) {
    const uint32_t ex = (uint32_t)(d2u(*X) >> 52) & 0x7ffu;
    if (ex != 0) {
        const double sc = u2d((uint64_t)(2046u - ex) << 52);
        *X *= sc;
        *S *= sc;
        *expo += (int32_t)ex - 1023;
    }
}

/* N2 + N3 for a block of up to ORC_EB energies.  Division-free Numerov in the
 * 4-operation "X form" (DESIGN.md section 3.3):  with f_k = 1 - (q_k - e),
 * the textbook recurrence  f_{k+1} psi_{k+1} = (12 - 10 f_k) psi_k - f_{k-1} psi_{k-1}
 * becomes, for  X_k = psi_k f_k prod_{j<k} f_j / 12^k  and  S_k = (f_k/12) X_k,
 *     fp   = F_k + e/12                    (1 add)
 *     Q    = fma(10, X_k, S_{k-1})         (1 fma)
 *     X'   = fma(-fp, Q, X_k)              (1 fma)
 *     S_k  = fp * X_k                      (1 mul)
 * X_k has the sign of psi_k while every f_j > 0; node = sign-bit flip between
 * consecutive X.  Start: X = 1, S = 0 (psi = 0 one point to the left). */
/* The ACCURATE recurrence, "D form" (form 1) -- 5 operations per step:
 * with Pi_k = prod_{j<k} f_j,  Y_k = f_k psi_k Pi_k  and  D_k = Y_{k+1} - f_k Y_k  (the scaled
 * first difference), the same textbook recurrence reads
 *     T12 = A_k - 12 e                      (1 add;  A_k = 12 q_k,  12 T_k = 12 (q_k - e))
 *     f   = fma(-1/12, T12, 1)              (1 fma)
 *     t   = T12 * Y_k                       (1 mul)
 *     D_k = fma(f, D_{k-1}, t)              (1 fma)
 *     Y'  = fma(f, Y_k, D_k)                (1 fma)
 * Start Y = 1, D = 1 (psi = 0 one point to the left).  Carrying the first difference instead of
 * the previous value keeps the rounding of one step from entering the SLOPE of the solution, which
 * in any value-carrying form (X, Y) is amplified by 1/sqrt(12 T) ~ 10^2..10^3 on fine grids:
 * measured eigenvalue noise (vs binary128, tests/test_accuracy_floor.py) falls from 2e-9 (C5) /
 * 2e-8 (C3) to ~1e-11.  Y_k has the sign of psi_k while every f_j > 0. */
static void sweep_block(const double* F, uint32_t n_steps, double s, const double* E, int nb, int form,
                        uint32_t* nodes, double* tail_mant, int32_t* tail_exp) {
    double   ep[ORC_EB], X[ORC_EB], S[ORC_EB];
    int64_t  cnt[ORC_EB];
    int32_t  ex[ORC_EB];
    for (int b = 0; b < ORC_EB; b++) {
        const double Eb = E[b < nb ? b : nb - 1];
        ep[b]  = form == 0 ? (s * Eb) / 12.0 : 12.0 * (s * Eb);
        X[b]   = 1.0;
        S[b]   = form == 0 ? 0.0 : 1.0;
        cnt[b] = 0;
        ex[b]  = 0;
    }
    if (form == 0) {
        for (uint32_t k = 0; k < n_steps; k++) {
            const double Fk = F[k];
#pragma omp simd
            for (int b = 0; b < ORC_EB; b++) {
                const double fp = Fk + ep[b];
                const double Q  = fma(10.0, X[b], S[b]);
                const double Xn = fma(-fp, Q, X[b]);
                S[b]            = fp * X[b];
                cnt[b] += (int64_t)((d2u(Xn) ^ d2u(X[b])) >> 63);
                X[b] = Xn;
            }
            if (((k + 1) % ORC_RENORM_PERIOD) == 0)
                for (int b = 0; b < ORC_EB; b++) renorm(&X[b], &S[b], &ex[b]);
        }
    } else {
        const double nC = -(1.0 / 12.0);
        for (uint32_t k = 0; k < n_steps; k++) {
            const double Ak = F[k];
#pragma omp simd
            for (int b = 0; b < ORC_EB; b++) {
                const double T12 = Ak - ep[b];
                const double f   = fma(nC, T12, 1.0);
                const double t   = T12 * X[b];
                S[b]             = fma(f, S[b], t);
                const double Yn  = fma(f, X[b], S[b]);
                cnt[b] += (int64_t)((d2u(Yn) ^ d2u(X[b])) >> 63);
                X[b] = Yn;
            }
            if (((k + 1) % ORC_RENORM_PERIOD) == 0)
                for (int b = 0; b < ORC_EB; b++) renorm(&X[b], &S[b], &ex[b]);
        }
    }
    for (int b = 0; b < nb; b++) {
        renorm(&X[b], &S[b], &ex[b]);
        if (nodes) nodes[b] = (uint32_t)cnt[b];
        if (tail_mant) tail_mant[b] = X[b];
        if (tail_exp) tail_exp[b] = ex[b];
    }
}

/* Sweep explicit energies E[0..nE). Any output pointer may be NULL. */
void orc_sweep_form(const double* F, uint32_t n_steps, double s, int form, const double* E, uint64_t nE,
                    uint32_t* nodes, double* tail_mant, int32_t* tail_exp);
void orc_sweep(const double* F, uint32_t n_steps, double s, const double* E, uint64_t nE,
               uint32_t* nodes, double* tail_mant, int32_t* tail_exp) {
    orc_sweep_form(F, n_steps, s, 0, E, nE, nodes, tail_mant, tail_exp);
}
void orc_sweep_form(const double* F, uint32_t n_steps, double s, int form, const double* E, uint64_t nE,
                    uint32_t* nodes, double* tail_mant, int32_t* tail_exp) {
    const int64_t nblk = (int64_t)((nE + ORC_EB - 1) / ORC_EB);
#pragma omp parallel for schedule(static)
    for (int64_t blk = 0; blk < nblk; blk++) {
        const uint64_t o  = (uint64_t)blk * ORC_EB;
        const int      nb = (int)((nE - o) < ORC_EB ? (nE - o) : ORC_EB);
        sweep_block(F, n_steps, s, E + o, nb, form, nodes ? nodes + o : 0, tail_mant ? tail_mant + o : 0,
                    tail_exp ? tail_exp + o : 0);
    }
}

/* Uniform grid E_j = E0 + j*dE, j = j0 .. j0+nE-1 (one multiply, one add). */
void orc_sweep_uniform_form(const double* F, uint32_t n_steps, double s, int form, double E0, double dE,
                            uint64_t j0, uint64_t nE, uint32_t* nodes, double* tail_mant, int32_t* tail_exp);
void orc_sweep_uniform(const double* F, uint32_t n_steps, double s, double E0, double dE,
                       uint64_t j0, uint64_t nE, uint32_t* nodes, double* tail_mant,
                       int32_t* tail_exp) {
    orc_sweep_uniform_form(F, n_steps, s, 0, E0, dE, j0, nE, nodes, tail_mant, tail_exp);
}
void orc_sweep_uniform_form(const double* F, uint32_t n_steps, double s, int form, double E0, double dE,
                            uint64_t j0, uint64_t nE, uint32_t* nodes, double* tail_mant, int32_t* tail_exp) {
    const int64_t nblk = (int64_t)((nE + ORC_EB - 1) / ORC_EB);
#pragma omp parallel for schedule(static)
    for (int64_t blk = 0; blk < nblk; blk++) {
        const uint64_t o  = (uint64_t)blk * ORC_EB;
        const int      nb = (int)((nE - o) < ORC_EB ? (nE - o) : ORC_EB);
        double         Eb[ORC_EB];
        for (int b = 0; b < nb; b++) Eb[b] = E0 + (double)(j0 + o + (uint64_t)b) * dE;
        sweep_block(F, n_steps, s, Eb, nb, form, nodes ? nodes + o : 0, tail_mant ? tail_mant + o : 0,
                    tail_exp ? tail_exp + o : 0);
    }
}

/* N5 + N6.  Locate levels vmin..vmax of one curve in [E_lo, E_hi]:
 *   coarse:  n_coarse energies on the affine grid E_j = E0 + (j0 + j)*dE (a slice of a global
 *            uniform grid; orc_solve_levels below passes E0 = E_lo, dE=(E_hi-E_lo)/(n_coarse-1), j0 = 0);
 *   bracket: j* = first j with nodes_j > v  ->  [E_{j*-1}, E_{j*}]
 *            (level absent if nodes_last <= v or nodes_0 > v -> NaN);
 *   refine:  rounds of M interior points E_m = lo + m*step, step=(hi-lo)/(M+1),
 *            m* = first m with nodes_m > v (M+1 if none), new bracket
 *            [E_{m*-1}, E_{m*}] with E_0 = lo, E_{M+1} = hi; a level is
 *            converged when hi-lo <= rel_tol*max(|lo|,|hi|) or the bracket
 *            stopped shrinking; at most max_rounds rounds;
 *   result:  0.5*(lo+hi).
 * All decisions are integer comparisons of node counts.  Returns the number of
 * refinement rounds used; *steps_done gets grid-steps x energies executed. */
int orc_solve_levels_grid_form(const double* F, uint32_t n_steps, double s, int form, double E0, double dE, uint64_t j0,
                               uint32_t n_coarse, uint32_t vmin, uint32_t vmax, uint32_t M, double rel_tol,
                               uint32_t max_rounds, double* levels, double* widths, uint32_t* n_below_hi,
                               uint32_t* n_first, uint64_t* steps_done);
int orc_solve_levels_grid(const double* F, uint32_t n_steps, double s, double E0, double dE, uint64_t j0,
                          uint32_t n_coarse, uint32_t vmin, uint32_t vmax, uint32_t M, double rel_tol,
                          uint32_t max_rounds, double* levels, double* widths, uint32_t* n_below_hi,
                          uint32_t* n_first, uint64_t* steps_done) {
    return orc_solve_levels_grid_form(F, n_steps, s, 0, E0, dE, j0, n_coarse, vmin, vmax, M, rel_tol, max_rounds, levels,
                                      widths, n_below_hi, n_first, steps_done);
}
int orc_solve_levels_grid_form(const double* F, uint32_t n_steps, double s, int form, double E0, double dE, uint64_t j0,
                               uint32_t n_coarse, uint32_t vmin, uint32_t vmax, uint32_t M, double rel_tol,
                               uint32_t max_rounds, double* levels, double* widths, uint32_t* n_below_hi,
                               uint32_t* n_first, uint64_t* steps_done) {
    const uint32_t nlev  = vmax - vmin + 1;
    uint64_t       steps = 0;
    uint32_t*      nodes = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)(n_coarse > M ? n_coarse : M));
    double*        lo    = (double*)malloc(sizeof(double) * nlev);
    double*        hi    = (double*)malloc(sizeof(double) * nlev);
    uint8_t*       act   = (uint8_t*)malloc(nlev);

    orc_sweep_uniform_form(F, n_steps, s, form, E0, dE, j0, n_coarse, nodes, 0, 0);
    steps += (uint64_t)n_steps * n_coarse;
    if (n_below_hi) *n_below_hi = nodes[n_coarse - 1];
    if (n_first) *n_first = nodes[0];
    for (uint32_t l = 0; l < nlev; l++) {
        const uint32_t v = vmin + l;
        act[l]           = 0;
        lo[l] = hi[l] = NAN;
        if (nodes[n_coarse - 1] <= v || nodes[0] > v) continue;
        uint32_t j = 1;
        while (nodes[j] <= v) j++;
        lo[l]  = E0 + (double)(j0 + j - 1) * dE;
        hi[l]  = E0 + (double)(j0 + j) * dE;
        act[l] = 1;
    }
    uint32_t round = 0;
    for (; round < max_rounds; round++) {
        int any = 0;
        for (uint32_t l = 0; l < nlev; l++) {
            if (!act[l]) continue;
            const double w   = hi[l] - lo[l];
            const double mag = fabs(lo[l]) > fabs(hi[l]) ? fabs(lo[l]) : fabs(hi[l]);
            if (w <= rel_tol * mag) act[l] = 0;
            else any = 1;
        }
        if (!any) break;
        for (uint32_t l = 0; l < nlev; l++) {
            if (!act[l]) continue;
            const uint32_t v    = vmin + l;
            const double   step = (hi[l] - lo[l]) / (double)(M + 1);
            /* points m = 1..M are j = 1..M of the grid E0=lo, dE=step */
            orc_sweep_uniform_form(F, n_steps, s, form, lo[l], step, 1, M, nodes, 0, 0);
            steps += (uint64_t)n_steps * M;
            uint32_t m = 1;
            while (m <= M && nodes[m - 1] <= v) m++;
            const double nlo = (m == 1) ? lo[l] : lo[l] + (double)(m - 1) * step;
            const double nhi = (m == M + 1) ? hi[l] : lo[l] + (double)m * step;
            if (!((nhi - nlo) < (hi[l] - lo[l]))) act[l] = 0; /* no progress */
            lo[l] = nlo;
            hi[l] = nhi;
        }
    }
    for (uint32_t l = 0; l < nlev; l++) {
        levels[l] = 0.5 * (lo[l] + hi[l]);
        if (widths) widths[l] = hi[l] - lo[l];
    }
    if (steps_done) *steps_done = steps;
    free(nodes);
    free(lo);
    free(hi);
    free(act);
    return (int)round;
}

/* Uniform coarse grid over [E_lo, E_hi]: dE = (E_hi - E_lo)/(n_coarse - 1), j0 = 0. */
int orc_solve_levels(const double* F, uint32_t n_steps, double s, double E_lo, double E_hi,
                     uint32_t n_coarse, uint32_t vmin, uint32_t vmax, uint32_t M, double rel_tol,
                     uint32_t max_rounds, double* levels, double* widths, uint32_t* n_below_hi,
                     uint64_t* steps_done) {
    const double dE = (E_hi - E_lo) / (double)(n_coarse - 1);
    return orc_solve_levels_grid(F, n_steps, s, E_lo, dE, 0, n_coarse, vmin, vmax, M, rel_tol, max_rounds,
                                 levels, widths, n_below_hi, 0, steps_done);
}

/* ---------------------------------------------------------------------------------------------
 * Outward/inward matching level search (Cooley; spec DESIGN.md section 3.9).  Mirrors
 * csrc/cooley_search.cuh operation for operation: the window is cut into segments of
 * L = max(64, ceil(n/256)) steps; pass 1 = two basis solutions per segment, a serial chain gives
 * the true entry state of every segment, pass 2 re-marches every segment counting sign flips and
 * accumulating the norm; the sums are added in segment order.  A is the D-form table (12 q_k).
 * ------------------------------------------------------------------------------------------- */
#define ORC_COOLEY_SEGS 256u
#define ORC_COOLEY_MINSEG 64u
typedef struct {
    double  u, d;
    int32_t e;
} cst;

static void c_renorm(cst* x, double* acc, double* psi_prev) {
    uint32_t ex = (uint32_t)(d2u(x->u) >> 52) & 0x7ffu;
    if (ex == 0) ex = (uint32_t)(d2u(x->d) >> 52) & 0x7ffu;
    if (ex != 0 && ex != 1023u) {
        const double sc = u2d((uint64_t)(2046u - ex) << 52);
        x->u *= sc;
        x->d *= sc;
        x->e += (int32_t)ex - 1023;
        if (acc) *acc = (*acc * sc) * sc;
        if (psi_prev) *psi_prev = *psi_prev * sc;
    }
}
static void c_coef(double Ak, double e12, double* g, double* r) {
    const double T12 = Ak - e12;
    const double f   = fma(-(1.0 / 12.0), T12, 1.0);
    *r               = 1.0 / f;
    *g               = T12 * *r;
}
static void c_step(cst* x, double g) {
    x->d = fma(g, x->u, x->d);
    x->u = x->u + x->d;
}
static void c_basis(const double* A, double e12, uint32_t k0, uint32_t k1, int dir, cst* a, cst* b) {
    *a = (cst){1.0, 0.0, 0};
    *b = (cst){0.0, 1.0, 0};
    for (uint32_t i = 0; i < k1 - k0; i++) {
        const uint32_t k = dir > 0 ? k0 + i : k1 - 1 - i;
        double         g, r;
        c_coef(A[k], e12, &g, &r);
        c_step(a, g);
        c_step(b, g);
        if ((i & 127u) == 127u) {
            c_renorm(a, 0, 0);
            c_renorm(b, 0, 0);
        }
    }
    c_renorm(a, 0, 0);
    c_renorm(b, 0, 0);
}
static cst c_apply(const cst* a, const cst* b, const cst* x) {
    int sh = b->e - a->e;
    sh     = sh > 1000 ? 1000 : (sh < -1000 ? -1000 : sh);
    const double sc = u2d((uint64_t)(1023 + sh) << 52); /* 2^sh, exact */
    const double bu = b->u * sc, bd = b->d * sc;
    cst          y;
    y.u = fma(a->u, x->u, bu * x->d);
    y.d = fma(a->d, x->u, bd * x->d);
    y.e = x->e + a->e;
    c_renorm(&y, 0, 0);
    return y;
}
static double c_pow2mul(double x, int sh) {
    sh = sh > 2000 ? 2000 : (sh < -2000 ? -2000 : sh);
    const int h1 = sh / 2, h2 = sh - h1;
    x = x * u2d((uint64_t)(1023 + h1) << 52);
    return x * u2d((uint64_t)(1023 + h2) << 52);
}
static void c_remarch(const double* A, double e12, uint32_t k0, uint32_t k1, int dir, cst* x, double psi_prev,
                      uint32_t* flips, double* acc_out) {
    uint32_t fl  = 0;
    double   acc = 0.0;
    for (uint32_t i = 0; i < k1 - k0; i++) {
        const uint32_t k = dir > 0 ? k0 + i : k1 - 1 - i;
        double         g, r;
        c_coef(A[k], e12, &g, &r);
        const double psi = x->u * r;
        acc              = fma(psi, fma(10.0, psi, 2.0 * psi_prev), acc);
        psi_prev         = psi;
        const uint64_t before = d2u(x->u);
        c_step(x, g);
        fl += (uint32_t)((before ^ d2u(x->u)) >> 63);
        if ((i & 127u) == 127u) c_renorm(x, &acc, &psi_prev);
    }
    c_renorm(x, &acc, &psi_prev);
    *flips   = fl;
    *acc_out = acc;
}

/* One level: bracket [lo, hi] with nodes(lo) <= v < nodes(hi).  Returns the iterations used;
 * *E_out the level, *width_out the magnitude of the last correction. */
int orc_cooley_level(const double* A, uint32_t n, double s, uint32_t v, double lo, double hi, double rel_tol,
                     uint32_t max_iter, int open_tail, uint32_t seg_len, double* E_out, double* width_out) {
    /* seg_len: the segment length the device chose (eps_cooley_segment_length); 0 = the few-item rule */
    const uint32_t L0 = (n + ORC_COOLEY_SEGS - 1) / ORC_COOLEY_SEGS;
    uint32_t       L  = seg_len ? seg_len : L0;
    L                 = L > ORC_COOLEY_MINSEG ? L : ORC_COOLEY_MINSEG;
    const uint32_t S  = (n + L - 1) / L;
    cst      fa[ORC_COOLEY_SEGS], fb[ORC_COOLEY_SEGS], ba[ORC_COOLEY_SEGS], bb[ORC_COOLEY_SEGS];
    cst      fend[ORC_COOLEY_SEGS], bend[ORC_COOLEY_SEGS];
    double   accs[ORC_COOLEY_SEGS];
    uint32_t flips[ORC_COOLEY_SEGS];
    double   E = 0.5 * (lo + hi), last = NAN;
    uint32_t it = 0;
    if (open_tail) lo = lo - (hi - lo); /* soft window for the level of the unbounded problem */
    for (; it < max_iter; it++) {
        const double e12 = 12.0 * (s * E);
        uint32_t     best = 0;
        for (uint32_t k = 0; k < n; k++)
            if (A[k] - e12 < 0.0) best = k + 1;
        const uint32_t ktp = best ? best - 1 : 0;
        uint32_t       sm  = ktp / L;
        sm                 = sm < 1 ? 1 : sm;
        const uint32_t smx = (n - 2) / L;
        sm                 = sm > smx ? smx : sm;
        const uint32_t m   = sm * L;
        for (uint32_t t = 0; t < S; t++) {
            const uint32_t k0 = t * L < n ? t * L : n, k1 = k0 + L < n ? k0 + L : n;
            c_basis(A, e12, k0, k1, +1, &fa[t], &fb[t]);
            if (t >= sm) c_basis(A, e12, t == sm ? m + 1 : k0, k1, -1, &ba[t], &bb[t]);
        }
        cst x = {1.0, 1.0, 0};
        for (uint32_t t = 0; t < S; t++) {
            const cst y = c_apply(&fa[t], &fb[t], &x);
            fa[t]       = x;
            x           = y;
        }
        double rho = 0.0;
        if (open_tail) { /* decaying solution of the recurrence with the last coefficient frozen */
            double gt, rt;
            c_coef(A[n - 1], e12, &gt, &rt);
            if (gt > 0.0) rho = fma(0.5, gt, 1.0) - sqrt(gt * fma(0.25, gt, 1.0));
        }
        cst z = {1.0, 1.0 - rho, 0};
        for (uint32_t t = S; t-- > sm;) {
            const cst y = c_apply(&ba[t], &bb[t], &z);
            ba[t]       = z;
            z           = y;
        }
        for (uint32_t t = 0; t < S; t++) {
            const uint32_t k0 = t * L < n ? t * L : n, k1 = k0 + L < n ? k0 + L : n;
            cst            xe = fa[t];
            double         gp, rp, pp = 0.0, acc = 0.0, acc_in = 0.0;
            uint32_t       fl = 0, fdummy;
            if (k0 > 0) {
                c_coef(A[k0 - 1], e12, &gp, &rp);
                pp = (xe.u - xe.d) * rp;
            }
            c_remarch(A, e12, k0, k1, +1, &xe, pp, &fl, &acc);
            flips[t] = fl;
            fend[t]  = xe;
            accs[t]  = acc;
            if (t >= sm) {
                cst    ze = ba[t];
                double pq = 0.0;
                if (k1 < n) {
                    c_coef(A[k1], e12, &gp, &rp);
                    pq = (ze.u - ze.d) * rp;
                }
                c_remarch(A, e12, t == sm ? m + 1 : k0, k1, -1, &ze, pq, &fdummy, &acc_in);
                accs[t] = acc_in;
                bend[t] = ze;
            }
        }
        uint32_t nodes = 0;
        for (uint32_t t = 0; t < S; t++) nodes += flips[t];
        const cst xo = fend[sm - 1], xi = bend[sm];
        double    gm, rm, gl, rl, gr, rr;
        c_coef(A[m], e12, &gm, &rm);
        c_coef(A[m - 1], e12, &gl, &rl);
        c_coef(A[m + 1], e12, &gr, &rr);
        const double a_out = xo.d / xo.u, b_in = xi.d / xi.u;
        const double R     = (-a_out - b_in) - gm;
        double       Nrm   = 0.0;
        for (uint32_t t = 0; t < sm; t++) {
            Nrm = Nrm + c_pow2mul(accs[t], 2 * (fend[t].e - xo.e));
        }
        Nrm = Nrm / (xo.u * xo.u);
        double Nin = 0.0;
        for (uint32_t t = S; t-- > sm;) {
            Nin = Nin + c_pow2mul(accs[t], 2 * (bend[t].e - xi.e));
        }
        Nin = Nin / (xi.u * xi.u);
        const double pm = rm, pl = (1.0 - a_out) * rl, pr = (1.0 - b_in) * rr;
        const double Nm = fma(pm, fma(10.0, pm, 2.0 * (pl + pr)), 0.0);
        const double N  = (Nrm + Nin) + Nm;
        const double de = -((rm * R) / N);
        double       dE = de / s;
        if (!open_tail) {
            if (nodes > v) hi = E;
            else lo = E;
        }
        double En = E + dE;
        int    done;
        if (dE == dE && fabs(dE) <= rel_tol * fabs(E)) { /* converged: the correction is below the tolerance */
            if (!(En >= lo && En <= hi)) En = E;
            done = 1;
        } else if (!(dE == dE) || !(En > lo && En < hi)) { /* NaN, or the step leaves the bracket: bisect */
            En   = 0.5 * (lo + hi);
            dE   = En - E;
            done = (fabs(dE) <= rel_tol * fabs(En)) || !(En != E);
        } else {
            done = 0;
        }
        last = fabs(dE);
        E    = En;
        if (done) {
            it++;
            break;
        }
    }
    *E_out = E;
    if (width_out) *width_out = last;
    return (int)it;
}

/* N7.  Normalised wavefunction of one level at energy E on the integration
 * window (psi[0..n_steps), psi_k = psi(r_{i0+k}); Dirichlet zeros at k = -1 and
 * k = n_steps).  DESIGN.md section 3.6:
 *   fp_k = F_k + e/12,  r_k = 1/fp_k,  c_k = r_k - 10;
 *   the Numerov variable u_k = fp_k psi_k obeys  u_{k+1} = fma(c_k, u_k, -u_{k-1});
 *   outward march k = 0..m from (u_{-1}, u_0) = (0, fp_0), inward march
 *   k = n-1..m from (u_n, u_{n-1}) = (0, fp_{n-1}), psi_k = u_k * r_k;
 *   m = outer classical turning point = max{k : fp_k > 1/12}, clamped to [1, n-2];
 *   (u_cur, u_prev) rescaled by an exact power of two every 128 march steps,
 *   the cumulative exponent kept per 128-block;
 *   the inward branch is scaled by psi_out(m)/psi_in(m); all values are
 *   referred to the largest binary exponent, then  psi /= sqrt(h * sum psi^2).
 * The outward branch starts positive, so psi > 0 on its first lobe.
 * Returns m, or -1 when E lies below the whole table (psi filled with zeros). */
int64_t orc_wavefunction(const double* F, uint32_t n_steps, double s, double E, double h,
                         double* psi) {
    const uint32_t n  = n_steps;
    const double   ep = (s * E) / 12.0;
    int64_t        m  = -1;
    for (uint32_t k = 0; k < n; k++) psi[k] = 0.0;
    if (!(E == E) || n < 3) return -1;
    for (uint32_t k = 0; k < n; k++)
        if (F[k] + ep > 1.0 / 12.0) m = k;
    if (m < 0) return -1;
    if (m < 1) m = 1;
    if (m > (int64_t)n - 2) m = (int64_t)n - 2;
    int32_t* cexp = (int32_t*)malloc(sizeof(int32_t) * n);
    double   a_in_m = 0.0;
    int32_t  c_in_m = 0, c_out_m = 0;
    for (int dir = 0; dir < 2; dir++) {
        const uint32_t count = dir == 0 ? (uint32_t)m + 1 : n - (uint32_t)m;
        double         u_cur = 0.0, u_prev = 0.0;
        int32_t        cum = 0;
        for (uint32_t j = 0; j < count; j++) {
            const uint32_t k  = dir == 0 ? j : n - 1 - j;
            const double   fp = F[k] + ep;
            const double   r  = 1.0 / fp;
            const double   c  = r - 10.0;
            if (j == 0) u_cur = fp;
            if (j > 0 && (j % ORC_RENORM_PERIOD) == 0) renorm(&u_cur, &u_prev, &cum);
            const double a = u_cur * r;
            if (dir == 1 && k == (uint32_t)m) {
                a_in_m = a;
                c_in_m = cum;
            } else {
                psi[k]  = a;
                cexp[k] = cum;
            }
            if (dir == 0 && k == (uint32_t)m) c_out_m = cum;
            const double un = fma(c, u_cur, -u_prev);
            u_prev          = u_cur;
            u_cur           = un;
        }
    }
    const double rho = (a_in_m != 0.0) ? psi[m] / a_in_m : 1.0;
    int32_t      emax = INT32_MIN;
    for (uint32_t k = 0; k < n; k++) {
        if (k > (uint32_t)m) {
            psi[k]  = psi[k] * rho;
            cexp[k] = cexp[k] - c_in_m + c_out_m;
        }
        if (psi[k] != 0.0) {
            const int32_t e = cexp[k] + (int32_t)ilogb(psi[k]);
            if (e > emax) emax = e;
        }
    }
    double sum = 0.0;
    for (uint32_t k = 0; k < n; k++) {
        int64_t sh = (int64_t)cexp[k] - (int64_t)emax;
        if (sh < -2200) sh = -2200;
        psi[k] = scalbn(psi[k], (int)sh);
        sum    = fma(psi[k], psi[k], sum);
    }
    const double nrm = sqrt(h * sum);
    for (uint32_t k = 0; k < n; k++) psi[k] = psi[k] / nrm;
    free(cexp);
    return m;
}

/* A-posteriori energy correction of a located level (Cooley's correction
 * written as the Rayleigh quotient of the Numerov pencil).  On the window, with u_k = fp_k psi_k,
 * fp_k = F_k + e/12 and psi_{-1} = psi_n = 0, the matched solution satisfies the Numerov equation
 *     u_{k+1} - 2 u_k + u_{k-1} = (1 - 12 fp_k) psi_k
 * at every k except the matching point m, where it leaves the residual r_m.  The pencil is
 * (A + e B) psi = 0 with B = tridiag(1, 10, 1)/12, so to first order the eigenvalue is
 *     e + de,   de = - psi_m r_m / D,   D = sum_k psi_k (psi_{k-1} + 10 psi_k + psi_{k+1}) / 12,
 * and dE = de / s with an error quadratic in the energy error.  psi is the output of
 * orc_wavefunction (window part), m its return value.  NaN when there is no matching point. */
double orc_level_correction(const double* F, uint32_t n_steps, double s, double E, const double* psi,
                            int64_t m) {
    const uint32_t n = n_steps;
    if (m < 1 || m > (int64_t)n - 2) return NAN;
    const double ep = (s * E) / 12.0;
    const double f0 = F[m] + ep, fl = F[m - 1] + ep, fr = F[m + 1] + ep;
    const double u0 = f0 * psi[m], ul = fl * psi[m - 1], ur = fr * psi[m + 1];
    const double r  = ((ur - u0) - (u0 - ul)) - (1.0 - 12.0 * f0) * psi[m];
    double       D  = 0.0;
    for (uint32_t k = 0; k < n; k++) {
        const double left = k > 0 ? psi[k - 1] : 0.0, right = k + 1 < n ? psi[k + 1] : 0.0;
        D = fma(psi[k], fma(10.0, psi[k], left + right), D);
    }
    D = D / 12.0;
    return -((psi[m] * r) / D) / s;
}

/* N1 (tabulated sources).  Natural cubic spline through the knots (r_k, V_k), k = 0..K-1 (r strictly
 * increasing), resampled on r_i = rmin + i*h, h = (rmax - rmin)/(N - 1).  Fills the slot of
 * PotentialFileLoader::get_potential_data (potential_source.hpp:89-91) for ab initio tables on
 * non-uniform grids.  Second derivatives m_k from the tridiagonal system
 *     h_{k-1} m_{k-1} + 2 (h_{k-1} + h_k) m_k + h_k m_{k+1} = 6 ((V_{k+1}-V_k)/h_k - (V_k-V_{k-1})/h_{k-1}),
 * m_0 = m_{K-1} = 0 (Thomas algorithm, forward sweep then back substitution); on [r_k, r_{k+1}]
 *     S = a + dx (b + dx (c + dx d)),  dx = r - r_k,
 *     a = V_k, b = (V_{k+1}-V_k)/h_k - h_k (2 m_k + m_{k+1})/6, c = m_k/2, d = (m_{k+1}-m_k)/(6 h_k),
 * evaluated as fma(fma(fma(d,dx,c),dx,b),dx,a).  Points left of r_0 / right of r_{K-1} use the first /
 * last interval (extrapolation).  coef (4*(K-1) doubles, may be NULL) receives a,b,c,d per interval.
 * Returns 0, or -1 on bad input. */
int orc_spline_coefficients(const double* r, const double* V, uint32_t K, double* coef) {
    if (K < 3) return -1;
    for (uint32_t k = 0; k + 1 < K; k++)
        if (!(r[k + 1] > r[k])) return -1;
    double* m  = (double*)calloc(K, sizeof(double));
    double* cp = (double*)calloc(K, sizeof(double));
    double* dp = (double*)calloc(K, sizeof(double));
    /* interior rows k = 1..K-2; unknowns m_1..m_{K-2} */
    for (uint32_t k = 1; k + 1 < K; k++) {
        const double hl = r[k] - r[k - 1], hr = r[k + 1] - r[k];
        const double diag = 2.0 * (hl + hr);
        const double rhs  = 6.0 * ((V[k + 1] - V[k]) / hr - (V[k] - V[k - 1]) / hl);
        const double sub  = (k == 1) ? 0.0 : hl;
        const double den  = diag - sub * cp[k - 1];
        cp[k]             = (k + 2 < K) ? hr / den : 0.0;
        dp[k]             = (rhs - sub * dp[k - 1]) / den;
    }
    for (uint32_t k = K - 2; k >= 1; k--) m[k] = dp[k] - cp[k] * m[k + 1];
    for (uint32_t k = 0; k + 1 < K; k++) {
        const double h = r[k + 1] - r[k];
        coef[4 * k + 0] = V[k];
        coef[4 * k + 1] = (V[k + 1] - V[k]) / h - (h * (2.0 * m[k] + m[k + 1])) / 6.0;
        coef[4 * k + 2] = m[k] / 2.0;
        coef[4 * k + 3] = (m[k + 1] - m[k]) / (6.0 * h);
    }
    free(m);
    free(cp);
    free(dp);
    return 0;
}

int orc_spline_resample(const double* r, const double* V, uint32_t K, double rmin, double rmax,
                        uint32_t N, double* out) {
    if (N < 2) return -1;
    double* coef = (double*)malloc(sizeof(double) * 4 * (K > 1 ? K - 1 : 1));
    if (orc_spline_coefficients(r, V, K, coef) != 0) {
        free(coef);
        return -1;
    }
    const double h = (rmax - rmin) / (double)(N - 1);
    for (uint32_t i = 0; i < N; i++) {
        const double x = rmin + (double)i * h;
        /* interval: largest k <= K-2 with r_k <= x (k = 0 when x < r_0) */
        uint32_t lo = 0, hi = K - 1;
        while (hi - lo > 1) {
            const uint32_t mid = (lo + hi) / 2;
            if (r[mid] <= x) lo = mid;
            else hi = mid;
        }
        const double  dx = x - r[lo];
        const double* c  = coef + 4 * lo;
        out[i]           = fma(fma(fma(c[3], dx, c[2]), dx, c[1]), dx, c[0]);
    }
    free(coef);
    return 0;
}
