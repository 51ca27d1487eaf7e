// cooley_search.cuh -- outward/inward matching level search (spec DESIGN.md section 3.9;
// oracle: orc_cooley_level).  Fills, like the rest of this directory, the compute slot the reference
// leaves empty (VibwaAlgorithm<FP>::run, vibwa.hpp:605-637); the parameter it serves is
// VibwaAlgorithmConfig::min_distance_to_asymptote / min_level..max_level (algorithm_config.hpp:78-83).
//
// The k-section search refines a level with full sweeps of hundreds of trial energies per round.
// Cooley's method converges quadratically from the coarse bracket instead: at the trial energy the
// solution is marched OUTWARD from the left wall and INWARD from the right wall to a matching point
// m near the outer classical turning point; the mismatch of the two branches is the residual R of the
// Numerov equation at m, and the Rayleigh quotient of the pencil gives the correction
//
//     dE = -(u_m^2 / f_m) R / (10 sum psi_k^2 + 2 sum psi_k psi_{k+1}) / s        (section 3.8).
//
// One CTA per (curve, level) iterates to convergence WITHOUT host round trips.  The marches are not
// serial: the window is cut into one segment per thread; pass 1 marches two basis solutions per
// segment (transfer matrix), thread 0 chains the matrices into the true entry state of every
// segment, pass 2 re-marches every segment from its entry state counting sign flips and
// accumulating the norm.  The FULL outward solution is also chained, so nodes(E) -- the number of
// eigenvalues below E exactly as the sweep defines it -- updates a rigorous bracket each iteration:
// a step that leaves the bracket is replaced by its midpoint, and the result is the level the
// coarse sweep bracketed, never a neighbour.  The marches use the difference form with one division
// per step (u'' = g u, g = 12T/(1 - T); state (u_k, d_{k-1} = u_k - u_{k-1})): the work is tiny next
// to a sweep, and the difference form keeps the slope free of rounding noise (section 3.3).
// Every operation and every summation order is fixed, so the oracle mirrors it bit for bit.
#pragma once
#include "numerov_kernels.cuh"

namespace eps {

constexpr int      kCooleyThreads = 256;  // most segments (= threads) per (curve, level)
constexpr uint32_t kCooleyMinSeg  = 64;   // shortest segment (steps)

// Segment length for a window of n steps.  Two regimes (measured, profiles/r2_cooley.md):
//   few (curve, level) items  -- latency-bound: as many segments as there are threads,
//                                L = ceil(n / kCooleyThreads);
//   many items (a batch)      -- throughput-bound: the marches cost ~4 L steps per thread, the serial
//                                chain of thread 0 ~2 n/L matrix applications and a barrier each:
//                                L ~ sqrt(2.5 n) balances them (fewer, longer segments).
// At least kCooleyMinSeg steps per segment.  Integer arithmetic: the oracle is handed the same L.
__host__ __device__ inline uint32_t cooley_segment_length(uint32_t n, uint64_t n_items) {
    const uint32_t by_threads = (n + kCooleyThreads - 1) / kCooleyThreads;
    uint32_t       L          = by_threads;
    if (n_items >= 1024) {
        uint32_t r = 1;
        while (static_cast<uint64_t>(r) * r * 2u < 5ull * n) r++;  // r = ceil(sqrt(2.5 n))
        r = (r + 7u) & ~7u;
        L = r > by_threads ? r : by_threads;
    }
    return L > kCooleyMinSeg ? L : kCooleyMinSeg;
}

struct CSt {  // value = (u, d) * 2^e
    double u, d;
    int    e;
};

__device__ __forceinline__ void cooley_renorm(CSt& x, double* acc, double* psi_prev) {
    uint32_t ex = (static_cast<uint32_t>(__double2hiint(x.u)) >> 20) & 0x7ffu;
    if (ex == 0) ex = (static_cast<uint32_t>(__double2hiint(x.d)) >> 20) & 0x7ffu;
    if (ex != 0 && ex != 1023u) {
        const double sc = __hiloint2double(static_cast<int>((2046u - ex) << 20), 0);
        x.u = __dmul_rn(x.u, sc);
        x.d = __dmul_rn(x.d, sc);
        x.e += static_cast<int>(ex) - 1023;
        if (acc) *acc = __dmul_rn(__dmul_rn(*acc, sc), sc);
        if (psi_prev) *psi_prev = __dmul_rn(*psi_prev, sc);
    }
}

// g_k = 12 T_k / f_k and r_k = 1 / f_k at trial energy e12 = 12 s E
__device__ __forceinline__ void cooley_coef(const double Ak, const double e12, double& g, double& r) {
    const double T12 = __dsub_rn(Ak, e12);
    const double f   = __fma_rn(-(1.0 / 12.0), T12, 1.0);
    r                = __ddiv_rn(1.0, f);
    g                = __dmul_rn(T12, r);
}

// One step k -> k+1 (or, on the reversed grid, k -> k-1) of the difference form.
__device__ __forceinline__ void cooley_step(CSt& x, const double g) {
    x.d = __fma_rn(g, x.u, x.d);
    x.u = __dadd_rn(x.u, x.d);
}

// The coefficients g_k, r_k do not depend on the solution: they are computed kCoefBlock at a time
// (independent divisions pipeline), and only the two-operation step stays on the dependent chain.
constexpr int kCoefBlock = 8;

// x * 2^sh, |sh| <= 2000, in one or two exact-power-of-two multiplications (deterministic; the oracle
// does the same).
__device__ __forceinline__ double cooley_pow2mul(double x, int sh) {
    sh = sh > 2000 ? 2000 : (sh < -2000 ? -2000 : sh);
    const int h1 = sh / 2, h2 = sh - h1;
    x = __dmul_rn(x, __hiloint2double((1023 + h1) << 20, 0));
    return __dmul_rn(x, __hiloint2double((1023 + h2) << 20, 0));
}

// Pass 1: transfer matrix of steps k = k0 .. k1-1 taken in direction dir (+1: k0 upwards; -1: from
// k1-1 downwards), as two basis solutions alpha = (1,0), beta = (0,1).
__device__ __forceinline__ void cooley_basis(const double* __restrict__ A, const double e12, uint32_t k0, uint32_t k1, int dir,
                                             CSt& a, CSt& b) {
    a = CSt{1.0, 0.0, 0};
    b = CSt{0.0, 1.0, 0};
    const uint32_t len = k1 - k0;
    for (uint32_t i0 = 0; i0 < len; i0 += kCoefBlock) {
        double g[kCoefBlock];
#pragma unroll
        for (int j = 0; j < kCoefBlock; j++) {
            const uint32_t i = min(i0 + j, len - 1);
            double         r;
            cooley_coef(A[dir > 0 ? k0 + i : k1 - 1 - i], e12, g[j], r);
        }
#pragma unroll
        for (int j = 0; j < kCoefBlock; j++)
            if (i0 + j < len) {
                cooley_step(a, g[j]);
                cooley_step(b, g[j]);
            }
        if (((i0 + kCoefBlock) & 127u) == 0u && i0 + kCoefBlock <= len) {
            cooley_renorm(a, nullptr, nullptr);
            cooley_renorm(b, nullptr, nullptr);
        }
    }
    cooley_renorm(a, nullptr, nullptr);
    cooley_renorm(b, nullptr, nullptr);
}

// entry' = [a b] entry  (2x2 matrix-vector with per-column exponents)
__device__ __forceinline__ CSt cooley_apply(const CSt& a, const CSt& b, const CSt& x) {
    // value = x.u * a + x.d * b, columns scaled by 2^a.e / 2^b.e: bring b to a's exponent
    int sh = b.e - a.e;
    sh     = sh > 1000 ? 1000 : (sh < -1000 ? -1000 : sh);
    const double sc = __hiloint2double((1023 + sh) << 20, 0);  // 2^sh, exact
    const double bu = __dmul_rn(b.u, sc), bd = __dmul_rn(b.d, sc);
    CSt          y;
    y.u = __fma_rn(a.u, x.u, __dmul_rn(bu, x.d));
    y.d = __fma_rn(a.d, x.u, __dmul_rn(bd, x.d));
    y.e = x.e + a.e;
    cooley_renorm(y, nullptr, nullptr);
    return y;
}

// Pass 2: re-march steps k0 .. k1-1 in direction dir from the true entry state; counts the sign flips
// of u and accumulates  acc += psi_k (10 psi_k + 2 psi_prev)  over the points k it leaves behind
// (psi_k = u_k r_k; psi_prev = the previously visited point, psi_entry_prev on entry).
__device__ __forceinline__ void cooley_remarch(const double* __restrict__ A, const double e12, uint32_t k0, uint32_t k1, int dir,
                                               CSt& x, double psi_prev, uint32_t& flips, double& acc) {
    flips = 0;
    acc   = 0.0;
    const uint32_t len = k1 - k0;
    for (uint32_t i0 = 0; i0 < len; i0 += kCoefBlock) {
        double g[kCoefBlock], r[kCoefBlock];
#pragma unroll
        for (int j = 0; j < kCoefBlock; j++) {
            const uint32_t i = min(i0 + j, len - 1);
            cooley_coef(A[dir > 0 ? k0 + i : k1 - 1 - i], e12, g[j], r[j]);
        }
#pragma unroll
        for (int j = 0; j < kCoefBlock; j++)
            if (i0 + j < len) {
                const double psi = __dmul_rn(x.u, r[j]);
                acc              = __fma_rn(psi, __fma_rn(10.0, psi, __dmul_rn(2.0, psi_prev)), acc);
                psi_prev         = psi;
                const uint32_t before = static_cast<uint32_t>(__double2hiint(x.u));
                cooley_step(x, g[j]);
                flips += (before ^ static_cast<uint32_t>(__double2hiint(x.u))) >> 31;
            }
        if (((i0 + kCoefBlock) & 127u) == 0u && i0 + kCoefBlock <= len) cooley_renorm(x, &acc, &psi_prev);
    }
    cooley_renorm(x, &acc, &psi_prev);
}

// state: 1 = active bracket [lo, hi] from the coarse sweep; on return levels[idx] = E (NaN for
// absent levels), widths[idx] = |last correction|, state 2, iters[idx] = iterations used.
__global__ void __launch_bounds__(kCooleyThreads, 2)
cooley_search_kernel(const double* __restrict__ Atab, const CurveDev* __restrict__ curves, uint32_t n_lev, uint32_t v_min,
                     double rel_tol, uint32_t max_iter, double* __restrict__ lo_g, double* __restrict__ hi_g,
                     uint32_t* __restrict__ state, double* __restrict__ levels, double* __restrict__ widths,
                     uint32_t* __restrict__ iters_out, unsigned long long* __restrict__ steps_done,
                     const int* __restrict__ stop_flag, const int open_tail) {
    __shared__ CSt      fa[kCooleyThreads], fb[kCooleyThreads];  // forward basis; reused: forward entry / end states
    __shared__ CSt      ba[kCooleyThreads], bb[kCooleyThreads];  // backward basis; reused: backward entry / end states
    __shared__ double   acc_s[kCooleyThreads];
    __shared__ uint32_t flips_s[kCooleyThreads];
    __shared__ uint32_t red[kCooleyThreads / 32];
    __shared__ double   E_sh, dE_sh;
    __shared__ uint32_t sm_sh, done_sh;

    const uint32_t idx = blockIdx.x, tid = threadIdx.x;
    const double   nan = __longlong_as_double(0x7ff8000000000000LL);
    if (state[idx] != 1u) {
        if (tid == 0) {
            levels[idx] = nan;
            widths[idx] = nan;
            if (iters_out) iters_out[idx] = 0;
        }
        return;
    }
    const CurveDev cv = curves[idx / n_lev];
    const uint32_t v  = v_min + idx % n_lev;
    const uint32_t n  = cv.n_steps;
    const double*  A  = Atab + cv.f_off;
    const uint32_t L  = cooley_segment_length(n, gridDim.x);
    const uint32_t S  = (n + L - 1) / L;  // segments in use (<= kCooleyThreads)
    const uint32_t k0 = min(n, tid * L), k1 = min(n, k0 + L);
    double         lo = lo_g[idx], hi = hi_g[idx];
    double         E  = __dmul_rn(0.5, __dadd_rn(lo, hi));
    // Open tail (EPS_SOLVE_OPEN_TAIL): the level of the UNBOUNDED problem lies below the box level the
    // coarse sweep bracketed; node counts (a property of the box problem) no longer bound it, so the
    // bracket is only a soft window [lo - (hi - lo), hi] that clamps the Newton steps.
    if (open_tail) lo = __dsub_rn(lo, __dsub_rn(hi, lo));
    double         last = nan;
    uint32_t       it = 0;
    if (tid == 0) {
        E_sh  = E;
        dE_sh = nan;
    }
    __syncthreads();

    for (; it < max_iter; it++) {
        if (*stop_flag != 0) break;
        const double e12 = __dmul_rn(12.0, __dmul_rn(cv.s, E));
        // ---- matching segment: the one holding the outer classical turning point max{k : 12 T_k < 0}
        uint32_t best = 0;
        for (uint32_t k = k0; k < k1; k++)
            if (__dsub_rn(A[k], e12) < 0.0) best = k + 1;
        for (int o = 16; o > 0; o >>= 1) best = max(best, __shfl_xor_sync(0xffffffffu, best, o));
        if ((tid & 31) == 0) red[tid >> 5] = best;
        __syncthreads();
        if (tid == 0) {
            uint32_t b = 0;
            for (uint32_t w = 0; w < (blockDim.x >> 5); w++) b = max(b, red[w]);
            const uint32_t ktp = b ? b - 1 : 0;
            uint32_t       sm  = ktp / L;
            sm                 = sm < 1 ? 1 : sm;
            const uint32_t smx = (n - 2) / L;  // m = sm * L <= n - 2
            sm                 = sm > smx ? smx : sm;
            sm_sh              = sm;           // (n - 2) / L >= 1 for every n >= 3 * kCooleyMinSeg; shorter windows: see host
        }
        __syncthreads();
        const uint32_t sm = sm_sh, m = sm * L;
        // ---- pass 1: transfer matrices (forward for every segment, backward right of m)
        if (tid < S) {
            cooley_basis(A, e12, k0, k1, +1, fa[tid], fb[tid]);
            if (tid >= sm) cooley_basis(A, e12, tid == sm ? m + 1 : k0, k1, -1, ba[tid], bb[tid]);
        }
        __syncthreads();
        // ---- chain: true entry state of every segment (overwrites the a-columns)
        if (tid == 0) {
            CSt x{1.0, 1.0, 0};  // (u_0, d_{-1}): psi = 0 one point to the left
            for (uint32_t s = 0; s < S; s++) {
                const CSt y = cooley_apply(fa[s], fb[s], x);
                fa[s]       = x;
                x           = y;
            }
            // (u_{n-1}, u_{n-1} - u_n).  Box: psi = 0 one point to the right, u_n = 0.  Open tail: the
            // decaying solution of the recurrence with the last coefficient frozen,
            // u_n / u_{n-1} = rho,  rho + 1/rho = 2 + g  ->  rho = 1 + g/2 - sqrt(g (1 + g/4)),  g = g_{n-1} > 0.
            double rho = 0.0;
            if (open_tail) {
                double gt, rt;
                cooley_coef(A[n - 1], e12, gt, rt);
                if (gt > 0.0)
                    rho = __dsub_rn(__fma_rn(0.5, gt, 1.0), __dsqrt_rn(__dmul_rn(gt, __fma_rn(0.25, gt, 1.0))));
            }
            CSt z{1.0, __dsub_rn(1.0, rho), 0};
            for (uint32_t s = S; s-- > sm;) {
                const CSt y = cooley_apply(ba[s], bb[s], z);
                ba[s]       = z;
                z           = y;
            }
        }
        __syncthreads();
        // ---- pass 2: flips of the full outward solution, norm of the matched solution
        double   acc = 0.0;
        uint32_t fl  = 0;
        CSt      xe{0.0, 0.0, 0}, ze{0.0, 0.0, 0};
        double   acc_in = 0.0;
        if (tid < S) {
            xe = fa[tid];
            double gp, rp, pp = 0.0;
            if (k0 > 0) {  // psi_{k0-1} = (u_{k0} - d_{k0-1}) r_{k0-1}
                cooley_coef(A[k0 - 1], e12, gp, rp);
                pp = __dmul_rn(__dsub_rn(xe.u, xe.d), rp);
            }
            cooley_remarch(A, e12, k0, k1, +1, xe, pp, fl, acc);
            if (tid >= sm) {
                ze = ba[tid];
                double pq = 0.0;
                if (k1 < n) {  // psi_{k1} = (u_{k1-1} - dtilde) r_{k1}
                    cooley_coef(A[k1], e12, gp, rp);
                    pq = __dmul_rn(__dsub_rn(ze.u, ze.d), rp);
                }
                uint32_t fdummy;
                cooley_remarch(A, e12, tid == sm ? m + 1 : k0, k1, -1, ze, pq, fdummy, acc_in);
            }
        }
        __syncthreads();  // everyone has read its entry state
        if (tid < S) {
            flips_s[tid] = fl;
            fa[tid]      = xe;  // end state of the forward march of this segment (state at k1)
            if (tid < sm) acc_s[tid] = acc;
            else {
                acc_s[tid] = acc_in;
                ba[tid]    = ze;  // end state of the backward march (state at k0, or at m for tid == sm)
            }
        }
        __syncthreads();
        if (tid == 0) {
            uint32_t nodes = 0;
            for (uint32_t s = 0; s < S; s++) nodes += flips_s[s];
            // the matched solution in units of u_m: outward branch / u_m^out, inward branch / u_m^in
            const CSt xo = fa[sm - 1];  // (u_m, d_{m-1}) of the outward branch
            const CSt xi = ba[sm];      // (u_m, u_m - u_{m+1}) of the inward branch
            double    gm, rm, gl, rl, gr, rr;
            cooley_coef(A[m], e12, gm, rm);
            cooley_coef(A[m - 1], e12, gl, rl);
            cooley_coef(A[m + 1], e12, gr, rr);
            const double a_out = __ddiv_rn(xo.d, xo.u), b_in = __ddiv_rn(xi.d, xi.u);
            const double R     = __dsub_rn(__dsub_rn(-a_out, b_in), gm);  // (u_{m+1} - 2 u_m + u_{m-1}) / u_m - g_m
            double       Nrm   = 0.0;
            for (uint32_t s = 0; s < sm; s++) {  // outward segments: units of their end scale -> units of u_m
                Nrm = __dadd_rn(Nrm, cooley_pow2mul(acc_s[s], 2 * (fa[s].e - xo.e)));
            }
            Nrm = __ddiv_rn(Nrm, __dmul_rn(xo.u, xo.u));
            double Nin = 0.0;
            for (uint32_t s = S; s-- > sm;) {
                Nin = __dadd_rn(Nin, cooley_pow2mul(acc_s[s], 2 * (ba[s].e - xi.e)));
            }
            Nin = __ddiv_rn(Nin, __dmul_rn(xi.u, xi.u));
            // the point m itself and its two cross terms: psi_m = r_m, psi_{m-1} = (1 - a_out) r_{m-1}, psi_{m+1} = (1 - b_in) r_{m+1}
            const double pm = rm, pl = __dmul_rn(__dsub_rn(1.0, a_out), rl), pr = __dmul_rn(__dsub_rn(1.0, b_in), rr);
            const double Nm = __fma_rn(pm, __fma_rn(10.0, pm, __dmul_rn(2.0, __dadd_rn(pl, pr))), 0.0);
            const double N  = __dadd_rn(__dadd_rn(Nrm, Nin), Nm);
            // f_m = 1 / r_m:  de = -(u_m^2 / f_m) R / N with u_m = 1  ->  -(r_m R) / N
            const double de = -__ddiv_rn(__dmul_rn(rm, R), N);
            double       dE = __ddiv_rn(de, cv.s);
            // rigorous bracket from the node count of the full outward solution (box problem only)
            if (!open_tail) {
                if (nodes > v) hi = E;
                else lo = E;
            }
            double   En = __dadd_rn(E, dE);
            uint32_t done;
            if (dE == dE && fabs(dE) <= __dmul_rn(rel_tol, fabs(E))) {  // converged: the correction is below the tolerance
                if (!(En >= lo && En <= hi)) En = E;
                done = 1;
            } else if (!(dE == dE) || !(En > lo && En < hi)) {  // NaN, or the step leaves the bracket: bisect
                En   = __dmul_rn(0.5, __dadd_rn(lo, hi));              // (open tail: back to the middle of the window)
                dE   = __dsub_rn(En, E);
                done = (fabs(dE) <= __dmul_rn(rel_tol, fabs(En)) || !(En != E)) ? 1u : 0u;
            } else {
                done = 0;
            }
            last    = fabs(dE);
            done_sh = done;
            E_sh    = En;
            dE_sh   = last;
        }
        __syncthreads();
        E = E_sh;
        if (done_sh) {
            it++;
            break;
        }
    }
    if (tid == 0) {
        levels[idx] = E;
        widths[idx] = dE_sh;
        lo_g[idx]   = lo;
        hi_g[idx]   = hi;
        state[idx]  = 2;
        if (iters_out) iters_out[idx] = it;
        atomicAdd(steps_done, static_cast<unsigned long long>(it) * n);  // one trial energy per iteration
    }
}

}  // namespace eps
