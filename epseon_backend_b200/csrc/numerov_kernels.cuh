// numerov_kernels.cuh -- sm_100a device code of the Numerov hot path.
//
// Fills the compute slot the reference leaves empty in
// VibwaAlgorithm<FP>::run (cpp/gpu/include/epseon/gpu/algorithms/vibwa.hpp:605-637):
// the reference allocates "5 device arrays of potential_buffer_size + level_count
// outputs per work item" (algorithm_config.hpp:177-190) and dispatches nothing.
// The numerical specification is DESIGN.md section 3 (the build's own; the
// reference states none) and is mirrored operation-for-operation by
// oracle/numerov_oracle.c.
//
// Arithmetic contract: every FP64 operation is an explicit round-to-nearest
// intrinsic (__dadd_rn/__dsub_rn/__dmul_rn/__fma_rn), which nvcc never
// contracts or re-associates, so node counts AND tails are bit-identical to
// the oracle.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace eps {

constexpr int      kTile          = 1024;  // grid steps per shared-memory stage (16 KiB of (A,B))
constexpr int      kStages        = 4;     // TMA ring depth
constexpr int      kConsumerWarps = 8;     // 256 trial energies per CTA
constexpr int      kSweepThreads  = (kConsumerWarps + 1) * 32;  // + 1 TMA producer warp
constexpr int      kEnergiesPerCta = kConsumerWarps * 32;
constexpr int      kRenorm        = 128;   // exponent renormalisation period (steps)
constexpr uint32_t kNone          = 0xffffffffu;

// One potential curve resident in HBM: its (A_k, B_k) coefficient pairs live at
// AB[ab_off .. ab_off + slot), slot a multiple of kTile, padded with (2, 1).
struct CurveDev {
    uint64_t ab_off;   // in double2 units
    uint32_t n_steps;  // recurrence steps
    uint32_t i0;
    double   s;        // energy scale
    double   v_min;
};

// One row of trial energies E_j = E0 + (j0 + j) * dE, j = 0..nE-1, on one curve
// (or explicit energies at Eexp[e_off + j]).  `slot` links a refinement row back
// to its (curve, level) bracket.
struct alignas(16) Job {
    double   E0;
    double   dE;
    uint64_t e_off;
    uint32_t curve;
    uint32_t j0;
    uint32_t nE;
    uint32_t level;
    uint32_t slot;
    uint32_t pad;
};

// ---------------------------------------------------------------------------
// mbarrier / TMA (cp.async.bulk) primitives -- inline PTX, sm_90+/sm_100a.
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// 1-D bulk copy global -> shared, completion signalled on an mbarrier (TMA
// engine; SASS: UBLKCP).  dst/src 16-byte aligned, bytes a multiple of 16.
__device__ __forceinline__ void tma_bulk_g2s(void* dst, const void* src, uint32_t bytes,
                                             uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
            "r"(smem_u32(dst)),
        "l"(src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

// ---------------------------------------------------------------------------
// Numerov recurrence pieces (spec: DESIGN.md section 3; oracle: sweep_block()).
// ---------------------------------------------------------------------------
struct Chain {
    double Y, Yp, fp;  // Y_k, Y_{k-1}, f_{k-1}
};

// One step: u = A - 10e, f = B + e, g = f f_prev, t = g Y_prev, Y' = fma(u, Y, -t).
// 5 FP64 instructions (2 DADD, 2 DMUL, 1 DFMA) -- counted as 7 FLOP.
__device__ __forceinline__ void numerov_step(Chain& c, const double2 ab, const double e,
                                             const double e10) {
    const double u  = __dsub_rn(ab.x, e10);
    const double f  = __dadd_rn(ab.y, e);
    const double g  = __dmul_rn(f, c.fp);
    const double t  = __dmul_rn(g, c.Yp);
    const double Yn = __fma_rn(u, c.Y, -t);
    c.Yp = c.Y;
    c.Y  = Yn;
    c.fp = f;
}

// Scale (Y, Yp) by the power of two that brings |Y| into [1,2); exact.
__device__ __forceinline__ void renorm(Chain& c, int& expo) {
    const uint32_t ex = (static_cast<uint32_t>(__double2hiint(c.Y)) >> 20) & 0x7ffu;
    if (ex != 0) {
        const double sc = __hiloint2double(static_cast<int>((2046u - ex) << 20), 0);
        c.Y  = __dmul_rn(c.Y, sc);
        c.Yp = __dmul_rn(c.Yp, sc);
        expo += static_cast<int>(ex) - 1023;
    }
}

// ---------------------------------------------------------------------------
// Many-energy sweep: one FP64 recurrence per thread, the curve's (A,B) table
// streamed through a kStages-deep shared-memory ring by a TMA producer warp and
// read by every consumer thread as a warp-broadcast LDS.128.
//   grid  = n_jobs * chunks_per_job CTAs,  CTA = 8 consumer warps + 1 producer
//   smem  = kStages * kTile * 16 B ring + 2*kStages mbarriers
// ---------------------------------------------------------------------------
template <bool kTails>
__global__ void __launch_bounds__(kSweepThreads, 2)
numerov_sweep_kernel(const double2* __restrict__ AB, const CurveDev* __restrict__ curves,
                     const Job* __restrict__ jobs, const uint32_t chunks_per_job,
                     const double* __restrict__ Eexp, const uint64_t out_stride,
                     uint32_t* __restrict__ nodes_out, double* __restrict__ mant_out,
                     int32_t* __restrict__ exp_out, unsigned long long* __restrict__ steps_done) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double2*  ring  = reinterpret_cast<double2*>(smem_raw);
    uint64_t* full  = reinterpret_cast<uint64_t*>(smem_raw + sizeof(double2) * kTile * kStages);
    uint64_t* empty = full + kStages;

    const uint32_t job_idx = blockIdx.x / chunks_per_job;
    const uint32_t chunk   = blockIdx.x - job_idx * chunks_per_job;
    const Job      job     = jobs[job_idx];
    const CurveDev cv      = curves[job.curve];
    const uint32_t n_steps = cv.n_steps;
    const uint32_t n_tiles = (n_steps + kTile - 1) / kTile;
    const uint32_t warp    = threadIdx.x >> 5;
    const uint32_t lane    = threadIdx.x & 31;

    const uint32_t e_base = chunk * kEnergiesPerCta;
    if (e_base >= job.nE) return;  // whole CTA past the end of the row (uniform)

    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; s++) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], kConsumerWarps);
        }
        mbar_fence_init();
    }
    __syncthreads();

    if (warp == kConsumerWarps) {
        // ===== TMA producer: one elected lane streams the curve through the ring =====
        if (lane == 0) {
            const double2* src = AB + cv.ab_off;
            for (uint32_t t = 0; t < n_tiles; t++) {
                const uint32_t s = t % kStages;
                if (t >= kStages) mbar_wait(&empty[s], ((t / kStages) - 1) & 1);
                mbar_arrive_expect_tx(&full[s], kTile * sizeof(double2));
                tma_bulk_g2s(ring + s * kTile, src + static_cast<uint64_t>(t) * kTile,
                             kTile * sizeof(double2), &full[s]);
            }
            const uint32_t in_cta = min(job.nE - e_base, static_cast<uint32_t>(kEnergiesPerCta));
            atomicAdd(steps_done, static_cast<unsigned long long>(n_steps) * in_cta);
        }
        return;
    }

    // ===== consumers: one trial energy per thread =====
    uint32_t j = e_base + warp * 32 + lane;
    const bool live = j < job.nE;
    if (!live) j = job.nE - 1;  // keep the warp converged; result discarded
    double E;
    if (Eexp != nullptr) E = Eexp[job.e_off + j];
    else E = __dadd_rn(job.E0, __dmul_rn(__ull2double_rn(static_cast<unsigned long long>(job.j0) + j), job.dE));
    const double e   = __dmul_rn(cv.s, E);
    const double e10 = __dmul_rn(10.0, e);

    Chain    c{1.0, 0.0, 1.0};
    int      expo     = 0;
    uint32_t n_nodes  = 0;
    uint32_t prevmask = 0;  // bit0 = sign of the latest Y of the previous 32-step group

    for (uint32_t t = 0; t < n_tiles; t++) {
        const uint32_t s = t % kStages;
        mbar_wait(&full[s], (t / kStages) & 1);
        const double2* __restrict__ tile = ring + s * kTile;
        const uint32_t n_valid = min(static_cast<uint32_t>(kTile), n_steps - t * kTile);
        const uint32_t n_full  = n_valid / kRenorm;

        uint32_t k = 0;
#pragma unroll 1
        for (uint32_t r = 0; r < n_full; r++) {
#pragma unroll 1
            for (int q = 0; q < kRenorm / 32; q++) {
                uint32_t mask = 0;
#pragma unroll
                for (int i = 0; i < 32; i++) {
                    numerov_step(c, tile[k + i], e, e10);
                    mask = __funnelshift_l(static_cast<uint32_t>(__double2hiint(c.Y)), mask, 1);
                }
                n_nodes += __popc(mask ^ __funnelshift_r(mask, prevmask, 1));
                prevmask = mask;
                k += 32;
            }
            renorm(c, expo);
        }
        // ragged tail of the last tile (< kRenorm steps): plain per-step counting
        for (; k < n_valid; k++) {
            const uint32_t before = static_cast<uint32_t>(__double2hiint(c.Y));
            numerov_step(c, tile[k], e, e10);
            const uint32_t after = static_cast<uint32_t>(__double2hiint(c.Y));
            n_nodes += (before ^ after) >> 31;
            prevmask = after >> 31;
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[s]);
    }
    renorm(c, expo);

    if (live) {
        const uint64_t o = static_cast<uint64_t>(job_idx) * out_stride + j;
        nodes_out[o] = n_nodes;
        if (kTails) {
            mant_out[o] = c.Y;
            exp_out[o]  = expo;
        }
    }
}

// ---------------------------------------------------------------------------
// Bracketing helpers (N5/N6).  All decisions are integer comparisons of node
// counts, so they reproduce the oracle's orc_solve_levels exactly.
// ---------------------------------------------------------------------------

// Row c of the coarse sweep: E_j = E_lo[c] + j*dE.
__global__ void make_coarse_jobs_kernel(const double* __restrict__ E_lo,
                                        const double* __restrict__ E_hi, uint32_t n_curves,
                                        uint32_t nE, Job* __restrict__ jobs) {
    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_curves) return;
    Job jb;
    jb.E0    = E_lo[c];
    jb.dE    = nE > 1 ? __ddiv_rn(__dsub_rn(E_hi[c], E_lo[c]), static_cast<double>(nE - 1)) : 0.0;
    jb.e_off = static_cast<uint64_t>(c) * nE;
    jb.curve = c;
    jb.j0    = 0;
    jb.nE    = nE;
    jb.level = 0;
    jb.slot  = c;
    jb.pad   = 0;
    jobs[c]  = jb;
}

// First index j of each row with nodes[j] > v, for every wanted level v of the
// row: adjacent energies are compared with a warp shuffle, warps without any
// increase leave on one ballot, the rest atomicMin their candidate index.
__global__ void crossing_kernel(const uint32_t* __restrict__ nodes, uint64_t stride, uint32_t nE,
                                uint32_t blocks_per_row, const Job* __restrict__ jobs, int coarse,
                                uint32_t v_first_all, uint32_t v_count_all,
                                uint32_t* __restrict__ jstar) {
    const uint32_t  row  = blockIdx.x / blocks_per_row;
    const uint32_t  j    = (blockIdx.x - row * blocks_per_row) * blockDim.x + threadIdx.x;
    const uint32_t* r    = nodes + static_cast<uint64_t>(row) * stride;
    const uint32_t  lane = threadIdx.x & 31;
    const uint32_t  b    = (j < nE) ? r[j] : 0u;
    uint32_t        a    = __shfl_up_sync(0xffffffffu, b, 1);
    if (lane == 0) a = (j == 0 || j >= nE) ? 0u : r[j - 1];
    const bool up = (j < nE) && (b > a);
    if (__ballot_sync(0xffffffffu, up) == 0u) return;
    if (up) {
        const uint32_t vf = coarse ? v_first_all : jobs[row].level;
        const uint32_t vc = coarse ? v_count_all : 1u;
        const uint32_t ob = coarse ? row * v_count_all : row;
        const uint32_t v0 = a > vf ? a : vf;
        const uint32_t v1 = (b - 1 < vf + vc - 1) ? b - 1 : vf + vc - 1;
        for (uint32_t v = v0; v <= v1 && v1 != kNone; v++) atomicMin(&jstar[ob + (v - vf)], j);
    }
}

// state: 0 absent, 1 active, 2 converged/frozen
__global__ void bracket_init_kernel(const uint32_t* __restrict__ nodes, uint64_t stride,
                                    const Job* __restrict__ jobs, const uint32_t* __restrict__ jstar,
                                    uint32_t n_curves, uint32_t v_min, uint32_t n_lev,
                                    double* __restrict__ lo, double* __restrict__ hi,
                                    uint32_t* __restrict__ state, uint32_t* __restrict__ n_below) {
    const uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n_curves * n_lev) return;
    const uint32_t c = idx / n_lev, l = idx - c * n_lev, v = v_min + l;
    const Job      jb    = jobs[c];
    const uint32_t first = nodes[static_cast<uint64_t>(c) * stride];
    const uint32_t last  = nodes[static_cast<uint64_t>(c) * stride + jb.nE - 1];
    if (l == 0) n_below[c] = last;
    const uint32_t j = jstar[idx];
    if (last <= v || first > v || j == kNone || j == 0) {
        lo[idx]    = __longlong_as_double(0x7ff8000000000000LL);
        hi[idx]    = __longlong_as_double(0x7ff8000000000000LL);
        state[idx] = 0;
        return;
    }
    lo[idx]    = __dadd_rn(jb.E0, __dmul_rn(static_cast<double>(j - 1), jb.dE));
    hi[idx]    = __dadd_rn(jb.E0, __dmul_rn(static_cast<double>(j), jb.dE));
    state[idx] = 1;
}

// Convergence test + ordered compaction of the still-active brackets into the
// refinement job list (single CTA; a few thousand entries at most).
__global__ void check_compact_kernel(const double* __restrict__ lo, const double* __restrict__ hi,
                                     uint32_t* __restrict__ state, uint32_t total, uint32_t n_lev,
                                     uint32_t v_min, double rel_tol, uint32_t M,
                                     Job* __restrict__ jobs_out, uint32_t* __restrict__ n_active) {
    __shared__ uint32_t warp_cnt[32];
    __shared__ uint32_t running;
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) running = 0;
    __syncthreads();
    for (uint32_t base = 0; base < total; base += blockDim.x) {
        const uint32_t idx  = base + threadIdx.x;
        bool           flag = false;
        double         l = 0, h = 0;
        if (idx < total && state[idx] == 1) {
            l = lo[idx];
            h = hi[idx];
            const double w   = __dsub_rn(h, l);
            const double mag = fmax(fabs(l), fabs(h));
            if (w <= __dmul_rn(rel_tol, mag)) state[idx] = 2;
            else flag = true;
        }
        const uint32_t bal = __ballot_sync(0xffffffffu, flag);
        if (lane == 0) warp_cnt[warp] = __popc(bal);
        __syncthreads();
        uint32_t off = running;
        for (uint32_t w = 0; w < warp; w++) off += warp_cnt[w];
        if (flag) {
            const uint32_t pos = off + __popc(bal & ((1u << lane) - 1u));
            Job            jb;
            jb.E0    = l;
            jb.dE    = __ddiv_rn(__dsub_rn(h, l), static_cast<double>(M + 1));
            jb.e_off = 0;
            jb.curve = idx / n_lev;
            jb.j0    = 1;
            jb.nE    = M;
            jb.level = v_min + (idx - (idx / n_lev) * n_lev);
            jb.slot  = idx;
            jb.pad   = 0;
            jobs_out[pos] = jb;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            uint32_t tot = 0;
            for (uint32_t w = 0; w < (blockDim.x >> 5); w++) tot += warp_cnt[w];
            running += tot;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) *n_active = running;
}

__global__ void bracket_update_kernel(const Job* __restrict__ jobs,
                                      const uint32_t* __restrict__ jstar, uint32_t n_jobs,
                                      uint32_t M, double* __restrict__ lo, double* __restrict__ hi,
                                      uint32_t* __restrict__ state) {
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_jobs) return;
    const Job      jb  = jobs[r];
    const uint32_t js  = jstar[r];
    const uint32_t m   = (js == kNone) ? M + 1 : js + 1;
    const double   lo0 = jb.E0, hi0 = hi[jb.slot];
    const double   nlo = (m == 1) ? lo0 : __dadd_rn(lo0, __dmul_rn(static_cast<double>(m - 1), jb.dE));
    const double   nhi = (m == M + 1) ? hi0 : __dadd_rn(lo0, __dmul_rn(static_cast<double>(m), jb.dE));
    if (!(__dsub_rn(nhi, nlo) < __dsub_rn(hi0, lo0))) state[jb.slot] = 2;
    lo[jb.slot] = nlo;
    hi[jb.slot] = nhi;
}

__global__ void finalize_levels_kernel(const double* __restrict__ lo, const double* __restrict__ hi,
                                       uint32_t total, double* __restrict__ levels,
                                       double* __restrict__ widths) {
    const uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    levels[idx] = __dmul_rn(0.5, __dadd_rn(lo[idx], hi[idx]));
    widths[idx] = __dsub_rn(hi[idx], lo[idx]);
}

// ---------------------------------------------------------------------------
// DFMA-saturating probe: measures the FP64 (non-tensor) roofline denominator,
// which MEASURED_PEAKS.json does not carry.  8 independent chains per thread.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) fp64_probe_kernel(double* __restrict__ out, int iters,
                                                         double x, double y) {
    double a0 = threadIdx.x * 1e-3, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5,
           a6 = a0 + 6, a7 = a0 + 7;
#pragma unroll 1
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int u = 0; u < 8; u++) {
            a0 = __fma_rn(a0, x, y);
            a1 = __fma_rn(a1, x, y);
            a2 = __fma_rn(a2, x, y);
            a3 = __fma_rn(a3, x, y);
            a4 = __fma_rn(a4, x, y);
            a5 = __fma_rn(a5, x, y);
            a6 = __fma_rn(a6, x, y);
            a7 = __fma_rn(a7, x, y);
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
}

}  // namespace eps
