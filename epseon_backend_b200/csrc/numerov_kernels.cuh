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
#ifndef EPS_STEP_NEGATED
#define EPS_STEP_NEGATED 1
#endif
#include <climits>
#include <cstdint>
#include <cuda_runtime.h>

namespace eps {

constexpr int      kTile          = 2048;  // grid steps per shared-memory stage (16 KiB of F_k)
constexpr int      kStages        = 4;     // TMA ring depth of the big CTA shapes
// Ring depth by CTA shape: the 128-energy CTAs (1 chain x 4 warps, 2 chains x 2 warps) run four to a SM, so two
// stages (32 KiB) per CTA keep as much table in flight per SM as one 4-stage CTA does.
template <int kEpt, int kWarps>
__host__ __device__ constexpr int sweep_stages() {
    return (kEpt * kWarps <= 4) ? 2 : kStages;
}
// Resident CTAs per SM the launch bounds ask for: 1 (512 energies), 2 (256), 4 (128).
template <int kEpt, int kWarps>
__host__ __device__ constexpr int sweep_min_ctas() {
    return (kEpt * kWarps >= 32) ? 1 : ((kEpt * kWarps <= 4) ? 4 : 2);
}
constexpr uint32_t kProducerSuspendNs = 1000000;  // try_wait suspend-time hint of the TMA producer
// CTA shape of the sweep: kWarps consumer warps (template parameter) + 1 TMA producer warp,
// 32 * kWarps * kEpt trial energies per CTA.
constexpr int      kRenorm        = 128;   // exponent renormalisation period (steps)
constexpr uint32_t kNone          = 0xffffffffu;
constexpr uint32_t kFlatRows      = 31;    // pack_log2 value selecting the flat-row mapping

// One potential curve resident in HBM: its coefficient table F_k = (1 - q_k)/12
// lives at F[f_off .. f_off + slot), slot a multiple of kTile, padded with 1/12.
struct CurveDev {
    uint64_t f_off;    // in doubles
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
// Producer-side wait.  The (otherwise idle) producer warp shares a scheduler with two consumer
// warps, so it must not spin: try_wait carries a suspend-time hint, i.e. the hardware parks the
// thread until the phase completes or the hint (ns) expires -- no issue slots are taken meanwhile.
// (profiles/r1c_ncu.md: a NANOSLEEP/try_wait/BRA poll loop still executed 26 M times per launch,
// 7.5 % of all warp instructions, all on the producer's scheduler.)
__device__ __forceinline__ void mbar_wait_backoff(uint64_t* bar, uint32_t parity) {
    uint32_t done = 0;
    while (true) {
        asm volatile(
            "{\n"
            ".reg .pred P1;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2, %3;\n"
            "selp.u32 %0, 1, 0, P1;\n"
            "}"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity), "r"(kProducerSuspendNs)
            : "memory");
        if (done) break;
    }
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
// Numerov recurrence pieces (spec: DESIGN.md section 3.3; oracle: sweep_block()).
// ---------------------------------------------------------------------------
struct Chain {
    double X, S;  // X_k, S_{k-1} = (f_{k-1}/12) X_{k-1}
};

// One step of the 4-operation X form (1 DADD, 1 DMUL, 2 DFMA = 6 FLOP):
//   fp = F_k + e/12;  Q = fma(10, X, S);  X' = fma(-fp, Q, X);  S' = fp * X.
// The ORDER OF THE STATEMENTS matters to ptxas, not to the arithmetic.  A DFMA with three distinct
// register operands holds the FP64 pipe for 3 cycles instead of 2 (one 64-bit register read per
// cycle; scripts/microbench3.cu) unless one operand comes from the operand reuse cache, i.e. was
// read in the same operand slot by the warp's previous instruction.  ptxas orders the operands of
// commutative instructions by the age of their virtual registers: with Q defined BEFORE fp the
// pair becomes  DMUL S', X, fp.reuse ; DFMA X', Q, fp, X  -- fp in slot B of both -- and the DFMA
// reads two registers (profiles/r1_microbench4_order.log; SASS check: scripts/sass_reuse.py).
__device__ __forceinline__ void numerov_step(Chain& c, const double Fk, const double ep) {
    const double Q  = __fma_rn(10.0, c.X, c.S);
#if EPS_STEP_NEGATED
    const double fn = __dsub_rn(-Fk, ep);  // -(F_k + e/12): round-to-nearest is sign-symmetric, same bits
    const double Sn = -__dmul_rn(fn, c.X);
    c.X = __fma_rn(fn, Q, c.X);
#else
    const double fp = __dadd_rn(Fk, ep);
    const double Sn = __dmul_rn(fp, c.X);
    c.X = __fma_rn(-fp, Q, c.X);
#endif
    c.S = Sn;
}

// The ACCURATE recurrence, "D form" (spec DESIGN.md section 3.3; oracle sweep_block, form 1): the
// chain carries Y_k (in Chain::X) and the scaled first difference D_{k-1} (in Chain::S); the table
// holds A_k = 12 q_k and the per-energy constant is 12 e.  5 operations, 8 FLOP per step:
//   T12 = A_k - 12e;  f = fma(-1/12, T12, 1);  t = T12 * Y;  D = fma(f, D, t);  Y' = fma(f, Y, D).
// Carrying the difference keeps one step's rounding out of the SLOPE of the solution, which every
// value-carrying form amplifies by 1/sqrt(12 T): eigenvalues agree with the binary128 solution of
// the discrete problem to ~1e-14 instead of 1e-9 (tests/test_accuracy_floor.py).
__device__ __forceinline__ void numerov_step_d(Chain& c, const double Ak, const double e12) {
    const double T12 = __dsub_rn(Ak, e12);
    const double t   = __dmul_rn(T12, c.X);
    const double f   = __fma_rn(-(1.0 / 12.0), T12, 1.0);
    c.S = __fma_rn(f, c.S, t);
    c.X = __fma_rn(f, c.X, c.S);
}

template <int kForm>
__device__ __forceinline__ void step_form(Chain& c, const double Tk, const double ec) {
    if constexpr (kForm == 0) numerov_step(c, Tk, ec);
    else numerov_step_d(c, Tk, ec);
}

// Per-energy constant of a form: e/12 = (s E)/12 (X form) or 12 e = 12 (s E) (D form).
template <int kForm>
__device__ __forceinline__ double energy_const(const double s, const double E) {
    if constexpr (kForm == 0) return __ddiv_rn(__dmul_rn(s, E), 12.0);
    else return __dmul_rn(12.0, __dmul_rn(s, E));
}

// Scale (X, S) by the power of two that brings |X| into [1,2); exact.
__device__ __forceinline__ void renorm(Chain& c, int& expo) {
    const uint32_t ex = (static_cast<uint32_t>(__double2hiint(c.X)) >> 20) & 0x7ffu;
    if (ex != 0) {
        const double sc = __hiloint2double(static_cast<int>((2046u - ex) << 20), 0);
        c.X = __dmul_rn(c.X, sc);
        c.S = __dmul_rn(c.S, sc);
        expo += static_cast<int>(ex) - 1023;
    }
}

// ---------------------------------------------------------------------------
// Many-energy sweep: kEpt independent FP64 recurrences per thread, the curve's
// F table streamed through a kStages-deep shared-memory ring by a TMA producer
// warp and read by every consumer thread as a warp-broadcast LDS.128 (2 steps).
//   grid  = n_jobs * chunks_per_job CTAs,  CTA = kWarps consumer warps + 1 producer
//   smem  = kStages * kTile * 8 B ring + 2*kStages mbarriers
// Energy j of a CTA's chunk sits in thread (j % (32 kWarps)), chain (j / (32 kWarps)), so the
// result stores of every chain are coalesced.
//
// Node counting.  An integer SHF per step between the FP64 instructions costs 16 % of the
// DFMA rate (it takes issue cycles, and it separates the DMUL/DFMA pairs that share an
// operand through the reuse cache, see numerov_step), so the sign of X is sampled every
// kStride steps.  That is EXACTLY the per-step
// sign-flip count of the spec whenever two zeros of a solution are more than
// kStride steps apart, which Sturm separation guarantees for
//     kStride * theta_max < pi,   theta_max^2 = 12 * s * (E_max - V_min);
// the host only selects kStride = 32 / 8 with a factor-2 margin on theta_max
// (launch_sweep) and kStride = 1 (per-step bit mask) otherwise.
//
// Row packing (pack_log2 = g > 0, sequential mode): rows shorter than the CTA (refinement rounds
// with few points per level) are packed 2^g to a CTA, every row in a slot of kPerCta >> g
// energies; the caller lays the rows out so that the 2^g rows of a CTA share one curve
// (jobs[] dense by [curve][level], level count padded to a multiple of 2^g, nE = 0 for idle
// rows), so the CTA still streams ONE table.  chunks_per_job is 1 in that mode.
// Flat rows (pack_log2 == kFlatRows, launches whose rows all sit on ONE curve and all have nE
// energies): the CTA takes kPerCta consecutive energies of the flattened (row, j) index space, so
// rows may straddle CTAs and every CTA but the last is full -- 17 rows x 4457 points fill 148 CTAs.
//
// kScan (transfer-matrix mode, N4): the grid is cut into n_seg segments of whole tiles and a
// CTA marches ONE segment for kWarps*32 energies; each thread carries two basis solutions of
// its energy (kEpt == 2, shared fp): A from (X,S) = (1,1) ("flat": psi_a = psi_{a-1}) and B from
// (0,1) ("unit slope"), plus the sign-flip count of A.  The segment's transfer matrix is stored in
// the coordinates (X, D = S - X): consecutive values of a smooth solution are nearly equal, so
// the naive (X,S) columns (1,0)/(0,1) would make every combination a cancellation of two
// columns ~n_steps/n_seg times larger than the result (measured: 150x larger tail error).
// segment_combine_kernel chains the segments.  Results go to SegOut instead of nodes/tails.
// ---------------------------------------------------------------------------
struct SegOut {
    double*   XA;  // [row][seg][energy] column A = P (1,1)^T: X and D = S - X, exponent eA
    double*   SA;  //   (holds D_A)
    double*   XB;  // column B = P (0,1)^T
    double*   SB;  //   (holds D_B)
    int32_t*  eA;
    int32_t*  eB;
    uint32_t* nA;  // sign flips of X along column A
};

template <int kEpt, int kWarps, int kStride, bool kTails, bool kScan, int kForm>
__global__ void __launch_bounds__((kWarps + 1) * 32, sweep_min_ctas<kEpt, kWarps>())
numerov_sweep_kernel(const double* __restrict__ F, const CurveDev* __restrict__ curves,
                     const Job* __restrict__ jobs, const uint32_t chunks_per_job,
                     const double* __restrict__ Eexp, const uint64_t out_stride,
                     uint32_t* __restrict__ nodes_out, double* __restrict__ mant_out,
                     int32_t* __restrict__ exp_out, unsigned long long* __restrict__ steps_done,
                     const uint32_t n_seg, const uint32_t tiles_per_seg, const SegOut seg_out,
                     const uint32_t pack_log2, const int* __restrict__ stop_flag,
                     const uint32_t* __restrict__ flat_rows_dev) {
    static_assert(kStride == 1 || kStride == 8 || kStride == 32, "sign sampling stride");
    static_assert(!kScan || (kEpt == 2 && !kTails), "scan mode: two basis chains per energy");
    constexpr uint32_t kPerCta = kScan ? kWarps * 32 : kWarps * 32 * kEpt;
    constexpr int      kCnt    = kScan ? 1 : kEpt;  // chains whose sign flips are counted
    constexpr int      kSt     = sweep_stages<kEpt, kWarps>();  // ring depth of this shape
    if (*stop_flag != 0) return;  // eps_request_stop: queued sweeps drain without marching (uniform per CTA)
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double*   ring  = reinterpret_cast<double*>(smem_raw);
    uint64_t* full  = reinterpret_cast<uint64_t*>(smem_raw + sizeof(double) * kTile * kSt);
    uint64_t* empty = full + kSt;

    const uint32_t seg     = kScan ? blockIdx.x % n_seg : 0u;
    const uint32_t cta     = kScan ? blockIdx.x / n_seg : blockIdx.x;
    const bool     flat    = !kScan && pack_log2 == kFlatRows;
    const uint32_t job_idx = flat ? 0u : (cta / chunks_per_job) << pack_log2;  // first row of this CTA
    const uint32_t chunk   = flat ? 0u : cta - (cta / chunks_per_job) * chunks_per_job;
    const Job      job     = jobs[job_idx];
    const CurveDev cv      = curves[job.curve];
    const uint32_t slot_sz = flat ? job.nE : kPerCta >> pack_log2;  // energies per row (flat) / packed row slot
    const uint64_t flat0   = static_cast<uint64_t>(cta) * kPerCta;  // flat mode: first (row * nE + j) of this CTA
    // flat mode: chunks_per_job carries the number of rows -- or an upper bound of it, with the true count on
    // the device (flat_rows_dev): refinement rounds enqueued before the host knows how many brackets are open
    const uint64_t flat_n  = static_cast<uint64_t>((flat && flat_rows_dev != nullptr) ? *flat_rows_dev : chunks_per_job) * job.nE;
    const uint32_t n_steps = cv.n_steps;
    const uint32_t n_tiles_all = (n_steps + kTile - 1) / kTile;
    const uint32_t t_begin = kScan ? min(seg * tiles_per_seg, n_tiles_all) : 0u;
    const uint32_t t_end   = kScan ? min(t_begin + tiles_per_seg, n_tiles_all) : n_tiles_all;
    const uint32_t n_tiles = t_end - t_begin;  // tiles this CTA marches (may be 0: identity segment)
    const uint32_t warp    = threadIdx.x >> 5;
    const uint32_t lane    = threadIdx.x & 31;

    const uint32_t e_base = chunk * kPerCta;
    uint32_t       cta_energies;  // valid trial energies of this CTA
    if (flat) {
        if (flat0 >= flat_n) return;
        cta_energies = static_cast<uint32_t>(min(static_cast<uint64_t>(kPerCta), flat_n - flat0));
    } else if (pack_log2 == 0) {
        if (e_base >= job.nE) return;  // whole CTA past the end of the row (uniform)
        cta_energies = min(job.nE - e_base, kPerCta);
    } else {
        cta_energies = 0;
        for (uint32_t r = 0; r < (1u << pack_log2); r++) cta_energies += min(jobs[job_idx + r].nE, slot_sz);
        if (cta_energies == 0) return;  // every packed row idle (uniform)
    }

    if (threadIdx.x == 0) {
        for (int s = 0; s < kSt; s++) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], kWarps * 32);  // every consumer THREAD releases the stage (see the arrive below)
        }
        mbar_fence_init();
    }
    __syncthreads();

    if (warp == kWarps) {
        // ===== TMA producer: one elected lane streams the curve through the ring =====
        if (lane == 0) {
            const double* src = F + cv.f_off;
            for (uint32_t t = 0; t < n_tiles; t++) {
                const uint32_t s = t % kSt;
                if (t >= kSt) mbar_wait_backoff(&empty[s], ((t / kSt) - 1) & 1);
                mbar_arrive_expect_tx(&full[s], kTile * sizeof(double));
                tma_bulk_g2s(ring + s * kTile, src + static_cast<uint64_t>(t_begin + t) * kTile,
                             kTile * sizeof(double), &full[s]);
            }
            const uint32_t in_cta   = cta_energies;
            const uint32_t my_steps = min(t_end * kTile, n_steps) - min(t_begin * kTile, n_steps);
            atomicAdd(steps_done, static_cast<unsigned long long>(my_steps) * in_cta);
        }
        return;
    }

    // ===== consumers: kEpt trial energies per thread =====
    Chain    c[kEpt];
    double   ep[kEpt];
    int      expo[kEpt];
    uint32_t n_nodes[kEpt], prev[kEpt];
#pragma unroll
    for (int i = 0; i < kEpt; i++) {
        const uint32_t t  = (kScan ? 0 : i) * (kWarps * 32) + warp * 32 + lane;  // energy slot in the CTA
        uint32_t row_i = job_idx, j = e_base + t;
        if (flat) {
            const uint64_t f = min(flat0 + t, flat_n - 1);  // clamp: keeps the warp converged; result discarded
            row_i            = static_cast<uint32_t>(f / slot_sz);
            j                = static_cast<uint32_t>(f - static_cast<uint64_t>(row_i) * slot_sz);
        } else if (pack_log2) {
            row_i = job_idx + t / slot_sz;
            j     = t % slot_sz;
        }
        const Job jb = (flat || pack_log2) ? jobs[row_i] : job;
        if (j >= jb.nE) j = jb.nE ? jb.nE - 1 : 0;  // keep the warp converged; result discarded
        double E;
        if (Eexp != nullptr) E = Eexp[jb.e_off + j];
        else E = __dadd_rn(jb.E0, __dmul_rn(__ull2double_rn(static_cast<unsigned long long>(jb.j0) + j), jb.dE));
        ep[i]      = energy_const<kForm>(cv.s, E);
        // start: psi = 0 one point to the left.  X form (X, S) = (1, 0); D form (Y, D) = (1, 1).
        // Scan basis: "flat" / "unit slope" = (X, D) = (1, 0) / (0, 1), i.e. (X, S) = (1, 1) / (0, 1) in the X form.
        if (kForm == 0) c[i] = kScan ? (i == 1 ? Chain{0.0, 1.0} : Chain{1.0, 1.0}) : Chain{1.0, 0.0};
        else c[i] = kScan ? (i == 1 ? Chain{0.0, 1.0} : Chain{1.0, 0.0}) : Chain{1.0, 1.0};
        expo[i]    = 0;
        n_nodes[i] = 0;
        prev[i]    = 0;  // kStride==1: bit0 = sign of the last X; else: hi word of the last sampled X
    }

    for (uint32_t t = 0; t < n_tiles; t++) {
        const uint32_t s = t % kSt;
        mbar_wait(&full[s], (t / kSt) & 1);
        const double* __restrict__ tile = ring + s * kTile;
        const uint32_t n_valid = min(static_cast<uint32_t>(kTile), n_steps - (t_begin + t) * kTile);
        const uint32_t n_full  = n_valid / kRenorm;

        uint32_t k = 0;
#pragma unroll 1
        for (uint32_t r = 0; r < n_full; r++) {
#pragma unroll 1
            for (int q = 0; q < kRenorm / 32; q++) {
                uint32_t mask[kEpt];
#pragma unroll
                for (int i = 0; i < kEpt; i++) mask[i] = 0;
                const double2* __restrict__ t2 = reinterpret_cast<const double2*>(tile + k);
#pragma unroll
                for (int p = 0; p < 16; p++) {
                    const double2 ff = t2[p];  // two consecutive grid steps, warp-broadcast
                    if (kScan) {  // the two basis chains share fp: both steps of a chain back to back
#pragma unroll              // (30 instead of 3 of the 83 three-register DFMAs then take fp from the reuse cache)
                        for (int i = 0; i < kEpt; i++) {
                            step_form<kForm>(c[i], ff.x, ep[0]);
                            if (kStride == 1 && i < kCnt) mask[i] = __funnelshift_l(static_cast<uint32_t>(__double2hiint(c[i].X)), mask[i], 1);
                            step_form<kForm>(c[i], ff.y, ep[0]);
                            if (kStride == 1 && i < kCnt) mask[i] = __funnelshift_l(static_cast<uint32_t>(__double2hiint(c[i].X)), mask[i], 1);
                        }
                    } else {
#pragma unroll
                        for (int i = kEpt - 1; i >= 0; i--) {
                            step_form<kForm>(c[i], ff.x, ep[i]);
                            if (kStride == 1 && i < kCnt) mask[i] = __funnelshift_l(static_cast<uint32_t>(__double2hiint(c[i].X)), mask[i], 1);
                        }
#pragma unroll
                        for (int i = 0; i < kEpt; i++) {
                            step_form<kForm>(c[i], ff.y, ep[i]);
                            if (kStride == 1 && i < kCnt) mask[i] = __funnelshift_l(static_cast<uint32_t>(__double2hiint(c[i].X)), mask[i], 1);
                        }
                    }
                    if (kStride == 8 && (p & 3) == 3) {
#pragma unroll
                        for (int i = 0; i < kCnt; i++) {
                            const uint32_t cur = static_cast<uint32_t>(__double2hiint(c[i].X));
                            n_nodes[i] += (cur ^ prev[i]) >> 31;
                            prev[i] = cur;
                        }
                    }
                }
#pragma unroll
                for (int i = 0; i < kCnt; i++) {
                    if (kStride == 1) {
                        n_nodes[i] += __popc(mask[i] ^ __funnelshift_r(mask[i], prev[i], 1));
                        prev[i] = mask[i];
                    } else if (kStride == 32) {
                        const uint32_t cur = static_cast<uint32_t>(__double2hiint(c[i].X));
                        n_nodes[i] += (cur ^ prev[i]) >> 31;
                        prev[i] = cur;
                    }
                }
                k += 32;
            }
#pragma unroll
            for (int i = 0; i < kEpt; i++) renorm(c[i], expo[i]);
        }
        // ragged tail of the last tile (< kRenorm steps): plain per-step counting
        for (; k < n_valid; k++) {
            const double Fk = tile[k];
#pragma unroll
            for (int i = 0; i < kEpt; i++) {
                const uint32_t before = static_cast<uint32_t>(__double2hiint(c[i].X));
                step_form<kForm>(c[i], Fk, ep[kScan ? 0 : i]);
                const uint32_t after = static_cast<uint32_t>(__double2hiint(c[i].X));
                n_nodes[i] += (before ^ after) >> 31;
                prev[i] = (kStride == 1) ? (after >> 31) : after;
            }
        }
        // Every consumer thread arrives for itself (32 arrivals per warp per 2048-step tile: noise).  One
        // elected lane behind a __syncwarp() is equally ordered by the memory model, but compute-sanitizer's
        // racecheck does not follow the order through the warp barrier and reports the producer's next TMA
        // write into the stage against the other lanes' reads (scripts/sanitize_ring.py: clean this way).
        mbar_arrive(&empty[s]);
    }

    if constexpr (kScan) {
        renorm(c[0], expo[0]);
        renorm(c[1], expo[1]);
        const uint32_t j = e_base + warp * 32 + lane;
        if (j < job.nE) {
            const uint64_t o = (static_cast<uint64_t>(job_idx) * n_seg + seg) * out_stride + j;
            seg_out.XA[o] = c[0].X;
            seg_out.SA[o] = kForm == 0 ? __dsub_rn(c[0].S, c[0].X) : c[0].S;  // D form: S already holds D
            seg_out.XB[o] = c[1].X;
            seg_out.SB[o] = kForm == 0 ? __dsub_rn(c[1].S, c[1].X) : c[1].S;
            seg_out.eA[o] = expo[0];
            seg_out.eB[o] = expo[1];
            seg_out.nA[o] = n_nodes[0];
        }
    } else {
#pragma unroll
    for (int i = 0; i < kEpt; i++) {
        renorm(c[i], expo[i]);
        const uint32_t t   = i * (kWarps * 32) + warp * 32 + lane;
        uint32_t row = job_idx, j = e_base + t, lim = job.nE;
        if (flat) {
            const uint64_t f = flat0 + t;
            row              = static_cast<uint32_t>(f / slot_sz);
            j                = static_cast<uint32_t>(f - static_cast<uint64_t>(row) * slot_sz);
            lim              = f < flat_n ? slot_sz : 0u;
        } else if (pack_log2) {
            row = job_idx + t / slot_sz;
            j   = t % slot_sz;
            lim = min(jobs[row].nE, slot_sz);
        }
        if (j < lim) {
            const uint64_t o = static_cast<uint64_t>(row) * out_stride + j;
            nodes_out[o] = n_nodes[i];
            if (kTails) {
                mant_out[o] = c[i].X;
                exp_out[o]  = expo[i];
            }
        }
    }
    }
}

// ---------------------------------------------------------------------------
// N4 combine: chains the segment transfer matrices of one trial energy, in (X, D = S - X)
// coordinates:   v_{s+1} = P_s v_s,  v_0 = (X,S) = (1,0) i.e. (X,D) = (1,-1);
//     nodes += nA_s + sign(D_s) * (D_b - D_a),
// D_a = [X_s < 0] (column A starts at X = +1), D_b = [sign X_{s+1} != sign XA_s].  Zeros of two
// solutions of a Sturm recurrence interlace and det[A, v] = D_s keeps its sign along the
// segment, so the flip count of ANY solution differs from column A's only by the change of "v
// and A lie on opposite sides of X = 0", signed by the orientation.  No second pass over the
// grid is needed.
// Guard: rho estimates the relative error of v's direction; a combine that cancels (both
// components shrink by c = max(|X'|/mag_X, |D'|/mag_D) -- the solution decays across the
// segment) amplifies it, rho' = (rho + 2^-34) / c; for the last segment only the X component
// counts.  Energies ending with rho >= 2^-10 (E within ~1e-6 level spacings of an eigenvalue,
// where the tail is a pure cancellation) are appended to `flagged` for the sequential kernel.
// One thread per (row, energy); segments are a serial loop of n_seg 2x2 mat-vecs (n_seg <= 64:
// a parallel prefix would save nothing next to the n_steps/n_seg marches it follows).
// ---------------------------------------------------------------------------
// State of one trial energy between segments: v = (X, D) 2^ex, the running node count and the guard.
struct CombState {
    double   X, D, rho;
    int      ex;
    uint32_t nodes;
};

// One segment applied to the state (the body both combine kernels share; operation order fixed).
__device__ __forceinline__ void combine_segment(CombState& st, const SegOut& so, const uint64_t o, const bool last_seg,
                                                const double D_start) {
    constexpr double kEta = 5.820766091346741e-11;  // 2^-34
    const double   XA = so.XA[o], DA = so.SA[o];
    int            sh = so.eB[o] - so.eA[o];
    sh                = sh > 1000 ? 1000 : (sh < -1000 ? -1000 : sh);
    const double XB = scalbn(so.XB[o], sh), DB = scalbn(so.SB[o], sh);
    const double X = st.X, D = st.D;
    const double pX = __dmul_rn(XB, D), pD = __dmul_rn(DB, D);
    const double Xn = __fma_rn(XA, X, pX);
    const double Dn = __fma_rn(DA, X, pD);
    const int    fa = static_cast<int>(static_cast<uint32_t>(__double2hiint(X)) >> 31);
    const int    fb = static_cast<int>((static_cast<uint32_t>(__double2hiint(Xn)) ^
                                        static_cast<uint32_t>(__double2hiint(XA))) >> 31);
    // orientation: the sign of the discrete Wronskian of (A, v).  In the X form's coordinates D is
    // MINUS the backward difference (D = S - X), in the D form's it is PLUS (D = Y_k - f Y_{k-1}),
    // so the two forms need opposite signs; D_start = -1 / +1 carries exactly that.
    const int    sg0 = D > 0.0 ? 1 : (D < 0.0 ? -1 : 0);
    const int    sg  = D_start < 0.0 ? sg0 : -sg0;
    st.nodes += so.nA[o] + static_cast<uint32_t>(sg * (fb - fa));
    const double magX = fabs(__dmul_rn(XA, X)) + fabs(pX), magD = fabs(__dmul_rn(DA, X)) + fabs(pD);
    const double cX = magX > 0.0 ? fabs(Xn) / magX : 1.0, cD = magD > 0.0 ? fabs(Dn) / magD : 1.0;
    const double c  = last_seg ? cX : fmax(cX, cD);
    st.rho          = (st.rho + kEta) / c;  // c == 0 -> inf -> flagged
    st.X            = Xn;
    st.D            = Dn;
    st.ex += so.eA[o];
    uint32_t e11 = (static_cast<uint32_t>(__double2hiint(st.X)) >> 20) & 0x7ffu;
    if (e11 == 0) e11 = (static_cast<uint32_t>(__double2hiint(st.D)) >> 20) & 0x7ffu;
    if (e11 != 0 && !last_seg) {
        const double sc = __hiloint2double(static_cast<int>((2046u - e11) << 20), 0);
        st.X = __dmul_rn(st.X, sc);
        st.D = __dmul_rn(st.D, sc);
        st.ex += static_cast<int>(e11) - 1023;
    }
}

// Result of one energy: node count, tail in the sweep's (mantissa in [1,2), exponent) form, flag.
__device__ __forceinline__ void combine_emit(CombState st, const uint64_t o, uint32_t* __restrict__ nodes_out,
                                             double* __restrict__ mant_out, int32_t* __restrict__ exp_out,
                                             uint32_t* __restrict__ n_flagged, uint2* __restrict__ flagged,
                                             const uint32_t flagged_cap, const uint32_t row, const uint32_t j) {
    const uint32_t e11 = (static_cast<uint32_t>(__double2hiint(st.X)) >> 20) & 0x7ffu;
    if (e11 != 0) {
        st.X = __dmul_rn(st.X, __hiloint2double(static_cast<int>((2046u - e11) << 20), 0));
        st.ex += static_cast<int>(e11) - 1023;
    }
    nodes_out[o] = st.nodes;
    if (mant_out) mant_out[o] = st.X;
    if (exp_out) exp_out[o] = st.ex;
    if (!(st.rho < 9.765625e-4)) {
        const uint32_t pos = atomicAdd(n_flagged, 1u);
        if (pos < flagged_cap) flagged[pos] = make_uint2(row, j);
    }
}

__global__ void segment_combine_kernel(const SegOut so, const Job* __restrict__ jobs, uint32_t n_jobs,
                                       uint32_t n_seg, uint64_t out_stride,
                                       uint32_t* __restrict__ nodes_out, double* __restrict__ mant_out,
                                       int32_t* __restrict__ exp_out, uint32_t* __restrict__ n_flagged,
                                       uint2* __restrict__ flagged, uint32_t flagged_cap, const double D_start) {
    const uint32_t row = blockIdx.x;  // rows on grid.x (up to 2^31 - 1 of them), energy blocks on grid.y
    const uint32_t j   = blockIdx.y * blockDim.x + threadIdx.x;
    if (row >= n_jobs || j >= jobs[row].nE) return;
    CombState st{1.0, D_start, 0.0, 0, 0u};  // start state in (X, D): X form (1, S - X = -1), D form (1, 1)
    for (uint32_t sgm = 0; sgm < n_seg; sgm++)
        combine_segment(st, so, (static_cast<uint64_t>(row) * n_seg + sgm) * out_stride + j, sgm + 1 == n_seg, D_start);
    combine_emit(st, static_cast<uint64_t>(row) * out_stride + j, nodes_out, mant_out, exp_out, n_flagged, flagged,
                 flagged_cap, row, j);
}

// ---------------------------------------------------------------------------
// N4 combine as a BLOCK-LEVEL PARALLEL PREFIX over the 2x2 transfer matrices (many segments).
//   CTA = 32 trial energies (threadIdx.x: coalesced loads) x kLanes segment lanes (threadIdx.y);
//   lane l owns the m = ceil(n_seg / kLanes) consecutive segments [l m, (l+1) m).
//   1. every lane multiplies its m matrices (entries with one shared binary exponent);
//   2. Hillis-Steele inclusive scan of the lane products through shared memory, log2(kLanes)
//      rounds of one 2x2 product each: Pi_l = M_l ... M_0;
//   3. lane l enters its segments with v = Pi_{l-1} v_0 and applies them one by one with the very
//      body of the serial kernel (node corrections, cancellation guard), so the signs that decide
//      the node count are taken at EVERY segment boundary, not only at lane boundaries;
//   4. the guard travels across lanes as the affine map rho -> alpha rho + beta of each lane, plus
//      the MEASURED mismatch between lane l's exit state and the prefix's entry state of lane
//      l + 1 (relative difference of their D/X ratios): rounding inside the matrix products shows
//      up there and nowhere else, so no a-priori bound on the tree's error is needed.
// Node counts are the serial kernel's (integers decided by the same signs; whatever the guard
// does not vouch for is recomputed by the sequential sweep); tails agree to rounding.
// ---------------------------------------------------------------------------
struct Mat2 {
    double a, b, c, d;  // [[a, b], [c, d]] 2^e acting on (X, D)
    int    e;
};
__device__ __forceinline__ void mat2_renorm(Mat2& m) {
    const double   mx  = fmax(fmax(fabs(m.a), fabs(m.b)), fmax(fabs(m.c), fabs(m.d)));
    const uint32_t e11 = (static_cast<uint32_t>(__double2hiint(mx)) >> 20) & 0x7ffu;
    if (e11 != 0 && e11 != 0x7ffu) {
        const double sc = __hiloint2double(static_cast<int>((2046u - e11) << 20), 0);
        m.a *= sc; m.b *= sc; m.c *= sc; m.d *= sc;
        m.e += static_cast<int>(e11) - 1023;
    }
}
__device__ __forceinline__ Mat2 mat2_mul(const Mat2& L, const Mat2& R) {  // L after R
    Mat2 m;
    m.a = fma(L.a, R.a, L.b * R.c);
    m.b = fma(L.a, R.b, L.b * R.d);
    m.c = fma(L.c, R.a, L.d * R.c);
    m.d = fma(L.c, R.b, L.d * R.d);
    m.e = L.e + R.e;
    mat2_renorm(m);
    return m;
}
__device__ __forceinline__ Mat2 mat2_load(const SegOut& so, const uint64_t o) {
    int sh = so.eB[o] - so.eA[o];
    sh     = sh > 1000 ? 1000 : (sh < -1000 ? -1000 : sh);
    Mat2 m{so.XA[o], scalbn(so.XB[o], sh), so.SA[o], scalbn(so.SB[o], sh), so.eA[o]};
    return m;
}

template <int kLanes>
__global__ void __launch_bounds__(32 * kLanes)
segment_prefix_kernel(const SegOut so, const Job* __restrict__ jobs, uint32_t n_jobs, uint32_t n_seg,
                      uint64_t out_stride, uint32_t* __restrict__ nodes_out, double* __restrict__ mant_out,
                      int32_t* __restrict__ exp_out, uint32_t* __restrict__ n_flagged, uint2* __restrict__ flagged,
                      uint32_t flagged_cap, const double D_start) {
    __shared__ double sh_m[4][kLanes][32];
    __shared__ int    sh_e[kLanes][32];
    __shared__ double sh_v[2][kLanes][32];       // exit state of every lane (X, D), normalised
    __shared__ double sh_g[3][kLanes][32];       // alpha, beta, mismatch
    __shared__ int    sh_x[kLanes][32];          // exit exponent
    __shared__ uint32_t sh_n[kLanes][32];
    const uint32_t row = blockIdx.x;
    const uint32_t tx = threadIdx.x, l = threadIdx.y;
    const uint32_t j   = blockIdx.y * 32 + tx;
    const uint32_t nE  = row < n_jobs ? jobs[row].nE : 0u;
    const bool     live = j < nE;
    const uint32_t jj  = live ? j : (nE ? nE - 1 : 0);  // clamp: every thread takes part in the barriers
    const uint32_t m   = (n_seg + kLanes - 1) / kLanes;
    const uint32_t s0  = min(l * m, n_seg), s1 = min(s0 + m, n_seg);
    if (nE == 0) return;  // uniform per CTA

    // 1. product of this lane's segments
    Mat2 M{1.0, 0.0, 0.0, 1.0, 0};
    for (uint32_t sgm = s0; sgm < s1; sgm++)
        M = mat2_mul(mat2_load(so, (static_cast<uint64_t>(row) * n_seg + sgm) * out_stride + jj), M);
    // 2. inclusive scan over the lanes
#pragma unroll
    for (int off = 1; off < kLanes; off <<= 1) {
        sh_m[0][l][tx] = M.a; sh_m[1][l][tx] = M.b; sh_m[2][l][tx] = M.c; sh_m[3][l][tx] = M.d; sh_e[l][tx] = M.e;
        __syncthreads();
        if (l >= static_cast<uint32_t>(off)) {
            const Mat2 R{sh_m[0][l - off][tx], sh_m[1][l - off][tx], sh_m[2][l - off][tx], sh_m[3][l - off][tx], sh_e[l - off][tx]};
            M = mat2_mul(M, R);
        }
        __syncthreads();
    }
    sh_m[0][l][tx] = M.a; sh_m[1][l][tx] = M.b; sh_m[2][l][tx] = M.c; sh_m[3][l][tx] = M.d; sh_e[l][tx] = M.e;
    __syncthreads();
    // 3. entry state of this lane: Pi_{l-1} (1, D_start); then its own segments, serially
    CombState st{1.0, D_start, 0.0, 0, 0u};
    if (l > 0) {
        st.X  = fma(sh_m[0][l - 1][tx], 1.0, sh_m[1][l - 1][tx] * D_start);
        st.D  = fma(sh_m[2][l - 1][tx], 1.0, sh_m[3][l - 1][tx] * D_start);
        st.ex = sh_e[l - 1][tx];
        uint32_t e11 = (static_cast<uint32_t>(__double2hiint(st.X)) >> 20) & 0x7ffu;
        if (e11 == 0) e11 = (static_cast<uint32_t>(__double2hiint(st.D)) >> 20) & 0x7ffu;
        if (e11 != 0 && e11 != 0x7ffu) {
            const double sc = __hiloint2double(static_cast<int>((2046u - e11) << 20), 0);
            st.X *= sc;
            st.D *= sc;
            st.ex += static_cast<int>(e11) - 1023;
        }
    }
    const double Xin = st.X, Din = st.D;
    // the guard of this lane as an affine map of the incoming rho: rho_out = alpha rho_in + beta, with
    // beta = the lane's own run from rho_in = 0 and alpha = prod 1/c (1/c read off rho' = (rho + eta) / c)
    double alpha = 1.0;
    for (uint32_t sgm = s0; sgm < s1; sgm++) {
        const double r0 = st.rho;
        combine_segment(st, so, (static_cast<uint64_t>(row) * n_seg + sgm) * out_stride + jj, sgm + 1 == n_seg, D_start);
        alpha *= st.rho / (r0 + 5.820766091346741e-11);
    }
    sh_v[0][l][tx] = st.X; sh_v[1][l][tx] = st.D; sh_x[l][tx] = st.ex; sh_n[l][tx] = st.nodes;
    sh_g[0][l][tx] = alpha;
    sh_g[1][l][tx] = st.rho;  // beta
    // 4. mismatch between the previous lane's exit state and this lane's entry state (set by lane l for l-1)
    __syncthreads();
    double mis = 0.0;
    if (l > 0 && s0 < n_seg) {
        const double Xp = sh_v[0][l - 1][tx], Dp = sh_v[1][l - 1][tx];
        const double u = Xp * Din, w = Dp * Xin;
        const double den = fmax(fabs(u), fabs(w));
        mis = den > 0.0 ? fabs(u - w) / den : 1.0;
    }
    sh_g[2][l][tx] = mis;
    __syncthreads();
    if (l == 0 && live) {
        double   rho = 0.0;
        uint32_t nodes = 0;
        uint32_t last = 0;
        for (uint32_t q = 0; q < static_cast<uint32_t>(kLanes) && q * m < n_seg; q++) {
            rho   = sh_g[0][q][tx] * (rho + sh_g[2][q][tx]) + sh_g[1][q][tx];
            nodes += sh_n[q][tx];
            last = q;
        }
        CombState fin{sh_v[0][last][tx], sh_v[1][last][tx], rho, sh_x[last][tx], nodes};
        combine_emit(fin, static_cast<uint64_t>(row) * out_stride + j, nodes_out, mant_out, exp_out, n_flagged, flagged,
                     flagged_cap, row, j);
    }
}

// Fix-up of flagged energies: one single-energy Job per flagged (row, j), E reproduced with the
// very operations the sweep uses, so the sequential kernel returns the oracle's bits for it.
__global__ void make_fixup_jobs_kernel(const Job* __restrict__ jobs, const double* __restrict__ Eexp,
                                       const uint2* __restrict__ flagged, uint32_t n, int packed,
                                       Job* __restrict__ out, double* __restrict__ E_out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const Job    jb = jobs[flagged[i].x];
    const uint32_t j = flagged[i].y;
    double       E;
    if (Eexp != nullptr) E = Eexp[jb.e_off + j];
    else E = __dadd_rn(jb.E0, __dmul_rn(__ull2double_rn(static_cast<unsigned long long>(jb.j0) + j), jb.dE));
    E_out[i] = E;
    if (packed && i > 0) return;
    Job f;
    f.E0    = packed ? 0.0 : E;
    f.dE    = 0.0;
    f.e_off = 0;
    f.curve = jb.curve;
    f.j0    = 0;
    f.nE    = packed ? n : 1;
    f.level = jb.level;
    f.slot  = jb.slot;
    f.pad   = 0;
    out[i]  = f;
}

__global__ void scatter_fixup_kernel(const uint2* __restrict__ flagged, uint32_t n, uint64_t out_stride,
                                     const uint32_t* __restrict__ fn, const double* __restrict__ fm,
                                     const int32_t* __restrict__ fe, uint32_t* __restrict__ nodes_out,
                                     double* __restrict__ mant_out, int32_t* __restrict__ exp_out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint64_t o = static_cast<uint64_t>(flagged[i].x) * out_stride + flagged[i].y;
    nodes_out[o]     = fn[i];
    if (mant_out) mant_out[o] = fm[i];
    if (exp_out) exp_out[o] = fe[i];
}

// ---------------------------------------------------------------------------
// Bracketing helpers (N5/N6).  All decisions are integer comparisons of node
// counts, so they reproduce the oracle's orc_solve_levels exactly.
// ---------------------------------------------------------------------------

// Row c of the coarse sweep: E_j = E_lo[c] + j*dE with dE = (E_hi-E_lo)/(nE-1) (grid == 0), or the
// caller's affine grid E_j = E0[c] + (j0 + j) * dE[c] (grid == 1: a rank's slice of a global
// grid, reproducing the global grid's energies bit for bit).
__global__ void make_coarse_jobs_kernel(const double* __restrict__ E_lo,
                                        const double* __restrict__ E_hi, uint32_t n_curves,
                                        uint32_t nE, int grid, uint32_t j0, Job* __restrict__ jobs) {
    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_curves) return;
    Job jb;
    jb.E0 = E_lo[c];
    if (grid) jb.dE = E_hi[c];
    else jb.dE = nE > 1 ? __ddiv_rn(__dsub_rn(E_hi[c], E_lo[c]), static_cast<double>(nE - 1)) : 0.0;
    jb.e_off = static_cast<uint64_t>(c) * nE;
    jb.curve = c;
    jb.j0    = grid ? j0 : 0u;
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
    if (jobs[row].nE == 0) return;  // idle refinement row (uniform per block)
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
                                    uint32_t* __restrict__ state, uint32_t* __restrict__ n_below,
                                    uint32_t* __restrict__ n_first) {
    const uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n_curves * n_lev) return;
    const uint32_t c = idx / n_lev, l = idx - c * n_lev, v = v_min + l;
    const Job      jb    = jobs[c];
    const uint32_t first = nodes[static_cast<uint64_t>(c) * stride];
    const uint32_t last  = nodes[static_cast<uint64_t>(c) * stride + jb.nE - 1];
    if (l == 0) {
        n_below[c] = last;
        n_first[c] = first;
    }
    const uint32_t j = jstar[idx];
    if (last <= v || first > v || j == kNone || j == 0) {
        lo[idx]    = __longlong_as_double(0x7ff8000000000000LL);
        hi[idx]    = __longlong_as_double(0x7ff8000000000000LL);
        state[idx] = 0;
        return;
    }
    const unsigned long long jg = static_cast<unsigned long long>(jb.j0) + j;  // index on the global grid
    lo[idx]    = __dadd_rn(jb.E0, __dmul_rn(__ull2double_rn(jg - 1), jb.dE));
    hi[idx]    = __dadd_rn(jb.E0, __dmul_rn(__ull2double_rn(jg), jb.dE));
    state[idx] = 1;
}

// Convergence test + ORDERED COMPACTION of the still-active brackets into refinement rows (used when
// one curve is resident: the compact rows are then swept with the flat-row mapping).  Single CTA;
// a few thousand (curve, level) entries at most.
__global__ void compact_refine_jobs_kernel(const double* __restrict__ lo, const double* __restrict__ hi,
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
    __syncthreads();
    for (uint32_t pos = running + threadIdx.x; pos < total; pos += blockDim.x) jobs_out[pos].nE = 0;  // idle tail rows
    if (threadIdx.x == 0) *n_active = running;
}

// Convergence test + refinement rows.  Rows are DENSE: row = curve * n_lev_pad + l (n_lev_pad a
// multiple of the sweep's rows-per-CTA so that packed rows of a CTA share a curve); a row that is
// absent, converged or padding gets nE = 0 and costs nothing downstream.
__global__ void make_refine_jobs_kernel(const double* __restrict__ lo, const double* __restrict__ hi,
                                        uint32_t* __restrict__ state, uint32_t n_curves, uint32_t n_lev,
                                        uint32_t n_lev_pad, uint32_t v_min, double rel_tol, uint32_t M,
                                        Job* __restrict__ jobs_out, uint32_t* __restrict__ n_active) {
    const uint32_t row = blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= n_curves * n_lev_pad) return;
    const uint32_t c = row / n_lev_pad, l = row - c * n_lev_pad;
    Job            jb;
    jb.E0    = 0.0;
    jb.dE    = 0.0;
    jb.e_off = 0;
    jb.curve = c;
    jb.j0    = 1;
    jb.nE    = 0;
    jb.level = v_min + l;
    jb.slot  = c * n_lev + l;
    jb.pad   = 0;
    if (l < n_lev) {
        const uint32_t idx = jb.slot;
        if (state[idx] == 1) {
            const double a = lo[idx], b = hi[idx];
            const double w   = __dsub_rn(b, a);
            const double mag = fmax(fabs(a), fabs(b));
            if (w <= __dmul_rn(rel_tol, mag)) {
                state[idx] = 2;
            } else {
                jb.E0 = a;
                jb.dE = __ddiv_rn(__dsub_rn(b, a), static_cast<double>(M + 1));
                jb.nE = M;
                atomicAdd(n_active, 1u);
            }
        }
    }
    jobs_out[row] = jb;
}

__global__ void bracket_update_kernel(const Job* __restrict__ jobs,
                                      const uint32_t* __restrict__ jstar, uint32_t n_jobs,
                                      uint32_t M, double* __restrict__ lo, double* __restrict__ hi,
                                      uint32_t* __restrict__ state) {
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_jobs) return;
    const Job      jb  = jobs[r];
    if (jb.nE == 0) return;  // idle row
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
// N1 for tabulated sources: evaluate the natural cubic spline (coefficients a,b,c,d per knot
// interval, computed on the host by eps_spline_coefficients) on the uniform grid
// r_i = rmin + i*h.  Same operations as the oracle's orc_spline_resample: binary search for the
// interval, dx = r - r_k, fma(fma(fma(d,dx,c),dx,b),dx,a).
// ---------------------------------------------------------------------------
__global__ void spline_eval_kernel(const double* __restrict__ knots, const double* __restrict__ coef,
                                   uint32_t K, double rmin, double h, uint32_t N, double* __restrict__ out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const double x  = __dadd_rn(rmin, __dmul_rn(static_cast<double>(i), h));
    uint32_t     lo = 0, hi = K - 1;
    while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) / 2;
        if (knots[mid] <= x) lo = mid;
        else hi = mid;
    }
    const double  dx = __dsub_rn(x, knots[lo]);
    const double* c  = coef + 4 * lo;
    out[i]           = __fma_rn(__fma_rn(__fma_rn(c[3], dx, c[2]), dx, c[1]), dx, c[0]);
}

// ---------------------------------------------------------------------------
// Rotational states (spec DESIGN.md section 3.7; oracle: orc_centrifugal).  Expands n_curves raw
// tables into n_curves * n_J effective ones, V_J = V + J(J+1) (h^2 / 12 s) / r^2, row c*n_J + j.
// Pure streaming: 8 B read (L2 hit for all but the first J of a curve) + 8 B written per point.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
centrifugal_kernel(const double* __restrict__ Vraw, const double* __restrict__ scale, const double* __restrict__ rmin,
                   const double* __restrict__ hstep, const uint32_t* __restrict__ J, uint32_t n_J, uint32_t N,
                   double* __restrict__ Vout, double* __restrict__ scale_out) {
    const uint32_t row = blockIdx.x, c = row / n_J, j = row - c * n_J;  // rows on grid.x: up to 2^31 - 1 of them
    const uint32_t Jv  = J[j];
    const double   s = scale[c], h = hstep[c], r0 = rmin[c];
    const double   jj = static_cast<double>(static_cast<unsigned long long>(Jv) * (static_cast<unsigned long long>(Jv) + 1ull));
    const double   cj = __ddiv_rn(__dmul_rn(jj, __dmul_rn(h, h)), __dmul_rn(12.0, s));
    const double*  src = Vraw + static_cast<uint64_t>(c) * N;
    double*        dst = Vout + static_cast<uint64_t>(row) * N;
    if (blockIdx.y == 0 && threadIdx.x == 0) scale_out[row] = s;
    for (uint32_t i = blockIdx.y * blockDim.x + threadIdx.x; i < N; i += gridDim.y * blockDim.x) {
        double v = src[i];
        if (Jv != 0) {
            const double r = __dadd_rn(r0, __dmul_rn(static_cast<double>(i), h));
            v              = __dadd_rn(v, __ddiv_rn(cj, __dmul_rn(r, r)));
            if (v > 1e300) v = 1e300;
        }
        dst[i] = v;
    }
}

// ---------------------------------------------------------------------------
// Preparation on the device (spec DESIGN.md section 3.2; oracle: orc_prep).  One CTA per curve:
// first argmin of q = s V, the window around it with q - q_min <= T_MAX, the coefficient table
// F_k = (1 - q_{i0+k}) / 12 written into the curve's slot (padded with 1/12).  Every value is
// produced by the same single IEEE operations as on the host, so the table is bit-identical to
// the oracle's.  status[c]: 0 ok, 1 non-finite table value, 2 window shorter than 2 steps.
// ---------------------------------------------------------------------------
struct PrepOut {  // one per curve, read back by the host
    uint32_t i0, n_steps, status, pad;
    double   v_min, v_last;
};

constexpr int kPrepThreads = 512;

__global__ void __launch_bounds__(kPrepThreads)
prep_curves_kernel(const double* __restrict__ V, const double* __restrict__ scale, uint32_t N, uint64_t slot,
                   double t_max, double* __restrict__ F, double* __restrict__ A, CurveDev* __restrict__ curves,
                   PrepOut* __restrict__ out) {
    __shared__ double   red_q[kPrepThreads / 32];
    __shared__ uint32_t red_i[kPrepThreads / 32];
    __shared__ uint32_t red_a[kPrepThreads / 32], red_b[kPrepThreads / 32], red_bad[kPrepThreads / 32];
    __shared__ double   qmin_sh;
    __shared__ uint32_t m_sh, ilo_sh, ihi_sh, bad_sh;
    const uint32_t c = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const double*  v = V + static_cast<uint64_t>(c) * N;
    const double   s = scale[c];

    // ---- first argmin of q (ties -> smallest index), non-finite check ----
    double   bq  = __longlong_as_double(0x7ff0000000000000LL);  // +inf
    uint32_t bi  = kNone, bad = 0;
    for (uint32_t i = tid; i < N; i += kPrepThreads) {
        const double vi = v[i];
        if (!isfinite(vi)) bad = 1;
        const double q = __dmul_rn(s, vi);
        if (q < bq) {  // strided ascending i per thread: strict < keeps the first index
            bq = q;
            bi = i;
        }
    }
    for (int o = 16; o > 0; o >>= 1) {
        const double   oq = __shfl_xor_sync(0xffffffffu, bq, o);
        const uint32_t oi = __shfl_xor_sync(0xffffffffu, bi, o);
        bad |= __shfl_xor_sync(0xffffffffu, bad, o);
        if (oq < bq || (oq == bq && oi < bi)) {
            bq = oq;
            bi = oi;
        }
    }
    if (lane == 0) {
        red_q[warp]   = bq;
        red_i[warp]   = bi;
        red_bad[warp] = bad;
    }
    __syncthreads();
    if (tid == 0) {
        double   q = red_q[0];
        uint32_t i = red_i[0], b = red_bad[0];
        for (int w = 1; w < kPrepThreads / 32; w++) {
            b |= red_bad[w];
            if (red_q[w] < q || (red_q[w] == q && red_i[w] < i)) {
                q = red_q[w];
                i = red_i[w];
            }
        }
        qmin_sh = q;
        m_sh    = i;
        bad_sh  = b;
    }
    __syncthreads();
    const uint32_t m   = m_sh;
    const double   thr = __dadd_rn(qmin_sh, t_max);

    // ---- window: ilo = 1 + last j < m with q_j > thr (else 0); ihi = first j > m with q_j > thr, minus 1 (else N-1) ----
    uint32_t a = 0, b = N;  // a: (last j<m above thr) + 1;  b: first j>m above thr (N = none)
    for (uint32_t i = tid; i < N; i += kPrepThreads) {
        const double q = __dmul_rn(s, v[i]);
        if (q > thr) {
            if (i < m) a = max(a, i + 1);
            else if (i > m) b = min(b, i);
        }
    }
    for (int o = 16; o > 0; o >>= 1) {
        a = max(a, __shfl_xor_sync(0xffffffffu, a, o));
        b = min(b, __shfl_xor_sync(0xffffffffu, b, o));
    }
    if (lane == 0) {
        red_a[warp] = a;
        red_b[warp] = b;
    }
    __syncthreads();
    if (tid == 0) {
        uint32_t aa = 0, bb = N;
        for (int w = 0; w < kPrepThreads / 32; w++) {
            aa = max(aa, red_a[w]);
            bb = min(bb, red_b[w]);
        }
        ilo_sh = aa;
        ihi_sh = bb == N ? N - 1 : bb - 1;
    }
    __syncthreads();
    const uint32_t i0   = ilo_sh < 1 ? 1 : ilo_sh;
    const uint32_t iend = (ihi_sh + 1 < N - 1) ? ihi_sh + 1 : N - 1;
    uint32_t       status = bad_sh ? 1u : (iend >= i0 + 2 ? 0u : 2u);
    const uint32_t n      = status ? 0u : iend - i0;

    double* dst = F + static_cast<uint64_t>(c) * slot;
    for (uint64_t k = tid; k < slot; k += kPrepThreads)
        dst[k] = (k < n) ? __ddiv_rn(__dsub_rn(1.0, __dmul_rn(s, v[i0 + k])), 12.0) : 1.0 / 12.0;
    if (A != nullptr) {  // D-form table A_k = 12 q_k in the same slot layout (pad 0)
        double* dsa = A + static_cast<uint64_t>(c) * slot;
        for (uint64_t k = tid; k < slot; k += kPrepThreads) dsa[k] = (k < n) ? __dmul_rn(12.0, __dmul_rn(s, v[i0 + k])) : 0.0;
    }
    if (tid == 0) {
        const double vm = (m != kNone) ? v[m] : __longlong_as_double(0x7ff8000000000000LL);  // no finite value at all
        curves[c] = CurveDev{static_cast<uint64_t>(c) * slot, n, i0, s, vm};
        out[c]    = PrepOut{i0, n, status, 0u, vm, v[N - 1]};
    }
}

// ---------------------------------------------------------------------------
// Preparation of a FEW LONG curves (C3: one curve of 10^6 points): prep_curves_kernel gives one CTA
// per curve, i.e. one SM for the whole table (0.84 ms for 10^6 points).  The same three passes are
// cut into `parts` contiguous chunks per curve, one CTA each, with the tiny cross-chunk reductions
// redone by every CTA of the next pass (parts <= 128).  Same operations per element, same tie rules
// (first minimum; window bounds as max / min of indices) => the same bits as prep_curves_kernel.
// ---------------------------------------------------------------------------
struct PrepPart {
    double   q;         // chunk minimum of q = s V (+inf when the chunk is empty)
    uint32_t idx, bad;  // its first index; non-finite value seen
    uint32_t a, b;      // window partials: (last j < m above thr) + 1, first j > m above thr (N = none)
    uint32_t pad[2];
};

constexpr uint32_t kPrepPartsMax = 128;

__device__ __forceinline__ void prep_chunk(uint32_t N, uint32_t parts, uint32_t part, uint32_t& lo, uint32_t& hi) {
    const uint32_t chunk = (N + parts - 1) / parts;
    lo = min(N, part * chunk);
    hi = min(N, lo + chunk);
}

// first argmin over the chunk partials of one curve (every thread computes the same value)
__device__ __forceinline__ void prep_reduce_argmin(const PrepPart* __restrict__ pp, uint32_t parts, double& q, uint32_t& m,
                                                   uint32_t& bad) {
    q   = pp[0].q;
    m   = pp[0].idx;
    bad = pp[0].bad;
    for (uint32_t k = 1; k < parts; k++) {
        bad |= pp[k].bad;
        if (pp[k].q < q || (pp[k].q == q && pp[k].idx < m)) {
            q = pp[k].q;
            m = pp[k].idx;
        }
    }
}

__global__ void __launch_bounds__(kPrepThreads)
prep_part_argmin_kernel(const double* __restrict__ V, const double* __restrict__ scale, uint32_t N, uint32_t parts,
                        PrepPart* __restrict__ out) {
    __shared__ double   red_q[kPrepThreads / 32];
    __shared__ uint32_t red_i[kPrepThreads / 32], red_bad[kPrepThreads / 32];
    const uint32_t c = blockIdx.y, part = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const double*  v = V + static_cast<uint64_t>(c) * N;
    const double   s = scale[c];
    uint32_t       lo, hi;
    prep_chunk(N, parts, part, lo, hi);
    double   bq = __longlong_as_double(0x7ff0000000000000LL);
    uint32_t bi = kNone, bad = 0;
    for (uint32_t i = lo + tid; i < hi; i += kPrepThreads) {
        const double vi = v[i];
        if (!isfinite(vi)) bad = 1;
        const double q = __dmul_rn(s, vi);
        if (q < bq) {
            bq = q;
            bi = i;
        }
    }
    for (int o = 16; o > 0; o >>= 1) {
        const double   oq = __shfl_xor_sync(0xffffffffu, bq, o);
        const uint32_t oi = __shfl_xor_sync(0xffffffffu, bi, o);
        bad |= __shfl_xor_sync(0xffffffffu, bad, o);
        if (oq < bq || (oq == bq && oi < bi)) {
            bq = oq;
            bi = oi;
        }
    }
    if (lane == 0) {
        red_q[warp]   = bq;
        red_i[warp]   = bi;
        red_bad[warp] = bad;
    }
    __syncthreads();
    if (tid == 0) {
        double   q = red_q[0];
        uint32_t i = red_i[0], b = red_bad[0];
        for (int w = 1; w < kPrepThreads / 32; w++) {
            b |= red_bad[w];
            if (red_q[w] < q || (red_q[w] == q && red_i[w] < i)) {
                q = red_q[w];
                i = red_i[w];
            }
        }
        PrepPart& o = out[static_cast<uint64_t>(c) * parts + part];
        o.q   = q;
        o.idx = i;
        o.bad = b;
    }
}

__global__ void __launch_bounds__(kPrepThreads)
prep_part_window_kernel(const double* __restrict__ V, const double* __restrict__ scale, uint32_t N, uint32_t parts,
                        double t_max, PrepPart* __restrict__ pp_all) {
    __shared__ uint32_t red_a[kPrepThreads / 32], red_b[kPrepThreads / 32];
    const uint32_t c = blockIdx.y, part = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const double*  v  = V + static_cast<uint64_t>(c) * N;
    const double   s  = scale[c];
    PrepPart*      pp = pp_all + static_cast<uint64_t>(c) * parts;
    double         qmin;
    uint32_t       m, bad;
    prep_reduce_argmin(pp, parts, qmin, m, bad);
    const double thr = __dadd_rn(qmin, t_max);
    uint32_t     lo, hi;
    prep_chunk(N, parts, part, lo, hi);
    uint32_t a = 0, b = N;
    for (uint32_t i = lo + tid; i < hi; i += kPrepThreads) {
        const double q = __dmul_rn(s, v[i]);
        if (q > thr) {
            if (i < m) a = max(a, i + 1);
            else if (i > m) b = min(b, i);
        }
    }
    for (int o = 16; o > 0; o >>= 1) {
        a = max(a, __shfl_xor_sync(0xffffffffu, a, o));
        b = min(b, __shfl_xor_sync(0xffffffffu, b, o));
    }
    if (lane == 0) {
        red_a[warp] = a;
        red_b[warp] = b;
    }
    __syncthreads();  // (also orders every thread's reads of pp[].q/idx before the writes below: disjoint fields)
    if (tid == 0) {
        uint32_t aa = 0, bb = N;
        for (int w = 0; w < kPrepThreads / 32; w++) {
            aa = max(aa, red_a[w]);
            bb = min(bb, red_b[w]);
        }
        pp[part].a = aa;
        pp[part].b = bb;
    }
}

__global__ void __launch_bounds__(kPrepThreads)
prep_part_finish_kernel(const double* __restrict__ V, const double* __restrict__ scale, uint32_t N, uint64_t slot,
                        uint32_t parts, const PrepPart* __restrict__ pp_all, double* __restrict__ F,
                        double* __restrict__ A, CurveDev* __restrict__ curves, PrepOut* __restrict__ out) {
    const uint32_t  c = blockIdx.y, part = blockIdx.x, tid = threadIdx.x;
    const double*   v  = V + static_cast<uint64_t>(c) * N;
    const double    s  = scale[c];
    const PrepPart* pp = pp_all + static_cast<uint64_t>(c) * parts;
    double          qmin;
    uint32_t        m, bad;
    prep_reduce_argmin(pp, parts, qmin, m, bad);
    uint32_t aa = 0, bb = N;
    for (uint32_t k = 0; k < parts; k++) {
        aa = max(aa, pp[k].a);
        bb = min(bb, pp[k].b);
    }
    const uint32_t ilo = aa, ihi = bb == N ? N - 1 : bb - 1;
    const uint32_t i0   = ilo < 1 ? 1 : ilo;
    const uint32_t iend = (ihi + 1 < N - 1) ? ihi + 1 : N - 1;
    const uint32_t status = bad ? 1u : (iend >= i0 + 2 ? 0u : 2u);
    const uint32_t n      = status ? 0u : iend - i0;
    const uint64_t chunk = (slot + parts - 1) / parts, k_lo = min(slot, part * chunk), k_hi = min(slot, k_lo + chunk);
    double*        dst = F + static_cast<uint64_t>(c) * slot;
    for (uint64_t k = k_lo + tid; k < k_hi; k += kPrepThreads)
        dst[k] = (k < n) ? __ddiv_rn(__dsub_rn(1.0, __dmul_rn(s, v[i0 + k])), 12.0) : 1.0 / 12.0;
    if (A != nullptr) {
        double* dsa = A + static_cast<uint64_t>(c) * slot;
        for (uint64_t k = k_lo + tid; k < k_hi; k += kPrepThreads) dsa[k] = (k < n) ? __dmul_rn(12.0, __dmul_rn(s, v[i0 + k])) : 0.0;
    }
    if (part == 0 && tid == 0) {
        const double vm = (m != kNone) ? v[m] : __longlong_as_double(0x7ff8000000000000LL);
        curves[c] = CurveDev{static_cast<uint64_t>(c) * slot, n, i0, s, vm};
        out[c]    = PrepOut{i0, n, status, 0u, vm, v[N - 1]};
    }
}

// ---------------------------------------------------------------------------
// N7: wavefunctions of located levels (spec DESIGN.md section 3.6; oracle:
// orc_wavefunction).  Few (curve, level) items, each a strictly serial three-term
// recurrence u_{k+1} = fma(c_k, u_k, -u_{k-1}), c_k = 1/fp_k - 10, so the design
// keeps ONE dependent DFMA per grid step on the critical path: the 256 threads of
// the CTA turn a 1024-step tile of the F table into (c_k, r_k = 1/fp_k) in shared
// memory (the FP64 divisions run in parallel, off the chain), one thread marches
// the tile, and all threads form psi_k = u_k r_k and store it coalesced.
//   grid = (n_items, 2): blockIdx.y = 0 outward (k = 0..m), 1 inward (k = n-1..m).
// ---------------------------------------------------------------------------
constexpr int kWfTile    = 1024;
constexpr int kWfThreads = 256;

__device__ __forceinline__ void renorm_pair(double& a, double& b, int& expo) {
    const uint32_t ex = (static_cast<uint32_t>(__double2hiint(a)) >> 20) & 0x7ffu;
    if (ex != 0) {
        const double sc = __hiloint2double(static_cast<int>((2046u - ex) << 20), 0);
        a = __dmul_rn(a, sc);
        b = __dmul_rn(b, sc);
        expo += static_cast<int>(ex) - 1023;
    }
}

// raw[item][k]: psi_k in the scale of its 128-block; bexp[item][dir][block]: cumulative
// exponent of that block; match[item]: m (kNone: no classically allowed point);
// in_at_m[item]: the inward branch's value at m, in_exp_m[item] its block exponent.
__global__ void __launch_bounds__(kWfThreads)
wavefunction_march_kernel(const double* __restrict__ F, const CurveDev* __restrict__ curves,
                          const double* __restrict__ E, uint32_t n_lev, uint64_t raw_stride,
                          uint32_t bexp_stride, double* __restrict__ raw, int32_t* __restrict__ bexp,
                          uint32_t* __restrict__ match, double* __restrict__ in_at_m,
                          int32_t* __restrict__ in_exp_m) {
    __shared__ double   c_s[kWfTile];  // c_k, overwritten by u_k during the march
    __shared__ double   r_s[kWfTile];
    __shared__ int32_t  e_s[kWfTile / kRenorm];
    __shared__ uint32_t red[kWfThreads / 32];
    __shared__ uint32_t m_sh;

    const uint32_t item = blockIdx.x, dir = blockIdx.y, tid = threadIdx.x;
    const CurveDev cv   = curves[item / n_lev];
    const double   En   = E[item];
    const uint32_t n    = cv.n_steps;
    const double*  Fc   = F + cv.f_off;
    const double   ep   = __ddiv_rn(__dmul_rn(cv.s, En), 12.0);

    // m = max{k : F_k + ep > 1/12} (+1 so that 0 means "none")
    uint32_t best = 0;
    if (En == En && n >= 3)
        for (uint32_t k = tid; k < n; k += kWfThreads)
            if (__dadd_rn(Fc[k], ep) > 1.0 / 12.0) best = k + 1;
    for (int o = 16; o > 0; o >>= 1) best = max(best, __shfl_xor_sync(0xffffffffu, best, o));
    if ((tid & 31) == 0) red[tid >> 5] = best;
    __syncthreads();
    if (tid == 0) {
        uint32_t b = 0;
        for (int w = 0; w < kWfThreads / 32; w++) b = max(b, red[w]);
        m_sh = b;
    }
    __syncthreads();
    if (m_sh == 0) {
        if (tid == 0 && dir == 0) match[item] = kNone;
        return;
    }
    uint32_t m = m_sh - 1;
    m          = m < 1 ? 1 : m;
    m          = m > n - 2 ? n - 2 : m;
    if (tid == 0 && dir == 0) match[item] = m;

    const uint32_t count = dir == 0 ? m + 1 : n - m;
    double*        out   = raw + static_cast<uint64_t>(item) * raw_stride;
    int32_t*       bx    = bexp + (static_cast<uint64_t>(item) * 2 + dir) * bexp_stride;
    double         u_cur = 0.0, u_prev = 0.0;  // live in thread 0 only
    int            cum   = 0;

    for (uint32_t j0 = 0; j0 < count; j0 += kWfTile) {
        const uint32_t nj = min(static_cast<uint32_t>(kWfTile), count - j0);
        for (uint32_t jj = tid; jj < nj; jj += kWfThreads) {
            const uint32_t k  = dir == 0 ? j0 + jj : n - 1 - (j0 + jj);
            const double   fp = __dadd_rn(Fc[k], ep);
            const double   r  = __ddiv_rn(1.0, fp);
            r_s[jj]           = r;
            c_s[jj]           = __dsub_rn(r, 10.0);
        }
        __syncthreads();
        if (tid == 0) {
            if (j0 == 0) {
                const uint32_t k0 = dir == 0 ? 0 : n - 1;
                u_cur             = __dadd_rn(Fc[k0], ep);
            }
            for (uint32_t b0 = 0; b0 < nj; b0 += kRenorm) {
                if (j0 + b0 > 0) renorm_pair(u_cur, u_prev, cum);
                e_s[b0 / kRenorm] = cum;
                const uint32_t nb = min(static_cast<uint32_t>(kRenorm), nj - b0);
#pragma unroll 8
                for (uint32_t q = 0; q < nb; q++) {
                    const double c  = c_s[b0 + q];
                    c_s[b0 + q]     = u_cur;
                    const double un = __fma_rn(c, u_cur, -u_prev);
                    u_prev          = u_cur;
                    u_cur           = un;
                }
            }
        }
        __syncthreads();
        for (uint32_t jj = tid; jj < nj; jj += kWfThreads) {
            const uint32_t k = dir == 0 ? j0 + jj : n - 1 - (j0 + jj);
            const double   a = __dmul_rn(c_s[jj], r_s[jj]);
            if (dir == 1 && k == m) {
                in_at_m[item]  = a;
                in_exp_m[item] = e_s[jj / kRenorm];
            } else {
                out[k] = a;
            }
        }
        if (tid < (nj + kRenorm - 1) / kRenorm) bx[j0 / kRenorm + tid] = e_s[tid];
        __syncthreads();
    }
}

// Matching, common binary scale, normalisation; psi[item][0..n_points) on the full grid.
__global__ void __launch_bounds__(kWfThreads)
wavefunction_finish_kernel(const CurveDev* __restrict__ curves, uint32_t n_lev, uint32_t n_points,
                           const double* __restrict__ grid_step, uint64_t raw_stride,
                           uint32_t bexp_stride, const double* __restrict__ raw,
                           const int32_t* __restrict__ bexp, const uint32_t* __restrict__ match,
                           const double* __restrict__ in_at_m, const int32_t* __restrict__ in_exp_m,
                           double* __restrict__ psi) {
    __shared__ int    red_i[kWfThreads / 32];
    __shared__ double red_d[kWfThreads / 32];
    __shared__ int    emax_sh;
    __shared__ double nrm_sh;
    const uint32_t item = blockIdx.x, tid = threadIdx.x;
    const uint32_t curve = item / n_lev;
    const CurveDev cv    = curves[curve];
    const uint32_t n = cv.n_steps, i0 = cv.i0, m = match[item];
    double*        dst = psi + static_cast<uint64_t>(item) * n_points;
    if (m == kNone) {
        for (uint32_t i = tid; i < n_points; i += kWfThreads) dst[i] = 0.0;
        return;
    }
    const double*  a_raw = raw + static_cast<uint64_t>(item) * raw_stride;
    const int32_t* bx_o  = bexp + static_cast<uint64_t>(item) * 2 * bexp_stride;
    const int32_t* bx_i  = bx_o + bexp_stride;
    const int      c_out_m = bx_o[m / kRenorm], c_in_m = in_exp_m[item];
    const double   a_in = in_at_m[item];
    const double   rho  = (a_in != 0.0) ? __ddiv_rn(a_raw[m], a_in) : 1.0;

    auto value = [&](uint32_t k, int& c) -> double {
        if (k <= m) {
            c = bx_o[k / kRenorm];
            return a_raw[k];
        }
        c = bx_i[(n - 1 - k) / kRenorm] - c_in_m + c_out_m;
        return __dmul_rn(a_raw[k], rho);
    };
    // largest binary exponent
    int emax = INT_MIN;
    for (uint32_t k = tid; k < n; k += kWfThreads) {
        int          c;
        const double a = value(k, c);
        if (a != 0.0) emax = max(emax, c + ilogb(a));
    }
    for (int o = 16; o > 0; o >>= 1) emax = max(emax, __shfl_xor_sync(0xffffffffu, emax, o));
    if ((tid & 31) == 0) red_i[tid >> 5] = emax;
    __syncthreads();
    if (tid == 0) {
        int e = INT_MIN;
        for (int w = 0; w < kWfThreads / 32; w++) e = max(e, red_i[w]);
        emax_sh = e;
    }
    __syncthreads();
    emax = emax_sh;
    // common scale + sum of squares
    double sum = 0.0;
    for (uint32_t k = tid; k < n; k += kWfThreads) {
        int          c;
        const double a  = value(k, c);
        long long    sh = static_cast<long long>(c) - emax;
        if (sh < -2200) sh = -2200;
        const double v = scalbn(a, static_cast<int>(sh));
        dst[i0 + k]    = v;
        sum            = __fma_rn(v, v, sum);
    }
    for (int o = 16; o > 0; o >>= 1) sum = __dadd_rn(sum, __shfl_xor_sync(0xffffffffu, sum, o));
    if ((tid & 31) == 0) red_d[tid >> 5] = sum;
    __syncthreads();
    if (tid == 0) {
        double t = 0.0;
        for (int w = 0; w < kWfThreads / 32; w++) t = __dadd_rn(t, red_d[w]);
        nrm_sh = __dsqrt_rn(__dmul_rn(grid_step[curve], t));
    }
    __syncthreads();
    const double nrm = nrm_sh;
    for (uint32_t i = tid; i < n_points; i += kWfThreads)
        dst[i] = (i >= i0 && i < i0 + n) ? __ddiv_rn(dst[i], nrm) : 0.0;
}

// A-posteriori energy correction of located levels (spec DESIGN.md section 3.8; oracle:
// orc_level_correction): residual of the Numerov equation at the matching point times psi_m over
// the B-norm of psi.  One CTA per (curve, level); reads the normalised psi the finish kernel left on
// the full grid.  The B-norm is a tree reduction, so dE agrees with the oracle's serial sum to
// rounding, not bit for bit.
__global__ void __launch_bounds__(kWfThreads)
level_correction_kernel(const double* __restrict__ F, const CurveDev* __restrict__ curves, const double* __restrict__ E,
                        uint32_t n_lev, uint32_t n_points, const uint32_t* __restrict__ match,
                        const double* __restrict__ psi, double* __restrict__ dE) {
    __shared__ double red_d[kWfThreads / 32];
    const uint32_t item = blockIdx.x, tid = threadIdx.x;
    const CurveDev cv = curves[item / n_lev];
    const uint32_t n = cv.n_steps, m = match[item];
    if (m == kNone || m < 1 || m + 2 > n) {
        if (tid == 0) dE[item] = __longlong_as_double(0x7ff8000000000000LL);
        return;
    }
    const double* p = psi + static_cast<uint64_t>(item) * n_points + cv.i0;
    double        D = 0.0;
    for (uint32_t k = tid; k < n; k += kWfThreads) {
        const double left = k > 0 ? p[k - 1] : 0.0, right = k + 1 < n ? p[k + 1] : 0.0;
        D = __fma_rn(p[k], __fma_rn(10.0, p[k], __dadd_rn(left, right)), D);
    }
    for (int o = 16; o > 0; o >>= 1) D = __dadd_rn(D, __shfl_xor_sync(0xffffffffu, D, o));
    if ((tid & 31) == 0) red_d[tid >> 5] = D;
    __syncthreads();
    if (tid == 0) {
        double t = 0.0;
        for (int w = 0; w < kWfThreads / 32; w++) t = __dadd_rn(t, red_d[w]);
        t = __ddiv_rn(t, 12.0);
        const double* f  = F + cv.f_off;
        const double  ep = __ddiv_rn(__dmul_rn(cv.s, E[item]), 12.0);
        const double  f0 = __dadd_rn(f[m], ep), fl = __dadd_rn(f[m - 1], ep), fr = __dadd_rn(f[m + 1], ep);
        const double  u0 = __dmul_rn(f0, p[m]), ul = __dmul_rn(fl, p[m - 1]), ur = __dmul_rn(fr, p[m + 1]);
        const double  r  = __dsub_rn(__dsub_rn(__dsub_rn(ur, u0), __dsub_rn(u0, ul)),
                                     __dmul_rn(__dsub_rn(1.0, __dmul_rn(12.0, f0)), p[m]));
        dE[item] = -__ddiv_rn(__ddiv_rn(__dmul_rn(p[m], r), t), cv.s);
    }
}

// ---------------------------------------------------------------------------
// DFMA-saturating probe: measures the FP64 (non-tensor) roofline denominator,
// which MEASURED_PEAKS.json does not carry.  8 independent chains per thread.
// ---------------------------------------------------------------------------
constexpr int kProbeIters = 4096;
__global__ void __launch_bounds__(256) fp64_probe_kernel(double* __restrict__ out, double x, double y) {
    double a[8];
#pragma unroll
    for (int i = 0; i < 8; i++) a[i] = threadIdx.x * 1e-3 + i;
#pragma unroll 1
    for (int it = 0; it < kProbeIters; it++) {
#pragma unroll
        for (int u = 0; u < 8; u++)
#pragma unroll
            for (int i = 0; i < 8; i++) a[i] = __fma_rn(a[i], x, y);
    }
    double sum = 0.0;
#pragma unroll
    for (int i = 0; i < 8; i++) sum += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = sum;
}

}  // namespace eps
