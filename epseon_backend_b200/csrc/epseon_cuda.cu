// epseon_cuda.cu -- C ABI (include/epseon_cuda.h) over the sm_100a kernels.
//
// Host side of the hot path that replaces the Vulkan body of
// VibwaAlgorithm<FP>::run (cpp/gpu/include/epseon/gpu/algorithms/vibwa.hpp:605-637):
// logical device -> eps_ctx (stream + device buffers), VMA staging/device/output
// buffer triplets (vibwa.hpp:57-234) -> one resident coefficient table per curve
// plus node/tail/level output buffers, descriptor sets -> kernel arguments.
// No CPU fallback: every compute entry point fails with EPS_ERR_CUDA when the
// CUDA runtime cannot provide the device.
#include "../../include/epseon_cuda.h"
#include "numerov_kernels.cuh"
#include "numerov_cbank.cuh"
#include "cooley_search.cuh"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

using namespace eps;

namespace {

thread_local std::string g_last_error;

constexpr double kTMax       = 0.5;  // window rule / validity bound on |q - e|
constexpr size_t kFlushBytes = 256u << 20;
constexpr int    kEventPairs = 64;

template <typename T>
struct DevBuf {
    T*     p   = nullptr;
    size_t cap = 0;  // elements
    cudaError_t reserve(size_t n) {
        if (n <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p   = nullptr;
        cap = 0;
        cudaError_t e = cudaMalloc(reinterpret_cast<void**>(&p), n * sizeof(T));
        if (e == cudaSuccess) cap = n;
        return e;
    }
    void release() {
        if (p) cudaFree(p);
        p   = nullptr;
        cap = 0;
    }
};

}  // namespace

struct eps_ctx {
    int          dev      = -1;
    int          sm_count = 0;
    cudaStream_t stream   = nullptr;
    std::string  err;

    // resident potentials
    uint32_t                    nC = 0, N = 0;
    uint64_t                    slot = 0;  // doubles per curve
    DevBuf<double>              d_F, d_A, d_V, d_scale, d_spl, d_Vraw, d_rot;
    int                         opt_pack128 = 0;  // EPS_OPT_PACK128: 0 auto, 1 always (when rows fit), 2 never
    int                         opt_form  = 0;  // EPS_OPT_FORM: 0 = X form (4 operations), 1 = D form (accurate, 5 operations)
    int                         form_resident = 0;  // form the resident tables were prepared for
    DevBuf<uint32_t>            d_J;
    DevBuf<PrepOut>             d_prep;
    DevBuf<PrepPart>            d_prep_parts;
    int64_t                     opt_prep_parts = 0;  // 0 auto (chunked prep for few long curves), 1 never
    DevBuf<CurveDev>            d_curves;
    std::vector<eps_curve_info> curves;

    // sweep buffers
    DevBuf<Job>      d_jobs, d_jobs_ref;
    DevBuf<double>   d_E, d_mant, d_Elo, d_Ehi;
    DevBuf<uint32_t> d_nodes;
    DevBuf<int32_t>  d_exp;
    unsigned long long* d_steps = nullptr;

    // level-search state
    uint32_t         last_total = 0, last_nC = 0;  // (curve, level) pairs / curves of the last solve (d_levels, d_widths, d_nbelow)
    DevBuf<double>   d_lo, d_hi, d_levels, d_widths;
    DevBuf<uint32_t> d_state, d_jstar, d_nbelow, d_nactive;
    uint32_t*        h_pinned = nullptr;  // small pinned scratch (readbacks)

    // transfer-matrix (scan) path scratch + policy
    DevBuf<double>   d_segXA, d_segSA, d_segXB, d_segSB, d_fixm, d_fixE;
    DevBuf<int32_t>  d_segeA, d_segeB, d_fixe;
    DevBuf<uint32_t> d_segnA, d_fixn, d_nflag;
    DevBuf<uint2>    d_flagged;
    DevBuf<Job>      d_jobs_fix;
    uint32_t         n_tiles_max  = 0;   // over the resident curves
    int64_t          opt_scan_segments = 0;  // 0 auto, 1 never, >= 2 forced segment count
    int64_t          opt_scan_exact    = 1;  // 1: eps_solve_levels also recomputes flagged energies sequentially
    const uint32_t*  flat_rows_dev     = nullptr;  // flat refinement rows: the row count lives on the device (blind rounds)
    int              opt_scan_combine  = 0;  // 0 auto (prefix from kScanPrefixMin segments), 1 serial loop, 2 prefix always
    uint64_t         scan_launches = 0, scan_flagged = 0;

    // constant-bank sweep: host copy of the (single) curve's table + per-energy carry
    std::vector<double> h_F;
    DevBuf<double>      d_cbX, d_cbS;
    DevBuf<int32_t>     d_cbexp;
    DevBuf<uint32_t>    d_cbnodes, d_cbprev;
    int64_t             opt_cbank = 0;  // 0 auto (large single-curve sweeps), 1 always when the launch qualifies, 2 never
    int                 cb_ept = 4, cb_threads = 128, cb_pdl = 1, cb_group = 2;
    uint64_t            cbank_launches = 0;

    // wavefunction scratch
    DevBuf<double>   d_wfE, d_wfraw, d_wfin, d_wfpsi, d_wfh, d_wfdE;
    DevBuf<int32_t>  d_wfbexp, d_wfinexp;
    DevBuf<uint32_t> d_wfmatch;

    // cooperative cancellation (eps_request_stop): host flag + a device copy the sweep kernels test
    // when a CTA starts, so the launches already queued behind a stop request drain in microseconds
    std::atomic<int> stop{0};
    int*             d_stop      = nullptr;
    int*             h_stop_src  = nullptr;  // pinned {0, 1}
    cudaStream_t     stop_stream = nullptr;

    // measurement
    cudaEvent_t t0 = nullptr, t1 = nullptr;
    cudaEvent_t ev_round[2] = {nullptr, nullptr};  // refinement rounds: the active-bracket count of round r has reached the host
    cudaEvent_t ev[kEventPairs][2];
    int         ev_used = 0;
    eps_stats   stats{};
    void*       d_flush = nullptr;
    int         force_ept = 0, force_stride = 0, force_warps = 0;  // tuning overrides (EPS_FORCE_EPT / EPS_FORCE_STRIDE)
};

// Device buffers of a context.  RESIDENT: what the uploaded curves need to stay usable (tables,
// curve descriptors, small bookkeeping).  SCRATCH: grow-only work areas that any later call
// re-reserves on demand (eps_ctx_trim releases them).
#define EPS_RESIDENT_BUFS(X) \
    X(d_F) X(d_A) X(d_scale) X(d_prep) X(d_prep_parts) X(d_curves) X(d_J) X(d_rot) X(d_jobs) X(d_jobs_ref) X(d_Elo) X(d_Ehi) \
    X(d_lo) X(d_hi) X(d_levels) X(d_widths) X(d_state) X(d_jstar) X(d_nbelow) X(d_nactive) X(d_nflag)
#define EPS_SCRATCH_BUFS(X) \
    X(d_V) X(d_Vraw) X(d_spl) X(d_E) X(d_mant) X(d_nodes) X(d_exp) \
    X(d_segXA) X(d_segSA) X(d_segXB) X(d_segSB) X(d_segeA) X(d_segeB) X(d_segnA) \
    X(d_fixm) X(d_fixE) X(d_fixe) X(d_fixn) X(d_flagged) X(d_jobs_fix) \
    X(d_cbX) X(d_cbS) X(d_cbexp) X(d_cbnodes) X(d_cbprev) \
    X(d_wfE) X(d_wfraw) X(d_wfin) X(d_wfpsi) X(d_wfh) X(d_wfdE) X(d_wfbexp) X(d_wfinexp) X(d_wfmatch)

namespace {

int fail(eps_ctx* ctx, int code, const std::string& msg) {
    g_last_error = msg;
    if (ctx) ctx->err = msg;
    return code;
}

#define EPS_CUDA(ctx, call)                                                              \
    do {                                                                                 \
        cudaError_t e__ = (call);                                                        \
        if (e__ != cudaSuccess)                                                          \
            return fail(ctx, EPS_ERR_CUDA,                                               \
                        std::string(#call) + ": " + cudaGetErrorString(e__));            \
    } while (0)

#define EPS_REQUIRE(ctx, cond, code, msg) \
    do {                                  \
        if (!(cond)) return fail(ctx, code, msg); \
    } while (0)

bool stop_requested(const eps_ctx* ctx) { return ctx->stop.load(std::memory_order_acquire) != 0; }

#define EPS_CHECK_STOP(ctx)                                                                     \
    do {                                                                                        \
        if (stop_requested(ctx)) return fail(ctx, EPS_ERR_CANCELLED, "stopped by eps_request_stop"); \
    } while (0)

int bind(eps_ctx* ctx) {
    if (!ctx) return fail(nullptr, EPS_ERR_INVALID, "null context");
    EPS_CUDA(ctx, cudaSetDevice(ctx->dev));
    return EPS_OK;
}

// Fold finished sweep event pairs into stats.sweep_ms (synchronises the stream).
int fold_events(eps_ctx* ctx) {
    if (ctx->ev_used == 0) return EPS_OK;
    EPS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    const int n  = ctx->ev_used;
    ctx->ev_used = 0;  // whatever happens below, the pairs are free again
    for (int i = 0; i < n; i++) {
        float ms = 0.f;
        EPS_CUDA(ctx, cudaEventElapsedTime(&ms, ctx->ev[i][0], ctx->ev[i][1]));
        ctx->stats.sweep_ms += ms;
    }
    return EPS_OK;
}

size_t sweep_smem_bytes(int stages) { return sizeof(double) * kTile * stages + 2 * stages * sizeof(uint64_t); }

struct SweepOut {  // device pointers of one sweep's results
    uint32_t* nodes;
    double*   mant;
    int32_t*  expo;
};

template <int kEpt, int kWarps, int kStride, bool kTails, bool kScan, int kForm = 0>
cudaError_t launch_sweep_variant(eps_ctx* ctx, const Job* d_jobs, uint32_t n_jobs, uint32_t nE, const double* d_Eexp,
                                 const SweepOut& out, uint32_t n_seg, uint32_t tiles_per_seg, const SegOut& so,
                                 uint32_t pack_log2 = 0) {
    constexpr uint32_t per_cta = kScan ? kWarps * 32 : kWarps * 32 * kEpt;
    const bool     flat   = pack_log2 == kFlatRows;
    const uint64_t chunks = flat ? n_jobs : pack_log2 ? 1 : (static_cast<uint64_t>(nE) + per_cta - 1) / per_cta;
    const uint64_t grid   = flat        ? (static_cast<uint64_t>(n_jobs) * nE + per_cta - 1) / per_cta
                            : pack_log2 ? ((n_jobs + (1u << pack_log2) - 1) >> pack_log2)
                                        : chunks * n_jobs * (kScan ? n_seg : 1u);
    if (grid == 0 || grid >= (1ull << 31)) return cudaErrorInvalidConfiguration;
    auto kern = numerov_sweep_kernel<kEpt, kWarps, kStride, kTails, kScan, kForm>;
    const size_t smem = sweep_smem_bytes(sweep_stages<kEpt, kWarps>());
    static thread_local int configured_dev = -1;  // opt in to > 48 KiB dynamic smem once per device
    if (configured_dev != ctx->dev) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
        if (e != cudaSuccess) return e;
        configured_dev = ctx->dev;
    }
    ctx->stats.kernel_launches++;
    kern<<<static_cast<unsigned>(grid), (kWarps + 1) * 32, smem, ctx->stream>>>(
        kForm == 0 ? ctx->d_F.p : ctx->d_A.p, ctx->d_curves.p, d_jobs, static_cast<uint32_t>(chunks), d_Eexp, nE, out.nodes,
        kTails ? out.mant : nullptr, kTails ? out.expo : nullptr, ctx->d_steps, n_seg, tiles_per_seg, so, pack_log2, ctx->d_stop,
        pack_log2 == kFlatRows ? ctx->flat_rows_dev : nullptr);
    return cudaGetLastError();
}

template <int kEpt, int kWarps, int kStride, int kForm>
cudaError_t launch_sweep_t(eps_ctx* ctx, const Job* j, uint32_t n, uint32_t nE, const double* E, bool tails, const SweepOut& out,
                           uint32_t pack_log2) {
    const SegOut none{};
    return tails ? launch_sweep_variant<kEpt, kWarps, kStride, true, false, kForm>(ctx, j, n, nE, E, out, 1, 0, none, pack_log2)
                 : launch_sweep_variant<kEpt, kWarps, kStride, false, false, kForm>(ctx, j, n, nE, E, out, 1, 0, none, pack_log2);
}

template <int kEpt, int kWarps, int kForm = 0>
cudaError_t launch_sweep_s(eps_ctx* ctx, int stride, const Job* j, uint32_t n, uint32_t nE, const double* E, bool tails, const SweepOut& out,
                           uint32_t pack_log2) {
    if (stride == 32) return launch_sweep_t<kEpt, kWarps, 32, kForm>(ctx, j, n, nE, E, tails, out, pack_log2);
    if (stride == 8) return launch_sweep_t<kEpt, kWarps, 8, kForm>(ctx, j, n, nE, E, tails, out, pack_log2);
    return launch_sweep_t<kEpt, kWarps, 1, kForm>(ctx, j, n, nE, E, tails, out, pack_log2);
}

// CTA shape (energies per thread, consumer warps).  512 energies per CTA as 4 chains x 4 warps:
// with four chains per thread ptxas feeds 65 % of the three-register DFMAs from the operand reuse
// cache (50 % with two), measured 0.93 against 0.89 of the FP64 pipe on many-wave sweeps and a tie
// on one wave (profiles/r1_quick3_shapes.log); rows too short to fill such CTAs use 256-energy CTAs.
struct Shape { int ept, warps; };
Shape pick_shape(const eps_ctx* ctx, uint32_t n_jobs, uint32_t nE) {
    if (ctx->force_ept) return Shape{ctx->force_ept, ctx->force_warps ? ctx->force_warps : 8};
    (void)n_jobs;
    return nE > 256u ? Shape{4, 4} : Shape{2, 4};
}

// Sign-sampling stride from theta_max^2 = 12 * t_max, t_max = max over rows of s*(E_max - V_min):
// stride m is exact while m * theta_max < pi; selected with a factor-2 margin (m * theta_max < pi/2).
int pick_stride(const eps_ctx* ctx, double t_max) {
    if (ctx->force_stride) return ctx->force_stride;
    if (!(t_max == t_max)) return 1;
    if (t_max <= 2.0e-4) return 32;   // (pi/64)^2 / 12 = 2.008e-4
    if (t_max <= 3.2e-3) return 8;    // (pi/16)^2 / 12 = 3.213e-3
    return 1;
}

// Rows-per-CTA packing of short rows (2^g rows of nE <= 512 >> g energies in one 512-energy CTA);
// only for callers that lay their rows out [curve][level padded to 2^g] (pack_rows > 1).
uint32_t pack_log2_for(uint32_t nE, uint32_t pack_rows, uint32_t per_cta = 512) {
    uint32_t g = 0;
    while ((2u << g) <= pack_rows && (per_cta >> (g + 1)) >= nE) g++;
    return g;
}

cudaError_t launch_sequential(eps_ctx* ctx, const Job* d_jobs, uint32_t n_jobs, uint32_t nE, const double* d_Eexp,
                              bool tails, int stride, const SweepOut& out, uint32_t pack_log2 = 0, uint32_t pack_cta = 512) {
    // pack_cta (packed or flat rows): energies per CTA = 128 / 256 / 512 as 1 / 2 / 4 chains x 4 warps.  Packed rows:
    // what one curve's rows fill.  Flat rows: few energies are spread over all SMs with fewer chains per
    // scheduler instead of filling a few SMs with four (a sweep is latency-bound below ~3 chains).
    // 128 energies as 2 chains x 2 warps or as 1 chain x 4 warps.  With one chain per thread ptxas feeds none
    // of the three-register DFMAs from the reuse cache (9.06 modelled pipe cycles per step against 8.33,
    // scripts/sass_pipe_model.py), and a launch of at most one resident wave is faster with two (512 curves:
    // 3.09 against 3.27 ms per solve); over many waves the four-warp CTA wins (4096 curves: 20.1 against
    // 20.7 ms): profiles/r2_refine_points2.log, r2_refine_points3.log.
    if (pack_log2 && pack_cta == 128) {
        const uint64_t ctas = pack_log2 == kFlatRows ? (static_cast<uint64_t>(n_jobs) * nE + 127) / 128
                                                     : ((static_cast<uint64_t>(n_jobs) + (1u << pack_log2) - 1) >> pack_log2);
        if (ctas <= 4ull * ctx->sm_count)
            return ctx->form_resident == 1 ? launch_sweep_s<2, 2, 1>(ctx, stride, d_jobs, n_jobs, nE, d_Eexp, tails, out, pack_log2)
                                           : launch_sweep_s<2, 2, 0>(ctx, stride, d_jobs, n_jobs, nE, d_Eexp, tails, out, pack_log2);
        return ctx->form_resident == 1 ? launch_sweep_s<1, 4, 1>(ctx, stride, d_jobs, n_jobs, nE, d_Eexp, tails, out, pack_log2)
                                       : launch_sweep_s<1, 4, 0>(ctx, stride, d_jobs, n_jobs, nE, d_Eexp, tails, out, pack_log2);
    }
    if (ctx->form_resident == 1) {  // D form: the two product shapes only (the EPS_FORCE_* tuning shapes are X-form)
        if ((pack_log2 && pack_cta == 256) || (!pack_log2 && nE <= 256u))
            return launch_sweep_s<2, 4, 1>(ctx, stride, d_jobs, n_jobs, nE, d_Eexp, tails, out, pack_log2);
        return launch_sweep_s<4, 4, 1>(ctx, stride, d_jobs, n_jobs, nE, d_Eexp, tails, out, pack_log2);
    }
    if (pack_log2 && pack_cta == 256) return launch_sweep_s<2, 4>(ctx, stride, d_jobs, n_jobs, nE, d_Eexp, tails, out, pack_log2);
    if (pack_log2) return launch_sweep_s<4, 4>(ctx, stride, d_jobs, n_jobs, nE, d_Eexp, tails, out, pack_log2);
    const Shape sh = pick_shape(ctx, n_jobs, nE);
    if (sh.ept == 4 && sh.warps == 4) return launch_sweep_s<4, 4>(ctx, stride, d_jobs, n_jobs, nE, d_Eexp, tails, out, 0);
    if (sh.ept == 4) return launch_sweep_s<4, 8>(ctx, stride, d_jobs, n_jobs, nE, d_Eexp, tails, out, 0);
    if (sh.ept == 2 && sh.warps == 4) return launch_sweep_s<2, 4>(ctx, stride, d_jobs, n_jobs, nE, d_Eexp, tails, out, 0);
    if (sh.ept == 2) return launch_sweep_s<2, 8>(ctx, stride, d_jobs, n_jobs, nE, d_Eexp, tails, out, 0);
    return launch_sweep_s<1, 8>(ctx, stride, d_jobs, n_jobs, nE, d_Eexp, tails, out, 0);
}

// Constant-bank sweep (numerov_cbank.cuh): one launch per chunk of kCbChunk steps, the chunk passed
// by value through the kernel-parameter constant bank.
template <int kEpt, int kThreads, int kStride, bool kTails, int kForm = 0>
cudaError_t launch_cbank_variant(eps_ctx* ctx, const Job* d_jobs, uint32_t n_jobs, uint32_t nE, const double* d_Eexp,
                                 const SweepOut& out) {
    constexpr uint32_t per_cta = kThreads * kEpt;
    const uint64_t chunks = (static_cast<uint64_t>(nE) + per_cta - 1) / per_cta;
    const uint64_t grid   = chunks * n_jobs;
    if (grid == 0 || grid >= (1ull << 31)) return cudaErrorInvalidConfiguration;
    const uint32_t n_steps = ctx->curves[0].n_steps;
    const CbState  st{ctx->d_cbX.p, ctx->d_cbS.p, ctx->d_cbexp.p, ctx->d_cbnodes.p, ctx->d_cbprev.p};
    static thread_local FChunk chunk;  // 31 KiB staging of the by-value parameter
    // Energy groups.  The per-energy state (X, S, exponent, node count, last sign: 28 B) is written by
    // every chunk launch and read by the next.  With all 2^24 energies of C5 in every launch that is
    // 470 MB out and 470 MB in per chunk -- 45 GB of HBM traffic per sweep for 1.6 MB of table.  Running
    // the chunk launches group by group, a group being as many CTAs as keep its state (and the L2's
    // other tenants) resident, leaves the state in L2: HBM then sees the table and the node counts.
    // cb_group = CTAs per group in resident waves of the device (0: one group, the round-1 order).  Default 2
    // (2368 CTAs, 34 MB of state on a B200).  Measured on C5 with the L2 as the sweep leaves it (ncu
    // application replay, profiles/r2_cbank_traffic.log, r2_cbank_group2.log): per chunk launch 0.7 MB of
    // DRAM traffic with 2 waves, 8 MB with 3, 46 MB with 4 (ungrouped: 0.9 GB, the round-1 capture's 45 GB per
    // sweep); sweep 758.2 / 758.5 / 755.9 / 753.2 ms -- 715 instead of 52 launch boundaries cost 0.7 %.
    const int  resident_ctas = std::max(1, 2048 / kThreads > 8 ? 8 : 2048 / kThreads);  // 64 regs x 128 threads: 8 per SM
    uint64_t   group         = grid;
    if (ctx->cb_group > 0) group = std::min<uint64_t>(grid, static_cast<uint64_t>(ctx->cb_group) * resident_ctas * ctx->sm_count);
    for (uint64_t g0 = 0; g0 < grid; g0 += group) {
        const uint64_t g_ctas = std::min<uint64_t>(group, grid - g0);
        for (uint32_t k0 = 0; k0 < n_steps; k0 += kCbChunk) {
            const uint32_t len = std::min<uint32_t>(kCbChunk, n_steps - k0);
            std::memcpy(&chunk, ctx->h_F.data() + k0, len * sizeof(double));
            std::memset(reinterpret_cast<double*>(&chunk) + len, 0, (kCbChunk + 2 - len) * sizeof(double));
            cudaLaunchConfig_t cfg{};
            cfg.gridDim  = dim3(static_cast<unsigned>(g_ctas));
            cfg.blockDim = dim3(kThreads);
            cfg.stream   = ctx->stream;
            cudaLaunchAttribute attr{};
            attr.id                                         = cudaLaunchAttributeProgrammaticStreamSerialization;
            attr.val.programmaticStreamSerializationAllowed = 1;
            cfg.attrs    = &attr;
            cfg.numAttrs = (ctx->cb_pdl >= 2 || (ctx->cb_pdl == 1 && g_ctas >= 2ull * ctx->sm_count)) ? 1 : 0;
            const int pdl_late = ctx->cb_pdl > 2 ? ctx->cb_pdl - 2 : 0;  // trigger that many 128-step blocks before the chunk's end
            cudaError_t e = cudaLaunchKernelEx(&cfg, numerov_cbank_kernel<kEpt, kThreads, kStride, kTails, kForm>, chunk, d_jobs,
                                               static_cast<uint32_t>(chunks), d_Eexp, static_cast<uint64_t>(nE), ctx->curves[0].scale, len,
                                               k0 == 0 ? 1 : 0, k0 + len >= n_steps ? 1 : 0, pdl_late, static_cast<uint32_t>(g0), st,
                                               out.nodes, kTails ? out.mant : nullptr, kTails ? out.expo : nullptr, ctx->d_steps,
                                               static_cast<const int*>(ctx->d_stop));
            if (e != cudaSuccess) return e;
            ctx->cbank_launches++;
            ctx->stats.kernel_launches++;
        }
    }
    return cudaSuccess;
}

template <int kEpt, int kThreads, int kForm = 0>
cudaError_t launch_cbank_s(eps_ctx* ctx, int stride, const Job* j, uint32_t n, uint32_t nE, const double* E, bool tails, const SweepOut& out) {
    if (stride == 32) return tails ? launch_cbank_variant<kEpt, kThreads, 32, true, kForm>(ctx, j, n, nE, E, out) : launch_cbank_variant<kEpt, kThreads, 32, false, kForm>(ctx, j, n, nE, E, out);
    if (stride == 8) return tails ? launch_cbank_variant<kEpt, kThreads, 8, true, kForm>(ctx, j, n, nE, E, out) : launch_cbank_variant<kEpt, kThreads, 8, false, kForm>(ctx, j, n, nE, E, out);
    return tails ? launch_cbank_variant<kEpt, kThreads, 1, true, kForm>(ctx, j, n, nE, E, out) : launch_cbank_variant<kEpt, kThreads, 1, false, kForm>(ctx, j, n, nE, E, out);
}

// Host copy of the (single) curve's coefficient table: the constant-bank kernel takes its chunks by
// value.  Fetched once per eps_set_potentials, on the first launch that needs it.
int fetch_host_table(eps_ctx* ctx) {
    if (!ctx->h_F.empty()) return EPS_OK;
    const uint32_t n = ctx->curves[0].n_steps;
    ctx->h_F.resize(n);
    EPS_CUDA(ctx, cudaMemcpyAsync(ctx->h_F.data(), ctx->form_resident == 1 ? ctx->d_A.p : ctx->d_F.p, n * sizeof(double),
                                  cudaMemcpyDeviceToHost, ctx->stream));
    EPS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->stats.d2h_bytes += n * sizeof(double);
    return EPS_OK;
}

int launch_cbank(eps_ctx* ctx, const Job* d_jobs, uint32_t n_jobs, uint32_t nE, const double* d_Eexp, bool tails, int stride,
                 const SweepOut& out) {
    if (int rc = fetch_host_table(ctx)) return rc;
    const size_t n_out = static_cast<size_t>(n_jobs) * nE;
    EPS_CUDA(ctx, ctx->d_cbX.reserve(n_out));
    EPS_CUDA(ctx, ctx->d_cbS.reserve(n_out));
    EPS_CUDA(ctx, ctx->d_cbexp.reserve(n_out));
    EPS_CUDA(ctx, ctx->d_cbnodes.reserve(n_out));
    EPS_CUDA(ctx, ctx->d_cbprev.reserve(n_out));
    cudaError_t e;
    if (ctx->form_resident == 1) e = launch_cbank_s<4, 128, 1>(ctx, stride, d_jobs, n_jobs, nE, d_Eexp, tails, out);
    else if (ctx->cb_ept == 4 && ctx->cb_threads == 128) e = launch_cbank_s<4, 128>(ctx, stride, d_jobs, n_jobs, nE, d_Eexp, tails, out);
    else if (ctx->cb_ept == 4) e = launch_cbank_s<4, 256>(ctx, stride, d_jobs, n_jobs, nE, d_Eexp, tails, out);
    else if (ctx->cb_threads == 128) e = launch_cbank_s<2, 128>(ctx, stride, d_jobs, n_jobs, nE, d_Eexp, tails, out);
    else e = launch_cbank_s<2, 256>(ctx, stride, d_jobs, n_jobs, nE, d_Eexp, tails, out);
    EPS_CUDA(ctx, e);
    return EPS_OK;
}

// Constant-bank kernel policy: measured on the C2 table (profiles/r1_cbank.log) it overtakes the TMA
// kernel once a launch carries >= ~2 CTAs of 512 energies per SM (+1 % at 151 552 energies, +5 % at
// 303 104, +9 % at 2^20); below that the per-chunk launch overhead (26 launches per 100k steps) loses.
bool use_cbank(const eps_ctx* ctx, uint32_t n_jobs, uint32_t nE, uint32_t pack_rows, uint32_t pack_cta) {
    if (ctx->opt_cbank == 2 || ctx->nC != 1 || ctx->force_ept) return false;
    if (pack_rows != kFlatRows && pack_log2_for(nE, pack_rows, pack_cta) != 0) return false;  // (flat rows: cbank uses per-row CTAs instead)
    if (ctx->opt_cbank == 1) return true;
    return static_cast<uint64_t>(n_jobs) * nE >= 2ull * 512 * ctx->sm_count;
}

// Transfer-matrix (scan) policy.  A sweep of E_tot = n_jobs * nE energies occupies
// ceil(E_tot / 512) SMs with the sequential kernel; when that leaves most of the GPU idle on a
// long grid, the grid is cut into n_seg segments of whole tiles so that about two 256-energy CTAs
// run per SM.  The scan path executes 7 instead of 4 FP64 instructions per (energy, step) and
// gains n_seg-fold parallelism.  Returns 1 for "sequential".
constexpr uint32_t kScanMinTiles = 32;   // auto mode: grids of >= 65 536 steps only
constexpr uint32_t kScanMaxSeg   = 512;  // (64 while the combine was a serial loop only)
constexpr uint32_t kScanPrefixMin = 8;   // segments from which the combine runs as a block-level prefix
constexpr int      kScanLanes     = 16;  // segment lanes of segment_prefix_kernel
uint32_t pick_segments(const eps_ctx* ctx, uint32_t n_jobs, uint32_t nE, bool tails) {  // n_jobs: rows that carry energies
    const uint32_t n_tiles = ctx->n_tiles_max;
    if (ctx->opt_scan_segments == 1 || n_tiles < 2) return 1;
    if (ctx->opt_scan_segments >= 2)
        return static_cast<uint32_t>(std::min<int64_t>({ctx->opt_scan_segments, n_tiles, kScanMaxSeg}));
    if (tails || n_tiles < kScanMinTiles) return 1;  // tails: only the sequential march reproduces the oracle's bits
    const uint64_t e_tot = static_cast<uint64_t>(n_jobs) * nE;
    if ((e_tot + 511) / 512 >= static_cast<uint64_t>(ctx->sm_count) / 2) return 1;
    // Segment length t (whole tiles): the launch runs ceil(n_seg * ctas256 / slots) waves of t tiles on
    // the 2 * SM-count resident 256-energy CTAs; a wave costs t tile-times plus a small fixed part
    // (CTA prologue, measured ~0.03 tile).  One tile per segment often wins on granularity: C3's 489
    // tiles x 16 CTAs are 26.4 -> 27 tile-times, 18 segments of 28 tiles are 28 (1.834 against 1.899 ms,
    // profiles/r2_scan_few.log).  Ties go to the longer segment (less scratch, a shorter combine).
    const uint64_t ctas256 = static_cast<uint64_t>(n_jobs) * ((nE + 255) / 256);
    const uint64_t slots   = 2ull * ctx->sm_count;
    const uint64_t cap_mem = std::max<uint64_t>(3, (1ull << 30) / (44ull * std::max<uint64_t>(1, static_cast<uint64_t>(n_jobs) * nE)));
    double   best_cost = 0.0;
    uint32_t best_seg  = 1;
    for (uint32_t t = 1; t <= n_tiles; t++) {
        const uint64_t n_seg = (n_tiles + t - 1) / t;
        if (n_seg < 3) break;
        if (n_seg > kScanMaxSeg || n_seg > cap_mem) continue;
        const uint64_t waves = (n_seg * ctas256 + slots - 1) / slots;
        const double   cost  = static_cast<double>(waves) * (static_cast<double>(t) + 0.03);
        if (best_seg == 1 || cost <= best_cost) {
            best_cost = cost;
            best_seg  = static_cast<uint32_t>(n_seg);
        }
    }
    return best_seg;
}

int launch_scan(eps_ctx* ctx, const Job* d_jobs, uint32_t n_jobs, uint32_t nE, const double* d_Eexp, bool tails,
                int stride, uint32_t n_seg, bool fix_flagged, const SweepOut& out) {
    EPS_REQUIRE(ctx, (nE + 127) / 128 <= 65535u, EPS_ERR_INVALID, "scan path: at most 8 388 480 energies per row");
    const uint32_t tiles_per_seg = (ctx->n_tiles_max + n_seg - 1) / n_seg;
    const size_t   n_so          = static_cast<size_t>(n_jobs) * n_seg * nE;
    EPS_CUDA(ctx, ctx->d_segXA.reserve(n_so));
    EPS_CUDA(ctx, ctx->d_segSA.reserve(n_so));
    EPS_CUDA(ctx, ctx->d_segXB.reserve(n_so));
    EPS_CUDA(ctx, ctx->d_segSB.reserve(n_so));
    EPS_CUDA(ctx, ctx->d_segeA.reserve(n_so));
    EPS_CUDA(ctx, ctx->d_segeB.reserve(n_so));
    EPS_CUDA(ctx, ctx->d_segnA.reserve(n_so));
    EPS_CUDA(ctx, ctx->d_nflag.reserve(1));
    const uint32_t cap = 1u << 16;
    EPS_CUDA(ctx, ctx->d_flagged.reserve(cap));
    const SegOut so{ctx->d_segXA.p, ctx->d_segSA.p, ctx->d_segXB.p, ctx->d_segSB.p, ctx->d_segeA.p, ctx->d_segeB.p, ctx->d_segnA.p};
    cudaError_t e;
    const bool dform = ctx->form_resident == 1;
    if (stride == 32) e = dform ? launch_sweep_variant<2, 8, 32, false, true, 1>(ctx, d_jobs, n_jobs, nE, d_Eexp, out, n_seg, tiles_per_seg, so)
                                : launch_sweep_variant<2, 8, 32, false, true, 0>(ctx, d_jobs, n_jobs, nE, d_Eexp, out, n_seg, tiles_per_seg, so);
    else if (stride == 8) e = dform ? launch_sweep_variant<2, 8, 8, false, true, 1>(ctx, d_jobs, n_jobs, nE, d_Eexp, out, n_seg, tiles_per_seg, so)
                                    : launch_sweep_variant<2, 8, 8, false, true, 0>(ctx, d_jobs, n_jobs, nE, d_Eexp, out, n_seg, tiles_per_seg, so);
    else e = dform ? launch_sweep_variant<2, 8, 1, false, true, 1>(ctx, d_jobs, n_jobs, nE, d_Eexp, out, n_seg, tiles_per_seg, so)
                   : launch_sweep_variant<2, 8, 1, false, true, 0>(ctx, d_jobs, n_jobs, nE, d_Eexp, out, n_seg, tiles_per_seg, so);
    EPS_CUDA(ctx, e);
    EPS_CUDA(ctx, cudaMemsetAsync(ctx->d_nflag.p, 0, sizeof(uint32_t), ctx->stream));
    // the chain of segment matrices: a serial loop per energy for a handful of segments, the block-level
    // parallel prefix (16 segment lanes x 32 energies per CTA) from kScanPrefixMin segments on
    if ((n_seg >= kScanPrefixMin || ctx->opt_scan_combine == 2) && ctx->opt_scan_combine != 1 && (nE + 31) / 32 <= 65535u)
        segment_prefix_kernel<kScanLanes><<<dim3(n_jobs, (nE + 31) / 32), dim3(32, kScanLanes), 0, ctx->stream>>>(
            so, d_jobs, n_jobs, n_seg, nE, out.nodes, tails ? out.mant : nullptr, tails ? out.expo : nullptr,
            ctx->d_nflag.p, ctx->d_flagged.p, cap, dform ? 1.0 : -1.0);
    else
        segment_combine_kernel<<<dim3(n_jobs, (nE + 127) / 128), 128, 0, ctx->stream>>>(
            so, d_jobs, n_jobs, n_seg, nE, out.nodes, tails ? out.mant : nullptr, tails ? out.expo : nullptr,
            ctx->d_nflag.p, ctx->d_flagged.p, cap, dform ? 1.0 : -1.0);
    EPS_CUDA(ctx, cudaGetLastError());
    ctx->stats.other_launches++;
    ctx->scan_launches++;
    if (!fix_flagged) return EPS_OK;
    EPS_CUDA(ctx, cudaMemcpyAsync(ctx->h_pinned + 8, ctx->d_nflag.p, sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
    EPS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->stats.d2h_bytes += sizeof(uint32_t);
    const uint32_t n_flag = ctx->h_pinned[8];
    if (n_flag == 0) return EPS_OK;
    ctx->scan_flagged += n_flag;
    if (n_flag > cap) {  // pathological: every energy ill-conditioned -> redo the whole sweep sequentially
        EPS_CUDA(ctx, launch_sequential(ctx, d_jobs, n_jobs, nE, d_Eexp, tails, stride, out));
        return EPS_OK;
    }
    EPS_CUDA(ctx, ctx->d_jobs_fix.reserve(n_flag));
    EPS_CUDA(ctx, ctx->d_fixn.reserve(n_flag));
    EPS_CUDA(ctx, ctx->d_fixm.reserve(n_flag));
    EPS_CUDA(ctx, ctx->d_fixe.reserve(n_flag));
    EPS_CUDA(ctx, ctx->d_fixE.reserve(n_flag));
    // one curve resident: pack the flagged energies into ONE explicit-energy row (512 per CTA);
    // several curves: one single-energy row per flagged item (rare, small)
    const bool packed = ctx->nC == 1;
    make_fixup_jobs_kernel<<<(n_flag + 127) / 128, 128, 0, ctx->stream>>>(d_jobs, d_Eexp, ctx->d_flagged.p, n_flag, packed ? 1 : 0,
                                                                         ctx->d_jobs_fix.p, ctx->d_fixE.p);
    EPS_CUDA(ctx, cudaGetLastError());
    const SweepOut fix{ctx->d_fixn.p, ctx->d_fixm.p, ctx->d_fixe.p};
    if (packed) EPS_CUDA(ctx, launch_sequential(ctx, ctx->d_jobs_fix.p, 1, n_flag, ctx->d_fixE.p, tails, stride, fix));
    else EPS_CUDA(ctx, launch_sequential(ctx, ctx->d_jobs_fix.p, n_flag, 1, nullptr, tails, stride, fix));
    scatter_fixup_kernel<<<(n_flag + 127) / 128, 128, 0, ctx->stream>>>(ctx->d_flagged.p, n_flag, nE, ctx->d_fixn.p, ctx->d_fixm.p, ctx->d_fixe.p,
                                                                       out.nodes, tails ? out.mant : nullptr, tails ? out.expo : nullptr);
    EPS_CUDA(ctx, cudaGetLastError());
    ctx->stats.other_launches += 2;
    return EPS_OK;
}

// Launch the sweep over n_jobs rows of nE energies each (jobs already on device).
// t_max = max over the rows of s * (E_max - V_min) (see pick_stride).  fix_flagged: on the scan
// path, recompute ill-conditioned energies with the sequential kernel (one host sync).
int launch_sweep(eps_ctx* ctx, const Job* d_jobs, uint32_t n_jobs, uint32_t nE, const double* d_Eexp,
                 bool tails, double t_max, bool fix_flagged = true, uint32_t n_rows_active = 0, uint32_t pack_rows = 1,
                 uint32_t pack_cta = 512, bool allow_scan = true) {
    const size_t n_out = static_cast<size_t>(n_jobs) * nE;
    EPS_CUDA(ctx, ctx->d_nodes.reserve(n_out));
    if (tails) {
        EPS_CUDA(ctx, ctx->d_mant.reserve(n_out));
        EPS_CUDA(ctx, ctx->d_exp.reserve(n_out));
    }
    const uint32_t n_seg = (allow_scan || ctx->opt_scan_segments >= 2) ? pick_segments(ctx, n_rows_active ? n_rows_active : n_jobs, nE, tails) : 1u;
    if (n_seg >= 2) {  // reserve before the timed region
        const size_t n_so = n_out * n_seg;
        EPS_CUDA(ctx, ctx->d_segXA.reserve(n_so));
        EPS_CUDA(ctx, ctx->d_segSA.reserve(n_so));
        EPS_CUDA(ctx, ctx->d_segXB.reserve(n_so));
        EPS_CUDA(ctx, ctx->d_segSB.reserve(n_so));
        EPS_CUDA(ctx, ctx->d_segeA.reserve(n_so));
        EPS_CUDA(ctx, ctx->d_segeB.reserve(n_so));
        EPS_CUDA(ctx, ctx->d_segnA.reserve(n_so));
    }
    if (ctx->ev_used == kEventPairs) {
        int rc = fold_events(ctx);
        if (rc) return rc;
    }
    cudaEvent_t* pair = ctx->ev[ctx->ev_used];
    EPS_CUDA(ctx, cudaEventRecord(pair[0], ctx->stream));
    const int      stride = pick_stride(ctx, t_max);
    const SweepOut out{ctx->d_nodes.p, ctx->d_mant.p, ctx->d_exp.p};
    // the pair only counts once BOTH events are recorded: a failure in between leaves ev_used alone
    int rc = EPS_OK;
    if (n_seg >= 2) {
        rc = launch_scan(ctx, d_jobs, n_jobs, nE, d_Eexp, tails, stride, n_seg, fix_flagged, out);
    } else if (use_cbank(ctx, n_jobs, nE, pack_rows, pack_cta)) {
        rc = launch_cbank(ctx, d_jobs, n_jobs, nE, d_Eexp, tails, stride, out);
    } else {
        const uint32_t pk = ctx->force_ept ? 0 : pack_rows == kFlatRows ? kFlatRows : pack_log2_for(nE, pack_rows, pack_cta);
        const cudaError_t e = launch_sequential(ctx, d_jobs, n_jobs, nE, d_Eexp, tails, stride, out, pk, pack_cta);
        if (e != cudaSuccess) rc = fail(ctx, EPS_ERR_CUDA, std::string("sweep launch: ") + cudaGetErrorString(e));
    }
    if (rc) return rc;
    EPS_CUDA(ctx, cudaEventRecord(pair[1], ctx->stream));
    ctx->ev_used++;
    ctx->stats.sweep_launches++;
    return EPS_OK;
}

// t_max of a curve for trial energies up to E_max (see pick_stride).
double curve_tmax(const eps_curve_info& ci, double E_max) { return ci.scale * (E_max - ci.v_min); }

// Validity of a trial-energy range on a curve: |s (E - V_min)| <= T_MAX keeps every
// f_i = 1 - (q_i - e) positive inside the window (Sturm property, bounded growth).
bool range_ok(const eps_curve_info& ci, double E_a, double E_b) {
    const double a = ci.scale * (E_a - ci.v_min), b = ci.scale * (E_b - ci.v_min);
    return std::isfinite(a) && std::isfinite(b) && std::fabs(a) <= kTMax && std::fabs(b) <= kTMax;
}

int fetch_sweep(eps_ctx* ctx, size_t n, uint32_t* nodes, double* mant, int32_t* expo) {
    if (nodes) {
        EPS_CUDA(ctx, cudaMemcpyAsync(nodes, ctx->d_nodes.p, n * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
        ctx->stats.d2h_bytes += n * sizeof(uint32_t);
    }
    if (mant) {
        EPS_CUDA(ctx, cudaMemcpyAsync(mant, ctx->d_mant.p, n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        ctx->stats.d2h_bytes += n * sizeof(double);
    }
    if (expo) {
        EPS_CUDA(ctx, cudaMemcpyAsync(expo, ctx->d_exp.p, n * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
        ctx->stats.d2h_bytes += n * sizeof(int32_t);
    }
    if (nodes || mant || expo) {
        EPS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        EPS_CHECK_STOP(ctx);
    }
    return EPS_OK;
}

// The resident state is invalid from the moment a new upload starts until it has fully succeeded:
// DevBuf::reserve frees before it allocates and the prep kernels overwrite d_F / d_curves, so a
// failure half-way must not leave the old curve count and curve_info behind.
void invalidate_resident(eps_ctx* ctx) {
    ctx->nC = 0;
    ctx->curves.clear();
    ctx->h_F.clear();
    ctx->n_tiles_max = 0;
}

// Second half of eps_set_potentials*: the tables are in d_V, the scales in d_scale (both on the
// device already); prep_curves_kernel derives window + coefficient table per curve (spec
// DESIGN.md section 3.2) and 32 B per curve come back.
template <typename ScaleOf>
int prep_resident(eps_ctx* ctx, uint32_t n_curves, uint32_t n_points, ScaleOf scale_of) {
    const uint32_t N    = n_points;
    const uint64_t slot = (static_cast<uint64_t>(N) + kTile - 1) / kTile * kTile;
    const size_t   n_f  = static_cast<size_t>(slot) * n_curves;
    invalidate_resident(ctx);
    EPS_CUDA(ctx, ctx->d_prep.reserve(n_curves));
    EPS_CUDA(ctx, ctx->d_F.reserve(n_f));
    const int form = ctx->opt_form;
    if (form == 1) EPS_CUDA(ctx, ctx->d_A.reserve(n_f));
    double* d_A = form == 1 ? ctx->d_A.p : nullptr;
    EPS_CUDA(ctx, ctx->d_curves.reserve(n_curves));
    // few long curves: cut every curve into chunks (one CTA each) instead of one CTA per curve
    const uint32_t parts = (N >= 65536 && n_curves <= 64) ? std::min<uint32_t>(kPrepPartsMax, (N + 8191) / 8192) : 1;
    if (parts > 1 && ctx->opt_prep_parts != 1) {
        EPS_CUDA(ctx, ctx->d_prep_parts.reserve(static_cast<size_t>(parts) * n_curves));
        const dim3 grid(parts, n_curves);
        prep_part_argmin_kernel<<<grid, kPrepThreads, 0, ctx->stream>>>(ctx->d_V.p, ctx->d_scale.p, N, parts, ctx->d_prep_parts.p);
        prep_part_window_kernel<<<grid, kPrepThreads, 0, ctx->stream>>>(ctx->d_V.p, ctx->d_scale.p, N, parts, kTMax, ctx->d_prep_parts.p);
        prep_part_finish_kernel<<<grid, kPrepThreads, 0, ctx->stream>>>(ctx->d_V.p, ctx->d_scale.p, N, slot, parts, ctx->d_prep_parts.p,
                                                                       ctx->d_F.p, d_A, ctx->d_curves.p, ctx->d_prep.p);
        EPS_CUDA(ctx, cudaGetLastError());
        ctx->stats.other_launches += 3;
    } else {
        prep_curves_kernel<<<n_curves, kPrepThreads, 0, ctx->stream>>>(ctx->d_V.p, ctx->d_scale.p, N, slot, kTMax, ctx->d_F.p, d_A, ctx->d_curves.p, ctx->d_prep.p);
        EPS_CUDA(ctx, cudaGetLastError());
        ctx->stats.other_launches++;
    }
    std::vector<PrepOut> po(n_curves);
    EPS_CUDA(ctx, cudaMemcpyAsync(po.data(), ctx->d_prep.p, n_curves * sizeof(PrepOut), cudaMemcpyDeviceToHost, ctx->stream));
    EPS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->stats.d2h_bytes += n_curves * sizeof(PrepOut);
    for (uint32_t c = 0; c < n_curves; c++) {
        EPS_REQUIRE(ctx, po[c].status != 1, EPS_ERR_RANGE, "potential table holds a non-finite value");
        EPS_REQUIRE(ctx, po[c].status != 2, EPS_ERR_RANGE, "integration window has fewer than 2 steps");
    }
    std::vector<eps_curve_info> infos(n_curves);
    for (uint32_t c = 0; c < n_curves; c++) infos[c] = eps_curve_info{po[c].i0, po[c].n_steps, scale_of(c), po[c].v_min, po[c].v_last};
    ctx->nC     = n_curves;
    ctx->form_resident = form;
    ctx->N      = N;
    ctx->slot   = slot;
    ctx->curves = std::move(infos);
    ctx->h_F.clear();  // host copy for the constant-bank kernel: fetched on first use (fetch_host_table)
    ctx->n_tiles_max = 0;
    for (const auto& ci : ctx->curves) ctx->n_tiles_max = std::max(ctx->n_tiles_max, (ci.n_steps + kTile - 1) / kTile);
    return EPS_OK;
}

}  // namespace

extern "C" {

int eps_abi_version(void) { return EPS_ABI_VERSION; }

const char* eps_last_error(const eps_ctx* ctx) { return ctx ? ctx->err.c_str() : g_last_error.c_str(); }

int eps_device_count(int* count) {
    if (!count) return fail(nullptr, EPS_ERR_INVALID, "count is null");
    *count = 0;
    int         n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return fail(nullptr, EPS_ERR_CUDA, std::string("cudaGetDeviceCount: ") + cudaGetErrorString(e));
    }
    *count = n;
    return EPS_OK;
}

int eps_device_get_props(int device, eps_device_props* out) {
    if (!out) return fail(nullptr, EPS_ERR_INVALID, "out is null");
    cudaDeviceProp p;
    EPS_CUDA(nullptr, cudaGetDeviceProperties(&p, device));
    std::memset(out, 0, sizeof(*out));
    std::snprintf(out->name, sizeof(out->name), "%s", p.name);
    out->ordinal  = device;
    out->cc_major = p.major;
    out->cc_minor = p.minor;
    out->sm_count = p.multiProcessorCount;
    int khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, device);
    out->clock_khz = khz;
    cudaDriverGetVersion(&out->driver_version);
    cudaRuntimeGetVersion(&out->runtime_version);
    out->pci_domain            = p.pciDomainID;
    out->pci_bus               = p.pciBusID;
    out->pci_device            = p.pciDeviceID;
    out->integrated            = p.integrated;
    out->max_threads_per_block = p.maxThreadsPerBlock;
    for (int i = 0; i < 3; i++) {
        out->max_grid[i]  = p.maxGridSize[i];
        out->max_block[i] = p.maxThreadsDim[i];
    }
    out->l2_bytes                   = p.l2CacheSize;
    out->total_global_mem           = p.totalGlobalMem;
    out->shared_mem_per_block_optin = p.sharedMemPerBlockOptin;
    out->shared_mem_per_sm          = p.sharedMemPerMultiprocessor;
    std::memcpy(out->uuid, p.uuid.bytes, 16);
    return EPS_OK;
}

int eps_ctx_create(int device, eps_ctx** out) {
    if (!out) return fail(nullptr, EPS_ERR_INVALID, "out is null");
    *out  = nullptr;
    int n = 0, rc = eps_device_count(&n);
    if (rc) return rc;
    if (device < 0 || device >= n) return fail(nullptr, EPS_ERR_INVALID, "no such CUDA device");
    eps_ctx* ctx = new (std::nothrow) eps_ctx();
    if (!ctx) return fail(nullptr, EPS_ERR_NOMEM, "out of host memory");
    ctx->dev = device;
    auto bail = [&](cudaError_t e, const char* what) {
        std::string m = std::string(what) + ": " + cudaGetErrorString(e);
        eps_ctx_destroy(ctx);
        return fail(nullptr, EPS_ERR_CUDA, m);
    };
    cudaError_t e;
    if ((e = cudaSetDevice(device)) != cudaSuccess) return bail(e, "cudaSetDevice");
    int cc_major = 0;
    cudaDeviceGetAttribute(&cc_major, cudaDevAttrComputeCapabilityMajor, device);
    cudaDeviceGetAttribute(&ctx->sm_count, cudaDevAttrMultiProcessorCount, device);
    if (cc_major < 10) {
        eps_ctx_destroy(ctx);
        return fail(nullptr, EPS_ERR_CUDA, "device is not sm_100-class (kernels are built for sm_100a only)");
    }
    if ((e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking)) != cudaSuccess) return bail(e, "cudaStreamCreate");
    if ((e = cudaEventCreate(&ctx->t0)) != cudaSuccess) return bail(e, "cudaEventCreate");
    if ((e = cudaEventCreate(&ctx->t1)) != cudaSuccess) return bail(e, "cudaEventCreate");
    for (auto& evt : ctx->ev_round)
        if ((e = cudaEventCreateWithFlags(&evt, cudaEventDisableTiming)) != cudaSuccess) return bail(e, "cudaEventCreate");
    for (auto& pr : ctx->ev)
        for (auto& evt : pr)
            if ((e = cudaEventCreate(&evt)) != cudaSuccess) return bail(e, "cudaEventCreate");
    if ((e = cudaMalloc(reinterpret_cast<void**>(&ctx->d_steps), sizeof(unsigned long long))) != cudaSuccess) return bail(e, "cudaMalloc");
    if ((e = cudaMemsetAsync(ctx->d_steps, 0, sizeof(unsigned long long), ctx->stream)) != cudaSuccess) return bail(e, "cudaMemset");
    if ((e = cudaMallocHost(reinterpret_cast<void**>(&ctx->h_pinned), 4096)) != cudaSuccess) return bail(e, "cudaMallocHost");
    if ((e = cudaMalloc(reinterpret_cast<void**>(&ctx->d_stop), sizeof(int))) != cudaSuccess) return bail(e, "cudaMalloc");
    if ((e = cudaMemsetAsync(ctx->d_stop, 0, sizeof(int), ctx->stream)) != cudaSuccess) return bail(e, "cudaMemset");
    if ((e = cudaMallocHost(reinterpret_cast<void**>(&ctx->h_stop_src), 2 * sizeof(int))) != cudaSuccess) return bail(e, "cudaMallocHost");
    ctx->h_stop_src[0] = 0;
    ctx->h_stop_src[1] = 1;
    if ((e = cudaStreamCreateWithFlags(&ctx->stop_stream, cudaStreamNonBlocking)) != cudaSuccess) return bail(e, "cudaStreamCreate");
    if (const char* fe = std::getenv("EPS_FORCE_EPT")) {
        const int v = std::atoi(fe);
        ctx->force_ept = (v == 1 || v == 2 || v == 4) ? v : 0;
    }
    if (const char* fw = std::getenv("EPS_FORCE_WARPS")) ctx->force_warps = std::atoi(fw) == 4 ? 4 : 8;
    if (const char* fs = std::getenv("EPS_FORCE_STRIDE")) {
        const int v = std::atoi(fs);
        ctx->force_stride = (v == 1 || v == 8 || v == 32) ? v : 0;
    }
    *out = ctx;
    return EPS_OK;
}

int eps_ctx_destroy(eps_ctx* ctx) {
    if (!ctx) return EPS_OK;
    if (ctx->dev >= 0 && cudaSetDevice(ctx->dev) == cudaSuccess) {
        if (ctx->stream) cudaStreamSynchronize(ctx->stream);
#define EPS_RELEASE(name) ctx->name.release();
        EPS_RESIDENT_BUFS(EPS_RELEASE)
        EPS_SCRATCH_BUFS(EPS_RELEASE)
#undef EPS_RELEASE
        if (ctx->d_steps) cudaFree(ctx->d_steps);
        if (ctx->d_flush) cudaFree(ctx->d_flush);
        if (ctx->h_pinned) cudaFreeHost(ctx->h_pinned);
        if (ctx->stop_stream) {
            cudaStreamSynchronize(ctx->stop_stream);
            cudaStreamDestroy(ctx->stop_stream);
        }
        if (ctx->d_stop) cudaFree(ctx->d_stop);
        if (ctx->h_stop_src) cudaFreeHost(ctx->h_stop_src);
        if (ctx->t0) cudaEventDestroy(ctx->t0);
        if (ctx->t1) cudaEventDestroy(ctx->t1);
        for (auto evt : ctx->ev_round)
            if (evt) cudaEventDestroy(evt);
        for (auto& pr : ctx->ev)
            for (auto& evt : pr)
                if (evt) cudaEventDestroy(evt);
        if (ctx->stream) cudaStreamDestroy(ctx->stream);
    }
    delete ctx;
    return EPS_OK;
}

// Cancellation.  eps_request_stop may be called from ANY thread while another thread is inside a
// compute call on the same context (it is what TaskHandle::cancel binds, reference
// task_handle.hpp:136-144): it raises the host flag -- tested by eps_solve_levels* between
// refinement rounds and by every sweep before it returns -- and copies it to the device on a side
// stream, where every sweep CTA tests it on entry, so sweeps already queued (the 51 chunk launches
// of a 2^24-energy sweep) drain at once.  The interrupted call returns EPS_ERR_CANCELLED and its
// outputs are unspecified.  The flag stays up until eps_reset_stop.
int eps_request_stop(eps_ctx* ctx) {
    if (!ctx) return fail(nullptr, EPS_ERR_INVALID, "null context");
    ctx->stop.store(1, std::memory_order_release);
    if (cudaSetDevice(ctx->dev) == cudaSuccess && ctx->d_stop && ctx->stop_stream)
        cudaMemcpyAsync(ctx->d_stop, ctx->h_stop_src + 1, sizeof(int), cudaMemcpyHostToDevice, ctx->stop_stream);
    return EPS_OK;
}

int eps_reset_stop(eps_ctx* ctx) {
    if (int rc = bind(ctx)) return rc;
    EPS_CUDA(ctx, cudaStreamSynchronize(ctx->stop_stream));
    ctx->stop.store(0, std::memory_order_release);
    EPS_CUDA(ctx, cudaMemcpyAsync(ctx->d_stop, ctx->h_stop_src, sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    return EPS_OK;
}

// Memory held by a context's grow-only device buffers, and a way to give it back: after one large
// or wavefunction task a context keeps GBs of scratch.  eps_ctx_trim releases every scratch buffer
// (the next call re-reserves what it needs); with drop_potentials != 0 also the resident tables
// (the context then needs a new eps_set_potentials*).
int eps_ctx_device_bytes(eps_ctx* ctx, uint64_t* bytes) {
    if (!ctx || !bytes) return fail(ctx, EPS_ERR_INVALID, "null argument");
    uint64_t n = 0;
#define EPS_COUNT(name) n += static_cast<uint64_t>(ctx->name.cap) * sizeof(*ctx->name.p);
    EPS_RESIDENT_BUFS(EPS_COUNT)
    EPS_SCRATCH_BUFS(EPS_COUNT)
#undef EPS_COUNT
    if (ctx->d_flush) n += kFlushBytes;
    *bytes = n;
    return EPS_OK;
}

int eps_ctx_trim(eps_ctx* ctx, int drop_potentials) {
    if (int rc = bind(ctx)) return rc;
    EPS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
#define EPS_RELEASE(name) ctx->name.release();
    EPS_SCRATCH_BUFS(EPS_RELEASE)
    if (drop_potentials) {
        invalidate_resident(ctx);
        EPS_RESIDENT_BUFS(EPS_RELEASE)
    }
#undef EPS_RELEASE
    if (ctx->d_flush) {
        cudaFree(ctx->d_flush);
        ctx->d_flush = nullptr;
    }
    return EPS_OK;
}

int eps_sync(eps_ctx* ctx) {
    if (int rc = bind(ctx)) return rc;
    EPS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return EPS_OK;
}

// Preparation (spec DESIGN.md section 3.2): q = s V, window [i0, iend] around the
// minimum with q - q_min <= T_MAX, coefficient table F_k = (1 - q_{i0+k}) / 12 -- on the device:
// the raw table goes up once, prep_resident() does the rest.
int eps_set_potentials(eps_ctx* ctx, const double* V, uint32_t n_curves, uint32_t n_points,
                       const double* scale) {
    if (int rc = bind(ctx)) return rc;
    EPS_REQUIRE(ctx, V && scale, EPS_ERR_INVALID, "V/scale is null");
    EPS_REQUIRE(ctx, n_curves >= 1 && n_points >= 3, EPS_ERR_INVALID, "need >=1 curve of >=3 points");
    for (uint32_t c = 0; c < n_curves; c++)
        EPS_REQUIRE(ctx, std::isfinite(scale[c]) && scale[c] > 0.0, EPS_ERR_INVALID, "scale must be finite and positive");
    const size_t n_v = static_cast<size_t>(n_points) * n_curves;
    invalidate_resident(ctx);
    EPS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    EPS_CUDA(ctx, ctx->d_V.reserve(n_v));
    EPS_CUDA(ctx, ctx->d_scale.reserve(n_curves));
    EPS_CUDA(ctx, cudaMemcpyAsync(ctx->d_V.p, V, n_v * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    EPS_CUDA(ctx, cudaMemcpyAsync(ctx->d_scale.p, scale, n_curves * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    ctx->stats.h2d_bytes += (n_v + n_curves) * sizeof(double);
    return prep_resident(ctx, n_curves, n_points, [&](uint32_t c) { return scale[c]; });
}

int eps_set_potentials_rot(eps_ctx* ctx, const double* V, uint32_t n_curves, uint32_t n_points, const double* scale,
                           const double* r_min, const double* grid_step, const uint32_t* J, uint32_t n_J) {
    if (int rc = bind(ctx)) return rc;
    EPS_REQUIRE(ctx, V && scale && r_min && grid_step && J, EPS_ERR_INVALID, "null argument");
    EPS_REQUIRE(ctx, n_curves >= 1 && n_points >= 3 && n_J >= 1, EPS_ERR_INVALID, "need >=1 curve of >=3 points and >=1 rotational state");
    EPS_REQUIRE(ctx, static_cast<uint64_t>(n_curves) * n_J <= 65535u * 64u, EPS_ERR_INVALID, "too many (curve, J) pairs");
    for (uint32_t c = 0; c < n_curves; c++) {
        EPS_REQUIRE(ctx, std::isfinite(scale[c]) && scale[c] > 0.0, EPS_ERR_INVALID, "scale must be finite and positive");
        EPS_REQUIRE(ctx, std::isfinite(r_min[c]) && std::isfinite(grid_step[c]) && grid_step[c] > 0.0, EPS_ERR_INVALID,
                    "r_min must be finite and grid_step finite and positive");
    }
    for (uint32_t j = 0; j < n_J; j++) EPS_REQUIRE(ctx, J[j] < (1u << 26), EPS_ERR_INVALID, "J must be below 2^26");
    const uint32_t n_eff = n_curves * n_J;
    const size_t   n_raw = static_cast<size_t>(n_points) * n_curves, n_v = static_cast<size_t>(n_points) * n_eff;
    invalidate_resident(ctx);
    EPS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    EPS_CUDA(ctx, ctx->d_Vraw.reserve(n_raw));
    EPS_CUDA(ctx, ctx->d_rot.reserve(3 * static_cast<size_t>(n_curves)));
    EPS_CUDA(ctx, ctx->d_J.reserve(n_J));
    EPS_CUDA(ctx, ctx->d_V.reserve(n_v));
    EPS_CUDA(ctx, ctx->d_scale.reserve(n_eff));
    double* d_s = ctx->d_rot.p, *d_r0 = d_s + n_curves, *d_h = d_r0 + n_curves;
    EPS_CUDA(ctx, cudaMemcpyAsync(ctx->d_Vraw.p, V, n_raw * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    EPS_CUDA(ctx, cudaMemcpyAsync(d_s, scale, n_curves * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    EPS_CUDA(ctx, cudaMemcpyAsync(d_r0, r_min, n_curves * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    EPS_CUDA(ctx, cudaMemcpyAsync(d_h, grid_step, n_curves * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    EPS_CUDA(ctx, cudaMemcpyAsync(ctx->d_J.p, J, n_J * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream));
    ctx->stats.h2d_bytes += (n_raw + 3 * static_cast<size_t>(n_curves)) * sizeof(double) + n_J * sizeof(uint32_t);
    const uint32_t bx = std::min<uint32_t>((n_points + 255) / 256, 64u);
    centrifugal_kernel<<<dim3(n_eff, bx), 256, 0, ctx->stream>>>(ctx->d_Vraw.p, d_s, d_r0, d_h, ctx->d_J.p, n_J, n_points,
                                                                 ctx->d_V.p, ctx->d_scale.p);
    EPS_CUDA(ctx, cudaGetLastError());
    ctx->stats.other_launches++;
    return prep_resident(ctx, n_eff, n_points, [&](uint32_t c) { return scale[c / n_J]; });
}

int eps_get_curve_info(eps_ctx* ctx, uint32_t curve, eps_curve_info* out) {
    if (!ctx) return fail(nullptr, EPS_ERR_INVALID, "null context");
    EPS_REQUIRE(ctx, out, EPS_ERR_INVALID, "out is null");
    EPS_REQUIRE(ctx, curve < ctx->nC, EPS_ERR_INVALID, "no such curve");
    *out = ctx->curves[curve];
    return EPS_OK;
}

int eps_sweep(eps_ctx* ctx, const double* E, uint64_t n_energies, uint32_t* nodes, double* tail_mant,
              int32_t* tail_exp) {
    if (int rc = bind(ctx)) return rc;
    EPS_REQUIRE(ctx, ctx->nC > 0, EPS_ERR_STATE, "eps_set_potentials has not been called");
    EPS_REQUIRE(ctx, E && n_energies >= 1 && n_energies < (1ull << 32), EPS_ERR_INVALID, "bad energies");
    const uint32_t nE = static_cast<uint32_t>(n_energies);
    std::vector<Job> jobs(ctx->nC);
    double           t_max = -1.0;
    for (uint32_t c = 0; c < ctx->nC; c++) {
        const double* row = E + static_cast<size_t>(c) * nE;
        double        mn = row[0], mx = row[0];
        for (uint32_t j = 1; j < nE; j++) {
            mn = std::min(mn, row[j]);
            mx = std::max(mx, row[j]);
        }
        bool finite = true;
        for (uint32_t j = 0; j < nE && finite; j++) finite = std::isfinite(row[j]);
        EPS_REQUIRE(ctx, finite && range_ok(ctx->curves[c], mn, mx), EPS_ERR_RANGE,
                    "trial energy outside the validity window |s (E - V_min)| <= 0.5 (grid too coarse)");
        t_max   = std::max(t_max, curve_tmax(ctx->curves[c], mx));
        jobs[c] = Job{0.0, 0.0, static_cast<uint64_t>(c) * nE, c, 0, nE, 0, c, 0};
    }
    const size_t n = static_cast<size_t>(ctx->nC) * nE;
    EPS_CUDA(ctx, ctx->d_E.reserve(n));
    EPS_CUDA(ctx, ctx->d_jobs.reserve(ctx->nC));
    EPS_CUDA(ctx, cudaMemcpyAsync(ctx->d_E.p, E, n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    EPS_CUDA(ctx, cudaMemcpyAsync(ctx->d_jobs.p, jobs.data(), jobs.size() * sizeof(Job), cudaMemcpyHostToDevice, ctx->stream));
    ctx->stats.h2d_bytes += n * sizeof(double) + jobs.size() * sizeof(Job);
    // jobs lives on the host stack of this call: make sure the copy is done before returning
    if (int rc = launch_sweep(ctx, ctx->d_jobs.p, ctx->nC, nE, ctx->d_E.p, tail_mant || tail_exp, t_max)) return rc;
    EPS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return fetch_sweep(ctx, n, nodes, tail_mant, tail_exp);
}

}  // extern "C" (reopened below)

namespace {

// Coarse grid of one call: E_j = A[c] + (j0 + j) * B[c] (grid) or A = E_lo, B = E_hi (uniform).
struct GridSpec {
    const double* A;
    const double* B;
    bool          grid;
    uint32_t      j0;
};

// First / last trial energy of row c, for the validity check and the stride choice.
void grid_ends(const GridSpec& g, uint32_t c, uint32_t nE, double& e_first, double& e_last) {
    if (!g.grid) {
        e_first = g.A[c];
        e_last  = g.B[c];
    } else {
        e_first = g.A[c] + static_cast<double>(g.j0) * g.B[c];
        e_last  = g.A[c] + static_cast<double>(static_cast<uint64_t>(g.j0) + nE - 1) * g.B[c];
    }
}

int upload_coarse_jobs(eps_ctx* ctx, const GridSpec& g, uint32_t nE, double& t_max) {
    const uint32_t nC = ctx->nC;
    t_max = -1.0;
    for (uint32_t c = 0; c < nC; c++) {
        double a, b;
        grid_ends(g, c, nE, a, b);
        EPS_REQUIRE(ctx, range_ok(ctx->curves[c], a, b), EPS_ERR_RANGE,
                    "trial energy outside the validity window |s (E - V_min)| <= 0.5 (grid too coarse)");
        t_max = std::max(t_max, curve_tmax(ctx->curves[c], std::max(a, b)));
    }
    EPS_CUDA(ctx, ctx->d_Elo.reserve(nC));
    EPS_CUDA(ctx, ctx->d_Ehi.reserve(nC));
    EPS_CUDA(ctx, ctx->d_jobs.reserve(nC));
    // A/B are pageable host memory: the async copies are staged before the call returns
    EPS_CUDA(ctx, cudaMemcpyAsync(ctx->d_Elo.p, g.A, nC * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    EPS_CUDA(ctx, cudaMemcpyAsync(ctx->d_Ehi.p, g.B, nC * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    ctx->stats.h2d_bytes += 2ull * nC * sizeof(double);
    make_coarse_jobs_kernel<<<(nC + 127) / 128, 128, 0, ctx->stream>>>(ctx->d_Elo.p, ctx->d_Ehi.p, nC, nE, g.grid ? 1 : 0, g.j0, ctx->d_jobs.p);
    EPS_CUDA(ctx, cudaGetLastError());
    ctx->stats.other_launches++;
    return EPS_OK;
}

int sweep_rows(eps_ctx* ctx, const GridSpec& g, uint64_t n_energies, uint32_t* nodes, double* tail_mant, int32_t* tail_exp) {
    if (int rc = bind(ctx)) return rc;
    EPS_REQUIRE(ctx, ctx->nC > 0, EPS_ERR_STATE, "eps_set_potentials has not been called");
    EPS_REQUIRE(ctx, g.A && g.B && n_energies >= 1 && n_energies < (1ull << 32), EPS_ERR_INVALID, "bad energies");
    const uint32_t nE = static_cast<uint32_t>(n_energies);
    double         t_max;
    if (int rc = upload_coarse_jobs(ctx, g, nE, t_max)) return rc;
    if (int rc = launch_sweep(ctx, ctx->d_jobs.p, ctx->nC, nE, nullptr, tail_mant || tail_exp, t_max)) return rc;
    return fetch_sweep(ctx, static_cast<size_t>(ctx->nC) * nE, nodes, tail_mant, tail_exp);
}

int solve_rows(eps_ctx* ctx, const eps_solve_params* p, const GridSpec& g, double* levels, double* widths,
               uint32_t* n_below, uint32_t* n_first) {
    if (int rc = bind(ctx)) return rc;
    EPS_REQUIRE(ctx, ctx->nC > 0, EPS_ERR_STATE, "eps_set_potentials has not been called");
    EPS_REQUIRE(ctx, p && g.A && g.B, EPS_ERR_INVALID, "null argument");  // levels == NULL: results stay on the device
    EPS_REQUIRE(ctx, p->v_max >= p->v_min, EPS_ERR_INVALID, "v_max < v_min");
    EPS_REQUIRE(ctx, p->n_coarse >= 2 && p->refine_points >= 1, EPS_ERR_INVALID, "n_coarse >= 2 and refine_points >= 1 required");
    const uint32_t nC = ctx->nC, nlev = p->v_max - p->v_min + 1, M = p->refine_points;
    const uint64_t total64 = static_cast<uint64_t>(nC) * nlev;
    EPS_REQUIRE(ctx, total64 < (1ull << 31), EPS_ERR_INVALID, "too many (curve, level) pairs");
    const uint32_t total = static_cast<uint32_t>(total64);
    const uint32_t nE    = p->n_coarse;
    for (uint32_t c = 0; c < nC; c++) {
        double a, b;
        grid_ends(g, c, nE, a, b);
        EPS_REQUIRE(ctx, b >= a, EPS_ERR_INVALID, "E_hi < E_lo");
    }
    EPS_CUDA(ctx, ctx->d_lo.reserve(total));
    EPS_CUDA(ctx, ctx->d_hi.reserve(total));
    EPS_CUDA(ctx, ctx->d_levels.reserve(total));
    EPS_CUDA(ctx, ctx->d_widths.reserve(total));
    EPS_CUDA(ctx, ctx->d_state.reserve(total));
    EPS_CUDA(ctx, ctx->d_jstar.reserve(total));
    EPS_CUDA(ctx, ctx->d_nbelow.reserve(2ull * nC));
    EPS_CUDA(ctx, ctx->d_nactive.reserve(1));

    // ---- coarse sweep + bracketing ----
    double t_max;
    if (int rc = upload_coarse_jobs(ctx, g, nE, t_max)) return rc;
    if (int rc = launch_sweep(ctx, ctx->d_jobs.p, nC, nE, nullptr, false, t_max, ctx->opt_scan_exact != 0)) return rc;
    EPS_CUDA(ctx, cudaMemsetAsync(ctx->d_jstar.p, 0xff, total * sizeof(uint32_t), ctx->stream));
    {
        const uint32_t bpr = (nE + 255) / 256;
        crossing_kernel<<<bpr * nC, 256, 0, ctx->stream>>>(ctx->d_nodes.p, nE, nE, bpr, ctx->d_jobs.p, 1, p->v_min, nlev, ctx->d_jstar.p);
        EPS_CUDA(ctx, cudaGetLastError());
        bracket_init_kernel<<<(total + 127) / 128, 128, 0, ctx->stream>>>(ctx->d_nodes.p, nE, ctx->d_jobs.p, ctx->d_jstar.p, nC, p->v_min, nlev,
                                                                         ctx->d_lo.p, ctx->d_hi.p, ctx->d_state.p, ctx->d_nbelow.p,
                                                                         ctx->d_nbelow.p + nC);
        EPS_CUDA(ctx, cudaGetLastError());
        ctx->stats.other_launches += 2;
    }
    // ---- Cooley (outward/inward matching) iteration instead of k-section sweeps ----
    bool cooley = (p->flags & EPS_SOLVE_COOLEY) != 0;
    if (cooley) {
        EPS_REQUIRE(ctx, ctx->form_resident == 1, EPS_ERR_STATE, "EPS_SOLVE_COOLEY needs the accurate tables (EPS_OPT_FORM = 1 before eps_set_potentials)");
        for (const auto& ci : ctx->curves) cooley = cooley && ci.n_steps >= 256u;  // short windows: k-section
    }
    if (cooley) {
        // one thread per segment of the longest window, rounded up to whole warps
        uint32_t seg_max = 1;
        for (const auto& ci : ctx->curves) {
            const uint32_t L = cooley_segment_length(ci.n_steps, total);
            seg_max          = std::max(seg_max, (ci.n_steps + L - 1) / L);
        }
        const uint32_t threads = std::min<uint32_t>(kCooleyThreads, (seg_max + 31u) / 32u * 32u);
        cooley_search_kernel<<<total, threads, 0, ctx->stream>>>(ctx->d_A.p, ctx->d_curves.p, nlev, p->v_min, p->rel_tol,
                                                                       std::max<uint32_t>(p->max_rounds, 1u), ctx->d_lo.p, ctx->d_hi.p,
                                                                       ctx->d_state.p, ctx->d_levels.p, ctx->d_widths.p, nullptr,
                                                                       ctx->d_steps, ctx->d_stop, (p->flags & EPS_SOLVE_OPEN_TAIL) ? 1 : 0);
        EPS_CUDA(ctx, cudaGetLastError());
        ctx->stats.other_launches++;
    }
    // ---- k-section refinement rounds ----
    // One curve resident: the active brackets are compacted into rows that are swept with the
    // flat-row mapping (full CTAs).  Several curves: dense rows [curve][level padded to the CTA
    // packing] so that the 2^g short rows packed into a CTA share a curve.
    // CTAs of 256 instead of 512 energies when the rows of one curve (nlev x M energies) would leave
    // half of a 512-energy CTA idle: a round then costs half as much, so fewer points per level per
    // round (less total work for the same final width) pay off on many-curve batches.
    const bool     flat      = nC == 1 && !ctx->force_ept;
    const uint32_t rows512   = 1u << pack_log2_for(M, 512, 512);
    uint32_t       pack_cta  = (!flat && M <= 128 && nlev <= rows512 / 2) ? 256u : 512u;
    // Balance of small per-device batches.  Equal CTAs over 148 SMs leave a tail: 512 curves in
    // 256-energy CTAs (2 resident per SM) are 3.46 per SM -> two rounds of 2, i.e. 4 units of SM time
    // for 3.46 of work.  128-energy CTAs (1 chain x 4 warps or 2 x 2, 4 resident per SM: the same four chains
    // per scheduler) halve the unit: 6.92 per SM -> a round of 4 and one of 3.  Cost model in units of
    // (one chain per scheduler x the grid): full rounds cost `resident x chains`, the last one what is
    // left; the smaller shape is taken when it saves more than 5 %.
    if (pack_cta == 256 && M <= 64 && ctx->opt_pack128 != 2) {
        const auto cost = [&](uint64_t ctas, uint32_t resident, uint32_t chains) {
            const uint64_t per_round = static_cast<uint64_t>(ctx->sm_count) * resident;
            const uint64_t rem       = ctas % per_round;
            return static_cast<double>(ctas / per_round) * resident * chains +
                   static_cast<double>((rem + ctx->sm_count - 1) / ctx->sm_count) * chains;
        };
        const uint32_t rows256 = 1u << pack_log2_for(M, 512, 256), rows128 = 1u << pack_log2_for(M, 512, 128);
        const uint64_t c256 = static_cast<uint64_t>(nC) * ((nlev + rows256 - 1) / rows256);
        const uint64_t c128 = static_cast<uint64_t>(nC) * ((nlev + rows128 - 1) / rows128);
        if (ctx->opt_pack128 == 1 || cost(c128, 4, 1) < 0.95 * cost(c256, 2, 2)) pack_cta = 128u;
    }
    const uint32_t pack_rows = flat ? kFlatRows : 1u << pack_log2_for(M, 512, pack_cta);
    const uint32_t nlev_pad  = flat ? nlev : (nlev + pack_rows - 1) / pack_rows * pack_rows;
    const uint32_t n_dense   = nC * nlev_pad;
    EPS_CUDA(ctx, ctx->d_jobs_ref.reserve(n_dense));
    EPS_CUDA(ctx, ctx->d_jstar.reserve(std::max(n_dense, total)));
    // Rounds without the host.  A round's launches do not need the number of open brackets on the host:
    // dense rows (several curves) idle where nothing is left (nE = 0, their CTAs leave at once), and flat
    // rows (one curve) read the row count from the device (flat_rows_dev; the grid is sized for all rows).
    // The first `blind` rounds are therefore enqueued back to back, before the coarse sweep has even
    // finished -- `blind` = the rounds the WIDEST tolerance of the job certainly needs, from the coarse
    // spacing, rel_tol and M -- so a host thread that is descheduled for a few milliseconds (shared
    // hosts: observed as rare 15 ms .. 1 s spikes in 4 ms solves that synced every round) no longer
    // idles the device.  After them the count is read: dense rows one round late (the host stays a
    // round ahead), flat rows before every further round (rare: brackets that needed one more).
    uint32_t blind = 0;
    if (!cooley && ctx->opt_scan_exact != 0 && ctx->opt_scan_segments < 2) {  // (the scan path's fix-up reads back per launch anyway)
        double need = 1e300;  // min over the curves of log(spacing / (rel_tol |E|max)) / log(M + 1)
        for (uint32_t c = 0; c < nC; c++) {
            double a, b;
            grid_ends(g, c, nE, a, b);
            const double spacing = (b - a) / static_cast<double>(nE - 1), mag = std::max(std::fabs(a), std::fabs(b));
            const double tol = p->rel_tol * mag;
            need = std::min(need, (tol > 0.0 && spacing > tol) ? std::log(spacing / tol) / std::log(static_cast<double>(M) + 1.0) : 0.0);
        }
        blind = static_cast<uint32_t>(std::min<double>(p->max_rounds, std::ceil(std::max(need, 0.0))));
    }
    uint32_t n_prev_pending = 0;  // dense rows: a count copy of the previous round is in flight
    for (uint32_t round = 0; round < p->max_rounds && !cooley; round++) {
        const bool is_blind = round < blind;
        if (flat) {
            compact_refine_jobs_kernel<<<1, 1024, 0, ctx->stream>>>(ctx->d_lo.p, ctx->d_hi.p, ctx->d_state.p, total, nlev, p->v_min, p->rel_tol, M,
                                                                  ctx->d_jobs_ref.p, ctx->d_nactive.p);
        } else {
            EPS_CUDA(ctx, cudaMemsetAsync(ctx->d_nactive.p, 0, sizeof(uint32_t), ctx->stream));
            make_refine_jobs_kernel<<<(n_dense + 127) / 128, 128, 0, ctx->stream>>>(ctx->d_lo.p, ctx->d_hi.p, ctx->d_state.p, nC, nlev, nlev_pad, p->v_min,
                                                                                   p->rel_tol, M, ctx->d_jobs_ref.p, ctx->d_nactive.p);
        }
        EPS_CUDA(ctx, cudaGetLastError());
        ctx->stats.other_launches++;
        uint32_t n_active = flat ? total : n_dense;  // upper bound while the count is not read
        if (!is_blind) {
            EPS_CUDA(ctx, cudaMemcpyAsync(ctx->h_pinned + 16 + (round & 1), ctx->d_nactive.p, sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
            ctx->stats.d2h_bytes += sizeof(uint32_t);
            if (flat || (blind > 0 && round == blind)) {  // (after the blind rounds: normally the end -- look before another sweep)
                EPS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
                n_active = ctx->h_pinned[16 + (round & 1)];
                if (!flat && n_active != 0) n_active = n_dense;
            } else {
                EPS_CUDA(ctx, cudaEventRecord(ctx->ev_round[round & 1], ctx->stream));
                if (n_prev_pending) {
                    EPS_CUDA(ctx, cudaEventSynchronize(ctx->ev_round[(round - 1) & 1]));
                    if (ctx->h_pinned[16 + ((round - 1) & 1)] == 0) { n_prev_pending = 0; break; }
                }
                n_prev_pending = 1;
            }
            EPS_CHECK_STOP(ctx);
            if (n_active == 0) break;
        }
        const uint32_t n_rows = flat ? n_active : n_dense;
        uint32_t       cta    = pack_cta;
        if (flat) {  // one curve: the round's energies over all SMs, with as few chains per scheduler as that takes
            const uint64_t e = static_cast<uint64_t>(n_active) * M;
            cta              = e <= 128ull * ctx->sm_count ? 128u : (e <= 256ull * ctx->sm_count ? 256u : 512u);
        }
        // Refinement energies crowd the eigenvalues, where the scan path's tails are cancellations: with the
        // exact fix-up on, a good part of every round is flagged and marched again sequentially, so the
        // scan can only add to the round (C2 with a third round: 9.2 instead of 4.9 ms).  It is left to the
        // coarse sweep, and to refinement in the fast mode (EPS_OPT_SCAN_EXACT = 0).
        ctx->flat_rows_dev = (flat && is_blind) ? ctx->d_nactive.p : nullptr;
        const int rc_sweep = launch_sweep(ctx, ctx->d_jobs_ref.p, n_rows, M, nullptr, false, t_max, ctx->opt_scan_exact != 0, n_active, pack_rows,
                                          cta, ctx->opt_scan_exact == 0);
        ctx->flat_rows_dev = nullptr;
        if (rc_sweep) return rc_sweep;
        EPS_CUDA(ctx, cudaMemsetAsync(ctx->d_jstar.p, 0xff, n_rows * sizeof(uint32_t), ctx->stream));
        const uint32_t bpr = (M + 255) / 256;
        crossing_kernel<<<bpr * n_rows, 256, 0, ctx->stream>>>(ctx->d_nodes.p, M, M, bpr, ctx->d_jobs_ref.p, 0, 0, 1, ctx->d_jstar.p);
        EPS_CUDA(ctx, cudaGetLastError());
        bracket_update_kernel<<<(n_rows + 127) / 128, 128, 0, ctx->stream>>>(ctx->d_jobs_ref.p, ctx->d_jstar.p, n_rows, M, ctx->d_lo.p, ctx->d_hi.p, ctx->d_state.p);
        EPS_CUDA(ctx, cudaGetLastError());
        ctx->stats.other_launches += 2;
    }
    if (!cooley) {
        finalize_levels_kernel<<<(total + 127) / 128, 128, 0, ctx->stream>>>(ctx->d_lo.p, ctx->d_hi.p, total, ctx->d_levels.p, ctx->d_widths.p);
        EPS_CUDA(ctx, cudaGetLastError());
        ctx->stats.other_launches++;
    }
    ctx->last_total = total;
    ctx->last_nC    = nC;
    if (levels) {
        EPS_CUDA(ctx, cudaMemcpyAsync(levels, ctx->d_levels.p, total * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        ctx->stats.d2h_bytes += total * sizeof(double);
    }
    if (widths) {
        EPS_CUDA(ctx, cudaMemcpyAsync(widths, ctx->d_widths.p, total * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        ctx->stats.d2h_bytes += total * sizeof(double);
    }
    if (n_below) {
        EPS_CUDA(ctx, cudaMemcpyAsync(n_below, ctx->d_nbelow.p, nC * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
        ctx->stats.d2h_bytes += nC * sizeof(uint32_t);
    }
    if (n_first) {
        EPS_CUDA(ctx, cudaMemcpyAsync(n_first, ctx->d_nbelow.p + nC, nC * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
        ctx->stats.d2h_bytes += nC * sizeof(uint32_t);
    }
    EPS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    EPS_CHECK_STOP(ctx);
    return EPS_OK;
}

}  // namespace

extern "C" {

int eps_sweep_uniform(eps_ctx* ctx, const double* E_lo, const double* E_hi, uint64_t n_energies,
                      uint32_t* nodes, double* tail_mant, int32_t* tail_exp) {
    return sweep_rows(ctx, GridSpec{E_lo, E_hi, false, 0}, n_energies, nodes, tail_mant, tail_exp);
}

int eps_sweep_grid(eps_ctx* ctx, const double* E0, const double* dE, uint32_t j0, uint64_t n_energies,
                   uint32_t* nodes, double* tail_mant, int32_t* tail_exp) {
    if (ctx && n_energies + j0 >= (1ull << 32)) return fail(ctx, EPS_ERR_INVALID, "j0 + n_energies must stay below 2^32");
    return sweep_rows(ctx, GridSpec{E0, dE, true, j0}, n_energies, nodes, tail_mant, tail_exp);
}

int eps_solve_levels(eps_ctx* ctx, const eps_solve_params* p, const double* E_lo, const double* E_hi,
                     double* levels, double* widths, uint32_t* n_below) {
    return solve_rows(ctx, p, GridSpec{E_lo, E_hi, false, 0}, levels, widths, n_below, nullptr);
}

int eps_solve_levels_grid(eps_ctx* ctx, const eps_solve_params* p, const double* E0, const double* dE,
                          uint32_t j0, double* levels, double* widths, uint32_t* n_last, uint32_t* n_first) {
    if (ctx && p && static_cast<uint64_t>(p->n_coarse) + j0 >= (1ull << 32))
        return fail(ctx, EPS_ERR_INVALID, "j0 + n_coarse must stay below 2^32");
    return solve_rows(ctx, p, GridSpec{E0, dE, true, j0}, levels, widths, n_last, n_first);
}

}  // extern "C" (reopened below)

namespace {

// Validates, uploads E / grid_step and runs the two wavefunction kernels; the normalised psi stays
// in ctx->d_wfpsi ([item][n_points]), the window-relative matching indices in ctx->d_wfmatch.
int run_wavefunctions(eps_ctx* ctx, const double* E, uint32_t n_levels, const double* grid_step, uint32_t& items_out) {
    EPS_REQUIRE(ctx, ctx->nC > 0, EPS_ERR_STATE, "eps_set_potentials has not been called");
    EPS_REQUIRE(ctx, E && grid_step && n_levels >= 1, EPS_ERR_INVALID, "null argument / no levels");
    const uint64_t items64 = static_cast<uint64_t>(ctx->nC) * n_levels;
    EPS_REQUIRE(ctx, items64 < 65536ull * 16, EPS_ERR_INVALID, "too many (curve, level) pairs in one call");
    const uint32_t items = static_cast<uint32_t>(items64);
    for (uint32_t c = 0; c < ctx->nC; c++) {
        EPS_REQUIRE(ctx, std::isfinite(grid_step[c]) && grid_step[c] > 0.0, EPS_ERR_INVALID, "grid_step must be positive");
        for (uint32_t l = 0; l < n_levels; l++) {
            const double e = E[static_cast<size_t>(c) * n_levels + l];
            if (e != e) continue;
            EPS_REQUIRE(ctx, range_ok(ctx->curves[c], e, e), EPS_ERR_RANGE,
                        "level energy outside the validity window |s (E - V_min)| <= 0.5");
        }
    }
    const uint32_t bstride = static_cast<uint32_t>(ctx->slot / kRenorm) + 8;
    const size_t   n_psi   = static_cast<size_t>(items) * ctx->N;
    EPS_REQUIRE(ctx, n_psi * sizeof(double) <= (64ull << 30), EPS_ERR_NOMEM, "wavefunction output above 64 GiB: split the call");
    EPS_CUDA(ctx, ctx->d_wfE.reserve(items));
    EPS_CUDA(ctx, ctx->d_wfh.reserve(ctx->nC));
    EPS_CUDA(ctx, ctx->d_wfraw.reserve(static_cast<size_t>(items) * ctx->slot));
    EPS_CUDA(ctx, ctx->d_wfbexp.reserve(static_cast<size_t>(items) * 2 * bstride));
    EPS_CUDA(ctx, ctx->d_wfmatch.reserve(items));
    EPS_CUDA(ctx, ctx->d_wfin.reserve(items));
    EPS_CUDA(ctx, ctx->d_wfinexp.reserve(items));
    EPS_CUDA(ctx, ctx->d_wfpsi.reserve(n_psi));
    EPS_CUDA(ctx, cudaMemcpyAsync(ctx->d_wfE.p, E, items * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    EPS_CUDA(ctx, cudaMemcpyAsync(ctx->d_wfh.p, grid_step, ctx->nC * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    ctx->stats.h2d_bytes += (items + static_cast<uint64_t>(ctx->nC)) * sizeof(double);
    wavefunction_march_kernel<<<dim3(items, 2), kWfThreads, 0, ctx->stream>>>(
        ctx->d_F.p, ctx->d_curves.p, ctx->d_wfE.p, n_levels, ctx->slot, bstride, ctx->d_wfraw.p, ctx->d_wfbexp.p,
        ctx->d_wfmatch.p, ctx->d_wfin.p, ctx->d_wfinexp.p);
    EPS_CUDA(ctx, cudaGetLastError());
    wavefunction_finish_kernel<<<items, kWfThreads, 0, ctx->stream>>>(
        ctx->d_curves.p, n_levels, ctx->N, ctx->d_wfh.p, ctx->slot, bstride, ctx->d_wfraw.p, ctx->d_wfbexp.p,
        ctx->d_wfmatch.p, ctx->d_wfin.p, ctx->d_wfinexp.p, ctx->d_wfpsi.p);
    EPS_CUDA(ctx, cudaGetLastError());
    ctx->stats.other_launches += 2;
    items_out = items;
    return EPS_OK;
}

}  // namespace

extern "C" {

int eps_wavefunctions(eps_ctx* ctx, const double* E, uint32_t n_levels, const double* grid_step,
                      double* psi, uint32_t* match_index) {
    if (int rc = bind(ctx)) return rc;
    EPS_REQUIRE(ctx, psi, EPS_ERR_INVALID, "psi is null");
    uint32_t items = 0;
    if (int rc = run_wavefunctions(ctx, E, n_levels, grid_step, items)) return rc;
    const size_t n_psi = static_cast<size_t>(items) * ctx->N;
    EPS_CUDA(ctx, cudaMemcpyAsync(psi, ctx->d_wfpsi.p, n_psi * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    ctx->stats.d2h_bytes += n_psi * sizeof(double);
    if (match_index) {
        EPS_CUDA(ctx, cudaMemcpyAsync(match_index, ctx->d_wfmatch.p, items * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
        ctx->stats.d2h_bytes += items * sizeof(uint32_t);
    }
    EPS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (match_index)  // report the matching point as an index of the full r grid
        for (uint32_t it = 0; it < items; it++)
            if (match_index[it] != kNone) match_index[it] += ctx->curves[it / n_levels].i0;
    return EPS_OK;
}

int eps_level_corrections(eps_ctx* ctx, const double* E, uint32_t n_levels, const double* grid_step, double* dE) {
    if (int rc = bind(ctx)) return rc;
    EPS_REQUIRE(ctx, dE, EPS_ERR_INVALID, "dE is null");
    uint32_t items = 0;
    if (int rc = run_wavefunctions(ctx, E, n_levels, grid_step, items)) return rc;
    EPS_CUDA(ctx, ctx->d_wfdE.reserve(items));
    level_correction_kernel<<<items, kWfThreads, 0, ctx->stream>>>(ctx->d_F.p, ctx->d_curves.p, ctx->d_wfE.p, n_levels, ctx->N,
                                                                  ctx->d_wfmatch.p, ctx->d_wfpsi.p, ctx->d_wfdE.p);
    EPS_CUDA(ctx, cudaGetLastError());
    ctx->stats.other_launches++;
    EPS_CUDA(ctx, cudaMemcpyAsync(dE, ctx->d_wfdE.p, items * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    ctx->stats.d2h_bytes += items * sizeof(double);
    EPS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return EPS_OK;
}

// Natural cubic spline coefficients (a, b, c, d per knot interval); host arithmetic, same operation
// order as the oracle counterpart; this file is compiled with -ffp-contract=off.
int eps_spline_coefficients(const double* r, const double* V, uint32_t K, double* coef) {
    if (!r || !V || !coef) return fail(nullptr, EPS_ERR_INVALID, "null argument");
    if (K < 3) return fail(nullptr, EPS_ERR_INVALID, "a cubic spline needs at least 3 knots");
    for (uint32_t k = 0; k < K; k++)
        if (!std::isfinite(r[k]) || !std::isfinite(V[k])) return fail(nullptr, EPS_ERR_RANGE, "non-finite knot");
    for (uint32_t k = 0; k + 1 < K; k++)
        if (!(r[k + 1] > r[k])) return fail(nullptr, EPS_ERR_INVALID, "knot abscissae must be strictly increasing");
    std::vector<double> m(K, 0.0), cp(K, 0.0), dp(K, 0.0);
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
    return EPS_OK;
}

int eps_spline_resample(eps_ctx* ctx, const double* r, const double* V, uint32_t n_knots, double r_min,
                        double r_max, uint32_t n_points, double* V_out) {
    if (int rc = bind(ctx)) return rc;
    EPS_REQUIRE(ctx, V_out && n_points >= 2 && std::isfinite(r_min) && std::isfinite(r_max) && r_max > r_min, EPS_ERR_INVALID,
                "bad output grid");
    std::vector<double> coef(4 * static_cast<size_t>(n_knots > 1 ? n_knots - 1 : 1));
    if (int rc = eps_spline_coefficients(r, V, n_knots, coef.data())) {
        ctx->err = g_last_error;
        return rc;
    }
    EPS_CUDA(ctx, ctx->d_V.reserve(n_points));
    EPS_CUDA(ctx, ctx->d_spl.reserve(coef.size() + n_knots));
    EPS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    EPS_CUDA(ctx, cudaMemcpyAsync(ctx->d_spl.p, r, n_knots * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    EPS_CUDA(ctx, cudaMemcpyAsync(ctx->d_spl.p + n_knots, coef.data(), coef.size() * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    ctx->stats.h2d_bytes += (n_knots + coef.size()) * sizeof(double);
    const double h = (r_max - r_min) / static_cast<double>(n_points - 1);
    spline_eval_kernel<<<(n_points + 255) / 256, 256, 0, ctx->stream>>>(ctx->d_spl.p, ctx->d_spl.p + n_knots, n_knots, r_min, h, n_points, ctx->d_V.p);
    EPS_CUDA(ctx, cudaGetLastError());
    ctx->stats.other_launches++;
    EPS_CUDA(ctx, cudaMemcpyAsync(V_out, ctx->d_V.p, n_points * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    ctx->stats.d2h_bytes += n_points * sizeof(double);
    EPS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return EPS_OK;
}

int eps_set_option(eps_ctx* ctx, int option, int64_t value) {
    if (!ctx) return fail(nullptr, EPS_ERR_INVALID, "null context");
    switch (option) {
        case EPS_OPT_SCAN_SEGMENTS:
            EPS_REQUIRE(ctx, value >= 0 && value <= 4096, EPS_ERR_INVALID, "scan segments: 0 (auto), 1 (off) or a count");
            ctx->opt_scan_segments = value;
            return EPS_OK;
        case EPS_OPT_SCAN_EXACT:
            ctx->opt_scan_exact = value != 0;
            return EPS_OK;
        case EPS_OPT_CBANK:
            EPS_REQUIRE(ctx, value >= 0 && value <= 2, EPS_ERR_INVALID, "cbank: 0 auto, 1 always, 2 never");
            ctx->opt_cbank = value;
            return EPS_OK;
        case EPS_OPT_FORM:
            EPS_REQUIRE(ctx, value == 0 || value == 1, EPS_ERR_INVALID, "form: 0 (X form, 4 operations) or 1 (D form, accurate, 5 operations)");
            ctx->opt_form = static_cast<int>(value);  // takes effect at the next eps_set_potentials*
            return EPS_OK;
        case EPS_OPT_PACK128:
            EPS_REQUIRE(ctx, value >= 0 && value <= 2, EPS_ERR_INVALID, "pack128: 0 auto, 1 always, 2 never");
            ctx->opt_pack128 = static_cast<int>(value);
            return EPS_OK;
        case EPS_OPT_PREP_PARTS:
            EPS_REQUIRE(ctx, value == 0 || value == 1, EPS_ERR_INVALID, "prep parts: 0 auto, 1 never");
            ctx->opt_prep_parts = value;
            return EPS_OK;
        case EPS_OPT_CBANK_PDL:
            ctx->cb_pdl = value < 0 ? 0 : value > 18 ? 18 : static_cast<int>(value);
            return EPS_OK;
        case EPS_OPT_SCAN_COMBINE:
            EPS_REQUIRE(ctx, value >= 0 && value <= 2, EPS_ERR_INVALID, "scan combine: 0 auto, 1 serial, 2 prefix");
            ctx->opt_scan_combine = static_cast<int>(value);
            return EPS_OK;
        case EPS_OPT_CBANK_GROUP:
            ctx->cb_group = value < 0 ? 0 : value > 4096 ? 4096 : static_cast<int>(value);
            return EPS_OK;
        case EPS_OPT_CBANK_SHAPE:  // tuning: energies per thread * 1000 + threads per CTA
            EPS_REQUIRE(ctx, (value / 1000 == 2 || value / 1000 == 4) && (value % 1000 == 128 || value % 1000 == 256), EPS_ERR_INVALID,
                        "cbank shape: (2|4)*1000 + (128|256)");
            ctx->cb_ept     = static_cast<int>(value / 1000);
            ctx->cb_threads = static_cast<int>(value % 1000);
            return EPS_OK;
        default: return fail(ctx, EPS_ERR_INVALID, "unknown option");
    }
}

int eps_get_counter(eps_ctx* ctx, int counter, uint64_t* value) {
    if (!ctx || !value) return fail(ctx, EPS_ERR_INVALID, "null argument");
    switch (counter) {
        case EPS_CNT_SCAN_LAUNCHES: *value = ctx->scan_launches; return EPS_OK;
        case EPS_CNT_SCAN_FLAGGED: *value = ctx->scan_flagged; return EPS_OK;
        case EPS_CNT_CBANK_LAUNCHES: *value = ctx->cbank_launches; return EPS_OK;
        default: return fail(ctx, EPS_ERR_INVALID, "unknown counter");
    }
}

int eps_timer_start(eps_ctx* ctx) {
    if (int rc = bind(ctx)) return rc;
    EPS_CUDA(ctx, cudaEventRecord(ctx->t0, ctx->stream));
    return EPS_OK;
}

int eps_timer_stop(eps_ctx* ctx, float* ms) {
    if (int rc = bind(ctx)) return rc;
    EPS_REQUIRE(ctx, ms, EPS_ERR_INVALID, "ms is null");
    EPS_CUDA(ctx, cudaEventRecord(ctx->t1, ctx->stream));
    EPS_CUDA(ctx, cudaEventSynchronize(ctx->t1));
    EPS_CUDA(ctx, cudaEventElapsedTime(ms, ctx->t0, ctx->t1));
    return EPS_OK;
}

int eps_stats_get(eps_ctx* ctx, eps_stats* out) {
    if (int rc = bind(ctx)) return rc;
    EPS_REQUIRE(ctx, out, EPS_ERR_INVALID, "out is null");
    if (int rc = fold_events(ctx)) return rc;
    unsigned long long steps = 0;
    EPS_CUDA(ctx, cudaMemcpyAsync(&steps, ctx->d_steps, sizeof(steps), cudaMemcpyDeviceToHost, ctx->stream));
    EPS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->stats.grid_steps = steps;
    *out                  = ctx->stats;
    out->kernel_launches += ctx->stats.other_launches;  // every kernel this library launched: sweeps (each chunk) + the rest
    return EPS_OK;
}

int eps_stats_reset(eps_ctx* ctx) {
    if (int rc = bind(ctx)) return rc;
    if (int rc = fold_events(ctx)) return rc;
    EPS_CUDA(ctx, cudaMemsetAsync(ctx->d_steps, 0, sizeof(unsigned long long), ctx->stream));
    EPS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->stats = eps_stats{};
    return EPS_OK;
}

int eps_host_alloc(eps_ctx* ctx, size_t bytes, void** out) {
    if (int rc = bind(ctx)) return rc;
    EPS_REQUIRE(ctx, out && bytes > 0, EPS_ERR_INVALID, "out is null or bytes == 0");
    *out = nullptr;
    EPS_CUDA(ctx, cudaHostAlloc(out, bytes, cudaHostAllocPortable));
    return EPS_OK;
}

int eps_host_free(eps_ctx* ctx, void* p) {
    if (!p) return EPS_OK;
    const cudaError_t e = cudaFreeHost(p);
    if (e != cudaSuccess) return fail(ctx, EPS_ERR_CUDA, cudaGetErrorString(e));
    return EPS_OK;
}

int eps_l2_flush(eps_ctx* ctx) {
    if (int rc = bind(ctx)) return rc;
    if (!ctx->d_flush) EPS_CUDA(ctx, cudaMalloc(&ctx->d_flush, kFlushBytes));
    EPS_CUDA(ctx, cudaMemsetAsync(ctx->d_flush, 0x5a, kFlushBytes, ctx->stream));
    return EPS_OK;
}

int eps_fp64_probe(eps_ctx* ctx, double* tflops, float* ms_out) {
    if (int rc = bind(ctx)) return rc;
    EPS_REQUIRE(ctx, tflops, EPS_ERR_INVALID, "tflops is null");
    const int    blocks = ctx->sm_count * 8, threads = 256, iters = kProbeIters;
    DevBuf<double> out;
    EPS_CUDA(ctx, out.reserve(static_cast<size_t>(blocks) * threads));
    float best = 1e30f;
    for (int rep = 0; rep < 12; rep++) {
        EPS_CUDA(ctx, cudaEventRecord(ctx->t0, ctx->stream));
        fp64_probe_kernel<<<blocks, threads, 0, ctx->stream>>>(out.p, 1.0000001, 1e-9);
        EPS_CUDA(ctx, cudaGetLastError());
        EPS_CUDA(ctx, cudaEventRecord(ctx->t1, ctx->stream));
        EPS_CUDA(ctx, cudaEventSynchronize(ctx->t1));
        float ms = 0.f;
        EPS_CUDA(ctx, cudaEventElapsedTime(&ms, ctx->t0, ctx->t1));
        if (rep > 0) best = std::min(best, ms);
    }
    out.release();
    const double flops = 2.0 * 64.0 * iters * static_cast<double>(blocks) * threads;
    *tflops            = flops / (best * 1e-3) / 1e12;
    if (ms_out) *ms_out = best;
    return EPS_OK;
}

}  // extern "C"

// =====================================================================================================
// Multi-device group: one process, one host thread and one eps_ctx per device.  The path
// shards without any exchange during compute -- by potential curve or by energy range -- and only the
// located levels travel: every device copies its level array into a gather buffer on the group's
// first device with cudaMemcpyPeerAsync (NVLink / NVSwitch when peer access is available), device 0
// waits on the peers' events and ONE device->host copy returns the lot; the per-level merge of an
// energy-sharded search (each level is finite on exactly one device) runs on the host.
// No torch, no NCCL: the payload is a few hundred bytes per device.
// =====================================================================================================
struct eps_group {
    struct Worker {
        eps_ctx*                ctx = nullptr;
        std::thread             th;
        std::mutex              m;
        std::condition_variable cv;
        std::function<int()>    job;
        bool                    has_job = false, done = false, quit = false;
        int                     rc = EPS_OK;
        float                   ms = 0.f;
        cudaEvent_t             posted = nullptr;  // levels of the last solve have landed in the gather buffer
    };
    std::vector<Worker*> w;
    std::string          err;
    int                  shard = EPS_SHARD_CURVES;
    uint32_t             nC = 0, N = 0;          // whole job
    std::vector<uint32_t> c0, cn;                 // curve block of every device (curve sharding)
    double*              d_gather = nullptr;      // on device w[0]: [n_dev][2 * total + pad] levels + widths
    size_t               gather_cap = 0;          // doubles per device slot
    double*              h_gather = nullptr;      // pinned mirror
    size_t               h_cap = 0;
    float                last_ms = 0.f;
};

namespace {

int gfail(eps_group* g, int code, const std::string& msg) {
    g_last_error = msg;
    if (g) g->err = msg;
    return code;
}

void worker_loop(eps_group::Worker* w) {
    cudaSetDevice(w->ctx->dev);
    for (;;) {
        std::function<int()> job;
        {
            std::unique_lock<std::mutex> lk(w->m);
            w->cv.wait(lk, [&] { return w->has_job || w->quit; });
            if (w->quit) return;
            job        = std::move(w->job);
            w->has_job = false;
        }
        eps_ctx* ctx = w->ctx;
        int      rc  = eps_timer_start(ctx);
        if (rc == EPS_OK) rc = job();
        float ms = 0.f;
        if (rc == EPS_OK) rc = eps_timer_stop(ctx, &ms);
        {
            std::lock_guard<std::mutex> lk(w->m);
            w->rc   = rc;
            w->ms   = ms;
            w->done = true;
        }
        w->cv.notify_all();
    }
}

// Run job(rank) on every device concurrently; -> first failing status, group->last_ms = max over devices.
int run_all(eps_group* g, const std::function<int(uint32_t)>& job) {
    for (uint32_t r = 0; r < g->w.size(); r++) {
        eps_group::Worker* w = g->w[r];
        std::lock_guard<std::mutex> lk(w->m);
        w->job     = [job, r] { return job(r); };
        w->has_job = true;
        w->done    = false;
        w->cv.notify_all();
    }
    int   rc = EPS_OK;
    float ms = 0.f;
    for (uint32_t r = 0; r < g->w.size(); r++) {
        eps_group::Worker* w = g->w[r];
        std::unique_lock<std::mutex> lk(w->m);
        w->cv.wait(lk, [&] { return w->done; });
        if (w->rc != EPS_OK && rc == EPS_OK) {
            rc     = w->rc;
            g->err = "device " + std::to_string(w->ctx->dev) + ": " + w->ctx->err;
            g_last_error = g->err;
        }
        ms = std::max(ms, w->ms);
    }
    g->last_ms = ms;
    return rc;
}

void shard_block(uint32_t n, uint32_t world, uint32_t rank, uint32_t& start, uint32_t& count) {
    const uint32_t base = n / world, rem = n % world;  // contiguous blocks, the remainder to the first ranks
    start = rank * base + std::min(rank, rem);
    count = base + (rank < rem ? 1u : 0u);
}

}  // namespace

extern "C" {

int eps_group_create(const int* devices, uint32_t n_devices, eps_group** out) {
    if (!out) return gfail(nullptr, EPS_ERR_INVALID, "out is null");
    *out = nullptr;
    if (!devices || n_devices == 0 || n_devices > 64) return gfail(nullptr, EPS_ERR_INVALID, "need 1..64 devices");
    eps_group* g = new (std::nothrow) eps_group();
    if (!g) return gfail(nullptr, EPS_ERR_NOMEM, "out of host memory");
    for (uint32_t r = 0; r < n_devices; r++) {
        eps_ctx* ctx = nullptr;
        int      rc  = eps_ctx_create(devices[r], &ctx);
        if (rc != EPS_OK) {
            const std::string m = g_last_error;
            eps_group_destroy(g);
            return gfail(nullptr, rc, m);
        }
        auto* w = new eps_group::Worker();
        w->ctx  = ctx;
        cudaSetDevice(ctx->dev);
        cudaEventCreateWithFlags(&w->posted, cudaEventDisableTiming);
        g->w.push_back(w);
    }
    // peer access towards the gather device (errors are not fatal: cudaMemcpyPeerAsync then stages through the host)
    for (uint32_t r = 1; r < n_devices; r++) {
        if (devices[r] == devices[0]) continue;
        int can = 0;
        if (cudaDeviceCanAccessPeer(&can, devices[r], devices[0]) == cudaSuccess && can) {
            cudaSetDevice(devices[r]);
            cudaDeviceEnablePeerAccess(devices[0], 0);
        }
        cudaGetLastError();
    }
    for (auto* w : g->w) w->th = std::thread(worker_loop, w);
    *out = g;
    return EPS_OK;
}

int eps_group_destroy(eps_group* g) {
    if (!g) return EPS_OK;
    for (auto* w : g->w) {
        if (w->th.joinable()) {
            {
                std::lock_guard<std::mutex> lk(w->m);
                w->quit = true;
            }
            w->cv.notify_all();
            w->th.join();
        }
        if (w->ctx) {
            cudaSetDevice(w->ctx->dev);
            if (w->posted) cudaEventDestroy(w->posted);
            eps_ctx_destroy(w->ctx);
        }
        delete w;
    }
    if (!g->w.empty() || g->d_gather) {
        if (g->d_gather) cudaFree(g->d_gather);
        if (g->h_gather) cudaFreeHost(g->h_gather);
    }
    delete g;
    return EPS_OK;
}

uint32_t    eps_group_size(const eps_group* g) { return g ? static_cast<uint32_t>(g->w.size()) : 0u; }
eps_ctx*    eps_group_ctx(eps_group* g, uint32_t rank) { return (g && rank < g->w.size()) ? g->w[rank]->ctx : nullptr; }
const char* eps_group_last_error(const eps_group* g) { return g ? g->err.c_str() : g_last_error.c_str(); }

int eps_group_last_ms(eps_group* g, float* ms) {
    if (!g || !ms) return gfail(g, EPS_ERR_INVALID, "null argument");
    *ms = g->last_ms;
    return EPS_OK;
}

int eps_group_set_option(eps_group* g, int option, int64_t value) {
    if (!g) return gfail(nullptr, EPS_ERR_INVALID, "null group");
    for (auto* w : g->w)
        if (int rc = eps_set_option(w->ctx, option, value)) return gfail(g, rc, w->ctx->err);
    return EPS_OK;
}

int eps_group_set_potentials(eps_group* g, const double* V, uint32_t n_curves, uint32_t n_points, const double* scale,
                             int shard) {
    if (!g || !V || !scale) return gfail(g, EPS_ERR_INVALID, "null argument");
    if (shard != EPS_SHARD_CURVES && shard != EPS_SHARD_ENERGY) return gfail(g, EPS_ERR_INVALID, "shard: EPS_SHARD_CURVES or EPS_SHARD_ENERGY");
    const uint32_t G = static_cast<uint32_t>(g->w.size());
    if (shard == EPS_SHARD_CURVES && n_curves < G) return gfail(g, EPS_ERR_INVALID, "curve sharding needs at least one curve per device");
    g->shard = shard;
    g->nC    = 0;
    g->c0.assign(G, 0);
    g->cn.assign(G, n_curves);
    if (shard == EPS_SHARD_CURVES)
        for (uint32_t r = 0; r < G; r++) shard_block(n_curves, G, r, g->c0[r], g->cn[r]);
    const int rc = run_all(g, [&](uint32_t r) {
        return eps_set_potentials(g->w[r]->ctx, V + static_cast<size_t>(g->c0[r]) * n_points, g->cn[r], n_points, scale + g->c0[r]);
    });
    if (rc) return rc;
    g->nC = n_curves;
    g->N  = n_points;
    return EPS_OK;
}

int eps_group_sweep_uniform(eps_group* g, const double* E_lo, const double* E_hi, uint64_t n_energies, uint32_t* nodes) {
    if (!g || !E_lo || !E_hi) return gfail(g, EPS_ERR_INVALID, "null argument");
    if (g->nC == 0) return gfail(g, EPS_ERR_STATE, "eps_group_set_potentials has not been called");
    if (n_energies < 2 || n_energies >= (1ull << 32)) return gfail(g, EPS_ERR_INVALID, "bad energies");
    const uint32_t G = static_cast<uint32_t>(g->w.size()), nC = g->nC, nE = static_cast<uint32_t>(n_energies);
    if (g->shard == EPS_SHARD_CURVES)
        return run_all(g, [&](uint32_t r) {
            return eps_sweep_uniform(g->w[r]->ctx, E_lo + g->c0[r], E_hi + g->c0[r], nE,
                                     nodes ? nodes + static_cast<size_t>(g->c0[r]) * nE : nullptr, nullptr, nullptr);
        });
    // energy range: every device sweeps a contiguous slice of the ONE global grid E_j = E_lo + j dE,
    // expressed as (E0, dE, j0) so that the global energies are reproduced bit for bit
    std::vector<double> dE(nC);
    for (uint32_t c = 0; c < nC; c++) dE[c] = (E_hi[c] - E_lo[c]) / static_cast<double>(nE - 1);
    std::vector<std::vector<uint32_t>> tmp(G);
    const int rc = run_all(g, [&](uint32_t r) {
        uint32_t j0, n;
        shard_block(nE, G, r, j0, n);
        if (n == 0) return static_cast<int>(EPS_OK);
        uint32_t* dst = nullptr;
        if (nodes) {
            if (nC == 1) dst = nodes + j0;  // one curve: the slice is contiguous in the caller's array
            else {
                tmp[r].resize(static_cast<size_t>(nC) * n);
                dst = tmp[r].data();
            }
        }
        return eps_sweep_grid(g->w[r]->ctx, E_lo, dE.data(), j0, n, dst, nullptr, nullptr);
    });
    if (rc) return rc;
    if (nodes && nC > 1)
        for (uint32_t r = 0; r < G; r++) {
            uint32_t j0, n;
            shard_block(nE, G, r, j0, n);
            for (uint32_t c = 0; c < nC && n; c++)
                std::memcpy(nodes + static_cast<size_t>(c) * nE + j0, tmp[r].data() + static_cast<size_t>(c) * n, n * sizeof(uint32_t));
        }
    return EPS_OK;
}

int eps_group_solve_levels(eps_group* g, const eps_solve_params* p, const double* E_lo, const double* E_hi, double* levels,
                           double* widths, uint32_t* n_below) {
    if (!g || !p || !E_lo || !E_hi || !levels) return gfail(g, EPS_ERR_INVALID, "null argument");
    if (g->nC == 0) return gfail(g, EPS_ERR_STATE, "eps_group_set_potentials has not been called");
    if (p->v_max < p->v_min || p->n_coarse < 2) return gfail(g, EPS_ERR_INVALID, "v_max >= v_min and n_coarse >= 2 required");
    const uint32_t G = static_cast<uint32_t>(g->w.size()), nC = g->nC, nlev = p->v_max - p->v_min + 1;
    const bool     by_energy = g->shard == EPS_SHARD_ENERGY;
    // gather buffer on device 0: per device [levels | widths] of its (curve, level) pairs
    size_t slot = 0;
    for (uint32_t r = 0; r < G; r++) slot = std::max(slot, 2 * static_cast<size_t>(g->cn[r]) * nlev);
    cudaSetDevice(g->w[0]->ctx->dev);
    if (slot > g->gather_cap) {
        if (g->d_gather) cudaFree(g->d_gather);
        if (g->h_gather) cudaFreeHost(g->h_gather);
        g->d_gather = nullptr;
        g->h_gather = nullptr;
        g->gather_cap = 0;
        if (cudaMalloc(reinterpret_cast<void**>(&g->d_gather), slot * G * sizeof(double)) != cudaSuccess ||
            cudaMallocHost(reinterpret_cast<void**>(&g->h_gather), slot * G * sizeof(double)) != cudaSuccess)
            return gfail(g, EPS_ERR_NOMEM, "gather buffer");
        g->gather_cap = slot;
    }
    std::vector<double>   dE(nC);
    std::vector<uint32_t> nl(static_cast<size_t>(G) * nC, 0u);
    if (by_energy)
        for (uint32_t c = 0; c < nC; c++) dE[c] = (E_hi[c] - E_lo[c]) / static_cast<double>(p->n_coarse - 1);
    std::vector<uint8_t> active(G, 1);
    const int dev0 = g->w[0]->ctx->dev;
    int rc = run_all(g, [&](uint32_t r) {
        eps_ctx* ctx = g->w[r]->ctx;
        int      st;
        if (by_energy) {
            // the n_coarse - 1 grid intervals are split contiguously; a device owning intervals [a, b)
            // sweeps the points a..b, i.e. shares point b with its right neighbour: every bracket of
            // the global grid lies in exactly one slice
            uint32_t a, cnt;
            shard_block(p->n_coarse - 1, G, r, a, cnt);
            if (cnt == 0) {
                active[r] = 0;
                return static_cast<int>(EPS_OK);
            }
            eps_solve_params q = *p;
            q.n_coarse         = cnt + 1;
            st = eps_solve_levels_grid(ctx, &q, E_lo, dE.data(), a, nullptr, nullptr, nl.data() + static_cast<size_t>(r) * nC, nullptr);
        } else {
            st = eps_solve_levels(ctx, p, E_lo + g->c0[r], E_hi + g->c0[r], nullptr, nullptr, n_below ? n_below + g->c0[r] : nullptr);
        }
        if (st) return st;
        const size_t tot = static_cast<size_t>(g->cn[r]) * nlev;
        double*      dst = g->d_gather + static_cast<size_t>(r) * g->gather_cap;
        EPS_CUDA(ctx, cudaMemcpyPeerAsync(dst, dev0, ctx->d_levels.p, ctx->dev, tot * sizeof(double), ctx->stream));
        EPS_CUDA(ctx, cudaMemcpyPeerAsync(dst + tot, dev0, ctx->d_widths.p, ctx->dev, tot * sizeof(double), ctx->stream));
        EPS_CUDA(ctx, cudaEventRecord(g->w[r]->posted, ctx->stream));
        return static_cast<int>(EPS_OK);
    });
    if (rc) return rc;
    // device 0: wait for every peer's copy, then one device->host transfer
    eps_ctx* c0 = g->w[0]->ctx;
    cudaSetDevice(c0->dev);
    for (uint32_t r = 0; r < G; r++)
        if (active[r]) EPS_CUDA(c0, cudaStreamWaitEvent(c0->stream, g->w[r]->posted, 0));
    EPS_CUDA(c0, cudaMemcpyAsync(g->h_gather, g->d_gather, g->gather_cap * G * sizeof(double), cudaMemcpyDeviceToHost, c0->stream));
    EPS_CUDA(c0, cudaStreamSynchronize(c0->stream));
    c0->stats.d2h_bytes += g->gather_cap * G * sizeof(double);
    const double nan = std::nan("");
    if (!by_energy) {
        for (uint32_t r = 0; r < G; r++) {
            const size_t tot = static_cast<size_t>(g->cn[r]) * nlev;
            const double* src = g->h_gather + static_cast<size_t>(r) * g->gather_cap;
            std::memcpy(levels + static_cast<size_t>(g->c0[r]) * nlev, src, tot * sizeof(double));
            if (widths) std::memcpy(widths + static_cast<size_t>(g->c0[r]) * nlev, src + tot, tot * sizeof(double));
        }
        return EPS_OK;
    }
    const size_t tot = static_cast<size_t>(nC) * nlev;
    for (size_t i = 0; i < tot; i++) {
        levels[i] = nan;
        if (widths) widths[i] = nan;
    }
    uint32_t last = 0;
    for (uint32_t r = 0; r < G; r++) {
        if (!active[r]) continue;
        last = r;
        const double* src = g->h_gather + static_cast<size_t>(r) * g->gather_cap;
        for (size_t i = 0; i < tot; i++) {
            if (src[i] != src[i]) continue;
            if (levels[i] == levels[i]) return gfail(g, EPS_ERR_STATE, "a level was located by two devices: energy slices overlap");
            levels[i] = src[i];
            if (widths) widths[i] = src[tot + i];
        }
    }
    if (n_below) std::memcpy(n_below, nl.data() + static_cast<size_t>(last) * nC, nC * sizeof(uint32_t));
    return EPS_OK;
}

// =====================================================================================================
// Cross-process mailbox (one process per GPU, e.g. under torchrun): rank 0 owns a device buffer of
// `world` slots and exports its CUDA IPC handle (64 bytes, moved by whatever rendezvous channel the
// launcher has); every rank copies its small result into its slot device-to-device (NVLink peer
// write through the IPC mapping), followed in stream order by a sequence number; rank 0 polls the
// sequence numbers and fetches all slots with one device->host copy.  Replaces an all-gather + two
// staging copies + a stream sync through a framework by one peer write.
// =====================================================================================================
}  // extern "C"

struct eps_mailbox {
    eps_ctx*  ctx = nullptr;
    uint32_t  world = 0, rank = 0;
    size_t    slot_bytes = 0;   // per rank, multiple of 256
    bool      owner = false;
    char*     d_base = nullptr;  // [world][slot_bytes] data, then [world] uint32 sequence numbers (in rank 0's memory)
    uint32_t* h_seq = nullptr;   // pinned ring of outgoing sequence numbers
    uint32_t  ring = 0;
    uint32_t* h_flags = nullptr; // pinned, owner: polled copy of the sequence numbers
    char*     h_stage = nullptr; // pinned staging for host payloads
};

extern "C" {

int eps_mailbox_create(eps_ctx* ctx, uint32_t world, size_t bytes_per_rank, eps_mailbox** out, unsigned char* handle64) {
    if (int rc = bind(ctx)) return rc;
    EPS_REQUIRE(ctx, out && handle64 && world >= 1 && bytes_per_rank >= 1, EPS_ERR_INVALID, "bad mailbox arguments");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handle size");
    auto* mb       = new (std::nothrow) eps_mailbox();
    EPS_REQUIRE(ctx, mb, EPS_ERR_NOMEM, "out of host memory");
    mb->ctx        = ctx;
    mb->world      = world;
    mb->rank       = 0;
    mb->owner      = true;
    mb->slot_bytes = (bytes_per_rank + 255) / 256 * 256;
    const size_t total = mb->slot_bytes * world + 256 * (((world + 1) * sizeof(uint32_t) + 255) / 256);  // slots, flags, ack
    cudaError_t  e     = cudaMalloc(reinterpret_cast<void**>(&mb->d_base), total);
    if (e == cudaSuccess) e = cudaMemset(mb->d_base, 0, total);
    cudaIpcMemHandle_t h;
    if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, mb->d_base);
    if (e == cudaSuccess) e = cudaMallocHost(reinterpret_cast<void**>(&mb->h_seq), 64 * sizeof(uint32_t));
    if (e == cudaSuccess) e = cudaMallocHost(reinterpret_cast<void**>(&mb->h_flags), world * sizeof(uint32_t));
    if (e == cudaSuccess) e = cudaMallocHost(reinterpret_cast<void**>(&mb->h_stage), mb->slot_bytes);
    if (e != cudaSuccess) {
        eps_mailbox_destroy(mb);
        return fail(ctx, EPS_ERR_CUDA, std::string("eps_mailbox_create: ") + cudaGetErrorString(e));
    }
    std::memcpy(handle64, &h, 64);
    *out = mb;
    return EPS_OK;
}

int eps_mailbox_open(eps_ctx* ctx, const unsigned char* handle64, uint32_t world, uint32_t rank, size_t bytes_per_rank,
                     eps_mailbox** out) {
    if (int rc = bind(ctx)) return rc;
    EPS_REQUIRE(ctx, out && handle64 && rank >= 1 && rank < world && bytes_per_rank >= 1, EPS_ERR_INVALID, "bad mailbox arguments");
    auto* mb       = new (std::nothrow) eps_mailbox();
    EPS_REQUIRE(ctx, mb, EPS_ERR_NOMEM, "out of host memory");
    mb->ctx        = ctx;
    mb->world      = world;
    mb->rank       = rank;
    mb->slot_bytes = (bytes_per_rank + 255) / 256 * 256;
    cudaIpcMemHandle_t h;
    std::memcpy(&h, handle64, 64);
    cudaError_t e = cudaIpcOpenMemHandle(reinterpret_cast<void**>(&mb->d_base), h, cudaIpcMemLazyEnablePeerAccess);
    if (e == cudaSuccess) e = cudaMallocHost(reinterpret_cast<void**>(&mb->h_seq), 64 * sizeof(uint32_t));
    if (e == cudaSuccess) e = cudaMallocHost(reinterpret_cast<void**>(&mb->h_stage), mb->slot_bytes);
    if (e != cudaSuccess) {
        const std::string m = std::string("eps_mailbox_open: ") + cudaGetErrorString(e);
        cudaGetLastError();
        eps_mailbox_destroy(mb);
        return fail(ctx, EPS_ERR_CUDA, m);
    }
    *out = mb;
    return EPS_OK;
}

int eps_mailbox_destroy(eps_mailbox* mb) {
    if (!mb) return EPS_OK;
    if (mb->ctx && cudaSetDevice(mb->ctx->dev) == cudaSuccess) {
        if (mb->ctx->stream) cudaStreamSynchronize(mb->ctx->stream);
        if (mb->d_base) {
            if (mb->owner) cudaFree(mb->d_base);
            else cudaIpcCloseMemHandle(mb->d_base);
        }
        if (mb->h_seq) cudaFreeHost(mb->h_seq);
        if (mb->h_flags) cudaFreeHost(mb->h_flags);
        if (mb->h_stage) cudaFreeHost(mb->h_stage);
    }
    delete mb;
    return EPS_OK;
}

namespace {
// Flow control: a rank may not overwrite its slot before rank 0 has fetched the previous payload.
// Rank 0 acknowledges every collect by writing the sequence number behind the flags; a sender polls
// that word (one small peer read, normally satisfied at once) before posting seq > ack + 1.
int mailbox_wait_ack(eps_mailbox* mb, uint32_t seq) {
    if (mb->owner || seq <= 1) return EPS_OK;
    eps_ctx*        ctx = mb->ctx;
    const uint32_t* ack = reinterpret_cast<const uint32_t*>(mb->d_base + mb->slot_bytes * mb->world) + mb->world;
    const auto      t0  = std::chrono::steady_clock::now();
    for (;;) {
        EPS_CUDA(ctx, cudaMemcpyAsync(mb->h_seq + 63, ack, sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
        EPS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        if (static_cast<int32_t>(mb->h_seq[63] - (seq - 1)) >= 0) return EPS_OK;
        const double waited = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        if (waited > 120.0) return fail(ctx, EPS_ERR_STATE, "eps_mailbox_post: rank 0 never collected the previous payload");
        // rank 0 is busy elsewhere: do not hammer its memory with peer reads (each poll is a copy + a sync)
        if (waited > 200e-6) std::this_thread::sleep_for(std::chrono::microseconds(waited > 5e-3 ? 200 : 20));
    }
}

int mailbox_flag(eps_mailbox* mb, uint32_t seq) {  // sequence number after the payload, in stream order
    eps_ctx* ctx = mb->ctx;
    uint32_t* src = mb->h_seq + (mb->ring++ % 62);  // (slots 62 / 63: acknowledgement out / in)
    *src          = seq;
    uint32_t* flags = reinterpret_cast<uint32_t*>(mb->d_base + mb->slot_bytes * mb->world);
    EPS_CUDA(ctx, cudaMemcpyAsync(flags + mb->rank, src, sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream));
    return EPS_OK;
}
}  // namespace

/* This rank's levels and widths of its last eps_solve_levels* (device resident), copied
 * device-to-device into its slot: [levels (n) | widths (n)], n = curves x levels of the last solve. */
int eps_mailbox_post_levels(eps_mailbox* mb, uint32_t seq) {
    if (!mb) return fail(nullptr, EPS_ERR_INVALID, "null mailbox");
    eps_ctx* ctx = mb->ctx;
    if (int rc = bind(ctx)) return rc;
    const size_t n = ctx->last_total;
    EPS_REQUIRE(ctx, n > 0, EPS_ERR_STATE, "no level search has run on this context");
    EPS_REQUIRE(ctx, 2 * n * sizeof(double) <= mb->slot_bytes, EPS_ERR_INVALID, "mailbox slot too small for the levels");
    if (int rc = mailbox_wait_ack(mb, seq)) return rc;
    char* dst = mb->d_base + mb->slot_bytes * mb->rank;
    EPS_CUDA(ctx, cudaMemcpyAsync(dst, ctx->d_levels.p, n * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
    EPS_CUDA(ctx, cudaMemcpyAsync(dst + n * sizeof(double), ctx->d_widths.p, n * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
    return mailbox_flag(mb, seq);
}

/* A small host payload (e.g. a checksum) into this rank's slot. */
int eps_mailbox_post(eps_mailbox* mb, const void* src, size_t bytes, uint32_t seq) {
    if (!mb) return fail(nullptr, EPS_ERR_INVALID, "null mailbox");
    eps_ctx* ctx = mb->ctx;
    if (int rc = bind(ctx)) return rc;
    EPS_REQUIRE(ctx, src && bytes <= mb->slot_bytes, EPS_ERR_INVALID, "payload larger than the mailbox slot");
    if (int rc = mailbox_wait_ack(mb, seq)) return rc;
    EPS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // the staging buffer may still be in flight
    std::memcpy(mb->h_stage, src, bytes);
    EPS_CUDA(ctx, cudaMemcpyAsync(mb->d_base + mb->slot_bytes * mb->rank, mb->h_stage, bytes, cudaMemcpyHostToDevice, ctx->stream));
    return mailbox_flag(mb, seq);
}

/* Rank 0: wait until every rank's slot carries `seq`, then fetch the first bytes_per_rank bytes of
 * every slot (out: [world][bytes_per_rank]) with one strided device->host copy, and acknowledge.
 * EPS_ERR_STATE after timeout_s. */
int eps_mailbox_collect(eps_mailbox* mb, uint32_t seq, void* out, size_t bytes_per_rank, double timeout_s) {
    if (!mb) return fail(nullptr, EPS_ERR_INVALID, "null mailbox");
    eps_ctx* ctx = mb->ctx;
    if (int rc = bind(ctx)) return rc;
    EPS_REQUIRE(ctx, mb->owner && out, EPS_ERR_INVALID, "collect is for the mailbox's owner (rank 0)");
    EPS_REQUIRE(ctx, bytes_per_rank >= 1 && bytes_per_rank <= mb->slot_bytes, EPS_ERR_INVALID, "bytes_per_rank exceeds the slot");
    const uint32_t* flags = reinterpret_cast<const uint32_t*>(mb->d_base + mb->slot_bytes * mb->world);
    const auto      t0    = std::chrono::steady_clock::now();
    for (;;) {
        EPS_CUDA(ctx, cudaMemcpyAsync(mb->h_flags, flags, mb->world * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
        EPS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        bool all = true;
        for (uint32_t r = 0; r < mb->world; r++) all = all && mb->h_flags[r] == seq;
        if (all) break;
        const double waited = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        if (waited > timeout_s) return fail(ctx, EPS_ERR_STATE, "eps_mailbox_collect: timed out waiting for a rank");
        if (waited > 5e-3) std::this_thread::sleep_for(std::chrono::microseconds(100));  // a rank is far behind: poll gently
    }
    // the first bytes_per_rank bytes of every slot, packed [world][bytes_per_rank], in one strided copy
    EPS_CUDA(ctx, cudaMemcpy2DAsync(out, bytes_per_rank, mb->d_base, mb->slot_bytes, bytes_per_rank, mb->world,
                                    cudaMemcpyDeviceToHost, ctx->stream));
    mb->h_seq[62] = seq;  // acknowledge: the senders may post again
    EPS_CUDA(ctx, cudaMemcpyAsync(const_cast<uint32_t*>(flags) + mb->world, mb->h_seq + 62, sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream));
    EPS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->stats.d2h_bytes += bytes_per_rank * mb->world;
    return EPS_OK;
}

size_t eps_mailbox_slot_bytes(const eps_mailbox* mb) { return mb ? mb->slot_bytes : 0; }

uint32_t eps_cooley_segment_length(uint32_t n_steps, uint64_t n_items) { return cooley_segment_length(n_steps, n_items); }

}  // extern "C"
