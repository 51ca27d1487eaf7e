// VibwaAlgorithm<FP>::run -- the hot path behind the reference's slot (vibwa.hpp:605-637).
// C++ host code -> thin C ABI (include/epseon_cuda.h) -> sm_100a kernels.  No Vulkan, no PyTorch,
// no CPU fallback: if the CUDA library cannot run, the task fails and says so in its status.
#pragma once
#include "epseon/gpu/algorithms/vibwa.hpp"
#include "epseon/gpu/task_handle.hpp"

#include "epseon_cuda.h"

#include <algorithm>
#include <cmath>
#include <limits>
#include <map>
#include <mutex>
#include <stdexcept>
#include <string>
#include <vector>

namespace epseon::gpu::cpp {

    namespace detail {
        constexpr double kHbar2Over2 = 16.857629206; // amu * Angstrom^2 * cm^-1

        // Contexts are pooled per device: creating one (stream, events, a dozen device buffers) and
        // tearing it down costs ~10 ms per task against ~2 ms of solve for a small task.  A task
        // takes an idle context of its device or creates one, and hands it back when it is done;
        // at most kMaxIdle stay parked per device, so concurrent tasks still get a context each.
        // Idle contexts are deliberately not destroyed at process exit (the CUDA runtime may
        // already be gone when static destructors run).
        class CtxPool {
            static constexpr size_t kMaxIdle = 2;
            std::mutex                                 mutex;
            std::map<int, std::vector<eps_ctx*>>       idle;

          public:
            static CtxPool& instance() {
                static CtxPool* pool = new CtxPool(); // never destroyed, see above
                return *pool;
            }
            eps_ctx* acquire(int device) {
                {
                    std::lock_guard<std::mutex> g(mutex);
                    auto&                       v = idle[device];
                    if (!v.empty()) {
                        eps_ctx* ctx = v.back();
                        v.pop_back();
                        return ctx;
                    }
                }
                eps_ctx* ctx = nullptr;
                if (eps_ctx_create(device, &ctx) != EPS_OK)
                    throw std::runtime_error(std::string("eps_ctx_create: ") + eps_last_error(nullptr));
                return ctx;
            }
            // Bytes a parked context may keep (its grow-only scratch is trimmed above that): a small
            // task re-reserves a few MB in microseconds, a wavefunction task's GBs must not stay pinned.
            static constexpr uint64_t kMaxParkedBytes = 64ull << 20;

            void clear() {
                std::map<int, std::vector<eps_ctx*>> all;
                {
                    std::lock_guard<std::mutex> g(mutex);
                    all.swap(idle);
                }
                for (auto& [dev, v] : all)
                    for (eps_ctx* ctx : v) eps_ctx_destroy(ctx);
            }
            void release(int device, eps_ctx* ctx, bool healthy) {
                if (ctx == nullptr) return;
                if (healthy) {
                    uint64_t bytes = 0;
                    if (eps_ctx_device_bytes(ctx, &bytes) != EPS_OK ||
                        (bytes > kMaxParkedBytes && eps_ctx_trim(ctx, /*drop_potentials=*/1) != EPS_OK))
                        healthy = false;
                }
                if (healthy) {
                    std::lock_guard<std::mutex> g(mutex);
                    auto&                       v = idle[device];
                    if (v.size() < kMaxIdle) {
                        v.push_back(ctx);
                        return;
                    }
                }
                eps_ctx_destroy(ctx);
            }
        };

        // SM count per device, asked once (cudaGetDeviceProperties costs milliseconds per call).
        inline int sm_count_of(int device) {
            static std::mutex         mutex;
            static std::map<int, int> cache;
            std::lock_guard<std::mutex> g(mutex);
            auto                        it = cache.find(device);
            if (it != cache.end()) return it->second;
            eps_device_props props{};
            if (eps_device_get_props(device, &props) != EPS_OK)
                throw std::runtime_error(std::string("eps_device_get_props: ") + eps_last_error(nullptr));
            cache[device] = props.sm_count;
            return props.sm_count;
        }

        struct CtxGuard {
            int      device  = 0;
            eps_ctx* ctx     = nullptr;
            bool     healthy = false; // set once the task ran to its end: only then is the context reused
            ~CtxGuard() { CtxPool::instance().release(device, ctx, healthy); }
        };

        // Releases every idle pooled context (additive; Python: release_device_memory()).
        inline void trim_context_pool() { CtxPool::instance().clear(); }

        inline void check(int rc, eps_ctx* ctx, const char* what) {
            if (rc != EPS_OK) throw std::runtime_error(std::string(what) + ": " + eps_last_error(ctx));
        }
    } // namespace detail

    template <typename FP>
    void VibwaAlgorithm<FP>::run(const std::stop_token& stop_token, TaskHandle<FP>* handle) {
        if (stop_token.stop_requested()) return handle->setCancelled();

        const TaskConfigurator<FP>& configurator = handle->getTaskConfigurator();
        const auto                  hardware     = configurator.getHardwareConfig();
        const auto                  source       = configurator.getPotentialSource();
        const auto algorithm = std::dynamic_pointer_cast<VibwaAlgorithmConfig<FP>>(configurator.getAlgorithmConfig());
        if (!algorithm) throw std::runtime_error("VibwaAlgorithm needs a VibwaAlgorithmConfig");
        if (algorithm->getMaxLevel() < algorithm->getMinLevel()) throw std::runtime_error("max_level < min_level");

        // ---- N1: tabulate V(r) on the host (double), one row per curve ----
        handle->setStatus("tabulating potentials");
        const auto                table = source->get_potential_data();
        const std::vector<double> steps = source->get_grid_steps();
        const uint32_t            nT    = static_cast<uint32_t>(table.size()); // tables (one per configured curve)
        if (nT == 0) {
            const std::string why = source->get_last_error();
            throw std::runtime_error("potential source holds no curves" + (why.empty() ? std::string() : ": " + why));
        }
        // additive: every table is solved once per rotational state J; row = table*nJ + j
        const std::vector<uint32_t>  J  = configurator.getRotationalStates(); // by value
        const uint32_t               nJ = static_cast<uint32_t>(J.size());
        const bool     rotating = std::any_of(J.begin(), J.end(), [](uint32_t j) { return j != 0; }) || nJ > 1;
        const uint32_t nC       = nT * nJ;
        const uint32_t N = static_cast<uint32_t>(table.front().size());
        if (hardware->getPotentialBufferSize() < N)
            throw std::runtime_error("potential_buffer_size is smaller than the potential's point count");
        std::vector<double> V(static_cast<size_t>(nT) * N), scale(nT);
        const double m0 = algorithm->getMassAtom0(), m1 = algorithm->getMassAtom1();
        const double mu = (m0 * m1) / (m0 + m1);
        const double c  = mu / detail::kHbar2Over2;
        for (uint32_t k = 0; k < nT; k++) {
            if (table[k].size() != N) throw std::runtime_error("all curves must have the same point count");
            for (uint32_t i = 0; i < N; i++) V[static_cast<size_t>(k) * N + i] = static_cast<double>(table[k][i]);
            scale[k] = ((steps[k] * steps[k]) * c) / 12.0;
        }
        if (stop_token.stop_requested()) return handle->setCancelled();

        // ---- device: resident coefficient tables ----
        handle->setStatus("uploading potentials");
        detail::CtxGuard guard;
        guard.device = handle->getDeviceInterface().getCudaOrdinal();
        guard.ctx    = detail::CtxPool::instance().acquire(guard.device);
        eps_ctx* ctx = guard.ctx;
        // cancel() -> stop_token -> eps_request_stop: the C ABI tests the flag between refinement
        // rounds and the sweep CTAs test it on entry, so a running solve stops within a round.
        detail::check(eps_reset_stop(ctx), ctx, "eps_reset_stop");
        const std::stop_callback on_stop(stop_token, [ctx] { eps_request_stop(ctx); });
        const auto run_step = [&](int rc, const char* what) { // -> true when the step was cancelled
            if (rc == EPS_ERR_CANCELLED) return true;
            detail::check(rc, ctx, what);
            return false;
        };
        // The drop-in path runs the ACCURATE recurrence (D form): its
        // eigenvalues sit within ~1e-14 of the discrete problem's at every grid size, which is what
        // makes the 1e-12 tolerance below meaningful (the 4-operation X form carries a rounding-noise
        // floor of 1e-9 .. 2e-8 on 2e5 .. 1e6-point grids); it costs 5 instead of 4 FP64 operations
        // per grid step.
        detail::check(eps_set_option(ctx, EPS_OPT_FORM, 1), ctx, "eps_set_option(EPS_OPT_FORM)");
        if (rotating) {
            const std::vector<double> origins = source->get_grid_origins();
            detail::check(eps_set_potentials_rot(ctx, V.data(), nT, N, scale.data(), origins.data(), steps.data(),
                                                 J.data(), nJ),
                          ctx, "eps_set_potentials_rot");
        } else {
            detail::check(eps_set_potentials(ctx, V.data(), nT, N, scale.data()), ctx, "eps_set_potentials");
        }

        // ---- search window per curve: [V_min, V_last - min_distance_to_asymptote] ----
        std::vector<double> E_lo(nC), E_hi(nC);
        const double        margin = algorithm->getMinDistanceToAsymptote();
        for (uint32_t k = 0; k < nC; k++) {
            eps_curve_info info{};
            detail::check(eps_get_curve_info(ctx, k, &info), ctx, "eps_get_curve_info");
            E_lo[k] = info.v_min;
            E_hi[k] = std::max(info.v_min, info.v_last - margin);
        }
        if (stop_token.stop_requested()) return handle->setCancelled();

        // ---- N2..N6: coarse sweep, bracketing, k-section refinement ----
        handle->setStatus("solving levels");
        eps_solve_params p{};
        p.v_min         = algorithm->getMinLevel();
        p.v_max         = algorithm->getMaxLevel();
        p.n_coarse      = std::clamp<uint32_t>(std::max<uint32_t>(hardware->getGroupSize(), 1024u), 1024u, 65536u);
        p.refine_points = 256;
        p.max_rounds    = 16;
        p.rel_tol       = std::is_same_v<FP, float> ? 1e-8 : 1e-12;
        // One curve, a large coarse grid: the sweep costs whole waves of 512-energy CTAs, so the grid is
        // rounded up to the next full wave (65 536 -> 148 x 512 = 75 776 on a B200): finer brackets for
        // the same sweep time (profiles/r1g_ncu.md: 128 and 148 CTAs take the same 1.87 ms).
        if (nC == 1 && p.n_coarse >= 32768u) {
            const uint32_t wave = static_cast<uint32_t>(detail::sm_count_of(guard.device)) * 512u;
            p.n_coarse          = (p.n_coarse + wave - 1) / wave * wave;
        }
        const uint32_t        nlev = p.v_max - p.v_min + 1;
        // A handful of curves: a refinement round is latency-bound, so many points per round (few
        // rounds) win.  Hundreds of curves: rounds are throughput-bound and k-section with fewer
        // points per round needs less total work -- take just enough points for one curve's rows to
        // fill a 256-energy CTA ("small packed CTAs").
        {
            const uint64_t sms  = static_cast<uint64_t>(detail::sm_count_of(guard.device));
            const uint64_t wave = sms * 512u;
            if (nC > 1 && static_cast<uint64_t>(nC) * nlev * p.refine_points >= 4u * wave) {
                // many curves: one curve's rows fill a 128-energy CTA (8 levels -> 16 points; k-section needs
                // M / log2(M + 1) sweeps-worth of energies per bit: C4 20.1 ms with 16 points, 25.5 with 32)
                uint32_t rows = 1;
                while (rows < nlev) rows <<= 1;
                p.refine_points = std::max<uint32_t>(4u, 128u / std::min<uint32_t>(rows, 128u));
                p.max_rounds    = 32;
            } else if (nC == 1) {
                // one curve: a round of <= 256 energies per SM (2 chains per thread) is latency-bound and costs
                // about half a full wave, so it is filled exactly: C2 3.97 ms with 17 x 2228 points, 5.64 with 4457
                p.refine_points = static_cast<uint32_t>(std::clamp<uint64_t>(sms * 256u / nlev, 256u, 32768u));
            }
        }
        if (configurator.getLevelSearch() != 0) { // Cooley iteration instead of k-section sweeps
            p.flags      = EPS_SOLVE_COOLEY | (configurator.getLevelSearch() == 2 ? EPS_SOLVE_OPEN_TAIL : 0);
            p.max_rounds = 40;
        }
        handle->setSearchParameters(p.n_coarse, p.refine_points, p.max_rounds, p.rel_tol);
        std::vector<double>   lev(static_cast<size_t>(nC) * nlev);
        std::vector<uint32_t> below(nC);
        detail::check(eps_timer_start(ctx), ctx, "eps_timer_start");
        int solve_rc;
        if (configurator.getEnergyShardWorld() > 1) {
            // energy-range shard: this task owns a contiguous block of the n_coarse - 1 intervals of the
            // global coarse grid (the remainder to the first ranks) and sweeps its end points, sharing
            // one point with its right neighbour -- every bracket lies in exactly one slice, and the
            // affine grid (E0, dE, j0) reproduces the global energies bit for bit.
            const uint32_t world = configurator.getEnergyShardWorld(), rank = configurator.getEnergyShardRank();
            const uint32_t n_int = p.n_coarse - 1, base = n_int / world, rem = n_int % world;
            const uint32_t a = rank * base + std::min(rank, rem), cnt = base + (rank < rem ? 1u : 0u);
            if (cnt == 0) { // more ranks than intervals: nothing to search here
                std::fill(lev.begin(), lev.end(), std::numeric_limits<double>::quiet_NaN());
                std::fill(below.begin(), below.end(), 0u);
                solve_rc = EPS_OK;
            } else {
                std::vector<double> dE(nC);
                for (uint32_t k = 0; k < nC; k++) dE[k] = (E_hi[k] - E_lo[k]) / static_cast<double>(p.n_coarse - 1);
                eps_solve_params q = p;
                q.n_coarse         = cnt + 1;
                solve_rc = eps_solve_levels_grid(ctx, &q, E_lo.data(), dE.data(), a, lev.data(), nullptr, below.data(), nullptr);
            }
        } else {
            solve_rc = eps_solve_levels(ctx, &p, E_lo.data(), E_hi.data(), lev.data(), nullptr, below.data());
        }
        if (run_step(solve_rc, "eps_solve_levels")) {
            eps_sync(ctx); // let the drained launches finish before the context is parked
            guard.healthy = eps_reset_stop(ctx) == EPS_OK;
            return handle->setCancelled();
        }
        float ms = 0.f;
        detail::check(eps_timer_stop(ctx, &ms), ctx, "eps_timer_stop");

        std::vector<std::vector<FP>> out(nC, std::vector<FP>(nlev));
        for (uint32_t k = 0; k < nC; k++)
            for (uint32_t l = 0; l < nlev; l++) out[k][l] = static_cast<FP>(lev[static_cast<size_t>(k) * nlev + l]);
        handle->setResults(std::move(out), std::move(below), ms);

        // ---- N7 (on request): normalised wavefunctions of the located levels ----
        if (configurator.getWavefunctionOutput() && !stop_token.stop_requested()) {
            handle->setStatus("computing wavefunctions");
            std::vector<double> psi(static_cast<size_t>(nC) * nlev * N), steps_eff(nC);
            for (uint32_t k = 0; k < nC; k++) steps_eff[k] = steps[k / nJ];
            detail::check(eps_wavefunctions(ctx, lev.data(), nlev, steps_eff.data(), psi.data(), nullptr), ctx,
                          "eps_wavefunctions");
            handle->setWavefunctions(std::move(psi), nC, nlev, N);
        }
        guard.healthy = true;
        handle->setStatus("done");
    }
} // namespace epseon::gpu::cpp
