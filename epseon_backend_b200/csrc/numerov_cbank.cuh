// numerov_cbank.cuh -- constant-bank variant of the many-energy Numerov sweep (sm_100a).
//
// Same recurrence, same operation order and therefore the same bits as
// numerov_sweep_kernel (numerov_kernels.cuh; spec DESIGN.md section 3.3; oracle sweep_block()).
// What changes is where the coefficient F_k comes from.  On B200 an FP64 instruction reads one
// 64-bit register operand per cycle (scripts/microbench.cu, microbench3.cu; profiles/): with F_k in a
// vector register (warp-broadcast LDS) the DADD reads two of them; with F_k in a UNIFORM register it
// reads one and the step's mix runs at 97-101 % of the pipe in isolation.  The only road into
// uniform registers is the constant bank (LDCU), so the table is fed through the kernel-parameter
// constant bank: the
// host cuts the curve into chunks of kCbChunk steps, every launch carries its chunk BY VALUE
// (__grid_constant__, 31 KiB of the 32 764-byte parameter space) and the per-energy state
// (X, S, exponent, node count, last sign) is carried from launch to launch in HBM
// (28 B per energy per 3968 steps: noise).  Only valid when every row of the launch is on the
// same curve (C2 / C5 shape); multi-curve batches keep the TMA/shared-memory kernel.
// (Tried and dropped: a divergent vector LDC per 256 steps to prefetch the next parameter lines
// into the SM's constant cache -- it serialises the constant pipe and cost 10 % instead;
// profiles/r1_cbank.log.)
#pragma once
#include "numerov_kernels.cuh"

namespace eps {

constexpr int kCbChunk = 3968;  // steps per launch: 31 renormalisation blocks, 31 744 B of parameters

struct alignas(16) FChunk {
    double2 f2[kCbChunk / 2 + 1];  // + one pad pair: the march loads one pair ahead
};

struct CbState {  // per-energy carry between chunk launches, [row * out_stride + j]
    double*   X;
    double*   S;
    int32_t*  expo;
    uint32_t* nodes;
    uint32_t* prev;
};

// grid = n_jobs * chunks_per_job CTAs of kThreads threads, kEpt energies per thread -- or a GROUP of
// them (cta_base .. cta_base + gridDim.x): the host may run all chunk launches of one group of
// energies before the next group's, so that the group's carried state (28 B per energy) stays in L2
// between launches instead of making a round trip through HBM per chunk (launch_cbank_variant).
//   len    steps of this chunk (multiple of 128 except for the curve's last chunk)
//   first  != 0: initialise the state instead of loading it;  last != 0: emit results.
template <int kEpt, int kThreads, int kStride, bool kTails, int kForm>
__global__ void __launch_bounds__(kThreads)
numerov_cbank_kernel(const __grid_constant__ FChunk P, const Job* __restrict__ jobs,
                     const uint32_t chunks_per_job, const double* __restrict__ Eexp,
                     const uint64_t out_stride, const double scale, const uint32_t len, const int first,
                     const int last, const int pdl_late, const uint32_t cta_base, const CbState st,
                     uint32_t* __restrict__ nodes_out,
                     double* __restrict__ mant_out, int32_t* __restrict__ exp_out,
                     unsigned long long* __restrict__ steps_done, const int* __restrict__ stop_flag) {
    static_assert(kStride == 1 || kStride == 8 || kStride == 32, "sign sampling stride");
    constexpr uint32_t kPerCta = kThreads * kEpt;
    // Programmatic dependent launch: the chunk launches of one sweep are chained with
    // cudaLaunchAttributeProgrammaticStreamSerialization.  griddepcontrol.wait returns once the
    // previous chunk has completed and its state stores are visible; only THEN does this chunk let
    // its own dependent become resident, so exactly one chunk is ever parked behind the running one
    // (triggering before the wait lets every queued chunk of the sweep pile up on the SMs:
    // measured 2.21 ms instead of 1.97 ms on a one-wave sweep).  pdl_late > 0 moves the trigger to
    // pdl_late renormalisation blocks (128 steps, ~2 us) before the end of the chunk: the dependent
    // is then resident only while this chunk drains, which is all a one-wave sweep can overlap.
    asm volatile("griddepcontrol.wait;" ::: "memory");
    if (pdl_late == 0) asm volatile("griddepcontrol.launch_dependents;");
    if (*stop_flag != 0) return;  // eps_request_stop: the remaining chunk launches of the sweep drain at once
    const uint32_t cta     = cta_base + blockIdx.x;  // the launch covers CTAs [cta_base, cta_base + gridDim.x) of the sweep
    const uint32_t job_idx = cta / chunks_per_job;
    const uint32_t chunk   = cta - job_idx * chunks_per_job;
    const Job      job     = jobs[job_idx];
    const uint32_t e_base  = chunk * kPerCta;
    if (e_base >= job.nE) return;
    if (threadIdx.x == 0)
        atomicAdd(steps_done, static_cast<unsigned long long>(len) * min(job.nE - e_base, kPerCta));

    Chain    c[kEpt];
    double   ep[kEpt];
    int      expo[kEpt];
    uint32_t n_nodes[kEpt], prev[kEpt];
#pragma unroll
    for (int i = 0; i < kEpt; i++) {
        uint32_t j = e_base + i * kThreads + threadIdx.x;
        if (j >= job.nE) j = job.nE - 1;  // keep warps converged; result discarded
        double E;
        if (Eexp != nullptr) E = Eexp[job.e_off + j];
        else E = __dadd_rn(job.E0, __dmul_rn(__ull2double_rn(static_cast<unsigned long long>(job.j0) + j), job.dE));
        ep[i] = energy_const<kForm>(scale, E);
        if (first) {
            c[i]       = kForm == 0 ? Chain{1.0, 0.0} : Chain{1.0, 1.0};
            expo[i]    = 0;
            n_nodes[i] = 0;
            prev[i]    = 0;
        } else {
            const uint64_t o = static_cast<uint64_t>(job_idx) * out_stride + j;
            c[i]       = Chain{st.X[o], st.S[o]};
            expo[i]    = st.expo[o];
            n_nodes[i] = st.nodes[o];
            prev[i]    = st.prev[o];
        }
    }

    const uint32_t n_full = len / kRenorm;
    uint32_t       k      = 0;
    double2        nxt    = P.f2[0];  // software pipeline: the pair of steps (k, k+1) is loaded one pair ahead,
                                      // across the loop back-edges ptxas will not hoist an LDCU over
#pragma unroll 1
    for (uint32_t r = 0; r < n_full; r++) {
        if (pdl_late != 0 && r + pdl_late == n_full) asm volatile("griddepcontrol.launch_dependents;");
#pragma unroll 1
        for (int q = 0; q < kRenorm / 32; q++) {
            uint32_t mask[kEpt];
#pragma unroll
            for (int i = 0; i < kEpt; i++) mask[i] = 0;
#pragma unroll
            for (int p = 0; p < 16; p++) {
                const double2 ff = nxt;
                nxt              = P.f2[(k >> 1) + p + 1];  // uniform index: LDCU into uniform registers
#pragma unroll
                for (int i = 0; i < kEpt; i++) {
                    step_form<kForm>(c[i], ff.x, ep[i]);
                    if (kStride == 1) mask[i] = __funnelshift_l(static_cast<uint32_t>(__double2hiint(c[i].X)), mask[i], 1);
                }
#pragma unroll
                for (int i = 0; i < kEpt; i++) {
                    step_form<kForm>(c[i], ff.y, ep[i]);
                    if (kStride == 1) mask[i] = __funnelshift_l(static_cast<uint32_t>(__double2hiint(c[i].X)), mask[i], 1);
                }
                if (kStride == 8 && (p & 3) == 3) {
#pragma unroll
                    for (int i = 0; i < kEpt; i++) {
                        const uint32_t cur = static_cast<uint32_t>(__double2hiint(c[i].X));
                        n_nodes[i] += (cur ^ prev[i]) >> 31;
                        prev[i] = cur;
                    }
                }
            }
#pragma unroll
            for (int i = 0; i < kEpt; i++) {
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
    // ragged tail of the curve's last chunk (< 128 steps): plain per-step counting
    for (; k < len; k++) {
        const double2 ff = P.f2[k >> 1];
        const double  Fk = (k & 1) ? ff.y : ff.x;
#pragma unroll
        for (int i = 0; i < kEpt; i++) {
            const uint32_t before = static_cast<uint32_t>(__double2hiint(c[i].X));
            step_form<kForm>(c[i], Fk, ep[i]);
            const uint32_t after = static_cast<uint32_t>(__double2hiint(c[i].X));
            n_nodes[i] += (before ^ after) >> 31;
            prev[i] = (kStride == 1) ? (after >> 31) : after;
        }
    }

#pragma unroll
    for (int i = 0; i < kEpt; i++) {
        const uint32_t j = e_base + i * kThreads + threadIdx.x;
        if (j >= job.nE) continue;
        const uint64_t o = static_cast<uint64_t>(job_idx) * out_stride + j;
        if (last) {
            renorm(c[i], expo[i]);
            nodes_out[o] = n_nodes[i];
            if (kTails) {
                mant_out[o] = c[i].X;
                exp_out[o]  = expo[i];
            }
        } else {
            st.X[o]     = c[i].X;
            st.S[o]     = c[i].S;
            st.expo[o]  = expo[i];
            st.nodes[o] = n_nodes[i];
            st.prev[o]  = prev[i];
        }
    }
}

}  // namespace eps
