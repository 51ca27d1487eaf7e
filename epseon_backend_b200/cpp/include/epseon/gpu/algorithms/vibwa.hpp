// VibwaAlgorithm<FP> -- vibrational levels of diatomic potential curves ("VIBWA").
// Reference: cpp/gpu/include/epseon/gpu/algorithms/vibwa.hpp.  There, run() (:605-637) builds a
// Vulkan logical device (:654-674), a VMA allocator and group_size x 7 buffers (:57-234, :236-603),
// one descriptor set, prints "Thread finished" and returns -- no compute.  Here run() is the real
// hot path: tabulate -> C ABI (include/epseon_cuda.h) -> CUDA sweep / bracketing / refinement ->
// results stored on the TaskHandle.  The definition lives in vibwa_run.hpp (it needs the complete
// TaskHandle type); like the reference (vibwa.hpp:8,11) this header pulls in algorithm_config.hpp
// and task_handle.hpp, so including it alone is enough to call run().
#pragma once
#include "epseon/gpu/predecl.hpp"

#include "epseon/gpu/algorithms/algorithm.hpp"

namespace epseon::gpu::cpp {

    template <typename FP>
    class VibwaAlgorithm : public Algorithm<FP> {
      public:
        VibwaAlgorithm()           = default;
        ~VibwaAlgorithm() override = default;

        void run(const std::stop_token& stop_token, TaskHandle<FP>* handle) override;
    };
} // namespace epseon::gpu::cpp

#include "epseon/gpu/task_configurator/algorithm_config.hpp"
#include "epseon/gpu/task_handle.hpp"
