// Algorithm<FP> -- the abstract hot-path slot.
// Reference: cpp/gpu/include/epseon/gpu/algorithms/algorithm.hpp:34 -- identical signature:
//     virtual void run(const std::stop_token&, TaskHandle<FP>*) = 0;
// run() executes on the TaskHandle's worker jthread (task_handle.hpp).
#pragma once
#include "epseon/gpu/predecl.hpp"

#include <memory>
#include <stop_token>
#include <type_traits>

namespace epseon::gpu::cpp {

    template <typename FP>
    class Algorithm : public std::enable_shared_from_this<Algorithm<FP>> {
        static_assert(std::is_floating_point_v<FP>, "FP must be an floating-point type.");

      public:
        Algorithm() = default;
        Algorithm(const Algorithm&) : std::enable_shared_from_this<Algorithm<FP>>() {}
        Algorithm& operator=(const Algorithm&) { return *this; }
        virtual ~Algorithm() = default;

        virtual void run(const std::stop_token&, TaskHandle<FP>*) = 0;
    };
} // namespace epseon::gpu::cpp
