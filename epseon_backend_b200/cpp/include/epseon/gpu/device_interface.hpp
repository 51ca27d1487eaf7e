// ComputeDeviceInterface -- handle to one GPU: hands out configurators and accepts tasks.
// Reference: cpp/gpu/include/epseon/gpu/device_interface.hpp:19-50 -- same methods and the same
// "not fully configured" std::runtime_error from submitTask (:40-43).  submitTask does NOT start the
// worker (C++ callers call startWorker(); the Python wrapper starts it, python/api.hpp:32-36).
#pragma once
#include "epseon/gpu/predecl.hpp"

#include "epseon/gpu/compute_context.hpp"
#include "epseon/gpu/task_configurator/task_configurator.hpp"

#include <memory>
#include <stdexcept>

namespace epseon::gpu::cpp {

    class ComputeDeviceInterface : public std::enable_shared_from_this<ComputeDeviceInterface> {
        std::shared_ptr<ComputeContextState> computeContextState;
        std::shared_ptr<PhysicalDevice>      physicalDevice;

      public:
        ComputeDeviceInterface(std::shared_ptr<ComputeContextState> state, std::shared_ptr<PhysicalDevice> device) :
            computeContextState(std::move(state)), physicalDevice(std::move(device)) {}

        template <typename FP>
        std::shared_ptr<TaskConfigurator<FP>> getTaskConfigurator() {
            return std::make_shared<TaskConfigurator<FP>>();
        }

        [[nodiscard]] const PhysicalDevice& getPhysicalDevice() const { return *physicalDevice; }
        [[nodiscard]] int                   getCudaOrdinal() const { return physicalDevice->ordinal; }

        template <typename FP>
        std::shared_ptr<epseon::gpu::cpp::TaskHandle<FP>> submitTask(std::shared_ptr<TaskConfigurator<FP>> task_config);

        [[nodiscard]] const ComputeContextState& getComputeContextState() const { return *computeContextState; }
    };
} // namespace epseon::gpu::cpp
