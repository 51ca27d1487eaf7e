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

// submitTask needs the complete TaskHandle; the reference's device_interface.hpp includes
// task_handle.hpp too (:9).  Placed after the class so that either header may come first.
#include "epseon/gpu/task_handle.hpp"

namespace epseon::gpu::cpp {
    template <typename FP>
    std::shared_ptr<epseon::gpu::cpp::TaskHandle<FP>>
    ComputeDeviceInterface::submitTask(std::shared_ptr<TaskConfigurator<FP>> task_config) {
        if (!task_config->isConfigured())
            throw std::runtime_error("TaskConfigurator wasn't fully configured before submitting for execution.");
        // Snapshot: the task owns a deep copy of the configuration (the copy constructor clones all
        // three parts), so the caller may go on mutating / reusing its builder while the worker runs.
        // (The reference hands the live builder to the handle and copies it inside run(),
        // vibwa.hpp:615 -- on the worker thread, i.e. after submit has returned.)
        return std::make_shared<TaskHandle<FP>>(this->shared_from_this(),
                                                std::make_shared<TaskConfigurator<FP>>(*task_config));
    }
} // namespace epseon::gpu::cpp
