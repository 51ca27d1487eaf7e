// Algorithm configuration of a task.
// Reference: cpp/gpu/include/epseon/gpu/task_configurator/algorithm_config.hpp --
// ShaderBuffersRequirements<FP> (:16-37), AlgorithmConfig<FP> (:40-70), VibwaAlgorithmConfig<FP>
// (:76-199): same members, getters, equals/clone helpers and buffer-plan arithmetic.  The six Vibwa
// parameters are the algorithm's inputs: masses in amu, energies in cm^-1;
// integration_step is stored but the integration runs on the potential's own grid.
#pragma once
#include "epseon/gpu/predecl.hpp"

#include "epseon/gpu/algorithms/algorithm.hpp"
#include "epseon/gpu/algorithms/vibwa.hpp"

#include <cstdint>
#include <memory>
#include <type_traits>
#include <vector>

namespace epseon::gpu::cpp {

    // Buffer plan of one work item (kept for API compatibility; the CUDA path sizes its own buffers).
    template <typename FP>
    struct ShaderBuffersRequirements {
        uint32_t stagingBuffersCount               = {};
        uint32_t stagingBuffersElementCount        = {};
        uint32_t gpuOnlyStorageBuffersCount        = {};
        uint32_t gpuOnlyStorageBuffersElementCount = {};
        uint32_t outputBuffersCount                = {};
        uint32_t outputBuffersElementCount         = {};

        [[nodiscard]] uint32_t getStagingBuffersSizeBytes() const {
            return stagingBuffersCount * stagingBuffersElementCount * sizeof(FP);
        }
        [[nodiscard]] uint32_t getGpuOnlyStorageBufferSizeBytes() const {
            return gpuOnlyStorageBuffersCount * gpuOnlyStorageBuffersElementCount * sizeof(FP);
        }
        [[nodiscard]] uint32_t getOutputBufferSizeBytes() const {
            return outputBuffersCount * outputBuffersElementCount * sizeof(FP);
        }
    };

    template <typename FP>
    class AlgorithmConfig : public std::enable_shared_from_this<AlgorithmConfig<FP>> {
        static_assert(std::is_floating_point_v<FP>, "FP must be an floating-point type.");

      public:
        AlgorithmConfig() noexcept = default;
        AlgorithmConfig(const AlgorithmConfig&) noexcept : std::enable_shared_from_this<AlgorithmConfig<FP>>() {}
        AlgorithmConfig& operator=(const AlgorithmConfig&) noexcept { return *this; }
        virtual ~AlgorithmConfig() = default;

        virtual bool equals(const AlgorithmConfig<FP>& other) const                          = 0;
        [[nodiscard]] virtual std::shared_ptr<Algorithm<FP>>       getImplementation() const = 0;
        [[nodiscard]] virtual std::shared_ptr<AlgorithmConfig<FP>> shared_clone() const      = 0;
        [[nodiscard]] virtual std::unique_ptr<AlgorithmConfig<FP>> unique_clone() const      = 0;
        virtual std::vector<ShaderBuffersRequirements<FP>>
        getShaderBufferRequirements(const TaskConfigurator<FP>& config) const = 0;
    };

    template <typename FP>
    class VibwaAlgorithmConfig : public AlgorithmConfig<FP> {
        FP       mass_atom_0               = 0;
        FP       mass_atom_1               = 0;
        FP       integration_step          = 0;
        FP       min_distance_to_asymptote = 0;
        uint32_t min_level                 = 0;
        uint32_t max_level                 = 0;

      public:
        VibwaAlgorithmConfig() noexcept = default;
        VibwaAlgorithmConfig(FP mass_atom_0_, FP mass_atom_1_, FP integration_step_, FP min_distance_to_asymptote_,
                             uint32_t min_level_, uint32_t max_level_) :
            mass_atom_0(mass_atom_0_),
            mass_atom_1(mass_atom_1_),
            integration_step(integration_step_),
            min_distance_to_asymptote(min_distance_to_asymptote_),
            min_level(min_level_),
            max_level(max_level_) {}
        ~VibwaAlgorithmConfig() override = default;

        bool equals(const AlgorithmConfig<FP>& other) const override {
            const auto* o = dynamic_cast<const VibwaAlgorithmConfig<FP>*>(&other);
            return o != nullptr && mass_atom_0 == o->mass_atom_0 && mass_atom_1 == o->mass_atom_1 &&
                   integration_step == o->integration_step &&
                   min_distance_to_asymptote == o->min_distance_to_asymptote && min_level == o->min_level &&
                   max_level == o->max_level;
        }

        [[nodiscard]] std::shared_ptr<Algorithm<FP>> getImplementation() const override {
            return std::make_shared<VibwaAlgorithm<FP>>();
        }
        [[nodiscard]] std::shared_ptr<AlgorithmConfig<FP>> shared_clone() const override {
            return std::make_shared<VibwaAlgorithmConfig<FP>>(*this);
        }
        [[nodiscard]] std::unique_ptr<AlgorithmConfig<FP>> unique_clone() const override {
            return std::make_unique<VibwaAlgorithmConfig<FP>>(*this);
        }

        FP                     getMassAtom0() const { return mass_atom_0; }
        FP                     getMassAtom1() const { return mass_atom_1; }
        FP                     getIntegrationStep() const { return integration_step; }
        FP                     getMinDistanceToAsymptote() const { return min_distance_to_asymptote; }
        [[nodiscard]] uint32_t getMinLevel() const { return min_level; }
        [[nodiscard]] uint32_t getMaxLevel() const { return max_level; }
        // level_count = max_level - min_level + 1: length of the per-curve output (reference :177).
        [[nodiscard]] uint32_t getLevelCount() const { return (max_level - min_level) + 1; }

        // group_size identical plans: 1 staging + 5 device arrays of potential_buffer_size elements and
        // one output of level_count elements per work item (reference :173-198).
        std::vector<ShaderBuffersRequirements<FP>>
        getShaderBufferRequirements(const TaskConfigurator<FP>& config) const override;
    };

    template <typename FP>
    bool operator==(const AlgorithmConfig<FP>& lhs, const AlgorithmConfig<FP>& rhs) {
        return lhs.equals(rhs);
    }
    template <typename FP>
    bool operator==(const VibwaAlgorithmConfig<FP>& lhs, const VibwaAlgorithmConfig<FP>& rhs) {
        return lhs.equals(rhs);
    }
} // namespace epseon::gpu::cpp

// Include closure of the reference: there, this header pulls in algorithms/vibwa.hpp (:7), which
// pulls in task_handle.hpp (vibwa.hpp:11) and through it task_configurator.hpp and
// device_interface.hpp -- a translation unit that includes ONLY algorithm_config.hpp sees complete
// TaskConfigurator / TaskHandle types and the definitions of VibwaAlgorithm<FP>::run and
// getShaderBufferRequirements (the reference's test_task_configurator.cpp and
// test_algorithm_confgu.cpp rely on exactly that).  Same closure here, placed after the class
// definitions so that it is order-independent.
#include "epseon/gpu/task_configurator/task_configurator.hpp"
#include "epseon/gpu/task_handle.hpp"

namespace epseon::gpu::cpp {
    template <typename FP>
    std::vector<ShaderBuffersRequirements<FP>>
    VibwaAlgorithmConfig<FP>::getShaderBufferRequirements(const TaskConfigurator<FP>& config) const {
        const auto hw = config.getHardwareConfig();
        ShaderBuffersRequirements<FP> one{};
        one.stagingBuffersCount               = 1;
        one.stagingBuffersElementCount        = hw->getPotentialBufferSize();
        one.gpuOnlyStorageBuffersCount        = 5;
        one.gpuOnlyStorageBuffersElementCount = hw->getPotentialBufferSize();
        one.outputBuffersCount                = 1;
        one.outputBuffersElementCount         = getLevelCount();
        return std::vector<ShaderBuffersRequirements<FP>>(hw->getGroupSize(), one);
    }
} // namespace epseon::gpu::cpp
