// TaskConfigurator<FP> -- builder holding the three configuration objects of a task.
// Reference: cpp/gpu/include/epseon/gpu/task_configurator/task_configurator.hpp:17-134 -- setters
// deep-clone their argument (:88-113) and return *this for chaining; copies deep-clone (:56-85);
// isConfigured() (:124-128).
#pragma once
#include "epseon/gpu/predecl.hpp"

#include "epseon/gpu/task_configurator/algorithm_config.hpp"
#include "epseon/gpu/task_configurator/hardware_config.hpp"
#include "epseon/gpu/task_configurator/potential_source.hpp"

#include <memory>
#include <stdexcept>
#include <type_traits>
#include <vector>

namespace epseon::gpu::cpp {

    template <typename FP>
    class TaskConfigurator : public std::enable_shared_from_this<TaskConfigurator<FP>> {
        static_assert(std::is_floating_point_v<FP>, "FP must be an floating-point type.");

        std::shared_ptr<HardwareConfig<FP>>  hardware_config  = {};
        std::shared_ptr<PotentialSource<FP>> potential_source = {};
        std::shared_ptr<AlgorithmConfig<FP>> algorithm_config = {};
        // additive: also return the normalised wavefunctions of the located levels
        bool wavefunction_output = false;
        // additive: rotational quantum numbers J; every curve is solved once per J with
        // V_J = V + J(J+1) hbar^2/(2 mu r^2).  Default {0}: the reference's J-less problem.
        std::vector<uint32_t> rotational_states = {0};
        // additive: this task searches only slice `shard_rank` of `shard_world` contiguous
        // slices of the coarse energy grid (energy-range sharding of ONE problem over several devices;
        // the reference gives one device per task, device_interface.hpp:22).  Default: the whole grid.
        uint32_t shard_rank = 0, shard_world = 1;
        // additive: how located brackets are refined.  0 = k-section on node counts
        // (default: bit-reproducible decisions), 1 = Cooley outward/inward matching iteration,
        // 2 = Cooley with an open (decaying) tail for near-dissociation levels.
        uint32_t level_search = 0;

        template <typename T>
        static std::shared_ptr<T> clone_or_null(const std::shared_ptr<T>& p) {
            return p ? p->shared_clone() : nullptr;
        }

      public:
        TaskConfigurator() = default;
        TaskConfigurator(std::shared_ptr<HardwareConfig<FP>> hardware_config_,
                         std::shared_ptr<PotentialSource<FP>> potential_source_,
                         std::shared_ptr<AlgorithmConfig<FP>> algorithm_config_) :
            hardware_config(std::move(hardware_config_)),
            potential_source(std::move(potential_source_)),
            algorithm_config(std::move(algorithm_config_)) {}
        TaskConfigurator(TaskConfigurator&& o) noexcept :
            std::enable_shared_from_this<TaskConfigurator<FP>>(),
            hardware_config(std::move(o.hardware_config)),
            potential_source(std::move(o.potential_source)),
            algorithm_config(std::move(o.algorithm_config)),
            wavefunction_output(o.wavefunction_output),
            rotational_states(std::move(o.rotational_states)),
            shard_rank(o.shard_rank),
            shard_world(o.shard_world),
            level_search(o.level_search) {}
        TaskConfigurator& operator=(TaskConfigurator&& o) noexcept {
            if (this != &o) {
                hardware_config     = std::move(o.hardware_config);
                potential_source    = std::move(o.potential_source);
                algorithm_config    = std::move(o.algorithm_config);
                wavefunction_output = o.wavefunction_output;
                rotational_states   = std::move(o.rotational_states);
                shard_rank          = o.shard_rank;
                shard_world         = o.shard_world;
                level_search        = o.level_search;
            }
            return *this;
        }
        TaskConfigurator(const TaskConfigurator& o) :
            std::enable_shared_from_this<TaskConfigurator<FP>>(),
            hardware_config(clone_or_null(o.hardware_config)),
            potential_source(clone_or_null(o.potential_source)),
            algorithm_config(clone_or_null(o.algorithm_config)),
            wavefunction_output(o.wavefunction_output),
            rotational_states(o.rotational_states),
            shard_rank(o.shard_rank),
            shard_world(o.shard_world),
            level_search(o.level_search) {}
        TaskConfigurator& operator=(const TaskConfigurator& o) {
            if (this != &o) {
                hardware_config     = clone_or_null(o.hardware_config);
                potential_source    = clone_or_null(o.potential_source);
                algorithm_config    = clone_or_null(o.algorithm_config);
                wavefunction_output = o.wavefunction_output;
                rotational_states   = o.rotational_states;
                shard_rank          = o.shard_rank;
                shard_world         = o.shard_world;
                level_search        = o.level_search;
            }
            return *this;
        }
        ~TaskConfigurator() = default;

        TaskConfigurator& setHardwareConfig(std::shared_ptr<HardwareConfig<FP>> cfg) {
            hardware_config = cfg->shared_clone();
            return *this;
        }
        [[nodiscard]] std::shared_ptr<HardwareConfig<FP>> getHardwareConfig() const { return hardware_config; }

        TaskConfigurator& setPotentialSource(std::shared_ptr<PotentialSource<FP>> ps) {
            potential_source = ps->shared_clone();
            return *this;
        }
        [[nodiscard]] std::shared_ptr<PotentialSource<FP>> getPotentialSource() const { return potential_source; }

        TaskConfigurator& setAlgorithmConfig(std::shared_ptr<AlgorithmConfig<FP>> ac) {
            algorithm_config = ac->shared_clone();
            return *this;
        }
        [[nodiscard]] std::shared_ptr<AlgorithmConfig<FP>> getAlgorithmConfig() const { return algorithm_config; }

        TaskConfigurator& setWavefunctionOutput(bool enabled) {
            wavefunction_output = enabled;
            return *this;
        }
        [[nodiscard]] bool getWavefunctionOutput() const { return wavefunction_output; }

        TaskConfigurator& setRotationalStates(std::vector<uint32_t> j_values) {
            if (j_values.empty()) throw std::runtime_error("rotational_states must hold at least one J");
            for (const uint32_t j : j_values)
                if (j >= (1u << 26)) throw std::runtime_error("rotational quantum numbers must be below 2^26");
            rotational_states = std::move(j_values);
            return *this;
        }
        [[nodiscard]] const std::vector<uint32_t>& getRotationalStates() const { return rotational_states; }

        TaskConfigurator& setEnergyShard(uint32_t rank, uint32_t world) {
            if (world == 0 || rank >= world) throw std::runtime_error("energy shard: need rank < world");
            shard_rank  = rank;
            shard_world = world;
            return *this;
        }
        TaskConfigurator& setLevelSearch(uint32_t mode) {
            if (mode > 2) throw std::runtime_error("level search: 0 (k-section), 1 (cooley) or 2 (cooley, open tail)");
            level_search = mode;
            return *this;
        }
        [[nodiscard]] uint32_t getLevelSearch() const { return level_search; }
        [[nodiscard]] uint32_t getEnergyShardRank() const { return shard_rank; }
        [[nodiscard]] uint32_t getEnergyShardWorld() const { return shard_world; }

        [[nodiscard]] bool isConfigured() const {
            return static_cast<bool>(hardware_config) && static_cast<bool>(potential_source) &&
                   static_cast<bool>(algorithm_config);
        }

        [[nodiscard]] std::vector<ShaderBuffersRequirements<FP>> getShaderBufferRequirements() const {
            return algorithm_config->getShaderBufferRequirements(*this);
        }
    };
} // namespace epseon::gpu::cpp
