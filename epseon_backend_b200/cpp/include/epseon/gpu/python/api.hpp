// Python-facing wrapper classes of module _libepseon_gpu.
// Reference: cpp/gpu/include/epseon/gpu/python/api.hpp -- python::TaskHandle<FP> (:27-72),
// python::MorsePotentialConfig (:82-118), python::TaskConfigurator<FP> (:122-254),
// python::ComputeDeviceInterface (:265-290), python::EpseonComputeContext (:292-309): same class and
// method names, argument meaning and exception types.  Deviations: Q1 fixed
// (N configs stored, not 2N), additive result accessors.
#pragma once
#include "epseon/gpu/predecl.hpp"

#include "epseon/gpu/compute_context.hpp"
#include "epseon/gpu/device_interface.hpp"
#include "epseon/gpu/enums.hpp"
#include "epseon/gpu/task_configurator/task_configurator.hpp"
#include "epseon/gpu/task_handle.hpp"

#include <cstdint>
#include <memory>
#include <optional>
#include <stdexcept>
#include <string>
#include <variant>
#include <tuple>
#include <vector>

namespace epseon::gpu::python {

    template <typename FP>
    class TaskHandle {
        std::shared_ptr<cpp::TaskHandle<FP>> handle = {};

      public:
        // Submitting from Python starts the worker immediately (reference :32-36).
        explicit TaskHandle(std::shared_ptr<cpp::TaskHandle<FP>> handle_) : handle(std::move(handle_)) {
            if (!handle->isRunning()) handle->startWorker();
        }

        std::string get_status_message() { return handle->getStatusMessage(); }
        bool        is_done() { return handle->isDone(); }
        bool        is_running() { return handle->isRunning(); }
        bool        cancel() { return handle->cancel(); }
        bool        was_cancelled() { return handle->wasCancelled(); }
        void        wait() { handle->wait(); }
        // additive
        std::vector<std::vector<FP>> get_levels() { return handle->getLevels(); }
        std::vector<uint32_t>        get_level_counts() { return handle->getLevelCounts(); }
        std::shared_ptr<cpp::TaskHandle<FP>> getHandle() const { return handle; }
        bool                         has_failed() { return handle->hasFailed(); }
        double                       get_device_milliseconds() { return handle->getDeviceMilliseconds(); }
        // (n_coarse, refine_points, max_rounds, rel_tol) the level search ran with
        std::tuple<uint32_t, uint32_t, uint32_t, double> get_search_parameters() {
            uint32_t v[3] = {0, 0, 0};
            double   tol  = 0.0;
            handle->getSearchParameters(v, tol);
            return {v[0], v[1], v[2], tol};
        }
    };

    using TaskHandleFloat32 = TaskHandle<float>;
    using TaskHandleFloat64 = TaskHandle<double>;
    using TaskHandleVariant = std::variant<TaskHandleFloat32, TaskHandleFloat64>;

    class MorsePotentialConfig {
        cpp::MorsePotentialConfig<double> configuration;

      public:
        MorsePotentialConfig() = default;
        explicit MorsePotentialConfig(cpp::MorsePotentialConfig<double> configuration_) :
            configuration(std::move(configuration_)) {}

        static MorsePotentialConfig create(double dissociation_energy, double equilibrium_bond_distance,
                                           double well_width, double min_r, double max_r, uint32_t point_count) {
            return MorsePotentialConfig{cpp::MorsePotentialConfig<double>(
                dissociation_energy, equilibrium_bond_distance, well_width, min_r, max_r, point_count)};
        }

        [[nodiscard]] const cpp::MorsePotentialConfig<double>& getConfiguration() const { return configuration; }
    };

    template <typename FP>
    class TaskConfigurator {
        std::shared_ptr<cpp::TaskConfigurator<FP>> configurator = {};

      public:
        explicit TaskConfigurator(std::shared_ptr<cpp::TaskConfigurator<FP>> configurator_) :
            configurator(std::move(configurator_)) {}

        TaskConfigurator& set_hardware_config(uint32_t potential_buffer_size, uint32_t group_size,
                                              uint32_t allocation_block_size) {
            configurator->setHardwareConfig(
                std::make_shared<cpp::HardwareConfig<FP>>(potential_buffer_size, group_size, allocation_block_size));
            return *this;
        }

        // All curves must share one point_count (reference :185-199); values arrive as double and are
        // cast to FP (reference :202-210).
        TaskConfigurator& set_morse_potential(const std::vector<MorsePotentialConfig>& configurations) {
            std::optional<uint32_t>                    point_count;
            std::vector<cpp::MorsePotentialConfig<FP>> converted;
            converted.reserve(configurations.size());
            for (const auto& element : configurations) {
                const auto&    c = element.getConfiguration();
                const uint32_t n = c.getPointCount();
                if (point_count.has_value() && *point_count != n)
                    throw std::runtime_error("All Morse potentials must have same point count, but previous ones had " +
                                             std::to_string(*point_count) + " and current one has " +
                                             std::to_string(n) + ".");
                point_count = n;
                converted.emplace_back(static_cast<FP>(c.getDissociationEnergy()),
                                       static_cast<FP>(c.getEquilibriumBondDistance()),
                                       static_cast<FP>(c.getWellWidth()), static_cast<FP>(c.getMinR()),
                                       static_cast<FP>(c.getMaxR()), n);
            }
            configurator->setPotentialSource(std::make_shared<cpp::MorsePotentialGenerator<FP>>(std::move(converted)));
            return *this;
        }

        // Additive: ask the task to also return normalised wavefunctions.
        TaskConfigurator& set_wavefunction_output(bool enabled) {
            configurator->setWavefunctionOutput(enabled);
            return *this;
        }

        // Additive: solve every curve once per rotational quantum number J
        // (centrifugal term J(J+1) hbar^2 / 2 mu r^2); result rows are ordered [curve][J].
        TaskConfigurator& set_rotational_states(const std::vector<uint32_t>& j_values) {
            configurator->setRotationalStates(j_values);
            return *this;
        }

        // Additive: curves already in memory, rows of a [n_curves][point_count] table on a uniform grid.
        TaskConfigurator& set_potential_tables(std::vector<std::vector<double>> tables, double min_r, double max_r) {
            configurator->setPotentialSource(
                std::make_shared<cpp::TabulatedPotentialSource<FP>>(std::move(tables), min_r, max_r));
            return *this;
        }

        // Additive: tabulated curves from "r V" text files.
        // additive: "ksection" (default), "cooley", "cooley_open"
        TaskConfigurator& set_level_search(const std::string& mode) {
            if (mode == "ksection") configurator->setLevelSearch(0);
            else if (mode == "cooley") configurator->setLevelSearch(1);
            else if (mode == "cooley_open") configurator->setLevelSearch(2);
            else throw std::runtime_error("level search mode must be 'ksection', 'cooley' or 'cooley_open'");
            return *this;
        }

        // additive: energy-range sharding of one problem over several devices (multi.py)
        TaskConfigurator& set_energy_shard(uint32_t rank, uint32_t world) {
            configurator->setEnergyShard(rank, world);
            return *this;
        }

        TaskConfigurator& set_potential_files(const std::vector<std::string>& file_names, uint32_t point_count) {
            configurator->setPotentialSource(std::make_shared<cpp::PotentialFileLoader<FP>>(file_names, point_count));
            return *this;
        }

        TaskConfigurator& set_vibwa_algorithm(double mass_atom_0, double mass_atom_1, double integration_step,
                                              double min_distance_to_asymptote, uint32_t min_level,
                                              uint32_t max_level) {
            configurator->setAlgorithmConfig(std::make_shared<cpp::VibwaAlgorithmConfig<FP>>(
                static_cast<FP>(mass_atom_0), static_cast<FP>(mass_atom_1), static_cast<FP>(integration_step),
                static_cast<FP>(min_distance_to_asymptote), min_level, max_level));
            return *this;
        }

        [[nodiscard]] bool is_configured() const { return configurator->isConfigured(); }
        [[nodiscard]] std::shared_ptr<cpp::TaskConfigurator<FP>> getTaskConfigurator() const { return configurator; }
    };

    using TaskConfiguratorFloat32 = TaskConfigurator<float>;
    using TaskConfiguratorFloat64 = TaskConfigurator<double>;
    using TaskConfiguratorVariant = std::variant<TaskConfiguratorFloat32, TaskConfiguratorFloat64>;

    class ComputeDeviceInterface {
        std::shared_ptr<cpp::ComputeDeviceInterface> device;

      public:
        explicit ComputeDeviceInterface(std::shared_ptr<cpp::ComputeDeviceInterface> device_) :
            device(std::move(device_)) {}

        TaskConfiguratorVariant get_task_configurator(const std::string& precision);

        template <typename FP>
        TaskHandleVariant submit_task(const TaskConfigurator<FP>& task_config) {
            if (!task_config.is_configured())
                throw std::runtime_error("TaskConfigurator submitted for execution before fully configured.");
            return TaskHandleVariant{TaskHandle<FP>{device->submitTask(task_config.getTaskConfigurator())}};
        }
    };

    class EpseonComputeContext {
      public:
        std::shared_ptr<cpp::ComputeContext> application = {};

        explicit EpseonComputeContext(std::shared_ptr<cpp::ComputeContext> application_) :
            application(std::move(application_)) {}

        static EpseonComputeContext create();

        std::string                          get_vulkan_version();
        std::vector<cpp::PhysicalDeviceInfo> get_physical_device_info();
        ComputeDeviceInterface               get_device_interface(uint32_t device_id);
    };
} // namespace epseon::gpu::python
