// pybind11 module _libepseon_gpu -- the reference's Python surface (15 classes) over the CUDA path.
// Reference: cpp/gpu/source/epseon/gpu/python/api.cpp:114-494 (class names, property / method names,
// keyword arguments, exception types); tests: python/test/test_device/test_gpu/test_libepseon_gpu.py.
#include "epseon/gpu/python/api.hpp"

#include "epseon/gpu/common.hpp"

#include <pybind11/numpy.h>
#include <pybind11/pybind11.h>
#include <pybind11/stl.h>

#include <cstring>
#include <tuple>

namespace py = pybind11;

namespace epseon::gpu::python {

    TaskConfiguratorVariant ComputeDeviceInterface::get_task_configurator(const std::string& precision) {
        cpp::PrecisionType kind;
        try {
            kind = cpp::toPrecisionType(precision);
        } catch (const cpp::InvalidPrecisionTypeString& e) {
            throw py::value_error(e.what()); // ValueError, message asserted by the reference's tests
        }
        PrecisionTypeAssertValueCount(2);
        if (kind == cpp::PrecisionType::Float32) return TaskConfigurator<float>{device->getTaskConfigurator<float>()};
        return TaskConfigurator<double>{device->getTaskConfigurator<double>()};
    }

    EpseonComputeContext EpseonComputeContext::create() {
        auto application = cpp::ComputeContext::create();
        if (!application) throw std::runtime_error("Failed to create EpseonComputeContext.");
        return EpseonComputeContext{application};
    }

    std::string EpseonComputeContext::get_vulkan_version() { return application->getVulkanAPIVersion(); }

    std::vector<cpp::PhysicalDeviceInfo> EpseonComputeContext::get_physical_device_info() {
        return application->getPhysicalDevicesInfo();
    }

    ComputeDeviceInterface EpseonComputeContext::get_device_interface(uint32_t device_id) {
        return ComputeDeviceInterface{application->getDeviceInterface(device_id)};
    }

    namespace {
        template <typename H>
        void bind_task_handle(py::module_& m, const char* name) {
            py::class_<H>(m, name)
                .def("get_status_message", &H::get_status_message,
                     "Get status message explaining current execution stage.")
                .def("is_done", &H::is_done, "Check if task already finished execution.")
                .def("wait", &H::wait, py::call_guard<py::gil_scoped_release>(), "Block and wait for task to finish.")
                .def("is_running", &H::is_running, "Check if task is still running.")
                .def("cancel", &H::cancel,
                     "Request cooperative cancellation; True when the request reached a running task. The solve "
                     "stops between refinement rounds (queued sweeps drain), is_done() becomes True, the status "
                     "reads 'cancelled'.")
                .def("was_cancelled", &H::was_cancelled, "True when the task stopped on a cancel() request.")
                .def("get_levels", &H::get_levels,
                     "Vibrational level energies [curve][level - min_level] (NaN where a level was not found).")
                .def("get_level_counts", &H::get_level_counts, "Number of levels below the search ceiling, per curve.")
                .def("get_wavefunctions",
                     [](H& h) {
                         uint32_t   d[3] = {0, 0, 0};
                         const auto psi  = h.getHandle()->getWavefunctions(d);
                         py::array_t<double> out({static_cast<py::ssize_t>(d[0]), static_cast<py::ssize_t>(d[1]),
                                                  static_cast<py::ssize_t>(d[2])});
                         if (!psi.empty()) std::memcpy(out.mutable_data(), psi.data(), psi.size() * sizeof(double));
                         return out;
                     },
                     "Normalised wavefunctions as a numpy array [curve][level - min_level][grid point] "
                     "(empty unless set_wavefunction_output(True) was configured).")
                .def("has_failed", &H::has_failed, "True when the worker stopped with an error (see status message).")
                .def("get_device_milliseconds", &H::get_device_milliseconds, "CUDA-event time of the level solve.")
                .def("get_search_parameters", &H::get_search_parameters,
                     "(n_coarse, refine_points, max_rounds, rel_tol) of the level search the task ran: 1024-point "
                     "coarse grid and 256 points per level per round for a few curves, fewer points per round for "
                     "batches of hundreds of curves.")
                .doc() = "Handle object for referencing GPU compute task.";
        }

        template <typename C>
        void bind_task_configurator(py::module_& m, const char* name) {
            py::class_<C>(m, name)
                .def("set_hardware_config", &C::set_hardware_config, py::arg("potential_buffer_size"),
                     py::arg("group_size"), py::arg("allocation_block_size"), py::return_value_policy::reference,
                     "Set hardware configuration for a GPU compute task.")
                .def("set_morse_potential", &C::set_morse_potential, py::arg("configurations"),
                     py::return_value_policy::reference,
                     "Set potential data source configuration for GPU compute task.")
                .def("set_wavefunction_output", &C::set_wavefunction_output, py::arg("enabled"),
                     py::return_value_policy::reference,
                     "Also compute the normalised wavefunctions of the located levels (TaskHandle.get_wavefunctions).")
                .def("set_rotational_states", &C::set_rotational_states, py::arg("j_values"),
                     py::return_value_policy::reference,
                     "Solve every curve once per rotational quantum number J (effective potential "
                     "V + J(J+1) hbar^2 / (2 mu r^2)); get_levels() rows are then ordered curve-major, "
                     "row = curve * len(j_values) + j.  Default [0].")
                .def("set_potential_tables",
                     [](C& self, const py::array_t<double, py::array::c_style | py::array::forcecast>& tables, double min_r,
                        double max_r) -> C& {
                         if (tables.ndim() != 2) throw std::runtime_error("tables must be a 2-D array [n_curves][point_count]");
                         std::vector<std::vector<double>> rows(static_cast<size_t>(tables.shape(0)));
                         for (py::ssize_t c = 0; c < tables.shape(0); c++)
                             rows[static_cast<size_t>(c)].assign(tables.data(c, 0), tables.data(c, 0) + tables.shape(1));
                         return self.set_potential_tables(std::move(rows), min_r, max_r);
                     },
                     py::arg("tables"), py::arg("min_r"), py::arg("max_r"), py::return_value_policy::reference,
                     "Use curves held in memory: a float64 array [n_curves][point_count] of V(r_i) on the uniform grid "
                     "r_i = min_r + i (max_r - min_r)/(point_count - 1).")
                .def("set_level_search", &C::set_level_search, py::arg("mode"), py::return_value_policy::reference,
                     "How bracketed levels are refined: 'ksection' (default; sweeps of trial energies, decisions on node "
                     "counts), 'cooley' (outward/inward matching iteration: 3-5 iterations per level), 'cooley_open' (the "
                     "same with a decaying tail instead of a wall at max_r, for levels near dissociation).")
                .def("set_energy_shard", &C::set_energy_shard, py::arg("rank"), py::arg("world"),
                     py::return_value_policy::reference,
                     "Search only slice `rank` of `world` contiguous slices of the coarse energy grid (energy-range "
                     "sharding of one problem over several devices); the union over the ranks equals the unsharded task.")
                .def("set_potential_files", &C::set_potential_files, py::arg("file_names"), py::arg("point_count") = 0,
                     py::return_value_policy::reference,
                     "Use tabulated 'r V' text files (or NumPy .npy arrays of shape (n, 2)) as potential source; "
                     "tables on a non-uniform grid (or any "
                     "table when point_count > 0) are resampled with a natural cubic spline.")
                .def("set_vibwa_algorithm", &C::set_vibwa_algorithm, py::arg("mass_atom_0"), py::arg("mass_atom_1"),
                     py::arg("integration_step"), py::arg("min_distance_to_asymptote"), py::arg("min_level"),
                     py::arg("max_level"), py::return_value_policy::reference,
                     "Set algorithm configuration for a GPU compute task.")
                .def("is_configured", &C::is_configured,
                     "Check if this instance is fully configured, i.e. it has been assigned a valid hardware "
                     "configuration, potential source and algorithm config.")
                .doc() = "Builder for configuring GPU compute task.";
        }
    } // namespace

    PYBIND11_MODULE(_libepseon_gpu, m) {
        m.doc() = "Sub package for interacting with GPU compute capabilities (B200-native CUDA build).";

        py::class_<cpp::PhysicalDeviceSparseProperties>(m, "PhysicalDeviceSparseProperties").doc() =
            "Placeholder kept for API compatibility (no sparse resources on the CUDA path).";

        py::class_<cpp::PhysicalDeviceLimits>(m, "PhysicalDeviceLimits")
            .def_property_readonly("max_compute_shared_memory_size",
                                   [](const cpp::PhysicalDeviceLimits& l) { return l.maxComputeSharedMemorySize; })
            .def_property_readonly("max_compute_work_group_count",
                                   [](const cpp::PhysicalDeviceLimits& l) -> std::tuple<uint32_t, uint32_t, uint32_t> {
                                       return {l.maxComputeWorkGroupCount[0], l.maxComputeWorkGroupCount[1],
                                               l.maxComputeWorkGroupCount[2]};
                                   })
            .def_property_readonly("max_compute_work_group_invocations",
                                   [](const cpp::PhysicalDeviceLimits& l) { return l.maxComputeWorkGroupInvocations; })
            // the reference returns the shared-memory size here (api.cpp:147-152); the test only
            // requires an int -- report the largest CTA dimension instead
            .def_property_readonly("max_compute_work_group_size",
                                   [](const cpp::PhysicalDeviceLimits& l) { return l.maxComputeWorkGroupSize[0]; })
            .doc() = "Physical device limits - mostly max counts of different resources.";

        py::class_<cpp::PhysicalDeviceProperties>(m, "PhysicalDeviceProperties")
            .def_property_readonly("api_version",
                                   [](const cpp::PhysicalDeviceProperties& p) {
                                       return common::vulkan_version_to_string(p.apiVersion);
                                   })
            .def_property_readonly("driver_version",
                                   [](const cpp::PhysicalDeviceProperties& p) {
                                       return common::vulkan_version_to_string(p.driverVersion);
                                   })
            .def_property_readonly("vendor_id", [](const cpp::PhysicalDeviceProperties& p) { return p.vendorID; })
            .def_property_readonly("device_id", [](const cpp::PhysicalDeviceProperties& p) { return p.deviceID; })
            .def_property_readonly("device_type",
                                   [](const cpp::PhysicalDeviceProperties& p) -> std::string {
                                       switch (p.deviceType) {
                                           case cpp::PhysicalDeviceType::eIntegratedGpu: return "INTEGRATED_GPU";
                                           case cpp::PhysicalDeviceType::eDiscreteGpu: return "DISCRETE_GPU";
                                           case cpp::PhysicalDeviceType::eVirtualGpu: return "VIRTUAL_GPU";
                                           case cpp::PhysicalDeviceType::eCpu: return "CPU";
                                           case cpp::PhysicalDeviceType::eOther: return "OTHER";
                                       }
                                       throw std::runtime_error("Unknown physical device type.");
                                   })
            .def_property_readonly("device_name", [](const cpp::PhysicalDeviceProperties& p) { return p.deviceName; })
            .def_property_readonly("pipeline_cache_uuid",
                                   [](const cpp::PhysicalDeviceProperties& p) {
                                       return std::vector<uint8_t>(p.pipelineCacheUUID.begin(),
                                                                   p.pipelineCacheUUID.end());
                                   })
            .def_property_readonly("limits", [](const cpp::PhysicalDeviceProperties& p) { return p.limits; })
            .def_property_readonly("sparse_properties",
                                   [](const cpp::PhysicalDeviceProperties& p) { return p.sparseProperties; })
            .def_property_readonly("sm_count", [](const cpp::PhysicalDeviceProperties& p) { return p.smCount; })
            .def_property_readonly("compute_capability",
                                   [](const cpp::PhysicalDeviceProperties& p) -> std::tuple<uint32_t, uint32_t> {
                                       return {p.computeCapabilityMajor, p.computeCapabilityMinor};
                                   })
            .doc() = "Properties of physical device retrieved from the CUDA runtime.";

        py::class_<cpp::MemoryHeap>(m, "MemoryHeap")
            .def_property_readonly("size", [](const cpp::MemoryHeap& h) { return h.size; })
            .def_property_readonly("flags",
                                   [](const cpp::MemoryHeap& h) {
                                       std::vector<std::string> flags;
                                       if (h.flags & cpp::eHeapDeviceLocal) flags.emplace_back("DEVICE_LOCAL");
                                       if (h.flags & cpp::eHeapMultiInstance) flags.emplace_back("MULTI_INSTANCE");
                                       return flags;
                                   })
            .doc() = "Memory heap description.";

        py::class_<cpp::MemoryType>(m, "MemoryType")
            .def_property_readonly("heap_index", [](const cpp::MemoryType& t) { return t.heapIndex; })
            .def_property_readonly("flags",
                                   [](const cpp::MemoryType& t) {
                                       std::vector<std::string> flags;
                                       if (t.propertyFlags & cpp::eDeviceLocal) flags.emplace_back("DEVICE_LOCAL");
                                       if (t.propertyFlags & cpp::eHostVisible) flags.emplace_back("HOST_VISIBLE");
                                       if (t.propertyFlags & cpp::eHostCoherent) flags.emplace_back("HOST_COHERENT");
                                       if (t.propertyFlags & cpp::eHostCached) flags.emplace_back("HOST_CACHED");
                                       if (t.propertyFlags & cpp::eLazilyAllocated)
                                           flags.emplace_back("LAZILY_ALLOCATED");
                                       if (t.propertyFlags & cpp::eProtected) flags.emplace_back("PROTECTED");
                                       return flags;
                                   })
            .doc() = "Memory type description.";

        py::class_<cpp::PhysicalDeviceMemoryProperties>(m, "PhysicalDeviceMemoryProperties")
            .def_property_readonly("memory_heaps",
                                   [](const cpp::PhysicalDeviceMemoryProperties& p) {
                                       return std::vector<cpp::MemoryHeap>(p.memoryHeaps.begin(),
                                                                           p.memoryHeaps.begin() + p.memoryHeapCount);
                                   })
            .def_property_readonly("memory_types",
                                   [](const cpp::PhysicalDeviceMemoryProperties& p) {
                                       return std::vector<cpp::MemoryType>(p.memoryTypes.begin(),
                                                                           p.memoryTypes.begin() + p.memoryTypeCount);
                                   })
            .doc() = "Memory properties of a physical device.";

        py::class_<cpp::PhysicalDeviceInfo>(m, "PhysicalDeviceInfo")
            .def_property_readonly("device_properties", [](const cpp::PhysicalDeviceInfo& i) { return i.deviceProperties; })
            .def_property_readonly("memory_properties", [](const cpp::PhysicalDeviceInfo& i) { return i.memoryProperties; })
            .doc() = "Container for physical device info.";

        bind_task_handle<TaskHandleFloat32>(m, "TaskHandleFloat32");
        bind_task_handle<TaskHandleFloat64>(m, "TaskHandleFloat64");

        py::class_<MorsePotentialConfig>(m, "MorsePotentialConfig")
            .def(py::init(&MorsePotentialConfig::create), py::arg("dissociation_energy"),
                 py::arg("equilibrium_bond_distance"), py::arg("well_width"), py::arg("min_r"), py::arg("max_r"),
                 py::arg("point_count"), "Create instance of MorsePotentialConfig class.")
            .doc() = "Configuration of single Morse potential curve.";

        bind_task_configurator<TaskConfiguratorFloat32>(m, "TaskConfiguratorFloat32");
        bind_task_configurator<TaskConfiguratorFloat64>(m, "TaskConfiguratorFloat64");

        py::class_<ComputeDeviceInterface>(m, "ComputeDeviceInterface")
            .def("get_task_configurator", &ComputeDeviceInterface::get_task_configurator,
                 "Get builder instance for configuring GPU compute task.")
            .def("submit_task", &ComputeDeviceInterface::submit_task<float>,
                 "Submit task for execution. Will raise RuntimeError upon receiving not fully configured "
                 "TaskConfigurator.")
            .def("submit_task", &ComputeDeviceInterface::submit_task<double>,
                 "Submit task for execution. Will raise RuntimeError upon receiving not fully configured "
                 "TaskConfigurator.")
            .doc() = "Interface to particular CUDA device.";

        m.def("release_device_memory", &cpp::detail::trim_context_pool,
              "Destroy the idle pooled CUDA contexts of finished tasks (their device memory is returned).");

        py::class_<EpseonComputeContext>(m, "EpseonComputeContext")
            .def_static("create", &EpseonComputeContext::create, "Create instance of the compute context.")
            .def("get_vulkan_version", &EpseonComputeContext::get_vulkan_version,
                 "Get API version string (CUDA driver version in the reference's v.M.m.p format).")
            .def("get_physical_device_info", &EpseonComputeContext::get_physical_device_info,
                 "Get information about available physical devices.")
            .def("get_device_interface", &EpseonComputeContext::get_device_interface,
                 "Get interface for running algorithms on a CUDA device.")
            .doc() = "Compute context handle.";
    }
} // namespace epseon::gpu::python
