// Device enumeration over the C ABI (eps_device_count / eps_device_get_props) -- replaces the
// Vulkan instance of cpp/gpu/source/epseon/gpu/compute_context.cpp:41-117.
#include "epseon/gpu/compute_context.hpp"

#include "epseon/gpu/common.hpp"
#include "epseon/gpu/device_interface.hpp"
#include "epseon_cuda.h"

#include <cstring>
#include <filesystem>
#include <fstream>
#include <stdexcept>

namespace epseon::gpu::cpp {

    // The reference logs to ./log/epseon/gpu/log.txt through spdlog (compute_context.cpp:42-45);
    // same file, plain append, failures ignored.
    void ComputeContextState::log(const std::string& line) const {
        std::error_code ec;
        std::filesystem::create_directories("./log/epseon/gpu", ec);
        if (ec) return;
        std::ofstream out("./log/epseon/gpu/log.txt", std::ios::app);
        if (out) out << "[_libepseon_gpu] " << line << "\n";
    }

    namespace {
        PhysicalDeviceInfo describe(int ordinal) {
            eps_device_props p{};
            if (eps_device_get_props(ordinal, &p) != EPS_OK) throw std::runtime_error(eps_last_error(nullptr));
            PhysicalDeviceInfo info;
            auto&              d = info.deviceProperties;
            d.apiVersion         = common::cuda_version_to_word(p.runtime_version);
            d.driverVersion      = common::cuda_version_to_word(p.driver_version);
            d.deviceID           = static_cast<uint32_t>(ordinal);
            d.deviceType         = p.integrated ? PhysicalDeviceType::eIntegratedGpu : PhysicalDeviceType::eDiscreteGpu;
            d.deviceName         = p.name;
            std::memcpy(d.pipelineCacheUUID.data(), p.uuid, 16);
            d.limits.maxComputeSharedMemorySize     = static_cast<uint32_t>(p.shared_mem_per_block_optin);
            d.limits.maxComputeWorkGroupInvocations = static_cast<uint32_t>(p.max_threads_per_block);
            for (int i = 0; i < 3; i++) {
                d.limits.maxComputeWorkGroupCount[i] = static_cast<uint32_t>(p.max_grid[i]);
                d.limits.maxComputeWorkGroupSize[i]  = static_cast<uint32_t>(p.max_block[i]);
            }
            d.smCount                = static_cast<uint32_t>(p.sm_count);
            d.computeCapabilityMajor = static_cast<uint32_t>(p.cc_major);
            d.computeCapabilityMinor = static_cast<uint32_t>(p.cc_minor);
            auto& m                  = info.memoryProperties;
            m.memoryHeapCount        = 2; // HBM + pinned host memory
            m.memoryHeaps[0]         = MemoryHeap{p.total_global_mem, eHeapDeviceLocal};
            m.memoryHeaps[1]         = MemoryHeap{0, 0};
            m.memoryTypeCount        = 2;
            m.memoryTypes[0]         = MemoryType{eDeviceLocal, 0};
            m.memoryTypes[1]         = MemoryType{eHostVisible | eHostCoherent | eHostCached, 1};
            return info;
        }
    } // namespace

    ComputeContext::ComputeContext(std::shared_ptr<ComputeContextState> state_) : state(std::move(state_)) {}

    std::shared_ptr<ComputeContext> ComputeContext::create(uint32_t /*version*/) {
        int count = 0;
        if (eps_device_count(&count) != EPS_OK || count <= 0) return nullptr;
        auto st         = std::make_shared<ComputeContextState>();
        st->deviceCount = count;
        eps_device_props p{};
        if (eps_device_get_props(0, &p) != EPS_OK) return nullptr;
        st->runtimeVersion = common::cuda_version_to_word(p.runtime_version);
        st->driverVersion  = common::cuda_version_to_word(p.driver_version);
        st->log("Discovered CUDA driver " + common::vulkan_version_to_string(st->driverVersion) + ", " +
                std::to_string(count) + " device(s)");
        return std::make_shared<ComputeContext>(st);
    }

    std::string ComputeContext::getVulkanAPIVersion() { return common::vulkan_version_to_string(state->driverVersion); }

    std::vector<PhysicalDeviceInfo> ComputeContext::getPhysicalDevicesInfo() {
        std::vector<PhysicalDeviceInfo> out;
        for (int i = 0; i < state->deviceCount; i++) out.push_back(describe(i));
        return out;
    }

    std::shared_ptr<ComputeDeviceInterface> ComputeContext::getDeviceInterface(uint32_t deviceId) {
        if (deviceId >= static_cast<uint32_t>(state->deviceCount)) throw std::runtime_error("Device not available.");
        auto dev     = std::make_shared<PhysicalDevice>();
        dev->ordinal = static_cast<int>(deviceId);
        dev->info    = describe(dev->ordinal);
        return std::make_shared<ComputeDeviceInterface>(state, dev);
    }
} // namespace epseon::gpu::cpp
