// ComputeContext -- device enumeration and per-device interfaces, over the CUDA runtime (through the
// C ABI) instead of a Vulkan instance.
// Reference: cpp/gpu/include/epseon/gpu/compute_context.hpp:20-72 and
// cpp/gpu/source/epseon/gpu/compute_context.cpp:41-117 -- same class / method names
// (create, getVulkanAPIVersion, getPhysicalDevicesInfo, getDeviceInterface) and the same
// "Device not available." error.  The property structs keep the Vulkan field names the reference's
// bindings read (api.cpp:118-329) so that the Python surface is unchanged.
#pragma once
#include "epseon/gpu/predecl.hpp"

#include <array>
#include <cstdint>
#include <memory>
#include <string>
#include <vector>

namespace epseon::gpu::cpp {

    enum class PhysicalDeviceType { eOther, eIntegratedGpu, eDiscreteGpu, eVirtualGpu, eCpu };

    struct PhysicalDeviceLimits {
        uint32_t                maxComputeSharedMemorySize     = 0; // opt-in dynamic smem per CTA (bytes)
        std::array<uint32_t, 3> maxComputeWorkGroupCount       = {}; // max grid dimensions
        uint32_t                maxComputeWorkGroupInvocations = 0; // max threads per CTA
        std::array<uint32_t, 3> maxComputeWorkGroupSize        = {}; // max CTA dimensions
    };

    struct PhysicalDeviceSparseProperties {};

    struct PhysicalDeviceProperties {
        uint32_t                       apiVersion    = 0; // CUDA runtime version, packed like a Vulkan word
        uint32_t                       driverVersion = 0; // CUDA driver version, packed likewise
        uint32_t                       vendorID      = 0x10DE;
        uint32_t                       deviceID      = 0; // CUDA ordinal: unique per GPU
        PhysicalDeviceType             deviceType    = PhysicalDeviceType::eDiscreteGpu;
        std::string                    deviceName;
        std::array<uint8_t, 16>        pipelineCacheUUID = {}; // CUDA device UUID
        PhysicalDeviceLimits           limits;
        PhysicalDeviceSparseProperties sparseProperties;
        // CUDA-only extras (additive)
        uint32_t smCount = 0, computeCapabilityMajor = 0, computeCapabilityMinor = 0;
    };

    enum MemoryHeapFlagBits : uint32_t { eHeapDeviceLocal = 1u, eHeapMultiInstance = 2u };
    enum MemoryPropertyFlagBits : uint32_t {
        eDeviceLocal = 1u, eHostVisible = 2u, eHostCoherent = 4u, eHostCached = 8u, eLazilyAllocated = 16u,
        eProtected = 32u
    };

    struct MemoryHeap {
        uint64_t size  = 0;
        uint32_t flags = 0;
    };
    struct MemoryType {
        uint32_t propertyFlags = 0;
        uint32_t heapIndex     = 0;
    };
    struct PhysicalDeviceMemoryProperties {
        uint32_t                   memoryTypeCount = 0;
        std::array<MemoryType, 32> memoryTypes     = {};
        uint32_t                   memoryHeapCount = 0;
        std::array<MemoryHeap, 16> memoryHeaps     = {};
    };

    struct PhysicalDeviceInfo {
        PhysicalDeviceProperties       deviceProperties;
        PhysicalDeviceMemoryProperties memoryProperties;
    };

    // What a ComputeDeviceInterface keeps alive after the ComputeContext is gone (the reference
    // shares its Vulkan instance the same way, device_interface.hpp:22).
    struct ComputeContextState {
        int      deviceCount    = 0;
        uint32_t runtimeVersion = 0; // packed
        uint32_t driverVersion  = 0; // packed
        void     log(const std::string& line) const;
    };

    // Stand-in for vk::raii::PhysicalDevice: the CUDA ordinal plus its properties.
    struct PhysicalDevice {
        int                ordinal = 0;
        PhysicalDeviceInfo info;
    };

    class ComputeContext {
        std::shared_ptr<ComputeContextState> state;

      public:
        explicit ComputeContext(std::shared_ptr<ComputeContextState> state_);
        virtual ~ComputeContext() = default;

        // nullptr when no CUDA device / driver is available.
        static std::shared_ptr<ComputeContext> create(uint32_t version = (1u << 12));

        std::string                             getVulkanAPIVersion();
        std::vector<PhysicalDeviceInfo>         getPhysicalDevicesInfo();
        std::shared_ptr<ComputeDeviceInterface> getDeviceInterface(uint32_t deviceId);
    };
} // namespace epseon::gpu::cpp
