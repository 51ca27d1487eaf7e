// Version-word helpers (reference: cpp/gpu/include/epseon/gpu/common.hpp:9).
#pragma once
#include <cstdint>
#include <string>

namespace epseon::gpu::common {
    // Decodes a Vulkan-style packed version word (variant:3 | major:7 | minor:10 | patch:12) to
    // "variant.major.minor.patch".  Name kept for source compatibility; no Vulkan involved.
    std::string vulkan_version_to_string(uint32_t version);
    // Packs a CUDA version (1000*major + 10*minor) into the same word layout.
    uint32_t cuda_version_to_word(int cuda_version);
} // namespace epseon::gpu::common
