#include "epseon/gpu/common.hpp"

#include <string>

namespace epseon::gpu::common {
    std::string vulkan_version_to_string(uint32_t v) {
        return std::to_string(v >> 29) + "." + std::to_string((v >> 22) & 0x7Fu) + "." +
               std::to_string((v >> 12) & 0x3FFu) + "." + std::to_string(v & 0xFFFu);
    }
    uint32_t cuda_version_to_word(int cuda_version) {
        const uint32_t major = static_cast<uint32_t>(cuda_version / 1000);
        const uint32_t minor = static_cast<uint32_t>((cuda_version % 1000) / 10);
        return (major << 22) | (minor << 12);
    }
} // namespace epseon::gpu::common
