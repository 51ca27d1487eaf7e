// HardwareConfig -- launch-shape hints of a task.
// Reference: cpp/gpu/include/epseon/gpu/task_configurator/hardware_config.hpp:12-76 (same three public
// u32 members, getters, equality and clone helpers).  In this build potential_buffer_size is checked
// against the potential's point count, group_size bounds the trial energies per sweep row, and
// allocation_block_size is carried untouched (the reference never reads it either).
#pragma once
#include "epseon/gpu/predecl.hpp"

#include <cstdint>
#include <memory>
#include <type_traits>

namespace epseon::gpu::cpp {

    template <typename FP>
    struct HardwareConfig : public std::enable_shared_from_this<HardwareConfig<FP>> {
        static_assert(std::is_floating_point_v<FP>, "FP must be an floating-point type.");

        uint32_t potential_buffer_size = 0;
        uint32_t group_size            = 0;
        uint32_t allocation_block_size = 0;

        HardwareConfig() = default;
        HardwareConfig(uint32_t potential_buffer_size_, uint32_t group_size_, uint32_t allocation_block_size_) :
            potential_buffer_size(potential_buffer_size_),
            group_size(group_size_),
            allocation_block_size(allocation_block_size_) {}
        HardwareConfig(const HardwareConfig& o) :
            std::enable_shared_from_this<HardwareConfig<FP>>(),
            potential_buffer_size(o.potential_buffer_size),
            group_size(o.group_size),
            allocation_block_size(o.allocation_block_size) {}
        HardwareConfig& operator=(const HardwareConfig& o) {
            potential_buffer_size = o.potential_buffer_size;
            group_size            = o.group_size;
            allocation_block_size = o.allocation_block_size;
            return *this;
        }
        virtual ~HardwareConfig() = default;

        bool operator==(const HardwareConfig<FP>& o) const {
            return potential_buffer_size == o.potential_buffer_size && group_size == o.group_size &&
                   allocation_block_size == o.allocation_block_size;
        }

        [[nodiscard]] virtual std::shared_ptr<HardwareConfig> shared_clone() const {
            return std::make_shared<HardwareConfig>(*this);
        }
        [[nodiscard]] virtual std::unique_ptr<HardwareConfig> unique_clone() const {
            return std::make_unique<HardwareConfig>(*this);
        }

        [[nodiscard]] uint32_t getPotentialBufferSize() const { return potential_buffer_size; }
        [[nodiscard]] uint32_t getGroupSize() const { return group_size; }
        [[nodiscard]] uint32_t getAllocationBlockSize() const { return allocation_block_size; }
    };
} // namespace epseon::gpu::cpp
