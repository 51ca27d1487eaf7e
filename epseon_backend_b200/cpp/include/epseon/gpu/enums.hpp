// PrecisionType and its string conversions (reference: cpp/gpu/include/epseon/gpu/enums.hpp,
// cpp/gpu/source/epseon/gpu/enums.cpp:14,34-53 -- the error text is asserted by the reference's
// pytest, python/test/test_device/test_gpu/test_libepseon_gpu.py:219-223).
#pragma once
#include "epseon/libepseon.hpp"

#include <exception>
#include <string>
#include <string_view>

#define PrecisionTypeAssertValueCount(count)                                                       \
    static_assert(static_cast<int>(epseon::gpu::cpp::PrecisionType::_Last) == (count),            \
                  "The number of PrecisionTypes has changed.");

namespace epseon::gpu::cpp {

    enum class PrecisionType { Float32, Float64, _Last };

    class InvalidPrecisionTypeString : public std::exception {
        std::string message;

      public:
        explicit InvalidPrecisionTypeString(std::string_view literal);
        const char* what() const noexcept override;
    };

    std::string   toString(PrecisionType);
    PrecisionType toPrecisionType(std::string_view precision); // case-insensitive

    template <typename FP> PrecisionType getPrecisionType();
    template <> PrecisionType            getPrecisionType<float>();
    template <> PrecisionType            getPrecisionType<double>();
} // namespace epseon::gpu::cpp
