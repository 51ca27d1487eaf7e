// Reference behaviour: cpp/gpu/source/epseon/gpu/enums.cpp:14 (message), :34-53 (case-insensitive
// parse; the message carries the lower-cased literal).
#include "epseon/gpu/enums.hpp"

#include <algorithm>
#include <cctype>
#include <stdexcept>

namespace epseon::gpu::cpp {

    InvalidPrecisionTypeString::InvalidPrecisionTypeString(std::string_view literal) :
        message("Invalid PrecisionType literal in string: \"" + std::string(literal) + "\"") {}

    const char* InvalidPrecisionTypeString::what() const noexcept { return message.c_str(); }

    std::string toString(PrecisionType prec) {
        PrecisionTypeAssertValueCount(2);
        if (prec == PrecisionType::Float32) return "Float32";
        if (prec == PrecisionType::Float64) return "Float64";
        throw std::runtime_error("Unreachable");
    }

    PrecisionType toPrecisionType(std::string_view precision) {
        std::string lower(precision);
        std::transform(lower.begin(), lower.end(), lower.begin(), [](unsigned char ch) { return std::tolower(ch); });
        PrecisionTypeAssertValueCount(2);
        if (lower == "float32") return PrecisionType::Float32;
        if (lower == "float64") return PrecisionType::Float64;
        throw InvalidPrecisionTypeString(lower);
    }

    template <> PrecisionType getPrecisionType<float>() { return PrecisionType::Float32; }
    template <> PrecisionType getPrecisionType<double>() { return PrecisionType::Float64; }
} // namespace epseon::gpu::cpp
