// _libepseon_cpu -- the reference's CPU module is a hello-world stub exporting greet()
// (cpp/cpu/source/libcpu.cpp:8-19, python/test/test_device/test_cpu/test_libepseon_cpu.py:5-9).
// Kept as such: no numerics live here (the parity oracle is test infrastructure under oracle/, and the
// product has no CPU fallback).
#include <pybind11/pybind11.h>

#include <string>

namespace {
    std::string greet() { return "Hello, World from C++!"; }
} // namespace

PYBIND11_MODULE(_libepseon_cpu, m) {
    m.doc() = "CPU sub package placeholder (reference parity: greet only).";
    m.def("greet", &greet, "Temporary example function.");
}
