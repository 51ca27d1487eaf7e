// Device integration test (needs a GPU).  Restates the reference's cpp/gpu/test/test_libgpu.cpp:18-107
// and test_compute_context.cpp:18-21 against the CUDA build, and additionally checks the results that
// the reference cannot produce.
#include "epseon/gpu/libgpu.hpp"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <regex>

using namespace epseon::gpu::cpp;

#define CHECK(cond)                                                                  \
    do {                                                                             \
        if (!(cond)) {                                                               \
            std::fprintf(stderr, "%s:%d: CHECK failed: %s\n", __FILE__, __LINE__, #cond); \
            std::exit(1);                                                            \
        }                                                                            \
    } while (0)

template <typename FP>
std::shared_ptr<TaskHandle<FP>> prepare(uint32_t max_level) {
    auto ctx = ComputeContext::create();
    CHECK(ctx != nullptr);
    auto infos = ctx->getPhysicalDevicesInfo();
    CHECK(!infos.empty());
    auto device = ctx->getDeviceInterface(infos[0].deviceProperties.deviceID);
    CHECK(&device->getPhysicalDevice() != nullptr);
    auto cfg = device->template getTaskConfigurator<FP>();
    cfg->setHardwareConfig(std::make_shared<HardwareConfig<FP>>(500, 100, 16 * 1024 * 1024))
        .setAlgorithmConfig(std::make_shared<VibwaAlgorithmConfig<FP>>(87.62, 87.62, 0.1, 0.1, 0, max_level))
        .setPotentialSource(std::make_shared<MorsePotentialGenerator<FP>>(
            std::vector<MorsePotentialConfig<FP>>{MorsePotentialConfig<FP>(5500.0, 0.6, 10, 0.0, 10.0, 500)}));
    CHECK(cfg->isConfigured());
    return device->submitTask(cfg); // ctx and device go out of scope here, as in the reference test
}

int main() {
    {
        auto ctx = ComputeContext::create();
        CHECK(ctx != nullptr);
        CHECK(std::regex_match(ctx->getVulkanAPIVersion(), std::regex("\\d+\\.\\d+\\.\\d+\\.\\d+")));
        bool threw = false;
        try {
            ctx->getDeviceInterface(12345);
        } catch (const std::runtime_error& e) {
            threw = std::string(e.what()) == "Device not available.";
        }
        CHECK(threw);
        auto device = ctx->getDeviceInterface(0);
        auto empty  = device->getTaskConfigurator<double>();
        threw       = false;
        try {
            device->submitTask(empty);
        } catch (const std::runtime_error&) {
            threw = true;
        }
        CHECK(threw);
    }
    {
        auto handle = prepare<float>(0);
        CHECK(!handle->isDone() && !handle->isRunning());
        handle->startWorker();
        handle->wait();
        CHECK(handle->isDone());
        CHECK(!handle->hasFailed());
    }
    {
        auto handle = prepare<double>(11);
        handle->startWorker();
        bool threw = false;
        try {
            if (handle->isRunning()) handle->startWorker();
            else threw = true;
        } catch (const std::runtime_error&) {
            threw = true;
        }
        CHECK(threw);
        handle->wait();
        CHECK(handle->isDone() && !handle->hasFailed());
        CHECK(handle->getStatusMessage() == "done");
        const auto levels = handle->getLevels();
        const auto counts = handle->getLevelCounts();
        CHECK(levels.size() == 1 && levels[0].size() == 12 && counts.size() == 1);
        // analytic Morse: B = hbar^2/(2 mu), E_v = 2a sqrt(De B)(v+1/2) - a^2 B (v+1/2)^2; 12 bound levels
        const double B = 16.857629206 / (87.62 / 2.0), a = 10.0, De = 5500.0;
        // h = 0.02 resolves the a = 10 well with ~5 points: the discrete problem holds one spurious
        // level below De - 0.1 (the CPU oracle gives 13 on the same table), so only bound the count
        CHECK(counts[0] >= 12 && counts[0] <= 13);
        for (int v = 0; v < 12; v++) {
            const double exact = 2 * a * std::sqrt(De * B) * (v + 0.5) - a * a * B * (v + 0.5) * (v + 0.5);
            // N = 500 points over [0,10] is a coarse grid (h = 0.02): O(h^4) error ~1e-2 relative at the top
            CHECK(std::fabs(levels[0][v] - exact) / exact < 3e-2);
            if (v > 0) CHECK(levels[0][v] > levels[0][v - 1]);
        }
        std::printf("E0 = %.6f cm^-1 (analytic %.6f), solve %.3f ms\n", levels[0][0],
                    2 * a * std::sqrt(De * B) * 0.5 - a * a * B * 0.25, handle->getDeviceMilliseconds());
    }
    std::printf("OK\n");
    return 0;
}
