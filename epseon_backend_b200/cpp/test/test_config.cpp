// Host-only conformance tests of the configuration classes (no GPU needed).
// They restate what the reference's typed gtests assert for <float, double>
// (cpp/gpu/test/task_configurator/test_hardware_config.cpp, test_potential_source.cpp,
// test_algorithm_confgu.cpp, test_task_configurator.cpp): construction, copy / move, clone,
// equality, getImplementation() != nullptr, isConfigured.  One deliberate difference:
// get_potential_data() is no longer empty (the reference asserts `.empty()`,
// test_potential_source.cpp:24 -- that is the stub this build replaces).
#include "epseon/gpu/libgpu.hpp"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <regex>
#include <string>

using namespace epseon::gpu::cpp;

static int g_checks = 0;
#define CHECK(cond)                                                                  \
    do {                                                                             \
        ++g_checks;                                                                  \
        if (!(cond)) {                                                               \
            std::fprintf(stderr, "%s:%d: CHECK failed: %s\n", __FILE__, __LINE__, #cond); \
            std::exit(1);                                                            \
        }                                                                            \
    } while (0)

template <typename FP>
void hardware_config() {
    HardwareConfig<FP> a(16500, 512, 16 * 1024 * 1024), dflt;
    CHECK(a.getPotentialBufferSize() == 16500 && a.getGroupSize() == 512 && a.getAllocationBlockSize() == 16777216);
    CHECK(dflt.potential_buffer_size == 0 && dflt.group_size == 0 && dflt.allocation_block_size == 0);
    HardwareConfig<FP> b(a);
    CHECK(a == b);
    HardwareConfig<FP> c(std::move(b));
    CHECK(a == c);
    c.group_size = 7;
    CHECK(!(a == c));
    auto s = a.shared_clone();
    auto u = a.unique_clone();
    CHECK(*s == a && *u == a && s.get() != &a);
}

template <typename FP>
void potential_source() {
    MorsePotentialConfig<FP> m(5500.0, 0.6, 10, 0.0, 10.0, 500), m2(5500.0, 0.6, 10, 0.0, 10.0, 501), d;
    CHECK(m.getDissociationEnergy() == FP(5500.0) && m.getEquilibriumBondDistance() == FP(0.6));
    CHECK(m.getWellWidth() == FP(10) && m.getMinR() == FP(0) && m.getMaxR() == FP(10) && m.getPointCount() == 500);
    CHECK(d.getPointCount() == 0);
    CHECK(m == MorsePotentialConfig<FP>(m) && !(m == m2));

    MorsePotentialGenerator<FP> g(std::vector<MorsePotentialConfig<FP>>{m, m}), empty;
    CHECK(empty.get_potential_data().empty());
    auto data = g.get_potential_data();
    CHECK(data.size() == 2 && data[0].size() == 500);
    // Morse shape: huge at r=0, zero-ish at re (index 30 = 0.6/(10/499)), ~De at r_max
    CHECK(data[0][0] > FP(1e8) && std::fabs(double(data[0][499]) - 5500.0) < 1e-3);
    FP mn = data[0][0];
    for (FP v : data[0]) mn = v < mn ? v : mn;
    CHECK(mn >= FP(0) && mn < FP(30));
    CHECK(g.get_grid_steps().size() == 2 && std::fabs(g.get_grid_steps()[0] - 10.0 / 499) < 1e-15);
    MorsePotentialGenerator<FP> g2(g);
    CHECK(g == g2 && g.equals(g2) && !(g == empty));
    auto sc = g.shared_clone();
    auto uc = g.unique_clone();
    CHECK(sc->equals(g) && uc->equals(g));

    const std::vector<std::string> names{"a.txt", "b.txt"};
    PotentialFileLoader<FP>        f(names), f2(names), f3;
    CHECK(f == f2 && !(f == f3) && !f.equals(g) && !g.equals(f));
    CHECK(f.shared_clone()->equals(f) && f.unique_clone()->equals(f));
    CHECK(f3.get_potential_data().empty());
    {   // curves handed over in memory
        TabulatedPotentialSource<FP> t({{4.0, 1.0, 0.0, 1.0, 4.0}, {8.0, 2.0, 0.0, 2.0, 8.0}}, 1.0, 3.0), t0;
        auto                         d = t.get_potential_data();
        CHECK(d.size() == 2 && d[1].size() == 5 && d[1][0] == FP(8) && d[0][2] == FP(0));
        CHECK(t.get_grid_steps() == (std::vector<double>{0.5, 0.5}) && t.get_grid_origins() == (std::vector<double>{1.0, 1.0}));
        CHECK(t.equals(*t.shared_clone()) && t.equals(*t.unique_clone()) && !t.equals(t0) && !t.equals(g));
        bool refused = false;
        try {
            TabulatedPotentialSource<FP> bad({{1.0, 2.0, 3.0}, {1.0, 2.0}}, 0.0, 1.0);
        } catch (const std::runtime_error&) {
            refused = true;
        }
        CHECK(refused);
    }
    {   // a real table round-trips
        const char* path = "/tmp/epseon_b200_test_curve.txt";
        {
            std::ofstream out(path);
            out << "# r V\n";
            for (int i = 0; i < 11; i++) out << 1.0 + 0.5 * i << " " << (i - 5) * (i - 5) * 2.0 << "\n";
        }
        const std::vector<std::string> one{path};
        PotentialFileLoader<FP>        fl(one);
        auto                           t = fl.get_potential_data();
        CHECK(t.size() == 1 && t[0].size() == 11 && t[0][5] == FP(0) && t[0][0] == FP(50));
        CHECK(std::fabs(fl.get_grid_steps()[0] - 0.5) < 1e-15);
        CHECK(fl.get_grid_origins()[0] == 1.0);
    }
    {   // the same table as a NumPy .npy array of shape (11, 2) (format 1.0, what numpy.save writes)
        const char* path = "/tmp/epseon_b200_test_curve.npy";
        {
            std::string dict = "{'descr': '<f8', 'fortran_order': False, 'shape': (11, 2), }";
            while ((10 + dict.size() + 1) % 64 != 0) dict.push_back(' ');
            dict.push_back('\n');
            std::ofstream out(path, std::ios::binary);
            const unsigned char head[10] = {0x93, 'N', 'U', 'M', 'P', 'Y', 1, 0, static_cast<unsigned char>(dict.size() & 0xff),
                                            static_cast<unsigned char>(dict.size() >> 8)};
            out.write(reinterpret_cast<const char*>(head), 10);
            out.write(dict.data(), static_cast<std::streamsize>(dict.size()));
            for (int i = 0; i < 11; i++) {
                const double rv[2] = {1.0 + 0.5 * i, (i - 5) * (i - 5) * 2.0};
                out.write(reinterpret_cast<const char*>(rv), sizeof(rv));
            }
        }
        const std::vector<std::string> one{path};
        PotentialFileLoader<FP>        fl(one);
        auto                           t = fl.get_potential_data();
        CHECK(t.size() == 1 && t[0].size() == 11 && t[0][5] == FP(0) && t[0][0] == FP(50) && t[0][10] == FP(50));
        CHECK(std::fabs(fl.get_grid_steps()[0] - 0.5) < 1e-15);
        bool refused = false;
        try {
            const std::vector<std::string> bad{"/tmp/epseon_b200_test_curve.txt.npy"};
            PotentialFileLoader<FP> missing(bad);
            CHECK(missing.get_potential_data().empty() && !missing.get_last_error().empty());
            missing.load();
        } catch (const std::runtime_error&) {
            refused = true;
        }
        CHECK(refused);
    }
}

template <typename FP>
void algorithm_config() {
    VibwaAlgorithmConfig<FP> a(87.62, 87.62, 0.1, 0.1, 0, 0), b(87.62, 87.62, 0.1, 0.1, 0, 3), d;
    CHECK(a.getMassAtom0() == FP(87.62) && a.getMassAtom1() == FP(87.62) && a.getIntegrationStep() == FP(0.1));
    CHECK(a.getMinDistanceToAsymptote() == FP(0.1) && a.getMinLevel() == 0 && a.getMaxLevel() == 0);
    CHECK(d.getMassAtom0() == FP(0) && b.getLevelCount() == 4);
    CHECK(a == VibwaAlgorithmConfig<FP>(a) && !(a == b) && a.equals(*a.shared_clone()) && a.equals(*a.unique_clone()));
    CHECK(a.getImplementation() != nullptr);
    const AlgorithmConfig<FP>& base = a;
    CHECK(base == *a.shared_clone());
}

template <typename FP>
void task_configurator() {
    TaskConfigurator<FP> cfg;
    CHECK(!cfg.isConfigured());
    auto hw = std::make_shared<HardwareConfig<FP>>(500, 100, 16 * 1024 * 1024);
    auto ac = std::make_shared<VibwaAlgorithmConfig<FP>>(87.62, 87.62, 0.1, 0.1, 1, 3);
    auto ps = std::make_shared<MorsePotentialGenerator<FP>>(
        std::vector<MorsePotentialConfig<FP>>{MorsePotentialConfig<FP>(5500.0, 0.6, 10, 0.0, 10.0, 500)});
    cfg.setHardwareConfig(hw).setAlgorithmConfig(ac);
    CHECK(!cfg.isConfigured());
    cfg.setPotentialSource(ps);
    CHECK(cfg.isConfigured());
    // setters deep-clone
    CHECK(cfg.getHardwareConfig().get() != hw.get() && *cfg.getHardwareConfig() == *hw);
    CHECK(cfg.getAlgorithmConfig().get() != ac.get() && cfg.getAlgorithmConfig()->equals(*ac));
    CHECK(cfg.getPotentialSource().get() != ps.get() && cfg.getPotentialSource()->equals(*ps));
    TaskConfigurator<FP> copy(cfg);
    CHECK(copy.isConfigured() && copy.getHardwareConfig().get() != cfg.getHardwareConfig().get());
    TaskConfigurator<FP> moved(std::move(copy));
    CHECK(moved.isConfigured());
    const auto req = cfg.getShaderBufferRequirements();
    CHECK(req.size() == 100 && req[0].stagingBuffersCount == 1 && req[0].gpuOnlyStorageBuffersCount == 5);
    CHECK(req[0].outputBuffersElementCount == 3 && req[0].stagingBuffersElementCount == 500);
    CHECK(req[0].getStagingBuffersSizeBytes() == 500 * sizeof(FP));
    CHECK(req[0].getGpuOnlyStorageBufferSizeBytes() == 5 * 500 * sizeof(FP));
    // additive: rotational states (default {0}; copied and moved with the configurator; validated)
    CHECK(cfg.getRotationalStates() == std::vector<uint32_t>{0});
    cfg.setRotationalStates({0, 1, 7});
    TaskConfigurator<FP> rot(cfg);
    CHECK((rot.getRotationalStates() == std::vector<uint32_t>{0, 1, 7}));
    TaskConfigurator<FP> rot_moved(std::move(rot));
    CHECK(rot_moved.getRotationalStates().size() == 3);
    bool refused = false;
    try {
        cfg.setRotationalStates({});
    } catch (const std::runtime_error&) {
        refused = true;
    }
    CHECK(refused && cfg.getRotationalStates().size() == 3);
    CHECK(cfg.getPotentialSource()->get_grid_origins() == std::vector<double>{0.0});
}

int main() {
    hardware_config<float>();
    hardware_config<double>();
    potential_source<float>();
    potential_source<double>();
    algorithm_config<float>();
    algorithm_config<double>();
    task_configurator<float>();
    task_configurator<double>();

    CHECK(toPrecisionType("FLOAT32") == PrecisionType::Float32 && toPrecisionType("float64") == PrecisionType::Float64);
    CHECK(toString(PrecisionType::Float32) == "Float32" && getPrecisionType<double>() == PrecisionType::Float64);
    bool threw = false;
    try {
        toPrecisionType("Float80");
    } catch (const InvalidPrecisionTypeString& e) {
        threw = std::string(e.what()) == "Invalid PrecisionType literal in string: \"float80\"";
    }
    CHECK(threw);
    CHECK(epseon::gpu::common::vulkan_version_to_string((1u << 22) | (3u << 12) | 7u) == "0.1.3.7");
    CHECK(std::regex_match(epseon::gpu::common::vulkan_version_to_string(epseon::gpu::common::cuda_version_to_word(12090)),
                           std::regex("\\d+\\.\\d+\\.\\d+\\.\\d+")));
    std::printf("OK %d checks\n", g_checks);
    return 0;
}
