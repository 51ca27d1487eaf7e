// Potential sources: where the tabulated V(r) of every curve comes from.
// Reference: cpp/gpu/include/epseon/gpu/task_configurator/potential_source.hpp -- PotentialSource<FP>
// (:13-41), PotentialFileLoader<FP> (:44-100), MorsePotentialConfig<FP> (:103-183),
// MorsePotentialGenerator<FP> (:186-234).  Same names, constructors, equality and clone helpers.
// DIFFERENCE (the point of this build): get_potential_data() is REAL here.  In the reference both
// implementations `return {};` (:89-91, :223-225).
#pragma once
#include "epseon/gpu/predecl.hpp"

#include <cmath>
#include <cstdint>
#include <fstream>
#include <memory>
#include <span>
#include <sstream>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <vector>

namespace epseon::gpu::cpp {

    template <typename FP>
    class PotentialSource : public std::enable_shared_from_this<PotentialSource<FP>> {
        static_assert(std::is_floating_point_v<FP>, "FP must be an floating-point type.");

      public:
        PotentialSource() noexcept = default;
        PotentialSource(const PotentialSource&) noexcept : std::enable_shared_from_this<PotentialSource<FP>>() {}
        PotentialSource& operator=(const PotentialSource&) noexcept { return *this; }
        virtual ~PotentialSource() = default;

        virtual bool                         equals(const PotentialSource<FP>& other) const = 0;
        // One row of point_count values per curve, V(r_i) on r_i = min_r + i*h.
        virtual std::vector<std::vector<FP>> get_potential_data()                           = 0;
        // Grid spacing h of every curve (additive; needed to scale energies).
        virtual std::vector<double>          get_grid_steps() const                         = 0;
        [[nodiscard]] virtual std::shared_ptr<PotentialSource<FP>> shared_clone() const     = 0;
        [[nodiscard]] virtual std::unique_ptr<PotentialSource<FP>> unique_clone() const     = 0;
    };

    // Tabulated curves from text files: one "r V" pair per line ('#' comments allowed), r uniformly
    // spaced.  The reference stores the names and loads nothing.
    template <typename FP>
    class PotentialFileLoader : public PotentialSource<FP> {
        std::vector<std::string> file_names = {};

        static void read_table(const std::string& path, std::vector<double>& r, std::vector<FP>& v) {
            std::ifstream in(path);
            if (!in) throw std::runtime_error("PotentialFileLoader: cannot open '" + path + "'");
            std::string line;
            while (std::getline(in, line)) {
                const auto hash = line.find('#');
                if (hash != std::string::npos) line.erase(hash);
                std::istringstream ls(line);
                double             ri = 0, vi = 0;
                if (ls >> ri >> vi) {
                    r.push_back(ri);
                    v.push_back(static_cast<FP>(vi));
                }
            }
            if (r.size() < 3) throw std::runtime_error("PotentialFileLoader: '" + path + "' holds fewer than 3 points");
        }

      public:
        PotentialFileLoader() noexcept = default;
        PotentialFileLoader(const std::span<const std::string> names) : // NOLINT(hicpp-explicit-conversions)
            file_names(names.begin(), names.end()) {}
        ~PotentialFileLoader() override = default;

        bool equals(const PotentialSource<FP>& other) const override {
            const auto* o = dynamic_cast<const PotentialFileLoader<FP>*>(&other);
            return o != nullptr && file_names == o->file_names;
        }

        std::vector<std::vector<FP>> get_potential_data() override {
            std::vector<std::vector<FP>> out;
            for (const auto& name : file_names) {
                std::vector<double> r;
                std::vector<FP>     v;
                read_table(name, r, v);
                if (!out.empty() && out.front().size() != v.size())
                    throw std::runtime_error("PotentialFileLoader: all curves must have the same point count");
                out.push_back(std::move(v));
            }
            return out;
        }

        std::vector<double> get_grid_steps() const override {
            std::vector<double> out;
            for (const auto& name : file_names) {
                std::vector<double> r;
                std::vector<FP>     v;
                read_table(name, r, v);
                out.push_back((r.back() - r.front()) / static_cast<double>(r.size() - 1));
            }
            return out;
        }

        std::shared_ptr<PotentialSource<FP>> shared_clone() const override {
            return std::make_shared<PotentialFileLoader<FP>>(*this);
        }
        std::unique_ptr<PotentialSource<FP>> unique_clone() const override {
            return std::make_unique<PotentialFileLoader<FP>>(*this);
        }
    };

    template <typename FP>
    class MorsePotentialConfig {
        static_assert(std::is_floating_point_v<FP>, "FP must be an floating-point type.");

        FP       dissociation_energy       = {};
        FP       equilibrium_bond_distance = {};
        FP       well_width                = {};
        FP       min_r                     = {};
        FP       max_r                     = {};
        uint32_t point_count               = {};

      public:
        MorsePotentialConfig() = default;
        MorsePotentialConfig(FP dissociation_energy_, FP equilibrium_bond_distance_, FP well_width_, FP min_r_,
                             FP max_r_, uint32_t point_count_) :
            dissociation_energy(dissociation_energy_),
            equilibrium_bond_distance(equilibrium_bond_distance_),
            well_width(well_width_),
            min_r(min_r_),
            max_r(max_r_),
            point_count(point_count_) {}
        virtual ~MorsePotentialConfig() = default;

        bool operator==(const MorsePotentialConfig<FP>& o) const {
            return dissociation_energy == o.dissociation_energy &&
                   equilibrium_bond_distance == o.equilibrium_bond_distance && well_width == o.well_width &&
                   min_r == o.min_r && max_r == o.max_r && point_count == o.point_count;
        }

        FP                     getDissociationEnergy() const { return dissociation_energy; }
        FP                     getEquilibriumBondDistance() const { return equilibrium_bond_distance; }
        FP                     getWellWidth() const { return well_width; }
        FP                     getMinR() const { return min_r; }
        FP                     getMaxR() const { return max_r; }
        [[nodiscard]] uint32_t getPointCount() const { return point_count; }

        // Grid spacing h = (max_r - min_r)/(point_count - 1), in double.
        [[nodiscard]] double getGridStep() const {
            return (static_cast<double>(max_r) - static_cast<double>(min_r)) / static_cast<double>(point_count - 1);
        }

        // N1: V_i = De * (1 - exp(-a (r_i - re)))^2, a = well_width.  Evaluated in
        // double and cast to FP once, so the float32 instantiation sees the same curve.
        [[nodiscard]] std::vector<FP> tabulate() const {
            std::vector<FP> v(point_count);
            const double    h = getGridStep(), De = dissociation_energy, re = equilibrium_bond_distance,
                         a = well_width, r0 = min_r;
            for (uint32_t i = 0; i < point_count; i++) {
                const double r = r0 + static_cast<double>(i) * h;
                const double t = 1.0 - std::exp(-a * (r - re));
                v[i]           = static_cast<FP>((De * t) * t);
            }
            return v;
        }
    };

    template <typename FP>
    class MorsePotentialGenerator : public PotentialSource<FP> {
      public:
        std::vector<MorsePotentialConfig<FP>> configurations = {};

        MorsePotentialGenerator() = default;
        explicit MorsePotentialGenerator(std::vector<MorsePotentialConfig<FP>>&& configurations_) :
            configurations(std::move(configurations_)) {}
        ~MorsePotentialGenerator() override = default;

        bool equals(const PotentialSource<FP>& other) const override {
            const auto* o = dynamic_cast<const MorsePotentialGenerator<FP>*>(&other);
            return o != nullptr && configurations == o->configurations;
        }

        std::vector<std::vector<FP>> get_potential_data() override {
            std::vector<std::vector<FP>> out;
            out.reserve(configurations.size());
            for (const auto& cfg : configurations) {
                if (cfg.getPointCount() < 3) throw std::runtime_error("Morse potential needs at least 3 points");
                out.push_back(cfg.tabulate());
            }
            return out;
        }

        std::vector<double> get_grid_steps() const override {
            std::vector<double> out;
            for (const auto& cfg : configurations) out.push_back(cfg.getGridStep());
            return out;
        }

        std::shared_ptr<PotentialSource<FP>> shared_clone() const override {
            return std::make_shared<MorsePotentialGenerator<FP>>(*this);
        }
        std::unique_ptr<PotentialSource<FP>> unique_clone() const override {
            return std::make_unique<MorsePotentialGenerator<FP>>(*this);
        }
    };

    template <typename FP>
    bool operator==(const PotentialSource<FP>& lhs, const PotentialSource<FP>& rhs) {
        return lhs.equals(rhs);
    }
    template <typename FP>
    bool operator==(const PotentialFileLoader<FP>& lhs, const PotentialFileLoader<FP>& rhs) {
        return lhs.equals(rhs);
    }
    template <typename FP>
    bool operator==(const MorsePotentialGenerator<FP>& lhs, const MorsePotentialGenerator<FP>& rhs) {
        return lhs.equals(rhs);
    }
} // namespace epseon::gpu::cpp
