// Potential sources: where the tabulated V(r) of every curve comes from.
// Reference: cpp/gpu/include/epseon/gpu/task_configurator/potential_source.hpp -- PotentialSource<FP>
// (:13-41), PotentialFileLoader<FP> (:44-100), MorsePotentialConfig<FP> (:103-183),
// MorsePotentialGenerator<FP> (:186-234).  Same names, constructors, equality and clone helpers.
// DIFFERENCE (the point of this build): get_potential_data() is REAL here.  In the reference both
// implementations `return {};` (:89-91, :223-225).
#pragma once
#include "epseon/gpu/predecl.hpp"

#include "epseon_cuda.h"

#include <cmath>
#include <cstdio>
#include <cstdint>
#include <fstream>
#include <memory>
#include <mutex>
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
        // First grid point r_0 of every curve (additive; the centrifugal term needs r).
        virtual std::vector<double>          get_grid_origins() const                       = 0;
        [[nodiscard]] virtual std::shared_ptr<PotentialSource<FP>> shared_clone() const     = 0;
        [[nodiscard]] virtual std::unique_ptr<PotentialSource<FP>> unique_clone() const     = 0;
        // Why get_potential_data() came back empty, if it did (additive).
        [[nodiscard]] virtual std::string get_last_error() const { return {}; }
    };

    // Tabulated curves from files: text with one "r V" pair per line ('#' comments allowed), or a
    // NumPy .npy array of shape (n, 2) (float64, C order) when the name ends in ".npy".  A table on
    // a uniform r grid is used as it is; a table on a non-uniform grid (ab initio points), or any
    // table when `point_count` > 0, is resampled onto `point_count` uniform points of
    // [r_first, r_last] with a natural cubic spline (coefficients: eps_spline_coefficients of the
    // C ABI; evaluation fma(fma(fma(d,dx,c),dx,b),dx,a) -- DESIGN.md section 3.1).  The reference
    // stores the names and loads nothing.
    template <typename FP>
    class PotentialFileLoader : public PotentialSource<FP> {
        std::vector<std::string> file_names  = {};
        uint32_t                 point_count = 0; // 0: keep the file's own point count

        struct Table {
            std::vector<double> v;
            double              h  = 0.0;
            double              r0 = 0.0;
        };
        // Every file is read and resampled once per loader; copies share the result AND the lock, so
        // the same configuration submitted to two devices (two worker threads) loads once.
        struct Shared {
            std::mutex                                mutex;
            bool                                      tried = false;
            std::shared_ptr<const std::vector<Table>> tables;
            std::string                               error; // why loading failed
        };
        mutable std::shared_ptr<Shared> shared = std::make_shared<Shared>();

        Shared& state() const {
            if (!shared) shared = std::make_shared<Shared>(); // moved-from object
            return *shared;
        }

        // nullptr when the files cannot be loaded (reason in state().error)
        std::shared_ptr<const std::vector<Table>> try_tables() const {
            Shared&                     st = state();
            std::lock_guard<std::mutex> g(st.mutex);
            if (!st.tried) {
                st.tried = true;
                try {
                    auto all = std::make_shared<std::vector<Table>>();
                    all->reserve(file_names.size());
                    for (const auto& name : file_names) {
                        all->push_back(load(name));
                        if (all->front().v.size() != all->back().v.size())
                            throw std::runtime_error("PotentialFileLoader: all curves must have the same point count");
                    }
                    st.tables = std::move(all);
                } catch (const std::exception& e) {
                    st.error = e.what();
                }
            }
            return st.tables;
        }

        const std::vector<Table>& tables() const {
            const auto t = try_tables();
            if (!t) throw std::runtime_error(state().error);
            return *t;
        }

        // NumPy .npy (format 1.0 / 2.0 / 3.0): a C-ordered little-endian float64 array of shape (n, 2),
        // column 0 = r, column 1 = V.
        static void read_npy(const std::string& path, std::vector<double>& r, std::vector<double>& v) {
            std::ifstream in(path, std::ios::binary);
            if (!in) throw std::runtime_error("PotentialFileLoader: cannot open '" + path + "'");
            unsigned char head[10] = {};
            in.read(reinterpret_cast<char*>(head), 10);
            if (!in || std::string(reinterpret_cast<char*>(head), 6) != "\x93NUMPY")
                throw std::runtime_error("PotentialFileLoader: '" + path + "' is not a .npy file");
            size_t hlen = head[8] | (static_cast<size_t>(head[9]) << 8);
            if (head[6] >= 2) { // 4-byte header length
                unsigned char more[2] = {};
                in.read(reinterpret_cast<char*>(more), 2);
                hlen |= (static_cast<size_t>(more[0]) << 16) | (static_cast<size_t>(more[1]) << 24);
            }
            std::string header(hlen, '\0');
            in.read(header.data(), static_cast<std::streamsize>(hlen));
            if (!in) throw std::runtime_error("PotentialFileLoader: '" + path + "': truncated .npy header");
            const auto has = [&](const char* needle) { return header.find(needle) != std::string::npos; };
            if (!(has("'<f8'") || has("'=f8'") || has("'|f8'")) || !has("'fortran_order': False"))
                throw std::runtime_error("PotentialFileLoader: '" + path + "' must hold C-ordered little-endian float64");
            const auto lp = header.find('(', header.find("'shape'")), rp = header.find(')', lp);
            unsigned long long n = 0, cols = 0;
            if (lp == std::string::npos || rp == std::string::npos ||
                std::sscanf(header.substr(lp, rp - lp + 1).c_str(), "(%llu, %llu)", &n, &cols) != 2 || cols != 2)
                throw std::runtime_error("PotentialFileLoader: '" + path + "' must have shape (n, 2)");
            if (n < 3) throw std::runtime_error("PotentialFileLoader: '" + path + "' holds fewer than 3 points");
            std::vector<double> rv(2 * static_cast<size_t>(n));
            in.read(reinterpret_cast<char*>(rv.data()), static_cast<std::streamsize>(rv.size() * sizeof(double)));
            if (!in) throw std::runtime_error("PotentialFileLoader: '" + path + "': truncated .npy data");
            r.resize(n);
            v.resize(n);
            for (size_t i = 0; i < n; i++) {
                r[i] = rv[2 * i];
                v[i] = rv[2 * i + 1];
            }
        }

        static void read_table(const std::string& path, std::vector<double>& r, std::vector<double>& v) {
            if (path.size() >= 4 && path.compare(path.size() - 4, 4, ".npy") == 0) return read_npy(path, r, v);
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
                    v.push_back(vi);
                }
            }
            if (r.size() < 3) throw std::runtime_error("PotentialFileLoader: '" + path + "' holds fewer than 3 points");
        }

        static bool is_uniform(const std::vector<double>& r) {
            const double h = (r.back() - r.front()) / static_cast<double>(r.size() - 1);
            for (size_t i = 0; i < r.size(); i++)
                if (std::fabs(r[i] - (r.front() + static_cast<double>(i) * h)) > 1e-9 * std::fabs(h)) return false;
            return true;
        }

        [[nodiscard]] Table load(const std::string& path) const {
            std::vector<double> r, v;
            read_table(path, r, v);
            Table          t;
            const uint32_t n = point_count > 0 ? point_count : static_cast<uint32_t>(r.size());
            t.h              = (r.back() - r.front()) / static_cast<double>(n - 1);
            t.r0             = r.front();
            if (point_count == 0 && is_uniform(r)) {
                t.v = std::move(v);
                return t;
            }
            const uint32_t      K = static_cast<uint32_t>(r.size());
            std::vector<double> coef(4 * static_cast<size_t>(K - 1));
            if (eps_spline_coefficients(r.data(), v.data(), K, coef.data()) != EPS_OK)
                throw std::runtime_error("PotentialFileLoader: '" + path + "': " + eps_last_error(nullptr));
            t.v.resize(n);
            for (uint32_t i = 0; i < n; i++) {
                const double x  = r.front() + static_cast<double>(i) * t.h;
                uint32_t     lo = 0, hi = K - 1;
                while (hi - lo > 1) {
                    const uint32_t mid = (lo + hi) / 2;
                    if (r[mid] <= x) lo = mid;
                    else hi = mid;
                }
                const double  dx = x - r[lo];
                const double* c  = coef.data() + 4 * static_cast<size_t>(lo);
                t.v[i]           = std::fma(std::fma(std::fma(c[3], dx, c[2]), dx, c[1]), dx, c[0]);
            }
            return t;
        }

      public:
        PotentialFileLoader() noexcept = default;
        PotentialFileLoader(const std::span<const std::string> names, uint32_t point_count_ = 0) : // NOLINT(hicpp-explicit-conversions)
            file_names(names.begin(), names.end()), point_count(point_count_) {
            if (point_count_ == 1 || point_count_ == 2) throw std::runtime_error("PotentialFileLoader: point_count must be 0 or >= 3");
        }
        ~PotentialFileLoader() override = default;

        bool equals(const PotentialSource<FP>& other) const override {
            const auto* o = dynamic_cast<const PotentialFileLoader<FP>*>(&other);
            return o != nullptr && file_names == o->file_names && point_count == o->point_count;
        }

        // Like the reference's (potential_source.hpp:89-91) this never throws: files that cannot be
        // read give an empty result -- which is all the reference ever returns -- and the reason is
        // kept for get_last_error(); a task over such a source fails with that reason in its status.
        // load() is the strict variant.
        std::vector<std::vector<FP>> get_potential_data() override {
            std::vector<std::vector<FP>> out;
            const auto                   t = try_tables();
            if (!t) return out;
            for (const Table& tb : *t) out.emplace_back(tb.v.begin(), tb.v.end());
            return out;
        }

        // Read (once) and validate every file; throws std::runtime_error with the reason.
        void load() const { (void)tables(); }

        [[nodiscard]] std::string get_last_error() const override {
            Shared&                     st = state();
            std::lock_guard<std::mutex> g(st.mutex);
            return st.error;
        }

        std::vector<double> get_grid_steps() const override {
            std::vector<double> out;
            for (const Table& t : tables()) out.push_back(t.h);
            return out;
        }

        std::vector<double> get_grid_origins() const override {
            std::vector<double> out;
            for (const Table& t : tables()) out.push_back(t.r0);
            return out;
        }

        std::shared_ptr<PotentialSource<FP>> shared_clone() const override {
            return std::make_shared<PotentialFileLoader<FP>>(*this);
        }
        std::unique_ptr<PotentialSource<FP>> unique_clone() const override {
            return std::make_unique<PotentialFileLoader<FP>>(*this);
        }
    };

    // Curves handed over in memory (additive; the natural source for ab initio data already held in
    // an array): n_curves rows of point_count values on the uniform grid r_i = min_r + i*h,
    // h = (max_r - min_r)/(point_count - 1).  Values are kept in double and cast to FP on request, so
    // the float64 task sees the caller's bits.
    template <typename FP>
    class TabulatedPotentialSource : public PotentialSource<FP> {
        std::vector<std::vector<double>> tables = {};
        double                           min_r  = 0.0;
        double                           max_r  = 0.0;

      public:
        TabulatedPotentialSource() noexcept = default;
        TabulatedPotentialSource(std::vector<std::vector<double>> tables_, double min_r_, double max_r_) :
            tables(std::move(tables_)), min_r(min_r_), max_r(max_r_) {
            if (!(max_r > min_r)) throw std::runtime_error("TabulatedPotentialSource: max_r must exceed min_r");
            for (const auto& t : tables) {
                if (t.size() < 3) throw std::runtime_error("TabulatedPotentialSource: a curve needs at least 3 points");
                if (t.size() != tables.front().size())
                    throw std::runtime_error("TabulatedPotentialSource: all curves must have the same point count");
            }
        }
        ~TabulatedPotentialSource() override = default;

        bool equals(const PotentialSource<FP>& other) const override {
            const auto* o = dynamic_cast<const TabulatedPotentialSource<FP>*>(&other);
            return o != nullptr && tables == o->tables && min_r == o->min_r && max_r == o->max_r;
        }

        std::vector<std::vector<FP>> get_potential_data() override {
            std::vector<std::vector<FP>> out;
            out.reserve(tables.size());
            for (const auto& t : tables) out.emplace_back(t.begin(), t.end());
            return out;
        }

        std::vector<double> get_grid_steps() const override {
            if (tables.empty()) return {};
            return std::vector<double>(tables.size(), (max_r - min_r) / static_cast<double>(tables.front().size() - 1));
        }

        std::vector<double> get_grid_origins() const override { return std::vector<double>(tables.size(), min_r); }

        std::shared_ptr<PotentialSource<FP>> shared_clone() const override {
            return std::make_shared<TabulatedPotentialSource<FP>>(*this);
        }
        std::unique_ptr<PotentialSource<FP>> unique_clone() const override {
            return std::make_unique<TabulatedPotentialSource<FP>>(*this);
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

        std::vector<double> get_grid_origins() const override {
            std::vector<double> out;
            for (const auto& cfg : configurations) out.push_back(static_cast<double>(cfg.getMinR()));
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
