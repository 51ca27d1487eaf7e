// TaskHandle<FP> -- one submitted task: worker jthread, completion flags, and (additive) results.
// Reference: cpp/gpu/include/epseon/gpu/task_handle.hpp:18-162 -- same thread/flag protocol:
// startWorker() throws std::runtime_error when a worker is already running (:95-97), flags are
// release-stored / acquire-loaded (:69-87,:117-129), wait() joins only while running (:146-152),
// cancel() requests a cooperative stop (:136-144).  Additions: result storage filled by
// the algorithm, a real status message, and the worker never lets an exception escape (the
// reference would std::terminate, vibwa.hpp:352-354).
#pragma once
#include "epseon/gpu/predecl.hpp"

#include "epseon/gpu/device_interface.hpp"
#include "epseon/gpu/task_configurator/task_configurator.hpp"

#include <atomic>
#include <cstdint>
#include <exception>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <stop_token>
#include <string>
#include <thread>
#include <vector>

namespace epseon::gpu::cpp {

    template <typename FP>
    class TaskHandle : public std::enable_shared_from_this<TaskHandle<FP>> {
        std::shared_ptr<ComputeDeviceInterface> device            = {};
        std::shared_ptr<TaskConfigurator<FP>>   config            = {};
        std::atomic<bool>                       is_worker_done    = false;
        std::atomic<bool>                       is_worker_started = false;

        // results: written by the worker before the release-store of the done flag
        mutable std::mutex           result_mutex;
        std::string                  status = "created";
        bool                         failed = false;
        bool                         cancelled = false;
        std::vector<std::vector<FP>> levels;       // [curve][level - min_level], NaN = not found
        std::vector<uint32_t>        level_counts; // [curve] levels below the search ceiling
        double                       device_ms = 0.0;
        uint32_t                     search_n_coarse = 0, search_refine_points = 0, search_max_rounds = 0;
        double                       search_rel_tol = 0.0; // parameters eps_solve_levels ran with
        std::vector<double>          wavefunctions;           // [curve][level][point], flat; empty unless requested
        uint32_t                     wf_curves = 0, wf_levels = 0, wf_points = 0;

        std::jthread worker = {}; // last member: joins before the rest is destroyed

      public:
        TaskHandle() = default;
        TaskHandle(std::shared_ptr<ComputeDeviceInterface> device_, std::shared_ptr<TaskConfigurator<FP>> config_) :
            device(std::move(device_)), config(std::move(config_)) {}
        TaskHandle(const TaskHandle&)            = delete;
        TaskHandle& operator=(const TaskHandle&) = delete;
        ~TaskHandle()                            = default;

      protected:
        void setDoneFlag() { is_worker_done.store(true, std::memory_order_release); }
        void setStartedFlag() { is_worker_started.store(true, std::memory_order_release); }
        void setNotDoneFlag() { is_worker_done.store(false, std::memory_order_release); }
        void setNotStartedFlag() { is_worker_started.store(false, std::memory_order_release); }

        friend VibwaAlgorithm<FP>;

      public:
        void startWorker() {
            if (isRunning()) throw std::runtime_error("One worker is already running, can't start another one.");
            if (worker.joinable()) worker.join();
            setNotDoneFlag();
            setStartedFlag();
            worker = std::jthread(TaskHandle<FP>::run, this);
        }

        static void run(std::stop_token stop_token, TaskHandle<FP>* self) {
            try {
                self->setStatus("running");
                const auto implementation = self->config->getAlgorithmConfig()->getImplementation();
                implementation->run(stop_token, self);
            } catch (const std::exception& e) {
                self->setFailure(std::string("failed: ") + e.what());
            } catch (...) {
                self->setFailure("failed: unknown exception");
            }
            self->setDoneFlag();
            self->setNotStartedFlag();
        }

        [[nodiscard]] bool isDone() const { return is_worker_done.load(std::memory_order_acquire); }
        [[nodiscard]] bool isStarted() const { return is_worker_started.load(std::memory_order_acquire); }
        [[nodiscard]] bool isRunning() const { return isStarted() && !isDone(); }

        bool cancel() { return isRunning() ? worker.request_stop() : false; }

        void wait() {
            if (worker.joinable()) worker.join();
        }

        const TaskConfigurator<FP>&                 getTaskConfigurator() const { return *config; }
        [[nodiscard]] const ComputeDeviceInterface& getDeviceInterface() const { return *device; }

        // ---- additive result surface ----
        void setStatus(const std::string& s) {
            std::lock_guard<std::mutex> g(result_mutex);
            if (!failed) status = s;
        }
        // The task stopped on its stop_token (cancel()): is_done() becomes true, has_failed() stays
        // false, the status reads "cancelled" and was_cancelled() is true; results may be missing.
        void setCancelled() {
            std::lock_guard<std::mutex> g(result_mutex);
            if (!failed) status = "cancelled";
            cancelled = true;
        }
        [[nodiscard]] bool wasCancelled() const {
            std::lock_guard<std::mutex> g(result_mutex);
            return cancelled;
        }
        void setFailure(const std::string& s) {
            std::lock_guard<std::mutex> g(result_mutex);
            status = s;
            failed = true;
        }
        void setResults(std::vector<std::vector<FP>> levels_, std::vector<uint32_t> counts_, double ms) {
            std::lock_guard<std::mutex> g(result_mutex);
            levels       = std::move(levels_);
            level_counts = std::move(counts_);
            device_ms    = ms;
        }
        void setSearchParameters(uint32_t n_coarse, uint32_t refine_points, uint32_t max_rounds, double rel_tol) {
            std::lock_guard<std::mutex> g(result_mutex);
            search_n_coarse      = n_coarse;
            search_refine_points = refine_points;
            search_max_rounds    = max_rounds;
            search_rel_tol       = rel_tol;
        }
        // {n_coarse, refine_points, max_rounds} and rel_tol of the level search the task ran.
        void getSearchParameters(uint32_t out[3], double& rel_tol) const {
            std::lock_guard<std::mutex> g(result_mutex);
            out[0]  = search_n_coarse;
            out[1]  = search_refine_points;
            out[2]  = search_max_rounds;
            rel_tol = search_rel_tol;
        }
        void setWavefunctions(std::vector<double> psi, uint32_t n_curves, uint32_t n_levels, uint32_t n_points) {
            std::lock_guard<std::mutex> g(result_mutex);
            wavefunctions = std::move(psi);
            wf_curves     = n_curves;
            wf_levels     = n_levels;
            wf_points     = n_points;
        }
        // Normalised wavefunctions psi[curve][level - min_level][grid point] (h * sum psi^2 = 1, first
        // lobe positive, zero rows for levels that were not found); dims = {curves, levels, points}.
        [[nodiscard]] std::vector<double> getWavefunctions(uint32_t dims[3]) const {
            std::lock_guard<std::mutex> g(result_mutex);
            dims[0] = wf_curves;
            dims[1] = wf_levels;
            dims[2] = wf_points;
            return wavefunctions;
        }
        [[nodiscard]] std::string getStatusMessage() const {
            std::lock_guard<std::mutex> g(result_mutex);
            return status;
        }
        [[nodiscard]] bool hasFailed() const {
            std::lock_guard<std::mutex> g(result_mutex);
            return failed;
        }
        [[nodiscard]] std::vector<std::vector<FP>> getLevels() const {
            std::lock_guard<std::mutex> g(result_mutex);
            return levels;
        }
        [[nodiscard]] std::vector<uint32_t> getLevelCounts() const {
            std::lock_guard<std::mutex> g(result_mutex);
            return level_counts;
        }
        [[nodiscard]] double getDeviceMilliseconds() const {
            std::lock_guard<std::mutex> g(result_mutex);
            return device_ms;
        }
    };
} // namespace epseon::gpu::cpp

#include "epseon/gpu/algorithms/vibwa_run.hpp"
