// The device-plugin host: Register / ListAndWatch / Allocate / GetPreferredAllocation / PreStartContainer
// over a unix socket, on top of the C ABI (include/b200probe.h).  The native twin of
// k3s-nvidia_b200/plugin.py — same names, same argument meaning, same error texts — standing where the
// Go binary of the chart that /root/reference/README.md:116 installs stands, configured by
// /root/reference/values.yaml (resource nvidia.com/gpu :17, 4 time-sliced replicas :18, no rename :14,
// multi-replica requests allowed :15, MIG off :11).  Upstream behaviour is [RECALLED]:
//   - advertised IDs `<GPU-UUID>::<replica>`; Allocate strips the suffix, dedupes, returns
//     NVIDIA_VISIBLE_DEVICES=<uuid[,uuid…]>;
//   - Device.health is the PASSIVE verdict only (XID/ECC event loop, bit-exact with the oracle); active
//     probe outcomes go to NFD labels (labels.hpp), never into health;
//   - ListAndWatch re-sends the complete list on every change; there is no path back to Healthy;
//   - kubelet restart (socket re-created) => serve again and re-Register.
#pragma once
#include <sys/stat.h>

#include <algorithm>
#include <cstdarg>
#include <ctime>
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <set>

#include "../../include/b200probe.h"
#include "config.hpp"
#include "h2.hpp"
#include "v1beta1.hpp"

namespace plugin {

inline void logf(const char* fmt, ...) __attribute__((format(printf, 1, 2)));
inline void logf(const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    struct timespec ts;
    clock_gettime(CLOCK_REALTIME, &ts);
    struct tm tm;
    gmtime_r(&ts.tv_sec, &tm);
    fprintf(stderr, "%04d-%02d-%02dT%02d:%02d:%02d.%03ldZ b200-device-plugin %s\n", tm.tm_year + 1900, tm.tm_mon + 1, tm.tm_mday, tm.tm_hour, tm.tm_min,
            tm.tm_sec, ts.tv_nsec / 1000000, buf);
}

struct AdvertisedDevice {
    std::string id;        // annotated ID kubelet sees
    std::string uuid;      // physical GPU
    int index = 0;         // NVML index
    int numa_node = -1;
    std::string health = v1beta1::kHealthy;
};

struct AllocationError : std::runtime_error { using std::runtime_error::runtime_error; };

// Physical GPUs -> advertised devices (replica expansion of values.yaml:16-18).
inline std::vector<AdvertisedDevice> build_devices(const std::vector<b200probe_device_t>& infos, const config::PluginConfig& cfg,
                                                   const std::string& resource = config::kDefaultResource) {
    const int64_t replicas = cfg.replicas(resource);
    std::vector<AdvertisedDevice> out;
    for (const auto& d : infos) {
        if (cfg.mig_strategy == "none" || d.mig_enabled <= 0) {
            for (int64_t r = 0; r < std::max<int64_t>(replicas, 1); ++r) {
                AdvertisedDevice a;
                a.uuid = d.uuid;
                a.id = replicas <= 1 ? a.uuid : config::annotate(a.uuid, r);
                a.index = d.index;
                a.numa_node = d.numa_node;
                out.push_back(a);
            }
        }   // MIG strategies single/mixed are outside this path (values.yaml:11 selects "none")
    }
    return out;
}

// Preferred allocation for replicated devices [RECALLED upstream distributedAlloc]: spread the request over
// the physical GPUs with the fewest replicas already handed out; ties break by the order of `available`.
inline std::vector<std::string> distributed_alloc(const std::vector<std::string>& all_ids, const std::vector<std::string>& available,
                                                  const std::vector<std::string>& required, int size) {
    const std::set<std::string> known(all_ids.begin(), all_ids.end()), req(required.begin(), required.end());
    std::vector<std::string> candidates;
    for (const auto& i : available) if (known.count(i) && !req.count(i)) candidates.push_back(i);
    int needed = size;
    for (const auto& r : required) if (known.count(r)) --needed;
    if (needed < 0) needed = 0;
    if ((int)candidates.size() < needed) throw AllocationError("not enough available devices to satisfy allocation");
    std::map<std::string, int> total, free_;
    for (const auto& c : candidates) ++free_[config::strip_replica(c)];
    for (const auto& d : known) { const std::string u = config::strip_replica(d); if (free_.count(u)) ++total[u]; }
    std::vector<std::string> picked;
    for (int n = 0; n < needed; ++n) {
        std::stable_sort(candidates.begin(), candidates.end(), [&](const std::string& a, const std::string& b) {
            const std::string ua = config::strip_replica(a), ub = config::strip_replica(b);
            return total[ua] - free_[ua] < total[ub] - free_[ub];
        });
        const std::string c = candidates.front();
        candidates.erase(candidates.begin());
        --free_[config::strip_replica(c)];
        picked.push_back(c);
    }
    std::vector<std::string> out(required);
    out.insert(out.end(), picked.begin(), picked.end());
    return out;
}

class DevicePlugin {
public:
    DevicePlugin(const config::PluginConfig& cfg, const std::string& socket_dir, const std::string& kubelet_socket = "", int health_timeout_ms = 5000,
                 const char* disable_healthchecks = nullptr, const std::string& resource = config::kDefaultResource)
        : cfg_(cfg), base_resource_(resource), resource_(cfg.resource_name(resource)), socket_dir_(socket_dir), health_timeout_ms_(health_timeout_ms) {
        std::string leaf = resource_.substr(resource_.find('/') + 1);
        std::replace(leaf.begin(), leaf.end(), '.', '-');
        endpoint_ = "nvidia-" + leaf + ".sock";
        socket_path_ = join(socket_dir_, endpoint_);
        kubelet_socket_ = kubelet_socket.empty() ? join(socket_dir_, "kubelet.sock") : kubelet_socket;
        const char* env = getenv("DP_DISABLE_HEALTHCHECKS");
        disable_healthchecks_ = disable_healthchecks ? disable_healthchecks : (env ? env : "");
        refresh_devices();
    }
    ~DevicePlugin() { stop(); }

    const std::string& resource() const { return resource_; }
    const std::string& socket_path() const { return socket_path_; }
    int registrations() const { return registrations_.load(); }

    // ---- device list ------------------------------------------------------------------------------
    void refresh_devices() {
        int n = 0;
        int rc = b200probe_device_count(&n);
        if (rc) throw std::runtime_error(std::string("device_count: ") + b200probe_strerror(rc));
        std::vector<b200probe_device_t> infos((size_t)n);
        for (int i = 0; i < n; ++i) {
            rc = b200probe_device_info(i, &infos[(size_t)i]);
            if (rc) throw std::runtime_error(std::string("device_info: ") + b200probe_strerror(rc));
        }
        std::lock_guard<std::mutex> l(mu_);
        devices_ = build_devices(infos, cfg_, base_resource_);
        ++generation_;
        cv_.notify_all();
    }
    std::vector<AdvertisedDevice> devices() { std::lock_guard<std::mutex> l(mu_); return devices_; }

    // Bit i of mask = physical GPU with NVML index i.  Every replica of that GPU turns Unhealthy (health is a
    // property of the physical device).  Returns true on a change.
    bool mark_unhealthy_mask(uint64_t mask) {
        bool changed = false;
        std::lock_guard<std::mutex> l(mu_);
        for (auto& d : devices_) {
            if (d.index < 64 && ((mask >> d.index) & 1) && d.health != v1beta1::kUnhealthy) {
                d.health = v1beta1::kUnhealthy;
                changed = true;
                logf("'%s' device marked unhealthy: %s", resource_.c_str(), d.id.c_str());
            }
        }
        if (changed) { ++generation_; cv_.notify_all(); }
        return changed;
    }

    // ---- RPC bodies (also called directly by unit checks) ------------------------------------------------
    v1beta1::ContainerAllocateResponse allocate_container(const std::vector<std::string>& ids) {
        const std::vector<AdvertisedDevice> devs = devices();
        std::set<std::string> known;
        for (const auto& d : devs) known.insert(d.id);
        if (cfg_.is_shared(base_resource_) && cfg_.time_slicing.fail_requests_greater_than_one && ids.size() > 1)
            throw AllocationError("request for '" + resource_ + ": " + std::to_string(ids.size()) + "' too large: maximum request size for shared resources is 1");
        for (const auto& i : ids)
            if (!known.count(i)) throw AllocationError("invalid allocation request for '" + resource_ + "': unknown device: " + i);
        std::vector<std::string> uuids;
        for (const auto& i : ids) {
            const std::string u = config::strip_replica(i);
            if (std::find(uuids.begin(), uuids.end(), u) == uuids.end()) uuids.push_back(u);
        }
        std::string visible;
        for (const auto& u : uuids) {
            std::string v = u;
            if (cfg_.device_id_strategy == "index")
                for (const auto& d : devs) if (d.uuid == u) { v = std::to_string(d.index); break; }
            visible += (visible.empty() ? "" : ",") + v;
        }
        v1beta1::ContainerAllocateResponse r;
        r.envs["NVIDIA_VISIBLE_DEVICES"] = visible;
        return r;
    }

    std::vector<std::string> preferred_container(const v1beta1::ContainerPreferredAllocationRequest& req) {
        std::vector<std::string> all_ids;
        for (const auto& d : devices()) all_ids.push_back(d.id);
        if (cfg_.is_shared(base_resource_)) return distributed_alloc(all_ids, req.available, req.must_include, req.allocation_size);
        // Unshared: upstream packs by NVLink topology; behind NVSwitch every pair is equidistant, so
        // required-first then availability order is an equivalent choice.
        std::vector<std::string> ids(req.must_include);
        const std::set<std::string> rs(req.must_include.begin(), req.must_include.end());
        std::vector<std::string> rest;
        for (const auto& i : req.available) if (!rs.count(i)) rest.push_back(i);
        if ((int)(ids.size() + rest.size()) < req.allocation_size) throw AllocationError("not enough available devices to satisfy allocation");
        ids.insert(ids.end(), rest.begin(), rest.end());
        ids.resize(std::max<size_t>((size_t)std::max(req.allocation_size, 0), req.must_include.size()));
        return ids;
    }

    // ---- lifecycle ---------------------------------------------------------------------------------
    void serve() {
        ::mkdir(socket_dir_.c_str(), 0755);
        server_.reset(new h2::Server());
        server_->route("/v1beta1.DevicePlugin/GetDevicePluginOptions", [this](std::shared_ptr<h2::ServerCall> c) {
            v1beta1::DevicePluginOptions o;
            o.pre_start_required = false; o.get_preferred_allocation_available = true;
            c->send(o.encode());
            c->finish({});
        }, false);
        server_->route("/v1beta1.DevicePlugin/ListAndWatch", [this](std::shared_ptr<h2::ServerCall> c) { list_and_watch(c); }, true);
        server_->route("/v1beta1.DevicePlugin/GetPreferredAllocation", [this](std::shared_ptr<h2::ServerCall> c) {
            std::vector<v1beta1::ContainerPreferredAllocationRequest> reqs;
            if (!v1beta1::decode_preferred_request(c->request(), &reqs)) { c->finish({h2::INTERNAL, "malformed PreferredAllocationRequest"}); return; }
            std::vector<std::vector<std::string>> out;
            try {
                for (const auto& r : reqs) out.push_back(preferred_container(r));
            } catch (const AllocationError& e) {
                c->finish({h2::UNKNOWN, std::string("error getting list of preferred allocation devices: ") + e.what()});
                return;
            }
            c->send(v1beta1::encode_preferred_response(out));
            c->finish({});
        }, false);
        server_->route("/v1beta1.DevicePlugin/Allocate", [this](std::shared_ptr<h2::ServerCall> c) {
            std::vector<v1beta1::ContainerAllocateRequest> reqs;
            if (!v1beta1::decode_allocate_request(c->request(), &reqs)) { c->finish({h2::INTERNAL, "malformed AllocateRequest"}); return; }
            std::vector<v1beta1::ContainerAllocateResponse> out;
            try {
                for (const auto& r : reqs) out.push_back(allocate_container(r.devices_ids));
            } catch (const AllocationError& e) {
                c->finish({h2::UNKNOWN, e.what()});
                return;
            }
            c->send(v1beta1::encode_allocate_response(out));
            c->finish({});
        }, false);
        server_->route("/v1beta1.DevicePlugin/PreStartContainer", [](std::shared_ptr<h2::ServerCall> c) { c->send(""); c->finish({}); }, false);
        std::string err;
        if (!server_->listen_unix(socket_path_, &err)) throw std::runtime_error(err);
    }

    void do_register(int timeout_ms = 5000) {
        v1beta1::RegisterRequest r;
        r.version = v1beta1::kVersion; r.endpoint = endpoint_; r.resource_name = resource_;
        r.options.pre_start_required = false; r.options.get_preferred_allocation_available = true;
        const h2::Status st = h2::unary_call(kubelet_socket_, "/v1beta1.Registration/Register", r.encode(), nullptr, timeout_ms);
        if (!st.ok()) throw std::runtime_error("Register with kubelet failed: code " + std::to_string(st.code) + ": " + st.message);
        ++registrations_;
        logf("Registered device plugin for '%s' with Kubelet", resource_.c_str());
    }

    void start(double watch_kubelet_period_s = 1.0, bool health = true) {
        stop_ = false;
        serve();
        do_register();
        if (health) threads_.emplace_back([this] { health_loop(); });
        if (watch_kubelet_period_s > 0) threads_.emplace_back([this, watch_kubelet_period_s] { kubelet_watch(watch_kubelet_period_s); });
    }

    void stop() {
        {
            std::lock_guard<std::mutex> l(mu_);
            stop_ = true;
            cv_.notify_all();
        }
        for (auto& t : threads_) if (t.joinable()) t.join();
        threads_.clear();
        if (server_) { server_->stop(); server_.reset(); }
        b200probe_health_close();
    }

private:
    static std::string join(const std::string& dir, const std::string& leaf) { return dir.empty() || dir.back() == '/' ? dir + leaf : dir + "/" + leaf; }

    std::string list_message() {                       // caller holds mu_
        std::vector<v1beta1::Device> devs;
        for (const auto& d : devices_) {
            v1beta1::Device x;
            x.id = d.id; x.health = d.health; x.numa_node = d.numa_node;
            devs.push_back(x);
        }
        return v1beta1::encode_list_and_watch(devs);
    }

    void list_and_watch(const std::shared_ptr<h2::ServerCall>& c) {
        uint64_t gen;
        std::string msg;
        { std::lock_guard<std::mutex> l(mu_); gen = generation_; msg = list_message(); }
        if (!c->send(msg)) return;
        for (;;) {
            {
                std::unique_lock<std::mutex> l(mu_);
                cv_.wait_for(l, std::chrono::milliseconds(200), [&] { return generation_ != gen || stop_; });
                if (stop_ || c->cancelled()) break;
                if (generation_ == gen) continue;
                gen = generation_;
                msg = list_message();
            }
            if (!c->send(msg)) return;                  // the complete list, every time
        }
        c->finish({});
    }

    void health_loop() {
        uint64_t at_open = 0;
        int rc = b200probe_health_open(disable_healthchecks_.c_str(), &at_open);
        if (rc) { logf("health_open failed: %s", b200probe_strerror(rc)); return; }
        if (at_open) mark_unhealthy_mask(at_open);
        while (!stopped()) {
            b200probe_health_event_t ev;
            memset(&ev, 0, sizeof(ev));
            // short waits so stop() is prompt; the event set itself is level-triggered and loses nothing
            rc = b200probe_health_wait(std::min(health_timeout_ms_, 200), &ev);
            if (rc) { logf("health_wait failed: %s", b200probe_strerror(rc)); break; }
            if (ev.newly_unhealthy) {
                logf("XidCriticalError: Xid=%llu on device %d; marking device as unhealthy", (unsigned long long)ev.event_data, ev.device_index);
                mark_unhealthy_mask(ev.newly_unhealthy);
            }
            // a wait that keeps failing (GPU_IS_LOST, UNINITIALIZED ...) returns at once: every device is already Unhealthy and
            // there is nothing left to learn fast, so do not spin a core on it
            if (ev.rc_wait != 0 /*NVML_SUCCESS*/ && ev.rc_wait != 10 /*NVML_ERROR_TIMEOUT*/)
                for (int i = 0; i < 20 && !stopped(); ++i) std::this_thread::sleep_for(std::chrono::milliseconds(100));
        }
    }

    // kubelet restart re-creates its socket: serve again and re-Register.
    void kubelet_watch(double period_s) {
        auto ident = [this](uint64_t* ino, int64_t* ctime_ns) {
            struct stat st;
            if (::stat(kubelet_socket_.c_str(), &st) != 0) return false;
            *ino = st.st_ino;
            *ctime_ns = (int64_t)st.st_ctim.tv_sec * 1000000000ll + st.st_ctim.tv_nsec;
            return true;
        };
        uint64_t last_ino = 0; int64_t last_ct = 0;
        bool have_last = ident(&last_ino, &last_ct);
        while (!stopped()) {
            {
                std::unique_lock<std::mutex> l(mu_);
                cv_.wait_for(l, std::chrono::milliseconds((int)(period_s * 1000)), [&] { return stop_; });
                if (stop_) return;
            }
            uint64_t ino = 0; int64_t ct = 0;
            if (!ident(&ino, &ct)) continue;
            if (!have_last || ino != last_ino || ct != last_ct) {
                logf("inotify: %s created, restarting.", kubelet_socket_.c_str());
                try {
                    if (server_) server_->stop();
                    serve();
                    do_register();
                    last_ino = ino; last_ct = ct; have_last = true;
                } catch (const std::exception& e) {
                    logf("restart after kubelet restart failed: %s", e.what());
                }
            }
        }
    }
    bool stopped() { std::lock_guard<std::mutex> l(mu_); return stop_; }

    config::PluginConfig cfg_;
    std::string base_resource_, resource_, socket_dir_, endpoint_, socket_path_, kubelet_socket_, disable_healthchecks_;
    int health_timeout_ms_;
    std::mutex mu_;
    std::condition_variable cv_;
    uint64_t generation_ = 0;
    bool stop_ = false;
    std::vector<AdvertisedDevice> devices_;
    std::unique_ptr<h2::Server> server_;
    std::vector<std::thread> threads_;
    std::atomic<int> registrations_{0};
};

}  // namespace plugin
