// Probe results -> NFD node labels that gate scheduling (SURVEY.md §5 "NFD label hand-off", §8f.1).
// Native twin of k3s-nvidia_b200/labels.py: same keys, same rounding, same thresholds, same file format.
//
// GFD (enabled by /root/reference/values.yaml:1-2) hands labels to Node Feature Discovery by writing
// key=value lines into a file under /etc/kubernetes/node-feature-discovery/features.d/ [RECALLED]; the
// reference hints at label gating in the commented selector of /root/reference/nvidia-smi.yaml:6-7 and says
// the plugin "needs these labels for scheduling" (/root/reference/README.md:99).  We write a SECOND file in
// the same directory with nvidia.com/b200probe.* keys, so no chart value changes.
#pragma once
#include <sys/stat.h>
#include <unistd.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <ctime>
#include <map>
#include <string>
#include <thread>
#include <vector>

#include "../../include/b200probe.h"
#include "plugin.hpp"

namespace labels {

static const char kPrefix[] = "nvidia.com/b200probe.";
static const double kHbmNominal = 8000.0, kHbmMeasured = 6565.8;            // north_star's denominator; MEASURED_PEAKS.json hbm_gbs
static const double kNvlinkNominal = 900.0;
static const double kNvlinkHealthyPair = 692.0, kNvlinkHealthyBox = 700.0;  // profiles/a2a_tune_r01_2gpu.txt; profiles/a2a_sync_relaxed_r01_g8.txt
static const double kGemmMeasured = 1670.2;

using Labels = std::map<std::string, std::string>;

inline double env_double(const char* name, double dflt) {
    const char* e = getenv(name);
    if (!e || !*e) return dflt;
    char* end = nullptr;
    const double v = strtod(e, &end);
    return (end && *end == 0) ? v : dflt;
}

struct Thresholds {
    double hbm_min_gbs = env_double("B200PROBE_HBM_MIN_GBS", 0.90 * kHbmMeasured);
    double nvlink_min_gbs = env_double("B200PROBE_NVLINK_MIN_GBS", 0.0);     // 0 = 97 % of the healthy figure for the number of GPUs exchanged
    double gemm_min_tflops = env_double("B200PROBE_GEMM_MIN_TFLOPS", 0.70 * kGemmMeasured);
    uint64_t verdict_min_bytes = 256ull << 20;                                // sizes below L2 are cache-resident: never used for the verdict
};

inline std::string b(bool x) { return x ? "true" : "false"; }
// Python's round(): half to even
inline std::string rint_str(double v) { return std::to_string((long long)std::nearbyint(v)); }
inline std::string key(int gpu, const char* leaf) { return std::string(kPrefix) + "gpu" + std::to_string(gpu) + "." + leaf; }
inline std::string key(const char* leaf) { return std::string(kPrefix) + leaf; }

inline bool valid_label(const std::string& k, const std::string& v) {
    auto name_ok = [](const std::string& s, bool allow_empty) {
        if (s.empty()) return allow_empty;
        if (s.size() > 63 || !isalnum((unsigned char)s.front()) || !isalnum((unsigned char)s.back())) return false;
        for (unsigned char c : s) if (!isalnum(c) && c != '-' && c != '_' && c != '.') return false;
        return true;
    };
    std::string name = k;
    const size_t slash = k.find('/');
    if (slash != std::string::npos) { if (slash > 253) return false; name = k.substr(slash + 1); }
    return name_ok(name, false) && name_ok(v, true);
}

inline Labels hbm_labels(const std::map<int, std::vector<b200probe_hbm_result_t>>& per_gpu, const Thresholds& th) {
    Labels out;
    bool all_ok = true, have_worst = false;
    double worst = 0;
    for (const auto& kv : per_gpu) {
        const int idx = kv.first;
        const b200probe_hbm_result_t* best[3] = {nullptr, nullptr, nullptr};   // read, write, copy at the largest HBM-resident size
        static const int modes[3] = {B200PROBE_HBM_READ, B200PROBE_HBM_WRITE, B200PROBE_HBM_COPY};
        static const char* names[3] = {"read", "write", "copy"};
        bool data_ok = true;
        for (const auto& p : kv.second) {
            if (p.verified == 0) data_ok = false;
            if (p.cache_resident || p.bytes < th.verdict_min_bytes) continue;
            for (int m = 0; m < 3; ++m) if (p.mode == modes[m] && (!best[m] || p.bytes > best[m]->bytes)) best[m] = &p;
        }
        for (int m = 0; m < 3; ++m) if (best[m]) out[key(idx, (std::string("hbm-") + names[m] + "-gbs").c_str())] = rint_str(best[m]->gbs_median);
        const bool ok = data_ok && best[2] && best[2]->gbs_median >= th.hbm_min_gbs;
        if (best[2]) {
            const double g = best[2]->gbs_median;
            out[key(idx, "hbm-copy-pct-of-nominal")] = rint_str(100.0 * g / kHbmNominal);
            out[key(idx, "hbm-copy-pct-of-measured")] = rint_str(100.0 * g / kHbmMeasured);
            worst = have_worst ? std::min(worst, g) : g;
            have_worst = true;
        }
        out[key(idx, "hbm-data-ok")] = b(data_ok);
        out[key(idx, "hbm-healthy")] = b(ok);
        all_ok = all_ok && ok;
    }
    out[key("hbm-healthy")] = b(all_ok && !per_gpu.empty());
    if (have_worst) out[key("hbm-copy-min-gbs")] = rint_str(worst);
    return out;
}

inline Labels nvlink_labels(const b200probe_a2a_result_t& rep, const std::vector<double>& pair_gbs, const Thresholds& th) {
    Labels out;
    const int G = rep.g;
    bool ok = rep.verified != 0;
    const double min_gbs = th.nvlink_min_gbs > 0 ? th.nvlink_min_gbs : (G <= 2 ? 0.97 * kNvlinkHealthyPair : 0.96 * kNvlinkHealthyBox);
    double min_egress = 1e300;
    for (int g = 0; g < G; ++g) {
        out[key(g, "nvlink-egress-gbs")] = rint_str(rep.egress_gbs[g]);
        out[key(g, "nvlink-ingress-gbs")] = rint_str(rep.ingress_gbs[g]);
        const bool good = rep.egress_gbs[g] >= min_gbs;
        out[key(g, "nvlink-healthy")] = b(good && rep.verified != 0);
        ok = ok && good;
        min_egress = std::min(min_egress, rep.egress_gbs[g]);
        for (int p = 0; p < G; ++p)
            if (p != g && pair_gbs[(size_t)(g * G + p)] > 0) out[key(g, ("nvlink-to-gpu" + std::to_string(p) + "-gbs").c_str())] = rint_str(pair_gbs[(size_t)(g * G + p)]);
    }
    out[key("nvlink-min-pair-gbs")] = rint_str(rep.min_pair_gbs);
    out[key("nvlink-egress-pct-of-nominal")] = rint_str(100.0 * min_egress / kNvlinkNominal);
    out[key("nvlink-data-ok")] = b(rep.verified != 0);
    out[key("nvlink-healthy")] = b(ok);
    return out;
}

// Every link NVML knows must be up and the fabric registration completed with a clean health mask (§8f.3).
inline Labels nvlink_passive_labels(const std::map<int, b200probe_nvlink_status_t>& per_gpu, int expected_links = 18) {
    Labels out;
    bool all_ok = true;
    for (const auto& kv : per_gpu) {
        const auto& st = kv.second;
        if (st.links_total == 0) continue;                         // no NVLink on this part: nothing to assert
        bool ok = st.links_active == st.links_total && st.links_total >= expected_links;
        if (st.fabric_state == 3 && st.fabric_status != 0) ok = false;
        if ((st.fabric_health_mask & 0x3) == 1) ok = false;        // DEGRADED_BW == TRUE (nvml.h:3453)
        out[key(kv.first, "nvlink-links-active")] = std::to_string(st.links_active);
        out[key(kv.first, "nvlink-links-total")] = std::to_string(st.links_total);
        out[key(kv.first, "nvlink-links-ok")] = b(ok);
        all_ok = all_ok && ok;
    }
    if (!out.empty()) out[key("nvlink-links-ok")] = b(all_ok);
    return out;
}

inline Labels gemm_labels(const std::map<int, b200probe_gemm_result_t>& per_gpu, const Thresholds& th) {
    Labels out;
    bool all_ok = true;
    for (const auto& kv : per_gpu) {
        const auto& r = kv.second;
        const bool ok = r.verified == 1 && r.tflops_median >= th.gemm_min_tflops;
        out[key(kv.first, "gemm-tflops")] = rint_str(r.tflops_median);
        out[key(kv.first, "gemm-data-ok")] = b(r.verified == 1);
        out[key(kv.first, "gemm-healthy")] = b(ok);
        all_ok = all_ok && ok;
    }
    out[key("gemm-healthy")] = b(all_ok && !per_gpu.empty());
    return out;
}

// The one label manifests select on: every probe that ran is healthy.
inline void gate_label(Labels* l) {
    bool any = false, all = true;
    for (const char* leaf : {"hbm-healthy", "nvlink-healthy", "gemm-healthy"}) {
        auto it = l->find(key(leaf));
        if (it != l->end()) { any = true; all = all && it->second == "true"; }
    }
    (*l)[key("healthy")] = b(any && all);
}

inline std::string render(const Labels& l) {
    std::string text;
    for (const auto& kv : l) {                                         // std::map iterates sorted by key, like sorted(labels.items())
        if (!valid_label(kv.first, kv.second)) throw std::runtime_error("invalid label: " + kv.first + "=" + kv.second);
        text += kv.first + "=" + kv.second + "\n";
    }
    return text;
}

// Atomic replace (NFD may read at any time): write a hidden temp file, then rename.
inline std::string write_feature_file(const Labels& l, const std::string& dir, const std::string& name = "b200probe") {
    std::string cur;
    for (size_t i = 1; i <= dir.size(); ++i)
        if (i == dir.size() || dir[i] == '/') { cur = dir.substr(0, i); ::mkdir(cur.c_str(), 0755); }
    const std::string text = render(l);
    std::string tmpl = dir + "/.b200probe.XXXXXX";                       // dot-files are ignored by NFD
    std::vector<char> tmp(tmpl.begin(), tmpl.end());
    tmp.push_back(0);
    const int fd = ::mkstemp(tmp.data());
    if (fd < 0) throw std::runtime_error("mkstemp in " + dir + ": " + strerror(errno));
    bool ok = true;
    for (size_t off = 0; off < text.size() && ok;) {
        const ssize_t w = ::write(fd, text.data() + off, text.size() - off);
        if (w < 0) { if (errno == EINTR) continue; ok = false; }
        else off += (size_t)w;
    }
    ::fchmod(fd, 0644);
    ::close(fd);
    const std::string path = dir + "/" + name;
    if (!ok || ::rename(tmp.data(), path.c_str()) != 0) { ::unlink(tmp.data()); throw std::runtime_error("writing " + path + " failed"); }
    return path;
}

// Runs the active probes on every enumerated GPU at an interval and publishes the labels.  A probe that
// fails (CUDA error, data mismatch) publishes ...healthy=false; it never blocks ListAndWatch.
class ActiveProbeRunner {
public:
    ActiveProbeRunner(const std::string& features_dir, double interval_s, bool run_nvlink = true, bool run_gemm = true)
        : dir_(features_dir), interval_s_(interval_s), run_nvlink_(run_nvlink), run_gemm_(run_gemm) {}
    ~ActiveProbeRunner() { stop(); }

    Labels run_once() {
        Labels l;
        int n = 0;
        b200probe_device_count(&n);
        std::vector<b200probe_device_t> infos((size_t)n);
        for (int i = 0; i < n; ++i) b200probe_device_info(i, &infos[(size_t)i]);

        std::map<int, std::vector<b200probe_hbm_result_t>> hbm;
        bool hbm_failed = false;
        for (const auto& d : infos) {
            b200probe_hbm_cfg_t cfg;
            memset(&cfg, 0, sizeof(cfg));
            cfg.min_bytes = 1ull << 28; cfg.max_bytes = 1ull << 30; cfg.warmup = 2; cfg.reps = 5; cfg.verify = 1;
            std::vector<b200probe_hbm_result_t> pts(128);
            int got = 0;
            const int rc = b200probe_hbm_sweep(d.index, &cfg, pts.data(), (int)pts.size(), &got);
            if (rc) {
                plugin::logf("HBM probe failed on GPU %d: %s", d.index, b200probe_strerror(rc));
                l[key(d.index, "hbm-healthy")] = "false";
                hbm_failed = true;
            } else { pts.resize((size_t)got); hbm[d.index] = pts; }
        }
        if (!hbm.empty()) {
            Labels got = hbm_labels(hbm, th_);
            for (const auto& kv : got) if (!(l.count(kv.first) && l[kv.first] == "false")) l[kv.first] = kv.second;
        }
        if (hbm_failed) l[key("hbm-healthy")] = "false";

        if (run_gemm_) {
            std::map<int, b200probe_gemm_result_t> gemm;
            bool failed = false;
            for (const auto& d : infos) {
                b200probe_gemm_cfg_t cfg;
                memset(&cfg, 0, sizeof(cfg));
                cfg.warmup = 2; cfg.reps = 5;
                b200probe_gemm_result_t r;
                const int rc = b200probe_gemm(d.index, &cfg, &r);
                if (rc) {
                    plugin::logf("GEMM probe failed on GPU %d: %s", d.index, b200probe_strerror(rc));
                    l[key(d.index, "gemm-healthy")] = "false";
                    failed = true;
                } else gemm[d.index] = r;
            }
            if (!gemm.empty()) { Labels got = gemm_labels(gemm, th_); for (const auto& kv : got) l[kv.first] = kv.second; }
            if (failed) l[key("gemm-healthy")] = "false";
        }

        std::vector<int> ords;
        for (int i = 0; i < n; ++i) { b200probe_device_t d; if (b200probe_device_info(i, &d) == 0 && d.cuda_ordinal >= 0) ords.push_back(d.cuda_ordinal); }
        std::map<int, b200probe_nvlink_status_t> before;
        for (const auto& d : infos) { b200probe_nvlink_status_t st; if (b200probe_nvlink_passive(d.index, &st) == 0) before[d.index] = st; }
        { Labels got = nvlink_passive_labels(before); for (const auto& kv : got) l[kv.first] = kv.second; }
        if (run_nvlink_ && ords.size() >= 2) {
            const int G = (int)ords.size();
            b200probe_a2a_cfg_t cfg;
            memset(&cfg, 0, sizeof(cfg));
            cfg.warmup = 1; cfg.reps = 3; cfg.verify = 1;
            std::vector<double> pair((size_t)(G * G), 0.0);
            b200probe_a2a_result_t rep;
            const int rc = b200probe_nvlink_a2a(ords.data(), G, &cfg, pair.data(), &rep);
            if (rc) {
                plugin::logf("NVLink probe failed: %s", b200probe_strerror(rc));
                l[key("nvlink-healthy")] = "false";
            } else {
                Labels got = nvlink_labels(rep, pair, th_);
                for (const auto& kv : got) l[kv.first] = kv.second;
                auto it = l.find(key("nvlink-links-ok"));
                if (it != l.end() && it->second == "false") l[key("nvlink-healthy")] = "false";   // a dead link fails the gate even if the matrix clears the bar
                double min_eff = 2.0;
                for (const auto& kv : before) {
                    b200probe_nvlink_status_t after;
                    if (b200probe_nvlink_passive(kv.first, &after) != 0 || !kv.second.counters_ok || !after.counters_ok) continue;
                    const double dd = (double)(after.data_tx_kib - kv.second.data_tx_kib), rr = (double)(after.raw_tx_kib - kv.second.raw_tx_kib);
                    if (rr > 0 && dd > 0) min_eff = std::min(min_eff, dd / rr);
                }
                if (min_eff <= 1.0) l[key("nvlink-data-over-raw-pct")] = rint_str(100.0 * min_eff);
            }
        }
        gate_label(&l);
        l[key("timestamp")] = std::to_string((long long)time(nullptr));
        write_feature_file(l, dir_);
        return l;
    }

    void start() {
        stop_ = false;
        thread_ = std::thread([this] {
            while (!stop_) {
                try { run_once(); } catch (const std::exception& e) { plugin::logf("active probe round failed: %s", e.what()); }
                std::unique_lock<std::mutex> lk(mu_);
                cv_.wait_for(lk, std::chrono::milliseconds((long long)(interval_s_ * 1000)), [this] { return stop_.load(); });
            }
        });
    }
    void stop() {
        stop_ = true;
        cv_.notify_all();
        if (thread_.joinable()) thread_.join();
    }

private:
    std::string dir_;
    double interval_s_;
    bool run_nvlink_, run_gemm_;
    Thresholds th_;
    std::atomic<bool> stop_{false};
    std::mutex mu_;
    std::condition_variable cv_;
    std::thread thread_;
};

}  // namespace labels
