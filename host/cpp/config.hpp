// The plugin configuration document and the time-slicing replica semantics
// (/root/reference/values.yaml:9-18; "treat that one GPU as if it were actually four",
// /root/reference/README.md:112).  Same rules, names and error texts as the Python twin
// k3s-nvidia_b200/config.py; how upstream interprets the document is [RECALLED] (un-pinned chart).
#pragma once
#include <string>
#include <vector>

#include "yaml.hpp"

namespace config {

static const char kResourcePrefix[] = "nvidia.com/";
static const char kDefaultResource[] = "nvidia.com/gpu";
static const char kSharedSuffix[] = ".shared";
static const char kReplicaSep[] = "::";

struct Error : std::runtime_error { using std::runtime_error::runtime_error; };

struct ReplicatedResource {
    std::string name;
    int64_t replicas = 1;
    std::string rename;                 // empty = none
};

struct TimeSlicing {
    bool rename_by_default = false;
    bool fail_requests_greater_than_one = false;
    std::vector<ReplicatedResource> resources;
};

struct PluginConfig {
    std::string version = "v1";
    std::string mig_strategy = "none";
    std::string device_list_strategy = "envvar";
    std::string device_id_strategy = "uuid";
    bool pass_device_specs = false;
    TimeSlicing time_slicing;

    const ReplicatedResource* replicated(const std::string& resource = kDefaultResource) const {
        for (const auto& r : time_slicing.resources) if (r.name == resource) return &r;
        return nullptr;
    }
    // advertised extended-resource name: unchanged unless renamed (values.yaml:14 false)
    std::string resource_name(const std::string& resource = kDefaultResource) const {
        const ReplicatedResource* r = replicated(resource);
        if (!r) return resource;
        if (!r->rename.empty()) return r->rename;
        if (time_slicing.rename_by_default) return resource + kSharedSuffix;
        return resource;
    }
    int64_t replicas(const std::string& resource = kDefaultResource) const { const auto* r = replicated(resource); return r ? r->replicas : 1; }
    bool is_shared(const std::string& resource = kDefaultResource) const { const auto* r = replicated(resource); return r && r->replicas > 1; }
};

inline std::string repr(const yaml::Node& n) {
    switch (n.kind) {
        case yaml::Node::Null: return "None";
        case yaml::Node::Bool: return n.b ? "True" : "False";
        case yaml::Node::Int: return std::to_string(n.i);
        case yaml::Node::Str: return "'" + n.s + "'";
        case yaml::Node::Map: return "{...}";
        default: return "[...]";
    }
}

inline bool as_bool(const yaml::Node& n, const std::string& what) {
    if (n.kind != yaml::Node::Bool) throw Error(what + " must be a boolean, got " + repr(n));
    return n.b;
}

inline std::string resource_name_checked(const yaml::Node& n) {
    if (n.kind != yaml::Node::Str || n.s.empty()) throw Error("resource name must be a non-empty string, got " + repr(n));
    std::string name = n.s;
    if (name.find('/') == std::string::npos) name = kResourcePrefix + name;
    if (name.compare(0, sizeof(kResourcePrefix) - 1, kResourcePrefix) != 0) throw Error("resource name '" + name + "' must start with 'nvidia.com/'");
    if (name.size() > 63) throw Error("resource name '" + name + "' longer than 63 characters");
    return name;
}

inline const yaml::Node* child_map(const yaml::Node& parent, const char* key, const char* what) {
    const yaml::Node* n = parent.get(key);
    if (!n || n->is_null()) return nullptr;
    if (n->kind != yaml::Node::Map) {
        // Python's `x.get(k) or {}` accepts any falsy value (0, "", [], false) as "absent"
        const bool falsy = (n->kind == yaml::Node::Bool && !n->b) || (n->kind == yaml::Node::Int && n->i == 0) || (n->kind == yaml::Node::Str && n->s.empty()) ||
                           (n->kind == yaml::Node::Seq && n->seq.empty());
        if (falsy) return nullptr;
        throw Error(std::string(what) + " must be a mapping");
    }
    return n->map.empty() ? nullptr : n;
}

inline PluginConfig parse_plugin_config(const std::string& text) {
    yaml::Node doc;
    try { doc = yaml::parse(text); } catch (const yaml::Error& e) { throw Error(std::string("plugin config is not valid YAML: ") + e.what()); }
    if (doc.is_null()) { doc.kind = yaml::Node::Map; }
    if (doc.kind != yaml::Node::Map) throw Error("plugin config must be a mapping");
    const yaml::Node* version = doc.get("version");
    if (!version || version->kind != yaml::Node::Str || version->s != "v1")
        throw Error("unknown version: " + (version ? repr(*version) : std::string("None")) + " (expected 'v1')");
    PluginConfig cfg;
    if (const yaml::Node* flags = child_map(doc, "flags", "flags")) {
        if (const yaml::Node* mig = flags->get("migStrategy")) {
            if (mig->kind != yaml::Node::Str || (mig->s != "none" && mig->s != "single" && mig->s != "mixed")) throw Error("invalid migStrategy " + repr(*mig));
            cfg.mig_strategy = mig->s;
        }
        if (const yaml::Node* plugin = child_map(*flags, "plugin", "flags.plugin")) {
            if (const yaml::Node* v = plugin->get("deviceListStrategy")) cfg.device_list_strategy = v->kind == yaml::Node::Str ? v->s : repr(*v);
            if (const yaml::Node* v = plugin->get("deviceIDStrategy")) cfg.device_id_strategy = v->kind == yaml::Node::Str ? v->s : repr(*v);
            if (const yaml::Node* v = plugin->get("passDeviceSpecs"))
                cfg.pass_device_specs = (v->kind == yaml::Node::Bool && v->b) || (v->kind == yaml::Node::Int && v->i) || (v->kind == yaml::Node::Str && !v->s.empty());
            if (cfg.device_id_strategy != "uuid" && cfg.device_id_strategy != "index") throw Error("invalid deviceIDStrategy '" + cfg.device_id_strategy + "'");
        }
    }
    const yaml::Node* sharing = child_map(doc, "sharing", "sharing");
    const yaml::Node* ts = sharing ? child_map(*sharing, "timeSlicing", "sharing.timeSlicing") : nullptr;
    if (ts) {
        TimeSlicing t;
        if (const yaml::Node* v = ts->get("renameByDefault")) t.rename_by_default = as_bool(*v, "renameByDefault");
        if (const yaml::Node* v = ts->get("failRequestsGreaterThanOne")) t.fail_requests_greater_than_one = as_bool(*v, "failRequestsGreaterThanOne");
        const yaml::Node* res = ts->get("resources");
        if (res && !res->is_null()) {
            if (res->kind != yaml::Node::Seq) throw Error("sharing.timeSlicing.resources must be a list");
            for (const yaml::Node& r : res->seq) {
                if (r.kind != yaml::Node::Map) throw Error("each replicated resource must be a mapping");
                const yaml::Node* name = r.get("name");
                const yaml::Node* rep = r.get("replicas");
                if (!name) throw Error("replicated resource is missing a 'name' field");
                if (!rep) throw Error("replicated resource is missing a 'replicas' field");
                ReplicatedResource rr;
                rr.name = resource_name_checked(*name);
                if (rep->kind != yaml::Node::Int) throw Error("replicas must be an integer, got " + repr(*rep));
                if (rep->i < 1) throw Error("number of replicas must be >= 1, got " + std::to_string(rep->i));
                rr.replicas = rep->i;
                for (const auto& seen : t.resources) if (seen.name == rr.name) throw Error("duplicate replicated resource '" + rr.name + "'");
                if (const yaml::Node* rename = r.get("rename")) if (!rename->is_null()) rr.rename = resource_name_checked(*rename);
                t.resources.push_back(rr);
            }
        }
        cfg.time_slicing = t;
    }
    return cfg;
}

// The Helm values file exactly as the reference ships it (/root/reference/values.yaml:1-18): gfd.enabled (:1-2),
// runtimeClassName (:4), config.map.<name> = the plugin config documents as YAML block scalars (:6-18).
struct HelmValues {
    bool gfd_enabled = false;
    std::string runtime_class_name;                                   // empty = unset
    std::vector<std::pair<std::string, std::string>> raw_configs;      // name -> document text, file order
    std::vector<std::pair<std::string, PluginConfig>> configs;

    // "default" if present, else the only one, else the built-in defaults (config.py HelmValues.default)
    PluginConfig default_config() const {
        for (const auto& kv : configs) if (kv.first == "default") return kv.second;
        if (configs.size() == 1) return configs[0].second;
        return PluginConfig();
    }
};

inline HelmValues parse_helm_values(const std::string& text) {
    yaml::Node doc;
    try { doc = yaml::parse(text); } catch (const yaml::Error& e) { throw Error(std::string("values.yaml is not valid YAML: ") + e.what()); }
    if (doc.is_null()) doc.kind = yaml::Node::Map;
    if (doc.kind != yaml::Node::Map) throw Error("values.yaml must be a mapping");
    HelmValues v;
    if (const yaml::Node* gfd = child_map(doc, "gfd", "gfd"))
        if (const yaml::Node* en = gfd->get("enabled")) v.gfd_enabled = (en->kind == yaml::Node::Bool && en->b) || (en->kind == yaml::Node::Int && en->i) || (en->kind == yaml::Node::Str && !en->s.empty());
    if (const yaml::Node* rc = doc.get("runtimeClassName")) if (rc->kind == yaml::Node::Str) v.runtime_class_name = rc->s;
    if (const yaml::Node* cfg = child_map(doc, "config", "config"))
        if (const yaml::Node* map = child_map(*cfg, "map", "config.map"))
            for (const auto& kv : map->map) {
                if (kv.second.kind != yaml::Node::Str) throw Error("config.map." + kv.first + " must be a YAML string (block scalar)");
                v.raw_configs.emplace_back(kv.first, kv.second.s);
                v.configs.emplace_back(kv.first, parse_plugin_config(kv.second.s));
            }
    return v;
}

// ---- replica annotation [RECALLED upstream AnnotatedID] -----------------------------------------
inline std::string annotate(const std::string& uuid, int64_t replica) { return uuid + kReplicaSep + std::to_string(replica); }
inline bool has_replica(const std::string& id) { return id.find(kReplicaSep) != std::string::npos; }
inline std::string strip_replica(const std::string& id) { const size_t p = id.find(kReplicaSep); return p == std::string::npos ? id : id.substr(0, p); }

}  // namespace config
