// kubelet device-plugin API v1beta1 — the messages this host reads and writes, hand-encoded in the
// protobuf wire format (no protoc / libprotobuf in the image).  Package `v1beta1`, message and field
// NUMBERS as in k8s.io/kubelet/pkg/apis/deviceplugin/v1beta1/api.proto [RECALLED; the .proto is not
// in /root/reference and cannot be fetched — SURVEY.md §8b.  The Python twin k3s-nvidia_b200/api.py
// restates the same numbers as run-time descriptors; tests/test_native_plugin.py drives this encoder
// with that decoder and vice versa].
#pragma once
#include <cstdint>
#include <cstring>
#include <map>
#include <string>
#include <vector>

namespace pb {

// ---- wire primitives ----------------------------------------------------------------------------
inline void put_varint(std::string* o, uint64_t v) {
    while (v >= 0x80) { o->push_back((char)(v | 0x80)); v >>= 7; }
    o->push_back((char)v);
}
inline void put_tag(std::string* o, int field, int wt) { put_varint(o, ((uint64_t)field << 3) | (uint64_t)wt); }
inline void put_bytes(std::string* o, int field, const std::string& s) { put_tag(o, field, 2); put_varint(o, s.size()); o->append(s); }
inline void put_string(std::string* o, int field, const std::string& s) { if (!s.empty()) put_bytes(o, field, s); }          // proto3: defaults are not sent
inline void put_bool(std::string* o, int field, bool v) { if (v) { put_tag(o, field, 0); put_varint(o, 1); } }
inline void put_int(std::string* o, int field, int64_t v) { if (v) { put_tag(o, field, 0); put_varint(o, (uint64_t)v); } }

struct Reader {
    const uint8_t* p;
    const uint8_t* end;
    explicit Reader(const std::string& s) : p((const uint8_t*)s.data()), end((const uint8_t*)s.data() + s.size()) {}
    bool done() const { return p >= end; }
    bool varint(uint64_t* v) {
        uint64_t x = 0;
        for (int shift = 0; shift < 64; shift += 7) {
            if (p >= end) return false;
            const uint8_t b = *p++;
            x |= (uint64_t)(b & 0x7f) << shift;
            if (!(b & 0x80)) { *v = x; return true; }
        }
        return false;
    }
    // next field: number, wire type, and for length-delimited fields the bytes; other types land in *val
    bool next(int* field, int* wt, uint64_t* val, std::string* bytes) {
        uint64_t tag;
        if (!varint(&tag)) return false;
        *field = (int)(tag >> 3);
        *wt = (int)(tag & 7);
        if (*field == 0) return false;
        switch (*wt) {
            case 0: return varint(val);
            case 1: if (end - p < 8) return false; memcpy(val, p, 8); p += 8; return true;
            case 2: {
                uint64_t n;
                if (!varint(&n) || (uint64_t)(end - p) < n) return false;
                bytes->assign((const char*)p, (size_t)n);
                p += n;
                return true;
            }
            case 5: if (end - p < 4) return false; *val = 0; memcpy(val, p, 4); p += 4; return true;
            default: return false;                          // groups are not used by this API
        }
    }
};

}  // namespace pb

namespace v1beta1 {

static const char kVersion[] = "v1beta1";
static const char kHealthy[] = "Healthy";
static const char kUnhealthy[] = "Unhealthy";
static const char kDevicePluginPath[] = "/var/lib/kubelet/device-plugins/";

struct DevicePluginOptions {
    bool pre_start_required = false, get_preferred_allocation_available = false;
    std::string encode() const { std::string o; pb::put_bool(&o, 1, pre_start_required); pb::put_bool(&o, 2, get_preferred_allocation_available); return o; }
};

struct RegisterRequest {
    std::string version, endpoint, resource_name;
    DevicePluginOptions options;
    std::string encode() const {
        std::string o;
        pb::put_string(&o, 1, version); pb::put_string(&o, 2, endpoint); pb::put_string(&o, 3, resource_name);
        pb::put_bytes(&o, 4, options.encode());
        return o;
    }
};

struct Device {
    std::string id, health;
    int64_t numa_node = -1;                 // < 0: no topology
    std::string encode() const {
        std::string o;
        pb::put_string(&o, 1, id); pb::put_string(&o, 2, health);
        if (numa_node >= 0) {
            std::string node, topo;
            pb::put_int(&node, 1, numa_node);               // NUMANode.ID; node 0 encodes as an empty message
            pb::put_bytes(&topo, 1, node);                  // TopologyInfo.nodes
            pb::put_bytes(&o, 3, topo);
        }
        return o;
    }
};

inline std::string encode_list_and_watch(const std::vector<Device>& devs) {
    std::string o;
    for (const Device& d : devs) pb::put_bytes(&o, 1, d.encode());
    return o;
}

// repeated string field `field` of a message
inline bool repeated_strings(const std::string& msg, int want, std::vector<std::string>* out, std::map<int, int64_t>* ints = nullptr) {
    pb::Reader r(msg);
    while (!r.done()) {
        int f, wt; uint64_t v = 0; std::string b;
        if (!r.next(&f, &wt, &v, &b)) return false;
        if (wt == 2 && f == want) out->push_back(b);
        else if (wt == 0 && ints) (*ints)[f] = (int64_t)v;
    }
    return true;
}

struct ContainerAllocateRequest { std::vector<std::string> devices_ids; };
inline bool decode_allocate_request(const std::string& msg, std::vector<ContainerAllocateRequest>* out) {
    std::vector<std::string> containers;
    if (!repeated_strings(msg, 1, &containers)) return false;
    for (const std::string& c : containers) {
        ContainerAllocateRequest r;
        if (!repeated_strings(c, 1, &r.devices_ids)) return false;
        out->push_back(std::move(r));
    }
    return true;
}

struct ContainerAllocateResponse {
    std::map<std::string, std::string> envs;
    std::string encode() const {
        std::string o;
        for (const auto& kv : envs) {                       // map<string,string> = repeated entry {key = 1, value = 2}
            std::string e;
            pb::put_string(&e, 1, kv.first); pb::put_string(&e, 2, kv.second);
            pb::put_bytes(&o, 1, e);
        }
        return o;
    }
};
inline std::string encode_allocate_response(const std::vector<ContainerAllocateResponse>& rs) {
    std::string o;
    for (const auto& r : rs) pb::put_bytes(&o, 1, r.encode());
    return o;
}

struct ContainerPreferredAllocationRequest {
    std::vector<std::string> available, must_include;
    int32_t allocation_size = 0;
};
inline bool decode_preferred_request(const std::string& msg, std::vector<ContainerPreferredAllocationRequest>* out) {
    std::vector<std::string> containers;
    if (!repeated_strings(msg, 1, &containers)) return false;
    for (const std::string& c : containers) {
        ContainerPreferredAllocationRequest r;
        std::map<int, int64_t> ints;
        if (!repeated_strings(c, 1, &r.available, &ints) || !repeated_strings(c, 2, &r.must_include)) return false;
        r.allocation_size = ints.count(3) ? (int32_t)ints[3] : 0;
        out->push_back(std::move(r));
    }
    return true;
}
inline std::string encode_preferred_response(const std::vector<std::vector<std::string>>& per_container) {
    std::string o;
    for (const auto& ids : per_container) {
        std::string c;
        for (const auto& id : ids) pb::put_bytes(&c, 1, id);
        pb::put_bytes(&o, 1, c);
    }
    return o;
}

}  // namespace v1beta1
