// HPACK (RFC 7541) for the device-plugin host's HTTP/2 transport.
//   Decoder: complete (static + dynamic table, table-size updates, Huffman strings) — gRPC peers
//            (grpc-go in kubelet, grpc C-core in the tests) index and Huffman-code freely.
//   Encoder: literal-without-indexing, raw strings.  Always legal, needs no peer state.
#pragma once
#include <cstdint>
#include <deque>
#include <string>
#include <utility>
#include <vector>

namespace hpack {

using Header = std::pair<std::string, std::string>;

struct HuffSym { uint32_t code; uint8_t len; };
static const HuffSym kHuff[257] = {
#include "hpack_huffman.inc"
};

// Binary trie over the code, built once: node = {child0, child1}; leaf = -(symbol + 1).
class HuffTree {
public:
    HuffTree() {
        nodes_.push_back({0, 0});
        for (int s = 0; s < 257; ++s) {
            int cur = 0;
            for (int b = kHuff[s].len - 1; b >= 0; --b) {
                const int bit = (kHuff[s].code >> b) & 1;
                if (b == 0) { child(cur, bit) = -(s + 1); break; }
                if (child(cur, bit) == 0) {
                    const int fresh = (int)nodes_.size();
                    nodes_.push_back({0, 0});
                    child(cur, bit) = fresh;
                }
                cur = child(cur, bit);
            }
        }
    }
    // RFC 7541 §5.2: padding is < 8 bits, all ones (a prefix of EOS); EOS inside the string is an error.
    bool decode(const uint8_t* p, size_t n, std::string* out) const {
        int cur = 0, pad_bits = 0;
        bool pad_ones = true;
        for (size_t i = 0; i < n; ++i) {
            for (int b = 7; b >= 0; --b) {
                const int bit = (p[i] >> b) & 1;
                const int next = bit ? nodes_[cur].c1 : nodes_[cur].c0;
                if (next == 0) return false;
                if (next < 0) {
                    const int sym = -next - 1;
                    if (sym == 256) return false;
                    out->push_back((char)sym);
                    cur = 0; pad_bits = 0; pad_ones = true;
                } else {
                    cur = next; ++pad_bits; pad_ones = pad_ones && bit;
                }
            }
        }
        return pad_bits < 8 && pad_ones;
    }
private:
    struct Node { int c0, c1; };
    int& child(int node, int bit) { return bit ? nodes_[node].c1 : nodes_[node].c0; }
    std::vector<Node> nodes_;
};

inline const HuffTree& huff_tree() { static const HuffTree t; return t; }

static const Header kStatic[61] = {
    {":authority", ""}, {":method", "GET"}, {":method", "POST"}, {":path", "/"}, {":path", "/index.html"}, {":scheme", "http"},
    {":scheme", "https"}, {":status", "200"}, {":status", "204"}, {":status", "206"}, {":status", "304"}, {":status", "400"},
    {":status", "404"}, {":status", "500"}, {"accept-charset", ""}, {"accept-encoding", "gzip, deflate"}, {"accept-language", ""},
    {"accept-ranges", ""}, {"accept", ""}, {"access-control-allow-origin", ""}, {"age", ""}, {"allow", ""}, {"authorization", ""},
    {"cache-control", ""}, {"content-disposition", ""}, {"content-encoding", ""}, {"content-language", ""}, {"content-length", ""},
    {"content-location", ""}, {"content-range", ""}, {"content-type", ""}, {"cookie", ""}, {"date", ""}, {"etag", ""}, {"expect", ""},
    {"expires", ""}, {"from", ""}, {"host", ""}, {"if-match", ""}, {"if-modified-since", ""}, {"if-none-match", ""}, {"if-range", ""},
    {"if-unmodified-since", ""}, {"last-modified", ""}, {"link", ""}, {"location", ""}, {"max-forwards", ""}, {"proxy-authenticate", ""},
    {"proxy-authorization", ""}, {"range", ""}, {"referer", ""}, {"refresh", ""}, {"retry-after", ""}, {"server", ""}, {"set-cookie", ""},
    {"strict-transport-security", ""}, {"transfer-encoding", ""}, {"user-agent", ""}, {"vary", ""}, {"via", ""}, {"www-authenticate", ""},
};

class Decoder {
public:
    // max_size_limit = the SETTINGS_HEADER_TABLE_SIZE we advertised (default 4096)
    explicit Decoder(size_t max_size_limit = 4096) : limit_(max_size_limit), max_size_(max_size_limit) {}

    bool decode(const uint8_t* p, size_t n, std::vector<Header>* out) {
        size_t i = 0;
        while (i < n) {
            const uint8_t b = p[i];
            if (b & 0x80) {                                   // indexed header field
                uint64_t idx;
                if (!read_int(p, n, &i, 7, &idx) || idx == 0) return false;
                Header h;
                if (!lookup(idx, &h)) return false;
                out->push_back(std::move(h));
            } else if (b & 0x40) {                            // literal with incremental indexing
                Header h;
                if (!read_literal(p, n, &i, 6, &h)) return false;
                insert(h);
                out->push_back(std::move(h));
            } else if (b & 0x20) {                            // dynamic table size update
                uint64_t sz;
                if (!read_int(p, n, &i, 5, &sz) || sz > limit_) return false;
                max_size_ = (size_t)sz;
                evict();
            } else {                                          // literal without indexing / never indexed
                Header h;
                if (!read_literal(p, n, &i, 4, &h)) return false;
                out->push_back(std::move(h));
            }
        }
        return true;
    }

private:
    static bool read_int(const uint8_t* p, size_t n, size_t* i, int prefix, uint64_t* v) {
        if (*i >= n) return false;
        const uint32_t mask = (1u << prefix) - 1;
        uint64_t x = p[(*i)++] & mask;
        if (x == mask) {
            int shift = 0;
            for (;;) {
                if (*i >= n || shift > 56) return false;
                const uint8_t b = p[(*i)++];
                x += (uint64_t)(b & 0x7f) << shift;
                shift += 7;
                if (!(b & 0x80)) break;
            }
        }
        *v = x;
        return true;
    }
    static bool read_string(const uint8_t* p, size_t n, size_t* i, std::string* s) {
        if (*i >= n) return false;
        const bool huff = p[*i] & 0x80;
        uint64_t len;
        if (!read_int(p, n, i, 7, &len) || len > n - *i) return false;
        if (huff) { if (!huff_tree().decode(p + *i, (size_t)len, s)) return false; }
        else s->assign((const char*)p + *i, (size_t)len);
        *i += (size_t)len;
        return true;
    }
    bool read_literal(const uint8_t* p, size_t n, size_t* i, int prefix, Header* h) {
        uint64_t idx;
        if (!read_int(p, n, i, prefix, &idx)) return false;
        if (idx) { Header ref; if (!lookup(idx, &ref)) return false; h->first = ref.first; }
        else if (!read_string(p, n, i, &h->first)) return false;
        return read_string(p, n, i, &h->second);
    }
    bool lookup(uint64_t idx, Header* h) const {
        if (idx >= 1 && idx <= 61) { *h = kStatic[idx - 1]; return true; }
        const uint64_t d = idx - 62;
        if (d >= dyn_.size()) return false;
        *h = dyn_[(size_t)d];
        return true;
    }
    void insert(const Header& h) {
        const size_t sz = h.first.size() + h.second.size() + 32;
        dyn_.push_front(h);
        size_ += sz;
        evict();
    }
    void evict() {
        while (size_ > max_size_ && !dyn_.empty()) {
            size_ -= dyn_.back().first.size() + dyn_.back().second.size() + 32;
            dyn_.pop_back();
        }
    }
    size_t limit_, max_size_, size_ = 0;
    std::deque<Header> dyn_;
};

inline void encode_int(std::string* out, uint64_t v, int prefix, uint8_t flags) {
    const uint32_t mask = (1u << prefix) - 1;
    if (v < mask) { out->push_back((char)(flags | v)); return; }
    out->push_back((char)(flags | mask));
    v -= mask;
    while (v >= 128) { out->push_back((char)((v & 0x7f) | 0x80)); v >>= 7; }
    out->push_back((char)v);
}

inline void encode(std::string* out, const std::vector<Header>& hs) {
    for (const Header& h : hs) {
        out->push_back(0x00);                              // literal header field without indexing, new name
        encode_int(out, h.first.size(), 7, 0x00);
        out->append(h.first);
        encode_int(out, h.second.size(), 7, 0x00);
        out->append(h.second);
    }
}

}  // namespace hpack
