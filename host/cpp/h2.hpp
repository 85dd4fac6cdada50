// Minimal HTTP/2 (RFC 9113) + gRPC framing over a unix-domain socket: exactly what the kubelet
// device-plugin API needs (SURVEY.md §8b outer boundary) and nothing else.  No TLS, no push, no priorities.
//
//   Server side: one reader thread per connection; every RPC is answered on its own worker thread (the
//   reader must stay free for WINDOW_UPDATE frames).  Send-side flow control is honoured (connection + stream windows);
//   received DATA is credited back immediately (messages on this API are a few KB).
//   Client side: one blocking unary call per connection (Registration.Register).
#pragma once
#include <errno.h>
#include <poll.h>
#include <sys/socket.h>
#include <sys/un.h>
#include <unistd.h>

#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdint>
#include <cstring>
#include <functional>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "hpack.hpp"

namespace h2 {

enum FrameType : uint8_t { DATA = 0, HEADERS = 1, PRIORITY = 2, RST_STREAM = 3, SETTINGS = 4, PUSH_PROMISE = 5, PING = 6, GOAWAY = 7, WINDOW_UPDATE = 8, CONTINUATION = 9 };
enum Flags : uint8_t { END_STREAM = 0x1, ACK = 0x1, END_HEADERS = 0x4, PADDED = 0x8, PRIORITY_FLAG = 0x20 };
enum Settings : uint16_t { HEADER_TABLE_SIZE = 1, ENABLE_PUSH = 2, MAX_CONCURRENT_STREAMS = 3, INITIAL_WINDOW_SIZE = 4, MAX_FRAME_SIZE = 5, MAX_HEADER_LIST_SIZE = 6 };

// gRPC status codes used by this host
enum GrpcCode { OK = 0, CANCELLED = 1, UNKNOWN = 2, INVALID_ARGUMENT = 3, DEADLINE_EXCEEDED = 4, UNIMPLEMENTED = 12, INTERNAL = 13, UNAVAILABLE = 14 };

struct Status {
    int code = OK;
    std::string message;
    bool ok() const { return code == OK; }
};

static const char kPreface[] = "PRI * HTTP/2.0\r\n\r\nSM\r\n\r\n";

inline bool write_all(int fd, const void* buf, size_t n) {
    const char* p = (const char*)buf;
    while (n) {
        ssize_t w = ::send(fd, p, n, MSG_NOSIGNAL);
        if (w < 0) { if (errno == EINTR) continue; return false; }
        p += w; n -= (size_t)w;
    }
    return true;
}
// deadline_ms < 0: block.  Returns false on EOF / error / timeout.
inline bool read_all(int fd, void* buf, size_t n, int deadline_ms = -1) {
    char* p = (char*)buf;
    while (n) {
        if (deadline_ms >= 0) {
            struct pollfd pf = {fd, POLLIN, 0};
            int r = ::poll(&pf, 1, deadline_ms);
            if (r == 0) return false;
            if (r < 0) { if (errno == EINTR) continue; return false; }
        }
        ssize_t r = ::recv(fd, p, n, 0);
        if (r == 0) return false;
        if (r < 0) { if (errno == EINTR) continue; return false; }
        p += r; n -= (size_t)r;
    }
    return true;
}

struct Frame {
    uint8_t type = 0, flags = 0;
    uint32_t stream = 0;
    std::string payload;
    bool oversize = false;
};

inline std::string frame_bytes(uint8_t type, uint8_t flags, uint32_t stream, const std::string& payload) {
    std::string f;
    f.resize(9);
    const uint32_t n = (uint32_t)payload.size();
    f[0] = (char)(n >> 16); f[1] = (char)(n >> 8); f[2] = (char)n;
    f[3] = (char)type; f[4] = (char)flags;
    f[5] = (char)((stream >> 24) & 0x7f); f[6] = (char)(stream >> 16); f[7] = (char)(stream >> 8); f[8] = (char)stream;
    f += payload;
    return f;
}
constexpr uint32_t kMaxFrameSize = 16384;          // RFC 9113 §4.2: the initial (and our only) SETTINGS_MAX_FRAME_SIZE
inline bool read_frame(int fd, Frame* f, int deadline_ms = -1) {
    uint8_t h[9];
    if (!read_all(fd, h, 9, deadline_ms)) return false;
    const uint32_t n = ((uint32_t)h[0] << 16) | ((uint32_t)h[1] << 8) | h[2];
    f->type = h[3]; f->flags = h[4];
    f->oversize = n > kMaxFrameSize;      // we never raise SETTINGS_MAX_FRAME_SIZE: a larger frame is a FRAME_SIZE_ERROR, and is not read
    if (f->oversize) return false;
    f->stream = (((uint32_t)h[5] & 0x7f) << 24) | ((uint32_t)h[6] << 16) | ((uint32_t)h[7] << 8) | h[8];
    f->payload.resize(n);
    return n == 0 || read_all(fd, &f->payload[0], n, deadline_ms);
}
inline std::string u32be(uint32_t v) { char b[4] = {(char)(v >> 24), (char)(v >> 16), (char)(v >> 8), (char)v}; return std::string(b, 4); }
inline uint32_t rd32(const std::string& s, size_t off) {
    return ((uint32_t)(uint8_t)s[off] << 24) | ((uint32_t)(uint8_t)s[off + 1] << 16) | ((uint32_t)(uint8_t)s[off + 2] << 8) | (uint8_t)s[off + 3];
}

// gRPC length-prefixed message: 1 byte compressed flag, 4 bytes big-endian length
inline std::string grpc_frame(const std::string& msg) { std::string f(1, '\0'); f += u32be((uint32_t)msg.size()); f += msg; return f; }
inline bool grpc_unframe(const std::string& body, std::vector<std::string>* msgs) {
    size_t i = 0;
    while (i < body.size()) {
        if (body.size() - i < 5 || body[i] != 0) return false;     // compressed messages are never negotiated
        const uint32_t n = rd32(body, i + 1);
        if (body.size() - i - 5 < n) return false;
        msgs->emplace_back(body, i + 5, n);
        i += 5 + n;
    }
    return true;
}
// grpc-message is percent-encoded (gRPC over HTTP/2 spec)
inline std::string pct_encode(const std::string& s) {
    static const char* hex = "0123456789ABCDEF";
    std::string o;
    for (unsigned char c : s) {
        if (c >= 0x20 && c <= 0x7e && c != '%') o.push_back((char)c);
        else { o.push_back('%'); o.push_back(hex[c >> 4]); o.push_back(hex[c & 15]); }
    }
    return o;
}
inline std::string pct_decode(const std::string& s) {
    std::string o;
    for (size_t i = 0; i < s.size(); ++i) {
        if (s[i] == '%' && i + 2 < s.size() + 0 && isxdigit((unsigned char)s[i + 1]) && isxdigit((unsigned char)s[i + 2])) {
            o.push_back((char)strtol(s.substr(i + 1, 2).c_str(), nullptr, 16));
            i += 2;
        } else o.push_back(s[i]);
    }
    return o;
}

class Connection;

// What a handler sees of one RPC.
class ServerCall {
public:
    ServerCall(Connection* c, uint32_t id, std::string path) : conn_(c), id_(id), path_(std::move(path)) {}
    const std::string& path() const { return path_; }
    const std::string& request() const { return request_; }
    bool cancelled() const { return cancelled_.load(); }
    // server-streaming: one response message (response headers go out with the first one)
    bool send(const std::string& msg);
    // end of the RPC (trailers).  Unary: send(msg) then finish(OK), or finish(error) alone.
    void finish(const Status& st);

private:
    friend class Connection;
    Connection* conn_;
    uint32_t id_;
    std::string path_, request_, header_block_;
    bool headers_done_ = false, request_done_ = false;
    std::atomic<bool> dispatched_{false};        // handed to a worker thread: the reader must not touch request_ / header state any more
    bool headers_sent_ = false, finished_ = false;
    std::atomic<bool> cancelled_{false};
    int64_t send_window_ = 65535;
};

using Handler = std::function<void(std::shared_ptr<ServerCall>)>;
struct Route { Handler fn; bool streaming; };

class Connection {
public:
    static constexpr size_t kMaxConcurrentStreams = 100;     // advertised in SETTINGS and enforced
    static constexpr size_t kMaxHeaderBlock = 64 * 1024;     // gRPC request headers are a few hundred bytes
    Connection(int fd, const std::map<std::string, Route>* routes) : fd_(fd), routes_(routes) {}
    // The descriptor is closed only here, after the reader thread was joined by the owner and the workers by
    // join_streams(): closing a descriptor another thread is blocked on is a race (and a reuse hazard).
    ~Connection() {
        close_fd();
        join_streams();
        const int fd = fd_.exchange(-1);
        if (fd < 0) return;
        // Closing a unix socket with unread bytes resets the peer and drops what we queued for it (a GOAWAY it has
        // not read yet): swallow what is left, briefly, before closing.
        char sink[4096];
        for (int spins = 0; spins < 10; ++spins) {
            struct pollfd pf = {fd, POLLIN, 0};
            if (::poll(&pf, 1, 20) <= 0) break;
            if (::recv(fd, sink, sizeof(sink), MSG_DONTWAIT) <= 0) break;
        }
        ::close(fd);
    }

    // Any thread: wake everything that waits on this connection and make further I/O fail.  Does not close.
    void close_fd() {
        if (!shut_.exchange(true)) { const int fd = fd_.load(); if (fd >= 0) ::shutdown(fd, SHUT_RDWR); }
        { std::lock_guard<std::mutex> l(mu_); dead_ = true; }
        cv_.notify_all();
    }

    // Server loop: returns when the peer goes away or violates the protocol.
    void serve() {
        char pre[24];
        int fd = fd_.load();
        if (fd < 0) return;
        std::string settings;                                           // SETTINGS: MAX_CONCURRENT_STREAMS, MAX_HEADER_LIST_SIZE
        settings += (char)0; settings += (char)MAX_CONCURRENT_STREAMS; settings += u32be((uint32_t)kMaxConcurrentStreams);
        settings += (char)0; settings += (char)MAX_HEADER_LIST_SIZE; settings += u32be((uint32_t)kMaxHeaderBlock);
        if (!write_frame(SETTINGS, 0, 0, settings)) return;
        if (!read_all(fd, pre, 24) || memcmp(pre, kPreface, 24) != 0) return;
        Frame f;
        uint32_t continuing = 0;
        while (!shut_.load() && read_frame(fd, &f)) {
            if (continuing && (f.type != CONTINUATION || f.stream != continuing)) { goaway(1); break; }
            switch (f.type) {
                case SETTINGS:
                    if (f.flags & ACK) break;
                    if (f.payload.size() % 6) { goaway(6); return; }
                    apply_settings(f.payload);
                    write_frame(SETTINGS, ACK, 0, "");
                    break;
                case PING:
                    if (!(f.flags & ACK)) write_frame(PING, ACK, 0, f.payload);
                    break;
                case WINDOW_UPDATE: {
                    if (f.payload.size() != 4) { goaway(6); return; }
                    const uint32_t inc = rd32(f.payload, 0) & 0x7fffffff;
                    std::lock_guard<std::mutex> l(mu_);
                    if (f.stream == 0) conn_window_ += inc;
                    else { auto it = calls_.find(f.stream); if (it != calls_.end()) it->second->send_window_ += inc; }
                    cv_.notify_all();
                    break;
                }
                case HEADERS: {
                    if (f.stream == 0 || !(f.stream & 1)) { goaway(1); return; }
                    size_t off = 0, pad = 0;
                    if (f.flags & PADDED) { if (f.payload.empty()) { goaway(1); return; } pad = (uint8_t)f.payload[0]; off = 1; }
                    if (f.flags & PRIORITY_FLAG) off += 5;
                    if (off + pad > f.payload.size()) { goaway(1); return; }
                    std::shared_ptr<ServerCall> call;
                    bool refused = false;
                    {
                        std::lock_guard<std::mutex> l(mu_);
                        auto it = calls_.find(f.stream);
                        if (it == calls_.end() && calls_.size() >= kMaxConcurrentStreams) refused = true;
                        else if (it == calls_.end()) {
                            call = std::make_shared<ServerCall>(this, f.stream, "");
                            call->send_window_ = peer_initial_window_;
                            calls_[f.stream] = call;
                        } else call = it->second;                         // a second header block on an open stream
                    }
                    // gRPC clients send no trailers.  HEADERS on a stream whose request is complete (its worker may be reading
                    // request_ and header state right now) is a protocol error; the block cannot simply be dropped either — it
                    // would have to pass through the HPACK decoder — so the connection goes.
                    if (call && (call->request_done_ || call->dispatched_.load())) { goaway(1); return; }
                    if (refused) {
                        // the block still has to pass through the HPACK decoder (it may update the dynamic table);
                        // simplest correct answer for a peer that ignores our SETTINGS is to drop the connection
                        goaway(7);                                            // REFUSED_STREAM
                        return;
                    }
                    call->header_block_.append(f.payload, off, f.payload.size() - off - pad);
                    if (call->header_block_.size() > kMaxHeaderBlock) { goaway(11); return; }   // ENHANCE_YOUR_CALM
                    if (f.flags & END_STREAM) call->request_done_ = true;
                    if (f.flags & END_HEADERS) { if (!headers_complete(call)) return; }
                    else continuing = f.stream;
                    break;
                }
                case CONTINUATION: {
                    std::shared_ptr<ServerCall> call = find(f.stream);
                    if (!call || continuing != f.stream) { goaway(1); return; }
                    call->header_block_ += f.payload;
                    if (call->header_block_.size() > kMaxHeaderBlock) { goaway(11); return; }
                    if (f.flags & END_HEADERS) { continuing = 0; if (!headers_complete(call)) return; }
                    break;
                }
                case DATA: {
                    size_t off = 0, pad = 0;
                    if (f.flags & PADDED) { if (f.payload.empty()) { goaway(1); return; } pad = (uint8_t)f.payload[0]; off = 1; }
                    if (off + pad > f.payload.size()) { goaway(1); return; }
                    std::shared_ptr<ServerCall> call = find(f.stream);
                    if (!f.payload.empty()) {                                // credit the bytes back at once: connection, then stream
                        write_frame(WINDOW_UPDATE, 0, 0, u32be((uint32_t)f.payload.size()));
                        if (call && !(f.flags & END_STREAM)) write_frame(WINDOW_UPDATE, 0, f.stream, u32be((uint32_t)f.payload.size()));
                    }
                    if (!call) break;                                         // stream already reset/finished
                    if (call->request_done_ || call->dispatched_.load()) {    // DATA after END_STREAM: the worker owns request_ now
                        reset(f.stream, 5);                                   // STREAM_CLOSED
                        break;
                    }
                    call->request_.append(f.payload, off, f.payload.size() - off - pad);
                    if (call->request_.size() > (4u << 20)) { reset(f.stream, 11); drop(f.stream); break; }
                    if (f.flags & END_STREAM) { call->request_done_ = true; dispatch(call); }
                    break;
                }
                case RST_STREAM: {
                    std::shared_ptr<ServerCall> call = find(f.stream);
                    if (call) { call->cancelled_ = true; cv_.notify_all(); drop(f.stream); }
                    break;
                }
                case GOAWAY:
                    break;                                                    // the peer will close when its streams are done
                case PRIORITY:
                    break;
                case PUSH_PROMISE:
                    goaway(1);
                    return;
                default:
                    break;                                                    // unknown frame types are ignored (RFC 9113 §4.1)
            }
        }
        if (f.oversize) goaway(6);                                            // FRAME_SIZE_ERROR
        cancel_all();
    }

    bool write_frame(uint8_t type, uint8_t flags, uint32_t stream, const std::string& payload) {
        const std::string bytes = frame_bytes(type, flags, stream, payload);
        std::lock_guard<std::mutex> l(wmu_);
        const int fd = fd_.load();
        return fd >= 0 && !shut_.load() && write_all(fd, bytes.data(), bytes.size());
    }

    // DATA under flow control; blocks while the peer's windows are closed.  false = connection or stream gone.
    bool send_data(ServerCall* call, const std::string& data, bool end_stream) {
        size_t off = 0;
        do {
            size_t n;
            {
                std::unique_lock<std::mutex> l(mu_);
                cv_.wait(l, [&] { return dead_ || call->cancelled() || data.size() == off || (conn_window_ > 0 && call->send_window_ > 0); });
                if (dead_ || call->cancelled()) return false;
                n = std::min<size_t>(data.size() - off, std::min<int64_t>(std::min<int64_t>(conn_window_, call->send_window_), peer_max_frame_));
                conn_window_ -= (int64_t)n;
                call->send_window_ -= (int64_t)n;
            }
            const bool last = off + n == data.size();
            if (!write_frame(DATA, (last && end_stream) ? END_STREAM : 0, call->id_, data.substr(off, n))) return false;
            off += n;
        } while (off < data.size());
        return true;
    }
    bool send_headers(uint32_t stream, const std::vector<hpack::Header>& hs, bool end_stream) {
        std::string block;
        hpack::encode(&block, hs);
        return write_frame(HEADERS, END_HEADERS | (end_stream ? END_STREAM : 0), stream, block);   // our header blocks are far below 16 KiB
    }
    void drop(uint32_t stream) { std::lock_guard<std::mutex> l(mu_); calls_.erase(stream); }
    bool finished() const { return finished_.load(); }
    void mark_finished() { finished_ = true; }
    void reset(uint32_t stream, uint32_t code) { write_frame(RST_STREAM, 0, stream, u32be(code)); }

private:
    std::shared_ptr<ServerCall> find(uint32_t stream) {
        std::lock_guard<std::mutex> l(mu_);
        auto it = calls_.find(stream);
        return it == calls_.end() ? nullptr : it->second;
    }
    void apply_settings(const std::string& p) {
        std::lock_guard<std::mutex> l(mu_);
        for (size_t i = 0; i + 6 <= p.size(); i += 6) {
            const uint16_t id = (uint16_t)(((uint8_t)p[i] << 8) | (uint8_t)p[i + 1]);
            const uint32_t v = rd32(p, i + 2);
            if (id == INITIAL_WINDOW_SIZE) {
                const int64_t delta = (int64_t)v - peer_initial_window_;
                peer_initial_window_ = v;
                for (auto& kv : calls_) kv.second->send_window_ += delta;
            } else if (id == MAX_FRAME_SIZE) {
                if (v >= 16384 && v <= 16777215) peer_max_frame_ = v;
            }
        }
        cv_.notify_all();
    }
    bool headers_complete(const std::shared_ptr<ServerCall>& call) {
        std::vector<hpack::Header> hs;
        if (!dec_.decode((const uint8_t*)call->header_block_.data(), call->header_block_.size(), &hs)) { goaway(9); return false; }   // COMPRESSION_ERROR
        call->header_block_.clear();
        if (call->headers_done_) return true;                       // request trailers: ignored
        call->headers_done_ = true;
        std::string method, ctype;
        for (auto& h : hs) {
            if (h.first == ":path") call->path_ = h.second;
            else if (h.first == ":method") method = h.second;
            else if (h.first == "content-type") ctype = h.second;
        }
        if (method != "POST" || ctype.compare(0, 16, "application/grpc") != 0) {
            send_headers(call->id_, {{":status", ctype.compare(0, 16, "application/grpc") ? "415" : "405"}}, true);
            drop(call->id_);
            return true;
        }
        if (call->request_done_) dispatch(call);                    // no body at all: still a (malformed) request
        return true;
    }
    void dispatch(const std::shared_ptr<ServerCall>& call) {
        if (call->dispatched_.exchange(true)) return;                      // one worker per call, whatever the peer sends afterwards
        auto it = routes_->find(call->path_);
        if (it == routes_->end()) { call->finish({UNIMPLEMENTED, "unknown method " + call->path_}); return; }
        std::vector<std::string> msgs;
        if (!grpc_unframe(call->request_, &msgs) || msgs.size() != 1) { call->finish({INTERNAL, "malformed gRPC request body"}); return; }
        call->request_ = msgs[0];
        // Every RPC runs on its own thread: the reader must stay free to process WINDOW_UPDATE frames, or a
        // response larger than the peer's flow-control window would wait on an update nobody reads.
        Handler fn = it->second.fn;
        auto done = std::make_shared<std::atomic<bool>>(false);
        std::lock_guard<std::mutex> l(mu_);
        for (size_t i = 0; i < workers_.size();) {                       // reap finished workers
            if (workers_[i].done->load()) { workers_[i].t.join(); workers_.erase(workers_.begin() + (long)i); }
            else ++i;
        }
        workers_.push_back(Worker{std::thread([fn, call, done] { fn(call); done->store(true); }), done});
    }
    void goaway(uint32_t code) {
        write_frame(GOAWAY, 0, 0, u32be(0x7fffffff) + u32be(code));
        close_fd();
    }
    void cancel_all() {
        std::lock_guard<std::mutex> l(mu_);
        dead_ = true;
        for (auto& kv : calls_) kv.second->cancelled_ = true;
        cv_.notify_all();
    }
    void join_streams() {
        cancel_all();
        std::vector<Worker> ws;
        { std::lock_guard<std::mutex> l(mu_); ws.swap(workers_); }
        for (auto& w : ws) if (w.t.joinable()) w.t.join();
    }

    std::atomic<int> fd_;
    std::atomic<bool> shut_{false};
    std::atomic<bool> finished_{false};
    const std::map<std::string, Route>* routes_;
    std::mutex mu_, wmu_;
    std::condition_variable cv_;
    bool dead_ = false;
    int64_t conn_window_ = 65535;
    int64_t peer_initial_window_ = 65535;
    int64_t peer_max_frame_ = 16384;
    hpack::Decoder dec_;
    std::map<uint32_t, std::shared_ptr<ServerCall>> calls_;
    struct Worker { std::thread t; std::shared_ptr<std::atomic<bool>> done; };
    std::vector<Worker> workers_;
};

inline bool ServerCall::send(const std::string& msg) {
    if (finished_ || cancelled()) return false;
    if (!headers_sent_) {
        headers_sent_ = true;
        if (!conn_->send_headers(id_, {{":status", "200"}, {"content-type", "application/grpc"}}, false)) return false;
    }
    return conn_->send_data(this, grpc_frame(msg), false);
}
inline void ServerCall::finish(const Status& st) {
    if (finished_) return;
    finished_ = true;
    std::vector<hpack::Header> hs;
    if (!headers_sent_) { hs.push_back({":status", "200"}); hs.push_back({"content-type", "application/grpc"}); }   // Trailers-Only
    hs.push_back({"grpc-status", std::to_string(st.code)});
    if (!st.message.empty()) hs.push_back({"grpc-message", pct_encode(st.message)});
    if (!cancelled()) conn_->send_headers(id_, hs, true);
    conn_->drop(id_);
}

// ---- server: accept loop over a unix socket ----------------------------------------------------------
class Server {
public:
    ~Server() { stop(); }
    void route(const std::string& path, Handler fn, bool streaming) { routes_[path] = Route{std::move(fn), streaming}; }

    bool listen_unix(const std::string& path, std::string* err) {
        ::unlink(path.c_str());
        int fd = ::socket(AF_UNIX, SOCK_STREAM | SOCK_CLOEXEC, 0);
        if (fd < 0) { *err = std::string("socket: ") + strerror(errno); return false; }
        struct sockaddr_un sa;
        memset(&sa, 0, sizeof(sa));
        sa.sun_family = AF_UNIX;
        if (path.size() >= sizeof(sa.sun_path)) { ::close(fd); *err = "socket path too long: " + path; return false; }
        memcpy(sa.sun_path, path.data(), path.size());
        if (::bind(fd, (struct sockaddr*)&sa, sizeof(sa)) < 0 || ::listen(fd, 16) < 0) {
            *err = "bind/listen " + path + ": " + strerror(errno);
            ::close(fd);
            return false;
        }
        lfd_ = fd;
        path_ = path;
        stopping_ = false;
        acceptor_ = std::thread([this] { accept_loop(); });
        return true;
    }
    void stop() {
        stopping_ = true;
        const int lfd = lfd_.load();
        if (lfd >= 0) ::shutdown(lfd, SHUT_RDWR);           // wakes the acceptor's poll; closed after it has exited
        if (acceptor_.joinable()) acceptor_.join();
        if (lfd_.exchange(-1) >= 0) ::close(lfd);
        std::vector<std::shared_ptr<Connection>> conns;
        std::vector<std::thread> ts;
        { std::lock_guard<std::mutex> l(mu_); conns.swap(conns_); ts.swap(threads_); }
        for (auto& c : conns) c->close_fd();
        for (auto& t : ts) if (t.joinable()) t.join();
        conns.clear();
        if (!path_.empty()) { ::unlink(path_.c_str()); path_.clear(); }
    }

private:
    void accept_loop() {
        for (;;) {
            const int lfd = lfd_.load();
            if (lfd < 0) return;
            struct pollfd pf = {lfd, POLLIN, 0};
            const int r = ::poll(&pf, 1, 200);
            if (stopping_) return;
            if (r <= 0) continue;
            const int fd = ::accept4(lfd, nullptr, nullptr, SOCK_CLOEXEC);
            if (fd < 0) { if (errno == EINTR || errno == EAGAIN) continue; return; }
            auto conn = std::make_shared<Connection>(fd, &routes_);
            std::lock_guard<std::mutex> l(mu_);
            for (size_t i = 0; i < conns_.size();) {                 // reap connections whose peer went away
                if (conns_[i]->finished()) {
                    if (threads_[i].joinable()) threads_[i].join();
                    conns_.erase(conns_.begin() + i);
                    threads_.erase(threads_.begin() + i);
                } else ++i;
            }
            conns_.push_back(conn);
            threads_.emplace_back([conn] { conn->serve(); conn->close_fd(); conn->mark_finished(); });
        }
    }
    std::map<std::string, Route> routes_;
    std::atomic<int> lfd_{-1};
    std::atomic<bool> stopping_{false};
    std::string path_;
    std::thread acceptor_;
    std::mutex mu_;
    std::vector<std::shared_ptr<Connection>> conns_;
    std::vector<std::thread> threads_;
};

// ---- client: one unary call over a fresh connection (Registration.Register) -----------------------------
inline Status unary_call(const std::string& socket_path, const std::string& path, const std::string& request, std::string* response, int timeout_ms) {
    int fd = ::socket(AF_UNIX, SOCK_STREAM | SOCK_CLOEXEC, 0);
    if (fd < 0) return {UNAVAILABLE, std::string("socket: ") + strerror(errno)};
    struct sockaddr_un sa;
    memset(&sa, 0, sizeof(sa));
    sa.sun_family = AF_UNIX;
    if (socket_path.size() >= sizeof(sa.sun_path)) { ::close(fd); return {UNAVAILABLE, "socket path too long"}; }
    memcpy(sa.sun_path, socket_path.data(), socket_path.size());
    if (::connect(fd, (struct sockaddr*)&sa, sizeof(sa)) < 0) {
        Status st{UNAVAILABLE, "connect " + socket_path + ": " + strerror(errno)};
        ::close(fd);
        return st;
    }
    struct Closer { int fd; ~Closer() { ::close(fd); } } closer{fd};
    std::string out(kPreface, 24);
    std::string settings;                                            // ENABLE_PUSH = 0
    settings += (char)0; settings += (char)ENABLE_PUSH; settings += u32be(0);
    out += frame_bytes(SETTINGS, 0, 0, settings);
    std::string block;
    hpack::encode(&block, {{":method", "POST"}, {":scheme", "http"}, {":path", path}, {":authority", "localhost"},
                           {"content-type", "application/grpc"}, {"te", "trailers"}, {"user-agent", "b200-device-plugin/1"},
                           {"grpc-timeout", std::to_string(timeout_ms) + "m"}});
    out += frame_bytes(HEADERS, END_HEADERS, 1, block);
    out += frame_bytes(DATA, END_STREAM, 1, grpc_frame(request));     // requests on this API are < 200 bytes: inside every initial window
    if (!write_all(fd, out.data(), out.size())) return {UNAVAILABLE, "write to " + socket_path + " failed"};

    hpack::Decoder dec;
    std::string body, hblock;
    std::vector<hpack::Header> all;
    bool ended = false, end_after_headers = false;
    const auto t_end = std::chrono::steady_clock::now() + std::chrono::milliseconds(timeout_ms);
    Frame f;
    while (!ended) {
        const int left = (int)std::chrono::duration_cast<std::chrono::milliseconds>(t_end - std::chrono::steady_clock::now()).count();
        if (left <= 0) return {DEADLINE_EXCEEDED, "deadline exceeded waiting for " + path};
        if (!read_frame(fd, &f, left)) return {UNAVAILABLE, "connection to " + socket_path + " closed before the response completed"};
        switch (f.type) {
            case SETTINGS:
                if (!(f.flags & ACK)) { const std::string ack = frame_bytes(SETTINGS, ACK, 0, ""); write_all(fd, ack.data(), ack.size()); }
                break;
            case PING:
                if (!(f.flags & ACK)) { const std::string ack = frame_bytes(PING, ACK, 0, f.payload); write_all(fd, ack.data(), ack.size()); }
                break;
            case HEADERS:
            case CONTINUATION: {
                size_t off = 0, pad = 0;
                if (f.type == HEADERS) {
                    if (f.flags & PADDED) { if (f.payload.empty()) return {INTERNAL, "bad padding"}; pad = (uint8_t)f.payload[0]; off = 1; }
                    if (f.flags & PRIORITY_FLAG) off += 5;
                    if (f.flags & END_STREAM) end_after_headers = true;
                }
                if (off + pad > f.payload.size()) return {INTERNAL, "bad HEADERS frame"};
                hblock.append(f.payload, off, f.payload.size() - off - pad);
                if (f.flags & END_HEADERS) {
                    if (!dec.decode((const uint8_t*)hblock.data(), hblock.size(), &all)) return {INTERNAL, "HPACK decoding failed"};
                    hblock.clear();
                    if (end_after_headers) ended = true;
                }
                break;
            }
            case DATA: {
                size_t off = 0, pad = 0;
                if (f.flags & PADDED) { if (f.payload.empty()) return {INTERNAL, "bad padding"}; pad = (uint8_t)f.payload[0]; off = 1; }
                if (off + pad > f.payload.size()) return {INTERNAL, "bad DATA frame"};
                body.append(f.payload, off, f.payload.size() - off - pad);
                if (f.flags & END_STREAM) ended = true;
                break;
            }
            case RST_STREAM:
                return {UNAVAILABLE, "stream reset by peer (code " + std::to_string(f.payload.size() == 4 ? rd32(f.payload, 0) : 0) + ")"};
            case GOAWAY:
                return {UNAVAILABLE, "peer sent GOAWAY (error " + std::to_string(f.payload.size() >= 8 ? rd32(f.payload, 4) : 0) + ": " +
                                         (f.payload.size() > 8 ? f.payload.substr(8) : std::string()) + ")"};
            default:
                break;
        }
    }
    Status st;
    bool have_status = false;
    std::string http_status;
    for (auto& h : all) {
        if (h.first == "grpc-status") { st.code = atoi(h.second.c_str()); have_status = true; }
        else if (h.first == "grpc-message") st.message = pct_decode(h.second);
        else if (h.first == ":status") http_status = h.second;
    }
    if (!have_status) return {UNKNOWN, "response carried no grpc-status (HTTP status " + http_status + ")"};
    if (st.ok()) {
        std::vector<std::string> msgs;
        if (!grpc_unframe(body, &msgs) || msgs.size() != 1) return {INTERNAL, "malformed gRPC response body"};
        if (response) *response = msgs[0];
    }
    return st;
}

}  // namespace h2
