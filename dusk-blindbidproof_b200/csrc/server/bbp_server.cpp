// Unix-socket server shell of the B200 blind-bid backend: the process boundary the Go node talks to.
//
// Mirrors the reference's binary (src/main.rs:13-58: flags -b / --bind-path, -l / --log-level, default socket
// $TMPDIR/dusk-uds-blindbid) and its per-connection protocol (src/futures/main.rs:64-110): one TLV request frame per
// connection, payload byte 0 = opcode, 1 = prove, 2 = verify; the reply is one TLV frame, or nothing at all for an unknown
// opcode / a prove-side error, after which the connection is closed.
//
// What changes behind that boundary is the execution model. The reference runs every connection start-to-finish on one pool
// thread (src/futures/prove.rs:21-25); here a few epoll reader threads accept, buffer and parse (no thread per connection: at
// 40 k connections/s thread creation alone was the bound), and ONE executor drains everything that is pending into
// a single bbp_wire_execute call — all prove requests as one batched GPU pass, all verify requests as one random linear
// combination — so N concurrent clients cost about as much as one. A short gathering window (--window-us) trades a little
// latency for batch size; every client still gets exactly the reply the per-request path would have produced.
//
// Plain C++ over the C ABI of include/bbp.h (no CUDA in this file). Byte-level TLV framing: see csrc/wire.h (UNPINNED).
#include <algorithm>
#include <atomic>
#include <cerrno>
#include <chrono>
#include <condition_variable>
#include <csignal>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <mutex>
#include <string>
#include <thread>
#include <unordered_set>
#include <vector>
#include <fcntl.h>
#include <sys/epoll.h>
#include <sys/socket.h>
#include <sys/stat.h>
#include <sys/un.h>
#include <unistd.h>
#include "../../../include/bbp.h"

namespace {

enum level { L_ERROR = 0, L_WARN, L_INFO, L_DEBUG, L_TRACE };
int g_level = L_INFO;
#define LOG(lvl, ...)                                                                  \
    do {                                                                               \
        if ((lvl) <= g_level) { fprintf(stderr, "[bbp-server] " __VA_ARGS__); fputc('\n', stderr); } \
    } while (0)

struct pending {
    int fd;
    bbp_wire_request *req;
};

std::mutex g_mu;
std::condition_variable g_cv;
std::deque<pending> g_queue;
std::atomic<bool> g_stop{false};
std::atomic<uint64_t> g_served{0}, g_batches{0};
std::atomic<int> g_open{0};                 // connections being read or waiting for their reply
const size_t MAX_REQUEST = 16u << 20;       // a request is a few kilobytes (L = 202: ~10 KB); refuse absurd frames before buffering them
const int MAX_OPEN = 8192;                  // connections being read or waiting for their reply: bounded
const int READ_TIMEOUT_S = 30;              // a client that connects and stays silent is dropped

bool write_all(int fd, const uint8_t *buf, size_t n) {
    while (n) {
        ssize_t r = write(fd, buf, n);
        if (r <= 0) return false;
        buf += r;
        n -= (size_t)r;
    }
    return true;
}

// one connection being read: the request frame arrives in up to three steps — the tag byte, the length field it announces
// (csrc/wire.h: tlv_header), the payload (TlvReader::next, futures/main.rs:68-76)
struct conn {
    int fd;
    std::vector<uint8_t> buf;
    size_t need = 1, hdr = 0, pl = 0;
    int stage = 0;
    std::chrono::steady_clock::time_point last;
};

void drop(conn *c, const char *why) {
    if (why) LOG(L_ERROR, "Error resolving the request: %s", why);
    close(c->fd);
    g_open--;
    delete c;
}

// returns true when the connection is finished with (queued or dropped), false when it waits for more bytes
bool on_readable(conn *c, int ep) {   // ep: the epoll set the descriptor is registered in, -1 if none yet
    for (;;) {
        if (c->buf.size() < c->need) {
            const size_t have = c->buf.size();
            c->buf.resize(c->need);
            ssize_t r = read(c->fd, c->buf.data() + have, c->need - have);
            if (r < 0 && (errno == EAGAIN || errno == EWOULDBLOCK)) { c->buf.resize(have); return false; }
            if (r <= 0) {
                c->buf.resize(have);
                drop(c, have == 0 ? "The request was not provided" : "unexpected end of the request frame");
                return true;
            }
            c->buf.resize(have + (size_t)r);
            c->last = std::chrono::steady_clock::now();
            continue;
        }
        if (c->stage == 0) {
            const int w = c->buf[0];
            if (w != 1 && w != 2 && w != 4 && w != 8) { drop(c, "malformed frame"); return true; }
            c->need = 1 + (size_t)w;
            c->stage = 1;
        } else if (c->stage == 1) {
            if (bbp_wire_frame_len(c->buf.data(), c->buf.size(), &c->hdr, &c->pl) < 0 || c->hdr == 0) { drop(c, "malformed frame"); return true; }
            if (c->pl > MAX_REQUEST) { LOG(L_ERROR, "Error resolving the request: frame of %zu bytes refused", c->pl); drop(c, nullptr); return true; }
            c->need = c->hdr + c->pl;
            c->stage = 2;
        } else {
            bbp_wire_request *req = nullptr;
            int op = c->pl ? bbp_wire_parse(c->buf.data() + c->hdr, c->pl, &req) : BBP_ERR_FORMAT;
            if (op <= 0) {   // Message::Error: nothing is written
                drop(c, op == 0 ? "Undefined operation code" : "malformed request");
                return true;
            }
            LOG(L_TRACE, "request queued (opcode %d)", op);
            if (ep >= 0) epoll_ctl(ep, EPOLL_CTL_DEL, c->fd, nullptr);    // before the executor can close the descriptor
            fcntl(c->fd, F_SETFL, fcntl(c->fd, F_GETFL) & ~O_NONBLOCK);   // the executor writes the reply with blocking calls
            {
                std::lock_guard<std::mutex> lock(g_mu);
                g_queue.push_back({c->fd, req});
            }
            g_cv.notify_one();
            delete c;   // the executor owns the descriptor (and counts the connection out) from here
            return true;
        }
    }
}

// accepts and reads: every reader has its own epoll set and shares the listening socket (EPOLLEXCLUSIVE: one wake-up per
// connection burst); a connection stays with the reader that accepted it until its request is queued
void reader_thread(int listen_fd) {
    int ep = epoll_create1(EPOLL_CLOEXEC);
    if (ep < 0) { perror("epoll_create1"); return; }
    epoll_event ev;
    memset(&ev, 0, sizeof ev);
    ev.events = EPOLLIN | EPOLLEXCLUSIVE;
    ev.data.ptr = nullptr;   // nullptr marks the listening socket
    epoll_ctl(ep, EPOLL_CTL_ADD, listen_fd, &ev);
    std::unordered_set<conn *> live;
    std::vector<epoll_event> evs(256);
    auto last_sweep = std::chrono::steady_clock::now();
    while (!g_stop) {
        int n = epoll_wait(ep, evs.data(), (int)evs.size(), 500);
        for (int i = 0; i < n; i++) {
            if (evs[i].data.ptr == nullptr) {
                for (int k = 0; k < 64; k++) {   // a bounded burst, so that reading is not starved by a connect flood
                    int fd = accept4(listen_fd, nullptr, nullptr, SOCK_NONBLOCK | SOCK_CLOEXEC);
                    if (fd < 0) break;
                    if (g_open.load() >= MAX_OPEN) { LOG(L_WARN, "too many open connections: refused"); close(fd); continue; }
                    g_open++;
                    conn *c = new conn;
                    c->fd = fd;
                    c->last = std::chrono::steady_clock::now();
                    if (on_readable(c, -1)) continue;   // the request usually arrives with the connection
                    epoll_event ce;
                    memset(&ce, 0, sizeof ce);
                    ce.events = EPOLLIN | EPOLLRDHUP;
                    ce.data.ptr = c;
                    if (epoll_ctl(ep, EPOLL_CTL_ADD, fd, &ce) != 0) { drop(c, "epoll registration failed"); continue; }
                    live.insert(c);
                }
            } else {
                conn *c = (conn *)evs[i].data.ptr;
                if (!live.count(c)) continue;
                if (on_readable(c, ep)) live.erase(c);   // a dropped connection leaves the set when its descriptor is closed
            }
        }
        // a client that connects and stays silent does not hold its slot forever
        auto now = std::chrono::steady_clock::now();
        if (now - last_sweep > std::chrono::seconds(1)) {
            last_sweep = now;
            for (auto it = live.begin(); it != live.end();) {
                if (now - (*it)->last > std::chrono::seconds(READ_TIMEOUT_S)) {
                    drop(*it, "timed out waiting for the request frame");
                    it = live.erase(it);
                } else ++it;
            }
        }
    }
    for (conn *c : live) drop(c, nullptr);
    close(ep);
}

// a resolved request on its way back to the client
struct done {
    int fd;
    bbp_wire_request *req;
    uint8_t *reply;     // nullptr: nothing is written (Message::Error)
    size_t len;
};
const size_t REPLY_CHUNK = 64;
std::mutex g_done_mu;
std::condition_variable g_done_cv;
std::deque<std::vector<done>> g_done;
std::atomic<bool> g_exec_finished{false};

void replier_thread() {
    for (;;) {
        std::vector<done> chunk;
        {
            std::unique_lock<std::mutex> lock(g_done_mu);
            g_done_cv.wait_for(lock, std::chrono::milliseconds(100), [] { return !g_done.empty() || g_exec_finished.load(); });
            if (g_done.empty()) {
                if (g_exec_finished) return;
                continue;
            }
            chunk = std::move(g_done.front());
            g_done.pop_front();
        }
        for (const done &d : chunk) {
            if (d.reply) {
                if (!write_all(d.fd, d.reply, d.len)) LOG(L_WARN, "client went away before the reply");
                else LOG(L_TRACE, "Request resolved");
            } else {
                LOG(L_ERROR, "Error resolving the request: no reply is written");
            }
            close(d.fd);
            g_open--;
            bbp_wire_reply_free(d.reply);
            bbp_wire_request_free(d.req);
        }
    }
}

void executor_thread(bbp_ctx *ctx, unsigned window_us, size_t max_batch) {
    while (!g_stop) {
        std::vector<pending> batch;
        {
            std::unique_lock<std::mutex> lock(g_mu);
            g_cv.wait_for(lock, std::chrono::milliseconds(100), [] { return !g_queue.empty() || g_stop.load(); });
            if (g_queue.empty()) continue;
            if (window_us) {   // gathering window: let concurrent clients land in the same batch
                lock.unlock();
                std::this_thread::sleep_for(std::chrono::microseconds(window_us));
                lock.lock();
            }
            while (!g_queue.empty() && batch.size() < max_batch) { batch.push_back(g_queue.front()); g_queue.pop_front(); }
        }
        const size_t n = batch.size();
        std::vector<bbp_wire_request *> reqs(n);
        std::vector<uint8_t *> replies(n, nullptr);
        std::vector<size_t> lens(n, 0);
        for (size_t i = 0; i < n; i++) reqs[i] = batch[i].req;
        auto t0 = std::chrono::steady_clock::now();
        int rc = bbp_wire_execute(ctx, n, reqs.data(), nullptr, replies.data(), lens.data());
        double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        if (rc) LOG(L_ERROR, "Error resolving %zu requests: backend status %d", n, rc);
        else LOG(L_DEBUG, "batch of %zu requests resolved in %.2f ms", n, ms);
        if (rc) for (size_t i = 0; i < n; i++) bbp_wire_reply_free(replies[i]);
        // replies are written (and connections closed) by the replier threads while this thread starts on the next batch:
        // ~15 us of system calls per connection, which serialised here capped the server at ~45 k requests/s
        for (size_t lo = 0; lo < n; lo += REPLY_CHUNK) {
            std::vector<done> chunk;
            for (size_t i = lo; i < std::min(n, lo + REPLY_CHUNK); i++) chunk.push_back({batch[i].fd, batch[i].req, rc ? nullptr : replies[i], lens[i]});
            {
                std::lock_guard<std::mutex> lock(g_done_mu);
                g_done.push_back(std::move(chunk));
            }
            g_done_cv.notify_one();
        }
        g_served += n;
        g_batches++;
    }
}

int g_listen_fd = -1;
void on_signal(int) {
    g_stop = true;
    if (g_listen_fd >= 0) shutdown(g_listen_fd, SHUT_RDWR);   // wakes the readers
}

}  // namespace

int main(int argc, char **argv) {
    const char *tmp = getenv("TMPDIR");
    std::string bind_path = std::string(tmp && *tmp ? tmp : "/tmp") + "/dusk-uds-blindbid";   // env::temp_dir() + "dusk-uds-blindbid"
    std::string lvl = "info";
    int device = 0;
    unsigned window_us = 200;
    size_t max_batch = 4096;
    int readers = 4;
    for (int i = 1; i < argc; i++) {
        std::string a = argv[i];
        auto val = [&](const char *name) -> const char * {
            if (i + 1 >= argc) { fprintf(stderr, "missing value for %s\n", name); exit(2); }
            return argv[++i];
        };
        if (a == "-b" || a == "--bind-path") bind_path = val("--bind-path");
        else if (a == "-l" || a == "--log-level") lvl = val("--log-level");
        else if (a == "--device") device = atoi(val("--device"));
        else if (a == "--window-us") window_us = (unsigned)atoi(val("--window-us"));
        else if (a == "--max-batch") max_batch = (size_t)atol(val("--max-batch"));
        else if (a == "--readers") readers = std::max(1, atoi(val("--readers")));
        else if (a == "-h" || a == "--help") {
            printf("bbp-blindbid-server [-b|--bind-path BIND] [-l|--log-level error|warn|info|debug|trace] [--device N] [--window-us US] [--max-batch N] [--readers N]\n");
            return 0;
        } else { fprintf(stderr, "unknown argument %s\n", a.c_str()); return 2; }
    }
    const char *names[] = {"error", "warn", "info", "debug", "trace"};
    bool known = false;
    for (int k = 0; k < 5; k++)
        if (lvl == names[k]) { g_level = k; known = true; }
    if (!known) { fprintf(stderr, "invalid log level %s\n", lvl.c_str()); return 2; }

    bbp_ctx *ctx = nullptr;
    int rc = bbp_init(&ctx, device, 2048, 1);   // BulletproofGens::new(2048, 1), built once instead of per request (mod.rs:36)
    if (rc) { LOG(L_ERROR, "bbp_init failed with status %d (no CUDA device?)", rc); return 1; }

    int srv = socket(AF_UNIX, SOCK_STREAM, 0);
    if (srv < 0) { perror("socket"); return 1; }
    sockaddr_un addr;
    memset(&addr, 0, sizeof addr);
    addr.sun_family = AF_UNIX;
    if (bind_path.size() >= sizeof addr.sun_path) { fprintf(stderr, "bind path too long\n"); return 2; }
    strcpy(addr.sun_path, bind_path.c_str());
    unlink(bind_path.c_str());
    if (bind(srv, (sockaddr *)&addr, sizeof addr) != 0 || listen(srv, 1024) != 0) { perror("Failed binding socket"); return 1; }
    g_listen_fd = srv;
    signal(SIGINT, on_signal);
    signal(SIGTERM, on_signal);
    signal(SIGPIPE, SIG_IGN);
    LOG(L_INFO, "listening on %s (device %d, gathering window %u us)", bind_path.c_str(), device, window_us);

    std::thread exec(executor_thread, ctx, window_us, max_batch);
    std::vector<std::thread> rp;
    for (int k = 0; k < readers; k++) rp.emplace_back(replier_thread);
    fcntl(srv, F_SETFL, fcntl(srv, F_GETFL) | O_NONBLOCK);
    std::vector<std::thread> rd;
    for (int k = 0; k < readers; k++) rd.emplace_back(reader_thread, srv);
    while (!g_stop) std::this_thread::sleep_for(std::chrono::milliseconds(50));
    for (auto &t : rd) t.join();
    g_stop = true;
    g_cv.notify_all();
    exec.join();
    g_exec_finished = true;
    g_done_cv.notify_all();
    for (auto &t : rp) t.join();
    close(srv);
    unlink(bind_path.c_str());
    LOG(L_INFO, "served %llu requests in %llu batches", (unsigned long long)g_served.load(), (unsigned long long)g_batches.load());
    bbp_free(ctx);
    return 0;
}
