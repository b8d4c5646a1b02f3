// Unix-socket server shell of the B200 blind-bid backend: the process boundary the Go node talks to.
//
// Mirrors the reference's binary (src/main.rs:13-58: flags -b / --bind-path, -l / --log-level, default socket
// $TMPDIR/dusk-uds-blindbid) and its per-connection protocol (src/futures/main.rs:64-110): one TLV request frame per
// connection, payload byte 0 = opcode, 1 = prove, 2 = verify; the reply is one TLV frame, or nothing at all for an unknown
// opcode / a prove-side error, after which the connection is closed.
//
// What changes behind that boundary is the execution model. The reference runs every connection start-to-finish on one pool
// thread (src/futures/prove.rs:21-25); here reader threads only parse, and ONE executor drains everything that is pending into
// a single bbp_wire_execute call — all prove requests as one batched GPU pass, all verify requests as one random linear
// combination — so N concurrent clients cost about as much as one. A short gathering window (--window-us) trades a little
// latency for batch size; every client still gets exactly the reply the per-request path would have produced.
//
// Plain C++ over the C ABI of include/bbp.h (no CUDA in this file). Byte-level TLV framing: see csrc/wire.h (UNPINNED).
#include <atomic>
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
#include <vector>
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
const int MAX_OPEN = 8192;                  // one reader thread per connection: bounded
const int READ_TIMEOUT_S = 30;              // a client that connects and stays silent does not hold a thread forever

bool read_exact(int fd, uint8_t *buf, size_t n) {
    while (n) {
        ssize_t r = read(fd, buf, n);
        if (r <= 0) return false;
        buf += r;
        n -= (size_t)r;
    }
    return true;
}
bool write_all(int fd, const uint8_t *buf, size_t n) {
    while (n) {
        ssize_t r = write(fd, buf, n);
        if (r <= 0) return false;
        buf += r;
        n -= (size_t)r;
    }
    return true;
}

// one connection: read the request frame, parse it, queue it (TlvReader::next, futures/main.rs:68-76)
struct open_guard {   // counts the connection out when its reader gives up (queued requests are counted out by the executor)
    bool armed = true;
    ~open_guard() { if (armed) g_open--; }
};

void reader_thread(int fd) {
    open_guard guard;
    timeval tv = {READ_TIMEOUT_S, 0};
    setsockopt(fd, SOL_SOCKET, SO_RCVTIMEO, &tv, sizeof tv);
    uint8_t head[9];
    size_t hdr = 0, pl = 0;
    // tag byte, then the length field it announces (csrc/wire.h: tlv_header)
    if (!read_exact(fd, head, 1)) { LOG(L_ERROR, "Error resolving the request: The request was not provided"); close(fd); return; }
    const int w = head[0];
    if ((w != 1 && w != 2 && w != 4 && w != 8) || !read_exact(fd, head + 1, (size_t)w) || bbp_wire_frame_len(head, 1 + (size_t)w, &hdr, &pl) < 0 || hdr == 0) {
        LOG(L_ERROR, "Error resolving the request: malformed frame");
        close(fd);
        return;
    }
    if (pl > MAX_REQUEST) { LOG(L_ERROR, "Error resolving the request: frame of %zu bytes refused", pl); close(fd); return; }
    std::vector<uint8_t> payload(pl);
    if (pl && !read_exact(fd, payload.data(), pl)) {
        LOG(L_ERROR, "Error resolving the request: unexpected end of the request frame");
        close(fd);
        return;
    }
    bbp_wire_request *req = nullptr;
    int op = pl ? bbp_wire_parse(payload.data(), pl, &req) : BBP_ERR_FORMAT;
    if (op <= 0) {
        LOG(L_ERROR, "Error resolving the request: %s", op == 0 ? "Undefined operation code" : "malformed request");
        close(fd);   // Message::Error: nothing is written
        return;
    }
    LOG(L_TRACE, "request queued (opcode %d)", op);
    guard.armed = false;
    {
        std::lock_guard<std::mutex> lock(g_mu);
        g_queue.push_back({fd, req});
    }
    g_cv.notify_one();
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
        for (size_t i = 0; i < n; i++) {
            if (!rc && replies[i]) {
                if (!write_all(batch[i].fd, replies[i], lens[i])) LOG(L_WARN, "client went away before the reply");
                else LOG(L_TRACE, "Request resolved");
            } else {
                LOG(L_ERROR, "Error resolving the request: no reply is written");
            }
            close(batch[i].fd);
            g_open--;
            bbp_wire_reply_free(replies[i]);
            bbp_wire_request_free(batch[i].req);
        }
        g_served += n;
        g_batches++;
    }
}

int g_listen_fd = -1;
void on_signal(int) {
    g_stop = true;
    if (g_listen_fd >= 0) shutdown(g_listen_fd, SHUT_RDWR);   // wakes accept()
}

}  // namespace

int main(int argc, char **argv) {
    const char *tmp = getenv("TMPDIR");
    std::string bind_path = std::string(tmp && *tmp ? tmp : "/tmp") + "/dusk-uds-blindbid";   // env::temp_dir() + "dusk-uds-blindbid"
    std::string lvl = "info";
    int device = 0;
    unsigned window_us = 200;
    size_t max_batch = 4096;
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
        else if (a == "-h" || a == "--help") {
            printf("bbp-blindbid-server [-b|--bind-path BIND] [-l|--log-level error|warn|info|debug|trace] [--device N] [--window-us US] [--max-batch N]\n");
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
    while (!g_stop) {
        int fd = accept(srv, nullptr, nullptr);
        if (fd < 0) { if (g_stop) break; continue; }
        if (g_open.load() >= MAX_OPEN) { LOG(L_WARN, "too many open connections: refused"); close(fd); continue; }
        g_open++;
        std::thread(reader_thread, fd).detach();
    }
    g_stop = true;
    g_cv.notify_all();
    exec.join();
    close(srv);
    unlink(bind_path.c_str());
    LOG(L_INFO, "served %llu requests in %llu batches", (unsigned long long)g_served.load(), (unsigned long long)g_batches.load());
    bbp_free(ctx);
    return 0;
}
