// Load generator for bbp-blindbid-server: K client threads, each opening ONE CONNECTION PER REQUEST as the reference's Go
// client does (src/futures/main.rs:64-110: one TLV frame in, one reply frame — or nothing — out, connection closed), replaying
// pre-encoded request frames from a file. Prints one JSON line: requests/s and latency percentiles.
//
// frames file: u32 count, then per frame u32 length + bytes (little-endian; written by tools/server_bench.py).
// usage: bbp-loadgen SOCKET FRAMES THREADS REQUESTS [expect-first-reply-byte]
#include <sys/socket.h>
#include <sys/un.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

static bool roundtrip(const sockaddr_un &addr, const std::vector<uint8_t> &frame, std::vector<uint8_t> &reply) {
    int fd = socket(AF_UNIX, SOCK_STREAM, 0);
    if (fd < 0) return false;
    int tries = 0;
    while (connect(fd, (const sockaddr *)&addr, sizeof addr) != 0) {   // a full listen backlog: back off and retry
        if (++tries > 2000) { close(fd); return false; }
        usleep(200);
    }
    size_t off = 0;
    while (off < frame.size()) {
        ssize_t w = write(fd, frame.data() + off, frame.size() - off);
        if (w <= 0) { close(fd); return false; }
        off += (size_t)w;
    }
    reply.clear();
    uint8_t buf[4096];
    for (;;) {
        ssize_t r = read(fd, buf, sizeof buf);
        if (r <= 0) break;
        reply.insert(reply.end(), buf, buf + r);
    }
    close(fd);
    return true;
}

int main(int argc, char **argv) {
    if (argc < 5) { fprintf(stderr, "usage: bbp-loadgen SOCKET FRAMES THREADS REQUESTS [min-reply-bytes]\n"); return 2; }
    sockaddr_un addr;
    memset(&addr, 0, sizeof addr);
    addr.sun_family = AF_UNIX;
    strncpy(addr.sun_path, argv[1], sizeof addr.sun_path - 1);
    FILE *f = fopen(argv[2], "rb");
    if (!f) { perror("frames"); return 2; }
    uint32_t count = 0;
    if (fread(&count, 4, 1, f) != 1 || !count) { fprintf(stderr, "empty frames file\n"); return 2; }
    std::vector<std::vector<uint8_t>> frames(count);
    for (auto &fr : frames) {
        uint32_t len = 0;
        if (fread(&len, 4, 1, f) != 1) { fprintf(stderr, "truncated frames file\n"); return 2; }
        fr.resize(len);
        if (len && fread(fr.data(), 1, len, f) != len) { fprintf(stderr, "truncated frames file\n"); return 2; }
    }
    fclose(f);
    const int threads = atoi(argv[3]);
    const long total = atol(argv[4]);
    const size_t min_reply = argc > 5 ? (size_t)atol(argv[5]) : 1;
    std::atomic<long> next{0}, failed{0}, short_reply{0};
    std::vector<std::vector<float>> lat(threads);
    auto t0 = std::chrono::steady_clock::now();
    std::vector<std::thread> th;
    for (int k = 0; k < threads; k++)
        th.emplace_back([&, k] {
            std::vector<uint8_t> reply;
            for (;;) {
                long i = next++;
                if (i >= total) break;
                auto a = std::chrono::steady_clock::now();
                bool ok = roundtrip(addr, frames[(size_t)i % count], reply);
                float ms = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - a).count();
                if (!ok) failed++;
                else if (reply.size() < min_reply) short_reply++;
                lat[k].push_back(ms);
            }
        });
    for (auto &t : th) t.join();
    double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    std::vector<float> all;
    for (auto &v : lat) all.insert(all.end(), v.begin(), v.end());
    std::sort(all.begin(), all.end());
    auto pct = [&](double p) { return all.empty() ? 0.f : all[std::min(all.size() - 1, (size_t)(p * all.size()))]; };
    printf("{\"requests\": %ld, \"threads\": %d, \"seconds\": %.4f, \"requests_per_s\": %.1f, \"latency_ms\": {\"p50\": %.3f, \"p90\": %.3f, \"p99\": %.3f, \"max\": %.3f}, "
           "\"failed\": %ld, \"short_replies\": %ld}\n",
           total, threads, s, total / s, pct(0.5), pct(0.9), pct(0.99), all.empty() ? 0.f : all.back(), failed.load(), short_reply.load());
    return failed.load() ? 1 : 0;
}
