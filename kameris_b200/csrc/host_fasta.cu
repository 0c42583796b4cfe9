// host_fasta.cu -- FASTA ingest for the `kmers` step: record[0] of every file of a directory listing,
// read by a pool of host threads into ONE contiguous buffer (what kam_kmers_host / kam_pack_ascii take).
//
// Replaces the per-file front half of generation_cgr's task (ifstream + getline + trim + first-record
// selection, generation_cgr @0x10000d691-0x10000d83e; task scheme of mmg::ParallelExecutor
// @0x10001a6f0: hardware_concurrency workers pulling file indices from one atomic counter).  A Python
// thread pool calling kam_fasta_record0 per file manages ~8 k files/s; here a file costs one open / fstat /
// read / close and one pass over its bytes.
//
// Layout: file i may only produce a record no longer than the file, so it is extracted straight into
// chars[file_offsets[i] ...] (file_offsets = exclusive prefix sum of the file sizes, capacity
// file_offsets[n]); when every worker is done the records are slid together in file order (dest <= src,
// so one in-order memmove pass is safe) and start[] receives their offsets.
#include <errno.h>
#include <fcntl.h>
#include <string.h>
#include <sys/stat.h>
#include <unistd.h>

#include <atomic>
#include <thread>
#include <vector>

#include "common.cuh"

namespace {

struct Ingest {
    const char *const *paths;
    uint32_t n;
    const uint64_t *off;
    uint8_t *chars;
    uint64_t *len;
    std::atomic<uint32_t> next{0};
    std::atomic<uint32_t> bad{0xffffffffu};   // smallest failing index
    std::atomic<int> bad_errno{0};
};

void note_bad(Ingest &g, uint32_t i, int err) {
    uint32_t cur = g.bad.load();
    while (i < cur && !g.bad.compare_exchange_weak(cur, i)) {
    }
    if (g.bad.load() == i) g.bad_errno.store(err);
}

void ingest_worker(Ingest *gp) {
    Ingest &g = *gp;
    std::vector<char> buf;
    for (;;) {
        const uint32_t i = g.next.fetch_add(1);
        if (i >= g.n) break;
        g.len[i] = 0;
        const uint64_t cap = g.off[i + 1] - g.off[i];
        const int fd = open(g.paths[i], O_RDONLY | O_CLOEXEC | O_NONBLOCK)   /* a FIFO in the directory must not block the step */;
        if (fd < 0) {
            note_bad(g, i, errno);
            continue;
        }
        struct stat st;
        if (fstat(fd, &st) != 0) {
            note_bad(g, i, errno);
            close(fd);
            continue;
        }
        if (!S_ISREG(st.st_mode) || (uint64_t)st.st_size > cap) {
            note_bad(g, i, S_ISREG(st.st_mode) ? EFBIG : EISDIR);   // it grew since it was sized / not a file
            close(fd);
            continue;
        }
        const size_t size = (size_t)st.st_size;
        if (buf.size() < size) buf.resize(size);
        size_t got = 0;
        while (got < size) {
            const ssize_t r = read(fd, buf.data() + got, size - got);
            if (r < 0 && errno == EINTR) continue;
            if (r <= 0) break;
            got += (size_t)r;
        }
        close(fd);
        const int64_t rl = kam_fasta_record0(buf.data(), got, (char *)g.chars + g.off[i]);
        if (rl < 0) {
            note_bad(g, i, 0);   // no record in the file: the reference dereferences begin() of an empty vector
            continue;
        }
        g.len[i] = (uint64_t)rl;
    }
}

int clamp_threads(int threads, uint32_t n) {
    if (threads <= 0) threads = (int)std::thread::hardware_concurrency();
    if (threads < 1) threads = 1;
    if (threads > 256) threads = 256;
    if ((uint32_t)threads > n) threads = (int)(n ? n : 1);
    return threads;
}

}  // namespace

extern "C" {

int kam_fasta_file_sizes(const char *const *paths, uint32_t n, uint64_t *sizes, int threads) {
    KAM_REQUIRE(n == 0 || (paths && sizes), "null pointer");
    std::atomic<uint32_t> next{0};
    std::atomic<uint32_t> bad{0xffffffffu};
    auto work = [&]() {
        for (;;) {
            const uint32_t i = next.fetch_add(1);
            if (i >= n) break;
            struct stat st;
            if (stat(paths[i], &st) != 0) {
                uint32_t cur = bad.load();
                while (i < cur && !bad.compare_exchange_weak(cur, i)) {
                }
                sizes[i] = 0;
            } else {
                sizes[i] = S_ISREG(st.st_mode) ? (uint64_t)st.st_size : 0;
            }
        }
    };
    const int t = clamp_threads(threads, n);
    std::vector<std::thread> pool;
    for (int k = 1; k < t; ++k) pool.emplace_back(work);
    work();
    for (auto &th : pool) th.join();
    if (bad.load() != 0xffffffffu) {
        kam_set_error("cannot stat \"%s\"", paths[bad.load()]);
        return KAM_E_INVALID;
    }
    return KAM_OK;
}

int kam_fasta_read_files(const char *const *paths, uint32_t n, const uint64_t *file_offsets, uint8_t *chars,
                         uint64_t *start, int threads, uint32_t *bad_index) {
    KAM_REQUIRE(start, "null pointer");
    KAM_REQUIRE(n == 0 || (paths && file_offsets && chars), "null pointer");
    if (bad_index) *bad_index = 0xffffffffu;
    start[0] = 0;
    if (n == 0) return KAM_OK;
    std::vector<uint64_t> len(n);
    Ingest g;
    g.paths = paths;
    g.n = n;
    g.off = file_offsets;
    g.chars = chars;
    g.len = len.data();
    const int t = clamp_threads(threads, n);
    std::vector<std::thread> pool;
    for (int k = 1; k < t; ++k) pool.emplace_back(ingest_worker, &g);
    ingest_worker(&g);
    for (auto &th : pool) th.join();
    const uint32_t bad = g.bad.load();
    if (bad != 0xffffffffu) {
        if (bad_index) *bad_index = bad;
        const int e = g.bad_errno.load();
        if (e) kam_set_error("cannot read \"%s\": %s", paths[bad], strerror(e));
        else kam_set_error("no FASTA record in \"%s\"", paths[bad]);
        return KAM_E_INVALID;
    }
    uint64_t pos = 0;
    for (uint32_t i = 0; i < n; ++i) {          // slide the records together, in file order
        if (len[i] && pos != file_offsets[i]) memmove(chars + pos, chars + file_offsets[i], len[i]);
        pos += len[i];
        start[i + 1] = pos;
    }
    return KAM_OK;
}

}  // extern "C"
