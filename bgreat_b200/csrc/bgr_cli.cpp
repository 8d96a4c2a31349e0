// bgreat -- drop-in command line for the reference's `bgreat` binary (reference bgreat.cpp:54-130),
// with the mapping done on the GPU through the C ABI (include/bgreat_b200.h).
//
// Same getopt string, same defaults, same stdout lines, same `paths` / `notAligned.fa` record formats
// (reference alignerGreedy.cpp:393-428, aligner.cpp:588-609). Host pipeline: one parser thread
// (the reference's getReads semantics, exactly) -> GPU thread (double-buffered batch slots, pinned
// staging) -> writer (formats on -t threads, writes in file order).
//
// Deliberate, documented differences (DESIGN.md "CLI"):
//   * a read file that cannot be opened is an error (the reference spins forever);
//   * -o FILE writes the reads without any overlap to FILE (the reference accepts -o and ignores it);
//   * -b writes its paths and unmapped reads (the reference computes them and writes nothing);
//   * -c / -G (and -i together with -b) are not part of this path and are refused.
#include <getopt.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#include <chrono>
#include <condition_variable>
#include <deque>
#include <iostream>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/bgreat_b200.h"
#include "bgr_host.h"

using namespace bgr;

namespace {

static void* pinned_alloc(size_t bytes) {
    void* p = nullptr;
    if (bgr_host_alloc(bytes, &p) != BGR_OK) {
        fprintf(stderr, "bgreat: cannot allocate %zu bytes of pinned host memory: %s\n", bytes, bgr_last_error(nullptr));
        exit(1);
    }
    return p;
}
static void pinned_free(void* p) { bgr_host_free(p); }

struct Batch {
    ReadBatch rb;
    PodBuf<bgr_result> res;
    PodBuf<int32_t> pool;
    uint64_t pool_used = 0;
    int slot = -1;
    bool last_of_file = false;
    Batch() {
        rb.set_allocator(pinned_alloc, pinned_free);
        res.alloc = pool.alloc = pinned_alloc;
        res.dealloc = pool.dealloc = pinned_free;
    }
};

template <typename T>
class Queue {
   public:
    void push(T x) {
        std::unique_lock<std::mutex> l(m_);
        q_.push_back(x);
        cv_.notify_one();
    }
    T pop() {
        std::unique_lock<std::mutex> l(m_);
        cv_.wait(l, [&] { return !q_.empty(); });
        T x = q_.front();
        q_.pop_front();
        return x;
    }

   private:
    std::mutex m_;
    std::condition_variable cv_;
    std::deque<T> q_;
};

struct Options {
    std::string reads, unitigs = "unitig.fa", path_file = "paths", no_overlap_file, not_aligned_file = "notAligned.fa";
    int errors = 2, threads = 1, k = 30, effort = 2;
    bool brute = false, incomplete = false, fastq = false, correction = false, dog = false, have_o = false;
};

static inline char* put_int(char* p, int32_t v) {
    char tmp[12];
    uint32_t u = v < 0 ? (uint32_t)(-(int64_t)v) : (uint32_t)v;
    int n = 0;
    do {
        tmp[n++] = (char)('0' + u % 10);
        u /= 10;
    } while (u);
    if (v < 0) *p++ = '-';
    while (n) *p++ = tmp[--n];
    return p;
}

struct Outputs {
    FILE* paths = nullptr;
    FILE* not_aligned = nullptr;
    FILE* no_overlap = nullptr;
};

// header\npath\n for mapped reads (printPath, aligner.cpp:600-609), header\nread\n otherwise
static void format_range(const Batch& b, size_t lo, size_t hi, bool split_no_overlap, std::string& paths,
                         std::string& na, std::string& nov) {
    for (size_t i = lo; i < hi; ++i) {
        const char* h = b.rb.headers.p + b.rb.hoff.p[i];
        size_t hn = b.rb.hoff.p[i + 1] - b.rb.hoff.p[i];
        const bgr_result r = b.res.p[i];
        uint32_t flags = BGR_RESULT_FLAGS(r), len = BGR_RESULT_LEN(r);
        if (flags & BGR_F_MAPPED) {
            size_t at = paths.size();
            paths.resize(at + hn + 2 + 12 * (size_t)len);
            char* p = &paths[at];
            memcpy(p, h, hn);
            p += hn;
            *p++ = '\n';
            const int32_t* ids = b.pool.p + r.path_off;
            for (uint32_t j = 0; j < len; ++j) {
                p = put_int(p, ids[j]);
                *p++ = '.';
            }
            *p++ = '\n';
            paths.resize((size_t)(p - paths.data()));
        } else {
            std::string& o = (split_no_overlap && !(flags & BGR_F_OVERLAP_FOUND)) ? nov : na;
            const char* s = b.rb.seqs.p + b.rb.soff.p[i];
            size_t sn = b.rb.soff.p[i + 1] - b.rb.soff.p[i];
            o.append(h, hn);
            o.push_back('\n');
            o.append(s, sn);
            o.push_back('\n');
        }
    }
}

static void write_all(FILE* f, const std::string& s) {
    if (f && !s.empty()) fwrite(s.data(), 1, s.size(), f);
}

static void usage() {
    // reference bgreat.cpp:117-127
    std::cout << "-r read_file" << std::endl
              << "-k k_value (30)" << std::endl
              << "-g unitig_file (unitig.dot)" << std::endl
              << "-m n_missmatch (2)" << std::endl
              << "-t n_thread (1)" << std::endl
              << "-e effort put in mapping (2)" << std::endl
              << "-f path_file (paths)" << std::endl
              << "-a not_aligned_file (notAligned.fa)" << std::endl
              << "-p to align on paths instead of walks" << std::endl
              << "-q for fastq read file" << std::endl
              << "-c to output corrected reads" << std::endl;
}

}  // namespace

int main(int argc, char** argv) {
    Options o;
    int c;
    // reference bgreat.cpp:67: "r:k:g:m:t:e:f:o:a:biqpcG"
    while ((c = getopt(argc, argv, "r:k:g:m:t:e:f:o:a:biqpcG")) != -1) {
        switch (c) {
            case 'r': o.reads = optarg; break;
            case 'k': o.k = std::stoi(optarg); break;
            case 'g': o.unitigs = optarg; break;
            case 'm': o.errors = std::stoi(optarg); break;
            case 't': o.threads = std::stoi(optarg); break;
            case 'e': o.effort = std::stoi(optarg); break;
            case 'f': o.path_file = optarg; break;
            case 'a': o.not_aligned_file = optarg; break;
            case 'o':
                o.no_overlap_file = optarg;
                o.have_o = true;
                break;
            case 'b': o.brute = true; break;
            case 'i': o.incomplete = true; break;
            case 'q': o.fastq = true; break;
            case 'G': o.dog = true; break;
            case 'c': o.correction = true; break;
            case 'p': break;  // accepted, no effect (no `case 'p'` in the reference either)
        }
    }
    if (o.reads.empty()) {
        usage();
        return 0;
    }
    if (o.correction || o.dog || (o.incomplete && o.brute)) {
        fprintf(stderr, "bgreat: -c, -G and -b -i are outside the GPU mapping path and not supported\n");
        return 2;
    }
    // the reference keeps the thread count in an unsigned char (aligner.h:80)
    unsigned threads = (unsigned char)o.threads;
    if (threads == 0) threads = 1;

    const int n_slots = 2;
    bgr_ctx* ctx = nullptr;
    if (bgr_ctx_create(0, n_slots, &ctx) != BGR_OK) {
        fprintf(stderr, "bgreat: %s\n", bgr_last_error(nullptr));
        return 1;
    }

    // Aligner ctor opens the outputs first (aligner.h:85-86)
    Outputs out;
    out.paths = fopen(o.path_file.c_str(), "wb");
    out.not_aligned = fopen(o.not_aligned_file.c_str(), "wb");
    if (o.have_o) out.no_overlap = fopen(o.no_overlap_file.c_str(), "wb");

    // ---- indexUnitigs (aligner.cpp:537-547)
    auto t0 = std::chrono::system_clock::now();
    int rc = bgr_index_build_file(ctx, o.unitigs.c_str(), (uint32_t)o.k);
    if (rc != BGR_OK && rc != BGR_ERR_IO) {  // a missing unitig file means an empty index, like the reference
        fprintf(stderr, "bgreat: %s\n", bgr_last_error(ctx));
        return 1;
    }
    auto t1 = std::chrono::system_clock::now();
    std::cout << "Indexing in seconds : " << std::chrono::duration_cast<std::chrono::seconds>(t1 - t0).count()
              << std::endl;

    // ---- alignAll (aligner.cpp:550-597)
    auto start = std::chrono::system_clock::now();
    bgr_params params;
    params.mode = o.brute ? BGR_MODE_EXHAUSTIVE : BGR_MODE_GREEDY;
    params.errors = (uint32_t)o.errors;
    params.effort = (uint32_t)o.effort;
    params.reserved = 0;

    const size_t n_batches = n_slots + 3;
    std::vector<Batch> batches(n_batches);
    Queue<Batch*> free_q, parsed_q, done_q;
    for (auto& b : batches) free_q.push(&b);
    const size_t max_reads = 1u << 20, max_chars = 160u << 20;
    bool io_error = false;

    // files of the comma separated list, in order (aligner.cpp:554-576)
    std::vector<std::string> files;
    {
        size_t last = 0;
        for (size_t i = 0; i < o.reads.size(); ++i)
            if (o.reads[i] == ',') {
                files.push_back(o.reads.substr(last, i - last));
                last = i + 1;
            }
        files.push_back(o.reads.substr(last));
    }

    std::thread parser([&] {
        for (const std::string& file : files) {
            MappedFile mf;
            if (!mf.open(file.c_str())) {
                fprintf(stderr, "bgreat: cannot open read file %s\n", file.c_str());
                io_error = true;
                continue;
            }
            ReadParser rp(mf.data, mf.size, (uint32_t)o.k, o.fastq);
            // alignPartGreedy's loop (alignerGreedy.cpp:372-377): getReads(10000) until eof
            Batch* b = free_q.pop();
            b->rb.clear();
            while (!rp.done()) {
                rp.next_chunk(10000, b->rb);
                if (b->rb.size() >= max_reads || b->rb.seqs.n >= max_chars) {
                    parsed_q.push(b);
                    b = free_q.pop();
                    b->rb.clear();
                }
            }
            parsed_q.push(b);  // possibly empty: still flows through so that order is kept
        }
        parsed_q.push(nullptr);
    });

    std::thread gpu([&] {
        std::deque<Batch*> inflight;
        int next_slot = 0;
        bool input_done = false;
        auto finish = [&](Batch* b) {
            uint64_t used = 0;
            int r = bgr_slot_wait(ctx, b->slot, &used);
            if (r == BGR_ERR_LIMIT) {  // pool buffer too small: come back with a bigger one
                b->pool.reserve(used + 1024);
                r = bgr_slot_download(ctx, b->slot, b->res.p, b->pool.p, b->pool.cap);
                if (r == BGR_OK) r = bgr_slot_wait(ctx, b->slot, &used);
            }
            if (r != BGR_OK) {
                fprintf(stderr, "bgreat: %s\n", bgr_last_error(ctx));
                exit(1);
            }
            b->pool_used = used;
            done_q.push(b);
        };
        while (!input_done || !inflight.empty()) {
            if (!input_done && (int)inflight.size() < n_slots) {
                Batch* b = parsed_q.pop();
                if (!b) {
                    input_done = true;
                    continue;
                }
                b->slot = next_slot;
                next_slot = (next_slot + 1) % n_slots;
                size_t n = b->rb.size();
                b->res.reserve(n + 1);
                b->pool.reserve(n * 8 + 1024);
                int r = bgr_slot_upload(ctx, b->slot, b->rb.seqs.p, b->rb.soff.p, (uint32_t)n);
                if (r == BGR_OK) r = bgr_slot_run(ctx, b->slot, &params);
                if (r == BGR_OK) r = bgr_slot_download(ctx, b->slot, b->res.p, b->pool.p, b->pool.cap);
                if (r != BGR_OK) {
                    fprintf(stderr, "bgreat: %s\n", bgr_last_error(ctx));
                    exit(1);
                }
                inflight.push_back(b);
            } else {
                Batch* b = inflight.front();
                inflight.pop_front();
                finish(b);
            }
        }
        done_q.push(nullptr);
    });

    // writer: this thread
    {
        std::vector<std::string> sp(threads), sn(threads), so(threads);
        for (;;) {
            Batch* b = done_q.pop();
            if (!b) break;
            size_t n = b->rb.size();
            unsigned T = (unsigned)std::min<size_t>(threads, n / 4096 + 1);
            std::vector<std::thread> th;
            for (unsigned t = 0; t < T; ++t) {
                sp[t].clear();
                sn[t].clear();
                so[t].clear();
                size_t lo = n * t / T, hi = n * (t + 1) / T;
                if (t + 1 == T) format_range(*b, lo, hi, o.have_o, sp[t], sn[t], so[t]);
                else th.emplace_back([&, t, lo, hi] { format_range(*b, lo, hi, o.have_o, sp[t], sn[t], so[t]); });
            }
            for (auto& x : th) x.join();
            for (unsigned t = 0; t < T; ++t) {
                write_all(out.paths, sp[t]);
                write_all(out.not_aligned, sn[t]);
                write_all(out.no_overlap, so[t]);
            }
            free_q.push(b);
        }
    }
    parser.join();
    gpu.join();
    for (const std::string& file : files) std::cout << file << std::endl;

    uint64_t cnt[BGR_C_COUNT];
    bgr_counters(ctx, cnt);
    const unsigned readNumber = (unsigned)cnt[BGR_C_READS], noOverlapRead = (unsigned)cnt[BGR_C_NOOVERLAP],
                   alignedRead = (unsigned)cnt[BGR_C_ALIGNED], notAligned = (unsigned)cnt[BGR_C_NOTALIGNED];
    // aligner.cpp:588-596, same expressions (float arithmetic, -nan on 0/0 included)
    std::cout << "The End" << std::endl;
    std::cout << "Reads : " << readNumber << std::endl;
    std::cout << "No overlap : " << noOverlapRead << " Percent : " << (100 * float(noOverlapRead)) / readNumber << std::endl;
    std::cout << "Got overlap : " << alignedRead + notAligned << " Percent : "
              << (100 * float(alignedRead + notAligned)) / readNumber << std::endl;
    std::cout << "Overlap and aligned : " << alignedRead << " Percent : "
              << (100 * float(alignedRead)) / (alignedRead + notAligned) << std::endl;
    std::cout << "Overlap but not aligned : " << notAligned << " Percent : "
              << (100 * float(notAligned)) / (alignedRead + notAligned) << std::endl;
    auto end = std::chrono::system_clock::now();
    auto secs = std::chrono::duration_cast<std::chrono::seconds>(end - start).count();
    std::cout << "Reads/seconds : " << readNumber / (secs + 1) << std::endl;
    std::cout << "Mapping in seconds : " << secs << std::endl;

    if (out.paths) fclose(out.paths);
    if (out.not_aligned) fclose(out.not_aligned);
    if (out.no_overlap) fclose(out.no_overlap);
    batches.clear();
    bgr_ctx_destroy(ctx);
    return io_error ? 1 : 0;
}
