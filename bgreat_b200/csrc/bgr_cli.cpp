// bgreat -- drop-in command line for the reference's `bgreat` binary (reference bgreat.cpp:54-130),
// with the mapping done on the GPU through the C ABI (include/bgreat_b200.h).
//
// Same getopt string, same defaults, same stdout lines, same `paths` / `notAligned.fa` record formats
// (reference alignerGreedy.cpp:393-428, aligner.cpp:588-609). Host
// pipeline: reader (ParallelReader: -t workers cut and parse ranges of the memory-mapped input with the
// reference's getReads semantics, exactly) -> GPU thread (double-buffered batch slots, pinned staging) ->
// writer (-t workers format straight into the memory-mapped output files, in file order).
//
// Deliberate, documented differences (DESIGN.md "CLI"):
//   * a read file that cannot be opened is an error (the reference spins forever);
//   * -o FILE writes the reads without any overlap to FILE (the reference accepts -o and ignores it);
//   * -b writes its paths and unmapped reads (the reference computes them and writes nothing);
//   * -G (and -i together with -b) are not part of this path and are refused; -c (corrected reads instead
//     of paths, alignerGreedy.cpp:394-404) is supported for the greedy aligner, where the reference has it.
#include <fcntl.h>
#include <getopt.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <atomic>
#include <chrono>
#include <condition_variable>
#include <deque>
#include <iostream>
#include <memory>
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
    ParsedBatch pb;
    PodBuf<bgr_result> res;
    PodBuf<int32_t> pool;
    PodBuf<uint64_t> corr;  // -c: corrected reads, 2 bits per base (layout: BGR_CORRECTED_BASE)
    uint64_t pool_used = 0;
    int slot = -1;
    uint64_t seq = 0;  // position in the input: the writer emits batches in this order
    // text mode (regular FASTA): the batch is a range of the input text, parsed and formatted on the device
    // (bgr_slot_upload_text / bgr_slot_format_text); the host only copies text in and out
    bool text_mode = false;
    PodBuf<char> text;     // the range's text (pinned)
    const char* src = nullptr;  // the mapped input and the range in it: parsed here if the device reports it irregular
    size_t src_a = 0, src_b = 0;
    PodBuf<char> otext[3];  // the three output streams (pinned)
    uint64_t oused[3] = {0, 0, 0};
    Batch() {
        pb.set_allocator(pinned_alloc, pinned_free);
        res.alloc = pool.alloc = corr.alloc = text.alloc = pinned_alloc;
        res.dealloc = pool.dealloc = corr.dealloc = text.dealloc = pinned_free;
        for (auto& o : otext) o.alloc = pinned_alloc, o.dealloc = pinned_free;
    }
    // Buffers for a batch cut from `range_bytes` of text, sized once: pinning memory costs about as much per
    // byte as parsing it, and growing a pinned buffer pins it twice. Reads shorter than ~60 bases make the
    // per-read arrays grow later, which is only slower.
    void prealloc_text(size_t range_bytes, bool split_no_overlap) {
        text.reserve(range_bytes + 4096);
        otext[0].reserve(range_bytes / 2 + 4096);  // paths are shorter than the reads they stand for
        otext[1].reserve(range_bytes * 5 / 8 + 4096);
        if (split_no_overlap) otext[2].reserve(range_bytes / 2 + 4096);
    }
    void prealloc(size_t range_bytes) {
        const size_t reads = range_bytes / 96 + 1024;
        pb.seqs.reserve(range_bytes + 4096);
        pb.soff.reserve(reads + 2);
        pb.hdr.reserve(reads + 1);
        res.reserve(reads + 1);
        pool.reserve(reads * 8 + 1024);
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
    bool try_pop(T& x) {
        std::unique_lock<std::mutex> l(m_);
        if (q_.empty()) return false;
        x = q_.front();
        q_.pop_front();
        return true;
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

// printPath's to_string(x) (aligner.cpp:600-609) is where the writer's time goes (about nine integers per
// mapped read): digit count by comparisons, digits two at a time from a table, no division loop per digit.
struct DigitPairs {
    char t[200];
    DigitPairs() {
        for (int i = 0; i < 100; ++i) {
            t[2 * i] = (char)('0' + i / 10);
            t[2 * i + 1] = (char)('0' + i % 10);
        }
    }
};
static const DigitPairs g_pairs;

static inline unsigned u32_digits(uint32_t u) {
    return 1u + (u >= 10u) + (u >= 100u) + (u >= 1000u) + (u >= 10000u) + (u >= 100000u) + (u >= 1000000u) +
           (u >= 10000000u) + (u >= 100000000u) + (u >= 1000000000u);
}

static inline char* put_int(char* p, int32_t v) {
    uint32_t u = v < 0 ? (uint32_t)(-(int64_t)v) : (uint32_t)v;
    if (v < 0) *p++ = '-';
    char* const e = p + u32_digits(u);
    char* q = e;
    while (u >= 100) {
        const uint32_t r = u % 100;
        u /= 100;
        q -= 2;
        memcpy(q, g_pairs.t + 2 * r, 2);
    }
    if (u >= 10) memcpy(q - 2, g_pairs.t + 2 * u, 2);
    else q[-1] = (char)('0' + u);
    return e;
}

static inline unsigned int_chars(int32_t v) {
    const uint32_t u = v < 0 ? (uint32_t)(-(int64_t)v) : (uint32_t)v;
    return u32_digits(u) + (v < 0 ? 1u : 0u);
}

// One output file, written batch by batch. A regular file is extended and memory-mapped per batch so
// that the workers format straight into the page cache in parallel (concurrent write()s to one file
// serialise on the inode lock); anything else (pipe, /dev/null, BGR_OUT=write) gets plain write()s.
struct OutFile {
    int fd = -1;
    bool mappable = false;
    uint64_t pos = 0;
    // current batch
    char* map = nullptr;
    size_t map_len = 0;
    char* base = nullptr;  // where this batch's bytes start (inside map, or in `stage`)
    std::vector<char> stage;

    bool open(const std::string& path) {
        bool rw = true;
        fd = ::open(path.c_str(), O_RDWR | O_CREAT | O_TRUNC, 0666);  // a shared mapping needs read access too
        if (fd < 0) {
            rw = false;
            fd = ::open(path.c_str(), O_WRONLY | O_CREAT | O_TRUNC, 0666);
        }
        if (fd < 0) return false;
        struct stat st;
        const char* e = getenv("BGR_OUT");
        mappable = rw && fstat(fd, &st) == 0 && S_ISREG(st.st_mode) && !(e && !strcmp(e, "write"));
        return true;
    }
    bool begin(size_t bytes) {
        base = nullptr;
        if (bytes == 0) return true;
        if (fd >= 0 && mappable) {
            const uint64_t page = 4096, lo = pos & ~(page - 1);
            if (ftruncate(fd, (off_t)(pos + bytes)) == 0) {
                map_len = (size_t)(pos + bytes - lo);
                void* m = mmap(nullptr, map_len, PROT_READ | PROT_WRITE, MAP_SHARED, fd, (off_t)lo);
                if (m != MAP_FAILED) {
                    map = (char*)m;
                    base = map + (pos - lo);
                    return true;
                }
            }
            mappable = false;  // fall back to write() from here on
            if (ftruncate(fd, (off_t)pos) != 0 || lseek(fd, (off_t)pos, SEEK_SET) < 0) return false;
        }
        stage.resize(bytes);  // also the sink when the file could not be opened
        base = stage.data();
        return true;
    }
    bool end(size_t bytes) {
        if (fd < 0 || bytes == 0) return true;
        bool ok = true;
        if (map) {
            munmap(map, map_len);
            map = nullptr;
        } else {
            const char* p = stage.data();
            size_t left = bytes;
            while (left) {
                ssize_t w = ::write(fd, p, left);
                if (w <= 0) {
                    ok = false;
                    break;
                }
                p += w;
                left -= (size_t)w;
            }
        }
        pos += bytes;
        return ok;
    }
    void close() {
        if (fd >= 0) ::close(fd);
        fd = -1;
    }
};

enum { OUT_PATHS = 0, OUT_NOT_ALIGNED = 1, OUT_NO_OVERLAP = 2 };

// which file a read goes to (alignerGreedy.cpp:393-428; -o: see the header)
static inline int out_of(uint32_t flags, bool split_no_overlap) {
    if (flags & BGR_F_MAPPED) return OUT_PATHS;
    return (split_no_overlap && !(flags & BGR_F_OVERLAP_FOUND)) ? OUT_NO_OVERLAP : OUT_NOT_ALIGNED;
}

// bytes reads [lo, hi) of the batch put into each of the three files
static void measure_range(const Batch& b, size_t lo, size_t hi, bool split_no_overlap, bool correction, size_t out_sz[3]) {
    // summed in locals: the workers' result slots are neighbours in one array, and adding to them per read made
    // the cores fight over the cache line
    size_t sz[3] = {0, 0, 0};
    for (size_t i = lo; i < hi; ++i) {
        const bgr_result r = b.res.p[i];
        const uint32_t flags = BGR_RESULT_FLAGS(r), len = BGR_RESULT_LEN(r);
        const int o = out_of(flags, split_no_overlap);
        size_t n = b.pb.hdr.p[i].n + 2;
        if (o == OUT_PATHS && correction) {
            n += b.pb.soff.p[i + 1] - b.pb.soff.p[i];  // the corrected read is as long as the read
        } else if (o == OUT_PATHS) {
            const int32_t* ids = b.pool.p + r.path_off;
            for (uint32_t j = 0; j < len; ++j) n += int_chars(ids[j]) + 1;
        } else {
            n += b.pb.soff.p[i + 1] - b.pb.soff.p[i];
        }
        sz[o] += n;
    }
    out_sz[0] = sz[0];
    out_sz[1] = sz[1];
    out_sz[2] = sz[2];
}

// header\npath\n for mapped reads (printPath, aligner.cpp:600-609), header\nread\n otherwise
// 2-bit words -> characters, four bases per table entry
struct DecodeTable {
    char t[256][4];
    DecodeTable() {
        for (int v = 0; v < 256; ++v)
            for (int j = 0; j < 4; ++j) t[v][j] = "ACGT"[(v >> (6 - 2 * j)) & 3];
    }
};
static const DecodeTable g_decode;

static inline char* put_bases(char* p, const uint64_t* w, size_t n) {
    size_t i = 0;
    for (; i + 32 <= n; i += 32, ++w)
        for (int b = 7; b >= 0; --b, p += 4) memcpy(p, g_decode.t[(*w >> (8 * b)) & 0xFF], 4);
    for (; i < n; ++i) *p++ = "ACGT"[(*w >> (62 - 2 * (i & 31))) & 3];
    return p;
}

static void format_range(const Batch& b, size_t lo, size_t hi, bool split_no_overlap, bool correction, char* dst[3]) {
    for (size_t i = lo; i < hi; ++i) {
        const bgr_result r = b.res.p[i];
        const uint32_t flags = BGR_RESULT_FLAGS(r), len = BGR_RESULT_LEN(r);
        const int o = out_of(flags, split_no_overlap);
        char* p = dst[o];
        memcpy(p, b.pb.hdr.p[i].p, b.pb.hdr.p[i].n);
        p += b.pb.hdr.p[i].n;
        *p++ = '\n';
        if (o == OUT_PATHS && correction) {  // header\ncorrected\n (alignerGreedy.cpp:399)
            p = put_bases(p, b.corr.p + BGR_CORRECTED_BASE(b.pb.soff.p[i], i), b.pb.soff.p[i + 1] - b.pb.soff.p[i]);
        } else if (o == OUT_PATHS) {
            const int32_t* ids = b.pool.p + r.path_off;
            for (uint32_t j = 0; j < len; ++j) {
                p = put_int(p, ids[j]);
                *p++ = '.';
            }
        } else {
            const size_t sn = b.pb.soff.p[i + 1] - b.pb.soff.p[i];
            memcpy(p, b.pb.seqs.p + b.pb.soff.p[i], sn);
            p += sn;
        }
        *p++ = '\n';
        dst[o] = p;
    }
}

// the comma separated read file list (aligner.cpp:554-576)
static std::vector<std::string> split_list(const std::string& csv) {
    std::vector<std::string> files;
    size_t last = 0;
    for (size_t i = 0; i < csv.size(); ++i)
        if (csv[i] == ',') {
            files.push_back(csv.substr(last, i - last));
            last = i + 1;
        }
    files.push_back(csv.substr(last));
    return files;
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
    if (o.dog || (o.incomplete && o.brute)) {
        fprintf(stderr, "bgreat: -G and -b -i are outside the GPU mapping path and not supported\n");
        return 2;
    }
    // correctionMode only acts in alignPartGreedy (alignerGreedy.cpp:394); -b ignores it
    const bool correction = o.correction && !o.brute;
    // the reference keeps the thread count in an unsigned char (aligner.h:80)
    unsigned threads = (unsigned char)o.threads;
    if (threads == 0) threads = 1;

    // BGR_VERBOSE=1: phase timings on stderr (not part of the reference's output)
    const bool verbose = getenv("BGR_VERBOSE") != nullptr;
    const auto t_start = std::chrono::steady_clock::now();
    auto lap = [&](const char* what) {
        if (verbose)
            fprintf(stderr, "[bgreat] %-28s t=%.3f s\n", what,
                    std::chrono::duration<double>(std::chrono::steady_clock::now() - t_start).count());
    };
    const int n_slots = 2;
    // one context per visible device (reads shard across the GPUs of a box, the index is replicated)
    int n_dev = 0;
    if (bgr_device_count(&n_dev) != BGR_OK || n_dev < 1) {
        fprintf(stderr, "bgreat: %s\n", n_dev < 1 && !*bgr_last_error(nullptr) ? "no CUDA device" : bgr_last_error(nullptr));
        return 1;
    }
    // Devices used: all of them for large inputs, fewer for small ones (a CUDA context takes about a second to
    // create, and 4 GB of reads keep one device busy for about as long); BGR_GPUS=N says otherwise.
    size_t input_bytes = 0;
    for (const std::string& f : split_list(o.reads)) {
        struct stat st;
        if (stat(f.c_str(), &st) == 0 && S_ISREG(st.st_mode)) input_bytes += (size_t)st.st_size;
    }
    int n_gpus = (int)std::min<size_t>((size_t)n_dev, 1 + input_bytes / ((size_t)4 << 30));
    if (const char* e = getenv("BGR_GPUS")) n_gpus = std::max(1, std::min(16, atoi(e)));
    std::vector<bgr_ctx*> ctxs((size_t)n_gpus, nullptr);
    {
        // contexts are created side by side: the driver initialises the devices concurrently
        std::vector<std::thread> mk;
        std::vector<int> rcs((size_t)n_gpus, BGR_OK);
        std::vector<std::string> errs((size_t)n_gpus);
        for (int g = 0; g < n_gpus; ++g)
            mk.emplace_back([&, g] {
                rcs[g] = bgr_ctx_create(g % n_dev, n_slots, &ctxs[g]);
                if (rcs[g] != BGR_OK) errs[g] = bgr_last_error(nullptr);
            });
        for (auto& t : mk) t.join();
        for (int g = 0; g < n_gpus; ++g)
            if (rcs[g] != BGR_OK) {
                fprintf(stderr, "bgreat: %s\n", errs[g].c_str());
                return 1;
            }
    }
    bgr_ctx* ctx = ctxs[0];

    lap("context created");
    // Aligner ctor opens the outputs first (aligner.h:85-86)
    OutFile out[3];
    out[OUT_PATHS].open(o.path_file);
    out[OUT_NOT_ALIGNED].open(o.not_aligned_file);
    if (o.have_o) out[OUT_NO_OVERLAP].open(o.no_overlap_file);

    // ---- indexUnitigs (aligner.cpp:537-547)
    auto t0 = std::chrono::system_clock::now();
    // BGR_INDEX_CACHE=FILE: the index blob is taken from FILE when it is there and fits this k, and written to
    // FILE after a build otherwise (keeping FILE in step with the unitig file is the user's business)
    const char* cache = getenv("BGR_INDEX_CACHE");
    int rc = (cache && *cache) ? bgr_index_load_file(ctx, cache, (uint32_t)o.k) : BGR_ERR_IO;
    if (rc != BGR_OK) {
        if (cache && *cache && rc != BGR_ERR_IO) fprintf(stderr, "bgreat: index cache not used: %s\n", bgr_last_error(ctx));
        rc = bgr_index_build_file(ctx, o.unitigs.c_str(), (uint32_t)o.k);
        if (rc != BGR_OK && rc != BGR_ERR_IO) {  // a missing unitig file means an empty index, like the reference
            fprintf(stderr, "bgreat: %s\n", bgr_last_error(ctx));
            return 1;
        }
        if (rc == BGR_OK && cache && *cache && bgr_index_save_file(ctx, cache) != BGR_OK)
            fprintf(stderr, "bgreat: %s\n", bgr_last_error(ctx));
        rc = BGR_OK;
    }
    {
        std::vector<std::thread> rep;
        std::vector<int> rcs((size_t)n_gpus, BGR_OK);
        for (int g = 1; g < n_gpus; ++g) rep.emplace_back([&, g] { rcs[g] = bgr_index_replicate(ctxs[g], ctx); });
        for (auto& t : rep) t.join();
        for (int g = 1; g < n_gpus; ++g)
            if (rcs[g] != BGR_OK) {
                fprintf(stderr, "bgreat: %s\n", bgr_last_error(ctxs[g]));
                return 1;
            }
    }
    auto t1 = std::chrono::system_clock::now();
    std::cout << "Indexing in seconds : " << std::chrono::duration_cast<std::chrono::seconds>(t1 - t0).count()
              << std::endl;

    lap("index built");
    // ---- alignAll (aligner.cpp:550-597)
    auto start = std::chrono::system_clock::now();
    bgr_params params;
    params.mode = o.brute ? BGR_MODE_EXHAUSTIVE : BGR_MODE_GREEDY;
    params.errors = (uint32_t)o.errors;
    params.effort = (uint32_t)o.effort;
    params.flags = correction ? BGR_P_CORRECT : 0u;

    const size_t n_batches = (size_t)n_gpus * (n_slots + 1) + 3;
    std::vector<Batch> batches(n_batches);
    Queue<Batch*> free_q, parsed_q, done_q;
    // text per batch: about 100 k reads of 150 bases (measured on 20 M reads: 16-32 MB beats 8 and 64 MB: the
    // pipeline fills sooner and there is less memory to pin); BGR_RANGE_BYTES overrides (tests cut small files)
    size_t range_bytes = 16u << 20;
    if (const char* e = getenv("BGR_RANGE_BYTES")) range_bytes = std::max<size_t>(1, strtoull(e, nullptr, 10));
    // the batches' pinned buffers are allocated beside the pipeline: the reader starts on the first batch while
    // the others are still being pinned
    size_t total_text = 0;
    for (const std::string& f : split_list(o.reads)) {
        struct stat st;
        if (stat(f.c_str(), &st) == 0 && S_ISREG(st.st_mode)) total_text += (size_t)st.st_size;
        else total_text += range_bytes;  // pipes: size unknown
    }
    const size_t batch_text = std::min(range_bytes, total_text + 4096);
    // regular FASTA / FASTQ text goes through the device's parser and formatter (BGR_TEXT=0: the host stages only)
    const bool text_path = !(getenv("BGR_TEXT") && !strcmp(getenv("BGR_TEXT"), "0"));
    std::thread allocator([&] {
        for (auto& b : batches) {
            if (text_path) b.prealloc_text(batch_text, o.have_o);
            else b.prealloc(batch_text);
            free_q.push(&b);
        }
    });
    bool io_error = false;
    double parse_s = 0, gpu_wait_s = 0, write_s = 0;  // busy time per stage (BGR_VERBOSE)
    double reader_phase[3] = {0, 0, 0}, writer_phase[4] = {0, 0, 0, 0}, free_wait_s = 0, done_wait_s = 0;
    auto now = [] { return std::chrono::steady_clock::now(); };
    auto since = [&](std::chrono::steady_clock::time_point t) { return std::chrono::duration<double>(now() - t).count(); };

    // files of the comma separated list, in order (aligner.cpp:554-576)
    const std::vector<std::string> files = split_list(o.reads);
    // the header lines of a batch are views into the mapped input: every file stays mapped until the end
    std::vector<std::unique_ptr<MappedFile>> mapped;

    // cleared by the first range the device reports irregular: the rest of the input is then parsed on the host
    std::atomic<bool> text_ok{true};
    std::thread parser([&] {
        uint64_t seq = 0;
        WorkerPool copy_pool(threads - 1);
        for (const std::string& file : files) {
            mapped.emplace_back(new MappedFile());
            MappedFile& mf = *mapped.back();
            if (!mf.open(file.c_str())) {
                fprintf(stderr, "bgreat: cannot open read file %s\n", file.c_str());
                io_error = true;
                continue;
            }
            // alignPartGreedy's loop (alignerGreedy.cpp:372-377): getReads(10000) until eof
            ParallelReader pr(mf.data, mf.size, (uint32_t)o.k, o.fastq, threads, range_bytes);
            do {
                const auto tw = now();
                Batch* b = free_q.pop();
                free_wait_s += since(tw);
                const auto t = now();
                b->pb.clear();
                b->text_mode = false;
                size_t ra = 0, rb = 0;
                bool final = false, have = false;
                std::vector<size_t> cuts;
                if (!text_path) {
                    pr.next(b->pb);
                } else if (!o.fastq) {
                    have = pr.fasta_next_range(ra, rb);
                } else if (pr.fastq_tail_pending()) {
                    pr.fastq_tail(b->pb);  // the end of the stream: the sequential parser, as the reference sees it
                } else {
                    have = pr.fastq_next_range(ra, rb, final, &cuts);
                }
                if (have) {
                    if (text_ok.load()) {  // hand the text over as it is: copied into the pinned buffer by all workers
                        b->text_mode = true;
                        b->src = mf.data;
                        b->src_a = ra;
                        b->src_b = rb;
                        const size_t nb = rb - ra;
                        b->text.reserve(nb + 64);
                        b->text.n = nb;
                        const unsigned T = (unsigned)std::max<size_t>(1, std::min<size_t>(threads, nb / (1u << 20) + 1));
                        copy_pool.run(T, [&](unsigned w) {
                            const size_t lo = nb * w / T, hi = nb * (w + 1) / T;
                            memcpy(b->text.p + lo, mf.data + ra + lo, hi - lo);
                        });
                    } else if (o.fastq) {
                        pr.fastq_parse_range(ra, rb, b->pb, &cuts);
                    } else {
                        pr.fasta_parse_range(ra, rb, b->pb);
                    }
                }
                b->seq = seq++;
                parse_s += since(t);
                parsed_q.push(b);  // possibly empty: still flows through so that order is kept
            } while (!pr.done());
            for (int i = 0; i < 3; ++i) reader_phase[i] += pr.phase_s[i];
        }
        for (int g = 0; g < n_gpus; ++g) parsed_q.push(nullptr);
    });

    std::mutex stat_m;
    auto gpu_loop = [&](bgr_ctx* ctx) {
        std::deque<Batch*> inflight;
        int next_slot = 0;
        bool input_done = false;
        auto finish = [&](Batch* b) {
            uint64_t used = 0;
            const auto t = now();
            if (b->text_mode) {
                char* outp[3] = {b->otext[0].p, b->otext[1].p, b->otext[2].p};
                uint64_t cap[3] = {b->otext[0].cap, b->otext[1].cap, b->otext[2].cap};
                int r = bgr_slot_format_text(ctx, b->slot, o.have_o ? 1 : 0, outp, cap, b->oused);
                if (r == BGR_ERR_LIMIT) {  // an output buffer too small: come back with bigger ones
                    for (int f = 0; f < 3; ++f) {
                        if (b->oused[f] > b->otext[f].cap) b->otext[f].reserve(b->oused[f] + b->oused[f] / 8 + 4096);
                        outp[f] = b->otext[f].p;
                        cap[f] = b->otext[f].cap;
                    }
                    r = bgr_slot_format_text(ctx, b->slot, o.have_o ? 1 : 0, outp, cap, b->oused);
                }
                if (r != BGR_OK) {
                    fprintf(stderr, "bgreat: %s\n", bgr_last_error(ctx));
                    exit(1);
                }
                {
                    std::lock_guard<std::mutex> l(stat_m);
                    gpu_wait_s += since(t);
                }
                done_q.push(b);
                return;
            }
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
            {
                std::lock_guard<std::mutex> l(stat_m);
                gpu_wait_s += since(t);
            }
            done_q.push(b);
        };
        while (!input_done || !inflight.empty()) {
            // take the next parsed batch if a slot is free -- but only wait for one when nothing is in flight:
            // otherwise a finished batch would sit here, kept from the writer, until the reader delivers
            Batch* b = nullptr;
            bool got = false;
            if (!input_done && (int)inflight.size() < n_slots) {
                if (inflight.empty()) {
                    b = parsed_q.pop();
                    got = true;
                } else {
                    got = parsed_q.try_pop(b);
                }
            }
            if (got) {
                if (!b) {
                    input_done = true;
                    continue;
                }
                b->slot = next_slot;
                next_slot = (next_slot + 1) % n_slots;
                if (b->text_mode) {
                    uint32_t n_reads = 0;
                    int regular = 1;
                    int r = bgr_slot_upload_text(ctx, b->slot, b->text.p, b->text.n, o.fastq ? 1 : 0, &n_reads, &regular);
                    if (r == BGR_OK && regular) r = bgr_slot_run(ctx, b->slot, &params);
                    if (r != BGR_OK) {
                        fprintf(stderr, "bgreat: %s\n", bgr_last_error(ctx));
                        exit(1);
                    }
                    if (regular) {
                        inflight.push_back(b);
                        continue;
                    }
                    // continuation lines, a missing final newline ...: the host parser knows what the reference
                    // does with those; this range here and now, the rest of the input in the reader
                    text_ok.store(false);
                    b->text_mode = false;
                    b->pb.clear();
                    if (o.fastq) parse_fastq_text_sequential(b->src, b->src_a, b->src_b, b->pb);
                    else parse_fasta_text_sequential(b->src, b->src_a, b->src_b, (uint32_t)o.k, b->pb);
                }
                size_t n = b->pb.size();
                b->res.reserve(n + 1);
                b->pool.reserve(n * 8 + 1024);
                int r = bgr_slot_upload(ctx, b->slot, b->pb.seqs.p, b->pb.soff.p, (uint32_t)n);
                if (r == BGR_OK) r = bgr_slot_run(ctx, b->slot, &params);
                if (r == BGR_OK) r = bgr_slot_download(ctx, b->slot, b->res.p, b->pool.p, b->pool.cap);
                if (r == BGR_OK && correction) {
                    b->corr.reserve(BGR_CORRECTED_WORDS(b->pb.seqs.n, n));
                    r = bgr_slot_download_corrected(ctx, b->slot, b->corr.p, b->corr.cap);
                }
                if (r != BGR_OK) {
                    fprintf(stderr, "bgreat: %s\n", bgr_last_error(ctx));
                    exit(1);
                }
                inflight.push_back(b);
            } else {
                Batch* oldest = inflight.front();
                inflight.pop_front();
                finish(oldest);
            }
        }
        done_q.push(nullptr);
    };
    std::vector<std::thread> gpu_threads;
    for (int g = 0; g < n_gpus; ++g) gpu_threads.emplace_back(gpu_loop, ctxs[g]);

    // writer: this thread and its workers; batches leave in input order (a small reorder buffer: the devices
    // finish in any order)
    bool write_error = false;
    {
        std::vector<size_t> sz(3 * (size_t)threads), at(3 * (size_t)threads);
        WorkerPool pool(threads - 1);
        std::vector<Batch*> held;
        uint64_t next_seq = 0;
        int ended = 0;
        for (;;) {
            Batch* b = nullptr;
            for (size_t i = 0; i < held.size(); ++i)
                if (held[i]->seq == next_seq) {
                    b = held[i];
                    held.erase(held.begin() + (long)i);
                    break;
                }
            if (!b) {
                if (ended == n_gpus) break;  // every device thread is through and nothing is held back
                const auto tw = now();
                Batch* x = done_q.pop();
                done_wait_s += since(tw);
                if (!x) ++ended;
                else held.push_back(x);
                continue;
            }
            ++next_seq;
            if (b->text_mode) {  // already text: the three streams are copied into the files by all workers
                const auto t0w = now();
                for (int f = 0; f < 3; ++f)
                    if (!out[f].begin((size_t)b->oused[f])) write_error = true;
                writer_phase[1] += since(t0w);
                const auto t2w = now();
                size_t total_b = (size_t)(b->oused[0] + b->oused[1] + b->oused[2]);
                const unsigned T = (unsigned)std::max<size_t>(1, std::min<size_t>(threads, total_b / (1u << 20) + 1));
                pool.run(T, [&](unsigned t) {
                    for (int f = 0; f < 3; ++f) {
                        if (!out[f].base || !b->oused[f]) continue;
                        const size_t lo = (size_t)b->oused[f] * t / T, hi = (size_t)b->oused[f] * (t + 1) / T;
                        memcpy(out[f].base + lo, b->otext[f].p + lo, hi - lo);
                    }
                });
                writer_phase[2] += since(t2w);
                const auto t3w = now();
                for (int f = 0; f < 3; ++f)
                    if (!out[f].end((size_t)b->oused[f])) write_error = true;
                writer_phase[3] += since(t3w);
                write_s += since(t0w);
                free_q.push(b);
                continue;
            }
            const auto t0w = now();
            const size_t n = b->pb.size();
            const unsigned T = (unsigned)std::min<size_t>(threads, n / 4096 + 1);
            pool.run(T, [&](unsigned t) { measure_range(*b, n * t / T, n * (t + 1) / T, o.have_o, correction, &sz[3 * t]); });
            writer_phase[0] += since(t0w);
            const auto t1w = now();
            size_t total[3] = {0, 0, 0};
            for (unsigned t = 0; t < T; ++t)
                for (int f = 0; f < 3; ++f) {
                    at[3 * t + f] = total[f];
                    total[f] += sz[3 * t + f];
                }
            for (int f = 0; f < 3; ++f)
                if (!out[f].begin(total[f])) write_error = true;
            writer_phase[1] += since(t1w);
            const auto t2w = now();
            pool.run(T, [&](unsigned t) {
                char* dst[3];
                for (int f = 0; f < 3; ++f) dst[f] = out[f].base ? out[f].base + at[3 * t + f] : nullptr;
                format_range(*b, n * t / T, n * (t + 1) / T, o.have_o, correction, dst);
            });
            writer_phase[2] += since(t2w);
            const auto t3w = now();
            for (int f = 0; f < 3; ++f)
                if (!out[f].end(total[f])) write_error = true;
            writer_phase[3] += since(t3w);
            write_s += since(t0w);
            free_q.push(b);
        }
    }
    parser.join();
    for (auto& t : gpu_threads) t.join();
    allocator.join();
    lap("mapping done");
    if (verbose) {
        fprintf(stderr, "[bgreat] busy: reader %.3f s, waiting for the GPU %.3f s, writer %.3f s\n", parse_s, gpu_wait_s,
                write_s);
        fprintf(stderr, "[bgreat] reader: cut %.3f parse %.3f copy %.3f, waited for a free batch %.3f s\n", reader_phase[0],
                reader_phase[1], reader_phase[2], free_wait_s);
        fprintf(stderr, "[bgreat] writer: measure %.3f open %.3f format %.3f close %.3f, waited for results %.3f s\n",
                writer_phase[0], writer_phase[1], writer_phase[2], writer_phase[3], done_wait_s);
    }
    for (const std::string& file : files) std::cout << file << std::endl;

    uint64_t cnt[BGR_C_COUNT] = {0};
    for (bgr_ctx* c : ctxs) {  // the reference's atomic counters (aligner.h:68), summed over the devices
        uint64_t one[BGR_C_COUNT];
        bgr_counters(c, one);
        for (int i = 0; i < BGR_C_COUNT; ++i) cnt[i] += one[i];
    }
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

    for (int f = 0; f < 3; ++f) out[f].close();
    if (write_error) fprintf(stderr, "bgreat: error while writing the output files\n");
    lap("outputs closed");
    const int status = (io_error || write_error) ? 1 : 0;
    if (!getenv("BGR_TEARDOWN")) {
        // Everything is written; unpinning the buffers, unmapping the inputs and tearing the CUDA context down
        // takes 0.1 - 3 s that the operating system does faster (the reference leaves its files to exit() too,
        // aligner.h:85). BGR_TEARDOWN=1 runs the orderly path (leak checkers).
        std::cout.flush();
        fflush(nullptr);
        _exit(status);
    }
    batches.clear();
    mapped.clear();
    for (bgr_ctx* c : ctxs) bgr_ctx_destroy(c);
    lap("context destroyed");
    return status;
}
