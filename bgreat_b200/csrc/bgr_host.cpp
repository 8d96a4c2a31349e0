// Host side of the mapping path: reader semantics, unitig loading, index blob construction.
#include "bgr_host.h"

#include <fcntl.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#if defined(__x86_64__)
#include <immintrin.h>
#endif

#include <algorithm>
#include <chrono>
#include <condition_variable>
#include <deque>
#include <mutex>
#include <thread>

#include "../../include/bgreat_b200.h"

namespace bgr {

// ------------------------------------------------------------------------------ LineStream

void LineStream::getline(std::string& s) {
    if (!good()) {  // sentry fails: failbit, string untouched
        fail_ = true;
        return;
    }
    s.clear();
    const char* p;
    size_t len;
    if (getline_view(p, len)) s.assign(p, len);
}

bool LineStream::getline_view(const char*& p, size_t& len) {
    if (!good()) {
        fail_ = true;
        return false;
    }
    p = d_ + pos_;
    if (pos_ >= n_) {  // nothing extracted at all: eofbit and failbit, string erased
        eof_ = true;
        fail_ = true;
        len = 0;
        return true;
    }
    const char* nl = (const char*)memchr(p, '\n', n_ - pos_);
    if (nl) {
        len = (size_t)(nl - p);
        pos_ += len + 1;
    } else {  // ran into the end of input: eofbit only
        len = n_ - pos_;
        pos_ = n_;
        eof_ = true;
    }
    return true;
}

int LineStream::peek() {
    if (!good()) return -1;
    if (pos_ >= n_) {
        eof_ = true;
        return -1;
    }
    return (unsigned char)d_[pos_];
}

// ------------------------------------------------------------------------------ MappedFile

bool MappedFile::open(const char* path) {
    close();
    fd = ::open(path, O_RDONLY);
    if (fd < 0) return false;
    struct stat st;
    if (fstat(fd, &st) != 0 || !S_ISREG(st.st_mode)) {
        // not a regular file (pipe, /dev/stdin ...): slurp it
        std::string* buf = new std::string();
        char tmp[1 << 16];
        ssize_t r;
        while ((r = ::read(fd, tmp, sizeof tmp)) > 0) buf->append(tmp, (size_t)r);
        ::close(fd);
        fd = -2;
        data = buf->data();
        size = buf->size();
        ok = true;
        // ownership of buf is kept until close(); stash the pointer behind data via a side table
        // (simple: leak-free because close() reconstructs it from data)
        slurped_ = buf;
        return true;
    }
    size = (size_t)st.st_size;
    if (size == 0) {
        data = "";
        ok = true;
        return true;
    }
    void* m = mmap(nullptr, size, PROT_READ, MAP_PRIVATE, fd, 0);
    if (m == MAP_FAILED) {
        ::close(fd);
        fd = -1;
        return false;
    }
    madvise(m, size, MADV_SEQUENTIAL);
    data = (const char*)m;
    ok = true;
    return true;
}

void MappedFile::close() {
    if (slurped_) {
        delete (std::string*)slurped_;
        slurped_ = nullptr;
    } else if (ok && size > 0 && data) {
        munmap((void*)data, size);
    }
    if (fd >= 0) ::close(fd);
    fd = -1;
    data = nullptr;
    size = 0;
    ok = false;
}

// ------------------------------------------------------------------------------ unitigs

bool load_unitig_file(const char* path, uint32_t k, std::string& seqs, std::vector<uint64_t>& off) {
    seqs.clear();
    off.assign(1, 0);
    MappedFile f;
    bool opened = f.open(path);
    LineStream s(f.data, f.size, opened);
    // Both getlines of a record write the same string in the reference, so a failed second
    // getline leaves the HEADER text in it: keep one string to reproduce that.
    std::string line;
    while (!s.eof()) {
        s.getline(line);
        s.getline(line);
        if (line.size() < k) break;
        seqs += line;
        off.push_back(seqs.size());
    }
    return opened;
}

static inline u64 pool_bases(const u64* pool, u64 p, u32 n) {
    // n (1..32) bases starting at base p, right-aligned (str2num value)
    u64 w = p >> 5;
    u32 s = (u32)(p & 31) * 2;
    u64 hi = pool[w], lo = pool[w + 1];
    u64 x = (hi << s) | ((lo >> 1) >> (63 - s));
    return x >> (64 - 2 * n);
}

namespace {
struct KeyRef {
    u64 key;
    u32 id;       // unitig id
    u32 side_flg;  // bit0: 1 = right table; bits 30/31: orientation flags for the slot
};
}  // namespace

int build_index_blob(const char* seqs, const uint64_t* off, uint32_t n, uint32_t k, std::vector<uint64_t>& blob,
                     std::string& err) {
    if (k < 2 || k > 32) {
        err = "k must be in [2,32] (kmer is a 64-bit word in the reference, utils.h:27)";
        return BGR_ERR_ARG;
    }
    // accepted unitigs: up to the first one shorter than k (aligner.cpp:418)
    u64 nu = 0, total_bases = 0;
    for (; nu < n; ++nu) {
        u64 len = off[nu + 1] - off[nu];
        if (len < k) break;
        if (len >> SLOT_LEN_BITS) {
            err = "unitig longer than 2^28-1 bases";
            return BGR_ERR_LIMIT;
        }
        total_bases += len;
    }
    if (nu >= SLOT_ID_MASK) {
        err = "more than 2^30-1 unitigs";
        return BGR_ERR_LIMIT;
    }
    if (total_bases >> SLOT_START_BITS) {
        err = "more than 2^36 unitig bases";
        return BGR_ERR_LIMIT;
    }
    const u64 pool_words = (total_bases + 31) / 32 + 2;  // + lookahead padding

    // ---- pack + per-unitig border codes
    std::vector<u64> pool(pool_words, 0);
    std::vector<UnitigMeta> meta(nu + 1);
    meta[0].start = 0;
    meta[0].len = 0;
    meta[0].pad = 0;
    u64 base = 0;
    for (u64 i = 0; i < nu; ++i) {
        const char* s = seqs + off[i];
        u32 len = (u32)(off[i + 1] - off[i]);
        meta[i + 1].start = base;
        meta[i + 1].len = len;
        meta[i + 1].pad = 0;
        for (u32 j = 0; j < len; ++j) {
            u32 c = base_code((unsigned char)s[j]);
            if (c > 3) {
                err = "unitig " + std::to_string(i + 1) + " contains a character outside ACGT";
                return BGR_ERR_FORMAT;
            }
            u64 p = base + j;
            pool[p >> 5] |= (u64)c << (62 - 2 * (p & 31));
        }
        base += len;
    }

    // ---- key occurrences in file order (aligner.cpp:466-533)
    std::vector<KeyRef> refs;
    refs.reserve(2 * nu);
    for (u64 i = 1; i <= nu; ++i) {
        u64 beg = pool_bases(pool.data(), meta[i].start, k - 1);
        u64 end = pool_bases(pool.data(), meta[i].start + meta[i].len - (k - 1), k - 1);
        u64 rcbeg = rcb(beg, k - 1), rcend = rcb(end, k - 1);
        KeyRef a, b;
        a.id = b.id = (u32)i;
        if (beg <= rcbeg) {  // left[beg]
            a.key = beg;
            a.side_flg = 0u | SLOT_FWD_CANON /* begins with key */ | (rcend == beg ? SLOT_FWD_RC : 0u);
        } else {  // right[rcbeg]
            a.key = rcbeg;
            a.side_flg = 1u | SLOT_FWD_RC /* rc(begin) == key */ | (end == rcbeg ? SLOT_FWD_CANON : 0u);
        }
        if (end <= rcend) {  // right[end]
            b.key = end;
            b.side_flg = 1u | SLOT_FWD_CANON /* ends with key */ | (rcbeg == end ? SLOT_FWD_RC : 0u);
        } else {  // left[rcend]
            b.key = rcend;
            b.side_flg = 0u | SLOT_FWD_RC /* rc(end) == key */ | (beg == rcend ? SLOT_FWD_CANON : 0u);
        }
        refs.push_back(a);
        refs.push_back(b);
    }

    // ---- distinct keys
    std::vector<u64> keys(refs.size());
    for (size_t i = 0; i < refs.size(); ++i) keys[i] = refs[i].key;
    std::sort(keys.begin(), keys.end());
    keys.erase(std::unique(keys.begin(), keys.end()), keys.end());
    const u64 n_keys = keys.size();
    const u64 n_buckets = std::max<u64>(4, (n_keys + 1) / 2);  // <= 50% load: 2 of 4 slots on average
    if (n_buckets * 4 > SLOT_NEXT_MASK) {
        err = "too many distinct border k-mers for 30-bit key slots";
        return BGR_ERR_LIMIT;
    }

    // ---- blob
    BlobHeader h;
    memset(&h, 0, sizeof h);
    h.magic = BLOB_MAGIC;
    h.version = BLOB_VERSION;
    h.k = k;
    h.n_unitigs = nu;
    h.n_keys = n_keys;
    h.n_buckets = n_buckets;
    h.pool_words = pool_words;
    auto align_up = [](u64 x) { return (x + 255) & ~255ull; };
    h.off_keys = align_up(sizeof(BlobHeader));
    h.off_entries = align_up(h.off_keys + n_buckets * 4 * sizeof(u64));
    h.off_meta = align_up(h.off_entries + n_buckets * 4 * sizeof(Entry));
    h.off_pool = align_up(h.off_meta + (nu + 1) * sizeof(UnitigMeta));
    u32 bloom_mode = 0;
    h.n_bloom = bloom_layout(n_keys, k, bloom_mode);
    h.bloom_mode = bloom_mode;
    if (h.n_bloom >= 0xFFFFFFFFull) {
        err = "too many distinct border k-mers for the 32-bit filter hash";
        return BGR_ERR_LIMIT;
    }
    h.off_bloom = align_up(h.off_pool + pool_words * sizeof(u64));
    h.total_bytes = align_up(h.off_bloom + h.n_bloom * sizeof(u32));
    blob.assign(h.total_bytes / 8, 0);
    char* B = (char*)blob.data();
    memcpy(B, &h, sizeof h);
    u64* tkeys = (u64*)(B + h.off_keys);
    Entry* tent = (Entry*)(B + h.off_entries);
    memcpy(B + h.off_meta, meta.data(), (nu + 1) * sizeof(UnitigMeta));
    memcpy(B + h.off_pool, pool.data(), pool_words * sizeof(u64));
    for (u64 i = 0; i < n_buckets * 4; ++i) tkeys[i] = EMPTY_KEY;

    auto find_or_insert = [&](u64 key, bool insert) -> u64 {
        u64 b = bucket_of(key, n_buckets);
        for (;;) {
            u64* bk = tkeys + b * 4;
            for (int j = 0; j < 4; ++j) {
                if (bk[j] == key) return b * 4 + j;
                if (bk[j] == EMPTY_KEY) {
                    if (!insert) return ~0ull;
                    bk[j] = key;
                    return b * 4 + j;
                }
            }
            if (++b == n_buckets) b = 0;
        }
    };
    for (u64 i = 0; i < n_keys; ++i) find_or_insert(keys[i], true);

    // ---- Bloom filter of the seed scan: every key sets its three bits
    {
        memcpy(B, &h, sizeof h);
        u32* bloom = (u32*)(B + h.off_bloom);
        for (u64 i = 0; i < n_keys; ++i) {
            if (bloom_mode == BLOOM_MINIMIZER) {
                u32 block;
                u64 pattern;
                bloom_slot_mm(keys[i], (u32)(h.n_bloom / 2), block, pattern);
                ((u64*)bloom)[block] |= pattern;
            } else {
                u32 word, pattern;
                bloom_slot(keys[i], k - 1, bloom_mode == BLOOM_WORDS_2HASH, (u32)h.n_bloom, word, pattern);
                bloom[word] |= pattern;
            }
        }
    }

    // ---- slot fill in file order: first free of 4, the 5th and later overwrite slot 4
    // (aligner.cpp:481-489)
    for (size_t i = 0; i < refs.size(); ++i) {
        u64 slot = find_or_insert(refs[i].key, false);
        const bool right = (refs[i].side_flg & 1u) != 0;
        USlot* s = right ? tent[slot].right : tent[slot].left;
        USlot v;
        v.start_len = meta[refs[i].id].start | ((u64)meta[refs[i].id].len << SLOT_START_BITS);
        v.id_flags = refs[i].id | (refs[i].side_flg & (SLOT_FWD_CANON | SLOT_FWD_RC));
        // the entry the walk goes on with: the key slot of the (k-1)-mer at the unitig's other end
        const u64 c = slot_far_code(pool.data(), meta[refs[i].id].start, meta[refs[i].id].len, k - 1, v.id_flags, right);
        const u64 rc = rcb(c, k - 1);
        v.next = (u32)find_or_insert(c <= rc ? c : rc, false) | (c <= rc ? SLOT_NEXT_CANON : 0u) | (c == rc ? SLOT_NEXT_PAL : 0u);
        if (s[0].id_flags == 0) s[0] = v;
        else if (s[1].id_flags == 0) s[1] = v;
        else if (s[2].id_flags == 0) s[2] = v;
        else s[3] = v;
    }
    return BGR_OK;
}

// ------------------------------------------------------------------------------ reads

// Read validation touches every input byte once: 32 bytes per step with AVX2 where the host has it
// (checked at run time: the binary is built on another machine), else a table walk.
namespace {
struct AcgtnTable {
    unsigned char bad[256];
    AcgtnTable() {
        memset(bad, 1, sizeof bad);
        bad[(unsigned char)'A'] = bad[(unsigned char)'C'] = bad[(unsigned char)'G'] = bad[(unsigned char)'T'] =
            bad[(unsigned char)'N'] = 0;
    }
};
const AcgtnTable g_acgtn;

bool acgtn_only_scalar(const char* p, size_t n) {
    const unsigned char* u = (const unsigned char*)p;
    unsigned bad = 0;
    size_t i = 0;
    for (; i + 8 <= n; i += 8)
        bad |= g_acgtn.bad[u[i]] | g_acgtn.bad[u[i + 1]] | g_acgtn.bad[u[i + 2]] | g_acgtn.bad[u[i + 3]] |
               g_acgtn.bad[u[i + 4]] | g_acgtn.bad[u[i + 5]] | g_acgtn.bad[u[i + 6]] | g_acgtn.bad[u[i + 7]];
    for (; i < n; ++i) bad |= g_acgtn.bad[u[i]];
    return bad == 0;
}

#if defined(__x86_64__) && defined(__GNUC__)
#define BGR_HAVE_AVX2_PATH 1
__attribute__((target("avx2"))) bool acgtn_only_avx2(const char* p, size_t n) {
    const __m256i cA = _mm256_set1_epi8('A'), cC = _mm256_set1_epi8('C'), cG = _mm256_set1_epi8('G'),
                  cT = _mm256_set1_epi8('T'), cN = _mm256_set1_epi8('N');
    __m256i all = _mm256_set1_epi8((char)0xFF);
    size_t i = 0;
    for (; i + 32 <= n; i += 32) {
        const __m256i x = _mm256_loadu_si256((const __m256i*)(p + i));
        const __m256i ok = _mm256_or_si256(
            _mm256_or_si256(_mm256_or_si256(_mm256_cmpeq_epi8(x, cA), _mm256_cmpeq_epi8(x, cC)),
                            _mm256_or_si256(_mm256_cmpeq_epi8(x, cG), _mm256_cmpeq_epi8(x, cT))),
            _mm256_cmpeq_epi8(x, cN));
        all = _mm256_and_si256(all, ok);
    }
    if (i < n && n >= 32) {  // the last, overlapping 32 bytes
        const __m256i x = _mm256_loadu_si256((const __m256i*)(p + n - 32));
        const __m256i ok = _mm256_or_si256(
            _mm256_or_si256(_mm256_or_si256(_mm256_cmpeq_epi8(x, cA), _mm256_cmpeq_epi8(x, cC)),
                            _mm256_or_si256(_mm256_cmpeq_epi8(x, cG), _mm256_cmpeq_epi8(x, cT))),
            _mm256_cmpeq_epi8(x, cN));
        all = _mm256_and_si256(all, ok);
        i = n;
    }
    if (_mm256_movemask_epi8(all) != -1) return false;
    return acgtn_only_scalar(p + i, n - i);
}
#endif
}  // namespace

bool acgtn_only(const char* p, size_t n) {
#ifdef BGR_HAVE_AVX2_PATH
    static const bool avx2 = __builtin_cpu_supports("avx2");
    if (avx2) return acgtn_only_avx2(p, n);
#endif
    return acgtn_only_scalar(p, n);
}
static inline bool acgtn_only(const std::string& r) { return acgtn_only(r.data(), r.size()); }

void ReadParser::next_chunk(uint32_t n, ReadBatch& out) {
    // The three strings are locals of getReads: fresh on every call (this is what makes the FASTQ
    // phantom read depend on where the 10000-read chunk boundaries fall).
    std::string read, header, inter;
    if (fastq_) {
        for (uint32_t i = 0; i < n; ++i) {
            s_.getline(header);
            s_.getline(read);
            if (read.size() > 2 && acgtn_only(read)) out.push(header.data(), header.size(), read.data(), read.size());
            s_.getline(header);
            s_.getline(header);
            if (s_.eof()) return;
        }
    } else {
        for (uint32_t i = 0; i < n; ++i) {
            s_.getline(header);
            s_.getline(read);
            for (;;) {
                int c = s_.peek();
                if (c == '>') {
                    if (read.size() > 2 && acgtn_only(read) && read.size() > k_)
                        out.push(header.data(), header.size(), read.data(), read.size());
                    read.clear();  // aligner.cpp:93: a following getline that fails must find it empty
                    break;
                } else if (!s_.eof()) {
                    s_.getline(inter);
                    read += inter;
                } else {
                    if (read.size() > 2 && acgtn_only(read) && read.size() > k_)
                        out.push(header.data(), header.size(), read.data(), read.size());
                    return;
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------ parallel reader

struct WorkerPool::State {
    std::mutex m;
    std::condition_variable cv_job, cv_done;
    std::vector<std::thread> th;
    const std::function<void(unsigned)>* fn = nullptr;
    unsigned T = 0, next = 0, pending = 0;
    uint64_t epoch = 0;
    bool stop = false;
    void work(std::unique_lock<std::mutex>& l) {  // take indices until the job is used up
        while (next < T) {
            const unsigned t = next++;
            const std::function<void(unsigned)>* f = fn;
            l.unlock();
            (*f)(t);
            l.lock();
            if (--pending == 0) cv_done.notify_all();
        }
    }
};

WorkerPool::WorkerPool(unsigned workers) : st_(new State()) {
    for (unsigned i = 0; i < workers; ++i)
        st_->th.emplace_back([s = st_] {
            std::unique_lock<std::mutex> l(s->m);
            uint64_t seen = 0;
            for (;;) {
                s->cv_job.wait(l, [&] { return s->stop || s->epoch != seen; });
                if (s->stop) return;
                seen = s->epoch;
                s->work(l);
            }
        });
}

WorkerPool::~WorkerPool() {
    {
        std::unique_lock<std::mutex> l(st_->m);
        st_->stop = true;
    }
    st_->cv_job.notify_all();
    for (auto& t : st_->th) t.join();
    delete st_;
}

void WorkerPool::run(unsigned T, const std::function<void(unsigned)>& fn) {
    if (T == 0) return;
    if (T == 1 || st_->th.empty()) {
        for (unsigned t = 0; t < T; ++t) fn(t);
        return;
    }
    std::unique_lock<std::mutex> l(st_->m);
    st_->fn = &fn;
    st_->T = T;
    st_->next = 0;
    st_->pending = T;
    ++st_->epoch;
    st_->cv_job.notify_all();
    st_->work(l);
    st_->cv_done.wait(l, [&] { return st_->pending == 0; });
    st_->fn = nullptr;
    st_->T = st_->next = 0;
}

namespace {

struct Rec {
    const char* h;
    const char* s;
    uint32_t hn, sn;
};

// one per worker; padded to its own cache lines: the workers push records and add to `chars` for every read, and
// two shards sharing a line made two workers slower than one (parse phase 0.158 s -> 0.208 s on 4 M reads)
struct alignas(128) Shard {
    std::vector<Rec> recs;
    std::deque<std::string> joined;  // multi-line FASTA reads (stable addresses)
    uint64_t chars = 0;
    void clear() {
        recs.clear();
        joined.clear();
        chars = 0;
    }
};

inline size_t line_end(const char* d, size_t pos, size_t end) {
    const void* nl = memchr(d + pos, '\n', end - pos);
    return nl ? (size_t)((const char*)nl - d) : end;
}

// the reference's FASTA loop (aligner.cpp:69-113) over [a, b): a is a record start, b a record start
// or the end of the input
void parse_fasta_range(const char* d, size_t a, size_t b, uint32_t k, Shard& out) {
    size_t pos = a;
    while (pos < b) {
        const size_t h = pos;
        size_t e = line_end(d, pos, b);
        const size_t hn = e - h;
        pos = e < b ? e + 1 : b;
        if (pos >= b) break;  // header without a sequence line: the second getline fails, read stays empty
        const size_t s = pos;
        e = line_end(d, pos, b);
        size_t sn = e - s;
        pos = e < b ? e + 1 : b;
        const char* sp = d + s;
        if (pos < b && d[pos] != '>') {  // continuation lines: read += inter until peek() sees '>' or the end
            out.joined.emplace_back(sp, sn);
            std::string& j = out.joined.back();
            while (pos < b && d[pos] != '>') {
                e = line_end(d, pos, b);
                j.append(d + pos, e - pos);
                pos = e < b ? e + 1 : b;
            }
            sp = j.data();
            sn = j.size();
        }
        if (sn > 2 && sn > k && sn <= 0xFFFFFFFFull && acgtn_only(sp, sn)) {
            out.recs.push_back(Rec{d + h, sp, (uint32_t)hn, (uint32_t)sn});
            out.chars += sn;
        }
    }
}

// whole 4-line FASTQ records in [a, b), every line newline-terminated (aligner.cpp:49-67 away from
// the end of the stream)
void parse_fastq_range_scalar(const char* d, size_t a, size_t b, Shard& out) {
    size_t pos = a;
    while (pos < b) {
        const size_t h = pos;
        size_t e = line_end(d, pos, b);
        const size_t hn = e - h;
        pos = e + 1;
        const size_t s = pos;
        e = line_end(d, pos, b);
        const size_t sn = e - s;
        pos = e + 1;
        pos = line_end(d, pos, b) + 1;
        pos = line_end(d, pos, b) + 1;
        if (sn > 2 && sn <= 0xFFFFFFFFull && acgtn_only(d + s, sn)) {
            out.recs.push_back(Rec{d + h, d + s, (uint32_t)hn, (uint32_t)sn});
            out.chars += sn;
        }
    }
}

size_t count_newlines_scalar(const char* d, size_t a, size_t b) {
    size_t c = 0;
    for (size_t pos = a; pos < b;) {
        const void* nl = memchr(d + pos, '\n', b - pos);
        if (!nl) break;
        ++c;
        pos = (size_t)((const char*)nl - d) + 1;
    }
    return c;
}

#ifdef BGR_HAVE_AVX2_PATH
// newline bitmask of the 32 bytes at p
__attribute__((target("avx2"))) inline uint32_t newline_mask32(const char* p) {
    const __m256i x = _mm256_loadu_si256((const __m256i*)p);
    return (uint32_t)_mm256_movemask_epi8(_mm256_cmpeq_epi8(x, _mm256_set1_epi8('\n')));
}
__attribute__((target("avx2,popcnt"))) size_t count_newlines_avx2(const char* d, size_t a, size_t b) {
    size_t c = 0, pos = a;
    for (; pos + 32 <= b; pos += 32) c += (size_t)__builtin_popcount(newline_mask32(d + pos));
    return c + count_newlines_scalar(d, pos, b);
}
// position right after the j-th newline (1-based) at or after a; the caller guarantees that it exists
__attribute__((target("avx2,popcnt"))) size_t after_newline_avx2(const char* d, size_t a, size_t j) {
    size_t pos = a;
    for (;;) {
        // 32 bytes at a time while the j-th newline cannot be in them; the newline exists, so the bytes up to it
        // are readable -- a block is only loaded whole when even all-newlines would not overshoot
        if (j > 32) {
            j -= (size_t)__builtin_popcount(newline_mask32(d + pos));
            pos += 32;
            continue;
        }
        const char* nl = (const char*)rawmemchr(d + pos, '\n');
        pos = (size_t)(nl - d) + 1;
        if (--j == 0) return pos;
    }
}

// parse_fastq_range with the newlines found 32 bytes at a time: a record is four lines whatever they hold, so
// every fourth newline closes one; no search call per line (four memchr calls per record were 60 of its 80 ns)
__attribute__((target("avx2"))) void parse_fastq_range_avx2(const char* d, size_t a, size_t b, Shard& out) {
    size_t start = a;       // first byte of the current line
    size_t h = a, hn = 0, s = 0, sn = 0;
    unsigned li = 0;        // line of the record the next newline ends: 0 header, 1 sequence, 2 '+', 3 quality
    uint64_t chars = 0;
    auto newline_at = [&](size_t nl) {
        if (li == 0) {
            h = start;
            hn = nl - start;
        } else if (li == 1) {
            s = start;
            sn = nl - start;
        } else if (li == 3) {
            if (sn > 2 && sn <= 0xFFFFFFFFull && acgtn_only(d + s, sn)) {
                out.recs.push_back(Rec{d + h, d + s, (uint32_t)hn, (uint32_t)sn});
                chars += sn;
            }
        }
        li = (li + 1) & 3;
        start = nl + 1;
    };
    size_t pos = a;
    for (; pos + 32 <= b; pos += 32) {
        uint32_t m = newline_mask32(d + pos);
        while (m) {
            newline_at(pos + (size_t)__builtin_ctz(m));
            m &= m - 1;
        }
    }
    for (; pos < b; ++pos)
        if (d[pos] == '\n') newline_at(pos);
    out.chars += chars;
}
#endif

void parse_fastq_range(const char* d, size_t a, size_t b, Shard& out) {
#ifdef BGR_HAVE_AVX2_PATH
    static const bool avx2 = __builtin_cpu_supports("avx2");
    if (avx2) return parse_fastq_range_avx2(d, a, b, out);
#endif
    parse_fastq_range_scalar(d, a, b, out);
}

// newlines in [a, b): the FASTQ reader and the unitig loader cut their input by line numbers, so this pass runs
// over every byte of such a file; one 32-byte compare + popcount per step instead of a memchr call per line
size_t count_newlines(const char* d, size_t a, size_t b) {
#ifdef BGR_HAVE_AVX2_PATH
    static const bool avx2 = __builtin_cpu_supports("avx2") && __builtin_cpu_supports("popcnt");
    if (avx2) return count_newlines_avx2(d, a, b);
#endif
    return count_newlines_scalar(d, a, b);
}

// position right after the j-th newline (1-based) at or after a; j newlines must exist
size_t after_newline(const char* d, size_t a, size_t j) {
#ifdef BGR_HAVE_AVX2_PATH
    static const bool avx2 = __builtin_cpu_supports("avx2") && __builtin_cpu_supports("popcnt");
    if (avx2) return after_newline_avx2(d, a, j);
#endif
    size_t pos = a;
    while (j--) pos = (size_t)((const char*)rawmemchr(d + pos, '\n') - d) + 1;
    return pos;
}

// Line numbers -> positions for a text [a, b): the newlines of T slices are counted in parallel, per 64 KB block,
// so that "right after the j-th newline" is found by skipping whole blocks and scanning one (a second pass over
// the text otherwise, and the last cut was made by one thread).
struct LineIndex {
    static const size_t BLK = (size_t)1 << 16;
    std::vector<size_t> lo, cnt;  // T+1 slice bounds; cnt[t] = newlines before slice t
    std::vector<std::vector<uint32_t>> blk;
    void build(const char* d, size_t a, size_t b, unsigned T, WorkerPool& pool) {
        lo.assign(T + 1, 0);
        cnt.assign(T + 1, 0);
        blk.assign(T, std::vector<uint32_t>());
        for (unsigned t = 0; t <= T; ++t) lo[t] = a + (size_t)((b - a) / T) * t;
        lo[T] = b;
        pool.run(T, [&](unsigned t) {
            std::vector<uint32_t> mine;  // local: the slices' vectors are neighbours in memory
            mine.reserve((lo[t + 1] - lo[t]) / BLK + 1);
            size_t c = 0;
            for (size_t pos = lo[t]; pos < lo[t + 1]; pos += BLK) {
                const size_t n = count_newlines(d, pos, std::min(pos + BLK, lo[t + 1]));
                mine.push_back((uint32_t)n);
                c += n;
            }
            blk[t].swap(mine);
            cnt[t + 1] = c;
        });
        for (unsigned t = 0; t < T; ++t) cnt[t + 1] += cnt[t];
    }
    size_t total() const { return cnt.back(); }
    // position right after newline number j (1-based, j <= total())
    size_t after(const char* d, size_t j) const {
        const unsigned t = (unsigned)(std::upper_bound(cnt.begin(), cnt.end(), j - 1) - cnt.begin()) - 1;  // cnt[t] < j <= cnt[t+1]
        size_t left = j - cnt[t], pos = lo[t];
        for (uint32_t n : blk[t]) {
            if (left <= n) break;  // the newline is in this block
            left -= n;
            pos += BLK;
        }
        return after_newline(d, pos, left);
    }
};

// append the shards' records to the batch, in shard order, copying in parallel
void gather(std::vector<Shard>& sh, size_t T, WorkerPool& pool, ParsedBatch& out) {
    std::vector<size_t> r0(T + 1, 0), c0(T + 1, 0);
    for (size_t t = 0; t < T; ++t) {
        r0[t + 1] = r0[t] + sh[t].recs.size();
        c0[t + 1] = c0[t] + sh[t].chars;
    }
    const size_t base_r = out.hdr.n, base_c = out.seqs.n;
    out.seqs.reserve(base_c + c0[T] + 64);
    out.soff.reserve(base_r + r0[T] + 2);
    out.hdr.reserve(base_r + r0[T] + 1);
    pool.run((unsigned)T, [&](unsigned t) {
        size_t r = base_r + r0[t], c = base_c + c0[t];
        for (const Rec& x : sh[t].recs) {
            memcpy(out.seqs.p + c, x.s, x.sn);
            c += x.sn;
            out.soff.p[r + 1] = c;
            out.hdr.p[r] = HdrView{x.h, x.hn};
            ++r;
        }
    });
    out.seqs.n = base_c + c0[T];
    out.soff.n = base_r + r0[T] + 1;
    out.hdr.n = base_r + r0[T];
}

}  // namespace

struct ParallelReader::Impl {
    WorkerPool pool;
    std::vector<Shard> sh;
    explicit Impl(unsigned threads) : pool(threads - 1), sh(threads) {}
    std::vector<Shard>& shards(unsigned T) {
        for (unsigned t = 0; t < T; ++t) sh[t].clear();
        return sh;
    }
};

ParallelReader::ParallelReader(const char* data, size_t n, uint32_t k, bool fastq, unsigned threads,
                               size_t range_bytes, size_t min_slice)
    : d_(data),
      n_(n),
      range_(range_bytes ? range_bytes : 1),
      min_slice_(min_slice ? min_slice : 1),
      k_(k),
      fastq_(fastq),
      threads_(threads ? threads : 1),
      impl_(new Impl(threads_)) {}

ParallelReader::~ParallelReader() { delete impl_; }

// smallest unambiguous record start >= x (see the header), or n_
size_t ParallelReader::fasta_boundary(size_t x) const {
    if (x == 0) return 0;
    size_t p = x - 1;
    for (;;) {
        if (p >= n_) return n_;
        const void* nl = memchr(d_ + p, '\n', n_ - p);
        if (!nl) return n_;
        const size_t c = (size_t)((const char*)nl - d_) + 1;  // start of the line after that newline
        if (c >= n_) return n_;
        if (d_[c] == '>') {
            // previous line = [ps, c-1)
            const void* q = c >= 2 ? memrchr(d_, '\n', c - 1) : nullptr;
            const size_t ps = q ? (size_t)((const char*)q - d_) + 1 : 0;
            if (ps != 0 && d_[ps] != '>') return c;
        }
        p = c;
    }
}

void ParallelReader::next(ParsedBatch& out) {
    if (finished_) return;
    if (fastq_) next_fastq(out);
    else next_fasta(out);
}

bool ParallelReader::fasta_next_range(size_t& a, size_t& b) {
    if (finished_) return false;
    a = pos_;
    const size_t want = a + range_;
    b = (want >= n_ || want < a) ? n_ : fasta_boundary(want);
    pos_ = b;
    if (pos_ >= n_) finished_ = true;
    return true;
}

void ParallelReader::fasta_parse_range(size_t a, size_t b, ParsedBatch& out) {
    const unsigned T = (unsigned)std::max<size_t>(1, std::min<size_t>(threads_, (b - a) / min_slice_ + 1));
    std::vector<size_t> cut(T + 1);
    cut[0] = a;
    cut[T] = b;
    WorkerPool& pool = impl_->pool;
    const auto t0 = std::chrono::steady_clock::now();
    pool.run(T, [&](unsigned t) {
        if (t) cut[t] = std::min(b, fasta_boundary(a + (size_t)((b - a) / T) * t));
    });
    const auto t1 = std::chrono::steady_clock::now();
    std::vector<Shard>& sh = impl_->shards(T);
    pool.run(T, [&](unsigned t) {
        if (cut[t] < cut[t + 1]) parse_fasta_range(d_, cut[t], cut[t + 1], k_, sh[t]);
    });
    const auto t2 = std::chrono::steady_clock::now();
    gather(sh, T, pool, out);
    const auto t3 = std::chrono::steady_clock::now();
    phase_s[0] += std::chrono::duration<double>(t1 - t0).count();
    phase_s[1] += std::chrono::duration<double>(t2 - t1).count();
    phase_s[2] += std::chrono::duration<double>(t3 - t2).count();
}

void ParallelReader::next_fasta(ParsedBatch& out) {
    size_t a, b;
    if (fasta_next_range(a, b)) fasta_parse_range(a, b, out);
}

void parse_fasta_text_sequential(const char* d, size_t a, size_t b, uint32_t k, ParsedBatch& out) {
    std::vector<Shard> sh(1);
    if (a < b) parse_fasta_range(d, a, b, k, sh[0]);
    WorkerPool pool(0);
    gather(sh, 1, pool, out);
}

// The next window of a FASTQ input: [a, b) = whole newline-terminated records (a multiple of getReads' chunk of
// 10000, so that what follows starts where a getReads call would), cut for T workers; `final`: this window reaches
// the end of the input, whose last <= 2 chunks are left to the sequential parser (fastq_tail).
bool ParallelReader::fastq_next_range(size_t& a, size_t& b, bool& final, std::vector<size_t>* cuts) {
    if (finished_ || tail_pending_) return false;
    a = pos_;
    const size_t CH = 10000;  // reads per getReads call (alignerGreedy.cpp:376)
    size_t win = range_;
    for (;;) {
        final = a + win >= n_ || a + win < a;
        const size_t e = final ? n_ : a + win;
        const unsigned T = (unsigned)std::max<size_t>(1, std::min<size_t>(threads_, (e - a) / min_slice_ + 1));
        // newline counts per slice
        WorkerPool& pool = impl_->pool;
        const auto t0 = std::chrono::steady_clock::now();
        LineIndex li;
        li.build(d_, a, e, T, pool);
        const size_t q = li.total() / 4;  // whole newline-terminated records in [a, e)
        size_t P;                     // records parsed by the workers (or the device)
        if (final) P = q ? ((q - 1) / CH) * CH : 0;  // the last <= 2*CH records go to the sequential parser
        else P = (q / CH) * CH;
        if (!final && P == 0) {  // not even one getReads chunk in the window: look further
            win *= 2;
            continue;
        }
        // start of record r = right after newline number 4r
        auto record_start = [&](size_t r) -> size_t {
            if (r == 0) return a;
            return li.after(d_, 4 * r);  // 1-based newline number
        };
        if (cuts) {
            cuts->assign(T + 1, 0);
            pool.run(T, [&](unsigned t) { (*cuts)[t] = record_start(P / T * t); });
            (*cuts)[T] = record_start(P);
            b = (*cuts)[T];
        } else {
            b = record_start(P);
        }
        phase_s[0] += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        pos_ = b;
        if (final) tail_pending_ = true;
        return true;
    }
}

// the end of the stream, exactly as the reference sees it (phantom duplicate of the last read, truncated record)
void ParallelReader::fastq_tail(ParsedBatch& out) {
    if (!tail_pending_) return;
    const size_t CH = 10000;
    ReadParser rp(d_ + pos_, n_ - pos_, k_, true);
    ReadBatch tail;
    while (!rp.done()) rp.next_chunk((uint32_t)CH, tail);
    const size_t nt = tail.size();
    out.arena.assign(tail.headers.p, tail.headers.p + tail.headers.n);
    out.seqs.reserve(out.seqs.n + tail.seqs.n + 64);
    out.soff.reserve(out.soff.n + nt + 1);
    for (size_t i = 0; i < nt; ++i) {
        const size_t sn = tail.soff.p[i + 1] - tail.soff.p[i];
        memcpy(out.seqs.p + out.seqs.n, tail.seqs.p + tail.soff.p[i], sn);
        out.seqs.n += sn;
        out.soff.p[out.soff.n++] = out.seqs.n;
        out.hdr.push(HdrView{out.arena.data() + tail.hoff.p[i], (uint32_t)(tail.hoff.p[i + 1] - tail.hoff.p[i])});
    }
    pos_ = n_;
    tail_pending_ = false;
    finished_ = true;
}

void ParallelReader::fastq_parse_range(size_t a, size_t b, ParsedBatch& out, const std::vector<size_t>* cuts) {
    WorkerPool& pool = impl_->pool;
    std::vector<size_t> one{a, b};
    const std::vector<size_t>& cut = cuts ? *cuts : one;
    const unsigned T = (unsigned)cut.size() - 1;
    const auto t1 = std::chrono::steady_clock::now();
    std::vector<Shard>& sh = impl_->shards(T);
    pool.run(T, [&](unsigned t) {
        if (cut[t] < cut[t + 1]) parse_fastq_range(d_, cut[t], cut[t + 1], sh[t]);
    });
    const auto t2 = std::chrono::steady_clock::now();
    gather(sh, T, pool, out);
    const auto t3 = std::chrono::steady_clock::now();
    phase_s[1] += std::chrono::duration<double>(t2 - t1).count();
    phase_s[2] += std::chrono::duration<double>(t3 - t2).count();
}

void ParallelReader::next_fastq(ParsedBatch& out) {
    size_t a, b;
    bool final = false;
    std::vector<size_t> cuts;
    if (fastq_next_range(a, b, final, &cuts)) fastq_parse_range(a, b, out, &cuts);
    if (final) fastq_tail(out);
}

void parse_fastq_text_sequential(const char* d, size_t a, size_t b, ParsedBatch& out) {
    std::vector<Shard> sh(1);
    if (a < b) parse_fastq_range(d, a, b, sh[0]);
    WorkerPool pool(0);
    gather(sh, 1, pool, out);
}

// ------------------------------------------------------------------------------ parallel unitig loader

bool load_unitig_file_parallel(const char* path, uint32_t k, unsigned threads, PodBuf<char>& seqs,
                               PodBuf<uint64_t>& off, size_t min_slice) {
    seqs.n = off.n = 0;
    off.push(0);
    MappedFile f;
    if (!f.open(path)) return false;
    const char* d = f.data;
    const size_t n = f.size;
    const unsigned T = (unsigned)std::max<size_t>(1, std::min<size_t>(threads ? threads : 1, n / (min_slice ? min_slice : 1) + 1));
    WorkerPool pool(T - 1);
    // newline counts per slice -> start of record r (= right after newline number 2r)
    LineIndex li;
    li.build(d, 0, n, T, pool);
    const size_t q = li.total() / 2;    // whole newline-terminated records
    const size_t P = q ? q - 1 : 0;     // records the workers take; the rest is the sequential tail
    auto record_start = [&](size_t r) -> size_t {
        if (r == 0) return 0;
        return li.after(d, 2 * r);
    };
    std::vector<size_t> cut(T + 1), first_rec(T + 1);
    for (unsigned t = 0; t <= T; ++t) first_rec[t] = P / T * t;
    first_rec[T] = P;
    pool.run(T, [&](unsigned t) { cut[t] = record_start(first_rec[t]); });
    cut[T] = record_start(P);
    // pass 1: sequence line of every record (view), up to the worker's first line shorter than k
    struct alignas(128) Part {  // own cache lines: see Shard
        std::vector<std::pair<const char*, size_t>> recs;
        uint64_t chars = 0;
        bool stopped = false;
    };
    std::vector<Part> part(T);
    pool.run(T, [&](unsigned t) {
        size_t pos = cut[t];
        const size_t end = cut[t + 1];
        Part& p = part[t];
        p.recs.reserve(first_rec[t + 1] - first_rec[t]);
        while (pos < end) {
            pos = line_end(d, pos, end) + 1;  // header line
            const size_t e = line_end(d, pos, end);
            const size_t len = e - pos;
            if (len < k) {  // aligner.cpp:418
                p.stopped = true;
                break;
            }
            p.recs.emplace_back(d + pos, len);
            p.chars += len;
            pos = e + 1;
        }
    });
    // loading stops at the first short line: later workers' records are dropped
    unsigned used = T;
    bool stopped = false;
    for (unsigned t = 0; t < T; ++t)
        if (part[t].stopped) {
            used = t + 1;
            stopped = true;
            break;
        }
    std::vector<uint64_t> r0(used + 1, 0), c0(used + 1, 0);
    for (unsigned t = 0; t < used; ++t) {
        r0[t + 1] = r0[t] + part[t].recs.size();
        c0[t + 1] = c0[t] + part[t].chars;
    }
    seqs.reserve(c0[used] + 64);
    off.reserve(r0[used] + 4);
    pool.run(used, [&](unsigned t) {
        uint64_t r = r0[t], c = c0[t];
        for (const auto& x : part[t].recs) {
            memcpy(seqs.p + c, x.first, x.second);
            c += x.second;
            off.p[++r] = c;
        }
    });
    seqs.n = c0[used];
    off.n = r0[used] + 1;
    if (!stopped) {
        // the end of the file, exactly as the reference's loop sees it (aligner.cpp:415-421)
        LineStream s(d + cut[T], n - cut[T], true);
        std::string line;
        while (!s.eof()) {
            s.getline(line);
            s.getline(line);
            if (line.size() < k) break;
            seqs.append(line.data(), line.size());
            off.push(seqs.n);
        }
    }
    return true;
}

}  // namespace bgr
