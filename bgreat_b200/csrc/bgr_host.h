// Host-side pieces of the mapping path (C++): unitig loading, 2-bit packing, border-index
// construction, read parsing. None of this touches CUDA; bgr_runtime.cu uploads the blob.
#pragma once
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <functional>
#include <string>
#include <vector>

#include "bgr_common.h"

namespace bgr {

// std::istream look-alike over a memory buffer with exactly the state-bit behaviour the
// reference's parsers depend on (std::getline / peek / eof on an ifstream; reference
// aligner.cpp:46-117 and :415-421): see DESIGN.md "reader semantics".
class LineStream {
   public:
    LineStream(const char* data, size_t n, bool open_ok = true) : d_(data), n_(n), fail_(!open_ok) {}
    bool good() const { return !eof_ && !fail_; }
    bool eof() const { return eof_; }
    // getline(stream, s): s is left untouched when the stream is not good() on entry.
    void getline(std::string& s);
    // same, but yields a view into the buffer instead of copying (valid = line was extracted)
    bool getline_view(const char*& p, size_t& len);
    int peek();

   private:
    const char* d_;
    size_t n_, pos_ = 0;
    bool eof_ = false, fail_ = false;
};

struct MappedFile {
    const char* data = nullptr;
    size_t size = 0;
    bool ok = false;
    int fd = -1;
    void* slurped_ = nullptr;  // owns the bytes when the path is not a regular file
    bool open(const char* path);
    void close();
    ~MappedFile() { close(); }
};

// Unitig loader: the reference's record loop (two getlines per record, keep the second, stop at
// the first kept line shorter than k; aligner.cpp:415-421). Returns false if the file cannot be
// opened (the reference then runs with an empty index).
bool load_unitig_file(const char* path, uint32_t k, std::string& seqs, std::vector<uint64_t>& off);

// Build the flat index blob (layout in bgr_common.h). Stops at the first sequence shorter than k.
// Returns 0, or a BGR_ERR_* code with `err` set.
int build_index_blob(const char* seqs, const uint64_t* off, uint32_t n, uint32_t k, std::vector<uint64_t>& blob,
                     std::string& err);

// Growable POD buffer over a caller-chosen allocator (the CLI plugs in pinned host memory so that
// the batch can be copied to the GPU asynchronously).
typedef void* (*AllocFn)(size_t);
typedef void (*FreeFn)(void*);
template <typename T>
struct PodBuf {
    T* p = nullptr;
    size_t n = 0, cap = 0;
    AllocFn alloc = nullptr;
    FreeFn dealloc = nullptr;
    PodBuf() {}
    PodBuf(const PodBuf&) = delete;
    PodBuf& operator=(const PodBuf&) = delete;
    ~PodBuf() { release(); }
    void release() {
        if (p) (dealloc ? dealloc : ::free)(p);
        p = nullptr;
        n = cap = 0;
    }
    void reserve(size_t need) {
        if (need <= cap) return;
        size_t c = cap ? cap : 1024;
        while (c < need) c += c / 2 + 1;
        T* q = (T*)(alloc ? alloc(c * sizeof(T)) : ::malloc(c * sizeof(T)));
        if (p) {
            memcpy(q, p, n * sizeof(T));
            (dealloc ? dealloc : ::free)(p);
        }
        p = q;
        cap = c;
    }
    void append(const T* s, size_t m) {
        reserve(n + m + 16);
        memcpy(p + n, s, m * sizeof(T));
        n += m;
    }
    void push(T x) {
        reserve(n + 1);
        p[n++] = x;
    }
};

// The reads one or more getReads calls delivered: headers and sequences, concatenated.
struct ReadBatch {
    PodBuf<char> headers, seqs;
    PodBuf<uint64_t> hoff, soff;
    ReadBatch() { clear(); }
    void set_allocator(AllocFn a, FreeFn f) {
        seqs.alloc = soff.alloc = a;
        seqs.dealloc = soff.dealloc = f;
    }
    size_t size() const { return soff.n - 1; }
    void clear() {
        headers.n = seqs.n = hoff.n = soff.n = 0;
        hoff.push(0);
        soff.push(0);
    }
    void push(const char* h, size_t hn, const char* s, size_t sn) {
        headers.append(h, hn);
        seqs.append(s, sn);
        hoff.push(headers.n);
        soff.push(seqs.n);
    }
};

// Reader with the reference's getReads semantics (aligner.cpp:46-117) driven like
// alignPartGreedy drives it (alignerGreedy.cpp:372-377): next_chunk(n) appends what one
// getReads(reads, n) call would return; done() is readFile.eof().
class ReadParser {
   public:
    ReadParser(const char* data, size_t n, uint32_t k, bool fastq) : s_(data, n), k_(k), fastq_(fastq) {}
    bool done() const { return s_.eof(); }
    void next_chunk(uint32_t n, ReadBatch& out);

   private:
    LineStream s_;
    uint32_t k_;
    bool fastq_;
};

// ------------------------------------------------------------------------------ parallel reader
//
// Same records as ReadParser (= the reference's getReads driven in 10000-read chunks), produced by
// `threads` workers over ranges of the input text. SURVEY 8(f) rank 1: the parser must keep up with
// the GPU, which a single getline loop cannot.
//
//   FASTA  the reference's state machine (header line, first sequence line, continuation lines until
//          a line that starts with '>') does not depend on the chunking of getReads, so the text is cut
//          at record starts that are unambiguous -- a line starting with '>' whose previous line does not
//          (and is not the first line of the file: every other header is reached through peek()=='>').
//   FASTQ  records are 4 lines whatever their content: newlines are counted in parallel and the text is
//          cut at line numbers that are multiples of 4 x 10000 (getReads boundaries). What depends on
//          the stream state at the end of the file (partial last record, phantom duplicate of the last
//          read after a trailing newline, aligner.cpp:52-53) is left to the sequential ReadParser,
//          which runs over the last <= 20000 records only.
struct HdrView {
    const char* p;
    uint32_t n;
};

struct ParsedBatch {
    PodBuf<char> seqs;      // sequences, concatenated (allocator: pinned host memory in the CLI)
    PodBuf<uint64_t> soff;  // n+1 offsets into seqs
    PodBuf<HdrView> hdr;    // header lines: views into the input text, or into `arena`
    std::vector<char> arena;   // headers that do not live in the input text (sequential tail)
    ParsedBatch() { clear(); }
    void set_allocator(AllocFn a, FreeFn f) {
        seqs.alloc = soff.alloc = a;
        seqs.dealloc = soff.dealloc = f;
    }
    size_t size() const { return hdr.n; }
    void clear() {
        seqs.n = soff.n = hdr.n = 0;
        soff.push(0);
        arena.clear();
    }
};

// Persistent workers for the data-parallel phases of the reader and the writer: run(T, fn) calls
// fn(t) for t in [0, T) on up to size()+1 threads (the caller is one of them) and returns when all are
// done. One pool serves one caller at a time.
class WorkerPool {
   public:
    explicit WorkerPool(unsigned workers);
    ~WorkerPool();
    WorkerPool(const WorkerPool&) = delete;
    WorkerPool& operator=(const WorkerPool&) = delete;
    void run(unsigned T, const std::function<void(unsigned)>& fn);

   private:
    struct State;
    State* st_;
};

// every character one of A C G T N (aligner.cpp:55-60)
bool acgtn_only(const char* p, size_t n);

// 2-bit packing of a batch on `threads` host threads (bgr_pack.cpp): words[w] = characters [32w, 32w+32) of
// the batch's character stream, first character most significant, A0 C1 G2 T3; (off[n]+31)/32 words are
// written. bad_bits: bitmap over the reads ((n+31)/32 words), bit set = the read holds a character outside
// ACGT (its 2-bit codes are then meaningless). Returns the number of such reads.
uint32_t pack_reads_2bit(const char* seqs, const uint64_t* off, uint32_t n, unsigned threads, uint64_t* words,
                         uint32_t* bad_bits);

class ParallelReader {
   public:
    // range_bytes: text handed out per next() call (a batch); FASTQ ranges are rounded to whole
    // getReads chunks and grow when a chunk does not fit
    // min_slice: least text per worker (tests set it to 1 to cut tiny inputs everywhere)
    ParallelReader(const char* data, size_t n, uint32_t k, bool fastq, unsigned threads, size_t range_bytes,
                   size_t min_slice = 4096);
    ~ParallelReader();
    ParallelReader(const ParallelReader&) = delete;
    ParallelReader& operator=(const ParallelReader&) = delete;
    bool done() const { return finished_; }
    // appends the records of the next range to `out` (which the caller has cleared)
    void next(ParsedBatch& out);
    // FASTA, in two steps, for a caller that may hand the range's text to the device instead of parsing it here
    // (bgr_slot_upload_text): the next range [a, b) of whole records (false: the input is used up), and its parse
    bool fasta_next_range(size_t& a, size_t& b);
    void fasta_parse_range(size_t a, size_t b, ParsedBatch& out);
    // FASTQ likewise: [a, b) = whole newline-terminated records; `final`: the input's last records (at most two
    // getReads chunks) are still to come and are parsed by fastq_tail, sequentially, with the reference's
    // end-of-stream behaviour. cuts: the range cut for the workers of fastq_parse_range (optional).
    bool fastq_next_range(size_t& a, size_t& b, bool& final, std::vector<size_t>* cuts = nullptr);
    void fastq_parse_range(size_t a, size_t b, ParsedBatch& out, const std::vector<size_t>* cuts = nullptr);
    bool fastq_tail_pending() const { return tail_pending_; }
    void fastq_tail(ParsedBatch& out);
    // wall time spent so far in the phases of next(): cutting (FASTQ: counting lines), parsing, copying (diagnostics)
    double phase_s[3] = {0, 0, 0};

   private:
    void next_fasta(ParsedBatch& out);
    void next_fastq(ParsedBatch& out);
    size_t fasta_boundary(size_t x) const;
    struct Impl;  // worker pool + per-worker record lists, reused from range to range

    const char* d_;
    size_t n_, pos_ = 0, range_, min_slice_;
    uint32_t k_;
    bool fastq_, finished_ = false, tail_pending_ = false;
    unsigned threads_;
    Impl* impl_;
};

// the FASTA records of d[a, b) (a a record start, b a record start or the end of the input) on the calling thread
void parse_fasta_text_sequential(const char* d, size_t a, size_t b, uint32_t k, ParsedBatch& out);
// the same for whole 4-line FASTQ records
void parse_fastq_text_sequential(const char* d, size_t a, size_t b, ParsedBatch& out);

// load_unitig_file on `threads` workers: records are two lines whatever they contain, so newlines are
// counted in parallel and every worker takes whole records; the last record goes through the sequential
// loop, which knows what a missing final line does (a failed second getline keeps the header text).
// seqs/off as bgr_index_build wants them. Returns false if the file cannot be opened.
bool load_unitig_file_parallel(const char* path, uint32_t k, unsigned threads, PodBuf<char>& seqs,
                               PodBuf<uint64_t>& off, size_t min_slice = 1 << 16);

}  // namespace bgr
