// Host-side pieces of the mapping path (C++): unitig loading, 2-bit packing, border-index
// construction, read parsing. None of this touches CUDA; bgr_runtime.cu uploads the blob.
#pragma once
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

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

}  // namespace bgr
