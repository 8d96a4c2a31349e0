// Host-side 2-bit packing of a read batch (the "host side 2-bit packs reads" step of the design;
// the reference keeps reads as std::string and encodes (k-1)-mers on the fly with str2num / nuc2int,
// utils.cpp:117-151). The packed stream is what travels over PCIe: 2 bits per base instead of 8.
//
// Layout: the batch's characters are one contiguous stream (read i = chars [off[i], off[i+1])), and
// the packed stream is the same stream at 2 bits per character: word w holds characters
// [32w, 32w+32), first character in the most significant bits (str2num digit order), A0 C1 G2 T3.
// Packing is therefore independent of the read boundaries and splits over threads at word boundaries.
// A character outside ACGT has no 2-bit code: the reads that contain one are flagged in a bitmap and
// travel as characters beside the stream (bgr_slot_upload_packed), because the reference treats such a
// character three different ways depending on where it is looked at.
#include <immintrin.h>
#include <string.h>

#include <algorithm>
#include <mutex>
#include <vector>

#include "bgr_host.h"

namespace bgr {
namespace {

inline uint64_t pack32_scalar(const unsigned char* p, size_t n, bool& bad) {
    uint64_t w = 0;
    for (size_t j = 0; j < n; ++j) {
        const uint32_t c = base_code(p[j]);
        if (c > 3) bad = true;
        w |= (uint64_t)(c & 3) << (62 - 2 * j);
    }
    return w;
}

// characters [32*w0, 32*w1) of s (all inside the stream) -> words; returns true if any is not A/C/G/T
bool pack_words_scalar(const unsigned char* s, uint64_t w0, uint64_t w1, uint64_t* words) {
    bool bad = false;
    for (uint64_t w = w0; w < w1; ++w) words[w] = pack32_scalar(s + 32 * w, 32, bad);
    return bad;
}

#if defined(__x86_64__) && defined(__GNUC__)
#define BGR_PACK_AVX2 1
__attribute__((target("avx2"))) bool pack_words_avx2(const unsigned char* s, uint64_t w0, uint64_t w1, uint64_t* words) {
    const __m256i m3 = _mm256_set1_epi8(3);
    // the character each code stands for, to validate: code = ((c>>1)^(c>>2))&3 maps A C G T to 0 1 2 3
    const __m256i lut = _mm256_setr_epi8('A', 'C', 'G', 'T', 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 'A', 'C', 'G', 'T', 0, 0, 0,
                                         0, 0, 0, 0, 0, 0, 0, 0, 0);
    const __m256i mul1 = _mm256_set1_epi16(0x0104);      // bytes (4, 1): c0*4 + c1
    const __m256i mul2 = _mm256_set1_epi32(0x00010010);  // words (16, 1): (c0c1)*16 + (c2c3)
    // low byte of every 32-bit lane, last lane first (first character most significant)
    const __m256i gather = _mm256_setr_epi8(12, 8, 4, 0, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, 12, 8, 4, 0, -1, -1,
                                            -1, -1, -1, -1, -1, -1, -1, -1, -1, -1);
    __m256i badacc = _mm256_setzero_si256();
#define BGR_PACK_ONE(W)                                                                                             \
    do {                                                                                                            \
        const __m256i x = _mm256_loadu_si256((const __m256i*)(s + 32 * (W)));                                       \
        const __m256i t = _mm256_and_si256(_mm256_xor_si256(_mm256_srli_epi16(x, 1), _mm256_srli_epi16(x, 2)), m3); \
        badacc = _mm256_or_si256(badacc, _mm256_xor_si256(_mm256_shuffle_epi8(lut, t), x));                         \
        const __m256i b = _mm256_madd_epi16(_mm256_maddubs_epi16(t, mul1), mul2); /* a byte value per 4 characters */ \
        const __m256i g = _mm256_shuffle_epi8(b, gather);                                                           \
        const uint32_t hi = (uint32_t)_mm256_extract_epi32(g, 0); /* characters 0..15 */                            \
        const uint32_t lo = (uint32_t)_mm256_extract_epi32(g, 4); /* characters 16..31 */                           \
        words[(W)] = ((uint64_t)hi << 32) | lo;                                                                     \
    } while (0)
    uint64_t w = w0;
    for (; w + 4 <= w1; w += 4) {
        _mm_prefetch((const char*)(s + 32 * w + 1024), _MM_HINT_T0);
        BGR_PACK_ONE(w);
        BGR_PACK_ONE(w + 1);
        BGR_PACK_ONE(w + 2);
        BGR_PACK_ONE(w + 3);
    }
    for (; w < w1; ++w) BGR_PACK_ONE(w);
#undef BGR_PACK_ONE
    return !_mm256_testz_si256(badacc, badacc);
}
#endif

bool pack_words(const unsigned char* s, uint64_t w0, uint64_t w1, uint64_t* words) {
#ifdef BGR_PACK_AVX2
    static const bool avx2 = __builtin_cpu_supports("avx2");
    if (avx2) return pack_words_avx2(s, w0, w1, words);
#endif
    return pack_words_scalar(s, w0, w1, words);
}

std::mutex g_pool_mutex;
WorkerPool* g_pool = nullptr;
unsigned g_pool_workers = 0;

}  // namespace

uint32_t pack_reads_2bit(const char* seqs, const uint64_t* off, uint32_t n, unsigned threads, uint64_t* words,
                         uint32_t* bad_bits) {
    const unsigned char* s = (const unsigned char*)seqs;
    const uint64_t chars = off[n];
    const uint64_t full = chars >> 5;  // whole words
    memset(bad_bits, 0, (((size_t)n + 31) / 32) * sizeof(uint32_t));
    // blocks of 128 words (4096 characters): a block that reports a bad character is looked at again
    const uint64_t BLOCK = 128;
    const uint64_t nblocks = (full + BLOCK - 1) / BLOCK;
    if (threads < 1) threads = 1;
    unsigned T = (unsigned)std::min<uint64_t>(threads, nblocks / 8 + 1);
    std::vector<std::vector<uint64_t>> bad_blocks(T);
    auto work = [&](unsigned t) {
        const uint64_t b0 = nblocks * t / T, b1 = nblocks * (t + 1) / T;
        for (uint64_t b = b0; b < b1; ++b) {
            const uint64_t w0 = b * BLOCK, w1 = std::min(full, w0 + BLOCK);
            if (pack_words(s, w0, w1, words)) bad_blocks[t].push_back(b);
        }
    };
    if (T == 1) {
        work(0);
    } else {
        std::lock_guard<std::mutex> l(g_pool_mutex);
        if (!g_pool || g_pool_workers < T - 1) {
            delete g_pool;
            g_pool = new WorkerPool(T - 1);
            g_pool_workers = T - 1;
        }
        g_pool->run(T, work);
    }
    // the tail word (zero padded)
    std::vector<uint64_t> bad_pos;
    if (chars & 31) {
        bool bad = false;
        words[full] = pack32_scalar(s + 32 * full, (size_t)(chars & 31), bad);
        if (bad)
            for (uint64_t p = 32 * full; p < chars; ++p)
                if (base_code(s[p]) > 3) bad_pos.push_back(p);
    }
    for (unsigned t = 0; t < T; ++t)
        for (uint64_t b : bad_blocks[t]) {
            const uint64_t p0 = b * BLOCK * 32, p1 = std::min(full * 32, p0 + BLOCK * 32);
            for (uint64_t p = p0; p < p1; ++p)
                if (base_code(s[p]) > 3) bad_pos.push_back(p);
        }
    // character position -> read (empty reads own no character: upper_bound skips them)
    uint32_t n_bad = 0;
    for (uint64_t p : bad_pos) {
        const uint64_t* it = std::upper_bound(off, off + n + 1, p);
        const uint32_t r = (uint32_t)(it - off) - 1;
        if (!(bad_bits[r >> 5] >> (r & 31) & 1u)) {
            bad_bits[r >> 5] |= 1u << (r & 31);
            ++n_bad;
        }
    }
    return n_bad;
}

}  // namespace bgr
