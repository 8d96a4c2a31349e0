// Record parsing and output formatting on the device: the two host stages either side of the mapping path
//  for input text of the regular shape -- every line newline-terminated, a record = a header
// line and a sequence line (FASTA; the header starts with '>' and the sequence does not) or four lines (FASTQ).
//   getReads (aligner.cpp:46-117): a read is kept when its sequence is longer than 2 (and, FASTA, than k) and
//     consists of A C G T N; anything else is dropped silently and not counted
//   alignPartGreedy's output (alignerGreedy.cpp:393-428) + printPath (aligner.cpp:600-609): header\npath\n with
//     every int followed by '.', or header\nread\n for a read that does not map (header\ncorrected\n with -c)
// Text that is not of that shape (continuation lines, a missing final newline, a header without '>') is reported
// back and parsed by the host's ParallelReader / ReadParser, which follow the reference's stream semantics to the
// letter. Part of bgr_kernels.cuh; kernels only.
#pragma once
#include "bgr_dev.cuh"

namespace bgr {

#if defined(__CUDACC__)

struct TextMeta {
    unsigned long long n_lines, n_reads, chars;
    u32 irregular;  // some record is not of the regular shape
    u32 pad;
};

static const int TX_THREADS = 256;

// newlines in bytes [16 t, 16 t + 16)
__global__ void __launch_bounds__(TX_THREADS) k_text_count(const char* text, u64 bytes, u32* cnt) {
    const u64 t = (u64)blockIdx.x * TX_THREADS + threadIdx.x;
    const u64 lo = 16 * t;
    if (lo >= bytes) return;
    u32 c = 0;
    if (lo + 16 <= bytes) {
        const uint4 v = *(const uint4*)(text + lo);  // the buffer is 16-byte aligned
        const u32 w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const u32 x = w[j] ^ 0x0A0A0A0Au;  // zero bytes where the text has '\n'
            c += __popc(((x - 0x01010101u) & ~x & 0x80808080u));
        }
    } else {
        for (u64 i = lo; i < bytes; ++i) c += text[i] == '\n';
    }
    cnt[t] = c;
}

// position of every newline, in order: line j ends at nl[j]
__global__ void __launch_bounds__(TX_THREADS) k_text_lines(const char* text, u64 bytes, const u32* base, u32* nl) {
    const u64 t = (u64)blockIdx.x * TX_THREADS + threadIdx.x;
    const u64 lo = 16 * t;
    if (lo >= bytes) return;
    u32 at = base[t];
    const u64 hi = lo + 16 < bytes ? lo + 16 : bytes;
    for (u64 i = lo; i < hi; ++i)
        if (text[i] == '\n') nl[at++] = (u32)i;
}

// one thread per record: its header and sequence lines, whether the read is kept (getReads' filters)
__global__ void __launch_bounds__(TX_THREADS) k_text_records(const char* text, u64 bytes, const u32* nl, u32 n_lines, u32 lpr,
                                                             u32 k, u32* hoff, u32* hlen, u32* soff, u32* slen, u32* keep,
                                                             unsigned long long* klen, TextMeta* meta) {
    const u32 i = blockIdx.x * TX_THREADS + threadIdx.x;
    const u32 n_rec = n_lines / lpr;
    if (i == 0 && (n_lines % lpr != 0 || (bytes && (n_lines == 0 || (u64)nl[n_lines - 1] + 1 != bytes)))) meta->irregular = 1;
    if (i >= n_rec) return;
    const u32 l0 = lpr * i;
    const u32 hs = l0 ? nl[l0 - 1] + 1 : 0, he = nl[l0], ss = he + 1, se = nl[l0 + 1];
    bool ok = true;
    if (lpr == 2) ok = text[hs] == '>' && text[ss] != '>';  // (an empty line has '\n' there)
    if (!ok) meta->irregular = 1;
    const u32 sn = se - ss;
    bool good = sn > 2 && (lpr == 4 || sn > k);
    for (u32 j = 0; good && j < sn; ++j) {
        const char c = text[ss + j];
        good = c == 'A' || c == 'C' || c == 'G' || c == 'T' || c == 'N';
    }
    hoff[i] = hs;
    hlen[i] = he - hs;
    soff[i] = ss;
    slen[i] = sn;
    keep[i] = good ? 1u : 0u;
    klen[i] = good ? sn : 0ull;
}

// warp per record: the kept reads' characters, concatenated in input order (what bgr_slot_upload takes), their
// offsets, and the record each read came from
__global__ void __launch_bounds__(TX_THREADS) k_text_gather(const char* text, const u32* soff, const u32* slen, const u32* keep,
                                                            const u32* ridx, const unsigned long long* roff, u32 n_rec,
                                                            char* seq, u64* off, u32* rec_of, TextMeta* meta) {
    const u32 w = (blockIdx.x * TX_THREADS + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (w >= n_rec) return;
    if (w == n_rec - 1 && lane == 0) {
        const unsigned long long n = ridx[w] + keep[w], chars = roff[w] + (keep[w] ? slen[w] : 0u);
        meta->n_reads = n;
        meta->chars = chars;
        off[n] = chars;
    }
    if (!keep[w]) return;
    const u32 r = ridx[w];
    const u64 o = roff[w];
    const char* src = text + soff[w];
    const u32 n = slen[w];
    for (u32 j = lane; j < n; j += 32) seq[o + j] = src[j];
    if (lane == 0) {
        off[r] = o;
        rec_of[r] = w;
    }
}

// ---- output
__device__ __forceinline__ u32 dec_digits(u32 u) {
    return 1u + (u >= 10u) + (u >= 100u) + (u >= 1000u) + (u >= 10000u) + (u >= 100000u) + (u >= 1000000u) +
           (u >= 10000000u) + (u >= 100000000u) + (u >= 1000000000u);
}
__device__ __forceinline__ int text_stream(u32 flags, bool split_no_overlap) {
    if (flags & BGR_F_MAPPED) return 0;
    return (split_no_overlap && !(flags & BGR_F_OVERLAP_FOUND)) ? 2 : 1;
}

// bytes every read puts into each of the three files
__global__ void __launch_bounds__(TX_THREADS) k_text_sizes(BatchArgs a, const u32* rec_of, const u32* hlen, u32 split_no_overlap,
                                                           u32 correction, unsigned long long* sz /* [3][n] */) {
    const u32 r = blockIdx.x * TX_THREADS + threadIdx.x;
    if (r >= a.n) return;
    const bgr_result res = a.res[r];
    const u32 flags = res.flags_len & 0xFFu, len = res.flags_len >> 8;
    const int o = text_stream(flags, split_no_overlap != 0);
    const u32 rl = (u32)(a.off[r + 1] - a.off[r]);
    unsigned long long n = (unsigned long long)hlen[rec_of[r]] + 2;
    if (o == 0 && !correction) {
        const i32* ids = a.pool + res.path_off;
        for (u32 j = 0; j < len; ++j) {
            const i32 v = ids[j];
            n += dec_digits(v < 0 ? (u32)(-(i64)v) : (u32)v) + (v < 0 ? 2u : 1u);
        }
    } else {
        n += rl;
    }
#pragma unroll
    for (int f = 0; f < 3; ++f) sz[(size_t)f * a.n + r] = (f == o) ? n : 0ull;
}

__global__ void __launch_bounds__(TX_THREADS) k_text_format(BatchArgs a, const char* text, const u32* rec_of, const u32* hoff,
                                                            const u32* hlen, u32 split_no_overlap, u32 correction,
                                                            const unsigned long long* at /* [3][n] */, char* out0, char* out1,
                                                            char* out2) {
    const u32 r = blockIdx.x * TX_THREADS + threadIdx.x;
    if (r >= a.n) return;
    const bgr_result res = a.res[r];
    const u32 flags = res.flags_len & 0xFFu, len = res.flags_len >> 8;
    const int o = text_stream(flags, split_no_overlap != 0);
    char* p = (o == 0 ? out0 : o == 1 ? out1 : out2) + at[(size_t)o * a.n + r];
    const u32 rec = rec_of[r];
    const char* h = text + hoff[rec];
    const u32 hn = hlen[rec];
    for (u32 j = 0; j < hn; ++j) *p++ = h[j];
    *p++ = '\n';
    const u64 ro = a.off[r];
    const u32 rl = (u32)(a.off[r + 1] - ro);
    if (o == 0 && correction) {  // header\ncorrected\n (alignerGreedy.cpp:399)
        const u64* w = a.corr + corrected_word_base(ro, r);
        for (u32 j = 0; j < rl; ++j) *p++ = "ACGT"[(w[j >> 5] >> (62 - 2 * (j & 31))) & 3];
    } else if (o == 0) {  // printPath: every int followed by '.'
        const i32* ids = a.pool + res.path_off;
        for (u32 j = 0; j < len; ++j) {
            const i32 v = ids[j];
            u32 u = v < 0 ? (u32)(-(i64)v) : (u32)v;
            if (v < 0) *p++ = '-';
            const u32 d = dec_digits(u);
            for (u32 q = d; q-- > 0;) {
                p[q] = (char)('0' + u % 10);
                u /= 10;
            }
            p += d;
            *p++ = '.';
        }
    } else {
        const char* s = a.seq + ro;
        for (u32 j = 0; j < rl; ++j) *p++ = s[j];
    }
    *p = '\n';
}
#endif  // __CUDACC__

}  // namespace bgr
