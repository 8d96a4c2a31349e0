// sm_100a kernels of the read-to-graph mapping path.
//
//   k_scan_tiles    the seed scan (replaces getNOverlap / getListOverlap + the BooPHF lookups, reference
//                   aligner.cpp:318-378, BooPHF.h:783). Persistent CTAs walk tiles of consecutive reads:
//                   the tile's characters arrive in shared memory by one TMA bulk copy (double
//                   buffered), every thread takes one CHUNK = 32 consecutive windows of one read: packs
//                   the 64 characters under it to 2-bit words, rolls the forward / reverse-complement
//                   (k-1)-mer codes over them, tests each canonical code in the Bloom filter (in shared
//                   memory when the index is small, else global/L2) and verifies the few survivors in the
//                   key table. Output: packed words + an EXACT hit bitmask per read.
//   k_pack_scan     warp per read, any length, N-aware: the general form of the same scan, used for the
//                   reads the tile kernel hands over (reads containing non-ACGT characters, tiles that do
//                   not fit the shared-memory buffer).
//   k_extend_greedy thread per read: alignReadGreedy and its four walkers (reference
//                   alignerGreedy.cpp:35-57,167-364) on packed words, mismatches by XOR + popc.
//   k_extend_exhaustive  thread per read: alignReadExhaustive / alignReadExhaustivePath
//                   (alignerExhaustive.cpp:35-259, alignerPaths.cpp:66-88,127-175,289-460) as an
//                   iterative branch-and-bound depth-first search.
//
// The per-read bodies are __host__ __device__ so that tests/emu can single-step the very same
// code on the CPU (there is no GPU in the build container). The product never runs them on the
// host: libbgreat_b200.so only ever launches the kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/bgreat_b200.h"
#include "bgr_common.h"

namespace bgr {

#if defined(__CUDA_ARCH__)
#define BGR_POPCLL(x) __popcll(x)
#define BGR_FFS(x) __ffs(x)
#define BGR_CLZ(x) __clz(x)
#else
#define BGR_POPCLL(x) __builtin_popcountll(x)
#define BGR_FFS(x) __builtin_ffs(x)
#define BGR_CLZ(x) __builtin_clz(x)
#endif

struct DevIndex {
    const u64* keys;
    const Payload* payload;
    const UnitigMeta* meta;
    const u64* pool;
    u64 n_buckets;
    const u32* bloom;
    u32 n_bloom;
    u32 bloom_two_hashes;
    u32 k;
};

// Written by the extension kernel; reset before every (re-)run of it.
struct DevCounters {
    unsigned long long pool_cursor;  // ints the batch's paths need (written by the compaction)
    unsigned long long c[BGR_C_COUNT];
    u32 scratch_overflow;  // some path / search stack did not fit the per-read scratch
    u32 pool_overflow;     // the path pool was too small (pool_cursor = ints needed)
    u32 ext_cursor;        // next hit-list entry to hand to a warp of the extension kernel
    u32 correct_error;     // correction mode: a path does not spell a sequence as long as its read
};
// Written by the pack + seed-scan kernel.
struct ScanCounters {
    unsigned long long c[BGR_C_COUNT];
    u32 list_n;  // reads handed to the extension kernel (those with at least one border hit)
    u32 slow_n;  // reads the tile kernel handed to the general scan kernel
    u32 pad[2];
};

struct BatchArgs {
    const char* seq;  // concatenated read characters (null when the batch was uploaded packed)
    const u64* off;   // n+1 character offsets
    u32 n;
    // batch uploaded as a 2-bit stream (bgr_slot_upload_packed): word w = characters [32w, 32w+32) of the
    // concatenation, first character most significant. Reads holding a character outside ACGT are flagged in
    // `bad` (bit r%32 of word r/32) and come as characters on the side: exception j is read exc_idx[j], its
    // characters are exc_seq[exc_off[j] .. exc_off[j+1]).
    const u64* pk;
    const u32* bad;
    const u32* exc_idx;
    const u64* exc_off;
    const char* exc_seq;
    u32 n_exc;
    u64* F;     // 2-bit words, per read: [pad][ceil(len/32) data words][pad], see read_word_base
    u64* NM;    // N mask in the same layout (0b11 per non-ACGT base); written only for reads with N
    u32* H1;    // hit mask of the forward scan, bit i of word (i/32): window i is a border
    u32* H2;    // hit mask for the scan of the reverse-complemented read (only for reads with N)
    u32* info;  // per read: bit0 has N, bit1 any H1 hit, bit2 any H2 hit
    bgr_result* res;
    // extension scratch, scratch_ints per listed read: entry j of list row t at scratch[t*sc_row + j*sc_ent].
    // Interleaved (sc_row 1, sc_ent n: a warp's accesses to the same entry of 32 rows coalesce; the greedy walk
    // only appends) or row by row (sc_row scratch_ints, sc_ent 1: a row stays within a few cache lines; the
    // exhaustive search reads and rewrites its frames and id lists all the time).
    i32* scratch;
    u32 scratch_ints;
    u32 sc_row, sc_ent;
    i32* pool;
    u64 pool_cap;
    DevCounters* cnt;
    ScanCounters* scnt;
    u32* list;       // indices of the reads the extension kernel has to look at (unordered)
    u32* slow_list;  // reads left to the general scan kernel (null: it scans every read)
    u32 tile_r;      // reads per tile of the tile kernel (even, <= ST_MAX_R)
    u32 tile_cb;     // bytes of one character buffer of the tile kernel
    u32 errors, effort, mode;
    u32 ext_lazy;  // greedy extension: lanes that must wait before hand-in / refill / seed selection run
    u64* corr;  // correction mode (-c): corrected reads, 2 bits per base, read r at word corrected_word_base(off[r], r)
};

static const u32 INFO_HAS_N = 1u, INFO_HIT1 = 2u, INFO_HIT2 = 4u;
// internal bit of bgr_result.flags_len between the greedy kernel and the path compaction: the first
// entries of the scratch row are pathBegin in walk order (count in the row's last entry) and must be
// reversed on the way into the pool. Never leaves the device.
static const u32 RES_REV_HEAD = 0x08u;

// ------------------------------------------------------------------------------ helpers

struct __align__(32) u64x4 {
    u64 a, b, c, d;
};

// one 32-byte sector = one bucket of 4 keys, read-only path
BGR_HD u64x4 ld_bucket(const u64* p) {
    u64x4 r;
#if defined(__CUDA_ARCH__)
    asm volatile("ld.global.nc.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(r.a), "=l"(r.b), "=l"(r.c), "=l"(r.d) : "l"(p));
#else
    r.a = p[0];
    r.b = p[1];
    r.c = p[2];
    r.d = p[3];
#endif
    return r;
}

// 32 bases starting at base p (left-aligned). W[p/32 + 1] must be readable.
// the same for two words already loaded: hi holds base 32*(p/32), lo the 32 bases after it
BGR_HD u64 align_bases(u64 hi, u64 lo, u32 s /* 2 * (p % 32) */) {
#if defined(__CUDA_ARCH__)
    // 32-bit funnel shifts over the four half words: a variable 64-bit shift costs twice as much
    const u32 h1 = (u32)(hi >> 32), h0 = (u32)hi, l1 = (u32)(lo >> 32), l0 = (u32)lo;
    const bool big = s >= 32;
    const u32 a = big ? h0 : h1, b = big ? l1 : h0, c = big ? l0 : l1;
    return ((u64)__funnelshift_l(b, a, s) << 32) | __funnelshift_l(c, b, s);
#else
    return (hi << s) | ((lo >> 1) >> (63 - s));
#endif
}
BGR_HD u64 ld_bases(const u64* W, i64 p) {
    const i64 w = p >> 5;
    return align_bases(W[w], W[w + 1], (u32)(p & 31) * 2);
}

// slot index of `key`, or -1. A key lives in the first bucket at/after its home bucket that had a
// free slot when it was inserted; slots of a bucket fill in order, so "last slot empty" ends the probe.
BGR_HD i64 index_find(const DevIndex& ix, u64 key) {
    u64 b = bucket_of(key, ix.n_buckets);
    for (;;) {
        u64x4 v = ld_bucket(ix.keys + b * 4);
        if (v.a == key) return (i64)(b * 4);
        if (v.b == key) return (i64)(b * 4 + 1);
        if (v.c == key) return (i64)(b * 4 + 2);
        if (v.d == key) return (i64)(b * 4 + 3);
        if (v.d == EMPTY_KEY) return -1;
        if (++b == ix.n_buckets) b = 0;
    }
}

// exact membership of a canonical (k-1)-mer in the border key set
BGR_HD bool index_has(const DevIndex& ix, u64 key) { return index_find(ix, key) >= 0; }

// ------------------------------------------------------------------------------ pack + seed scan

static const int PS_WARPS = 8;    // warps per block
static const int PS_SWORDS = 21;  // shared words per warp and array: reads up to 576 bases stay on chip

// 4 characters (one 32-bit word, first character in the low byte) -> 8 bits of 2-bit codes (first
// character most significant) and 8 bits of N mask (0b11 per character that is not A/C/G/T; in the
// reference every such character behaves like N: str2num -> 3, nuc2int/nuc2intrc -> 0,
// revCompChar -> 'A', and it never equals a unitig character).
BGR_HD void pack4(u32 x, u32& code8, u32& nm8) {
    // A=0x41 C=0x43 G=0x47 T=0x54: ((c>>1)^(c>>2))&3 = 0,1,2,3
    u32 t = ((x >> 1) ^ (x >> 2)) & 0x03030303u;
    // expected character for that code: 0x41 + {0,2,6,0x13}
    u32 b0 = t & 0x01010101u, b1 = (t >> 1) & 0x01010101u;
    u32 expect = 0x41414141u + 2u * b0 + 6u * b1 + 11u * (b0 & b1);
    u32 d = x ^ expect;
    // 0x80 in every byte of d that is non-zero
    u32 nz = (((d & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | d) & 0x80808080u;
    u32 n3 = (nz >> 7) * 3u;  // 0 or 3 per byte
    t &= ~n3;                 // non-ACGT encodes as 0 (the rolling encoders' default, utils.cpp:132-151)
    // gather the four 2-bit fields: byte0 -> bits 7:6 ... byte3 -> bits 1:0
    code8 = (t * 0x40100401u) >> 24;
    nm8 = (n3 * 0x40100401u) >> 24;
}

// byte of the packed arrays that holds characters [c0, c0+4) of a read (c0 % 4 == 0): word 0 is the
// front pad, bases are big-endian inside a word
BGR_HD u32 packed_byte(u32 c0) { return 8 * (1 + (c0 >> 5)) + 7 - ((c0 >> 2) & 7); }

// Seed-scan test for window i (bases [i, i+k-1)) of a read: is its canonical (k-1)-mer a unitig
// border? D / ND: data words of the read (D[-1] is the front pad).
//   greedy modes: getNOverlap's rolling encoders (aligner.cpp:345-378). The forward code takes N as
//     3 in the first window (str2num) and as 0 when rolled in later (nuc2int), i.e. N reads 3 only at
//     absolute positions < k-1; the rolling RC stream takes N as 3 everywhere (nuc2intrc default 0 =
//     complement of 3). hit1 is that test; hit2 is the test for the same window of the
//     reverse-complemented read (N->A, utils.cpp:52-73), needed only when the read has N.
//   exhaustive modes: getListOverlap stores the same forward code; getBegin/getEnd canonicalise it
//     with a true rcb (aligner.cpp:149,211), so hit1 tests min(num, rcb(num)).
BGR_HD void scan_window(const DevIndex& ix, const u64* D, const u64* ND, bool anyN, bool exhaustive, u32 i,
                        bool& hit1, bool& hit2) {
    const u32 km1 = ix.k - 1, sh = 64 - 2 * km1;
    u64 x0 = ld_bases(D, i) >> sh;
    u64 num = x0, x2 = x0;
    if (anyN) {
        u64 nm = ld_bases(ND, i) >> sh;
        x2 = x0 | nm;
        u64 first = (i < km1) ? (nm & ~((1ull << (2 * i)) - 1)) : 0ull;
        num = x0 | first;
    }
    u64 rc = rcb(exhaustive ? num : x2, km1);
    u64 rep = num < rc ? num : rc;
    hit1 = index_has(ix, rep);
    hit2 = false;
    if (anyN && !exhaustive) {
        u64 rep2 = x2 < rc ? x2 : rc;
        hit2 = (rep2 == rep) ? hit1 : index_has(ix, rep2);
    }
}

// What the seed scan already knows about a read that has no border hit at all: greedy ->
// getNOverlap returns an empty list -> ++noOverlapRead (alignerGreedy.cpp:37); exhaustive -> every
// window fails -> ++notAligned (alignerExhaustive.cpp:55). A read of exactly k-1 bases maps trivially
// in exhaustive mode ([0,0]) and is therefore always passed on.
BGR_HD u32 no_hit_flags(bool exhaustive) {
    return exhaustive ? (BGR_F_OVERLAP_FOUND | BGR_F_CNT_NOTALIGNED) : BGR_F_CNT_NOOVERLAP;
}

// spread the 8 bits of x to bit positions 0,4,8,...,28
BGR_HD u32 spread8x4(u32 x) {
    x = (x | (x << 12)) & 0x000F000Fu;
    x = (x | (x << 6)) & 0x03030303u;
    x = (x | (x << 3)) & 0x11111111u;
    return x;
}

static const u32 PS_WPL = 4;  // consecutive windows per lane and round: 128 windows per round

// Fast path of the seed scan (read without N): PS_WPL consecutive windows starting at w0. The first
// one is extracted, the following ones roll the forward and the reverse-complement code like
// update / updateRC do (aligner.cpp:305-315). Bit j of the result = window w0+j is a border.
BGR_HD u32 scan_lane_rolling(const DevIndex& ix, const u64* D, u32 w0, u32 nwin) {
    const u32 km1 = ix.k - 1, sh = 64 - 2 * km1, top = 2 * (km1 - 1);
    const u64 mask = ~0ull >> sh;
    u32 hits = 0;
    if (w0 < nwin) {
        u64 num = ld_bases(D, w0) >> sh;
        u64 nxt = ld_bases(D, w0 + km1);  // the bases that follow the first window
        u64 rc = rcb(num, km1);
#pragma unroll
        for (u32 j = 0; j < PS_WPL; ++j) {
            if (w0 + j < nwin) {
                u64 rep = num < rc ? num : rc;
                hits |= (u32)index_has(ix, rep) << j;
            }
            u64 nb = nxt >> 62;
            nxt <<= 2;
            num = ((num << 2) | nb) & mask;
            rc = (rc >> 2) | ((3 - nb) << top);
        }
    }
    return hits;
}

// Lane l holds windows 4l..4l+3; ballot[j] has window 4l+j in bit l. The hit mask wants window i in
// bit i%32 of word i/32: word w = OR_j spread(byte w of ballot[j]) << j.
BGR_HD u32 transpose_hits(const u32 ballot[PS_WPL], u32 w) {
    u32 word = 0;
#pragma unroll
    for (u32 j = 0; j < PS_WPL; ++j) word |= spread8x4((ballot[j] >> (8 * w)) & 0xFFu) << j;
    return word;
}

#if defined(__CUDACC__)
// Pack one read into F/N (shared or global words, one call site each so that the address space is
// known at compile time), then scan its windows. Returns (any hit in H1, any in H2, has N).
// src: the read's characters, or null = take the bases from the batch's 2-bit stream at base o (such a
// read has no N: the host flags those and sends them as characters).
__device__ __forceinline__ u32 pack_and_scan(const DevIndex& ix, const BatchArgs& a, u64* F, u64* N,
                                             const unsigned char* src, const u64 o, const u32 len, const u64 wb,
                                             const bool exhaustive, const u32 lane) {
    const u32 km1 = ix.k - 1;
    const u32 nw = (len + 31) >> 5;
    if (lane == 0) {  // pad words: one in front, two behind (the rolling scan looks one word ahead)
        F[0] = 0;
        F[nw + 1] = 0;
        F[nw + 2] = 0;
    }
    if (!src) {  // packed upload: realign the read's words out of the stream
        for (u32 j = lane; j < nw; j += 32) {
            u64 w = ld_bases(a.pk, (i64)(o + 32ull * j));
            const u32 rem = len - 32 * j;
            if (rem < 32) w &= ~(~0ull >> (2 * rem));
            F[1 + j] = w;
        }
    }
    // ---- pack: 256 characters per pass, 8 per lane (a 150-base read takes one pass on 19 lanes)
    bool anyN = false;
    for (u32 base = 0; src && base < nw * 32; base += 256) {
        const u32 c0 = base + 8 * lane;
        if (c0 < nw * 32) {
            u32 x0 = 0x41414141u, x1 = 0x41414141u;
            if (c0 < len) {
                // 8 unaligned bytes out of three aligned words
                const uintptr_t ad = (uintptr_t)(src + c0);
                const u32* wp = (const u32*)(ad & ~(uintptr_t)3);
                const u32 bs = (u32)(ad & 3) * 8;
                const u32 w0 = __ldg(wp), w1 = __ldg(wp + 1), w2 = bs ? __ldg(wp + 2) : 0u;
                x0 = __funnelshift_r(w0, w1, bs);
                x1 = __funnelshift_r(w1, w2, bs);
                const u32 left = len - c0;  // characters from c0 on that belong to the read
                if (left < 8) {
                    // bytes past the end of the read become 'A' (code 0, not N)
                    const u32 k0 = left >= 4 ? 0xFFFFFFFFu : (1u << (8 * left)) - 1;
                    const u32 k1 = left <= 4 ? 0u : (1u << (8 * (left - 4))) - 1;
                    x0 = (x0 & k0) | (0x41414141u & ~k0);
                    x1 = (x1 & k1) | (0x41414141u & ~k1);
                }
            }
            u32 ca, na, cb, nb;
            pack4(x0, ca, na);
            pack4(x1, cb, nb);
            // chars c0..c0+3 go to the higher byte, c0+4..c0+7 to the byte below it
            *(unsigned short*)((unsigned char*)F + packed_byte(c0 + 4)) = (unsigned short)((ca << 8) | cb);
            anyN |= (na | nb) != 0;
        }
    }
    anyN = __any_sync(0xFFFFFFFFu, anyN);
    if (anyN) {
        // rare: write the N mask words too (same walk, mask bytes only)
        if (lane == 0) {
            N[0] = 0;
            N[nw + 1] = 0;
            N[nw + 2] = 0;
        }
        for (u32 base = 0; base < nw * 32; base += 128) {
            const u32 c0 = base + 4 * lane;
            if (c0 < nw * 32) {
                u32 x = 0x41414141u;
                if (c0 < len) {
                    const uintptr_t ad = (uintptr_t)(src + c0);
                    const u32* wp = (const u32*)(ad & ~(uintptr_t)3);
                    const u32 bs = (u32)(ad & 3) * 8;
                    const u32 w0 = __ldg(wp), w1 = bs ? __ldg(wp + 1) : 0u;
                    x = __funnelshift_r(w0, w1, bs);
                    const u32 left = len - c0;
                    if (left < 4) {
                        const u32 keep = (1u << (8 * left)) - 1;
                        x = (x & keep) | (0x41414141u & ~keep);
                    }
                }
                u32 code8, nm8;
                pack4(x, code8, nm8);
                ((unsigned char*)N)[packed_byte(c0)] = (unsigned char)nm8;
            }
        }
    }
    __syncwarp();
    // ---- seed scan
    const u64* D = F + 1;
    const u64* ND = N + 1;
    const u32 nwin = (len >= km1) ? len - km1 + 1 : 0;
    u32 any1 = 0, any2 = 0;
    if (!anyN) {
        for (u32 base = 0; base < nwin; base += 32 * PS_WPL) {
            const u32 hits = scan_lane_rolling(ix, D, base + PS_WPL * lane, nwin);
            u32 ballot[PS_WPL];
#pragma unroll
            for (u32 j = 0; j < PS_WPL; ++j) ballot[j] = __ballot_sync(0xFFFFFFFFu, (hits >> j) & 1u);
            // lanes 0..3 hold the words for windows base+32*lane ..
            const u32 word = transpose_hits(ballot, lane & 3);
            if (lane < 4 && base + 32 * lane < nwin) a.H1[wb + (base >> 5) + lane] = word;
            any1 |= ballot[0] | ballot[1] | ballot[2] | ballot[3];
        }
    } else {
        for (u32 base = 0; base < nwin; base += 32) {
            u32 i = base + lane;
            bool hit1 = false, hit2 = false;
            if (i < nwin) scan_window(ix, D, ND, true, exhaustive, i, hit1, hit2);
            u32 m1 = __ballot_sync(0xFFFFFFFFu, hit1);
            u32 m2 = __ballot_sync(0xFFFFFFFFu, hit2);
            any1 |= m1;
            any2 |= m2;
            if (lane == 0) {
                a.H1[wb + (base >> 5)] = m1;
                a.H2[wb + (base >> 5)] = m2;
            }
        }
    }
    return (anyN ? INFO_HAS_N : 0u) | (any1 ? INFO_HIT1 : 0u) | (any2 ? INFO_HIT2 : 0u);
}

template <int MINB>
__global__ void __launch_bounds__(PS_WARPS * 32, MINB) k_pack_scan(DevIndex ix, BatchArgs a) {
    __shared__ u64 sF[PS_WARPS][PS_SWORDS];
    __shared__ u64 sN[PS_WARPS][PS_SWORDS];
    __shared__ u32 s_list[PS_WARPS][32];
    const u32 lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const u32 nwarps = gridDim.x * PS_WARPS;
    const u32 km1 = ix.k - 1;
    const bool exhaustive = a.mode != BGR_MODE_GREEDY;
    unsigned long long windows = 0, n_reads = 0, n_nohit = 0;
    u32 n_list = 0;  // reads buffered in s_list[warp]

    // the reads the tile kernel handed over (slow_list), or every read of the batch (no list); after a
    // packed upload also the exception reads, whose characters came on the side
    const u32 n_slow = a.slow_list ? a.scnt->slow_n : a.n;
    const u32 n_items = n_slow + a.n_exc;
    for (u32 it = blockIdx.x * PS_WARPS + warp; it < n_items; it += nwarps) {
        u32 r;
        const unsigned char* src;
        if (it < n_slow) {
            r = a.slow_list ? a.slow_list[it] : it;
            if (a.pk && (a.bad[r >> 5] >> (r & 31) & 1u)) continue;  // an exception read: see below
            src = a.pk ? nullptr : (const unsigned char*)a.seq + a.off[r];
        } else {
            r = a.exc_idx[it - n_slow];
            src = (const unsigned char*)a.exc_seq + a.exc_off[it - n_slow];
        }
        const u64 o = a.off[r];
        const u32 len = (u32)(a.off[r + 1] - o);
        const u64 wb = read_word_base(o, r);
        const u32 nw = (len + 31) >> 5;
        const u32 nwin = (len >= km1) ? len - km1 + 1 : 0;
        u64* gF = a.F + wb;
        u64* gN = a.NM + wb;
        u32 info;
        __syncwarp();  // the previous read's shared words are no longer in use
        if (nw + 3 <= (u32)PS_SWORDS) {
            info = pack_and_scan(ix, a, sF[warp], sN[warp], src, o, len, wb, exhaustive, lane);
            // keep the packed words for the extension kernel
            const bool hasN = (info & INFO_HAS_N) != 0;
            const bool needed = (info & (INFO_HIT1 | INFO_HIT2)) != 0;
            if (needed)
                for (u32 j = lane; j < nw + 3; j += 32) {
                    gF[j] = sF[warp][j];
                    if (hasN) gN[j] = sN[warp][j];
                }
        } else {
            // long read: the words live in HBM; the warp reads back what it wrote
            info = pack_and_scan(ix, a, gF, gN, src, o, len, wb, exhaustive, lane);
        }
        const bool pass_on = (info & INFO_HIT1) || (exhaustive && len == km1);
        if (lane == 0) {
            a.info[r] = info;
            windows += nwin;
            n_reads += 1;
            if (pass_on) {
                s_list[warp][n_list] = r;
            } else {
                bgr_result res;
                res.flags_len = no_hit_flags(exhaustive);
                res.path_off = 0;
                a.res[r] = res;
                n_nohit += 1;
            }
        }
        if (pass_on) ++n_list;
        if (n_list == 32) {  // flush: one atomic per 32 listed reads
            u32 base = 0;
            if (lane == 0) base = atomicAdd(&a.scnt->list_n, 32u);
            base = __shfl_sync(0xFFFFFFFFu, base, 0);
            __syncwarp();
            a.list[base + lane] = s_list[warp][lane];
            n_list = 0;
        }
    }
    __syncwarp();
    if (n_list) {
        u32 base = 0;
        if (lane == 0) base = atomicAdd(&a.scnt->list_n, n_list);
        base = __shfl_sync(0xFFFFFFFFu, base, 0);
        if (lane < n_list) a.list[base + lane] = s_list[warp][lane];
    }
    if (lane == 0) {
        if (windows) atomicAdd(&a.scnt->c[BGR_C_WINDOWS], windows);
        if (n_reads) atomicAdd(&a.scnt->c[BGR_C_READS], n_reads);
        if (exhaustive) {
            if (windows) atomicAdd(&a.scnt->c[BGR_C_OVERLAPS], windows);  // overlaps += listOverlap.size()
            if (n_nohit) atomicAdd(&a.scnt->c[BGR_C_NOTALIGNED], n_nohit);
        } else if (n_nohit) {
            atomicAdd(&a.scnt->c[BGR_C_NOOVERLAP], n_nohit);
        }
    }
}
#endif  // __CUDACC__

// ------------------------------------------------------------------------------ tile scan: chunk logic
//
// A chunk is 32 consecutive windows of one read: windows [32l, 32l+count). It needs the bases
// [32l, 32l + count + k-2), at most 62 <= 64 of them: two packed words w0 (bases 32l..32l+31) and w1.

static const u32 TS_FLAG_HIT = 1u, TS_FLAG_BAD = 2u;

// 64 characters starting at p -> w0, w1 (2 bits per base, first base most significant). `bad` != 0
// iff some character is not A/C/G/T (then the codes are meaningless: the caller hands the read to
// the general kernel, which knows how the reference treats such characters).
BGR_HD void pack_span(const unsigned char* p, u64& w0, u64& w1, u32& bad) {
    u32 acc[4] = {0, 0, 0, 0};
    bad = 0;
#if defined(__CUDA_ARCH__)
    // aligned 32-bit loads + funnel shift: 17 loads for 16 groups of 4 characters
    const u32* wp = (const u32*)((uintptr_t)p & ~(uintptr_t)3);
    const u32 bs = (u32)((uintptr_t)p & 3) * 8;
    u32 prev = wp[0];
#pragma unroll
    for (int g = 0; g < 16; ++g) {
        const u32 nxt = wp[g + 1];
        const u32 x = __funnelshift_r(prev, nxt, bs);
        prev = nxt;
        const u32 t = ((x >> 1) ^ (x >> 2)) & 0x03030303u;  // A0 C1 G2 T3 per byte
        u32 sel = (t | (t >> 4)) & 0x00FF00FFu;
        sel = (sel | (sel >> 8)) & 0xFFFFu;                 // nibble i = code of character i
        bad |= x ^ __byte_perm(0x54474341u, 0u, sel);       // the character that code stands for
        acc[g >> 2] = (acc[g >> 2] << 8) | ((t * 0x40100401u) >> 24);
    }
#else
    for (int g = 0; g < 16; ++g) {
        u32 code8 = 0;
        for (int j = 0; j < 4; ++j) {
            const u32 c = base_code(p[4 * g + j]);
            if (c > 3) bad = 1;
            code8 = (code8 << 2) | (c & 3);
        }
        acc[g >> 2] = (acc[g >> 2] << 8) | code8;
    }
#endif
    w0 = ((u64)acc[0] << 32) | acc[1];
    w1 = ((u64)acc[2] << 32) | acc[3];
}

// zero the bases at and after `rem` (bases of the read left from the chunk's first base on)
BGR_HD void clip_span(u32 rem, u64& w0, u64& w1) {
    if (rem < 32) {
        w0 &= ~(~0ull >> (2 * rem));
        w1 = 0;
    } else if (rem < 64) {
        w1 = (rem == 32) ? 0ull : (w1 & ~(~0ull >> (2 * (rem - 32))));
    }
}

// 64 bases starting at base p of a 2-bit stream W (first base most significant): the packed-upload
// counterpart of pack_span. W[p/32 + 2] must be readable.
BGR_HD void packed_span(const u64* W, u64 p, u64& w0, u64& w1) {
    const u64 i = p >> 5;
    const u32 s = (u32)(p & 31) * 2;
    const u64 a = W[i], b = W[i + 1], c = W[i + 2];
    w0 = (a << s) | ((b >> 1) >> (63 - s));
    w1 = (b << s) | ((c >> 1) >> (63 - s));
}

// Bloom test of the chunk's windows: bit j = window 32l+j may be a border. The first window is taken
// from w0, the following ones roll the forward and the reverse-complement code like update / updateRC
// (aligner.cpp:305-315). TWO: word index and bit pattern come from two hashes (large filters). pat: the table
// of bloom_pattern values (shared memory on the device). KM1S: k-1 known at compile time (0: use km1_rt):
// with the loop unrolled every shift count and mask below is then a constant, which takes about a third of
// the integer instructions out of the loop (the scan is bound by the integer ALU pipe).
template <bool TWO, u32 KM1S>
BGR_HD u32 scan_chunk(const u32* bloom, u32 n_bloom, const u32* pat, u32 km1_rt, u64 w0, u64 w1, u32 count) {
    const u32 km1 = KM1S ? KM1S : km1_rt;
    const u32 sh = 64 - 2 * km1, top = 2 * (km1 - 1);
    const u64 mask = ~0ull >> sh;
    u64 num = w0 >> sh;
    u64 rc = rcb(num, km1);
    const u64 inc = (w0 << (2 * km1)) | (w1 >> sh);  // the bases that follow the first window
    u32 hits = 0;
    // All 32 window slots are tested, the slots past `count` are dropped at the end: no branch per
    // window, so the filter loads of several windows are in flight together (a slot past the end reads
    // some valid filter word: the index is always < n_bloom).
#pragma unroll
    for (u32 j = 0; j < 32; ++j) {
        // symmetric in (num, rc): no canonical minimum needed
        const u32 h1 = bloom_h1(num, rc);
        const u32 m = pat[bloom_pidx(TWO ? bloom_h2(num, rc) : h1)];
        const u32 w = bloom[bloom_word_of(h1, n_bloom)];
        hits |= ((w & m) == m ? 1u : 0u) << j;
        const u64 nb = (inc >> (62 - 2 * j)) & 3;
        num = ((num << 2) | nb) & mask;
        rc = (rc >> 2) | ((3 - nb) << top);
    }
    return count >= 32 ? hits : hits & ((1u << count) - 1);
}

// The windows that passed the filter, checked against the key table: the returned mask is exact.
BGR_HD u32 verify_chunk(const DevIndex& ix, u64 w0, u64 w1, u32 hits) {
    const u32 km1 = ix.k - 1, sh = 64 - 2 * km1;
    u32 m = hits;
    while (m) {
        const u32 j = (u32)BGR_FFS(m) - 1;
        m &= m - 1;
        const u64 x = (j ? ((w0 << (2 * j)) | (w1 >> (64 - 2 * j))) : w0) >> sh;
        const u64 rc = rcb(x, km1);
        if (!index_has(ix, x < rc ? x : rc)) hits &= ~(1u << j);
    }
    return hits;
}

#if defined(__CUDACC__)
// ------------------------------------------------------------------------------ tile scan kernel

static const int ST_THREADS = 1024;
static const u32 ST_MAX_R = 256;    // reads per tile (one prologue / epilogue thread each)
static const u32 ST_CHAR_SLACK = 128;  // bytes readable behind a character buffer (look-ahead of the last span)

__device__ __forceinline__ u32 smem_u32(const void* p) { return (u32)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(u64* bar, u32 count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(u64* bar, u32 bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(u64* bar, u32 parity) {
    u32 done;
    do {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    } while (!done);
}
// TMA bulk copy global -> shared (1-D, 16-byte granules), completion counted on an mbarrier; the read
// characters are used once, so they are marked evict-first in L2 (the filter and the index stay).
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, u32 bytes, u64* bar, u64 policy) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
            smem_u32(dst)),
        "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
        : "memory");
}

struct TileSmem {
    u32* bloom;
    u32* pat;  // BLOOM_PATTERNS filter bit patterns
    unsigned char* chars[2];
    u64* off;     // ST_MAX_R + 2
    u32* prefix;  // ST_MAX_R + 2
    u32* flags;   // ST_MAX_R + 2
    u64* mbar;    // 2
    u32* wsum;    // 32
};
__host__ __device__ inline size_t tile_smem_bytes(u32 bloom_words_in_smem, u32 tile_cb) {
    size_t b = ((size_t)bloom_words_in_smem * 4 + 127) & ~(size_t)127;
    b += BLOOM_PATTERNS * 4;
    b += 2 * ((size_t)tile_cb + ST_CHAR_SLACK);
    b += (ST_MAX_R + 2) * 8 + 2 * (ST_MAX_R + 2) * 4 + 2 * 8 + 32 * 4;
    return b + 128;
}

// One chunk of a tile (both schedules of the tile scan): chunk j of the tile whose reads start at r0, with the
// tile's offsets / chunk prefix sums / flag words in shared memory and its characters (or 2-bit words) at cb,
// whose first base is base gs of the batch. Packs, stores the words for the extension kernel, scans, verifies,
// stores the hit mask and flags the read. Returns the number of windows that passed the filter.
template <bool BLOOM_SMEM, bool PACKED, u32 KM1S>
__device__ __forceinline__ u32 tile_chunk(const DevIndex& ix, const BatchArgs& a, const u32* bloom, const u32* pat,
                                          const u32 km1, const unsigned char* cb, const u64 gs, const u64* s_off,
                                          const u32* s_prefix, u32* s_flags, const u32 cnt, const u32 r0, const u32 j) {
    // the read this chunk belongs to: last rd with prefix[rd] <= j
    u32 rd = 0;
#pragma unroll
    for (u32 sft = ST_MAX_R / 2; sft; sft >>= 1) {
        const u32 c = rd + sft;
        if (c < cnt && s_prefix[c] <= j) rd = c;
    }
    const u32 l = j - s_prefix[rd];
    const u64 ro = s_off[rd];
    const u32 rlen = (u32)(s_off[rd + 1] - ro);
    const u32 rnwin = rlen - km1 + 1;
    const u32 rchunks = (rnwin + 31) >> 5;
    const u32 r = r0 + rd;
    const u64 wb = read_word_base(ro, r);
    u64 w0, w1;
    u32 bad = 0;
    if (PACKED) packed_span((const u64*)cb, ro - gs + 32 * l, w0, w1);
    else pack_span(cb + (u32)(ro - gs) + 32 * l, w0, w1, bad);
    clip_span(rlen - 32 * l, w0, w1);
    // packed words for the extension kernel (layout of read_word_base: pad, words, 2 pads)
    u64* gF = a.F + wb;
    const u32 nw = (rlen + 31) >> 5;
    gF[1 + l] = w0;
    if (l == 0) gF[0] = 0;
    if (l + 1 == rchunks) {
        if (l + 1 < nw) gF[2 + l] = w1;
        gF[nw + 1] = 0;
        gF[nw + 2] = 0;
    }
    const u32 count = min(32u, rnwin - 32 * l);
    // the shared-memory tier is the single-hash tier (bloom_words_for)
    u32 hits = scan_chunk<!BLOOM_SMEM, KM1S>(bloom, ix.n_bloom, pat, km1, w0, w1, count);
    const u32 passed = (u32)__popc(hits);
    hits = verify_chunk(ix, w0, w1, hits);
    a.H1[wb + l] = hits;
    const u32 fl = (hits ? TS_FLAG_HIT : 0u) | (bad ? TS_FLAG_BAD : 0u);
    if (fl) atomicOr(&s_flags[rd], fl);
    return passed;
}

// PACKED: the batch came as a 2-bit stream (bgr_slot_upload_packed); a tile is then the stream's words under
// its reads, a chunk's two words are cut out of them with funnel shifts (packed_span) instead of being packed
// from characters, and the reads the host flagged as holding a non-ACGT character are left to k_pack_scan.
// KM1S: k-1 as a compile-time constant for the usual k = 31 (0: any k, taken from the index at run time).
template <bool BLOOM_SMEM, bool PACKED, u32 KM1S>
__global__ void __launch_bounds__(ST_THREADS, 1) k_scan_tiles(DevIndex ix, BatchArgs a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    TileSmem S;
    {
        unsigned char* q = smem_raw;
        S.bloom = (u32*)q;
        q += BLOOM_SMEM ? (((size_t)ix.n_bloom * 4 + 127) & ~(size_t)127) : 0;
        S.pat = (u32*)q;
        q += BLOOM_PATTERNS * 4;
        S.chars[0] = q;
        q += a.tile_cb + ST_CHAR_SLACK;
        S.chars[1] = q;
        q += a.tile_cb + ST_CHAR_SLACK;
        S.off = (u64*)q;
        q += (ST_MAX_R + 2) * 8;
        S.prefix = (u32*)q;
        q += (ST_MAX_R + 2) * 4;
        S.flags = (u32*)q;
        q += (ST_MAX_R + 2) * 4;
        S.mbar = (u64*)q;
        q += 16;
        S.wsum = (u32*)q;
    }
    const u32 tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const u32 R = a.tile_r, n = a.n, G = gridDim.x;
    const u32 ntiles = (n + R - 1) / R;
    const u32 km1 = KM1S ? KM1S : ix.k - 1;
    const bool exhaustive = a.mode != BGR_MODE_GREEDY;
    const u32* bloom = BLOOM_SMEM ? S.bloom : ix.bloom;
    for (u32 i = threadIdx.x; i < BLOOM_PATTERNS; i += ST_THREADS) S.pat[i] = bloom_pattern(i);

    u64 policy;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
    if (tid == 0) {
        mbar_init(&S.mbar[0], 1);
        mbar_init(&S.mbar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (BLOOM_SMEM) {
        const uint4* src = (const uint4*)ix.bloom;  // the blob pads the filter to a 256-byte boundary
        uint4* dst = (uint4*)S.bloom;
        for (u32 i = tid; i < (ix.n_bloom + 3) / 4; i += ST_THREADS) dst[i] = __ldg(src + i);
    }
    __syncthreads();

    // bytes the tile [lo, hi) of the character stream occupies in a buffer (16-byte granules); first
    // character of the buffer
    auto tile_start = [](u64 lo) -> u64 { return PACKED ? ((lo >> 5) & ~1ull) << 5 : (lo & ~15ull); };
    auto tile_bytes = [](u64 lo, u64 hi) -> u64 {
        if (PACKED) return (((((hi + 31) >> 5) + 1) & ~1ull) - ((lo >> 5) & ~1ull)) * 8;
        return ((hi + 15) & ~15ull) - (lo & ~15ull);
    };
    auto load_bounds = [&](u32 t, u64& lo, u64& hi) {
        lo = hi = 0;
        if (t < ntiles) {
            const u32 r0 = t * R;
            lo = a.off[r0];
            hi = a.off[min(n, r0 + R)];
        }
    };
    auto issue = [&](u32 buf, u64 lo, u64 hi) {
        const u64 bytes = tile_bytes(lo, hi);
        if (bytes != 0 && bytes <= a.tile_cb) {
            mbar_expect_tx(&S.mbar[buf], (u32)bytes);
            const void* src = PACKED ? (const void*)(a.pk + (tile_start(lo) >> 5)) : (const void*)(a.seq + tile_start(lo));
            bulk_g2s(S.chars[buf], src, (u32)bytes, &S.mbar[buf], policy);
        }
    };

    u32 tile = blockIdx.x;
    u64 nx_lo = 0, nx_hi = 0;  // thread 0: bounds of the tile after the current one
    if (tid == 0) {
        u64 lo, hi;
        load_bounds(tile, lo, hi);
        if (tile < ntiles) issue(0, lo, hi);
        load_bounds(tile + G, nx_lo, nx_hi);
    }
    u64 my_off = 0;  // off[first read of the tile + tid], fetched one tile ahead
    if (tile < ntiles && tid <= R) {
        const u64 idx = (u64)tile * R + tid;
        if (idx <= n) my_off = a.off[idx];
    }
    u32 uses[2] = {0, 0};  // completed TMA rounds per buffer (mbarrier phase parity)
    unsigned long long c_windows = 0;
    u32 c_reads = 0, c_nohit = 0, c_verify = 0;

    for (u32 it = 0; tile < ntiles; ++it, tile += G) {
        const u32 buf = it & 1;
        const u32 r0 = tile * R;
        const u32 cnt = min(R, n - r0);
        if (tid == 0) {  // next tile's characters start flowing now; its buffer was released by the last barrier
            if (tile + G < ntiles) issue(buf ^ 1, nx_lo, nx_hi);
            load_bounds(tile + 2 * G, nx_lo, nx_hi);
        }
        if (tid <= cnt) S.off[tid] = my_off;
        u64 next_off = 0;
        if (tile + G < ntiles && tid <= R) {
            const u64 idx = (u64)(tile + G) * R + tid;
            if (idx <= n) next_off = a.off[idx];
        }
        __syncthreads();

        // ---- per read: length, windows, chunks; exclusive scan of the chunk counts
        u64 o = 0;
        u32 len = 0, nwin = 0, chunks = 0;
        if (tid < cnt) {
            o = S.off[tid];
            len = (u32)(S.off[tid + 1] - o);
            nwin = (len >= km1) ? len - km1 + 1 : 0;
            chunks = (nwin + 31) >> 5;
            const u32 r = r0 + tid;
            S.flags[tid] = (PACKED && (a.bad[r >> 5] >> (r & 31) & 1u)) ? TS_FLAG_BAD : 0u;
        }
        if (tid < ST_MAX_R) {
            u32 incl = chunks;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const u32 x = __shfl_up_sync(0xFFFFFFFFu, incl, d);
                if (lane >= (u32)d) incl += x;
            }
            if (lane == 31) S.wsum[warp] = incl;
            chunks = incl - chunks;  // exclusive within the warp
        }
        __syncthreads();
        if (tid < ST_MAX_R) {
            u32 basew = 0;
#pragma unroll
            for (u32 w = 0; w < ST_MAX_R / 32; ++w)
                if (w < warp) basew += S.wsum[w];
            if (tid < cnt) S.prefix[tid] = basew + chunks;
            if (tid == cnt - 1) S.prefix[cnt] = basew + chunks + ((nwin + 31) >> 5);
        }
        const u64 t_lo = S.off[0], t_hi = S.off[cnt];
        const u64 t_bytes = tile_bytes(t_lo, t_hi);
        const bool fits = t_bytes <= a.tile_cb;
        __syncthreads();

        if (fits) {
            if (t_bytes != 0) {
                mbar_wait(&S.mbar[buf], uses[buf] & 1);
                ++uses[buf];
            }
            const u32 T = S.prefix[cnt];
            const unsigned char* cb = S.chars[buf];
            const u64 gs = tile_start(t_lo);
            for (u32 j = tid; j < T; j += ST_THREADS)
                c_verify += tile_chunk<BLOOM_SMEM, PACKED, KM1S>(ix, a, bloom, S.pat, km1, cb, gs, S.off, S.prefix, S.flags, cnt,
                                                                  r0, j);
        }
        __syncthreads();

        // ---- per read: classify, hand on
        if (tid < ST_MAX_R) {
            const bool valid = tid < cnt;
            const u32 r = r0 + tid;
            const u32 fl = valid ? S.flags[tid] : 0u;
            const bool slow = valid && (!fits || (fl & TS_FLAG_BAD));
            const bool hit = (fl & TS_FLAG_HIT) != 0;
            const bool pass_on = valid && !slow && (hit || (exhaustive && len == km1));
            if (valid && !slow) {
                a.info[r] = hit ? INFO_HIT1 : 0u;
                c_windows += nwin;
                c_reads += 1;
                if (!pass_on) {
                    bgr_result res;
                    res.flags_len = no_hit_flags(exhaustive);
                    res.path_off = 0;
                    a.res[r] = res;
                    c_nohit += 1;
                }
            }
            const u32 mp = __ballot_sync(0xFFFFFFFFu, pass_on);
            // reads for the general kernel's list; flagged reads of a packed batch reach it as exceptions
            const bool to_slow = PACKED ? (slow && !(fl & TS_FLAG_BAD)) : slow;
            const u32 ms = __ballot_sync(0xFFFFFFFFu, to_slow);
            u32 bp = 0, bs = 0;
            if (lane == 0) {
                if (mp) bp = atomicAdd(&a.scnt->list_n, (u32)__popc(mp));
                if (ms) bs = atomicAdd(&a.scnt->slow_n, (u32)__popc(ms));
            }
            bp = __shfl_sync(0xFFFFFFFFu, bp, 0);
            bs = __shfl_sync(0xFFFFFFFFu, bs, 0);
            const u32 lt = (1u << lane) - 1;
            if (pass_on) a.list[bp + __popc(mp & lt)] = r;
            if (to_slow) a.slow_list[bs + __popc(ms & lt)] = r;
        }
        my_off = next_off;
        // the next iteration's first barrier orders these reads of S.flags / S.off before they are rewritten
    }

    // ---- counters: one atomic per warp and counter
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        c_windows += __shfl_xor_sync(0xFFFFFFFFu, c_windows, d);
        c_reads += __shfl_xor_sync(0xFFFFFFFFu, c_reads, d);
        c_nohit += __shfl_xor_sync(0xFFFFFFFFu, c_nohit, d);
        c_verify += __shfl_xor_sync(0xFFFFFFFFu, c_verify, d);
    }
    if (lane == 0) {
        if (c_windows) atomicAdd(&a.scnt->c[BGR_C_WINDOWS], c_windows);
        if (c_reads) atomicAdd(&a.scnt->c[BGR_C_READS], (unsigned long long)c_reads);
        if (c_verify) atomicAdd(&a.scnt->c[BGR_C_VERIFIED], (unsigned long long)c_verify);
        if (exhaustive) {
            if (c_windows) atomicAdd(&a.scnt->c[BGR_C_OVERLAPS], c_windows);  // overlaps += listOverlap.size()
            if (c_nohit) atomicAdd(&a.scnt->c[BGR_C_NOTALIGNED], (unsigned long long)c_nohit);
        } else if (c_nohit) {
            atomicAdd(&a.scnt->c[BGR_C_NOOVERLAP], (unsigned long long)c_nohit);
        }
    }
}

// ------------------------------------------------------------------------------ tile scan, warp-specialised
//
// Same work as k_scan_tiles, other schedule. In k_scan_tiles all 32 warps meet at four barriers per tile, and
// between them the per-tile bookkeeping (offsets, chunk prefix sums, classification, hit-list append with its
// global atomic) runs on 8 warps while 24 wait: a quarter of all warp time is spent at those barriers
// (profiles/r01_ncu_k_scan_tiles_v8_5Mbp.txt). Here warp 0 does nothing but the bookkeeping, one tile ahead
// (prologue of tile i+1, then epilogue of tile i), and the other 31 warps do nothing but chunks. They meet
// through mbarriers per buffer -- full[b]: tile metadata (and TMA issued) in buffer b; done[b]: the workers are
// through with buffer b -- which a warp passes as soon as ITS condition holds, so a fast warp runs into the
// next tile instead of waiting for the slowest one.
struct TileMeta {
    u64 off[ST_MAX_R + 2];
    u32 prefix[ST_MAX_R + 2];
    u32 flags[ST_MAX_R + 2];
    u64 t_lo;
    u32 cnt, r0, T, fits, wait_tma, tma_parity;
};
static const u32 WS_WORKERS = ST_THREADS - 32;

__host__ __device__ inline size_t tile_ws_smem_bytes(u32 bloom_words_in_smem, u32 tile_cb) {
    size_t b = ((size_t)bloom_words_in_smem * 4 + 127) & ~(size_t)127;
    b += BLOOM_PATTERNS * 4;
    b += 2 * ((size_t)tile_cb + ST_CHAR_SLACK);
    b += 2 * ((sizeof(TileMeta) + 15) & ~(size_t)15) + 6 * 8;
    return b + 128;
}

__device__ __forceinline__ void mbar_arrive(u64* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

template <bool BLOOM_SMEM, bool PACKED, u32 KM1S>
__global__ void __launch_bounds__(ST_THREADS, 1) k_scan_tiles_ws(DevIndex ix, BatchArgs a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    u32 *s_bloom, *s_pat;
    unsigned char* s_chars[2];
    TileMeta* M;
    u64 *bar_tma, *bar_full, *bar_done;
    {
        unsigned char* q = smem_raw;
        s_bloom = (u32*)q;
        q += BLOOM_SMEM ? (((size_t)ix.n_bloom * 4 + 127) & ~(size_t)127) : 0;
        s_pat = (u32*)q;
        q += BLOOM_PATTERNS * 4;
        s_chars[0] = q;
        q += a.tile_cb + ST_CHAR_SLACK;
        s_chars[1] = q;
        q += a.tile_cb + ST_CHAR_SLACK;
        M = (TileMeta*)q;  // sizeof(TileMeta) is a multiple of 8; two of them, each padded to 16
        q += 2 * ((sizeof(TileMeta) + 15) & ~(size_t)15);
        bar_tma = (u64*)q;
        bar_full = bar_tma + 2;
        bar_done = bar_tma + 4;
    }
    auto meta = [&](u32 b) -> TileMeta& { return *(TileMeta*)((unsigned char*)M + b * ((sizeof(TileMeta) + 15) & ~(size_t)15)); };
    const u32 tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const u32 R = a.tile_r, n = a.n, G = gridDim.x;
    const u32 ntiles = (n + R - 1) / R;
    const u32 km1 = KM1S ? KM1S : ix.k - 1;
    const bool exhaustive = a.mode != BGR_MODE_GREEDY;
    const u32* bloom = BLOOM_SMEM ? s_bloom : ix.bloom;
    for (u32 i = tid; i < BLOOM_PATTERNS; i += ST_THREADS) s_pat[i] = bloom_pattern(i);
    if (tid == 0) {
        for (int b = 0; b < 2; ++b) {
            mbar_init(&bar_tma[b], 1);
            mbar_init(&bar_full[b], 1);
            mbar_init(&bar_done[b], WS_WORKERS / 32);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (BLOOM_SMEM) {
        const uint4* src = (const uint4*)ix.bloom;  // the blob pads the filter to a 256-byte boundary
        uint4* dst = (uint4*)s_bloom;
        for (u32 i = tid; i < (ix.n_bloom + 3) / 4; i += ST_THREADS) dst[i] = __ldg(src + i);
    }
    __syncthreads();
    const u32 my_tiles = blockIdx.x < ntiles ? (ntiles - blockIdx.x + G - 1) / G : 0;
    auto tile_start = [](u64 lo) -> u64 { return PACKED ? ((lo >> 5) & ~1ull) << 5 : (lo & ~15ull); };
    auto tile_bytes = [](u64 lo, u64 hi) -> u64 {
        if (PACKED) return (((((hi + 31) >> 5) + 1) & ~1ull) - ((lo >> 5) & ~1ull)) * 8;
        return ((hi + 15) & ~15ull) - (lo & ~15ull);
    };

    if (warp == 0) {
        // ================================================================== bookkeeper
        u64 policy;
        asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
        u32 tma_uses[2] = {0, 0};  // lane 0: TMA rounds issued per buffer (mbarrier phase parity)
        unsigned long long c_windows = 0;
        u32 c_reads = 0, c_nohit = 0;

        // metadata of local tile i into buffer i&1, its characters on their way, then full[i&1]
        auto prologue = [&](u32 i) {
            const u32 b = i & 1;
            TileMeta& m = meta(b);
            const u32 r0 = (blockIdx.x + i * G) * R;
            const u32 cnt = min(R, n - r0);
            for (u32 t = lane; t <= cnt; t += 32) m.off[t] = a.off[(u64)r0 + t];
            __syncwarp();
            u32 carry = 0;
            for (u32 base = 0; base < cnt; base += 32) {
                const u32 t = base + lane;
                u32 chunks = 0;
                if (t < cnt) {
                    const u32 len = (u32)(m.off[t + 1] - m.off[t]);
                    const u32 nwin = (len >= km1) ? len - km1 + 1 : 0;
                    chunks = (nwin + 31) >> 5;
                    const u32 r = r0 + t;
                    m.flags[t] = (PACKED && (a.bad[r >> 5] >> (r & 31) & 1u)) ? TS_FLAG_BAD : 0u;
                }
                u32 incl = chunks;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const u32 x = __shfl_up_sync(0xFFFFFFFFu, incl, d);
                    if (lane >= (u32)d) incl += x;
                }
                if (t < cnt) m.prefix[t] = carry + incl - chunks;
                carry += __shfl_sync(0xFFFFFFFFu, incl, 31);
            }
            if (lane == 0) {
                m.prefix[cnt] = carry;
                m.T = carry;
                m.cnt = cnt;
                m.r0 = r0;
                const u64 lo = m.off[0], hi = m.off[cnt];
                m.t_lo = lo;
                const u64 bytes = tile_bytes(lo, hi);
                const bool fits = bytes <= a.tile_cb;
                m.fits = fits ? 1u : 0u;
                m.wait_tma = 0;
                if (fits && bytes != 0) {
                    mbar_expect_tx(&bar_tma[b], (u32)bytes);
                    const void* src =
                        PACKED ? (const void*)(a.pk + (tile_start(lo) >> 5)) : (const void*)(a.seq + tile_start(lo));
                    bulk_g2s(s_chars[b], src, (u32)bytes, &bar_tma[b], policy);
                    m.wait_tma = 1;
                    m.tma_parity = tma_uses[b] & 1;
                    ++tma_uses[b];
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar_full[b]);
        };

        // classification of the tile's reads and the hit-list append: one atomic per tile
        auto epilogue = [&](u32 i) {
            TileMeta& m = meta(i & 1);
            const u32 cnt = m.cnt, r0 = m.r0;
            const bool fits = m.fits != 0;
            const u32 lt = (1u << lane) - 1;
            u32 mp[ST_MAX_R / 32];
            u32 total = 0;
#pragma unroll
            for (u32 q = 0; q < ST_MAX_R / 32; ++q) {
                const u32 t = 32 * q + lane;
                const bool valid = t < cnt;
                const u32 r = r0 + t;
                const u32 fl = valid ? m.flags[t] : 0u;
                const u32 len = valid ? (u32)(m.off[t + 1] - m.off[t]) : 0u;
                const u32 nwin = (len >= km1) ? len - km1 + 1 : 0;
                const bool slow = valid && (!fits || (fl & TS_FLAG_BAD));
                const bool hit = (fl & TS_FLAG_HIT) != 0;
                const bool pass_on = valid && !slow && (hit || (exhaustive && len == km1));
                if (valid && !slow) {
                    a.info[r] = hit ? INFO_HIT1 : 0u;
                    c_windows += nwin;
                    c_reads += 1;
                    if (!pass_on) {
                        bgr_result res;
                        res.flags_len = no_hit_flags(exhaustive);
                        res.path_off = 0;
                        a.res[r] = res;
                        c_nohit += 1;
                    }
                }
                mp[q] = __ballot_sync(0xFFFFFFFFu, pass_on);
                total += __popc(mp[q]);
                // reads for the general kernel's list (rare); flagged reads of a packed batch reach it as exceptions
                const bool to_slow = PACKED ? (slow && !(fl & TS_FLAG_BAD)) : slow;
                const u32 ms = __ballot_sync(0xFFFFFFFFu, to_slow);
                if (ms) {
                    u32 bs = 0;
                    if (lane == 0) bs = atomicAdd(&a.scnt->slow_n, (u32)__popc(ms));
                    bs = __shfl_sync(0xFFFFFFFFu, bs, 0);
                    if (to_slow) a.slow_list[bs + __popc(ms & lt)] = r;
                }
            }
            if (total) {
                u32 base = 0;
                if (lane == 0) base = atomicAdd(&a.scnt->list_n, total);
                base = __shfl_sync(0xFFFFFFFFu, base, 0);
#pragma unroll
                for (u32 q = 0; q < ST_MAX_R / 32; ++q) {
                    if ((mp[q] >> lane) & 1u) a.list[base + __popc(mp[q] & lt)] = r0 + 32 * q + lane;
                    base += __popc(mp[q]);
                }
            }
        };

        if (my_tiles) prologue(0);
        for (u32 i = 0; i < my_tiles; ++i) {
            // buffer (i+1)&1 is free: its last tile (i-1) was waited for and classified in the previous turn
            if (i + 1 < my_tiles) prologue(i + 1);
            mbar_wait(&bar_done[i & 1], (i >> 1) & 1);
            epilogue(i);
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            c_windows += __shfl_xor_sync(0xFFFFFFFFu, c_windows, d);
            c_reads += __shfl_xor_sync(0xFFFFFFFFu, c_reads, d);
            c_nohit += __shfl_xor_sync(0xFFFFFFFFu, c_nohit, d);
        }
        if (lane == 0) {
            if (c_windows) atomicAdd(&a.scnt->c[BGR_C_WINDOWS], c_windows);
            if (c_reads) atomicAdd(&a.scnt->c[BGR_C_READS], (unsigned long long)c_reads);
            if (exhaustive) {
                if (c_windows) atomicAdd(&a.scnt->c[BGR_C_OVERLAPS], c_windows);  // overlaps += listOverlap.size()
                if (c_nohit) atomicAdd(&a.scnt->c[BGR_C_NOTALIGNED], (unsigned long long)c_nohit);
            } else if (c_nohit) {
                atomicAdd(&a.scnt->c[BGR_C_NOOVERLAP], (unsigned long long)c_nohit);
            }
        }
    } else {
        // ================================================================== workers
        const u32 wtid = tid - 32;
        u32 c_verify = 0;
        for (u32 i = 0; i < my_tiles; ++i) {
            const u32 b = i & 1;
            TileMeta& m = meta(b);
            mbar_wait(&bar_full[b], (i >> 1) & 1);
            if (m.fits) {
                if (m.wait_tma) mbar_wait(&bar_tma[b], m.tma_parity);
                const u32 T = m.T, cnt = m.cnt, r0 = m.r0;
                const unsigned char* cb = s_chars[b];
                const u64 gs = tile_start(m.t_lo);
                for (u32 j = wtid; j < T; j += WS_WORKERS)
                    c_verify += tile_chunk<BLOOM_SMEM, PACKED, KM1S>(ix, a, bloom, s_pat, km1, cb, gs, m.off, m.prefix, m.flags,
                                                                      cnt, r0, j);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar_done[b]);
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) c_verify += __shfl_xor_sync(0xFFFFFFFFu, c_verify, d);
        if (lane == 0 && c_verify) atomicAdd(&a.scnt->c[BGR_C_VERIFIED], (unsigned long long)c_verify);
    }
}

#endif  // __CUDACC__

// ------------------------------------------------------------------------------ extension: shared pieces

// One pass over a read: pass 0 is the read as given (N mismatches everything: missmatchNumber
// compares characters, utils.cpp:154); pass 1 is reverseComplements(read) (utils.cpp:66), in which
// every N has become 'A'. Both are views of the same stored words:
//   pass 0 sequence  = F0 (N->0) with mask NM
//   pass 1 sequence  = reverse complement of F2 = F0|NM (N->3), no mask
struct ReadView {
    const u64* D;   // data words (D[-1] and D[nw] are zero pads)
    const u64* ND;  // N mask words
    i32 len;
    bool hasN;
    bool pass;
};

struct Oriented {
    u64 start;  // first base of the FORWARD unitig in the pool
    i64 len;
    i32 sid;  // signed id: negative = reverse complemented
    bool rc;
};

struct Cands {
    u32 raw[4];
    int n;
    u32 fwd_bit;
};

// getBegin / getEnd (aligner.cpp:147-267). The key's bucket and the payloads of that bucket's four slots
// (one 128-byte line) are requested together: a key sits in its home bucket nearly always, and the walk
// is a chain of dependent round trips, so the extra sectors are cheaper than a second trip.
BGR_HD Cands index_query(const DevIndex& ix, u64 bin, bool want_end) {
    Cands c;
    c.n = 0;
    c.fwd_bit = 0;
    u64 rc = rcb(bin, ix.k - 1);
    bool canon = bin <= rc;
    const u64 key = canon ? bin : rc;
    // getEnd: canonical -> right table, else left; getBegin: canonical -> left, else right
    bool right = (want_end == canon);
    const u64 b0 = bucket_of(key, ix.n_buckets);
    const u64x4 v = ld_bucket(ix.keys + b0 * 4);
    const u32* line = right ? ix.payload[b0 * 4].right : ix.payload[b0 * 4].left;
    i64 slot;
#if defined(__CUDA_ARCH__)
    const uint4 p0 = __ldg((const uint4*)line), p1 = __ldg((const uint4*)(line + 8)), p2 = __ldg((const uint4*)(line + 16)),
                p3 = __ldg((const uint4*)(line + 24));
    uint4 p;
    if (v.a == key) p = p0, slot = 0;
    else if (v.b == key) p = p1, slot = 1;
    else if (v.c == key) p = p2, slot = 2;
    else if (v.d == key) p = p3, slot = 3;
    else if (v.d == EMPTY_KEY) return c;
    else {
        slot = index_find(ix, key);  // overflowed into a later bucket
        if (slot < 0) return c;
        const u32* q = right ? ix.payload[slot].right : ix.payload[slot].left;
        p = __ldg((const uint4*)q);
    }
    c.raw[0] = p.x;
    c.raw[1] = p.y;
    c.raw[2] = p.z;
    c.raw[3] = p.w;
#else
    (void)v;
    (void)line;
    slot = index_find(ix, key);
    if (slot < 0) return c;
    const u32* q = right ? ix.payload[slot].right : ix.payload[slot].left;
    for (int j = 0; j < 4; ++j) c.raw[j] = q[j];
#endif
    c.fwd_bit = canon ? SLOT_FWD_CANON : SLOT_FWD_RC;
    // slots are scanned in order up to the first empty one (aligner.cpp:172-203)
    const bool s0 = (c.raw[0] & SLOT_ID_MASK) != 0, s1 = s0 && (c.raw[1] & SLOT_ID_MASK) != 0,
               s2 = s1 && (c.raw[2] & SLOT_ID_MASK) != 0, s3 = s2 && (c.raw[3] & SLOT_ID_MASK) != 0;
    c.n = (int)s0 + (int)s1 + (int)s2 + (int)s3;
    return c;
}

BGR_HD Oriented orient(const DevIndex& ix, const Cands& c, int i) {
    Oriented u;
    u32 id = c.raw[i] & SLOT_ID_MASK;
#if defined(__CUDA_ARCH__)
    const ulonglong2 m = __ldg((const ulonglong2*)&ix.meta[id]);
    u.start = m.x;
    u.len = (i64)(u32)m.y;
#else
    u.start = ix.meta[id].start;
    u.len = ix.meta[id].len;
#endif
    u.rc = (c.raw[i] & c.fwd_bit) == 0;
    u.sid = u.rc ? -(i32)id : (i32)id;
    return u;
}

// str2num of the first / last k-1 bases of the oriented unitig
BGR_HD u64 first_code(const DevIndex& ix, const Oriented& u) {
    u32 km1 = ix.k - 1;
    if (!u.rc) return ld_bases(ix.pool, (i64)u.start) >> (64 - 2 * km1);
    return rcb(ld_bases(ix.pool, (i64)u.start + u.len - km1) >> (64 - 2 * km1), km1);
}
BGR_HD u64 last_code(const DevIndex& ix, const Oriented& u) {
    u32 km1 = ix.k - 1;
    if (!u.rc) return ld_bases(ix.pool, (i64)u.start + u.len - km1) >> (64 - 2 * km1);
    return rcb(ld_bases(ix.pool, (i64)u.start) >> (64 - 2 * km1), km1);
}

// missmatchNumber (utils.cpp:154-168) between n bases of the oriented unitig starting at upos and n
// bases of the pass's read starting at rpos; the value is capped at cap+1 like the reference's
// early exit. XOR + popc on 2-bit words, N mask ORed in.
// HASN_: 1 = the read holds non-ACGT characters (rare): only then are the mask words loaded and applied; 0 =
// it does not; 2 = decided at run time (one copy of the loop: the exhaustive search calls this from four
// places and runs slower with two specialised copies at each).
template <int HASN_>
BGR_HD u32 mismatches_t(const DevIndex& ix, const ReadView& v, const Oriented& u, i64 upos, i32 rpos, i32 n, u32 cap) {
    // forward pool coordinates of the unitig segment, stored coordinates of the read segment
    const i64 ua = (i64)u.start + (u.rc ? u.len - upos - n : upos);
    const i32 sa = v.pass ? v.len - rpos - n : rpos;
    const bool flip = u.rc != v.pass;    // the read segment is reverse-complemented before comparing
    const bool HASN = HASN_ == 2 ? v.hasN : (HASN_ == 1);
    const bool orN = HASN && v.pass;    // pass 1 sequence is F0|NM
    const bool useM = HASN && !v.pass;  // pass 0 carries the N mask
    // The words of chunk c+1 are requested before chunk c is judged: the early exit below would otherwise
    // put one memory round trip per 32 bases on the walk's critical path (the kernel waits on exactly these
    // loads: profiles/r01_ncu_k_extend_greedy_v8_5Mbp.txt, long scoreboard).
    //   chunk at `done`: unitig bases [ua+done, +m); read bases [sa+done, +m), or, flipped, the m bases that
    //   end at sa+n-done
    auto read_at = [&](i32 done, i32 m) -> i32 { return flip ? sa + n - done - m : sa + done; };
    i32 m = n < 32 ? n : 32;
    u64 xu = n > 0 ? ld_bases(ix.pool, ua) : 0ull;
    u64 y = n > 0 ? ld_bases(v.D, read_at(0, m)) : 0ull;
    u64 ymask = (HASN && n > 0) ? ld_bases(v.ND, read_at(0, m)) : 0ull;
    u32 miss = 0;
    for (i32 done = 0; done < n; done += 32) {
        const i32 left = n - done - 32;  // bases after this chunk
        const i32 m_next = left < 32 ? left : 32;
        u64 xu_n = 0, y_n = 0, ymask_n = 0;
        if (left > 0) {
            xu_n = ld_bases(ix.pool, ua + done + 32);
            y_n = ld_bases(v.D, read_at(done + 32, m_next));
            if (HASN) ymask_n = ld_bases(v.ND, read_at(done + 32, m_next));
        }
        u64 xr, nm = 0;
        if (!flip) {
            xr = orN ? (y | ymask) : y;
            if (useM) nm = ymask;
        } else {
            xr = rc64(orN ? (y | ymask) : y) << (64 - 2 * m);
            if (useM) nm = rev2(ymask) << (64 - 2 * m);
        }
        u64 d = xu ^ xr;
        d = (d | (d >> 1) | nm) & 0x5555555555555555ull;
        if (m < 32) d &= ~(~0ull >> (2 * m));
        miss += (u32)BGR_POPCLL(d);
        if (miss > cap) return cap + 1;
        xu = xu_n;
        y = y_n;
        ymask = ymask_n;
        m = m_next;
    }
    return miss;
}
BGR_HD u32 mismatches(const DevIndex& ix, const ReadView& v, const Oriented& u, i64 upos, i32 rpos, i32 n, u32 cap) {
    return mismatches_t<2>(ix, v, u, upos, rpos, n, cap);
}
// the greedy walk's call site: N-free reads take a loop without any mask handling
BGR_HD u32 mismatches_split(const DevIndex& ix, const ReadView& v, const Oriented& u, i64 upos, i32 rpos, i32 n, u32 cap) {
    return v.hasN ? mismatches_t<1>(ix, v, u, upos, rpos, n, cap) : mismatches_t<0>(ix, v, u, upos, rpos, n, cap);
}

// forward code of window i of the pass's read, as the seed scans store it
// (getNOverlap / getListOverlap push `num`, aligner.cpp:326,354)
BGR_HD u64 seed_code(const DevIndex& ix, const ReadView& v, i32 i) {
    const u32 km1 = ix.k - 1, sh = 64 - 2 * km1;
    if (!v.pass) {
        u64 x = ld_bases(v.D, i) >> sh;
        if (v.hasN && (u32)i < km1) {
            u64 nm = ld_bases(v.ND, i) >> sh;
            x |= nm & ~((1ull << (2 * i)) - 1);
        }
        return x;
    }
    i32 j = v.len - (i32)km1 - i;  // same window in stored coordinates
    u64 x = ld_bases(v.D, j);
    if (v.hasN) x |= ld_bases(v.ND, j);
    return rcb(x >> sh, km1);
}

// per-read scratch: entry j of this read is base[j*stride]
struct Scratch {
    i32* base;
    u32 stride, cap;
    bool overflow;
    BGR_HD void set(u32 j, i32 x) {
        if (j < cap) base[(size_t)j * stride] = x;
        else overflow = true;
    }
    BGR_HD i32 get(u32 j) const { return j < cap ? base[(size_t)j * stride] : 0; }
};

// Iterator over the seed windows of one pass in the order the reference visits them. Pass 0 walks
// the set bits of the hit mask upwards. The windows of the reverse-complemented read are the same
// (k-1)-mers mirrored, so pass 1 walks the stored bits downwards (of H2 when the read has N: its RC
// copy has none and is scanned with the plain encoders). Positions are returned in pass coordinates.
struct SeedIter {
    const u32* H;
    i32 nwin, lim, cur;
    bool down;
    BGR_HD void init(const u32* h, i32 nwin_, bool down_, i32 lim_) {
        H = h;
        nwin = nwin_;
        down = down_;
        cur = 0;
        lim = lim_ < nwin_ ? lim_ : nwin_;
    }
    BGR_HD i32 next() {
        while (cur < lim) {
            i32 j = down ? nwin - 1 - cur : cur;
            if (!down) {
                u32 w = H[j >> 5] >> (j & 31);
                if (w == 0) {
                    cur += 32 - (j & 31);
                    continue;
                }
                cur += BGR_FFS(w) - 1;
            } else {
                u32 w = H[j >> 5] << (31 - (j & 31));
                if (w == 0) {
                    cur += (j & 31) + 1;
                    continue;
                }
                cur += BGR_CLZ(w);
            }
            if (cur >= lim) return -1;
            return cur++;
        }
        return -1;
    }
};

// ------------------------------------------------------------------------------ greedy extension

// checkBeginGreedy == mapOnLeftEndGreedy (alignerGreedy.cpp:268-319,167-218), tail recursion
// unrolled into a loop. Pushes to sc[n...]; returns false when the walk fails.
BGR_HD bool greedy_left(const DevIndex& ix, const ReadView& v, u64 code, i32 pos, u32 budget, Scratch& sc, u32& n,
                        u32& total) {
    const i64 km1 = ix.k - 1;
    u32 acc = 0;
    for (;;) {
        if (pos == 0) {
            sc.set(n++, 0);
            total = acc;
            return true;
        }
        Cands c = index_query(ix, code, true);
        u32 best = budget + 1;
        i32 best_sid = 0, offset = 0;
        bool ended = false, stepped = false;
        u64 next_code = 0;
        i64 next_L = 0;
        for (int i = 0; i < c.n; ++i) {
            Oriented u = orient(ix, c, i);
            const i64 L = u.len - km1;  // bases left of the overlap
            if (L >= pos) {
                u32 miss = mismatches(ix, v, u, L - pos, 0, pos, budget);
                if (miss == 0) {  // accepted at once (quirk 9)
                    sc.set(n++, u.sid);
                    sc.set(n++, (i32)(L - pos));
                    total = acc;
                    return true;
                } else if (miss < best) {
                    best = miss;
                    best_sid = u.sid;
                    ended = true;
                    offset = (i32)(L - pos);
                }
            } else {
                u32 miss = mismatches(ix, v, u, 0, pos - (i32)L, (i32)L, budget);
                if (miss == 0) {
                    sc.set(n++, u.sid);
                    code = first_code(ix, u);
                    pos -= (i32)L;
                    stepped = true;  // continue the walk with the budget untouched
                    break;
                } else if (miss < best) {
                    ended = false;
                    best = miss;
                    best_sid = u.sid;
                    next_L = L;
                    next_code = first_code(ix, u);
                }
            }
        }
        if (stepped) continue;
        if (best > budget) return false;
        sc.set(n++, best_sid);
        if (ended) {
            sc.set(n++, offset);
            total = acc + best;
            return true;
        }
        acc += best;
        budget -= best;
        code = next_code;
        pos -= (i32)next_L;
    }
}

// checkEndGreedy (first step, alignerGreedy.cpp:322-364) then mapOnRightEndGreedy (:221-265).
// In the first step comparisons start after the anchor overlap; in later steps the remaining read
// starts AT the overlap and the whole unitig is compared, so mismatches inside an overlap count
// again (quirk 8).
BGR_HD bool greedy_right(const DevIndex& ix, const ReadView& v, u64 code, i32 pos, u32 budget, Scratch& sc, u32& n,
                         u32& total) {
    const i64 km1 = ix.k - 1;
    u32 acc = 0;
    bool first = true;
    for (;;) {
        // first: readLeft = read[pos+k-1:], later: readLeft = read[pos:]
        const i32 rstart = first ? pos + (i32)km1 : pos;
        const i32 rl = v.len - rstart;
        if (first ? (rl == 0) : (rl < (i32)ix.k)) {
            total = acc;
            return true;
        }
        Cands c = index_query(ix, code, false);
        u32 best = budget + 1;
        i32 best_sid = 0;
        bool ended = false, stepped = false;
        u64 next_code = 0;
        i64 next_L = 0;
        for (int i = 0; i < c.n; ++i) {
            Oriented u = orient(ix, c, i);
            const i64 L = u.len - km1;
            const bool ends_here = L >= rl;
            u32 miss;
            if (first) {
                // unitig[k-1 : k-1+n) against readLeft[0:n)
                i32 n2 = ends_here ? rl : (i32)L;
                miss = mismatches(ix, v, u, km1, rstart, n2, budget);
            } else {
                // unitig[0:n) against read[pos : pos+n): n = |readLeft| if it ends here, else
                // min(|unitig|, |readLeft|) (substr is cut at the end of the read)
                i32 n2 = ends_here ? rl : (i32)(u.len < rl ? u.len : rl);
                miss = mismatches(ix, v, u, 0, rstart, n2, budget);
            }
            if (ends_here) {
                if (miss == 0) {
                    sc.set(n++, u.sid);
                    total = acc;
                    return true;
                } else if (miss < best) {
                    best = miss;
                    best_sid = u.sid;
                    ended = true;
                }
            } else {
                if (miss == 0) {
                    sc.set(n++, u.sid);
                    code = last_code(ix, u);
                    pos += (i32)L;
                    stepped = true;
                    break;
                } else if (miss < best) {
                    ended = false;
                    best = miss;
                    best_sid = u.sid;
                    next_L = L;
                    next_code = last_code(ix, u);
                }
            }
        }
        first = false;
        if (stepped) continue;
        if (best > budget) return false;
        sc.set(n++, best_sid);
        if (ended) {
            total = acc + best;
            return true;
        }
        acc += best;
        budget -= best;
        code = next_code;
        pos += (i32)next_L;
    }
}

struct ReadSetup {
    ReadView v;
    u64 wb;
    u32 info;
    i32 nwin;
};

BGR_HD ReadSetup setup_read(const DevIndex& ix, const BatchArgs& a, u32 r) {
    ReadSetup s;
    const u64 o = a.off[r];
    const u32 len = (u32)(a.off[r + 1] - o);
    s.wb = read_word_base(o, r);
    s.info = a.info[r];
    s.v.D = a.F + s.wb + 1;
    s.v.ND = a.NM + s.wb + 1;
    s.v.len = (i32)len;
    s.v.hasN = (s.info & INFO_HAS_N) != 0;
    s.v.pass = false;
    const u32 km1 = ix.k - 1;
    s.nwin = (len >= km1) ? (i32)(len - km1 + 1) : 0;
    return s;
}

// alignReadGreedy (alignerGreedy.cpp:35-57) for read r. The path is left in scratch as
// [0, nl) = pathBegin in push order (to be reversed) followed by [nl, n) = pathEnd.
//
// Written as ONE loop whose body is "settle the current level / pick the next seed / query the index"
// followed by "compare one candidate unitig": threads of a warp work on different reads and are at
// different points of their walks, but they all execute this same body, so the warp stays converged
// (the first version, a call tree of walkers, ran with 1.65 of 32 lanes active:
// profiles/r01_ncu_k_extend_greedy_v1.txt). The walkers greedy_left / greedy_right above are the
// readable statement of the same procedure and are kept for the emulator's cross-check.
enum { GP_SEED = 0, GP_LEFT = 1, GP_RIGHT_FIRST = 2, GP_RIGHT_MORE = 3, GP_DONE = 4 };

struct GreedyState {
    ReadSetup s;
    Scratch sc;
    SeedIter it;
    u32 flags, n, nl, tried, phase;
    // walk state
    u64 code, seed_c;
    i32 pos, seed_pos;
    u32 budget, acc, eb;

    BGR_HD bool done() const { return phase == GP_DONE; }

    BGR_HD void start(const DevIndex& ix, const BatchArgs& a, u32 r, i32* scratch_row) {
        s = setup_read(ix, a, r);
        sc.base = scratch_row;
        sc.stride = a.sc_ent;
        sc.cap = a.scratch_ints;
        sc.overflow = false;
        flags = 0;
        n = nl = 0;
        tried = 0;
        phase = GP_SEED;
        // effort 0: getNOverlap returns right after the first window, hit or not (aligner.cpp:366)
        it.init(a.H1 + s.wb, (s.info & INFO_HIT1) ? s.nwin : 0, false, a.effort ? s.nwin : 1);
        code = seed_c = 0;
        pos = seed_pos = 0;
        budget = acc = eb = 0;
    }

    BGR_HD void succeed() {
        flags |= BGR_F_MAPPED | BGR_F_CNT_ALIGNED | BGR_F_OVERLAP_FOUND | (s.v.pass ? BGR_F_RC : 0u);
        phase = GP_DONE;
    }
    BGR_HD void go_right(const BatchArgs& a) {  // the left walk is complete with `eb` mismatches
        nl = n;
        phase = GP_RIGHT_FIRST;
        budget = a.errors - eb;
        acc = 0;
        code = seed_c;
        pos = seed_pos;
    }

    // Seed selection, possibly switching to the reverse-complement pass: needed once per seed, i.e. rarely
    // compared with the levels of a walk, so the kernel runs it for several lanes at once (see k_extend_greedy).
    BGR_HD void pick_seed(const DevIndex& ix, const BatchArgs& a) {
        ReadView& v = s.v;
        const u32 max_seeds = a.effort ? a.effort : 1;
        for (int rep = 0; rep < 2 && phase == GP_SEED; ++rep) {
            i32 p = tried < max_seeds ? it.next() : -1;
            if (p >= 0) {
                ++tried;
                seed_c = code = seed_code(ix, v, p);
                seed_pos = pos = p;
                budget = a.errors;
                acc = 0;
                n = 0;
                phase = GP_LEFT;
            } else if (tried == 0) {  // getNOverlap found nothing: ++noOverlapRead, return {}
                flags |= BGR_F_CNT_NOOVERLAP;
                phase = GP_DONE;
            } else if (!v.pass) {  // every seed failed: rc=true, retry on the reverse complement
                flags |= BGR_F_OVERLAP_FOUND | BGR_F_RC;
                v.pass = true;
                const bool useH2 = v.hasN;
                const bool anyhit = (s.info & (useH2 ? INFO_HIT2 : INFO_HIT1)) != 0;
                it.init((useH2 ? a.H2 : a.H1) + s.wb, anyhit ? s.nwin : 0, true, a.effort ? s.nwin : 1);
                tried = 0;
            } else {
                flags |= BGR_F_CNT_NOTALIGNED;
                phase = GP_DONE;
            }
        }
    }

    // One turn of the loop = one level of a walk: query the index -> compare the <= 4 candidates ->
    // settle. Every lane of a warp that holds a seed runs these same stages every turn.
    BGR_HD void step(const DevIndex& ix, const BatchArgs& a) {
        if (phase == GP_SEED) pick_seed(ix, a);
        if (phase != GP_SEED && phase != GP_DONE) step_level(ix, a);
    }

    BGR_HD void step_level(const DevIndex& ix, const BatchArgs& a) {
        ReadView& v = s.v;
        const i64 km1 = ix.k - 1;
        // ---- ends of walks that need no index query
        if (phase == GP_LEFT && pos == 0) {  // checkBeginGreedy:269 / mapOnLeftEndGreedy:168
            sc.set(n++, 0);
            eb = acc;
            go_right(a);
        }
        if (phase == GP_RIGHT_FIRST && v.len - (pos + (i32)km1) == 0) succeed();  // checkEndGreedy:324
        if (phase == GP_RIGHT_MORE && v.len - pos < (i32)ix.k) succeed();         // mapOnRightEndGreedy:223
        if (phase == GP_DONE) return;
        // ---- candidates of this level: getEnd on the way left, getBegin on the way right
        const Cands c = index_query(ix, code, phase == GP_LEFT);
        u32 best = budget + 1;
        i32 best_sid = 0, offset = 0;
        i64 next_L = 0;
        u64 next_code = 0;
        bool ended = false, accepted = false;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            if (i < c.n && !accepted) {
                const Oriented u = orient(ix, c, i);
                const i64 L = u.len - km1;
                bool ends_here;
                i64 upos;
                i32 rpos, nn;
                if (phase == GP_LEFT) {
                    // unitig covers the rest: unitig[L-pos, L) vs read[0,pos); else unitig[0,L) vs read[pos-L,pos)
                    ends_here = L >= pos;
                    upos = ends_here ? L - pos : 0;
                    rpos = ends_here ? 0 : pos - (i32)L;
                    nn = ends_here ? pos : (i32)L;
                } else if (phase == GP_RIGHT_FIRST) {
                    // comparisons start after the anchor overlap
                    const i32 rstart = pos + (i32)km1, rl = v.len - rstart;
                    ends_here = L >= rl;
                    upos = km1;
                    rpos = rstart;
                    nn = ends_here ? rl : (i32)L;
                } else {
                    // the remaining read starts AT the overlap; whole unitig compared, cut at the read end
                    const i32 rl = v.len - pos;
                    ends_here = L >= rl;
                    upos = 0;
                    rpos = pos;
                    nn = ends_here ? rl : (i32)(u.len < rl ? u.len : rl);
                }
                const u32 miss = mismatches_split(ix, v, u, upos, rpos, nn, budget);
                if (miss < best) {  // first strictly smallest; a zero is taken at once (quirk 9)
                    best = miss;
                    best_sid = u.sid;
                    ended = ends_here;
                    offset = (i32)(L - pos);
                    next_L = L;
                    // (asking for this code together with the comparison's words, for every candidate, was
                    // measured and lost: 3.91 vs 3.66 ms at 5 Mbp, 6.65 vs 5.35 ms at 1 Gbp -- the extra
                    // random sectors cost more than the saved round trip)
                    if (!ends_here) next_code = (phase == GP_LEFT) ? first_code(ix, u) : last_code(ix, u);
                    accepted = miss == 0;
                }
            }
        }
        // ---- settle the level
        if (best > budget) {
            phase = GP_SEED;  // this seed failed
            return;
        }
        sc.set(n++, best_sid);
        // an exact candidate leaves the budget untouched, an inexact one pays for its mismatches
        if (ended) {
            if (phase == GP_LEFT) {
                sc.set(n++, offset);
                eb = acc + best;
                go_right(a);
            } else {
                succeed();
            }
        } else {
            acc += best;
            budget -= best;
            code = next_code;
            if (phase == GP_LEFT) pos -= (i32)next_L;
            else {
                pos += (i32)next_L;
                phase = GP_RIGHT_MORE;
            }
        }
    }

    // Path in final order in scratch [0, plen): pathBegin is reversed in place (alignerGreedy.cpp:48).
    // Returns the path length (0 when unmapped or when the scratch overflowed).
    BGR_HD u32 finish() {
        if (!(flags & BGR_F_MAPPED) || sc.overflow) return 0;
        for (u32 i = 0, j = nl; i + 1 < j; ++i) {
            --j;
            i32 x = sc.get(i), y = sc.get(j);
            sc.set(i, y);
            sc.set(j, x);
        }
        return n;
    }
    // Kernel form: pathBegin stays in walk order; its length goes into the last scratch entry and
    // k_path_gather reverses it while copying the path into the pool (the in-place loop above runs with
    // a few lanes of a warp active). The caller sets RES_REV_HEAD in the result's flags.
    BGR_HD u32 finish_deferred() {
        if (!(flags & BGR_F_MAPPED) || sc.overflow) return 0;
        if (n + 1 > sc.cap) {
            sc.overflow = true;
            return 0;
        }
        sc.set(sc.cap - 1, (i32)nl);
        return n;
    }
};

// whole read in one go (emulator, and the shape the persistent kernel unrolls over time)
BGR_HD u32 greedy_read(const DevIndex& ix, const BatchArgs& a, u32 r, Scratch& sc, u32& n, u32& nl) {
    GreedyState g;
    g.start(ix, a, r, sc.base);
    g.sc.stride = sc.stride;
    while (!g.done()) g.step(ix, a);
    sc.overflow = g.sc.overflow;
    n = (g.flags & BGR_F_MAPPED) ? g.n : 0;
    nl = g.nl;
    return g.flags;
}

// Reference-shaped version (walkers called in sequence); tests/emu checks that both agree.
BGR_HD u32 greedy_read_walkers(const DevIndex& ix, const BatchArgs& a, u32 r, Scratch& sc, u32& n, u32& nl) {
    ReadSetup s = setup_read(ix, a, r);
    ReadView& v = s.v;
    u32 flags = 0;
    n = nl = 0;
    const i32 lim = a.effort ? s.nwin : 1;
    const u32 max_seeds = a.effort ? a.effort : 1;
    for (int pass = 0; pass < 2; ++pass) {
        v.pass = pass != 0;
        const bool useH2 = v.pass && v.hasN;
        const bool anyhit = (s.info & (useH2 ? INFO_HIT2 : INFO_HIT1)) != 0;
        SeedIter it;
        it.init((useH2 ? a.H2 : a.H1) + s.wb, anyhit ? s.nwin : 0, v.pass, lim);
        u32 tried = 0;
        while (tried < max_seeds) {
            i32 pos = it.next();
            if (pos < 0) break;
            ++tried;
            u64 code = seed_code(ix, v, pos);
            n = 0;
            u32 eb = 0, ee = 0;
            if (!greedy_left(ix, v, code, pos, a.errors, sc, n, eb)) continue;
            nl = n;
            if (!greedy_right(ix, v, code, pos, a.errors - eb, sc, n, ee)) continue;
            flags |= BGR_F_MAPPED | BGR_F_CNT_ALIGNED | BGR_F_OVERLAP_FOUND | (pass ? BGR_F_RC : 0u);
            return flags;
        }
        if (tried == 0) {
            flags |= BGR_F_CNT_NOOVERLAP;
            n = 0;
            return flags;
        }
        flags |= BGR_F_OVERLAP_FOUND | BGR_F_RC;
    }
    flags |= BGR_F_CNT_NOTALIGNED;
    n = 0;
    return flags;
}

// ------------------------------------------------------------------------------ exhaustive extension
//
// The reference's recursion (alignerExhaustive.cpp:61-259) returns, at every level, the FIRST
// candidate in slot order whose subtree reaches the minimal total mismatch count: equivalently the
// first complete walk in depth-first order that attains the global minimum, provided it is within the
// budget. That is what the iterative branch-and-bound search below keeps (strict < on update;
// subtrees whose partial sum cannot beat the incumbent are skipped, which the reference also does
// level by level with `miss < miniMiss`). alignReadExhaustivePath (alignerPaths.cpp:66-88) calls the
// *Path helpers with empty vectors, so their "unitig already in path" filters never fire and the
// results are the same.
//
// Scratch layout (ints): [0, 2*D+2) final path | cur ids [D] | best ids [D] | frames 5*[D].
struct DfsLayout {
    u32 D, out0, cur0, best0, fr0;
    BGR_HD void init(u32 scratch_ints) {
        D = scratch_ints > 11 ? (scratch_ints - 2) / 9 : 1;
        out0 = 0;
        cur0 = 2 * D + 2;
        best0 = cur0 + D;
        fr0 = best0 + D;
    }
};

// Search one side. Returns true and (best_total, best_n ids at best0.., term) when a walk within
// `budget` exists. term: left side = start offset or -1 (ended exactly on a unitig start: nothing
// emitted); right side = end marker (|readLeft|+k-1, or 0 when the read ends on a unitig boundary).
// avoid_n > 0 (greedy "paths" aligner only): the first avoid_n ints of the scratch row are the path built
// so far, and a first-level candidate found in it is skipped (alignerPaths.cpp:136,307,370: the deeper
// levels of the reference test against vectors that are still empty, so only depth 0 filters).
BGR_HD bool exhaustive_side(const DevIndex& ix, const ReadView& v, bool right, u64 code0, i32 pos0, u32 budget,
                            Scratch& sc, const DfsLayout& lay, u32& best_total, u32& best_n, i32& term,
                            u32 avoid_n = 0) {
    const i64 km1 = ix.k - 1;
    best_total = budget + 1;
    best_n = 0;
    term = -1;
    i32 depth = 0;
    // frame: code (2 ints), pos, acc, next candidate
    auto fr_set = [&](u32 d, u64 code, i32 pos, u32 acc, u32 next) {
        u32 b = lay.fr0 + 5 * d;
        sc.set(b, (i32)(u32)code);
        sc.set(b + 1, (i32)(u32)(code >> 32));
        sc.set(b + 2, pos);
        sc.set(b + 3, (i32)acc);
        sc.set(b + 4, (i32)next);
    };
    fr_set(0, code0, pos0, 0, 0);
    while (depth >= 0) {
        const u32 fb = lay.fr0 + 5 * (u32)depth;
        const u64 code = (u64)(u32)sc.get(fb) | ((u64)(u32)sc.get(fb + 1) << 32);
        const i32 pos = sc.get(fb + 2);
        const u32 acc = (u32)sc.get(fb + 3);
        const u32 next = (u32)sc.get(fb + 4);
        // how much of the read is left on this side
        const i32 rl = right ? v.len - (pos + (i32)km1) : pos;
        if (rl == 0) {
            // the read is used up exactly at a unitig boundary (mapOnLeftEndExhaustive:112 returns 0
            // pushing nothing; mapOnRightEndExhaustive:64 pushes 0)
            if (acc < best_total) {
                best_total = acc;
                best_n = (u32)depth;
                term = right ? 0 : -1;
                for (u32 j = 0; j < best_n; ++j) sc.set(lay.best0 + j, sc.get(lay.cur0 + j));
            }
            --depth;
            continue;
        }
        Cands c = index_query(ix, code, !right);
        bool descended = false;
        for (u32 i = next; i < (u32)c.n; ++i) {
            if (acc >= best_total) break;  // nothing below can be strictly better
            Oriented u = orient(ix, c, (int)i);
            if (avoid_n && depth == 0) {
                bool seen = false;
                for (u32 j = 0; j < avoid_n; ++j) seen |= sc.get(j) == u.sid;
                if (seen) continue;
            }
            const i64 L = u.len - km1;
            const u32 cap = best_total - 1 - acc;  // a candidate is useful only if acc+miss < best_total
            if (L >= rl) {
                u32 miss = right ? mismatches(ix, v, u, km1, pos + (i32)km1, rl, cap)
                                 : mismatches(ix, v, u, L - rl, 0, rl, cap);
                if (miss <= cap) {
                    best_total = acc + miss;
                    best_n = (u32)depth + 1;
                    term = right ? rl + (i32)km1 : (i32)(L - rl);
                    for (u32 j = 0; j < (u32)depth; ++j) sc.set(lay.best0 + j, sc.get(lay.cur0 + j));
                    sc.set(lay.best0 + (u32)depth, u.sid);
                }
            } else {
                u32 miss = right ? mismatches(ix, v, u, km1, pos + (i32)km1, (i32)L, cap)
                                 : mismatches(ix, v, u, 0, rl - (i32)L, (i32)L, cap);
                if (miss <= cap) {
                    if ((u32)depth + 1 >= lay.D) {
                        // search stack too small: give up on this read at once, the host re-runs the
                        // batch with a larger scratch (carrying on would only enumerate dead ends)
                        sc.overflow = true;
                        return false;
                    }
                    sc.set(lay.cur0 + (u32)depth, u.sid);
                    sc.set(fb + 4, (i32)(i + 1));
                    u64 nc = right ? last_code(ix, u) : first_code(ix, u);
                    fr_set((u32)depth + 1, nc, right ? pos + (i32)L : pos - (i32)L, acc + miss, 0);
                    ++depth;
                    descended = true;
                    break;
                }
            }
        }
        if (!descended) --depth;
    }
    return best_total <= budget;
}

// alignReadExhaustive (alignerExhaustive.cpp:35-58) for read r; the final path is left in scratch
// [0, n) in output order.
BGR_HD u32 exhaustive_read(const DevIndex& ix, const BatchArgs& a, u32 r, Scratch& sc, u32& n, u32& overlaps) {
    ReadSetup s = setup_read(ix, a, r);
    ReadView& v = s.v;
    const u32 km1 = ix.k - 1;
    DfsLayout lay;
    lay.init(a.scratch_ints);
    n = 0;
    overlaps = (u32)s.nwin;  // overlaps += listOverlap.size(): every window is listed
    u32 flags = BGR_F_OVERLAP_FOUND;
    if (s.nwin == 0) return flags | BGR_F_CNT_NOTALIGNED;  // shorter than k-1: see DESIGN.md
    if ((u32)v.len == km1) {
        // single window, nothing on either side: checkBegin pushes 0, checkEnd pushes 0
        sc.set(0, 0);
        sc.set(1, 0);
        n = 2;
        return flags | BGR_F_MAPPED | BGR_F_CNT_ALIGNED;
    }
    // Every window is tried in order, but only windows whose (k-1)-mer is a border can succeed: a
    // window at pos>0 needs getEnd non-empty, window 0 needs getBegin non-empty (the read is longer
    // than the window here).
    SeedIter it;
    it.init(a.H1 + s.wb, (s.info & INFO_HIT1) ? s.nwin : 0, false, s.nwin);
    for (;;) {
        i32 pos = it.next();
        if (pos < 0) break;
        u64 code = seed_code(ix, v, pos);
        u32 eb = 0, nb = 0, ee = 0, ne = 0;
        i32 tb = 0, te = 0;
        u32 w = 0;
        if (pos == 0) {
            sc.set(w++, 0);  // checkBeginExhaustive:159
        } else {
            if (!exhaustive_side(ix, v, false, code, pos, a.errors, sc, lay, eb, nb, tb)) {
                if (sc.overflow) break;
                continue;
            }
            // left part in output order: [offset] then the ids from the far end to the anchor
            if (tb >= 0) sc.set(w++, tb);
            for (u32 j = 0; j < nb; ++j) sc.set(w++, sc.get(lay.best0 + nb - 1 - j));
        }
        if (!exhaustive_side(ix, v, true, code, pos, a.errors - eb, sc, lay, ee, ne, te)) {
            if (sc.overflow) break;
            continue;
        }
        for (u32 j = 0; j < ne; ++j) sc.set(w++, sc.get(lay.best0 + j));
        sc.set(w++, te);
        n = w;
        return flags | BGR_F_MAPPED | BGR_F_CNT_ALIGNED;
    }
    n = 0;
    return flags | BGR_F_CNT_NOTALIGNED;
}

// ---- the same search as a resumable state machine: one turn = one node of the depth-first search
// (index query -> candidate compares until one is worth descending into -> descend or return). This is
// what the persistent kernel runs: lanes of a warp are at different nodes of different searches but
// execute the same body every turn. exhaustive_read above is the loop-shaped statement of the same
// procedure; tests/emu runs both and compares.
enum { XP_SEED = 0, XP_LEFT = 1, XP_RIGHT = 2, XP_DONE = 3 };

struct ExhaustiveState {
    ReadSetup s;
    Scratch sc;
    SeedIter it;
    DfsLayout lay;
    u32 flags, phase, n;
    // current seed
    u64 seed_c;
    i32 seed_pos;
    u32 w, eb;  // ints of the final path written so far; mismatches spent on the left side
    // current search (one side)
    u32 budget, best_total, best_n;
    i32 term, depth;
    // top frame, kept in registers; the frames below it live in the scratch row
    u64 f_code;
    i32 f_pos;
    u32 f_acc, f_next;

    BGR_HD bool done() const { return phase == XP_DONE; }

    BGR_HD void start(const DevIndex& ix, const BatchArgs& a, u32 r, i32* scratch_row) {
        s = setup_read(ix, a, r);
        sc.base = scratch_row;
        sc.stride = a.sc_ent;
        sc.cap = a.scratch_ints;
        sc.overflow = false;
        lay.init(a.scratch_ints);
        flags = BGR_F_OVERLAP_FOUND;
        n = 0;
        w = eb = 0;
        phase = XP_SEED;
        seed_c = 0;
        seed_pos = 0;
        budget = best_total = best_n = 0;
        term = -1;
        depth = -1;
        f_code = 0;
        f_pos = 0;
        f_acc = f_next = 0;
        const u32 km1 = ix.k - 1;
        it.init(a.H1 + s.wb, (s.info & INFO_HIT1) ? s.nwin : 0, false, s.nwin);
        if (s.nwin == 0) {  // shorter than k-1: see DESIGN.md
            flags |= BGR_F_CNT_NOTALIGNED;
            phase = XP_DONE;
        } else if ((u32)s.v.len == km1) {  // single window, nothing on either side: [0, 0]
            sc.set(0, 0);
            sc.set(1, 0);
            n = 2;
            flags |= BGR_F_MAPPED | BGR_F_CNT_ALIGNED;
            phase = XP_DONE;
        }
    }

    BGR_HD void begin_side(bool right, u32 budget_) {
        phase = right ? XP_RIGHT : XP_LEFT;
        budget = budget_;
        best_total = budget_ + 1;
        best_n = 0;
        term = -1;
        depth = 0;
        f_code = seed_c;
        f_pos = seed_pos;
        f_acc = 0;
        f_next = 0;
    }

    // alignReadExhaustive's loop header (alignerExhaustive.cpp:41): the next window that can be a seed
    BGR_HD void pick_seed(const DevIndex& ix, const BatchArgs& a) {
        if (phase != XP_SEED) return;
        const i32 p = it.next();
        if (p < 0) {
            n = 0;
            flags |= BGR_F_CNT_NOTALIGNED;
            phase = XP_DONE;
            return;
        }
        seed_pos = p;
        seed_c = seed_code(ix, s.v, p);
        w = 0;
        eb = 0;
        if (p == 0) {
            sc.set(w++, 0);  // checkBeginExhaustive:159
            begin_side(true, a.errors);
        } else {
            begin_side(false, a.errors);
        }
    }

    BGR_HD void pop_frame() {
        --depth;
        if (depth >= 0) {
            const u32 b = lay.fr0 + 5 * (u32)depth;
            f_code = (u64)(u32)sc.get(b) | ((u64)(u32)sc.get(b + 1) << 32);
            f_pos = sc.get(b + 2);
            f_acc = (u32)sc.get(b + 3);
            f_next = (u32)sc.get(b + 4);
        }
    }

    // the side's search is over: left -> store its part of the path and search the right side;
    // right -> the read is mapped. A side without a walk within the budget fails the seed.
    BGR_HD void side_finished(const BatchArgs& a) {
        const bool ok = best_total <= budget;
        if (!ok || sc.overflow) {
            if (sc.overflow) {  // the host re-runs the batch with a larger scratch
                n = 0;
                flags |= BGR_F_CNT_NOTALIGNED;
                phase = XP_DONE;
            } else {
                phase = XP_SEED;
            }
            return;
        }
        if (phase == XP_LEFT) {
            // left part in output order: [offset] then the ids from the far end to the anchor
            if (term >= 0) sc.set(w++, term);
            for (u32 j = 0; j < best_n; ++j) sc.set(w++, sc.get(lay.best0 + best_n - 1 - j));
            eb = best_total;
            begin_side(true, a.errors - eb);
        } else {
            for (u32 j = 0; j < best_n; ++j) sc.set(w++, sc.get(lay.best0 + j));
            sc.set(w++, term);
            n = w;
            flags |= BGR_F_MAPPED | BGR_F_CNT_ALIGNED;
            phase = XP_DONE;
        }
    }

    // one node of the search (the body of exhaustive_side's while loop)
    BGR_HD void step_node(const DevIndex& ix, const BatchArgs& a) {
        const ReadView& v = s.v;
        const i64 km1 = ix.k - 1;
        const bool right = phase == XP_RIGHT;
        const i32 rl = right ? v.len - (f_pos + (i32)km1) : f_pos;
        if (rl == 0) {
            if (f_acc < best_total) {
                best_total = f_acc;
                best_n = (u32)depth;
                term = right ? 0 : -1;
                for (u32 j = 0; j < best_n; ++j) sc.set(lay.best0 + j, sc.get(lay.cur0 + j));
            }
            pop_frame();
        } else {
            const Cands c = index_query(ix, f_code, !right);
            bool descended = false;
            for (u32 i = f_next; i < (u32)c.n; ++i) {
                if (f_acc >= best_total) break;  // nothing below can be strictly better
                const Oriented u = orient(ix, c, (int)i);
                const i64 L = u.len - km1;
                const u32 cap = best_total - 1 - f_acc;
                if (L >= rl) {
                    const u32 miss = right ? mismatches(ix, v, u, km1, f_pos + (i32)km1, rl, cap)
                                           : mismatches(ix, v, u, L - rl, 0, rl, cap);
                    if (miss <= cap) {
                        best_total = f_acc + miss;
                        best_n = (u32)depth + 1;
                        term = right ? rl + (i32)km1 : (i32)(L - rl);
                        for (u32 j = 0; j < (u32)depth; ++j) sc.set(lay.best0 + j, sc.get(lay.cur0 + j));
                        sc.set(lay.best0 + (u32)depth, u.sid);
                    }
                } else {
                    const u32 miss = right ? mismatches(ix, v, u, km1, f_pos + (i32)km1, (i32)L, cap)
                                           : mismatches(ix, v, u, 0, rl - (i32)L, (i32)L, cap);
                    if (miss <= cap) {
                        if ((u32)depth + 1 >= lay.D) {  // search stack too small: give up, the host grows the scratch
                            sc.overflow = true;
                            depth = -1;
                            side_finished(a);
                            return;
                        }
                        // park this frame (resume with candidate i+1), make the child the top frame
                        const u32 b = lay.fr0 + 5 * (u32)depth;
                        sc.set(b, (i32)(u32)f_code);
                        sc.set(b + 1, (i32)(u32)(f_code >> 32));
                        sc.set(b + 2, f_pos);
                        sc.set(b + 3, (i32)f_acc);
                        sc.set(b + 4, (i32)(i + 1));
                        sc.set(lay.cur0 + (u32)depth, u.sid);
                        f_code = right ? last_code(ix, u) : first_code(ix, u);
                        f_pos = right ? f_pos + (i32)L : f_pos - (i32)L;
                        f_acc += miss;
                        f_next = 0;
                        ++depth;
                        descended = true;
                        break;
                    }
                }
            }
            if (!descended) pop_frame();
        }
        if (depth < 0) side_finished(a);
    }

    BGR_HD void step(const DevIndex& ix, const BatchArgs& a) {
        if (phase == XP_SEED) pick_seed(ix, a);
        if (phase == XP_LEFT || phase == XP_RIGHT) step_node(ix, a);
    }
};

// ------------------------------------------------------------------------------ greedy "paths" aligner
//
// alignReadGreedyPath and its helpers (alignerPaths.cpp:35-63 driver, :91-124 coverGreedyPath,
// :178-234 mapOnRightPath, :237-286 checkPairPaths; the two ends use the exhaustive path walkers,
// :289-411). Unreachable from the reference's command line (pathOption is never set) but part of
// class Aligner's public interface. ONE path vector collects everything in call order: the left walk
// of the start seed, the right walk of the END seed, then the cover between them, and it is not
// cleared between failed (start, end) attempts: reproduced as is. "Already in the path" compares a
// signed unitig id with every int of that vector, offsets included (std::find over vector<uNumber>).
//
// Scratch row: [0, P) the path | cur ids [D] | best ids [D] | search frames 5*[D].
struct PathLayout {
    u32 P;
    DfsLayout dfs;
    BGR_HD void init(u32 scratch_ints) {
        P = scratch_ints / 2;
        const u32 rest = scratch_ints - P;
        dfs.D = rest >= 14 ? rest / 7 : 1;
        dfs.out0 = 0;
        dfs.cur0 = P;
        dfs.best0 = dfs.cur0 + dfs.D;
        dfs.fr0 = dfs.best0 + dfs.D;
    }
};

struct PathBuf {
    Scratch* sc;
    u32 n, cap;
    BGR_HD void push(i32 x) {
        if (n < cap) sc->set(n, x);
        else sc->overflow = true;
        ++n;
    }
    BGR_HD bool contains(i32 x) const {
        bool seen = false;
        const u32 m = n < cap ? n : cap;
        for (u32 j = 0; j < m; ++j) seen |= sc->get(j) == x;
        return seen;
    }
};

// checkPairPaths (alignerPaths.cpp:237-286): the unitig that begins with the (k-1)-mer of window p1
// and ends with the one of window p2. Returns the mismatch count (allowed+1 = none) and the id in `number`.
BGR_HD u32 check_pair_paths(const DevIndex& ix, const ReadView& v, const PathBuf& path, i32 p1, i32 p2, i32& number,
                            u32 allowed) {
    const i32 km1 = (i32)ix.k - 1;
    const Cands c1 = index_query(ix, seed_code(ix, v, p1), false);  // getBegin(overlap1)
    const Cands c2 = index_query(ix, seed_code(ix, v, p2), true);   // getEnd(overlap2)
    if (p2 - p1 < (i32)ix.k) {
        // the two windows overlap: any common unitig is taken without looking at the read
        for (int i = 0; i < c1.n; ++i) {
            const i32 id1 = orient(ix, c1, i).sid;
            for (int j = 0; j < c2.n; ++j)
                if (orient(ix, c2, j).sid == id1) {
                    if (!path.contains(id1)) number = id1;
                    return 0;
                }
        }
        return allowed + 1;
    }
    const i32 sub_len = p2 - p1 - km1;
    u32 best = allowed + 1;
    i32 best_sid = 0;
    for (int i = 0; i < c1.n; ++i) {
        const Oriented u = orient(ix, c1, i);
        for (int j = 0; j < c2.n; ++j) {
            if (orient(ix, c2, j).sid == u.sid && !path.contains(u.sid)) {
                if (u.len - 2 * (i64)km1 == (i64)sub_len) {
                    const u32 miss = mismatches(ix, v, u, km1, p1 + km1, sub_len, allowed);
                    if (miss < best) {
                        best = miss;
                        best_sid = u.sid;
                    }
                }
            }
        }
    }
    if (best <= allowed) number = best_sid;
    return best;
}

// mapOnRightPath (alignerPaths.cpp:178-234), tail recursion as a loop. Walks right from window
// `start` until the read ends (ended) or the walk lands exactly on a later listed window (first = its
// index). Returns false when no walk within `budget` exists; second = mismatches spent.
BGR_HD bool map_on_right_path(const DevIndex& ix, const ReadView& v, PathBuf& path, i32 start, i32 nwin, u32 budget,
                              bool& ended, i32& first, u32& second) {
    const i32 km1 = (i32)ix.k - 1;
    u64 code = seed_code(ix, v, start);
    i32 pos = start;
    u32 total = 0;
    for (;;) {
        const i32 rl = v.len - (pos + km1);
        if (rl <= 0) {  // the reference prints "should not appears" and exits; unreachable (start < end)
            first = start;
            second = total;
            return true;
        }
        const Cands c = index_query(ix, code, false);
        u32 best = budget + 1;
        i32 best_sid = 0, next = start;
        bool have_best = false;
        u64 next_code = 0;
        i64 next_L = 0;
        for (int i = 0; i < c.n; ++i) {
            const Oriented u = orient(ix, c, i);
            if (path.contains(u.sid)) continue;
            const i64 L = u.len - km1;
            if ((i64)rl <= L) {
                const u32 miss = mismatches(ix, v, u, km1, pos + km1, rl, budget);
                if (miss < best) {
                    ended = true;
                    best = miss;
                    best_sid = u.sid;
                    have_best = true;
                }
            } else {
                const u32 miss = mismatches(ix, v, u, km1, pos + km1, (i32)L, budget);
                if (miss < best) {
                    const u64 oc = last_code(ix, u);
                    ended = false;
                    best = miss;
                    best_sid = u.sid;
                    have_best = true;
                    next_code = oc;
                    next_L = L;
                    next = start;
                    // a later listed overlap at exactly that position with that code (list[j].pos == j)
                    const i64 j = (i64)pos + L;
                    if (j >= (i64)start + 1 && j < (i64)nwin && seed_code(ix, v, (i32)j) == oc) next = (i32)j;
                }
            }
        }
        if (ended) {
            if (have_best) path.push(best_sid);
            first = start;
            second = total + best;
            return true;
        }
        if (best > budget) return false;
        path.push(best_sid);
        total += best;
        if (next > start) {
            first = next;
            second = total;
            return true;
        }
        budget -= best;
        code = next_code;
        pos += (i32)next_L;
    }
}

// coverGreedyPath (alignerPaths.cpp:91-124), recursion as a loop: joins window `start` to window
// `end` by single unitigs between listed windows (checkPairPaths) or, failing that, by a rightward walk.
BGR_HD bool cover_greedy_path(const DevIndex& ix, const ReadView& v, PathBuf& path, i32 start, i32 end, i32 nwin,
                              u32 budget, u32 effort) {
    bool ended = false;
    for (;;) {
        if (start == end) return true;
        i32 indice = 0, keep = 0;
        u32 best = budget + 1;
        for (u32 i = 1; (i64)i + start <= (i64)end && i <= effort; ++i) {
            i32 number = 0;
            const u32 miss = check_pair_paths(ix, v, path, start, start + (i32)i, number, budget);
            if (miss < best) {
                indice = (i32)i;
                keep = number;
                best = miss;
            }
        }
        if (best <= budget) {
            path.push(keep);
            budget -= best;
            start += indice;
            continue;
        }
        i32 first = start;
        u32 second = 0;
        if (!map_on_right_path(ix, v, path, start, nwin, budget, ended, first, second)) return false;
        if (second > budget) return false;
        if (ended) return true;
        if (first <= start) return false;
        budget -= second;
        start = first;
    }
}

// alignReadGreedyPath (alignerPaths.cpp:35-63) for read r; the path is left in scratch [0, n).
BGR_HD u32 greedy_path_read(const DevIndex& ix, const BatchArgs& a, u32 r, Scratch& sc, u32& n) {
    ReadSetup s = setup_read(ix, a, r);
    ReadView& v = s.v;
    PathLayout lay;
    lay.init(a.scratch_ints);
    n = 0;
    const u32 flags = BGR_F_OVERLAP_FOUND;
    if (s.nwin == 0) return flags | BGR_F_CNT_NOTALIGNED;  // shorter than k-1: see DESIGN.md
    const i32 nwin = s.nwin;  // listOverlap holds every window: list[j].second == j
    const u32 last_start = (a.effort - 1u) < (u32)(nwin - 1) ? (a.effort - 1u) : (u32)(nwin - 1);
    PathBuf path;
    path.sc = &sc;
    path.cap = lay.P;
    for (u32 start = 0; start <= last_start; ++start) {
        path.n = 0;
        u32 eb = 0, nb = 0;
        i32 tb = 0;
        if (start == 0) {
            path.push(0);  // checkBeginExhaustivePath:352
        } else {
            if (!exhaustive_side(ix, v, false, seed_code(ix, v, (i32)start), (i32)start, a.errors, sc, lay.dfs, eb, nb,
                                 tb)) {
                if (sc.overflow) return flags | BGR_F_CNT_NOTALIGNED;
                continue;
            }
            if (tb >= 0) path.push(tb);
            for (u32 j = 0; j < nb; ++j) path.push(sc.get(lay.dfs.best0 + nb - 1 - j));
        }
        const i32 lo = (i32)start > nwin - (i32)a.effort ? (i32)start : nwin - (i32)a.effort;
        for (i32 end = nwin - 1; end >= lo; --end) {
            u32 ee = 0, ne = 0;
            i32 te = 0;
            const u32 avoid = path.n < path.cap ? path.n : path.cap;  // >= 1: the left part is never empty
            if (!exhaustive_side(ix, v, true, seed_code(ix, v, end), end, a.errors - eb, sc, lay.dfs, ee, ne, te,
                                 avoid)) {
                if (sc.overflow) return flags | BGR_F_CNT_NOTALIGNED;
                continue;
            }
            for (u32 j = 0; j < ne; ++j) path.push(sc.get(lay.dfs.best0 + j));
            path.push(te);
            if (cover_greedy_path(ix, v, path, (i32)start, end, nwin, a.errors - ee - eb, a.effort)) {
                n = path.n;
                return flags | BGR_F_MAPPED | BGR_F_CNT_ALIGNED;
            }
            if (sc.overflow) return flags | BGR_F_CNT_NOTALIGNED;
        }
    }
    n = 0;
    return flags | BGR_F_CNT_NOTALIGNED;
}

// ------------------------------------------------------------------------------ correction mode
//
// recoverPath (aligner.cpp:268-290) on packed words. path = [offset, +-u1, ..., +-um] as alignReadGreedy
// returns it. S = U1 ++ U2[k-1:] ++ ... ++ Um[k-1:] (oriented unitigs; compactionEnd, utils.cpp:171-179,
// always finds the forward overlap on such a path); the corrected read is S[offset, offset+len), or its
// reverse complement when `rc` -- produced directly by walking the path backwards with every unitig
// flipped. Writes ceil(len/32) words (first base most significant); returns the number of bases written
// (len, unless the path spells less than the read: the caller reports that).
struct BaseWriter {
    u64* out;
    u64 acc;
    u32 fill, wi;
    BGR_HD void push(u64 x, u32 m) {  // m (1..32) bases, left aligned, zero behind
        if (fill == 0) {
            if (m == 32) out[wi++] = x;
            else {
                acc = x;
                fill = m;
            }
        } else {
            acc |= x >> (2 * fill);
            if (fill + m >= 32) {
                out[wi++] = acc;
                const u32 used = 32 - fill;
                acc = (m > used) ? x << (2 * used) : 0ull;
                fill = m - used;
            } else {
                fill += m;
            }
        }
    }
    BGR_HD void flush() {
        if (fill) out[wi++] = acc;
        fill = 0;
    }
};

// bases [from, from+take) of unitig (start, ulen) read in orientation `neg` (true = reverse complement)
BGR_HD void emit_unitig(const DevIndex& ix, u64 start, u32 ulen, bool neg, u32 from, u32 take, BaseWriter& w) {
    for (u32 done = 0; done < take; done += 32) {
        const u32 m = (take - done < 32) ? take - done : 32;
        const u32 p = from + done;
        u64 x;
        if (!neg) {
            x = ld_bases(ix.pool, (i64)(start + p));
        } else {
            const u64 y = ld_bases(ix.pool, (i64)(start + ulen - p - m));  // its first m bases, reversed, are wanted
            x = rc64(y) << ((2 * (32 - m)) & 63);
        }
        if (m < 32) x &= ~(~0ull >> (2 * m));
        w.push(x, m);
    }
}

BGR_HD u32 correct_read(const DevIndex& ix, const i32* path, u32 n, u32 len, bool rc, u64* out) {
    if (n < 2) return 0;  // the reference reads numbers[1] of a one-element path here
    const u32 km1 = ix.k - 1;
    u64 S = 0;
    for (u32 j = 1; j < n; ++j) {
        const i32 id = path[j];
        const u32 ul = ix.meta[id < 0 ? -id : id].len;
        S += (j == 1) ? ul : ul - km1;
    }
    const i64 offset = path[0];
    if (offset < 0 || (u64)offset > S) return 0;
    const u64 room = S - (u64)offset;
    const u32 emit = room < len ? (u32)room : len;
    BaseWriter w;
    w.out = out;
    w.acc = 0;
    w.fill = 0;
    w.wi = 0;
    u64 skip = rc ? S - (u64)offset - emit : (u64)offset;
    u32 need = emit;
    for (u32 t = 1; t < n && need; ++t) {
        const u32 j = rc ? n - t : t;  // backwards for the reverse complement
        const i32 id = path[j];
        const UnitigMeta m = ix.meta[id < 0 ? -id : id];
        const bool neg = (id < 0) != rc;
        const u32 base0 = (t == 1) ? 0 : km1;
        const u32 avail = m.len - base0;
        if (skip >= avail) {
            skip -= avail;
            continue;
        }
        const u32 from = base0 + (u32)skip;
        const u32 take = (avail - (u32)skip < need) ? avail - (u32)skip : need;
        emit_unitig(ix, m.start, m.len, neg, from, take, w);
        skip = 0;
        need -= take;
    }
    w.flush();
    return emit - need;
}

#if defined(__CUDACC__)
// ---- extension kernels. A finished read leaves its path (final order) in its scratch row, its flags
// and path length in res[r].flags_len and the scratch row number in res[r].path_off; the compaction
// kernels below then build the pool in read order (no atomics, deterministic layout).

__device__ __forceinline__ void count_classes(const BatchArgs& a, u32 n_al, u32 n_na, u32 n_no) {
    // per-thread tallies -> one atomic per warp and class
    const u32 lane = threadIdx.x & 31;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        n_al += __shfl_xor_sync(0xFFFFFFFFu, n_al, d);
        n_na += __shfl_xor_sync(0xFFFFFFFFu, n_na, d);
        n_no += __shfl_xor_sync(0xFFFFFFFFu, n_no, d);
    }
    if (lane == 0) {
        if (n_al) atomicAdd(&a.cnt->c[BGR_C_ALIGNED], (unsigned long long)n_al);
        if (n_na) atomicAdd(&a.cnt->c[BGR_C_NOTALIGNED], (unsigned long long)n_na);
        if (n_no) atomicAdd(&a.cnt->c[BGR_C_NOOVERLAP], (unsigned long long)n_no);
    }
}

static const int EXT_THREADS = 128;

// Persistent warps: every lane walks one read with GreedyState::step; lanes that finish hand in their
// result and take the next read of the hit list (a warp draws 32 list entries per atomic), so the
// warp stays full until the list runs dry.
template <int MINB>
__global__ void __launch_bounds__(EXT_THREADS, MINB) k_extend_greedy(DevIndex ix, BatchArgs a) {
    const u32 lane = threadIdx.x & 31;
    const u32 list_n = a.scnt->list_n;
    GreedyState g;
    g.phase = GP_DONE;
    g.sc.overflow = false;
    bool have = false;  // this lane holds a read
    u32 r = 0, t = 0;
    u32 chunk_base = 0, chunk_left = 0;  // warp-uniform: list entries drawn but not handed out yet
    bool dry = false;                    // the list has no more chunks
    u32 n_al = 0, n_na = 0, n_no = 0;
    bool any_overflow = false;
    // Hand-in, refill and seed selection are needed once or twice per read while a walk has several
    // levels: run per lane as soon as it is due, every turn pays for them with a handful of lanes
    // active (profiles/r01_ncu_k_extend_greedy_v5_5Mbp.txt: 8 of 32 threads per instruction). They are
    // therefore serviced together, once `lazy` lanes wait for them (or nobody can walk).
    const u32 lazy = a.ext_lazy ? a.ext_lazy : 1;
    for (;;) {
        const bool walking = have && g.phase != GP_DONE && g.phase != GP_SEED;
        const u32 waiting = __ballot_sync(0xFFFFFFFFu, !walking);
        if ((u32)__popc(waiting) >= lazy || waiting == 0xFFFFFFFFu) {
            if (have && g.done()) {  // hand in
                const u32 plen = g.finish_deferred();
                any_overflow |= g.sc.overflow;
                bgr_result res;
                res.flags_len = (g.flags & 0xFFu) | (plen ? RES_REV_HEAD : 0u) | (plen << 8);
                res.path_off = t;
                a.res[r] = res;
                n_al += (g.flags & BGR_F_CNT_ALIGNED) != 0;
                n_na += (g.flags & BGR_F_CNT_NOTALIGNED) != 0;
                n_no += (g.flags & BGR_F_CNT_NOOVERLAP) != 0;
                have = false;
            }
            const u32 want = __ballot_sync(0xFFFFFFFFu, !have);
            if (want) {
                if (chunk_left == 0 && !dry) {
                    u32 base = 0;
                    if (lane == 0) base = atomicAdd(&a.cnt->ext_cursor, 32u);
                    base = __shfl_sync(0xFFFFFFFFu, base, 0);
                    if (base >= list_n) dry = true;
                    else {
                        chunk_base = base;
                        chunk_left = min(32u, list_n - base);
                    }
                }
                const u32 rank = __popc(want & ((1u << lane) - 1));
                if (!have && rank < chunk_left) {
                    t = chunk_base + rank;
                    r = a.list[t];
                    g.start(ix, a, r, a.scratch + (size_t)t * a.sc_row);
                    have = true;
                }
                const u32 taken = min((u32)__popc(want), chunk_left);
                chunk_base += taken;
                chunk_left -= taken;
                if (dry && want == 0xFFFFFFFFu) break;  // nobody holds a read and nothing is left
            }
            if (have && g.phase == GP_SEED) g.pick_seed(ix, a);
        }
        if (have && g.phase != GP_DONE && g.phase != GP_SEED) g.step_level(ix, a);
    }
    if (any_overflow) atomicExch(&a.cnt->scratch_overflow, 1u);
    count_classes(a, n_al, n_na, n_no);
}

// alignReadExhaustive / alignReadExhaustivePath as persistent warps (same life cycle as k_extend_greedy):
// every lane runs one node of its read's search per turn.
template <int MINB>
__global__ void __launch_bounds__(EXT_THREADS, MINB) k_extend_exhaustive_nodes(DevIndex ix, BatchArgs a) {
    const u32 lane = threadIdx.x & 31;
    const u32 list_n = a.scnt->list_n;
    ExhaustiveState g;
    g.phase = XP_DONE;
    g.sc.overflow = false;
    bool have = false;
    u32 r = 0, t = 0;
    u32 chunk_base = 0, chunk_left = 0;
    bool dry = false;
    u32 n_al = 0, n_na = 0;
    bool any_overflow = false;
    const u32 lazy = a.ext_lazy ? a.ext_lazy : 1;
    for (;;) {
        const bool searching = have && (g.phase == XP_LEFT || g.phase == XP_RIGHT);
        const u32 waiting = __ballot_sync(0xFFFFFFFFu, !searching);
        if ((u32)__popc(waiting) >= lazy || waiting == 0xFFFFFFFFu) {
            if (have && g.done()) {  // hand in
                any_overflow |= g.sc.overflow;
                const u32 plen = ((g.flags & BGR_F_MAPPED) && !g.sc.overflow) ? g.n : 0;
                bgr_result res;
                res.flags_len = (g.flags & 0xFFu) | (plen << 8);
                res.path_off = t;
                a.res[r] = res;
                n_al += (g.flags & BGR_F_CNT_ALIGNED) != 0;
                n_na += (g.flags & BGR_F_CNT_NOTALIGNED) != 0;
                have = false;
            }
            const u32 want = __ballot_sync(0xFFFFFFFFu, !have);
            if (want) {
                if (chunk_left == 0 && !dry) {
                    u32 base = 0;
                    if (lane == 0) base = atomicAdd(&a.cnt->ext_cursor, 32u);
                    base = __shfl_sync(0xFFFFFFFFu, base, 0);
                    if (base >= list_n) dry = true;
                    else {
                        chunk_base = base;
                        chunk_left = min(32u, list_n - base);
                    }
                }
                const u32 rank = __popc(want & ((1u << lane) - 1));
                if (!have && rank < chunk_left) {
                    t = chunk_base + rank;
                    r = a.list[t];
                    g.start(ix, a, r, a.scratch + (size_t)t * a.sc_row);
                    have = true;
                }
                const u32 taken = min((u32)__popc(want), chunk_left);
                chunk_base += taken;
                chunk_left -= taken;
                if (dry && want == 0xFFFFFFFFu) break;
            }
            if (have && g.phase == XP_SEED) g.pick_seed(ix, a);
        }
        if (have && (g.phase == XP_LEFT || g.phase == XP_RIGHT)) g.step_node(ix, a);
    }
    if (any_overflow) atomicExch(&a.cnt->scratch_overflow, 1u);
    count_classes(a, n_al, n_na, 0);
}

__global__ void __launch_bounds__(EXT_THREADS) k_extend_exhaustive(DevIndex ix, BatchArgs a) {
    const u32 t = blockIdx.x * EXT_THREADS + threadIdx.x;
    const bool valid = t < a.scnt->list_n;
    const u32 r = valid ? a.list[t] : 0;
    Scratch sc;
    sc.base = a.scratch + (size_t)(valid ? t : 0) * a.sc_row;
    sc.stride = a.sc_ent;
    sc.cap = a.scratch_ints;
    sc.overflow = false;
    u32 flags = 0, n = 0, overlaps = 0;
    if (valid) {
        flags = (a.mode == BGR_MODE_GREEDY_PATH) ? greedy_path_read(ix, a, r, sc, n)
                                                 : exhaustive_read(ix, a, r, sc, n, overlaps);
        u32 plen = (flags & BGR_F_MAPPED) && !sc.overflow ? n : 0;
        if (sc.overflow) atomicExch(&a.cnt->scratch_overflow, 1u);
        bgr_result res;
        res.flags_len = (flags & 0xFFu) | (plen << 8);
        res.path_off = t;
        a.res[r] = res;
    }
    count_classes(a, valid && (flags & BGR_F_CNT_ALIGNED), valid && (flags & BGR_F_CNT_NOTALIGNED),
                  valid && (flags & BGR_F_CNT_NOOVERLAP));
}

// ---- path compaction: exclusive scan of the path lengths in read order, then gather
static const int CP_THREADS = 256, CP_PER_THREAD = 4, CP_BLOCK = CP_THREADS * CP_PER_THREAD;

__device__ __forceinline__ u32 block_exclusive_scan(u32 v, u32* total) {
    // 256 threads; returns the exclusive prefix of v, *total = block sum
    __shared__ u32 warp_sums[CP_THREADS / 32];
    const u32 lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    u32 incl = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        u32 x = __shfl_up_sync(0xFFFFFFFFu, incl, d);
        if (lane >= (u32)d) incl += x;
    }
    if (lane == 31) warp_sums[warp] = incl;
    __syncthreads();
    u32 base = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < CP_THREADS / 32; ++w) {
        u32 sw = warp_sums[w];
        if ((u32)w < warp) base += sw;
        tot += sw;
    }
    __syncthreads();
    *total = tot;
    return base + incl - v;
}

__global__ void __launch_bounds__(CP_THREADS) k_path_block_sums(BatchArgs a, unsigned long long* bsum) {
    const u32 r0 = blockIdx.x * CP_BLOCK + threadIdx.x * CP_PER_THREAD;
    u32 v = 0;
#pragma unroll
    for (int j = 0; j < CP_PER_THREAD; ++j)
        if (r0 + j < a.n) v += a.res[r0 + j].flags_len >> 8;
    u32 total;
    block_exclusive_scan(v, &total);
    if (threadIdx.x == 0) bsum[blockIdx.x] = total;
}

// single block: exclusive scan of the block sums in place; total -> pool_cursor
__global__ void __launch_bounds__(CP_THREADS) k_path_scan_blocks(BatchArgs a, unsigned long long* bsum, u32 nb) {
    __shared__ unsigned long long carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (u32 base = 0; base < nb; base += CP_THREADS) {
        const u32 i = base + threadIdx.x;
        const u32 v = i < nb ? (u32)bsum[i] : 0;  // a block sum is < 2^32 (1024 reads)
        u32 total;
        const u32 ex = block_exclusive_scan(v, &total);
        const unsigned long long c = carry;
        if (i < nb) bsum[i] = c + ex;
        __syncthreads();
        if (threadIdx.x == 0) carry = c + total;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        a.cnt->pool_cursor = carry;
        if (carry > a.pool_cap) a.cnt->pool_overflow = 1;
    }
}

__global__ void __launch_bounds__(CP_THREADS) k_path_gather(BatchArgs a, const unsigned long long* bsum) {
    if (a.cnt->pool_overflow) return;  // the host grows the pool and runs the compaction again
    const u32 r0 = blockIdx.x * CP_BLOCK + threadIdx.x * CP_PER_THREAD;
    u32 len[CP_PER_THREAD], v = 0;
    bgr_result rs[CP_PER_THREAD];
#pragma unroll
    for (int j = 0; j < CP_PER_THREAD; ++j) {
        rs[j].flags_len = 0;
        rs[j].path_off = 0;
        if (r0 + j < a.n) rs[j] = a.res[r0 + j];
        len[j] = rs[j].flags_len >> 8;
        v += len[j];
    }
    u32 total;
    unsigned long long off = bsum[blockIdx.x] + block_exclusive_scan(v, &total);
#pragma unroll
    for (int j = 0; j < CP_PER_THREAD; ++j) {
        const u32 r = r0 + j;
        if (len[j]) {
            bgr_result res = rs[j];
            const i32* src = a.scratch + (size_t)res.path_off * a.sc_row;  // scratch row
            i32* dst = a.pool + off;
            // pathBegin was pushed walking away from the seed: reverse(pathBegin) (alignerGreedy.cpp:48)
            const u32 nl = (res.flags_len & RES_REV_HEAD) ? (u32)src[(size_t)(a.scratch_ints - 1) * a.sc_ent] : 0u;
            for (u32 q = 0; q < len[j]; ++q) dst[q] = src[(size_t)(q < nl ? nl - 1 - q : q) * a.sc_ent];
            res.flags_len &= ~RES_REV_HEAD;
            res.path_off = (u32)off;
            a.res[r] = res;
            off += len[j];
        } else if (rs[j].path_off != 0) {
            // an unmapped read that went through the extension kernel still carries its scratch row
            // number (which depends on the order of the hit list): results are deterministic without it
            a.res[r].path_off = 0;
        }
    }
}

// ---- correction mode (-c): the sequence a greedy path spells, cut to the read (recoverPath,
// aligner.cpp:268-290: oriented unitigs glued on their k-1 overlaps, substr(offset, |read|)), reverse
// complemented when the read was mapped through its reverse complement (alignerGreedy.cpp:396)
__global__ void __launch_bounds__(128) k_correct(DevIndex ix, BatchArgs a) {
    const u32 r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= a.n) return;
    const bgr_result res = a.res[r];
    if (!(res.flags_len & BGR_F_MAPPED)) return;
    const u64 o = a.off[r];
    const u32 len = (u32)(a.off[r + 1] - o);
    const u32 got = correct_read(ix, a.pool + res.path_off, res.flags_len >> 8, len, (res.flags_len & BGR_F_RC) != 0,
                                 a.corr + corrected_word_base(o, r));
    if (got != len) atomicExch(&a.cnt->correct_error, 1u);
}

// getBegin / getEnd for one code, for tests (bgr_index_query)
__global__ void k_index_query(DevIndex ix, u64 code, int want_end, i32* out) {
    Cands c = index_query(ix, code, want_end != 0);
    out[0] = c.n;
    for (int i = 0; i < c.n; ++i) out[1 + i] = orient(ix, c, i).sid;
}
#endif  // __CUDACC__

}  // namespace bgr
