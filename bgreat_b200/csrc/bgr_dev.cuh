// Device-side vocabulary shared by every kernel of the mapping path: the index view, the per-batch argument
// block, counters, and the small helpers (bucket loads, base extraction, exact index lookup).
// Part of bgr_kernels.cuh, which is the header to include.
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
    const Entry* entries;
    const UnitigMeta* meta;
    const u64* pool;
    u64 n_buckets;
    const u32* bloom;
    u32 n_bloom;
    u32 bloom_mode;  // BLOOM_* (bgr_common.h)
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

}  // namespace bgr
