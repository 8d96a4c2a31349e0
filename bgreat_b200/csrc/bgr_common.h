// Shared host/device definitions: 2-bit encoding, the border-index layout and its hash.
//
// Data layout (everything lives in ONE flat blob so that it can be exported / broadcast /
// imported with a single copy, see bgr_index_export_device):
//
//   BlobHeader
//   keys    : u64[n_buckets * 4]      bucketed open addressing, one bucket = 4 keys = one 32-byte
//                                     DRAM sector; EMPTY_KEY marks a free slot; a key lives in the
//                                     first bucket at or after hash(key) that had room
//   entries : Entry[n_buckets * 4]    per key slot, one 128-byte line: left[4] + right[4] unitig slots, the
//                                     reference's two unitigIndices tables merged (reference aligner.h:49-55,
//                                     aligner.cpp:463-533). A slot is 16 bytes and carries everything a walk
//                                     needs to go on without another dependent load: where the unitig lies in
//                                     `pool` and how long it is, its id and orientation flags, and the key slot
//                                     of the (k-1)-mer at its FAR end (the next level's entry). One side of an
//                                     entry = 64 bytes = two sectors = one request.
//   bloom   : u32[n_bloom]            negative filter in front of `keys` for the seed scan: a blocked Bloom
//                                     filter whose block is ONE 32-bit word; every key sets 3 bits of its
//                                     word. "Some bit clear" proves absence; a window that passes is
//                                     verified against `keys` in the same kernel, so the hit masks the scan
//                                     emits are exact. Sized by the build (bloom_words_for): small indexes
//                                     get a filter that fits shared memory (<= 128 KB, one hash), larger
//                                     ones 2..16 bytes per key kept L2-resident as far as possible (two
//                                     hashes).
//   meta    : UnitigMeta[n_unitigs+1] first base of the unitig in `pool`, and its length; ids are
//                                     1-based like the reference's (aligner.cpp:408)
//   pool    : u64[pool_words]         all unitigs, 2 bits per base, first base in the most
//                                     significant bits (same digit order as str2num, utils.cpp:117)
#pragma once
#include <stdint.h>
#include <stdlib.h>

#if defined(__CUDACC__)
#define BGR_HD __host__ __device__ __forceinline__
#else
#define BGR_HD inline
#endif

namespace bgr {

typedef uint64_t u64;
typedef uint32_t u32;
typedef int64_t i64;
typedef int32_t i32;

static const u64 EMPTY_KEY = ~0ull;
static const u32 SLOT_ID_MASK = 0x3FFFFFFFu;  // unitig ids < 2^30
// Orientation flags of a slot, decided at build time by the reference's "forward test"
// (aligner.cpp:174,235): bit 30 answers the query made with the canonical key itself, bit 31 the
// query made with its reverse complement.
//   left  table: bit30 = unitig begins with key        (getBegin(key))
//                bit31 = rc(unitig end) == key          (getEnd(rc(key)))
//   right table: bit30 = unitig ends with key           (getEnd(key))
//                bit31 = rc(unitig begin) == key        (getBegin(rc(key)))
static const u32 SLOT_FWD_CANON = 1u << 30;
static const u32 SLOT_FWD_RC = 1u << 31;

// One candidate of a border: 16 bytes.
//   start_len : first base of the (forward) unitig in `pool`, bits 0..35; its length in bases, bits 36..63
//   id_flags  : unitig id (1-based, < 2^30) | SLOT_FWD_CANON | SLOT_FWD_RC; 0 = empty slot
//   next      : key slot of the canonical (k-1)-mer at the unitig's other end, bits 0..29 -- the entry the
//               walk queries after stepping through this unitig, whichever way it was entered (entered through
//               the canonical key the next code is some c, entered through its reverse complement it is rc(c):
//               both have the same canonical key). SLOT_NEXT_CANON: c <= rc(c); SLOT_NEXT_PAL: c == rc(c).
struct USlot {
    u64 start_len;
    u32 id_flags;
    u32 next;
};
static const u32 SLOT_START_BITS = 36, SLOT_LEN_BITS = 28;
static const u64 SLOT_START_MASK = (1ull << SLOT_START_BITS) - 1;
static const u32 SLOT_NEXT_MASK = 0x3FFFFFFFu;
static const u32 SLOT_NEXT_CANON = 1u << 30;
static const u32 SLOT_NEXT_PAL = 1u << 31;

struct Entry {
    USlot left[4];
    USlot right[4];
};

struct UnitigMeta {
    u64 start;  // base offset in pool
    u32 len;    // bases
    u32 pad;
};

struct BlobHeader {
    u64 magic;
    u32 version;
    u32 k;
    u64 n_unitigs;
    u64 n_keys;
    u64 n_buckets;
    u64 pool_words;
    u64 off_keys, off_entries, off_meta, off_pool;  // byte offsets from the blob start
    u64 total_bytes;
    u64 n_bloom;  // words of the Bloom filter
    u64 off_bloom;
    u32 bloom_mode;  // BLOOM_SMEM_1HASH / BLOOM_WORDS_2HASH / BLOOM_MINIMIZER
    u32 pad;
};
static const u64 BLOB_MAGIC = 0x3030324252474221ull;  // "!BGRB200"
static const u32 BLOB_VERSION = 8;  // 6: minimizer filter's bit pattern from mm_pat_hash; 7: so the single-hash filter of 31-mers; 8: four bits per key in a minimizer block

// 2-bit code of a base character, A0 C1 G2 T3 (utils.cpp:117-129). Returns 4 for anything else.
BGR_HD u32 base_code(unsigned char c) {
    switch (c) {
        case 'A': return 0;
        case 'C': return 1;
        case 'G': return 2;
        case 'T': return 3;
    }
    return 4;
}

// reverse complement of all 32 bases of a word (complement = 3 - code = ~code)
BGR_HD u64 rc64(u64 x) {
    x = ~x;
#if defined(__CUDA_ARCH__)
    x = __brevll(x);
#else
    x = ((x >> 1) & 0x5555555555555555ull) | ((x & 0x5555555555555555ull) << 1);
    x = ((x >> 2) & 0x3333333333333333ull) | ((x & 0x3333333333333333ull) << 2);
    x = ((x >> 4) & 0x0F0F0F0F0F0F0F0Full) | ((x & 0x0F0F0F0F0F0F0F0Full) << 4);
    x = __builtin_bswap64(x);
#endif
    // the bit reversal also swapped the two bits of every base: swap them back
    return ((x >> 1) & 0x5555555555555555ull) | ((x & 0x5555555555555555ull) << 1);
}

// reverse the order of the 32 two-bit groups (no complement); used for N masks
BGR_HD u64 rev2(u64 x) { return rc64(~x); }

// rcb(x, n) of the reference (utils.cpp:182-192): reverse complement of an n-base code
BGR_HD u64 rcb(u64 x, u32 n) { return rc64(x) >> (64 - 2 * n); }

// n (1..32) bases of the 2-bit stream W starting at base p, right aligned (str2num's value). W[p/32 + 1] must
// be readable.
BGR_HD u64 bases_at(const u64* W, u64 p, u32 n) {
    const u64 w = p >> 5;
    const u32 s = (u32)(p & 31) * 2;
    const u64 x = (W[w] << s) | ((W[w + 1] >> 1) >> (63 - s));
    return x >> (64 - 2 * n);
}

// The code a walk continues with after stepping through the unitig of a slot that it entered through the
// CANONICAL key (entered through the key's reverse complement it is the reverse complement of this):
//   left table  = getBegin(key): the unitig oriented to begin with key -> its last k-1 bases
//   right table = getEnd(key):   the unitig oriented to end with key   -> its first k-1 bases
// (aligner.cpp:147-267 orient by the forward test only; alignerGreedy.cpp:200,246 take the other end).
BGR_HD u64 slot_far_code(const u64* pool, u64 start, u64 len, u32 km1, u32 id_flags, bool right_table) {
    const bool rc = (id_flags & SLOT_FWD_CANON) == 0;
    const u64 ubeg = bases_at(pool, start, km1), uend = bases_at(pool, start + len - km1, km1);
    if (!right_table) return rc ? rcb(ubeg, km1) : uend;
    return rc ? rcb(uend, km1) : ubeg;
}

// hash of a canonical (k-1)-mer -> bucket. Kept to 32-bit integer ops on purpose: the seed-scan
// kernel is instruction-issue bound (profiles/r01_ncu_k_pack_scan_v1.txt) and computes one per window.
// n_buckets < 2^32.
BGR_HD u32 key_hash(u64 key) {
    u32 lo = (u32)key, hi = (u32)(key >> 32);
    u32 h = lo * 0x9E3779B1u + hi * 0x85EBCA77u;
    h ^= h >> 15;
    h *= 0x2C1B3C6Du;
    h ^= h >> 13;
    return h;
}
BGR_HD u64 bucket_of(u64 key, u64 n_buckets) {
#if defined(__CUDA_ARCH__)
    return __umulhi(key_hash(key), (u32)n_buckets);
#else
    return ((u64)key_hash(key) * (u64)(u32)n_buckets) >> 32;
#endif
}

// ---- Bloom filter of the seed scan. One 32-bit word per probe, three bits per key.
//   word  = umulhi(h1, n_bloom)
//   bits  = bloom_pattern(bloom_pidx(p)): one of 1024 fixed three-bit patterns, where p = h1 (single-hash mode:
//           the shared-memory tier) or p = h2 (two-hash mode: large filters).
// h1 / h2 take a (k-1)-mer code AND its reverse complement and are symmetric in the two, so the scan does not
// need the canonical minimum of the pair; they are multiply-adds only, and the pattern comes out of a table
// (shared memory in the scan kernel): the scan computes one filter probe per window and is bound by the
// integer ALU pipe, while the multiply-add pipe and the load/store pipe have room
// (profiles/r01_ncu_k_scan_tiles_v6_5Mbp.txt: ALU 67 %, FMA 8 %, LSU 9 %).
static const u32 BLOOM_SMEM_WORDS = 32768;  // 128 KB: the largest filter the scan keeps in shared memory
static const u32 BLOOM_PATTERNS = 1024;
BGR_HD u32 bloom_h1(u64 x, u64 r) {
    u32 h = (u32)x * 0x9E3779B1u;
    h += (u32)(x >> 32) * 0x85EBCA77u;
    h += (u32)r * 0x9E3779B1u;
    h += (u32)(r >> 32) * 0x85EBCA77u;
    return h;
}
BGR_HD u32 bloom_h2(u64 x, u64 r) {
    u32 h = (u32)x * 0xC2B2AE3Du;
    h += (u32)(x >> 32) * 0x27D4EB2Fu;
    h += (u32)r * 0xC2B2AE3Du;
    h += (u32)(r >> 32) * 0x27D4EB2Fu;
    return h;
}
BGR_HD u32 bloom_pidx(u32 p) { return (p * 0x2C1B3C6Du) >> 22; }
// pattern i: three DISTINCT bits of a 32-bit word
BGR_HD u32 bloom_pattern(u32 i) {
    u32 r = (i + 1) * 0x9E3779B1u;
    r ^= r >> 16;
    r *= 0x85EBCA6Bu;
    r ^= r >> 13;
    const u32 b0 = r & 31;
    const u32 b1 = (b0 + 1 + ((r >> 5) & 1023) % 31) & 31;
    const u32 lo = b0 < b1 ? b0 : b1, hi = b0 < b1 ? b1 : b0;
    u32 b2 = ((r >> 15) & 1023) % 30;
    if (b2 >= lo) ++b2;
    if (b2 >= hi) ++b2;
    return (1u << b0) | (1u << b1) | (1u << b2);
}
BGR_HD u32 bloom_word_of(u32 h1, u32 n_bloom) {
#if defined(__CUDA_ARCH__)
    return __umulhi(h1, n_bloom);
#else
    return (u32)(((u64)h1 * (u64)n_bloom) >> 32);
#endif
}
// Filter modes (BlobHeader::bloom_mode):
//   BLOOM_SMEM_1HASH  small indexes: <= 32768 words, copied into shared memory by every CTA of the scan; word and
//                     bit pattern from one hash
//   BLOOM_WORDS_2HASH large indexes, any k: one 32-bit word per probe somewhere in a filter of up to 64 MB, word
//                     from one hash, pattern from another: one random L2 sector per WINDOW
//   BLOOM_MINIMIZER   large indexes, k = 31: the filter is an array of 64-bit blocks and a (k-1)-mer's block is
//                     chosen by its MINIMIZER -- the smallest hash among its 15 16-mers (hash symmetric in a
//                     16-mer and its reverse complement). Consecutive windows of a read mostly share their
//                     minimizer, so the scan fetches a block once per run of ~8 windows instead of once per window
//                     (the L2 sector rate was what bound the scan on large indexes: 1.6 G sectors per 10 M reads,
//                     profiles/r01_ncu_k_scan_tiles_v9_1Gbp.txt). Inside the block a key sets 4 of 64 bits, chosen
//                     by mm_pat_hash of its first and last 16-mer.
enum { BLOOM_SMEM_1HASH = 0, BLOOM_WORDS_2HASH = 1, BLOOM_MINIMIZER = 2 };
static const u32 MM_M = 16;       // minimizer length: a 16-mer is one 32-bit word
static const u32 MM_KM1 = 30;     // the (k-1) the minimizer mode is built for
static const u32 MM_W = MM_KM1 - MM_M + 1;  // 16-mers per (k-1)-mer
// hash of a 16-mer s and its reverse complement r (symmetric: both strands of a (k-1)-mer see the same values).
// Multiply-adds only (the scan's integer ALU pipe is the busy one). The product term matters: s + r alone takes
// few distinct values for a word and its reversed complement, and a filter addressed by it had a third of its
// blocks empty and 31 % false positives.
BGR_HD u32 mm_hash(u32 s, u32 r) { return (s * r) * 0x9E3779B1u + s * 0x85EBCA77u + r * 0x85EBCA77u; }
// which three bits a (k-1)-mer sets in its block: a hash of its FIRST 16-mer s and of the reverse complement r of its
// LAST 16-mer. The reverse complement of the (k-1)-mer swaps the two, so this too is the same on both strands, and
// the scan gets both words as slices of the chunk's packed words (no rolling 60-bit codes: mm_test_windows).
BGR_HD u32 mm_pat_hash(u32 s, u32 r) { return (s * r) * 0xC2B2AE3Du + s * 0x27D4EB2Fu + r * 0x27D4EB2Fu; }
// block of a (k-1)-mer whose smallest 16-mer hash is hmin (the multiplication spreads the small values a
// minimum takes over the whole range)
BGR_HD u32 mm_block(u32 hmin, u32 n_blocks) {
#if defined(__CUDA_ARCH__)
    return __umulhi(hmin * 0x85EBCA77u, n_blocks);
#else
    return (u32)(((u64)(u32)(hmin * 0x85EBCA77u) * (u64)n_blocks) >> 32);
#endif
}
// pattern i: four distinct bits of a 64-bit block
BGR_HD u64 bloom_pattern64(u32 i) {
    u32 r = (i + 1) * 0x9E3779B1u;
    r ^= r >> 16;
    r *= 0x85EBCA6Bu;
    r ^= r >> 13;
    const u32 b0 = r & 63;
    const u32 b1 = (b0 + 1 + ((r >> 6) & 1023) % 63) & 63;
    const u32 lo = b0 < b1 ? b0 : b1, hi = b0 < b1 ? b1 : b0;
    u32 b2 = ((r >> 16) & 1023) % 62;
    if (b2 >= lo) ++b2;
    if (b2 >= hi) ++b2;
    u64 p = (1ull << b0) | (1ull << b1) | (1ull << b2);
    // a fourth bit, distinct from the three: the ((r >> 26) % 61)-th clear bit of p. Four bits instead of three take
    // the windows of a 3 Gbp index that pass the filter from 3.34 to 2.96 per read (each one a random read of HBM).
    u32 skip = (r >> 26) % 61;
    for (u32 b = 0; b < 64; ++b) {
        if (!((p >> b) & 1)) {
            if (skip == 0) {
                p |= 1ull << b;
                break;
            }
            --skip;
        }
    }
    return p;
}
// smallest 16-mer hash of a 30-mer code
BGR_HD u32 mm_min_hash(u64 key) {
    u32 best = 0xFFFFFFFFu;
    for (u32 i = 0; i < MM_W; ++i) {
        const u32 sft = 2 * (MM_W - 1 - i);
        const u32 s = (u32)(key >> sft);
        const u32 r = (u32)rcb((u64)s, MM_M);
        const u32 h = mm_hash(s, r);
        best = h < best ? h : best;
    }
    return best;
}

// word index and bit pattern of a canonical key (the index builders; km1 = k-1): word modes
// Single-hash tier of 31-mers: the hash is mm_pat_hash of the (k-1)-mer's first 16-mer and of the reverse complement
// of its last one instead of bloom_h1 of the two 60-bit codes -- the same on both strands, and the tile scan gets the
// two words as slices of a chunk's packed words without rolling the codes (scan_chunk_slices).
BGR_HD u32 bloom_h1_of(u64 x, u64 r, u32 km1, bool two_hashes) {
    if (!two_hashes && km1 == MM_KM1) return mm_pat_hash((u32)(x >> (2 * (MM_KM1 - MM_M))), (u32)(r >> (2 * (MM_KM1 - MM_M))));
    return bloom_h1(x, r);
}
BGR_HD void bloom_slot(u64 key, u32 km1, bool two_hashes, u32 n_bloom, u32& word, u32& pattern) {
    const u64 r = rcb(key, km1);
    const u32 h1 = bloom_h1_of(key, r, km1, two_hashes);
    word = bloom_word_of(h1, n_bloom);
    pattern = bloom_pattern(bloom_pidx(two_hashes ? bloom_h2(key, r) : h1));
}
// block index and bit pattern of a canonical 30-mer: minimizer mode (n_blocks = n_bloom / 2)
BGR_HD void bloom_slot_mm(u64 key, u32 n_blocks, u32& block, u64& pattern) {
    block = mm_block(mm_min_hash(key), n_blocks);
    const u32 first = (u32)(key >> (2 * (MM_KM1 - MM_M))), last = (u32)key;
    pattern = bloom_pattern64(bloom_pidx(mm_pat_hash(first, (u32)rcb((u64)last, MM_M))));
}
// filter layout for n_keys keys of k-mers: mode and number of 32-bit words (power of two not required)
inline u64 bloom_layout(u64 n_keys, u32 k, u32& mode) {
    bool force_global = false;
    if (const char* e = getenv("BGR_BLOOM")) force_global = e[0] == 'g';  // test hook: "global" = the large-index tier
    if (n_keys <= 4ull * BLOOM_SMEM_WORDS && !force_global) {  // shared-memory tier: <= 4 keys per word
        mode = BLOOM_SMEM_1HASH;
        u64 w = 4 * n_keys;
        if (w < 1024) w = 1024;
        if (w > BLOOM_SMEM_WORDS) w = BLOOM_SMEM_WORDS;
        return w;
    }
    // global tier: 16 bytes per key while that stays within the cap, never below 2/3 byte per key. The cap is
    // what stays L2-resident next to the read stream: the B200's 126 MB L2 is two partitions that each hold
    // their own copy of read-shared lines (profiles/r01_ncu_k_scan_tiles_v5_1Gbp.txt).
    mode = (k - 1 == MM_KM1) ? BLOOM_MINIMIZER : BLOOM_WORDS_2HASH;
    u64 cap_mb = 32;  // BGR_BLOOM_MB: experiment hook
    if (n_keys > (20ull << 20)) cap_mb = 64;  // beyond ~20 M keys the false positives cost more than the misses
    if (const char* e = getenv("BGR_BLOOM_MB")) {
        const long v = atol(e);
        if (v > 0) cap_mb = (u64)v;
    }
    u64 w = 4 * n_keys;
    const u64 cap_words = cap_mb << 18;
    if (w > cap_words) w = cap_words;
    if (w < n_keys / 6) w = n_keys / 6;
    if (w < 1024) w = 1024;
    return (w + 1) & ~1ull;  // whole 64-bit blocks
}

// word offset of read r's packed words inside a batch whose reads are concatenated characters:
// read r occupies chars [o, o') and gets ceil(len/32)+3 words (one zero pad word in front, two
// behind); floor(o/32) + 4r is a closed form that leaves at least that much room.
BGR_HD u64 read_word_base(u64 char_off, u64 r) { return (char_off >> 5) + 4 * r; }
BGR_HD u64 batch_words(u64 chars, u64 n_reads) { return (chars >> 5) + 4 * n_reads + 8; }

// correction mode: read r's corrected bases start at this word of the batch's corrected buffer
// (ceil(len/32) <= floor((o+len)/32) - floor(o/32) + 1, so one spare word per read is enough)
BGR_HD u64 corrected_word_base(u64 char_off, u64 r) { return (char_off >> 5) + r; }
BGR_HD u64 corrected_words(u64 chars, u64 n_reads) { return (chars >> 5) + n_reads + 2; }

}  // namespace bgr
