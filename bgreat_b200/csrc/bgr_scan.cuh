// Seed scan (replaces getNOverlap / getListOverlap + the BooPHF lookups, reference aligner.cpp:318-378,
// BooPHF.h:783): k_pack_scan (general form), the chunk logic and the two schedules of the tile scan
// (k_scan_tiles_ws, k_scan_tiles). Part of bgr_kernels.cuh.
#pragma once
#include "bgr_dev.cuh"

namespace bgr {

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
        const u32 h1 = bloom_h1_of(num, rc, km1, TWO);
        const u32 m = pat[bloom_pidx(TWO ? bloom_h2(num, rc) : h1)];
        const u32 w = bloom[bloom_word_of(h1, n_bloom)];
        hits |= ((w & m) == m ? 1u : 0u) << j;
        const u64 nb = (inc >> (62 - 2 * j)) & 3;
        num = ((num << 2) | nb) & mask;
        rc = (rc >> 2) | ((3 - nb) << top);
    }
    return count >= 32 ? hits : hits & ((1u << count) - 1);
}

// The same test against the minimizer-addressed filter (BLOOM_MINIMIZER, bgr_common.h; k-1 = 30): the block of a
// window is chosen by the smallest hash among its 15 16-mers, and is looked at only when it differs from the
// previous window's (a run of ~8 windows shares it). The hash of the 16-mer that enters with a window is a few
// multiply-adds on words the rolling codes already hold (its code = the low word of the forward code, its reverse
// complement = the top word of the reverse-complement code); the minimum over the window comes from the van Herk
// scheme over blocks of 15 offsets -- suffix minima of the block a window starts in, a running prefix minimum of
// the block it ends in: three min operations per window. Everything is unrolled, the arrays live in registers.
// Measured, 10 M reads: 3.7 ms at 100 Mbp and 4.4 ms at 1 Gbp against 5.7 / 5.9 ms for the word filter (one L2 sector
// per window); at 3 Gbp (40 M keys, 64 MB of blocks) 6.2 against 6.0 ms: there a warp waits out a memory round trip at
// nearly every window (profiles/r02_ncu_k_scan_tiles_d_3Gbp.txt), which is what the staged form below (mm_stage_blocks /
// mm_test_windows / verify_warp, used by k_scan_tiles_ws) removes. This function is the form with the loads in place:
// k_scan_tiles (BGR_SCAN=classic), chunks with more runs than slots, and the host emulator.
// Measured at 3 Gbp and lost: a first pass that only computes the minimizers, prefetches into L2 and lists the block
// numbers (6.8 ms, twice the L2 sectors); a bit map over a sparser block array so that only occupied blocks are fetched
// (8.8 ms: two dependent loads per run of windows); L2 cache hints (filter loads evict-last, the packed words and hit
// masks stored streaming: nothing at any graph size, with either form of the scan).
// blocks: the filter as 64-bit blocks; pat: bloom_pattern64 table.
BGR_HD u32 scan_chunk_mm(const u64* blocks, u32 n_blocks, const u64* pat, u64 w0, u64 w1, u32 count) {
    const u32 sh = 64 - 2 * MM_KM1, top = 2 * (MM_KM1 - 1);
    const u64 mask = ~0ull >> sh;
    u64 num = w0 >> sh;
    u64 rc = rcb(num, MM_KM1);
    const u64 inc = (w0 << (2 * MM_KM1)) | (w1 >> sh);  // the bases that follow the first window
    // 16-mers at offsets 0..13 lie inside w0: codes are slices of w0, their reverse complements slices of rc64(w0)
    const u64 R = rc64(w0);
    u32 S[MM_W], hb[MM_W];
#pragma unroll
    for (u32 i = 0; i + 1 < MM_W; ++i) S[i] = mm_hash((u32)(w0 >> (32 - 2 * i)), (u32)(R >> (2 * i)));
    u32 P = 0, Q = 0xFFFFFFFFu, hits = 0, cur = 0xFFFFFFFFu;
    u64 blk = 0;
#pragma unroll
    for (u32 j = 0; j < 32; ++j) {
        const u32 h = mm_hash((u32)num, (u32)(rc >> (2 * (MM_KM1 - MM_M))));  // the 16-mer at offset j + 14
        u32 wmin;
        if (j == 0) {
            S[MM_W - 1] = h;
#pragma unroll
            for (int i = (int)MM_W - 2; i >= 0; --i) S[i] = S[i] < S[i + 1] ? S[i] : S[i + 1];
            wmin = S[0];
        } else if (j <= MM_W) {  // offsets 15..29
            hb[j - 1] = h;
            P = (j == 1 || h < P) ? h : P;
            wmin = (j < MM_W) ? (S[j] < P ? S[j] : P) : P;
            if (j == MM_W) {  // the block is complete: its suffix minima replace the first block's
                S[MM_W - 1] = hb[MM_W - 1];
#pragma unroll
                for (int i = (int)MM_W - 2; i >= 0; --i) S[i] = hb[i] < S[i + 1] ? hb[i] : S[i + 1];
            }
        } else {  // offsets 30..45
            if (j <= 2 * MM_W) P = (j == MM_W + 1 || h < P) ? h : P;
            if (j >= MM_W + 2) Q = h < Q ? h : Q;  // offsets 31..: the last window starts one past a block start
            wmin = (j < 2 * MM_W) ? (S[j - MM_W] < P ? S[j - MM_W] : P) : (j == 2 * MM_W ? P : Q);
        }
        const u32 b = mm_block(wmin, n_blocks);
        if (b != cur) {
            cur = b;
#if defined(__CUDA_ARCH__)
            blk = __ldg((const unsigned long long*)blocks + b);
#else
            blk = blocks[b];
#endif
        }
        const u64 m = pat[bloom_pidx(mm_pat_hash((u32)(num >> (2 * (MM_KM1 - MM_M))), (u32)(rc >> (2 * (MM_KM1 - MM_M)))))];
        hits |= ((blk & m) == m ? 1u : 0u) << j;
        const u64 nb = (inc >> (62 - 2 * j)) & 3;
        num = ((num << 2) | nb) & mask;
        rc = (rc >> 2) | ((3 - nb) << top);
    }
    return count >= 32 ? hits : hits & ((1u << count) - 1);
}

// The windows that passed the filter, checked against the key table: the returned mask is exact. The home
// buckets of up to NB survivors are requested together (a chunk has one or two on average, and on a large index
// each is a random sector in HBM: waiting for them one after the other costs a round trip each).
template <int NB>
BGR_HD u32 verify_chunk(const DevIndex& ix, u64 w0, u64 w1, u32 hits) {
    const u32 km1 = ix.k - 1, sh = 64 - 2 * km1;
    u32 m = hits;
    while (m) {
        u32 jj[NB];
        u64 key[NB];
        u64x4 v[NB];
        int n = 0;
#pragma unroll
        for (int t = 0; t < NB; ++t) {
            if (m) {
                const u32 j = (u32)BGR_FFS(m) - 1;
                m &= m - 1;
                const u64 x = (j ? ((w0 << (2 * j)) | (w1 >> (64 - 2 * j))) : w0) >> sh;
                const u64 rc = rcb(x, km1);
                jj[t] = j;
                key[t] = x < rc ? x : rc;
                v[t] = ld_bucket(ix.keys + bucket_of(key[t], ix.n_buckets) * 4);
                n = t + 1;
            }
        }
#pragma unroll
        for (int t = 0; t < NB; ++t) {
            if (t < n) {
                bool found = v[t].a == key[t] || v[t].b == key[t] || v[t].c == key[t] || v[t].d == key[t];
                // a full home bucket: the key may have overflowed into a later one
                if (!found && v[t].d != EMPTY_KEY) found = index_has(ix, key[t]);
                if (!found) hits &= ~(1u << jj[t]);
            }
        }
    }
    return hits;
}

#if defined(__CUDACC__)
// ------------------------------------------------------------------------------ tile scan kernel

static const int ST_THREADS = 1024;
static const u32 WS_WORKERS = ST_THREADS - 32;  // worker threads of the warp-specialised tile kernel

// ------------------------------------------------------------------------------ minimizer tier, loads off the critical path
//
// scan_chunk_mm asks for a filter block at the window where a run starts and tests that window with it at once: in a
// warp some lane starts a run at nearly every one of the 32 unrolled windows, so the warp waits out a memory round
// trip per WINDOW, and then one more per survivor in verify_chunk. On an index far larger than L2 that chain, not
// the rate at which HBM serves random reads, was what the scan took its time from (profiles/r02_scan_experiments.txt:
// 6.2 ms per 10 M reads at 3 Gbp; 4.1 ms without the key-table checks, 2.9 ms without any index load, while a
// micro-benchmark of the same mix of accesses runs at 2.5 times the kernel's rate). The tile kernel's workers
// therefore take a chunk in three steps, none of which waits for memory inside a loop:
//   mm_stage_blocks  the minimizers of all 32 windows (16-mer hashes straight from slices of the chunk's two
//                    words and of their reverse complements: no rolling codes yet); at every run start a
//                    cp.async carries the run's block into a shared-memory slot of the thread -- no register waits
//                    for it. One bit per run start is all that is kept.
//   mm_test_windows  after ONE wait for the thread's copies: the rolling codes, the bit pattern of every window,
//                    tested against the block of its run (shared memory).
//   verify_warp      the survivors of all 32 lanes are listed in shared memory and checked against the key table
//                    side by side, one bucket load per lane: one round trip per chunk of the warp instead of one
//                    per survivor of its unluckiest lane.
static const u32 MM_SLOTS = 10;  // runs of a chunk with a slot (a chunk has 4.8 on average; the others are loaded in place)
static const u32 VQ_CAP = 128;   // survivors listed per round: four per lane
struct VerifyQueue {
    u32 hits[32];
    unsigned short tag[VQ_CAP];  // lane << 5 | window
};

__device__ __forceinline__ void cp_async8(void* dst_shared, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((u32)__cvta_generic_to_shared(dst_shared)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async8_if(u32 dst_shared, const void* src, bool p) {
    asm volatile(
        "{\n"
        ".reg .pred q;\n"
        "setp.ne.u32 q, %2, 0;\n"
        "@q cp.async.ca.shared.global [%0], [%1], 8;\n"
        "}" ::"r"(dst_shared),
        "l"(src), "r"((u32)p)
        : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// The 16-mer at offset T of the 64 bases (A:B:C:D), and its reverse complement: the 16-mer at offset 48 - T of the
// reverse-complemented 64 bases (E:F:G:H). T is a compile-time constant: a slice is one funnel shift.
template <u32 T>
__device__ __forceinline__ u32 mm_s_at(u32 A, u32 B, u32 C, u32 D) {
    return T == 0 ? A : T < 16 ? __funnelshift_l(B, A, 2 * T) : T == 16 ? B : T < 32 ? __funnelshift_l(C, B, 2 * T - 32)
                                                                                 : T == 32 ? C : __funnelshift_l(D, C, 2 * T - 64);
}
template <u32 T>
__device__ __forceinline__ u32 mm_r_at(u32 E, u32 F, u32 G, u32 H) {
    constexpr u32 U = 48 - T;
    return U == 48 ? H : U > 32 ? __funnelshift_l(H, G, 2 * U - 64) : U == 32 ? G : U > 16 ? __funnelshift_l(G, F, 2 * U - 32)
                                                                                   : U == 16 ? F : __funnelshift_l(F, E, 2 * U);
}
template <u32 T>
__device__ __forceinline__ u32 mm_hash_at(u32 A, u32 B, u32 C, u32 D, u32 E, u32 F, u32 G, u32 H) {
    return mm_hash(mm_s_at<T>(A, B, C, D), mm_r_at<T>(E, F, G, H));
}

template <u32 J>
struct MMStage {
    // window J of the chunk (and, recursively, the windows after it): the sliding minimum as in scan_chunk_mm
    static __device__ __forceinline__ void run(const u64* blocks, u32 n_blocks, u32 count, u32 sp_last, u32 A, u32 B, u32 C, u32 D,
                                               u32 E, u32 F, u32 G, u32 H, u32 (&S)[MM_W], u32 (&hb)[MM_W], u32& P, u32& Q,
                                               u32& cur, u32& runs, u32& sp) {
        u32 wmin;
        if (J == 0) {
            wmin = S[0];
        } else if (J <= MM_W) {  // 16-mer offsets 15..29
            const u32 h = mm_hash_at<J + 14>(A, B, C, D, E, F, G, H);
            hb[J - 1] = h;
            P = (J == 1 || h < P) ? h : P;
            constexpr u32 si = J < MM_W ? J : 0;
            wmin = (J < MM_W) ? (S[si] < P ? S[si] : P) : P;
            if (J == MM_W) {  // the block is complete: its suffix minima replace the first block's
                S[MM_W - 1] = hb[MM_W - 1];
#pragma unroll
                for (int i = (int)MM_W - 2; i >= 0; --i) S[i] = hb[i] < S[i + 1] ? hb[i] : S[i + 1];
            }
        } else {  // offsets 30..45
            const u32 h = mm_hash_at<J + 14>(A, B, C, D, E, F, G, H);
            if (J <= 2 * MM_W) P = (J == MM_W + 1 || h < P) ? h : P;
            if (J >= MM_W + 2) Q = h < Q ? h : Q;  // offsets 31..: the last window starts one past a block start
            constexpr u32 si = J > MM_W && J < 2 * MM_W ? J - MM_W : 0;
            wmin = (J < 2 * MM_W) ? (S[si] < P ? S[si] : P) : (J == 2 * MM_W ? P : Q);
        }
        // no branch: a predicated copy, and the slot pointer stops at the last slot (a chunk with more runs than slots
        // is redone by scan_chunk_mm, see tile_chunk_mm)
        const u32 b = mm_block(wmin, n_blocks);
        const bool isnew = (J == 0 || b != cur) && J < count;
        cur = b;
        cp_async8_if(sp, blocks + b, isnew);
        runs |= (isnew ? 1u : 0u) << J;
        sp = min(sp + (isnew ? WS_WORKERS * 8 : 0u), sp_last);
        if constexpr (J + 1 < 32) MMStage<J + 1>::run(blocks, n_blocks, count, sp_last, A, B, C, D, E, F, G, H, S, hb, P, Q, cur, runs, sp);
    }
};

// slot: the thread's first slot; its slot i is slot[i * WS_WORKERS]. Returns the mask of the windows that start a run.
__device__ __forceinline__ u32 mm_stage_blocks(const u64* blocks, u32 n_blocks, u64 w0, u64 w1, u32 count, u64* slot) {
    const u32 A = (u32)(w0 >> 32), B = (u32)w0, C = (u32)(w1 >> 32), D = (u32)w1;
    const u64 R0 = rc64(w0), R1 = rc64(w1);
    const u32 E = (u32)(R1 >> 32), F = (u32)R1, G = (u32)(R0 >> 32), H = (u32)R0;
    u32 S[MM_W], hb[MM_W];
    S[0] = mm_hash_at<0>(A, B, C, D, E, F, G, H);
    S[1] = mm_hash_at<1>(A, B, C, D, E, F, G, H);
    S[2] = mm_hash_at<2>(A, B, C, D, E, F, G, H);
    S[3] = mm_hash_at<3>(A, B, C, D, E, F, G, H);
    S[4] = mm_hash_at<4>(A, B, C, D, E, F, G, H);
    S[5] = mm_hash_at<5>(A, B, C, D, E, F, G, H);
    S[6] = mm_hash_at<6>(A, B, C, D, E, F, G, H);
    S[7] = mm_hash_at<7>(A, B, C, D, E, F, G, H);
    S[8] = mm_hash_at<8>(A, B, C, D, E, F, G, H);
    S[9] = mm_hash_at<9>(A, B, C, D, E, F, G, H);
    S[10] = mm_hash_at<10>(A, B, C, D, E, F, G, H);
    S[11] = mm_hash_at<11>(A, B, C, D, E, F, G, H);
    S[12] = mm_hash_at<12>(A, B, C, D, E, F, G, H);
    S[13] = mm_hash_at<13>(A, B, C, D, E, F, G, H);
    S[14] = mm_hash_at<14>(A, B, C, D, E, F, G, H);
    static_assert(MM_W == 15, "the offsets above");
#pragma unroll
    for (int i = (int)MM_W - 2; i >= 0; --i) S[i] = S[i] < S[i + 1] ? S[i] : S[i + 1];
    u32 P = 0, Q = 0xFFFFFFFFu, cur = 0xFFFFFFFFu, runs = 0;
    u32 sp = (u32)__cvta_generic_to_shared(slot);  // shared-memory address of the next free slot
    const u32 sp_last = sp + (MM_SLOTS - 1) * WS_WORKERS * 8;
    MMStage<0>::run(blocks, n_blocks, count, sp_last, A, B, C, D, E, F, G, H, S, hb, P, Q, cur, runs, sp);
    return runs;
}

// a chunk with more runs than slots (rare): the blocks loaded where they are tested
__device__ __noinline__ u32 scan_chunk_mm_far(const u64* blocks, u32 n_blocks, const u64* pat, u64 w0, u64 w1, u32 count) {
    return scan_chunk_mm(blocks, n_blocks, pat, w0, w1, count);
}

template <u32 J>
struct MMTest {
    // window J (and the ones after it): the block of its run, its bit pattern from its first 16-mer and the reverse
    // complement of its last one (mm_pat_hash), the test
    static __device__ __forceinline__ void run(const u64* pat, u32 runs, u32 A, u32 B, u32 C, u32 D, u32 E, u32 F, u32 G, u32 H,
                                               const u64*& sp, u64& blk, u32& hits) {
        if ((runs >> J) & 1u) {
            blk = *sp;
            sp += WS_WORKERS;
        }
        const u64 m = pat[bloom_pidx(mm_pat_hash(mm_s_at<J>(A, B, C, D), mm_r_at<J + MM_W - 1>(E, F, G, H)))];
        hits |= ((blk & m) == m ? 1u : 0u) << J;
        if constexpr (J + 1 < 32) MMTest<J + 1>::run(pat, runs, A, B, C, D, E, F, G, H, sp, blk, hits);
    }
};

// the chunk's runs must all have a slot: popc(runs) <= MM_SLOTS
__device__ __forceinline__ u32 mm_test_windows(const u64* pat, u64 w0, u64 w1, u32 count, u32 runs, const u64* slot) {
    const u32 A = (u32)(w0 >> 32), B = (u32)w0, C = (u32)(w1 >> 32), D = (u32)w1;
    const u64 R0 = rc64(w0), R1 = rc64(w1);
    const u32 E = (u32)(R1 >> 32), F = (u32)R1, G = (u32)(R0 >> 32), H = (u32)R0;
    u32 hits = 0;
    const u64* sp = slot;
    u64 blk = 0;
    MMTest<0>::run(pat, runs, A, B, C, D, E, F, G, H, sp, blk, hits);
    return count >= 32 ? hits : hits & ((1u << count) - 1);
}

__device__ __noinline__ bool index_has_far(const DevIndex& ix, u64 key) { return index_has(ix, key); }

// scan_chunk<false, 30> without the rolling codes: the single-hash filter of 31-mers is addressed by mm_pat_hash of a
// window's first 16-mer and of the reverse complement of its last one (bloom_h1_of), both slices of the chunk's two
// packed words and of their reverse complements. 14 instructions per window against 24 (the shared-memory tier is
// bound by the integer pipes).
template <u32 J>
struct SliceTest {
    static __device__ __forceinline__ void run(const u32* bloom, u32 n_bloom, const u32* pat, u32 A, u32 B, u32 C, u32 D, u32 E,
                                               u32 F, u32 G, u32 H, u32& hits) {
        const u32 h1 = mm_pat_hash(mm_s_at<J>(A, B, C, D), mm_r_at<J + MM_W - 1>(E, F, G, H));
        const u32 m = pat[bloom_pidx(h1)];
        const u32 w = bloom[bloom_word_of(h1, n_bloom)];
        hits |= ((w & m) == m ? 1u : 0u) << J;
        if constexpr (J + 1 < 32) SliceTest<J + 1>::run(bloom, n_bloom, pat, A, B, C, D, E, F, G, H, hits);
    }
};
__device__ __forceinline__ u32 scan_chunk_slices(const u32* bloom, u32 n_bloom, const u32* pat, u64 w0, u64 w1, u32 count) {
    const u32 A = (u32)(w0 >> 32), B = (u32)w0, C = (u32)(w1 >> 32), D = (u32)w1;
    const u64 R0 = rc64(w0), R1 = rc64(w1);
    const u32 E = (u32)(R1 >> 32), F = (u32)R1, G = (u32)(R0 >> 32), H = (u32)R0;
    u32 hits = 0;
    SliceTest<0>::run(bloom, n_bloom, pat, A, B, C, D, E, F, G, H, hits);
    return count >= 32 ? hits : hits & ((1u << count) - 1);
}

// The windows of the warp's 32 chunks that passed the filter, checked against the key table: all lanes call this
// (hits = 0 for a lane without a chunk); the returned mask is exact.
__device__ __forceinline__ u32 verify_warp(const DevIndex& ix, const u32 km1, VerifyQueue& Q, const u32 lane, const u64 w0,
                                           const u64 w1, u32 hits) {
    const u32 FULL = 0xFFFFFFFFu;
    if (!__any_sync(FULL, hits != 0)) return hits;
    const u32 sh = 64 - 2 * km1;
    Q.hits[lane] = hits;
    u32 m = hits;
    for (;;) {
        const u32 have = (u32)__popc(m);
        const u32 c = have < 4 ? have : 4;
        u32 incl = c;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const u32 x = __shfl_up_sync(FULL, incl, d);
            if (lane >= (u32)d) incl += x;
        }
        const u32 total = __shfl_sync(FULL, incl, 31);
        u32 at = incl - c;
#pragma unroll 1
        for (u32 t = 0; t < c; ++t) {
            const u32 j = (u32)__ffs(m) - 1;
            m &= m - 1;
            Q.tag[at++] = (unsigned short)((lane << 5) | j);
        }
        __syncwarp();
#pragma unroll 1
        for (u32 e0 = 0; e0 < total; e0 += 32) {
            const u32 e = e0 + lane;
            const bool valid = e < total;
            const u32 tag = valid ? Q.tag[e] : 0u;
            const u32 owner = tag >> 5, j = tag & 31;
            const u64 o0 = __shfl_sync(FULL, w0, owner), o1 = __shfl_sync(FULL, w1, owner);
            if (valid) {
                const u64 x = (j ? ((o0 << (2 * j)) | (o1 >> (64 - 2 * j))) : o0) >> sh;
                const u64 rc = rcb(x, km1);
                const u64 key = x < rc ? x : rc;
                const u64x4 v = ld_bucket(ix.keys + bucket_of(key, ix.n_buckets) * 4);
                bool found = v.a == key || v.b == key || v.c == key || v.d == key;
                // a full home bucket: the key may have overflowed into a later one
                if (!found && v.d != EMPTY_KEY) found = index_has_far(ix, key);
                if (!found) atomicAnd(&Q.hits[owner], ~(1u << j));
            }
        }
        __syncwarp();
        if (!__any_sync(FULL, m != 0)) break;
    }
    return Q.hits[lane];
}

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
        // a waiting warp must not take issue slots from the working ones: the bookkeeper warp of k_scan_tiles_ws spent
        // 18 % of the kernel's issued instructions in this loop (profiles/r02_lines_k_scan_tiles_async_3Gbp.txt)
        if (!done) __nanosleep(100);
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
    u32* pat;  // BLOOM_PATTERNS filter bit patterns (64-bit ones in minimizer mode)
    unsigned char* chars[2];
    u64* off;     // ST_MAX_R + 2
    u32* prefix;  // ST_MAX_R + 2
    u32* flags;   // ST_MAX_R + 2
    u64* mbar;    // 2
    u32* wsum;    // 32
};
__host__ __device__ inline size_t tile_smem_bytes(u32 bloom_words_in_smem, u32 tile_cb) {
    size_t b = ((size_t)bloom_words_in_smem * 4 + 127) & ~(size_t)127;
    b += BLOOM_PATTERNS * 8;
    b += 2 * ((size_t)tile_cb + ST_CHAR_SLACK);
    b += (ST_MAX_R + 2) * 8 + 2 * (ST_MAX_R + 2) * 4 + 2 * 8 + 32 * 4;
    return b + 128;
}

// One chunk of a tile (both schedules of the tile scan): chunk j of the tile whose reads start at r0, with the
// tile's offsets / chunk prefix sums / flag words in shared memory and its characters (or 2-bit words) at cb,
// whose first base is base gs of the batch. Packs, stores the words for the extension kernel, scans, verifies,
// stores the hit mask and flags the read. Returns the number of windows that passed the filter.
// SLICES: the single-hash filter of 31-mers tested from slices (scan_chunk_slices) instead of rolling codes
template <bool BLOOM_SMEM, bool PACKED, u32 KM1S, bool SLICES = false>
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
    // the shared-memory tier is the single-hash tier; large indexes of 31-mers have the minimizer filter, other
    // k the two-hash word filter (bloom_layout)
    u32 hits;
    if (!BLOOM_SMEM && KM1S == MM_KM1) hits = scan_chunk_mm((const u64*)bloom, ix.n_bloom / 2, (const u64*)pat, w0, w1, count);
    else if (SLICES && BLOOM_SMEM && KM1S == MM_KM1) hits = scan_chunk_slices(bloom, ix.n_bloom, pat, w0, w1, count);
    else hits = scan_chunk<!BLOOM_SMEM, KM1S>(bloom, ix.n_bloom, pat, km1, w0, w1, count);
    const u32 passed = (u32)__popc(hits);
    hits = verify_chunk<1>(ix, w0, w1, hits);
    a.H1[wb + l] = hits;
    const u32 fl = (hits ? TS_FLAG_HIT : 0u) | (bad ? TS_FLAG_BAD : 0u);
    if (fl) atomicOr(&s_flags[rd], fl);
    return passed;
}

// The same chunk on the minimizer tier of the warp-specialised kernel, in the three steps of mm_stage_blocks /
// mm_test_windows / verify_warp. Called by whole warps; `active`: the lane has a chunk (j < chunks of the tile).
template <bool PACKED>
__device__ __forceinline__ u32 tile_chunk_mm(const DevIndex& ix, const BatchArgs& a, const u64* pat, const unsigned char* cb,
                                             const u64 gs, const u64* s_off, const u32* s_prefix, u32* s_flags, const u32 cnt,
                                             const u32 r0, const u32 j, const bool active, u64* slot, VerifyQueue& Q,
                                             const u32 lane) {
    u64 w0 = 0, w1 = 0, wb = 0;
    u32 hits = 0, passed = 0, bad = 0, rd = 0, l = 0;
    if (active) {
#pragma unroll
        for (u32 sft = ST_MAX_R / 2; sft; sft >>= 1) {
            const u32 c = rd + sft;
            if (c < cnt && s_prefix[c] <= j) rd = c;
        }
        l = j - s_prefix[rd];
        const u64 ro = s_off[rd];
        const u32 rlen = (u32)(s_off[rd + 1] - ro);
        const u32 rnwin = rlen - MM_KM1 + 1;
        const u32 rchunks = (rnwin + 31) >> 5;
        wb = read_word_base(ro, r0 + rd);
        if (PACKED) packed_span((const u64*)cb, ro - gs + 32 * l, w0, w1);
        else pack_span(cb + (u32)(ro - gs) + 32 * l, w0, w1, bad);
        clip_span(rlen - 32 * l, w0, w1);
        const u32 count = min(32u, rnwin - 32 * l);
        const u64* blocks = (const u64*)ix.bloom;
        const u32 n_blocks = ix.n_bloom / 2;
        const u32 runs = mm_stage_blocks(blocks, n_blocks, w0, w1, count, slot);
        // packed words for the extension kernel (layout of read_word_base: pad, words, 2 pads), while the blocks travel
        u64* gF = a.F + wb;
        const u32 nw = (rlen + 31) >> 5;
        gF[1 + l] = w0;
        if (l == 0) gF[0] = 0;
        if (l + 1 == rchunks) {
            if (l + 1 < nw) gF[2 + l] = w1;
            gF[nw + 1] = 0;
            gF[nw + 2] = 0;
        }
        // (Writing the words between this read's and the next one's as well -- so that every sector of F and H1 is
        // written in full and L2 need not fill it from HBM -- was measured: nothing at 3 Gbp, 3 % slower at 100 Mbp.)
        cp_async_wait_all();
        if ((u32)__popc(runs) <= MM_SLOTS) hits = mm_test_windows(pat, w0, w1, count, runs, slot);
        else hits = scan_chunk_mm_far(blocks, n_blocks, pat, w0, w1, count);
        passed = (u32)__popc(hits);
    }
    hits = verify_warp(ix, MM_KM1, Q, lane, w0, w1, hits);
    if (active) {
        a.H1[wb + l] = hits;
        const u32 fl = (hits ? TS_FLAG_HIT : 0u) | (bad ? TS_FLAG_BAD : 0u);
        if (fl) atomicOr(&s_flags[rd], fl);
    }
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
        q += BLOOM_PATTERNS * 8;
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
    for (u32 i = threadIdx.x; i < BLOOM_PATTERNS; i += ST_THREADS) {
        if (!BLOOM_SMEM && KM1S == MM_KM1) ((u64*)S.pat)[i] = bloom_pattern64(i);
        else S.pat[i] = bloom_pattern(i);
    }

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

// mm_async: the minimizer tier's block slots and verify queues (mm_stage_blocks / verify_warp)
__host__ __device__ inline size_t tile_ws_smem_bytes(u32 bloom_words_in_smem, u32 tile_cb, bool mm_async) {
    size_t b = ((size_t)bloom_words_in_smem * 4 + 127) & ~(size_t)127;
    b += BLOOM_PATTERNS * 8;
    b += 2 * ((size_t)tile_cb + ST_CHAR_SLACK);
    b += 2 * ((sizeof(TileMeta) + 15) & ~(size_t)15) + 6 * 8;
    if (mm_async) b += (size_t)MM_SLOTS * WS_WORKERS * 8 + (WS_WORKERS / 32) * sizeof(VerifyQueue);
    return b + 128;
}

__device__ __forceinline__ void mbar_arrive(u64* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// STAGED: the minimizer tier in the three steps of mm_stage_blocks / mm_test_windows / verify_warp (see there)
template <bool BLOOM_SMEM, bool PACKED, u32 KM1S, bool STAGED>
__global__ void __launch_bounds__(ST_THREADS, 1) k_scan_tiles_ws(DevIndex ix, BatchArgs a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    u32 *s_bloom, *s_pat;
    unsigned char* s_chars[2];
    TileMeta* M;
    u64 *bar_tma, *bar_full, *bar_done, *s_slots;
    VerifyQueue* s_vq;
    constexpr bool MM_ASYNC = STAGED && !BLOOM_SMEM && KM1S == MM_KM1;  // the minimizer tier (bloom_layout)
    {
        unsigned char* q = smem_raw;
        s_bloom = (u32*)q;
        q += BLOOM_SMEM ? (((size_t)ix.n_bloom * 4 + 127) & ~(size_t)127) : 0;
        s_pat = (u32*)q;
        q += BLOOM_PATTERNS * 8;
        s_chars[0] = q;
        q += a.tile_cb + ST_CHAR_SLACK;
        s_chars[1] = q;
        q += a.tile_cb + ST_CHAR_SLACK;
        M = (TileMeta*)q;  // sizeof(TileMeta) is a multiple of 8; two of them, each padded to 16
        q += 2 * ((sizeof(TileMeta) + 15) & ~(size_t)15);
        bar_tma = (u64*)q;
        bar_full = bar_tma + 2;
        bar_done = bar_tma + 4;
        q += 6 * 8;
        s_slots = (u64*)q;
        q += MM_ASYNC ? (size_t)MM_SLOTS * WS_WORKERS * 8 : 0;
        s_vq = (VerifyQueue*)q;
    }
    auto meta = [&](u32 b) -> TileMeta& { return *(TileMeta*)((unsigned char*)M + b * ((sizeof(TileMeta) + 15) & ~(size_t)15)); };
    const u32 tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const u32 R = a.tile_r, n = a.n, G = gridDim.x;
    const u32 ntiles = (n + R - 1) / R;
    const u32 km1 = KM1S ? KM1S : ix.k - 1;
    const bool exhaustive = a.mode != BGR_MODE_GREEDY;
    const u32* bloom = BLOOM_SMEM ? s_bloom : ix.bloom;
    for (u32 i = tid; i < BLOOM_PATTERNS; i += ST_THREADS) {
        if (!BLOOM_SMEM && KM1S == MM_KM1) ((u64*)s_pat)[i] = bloom_pattern64(i);
        else s_pat[i] = bloom_pattern(i);
    }
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
                if (MM_ASYNC) {
                    for (u32 jb = 0; jb < T; jb += WS_WORKERS)  // whole warps: verify_warp is a warp's joint step
                        c_verify += tile_chunk_mm<PACKED>(ix, a, (const u64*)s_pat, cb, gs, m.off, m.prefix, m.flags, cnt, r0,
                                                          jb + wtid, jb + wtid < T, s_slots + wtid, s_vq[warp - 1], lane);
                } else {
                    for (u32 j = wtid; j < T; j += WS_WORKERS)
                        c_verify += tile_chunk<BLOOM_SMEM, PACKED, KM1S, true>(ix, a, bloom, s_pat, km1, cb, gs, m.off, m.prefix,
                                                                          m.flags, cnt, r0, j);
                }
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

}  // namespace bgr
