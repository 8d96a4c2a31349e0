// TEST INFRASTRUCTURE ONLY -- host single-stepper for the device code.
//
// The build container has no GPU, so the per-read __host__ __device__ bodies of
// bgreat_b200/csrc/bgr_kernels.cuh (pack4, scan_window, greedy_read, exhaustive_read and
// everything under them) are compiled here for the host and driven by plain loops that
// stand in for the kernels' thread/warp structure. This lets `pytest -m "not gpu"` check
// the device LOGIC against the oracle. It is not a fallback: nothing under bgreat_b200/
// links, loads or calls this library, and the product fails loudly without a GPU.
#include <string.h>

#include <string>
#include <vector>

#include "../../bgreat_b200/csrc/bgr_host.h"
#include "../../bgreat_b200/csrc/bgr_kernels.cuh"

using namespace bgr;

// how: bit 0 = the batch goes through the host packer (bgr_pack.cpp) and the scan takes its bases from the
// 2-bit stream (k_scan_tiles<.., true>, packed_span), flagged reads through the character form; bit 1 = no
// tile kernel, every read through the general form (k_pack_scan), as for batches of very long reads.
extern "C" int emu_map2(const char* useqs, const uint64_t* uoff, uint32_t nu, uint32_t k, const char* seqs,
                        const uint64_t* off, uint32_t n, uint32_t mode, uint32_t errors, uint32_t effort,
                        uint32_t scratch_ints, bgr_result* res, int32_t* pool, uint64_t pool_cap, uint64_t* used,
                        uint64_t counters[BGR_C_COUNT], int* scratch_overflow, uint32_t how) {
    const bool packed = (how & 1u) != 0, general_only = (how & 2u) != 0;
    std::vector<uint64_t> blob;
    std::string err;
    int rc = build_index_blob(useqs, uoff, nu, k, blob, err);
    if (rc) return rc;
    BlobHeader h;
    memcpy(&h, blob.data(), sizeof h);
    char* B = (char*)blob.data();
    DevIndex ix;
    ix.keys = (const u64*)(B + h.off_keys);
    ix.entries = (const Entry*)(B + h.off_entries);
    ix.meta = (const UnitigMeta*)(B + h.off_meta);
    ix.pool = (const u64*)(B + h.off_pool);
    ix.n_buckets = h.n_buckets;
    ix.k = h.k;
    ix.bloom = (const u32*)(B + h.off_bloom);
    ix.n_bloom = (u32)h.n_bloom;
    ix.bloom_mode = h.bloom_mode;

    const u64 chars = off[n];
    const u64 words = batch_words(chars, n);
    std::vector<u64> F(words, 0xDEADBEEFDEADBEEFull), NM(words, 0xDEADBEEFDEADBEEFull);
    std::vector<u32> H1(words, 0xFFFFFFFFu), H2(words, 0xFFFFFFFFu), info(n + 1, 0), list(n + 1, 0);
    std::vector<i32> scratch((size_t)n * scratch_ints + 1, 0);
    DevCounters cnt;
    ScanCounters scnt;
    memset(&cnt, 0, sizeof cnt);
    memset(&scnt, 0, sizeof scnt);
    BatchArgs a;
    a.seq = seqs;
    a.off = off;
    a.n = n;
    a.F = F.data();
    a.NM = NM.data();
    a.H1 = H1.data();
    a.H2 = H2.data();
    a.info = info.data();
    a.res = res;
    a.scratch = scratch.data();
    a.scratch_ints = scratch_ints;
    a.sc_row = 1;
    a.sc_ent = n;
    a.pool = pool;
    a.pool_cap = pool_cap;
    a.cnt = &cnt;
    a.scnt = &scnt;
    a.list = list.data();
    a.slow_list = nullptr;
    a.tile_r = 0;
    a.tile_cb = 0;
    a.ext_lazy = 8;
    a.errors = errors;
    a.effort = effort;
    a.mode = mode;
    const u32 km1 = k - 1;
    const bool exhaustive = mode != BGR_MODE_GREEDY;

    // ---- k_scan_tiles: one "thread" per chunk of 32 windows (pack_span / scan_chunk / verify_chunk);
    //      a read with a non-ACGT character goes to the general form below, like on the device
    std::vector<unsigned char> padded(seqs, seqs + chars);
    padded.resize(chars + 128, 'A');
    // packed upload: the stream and the flag bitmap as the host packer makes them (3 threads: the cuts between
    // workers are part of what is tested); the words behind the stream hold garbage on the device
    std::vector<u64> pk(((chars + 31) >> 5) + 8, 0xDEADBEEFDEADBEEFull);
    std::vector<u32> bad_bits((n + 31) / 32 + 1, 0xFFFFFFFFu);
    if (packed) pack_reads_2bit(seqs, off, n, 3, pk.data(), bad_bits.data());
    a.pk = packed ? pk.data() : nullptr;
    std::vector<u32> pat(BLOOM_PATTERNS);
    for (u32 i = 0; i < BLOOM_PATTERNS; ++i) pat[i] = bloom_pattern(i);
    std::vector<u64> pat64(BLOOM_PATTERNS);
    for (u32 i = 0; i < BLOOM_PATTERNS; ++i) pat64[i] = bloom_pattern64(i);
    for (u32 r = 0; r < n; ++r) {
        const u64 o = off[r];
        const u32 len = (u32)(off[r + 1] - o);
        const u64 wb = read_word_base(o, r);
        const u32 nw = (len + 31) >> 5;
        u64* gF = a.F + wb;
        u64* gN = a.NM + wb;
        const u32 nwin_t = (len >= km1) ? len - km1 + 1 : 0;
        const u32 chunks = (nwin_t + 31) >> 5;
        const bool flagged = packed && (bad_bits[r >> 5] >> (r & 31) & 1u);
        bool tile_bad = flagged || general_only, tile_hit = false;
        for (u32 l = 0; l < chunks && !general_only; ++l) {
            u64 w0, w1;
            u32 bad = 0;
            // on the device the span may run into the next read's characters: only the part inside the
            // read can flag it here, which is the exact rule (a spurious flag just takes the general path)
            unsigned char span[64];
            for (u32 j = 0; j < 64; ++j) span[j] = (32 * l + j < len) ? padded[o + 32 * l + j] : (unsigned char)'A';
            if (packed) packed_span(pk.data(), o + 32 * l, w0, w1);
            else pack_span(span, w0, w1, bad);
            clip_span(len - 32 * l, w0, w1);
            gF[1 + l] = w0;
            if (l == 0) gF[0] = 0;
            if (l + 1 == chunks) {
                if (l + 1 < nw) gF[2 + l] = w1;
                gF[nw + 1] = 0;
                gF[nw + 2] = 0;
            }
            const u32 count = nwin_t - 32 * l < 32u ? nwin_t - 32 * l : 32u;
            // both the k-1 = 30 specialisation and the run-time form, like the kernel's dispatch
            u32 hits;
            if (ix.bloom_mode == BLOOM_MINIMIZER) {
                hits = scan_chunk_mm((const u64*)ix.bloom, ix.n_bloom / 2, pat64.data(), w0, w1, count);
                // window by window from the definition (bloom_slot_mm of the window's canonical code)
                u32 want = 0;
                for (u32 j = 0; j < count; ++j) {
                    const u64 x = (j ? ((w0 << (2 * j)) | (w1 >> (64 - 2 * j))) : w0) >> (64 - 2 * km1);
                    const u64 rcx = rcb(x, km1);
                    u32 block;
                    u64 pattern;
                    bloom_slot_mm(x < rcx ? x : rcx, ix.n_bloom / 2, block, pattern);
                    u32 block2;
                    u64 pattern2;
                    bloom_slot_mm(x < rcx ? rcx : x, ix.n_bloom / 2, block2, pattern2);  // strand symmetry
                    if (block2 != block || pattern2 != pattern) return 1004;
                    if ((((const u64*)ix.bloom)[block] & pattern) == pattern) want |= 1u << j;
                }
                if (want != hits) return 1003;
            } else if (ix.bloom_mode == BLOOM_WORDS_2HASH)
                hits = km1 == 30 ? scan_chunk<true, 30>(ix.bloom, ix.n_bloom, pat.data(), km1, w0, w1, count)
                                 : scan_chunk<true, 0>(ix.bloom, ix.n_bloom, pat.data(), km1, w0, w1, count);
            else
                hits = km1 == 30 ? scan_chunk<false, 30>(ix.bloom, ix.n_bloom, pat.data(), km1, w0, w1, count)
                                 : scan_chunk<false, 0>(ix.bloom, ix.n_bloom, pat.data(), km1, w0, w1, count);
            if (km1 == 30 && ix.bloom_mode != BLOOM_MINIMIZER) {  // the two forms must agree
                const u32 h0 = ix.bloom_mode == BLOOM_WORDS_2HASH
                                   ? scan_chunk<true, 0>(ix.bloom, ix.n_bloom, pat.data(), km1, w0, w1, count)
                                   : scan_chunk<false, 0>(ix.bloom, ix.n_bloom, pat.data(), km1, w0, w1, count);
                if (h0 != hits) return 1003;
            }
            scnt.c[BGR_C_VERIFIED] += (u64)__builtin_popcount(hits);
            hits = (r & 1) ? verify_chunk<1>(ix, w0, w1, hits) : verify_chunk<2>(ix, w0, w1, hits);
            a.H1[wb + l] = hits;
            tile_bad |= bad != 0;
            tile_hit |= hits != 0;
        }
        if (!tile_bad) {
            a.info[r] = tile_hit ? INFO_HIT1 : 0u;
            scnt.c[BGR_C_WINDOWS] += nwin_t;
            scnt.c[BGR_C_READS] += 1;
            if (exhaustive) scnt.c[BGR_C_OVERLAPS] += nwin_t;
            if (tile_hit || (exhaustive && len == km1)) {
                a.list[scnt.list_n++] = r;
            } else {
                res[r].flags_len = no_hit_flags(exhaustive);
                res[r].path_off = 0;
                scnt.c[exhaustive ? BGR_C_NOTALIGNED : BGR_C_NOOVERLAP] += 1;
            }
            continue;
        }
        // ---- k_pack_scan, one "warp" per read handed over
        gF[0] = gN[0] = 0;
        gF[nw + 1] = gN[nw + 1] = 0;
        gF[nw + 2] = gN[nw + 2] = 0;
        bool anyN = false;
        if (packed && !flagged) {  // pack_and_scan with src == null: the words come out of the stream
            for (u32 j = 0; j < nw; ++j) {
                u64 w = ld_bases(a.pk, (i64)(o + 32ull * j));
                const u32 rem = len - 32 * j;
                if (rem < 32) w &= ~(~0ull >> (2 * rem));
                gF[1 + j] = w;
            }
        }
        for (u32 c0 = 0; (!packed || flagged) && c0 < nw * 32; c0 += 4) {
            u32 x = 0x41414141u;
            if (c0 < len) {
                u32 left = len - c0;
                unsigned char b[4] = {'A', 'A', 'A', 'A'};
                for (u32 j = 0; j < 4 && j < left; ++j) b[j] = (unsigned char)seqs[o + c0 + j];
                x = b[0] | (b[1] << 8) | (b[2] << 16) | ((u32)b[3] << 24);
            }
            u32 code8, nm8;
            pack4(x, code8, nm8);
            u32 byte = packed_byte(c0);
            ((unsigned char*)gF)[byte] = (unsigned char)code8;
            ((unsigned char*)gN)[byte] = (unsigned char)nm8;
            anyN |= nm8 != 0;
        }
        const u32 nwin = (len >= km1) ? len - km1 + 1 : 0;
        u32 any1 = 0, any2 = 0;
        if (!anyN) {
            for (u32 base = 0; base < nwin; base += 32 * PS_WPL) {
                u32 ballot[PS_WPL] = {0, 0, 0, 0};
                for (u32 lane = 0; lane < 32; ++lane) {
                    u32 hits = scan_lane_rolling(ix, gF + 1, base + PS_WPL * lane, nwin);
                    for (u32 j = 0; j < PS_WPL; ++j) ballot[j] |= ((hits >> j) & 1u) << lane;
                }
                for (u32 lane = 0; lane < 4; ++lane)
                    if (base + 32 * lane < nwin) a.H1[wb + (base >> 5) + lane] = transpose_hits(ballot, lane);
                any1 |= ballot[0] | ballot[1] | ballot[2] | ballot[3];
            }
        } else {
            for (u32 base = 0; base < nwin; base += 32) {
                u32 m1 = 0, m2 = 0;
                for (u32 lane = 0; lane < 32; ++lane) {
                    u32 i = base + lane;
                    bool h1 = false, h2 = false;
                    if (i < nwin) scan_window(ix, gF + 1, gN + 1, true, exhaustive, i, h1, h2);
                    m1 |= (u32)h1 << lane;
                    m2 |= (u32)h2 << lane;
                }
                any1 |= m1;
                any2 |= m2;
                a.H1[wb + (base >> 5)] = m1;
                a.H2[wb + (base >> 5)] = m2;
            }
        }
        const u32 inf = (anyN ? INFO_HAS_N : 0u) | (any1 ? INFO_HIT1 : 0u) | (any2 ? INFO_HIT2 : 0u);
        a.info[r] = inf;
        scnt.c[BGR_C_WINDOWS] += nwin;
        scnt.c[BGR_C_READS] += 1;
        if (exhaustive) scnt.c[BGR_C_OVERLAPS] += nwin;
        const bool pass_on = (inf & INFO_HIT1) || (exhaustive && len == km1);
        if (pass_on) {
            a.list[scnt.list_n++] = r;
        } else {
            res[r].flags_len = no_hit_flags(exhaustive);
            res[r].path_off = 0;
            scnt.c[exhaustive ? BGR_C_NOTALIGNED : BGR_C_NOOVERLAP] += 1;
        }
    }

    // ---- k_extend_*, one "thread" per listed read, then emit_result / count_classes
    *scratch_overflow = 0;
    for (u32 t = 0; t < scnt.list_n; ++t) {
        const u32 r = a.list[t];
        Scratch sc;
        sc.base = a.scratch + t;
        sc.stride = n;
        sc.cap = scratch_ints;
        sc.overflow = false;
        u32 flags = 0, np = 0, nl = 0, overlaps = 0;
        if (mode == BGR_MODE_GREEDY) {
            // the persistent kernel's per-lane life cycle: start, step until done, finish
            GreedyState g;
            g.start(ix, a, r, sc.base);
            while (!g.done()) g.step(ix, a);
            flags = g.flags;
            np = g.finish();  // pathBegin reversed in place: scratch [0, np) is the final path
            nl = 0;
            sc.overflow = g.sc.overflow;
            // cross-check: the loop-shaped kernel body against the walker-shaped statement
            Scratch sc2 = sc;
            std::vector<i32> tmp((size_t)scratch_ints + 1, 0);
            sc2.base = tmp.data();
            sc2.stride = 1;
            sc2.overflow = false;
            u32 np2 = 0, nl2 = 0;
            u32 flags2 = greedy_read_walkers(ix, a, r, sc2, np2, nl2);
            bool same = flags2 == flags && (!(flags & BGR_F_MAPPED) || np2 == np);
            if (same && (flags & BGR_F_MAPPED) && !sc.overflow && !sc2.overflow)
                for (u32 j = 0; j < np; ++j) same = same && sc.get(j) == (j < nl2 ? sc2.get(nl2 - 1 - j) : sc2.get(j));
            if (!same && !sc.overflow && !sc2.overflow) return 1000;  // the two formulations disagree
        } else if (mode == BGR_MODE_EXHAUSTIVE || mode == BGR_MODE_EXHAUSTIVE_PATH) {
            // the persistent kernel's per-lane life cycle: start, one search node per step until done
            ExhaustiveState x;
            // the kernel's shared-memory list heads (ExhaustiveState::fast): a host buffer for every other read, none
            // for the rest, so that both placements of the lists are single-stepped
            std::vector<i32> fastbuf((size_t)XF_INTS * XF_STRIDE, 0);
            x.fast = (r & 1u) ? fastbuf.data() : nullptr;
            x.start(ix, a, r, sc.base);
            while (!x.done()) x.step(ix, a);
            flags = x.flags;
            np = x.n;
            sc.overflow = x.sc.overflow;
            // cross-check against the loop-shaped statement of the same search
            std::vector<i32> tmp((size_t)scratch_ints + 1, 0);
            Scratch sc2 = sc;
            sc2.base = tmp.data();
            sc2.stride = 1;
            sc2.overflow = false;
            u32 np2 = 0;
            const u32 flags2 = exhaustive_read(ix, a, r, sc2, np2, overlaps);
            if (!sc.overflow && !sc2.overflow) {
                bool same = flags2 == flags && np2 == np;
                for (u32 j = 0; same && j < np; ++j) same = sc.get(j) == sc2.get(j);
                if (!same) return 1001;  // the two formulations disagree
            } else if (sc.overflow != sc2.overflow) {
                return 1002;
            }
        } else if (mode == BGR_MODE_GREEDY_PATH) {
            flags = greedy_path_read(ix, a, r, sc, np);
        } else {
            return BGR_ERR_ARG;
        }
        if (flags & BGR_F_CNT_ALIGNED) cnt.c[BGR_C_ALIGNED]++;
        if (flags & BGR_F_CNT_NOTALIGNED) cnt.c[BGR_C_NOTALIGNED]++;
        if (flags & BGR_F_CNT_NOOVERLAP) cnt.c[BGR_C_NOOVERLAP]++;
        u32 plen = (flags & BGR_F_MAPPED) ? np : 0;
        if (sc.overflow) {
            *scratch_overflow = 1;
            plen = 0;
        }
        u64 offp = cnt.pool_cursor;
        cnt.pool_cursor += plen;
        if (offp + plen > pool_cap) return BGR_ERR_LIMIT;
        for (u32 j = 0; j < plen; ++j) pool[offp + j] = (j < nl) ? sc.get(nl - 1 - j) : sc.get(j);
        res[r].flags_len = (flags & 0xFFu) | (plen << 8);
        res[r].path_off = (u32)offp;
    }
    *used = cnt.pool_cursor;
    for (int i = 0; i < BGR_C_COUNT; ++i) counters[i] = cnt.c[i] + scnt.c[i];
    return BGR_OK;
}

extern "C" int emu_map(const char* useqs, const uint64_t* uoff, uint32_t nu, uint32_t k, const char* seqs,
                       const uint64_t* off, uint32_t n, uint32_t mode, uint32_t errors, uint32_t effort,
                       uint32_t scratch_ints, bgr_result* res, int32_t* pool, uint64_t pool_cap, uint64_t* used,
                       uint64_t counters[BGR_C_COUNT], int* scratch_overflow) {
    return emu_map2(useqs, uoff, nu, k, seqs, off, n, mode, errors, effort, scratch_ints, res, pool, pool_cap, used,
                    counters, scratch_overflow, 0);
}

// the host packer alone (bgr_pack.cpp): words + flag bitmap for a batch
extern "C" uint32_t emu_pack_reads(const char* seqs, const uint64_t* off, uint32_t n, uint32_t threads, uint64_t* words,
                                   uint32_t* bad_bits) {
    return pack_reads_2bit(seqs, off, n, threads, words, bad_bits);
}

// ---- host-logic hooks: the product's own parser / unitig loader (bgr_host.cpp), for CPU tests

static char* dup_bytes(const char* p, size_t n) {
    char* q = (char*)malloc(n + 1);
    if (n) memcpy(q, p, n);
    q[n] = 0;
    return q;
}
static uint64_t* dup_u64(const uint64_t* p, size_t n) {
    uint64_t* q = (uint64_t*)malloc(n * sizeof(uint64_t));
    memcpy(q, p, n * sizeof(uint64_t));
    return q;
}

extern "C" long emu_parse_reads(const char* path, uint32_t k, int fastq, uint32_t chunk, char** headers,
                                uint64_t** hoff, char** seqs, uint64_t** soff) {
    MappedFile mf;
    if (!mf.open(path)) return -1;
    ReadParser rp(mf.data, mf.size, k, fastq != 0);
    ReadBatch rb;
    while (!rp.done()) rp.next_chunk(chunk, rb);
    *headers = dup_bytes(rb.headers.p, rb.headers.n);
    *seqs = dup_bytes(rb.seqs.p, rb.seqs.n);
    *hoff = dup_u64(rb.hoff.p, rb.hoff.n);
    *soff = dup_u64(rb.soff.p, rb.soff.n);
    return (long)rb.size();
}

// the parallel reader (bgr_host.cpp ParallelReader) over the whole file, range by range
extern "C" long emu_parse_reads_parallel(const char* path, uint32_t k, int fastq, uint32_t threads, uint64_t range_bytes,
                                         uint64_t min_slice, char** headers, uint64_t** hoff, char** seqs, uint64_t** soff) {
    MappedFile mf;
    if (!mf.open(path)) return -1;
    ParallelReader pr(mf.data, mf.size, k, fastq != 0, threads, (size_t)range_bytes, (size_t)min_slice);
    ReadBatch rb;
    ParsedBatch pb;
    do {
        pb.clear();
        pr.next(pb);
        for (size_t i = 0; i < pb.size(); ++i)
            rb.push(pb.hdr.p[i].p, pb.hdr.p[i].n, pb.seqs.p + pb.soff.p[i], pb.soff.p[i + 1] - pb.soff.p[i]);
    } while (!pr.done());
    *headers = dup_bytes(rb.headers.p, rb.headers.n);
    *seqs = dup_bytes(rb.seqs.p, rb.seqs.n);
    *hoff = dup_u64(rb.hoff.p, rb.hoff.n);
    *soff = dup_u64(rb.soff.p, rb.soff.n);
    return (long)rb.size();
}

extern "C" long emu_load_unitigs(const char* path, uint32_t k, char** seqs, uint64_t** off, int* opened) {
    std::string s;
    std::vector<uint64_t> o;
    *opened = load_unitig_file(path, k, s, o) ? 1 : 0;
    *seqs = dup_bytes(s.data(), s.size());
    *off = dup_u64(o.data(), o.size());
    return (long)o.size() - 1;
}

extern "C" long emu_load_unitigs_parallel(const char* path, uint32_t k, uint32_t threads, uint64_t min_slice, char** seqs,
                                          uint64_t** off, int* opened) {
    PodBuf<char> s;
    PodBuf<uint64_t> o;
    *opened = load_unitig_file_parallel(path, k, threads, s, o, (size_t)min_slice) ? 1 : 0;
    *seqs = dup_bytes(s.p, s.n);
    *off = dup_u64(o.p, o.n);
    return (long)o.n - 1;
}

// correct_read (bgr_kernels.cuh) for a batch of greedy paths over one index: paths are concatenated,
// path i = ints[poff[i] .. poff[i+1]); out receives lens[i] characters per read, concatenated; got[i] =
// bases written
extern "C" int emu_correct(const char* useqs, const uint64_t* uoff, uint32_t nu, uint32_t k, const int32_t* ints,
                           const uint64_t* poff, const uint32_t* lens, const int* rcs, uint32_t n, char* out,
                           uint32_t* got) {
    std::vector<uint64_t> blob;
    std::string err;
    int rc = build_index_blob(useqs, uoff, nu, k, blob, err);
    if (rc) return rc;
    BlobHeader h;
    memcpy(&h, blob.data(), sizeof h);
    char* B = (char*)blob.data();
    DevIndex ix;
    memset(&ix, 0, sizeof ix);
    ix.meta = (const UnitigMeta*)(B + h.off_meta);
    ix.pool = (const u64*)(B + h.off_pool);
    ix.k = h.k;
    size_t o = 0;
    for (uint32_t i = 0; i < n; ++i) {
        std::vector<u64> w((lens[i] + 31) / 32 + 2, 0xDEADBEEFDEADBEEFull);
        got[i] = correct_read(ix, ints + poff[i], (u32)(poff[i + 1] - poff[i]), lens[i], rcs[i] != 0, w.data());
        for (uint32_t j = 0; j < lens[i]; ++j) out[o + j] = "ACGT"[(w[j >> 5] >> (62 - 2 * (j & 31))) & 3];
        o += lens[i];
    }
    return 0;
}

extern "C" int emu_acgtn_only(const char* p, uint64_t n) { return acgtn_only(p, (size_t)n) ? 1 : 0; }

extern "C" void emu_free(void* p) { free(p); }
