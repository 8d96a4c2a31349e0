// Correction mode (recoverPath, aligner.cpp:268-290), the extension kernels around the per-read bodies of
// bgr_extend.cuh, path compaction in read order, k_correct, k_index_query. Part of bgr_kernels.cuh.
#pragma once
#include "bgr_dev.cuh"

namespace bgr {

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

// ---- greedy extension: lane per read for the walk's state, the whole warp for its comparisons.
//
// Every lane owns one read and runs GreedyState's stages (bgr_extend.cuh). What differs from lane to lane is how
// much comparing a level needs -- 1..4 candidates of 1..5 (or, for long reads, hundreds of) 32-base chunks -- and
// run lane by lane that left 9 of 32 threads active per instruction (profiles/r01_ncu_k_extend_greedy_v11_5Mbp.txt).
// So a turn of the warp is:
//   1. every walking lane reads its level's entry side (64 bytes) and lays out <= 4 comparison jobs in shared
//      memory, at job slots 4*lane .. 4*lane+3, with an exclusive prefix sum of their chunk counts;
//   2. the chunks of ALL jobs are dealt round-robin to the 32 lanes: chunk j finds its job by binary search in the
//      prefix sums, loads two pool words and two read words, XOR + popc, and adds to the job's counter;
//   3. every walking lane reads its counters and settles the level (GreedyState::level_settle).
// The loads of step 2 are all independent and issued together: a level is two dependent round trips.
// Measured (10 M reads, 6.4 M of them listed): 15 of 32 threads per instruction, 3.48 ms at 5 Mbp, 4.68 ms at
// 3 Gbp against 3.49 / 5.9 ms for thread-per-read (profiles/r02_ncu_k_extend_greedy_a_warp_*.txt). A CTA-wide
// form of the same stages -- read states in shared memory, per-round lists of the reads at each stage, so that
// every stage runs on dense warps -- reached 21.6 of 32 threads but lost the time to its seven barriers per
// round and to moving state in and out of shared memory: 5.0 / 6.6 ms (profiles/r02_ncu_k_extend_greedy_b_cta_*.txt).
struct ExtJob {   // 32 bytes
    u64 ua;       // first unitig base, forward pool coordinates
    u64 wofs;     // the read's first data word in BatchArgs::F / NM
    i32 sa, n;    // first read base (stored coordinates), bases
    u32 flags;    // bit0 flip, bit1 read has N, bit2 pass
    u32 pad;
};
struct ExtWarpShared {
    ExtJob job[128];
    u32 start[128];  // first chunk of job q among the warp's chunks of this turn
    u32 miss[128];
};

// Steps 1b and 2 of a turn, for both extension kernels: lane's jobs lv.job[0..nc) (q: the lane has any) go to shared
// memory with a prefix sum of their chunk counts, the chunks of all jobs are dealt to the 32 lanes, and every lane
// gets back the mismatch counts of its own jobs (exact: not capped).
__device__ __forceinline__ void warp_compare(ExtWarpShared& W, const DevIndex& ix, const BatchArgs& a, const u32 lane, const bool q,
                                             const int nc, const GreedyLevel& lv, const u64 wofs, const bool hasN, const bool pass,
                                             u32 miss[4]) {
    u32 cnt[4], mine = 0;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        cnt[c] = (q && c < nc) ? (u32)(lv.job[c].n + 31) >> 5 : 0u;
        mine += cnt[c];
    }
    u32 incl = mine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const u32 x = __shfl_up_sync(0xFFFFFFFFu, incl, d);
        if (lane >= (u32)d) incl += x;
    }
    const u32 total = __shfl_sync(0xFFFFFFFFu, incl, 31);
    {
        u32 at = incl - mine;
        const u32 jf = (hasN ? 2u : 0u) | (pass ? 4u : 0u);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const u32 qi = 4 * lane + c;
            W.start[qi] = at;
            W.miss[qi] = 0;
            if (cnt[c]) {
                ExtJob j;
                j.ua = lv.job[c].ua;
                j.wofs = wofs;
                j.sa = lv.job[c].sa;
                j.n = lv.job[c].n;
                j.flags = (lv.job[c].flip ? 1u : 0u) | jf;
                j.pad = 0;
                W.job[qi] = j;
            }
            at += cnt[c];
        }
    }
    __syncwarp();
    for (u32 j = lane; j < total; j += 32) {
        // its job: the last one whose first chunk is at or before j (a table chunk -> job, filled by the owners,
        // was measured and was no faster than these seven steps)
        u32 qi = 0;
#pragma unroll
        for (u32 sft = 64; sft; sft >>= 1)
            if (W.start[qi + sft] <= j) qi += sft;
        const ExtJob job = W.job[qi];
        const i32 done = (i32)(j - W.start[qi]) * 32;
        const u32 m = job_chunk_miss(ix.pool, a.F + job.wofs, a.NM + job.wofs, (job.flags & 2u) != 0, (job.flags & 4u) != 0,
                                     job.ua, job.sa, job.n, (job.flags & 1u) != 0, done);
        if (m) atomicAdd(&W.miss[qi], m);
    }
    __syncwarp();
#pragma unroll
    for (int c = 0; c < 4; ++c) miss[c] = W.miss[4 * lane + c];
    __syncwarp();  // the next turn overwrites the tables
}

template <int MINB>
__global__ void __launch_bounds__(EXT_THREADS, MINB) k_extend_greedy(DevIndex ix, BatchArgs a) {
    __shared__ __align__(16) ExtWarpShared wsh_all[EXT_THREADS / 32];
    ExtWarpShared& W = wsh_all[threadIdx.x >> 5];
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
    // Hand-in, refill and seed selection are needed once or twice per read while a walk has several levels:
    // they are serviced together, once `lazy` lanes wait for them (or nobody can walk).
    const u32 lazy = a.ext_lazy ? a.ext_lazy : 1;
    for (;;) {
        const bool walking = have && g.walking();
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
        // ---- 1. entry side -> comparison jobs
        const bool q = have && g.walking() && g.level_begin(ix, a);
        USlot sl[4];
        int nc = 0;
        GreedyLevel lv;
        lv.n = 0;
        lv.ends = 0;
        if (q) {
            if (const USlot* side = g.level_side(ix)) {
                load_side(side, sl);
                nc = count_slots(sl);
            }
            g.level_jobs(ix, sl, nc, lv);
        }
        u32 miss[4];
        warp_compare(W, ix, a, lane, q, nc, lv, g.s.wb + 1, g.s.v.hasN, g.s.v.pass, miss);
        // ---- 3. settle
        if (q) g.level_settle(a, sl, lv, miss);
    }
    if (any_overflow) atomicExch(&a.cnt->scratch_overflow, 1u);
    count_classes(a, n_al, n_na, n_no);
}

// alignReadExhaustive / alignReadExhaustivePath as persistent warps (same life cycle as k_extend_greedy): every lane
// runs one node of its read's search per turn, in the stages of ExhaustiveState, and the candidates of all the
// lanes' nodes are compared by the whole warp (warp_compare).
template <int MINB>
__global__ void __launch_bounds__(EXT_THREADS, MINB) k_extend_exhaustive_nodes(DevIndex ix, BatchArgs a) {
    __shared__ __align__(16) ExtWarpShared wsh_all[EXT_THREADS / 32];
    ExtWarpShared& W = wsh_all[threadIdx.x >> 5];
    const u32 lane = threadIdx.x & 31;
    const u32 list_n = a.scnt->list_n;
    __shared__ i32 xf_all[XF_INTS * EXT_THREADS];  // the lanes' list heads and parked frames (ExhaustiveState::fast)
    static_assert(XF_STRIDE == (u32)EXT_THREADS, "a lane's entries are one CTA width apart");
    ExhaustiveState g;
    g.phase = XP_DONE;
    g.sc.overflow = false;
    g.fast = xf_all + threadIdx.x;
    bool have = false;
    u32 r = 0, t = 0;
    u32 chunk_base = 0, chunk_left = 0;
    bool dry = false;
    u32 n_al = 0, n_na = 0;
    bool any_overflow = false;
    const u32 lazy = a.ext_lazy ? a.ext_lazy : 1;
    for (;;) {
        const bool searching = have && g.searching();
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
        // ---- one node: entry side -> comparison jobs -> warp-wide comparing -> settle
        const bool q = have && g.searching() && g.node_begin(ix, a);
        USlot sl[4];
        int nc = 0;
        GreedyLevel lv;
        lv.n = 0;
        lv.ends = 0;
        if (q) {
            if (const USlot* side = g.node_side(ix)) {
                load_side(side, sl);
                nc = count_slots(sl);
            }
            g.node_jobs(ix, sl, nc, lv);
        }
        u32 miss[4];
        warp_compare(W, ix, a, lane, q, nc, lv, g.s.wb + 1, g.s.v.hasN, g.s.v.pass, miss);
        if (q) g.node_settle(ix, a, sl, lv, miss);
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
