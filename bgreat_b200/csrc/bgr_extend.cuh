// Overlap extension: what every aligner shares (read views, index queries, mismatch counting), the greedy
// aligner (alignerGreedy.cpp:35-57,167-364), the exhaustive aligner (alignerExhaustive.cpp:35-259) and the
// greedy "paths" aligner (alignerPaths.cpp). Per-read bodies only; the kernels are in bgr_emit.cuh.
// Part of bgr_kernels.cuh.
#pragma once
#include "bgr_dev.cuh"

namespace bgr {

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

// The candidates of one border: the <= 4 slots of one side of a key's entry (bgr_common.h), scanned in order up
// to the first empty one (aligner.cpp:172-203).
struct Cands {
    USlot s[4];
    int n;
    u32 fwd_bit;
    bool canon;  // the query's code was the canonical form of its key
};

static const u32 NO_SLOT = 0xFFFFFFFFu;

// one side of an entry: 64 bytes, two sectors of one line, one request
BGR_HD void load_side(const USlot* side, USlot out[4]) {
#if defined(__CUDA_ARCH__)
    const u64x4 a = ld_bucket((const u64*)side), b = ld_bucket((const u64*)side + 4);
    out[0].start_len = a.a;
    out[0].id_flags = (u32)a.b;
    out[0].next = (u32)(a.b >> 32);
    out[1].start_len = a.c;
    out[1].id_flags = (u32)a.d;
    out[1].next = (u32)(a.d >> 32);
    out[2].start_len = b.a;
    out[2].id_flags = (u32)b.b;
    out[2].next = (u32)(b.b >> 32);
    out[3].start_len = b.c;
    out[3].id_flags = (u32)b.d;
    out[3].next = (u32)(b.d >> 32);
#else
    for (int j = 0; j < 4; ++j) out[j] = side[j];
#endif
}

BGR_HD int count_slots(const USlot s[4]) {
    const bool s0 = (s[0].id_flags & SLOT_ID_MASK) != 0, s1 = s0 && (s[1].id_flags & SLOT_ID_MASK) != 0,
               s2 = s1 && (s[2].id_flags & SLOT_ID_MASK) != 0, s3 = s2 && (s[3].id_flags & SLOT_ID_MASK) != 0;
    return (int)s0 + (int)s1 + (int)s2 + (int)s3;
}

// getBegin / getEnd for the key in key slot `slot`, queried through its canonical form (canon) or through the
// reverse complement of that: getEnd takes the right table for a canonical code, else the left one; getBegin the
// other way round (aligner.cpp:147-267).
BGR_HD const USlot* entry_side(const DevIndex& ix, u32 slot, bool canon, bool want_end) {
    const bool right = (want_end == canon);
    return right ? ix.entries[slot].right : ix.entries[slot].left;
}
BGR_HD Cands index_query_at(const DevIndex& ix, u32 slot, bool canon, bool want_end) {
    Cands c;
    c.n = 0;
    c.canon = canon;
    c.fwd_bit = canon ? SLOT_FWD_CANON : SLOT_FWD_RC;
    if (slot == NO_SLOT) return c;
    load_side(entry_side(ix, slot, canon, want_end), c.s);
    c.n = count_slots(c.s);
    return c;
}
// the same for a (k-1)-mer code: canonical form -> key table -> entry
BGR_HD Cands index_query(const DevIndex& ix, u64 bin, bool want_end) {
    const u64 rc = rcb(bin, ix.k - 1);
    const bool canon = bin <= rc;
    const i64 slot = index_find(ix, canon ? bin : rc);
    return index_query_at(ix, slot < 0 ? NO_SLOT : (u32)slot, canon, want_end);
}

BGR_HD Oriented orient_slot(const USlot& s, u32 fwd_bit) {
    Oriented u;
    const u32 id = s.id_flags & SLOT_ID_MASK;
    u.start = s.start_len & SLOT_START_MASK;
    u.len = (i64)(s.start_len >> SLOT_START_BITS);
    u.rc = (s.id_flags & fwd_bit) == 0;
    u.sid = u.rc ? -(i32)id : (i32)id;
    return u;
}
BGR_HD Oriented orient(const DevIndex&, const Cands& c, int i) { return orient_slot(c.s[i], c.fwd_bit); }

// where a walk goes on after stepping through slot s, which it reached by a query of form `canon`: the key
// slot of the unitig's other end, and the form of the next query (bgr_common.h, USlot::next)
BGR_HD void next_query(const USlot& s, bool canon, u32& slot, bool& next_canon) {
    slot = s.next & SLOT_NEXT_MASK;
    const bool c1 = (s.next & SLOT_NEXT_CANON) != 0, pal = (s.next & SLOT_NEXT_PAL) != 0;
    next_canon = canon ? c1 : (!c1 || pal);
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

// One missmatchNumber call laid out for chunk-wise evaluation: n bases of the oriented unitig from upos against
// n bases of the pass's read from rpos, in forward pool coordinates and stored read coordinates. The greedy
// kernel spreads the 32-base chunks of all the comparisons a warp has pending over its lanes (k_extend_greedy);
// the sum of the chunk counts, capped at cap+1, is what mismatches() returns.
struct CmpJob {
    u64 ua;     // first unitig base, forward pool coordinates
    i32 sa;     // first read base, stored coordinates
    i32 n;      // bases
    bool flip;  // the read segment is reverse-complemented before comparing
};
BGR_HD CmpJob make_job(const ReadView& v, const Oriented& u, i64 upos, i32 rpos, i32 n) {
    CmpJob j;
    j.ua = (u64)((i64)u.start + (u.rc ? u.len - upos - n : upos));
    j.sa = v.pass ? v.len - rpos - n : rpos;
    j.n = n;
    j.flip = u.rc != v.pass;
    return j;
}
// mismatches among bases [done, done+32) of the job (fewer at its end)
BGR_HD u32 job_chunk_miss(const u64* pool, const u64* D, const u64* ND, bool hasN, bool pass, u64 ua, i32 sa, i32 n,
                          bool flip, i32 done) {
    const i32 left = n - done;
    const i32 m = left < 32 ? left : 32;
    const i32 rp = flip ? sa + n - done - m : sa + done;
    const u64 xu = ld_bases(pool, (i64)ua + done);
    u64 y = ld_bases(D, rp), nm = 0;
    if (hasN) {
        const u64 ymask = ld_bases(ND, rp);
        if (pass) y |= ymask;  // pass 1 sequence is F0|NM, without a mask
        else nm = ymask;       // pass 0 carries the N mask
    }
    if (flip) {
        y = rc64(y) << ((64 - 2 * m) & 63);
        if (hasN) nm = rev2(nm) << ((64 - 2 * m) & 63);
    }
    u64 d = xu ^ y;
    d = (d | (d >> 1) | nm) & 0x5555555555555555ull;
    if (m < 32) d &= ~(~0ull >> (2 * m));
    return (u32)BGR_POPCLL(d);
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

// The same walk for reads of at most 128 windows whose hit mask sits in two registers (the usual read: 150 bases,
// 121 windows): no loop, no load. lo = windows 0..63, hi = 64..127; cur / lim / nwin / down as in SeedIter.
// In the greedy kernel seed selection is run for a few lanes at a time, and SeedIter's loop -- one trip per empty
// mask word, a load each -- ran with 5 of 32 lanes on average (profiles/r02_lines_k_extend_greedy_5Mbp.txt).
BGR_HD i32 seed_next128(u64 lo, u64 hi, i32 nwin, i32 lim, bool down, i32& cur) {
    if (cur >= lim) return -1;
    i32 p;
    if (!down) {
        // first set bit at or after cur
        const u32 s = (u32)cur;
        const u64 a = s >= 64 ? (hi >> (s & 63)) : ((lo >> s) | (s ? hi << (64 - s) : 0ull));
        const u64 b = s >= 64 ? 0ull : (hi >> s);
        if ((a | b) == 0) {
            cur = lim;
            return -1;
        }
#if defined(__CUDA_ARCH__)
        const i32 z = a ? __ffsll((long long)a) - 1 : 64 + __ffsll((long long)b) - 1;
#else
        const i32 z = a ? __builtin_ctzll(a) : 64 + __builtin_ctzll(b);
#endif
        p = cur + z;
    } else {
        // highest set bit at or below j = nwin-1-cur, as a distance from j
        const u32 s = 127u - (u32)(nwin - 1 - cur);  // left shift that brings bit j to bit 127
        const u64 a = s >= 64 ? (lo << (s & 63)) : ((hi << s) | (s ? lo >> (64 - s) : 0ull));
        const u64 b = s >= 64 ? 0ull : (lo << s);
        if ((a | b) == 0) {
            cur = lim;
            return -1;
        }
#if defined(__CUDA_ARCH__)
        const i32 z = a ? __clzll((long long)a) : 64 + __clzll((long long)b);
#else
        const i32 z = a ? __builtin_clzll(a) : 64 + __builtin_clzll(b);
#endif
        p = cur + z;
    }
    if (p >= lim) {
        cur = lim;
        return -1;
    }
    cur = p + 1;
    return p;
}

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
// A resumable state machine whose turn is one LEVEL of a walk, cut into three stages so that the kernel can run
// the middle one cooperatively:
//   level_begin : ends of walks that need no index query; says which entry side the level reads
//   level_jobs  : the <= 4 candidates of that side -> one comparison job each (geometry only, no memory access
//                 beyond the 64 bytes of the side: a slot carries its unitig's position and length)
//   level_settle: given the mismatch counts, the reference's choice -- a zero-miss candidate is taken at once,
//                 otherwise the first strictly smallest within the budget (quirk 9) -- and the next level's key
//                 slot straight from the chosen slot (USlot::next): no hashing, no load of the overlap code
// A level therefore costs two dependent round trips (entry side, then the words of all its comparisons at once)
// where the first form of this walk took four (bucket + payload, unitig meta, words, next overlap code).
// The walkers greedy_left / greedy_right above are the readable statement of the same procedure (they go through
// the key table at every level, which also cross-checks the next-slot links) and are kept for the emulator.
enum { GP_SEED = 0, GP_LEFT = 1, GP_RIGHT_FIRST = 2, GP_RIGHT_MORE = 3, GP_DONE = 4 };

struct GreedyLevel {
    CmpJob job[4];
    i32 L[4];         // bases of the candidate beyond the overlap
    u32 ends;         // bit i: candidate i covers the rest of the read
    int n;
};

struct GreedyState {
    ReadSetup s;
    Scratch sc;
    SeedIter it;
    u32 flags, n, nl, tried, phase;
    u64 hm_lo, hm_hi;  // the pass's hit mask, when it fits (hm_ok): seed_next128 instead of the iterator's loop
    bool hm_ok;
    // walk state: the key slot and query form of the current level, and of the seed
    u32 slot, seed_slot;
    bool canon, seed_canon;
    i32 pos, seed_pos;
    u32 budget, acc, eb;

    BGR_HD bool done() const { return phase == GP_DONE; }
    BGR_HD bool walking() const { return phase == GP_LEFT || phase == GP_RIGHT_FIRST || phase == GP_RIGHT_MORE; }

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
        // reads without N use the same mask for both passes (pass 1 walks it from the top)
        hm_ok = !s.v.hasN && s.nwin <= 128;
        hm_lo = hm_hi = 0;
        if (hm_ok && (s.info & INFO_HIT1)) {
            const u32* h = a.H1 + s.wb;
            const i32 nw32 = (s.nwin + 31) >> 5;
            const u64 w0 = nw32 > 0 ? h[0] : 0u, w1 = nw32 > 1 ? h[1] : 0u, w2 = nw32 > 2 ? h[2] : 0u, w3 = nw32 > 3 ? h[3] : 0u;
            hm_lo = w0 | (w1 << 32);
            hm_hi = w2 | (w3 << 32);
        }
        slot = seed_slot = NO_SLOT;
        canon = seed_canon = false;
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
        slot = seed_slot;
        canon = seed_canon;
        pos = seed_pos;
    }

    // (Handing the first seed over from the tile scan -- window, key slot and query form of a read's first border,
    // collected per tile with a shared-memory atomicMin -- was measured: the extension gained 0.1 ms per 10 M reads
    // and the scan, at its 64-register cap, lost 0.2-0.35 ms.)
    // Seed selection, possibly switching to the reverse-complement pass: needed once per seed, i.e. rarely
    // compared with the levels of a walk, so the kernel runs it for several lanes at once (see k_extend_greedy).
    // The seed's (k-1)-mer is looked up in the key table here; every later level follows the slots' links.
    BGR_HD void pick_seed(const DevIndex& ix, const BatchArgs& a) {
        ReadView& v = s.v;
        const u32 max_seeds = a.effort ? a.effort : 1;
        for (int rep = 0; rep < 2 && phase == GP_SEED; ++rep) {
            i32 p = -1;
            if (tried < max_seeds) p = hm_ok ? seed_next128(hm_lo, hm_hi, it.nwin, it.lim, it.down, it.cur) : it.next();
            if (p >= 0) {
                ++tried;
                const u64 code = seed_code(ix, v, p);
                const u64 rc = rcb(code, ix.k - 1);
                seed_canon = canon = code <= rc;
                // (a read with N can name a window whose code, canonicalised the way getBegin / getEnd do it, is
                // no key: the walk then finds no candidate, like the reference's)
                const i64 sl = index_find(ix, canon ? code : rc);
                seed_slot = slot = sl < 0 ? NO_SLOT : (u32)sl;
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

    // ---- stage 1: ends of walks that need no index query. Returns true when the level goes on to a query.
    BGR_HD bool level_begin(const DevIndex& ix, const BatchArgs& a) {
        const i32 km1 = (i32)ix.k - 1;
        if (phase == GP_LEFT && pos == 0) {  // checkBeginGreedy:269 / mapOnLeftEndGreedy:168
            sc.set(n++, 0);
            eb = acc;
            go_right(a);
        }
        if (phase == GP_RIGHT_FIRST && s.v.len - (pos + km1) == 0) succeed();  // checkEndGreedy:324
        if (phase == GP_RIGHT_MORE && s.v.len - pos < (i32)ix.k) succeed();    // mapOnRightEndGreedy:223
        return phase != GP_DONE;
    }
    // the 64 bytes this level reads: getEnd on the way left, getBegin on the way right (null: no such key)
    BGR_HD const USlot* level_side(const DevIndex& ix) const {
        return slot == NO_SLOT ? nullptr : entry_side(ix, slot, canon, phase == GP_LEFT);
    }

    // ---- stage 2: comparison geometry of the level's candidates. One formula for the three kinds of level (lanes
    // of a warp are at different kinds; as three branches this stage ran with 10 of 32 lanes):
    //   left walk   rem = pos                 unitig covers the rest: unitig[L-pos, L) vs read[0, pos)
    //                                         else:                   unitig[0, L)     vs read[pos-L, pos)
    //   right first rem = len - (pos + k-1)   comparisons start after the anchor overlap: unitig[k-1, ..) vs read[pos+k-1, ..)
    //   right more  rem = len - pos           the remaining read starts AT the overlap; the whole unitig is compared,
    //                                         cut at the read end (quirk 8): unitig[0, ..) vs read[pos, ..)
    BGR_HD void level_jobs(const DevIndex& ix, const USlot sl[4], int nc, GreedyLevel& lv) const {
        const ReadView& v = s.v;
        const i32 km1 = (i32)ix.k - 1;
        const u32 fwd_bit = canon ? SLOT_FWD_CANON : SLOT_FWD_RC;
        const bool left = phase == GP_LEFT, first = phase == GP_RIGHT_FIRST, more = phase == GP_RIGHT_MORE;
        const i32 rem = left ? pos : (first ? v.len - pos - km1 : v.len - pos);
        lv.n = nc;
        lv.ends = 0;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            lv.job[i].n = 0;
            lv.L[i] = 0;
            if (i < nc) {
                const Oriented u = orient_slot(sl[i], fwd_bit);
                const i32 L = (i32)(u.len - km1);
                const bool ends_here = L >= rem;
                const i64 upos = left ? (ends_here ? (i64)(L - pos) : 0) : (first ? (i64)km1 : 0);
                const i32 rpos = left ? (ends_here ? 0 : pos - L) : (first ? pos + km1 : pos);
                const i32 whole = more ? (i32)(u.len < rem ? u.len : rem) : L;
                const i32 nn = ends_here ? rem : whole;
                lv.job[i] = make_job(v, u, upos, rpos, nn);
                lv.L[i] = L;
                lv.ends |= (ends_here ? 1u : 0u) << i;
            }
        }
    }

    // (Asking the next level's entry sides into L2 while the comparisons run -- prefetch.global.L2 through the links of
    // the candidates that do not end the walk, for single-candidate levels only or for all -- was measured at 3 Gbp:
    // 3.96 -> 4.06 ms either way, 100 Mbp 3.64 -> 3.75: the extra requests cost more than the shorter wait saves.)
    // ---- stage 3: settle the level. miss[i] = mismatches of candidate i (any value above the budget counts as
    // budget+1, the cap of missmatchNumber)
    BGR_HD void level_settle(const BatchArgs& a, const USlot sl[4], const GreedyLevel& lv, const u32 miss[4]) {
        // the reference's choice, without branches: first strictly smallest; a zero is taken at once (quirk 9)
        u32 best = budget + 1, b_idf = 0, b_next = 0;
        i32 bL = 0;
        int bi = -1;
        bool accepted = false;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const u32 m = miss[i] < budget + 1 ? miss[i] : budget + 1;
            const bool take = i < lv.n && !accepted && m < best;
            best = take ? m : best;
            bi = take ? i : bi;
            b_idf = take ? sl[i].id_flags : b_idf;
            b_next = take ? sl[i].next : b_next;
            bL = take ? lv.L[i] : bL;
            accepted = accepted || (take && m == 0);
        }
        if (bi < 0) {
            phase = GP_SEED;  // this seed failed
            return;
        }
        const u32 id = b_idf & SLOT_ID_MASK;
        const bool urc = (b_idf & (canon ? SLOT_FWD_CANON : SLOT_FWD_RC)) == 0;
        const bool ended = (lv.ends >> bi) & 1u;
        const bool left = phase == GP_LEFT;
        sc.set(n++, urc ? -(i32)id : (i32)id);
        // an exact candidate leaves the budget untouched, an inexact one pays for its mismatches
        if (ended) {
            if (left) {
                sc.set(n++, bL - pos);
                eb = acc + best;
                go_right(a);
            } else {
                succeed();
            }
        } else {
            acc += best;
            budget -= best;
            USlot bs;
            bs.start_len = 0;
            bs.id_flags = b_idf;
            bs.next = b_next;
            next_query(bs, canon, slot, canon);
            pos += left ? -bL : bL;
            phase = left ? (u32)GP_LEFT : (u32)GP_RIGHT_MORE;
        }
    }

    // One turn, start to finish, by one thread (emulator; the kernel runs the stages itself).
    BGR_HD void step(const DevIndex& ix, const BatchArgs& a) {
        if (phase == GP_SEED) pick_seed(ix, a);
        if (!walking() || !level_begin(ix, a)) return;
        USlot sl[4];
        int nc = 0;
        if (const USlot* side = level_side(ix)) {
            load_side(side, sl);
            nc = count_slots(sl);
        }
        GreedyLevel lv;
        level_jobs(ix, sl, nc, lv);
        u32 miss[4] = {0, 0, 0, 0};
        for (int i = 0; i < nc; ++i)
            for (i32 d = 0; d < lv.job[i].n; d += 32)
                miss[i] += job_chunk_miss(ix.pool, s.v.D, s.v.ND, s.v.hasN, s.v.pass, lv.job[i].ua, lv.job[i].sa, lv.job[i].n,
                                          lv.job[i].flip, d);
        level_settle(a, sl, lv, miss);
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

// shared-memory part of ExhaustiveState's lists (see there): ints per lane, and the distance between a lane's entries
static const u32 XF_LIST = 8, XF_FRAMES = 4, XF_INTS = 2 * XF_LIST + 4 * XF_FRAMES, XF_STRIDE = 128;

struct ExhaustiveState {
    ReadSetup s;
    Scratch sc;
    SeedIter it;
    DfsLayout lay;
    u32 flags, phase, n;
    // current seed: the key slot of its (k-1)-mer and the form of the query (entered through the canonical code
    // or its reverse complement); every node below it follows the slots' links (USlot::next) instead of hashing
    // the overlap code again
    u32 seed_slot;
    bool seed_canon;
    i32 seed_pos;
    u32 w, eb;  // ints of the final path written so far; mismatches spent on the left side
    // current search (one side)
    u32 budget, best_total, best_n;
    i32 term, depth;
    // top frame, kept in registers; the frames below it live in the scratch row (4 ints each: slot | canon bit,
    // pos, acc, next candidate; DfsLayout still reserves the five of the loop-shaped form)
    u32 f_slot;
    bool f_canon;
    i32 f_pos;
    u32 f_acc, f_next;
    // The first XF_LIST entries of the current / best id lists and the first XF_FRAMES parked frames live in shared
    // memory when the kernel provides it (fast: this lane's column, entry e at fast[e * XF_STRIDE]); deeper entries,
    // and everything in the host emulator (fast == nullptr), in the scratch row. In global memory these few ints
    // were a quarter of the kernel's instructions and a fifth of its stalls: copies of a handful of ids, with two
    // or three lanes of the warp active and an L2 round trip per int (profiles/r02b_lines_k_extend_exhaustive_*.txt).
    i32* fast;
    BGR_HD i32 cur_get(u32 j) const { return (fast && j < XF_LIST) ? fast[j * XF_STRIDE] : sc.get(lay.cur0 + j); }
    BGR_HD void cur_set(u32 j, i32 x) {
        if (fast && j < XF_LIST) fast[j * XF_STRIDE] = x;
        else sc.set(lay.cur0 + j, x);
    }
    BGR_HD i32 best_get(u32 j) const {
        return (fast && j < XF_LIST) ? fast[(XF_LIST + j) * XF_STRIDE] : sc.get(lay.best0 + j);
    }
    BGR_HD void best_set(u32 j, i32 x) {
        if (fast && j < XF_LIST) fast[(XF_LIST + j) * XF_STRIDE] = x;
        else sc.set(lay.best0 + j, x);
    }
    // frame d, field f (0: slot | canon bit, 1: pos, 2: acc, 3: next candidate)
    BGR_HD i32 frame_get(u32 d, u32 f) const {
        return (fast && d < XF_FRAMES) ? fast[(2 * XF_LIST + 4 * d + f) * XF_STRIDE] : sc.get(lay.fr0 + 5 * d + (f ? f + 1 : 0));
    }
    BGR_HD void frame_set(u32 d, u32 f, i32 x) {
        if (fast && d < XF_FRAMES) fast[(2 * XF_LIST + 4 * d + f) * XF_STRIDE] = x;
        else sc.set(lay.fr0 + 5 * d + (f ? f + 1 : 0), x);
    }

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
        seed_slot = NO_SLOT;
        seed_canon = false;
        seed_pos = 0;
        budget = best_total = best_n = 0;
        term = -1;
        depth = -1;
        f_slot = NO_SLOT;
        f_canon = false;
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
        f_slot = seed_slot;
        f_canon = seed_canon;
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
        {
            const u64 code = seed_code(ix, s.v, p);
            const u64 rc = rcb(code, ix.k - 1);
            seed_canon = code <= rc;
            const i64 sl = index_find(ix, seed_canon ? code : rc);
            seed_slot = sl < 0 ? NO_SLOT : (u32)sl;
        }
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
            const u32 d = (u32)depth;
            const u32 sc0 = (u32)frame_get(d, 0);
            f_slot = (sc0 & 0x7FFFFFFFu) == 0x7FFFFFFFu ? NO_SLOT : (sc0 & 0x7FFFFFFFu);
            f_canon = (sc0 >> 31) != 0;
            f_pos = frame_get(d, 1);
            f_acc = (u32)frame_get(d, 2);
            f_next = (u32)frame_get(d, 3);
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
            for (u32 j = 0; j < best_n; ++j) sc.set(w++, best_get(best_n - 1 - j));
            eb = best_total;
            begin_side(true, a.errors - eb);
        } else {
            for (u32 j = 0; j < best_n; ++j) sc.set(w++, best_get(j));
            sc.set(w++, term);
            n = w;
            flags |= BGR_F_MAPPED | BGR_F_CNT_ALIGNED;
            phase = XP_DONE;
        }
    }

    // One node of the search (the body of exhaustive_side's while loop), in the three stages of the greedy level so
    // that the kernel can compare the candidates of all its lanes' nodes together:
    //   node_begin : nodes that need no query (the read is used up: record the walk; or nothing below can beat the
    //                incumbent); false = the node is finished
    //   node_jobs  : one comparison job per candidate from f_next on (geometry only)
    //   node_settle: the loop over the candidates with their exact mismatch counts: a candidate that covers the
    //                rest of the read may become the incumbent, the first other one within reach is descended into
    // The comparisons of the candidates behind the one descended into are wasted (and made again when the search
    // comes back to this node): the price of comparing them side by side.
    BGR_HD bool searching() const { return phase == XP_LEFT || phase == XP_RIGHT; }
    BGR_HD bool node_begin(const DevIndex& ix, const BatchArgs& a) {
        const bool right = phase == XP_RIGHT;
        const i32 rl = right ? s.v.len - (f_pos + (i32)ix.k - 1) : f_pos;
        bool query = true;
        if (rl == 0) {
            if (f_acc < best_total) {
                best_total = f_acc;
                best_n = (u32)depth;
                term = right ? 0 : -1;
                for (u32 j = 0; j < best_n; ++j) best_set(j, cur_get(j));
            }
            query = false;
        } else if (f_acc >= best_total) {
            query = false;  // nothing below can be strictly better
        }
        if (!query) {
            pop_frame();
            if (depth < 0) side_finished(a);
        }
        return query;
    }
    BGR_HD const USlot* node_side(const DevIndex& ix) const {
        return f_slot == NO_SLOT ? nullptr : entry_side(ix, f_slot, f_canon, phase != XP_RIGHT);
    }
    BGR_HD void node_jobs(const DevIndex& ix, const USlot sl[4], int nc, GreedyLevel& lv) const {
        const ReadView& v = s.v;
        const i32 km1 = (i32)ix.k - 1;
        const bool right = phase == XP_RIGHT;
        const u32 fwd_bit = f_canon ? SLOT_FWD_CANON : SLOT_FWD_RC;
        const i32 rl = right ? v.len - (f_pos + km1) : f_pos;
        lv.n = nc;
        lv.ends = 0;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            lv.job[i].n = 0;
            lv.L[i] = 0;
            if (i < nc && (u32)i >= f_next) {
                const Oriented u = orient_slot(sl[i], fwd_bit);
                const i32 L = (i32)(u.len - km1);
                const bool ends_here = L >= rl;
                // right: unitig[k-1, ..) vs read[pos+k-1, ..); left: unitig[L-rl, L) vs read[0, rl) when it covers the
                // rest, else unitig[0, L) vs read[rl-L, rl)
                const i64 upos = right ? (i64)km1 : (ends_here ? (i64)(L - rl) : 0);
                const i32 rpos = right ? f_pos + km1 : (ends_here ? 0 : rl - L);
                lv.job[i] = make_job(v, u, upos, rpos, ends_here ? rl : L);
                lv.L[i] = L;
                lv.ends |= (ends_here ? 1u : 0u) << i;
            }
        }
    }
    BGR_HD void node_settle(const DevIndex& ix, const BatchArgs& a, const USlot sl[4], const GreedyLevel& lv, const u32 miss[4]) {
        const bool right = phase == XP_RIGHT;
        const u32 fwd_bit = f_canon ? SLOT_FWD_CANON : SLOT_FWD_RC;
        const i32 km1 = (i32)ix.k - 1;
        const i32 rl = right ? s.v.len - (f_pos + km1) : f_pos;
        bool descended = false;
        for (u32 i = f_next; i < (u32)lv.n; ++i) {
            if (f_acc >= best_total) break;  // nothing below can be strictly better
            const u32 cap = best_total - 1 - f_acc;
            const Oriented u = orient_slot(sl[i], fwd_bit);
            const i32 L = lv.L[i];
            if ((lv.ends >> i) & 1u) {
                if (miss[i] <= cap) {
                    best_total = f_acc + miss[i];
                    best_n = (u32)depth + 1;
                    term = right ? rl + km1 : L - rl;
                    for (u32 j = 0; j < (u32)depth; ++j) best_set(j, cur_get(j));
                    best_set((u32)depth, u.sid);
                }
            } else if (miss[i] <= cap) {
                if ((u32)depth + 1 >= lay.D) {  // search stack too small: give up, the host grows the scratch
                    sc.overflow = true;
                    depth = -1;
                    side_finished(a);
                    return;
                }
                // park this frame (resume with candidate i+1), make the child the top frame
                const u32 d = (u32)depth;
                frame_set(d, 0, (i32)((f_slot & 0x7FFFFFFFu) | (f_canon ? 0x80000000u : 0u)));
                frame_set(d, 1, f_pos);
                frame_set(d, 2, (i32)f_acc);
                frame_set(d, 3, (i32)(i + 1));
                cur_set(d, u.sid);
                next_query(sl[i], f_canon, f_slot, f_canon);
                f_pos = right ? f_pos + L : f_pos - L;
                f_acc += miss[i];
                f_next = 0;
                ++depth;
                descended = true;
                break;
            }
        }
        if (!descended) pop_frame();
        if (depth < 0) side_finished(a);
    }

    // one node, start to finish, by one thread (emulator; the kernel runs the stages itself)
    BGR_HD void step_node(const DevIndex& ix, const BatchArgs& a) {
        if (!node_begin(ix, a)) return;
        USlot sl[4];
        int nc = 0;
        if (const USlot* side = node_side(ix)) {
            load_side(side, sl);
            nc = count_slots(sl);
        }
        GreedyLevel lv;
        node_jobs(ix, sl, nc, lv);
        u32 miss[4] = {0, 0, 0, 0};
        for (int i = 0; i < nc; ++i)
            for (i32 d = 0; d < lv.job[i].n; d += 32)
                miss[i] += job_chunk_miss(ix.pool, s.v.D, s.v.ND, s.v.hasN, s.v.pass, lv.job[i].ua, lv.job[i].sa, lv.job[i].n,
                                          lv.job[i].flip, d);
        node_settle(ix, a, sl, lv, miss);
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

}  // namespace bgr
