// Index construction on the GPU: replaces Aligner::indexUnitigsAux
// (reference aligner.cpp:407-534) and the three BooPHF builds for graphs whose host-side build takes
// tens of seconds (3 Gbp: 60 M unitigs, 40 M border keys).
//
//   characters + offsets (uploaded once)
//     -> k_border_refs : per unitig, its two border (k-1)-mers, canonical key, table side, orientation
//                        flags (the same rules as build_index_blob, bgr_host.cpp)
//     -> stable radix sort by key (cub::DeviceRadixSort; a library primitive, not a hot-path kernel):
//        equal keys keep FILE ORDER, which is what the reference's slot rules depend on
//     -> k_mark_runs / exclusive sum / k_run_starts : distinct keys and their runs
//     -> k_insert_keys : bucketed open addressing by atomicCAS, slots of a bucket fill in order; Bloom bits
//     -> k_pack_pool / k_unitig_meta : 2-bit pool, ACGT validation, per-unitig (start, length)
//     -> k_fill_entries: per key, left/right slots in file order, the 5th and later arrivals overwrite
//                        slot 4 (aligner.cpp:481-489); every slot also gets where its unitig lies and the key
//                        slot of the unitig's other end (the next level of a walk)
//
// The blob it leaves in HBM has the layout of bgr_common.h and answers every query exactly like the
// host-built one (tests/test_gpu_parity.py compares the two); only the slot a key occupies inside its
// probe sequence may differ, which no result depends on.
#include <cuda_runtime.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <string>

#include "../../include/bgreat_b200.h"
#include "bgr_common.h"
#include "bgr_index_gpu.h"

namespace bgr {

namespace {

struct BuildStatus {
    unsigned long long bad_pos;  // smallest character position (in the uploaded text) that is not A/C/G/T
};

// str2num of n characters (n <= 31): A0 C1 G2 T3; anything else is reported through bad
__device__ __forceinline__ u64 code_of(const char* s, u32 n, bool& bad) {
    u64 x = 0;
    for (u32 j = 0; j < n; ++j) {
        const u32 c = base_code((unsigned char)s[j]);
        bad |= c > 3;
        x = (x << 2) | (c & 3);
    }
    return x;
}

// Two references per unitig, in file order: index 2i = the begin border, 2i+1 = the end border.
// flags: bit0 = right table, bits 30/31 = orientation flags of the slot (bgr_common.h).
__global__ void k_border_refs(const char* seq, const u64* off, u32 nu, u32 k, u64* keys, u32* idx, u32* flags) {
    const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nu) return;
    const u64 o = off[i] - off[0];
    const u32 len = (u32)(off[i + 1] - off[i]);
    bool bad = false;  // reported with its position by k_pack_pool
    const u64 beg = code_of(seq + o, k - 1, bad);
    const u64 end = code_of(seq + o + len - (k - 1), k - 1, bad);
    const u64 rcbeg = rcb(beg, k - 1), rcend = rcb(end, k - 1);
    u64 ka, kb;
    u32 fa, fb;
    if (beg <= rcbeg) {  // left[beg]
        ka = beg;
        fa = 0u | SLOT_FWD_CANON | (rcend == beg ? SLOT_FWD_RC : 0u);
    } else {  // right[rcbeg]
        ka = rcbeg;
        fa = 1u | SLOT_FWD_RC | (end == rcbeg ? SLOT_FWD_CANON : 0u);
    }
    if (end <= rcend) {  // right[end]
        kb = end;
        fb = 1u | SLOT_FWD_CANON | (rcbeg == end ? SLOT_FWD_RC : 0u);
    } else {  // left[rcend]
        kb = rcend;
        fb = 0u | SLOT_FWD_RC | (beg == rcend ? SLOT_FWD_CANON : 0u);
    }
    keys[2 * (u64)i] = ka;
    keys[2 * (u64)i + 1] = kb;
    idx[2 * (u64)i] = 2 * i;
    idx[2 * (u64)i + 1] = 2 * i + 1;
    flags[2 * (u64)i] = fa;
    flags[2 * (u64)i + 1] = fb;
}

__global__ void k_mark_runs(const u64* skeys, u64 n, u32* head) {
    const u64 j = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    head[j] = (j == 0 || skeys[j] != skeys[j - 1]) ? 1u : 0u;
}

// run r starts at sorted position run_start[r]; run_start[n_keys] = n
__global__ void k_run_starts(const u32* head, const u32* run_id, u64 n, u32 n_keys, u32* run_start) {
    const u64 j = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (j == 0) run_start[n_keys] = (u32)n;
    if (j >= n) return;
    if (head[j]) run_start[run_id[j]] = (u32)j;
}

__global__ void k_insert_keys(const u64* skeys, const u32* run_start, u32 n_keys, u64* tkeys, u64 n_buckets, u32* bloom,
                              u32 n_bloom, u32 bloom_mode, u32 km1, u32* slot_of) {
    const u32 r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_keys) return;
    const u64 key = skeys[run_start[r]];
    u64 b = bucket_of(key, n_buckets);
    for (;;) {
        bool placed = false;
        for (int j = 0; j < 4 && !placed; ++j) {
            // slots fill in order: a later slot is tried only when the earlier one is taken
            const unsigned long long old =
                atomicCAS((unsigned long long*)&tkeys[b * 4 + j], (unsigned long long)EMPTY_KEY, (unsigned long long)key);
            if (old == (unsigned long long)EMPTY_KEY) {
                slot_of[r] = (u32)(b * 4 + j);
                placed = true;
            }
        }
        if (placed) break;
        if (++b == n_buckets) b = 0;
    }
    if (bloom_mode == BLOOM_MINIMIZER) {
        u32 block;
        u64 pattern;
        bloom_slot_mm(key, n_bloom / 2, block, pattern);
        atomicOr((unsigned long long*)bloom + block, (unsigned long long)pattern);
    } else {
        u32 word, pattern;
        bloom_slot(key, km1, bloom_mode == BLOOM_WORDS_2HASH, n_bloom, word, pattern);
        atomicOr(&bloom[word], pattern);
    }
}

// slot of an existing key (every border of an accepted unitig is one)
__device__ __forceinline__ u32 find_key_slot(const u64* tkeys, u64 n_buckets, u64 key) {
    u64 b = bucket_of(key, n_buckets);
    for (u64 tries = 0; tries < n_buckets; ++tries) {
        for (int j = 0; j < 4; ++j)
            if (tkeys[b * 4 + j] == key) return (u32)(b * 4 + j);
        if (++b == n_buckets) b = 0;
    }
    return 0;  // unreachable for a well-formed table
}

__global__ void k_fill_entries(const u32* sidx, const u32* flags, const u32* run_start, const u32* slot_of, u32 n_keys,
                               const u64* off, const u64* pool, const u64* tkeys, u64 n_buckets, u32 km1, Entry* entries) {
    const u32 r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_keys) return;
    Entry e;
    for (int j = 0; j < 4; ++j) {
        e.left[j].start_len = e.right[j].start_len = 0;
        e.left[j].id_flags = e.right[j].id_flags = 0;
        e.left[j].next = e.right[j].next = 0;
    }
    u32 nl = 0, nr = 0;
    for (u32 j = run_start[r]; j < run_start[r + 1]; ++j) {  // file order (the sort is stable)
        const u32 ix = sidx[j];
        const u32 f = flags[ix];
        const u32 id = ix / 2 + 1;
        const u64 start = off[id - 1] - off[0], len = off[id] - off[id - 1];
        const bool right = (f & 1u) != 0;
        USlot v;
        v.start_len = start | (len << SLOT_START_BITS);
        v.id_flags = id | (f & (SLOT_FWD_CANON | SLOT_FWD_RC));
        // the entry the walk goes on with: the key slot of the (k-1)-mer at the unitig's other end
        const u64 c = slot_far_code(pool, start, len, km1, v.id_flags, right);
        const u64 rc = rcb(c, km1);
        v.next = find_key_slot(tkeys, n_buckets, c <= rc ? c : rc) | (c <= rc ? SLOT_NEXT_CANON : 0u) |
                 (c == rc ? SLOT_NEXT_PAL : 0u);
        if (right) {
            e.right[nr < 3 ? nr : 3] = v;  // first free of 4; the 5th and later overwrite slot 4
            ++nr;
        } else {
            e.left[nl < 3 ? nl : 3] = v;
            ++nl;
        }
    }
    entries[slot_of[r]] = e;
}

// pool word w = characters [32w, 32w+32) of the accepted text, first base most significant
__global__ void k_pack_pool(const char* seq, u64 total_bases, u64* pool, u64 pool_words, BuildStatus* st) {
    const u64 w = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= pool_words) return;
    u64 x = 0;
    const u64 c0 = w * 32;
    if (c0 < total_bases) {
        const u32 m = (u32)min((u64)32, total_bases - c0);
        // the text is uploaded to a 16-byte aligned buffer and c0 is a multiple of 32
        const uint4* q = (const uint4*)(seq + c0);
        uint4 v[2];
        v[0] = q[0];
        v[1] = (m > 16) ? q[1] : make_uint4(0, 0, 0, 0);
        const unsigned char* ch = (const unsigned char*)v;
        for (u32 j = 0; j < m; ++j) {
            const u32 c = base_code(ch[j]);
            if (c > 3) atomicMin(&st->bad_pos, (unsigned long long)(c0 + j));
            x |= (u64)(c & 3) << (62 - 2 * j);
        }
    }
    pool[w] = x;
}

__global__ void k_unitig_meta(const u64* off, u32 nu, UnitigMeta* meta) {
    const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i > nu) return;
    UnitigMeta m;
    m.pad = 0;
    if (i == 0) {
        m.start = 0;
        m.len = 0;
    } else {
        m.start = off[i - 1] - off[0];
        m.len = (u32)(off[i] - off[i - 1]);
    }
    meta[i] = m;
}

struct DevBuf {
    void* p = nullptr;
    ~DevBuf() {
        if (p) cudaFree(p);
    }
    cudaError_t alloc(size_t bytes) { return cudaMalloc(&p, bytes ? bytes : 16); }
    template <typename T>
    T* as() {
        return (T*)p;
    }
};

#define GCK(call)                                                        \
    do {                                                                 \
        cudaError_t e_ = (call);                                         \
        if (e_ != cudaSuccess) {                                         \
            err = std::string(#call) + ": " + cudaGetErrorString(e_);    \
            return BGR_ERR_CUDA;                                         \
        }                                                                \
    } while (0)

inline unsigned grid_for(u64 n, unsigned block) { return (unsigned)std::max<u64>(1, (n + block - 1) / block); }

}  // namespace

int build_index_on_device(const char* seqs, const uint64_t* off, uint32_t n, uint32_t k, void** d_blob_out,
                          BlobHeader* hdr_out, uint64_t* launches, std::string& err) {
    *d_blob_out = nullptr;
    if (k < 2 || k > 32) {
        err = "k must be in [2,32] (kmer is a 64-bit word in the reference, utils.h:27)";
        return BGR_ERR_ARG;
    }
    // accepted unitigs: up to the first one shorter than k (aligner.cpp:418)
    u64 nu = 0;
    for (; nu < n; ++nu) {
        const u64 len = off[nu + 1] - off[nu];
        if (len < k) break;
        if (len >> SLOT_LEN_BITS) {
            err = "unitig longer than 2^28-1 bases";
            return BGR_ERR_LIMIT;
        }
    }
    if (nu >= SLOT_ID_MASK) {
        err = "more than 2^30-1 unitigs";
        return BGR_ERR_LIMIT;
    }
    const u64 total_bases = nu ? off[nu] - off[0] : 0;
    if (total_bases >> SLOT_START_BITS) {
        err = "more than 2^36 unitig bases";
        return BGR_ERR_LIMIT;
    }
    const u64 pool_words = (total_bases + 31) / 32 + 2;
    const u64 n_refs = 2 * nu;
    if (n_refs >= 0xFFFFFFFFull) {
        err = "too many unitigs for 32-bit reference indices";
        return BGR_ERR_LIMIT;
    }
    const unsigned B = 256;

    // ---- upload the accepted text and its offsets
    DevBuf d_seq, d_off, d_st;
    GCK(d_seq.alloc(total_bases + 64));
    GCK(d_off.alloc((nu + 1) * sizeof(u64)));
    GCK(d_st.alloc(sizeof(BuildStatus)));
    if (total_bases) GCK(cudaMemcpy(d_seq.p, seqs + off[0], total_bases, cudaMemcpyHostToDevice));
    GCK(cudaMemcpy(d_off.p, off, (nu + 1) * sizeof(u64), cudaMemcpyHostToDevice));
    GCK(cudaMemset(d_st.p, 0xFF, sizeof(BuildStatus)));

    // ---- border references, sorted by key (stable: file order within a key)
    DevBuf d_keys, d_skeys, d_idx, d_sidx, d_flags, d_head, d_runid, d_tmp;
    GCK(d_keys.alloc(n_refs * sizeof(u64)));
    GCK(d_skeys.alloc(n_refs * sizeof(u64)));
    GCK(d_idx.alloc(n_refs * sizeof(u32)));
    GCK(d_sidx.alloc(n_refs * sizeof(u32)));
    GCK(d_flags.alloc(n_refs * sizeof(u32)));
    u32 n_keys = 0;
    DevBuf d_runstart, d_slot;
    if (nu) {
        k_border_refs<<<grid_for(nu, B), B>>>(d_seq.as<char>(), d_off.as<u64>(), (u32)nu, k, d_keys.as<u64>(),
                                               d_idx.as<u32>(), d_flags.as<u32>());
        ++*launches;
        size_t tmp_bytes = 0;
        const int end_bit = (int)std::min<u32>(64, 2 * (k - 1));
        GCK(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, d_keys.as<u64>(), d_skeys.as<u64>(), d_idx.as<u32>(),
                                            d_sidx.as<u32>(), (int)n_refs, 0, end_bit));
        size_t scan_bytes = 0;
        GCK(cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, (u32*)nullptr, (u32*)nullptr, (int)n_refs));
        GCK(d_tmp.alloc(std::max(tmp_bytes, scan_bytes)));
        GCK(cub::DeviceRadixSort::SortPairs(d_tmp.p, tmp_bytes, d_keys.as<u64>(), d_skeys.as<u64>(), d_idx.as<u32>(),
                                            d_sidx.as<u32>(), (int)n_refs, 0, end_bit));
        // ---- distinct keys
        GCK(d_head.alloc(n_refs * sizeof(u32)));
        GCK(d_runid.alloc(n_refs * sizeof(u32)));
        k_mark_runs<<<grid_for(n_refs, B), B>>>(d_skeys.as<u64>(), n_refs, d_head.as<u32>());
        ++*launches;
        GCK(cub::DeviceScan::ExclusiveSum(d_tmp.p, scan_bytes, d_head.as<u32>(), d_runid.as<u32>(), (int)n_refs));
        u32 last_id = 0, last_head = 0;
        GCK(cudaMemcpy(&last_id, d_runid.as<u32>() + (n_refs - 1), sizeof(u32), cudaMemcpyDeviceToHost));
        GCK(cudaMemcpy(&last_head, d_head.as<u32>() + (n_refs - 1), sizeof(u32), cudaMemcpyDeviceToHost));
        n_keys = last_id + last_head;
        GCK(d_runstart.alloc(((size_t)n_keys + 1) * sizeof(u32)));
        GCK(d_slot.alloc((size_t)n_keys * sizeof(u32)));
        k_run_starts<<<grid_for(n_refs, B), B>>>(d_head.as<u32>(), d_runid.as<u32>(), n_refs, n_keys, d_runstart.as<u32>());
        ++*launches;
    }
    const u64 n_buckets = std::max<u64>(4, ((u64)n_keys + 1) / 2);  // <= 50 % load
    if (n_buckets * 4 > SLOT_NEXT_MASK) {
        err = "too many distinct border k-mers for 30-bit key slots";
        return BGR_ERR_LIMIT;
    }

    // ---- blob layout (identical to build_index_blob)
    BlobHeader h;
    memset(&h, 0, sizeof h);
    h.magic = BLOB_MAGIC;
    h.version = BLOB_VERSION;
    h.k = k;
    h.n_unitigs = nu;
    h.n_keys = n_keys;
    h.n_buckets = n_buckets;
    h.pool_words = pool_words;
    auto align_up = [](u64 x) { return (x + 255) & ~255ull; };
    h.off_keys = align_up(sizeof(BlobHeader));
    h.off_entries = align_up(h.off_keys + n_buckets * 4 * sizeof(u64));
    h.off_meta = align_up(h.off_entries + n_buckets * 4 * sizeof(Entry));
    h.off_pool = align_up(h.off_meta + (nu + 1) * sizeof(UnitigMeta));
    u32 bloom_mode = 0;
    h.n_bloom = bloom_layout(n_keys, k, bloom_mode);
    h.bloom_mode = bloom_mode;
    if (h.n_bloom >= 0xFFFFFFFFull) {
        err = "too many distinct border k-mers for the 32-bit filter hash";
        return BGR_ERR_LIMIT;
    }
    h.off_bloom = align_up(h.off_pool + pool_words * sizeof(u64));
    h.total_bytes = align_up(h.off_bloom + h.n_bloom * sizeof(u32));

    DevBuf blob;
    GCK(blob.alloc(h.total_bytes));
    char* Bp = blob.as<char>();
    GCK(cudaMemset(Bp, 0, h.total_bytes));
    GCK(cudaMemset(Bp + h.off_keys, 0xFF, n_buckets * 4 * sizeof(u64)));  // EMPTY_KEY
    GCK(cudaMemcpy(Bp, &h, sizeof h, cudaMemcpyHostToDevice));
    if (n_keys) {
        k_insert_keys<<<grid_for(n_keys, B), B>>>(d_skeys.as<u64>(), d_runstart.as<u32>(), n_keys, (u64*)(Bp + h.off_keys),
                                                   n_buckets, (u32*)(Bp + h.off_bloom), (u32)h.n_bloom, h.bloom_mode,
                                                   k - 1, d_slot.as<u32>());
        ++*launches;
    }
    k_pack_pool<<<grid_for(pool_words, B), B>>>(d_seq.as<char>(), total_bases, (u64*)(Bp + h.off_pool), pool_words,
                                                 d_st.as<BuildStatus>());
    k_unitig_meta<<<grid_for(nu + 1, B), B>>>(d_off.as<u64>(), (u32)nu, (UnitigMeta*)(Bp + h.off_meta));
    *launches += 2;
    if (n_keys) {  // needs the finished key table and the packed pool (far-end codes)
        k_fill_entries<<<grid_for(n_keys, B), B>>>(d_sidx.as<u32>(), d_flags.as<u32>(), d_runstart.as<u32>(),
                                                    d_slot.as<u32>(), n_keys, d_off.as<u64>(), (const u64*)(Bp + h.off_pool),
                                                    (const u64*)(Bp + h.off_keys), n_buckets, k - 1,
                                                    (Entry*)(Bp + h.off_entries));
        ++*launches;
    }
    GCK(cudaGetLastError());
    BuildStatus st;
    GCK(cudaMemcpy(&st, d_st.p, sizeof st, cudaMemcpyDeviceToHost));  // synchronises with the kernels
    if (st.bad_pos != ~0ull) {
        const u64 pos = off[0] + st.bad_pos;
        const u64 i = (u64)(std::upper_bound(off, off + nu + 1, pos) - off) - 1;
        err = "unitig " + std::to_string(i + 1) + " contains a character outside ACGT";
        return BGR_ERR_FORMAT;
    }
    *hdr_out = h;
    *d_blob_out = blob.p;
    blob.p = nullptr;  // ownership passes to the caller
    return BGR_OK;
}

}  // namespace bgr
