// C ABI implementation (include/bgreat_b200.h): context, index residency, batch slots, launches.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <cub/device/device_scan.cuh>
#include <string>
#include <thread>
#include <vector>

#include "../../include/bgreat_b200.h"
#include "bgr_host.h"
#include "bgr_index_gpu.h"
#include "bgr_kernels.cuh"

using namespace bgr;

namespace {

thread_local std::string g_err;

// bytes of one character buffer of the tile scan: 256 reads of 150 bases and their alignment slack
const u32 TILE_CB = 40960;

struct Slot {
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[3] = {nullptr, nullptr, nullptr};
    // device buffers and their capacities
    char* d_seq = nullptr;
    size_t cap_seq = 0;
    // packed upload (bgr_slot_upload_packed): 2-bit stream, flag bitmap, exception reads as characters
    u64* d_pk = nullptr;
    size_t cap_pk = 0;
    u32* d_bad = nullptr;
    size_t cap_bad = 0;
    u32* d_exc_idx = nullptr;
    size_t cap_exc_idx = 0;
    u64* d_exc_off = nullptr;
    size_t cap_exc_off = 0;
    char* d_exc_seq = nullptr;
    size_t cap_exc_seq = 0;
    char* h_exc = nullptr;  // pinned staging: [exc_off u64 x (n_exc+1)][exc_idx u32 x n_exc][characters]
    size_t cap_h_exc = 0;
    bool packed = false;
    u32 n_exc = 0;
    u64* d_off = nullptr;
    bgr_result* d_res = nullptr;
    u32* d_info = nullptr;
    size_t cap_reads = 0;
    u64 *d_F = nullptr, *d_NM = nullptr;
    u32 *d_H1 = nullptr, *d_H2 = nullptr;
    size_t cap_words = 0;
    i32* d_scratch = nullptr;
    size_t cap_scratch = 0;  // ints
    i32* d_pool = nullptr;
    size_t cap_pool = 0;  // ints
    DevCounters* d_cnt = nullptr;
    DevCounters* h_cnt = nullptr;  // pinned
    ScanCounters* d_scnt = nullptr;
    ScanCounters* h_scnt = nullptr;  // pinned
    u32* d_list = nullptr;
    u32* d_slow = nullptr;  // reads the tile scan hands to the general scan kernel
    unsigned long long* d_bsum = nullptr;  // per-1024-read path length sums (compaction)
    size_t cap_bsum = 0;
    u64* d_corr = nullptr;  // correction mode: corrected reads, 2 bits per base
    size_t cap_corr = 0;
    u64* h_corr = nullptr;
    u64 h_corr_cap = 0;
    bool want_corrected = false;
    // current batch
    u32 n = 0;
    u64 chars = 0, words = 0;
    bgr_params params{};
    u32 scratch_ints = 0;
    u32 tile_r = 0;  // reads per tile of the tile scan for this batch (0: general kernel only)
    bool uploaded = false, ran = false, timed = false, counted = false;
    bgr_result* h_res = nullptr;
    i32* h_pool = nullptr;
    u64 h_pool_cap = 0;
    bool want_download = false;
    // text batches (bgr_slot_upload_text / bgr_slot_format_text): the raw text, its line and record tables
    char* d_text = nullptr;
    size_t cap_text = 0;
    u64 text_bytes = 0;
    bool is_text = false;
    u32 n_rec = 0;
    u32 *d_tcnt = nullptr, *d_tbase = nullptr;  // newlines per 16 bytes, and their exclusive scan
    size_t cap_tcnt = 0;
    u32* d_nl = nullptr;
    size_t cap_nl = 0;
    u32 *d_hoff = nullptr, *d_hlen = nullptr, *d_soff = nullptr, *d_slen = nullptr, *d_keep = nullptr, *d_ridx = nullptr;
    unsigned long long *d_klen = nullptr, *d_roff = nullptr;
    size_t cap_rec = 0;
    u32* d_rec_of = nullptr;
    size_t cap_rec_of = 0;
    void* d_scan_tmp = nullptr;
    size_t cap_scan_tmp = 0;
    TextMeta* d_tmeta = nullptr;
    TextMeta* h_tmeta = nullptr;  // pinned
    unsigned long long *d_osz = nullptr, *d_oat = nullptr;  // [3][n] bytes per read and file, and their scans
    size_t cap_osz = 0;
    char* d_out[3] = {nullptr, nullptr, nullptr};
    size_t cap_out[3] = {0, 0, 0};
    unsigned long long* h_ototal = nullptr;  // pinned: [3] last offset + last size per file
};

}  // namespace

struct bgr_ctx {
    int device = 0;
    int n_slots = 1;
    int sm_count = 148;
    Slot slots[8];
    void* d_blob = nullptr;
    BlobHeader hdr{};
    DevIndex ix{};
    bool has_index = false;
    unsigned long long totals[BGR_C_COUNT] = {0};
    uint64_t launches = 0;
    u32 scratch_greedy = 32, scratch_exhaustive = 128;  // ints per read, grown on overflow
    // resident CTAs per SM the kernels are compiled for (register budget); tuned on B200, see
    // profiles/. BGR_SCAN_MINB / BGR_EXT_MINB override for experiments.
    u32 scan_blocks_per_sm = 4, ext_blocks_per_sm = 6, exh_blocks_per_sm = 6;  // both extension kernels (exhaustive: 6.65 / 6.25 / 6.11 / 6.99 ms with 4 / 5 / 6 / 8 at configs[3]); greedy kernel: 6 CTAs (24 warps, 80 registers) measured 2 % ahead of 5
    bool scan_ws = true;  // tile scan in its warp-specialised form (BGR_SCAN=classic: the barrier-per-phase form)
    std::string err;
};

namespace {

int fail(bgr_ctx* c, int code, const std::string& msg) {
    g_err = msg;
    if (c) c->err = msg;
    return code;
}

#define CK(call)                                                                                     \
    do {                                                                                             \
        cudaError_t e_ = (call);                                                                     \
        if (e_ != cudaSuccess)                                                                       \
            return fail(ctx, BGR_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));      \
    } while (0)

template <typename T>
int grow(bgr_ctx* ctx, T*& p, size_t& cap, size_t need, size_t slack_num = 5, size_t slack_den = 4) {
    if (need <= cap && p) return BGR_OK;
    if (p) CK(cudaFree(p));
    p = nullptr;
    size_t ncap = std::max<size_t>(need * slack_num / slack_den, 256);
    CK(cudaMalloc((void**)&p, ncap * sizeof(T)));
    cap = ncap;
    return BGR_OK;
}

int set_index_from_device_blob(bgr_ctx* ctx) {
    CK(cudaMemcpy(&ctx->hdr, ctx->d_blob, sizeof(BlobHeader), cudaMemcpyDeviceToHost));
    const BlobHeader& h = ctx->hdr;
    if (h.magic != BLOB_MAGIC || h.version != BLOB_VERSION) return fail(ctx, BGR_ERR_FORMAT, "index blob: bad magic/version");
    char* B = (char*)ctx->d_blob;
    ctx->ix.keys = (const u64*)(B + h.off_keys);
    ctx->ix.entries = (const Entry*)(B + h.off_entries);
    ctx->ix.meta = (const UnitigMeta*)(B + h.off_meta);
    ctx->ix.pool = (const u64*)(B + h.off_pool);
    ctx->ix.n_buckets = h.n_buckets;
    ctx->ix.bloom = (const u32*)(B + h.off_bloom);
    ctx->ix.n_bloom = (u32)h.n_bloom;
    ctx->ix.bloom_mode = h.bloom_mode;
    if (h.bloom_mode == BLOOM_SMEM_1HASH && h.n_bloom > BLOOM_SMEM_WORDS)
        return fail(ctx, BGR_ERR_FORMAT, "index blob: single-hash filter larger than the shared-memory tier");
    if (h.bloom_mode > BLOOM_MINIMIZER || (h.bloom_mode == BLOOM_MINIMIZER) != (h.bloom_mode != BLOOM_SMEM_1HASH && h.k - 1 == MM_KM1))
        return fail(ctx, BGR_ERR_FORMAT, "index blob: filter mode does not match k");
    ctx->ix.k = h.k;
    ctx->has_index = true;
    // (A persisting L2 access window over the filter -- cudaStreamAttributeAccessPolicyWindow -- was measured twice
    // and dropped. With the carve-out at its maximum the 64 MB filter of a 3 Gbp index gained 1 % while the carve-out
    // took L2 away from everything else: 5 Mbp scan 2.68 -> 2.97 ms, extension +10 %. With the staged scan, a window
    // for the scan kernel only and a carve-out of 64 / 72 MB: scan 4.79 -> 5.28 / 7.45 ms, extension 3.94 -> 4.29 / 4.67.)
    return BGR_OK;
}

int drop_index(bgr_ctx* ctx) {
    if (ctx->d_blob) CK(cudaFree(ctx->d_blob));
    ctx->d_blob = nullptr;
    ctx->has_index = false;
    return BGR_OK;
}

int install_host_blob(bgr_ctx* ctx, const void* blob, u64 bytes) {
    if (bytes < sizeof(BlobHeader)) return fail(ctx, BGR_ERR_FORMAT, "index blob too small");
    BlobHeader h;
    memcpy(&h, blob, sizeof h);
    if (h.magic != BLOB_MAGIC || h.version != BLOB_VERSION || h.total_bytes != bytes)
        return fail(ctx, BGR_ERR_FORMAT, "index blob: bad magic/version/size");
    CK(cudaSetDevice(ctx->device));
    int rc = drop_index(ctx);
    if (rc) return rc;
    CK(cudaMalloc(&ctx->d_blob, bytes));
    CK(cudaMemcpy(ctx->d_blob, blob, bytes, cudaMemcpyHostToDevice));
    return set_index_from_device_blob(ctx);
}

int launch_extension(bgr_ctx* ctx, Slot& s, const BatchArgs& a) {
    switch (s.params.mode) {
        case BGR_MODE_GREEDY: {
            // persistent warps: enough CTAs to fill the SMs, never more than the batch can use
            u32 blocks = (s.n + EXT_THREADS - 1) / EXT_THREADS;
            const u32 maxb = (u32)ctx->sm_count * ctx->ext_blocks_per_sm;
            if (blocks > maxb) blocks = maxb;
            switch (ctx->ext_blocks_per_sm) {
                case 8: k_extend_greedy<8><<<blocks, EXT_THREADS, 0, s.stream>>>(ctx->ix, a); break;
                case 4: k_extend_greedy<4><<<blocks, EXT_THREADS, 0, s.stream>>>(ctx->ix, a); break;
                case 5: k_extend_greedy<5><<<blocks, EXT_THREADS, 0, s.stream>>>(ctx->ix, a); break;
                default: k_extend_greedy<6><<<blocks, EXT_THREADS, 0, s.stream>>>(ctx->ix, a); break;
            }
            break;
        }
        case BGR_MODE_EXHAUSTIVE:
        case BGR_MODE_EXHAUSTIVE_PATH: {
            // persistent warps, one node of the depth-first search per lane and turn
            u32 blocks = (s.n + EXT_THREADS - 1) / EXT_THREADS;
            const u32 maxb = (u32)ctx->sm_count * ctx->exh_blocks_per_sm;
            if (blocks > maxb) blocks = maxb;
            switch (ctx->exh_blocks_per_sm) {
                case 4: k_extend_exhaustive_nodes<4><<<blocks, EXT_THREADS, 0, s.stream>>>(ctx->ix, a); break;
                case 5: k_extend_exhaustive_nodes<5><<<blocks, EXT_THREADS, 0, s.stream>>>(ctx->ix, a); break;
                case 8: k_extend_exhaustive_nodes<8><<<blocks, EXT_THREADS, 0, s.stream>>>(ctx->ix, a); break;
                default: k_extend_exhaustive_nodes<6><<<blocks, EXT_THREADS, 0, s.stream>>>(ctx->ix, a); break;
            }
            break;
        }
        case BGR_MODE_GREEDY_PATH: {
            const u32 blocks = (s.n + EXT_THREADS - 1) / EXT_THREADS;
            k_extend_exhaustive<<<blocks, EXT_THREADS, 0, s.stream>>>(ctx->ix, a);
            break;
        }
        default: return fail(ctx, BGR_ERR_ARG, "mode not supported");
    }
    ctx->launches++;
    CK(cudaGetLastError());
    return BGR_OK;
}

BatchArgs make_args(Slot& s) {
    BatchArgs a;
    a.seq = s.packed ? nullptr : s.d_seq;
    a.pk = s.packed ? s.d_pk : nullptr;
    a.bad = s.packed ? s.d_bad : nullptr;
    a.exc_idx = s.d_exc_idx;
    a.exc_off = s.d_exc_off;
    a.exc_seq = s.d_exc_seq;
    a.n_exc = s.packed ? s.n_exc : 0;
    a.off = s.d_off;
    a.n = s.n;
    a.F = s.d_F;
    a.NM = s.d_NM;
    a.H1 = s.d_H1;
    a.H2 = s.d_H2;
    a.info = s.d_info;
    a.res = s.d_res;
    a.scratch = s.d_scratch;
    a.scratch_ints = s.scratch_ints;
    {
        // the node-per-turn exhaustive search keeps its rows compact, everything else interleaves them
        bool rows = s.params.mode == BGR_MODE_EXHAUSTIVE || s.params.mode == BGR_MODE_EXHAUSTIVE_PATH;
        if (const char* e = getenv("BGR_SCRATCH")) rows = !strcmp(e, "rows");
        a.sc_row = rows ? s.scratch_ints : 1u;
        a.sc_ent = rows ? 1u : s.n;
    }
    a.pool = s.d_pool;
    a.pool_cap = s.cap_pool;
    a.cnt = s.d_cnt;
    a.scnt = s.d_scnt;
    a.list = s.d_list;
    a.slow_list = s.d_slow;
    a.tile_r = s.tile_r;
    a.tile_cb = TILE_CB;
    a.errors = s.params.errors;
    a.effort = s.params.effort;
    a.mode = s.params.mode;
    a.ext_lazy = 24;  // measured: 8 -> 24 takes 10 % off the greedy kernel (fewer, fuller service rounds)
    if (const char* e = getenv("BGR_EXT_LAZY")) a.ext_lazy = (u32)std::max(1, std::min(32, atoi(e)));
    a.corr = s.d_corr;
    return a;
}

// scratch + pool sized for the current batch and mode
int ensure_work_buffers(bgr_ctx* ctx, Slot& s) {
    int rc;
    if ((rc = grow(ctx, s.d_scratch, s.cap_scratch, (size_t)s.n * s.scratch_ints, 1, 1))) return rc;
    if ((rc = grow(ctx, s.d_pool, s.cap_pool, std::max<size_t>((size_t)s.n * 8, 1024), 1, 1))) return rc;
    return BGR_OK;
}

// path compaction: lengths -> offsets (read order) -> gather into the pool; then counters to the host
int enqueue_compaction(bgr_ctx* ctx, Slot& s) {
    const u32 nb = (s.n + CP_BLOCK - 1) / CP_BLOCK;
    int rc = grow(ctx, s.d_bsum, s.cap_bsum, (size_t)nb + 1);
    if (rc) return rc;
    CK(cudaMemsetAsync(&s.d_cnt->pool_cursor, 0, sizeof(unsigned long long), s.stream));
    CK(cudaMemsetAsync(&s.d_cnt->pool_overflow, 0, sizeof(u32), s.stream));
    BatchArgs a = make_args(s);
    k_path_block_sums<<<nb, CP_THREADS, 0, s.stream>>>(a, s.d_bsum);
    k_path_scan_blocks<<<1, CP_THREADS, 0, s.stream>>>(a, s.d_bsum, nb);
    k_path_gather<<<nb, CP_THREADS, 0, s.stream>>>(a, s.d_bsum);
    ctx->launches += 3;
    if (s.params.flags & BGR_P_CORRECT) {  // needs the final pool offsets the gather just wrote
        k_correct<<<(s.n + 127) / 128, 128, 0, s.stream>>>(ctx->ix, a);
        ctx->launches++;
    }
    CK(cudaGetLastError());
    CK(cudaEventRecord(s.ev[2], s.stream));
    CK(cudaMemcpyAsync(s.h_cnt, s.d_cnt, sizeof(DevCounters), cudaMemcpyDeviceToHost, s.stream));
    return BGR_OK;
}

int enqueue_extension(bgr_ctx* ctx, Slot& s) {
    int rc = ensure_work_buffers(ctx, s);
    if (rc) return rc;
    CK(cudaMemsetAsync(s.d_cnt, 0, sizeof(DevCounters), s.stream));
    BatchArgs a = make_args(s);
    if ((rc = launch_extension(ctx, s, a))) return rc;
    return enqueue_compaction(ctx, s);
}

}  // namespace

namespace {
template <typename T>
int scan_exclusive(bgr_ctx* ctx, Slot& s, const T* in, T* out, size_t n) {
    size_t need = 0;
    CK(cub::DeviceScan::ExclusiveSum(nullptr, need, in, out, (int)n, s.stream));
    if (need > s.cap_scan_tmp) {
        if (s.d_scan_tmp) CK(cudaFree(s.d_scan_tmp));
        s.d_scan_tmp = nullptr;
        CK(cudaMalloc(&s.d_scan_tmp, need * 2));
        s.cap_scan_tmp = need * 2;
    }
    CK(cub::DeviceScan::ExclusiveSum(s.d_scan_tmp, need, in, out, (int)n, s.stream));
    return BGR_OK;
}
}  // namespace

extern "C" {

int bgr_abi_version(void) { return BGR_ABI_VERSION; }

const char* bgr_last_error(const bgr_ctx* ctx) { return ctx ? ctx->err.c_str() : g_err.c_str(); }

int bgr_ctx_create(int device, int n_slots, bgr_ctx** out) {
    bgr_ctx* ctx = nullptr;
    if (!out) return fail(nullptr, BGR_ERR_ARG, "out is null");
    *out = nullptr;
    if (n_slots < 1 || n_slots > 8) return fail(nullptr, BGR_ERR_ARG, "n_slots must be in [1,8]");
    int ndev = 0;
    CK(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) return fail(nullptr, BGR_ERR_CUDA, "no such CUDA device");
    CK(cudaSetDevice(device));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10)
        return fail(nullptr, BGR_ERR_CUDA,
                    std::string("device is not sm_100 class (") + prop.name + "): the kernels are built for sm_100a only");
    ctx = new bgr_ctx();
    ctx->device = device;
    ctx->n_slots = n_slots;
    ctx->sm_count = prop.multiProcessorCount;
    if (const char* e = getenv("BGR_SCAN_MINB")) {
        int v = atoi(e);
        if (v == 4 || v == 6 || v == 8) ctx->scan_blocks_per_sm = (u32)v;
    }
    if (const char* e = getenv("BGR_SCAN")) ctx->scan_ws = strcmp(e, "classic") != 0;
    if (const char* e = getenv("BGR_EXT_MINB")) {
        int v = atoi(e);
        if (v == 4 || v == 5 || v == 6 || v == 8) ctx->ext_blocks_per_sm = (u32)v;
    }
    if (const char* e = getenv("BGR_EXH_MINB")) {
        int v = atoi(e);
        if (v == 4 || v == 5 || v == 6 || v == 8) ctx->exh_blocks_per_sm = (u32)v;
    }
    for (int i = 0; i < n_slots; ++i) {
        Slot& s = ctx->slots[i];
        cudaError_t e = cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking);
        for (int j = 0; j < 3 && e == cudaSuccess; ++j) e = cudaEventCreate(&s.ev[j]);
        if (e == cudaSuccess) e = cudaMalloc((void**)&s.d_cnt, sizeof(DevCounters));
        if (e == cudaSuccess) e = cudaHostAlloc((void**)&s.h_cnt, sizeof(DevCounters), cudaHostAllocDefault);
        if (e == cudaSuccess) e = cudaMalloc((void**)&s.d_scnt, sizeof(ScanCounters));
        if (e == cudaSuccess) e = cudaHostAlloc((void**)&s.h_scnt, sizeof(ScanCounters), cudaHostAllocDefault);
        if (e != cudaSuccess) {
            std::string m = std::string("slot setup: ") + cudaGetErrorString(e);
            bgr_ctx_destroy(ctx);
            return fail(nullptr, BGR_ERR_CUDA, m);
        }
    }
    *out = ctx;
    return BGR_OK;
}

void bgr_ctx_destroy(bgr_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    for (int i = 0; i < ctx->n_slots; ++i) {
        Slot& s = ctx->slots[i];
        if (s.stream) cudaStreamSynchronize(s.stream);
        cudaFree(s.d_seq);
        cudaFree(s.d_pk);
        cudaFree(s.d_bad);
        cudaFree(s.d_exc_idx);
        cudaFree(s.d_exc_off);
        cudaFree(s.d_exc_seq);
        if (s.h_exc) cudaFreeHost(s.h_exc);
        cudaFree(s.d_off);
        cudaFree(s.d_res);
        cudaFree(s.d_info);
        cudaFree(s.d_F);
        cudaFree(s.d_NM);
        cudaFree(s.d_H1);
        cudaFree(s.d_H2);
        cudaFree(s.d_scratch);
        cudaFree(s.d_pool);
        cudaFree(s.d_cnt);
        cudaFree(s.d_scnt);
        cudaFree(s.d_list);
        cudaFree(s.d_slow);
        cudaFree(s.d_bsum);
        cudaFree(s.d_corr);
        cudaFree(s.d_text);
        cudaFree(s.d_tcnt);
        cudaFree(s.d_tbase);
        cudaFree(s.d_nl);
        cudaFree(s.d_hoff);
        cudaFree(s.d_hlen);
        cudaFree(s.d_soff);
        cudaFree(s.d_slen);
        cudaFree(s.d_keep);
        cudaFree(s.d_ridx);
        cudaFree(s.d_klen);
        cudaFree(s.d_roff);
        cudaFree(s.d_rec_of);
        cudaFree(s.d_scan_tmp);
        cudaFree(s.d_tmeta);
        cudaFree(s.d_osz);
        cudaFree(s.d_oat);
        for (int f = 0; f < 3; ++f) cudaFree(s.d_out[f]);
        if (s.h_tmeta) cudaFreeHost(s.h_tmeta);
        if (s.h_ototal) cudaFreeHost(s.h_ototal);
        if (s.h_cnt) cudaFreeHost(s.h_cnt);
        if (s.h_scnt) cudaFreeHost(s.h_scnt);
        for (int j = 0; j < 3; ++j)
            if (s.ev[j]) cudaEventDestroy(s.ev[j]);
        if (s.stream) cudaStreamDestroy(s.stream);
    }
    cudaFree(ctx->d_blob);
    delete ctx;
}

int bgr_device_count(int* n) {
    bgr_ctx* ctx = nullptr;
    if (!n) return fail(nullptr, BGR_ERR_ARG, "n is null");
    CK(cudaGetDeviceCount(n));
    return BGR_OK;
}

int bgr_host_alloc(size_t bytes, void** out) {
    bgr_ctx* ctx = nullptr;
    if (!out) return fail(nullptr, BGR_ERR_ARG, "out is null");
    CK(cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocPortable));
    return BGR_OK;
}

void bgr_host_free(void* p) {
    if (p) cudaFreeHost(p);
}

// ---------------------------------------------------------------------------------- index

int bgr_index_build(bgr_ctx* ctx, const char* seqs, const uint64_t* off, uint32_t n_unitigs, uint32_t k) {
    if (!ctx || !off || (!seqs && n_unitigs)) return fail(ctx, BGR_ERR_ARG, "null argument");
    std::string err;
    const char* where = getenv("BGR_INDEX_BUILD");
    if (where && !strcmp(where, "host")) {
        // the original builder: table filled on the host, one upload (kept for comparison and debugging)
        std::vector<uint64_t> blob;
        int rc = build_index_blob(seqs, off, n_unitigs, k, blob, err);
        if (rc) return fail(ctx, rc, err);
        return install_host_blob(ctx, blob.data(), blob.size() * 8);
    }
    CK(cudaSetDevice(ctx->device));
    void* d_blob = nullptr;
    BlobHeader h;
    int rc = build_index_on_device(seqs, off, n_unitigs, k, &d_blob, &h, &ctx->launches, err);
    if (rc) return fail(ctx, rc, err);
    if ((rc = drop_index(ctx))) {
        cudaFree(d_blob);
        return rc;
    }
    ctx->d_blob = d_blob;
    return set_index_from_device_blob(ctx);
}

int bgr_index_build_file(bgr_ctx* ctx, const char* unitig_fasta, uint32_t k) {
    if (!ctx || !unitig_fasta) return fail(ctx, BGR_ERR_ARG, "null argument");
    // host threads for the loader: BGR_HOST_THREADS, else the machine's cores (at most 32)
    unsigned threads = std::min(32u, std::max(1u, std::thread::hardware_concurrency()));
    if (const char* e = getenv("BGR_HOST_THREADS")) threads = (unsigned)std::max(1, atoi(e));
    PodBuf<char> seqs;
    PodBuf<uint64_t> off;
    const bool opened = load_unitig_file_parallel(unitig_fasta, k, threads, seqs, off);
    if (off.n - 1 > 0xFFFFFFFFull) return fail(ctx, BGR_ERR_LIMIT, "more than 2^32-1 unitig records");
    int rc = bgr_index_build(ctx, seqs.p ? seqs.p : "", off.p, (uint32_t)(off.n - 1), k);
    if (rc) return rc;
    if (!opened) return fail(ctx, BGR_ERR_IO, std::string("cannot open unitig file ") + unitig_fasta);
    return BGR_OK;
}

int bgr_index_stats(const bgr_ctx* ctx, uint64_t* n_unitigs, uint64_t* n_keys, uint64_t* device_bytes) {
    if (!ctx || !ctx->has_index) return BGR_ERR_NO_INDEX;
    if (n_unitigs) *n_unitigs = ctx->hdr.n_unitigs;
    if (n_keys) *n_keys = ctx->hdr.n_keys;
    if (device_bytes) *device_bytes = ctx->hdr.total_bytes;
    return BGR_OK;
}

int bgr_index_blob_size(const bgr_ctx* ctx, uint64_t* bytes) {
    if (!ctx || !ctx->has_index || !bytes) return BGR_ERR_NO_INDEX;
    *bytes = ctx->hdr.total_bytes;
    return BGR_OK;
}

int bgr_index_export_device(const bgr_ctx* cctx, void* dev_dst, uint64_t bytes) {
    bgr_ctx* ctx = const_cast<bgr_ctx*>(cctx);
    if (!ctx || !ctx->has_index) return fail(ctx, BGR_ERR_NO_INDEX, "no index");
    if (!dev_dst || bytes < ctx->hdr.total_bytes) return fail(ctx, BGR_ERR_ARG, "destination too small");
    CK(cudaSetDevice(ctx->device));
    CK(cudaMemcpy(dev_dst, ctx->d_blob, ctx->hdr.total_bytes, cudaMemcpyDeviceToDevice));
    return BGR_OK;
}

int bgr_index_import_device(bgr_ctx* ctx, const void* dev_src, uint64_t bytes) {
    if (!ctx || !dev_src || bytes < sizeof(BlobHeader)) return fail(ctx, BGR_ERR_ARG, "bad argument");
    CK(cudaSetDevice(ctx->device));
    int rc = drop_index(ctx);
    if (rc) return rc;
    CK(cudaMalloc(&ctx->d_blob, bytes));
    CK(cudaMemcpy(ctx->d_blob, dev_src, bytes, cudaMemcpyDeviceToDevice));
    rc = set_index_from_device_blob(ctx);
    if (rc == BGR_OK && ctx->hdr.total_bytes != bytes) {
        drop_index(ctx);
        return fail(ctx, BGR_ERR_FORMAT, "index blob: size mismatch");
    }
    return rc;
}

int bgr_index_replicate(bgr_ctx* ctx, const bgr_ctx* src) {
    if (!ctx || !src || ctx == src) return fail(ctx, BGR_ERR_ARG, "bad argument");
    if (!src->has_index) return fail(ctx, BGR_ERR_NO_INDEX, "source context has no index");
    const u64 bytes = src->hdr.total_bytes;
    CK(cudaSetDevice(ctx->device));
    int rc = drop_index(ctx);
    if (rc) return rc;
    CK(cudaMalloc(&ctx->d_blob, bytes));
    CK(cudaMemcpyPeer(ctx->d_blob, ctx->device, src->d_blob, src->device, bytes));
    return set_index_from_device_blob(ctx);
}

int bgr_index_export_host(const bgr_ctx* cctx, void* host_dst, uint64_t bytes) {
    bgr_ctx* ctx = const_cast<bgr_ctx*>(cctx);
    if (!ctx || !ctx->has_index) return fail(ctx, BGR_ERR_NO_INDEX, "no index");
    if (!host_dst || bytes < ctx->hdr.total_bytes) return fail(ctx, BGR_ERR_ARG, "destination too small");
    CK(cudaSetDevice(ctx->device));
    CK(cudaMemcpy(host_dst, ctx->d_blob, ctx->hdr.total_bytes, cudaMemcpyDeviceToHost));
    return BGR_OK;
}

int bgr_index_import_host(bgr_ctx* ctx, const void* host_src, uint64_t bytes) {
    if (!ctx || !host_src) return fail(ctx, BGR_ERR_ARG, "null argument");
    return install_host_blob(ctx, host_src, bytes);
}

int bgr_index_save_file(const bgr_ctx* cctx, const char* path) {
    bgr_ctx* ctx = const_cast<bgr_ctx*>(cctx);
    if (!ctx || !path) return fail(ctx, BGR_ERR_ARG, "null argument");
    if (!ctx->has_index) return fail(ctx, BGR_ERR_NO_INDEX, "no index");
    const u64 bytes = ctx->hdr.total_bytes;
    std::vector<char> host(bytes);
    int rc = bgr_index_export_host(ctx, host.data(), bytes);
    if (rc) return rc;
    // written under a temporary name and renamed: a reader never sees half a blob
    const std::string tmp = std::string(path) + ".tmp";
    FILE* f = fopen(tmp.c_str(), "wb");
    if (!f) return fail(ctx, BGR_ERR_IO, std::string("cannot write ") + tmp);
    const bool ok = fwrite(host.data(), 1, bytes, f) == bytes;
    if (fclose(f) != 0 || !ok || rename(tmp.c_str(), path) != 0) {
        remove(tmp.c_str());
        return fail(ctx, BGR_ERR_IO, std::string("cannot write ") + path);
    }
    return BGR_OK;
}

int bgr_index_load_file(bgr_ctx* ctx, const char* path, uint32_t k) {
    if (!ctx || !path) return fail(ctx, BGR_ERR_ARG, "null argument");
    MappedFile mf;
    if (!mf.open(path)) return fail(ctx, BGR_ERR_IO, std::string("cannot read ") + path);
    if (mf.size < sizeof(BlobHeader)) return fail(ctx, BGR_ERR_FORMAT, "index file too small");
    BlobHeader h;
    memcpy(&h, mf.data, sizeof h);
    if (h.magic != BLOB_MAGIC || h.version != BLOB_VERSION || h.total_bytes != mf.size)
        return fail(ctx, BGR_ERR_FORMAT, "index file: bad magic/version/size");
    if (h.k != k) return fail(ctx, BGR_ERR_FORMAT, "index file was built for another k");
    return install_host_blob(ctx, mf.data, mf.size);
}

int bgr_index_query(const bgr_ctx* cctx, uint64_t kmer_code, int want_end, int32_t ids[4]) {
    bgr_ctx* ctx = const_cast<bgr_ctx*>(cctx);
    if (!ctx || !ctx->has_index) return -BGR_ERR_NO_INDEX;
    if (cudaSetDevice(ctx->device) != cudaSuccess) return -BGR_ERR_CUDA;
    i32* d = nullptr;
    if (cudaMalloc((void**)&d, 5 * sizeof(i32)) != cudaSuccess) return -BGR_ERR_CUDA;
    k_index_query<<<1, 1>>>(ctx->ix, kmer_code, want_end, d);
    ctx->launches++;
    i32 h[5] = {0, 0, 0, 0, 0};
    cudaError_t e = cudaMemcpy(h, d, sizeof h, cudaMemcpyDeviceToHost);
    cudaFree(d);
    if (e != cudaSuccess) {
        fail(ctx, BGR_ERR_CUDA, cudaGetErrorString(e));
        return -BGR_ERR_CUDA;
    }
    for (int i = 0; i < h[0]; ++i) ids[i] = h[1 + i];
    return h[0];
}

// ---------------------------------------------------------------------------------- mapping

namespace {
// per-read and per-word device buffers every upload form needs; the offsets go up with it
// off: the n+1 offsets on the host, or (off_on_host false) null with `chars_known` characters: the device fills d_off
int upload_common(bgr_ctx* ctx, Slot& s, const uint64_t* off, uint32_t n, bool off_on_host = true, u64 chars_known = 0) {
    // the kernels index reads, tiles and list entries with 32-bit numbers
    if (n > (1u << 30)) return fail(ctx, BGR_ERR_LIMIT, "more than 2^30 reads in one batch: split it");
    const u64 chars = off_on_host ? off[n] : chars_known;
    const u64 words = batch_words(chars, n);
    int rc;
    if (n + 1 > s.cap_reads || !s.d_off) {
        size_t cap = 0, c2 = 0, c3 = 0, c4 = 0, c5 = 0;
        if ((rc = grow(ctx, s.d_off, cap, (size_t)n + 4))) return rc;
        if ((rc = grow(ctx, s.d_res, c2, (size_t)n + 1))) return rc;
        if ((rc = grow(ctx, s.d_info, c3, (size_t)n + 1))) return rc;
        if ((rc = grow(ctx, s.d_list, c4, (size_t)n + 1))) return rc;
        if ((rc = grow(ctx, s.d_slow, c5, (size_t)n + 1))) return rc;
        s.cap_reads = std::min(std::min(std::min(cap, c4), c5), std::min(c2, c3)) - 3;
    }
    if (words > s.cap_words || !s.d_F) {
        size_t c0 = 0, c1 = 0, c2 = 0, c3 = 0;
        if ((rc = grow(ctx, s.d_F, c0, (size_t)words))) return rc;
        if ((rc = grow(ctx, s.d_NM, c1, (size_t)words))) return rc;
        if ((rc = grow(ctx, s.d_H1, c2, (size_t)words))) return rc;
        if ((rc = grow(ctx, s.d_H2, c3, (size_t)words))) return rc;
        s.cap_words = std::min(std::min(c0, c1), std::min(c2, c3));
    }
    if (off_on_host) CK(cudaMemcpyAsync(s.d_off, off, ((size_t)n + 1) * sizeof(u64), cudaMemcpyHostToDevice, s.stream));
    s.n = n;
    s.chars = chars;
    s.words = words;
    s.uploaded = true;
    s.ran = false;
    s.want_download = false;
    s.want_corrected = false;
    s.is_text = false;
    return BGR_OK;
}
}  // namespace

int bgr_slot_upload(bgr_ctx* ctx, int slot, const char* seqs, const uint64_t* off, uint32_t n) {
    if (!ctx || slot < 0 || slot >= ctx->n_slots) return fail(ctx, BGR_ERR_ARG, "bad slot");
    if (!off || (!seqs && n)) return fail(ctx, BGR_ERR_ARG, "null argument");
    if (off[0] != 0) return fail(ctx, BGR_ERR_ARG, "off[0] must be 0");
    Slot& s = ctx->slots[slot];
    CK(cudaSetDevice(ctx->device));
    const u64 chars = off[n];
    int rc;
    if ((rc = grow(ctx, s.d_seq, s.cap_seq, (size_t)chars + 16))) return rc;
    if (chars) CK(cudaMemcpyAsync(s.d_seq, seqs, chars, cudaMemcpyHostToDevice, s.stream));
    if ((rc = upload_common(ctx, s, off, n))) return rc;
    s.packed = false;
    s.n_exc = 0;
    return BGR_OK;
}

int bgr_pack_reads(const char* seqs, const uint64_t* off, uint32_t n, uint32_t threads, uint64_t* words,
                   uint32_t* bad_bits, uint32_t* n_bad) {
    if (!off || (!seqs && n && off[n]) || !words || !bad_bits) return fail(nullptr, BGR_ERR_ARG, "null argument");
    if (off[0] != 0) return fail(nullptr, BGR_ERR_ARG, "off[0] must be 0");
    const uint32_t nb = pack_reads_2bit(seqs ? seqs : "", off, n, threads, words, bad_bits);
    if (n_bad) *n_bad = nb;
    return BGR_OK;
}

int bgr_slot_upload_packed(bgr_ctx* ctx, int slot, const uint64_t* words, const uint64_t* off, uint32_t n,
                           const uint32_t* bad_bits, uint32_t n_bad, const char* seqs) {
    if (!ctx || slot < 0 || slot >= ctx->n_slots) return fail(ctx, BGR_ERR_ARG, "bad slot");
    if (!off || (!words && n && off[n]) || (!bad_bits && n)) return fail(ctx, BGR_ERR_ARG, "null argument");
    if (n_bad && !seqs) return fail(ctx, BGR_ERR_ARG, "flagged reads need their characters (seqs)");
    if (off[0] != 0) return fail(ctx, BGR_ERR_ARG, "off[0] must be 0");
    Slot& s = ctx->slots[slot];
    CK(cudaSetDevice(ctx->device));
    const u64 chars = off[n];
    const size_t pk_words = (size_t)((chars + 31) >> 5), bad_words = ((size_t)n + 31) / 32;
    int rc;
    // the scan reads up to three words past a read's last word (tile granules, funnel shifts)
    if ((rc = grow(ctx, s.d_pk, s.cap_pk, pk_words + 8))) return rc;
    if ((rc = grow(ctx, s.d_bad, s.cap_bad, bad_words + 1))) return rc;
    if (pk_words) CK(cudaMemcpyAsync(s.d_pk, words, pk_words * sizeof(u64), cudaMemcpyHostToDevice, s.stream));
    if (bad_words) CK(cudaMemcpyAsync(s.d_bad, bad_bits, bad_words * sizeof(u32), cudaMemcpyHostToDevice, s.stream));
    u32 n_exc = 0;
    if (n_bad) {
        // the flagged reads, in read order: indices, offsets and characters through one pinned staging buffer
        u64 exc_chars = 0;
        for (size_t w = 0; w < bad_words; ++w)
            for (u32 m = bad_bits[w]; m; m &= m - 1) {
                const u32 r = (u32)(w * 32) + (u32)__builtin_ctz(m);
                if (r >= n) return fail(ctx, BGR_ERR_ARG, "flag bitmap has bits beyond the batch");
                exc_chars += off[r + 1] - off[r];
                ++n_exc;
            }
        const size_t off_bytes = ((size_t)n_exc + 1) * sizeof(u64), idx_bytes = (((size_t)n_exc * sizeof(u32)) + 7) & ~(size_t)7;
        const size_t need = off_bytes + idx_bytes + (size_t)exc_chars + 16;
        if (need > s.cap_h_exc) {
            CK(cudaStreamSynchronize(s.stream));  // the previous batch's copies out of the old buffer
            if (s.h_exc) CK(cudaFreeHost(s.h_exc));
            s.h_exc = nullptr;
            s.cap_h_exc = 0;
            CK(cudaHostAlloc((void**)&s.h_exc, need * 2, cudaHostAllocDefault));
            s.cap_h_exc = need * 2;
        }
        u64* h_off = (u64*)s.h_exc;
        u32* h_idx = (u32*)(s.h_exc + off_bytes);
        char* h_seq = s.h_exc + off_bytes + idx_bytes;
        u32 j = 0;
        u64 at = 0;
        for (size_t w = 0; w < bad_words; ++w)
            for (u32 m = bad_bits[w]; m; m &= m - 1) {
                const u32 r = (u32)(w * 32) + (u32)__builtin_ctz(m);
                const u64 len = off[r + 1] - off[r];
                h_off[j] = at;
                h_idx[j] = r;
                memcpy(h_seq + at, seqs + off[r], len);
                at += len;
                ++j;
            }
        h_off[n_exc] = at;
        if ((rc = grow(ctx, s.d_exc_off, s.cap_exc_off, (size_t)n_exc + 1))) return rc;
        if ((rc = grow(ctx, s.d_exc_idx, s.cap_exc_idx, (size_t)n_exc))) return rc;
        if ((rc = grow(ctx, s.d_exc_seq, s.cap_exc_seq, (size_t)exc_chars + 16))) return rc;
        CK(cudaMemcpyAsync(s.d_exc_off, h_off, off_bytes, cudaMemcpyHostToDevice, s.stream));
        CK(cudaMemcpyAsync(s.d_exc_idx, h_idx, (size_t)n_exc * sizeof(u32), cudaMemcpyHostToDevice, s.stream));
        if (exc_chars) CK(cudaMemcpyAsync(s.d_exc_seq, h_seq, exc_chars, cudaMemcpyHostToDevice, s.stream));
    }
    if ((rc = upload_common(ctx, s, off, n))) return rc;
    s.packed = true;
    s.n_exc = n_exc;
    return BGR_OK;
}

// ---- text batches: records parsed and output formatted on the device (bgr_text.cuh)


int bgr_slot_upload_text(bgr_ctx* ctx, int slot, const char* text, uint64_t bytes, int fastq, uint32_t* n_reads,
                         int* regular) {
    if (!ctx || slot < 0 || slot >= ctx->n_slots) return fail(ctx, BGR_ERR_ARG, "bad slot");
    if ((!text && bytes) || !n_reads || !regular) return fail(ctx, BGR_ERR_ARG, "null argument");
    if (!ctx->has_index) return fail(ctx, BGR_ERR_NO_INDEX, "no index built or imported (the FASTA length filter needs k)");
    if (bytes >= 0xFFFFFF00ull) return fail(ctx, BGR_ERR_LIMIT, "more than 4 GB of text in one batch: split it");
    Slot& s = ctx->slots[slot];
    CK(cudaSetDevice(ctx->device));
    *n_reads = 0;
    *regular = 1;
    int rc;
    if (!s.d_tmeta) {
        CK(cudaMalloc((void**)&s.d_tmeta, sizeof(TextMeta)));
        CK(cudaHostAlloc((void**)&s.h_tmeta, sizeof(TextMeta), cudaHostAllocPortable));
        CK(cudaHostAlloc((void**)&s.h_ototal, 8 * sizeof(unsigned long long), cudaHostAllocPortable));
    }
    const size_t n16 = (size_t)((bytes + 15) / 16);
    if ((rc = grow(ctx, s.d_text, s.cap_text, (size_t)bytes + 64))) return rc;
    if (n16 + 1 > s.cap_tcnt || !s.d_tcnt) {
        size_t c0 = 0, c1 = 0;
        if ((rc = grow(ctx, s.d_tcnt, c0, n16 + 1))) return rc;
        if ((rc = grow(ctx, s.d_tbase, c1, n16 + 1))) return rc;
        s.cap_tcnt = std::min(c0, c1);
    }
    CK(cudaMemsetAsync(s.d_tmeta, 0, sizeof(TextMeta), s.stream));
    if (bytes) CK(cudaMemcpyAsync(s.d_text, text, bytes, cudaMemcpyHostToDevice, s.stream));
    s.text_bytes = bytes;
    u32 n_lines = 0;
    if (bytes) {
        k_text_count<<<(unsigned)((n16 + TX_THREADS - 1) / TX_THREADS), TX_THREADS, 0, s.stream>>>(s.d_text, bytes, s.d_tcnt);
        if ((rc = scan_exclusive(ctx, s, s.d_tcnt, s.d_tbase, n16))) return rc;
        u32 last_base = 0, last_cnt = 0;
        CK(cudaMemcpyAsync(&last_base, s.d_tbase + (n16 - 1), sizeof(u32), cudaMemcpyDeviceToHost, s.stream));
        CK(cudaMemcpyAsync(&last_cnt, s.d_tcnt + (n16 - 1), sizeof(u32), cudaMemcpyDeviceToHost, s.stream));
        CK(cudaStreamSynchronize(s.stream));
        n_lines = last_base + last_cnt;
        ctx->launches += 1;
    }
    const u32 lpr = fastq ? 4u : 2u;
    const u32 n_rec = n_lines / lpr;
    s.n_rec = n_rec;
    if (n_lines % lpr != 0 || (bytes && n_lines == 0)) {  // a partial record: not the regular shape
        *regular = 0;
        return BGR_OK;
    }
    if ((rc = grow(ctx, s.d_nl, s.cap_nl, (size_t)n_lines + 1))) return rc;
    if ((size_t)n_rec + 1 > s.cap_rec || !s.d_hoff) {
        size_t c[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        if ((rc = grow(ctx, s.d_hoff, c[0], (size_t)n_rec + 1))) return rc;
        if ((rc = grow(ctx, s.d_hlen, c[1], (size_t)n_rec + 1))) return rc;
        if ((rc = grow(ctx, s.d_soff, c[2], (size_t)n_rec + 1))) return rc;
        if ((rc = grow(ctx, s.d_slen, c[3], (size_t)n_rec + 1))) return rc;
        if ((rc = grow(ctx, s.d_keep, c[4], (size_t)n_rec + 1))) return rc;
        if ((rc = grow(ctx, s.d_ridx, c[5], (size_t)n_rec + 1))) return rc;
        if ((rc = grow(ctx, s.d_klen, c[6], (size_t)n_rec + 1))) return rc;
        if ((rc = grow(ctx, s.d_roff, c[7], (size_t)n_rec + 1))) return rc;
        s.cap_rec = *std::min_element(c, c + 8);
    }
    if (n_rec) {
        k_text_lines<<<(unsigned)((n16 + TX_THREADS - 1) / TX_THREADS), TX_THREADS, 0, s.stream>>>(s.d_text, bytes, s.d_tbase, s.d_nl);
        k_text_records<<<(n_rec + TX_THREADS - 1) / TX_THREADS, TX_THREADS, 0, s.stream>>>(
            s.d_text, bytes, s.d_nl, n_lines, lpr, ctx->ix.k, s.d_hoff, s.d_hlen, s.d_soff, s.d_slen, s.d_keep, s.d_klen, s.d_tmeta);
        if ((rc = scan_exclusive(ctx, s, s.d_keep, s.d_ridx, n_rec))) return rc;
        if ((rc = scan_exclusive(ctx, s, s.d_klen, s.d_roff, n_rec))) return rc;
        // how many reads, how many characters: needed to size the batch's buffers
        u32 last[2] = {0, 0};
        unsigned long long lastc[2] = {0, 0};
        CK(cudaMemcpyAsync(&last[0], s.d_ridx + (n_rec - 1), sizeof(u32), cudaMemcpyDeviceToHost, s.stream));
        CK(cudaMemcpyAsync(&last[1], s.d_keep + (n_rec - 1), sizeof(u32), cudaMemcpyDeviceToHost, s.stream));
        CK(cudaMemcpyAsync(&lastc[0], s.d_roff + (n_rec - 1), sizeof(unsigned long long), cudaMemcpyDeviceToHost, s.stream));
        CK(cudaMemcpyAsync(&lastc[1], s.d_klen + (n_rec - 1), sizeof(unsigned long long), cudaMemcpyDeviceToHost, s.stream));
        CK(cudaMemcpyAsync(s.h_tmeta, s.d_tmeta, sizeof(TextMeta), cudaMemcpyDeviceToHost, s.stream));
        CK(cudaStreamSynchronize(s.stream));
        ctx->launches += 2;
        if (s.h_tmeta->irregular) {
            *regular = 0;
            return BGR_OK;
        }
        const u32 n = last[0] + last[1];
        const u64 chars = lastc[0] + lastc[1];
        if ((rc = grow(ctx, s.d_seq, s.cap_seq, (size_t)chars + 16))) return rc;
        if ((rc = grow(ctx, s.d_rec_of, s.cap_rec_of, (size_t)n + 1))) return rc;
        if ((rc = upload_common(ctx, s, nullptr, n, false, chars))) return rc;
        k_text_gather<<<(unsigned)(((size_t)n_rec * 32 + TX_THREADS - 1) / TX_THREADS), TX_THREADS, 0, s.stream>>>(
            s.d_text, s.d_soff, s.d_slen, s.d_keep, s.d_ridx, s.d_roff, n_rec, s.d_seq, s.d_off, s.d_rec_of, s.d_tmeta);
        ctx->launches += 1;
        CK(cudaGetLastError());
        *n_reads = n;
    } else {
        if ((rc = upload_common(ctx, s, nullptr, 0, false, 0))) return rc;
        CK(cudaMemsetAsync(s.d_off, 0, sizeof(u64), s.stream));
    }
    s.packed = false;
    s.n_exc = 0;
    s.is_text = true;
    return BGR_OK;
}

int bgr_slot_format_text(bgr_ctx* ctx, int slot, int split_no_overlap, char* const out[3], const uint64_t cap[3],
                         uint64_t used[3]) {
    if (!ctx || slot < 0 || slot >= ctx->n_slots || !out || !cap || !used) return fail(ctx, BGR_ERR_ARG, "bad argument");
    Slot& s = ctx->slots[slot];
    if (!s.is_text || !s.ran) return fail(ctx, BGR_ERR_ARG, "slot holds no text batch that has been run");
    // the mapping itself (scratch / pool growth included), then the text of the three files
    int rc = bgr_slot_wait(ctx, slot, nullptr);
    if (rc) return rc;
    used[0] = used[1] = used[2] = 0;
    if (s.n == 0) return BGR_OK;
    const u32 n = s.n;
    if ((rc = grow(ctx, s.d_osz, s.cap_osz, (size_t)3 * n + 3))) return rc;
    {
        size_t c = s.d_oat ? s.cap_osz : 0;
        if (!s.d_oat || (size_t)3 * n + 3 > c) {
            if (s.d_oat) CK(cudaFree(s.d_oat));
            s.d_oat = nullptr;
            CK(cudaMalloc((void**)&s.d_oat, s.cap_osz * sizeof(unsigned long long)));
        }
    }
    BatchArgs a = make_args(s);
    const u32 correction = (s.params.flags & BGR_P_CORRECT) ? 1u : 0u;
    const unsigned blocks = (n + TX_THREADS - 1) / TX_THREADS;
    k_text_sizes<<<blocks, TX_THREADS, 0, s.stream>>>(a, s.d_rec_of, s.d_hlen, split_no_overlap ? 1u : 0u, correction, s.d_osz);
    for (int f = 0; f < 3; ++f)
        if ((rc = scan_exclusive(ctx, s, s.d_osz + (size_t)f * n, s.d_oat + (size_t)f * n, n))) return rc;
    for (int f = 0; f < 3; ++f) {
        CK(cudaMemcpyAsync(&s.h_ototal[2 * f], s.d_oat + (size_t)f * n + (n - 1), sizeof(unsigned long long), cudaMemcpyDeviceToHost, s.stream));
        CK(cudaMemcpyAsync(&s.h_ototal[2 * f + 1], s.d_osz + (size_t)f * n + (n - 1), sizeof(unsigned long long), cudaMemcpyDeviceToHost, s.stream));
    }
    CK(cudaStreamSynchronize(s.stream));
    bool fits = true;
    for (int f = 0; f < 3; ++f) {
        used[f] = s.h_ototal[2 * f] + s.h_ototal[2 * f + 1];
        if (used[f] > cap[f] || (used[f] && !out[f])) fits = false;
    }
    if (!fits) return fail(ctx, BGR_ERR_LIMIT, "output buffer too small (used[] holds the bytes needed)");
    for (int f = 0; f < 3; ++f)
        if ((rc = grow(ctx, s.d_out[f], s.cap_out[f], (size_t)used[f] + 16))) return rc;
    k_text_format<<<blocks, TX_THREADS, 0, s.stream>>>(a, s.d_text, s.d_rec_of, s.d_hoff, s.d_hlen, split_no_overlap ? 1u : 0u,
                                                        correction, s.d_oat, s.d_out[0], s.d_out[1], s.d_out[2]);
    ctx->launches += 2;
    CK(cudaGetLastError());
    for (int f = 0; f < 3; ++f)
        if (used[f]) CK(cudaMemcpyAsync(out[f], s.d_out[f], used[f], cudaMemcpyDeviceToHost, s.stream));
    CK(cudaStreamSynchronize(s.stream));
    return BGR_OK;
}

int bgr_slot_run(bgr_ctx* ctx, int slot, const bgr_params* p) {
    if (!ctx || slot < 0 || slot >= ctx->n_slots || !p) return fail(ctx, BGR_ERR_ARG, "bad argument");
    if (!ctx->has_index) return fail(ctx, BGR_ERR_NO_INDEX, "no index built or imported");
    Slot& s = ctx->slots[slot];
    if (!s.uploaded) return fail(ctx, BGR_ERR_ARG, "slot has no batch uploaded");
    if (p->mode > BGR_MODE_EXHAUSTIVE_PATH) return fail(ctx, BGR_ERR_ARG, "unknown mode");
    if ((p->flags & BGR_P_CORRECT) && p->mode != BGR_MODE_GREEDY)
        return fail(ctx, BGR_ERR_ARG, "correction mode applies to the greedy aligner only (alignerGreedy.cpp:394)");
    if (p->flags & ~BGR_P_CORRECT) return fail(ctx, BGR_ERR_ARG, "unknown flag in bgr_params.flags");
    CK(cudaSetDevice(ctx->device));
    s.params = *p;
    if (p->flags & BGR_P_CORRECT) {
        int rc = grow(ctx, s.d_corr, s.cap_corr, (size_t)corrected_words(s.chars, s.n));
        if (rc) return rc;
    }
    s.scratch_ints = (p->mode == BGR_MODE_GREEDY) ? ctx->scratch_greedy : ctx->scratch_exhaustive;
    CK(cudaMemsetAsync(s.d_cnt, 0, sizeof(DevCounters), s.stream));
    CK(cudaMemsetAsync(s.d_scnt, 0, sizeof(ScanCounters), s.stream));
    CK(cudaEventRecord(s.ev[0], s.stream));
    if (s.n) {
        // reads per tile from the batch's mean read length: a tile's characters must fit one buffer
        const u64 mean_len = (s.chars + s.n - 1) / s.n;
        // (a packed batch takes a quarter of a byte per base)
        u64 tr = (u64)(TILE_CB - 32) * (s.packed ? 4 : 1) / std::max<u64>(mean_len, 1);
        if (tr > ST_MAX_R) tr = ST_MAX_R;
        if (ctx->scan_ws) {
            // one chunk (32 windows) per worker thread and tile where the reads allow it
            const u64 km1 = ctx->ix.k - 1;
            const u64 chunks = mean_len >= km1 ? (mean_len - km1 + 1 + 31) / 32 : 1;
            tr = std::min<u64>(tr, std::max<u64>(WS_WORKERS / std::max<u64>(chunks, 1), 2));
        }
        s.tile_r = (u32)(tr & ~1ull);
        BatchArgs a = make_args(s);
        if (s.tile_r >= 2) {
            // tile scan: one persistent CTA per SM; the Bloom filter rides in shared memory when it is small
            const bool bloom_smem = ctx->ix.bloom_mode == BLOOM_SMEM_1HASH;
            bool mm_async = ctx->scan_ws && ctx->ix.bloom_mode == BLOOM_MINIMIZER && ctx->ix.k == 31;  // k_scan_tiles_ws' STAGED
            if (const char* e = getenv("BGR_SCAN_MM")) mm_async = mm_async && strcmp(e, "inline") != 0;  // experiment hook
            const size_t smem = ctx->scan_ws ? tile_ws_smem_bytes(bloom_smem ? ctx->ix.n_bloom : 0, TILE_CB, mm_async)
                                             : tile_smem_bytes(bloom_smem ? ctx->ix.n_bloom : 0, TILE_CB);
            const u32 ntiles = (s.n + s.tile_r - 1) / s.tile_r;
            const u32 blocks = std::min<u32>(ntiles, (u32)ctx->sm_count);
#define BGR_LAUNCH_TILES(BS, PK, KS)                                                                                 \
    do {                                                                                                             \
        if (ctx->scan_ws) {                                                                                          \
            if (!BS && KS == 30 && mm_async) {                                                                       \
                CK(cudaFuncSetAttribute(k_scan_tiles_ws<BS, PK, KS, !BS && KS == 30>,                                \
                                        cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));                    \
                k_scan_tiles_ws<BS, PK, KS, !BS && KS == 30><<<blocks, ST_THREADS, smem, s.stream>>>(ctx->ix, a);      \
            } else {                                                                                                 \
                CK(cudaFuncSetAttribute(k_scan_tiles_ws<BS, PK, KS, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                        (int)smem));                                                                 \
                k_scan_tiles_ws<BS, PK, KS, false><<<blocks, ST_THREADS, smem, s.stream>>>(ctx->ix, a);                \
            }                                                                                                        \
        } else {                                                                                                     \
            CK(cudaFuncSetAttribute(k_scan_tiles<BS, PK, KS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
            k_scan_tiles<BS, PK, KS><<<blocks, ST_THREADS, smem, s.stream>>>(ctx->ix, a);                              \
        }                                                                                                            \
    } while (0)
#define BGR_LAUNCH_TILES_K(BS, PK)               \
    do {                                         \
        if (ctx->ix.k == 31) BGR_LAUNCH_TILES(BS, PK, 30); \
        else BGR_LAUNCH_TILES(BS, PK, 0);        \
    } while (0)
            if (bloom_smem && s.packed) BGR_LAUNCH_TILES_K(true, true);
            else if (bloom_smem) BGR_LAUNCH_TILES_K(true, false);
            else if (s.packed) BGR_LAUNCH_TILES_K(false, true);
            else BGR_LAUNCH_TILES_K(false, false);
#undef BGR_LAUNCH_TILES_K
#undef BGR_LAUNCH_TILES
            ctx->launches++;
            CK(cudaGetLastError());
        } else {
            a.slow_list = nullptr;  // reads too long for a tile: the general kernel takes the whole batch
        }
        // general scan for what the tile kernel handed over (usually nothing: the kernel then exits at once)
        u32 blocks = (s.n + PS_WARPS - 1) / PS_WARPS;
        u32 maxb = (u32)ctx->sm_count * ctx->scan_blocks_per_sm;
        if (blocks > maxb) blocks = maxb;
        switch (ctx->scan_blocks_per_sm) {
            case 8: k_pack_scan<8><<<blocks, PS_WARPS * 32, 0, s.stream>>>(ctx->ix, a); break;
            case 6: k_pack_scan<6><<<blocks, PS_WARPS * 32, 0, s.stream>>>(ctx->ix, a); break;
            default: k_pack_scan<4><<<blocks, PS_WARPS * 32, 0, s.stream>>>(ctx->ix, a); break;
        }
        ctx->launches++;
        CK(cudaGetLastError());
    }
    CK(cudaEventRecord(s.ev[1], s.stream));
    CK(cudaMemcpyAsync(s.h_scnt, s.d_scnt, sizeof(ScanCounters), cudaMemcpyDeviceToHost, s.stream));
    if (s.n) {
        int rc = enqueue_extension(ctx, s);
        if (rc) return rc;
    } else {
        CK(cudaEventRecord(s.ev[2], s.stream));
        CK(cudaMemcpyAsync(s.h_cnt, s.d_cnt, sizeof(DevCounters), cudaMemcpyDeviceToHost, s.stream));
    }
    s.ran = true;
    s.timed = true;
    s.counted = false;
    return BGR_OK;
}

int bgr_slot_download(bgr_ctx* ctx, int slot, bgr_result* results, int32_t* path_pool, uint64_t pool_cap) {
    if (!ctx || slot < 0 || slot >= ctx->n_slots) return fail(ctx, BGR_ERR_ARG, "bad slot");
    Slot& s = ctx->slots[slot];
    if (!s.ran) return fail(ctx, BGR_ERR_ARG, "slot has not been run");
    if (!results || (!path_pool && pool_cap)) return fail(ctx, BGR_ERR_ARG, "null argument");
    s.h_res = results;
    s.h_pool = path_pool;
    s.h_pool_cap = pool_cap;
    s.want_download = true;
    return BGR_OK;
}

int bgr_slot_download_corrected(bgr_ctx* ctx, int slot, uint64_t* words, uint64_t cap_words) {
    if (!ctx || slot < 0 || slot >= ctx->n_slots) return fail(ctx, BGR_ERR_ARG, "bad slot");
    Slot& s = ctx->slots[slot];
    if (!s.ran) return fail(ctx, BGR_ERR_ARG, "slot has not been run");
    if (!(s.params.flags & BGR_P_CORRECT)) return fail(ctx, BGR_ERR_ARG, "the run was not made with BGR_P_CORRECT");
    if (!words || cap_words < corrected_words(s.chars, s.n)) return fail(ctx, BGR_ERR_ARG, "corrected buffer too small");
    s.h_corr = words;
    s.h_corr_cap = cap_words;
    s.want_corrected = true;
    return BGR_OK;
}

int bgr_slot_wait(bgr_ctx* ctx, int slot, uint64_t* pool_used) {
    if (!ctx || slot < 0 || slot >= ctx->n_slots) return fail(ctx, BGR_ERR_ARG, "bad slot");
    Slot& s = ctx->slots[slot];
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(s.stream));
    if (!s.ran) return BGR_OK;
    // scratch / pool too small: grow and redo the extension kernel (the seed scan's output is kept)
    for (int attempt = 0; s.n && (s.h_cnt->scratch_overflow || s.h_cnt->pool_overflow); ++attempt) {
        if (attempt > 24) return fail(ctx, BGR_ERR_LIMIT, "extension scratch keeps overflowing");
        int rc;
        if (s.h_cnt->scratch_overflow) {
            s.scratch_ints *= 4;
            if (s.params.mode == BGR_MODE_GREEDY) ctx->scratch_greedy = s.scratch_ints;
            else ctx->scratch_exhaustive = s.scratch_ints;
            rc = enqueue_extension(ctx, s);  // the seed scan's output is kept
        } else {
            rc = grow(ctx, s.d_pool, s.cap_pool, (size_t)s.h_cnt->pool_cursor + 1024, 5, 4);
            if (rc == BGR_OK) rc = enqueue_compaction(ctx, s);  // paths are still in the scratch rows
        }
        if (rc) return rc;
        CK(cudaStreamSynchronize(s.stream));
        s.timed = false;  // event pair no longer brackets a single pass
    }
    if (s.n && s.h_cnt->correct_error)
        return fail(ctx, BGR_ERR_FORMAT, "correction mode: a path spells a sequence shorter than its read");
    const u64 used = s.h_cnt->pool_cursor;
    if (pool_used) *pool_used = used;
    if (!s.counted) {
        for (int i = 0; i < BGR_C_COUNT; ++i) ctx->totals[i] += s.h_cnt->c[i] + s.h_scnt->c[i];
        s.counted = true;
    }
    if (s.want_download) {
        s.want_download = false;
        // the caller may come back with a larger pool: bgr_slot_download + bgr_slot_wait again
        if (used > s.h_pool_cap) return fail(ctx, BGR_ERR_LIMIT, "path pool buffer too small");
        if (s.n) CK(cudaMemcpyAsync(s.h_res, s.d_res, (size_t)s.n * sizeof(bgr_result), cudaMemcpyDeviceToHost, s.stream));
        if (used) CK(cudaMemcpyAsync(s.h_pool, s.d_pool, (size_t)used * sizeof(i32), cudaMemcpyDeviceToHost, s.stream));
        CK(cudaStreamSynchronize(s.stream));
    }
    if (s.want_corrected) {
        s.want_corrected = false;
        if (s.n)
            CK(cudaMemcpyAsync(s.h_corr, s.d_corr, (size_t)corrected_words(s.chars, s.n) * sizeof(u64), cudaMemcpyDeviceToHost,
                               s.stream));
        CK(cudaStreamSynchronize(s.stream));
    }
    return BGR_OK;
}

int bgr_slot_kernel_ms(bgr_ctx* ctx, int slot, float ms[3]) {
    if (!ctx || slot < 0 || slot >= ctx->n_slots || !ms) return fail(ctx, BGR_ERR_ARG, "bad argument");
    Slot& s = ctx->slots[slot];
    if (!s.timed) return fail(ctx, BGR_ERR_ARG, "no timed run on this slot");
    CK(cudaEventSynchronize(s.ev[2]));
    CK(cudaEventElapsedTime(&ms[0], s.ev[0], s.ev[1]));
    CK(cudaEventElapsedTime(&ms[1], s.ev[1], s.ev[2]));
    CK(cudaEventElapsedTime(&ms[2], s.ev[0], s.ev[2]));
    return BGR_OK;
}

int bgr_map_batch(bgr_ctx* ctx, const char* seqs, const uint64_t* off, uint32_t n, const bgr_params* p,
                  bgr_result* results, int32_t* path_pool, uint64_t pool_cap, uint64_t* pool_used) {
    int rc;
    if ((rc = bgr_slot_upload(ctx, 0, seqs, off, n))) return rc;
    if ((rc = bgr_slot_run(ctx, 0, p))) return rc;
    if ((rc = bgr_slot_download(ctx, 0, results, path_pool, pool_cap))) return rc;
    return bgr_slot_wait(ctx, 0, pool_used);
}

int bgr_counters(bgr_ctx* ctx, uint64_t out[BGR_C_COUNT]) {
    if (!ctx || !out) return BGR_ERR_ARG;
    for (int i = 0; i < BGR_C_COUNT; ++i) out[i] = ctx->totals[i];
    return BGR_OK;
}

int bgr_counters_reset(bgr_ctx* ctx) {
    if (!ctx) return BGR_ERR_ARG;
    for (int i = 0; i < BGR_C_COUNT; ++i) ctx->totals[i] = 0;
    return BGR_OK;
}

uint64_t bgr_launch_count(const bgr_ctx* ctx) { return ctx ? ctx->launches : 0; }

}  // extern "C"
