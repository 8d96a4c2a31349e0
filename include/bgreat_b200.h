/* bgreat_b200 -- C ABI of the B200-native read-to-de-Bruijn-graph mapping path.
 *
 * This is the drop-in boundary for BGREAT's hot path. The reference (Malfoy/BGREAT) has no
 * FFI layer; the path sits behind the public methods of `class Aligner` (reference
 * aligner.h:58-179). Each entry point below names the reference interface it replaces.
 * Plain C types only: pointers + sizes, int status codes (0 = ok), no exceptions cross the
 * boundary. One bgr_ctx per GPU, used by one host thread at a time. INTEGRATION.md shows
 * the binding a maintainer of the reference would add.
 *
 * There is no CPU fallback: every mapping call runs the sm_100a kernels or fails.
 */
#ifndef BGREAT_B200_H
#define BGREAT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BGR_ABI_VERSION 2

/* ---- status codes ---------------------------------------------------------------------- */
enum {
    BGR_OK = 0,
    BGR_ERR_CUDA = 1,      /* a CUDA runtime call failed (no device, out of memory, launch error) */
    BGR_ERR_ARG = 2,       /* invalid argument (k out of range, null pointer, bad mode ...) */
    BGR_ERR_IO = 3,        /* file cannot be opened / read */
    BGR_ERR_NO_INDEX = 4,  /* mapping requested before an index was built or imported */
    BGR_ERR_FORMAT = 5,    /* unitig contains a character outside ACGT, or index blob malformed */
    BGR_ERR_LIMIT = 6,     /* a documented limit was exceeded (>= 2^30 unitigs, batch too large) */
    BGR_ERR_BUSY = 7       /* slot still has work in flight */
};

/* ---- aligner modes: which reference per-read entry point a batch call stands for ------- */
enum {
    BGR_MODE_GREEDY = 0,         /* Aligner::alignReadGreedy          aligner.h:127, alignerGreedy.cpp:35 */
    BGR_MODE_EXHAUSTIVE = 1,     /* Aligner::alignReadExhaustive      aligner.h:135, alignerExhaustive.cpp:35 */
    BGR_MODE_GREEDY_PATH = 2,    /* Aligner::alignReadGreedyPath      aligner.h:152, alignerPaths.cpp:35 */
    BGR_MODE_EXHAUSTIVE_PATH = 3 /* Aligner::alignReadExhaustivePath  aligner.h:158, alignerPaths.cpp:66 */
};

/* ---- per-read result record ------------------------------------------------------------ */
/* flags */
#define BGR_F_MAPPED 0x01u        /* path non-empty (alignerGreedy.cpp:393) */
#define BGR_F_OVERLAP_FOUND 0x02u /* the reference's `overlapFound` out-parameter */
#define BGR_F_RC 0x04u            /* the reference's `rc` out-parameter: path describes the RC read */
/* which reference counter this read incremented (aligner.h:68) */
#define BGR_F_CNT_ALIGNED 0x10u
#define BGR_F_CNT_NOTALIGNED 0x20u
#define BGR_F_CNT_NOOVERLAP 0x40u

typedef struct bgr_result {
    uint32_t flags_len; /* bits 0..7 flags, bits 8..31 path length in ints */
    uint32_t path_off;  /* offset (in ints) of the path in the batch's path pool */
} bgr_result;

#define BGR_RESULT_FLAGS(r) ((r).flags_len & 0xFFu)
#define BGR_RESULT_LEN(r) ((r).flags_len >> 8)

/* counters, in the order bgr_counters() fills them (reference aligner.h:68, printed by
 * alignAll, aligner.cpp:589-593) */
enum {
    BGR_C_READS = 0,      /* readNumber */
    BGR_C_ALIGNED = 1,    /* alignedRead */
    BGR_C_NOTALIGNED = 2, /* notAligned */
    BGR_C_NOOVERLAP = 3,  /* noOverlapRead */
    BGR_C_OVERLAPS = 4,   /* overlaps (exhaustive / paths modes: windows listed) */
    BGR_C_WINDOWS = 5,    /* instrumentation: (k-1)-mer windows probed by the seed-scan kernel */
    BGR_C_VERIFIED = 6,   /* instrumentation: windows that passed the Bloom filter and were looked up in the key table */
    BGR_C_COUNT = 8
};

/* parameters of one mapping call: Aligner ctor arguments that act per read (aligner.h:80) */
/* bgr_params.flags */
#define BGR_P_CORRECT 0x1u /* correctionMode (-c, aligner.h:97): also produce the corrected reads; BGR_MODE_GREEDY only */

typedef struct bgr_params {
    uint32_t mode;   /* BGR_MODE_* */
    uint32_t errors; /* errorsMax, -m */
    uint32_t effort; /* tryNumber, -e */
    uint32_t flags;  /* BGR_P_* (0: paths only) */
} bgr_params;

typedef struct bgr_ctx bgr_ctx;

/* ---- context ---------------------------------------------------------------------------- */
/* Replaces Aligner::Aligner (aligner.h:80) as far as device state goes. n_slots (1..8)
 * independent batch slots, each with its own CUDA stream and device buffers, let the caller
 * double-buffer: upload slot B while slot A computes. */
int bgr_ctx_create(int device, int n_slots, bgr_ctx** out);
void bgr_ctx_destroy(bgr_ctx* ctx);
/* Human-readable description of the last error on this context (or of a create failure when
 * ctx == NULL). The pointer stays valid until the next failing call on the same thread. */
const char* bgr_last_error(const bgr_ctx* ctx);
int bgr_abi_version(void);

/* number of CUDA devices this process sees (the reference's parallelism is inside the binary -- coreNumber threads,
 * aligner.cpp:560-586; here a caller makes one context per device and deals batches to them) */
int bgr_device_count(int* n);

/* pinned host memory for asynchronous transfers (cudaHostAlloc, portable: usable by the contexts of every device /
 * cudaFreeHost) */
int bgr_host_alloc(size_t bytes, void** out);
void bgr_host_free(void* p);

/* ---- index ------------------------------------------------------------------------------ */
/* Replaces Aligner::indexUnitigs / indexUnitigsAux (aligner.cpp:407-547) and the BooPHF tables.
 * Unitig i (1-based id i+1... i.e. ids follow the order given) = seqs[off[i] .. off[i+1]).
 * Loading stops at the first sequence shorter than k, like the reference (aligner.cpp:418).
 * Characters must be A/C/G/T. 2 <= k <= 32. */
int bgr_index_build(bgr_ctx* ctx, const char* seqs, const uint64_t* off, uint32_t n_unitigs, uint32_t k);
/* Same from a unitig FASTA file, read with the reference's two-lines-per-record rule
 * (aligner.cpp:415-421). A missing file yields an empty index like the reference does, and
 * returns BGR_ERR_IO so that the caller can decide. */
int bgr_index_build_file(bgr_ctx* ctx, const char* unitig_fasta, uint32_t k);
/* number of unitigs / distinct border (k-1)-mers / bytes resident in HBM */
int bgr_index_stats(const bgr_ctx* ctx, uint64_t* n_unitigs, uint64_t* n_keys, uint64_t* device_bytes);

/* Flat index blob for replication across GPUs (build once, broadcast, import).
 * export/import work on DEVICE pointers so the blob can travel by NCCL without touching the
 * host; the *_host variants are for files and tests. */
int bgr_index_blob_size(const bgr_ctx* ctx, uint64_t* bytes);
int bgr_index_export_device(const bgr_ctx* ctx, void* dev_dst, uint64_t bytes);
int bgr_index_import_device(bgr_ctx* ctx, const void* dev_src, uint64_t bytes);
int bgr_index_export_host(const bgr_ctx* ctx, void* host_dst, uint64_t bytes);
int bgr_index_import_host(bgr_ctx* ctx, const void* host_src, uint64_t bytes);

/* The same replication inside one process: dst gets a copy of src's index, device to device (cudaMemcpyPeer: over
 * NVLink where the devices are peers). dst and src may sit on the same device. */
int bgr_index_replicate(bgr_ctx* dst, const bgr_ctx* src);

/* On-disk index cache (cf. the reference's unused mphf::save / load, BooPHF.h:927-1005):
 * the flat blob as a file. bgr_index_load_file checks magic, version, size and that the blob was built for
 * this k (BGR_ERR_FORMAT otherwise; BGR_ERR_IO if the file cannot be read); whether the file still matches the
 * unitig file it was built from is the caller's business. */
int bgr_index_save_file(const bgr_ctx* ctx, const char* path);
int bgr_index_load_file(bgr_ctx* ctx, const char* path, uint32_t k);

/* Debug/test view of the index: getBegin / getEnd (aligner.cpp:147-267) for one (k-1)-mer code
 * (str2num encoding, utils.cpp:117). Writes up to 4 signed unitig ids, returns the count (<0 on error). */
int bgr_index_query(const bgr_ctx* ctx, uint64_t kmer_code, int want_end, int32_t ids[4]);

/* ---- mapping ---------------------------------------------------------------------------- */
/* One batch = n reads given as the reference's per-read API takes them: character strings
 * (read i = seqs[off[i] .. off[i+1])). Bytes other than A/C/G/T behave like 'N' does in the
 * reference. The batch replaces n calls of Aligner::alignRead* (mode) plus the counter updates
 * of alignPartGreedy / alignPartExhaustive (alignerGreedy.cpp:382, alignerExhaustive.cpp:48-56).
 *
 * A batch holds at most 2^30 reads (BGR_ERR_LIMIT beyond).
 *
 * bgr_map_batch = upload + run + download + wait on slot 0, for callers that do not pipeline.
 * results: n records. path_pool: caller buffer of pool_cap ints; *pool_used receives the ints
 * written. If the paths do not fit, BGR_ERR_LIMIT is returned and *pool_used holds the size needed. */
int bgr_map_batch(bgr_ctx* ctx, const char* seqs, const uint64_t* off, uint32_t n, const bgr_params* p,
                  bgr_result* results, int32_t* path_pool, uint64_t pool_cap, uint64_t* pool_used);

/* Pipelined form. All three are asynchronous on the slot's stream; host buffers should be pinned
 * (bgr_host_alloc) for the copies to overlap. bgr_slot_wait blocks until the slot is idle. */
int bgr_slot_upload(bgr_ctx* ctx, int slot, const char* seqs, const uint64_t* off, uint32_t n);
int bgr_slot_run(bgr_ctx* ctx, int slot, const bgr_params* p);

/* Packed upload: the batch travels over PCIe at 2 bits per base instead of 8 (the reference encodes
 * (k-1)-mers from characters on the fly, str2num / nuc2int, utils.cpp:117-151; here the host packs once).
 *
 * bgr_pack_reads runs on `threads` host threads and needs no GPU: words[w] = characters [32w, 32w+32) of
 * the batch's character stream (the concatenation seqs[0 .. off[n])), first character in the most
 * significant bits, A0 C1 G2 T3 (str2num's digit order); BGR_PACKED_WORDS(off[n]) words are written.
 * bad_bits is a bitmap over the reads (BGR_BAD_WORDS(n) words, bit i%32 of word i/32): set = read i holds
 * a character outside ACGT, which has no 2-bit code and which the reference treats three different ways
 * depending on where it is looked at (str2num -> 3, nuc2int / nuc2intrc -> 0, missmatchNumber -> never
 * equal). *n_bad receives the number of such reads.
 *
 * bgr_slot_upload_packed replaces bgr_slot_upload for such a batch. It copies the 2-bit stream, the
 * offsets and the bitmap, and -- only for the flagged reads -- their characters out of `seqs` (may be NULL
 * when n_bad == 0), through a staging buffer of the slot, so every buffer passed in may be reused once
 * bgr_slot_wait has returned for this slot. Results are identical to those of the character upload. */
#define BGR_PACKED_WORDS(chars) (((uint64_t)(chars) + 31) >> 5)
#define BGR_BAD_WORDS(n) (((uint64_t)(n) + 31) >> 5)
int bgr_pack_reads(const char* seqs, const uint64_t* off, uint32_t n, uint32_t threads, uint64_t* words,
                   uint32_t* bad_bits, uint32_t* n_bad);
int bgr_slot_upload_packed(bgr_ctx* ctx, int slot, const uint64_t* words, const uint64_t* off, uint32_t n,
                           const uint32_t* bad_bits, uint32_t n_bad, const char* seqs);
int bgr_slot_download(bgr_ctx* ctx, int slot, bgr_result* results, int32_t* path_pool, uint64_t pool_cap);
int bgr_slot_wait(bgr_ctx* ctx, int slot, uint64_t* pool_used);

/* Correction mode (reference alignerGreedy.cpp:394-404: recoverPath, aligner.cpp:268-290, then the reverse
 * complement when rc is set). With BGR_P_CORRECT in bgr_params.flags the run also spells every mapped
 * read's path as a sequence of the read's length: 2 bits per base (A0 C1 G2 T3), first base in the most
 * significant bits of a word, read i starting at word BGR_CORRECTED_BASE(off[i], i) of a buffer of
 * BGR_CORRECTED_WORDS(off[n], n) words. Call between bgr_slot_run and bgr_slot_wait, like
 * bgr_slot_download; words of unmapped reads are left untouched. */
#define BGR_CORRECTED_BASE(char_off, i) (((uint64_t)(char_off) >> 5) + (uint64_t)(i))
#define BGR_CORRECTED_WORDS(chars, n) (((uint64_t)(chars) >> 5) + (uint64_t)(n) + 2)
int bgr_slot_download_corrected(bgr_ctx* ctx, int slot, uint64_t* words, uint64_t cap_words);
/* Text batches: record parsing and output formatting on the device, for callers that hold the input as text and
 * want the output as text (the command line). They replace, for input of the regular shape, Aligner::getReads
 * (aligner.cpp:46-117) and the output half of alignPartGreedy / alignPartExhaustive with printPath
 * (alignerGreedy.cpp:393-428, aligner.cpp:600-609).
 *
 * bgr_slot_upload_text: text[0 .. bytes) is a run of whole records starting at a record start -- FASTA (fastq = 0:
 * header line starting with '>', one sequence line) or FASTQ (fastq != 0: four lines), every line terminated by a
 * newline. The reads getReads would return (sequence longer than 2 -- FASTA: and than k -- and made of A C G T N;
 * the others are dropped and not counted) become the slot's batch, in order, exactly as if passed to
 * bgr_slot_upload; *n_reads receives their number. *regular = 0 reports text that is not of that shape (continuation
 * lines, a missing final newline, a partial record, a header without '>'): nothing is uploaded and the caller
 * parses this text itself (the reference's stream semantics for such input are the host parser's business).
 * Blocks until the batch is parsed (the host has to know its size). The text buffer may be reused on return. Needs
 * an index (k). At most 4 GB of text per batch.
 *
 * bgr_slot_format_text (after bgr_slot_run on a text batch; blocks): the three output streams of the batch, in
 * input order -- out[0] header\npath\n of every mapped read (header\ncorrected\n when the run had BGR_P_CORRECT),
 * out[1] header\nread\n of the others, out[2] (split_no_overlap != 0 only: the command line's -o) those of them
 * without any overlap. cap[f] = capacity of out[f]; used[f] receives the bytes written. If a buffer is too small
 * BGR_ERR_LIMIT is returned with used[] = the sizes needed, and the call may be repeated. Counters are added to the
 * context's totals as by bgr_slot_wait. */
int bgr_slot_upload_text(bgr_ctx* ctx, int slot, const char* text, uint64_t bytes, int fastq, uint32_t* n_reads,
                         int* regular);
int bgr_slot_format_text(bgr_ctx* ctx, int slot, int split_no_overlap, char* const out[3], const uint64_t cap[3],
                         uint64_t used[3]);

/* device time of the last bgr_slot_run on this slot, per kernel (ms; CUDA events on the slot's
 * stream): [0] pack+seed-scan kernel, [1] extension kernel, [2] whole run */
int bgr_slot_kernel_ms(bgr_ctx* ctx, int slot, float ms[3]);

/* Counters accumulated over all batches since the context was created / last reset
 * (sum of the per-batch device counters). */
int bgr_counters(bgr_ctx* ctx, uint64_t out[BGR_C_COUNT]);
int bgr_counters_reset(bgr_ctx* ctx);

/* number of kernels launched by this context so far (bench.py's gpu_launches) */
uint64_t bgr_launch_count(const bgr_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif /* BGREAT_B200_H */
