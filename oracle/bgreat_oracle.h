/* TEST INFRASTRUCTURE ONLY -- CPU oracle for the BGREAT read-to-graph mapping path.
 *
 * A plain-C restatement of the reference algorithm (Malfoy/BGREAT), working on
 * character strings exactly like the reference does. It exists to CHECK the
 * CUDA product path (tests/, __graft_entry__.smoke(), bench.py's cpu_baseline
 * leg) and must never be imported, linked or called by anything under
 * bgreat_b200/ -- the product path has no CPU fallback.
 *
 * Parity status: PINNED. tests/test_oracle_golden.py checks this restatement
 * against golden vectors produced by the compiled, unmodified reference
 * (oracle/_ref/bgreat and oracle/_ref/ref_harness; generator:
 * tests/golden/make_golden.py), and tests/test_oracle_vs_ref.py fuzzes it
 * against the reference directly when oracle/_ref is present.
 */
#ifndef BGREAT_ORACLE_H
#define BGREAT_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct bo_aligner bo_aligner;

enum { BO_GREEDY = 0, BO_EXHAUSTIVE = 1, BO_GREEDY_PATH = 2, BO_EXHAUSTIVE_PATH = 3 };

/* counters, in the order bo_counters() fills them */
enum {
    BO_C_READS = 0,      /* readNumber      aligner.h:68 */
    BO_C_ALIGNED = 1,    /* alignedRead */
    BO_C_NOTALIGNED = 2, /* notAligned */
    BO_C_NOOVERLAP = 3,  /* noOverlapRead */
    BO_C_OVERLAPS = 4,   /* overlaps */
    BO_C_WINDOWS = 5,    /* instrumentation: (k-1)-mer windows examined by the seed scans */
    BO_C_LOOKUPS = 6,    /* instrumentation: dictionary lookups (MPHF.lookup calls) */
    BO_C_QUERIES = 7,    /* instrumentation: getBegin + getEnd calls */
    BO_C_COMPARES = 8,   /* instrumentation: missmatchNumber calls */
    BO_C_CMPBASES = 9,   /* instrumentation: characters handed to missmatchNumber */
    BO_C_COUNT = 10
};

/* Aligner::Aligner (aligner.h:80): k, errorsMax (-m), tryNumber (-e). */
bo_aligner* bo_new(uint32_t k, uint32_t errors, uint32_t effort);
/* other -m / -e for the graph already indexed */
void bo_set_params(bo_aligner* a, uint32_t errors, uint32_t effort);
void bo_free(bo_aligner* a);

/* indexUnitigsAux (aligner.cpp:407-534) on a unitig FASTA file. Returns the number of
 * unitigs indexed, or -1 when the file cannot be opened (the reference then indexes nothing). */
int64_t bo_index_file(bo_aligner* a, const char* unitig_path);
/* Same, unitigs given in memory: n sequences, sequence i = seqs[off[i] .. off[i+1]).
 * Stops at the first sequence shorter than k, like the file loop (aligner.cpp:418). */
int64_t bo_index_mem(bo_aligner* a, const char* seqs, const uint64_t* off, uint32_t n);

/* Per-read entry points. `path` receives up to `cap` ints; the return value is the path
 * length (0 = unmapped), or -1 if cap was too small. *overlap_found / *rc as in the
 * reference signatures (aligner.h:127,135,152,158). bo_align() bumps readNumber for
 * BO_GREEDY like the caller alignPartGreedy does (alignerGreedy.cpp:382). */
int bo_align(bo_aligner* a, int mode, const char* read, uint32_t len, int* overlap_found, int* rc,
             int32_t* path, int cap);

void bo_counters(const bo_aligner* a, uint64_t out[BO_C_COUNT]);
void bo_reset_counters(bo_aligner* a);

/* getReads (aligner.cpp:46-117) driven by the loop of alignPartGreedy (alignerGreedy.cpp:372):
 * parses `reads_path` and returns every (header, read) pair the reference would process, in order.
 * Output buffers are malloc'd; free with bo_free_buf. Returns number of reads or -1 (cannot open;
 * the reference spins forever on a missing file). Layout: headers and reads are concatenated,
 * item i = buf[off[i] .. off[i+1]). */
int64_t bo_parse_reads(const char* reads_path, uint32_t k, int fastq, uint32_t chunk, char** headers,
                       uint64_t** header_off, char** reads, uint64_t** read_off);
void bo_free_buf(void* p);

/* Whole-run restatement: alignAll (aligner.cpp:550-597) with one worker thread, writing the
 * same files the reference binary writes (greedy: paths + notAligned; exhaustive: nothing is
 * written by the reference -- here the per-read paths are written so that they can be compared
 * with the harness). reads_csv is the -r argument. Returns 0 or -1. */
int bo_run(bo_aligner* a, int mode, const char* reads_csv, int fastq, const char* paths_file,
           const char* not_aligned_file);

/* Correction mode (-c): recoverPath (aligner.cpp:268-290) + the reverse complement alignPartGreedy applies
 * when rc is set (alignerGreedy.cpp:394-398). `numbers` is a greedy path [offset, ids...]; `out` needs
 * `size` bytes. Returns the corrected length, or < 0 where the reference aborts. */
int64_t bo_recover_path(const bo_aligner* a, const int32_t* numbers, int n, uint32_t size, int rc, char* out);
/* bo_run in correction mode: greedy aligner, corrected reads instead of paths in paths_file. */
int bo_run_correct(bo_aligner* a, const char* reads_csv, int fastq, const char* paths_file,
                   const char* not_aligned_file);

/* Sequence primitives, exported for unit tests (utils.cpp:52-73,117-192; aligner.cpp:305-315). */
uint64_t bo_str2num(const char* s, uint32_t n);
uint64_t bo_rcb(uint64_t x, uint32_t n);
void bo_reverse_complement(const char* s, uint32_t n, char* out);
uint32_t bo_mismatch_number(const char* seq1, const char* seq2, uint32_t n2, uint32_t n);

#ifdef __cplusplus
}
#endif
#endif
