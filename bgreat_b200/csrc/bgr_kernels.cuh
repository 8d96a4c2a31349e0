// sm_100a kernels of the read-to-graph mapping path.
//
//   k_scan_tiles    the seed scan (replaces getNOverlap / getListOverlap + the BooPHF lookups, reference
//                   aligner.cpp:318-378, BooPHF.h:783). Persistent CTAs walk tiles of consecutive reads:
//                   the tile's characters arrive in shared memory by one TMA bulk copy (double
//                   buffered), every thread takes one CHUNK = 32 consecutive windows of one read: packs
//                   the 64 characters under it to 2-bit words, rolls the forward / reverse-complement
//                   (k-1)-mer codes over them, tests each canonical code in the Bloom filter (in shared
//                   memory when the index is small, else global/L2) and verifies the few survivors in the
//                   key table. Output: packed words + an EXACT hit bitmask per read.
//   k_pack_scan     warp per read, any length, N-aware: the general form of the same scan, used for the
//                   reads the tile kernel hands over (reads containing non-ACGT characters, tiles that do
//                   not fit the shared-memory buffer).
//   k_extend_greedy thread per read: alignReadGreedy and its four walkers (reference
//                   alignerGreedy.cpp:35-57,167-364) on packed words, mismatches by XOR + popc.
//   k_extend_exhaustive  thread per read: alignReadExhaustive / alignReadExhaustivePath
//                   (alignerExhaustive.cpp:35-259, alignerPaths.cpp:66-88,127-175,289-460) as an
//                   iterative branch-and-bound depth-first search.
//
// The per-read bodies are __host__ __device__ so that tests/emu can single-step the very same
// code on the CPU (there is no GPU in the build container). The product never runs them on the
// host: libbgreat_b200.so only ever launches the kernels.
//
// This header is the one to include; the code lives in bgr_dev.cuh (shared vocabulary), bgr_scan.cuh (seed scan),
// bgr_extend.cuh (per-read aligners), bgr_emit.cuh (extension kernels, compaction, correction) and bgr_text.cuh
// (record parsing and output formatting for text batches).
#pragma once
#include "bgr_dev.cuh"
#include "bgr_scan.cuh"
#include "bgr_extend.cuh"
#include "bgr_emit.cuh"
#include "bgr_text.cuh"
