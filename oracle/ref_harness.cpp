// TEST INFRASTRUCTURE ONLY -- never linked into or called from the product path.
//
// Per-read harness around the UNMODIFIED reference (Malfoy/BGREAT). This file is
// our own code; it only #includes the reference's public header and is linked
// against the reference's aligner.cpp + utils.cpp where they lie under
// /root/reference (see oracle/Makefile). Every member of `class Aligner` is
// public (reference aligner.h:59), so the per-read entry points can be called
// directly:
//   alignReadGreedy          aligner.h:127  (alignerGreedy.cpp:35)
//   alignReadExhaustive      aligner.h:135  (alignerExhaustive.cpp:35)
//   alignReadGreedyPath      aligner.h:152  (alignerPaths.cpp:35)
//   alignReadExhaustivePath  aligner.h:158  (alignerPaths.cpp:66)
//   getReads                 aligner.h:126  (aligner.cpp:46)
//
// Usage:
//   ref_harness align <unitigs.fa> <k> <m> <effort> <mode> <reads.txt>
//        mode in {greedy, exhaustive, greedy_path, exhaustive_path}
//        reads.txt: one read per line (no headers; fed verbatim to alignRead*)
//        output  : one line per read: "<overlapFound> <rc> <n> <p0> <p1> ..."
//        last line: "# readNumber alignedRead notAligned noOverlapRead overlaps"
//   ref_harness parse <reads file> <k> <fastq 0|1> <chunk>
//        output  : the (header, read) pairs getReads yields, as "header\tread"
//        lines, chunk after chunk (a line "#chunk <n>" precedes each chunk).
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <string>
#include <vector>

#include "aligner.h"

static int run_align(int argc, char** argv) {
    if (argc < 8) return 2;
    std::string unitigs = argv[2];
    uint k = atoi(argv[3]);
    uint m = atoi(argv[4]);
    uint effort = atoi(argv[5]);
    std::string mode = argv[6];
    std::ifstream in(argv[7]);
    bool bpath = (mode == "greedy_path" || mode == "exhaustive_path");
    // The ctor fopen()s both output files; the harness never writes to them.
    Aligner al(unitigs, "/dev/null", "/dev/null", k, 1, m, false, false, bpath, false, effort, false);
    {
        // indexUnitigs() prints a timing line on stdout; silence it.
        std::streambuf* old = std::cout.rdbuf(nullptr);
        al.indexUnitigs();
        std::cout.rdbuf(old);
    }
    std::string read;
    while (std::getline(in, read)) {
        bool overlapFound = false, rc = false;
        std::vector<uNumber> path;
        if (mode == "greedy") {
            ++al.readNumber;  // done by the caller alignPartGreedy (alignerGreedy.cpp:382)
            path = al.alignReadGreedy(read, overlapFound, m, rc);
        } else if (mode == "exhaustive") {
            path = al.alignReadExhaustive(read, overlapFound, m);
        } else if (mode == "greedy_path") {
            path = al.alignReadGreedyPath(read, overlapFound, m, rc);
        } else if (mode == "exhaustive_path") {
            path = al.alignReadExhaustivePath(read, overlapFound, m);
        } else {
            return 2;
        }
        printf("%d %d %zu", (int)overlapFound, (int)rc, path.size());
        for (size_t i = 0; i < path.size(); ++i) printf(" %d", path[i]);
        printf("\n");
    }
    printf("# %u %u %u %u %u\n", (uint)al.readNumber, (uint)al.alignedRead, (uint)al.notAligned,
           (uint)al.noOverlapRead, (uint)al.overlaps);
    return 0;
}

static int run_parse(int argc, char** argv) {
    if (argc < 6) return 2;
    uint k = atoi(argv[3]);
    bool fastq = atoi(argv[4]) != 0;
    uint chunk = atoi(argv[5]);
    Aligner al("/dev/null", "/dev/null", "/dev/null", k, 1, 0, false, fastq, false, false, 2, false);
    al.readFile.open(argv[2]);
    std::vector<std::pair<std::string, std::string>> reads;
    // Same loop shape as alignPartGreedy (alignerGreedy.cpp:372-377).
    while (!al.readFile.eof()) {
        al.getReads(reads, chunk);
        printf("#chunk %zu\n", reads.size());
        for (auto& hr : reads) printf("%s\t%s\n", hr.first.c_str(), hr.second.c_str());
    }
    return 0;
}

int main(int argc, char** argv) {
    if (argc < 2) return 2;
    std::string cmd = argv[1];
    if (cmd == "align") return run_align(argc, argv);
    if (cmd == "parse") return run_parse(argc, argv);
    return 2;
}
