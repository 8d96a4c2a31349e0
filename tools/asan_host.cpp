// ASan/UBSan driver for the host code (tests/test_host_sanitizers.py builds and runs it): the parallel reader (FASTA + FASTQ,
// several thread counts, random batch and slice sizes) against the sequential parser, and the 2-bit packer.
#include "bgr_host.h"
#include <stdio.h>
#include <stdlib.h>
#include <unistd.h>
#include <algorithm>
#include <random>
#include <string>
#include <vector>
using namespace bgr;
static std::string rnd_seq(std::mt19937& g, size_t n, double pn) { std::string s(n, 'A'); for (auto& c : s) { c = "ACGT"[g() & 3]; if ((g() % 10000) < pn * 10000) c = 'N'; } return s; }
int main() {
    std::mt19937 g(5);
    int checks = 0;
    for (int iter = 0; iter < 60; ++iter) {
        const bool fq = iter & 1;
        std::string text;
        const size_t nrec = g() % 3000;
        for (size_t i = 0; i < nrec; ++i) {
            const size_t len = (g() % 20 == 0) ? g() % 40 : 30 + g() % 200;
            std::string s = rnd_seq(g, len, 0.001);
            if (fq) { text += "@r" + std::to_string(i) + "\n" + s + "\n+\n" + std::string(len, 'I') + "\n"; }
            else { text += ">r" + std::to_string(i) + "\n"; if (g() % 30 == 0 && len > 10) { text += s.substr(0, len / 2) + "\n" + s.substr(len / 2) + "\n"; } else text += s + "\n"; }
        }
        if (g() % 3 == 0 && !text.empty()) text.pop_back();
        // sequential
        ReadParser rp(text.data(), text.size(), 31, fq);
        ReadBatch rb;
        while (!rp.done()) rp.next_chunk(10000, rb);
        for (unsigned T : {1u, 3u, 8u}) {
            ParallelReader pr(text.data(), text.size(), 31, fq, T, 1 + g() % 200000, 1 + g() % 5000);
            ParsedBatch pb; size_t n = 0; std::string seqs; std::vector<std::string> hdrs;
            do { pb.clear(); pr.next(pb); for (size_t i = 0; i < pb.size(); ++i) { seqs.append(pb.seqs.p + pb.soff.p[i], pb.soff.p[i + 1] - pb.soff.p[i]); hdrs.emplace_back(pb.hdr.p[i].p, pb.hdr.p[i].n); } n += pb.size(); } while (!pr.done());
            if (n != rb.size() || seqs != std::string(rb.seqs.p, rb.seqs.n)) { printf("MISMATCH iter %d T %u: %zu vs %zu\n", iter, T, n, rb.size()); return 1; }
            for (size_t i = 0; i < n; ++i) if (hdrs[i] != std::string(rb.headers.p + rb.hoff.p[i], rb.hoff.p[i + 1] - rb.hoff.p[i])) { printf("HDR MISMATCH\n"); return 1; }
            ++checks;
        }
        // packer on the parsed batch
        std::vector<uint64_t> words((rb.seqs.n + 31) / 32 + 1), w2 = words;
        std::vector<uint32_t> bad((rb.size() + 31) / 32 + 1), b2 = bad;
        uint32_t n1 = pack_reads_2bit(rb.seqs.p ? rb.seqs.p : "", rb.soff.p, (uint32_t)rb.size(), 1, words.data(), bad.data());
        uint32_t n2 = pack_reads_2bit(rb.seqs.p ? rb.seqs.p : "", rb.soff.p, (uint32_t)rb.size(), 7, w2.data(), b2.data());
        if (n1 != n2 || bad != b2) { printf("PACK MISMATCH\n"); return 1; }
    }
    // unitig loader: parallel (line index, per-worker parts) against the sequential record loop
    for (int iter = 0; iter < 40; ++iter) {
        std::string text;
        const size_t nrec = g() % 2000;
        for (size_t i = 0; i < nrec; ++i) {
            const size_t len = (g() % 400 == 0) ? g() % 31 : 31 + g() % 300;  // a short line stops the loading
            text += ">" + std::to_string(i) + "\n" + rnd_seq(g, len, 0.0) + "\n";
        }
        if (g() % 4 == 0 && !text.empty()) text.pop_back();
        if (g() % 7 == 0) text += ">dangling header";
        char path[] = "/tmp/bgr_asan_unitigs_XXXXXX";
        const int fd = mkstemp(path);
        if (fd < 0) return 2;
        FILE* f = fdopen(fd, "wb");
        fwrite(text.data(), 1, text.size(), f);
        fclose(f);
        std::string s1;
        std::vector<uint64_t> o1;
        const bool ok1 = load_unitig_file(path, 31, s1, o1);
        for (unsigned T : {1u, 3u, 8u}) {
            PodBuf<char> s2;
            PodBuf<uint64_t> o2;
            const bool ok2 = load_unitig_file_parallel(path, 31, T, s2, o2, 1 + g() % 3000);
            if (ok1 != ok2 || o2.n != o1.size() || std::string(s2.p ? s2.p : "", s2.n) != s1 ||
                !std::equal(o1.begin(), o1.end(), o2.p)) {
                printf("UNITIG LOADER MISMATCH iter %d T %u\n", iter, T);
                remove(path);
                return 1;
            }
            ++checks;
        }
        remove(path);
    }
    printf("ok %d checks\n", checks);
    return 0;
}
