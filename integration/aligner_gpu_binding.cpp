// The binding INTEGRATION.md describes, as code: Aligner::indexUnitigs and Aligner::alignPartGreedy of the REFERENCE's
// own class, re-implemented on the C ABI of libbgreat_b200.so (include/bgreat_b200.h). integration/Makefile compiles it
// together with the reference's unmodified bgreat.cpp / utils.cpp and a build-time copy of aligner.h / aligner.cpp /
// alignerGreedy.cpp in which only (a) this header is included and a `bgr_ctx* gpu` member added, (b) the two original
// definitions are renamed to *_ref. Everything else -- main(), option parsing, Aligner's constructor, alignAll and its
// thread pool, getReads, printPath, the output files and the summary lines -- is the reference's code, driving the GPU
// path through the seam. TEST / DOCUMENTATION ARTEFACT: the product's own host side is bgreat_b200/csrc/bgr_cli.cpp.
#include <stdexcept>

#include "aligner.h"  // the patched copy: declares gpu, indexUnitigs_ref, alignPartGreedy_ref

using namespace std;

static mutex gpuMutex;  // a bgr_ctx is used by one host thread at a time; alignAll runs coreNumber of these

// aligner.cpp:537-547
void Aligner::indexUnitigs() {
    auto start = chrono::system_clock::now();
    if (!gpu && bgr_ctx_create(0, 2, &gpu) != BGR_OK) throw runtime_error(bgr_last_error(nullptr));
    const int rc = bgr_index_build_file(gpu, unitigFileName.c_str(), k);
    if (rc != BGR_OK && rc != BGR_ERR_IO) throw runtime_error(bgr_last_error(gpu));  // BGR_ERR_IO == empty index
    auto end = chrono::system_clock::now();
    cout << "Indexing in seconds : " << (chrono::duration_cast<chrono::seconds>(end - start).count()) << endl;
}

// alignerGreedy.cpp:367-431: the per-read loop becomes one batch call per 10 000-read chunk
void Aligner::alignPartGreedy() {
    vector<pair<string, string>> multiread;
    while (!readFile.eof()) {
        readMutex.lock();
        getReads(multiread, 10000);
        readMutex.unlock();
        // flatten the chunk the way the ABI wants it: characters + n+1 offsets
        string seqs;
        vector<uint64_t> off{0};
        for (auto& hr : multiread) {
            seqs += hr.second;
            off.push_back(seqs.size());
        }
        const uint32_t n = (uint32_t)multiread.size();
        vector<bgr_result> res(n + 1);
        vector<int32_t> pool(seqs.size() + 2 * (size_t)n + 16);  // a path never exceeds |read| + 2 ints
        vector<uint64_t> corr(BGR_CORRECTED_WORDS(seqs.size(), n));
        uint64_t used = 0;
        bgr_params p{BGR_MODE_GREEDY, errorsMax, tryNumber, correctionMode ? BGR_P_CORRECT : 0u};
        {
            lock_guard<mutex> l(gpuMutex);
            int rc = bgr_slot_upload(gpu, 0, seqs.data(), off.data(), n);
            if (rc == BGR_OK) rc = bgr_slot_run(gpu, 0, &p);
            if (rc == BGR_OK) rc = bgr_slot_download(gpu, 0, res.data(), pool.data(), pool.size());
            if (rc == BGR_OK && correctionMode) rc = bgr_slot_download_corrected(gpu, 0, corr.data(), corr.size());
            if (rc == BGR_OK) rc = bgr_slot_wait(gpu, 0, &used);
            if (rc != BGR_OK) throw runtime_error(bgr_last_error(gpu));
        }
        for (uint32_t i = 0; i < n; ++i) {
            string header = multiread[i].first;
            const uint32_t flags = BGR_RESULT_FLAGS(res[i]);
            // the counters alignReadGreedy bumps (alignerGreedy.cpp:37,52,55), read back from the records
            ++readNumber;
            if (flags & BGR_F_CNT_ALIGNED) ++alignedRead;
            if (flags & BGR_F_CNT_NOTALIGNED) ++notAligned;
            if (flags & BGR_F_CNT_NOOVERLAP) ++noOverlapRead;
            if (flags & BGR_F_MAPPED) {
                if (correctionMode) {  // alignerGreedy.cpp:394-404
                    const string& read = multiread[i].second;
                    const uint64_t* w = corr.data() + BGR_CORRECTED_BASE(off[i], i);
                    string corrected(read.size(), 'A');
                    for (size_t j = 0; j < read.size(); ++j) corrected[j] = "ACGT"[(w[j >> 5] >> (62 - 2 * (j & 31))) & 3];
                    header += '\n' + corrected + '\n';
                } else {
                    vector<uNumber> path(pool.begin() + res[i].path_off, pool.begin() + res[i].path_off + BGR_RESULT_LEN(res[i]));
                    header += '\n' + printPath(path);  // aligner.cpp:600
                }
                pathMutex.lock();
                fwrite(header.c_str(), sizeof(char), header.size(), pathFilef);
                pathMutex.unlock();
            } else {
                header += '\n' + multiread[i].second + '\n';
                notMappedMutex.lock();
                fwrite(header.c_str(), sizeof(char), header.size(), notMappedFilef);
                notMappedMutex.unlock();
            }
        }
    }
}
