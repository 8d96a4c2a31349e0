// Index construction on the GPU (bgr_index_gpu.cu): same blob layout as build_index_blob (bgr_host.h).
#pragma once
#include <stdint.h>

#include <string>

#include "bgr_common.h"

namespace bgr {

// Builds the index blob in device memory from unitig characters + offsets given in HOST memory.
// On success *d_blob_out owns a cudaMalloc'ed blob of hdr_out->total_bytes. `launches` is advanced by
// the number of kernels launched. Returns 0 or a BGR_ERR_* code with `err` set.
int build_index_on_device(const char* seqs, const uint64_t* off, uint32_t n, uint32_t k, void** d_blob_out,
                          BlobHeader* hdr_out, uint64_t* launches, std::string& err);

}  // namespace bgr
