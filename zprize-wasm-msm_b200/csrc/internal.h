// internal.h -- what the library's translation units share besides the public C ABI.
// b200msm.cu owns the context (stream, device, error string, launch counter); ntt.cu keeps its own per-context state
// (twiddle tables, scratch) and releases it from b200msm_destroy through b200ntt_release.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/b200msm.h"

cudaStream_t b200msm_internal_stream(b200msm_ctx* ctx);
int b200msm_internal_device(b200msm_ctx* ctx);
void b200msm_internal_count_launches(b200msm_ctx* ctx, uint64_t k);
void b200msm_internal_set_error(b200msm_ctx* ctx, const char* msg);
void b200ntt_release(b200msm_ctx* ctx);
