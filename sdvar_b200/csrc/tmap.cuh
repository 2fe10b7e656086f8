// tmap.cuh -- host-side TMA tensor-map construction.  cuTensorMapEncodeTiled is resolved at run time via
// cudaGetDriverEntryPoint, so libsdvar_b200.so has no link-time dependency on libcuda.so.1.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include "common.cuh"

namespace sdvar {

// bf16 tensor, up to 3 dims (dims[0] innermost, contiguous), 128-byte swizzle, zero fill out of bounds.
// strides_bytes[i] is the byte stride of dims[i+1].  box[0] must be 64 elements (=128 bytes).
int make_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                   const uint32_t* box);
// same as make_tmap_bf16 for up to 4 dims
int make_tmap_bf16_nd(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                      const uint32_t* box);
// bf16 tensor, up to 4 dims, 64-byte swizzle: box[0] must be 32 elements (the convolution's channel chunks: 160/320/640 input
// channels are multiples of 32, not of 64)
int make_tmap_bf16_sw64(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                        const uint32_t* box);
// fp32 tensor, same conventions; box[0] must be 32 elements (=128 bytes).
int make_tmap_f32(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                  const uint32_t* box);

}  // namespace sdvar
