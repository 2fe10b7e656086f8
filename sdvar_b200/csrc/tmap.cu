#include "tmap.cuh"

namespace sdvar {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
      q != cudaDriverEntryPointSuccess || p == nullptr)
    return nullptr;
  fn = reinterpret_cast<EncodeTiledFn>(p);
  return fn;
}

// ---- cache: the same (buffer, shape, box) is encoded again on every launch of the loop (3 maps per GEMM / attention launch,
// ~4.5 k launches per bench step); a tensor map is a pure function of the encode arguments, so each thread keeps the last
// kCacheSlots encodings in a direct-mapped table and replays them.
struct TmapKey {
  const void* base;
  uint64_t dims[4], strides[3];
  uint32_t box[4];
  int dt, rank, swz;
  bool operator==(const TmapKey& o) const {
    if (base != o.base || dt != o.dt || rank != o.rank || swz != o.swz) return false;
    for (int i = 0; i < 4; ++i) if (dims[i] != o.dims[i] || box[i] != o.box[i]) return false;
    return strides[0] == o.strides[0] && strides[1] == o.strides[1] && strides[2] == o.strides[2];
  }
};
constexpr int kCacheSlots = 1024;
struct TmapSlot { TmapKey key; CUtensorMap map; bool valid; };
static thread_local TmapSlot* g_cache = nullptr;
static inline uint64_t mix64(uint64_t h, uint64_t v) { h ^= v + 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2); return h; }

static int make_tmap_uncached(CUtensorMap* out, CUtensorMapDataType dt, int esize, const void* base, int rank, const uint64_t* dims,
                              const uint64_t* strides_bytes, const uint32_t* box, int swz);

static int make_tmap(CUtensorMap* out, CUtensorMapDataType dt, int esize, const void* base, int rank, const uint64_t* dims,
                     const uint64_t* strides_bytes, const uint32_t* box, int swz = 128) {
  if (rank < 2 || rank > 4) return make_tmap_uncached(out, dt, esize, base, rank, dims, strides_bytes, box, swz);
  TmapKey k{};
  k.base = base; k.dt = (int)dt; k.rank = rank; k.swz = swz;
  uint64_t h = mix64((uint64_t)(uintptr_t)base, (uint64_t)dt * 8 + rank + 64 * swz);
  for (int i = 0; i < rank; ++i) { k.dims[i] = dims[i]; k.box[i] = box[i]; h = mix64(h, dims[i] * 1315423911ull + box[i]); }
  for (int i = 0; i + 1 < rank; ++i) { k.strides[i] = strides_bytes[i]; h = mix64(h, strides_bytes[i]); }
  if (!g_cache) g_cache = new TmapSlot[kCacheSlots]();
  TmapSlot& sl = g_cache[h % kCacheSlots];
  if (sl.valid && sl.key == k) { *out = sl.map; return SDVAR_OK; }
  if (int rc = make_tmap_uncached(out, dt, esize, base, rank, dims, strides_bytes, box, swz)) return rc;
  sl.key = k; sl.map = *out; sl.valid = true;
  return SDVAR_OK;
}

static int make_tmap_uncached(CUtensorMap* out, CUtensorMapDataType dt, int esize, const void* base, int rank, const uint64_t* dims,
                              const uint64_t* strides_bytes, const uint32_t* box, int swz) {
  EncodeTiledFn enc = get_encode();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled is not available (driver too old or no driver)");
    return SDVAR_ERR_CUDA;
  }
  SDVAR_REQUIRE(rank >= 2 && rank <= 4, "tensor map rank %d", rank);
  SDVAR_REQUIRE(swz == 128 || swz == 64, "tensor map swizzle %d", swz);
  SDVAR_REQUIRE(((uintptr_t)base & 15) == 0, "TMA base must be 16-byte aligned");
  cuuint64_t gdims[4];
  cuuint64_t gstr[3];
  cuuint32_t gbox[4], estr[4];
  for (int i = 0; i < rank; ++i) {
    gdims[i] = dims[i];
    gbox[i] = box[i];
    estr[i] = 1;
    SDVAR_REQUIRE(box[i] >= 1 && box[i] <= 256, "TMA box[%d]=%u", i, box[i]);
  }
  for (int i = 0; i + 1 < rank; ++i) {
    gstr[i] = strides_bytes[i];
    SDVAR_REQUIRE(strides_bytes[i] % 16 == 0, "TMA stride %llu not a multiple of 16 bytes", (unsigned long long)strides_bytes[i]);
  }
  SDVAR_REQUIRE((int)box[0] * esize == swz, "inner box must span the swizzle width (%d bytes)", swz);
  const CUresult r = enc(out, dt, (cuuint32_t)rank, const_cast<void*>(base), gdims, gstr, gbox,
                         estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult %d (rank %d, dims %llu x %llu)", (int)r, rank,
              (unsigned long long)dims[0], (unsigned long long)dims[1]);
    return SDVAR_ERR_CUDA;
  }
  return SDVAR_OK;
}

int make_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                   const uint32_t* box) {
  return make_tmap(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base, rank, dims, strides_bytes, box);
}
int make_tmap_bf16_nd(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                      const uint32_t* box) {
  return make_tmap(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base, rank, dims, strides_bytes, box, 128);
}
int make_tmap_bf16_sw64(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                        const uint32_t* box) {
  return make_tmap(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base, rank, dims, strides_bytes, box, 64);
}
int make_tmap_f32(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                  const uint32_t* box) {
  return make_tmap(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, base, rank, dims, strides_bytes, box);
}

}  // namespace sdvar
