// micro-test: can a K-major UMMA shared-memory descriptor start at an arbitrary ROW of a TMA-swizzled tile?
// (the halo trick of the convolution: one activation strip in shared memory serves the taps dx = -1, 0, +1 as three operand
// views shifted by one pixel = one 64/128-byte row)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_shift umma_shift.cu && ./umma_shift
// For swizzle 64B (KC=32) and 128B (KC=64): loads R=144 rows with ONE TMA box, then for shift = 0..9 runs M=128,N=32 MMAs whose
// A descriptor starts at row `shift`, with base_offset = 0 and base_offset = (addr >> 7) & 7, and prints the max error against
// the host result of rows [shift, shift+128).
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int KC>
__global__ void __launch_bounds__(128, 1)
shift_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, int shift, int use_base_offset, float* out) {
  constexpr int R = 144, N = 32, ROWB = KC * 2;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;
  uint8_t* sB = smem + ((R * ROWB + 1023) & ~1023);
  uint64_t* bar = reinterpret_cast<uint64_t*>(sB + 4096);
  uint64_t* dbar = bar + 1;
  uint32_t* tslot = reinterpret_cast<uint32_t*>(bar + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(dbar)));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 32;" ::"r"(smem_u32(tslot)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tmem = *tslot;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(R * ROWB + N * ROWB));
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(sA)),
                 "l"(reinterpret_cast<uint64_t>(&tmA)), "r"(smem_u32(bar)), "r"(0), "r"(0)
                 : "memory");
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(sB)),
                 "l"(reinterpret_cast<uint64_t>(&tmB)), "r"(smem_u32(bar)), "r"(0), "r"(0)
                 : "memory");
    asm volatile(
        "{\n.reg .pred p;\nW1:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n@p bra D1;\nbra W1;\nD1:\n}\n" ::"r"(smem_u32(bar))
        : "memory");
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    for (int k = 0; k < KC / 16; ++k) {
      const uint32_t a_addr = smem_u32(sA) + shift * ROWB + k * 32, b_addr = smem_u32(sB) + k * 32;
      auto desc = [&](uint32_t addr, bool is_a) {
        uint64_t d = 0;
        d |= (uint64_t)((addr & 0x3FFFF) >> 4);
        d |= (uint64_t)1 << 16;
        d |= (uint64_t)((8 * ROWB) >> 4) << 32;
        d |= (uint64_t)1 << 46;
        if (is_a && use_base_offset) d |= (uint64_t)((addr >> 7) & 7) << 49;
        d |= (uint64_t)(KC == 64 ? 2 : 4) << 61;
        return d;
      };
      const uint64_t ad = desc(a_addr, true), bd = desc(b_addr, false);
      const uint32_t acc = k != 0;
      asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem), "l"(ad),
                   "l"(bd), "r"(idesc), "r"(acc)
                   : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(dbar)) : "memory");
  }
  asm volatile("{\n.reg .pred p;\nW2:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n@p bra D2;\nbra W2;\nD2:\n}\n" ::"r"(smem_u32(dbar))
               : "memory");
  asm volatile("tcgen05.fence::after_thread_sync;");
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
        "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
        "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]),
        "=r"(r[31])
      : "r"(tmem + ((uint32_t)(warp * 32) << 16))
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  for (int j = 0; j < 32; ++j) out[(warp * 32 + lane) * 32 + j] = __uint_as_float(r[j]);
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 32;" ::"r"(tmem));
}

template <int KC>
static void run(EncodeTiledFn enc) {
  constexpr int R = 144, N = 32;
  std::vector<__nv_bfloat16> hA(R * KC), hB(N * KC);
  std::vector<float> fA(R * KC), fB(N * KC);
  srand(1);
  for (int i = 0; i < R * KC; ++i) { float v = (rand() % 17 - 8) / 8.0f; hA[i] = __float2bfloat16(v); fA[i] = v; }
  for (int i = 0; i < N * KC; ++i) { float v = (rand() % 13 - 6) / 4.0f; hB[i] = __float2bfloat16(v); fB[i] = v; }
  __nv_bfloat16 *dA, *dB;
  float* dO;
  cudaMalloc(&dA, R * KC * 2); cudaMalloc(&dB, N * KC * 2); cudaMalloc(&dO, 128 * 32 * 4);
  cudaMemcpy(dA, hA.data(), R * KC * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, hB.data(), N * KC * 2, cudaMemcpyHostToDevice);
  CUtensorMap tmA, tmB;
  const CUtensorMapSwizzle sw = KC == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
  cuuint64_t dims[2] = {(cuuint64_t)KC, (cuuint64_t)R}, str[1] = {(cuuint64_t)KC * 2};
  cuuint32_t box[2] = {(cuuint32_t)KC, (cuuint32_t)R}, es[2] = {1, 1};
  CUresult r1 = enc(&tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dA, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  cuuint64_t dimsB[2] = {(cuuint64_t)KC, (cuuint64_t)N};
  cuuint32_t boxB[2] = {(cuuint32_t)KC, (cuuint32_t)N};
  CUresult r2 = enc(&tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dB, dimsB, str, boxB, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r1 != CUDA_SUCCESS || r2 != CUDA_SUCCESS) { printf("encode failed %d %d\n", (int)r1, (int)r2); return; }
  const size_t smem = 1024 + ((R * KC * 2 + 1023) & ~1023) + 4096 + 64;
  cudaFuncSetAttribute(shift_kernel<KC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  std::vector<float> hO(128 * 32);
  for (int bo = 0; bo < 2; ++bo)
    for (int shift = 0; shift < 10; ++shift) {
      shift_kernel<KC><<<1, 128, smem>>>(tmA, tmB, shift, bo, dO);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("KC=%d shift=%d bo=%d: CUDA error %s\n", KC, shift, bo, cudaGetErrorString(e)); return; }
      cudaMemcpy(hO.data(), dO, 128 * 32 * 4, cudaMemcpyDeviceToHost);
      double maxerr = 0;
      for (int m = 0; m < 128; ++m)
        for (int n = 0; n < 32; ++n) {
          double ref = 0;
          for (int k = 0; k < KC; ++k) ref += (double)fA[(m + shift) * KC + k] * fB[n * KC + k];
          maxerr = fmax(maxerr, fabs(ref - hO[m * 32 + n]));
        }
      printf("swizzle %3dB shift=%d rows base_offset=%s : max |err| = %g %s\n", KC * 2, shift, bo ? "(addr>>7)&7" : "0", maxerr,
             maxerr < 1e-3 ? "OK" : "WRONG");
    }
  cudaFree(dA); cudaFree(dB); cudaFree(dO);
}

int main() {
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaFree(0);
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || !p) { printf("no encode fn\n"); return 1; }
  run<32>(reinterpret_cast<EncodeTiledFn>(p));
  run<64>(reinterpret_cast<EncodeTiledFn>(p));
  return 0;
}
