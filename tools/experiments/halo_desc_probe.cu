// Round-2 groundwork (NOT part of the product build): how does a K-major SWIZZLE_128B UMMA operand descriptor behave when
// its start address is a 128-byte row inside a TMA-written tile (not 1024-byte aligned), and when the stride between
// 8-row core groups (SBO) is not a multiple of 1024 bytes?
//
// Why: the tap GEMMs are bound by the number of TMA box rows they load.  A 3x3 / 2x2-tap
// convolution loads the same input pixels once per tap.  If the A operand of tap (dy,dx) can be a SHIFTED WINDOW of
// one halo tile in shared memory -- tile 8 pixels wide, halo row pitch 10 pixels, so core group g starts at
// (g*10 + dy*10 + dx) * 128 bytes -- every tile loads its input once instead of 4 / 9 times.
//
// Probe: A = [160 rows][64 bf16] with recognisable integers, loaded by ONE TMA box (SWIZZLE_128B) to a 1024-aligned
// buffer; B = 64x64 identity (K-major), so D[m][n] = A[row(m)][n] shows exactly which bytes the MMA read for row m.
// For each (row offset, SBO in rows, base_offset policy) the kernel runs M=128,N=64,K=64 and the host reports whether
//   D[m][:] == A[off + (m/8)*sbo_rows + (m%8)][:]   ("absolute-address swizzle": the halo plan works as is), or
//   a column-permuted variant (swizzle phase taken relative to the descriptor base), or neither.
//
// Build / run on a B200:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -I simple_b200/csrc -o /tmp/halo_probe \
//        tools/experiments/halo_desc_probe.cu -lcuda && /tmp/halo_probe
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "common.cuh"

using namespace sb;

constexpr int kRows = 160, kCols = 64, kN = 64;

__device__ __forceinline__ uint64_t desc_sw128(uint32_t addr, uint32_t sbo_bytes, uint32_t base_offset) {
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3FFF);
  d |= (uint64_t)(16u >> 4) << 16;                    // LBO (unused for one 64-column K chunk)
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;   // stride between 8-row core groups
  d |= (uint64_t)1 << 46;                             // descriptor version (sm_100)
  d |= (uint64_t)(base_offset & 7u) << 49;
  d |= (uint64_t)2 << 61;                             // SWIZZLE_128B
  return d;
}

__global__ void __launch_bounds__(128)
    probe_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, int off_rows,
                 int sbo_rows, int use_base_offset, float* __restrict__ out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sA = smem;                        // 160 x 128 B
  uint8_t* sB = smem + 20 * 1024;            // 64 x 128 B
  uint64_t* full = reinterpret_cast<uint64_t*>(sB + 8 * 1024);
  uint64_t* done = full + 1;
  uint32_t* slot = reinterpret_cast<uint32_t*>(done + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(full, 1);
    mbar_init(done, 1);
    fence_mbar_init();
  }
  if (warp == 0) {
    tmem_alloc(slot, 64);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *slot;
  if (threadIdx.x == 0) {
    mbar_expect_tx(full, kRows * 128 + kN * 128);
    tma_load_2d(sA, &tmA, full, 0, 0);
    tma_load_2d(sB, &tmB, full, 0, 0);
    mbar_wait(full, 0);
    tc_fence_after();
    const uint32_t idesc = make_idesc_bf16(128, kN, 0, 0);
    const uint32_t a0 = smem_u32(sA) + (uint32_t)off_rows * 128u;
    const uint32_t bo = use_base_offset ? ((a0 >> 7) & 7u) : 0u;
    for (int k = 0; k < 4; ++k) {
      const uint64_t ad = desc_sw128(a0 + k * 32, (uint32_t)sbo_rows * 128u, bo);
      const uint64_t bd = desc_sw128(smem_u32(sB) + k * 32, 1024u, 0u);
      umma_bf16(tmem, ad, bd, idesc, (uint32_t)(k != 0));
    }
    umma_commit(done);
  }
  mbar_wait(done, 0);
  tc_fence_after();
  const int m = warp * 32 + lane;
  for (int c = 0; c < kN; c += 32) {
    uint32_t v[32];
    tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c, v);
    tmem_ld_wait();
    for (int j = 0; j < 32; ++j) out[m * kN + c + j] = __uint_as_float(v[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 64);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int make_map(EncodeTiledFn enc, CUtensorMap* m, void* ptr, int rows) {
  cuuint64_t dims[2] = {(cuuint64_t)kCols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)kCols * 2};
  cuuint32_t box[2] = {(cuuint32_t)kCols, (cuuint32_t)rows};
  cuuint32_t es[2] = {1, 1};
  return enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, ptr, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS
             ? 0
             : 1;
}

static float a_val(int r, int c) { return (float)(((r * 7 + c * 3) % 13) - 6 + (r % 5) * 16); }  // exact in bf16

int main() {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn) {
    printf("no cuTensorMapEncodeTiled\n");
    return 1;
  }
  EncodeTiledFn enc = (EncodeTiledFn)fn;
  std::vector<__nv_bfloat16> hA(kRows * kCols), hB(kN * kCols);
  for (int r = 0; r < kRows; ++r)
    for (int c = 0; c < kCols; ++c) hA[r * kCols + c] = __float2bfloat16(a_val(r, c));
  for (int n = 0; n < kN; ++n)
    for (int k = 0; k < kCols; ++k) hB[n * kCols + k] = __float2bfloat16(n == k ? 1.f : 0.f);
  __nv_bfloat16 *dA, *dB;
  float* dOut;
  cudaMalloc(&dA, hA.size() * 2);
  cudaMalloc(&dB, hB.size() * 2);
  cudaMalloc(&dOut, 128 * kN * 4);
  cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice);
  CUtensorMap tmA, tmB;
  if (make_map(enc, &tmA, dA, kRows) || make_map(enc, &tmB, dB, kN)) {
    printf("tensor map encode failed\n");
    return 1;
  }
  const int smem = 20 * 1024 + 8 * 1024 + 64 + 1024;
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  std::vector<float> out(128 * kN);
  const int offs[] = {0, 1, 3, 8, 10, 11, 21};
  const int sbos[] = {8, 10};
  for (int sbo : sbos)
    for (int off : offs)
      for (int ubo = 0; ubo < 2; ++ubo) {
        if (off + 15 * sbo + 8 > kRows) continue;
        probe_kernel<<<1, 128, smem>>>(tmA, tmB, off, sbo, ubo, dOut);
        if (cudaDeviceSynchronize() != cudaSuccess) {
          printf("off=%d sbo=%d base_offset=%d: launch failed (%s)\n", off, sbo, ubo, cudaGetErrorString(cudaGetLastError()));
          return 2;
        }
        cudaMemcpy(out.data(), dOut, out.size() * 4, cudaMemcpyDeviceToHost);
        int exact = 0, row_ok_cols_permuted = 0, other = 0;
        for (int m = 0; m < 128; ++m) {
          const int r = off + (m / 8) * sbo + (m % 8);
          bool same = true;
          for (int n = 0; n < kN; ++n) same = same && out[m * kN + n] == a_val(r, n);
          if (same) {
            ++exact;
            continue;
          }
          // same row, 16-byte chunks (8 columns) XOR-permuted?
          bool perm = false;
          for (int x = 1; x < 8 && !perm; ++x) {
            bool ok = true;
            for (int n = 0; n < kN; ++n) ok = ok && out[m * kN + n] == a_val(r, (((n / 8) ^ x) * 8) + n % 8);
            perm = ok;
          }
          if (perm) ++row_ok_cols_permuted; else ++other;
        }
        printf("off=%2d rows  SBO=%2d rows  base_offset=%s : %3d/128 rows exact, %3d chunk-permuted, %3d other%s\n", off,
               sbo, ubo ? "(addr>>7)&7" : "0", exact, row_ok_cols_permuted, other,
               exact == 128 ? "   <-- usable as is" : "");
      }
  return 0;
}
