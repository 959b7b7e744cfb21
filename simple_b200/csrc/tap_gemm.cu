// Implicit-GEMM convolution ("tap GEMM") and its weight gradient on the sm_100a tensor cores.
//
//   forward / dgrad:  out[b,y,x,n] (+)= sum_t sum_c A[b, y+dy_t, x+dx_t, c] * Wp[n, t*C + c]
//   wgrad:            dw[n, t*C + c] += sum_{b,y,x} dY[b,y,x,n] * A[b, y+dy_t, x+dx_t, c]
//
// Warp-specialised, one output tile per CTA:  warp 0 = TMA producer (4-D box loads of the NHWC activation, shifted by
// the tap, zero-filled outside the image; 2-D loads of the packed weights), warp 1 = tcgen05.mma issuer (single
// elected thread, accumulators in TMEM), warps 2..5 = epilogue (tcgen05.ld -> bias/ReLU -> global store or red.add).
// Shared-memory operand tiles are 128-byte-swizzled rows of 64 bf16, the layout TMA writes and UMMA reads.
// Reference call sites replaced: simple/next_frame_predictor.py:225,229,245-246,137-139,45,266-268,275.
#include <cstdlib>

#include "common.cuh"
#include "../../include/simple_b200.h"

namespace sb {

long long g_launches = 0;
int g_sm_reserve = 0;  // SMs left to a concurrent collective (sb_set_sm_reserve): persistent grids use 148 - reserve
int g_wgrad_tap_groups = 1;  // 1: 3-tap groups in the weight-gradient kernel where eligible (A/B switch for tests)
int g_wgrad_mt2 = 1;  // 1: two 128-row M tiles per weight-gradient CTA where the M side has >= 256 channels (A/B switch)
int g_pair_mode = 1;  // 1: use the CTA-pair (cta_group::2) tap GEMM where eligible; 0: single-CTA kernel only

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn g_encode = nullptr;

int load_driver_entry() {
  if (g_encode) return SB_OK;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
  if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || fn == nullptr) return SB_ERR_DRIVER;
  g_encode = (EncodeTiledFn)fn;
  return SB_OK;
}

// NHWC bf16 tensor [B,H,W,C] -> 4-D map (C, W, H, B), box (64, bw, bh, bb), 128B swizzle, zero OOB fill.
static int make_map_nhwc(CUtensorMap* m, const void* ptr, int B, int H, int W, int C, int bw, int bh, int bb) {
  if (((uintptr_t)ptr & 15) || (C % 8)) return SB_ERR_ARG;
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
  cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
  cuuint32_t box[4] = {64, (cuuint32_t)bw, (cuuint32_t)bh, (cuuint32_t)bb};
  cuuint32_t es[4] = {1, 1, 1, 1};
  CUresult r = g_encode(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, es,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? SB_OK : SB_ERR_DRIVER;
}
// bf16 matrix [rows, ld] (K contiguous) -> 2-D map (K, rows), box (64, box_rows).
static int make_map_2d(CUtensorMap* m, const void* ptr, int rows, int k, int ld, int box_rows) {
  if (((uintptr_t)ptr & 15) || (ld % 8)) return SB_ERR_ARG;
  cuuint64_t dims[2] = {(cuuint64_t)k, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
  cuuint32_t es[2] = {1, 1};
  CUresult r = g_encode(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, es,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? SB_OK : SB_ERR_DRIVER;
}

// bf16 output [B,Ho,Wo,ldo] -> 4-D map (N, Wo, Ho, B), box (32 channels, bw, bh, bb), 64B swizzle: the epilogue
// stages 128 rows x 32 columns (64-byte rows) in shared memory and TMA-stores them (full-line writes, OOB clipped).
static int make_map_out(CUtensorMap* m, const void* ptr, int B, int H, int W, int N, int ld, int bw, int bh, int bb) {
  if (((uintptr_t)ptr & 15) || (ld % 8)) return SB_ERR_ARG;
  cuuint64_t dims[4] = {(cuuint64_t)N, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
  cuuint64_t strides[3] = {(cuuint64_t)ld * 2, (cuuint64_t)W * ld * 2, (cuuint64_t)H * W * ld * 2};
  cuuint32_t box[4] = {32, (cuuint32_t)bw, (cuuint32_t)bh, (cuuint32_t)bb};
  cuuint32_t es[4] = {1, 1, 1, 1};
  CUresult r = g_encode(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, es,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? SB_OK : SB_ERR_DRIVER;
}

// Same, 64 channels x 128-byte rows / 128B swizzle (CTA-pair kernel): half as many row requests per byte for the TMA
// store unit, which is request-rate bound on the 64-byte rows above (measured: ~1400 clk per 128-row store).
static int make_map_out64(CUtensorMap* m, const void* ptr, int B, int H, int W, int N, int ld, int bw, int bh, int bb) {
  if (((uintptr_t)ptr & 15) || (ld % 8)) return SB_ERR_ARG;
  cuuint64_t dims[4] = {(cuuint64_t)N, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
  cuuint64_t strides[3] = {(cuuint64_t)ld * 2, (cuuint64_t)W * ld * 2, (cuuint64_t)H * W * ld * 2};
  cuuint32_t box[4] = {64, (cuuint32_t)bw, (cuuint32_t)bh, (cuuint32_t)bb};
  cuuint32_t es[4] = {1, 1, 1, 1};
  CUresult r = g_encode(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, es,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? SB_OK : SB_ERR_DRIVER;
}

struct TapGemmDev {
  int B, Ho, Wo;
  int bw, bh, bb, rows;
  int tiles_x, tiles_y;
  int m_tiles, n_tiles;
  int C, kchunks, ntaps;
  int tap_dy[SB_MAX_TAPS], tap_dx[SB_MAX_TAPS];
  int N, splits;
  void* out;
  int out_f32, ldo;
  const float* bias;
  int bias_mod;  // bias index = column % bias_mod (transposed convs: N = (phase, channel), one bias per channel)
  int relu, atomic;
  int tma_out;  // bf16 output through the shared-memory staged TMA store path
};

constexpr int kGemmThreads = 192;        // wgrad kernel: TMA warp, MMA warp, 4 epilogue warps
constexpr int kFwdThreads = 320;         // tap GEMM: TMA warp, MMA warp, 8 epilogue warps
constexpr int kATileBytes = 128 * 128;   // 128 rows x 64 bf16
constexpr int kStageOutBytes = 128 * 64;  // epilogue staging buffer: 128 rows x 32 bf16
constexpr int kOutBufs = 4;               // 2 column halves x double buffering
constexpr int kStageOut2Bytes = 128 * 128;  // CTA-pair kernel: 128 rows x 64 bf16 (128-byte rows)

__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d)
               : "memory");
}

// Persistent, warp-specialised tap GEMM.  Each CTA (one per SM) walks the tile list tile = blockIdx.x + i*gridDim.x.
// The accumulator is double-buffered in TMEM (2 x BN columns): while the 8 epilogue warps drain tile i, the MMA warp
// already accumulates tile i+1, and the TMA warp runs ahead through the shared-memory ring across tile boundaries.
template <int BN, int STAGES>
__global__ void __launch_bounds__(kFwdThreads, 1)
    tap_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    const __grid_constant__ CUtensorMap tmO, const __grid_constant__ TapGemmDev p) {
  constexpr int B_BYTES = BN * 128;
  constexpr int TMEM_COLS = 2 * BN <= 256 ? 256 : 512;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * kATileBytes;
  uint8_t* sO = sB + STAGES * B_BYTES;  // [kOutBufs][128 rows][64 B], 64B-swizzled
  uint64_t* full = reinterpret_cast<uint64_t*>(sO + kOutBufs * kStageOutBytes);
  uint64_t* empty = full + STAGES;
  uint64_t* tfull = empty + STAGES;   // [2] accumulator ready
  uint64_t* tempty = tfull + 2;       // [2] accumulator drained
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int KT = p.ntaps * p.kchunks;
  const int total_tiles = p.m_tiles * p.n_tiles * p.splits;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull[s], 1);
      mbar_init(&tempty[s], 8);  // one arrival per epilogue warp
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      tma_prefetch_desc(&tmA);
      tma_prefetch_desc(&tmB);
      const uint32_t tx_bytes = (uint32_t)p.rows * 128u + (uint32_t)B_BYTES;
      int it = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int nt = tile % p.n_tiles;
        const int mt = (tile / p.n_tiles) % p.m_tiles;
        const int sp = tile / (p.n_tiles * p.m_tiles);
        const int tx = mt % p.tiles_x, ty = (mt / p.tiles_x) % p.tiles_y, tb = mt / (p.tiles_x * p.tiles_y);
        const int x0 = tx * p.bw, y0 = ty * p.bh, b0 = tb * p.bb, n0 = nt * BN;
        const int k_begin = (int)((long long)KT * sp / p.splits);
        const int k_end = (int)((long long)KT * (sp + 1) / p.splits);
        for (int kk = k_begin; kk < k_end; ++kk, ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (uint32_t)(it / STAGES) & 1u;
          mbar_wait(&empty[s], ph ^ 1u);
          const int t = kk / p.kchunks;
          const int c0 = (kk - t * p.kchunks) * 64;
          mbar_expect_tx(&full[s], tx_bytes);
          tma_load_4d(sA + s * kATileBytes, &tmA, &full[s], c0, x0 + p.tap_dx[t], y0 + p.tap_dy[t], b0);
          tma_load_2d(sB + s * B_BYTES, &tmB, &full[s], t * p.C + c0, n0);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(128, BN, 0, 0);
      const int last_steps = (p.C - (p.kchunks - 1) * 64 + 15) >> 4;  // K=16 steps of the last channel chunk (1..4)
      int it = 0, tcount = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++tcount) {
        const int sp = tile / (p.n_tiles * p.m_tiles);
        const int k_begin = (int)((long long)KT * sp / p.splits);
        const int nk = (int)((long long)KT * (sp + 1) / p.splits) - k_begin;
        const int as = tcount & 1;
        const uint32_t aph = (uint32_t)(tcount >> 1) & 1u;
        mbar_wait(&tempty[as], aph ^ 1u);  // epilogue has drained this accumulator stage
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(as * BN);
        int chunk = k_begin % p.kchunks;  // channel chunk of the next stage (no division in the issue loop)
        for (int i = 0; i < nk; ++i, ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (uint32_t)(it / STAGES) & 1u;
          mbar_wait(&full[s], ph);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(sA + s * kATileBytes);
          const uint32_t b_addr = smem_u32(sB + s * B_BYTES);
          // channels past C in the last 64-wide chunk are TMA zero fill: skip their K=16 steps (gate/embed: C = 80)
          const int ksteps = (chunk == p.kchunks - 1) ? last_steps : 4;
          chunk = (chunk + 1 == p.kchunks) ? 0 : chunk + 1;
          if (ksteps == 4) {  // (the common case keeps its straight-line issue sequence: this thread is the critical path)
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const uint64_t ad = make_smem_desc_sw128(a_addr + k * 32, 16, 1024);
              const uint64_t bd = make_smem_desc_sw128(b_addr + k * 32, 16, 1024);
              umma_bf16(d_tmem, ad, bd, idesc, (uint32_t)((i | k) != 0));
            }
          } else {
            for (int k = 0; k < ksteps; ++k) {
              const uint64_t ad = make_smem_desc_sw128(a_addr + k * 32, 16, 1024);
              const uint64_t bd = make_smem_desc_sw128(b_addr + k * 32, 16, 1024);
              umma_bf16(d_tmem, ad, bd, idesc, (uint32_t)((i | k) != 0));
            }
          }
          umma_commit(&empty[s]);
        }
        umma_commit(&tfull[as]);
      }
    }
  } else {
    const int e = warp - 2;          // 0..7
    const int lg = warp & 3;         // TMEM lane quarter this warp may access
    const int half = e >> 2;         // interleaved 32-column chunks: half 0 takes chunks 0,2,4.., half 1 takes 1,3,5..
    const int r = lg * 32 + lane;
    const int xl = r % p.bw, yl = (r / p.bw) % p.bh, bl = r / (p.bw * p.bh);
    const bool issuer = ((e & 3) == 0) && lane == 0;  // one TMA-store issuing thread per column half
    const uint32_t bar_id = 1u + (uint32_t)half;       // named barrier of the 4 warps (128 threads) of this half
    uint32_t ochunk = 0;                                // staged chunks so far (staging-buffer parity)
    int tcount = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++tcount) {
      const int nt = tile % p.n_tiles;
      const int mt = (tile / p.n_tiles) % p.m_tiles;
      const int tx = mt % p.tiles_x, ty = (mt / p.tiles_x) % p.tiles_y, tb = mt / (p.tiles_x * p.tiles_y);
      const int x0 = tx * p.bw, y0 = ty * p.bh, b0 = tb * p.bb;
      const int x = x0 + xl, y = y0 + yl, b = b0 + bl, n0 = nt * BN;
      const bool valid = (r < p.rows) && (x < p.Wo) && (y < p.Ho) && (b < p.B);
      const long long pix = ((long long)b * p.Ho + y) * p.Wo + x;
      const int as = tcount & 1;
      const uint32_t aph = (uint32_t)(tcount >> 1) & 1u;
      mbar_wait(&tfull[as], aph);
      tc_fence_after();
      if (p.tma_out) {
        // bf16 output: TMEM -> registers -> (bias, ReLU, bf16) -> 64B-swizzled staging tile -> TMA store.
        // Every output line is written in full 64-byte rows by the TMA engine instead of 16-byte scattered stores.
#pragma unroll 1
        for (int c = half * 32; c < BN; c += 64) {
          if (n0 + c >= p.N) break;  // uniform over the 128 threads of this half
          uint32_t v[32];
          tmem_ld32(tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)(as * BN + c), v);
          if (issuer) bulk_wait_read<1>();  // the store that used this staging buffer two chunks ago has drained
          tmem_ld_wait();
          const int nb = n0 + c;
          float f[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
          if (p.bias) {
            const int nbm = nb % p.bias_mod;  // (bias_mod is a multiple of 32: a 32-column chunk never wraps)
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 bv = __ldg(reinterpret_cast<const float4*>(p.bias + nbm + j));
              f[j] += bv.x; f[j + 1] += bv.y; f[j + 2] += bv.z; f[j + 3] += bv.w;
            }
          }
          if (p.relu) {
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] = fmaxf(f[j], 0.f);
          }
          named_bar_sync(bar_id, 128);  // staging buffer is free (issuer waited above)
          uint8_t* buf = sO + (half * 2 + (int)(ochunk & 1u)) * kStageOutBytes;
          if (r < p.rows) {
            uint8_t* row = buf + r * 64;
            const int sw = (r >> 1) & 3;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              uint4 pk;
              __nv_bfloat162 t0 = __floats2bfloat162_rn(f[8 * j], f[8 * j + 1]);
              __nv_bfloat162 t1 = __floats2bfloat162_rn(f[8 * j + 2], f[8 * j + 3]);
              __nv_bfloat162 t2 = __floats2bfloat162_rn(f[8 * j + 4], f[8 * j + 5]);
              __nv_bfloat162 t3 = __floats2bfloat162_rn(f[8 * j + 6], f[8 * j + 7]);
              pk.x = *reinterpret_cast<uint32_t*>(&t0);
              pk.y = *reinterpret_cast<uint32_t*>(&t1);
              pk.z = *reinterpret_cast<uint32_t*>(&t2);
              pk.w = *reinterpret_cast<uint32_t*>(&t3);
              *reinterpret_cast<uint4*>(row + ((j ^ sw) << 4)) = pk;
            }
          }
          fence_proxy_async_smem();
          named_bar_sync(bar_id, 128);  // all 128 rows staged
          if (issuer) {
            tma_store_4d(&tmO, buf, nb, x0, y0, b0);
            bulk_commit();
          }
          ++ochunk;
        }
      } else {
#pragma unroll 1
        for (int c = half * 32; c < BN; c += 64) {
          if (n0 + c >= p.N) break;  // warp-uniform
          uint32_t v[32];
          tmem_ld32(tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)(as * BN + c), v);
          tmem_ld_wait();
          if (!valid) continue;
          float f[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
          const int nb = n0 + c;
          if (p.atomic) {
            float* o = reinterpret_cast<float*>(p.out) + pix * p.ldo + nb;
#pragma unroll
            for (int j = 0; j < 32; j += 4)
              if (nb + j < p.N) red_add_v4(o + j, f[j], f[j + 1], f[j + 2], f[j + 3]);
          } else {
            if (p.bias) {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (nb + j < p.N) f[j] += __ldg(p.bias + (nb + j) % p.bias_mod);
            }
            if (p.relu) {
#pragma unroll
              for (int j = 0; j < 32; ++j) f[j] = fmaxf(f[j], 0.f);
            }
            if (p.out_f32) {
              float* o = reinterpret_cast<float*>(p.out) + pix * p.ldo + nb;
#pragma unroll
              for (int j = 0; j < 32; j += 4)
                if (nb + j < p.N) *reinterpret_cast<float4*>(o + j) = make_float4(f[j], f[j + 1], f[j + 2], f[j + 3]);
            } else {
              __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(p.out) + pix * p.ldo + nb;
#pragma unroll
              for (int j = 0; j < 32; j += 8) {
                if (nb + j < p.N) {
                  uint4 pk;
                  __nv_bfloat162 t0 = __floats2bfloat162_rn(f[j], f[j + 1]);
                  __nv_bfloat162 t1 = __floats2bfloat162_rn(f[j + 2], f[j + 3]);
                  __nv_bfloat162 t2 = __floats2bfloat162_rn(f[j + 4], f[j + 5]);
                  __nv_bfloat162 t3 = __floats2bfloat162_rn(f[j + 6], f[j + 7]);
                  pk.x = *reinterpret_cast<uint32_t*>(&t0);
                  pk.y = *reinterpret_cast<uint32_t*>(&t1);
                  pk.z = *reinterpret_cast<uint32_t*>(&t2);
                  pk.w = *reinterpret_cast<uint32_t*>(&t3);
                  *reinterpret_cast<uint4*>(o + j) = pk;
                }
              }
            }
          }
        }
      }
      // this warp is done reading the accumulator stage: hand it back to the MMA warp
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[as]);
    }
    if (issuer) bulk_wait_all();  // staged tiles must be read out before the CTA (and its shared memory) retires
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, TMEM_COLS);
}

template <int BN, int STAGES>
static int launch_tap_gemm(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmO, const TapGemmDev& p,
                           cudaStream_t stream) {
  constexpr int smem = STAGES * (kATileBytes + BN * 128) + kOutBufs * kStageOutBytes + (2 * STAGES + 4) * 8 + 16 + 1024;
  static_assert(smem <= 232448, "tap_gemm_kernel: shared memory budget exceeded");
  static bool configured = false;
  if (!configured) {
    if (cudaFuncSetAttribute(tap_gemm_kernel<BN, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) !=
        cudaSuccess)
      return SB_ERR_CUDA;
    configured = true;
  }
  const int total = p.m_tiles * p.n_tiles * p.splits;
  const int sms = 148 - g_sm_reserve;
  const int grid = total < sms ? total : sms;
  tap_gemm_kernel<BN, STAGES><<<grid, kFwdThreads, smem, stream>>>(tmA, tmB, tmO, p);
  ++g_launches;
  SB_CHECK_LAUNCH();
  return SB_OK;
}

// ------------------------------------------------------------------------------------------------------------------
// CTA-pair variant (cta_group::2) for the large bf16-output layers.  One thread-block cluster of two CTAs (one SM
// pair) owns a 256-row x BN tile: CTA r of the pair TMA-loads the A box of M tile 2*pair+r (128 rows) and HALF of the
// weight tile (rows n0 + r*BN/2 ..), the leader CTA issues tcgen05.mma.cta_group::2 (M = 256) which reads A from each
// CTA's own shared memory and the two B halves from both, and accumulates 128 lanes x BN columns in each CTA's TMEM.
// Why: the single-CTA kernel is bound by L2 -> SM operand traffic, not by the tensor pipe (85 FLOP per L2 byte at
// 128 x 256 tiles against the ~6300 B/clk L2 cap, i.e. ~1.05 PFLOP/s at 1.96 GHz -- exactly where it measures).  The
// pair halves the B traffic per FLOP (128 FLOP/B at 256 x 256).  Everything else (persistent tile loop, TMEM double
// buffering, 8 epilogue warps with the 64B-swizzled TMA-store staging) is the single-CTA design.
// Barrier protocol: full[s] lives in the LEADER (2 arrivals: both producers' arrive.expect_tx, the peer's remotely;
// the peer's TMA completes its bytes on the leader's barrier through the .cta_group::2 load form); empty[s] / tfull[a]
// are signalled in BOTH CTAs by one multicast tcgen05.commit; tempty[a] lives in the leader (16 arrivals: the 8
// epilogue warps of each CTA, the peer's remotely).
// ------------------------------------------------------------------------------------------------------------------
// EPI = 1 / 2: fused softmax-CE epilogue of the LOGITS layer (N = 3 x 256, colour-major: an N tile IS one softmax group).
// Reference: simple/next_frame_predictor.py:356-357 (logits conv) + simple/trainer.py:130,134-137 (CE, clip, arg-max).
// The 826 MB logits tensor of a B = 64 step is never written: every epilogue thread owns half of one pixel's 256 logits
// (its two interleaved 64-column chunks, rounded to bf16 exactly like the materialised path), the two halves of a row
// meet through shared memory for the max / first arg-max / exp-sum, and what leaves the SM is the GRADIENT tile
// (EPI = 1: training) through the usual swizzled staging + TMA store, or nothing but the arg-max byte (EPI = 2: rollout).
struct CeParams {
  const uint8_t* target;  // u8 [B,3,HW] (EPI = 1)
  uint8_t* amax;          // u8 [B,3,HW] or NULL
  double* loss_sum;       // += sum max(ce, clip) (EPI = 1)
  float clip, gscale;
  int HW;
};
constexpr int kCeExchangeBytes = 128 * 11 * 4;  // per row: max[2][2], argmax[2][2] (tile parity), sum[2], target logit

template <int BN, int STAGES, int EPI = 0>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kFwdThreads, 1)
    tap_gemm2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                     const __grid_constant__ CUtensorMap tmO, const __grid_constant__ TapGemmDev p,
                     const __grid_constant__ CeParams ce) {
  static_assert(EPI == 0 || BN == 256, "the CE epilogue needs one 256-way softmax group per N tile");
  constexpr int BH = BN / 2;           // weight-tile rows held by each CTA of the pair
  constexpr int B_BYTES = BH * 128;
  constexpr int TMEM_COLS = 2 * BN <= 256 ? 256 : 512;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * kATileBytes;
  uint8_t* sO = sB + STAGES * B_BYTES;
  uint64_t* full = reinterpret_cast<uint64_t*>(sO + kOutBufs * kStageOut2Bytes);
  uint64_t* empty = full + STAGES;
  uint64_t* tfull = empty + STAGES;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;
  const int KT = p.ntaps * p.kchunks;
  const int m_pairs = (p.m_tiles + 1) >> 1;
  const int total_tiles = m_pairs * p.n_tiles;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 2);
      mbar_init(&empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull[s], 1);
      mbar_init(&tempty[s], 16);
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc_2cta(tmem_slot, TMEM_COLS);
    tmem_relinquish_2cta();
  }
  tc_fence_before();
  cluster_sync_all();  // barriers of both CTAs initialised, TMEM allocated, before any remote arrival / multicast commit
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      tma_prefetch_desc(&tmA);
      tma_prefetch_desc(&tmB);
      const uint32_t tx_bytes = (uint32_t)p.rows * 128u + (uint32_t)B_BYTES;
      int it = 0;
      for (int tile = pair; tile < total_tiles; tile += npairs) {
        const int nt = tile % p.n_tiles;
        const int mt = 2 * (tile / p.n_tiles) + (int)rank;  // may be one past the last M tile: TMA zero-fills / clips
        const int tx = mt % p.tiles_x, ty = (mt / p.tiles_x) % p.tiles_y, tb = mt / (p.tiles_x * p.tiles_y);
        const int x0 = tx * p.bw, y0 = ty * p.bh, b0 = tb * p.bb, n0 = nt * BN + (int)rank * BH;
        for (int kk = 0; kk < KT; ++kk, ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (uint32_t)(it / STAGES) & 1u;
          mbar_wait(&empty[s], ph ^ 1u);
          const int t = kk / p.kchunks;
          const int c0 = (kk - t * p.kchunks) * 64;
          const uint32_t lead_full = mapa_u32(&full[s], 0);
          mbar_expect_tx_cluster(lead_full, tx_bytes);
          tma_load_4d_2cta(sA + s * kATileBytes, &tmA, lead_full, c0, x0 + p.tap_dx[t], y0 + p.tap_dy[t], b0);
          tma_load_2d_2cta(sB + s * B_BYTES, &tmB, lead_full, t * p.C + c0, n0);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && rank == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(256, BN, 0, 0);
      const int last_steps = (p.C - (p.kchunks - 1) * 64 + 15) >> 4;
      int it = 0, tcount = 0;
      for (int tile = pair; tile < total_tiles; tile += npairs, ++tcount) {
        const int as = tcount & 1;
        const uint32_t aph = (uint32_t)(tcount >> 1) & 1u;
        mbar_wait(&tempty[as], aph ^ 1u);  // both CTAs' epilogues have drained this accumulator stage
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(as * BN);
        int chunk = 0;
        for (int i = 0; i < KT; ++i, ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (uint32_t)(it / STAGES) & 1u;
          mbar_wait(&full[s], ph);  // both CTAs' operand tiles of this stage have landed
          tc_fence_after();
          const uint32_t a_addr = smem_u32(sA + s * kATileBytes);
          const uint32_t b_addr = smem_u32(sB + s * B_BYTES);
          const int ksteps = (chunk == p.kchunks - 1) ? last_steps : 4;  // see the single-CTA kernel
          chunk = (chunk + 1 == p.kchunks) ? 0 : chunk + 1;
          if (ksteps == 4) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const uint64_t ad = make_smem_desc_sw128(a_addr + k * 32, 16, 1024);
              const uint64_t bd = make_smem_desc_sw128(b_addr + k * 32, 16, 1024);
              umma_bf16_2cta(d_tmem, ad, bd, idesc, (uint32_t)((i | k) != 0));
            }
          } else {
            for (int k = 0; k < ksteps; ++k) {
              const uint64_t ad = make_smem_desc_sw128(a_addr + k * 32, 16, 1024);
              const uint64_t bd = make_smem_desc_sw128(b_addr + k * 32, 16, 1024);
              umma_bf16_2cta(d_tmem, ad, bd, idesc, (uint32_t)((i | k) != 0));
            }
          }
          umma_commit_2cta(&empty[s], 3);
        }
        umma_commit_2cta(&tfull[as], 3);
      }
    }
  } else if (EPI != 0) {
    // ---------------- fused softmax-CE epilogue ----------------
    constexpr float LOG2E = 1.4426950408889634f, LN2 = 0.6931471805599453f;
    const int e = warp - 2;
    const int lg = warp & 3;
    const int half = e >> 2;           // half 0 owns columns [0,64) u [128,192), half 1 owns [64,128) u [192,256)
    const int r = lg * 32 + lane;      // row of the 128-row tile = TMEM lane
    const bool issuer = ((e & 3) == 0) && lane == 0;
    const uint32_t bar_id = 1u + (uint32_t)half;
    // exchange area of the two column halves of a row; max / arg-max are double-buffered by tile parity because the
    // rollout variant has a single barrier per tile (a fast thread may already write tile i+1 while its partner reads i)
    float* ex_m_all = reinterpret_cast<float*>(tmem_slot + 4);  // [2 parity][2 halves][128] row maxima
    int* ex_i_all = reinterpret_cast<int*>(ex_m_all + 512);     // [2][2][128] first arg-max columns
    float* ex_s = reinterpret_cast<float*>(ex_i_all + 512);     // [2][128] exp-sums
    float* ex_t = ex_s + 256;                                   // [128] target logit
    const int xl = r % p.bw, yl = (r / p.bw) % p.bh, bl = r / (p.bw * p.bh);
    uint32_t ochunk = 0;
    float loss_acc = 0.f;
    int tcount = 0;
    for (int tile = pair; tile < total_tiles; tile += npairs, ++tcount) {
      const int nt = tile % p.n_tiles;  // colour
      const int mt = 2 * (tile / p.n_tiles) + (int)rank;
      const int tx = mt % p.tiles_x, ty = (mt / p.tiles_x) % p.tiles_y, tb = mt / (p.tiles_x * p.tiles_y);
      const int x0 = tx * p.bw, y0 = ty * p.bh, b0 = tb * p.bb, n0 = nt * BN;
      const int x = x0 + xl, y = y0 + yl, b = b0 + bl;
      const bool valid = (r < p.rows) && (x < p.Wo) && (y < p.Ho) && (b < p.B) && (mt < p.m_tiles);
      const long long toff = ((long long)b * 3 + nt) * ce.HW + (long long)y * p.Wo + x;  // u8 [B,3,HW] offset
      int t = 0;
      if (EPI == 1 && valid) t = ce.target[toff];  // issued before the wait: overlaps the MMA of this tile
      const int as = tcount & 1;
      const uint32_t aph = (uint32_t)(tcount >> 1) & 1u;
      mbar_wait(&tfull[as], aph);
      tc_fence_after();
      // ---- pass 1: my 128 logits -> +bias -> bf16 (the rounding of the materialised path), packed in 64 registers
      uint32_t xr[64];
      __nv_bfloat162 m2 = __floats2bfloat162_rn(-INFINITY, -INFINITY);
#pragma unroll
      for (int q4 = 0; q4 < 4; ++q4) {
        const int col = half * 64 + (q4 >> 1) * 128 + (q4 & 1) * 32;  // first column of this 32-column piece
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)(as * BN + col), v);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          const float4 bv = __ldg(reinterpret_cast<const float4*>(p.bias + n0 + col + j));
          const __nv_bfloat162 a0 = __floats2bfloat162_rn(__uint_as_float(v[j]) + bv.x, __uint_as_float(v[j + 1]) + bv.y);
          const __nv_bfloat162 a1 = __floats2bfloat162_rn(__uint_as_float(v[j + 2]) + bv.z, __uint_as_float(v[j + 3]) + bv.w);
          xr[q4 * 16 + (j >> 1)] = *reinterpret_cast<const uint32_t*>(&a0);
          xr[q4 * 16 + (j >> 1) + 1] = *reinterpret_cast<const uint32_t*>(&a1);
          m2 = __hmax2(m2, __hmax2(a0, a1));
        }
      }
      // all TMEM reads of this tile are done: hand the accumulator stage back, the next tile's MMAs overlap the rest
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (rank == 0) mbar_arrive(&tempty[as]);
        else mbar_arrive_cluster(mapa_u32(&tempty[as], 0));
      }
      const float2 mf = __bfloat1622float2(m2);
      const float lm = fmaxf(mf.x, mf.y);
      // first column (in this thread's ascending column order) holding the local maximum
      int first = 256;
      {
        const __nv_bfloat162 g2 = __float2bfloat162_rn(lm);
#pragma unroll
        for (int k = 63; k >= 0; --k) {
          const uint32_t eq = __heq2_mask(*reinterpret_cast<const __nv_bfloat162*>(&xr[k]), g2);
          const int colk = half * 64 + (k >> 5) * 128 + ((k & 31) << 1);  // column of the low element of pair k
          if (eq & 0xffff0000u) first = colk + 1;
          if (eq & 0x0000ffffu) first = colk;
        }
      }
      float* ex_m = ex_m_all + (tcount & 1) * 256;
      int* ex_i = ex_i_all + (tcount & 1) * 256;
      ex_m[half * 128 + r] = lm;
      ex_i[half * 128 + r] = first;
      named_bar_sync(3, 256);
      const float om = ex_m[(half ^ 1) * 128 + r];
      const int oi = ex_i[(half ^ 1) * 128 + r];
      const float gm = fmaxf(lm, om);
      if (ce.amax && half == 0 && valid) {
        int best = first;
        if (om > lm || (om == lm && oi < first)) best = oi;
        ce.amax[toff] = (uint8_t)best;
      }
      if (EPI == 1) {
        // ---- pass 2: exp-sum (same fp32 formula as pixel_ce_kernel) and the target logit
        const float moff = -gm * LOG2E;
        float s = 0.f;
        const int tq = t - half * 64;                     // position of the target inside my columns, if it is mine
        const bool mine = ((t >> 6) & 1) == half;
        const int tk = mine ? (((t >> 7) << 5) + ((tq & 63) >> 1)) : -1;  // pair index holding the target
        float xt = 0.f;
#pragma unroll
        for (int k = 0; k < 64; ++k) {
          const float x0f = __uint_as_float(xr[k] << 16), x1f = __uint_as_float(xr[k] & 0xffff0000u);
          s += exp2f(fmaf(x0f, LOG2E, moff)) + exp2f(fmaf(x1f, LOG2E, moff));
          if (k == tk) xt = (t & 1) ? x1f : x0f;
        }
        ex_s[half * 128 + r] = s;
        if (mine) ex_t[r] = xt;
        named_bar_sync(3, 256);
        const float st = ex_s[r] + ex_s[128 + r];
        xt = ex_t[r];
        const float cev = gm + log2f(st) * LN2 - xt;
        if (half == 0 && valid) loss_acc += fmaxf(cev, ce.clip);
        const float scale = (valid && cev > ce.clip) ? ce.gscale : 0.f;
        const float inv = scale / st;
        // ---- pass 3: gradient tile -> bf16 -> swizzled staging -> TMA store (two 64-column chunks per thread)
        const int sw = r & 7;
#pragma unroll
        for (int cc = 0; cc < 2; ++cc) {  // (unrolled: `xr` must stay in registers)
          const int c = half * 64 + cc * 128;
          if (issuer) bulk_wait_read<1>();  // the store that used this staging buffer two chunks ago has drained
          named_bar_sync(bar_id, 128);
          uint8_t* buf = sO + (half * 2 + (int)(ochunk & 1u)) * kStageOut2Bytes;
          uint8_t* row = buf + r * 128;
          if (r < p.rows) {
#pragma unroll
            for (int pc = 0; pc < 8; ++pc) {  // 16-byte pieces of 8 columns
              uint32_t o[4];
#pragma unroll
              for (int u = 0; u < 4; ++u) {
                const uint32_t w = xr[cc * 32 + pc * 4 + u];
                const float d0 = exp2f(fmaf(__uint_as_float(w << 16), LOG2E, moff)) * inv;
                const float d1 = exp2f(fmaf(__uint_as_float(w & 0xffff0000u), LOG2E, moff)) * inv;
                const __nv_bfloat162 h2 = __floats2bfloat162_rn(d0, d1);
                o[u] = *reinterpret_cast<const uint32_t*>(&h2);
              }
              *reinterpret_cast<uint4*>(row + ((pc ^ sw) << 4)) = make_uint4(o[0], o[1], o[2], o[3]);
            }
            if (mine && (t >> 7) == cc) {  // one-hot term: patch the target element (same thread: ordered after its vector store)
              const int tl = t & 63;
              const float et = exp2f(fmaf(xt, LOG2E, moff));
              *reinterpret_cast<__nv_bfloat16*>(row + (((tl >> 3) ^ sw) << 4) + ((tl & 7) << 1)) =
                  __float2bfloat16_rn(et * inv - scale);
            }
          }
          fence_proxy_async_smem();
          named_bar_sync(bar_id, 128);
          if (issuer) {
            tma_store_4d(&tmO, buf, n0 + c, x0, y0, b0);
            bulk_commit();
          }
          ++ochunk;
        }
      }
    }
    if (EPI == 1) {
      loss_acc = warp_sum(loss_acc);
      if (lane == 0 && half == 0 && loss_acc != 0.f) atomicAdd(ce.loss_sum, (double)loss_acc);
      if (issuer) bulk_wait_all();
    }
  } else {
    const int e = warp - 2;
    const int lg = warp & 3;
    const int half = e >> 2;
    const int r = lg * 32 + lane;
    const bool issuer = ((e & 3) == 0) && lane == 0;
    const uint32_t bar_id = 1u + (uint32_t)half;
    uint32_t ochunk = 0;
    int tcount = 0;
    for (int tile = pair; tile < total_tiles; tile += npairs, ++tcount) {
      const int nt = tile % p.n_tiles;
      const int mt = 2 * (tile / p.n_tiles) + (int)rank;
      const int tx = mt % p.tiles_x, ty = (mt / p.tiles_x) % p.tiles_y, tb = mt / (p.tiles_x * p.tiles_y);
      const int x0 = tx * p.bw, y0 = ty * p.bh, b0 = tb * p.bb, n0 = nt * BN;
      const int as = tcount & 1;
      const uint32_t aph = (uint32_t)(tcount >> 1) & 1u;
      mbar_wait(&tfull[as], aph);
      tc_fence_after();
#pragma unroll 1
      for (int c = half * 64; c < BN; c += 128) {  // 64-column chunks, interleaved between the two halves
        if (n0 + c >= p.N) break;
        if (issuer) bulk_wait_read<1>();  // the store that used this staging buffer two chunks ago has drained
        named_bar_sync(bar_id, 128);      // ... and every thread of the half knows it
        uint8_t* buf = sO + (half * 2 + (int)(ochunk & 1u)) * kStageOut2Bytes;
        uint8_t* row = buf + r * 128;
        const int sw = r & 7;
#pragma unroll
        for (int sub = 0; sub < 2; ++sub) {
          const int nb = n0 + c + sub * 32;
          if (nb >= p.N) break;  // ragged N (multiple of 32, not of 64): the TMA store clips the rest of the box
          uint32_t v[32];
          tmem_ld32(tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)(as * BN + c + sub * 32), v);
          tmem_ld_wait();
          float f[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
          if (p.bias) {
            const int nbm = nb % p.bias_mod;
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 bv = __ldg(reinterpret_cast<const float4*>(p.bias + nbm + j));
              f[j] += bv.x; f[j + 1] += bv.y; f[j + 2] += bv.z; f[j + 3] += bv.w;
            }
          }
          if (p.relu) {
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] = fmaxf(f[j], 0.f);
          }
          if (r < p.rows) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              uint4 pk;
              __nv_bfloat162 t0 = __floats2bfloat162_rn(f[8 * j], f[8 * j + 1]);
              __nv_bfloat162 t1 = __floats2bfloat162_rn(f[8 * j + 2], f[8 * j + 3]);
              __nv_bfloat162 t2 = __floats2bfloat162_rn(f[8 * j + 4], f[8 * j + 5]);
              __nv_bfloat162 t3 = __floats2bfloat162_rn(f[8 * j + 6], f[8 * j + 7]);
              pk.x = *reinterpret_cast<uint32_t*>(&t0);
              pk.y = *reinterpret_cast<uint32_t*>(&t1);
              pk.z = *reinterpret_cast<uint32_t*>(&t2);
              pk.w = *reinterpret_cast<uint32_t*>(&t3);
              *reinterpret_cast<uint4*>(row + (((sub * 4 + j) ^ sw) << 4)) = pk;  // 128B swizzle: chunk ^ (row & 7)
            }
          }
        }
        fence_proxy_async_smem();
        named_bar_sync(bar_id, 128);  // all 128 rows staged
        if (issuer) {
          tma_store_4d(&tmO, buf, n0 + c, x0, y0, b0);  // rows / channels outside the tensor are clipped
          bulk_commit();
        }
        ++ochunk;
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (rank == 0) mbar_arrive(&tempty[as]);
        else mbar_arrive_cluster(mapa_u32(&tempty[as], 0));
      }
    }
    if (issuer) bulk_wait_all();
  }
  tc_fence_before();
  cluster_sync_all();  // no CTA retires (or frees TMEM) while its peer may still signal it or read its operands
  if (warp == 1) tmem_dealloc_2cta(tmem_base, TMEM_COLS);
}

template <int BN, int STAGES, int EPI = 0>
static int launch_tap_gemm2(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmO, const TapGemmDev& p,
                            cudaStream_t stream, const CeParams* cep = nullptr) {
  constexpr int smem = STAGES * (kATileBytes + (BN / 2) * 128) + kOutBufs * kStageOut2Bytes + (2 * STAGES + 4) * 8 + 16 +
                       (EPI ? kCeExchangeBytes : 0) + 1024;
  static_assert(smem <= 232448, "tap_gemm2_kernel: shared memory budget exceeded");
  static bool configured = false;
  if (!configured) {
    if (cudaFuncSetAttribute(tap_gemm2_kernel<BN, STAGES, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) !=
        cudaSuccess)
      return SB_ERR_CUDA;
    configured = true;
  }
  const int total = ((p.m_tiles + 1) / 2) * p.n_tiles;
  const int max_pairs = (148 - g_sm_reserve) / 2;
  const int pairs = total < max_pairs ? total : max_pairs;
  CeParams ce{};
  if (cep) ce = *cep;
  tap_gemm2_kernel<BN, STAGES, EPI><<<2 * pairs, kFwdThreads, smem, stream>>>(tmA, tmB, tmO, p, ce);
  ++g_launches;
  SB_CHECK_LAUNCH();
  return SB_OK;
}

// ------------------------------------------------------------------------------------------------------------------
// wgrad: D[m = dY channel (128/tile), n = A channel (BN/tile)] += sum over a 64-pixel box; both operands MN-major.
// ------------------------------------------------------------------------------------------------------------------
struct TapWgradDev {
  int tiles_x, tiles_y, tiles_b;
  int bw, bh, bb;
  int C, N, ntaps, c_tiles, m_tiles;
  int tap_dy[SB_MAX_TAPS], tap_dx[SB_MAX_TAPS];
  float* dw;
  int ldw, splits;
  // Orientation.  Normal: M side (128-row tiles) = dY channels, N side (BN columns) = A channels, shifted by the tap.
  // Swapped: M side = A channels (shifted), N side = dY channels -- chosen by the launcher when it pads less (a 192 x 384
  // gradient wastes 44 % of a 256 x 512 tiling but none of a 384 x 192 one).  `C`/`N` above always name the N-side / M-side
  // channel counts of the launch; Ctrue is the A-channel count that addresses dw.
  int swapped, Ctrue;
  // Optional bias gradient (normal orientation only): dbias[m] += sum over pixels of dY[pixel, m].  It rides on the
  // tensor core: one more MMA per K step multiplies the dY tile (already in shared memory) with a constant tile of ones
  // into 16 spare TMEM columns, so the column sums cost no extra pass over dY (the logits gradient is 826 MB).
  float* dbias;
};
constexpr int kWgOnesBytes = 2048;            // bf16 ones: 16 pixels x 128-byte rows, read as an MN-major N = 16 operand
constexpr int kWgRows = 64;                   // pixels per stage (K of the contraction)
constexpr int kWgChunkBytes = kWgRows * 128;  // 64 pixels x 64 channels bf16

// TG = taps per CTA: one dY tile per stage feeds TG shifted A tiles and TG accumulators (TMEM columns j*BN).  A 3x3
// layer with few channels (the gate conv: 128 x 80 per tap) is bound by re-reading dY once per tap; TG = 3 reads it 3x
// instead of 9x.  TG > 1 needs the normal orientation (the shift is on the N-side operand).
// MT = 128-row M tiles per CTA.  The kernel is bound by the TMA unit's box-ROW rate (every loaded row is one pixel x 64
// channels = 128 bytes, ~3.5 clk per row per SM), not by bytes: at MT = 1 a stage of (2 + BN/64) chunks x 64 rows feeds
// 2*128*BN*64 flops, i.e. 10.9 kflop per row at BN = 256 -- 40 % of the tensor peak, which is what ncu shows (42.8 %
// tensor-pipe active).  MT = 2 keeps TWO dY tiles per stage against the same A chunks (two TMEM accumulators): 16.4 kflop
// per loaded row.
template <int BN, int STAGES, int TG, int MT>
__global__ void __launch_bounds__(kGemmThreads)
    tap_wgrad_kernel(const __grid_constant__ CUtensorMap tmDY, const __grid_constant__ CUtensorMap tmA,
                     const __grid_constant__ TapWgradDev p) {
  constexpr int NCH = BN / 64;
  constexpr int DY_BYTES = MT * 2 * kWgChunkBytes;
  constexpr int A_BYTES = NCH * kWgChunkBytes;  // per tap
  constexpr bool BIAS_ROOM = MT * TG * BN + 16 * MT <= 512;  // 16 spare TMEM columns per M tile for the bias gradient
  constexpr int ACC_COLS = MT * TG * BN + (BIAS_ROOM ? 16 * MT : 0);
  constexpr int TMEM_COLS = ACC_COLS <= 64 ? 64 : ACC_COLS <= 128 ? 128 : ACC_COLS <= 256 ? 256 : 512;
  static_assert(BN % 64 == 0, "N-side tile = whole 64-channel chunks");
  static_assert(ACC_COLS <= 512, "accumulators of one tap group must fit TMEM");
  static_assert(MT == 1 || TG == 1, "tap groups and multiple M tiles are not combined");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sDY = smem;
  uint8_t* sA = smem + STAGES * DY_BYTES;
  uint8_t* sOnes = sA + STAGES * TG * A_BYTES;
  uint64_t* full = reinterpret_cast<uint64_t*>(sOnes + kWgOnesBytes);
  uint64_t* empty = full + STAGES;
  uint64_t* tmem_full = empty + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int w = blockIdx.x;
  const int ct = w % p.c_tiles;
  w /= p.c_tiles;
  const int groups = (p.ntaps + TG - 1) / TG;
  const int t = (w % groups) * TG;             // first tap of this CTA's group
  const int ntg = min(TG, p.ntaps - t);        // taps in the group
  const int mtile = w / groups;
  const int m0 = mtile * 128 * MT, c0 = ct * BN;
  const int PT = p.tiles_x * p.tiles_y * p.tiles_b;
  const int pt_begin = (int)((long long)PT * blockIdx.y / p.splits);
  const int pt_end = (int)((long long)PT * (blockIdx.y + 1) / p.splits);
  const int nk = pt_end - pt_begin;
  const bool do_bias = BIAS_ROOM && p.dbias != nullptr && !p.swapped && ct == 0 && t == 0;
  if (do_bias) {
    for (int i = threadIdx.x; i < kWgOnesBytes / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(sOnes)[i] = 0x3F803F80u;
    fence_proxy_async_smem();  // generic-proxy writes -> visible to the tensor core's shared-memory reads
  }

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    mbar_init(tmem_full, 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      tma_prefetch_desc(&tmDY);
      tma_prefetch_desc(&tmA);
      const int dy = p.tap_dy[t], dx = p.tap_dx[t];
      for (int i = 0; i < nk; ++i) {
        const int s = i % STAGES;
        const uint32_t ph = (uint32_t)(i / STAGES) & 1u;
        mbar_wait(&empty[s], ph ^ 1u);
        const int pt = pt_begin + i;
        const int x0 = (pt % p.tiles_x) * p.bw;
        const int y0 = ((pt / p.tiles_x) % p.tiles_y) * p.bh;
        const int b0 = (pt / (p.tiles_x * p.tiles_y)) * p.bb;
        mbar_expect_tx(&full[s], (uint32_t)(DY_BYTES + ntg * A_BYTES));
        const int mx = (TG == 1 && p.swapped) ? dx : 0, my = (TG == 1 && p.swapped) ? dy : 0;  // M-side pixel shift
#pragma unroll
        for (int j = 0; j < 2 * MT; ++j)  // (channels past the tensor are zero-filled: an odd last M tile costs nothing but time)
          tma_load_4d(sDY + s * DY_BYTES + j * kWgChunkBytes, &tmDY, &full[s], m0 + j * 64, x0 + mx, y0 + my, b0);
#pragma unroll
        for (int g = 0; g < TG; ++g) {
          if (g >= ntg) break;
          // pixel shift of the N-side operand (normal orientation)
          const int nx = (TG == 1 && p.swapped) ? 0 : p.tap_dx[t + g], ny = (TG == 1 && p.swapped) ? 0 : p.tap_dy[t + g];
#pragma unroll
          for (int j = 0; j < NCH; ++j)
            tma_load_4d(sA + (s * TG + g) * A_BYTES + j * kWgChunkBytes, &tmA, &full[s], c0 + j * 64, x0 + nx, y0 + ny,
                        b0);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // A-channel (N) extent of this tile rounded to the MMA granularity: C = 80 runs N = 80, not the full 128
      const int n_eff = min(BN, ((p.C - c0 + 15) >> 4) << 4);
      const uint32_t idesc = make_idesc_bf16(128, n_eff, 1, 1);
      for (int i = 0; i < nk; ++i) {
        const int s = i % STAGES;
        const uint32_t ph = (uint32_t)(i / STAGES) & 1u;
        mbar_wait(&full[s], ph);
        tc_fence_after();
        const uint32_t a_addr = smem_u32(sDY + s * DY_BYTES);
#pragma unroll
        for (int g = 0; g < TG; ++g) {
          if (g >= ntg) break;
          const uint32_t b_addr = smem_u32(sA + (s * TG + g) * A_BYTES);
#pragma unroll
          for (int mi = 0; mi < MT; ++mi) {
#pragma unroll
            for (int k = 0; k < kWgRows / 16; ++k) {
              // MN-major, SW128: LBO = byte stride between 64-channel chunks, SBO = byte stride between 8-pixel groups
              const uint64_t ad = make_smem_desc_sw128(a_addr + mi * 2 * kWgChunkBytes + k * 16 * 128, kWgChunkBytes, 1024);
              const uint64_t bd = make_smem_desc_sw128(b_addr + k * 16 * 128, kWgChunkBytes, 1024);
              umma_bf16(tmem_base + (uint32_t)((mi * TG + g) * BN), ad, bd, idesc, (uint32_t)((i | k) != 0));
            }
          }
        }
        if (do_bias) {  // D[128 x 16] += dY^T (128 x 16 pixels) * ones (16 pixels x 16): every column = sum over the pixels
          const uint32_t idesc1 = make_idesc_bf16(128, 16, 1, 1);
#pragma unroll
          for (int mi = 0; mi < MT; ++mi) {
#pragma unroll
            for (int k = 0; k < kWgRows / 16; ++k) {
              const uint64_t ad = make_smem_desc_sw128(a_addr + mi * 2 * kWgChunkBytes + k * 16 * 128, kWgChunkBytes, 1024);
              const uint64_t od = make_smem_desc_sw128(smem_u32(sOnes), kWgChunkBytes, 1024);
              umma_bf16(tmem_base + (uint32_t)(MT * TG * BN + mi * 16), ad, od, idesc1, (uint32_t)((i | k) != 0));
            }
          }
        }
        umma_commit(&empty[s]);
      }
      umma_commit(tmem_full);
    }
  } else {
    mbar_wait(tmem_full, 0);
    tc_fence_after();
    const int lg = warp & 3;
#pragma unroll 1
    for (int mi = 0; mi < MT; ++mi) {
      const int m = m0 + mi * 128 + lg * 32 + lane;
      const bool valid = m < p.N;
      if (do_bias) {
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)(MT * TG * BN), v);  // identical column sums
        tmem_ld_wait();
        if (valid) atomicAdd(p.dbias + m, __uint_as_float(v[mi * 16]));
      }
      // normal: dw[m][t*C + c]; swapped (row = A channel, column = dY channel): dw[column][t*Ctrue + row], the 32 lanes
      // of a warp still hit 32 consecutive floats
#pragma unroll 1
      for (int gc = 0; gc < ntg * BN; gc += 32) {
        const int g = gc / BN, c = gc - g * BN;  // tap of the group, column inside its accumulator
        if (c0 + c >= p.C) continue;
        float* orow = p.swapped ? p.dw + (long long)(t + g) * p.Ctrue + m
                                : p.dw + (long long)m * p.ldw + (long long)(t + g) * p.C + c0;
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)(mi * TG * BN + gc), v);
        tmem_ld_wait();
        if (!valid) continue;
        if (p.swapped) {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (c0 + c + j < p.C) atomicAdd(orow + (long long)(c0 + c + j) * p.ldw, __uint_as_float(v[j]));
        } else {
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            if (c0 + c + j < p.C)
              red_add_v4(orow + c + j, __uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]),
                         __uint_as_float(v[j + 3]));
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, TMEM_COLS);
}

template <int BN, int STAGES, int TG = 1, int MT = 1>
static int launch_tap_wgrad(const CUtensorMap& tmDY, const CUtensorMap& tmA, const TapWgradDev& p, dim3 grid,
                            cudaStream_t stream) {
  constexpr int smem = STAGES * (2 * MT + TG * (BN / 64)) * kWgChunkBytes + kWgOnesBytes + (2 * STAGES + 1) * 8 + 16 + 1024;
  static_assert(smem <= 232448, "tap_wgrad_kernel: shared memory budget exceeded");
  static bool configured = false;
  if (!configured) {
    if (cudaFuncSetAttribute(tap_wgrad_kernel<BN, STAGES, TG, MT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) !=
        cudaSuccess)
      return SB_ERR_CUDA;
    configured = true;
  }
  tap_wgrad_kernel<BN, STAGES, TG, MT><<<grid, kGemmThreads, smem, stream>>>(tmDY, tmA, p);
  ++g_launches;
  SB_CHECK_LAUNCH();
  return SB_OK;
}

}  // namespace sb

using namespace sb;

extern "C" int sb_tap_gemm(const sb_tap_gemm_args* a, sb_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!a || !a->a || !a->w || !a->out) return SB_ERR_ARG;
  if (a->ntaps < 1 || a->ntaps > SB_MAX_TAPS) return SB_ERR_ARG;
  const int rows = a->bw * a->bh * a->bb;
  if (rows < 1 || rows > 128 || a->bw > 256 || a->bh > 256 || a->bb > 256) return SB_ERR_ARG;
  if (a->N % 8 || a->N < 8) return SB_ERR_ARG;
  const bool atomic = a->splits > 1 || a->accumulate;
  if (atomic && !a->out_f32) return SB_ERR_ARG;
  if (a->ldo % (a->out_f32 ? 4 : 8)) return SB_ERR_ARG;
  if (a->ldw < a->ntaps * a->C) return SB_ERR_ARG;
  int rc = load_driver_entry();
  if (rc) return rc;
  TapGemmDev p;
  p.B = a->B; p.Ho = a->Ho; p.Wo = a->Wo;
  p.bw = a->bw; p.bh = a->bh; p.bb = a->bb; p.rows = rows;
  p.tiles_x = (a->Wo + a->bw - 1) / a->bw;
  p.tiles_y = (a->Ho + a->bh - 1) / a->bh;
  const int tiles_b = (a->B + a->bb - 1) / a->bb;
  p.C = a->C; p.kchunks = (a->C + 63) / 64; p.ntaps = a->ntaps;
  for (int i = 0; i < SB_MAX_TAPS; ++i) {
    p.tap_dy[i] = i < a->ntaps ? a->tap_dy[i] : 0;
    p.tap_dx[i] = i < a->ntaps ? a->tap_dx[i] : 0;
  }
  p.N = a->N;
  const int KT = p.ntaps * p.kchunks;
  p.splits = a->splits < 1 ? 1 : (a->splits > KT ? KT : a->splits);
  p.out = a->out; p.out_f32 = a->out_f32; p.ldo = a->ldo;
  p.bias = a->bias; p.relu = a->relu; p.atomic = atomic ? 1 : 0;
  p.bias_mod = a->bias_mod > 0 ? a->bias_mod : a->N;
  if (a->bias && (p.bias_mod % 32 || a->N % p.bias_mod)) return SB_ERR_ARG;

  // N tile: 256 when it divides N (fewer A re-reads), else 192 / 128 / 96
  int BN;
  if (a->N % 256 == 0) BN = 256;
  else if (a->N % 192 == 0) BN = 192;
  else if (a->N % 128 == 0) BN = 128;
  else if (a->N <= 96) BN = 96;
  else BN = 128;
  CUtensorMap tmA, tmB, tmO;
  rc = make_map_nhwc(&tmA, a->a, a->B, a->H, a->W, a->C, a->bw, a->bh, a->bb);
  if (rc) return rc;
  rc = make_map_2d(&tmB, a->w, a->N, a->ntaps * a->C, a->ldw, BN);
  if (rc) return rc;
  // TMA-store epilogue: bf16 outputs; a ragged last 32-column chunk is clipped by the TMA unit (the vectorised bias loads
  // of the epilogue would read past the end, so ragged N is only taken without a bias: dgrad)
  p.tma_out = (!atomic && !a->out_f32 && (a->N % 32 == 0 || a->bias == nullptr)) ? 1 : 0;
  if (p.tma_out) {
    rc = make_map_out(&tmO, a->out, a->B, a->Ho, a->Wo, a->N, a->ldo, a->bw, a->bh, a->bb);
    if (rc) return rc;
  } else {
    tmO = tmA;  // unused
  }
  p.m_tiles = p.tiles_x * p.tiles_y * tiles_b;
  p.n_tiles = (a->N + BN - 1) / BN;
  // CTA-pair kernel: bf16 TMA-store outputs, no split-K, enough 256-row tiles to fill the 74 SM pairs, full N tiles
  // (N = 96 tiles measured slower on the pair kernel: 181 vs 162 us for the logits dgrad, which is HBM-bound anyway)
  const bool pair_ok = p.tma_out && p.splits == 1 && a->N % BN == 0 && BN != 96 && rows >= 64 &&
                       ((p.m_tiles + 1) / 2) * p.n_tiles >= 74 && g_pair_mode != 0;
  if (pair_ok) {
    CUtensorMap tmBh, tmO64;
    rc = make_map_2d(&tmBh, a->w, a->N, a->ntaps * a->C, a->ldw, BN / 2);
    if (rc) return rc;
    rc = make_map_out64(&tmO64, a->out, a->B, a->Ho, a->Wo, a->N, a->ldo, a->bw, a->bh, a->bb);
    if (rc) return rc;
    switch (BN) {
      case 256: return launch_tap_gemm2<256, 5>(tmA, tmBh, tmO64, p, stream);
      case 192: return launch_tap_gemm2<192, 5>(tmA, tmBh, tmO64, p, stream);
      case 128: return launch_tap_gemm2<128, 6>(tmA, tmBh, tmO64, p, stream);
      default: return launch_tap_gemm2<96, 6>(tmA, tmBh, tmO64, p, stream);
    }
  }
  switch (BN) {
    case 256: return launch_tap_gemm<256, 4>(tmA, tmB, tmO, p, stream);
    case 192: return launch_tap_gemm<192, 4>(tmA, tmB, tmO, p, stream);
    case 128: return launch_tap_gemm<128, 6>(tmA, tmB, tmO, p, stream);
    default: return launch_tap_gemm<96, 6>(tmA, tmB, tmO, p, stream);
  }
}

// Logits layer fused with the per-pixel softmax cross-entropy (see the EPI = 1 / 2 epilogue of tap_gemm2_kernel).
extern "C" int sb_logits_ce(const sb_tap_gemm_args* a, const uint8_t* target, float clip, float gscale,
                            uint8_t* argmax_out, double* loss_sum, sb_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!a || !a->a || !a->w || !a->bias) return SB_ERR_ARG;
  const bool train = a->out != nullptr;
  if (train && (!target || !loss_sum || a->out_f32)) return SB_ERR_ARG;
  if (!train && !argmax_out) return SB_ERR_ARG;
  if (a->N != 768 || a->ntaps != 1 || a->tap_dy[0] != 0 || a->tap_dx[0] != 0 || a->C % 8 || a->ldw < a->C) return SB_ERR_ARG;
  if (a->H != a->Ho || a->W != a->Wo || (train && (a->ldo != 768))) return SB_ERR_ARG;
  const int rows = a->bw * a->bh * a->bb;
  if (rows < 1 || rows > 128 || a->bw > 256 || a->bh > 256 || a->bb > 256) return SB_ERR_ARG;
  if ((long long)a->B * a->Ho * a->Wo * 3 >= (1LL << 31)) return SB_ERR_ARG;
  int rc = load_driver_entry();
  if (rc) return rc;
  TapGemmDev p;
  p.B = a->B; p.Ho = a->Ho; p.Wo = a->Wo;
  p.bw = a->bw; p.bh = a->bh; p.bb = a->bb; p.rows = rows;
  p.tiles_x = (a->Wo + a->bw - 1) / a->bw;
  p.tiles_y = (a->Ho + a->bh - 1) / a->bh;
  const int tiles_b = (a->B + a->bb - 1) / a->bb;
  p.C = a->C; p.kchunks = (a->C + 63) / 64; p.ntaps = 1;
  for (int i = 0; i < SB_MAX_TAPS; ++i) p.tap_dy[i] = p.tap_dx[i] = 0;
  p.N = 768; p.splits = 1;
  p.out = a->out; p.out_f32 = 0; p.ldo = 768;
  p.bias = a->bias; p.bias_mod = 768; p.relu = 0; p.atomic = 0; p.tma_out = 1;
  p.m_tiles = p.tiles_x * p.tiles_y * tiles_b;
  p.n_tiles = 3;
  CUtensorMap tmA, tmBh, tmO64;
  rc = make_map_nhwc(&tmA, a->a, a->B, a->H, a->W, a->C, a->bw, a->bh, a->bb);
  if (rc) return rc;
  rc = make_map_2d(&tmBh, a->w, 768, a->C, a->ldw, 128);
  if (rc) return rc;
  if (train) {
    rc = make_map_out64(&tmO64, a->out, a->B, a->Ho, a->Wo, 768, 768, a->bw, a->bh, a->bb);
    if (rc) return rc;
  } else {
    tmO64 = tmA;  // unused
  }
  CeParams ce;
  ce.target = target; ce.amax = argmax_out; ce.loss_sum = loss_sum; ce.clip = clip; ce.gscale = gscale;
  ce.HW = a->Ho * a->Wo;
  if (train) return launch_tap_gemm2<256, 4, 1>(tmA, tmBh, tmO64, p, stream, &ce);
  return launch_tap_gemm2<256, 4, 2>(tmA, tmBh, tmO64, p, stream, &ce);
}

// Tile the [dY channels N] x [A channels C] gradient: least padded MMA work; ties: fewer tiles (fewer re-reads of the M
// side), then the smaller tile.
static int wgrad_pad_n(int ch, int* bn_out) {
  int best = 1 << 30, best_tiles = 1 << 30, bn_best = 64;
  const int opts[4] = {64, 128, 192, 256};
  for (int bn : opts) {
    const int full = ch / bn, rem = ch - full * bn;
    const int cost = full * bn + ((rem + 15) / 16) * 16, tiles = full + (rem ? 1 : 0);
    if (cost < best || (cost == best && tiles < best_tiles)) {
      best = cost;
      best_tiles = tiles;
      bn_best = bn;
    }
  }
  *bn_out = bn_best;
  return best;
}
// 1: run with the A channels on the 128-row M side (the swapped epilogue issues scalar atomics: only for a clear win)
static int wgrad_swapped(int N, int C, int* bn_norm, int* bn_swap) {
  const long long cost_norm = (long long)((N + 127) / 128) * 128 * wgrad_pad_n(C, bn_norm);
  const long long cost_swap = (long long)((C + 127) / 128) * 128 * wgrad_pad_n(N, bn_swap);
  return (cost_swap * 100 < cost_norm * 85) ? 1 : 0;
}
extern "C" int sb_tap_wgrad_bias_ok(int N, int C) {
  int a, b;
  return wgrad_swapped(N, C, &a, &b) ? 0 : 1;
}

extern "C" int sb_tap_wgrad(const sb_tap_wgrad_args* a, sb_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!a || !a->a || !a->dy || !a->dw) return SB_ERR_ARG;
  if (a->ntaps < 1 || a->ntaps > SB_MAX_TAPS) return SB_ERR_ARG;
  if (a->bw * a->bh * a->bb != kWgRows) return SB_ERR_ARG;
  if (a->ldw % 4 || a->ldw < a->ntaps * a->C || a->C % 8 || a->N % 8) return SB_ERR_ARG;
  int rc = load_driver_entry();
  if (rc) return rc;
  TapWgradDev p;
  p.bw = a->bw; p.bh = a->bh; p.bb = a->bb;
  p.tiles_x = (a->Wo + a->bw - 1) / a->bw;
  p.tiles_y = (a->Ho + a->bh - 1) / a->bh;
  p.tiles_b = (a->B + a->bb - 1) / a->bb;
  p.C = a->C; p.N = a->N; p.ntaps = a->ntaps;
  for (int i = 0; i < SB_MAX_TAPS; ++i) {
    p.tap_dy[i] = i < a->ntaps ? a->tap_dy[i] : 0;
    p.tap_dx[i] = i < a->ntaps ? a->tap_dx[i] : 0;
  }
  p.dw = a->dw; p.ldw = a->ldw;
  p.dbias = a->dbias;
  const int PT = p.tiles_x * p.tiles_y * p.tiles_b;
  // Tile the [dY channels N] x [A channels C] gradient: M side in 128-row tiles, N side in BN-column tiles
  // (BN in {64,128,192,256}; the last tile runs the MMA at the 16-rounded remainder).  Pick the orientation and BN that
  // waste the fewest tensor-core cycles.
  int bn_norm, bn_swap;
  p.swapped = wgrad_swapped(a->N, a->C, &bn_norm, &bn_swap);
  p.Ctrue = a->C;
  int BN;
  const void *m_ptr, *n_ptr;
  int mH, mW, mC, nH, nW, nC;
  if (!p.swapped) {
    BN = bn_norm;
    p.N = a->N; p.C = a->C;
    m_ptr = a->dy; mH = a->Ho; mW = a->Wo; mC = a->N;
    n_ptr = a->a; nH = a->H; nW = a->W; nC = a->C;
  } else {
    BN = bn_swap;
    p.N = a->C; p.C = a->N;  // kernel naming: N = M-side channels, C = N-side channels
    m_ptr = a->a; mH = a->H; mW = a->W; mC = a->C;
    n_ptr = a->dy; nH = a->Ho; nW = a->Wo; nC = a->N;
  }
  p.c_tiles = (p.C + BN - 1) / BN;
  p.m_tiles = (p.N + 127) / 128;
  // tap groups (3 taps share one dY tile): multi-tap layers with one narrow N-side tile and many pixels -- the gate conv
  const bool grouped = g_wgrad_tap_groups != 0 && !p.swapped && BN == 128 && p.c_tiles == 1 && p.ntaps % 3 == 0 &&
                       PT >= 1024;
  const int groups = grouped ? p.ntaps / 3 : p.ntaps;
  // two M tiles per CTA (see the kernel comment) when the M side has at least 256 channels; with a bias gradient
  // requested the 16 extra TMEM columns must still fit (BN <= 192)
  const bool mt2 = g_wgrad_mt2 != 0 && !grouped && p.m_tiles >= 2 && (p.dbias == nullptr || p.swapped || BN <= 192);
  if (mt2) p.m_tiles = (p.m_tiles + 1) / 2;
  const int ctas = p.m_tiles * groups * p.c_tiles;
  int splits = a->splits;
  static int slots = 0;
  if (!slots) {
    const char* e = getenv("SB_WGRAD_SLOTS");
    slots = e ? atoi(e) : 148;
    if (slots < 1) slots = 148;
  }
  // auto: ONE co-resident wave (one CTA per SM).  Measured over all 18 weight gradients of the training step: one wave
  // 1.17 ms, two waves 1.28 ms, four 1.53 ms -- longer K loops amortise the prologue and halve the fp32 atomics
  if (splits < 1) splits = (slots > g_sm_reserve ? slots - g_sm_reserve : 1) / ctas;
  p.splits = splits < 1 ? 1 : (splits > PT ? PT : splits);
  CUtensorMap tmDY, tmA;
  rc = make_map_nhwc(&tmDY, m_ptr, a->B, mH, mW, mC, a->bw, a->bh, a->bb);
  if (rc) return rc;
  rc = make_map_nhwc(&tmA, n_ptr, a->B, nH, nW, nC, a->bw, a->bh, a->bb);
  if (rc) return rc;
  dim3 grid(ctas, p.splits, 1);
  if (grouped) return launch_tap_wgrad<128, 3, 3>(tmDY, tmA, p, grid, stream);
  if (mt2) {
    switch (BN) {
      case 256: return launch_tap_wgrad<256, 3, 1, 2>(tmDY, tmA, p, grid, stream);
      case 192: return launch_tap_wgrad<192, 3, 1, 2>(tmDY, tmA, p, grid, stream);
      case 128: return launch_tap_wgrad<128, 4, 1, 2>(tmDY, tmA, p, grid, stream);
      default: return launch_tap_wgrad<64, 4, 1, 2>(tmDY, tmA, p, grid, stream);
    }
  }
  switch (BN) {
    case 256: return launch_tap_wgrad<256, 4>(tmDY, tmA, p, grid, stream);
    case 192: return launch_tap_wgrad<192, 4>(tmDY, tmA, p, grid, stream);
    case 128: return launch_tap_wgrad<128, 6>(tmDY, tmA, p, grid, stream);
    default: return launch_tap_wgrad<64, 6>(tmDY, tmA, p, grid, stream);
  }
}

extern "C" int sb_init(int device) {
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return SB_ERR_CUDA;
  if (prop.major != 10) return SB_ERR_ARCH;
  return load_driver_entry();
}
extern "C" long long sb_launch_count(void) { return sb::g_launches; }
extern "C" int sb_set_wgrad_tap_groups(int enabled) {
  const int prev = sb::g_wgrad_tap_groups;
  sb::g_wgrad_tap_groups = enabled ? 1 : 0;
  return prev;
}
// Leave `n` SMs free in every persistent / one-wave grid launched from now on.  Data-parallel training sets it for the
// backward pass: the gradient all-reduce of a finished bucket runs on a side stream meanwhile, and a CTA that finds its
// SM taken by a collective's CTA would otherwise run as a second wave behind the other 140 (measured: +37 % on the
// pair GEMMs that overlap an all-reduce).  Returns the previous value.
extern "C" int sb_set_sm_reserve(int n) {
  const int prev = sb::g_sm_reserve;
  sb::g_sm_reserve = n < 0 ? 0 : (n > 64 ? 64 : n);
  return prev;
}
extern "C" int sb_set_wgrad_mt2(int enabled) {
  const int prev = sb::g_wgrad_mt2;
  sb::g_wgrad_mt2 = enabled ? 1 : 0;
  return prev;
}
extern "C" int sb_set_pair_mode(int enabled) {
  const int prev = sb::g_pair_mode;
  sb::g_pair_mode = enabled ? 1 : 0;
  return prev;
}
extern "C" const char* sb_version(void) { return "simple_b200 0.1 (sm_100a)"; }
