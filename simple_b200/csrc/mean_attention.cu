// MeanAttention of the posterior encoder (simple/utils.py:72-88), forward and backward, one CTA per sample.
//
//   m[c]     = mean_hw x[c,hw]
//   a[k,hw]  = We[k,:] . x[:,hw] + be[k]                       (1x1 conv to 4 maps)
//   s        = softmax over dim 1 of a.view(B, HW, 4)          (reference quirk: the view reinterprets the [4,HW] buffer, so
//                                                               softmax group j holds the FLAT indices f = k*HW + hw with f % 4 == j)
//   am[c,k]  = mean_hw x[c,hw] * s[k,hw]
//   y        = Wd . cat(am, m)[c*5 + k] + bd
//
// The reference materialises x.unsqueeze(2) * s as a [B,C,4,H,W] fp32 tensor (68 MB at B = 64) in both passes and runs
// ~25 ATen kernels per call; here the sample's activations (<= 135 KB bf16) are staged once in shared memory and every
// intermediate lives on chip.  Input is the conv output as the tap GEMM wrote it: NHWC bf16 [B, HW, C].
#include "common.cuh"
#include "../../include/simple_b200.h"

namespace sb {
extern long long g_launches;

constexpr int kMaThreads = 512;

struct MaSmem {
  __nv_bfloat16* x;  // [HW][C + 2]  (row pitch C/2 + 1 words: per-row accesses by consecutive threads hit distinct banks)
  float* a;          // [4 * HW] flat (k * HW + hw): pre-softmax, then softmax, then its gradient
  float* ds;         // [4 * HW] (backward)
  float* l;          // [5 * C]: cat(am, m) laid out c*5 + k (forward) / its gradient (backward)
  float* we;         // [4][C]
  float* red;        // [32]
};
__device__ __forceinline__ MaSmem ma_carve(uint8_t* base, int HW, int C) {
  MaSmem m;
  m.x = reinterpret_cast<__nv_bfloat16*>(base);
  size_t off = ((size_t)HW * (C + 2) * 2 + 15) & ~(size_t)15;
  m.a = reinterpret_cast<float*>(base + off); off += (size_t)4 * HW * 4;
  m.ds = reinterpret_cast<float*>(base + off); off += (size_t)4 * HW * 4;
  m.l = reinterpret_cast<float*>(base + off); off += (size_t)5 * C * 4;
  m.we = reinterpret_cast<float*>(base + off); off += (size_t)4 * C * 4;
  m.red = reinterpret_cast<float*>(base + off);
  return m;
}
static size_t ma_smem_bytes(int HW, int C) {
  return (((size_t)HW * (C + 2) * 2 + 15) & ~(size_t)15) + (size_t)(8 * HW + 9 * C + 32) * 4;
}

__device__ __forceinline__ void ma_stage(const MaSmem& sm, const __nv_bfloat16* __restrict__ xg, const float* __restrict__ We,
                                         int HW, int C) {
  const int CV = C / 8;
  for (int i = threadIdx.x; i < HW * CV; i += kMaThreads) {
    const int hw = i / CV, cv = i - hw * CV;
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(xg + (size_t)hw * C + cv * 8));
    uint32_t* dst = reinterpret_cast<uint32_t*>(sm.x + (size_t)hw * (C + 2) + cv * 8);  // 4-byte aligned (pitch is even)
    dst[0] = v.x; dst[1] = v.y; dst[2] = v.z; dst[3] = v.w;
  }
  for (int i = threadIdx.x; i < 4 * C; i += kMaThreads) sm.we[i] = We[i];
}

// dot products of one pixel row with the four rows of a [4][C] fp32 matrix in shared memory
__device__ __forceinline__ void ma_row_dots(const MaSmem& sm, const float* w4, int hw, int C, float* out) {
  const __nv_bfloat162* row = reinterpret_cast<const __nv_bfloat162*>(sm.x + (size_t)hw * (C + 2));
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  for (int c2 = 0; c2 < C / 2; ++c2) {
    const float2 v = __bfloat1622float2(row[c2]);
    const int c = 2 * c2;
    a0 = fmaf(v.x, w4[c], fmaf(v.y, w4[c + 1], a0));
    a1 = fmaf(v.x, w4[C + c], fmaf(v.y, w4[C + c + 1], a1));
    a2 = fmaf(v.x, w4[2 * C + c], fmaf(v.y, w4[2 * C + c + 1], a2));
    a3 = fmaf(v.x, w4[3 * C + c], fmaf(v.y, w4[3 * C + c + 1], a3));
  }
  out[0] = a0; out[1] = a1; out[2] = a2; out[3] = a3;
}

// Same four dot products with the channels of the row split over the 32 lanes of a warp (few pixels, many channels: the
// posterior's second MeanAttention has HW = 30, C = 512 -- one thread per row would leave 480 of 512 threads idle).
__device__ __forceinline__ void ma_row_dots_warp(const MaSmem& sm, const float* w4, int hw, int C, int lane, float* out) {
  const __nv_bfloat162* row = reinterpret_cast<const __nv_bfloat162*>(sm.x + (size_t)hw * (C + 2));
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  for (int c2 = lane; c2 < C / 2; c2 += 32) {
    const float2 v = __bfloat1622float2(row[c2]);
    const int c = 2 * c2;
    a0 = fmaf(v.x, w4[c], fmaf(v.y, w4[c + 1], a0));
    a1 = fmaf(v.x, w4[C + c], fmaf(v.y, w4[C + c + 1], a1));
    a2 = fmaf(v.x, w4[2 * C + c], fmaf(v.y, w4[2 * C + c + 1], a2));
    a3 = fmaf(v.x, w4[3 * C + c], fmaf(v.y, w4[3 * C + c + 1], a3));
  }
  out[0] = warp_sum(a0); out[1] = warp_sum(a1); out[2] = warp_sum(a2); out[3] = warp_sum(a3);
}

__global__ void __launch_bounds__(kMaThreads)
    mean_attention_fwd_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ We,
                              const float* __restrict__ be, const float* __restrict__ Wd, const float* __restrict__ bd,
                              int HW, int C, int O, float* __restrict__ y, float* __restrict__ s_save,
                              float* __restrict__ l_save) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  const MaSmem sm = ma_carve(smem_raw, HW, C);
  const int b = blockIdx.x, t = threadIdx.x, lane = t & 31, warp = t >> 5;
  ma_stage(sm, x + (size_t)b * HW * C, We, HW, C);
  __syncthreads();
  if (HW * 4 < kMaThreads) {
    for (int hw = warp; hw < HW; hw += kMaThreads / 32) {
      float d[4];
      ma_row_dots_warp(sm, sm.we, hw, C, lane, d);
      if (lane < 4) sm.a[lane * HW + hw] = d[lane] + be[lane];
    }
  } else {
    for (int hw = t; hw < HW; hw += kMaThreads) {
      float d[4];
      ma_row_dots(sm, sm.we, hw, C, d);
#pragma unroll
      for (int k = 0; k < 4; ++k) sm.a[k * HW + hw] = d[k] + be[k];
    }
  }
  __syncthreads();
  if (warp < 4) {  // softmax group j = warp: flat indices f = 4 i + j
    const int j = warp;
    float mx = -3.0e38f;
    for (int i = lane; i < HW; i += 32) mx = fmaxf(mx, sm.a[4 * i + j]);
    mx = warp_max(mx);
    float s = 0.f;
    for (int i = lane; i < HW; i += 32) {
      const float e = expf(sm.a[4 * i + j] - mx);
      sm.a[4 * i + j] = e;
      s += e;
    }
    s = warp_sum(s);
    const float inv = 1.0f / s;
    for (int i = lane; i < HW; i += 32) sm.a[4 * i + j] *= inv;
  }
  __syncthreads();
  const float invP = 1.0f / (float)HW;
  for (int c = t; c < C; c += kMaThreads) {
    float acc[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
    for (int hw = 0; hw < HW; ++hw) {
      const float v = __bfloat162float(sm.x[(size_t)hw * (C + 2) + c]);
      acc[0] = fmaf(v, sm.a[hw], acc[0]);
      acc[1] = fmaf(v, sm.a[HW + hw], acc[1]);
      acc[2] = fmaf(v, sm.a[2 * HW + hw], acc[2]);
      acc[3] = fmaf(v, sm.a[3 * HW + hw], acc[3]);
      acc[4] += v;
    }
#pragma unroll
    for (int k = 0; k < 5; ++k) sm.l[c * 5 + k] = acc[k] * invP;
  }
  __syncthreads();
  for (int i = t; i < 4 * HW; i += kMaThreads) s_save[(size_t)b * 4 * HW + i] = sm.a[i];
  for (int i = t; i < 5 * C; i += kMaThreads) l_save[(size_t)b * 5 * C + i] = sm.l[i];
  for (int o = warp; o < O; o += kMaThreads / 32) {
    float acc = 0.f;
    const float* w = Wd + (size_t)o * 5 * C;
    for (int i = lane; i < 5 * C; i += 32) acc = fmaf(__ldg(w + i), sm.l[i], acc);
    acc = warp_sum(acc);
    if (lane == 0) y[(size_t)b * O + o] = acc + bd[o];
  }
}

__global__ void __launch_bounds__(kMaThreads)
    mean_attention_bwd_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ We,
                              const float* __restrict__ Wd, const float* __restrict__ s_save,
                              const float* __restrict__ dy, int HW, int C, int O, __nv_bfloat16* __restrict__ dx,
                              float* dWe_part, float* dbe_part, int accumulate) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  const MaSmem sm = ma_carve(smem_raw, HW, C);
  const int b = blockIdx.x, t = threadIdx.x, lane = t & 31, warp = t >> 5;
  ma_stage(sm, x + (size_t)b * HW * C, We, HW, C);
  for (int i = t; i < 4 * HW; i += kMaThreads) sm.a[i] = s_save[(size_t)b * 4 * HW + i];
  // dl[i] = sum_o dy[o] Wd[o,i]   (gradient of cat(am, m); consecutive threads read consecutive i: coalesced)
  for (int i = t; i < 5 * C; i += kMaThreads) {
    float acc = 0.f;
    for (int o = 0; o < O; ++o) acc = fmaf(dy[(size_t)b * O + o], __ldg(Wd + (size_t)o * 5 * C + i), acc);
    sm.l[i] = acc;
  }
  __syncthreads();
  const float invP = 1.0f / (float)HW;
  // ds[k,hw] = (1/HW) sum_c dam[c,k] x[hw,c]: gather d(am) into a [4][C] matrix first (reuse of the row-dot routine);
  // the matrix lives after `red` (the launcher reserves 4*C floats there)
  float* dam = sm.red + 32;
  for (int i = t; i < 4 * C; i += kMaThreads) {
    const int k = i / C, c = i - k * C;
    dam[i] = sm.l[c * 5 + k];
  }
  __syncthreads();
  if (HW * 4 < kMaThreads) {
    for (int hw = warp; hw < HW; hw += kMaThreads / 32) {
      float d[4];
      ma_row_dots_warp(sm, dam, hw, C, lane, d);
      if (lane < 4) sm.ds[lane * HW + hw] = d[lane] * invP;
    }
  } else {
    for (int hw = t; hw < HW; hw += kMaThreads) {
      float d[4];
      ma_row_dots(sm, dam, hw, C, d);
#pragma unroll
      for (int k = 0; k < 4; ++k) sm.ds[k * HW + hw] = d[k] * invP;
    }
  }
  __syncthreads();
  if (warp < 4) {  // softmax backward of group j = warp: da = s * (ds - sum s ds); stored over ds
    const int j = warp;
    float dot = 0.f;
    for (int i = lane; i < HW; i += 32) dot = fmaf(sm.a[4 * i + j], sm.ds[4 * i + j], dot);
    dot = warp_sum(dot);
    for (int i = lane; i < HW; i += 32) sm.ds[4 * i + j] = sm.a[4 * i + j] * (sm.ds[4 * i + j] - dot);
  }
  __syncthreads();
  // per-sample parts of the 1x1-conv gradients: dWe[k,c] = sum_hw da[k,hw] x[hw,c], dbe[k] = sum_hw da[k,hw]
  for (int c = t; c < C; c += kMaThreads) {
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int hw = 0; hw < HW; ++hw) {
      const float v = __bfloat162float(sm.x[(size_t)hw * (C + 2) + c]);
      acc[0] = fmaf(v, sm.ds[hw], acc[0]);
      acc[1] = fmaf(v, sm.ds[HW + hw], acc[1]);
      acc[2] = fmaf(v, sm.ds[2 * HW + hw], acc[2]);
      acc[3] = fmaf(v, sm.ds[3 * HW + hw], acc[3]);
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (accumulate) atomicAdd(dWe_part + (size_t)k * C + c, acc[k]);  // batch sum [4,C] (caller zeroes)
      else dWe_part[((size_t)b * 4 + k) * C + c] = acc[k];
    }
  }
  if (warp < 4) {
    float s = 0.f;
    for (int hw = lane; hw < HW; hw += 32) s += sm.ds[warp * HW + hw];
    s = warp_sum(s);
    if (lane == 0) {
      if (accumulate) atomicAdd(dbe_part + warp, s);
      else dbe_part[b * 4 + warp] = s;
    }
  }
  // dx[hw,c] = (dm[c] + sum_k dam[c,k] s[k,hw]) / HW + sum_k da[k,hw] We[k,c]
  const int CV = C / 8;
  for (int i = t; i < HW * CV; i += kMaThreads) {
    const int hw = i / CV, c0 = (i - hw * CV) * 8;
    float s4[4], d4[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      s4[k] = sm.a[k * HW + hw] * invP;
      d4[k] = sm.ds[k * HW + hw];
    }
    uint32_t o[4];
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) {
      float v[2];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int c = c0 + 2 * jj + h;
        float acc = sm.l[c * 5 + 4] * invP;
#pragma unroll
        for (int k = 0; k < 4; ++k) acc = fmaf(dam[k * C + c], s4[k], fmaf(d4[k], sm.we[k * C + c], acc));
        v[h] = acc;
      }
      const __nv_bfloat162 pk = __floats2bfloat162_rn(v[0], v[1]);
      o[jj] = *reinterpret_cast<const uint32_t*>(&pk);
    }
    *reinterpret_cast<uint4*>(dx + ((size_t)b * HW + hw) * C + c0) = make_uint4(o[0], o[1], o[2], o[3]);
  }
}

}  // namespace sb

using namespace sb;

extern "C" int sb_mean_attention_fwd(const void* x, const float* We, const float* be, const float* Wd, const float* bd,
                                     int B, int HW, int C, int O, float* y, float* s_save, float* l_save,
                                     sb_stream_t stream) {
  if (!x || !We || !be || !Wd || !bd || !y || !s_save || !l_save || B < 1 || HW < 1 || C % 8 || C < 8 || O < 1)
    return SB_ERR_ARG;
  const size_t smem = ma_smem_bytes(HW, C);
  if (smem > 227 * 1024) return SB_ERR_ARG;
  static size_t configured = 0;
  if (smem > configured) {
    if (cudaFuncSetAttribute(mean_attention_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) !=
        cudaSuccess)
      return SB_ERR_CUDA;
    configured = smem;
  }
  mean_attention_fwd_kernel<<<B, kMaThreads, smem, (cudaStream_t)stream>>>((const __nv_bfloat16*)x, We, be, Wd, bd, HW, C,
                                                                          O, y, s_save, l_save);
  ++g_launches;
  SB_CHECK_LAUNCH();
  return SB_OK;
}

extern "C" int sb_mean_attention_bwd(const void* x, const float* We, const float* Wd, const float* s_save,
                                     const float* dy, int B, int HW, int C, int O, void* dx, float* dWe_part,
                                     float* dbe_part, int accumulate, sb_stream_t stream) {
  if (!x || !We || !Wd || !s_save || !dy || !dx || !dWe_part || !dbe_part || B < 1 || HW < 1 || C % 8 || C < 8 || O < 1)
    return SB_ERR_ARG;
  const size_t smem = ma_smem_bytes(HW, C) + (size_t)4 * C * 4;  // + the [4][C] gather of d(am) after `red`
  if (smem > 227 * 1024) return SB_ERR_ARG;
  static size_t configured = 0;
  if (smem > configured) {
    if (cudaFuncSetAttribute(mean_attention_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) !=
        cudaSuccess)
      return SB_ERR_CUDA;
    configured = smem;
  }
  mean_attention_bwd_kernel<<<B, kMaThreads, smem, (cudaStream_t)stream>>>((const __nv_bfloat16*)x, We, Wd, s_save, dy, HW,
                                                                          C, O, (__nv_bfloat16*)dx, dWe_part, dbe_part, accumulate);
  ++g_launches;
  SB_CHECK_LAUNCH();
  return SB_OK;
}
