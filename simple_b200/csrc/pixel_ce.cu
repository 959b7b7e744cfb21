// Fused per-pixel 256-way softmax cross-entropy (+ clipping, gradient, argmax decode) -- K14 of SURVEY 2.3.
// Replaces simple/trainer.py:134-137 (CrossEntropyLoss(reduction='none') -> max(.,0.03) -> mean - 0.03) and its
// autograd backward, plus torch.argmax at simple/trainer.py:130 and simple/subproc_vec_env.py:428.
//
// Logits are bf16 [NPIX, 3*256] with the 768 channels permuted colour-major (col = colour*256 + value) so that one
// softmax group is 512 contiguous bytes: one warp reads one group with a single 16-byte load per lane, reduces with
// shuffles, and writes the gradient back with one 16-byte store per lane.  Every logit is read from HBM exactly once
// and the gradient written once (25.85 MB per frame).
#include "common.cuh"
#include "../../include/simple_b200.h"

namespace sb {
extern long long g_launches;

struct __align__(16) BF8 {
  __nv_bfloat162 v[4];
};

__device__ __forceinline__ uint4 ld_stream(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}

template <bool TRAIN>
__global__ void __launch_bounds__(256)
    pixel_ce_kernel(const __nv_bfloat16* __restrict__ logits, const uint8_t* __restrict__ target, int B, int HW,
                    float clip, float gscale, __nv_bfloat16* dlogits, uint8_t* __restrict__ argmax_out,
                    double* __restrict__ loss_sum, float* __restrict__ dbias) {
  const int lane = threadIdx.x & 31;
  const int warp_in_block = threadIdx.x >> 5;
  const long long npix = (long long)B * HW;
  const long long warp_global = (long long)blockIdx.x * (blockDim.x >> 5) + warp_in_block;
  const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
  __shared__ float s_db[768];
  __shared__ float s_loss[8];
  if (TRAIN && dbias) {
    for (int i = threadIdx.x; i < 768; i += blockDim.x) s_db[i] = 0.f;
    __syncthreads();
  }
  float db[24];
#pragma unroll
  for (int i = 0; i < 24; ++i) db[i] = 0.f;
  float loss_acc = 0.f;
  constexpr float LOG2E = 1.4426950408889634f, LN2 = 0.6931471805599453f;

  for (long long pix = warp_global; pix < npix; pix += nwarps) {
    const int b = (int)(pix / HW);
    const int hw = (int)(pix - (long long)b * HW);
    const __nv_bfloat16* row = logits + pix * 768;
    uint4 raw[3];
#pragma unroll
    for (int g = 0; g < 3; ++g) raw[g] = ld_stream(row + g * 256 + lane * 8);
#pragma unroll
    for (int g = 0; g < 3; ++g) {
      float x[8];
      {
        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&raw[g]);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float2 f = __bfloat1622float2(h[j]);
          x[2 * j] = f.x;
          x[2 * j + 1] = f.y;
        }
      }
      // local max + first index
      float m = x[0];
      int mi = 0;
#pragma unroll
      for (int j = 1; j < 8; ++j)
        if (x[j] > m) {
          m = x[j];
          mi = j;
        }
      mi += lane * 8;
      // warp argmax with first-index tie-break (torch.argmax semantics)
      float wm = m;
      int wi = mi;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float om = __shfl_xor_sync(0xffffffffu, wm, o);
        const int oi = __shfl_xor_sync(0xffffffffu, wi, o);
        if (om > wm || (om == wm && oi < wi)) {
          wm = om;
          wi = oi;
        }
      }
      if (argmax_out && lane == 0) argmax_out[((long long)b * 3 + g) * HW + hw] = (uint8_t)wi;
      if (TRAIN) {
        float e[8];
        float s = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          e[j] = exp2f((x[j] - wm) * LOG2E);
          s += e[j];
        }
        s = warp_sum(s);
        const int t = target[((long long)b * 3 + g) * HW + hw];
        // value of the target logit: owned by lane t/8, slot t%8
        float xt = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) xt = ((t & 7) == j) ? x[j] : xt;
        xt = __shfl_sync(0xffffffffu, xt, t >> 3);
        const float ce = wm + log2f(s) * LN2 - xt;
        loss_acc += (lane == 0) ? fmaxf(ce, clip) : 0.f;
        const float scale = (ce > clip) ? gscale : 0.f;
        const float inv = scale / s;
        float d[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) d[j] = e[j] * inv;
        if ((t >> 3) == lane) {
#pragma unroll
          for (int j = 0; j < 8; ++j) d[j] -= ((t & 7) == j) ? scale : 0.f;
        }
        BF8 o;
#pragma unroll
        for (int j = 0; j < 4; ++j) o.v[j] = __floats2bfloat162_rn(d[2 * j], d[2 * j + 1]);
        *reinterpret_cast<BF8*>(dlogits + pix * 768 + g * 256 + lane * 8) = o;
#pragma unroll
        for (int j = 0; j < 8; ++j) db[g * 8 + j] += d[j];
      }
    }
  }
  if (TRAIN) {
    if (dbias) {
#pragma unroll
      for (int g = 0; g < 3; ++g)
#pragma unroll
        for (int j = 0; j < 8; ++j) atomicAdd(&s_db[g * 256 + lane * 8 + j], db[g * 8 + j]);
    }
    loss_acc = warp_sum(loss_acc);
    if (lane == 0) s_loss[warp_in_block] = loss_acc;
    __syncthreads();
    if (dbias)
      for (int i = threadIdx.x; i < 768; i += blockDim.x) atomicAdd(&dbias[i], s_db[i]);
    if (threadIdx.x == 0) {
      float t = 0.f;
      for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += s_loss[i];
      atomicAdd(loss_sum, (double)t);
    }
  }
}

}  // namespace sb

using namespace sb;

extern "C" int sb_pixel_ce(const void* logits, const uint8_t* target, int B, int HW, float clip, float gscale,
                           void* dlogits, uint8_t* argmax_out, double* loss_sum, float* dbias, sb_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!logits || B < 1 || HW < 1) return SB_ERR_ARG;
  const bool train = dlogits != nullptr;
  if (train && (!target || !loss_sum)) return SB_ERR_ARG;
  if (!train && !argmax_out) return SB_ERR_ARG;
  const long long npix = (long long)B * HW;
  long long blocks = (npix + 7) / 8;
  const long long cap = 148LL * 8;  // 8 CTAs of 256 threads per SM
  if (blocks > cap) blocks = cap;
  if (train)
    pixel_ce_kernel<true><<<(int)blocks, 256, 0, stream>>>((const __nv_bfloat16*)logits, target, B, HW, clip, gscale,
                                                            (__nv_bfloat16*)dlogits, argmax_out, loss_sum, dbias);
  else
    pixel_ce_kernel<false><<<(int)blocks, 256, 0, stream>>>((const __nv_bfloat16*)logits, target, B, HW, clip, gscale,
                                                             nullptr, argmax_out, nullptr, nullptr);
  ++g_launches;
  SB_CHECK_LAUNCH();
  return SB_OK;
}
