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

// One softmax group (256 bf16 = 512 contiguous bytes) per warp and iteration; a warp stays on ONE colour so that its
// bias-gradient accumulators are 8 registers; two pixels are in flight per iteration (memory-level parallelism).
template <bool TRAIN>
__device__ __forceinline__ void ce_group(const uint4 raw, int lane, int t, float clip, float gscale, uint8_t* amax_dst,
                                         __nv_bfloat16* dl_dst, float* db, float& loss_acc) {
  constexpr float LOG2E = 1.4426950408889634f, LN2 = 0.6931471805599453f;
  float x[8];
  {
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 f = __bfloat1622float2(h[j]);
      x[2 * j] = f.x;
      x[2 * j + 1] = f.y;
    }
  }
  float m = fmaxf(fmaxf(fmaxf(x[0], x[1]), fmaxf(x[2], x[3])), fmaxf(fmaxf(x[4], x[5]), fmaxf(x[6], x[7])));
  const float wm = warp_max(m);
  if (amax_dst) {
    // first index of the maximum (torch.argmax): lowest lane holding it, then its lowest slot
    const unsigned who = __ballot_sync(0xffffffffu, m == wm);
    const int src = __ffs(who) - 1;
    int slot = 7;
#pragma unroll
    for (int j = 6; j >= 0; --j) slot = (x[j] == wm) ? j : slot;
    slot = __shfl_sync(0xffffffffu, slot, src);
    if (lane == 0) *amax_dst = (uint8_t)(src * 8 + slot);
  }
  if (TRAIN) {
    float e[8];
    float s = 0.f;
    const float moff = wm * LOG2E;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      e[j] = exp2f(fmaf(x[j], LOG2E, -moff));
      s += e[j];
    }
    s = warp_sum(s);
    float xt = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) xt = ((t & 7) == j) ? x[j] : xt;
    xt = __shfl_sync(0xffffffffu, xt, t >> 3);
    const float ce = wm + log2f(s) * LN2 - xt;
    loss_acc += fmaxf(ce, clip);  // identical in every lane; lane 0's copy is used
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
    *reinterpret_cast<BF8*>(dl_dst) = o;
#pragma unroll
    for (int j = 0; j < 8; ++j) db[j] += d[j];
  }
}

template <bool TRAIN>
__global__ void __launch_bounds__(256, 4)
    pixel_ce_kernel(const __nv_bfloat16* __restrict__ logits, const uint8_t* __restrict__ target, int B, int HW,
                    float clip, float gscale, __nv_bfloat16* dlogits, uint8_t* __restrict__ argmax_out,
                    double* __restrict__ loss_sum, float* __restrict__ dbias) {
  const int lane = threadIdx.x & 31;
  const int warp_in_block = threadIdx.x >> 5;
  const long long npix = (long long)B * HW;
  const long long warp_global = (long long)blockIdx.x * (blockDim.x >> 5) + warp_in_block;
  const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);  // multiple of 3 (host guarantees it)
  const int g = (int)(warp_global % 3);                                // this warp's colour
  const long long wslot = warp_global / 3, nslots = nwarps / 3;
  __shared__ float s_db[768];
  __shared__ float s_loss[8];
  if (TRAIN && dbias) {
    for (int i = threadIdx.x; i < 768; i += blockDim.x) s_db[i] = 0.f;
    __syncthreads();
  }
  float db[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) db[i] = 0.f;
  float loss_acc = 0.f;
  for (long long pix0 = wslot; pix0 < npix; pix0 += 2 * nslots) {
    const long long pix1 = pix0 + nslots;
    const bool has1 = pix1 < npix;
    const uint4 r0 = ld_stream(logits + pix0 * 768 + g * 256 + lane * 8);
    uint4 r1 = make_uint4(0, 0, 0, 0);
    if (has1) r1 = ld_stream(logits + pix1 * 768 + g * 256 + lane * 8);
    int b0 = (int)(pix0 / HW), hw0 = (int)(pix0 - (long long)b0 * HW);
    const long long o0 = ((long long)b0 * 3 + g) * HW + hw0;
    int t0 = 0, t1 = 0;
    long long o1 = 0;
    if (TRAIN) t0 = target[o0];
    if (has1) {
      int b1 = (int)(pix1 / HW), hw1 = (int)(pix1 - (long long)b1 * HW);
      o1 = ((long long)b1 * 3 + g) * HW + hw1;
      if (TRAIN) t1 = target[o1];
    }
    ce_group<TRAIN>(r0, lane, t0, clip, gscale, argmax_out ? argmax_out + o0 : nullptr,
                    TRAIN ? dlogits + pix0 * 768 + g * 256 + lane * 8 : nullptr, db, loss_acc);
    if (has1)
      ce_group<TRAIN>(r1, lane, t1, clip, gscale, argmax_out ? argmax_out + o1 : nullptr,
                      TRAIN ? dlogits + pix1 * 768 + g * 256 + lane * 8 : nullptr, db, loss_acc);
  }
  if (TRAIN) {
    if (dbias) {
#pragma unroll
      for (int j = 0; j < 8; ++j) atomicAdd(&s_db[g * 256 + lane * 8 + j], db[j]);
    }
    if (lane == 0) s_loss[warp_in_block] = loss_acc;
    __syncthreads();
    if (dbias)
      for (int i = threadIdx.x; i < 768; i += blockDim.x) atomicAdd(&dbias[i], s_db[i]);
    if (threadIdx.x == 0) {
      float tt = 0.f;
      for (int i = 0; i < (int)(blockDim.x >> 5); ++i) tt += s_loss[i];
      atomicAdd(loss_sum, (double)tt);
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
  // 8 warps per CTA; the total warp count must be a multiple of 3 (one colour per warp): CTA count multiple of 3
  long long blocks = (npix * 3 + 7) / 8;
  const long long cap = 148LL * 6;  // 888 = 3 * 296 CTAs of 256 threads (4 resident per SM)
  if (blocks > cap) blocks = cap;
  blocks = ((blocks + 2) / 3) * 3;
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
