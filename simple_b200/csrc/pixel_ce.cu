// Fused per-pixel 256-way softmax cross-entropy (+ clipping, gradient, argmax decode) -- K14 of SURVEY 2.3.
// Replaces simple/trainer.py:134-137 (CrossEntropyLoss(reduction='none') -> max(.,0.03) -> mean - 0.03) and its
// autograd backward, plus torch.argmax at simple/trainer.py:130 and simple/subproc_vec_env.py:428.
//
// Logits are bf16 [NPIX, 3*256] with the 768 channels permuted colour-major (col = colour*256 + value) so that one
// softmax group is 512 contiguous bytes.  A HALF-warp owns one group: lane q of the half holds the 16-byte chunks q and
// q+16 (two fully coalesced 256-byte loads per half-warp), i.e. 16 logits per thread, so that the three reductions
// of a group (max, sum, first arg-max) cost 4 shuffle steps each and are amortised over 16 elements per thread.
// The arithmetic uses the packed Blackwell pipes (HMNMX2.BF16 for the max, FFMA2/FADD2/FMUL2 for the exp argument,
// the sum, the gradient scaling and the bias-gradient accumulation) because this kernel is issue-bound as much as it
// is HBM-bound: every logit is read from HBM exactly once and the gradient written once (25.85 MB per frame).
// The one-hot term of the gradient touches one element of 256: it is applied by a 2-byte store of the owning lane
// after its vector store, and by one shared-memory atomic into the bias-gradient histogram.
#include "common.cuh"
#include "../../include/simple_b200.h"

namespace sb {
extern long long g_launches;

__device__ __forceinline__ uint4 ld_stream(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ unsigned long long as_u64(float2 v) { return *reinterpret_cast<unsigned long long*>(&v); }
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
  float2 d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;"
      : "=l"(*reinterpret_cast<unsigned long long*>(&d))
      : "l"(as_u64(a)), "l"(as_u64(b)), "l"(as_u64(c)));
  return d;
}
__device__ __forceinline__ float2 fadd2(float2 a, float2 b) {
  float2 d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(*reinterpret_cast<unsigned long long*>(&d)) : "l"(as_u64(a)), "l"(as_u64(b)));
  return d;
}
__device__ __forceinline__ float2 fmul2(float2 a, float2 b) {
  float2 d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(*reinterpret_cast<unsigned long long*>(&d)) : "l"(as_u64(a)), "l"(as_u64(b)));
  return d;
}
// xor-butterflies over offsets 8,4,2,1 stay inside a half-warp
__device__ __forceinline__ float half_max(float v) {
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float half_sum(float v) {
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ int half_min(int v) {
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) v = min(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// One softmax group held by a half-warp.  w[0..3] = chunk q (columns 8q..8q+7), w[4..7] = chunk q+16 (columns 128+8q..).
template <bool TRAIN>
__device__ __forceinline__ void ce_group(const uint4 ra, const uint4 rb, int q, int t, float xt, float clip, float gscale,
                                         uint8_t* amax_dst, __nv_bfloat16* dl_group, float* s_db_group, float2* db,
                                         float& loss_acc) {
  constexpr float LOG2E = 1.4426950408889634f, LN2 = 0.6931471805599453f;
  const uint32_t w[8] = {ra.x, ra.y, ra.z, ra.w, rb.x, rb.y, rb.z, rb.w};
  // ---- max in packed bf16 (exact: the inputs are bf16)
  __nv_bfloat162 m2 = *reinterpret_cast<const __nv_bfloat162*>(&w[0]);
#pragma unroll
  for (int k = 1; k < 8; ++k) m2 = __hmax2(m2, *reinterpret_cast<const __nv_bfloat162*>(&w[k]));
  const float2 mf = __bfloat1622float2(m2);
  const float gm = half_max(fmaxf(mf.x, mf.y));
  if (amax_dst) {
    // first index of the maximum (torch.argmax): equality masks per bf16 pair -> bit i = element i of this lane
    const __nv_bfloat162 g2 = __float2bfloat162_rn(gm);
    uint32_t acc = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const uint32_t eq = __heq2_mask(*reinterpret_cast<const __nv_bfloat162*>(&w[k]), g2);
      acc |= eq & ((1u << (2 * k)) | (1u << (17 + 2 * k)));
    }
    const uint32_t bits = (acc & 0xffffu) | (acc >> 16);
    const int i = __ffs(bits) - 1;  // -1 when this lane does not hold the maximum
    const int col = (i < 0) ? 256 : ((q + ((i >> 3) << 4)) << 3) + (i & 7);
    const int first = half_min(col);
    if (q == 0) *amax_dst = (uint8_t)first;
  }
  if (TRAIN) {
    const float moff = -gm * LOG2E;
    const float2 l2 = make_float2(LOG2E, LOG2E), mo2 = make_float2(moff, moff);
    float2 e[8];
    float2 s2 = make_float2(0.f, 0.f);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float2 x = make_float2(__uint_as_float(w[k] << 16), __uint_as_float(w[k] & 0xffff0000u));
      const float2 a = ffma2(x, l2, mo2);
      e[k] = make_float2(exp2f(a.x), exp2f(a.y));
      s2 = fadd2(s2, e[k]);
    }
    const float s = half_sum(s2.x + s2.y);
    const float ce = gm + log2f(s) * LN2 - xt;
    loss_acc += fmaxf(ce, clip);  // identical in the 16 lanes; lane 0's copy is used
    const float scale = (ce > clip) ? gscale : 0.f;
    const float inv = scale / s;
    const float2 inv2 = make_float2(inv, inv);
    uint32_t o[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float2 d = fmul2(e[k], inv2);
      db[k] = fadd2(db[k], d);
      const __nv_bfloat162 h = __floats2bfloat162_rn(d.x, d.y);
      o[k] = *reinterpret_cast<const uint32_t*>(&h);
    }
    *reinterpret_cast<uint4*>(dl_group + q * 8) = make_uint4(o[0], o[1], o[2], o[3]);
    *reinterpret_cast<uint4*>(dl_group + 128 + q * 8) = make_uint4(o[4], o[5], o[6], o[7]);
    // one-hot term: the lane that owns column t overwrites that element (same thread => ordered after its vector store)
    if (((t >> 3) & 15) == q) {
      const float et = exp2f(fmaf(xt, LOG2E, moff));
      dl_group[t] = __float2bfloat16_rn(et * inv - scale);
      if (s_db_group) atomicAdd(s_db_group + t, -scale);
    }
  }
}

template <bool TRAIN>
__global__ void __launch_bounds__(256, 3)
    pixel_ce_kernel(const __nv_bfloat16* __restrict__ logits, const uint8_t* __restrict__ target, int B, int HW,
                    float clip, float gscale, __nv_bfloat16* dlogits, uint8_t* __restrict__ argmax_out,
                    double* __restrict__ loss_sum, float* __restrict__ dbias) {
  const int lane = threadIdx.x & 31;
  const int q = lane & 15;
  const long long npix = (long long)B * HW;
  // half-warps: 16 per CTA; the total count is a multiple of 3 (host guarantees it) and a half-warp keeps ONE colour,
  // so its bias-gradient accumulators are 16 registers holding fixed columns
  const long long hw_global = (long long)blockIdx.x * (blockDim.x >> 4) + (threadIdx.x >> 4);
  const long long nhw = (long long)gridDim.x * (blockDim.x >> 4);
  const int g = (int)(hw_global % 3);
  const long long slot = hw_global / 3, nslots = nhw / 3;
  __shared__ float s_db[768];
  __shared__ float s_loss[16];
  if (TRAIN) {
    for (int i = threadIdx.x; i < 768; i += blockDim.x) s_db[i] = 0.f;
    __syncthreads();
  }
  float2 db[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) db[i] = make_float2(0.f, 0.f);
  float loss_acc = 0.f;
  const unsigned short* lraw = reinterpret_cast<const unsigned short*>(logits);
  for (long long pix0 = slot; pix0 < npix; pix0 += 2 * nslots) {
    const long long pix1 = pix0 + nslots;
    const bool has1 = pix1 < npix;
    const __nv_bfloat16* g0 = logits + pix0 * 768 + g * 256;
    const __nv_bfloat16* g1 = logits + pix1 * 768 + g * 256;
    // four independent 16-byte loads in flight per thread
    const uint4 a0 = ld_stream(g0 + q * 8);
    const uint4 b0 = ld_stream(g0 + 128 + q * 8);
    uint4 a1 = make_uint4(0, 0, 0, 0), b1 = a1;
    if (has1) {
      a1 = ld_stream(g1 + q * 8);
      b1 = ld_stream(g1 + 128 + q * 8);
    }
    const int bi0 = (int)(pix0 / HW), hw0 = (int)(pix0 - (long long)bi0 * HW);
    const long long o0 = ((long long)bi0 * 3 + g) * HW + hw0;
    long long o1 = 0;
    int t0 = 0, t1 = 0;
    float xt0 = 0.f, xt1 = 0.f;
    if (TRAIN) {
      t0 = target[o0];
      xt0 = __uint_as_float((uint32_t)__ldg(lraw + (pix0 * 768 + g * 256 + t0)) << 16);
    }
    if (has1) {
      const int bi1 = (int)(pix1 / HW), hw1 = (int)(pix1 - (long long)bi1 * HW);
      o1 = ((long long)bi1 * 3 + g) * HW + hw1;
      if (TRAIN) {
        t1 = target[o1];
        xt1 = __uint_as_float((uint32_t)__ldg(lraw + (pix1 * 768 + g * 256 + t1)) << 16);
      }
    }
    ce_group<TRAIN>(a0, b0, q, t0, xt0, clip, gscale, argmax_out ? argmax_out + o0 : nullptr,
                    TRAIN ? dlogits + pix0 * 768 + g * 256 : nullptr, dbias ? s_db + g * 256 : nullptr, db, loss_acc);
    if (has1)
      ce_group<TRAIN>(a1, b1, q, t1, xt1, clip, gscale, argmax_out ? argmax_out + o1 : nullptr,
                      TRAIN ? dlogits + pix1 * 768 + g * 256 : nullptr, dbias ? s_db + g * 256 : nullptr, db,
                      loss_acc);
  }
  if (TRAIN) {
    if (dbias) {
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int col = g * 256 + ((k >> 2) << 7) + q * 8 + (k & 3) * 2;
        atomicAdd(&s_db[col], db[k].x);
        atomicAdd(&s_db[col + 1], db[k].y);
      }
    }
    if (q == 0) s_loss[threadIdx.x >> 4] = loss_acc;
    __syncthreads();
    if (dbias)
      for (int i = threadIdx.x; i < 768; i += blockDim.x) atomicAdd(&dbias[i], s_db[i]);
    if (threadIdx.x == 0) {
      float tt = 0.f;
      for (int i = 0; i < (int)(blockDim.x >> 4); ++i) tt += s_loss[i];
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
  // 16 half-warps per CTA, one colour per half-warp: the CTA count is a multiple of 3 so that the half-warp count is
  long long blocks = (npix * 3 + 15) / 16;
  const long long cap = 148LL * 3;  // 444 = 3 * 148 CTAs of 256 threads: one resident wave at 3 CTAs per SM
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
