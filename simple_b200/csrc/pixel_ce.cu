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
                                         float& loss_acc, bool valid = true) {
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
    if (q == 0 && valid) *amax_dst = (uint8_t)first;
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
    if (valid) loss_acc += fmaxf(ce, clip);  // identical in the 16 lanes; lane 0's copy is used
    const float scale = (valid && ce > clip) ? gscale : 0.f;
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
    if (valid) {
      *reinterpret_cast<uint4*>(dl_group + q * 8) = make_uint4(o[0], o[1], o[2], o[3]);
      *reinterpret_cast<uint4*>(dl_group + 128 + q * 8) = make_uint4(o[4], o[5], o[6], o[7]);
    }
    // one-hot term: the lane that owns column t overwrites that element (same thread => ordered after its vector store)
    if (valid && ((t >> 3) & 15) == q) {
      const float et = exp2f(fmaf(xt, LOG2E, moff));
      dl_group[t] = __float2bfloat16_rn(et * inv - scale);
      if (s_db_group) atomicAdd(s_db_group + t, -scale);
    }
  }
}

// Register-staged kernel.  A warp keeps ONE colour (fixed bias-gradient columns per lane) and walks the pixels
// wslot, wslot + nws, ...; each iteration gives two pixels to each half-warp (four independent 16-byte loads per thread
// in flight), with the target bytes of the NEXT iteration prefetched so that no dependent load sits on the critical path.
// All loop bounds are warp-uniform (the reductions are full-mask shuffles).
template <bool TRAIN, int MINB>
__global__ void __launch_bounds__(256, MINB)
    pixel_ce_kernel(const __nv_bfloat16* __restrict__ logits, const uint8_t* __restrict__ target, int B, int HW,
                    float clip, float gscale, __nv_bfloat16* dlogits, uint8_t* __restrict__ argmax_out,
                    double* __restrict__ loss_sum, float* __restrict__ dbias) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int hh = lane >> 4, q = lane & 15;
  const int npix = B * HW;  // (the launcher rejects B*HW*3 >= 2^31)
  const int w_global = blockIdx.x * 8 + warp;
  const int nwarps = gridDim.x * 8;  // multiple of 3 (host guarantees it)
  const int g = w_global % 3;
  const int wslot = w_global / 3, nws = nwarps / 3;
  const int n_mine = wslot < npix ? (npix - wslot + nws - 1) / nws : 0;  // pixels of this warp
  const int J = (n_mine + 3) / 4;                                        // 4 pixels per iteration
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
  // pixel k of the warp's list -> (valid, flat pixel, target/argmax offset)
  auto locate = [&](int k, bool& valid, int& pix, int& o) {
    valid = k < n_mine;
    pix = wslot + (valid ? k : 0) * nws;
    const int bi = pix / HW, hw = pix - bi * HW;
    o = (bi * 3 + g) * HW + hw;
  };
  bool v0, v1;
  int p0, p1, o0, o1;
  int t0 = 0, t1 = 0;
  locate(hh, v0, p0, o0);
  locate(2 + hh, v1, p1, o1);
  if (TRAIN) {
    if (v0) t0 = target[o0];
    if (v1) t1 = target[o1];
  }
  for (int j = 0; j < J; ++j) {
    const __nv_bfloat16* g0 = logits + (size_t)p0 * 768 + g * 256;
    const __nv_bfloat16* g1 = logits + (size_t)p1 * 768 + g * 256;
    uint4 a0 = make_uint4(0, 0, 0, 0), b0 = a0, a1 = a0, b1 = a0;
    if (v0) {
      a0 = ld_stream(g0 + q * 8);
      b0 = ld_stream(g0 + 128 + q * 8);
    }
    if (v1) {
      a1 = ld_stream(g1 + q * 8);
      b1 = ld_stream(g1 + 128 + q * 8);
    }
    // target logits: the target bytes were prefetched one iteration ago, so these two 2-byte loads are issued together
    // with the group loads (one more sector of the same lines) instead of a 15-instruction register select + shuffle
    float xt0 = 0.f, xt1 = 0.f;
    if (TRAIN) {
      const unsigned short* r0 = reinterpret_cast<const unsigned short*>(g0);
      const unsigned short* r1 = reinterpret_cast<const unsigned short*>(g1);
      if (v0) xt0 = __uint_as_float((uint32_t)__ldg(r0 + t0) << 16);
      if (v1) xt1 = __uint_as_float((uint32_t)__ldg(r1 + t1) << 16);
    }
    // next iteration's pixels and target bytes (addresses do not depend on data)
    bool nv0, nv1;
    int np0, np1, no0, no1;
    int nt0 = 0, nt1 = 0;
    locate(4 * (j + 1) + hh, nv0, np0, no0);
    locate(4 * (j + 1) + 2 + hh, nv1, np1, no1);
    if (TRAIN) {
      if (nv0) nt0 = target[no0];
      if (nv1) nt1 = target[no1];
    }
    ce_group<TRAIN>(a0, b0, q, t0, xt0, clip, gscale, argmax_out ? argmax_out + o0 : nullptr,
                    TRAIN ? dlogits + (size_t)p0 * 768 + g * 256 : nullptr, (TRAIN && dbias) ? s_db + g * 256 : nullptr, db,
                    loss_acc, v0);
    ce_group<TRAIN>(a1, b1, q, t1, xt1, clip, gscale, argmax_out ? argmax_out + o1 : nullptr,
                    TRAIN ? dlogits + (size_t)p1 * 768 + g * 256 : nullptr, (TRAIN && dbias) ? s_db + g * 256 : nullptr, db,
                    loss_acc, v1);
    v0 = nv0; v1 = nv1; p0 = np0; p1 = np1; o0 = no0; o1 = no1; t0 = nt0; t1 = nt1;
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

// ---------------------------------------------------------------------------------------------------------------
// TMA-staged variant: every warp owns a ring of kCeStages shared-memory stages (2 groups = 1 KB each) filled by 1-D bulk
// copies (cp.async.bulk, completion on an mbarrier) that its lane 0 keeps kCeStages - 1 iterations ahead.  The bytes in
// flight per SM (~96 KB at 4 CTAs) no longer depend on registers or occupancy -- the register-staged kernel above sits
// at ~50 KB and 52 % issue utilisation with long-scoreboard stalls -- and the target logit comes out of shared memory.
// ---------------------------------------------------------------------------------------------------------------
constexpr int kCeStages = 4;
__device__ __forceinline__ void bulk_load_1d(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

template <bool TRAIN>
__global__ void __launch_bounds__(256, 4)
    pixel_ce_tma_kernel(const __nv_bfloat16* __restrict__ logits, const uint8_t* __restrict__ target, int B, int HW,
                        float clip, float gscale, __nv_bfloat16* dlogits, uint8_t* __restrict__ argmax_out,
                        double* __restrict__ loss_sum, float* __restrict__ dbias) {
  __shared__ __align__(128) uint8_t ring[8][kCeStages][1024];
  __shared__ uint64_t full[8][kCeStages];
  __shared__ float s_db[768];
  __shared__ float s_loss[16];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int hh = lane >> 4, q = lane & 15;
  const long long npix = (long long)B * HW;
  // a warp keeps ONE colour (fixed bias-gradient columns per lane) and walks the pixels wslot, wslot + nws, ...;
  // iteration j gives pixel 2j to its lower half-warp and pixel 2j+1 to the upper one
  const long long w_global = (long long)blockIdx.x * 8 + warp;
  const long long nwarps = (long long)gridDim.x * 8;  // multiple of 3 (host guarantees it)
  const int g = (int)(w_global % 3);
  const long long wslot = w_global / 3, nws = nwarps / 3;
  const long long n_mine = wslot < npix ? (npix - wslot + nws - 1) / nws : 0;
  const long long J = (n_mine + 1) / 2;
  if (lane == 0) {
    for (int s = 0; s < kCeStages; ++s) mbar_init(&full[warp][s], 1);
    fence_mbar_init();
  }
  if (TRAIN)
    for (int i = threadIdx.x; i < 768; i += blockDim.x) s_db[i] = 0.f;
  __syncthreads();
  auto issue = [&](long long j) {  // lane 0 only
    const int s = (int)(j % kCeStages);
    const long long k0 = 2 * j, k1 = 2 * j + 1;
    const bool v1 = k1 < n_mine;
    mbar_expect_tx(&full[warp][s], v1 ? 1024u : 512u);
    bulk_load_1d(&ring[warp][s][0], logits + (wslot + k0 * nws) * 768 + g * 256, 512u, &full[warp][s]);
    if (v1) bulk_load_1d(&ring[warp][s][512], logits + (wslot + k1 * nws) * 768 + g * 256, 512u, &full[warp][s]);
  };
  if (lane == 0)
    for (long long j = 0; j < kCeStages && j < J; ++j) issue(j);
  float2 db[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) db[i] = make_float2(0.f, 0.f);
  float loss_acc = 0.f;
  for (long long j = 0; j < J; ++j) {
    const int s = (int)(j % kCeStages);
    const uint32_t ph = (uint32_t)(j / kCeStages) & 1u;
    const long long k = 2 * j + hh;
    const bool valid = k < n_mine;
    const long long pix = wslot + (valid ? k : 0) * nws;
    const int bi = (int)(pix / HW), hw = (int)(pix - (long long)bi * HW);
    const long long o = ((long long)bi * 3 + g) * HW + hw;
    int t = 0;
    if (TRAIN && valid) t = target[o];  // issued before the wait: overlaps the stage's arrival
    mbar_wait(&full[warp][s], ph);
    const uint8_t* grp = &ring[warp][s][hh * 512];
    const uint4 ra = *reinterpret_cast<const uint4*>(grp + q * 16);
    const uint4 rb = *reinterpret_cast<const uint4*>(grp + 256 + q * 16);
    float xt = 0.f;
    if (TRAIN) xt = __uint_as_float((uint32_t)reinterpret_cast<const unsigned short*>(grp)[t] << 16);
    ce_group<TRAIN>(ra, rb, q, t, xt, clip, gscale, argmax_out ? argmax_out + o : nullptr,
                    TRAIN ? dlogits + pix * 768 + g * 256 : nullptr, (TRAIN && dbias) ? s_db + g * 256 : nullptr, db,
                    loss_acc, valid);
    __syncwarp();  // every lane has read the stage (and its values are in registers) before lane 0 refills it
    if (lane == 0 && j + kCeStages < J) issue(j + kCeStages);
  }
  if (TRAIN) {
    if (dbias) {
#pragma unroll
      for (int kk = 0; kk < 8; ++kk) {
        const int col = g * 256 + ((kk >> 2) << 7) + q * 8 + (kk & 3) * 2;
        atomicAdd(&s_db[col], db[kk].x);
        atomicAdd(&s_db[col + 1], db[kk].y);
      }
    }
    if (q == 0) s_loss[threadIdx.x >> 4] = loss_acc;
    __syncthreads();
    if (dbias)
      for (int i = threadIdx.x; i < 768; i += blockDim.x) atomicAdd(&dbias[i], s_db[i]);
    if (threadIdx.x == 0) {
      float tt = 0.f;
      for (int i = 0; i < 16; ++i) tt += s_loss[i];
      atomicAdd(loss_sum, (double)tt);
    }
  }
}

}  // namespace sb

using namespace sb;

static int g_ce_tma = 0;  // 0: register-staged kernel (default, faster as measured); 1: TMA-staged kernel
extern "C" int sb_set_ce_tma(int enabled) {
  const int prev = g_ce_tma;
  g_ce_tma = enabled ? 1 : 0;
  return prev;
}

extern "C" int sb_pixel_ce(const void* logits, const uint8_t* target, int B, int HW, float clip, float gscale,
                           void* dlogits, uint8_t* argmax_out, double* loss_sum, float* dbias, sb_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!logits || B < 1 || HW < 1 || (long long)B * HW * 3 >= (1LL << 31)) return SB_ERR_ARG;
  const bool train = dlogits != nullptr;
  if (train && (!target || !loss_sum)) return SB_ERR_ARG;
  if (!train && !argmax_out) return SB_ERR_ARG;
  const long long npix = (long long)B * HW;
  // 16 half-warps per CTA, one colour per half-warp: the CTA count is a multiple of 3 so that the half-warp count is
  long long blocks = (npix * 3 + 15) / 16;
  const long long cap = 148LL * 3;  // 444 = 3 * 148 CTAs of 256 threads: one resident wave at 3 CTAs per SM
  if (blocks > cap) blocks = cap;
  blocks = ((blocks + 2) / 3) * 3;
  if (g_ce_tma) {
    // 8 warps per CTA, one colour per warp: CTA count multiple of 3; 4 CTAs per SM resident
    long long tb = (npix * 3 + 15) / 16;
    const long long tcap = 148LL * 4 - (148LL * 4) % 3;
    if (tb > tcap) tb = tcap;
    tb = ((tb + 2) / 3) * 3;
    if (train)
      pixel_ce_tma_kernel<true><<<(int)tb, 256, 0, stream>>>((const __nv_bfloat16*)logits, target, B, HW, clip, gscale,
                                                             (__nv_bfloat16*)dlogits, argmax_out, loss_sum, dbias);
    else
      pixel_ce_tma_kernel<false><<<(int)tb, 256, 0, stream>>>((const __nv_bfloat16*)logits, target, B, HW, clip, gscale,
                                                              nullptr, argmax_out, nullptr, nullptr);
  } else if (train)
    pixel_ce_kernel<true, 3><<<(int)blocks, 256, 0, stream>>>((const __nv_bfloat16*)logits, target, B, HW, clip, gscale,
                                                             (__nv_bfloat16*)dlogits, argmax_out, loss_sum, dbias);
  else
    pixel_ce_kernel<false, 3><<<(int)blocks, 256, 0, stream>>>((const __nv_bfloat16*)logits, target, B, HW, clip, gscale,
                                                              nullptr, argmax_out, nullptr, nullptr);
  ++g_launches;
  SB_CHECK_LAUNCH();
  return SB_OK;
}
