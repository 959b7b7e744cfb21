// Scheduled-sampling frame feedback of the world-model trainer (simple/trainer.py:58-65,125-132), one 4-CTA cluster per
// sample:
//   next  = ground-truth frame w.p. epsilon, else the model's own arg-max decode            (u8 [3,HW])
//   next  = next / 255, each value replaced w.p. input_noise by the frame's (lower) median  (torch.median semantics)
//   stack = cat(stack[3:], next)                                                            (fp32 [12,HW], in place)
// The reference does this per sample in Python (median through a sort, CPU multinomial, torch.cat); the previous
// revision here used ~17 ATen launches including a 118 us gatherMedian.  The pixel values are bytes, so the median is
// an exact 256-bin histogram walk in shared memory.
#include "common.cuh"
#include "../../include/simple_b200.h"

namespace sb {
extern long long g_launches;

constexpr int kFfCluster = 4;  // CTAs per sample: each histograms and rewrites a quarter of the frame

__device__ __forceinline__ uint32_t ld_cluster_u32(uint32_t cluster_addr) {
  uint32_t v;
  asm volatile("ld.shared::cluster.u32 %0, [%1];" : "=r"(v) : "r"(cluster_addr) : "memory");
  return v;
}

// grid = (kFfCluster, B), one 4-CTA cluster per sample.  Each CTA histograms a quarter of the sample's bytes, the four
// partial histograms are summed through distributed shared memory, the median is found with a block-wide scan, and
// each CTA rewrites its quarter of the pixels.
__global__ void __cluster_dims__(kFfCluster, 1, 1) __launch_bounds__(256)
    frame_feedback_kernel(float* __restrict__ stack, const uint8_t* __restrict__ gt, const uint8_t* __restrict__ pred,
                          const float* __restrict__ eps_ptr, float p_noise, const unsigned long long* seed_ptr,
                          unsigned long long seed, int C_stack, int C_new, int HW) {
  __shared__ unsigned int hist[256];
  __shared__ unsigned int warp_tot[8];
  __shared__ int s_median;
  const int b = blockIdx.y, t = threadIdx.x;
  const uint32_t rank = cluster_ctarank();
  const unsigned long long sd = seed_ptr ? seed + *seed_ptr : seed;
  const uint2 key = hash_key(sd, 0x46454544u);
  // scheduled sampling: one uniform per sample
  const float eps = eps_ptr ? *eps_ptr : 0.f;
  const bool use_gt = u01(hash4x32(0x7fffff00u + (uint32_t)b, key).x) < eps;
  const uint8_t* src = (use_gt ? gt : pred) + (size_t)b * C_new * HW;
  const int n = C_new * HW;
  hist[t] = 0u;
  if (t == 0) s_median = 0;
  __syncthreads();
  if (p_noise > 0.f) {  // (uniform over the grid: every CTA takes the same branch, cluster barriers included)
    // Histogram of this CTA's share of the frame's bytes.  16 bytes per thread and load, warp-aggregated increments:
    // Atari frames are mostly one background colour, so every lane hits the same bin and plain shared-memory atomics
    // would serialise on one address.
    const int lane = t & 31;
    const bool aligned = (reinterpret_cast<uintptr_t>(src) & 15) == 0;
    const int nvec = aligned ? n / 16 : 0;
    const int vper = (nvec + kFfCluster - 1) / kFfCluster;
    const int v_begin = (int)rank * vper, v_end = min(nvec, v_begin + vper);
    for (int i0 = v_begin; i0 < v_end; i0 += blockDim.x) {
      const int i = i0 + t;
      uint4 q = make_uint4(0, 0, 0, 0);
      if (i < v_end) q = __ldg(reinterpret_cast<const uint4*>(src) + i);
      const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
      for (int k = 0; k < 16; ++k) {
        const int v = i < v_end ? (int)((w[k >> 2] >> (8 * (k & 3))) & 0xffu) : 256 + lane;  // idle lanes: unique values
        const unsigned peers = __match_any_sync(0xffffffffu, v);
        if (v < 256 && lane == __ffs(peers) - 1) atomicAdd(&hist[v], (unsigned)__popc(peers));
      }
    }
    if (rank == 0) {  // tail / unaligned fallback
      for (int i0 = nvec * 16; i0 < n; i0 += blockDim.x) {
        const int i = i0 + t;
        const int v = i < n ? (int)src[i] : 256 + lane;
        const unsigned peers = __match_any_sync(0xffffffffu, v);
        if (v < 256 && lane == __ffs(peers) - 1) atomicAdd(&hist[v], (unsigned)__popc(peers));
      }
    }
    cluster_sync_all();  // all four partial histograms complete and visible
    unsigned int cnt = 0;
#pragma unroll
    for (uint32_t r = 0; r < kFfCluster; ++r) cnt += ld_cluster_u32(mapa_u32(&hist[t], r));
    // lower median = value at sorted position (n - 1) / 2: first bin whose inclusive prefix count exceeds it
    unsigned int incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned int up = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += up;
    }
    if (lane == 31) warp_tot[t >> 5] = incl;
    __syncthreads();
    unsigned int base = 0;
    for (int wi = 0; wi < (t >> 5); ++wi) base += warp_tot[wi];
    incl += base;
    const unsigned int target = (unsigned int)((n - 1) / 2);
    if (incl > target && incl - cnt <= target) s_median = t;  // exactly one bin satisfies this
    __syncthreads();
    cluster_sync_all();  // nobody leaves (or reuses hist) while a peer may still read its partial histogram
  }
  const float med = (float)s_median / 255.0f;
  float* st = stack + (size_t)b * C_stack * HW;
  const uint32_t thr = (uint32_t)(p_noise * 65536.0f);
  const int per = (HW + kFfCluster - 1) / kFfCluster;
  const int hw_begin = (int)rank * per, hw_end = min(HW, hw_begin + per);
  for (int hw = hw_begin + t; hw < hw_end; hw += blockDim.x) {
    // shift: channel c <- channel c + C_new, ascending c (each element is read before this thread overwrites it)
    for (int c = 0; c < C_stack - C_new; ++c) st[(size_t)c * HW + hw] = st[(size_t)(c + C_new) * HW + hw];
    uint4 r = make_uint4(0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu);
    if (p_noise > 0.f) r = hash4x32((uint32_t)(b * HW + hw), key);
    const uint32_t bits[4] = {r.x, r.y, r.z, r.w};
    for (int c = 0; c < C_new; ++c) {
      const uint32_t rb = (bits[(c >> 1) & 3] >> (16 * (c & 1))) & 0xffffu;  // 16 random bits per value (C_new <= 8)
      const bool keep = p_noise <= 0.f || rb >= thr;
      const float v = (float)src[(size_t)c * HW + hw] / 255.0f;
      st[(size_t)(C_stack - C_new + c) * HW + hw] = keep ? v : med;
    }
  }
}


// ---------------------------------------------------------------------------------------------------------------
// Simulated-env step tail (simple/subproc_vec_env.py:428-441): u8 frame stack shifted by one frame with the decoded
// frame appended (in place), the float observation the policy reads, reward = argmax(reward logits) - 1 and the value
// estimate -- one launch instead of cat / copy / argmax / sub / cast kernels.  One thread owns 16 pixels of one agent
// across all channels (reads channel k + c before it overwrites channel k: in place is safe).
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
    env_post_kernel(uint8_t* __restrict__ frames, const uint8_t* __restrict__ fresh, float* __restrict__ obs, int SC, int c,
                    int HW, const float* __restrict__ reward_pred, int R, const float* __restrict__ value_pred,
                    float* __restrict__ rew_out, float* __restrict__ val_out, int A) {
  const int a = blockIdx.y;
  const int v = blockIdx.x * blockDim.x + threadIdx.x;  // 16-pixel group
  if (v * 16 < HW) {
    for (int k = 0; k < SC; ++k) {
      const uint8_t* src = (k + c < SC) ? frames + ((size_t)a * SC + k + c) * HW : fresh + ((size_t)a * c + (k + c - SC)) * HW;
      const uint4 q = *reinterpret_cast<const uint4*>(src + (size_t)v * 16);
      *reinterpret_cast<uint4*>(frames + ((size_t)a * SC + k) * HW + (size_t)v * 16) = q;
      if (obs) {
        float4* o = reinterpret_cast<float4*>(obs + ((size_t)a * SC + k) * HW + (size_t)v * 16);
        const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
        for (int j = 0; j < 4; ++j)
          o[j] = make_float4((float)(w[j] & 255u), (float)((w[j] >> 8) & 255u), (float)((w[j] >> 16) & 255u),
                             (float)(w[j] >> 24));
      }
    }
  }
  if (blockIdx.x == 0 && blockIdx.y == 0) {
    for (int i = threadIdx.x; i < A; i += blockDim.x) {
      if (reward_pred) {
        int best = 0;
        float bv = reward_pred[(size_t)i * R];
        for (int r = 1; r < R; ++r) {
          const float x = reward_pred[(size_t)i * R + r];
          if (x > bv) { bv = x; best = r; }  // first maximum, like torch.argmax
        }
        rew_out[i] = (float)(best - 1);
      }
      if (value_pred) val_out[i] = value_pred[i];
    }
  }
}

}  // namespace sb

using namespace sb;

extern "C" int sb_frame_feedback(float* stack, const uint8_t* gt, const uint8_t* pred, const float* eps_ptr,
                                 float p_noise, const unsigned long long* seed_ptr, unsigned long long seed, int B,
                                 int C_stack, int C_new, int HW, sb_stream_t stream) {
  if (!stack || !gt || !pred || B < 1 || C_new < 1 || C_new > 8 || C_stack <= C_new || HW < 1) return SB_ERR_ARG;
  frame_feedback_kernel<<<dim3(kFfCluster, B), 256, 0, (cudaStream_t)stream>>>(stack, gt, pred, eps_ptr, p_noise, seed_ptr,
                                                                               seed, C_stack, C_new, HW);
  ++g_launches;
  SB_CHECK_LAUNCH();
  return SB_OK;
}

extern "C" int sb_env_post(uint8_t* frames, const uint8_t* fresh, float* obs, int A, int stack_channels, int c, int HW,
                           const float* reward_pred, int R, const float* value_pred, float* rew_out, float* val_out,
                           sb_stream_t stream) {
  if (!frames || !fresh || A < 1 || c < 1 || stack_channels <= c || HW < 16 || (HW & 15)) return SB_ERR_ARG;
  if ((reward_pred && (R < 1 || !rew_out)) || (value_pred && !val_out)) return SB_ERR_ARG;
  if ((reinterpret_cast<uintptr_t>(frames) | reinterpret_cast<uintptr_t>(fresh) | reinterpret_cast<uintptr_t>(obs)) & 15)
    return SB_ERR_ARG;
  const int groups = HW / 16;
  env_post_kernel<<<dim3((groups + 255) / 256, A), 256, 0, (cudaStream_t)stream>>>(
      frames, fresh, obs, stack_channels, c, HW, reward_pred, R, value_pred, rew_out, val_out, A);
  ++g_launches;
  SB_CHECK_LAUNCH();
  return SB_OK;
}
