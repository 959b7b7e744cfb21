// Latent-bit predictor kernels -- K11/K12 of SURVEY 2.3 (simple/next_frame_predictor.py:66-121).
//
//  * sb_lstm_sample2: the whole 16-step autoregressive sampling pass of BitsPredictor.forward (:109-121) in ONE launch:
//    LSTMCell(128) -> Linear 128->256 -> multinomial(exp(logits)) -> one-hot -> Linear 256->128, then ints -> bits
//    (LSB first, simple/utils.py:115-119) -> {-1,+1}; fp32, state in shared memory; the reference issues ~20 ATen
//    kernels per step (320 per forward).  multinomial(w, 1) == argmax(w / q), q ~ Exp(1) (ATen's implementation for one
//    sample): q is either supplied (parity mode) or drawn in-kernel with Philox.
//  * sb_lstm_teacher_fwd / _bwd: all 16 teacher-forced LSTMCell steps (:97-100) and their backward through time, one
//    launch each.
#include "common.cuh"
#include "../../include/simple_b200.h"

namespace sb {
extern long long g_launches;

constexpr int kH = 128;       // latent_state_size
constexpr int kV = 256;       // 2^bits_at_once symbols
constexpr int kSteps = 16;    // bottleneck_bits / bits_at_once


// One warp computes one output row: lane l holds elements 4l..4l+3 of the 128-vector, reads the matching float4 of the
// weight row (a fully coalesced 512-byte row per warp instruction) and the partial sums are reduced with shuffles.
__device__ __forceinline__ float row_dot(const float* __restrict__ w_row, const float4 v, int lane) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(w_row) + lane);
  return a.x * v.x + a.y * v.y + a.z * v.z + a.w * v.w;
}

// ---------------------------------------------------------------------------------------------------------------
// Second-generation operands: TRANSPOSED weights (thread-per-output-row matrix-vector products, coalesced, no shuffle
// reductions), input gates by table lookup, S samples per CTA pair sharing every weight.
//   whhT [128,512] = W_hh^T, w5T [128,256] = W5^T,
//   E [256,512]    = input-gate pre-activations of every symbol: W_ih (W4[:,s] + b4) + b_ih + b_hh  (dense4 of a one-hot
//                    vector is a column lookup, so the input half of the LSTM gates is a table lookup as well),
//   G0 [B,512]     = W_ih x0 + b_ih + b_hh for the first step.
// The small GEMMs that build E / G0 and the transposes are plain library calls on the host side.
// ---------------------------------------------------------------------------------------------------------------
// The sampler: a thread-block CLUSTER of two CTAs per S samples keeps the recurrent weights
// RESIDENT in shared memory for all 16 steps (fp32: W_hh^T is 256 KB, W5^T 128 KB -- one SM's 227 KB cannot hold them,
// a pair can).  CTA r owns hidden units [64r, 64r+64): the 4 x 64 gate rows of those units (128 KB of W_hh^T) and the
// logits [128r, 128r+128) (64 KB of W5^T).  Per step the halves of h and the two local arg-max candidates cross the
// pair through distributed shared memory (two cluster barriers); no weight byte is re-read from L2 after the prologue,
// which was the whole cost of the previous kernels (a dependent chain of L2 round trips per step).
__device__ __forceinline__ void st_cluster_f32(uint32_t addr, float v) {
  asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}
__device__ __forceinline__ void st_cluster_u32(uint32_t addr, uint32_t v) {
  asm volatile("st.shared::cluster.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}

template <int S>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(512)
    lstm_sample3_kernel(const float* __restrict__ G0, const float* __restrict__ h0, const float* __restrict__ c0,
                        const float* __restrict__ E, const float* __restrict__ whhT, const float* __restrict__ w5T,
                        const float* __restrict__ b5, const float* __restrict__ noise,
                        const unsigned long long* seed_ptr, int B, int hc_ld, float* __restrict__ bits, int* __restrict__ ints) {
  extern __shared__ float wsm[];
  float* wh = wsm;                    // [128 k][256 local gate rows]: local row g*64+jl <-> global row g*128 + 64r + jl
  float* w5 = wsm + kH * 256;         // [128 k][128 local logits]:    local ln <-> global logit 128r + ln
  __shared__ float hbuf[2][S][kH], c[S][kH / 2], gates[S][256], part_g[S][256], part_l[3][S][kH], score[S][kH];
  __shared__ float slot_v[2][S];
  __shared__ int slot_i[2][S];
  __shared__ int sym[S][kSteps];
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const uint32_t rank = cluster_ctarank(), peer = rank ^ 1u;
  const int b0 = (blockIdx.x >> 1) * S;
  for (int i = t; i < kH * 256; i += 512) {
    const int k = i >> 8, lr = i & 255;
    wh[i] = __ldg(whhT + k * (4 * kH) + (lr >> 6) * kH + 64 * (int)rank + (lr & 63));
  }
  for (int i = t; i < kH * 128; i += 512) {
    const int k = i >> 7, ln = i & 127;
    w5[i] = __ldg(w5T + k * kV + 128 * (int)rank + ln);
  }
  for (int i = t; i < S * kH; i += 512) {
    const int s = i / kH, j = i - s * kH;
    hbuf[0][s][j] = h0[(size_t)min(b0 + s, B - 1) * hc_ld + j];
  }
  for (int i = t; i < S * 64; i += 512) {
    const int s = i >> 6, jl = i & 63;
    c[s][jl] = c0[(size_t)min(b0 + s, B - 1) * hc_ld + 64 * (int)rank + jl];
  }
  const unsigned long long seed = seed_ptr ? *seed_ptr : 0ull;
  cluster_sync_all();  // both CTAs are running (remote shared-memory writes below) and the local weights are staged
  for (int step = 0; step < kSteps; ++step) {
    // h is double-buffered: this step reads h_old, the cells write h_new in BOTH CTAs -- the peer may still be reading
    // h_old in its gate phase when this CTA's cells are done
    const float (*h_old)[kH] = hbuf[step & 1];
    float (*h_new)[kH] = hbuf[(step & 1) ^ 1];
    {  // gate pre-activations of the local rows: thread = (k half, local row)
      const int lr = t & 255, kh = t >> 8;
      const int grow = (lr >> 6) * kH + 64 * (int)rank + (lr & 63);
      float acc[S];
#pragma unroll
      for (int s = 0; s < S; ++s) {
        acc[s] = 0.f;
        if (kh == 0) {
          const int b = min(b0 + s, B - 1);
          acc[s] = step == 0 ? G0[(size_t)b * 4 * kH + grow] : E[(size_t)sym[s][step - 1] * 4 * kH + grow];
        }
      }
#pragma unroll 8
      for (int k = kh * 64; k < kh * 64 + 64; ++k) {
        const float w = wh[k * 256 + lr];
#pragma unroll
        for (int s = 0; s < S; ++s) acc[s] = fmaf(w, h_old[s][k], acc[s]);
      }
      if (kh == 1) {
#pragma unroll
        for (int s = 0; s < S; ++s) part_g[s][lr] = acc[s];
      }
      __syncthreads();
      if (kh == 0) {
#pragma unroll
        for (int s = 0; s < S; ++s) gates[s][lr] = acc[s] + part_g[s][lr];
      }
    }
    __syncthreads();
    if (t < S * 64) {  // LSTM cell of the local hidden units; the new h goes to both CTAs
      const int s = t >> 6, jl = t & 63;
      const float ig = sigmoidf_(gates[s][jl]), fg = sigmoidf_(gates[s][64 + jl]);
      const float gg = tanh_acc(gates[s][128 + jl]), og = sigmoidf_(gates[s][192 + jl]);
      const float cn = fg * c[s][jl] + ig * gg;
      const float hn = og * tanh_acc(cn);
      c[s][jl] = cn;
      float* dst = &h_new[s][64 * (int)rank + jl];
      *dst = hn;
      st_cluster_f32(mapa_u32(dst, peer), hn);
    }
    cluster_sync_all();
    {  // local logits: thread = (k quarter, local logit)
      const int ln = t & 127, kq = t >> 7;
      float acc[S];
#pragma unroll
      for (int s = 0; s < S; ++s) acc[s] = 0.f;
#pragma unroll 8
      for (int k = kq * 32; k < kq * 32 + 32; ++k) {
        const float w = w5[k * 128 + ln];
#pragma unroll
        for (int s = 0; s < S; ++s) acc[s] = fmaf(w, h_new[s][k], acc[s]);
      }
      if (kq > 0) {
#pragma unroll
        for (int s = 0; s < S; ++s) part_l[kq - 1][s][ln] = acc[s];
      }
      __syncthreads();
      if (kq == 0) {
        const int n = 128 * (int)rank + ln;
#pragma unroll
        for (int s = 0; s < S; ++s) {
          const int b = min(b0 + s, B - 1);
          const float logit = acc[s] + part_l[0][s][ln] + part_l[1][s][ln] + part_l[2][s][ln] + b5[n];
          float q;
          if (noise) {
            q = noise[((size_t)b * kSteps + step) * kV + n];
          } else {
            const uint4 rr = philox4x32(make_uint4((uint32_t)n, (uint32_t)step, (uint32_t)b, 0x4c53544du),
                                        make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
            q = -logf(((float)(rr.x >> 8) + 0.5f) * (1.0f / 16777216.0f));
          }
          score[s][ln] = expf(logit) / q;  // sample_with_temperature(T=1): multinomial(exp(logits))
        }
      }
    }
    __syncthreads();
    if (warp < S) {  // arg-max of the 128 local scores of sample `warp` (first index on ties)
      float v = -1.0f;
      int i = 0;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float sv = score[warp][lane + 32 * j];
        if (sv > v) {
          v = sv;
          i = lane + 32 * j;
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, v, o);
        const int oi = __shfl_xor_sync(0xffffffffu, i, o);
        if (ov > v || (ov == v && oi < i)) {
          v = ov;
          i = oi;
        }
      }
      if (lane == 0) {
        i += 128 * (int)rank;
        slot_v[rank][warp] = v;
        slot_i[rank][warp] = i;
        st_cluster_f32(mapa_u32(&slot_v[rank][warp], peer), v);
        st_cluster_u32(mapa_u32(&slot_i[rank][warp], peer), (uint32_t)i);
      }
    }
    cluster_sync_all();
    if (t < S)  // CTA 0 holds the lower symbol indices: it wins ties, like a single arg-max over all 256 scores
      sym[t][step] = (slot_v[1][t] > slot_v[0][t]) ? slot_i[1][t] : slot_i[0][t];
    __syncthreads();
  }
  if (rank == 0) {
    for (int i = t; i < S * kSteps; i += 512) {
      const int s = i / kSteps, k = i - s * kSteps;
      if (b0 + s < B) ints[(b0 + s) * kSteps + k] = sym[s][k];
    }
    for (int i = t; i < S * kSteps * 8; i += 512) {
      const int s = i / (kSteps * 8), k = i - s * (kSteps * 8);
      if (b0 + s < B) bits[(size_t)(b0 + s) * (kSteps * 8) + k] = ((sym[s][k >> 3] >> (k & 7)) & 1) ? 1.0f : -1.0f;
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Cluster versions of the teacher-forced pass and its backward through time (same split as lstm_sample3_kernel: CTA r
// of the pair owns hidden units [64r, 64r+64) and keeps the 256 matching rows of W_hh resident in shared memory).
// ---------------------------------------------------------------------------------------------------------------
template <int S>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(512)
    lstm_teacher_fwd2_kernel(const float* __restrict__ xw, const float* __restrict__ h0, const float* __restrict__ c0,
                             const float* __restrict__ whhT, int B, int hc_ld, float* __restrict__ hs, float* __restrict__ act,
                             float* __restrict__ cs) {
  extern __shared__ float wsm[];
  float* wh = wsm;  // [128 k][256 local gate rows]
  __shared__ float hbuf[2][S][kH], c[S][kH / 2], gates[S][256], part_g[S][256];
  const int t = threadIdx.x;
  const uint32_t rank = cluster_ctarank(), peer = rank ^ 1u;
  const int b0 = (blockIdx.x >> 1) * S;
  for (int i = t; i < kH * 256; i += 512) {
    const int k = i >> 8, lr = i & 255;
    wh[i] = __ldg(whhT + k * (4 * kH) + (lr >> 6) * kH + 64 * (int)rank + (lr & 63));
  }
  for (int i = t; i < S * kH; i += 512) hbuf[0][i / kH][i % kH] = h0[(size_t)min(b0 + i / kH, B - 1) * hc_ld + i % kH];
  for (int i = t; i < S * 64; i += 512) c[i >> 6][i & 63] = c0[(size_t)min(b0 + (i >> 6), B - 1) * hc_ld + 64 * (int)rank + (i & 63)];
  cluster_sync_all();
  for (int step = 0; step < kSteps; ++step) {
    const float (*h_old)[kH] = hbuf[step & 1];
    float (*h_new)[kH] = hbuf[(step & 1) ^ 1];
    {
      const int lr = t & 255, kh = t >> 8;
      const int grow = (lr >> 6) * kH + 64 * (int)rank + (lr & 63);
      float acc[S];
#pragma unroll
      for (int s = 0; s < S; ++s)
        acc[s] = kh == 0 ? xw[((size_t)min(b0 + s, B - 1) * kSteps + step) * 4 * kH + grow] : 0.f;
#pragma unroll 8
      for (int k = kh * 64; k < kh * 64 + 64; ++k) {
        const float w = wh[k * 256 + lr];
#pragma unroll
        for (int s = 0; s < S; ++s) acc[s] = fmaf(w, h_old[s][k], acc[s]);
      }
      if (kh == 1) {
#pragma unroll
        for (int s = 0; s < S; ++s) part_g[s][lr] = acc[s];
      }
      __syncthreads();
      if (kh == 0) {
#pragma unroll
        for (int s = 0; s < S; ++s) gates[s][lr] = acc[s] + part_g[s][lr];
      }
    }
    __syncthreads();
    if (t < S * 64) {
      const int s = t >> 6, jl = t & 63, j = 64 * (int)rank + jl;
      const float ig = sigmoidf_(gates[s][jl]), fg = sigmoidf_(gates[s][64 + jl]);
      const float gg = tanh_acc(gates[s][128 + jl]), og = sigmoidf_(gates[s][192 + jl]);
      const float cn = fg * c[s][jl] + ig * gg;
      const float hn = og * tanh_acc(cn);
      c[s][jl] = cn;
      float* dst = &h_new[s][j];
      *dst = hn;
      st_cluster_f32(mapa_u32(dst, peer), hn);
      if (b0 + s < B) {
        const size_t row = (size_t)(b0 + s) * kSteps + step;
        float* a = act + row * 4 * kH;
        a[j] = ig; a[kH + j] = fg; a[2 * kH + j] = gg; a[3 * kH + j] = og;
        cs[row * kH + j] = cn;
        hs[row * kH + j] = hn;
      }
    }
    cluster_sync_all();
  }
}

template <int S>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(512)
    lstm_teacher_bwd2_kernel(const float* __restrict__ dhs, const float* __restrict__ act, const float* __restrict__ cs,
                             const float* __restrict__ c0, const float* __restrict__ whh, int B, int hc_ld,
                             float* __restrict__ dpre_all, float* __restrict__ dh0, float* __restrict__ dc0) {
  extern __shared__ float wsm[];
  float* wr = wsm;  // [256 local rows][128 j]: the W_hh rows of this CTA's hidden units (gate-major local order)
  __shared__ float dh[S][kH / 2], dc[S][kH / 2], dpre[S][256], part[S][4][kH], recv[2][S][kH / 2];
  const int t = threadIdx.x;
  const uint32_t rank = cluster_ctarank(), peer = rank ^ 1u;
  const int b0 = (blockIdx.x >> 1) * S;
  for (int i = t; i < 256 * kH; i += 512) {
    const int lr = i >> 7, j = i & 127;
    wr[i] = __ldg(whh + (size_t)((lr >> 6) * kH + 64 * (int)rank + (lr & 63)) * kH + j);
  }
  for (int i = t; i < S * 64; i += 512) {
    dh[i >> 6][i & 63] = 0.f;
    dc[i >> 6][i & 63] = 0.f;
  }
  cluster_sync_all();
  for (int step = kSteps - 1; step >= 0; --step) {
    if (t < S * 64) {  // cell backward of the local hidden units
      const int s = t >> 6, jl = t & 63, j = 64 * (int)rank + jl;
      const int b = min(b0 + s, B - 1);
      const size_t row = (size_t)b * kSteps + step;
      const float* a = act + row * 4 * kH;
      const float ig = a[j], fg = a[kH + j], gg = a[2 * kH + j], og = a[3 * kH + j];
      const float cprev = step > 0 ? cs[(row - 1) * kH + j] : c0[(size_t)b * hc_ld + j];
      const float tc = tanh_acc(cs[row * kH + j]);
      const float dhi = dhs[row * kH + j] + dh[s][jl];
      const float dct = dc[s][jl] + dhi * og * (1.0f - tc * tc);
      const float d0 = dct * gg * ig * (1.0f - ig), d1 = dct * cprev * fg * (1.0f - fg);
      const float d2 = dct * ig * (1.0f - gg * gg), d3 = dhi * tc * og * (1.0f - og);
      dpre[s][jl] = d0; dpre[s][64 + jl] = d1; dpre[s][128 + jl] = d2; dpre[s][192 + jl] = d3;
      dc[s][jl] = dct * fg;
      if (b0 + s < B) {
        float* d = dpre_all + row * 4 * kH;
        d[j] = d0; d[kH + j] = d1; d[2 * kH + j] = d2; d[3 * kH + j] = d3;
      }
    }
    __syncthreads();
    {  // partial dh_prev[j] over the 256 local rows: thread = (quarter of the rows, j)
      const int j = t & (kH - 1), qr = t >> 7;
      float acc[S];
#pragma unroll
      for (int s = 0; s < S; ++s) acc[s] = 0.f;
#pragma unroll 8
      for (int lr = qr * 64; lr < qr * 64 + 64; ++lr) {
        const float w = wr[lr * kH + j];
#pragma unroll
        for (int s = 0; s < S; ++s) acc[s] = fmaf(w, dpre[s][lr], acc[s]);
      }
#pragma unroll
      for (int s = 0; s < S; ++s) part[s][qr][j] = acc[s];
    }
    __syncthreads();
    if (t < S * kH) {  // own half stays, the peer's half goes over distributed shared memory
      const int s = t / kH, j = t - s * kH;
      const float v = part[s][0][j] + part[s][1][j] + part[s][2][j] + part[s][3][j];
      if ((j >> 6) == (int)rank) dh[s][j & 63] = v;
      else st_cluster_f32(mapa_u32(&recv[step & 1][s][j & 63], peer), v);
    }
    cluster_sync_all();
    if (t < S * 64) dh[t >> 6][t & 63] += recv[step & 1][t >> 6][t & 63];
    __syncthreads();
  }
  for (int i = t; i < S * 64; i += 512) {
    const int s = i >> 6, jl = i & 63;
    if (b0 + s < B) {
      dh0[(size_t)(b0 + s) * hc_ld + 64 * (int)rank + jl] = dh[s][jl];
      dc0[(size_t)(b0 + s) * hc_ld + 64 * (int)rank + jl] = dc[s][jl];
    }
  }
}

}  // namespace sb

using namespace sb;

template <int S>
static int launch_sample3(const float* G0, const float* h0, const float* c0, const float* E, const float* whhT,
                          const float* w5T, const float* b5, const float* noise, const unsigned long long* seed_ptr, int B,
                          int hc_ld, float* bits, int* ints, cudaStream_t st) {
  constexpr int smem = (kH * 256 + kH * 128) * sizeof(float);  // resident weight slices: 192 KB
  static bool configured = false;
  if (!configured) {
    if (cudaFuncSetAttribute(lstm_sample3_kernel<S>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess)
      return SB_ERR_CUDA;
    configured = true;
  }
  const int groups = (B + S - 1) / S;
  lstm_sample3_kernel<S><<<2 * groups, 512, smem, st>>>(G0, h0, c0, E, whhT, w5T, b5, noise, seed_ptr, B, hc_ld, bits,
                                                        ints);
  ++g_launches;
  SB_CHECK_LAUNCH();
  return SB_OK;
}

extern "C" int sb_lstm_sample2(const float* G0, const float* h0, const float* c0, const float* E, const float* whhT,
                               const float* w5T, const float* b5, const float* noise,
                               const unsigned long long* seed_ptr, int B, int hc_ld, float* bits, int* ints,
                               sb_stream_t stream) {
  if (!G0 || !h0 || !c0 || !E || !whhT || !w5T || !b5 || !bits || !ints || B < 1 || hc_ld < kH) return SB_ERR_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  // one CTA pair (cluster) per one or two samples, recurrent weights resident in the pair's shared memory
  if (B >= 64 && B % 2 == 0) return launch_sample3<2>(G0, h0, c0, E, whhT, w5T, b5, noise, seed_ptr, B, hc_ld, bits, ints, st);
  return launch_sample3<1>(G0, h0, c0, E, whhT, w5T, b5, noise, seed_ptr, B, hc_ld, bits, ints, st);
}

template <int S>
static int launch_teacher(bool bwd, const float* p0, const float* p1, const float* p2, const float* p3, const float* p4,
                          int B, int hc_ld, float* o0, float* o1, float* o2, cudaStream_t st) {
  constexpr int smem = kH * 256 * sizeof(float);  // resident W_hh slice: 128 KB
  static bool configured = false;
  if (!configured) {
    if (cudaFuncSetAttribute(lstm_teacher_fwd2_kernel<S>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) !=
            cudaSuccess ||
        cudaFuncSetAttribute(lstm_teacher_bwd2_kernel<S>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) !=
            cudaSuccess)
      return SB_ERR_CUDA;
    configured = true;
  }
  const int groups = (B + S - 1) / S;
  if (!bwd) lstm_teacher_fwd2_kernel<S><<<2 * groups, 512, smem, st>>>(p0, p1, p2, p3, B, hc_ld, o0, o1, o2);
  else lstm_teacher_bwd2_kernel<S><<<2 * groups, 512, smem, st>>>(p0, p1, p2, p3, p4, B, hc_ld, o0, o1, o2);
  ++g_launches;
  SB_CHECK_LAUNCH();
  return SB_OK;
}

extern "C" int sb_lstm_teacher_fwd(const float* xw, const float* h0, const float* c0, const float* whhT, int B, int hc_ld,
                                   float* hs, float* act, float* cs, sb_stream_t stream) {
  if (!xw || !h0 || !c0 || !whhT || !hs || !act || !cs || B < 1 || hc_ld < kH) return SB_ERR_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  if (B >= 64 && B % 2 == 0) return launch_teacher<2>(false, xw, h0, c0, whhT, nullptr, B, hc_ld, hs, act, cs, st);
  return launch_teacher<1>(false, xw, h0, c0, whhT, nullptr, B, hc_ld, hs, act, cs, st);
}

extern "C" int sb_lstm_teacher_bwd(const float* dhs, const float* act, const float* cs, const float* c0,
                                   const float* whh, int B, int hc_ld, float* dpre, float* dh0, float* dc0,
                                   sb_stream_t stream) {
  if (!dhs || !act || !cs || !c0 || !whh || !dpre || !dh0 || !dc0 || B < 1 || hc_ld < kH) return SB_ERR_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  if (B >= 64 && B % 2 == 0) return launch_teacher<2>(true, dhs, act, cs, c0, whh, B, hc_ld, dpre, dh0, dc0, st);
  return launch_teacher<1>(true, dhs, act, cs, c0, whh, B, hc_ld, dpre, dh0, dc0, st);
}
