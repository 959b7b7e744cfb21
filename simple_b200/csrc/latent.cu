// Fused kernels of the world model's LATENT / HEADS block -- everything between the bottleneck activation and the middle
// network, plus the reward / value heads and the scalar losses (north_star item 3: "discrete-latent bit inference and
// straight-through sampling ... as fused elementwise and reduction kernels").
//
// Reference arithmetic replaced (file:line under /root/reference):
//   simple/next_frame_predictor.py:328          ValueEstimator (4608-wide dot product, C,H,W flatten order)
//   simple/utils.py:98-105                      ActionInjector scale / shift (sigmoid(dense1 a), dense2 a), all 7 at once
//   simple/next_frame_predictor.py:176-197      posterior bits: sign, truncated-normal noise, tanh, straight-through, flip
//                                               noise, the three `mix` calls (simple/utils.py:157-163)
//   simple/next_frame_predictor.py:87-107       teacher forcing: bits -> bytes (LSB first), dense4(one_hot) as a column
//                                               lookup, dropout 0.1, (LSTM: lstm.cu), dropout 0.1, CE over 256 symbols
//   simple/next_frame_predictor.py:153-159      add_bits
//   simple/next_frame_predictor.py:11-24,339,352-354 + simple/trainer.py:138-141   reward MLP tail, reward CE, value MSE
// The dense layers between these stages are [B,K]x[K,N] sb_dense launches (dense.cu) issued by the host wrapper; every
// elementwise / reduction / gather / scatter piece around them, forward AND backward, lives here, so that one training
// step launches ~45 kernels for this block instead of ~300 ATen kernels.
// All tensors fp32 unless stated.  "flat" = the reference's flatten order of [768,3,2]: index c*6 + p, p = h*2 + w.
#include "common.cuh"
#include "../../include/simple_b200.h"

namespace sb {
extern long long g_launches;

constexpr int kLC = 768;    // bottleneck channels
constexpr int kLP = 6;      // bottleneck pixels (3 x 2)
constexpr int kLF = kLC * kLP;
constexpr int kBits = 128;  // bottleneck_bits

__device__ __forceinline__ float block_sum256(float v, float* red) {
  v = warp_sum(v);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) red[w] = v;
  __syncthreads();
  float t = 0.f;
  const int nw = (blockDim.x + 31) >> 5;
  for (int i = 0; i < nw; ++i) t += red[i];
  return t;
}

// counter-based uniform in [0,1): independent of launch geometry, regenerated identically by the backward kernels
__device__ __forceinline__ float rnd_u01(uint2 key, uint32_t idx) { return u01(pcg_word(idx ^ key.x) ^ key.y); }
__device__ __forceinline__ uint2 rnd_key(const unsigned long long* seed_ptr, unsigned long long seed, uint32_t stream) {
  return hash_key(seed_ptr ? seed + *seed_ptr : seed, stream);
}
__device__ __forceinline__ float read_eps(const float* eps_ptr, float eps_val) { return eps_ptr ? *eps_ptr : eps_val; }

// ---------------------------------------------------------------------------------------------------------------
// 1. heads-in: value head + ActionInjector bank + injection of the bottleneck
//    bank [B, 2S]: row b = (dense1 pre-activations of all injectors | dense2 outputs), S = sum of injector widths.
//    sc_flat / sh_flat: per-injector contiguous blocks [B, C_i] at element offset offs[i] * B.
// ---------------------------------------------------------------------------------------------------------------
struct InjOffsets {
  int n;
  int offs[9];  // prefix sums of the injector widths; offs[n] = S
};
__device__ __forceinline__ int inj_of(const InjOffsets& o, int j) {
  int i = 0;
  while (i + 1 < o.n && j >= o.offs[i + 1]) ++i;
  return i;
}

__global__ void __launch_bounds__(256)
    heads_fwd_kernel(const __nv_bfloat16* __restrict__ x5, const float* __restrict__ bank, InjOffsets o,
                     const float* __restrict__ wv, const float* __restrict__ bv, int B, float* __restrict__ sc_flat,
                     float* __restrict__ sh_flat, float* __restrict__ h_flat, float* __restrict__ value) {
  __shared__ float red[8];
  const int b = blockIdx.x, S = o.offs[o.n];
  const float* brow = bank + (size_t)b * 2 * S;
  for (int j = threadIdx.x; j < S; j += blockDim.x) {
    const int i = inj_of(o, j), c = j - o.offs[i], C = o.offs[i + 1] - o.offs[i];
    const size_t dst = (size_t)o.offs[i] * B + (size_t)b * C + c;
    sc_flat[dst] = sigmoidf_(brow[j]);
    sh_flat[dst] = brow[S + j];
  }
  float vacc = 0.f;
  for (int e = threadIdx.x; e < kLF; e += blockDim.x) {  // e = p*768 + c (NHWC): coalesced bf16 reads
    const int p = e / kLC, c = e - p * kLC;
    const float x = __bfloat162float(x5[(size_t)b * kLF + e]);
    const int f = c * kLP + p;
    vacc += x * __ldg(wv + f);
    if (h_flat) h_flat[(size_t)b * kLF + f] = x * sigmoidf_(brow[c]) + brow[S + c];  // injector 0 sits at offset 0
  }
  vacc = block_sum256(vacc, red);
  if (threadIdx.x == 0) value[b] = vacc + bv[0];
}

struct InjGrads {          // gradients w.r.t. the post-sigmoid scale / the shift of injectors 1..n-1 (strided views)
  const float* dsc[9];
  const float* dsh[9];
  int sc_sb[9], sc_sc[9], sh_sb[9], sh_sc[9];
};

__global__ void __launch_bounds__(256)
    heads_bwd_kernel(const __nv_bfloat16* __restrict__ x5, const float* __restrict__ bank, InjOffsets o, InjGrads g,
                     const float* __restrict__ wv, const float* __restrict__ dvalue, const float* __restrict__ dh_flat,
                     int B, __nv_bfloat16* __restrict__ dx5, float* __restrict__ dbank, float* dwv, float* dbv,
                     float* dbank_b) {
  const int b = blockIdx.x, S = o.offs[o.n];
  const float* brow = bank + (size_t)b * 2 * S;
  float* drow = dbank + (size_t)b * 2 * S;
  const float dv = dvalue ? dvalue[b] : 0.f;
  for (int c = threadIdx.x; c < kLC; c += blockDim.x) {
    const float s0 = sigmoidf_(brow[c]);
    float a_sc = 0.f, a_sh = 0.f;
#pragma unroll
    for (int p = 0; p < kLP; ++p) {
      const int f = c * kLP + p;
      const float dh = dh_flat ? dh_flat[(size_t)b * kLF + f] : 0.f;
      const float x = __bfloat162float(x5[(size_t)b * kLF + p * kLC + c]);
      a_sc += dh * x;
      a_sh += dh;
      const float w = __ldg(wv + f);
      dx5[(size_t)b * kLF + p * kLC + c] = __float2bfloat16_rn(dh * s0 + dv * w);
      if (dv != 0.f) atomicAdd(dwv + f, dv * x);
    }
    const float d1 = a_sc * s0 * (1.0f - s0);
    drow[c] = d1;
    drow[S + c] = a_sh;
    atomicAdd(dbank_b + c, d1);
    atomicAdd(dbank_b + S + c, a_sh);
  }
  for (int j = o.offs[1] + threadIdx.x; j < S; j += blockDim.x) {
    const int i = inj_of(o, j), c = j - o.offs[i];
    const float gsc = g.dsc[i] ? g.dsc[i][(size_t)b * g.sc_sb[i] + (size_t)c * g.sc_sc[i]] : 0.f;
    const float gsh = g.dsh[i] ? g.dsh[i][(size_t)b * g.sh_sb[i] + (size_t)c * g.sh_sc[i]] : 0.f;
    const float s = sigmoidf_(brow[j]);
    const float d1 = gsc * s * (1.0f - s);
    drow[j] = d1;
    drow[S + j] = gsh;
    if (d1 != 0.f) atomicAdd(dbank_b + j, d1);
    if (gsh != 0.f) atomicAdd(dbank_b + S + j, gsh);
  }
  if (threadIdx.x == 0 && dv != 0.f) atomicAdd(dbv, dv);
}

// ---------------------------------------------------------------------------------------------------------------
// 2. posterior bits (training branch of StochasticModel.forward)
//    flags bit 0: flip noise is -1; bit 1: first mix took x (m1); bit 2: second mix took the posterior bits (m2)
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float truncnorm02(float u) {
  // truncated normal on [-2, 2] sigma with sigma = 0.2 (scipy.stats.truncnorm.rvs(-2, 2, scale=0.2)), inverse CDF
  const float lo = 0.02275013194817921f, hi = 0.9772498680518208f;
  const float q = lo + u * (hi - lo);
  return 0.2f * 1.4142135623730951f * erfinvf(2.0f * q - 1.0f);
}

__global__ void __launch_bounds__(128)
    posterior_bits_fwd_kernel(const float* __restrict__ pre, const float* __restrict__ tn, const float* __restrict__ u_flip,
                              const float* __restrict__ u_mix1, const float* eps_ptr, float eps_val, float noise_p,
                              const unsigned long long* seed_ptr, unsigned long long seed, float* __restrict__ x_out,
                              float* __restrict__ bits2, float* __restrict__ bits_clean, uint8_t* __restrict__ flags,
                              int* __restrict__ tints) {
  __shared__ uint8_t sbit[kBits];
  const int b = blockIdx.x, j = threadIdx.x;
  const size_t i = (size_t)b * kBits + j;
  const float eps = read_eps(eps_ptr, eps_val);
  const float p = pre[i];
  const float t = tn ? tn[i] : truncnorm02(rnd_u01(rnd_key(seed_ptr, seed, 101u), (uint32_t)i));
  const float uf = u_flip ? u_flip[i] : rnd_u01(rnd_key(seed_ptr, seed, 102u), (uint32_t)i);
  const float u1 = u_mix1 ? u_mix1[i] : rnd_u01(rnd_key(seed_ptr, seed, 103u), (uint32_t)i);
  const float x = tanh_acc(p + t);
  const float hard = x > 0.f ? 1.0f : -1.0f;        // 2 * (0 < x) - 1
  const bool neg = !(noise_p < uf);                  // noise = 2 * (bottleneck_noise < u) - 1
  const float b1 = neg ? -hard : hard;
  const bool m1 = u1 < 1.0f - eps;                   // mix(bits, x, 1 - epsilon): x where u < 1 - epsilon
  x_out[i] = x;
  bits2[i] = m1 ? x : b1;
  const bool clean = p > 0.f;
  bits_clean[i] = clean ? 1.0f : -1.0f;
  flags[i] = (uint8_t)((neg ? 1 : 0) | (m1 ? 2 : 0));
  sbit[j] = clean ? 1 : 0;
  __syncthreads();
  if (j < 16) {  // bytes of the clean bits, LSB first (simple/utils.py:108-112)
    int v = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) v |= (int)sbit[j * 8 + k] << k;
    tints[b * 16 + j] = v;
  }
}

// bits3 = mix(bits_pred, bits2, p2), p2 = 1 - (1 - epsilon) * latent_rnn_max_sampling  (next_frame_predictor.py:194)
__global__ void __launch_bounds__(128)
    mix_bits_kernel(const float* __restrict__ bits_pred, const float* __restrict__ bits2, const float* __restrict__ u_mix2,
                    const float* eps_ptr, float eps_val, float max_sampling, const unsigned long long* seed_ptr,
                    unsigned long long seed, float* __restrict__ bits3, uint8_t* __restrict__ flags) {
  const size_t i = (size_t)blockIdx.x * kBits + threadIdx.x;
  const float eps = read_eps(eps_ptr, eps_val);
  const float u2 = u_mix2 ? u_mix2[i] : rnd_u01(rnd_key(seed_ptr, seed, 104u), (uint32_t)i);
  const bool m2 = u2 < 1.0f - (1.0f - eps) * max_sampling;
  bits3[i] = m2 ? bits2[i] : bits_pred[i];
  flags[i] = (uint8_t)((flags[i] & 3) | (m2 ? 4 : 0));
}

// d pre = d bits3 * m2 * ((1 - m1) * noise + m1) * (1 - x^2); db1 += column sums
__global__ void __launch_bounds__(128)
    posterior_bits_bwd_kernel(const float* __restrict__ dbits3, const float* __restrict__ x, const uint8_t* __restrict__ flags,
                              float* __restrict__ dpre, float* db1) {
  const int j = threadIdx.x;
  const size_t i = (size_t)blockIdx.x * kBits + j;
  const uint8_t f = flags[i];
  float d = (f & 4) ? dbits3[i] : 0.f;
  d *= (f & 2) ? 1.0f : ((f & 1) ? -1.0f : 1.0f);
  const float xv = x[i];
  d *= 1.0f - xv * xv;
  dpre[i] = d;
  if (d != 0.f) atomicAdd(db1 + j, d);
}

// ---------------------------------------------------------------------------------------------------------------
// 3. teacher forcing: LSTM inputs [first | dropout(dense4(one_hot(byte_t)))] and their backward
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
    teacher_embed_fwd_kernel(const float* __restrict__ proj, int proj_ld, const int* __restrict__ tints,
                             const float* __restrict__ w4, const float* __restrict__ b4, const float* __restrict__ keep,
                             float p, const unsigned long long* seed_ptr, unsigned long long seed,
                             const float* __restrict__ b_ih, const float* __restrict__ b_hh, float* __restrict__ bsum,
                             float* __restrict__ tin) {
  const int b = blockIdx.x >> 4, t = blockIdx.x & 15, j = threadIdx.x;
  float v;
  if (t == 0) {
    v = proj[(size_t)b * proj_ld + j];
  } else {
    const int s = tints[b * 16 + t - 1];
    const size_t ki = ((size_t)b * 16 + (t - 1)) * 128 + j;
    const float k = keep ? keep[ki] : (rnd_u01(rnd_key(seed_ptr, seed, 105u), (uint32_t)ki) >= p ? 1.0f : 0.0f);
    v = (__ldg(w4 + (size_t)j * 256 + s) + __ldg(b4 + j)) * k * (1.0f / (1.0f - p));
  }
  tin[((size_t)b * 16 + t) * 128 + j] = v;
  if (blockIdx.x == 0 && bsum)
    for (int q = j; q < 512; q += 128) bsum[q] = b_ih[q] + b_hh[q];
}

// dtin [B,16,128] -> dproj[:, :128] = dtin[:, 0];  dW4[:, byte_t] += dtin[:, t+1] * keep / (1-p), db4 likewise
__global__ void __launch_bounds__(128)
    teacher_embed_bwd_kernel(const float* __restrict__ dtin, const int* __restrict__ tints, const float* __restrict__ keep,
                             float p, const unsigned long long* seed_ptr, unsigned long long seed,
                             float* __restrict__ dproj, int dproj_ld, float* dw4, float* db4) {
  const int b = blockIdx.x >> 4, t = blockIdx.x & 15, j = threadIdx.x;
  const float d = dtin[((size_t)b * 16 + t) * 128 + j];
  if (t == 0) {
    dproj[(size_t)b * dproj_ld + j] = d;
    return;
  }
  const int s = tints[b * 16 + t - 1];
  const size_t ki = ((size_t)b * 16 + (t - 1)) * 128 + j;
  const float k = keep ? keep[ki] : (rnd_u01(rnd_key(seed_ptr, seed, 105u), (uint32_t)ki) >= p ? 1.0f : 0.0f);
  const float de = d * k * (1.0f / (1.0f - p));
  if (de != 0.f) {
    atomicAdd(dw4 + (size_t)j * 256 + s, de);
    atomicAdd(db4 + j, de);
  }
}

// out = x * keep / (1 - p)  (F.dropout with an explicit or regenerated mask); used forward and backward
__global__ void __launch_bounds__(256)
    drop_scale_kernel(const float* __restrict__ x, const float* __restrict__ keep, float p, const unsigned long long* seed_ptr,
                      unsigned long long seed, uint32_t stream, float* __restrict__ out, long long n) {
  const uint2 key = keep ? make_uint2(0u, 0u) : rnd_key(seed_ptr, seed, stream);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float k = keep ? keep[i] : (rnd_u01(key, (uint32_t)i) >= p ? 1.0f : 0.0f);
    out[i] = x[i] * k * (1.0f / (1.0f - p));
  }
}

// ---------------------------------------------------------------------------------------------------------------
// 4. row-wise softmax cross-entropy, loss + gradient in one pass (LSTM symbols: N = 256, R = 16 B)
//    loss_sum += sum_r (lse_r - x[r, t_r]);  dlogits = (softmax - onehot) * gscale;  dbias += column sums of dlogits
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
    softmax_ce_rows_kernel(const float* __restrict__ logits, const int* __restrict__ targets, int R, int N, float gscale,
                           float* loss_sum, float* __restrict__ dlogits, float* dbias) {
  __shared__ float sdb[256];  // per-CTA column sums of the gradient (N <= 256): one global atomic per column and CTA
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int r = blockIdx.x * (blockDim.x >> 5) + warp;
  if (dbias) {
    for (int c = threadIdx.x; c < N; c += blockDim.x) sdb[c] = 0.f;
    __syncthreads();
  }
  if (r < R) {
    const float* x = logits + (size_t)r * N;
    float m = -INFINITY;
    for (int c = lane; c < N; c += 32) m = fmaxf(m, x[c]);
    m = warp_max(m);
    float s = 0.f;
    for (int c = lane; c < N; c += 32) s += expf(x[c] - m);
    s = warp_sum(s);
    const int t = targets[r];
    if (lane == 0) atomicAdd(loss_sum, m + logf(s) - x[t]);
    const float inv = gscale / s;
    for (int c = lane; c < N; c += 32) {
      const float d = expf(x[c] - m) * inv - (c == t ? gscale : 0.f);
      dlogits[(size_t)r * N + c] = d;
      if (dbias) atomicAdd(&sdb[c], d);
    }
  }
  if (dbias) {
    __syncthreads();
    for (int c = threadIdx.x; c < N; c += blockDim.x) atomicAdd(dbias + c, sdb[c]);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// 5. latent out: add_bits + the third mix + spatial mean + bf16 NHWC operand of the middle network
//    zz [B, 1536] = (dense2(bits) | dense3(bits)) pre-activations / shifts; flags m3: the mix kept `layer`
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
    latent_out_fwd_kernel(const float* __restrict__ h_flat, const float* __restrict__ zz, const float* __restrict__ u_mix3,
                          int train, const float* eps_ptr, float eps_val, float use_max_p,
                          const unsigned long long* seed_ptr, unsigned long long seed, __nv_bfloat16* __restrict__ hm,
                          float* __restrict__ x_mid, uint8_t* __restrict__ m3) {
  const int b = blockIdx.x;
  const float p3 = 1.0f - (1.0f - read_eps(eps_ptr, eps_val)) * use_max_p;
  const uint2 key = (train && !u_mix3) ? rnd_key(seed_ptr, seed, 106u) : make_uint2(0u, 0u);
  for (int c = threadIdx.x; c < kLC; c += blockDim.x) {
    const float s = sigmoidf_(zz[(size_t)b * 2 * kLC + c]);
    const float a = zz[(size_t)b * 2 * kLC + kLC + c];
    float acc = 0.f;
#pragma unroll
    for (int p = 0; p < kLP; ++p) {
      const size_t f = (size_t)b * kLF + c * kLP + p;
      const float h = h_flat[f];
      bool keep_layer = false;
      if (train) {
        const float u = u_mix3 ? u_mix3[f] : rnd_u01(key, (uint32_t)f);
        keep_layer = u < p3;  // mix(res, layer, p3): layer where u < p3
        m3[f] = keep_layer ? 1 : 0;
      }
      const float o = keep_layer ? h : h * s + a;
      acc += o;
      hm[(size_t)b * kLF + p * kLC + c] = __float2bfloat16_rn(o);
    }
    x_mid[(size_t)b * kLC + c] = acc * (1.0f / kLP);
  }
}

__global__ void __launch_bounds__(256)
    latent_out_bwd_kernel(const __nv_bfloat16* __restrict__ dhm, const float* __restrict__ dx_mid, int dxm_ld,
                          const float* __restrict__ h_flat, const float* __restrict__ zz, const uint8_t* __restrict__ m3,
                          float* __restrict__ dh_flat, float* __restrict__ dzz, float* db23) {
  const int b = blockIdx.x;
  for (int c = threadIdx.x; c < kLC; c += blockDim.x) {
    const float s = sigmoidf_(zz[(size_t)b * 2 * kLC + c]);
    const float dm = dx_mid ? dx_mid[(size_t)b * dxm_ld + c] * (1.0f / kLP) : 0.f;
    float a_mul = 0.f, a_add = 0.f;
#pragma unroll
    for (int p = 0; p < kLP; ++p) {
      const size_t f = (size_t)b * kLF + c * kLP + p;
      const float dout = (dhm ? __bfloat162float(dhm[(size_t)b * kLF + p * kLC + c]) : 0.f) + dm;
      if (m3[f]) {
        dh_flat[f] = dout;
      } else {
        dh_flat[f] = dout * s;
        a_mul += dout * h_flat[f];
        a_add += dout;
      }
    }
    const float d1 = a_mul * s * (1.0f - s);
    dzz[(size_t)b * 2 * kLC + c] = d1;
    dzz[(size_t)b * 2 * kLC + kLC + c] = a_add;
    if (d1 != 0.f) atomicAdd(db23 + c, d1);
    if (a_add != 0.f) atomicAdd(db23 + kLC + c, a_add);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// 6. weight-only operands of the sampling kernel: W_hh^T, W_ih^T, W5^T, and emb[s] = W4[:, s] + b4 for all 256 symbols
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
    sampler_tables_kernel(const float* __restrict__ w_hh, const float* __restrict__ w_ih, const float* __restrict__ w5,
                          const float* __restrict__ w4, const float* __restrict__ b4, float* __restrict__ whhT,
                          float* __restrict__ wihT, float* __restrict__ w5T, float* __restrict__ emb) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;  // 3 x 65536 + 32768 outputs
  if (i < 65536) {                  // whhT [128,512] <- w_hh [512,128]
    const int k = i >> 9, r = i & 511;
    whhT[i] = w_hh[r * 128 + k];
  } else if (i < 131072) {          // wihT [128,512] <- w_ih [512,128]
    const int q = i - 65536, k = q >> 9, r = q & 511;
    wihT[q] = w_ih[r * 128 + k];
  } else if (i < 163840) {          // w5T [128,256] <- w5 [256,128]
    const int q = i - 131072, k = q >> 8, r = q & 255;
    w5T[q] = w5[r * 128 + k];
  } else if (i < 196608) {          // emb [256,128] <- w4 [128,256]^T + b4
    const int q = i - 163840, s = q >> 7, j = q & 127;
    emb[q] = w4[j * 256 + s] + b4[j];
  }
}

// ---------------------------------------------------------------------------------------------------------------
// 7. reward MLP tail + reward CE + value MSE, forward and backward in one launch (one CTA per sample)
//    hid_pre [B,128] = dense1(cat(x_mid, x_fin)); reward_pred = dense2(relu(hid_pre)) [B,3]
//    loss_out[0] += CE(reward_pred, rewards) / B;  loss_out[1] += (value_pred - values)^2 / B
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
    reward_value_kernel(const float* __restrict__ hid_pre, const float* __restrict__ w2, const float* __restrict__ b2,
                        const long long* __restrict__ rewards, const float* __restrict__ value_pred,
                        const float* __restrict__ values, int B, float* __restrict__ reward_pred, float* loss_out,
                        float* __restrict__ dhid_pre, float* dw2, float* db2, float* __restrict__ dvalue) {
  __shared__ float red[8];
  __shared__ float s_r[3];
  const int b = blockIdx.x, j = threadIdx.x;
  const float hp = hid_pre[(size_t)b * 128 + j];
  const float h = fmaxf(hp, 0.f);
  float w[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    w[k] = __ldg(w2 + k * 128 + j);
    const float t = block_sum256(w[k] * h, red);
    if (j == 0) s_r[k] = t + b2[k];
  }
  __syncthreads();
  const float r0 = s_r[0], r1 = s_r[1], r2 = s_r[2];
  if (j < 3) reward_pred[b * 3 + j] = s_r[j];
  if (!rewards) return;  // forward only (rollout)
  const float m = fmaxf(r0, fmaxf(r1, r2));
  const float e0 = expf(r0 - m), e1 = expf(r1 - m), e2 = expf(r2 - m);
  const float s = e0 + e1 + e2;
  const int t = (int)rewards[b];
  const float invB = 1.0f / (float)B;
  const float d[3] = {(e0 / s - (t == 0 ? 1.f : 0.f)) * invB, (e1 / s - (t == 1 ? 1.f : 0.f)) * invB,
                      (e2 / s - (t == 2 ? 1.f : 0.f)) * invB};
  dhid_pre[(size_t)b * 128 + j] = hp > 0.f ? d[0] * w[0] + d[1] * w[1] + d[2] * w[2] : 0.f;
#pragma unroll
  for (int k = 0; k < 3; ++k) atomicAdd(dw2 + k * 128 + j, d[k] * h);
  if (j < 3) atomicAdd(db2 + j, d[j]);
  if (j == 0) {
    const float rt = t == 0 ? r0 : (t == 1 ? r1 : r2);
    atomicAdd(loss_out, (m + logf(s) - rt) * invB);
    const float dv = value_pred[b] - values[b];
    atomicAdd(loss_out + 1, dv * dv * invB);
    dvalue[b] = 2.0f * dv * invB;
  }
}

// column sums of an fp32 matrix [M, N] -> out[N] (+=): bias gradients of the dense layers
__global__ void __launch_bounds__(256) colsum_f32_kernel(const float* __restrict__ x, int M, int N, int ld, float* out) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= N) return;
  const int rows_per = (M + gridDim.y - 1) / gridDim.y;
  const int r0 = blockIdx.y * rows_per, r1 = min(M, r0 + rows_per);
  float s = 0.f;
  for (int r = r0; r < r1; ++r) s += x[(size_t)r * ld + c];
  if (r1 > r0) atomicAdd(out + c, s);
}

// ---------------------------------------------------------------------------------------------------------------
// step bookkeeping: the losses of simple/trainer.py:134-144 from the kernels' raw sums, and the batch gather of
// trainer.py:108-121 (replay[idx + j] -> static step inputs).  Two tiny launches instead of ~15 ATen kernels.
// ---------------------------------------------------------------------------------------------------------------
__global__ void loss_total_kernel(const double* __restrict__ rec_sum, double inv_n, float clip,
                                  const float* __restrict__ l_value, const float* __restrict__ l_reward,
                                  const float* __restrict__ l_lstm, float* __restrict__ out5, float* running5) {
  if (threadIdx.x != 0) return;
  // max(ce, clip) summed by the logits kernel; trainer.py:136-137 subtracts the clip constant again after the mean
  const float rec = (float)(*rec_sum * inv_n - (double)clip);
  const float v = *l_value, r = *l_reward, l = l_lstm ? *l_lstm : 0.f;
  const float total = ((rec + v) + r) + l;  // the order of trainer.py:141-143
  const float o[5] = {total, rec, v, r, l};
#pragma unroll
  for (int i = 0; i < 5; ++i) {
    out5[i] = o[i];
    if (running5) running5[i] += o[i];
  }
}

__global__ void __launch_bounds__(256)
    gather_step_kernel(const long long* __restrict__ idx, long long nrec, const uint8_t* __restrict__ obs_src,
                       long long frame_bytes, const float* __restrict__ act_src, int n_action,
                       const long long* __restrict__ rew_src, const float* __restrict__ val_src,
                       uint8_t* __restrict__ obs, float* __restrict__ act, long long* __restrict__ rew,
                       float* __restrict__ val) {
  const int b = blockIdx.y;
  long long r = idx[b];
  r = r < 0 ? 0 : (r >= nrec ? nrec - 1 : r);  // (the trainer only samples valid starts; never read out of bounds)
  const uint4* src = reinterpret_cast<const uint4*>(obs_src + r * frame_bytes);
  uint4* dst = reinterpret_cast<uint4*>(obs + (long long)b * frame_bytes);
  const long long nv = frame_bytes >> 4;
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < nv; i += (long long)gridDim.x * 256)
    dst[i] = __ldg(src + i);
  if (blockIdx.x == 0) {
    if (threadIdx.x < n_action) act[(long long)b * n_action + threadIdx.x] = act_src[r * n_action + threadIdx.x];
    if (threadIdx.x == 32) rew[b] = rew_src[r];
    if (threadIdx.x == 33) val[b] = val_src[r];
  }
}

}  // namespace sb

using namespace sb;

static InjOffsets mk_offs(const int* widths, int n) {
  InjOffsets o;
  o.n = n;
  o.offs[0] = 0;
  for (int i = 0; i < 9; ++i) o.offs[i] = 0;
  for (int i = 0; i < n; ++i) o.offs[i + 1] = o.offs[i] + widths[i];
  return o;
}

extern "C" int sb_heads_fwd(const void* x5, const float* bank, const int* widths, int ninj, const float* wv,
                            const float* bv, int B, float* sc_flat, float* sh_flat, float* h_flat, float* value,
                            sb_stream_t stream) {
  if (!x5 || !bank || !widths || ninj < 1 || ninj > 8 || !wv || !bv || B < 1 || !sc_flat || !sh_flat || !value)
    return SB_ERR_ARG;
  if (widths[0] != kLC) return SB_ERR_ARG;
  heads_fwd_kernel<<<B, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)x5, bank, mk_offs(widths, ninj), wv, bv, B,
                                                        sc_flat, sh_flat, h_flat, value);
  ++g_launches;
  SB_CHECK_LAUNCH();
  return SB_OK;
}

extern "C" int sb_heads_bwd(const void* x5, const float* bank, const int* widths, int ninj, const sb_inj_grads* grads,
                            const float* wv, const float* dvalue, const float* dh_flat, int B, void* dx5, float* dbank,
                            float* dwv, float* dbv, float* dbank_b, sb_stream_t stream) {
  if (!x5 || !bank || !widths || ninj < 1 || ninj > 8 || !wv || B < 1 || !dx5 || !dbank || !dwv || !dbv || !dbank_b)
    return SB_ERR_ARG;
  InjGrads g;
  for (int i = 0; i < 9; ++i) {
    g.dsc[i] = (grads && i < ninj) ? grads->dsc[i] : nullptr;
    g.dsh[i] = (grads && i < ninj) ? grads->dsh[i] : nullptr;
    g.sc_sb[i] = (grads && i < ninj) ? grads->sc_sb[i] : 0;
    g.sc_sc[i] = (grads && i < ninj) ? grads->sc_sc[i] : 0;
    g.sh_sb[i] = (grads && i < ninj) ? grads->sh_sb[i] : 0;
    g.sh_sc[i] = (grads && i < ninj) ? grads->sh_sc[i] : 0;
  }
  heads_bwd_kernel<<<B, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)x5, bank, mk_offs(widths, ninj), g, wv, dvalue,
                                                        dh_flat, B, (__nv_bfloat16*)dx5, dbank, dwv, dbv, dbank_b);
  ++g_launches;
  SB_CHECK_LAUNCH();
  return SB_OK;
}

extern "C" int sb_posterior_bits_fwd(const float* pre, const float* tn, const float* u_flip, const float* u_mix1,
                                     const float* eps_ptr, float eps_val, float noise_p,
                                     const unsigned long long* seed_ptr, unsigned long long seed, int B, float* x,
                                     float* bits2, float* bits_clean, uint8_t* flags, int* tints, sb_stream_t stream) {
  if (!pre || B < 1 || !x || !bits2 || !bits_clean || !flags || !tints) return SB_ERR_ARG;
  posterior_bits_fwd_kernel<<<B, kBits, 0, (cudaStream_t)stream>>>(pre, tn, u_flip, u_mix1, eps_ptr, eps_val, noise_p,
                                                                   seed_ptr, seed, x, bits2, bits_clean, flags, tints);
  ++g_launches;
  SB_CHECK_LAUNCH();
  return SB_OK;
}

extern "C" int sb_mix_bits(const float* bits_pred, const float* bits2, const float* u_mix2, const float* eps_ptr,
                           float eps_val, float max_sampling, const unsigned long long* seed_ptr, unsigned long long seed,
                           int B, float* bits3, uint8_t* flags, sb_stream_t stream) {
  if (!bits_pred || !bits2 || B < 1 || !bits3 || !flags) return SB_ERR_ARG;
  mix_bits_kernel<<<B, kBits, 0, (cudaStream_t)stream>>>(bits_pred, bits2, u_mix2, eps_ptr, eps_val, max_sampling,
                                                         seed_ptr, seed, bits3, flags);
  ++g_launches;
  SB_CHECK_LAUNCH();
  return SB_OK;
}

extern "C" int sb_posterior_bits_bwd(const float* dbits3, const float* x, const uint8_t* flags, int B, float* dpre,
                                     float* db1, sb_stream_t stream) {
  if (!dbits3 || !x || !flags || B < 1 || !dpre || !db1) return SB_ERR_ARG;
  posterior_bits_bwd_kernel<<<B, kBits, 0, (cudaStream_t)stream>>>(dbits3, x, flags, dpre, db1);
  ++g_launches;
  SB_CHECK_LAUNCH();
  return SB_OK;
}

extern "C" int sb_teacher_embed_fwd(const float* proj, int proj_ld, const int* tints, const float* w4, const float* b4,
                                    const float* keep, float p, const unsigned long long* seed_ptr,
                                    unsigned long long seed, const float* b_ih, const float* b_hh, float* bsum, int B,
                                    float* tin, sb_stream_t stream) {
  if (!proj || !tints || !w4 || !b4 || B < 1 || !tin || (bsum && (!b_ih || !b_hh))) return SB_ERR_ARG;
  teacher_embed_fwd_kernel<<<B * 16, 128, 0, (cudaStream_t)stream>>>(proj, proj_ld, tints, w4, b4, keep, p, seed_ptr, seed,
                                                                     b_ih, b_hh, bsum, tin);
  ++g_launches;
  SB_CHECK_LAUNCH();
  return SB_OK;
}

extern "C" int sb_teacher_embed_bwd(const float* dtin, const int* tints, const float* keep, float p,
                                    const unsigned long long* seed_ptr, unsigned long long seed, int B, float* dproj,
                                    int dproj_ld, float* dw4, float* db4, sb_stream_t stream) {
  if (!dtin || !tints || B < 1 || !dproj || !dw4 || !db4) return SB_ERR_ARG;
  teacher_embed_bwd_kernel<<<B * 16, 128, 0, (cudaStream_t)stream>>>(dtin, tints, keep, p, seed_ptr, seed, dproj, dproj_ld,
                                                                     dw4, db4);
  ++g_launches;
  SB_CHECK_LAUNCH();
  return SB_OK;
}

extern "C" int sb_drop_scale(const float* x, const float* keep, float p, const unsigned long long* seed_ptr,
                             unsigned long long seed, unsigned int stream_id, float* out, long long n,
                             sb_stream_t stream) {
  if (!x || !out || n < 1 || n >= (1LL << 32) || p < 0.f || p >= 1.f) return SB_ERR_ARG;
  const long long blocks = (n + 255) / 256;
  drop_scale_kernel<<<(int)(blocks < 1184 ? blocks : 1184), 256, 0, (cudaStream_t)stream>>>(x, keep, p, seed_ptr, seed,
                                                                                           stream_id, out, n);
  ++g_launches;
  SB_CHECK_LAUNCH();
  return SB_OK;
}

extern "C" int sb_softmax_ce_rows(const float* logits, const int* targets, int R, int N, float gscale, float* loss_sum,
                                  float* dlogits, float* dbias, sb_stream_t stream) {
  if (!logits || !targets || R < 1 || N < 1 || N > 256 || !loss_sum || !dlogits) return SB_ERR_ARG;
  softmax_ce_rows_kernel<<<(R + 7) / 8, 256, 0, (cudaStream_t)stream>>>(logits, targets, R, N, gscale, loss_sum, dlogits,
                                                                        dbias);
  ++g_launches;
  SB_CHECK_LAUNCH();
  return SB_OK;
}

extern "C" int sb_latent_out_fwd(const float* h_flat, const float* zz, const float* u_mix3, int train,
                                 const float* eps_ptr, float eps_val, float use_max_p,
                                 const unsigned long long* seed_ptr, unsigned long long seed, int B, void* hm,
                                 float* x_mid, uint8_t* m3, sb_stream_t stream) {
  if (!h_flat || !zz || B < 1 || !hm || !x_mid || (train && !m3)) return SB_ERR_ARG;
  latent_out_fwd_kernel<<<B, 256, 0, (cudaStream_t)stream>>>(h_flat, zz, u_mix3, train, eps_ptr, eps_val, use_max_p,
                                                             seed_ptr, seed, (__nv_bfloat16*)hm, x_mid, m3);
  ++g_launches;
  SB_CHECK_LAUNCH();
  return SB_OK;
}

extern "C" int sb_latent_out_bwd(const void* dhm, const float* dx_mid, int dxm_ld, const float* h_flat, const float* zz,
                                 const uint8_t* m3, int B, float* dh_flat, float* dzz, float* db23, sb_stream_t stream) {
  if (!h_flat || !zz || !m3 || B < 1 || !dh_flat || !dzz || !db23 || (dx_mid && dxm_ld < kLC)) return SB_ERR_ARG;
  latent_out_bwd_kernel<<<B, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)dhm, dx_mid, dxm_ld, h_flat, zz, m3,
                                                             dh_flat, dzz, db23);
  ++g_launches;
  SB_CHECK_LAUNCH();
  return SB_OK;
}

extern "C" int sb_sampler_tables(const float* w_hh, const float* w_ih, const float* w5, const float* w4, const float* b4,
                                 float* whhT, float* wihT, float* w5T, float* emb, sb_stream_t stream) {
  if (!w_hh || !w_ih || !w5 || !w4 || !b4 || !whhT || !wihT || !w5T || !emb) return SB_ERR_ARG;
  sampler_tables_kernel<<<196608 / 256, 256, 0, (cudaStream_t)stream>>>(w_hh, w_ih, w5, w4, b4, whhT, wihT, w5T, emb);
  ++g_launches;
  SB_CHECK_LAUNCH();
  return SB_OK;
}

extern "C" int sb_reward_value(const float* hid_pre, const float* w2, const float* b2, const long long* rewards,
                               const float* value_pred, const float* values, int B, float* reward_pred, float* loss_out,
                               float* dhid_pre, float* dw2, float* db2, float* dvalue, sb_stream_t stream) {
  if (!hid_pre || !w2 || !b2 || B < 1 || !reward_pred) return SB_ERR_ARG;
  if (rewards && (!value_pred || !values || !loss_out || !dhid_pre || !dw2 || !db2 || !dvalue)) return SB_ERR_ARG;
  reward_value_kernel<<<B, 128, 0, (cudaStream_t)stream>>>(hid_pre, w2, b2, rewards, value_pred, values, B, reward_pred,
                                                           loss_out, dhid_pre, dw2, db2, dvalue);
  ++g_launches;
  SB_CHECK_LAUNCH();
  return SB_OK;
}

extern "C" int sb_colsum_f32(const float* x, int M, int N, int ld, float* out, sb_stream_t stream) {
  if (!x || M < 1 || N < 1 || ld < N || !out) return SB_ERR_ARG;
  const int ysplit = M >= 256 ? 8 : (M >= 32 ? 4 : 1);
  colsum_f32_kernel<<<dim3((N + 255) / 256, ysplit), 256, 0, (cudaStream_t)stream>>>(x, M, N, ld, out);
  ++g_launches;
  SB_CHECK_LAUNCH();
  return SB_OK;
}

extern "C" int sb_loss_total(const double* rec_sum, double n, float clip, const float* l_value, const float* l_reward,
                             const float* l_lstm, float* out5, float* running5, sb_stream_t stream) {
  if (!rec_sum || n <= 0 || !l_value || !l_reward || !out5) return SB_ERR_ARG;
  loss_total_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(rec_sum, 1.0 / n, clip, l_value, l_reward, l_lstm, out5, running5);
  ++g_launches;
  SB_CHECK_LAUNCH();
  return SB_OK;
}

extern "C" int sb_gather_step(const long long* idx, int B, long long nrec, const uint8_t* obs_src, long long frame_bytes,
                              const float* act_src, int n_action, const long long* rew_src, const float* val_src,
                              uint8_t* obs, float* act, long long* rew, float* val, sb_stream_t stream) {
  if (!idx || B < 1 || nrec < 1 || !obs_src || frame_bytes < 16 || (frame_bytes & 15) || !act_src || n_action < 1 ||
      n_action > 32 || !rew_src || !val_src || !obs || !act || !rew || !val)
    return SB_ERR_ARG;
  if ((reinterpret_cast<uintptr_t>(obs_src) | reinterpret_cast<uintptr_t>(obs)) & 15) return SB_ERR_ARG;
  const int chunks = (int)((frame_bytes / 16 + 255) / 256);
  gather_step_kernel<<<dim3(chunks < 8 ? chunks : 8, B), 256, 0, (cudaStream_t)stream>>>(
      idx, nrec, obs_src, frame_bytes, act_src, n_action, rew_src, val_src, obs, act, rew, val);
  ++g_launches;
  SB_CHECK_LAUNCH();
  return SB_OK;
}
