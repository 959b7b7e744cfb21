// Fused elementwise / reduction kernels around the tap GEMMs (HBM-bound; vectorised 16-byte accesses, NHWC bf16).
//
//  * sb_standardize      simple/utils.py:122-126 (+ loops next_frame_predictor.py:307,333-334)
//  * sb_stats            per-(b,c) sums for InstanceNorm2d (next_frame_predictor.py:247,269,49) and spatial means (:339,352)
//  * sb_prep             InstanceNorm apply + timing add + skip save + dropout + timing add + ActionInjector
//                        + layout change (plain <-> padded space-to-depth)  (next_frame_predictor.py:320-326,343-350)
//  * sb_prep_bwd_*       the matching backward
//  * sb_gru_blend[_bwd]  next_frame_predictor.py:298-303
//  * sb_pack_conv        fp32 (O,I,kh,kw) master weights -> bf16 K-major GEMM operands
#include <cstdlib>
#include <unordered_map>

#include "common.cuh"
#include "../../include/simple_b200.h"

namespace sb {
extern long long g_launches;
extern int g_sm_reserve;

struct __align__(16) BF8 {
  __nv_bfloat162 v[4];
};
__device__ __forceinline__ void bf8_to_f(const BF8& b, float* f) {
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    float2 t = __bfloat1622float2(b.v[j]);
    f[2 * j] = t.x;
    f[2 * j + 1] = t.y;
  }
}
__device__ __forceinline__ BF8 f_to_bf8(const float* f) {
  BF8 b;
#pragma unroll
  for (int j = 0; j < 4; ++j) b.v[j] = __floats2bfloat162_rn(f[2 * j], f[2 * j + 1]);
  return b;
}
__device__ __forceinline__ void ld8f(const float* p, float* f) {
  float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
  f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
}

// Compile-time feature masks.  The fused stages are issue-bound on integer / predicate / branch instructions when every
// optional piece is a runtime test, so each kernel is instantiated for the handful of feature combinations the network
// uses (dead pieces vanish) next to one fully-runtime instance (kRT) that serves everything else (tests, parity masks).
constexpr uint32_t kRT = 0x80000000u;
enum : uint32_t {
  F_RELU = 1u, F_NORM = 2u, F_TA = 4u, F_OUT1 = 8u, F_DROP = 16u, F_TB = 32u, F_INJ = 64u, F_ADD = 128u,
  F_INS2D = 256u, F_OUTS2D = 512u, F_G1 = 1024u, F_DADD = 2048u, F_DCOL = 4096u
};
template <uint32_t F, uint32_t BIT>
__device__ __forceinline__ bool has(bool runtime) {
  return F == kRT ? runtime : (F & BIT) != 0;
}

// A bf16 tensor holding a logical [B,H,W,C] activation, either plain NHWC or padded space-to-depth:
//   s2d element (b, y2, x2, (py*s+px)*C + c) = logical (b, s*y2+py-pad, s*x2+px-pad, c);  s is a power of two
struct View {
  void* p;
  int kind;  // 0 plain, 1 s2d
  int s, pad, H2, W2;
  int sh;    // log2(s)
};
// element offsets fit 32 bits (the launchers reject tensors of 2^31 elements or more)
template <uint32_t F, uint32_t BIT>
__device__ __forceinline__ uint32_t view_off(const View& v, int H, int W, int C, int b, int y, int x, int c) {
  if (!has<F, BIT>(v.kind != 0)) return (uint32_t)(((b * H + y) * W + x) * C + c);
  const int yp = y + v.pad, xp = x + v.pad;
  const int y2 = yp >> v.sh, x2 = xp >> v.sh;
  const int ph = ((yp & (v.s - 1)) << v.sh) + (xp & (v.s - 1));
  return (uint32_t)(((((b * v.H2 + y2) * v.W2 + x2) << (2 * v.sh)) + ph) * C + c);
}
__device__ __forceinline__ uint32_t plain_off(int H, int W, int C, int b, int y, int x, int c) {
  return (uint32_t)(((b * H + y) * W + x) * C + c);
}
// n / d for 0 <= n < 2^17, 2 <= d < 300 with m = ceil(2^32 / d) (exact in that range); m == 0 encodes d == 1
__device__ __forceinline__ int fast_div(int n, uint32_t m) { return m ? (int)__umulhi((uint32_t)n, m) : n; }

struct PrepParams {
  int B, H, W, C;
  int b_base;               // first sample of this launch (grid.y covers b_count samples): batch-chunked launches keep a
                            // stage's reduce -> apply working set inside the 126 MB L2
  int add_f32;              // `add` holds fp32
  int in_f32;               // main input holds fp32 instead of bf16 (tiny deep layers: InstanceNorm over <= 30 pixels)
  View in;                  // main input
  const __nv_bfloat16* add; // optional plain [B,H,W,C], added to main (skip connection)
  int relu_in;              // forward: apply ReLU to main first;  backward: mask d_main by (main > 0)
  const float* sums;        // [B,C,2] raw sum / sum of squares of (main+add), or NULL (no InstanceNorm)
  float eps;
  const float* gamma;
  const float* beta;
  const float* Ta;          // separable timing table [(H+W), C/2] or NULL
  __nv_bfloat16* out1;      // optional plain [B,H,W,C]: value after Ta (skip / h)
  float p;                  // dropout probability (0 = none)
  unsigned long long seed;
  const unsigned long long* seed_ptr;  // optional device-resident seed (added to `seed`): CUDA-graph replays
  unsigned int stream_id;
  const uint8_t* keep;      // parity mode: explicit keep mask, uint8 [B,H,W,C]
  const float* Tb;          // separable timing table [(H+W), C/2] or NULL
  const float* sc;          // [B,C] or NULL: x = x*sc + sh  (ActionInjector with sc = sigmoid(dense1(a)), sh = dense2(a))
  const float* sh;
  View out2;
  // backward only
  View g2;                  // grad wrt out2 (same layout as out2)
  const __nv_bfloat16* g1;  // grad wrt out1 (plain) or NULL
  float* bsums;             // [B,C,4]: sum gx, sum gx*xhat, sum g2*xd, sum g2
  View d_main;              // grad wrt main (same layout as in)
  __nv_bfloat16* d_add;     // grad wrt add (plain) or NULL
  float* dcol;              // optional [C]: += column sums of d_main (the producing conv's bias gradient)
  float* dgb;               // optional [2,C] (backward reduce): += sum over b of bsums[b,c,0..1] = (d beta | d gamma)
  // host-precomputed
  uint32_t w_magic;         // fast_div magic of W
  uint32_t nx_magic;        // fast_div magic of the iteration-space width (see iter_space)
  int moments;              // `sums` holds (mean, biased centred variance) from the two-pass small-plane kernel
};

__device__ __forceinline__ uint4 ldg16(const void* p) { return __ldg(reinterpret_cast<const uint4*>(p)); }
__device__ __forceinline__ void unpack8(const uint4& r, float* f) {
  f[0] = __uint_as_float(r.x << 16); f[1] = __uint_as_float(r.x & 0xffff0000u);
  f[2] = __uint_as_float(r.y << 16); f[3] = __uint_as_float(r.y & 0xffff0000u);
  f[4] = __uint_as_float(r.z << 16); f[5] = __uint_as_float(r.z & 0xffff0000u);
  f[6] = __uint_as_float(r.w << 16); f[7] = __uint_as_float(r.w & 0xffff0000u);
}

// keep decisions for 8 consecutive channels: 16 random bits each from one hash4x32 call (counter = vector index)
template <uint32_t F>
__device__ __forceinline__ void keep8(const PrepParams& q, const uint2 key, uint32_t vec_index, uint32_t elem_index,
                                      bool* k) {
  if (F == kRT && q.keep) {
    const uint2 m = *reinterpret_cast<const uint2*>(q.keep + elem_index);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      k[j] = ((m.x >> (8 * j)) & 0xff) != 0;
      k[4 + j] = ((m.y >> (8 * j)) & 0xff) != 0;
    }
    return;
  }
  const uint4 r = hash4x32(vec_index, key);
  const uint32_t thr = (uint32_t)(q.p * 65536.0f);
  k[0] = (r.x & 0xffff) >= thr; k[1] = (r.x >> 16) >= thr;
  k[2] = (r.y & 0xffff) >= thr; k[3] = (r.y >> 16) >= thr;
  k[4] = (r.z & 0xffff) >= thr; k[5] = (r.z >> 16) >= thr;
  k[6] = (r.w & 0xffff) >= thr; k[7] = (r.w >> 16) >= thr;
}
// per-launch dropout key (seed may live on the device so that CUDA-graph replays draw fresh masks)
template <uint32_t F>
__device__ __forceinline__ uint2 dropout_key(const PrepParams& q) {
  if (!has<F, F_DROP>(q.p > 0.f) || (F == kRT && q.keep)) return make_uint2(0u, 0u);
  const unsigned long long seed = q.seed_ptr ? q.seed + *q.seed_ptr : q.seed;
  return hash_key(seed, q.stream_id);
}

// Raw operands of one (pixel, 8-channel vector): loads are ISSUED for several pixels before any is consumed, so that
// a thread keeps 4-8 independent 16-byte loads in flight (these kernels are bound by bytes in flight per SM).
struct RawZ {
  uint4 m, a;
};
template <bool F32, uint32_t F>
__device__ __forceinline__ void issue_z(const PrepParams& q, int b, int y, int x, int c, RawZ& r) {
  if (F32) return;
  r.m = ldg16(reinterpret_cast<const __nv_bfloat16*>(q.in.p) + view_off<F, F_INS2D>(q.in, q.H, q.W, q.C, b, y, x, c));
  if (has<F, F_ADD>(q.add != nullptr)) r.a = ldg16(q.add + plain_off(q.H, q.W, q.C, b, y, x, c));
}
// value of (main [+relu] + add) for 8 channels; also returns raw main
template <bool F32, uint32_t F>
__device__ __forceinline__ void finish_z(const PrepParams& q, const RawZ& r, int b, int y, int x, int c, float* z,
                                         float* mainv) {
  float a[8];
  const bool has_add = has<F, F_ADD>(q.add != nullptr);
  if (F32) {
    const uint32_t off = view_off<F, F_INS2D>(q.in, q.H, q.W, q.C, b, y, x, c);
    if (q.in_f32) ld8f(reinterpret_cast<const float*>(q.in.p) + off, mainv);
    else bf8_to_f(*reinterpret_cast<const BF8*>(reinterpret_cast<const __nv_bfloat16*>(q.in.p) + off), mainv);
    if (has_add) {
      const uint32_t aoff = plain_off(q.H, q.W, q.C, b, y, x, c);
      if (q.add_f32) ld8f(reinterpret_cast<const float*>(q.add) + aoff, a);
      else bf8_to_f(*reinterpret_cast<const BF8*>(q.add + aoff), a);
    }
  } else {
    unpack8(r.m, mainv);
    if (has_add) unpack8(r.a, a);
  }
  const bool relu = has<F, F_RELU>(q.relu_in != 0);
#pragma unroll
  for (int j = 0; j < 8; ++j) z[j] = relu ? fmaxf(mainv[j], 0.f) : mainv[j];
  if (has_add) {
#pragma unroll
    for (int j = 0; j < 8; ++j) z[j] += a[j];
  }
}

// Timing signals are separable (simple/utils.py:129-154): the first C/2 channels depend on the row only, the others
// on the column only.  Table layout: [(H + W), C/2] = rows for y = 0..H-1, then rows for x = 0..W-1.
__device__ __forceinline__ void add_timing(const float* T, int H, int C, int y, int x, int c, float* v) {
  const int half = C >> 1;
  const float* t = (c < half) ? T + (long long)y * half + c : T + (long long)(H + x) * half + (c - half);
  float tv[8];
  ld8f(t, tv);
#pragma unroll
  for (int j = 0; j < 8; ++j) v[j] += tv[j];
}

// Per-(b,c) constants of one sample live in SHARED memory (8 arrays of C floats), not registers: the fused chain
// needs up to 64 of them per thread, which capped occupancy at 2 CTAs/SM and left the kernels latency-bound.
struct ChanCst {
  const float *mean, *rstd, *gamma, *beta, *sc, *sh, *m1, *m2;
};
__device__ __forceinline__ ChanCst load_cst(const PrepParams& q, int b, float* sm, bool bwd) {
  const int C = q.C;
  const float invP = 1.0f / (float)(q.H * q.W);
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float mean = 0.f, rstd = 1.f, g = 1.f, bt = 0.f, m1 = 0.f, m2 = 0.f;
    if (q.sums) {
      const float s1 = q.sums[((long long)b * C + c) * 2], s2 = q.sums[((long long)b * C + c) * 2 + 1];
      float var;
      if (q.moments) {  // two-pass statistics of a tiny plane (stats_small_kernel)
        mean = s1;
        var = s2;
      } else {
        mean = s1 * invP;
        var = fmaxf(s2 * invP - mean * mean, 0.f);
      }
      rstd = 1.0f / __fsqrt_rn(var + q.eps);
      g = q.gamma[c];
      bt = q.beta[c];
      if (bwd && q.bsums) {
        m1 = q.bsums[((long long)b * C + c) * 4] * invP;
        m2 = q.bsums[((long long)b * C + c) * 4 + 1] * invP;
      }
    }
    sm[c] = mean; sm[C + c] = rstd; sm[2 * C + c] = g; sm[3 * C + c] = bt;
    sm[4 * C + c] = q.sc ? q.sc[(long long)b * C + c] : 1.f;
    sm[5 * C + c] = q.sh ? q.sh[(long long)b * C + c] : 0.f;
    sm[6 * C + c] = m1; sm[7 * C + c] = m2;
  }
  __syncthreads();
  ChanCst k;
  k.mean = sm; k.rstd = sm + C; k.gamma = sm + 2 * C; k.beta = sm + 3 * C;
  k.sc = sm + 4 * C; k.sh = sm + 5 * C; k.m1 = sm + 6 * C; k.m2 = sm + 7 * C;
  return k;
}

template <uint32_t F>
__device__ __forceinline__ void keep_for(const PrepParams& q, const uint2 key, int b, int y, int x, int c, bool* keep) {
  if (has<F, F_DROP>(q.p > 0.f)) {
    const uint32_t e = plain_off(q.H, q.W, q.C, b, y, x, c);
    keep8<F>(q, key, e >> 3, e, keep);
  } else {
#pragma unroll
    for (int j = 0; j < 8; ++j) keep[j] = true;
  }
}

// forward chain: xhat (normalised), xn (after affine + Ta), xd (after dropout + Tb, before the injector)
template <uint32_t F>
__device__ __forceinline__ void fwd_chain(const PrepParams& q, const ChanCst& k, int y, int x, int c, const float* z,
                                          const bool* keep, float* xhat, float* xn, float* xd) {
  if (has<F, F_NORM>(q.sums != nullptr)) {
    float mean[8], rstd[8], g[8], bt[8];
    ld8f(k.mean + c, mean); ld8f(k.rstd + c, rstd); ld8f(k.gamma + c, g); ld8f(k.beta + c, bt);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      xhat[j] = (z[j] - mean[j]) * rstd[j];
      xn[j] = xhat[j] * g[j] + bt[j];
    }
  } else {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      xhat[j] = z[j];
      xn[j] = z[j];
    }
  }
  if (has<F, F_TA>(q.Ta != nullptr)) add_timing(q.Ta, q.H, q.C, y, x, c, xn);
  if (has<F, F_DROP>(q.p > 0.f)) {
    const float inv_keep = 1.0f / (1.0f - q.p);
#pragma unroll
    for (int j = 0; j < 8; ++j) xd[j] = keep[j] ? xn[j] * inv_keep : 0.f;
  } else {
#pragma unroll
    for (int j = 0; j < 8; ++j) xd[j] = xn[j];
  }
  if (has<F, F_TB>(q.Tb != nullptr)) add_timing(q.Tb, q.H, q.C, y, x, c, xd);
}

// iteration space: logical pixels, extended to the padded extent of `ext` when it is an s2d view
template <uint32_t F, uint32_t BIT>
__device__ __forceinline__ void iter_space(const View& ext, int H, int W, int* y0, int* ny, int* x0, int* nx) {
  if (!has<F, BIT>(ext.kind != 0)) {
    *y0 = 0; *ny = H; *x0 = 0; *nx = W;
  } else {
    *y0 = -ext.pad; *ny = ext.s * ext.H2; *x0 = -ext.pad; *nx = ext.s * ext.W2;
  }
}

// grid = (pixel slabs, B); block = CV * PL threads: thread = (8-channel vector cv, pixel lane pl)
template <bool F32, uint32_t F>
__global__ void __launch_bounds__(256, 2) prep_fwd_kernel(const __grid_constant__ PrepParams q, int PL, int pix_per_slab) {
  constexpr int U = F32 ? 1 : 4;
  extern __shared__ float sm[];
  int y0, ny, x0, nx;
  iter_space<F, F_OUTS2D>(q.out2, q.H, q.W, &y0, &ny, &x0, &nx);
  const int CV = q.C / 8;
  const int cv = threadIdx.x % CV, pl = threadIdx.x / CV;
  const int b = blockIdx.y + q.b_base, c = cv * 8;
  const int NP = ny * nx;
  const int p_begin = blockIdx.x * pix_per_slab;
  const int p_end = min(NP, p_begin + pix_per_slab);
  const ChanCst k = load_cst(q, b, sm, false);
  const uint2 dkey = dropout_key<F>(q);
  for (int pix0 = p_begin + pl; pix0 < p_end; pix0 += U * PL) {
    RawZ raw[U];
    int yy[U], xx[U];
    bool ins[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int pix = pix0 + u * PL;
      const int row = fast_div(pix, q.nx_magic);
      yy[u] = row + y0;
      xx[u] = pix - row * nx + x0;
      ins[u] = pix < p_end && yy[u] >= 0 && yy[u] < q.H && xx[u] >= 0 && xx[u] < q.W;
      if (ins[u]) issue_z<F32, F>(q, b, yy[u], xx[u], c, raw[u]);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (pix0 + u * PL >= p_end) break;
      __nv_bfloat16* o2 = reinterpret_cast<__nv_bfloat16*>(q.out2.p) +
                          view_off<F, F_OUTS2D>(q.out2, q.H, q.W, q.C, b, yy[u], xx[u], c);
      if (!ins[u]) {
        *reinterpret_cast<uint4*>(o2) = make_uint4(0, 0, 0, 0);
        continue;
      }
      float z[8], mainv[8], xhat[8], xn[8], xd[8];
      bool keep[8];
      finish_z<F32, F>(q, raw[u], b, yy[u], xx[u], c, z, mainv);
      keep_for<F>(q, dkey, b, yy[u], xx[u], c, keep);
      fwd_chain<F>(q, k, yy[u], xx[u], c, z, keep, xhat, xn, xd);
      if (has<F, F_OUT1>(q.out1 != nullptr))
        *reinterpret_cast<BF8*>(q.out1 + plain_off(q.H, q.W, q.C, b, yy[u], xx[u], c)) = f_to_bf8(xn);
      if (has<F, F_INJ>(q.sc != nullptr)) {
        float sc[8], sh[8];
        ld8f(k.sc + c, sc); ld8f(k.sh + c, sh);
#pragma unroll
        for (int j = 0; j < 8; ++j) xd[j] = xd[j] * sc[j] + sh[j];
      }
      *reinterpret_cast<BF8*>(o2) = f_to_bf8(xd);
    }
  }
}

// gradient wrt the normalised value (gx) for 8 channels, plus g2
struct RawG {
  uint4 g2, g1;
};
template <uint32_t F>
__device__ __forceinline__ void issue_g(const PrepParams& q, int b, int y, int x, int c, RawG& r) {
  r.g2 = ldg16(reinterpret_cast<const __nv_bfloat16*>(q.g2.p) + view_off<F, F_OUTS2D>(q.g2, q.H, q.W, q.C, b, y, x, c));
  if (has<F, F_G1>(q.g1 != nullptr)) r.g1 = ldg16(q.g1 + plain_off(q.H, q.W, q.C, b, y, x, c));
}
template <uint32_t F>
__device__ __forceinline__ void finish_gx(const PrepParams& q, const ChanCst& k, const RawG& r, int c, const bool* keep,
                                          float* g2, float* gx) {
  unpack8(r.g2, g2);
  const float inv_keep = has<F, F_DROP>(q.p > 0.f) ? 1.0f / (1.0f - q.p) : 1.0f;
  if (has<F, F_INJ>(q.sc != nullptr)) {
    float sc[8];
    ld8f(k.sc + c, sc);
#pragma unroll
    for (int j = 0; j < 8; ++j) gx[j] = keep[j] ? g2[j] * sc[j] * inv_keep : 0.f;
  } else {
#pragma unroll
    for (int j = 0; j < 8; ++j) gx[j] = keep[j] ? g2[j] * inv_keep : 0.f;
  }
  if (has<F, F_G1>(q.g1 != nullptr)) {
    float t[8];
    unpack8(r.g1, t);
#pragma unroll
    for (int j = 0; j < 8; ++j) gx[j] += t[j];
  }
}

// reduce over pixel lanes through shared memory, then one atomic per (channel, value): scratch[pl][cv][NV*8]
template <int NV>
__device__ __forceinline__ void lane_reduce_atomic(float (*acc)[8], float* scratch, int CV, int PL, int cv, int pl,
                                                   float* out, long long out_base, int out_stride_c,
                                                   float* out2 = nullptr) {
  float* mine = scratch + ((size_t)pl * CV + cv) * (NV * 8);
#pragma unroll
  for (int v = 0; v < NV; ++v)
#pragma unroll
    for (int j = 0; j < 8; ++j) mine[v * 8 + j] = acc[v][j];
  __syncthreads();
  for (int i = threadIdx.x; i < CV * NV * 8; i += blockDim.x) {
    const int cvv = i / (NV * 8), rem = i % (NV * 8);
    const int v = rem / 8, j = rem % 8;
    float s = 0.f;
    for (int l = 0; l < PL; ++l) s += scratch[((size_t)l * CV + cvv) * (NV * 8) + rem];
    atomicAdd(out + out_base + (long long)(cvv * 8 + j) * out_stride_c + v, s);
    if (out2 && v < 2) atomicAdd(out2 + v * (CV * 8) + cvv * 8 + j, s);  // batch-summed affine gradients [2][C]
  }
}

// ---------------------------------------------------------------------------------------------------------------
// per-(b,c) reductions over pixels.  MODE 0: sum z, sum z^2 (2 values).  MODE 1: backward sums (4 values).
// grid = (slabs, B); block = CV * PL threads (thread = one 8-channel vector x one pixel lane)
// shared memory: [8*C constants (MODE 1)] then the [PL][CV][NV*8] reduction scratch
// ---------------------------------------------------------------------------------------------------------------
template <int MODE, bool F32, uint32_t F>
__global__ void __launch_bounds__(256, (MODE == 0 ? 3 : 2))
    reduce_bc_kernel(const __grid_constant__ PrepParams q, float* out, int PL, int pix_per_slab) {
  constexpr int NV = MODE == 0 ? 2 : 4;
  constexpr int U = F32 ? 1 : (MODE == 0 ? 4 : 2);
  extern __shared__ float sm[];
  const int CV = q.C / 8;
  const int cv = threadIdx.x % CV, pl = threadIdx.x / CV;
  const int b = blockIdx.y + q.b_base;
  const int c = cv * 8;
  const int P = q.H * q.W;
  const int p_begin = blockIdx.x * pix_per_slab;
  const int p_end = min(P, p_begin + pix_per_slab);
  float acc[NV][8];
#pragma unroll
  for (int v = 0; v < NV; ++v)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[v][j] = 0.f;
  ChanCst k;
  float* scratch = sm;
  if (MODE == 1) {
    k = load_cst(q, b, sm, false);
    scratch = sm + 8 * q.C;
  }
  const uint2 dkey = MODE == 1 ? dropout_key<F>(q) : make_uint2(0u, 0u);
  for (int pix0 = p_begin + pl; pix0 < p_end; pix0 += U * PL) {
    RawZ raw[U];
    RawG rg[U];
    int yy[U], xx[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int pix = pix0 + u * PL;
      yy[u] = fast_div(pix, q.w_magic);
      xx[u] = pix - yy[u] * q.W;
      if (pix < p_end) {
        issue_z<F32, F>(q, b, yy[u], xx[u], c, raw[u]);
        if (MODE == 1) issue_g<F>(q, b, yy[u], xx[u], c, rg[u]);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (pix0 + u * PL >= p_end) break;
      float z[8], mainv[8];
      finish_z<F32, F>(q, raw[u], b, yy[u], xx[u], c, z, mainv);
      if (MODE == 0) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          acc[0][j] += z[j];
          acc[1][j] += z[j] * z[j];
        }
      } else {
        float xhat[8], xn[8], xd[8], g2[8], gx[8];
        bool keep[8];
        keep_for<F>(q, dkey, b, yy[u], xx[u], c, keep);
        fwd_chain<F>(q, k, yy[u], xx[u], c, z, keep, xhat, xn, xd);
        finish_gx<F>(q, k, rg[u], c, keep, g2, gx);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          acc[0][j] += gx[j];
          acc[1][j] += gx[j] * xhat[j];
          acc[2][j] += g2[j] * xd[j];
          acc[3][j] += g2[j];
        }
      }
    }
  }
  lane_reduce_atomic<NV>(acc, scratch, CV, PL, cv, pl, out, (long long)b * q.C * NV, NV, MODE == 1 ? q.dgb : nullptr);
}

template <bool F32, uint32_t F>
__global__ void __launch_bounds__(256, 2)
    prep_bwd_apply_kernel(const __grid_constant__ PrepParams q, int PL, int pix_per_slab) {
  constexpr int U = F32 ? 1 : 2;
  extern __shared__ float sm[];
  int y0, ny, x0, nx;
  iter_space<F, F_INS2D>(q.d_main, q.H, q.W, &y0, &ny, &x0, &nx);
  const int CV = q.C / 8;
  const int cv = threadIdx.x % CV, pl = threadIdx.x / CV;
  const int b = blockIdx.y + q.b_base, c = cv * 8;
  const int NP = ny * nx;
  const int p_begin = blockIdx.x * pix_per_slab;
  const int p_end = min(NP, p_begin + pix_per_slab);
  const ChanCst k = load_cst(q, b, sm, true);
  const uint2 dkey = dropout_key<F>(q);
  const bool has_dcol = has<F, F_DCOL>(q.dcol != nullptr);
  float dsum[1][8];
#pragma unroll
  for (int j = 0; j < 8; ++j) dsum[0][j] = 0.f;
  for (int pix0 = p_begin + pl; pix0 < p_end; pix0 += U * PL) {
    RawZ raw[U];
    RawG rg[U];
    int yy[U], xx[U];
    bool ins[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int pix = pix0 + u * PL;
      const int row = fast_div(pix, q.nx_magic);
      yy[u] = row + y0;
      xx[u] = pix - row * nx + x0;
      ins[u] = pix < p_end && yy[u] >= 0 && yy[u] < q.H && xx[u] >= 0 && xx[u] < q.W;
      if (ins[u]) {
        issue_z<F32, F>(q, b, yy[u], xx[u], c, raw[u]);
        issue_g<F>(q, b, yy[u], xx[u], c, rg[u]);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (pix0 + u * PL >= p_end) break;
      __nv_bfloat16* dm = reinterpret_cast<__nv_bfloat16*>(q.d_main.p) +
                          view_off<F, F_INS2D>(q.d_main, q.H, q.W, q.C, b, yy[u], xx[u], c);
      if (!ins[u]) {
        *reinterpret_cast<uint4*>(dm) = make_uint4(0, 0, 0, 0);
        continue;
      }
      float z[8], mainv[8], g2[8], gx[8], dz[8];
      bool keep[8];
      finish_z<F32, F>(q, raw[u], b, yy[u], xx[u], c, z, mainv);
      keep_for<F>(q, dkey, b, yy[u], xx[u], c, keep);
      finish_gx<F>(q, k, rg[u], c, keep, g2, gx);
      if (has<F, F_NORM>(q.sums != nullptr)) {  // InstanceNorm backward; the affine / timing / dropout values themselves are not needed here
        float mean[8], rstd[8], g[8], m1[8], m2[8];
        ld8f(k.mean + c, mean); ld8f(k.rstd + c, rstd); ld8f(k.gamma + c, g); ld8f(k.m1 + c, m1); ld8f(k.m2 + c, m2);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float xhat = (z[j] - mean[j]) * rstd[j];
          dz[j] = rstd[j] * g[j] * (gx[j] - m1[j] - xhat * m2[j]);
        }
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) dz[j] = gx[j];
      }
      if (has<F, F_DADD>(q.d_add != nullptr))
        *reinterpret_cast<BF8*>(q.d_add + plain_off(q.H, q.W, q.C, b, yy[u], xx[u], c)) = f_to_bf8(dz);
      if (has<F, F_RELU>(q.relu_in != 0)) {
#pragma unroll
        for (int j = 0; j < 8; ++j) dz[j] = mainv[j] > 0.f ? dz[j] : 0.f;
      }
      const BF8 o = f_to_bf8(dz);
      *reinterpret_cast<BF8*>(dm) = o;
      if (has_dcol) {  // bias gradient of the conv that produced `main`: sums of the bf16-rounded values it will see
        float r8[8];
        bf8_to_f(o, r8);
#pragma unroll
        for (int j = 0; j < 8; ++j) dsum[0][j] += r8[j];
      }
    }
  }
  if (has_dcol) {
    __syncthreads();  // the per-channel constants in shared memory are dead: reuse the space as reduction scratch
    lane_reduce_atomic<1>(dsum, sm, CV, PL, cv, pl, q.dcol, 0, 1);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// InstanceNorm statistics of TINY planes (H*W <= 64: the 6x5 / 3x2 bottleneck layers).  eps = 1e-6 and planes of 6
// pixels make these norms ill-conditioned: a ReLU-sparse channel has a variance of 1e-8..1e-6, where the single-pass
// E[z^2] - E[z]^2 of the streaming kernel (fp32 atomics over pixel lanes) is mostly cancellation noise -- and rstd ~ 1000
// turns that noise into O(1e-2) errors of the normalised values the value head reads.  Here one thread owns 8 channels of
// one sample and makes TWO passes over the <= 64 pixels (mean, then centred sum of squares), like ATen's instance norm:
// out[b,c] = (mean, biased variance).  No atomics: run-to-run deterministic.
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) stats_small_kernel(const __grid_constant__ PrepParams q, float* out, int nb) {
  // one WARP per (sample, 8 channels): the lanes own the <= 64 pixels (two each at most), read once and kept in
  // registers for both passes; fixed-order butterfly sums (round 2b: one THREAD per group took 13-29 us per launch on
  // these 6- and 30-pixel planes -- a serial chain of dependent loads; now ~3 us)
  const int CV = q.C / 8;
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (w >= nb * CV) return;
  const int b = w / CV + q.b_base, c = (w % CV) * 8;
  const int P = q.H * q.W;
  float z[2][8];
  RawZ raw;
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int pix = lane + 32 * h;
    if (pix < P) {
      const int y = pix / q.W, x = pix - y * q.W;
      float mainv[8];
      finish_z<true, kRT>(q, raw, b, y, x, c, z[h], mainv);
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) z[h][j] = 0.f;
    }
  }
  const float invP = 1.0f / (float)P;
  float mean[8], var[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) mean[j] = warp_sum(z[0][j] + z[1][j]) * invP;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float d0 = (lane < P) ? z[0][j] - mean[j] : 0.f, d1 = (lane + 32 < P) ? z[1][j] - mean[j] : 0.f;
    var[j] = warp_sum(d0 * d0 + d1 * d1) * invP;
  }
  if (lane < 8) {
    float m = mean[0], v = var[0];
#pragma unroll
    for (int j = 1; j < 8; ++j)
      if (lane == j) { m = mean[j]; v = var[j]; }
    *reinterpret_cast<float2*>(out + ((long long)b * q.C + c + lane) * 2) = make_float2(m, v);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// standardize_frame (simple/utils.py:122-126): (x - mean) / max(std_unbiased, 1/sqrt(P)) per (b, channel).
// Two kernels: per-plane statistics (one CTA per (b,c): a two-pass mean / centred-variance over the plane staged in
// shared memory, HBM read once), then an apply pass whose threads own PIXELS: the NHWC bf16 destination receives the C
// channels of a pixel as one contiguous run (the former one-kernel version scattered 2-byte stores at a 160-byte stride:
// 16x write amplification).
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
    standardize_stats_kernel(const float* __restrict__ in, int P, float2* __restrict__ stats) {
  extern __shared__ float plane[];
  __shared__ float red[8];
  __shared__ float s_mean;
  const int bc = blockIdx.x;
  const float* src = in + (long long)bc * P;
  float s = 0.f;
  for (int i = threadIdx.x; i < P; i += blockDim.x) {
    const float v = src[i];
    plane[i] = v;
    s += v;
  }
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < 8; ++i) t += red[i];
    s_mean = t / (float)P;
  }
  __syncthreads();
  const float mean = s_mean;
  float v2 = 0.f;
  for (int i = threadIdx.x; i < P; i += blockDim.x) {
    const float d = plane[i] - mean;
    v2 += d * d;
  }
  v2 = warp_sum(v2);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v2;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < 8; ++i) t += red[i];
    const float var = t / (float)(P - 1);  // unbiased (torch.var default)
    stats[bc] = make_float2(mean, 1.0f / fmaxf(__fsqrt_rn(var), rsqrtf((float)P)));
  }
}

// thread = one pixel of one sample, all CT channels: CT independent coalesced plane loads in flight, the CT bf16 results
// written with 16/8-byte stores when the destination allows (12 channels at channel offset 64: one 16 B + one 8 B
// store instead of twelve scattered 2-byte stores).  CT = 0: runtime channel count (<= 16), scalar stores.
template <int CT>
__global__ void __launch_bounds__(256)
    standardize_apply_kernel(const float* __restrict__ in, const float2* __restrict__ stats, int C, int P,
                             __nv_bfloat16* dst, int ld, int coff, float* out_f32, __nv_bfloat16* dst2, int ld2,
                             int coff2) {
  const int b = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P) return;
  if (CT == 0) {
    for (int c = 0; c < C; ++c) {
      const float2 st = stats[b * C + c];
      const float v = (in[((long long)b * C + c) * P + i] - st.x) * st.y;  // coalesced across the pixels of a plane
      if (dst) dst[((long long)b * P + i) * ld + coff + c] = __float2bfloat16_rn(v);
      if (dst2) dst2[((long long)b * P + i) * ld2 + coff2 + c] = __float2bfloat16_rn(v);
      if (out_f32) out_f32[((long long)b * C + c) * P + i] = v;
    }
    return;
  }
  constexpr int CC = CT == 0 ? 1 : CT;
  float v[CC];
#pragma unroll
  for (int c = 0; c < CC; ++c) v[c] = in[((long long)b * CC + c) * P + i];
#pragma unroll
  for (int c = 0; c < CC; ++c) {
    const float2 st = stats[b * CC + c];
    v[c] = (v[c] - st.x) * st.y;
  }
  if (out_f32) {
#pragma unroll
    for (int c = 0; c < CC; ++c) out_f32[((long long)b * CC + c) * P + i] = v[c];
  }
#pragma unroll
  for (int which = 0; which < 2; ++which) {
    __nv_bfloat16* base = which == 0 ? dst : dst2;
    if (!base) continue;
    __nv_bfloat16* o = base + ((long long)b * P + i) * (which == 0 ? ld : ld2) + (which == 0 ? coff : coff2);
    if (CC % 4 == 0 && ((reinterpret_cast<uintptr_t>(o) & 7) == 0)) {
#pragma unroll
      for (int c = 0; c < CC; c += 4) {
        __nv_bfloat162 t0 = __floats2bfloat162_rn(v[c], v[c + 1]), t1 = __floats2bfloat162_rn(v[c + 2], v[c + 3]);
        uint2 pk;
        pk.x = *reinterpret_cast<uint32_t*>(&t0);
        pk.y = *reinterpret_cast<uint32_t*>(&t1);
        *reinterpret_cast<uint2*>(o + c) = pk;
      }
    } else {
#pragma unroll
      for (int c = 0; c < CC; ++c) o[c] = __float2bfloat16_rn(v[c]);
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// conv-GRU blend of the recurrent state (next_frame_predictor.py:298-303)
//   G bf16 [NPIX, 128] = gate conv output (0..63 gate, 64..127 candidate); state fp32 [NPIX, 64]
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
    gru_blend_kernel(const __nv_bfloat16* __restrict__ G, const float* __restrict__ s_old, float* s_new,
                     __nv_bfloat16* xs, int xs_ld, long long npix) {
  const long long total = npix * 8;  // 8 vectors of 8 channels
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long pix = i >> 3;
    const int c = (int)(i & 7) * 8;
    float g[8], cand[8], s[8], o[8];
    bf8_to_f(*reinterpret_cast<const BF8*>(G + pix * 128 + c), g);
    bf8_to_f(*reinterpret_cast<const BF8*>(G + pix * 128 + 64 + c), cand);
    ld8f(s_old + pix * 64 + c, s);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float sg = 1.0f / (1.0f + __expf(-g[j]));
      const float tc = tanh_acc(cand[j]);
      o[j] = s[j] * sg + tc * (1.0f - sg);
    }
    *reinterpret_cast<float4*>(s_new + pix * 64 + c) = make_float4(o[0], o[1], o[2], o[3]);
    *reinterpret_cast<float4*>(s_new + pix * 64 + c + 4) = make_float4(o[4], o[5], o[6], o[7]);
    *reinterpret_cast<BF8*>(xs + pix * xs_ld + c) = f_to_bf8(o);
  }
}

// dG from d(state_new): d/dgate = sg*(1-sg)*(s - tanh(c)),  d/dcand = (1-sg)*(1-tanh(c)^2).  s_old carries no grad.
// The old state is read from its bf16 copy inside the previous step's gate operand (row stride s_ld): the fp32 state
// buffer has already been updated in place by the forward.
__global__ void __launch_bounds__(256)
    gru_blend_bwd_kernel(const __nv_bfloat16* __restrict__ G, const __nv_bfloat16* __restrict__ s_old, int s_ld,
                         const __nv_bfloat16* __restrict__ dxs, int dxs_ld, __nv_bfloat16* dG, long long npix) {
  const long long total = npix * 8;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long pix = i >> 3;
    const int c = (int)(i & 7) * 8;
    float g[8], cand[8], s[8], d[8], dg[8], dc[8];
    bf8_to_f(*reinterpret_cast<const BF8*>(G + pix * 128 + c), g);
    bf8_to_f(*reinterpret_cast<const BF8*>(G + pix * 128 + 64 + c), cand);
    bf8_to_f(*reinterpret_cast<const BF8*>(s_old + pix * s_ld + c), s);
    bf8_to_f(*reinterpret_cast<const BF8*>(dxs + pix * dxs_ld + c), d);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float sg = 1.0f / (1.0f + __expf(-g[j]));
      const float tc = tanh_acc(cand[j]);
      dg[j] = d[j] * sg * (1.0f - sg) * (s[j] - tc);
      dc[j] = d[j] * (1.0f - sg) * (1.0f - tc * tc);
    }
    *reinterpret_cast<BF8*>(dG + pix * 128 + c) = f_to_bf8(dg);
    *reinterpret_cast<BF8*>(dG + pix * 128 + 64 + c) = f_to_bf8(dc);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// weight packing.  W fp32 [O, I, KH, KW] (Conv2d layout; a ConvTranspose2d weight [in,out,kh,kw] is passed as-is with
// O=in, I=out).  s = space-to-depth factor (stride), taps T = (KH/s)*(KW/s), phases PH = s*s, ky = s*dy + py.
//   fwd[operm(o), (t*PH + ph)*Ipad + iperm(i)] = W[o,i,ky,kx]        (bf16, [Opad, T*PH*Ipad])
//   tr [ph*Ipad + iperm(i), t*Opad + operm(o)] = W[o,i,ky,kx]        (bf16, [PH*Ipad, T*Opad])
// Both destinations are zero-filled by the caller once (padding never changes).
// ---------------------------------------------------------------------------------------------------------------
struct PackParams {
  const float* w;
  int O, I, KH, KW, s;
  int Ipad, Opad;
  const int* iperm;
  const int* operm;
  __nv_bfloat16* fwd;
  __nv_bfloat16* tr;
};
// one thread per (o,i) patch: reads KH*KW contiguous fp32, writes bf16 at stride Ipad (coalesced across threads in i)
__global__ void __launch_bounds__(256) pack_conv_kernel(const __grid_constant__ PackParams q) {
  const int KK = q.KH * q.KW;
  const long long npatch = (long long)q.O * q.I;
  const int TW = q.KW / q.s, PH = q.s * q.s, T = (q.KH / q.s) * TW;
  const long long ldf = (long long)T * PH * q.Ipad;
  for (long long patch = (long long)blockIdx.x * blockDim.x + threadIdx.x; patch < npatch;
       patch += (long long)gridDim.x * blockDim.x) {
    const int o = (int)(patch / q.I), i = (int)(patch - (long long)o * q.I);
    const int oo = q.operm ? q.operm[o] : o;
    const int ii = q.iperm ? q.iperm[i] : i;
    const float* src = q.w + patch * KK;
    __nv_bfloat16* dst = q.fwd + (long long)oo * ldf + ii;
    for (int kk = 0; kk < KK; ++kk) {
      const int ky = kk / q.KW, kx = kk - ky * q.KW;
      const int dy = ky / q.s, py = ky - dy * q.s, dx = kx / q.s, px = kx - dx * q.s;
      dst[(long long)((dy * TW + dx) * PH + py * q.s + px) * q.Ipad] = __float2bfloat16_rn(src[kk]);
    }
  }
}

// inverse of the above for gradients: packed fp32 [Opad, T*PH*Ipad] -> reference layout [O, I, KH, KW]
__global__ void __launch_bounds__(256) unpack_conv_kernel(const __grid_constant__ PackParams q, const float* packed,
                                                          float* out) {
  const int KK = q.KH * q.KW;
  const long long npatch = (long long)q.O * q.I;
  const int TW = q.KW / q.s, PH = q.s * q.s, T = (q.KH / q.s) * TW;
  const long long ldf = (long long)T * PH * q.Ipad;
  for (long long patch = (long long)blockIdx.x * blockDim.x + threadIdx.x; patch < npatch;
       patch += (long long)gridDim.x * blockDim.x) {
    const int o = (int)(patch / q.I), i = (int)(patch - (long long)o * q.I);
    const int oo = q.operm ? q.operm[o] : o;
    const int ii = q.iperm ? q.iperm[i] : i;
    const float* src = packed + (long long)oo * ldf + ii;
    float* dst = out + patch * KK;
    for (int kk = 0; kk < KK; ++kk) {
      const int ky = kk / q.KW, kx = kk - ky * q.KW;
      const int dy = ky / q.s, py = ky - dy * q.s, dx = kx / q.s, px = kx - dx * q.s;
      dst[kk] = src[(long long)((dy * TW + dx) * PH + py * q.s + px) * q.Ipad];
    }
  }
}

// tr[:, t*Opad:(t+1)*Opad] = fwd[:, t*PHI:(t+1)*PHI]^T for every tap t.  64x64 bf16 tiles through shared memory, 16-byte
// global accesses on both sides (full 128-byte row segments per 8 threads); the previous 32x32 / 2-byte version moved
// 64-byte segments and ran at ~1.1 TB/s over the 137 MB of operands.
__device__ __forceinline__ void transpose_tile(unsigned short (*tile)[66], const __nv_bfloat16* __restrict__ fwd,
                                               __nv_bfloat16* __restrict__ tr, int Opad, int PHI, int T, int bx, int by,
                                               int t) {
  const unsigned short* src = reinterpret_cast<const unsigned short*>(fwd) + (long long)t * PHI;
  unsigned short* dst = reinterpret_cast<unsigned short*>(tr) + (long long)t * Opad;
  const long long lds = (long long)T * PHI, ldd = (long long)T * Opad;
  const int c0 = bx * 64, r0 = by * 64;
  const int cv = (threadIdx.x & 7) * 8, rr = threadIdx.x >> 3;  // 8 threads x 8 halves per 64-wide row, 32 rows per pass
  const bool vec_ok = (PHI % 8 == 0) && (Opad % 8 == 0);
#pragma unroll
  for (int pass = 0; pass < 2; ++pass) {
    const int r = r0 + rr + pass * 32, c = c0 + cv;
    unsigned short v[8];
    if (vec_ok && r < Opad && c + 8 <= PHI) {
      const uint4 q = *reinterpret_cast<const uint4*>(src + (long long)r * lds + c);
      const unsigned short* h = reinterpret_cast<const unsigned short*>(&q);
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = h[j];
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = (r < Opad && c + j < PHI) ? src[(long long)r * lds + c + j] : (unsigned short)0;
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) tile[rr + pass * 32][cv + j] = v[j];
  }
  __syncthreads();
#pragma unroll
  for (int pass = 0; pass < 2; ++pass) {
    const int c = c0 + rr + pass * 32;  // destination row = source column
    const int r = r0 + cv;              // destination columns r .. r+7 = source rows
    unsigned short v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = tile[cv + j][rr + pass * 32];
    if (c < PHI) {
      if (vec_ok && r + 8 <= Opad) {
        uint4 q;
        unsigned short* h = reinterpret_cast<unsigned short*>(&q);
#pragma unroll
        for (int j = 0; j < 8; ++j) h[j] = v[j];
        *reinterpret_cast<uint4*>(dst + (long long)c * ldd + r) = q;
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j)
          if (r + j < Opad) dst[(long long)c * ldd + r + j] = v[j];
      }
    }
  }
}

__global__ void __launch_bounds__(256)
    transpose_taps_kernel(const __nv_bfloat16* __restrict__ fwd, __nv_bfloat16* __restrict__ tr, int Opad, int PHI,
                          int T) {
  __shared__ unsigned short tile[64][66];  // [source row][source col], pitch 66 halves = 33 words (odd: conflict-free)
  transpose_tile(tile, fwd, tr, Opad, PHI, T, blockIdx.x, blockIdx.y, blockIdx.z);
}
// All layers of the network in ONE launch (the per-layer launches of the small layers were 2..24 CTAs each): CTA ->
// (layer, tile) through the prefix sums of the layers' tile counts.
static_assert(sizeof(sb_transpose_desc) == 32, "sb_transpose_desc: the Python side builds 32-byte records");
__global__ void __launch_bounds__(256)
    transpose_taps_multi_kernel(const sb_transpose_desc* __restrict__ descs, int n) {
  __shared__ unsigned short tile[64][66];
  int l = 0;
  while (l + 1 < n && (int)blockIdx.x >= descs[l + 1].first_tile) ++l;
  const sb_transpose_desc d = descs[l];
  int w = blockIdx.x - d.first_tile;
  const int nx = (d.PHI + 63) / 64, ny = (d.Opad + 63) / 64;
  const int bx = w % nx;
  w /= nx;
  const int by = w % ny, t = w / ny;
  transpose_tile(tile, (const __nv_bfloat16*)d.fwd, (__nv_bfloat16*)d.tr, d.Opad, d.PHI, d.T, bx, by, t);
}

// epilogue of a split-K tap GEMM (tiny deep-layer outputs): fp32 partial sums + bias -> ReLU -> bf16 or fp32 (in place
// allowed).  One launch instead of an ATen add / relu / cast chain.
__global__ void __launch_bounds__(256)
    bias_act_kernel(const float* acc, const float* __restrict__ bias, int bias_mod, int relu, float* out_f32,
                    __nv_bfloat16* out_bf16, long long nvec) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (long long)gridDim.x * blockDim.x) {
    float4 v = reinterpret_cast<const float4*>(acc)[i];
    if (bias) {
      const float4 bv = *reinterpret_cast<const float4*>(bias + (int)((i * 4) % bias_mod));
      v.x += bv.x; v.y += bv.y; v.z += bv.z; v.w += bv.w;
    }
    if (relu) {
      v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f);
    }
    if (out_f32) {
      reinterpret_cast<float4*>(out_f32)[i] = v;
    } else {
      __nv_bfloat162 t0 = __floats2bfloat162_rn(v.x, v.y), t1 = __floats2bfloat162_rn(v.z, v.w);
      uint2 pk;
      pk.x = *reinterpret_cast<uint32_t*>(&t0);
      pk.y = *reinterpret_cast<uint32_t*>(&t1);
      reinterpret_cast<uint2*>(out_bf16)[i] = pk;
    }
  }
}

static inline int grid_for(long long total, int block = 256) {
  long long g = (total + block - 1) / block;
  const long long cap = 148LL * 16;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace sb

using namespace sb;

// ---- C ABI ------------------------------------------------------------------------------------------------------
static View mk_view(const sb_view* v) {
  View r;
  r.p = v->p; r.kind = v->kind; r.s = v->s; r.pad = v->pad; r.H2 = v->H2; r.W2 = v->W2;
  r.sh = 0;
  while ((1 << r.sh) < r.s) ++r.sh;
  return r;
}
static bool view_ok(const View& v) { return v.kind == 0 || (v.s >= 1 && (1 << v.sh) == v.s); }
static int g_prep_specialize = 1;  // 0: always run the fully-runtime kernel instances (A/B testing)
// feature mask of a parameter block, restricted to the bits a kernel looks at
static uint32_t feat_mask(const PrepParams& q, uint32_t relevant) {
  if (!g_prep_specialize) return 0xffffffffu;  // matches no specialised instance
  uint32_t f = 0;
  if (q.relu_in) f |= F_RELU;
  if (q.sums) f |= F_NORM;
  if (q.Ta) f |= F_TA;
  if (q.out1) f |= F_OUT1;
  if (q.p > 0.f) f |= F_DROP;
  if (q.Tb) f |= F_TB;
  if (q.sc) f |= F_INJ;
  if (q.add) f |= F_ADD;
  if (q.in.kind != 0) f |= F_INS2D;
  if (q.out2.kind != 0 || q.g2.kind != 0) f |= F_OUTS2D;
  if (q.g1) f |= F_G1;
  if (q.d_add) f |= F_DADD;
  if (q.dcol) f |= F_DCOL;
  return f & relevant;
}
// the feature combinations of the world model's bf16 stages (anything else runs the fully-runtime instance)
#define SB_FWD_MASKS(X)                                              \
  X(F_TA | F_OUT1 | F_DROP | F_TB | F_OUTS2D)                        \
  X(F_RELU | F_NORM | F_OUT1 | F_DROP | F_TB | F_OUTS2D)             \
  X(F_RELU | F_NORM | F_TA | F_DROP | F_INJ | F_ADD | F_INS2D)       \
  X(F_RELU | F_NORM | F_TA | F_ADD | F_INS2D)                        \
  X(F_TA | F_INJ | F_OUTS2D)                                         \
  X(F_RELU | F_OUTS2D)
#define SB_STATS_MASKS(X) X(0u) X(F_RELU) X(F_RELU | F_ADD | F_INS2D)
#define SB_RED1_MASKS(X)                                             \
  X(F_RELU | F_NORM | F_DROP | F_TB | F_OUTS2D | F_G1)               \
  X(F_RELU | F_NORM | F_TA | F_DROP | F_INJ | F_ADD | F_INS2D)       \
  X(F_RELU | F_NORM | F_TA | F_ADD | F_INS2D)                        \
  X(F_TA | F_INJ | F_OUTS2D)
#define SB_APPLY_MASKS(X)                                                    \
  X(F_DROP | F_OUTS2D | F_G1 | F_DCOL)                                       \
  X(F_RELU | F_NORM | F_DROP | F_OUTS2D | F_G1 | F_DCOL)                     \
  X(F_RELU | F_NORM | F_DROP | F_INJ | F_ADD | F_INS2D | F_DADD | F_DCOL)    \
  X(F_RELU | F_NORM | F_ADD | F_INS2D | F_DADD | F_DCOL)                     \
  X(F_INJ | F_OUTS2D | F_DCOL)                                               \
  X(F_RELU | F_OUTS2D)
constexpr uint32_t kFwdBits = F_RELU | F_NORM | F_TA | F_OUT1 | F_DROP | F_TB | F_INJ | F_ADD | F_INS2D | F_OUTS2D;
constexpr uint32_t kStatsBits = F_RELU | F_ADD | F_INS2D;
constexpr uint32_t kRed1Bits = F_RELU | F_NORM | F_TA | F_DROP | F_TB | F_INJ | F_ADD | F_INS2D | F_OUTS2D | F_G1;
constexpr uint32_t kApplyBits = F_RELU | F_NORM | F_DROP | F_INJ | F_ADD | F_INS2D | F_OUTS2D | F_G1 | F_DADD | F_DCOL;

static uint32_t div_magic(int d) { return d <= 1 ? 0u : (uint32_t)((0x100000000ULL + (uint64_t)d - 1) / (uint64_t)d); }
static int fill(PrepParams& q, const sb_prep_args* a) {
  if (!a || a->C % 8 || a->C < 8 || a->C / 8 > 256) return SB_ERR_ARG;
  q.B = a->B; q.H = a->H; q.W = a->W; q.C = a->C;
  q.b_base = a->b_base;
  if (a->b_base < 0 || a->b_count < 0 || a->b_base + a->b_count > a->B) return SB_ERR_ARG;
  q.in_f32 = a->in_f32;
  q.add_f32 = a->add_f32;
  q.in = mk_view(&a->in);
  q.add = (const __nv_bfloat16*)a->add;
  q.relu_in = a->relu_in;
  q.sums = a->sums; q.eps = a->eps; q.gamma = a->gamma; q.beta = a->beta;
  q.Ta = a->Ta; q.out1 = (__nv_bfloat16*)a->out1;
  q.p = a->p; q.seed = a->seed; q.seed_ptr = a->seed_ptr; q.stream_id = a->stream_id; q.keep = a->keep;
  q.Tb = a->Tb; q.sc = a->sc; q.sh = a->sh;
  q.out2 = mk_view(&a->out2);
  q.g2 = mk_view(&a->g2);
  q.g1 = (const __nv_bfloat16*)a->g1;
  q.bsums = a->bsums;
  q.d_main = mk_view(&a->d_main);
  q.d_add = (__nv_bfloat16*)a->d_add;
  q.dcol = a->dcol;
  q.dgb = a->dgb;
  q.moments = a->moments;
  if (q.moments && (long long)q.H * q.W > 64) return SB_ERR_ARG;
  if (!view_ok(q.in) || !view_ok(q.out2) || !view_ok(q.g2) || !view_ok(q.d_main)) return SB_ERR_ARG;
  {  // 32-bit element offsets: the largest view (padded space-to-depth extent included) must stay below 2^31 elements
    long long worst = (long long)q.B * q.H * q.W * q.C;
    const View* vs[4] = {&q.in, &q.out2, &q.g2, &q.d_main};
    for (const View* v : vs)
      if (v->kind == 1) {
        const long long e = (long long)q.B * v->H2 * v->W2 * v->s * v->s * q.C;
        if (e > worst) worst = e;
      }
    if (worst >= (1LL << 31)) return SB_ERR_ARG;
  }
  if (q.H > 1 && (q.W >= 300 || (long long)q.H * q.W >= (1 << 17))) return SB_ERR_ARG;  // fast_div range
  q.w_magic = q.H > 1 ? div_magic(q.W) : 1u;  // single-row tensors (column sums of a matrix): row index is always 0
  q.nx_magic = 0;
  if (q.sums && (!q.gamma || !q.beta)) return SB_ERR_ARG;
  if ((q.sc == nullptr) != (q.sh == nullptr)) return SB_ERR_ARG;
  return SB_OK;
}
static int launch_samples(const sb_prep_args* a) { return a->b_count > 0 ? a->b_count : a->B; }
// Grid sizing.  A launch is (slabs, samples) CTAs; what matters is how that count sits against the number of CTAs the
// chip holds at once (`cap` per SM x 148 slots).  Measured on the training step (B = 64, ms per step): a fixed
// "6 CTAs per SM" rule gave 14 x 64 = 896 CTAs on 296 slots -- three full waves plus eight CTAs alone in a fourth --
// 7.86; grids cut to whole waves: 6 waves 8.03, 4 waves 7.86, 3 waves 7.76, 2 waves 7.63, ONE wave 7.61.  Every CTA
// pays a prologue (per-channel constants into shared memory) and an epilogue (block reduction + atomics), so the
// fewest, longest CTAs win: the slab count is the largest that keeps the whole grid co-resident.
// SB_PREP_WAVES=n forces n waves (experiments).
static int g_prep_waves = 0;
static void reduce_cfg(const PrepParams& q, int nb, int cap, int* PL, int* slabs, int* pps, int* threads, int P = -1) {
  const int CV = q.C / 8;
  *PL = 256 / CV < 1 ? 1 : 256 / CV;
  *threads = CV * (*PL);
  if (P < 0) P = q.H * q.W;
  static bool env_read = false;
  if (!env_read) {
    const char* e = getenv("SB_PREP_WAVES");
    if (e) g_prep_waves = atoi(e);
    env_read = true;
  }
  if (cap < 1) cap = 2;
  const int slots = cap * (148 - g_sm_reserve);
  int want = slots * (g_prep_waves > 0 ? g_prep_waves : 1) / nb;
  if (want < 1) want = 1;
  int per = (P + want - 1) / want;
  if (per < *PL) per = *PL;
  *pps = per;
  *slabs = (P + per - 1) / per;
}
// resident CTAs per SM of one kernel instance at this block size / dynamic shared memory (cached per instance)
template <typename K>
static int resident_ctas(K kernel, int threads, size_t smem) {
  struct Entry { int threads; size_t smem; int n; };
  static std::unordered_map<const void*, Entry> cache;
  const void* key = reinterpret_cast<const void*>(kernel);
  auto it = cache.find(key);
  if (it != cache.end() && it->second.threads == threads && it->second.smem == smem) return it->second.n;
  int n = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kernel, threads, smem) != cudaSuccess || n < 1) n = 2;
  cache[key] = Entry{threads, smem, n};
  return n;
}
static void block_cfg(const PrepParams& q, int* PL, int* threads) {
  const int CV = q.C / 8;
  *PL = 256 / CV < 1 ? 1 : 256 / CV;
  *threads = CV * (*PL);
}

extern "C" int sb_set_prep_specialize(int enabled) {
  const int prev = g_prep_specialize;
  g_prep_specialize = enabled ? 1 : 0;
  return prev;
}

extern "C" int sb_stats(const sb_prep_args* a, float* sums, sb_stream_t stream) {
  PrepParams q;
  int rc = fill(q, a);
  if (rc || !sums || !q.in.p) return SB_ERR_ARG;
  q.sums = nullptr;
  int PL, slabs, pps, threads;
  const int nb = launch_samples(a);
  cudaStream_t st = (cudaStream_t)stream;
  if (q.moments) {
    const int total = nb * (q.C / 8);  // one warp each
    stats_small_kernel<<<(total + 3) / 4, 128, 0, st>>>(q, sums, nb);
    ++g_launches;
    SB_CHECK_LAUNCH();
    return SB_OK;
  }
  block_cfg(q, &PL, &threads);
  const size_t smem = (size_t)threads * 2 * 8 * sizeof(float);
  auto go = [&](auto kern) {
    reduce_cfg(q, nb, resident_ctas(kern, threads, smem), &PL, &slabs, &pps, &threads);
    kern<<<dim3(slabs, nb), threads, smem, st>>>(q, sums, PL, pps);
  };
  if (q.in_f32 || q.add_f32) {
    go(reduce_bc_kernel<0, true, kRT>);
  } else {
    switch (feat_mask(q, kStatsBits)) {
#define SB_CASE(M) case (M): go(reduce_bc_kernel<0, false, (M)>); break;
      SB_STATS_MASKS(SB_CASE)
#undef SB_CASE
      default: go(reduce_bc_kernel<0, false, kRT>);
    }
  }
  ++g_launches;
  SB_CHECK_LAUNCH();
  return SB_OK;
}

extern "C" int sb_prep(const sb_prep_args* a, sb_stream_t stream) {
  PrepParams q;
  int rc = fill(q, a);
  if (rc || !q.in.p || !q.out2.p) return SB_ERR_ARG;
  int ny, nx;
  if (q.out2.kind == 0) { ny = q.H; nx = q.W; } else { ny = q.out2.s * q.out2.H2; nx = q.out2.s * q.out2.W2; }
  int PL, slabs, pps, threads;
  const int nb = launch_samples(a);
  block_cfg(q, &PL, &threads);
  if (nx >= 300 || ny * nx >= (1 << 17)) return SB_ERR_ARG;
  q.nx_magic = div_magic(nx);
  const size_t smem = 8 * q.C * sizeof(float);
  cudaStream_t st = (cudaStream_t)stream;
  auto go = [&](auto kern) {
    reduce_cfg(q, nb, resident_ctas(kern, threads, smem), &PL, &slabs, &pps, &threads, ny * nx);
    kern<<<dim3(slabs, nb), threads, smem, st>>>(q, PL, pps);
  };
  if (q.in_f32 || q.add_f32) {
    go(prep_fwd_kernel<true, kRT>);
  } else if (q.keep) {
    go(prep_fwd_kernel<false, kRT>);
  } else {
    switch (feat_mask(q, kFwdBits)) {
#define SB_CASE(M) case (M): go(prep_fwd_kernel<false, (M)>); break;
      SB_FWD_MASKS(SB_CASE)
#undef SB_CASE
      default: go(prep_fwd_kernel<false, kRT>);
    }
  }
  ++g_launches;
  SB_CHECK_LAUNCH();
  return SB_OK;
}

extern "C" int sb_prep_bwd_reduce(const sb_prep_args* a, sb_stream_t stream) {
  PrepParams q;
  int rc = fill(q, a);
  if (rc || !q.in.p || !q.g2.p || !q.bsums) return SB_ERR_ARG;
  int PL, slabs, pps, threads;
  const int nb = launch_samples(a);
  block_cfg(q, &PL, &threads);
  const size_t smem = ((size_t)threads * 4 * 8 + 8 * (size_t)q.C) * sizeof(float);
  cudaStream_t st = (cudaStream_t)stream;
  auto go = [&](auto kern) {
    reduce_cfg(q, nb, resident_ctas(kern, threads, smem), &PL, &slabs, &pps, &threads);
    kern<<<dim3(slabs, nb), threads, smem, st>>>(q, q.bsums, PL, pps);
  };
  if (q.in_f32 || q.add_f32) {
    go(reduce_bc_kernel<1, true, kRT>);
  } else if (q.keep) {
    go(reduce_bc_kernel<1, false, kRT>);
  } else {
    switch (feat_mask(q, kRed1Bits)) {
#define SB_CASE(M) case (M): go(reduce_bc_kernel<1, false, (M)>); break;
      SB_RED1_MASKS(SB_CASE)
#undef SB_CASE
      default: go(reduce_bc_kernel<1, false, kRT>);
    }
  }
  ++g_launches;
  SB_CHECK_LAUNCH();
  return SB_OK;
}

extern "C" int sb_prep_bwd_apply(const sb_prep_args* a, sb_stream_t stream) {
  PrepParams q;
  int rc = fill(q, a);
  if (rc || !q.in.p || !q.g2.p || !q.d_main.p || (q.sums && !q.bsums)) return SB_ERR_ARG;
  int ny, nx;
  if (q.d_main.kind == 0) { ny = q.H; nx = q.W; } else { ny = q.d_main.s * q.d_main.H2; nx = q.d_main.s * q.d_main.W2; }
  int PL, slabs, pps, threads;
  const int nb = launch_samples(a);
  block_cfg(q, &PL, &threads);
  if (nx >= 300 || ny * nx >= (1 << 17)) return SB_ERR_ARG;
  q.nx_magic = div_magic(nx);
  // shared memory: 8*C per-channel constants, reused as the [PL][CV][8] scratch of the bias-gradient reduction
  const size_t smem = (size_t)(8 * q.C > threads * 8 ? 8 * q.C : threads * 8) * sizeof(float);
  cudaStream_t st = (cudaStream_t)stream;
  auto go = [&](auto kern) {
    reduce_cfg(q, nb, resident_ctas(kern, threads, smem), &PL, &slabs, &pps, &threads, ny * nx);
    kern<<<dim3(slabs, nb), threads, smem, st>>>(q, PL, pps);
  };
  if (q.in_f32 || q.add_f32) {
    go(prep_bwd_apply_kernel<true, kRT>);
  } else if (q.keep) {
    go(prep_bwd_apply_kernel<false, kRT>);
  } else {
    switch (feat_mask(q, kApplyBits)) {
#define SB_CASE(M) case (M): go(prep_bwd_apply_kernel<false, (M)>); break;
      SB_APPLY_MASKS(SB_CASE)
#undef SB_CASE
      default: go(prep_bwd_apply_kernel<false, kRT>);
    }
  }
  ++g_launches;
  SB_CHECK_LAUNCH();
  return SB_OK;
}

extern "C" int sb_standardize(const float* in, int B, int C, int P, void* dst_bf16, int ld, int coff, float* out_f32,
                              float* stats_scratch, void* dst2_bf16, int ld2, int coff2, sb_stream_t stream) {
  if (!in || !stats_scratch || B < 1 || C < 1 || C > 16 || P < 2 || P * sizeof(float) > 200 * 1024) return SB_ERR_ARG;
  static bool configured = false;
  if (!configured) {
    if (cudaFuncSetAttribute(standardize_stats_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) !=
        cudaSuccess)
      return SB_ERR_CUDA;
    configured = true;
  }
  cudaStream_t st = (cudaStream_t)stream;
  float2* stats = reinterpret_cast<float2*>(stats_scratch);
  standardize_stats_kernel<<<B * C, 256, P * sizeof(float), st>>>(in, P, stats);
  ++g_launches;
  SB_CHECK_LAUNCH();
  const dim3 grid((P + 255) / 256, B);
  if (C == 12)
    standardize_apply_kernel<12><<<grid, 256, 0, st>>>(in, stats, C, P, (__nv_bfloat16*)dst_bf16, ld, coff, out_f32,
                                                         (__nv_bfloat16*)dst2_bf16, ld2, coff2);
  else if (C == 3)
    standardize_apply_kernel<3><<<grid, 256, 0, st>>>(in, stats, C, P, (__nv_bfloat16*)dst_bf16, ld, coff, out_f32,
                                                         (__nv_bfloat16*)dst2_bf16, ld2, coff2);
  else
    standardize_apply_kernel<0><<<grid, 256, 0, st>>>(in, stats, C, P, (__nv_bfloat16*)dst_bf16, ld, coff, out_f32,
                                                         (__nv_bfloat16*)dst2_bf16, ld2, coff2);
  ++g_launches;
  SB_CHECK_LAUNCH();
  return SB_OK;
}

extern "C" int sb_bias_act(const float* acc, const float* bias, int bias_mod, int relu, void* out, int out_is_f32,
                           long long rows, int N, sb_stream_t stream) {
  if (!acc || !out || rows < 1 || N < 4 || N % 4) return SB_ERR_ARG;
  if (bias_mod <= 0) bias_mod = N;
  if (bias_mod % 4 || N % bias_mod) return SB_ERR_ARG;
  const long long nvec = rows * N / 4;
  bias_act_kernel<<<grid_for(nvec), 256, 0, (cudaStream_t)stream>>>(acc, bias, bias_mod, relu,
                                                                   out_is_f32 ? (float*)out : nullptr,
                                                                   out_is_f32 ? nullptr : (__nv_bfloat16*)out, nvec);
  ++g_launches;
  SB_CHECK_LAUNCH();
  return SB_OK;
}

extern "C" int sb_gru_blend(const void* G, const float* s_old, float* s_new, void* xs, int xs_ld, long long npix,
                            sb_stream_t stream) {
  if (!G || !s_old || !s_new || !xs || xs_ld % 8) return SB_ERR_ARG;
  gru_blend_kernel<<<grid_for(npix * 8), 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)G, s_old, s_new,
                                                                         (__nv_bfloat16*)xs, xs_ld, npix);
  ++g_launches;
  SB_CHECK_LAUNCH();
  return SB_OK;
}

extern "C" int sb_gru_blend_bwd(const void* G, const void* s_old_bf16, int s_ld, const void* dxs, int dxs_ld, void* dG,
                                long long npix, sb_stream_t stream) {
  if (!G || !s_old_bf16 || !dxs || !dG || dxs_ld % 8 || s_ld % 8 || s_ld < 64) return SB_ERR_ARG;
  gru_blend_bwd_kernel<<<grid_for(npix * 8), 256, 0, (cudaStream_t)stream>>>(
      (const __nv_bfloat16*)G, (const __nv_bfloat16*)s_old_bf16, s_ld, (const __nv_bfloat16*)dxs, dxs_ld,
      (__nv_bfloat16*)dG, npix);
  ++g_launches;
  SB_CHECK_LAUNCH();
  return SB_OK;
}

static int fill_pack(PackParams& q, const float* w, int O, int I, int KH, int KW, int s, int Ipad, int Opad,
                     const int* iperm, const int* operm) {
  if (s < 1 || KH % s || KW % s || Ipad < I || Opad < O) return SB_ERR_ARG;
  q.w = w; q.O = O; q.I = I; q.KH = KH; q.KW = KW; q.s = s; q.Ipad = Ipad; q.Opad = Opad;
  q.iperm = iperm; q.operm = operm; q.fwd = nullptr; q.tr = nullptr;
  return SB_OK;
}

extern "C" int sb_pack_conv(const float* w, int O, int I, int KH, int KW, int s, int Ipad, int Opad, const int* iperm,
                            const int* operm, void* fwd, void* tr, sb_stream_t stream) {
  PackParams q;
  if (!w || !fwd || fill_pack(q, w, O, I, KH, KW, s, Ipad, Opad, iperm, operm)) return SB_ERR_ARG;
  q.fwd = (__nv_bfloat16*)fwd;
  pack_conv_kernel<<<grid_for((long long)O * I), 256, 0, (cudaStream_t)stream>>>(q);
  ++g_launches;
  SB_CHECK_LAUNCH();
  if (tr) {
    const int T = (KH / s) * (KW / s), PHI = s * s * Ipad;
    dim3 grid((PHI + 63) / 64, (Opad + 63) / 64, T);
    transpose_taps_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)fwd, (__nv_bfloat16*)tr, Opad,
                                                                  PHI, T);
    ++g_launches;
    SB_CHECK_LAUNCH();
  }
  return SB_OK;
}

extern "C" int sb_transpose_taps(const void* fwd, void* tr, int Opad, int PHI, int T, sb_stream_t stream) {
  if (!fwd || !tr || Opad < 1 || PHI < 1 || T < 1) return SB_ERR_ARG;
  dim3 grid((PHI + 63) / 64, (Opad + 63) / 64, T);
  transpose_taps_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)fwd, (__nv_bfloat16*)tr, Opad,
                                                                PHI, T);
  ++g_launches;
  SB_CHECK_LAUNCH();
  return SB_OK;
}

extern "C" int sb_transpose_taps_multi(const sb_transpose_desc* descs_dev, int n, int total_tiles, sb_stream_t stream) {
  if (!descs_dev || n < 1 || total_tiles < 1) return SB_ERR_ARG;
  transpose_taps_multi_kernel<<<total_tiles, 256, 0, (cudaStream_t)stream>>>(descs_dev, n);
  ++g_launches;
  SB_CHECK_LAUNCH();
  return SB_OK;
}

extern "C" int sb_unpack_conv_grad(const float* packed, int O, int I, int KH, int KW, int s, int Ipad, int Opad,
                                   const int* iperm, const int* operm, float* out, sb_stream_t stream) {
  PackParams q;
  if (!packed || !out || fill_pack(q, nullptr, O, I, KH, KW, s, Ipad, Opad, iperm, operm)) return SB_ERR_ARG;
  unpack_conv_kernel<<<grid_for((long long)O * I), 256, 0, (cudaStream_t)stream>>>(q, packed, out);
  ++g_launches;
  SB_CHECK_LAUNCH();
  return SB_OK;
}
