// Multi-tensor fused clip_grad_norm_ + Adafactor (fairseq defaults) -- K16/K17 of SURVEY 2.3.
// Replaces simple/trainer.py:148 (clip_grad_norm_) and simple/adafactor.py:113-212 (~1800 ATen launches + 120 host
// syncs per step in the reference) by three launches with no host synchronisation:
//   1. sb_grad_sumsq        : sum g^2 over every gradient                     -> clip coefficient (device scalar)
//   2. sb_adafactor_conv    : 4-D weights (O,I,kh,kw).  The factored second moment of fairseq's Adafactor reduces
//                             over kw (row state (O,I,kh)) and kh (col state (O,I,kw)): everything is local to one
//                             (o,i) patch, one thread per patch; pass 1 updates the state and accumulates
//                             sum u^2 / sum p^2, pass 2 applies p -= lr * u / max(1, RMS(u)).
//   3. sb_adafactor_small   : 1-D / 2-D tensors, one CTA per tensor (block-local reductions).
// HBM-bound: ~12 B/param minimum (read p, g; write p) + state; see DESIGN.md.
#include <cooperative_groups.h>
#include "common.cuh"
#include "../../include/simple_b200.h"

namespace sb {
extern long long g_launches;

__device__ __forceinline__ float block_sum(float v, float* red) {
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

__device__ __forceinline__ float clip_coef(const float* total_sumsq, float max_norm) {
  if (max_norm <= 0.f) return 1.0f;
  const float c = max_norm / (__fsqrt_rn(*total_sumsq) + 1e-6f);
  return fminf(c, 1.0f);
}

// Fixed-order sum of `n` floats by one warp (lane-strided partial sums, then the xor-shuffle tree): the result depends on
// the values only, never on which CTA finished first.
__device__ __forceinline__ float warp_ordered_sum(const volatile float* v, int n, int lane) {
  float t = 0.f;
  for (int i = lane; i < n; i += 32) t += v[i];
  return warp_sum(t);
}

// ---- 1. global gradient norm ------------------------------------------------------------------------------------
// Persistent: a resident grid walks the (tensor, 4096-element chunk) list, 16-byte loads.  ORDER-DETERMINISTIC: every CTA
// publishes its partial sum, the CTA that arrives last (ticket counter in scratch[0]) adds the partials in index order.
// Data-parallel replicas therefore compute bit-identical clip coefficients from the bit-identical reduced gradient and
// never drift apart (round 1 used one fp32 atomic per CTA and re-broadcast the weights every window).
__global__ void __launch_bounds__(256)
    grad_sumsq_kernel(const sb_tensor_desc* __restrict__ descs, const long long* __restrict__ glen,
                      const int2* __restrict__ block_map, int nblocks, float* total, float* scratch,
                      float sumsq_scale) {
  __shared__ float red[8];
  __shared__ unsigned s_ticket;
  float s = 0.f;
  for (int blk = blockIdx.x; blk < nblocks; blk += gridDim.x) {
    const int2 bm = block_map[blk];  // (tensor, chunk)
    const float* g = descs[bm.x].g;
    const long long n = glen ? glen[bm.x] : descs[bm.x].n;  // a packed gradient buffer is longer than its parameter
    const long long begin = (long long)bm.y * 4096;
    const long long end = begin + 4096 < n ? begin + 4096 : n;
    if (end - begin == 4096 && (reinterpret_cast<uintptr_t>(g) & 15) == 0) {
      const float4* g4 = reinterpret_cast<const float4*>(g + begin);
      float4 v[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) v[k] = __ldg(g4 + threadIdx.x + 256 * k);
#pragma unroll
      for (int k = 0; k < 4; ++k) s += v[k].x * v[k].x + v[k].y * v[k].y + v[k].z * v[k].z + v[k].w * v[k].w;
    } else {
      for (long long i = begin + threadIdx.x; i < end; i += 256) {
        const float x = g[i];
        s += x * x;
      }
    }
  }
  s = block_sum(s, red);
  if (threadIdx.x == 0) {
    scratch[1 + blockIdx.x] = s;
    __threadfence();
    s_ticket = atomicAdd(reinterpret_cast<unsigned*>(scratch), 1u);
  }
  __syncthreads();
  if (s_ticket == gridDim.x - 1 && threadIdx.x < 32) {
    __threadfence();
    const float t = warp_ordered_sum(scratch + 1, (int)gridDim.x, (int)threadIdx.x);
    if (threadIdx.x == 0) {
      *total = t * sumsq_scale;
      *reinterpret_cast<unsigned*>(scratch) = 0u;  // ready for the next launch / graph replay
    }
  }
}

// ---- 2. conv-type weights ---------------------------------------------------------------------------------------
// PACKED: the gradient is read straight from the packed fp32 buffer the wgrad kernel accumulated into (layout of the
// bf16 GEMM operand `fwd`, see sb_pack_conv) and pass 2 also refreshes that bf16 operand, so that neither an
// unpack-gradient pass nor a re-pack pass over the 68 M weights exists.  Threads of a warp hold consecutive input
// channels i of one output channel o: packed reads / bf16 writes are contiguous across the warp for every tap.
template <int KH, int KW>
__device__ __forceinline__ int packed_tap_offset(int k, int sh, int s, int Ipad) {
  const int ky = k / KW, kx = k % KW;
  const int dy = ky >> sh, py = ky & (s - 1), dx = kx >> sh, px = kx & (s - 1);
  return ((dy * (KW >> sh) + dx) * (s * s) + py * s + px) * Ipad;
}

template <int KH, int KW, int PASS, bool PACKED>
__global__ void __launch_bounds__(128)
    adafactor_conv_kernel(const sb_tensor_desc* __restrict__ descs, const sb_pack_desc* __restrict__ packs,
                          const int2* __restrict__ block_map, const float* __restrict__ total_sumsq, float max_norm,
                          float* scratch, int ntensors) {
  constexpr int KK = KH * KW;
  __shared__ float red[8];
  __shared__ unsigned s_ticket;
  const int2 bm = block_map[blockIdx.x];  // (tensor, first patch)
  const sb_tensor_desc d = descs[bm.x];
  const float tstep = *d.stepf;
  const float beta2t = 1.0f - powf(tstep, -0.8f);            // adafactor.py:170
  const float rel_step = fminf(1e-2f, rsqrtf(tstep));        // adafactor.py:84-94 (relative_step, no warm-up)
  const long long patch = (long long)bm.y + threadIdx.x;
  const long long npatch = PACKED ? (long long)packs[bm.x].O * packs[bm.x].I : d.n / KK;
  const float coef = clip_coef(total_sumsq, max_norm);
  float su = 0.f, sp = 0.f;
  if (patch < npatch) {
    float g[KK], p[KK];
    long long gbase = 0;  // PACKED: offset of (o, i) in the packed buffer
    int sh = 0, s = 1, Ipad = 0;
    if (PACKED) {
      const sb_pack_desc pk = packs[bm.x];
      const int o = (int)(patch / pk.I), i = (int)(patch - (long long)o * pk.I);
      const int oo = pk.operm ? pk.operm[o] : o;
      const int ii = pk.iperm ? pk.iperm[i] : i;
      s = pk.s; Ipad = pk.Ipad;
      while ((1 << sh) < s) ++sh;
      gbase = (long long)oo * ((KH >> sh) * (KW >> sh) * s * s * Ipad) + ii;
#pragma unroll
      for (int k = 0; k < KK; ++k) g[k] = pk.gpacked[gbase + packed_tap_offset<KH, KW>(k, sh, s, Ipad)] * coef;
    } else if (KK % 4 == 0) {
#pragma unroll
      for (int k = 0; k < KK; k += 4) {
        const float4 v = *reinterpret_cast<const float4*>(d.g + patch * KK + k);
        g[k] = v.x * coef; g[k + 1] = v.y * coef; g[k + 2] = v.z * coef; g[k + 3] = v.w * coef;
      }
    } else {
#pragma unroll
      for (int k = 0; k < KK; ++k) g[k] = d.g[patch * KK + k] * coef;
    }
    // the 16 (64) weights of a patch are contiguous: 16-byte accesses (4x fewer L1 wavefronts than scalar ones)
    if (KK % 4 == 0) {
#pragma unroll
      for (int k = 0; k < KK; k += 4) {
        const float4 v = *reinterpret_cast<const float4*>(d.p + patch * KK + k);
        p[k] = v.x; p[k + 1] = v.y; p[k + 2] = v.z; p[k + 3] = v.w;
      }
    } else {
#pragma unroll
      for (int k = 0; k < KK; ++k) p[k] = d.p[patch * KK + k];
    }
    float row[KH], col[KW];
    if (PASS == 1) {
#pragma unroll
      for (int a = 0; a < KH; ++a) {
        float m = 0.f;
#pragma unroll
        for (int b = 0; b < KW; ++b) m += g[a * KW + b] * g[a * KW + b] + 1e-30f;
        row[a] = d.row[patch * KH + a] * beta2t + (m / KW) * (1.0f - beta2t);
        d.row[patch * KH + a] = row[a];
      }
#pragma unroll
      for (int b = 0; b < KW; ++b) {
        float m = 0.f;
#pragma unroll
        for (int a = 0; a < KH; ++a) m += g[a * KW + b] * g[a * KW + b] + 1e-30f;
        col[b] = d.col[patch * KW + b] * beta2t + (m / KH) * (1.0f - beta2t);
        d.col[patch * KW + b] = col[b];
      }
    } else {
#pragma unroll
      for (int a = 0; a < KH; ++a) row[a] = d.row[patch * KH + a];
#pragma unroll
      for (int b = 0; b < KW; ++b) col[b] = d.col[patch * KW + b];
    }
    float rmean = 0.f;
#pragma unroll
    for (int a = 0; a < KH; ++a) rmean += row[a];
    rmean /= KH;
    float rf[KH], cf[KW];
#pragma unroll
    for (int a = 0; a < KH; ++a) rf[a] = rsqrtf(row[a] / rmean);
#pragma unroll
    for (int b = 0; b < KW; ++b) cf[b] = rsqrtf(col[b]);
    if (PASS == 1) {
#pragma unroll
      for (int k = 0; k < KK; ++k) {
        const float u = rf[k / KW] * cf[k % KW] * g[k];
        su += u * u;
        sp += p[k] * p[k];
      }
    } else {
      const float n = (float)d.n;
      const float rms_u = __fsqrt_rn(d.acc[0] / n);
      const float rms_p = __fsqrt_rn(d.acc[1] / n);
      const float lr = fmaxf(1e-3f, rms_p) * rel_step;
      const float scale = lr / fmaxf(1.0f, rms_u);  // clip_threshold = 1
#pragma unroll
      for (int k = 0; k < KK; ++k) p[k] -= scale * (rf[k / KW] * cf[k % KW] * g[k]);
      if (KK % 4 == 0) {
#pragma unroll
        for (int k = 0; k < KK; k += 4)
          *reinterpret_cast<float4*>(d.p + patch * KK + k) = make_float4(p[k], p[k + 1], p[k + 2], p[k + 3]);
      } else {
#pragma unroll
        for (int k = 0; k < KK; ++k) d.p[patch * KK + k] = p[k];
      }
      if (PACKED) {
        __nv_bfloat16* fw = reinterpret_cast<__nv_bfloat16*>(packs[bm.x].fwd);
#pragma unroll
        for (int k = 0; k < KK; ++k)
          fw[gbase + packed_tap_offset<KH, KW>(k, sh, s, Ipad)] = __float2bfloat16_rn(p[k]);
      }
    }
  }
  if (PASS == 1) {
    // sum u^2 / sum p^2 of the whole tensor, ORDER-DETERMINISTIC (see grad_sumsq_kernel): scratch = [ntensors ticket
    // counters | 2 floats per CTA of this launch]; the CTAs of one tensor are consecutive in the block map, so the
    // tensor's first CTA is blockIdx.x - (first patch / 128) and the last arriver adds the partials in CTA order.
    su = block_sum(su, red);
    sp = block_sum(sp, red);
    unsigned* tickets = reinterpret_cast<unsigned*>(scratch);
    float* part = scratch + ntensors;
    const int nb = (int)((npatch + 127) / 128);
    const int first = (int)blockIdx.x - bm.y / 128;
    if (threadIdx.x == 0) {
      part[2 * blockIdx.x] = su;
      part[2 * blockIdx.x + 1] = sp;
      __threadfence();
      s_ticket = atomicAdd(&tickets[bm.x], 1u);
    }
    __syncthreads();
    if (s_ticket == (unsigned)(nb - 1) && threadIdx.x < 32) {
      __threadfence();
      float tu = 0.f, tp = 0.f;
      const volatile float* pv = part + 2 * first;
      for (int i = (int)threadIdx.x; i < nb; i += 32) {
        tu += pv[2 * i];
        tp += pv[2 * i + 1];
      }
      tu = warp_sum(tu);
      tp = warp_sum(tp);
      if (threadIdx.x == 0) {
        d.acc[0] = tu;
        d.acc[1] = tp;
        tickets[bm.x] = 0u;
      }
    }
  }
}

// ---- 3. small 1-D / 2-D tensors: one CTA per tensor ---------------------------------------------------------------
__global__ void __launch_bounds__(1024)
    adafactor_small_kernel(const sb_tensor_desc* __restrict__ descs, const int* __restrict__ index,
                           const float* __restrict__ total_sumsq, float max_norm) {
  extern __shared__ float sm[];  // rows [R] | cols [C]
  __shared__ float red[32];
  const sb_tensor_desc d = descs[index ? index[blockIdx.x] : blockIdx.x];
  const float coef = clip_coef(total_sumsq, max_norm);
  const long long n = d.n;
  const float tstep = *d.stepf;
  const float beta2t = 1.0f - powf(tstep, -0.8f);
  const float rel_step = fminf(1e-2f, rsqrtf(tstep));
  float su = 0.f, sp = 0.f;
  if (d.cols == 0) {  // 1-D: exp_avg_sq in d.row
    for (long long i = threadIdx.x; i < n; i += blockDim.x) {
      const float g = d.g[i] * coef;
      const float v = d.row[i] * beta2t + (g * g + 1e-30f) * (1.0f - beta2t);
      d.row[i] = v;
      const float u = g * rsqrtf(v);
      su += u * u;
      const float p = d.p[i];
      sp += p * p;
    }
    su = block_sum(su, red);
    sp = block_sum(sp, red);
    const float lr = fmaxf(1e-3f, __fsqrt_rn(sp / (float)n)) * rel_step;
    const float scale = lr / fmaxf(1.0f, __fsqrt_rn(su / (float)n));
    for (long long i = threadIdx.x; i < n; i += blockDim.x) {
      const float g = d.g[i] * coef;
      d.p[i] -= scale * g * rsqrtf(d.row[i]);
    }
    if (threadIdx.x == 0) d.acc[1] = sp;
    return;
  }
  const int R = d.rows, C = d.cols;
  float* srow = sm;
  float* scol = sm + R;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  // sweep 1 (warp per row, lanes stride the columns: coalesced, no atomics): row sums of g^2+eps and sum p^2
  for (int r = warp; r < R; r += nw) {
    float rs = 0.f;
    const float* gr = d.g + (long long)r * C;
    const float* pr = d.p + (long long)r * C;
    for (int c = lane; c < C; c += 32) {
      const float g = gr[c] * coef;
      rs += g * g + 1e-30f;
      const float p = pr[c];
      sp += p * p;
    }
    rs = warp_sum(rs);
    if (lane == 0) srow[r] = rs;
  }
  // sweep 2 (thread per column, rows looped: coalesced across threads): column sums
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float cs = 0.f;
    for (int r = 0; r < R; ++r) {
      const float g = d.g[(long long)r * C + c] * coef;
      cs += g * g + 1e-30f;
    }
    scol[c] = cs;
  }
  __syncthreads();
  float rpart = 0.f;
  for (int r = threadIdx.x; r < R; r += blockDim.x) {
    const float v = d.row[r] * beta2t + (srow[r] / C) * (1.0f - beta2t);
    d.row[r] = v;
    srow[r] = v;
    rpart += v;
  }
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const float v = d.col[c] * beta2t + (scol[c] / R) * (1.0f - beta2t);
    d.col[c] = v;
    scol[c] = rsqrtf(v);
  }
  const float rmean = block_sum(rpart, red) / R;
  sp = block_sum(sp, red);
  for (int r = threadIdx.x; r < R; r += blockDim.x) srow[r] = rsqrtf(srow[r] / rmean);
  __syncthreads();
  for (int r = warp; r < R; r += nw) {
    const float rf = srow[r];
    const float* gr = d.g + (long long)r * C;
    for (int c = lane; c < C; c += 32) {
      const float u = rf * scol[c] * gr[c] * coef;
      su += u * u;
    }
  }
  su = block_sum(su, red);
  const float lr = fmaxf(1e-3f, __fsqrt_rn(sp / (float)n)) * rel_step;
  const float scale = lr / fmaxf(1.0f, __fsqrt_rn(su / (float)n));
  for (int r = warp; r < R; r += nw) {
    const float rf = srow[r] * scale * coef;
    const float* gr = d.g + (long long)r * C;
    float* pr = d.p + (long long)r * C;
    for (int c = lane; c < C; c += 32) pr[c] -= rf * scol[c] * gr[c];
  }
  if (threadIdx.x == 0) d.acc[1] = sp;
}

// ---- 3b. large 2-D tensors (e.g. bits_predictor.dense1-3: [128, 4608]): one thread-block CLUSTER of 8 CTAs per
// tensor.  CTA k owns a band of rows; column sums, the row-state mean and the two RMS sums are combined through
// distributed shared memory (3 cluster barriers) instead of running 590 k parameters through a single SM.
static_assert(sizeof(sb_tensor_desc) == 64 && sizeof(sb_pack_desc) == 64, "descriptor ABI");
constexpr int kClusterCtas = 8;
namespace cg = cooperative_groups;
__global__ void __cluster_dims__(kClusterCtas, 1, 1) __launch_bounds__(1024)
    adafactor_2d_cluster_kernel(const sb_tensor_desc* __restrict__ descs, const int* __restrict__ index,
                                const float* __restrict__ total_sumsq, float max_norm) {
  extern __shared__ float sm[];  // colpart [C] | cf [C] | rowf [rows_per]
  __shared__ float red[32];
  __shared__ float part[3][kClusterCtas];  // per-CTA partials of: sum(new row state), sum p^2, sum u^2
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const sb_tensor_desc d = descs[index ? index[blockIdx.x / kClusterCtas] : blockIdx.x / kClusterCtas];
  const int R = d.rows, C = d.cols;
  const int rows_per = (R + kClusterCtas - 1) / kClusterCtas;
  const int r0 = min(R, rank * rows_per), r1 = min(R, r0 + rows_per);
  float* colpart = sm;
  float* cf = sm + C;
  float* rowf = sm + 2 * C;
  const float coef = clip_coef(total_sumsq, max_norm);
  const float tstep = *d.stepf;
  const float beta2t = 1.0f - powf(tstep, -0.8f);
  const float rel_step = fminf(1e-2f, rsqrtf(tstep));
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  cluster.sync();  // every CTA of the cluster is running before its shared memory is written remotely
  // sweep 1: row sums of g^2+eps (warp per row) -> new row state; sum p^2
  float sp = 0.f, rpart = 0.f;
  for (int r = r0 + warp; r < r1; r += nw) {
    float rs = 0.f;
    const float* gr = d.g + (long long)r * C;
    const float* pr = d.p + (long long)r * C;
    for (int c = lane; c < C; c += 32) {
      const float g = gr[c] * coef;
      rs += g * g + 1e-30f;
      const float p = pr[c];
      sp += p * p;
    }
    rs = warp_sum(rs);
    if (lane == 0) {
      const float v = d.row[r] * beta2t + (rs / C) * (1.0f - beta2t);
      d.row[r] = v;
      rowf[r - r0] = v;
      rpart += v;
    }
  }
  // sweep 2: partial column sums over this CTA's rows (thread per column, coalesced across threads)
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float cs = 0.f;
    for (int r = r0; r < r1; ++r) {
      const float g = d.g[(long long)r * C + c] * coef;
      cs += g * g + 1e-30f;
    }
    colpart[c] = cs;
  }
  rpart = block_sum(rpart, red);
  sp = block_sum(sp, red);
  if (threadIdx.x == 0) {
    for (int k = 0; k < kClusterCtas; ++k) {
      float(*rp)[kClusterCtas] = cluster.map_shared_rank(part, k);
      rp[0][rank] = rpart;
      rp[1][rank] = sp;
    }
  }
  cluster.sync();
  // column slice of this CTA: total over the 8 partials -> new col state -> rsqrt factor broadcast to every CTA
  {
    const int c0 = (int)((long long)C * rank / kClusterCtas), c1 = (int)((long long)C * (rank + 1) / kClusterCtas);
    for (int c = c0 + threadIdx.x; c < c1; c += blockDim.x) {
      float cs = 0.f;
      for (int k = 0; k < kClusterCtas; ++k) cs += cluster.map_shared_rank(colpart, k)[c];
      const float v = d.col[c] * beta2t + (cs / R) * (1.0f - beta2t);
      d.col[c] = v;
      const float f = rsqrtf(v);
      for (int k = 0; k < kClusterCtas; ++k) cluster.map_shared_rank(cf, k)[c] = f;
    }
  }
  float rsum = 0.f, spt = 0.f;
  for (int k = 0; k < kClusterCtas; ++k) {
    rsum += part[0][k];
    spt += part[1][k];
  }
  const float rmean = rsum / R;
  cluster.sync();
  for (int r = r0 + threadIdx.x; r < r1; r += blockDim.x) rowf[r - r0] = rsqrtf(rowf[r - r0] / rmean);
  __syncthreads();
  float su = 0.f;
  for (int r = r0 + warp; r < r1; r += nw) {
    const float rf = rowf[r - r0];
    const float* gr = d.g + (long long)r * C;
    for (int c = lane; c < C; c += 32) {
      const float u = rf * cf[c] * gr[c] * coef;
      su += u * u;
    }
  }
  su = block_sum(su, red);
  if (threadIdx.x == 0)
    for (int k = 0; k < kClusterCtas; ++k) cluster.map_shared_rank(part, k)[2][rank] = su;
  cluster.sync();
  float sut = 0.f;
  for (int k = 0; k < kClusterCtas; ++k) sut += part[2][k];
  const float n = (float)d.n;
  const float lr = fmaxf(1e-3f, __fsqrt_rn(spt / n)) * rel_step;
  const float scale = lr / fmaxf(1.0f, __fsqrt_rn(sut / n));
  for (int r = r0 + warp; r < r1; r += nw) {
    const float rf = rowf[r - r0] * scale * coef;
    const float* gr = d.g + (long long)r * C;
    float* pr = d.p + (long long)r * C;
    for (int c = lane; c < C; c += 32) pr[c] -= rf * cf[c] * gr[c];
  }
  if (threadIdx.x == 0 && rank == 0) d.acc[1] = spt;
}

// step counters live on the device (one float per tensor) so that a captured CUDA graph advances them by itself
__global__ void adafactor_tick_kernel(const sb_tensor_desc* __restrict__ descs, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) *descs[i].stepf += 1.0f;
}

}  // namespace sb

using namespace sb;

extern "C" int sb_adafactor_tick(const sb_tensor_desc* descs_dev, int ntensors, sb_stream_t stream) {
  if (!descs_dev || ntensors < 1) return SB_ERR_ARG;
  adafactor_tick_kernel<<<(ntensors + 127) / 128, 128, 0, (cudaStream_t)stream>>>(descs_dev, ntensors);
  ++g_launches;
  SB_CHECK_LAUNCH();
  return SB_OK;
}

extern "C" int sb_grad_sumsq(const sb_tensor_desc* descs_dev, const long long* glen_dev, const int* block_map_dev,
                             int nblocks, float* total, float* scratch, float sumsq_scale, sb_stream_t stream) {
  if (!descs_dev || !block_map_dev || nblocks < 1 || !total || !scratch) return SB_ERR_ARG;
  const int grid = nblocks < SB_SUMSQ_MAX_CTAS ? nblocks : SB_SUMSQ_MAX_CTAS;  // 8 CTAs of 256 threads per SM: one resident wave
  grad_sumsq_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(descs_dev, glen_dev, (const int2*)block_map_dev, nblocks,
                                                            total, scratch, sumsq_scale);
  ++g_launches;
  SB_CHECK_LAUNCH();
  return SB_OK;
}

template <int KH, int KW>
static int launch_conv(const sb_tensor_desc* descs, const sb_pack_desc* packs, const int* bm, int nblocks,
                       const float* total, float max_norm, int pass, float* scratch, int ntensors, cudaStream_t s) {
  const int2* m = (const int2*)bm;
  if (packs) {
    if (pass == 1)
      adafactor_conv_kernel<KH, KW, 1, true><<<nblocks, 128, 0, s>>>(descs, packs, m, total, max_norm, scratch, ntensors);
    else
      adafactor_conv_kernel<KH, KW, 2, true><<<nblocks, 128, 0, s>>>(descs, packs, m, total, max_norm, scratch, ntensors);
  } else {
    if (pass == 1)
      adafactor_conv_kernel<KH, KW, 1, false><<<nblocks, 128, 0, s>>>(descs, packs, m, total, max_norm, scratch, ntensors);
    else
      adafactor_conv_kernel<KH, KW, 2, false><<<nblocks, 128, 0, s>>>(descs, packs, m, total, max_norm, scratch, ntensors);
  }
  ++g_launches;
  SB_CHECK_LAUNCH();
  return SB_OK;
}

static int conv_dispatch(const sb_tensor_desc* descs_dev, const sb_pack_desc* packs_dev, const int* block_map_dev,
                         int nblocks, int KH, int KW, const float* total_sumsq, float max_norm, int pass,
                         float* scratch, int ntensors, sb_stream_t stream) {
  if (!descs_dev || !block_map_dev || nblocks < 1 || (pass != 1 && pass != 2)) return SB_ERR_ARG;
  if (pass == 1 && (!scratch || ntensors < 1)) return SB_ERR_ARG;
  cudaStream_t s = (cudaStream_t)stream;
  if (KH == 1 && KW == 1)
    return launch_conv<1, 1>(descs_dev, packs_dev, block_map_dev, nblocks, total_sumsq, max_norm, pass, scratch, ntensors, s);
  if (KH == 3 && KW == 3)
    return launch_conv<3, 3>(descs_dev, packs_dev, block_map_dev, nblocks, total_sumsq, max_norm, pass, scratch, ntensors, s);
  if (KH == 4 && KW == 4)
    return launch_conv<4, 4>(descs_dev, packs_dev, block_map_dev, nblocks, total_sumsq, max_norm, pass, scratch, ntensors, s);
  if (KH == 8 && KW == 8)
    return launch_conv<8, 8>(descs_dev, packs_dev, block_map_dev, nblocks, total_sumsq, max_norm, pass, scratch, ntensors, s);
  return SB_ERR_ARG;
}

extern "C" int sb_adafactor_conv(const sb_tensor_desc* descs_dev, const int* block_map_dev, int nblocks, int KH, int KW,
                                 const float* total_sumsq, float max_norm, int pass, float* scratch, int ntensors,
                                 sb_stream_t stream) {
  return conv_dispatch(descs_dev, nullptr, block_map_dev, nblocks, KH, KW, total_sumsq, max_norm, pass, scratch, ntensors,
                       stream);
}

extern "C" int sb_adafactor_conv_packed(const sb_tensor_desc* descs_dev, const sb_pack_desc* packs_dev,
                                        const int* block_map_dev, int nblocks, int KH, int KW,
                                        const float* total_sumsq, float max_norm, int pass, float* scratch, int ntensors,
                                        sb_stream_t stream) {
  if (!packs_dev) return SB_ERR_ARG;
  return conv_dispatch(descs_dev, packs_dev, block_map_dev, nblocks, KH, KW, total_sumsq, max_norm, pass, scratch,
                       ntensors, stream);
}

extern "C" int sb_adafactor_small(const sb_tensor_desc* descs_dev, const int* index_dev, int ntensors,
                                  int max_rows_plus_cols, const float* total_sumsq, float max_norm,
                                  sb_stream_t stream) {
  if (!descs_dev || ntensors < 1 || max_rows_plus_cols * sizeof(float) > 96 * 1024) return SB_ERR_ARG;
  const size_t smem = (size_t)(max_rows_plus_cols > 0 ? max_rows_plus_cols : 1) * sizeof(float);
  static bool configured = false;
  if (!configured) {
    if (cudaFuncSetAttribute(adafactor_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024) !=
        cudaSuccess)
      return SB_ERR_CUDA;
    configured = true;
  }
  adafactor_small_kernel<<<ntensors, 1024, smem, (cudaStream_t)stream>>>(descs_dev, index_dev, total_sumsq, max_norm);
  ++g_launches;
  SB_CHECK_LAUNCH();
  return SB_OK;
}

extern "C" int sb_adafactor_big2d(const sb_tensor_desc* descs_dev, const int* index_dev, int ntensors, int max_cols,
                                  int max_rows, const float* total_sumsq, float max_norm, sb_stream_t stream) {
  if (!descs_dev || ntensors < 1 || max_cols < 1 || max_rows < 1) return SB_ERR_ARG;
  const size_t smem = ((size_t)2 * max_cols + (max_rows + kClusterCtas - 1) / kClusterCtas) * sizeof(float);
  if (smem > 200 * 1024) return SB_ERR_ARG;
  static bool configured = false;
  if (!configured) {
    if (cudaFuncSetAttribute(adafactor_2d_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) !=
        cudaSuccess)
      return SB_ERR_CUDA;
    configured = true;
  }
  adafactor_2d_cluster_kernel<<<ntensors * kClusterCtas, 1024, smem, (cudaStream_t)stream>>>(descs_dev, index_dev,
                                                                                            total_sumsq, max_norm);
  ++g_launches;
  SB_CHECK_LAUNCH();
  return SB_OK;
}
