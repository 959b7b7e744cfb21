// Skinny fp32 GEMMs of the latent / heads block: the ~25 dense layers between the bottleneck and the middle network
// (ActionInjector bank, BitsPredictor dense1|2|3, LSTM input projection, dense5, add_bits, reward MLP) and their
// gradients.  Every one of them has a batch-sized dimension (64 rows, or 16 x 64 for the per-symbol ones) and at most
// 7 MB of weights: the library's fp32 kernels run them as 2-24 CTAs (43 us for the [64,4608] x [4608,384] projection,
// 5 TFLOP/s); here every problem is cut into enough (tile, K-split) CTAs to cover the chip, so they run at the speed
// the weights stream out of L2.
//
// Reference arithmetic replaced (file:line under /root/reference): the nn.Linear calls of
//   simple/utils.py:98-105 (ActionInjector), simple/next_frame_predictor.py:43-52,87-121 (BitsPredictor),
//   simple/next_frame_predictor.py:153-159,176 (StochasticModel dense1/2/3), :11-24 (RewardEstimator)
// and the matching autograd matmuls.  Plain fp32 FMAs (the reference runs them in fp32 too).
//
//   C[M,N] (+)= op(A)[M,K] . op(B)[K,N] (+ bias[N])
//     a_trans = 0: A stored [M,K] (row stride lda);  1: A stored [K,M]   (weight gradients: dY^T . X)
//     b_trans = 0: B stored [K,N] (row stride ldb);  1: B stored [N,K]   (nn.Linear forward: X . W^T)
//   accumulate = 0: one CTA per tile stores its result.  accumulate = 1: C already holds zeros (or an addend); the K
//   range is split over `splits` CTAs per tile which add their partial products with fp32 red.add.
#include "common.cuh"
#include "../../include/simple_b200.h"

namespace sb {
extern long long g_launches;

constexpr int kDM = 64, kDN = 64, kDK = 16, kDPad = 4;

struct DenseDev {
  const float* A;
  const float* B;
  float* C;
  const float* bias;
  int M, N, K;
  long long lda, ldb, ldc;
  int a_trans, b_trans, accumulate, kchunk;
};

// one operand tile, global -> registers (4 values per thread); `CONTIG_K`: global memory is contiguous along k (else along
// the row dimension).  Split from the shared-memory store so that the NEXT tile's loads are in flight while the current
// one is multiplied (these kernels run 2-6 K tiles per CTA: un-pipelined, every tile paid a full global-load latency).
template <bool CONTIG_K>
__device__ __forceinline__ float4 fetch_tile(const float* __restrict__ src, long long ld, int row0, int rows, int k0,
                                             int kend, bool vec_ok) {
  const int t = threadIdx.x;
  float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
  if (CONTIG_K) {
    const int r = t >> 2, kq = (t & 3) * 4;  // 64 rows x 4 quads of k
    const int gr = row0 + r, gk = k0 + kq;
    if (gr < rows && gk < kend) {
      const float* p = src + (long long)gr * ld + gk;
      if (vec_ok && gk + 3 < kend) {
        q = __ldg(reinterpret_cast<const float4*>(p));
      } else {
        q.x = __ldg(p);
        if (gk + 1 < kend) q.y = __ldg(p + 1);
        if (gk + 2 < kend) q.z = __ldg(p + 2);
        if (gk + 3 < kend) q.w = __ldg(p + 3);
      }
    }
  } else {
    const int k = t >> 4, rq = (t & 15) * 4;  // 16 k x 16 quads of rows
    const int gk = k0 + k, gr = row0 + rq;
    if (gk < kend && gr < rows) {
      const float* p = src + (long long)gk * ld + gr;
      if (vec_ok && gr + 3 < rows) {
        q = __ldg(reinterpret_cast<const float4*>(p));
      } else {
        q.x = __ldg(p);
        if (gr + 1 < rows) q.y = __ldg(p + 1);
        if (gr + 2 < rows) q.z = __ldg(p + 2);
        if (gr + 3 < rows) q.w = __ldg(p + 3);
      }
    }
  }
  return q;
}

// registers -> shared memory as [k][row]
template <bool CONTIG_K>
__device__ __forceinline__ void stash_tile(float (*dst)[kDM + kDPad], const float4 q) {
  const int t = threadIdx.x;
  if (CONTIG_K) {
    const int r = t >> 2, kq = (t & 3) * 4;
    dst[kq][r] = q.x; dst[kq + 1][r] = q.y; dst[kq + 2][r] = q.z; dst[kq + 3][r] = q.w;
  } else {
    const int k = t >> 4, rq = (t & 15) * 4;
    *reinterpret_cast<float4*>(&dst[k][rq]) = q;
  }
}

template <bool A_T, bool B_T>
__global__ void __launch_bounds__(256) dense_kernel(DenseDev p) {
  __shared__ __align__(16) float As[kDK][kDM + kDPad];
  __shared__ __align__(16) float Bs[kDK][kDN + kDPad];
  const int m0 = blockIdx.y * kDM, n0 = blockIdx.x * kDN;
  const int kbeg = blockIdx.z * p.kchunk;
  const int kend = min(p.K, kbeg + p.kchunk);
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  // 16-byte loads need the base and the row pitch aligned (tile origins are multiples of 4 elements)
  const bool a_vec = ((reinterpret_cast<uintptr_t>(p.A) | (uintptr_t)(p.lda * 4)) & 15) == 0;
  const bool b_vec = ((reinterpret_cast<uintptr_t>(p.B) | (uintptr_t)(p.ldb * 4)) & 15) == 0;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  // A tile as [k][m]: stored [M,K] -> contiguous along k; stored [K,M] -> contiguous along m
  // B tile as [k][n]: stored [N,K] (b_trans) -> contiguous along k; stored [K,N] -> contiguous along n
  float4 qa = fetch_tile<!A_T>(p.A, p.lda, m0, p.M, kbeg, kend, a_vec);
  float4 qb = fetch_tile<B_T>(p.B, p.ldb, n0, p.N, kbeg, kend, b_vec);
  for (int k0 = kbeg; k0 < kend; k0 += kDK) {
    stash_tile<!A_T>(As, qa);
    stash_tile<B_T>(Bs, qb);
    __syncthreads();
    if (k0 + kDK < kend) {  // next tile: in flight underneath the FMAs below
      qa = fetch_tile<!A_T>(p.A, p.lda, m0, p.M, k0 + kDK, kend, a_vec);
      qb = fetch_tile<B_T>(p.B, p.ldb, n0, p.N, k0 + kDK, kend, b_vec);
    }
#pragma unroll
    for (int kk = 0; kk < kDK; ++kk) {
      const float4 a = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w}, bw[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bw[j], acc[i][j]);
    }
    __syncthreads();
  }
  const bool add_bias = p.bias != nullptr && blockIdx.z == 0;
  const int n = n0 + tx * 4;
  const bool c_vec = ((reinterpret_cast<uintptr_t>(p.C) | (uintptr_t)(p.ldc * 4)) & 15) == 0 && n + 3 < p.N;
  float bv[4] = {0.f, 0.f, 0.f, 0.f};
  if (add_bias) {
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (n + j < p.N) bv[j] = __ldg(p.bias + n + j);
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= p.M) continue;
    float* c = p.C + (long long)m * p.ldc + n;
    if (c_vec) {
      if (p.accumulate)
        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(c), "f"(acc[i][0] + bv[0]),
                     "f"(acc[i][1] + bv[1]), "f"(acc[i][2] + bv[2]), "f"(acc[i][3] + bv[3])
                     : "memory");
      else
        *reinterpret_cast<float4*>(c) =
            make_float4(acc[i][0] + bv[0], acc[i][1] + bv[1], acc[i][2] + bv[2], acc[i][3] + bv[3]);
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (n + j >= p.N) continue;
        if (p.accumulate)
          atomicAdd(c + j, acc[i][j] + bv[j]);
        else
          c[j] = acc[i][j] + bv[j];
      }
    }
  }
}

}  // namespace sb

extern "C" int sb_dense_splits(int M, int N, int K) {
  // enough CTAs for two per SM, at least two K tiles per split
  const int tiles = ((M + sb::kDM - 1) / sb::kDM) * ((N + sb::kDN - 1) / sb::kDN);
  if (tiles >= 148) return 1;
  int s = (296 + tiles - 1) / tiles;
  const int kmax = (K + 2 * sb::kDK - 1) / (2 * sb::kDK);
  if (s > kmax) s = kmax;
  return s < 1 ? 1 : s;
}

extern "C" int sb_dense(const float* A, const float* B, float* C, const float* bias, int M, int N, int K, long long lda,
                        long long ldb, long long ldc, int a_trans, int b_trans, int accumulate, sb_stream_t stream) {
  using namespace sb;
  if (!A || !B || !C || M < 1 || N < 1 || K < 1) return SB_ERR_ARG;
  DenseDev p;
  p.A = A; p.B = B; p.C = C; p.bias = bias;
  p.M = M; p.N = N; p.K = K;
  p.lda = lda; p.ldb = ldb; p.ldc = ldc;
  p.a_trans = a_trans; p.b_trans = b_trans; p.accumulate = accumulate;
  const int splits = accumulate ? sb_dense_splits(M, N, K) : 1;
  p.kchunk = (((K + splits - 1) / splits + kDK - 1) / kDK) * kDK;
  const int zs = (K + p.kchunk - 1) / p.kchunk;
  const dim3 grid((N + kDN - 1) / kDN, (M + kDM - 1) / kDM, zs);
  cudaStream_t s = (cudaStream_t)stream;
  if (!a_trans && !b_trans)
    dense_kernel<false, false><<<grid, 256, 0, s>>>(p);
  else if (!a_trans && b_trans)
    dense_kernel<false, true><<<grid, 256, 0, s>>>(p);
  else if (a_trans && !b_trans)
    dense_kernel<true, false><<<grid, 256, 0, s>>>(p);
  else
    dense_kernel<true, true><<<grid, 256, 0, s>>>(p);
  ++g_launches;
  SB_CHECK_LAUNCH();
  return SB_OK;
}
