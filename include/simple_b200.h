/* simple_b200 -- C ABI of the B200 (sm_100a) kernels behind the SimPLe world-model hot path.
 *
 * The reference (thomas-schillaci/SimPLe) is pure Python/PyTorch and has NO native plugin / FFI interface
 * (SURVEY.md 8b): the drop-in boundary a user sees is the Python class surface (simple_b200.NextFramePredictor,
 * Adafactor, Trainer, make_simulated_env).  This header is the boundary UNDER that surface: what a reference
 * maintainer would bind with ctypes to replace the ATen call sites listed per entry point.
 *
 * Conventions: plain pointers and sizes, no torch types.  All pointers are DEVICE pointers owned by the caller
 * (PyTorch's allocator); kernels never allocate.  Every call is asynchronous on `stream`, performs no host
 * synchronisation and is CUDA-graph capturable.  Return 0 on success, negative SB_ERR_* otherwise; never throws.
 * Activations are NHWC bf16 ("[B,H,W,C]") unless stated; reference tensors are NCHW fp32.
 */
#ifndef SIMPLE_B200_H
#define SIMPLE_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* sb_stream_t; /* cudaStream_t */

#define SB_MAX_TAPS 16

/* One-time check: device is sm_100, driver exposes cuTensorMapEncodeTiled.  Returns SB_ERR_ARCH otherwise. */
int sb_init(int device);
/* Number of kernels launched by this library since load (bench.py's gpu_launches claim). */
long long sb_launch_count(void);
const char* sb_version(void);

/* ---------------------------------------------------------------------------------------------------------
 * Implicit-GEMM convolution on tcgen05 tensor cores ("tap GEMM").
 *   out[b,y,x,n] (+)= sum_t sum_c  A[b, y+dy[t], x+dx[t], c] * Wp[n, t*C + c]     (zero outside A)
 * A: NHWC bf16 [B,H,W,C] read by 4-D TMA boxes (bw x bh x bb pixels x 64 channels), Wp: bf16 [N, ntaps*C]
 * (K-major) read by 2-D TMA; fp32 accumulation in TMEM; epilogue adds bias, optional ReLU, stores bf16/fp32,
 * or (splits>1 / accumulate) red.add's fp32 partials.  Strided convolutions are expressed as stride-1 taps over a
 * space-to-depth ("s2d") layout of the input, transposed convolutions as taps with negative offsets whose N is
 * (phase, channel); see DESIGN.md.
 * Replaces: nn.Conv2d / nn.ConvTranspose2d forward and their input-gradient (cuDNN) at
 *   simple/next_frame_predictor.py:225,229,245-246,137-139,45,266-268,275 (K1,K2 of SURVEY 2.3).
 * --------------------------------------------------------------------------------------------------------- */
typedef struct {
  const void* a;      /* bf16 [B,H,W,C] */
  int B, H, W, C;     /* C % 8 == 0 */
  int Ho, Wo;         /* output pixel grid (per sample) */
  int bw, bh, bb;     /* tile box in output pixels, bw*bh*bb <= 128 */
  int ntaps;
  int tap_dy[SB_MAX_TAPS];
  int tap_dx[SB_MAX_TAPS];
  const void* w;      /* bf16 [N, ldw], ldw >= ntaps*C, ldw % 8 == 0 */
  int N, ldw;
  void* out;          /* [B,Ho,Wo,ldo] bf16 or fp32 */
  int out_f32;        /* 0: bf16, 1: fp32 */
  int ldo;            /* elements per output pixel */
  const float* bias;  /* [N] or NULL */
  int relu;
  int splits;         /* split-K factor; >1 forces fp32 red.add accumulation into `out` (caller zeroes it) */
  int accumulate;     /* 1: red.add into fp32 `out` even when splits == 1 */
} sb_tap_gemm_args;
int sb_tap_gemm(const sb_tap_gemm_args* args, sb_stream_t stream);

/* ---------------------------------------------------------------------------------------------------------
 * Weight gradient of a tap GEMM on tcgen05 (both operands MN-major, contraction over pixels):
 *   dw[n, t*C + c] += sum_{b,y,x} dy[b,y,x,n] * A[b, y+dy[t], x+dx[t], c]
 * dy: NHWC bf16 [B,Ho,Wo,N]; A as above; dw fp32 [N, ldw] accumulated with red.add (caller zeroes it).
 * Replaces: the weight-gradient halves of ConvolutionBackward0.
 * --------------------------------------------------------------------------------------------------------- */
typedef struct {
  const void* a;
  int B, H, W, C;
  const void* dy;     /* bf16 [B,Ho,Wo,N] */
  int Ho, Wo, N;
  int bw, bh, bb;     /* pixel box, bw*bh*bb == 128 exactly (may overhang; TMA zero-fills) */
  int ntaps;
  int tap_dy[SB_MAX_TAPS];
  int tap_dx[SB_MAX_TAPS];
  float* dw;          /* fp32 [N, ldw] */
  int ldw;
  int splits;         /* pixel-tile split factor */
} sb_tap_wgrad_args;
int sb_tap_wgrad(const sb_tap_wgrad_args* args, sb_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif
