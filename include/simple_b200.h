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
/* Debug / A-B switch: 1 (default) lets sb_tap_gemm use the CTA-pair (tcgen05 cta_group::2) kernel for large bf16-output
 * layers, 0 forces the single-CTA kernel.  Returns the previous setting. */
int sb_set_pair_mode(int enabled);
/* Same for sb_tap_wgrad's tap groups (3 taps of a 3x3 layer share one dY tile per pipeline stage; used for narrow
 * layers with many pixels: the gate conv).  0 = one tap per CTA everywhere. */
int sb_set_wgrad_tap_groups(int enabled);
/* Same for two 128-row M tiles per sb_tap_wgrad CTA (two dY tiles per pipeline stage against the same A chunks: 1.5x the
 * flops per TMA-loaded row; used when the M side has >= 256 channels).  0 = one M tile per CTA everywhere. */
int sb_set_wgrad_mt2(int enabled);
/* Leave n SMs out of every persistent / one-wave grid launched from now on (room for a concurrent NCCL kernel during
   the data-parallel backward pass); returns the previous value. */
int sb_set_sm_reserve(int n);

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
  int bias_mod;       /* 0: bias has N entries; else bias[column % bias_mod] (a multiple of 32 dividing N): transposed
                         convolutions run with N = (phase, channel) and share one bias per channel */
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
  int bw, bh, bb;     /* pixel box, bw*bh*bb == 64 exactly (may overhang; TMA zero-fills) */
  int ntaps;
  int tap_dy[SB_MAX_TAPS];
  int tap_dx[SB_MAX_TAPS];
  float* dw;          /* fp32 [N, ldw] */
  int ldw;
  int splits;         /* pixel-tile split factor; 0 = automatic (one co-resident wave of CTAs) */
  float* dbias;       /* optional fp32 [N] (caller zeroes): += sum over all pixels of dy[., n] -- the bias gradient, taken by
                         the tensor core from the dy tiles the kernel stages anyway (an extra N = 16 MMA against a tile of
                         ones).  Ignored (left untouched) when the launcher picks the swapped orientation: check
                         sb_tap_wgrad_bias_ok() first. */
} sb_tap_wgrad_args;
/* 1 if sb_tap_wgrad would fill `dbias` for a gradient of N dy-channels x C a-channels (normal tile orientation). */
int sb_tap_wgrad_bias_ok(int N, int C);
int sb_tap_wgrad(const sb_tap_wgrad_args* args, sb_stream_t stream);

/* ---------------------------------------------------------------------------------------------------------
 * Fused per-pixel 256-way softmax cross-entropy, clipped loss, gradient and argmax decode (one HBM pass).
 * logits: bf16 [B*HW, 768], channel = colour*256 + value (colour-major permutation of the reference's
 * value*3+colour, applied when the logits weights are packed).  target: u8 [B,3,HW] (reference NCHW order).
 * dlogits (may alias logits) = (softmax - onehot) * [ce > clip] * gscale; loss_sum += sum max(ce, clip);
 * dbias[768] += column sums of dlogits; argmax_out: u8 [B,3,HW] (first index on ties, like torch.argmax).
 * dlogits == NULL selects the decode-only (rollout) variant.
 * Replaces: simple/trainer.py:130,134-137 and simple/subproc_vec_env.py:428 (K14).
 * --------------------------------------------------------------------------------------------------------- */
/* Debug / A-B switch: 0 (default) = register-staged kernel, 1 = TMA-staged kernel (per-warp shared-memory ring filled by
 * 512-byte cp.async.bulk copies; measured slower: the copies are too small for the TMA unit).  Returns the previous setting. */
int sb_set_ce_tma(int enabled);
int sb_pixel_ce(const void* logits, const uint8_t* target, int B, int HW, float clip, float gscale, void* dlogits,
                uint8_t* argmax_out, double* loss_sum, float* dbias, sb_stream_t stream);
/* The logits layer (1x1 conv 96 -> 3 x 256, simple/next_frame_predictor.py:356-357) FUSED with the loss above: the
 * softmax-CE runs in the epilogue of the tcgen05 CTA-pair GEMM (one N tile = one colour = one softmax group per TMEM
 * lane), so the [B*HW, 768] logits tensor is never written to HBM.  `args` describes the GEMM like sb_tap_gemm (N = 768,
 * one tap, bias required); args->out = dlogits bf16 [B,Ho,Wo,768] (training: the only large tensor that leaves the SM)
 * or NULL (rollout: only argmax_out is produced).  Same results as sb_tap_gemm followed by sb_pixel_ce up to the
 * summation order of the fp32 exp-sums (arg-max identical). */
int sb_logits_ce(const sb_tap_gemm_args* args, const uint8_t* target, float clip, float gscale, uint8_t* argmax_out,
                 double* loss_sum, sb_stream_t stream);

/* ---------------------------------------------------------------------------------------------------------
 * Fused elementwise stage between two tap GEMMs ("prep"), its statistics and its backward.
 * A view is a bf16 tensor holding a logical [B,H,W,C] activation: kind 0 = plain NHWC; kind 1 = padded
 * space-to-depth [B,H2,W2,s*s*C] with element (b,y2,x2,(py*s+px)*C+c) = logical (b, s*y2+py-pad, s*x2+px-pad, c).
 * forward:  z = [relu](in) + add;  x = InstanceNorm(z; sums,gamma,beta) + Ta;  out1 = x;
 *           x = dropout(x; p) + Tb;  out2 = x * sc[b,c] + sh[b,c]
 * Replaces: nn.InstanceNorm2d, F.dropout, timing-signal adds, skip adds, ActionInjector at
 *   simple/next_frame_predictor.py:317,320-326,343-350 and simple/utils.py:98-105 (K3,K4,K7,K8).
 * --------------------------------------------------------------------------------------------------------- */
typedef struct {
  void* p;
  int kind, s, pad, H2, W2;
} sb_view;
typedef struct {
  int B, H, W, C;          /* logical dims; C % 8 == 0 */
  sb_view in;              /* main input */
  const void* add;         /* optional plain bf16 [B,H,W,C] */
  int relu_in;
  int in_f32;              /* main input is fp32 (outputs and gradients stay bf16) */
  int add_f32;             /* `add` is fp32 */
  const float* sums;       /* [B,C,2] from sb_stats, or NULL: no InstanceNorm */
  float eps;
  const float* gamma;
  const float* beta;
  const float* Ta;         /* separable timing table fp32 [(H+W), C/2] (rows y=0..H-1, then x=0..W-1) or NULL */
  void* out1;              /* optional plain bf16 */
  float p;                 /* dropout probability */
  unsigned long long seed; /* Philox key */
  const unsigned long long* seed_ptr; /* optional DEVICE seed added to `seed` (changes between CUDA-graph replays) */
  unsigned int stream_id;  /* Philox stream (one per dropout site) */
  const uint8_t* keep;     /* optional explicit keep mask u8 [B,H,W,C] ("randomness as input" parity mode) */
  const float* Tb;
  const float* sc;         /* fp32 [B,C] or NULL */
  const float* sh;
  sb_view out2;
  /* backward */
  sb_view g2;              /* grad wrt out2 (layout of out2) */
  const void* g1;          /* grad wrt out1 (plain) or NULL */
  float* bsums;            /* [B,C,4]: sum gx, sum gx*xhat, sum g2*xd (d sc), sum g2 (d sh); caller zeroes */
  sb_view d_main;          /* grad wrt in (layout of in), masked by in>0 when relu_in */
  void* d_add;             /* grad wrt add (plain) or NULL */
  float* dcol;             /* optional fp32 [C] (caller zeroes): += column sums of d_main over b,y,x = the bias gradient
                              of the convolution that produced `in` (sb_prep_bwd_apply only) */
  int b_base, b_count;     /* batch-chunked launch: process samples [b_base, b_base+b_count) only (all pointers and B
                              still describe the WHOLE batch; b_count == 0 = all samples).  Lets a caller run
                              stats -> prep or reduce -> apply chunk by chunk inside the L2 */
  float* dgb;              /* optional fp32 [2,C] (caller zeroes; sb_prep_bwd_reduce only): row 0 += sum over b of bsums[b,c,0]
                              (d beta), row 1 += sum over b of bsums[b,c,1] (d gamma) of the InstanceNorm affine */
  int moments;             /* 1 (planes of H*W <= 64 pixels only): `sums` holds (mean, biased centred variance) instead of
                              raw sums -- sb_stats then runs a two-pass, atomics-free kernel (the 6-pixel InstanceNorms with
                              eps = 1e-6 are ill-conditioned under single-pass statistics); same flag for all four calls */
} sb_prep_args;
/* Debug / A-B switch: 1 (default) dispatches to kernel instances specialised at compile time for the feature
 * combinations of the world model; 0 always runs the fully-runtime instances.  Returns the previous setting. */
int sb_set_prep_specialize(int enabled);
int sb_stats(const sb_prep_args* args, float* sums /* [B,C,2], caller zeroes */, sb_stream_t stream);
int sb_prep(const sb_prep_args* args, sb_stream_t stream);
int sb_prep_bwd_reduce(const sb_prep_args* args, sb_stream_t stream);
int sb_prep_bwd_apply(const sb_prep_args* args, sb_stream_t stream);

/* standardize_frame (simple/utils.py:122-126): in fp32 [B,C,P] -> (x-mean)/max(std_unbiased, 1/sqrt(P)) written to
 * dst_bf16[(b*P+i)*ld + coff + c] and/or out_f32 [B,C,P] (may alias `in`: the reference standardises targets in place).
 * stats_scratch: caller-owned fp32 [B*C*2] (per-plane mean / scale between the statistics and the apply kernel). C <= 16.
 * dst2_bf16 (optional): a second NHWC destination of the same values (the posterior's input operand). */
int sb_standardize(const float* in, int B, int C, int P, void* dst_bf16, int ld, int coff, float* out_f32,
                   float* stats_scratch, void* dst2_bf16, int ld2, int coff2, sb_stream_t stream);

/* Scheduled-sampling frame feedback (simple/trainer.py:58-65,125-132), in place on the fp32 frame stack [B,C_stack,HW]:
 * next = gt w.p. *eps_ptr else pred (u8 [B,C_new,HW]); next/255 with each value replaced w.p. p_noise by the frame's lower
 * median (exact 256-bin histogram); stack = cat(stack[:, C_new:], next).  Randomness: counter-based hash of
 * (seed + *seed_ptr). */
int sb_frame_feedback(float* stack, const uint8_t* gt, const uint8_t* pred, const float* eps_ptr, float p_noise,
                      const unsigned long long* seed_ptr, unsigned long long seed, int B, int C_stack, int C_new, int HW,
                      sb_stream_t stream);

/* Epilogue of a split-K sb_tap_gemm (fp32 partial sums [rows, N]): out = act(acc + bias[N]) as fp32 (out may alias acc)
 * or bf16 -- the bias add + ReLU of the reference's conv layers (next_frame_predictor.py:37-63,320-350) for the deep,
 * tiny feature maps whose GEMMs are split over K. */
int sb_bias_act(const float* acc, const float* bias, int bias_mod, int relu, void* out, int out_is_f32, long long rows,
                int N, sb_stream_t stream);

/* conv-GRU blend of the recurrent state (simple/next_frame_predictor.py:298-303) and its backward.
 * G bf16 [npix,128] (gate | candidate pre-activations), state fp32 [npix,64], xs/dxs bf16 with row stride ld.
 * s_new may alias s_old (in-place update).  The backward reads the old state from a bf16 copy with row stride s_ld
 * (channels 0..63 of the previous step's gate operand), so the fp32 state need not be kept. */
int sb_gru_blend(const void* G, const float* s_old, float* s_new, void* xs, int xs_ld, long long npix,
                 sb_stream_t stream);
int sb_gru_blend_bwd(const void* G, const void* s_old_bf16, int s_ld, const void* dxs, int dxs_ld, void* dG,
                     long long npix, sb_stream_t stream);

/* fp32 (O,I,KH,KW) master weight -> bf16 GEMM operands (see DESIGN.md "weight packing"). */
int sb_pack_conv(const float* w, int O, int I, int KH, int KW, int s, int Ipad, int Opad, const int* iperm,
                 const int* operm, void* fwd, void* tr, sb_stream_t stream);
/* tr[:, t*Opad:(t+1)*Opad] = fwd[:, t*PHI:(t+1)*PHI]^T per tap (the second half of sb_pack_conv, for operands whose
 * `fwd` was refreshed in place by sb_adafactor_conv_packed).  PHI = s*s*Ipad. */
int sb_transpose_taps(const void* fwd, void* tr, int Opad, int PHI, int T, sb_stream_t stream);
/* The same for a list of layers in one launch.  first_tile = prefix sum of ceil(PHI/64) * ceil(Opad/64) * T over the
 * preceding entries; total_tiles = the sum over all entries.  descs_dev: device array. */
typedef struct sb_transpose_desc {
  const void* fwd;
  void* tr;
  int Opad, PHI, T, first_tile;
} sb_transpose_desc;
int sb_transpose_taps_multi(const sb_transpose_desc* descs_dev, int n, int total_tiles, sb_stream_t stream);
/* inverse map for gradients: fp32 packed [Opad, T*s*s*Ipad] (layout of `fwd`) -> reference layout (O,I,KH,KW). */
int sb_unpack_conv_grad(const float* packed, int O, int I, int KH, int KW, int s, int Ipad, int Opad,
                        const int* iperm, const int* operm, float* out, sb_stream_t stream);

/* ---------------------------------------------------------------------------------------------------------
 * Multi-tensor fused clip_grad_norm_ + Adafactor (fairseq defaults: eps=(1e-30,1e-3), clip_threshold=1,
 * decay_rate=-0.8, relative_step, scale_parameter, no momentum, no weight decay).
 * Replaces: simple/trainer.py:148 and simple/adafactor.py:113-212 (K16, K17).
 * Descriptors and block maps live in DEVICE memory (built by the host wrapper each step; one small H2D copy).
 *   sb_grad_sumsq:      block_map[i] = (tensor, first 4096-element chunk); *total = sum g^2.  scratch: caller-owned fp32
 *                       [1 + SB_SUMSQ_MAX_CTAS], zeroed ONCE (word 0 is a ticket counter the kernel resets itself).
 *                       sumsq_scale: *total = sumsq_scale * sum g^2 -- 1 normally; 1/world^2 when g holds the SUM of
 *                       the data-parallel shards' gradients (Adafactor itself is invariant to a constant gradient
 *                       scale, only the clip coefficient needs the true norm).
 *   sb_adafactor_conv:  4-D weights (O,I,KH,KW) of one kernel size; block_map[i] = (descriptor index, first patch), 128 patches
 *                       per block, the blocks of one tensor consecutive; pass 1 updates row/col states and writes
 *                       acc[0]=sum u^2, acc[1]=sum p^2, pass 2 applies the update.  g is scaled by
 *                       min(1, max_norm/(sqrt(total)+1e-6)).  scratch (pass 1): caller-owned fp32 [ntensors + 2*nblocks],
 *                       its first ntensors words (ticket counters, reset by the kernel) zeroed ONCE.
 *   All cross-CTA sums above are ORDER-DETERMINISTIC (per-CTA partials, added in index order by the last arriver):
 *   data-parallel replicas that see the same reduced gradient apply bit-identical updates.
 *   sb_adafactor_small: 1-D (cols==0, state in `row`) and 2-D tensors, one CTA per descriptor.
 *   sb_adafactor_big2d: large 2-D tensors, one thread-block cluster of 8 CTAs per descriptor (row bands; column sums
 *                       and RMS partials combined through distributed shared memory).
 * --------------------------------------------------------------------------------------------------------- */
typedef struct {
  float* p;        /* parameter, fp32, reference layout */
  const float* g;  /* gradient */
  float* row;      /* exp_avg_sq_row  (or exp_avg_sq for 1-D) */
  float* col;      /* exp_avg_sq_col */
  float* acc;      /* 2 floats: sum u^2, sum p^2 (RMS state) */
  long long n;     /* elements */
  int rows, cols;  /* 2-D: shape; 1-D: cols = 0; 4-D: unused */
  float* stepf;    /* DEVICE step counter of this tensor (float); beta2 = 1 - step^-0.8, rel_step = min(1e-2, step^-1/2) */
} sb_tensor_desc;
/* Packed-layout companion of a 4-D weight's descriptor (sb_adafactor_conv_packed): the gradient is read from the fp32
 * buffer sb_tap_wgrad accumulated into (layout of sb_pack_conv's `fwd`) and pass 2 also rewrites the bf16 operand. */
typedef struct {
  const float* gpacked; /* fp32 [Opad, T*s*s*Ipad] */
  void* fwd;            /* bf16, same layout */
  const int* iperm;     /* or NULL */
  const int* operm;     /* or NULL */
  int O, I, s, Ipad, Opad;
  int pad_[3];
} sb_pack_desc;         /* 64 bytes */
/* step counters += 1 for every descriptor (first call of a step; keeps CUDA-graph replays self-contained) */
int sb_adafactor_tick(const sb_tensor_desc* descs_dev, int ntensors, sb_stream_t stream);
/* glen_dev: optional per-descriptor gradient length (a packed gradient buffer is longer than its parameter) */
#define SB_SUMSQ_MAX_CTAS (148 * 8)
int sb_grad_sumsq(const sb_tensor_desc* descs_dev, const long long* glen_dev, const int* block_map_dev, int nblocks,
                  float* total, float* scratch, float sumsq_scale, sb_stream_t stream);
int sb_adafactor_conv(const sb_tensor_desc* descs_dev, const int* block_map_dev, int nblocks, int KH, int KW,
                      const float* total_sumsq, float max_norm, int pass, float* scratch, int ntensors,
                      sb_stream_t stream);
int sb_adafactor_conv_packed(const sb_tensor_desc* descs_dev, const sb_pack_desc* packs_dev, const int* block_map_dev,
                             int nblocks, int KH, int KW, const float* total_sumsq, float max_norm, int pass,
                             float* scratch, int ntensors, sb_stream_t stream);
/* index_dev: optional list of descriptor indices handled by this launch (NULL = descriptors 0..ntensors-1) */
int sb_adafactor_small(const sb_tensor_desc* descs_dev, const int* index_dev, int ntensors, int max_rows_plus_cols,
                       const float* total_sumsq, float max_norm, sb_stream_t stream);
int sb_adafactor_big2d(const sb_tensor_desc* descs_dev, const int* index_dev, int ntensors, int max_cols, int max_rows,
                       const float* total_sumsq, float max_norm, sb_stream_t stream);

/* ---------------------------------------------------------------------------------------------------------
 * MeanAttention of the posterior encoder (simple/utils.py:72-88), one CTA per sample, activations staged in shared memory.
 * x: bf16 NHWC conv output [B, HW, C]; We [4,C], be [4]: the 1x1 conv; Wd [O, 5C], bd [O]: the dense layer on
 * cat(am, m) laid out c*5+k.  The softmax follows the reference's memory reinterpretation (groups = flat index mod 4).
 * fwd saves s [B, 4*HW] (softmax, flat k*HW+hw) and l [B, 5C]; bwd returns dx (bf16, layout of x) and the PER-SAMPLE
 * parts dWe_part [B,4,C], dbe_part [B,4] (summed over B by the caller; dWd = dy^T l and dbd = sum dy are plain GEMMs).
 * --------------------------------------------------------------------------------------------------------- */
int sb_mean_attention_fwd(const void* x, const float* We, const float* be, const float* Wd, const float* bd, int B,
                          int HW, int C, int O, float* y, float* s_save, float* l_save, sb_stream_t stream);
/* accumulate = 1: dWe_part [4,C] / dbe_part [4] receive the batch SUM through atomics (caller zeroes) instead of per-sample parts. */
int sb_mean_attention_bwd(const void* x, const float* We, const float* Wd, const float* s_save, const float* dy, int B,
                          int HW, int C, int O, void* dx, float* dWe_part, float* dbe_part, int accumulate,
                          sb_stream_t stream);

/* Latent-bit predictor (simple/next_frame_predictor.py:66-121; K11/K12): transposed weights (whhT [128,512] = W_hh^T, w5T [128,256] = W5^T), input-gate
 * pre-activations by table lookup (E [256,512] = W_ih (W4[:,s] + b4) + b_ih + b_hh; G0 [B,512] for the first step), one
 * or two samples per CTA.  sb_lstm_teacher_fwd/bwd run all 16 teacher-forced LSTMCell steps and their backward through
 * time in one launch each: xw [B,16,512] input-gate pre-activations -> hs [B,16,128]; saved act [B,16,512], cs [B,16,128];
 * backward emits dpre [B,16,512] (= d xw; dW_hh = dpre^T h_prev on the host side), dh0, dc0 [B,128]. */
/* hc_ld: row stride (floats, >= 128) of h0 / c0 / dh0 / dc0 -- they may be column slices of the [B,384] output of the
 * fused dense1|dense2|dense3 projection. */
int sb_lstm_sample2(const float* G0, const float* h0, const float* c0, const float* E, const float* whhT,
                    const float* w5T, const float* b5, const float* noise, const unsigned long long* seed_ptr, int B,
                    int hc_ld, float* bits, int* ints, sb_stream_t stream);
int sb_lstm_teacher_fwd(const float* xw, const float* h0, const float* c0, const float* whhT, int B, int hc_ld,
                        float* hs, float* act, float* cs, sb_stream_t stream);
int sb_lstm_teacher_bwd(const float* dhs, const float* act, const float* cs, const float* c0, const float* whh, int B,
                        int hc_ld, float* dpre, float* dh0, float* dc0, sb_stream_t stream);

/* ---------------------------------------------------------------------------------------------------------
 * Latent / heads block (csrc/latent.cu): every elementwise, reduction, gather and scatter piece between the bottleneck
 * and the middle network, forward and backward; the dense layers between them are sb_dense launches issued by the
 * caller.  "flat" = the reference's flatten order of the [768,3,2] bottleneck (index c*6 + h*2 + w).  Optional explicit
 * randomness pointers (NULL = drawn in-kernel from a counter hash of seed + *seed_ptr; the backward regenerates it).
 * Replaces: simple/next_frame_predictor.py:328 (value head), simple/utils.py:98-105 (ActionInjector scale/shift, all
 * injectors at once), next_frame_predictor.py:176-197 (posterior bits, straight-through, the three mix() calls of
 * simple/utils.py:157-163), :87-107 (teacher forcing: bytes, dense4 lookup, dropouts, CE), :153-159 (add_bits),
 * :11-24,352-354 + simple/trainer.py:138-141 (reward MLP tail, reward CE, value MSE).
 * --------------------------------------------------------------------------------------------------------- */
typedef struct {
  const float* dsc[9];   /* grad w.r.t. sigmoid(dense1(a)) of injector i (NULL = none); element (b,c) at b*sc_sb + c*sc_sc */
  const float* dsh[9];   /* grad w.r.t. dense2(a); element (b,c) at b*sh_sb + c*sh_sc */
  int sc_sb[9], sc_sc[9], sh_sb[9], sh_sc[9];
} sb_inj_grads;
/* x5: bf16 NHWC [B,3,2,768]; bank [B, 2S] = (dense1 pre-activations | dense2 outputs) of the `ninj` injectors (widths[i]
 * channels each, widths[0] == 768, S = sum); wv [4608] flat, bv [1].  Outputs: sc_flat / sh_flat (injector i: contiguous
 * [B, widths[i]] block at element offset B * sum(widths[:i])), h_flat [B,4608] = x5 * sc_0 + sh_0 (may be NULL), value [B]. */
int sb_heads_fwd(const void* x5, const float* bank, const int* widths, int ninj, const float* wv, const float* bv, int B,
                 float* sc_flat, float* sh_flat, float* h_flat, float* value, sb_stream_t stream);
/* dx5 bf16 NHWC; dbank [B,2S]; dwv [4608], dbv [1], dbank_b [2S]: accumulated (+=), caller zeroes. */
int sb_heads_bwd(const void* x5, const float* bank, const int* widths, int ninj, const sb_inj_grads* grads,
                 const float* wv, const float* dvalue, const float* dh_flat, int B, void* dx5, float* dbank, float* dwv,
                 float* dbv, float* dbank_b, sb_stream_t stream);
/* pre [B,128] -> x = tanh(pre + truncnorm), bits2 = mix(sign(x) * flip, x, 1 - eps), bits_clean = sign(pre),
 * flags u8 [B,128] (bit0 flip = -1, bit1 first mix took x), tints int32 [B,16] (bytes of bits_clean, LSB first). */
int sb_posterior_bits_fwd(const float* pre, const float* tn, const float* u_flip, const float* u_mix1,
                          const float* eps_ptr, float eps_val, float noise_p, const unsigned long long* seed_ptr,
                          unsigned long long seed, int B, float* x, float* bits2, float* bits_clean, uint8_t* flags,
                          int* tints, sb_stream_t stream);
/* bits3 = mix(bits_pred, bits2, 1 - (1 - eps) * max_sampling); sets flags bit2 where bits2 was taken. */
int sb_mix_bits(const float* bits_pred, const float* bits2, const float* u_mix2, const float* eps_ptr, float eps_val,
                float max_sampling, const unsigned long long* seed_ptr, unsigned long long seed, int B, float* bits3,
                uint8_t* flags, sb_stream_t stream);
/* dpre = dbits3 * m2 * ((1-m1)*flip + m1) * (1 - x^2); db1 [128] += column sums (caller zeroes). */
int sb_posterior_bits_bwd(const float* dbits3, const float* x, const uint8_t* flags, int B, float* dpre, float* db1,
                          sb_stream_t stream);
/* tin [B,16,128]: step 0 = proj[b, :128] (row stride proj_ld), step t = dropout(w4[:, byte_{t-1}] + b4; keep [B,16,128] or
 * hash, p); optionally bsum [512] = b_ih + b_hh.  Backward: dproj[b, :128] = dtin[:, 0]; dw4 [128,256], db4 [128] += scatter. */
int sb_teacher_embed_fwd(const float* proj, int proj_ld, const int* tints, const float* w4, const float* b4,
                         const float* keep, float p, const unsigned long long* seed_ptr, unsigned long long seed,
                         const float* b_ih, const float* b_hh, float* bsum, int B, float* tin, sb_stream_t stream);
int sb_teacher_embed_bwd(const float* dtin, const int* tints, const float* keep, float p,
                         const unsigned long long* seed_ptr, unsigned long long seed, int B, float* dproj, int dproj_ld,
                         float* dw4, float* db4, sb_stream_t stream);
/* out = x * keep / (1 - p) over n elements (keep: explicit 0/1 floats, or NULL = hash stream `stream_id`). */
int sb_drop_scale(const float* x, const float* keep, float p, const unsigned long long* seed_ptr, unsigned long long seed,
                  unsigned int stream_id, float* out, long long n, sb_stream_t stream);
/* Row-wise softmax CE (N <= 256): *loss_sum += sum_r (lse - x[r,t_r]); dlogits = (softmax - onehot) * gscale;
 * dbias [N] += column sums of dlogits (optional).  Caller zeroes loss_sum / dbias. */
int sb_softmax_ce_rows(const float* logits, const int* targets, int R, int N, float gscale, float* loss_sum,
                       float* dlogits, float* dbias, sb_stream_t stream);
/* zz [B,1536] = (dense2(bits) | dense3(bits)); out = mix(h*sigmoid(z_mul)+z_add, h, 1-(1-eps)*use_max_p) when train, else
 * h*sigmoid(z_mul)+z_add.  hm: bf16 NHWC [B,3,2,768]; x_mid [B,768] = mean over the 6 pixels; m3 u8 [B,4608] (train). */
int sb_latent_out_fwd(const float* h_flat, const float* zz, const float* u_mix3, int train, const float* eps_ptr,
                      float eps_val, float use_max_p, const unsigned long long* seed_ptr, unsigned long long seed, int B,
                      void* hm, float* x_mid, uint8_t* m3, sb_stream_t stream);
/* dx_mid: [B,768] with row stride dxm_ld (a column slice of the reward MLP's input gradient), or NULL. */
int sb_latent_out_bwd(const void* dhm, const float* dx_mid, int dxm_ld, const float* h_flat, const float* zz,
                      const uint8_t* m3, int B, float* dh_flat, float* dzz, float* db23, sb_stream_t stream);
/* whhT [128,512], wihT [128,512], w5T [128,256], emb [256,128] = w4^T + b4 (operands of sb_lstm_sample2) in one launch. */
int sb_sampler_tables(const float* w_hh, const float* w_ih, const float* w5, const float* w4, const float* b4, float* whhT,
                      float* wihT, float* w5T, float* emb, sb_stream_t stream);
/* hid_pre [B,128]; reward_pred [B,3] = w2 relu(hid_pre) + b2.  With rewards (int64 [B]): loss_out[0] += CE / B,
 * loss_out[1] += mean (value_pred - values)^2, dhid_pre, dw2 [3,128] / db2 [3] (+=), dvalue [B].  rewards == NULL: forward only. */
int sb_reward_value(const float* hid_pre, const float* w2, const float* b2, const long long* rewards,
                    const float* value_pred, const float* values, int B, float* reward_pred, float* loss_out,
                    float* dhid_pre, float* dw2, float* db2, float* dvalue, sb_stream_t stream);
/* out[N] += column sums of x [M,N] (row stride ld): bias gradients of the dense layers. */
int sb_colsum_f32(const float* x, int M, int N, int ld, float* out, sb_stream_t stream);
/* Simulated-env step tail (simple/subproc_vec_env.py:428-441), one launch: frames u8 [A, stack_channels, HW] shifted by c
 * channels in place with `fresh` u8 [A, c, HW] appended; obs f32 (same shape, values 0..255) or NULL;
 * rew_out[a] = argmax_r reward_pred[a, r] - 1 (first maximum), val_out = value_pred (either pair may be NULL).
 * HW must be a multiple of 16, the image pointers 16-byte aligned. */
int sb_env_post(uint8_t* frames, const uint8_t* fresh, float* obs, int A, int stack_channels, int c, int HW,
                const float* reward_pred, int R, const float* value_pred, float* rew_out, float* val_out,
                sb_stream_t stream);

/* ---- step bookkeeping (latent.cu) ------------------------------------------------------------------------------
 * sb_loss_total: out5 = (total, reconstruct, value, reward, lstm) with reconstruct = rec_sum / n - clip (the logits
 *   kernel sums max(ce, clip); simple/trainer.py:134-144), total = ((rec + value) + reward) + lstm; running5 (or NULL)
 *   += out5 (the per-window loss log, trainer.py:150-154).
 * sb_gather_step: this step's batch out of the device-resident replay (trainer.py:108-121): record idx[b] of
 *   obs_src u8 [nrec, frame_bytes], act_src f32 [nrec, n_action], rew_src i64 [nrec], val_src f32 [nrec]. */
int sb_loss_total(const double* rec_sum, double n, float clip, const float* l_value, const float* l_reward,
                  const float* l_lstm, float* out5, float* running5, sb_stream_t stream);
int sb_gather_step(const long long* idx, int B, long long nrec, const uint8_t* obs_src, long long frame_bytes,
                   const float* act_src, int n_action, const long long* rew_src, const float* val_src, uint8_t* obs,
                   float* act, long long* rew, float* val, sb_stream_t stream);

/* ---- skinny fp32 GEMMs of the latent / heads block (dense.cu) ---------------------------------------------------
 * C[M,N] (+)= op(A)[M,K] . op(B)[K,N] (+ bias[N]).  a_trans: A stored [K,M]; b_trans: B stored [N,K] (an nn.Linear
 * weight); lda / ldb / ldc: row pitches in elements (inner strides are 1).  accumulate = 0: stores C.  accumulate = 1:
 * C must hold zeros or an addend; the K range is split over sb_dense_splits(M,N,K) CTAs per 64x64 tile, partial
 * products are added with fp32 red.add.
 * Replaces the nn.Linear forwards and autograd matmuls of simple/utils.py:98-105 (ActionInjector),
 * simple/next_frame_predictor.py:11-24 (RewardEstimator), :43-52,:87-121 (BitsPredictor), :153-159,:176 (StochasticModel). */
int sb_dense_splits(int M, int N, int K);
int sb_dense(const float* A, const float* B, float* C, const float* bias, int M, int N, int K, long long lda,
             long long ldb, long long ldc, int a_trans, int b_trans, int accumulate, sb_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif
