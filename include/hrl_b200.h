/*
 * hrl_b200.h -- C ABI of the B200-native HandyRL learner hot path.
 *
 * The reference (DeNA/HandyRL) is pure Python and has no FFI layer; the seam this
 * header introduces is the operator level underneath these reference functions:
 *
 *   hrl_loss_fwd_bwd      <- handyrl/train.py:176-184 (mask epilogue of forward_prediction)
 *                            handyrl/train.py:218-267 (compute_loss)
 *                            handyrl/train.py:189-215 (compose_losses)
 *                            handyrl/losses.py:16-80  (monte_carlo / temporal_difference / upgo / vtrace)
 *                            + the autograd pass of train.py:369 restricted to those ops
 *   hrl_compute_target    <- handyrl/losses.py:63-80  (compute_target, stand-alone)
 *   hrl_peer_allreduce_sumsq <- the gradient exchange nn.DataParallel does implicitly (train.py:339-340), as a
 *                            fused peer-memory kernel
 *   hrl_grad_sumsq /
 *   hrl_clip_adam_step    <- handyrl/train.py:370-371 (clip_grad_norm_(params, 4.0) + Adam.step,
 *                            Adam(lr, weight_decay=1e-5) of train.py:331)
 *   hrl_gather_pad        <- handyrl/train.py:33-124  (make_batch: window slice + pad + collate)
 *   hrl_gemm_tf32x3       <- the Linear/Conv contractions of the user's net inside train.py:142-146 (+ autograd, :369)
 *
 * Conventions
 *   - plain C, no torch types; every pointer is a DEVICE pointer unless stated;
 *   - the caller owns every buffer (inputs, outputs, workspace); the library allocates nothing;
 *   - kernels are enqueued on the stream handed in (a cudaStream_t passed as void*), no
 *     internal synchronisation, safe to capture in a CUDA graph;
 *   - return value 0 = success, negative = HrlStatus error; hrl_last_error() gives the text
 *     for the calling thread.  Nothing throws across this boundary.
 *   - tensors are contiguous, batch-major exactly as make_batch emits them
 *     (B, T, P|Pa, ...) fp32, `action` int64 (train.py:44-45, 111-124).
 */
#ifndef HRL_B200_H_
#define HRL_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HRL_ABI_VERSION 2

typedef enum {
    HRL_OK = 0,
    HRL_ERR_BAD_ARG = -1,      /* null pointer, bad dimension, unknown algorithm id        */
    HRL_ERR_WORKSPACE = -2,    /* workspace missing or smaller than hrl_*_workspace_bytes   */
    HRL_ERR_UNSUPPORTED = -3,  /* shape outside what the kernels were built for             */
    HRL_ERR_CUDA = -4          /* a CUDA runtime call failed; see hrl_last_error()          */
} HrlStatus;

/* target algorithms, the strings of config.yaml `policy_target` / `value_target`
 * (reference config.yaml:26-27, losses.py:68-78) */
typedef enum { HRL_MC = 0, HRL_TD = 1, HRL_UPGO = 2, HRL_VTRACE = 3 } HrlAlgo;

/* index of each reduced scalar in HrlLossArgs.losses */
enum { HRL_LOSS_P = 0, HRL_LOSS_V = 1, HRL_LOSS_R = 2, HRL_LOSS_ENT = 3, HRL_LOSS_TOTAL = 4,
       HRL_LOSS_DCNT = 5, HRL_NUM_LOSS = 6 };

/*
 * One fused forward+backward pass of the loss over a replay batch.
 *
 * Dimensions: B windows, T = burn_in + forward steps, P players on the value side,
 * Pa players on the policy side (1 for the turn-alternating layout, else P; train.py:65-68),
 * A actions (A <= 1024 and P <= 64 are built; larger values return HRL_ERR_UNSUPPORTED).  Steps t < burn_in are
 * excluded from every loss term and receive zero gradients (train.py:220-222).
 */
/* Kernel selection and tuning of hrl_loss_fwd_bwd.  All zero = the library's own choice (what production uses);
 * the fields exist so that tests and profiling can force every code path WITHOUT process-global state (the library
 * reads no environment variables). */
typedef struct HrlLossTuning {
    int32_t variant;     /* 0 auto | 1 rows, direct loads | 2 rows, cp.async-staged | 3 bulk (TMA, wide rows) |
                            4 element-parallel | 5 lane-group (A <= 32); inapplicable variants fall back to rows */
    int32_t recurrence;  /* 0 auto (serial below 96 steps) | 1 serial loops | 2 parallel suffix scan           */
    int32_t cluster;     /* bulk kernel: CTAs per window, 0 auto | 1 | 2 | 4 | 8                               */
    int32_t consumers;   /* bulk kernel: row-reducing warps per CTA, 0 = 16                                    */
    int32_t threads;     /* threads per CTA for the other variants, 0 auto                                     */
    int32_t unstaged;    /* rows kernel: 1 = recompute masked logits instead of staging them in shared memory  */
    long long *trace;    /* device buffer of >= 32 clock64 stamps written by CTA 0 (debugging), or NULL        */
} HrlLossTuning;

typedef struct HrlLossArgs {
    int32_t B, T, P, Pa, A;
    int32_t burn_in;
    int32_t value_target;        /* HrlAlgo, used for targets (train.py:257-258)            */
    int32_t policy_target;       /* HrlAlgo, used for advantages (train.py:260-262)         */
    int32_t two_player_zero_sum; /* turn_based_training && P == 2 (train.py:243)            */
    float lambda;                /* args['lambda']                                           */
    float gamma;                 /* args['gamma'] (return stream; the value stream uses 1)   */
    float entropy_regularization;
    float entropy_regularization_decay;

    /* raw net outputs (before the mask epilogue) */
    const float *policy_raw;     /* (B,T,Pa,A)                                               */
    const float *value_raw;      /* (B,T,Pa)   or NULL when the net has no value head        */
    const float *return_raw;     /* (B,T,Pa)   or NULL when the net has no return head       */

    /* replay batch */
    const float *action_mask;    /* (B,T,Pa,A) 0 legal / 1e32 illegal                        */
    const int64_t *action;       /* (B,T,Pa)                                                 */
    const float *selected_prob;  /* (B,T,Pa)   behaviour probability                         */
    const float *reward;         /* (B,T,P)                                                  */
    const float *ret;            /* (B,T,P)    batch['return']                               */
    const float *turn_mask;      /* (B,T,P)                                                  */
    const float *observation_mask; /* (B,T,P)                                                */
    const float *episode_mask;   /* (B,T)                                                    */
    const float *progress;       /* (B,T)                                                    */
    const float *outcome;        /* (B,P)                                                    */

    /* outputs */
    float *dpolicy_raw;          /* (B,T,Pa,A) d total / d policy_raw                        */
    float *dvalue_raw;           /* (B,T,Pa)   or NULL iff value_raw is NULL                 */
    float *dreturn_raw;          /* (B,T,Pa)   or NULL iff return_raw is NULL                */
    float *losses;               /* [HRL_NUM_LOSS] p, v, r, ent, total, dcnt (sums)          */

    /* optional per-element taps for parity tests; each may be NULL */
    float *tap_target_value;     /* (B,T,P) targets['value']  (t >= burn_in, 0 before)       */
    float *tap_target_return;    /* (B,T,P) targets['return']                                */
    float *tap_advantage;        /* (B,T,P) total_advantages broadcast to P (train.py:265)   */
    float *tap_logp;             /* (B,T,Pa) log pi(a) * episode_mask (train.py:232)         */
    float *tap_rho;              /* (B,T,Pa) clipped importance ratio (train.py:237)         */
    float *tap_entropy;          /* (B,T,Pa) policy entropy per row (train.py:208)           */

    void *workspace;             /* >= hrl_loss_workspace_bytes(...) bytes, 256-byte aligned;
                                    its first 64 bytes must be zero on the first call and are
                                    left zero by every call                                  */
    size_t workspace_bytes;
    HrlLossTuning tuning;        /* zero-initialise for the defaults                         */
    int32_t io_bf16;             /* 1: policy_raw and dpolicy_raw hold bf16 (same shapes): 8 instead of 12 bytes per action
                                    move through HBM.  Wide rows only (256 < A <= 512, A % 8 == 0); everything in between --
                                    masks, softmax statistics, targets, the gradient before its final rounding -- stays fp32, so
                                    the losses equal those of the fp32 call on the widened logits bit for bit              */
} HrlLossArgs;

/* Bytes of workspace hrl_loss_fwd_bwd needs for these dimensions (host call, no GPU work). */
size_t hrl_loss_workspace_bytes(int32_t B, int32_t T, int32_t P, int32_t Pa, int32_t A);

/* Enqueue the fused loss pass.  `stream` is a cudaStream_t. */
int hrl_loss_fwd_bwd(const HrlLossArgs *args, void *stream);

/*
 * Stand-alone compute_target (losses.py:63-80) on (B,T,P) columns.
 *   values   (B,T,P) or NULL  -> targets = advantages = returns (losses.py:64-66)
 *   returns  (B,Tr,P) with Tr == T or Tr == 1 (the (B,1,P,1) outcome of train.py:254)
 *   rewards  (B,T,P) or NULL (= 0)
 *   rhos, cs (B,T,Pr) with Pr == P or Pr == 1 (broadcast over players)
 *   masks    (B,T,P)
 *   targets, advantages (B,T,P) outputs
 */
int hrl_compute_target(int32_t algo, int32_t B, int32_t T, int32_t P, int32_t Tr, int32_t Pr,
                       const float *values, const float *returns, const float *rewards,
                       float lambda, float gamma, const float *rhos, const float *cs,
                       const float *masks, float *targets, float *advantages, void *stream);

/*
 * Optimiser step on one flat fp32 parameter bucket (train.py:370-371).
 *
 * hrl_grad_sumsq writes per-block partial sums of grad^2 into `partials`
 * (hrl_sumsq_num_partials() floats); hrl_clip_adam_step reduces them in a fixed order,
 * applies clip_grad_norm_(max_norm) and one Adam step with L2 weight decay
 * (torch.optim.Adam semantics: grad += wd * param before the moments).
 *
 * `lr` and `step` live on the device so that a captured CUDA graph stays valid while
 * the learner changes the learning rate every epoch (train.py:383-384); the kernel
 * increments *step.
 */
int32_t hrl_sumsq_num_partials(void);
int hrl_grad_sumsq(const float *grad, int64_t n, float *partials, void *stream);
int hrl_clip_adam_step(float *param, const float *grad, float *exp_avg, float *exp_avg_sq,
                       int64_t n, const float *partials, const float *lr, int64_t *step,
                       double max_norm, double beta1, double beta2, double eps, double weight_decay,
                       float *grad_norm_out /* may be NULL */, void *stream);

/*
 * Multi-GPU form of the same step: one-shot all-reduce (SUM) of the flat gradient bucket over NVLink peer
 * memory, fused with the sum-of-squares partials hrl_clip_adam_step consumes (replaces NCCL all-reduce +
 * hrl_grad_sumsq).  Every rank reads every rank's bucket directly (P2P loads through NVSwitch), adds them in
 * rank order -- all ranks get bit-identical sums -- and writes the result to `out_sum`.
 *   peer_buckets  device array [world] of pointers: this process's mapping of rank r's bucket (index r);
 *                 each bucket holds n floats followed by 2*world uint32 flags (zero-initialised once)
 *   flag_offset   index (in 32-bit words from the bucket start) of the flags
 *   n, n_norm     floats to reduce; the first n_norm of them (the gradients proper, not the loss sums riding in
 *                 the tail) enter the sum of squares
 *   epoch, ticket device uint32, zero-initialised once; maintained by the kernel (CUDA-graph safe)
 *   status        device uint32, zero-initialised once; set to 1 (and left there) if a peer rank did not arrive
 *                 within 20 s -- the kernel then gives up instead of spinning forever and its sums are invalid
 * Ranks synchronise inside the kernel with release/acquire flags at system scope: "my gradients are ready"
 * before the loads, "I am done reading" after them, so the next step may overwrite the bucket.
 */
int hrl_peer_allreduce_sumsq(float *out_sum, const float *const *peer_buckets, int64_t flag_offset, int32_t world,
                             int32_t rank, int64_t n, int64_t n_norm, float *partials, uint32_t *epoch, uint32_t *ticket,
                             uint32_t *status, void *stream);

/*
 * fp32-accurate matrix product on the tensor cores (tcgen05.mma kind::tf32 with the 3xTF32 hi/lo split, fp32
 * accumulation in tensor memory) -- the dense contractions of the user's net (fastnet.py runs a convolution over a
 * tiny board as one such product per direction), i.e. what torch.nn.functional.linear / conv2d and their autograd
 * do inside reference train.py:142-146, 369.
 *     C[M x N] = A_op[M x K] * B_op[N x K]^T (+ bias[N])
 *   a_kmajor / b_kmajor  1: element (row, k) of the operand at  row * ld + k  (reduction dimension contiguous)
 *                        0: at  k * ld + row  (the operand is stored transposed, e.g. reduce over samples)
 *   splits               K slices computed by separate CTAs into `workspace` (hrl_gemm_workspace_floats floats) and
 *                        summed in a fixed order (deterministic); 1 = no split, workspace may be NULL.  A split
 *                        product takes no bias and needs ldc == N; with C == NULL the slice partials
 *                        (hrl_gemm_effective_splits of them, M*N floats apart) are left in the workspace.
 * Relative error ~1e-6 of sum_k |a||b| (plain fp32 summation is ~1e-7 * sqrt(K)); NOT the 1e-3 of single-pass TF32.
 */
size_t hrl_gemm_workspace_floats(int64_t M, int64_t N, int64_t K, int32_t splits);
int32_t hrl_gemm_effective_splits(int64_t K, int32_t splits);   /* slices really produced (whole 32-element chunks) */
int hrl_gemm_tf32x3(const float *A, int64_t lda, int32_t a_kmajor, const float *B, int64_t ldb, int32_t b_kmajor,
                    const float *bias, float *C, int64_t ldc, int64_t M, int64_t N, int64_t K, int32_t splits,
                    float *workspace, void *stream);

/*
 * The same product with the elementwise neighbours of a conv -> BatchNorm -> ReLU tower fused into it, so that a layer of
 * the user's net is ONE launch per direction (forward / input gradient / weight gradient) instead of a product plus
 * separate normalisation, activation and reduction passes over the (samples x features) activations:
 *   operand transform   v = x * p[f] + y * q[f] + r[f], optionally clamped at 0, applied while the operand is staged:
 *                       BatchNorm-apply + ReLU of the previous layer (x = its raw output, p = gamma*rstd, r = beta - mean*p),
 *                       or the BatchNorm backward dY = dZ*p + Y*q + r (x = dZ, y = Y).  f = the reduction index, or the
 *                       operand row when feature_is_row (transposed operands of the weight-gradient product).
 *   epilogues           RELU: C = max(acc + bias, 0).  STATS: C = acc and per-column sum / sum of squares of the tile
 *                       (BatchNorm batch statistics of this layer's output).  MASK_STATS: C = acc * (z > 0) with
 *                       z = y*scale + shift of the pre-activation y (ReLU backward) and per-column sums of C and
 *                       C * (y - mean) * rstd (the two batch sums of the BatchNorm backward).
 * Column sums land in col_partials[row_tile][2][N] (row_tile = ceil(M/128) tiles, summed by hrl_bn_finalize_*).
 */
typedef struct HrlGemmOperand {
    const float *ptr;            /* the operand as it lies in memory                                   */
    const float *ptr2;           /* optional second source with the same layout (y above), or NULL     */
    const float *p, *q, *r;      /* per-feature constants, NULL = plain operand (q only with ptr2)     */
    int64_t ld;
    int32_t kmajor;              /* 1: element (row,k) at row*ld + k; 0: at k*ld + row                 */
    int32_t relu;
    int32_t feature_is_row;
    int32_t packed;              /* B only: ptr is an hrl_board_pack image (weights pre-split into TF32 hi/lo halves and
                                    pre-swizzled, one contiguous block per 32-element chunk: staged by ONE bulk copy)  */
} HrlGemmOperand;

typedef enum { HRL_GEMM_EP_STORE = 0, HRL_GEMM_EP_RELU = 1, HRL_GEMM_EP_STATS = 2, HRL_GEMM_EP_MASK_STATS = 3 } HrlGemmEpilogue;

typedef struct HrlGemmArgs {
    HrlGemmOperand a, b;         /* C[M x N] = A_op[M x K] * B_op[N x K]^T                            */
    const float *bias;           /* per column, or NULL                                                */
    float *C;
    int64_t ldc, M, N, K;
    int32_t splits;              /* as hrl_gemm_tf32x3 (plain epilogue only)                           */
    int32_t epilogue;            /* HrlGemmEpilogue                                                    */
    float *workspace;
    const float *ep_y;           /* MASK_STATS: pre-activation tile (M x N)                            */
    int64_t ep_ldy;
    const float *ep_scale, *ep_shift;   /* per column; NULL = 1 / 0                                    */
    const float *ep_mean, *ep_rstd;     /* per column; NULL = second sum is sum(C * y)                 */
    float *col_partials;         /* [ceil(M/128)][2][N] floats, or NULL                                */
    /* Convolution over a board as an implicit product (replaces cuDNN's fp32 SIMT kernels for the stride-1 "same" / wrap-around
     * convolutions of the board nets: reference geister.py:18-56 ConvLSTM cells, hungry_geese.py:20-37 TorusConv2d).  Activations
     * are channels-last: a row is a pixel, ld its stride in floats.  conv_off[cell * taps + tap] = (cell the tap reads) - cell, or
     * HRL_CONV_OUTSIDE under zero padding; wrap-around boards have no outside (hrl_conv_geometry fills it).
     *   conv_mode 1  forward / input gradient: A = input pixels (M = pixels, a.ld >= conv_cin), B = hrl_conv_pack image, K = taps *
     *                (conv_cin padded to a multiple of 32); the input gradient is the same product of dy with the adjoint image.
     *   conv_mode 2  weight gradient: A = dy [pixels][Cout] (kmajor 0), B = input [pixels][conv_cin] (kmajor 0), N = taps * conv_cin,
     *                K = pixels, split over K slices; hrl_conv_wgrad_reduce sums the slices into the (Cout, Cin, kh, kw) gradient. */
    const int16_t *conv_off;
    int32_t conv_mode, conv_hw, conv_taps, conv_cin;
    /* conv_mode 2 over several (dy, x) pairs that share ONE weight -- a recurrent cell applied at every time step: segments > 0,
     * seg_a[i] / seg_b[i] (host arrays of device pointers, at most 64) replace a.ptr / b.ptr, every pair has K pixels and gets
     * `splits` K slices; C must be NULL, the (segments * effective splits) slice partials stay in the workspace for
     * hrl_conv_wgrad_reduce2.  conv_ones_row = 1 appends a B row of ones: N = taps * conv_cin + 1, and the last column of the result is
     * sum_pixels dy = the bias gradient. */
    const float *const *seg_a;
    const float *const *seg_b;
    int32_t segments, conv_ones_row;
} HrlGemmArgs;

#define HRL_CONV_OUTSIDE (-32768)
/* table (H*W*kh*kw int16, host memory) of the neighbour offsets above; wrap != 0: the board is a torus */
int hrl_conv_geometry(int32_t H, int32_t W, int32_t kh, int32_t kw, int32_t wrap, int16_t *table);
/* weights (Cout, Cin, kh, kw) -> packed B images (zero them once): forward (rows = Cout, reduction = tap-major, Cin padded to 32)
 * and adjoint (rows = Cin, reduction = flipped tap, Cout padded to 32).  Either may be NULL. */
size_t hrl_conv_pack_floats(int32_t rows, int32_t channels, int32_t taps);
int hrl_conv_pack(const float *w, int32_t Cout, int32_t Cin, int32_t kh, int32_t kw, float *image_fwd, float *image_adj, void *stream);
/* dw[co][ci][a][b] = sum over slices of partials[s][co][(a*kw+b)*Cin + ci]  (fixed order) */
int hrl_conv_wgrad_reduce(const float *partials, int32_t splits, float *dw, int32_t Cout, int32_t Cin, int32_t taps, void *stream);
/* the same over partial rows of `ncols` floats (taps*Cin, or taps*Cin + 1 with the ones row: column taps*Cin -> db[co], may be NULL);
 * accumulate != 0 adds to dw / db instead of overwriting them */
int hrl_conv_wgrad_reduce2(const float *partials, int32_t splits, int32_t ncols, float *dw, float *db, int32_t Cout, int32_t Cin, int32_t taps,
                           int32_t accumulate, void *stream);

int hrl_gemm_fused(const HrlGemmArgs *args, void *stream);

/*
 * Glue of a fused conv -> BatchNorm -> ReLU tower over a tiny board (handyrl_b200/tower.py: the architecture of the
 * reference's SimpleConv2dModel, envs/tictactoe.py:52-69, with every layer ONE hrl_gemm_fused launch per direction).
 *   hrl_bn_finalize_fwd   col_partials [tiles][2][C*HW] (column sum / sum of squares from the STATS epilogue) -> batch
 *                         statistics of nn.BatchNorm2d in training mode (running stats with `momentum`, unbiased running
 *                         variance, num_batches_tracked += 1) and per-COLUMN mean / rstd / scale = gamma*rstd /
 *                         shift = beta - mean*scale (each C*HW floats) for the next product's operand transform
 *   hrl_bn_finalize_bwd   col_partials (column sums of dZ and dZ*xhat from the MASK_STATS epilogue) -> dgamma, dbeta and
 *                         the per-column constants of dY = dZ*p + Y*q + r.  gamma == NULL: only dbeta (a plain bias).
 *   hrl_heads_fwd / _bwd  squeeze outputs pre (M, ld), columns [pmaps*cells | vmaps*cells | rmaps*cells] -> LeakyReLU(slope)
 *                         -> policy = . Wp^T (A x pmaps*cells), value = tanh(. Wv^T), return = . Wr^T; the backward
 *                         writes dpre and the gradients of Wp / Wv / Wr and of the squeeze biases (fixed-order sums;
 *                         workspace: hrl_heads_num_blocks(M) * (A*pin + vin + rin + maps) floats).
 */
int hrl_bn_finalize_fwd(const float *col_partials, int32_t tiles, int32_t C, int32_t HW, int64_t rows, const float *gamma,
                        const float *beta, float eps, float momentum, float *running_mean, float *running_var,
                        int64_t *batches_tracked, float *mean_col, float *rstd_col, float *scale_col, float *shift_col, void *stream);
int hrl_bn_finalize_bwd(const float *col_partials, int32_t tiles, int32_t C, int32_t HW, int64_t rows, const float *gamma,
                        const float *mean_col, const float *rstd_col, float *dgamma, float *dbeta, float *p_col, float *q_col,
                        float *r_col, void *stream);
int32_t hrl_heads_num_blocks(int64_t M);
int hrl_heads_fwd(const float *pre, int64_t ld, int64_t M, int32_t cells, int32_t pmaps, int32_t vmaps, int32_t rmaps, int32_t A,
                  float slope, const float *Wp, const float *Wv, const float *Wr, float *policy, float *value, float *ret, void *stream);
int hrl_heads_bwd(const float *pre, int64_t ld, int64_t M, int32_t cells, int32_t pmaps, int32_t vmaps, int32_t rmaps, int32_t A,
                  float slope, const float *Wp, const float *Wv, const float *Wr, const float *value, const float *dpolicy,
                  const float *dvalue, const float *dret, float *dpre, float *dWp, float *dWv, float *dWr, float *dbias_p,
                  float *dbias_v, float *dbias_r, float *workspace, void *stream);

/*
 * Weight of a stride-1 "same" convolution (Cout,Cin,kh,kw; odd kernel, zero padding) <-> the dense matrix
 * (Cout*H*W, Cin*H*W) that applies it to an H x W board stored NCHW (fastnet.BoardConv2d), and the adjoint map
 * dense-gradient -> weight-gradient.  dense[(o,q),(i,p)] = w[o,i,a,b] where tap (a,b) makes output cell q read input
 * cell p, 0 if no tap does.  hrl_board_fold sums `splits` dense gradients `split_stride` floats apart (the K-slice
 * partials a split hrl_gemm_tf32x3 leaves in its workspace when called with C == NULL; 1 for a single matrix).
 */
int hrl_board_expand(const float *w, float *dense, int32_t Cout, int32_t Cin, int32_t kh, int32_t kw, int32_t H, int32_t W,
                     void *stream);
/*
 * hrl_board_pack: the same dense matrix, written directly as the B-operand images of hrl_gemm_fused (HrlGemmOperand.packed)
 * for the forward product (rows = output features (o,q), reduction = input features (i,p)) and for the input-gradient
 * product (rows = input features, reduction = output features).  Image layout: [chunk of 32 reduction elements][hi | lo]
 * [n_pad rows][128 bytes, 16-byte slots XOR-swizzled by the row], n_pad = hrl_gemm_padded_rows(rows of the operand);
 * row0 / k0 place several convolutions side by side in one operand (e.g. the policy / value squeeze convolutions).
 * Padding rows and the reduction tail must be zero: zero the images once, they are never written.
 */
/* Profiling / test hook of hrl_gemm_fused, process-global and NOT thread safe (default 0 = the product path):
 * 1 = skip the MMAs, 2 = skip the operand loads (timing attribution, scripts/gemm_breakdown.py);
 * +64 = read the A operand straight from global memory even where it could be staged through shared memory (the two
 * paths must agree bit for bit: tests/test_gemm_gpu.py).  Results with 1 or 2 set are garbage by design. */
void hrl_gemm_set_debug(int mode);
int32_t hrl_gemm_padded_rows(int64_t N);
size_t hrl_board_pack_floats(int64_t rows, int64_t K);      /* floats of an image with `rows` operand rows over K */
int hrl_board_pack(const float *w, int32_t Cout, int32_t Cin, int32_t kh, int32_t kw, int32_t H, int32_t W, float *image_fwd,
                   int32_t fwd_rows, int32_t fwd_row0, float *image_bwd, int32_t bwd_rows, int32_t bwd_k0, void *stream);
/* Batched forms: up to HRL_MAX_BOARD_JOBS convolutions in ONE launch (a net packs all its layers' weights once per step, and
 * folds all its weight gradients once per backward pass).  bias / bias_cells (optional, both or neither): the convolution's
 * bias replicated per cell, bias_cells[c*H*W + q] = bias[c] -- the per-column bias of the dense product. */
#define HRL_MAX_BOARD_JOBS 8
typedef struct HrlPackJob {
    const float *w;
    int32_t Cout, Cin, kh, kw, H, W;
    float *image_fwd;
    int32_t fwd_rows, fwd_row0;
    float *image_bwd;
    int32_t bwd_rows, bwd_k0;
    const float *bias;
    float *bias_cells;
} HrlPackJob;
typedef struct HrlFoldJob {
    const float *ddense;
    int32_t splits;
    int64_t split_stride;
    float *dw;
    int32_t Cout, Cin, kh, kw, H, W;
} HrlFoldJob;
int hrl_board_pack_many(const HrlPackJob *jobs, int32_t n_jobs, void *stream);
int hrl_board_fold_many(const HrlFoldJob *jobs, int32_t n_jobs, void *stream);
int hrl_board_fold(const float *ddense, int32_t splits, int64_t split_stride, float *dw, int32_t Cout, int32_t Cin, int32_t kh,
                   int32_t kw, int32_t H, int32_t W, void *stream);

/*
 * Recurrent nets.  ConvLSTM gate arithmetic (reference geister.py:49-56): gates (N,4C,S) in the order
 * i, f, o, g are the cell's convolution output; c' = sig(f) c + sig(i) tanh(g), h' = sig(o) tanh(c').  The backward
 * recomputes the activations from `gates` and `c_prev`; dh / dc_out may be NULL (no gradient from that side).
 */
int hrl_lstm_gates_fwd(const float *gates, const float *c_prev, float *h_out, float *c_out, int64_t N, int32_t C, int32_t S,
                       void *stream);
int hrl_lstm_gates_bwd(const float *gates, const float *c_prev, const float *dh, const float *dc_out, float *dgates,
                       float *dc_prev, int64_t N, int32_t C, int32_t S, void *stream);

/*
 * Hidden-state masking of the recurrent time loop (reference train.py:152-158, 173) on a hidden leaf h (B,P,R):
 *   visible: out = h * om[b,p]            (sum_players = 0, out (B,P,R))
 *            out = sum_p h * om[b,p]      (sum_players = 1, out (B,R): turn-alternating batches)
 *   blend:   out = h (1 - om) + h_new om  (h_new (B,Pn,R), Pn == P or 1)
 * om points at observation_mask[:, t] = element (b,p) at om[b * om_stride + p].
 */
int hrl_hidden_visible_fwd(const float *h, const float *om, int64_t om_stride, float *out, int64_t B, int32_t P, int32_t R,
                           int32_t sum_players, void *stream);
int hrl_hidden_visible_bwd(const float *dout, const float *om, int64_t om_stride, float *dh, int64_t B, int32_t P, int32_t R,
                           int32_t sum_players, void *stream);
int hrl_hidden_blend_fwd(const float *h, const float *nh, const float *om, int64_t om_stride, float *out, int64_t B, int32_t P,
                         int32_t Pn, int32_t R, void *stream);
int hrl_hidden_blend_bwd(const float *dout, const float *om, int64_t om_stride, float *dh, float *dnh, int64_t B, int32_t P,
                         int32_t Pn, int32_t R, void *stream);

/*
 * Train-mode BatchNorm over (N, C, HW) fp32 activations with small HW -- used by the small-board rewrite of the user's
 * net (the nets of reference envs normalise (N,32,3,3) / (N,C,6,6) tensors; cuDNN / ATen launch one CTA per channel
 * there).  Semantics of nn.BatchNorm2d in training mode: biased variance for the normalisation, unbiased for
 * running_var, running stats updated with `momentum` (pass NULL for both to skip).  mean / rstd (C floats each) are
 * saved for the backward.  channels_last = 1: x / y / dy / dx are (N, H, W, C) in memory (torch.channels_last), e.g. the
 * activations between cuDNN's NHWC convolutions.  workspace: hrl_bn_workspace_floats(N, C, HW, channels_last) floats.
 */
size_t hrl_bn_workspace_floats(int64_t N, int32_t C, int32_t HW, int32_t channels_last);
int hrl_bn_train_fwd(const float *x, const float *gamma, const float *beta, float *y, float *mean, float *rstd,
                     float *running_mean, float *running_var, int64_t N, int32_t C, int32_t HW, int32_t channels_last, float eps, float momentum,
                     float *workspace, void *stream);
int hrl_bn_train_bwd(const float *x, const float *dy, const float *gamma, const float *mean, const float *rstd, float *dx,
                     float *dgamma, float *dbeta, int64_t N, int32_t C, int32_t HW, int32_t channels_last, float *workspace,
                     void *stream);

/*
 * Replay gather/pad: the device form of make_batch (train.py:33-124).
 *
 * Episodes are kept decoded in flat device arrays ("replay store"), one row per step:
 *   per step s and player p (value side, Ps players):  obs_mask, turn_mask(=selected_prob present),
 *   selected_prob, action, value, reward, return; per step: action_mask rows, observation rows.
 * A window descriptor selects [start, end) of one episode and where it lands in [0, T).
 */
typedef struct HrlWindow {
    int64_t first_step;   /* row of the episode's step 0 in the store                          */
    int32_t start, end;   /* window [start, end) in episode steps (train.py:305-306)           */
    int32_t train_start;  /* first trained step (train.py:304); pad_before = burn_in-(train_start-start) */
    int32_t total;        /* episode length, for progress = step / total (train.py:89)        */
    int32_t outcome_row;  /* row in the store's outcome table                                  */
    int32_t player;       /* solo training (train.py:57-58): the one store player batched; else 0 */
} HrlWindow;

typedef struct HrlGatherArgs {
    int32_t B, T, P, Pa, A;
    int32_t Ps;                   /* players per step in the store (P == Ps, or P == 1 for solo) */
    int32_t burn_in;
    int32_t obs_elems;            /* floats per observation (flattened leaf)                   */
    int32_t turn_alternating;     /* Pa == 1 layout: pick the turn player's row (train.py:65-66) */
    const HrlWindow *windows;     /* [B] device                                                */

    /* replay store, S = total stored steps */
    const float *st_obs;          /* (S,Ps,obs_elems) zero where the player did not observe     */
    const float *st_prob;         /* (S,Ps) 1.0 where absent                                   */
    const int32_t *st_action;     /* (S,Ps) 0 where absent                                     */
    const float *st_amask;        /* (S,Ps,A) 1e32 where absent                                 */
    const float *st_value;        /* (S,Ps)                                                    */
    const float *st_reward;       /* (S,Ps)                                                    */
    const float *st_return;       /* (S,Ps)                                                    */
    const uint8_t *st_flags;      /* (S,Ps) bit0 = acted (turn_mask), bit1 = observed           */
    const int32_t *st_turn;       /* (S)   index of the step's first turn player (train.py:66)         */
    const float *st_outcome;      /* (E,Ps)                                                    */

    /* batch outputs, layouts of train.py:114-124 */
    float *observation;           /* (B,T,Pa,obs_elems)                                        */
    float *selected_prob;         /* (B,T,Pa)                                                  */
    float *value;                 /* (B,T,P)                                                   */
    int64_t *action;              /* (B,T,Pa)                                                  */
    float *outcome;               /* (B,P)                                                     */
    float *reward;                /* (B,T,P)                                                   */
    float *ret;                   /* (B,T,P)                                                   */
    float *episode_mask;          /* (B,T)                                                     */
    float *turn_mask;             /* (B,T,P)                                                   */
    float *observation_mask;      /* (B,T,P)                                                   */
    float *action_mask;           /* (B,T,Pa,A)                                                */
    float *progress;              /* (B,T)                                                     */
} HrlGatherArgs;

int hrl_gather_pad(const HrlGatherArgs *args, void *stream);

/* Text of the last error raised on the calling thread ("" if none). */
const char *hrl_last_error(void);

/* HRL_ABI_VERSION the library was built with. */
int32_t hrl_abi_version(void);

#ifdef __cplusplus
}
#endif
#endif /* HRL_B200_H_ */
