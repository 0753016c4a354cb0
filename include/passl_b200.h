/* passl_b200 — C ABI of the B200-native PASSL hot path (libpassl_b200.so).
 *
 * The reference (PaddlePaddle/PASSL @ 5c7359b) has no C/FFI seam: its hot path calls Paddle library ops from
 * Python (SURVEY.md §2.2, §8 b).  This header is the seam a maintainer would bind instead: every entry point
 * replaces the Paddle op(s) cited next to it.  Conventions:
 *   - plain device pointers + sizes, no framework types; all tensors contiguous unless a leading dim is passed;
 *   - stream-ordered on `stream` (a cudaStream_t), no allocation, no hidden host sync;
 *   - return 0 on success, >0 = cudaError_t, <0 = contract violation (PB_ERR_* below);
 *   - activations are NHWC / [tokens, features] bf16, statistics / losses / embeddings fp32;
 *   - integer state (queue pointer, labels, indices) is int64 like the reference's Paddle tensors.
 */
#ifndef PASSL_B200_H_
#define PASSL_B200_H_

#ifdef __cplusplus
extern "C" {
#endif

#define PASSL_B200_ERR_BAD_ARG (-1)
#define PASSL_B200_ERR_UNSUPPORTED (-2)
#define PASSL_B200_ERR_TMAP (-3)
#define PASSL_B200_ERR_WORKSPACE (-4)

/* activation codes for fused epilogues */
#define PASSL_B200_ACT_NONE 0
#define PASSL_B200_ACT_RELU 1
#define PASSL_B200_ACT_GELU 2      /* exact erf GELU: paddle nn.GELU (passl/models/vision_transformer.py:98) */
#define PASSL_B200_ACT_QUICKGELU 3 /* x*sigmoid(1.702x): passl_v110/modeling/backbones/base_transformer.py:25-28 */

int passl_b200_version(void);

/* ---------------------------------------------------------------------------------------------------------------
 * Dense contraction (tcgen05 + TMA).  Replaces paddle nn.Linear / paddle.matmul (cuBLAS) at
 * passl/models/vision_transformer.py:107-113,145-153, passl_v110/modeling/necks/base_neck.py:83-97 and every
 * 1x1 convolution of resnetimagenet.py:112-131.
 *   out[M,N] = act(alpha * A.B^T + bias) + residual
 *   A: a_mn_major==0 -> [M,K] row-major (lda);  ==1 -> [K,M] row-major (lda)      bf16
 *   B: b_mn_major==0 -> [N,K] row-major (ldb);  ==1 -> [K,N] row-major (ldb)      bf16
 *   out: bf16 (out_fp32==0) or fp32; atomic_add (fp32 only) accumulates with red.add (needed for splits>1)
 *   col_sum / col_sqsum: optional fp32 [N] accumulators of per-column sum / sum of squares of the stored values
 * ------------------------------------------------------------------------------------------------------------- */
int passl_b200_gemm_bf16(const void* A, const void* B, void* out, int M, int N, int K, int a_mn_major,
                         int b_mn_major, long long lda, long long ldb, long long ldc, int out_fp32, int atomic_add,
                         const float* bias, const void* residual, int act, float alpha, int splits, float* col_sum,
                         float* col_sqsum, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * Convolution as implicit GEMM over NHWC bf16 (replaces paddle nn.Conv2D -> cuDNN fwd/dgrad/wgrad at
 * passl_v110/modeling/backbones/resnetimagenet.py:112-131,190-206).  Weights are [Cout, R, S, Cin] bf16.
 * Cin (fwd) / Cout (dgrad) must be a multiple of 64; stride 1 or 2 (H, W even when stride 2).
 * ------------------------------------------------------------------------------------------------------------- */
int passl_b200_conv2d_fwd_bf16(const void* x, const void* w, void* out, int N, int H, int W, int Cin, int Cout, int R,
                               int S, int stride, int pad, const float* bias, const void* residual, int act,
                               float* col_sum, float* col_sqsum, void* stream);
long long passl_b200_conv2d_dgrad_workspace_bytes(int Cin, int Cout, int R, int S);
int passl_b200_conv2d_dgrad_bf16(const void* dy, const void* w, void* dx, void* workspace, int N, int H, int W, int Cin,
                                 int Cout, int R, int S, int stride, int pad, int accumulate, void* stream);
int passl_b200_conv2d_wgrad_bf16(const void* x, const void* dy, float* dw, int N, int H, int W, int Cin, int Cout, int R,
                                 int S, int stride, int pad, int zero_first, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * Fused similarity -> scaled softmax -> cross-entropy (InfoNCE), fp32 SIMT variant.
 * Replaces matmul + concat + divide + CrossEntropyLoss + topk of
 *   passl_v110/modeling/architectures/moco.py:178-182 + heads/contrastive_head.py:37-60   (P != NULL: [l_pos|l_neg], label 0)
 *   passl/models/mocov3.py:187-198                                                       (label = arange(N)+N*rank)
 *   passl_v110/modeling/backbones/clip.py:331-335 + heads/clip_head.py:27-35              (label = arange(n))
 * A [N,D] fp32 queries; B [K,D] keys (fp32, or bf16 when b_is_bf16); P [N,D] optional positive keys;
 * label int64 [N] column index (ignored when P given); excl int32 [N] excluded column or -1 (may be NULL).
 * out_scalars = {loss_scale * mean_i loss_i, top-1 %, top-5 %};  lse/tgt [N] are saved for backward.
 * ------------------------------------------------------------------------------------------------------------- */
long long passl_b200_simce_workspace_bytes(int N, int K);
int passl_b200_simce_fwd_f32(const float* A, const void* B, int b_is_bf16, const float* P, const long long* label,
                             const int* excl, float scale, float loss_scale, int N, int K, int D, float* lse, float* tgt,
                             float* loss_rows, float* out_scalars, void* workspace, long long workspace_bytes,
                             void* stream);
int passl_b200_simce_bwd_f32(const float* A, const void* B, int b_is_bf16, const float* P, const long long* label,
                             const int* excl, float scale, float loss_scale, int N, int K, int D, const float* lse,
                             const float* tgt, const float* dloss, float* dA, void* workspace, long long workspace_bytes,
                             void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * Fused InfoNCE forward on tcgen05 (bf16 operands, fp32 accumulate in TMEM, online softmax out of TMEM).
 * Same contract as passl_b200_simce_fwd_f32 with Q and Kmat in bf16; the key matrix is streamed from HBM exactly
 * once.  D multiple of 64, <= 512.
 * ------------------------------------------------------------------------------------------------------------- */
long long passl_b200_infonce_tc_workspace_bytes(int N, int K, int D);
int passl_b200_infonce_tc_fwd(const void* Q, const void* Kmat, const float* P, const long long* label, const int* excl,
                              float scale, float loss_scale, int N, int K, int D, float* lse, float* tgt,
                              float* loss_rows, float* out_scalars, void* workspace, long long workspace_bytes,
                              void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * Embedding utilities.
 *   l2norm mode 0: x/max(||x||,eps) (paddle F.normalize, moco.py:159,170; mocov3.py:189-190)
 *          mode 1: x/sqrt(sum x^2+eps) (passl/nn/norm.py:18-40; simclr.py:58)
 *          mode 2: x/||x|| (clip.py:325-328)
 *   queue_enqueue: moco.py:92-105 (_dequeue_and_enqueue) on a key-major [K,D] ring buffer; queue_ptr int64[1]
 *                  lives on the device (the reference's int(queue_ptr[0]) D2H sync is gone). K % Bg != 0 -> BAD_ARG
 *                  (moco.py:99 assert).
 *   ema_update:    moco.py:82-90 over one flat fp32 parameter buffer (+ bf16 compute copy).
 * ------------------------------------------------------------------------------------------------------------- */
int passl_b200_l2norm_fwd(const float* x, float* y, void* y_bf16, float* inv_norm, int N, int D, int mode, float eps,
                          void* stream);
int passl_b200_l2norm_bwd(const float* dy, const float* y, const float* inv_norm, float* dx, void* dx_bf16, int N, int D,
                          int mode, float eps, void* stream);
int passl_b200_queue_enqueue(const float* keys, float* queue_f32, void* queue_bf16, long long* queue_ptr, int Bg, int D,
                             int K, void* stream);
int passl_b200_ema_update(float* k, const float* q, void* k_bf16, float m, long long n, void* stream);
int passl_b200_cast_f32_to_bf16(const float* x, void* y, long long n, void* stream);
int passl_b200_cast_bf16_to_f32(const void* x, float* y, long long n, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PASSL_B200_H_ */
