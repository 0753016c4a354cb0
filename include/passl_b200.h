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
/* number of kernels this library has launched so far in the process (bench.py's gpu_launches) */
long long passl_b200_launch_count(void);
long long passl_b200_launch_counter_add(long long n);

/* ---------------------------------------------------------------------------------------------------------------
 * Dense contraction (tcgen05 + TMA).  Replaces paddle nn.Linear / paddle.matmul (cuBLAS) at
 * passl/models/vision_transformer.py:107-113,145-153, passl_v110/modeling/necks/base_neck.py:83-97 and every
 * 1x1 convolution of resnetimagenet.py:112-131.
 *   out[M,N] = act(alpha * A.B^T + bias) + residual
 *   A: a_mn_major==0 -> [M,K] row-major (lda);  ==1 -> [K,M] row-major (lda)      bf16
 *   B: b_mn_major==0 -> [N,K] row-major (ldb);  ==1 -> [K,N] row-major (ldb)      bf16
 *   out: bf16 (out_fp32==0) or fp32; atomic_add (fp32 only) accumulates with red.add (needed for splits>1)
 *   col_sum: optional fp32 [gemm_stats_rows()][2][N] per-CTA partials of the per-column sum / sum of squares of the stored bf16
 *            values (BatchNorm batch statistics fused into the producing GEMM: bf16 out, splits == 1); zeroed and filled by the
 *            call, row b written only by CTA b (no atomics); feed it to bn_finalize with nblk = gemm_stats_rows().
 *            col_sqsum is reserved (pass NULL).
 * ------------------------------------------------------------------------------------------------------------- */
int passl_b200_gemm_stats_rows(void);
int passl_b200_gemm_bf16(const void* A, const void* B, void* out, int M, int N, int K, int a_mn_major,
                         int b_mn_major, long long lda, long long ldb, long long ldc, int out_fp32, int atomic_add,
                         const float* bias, const void* residual, int act, float alpha, int splits, float* col_sum,
                         float* col_sqsum, void* stream);
/* + gate after the activation: aux (bf16, addressed like out), aux_mode 1 = ReLU mask (aux>0), 2 = *GELU'(aux),
 *   3 = *QuickGELU'(aux) — fuses the activation backward into the dgrad GEMM;
 *   preact_out (bf16, addressed like out, optional): the value before the activation, kept for that backward. */
int passl_b200_gemm_bf16_ex(const void* A, const void* B, void* out, int M, int N, int K, int a_mn_major,
                            int b_mn_major, long long lda, long long ldb, long long ldc, int out_fp32, int atomic_add,
                            const float* bias, const void* residual, int act, float alpha, int splits, float* col_sum,
                            float* col_sqsum, const void* aux, int aux_mode, void* preact_out, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * Convolution as implicit GEMM over NHWC bf16 (replaces paddle nn.Conv2D -> cuDNN fwd/dgrad/wgrad at
 * passl_v110/modeling/backbones/resnetimagenet.py:112-131,190-206).  Weights are [Cout, R, S, Cin] bf16.
 * Cin (fwd) / Cout (dgrad) must be a multiple of 64; stride 1 or 2 (H, W even when stride 2).
 * ------------------------------------------------------------------------------------------------------------- */
int passl_b200_conv2d_fwd_bf16(const void* x, const void* w, void* out, int N, int H, int W, int Cin, int Cout, int R,
                               int S, int stride, int pad, const float* bias, const void* residual, int act,
                               float* col_sum, float* col_sqsum, void* stream);
/* 3x3 / stride 1 / pad 1 weight gradients use the halo-tile kernel (wgrad_halo.cu) when the problem is large enough to amortise
 * its per-item epilogue; tuning knob: 0 auto (default), 1 whenever the shape is supported, 2 never. */
int passl_b200_wgrad_halo_mode(int mode);
/* rectangular filter, separate row / column padding, explicit output extent, stride 1 (the repacked 7x7/2 stem below) */
int passl_b200_conv2d_fwd_rect_bf16(const void* x, const void* w, void* out, int N, int H, int W, int Cin, int Cout, int R,
                                    int S, int pad_h, int pad_w, int Ho, int Wo, const float* bias, int act, float* col_sum,
                                    void* stream);
int passl_b200_conv2d_wgrad_rect_bf16(const void* x, const void* dy, float* dw, int N, int H, int W, int Cin, int Cout, int R,
                                      int S, int pad_h, int pad_w, int Ho, int Wo, int zero_first, void* stream);
/* ResNet stem conv 7x7/2 pad 3, 3 -> 64 (resnetimagenet.py:190-198) without an im2col matrix:
 *   stem_pack_input : NCHW fp32 [N,3,H,W] -> xp bf16 [N,H/2,W/2,64], xp[n,i,q,((dj*2+a)*2+b)*4+c] = img[n,c,2i+a,2(q+dj-2)+b]
 *   stem_pack_weight: w fp32 [64,kpad] ((r,s,c) order) -> wp bf16 [64,4,64] so that
 *                     conv7x7/2(img, w) == conv2d_fwd_rect(xp, wp, R=4, S=1, pad_h=2, pad_w=0, Ho=H/2, Wo=W/2)
 *   stem_unpack_wgrad: dw[64,kpad] += fold(dwp fp32 [64,4,64])  (dwp from conv2d_wgrad_rect on xp) */
int passl_b200_stem_pack_input(const float* img, void* xp, int N, int H, int W, void* stream);
int passl_b200_stem_pack_weight(const float* w, void* wp, int kpad, void* stream);
int passl_b200_stem_unpack_wgrad(const float* dwp, float* dw, int kpad, void* stream);
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
 * Fused InfoNCE on tcgen05 (bf16 operands, fp32 accumulate in TMEM, softmax out of TMEM), ONE launch per direction.
 * Same contract as passl_b200_simce_fwd_f32 / _bwd_f32 with Q and Kmat in bf16 (rows L2-normalised: |<q,k>| <= ~1); the
 * key matrix is streamed from HBM exactly once per direction.  D multiple of 64, <= 512 (forward), <= 256 (backward).
 * Replaces matmul + concat + /T + CrossEntropyLoss + topk and their autograd backward
 * (passl_v110/modeling/architectures/moco.py:178-182, heads/contrastive_head.py:37-60, passl/models/mocov3.py:187-198).
 * `workspace` of the forward is PERSISTENT STATE (passl_b200_infonce_tc_workspace_bytes(N, K, D) bytes): the caller
 * zero-fills it once before its first use (and after a failed launch) and keeps one buffer per (N, stream); every launch
 * leaves it ready for the next.  Slice partial sums are merged with float atomics (results are order-dependent in the
 * last bit).  The backward overwrites dQ fp32 [N, D].
 * ------------------------------------------------------------------------------------------------------------- */
long long passl_b200_infonce_tc_workspace_bytes(int N, int K, int D);
/* developer hook: per-CTA timeline (uint64 [grid*16], %globaltimer ns) written by subsequent launches; NULL disables */
int passl_b200_infonce_tc_set_debug(void* buf);
int passl_b200_infonce_tc_fwd(const void* Q, const void* Kmat, const float* P, const long long* label, const int* excl,
                              float scale, float loss_scale, int N, int K, int D, float* lse, float* tgt,
                              float* loss_rows, float* out_scalars, void* workspace, long long workspace_bytes,
                              void* stream);
int passl_b200_infonce_tc_bwd(const void* Q, const void* Kmat, const float* P, const long long* label, const int* excl,
                              float scale, float loss_scale, int N, int K, int D, const float* lse, const float* tgt,
                              const float* dloss, float* dQ, void* stream);
/* Fused compute + collective: the gathered-key InfoNCE of MoCo v3 / CLIP (passl/models/mocov3.py:173-198: k_all =
 * concat_all_gather(k); logits = q k_all^T / T; labels = arange(N) + N*rank) WITHOUT the all-gather.  The key matrix is the
 * concatenation of `world` bf16 shards [shard_rows, D] (shard_rows % 64 == 0), each read IN PLACE from its owner's peer-mapped
 * buffer by TMA, tile by tile, so the NVLink transfer overlaps the MMAs / softmax of the tiles already on chip.
 * shard_ptrs: HOST array of `world` (<= 8) device pointers = this process's mappings of every rank's shard buffer (own rank
 * included); my_flags: this rank's flag row (uint32[world]); the kernel consumes shard q once my_flags[q] >= epoch
 * (passl_b200_peer_publish_keys_bf16 on rank q).  Label mode only; other arguments as passl_b200_infonce_tc_fwd / _bwd. */
int passl_b200_infonce_tc_fwd_peer(const void* Q, const void* const* shard_ptrs, const void* my_flags, int world, int shard_rows,
                                   unsigned epoch, const long long* label, const int* excl, float scale, float loss_scale, int N,
                                   int D, float* lse, float* tgt, float* loss_rows, float* out_scalars, void* workspace,
                                   long long workspace_bytes, void* stream);
int passl_b200_infonce_tc_bwd_peer(const void* Q, const void* const* shard_ptrs, const void* my_flags, int world, int shard_rows,
                                   unsigned epoch, const long long* label, const int* excl, float scale, float loss_scale, int N,
                                   int D, const float* lse, const float* tgt, const float* dloss, float* dQ, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * SimCLR NT-Xent + CO2 (passl_v110/modeling/heads/simclr_contrastive_head.py:42-102) on a similarity matrix
 * S = [h1;h2] [h1_all;h2_all]^T / T  (fp32 [2n, 2m], produced by passl_b200_gemm_bf16; m = world*n gathered negatives,
 * rank selects the diagonal block).  fwd: out = {loss, acc1 (fraction), contrast, co2}; bwd: dS (bf16 [2n, 2m]).
 * ------------------------------------------------------------------------------------------------------------- */
long long passl_b200_ntxent_workspace_bytes(int n);
int passl_b200_ntxent_co2_fwd(const float* S, int n, int m, int rank, float co2_weight, float* out, void* workspace,
                              long long workspace_bytes, void* stream);
int passl_b200_ntxent_co2_bwd(const float* S, int n, int m, int rank, float co2_weight, const float* dloss, void* dS_bf16,
                              const void* workspace, void* stream);

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

/* ---------------------------------------------------------------------------------------------------------------
 * BatchNorm (training mode) on channels-last bf16 [P, C].  Replaces paddle nn.BatchNorm2D / BatchNorm1D (cuDNN BN
 * fwd/bwd) at resnetimagenet.py:112-131 and necks/base_neck.py:221-227.  Paddle conventions: eps 1e-5,
 * running = momentum*running + (1-momentum)*batch (momentum 0.9), biased batch variance.
 *   bn_reduce_blocks : nblk = number of per-CTA partials the two reduce kernels write for a [P, C] tensor
 *   bn_stats      : part[b][0][c] = sum_p y, part[b][1][c] = sum_p y^2 over the rows of CTA b (fp32 [nblk,2,C]; no atomics,
 *                   deterministic); the GEMM / conv forward can emit the same partials from its epilogue (col_sum, gemm_stats_rows() rows)
 *   bn_finalize   : sums the partials -> mean, invstd, scale = gamma*invstd, shift = beta - mean*scale, running stats
 *   bn_apply      : z = relu?(y*scale + shift + residual)  (bf16 and/or fp32 output)
 *   bn_bwd_reduce : partials of sum_g = sum_p g, sum_gx = sum_p g*xhat with g = dz * (z>0 if relu).  relu: 0 none, 1 mask
 *                   read from z (bf16, needed when a residual was added before the ReLU), 2 mask recomputed from y as
 *                   fma(y, scale, shift) > 0 — `z` then points to the fp32 [2, C] (scale, shift) rows and the activation
 *                   tensor is never read; 3 mask from the 1-bit-per-element tensor written by bn_apply_mask (`z` points to it):
 *                   16x less traffic than reading z back for the residual units (same conventions in bn_bwd_apply)
 *   bn_bwd_finalize: sums [2,C] = totals (= dbeta, dgamma), accumulated into the gradient buffers when given
 *   bn_bwd_apply  : dy = gamma*invstd*(g - sum_g/P - xhat*sum_gx/P) = k1*g + k2*y + k3; dres = g (residual-branch gradient)
 * ------------------------------------------------------------------------------------------------------------- */
int passl_b200_bn_reduce_blocks(long long P, int C);
int passl_b200_bn_stats(const void* y, float* part, long long P, int C, void* stream);
int passl_b200_bn_finalize(const float* part, int nblk, const float* gamma, const float* beta, float* mean,
                           float* invstd, float* scale, float* shift, float* running_mean, float* running_var,
                           long long count, float eps, float momentum, int C, void* stream);
/* use_global_stats (passl_v110/modules/freeze.py:17-23, MoCo key encoder) / eval: affine from running statistics */
int passl_b200_bn_global_affine(const float* running_mean, const float* running_var, const float* gamma, const float* beta,
                                float* mean, float* invstd, float* scale, float* shift, float eps, int C, void* stream);
/* y += a*x on fp32 vectors (bias / BN parameter gradient accumulation) */
int passl_b200_axpy_f32(float* y, const float* x, float a, long long n, void* stream);
int passl_b200_bn_apply(const void* y, const void* residual, const float* scale, const float* shift, void* z, float* z_f32,
                        long long P, int C, int relu, void* stream);
int passl_b200_bn_apply_mask(const void* y, const void* residual, const float* scale, const float* shift, void* z, void* relu_mask,
                             long long P, int C, int relu, void* stream);
int passl_b200_bn_bwd_reduce(const void* y, const void* dz, const void* z, const float* mean, const float* invstd,
                             float* part, long long P, int C, int relu, void* stream);
/* gamma / mean / invstd / count / coef may be NULL/0 (LayerNorm and bias-gradient uses); with them, coef [3, C] receives the
 * per-channel coefficients of  dy = k1*g + k2*y + k3  consumed by bn_bwd_apply. */
int passl_b200_bn_bwd_finalize(const float* part, int nblk, float* sums, float* dgamma, float* dbeta, const float* gamma,
                               const float* mean, const float* invstd, long long count, float* coef, int C, void* stream);
int passl_b200_bn_bwd_apply(const void* y, const void* dz, const void* z, const float* coef, void* dy, void* dres, long long P,
                            int C, int relu, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * LayerNorm over token rows [T, D] bf16 (paddle nn.LayerNorm, passl/models/vision_transformer.py:174,204-205; eps 1e-6).
 * fwd saves mean / rstd (fp32 [T]); bwd writes dx and per-CTA partials [nblk, 2, D] = (dbeta, dgamma) that
 * passl_b200_bn_bwd_finalize sums into the gradient buffers.  D % 8 == 0, D <= 2048.
 * ------------------------------------------------------------------------------------------------------------- */
int passl_b200_layernorm_fwd(const void* x, const float* gamma, const float* beta, void* y, float* mean, float* rstd,
                             long long T, int D, float eps, void* stream);
int passl_b200_layernorm_bwd_blocks(long long T);
/* dx = LN'(dy) + dres   (dres: optional bf16 [T, D] gradient of the residual connection, fused add) */
int passl_b200_layernorm_bwd(const void* x, const void* dy, const float* gamma, const float* mean, const float* rstd,
                             const void* dres, void* dx, float* part, long long T, int D, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * Fused multi-head self-attention on tcgen05 (N <= 256 tokens, head dim 64 or 32): softmax(Q K^T * scale) V per
 * (batch, head) without materialising [B,H,N,N] (passl/models/vision_transformer.py:142-153; causal = CLIP text mask,
 * clip.py:288-290).  qkv: bf16 [B, N, 3, H, d] = the qkv Linear output as is; out: bf16 [B, N, H, d]; lse: fp32 [B, H, N].
 * Backward recomputes the probabilities and writes dqkv in the packed layout.
 * ------------------------------------------------------------------------------------------------------------- */
int passl_b200_attention_fwd(const void* qkv, void* out, float* lse, int B, int N, int H, int d, float scale, int causal,
                             void* stream);
/* delta_ws: caller-provided fp32 [B, H, N] workspace (row sums of dO * O, filled by a pre-pass on the same stream) */
int passl_b200_attention_bwd(const void* qkv, const void* dO, const void* O, const float* lse, void* dqkv, float* delta_ws, int B,
                             int N, int H, int d, float scale, int causal, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * MAE (passl/models/mae.py).
 *   mae_random_masking (:184-212): noise fp32 [B,L] -> ids_shuffle / ids_restore int64 [B,L] (argsort, stable), mask fp32
 *       [B,L] (0 keep / 1 remove).  The Paddle RNG stream is not reproducible, so the noise is an input.
 *   token_assemble_fwd/bwd (:214-266, vision_transformer.py:340-346): mode 0 ViT (cls + patches + pos), mode 1 MAE encoder
 *       input (gather kept patches by ids_shuffle, + pos, prepend cls), mode 2 MAE decoder input (un-shuffle by ids_restore
 *       with mask tokens, + decoder pos).  bwd: ids = ids_restore (mode 1) / ids_shuffle (mode 2); acc_tok / acc_pos are
 *       fp32 gradient accumulators of the cls / mask token and of a learnable positional table (ids_tok = ids_restore).
 *   mae_loss_fwd/bwd (:268-284): masked-patch MSE with optional norm_pix; pred rows may include the cls row
 *       (pred_tokens = L+1, pred_off = 1), imgs fp32 NCHW, mask_sum = sum(mask).
 * ------------------------------------------------------------------------------------------------------------- */
int passl_b200_mae_random_masking(const float* noise, long long* ids_shuffle, long long* ids_restore, float* mask, int B, int L,
                                  int len_keep, void* stream);
int passl_b200_token_assemble_fwd(const void* src, const long long* ids, const float* pos, const float* tok, void* out, int B,
                                  int Ls, int Lo, int D, int mode, int keep, void* stream);
int passl_b200_token_assemble_bwd(const void* dout, const long long* ids, void* dsrc, float* acc_tok, float* acc_pos,
                                  const long long* ids_tok, int B, int Ls, int Lo, int D, int mode, int keep, void* stream);
long long passl_b200_mae_loss_workspace_bytes(int B, int L);
int passl_b200_mae_loss_fwd(const void* pred, const float* imgs, const float* mask, float* loss, int B, int Hp, int P,
                            int pred_tokens, int pred_off, int norm_pix, float mask_sum, void* workspace, void* stream);
int passl_b200_mae_loss_bwd(const void* pred, const float* imgs, const float* mask, const float* dloss, void* dpred, int B,
                            int Hp, int P, int pred_tokens, int pred_off, int norm_pix, float mask_sum, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * CLIP (passl_v110/modeling/backbones/clip.py:299-338, heads/clip_head.py:27-35, architectures/CLIPWrapper.py:45-51).
 *   embedding_fwd   replaces nn.Embedding gather + "+ positional_embedding" (clip.py:300-303): ids int64 [T] (T = B*L),
 *                   table fp32 [V,D], pos fp32 [L,D] -> out bf16 [T,D].   embedding_bwd scatters dout into dtable
 *                   (fp32 L2 reductions) and accumulates dpos; either may be NULL.
 *   eot_gather      replaces text.argmax(-1) + the per-sample Python gather loop (clip.py:307-311): idx int32 [B] (first
 *                   maximum), out[b] = x[b*L + idx[b]].   eot_gather_bwd writes the full dx (zeros elsewhere).
 *   clip_ce         replaces exp(logit_scale), the two logits matmul scalings, both CrossEntropyLoss calls and the
 *                   logit_scale clip (clip.py:316-335, clip_head.py:29-35) on C = I_n T_n^T fp32, n x n valid in an ld x ld buffer (ld % 4 == 0):
 *                   out3 = {img_loss, text_loss, loss}; logit_scale is read on the device and, if clamp != 0, clipped to
 *                   [-4.6, 4.6] after use.  clip_ce_bwd: dC bf16 [ld,ld] = d loss / d C (zero in the padding), *dlogit_scale += d loss / d logit_scale.
 *                   The same workspace must be passed to fwd and bwd (it carries s and the row / column LSE).
 * ------------------------------------------------------------------------------------------------------------- */
int passl_b200_embedding_fwd(const long long* ids, const float* table, const float* pos, void* out, long long T, int L, int D,
                             int V, void* stream);
int passl_b200_embedding_bwd(const long long* ids, const void* dout, float* dtable, float* dpos, long long T, int L, int D, int V,
                             void* stream);
int passl_b200_eot_gather_fwd(const long long* ids, const void* x, void* out, int* idx, int B, int L, int D, void* stream);
int passl_b200_eot_gather_bwd(const int* idx, const void* dout, void* dx, int B, int L, int D, void* stream);
long long passl_b200_clip_ce_workspace_bytes(int ld);
int passl_b200_clip_ce_fwd(const float* C, float* logit_scale, float* out3, int n, int ld, int clamp, void* workspace,
                           long long workspace_bytes, void* stream);
int passl_b200_clip_ce_bwd(const float* C, const float* dloss, void* dC, float* dlogit_scale, int n, int ld, void* workspace,
                           long long workspace_bytes, void* stream);
/* mean softmax cross entropy on materialised fp32 logits [n,m] with int64 labels — nn.CrossEntropyLoss() as the heads call it
 * with explicit logits (clip_head.py:29-32, contrastive_head.py:52-53).  workspace >= 4*n bytes; row_lse [n] feeds the backward:
 * dlogits = dloss/n * (softmax - onehot). */
int passl_b200_rows_ce_fwd(const float* logits, const long long* labels, float* loss, float* row_lse, int n, int m,
                           void* workspace, long long workspace_bytes, void* stream);
int passl_b200_rows_ce_bwd(const float* logits, const long long* labels, const float* row_lse, const float* dloss, float* dlogits,
                           int n, int m, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * Stem / pooling.  im2col: reference NCHW fp32 images -> bf16 [N*Ho*Wo, Kpad] with K order (r, s, c), zero padded
 * (7x7/2 stem conv resnetimagenet.py:190-198; 16x16/16 patch embedding vision_transformer.py:231-236).
 * maxpool 3x3/2 pad 1 (resnetimagenet.py:198), arg-max tap saved as int8; global average pool (base_neck.py:52,79).
 * ------------------------------------------------------------------------------------------------------------- */
int passl_b200_im2col_nchw_f32(const float* x, void* out, int N, int C, int H, int W, int R, int S, int stride, int pad,
                               int Kpad, void* stream);
int passl_b200_maxpool3x3s2_fwd(const void* x, void* y, void* argmax, int N, int H, int W, int C, void* stream);
/* stem tail fused: y = maxpool3x3/2(relu(x*scale + shift)) (BatchNorm apply + ReLU + MaxPool2D of resnetimagenet.py:196-198) */
int passl_b200_bn_relu_maxpool3x3s2_fwd(const void* x, const float* scale, const float* shift, void* y, void* argmax, int N, int H,
                                       int W, int C, void* stream);
int passl_b200_maxpool3x3s2_bwd(const void* dy, const void* argmax, void* dx, int N, int H, int W, int C, void* stream);
int passl_b200_avgpool_fwd(const void* x, void* y_bf16, float* y_f32, int N, int HW, int C, void* stream);
int passl_b200_avgpool_bwd(const void* dy, void* dx, int N, int HW, int C, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * Fused optimizer steps over flat fp32 buffers (+ bf16 compute copy).  Replace the per-parameter Python loops of
 * passl/optimizer/momentum.py:60-158, momentum_lars.py:56-114, adamw.py:52-138.  Tensors start at multiples of 1024
 * elements inside the flat buffer; block_seg[b] = tensor id of 1024-element block b.
 * `ctrl` (may be NULL): the device-side control word {multiplier, found_inf, global_norm} written by
 * passl_b200_grad_norm_finite — the gradient is multiplied by ctrl[0] on the fly and the whole step is skipped when ctrl[1] != 0,
 * so gradient clipping / unscaling costs one read pass over the gradients and no host round trip.
 * ------------------------------------------------------------------------------------------------------------- */
int passl_b200_sgd_momentum(float* p, const float* g, float* v, void* p_bf16, float lr, float momentum, float wd,
                            float grad_scale, const float* ctrl, long long n, void* stream);
int passl_b200_lars_momentum(float* p, const float* g, float* v, void* p_bf16, const int* block_seg, const float* seg_wd,
                             float* norms, int num_segments, float lr, float momentum, float lars_coeff, float eps,
                             float grad_scale, const float* ctrl, long long n, void* stream);
int passl_b200_adamw(float* p, const float* g, float* m, float* v, void* p_bf16, const int* block_seg, const float* seg_wd,
                     const float* seg_lr_ratio, float lr, float beta1, float beta2, float eps, int step, float grad_scale,
                     const float* ctrl, long long n, void* stream);
/* ClipGradByGlobalNorm (passl/core/grad_clip.py:30-84) + check_finite_and_unscale (passl/core/grad_scaler.py:48-87) as ONE read
 * pass over the flat fp32 gradient buffer: ctrl fp32[3] = {unscale * clip_coef (0 when a gradient is non-finite), found_inf,
 * global norm of unscale*g}; clip_coef = 1 if (!always_clip && norm <= clip_norm) else min(clip_norm / (norm + 1e-6), coef_max);
 * clip_norm <= 0 disables clipping, coef_max <= 0 disables the cap.  scratch fp32[3]: zeroed once by the caller, left zeroed. */
int passl_b200_grad_norm_finite(const float* g, long long n, float unscale, float clip_norm, float coef_max, int always_clip,
                                float* ctrl, float* scratch, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * Embedding exchange over NVLink peer memory (replaces the NCCL all_gather / reduce_scatter behind
 * moco.py:198-210 and passl/distributed/nn/functional.py:100-127 when the ranks share a node): one kernel that signals through
 * flags in peer memory and reads every rank's shard with P2P loads.  data_ptrs / flag_ptrs are HOST arrays of `world` device
 * pointers = this process's mappings (peer_buffer_open) of each rank's data buffer and flag row (uint32[world], zeroed once); `epoch` increases by one per exchange on a slot; use two slots alternately (see passl_b200/distributed/peer.py).
 *   peer_allgather          out[q*shard_bytes ...] = rank q's shard
 *   peer_reduce_scatter_f32 out[i] = sum_q (rank q's buffer)[rank*shard_elems + i]
 * ------------------------------------------------------------------------------------------------------------- */
/* exchange buffers: the only device memory the library allocates itself (a CUDA IPC handle needs the base of a cudaMalloc
 * allocation); create exports a 64-byte handle, open maps a peer's buffer with the current device as accessor */
int passl_b200_peer_buffer_create(long long bytes, void** base, unsigned char* handle64);
int passl_b200_peer_buffer_open(const unsigned char* handle64, void** mapped);
int passl_b200_peer_buffer_close(void* mapped);
int passl_b200_peer_buffer_destroy(void* base);
/* `shard` / `grad_all` (local device memory) is first copied into this rank's slot, then the kernel signals and gathers */
/* publish this rank's key shard for passl_b200_infonce_tc_fwd_peer: keys fp32 [n, D] -> bf16 into data_ptrs[rank], then
 * (system-scope fence) flag[rank] = epoch in every rank's flag row.  done: uint32, zeroed once by the caller. */
int passl_b200_peer_publish_keys_bf16(const float* keys, int n, int D, const void* const* data_ptrs, void* const* flag_ptrs, int rank,
                                      int world, unsigned epoch, void* done, void* stream);
int passl_b200_peer_allgather(const void* shard, const void* const* data_ptrs, void* const* flag_ptrs, void* out, long long shard_bytes,
                              int rank, int world, unsigned epoch, void* stream);
int passl_b200_peer_reduce_scatter_f32(const float* grad_all, const void* const* data_ptrs, void* const* flag_ptrs, float* out,
                                       long long shard_elems, int rank, int world, unsigned epoch, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * Image input stage: the per-view image ops right before the backbone, on decoded uint8 HWC RGB images that are
 * already in device memory.  `src` packs the images back to back; image n starts at byte src_off[n] and is src_h[n] x src_w[n].
 * An item m is one output view: item_img[m] names its source image, item_box[4m..4m+3] = (top, left, crop_h, crop_w) is the box the
 * host drew (RandomResizedCrop.get_params, passl_v110/datasets/preprocess/transforms.py:517-557).  All index arrays are device int32.
 *   resized_crop_u8     RandomResizedCrop's image op = PIL crop + Image.resize((S, S), BILINEAR | BICUBIC) (configs/simclr/
 *                       simclr_r50_IM.yaml:35-39): Pillow's two-pass 8-bit resampler (libImaging/Resample.c) with its fixed-point
 *                       taps, bit-exact.  dst uint8 [items, S, S, 3].  interpolation 0 bilinear, 1 bicubic.  kmax >=
 *                       resample_kmax(largest crop side, S, interpolation); max_crop_h >= the largest crop_h.  After the call
 *                       the first int of the workspace is a status word: bit 0 = a box outside its image, bit 1 = kmax too small
 *                       (the affected items are written as zeros).
 *   views_finalize_f32  RandomGrayscale (img.convert('L') replicated, transforms.py:150-170) where gray[m], RandomHorizontalFlip
 *                       where flip[m], Transpose HWC->CHW and NormalizeImage (x * scale - mean) / std evaluated in double and
 *                       stored as float32 (transforms.py:462-467).  out fp32 [items, 3, S, S]; mean3 / std3 are HOST float[3].
 * ------------------------------------------------------------------------------------------------------------- */
int passl_b200_resample_kmax(int max_crop, int out_size, int interpolation);
long long passl_b200_resized_crop_workspace_bytes(int items, int out_size, int max_crop_h, int kmax);
int passl_b200_resized_crop_u8(const void* src, const long long* src_off, const int* src_h, const int* src_w, const int* item_img,
                               const int* item_box, void* dst, void* workspace, long long workspace_bytes, int items, int out_size,
                               int max_crop_h, int kmax, int interpolation, void* stream);
int passl_b200_views_finalize_f32(const void* img, const int* gray, const int* flip, float* out, int items, int size, double scale,
                                  const float* mean3, const float* std3, void* stream);
/*   color_jitter_u8     ColorJitter (configs/simclr/simclr_r50_IM.yaml:41-48; paddle.vision over Pillow): per view up to four ops
 *                       in the order the host drew, in place on uint8 [items, S, S, 3].  ops int32 [items][4]: 0 none, 1 brightness,
 *                       2 contrast, 3 saturation (ImageEnhance blends, libImaging/Blend.c), 4 hue (HSV round trip of Convert.c with
 *                       the H plane shifted), 5 grayscale (convert('L') replicated; for views that are blurred afterwards).
 *                       factors fp32 [items][4]: the blend factor, or for hue the shift uint8(hue_factor * 255) the host computed.
 *                       workspace >= 8 * items bytes (exact luma sums for the contrast mean); contrast_positions: bit p set when
 *                       some view has contrast at position p (0xF is always safe). */
int passl_b200_color_jitter_u8(void* img, const int* ops, const float* factors, void* workspace, long long workspace_bytes, int items,
                               int size, int contrast_positions, void* stream);
/*   gaussian_blur_u8    GaussianBlur (transforms.py:173-191: cv2.GaussianBlur(x, (23, 23), sigma) on uint8): OpenCV's fixed-point
 *                       filter, in place on uint8 [items, S, S, 3].  taps int32 [items][ksize]: the 8.8 fixed-point Gaussian row of
 *                       each view (sum 256; computed on the host from its sigma, passl_b200.data.gaussian_taps_fixed); apply int32
 *                       [items]: 0 leaves the view untouched.  Borders BORDER_REFLECT_101, one rounding at the end
 *                       ((sum + 2^15) >> 16).  workspace >= gaussian_blur_workspace_bytes (16-bit intermediate). */
long long passl_b200_gaussian_blur_workspace_bytes(int items, int size);
int passl_b200_gaussian_blur_u8(void* img, const int* taps, const int* apply, void* workspace, long long workspace_bytes, int items,
                                int size, int ksize, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * Developer probe (not a reference entry point): one tcgen05.mma over a row-shifted view of a SWIZZLE_128B tile, used by
 * tests/test_umma_probe_gpu.py to pin the shared-memory descriptor semantics (start address not 1024-aligned, SBO != 1024,
 * base_offset bits) that the halo-tile convolution kernels rely on.  A bf16 [256,64], B bf16 [64,64], out fp32 [128,64].
 * ------------------------------------------------------------------------------------------------------------- */
int passl_b200_umma_probe(const void* A, const void* B, float* out, int shift_rows, int sbo_bytes, int base_offset, int a_mn,
                          void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PASSL_B200_H_ */
