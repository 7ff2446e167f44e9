/* graphecho_b200 — C-ABI of the sm_100a kernels behind GraphEcho's data-parallel hot path.
 *
 * The reference (xmed-lab/GraphEcho) is pure Python/PyTorch and has no FFI of its own: its
 * boundary for this path is the nn.Module API (models.fpnseg / models.vig /
 * models.graph_matching / models.affinity_layer / models.TGCN / utils.sinkhorn_distance).
 * Each entry point below names the reference op sequence it replaces (file:line under the
 * reference tree); graphecho_b200/_cabi.py binds them with ctypes and
 * graphecho_b200/functional.py wraps them in torch.autograd.Function.
 *
 * Conventions
 *  - plain pointers + sizes; every pointer is DEVICE memory owned by the caller
 *    (inputs, outputs and workspace); the library never allocates, frees or keeps them.
 *  - tensors are dense in the documented layout; float = fp32.
 *  - launches are asynchronous on `stream` (a cudaStream_t); no internal sync,
 *    CUDA-graph capturable.
 *  - return 0 on success, <0 argument / shape / capacity error, >0 a cudaError_t;
 *    ge_last_error() gives the text (thread-local).  Nothing falls back to the CPU.
 */
#ifndef GRAPHECHO_B200_H
#define GRAPHECHO_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GE_ABI_VERSION 1

typedef void* ge_stream_t; /* cudaStream_t */

/* activation storage types for the NHWC feature-map kernels (math is always fp32) */
#define GE_DTYPE_F32 0
#define GE_DTYPE_BF16 1

int ge_version(void);
const char* ge_last_error(void);
int ge_device_sm_count(void);
unsigned long long ge_launch_count(void);

/* ---- K3: pairwise affinity --------------------------------------------------------------
 * M[b,i,j] = sum_k w2[k]*relu(A[b,i,k] + B[b,j,k]) + b2[0]
 * Separable form of Affinity.forward (models/affinity_layer.py:52-73): A = fc_M.0 applied to
 * the project_sr half, B = the project_tg half + bias (dense projections done by the caller).
 * A [batch,N1,H], B [batch,N2,H], w2 [H], b2 [1], M [batch,N1,N2].  H % 32 == 0, H <= 640. */
int ge_affinity_pairwise_fwd(const float* A, const float* B, const float* w2, const float* b2,
                             float* M, int batch, int N1, int N2, int H, ge_stream_t stream);
size_t ge_affinity_pairwise_bwd_workspace_bytes(int batch, int N1, int N2, int H);
/* dA [batch,N1,H], dB [batch,N2,H], dw2 [H], db2 [1] (all overwritten). */
int ge_affinity_pairwise_bwd(const float* A, const float* B, const float* w2, const float* dM,
                             float* dA, float* dB, float* dw2, float* db2,
                             void* workspace, size_t workspace_bytes,
                             int batch, int N1, int N2, int H, ge_stream_t stream);

/* ---- K4: instance-norm + slack Sinkhorn + exp --------------------------------------------
 * P = exp(sinkhorn_rpm(InstanceNorm2d(1)(M), n_iters, slack=True))
 * (models/graph_matching.py:574-575 and 637-676).  One thread-block cluster per problem, matrix on chip for the
 * whole loop.  Two kernels behind one entry point:
 *  - register-resident, exponent domain (K = exp(z) in a per-thread register tile, u = exp(-r), v = exp(-c); two
 *    matrix-vector products per iteration, no exp/log in the loop) for N2 <= 256, N1 <= 512;
 *  - log domain with the matrix in (distributed) shared memory for everything else, for cluster_size > 0, and --
 *    gated per problem on stats[2] -- for inputs whose max z (> 80) would overflow the exponent domain.
 * M, P [batch,N1,N2]; hist_r [batch,n_iters,N1], hist_c [batch,n_iters,N2] receive the row / column potentials after
 * every pass, saved for the backward (u_t, v_t when stats[2] == 0, log-potentials r_t, c_t when stats[2] == 1);
 * stats [batch,4] receives (mean, rstd, path, -).  apply_instnorm=0 skips the normalisation (plain sinkhorn_rpm).
 * cluster_size: 0 = choose, else 1/2/4/8/16 (log-domain kernel).  GE_ERR_CAPACITY if N1*N2 does not fit.
 * ge_sinkhorn_rpm_set_path: 0 automatic, 1 log-domain only, 2 / 3 register path with 4 / 8 rows per thread (tuning). */
int ge_sinkhorn_rpm_set_path(int path);
int ge_sinkhorn_rpm_cluster_size(int N1, int N2, int backward);
int ge_sinkhorn_rpm_fwd(const float* M, float* P, float* hist_r, float* hist_c, float* stats,
                        int batch, int N1, int N2, int n_iters, int apply_instnorm,
                        int cluster_size, ge_stream_t stream);
/* G = dLoss/dP; dM receives dLoss/dM (exact adjoint of the unrolled iterations). */
int ge_sinkhorn_rpm_bwd(const float* M, const float* G, const float* hist_r, const float* hist_c,
                        const float* stats, float* dM, int batch, int N1, int N2, int n_iters,
                        int apply_instnorm, int cluster_size, ge_stream_t stream);

/* ---- K4b: 'o2o' matching loss on the Sinkhorn-normalised affinity --------------------------
 * GModule._forward_aff (models/graph_matching.py:572-590) with BCEFocalLoss (:23-45):
 *   idx_i = argmax_j P_ij [lab1_i == lab2_j], tp_i = P[i,idx_i], tp_loss = mean_i(-alpha (1-tp_i)^gamma log tp_i) / N1,
 *   fp_loss = mean over different-class entries of -(1-alpha) P^gamma log(1-P), divided by their (detached) sum of P;
 *   loss = tp_loss + fp_loss.  One launch forward (one CTA walks the matrix), one element-wise launch backward.
 * P [N1,N2] fp32; lab1 [N1], lab2 [N2] fp32 class labels; loss [1]; idx int32 [N1] and stats fp32 [4] are saved for the
 * backward; gout [1] = dLoss/dloss on the device; dP [N1,N2]. */
int ge_matching_loss_fwd(const float* P, const float* lab1, const float* lab2, float* loss, int* idx, float* stats,
                         int N1, int N2, float alpha, float gamma, ge_stream_t stream);
int ge_matching_loss_bwd(const float* P, const float* lab1, const float* lab2, const int* idx, const float* stats,
                         const float* gout, float* dP, int N1, int N2, float alpha, float gamma, ge_stream_t stream);

/* ---- stem convolution --------------------------------------------------------------------
 * ResNet.conv1 = Conv2d(1, 64, 7, stride 2, padding 3, bias=False) on gray frames (models/fpnseg.py:229, 251), forward and
 * weight gradient as direct FP32-pipe kernels (cuDNN has no tensor-core path for one input channel: 0.46 + 0.64 ms of
 * conversions and sm80 kernels per step).  x fp32 [F,1,H,W]; w / dw fp32 [64,1,7,7]; y / dy [F,H/2,W/2,64] NHWC in `dtype`
 * (bf16: operands rounded to bf16, fp32 accumulation, as the library's bf16 convolution).  H even, W % 8 == 0, W <= 512.
 * The input receives no gradient (it is the data). */
int ge_stem_conv_supported(int H, int W);
int ge_stem_conv_fwd(const float* x, const float* w, void* y, int F, int H, int W, int dtype, ge_stream_t stream);
size_t ge_stem_conv_wgrad_workspace_bytes(int F, int H, int W);
int ge_stem_conv_wgrad(const float* x, const void* dy, float* dw, void* workspace, size_t workspace_bytes,
                       int F, int H, int W, int dtype, ge_stream_t stream);

/* ---- K5: SinkhornDistance ----------------------------------------------------------------
 * utils/sinkhorn_distance.py:27-86: C_ij = sum_d (x_id-y_jd)^2, <= max_iter log-domain updates
 * with the batch-mean early stop `err < thresh`, pi = exp((-C+u+v)/eps), cost_b = sum pi*C.
 * x [B,P1,D], y [B,P2,D]; outputs C, pi [B,P1,P2], cost [B]; saved for backward:
 * hist_u [B,max_iter,P1], hist_v [B,max_iter,P2], err [B,max_iter], nits [1] (int32, the
 * number of iterations the reference loop would have executed).  One CTA per batch element; when the P1 x P2
 * matrices do not fit its shared memory (node sets of 200-320 rows) they live in their global output arrays. */
int ge_sinkhorn_distance_fwd(const float* x, const float* y, float* C, float* pi, float* cost,
                             float* hist_u, float* hist_v, float* err, int* nits,
                             int B, int P1, int P2, int D, float eps, int max_iter, double thresh,
                             ge_stream_t stream);
/* gcost [B] = dLoss/dcost_b; dC [B,P1,P2] scratch/output; dx [B,P1,D], dy [B,P2,D] -- or both NULL: only dC is
 * produced and the caller forms dx = 2(rowsum(dC) x - dC y), dy = 2(colsum(dC) y - dC^T x) itself (GEMM-shaped). */
int ge_sinkhorn_distance_bwd(const float* x, const float* y, const float* C, const float* hist_u,
                             const float* hist_v, const int* nits, const float* gcost,
                             float* dC, float* dx, float* dy,
                             int B, int P1, int P2, int D, float eps, int max_iter, ge_stream_t stream);

/* ---- K1: dense dilated k-NN graph ----------------------------------------------------------
 * DenseDilatedKnnGraph.forward (models/vig.py:369-381) incl. F.normalize, (xy_)pairwise_distance
 * (:232-274), topk (:306/327), centre index + stack (:308-309/328-329) and the dilation stride
 * (:353).  x [B,C,N], y [B,C,M] or NULL (self graph, M == N), relative_pos [N,M] or NULL,
 * edge_index int64 [2,B,N,k]; k*dilation <= 64 and <= M.  Ties -> lower key index.
 * Two implementations behind the one entry point: the inner-product GEMM on tcgen05.mma
 * (3xTF32 split = fp32-level accuracy, TMA-staged operands, accumulator in TMEM, per-thread
 * top-k; used when relative_pos == NULL, C % 32 == 0, k*dilation <= 32 and N, M >= 128) and
 * fp32 FFMA kernels for everything else.  ge_knn_graph_set_path: 0 = choose (default),
 * 1 = FFMA kernels only, 2 = require the tcgen05 kernel (GE_ERR_SHAPE if it does not apply). */
int ge_knn_graph_set_path(int path);
size_t ge_knn_graph_workspace_bytes(int B, int C, int N, int M);
int ge_knn_graph(const float* x, const float* y, const float* relative_pos, long long* edge_index,
                 void* workspace, size_t workspace_bytes,
                 int B, int C, int N, int M, int k, int dilation, ge_stream_t stream);

/* Node-major (channels-last) entry for the Grapher: x [B,N,C], y [B,M,C] or NULL in `dtype` (fp32 / bf16) -- the
 * layout the FPN feature maps already have.  tcgen05 path only: ge_knn_graph_nmajor_supported() says whether
 * it covers the problem (else transpose and call ge_knn_graph).  Same workspace size as ge_knn_graph. */
int ge_knn_graph_nmajor_supported(int B, int C, int N, int M, int k, int dilation);
int ge_knn_graph_nmajor(const void* x, const void* y, int dtype, long long* edge_index,
                        void* workspace, size_t workspace_bytes,
                        int B, int C, int N, int M, int k, int dilation, ge_stream_t stream);

/* ---- K2: max-relative aggregation (gather half of MRConv2d) -------------------------------
 * models/vig.py:96-104 + batched_index_select (:209-229):
 * out[b,2c,n] = x[b,c,n]; out[b,2c+1,n] = max_k(y[b,c,idx_nbr[b,n,k]] - x[b,c,idx_ctr[b,n,k]]).
 * x [B,C,N], y [B,C,M] or NULL (= x), idx_* int64 [B,N,k] (idx_ctr NULL = the point itself),
 * out [B,2C,N], argk uint8 [B,C,N] (winning neighbour slot, saved for the backward). */
int ge_mrconv_gather_fwd(const float* x, const float* y, const long long* idx_nbr, const long long* idx_ctr,
                         float* out, unsigned char* argk, int B, int C, int N, int M, int k, ge_stream_t stream);
/* dx [B,C,N] overwritten; dy [B,C,M] must be zero-filled by the caller (NULL for a self graph:
 * neighbour gradients are accumulated into dx). */
int ge_mrconv_gather_bwd(const float* dout, const long long* idx_nbr, const long long* idx_ctr,
                         const unsigned char* argk, float* dx, float* dy,
                         int B, int C, int N, int M, int k, ge_stream_t stream);
/* Node-major variants (centre of every edge = the point itself): x [B,N,C], y [B,M,C] or NULL, out [B,N,2C] in
 * `dtype` with out[..,2c] = x[..,c], out[..,2c+1] = max_k(y[nbr_k,c] - x[n,c]); argk uint8 [B,N,C].
 * Backward: dx [B,N,C] fp32 overwritten, dy [B,M,C] fp32 zero-filled by the caller (NULL for a self-graph). */
int ge_mrconv_gather_nmajor_fwd(const void* x, const void* y, const long long* idx_nbr, void* out,
                                unsigned char* argk, int dtype, int B, int C, int N, int M, int k,
                                ge_stream_t stream);
int ge_mrconv_gather_nmajor_bwd(const void* dout, const long long* idx_nbr, const unsigned char* argk,
                                float* dx, float* dy, int dtype, int B, int C, int N, int M, int k,
                                ge_stream_t stream);
/* Self-graph (y == NULL) backward with the gradient slab [N][64 channels] resident in shared memory: one launch, dx
 * [B,N,C] written once in `dtype`.  GE_ERR_CAPACITY when N * 256 + N * k * 2 bytes exceed 220 KB (use the entry above). */
int ge_mrconv_gather_nmajor_bwd_self(const void* dout, const long long* idx_nbr, const unsigned char* argk,
                                     void* dx, int dtype, int B, int C, int N, int k, ge_stream_t stream);

/* ---- K6: TGCN pyramid pooling + concat ------------------------------------------------------
 * avg_pool2d(r) per level + channel concat of TGCN.DyGraphConv2d.forward (models/TGCN.py:62-70),
 * hoisted out of the time loop.  `in`: NHWC-dense frames [H,W,C] `frame_stride` elements apart;
 * out fp32 [frames, H/r, W/r, Ctot] (channels_last), this level fills channels [coff, coff+C). */
int ge_tgcn_pool_concat_fwd(const void* in, long long frame_stride, float* out, int dtype,
                            long long frames, int H, int W, int C, int r, int Ctot, int coff, ge_stream_t stream);
int ge_tgcn_pool_concat_bwd(const float* dout, void* din, int dtype,
                            long long frames, int H, int W, int C, int r, int Ctot, int coff, ge_stream_t stream);

/* ---- K7: FPN top-down / semantic head glue (NHWC feature maps, dtype = GE_DTYPE_*) ---------
 * (i) out = bilinear_up(top -> HxW, align_corners=True) + lateral   (models/fpnseg.py:371-388);
 *     lateral may be NULL (plain _upsample, :358-359).  top [N,h,w,C], lateral/out [N,H,W,C]. */
int ge_upsample_add_fwd(const void* top, const void* lateral, void* out, int dtype,
                        int N, int h, int w, int H, int W, int C, ge_stream_t stream);
/* adjoint of the up-sampling: dtop [N,h,w,C] from dout [N,H,W,C] (d lateral = dout). */
int ge_upsample_bwd(const void* dout, void* dtop, int dtype,
                    int N, int h, int w, int H, int W, int C, ge_stream_t stream);
/* (ii) GroupNorm + ReLU (+ _upsample): the semantic head's GroupNorm(C,C) (fpnseg.py:354-355, 428-442)
 *      and the Discriminator towers' GroupNorm(32,256)+ReLU (fpnseg.py:455-466).
 *      Group statistics over HW x channels_per_group, written per channel: mean, rstd fp32 [N,C]. */
int ge_group_stats(const void* x, float* mean, float* rstd, int dtype,
                   int N, int HW, int C, int channels_per_group, float eps, ge_stream_t stream);
/* Same statistics for x + pre_bias[c] (the bias of the convolution that produced x, folded into the norm so that
 * neither the add nor its gradient reduction is a separate pass).  mean receives the EFFECTIVE mean
 * (group mean - pre_bias[c]): ge_gn_relu_upsample_fwd/bwd then run unchanged on the raw x.  chan_sum [N,C] (or NULL)
 * receives sum_hw x[n,:,c]; the gradient of pre_bias follows from it and the backward's S/A arrays. */
int ge_group_stats_bias(const void* x, const float* pre_bias, float* mean, float* rstd, float* chan_sum,
                        int dtype, int N, int HW, int C, int channels_per_group, float eps,
                        ge_stream_t stream);
/* out [N,H,W,C] = bilinear_up(relu((x-mean)*rstd*gamma+beta)); x [N,h,w,C]; (h,w)==(H,W) = no up-sampling. */
int ge_gn_relu_upsample_fwd(const void* x, const float* mean, const float* rstd,
                            const float* gamma, const float* beta, void* out, int dtype,
                            int N, int h, int w, int H, int W, int C, ge_stream_t stream);
/* dx [N,h,w,C]; dyh fp32 scratch [N,h,w,C] (only for an up-sampling call, else may be NULL);
 * S1,S2 fp32 [N,C]: dbeta = sum_n S1, dgamma = sum_n S2; A1,A2 fp32 [N,C] scratch. */
int ge_gn_relu_upsample_bwd(const void* dout, const void* x, const float* mean, const float* rstd,
                            const float* gamma, const float* beta, float* dyh, float* S1, float* S2,
                            float* A1, float* A2, void* dx, int dtype, int N, int h, int w, int H, int W,
                            int C, int channels_per_group, ge_stream_t stream);
/* (iii) logits = bilinear_up_x4(conv3(s2+s3+s4+s5))  (fpnseg.py:444).  s* [N,h,w,C]; W3 [nc,C],
 *       b3 [nc]; q fp32 [N,h,w,nc] (conv3 output before up-sampling); logits fp32 NCHW [N,nc,H,W]. */
int ge_seg_tail_fwd(const void* s2, const void* s3, const void* s4, const void* s5,
                    const float* W3, const float* b3, float* q, float* logits, int dtype,
                    int N, int h, int w, int H, int W, int C, int nc, ge_stream_t stream);
/* ds [N,h,w,C] is the gradient of every branch; dW3 [nc,C], db3 [nc] must be zero-filled. */
int ge_seg_tail_bwd(const float* dlogits, const void* s2, const void* s3, const void* s4, const void* s5,
                    const float* W3, float* dq, void* ds, float* dW3, float* db3, int dtype,
                    int N, int h, int w, int H, int W, int C, int nc, ge_stream_t stream);

/* ---- fused BatchNorm2d (+ residual add) (+ ReLU) on NHWC maps ---------------------------------
 * nn.BatchNorm2d + `out += identity` + nn.ReLU of Bottleneck.forward / the ResNet stem / the VGG16
 * blocks (models/fpnseg.py:192-212, 251-255, 27-142).  x, residual (or NULL), out: [P,C] with
 * P = N*H*W, in `dtype`; gamma, beta, running_* fp32 [C].  Training mode uses batch statistics
 * (biased variance), updates running_* with `momentum` (unbiased variance) and num_batches_tracked exactly as
 * nn.BatchNorm2d, and saves mean / rstd for the backward.  Two-stage deterministic reductions through `workspace`.
 * P_split: 0 < P_split < P splits the batch into two SEGMENTS [0,P_split) and [P_split,P) with their own batch
 * statistics and one running-stat update each, in that order -- the semantics of the reference trainer's two
 * separate network calls on the source and the target batch (train_cardiac_uda.py:225, 234) for a batch that
 * holds both.  save_mean / save_rstd are then [2][C] (else [1][C]).
 * relu_mask (ge_bn_relu_mask_bytes(P,C) bytes; may be NULL when relu == 0): 1-bit-per-element ReLU mask written by
 * the forward and consumed by the backward, which therefore never re-reads `out`. */
size_t ge_bn_workspace_bytes(long long P, int C);
size_t ge_bn_relu_mask_bytes(long long P, int C);
/* 0 / 1 (default): the three-kernel streaming path.  2: maps that fit the chip's shared memory (<= ~28 MB; for the
 * backward x and dy, or x alone) take a single-launch cooperative kernel -- x read from HBM once, statistics /
 * finalize / apply separated by grid barriers.  Measured no faster alone and slower inside the training step (a
 * cooperative grid cannot overlap other streams), hence opt-in. */
int ge_bn_set_path(int path);
int ge_bn_fwd_train(const void* x, const void* residual, const float* gamma, const float* beta,
                    float* running_mean, float* running_var, long long* num_batches_tracked,
                    float momentum, float eps, void* out, float* save_mean, float* save_rstd,
                    void* relu_mask, void* workspace, size_t workspace_bytes,
                    int dtype, long long P, long long P_split, int C, int relu, ge_stream_t stream);
/* The same forward for an x whose per-CTA partial statistics already exist (the epilogue of ge_conv1x1_bn_stats):
 * finalize + apply only.  part fp32 [rows][2][C] = sums of (x - shift), (x - shift)^2 with shift = running_mean as it
 * is before this call (0 when NULL); rows [0, rows_segment0) belong to pixels [0, P_split). */
int ge_bn_fwd_train_prestat(const void* x, const void* residual, const float* gamma, const float* beta,
                            float* running_mean, float* running_var, long long* num_batches_tracked,
                            float momentum, float eps, void* out, float* save_mean, float* save_rstd,
                            void* relu_mask, const float* part, int rows, int rows_segment0,
                            int dtype, long long P, long long P_split, int C, int relu, ge_stream_t stream);
int ge_bn_fwd_eval(const void* x, const void* residual, const float* gamma, const float* beta,
                   const float* running_mean, const float* running_var, float eps, void* out,
                   int dtype, long long P, int C, int relu, ge_stream_t stream);
/* dx [P,C]; dres [P,C] or NULL (gradient of the residual input); dgamma, dbeta fp32 [C]. */
int ge_bn_bwd(const void* dy, const void* relu_mask, const void* x, const float* gamma,
              const float* mean, const float* rstd, void* dx, void* dres,
              float* dgamma, float* dbeta, void* workspace, size_t workspace_bytes,
              int dtype, long long P, long long P_split, int C, int relu, ge_stream_t stream);

/* SyncBatchNorm (the reference's intent under DDP, train_cardiac_uda.py:142): the same kernels, split around the two
 * cross-rank exchanges.  The caller all-reduces `sums` / `seg_sums` with op = AVG over `world` equally loaded ranks
 * between stage 1 and stage 2 of each direction (NCCL on the same stream).
 *   ge_bn_sync_stats      : sums fp32 [nseg][2][C] = per-segment sums of (x - shift), (x - shift)^2 over this rank's
 *                           pixels; shift = running_mean (identical on all ranks) or NULL.
 *   ge_bn_sync_fwd_apply  : as ge_bn_fwd_train from the averaged sums (running variance unbiased over P_seg * world).
 *   ge_bn_sync_bwd_reduce : seg_sums fp32 [2][nseg][C] = per-segment sums of dy*relu', dy*relu'*xhat; dgamma, dbeta of
 *                           this rank (the gradient exchange averages them like any other parameter gradient).
 *   ge_bn_sync_bwd_apply  : dx (+ dres) from the averaged seg_sums. */
int ge_bn_sync_stats(const void* x, const float* shift, float* sums, void* workspace, size_t workspace_bytes,
                     int dtype, long long P, long long P_split, int C, ge_stream_t stream);
int ge_bn_sync_fwd_apply(const void* x, const void* residual, const float* gamma, const float* beta,
                         float* running_mean, float* running_var, long long* num_batches_tracked,
                         float momentum, float eps, void* out, float* save_mean, float* save_rstd,
                         void* relu_mask, const float* sums_avg, int world,
                         int dtype, long long P, long long P_split, int C, int relu, ge_stream_t stream);
int ge_bn_sync_bwd_reduce(const void* dy, const void* relu_mask, const void* x, const float* mean,
                          const float* rstd, float* seg_sums, float* dgamma, float* dbeta,
                          void* workspace, size_t workspace_bytes,
                          int dtype, long long P, long long P_split, int C, int relu, ge_stream_t stream);
int ge_bn_sync_bwd_apply(const void* dy, const void* relu_mask, const void* x, const float* gamma,
                         const float* mean, const float* rstd, const float* seg_sums_avg, void* dx, void* dres,
                         int dtype, long long P, long long P_split, int C, int relu, ge_stream_t stream);

/* ---- update_seed: spectral bipartition ------------------------------------------------------
 * What GModule.update_seed asks sklearn's SpectralClustering(2, affinity='nearest_neighbors',
 * n_neighbors, assign_labels='kmeans') for (models/graph_matching.py:532-567), on the device:
 * pts fp32 [n,d] with row 0 = the class seed; keep uint8 [n-1] = 1 where point i+1 lands in the
 * seed's cluster.  One CTA, n <= ge_spectral_bipartition_max_points(). */
int ge_spectral_bipartition_max_points(void);
int ge_spectral_bipartition(const float* pts, unsigned char* keep, int n, int d, int n_neighbors,
                            int iterations, ge_stream_t stream);

/* ---- stem max pooling -----------------------------------------------------------------------
 * nn.MaxPool2d(3, 2, 1) of the ResNet stem (models/fpnseg.py:232, 254) on an NHWC map in `dtype`:
 * x [N,H,W,C] -> out [N,Ho,Wo,C], Ho = (H-1)/2+1; arg uint8 [N,Ho,Wo,C] = selected window tap (first maximum).
 * Backward is a deterministic gather (no atomics): dx [N,H,W,C] overwritten. */
int ge_maxpool3s2_fwd(const void* x, void* out, unsigned char* arg, int dtype,
                      int N, int H, int W, int C, ge_stream_t stream);
int ge_maxpool3s2_bwd(const void* dout, const unsigned char* arg, void* dx, int dtype,
                      int N, int H, int W, int C, ge_stream_t stream);

/* ---- f3 (first slice): 1x1 convolution as a tcgen05 GEMM with the BatchNorm statistics in its epilogue -------------
 * Bottleneck conv1/conv3 + BatchNorm (models/fpnseg.py:192-212), Grapher fc1 (models/vig.py:402-405) on NHWC bf16 maps:
 * y [P,N] = x [P,K] W[N,K]^T (fp32 accumulate in TMEM, TMA-staged operands, weight tile resident in shared memory,
 * persistent CTAs).  part (or NULL) receives [rows][2][N] fp32 partial sums of (y - shift[n]) and (y - shift[n])^2 taken
 * from the fp32 accumulators -- the input of ge_bn_fwd_train_prestat; rows = ge_conv1x1_tc_partial_rows(...).
 * Supported: K % 64 == 0, N % 64 == 0, min(N,256) * K * 2 <= 128 KB (or a narrower column tile that fits). */
int ge_conv1x1_tc_supported(long long P, int K, int N);
int ge_conv1x1_tc_partial_rows(long long P, long long P_split, int K, int N, int* rows_segment0);
int ge_conv1x1_bn_stats(const void* x, const void* w, void* y, const float* shift, float* part,
                        long long P, long long P_split, int K, int N, ge_stream_t stream);

/* ---- K6: the TGCN recurrence as one persistent launch ---------------------------------------------
 * TGCN.forward's time loop (models/TGCN.py:224-235) over DyGraphConv2d.forward (:62-78) AFTER the state-independent
 * part (pooling, MLP, position embedding) has been hoisted out: per step
 *   hidden_t = GELU(Conv1x1_{groups=4}(interleave[x_t ; max_k(hidden_{t-1}[:, nn_k] - x_t)]) + bias),
 * nn_k = the k nearest hidden_{t-1} nodes of every x_t node on channel-normalised vectors (vig.py:369-381), hidden_0 = 0.
 * One CTA per clip keeps hidden / x_t / the max-relative features in shared memory for all T steps.
 * emb fp32 [B,T,C,N]; Wt fp32 [2C/4, C] = TRANSPOSED conv weight; W fp32 [C, 2C/4]; bias fp32 [C].
 * hidden_all, z_all fp32 [B,T,C,N] (hidden_all[:,T-1] = result; z = pre-activation), idx_all int32 [B,T,N,k],
 * argk_all uint8 [B,T,C,N] are written by the forward and read by the backward.  Backward: dH fp32 [B,C,N] ->
 * d_emb fp32 [B,T,C,N], dWt_part fp32 [B,2C/4,C] and db_part fp32 [B,C] (ZERO-FILLED by the caller, summed over B by the
 * caller), scratch fp32 [B,C,N].  Supported: C = 256, N = 64, k <= 16, dilation 1, groups 4 (ge_tgcn_recurrence_supported). */
int ge_tgcn_recurrence_supported(int C, int Cout, int N, int k, int dilation, int groups);
int ge_tgcn_recurrence_fwd(const float* emb, const float* Wt, const float* bias, float* hidden_all, float* z_all,
                           int* idx_all, unsigned char* argk_all, int B, int T, int C, int N, int k, ge_stream_t stream);
int ge_tgcn_recurrence_bwd(const float* emb, const float* W, const float* hidden_all, const float* z_all,
                           const int* idx_all, const unsigned char* argk_all, const float* dH, float* d_emb,
                           float* dWt_part, float* db_part, float* scratch, int B, int T, int C, int N, int k,
                           ge_stream_t stream);

/* ---- segmentation loss and score-map boxes (the full-resolution passes around the network) -------
 * seg_loss = DiceLoss()(pred, masks) + BCEWithLogitsLoss()(pred, masks) (train_cardiac_uda.py:228, train_camus_echo.py:212;
 * utils/losses.py:64-95: softmax over classes, BinaryDiceLoss(smooth=1, p=2) per class averaged over frames, / nc).
 * logits, target fp32 NCHW [F,nc,H*W], nc <= 8.  Forward: one pass; numden fp32 [F,nc,2] is saved for the backward;
 * loss fp32 [3] = (dice + bce, dice, bce).  Backward: one pass, dlogits = gout[0] * d loss[0] / d logits. */
size_t ge_seg_loss_workspace_bytes(int F, int nc, int HW);
int ge_seg_loss_fwd(const float* logits, const float* target, float* numden, float* loss, void* workspace,
                    size_t workspace_bytes, int F, int nc, int HW, float smooth, ge_stream_t stream);
int ge_seg_loss_bwd(const float* logits, const float* target, const float* numden, const float* gout,
                    float* dlogits, int F, int nc, int HW, ge_stream_t stream);
/* GModule.find_bbox / masks_to_boxes (models/graph_matching.py:702-746): boxes fp32 [planes,4] = (xmin,ymin,xmax,ymax)
 * of the "on" pixels of every [H,W] plane, (0,0,W,H) for an empty plane.  dtype GE_DTYPE_F32 or 2 (int64).
 * mode 0: on = value != 0; mode 1: on = value > 0 -- the box of `where(sigmoid(pred) > 0.5, 1, 0)`
 * (train_cardiac_uda.py:235) straight from the logits, without materialising the score map. */
int ge_mask_boxes(const void* maps, float* boxes, int dtype, int planes, int H, int W, int mode, ge_stream_t stream);

/* ---- node sampler, device half ------------------------------------------------------------------------
 * PrototypeComputation.prepare_targets / compute_targets_for_locations over GModule.compute_locations
 * (models/graph_matching.py:609-635, 874-959): the label of every location of every pyramid level (smallest-area class
 * box that strictly contains the location and whose largest side distance lies in the level's size range; else 0) and the
 * per-level (positive, negative) counts, one launch per domain.  boxes fp32 [B,K,4] (ge_mask_boxes), K <= 8, levels <= 5;
 * labels int64 [B * sum_l h_l*w_l], level-major / image-major inside a level; counts int32 [levels][2], ZERO-FILLED by the
 * caller.  heights / widths / strides / size_lo / size_hi are HOST arrays of length `levels`. */
int ge_sampler_labels(const float* boxes, long long* labels, int* counts, const int* heights, const int* widths,
                      const int* strides, const float* size_lo, const float* size_hi, int levels, int B, int K,
                      ge_stream_t stream);

/* f1, second half (models/graph_matching.py:978-1013): positive / negative picks of every level and the row gather of
 * the picked locations, ONE launch for up to 10 (domain, level) entries; ge_sampler_scatter is its backward.
 * Per entry: feats = NHWC feature map of the level [images,h,w,C] in `dtype`; labels int64 [n_loc] from
 * ge_sampler_labels; shift = batch_offset*h*w (row of the first labelled image in feats); n_neg_all = #labels == 0;
 * step = positive stride (n_pos_all // 100, >= 1); n_pos_pick / n_neg_pick = nodes kept; neg_all = 1 keeps every negative
 * (more positives than negatives), else the negatives of rank floor(linspace(0, n_neg_all-2, n_neg_pick)) (fp64, as
 * numpy); out_pos / out_neg = first slot of the level's positives / negatives in the domain's arrays nodes fp32
 * [n_nodes,C], node_labels int64 [n_nodes], src_row int64 [n_nodes] (row of feats each node came from; consumed by
 * the scatter).  Entries of one domain pass the same three bases.  Node order = [negatives level 0.. | positives level 0..]. */
int ge_sampler_gather(const void* const* feats, const long long* const* labels, float* const* nodes,
                      long long* const* node_labels, long long* const* src_row, const long long* n_loc,
                      const long long* shift, const int* n_neg_all, const int* step, const int* n_pos_pick,
                      const int* n_neg_pick, const int* neg_all, const int* out_pos, const int* out_neg,
                      int n_entries, int C, int dtype, ge_stream_t stream);
/* dfeats[i]: ZERO-FILLED gradient of feats[i]; receives the rows of dnodes[i] at src_row[i] (distinct rows: a scatter). */
int ge_sampler_scatter(void* const* dfeats, const float* const* dnodes, const long long* const* src_row,
                       const int* n_pos_pick, const int* n_neg_pick, const int* out_pos, const int* out_neg,
                       int n_entries, int C, int dtype, ge_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* GRAPHECHO_B200_H */
