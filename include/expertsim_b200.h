/*
 * expertsim_b200.h — C-ABI of the B200-native ExpertSim hot path (libexpertsim_b200.so).
 *
 * The reference (patrick-bedkowski/Generative-DNN-for-Physics-Simulations-CERN) has no native code and no FFI:
 * every "kernel" is an ATen/cuDNN/cuBLAS call issued from eager PyTorch.  Each entry point
 * below therefore cites the reference Python call site (file:line under /root/reference) whose arithmetic it
 * replaces.  INTEGRATION.md shows the ctypes binding a maintainer of the reference would add.
 *
 * Conventions
 *  - extern "C", plain pointers and sizes, no torch types.  Every function returns 0 on success or a negative
 *    es_status; es_last_error() returns a thread-local message for the last failure.
 *  - All buffers are CALLER-OWNED DEVICE pointers.  Work is enqueued asynchronously on `stream`
 *    (a cudaStream_t passed as void*); no hidden synchronisation, no allocation, no global mutable state.
 *  - "rows" are samples in EXPERT-SORTED order (the stable token->expert permutation produced by
 *    es_router_partition).  A group table `grp` is a device array of es_group; group g uses weight slot
 *    grp[g].slot, i.e. parameter pointer + slot * slot_stride.  Groups with rows == 0 do no work, so the
 *    reference's "B_e <= 1 -> skip expert" rule (models/moe.py:126-135) needs no host sync.
 *  - fp32 tensors are NCHW as in the reference; bf16 generator activations are NHWC (channels-last).
 */
#ifndef EXPERTSIM_B200_H
#define EXPERTSIM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
  ES_OK = 0,
  ES_ERR_INVALID = -1,     /* bad argument (shape, null pointer, unsupported size) */
  ES_ERR_CUDA = -2,        /* CUDA runtime error at launch */
  ES_ERR_UNSUPPORTED = -3
} es_status;

/* One contiguous run of rows that shares one set of expert weights. */
typedef struct {
  int32_t row_start;   /* first row (sample) of the group in the sorted batch */
  int32_t rows;        /* number of rows; 0 = inactive (skipped expert) */
  int32_t slot;        /* expert index = weight slot */
  int32_t pass_rows;   /* rows per generator pass (rows == passes * pass_rows) */
} es_group;

const char* es_last_error(void);
int es_version(void);
/* returns 1 when the loaded library contains sm_100a code and a CUDA device of CC 10.x is present */
int es_device_ok(void);

/* ------------------------------------------------------------------------------------------------
 * K1  router gating + sort-free stable token->expert permutation
 *     replaces RouterNetwork.forward (expertsim/models/routers/router.py:21-26, F.gumbel_softmax) and the
 *     routing block of MoEWrapper.train_step (expertsim/models/moe.py:76-77,97-103,123-126).
 * ---------------------------------------------------------------------------------------------- */
/* cond[B,9], router params (fc_layers.{0,2,4,6}), gumbel[B,E] (injected -log(Exp(1)) noise), tau ->
 * logits[B,E], gates[B,E] (softmax((logits+gumbel)/tau)), idx[B] (int64 argmax, first max wins),
 * hidden activations h1[B,128], h2[B,64], h3[B,32] (kept for es_router_bwd),
 * blk_hist[ceil(B/256), E] per-block expert histogram (input of es_router_partition). */
int es_router_fwd(const float* cond, int B, int E,
                  const float* w0, const float* b0, const float* w2, const float* b2,
                  const float* w4, const float* b4, const float* w6, const float* b6,
                  const float* gumbel, float tau,
                  float* logits, float* gates, int64_t* idx,
                  float* h1, float* h2, float* h3, int32_t* blk_hist, void* stream);

/* Stable partition: counts[E], offsets[E+1], perm[B] (perm[j] = original index of the j-th sorted sample;
 * ascending inside an expert, exactly (idx==e).nonzero() of moe.py:123), plus the group tables used by every
 * grouped kernel:  grp_half[E]  rows = counts[e] (0 if counts[e] < min_rows)   row_start = offsets[e]
 *                  grp_gen[E]   the generator's two-pass batch: row_start = 2*offsets[e], rows = 2*counts[e]
 * min_rows = 2 in training (skip rule), 1 at inference.  scratch: int32[ceil(B/256)*E + E]. */
int es_router_partition(const int64_t* idx, int B, int E, int min_rows, const int32_t* blk_hist,
                        int32_t* counts, int32_t* offsets, int32_t* perm,
                        es_group* grp_half, es_group* grp_gen, int32_t* scratch, void* stream);

/* out[j, :] = in[perm[j], :] for j < B  (row gather, fp32, `width` floats per row) — cond[mask], real[mask]...
 * (moe.py:143,150,165-168) done once for all experts. */
int es_gather_rows(const float* in, const int32_t* perm, int B, int width, float* out, void* stream);
/* out[perm[j], :] = in[j, :]  (scatter back to original order; moe.py:197-198) */
int es_scatter_rows(const float* in, const int32_t* perm, int B, int width, float* out, void* stream);

/* Router loss (moe.py:407-435 with train/utils.py:398-419,623-642) and its backward through the gumbel
 * softmax and the MLP, accumulated into the router's gradient buffers (dw*, db* must be zeroed by the caller).
 *   alb  = alb_strength * mean_e exp(1/(sum_b gates[b,e] + 1e-6)), weighted by alb_weight (decreasing_weight)
 *   ent  = +util_strength * sum_e pbar_e log(pbar_e + 1e-9),  pbar = mean_b gates
 * gate_sums[E] (sum_b gates[b,e]) must have been produced by es_router_gate_sums (all-reduced across ranks
 * by the caller under data parallelism).  B_global = global batch for the mean in pbar.
 * extra_dgates (nullable) [B,E] is added to dL/dgates (the expert-distribution term).
 * losses_out[2] = {alb (unweighted by alb_weight), entropy term}. */
int es_router_gate_sums(const float* gates, int B, int E, float* gate_sums, void* stream);
int es_router_bwd(const float* cond, int B, int E, int B_global,
                  const float* w2, const float* w4, const float* w6,
                  const float* gates, const float* h1, const float* h2, const float* h3,
                  const float* gate_sums, float tau,
                  float alb_strength, float alb_weight, float util_strength, const float* extra_dgates,
                  float* dw0, float* db0, float* dw2, float* db2, float* dw4, float* db4, float* dw6, float* db6,
                  float* losses_out, void* stream);
/* expert-distribution loss (train/utils.py:372-395 as used at moe.py:264-268): gates are the straight-through
 * one-hot gates, m[B] the per-sample photon sums.  loss_out[1] += ed_strength*0.1*sum_{b,b'}[idx_b==idx_b']|m_b-m_b'|/B;
 * dgates[B,E] = d loss / d gates_soft. */
int es_router_ed_loss(const int64_t* idx, const float* m, int B, int E, float ed_strength,
                      float* dgates, float* loss_out, void* stream);

/* ------------------------------------------------------------------------------------------------
 * K4  fused loss tails (warp-shuffle reductions, vectorised loads)
 * ---------------------------------------------------------------------------------------------- */
/* Discriminator hinge loss (moe.py:518-523): per active group e
 *   loss[e] = (mean relu(1-real) + mean relu(1+fake)) * counts_global[e]/B_global ; d_real/d_fake = dloss/dscore.
 * counts_global[E] float (global B_e; equals the local count on one GPU). */
int es_hinge_d(const float* real_score, const float* fake_score, const es_group* grp, int E,
               const float* counts_global, int B_global,
               float* d_real, float* d_fake, float* loss, void* stream);

/* Generator loss tails, phase 1 (per-sample terms + per-expert partial sums):
 *   photon sums s[r] = sum_hw(exp(img)-1)                          (moe.py:611-616)
 *   SDI div[r] = mean|lat1-lat2| / (mean|z1-z2| + 1e-5)            (moe.py:576-580)
 *   sums[e][0..7] += {sum std, sum 1/(div+1e-5), sum s, sum s^2, sum|s-I|, sum logcosh-coord terms, sum score1, rows}
 * img[R,HW] fp32, lat[R,64], z[R,10], std[R], intensity[R], coords/pos[R,2], score1[R]. */
int es_gen_loss_reduce(const float* img, int HW, const float* lat1, const float* lat2,
                       const float* z1, const float* z2, const float* stdv, const float* intensity,
                       const float* coords, const float* pos, const float* score1,
                       const es_group* grp, int E, int total_rows, float* s_out, float* div_out, double* sums, void* stream);
/* phase 2: per-expert losses and every gradient that enters the backward pass (moe.py:544-563):
 *   losses[e][0..5] = {gen_loss (total, scaled by B_e/B), div_loss, intensity_loss, aux_loss, std(s), mean(s)}
 *   d_score1[R], d_lat1/d_lat2[R,64], d_coords[R,2], and d_img[R,HW] += in_strength*sign(s-I)/B_e * exp(img) * w_e
 * sums[E][8] are the (all-reduced) partial sums of phase 1. */
int es_gen_loss_grads(const float* img, int HW, const float* lat1, const float* lat2,
                      const float* z1, const float* z2, const float* stdv, const float* intensity,
                      const float* coords, const float* pos, const float* s, const float* divv,
                      const es_group* grp, int E, int total_rows, const double* sums, int B_global,
                      float di_strength, float in_strength, float aux_strength,
                      float* d_score1, float* d_lat1, float* d_lat2, float* d_coords, float* d_img,
                      float* losses, void* stream);
/* d_coords[r][j] = aux_strength * tanh(coords - pos) / (2 B_global) for rows of active groups, 0 elsewhere: the
 * regression-loss gradient of es_gen_loss_grads on its own (proton/aux_reg.py:42-45 scaled as moe.py:559-562), so that the
 * auxiliary regressor's backward does not have to wait for the discriminator passes of the generator step. */
int es_aux_loss_grad(const float* coords, const float* pos, const es_group* grp, int E, int total_rows, int B_global,
                     float aux_strength, float* d_coords, void* stream);

/* ------------------------------------------------------------------------------------------------
 * K2  generator: grouped (per-expert) bf16 tcgen05 GEMM / implicit GEMM
 *     replaces Generator.forward (expertsim/models/proton/generator.py:46-52) and its autograd backward.
 * ---------------------------------------------------------------------------------------------- */
/* Geometry of one implicit-GEMM convolution over NHWC bf16 activations.  The source map [Hs,Ws,C] is virtually
 * nearest-upsampled to [Hu,Wu] (src = floor(dst*Hs/Hu), torch 'nearest'), zero padded by `pad`, and convolved
 * with a KHxKW window, stride 1, giving [Ho,Wo,N].  A dense layer is Hs=Ws=Hu=Wu=Ho=Wo=KH=KW=1, pad=0. */
typedef struct {
  int32_t Hs, Ws, C;
  int32_t Hu, Wu;
  int32_t Ho, Wo;
  int32_t KH, KW, pad;
  int32_t N;            /* output channels */
} es_conv_geom;

/* y[row, oy, ox, n] = sum_{ky,kx,c} x_up[row, oy+ky-pad, ox+kx-pad, c] * w[slot][n][ky][kx][c] (+ bias[slot][n])
 * x: bf16 [rows,Hs,Ws,C]; w: bf16 packed [slots][N][KH*KW*C]; bias fp32, slot s at bias + s*bias_slot_stride, or NULL;
 * y: bf16 [rows,Ho,Wo,N].  Used for the forward convs, for their data gradients (with transposed/flipped packed
 * weights) and for fc2. */
int es_igemm_fwd(const void* x, const void* w, const float* bias, long bias_slot_stride, void* y, const es_conv_geom* g,
                 const es_group* grp, int n_groups, int total_rows, void* stream);
/* Same, with the following GroupNorm's statistics fused into the epilogue (proton/generator.py:27-29,33-35,38-40: every conv
 * is followed by GroupNorm(32)): pair_sums fp32 [total rows][N/2][2], ZEROED by the caller, receives per (row, channel pair)
 * the sum and the sum of squares of the bf16-rounded outputs over the row's pixels (atomic adds; several calls writing
 * disjoint pixels of one map accumulate).  *fused (HOST int) is set to 1 when the kernel variant chosen for this geometry
 * did accumulate (the TMA-fed CTA-pair variants 2 and 3), to 0 when it did not — the caller then runs es_gn_lrelu_fwd, which
 * computes its own statistics, instead of es_gn_lrelu_apply_fwd. */
int es_igemm_fwd_sums(const void* x, const void* w, const float* bias, long bias_slot_stride, void* y, const es_conv_geom* g,
                      const es_group* grp, int n_groups, int total_rows, float* pair_sums, int32_t* fused, void* stream);
/* Generalised reduction of the same kernel: a TAP TABLE instead of a dense KHxKW window.  Tap t reads the (virtually
 * nearest-upsampled [Hs,Ws]->[Hu,Wu]) source at (oy*my + tap_dy[t], ox*mx + tap_dx[t]) (outside [0,Hu)x[0,Wu) = zero) and
 * multiplies it with the C weights starting at column tap_koff[t] of the packed weight row of length KK.  (oy, ox) runs
 * over the M-space grid [Ho, Wo]; the result is stored at pixel ((oy*o_my + o_oy), (ox*o_mx + o_ox)) of a [Ho_full, Wo_full]
 * map.  This is how the x2 nearest upsample in front of a conv is folded away: output phase (py, px) of such a conv only
 * sees ceil((KH+1-py)/2)-ish distinct source rows, so four phase convs with pre-summed taps (es_fold_up2_weights) do
 * 25 instead of 64 tap-MACs per 4 outputs for k4/p1 (2.56x fewer), and the data gradient is ONE table-conv over dy
 * (my = mx = 2) writing the low-resolution gradient directly. */
/* Host-only: which kernel variant es_igemm_fwd picks for geometry g (no device work; callable without a GPU).
 * plan8 = {variant, BN, tap rows (variant 2: taps), taps per row (nx), M-axis row pitch (Wo, or Wo + nx - 1 for strips),
 *          pipeline stages, pipeline steps per tile, M tiles per row}.  variant 2 = TMA-fed CTA pair (A by TMA im2col, B halves
 * by TMA, tcgen05.mma.cta_group::2): the conv reads its source directly (no nearest upsample in between), N tile >= 64, not
 * the dense 1x1 product.  variant 3 = the same CTA pair with TAP-ROW STRIPS: N tile <= 128, taps form equal rows of
 * consecutive dx (mx = 1) — one im2col strip of 128 + nx - 1 pixels per tap row and channel block serves the nx taps as
 * row-shifted A descriptors, two 256-row sub-tiles share each weight box; plan8[7] counts 128-row blocks of the padded pitch.
 * variant 1 = strip (one cp.async-gathered strip per tap row, kx taps as row-shifted A descriptors):
 * no upsample, N <= 128, 2..4 taps per row, nx * BN <= 384 — taken where variant 2 does not apply (N = 32).  variant 0 =
 * single-CTA kernel with the cp.async gather (nearest upsample inside the conv, fc2). */
int es_igemm_fwd_plan(const es_conv_geom* g, int total_rows, int32_t* plan8);

typedef struct {
  int32_t Hs, Ws, C;
  int32_t Hu, Wu;
  int32_t Ho, Wo;
  int32_t my, mx;
  int32_t n_taps;
  int8_t tap_dy[32], tap_dx[32];
  int32_t tap_koff[32];
  int32_t KK;
  int32_t N;
  int32_t o_my, o_oy, o_mx, o_ox, Ho_full, Wo_full;
} es_tap_geom;
int es_igemm_taps_fwd(const void* x, const void* w, const float* bias, long bias_slot_stride, void* y, const es_tap_geom* g,
                      const es_group* grp, int n_groups, int total_rows, void* stream);
/* tap-table form of es_igemm_fwd_sums (the classes of a folded conv write disjoint pixels of one map: their sums add up) */
int es_igemm_taps_fwd_sums(const void* x, const void* w, const float* bias, long bias_slot_stride, void* y, const es_tap_geom* g,
                           const es_group* grp, int n_groups, int total_rows, float* pair_sums, int32_t* fused, void* stream);

/* folded-tap table of a conv behind a nearest upsample: folded tap t is the SUM of the original taps (ky, kx) whose bit
 * ky*KW + kx is set in mask[t] (they all read the same source pixel for the output class the tap belongs to; KH*KW <= 32).
 * x2 upsample, class = output phase (py, px), source offset (dy, dx): (ky, kx) is in the mask iff
 * floor((py+ky-pad)/2) == dy and floor((px+kx-pad)/2) == dx. */
typedef struct {
  int32_t n_taps;
  uint32_t mask[32];
} es_fold_table;
/* w fp32 [slots][N][C][KH][KW] -> w_fwd bf16 [slots][N][n_taps][C] and/or w_dgrad bf16 [slots][C][n_taps][N] (pre-summed) */
int es_fold_up2_weights(const float* w, long slot_stride, int slots, int N, int C, int KH, int KW,
                        const es_fold_table* t, void* w_fwd, void* w_dgrad, void* stream);
/* the same for several tables in ONE pass over the weights: outs[j] receives table j's folded weights, all in the forward
 * layout [slots][N][n_taps_j][C] (dgrad_layout == 0) or all in the data-gradient layout [slots][C][n_taps_j][N] (!= 0).
 * tables / outs are HOST arrays of n_tables <= 13 entries (copied into the launch). */
int es_fold_weights_multi(const float* w, long slot_stride, int slots, int N, int C, int KH, int KW,
                          const es_fold_table* tables, int n_tables, void* const* outs, int dgrad_layout, void* stream);
/* dw_ref[slot][n][c][ky][kx] += sum over the folded taps whose mask contains (ky, kx), of dw_folded[slot][n][t][c] */
int es_unfold_up2_wgrad(const float* dw_folded, int slots, int N, int C, int KH, int KW, const es_fold_table* t,
                        float* dw_ref, long slot_stride, void* stream);
/* all classes at once: dw_ref[slot][n][c][ky][kx] += sum_j sum over table j's folded taps ...  (dw_folded: HOST array of
 * n_tables <= 13 device pointers, each [slots][N][n_taps_j][C]) */
int es_unfold_wgrad_multi(const float* const* dw_folded, int slots, int N, int C, int KH, int KW,
                          const es_fold_table* tables, int n_tables, float* dw_ref, long slot_stride, void* stream);
/* dst[row, a*Wo + b, :] = src[row, (a*my + oy)*Wo_full + b*mx + ox, :]  (NHWC bf16; one output phase made contiguous) */
int es_pick_pixels(const void* src, int Ho_full, int Wo_full, int C, int my, int oy, int mx, int ox, int Ho, int Wo,
                   int total_rows, void* dst, void* stream);
/* weight gradient with a tap table (see es_igemm_taps_fwd): dy is [rows, Ho*Wo, N], x the source map;
 * dw[slot][n][tap_koff[t] + c] += sum_pix dy[pix, n] * x[(oy*my + tap_dy[t], ox*mx + tap_dx[t]), c]; a dw row has KK columns */
int es_igemm_taps_wgrad(const void* x, const void* dy, float* dw, const es_tap_geom* g, const es_group* grp, int n_groups,
                        int total_rows, void* stream);

/* dw[slot][n][ky][kx][c] += sum_{row,oy,ox} dy[row,oy,ox,n] * x_up[row,oy+ky-pad,ox+kx-pad,c]   (fp32, packed layout,
 * split-K with fp32 atomics; dw must be zeroed by the caller). */
int es_igemm_wgrad(const void* x, const void* dy, float* dw, const es_conv_geom* g,
                   const es_group* grp, int n_groups, int total_rows, void* stream);
/* dense data gradient with the weight read in its forward layout: dx[row, k] += sum_n dy[row, n] * w[slot][n][k]
 * (fc2: N=92160, K=256; split over n with fp32 atomics; dx fp32 [rows,K] zeroed by the caller). */
int es_dense_dgrad(const void* dy, const void* w, float* dx, int N, int K,
                   const es_group* grp, int n_groups, int total_rows, void* stream);
/* dense weight gradient: dw[slot*dw_slot_stride + row_map[n]*K + k] = sum_row dy[row,n] * x[row,k]  (fp32, direct store;
 * row_map (nullable) un-permutes the channels-last feature order back to the reference's NCHW flattening).
 * scratch (nullable): n_groups * round_up(total_rows, 64) * K bf16 — the zero-padded per-group copy of x that lets the
 * TMA-fed kernel (K = 256) stream dy boxes across group boundaries; without it the cp.async gather kernel runs. */
int es_dense_wgrad(const void* dy, const void* x, float* dw, long dw_slot_stride, int N, int K, const int32_t* row_map,
                   const es_group* grp, int n_groups, int total_rows, void* scratch, void* stream);
/* Test-only SIMT (CUDA-core, fp32 accumulate) versions of the two implicit GEMMs; same arguments.  They exist so the
 * tcgen05 path can be cross-checked on the device at sizes the CPU oracle cannot reach.  Never called by the product. */
int es_igemm_fwd_simt(const void* x, const void* w, const float* bias, long bias_slot_stride, void* y, const es_conv_geom* g,
                      const es_group* grp, int n_groups, int total_rows, void* stream);
int es_igemm_wgrad_simt(const void* x, const void* dy, float* dw, const es_conv_geom* g,
                        const es_group* grp, int n_groups, int total_rows, void* stream);

/* generator head: x0 = [z | cond] (19) -> Linear(19,256) + LayerNorm(256) + LeakyReLU(0.1) (proton/generator.py:13-17).
 * In the two-pass training batch, row r of group e takes z from z1 (first pass_rows rows) or z2 and cond from the
 * half-batch row offsets.  lin[rows,256] fp32 (pre-norm, kept for backward), h[rows,256] bf16.
 * gamma == NULL selects the linear head only (neutron: h = bf16(x0 W^T + b), BatchNorm follows as its own pass; in the
 * backward dh is then the gradient w.r.t. the linear output and lin/beta/dgamma/dbeta may be NULL). */
int es_gen_fc1_fwd(const float* z1, const float* z2, const float* cond, const float* w, const float* b,
                   const float* gamma, const float* beta, long slot_stride_w, long slot_stride_v,
                   const es_group* grp_gen, int E, int total_rows, int two_pass, float* x0, float* lin, void* h, void* stream);
int es_gen_fc1_bwd(const float* dh, const float* x0, const float* lin, const float* gamma, const float* beta,
                   long slot_stride_w, long slot_stride_v, const es_group* grp_gen, int E, int total_rows,
                   float* dw, float* db, float* dgamma, float* dbeta, void* stream);

/* LayerNorm over all F features of a row + LeakyReLU (fc2.1: F=92160; proton/generator.py:18-22), bf16 in/out,
 * fp32 statistics.  stats[row][2] = {mean, rstd}. */
int es_ln_lrelu_fwd(const void* x, const float* gamma, const float* beta, long slot_stride, int F,
                    const es_group* grp, int n_groups, int total_rows, void* y, float* stats, void* stream);
/* GroupNorm(groups) + LeakyReLU over NHWC bf16 [rows,P,C] (proton/generator.py:28-29,34-35,39-40). stats[row][groups][2]. */
int es_gn_lrelu_fwd(const void* x, const float* gamma, const float* beta, long slot_stride, int P, int C, int groups,
                    const es_group* grp, int n_groups, int total_rows, void* y, float* stats, void* stream);
/* Same, x is [rows,Hs,Ws,C] and y is stored NEAREST-UPSAMPLED ALONG x as [rows,Hs,Wu,C] (torch's nearest rule: source column
 * min(floor(xu*Ws/Wu), Ws-1)) — the x half of the Upsample in front of proton conv2 (proton/generator.py:30-32), materialised by
 * the producer so that the conv reads its source directly and both GEMM operands can come by TMA.  stats as above. */
int es_gn_lrelu_fwd_upx(const void* x, const float* gamma, const float* beta, long slot_stride, int Hs, int Ws, int Wu, int C,
                        int groups, const es_group* grp, int n_groups, int total_rows, void* y, float* stats, void* stream);
/* GroupNorm + LeakyReLU as one streaming pass over x [rows,Hs,Ws,C] when the producing conv accumulated pair_sums
 * (es_igemm_fwd_sums / es_igemm_taps_fwd_sums, all calls fused): mean / rstd per (row, group) come from the sums (variance =
 * E[x^2] - mean^2), are written to stats like es_gn_lrelu_fwd does, and y is stored as [rows,Hs,Wu,C] — Wu == Ws: plain;
 * Wu > Ws: nearest-upsampled along x as es_gn_lrelu_fwd_upx.  C/groups must be even. */
int es_gn_lrelu_apply_fwd(const void* x, const float* pair_sums, const float* gamma, const float* beta, long slot_stride, int Hs,
                          int Ws, int Wu, int C, int groups, const es_group* grp, int n_groups, int total_rows, void* y,
                          float* stats, void* stream);
/* Backward of norm+LeakyReLU.  `dy_up` is the gradient w.r.t. the (virtually upsampled) consumer input
 * [rows,Hu,Wu,C]; it is summed over the pixels that map to each source pixel [Hs,Ws] (nearest-upsample backward).
 * dx (bf16, gradient w.r.t. the pre-norm tensor), dgamma/dbeta/dbias_conv (fp32, atomically accumulated). */
int es_gn_lrelu_bwd(const void* dy_up, int Hs, int Ws, int Hu, int Wu, const void* x, const float* stats,
                    const float* gamma, const float* beta, long slot_stride, int C, int groups,
                    const es_group* grp, int n_groups, int total_rows,
                    void* dx, float* dgamma, float* dbeta, float* dbias_conv, void* stream);
int es_ln_lrelu_bwd(const void* dy_up, int Hs, int Ws, int Hu, int Wu, int C, const void* x, const float* stats,
                    const float* gamma, const float* beta, long slot_stride,
                    const es_group* grp, int n_groups, int total_rows, void* dx, void* stream);
/* column reductions of the big LayerNorm: dgamma[slot][f] += sum_row dyn*xhat, dbeta += sum_row dyn, dbias_lin += sum_row dx;
 * results are written at row_map[f] (reference NCHW feature order) when row_map is given */
int es_ln_affine_bwd(const void* dy_up, int Hs, int Ws, int Hu, int Wu, int C, const void* x, const void* dx,
                     const float* stats, const float* gamma, const float* beta, long slot_stride,
                     const es_group* grp, int n_groups, int total_rows, const int32_t* row_map, long out_slot_stride,
                     float* dgamma, float* dbeta, float* dbias_lin, void* stream);

/* last generator layer: Conv2d(64->1, k2, pad 1) + ReLU (proton/generator.py:42-43) on CUDA cores.
 * x bf16 [rows,Hs,Ws,64] -> img fp32; in the two-pass batch the image row goes to img1 or img2 [half rows, Ho*Wo]. */
int es_gen_out_fwd(const void* x, const float* w, const float* b, long slot_stride_w, long slot_stride_b, int Hs, int Ws, int C, int KH, int KW, int pad,
                   const es_group* grp_gen, int E, int total_rows, int two_pass, float* img1, float* img2, void* stream);
int es_gen_out_bwd(const void* x, const float* w, long slot_stride_w, long slot_stride_b, int Hs, int Ws, int C, int KH, int KW, int pad,
                   const float* img1, const float* img2, const float* dimg1, const float* dimg2,
                   const es_group* grp_gen, int E, int total_rows, int two_pass,
                   void* dx, float* dw, float* db, void* stream);

/* weight packing: fp32 reference layout -> bf16 kernel layout, all slots in one launch.
 *   conv  [N,C,KH,KW] -> fwd  [N][KH][KW][C]      and (transposed, flipped) dgrad [C][KH][KW][N]
 *   dense [N,K] with output-row permutation row_map (packed row n <- reference row row_map[n]) */
int es_pack_conv_weight(const float* w, long slot_stride, int slots, int N, int C, int KH, int KW,
                        void* w_fwd, void* w_dgrad, void* stream);
int es_pack_dense_weight(const float* w, long slot_stride, int slots, int N, int K, const int32_t* row_map,
                         void* w_packed, void* stream);
/* dw_ref[slot][n][c][ky][kx] = dw_packed[slot][n][ky][kx][c] */
int es_unpack_conv_wgrad(const float* dw_packed, int slots, int N, int C, int KH, int KW, float* dw_ref, long slot_stride, void* stream);
/* out[slot][row_map[f]] = in[slot][f]  (permute per-feature vectors between channels-last and NCHW feature order) */
int es_permute_features(const float* in, long in_stride, const int32_t* row_map, int slots, int F, float* out, long out_stride,
                        int inverse, void* stream);

/* ------------------------------------------------------------------------------------------------
 * K3  discriminator and auxiliary regressor building blocks (fp32, NCHW, grouped by expert)
 *     replace Discriminator.forward (proton/discriminator.py:148-155), AuxReg.forward (proton/aux_reg.py:33-40,
 *     84-96,123-131) and their autograd backward.
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
  int32_t Ci, Hi, Wi, Co, Ho, Wo, KH, KW, stride, pad;
} es_conv2d;

int es_conv2d_fwd(const float* x, const float* w, const float* b, long slot_stride_w, long slot_stride_b,
                  const es_conv2d* g, const es_group* grp, int n_groups, int total_rows, float* y, void* stream);
int es_conv2d_bwd_data(const float* dy, const float* w, long slot_stride_w, const es_conv2d* g,
                       const es_group* grp, int n_groups, int total_rows, float* dx, int accumulate, void* stream);
int es_conv2d_bwd_weight(const float* x, const float* dy, const es_conv2d* g, const es_group* grp, int n_groups,
                         int total_rows, float* dw, float* db, long slot_stride_w, long slot_stride_b, void* stream);

/* act: 0 none, 1 ReLU, 2 LeakyReLU(0.1).  GroupNorm over [C/groups, H*W] per sample; stats[row][groups][2]. */
int es_groupnorm_fwd(const float* x, const float* gamma, const float* beta, long slot_stride, int C, int HW, int groups, int act,
                     const es_group* grp, int n_groups, int total_rows, float* y, float* stats, void* stream);
int es_groupnorm_bwd(const float* dy, const float* x, const float* stats, const float* gamma, const float* beta, long slot_stride,
                     int C, int HW, int groups, int act, const es_group* grp, int n_groups, int total_rows,
                     float* dx, float* dgamma, float* dbeta, void* stream);
/* LayerNorm over the last dimension F (<=1024) + activation; stats[row][2]. */
int es_layernorm_fwd(const float* x, const float* gamma, const float* beta, long slot_stride, int F, int act,
                     const es_group* grp, int n_groups, int total_rows, float* y, float* stats, void* stream);
int es_layernorm_bwd(const float* dy, const float* x, const float* stats, const float* gamma, const float* beta, long slot_stride,
                     int F, int act, const es_group* grp, int n_groups, int total_rows,
                     float* dx, float* dgamma, float* dbeta, void* stream);
/* MaxPool2d(kernel (kh,kw), stride (sh,sw)), floor mode; idx (uint8) = winning tap for backward. */
int es_maxpool_fwd(const float* x, int C, int Hi, int Wi, int kh, int kw, int sh, int sw, int total_rows,
                   float* y, uint8_t* idx, void* stream);
int es_maxpool_bwd(const float* dy, const uint8_t* idx, int C, int Hi, int Wi, int kh, int kw, int sh, int sw, int total_rows,
                   float* dx, void* stream);
/* Fused discriminator trunk (expertsim/models/proton/discriminator.py:121-155, neutron/discriminator.py:11-48).
 * Stem = SN-Conv(1->32,k3,valid) -> GroupNorm(8) -> LeakyReLU(0.1) -> MaxPool(2): one launch; only the pooled map
 * p1 [rows][32][(H-2)/2][(W-2)/2] and the statistics stats [rows][8][2] are written.  w = normalised weights
 * [slots][32*9] (slot stride slot_stride_w); bias / gamma / beta live in the arena (slot strides given).
 * Backward re-computes the conv from the image: dp1 = gradient of p1; d_img [rows][H*W] is ACCUMULATED with atomics (the
 * caller zeroes it) or NULL; dw [slots][32*9] (+ dbias, dgamma, dbeta, all accumulated) or NULL. */
int es_disc_stem_fwd(const float* img, const float* w, long slot_stride_w, const float* bias, long slot_stride_b,
                     const float* gamma, const float* beta, long slot_stride_n, int H, int W, const es_group* grp,
                     int n_groups, int total_rows, float* p1, float* stats, void* stream);
int es_disc_stem_bwd(const float* dp1, const float* img, const float* w, long slot_stride_w, const float* bias,
                     long slot_stride_b, const float* gamma, const float* beta, long slot_stride_n, const float* stats, int H,
                     int W, const es_group* grp, int n_groups, int total_rows, float* d_img, float* dw, long slot_stride_dw,
                     float* dbias, float* dgamma, float* dbeta, void* stream);
/* Stage 2 = SN-Conv(32->16,k3,valid) -> GroupNorm(8) -> LeakyReLU -> MaxPool(2, pool_kw) -> flatten, written straight into
 * the fc1 input fcin [rows][ldf] (features 0..flat-1, then the 9 conditionals).  p1 [rows][32][H1][W1]; y2 (pre-norm conv
 * output, [rows][16][(H1-2)*(W1-2)]) and stats [rows][8][2] are kept for the backward.  Backward: dfc = gradient of
 * fcin (only the first `flat` columns are read); dp1 = gradient of p1 (written); dw [slots][16*288] etc. or NULL. */
int es_disc_stage2_fwd(const float* p1, const float* w, long slot_stride_w, const float* bias, long slot_stride_b,
                       const float* gamma, const float* beta, long slot_stride_n, const float* cond, int H1, int W1,
                       int pool_kw, const es_group* grp, int n_groups, int total_rows, float* y2, float* stats, float* fcin,
                       int ldf, void* stream);
int es_disc_stage2_bwd(const float* dfc, int ldf, const float* y2, const float* stats, const float* p1, const float* w,
                       long slot_stride_w, const float* gamma, const float* beta, long slot_stride_n, int H1, int W1,
                       int pool_kw, const es_group* grp, int n_groups, int total_rows, float* dp1, float* dw,
                       long slot_stride_dw, float* dbias, long slot_stride_b, float* dgamma, float* dbeta, void* stream);
/* y[row, 0:O] = x[row, 0:I] . W[slot][O,I]^T + b   /  dx = dy . W  /  dW += dy^T x, db += sum dy */
int es_linear_fwd(const float* x, int ldx, const float* w, const float* b, long slot_stride_w, long slot_stride_b, int I, int O,
                  const es_group* grp, int n_groups, int total_rows, float* y, void* stream);
int es_linear_bwd_data(const float* dy, const float* w, long slot_stride_w, int I, int O,
                       const es_group* grp, int n_groups, int total_rows, float* dx, int lddx, void* stream);
int es_linear_bwd_weight(const float* x, int ldx, const float* dy, int I, int O, const es_group* grp, int n_groups, int total_rows,
                         float* dw, float* db, long slot_stride_w, long slot_stride_b, void* stream);
/* Spectral norm (hook-based torch.nn.utils.spectral_norm, torch/nn/utils/spectral_norm.py:62-113): per slot, one power
 * iteration in place on u[O], v[I] (if do_power_iter), sigma = u.(W v), w_sn = w_orig / sigma.  sigma_out[slot].
 * scratch (nullable): slots*(I+O+2) floats (fwd) / slots floats (bwd); when given, matrices of >= 16384 elements run the
 * multi-CTA path (three launches, partial norms by atomics) instead of one CTA per slot. */
int es_spectral_norm_fwd(const float* w_orig, float* u, float* v, long slot_stride_w, long slot_stride_u, long slot_stride_v,
                         int slots, int O, int I, int do_power_iter, const es_group* grp, float* w_sn, long slot_stride_sn,
                         float* sigma_out, float* u_used, float* v_used, float* scratch, void* stream);
/* dw_orig += (dw_sn - <dw_sn, w_sn> u v^T) / sigma; slots whose grp[slot].rows == 0 are skipped (grp nullable: all slots,
 * where grp[s] must describe slot s as produced by es_router_partition) */
int es_spectral_norm_bwd(const float* dw_sn, const float* w_sn, const float* u_used, const float* v_used, const float* sigma,
                         long slot_stride_sn, int slots, int O, int I, float* dw_orig, long slot_stride_w,
                         const es_group* grp, float* scratch, void* stream);
/* elementwise helpers: y = a + b then ReLU (residual join), its backward mask, mean over HW, dropout with a given keep-mask */
int es_add_relu_fwd(const float* a, const float* b, long n, float* y, void* stream);
int es_relu_bwd(const float* dy, const float* y, long n, float* dx, void* stream);
int es_gap_fwd(const float* x, int C, int HW, int total_rows, float* y, void* stream);
int es_gap_bwd(const float* dy, int C, int HW, int total_rows, float* dx, void* stream);
int es_dropout(const float* x, const float* keep_mask, float p, long n, float* y, void* stream);
int es_axpy(float alpha, const float* x, long n, float* y, void* stream);   /* y += alpha * x */
int es_copy_cols(const float* src, int lds, int cols, int rows, float* dst, int ldd, int col0, void* stream);

/* ------------------------------------------------------------------------------------------------
 * BatchNorm (+ Dropout + LeakyReLU) of the neutron networks (expertsim/models/neutron/generator.py:11-40,
 * expertsim/models/neutron/aux_reg.py:11-49), forward and backward, grouped by expert.
 * A stat group sg = 2*slot + pass is what one reference forward call normalises over (the generator's two-pass batch
 * = two calls).  sums / sums2: double [2*slots][CS][2]  ((sum x, sum x^2) / (sum g, sum g*xhat)), zeroed by the caller and,
 * under data parallelism, all-reduced by the caller between the reduce and the apply call (SyncBN).
 * n_sg: float [2*slots] GLOBAL element count of a stat group (rows * pixels).  stats: float [2*slots][CS][2] (mean, rstd).
 * chmap (nullable): stat channel -> parameter index.  Dropout: keep_mask (nullable; reference NCHW activation layout,
 * element (row*C + c)*P + p) or a counter hash of (seed, element index); p_drop = 0 disables it.
 * NHWC entry points (generator, bf16): a row is [Hs*Ws][C]; feat_stats != 0 = BatchNorm1d over the flattened map
 * (CS = Hs*Ws*C), else BatchNorm2d (CS = C); order BN -> Dropout -> LeakyReLU.  dy_up lives on the (Hu,Wu) grid of the
 * consumer (nearest-upsample backward folded in).  NCHW entry points (aux regressor, fp32): BN -> LeakyReLU -> Dropout.
 * ---------------------------------------------------------------------------------------------- */
int es_bn_stats_nhwc(const void* x, int Hs, int Ws, int C, int feat_stats, const es_group* grp, int E, int total_rows,
                     int two_pass, double* sums, void* stream);
/* training: sums -> stats, running_mean/var <- momentum update per pass (unbiased variance), num_batches_tracked += passes;
 * eval (training == 0): stats <- running buffers.  Slots with grp[slot].rows == 0 are untouched. */
int es_bn_finalize(const double* sums, const float* n_sg, int CS, int passes, int training, float momentum,
                   const int32_t* chmap, float* running_mean, float* running_var, long buf_slot_stride,
                   int64_t* num_batches_tracked, long nbt_slot_stride, const es_group* grp, int slots, float* stats,
                   void* stream);
int es_bn_apply_fwd_nhwc(const void* x, int Hs, int Ws, int C, int feat_stats, const float* stats, const float* gamma,
                         const float* beta, long slot_stride, const int32_t* chmap, const float* keep_mask,
                         unsigned long long seed, float p_drop, const es_group* grp, int E, int total_rows, int two_pass,
                         void* y, void* stream);
int es_bn_bwd_reduce_nhwc(const void* dy_up, int Hs, int Ws, int Hu, int Wu, int C, int feat_stats, const void* x,
                          const float* stats, const float* gamma, const float* beta, long slot_stride, const int32_t* chmap,
                          const float* keep_mask, unsigned long long seed, float p_drop, const es_group* grp, int E,
                          int total_rows, int two_pass, double* sums2, void* stream);
int es_bn_bwd_apply_nhwc(const void* dy_up, int Hs, int Ws, int Hu, int Wu, int C, int feat_stats, const void* x,
                         const float* stats, const double* sums2, const float* n_sg, const float* gamma, const float* beta,
                         long slot_stride, const int32_t* chmap, const float* keep_mask, unsigned long long seed,
                         float p_drop, const es_group* grp, int E, int total_rows, int two_pass, void* dx, void* stream);
/* dgamma[slot][chmap[c]] += scale * sum_pass sums2[.][c][1],  dbeta += scale * sum_pass sums2[.][c][0]
 * (scale = 1/world when sums2 was all-reduced and the gradient arena is sum-all-reduced afterwards) */
int es_bn_affine_grads(const double* sums2, int CS, int passes, float scale, const int32_t* chmap, const es_group* grp,
                       int slots, float* dgamma, float* dbeta, long slot_stride, void* stream);
int es_bn2d_stats(const float* x, int C, int P, const es_group* grp, int E, int total_rows, double* sums, void* stream);
int es_bn2d_apply_fwd(const float* x, int C, int P, const float* stats, const float* gamma, const float* beta,
                      long slot_stride, const float* keep_mask, unsigned long long seed, float p_drop, const es_group* grp,
                      int E, int total_rows, float* y, void* stream);
int es_bn2d_bwd_reduce(const float* dy, const float* x, int C, int P, const float* stats, const float* gamma,
                       const float* beta, long slot_stride, const float* keep_mask, unsigned long long seed, float p_drop,
                       const es_group* grp, int E, int total_rows, double* sums2, void* stream);
int es_bn2d_bwd_apply(const float* dy, const float* x, int C, int P, const float* stats, const double* sums2,
                      const float* n_sg, const float* gamma, const float* beta, long slot_stride, const float* keep_mask,
                      unsigned long long seed, float p_drop, const es_group* grp, int E, int total_rows, float* dx,
                      void* stream);

/* ------------------------------------------------------------------------------------------------
 * fused multi-tensor Adam (torch.optim.Adam defaults; expertsim/train/training_setup.py:20-40, stepped at
 * moe.py:439,526,565-566).  One launch updates every parameter of every expert: the parameters of slot s occupy
 * p + s*slot_stride .. + n.  step_count[slot] (device, int32) is advanced only for slots whose grp[s].rows > 0
 * (an expert skipped by the B_e<=1 rule keeps its Adam step).  grp may be NULL (always step).
 * ---------------------------------------------------------------------------------------------- */
int es_adam_step(float* p, const float* g, float* m, float* v, long n, long slot_stride, int slots,
                 float lr, float beta1, float beta2, float eps, int32_t* step_count, const es_group* grp, void* stream);

/* batch inference tail: out[i] = expm1(img[i]) (train/utils.py:201) optionally scattered back to original sample order
 * and widened to float64 as the reference's numpy result. */
int es_expm1_scatter(const float* img, const int32_t* perm, int rows, int HW, double* out_f64, float* out_f32, void* stream);

/* ------------------------------------------------------------------------------------------------
 * evaluation metric front end: sum_channels_parallel (expertsim/train/utils.py:18-78) with the expm1
 * of the inference tail fused in, and the 1-D Wasserstein distance of two equally sized sorted samples
 * (scipy.stats.wasserstein_distance as called at train/utils.py:153-168).
 * ---------------------------------------------------------------------------------------------- */
/* out5[row][0..4] (fp64) = the five channel sums of img[row] ([H,W] fp32; expm1 applied first when apply_expm1 != 0) */
int es_channel_sums(const float* img, int H, int W, int rows, int apply_expm1, double* out5, void* stream);
/* a, b: row-major [n, n_cols] fp64, every column sorted ascending; out[c] = mean_i |a[i,c] - b[i,c]| */
int es_w1_sorted(const double* a, const double* b, int n, int n_cols, double* out, void* stream);

/* ------------------------------------------------------------------------------------------------
 * offline preprocessing: the producers of the `positions` and `std` inputs of the training step.
 * ---------------------------------------------------------------------------------------------- */
/* arg-max pixel of every image as (row, col) — np.unravel_index(np.argmax(img), img.shape), first maximum in row-major
 * order (expertsim/train/utils.py:81-82; notebooks/calculate_and_analysis_of_max_coordinates.ipynb cell 6).
 * img [rows][H*W] fp32; out_rowcol int32 [rows][2] and/or out_rowcol_f32 [rows][2] (the float targets the loader feeds). */
int es_argmax_coords(const float* img, int rows, int H, int W, int32_t* out_rowcol, float* out_rowcol_f32, void* stream);
/* per-condition-group pixel standard deviation (notebooks/calculating_diversity_for_data.ipynb cells 16-23):
 * order[seg[g] .. seg[g+1]) lists the rows of group g (rows with identical conditioning vectors), group_of_row[r] = g.
 * group_sums[g] (fp64, scratch + result) = sum over pixels of the population std (ddof 0) over the group's rows;
 * out_std[r] = group_sums[group_of_row[r]] / max_g group_sums[g]. */
int es_group_pixel_std(const float* img, int rows, int HW, const int32_t* order, const int32_t* seg, int n_groups,
                       const int32_t* group_of_row, double* group_sums, float* out_std, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* EXPERTSIM_B200_H */
