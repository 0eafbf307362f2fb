/*
 * hmvae_b200.h -- C ABI of the B200-native hm-vae hot path (libhmvae_b200.so).
 *
 * Plain pointers and sizes only; every pointer named `d_*` / documented "device" is a CUDA device
 * pointer to contiguous fp32 (or int32) data, 16-byte aligned.  `stream` is a cudaStream_t passed as
 * void*.  Every entry point returns 0 on success, a negative HMVAE_E_* code on argument errors and a
 * positive cudaError_t on CUDA failures; hmvae_last_error() returns a thread-local message.
 * There is NO CPU fallback: a build without the CUDA kernels does not exist.
 *
 * Each entry replaces a PyTorch call chain of the reference (file:line into lijiaman/hm-vae):
 *
 *   hmvae_conv_*            skeleton.py:95-105      SkeletonConv.forward  (weight*mask, F.pad, F.conv1d) + autograd
 *   hmvae_pool_* / unpool   skeleton.py:228-231, 258-261   SkeletonPool/Unpool.forward (matmul by 0/.5/1 matrix)
 *   hmvae_upsample2_*       seq_two_hier_sa_vae.py:233-240 nn.Upsample(x2, linear, align_corners=False)
 *   hmvae_lrelu_*           seq_two_hier_sa_vae.py:129,256 nn.LeakyReLU(0.2)
 *   hmvae_fk_*              fk_layer.py:47-93       ForwardKinematicsLayer.forward (+ my_tools 6D input path)
 *   hmvae_rot6d_*           my_tools.py:19-39       rotation_matrix_from_ortho6d
 *   hmvae_aa2rot_fwd        torchgeometry.angle_axis_to_rotation_matrix (call site seq_two_hier_sa_vae.py:650)
 *   hmvae_latent_*          seq_two_hier_sa_vae.py:419-428  reparametrize + kl_loss
 *   hmvae_recon_fwdbwd      seq_two_hier_sa_vae.py:343, 441-468, 395-411  GT FK, rot6d->R, FK, 3x MSE and their backward
 *   hmvae_mse_*             seq_two_hier_sa_vae.py:430-434  l2_criterion
 *   hmvae_traj_*            trajectory_pred_model.py:289-303, 237-244  gen_motion_w_trajectory + 2x MSE
 *   hmvae_adam_step         trainer_motion_vae.py:29-31, 92-93  torch.optim.Adam(lr, weight_decay) step
 *   hmvae_dp_adam_step      train_motion_vae.py:49-53 (nn.DataParallel) + trainer_motion_vae.py:29-31, 92-93: gradient
 *                           reduce-scatter + Adam + parameter all-gather over NVLink peer memory, one kernel per rank
 *   hmvae_dp_adam_step_units  the same step over a device table of work units that leaves out the always-masked blocks of the
 *                           SkeletonConv weights (skeleton.py:84-96: zero value, zero gradient)
 *   hmvae_linear_*          seq_two_hier_sa_vae.py:132-136, 159-164, 225-229, 267  latent nn.Linear heads
 *   hmvae_batch_assemble,   utils_motion_vae.py:140-187 (MotionSeqData.__getitem__ arithmetic) and :17-57
 *   hmvae_rand_rotation     (rand_rotation_matrix): the batch assembly right before the path
 */
#ifndef HMVAE_B200_H
#define HMVAE_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define HMVAE_E_ARG (-1)     /* bad argument (shape, alignment, unsupported mode) */
#define HMVAE_E_STATE (-2)   /* bad handle */

const char* hmvae_last_error(void);
int hmvae_version(void);
/* Number of kernels this library has launched in the calling process (all threads). */
long long hmvae_launch_count(void);

/* ------------------------------------------------------------------ skeleton-aware conv */

typedef struct hmvae_conv_plan hmvae_conv_plan;

/* Layer geometry + fused prologue/epilogue.  Channel index = joint * channels_per_joint + c on both sides. */
typedef struct {
  int joints;        /* J: edges at this level (in == out) */
  int ci, co;        /* channels per joint, input / output */
  int ksize;         /* K taps */
  int stride;        /* 1 or 2 */
  int pad;           /* zeros/reflect padding on both sides */
  int pad_mode;      /* 0 = zeros ('constant'), 1 = reflect */
  /* prologue (decoder): conv input = unpool(upsample2(src)) -- both optional */
  int upsample;      /* 1: src has T/2 frames, x2 linear upsample (align_corners=False) is applied on the fly */
  int src_joints;    /* joints of src tensor when unpool_src != NULL, else == joints */
  /* epilogue */
  int lrelu;         /* 1: LeakyReLU(0.2) on the output */
  int out_joint_stride; /* channels between consecutive output joints in y (>= co; == co for a plain tensor) */
  int out_chan_offset;  /* first channel inside each output joint block (for writing into a concat buffer) */
  int out_channels_last;/* 1: y is [B, T_out, J*co] instead of [B, J*co, T_out] */
} hmvae_conv_desc;

/* neighbours: CSR (nb_off[J+1], nb_idx[nnz]) host arrays, skeleton.py:34-39.  unpool_src: host [J] or NULL. */
int hmvae_conv_plan_create(const hmvae_conv_desc* desc, const int* nb_off, const int* nb_idx,
                           const int* unpool_src, hmvae_conv_plan** out);
void hmvae_conv_plan_destroy(hmvae_conv_plan* plan);

/* y[B, J*co, T_out] = epilogue(conv1d(pad(prologue(x)), W (.) mask, bias, stride)).
 * x: [B, src_joints*ci, T_src], w: dense [J*co, J*ci, K] (masked entries are never read), bias may be NULL.
 * T is the conv-input length (T_src*2 when upsample).  impl: 0 = auto, 1 = CUDA-core fp32, 2 = tcgen05 TF32. */
int hmvae_conv_fprop(const hmvae_conv_plan* plan, const float* x, const float* w, const float* bias, float* y,
                     int batch, int t_in, int impl, void* stream);
/* dxin[B, J*ci, T]: gradient w.r.t. the *virtual* conv input (after unpool/upsample, before padding); the
 * padding adjoint is folded in.  If the plan has lrelu, dy is first multiplied by lrelu'(y) using y. */
int hmvae_conv_dgrad(const hmvae_conv_plan* plan, const float* dy, const float* y, const float* w, float* dxin,
                     int batch, int t_in, int impl, void* stream);
/* dw: dense [J*co, J*ci, K]; only unmasked blocks are written (masked entries keep their previous value,
 * which must be 0).  dbias may be NULL.  accumulate != 0 adds into dw/dbias. */
int hmvae_conv_wgrad(const hmvae_conv_plan* plan, const float* x, const float* dy, const float* y, float* dw,
                     float* dbias, int batch, int t_in, int accumulate, int impl, void* stream);
/* ---- tensor-core (tcgen05, TF32 operands, FP32 accumulation in TMEM) variants.  They read a packed, tf32-rounded
 * copy of the weights; refresh it with hmvae_conv_pack_weights whenever the dense parameter changes.
 * mode: 0 = fprop, 1 = dgrad.  hmvae_conv_tc_supported returns 1 when (plan, batch, t_in) maps onto the kernel. */
int hmvae_conv_tc_supported(const hmvae_conv_plan* plan, int batch, int t_in, int mode);
int hmvae_conv_packed_size(const hmvae_conv_plan* plan, long* n_fprop, long* n_dgrad);   /* floats */
int hmvae_conv_pack_weights(const hmvae_conv_plan* plan, const float* w, float* wp_fprop, float* wp_dgrad, void* stream);
/* workspace: device scratch for the staged (padded / upsampled / tf32-rounded) activation tiles, at least
 * hmvae_conv_tc_workspace(plan, batch, t_in, mode) bytes, 16-byte aligned; contents are dead after the call returns
 * (stream-ordered). */
long hmvae_conv_tc_workspace(const hmvae_conv_plan* plan, int batch, int t_in, int mode);
int hmvae_conv_fprop_tc(const hmvae_conv_plan* plan, const float* x, const float* wp_fprop, const float* bias, float* y,
                        int batch, int t_in, void* workspace, long workspace_bytes, void* stream);
int hmvae_conv_dgrad_tc(const hmvae_conv_plan* plan, const float* dy, const float* y, const float* wp_dgrad, float* dxin,
                        int batch, int t_in, void* workspace, long workspace_bytes, void* stream);
/* ---- stack-level path: the three phases of a tensor-core conv as separate calls, and the inter-layer link kernel.
 *
 * hmvae_conv_tc_stage  : the staging pass only (padding / upsample / unpool gather, or zero insertion + LeakyReLU' for dgrad).
 * hmvae_conv_tc_run    : the tcgen05 kernel only: staged tiles -> raw accumulator dump (split-K partials, no bias).
 * hmvae_conv_tc_finish : dump -> result tensor (split-K sum, bias, LeakyReLU / reflect fold), as hmvae_conv_{fprop,dgrad}_tc do.
 * hmvae_conv_tc_sizes  : bytes of the staging buffer and of the dump for (plan, mode, batch, t_in); returns 0 if unsupported.
 * mode: 0 fprop, 1 dgrad.  hmvae_conv_{fprop,dgrad}_tc == stage + run + finish on one workspace. */
int hmvae_conv_tc_sizes(const hmvae_conv_plan* plan, int batch, int t_in, int mode, long* stage_bytes, long* dump_bytes);
int hmvae_conv_tc_stage(const hmvae_conv_plan* plan, int mode, const float* src, const float* yact, int batch, int t_in,
                        void* stage_ws, void* stream);
int hmvae_conv_tc_run(const hmvae_conv_plan* plan, int mode, const float* wp, int batch, int t_in, const void* stage_ws,
                      void* dump_ws, void* stream);
int hmvae_conv_tc_finish(const hmvae_conv_plan* plan, int mode, const void* dump_ws, const float* bias, float* dst, int batch,
                         int t_in, void* stream);

/* hmvae_conv_link: ONE kernel for everything between two convs of a stack (replaces finish of the producer, SkeletonPool +
 * LeakyReLU / nn.Upsample + SkeletonUnpool + the last decoder level's per-edge concat -- seq_two_hier_sa_vae.py:120-130,
 * 233-258, 278-288 -- or their adjoints, and the staging pass of the consumer).  It reads the producer's dump, writes the
 * boundary tensor S (NCW, [batch, s_joints * s_cpj, s_t]) and, when `cons` is given, the consumer's staged tiles.
 *   kind 0 (forward)  : S = act(mean over pool members of (conv_P + bias)); `aux` [batch, s_joints*(ojs_P - co_P), s_t] fills the
 *                       remaining channels of every joint when the producer plan has out_joint_stride > co (concat).
 *                       pool_off / pool_idx: HOST CSR pooled edge -> producer joints (NULL: identity), pool_joints rows.
 *                       cons = the next conv (its upsample / unpool / padding are applied on the fly), mode fprop.
 *   kind 1 (backward, decoder side): producer = dgrad of conv P; S = gradient of P's source tensor (adjoint of P's upsample /
 *                       unpool), [batch, src_joints_P * ci_P, T_src].  cons = the previous conv (mode dgrad); yact_c = its
 *                       activated output in the layout of S when it fuses LeakyReLU.
 *   kind 2 (backward, encoder side): producer = dgrad of conv P whose input is pool(+LeakyReLU) of conv C's output;
 *                       S = gradient of C's raw output = pool^T(lrelu'(sact) * (dgrad_P + add)); sact / add: [batch, J_P*ci_P, T_P].
 * s_out may be NULL when a consumer is given (inference: nothing needs the boundary tensor again).
 * stage_ws: the consumer's staging buffer -- PERSISTENT and ZERO-INITIALISED by the caller (hmvae_conv_tc_sizes bytes): the kernel
 * writes only real values, padding rows / channels and zero-inserted positions must stay 0.
 * hmvae_conv_link_supported: 1 if the descriptor's geometry can take this path (buffers may be NULL). */
typedef struct {
  int kind, batch;
  const hmvae_conv_plan* prod;
  int prod_t;                 /* t_in of the producer conv */
  const hmvae_conv_plan* cons;
  int cons_t;                 /* t_in of the consumer conv */
  int act;
  int pool_joints;
  const int* pool_off;
  const int* pool_idx;
  const float* dump;
  const float* bias;
  const float* aux;
  const float* add;
  const float* sact;
  const float* yact_c;
  float* s_out;
  void* stage_ws;
} hmvae_conv_link_desc;
int hmvae_conv_link_supported(const hmvae_conv_link_desc* desc);
int hmvae_conv_link(const hmvae_conv_link_desc* desc, void* stream);

/* Debug aid (tools/tc_phases.py): device buffer (8 x uint64 per CTA) that receives %globaltimer stamps from the following
 * hmvae_conv_{fprop,dgrad}_tc launches; NULL switches it off. */
int hmvae_conv_tc_debug(void* buf);

/* Weight gradient on the tensor cores.  Only unmasked blocks of dw are written (deterministically, one writer per element);
 * masked entries keep their previous value, which must be 0.  accumulate != 0 adds into dw / dbias. */
int hmvae_conv_wgrad_tc_supported(const hmvae_conv_plan* plan, int batch, int t_in);
long hmvae_conv_wgrad_tc_workspace(const hmvae_conv_plan* plan, int batch, int t_in);
int hmvae_conv_wgrad_tc(const hmvae_conv_plan* plan, const float* x, const float* dy, const float* y, float* dw, float* dbias,
                        int batch, int t_in, int accumulate, void* workspace, long workspace_bytes, void* stream);
/* The x operand of the weight gradient (the conv's source tensor, skeleton.py:95-105's `input`) depends on forward data only:
 * this stages its TF32 tiles into `workspace` ahead of time (during the forward pass); the backward pass then calls
 * hmvae_conv_wgrad_tc with x == NULL and the SAME workspace, and only dy is staged there. */
int hmvae_conv_wgrad_tc_stage_x(const hmvae_conv_plan* plan, const float* x, int batch, int t_in, void* workspace,
                                long workspace_bytes, void* stream);

/* adjoint of the prologue: dsrc[B, src_joints*ci, T_src] from dxin[B, J*ci, T]; if src_act != NULL the result is
 * multiplied by lrelu'(src_act) (src_act = the activation tensor that fed this layer). */
int hmvae_conv_prologue_bwd(const hmvae_conv_plan* plan, const float* dxin, const float* src_act, float* dsrc,
                            int batch, int t_in, void* stream);

/* ------------------------------------------------------------------ pool / unpool / upsample / lrelu */

/* pooling table: CSR over output edges (host arrays, <= 64 edges).  y = mean over members (skeleton.py:219-226). */
int hmvae_pool_fwd(const float* x, float* y, int batch, int in_edges, int out_edges, int c, int t,
                   const int* pool_off, const int* pool_idx, int lrelu, void* stream);
/* dx = pool^T(dy (.) lrelu'(y)) ; y may be NULL when lrelu == 0 */
int hmvae_pool_bwd(const float* dy, const float* y, float* dx, int batch, int in_edges, int out_edges, int c, int t,
                   const int* pool_off, const int* pool_idx, int lrelu, void* stream);
/* unpool: y[:, j*c + ch] = x[:, src[j]*c + ch] (skeleton.py:248-256); bwd sums over members. */
int hmvae_unpool_fwd(const float* x, float* y, int batch, int in_edges, int out_edges, int c, int t, const int* src,
                     void* stream);
int hmvae_unpool_bwd(const float* dy, float* dx, int batch, int in_edges, int out_edges, int c, int t, const int* src,
                     void* stream);
int hmvae_upsample2_fwd(const float* x, float* y, long rows, int t, void* stream);
int hmvae_upsample2_bwd(const float* dy, float* dx, long rows, int t, void* stream);
int hmvae_lrelu_fwd(const float* x, float* y, long n, float slope, void* stream);
int hmvae_lrelu_bwd(const float* dy, const float* y, float* dx, long n, float slope, void* stream);
/* [B, C, T] <-> [B, T, C] */
int hmvae_transpose_ct(const float* x, float* y, int batch, int c, int t, void* stream);

/* ------------------------------------------------------------------ rotations / forward kinematics */

/* rot: [N, J, 3, 3] (rot_dim 9) or [N, J, 6] (rot_dim 6).  offsets: device [J,3]; positions: device [N,J,3] or NULL
 * (fk_layer.py:82-89).  parents: HOST int[J], parents[0] ignored, parents[i] < i required.  pos: [N, J, 3].
 * rotmat_out (optional, rot_dim == 6 only): [N, J, 3, 3] = rot6d->R of the input. */
int hmvae_fk_fwd(const float* rot, int rot_dim, const float* offsets, const float* positions, const int* parents,
                 int joints, long n, float* pos, float* rotmat_out, void* stream);
/* drot: [N, J, rot_dim].  drotmat_extra (optional, rot_dim == 6): extra upstream gradient on R, [N,J,3,3]. */
int hmvae_fk_bwd(const float* rot, int rot_dim, const float* offsets, const float* positions, const int* parents,
                 int joints, long n, const float* dpos, const float* drotmat_extra, float* drot, void* stream);
int hmvae_rot6d_fwd(const float* x6, float* rotmat, long m, void* stream);
int hmvae_rot6d_bwd(const float* x6, const float* drotmat, float* dx6, long m, void* stream);
/* [M,3] -> [M,4,4] homogeneous (torchgeometry semantics, small-angle Taylor branch at theta^2 <= 1e-6) */
int hmvae_aa2rot_fwd(const float* aa, float* out44, long m, void* stream);

/* ------------------------------------------------------------------ fused VAE losses */

/* z = eps*exp(0.5*lv)+mu, kl_sum += sum_rows(-0.5*sum_d(1+lv-mu^2-exp(lv))).  dist: [rows, 2*d] = (mu|lv).
 * eps may be NULL (z = mu).  kl_out: device float[1], ATOMICALLY accumulated: caller zeroes it. */
int hmvae_latent_fwd(const float* dist, const float* eps, float* z, float* kl_out, long rows, int d, void* stream);
/* ddist = [dz + s*mu | dz*eps*0.5*exp(0.5 lv) + s*0.5*(exp(lv)-1)],  s = kl_scale * (dkl ? *dkl : 1).
 * kl_scale = kl_w/rows for a fused loss; dkl is an optional DEVICE scalar (upstream gradient of the KL sum).
 * dz may be NULL (detached z). */
int hmvae_latent_bwd(const float* dist, const float* eps, const float* dz, const float* dkl, float* ddist, long rows,
                     int d, float kl_scale, void* stream);

/* The latent bottleneck of the hierarchy, up to 4 levels per launch (seq_two_hier_sa_vae.py:159-164 encoder heads `nn.Linear`,
 * :357-391 / :419-428 reparametrise + kl_loss, :225-229 / :267 decoder heads `nn.Linear`), one row per (sequence, edge):
 *   fwd: dist = x enc_w^T + enc_b;  z = eps*exp(lv/2)+mu ((mu|lv) = dist; eps NULL: z = mu);  kl_acc[0] += sum of KL row sums
 *        (atomic; caller zeroes);  feat = z dec_w^T + dec_b.
 *   bwd: gz = gfeat dec_w;  gdist = [gz + s*mu | gz*eps*exp(lv/2)/2 + s*(exp(lv)-1)/2] with s = kl_scale;  gx = gdist enc_w.
 * Weight gradients are NOT produced here: use hmvae_linear_bwd on (x, gdist) and (z, gfeat).
 * x / feat / gfeat / gx: [rows, features]; dist / gdist: [rows, 2d]; z / eps: [rows, d]; enc_w [2d, features]; dec_w [features, d]. */
typedef struct {
  int rows, features, d;
  float kl_scale;
  const float* x;
  const float* enc_w;
  const float* enc_b;
  const float* eps;
  const float* dec_w;
  const float* dec_b;
  float* dist;
  float* z;
  float* feat;
  float* kl_acc;
  const float* gfeat;
  float* gdist;
  float* gx;
  long gfeat_stride;   /* floats between consecutive rows of gfeat; 0 = features (contiguous) */
} hmvae_head_level;
int hmvae_latent_heads_fwd(const hmvae_head_level* levels, int n_levels, void* stream);
int hmvae_latent_heads_bwd(const hmvae_head_level* levels, int n_levels, void* stream);

/* One kernel for: GT FK (no grad), rot6d->R, FK on the prediction, the three MSEs and d(total)/d(x6_pred).
 * x6_pred: decoder output, NCW [B, 24*6, T] (ncw=1) or [B, T, 24*6] (ncw=0); gt_6d [B,T,144]; gt_rotmat [B,T,216].
 * losses: device float[4], ATOMICALLY accumulated (caller zeroes): {sum sq 6d, sum sq rot, sum sq pos, unused}.
 * dx6 (may be NULL for validation): same layout as x6_pred, = w6d*2(d)/n6 + ... (weights already divided by the
 * element counts by the caller: s6 = 2*w_6d/numel etc).  pos_pred_out / gt_pos_out optional [B,T,72]. */
int hmvae_recon_fwdbwd(const float* x6_pred, int ncw, const float* gt_6d, const float* gt_rotmat,
                       const float* offsets, const int* parents, int joints, int batch, int t,
                       float s6, float srot, float spos, float* losses, float* dx6, float* pos_pred_out,
                       float* gt_pos_out, void* stream);
/* Same kernel with a per-(frame, joint) weight on the squared errors: l2_masked_criterion, seq_two_hier_sa_vae.py:717-735
 * (mask [B, T, joints], the mean still runs over ALL elements), as used by the latent-space optimisation loops
 * (:1356-1429, :1698-1757).  mask may be NULL (= hmvae_recon_fwdbwd).  rot_pred_out: optional [B,T,joints,3,3]. */
int hmvae_recon_masked_fwdbwd(const float* x6_pred, int ncw, const float* gt_6d, const float* gt_rotmat, const float* mask,
                              const float* offsets, const int* parents, int joints, int batch, int t, float s6, float srot,
                              float spos, float* losses, float* dx6, float* pos_pred_out, float* gt_pos_out,
                              float* rot_pred_out, void* stream);
/* Turns the atomically accumulated sums into loss values on the device, without a host sync:
 * out[i] = acc[i]*scale[i] (i<n), out[n] = sum w[i]*out[i] (total), out[n+1] = sum wk[i]*out[i] (weighted KL); acc is zeroed
 * for the next step.  scale / w / wk are HOST arrays of n <= 8 floats. */
int hmvae_loss_finalize(float* acc, float* out, const float* scale, const float* w, const float* wk, int n, void* stream);
/* sum((a-b)^2) atomically added to out[0]; grad: da = scale*(a-b) */
int hmvae_mse_fwd(const float* a, const float* b, float* out, long n, void* stream);
int hmvae_mse_bwd(const float* a, const float* b, float* da, long n, float scale, void* stream);

/* trajectory accumulation (prefix sum over T of de-standardised root velocity, added to every joint) + its MSE.
 * root_v_pred/gt: [B,T,3] standardised; mean/std: 3 floats each (host); returns in losses[0] sum sq of
 * (root_v_pred-root_v_gt), losses[1] sum sq of accumulated trajectories difference *per joint-coordinate*
 * (already multiplied by the 24 joints); d_root_v = gradient of  sv*L_v + st*L_trans  (sv, st pre-divided). */
int hmvae_traj_fwdbwd(const float* root_v_pred, const float* root_v_gt, const float* mean3, const float* std3,
                      int batch, int t, int joints, float sv, float st, float* losses, float* d_root_v,
                      void* stream);

/* ------------------------------------------------------------------ latent heads (nn.Linear) */

/* y[rows, out_f] = x[rows, in_f] w[out_f, in_f]^T + bias   (seq_two_hier_sa_vae.py:162, 267); fp32 CUDA cores. */
int hmvae_linear_fwd(const float* x, const float* w, const float* bias, float* y, int rows, int in_f, int out_f, void* stream);
/* dx = dy w, dw = dy^T x, db = column sums of dy; any of dx / dw / db may be NULL. */
int hmvae_linear_bwd(const float* x, const float* w, const float* dy, float* dx, float* dw, float* db, int rows, int in_f,
                     int out_f, void* stream);

/* ------------------------------------------------------------------ optimiser */

/* torch.optim.Adam semantics (L2 weight decay added to the gradient, bias correction, eps outside sqrt):
 *   g += wd*p; m = b1*m+(1-b1)*g; v = b2*v+(1-b2)*g*g; p -= lr/(1-b1^t) * m / (sqrt(v)/sqrt(1-b2^t) + eps)
 * tensors: HOST array of n_tensors {p,g,m,v device pointers, numel}; grad_scale multiplies g first (1/world). */
typedef struct {
  float* p;
  const float* g;
  float* m;
  float* v;
  long numel;
} hmvae_adam_tensor;
int hmvae_adam_step(const hmvae_adam_tensor* tensors, int n_tensors, float lr, double beta1, double beta2, float eps,
                    float weight_decay, int step, float grad_scale, void* stream);
/* Same, but the two step-dependent scalars {lr/(1-b1^t), 1/sqrt(1-b2^t)} are read from DEVICE memory (float[2]) so that
 * a CUDA graph of the whole training step can be replayed while the host advances t and the LR schedule. */
int hmvae_adam_step_dyn(const hmvae_adam_tensor* tensors, int n_tensors, const float* dyn2, double beta1, double beta2,
                        float eps, float weight_decay, float grad_scale, void* stream);

/* Decoder-weight L2 regulariser of the latent-space optimisation loops (seq_two_hier_sa_vae.py:1382-1387, 1717-1722; the
 * per-parameter l2_criterion(params, self.dec.state_dict()[name]) loop): loss[0] += sum_i mean((p_i - p0_i)^2) (atomic; caller
 * zeroes) and, where g_i is given, g_i (+)= weight * 2 (p_i - p0_i) / numel_i.  accumulate != 0: add to the gradient the
 * backward pass wrote; 0: store (parameters the backward pass does not reach).  One launch per 64 tensors. */
typedef struct {
  const float* p;
  const float* p0;
  float* g;
  long numel;
  int accumulate;
} hmvae_reg_tensor;
int hmvae_l2_reg_fwdbwd(const hmvae_reg_tensor* tensors, int n_tensors, float weight, float* loss, void* stream);

/* The step-dependent scalars produced ON THE DEVICE (replaces torch.optim.Adam's host-side bias corrections and
 * torch.optim.lr_scheduler.StepLR, trainer_motion_vae.py:29-33, 251-262): clock = device uint32[2] {completed Adam steps t,
 * scheduler iterations}; one call advances both and writes dyn2 = {lr_t / (1 - b1^t), 1 / sqrt(1 - b2^t)} with
 * lr_t = base_lr * gamma^(iterations / step_size) (step_size <= 0: constant).  Capturable: a replayed CUDA graph of the step
 * keeps its own time, whatever the host has queued ahead. */
int hmvae_opt_clock_tick(unsigned int* clock, float base_lr, float gamma, int step_size, double beta1, double beta2,
                         float* dyn2, void* stream);

/* ------------------------------------------------------------------ batch assembly (utils_motion_vae.py, the step before the path)
 *
 * rand_rotation_matrix (:17-57): randnums [n,3] float64 in [0,1] -> rot [n,9] float32 (row-major), computed in float64. */
int hmvae_rand_rotation(const double* randnums, double deflection, float* rot, long n, void* stream);
/* MotionSeqData.__getitem__ (:140-187) for a batch of already cropped windows raw [batch, t, 579]: the seven training tensors
 * rot6d [B,T,144], rotmat [B,T,216], rot_pos [B,T,72] (raw), joint_pos / linear_v / angular_v [B,T,72] and root_v [B,T,3]
 * (standardised with mean / std, float64 [579], zero stds already replaced by 1); any output may be NULL.  root_rot [batch, 9]
 * (one rotation per sequence, hmvae_rand_rotation) switches on the random-root-rotation augmentation (:167-185): root matrix
 * and root velocity are rotated, the 6D representation is re-derived from the matrices; NULL = no augmentation. */
int hmvae_batch_assemble(const float* raw, const float* root_rot, const double* mean, const double* stdv, int batch, int t,
                         float* rot6d, float* rotmat, float* rot_pos, float* joint_pos, float* linear_v, float* angular_v,
                         float* root_v, void* stream);

/* ------------------------------------------------------------------ data parallelism (train_motion_vae.py:49-53)
 *
 * The optimiser step fused with its collective over NVLink peer memory: every rank owns a share of the parameter elements;
 * for those it reads the gradient from EVERY rank's gradient arena (peer loads), applies Adam with its local slice of m / v,
 * and stores the new value into EVERY rank's parameter arena (peer stores) = reduce-scatter + sharded Adam + all-gather in one
 * kernel, ordered by two flag barriers in peer memory (no NCCL, no host sync, CUDA-graph capturable).
 *   peers  : device-accessible base pointers of all ranks' arenas (symmetric layout: same offset = same parameter element);
 *            flags[q] = >= 2*world zero-initialised uint32 in rank q's memory.
 *   ranges : HOST array [nranges][2] of element ranges [begin, end) owned by THIS rank (multiples of 4); m / v: this rank's
 *            full-size moment arenas (only the owned ranges are touched).   dyn2: device float[2] as in hmvae_adam_step_dyn.
 *   state  : device uint32[4], zero-initialised once: {epoch, CTA counter, timeout flag, -}.
 * max_ctas : 0 = size the grid for the whole GPU; > 0 caps it (a call issued in the middle of the backward pass, for the
 *            parameters whose gradients are already final, should leave SMs to the kernels it overlaps with).
 * Every rank must make the same sequence of calls.  world == 1 degenerates to a plain fused Adam over the ranges. */
#define HMVAE_DP_MAX_WORLD 8
#define HMVAE_DP_MAX_RANGES 64
typedef struct {
  int world, rank;
  const float* grad[HMVAE_DP_MAX_WORLD];
  float* param[HMVAE_DP_MAX_WORLD];
  unsigned int* flags[HMVAE_DP_MAX_WORLD];
  const float* mc_grad;   /* optional NVSwitch multicast (NVLS) mappings of the gradient / parameter arenas: when both are set the */
  float* mc_param;        /* kernel uses multimem.ld_reduce / multimem.st (in-switch reduction and replication); else NULL          */
} hmvae_dp_peers;
int hmvae_dp_adam_step(const hmvae_dp_peers* peers, float* m, float* v, const long* ranges, int nranges, const float* dyn2,
                       double beta1, double beta2, float eps, float weight_decay, float grad_scale, unsigned int* state,
                       int max_ctas, void* stream);
/* Same step over a DEVICE table of work units instead of host ranges -- the mask-aware variant: the masked
 * (joint, non-neighbour joint) blocks of a SkeletonConv weight are zero at initialisation (skeleton.py:84-93) and receive a
 * zero gradient on every step (the forward multiplies by the mask, skeleton.py:96), so torch.optim.Adam never moves them; the
 * table simply does not list them and neither the reduce-scatter, the optimiser nor the all-gather streams those bytes.
 *   units : device int32 [nunits][2] = {first float4 of the unit, number of float4 (1..32)}, owned by THIS rank, pairwise
 *           disjoint over all ranks (one warp handles one unit per step; 8-byte aligned).
 *   loads_in_flight : units per warp and iteration (1, 2, 4 or 8; 0 = default: 1 on one rank, 2 across ranks).  A call capped
 *           to few CTAs hides the NVLink latency with more loads per thread instead of more threads. */
int hmvae_dp_adam_step_units(const hmvae_dp_peers* peers, float* m, float* v, const int* units, long nunits, const float* dyn2,
                             double beta1, double beta2, float eps, float weight_decay, float grad_scale, unsigned int* state,
                             int max_ctas, int loads_in_flight, void* stream);
/* Peer-memory plumbing over CUDA IPC (used when torch's symmetric memory is unavailable): zero-filled device allocation, its
 * 64-byte handle, and mapping / unmapping of another process's handle. */
int hmvae_ipc_alloc(long bytes, void** ptr);
int hmvae_ipc_free(void* ptr);
int hmvae_ipc_get_handle(void* ptr, unsigned char* handle64);
int hmvae_ipc_open_handle(const unsigned char* handle64, void** ptr);
int hmvae_ipc_close_handle(void* ptr);

#ifdef __cplusplus
}
#endif
#endif
