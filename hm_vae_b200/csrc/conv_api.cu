// C ABI for the skeleton-aware conv: plan construction (index tables) and implementation dispatch.
#include <string.h>

#include <stdlib.h>

#include "conv_common.cuh"

namespace hmvae {
thread_local char g_err[512] = "";
std::atomic<long long> g_launches{0};

int env_int(const char* name, int dflt) {
  const char* e = getenv(name);
  return (e && e[0]) ? atoi(e) : dflt;
}

int pdl_level() {
  // Measured on B200 (profiles/r01_summary_v3.md): back-to-back eager launches gain ~12 % device time with level 1, but inside the
  // step's CUDA graph early-launched 200 KB-smem CTAs sit on SMs the other stream's kernels want (1.217 vs 1.205 ms) => opt-in.
  static int lvl = -1;
  if (lvl < 0) lvl = env_int("HMVAE_PDL", 0);
  return lvl;
}

int num_sms() {
  static int cached = 0;
  if (cached == 0) {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0)
      cached = n;
    else
      cached = 148;
  }
  return cached;
}

bool conv_tc_supported(const hmvae_conv_plan* plan, int B, int T, int mode);
void conv_packed_sizes(const hmvae_conv_plan* plan, long* n_fprop, long* n_dgrad);
int conv_pack(const hmvae_conv_plan* plan, const float* w, float* wp_f, float* wp_d, cudaStream_t st);
int conv_tc_launch(const hmvae_conv_plan* plan, int mode, const float* src, const float* yact, const float* wp,
                   const float* bias, float* dst, int B, int T, void* workspace, long workspace_bytes, cudaStream_t st);
long conv_tc_workspace_bytes(const hmvae_conv_plan* plan, int B, int T, int mode);
bool conv_tc_sizes(const hmvae_conv_plan* plan, int B, int T, int mode, long* stage_bytes, long* dump_bytes);
int conv_tc_stage(const hmvae_conv_plan* plan, int mode, const float* src, const float* yact, int B, int T, void* stage_ws,
                  cudaStream_t st);
int conv_tc_run(const hmvae_conv_plan* plan, int mode, const float* wp, int B, int T, const void* stage_ws, void* dump_ws,
                cudaStream_t st);
int conv_tc_finish(const hmvae_conv_plan* plan, int mode, const void* dump_ws, const float* bias, float* dst, int B, int T,
                   cudaStream_t st);
void conv_tc_release(const hmvae_conv_plan* plan);
bool conv_wgrad_tc_supported(const hmvae_conv_plan* plan, int B, int T);
long conv_wgrad_tc_workspace_bytes(const hmvae_conv_plan* plan, int B, int T);
int conv_wgrad_tc_stage_x(const hmvae_conv_plan* plan, const float* x, int B, int T, void* workspace, long workspace_bytes,
                          cudaStream_t st);
int conv_wgrad_tc_launch(const hmvae_conv_plan* plan, const float* x, const float* dy, const float* yact, float* dw,
                         float* dbias, int B, int T, int accumulate, void* workspace, long workspace_bytes, cudaStream_t st);
}  // namespace hmvae

using namespace hmvae;

extern "C" const char* hmvae_last_error(void) { return g_err; }
extern "C" int hmvae_version(void) { return 100; }
extern "C" long long hmvae_launch_count(void) { return g_launches.load(); }

extern "C" int hmvae_conv_plan_create(const hmvae_conv_desc* desc, const int* nb_off, const int* nb_idx,
                                      const int* unpool_src, hmvae_conv_plan** out) {
  if (!desc || !nb_off || !nb_idx || !out) return fail_arg("conv_plan_create: null pointer");
  const hmvae_conv_desc& d = *desc;
  if (d.joints < 1 || d.joints > 64 || d.ci < 1 || d.co < 1) return fail_arg("conv_plan_create: bad joints/channels");
  if (d.ksize < 1 || d.ksize > 32) return fail_arg("conv_plan_create: kernel_size must be in [1, 32]");
  if (d.stride < 1 || d.stride > 2) return fail_arg("conv_plan_create: stride must be 1 or 2");
  if (d.pad < 0 || (d.pad_mode != 0 && d.pad_mode != 1)) return fail_arg("conv_plan_create: bad padding");
  const int J = d.joints, nnz = nb_off[J];
  if (nb_off[0] != 0 || nnz < 0 || nnz > J * J) return fail_arg("conv_plan_create: bad neighbour CSR");
  const int src_J = unpool_src ? d.src_joints : J;
  if (src_J < 1 || src_J > 64) return fail_arg("conv_plan_create: bad src_joints");
  static std::atomic<unsigned long long> next_uid{1};
  hmvae_conv_plan* p = new hmvae_conv_plan();
  p->uid = next_uid.fetch_add(1);
  p->d = d;
  p->nb_off.assign(nb_off, nb_off + J + 1);
  p->nb_idx.assign(nb_idx, nb_idx + nnz);
  p->src.resize(J);
  p->max_nb = 0;
  for (int j = 0; j < J; ++j) {
    p->src[j] = unpool_src ? unpool_src[j] : j;
    if (p->src[j] < 0 || p->src[j] >= src_J) { delete p; return fail_arg("conv_plan_create: unpool source out of range"); }
    if (nb_off[j + 1] < nb_off[j]) { delete p; return fail_arg("conv_plan_create: bad neighbour CSR"); }
    if (nb_off[j + 1] - nb_off[j] > p->max_nb) p->max_nb = nb_off[j + 1] - nb_off[j];
    for (int m = nb_off[j]; m < nb_off[j + 1]; ++m)
      if (nb_idx[m] < 0 || nb_idx[m] >= J) { delete p; return fail_arg("conv_plan_create: neighbour out of range"); }
  }
  // transpose adjacency + block list
  std::vector<int> tOff(J + 1, 0), tIdx(nnz), bj(nnz), bn(nnz);
  for (int m = 0; m < nnz; ++m) tOff[nb_idx[m] + 1]++;
  for (int j = 0; j < J; ++j) tOff[j + 1] += tOff[j];
  std::vector<int> fill(J, 0);
  for (int j = 0; j < J; ++j)
    for (int m = nb_off[j]; m < nb_off[j + 1]; ++m) {
      const int n = nb_idx[m];
      tIdx[tOff[n] + fill[n]++] = j;
      bj[m] = j;
      bn[m] = n;
    }
  std::vector<int> host;
  auto push = [&](const std::vector<int>& v) { size_t o = host.size(); host.insert(host.end(), v.begin(), v.end()); while (host.size() % 4) host.push_back(0); return o; };
  const size_t o_nb_off = push(p->nb_off), o_nb_idx = push(p->nb_idx), o_t_off = push(tOff), o_t_idx = push(tIdx);
  const size_t o_bj = push(bj), o_bn = push(bn), o_src = push(p->src);
  cudaError_t e = cudaMalloc(&p->dev_tables, host.size() * sizeof(int));
  if (e == cudaSuccess) e = cudaMemcpy(p->dev_tables, host.data(), host.size() * sizeof(int), cudaMemcpyHostToDevice);
  if (e != cudaSuccess) {
    snprintf(g_err, sizeof(g_err), "hmvae: conv_plan_create: %s", cudaGetErrorString(e));
    delete p;
    return (int)e;
  }
  ConvArgs& a = p->a;
  a.J = J; a.ci = d.ci; a.co = d.co; a.K = d.ksize; a.s = d.stride; a.p = d.pad; a.pad_mode = d.pad_mode;
  a.upsample = d.upsample ? 1 : 0; a.src_J = src_J; a.lrelu = d.lrelu ? 1 : 0;
  a.ojs = d.out_joint_stride > 0 ? d.out_joint_stride : d.co;
  a.oco = d.out_chan_offset; a.cl = d.out_channels_last ? 1 : 0; a.nnz = nnz;
  if (a.ojs < a.oco + a.co) { cudaFree(p->dev_tables); delete p; return fail_arg("conv_plan_create: out_joint_stride too small"); }
  a.nb_off = p->dev_tables + o_nb_off; a.nb_idx = p->dev_tables + o_nb_idx;
  a.nbT_off = p->dev_tables + o_t_off; a.nbT_idx = p->dev_tables + o_t_idx;
  a.blk_j = p->dev_tables + o_bj; a.blk_n = p->dev_tables + o_bn; a.src = p->dev_tables + o_src;
  *out = p;
  return 0;
}

extern "C" void hmvae_conv_plan_destroy(hmvae_conv_plan* plan) {
  if (!plan) return;
  cudaFree(plan->dev_tables);
  conv_tc_release(plan);
  for (auto& kv : plan->tc_tables) cudaFree(kv.second);
  delete plan;
}

static int check_shape(const hmvae_conv_plan* plan, int batch, int t_in, const char* who) {
  if (!plan) { snprintf(g_err, sizeof(g_err), "hmvae: %s: null plan", who); return HMVAE_E_STATE; }
  const hmvae_conv_desc& d = plan->d;
  if (batch < 0 || t_in < 1) { snprintf(g_err, sizeof(g_err), "hmvae: %s: bad batch/time", who); return HMVAE_E_ARG; }
  if (d.upsample && (t_in & 1)) { snprintf(g_err, sizeof(g_err), "hmvae: %s: upsampled length must be even", who); return HMVAE_E_ARG; }
  if (d.pad_mode == 1 && d.pad > t_in - 1) {
    snprintf(g_err, sizeof(g_err), "hmvae: %s: reflect padding %d needs an input of at least %d frames (got %d)", who, d.pad, d.pad + 1, t_in);
    return HMVAE_E_ARG;
  }
  if (t_in + 2 * d.pad < d.ksize) { snprintf(g_err, sizeof(g_err), "hmvae: %s: input shorter than the kernel", who); return HMVAE_E_ARG; }
  return 0;
}

extern "C" int hmvae_conv_fprop(const hmvae_conv_plan* plan, const float* x, const float* w, const float* bias, float* y,
                                int batch, int t_in, int impl, void* stream) {
  int rc = check_shape(plan, batch, t_in, "conv_fprop");
  if (rc) return rc;
  if (!x || !w || !y) return fail_arg("conv_fprop: null pointer");
  if (batch == 0) return 0;
  (void)impl;
  return conv_fprop_simt(plan, x, w, bias, y, batch, t_in, (cudaStream_t)stream);
}

extern "C" int hmvae_conv_tc_supported(const hmvae_conv_plan* plan, int batch, int t_in, int mode) {
  if (!plan || batch < 1 || t_in < 1 || (mode != 0 && mode != 1)) return 0;
  if (check_shape(plan, batch, t_in, "conv_tc_supported")) return 0;
  return conv_tc_supported(plan, batch, t_in, mode) ? 1 : 0;
}

extern "C" int hmvae_conv_packed_size(const hmvae_conv_plan* plan, long* n_fprop, long* n_dgrad) {
  if (!plan || !n_fprop || !n_dgrad) return fail_arg("conv_packed_size: null pointer");
  conv_packed_sizes(plan, n_fprop, n_dgrad);
  return 0;
}

extern "C" int hmvae_conv_pack_weights(const hmvae_conv_plan* plan, const float* w, float* wp_fprop, float* wp_dgrad,
                                       void* stream) {
  if (!plan || !w) return fail_arg("conv_pack_weights: null pointer");
  return conv_pack(plan, w, wp_fprop, wp_dgrad, (cudaStream_t)stream);
}

extern "C" long hmvae_conv_tc_workspace(const hmvae_conv_plan* plan, int batch, int t_in, int mode) {
  if (!plan || batch < 1 || t_in < 1 || (mode != 0 && mode != 1)) return -1;
  return conv_tc_workspace_bytes(plan, batch, t_in, mode);
}

extern "C" int hmvae_conv_tc_sizes(const hmvae_conv_plan* plan, int batch, int t_in, int mode, long* stage_bytes,
                                   long* dump_bytes) {
  if (!plan || !stage_bytes || !dump_bytes || batch < 1 || t_in < 1 || (mode != 0 && mode != 1)) return 0;
  if (check_shape(plan, batch, t_in, "conv_tc_sizes")) return 0;
  return conv_tc_sizes(plan, batch, t_in, mode, stage_bytes, dump_bytes) ? 1 : 0;
}

extern "C" int hmvae_conv_tc_stage(const hmvae_conv_plan* plan, int mode, const float* src, const float* yact, int batch,
                                   int t_in, void* stage_ws, void* stream) {
  int rc = check_shape(plan, batch, t_in, "conv_tc_stage");
  if (rc) return rc;
  if (!src || (mode != 0 && mode != 1) || (mode == 1 && plan->d.lrelu && !yact)) return fail_arg("conv_tc_stage: bad arguments");
  return conv_tc_stage(plan, mode, src, yact, batch, t_in, stage_ws, (cudaStream_t)stream);
}

extern "C" int hmvae_conv_tc_run(const hmvae_conv_plan* plan, int mode, const float* wp, int batch, int t_in,
                                 const void* stage_ws, void* dump_ws, void* stream) {
  int rc = check_shape(plan, batch, t_in, "conv_tc_run");
  if (rc) return rc;
  if (!wp || (mode != 0 && mode != 1)) return fail_arg("conv_tc_run: bad arguments");
  return conv_tc_run(plan, mode, wp, batch, t_in, stage_ws, dump_ws, (cudaStream_t)stream);
}

extern "C" int hmvae_conv_tc_finish(const hmvae_conv_plan* plan, int mode, const void* dump_ws, const float* bias, float* dst,
                                    int batch, int t_in, void* stream) {
  int rc = check_shape(plan, batch, t_in, "conv_tc_finish");
  if (rc) return rc;
  if (!dump_ws || !dst || (mode != 0 && mode != 1)) return fail_arg("conv_tc_finish: bad arguments");
  return conv_tc_finish(plan, mode, dump_ws, bias, dst, batch, t_in, (cudaStream_t)stream);
}

extern "C" int hmvae_conv_fprop_tc(const hmvae_conv_plan* plan, const float* x, const float* wp_fprop, const float* bias,
                                   float* y, int batch, int t_in, void* workspace, long workspace_bytes, void* stream) {
  int rc = check_shape(plan, batch, t_in, "conv_fprop_tc");
  if (rc) return rc;
  if (!x || !wp_fprop || !y) return fail_arg("conv_fprop_tc: null pointer");
  if (batch == 0) return 0;
  return conv_tc_launch(plan, 0, x, nullptr, wp_fprop, bias, y, batch, t_in, workspace, workspace_bytes, (cudaStream_t)stream);
}

extern "C" int hmvae_conv_dgrad_tc(const hmvae_conv_plan* plan, const float* dy, const float* y, const float* wp_dgrad,
                                   float* dxin, int batch, int t_in, void* workspace, long workspace_bytes, void* stream) {
  int rc = check_shape(plan, batch, t_in, "conv_dgrad_tc");
  if (rc) return rc;
  if (!dy || !wp_dgrad || !dxin || (plan->d.lrelu && !y)) return fail_arg("conv_dgrad_tc: null pointer");
  if (batch == 0) return 0;
  return conv_tc_launch(plan, 1, dy, y, wp_dgrad, nullptr, dxin, batch, t_in, workspace, workspace_bytes, (cudaStream_t)stream);
}

extern "C" int hmvae_conv_dgrad(const hmvae_conv_plan* plan, const float* dy, const float* y, const float* w, float* dxin,
                                int batch, int t_in, int impl, void* stream) {
  int rc = check_shape(plan, batch, t_in, "conv_dgrad");
  if (rc) return rc;
  if (!dy || !w || !dxin || (plan->d.lrelu && !y)) return fail_arg("conv_dgrad: null pointer");
  if (batch == 0) return 0;
  (void)impl;
  return conv_dgrad_simt(plan, dy, y, w, dxin, batch, t_in, (cudaStream_t)stream);
}

extern "C" int hmvae_conv_wgrad(const hmvae_conv_plan* plan, const float* x, const float* dy, const float* y, float* dw,
                                float* dbias, int batch, int t_in, int accumulate, int impl, void* stream) {
  int rc = check_shape(plan, batch, t_in, "conv_wgrad");
  if (rc) return rc;
  if (!x || !dy || !dw || (plan->d.lrelu && !y)) return fail_arg("conv_wgrad: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  const hmvae_conv_desc& d = plan->d;
  if (!accumulate) {
    HMVAE_CUDA(cudaMemsetAsync(dw, 0, sizeof(float) * (size_t)d.joints * d.co * d.joints * d.ci * d.ksize, st));
    if (dbias) HMVAE_CUDA(cudaMemsetAsync(dbias, 0, sizeof(float) * (size_t)d.joints * d.co, st));
  }
  if (batch == 0) return 0;
  (void)impl;
  return conv_wgrad_simt(plan, x, dy, y, dw, dbias, batch, t_in, st);
}

/* mode 2 = wgrad in hmvae_conv_tc_supported / hmvae_conv_tc_workspace is handled by these dedicated entry points */
extern "C" int hmvae_conv_wgrad_tc_supported(const hmvae_conv_plan* plan, int batch, int t_in) {
  if (!plan || batch < 1 || t_in < 1) return 0;
  if (check_shape(plan, batch, t_in, "conv_wgrad_tc_supported")) return 0;
  return conv_wgrad_tc_supported(plan, batch, t_in) ? 1 : 0;
}

extern "C" long hmvae_conv_wgrad_tc_workspace(const hmvae_conv_plan* plan, int batch, int t_in) {
  if (!plan || batch < 1 || t_in < 1) return -1;
  return conv_wgrad_tc_workspace_bytes(plan, batch, t_in);
}

extern "C" int hmvae_conv_wgrad_tc(const hmvae_conv_plan* plan, const float* x, const float* dy, const float* y, float* dw,
                                   float* dbias, int batch, int t_in, int accumulate, void* workspace, long workspace_bytes,
                                   void* stream) {
  int rc = check_shape(plan, batch, t_in, "conv_wgrad_tc");
  if (rc) return rc;
  if (!dy || !dw || (plan->d.lrelu && !y)) return fail_arg("conv_wgrad_tc: null pointer");      // x == NULL: pre-staged
  if (batch == 0) return 0;
  return conv_wgrad_tc_launch(plan, x, dy, y, dw, dbias, batch, t_in, accumulate, workspace, workspace_bytes, (cudaStream_t)stream);
}

extern "C" int hmvae_conv_wgrad_tc_stage_x(const hmvae_conv_plan* plan, const float* x, int batch, int t_in, void* workspace,
                                           long workspace_bytes, void* stream) {
  int rc = check_shape(plan, batch, t_in, "conv_wgrad_tc_stage_x");
  if (rc) return rc;
  if (!x) return fail_arg("conv_wgrad_tc_stage_x: null pointer");
  if (batch == 0) return 0;
  return conv_wgrad_tc_stage_x(plan, x, batch, t_in, workspace, workspace_bytes, (cudaStream_t)stream);
}
