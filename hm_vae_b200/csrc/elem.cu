// Bandwidth-bound glue of the hot path: pool / unpool / upsample / LeakyReLU / transposes, the latent
// reparametrisation + KL, MSE, trajectory accumulation, and the multi-tensor Adam step.  sm_100a CUDA cores.
#include "common.cuh"

namespace hmvae {

constexpr int EW_TPB = 256;
constexpr int MAX_EDGES = 64;

static inline int ew_grid(long n, int per_thread = 1) {
  long blocks = (n + (long)EW_TPB * per_thread - 1) / ((long)EW_TPB * per_thread);
  long cap = (long)num_sms() * 16;
  if (blocks < 1) blocks = 1;
  return (int)(blocks < cap ? blocks : cap);
}

struct EdgeTab {
  int off[MAX_EDGES + 1];   // CSR over output edges
  int idx[MAX_EDGES];
  int owner[MAX_EDGES];     // pooled edge that input edge j belongs to
  float inv[MAX_EDGES];     // 1/len of that pooled edge
};

// ------------------------------------------------------------------------------------------------ pool / unpool
__global__ void pool_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, int E_in, int E_out, int c, int T,
                                long total, EdgeTab tab, int lrelu) {
  pdl_trigger();
  pdl_wait();
  for (long o = (long)blockIdx.x * blockDim.x + threadIdx.x; o < total; o += (long)gridDim.x * blockDim.x) {
    const int t = (int)(o % T);
    long r = o / T;
    const int ch = (int)(r % c); r /= c;
    const int e = (int)(r % E_out);
    const long b = r / E_out;
    const float inv = 1.f / (float)(tab.off[e + 1] - tab.off[e]);
    float acc = 0.f;
    for (int m = tab.off[e]; m < tab.off[e + 1]; ++m)
      acc += x[((b * E_in + tab.idx[m]) * c + ch) * (long)T + t] * inv;
    y[o] = lrelu ? lrelu_f(acc, 0.2f) : acc;
  }
}

__global__ void pool_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ y, float* __restrict__ dx,
                                int E_in, int E_out, int c, int T, long total, EdgeTab tab, int lrelu) {
  pdl_trigger();
  pdl_wait();
  for (long o = (long)blockIdx.x * blockDim.x + threadIdx.x; o < total; o += (long)gridDim.x * blockDim.x) {
    const int t = (int)(o % T);
    long r = o / T;
    const int ch = (int)(r % c); r /= c;
    const int j = (int)(r % E_in);
    const long b = r / E_in;
    const int e = tab.owner[j];
    float v = 0.f;
    if (e >= 0) {
      const long src = ((b * E_out + e) * c + ch) * (long)T + t;
      v = dy[src] * tab.inv[j];
      if (lrelu && !(y[src] > 0.f)) v *= 0.2f;
    }
    dx[o] = v;
  }
}

__global__ void unpool_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, int E_in, int E_out, int c, int T,
                                  long total, EdgeTab tab) {
  for (long o = (long)blockIdx.x * blockDim.x + threadIdx.x; o < total; o += (long)gridDim.x * blockDim.x) {
    const int t = (int)(o % T);
    long r = o / T;
    const int ch = (int)(r % c); r /= c;
    const int j = (int)(r % E_out);
    const long b = r / E_out;
    y[o] = x[((b * E_in + tab.owner[j]) * c + ch) * (long)T + t];
  }
}

__global__ void unpool_bwd_kernel(const float* __restrict__ dy, float* __restrict__ dx, int E_in, int E_out, int c,
                                  int T, long total, EdgeTab tab) {
  for (long o = (long)blockIdx.x * blockDim.x + threadIdx.x; o < total; o += (long)gridDim.x * blockDim.x) {
    const int t = (int)(o % T);
    long r = o / T;
    const int ch = (int)(r % c); r /= c;
    const int i = (int)(r % E_in);
    const long b = r / E_in;
    float acc = 0.f;
    for (int m = tab.off[i]; m < tab.off[i + 1]; ++m) acc += dy[((b * E_out + tab.idx[m]) * c + ch) * (long)T + t];
    dx[o] = acc;
  }
}

// ------------------------------------------------------------------------------------------------ upsample x2 (linear, align_corners=False)
__global__ void upsample2_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, long rows, int T) {
  const long total = rows * 2 * T;
  for (long o = (long)blockIdx.x * blockDim.x + threadIdx.x; o < total; o += (long)gridDim.x * blockDim.x) {
    const int u = (int)(o % (2 * T));
    const long r = o / (2 * T);
    const int i = u >> 1;
    const int nb = (u & 1) ? (i + 1 < T ? i + 1 : T - 1) : (i > 0 ? i - 1 : 0);
    y[o] = 0.75f * x[r * T + i] + 0.25f * x[r * T + nb];
  }
}

__global__ void upsample2_bwd_kernel(const float* __restrict__ dy, float* __restrict__ dx, long rows, int T) {
  const long total = rows * T;
  for (long o = (long)blockIdx.x * blockDim.x + threadIdx.x; o < total; o += (long)gridDim.x * blockDim.x) {
    const int i = (int)(o % T);
    const long r = o / T;
    const float* g = dy + r * 2 * T;
    float acc = 0.75f * (g[2 * i] + g[2 * i + 1]);
    acc += 0.25f * (i + 1 < T ? g[2 * i + 2] : g[2 * i + 1]);   // from out[2(i+1)] or the clamped right edge
    acc += 0.25f * (i > 0 ? g[2 * i - 1] : g[0]);               // from out[2(i-1)+1] or the clamped left edge
    dx[o] = acc;
  }
}

__global__ void lrelu_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, long n, float slope) {
  for (long o = (long)blockIdx.x * blockDim.x + threadIdx.x; o < n; o += (long)gridDim.x * blockDim.x)
    y[o] = lrelu_f(x[o], slope);
}
__global__ void lrelu_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ y, float* __restrict__ dx, long n,
                                 float slope) {
  for (long o = (long)blockIdx.x * blockDim.x + threadIdx.x; o < n; o += (long)gridDim.x * blockDim.x)
    dx[o] = y[o] > 0.f ? dy[o] : dy[o] * slope;
}

// [B, C, T] -> [B, T, C] through a padded 32x32 smem tile
__global__ void transpose_kernel(const float* __restrict__ x, float* __restrict__ y, int C, int T) {
  pdl_trigger();
  pdl_wait();
  __shared__ float tile[32][33];
  const long b = blockIdx.z;
  const int c0 = blockIdx.y * 32, t0 = blockIdx.x * 32;
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int c = c0 + r, t = t0 + threadIdx.x;
    if (c < C && t < T) tile[r][threadIdx.x] = x[(b * C + c) * (long)T + t];
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int t = t0 + r, c = c0 + threadIdx.x;
    if (c < C && t < T) y[(b * T + t) * (long)C + c] = tile[threadIdx.x][r];
  }
}

// ------------------------------------------------------------------------------------------------ block reduce
__device__ __forceinline__ float block_sum(float v) {
  __shared__ float red[32];
  v = warp_sum(v);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) red[w] = v;
  __syncthreads();
  const int nw = (blockDim.x + 31) >> 5;
  v = (threadIdx.x < nw) ? red[threadIdx.x] : 0.f;
  if (w == 0) v = warp_sum(v);
  return v;   // valid in thread 0
}

// ------------------------------------------------------------------------------------------------ latent: reparam + KL
// single block: the tensors are tiny ([B*E, 2d]) and a fixed reduction order keeps the KL deterministic
__global__ void __launch_bounds__(1024) latent_fwd_kernel(const float* __restrict__ dist, const float* __restrict__ eps,
                                                          float* __restrict__ z, float* __restrict__ kl_out, long rows,
                                                          int d) {
  pdl_trigger();
  pdl_wait();
  float acc = 0.f;
  const long total = rows * d;
  for (long o = threadIdx.x; o < total; o += blockDim.x) {
    const long r = o / d;
    const int k = (int)(o % d);
    const float mu = dist[r * 2 * d + k], lv = dist[r * 2 * d + d + k];
    if (z) z[o] = eps ? eps[o] * expf(0.5f * lv) + mu : mu;
    acc += -0.5f * (1.f + lv - mu * mu - expf(lv));
  }
  acc = block_sum(acc);
  if (threadIdx.x == 0 && kl_out) atomicAdd(kl_out, acc);
}

__global__ void latent_bwd_kernel(const float* __restrict__ dist, const float* __restrict__ eps,
                                  const float* __restrict__ dz, const float* __restrict__ dkl,
                                  float* __restrict__ ddist, long rows, int d, float kl_scale) {
  pdl_trigger();
  pdl_wait();
  const long total = rows * d;
  if (dkl) kl_scale *= dkl[0];
  for (long o = (long)blockIdx.x * blockDim.x + threadIdx.x; o < total; o += (long)gridDim.x * blockDim.x) {
    const long r = o / d;
    const int k = (int)(o % d);
    const float mu = dist[r * 2 * d + k], lv = dist[r * 2 * d + d + k];
    const float g = dz ? dz[o] : 0.f;
    float dmu = g + kl_scale * mu;
    float dlv = kl_scale * 0.5f * (expf(lv) - 1.f);
    if (eps) dlv += g * eps[o] * 0.5f * expf(0.5f * lv);
    ddist[r * 2 * d + k] = dmu;
    ddist[r * 2 * d + d + k] = dlv;
  }
}

// ------------------------------------------------------------------------------------------------ loss finalize
// out[i] = acc[i] * scale[i] for i < n; out[n] = sum_i w[i] * out[i]; out[n+1] = sum_i wk[i] * out[i]; then acc is zeroed
// (ready for the next step).  n <= 8.
struct LossCoef { float scale[8], w[8], wk[8]; int n; };
__global__ void loss_finalize_kernel(float* __restrict__ acc, float* __restrict__ out, LossCoef c) {
  pdl_trigger();
  pdl_wait();
  if (threadIdx.x == 0) {
    float total = 0.f, kl = 0.f;
    for (int i = 0; i < c.n; ++i) {
      const float v = acc[i] * c.scale[i];
      out[i] = v;
      total += c.w[i] * v;
      kl += c.wk[i] * v;
      acc[i] = 0.f;
    }
    out[c.n] = total;
    out[c.n + 1] = kl;
  }
}

// ------------------------------------------------------------------------------------------------ MSE
__global__ void __launch_bounds__(EW_TPB) mse_fwd_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                                         float* __restrict__ out, long n) {
  float acc = 0.f;
  for (long o = (long)blockIdx.x * blockDim.x + threadIdx.x; o < n; o += (long)gridDim.x * blockDim.x) {
    const float d = a[o] - b[o];
    acc += d * d;
  }
  acc = block_sum(acc);
  if (threadIdx.x == 0) atomicAdd(out, acc);
}
__global__ void mse_bwd_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ da, long n,
                               float scale) {
  for (long o = (long)blockIdx.x * blockDim.x + threadIdx.x; o < n; o += (long)gridDim.x * blockDim.x)
    da[o] = scale * (a[o] - b[o]);
}

// ------------------------------------------------------------------------------------------------ weight regulariser
// Decoder-weight L2 regulariser of the latent-space optimisation loops (seq_two_hier_sa_vae.py:1382-1387, 1717-1722):
//   l_reg = sum_i mean((p_i - p0_i)^2),   dl/dp_i = 2 (p_i - p0_i) / numel_i.
// One launch for all tensors (blockIdx.y = tensor); the loss sum is accumulated atomically, the gradient (times `weight`) is
// ADDED to g when `accumulate` is set for the tensor (a gradient written by the backward pass) or stored otherwise.
constexpr int REG_MAX_T = 64;
struct RegPack {
  hmvae_reg_tensor t[REG_MAX_T];
};
__global__ void __launch_bounds__(EW_TPB) l2_reg_kernel(RegPack pack, float weight, float* __restrict__ loss) {
  const hmvae_reg_tensor T = pack.t[blockIdx.y];
  const float inv_n = 1.f / (float)T.numel;
  const float gs = 2.f * weight * inv_n;
  float acc = 0.f;
  for (long o = (long)blockIdx.x * blockDim.x + threadIdx.x; o < T.numel; o += (long)gridDim.x * blockDim.x) {
    const float d = T.p[o] - T.p0[o];
    acc += d * d;
    if (T.g) T.g[o] = T.accumulate ? T.g[o] + gs * d : gs * d;
  }
  acc = block_sum(acc);
  if (threadIdx.x == 0 && loss) atomicAdd(loss, acc * inv_n);
}

// ------------------------------------------------------------------------------------------------ trajectory
// trajectory_pred_model.py:289-303 + :237-244.  One thread per (sequence, coordinate): forward prefix sum over T of the
// de-standardised velocity (t >= 1), then a reverse prefix sum for the gradient.  T <= 1024.
__global__ void traj_fwdbwd_kernel(const float* __restrict__ vp, const float* __restrict__ vg, float m0, float m1, float m2,
                                   float s0, float s1, float s2, int B, int T, int J, float sv, float st,
                                   float* __restrict__ losses, float* __restrict__ dv) {
  const int id = blockIdx.x * blockDim.x + threadIdx.x;
  float lv = 0.f, lt = 0.f;
  if (id < B * 3) {
    const int b = id / 3, k = id % 3;
    const float mean = k == 0 ? m0 : (k == 1 ? m1 : m2);
    const float sd = k == 0 ? s0 : (k == 1 ? s1 : s2);
    const float* p = vp + (long)b * T * 3 + k;
    const float* g = vg + (long)b * T * 3 + k;
    float cp = 0.f, cg = 0.f;
    for (int t = 0; t < T; ++t) {
      const float a = p[t * 3], c = g[t * 3];
      lv += (a - c) * (a - c);
      if (t >= 1) {
        cp += mean + sd * a;
        cg += mean + sd * c;
        const float d = cp - cg;
        lt += d * d;
      }
    }
    if (dv) {
      // second sweep in reverse: suffix sum of 2*(cp-cg); recompute the prefix difference incrementally
      float diff = cp - cg, suffix = 0.f;
      float* o = dv + (long)b * T * 3 + k;
      for (int t = T - 1; t >= 0; --t) {
        const float a = p[t * 3], c = g[t * 3];
        float gr = sv * (a - c);
        if (t >= 1) {
          suffix += diff;
          gr += st * sd * suffix * (float)J;
          diff -= sd * (a - c);
        }
        o[t * 3] = gr;
      }
    }
  }
  lv = block_sum(lv);
  __syncthreads();
  lt = block_sum(lt);
  if (threadIdx.x == 0) {
    atomicAdd(losses + 0, lv);
    atomicAdd(losses + 1, lt * (float)J);
  }
}

// ------------------------------------------------------------------------------------------------ Adam
constexpr int ADAM_MAX_T = 64;
struct AdamPack {
  hmvae_adam_tensor t[ADAM_MAX_T];
};

__global__ void __launch_bounds__(EW_TPB) adam_kernel(AdamPack pack, float lr_over_bc1, float inv_sqrt_bc2, float omb1,
                                                      float beta2, float omb2, float eps, float wd, float gscale,
                                                      const float* __restrict__ dyn2) {
  pdl_trigger();
  pdl_wait();
  if (dyn2) {
    lr_over_bc1 = dyn2[0];
    inv_sqrt_bc2 = dyn2[1];
  }
  const hmvae_adam_tensor T = pack.t[blockIdx.y];
  float* __restrict__ p = T.p;
  const float* __restrict__ g = T.g;
  float* __restrict__ m = T.m;
  float* __restrict__ v = T.v;
  const long n = T.numel;
  const long n4 = ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
                    reinterpret_cast<uintptr_t>(v)) & 15) == 0 ? n / 4 : 0;
  auto upd = [&](float& pp, float gg, float& mm, float& vv) {
    gg = gg * gscale + wd * pp;
    mm = mm + (gg - mm) * omb1;
    vv = vv * beta2 + omb2 * gg * gg;
    const float denom = sqrtf(vv) * inv_sqrt_bc2 + eps;
    pp = pp - lr_over_bc1 * (mm / denom);
  };
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long)gridDim.x * blockDim.x) {
    float4 P = reinterpret_cast<float4*>(p)[i], G = reinterpret_cast<const float4*>(g)[i];
    float4 M = reinterpret_cast<float4*>(m)[i], V = reinterpret_cast<float4*>(v)[i];
    upd(P.x, G.x, M.x, V.x); upd(P.y, G.y, M.y, V.y); upd(P.z, G.z, M.z, V.z); upd(P.w, G.w, M.w, V.w);
    reinterpret_cast<float4*>(p)[i] = P;
    reinterpret_cast<float4*>(m)[i] = M;
    reinterpret_cast<float4*>(v)[i] = V;
  }
  for (long i = n4 * 4 + (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    float P = p[i], M = m[i], V = v[i];
    upd(P, g[i], M, V);
    p[i] = P; m[i] = M; v[i] = V;
  }
}

static int make_tab(const int* off, const int* idx, int n_groups, int n_items, EdgeTab* t) {
  if (n_groups > MAX_EDGES || n_items > MAX_EDGES || n_groups < 1) return fail_arg("edge table: at most 64 edges");
  for (int j = 0; j < MAX_EDGES; ++j) { t->owner[j] = -1; t->inv[j] = 0.f; t->idx[j] = 0; }
  for (int e = 0; e <= n_groups; ++e) t->off[e] = off[e];
  if (off[0] != 0 || off[n_groups] > MAX_EDGES) return fail_arg("edge table: bad CSR offsets");
  for (int e = 0; e < n_groups; ++e)
    for (int m = off[e]; m < off[e + 1]; ++m) {
      const int j = idx[m];
      if (j < 0 || j >= n_items) return fail_arg("edge table: member index out of range");
      t->idx[m] = j;
      t->owner[j] = e;
      t->inv[j] = 1.f / (float)(off[e + 1] - off[e]);
    }
  return 0;
}

}  // namespace hmvae

using namespace hmvae;

extern "C" int hmvae_pool_fwd(const float* x, float* y, int batch, int in_edges, int out_edges, int c, int t,
                              const int* pool_off, const int* pool_idx, int lrelu, void* stream) {
  if (!x || !y || !pool_off || !pool_idx) return fail_arg("pool_fwd: null pointer");
  EdgeTab tab;
  int rc = make_tab(pool_off, pool_idx, out_edges, in_edges, &tab);
  if (rc) return rc;
  const long total = (long)batch * out_edges * c * t;
  if (total <= 0) return 0;
  launch_pdl(pool_fwd_kernel, dim3(ew_grid(total)), dim3(EW_TPB), 0, (cudaStream_t)stream, x, y, in_edges, out_edges, c, t, total, tab, lrelu);
  return check_launch("pool_fwd");
}

extern "C" int hmvae_pool_bwd(const float* dy, const float* y, float* dx, int batch, int in_edges, int out_edges, int c,
                              int t, const int* pool_off, const int* pool_idx, int lrelu, void* stream) {
  if (!dy || !dx || !pool_off || !pool_idx || (lrelu && !y)) return fail_arg("pool_bwd: null pointer");
  EdgeTab tab;
  int rc = make_tab(pool_off, pool_idx, out_edges, in_edges, &tab);
  if (rc) return rc;
  const long total = (long)batch * in_edges * c * t;
  if (total <= 0) return 0;
  launch_pdl(pool_bwd_kernel, dim3(ew_grid(total)), dim3(EW_TPB), 0, (cudaStream_t)stream, dy, y, dx, in_edges, out_edges, c, t, total, tab, lrelu);
  return check_launch("pool_bwd");
}

static int unpool_tab(const int* src, int in_edges, int out_edges, EdgeTab* tab) {
  if (in_edges > MAX_EDGES || out_edges > MAX_EDGES) return fail_arg("unpool: at most 64 edges");
  int cnt[MAX_EDGES] = {0};
  for (int j = 0; j < out_edges; ++j) {
    if (src[j] < 0 || src[j] >= in_edges) return fail_arg("unpool: source edge out of range");
    cnt[src[j]]++;
  }
  tab->off[0] = 0;
  for (int i = 0; i < in_edges; ++i) tab->off[i + 1] = tab->off[i] + cnt[i];
  int fill[MAX_EDGES] = {0};
  for (int j = 0; j < MAX_EDGES; ++j) { tab->owner[j] = 0; tab->inv[j] = 1.f; tab->idx[j] = 0; }
  for (int j = 0; j < out_edges; ++j) {
    tab->owner[j] = src[j];
    tab->idx[tab->off[src[j]] + fill[src[j]]++] = j;
  }
  return 0;
}

extern "C" int hmvae_unpool_fwd(const float* x, float* y, int batch, int in_edges, int out_edges, int c, int t,
                                const int* src, void* stream) {
  if (!x || !y || !src) return fail_arg("unpool_fwd: null pointer");
  EdgeTab tab;
  int rc = unpool_tab(src, in_edges, out_edges, &tab);
  if (rc) return rc;
  const long total = (long)batch * out_edges * c * t;
  if (total <= 0) return 0;
  unpool_fwd_kernel<<<ew_grid(total), EW_TPB, 0, (cudaStream_t)stream>>>(x, y, in_edges, out_edges, c, t, total, tab);
  return check_launch("unpool_fwd");
}

extern "C" int hmvae_unpool_bwd(const float* dy, float* dx, int batch, int in_edges, int out_edges, int c, int t,
                                const int* src, void* stream) {
  if (!dy || !dx || !src) return fail_arg("unpool_bwd: null pointer");
  EdgeTab tab;
  int rc = unpool_tab(src, in_edges, out_edges, &tab);
  if (rc) return rc;
  const long total = (long)batch * in_edges * c * t;
  if (total <= 0) return 0;
  unpool_bwd_kernel<<<ew_grid(total), EW_TPB, 0, (cudaStream_t)stream>>>(dy, dx, in_edges, out_edges, c, t, total, tab);
  return check_launch("unpool_bwd");
}

extern "C" int hmvae_upsample2_fwd(const float* x, float* y, long rows, int t, void* stream) {
  if (!x || !y) return fail_arg("upsample2_fwd: null pointer");
  if (rows * t <= 0) return 0;
  upsample2_fwd_kernel<<<ew_grid(rows * 2 * t), EW_TPB, 0, (cudaStream_t)stream>>>(x, y, rows, t);
  return check_launch("upsample2_fwd");
}
extern "C" int hmvae_upsample2_bwd(const float* dy, float* dx, long rows, int t, void* stream) {
  if (!dy || !dx) return fail_arg("upsample2_bwd: null pointer");
  if (rows * t <= 0) return 0;
  upsample2_bwd_kernel<<<ew_grid(rows * t), EW_TPB, 0, (cudaStream_t)stream>>>(dy, dx, rows, t);
  return check_launch("upsample2_bwd");
}
extern "C" int hmvae_lrelu_fwd(const float* x, float* y, long n, float slope, void* stream) {
  if (!x || !y) return fail_arg("lrelu_fwd: null pointer");
  if (n <= 0) return 0;
  lrelu_fwd_kernel<<<ew_grid(n), EW_TPB, 0, (cudaStream_t)stream>>>(x, y, n, slope);
  return check_launch("lrelu_fwd");
}
extern "C" int hmvae_lrelu_bwd(const float* dy, const float* y, float* dx, long n, float slope, void* stream) {
  if (!dy || !y || !dx) return fail_arg("lrelu_bwd: null pointer");
  if (n <= 0) return 0;
  lrelu_bwd_kernel<<<ew_grid(n), EW_TPB, 0, (cudaStream_t)stream>>>(dy, y, dx, n, slope);
  return check_launch("lrelu_bwd");
}
extern "C" int hmvae_transpose_ct(const float* x, float* y, int batch, int c, int t, void* stream) {
  if (!x || !y) return fail_arg("transpose_ct: null pointer");
  if (batch <= 0 || c <= 0 || t <= 0) return 0;
  if (batch > 65535) return fail_arg("transpose_ct: batch > 65535");
  dim3 grid((t + 31) / 32, (c + 31) / 32, batch), block(32, 8);
  launch_pdl(transpose_kernel, dim3(grid), dim3(block), 0, (cudaStream_t)stream, x, y, c, t);
  return check_launch("transpose_ct");
}

extern "C" int hmvae_latent_fwd(const float* dist, const float* eps, float* z, float* kl_out, long rows, int d,
                                void* stream) {
  if (!dist) return fail_arg("latent_fwd: null pointer");
  if (rows * d <= 0) return 0;
  launch_pdl(latent_fwd_kernel, dim3(1), dim3(1024), 0, (cudaStream_t)stream, dist, eps, z, kl_out, rows, d);
  return check_launch("latent_fwd");
}
extern "C" int hmvae_latent_bwd(const float* dist, const float* eps, const float* dz, const float* dkl, float* ddist,
                                long rows, int d, float kl_scale, void* stream) {
  if (!dist || !ddist) return fail_arg("latent_bwd: null pointer");
  if (rows * d <= 0) return 0;
  launch_pdl(latent_bwd_kernel, dim3(ew_grid(rows * d)), dim3(EW_TPB), 0, (cudaStream_t)stream, dist, eps, dz, dkl, ddist, rows, d, kl_scale);
  return check_launch("latent_bwd");
}
extern "C" int hmvae_loss_finalize(float* acc, float* out, const float* scale, const float* w, const float* wk, int n,
                                   void* stream) {
  if (!acc || !out || !scale || !w || !wk || n < 1 || n > 8) return fail_arg("loss_finalize: bad arguments");
  LossCoef c;
  for (int i = 0; i < 8; ++i) { c.scale[i] = i < n ? scale[i] : 0.f; c.w[i] = i < n ? w[i] : 0.f; c.wk[i] = i < n ? wk[i] : 0.f; }
  c.n = n;
  launch_pdl(loss_finalize_kernel, dim3(1), dim3(32), 0, (cudaStream_t)stream, acc, out, c);
  return check_launch("loss_finalize");
}

extern "C" int hmvae_mse_fwd(const float* a, const float* b, float* out, long n, void* stream) {
  if (!a || !b || !out) return fail_arg("mse_fwd: null pointer");
  if (n <= 0) return 0;
  mse_fwd_kernel<<<ew_grid(n, 4), EW_TPB, 0, (cudaStream_t)stream>>>(a, b, out, n);
  return check_launch("mse_fwd");
}
extern "C" int hmvae_mse_bwd(const float* a, const float* b, float* da, long n, float scale, void* stream) {
  if (!a || !b || !da) return fail_arg("mse_bwd: null pointer");
  if (n <= 0) return 0;
  mse_bwd_kernel<<<ew_grid(n), EW_TPB, 0, (cudaStream_t)stream>>>(a, b, da, n, scale);
  return check_launch("mse_bwd");
}
extern "C" int hmvae_traj_fwdbwd(const float* root_v_pred, const float* root_v_gt, const float* mean3, const float* std3,
                                 int batch, int t, int joints, float sv, float st, float* losses, float* d_root_v,
                                 void* stream) {
  if (!root_v_pred || !root_v_gt || !mean3 || !std3 || !losses) return fail_arg("traj_fwdbwd: null pointer");
  if (batch <= 0 || t <= 0) return 0;
  const int n = batch * 3;
  traj_fwdbwd_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(root_v_pred, root_v_gt, mean3[0], mean3[1], mean3[2],
                                                                       std3[0], std3[1], std3[2], batch, t, joints, sv, st,
                                                                       losses, d_root_v);
  return check_launch("traj_fwdbwd");
}

extern "C" int hmvae_l2_reg_fwdbwd(const hmvae_reg_tensor* tensors, int n_tensors, float weight, float* loss, void* stream) {
  if (n_tensors > 0 && !tensors) return fail_arg("l2_reg_fwdbwd: null pointer");
  for (int base = 0; base < n_tensors; base += REG_MAX_T) {
    RegPack pack;
    const int cnt = n_tensors - base < REG_MAX_T ? n_tensors - base : REG_MAX_T;
    long maxn = 0;
    for (int i = 0; i < cnt; ++i) {
      pack.t[i] = tensors[base + i];
      if (!pack.t[i].p || !pack.t[i].p0 || pack.t[i].numel <= 0) return fail_arg("l2_reg_fwdbwd: null tensor pointer / empty tensor");
      if (pack.t[i].numel > maxn) maxn = pack.t[i].numel;
    }
    long bx = (maxn + 4 * EW_TPB - 1) / (4 * EW_TPB), cap = (long)num_sms() * 2;
    if (bx < 1) bx = 1;
    if (bx > cap) bx = cap;
    l2_reg_kernel<<<dim3((unsigned)bx, (unsigned)cnt), EW_TPB, 0, (cudaStream_t)stream>>>(pack, weight, loss);
    int rc = check_launch("l2_reg_fwdbwd");
    if (rc) return rc;
  }
  return 0;
}

static int adam_launch(const hmvae_adam_tensor* tensors, int n_tensors, float lr_over_bc1, float inv_sqrt_bc2,
                       const float* dyn2, double beta1, double beta2, float eps, float weight_decay, float grad_scale,
                       void* stream) {
  for (int base = 0; base < n_tensors; base += ADAM_MAX_T) {
    AdamPack pack;
    const int cnt = n_tensors - base < ADAM_MAX_T ? n_tensors - base : ADAM_MAX_T;
    long maxn = 0;
    for (int i = 0; i < cnt; ++i) {
      pack.t[i] = tensors[base + i];
      if (!pack.t[i].p || !pack.t[i].g || !pack.t[i].m || !pack.t[i].v) return fail_arg("adam_step: null tensor pointer");
      if (pack.t[i].numel > maxn) maxn = pack.t[i].numel;
    }
    if (maxn == 0) continue;
    long bx = (maxn / 4 + EW_TPB - 1) / EW_TPB;
    long cap = (long)num_sms() * 4;
    if (bx < 1) bx = 1;
    if (bx > cap) bx = cap;
    dim3 grid((unsigned)bx, (unsigned)cnt);
    launch_pdl(adam_kernel, dim3(grid), dim3(EW_TPB), 0, (cudaStream_t)stream, pack, lr_over_bc1, inv_sqrt_bc2, (float)(1.0 - beta1), (float)beta2, (float)(1.0 - beta2), eps, weight_decay, grad_scale, dyn2);
    int rc = check_launch("adam_step");
    if (rc) return rc;
  }
  return 0;
}

// Optimiser clock on the DEVICE.  clock[0] = completed Adam steps t, clock[1] = scheduler iterations (StepLR position).  One
// thread advances both and writes {lr / (1 - b1^t), 1 / sqrt(1 - b2^t)} for the step that starts now, in double precision like
// torch.optim.Adam does on the host.  The step's CUDA graph contains this node instead of a pinned-host -> device copy, so a host
// that queues many replays ahead can no longer overwrite the scalars of a step that has not run yet.
__global__ void opt_clock_tick_kernel(unsigned int* __restrict__ clock, float base_lr, float gamma, int step_size, double beta1,
                                      double beta2, float* __restrict__ dyn2) {
  pdl_trigger();
  pdl_wait();
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const unsigned int t = clock[0] + 1u, it = clock[1];
  double lr = (double)base_lr;
  if (step_size > 0) lr *= pow((double)gamma, (double)(it / (unsigned int)step_size));
  dyn2[0] = (float)(lr / (1.0 - pow(beta1, (double)t)));
  dyn2[1] = (float)(1.0 / sqrt(1.0 - pow(beta2, (double)t)));
  clock[0] = t;
  clock[1] = it + 1u;
}

extern "C" int hmvae_opt_clock_tick(unsigned int* clock, float base_lr, float gamma, int step_size, double beta1, double beta2,
                                    float* dyn2, void* stream) {
  if (!clock || !dyn2) return fail_arg("opt_clock_tick: null pointer");
  launch_pdl(opt_clock_tick_kernel, dim3(1), dim3(32), 0, (cudaStream_t)stream, clock, base_lr, gamma, step_size, beta1, beta2, dyn2);
  return check_launch("opt_clock_tick");
}

extern "C" int hmvae_adam_step(const hmvae_adam_tensor* tensors, int n_tensors, float lr, double beta1, double beta2,
                               float eps, float weight_decay, int step, float grad_scale, void* stream) {
  if (!tensors || n_tensors < 0 || step < 1) return fail_arg("adam_step: bad arguments");
  const double bc1 = 1.0 - pow(beta1, (double)step);
  const double bc2 = 1.0 - pow(beta2, (double)step);
  return adam_launch(tensors, n_tensors, (float)((double)lr / bc1), (float)(1.0 / sqrt(bc2)), nullptr, beta1, beta2, eps,
                     weight_decay, grad_scale, stream);
}

extern "C" int hmvae_adam_step_dyn(const hmvae_adam_tensor* tensors, int n_tensors, const float* dyn2, double beta1,
                                   double beta2, float eps, float weight_decay, float grad_scale, void* stream) {
  if (!tensors || n_tensors < 0 || !dyn2) return fail_arg("adam_step_dyn: bad arguments");
  return adam_launch(tensors, n_tensors, 0.f, 0.f, dyn2, beta1, beta2, eps, weight_decay, grad_scale, stream);
}
