// The latent bottleneck of the hierarchy in one kernel per direction (seq_two_hier_sa_vae.py:159-164 encoder heads, :357-391 +
// :419-428 reparametrisation and KL, :267 / :225-229 decoder heads):
//
//     dist = x We^T + be            x: [rows, F]  (one row per (sequence, edge): the level's features, F = channels x frames)
//     z    = eps * exp(lv / 2) + mu (mu | lv) = dist,  KL row sums accumulated
//     feat = z Wd^T + bd            [rows, F] -> the decoder's hier_feats of that level
//
// At B=32 these were 4 tiny GEMM launches of ~12 us + 2 latent kernels in the middle of the forward critical path (70 us of a
// 0.88 ms step, tools/timeline.py) and the mirror image in the backward pass.  Here up to HD_MAXL levels go through ONE launch: a CTA
// owns HD_ROWS rows of one level, keeps them in shared memory, and every weight element it reads is reused for all its rows.
// Weight gradients stay on the generic linear kernels (side stream); this kernel writes gdist for them.
#include <string.h>

#include "common.cuh"

namespace hmvae {

constexpr int HD_ROWS = 4;
constexpr int HD_MAXL = 4;
constexpr int HD_THREADS = 256;
constexpr int HD_WCH = 12;           // weight elements per lane held in registers (covers 384 features per pass)
constexpr int HD_MAXD = 32;          // latent width per edge

struct HeadLevel {
  const float* x;      // [rows, F]
  const float* We;     // [2d, F]
  const float* be;     // [2d] or NULL
  const float* eps;    // [rows, d] or NULL (z = mu)
  const float* Wd;     // [F, d]
  const float* bd;     // [F] or NULL
  float* dist;         // [rows, 2d]
  float* z;            // [rows, d]
  float* feat;         // [rows, F]
  float* kl;           // device float[1], atomically accumulated
  // backward
  const float* gfeat;  // [rows, F]
  float* gdist;        // [rows, 2d]
  float* gx;           // [rows, F]
  float kl_scale;      // kl_w / rows
  long gfeat_stride;   // floats between consecutive rows of gfeat (>= F: the gradient may live inside a wider concat buffer)
  int rows, F, d;
  int cta0;            // first CTA of this level
};
struct HeadArgs {
  HeadLevel lv[HD_MAXL];
  int nlev;
};

__device__ __forceinline__ int hd_find_level(const HeadArgs& A, int cta) {
  int l = 0;
  while (l + 1 < A.nlev && cta >= A.lv[l + 1].cta0) ++l;
  return l;
}

__global__ void __launch_bounds__(HD_THREADS) heads_fwd_kernel(const __grid_constant__ HeadArgs A) {
  extern __shared__ float hs[];
  pdl_trigger();
  pdl_wait();
  const int l = hd_find_level(A, blockIdx.x);
  const HeadLevel& L = A.lv[l];
  const int F = L.F, d = L.d, d2 = 2 * L.d;
  const int r0 = (blockIdx.x - L.cta0) * HD_ROWS;
  const int nr = (L.rows - r0 < HD_ROWS) ? L.rows - r0 : HD_ROWS;
  float* xs = hs;                          // [HD_ROWS][F]
  float* ds = xs + HD_ROWS * F;            // [HD_ROWS][2d]
  float* zs = ds + HD_ROWS * d2;           // [HD_ROWS][d]
  __shared__ float kl_red[HD_THREADS / 32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int e = tid; e < HD_ROWS * F; e += HD_THREADS) {
    const int r = e / F;
    xs[e] = r < nr ? L.x[(size_t)(r0 + r) * F + (e - r * F)] : 0.f;
  }
  __syncthreads();
  // ---- dist = x We^T + be: one warp per output, lanes over the reduction, the weight element is reused for all rows
  for (int o = warp; o < d2; o += HD_THREADS / 32) {
    float acc[HD_ROWS];
#pragma unroll
    for (int r = 0; r < HD_ROWS; ++r) acc[r] = 0.f;
    const float* wrow = L.We + (size_t)o * F;
    for (int i0 = 0; i0 < F; i0 += 32 * HD_WCH) {
      // all loads of the pass are issued before the first use (a one-load-per-iteration loop is a chain of L2 latencies:
      // the first version of this kernel took 41 us)
      float w[HD_WCH];
#pragma unroll
      for (int k = 0; k < HD_WCH; ++k) {
        const int i = i0 + lane + 32 * k;
        w[k] = i < F ? wrow[i] : 0.f;
      }
#pragma unroll
      for (int k = 0; k < HD_WCH; ++k) {
        const int i = i0 + lane + 32 * k;
        if (i < F) {
#pragma unroll
          for (int r = 0; r < HD_ROWS; ++r) acc[r] = fmaf(w[k], xs[r * F + i], acc[r]);
        }
      }
    }
    const float b = L.be ? L.be[o] : 0.f;
#pragma unroll
    for (int r = 0; r < HD_ROWS; ++r) {
      const float v = warp_sum(acc[r]);
      if (lane == 0) ds[r * d2 + o] = v + b;
    }
  }
  __syncthreads();
  // ---- z = eps * exp(lv / 2) + mu, KL row sums
  float kl = 0.f;
  for (int e = tid; e < nr * d; e += HD_THREADS) {
    const int r = e / d, c = e - r * d;
    const float mu = ds[r * d2 + c], lv = ds[r * d2 + d + c];
    const size_t row = (size_t)(r0 + r);
    const float z = L.eps ? fmaf(L.eps[row * d + c], expf(0.5f * lv), mu) : mu;
    kl += -0.5f * (1.f + lv - mu * mu - expf(lv));
    zs[r * d + c] = z;
    L.z[row * d + c] = z;
  }
  for (int e = tid; e < nr * d2; e += HD_THREADS) {
    const int r = e / d2;
    L.dist[(size_t)(r0 + r) * d2 + (e - r * d2)] = ds[e];
  }
  kl = warp_sum(kl);
  if (lane == 0) kl_red[warp] = kl;
  __syncthreads();
  if (tid == 0 && L.kl) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < HD_THREADS / 32; ++w) s += kl_red[w];
    atomicAdd(L.kl, s);
  }
  // ---- feat = z Wd^T + bd: one thread per output column, the d weights of the column are reused for all rows
  for (int j = tid; j < F; j += HD_THREADS) {
    float acc[HD_ROWS];
    const float b = L.bd ? L.bd[j] : 0.f;
#pragma unroll
    for (int r = 0; r < HD_ROWS; ++r) acc[r] = b;
    const float* wrow = L.Wd + (size_t)j * d;
#pragma unroll 8
    for (int c = 0; c < d; ++c) {
      const float w = wrow[c];
#pragma unroll
      for (int r = 0; r < HD_ROWS; ++r) acc[r] = fmaf(w, zs[r * d + c], acc[r]);
    }
#pragma unroll
    for (int r = 0; r < HD_ROWS; ++r)
      if (r < nr) L.feat[(size_t)(r0 + r) * F + j] = acc[r];
  }
}

// gz = gfeat Wd;  gdist = [gz + s mu | gz eps exp(lv/2)/2 + s (exp(lv) - 1)/2];  gx = gdist We      (s = kl_w / rows)
__global__ void __launch_bounds__(HD_THREADS) heads_bwd_kernel(const __grid_constant__ HeadArgs A) {
  extern __shared__ float hs[];
  pdl_trigger();
  pdl_wait();
  const int l = hd_find_level(A, blockIdx.x);
  const HeadLevel& L = A.lv[l];
  const int F = L.F, d = L.d, d2 = 2 * L.d;
  const int r0 = (blockIdx.x - L.cta0) * HD_ROWS;
  const int nr = (L.rows - r0 < HD_ROWS) ? L.rows - r0 : HD_ROWS;
  float* gs = hs;                          // [HD_ROWS][F]
  float* gz = gs + HD_ROWS * F;            // [HD_ROWS][d]
  float* gd = gz + HD_ROWS * d;            // [HD_ROWS][2d]
  float* wds = gd + HD_ROWS * d2;          // [F][d]: the decoder head's weight, staged once per CTA (coalesced) -- read column-wise below
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  (void)lane; (void)warp;
  for (int e = tid; e < HD_ROWS * F; e += HD_THREADS) {
    const int r = e / F;
    gs[e] = r < nr ? L.gfeat[(size_t)(r0 + r) * L.gfeat_stride + (e - r * F)] : 0.f;
  }
  for (int e = tid; e < F * d; e += HD_THREADS) wds[e] = L.Wd[e];
  __syncthreads();
  // gz[r][c] = sum_j gfeat[r][j] Wd[j][c]: one thread per (row, c); threads of a warp read consecutive c (conflict-free), the
  // gfeat element is a broadcast.  (Reading Wd[j*d + c] from global memory, lanes over j, was a strided gather: 34 us per call.)
  for (int e = tid; e < HD_ROWS * d; e += HD_THREADS) {
    const int r = e / d, c = e - r * d;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    const float* g = gs + r * F;
    int jj = 0;
    for (; jj + 3 < F; jj += 4) {
      a0 = fmaf(g[jj], wds[jj * d + c], a0);
      a1 = fmaf(g[jj + 1], wds[(jj + 1) * d + c], a1);
      a2 = fmaf(g[jj + 2], wds[(jj + 2) * d + c], a2);
      a3 = fmaf(g[jj + 3], wds[(jj + 3) * d + c], a3);
    }
    for (; jj < F; ++jj) a0 = fmaf(g[jj], wds[jj * d + c], a0);
    gz[e] = (a0 + a1) + (a2 + a3);
  }
  __syncthreads();
  for (int e = tid; e < HD_ROWS * d; e += HD_THREADS) {
    const int r = e / d, c = e - r * d;
    float gmu = 0.f, glv = 0.f;
    if (r < nr) {
      const size_t row = (size_t)(r0 + r);
      const float mu = L.dist[row * d2 + c], lv = L.dist[row * d2 + d + c];
      const float g = gz[e];
      gmu = g + L.kl_scale * mu;
      glv = L.kl_scale * 0.5f * (expf(lv) - 1.f);
      if (L.eps) glv += g * L.eps[row * d + c] * 0.5f * expf(0.5f * lv);
      L.gdist[row * d2 + c] = gmu;
      L.gdist[row * d2 + d + c] = glv;
    }
    gd[r * d2 + c] = gmu;
    gd[r * d2 + d + c] = glv;
  }
  __syncthreads();
  for (int i = tid; i < F; i += HD_THREADS) {
    float acc[HD_ROWS];
#pragma unroll
    for (int r = 0; r < HD_ROWS; ++r) acc[r] = 0.f;
#pragma unroll 8
    for (int o = 0; o < d2; ++o) {
      const float w = L.We[(size_t)o * F + i];
#pragma unroll
      for (int r = 0; r < HD_ROWS; ++r) acc[r] = fmaf(w, gd[r * d2 + o], acc[r]);
    }
#pragma unroll
    for (int r = 0; r < HD_ROWS; ++r)
      if (r < nr) L.gx[(size_t)(r0 + r) * F + i] = acc[r];
  }
}

}  // namespace hmvae

using namespace hmvae;

static int heads_pack(const hmvae_head_level* levels, int n, bool bwd, HeadArgs* A, int* ctas, size_t* smem) {
  if (!levels || n < 1 || n > HD_MAXL) return fail_arg("latent_heads: 1..4 levels");
  memset(A, 0, sizeof(*A));
  A->nlev = n;
  int cta = 0;
  size_t sm = 0;
  for (int l = 0; l < n; ++l) {
    const hmvae_head_level& h = levels[l];
    HeadLevel& L = A->lv[l];
    if (h.rows < 1 || h.features < 1 || h.d < 1 || h.d > HD_MAXD) return fail_arg("latent_heads: bad rows / features / latent width");
    if (!h.dist || !h.enc_w || !h.dec_w) return fail_arg("latent_heads: null pointer");
    if (!bwd && (!h.x || !h.z || !h.feat)) return fail_arg("latent_heads_fwd: null pointer");
    if (bwd && (!h.gfeat || !h.gdist || !h.gx)) return fail_arg("latent_heads_bwd: null pointer");
    L.x = h.x; L.We = h.enc_w; L.be = h.enc_b; L.eps = h.eps; L.Wd = h.dec_w; L.bd = h.dec_b;
    L.dist = h.dist; L.z = h.z; L.feat = h.feat; L.kl = h.kl_acc;
    L.gfeat = h.gfeat; L.gdist = h.gdist; L.gx = h.gx; L.kl_scale = h.kl_scale;
    L.gfeat_stride = h.gfeat_stride > 0 ? h.gfeat_stride : h.features;
    if (L.gfeat_stride < h.features) return fail_arg("latent_heads: gfeat_stride smaller than the row");
    L.rows = h.rows; L.F = h.features; L.d = h.d;
    L.cta0 = cta;
    cta += (h.rows + HD_ROWS - 1) / HD_ROWS;
    size_t need = (size_t)HD_ROWS * (h.features + 3 * h.d) * sizeof(float);
    if (bwd) need += (size_t)h.features * h.d * sizeof(float);
    if (need > sm) sm = need;
  }
  if (sm > 200 * 1024) return fail_arg("latent_heads: feature rows too long for shared memory");
  *ctas = cta;
  *smem = sm;
  return 0;
}

extern "C" int hmvae_latent_heads_fwd(const hmvae_head_level* levels, int n_levels, void* stream) {
  HeadArgs A;
  int ctas = 0;
  size_t smem = 0;
  int rc = heads_pack(levels, n_levels, false, &A, &ctas, &smem);
  if (rc) return rc;
  if (smem > 48 * 1024) HMVAE_CUDA(cudaFuncSetAttribute(heads_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  launch_pdl(heads_fwd_kernel, dim3(ctas), dim3(HD_THREADS), smem, (cudaStream_t)stream, A);
  return check_launch("latent_heads_fwd");
}

extern "C" int hmvae_latent_heads_bwd(const hmvae_head_level* levels, int n_levels, void* stream) {
  HeadArgs A;
  int ctas = 0;
  size_t smem = 0;
  int rc = heads_pack(levels, n_levels, true, &A, &ctas, &smem);
  if (rc) return rc;
  if (smem > 48 * 1024) HMVAE_CUDA(cudaFuncSetAttribute(heads_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  launch_pdl(heads_bwd_kernel, dim3(ctas), dim3(HD_THREADS), smem, (cudaStream_t)stream, A);
  return check_launch("latent_heads_bwd");
}
