// Skeleton-aware conv: plan (immutable index tables) shared by the CUDA-core and the tcgen05 implementations.
#pragma once
#include <map>
#include <vector>

#include "common.cuh"

namespace hmvae {

struct ConvArgs {  // by-value kernel argument
  int J, ci, co, K, s, p, pad_mode, upsample, src_J, lrelu, ojs, oco, cl;
  int nnz;
  const int* nb_off;    // [J+1]  out joint -> in joints (skeleton.py:34-39)
  const int* nb_idx;    // [nnz]
  const int* nbT_off;   // [J+1]  in joint -> out joints (transpose adjacency, used by dgrad)
  const int* nbT_idx;   // [nnz]
  const int* blk_j;     // [nnz]  unmasked (j_out, j_in) weight blocks (used by wgrad)
  const int* blk_n;     // [nnz]
  const int* src;       // [J]    unpool source joint of conv-input joint n (identity when no unpool)
};

}  // namespace hmvae

struct hmvae_conv_plan {
  hmvae_conv_desc d;
  hmvae::ConvArgs a;
  std::vector<int> nb_off, nb_idx, src;
  int* dev_tables;
  int max_nb;   // largest neighbour-list length
  unsigned long long uid;   // unique per created plan (keys host-side caches; never reused, unlike the address)
  mutable std::map<int, void*> tc_tables;   // per (mode, joints-per-CTA) work tables of the tcgen05 kernels (device memory)
};

namespace hmvae {

// virtual conv input = unpool(upsample2(src tensor)); u in [0, T)
__device__ __forceinline__ float load_virtual(const float* __restrict__ x, const ConvArgs& a, long b, int n, int c, int u,
                                              int T) {
  const int ch = a.src[n] * a.ci + c;
  const long row = b * (long)(a.src_J * a.ci) + ch;
  if (a.upsample) {
    const int Ts = T >> 1, i = u >> 1;
    const int nbr = (u & 1) ? (i + 1 < Ts ? i + 1 : Ts - 1) : (i > 0 ? i - 1 : 0);
    const float* r = x + row * Ts;
    return 0.75f * r[i] + 0.25f * r[nbr];
  }
  return x[row * T + u];
}

// padded coordinate q in [0, T+2p) -> value (reflect / zeros), skeleton.py:18-19,100
__device__ __forceinline__ float load_padded(const float* __restrict__ x, const ConvArgs& a, long b, int n, int c, int q,
                                             int T) {
  int u = q - a.p;
  if (a.pad_mode == 1) {
    u = u < 0 ? -u : u;
    u = u >= T ? 2 * (T - 1) - u : u;
  } else if (u < 0 || u >= T) {
    return 0.f;
  }
  return load_virtual(x, a, b, n, c, u, T);
}

int conv_fprop_simt(const hmvae_conv_plan* plan, const float* x, const float* w, const float* bias, float* y, int B, int T,
                    cudaStream_t st);
int conv_dgrad_simt(const hmvae_conv_plan* plan, const float* dy, const float* y, const float* w, float* dxin, int B, int T,
                    cudaStream_t st);
int conv_wgrad_simt(const hmvae_conv_plan* plan, const float* x, const float* dy, const float* y, float* dw, float* dbias,
                    int B, int T, cudaStream_t st);

inline int conv_t_out(const hmvae_conv_desc& d, int T) { return (T + 2 * d.pad - d.ksize) / d.stride + 1; }

}  // namespace hmvae
