// Skeleton-aware masked conv1d, CUDA-core fp32 implementation (fprop / dgrad / wgrad / prologue adjoint).
//
// Restates skeleton.py:95-105 without ever forming W*mask: the neighbour lists drive the loops, so masked
// (j_out, j_in) blocks are never read (fprop, dgrad) or written (wgrad).  Reflect/zero padding, the decoder's
// x2 linear upsample and the unpool gather are index arithmetic inside the loaders (conv_common.cuh).
// This is the always-available exact-fp32 path and the on-device comparator for the tcgen05 kernels.
#include "conv_common.cuh"

namespace hmvae {

constexpr int CV_TPB = 256;
constexpr int CV_CC = 8;      // reduction-channel chunk staged in smem
constexpr int CV_PPT = 4;     // max positions per thread (dgrad)

__device__ __forceinline__ long out_index(const ConvArgs& a, long b, int j, int o, int t, int T_out) {
  const int ch = j * a.ojs + a.oco + o;
  const long ctot = (long)a.J * a.ojs;
  return a.cl ? (b * T_out + t) * ctot + ch : (b * ctot + ch) * T_out + t;
}

// ------------------------------------------------------------------------------------------------ fprop
template <int RO>
__global__ void __launch_bounds__(CV_TPB) conv_fprop_kernel(ConvArgs a, const float* __restrict__ x,
                                                            const float* __restrict__ w, const float* __restrict__ bias,
                                                            float* __restrict__ y, int B, int T, int T_out, int nt, int nb,
                                                            int npos, int og_cnt, int co_pad, int tq, int tq_pad) {
  extern __shared__ __align__(16) float sm[];
  float* xs = sm;                               // [CC][nb][tq_pad]
  float* ws = sm + CV_CC * nb * tq_pad;         // [CC][K][co_pad]
  const int j = blockIdx.y;
  const int ntt = (T_out + nt - 1) / nt;
  const int b0 = (blockIdx.x / ntt) * nb, t0 = (blockIdx.x % ntt) * nt;
  const int tid = threadIdx.x, pos = tid % npos, og = tid / npos;
  const int bl = pos / nt, tl = pos % nt;
  const bool act = og < og_cnt && bl < nb && (b0 + bl) < B && (t0 + tl) < T_out;
  const int q0 = t0 * a.s, Tq = T + 2 * a.p, Cin = a.J * a.ci;
  float acc[RO];
#pragma unroll
  for (int r = 0; r < RO; ++r) acc[r] = 0.f;

  for (int m = a.nb_off[j]; m < a.nb_off[j + 1]; ++m) {
    const int n = a.nb_idx[m];
    for (int c0 = 0; c0 < a.ci; c0 += CV_CC) {
      const int ccn = a.ci - c0 < CV_CC ? a.ci - c0 : CV_CC;
      for (int e = tid; e < ccn * nb * tq; e += CV_TPB) {
        const int ql = e % tq, r = e / tq, bb = r % nb, cc = r / nb;
        float v = 0.f;
        if (b0 + bb < B && q0 + ql < Tq) v = load_padded(x, a, b0 + bb, n, c0 + cc, q0 + ql, T);
        xs[(cc * nb + bb) * tq_pad + ql] = v;
      }
      for (int e = tid; e < a.co * ccn * a.K; e += CV_TPB) {
        const int k = e % a.K, r = e / a.K, cc = r % ccn, o = r / ccn;
        ws[(cc * a.K + k) * co_pad + o] = w[((long)(j * a.co + o) * Cin + n * a.ci + c0 + cc) * a.K + k];
      }
      __syncthreads();
      if (act) {
        const float* xr = xs + bl * tq_pad + tl * a.s;
        for (int cc = 0; cc < ccn; ++cc) {
          const float* xc = xr + cc * nb * tq_pad;
          const float* wc = ws + (cc * a.K) * co_pad + og * RO;
          for (int k = 0; k < a.K; ++k) {
            const float xv = xc[k];
            const float4* wv = reinterpret_cast<const float4*>(wc + k * co_pad);
#pragma unroll
            for (int r4 = 0; r4 < RO / 4; ++r4) {
              const float4 t = wv[r4];
              acc[r4 * 4 + 0] += xv * t.x; acc[r4 * 4 + 1] += xv * t.y;
              acc[r4 * 4 + 2] += xv * t.z; acc[r4 * 4 + 3] += xv * t.w;
            }
          }
        }
      }
      __syncthreads();
    }
  }
  if (act) {
#pragma unroll
    for (int r = 0; r < RO; ++r) {
      const int o = og * RO + r;
      if (o < a.co) {
        float v = acc[r] + (bias ? bias[j * a.co + o] : 0.f);
        if (a.lrelu) v = lrelu_f(v, 0.2f);
        y[out_index(a, b0 + bl, j, o, t0 + tl, T_out)] = v;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------ dgrad
// dxin[b, n*ci + c, u] for the virtual (pre-padding) conv input.  The CTA owns whole padded sequences so that the
// padding adjoint (fold of the reflected positions) stays on chip.
template <int RC>
__global__ void __launch_bounds__(CV_TPB) conv_dgrad_kernel(ConvArgs a, const float* __restrict__ dy,
                                                            const float* __restrict__ yact, const float* __restrict__ w,
                                                            float* __restrict__ dxin, int B, int T, int T_out, int nb,
                                                            int npos, int cg_cnt, int ci_pad, int zlen_pad, int ppt) {
  extern __shared__ __align__(16) float sm[];
  const int Tq = T + 2 * a.p, Cin = a.J * a.ci;
  const int zlen = Tq + a.K - 1;
  float* dyz = sm;                               // [OC][nb][zlen_pad]   zero-inserted / shifted dy
  float* ws = sm + CV_CC * nb * zlen_pad;        // [OC][K][ci_pad]
  float* dxp = ws + CV_CC * a.K * ci_pad;        // [ci][nb][Tq]
  const int n = blockIdx.y, b0 = blockIdx.x * nb;
  const int tid = threadIdx.x, p0 = tid % npos, cg = tid / npos;
  const int total_pos = nb * Tq;
  float acc[CV_PPT][RC];
#pragma unroll
  for (int i = 0; i < CV_PPT; ++i)
#pragma unroll
    for (int r = 0; r < RC; ++r) acc[i][r] = 0.f;

  for (int m = a.nbT_off[n]; m < a.nbT_off[n + 1]; ++m) {
    const int j = a.nbT_idx[m];
    for (int o0 = 0; o0 < a.co; o0 += CV_CC) {
      const int ocn = a.co - o0 < CV_CC ? a.co - o0 : CV_CC;
      for (int e = tid; e < ocn * nb * zlen; e += CV_TPB) {
        const int z = e % zlen, r = e / zlen, bb = r % nb, oc = r / nb;
        const int zz = z - (a.K - 1);
        float v = 0.f;
        if (zz >= 0 && (zz % a.s) == 0 && zz / a.s < T_out && b0 + bb < B) {
          const long oi = out_index(a, b0 + bb, j, o0 + oc, zz / a.s, T_out);
          v = dy[oi];
          if (a.lrelu && !(yact[oi] > 0.f)) v *= 0.2f;
        }
        dyz[(oc * nb + bb) * zlen_pad + z] = v;
      }
      for (int e = tid; e < ocn * a.ci * a.K; e += CV_TPB) {
        const int k = e % a.K, r = e / a.K, c = r % a.ci, oc = r / a.ci;
        ws[(oc * a.K + k) * ci_pad + c] = w[((long)(j * a.co + o0 + oc) * Cin + n * a.ci + c) * a.K + k];
      }
      __syncthreads();
      if (cg < cg_cnt) {
#pragma unroll
        for (int i = 0; i < CV_PPT; ++i) {
          const int pos = p0 + i * npos;
          if (i < ppt && pos < total_pos) {
            const int bl = pos / Tq, q = pos % Tq;
            for (int oc = 0; oc < ocn; ++oc) {
              const float* dr = dyz + (oc * nb + bl) * zlen_pad + q + (a.K - 1);
              const float* wc = ws + (oc * a.K) * ci_pad + cg * RC;
              for (int k = 0; k < a.K; ++k) {
                const float dv = dr[-k];
                const float4* wv = reinterpret_cast<const float4*>(wc + k * ci_pad);
#pragma unroll
                for (int r4 = 0; r4 < RC / 4; ++r4) {
                  const float4 t = wv[r4];
                  acc[i][r4 * 4 + 0] += dv * t.x; acc[i][r4 * 4 + 1] += dv * t.y;
                  acc[i][r4 * 4 + 2] += dv * t.z; acc[i][r4 * 4 + 3] += dv * t.w;
                }
              }
            }
          }
        }
      }
      __syncthreads();
    }
  }
  if (cg < cg_cnt) {
#pragma unroll
    for (int i = 0; i < CV_PPT; ++i) {
      const int pos = p0 + i * npos;
      if (i < ppt && pos < total_pos) {
        const int bl = pos / Tq, q = pos % Tq;
#pragma unroll
        for (int r = 0; r < RC; ++r) {
          const int c = cg * RC + r;
          if (c < a.ci) dxp[(c * nb + bl) * Tq + q] = acc[i][r];
        }
      }
    }
  }
  __syncthreads();
  for (int e = tid; e < a.ci * nb * T; e += CV_TPB) {
    const int u = e % T, r = e / T, bb = r % nb, c = r / nb;
    if (b0 + bb < B) {
      const float* row = dxp + (c * nb + bb) * Tq;
      float v = row[u + a.p];
      if (a.pad_mode == 1) {
        if (u >= 1 && u <= a.p) v += row[a.p - u];
        if (u <= T - 2 && u >= T - 1 - a.p) v += row[a.p + 2 * (T - 1) - u];
      }
      dxin[((long)(b0 + bb) * Cin + n * a.ci + c) * T + u] = v;
    }
  }
}

// ------------------------------------------------------------------------------------------------ wgrad
// One CTA per unmasked (j_out, j_in) block (x item chunk x batch slice).  A thread owns ROW output channels of one
// input channel and all K taps; when a block has fewer than 256 such items the CTA's thread groups split the batch
// and are reduced through shared memory.  Results are accumulated into dw with atomics (batch slices across CTAs).
template <int KMAX, int ROW>
__global__ void __launch_bounds__(CV_TPB) conv_wgrad_kernel(ConvArgs a, const float* __restrict__ x,
                                                            const float* __restrict__ dy, const float* __restrict__ yact,
                                                            float* __restrict__ dw, float* __restrict__ dbias, int B, int T,
                                                            int T_out, int items, int ipb, int groups, int bslice,
                                                            int co_pad, int tq_pad) {
  extern __shared__ __align__(16) float sm[];
  const int j = a.blk_j[blockIdx.x], n = a.blk_n[blockIdx.x];
  const int Tq = T + 2 * a.p, Cin = a.J * a.ci;
  const int tid = threadIdx.x, li = tid % ipb, grp = tid / ipb;
  const int item = blockIdx.y * ipb + li;
  const bool act = grp < groups && item < items;
  const int og = item / a.ci, c = item % a.ci;
  const int gstride = T_out * co_pad + a.ci * tq_pad;
  float* dys = sm + (grp < groups ? grp : 0) * gstride;   // [T_out][co_pad]
  float* xs = dys + T_out * co_pad;                        // [ci][tq_pad]
  const bool do_bias = dbias != nullptr && n == a.nb_idx[a.nb_off[j]] && c == 0;

  float acc[ROW][KMAX], bacc[ROW];
#pragma unroll
  for (int r = 0; r < ROW; ++r) {
    bacc[r] = 0.f;
#pragma unroll
    for (int k = 0; k < KMAX; ++k) acc[r][k] = 0.f;
  }
  const int bbeg = blockIdx.z * bslice, bend = (bbeg + bslice < B) ? bbeg + bslice : B;
  const int iters = (bslice + groups - 1) / groups;
  for (int it = 0; it < iters; ++it) {
    const int b = bbeg + it * groups + grp;
    const bool bval = grp < groups && b < bend;
    if (bval) {
      for (int e = li; e < a.co * T_out; e += ipb) {
        const int t = e % T_out, o = e / T_out;
        const long oi = out_index(a, b, j, o, t, T_out);
        float v = dy[oi];
        if (a.lrelu && !(yact[oi] > 0.f)) v *= 0.2f;
        dys[t * co_pad + o] = v;
      }
      for (int e = li; e < (co_pad - a.co) * T_out; e += ipb) dys[(e % T_out) * co_pad + a.co + e / T_out] = 0.f;
      for (int e = li; e < a.ci * Tq; e += ipb) {
        const int q = e % Tq, cc = e / Tq;
        xs[cc * tq_pad + q] = load_padded(x, a, b, n, cc, q, T);
      }
    }
    __syncthreads();
    if (bval && act) {
      const float* xr = xs + c * tq_pad;
      for (int t = 0; t < T_out; ++t) {
        float dv[ROW];
#pragma unroll
        for (int r = 0; r < ROW; ++r) dv[r] = dys[t * co_pad + og * ROW + r];
        if (do_bias) {
#pragma unroll
          for (int r = 0; r < ROW; ++r) bacc[r] += dv[r];
        }
        const float* xt = xr + t * a.s;
#pragma unroll
        for (int k = 0; k < KMAX; ++k) {
          if (k < a.K) {
            const float xv = xt[k];
#pragma unroll
            for (int r = 0; r < ROW; ++r) acc[r][k] += dv[r] * xv;
          }
        }
      }
    }
    __syncthreads();
  }
  // cross-group reduction through smem (reuses the staging area), then one atomic per output element
  float* red = sm;   // [groups][ipb][ROW*KMAX + ROW]
  constexpr int PER = ROW * KMAX + ROW;
  if (grp < groups) {
    float* mine = red + ((long)grp * ipb + li) * PER;
#pragma unroll
    for (int r = 0; r < ROW; ++r) {
#pragma unroll
      for (int k = 0; k < KMAX; ++k) mine[r * KMAX + k] = act ? acc[r][k] : 0.f;
      mine[ROW * KMAX + r] = (act && do_bias) ? bacc[r] : 0.f;
    }
  }
  __syncthreads();
  for (int e = tid; e < ipb * PER; e += CV_TPB) {
    const int f = e % PER, l2 = e / PER;
    const int it2 = blockIdx.y * ipb + l2;
    if (it2 >= items) continue;
    float s = 0.f;
    for (int g = 0; g < groups; ++g) s += red[((long)g * ipb + l2) * PER + f];
    const int og2 = it2 / a.ci, c2 = it2 % a.ci;
    if (f < ROW * KMAX) {
      const int r = f / KMAX, k = f % KMAX, o = og2 * ROW + r;
      if (k < a.K && o < a.co) atomicAdd(dw + ((long)(j * a.co + o) * Cin + n * a.ci + c2) * a.K + k, s);
    } else {
      const int o = og2 * ROW + (f - ROW * KMAX);
      if (dbias != nullptr && n == a.nb_idx[a.nb_off[j]] && c2 == 0 && o < a.co) atomicAdd(dbias + j * a.co + o, s);
    }
  }
}

// ------------------------------------------------------------------------------------------------ prologue adjoint
// dsrc[b, sj*ci + c, v] = sum_{n: src[n] = sj} upsample2^T(dxin[b, n*ci + c, :])[v]  (* lrelu'(src_act))
__global__ void conv_prologue_bwd_kernel(ConvArgs a, const float* __restrict__ dxin, const float* __restrict__ src_act,
                                         float* __restrict__ dsrc, int B, int T, long total) {
  pdl_trigger();
  pdl_wait();
  const int Ts = a.upsample ? T / 2 : T;
  const int Cin = a.J * a.ci;
  for (long o = (long)blockIdx.x * blockDim.x + threadIdx.x; o < total; o += (long)gridDim.x * blockDim.x) {
    const int v = (int)(o % Ts);
    long r = o / Ts;
    const int c = (int)(r % a.ci); r /= a.ci;
    const int sj = (int)(r % a.src_J);
    const long b = r / a.src_J;
    float acc = 0.f;
    for (int n = 0; n < a.J; ++n) {
      if (a.src[n] != sj) continue;
      const float* g = dxin + ((b * Cin) + n * a.ci + c) * (long)T;
      if (a.upsample) {
        float t = 0.75f * (g[2 * v] + g[2 * v + 1]);
        t += 0.25f * (v + 1 < Ts ? g[2 * v + 2] : g[2 * v + 1]);
        t += 0.25f * (v > 0 ? g[2 * v - 1] : g[0]);
        acc += t;
      } else {
        acc += g[v];
      }
    }
    if (src_act && !(src_act[o] > 0.f)) acc *= 0.2f;
    dsrc[o] = acc;
  }
}

// ------------------------------------------------------------------------------------------------ launchers
static inline int round4(int v) { return (v + 3) & ~3; }

int conv_fprop_simt(const hmvae_conv_plan* plan, const float* x, const float* w, const float* bias, float* y, int B, int T,
                    cudaStream_t st) {
  const ConvArgs& a = plan->a;
  const int T_out = conv_t_out(plan->d, T);
  const int co4 = round4(a.co);
  const int RO = co4 >= 48 ? 12 : (co4 >= 16 ? 8 : 4);
  const int og_cnt = (co4 + RO - 1) / RO, co_pad = og_cnt * RO;
  int npos = CV_TPB / og_cnt;
  if (npos < 1) return fail_arg("conv_fprop: too many output channels per joint (> 3072)");
  const int nt = T_out < npos ? T_out : npos;
  const int nb = npos / nt;
  const int tq = (nt - 1) * a.s + a.K, tq_pad = round4(tq) + 1;
  const size_t smem = ((size_t)CV_CC * nb * tq_pad + 4 + (size_t)CV_CC * a.K * co_pad) * 4;
  const int ntt = (T_out + nt - 1) / nt;
  dim3 grid(((B + nb - 1) / nb) * ntt, a.J);
  // CV_CC == 8 keeps the ws region 16-byte aligned for any nb * tq_pad
#define LAUNCH_F(R)                                                                                                   \
  {                                                                                                                   \
    HMVAE_CUDA(cudaFuncSetAttribute(conv_fprop_kernel<R>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));  \
    conv_fprop_kernel<R><<<grid, CV_TPB, smem, st>>>(a, x, w, bias, y, B, T, T_out, nt, nb, npos, og_cnt, co_pad, tq, \
                                                     tq_pad);                                                        \
  }
  if (smem > 220 * 1024) return fail_arg("conv_fprop: tile does not fit shared memory");
  if (RO == 12) LAUNCH_F(12) else if (RO == 8) LAUNCH_F(8) else LAUNCH_F(4)
#undef LAUNCH_F
  return check_launch("conv_fprop_simt");
}

int conv_dgrad_simt(const hmvae_conv_plan* plan, const float* dy, const float* y, const float* w, float* dxin, int B, int T,
                    cudaStream_t st) {
  const ConvArgs& a = plan->a;
  const int T_out = conv_t_out(plan->d, T);
  const int ci4 = round4(a.ci);
  const int RC = ci4 >= 48 ? 12 : (ci4 >= 16 ? 8 : 4);
  const int cg_cnt = (ci4 + RC - 1) / RC, ci_pad = cg_cnt * RC;
  const int npos = CV_TPB / cg_cnt;
  if (npos < 1) return fail_arg("conv_dgrad: too many input channels per joint");
  const int Tq = T + 2 * a.p;
  if (Tq > npos * CV_PPT) return fail_arg("conv_dgrad: padded sequence longer than the CTA tile");
  int nb = (npos * 2) / Tq;     // aim at <= 2 positions per thread
  if (nb < 1) nb = 1;
  const int ppt = (nb * Tq + npos - 1) / npos;
  const int zlen_pad = round4(Tq + a.K - 1) + 1;
  const size_t fl = (size_t)CV_CC * nb * zlen_pad + 4 + (size_t)CV_CC * a.K * ci_pad + (size_t)a.ci * nb * Tq;
  const size_t smem = fl * 4;
  if (smem > 220 * 1024) return fail_arg("conv_dgrad: tile does not fit shared memory");
  dim3 grid((B + nb - 1) / nb, a.J);
#define LAUNCH_D(R)                                                                                                   \
  {                                                                                                                   \
    HMVAE_CUDA(cudaFuncSetAttribute(conv_dgrad_kernel<R>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));  \
    conv_dgrad_kernel<R><<<grid, CV_TPB, smem, st>>>(a, dy, y, w, dxin, B, T, T_out, nb, npos, cg_cnt, ci_pad,        \
                                                     zlen_pad, ppt);                                                 \
  }
  if (RC == 12) LAUNCH_D(12) else if (RC == 8) LAUNCH_D(8) else LAUNCH_D(4)
#undef LAUNCH_D
  return check_launch("conv_dgrad_simt");
}

int conv_wgrad_simt(const hmvae_conv_plan* plan, const float* x, const float* dy, const float* y, float* dw, float* dbias,
                    int B, int T, cudaStream_t st) {
  const ConvArgs& a = plan->a;
  const int T_out = conv_t_out(plan->d, T);
  const int KMAX = a.K <= 4 ? 4 : (a.K <= 16 ? 16 : 32);
  if (a.K > 32) return fail_arg("conv_wgrad: kernel_size > 32 is not supported");
  const int ROW = KMAX == 32 ? 2 : 4;
  const int co_pad = ((a.co + ROW - 1) / ROW) * ROW;
  const int items = (co_pad / ROW) * a.ci;
  const int ipb = items < CV_TPB ? items : CV_TPB;
  int groups = CV_TPB / ipb;
  if (groups > B) groups = B;
  const int chunks = (items + ipb - 1) / ipb;
  const int Tq = T + 2 * a.p, tq_pad = Tq | 1;
  const int PER = ROW * KMAX + ROW;
  // batch slices: enough CTAs for ~2 waves, each slice at least `groups` sequences
  long ctas = (long)a.nnz * chunks;
  int slices = (int)((2L * num_sms() + ctas - 1) / ctas);
  int max_slices = (B + groups - 1) / groups;
  if (slices > max_slices) slices = max_slices;
  if (slices < 1) slices = 1;
  const int bslice = (B + slices - 1) / slices;
  slices = (B + bslice - 1) / bslice;
  const size_t stage = (size_t)groups * (T_out * co_pad + a.ci * tq_pad);
  const size_t red = (size_t)groups * ipb * PER;
  const size_t smem = (stage > red ? stage : red) * 4 + 16;
  if (smem > 220 * 1024) return fail_arg("conv_wgrad: tile does not fit shared memory");
  dim3 grid(a.nnz, chunks, slices);
#define LAUNCH_W(KM, RW)                                                                                                  \
  {                                                                                                                       \
    HMVAE_CUDA(cudaFuncSetAttribute(conv_wgrad_kernel<KM, RW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    conv_wgrad_kernel<KM, RW><<<grid, CV_TPB, smem, st>>>(a, x, dy, y, dw, dbias, B, T, T_out, items, ipb, groups, bslice, \
                                                          co_pad, tq_pad);                                               \
  }
  if (KMAX == 4) LAUNCH_W(4, 4) else if (KMAX == 16) LAUNCH_W(16, 4) else LAUNCH_W(32, 2)
#undef LAUNCH_W
  return check_launch("conv_wgrad_simt");
}

}  // namespace hmvae

using namespace hmvae;

extern "C" int hmvae_conv_prologue_bwd(const hmvae_conv_plan* plan, const float* dxin, const float* src_act, float* dsrc,
                                       int batch, int t_in, void* stream) {
  if (!plan || !dxin || !dsrc) return fail_arg("conv_prologue_bwd: null pointer");
  const ConvArgs& a = plan->a;
  if (a.upsample && (t_in & 1)) return fail_arg("conv_prologue_bwd: upsampled length must be even");
  const int Ts = a.upsample ? t_in / 2 : t_in;
  const long total = (long)batch * a.src_J * a.ci * Ts;
  if (total <= 0) return 0;
  long blocks = (total + 255) / 256, cap = (long)num_sms() * 16;
  launch_pdl(conv_prologue_bwd_kernel, dim3((int)(blocks < cap ? blocks : cap)), dim3(256), 0, (cudaStream_t)stream, a, dxin, src_act, dsrc, batch, t_in, total);
  return check_launch("conv_prologue_bwd");
}
