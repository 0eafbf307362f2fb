// Inter-layer link kernel: everything that sits BETWEEN two tcgen05 convs of a stack, in one launch.
//
// A conv on the tensor cores is  stage (gather + tf32 rounding into UMMA tiles) -> conv_tc_kernel (accumulator dump) -> finish
// (split-K sum, bias, activation, layout).  Between two consecutive convs the reference additionally runs SkeletonPool + LeakyReLU
// (encoder, seq_two_hier_sa_vae.py:120-130), nn.Upsample + SkeletonUnpool (+ the per-edge concat of the last decoder level,
// :233-258, :278-288), or their adjoints in the backward pass.  That used to be 3-4 dependent launches of ~5-8 us each per
// boundary (finish, pool / prologue_bwd, the next stage pass): more than half of the conv chain at B=32
// (tools/tc_phases.py: the MMA kernels themselves are 9-13 us).  Here ONE kernel reads the producer's accumulator dump once,
// forms the boundary tensor S, writes it (it is needed again: activation mask, weight-gradient operand) and scatters its
// tf32-rounded values straight into the consumer's staged tiles.
//
//   kind 0  forward        S = act( mean_{m in pool(e)} (sum_z dump_P[.., m, ..] + bias_P[m]) )  (+ per-edge concat of `aux`)
//                          consumer = fprop of the next conv: reflect/zero padding, x2 linear upsample and unpool fan-out are
//                          applied while scattering
//   kind 1  backward, dec  S = upsample2^T unpool^T ( G ),   G = sum_z dump_P rows (+ reflect-padding fold) = d/d(conv input)
//                          consumer = dgrad of the previous conv: LeakyReLU' (mask = its activated output) + zero insertion
//   kind 2  backward, enc  S[j] = pool^T( lrelu'(S_fwd) * (G + add) )  = gradient of the previous conv's (pre-pool) output
//                          consumer = dgrad of that conv
// One thread owns 4 consecutive channels of one joint of one sequence and LK_CH consecutive time steps of S: every dump element
// is read exactly once per S element that needs it, S is written once, every staged element is written exactly once.  The
// consumer's staging buffer is persistent and zero-initialised by the caller: padding rows / channels and zero-inserted
// positions are never written and stay 0.
#include <string.h>

#include "conv_tc.cuh"

namespace hmvae {

constexpr int LK_CH = 2;          // S time steps per item
constexpr int LK_MAXJ = 64;

struct LinkArgs {
  TcArgs P, C;                    // producer geometry (dump layout) / consumer geometry (staging layout); has_c == 0: no consumer
  int kind, has_c;
  int zg;                         // lanes per item: the split-K partials of an item are summed by zg lanes (power of two <= 8)
  int ES, cs, cp, TS;             // S: joints, channels per joint, channels per joint that come from the dump, time steps
  int act;                        // kind 0: LeakyReLU on S; kind 2: multiply by LeakyReLU'(sact)
  int EP;                         // kind 2: joints of `add` / `sact` (= producer input joints)
  const float4* dump;
  const float* bias;              // kind 0
  const float* aux;               // kind 0: [B, ES*(cs-cp), TS] second source of the per-edge concat
  const float* add;               // kind 2: [B, EP*cs, TS] extra gradient of the pooled activation (latent head)
  const float* sact;              // kind 2: [B, EP*cs, TS] pooled, activated forward tensor
  const float* yact_c;            // kinds 1/2: activated output of the consumer conv (layout of S), when the consumer fuses LeakyReLU
  float* S;
  float4* astage;
  unsigned char m_off[LK_MAXJ + 1], m_idx[LK_MAXJ];    // S joint -> producer-side joints (pool members / unpool fan-in / owner)
  float m_scale[LK_MAXJ];                              // per S joint: 1/|pool| (kinds 0, 2), 1 (kind 1)
  unsigned char f_off[LK_MAXJ + 1], f_idx[LK_MAXJ];    // S joint -> consumer K-side joints to scatter into
};

__device__ __forceinline__ float4 f4_add(float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
__device__ __forceinline__ float4 f4_scale(float4 a, float s) { return make_float4(a.x * s, a.y * s, a.z * s, a.w * s); }
__device__ __forceinline__ float4 f4_axpy(float s, float4 a, float4 b) {       // s*a + b
  return make_float4(fmaf(s, a.x, b.x), fmaf(s, a.y, b.y), fmaf(s, a.z, b.z), fmaf(s, a.w, b.w));
}

__device__ __forceinline__ float4 lk_tf32(float4 v) {
  return make_float4(__uint_as_float(to_tf32(v.x)), __uint_as_float(to_tf32(v.y)), __uint_as_float(to_tf32(v.z)),
                     __uint_as_float(to_tf32(v.w)));
}

// staged-tile address of (consumer tile mt, K-side joint n, channel group q, row)
__device__ __forceinline__ size_t lk_stage_index(const TcArgs& C, int mt, int n, int q, int row) {
  const int qpb = C.KC >> 2, nq = C.ck_pad >> 2;
  const int cb = q / qpb, h = q - cb * qpb;
  return ((((size_t)mt * C.a.J + n) * (nq / qpb) + cb) * qpb + h) * C.rows_alloc + row;
}

// dump rows one item accumulates side by side: forward CH + 2 (one-step halo for the consumer's upsample), decoder backward
// 2*CH + 2 (upsample adjoint), encoder backward CH.  The kernel is instantiated per kind so that each variant only carries its own
// accumulators (one generic kernel needed 154 registers: 12 warps per SM, and at B=512 ran at 1.3 TB/s).
template <int KIND> struct LkRows { static constexpr int N = KIND == 0 ? LK_CH + 2 : (KIND == 1 ? 2 * LK_CH + 2 : LK_CH); };

// acc[k] += sum over THIS LANE'S splits z = zl, zl + ZG, ... of dump[z][mt][g][row[k]][col4], for the k with row[k] >= 0.
// The LK_MAXK loads of one split are independent and issued back to back, and the splits of one item are spread over ZG lanes
// (an item at B=32 has up to 21 split-K partials x 2 pool members x 3 reflect folds: walked by one thread that is a ~10 us
// dependent-latency chain; measured 20-30 us per link before this).
template <int NK>
__device__ __forceinline__ void lk_accum(const TcArgs& P, const float4* __restrict__ dump, int mt, int g, int col4, int zl, int ZG,
                                         const int (&row)[NK], float4 (&acc)[NK]) {
  const int dc4 = P.dcols >> 2;
  const size_t zstride = (size_t)P.mtiles * P.groups * 128 * dc4;
  const float4* d = dump + ((size_t)mt * P.groups + g) * 128 * dc4 + col4;
#pragma unroll 2
  for (int z = zl; z < P.splits; z += ZG) {
    float4 t[NK];
#pragma unroll
    for (int k = 0; k < NK; ++k)
      t[k] = row[k] >= 0 ? d[(size_t)z * zstride + (size_t)row[k] * dc4] : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int k = 0; k < NK; ++k) acc[k] = f4_add(acc[k], t[k]);
  }
}

// gradient w.r.t. the producer conv's (virtual) input, joint n, channel group q, steps u0 .. u0 + nk - 1 (reflect-padding adjoint
// folded in): G[k] += (this lane's share of the splits)
template <int NK>
__device__ __forceinline__ void lk_dgrad_rows(const TcArgs& P, const float4* __restrict__ dump, int b, int n, int q, int u0, int nk,
                                              int zl, int ZG, float4 (&G)[NK]) {
  const ConvArgs& a = P.a;
  const int mt = b / P.Bt, bl = b - mt * P.Bt;
  const int g = n / P.GJ, col4 = ((n % P.GJ) * P.n_pad >> 2) + q;
  int row[NK];
#pragma unroll
  for (int k = 0; k < NK; ++k) {
    const int u = u0 + k;
    row[k] = (k < nk && u >= 0 && u < P.T) ? (u + a.p) * P.Bt + bl : -1;
  }
  lk_accum(P, dump, mt, g, col4, zl, ZG, row, G);
  if (a.pad_mode == 1) {
    bool any = false;
#pragma unroll
    for (int k = 0; k < NK; ++k) {
      const int u = u0 + k;
      row[k] = (k < nk && u >= 1 && u <= a.p && u < P.T) ? (a.p - u) * P.Bt + bl : -1;
      any |= row[k] >= 0;
    }
    if (any) lk_accum(P, dump, mt, g, col4, zl, ZG, row, G);
    any = false;
#pragma unroll
    for (int k = 0; k < NK; ++k) {
      const int u = u0 + k;
      row[k] = (k < nk && u >= 0 && u <= P.T - 2 && u >= P.T - 1 - a.p) ? (a.p + 2 * (P.T - 1) - u) * P.Bt + bl : -1;
      any |= row[k] >= 0;
    }
    if (any) lk_accum(P, dump, mt, g, col4, zl, ZG, row, G);
  }
}

// butterfly sum over the ZG lanes of an item (fixed order: deterministic); every lane ends up with the total
template <int NK>
__device__ __forceinline__ void lk_reduce(float4 (&acc)[NK], int ZG, unsigned mask) {
  for (int o = ZG >> 1; o > 0; o >>= 1) {
#pragma unroll
    for (int k = 0; k < NK; ++k) {
      acc[k].x += __shfl_xor_sync(mask, acc[k].x, o);
      acc[k].y += __shfl_xor_sync(mask, acc[k].y, o);
      acc[k].z += __shfl_xor_sync(mask, acc[k].z, o);
      acc[k].w += __shfl_xor_sync(mask, acc[k].w, o);
    }
  }
}

template <int KIND>
__global__ void __launch_bounds__(128, 4) conv_link_kernel(const __grid_constant__ LinkArgs L) {
  constexpr int NK = LkRows<KIND>::N;
  pdl_trigger();
  pdl_wait();
  const TcArgs& P = L.P;
  const TcArgs& C = L.C;
  const int ZG = L.zg;                                   // lanes per item (power of two <= 8)
  const int nq = (L.cs + 3) >> 2;
  const int nch = (L.TS + LK_CH - 1) / LK_CH;
  const long total = (long)P.B * L.ES * nq * nch;
  const long gtid = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const int zl = (int)(gtid & (ZG - 1));
  const unsigned mask = (ZG == 32) ? 0xffffffffu : (((1u << ZG) - 1u) << ((threadIdx.x & 31) & ~(ZG - 1)));
  for (long it = gtid / ZG; it < total; it += ((long)gridDim.x * blockDim.x) / ZG) {
    // time chunk fastest: the lanes of a warp write consecutive steps of the same channel rows of S (coalesced NCW stores; with
    // the channel group fastest every 8-byte store was its own sector: 80 us for the 19 MB of the last decoder level at B=512)
    const int ch = (int)(it % nch);
    long r = it / nch;
    const int q = (int)(r % nq); r /= nq;
    const int e = (int)(r % L.ES);
    const int b = (int)(r / L.ES);
    const int i0 = ch * LK_CH;
    const int i1 = (i0 + LK_CH < L.TS) ? i0 + LK_CH : L.TS;
    const int c0 = q << 2;
    const bool up_f = KIND == 0 && L.has_c && C.a.upsample;      // forward consumer blends neighbouring S steps: one-step halo
    float4 sv[LK_CH + 2];                                            // sv[k] = S at step i0 - 1 + k
#pragma unroll
    for (int k = 0; k < LK_CH + 2; ++k) sv[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    const int klo = (up_f && i0 > 0) ? 0 : 1, khi = (up_f && i1 < L.TS) ? (i1 - i0 + 2) : (i1 - i0 + 1);
    const float scale = L.m_scale[e];
    float4 acc[NK];
#pragma unroll
    for (int k = 0; k < NK; ++k) acc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    if constexpr (KIND == 0) {
      if (c0 < L.cp) {
        const int mt = b / P.Bt, bl = b - mt * P.Bt;
        int row[NK];
#pragma unroll
        for (int k = 0; k < NK; ++k) row[k] = (k >= klo && k < khi) ? (i0 - 1 + k) * P.Bt + bl : -1;
        float4 bsum = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int m = L.m_off[e]; m < L.m_off[e + 1]; ++m) {
          const int j = L.m_idx[m];
          lk_accum(P, L.dump, mt, j / P.GJ, ((j % P.GJ) * P.n_pad >> 2) + q, zl, ZG, row, acc);
          if (L.bias) {
            const float* bp = L.bias + j * P.a.co + c0;
            bsum.x += bp[0];
            if (c0 + 1 < P.a.co) bsum.y += bp[1];
            if (c0 + 2 < P.a.co) bsum.z += bp[2];
            if (c0 + 3 < P.a.co) bsum.w += bp[3];
          }
        }
        lk_reduce(acc, ZG, mask);
#pragma unroll
        for (int k = 0; k < LK_CH + 2; ++k) {
          float4 v = f4_scale(f4_add(acc[k], bsum), scale);          // mean over the pool members of (conv + bias)
          if (L.act) { v.x = lrelu_f(v.x, 0.2f); v.y = lrelu_f(v.y, 0.2f); v.z = lrelu_f(v.z, 0.2f); v.w = lrelu_f(v.w, 0.2f); }
          if (c0 + 1 >= L.cp) v.y = 0.f;
          if (c0 + 2 >= L.cp) v.z = 0.f;
          if (c0 + 3 >= L.cp) v.w = 0.f;
          sv[k] = v;
        }
      } else {
        const int ca = L.cs - L.cp;
#pragma unroll
        for (int k = 0; k < LK_CH + 2; ++k) {
          if (k < klo || k >= khi) continue;
          const float* ap = L.aux + (((size_t)b * L.ES + e) * ca + (c0 - L.cp)) * L.TS + (i0 - 1 + k);
          float4 v = make_float4(ap[0], 0.f, 0.f, 0.f);
          if (c0 + 1 < L.cs) v.y = ap[(size_t)L.TS];
          if (c0 + 2 < L.cs) v.z = ap[2 * (size_t)L.TS];
          if (c0 + 3 < L.cs) v.w = ap[3 * (size_t)L.TS];
          sv[k] = v;
        }
      }
    } else if constexpr (KIND == 1) {
      const bool up = P.a.upsample != 0;
      const int u0 = up ? 2 * i0 - 1 : i0;                           // acc[k] = gradient at conv-input step u0 + k
      const int nk = up ? 2 * (i1 - i0) + 2 : (i1 - i0);
      for (int m = L.m_off[e]; m < L.m_off[e + 1]; ++m) lk_dgrad_rows(P, L.dump, b, L.m_idx[m], q, u0, nk, zl, ZG, acc);
      lk_reduce(acc, ZG, mask);
#pragma unroll
      for (int kk = 0; kk < LK_CH; ++kk) {
        const int t = i0 + kk;
        if (t >= i1) break;
        if (up) {
          // upsample2^T: .75 (g[2t] + g[2t+1]) + .25 (t+1 < Ts ? g[2t+2] : g[2t+1]) + .25 (t > 0 ? g[2t-1] : g[0])
          float4 v = f4_scale(f4_add(acc[2 * kk + 1], acc[2 * kk + 2]), 0.75f);
          v = f4_axpy(0.25f, (t + 1 < L.TS) ? acc[2 * kk + 3] : acc[2 * kk + 2], v);
          v = f4_axpy(0.25f, (t > 0) ? acc[2 * kk] : acc[2 * kk + 1], v);
          sv[kk + 1] = v;
        } else {
          sv[kk + 1] = acc[kk];
        }
      }
    } else {
      const int ep = L.m_idx[L.m_off[e]];                        // the pooled joint that S joint e was averaged into
      lk_dgrad_rows(P, L.dump, b, ep, q, i0, i1 - i0, zl, ZG, acc);
      lk_reduce(acc, ZG, mask);
#pragma unroll
      for (int kk = 0; kk < LK_CH; ++kk) {
        const int t = i0 + kk;
        if (t >= i1) break;
        float4 v = acc[kk];
        const size_t base = (((size_t)b * L.EP + ep) * L.cs + c0) * L.TS + t;
        if (L.add) {
          v.x += L.add[base];
          if (c0 + 1 < L.cs) v.y += L.add[base + L.TS];
          if (c0 + 2 < L.cs) v.z += L.add[base + 2 * (size_t)L.TS];
          if (c0 + 3 < L.cs) v.w += L.add[base + 3 * (size_t)L.TS];
        }
        if (L.act) {
          if (!(L.sact[base] > 0.f)) v.x *= 0.2f;
          if (c0 + 1 < L.cs && !(L.sact[base + L.TS] > 0.f)) v.y *= 0.2f;
          if (c0 + 2 < L.cs && !(L.sact[base + 2 * (size_t)L.TS] > 0.f)) v.z *= 0.2f;
          if (c0 + 3 < L.cs && !(L.sact[base + 3 * (size_t)L.TS] > 0.f)) v.w *= 0.2f;
        }
        sv[kk + 1] = f4_scale(v, scale);
      }
    }
    // ---- the boundary tensor itself (NCW): 4 channel rows x (i1 - i0) consecutive steps, written by lane 0 of the item
    if (zl == 0 && L.S != nullptr) {
      float* sp = L.S + (((size_t)b * L.ES + e) * L.cs + c0) * L.TS + i0;
      const int len = i1 - i0;
      const float rowv[4][LK_CH] = {{sv[1].x, sv[2].x}, {sv[1].y, sv[2].y}, {sv[1].z, sv[2].z}, {sv[1].w, sv[2].w}};
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        if (c0 + c >= L.cs) break;
        float* rp = sp + (size_t)c * L.TS;
        if (len == LK_CH && (L.TS & 1) == 0) {
          *reinterpret_cast<float2*>(rp) = make_float2(rowv[c][0], rowv[c][1]);
        } else {
          for (int k = 0; k < len; ++k) rp[k] = rowv[c][k];
        }
      }
    }
    if (!L.has_c) continue;
    // ---- scatter into the consumer's staged tiles; the (fan-out joint, step) pairs of an item are dealt round-robin to its lanes
    const int mtc = b / C.Bt, blc = b - mtc * C.Bt;
    int turn = 0;
    if constexpr (KIND == 0) {
      const ConvArgs& ca = C.a;
      const int Tc = C.T;
      auto put = [&](int n, int u, float4 val) {
        const float4 tv = lk_tf32(val);
        auto at = [&](int qp) {
          const int row = (ca.s == 1) ? qp * C.Bt + blc : ((qp & 1) * C.Tp2 + (qp >> 1)) * C.Bt + blc;
          L.astage[lk_stage_index(C, mtc, n, q, row)] = tv;
        };
        at(u + ca.p);
        if (ca.pad_mode == 1) {
          if (u >= 1 && u <= ca.p) at(ca.p - u);
          if (u <= Tc - 2 && u >= Tc - 1 - ca.p) at(ca.p + 2 * (Tc - 1) - u);
        }
      };
      for (int f = L.f_off[e]; f < L.f_off[e + 1]; ++f) {
        const int n = L.f_idx[f];
#pragma unroll
        for (int kk = 0; kk < LK_CH; ++kk) {
          const int i = i0 + kk;
          if (i >= i1) break;
          if (((turn++) & (ZG - 1)) != zl) continue;
          const float4 cur = sv[kk + 1];
          if (!ca.upsample) {
            put(n, i, cur);
          } else {
            const float4 lo = (i > 0) ? sv[kk] : cur, hi = (i + 1 < L.TS) ? sv[kk + 2] : cur;
            put(n, 2 * i, f4_axpy(0.75f, cur, f4_scale(lo, 0.25f)));
            put(n, 2 * i + 1, f4_axpy(0.75f, cur, f4_scale(hi, 0.25f)));
          }
        }
      }
    } else {
      // dgrad staging of the consumer conv: its output gradient at step t sits at zero-inserted row t*s + K-1
      const ConvArgs& ca = C.a;
      if (c0 >= ca.co) continue;
      for (int f = L.f_off[e]; f < L.f_off[e + 1]; ++f) {
        const int n = L.f_idx[f];
#pragma unroll
        for (int kk = 0; kk < LK_CH; ++kk) {
          const int i = i0 + kk;
          if (i >= i1) break;
          if (((turn++) & (ZG - 1)) != zl) continue;
          float4 val = sv[kk + 1];
          if (ca.lrelu) {
            const size_t base = (((size_t)b * L.ES + e) * L.cs + c0) * L.TS + i;
            if (!(L.yact_c[base] > 0.f)) val.x *= 0.2f;
            if (c0 + 1 < ca.co && !(L.yact_c[base + L.TS] > 0.f)) val.y *= 0.2f;
            if (c0 + 2 < ca.co && !(L.yact_c[base + 2 * (size_t)L.TS] > 0.f)) val.z *= 0.2f;
            if (c0 + 3 < ca.co && !(L.yact_c[base + 3 * (size_t)L.TS] > 0.f)) val.w *= 0.2f;
          }
          if (c0 + 1 >= ca.co) val.y = 0.f;
          if (c0 + 2 >= ca.co) val.z = 0.f;
          if (c0 + 3 >= ca.co) val.w = 0.f;
          const int row = (i * ca.s + ca.K - 1) * C.Bt + blc;
          L.astage[lk_stage_index(C, mtc, n, q, row)] = lk_tf32(val);
        }
      }
    }
  }
}

}  // namespace hmvae

using namespace hmvae;

static const char* link_build(const hmvae_conv_link_desc& d, LinkArgs* out) {
  LinkArgs& L = *out;
  memset(&L, 0, sizeof(L));
  if (!d.prod || !d.dump || (!d.s_out && !d.cons)) return "null pointer";      // s_out may be NULL (inference: staging only)
  if (d.kind < 0 || d.kind > 2 || d.batch < 1) return "bad kind / batch";
  const int pmode = d.kind == 0 ? 0 : 1;
  if (!tc_geometry(d.prod, d.batch, d.prod_t, pmode, &L.P)) return "producer geometry not supported by the tcgen05 path";
  if (L.P.ntt > 1) return "time-tiled producer (T + 2p > 128)";
  L.kind = d.kind;
  L.has_c = d.cons ? 1 : 0;
  if (d.cons) {
    if (!d.stage_ws) return "consumer without a staging buffer";
    if (!tc_geometry(d.cons, d.batch, d.cons_t, pmode, &L.C)) return "consumer geometry not supported by the tcgen05 path";
    if (L.C.ntt > 1) return "time-tiled consumer";
  }
  const ConvArgs& pa = L.P.a;
  L.dump = reinterpret_cast<const float4*>(d.dump);
  L.bias = d.bias; L.aux = d.aux; L.add = d.add; L.sact = d.sact; L.yact_c = d.yact_c;
  L.S = d.s_out;
  L.astage = reinterpret_cast<float4*>(d.stage_ws);
  L.act = d.act ? 1 : 0;
  std::vector<std::vector<int>> mem, fan;
  if (d.kind == 0) {
    // S joints: pooled edges (or the producer's own joints); channels per joint = the producer's output joint stride
    if (pa.cl) return "channels-last producer";
    L.cp = pa.co;
    L.cs = d.aux ? pa.ojs : pa.co;
    if (pa.oco != 0 || (!d.aux && pa.ojs != pa.co)) return "producer output layout";
    if (d.aux && (L.cp & 3)) return "concat boundary must be a multiple of 4 channels";
    L.TS = L.P.T_out;
    if (d.pool_off) {
      L.ES = d.pool_joints;
      if (L.ES < 1 || L.ES > LK_MAXJ) return "bad pool table";
      mem.resize(L.ES);
      for (int e = 0; e < L.ES; ++e)
        for (int m = d.pool_off[e]; m < d.pool_off[e + 1]; ++m) {
          if (d.pool_idx[m] < 0 || d.pool_idx[m] >= pa.J) return "pool member out of range";
          mem[e].push_back(d.pool_idx[m]);
        }
    } else {
      L.ES = pa.J;
      mem.resize(L.ES);
      for (int e = 0; e < L.ES; ++e) mem[e].push_back(e);
    }
    if (d.pool_off && d.aux) return "pool + concat";
    fan.resize(L.ES);
    if (d.cons) {
      const ConvArgs& ca = L.C.a;
      if (ca.src_J != L.ES || ca.ci != L.cs) return "consumer input does not match the boundary tensor";
      if (L.C.T != (ca.upsample ? 2 * L.TS : L.TS)) return "consumer length does not match the boundary tensor";
      for (int n = 0; n < ca.J; ++n) fan[d.cons->src[n]].push_back(n);
    }
    for (int e = 0; e < L.ES; ++e) L.m_scale[e] = 1.f / (float)mem[e].size();
  } else if (d.kind == 1) {
    L.ES = pa.src_J;
    L.cs = L.cp = pa.ci;
    L.TS = pa.upsample ? L.P.T / 2 : L.P.T;
    mem.resize(L.ES);
    for (int n = 0; n < pa.J; ++n) mem[d.prod->src[n]].push_back(n);
    fan.resize(L.ES);
    for (int e = 0; e < L.ES; ++e) { L.m_scale[e] = 1.f; fan[e].push_back(e); }
    if (d.cons) {
      const ConvArgs& ca = L.C.a;
      if (ca.J != L.ES || ca.ojs != L.cs || ca.oco != 0 || ca.cl || L.C.T_out != L.TS) return "consumer output does not match the boundary tensor";
      if (ca.lrelu && !d.yact_c) return "consumer fuses LeakyReLU: its activated output is required";
    }
  } else {
    // S joints = the consumer conv's output joints (pre-pool); producer input joints = pooled edges
    if (!d.pool_off) return "kind 2 needs the pooling list";
    if (pa.upsample || pa.src_J != pa.J) return "encoder-side producer with a prologue";
    L.EP = pa.J;
    if (d.pool_joints != L.EP) return "pool table does not match the producer";
    L.cs = L.cp = pa.ci;
    L.TS = L.P.T;
    int es = 0;
    for (int e = 0; e < L.EP; ++e)
      for (int m = d.pool_off[e]; m < d.pool_off[e + 1]; ++m) es = d.pool_idx[m] + 1 > es ? d.pool_idx[m] + 1 : es;
    L.ES = es;
    if (L.ES < 1 || L.ES > LK_MAXJ) return "bad pool table";
    mem.assign(L.ES, {});
    for (int e = 0; e < L.EP; ++e)
      for (int m = d.pool_off[e]; m < d.pool_off[e + 1]; ++m) {
        const int j = d.pool_idx[m];
        if (j < 0 || !mem[j].empty()) return "pool table is not a partition";
        mem[j].push_back(e);
        L.m_scale[j] = 1.f / (float)(d.pool_off[e + 1] - d.pool_off[e]);
      }
    for (int j = 0; j < L.ES; ++j)
      if (mem[j].empty()) return "pool table is not a partition";
    if (d.act && !d.sact) return "kind 2 with activation needs the forward tensor";
    fan.resize(L.ES);
    for (int e = 0; e < L.ES; ++e) fan[e].push_back(e);
    if (d.cons) {
      const ConvArgs& ca = L.C.a;
      if (ca.J != L.ES || ca.co != L.cs || ca.ojs != L.cs || ca.oco != 0 || ca.cl || L.C.T_out != L.TS || ca.lrelu)
        return "consumer output does not match the boundary tensor";
    }
  }
  if (L.ES > LK_MAXJ) return "too many joints";
  int mo = 0, fo = 0;
  for (int e = 0; e < L.ES; ++e) {
    L.m_off[e] = (unsigned char)mo;
    for (int v : mem[e]) { if (mo >= LK_MAXJ) return "member table overflow"; L.m_idx[mo++] = (unsigned char)v; }
    L.f_off[e] = (unsigned char)fo;
    for (int v : fan[e]) { if (fo >= LK_MAXJ) return "fan-out table overflow"; L.f_idx[fo++] = (unsigned char)v; }
  }
  L.m_off[L.ES] = (unsigned char)mo;
  L.f_off[L.ES] = (unsigned char)fo;
  return nullptr;
}

extern "C" int hmvae_conv_link_supported(const hmvae_conv_link_desc* desc) {
  if (!desc) return 0;
  hmvae_conv_link_desc d = *desc;
  static float dummy;
  if (!d.dump) d.dump = &dummy;
  if (!d.s_out && !d.cons) d.s_out = &dummy;
  if (d.cons && !d.stage_ws) d.stage_ws = &dummy;
  LinkArgs L;
  return link_build(d, &L) == nullptr ? 1 : 0;
}

extern "C" int hmvae_conv_link(const hmvae_conv_link_desc* desc, void* stream) {
  if (!desc) return fail_arg("conv_link: null descriptor");
  LinkArgs L;
  const char* err = link_build(*desc, &L);
  if (err) {
    snprintf(g_err, sizeof(g_err), "hmvae: conv_link: %s", err);
    return HMVAE_E_ARG;
  }
  if (!aligned16(desc->dump) || (desc->s_out && !aligned16(desc->s_out)) || (desc->cons && !aligned16(desc->stage_ws)))
    return fail_arg("conv_link: buffers must be 16-byte aligned");
  const long items = (long)L.P.B * L.ES * ((L.cs + 3) / 4) * ((L.TS + LK_CH - 1) / LK_CH);
  // lanes per item: enough to cut the split-K walk to <= 2-3 partials per lane, but only while the grid stays small
  int zg = 1;
  // (measured, graph replay of the len64 stacks at B=32: 1 lane 85.9 / 147.1 us encoder / decoder forward, up to 8 lanes
  //  94.8 / 165.1 us -- the extra waves cost more than the shorter walks save; the lanes stay available as a knob)
  const int zg_max = env_int("HMVAE_LINK_ZG", 1), per_lane = env_int("HMVAE_LINK_ZPER", 2);
  while (zg < zg_max && zg * 2 * per_lane <= 2 * L.P.splits && items * zg * 2 <= (long)num_sms() * 2048) zg *= 2;
  L.zg = zg;
  const long total = items * zg;
  long blocks = (total + 127) / 128, cap = (long)num_sms() * 16;
  if (blocks < 1) blocks = 1;
  const dim3 grid((unsigned)(blocks < cap ? blocks : cap));
  if (L.kind == 0) launch_pdl(conv_link_kernel<0>, grid, dim3(128), 0, (cudaStream_t)stream, L);
  else if (L.kind == 1) launch_pdl(conv_link_kernel<1>, grid, dim3(128), 0, (cudaStream_t)stream, L);
  else launch_pdl(conv_link_kernel<2>, grid, dim3(128), 0, (cudaStream_t)stream, L);
  return check_launch("conv_link");
}
