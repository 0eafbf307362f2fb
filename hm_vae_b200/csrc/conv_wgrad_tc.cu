// Skeleton-aware conv, weight gradient on the tensor cores (tcgen05 kind::tf32, accumulators in TMEM).   sm_100a only.
//
//     dW[(j,o), (n,c), k] = sum_{b,t}  dy[b, (j,o), t] * xpad[b, (n,c), t*s + k]          for n in nb(j) only
//
// GEMM view per tap k:  D_k[M = 128 dy channels, N = channels of input joint n] += A^T B, reduction over positions.
// Both operands are K-major tiles (tools/umma_probe: MN-major tf32 operands in the no-swizzle layout return zeros on sm_100a):
// rows are channels, the reduction index is the position p = t*Bt + b (time-major, sequence-minor, Bt % 4 == 0) and one
// 16-byte element holds 4 consecutive sequences of the same time step, so
//   * the 128 M-rows are 128 consecutive dy channels (several consecutive output joints),
//   * the N columns are the channels of a WINDOW of consecutive input joints (N = wj * n_pad <= 128): a kind::tf32 M=128 K=8
//     MMA from shared memory costs ~55-63 cycles for every N <= 128 (tools/umma_probe/rate.cu), so wide N is free,
//   * tap k is a start-address shift of the x tile by k*Bt/4 K-chunks (stride-2 layers keep two input phases),
//   * one tcgen05.mma consumes 8 positions; the accumulators of a whole tap group (L taps x N columns <= 512) stay in TMEM
//     while the CTA streams over (batch group, time window) stages.
// One CTA owns (M-slab, input-joint window, tap group) => every dW element is written by exactly one thread (deterministic,
// no atomics, masked blocks are never touched).  (M-slab, window) pairs without any neighbour relation are not launched.
//
//   conv_wgrad_prep_kernel : stages dy (with LeakyReLU') and the padded/upsampled/unpooled x as tf32 tiles; x is staged once per
//                            (stage, window, phase) for ALL taps, a tap group bulk-copies the contiguous chunk range it needs.
//   conv_wgrad_tc_kernel   : warp 0 TMA producer, warp 1 MMA issuer, warps 2-5 epilogue (TMEM -> global dW).
#include <string.h>

#include "conv_common.cuh"

namespace hmvae {

constexpr int WG_THREADS = 192;
constexpr int WG_MAX_STAGES = 4;
constexpr int WG_MAX_WJ = 8;       // input joints per N window

struct WgItem {      // one CTA
  int slab, win, k0, L;
  unsigned long long nbmask[WG_MAX_WJ];    // [joint of the window]: bit j set <=> (j, n) is an unmasked block
};

struct WgArgs {
  ConvArgs a;
  int B, T, T_out;
  int ckd;          // dy channels per joint padded to 8   (M side)
  int n_pad;        // x channels per joint padded to 16   (N side)
  int wj, Nw, nwinN;  // input joints per N window, columns per window (= wj * n_pad <= 128), number of windows
  int dy_chunks;    // J * ckd / 4
  int slabs;        // ceil(dy_chunks / 32)
  int Bt, mtiles;   // sequences per stage (multiple of 4), number of batch groups
  int Tw, nwin;     // output time steps per stage, windows per sequence
  int Rd;           // dy positions per stage = Tw * Bt (multiple of 8)
  int RxAll;        // staged x positions per phase and stage (all taps) = (Tw + span - 1) * Bt
  int Rx;           // x positions per phase held in shared memory by one CTA (its tap group) = (Tw + win - 1) * Bt
  int nphase;       // 1 (stride 1) or 2
  int TG, Lmax;     // tap groups, taps per group
  int a_bytes, b_bytes, stage_bytes, stages;
  int tmem_cols;
  int nitems;
  int psplits, plen;     // position split: gridDim.y CTAs per item, each handles plen consecutive stages
  const WgItem* items;
};

__device__ __forceinline__ uint32_t wg_tf32(float x) {
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
  return u;
}
__device__ __forceinline__ long wg_out_index(const ConvArgs& a, long b, int j, int o, int t, int T_out) {
  const int ch = j * a.ojs + a.oco + o;
  const long ctot = (long)a.J * a.ojs;
  return a.cl ? (b * T_out + t) * ctot + ch : (b * ctot + ch) * T_out + t;
}
__device__ __forceinline__ uint64_t wg_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46);
}

// tap group geometry: first padded input row (per phase) and number of rows of the window
__device__ __host__ __forceinline__ void wg_window(int s, int k0, int L, int phase, int* lo, int* cnt) {
  if (s == 1) { *lo = k0; *cnt = L; return; }
  // taps k in [k0, k0+L) with k % 2 == phase  ->  shifts k/2
  int first = k0 + (((k0 & 1) != phase) ? 1 : 0);
  if (first >= k0 + L) { *lo = 0; *cnt = 0; return; }
  int last = k0 + L - 1;
  if ((last & 1) != phase) --last;
  *lo = first >> 1;
  *cnt = (last >> 1) - (first >> 1) + 1;
}

// ---------------------------------------------------------------------------------------------- staging
// stage st = mt * nwin + w  (batch group mt, time window w).  Position inside a stage: p = tl*Bt + b.
// dyw[st][slab][Rd/4][128 rows][4]           rows = dy channels of the slab (padded numbering j*ckd + o)
// xw [st][win][phase][RxAll/4][Nw rows][4]   rows = (joint of the window, channel); positions count from the stage's first
//                                            input index (stride 1: padded coordinate; stride 2: index inside the phase)
__global__ void __launch_bounds__(256) conv_wgrad_prep_kernel(WgArgs p, const float* __restrict__ x,
                                                              const float* __restrict__ dy, const float* __restrict__ yact,
                                                              float4* __restrict__ dyw, float4* __restrict__ xw, int part) {
  pdl_trigger();
  pdl_wait();
  const ConvArgs& a = p.a;
  const int nst = p.mtiles * p.nwin;
  const long n_dy = (long)nst * p.slabs * (p.Rd / 4) * 128;
  const long n_x = (long)nst * p.nwinN * p.nphase * (p.RxAll / 4) * p.Nw;
  const int Tq = p.T + 2 * a.p;
  const int bq = p.Bt / 4;
  // part 0: both operands; 1: dy only (x was staged ahead of time, during the forward pass); 2: x only
  const long it_lo = (part == 2) ? n_dy : 0, it_hi = (part == 1) ? n_dy : n_dy + n_x;
  for (long it = it_lo + (long)blockIdx.x * blockDim.x + threadIdx.x; it < it_hi; it += (long)gridDim.x * blockDim.x) {
    float v[4] = {0.f, 0.f, 0.f, 0.f};
    if (it < n_dy) {
      // thread order: time fastest inside (row, b-quad) so that the 4 scalar loads of neighbouring threads coalesce along t
      long r = it;
      const int tl = (int)(r % p.Tw); r /= p.Tw;
      const int b4 = (int)(r % bq); r /= bq;
      const int row = (int)(r % 128); r /= 128;
      const int slab = (int)(r % p.slabs);
      const int st = (int)(r / p.slabs);
      const int mt = st / p.nwin, w = st % p.nwin;
      const int t = w * p.Tw + tl;
      const int gch = slab * 128 + row;
      const int j = gch / p.ckd, o = gch % p.ckd;
      if (j < a.J && o < a.co && t < p.T_out) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const long bb = (long)mt * p.Bt + b4 * 4 + i;
          if (bb < p.B) {
            const long oi = wg_out_index(a, bb, j, o, t, p.T_out);
            float g = dy[oi];
            if (a.lrelu && !(yact[oi] > 0.f)) g *= 0.2f;
            v[i] = g;
          }
        }
      }
      const long dst = (((long)st * p.slabs + slab) * (p.Rd / 4) + (tl * bq + b4)) * 128 + row;
      dyw[dst] = make_float4(__uint_as_float(wg_tf32(v[0])), __uint_as_float(wg_tf32(v[1])), __uint_as_float(wg_tf32(v[2])),
                             __uint_as_float(wg_tf32(v[3])));
    } else {
      long r = it - n_dy;
      const int ulmax = p.RxAll / p.Bt;
      const int ul = (int)(r % ulmax); r /= ulmax;
      const int b4 = (int)(r % bq); r /= bq;
      const int row = (int)(r % p.Nw); r /= p.Nw;
      const int phase = (int)(r % p.nphase); r /= p.nphase;
      const int win = (int)(r % p.nwinN);
      const int st = (int)(r / p.nwinN);
      const int mt = st / p.nwin, w = st % p.nwin;
      const int n = win * p.wj + row / p.n_pad, c = row % p.n_pad;
      const int u = w * p.Tw + ul;            // stride 2: output step t reads index t + (k >> 1) of phase (k & 1)
      const int tp = (a.s == 1) ? u : 2 * u + phase;
      if (n < a.J && c < a.ci && tp < Tq) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const long bb = (long)mt * p.Bt + b4 * 4 + i;
          if (bb < p.B) v[i] = load_padded(x, a, bb, n, c, tp, p.T);
        }
      }
      const long dst = ((((long)st * p.nwinN + win) * p.nphase + phase) * (p.RxAll / 4) + (ul * bq + b4)) * p.Nw + row;
      xw[dst] = make_float4(__uint_as_float(wg_tf32(v[0])), __uint_as_float(wg_tf32(v[1])), __uint_as_float(wg_tf32(v[2])),
                            __uint_as_float(wg_tf32(v[3])));
    }
  }
}

// db[j*co + o] = sum_{b,t} dy * lrelu'(y)          (one 128-thread CTA per channel, coalesced along t, fixed-order reduction)
__global__ void __launch_bounds__(128) conv_bias_grad_kernel(ConvArgs a, const float* __restrict__ dy,
                                                             const float* __restrict__ yact, float* __restrict__ db, int B,
                                                             int T_out, int accumulate) {
  __shared__ float red[4];
  pdl_trigger();
  pdl_wait();
  const int ch = blockIdx.x;
  const int j = ch / a.co, o = ch % a.co;
  float acc = 0.f;
  for (int e = threadIdx.x; e < B * T_out; e += 128) {
    const long oi = wg_out_index(a, e / T_out, j, o, e % T_out, T_out);
    float g = dy[oi];
    if (a.lrelu && !(yact[oi] > 0.f)) g *= 0.2f;
    acc += g;
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    const float s = (red[0] + red[1]) + (red[2] + red[3]);
    db[ch] = accumulate ? db[ch] + s : s;
  }
}

// ---------------------------------------------------------------------------------------------- main kernel
__device__ __forceinline__ bool wg_elect() {      // one lane of a converged warp
  uint32_t pred;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "elect.sync _|p, 0xffffffff;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}

__global__ void __launch_bounds__(WG_THREADS, 1) conv_wgrad_tc_kernel(WgArgs p, const unsigned char* __restrict__ dyw,
                                                                      const unsigned char* __restrict__ xw,
                                                                      float* __restrict__ dw, int accumulate) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  __shared__ uint64_t full_bar[WG_MAX_STAGES], empty_bar[WG_MAX_STAGES], accum_bar;
  __shared__ uint32_t tmem_base_s;

  pdl_trigger();
  const ConvArgs& a = p.a;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const WgItem item = p.items[blockIdx.x];          // plan constant

  if (tid == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(&accum_bar, 1);
    mbar_fence_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(p.tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_s;
  pdl_wait();       // prologue done: now wait for the staging kernel
  const int nst_all = p.mtiles * p.nwin;
  const int st_beg = blockIdx.y * p.plen;
  const int st_end = (st_beg + p.plen < nst_all) ? st_beg + p.plen : nst_all;
  if (p.psplits > 1) dw += (size_t)blockIdx.y * a.J * a.co * a.J * a.ci * a.K;      // this split's partial buffer
  const uint32_t bq = (uint32_t)p.Bt / 4;
  // the tap group's window of staged input positions, per phase: first shift `lo`, `cnt` distinct shifts
  int lo[2] = {0, 0}, cnt[2] = {0, 0};
  for (int ph = 0; ph < p.nphase; ++ph) wg_window(a.s, item.k0, item.L, ph, &lo[ph], &cnt[ph]);
  const uint32_t ph_bytes_smem = (uint32_t)(p.Rx / 4) * p.Nw * 16;      // smem region of one phase

  if (warp == 0) {
    // =============================== producer ===============================
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      uint32_t xbytes[2] = {0, 0};
      for (int q = 0; q < p.nphase; ++q)
        if (cnt[q] > 0) xbytes[q] = (uint32_t)(p.Tw + cnt[q] - 1) * bq * p.Nw * 16;
      const size_t x_phase_stride = (size_t)(p.RxAll / 4) * p.Nw * 16;
      for (int mt = st_beg; mt < st_end; ++mt) {
        mbar_wait(&empty_bar[s], ph ^ 1);
        unsigned char* st = smem_raw + (size_t)s * p.stage_bytes;
        mbar_arrive_expect_tx(&full_bar[s], (uint32_t)p.a_bytes + xbytes[0] + xbytes[1]);
        bulk_g2s(st, dyw + ((size_t)mt * p.slabs + item.slab) * p.a_bytes, (uint32_t)p.a_bytes, &full_bar[s]);
        for (int q = 0; q < p.nphase; ++q)
          if (xbytes[q])
            bulk_g2s(st + p.a_bytes + (size_t)q * ph_bytes_smem,
                     xw + (((size_t)mt * p.nwinN + item.win) * p.nphase + q) * x_phase_stride + (size_t)lo[q] * bq * p.Nw * 16,
                     xbytes[q], &full_bar[s]);
        if (++s == p.stages) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    // =============================== MMA issuer ===============================
    // instruction descriptor: F32 accum, TF32 x TF32, both operands K-major, N = Nw, M = 128
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(p.Nw >> 3) << 17) | ((128u >> 4) << 24);
    const int ksteps = p.Rd / 8;
    int s = 0;
    uint32_t ph = 0;
    for (int mt = st_beg; mt < st_end; ++mt) {
      mbar_wait(&full_bar[s], ph);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t a_base = smem_u32(smem_raw + (size_t)s * p.stage_bytes);
      const uint32_t b_base = a_base + p.a_bytes;
      // K-major, no swizzle: LBO = stride between 16-byte K chunks (= rows*16), SBO = stride between 8-row groups (128 B)
      const uint64_t adesc0 = wg_desc(a_base, 128 * 16, 128);
      const uint64_t bdesc0 = wg_desc(b_base, (uint32_t)p.Nw * 16, 128);
      const uint32_t acc0 = (mt > st_beg) ? 1u : 0u;
      if (lane == 0) {
        // single-thread issue loop (a converged-warp elect.sync variant was measured slower inside the real kernels)
        for (int t = 0; t < item.L; ++t) {
          const int k = item.k0 + t;
          const int phase = (a.s == 1) ? 0 : (k & 1);
          const int shift = ((a.s == 1) ? k : (k >> 1)) - lo[phase];
          uint64_t ad = adesc0;
          uint64_t bd = bdesc0 + (uint32_t)(phase * (p.Rx / 4) + shift * (int)bq) * (uint32_t)p.Nw;
          const uint32_t d_addr = tmem_base + (uint32_t)(t * p.Nw);
          uint32_t acc = acc0;
#pragma unroll 4
          for (int ks = 0; ks < ksteps; ++ks) {
            asm volatile(
                "{\n"
                ".reg .pred p;\n"
                "setp.ne.b32 p, %4, 0;\n"
                "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
                "}\n" ::"r"(d_addr),
                "l"(ad), "l"(bd), "r"(idesc), "r"(acc)
                : "memory");
            acc = 1;
            ad += 2 * 128;                     // 8 positions = 2 K-chunks of 128 rows x 16 B (16-byte units)
            bd += 2 * (uint32_t)p.Nw;
          }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&empty_bar[s]))
                     : "memory");
      }
      __syncwarp();
      if (++s == p.stages) { s = 0; ph ^= 1; }
    }
    if (lane == 0)
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&accum_bar)) : "memory");
    __syncwarp();
  } else {
    // =============================== epilogue (warps 2..5): TMEM -> dW ===============================
    // (A shared-memory transpose that makes every warp write one dW row at a time was measured SLOWER than these direct
    //  per-thread stores -- 0.77 vs 0.55 ms per step -- the row loop serialises; kept simple.)
    mbar_wait(&accum_bar, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int lq = warp & 3;
    const int m = lq * 32 + lane;
    const int gch = item.slab * 128 + m;                 // dy channel (padded numbering)
    const int j = gch / p.ckd, o = gch % p.ckd;
    const bool rowok = j < a.J && o < a.co;
    const int Cin = a.J * a.ci;
    float* wrow0 = dw + ((long)(rowok ? j * a.co + o : 0) * Cin) * a.K + item.k0;
    for (int c16 = 0; c16 < p.Nw; c16 += 16) {
      const int wjl = c16 / p.n_pad;                     // n_pad is a multiple of 16: a 16-column group lies inside one joint
      const int n = item.win * p.wj + wjl;
      const int cbase = c16 - wjl * p.n_pad;
      if (n >= a.J || cbase >= a.ci) continue;           // warp-uniform: padding columns
      const bool valid = rowok && ((item.nbmask[wjl] >> j) & 1ull);      // warp-divergent only in the stores
      for (int t = 0; t < item.L; ++t) {
        uint32_t r[16];
        const uint32_t taddr = tmem_base + ((uint32_t)(lq * 32) << 16) + (uint32_t)(t * p.Nw + c16);
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
              "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
            : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (valid) {
          float* wrow = wrow0 + (long)n * a.ci * a.K;
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const int c = cbase + i;
            if (c < a.ci) {
              float* dst = wrow + (long)c * a.K + t;
              const float v = __uint_as_float(r[i]);
              *dst = (accumulate && p.psplits == 1) ? *dst + v : v;
            }
          }
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(p.tmem_cols) : "memory");
  }
}

// ---------------------------------------------------------------------------------------------- split reduce
// dw[e] (= or +=) sum over splits of part[s][e], for unmasked entries only, in a fixed order (deterministic)
__global__ void conv_wgrad_reduce_kernel(WgArgs p, const float* __restrict__ part, float* __restrict__ dw, int accumulate) {
  pdl_trigger();
  pdl_wait();
  const ConvArgs& a = p.a;
  const long per_row = (long)a.ci * a.K;
  const long total = (long)a.nnz * a.co * per_row;
  const long dense = (long)a.J * a.co * a.J * a.ci * a.K;
  const int Cin = a.J * a.ci;
  for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x) {
    const long x = e % per_row;
    long r = e / per_row;
    const int o = (int)(r % a.co);
    const int blk = (int)(r / a.co);
    const int j = a.blk_j[blk], n = a.blk_n[blk];
    const long idx = ((long)(j * a.co + o) * Cin + n * a.ci) * a.K + x;
    float v = 0.f;
    for (int s2 = 0; s2 < p.psplits; ++s2) v += part[s2 * dense + idx];
    dw[idx] = accumulate ? dw[idx] + v : v;
  }
}

// ---------------------------------------------------------------------------------------------- host side
static inline int rup(int v, int m) { return (v + m - 1) / m * m; }

static bool wg_geometry_build(const hmvae_conv_plan* plan, int B, int T, WgArgs* out, std::vector<WgItem>* items_out) {
  const ConvArgs& a = plan->a;
  if (a.J > 64 || a.K > 32) return false;
  WgArgs p;
  memset(&p, 0, sizeof(p));
  p.a = a;
  p.B = B;
  p.T = T;
  p.T_out = conv_t_out(plan->d, T);
  p.ckd = rup(a.co, 8);
  p.n_pad = rup(a.ci, 16);
  if (p.n_pad > 256) return false;
  // N window: as many consecutive input joints as fit 128 columns (an MMA costs the same for every N <= 128)
  p.wj = 128 / p.n_pad;
  if (p.wj < 1) p.wj = 1;
  if (p.wj > WG_MAX_WJ) p.wj = WG_MAX_WJ;
  if (p.wj > a.J) p.wj = a.J;
  p.Nw = p.wj * p.n_pad;
  p.nwinN = (a.J + p.wj - 1) / p.wj;
  p.dy_chunks = a.J * p.ckd / 4;
  p.slabs = (p.dy_chunks + 31) / 32;
  p.nphase = a.s;
  // taps per group: L * Nw accumulator columns <= 512
  // accumulator columns per CTA.  Measured: 128 columns (one tap of a 128-column window per CTA => 15x more, shorter CTAs that
  // co-reside with the dgrad kernels of the other stream) beats 256 / 384 / 512: 0.971 / 0.997 / 1.063 / 1.112 ms per step.
  // Round 2 (linked stack path, two weight-gradient streams; graph replay of the whole step on one box): 128 columns / 148 target
  // CTAs 773 us, 256 / 60: 750 us, 384 / 40: 751 us, 512 / 30: 795 us.  A CTA re-loads the dy tile and the x window of every stage
  // for each tap it owns (ncu: the kernels are bound by that L2 -> shared-memory traffic, 276 MB for the last decoder level with
  // one tap per CTA), so two taps per CTA halve it -- as long as the position split below does not grow in return.
  const int tmem_cap = env_int("HMVAE_WG_TMEM_COLS", 256);
  p.Lmax = tmem_cap / p.Nw;
  if (p.Lmax < 1) p.Lmax = 1;
  if (p.Lmax > 16) p.Lmax = 16;                     // epilogue transpose tile: 128 x (16 * L + 1) floats <= 132 KB
  if (p.Lmax > a.K) p.Lmax = a.K;
  p.TG = (a.K + p.Lmax - 1) / p.Lmax;
  p.Lmax = (a.K + p.TG - 1) / p.TG;                 // balance the groups
  int cols = p.Lmax * p.Nw, pow2 = 32;
  while (pow2 < cols) pow2 <<= 1;
  if (pow2 > 512) return false;
  p.tmem_cols = pow2;
  // a stage = Bt sequences (multiple of 4) x Tw output time steps, ~64 positions
  p.Tw = p.T_out < 16 ? p.T_out : 16;
  p.nwin = (p.T_out + p.Tw - 1) / p.Tw;
  const int span = (a.s == 1) ? a.K : ((a.K - 1) >> 1) + 1;          // distinct input shifts over all taps (per phase)
  int win = 1;                                                        // largest shift window of one (tap group, phase)
  for (int tg = 0; tg < p.TG; ++tg) {
    const int k0 = tg * p.Lmax;
    const int L = (a.K - k0 < p.Lmax) ? a.K - k0 : p.Lmax;
    for (int ph = 0; ph < p.nphase; ++ph) {
      int lo, cnt;
      wg_window(a.s, k0, L, ph, &lo, &cnt);
      if (cnt > win) win = cnt;
    }
  }
  // sequences per stage: start at ~64 positions, shrink (multiples of 4, Tw*Bt % 8 == 0) until >= 3 stages fit
  auto size_stage = [&](int bt) {
    p.Bt = bt;
    p.Rd = p.Tw * p.Bt;
    p.RxAll = (p.Tw + span - 1) * p.Bt;
    p.Rx = (p.Tw + win - 1) * p.Bt;
    p.a_bytes = (p.Rd / 4) * 128 * 16;
    p.b_bytes = p.nphase * (p.Rx / 4) * p.Nw * 16;
    p.stage_bytes = rup(p.a_bytes + p.b_bytes, 128);
    p.stages = (env_int("HMVAE_WG_SMEM_KB", 200) * 1024) / p.stage_bytes;
  };
  int bt0 = rup(64 / p.Tw > 4 ? 64 / p.Tw : 4, 4);
  if (bt0 > rup(B, 4)) bt0 = rup(B, 4);
  while ((p.Tw * bt0) % 8 != 0) bt0 += 4;
  size_stage(bt0);
  for (int bt = bt0 - 4; p.stages < 3 && bt >= 4; bt -= 4)
    if ((p.Tw * bt) % 8 == 0) size_stage(bt);
  p.mtiles = (B + p.Bt - 1) / p.Bt;
  {
    int cap = env_int("HMVAE_WG_STAGES", 2);
    if (cap > WG_MAX_STAGES) cap = WG_MAX_STAGES;
    if (cap < 2) cap = 2;
    if (p.stages > cap) p.stages = cap;
  }
  if (p.stages < 2) return false;
  if (p.Rd * 16 >= (1 << 18) || p.Rx * 16 >= (1 << 18)) return false;
  // work items
  std::vector<WgItem> items;
  for (int slab = 0; slab < p.slabs; ++slab) {
    const int ch0 = slab * 128, ch1 = ch0 + 128;
    for (int win_i = 0; win_i < p.nwinN; ++win_i) {
      WgItem it;
      memset(&it, 0, sizeof(it));
      bool any = false;
      for (int wl = 0; wl < p.wj; ++wl) {
        const int n = win_i * p.wj + wl;
        if (n >= a.J) break;
        unsigned long long mask = 0;
        for (int j = 0; j < a.J; ++j) {
          bool nb = false;
          for (int m = plan->nb_off[j]; m < plan->nb_off[j + 1]; ++m) nb |= plan->nb_idx[m] == n;
          if (!nb) continue;
          mask |= 1ull << j;
          const int c0 = j * p.ckd, c1 = c0 + a.co;    // channel range of joint j in the padded numbering
          if (c0 < ch1 && c1 > ch0) any = true;
        }
        it.nbmask[wl] = mask;
      }
      if (!any) continue;
      for (int tg = 0; tg < p.TG; ++tg) {
        it.slab = slab; it.win = win_i; it.k0 = tg * p.Lmax;
        it.L = (a.K - it.k0 < p.Lmax) ? a.K - it.k0 : p.Lmax;
        if (it.L > 0) items.push_back(it);
      }
    }
  }
  p.nitems = (int)items.size();
  {
    // split the (batch group, time window) stages over gridDim.y so that ~one wave of SMs is busy; >= 4 stages per CTA
    const int nst = p.mtiles * p.nwin;
    int splits = p.nitems > 0 ? env_int("HMVAE_WG_TARGET_CTAS", 60) / p.nitems : 1;
    if (splits > nst / 4) splits = nst / 4;
    if (splits > 8) splits = 8;
    const long dense_bytes = (long)a.J * a.co * a.J * a.ci * a.K * 4;
    if (dense_bytes * splits > (24L << 20)) splits = (int)((24L << 20) / dense_bytes);   // partial buffers + reduce cost too much
    if (splits < 1) splits = 1;
    p.plen = (nst + splits - 1) / splits;
    p.psplits = (nst + p.plen - 1) / p.plen;
  }
  *out = p;
  *items_out = items;
  return p.nitems > 0;
}

struct WgCached {
  bool ok;
  WgArgs p;
};

static bool wg_geometry(const hmvae_conv_plan* plan, int B, int T, WgArgs* out) {
  static thread_local std::map<std::pair<unsigned long long, long>, WgCached> cache;
  const auto key = std::make_pair(plan->uid, ((long)B << 20) | (long)T);
  auto it = cache.find(key);
  if (it == cache.end()) {
    WgCached c;
    std::vector<WgItem> items;
    c.ok = wg_geometry_build(plan, B, T, &c.p, &items);
    if (c.ok) {
      const int tkey = 900000 + (int)(cache.size() % 1000) * 64 + (int)(plan->tc_tables.size());
      void* dev = nullptr;
      if (cudaMalloc(&dev, items.size() * sizeof(WgItem)) != cudaSuccess ||
          cudaMemcpy(dev, items.data(), items.size() * sizeof(WgItem), cudaMemcpyHostToDevice) != cudaSuccess) {
        c.ok = false;
      } else {
        int k2 = tkey;
        while (plan->tc_tables.count(k2)) ++k2;
        plan->tc_tables[k2] = dev;                  // freed with the plan
        c.p.items = reinterpret_cast<const WgItem*>(dev);
      }
    }
    it = cache.emplace(key, c).first;
  }
  if (it->second.ok) *out = it->second.p;
  return it->second.ok;
}

static long wg_dy_bytes(const WgArgs& p) { return (long)p.mtiles * p.nwin * p.slabs * p.a_bytes; }
static long wg_x_bytes(const WgArgs& p) { return (long)p.mtiles * p.nwin * p.nwinN * p.nphase * (p.RxAll / 4) * p.Nw * 16; }
static long wg_part_bytes(const WgArgs& p) {
  return p.psplits > 1 ? (long)p.psplits * p.a.J * p.a.co * p.a.J * p.a.ci * p.a.K * 4 : 0;
}

bool conv_wgrad_tc_supported(const hmvae_conv_plan* plan, int B, int T) {
  WgArgs p;
  return wg_geometry(plan, B, T, &p);
}

long conv_wgrad_tc_workspace_bytes(const hmvae_conv_plan* plan, int B, int T) {
  WgArgs p;
  if (!wg_geometry(plan, B, T, &p)) return -1;
  return wg_dy_bytes(p) + wg_x_bytes(p) + wg_part_bytes(p);
}

// staging pass: part 0 = both operands, 1 = dy only, 2 = x only (see conv_wgrad_prep_kernel)
static int wg_stage(const WgArgs& p, const float* x, const float* dy, const float* yact, unsigned char* dyw, unsigned char* xw,
                    int part, cudaStream_t st) {
  const long items = ((part == 2 ? 0 : wg_dy_bytes(p)) + (part == 1 ? 0 : wg_x_bytes(p))) / 16;
  // resident CTAs per SM: the staging pass runs beside the data-gradient chain; a full complement of 8 x 256 threads per SM
  // leaves the chain's CTAs no thread slots until a staging CTA retires (tools/timeline.py)
  long blocks = (items + 255) / 256, cap = (long)num_sms() * env_int("HMVAE_WG_PREP_CTAS_PER_SM", 8);
  launch_pdl(conv_wgrad_prep_kernel, dim3((int)(blocks < cap ? blocks : cap)), dim3(256), 0, st, p, x, dy, yact,
             reinterpret_cast<float4*>(dyw), reinterpret_cast<float4*>(xw), part);
  return check_launch("conv_wgrad_prep");
}

// The x operand depends on forward data only: staged during the forward pass (where a training step leaves the GPU mostly idle
// beside its one chain of kernels) into the workspace that the weight-gradient call of the backward pass receives with x == NULL.
int conv_wgrad_tc_stage_x(const hmvae_conv_plan* plan, const float* x, int B, int T, void* workspace, long workspace_bytes,
                          cudaStream_t st) {
  WgArgs p;
  if (!wg_geometry(plan, B, T, &p)) return fail_arg("conv_wgrad_stage_x (tcgen05): unsupported geometry");
  if (!workspace || workspace_bytes < wg_dy_bytes(p) + wg_x_bytes(p) + wg_part_bytes(p) || !aligned16(workspace))
    return fail_arg("conv_wgrad_stage_x (tcgen05): workspace too small or misaligned");
  unsigned char* dyw = reinterpret_cast<unsigned char*>(workspace);
  return wg_stage(p, x, nullptr, nullptr, dyw, dyw + wg_dy_bytes(p), 2, st);
}

int conv_wgrad_tc_launch(const hmvae_conv_plan* plan, const float* x, const float* dy, const float* yact, float* dw,
                         float* dbias, int B, int T, int accumulate, void* workspace, long workspace_bytes,
                         cudaStream_t st) {
  WgArgs p;
  if (!wg_geometry(plan, B, T, &p)) return fail_arg("conv_wgrad (tcgen05): unsupported geometry");
  if (!workspace || workspace_bytes < wg_dy_bytes(p) + wg_x_bytes(p) + wg_part_bytes(p) || !aligned16(workspace))
    return fail_arg("conv_wgrad (tcgen05): workspace too small or misaligned");
  unsigned char* dyw = reinterpret_cast<unsigned char*>(workspace);
  unsigned char* xw = dyw + wg_dy_bytes(p);
  {
    int rc = wg_stage(p, x, dy, yact, dyw, xw, x ? 0 : 1, st);      // x == NULL: its tiles are already in the workspace
    if (rc) return rc;
  }
  if (dbias) {
    const int ch = p.a.J * p.a.co;
    launch_pdl(conv_bias_grad_kernel, dim3(ch), dim3(128), 0, st, p.a, dy, yact, dbias, B, p.T_out, accumulate);
    int rc = check_launch("conv_bias_grad");
    if (rc) return rc;
  }
  size_t smem = (size_t)p.stages * p.stage_bytes;
  const size_t epi = (size_t)128 * (16 * p.Lmax + 1) * 4;       // the epilogue's transpose tile reuses the stage buffers
  if (epi > smem) smem = epi;
  smem += 1024;
  // the limit is a per-function permission, sticky and process-wide (forward and autograd-engine threads both launch): raise it
  // once to the device maximum (opt-in limit minus the kernel's static shared memory) instead of per launch
  static std::atomic<long> smem_max{0};
  if (smem_max.load(std::memory_order_acquire) == 0) {
    int dev = 0, optin = 0;
    cudaFuncAttributes fa;
    HMVAE_CUDA(cudaGetDevice(&dev));
    HMVAE_CUDA(cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    HMVAE_CUDA(cudaFuncGetAttributes(&fa, conv_wgrad_tc_kernel));
    const long lim = (long)optin - (long)fa.sharedSizeBytes;
    HMVAE_CUDA(cudaFuncSetAttribute(conv_wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lim));
    smem_max.store(lim, std::memory_order_release);
  }
  if ((long)smem > smem_max.load(std::memory_order_acquire)) return fail_arg("conv_wgrad (tcgen05): stage ring exceeds the shared memory of an SM");
  float* part = reinterpret_cast<float*>(xw + wg_x_bytes(p));
  dim3 grid(p.nitems, p.psplits);
  launch_pdl<true>(conv_wgrad_tc_kernel, grid, dim3(WG_THREADS), smem, st, p, (const unsigned char*)dyw, (const unsigned char*)xw,
             p.psplits > 1 ? part : dw, accumulate);
  int rc = check_launch("conv_wgrad_tc");
  if (rc || p.psplits <= 1) return rc;
  const long total = (long)p.a.nnz * p.a.co * p.a.ci * p.a.K;
  long blocks = (total + 255) / 256, cap = (long)num_sms() * 8;
  launch_pdl(conv_wgrad_reduce_kernel, dim3((int)(blocks < cap ? blocks : cap)), dim3(256), 0, st, p, (const float*)part, dw, accumulate);
  return check_launch("conv_wgrad_reduce");
}

}  // namespace hmvae
