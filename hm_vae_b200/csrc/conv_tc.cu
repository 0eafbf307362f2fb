// Skeleton-aware conv on the 5th-gen tensor cores: tcgen05.mma kind::tf32, FP32 accumulators in TMEM.   sm_100a only.
//
// "Shift-GEMM" formulation (conv).  For one output joint j and one input joint n in its neighbour list
//     D_j[m, o] += sum_k  A_n[m + shift(k), c] * W_{j,n,k}[o, c]
// where the rows of A are (time, sequence) pairs laid out time-major/sequence-minor, 16 bytes (4 tf32 channels) per row
// and per K-chunk.  In the canonical no-swizzle K-major UMMA layout (8-row core matrices packed back to back,
// SBO = 128 B, LBO = chunk stride) the operand for tap k is the SAME tile with its start address advanced by shift(k)
// rows, so no im2col expansion is ever materialised: one staged activation tile feeds all K taps and every output joint
// that has n as a neighbour.  Masked (j, n) blocks are never visited.
//   * fprop : rows = (t_out, b); stride-2 layers keep even/odd input phases in two row ranges so taps stay shifts.
//   * dgrad : same kernel with the roles of the channel sets swapped, transposed neighbour lists, flipped taps and a
//             zero-inserting staging pass (stride 2); the reflect-padding adjoint is folded in the epilogue.
//
// Two kernels per conv:
//   1. conv_tc_prep_kernel  -- elementwise gather: applies reflect/zero padding, the decoder's x2 linear upsample and unpool
//      (fprop) or zero-insertion and LeakyReLU' (dgrad), rounds to tf32 and writes activation tiles to a staging buffer in
//      EXACTLY the shared-memory layout the MMA wants.  (HBM/L2-bound; its output stays L2-resident.)
//   2. conv_tc_kernel       -- pure TMA-fed tensor-core kernel.  Warp 0: one thread arms an mbarrier with expect_tx and issues
//      1-D bulk copies (cp.async.bulk) of the activation tile and of the packed weight pieces; warp 1: TMEM allocation and
//      single-thread tcgen05.mma issue, tcgen05.commit releases the smem stage; warps 2-5: epilogue (tcgen05.ld -> smem
//      transpose -> coalesced store with bias / LeakyReLU / reflect fold).
// Weights come from a packed, tf32-rounded copy ([block][K-chunk cb][tap][c/4][n_pad][4], i.e. one contiguous piece per
// pipeline stage and joint) refreshed by hmvae_conv_pack_weights.  Layers whose (tile, joint-group) grid cannot fill the
// 148 SMs (small B*T, big channel counts) are additionally split along the reduction (neighbour joints x channel blocks)
// over gridDim.z; the partial sums go to a workspace and conv_tc_finish_kernel adds them in a fixed order (deterministic)
// together with bias / LeakyReLU.
#include <string.h>

#include "conv_tc.cuh"

namespace hmvae {

static unsigned long long* g_tc_dbg = nullptr;

constexpr int TC_THREADS = 192;
constexpr int TC_MAX_STAGES = 6;
constexpr int TC_MAX_GJ = 32;
constexpr int TC_MAX_NB = 16;

// ---------------------------------------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// no-swizzle canonical layout descriptor: start address, LBO (K-chunk stride), SBO (8-row group stride); version 1
__device__ __forceinline__ uint64_t tc_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46);
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// ---------------------------------------------------------------------------------------------- weight packing
// Packed layout (per mode): for joint group g and K-side joint n with `cnt` consuming joints ("slots") in the group, one
// contiguous piece per K-chunk cb:   piece(g, n, cb) = [tap k][chunk h][slot][n_pad rows][4 reduction channels]
// so that (a) a pipeline stage needs ONE bulk copy for all its weights and (b) the rows of consecutive slots are contiguous,
// i.e. one tcgen05.mma can cover several output joints (N = len * n_pad).
//   fprop: rows = out channels o of joint j, reduction = in channels c of joint n.
//   dgrad: rows = in channels c of joint n,  reduction = out channels o of joint j.
// One CTA reads the contiguous ci*K runs of 4 consecutive output channels of block (j, n) once (coalesced) and writes both
// copies.  Padding rows / channels are never written: the packed buffers are zero-filled once at allocation.
struct TcPk { int base_f, cnt_f, slot_f, base_d, cnt_d, slot_d; };

__global__ void __launch_bounds__(128) conv_pack_kernel(ConvArgs a, const float* __restrict__ w, float4* __restrict__ wp_f,
                                                        float4* __restrict__ wp_d, const TcPk* __restrict__ pk, int npad_f,
                                                        int kc_f, int npad_d, int kc_d) {
  extern __shared__ float rows[];               // [4][ci*K]
  pdl_trigger();
  pdl_wait();
  const int Cin = a.J * a.ci;
  const int run = a.ci * a.K;
  const int no4 = (a.co + 3) / 4;
  const int blk = blockIdx.x / no4, o4 = blockIdx.x % no4;
  const int j = a.blk_j[blk], n = a.blk_n[blk];
  const TcPk e = pk[blk];
  // all loads of a pass are issued before the first use: the weights were just rewritten by the optimiser (HBM misses), and a
  // load -> convert -> store loop walks them one memory latency at a time (the 13.5 MB layers took 21-35 us)
  constexpr int PK_U = 6;
  for (int x0 = 0; x0 < run; x0 += PK_U * 128) {
    float v[4][PK_U];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int o = o4 * 4 + i;
      const float* src = w + ((long)(j * a.co + (o < a.co ? o : 0)) * Cin + n * a.ci) * a.K;
#pragma unroll
      for (int u = 0; u < PK_U; ++u) {
        const int x = x0 + u * 128 + threadIdx.x;
        v[i][u] = (o < a.co && x < run) ? src[x] : 0.f;
      }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int u = 0; u < PK_U; ++u) {
        const int x = x0 + u * 128 + threadIdx.x;
        if (x < run) rows[i * run + x] = __uint_as_float(to_tf32(v[i][u]));
      }
  }
  __syncthreads();
  // Thread order follows the DESTINATION's innermost index (o for the fprop copy: 4 x 16 B = 64 contiguous bytes per (tap,
  // chunk); c for the dgrad copy: ci x 16 contiguous bytes per tap), so the 16-byte stores of a warp coalesce; the shared-
  // memory reads are strided instead (stride K = 15 words: conflict-free, stride ci*K: 4-way at worst).
  if (wp_f) {   // rows = o (4 of them), 16-byte groups over c
    const int qpb = kc_f / 4, nq = (a.ci + 3) / 4;
    for (int it = threadIdx.x; it < 4 * nq * a.K; it += blockDim.x) {
      const int i = it & 3;
      const int q = (it >> 2) % nq;
      const int k = (it >> 2) / nq;
      const int o = o4 * 4 + i;
      if (o >= a.co) continue;
      float v[4];
#pragma unroll
      for (int c = 0; c < 4; ++c) v[c] = (q * 4 + c < a.ci) ? rows[i * run + (q * 4 + c) * a.K + k] : 0.f;
      const int cb = q / qpb, h = q % qpb;
      wp_f[(long)e.base_f + ((long)((cb * a.K + k) * qpb + h) * e.cnt_f + e.slot_f) * npad_f + o] = make_float4(v[0], v[1], v[2], v[3]);
    }
  }
  if (wp_d) {   // rows = c, 16-byte group = the 4 output channels of this CTA
    const int qpb = kc_d / 4;
    const int cb = o4 / qpb, h = o4 % qpb;
    for (int it = threadIdx.x; it < run; it += blockDim.x) {
      const int c = it % a.ci, k = it / a.ci;
      const int x = c * a.K + k;
      wp_d[(long)e.base_d + ((long)((cb * a.K + k) * qpb + h) * e.cnt_d + e.slot_d) * npad_d + c] =
          make_float4(rows[x], rows[run + x], rows[2 * run + x], rows[3 * run + x]);
    }
  }
}

// ---------------------------------------------------------------------------------------------- activation staging
// astage[mt][n][cb][h][rows_alloc][4]  (cb = KC-channel block, h = 16-byte chunk inside the block)
__global__ void __launch_bounds__(256) conv_tc_prep_kernel(TcArgs p, const float* __restrict__ src,
                                                           const float* __restrict__ yact, float4* __restrict__ astage) {
  pdl_trigger();
  pdl_wait();
  const ConvArgs& a = p.a;
  const int Tq = p.T + 2 * a.p;
  const int tfill = (p.mode == 0) ? (p.ntt > 1 ? 128 + a.K - 1 : Tq) : ((p.ntt > 1 ? 128 : Tq) + a.K - 1);
  const int nq = p.ck_pad / 4;             // 16-byte chunks per K-side joint
  const int qpb = p.KC / 4;                // chunks per stage block
  const long per_mt = (long)a.J * nq * p.Bt * tfill;
  const long total = per_mt * p.mtiles;
  for (long it = (long)blockIdx.x * blockDim.x + threadIdx.x; it < total; it += (long)gridDim.x * blockDim.x) {
    const int r = (int)(it % tfill);
    long rest = it / tfill;
    const int b = (int)(rest % p.Bt); rest /= p.Bt;
    const int q = (int)(rest % nq); rest /= nq;
    const int n = (int)(rest % a.J);
    const int mt = (int)(rest / a.J);
    const long bb = (p.ntt > 1) ? mt / p.ntt : (long)mt * p.Bt + b;
    const int c0 = q * 4;
    float v[4] = {0.f, 0.f, 0.f, 0.f};
    int row;
    if (p.mode == 0) {
      row = (a.s == 1) ? r * p.Bt + b : ((r & 1) * p.Tp2 + (r >> 1)) * p.Bt + b;
      const int qpos = (p.ntt > 1) ? r + (mt % p.ntt) * p.U : r;       // padded input position (time-tiled: window of tile tt)
      if (bb < p.B && qpos < Tq) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
          if (c0 + i < a.ci) v[i] = load_padded(src, a, bb, n, c0 + i, qpos, p.T);
      }
    } else {
      row = r * p.Bt + b;
      const int zz = r + tc_tile_t0(p.ntt > 1 ? mt % p.ntt : 0, p.ntt, p.U, a.p, Tq) - (a.K - 1);
      if (bb < p.B && zz >= 0 && (zz % a.s) == 0 && zz / a.s < p.T_out) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
          if (c0 + i < a.co) {
            const long oi = tc_out_index(a, bb, n, c0 + i, zz / a.s, p.T_out);
            float t = src[oi];
            if (a.lrelu && !(yact[oi] > 0.f)) t *= 0.2f;
            v[i] = t;
          }
      }
    }
    float4 o;
    o.x = __uint_as_float(to_tf32(v[0])); o.y = __uint_as_float(to_tf32(v[1]));
    o.z = __uint_as_float(to_tf32(v[2])); o.w = __uint_as_float(to_tf32(v[3]));
    const int cb = q / qpb, h = q % qpb;
    astage[((((long)mt * a.J + n) * (nq / qpb) + cb) * qpb + h) * p.rows_alloc + row] = o;
  }
}

// ---------------------------------------------------------------------------------------------- main kernel
// per joint group: which K-side joints' tiles are consumed, where their packed weights are, and the runs of consecutive
// local output joints (one MMA each)
struct TcWorkG {
  int cnt[64];
  int base[64];                 // float4 offset of piece (g, n, cb = 0)
  int nruns[64];
  unsigned int run[64][4];      // slot_first | jl_first << 8 | len << 16
};
struct TcWork {
  TcWorkG g;
};

__global__ void __launch_bounds__(TC_THREADS, 1) conv_tc_kernel(TcArgs p, const unsigned char* __restrict__ astage,
                                                                const float* __restrict__ wp,
                                                                const float* __restrict__ bias, float* __restrict__ dst) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  __shared__ uint64_t full_bar[TC_MAX_STAGES], empty_bar[TC_MAX_STAGES], accum_bar;
  __shared__ uint32_t tmem_base_s;
  __shared__ __align__(16) TcWork work;

  pdl_trigger();
  const ConvArgs& a = p.a;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) TC_STAMP(0);
  const int mt = blockIdx.x;
  const int j0 = blockIdx.y * p.GJ;
  const int gj = (a.J - j0 < p.GJ) ? a.J - j0 : p.GJ;
  const int ncb = p.ck_pad / p.KC;
  const int qpb = p.KC / 4;

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
  {
    const int4* wsrc = reinterpret_cast<const int4*>(p.wtab + blockIdx.y);
    int4* wdst = reinterpret_cast<int4*>(&work.g);
    for (int i = tid; i < (int)(sizeof(TcWorkG) / 16); i += TC_THREADS) wdst[i] = wsrc[i];
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  if (warp >= 2) {
    // zero the accumulators once: every MMA then accumulates, so one instruction may span several output joints
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    for (int c = 0; c < p.tmem_cols; c += 16) {
      asm volatile(
          "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1};" ::"r"(tmem_base + lane_base + c),
          "r"(0)
          : "memory");
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_wait();       // prologue done (barriers, TMEM, work table = plan constants): now wait for the producer of astage / wp
  if (tid == 0) TC_STAMP(1);
  const uint32_t slot_u = (uint32_t)p.n_pad;                 // 16-byte units per slot inside one (tap, chunk) row block
  const int si_beg = blockIdx.z * p.split_len, si_end = si_beg + p.split_len;

  if (warp == 0) {
    // =============================== producer: bulk copies of activation tile + weight pieces ===============================
    int s = 0, si = 0;
    uint32_t ph = 0;
    for (int n = 0; n < a.J; ++n) {
      const int cnt = work.g.cnt[n];
      if (cnt == 0) continue;
      for (int cb = 0; cb < ncb; ++cb, ++si) {
        if (si < si_beg || si >= si_end) continue;
        mbar_wait(&empty_bar[s], ph ^ 1);
        unsigned char* st = smem_raw + (size_t)s * p.stage_bytes;
        if (lane == 0) {
          const uint32_t wbytes = (uint32_t)a.K * qpb * cnt * p.n_pad * 16;      // all slots of this stage, contiguous
          mbar_arrive_expect_tx(&full_bar[s], (uint32_t)p.a_bytes + wbytes);
          bulk_g2s(st, astage + (((size_t)mt * a.J + n) * ncb + cb) * p.a_bytes, (uint32_t)p.a_bytes, &full_bar[s]);
          bulk_g2s(st + p.a_bytes, reinterpret_cast<const unsigned char*>(wp) + ((size_t)work.g.base[n] * 16 + (size_t)cb * wbytes),
                   wbytes, &full_bar[s]);
        }
        __syncwarp();
        if (++s == p.stages) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    // =============================== MMA issuer ===============================
    const uint32_t idesc0 = (1u << 4) | (2u << 7) | (2u << 10) | ((128u >> 4) << 24);      // N is filled in per run
    int s = 0, si = 0;
    uint32_t ph = 0;
    for (int n = 0; n < a.J; ++n) {
      const int cnt = work.g.cnt[n];
      if (cnt == 0) continue;
      for (int cb = 0; cb < ncb; ++cb, ++si) {
        if (si < si_beg || si >= si_end) continue;
        mbar_wait(&full_bar[s], ph);
        tc_fence_after();
        if (lane == 0) {
          if (si == si_beg) TC_STAMP(2);
          const uint32_t a_base = smem_u32(smem_raw + (size_t)s * p.stage_bytes);
          const uint32_t b_base = a_base + p.a_bytes;
          // Descriptors: only the 14-bit start-address field (16-byte units, low word) changes between MMAs, and between
          // consecutive taps it changes by a constant, so the issue loop is two adds + one tcgen05.mma per instruction.
          //   fprop stride 1: tap k reads rows shifted by k*Bt           -> one run  k = 0..K-1,      a += Bt
          //   fprop stride 2: even taps read phase 0, odd taps phase 1    -> two runs (k even / odd),  a += Bt, b += 2 taps
          //   dgrad         : tap k reads rows shifted by (K-1-k)*Bt      -> one run  k = 0..K-1,      a -= Bt
          // B rows of consecutive slots are contiguous: one MMA covers a whole run of consecutive output joints.
          const uint32_t row_u = (uint32_t)cnt * slot_u;                       // 16-byte units of one (tap, chunk) row block
          const uint64_t adesc0 = tc_desc(a_base, (uint32_t)p.rows_alloc * 16, 128);
          const uint64_t bdesc0 = tc_desc(b_base, row_u * 16, 128);
          const uint64_t a_kk = 2u * (uint32_t)p.rows_alloc, b_kk = 2u * row_u;   // next 8 reduction channels
          const uint32_t tap_u = (uint32_t)qpb * row_u;
          const int ntaprun = (p.mode == 0 && a.s == 2) ? 2 : 1;
          const int nkk = qpb / 2;
          for (int r = 0; r < work.g.nruns[n]; ++r) {
            const uint32_t rr = work.g.run[n][r];
            const uint32_t slot0 = rr & 0xff, jl0 = (rr >> 8) & 0xff, len = rr >> 16;
            const uint32_t d_addr = tmem_base + jl0 * (uint32_t)p.n_pad;
            const uint32_t idesc = idesc0 | (((len * (uint32_t)p.n_pad) >> 3) << 17);
            for (int run = 0; run < ntaprun; ++run) {
              uint64_t ad, bd = bdesc0 + (uint64_t)(slot0 * slot_u + (uint32_t)run * tap_u);
              long a_step;
              int ntap;
              if (p.mode == 1) { ad = adesc0 + (uint32_t)((a.K - 1) * p.Bt); a_step = -(long)p.Bt; ntap = a.K; }
              else if (ntaprun == 1) { ad = adesc0; a_step = p.Bt; ntap = a.K; }
              else { ad = adesc0 + (uint32_t)(run * p.Tp2 * p.Bt); a_step = p.Bt; ntap = (a.K - run + 1) / 2; }
              const uint64_t b_step = (uint64_t)tap_u * ntaprun;
              if (nkk == 1) {
#pragma unroll 5
                for (int tp = 0; tp < ntap; ++tp) {
                  tc_mma_tf32(d_addr, ad, bd, idesc, 1u);
                  ad += a_step;
                  bd += b_step;
                }
              } else {
                for (int tp = 0; tp < ntap; ++tp) {
                  uint64_t ad2 = ad, bd2 = bd;
#pragma unroll 4
                  for (int kk = 0; kk < nkk; ++kk) {
                    tc_mma_tf32(d_addr, ad2, bd2, idesc, 1u);
                    ad2 += a_kk;
                    bd2 += b_kk;
                  }
                  ad += a_step;
                  bd += b_step;
                }
              }
            }
          }
          tc_commit(&empty_bar[s]);
        }
        __syncwarp();
        if (++s == p.stages) { s = 0; ph ^= 1; }
      }
    }
    if (lane == 0) { tc_commit(&accum_bar); TC_STAMP(3); }
    __syncwarp();
  } else {
    // =============================== epilogue (warps 2..5): TMEM -> accumulator dump ===============================
    // Every (tile, group, split) CTA dumps its raw fp32 accumulators row-major, dump[split][tile][group][128 rows][dcols]:
    // a thread owns one accumulator row and writes 64 contiguous bytes per tcgen05.ld -- no shared-memory transpose, no index
    // arithmetic per element (the old NCW-scattering epilogue was ~300 SASS instructions per element and dominated the life of a
    // CTA at B=32).  Layout conversion, split-K sum, bias / LeakyReLU / pooling / reflect fold happen in conv_tc_finish_kernel.
    mbar_wait(&accum_bar, 0);
    tc_fence_after();
    if (tid == 64) TC_STAMP(4);
    const int lq = warp & 3;
    const int m = lq * 32 + lane;
    float4* drow = reinterpret_cast<float4*>(dst + ((((size_t)blockIdx.z * p.mtiles + mt) * p.groups + blockIdx.y) * 128 + m) * p.dcols);
    const int ncols = gj * p.n_pad;
    for (int c16 = 0; c16 < ncols; c16 += 16) {
      float v[16];
      tmem_ld16(tmem_base + ((uint32_t)(lq * 32) << 16) + (uint32_t)c16, v);
#pragma unroll
      for (int q = 0; q < 4; ++q) drow[(c16 >> 2) + q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
    }
  }
  // ---- teardown
  if (tid == 64) TC_STAMP(5);
  tc_fence_before();
  __syncthreads();
  if (tid == 0) TC_STAMP(6);
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(p.tmem_cols) : "memory");
  }
}

// ---------------------------------------------------------------------------------------------- finish
// Accumulator dump -> result tensor.  One thread per result element (time fastest, so the NCW stores coalesce):
//   fprop: y[b, j, o, t]  = act(sum_z dump[z][tile(b)][group(j)][t*Bt + b'][jl*n_pad + o] + bias)
//   dgrad: dx[b, n, c, u] = sum_z (row(u + p) + reflect-fold rows) of the same dump
// The split-K partial sums are added in a fixed order (deterministic).
__global__ void __launch_bounds__(256) conv_tc_finish_kernel(TcArgs p, const float* __restrict__ dump, const float* __restrict__ bias,
                                                             float* __restrict__ dst) {
  pdl_trigger();
  pdl_wait();
  const ConvArgs& a = p.a;
  const int Tr = (p.mode == 0) ? p.T_out : p.T;
  const int C = a.J * p.n_real;
  const long per = (long)p.B * C * Tr;
  for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < per; e += (long)gridDim.x * blockDim.x) {
    const int t = (int)(e % Tr);
    const long r = e / Tr;
    const int ch = (int)(r % C);
    const int b = (int)(r / C);
    const int j = ch / p.n_real, o = ch - j * p.n_real;
    if (p.mode == 0) {
      float v = tc_fprop_value(p, dump, bias, b, j, o, t);
      if (a.lrelu) v = lrelu_f(v, 0.2f);
      dst[tc_out_index(a, b, j, o, t, p.T_out)] = v;
    } else {
      dst[e] = tc_dgrad_value(p, dump, b, j, o, t);
    }
  }
}

// ---------------------------------------------------------------------------------------------- host side
static inline int rup(int v, int m) { return (v + m - 1) / m * m; }

// reduction channels per stage: a property of the layer (the packed weight layout depends on it), not of the batch
static int tc_pick_kc(int ck_pad, int n_pad, int K) {
  const int kcs[4] = {32, 24, 16, 8};
  for (int i = 0; i < 4; ++i)
    if (ck_pad % kcs[i] == 0 && K * (kcs[i] / 4) * n_pad * 16 <= 16 * 1024) return kcs[i];
  return 8;
}

// Batch-independent description of one (plan, mode): channel padding, joint groups, packed-weight layout, work tables.
struct TcLayer {
  bool ok;
  int n_real, n_pad, ck, ck_pad, KC, GJ, groups, nbmax, longest;
  long packed_f4;                       // size of the packed weights in float4
  TcWorkG* dev_work;                    // [groups]
  std::vector<std::vector<int>> slots;  // [g * J + n] -> local joints jl (ascending) consuming tile n
  std::vector<long> base;               // [g * J + n] -> float4 offset of piece (g, n, 0)
};
struct TcLayers {
  TcLayer m[2];
  TcPk* dev_pk;
};

static void tc_build_layer(const hmvae_conv_plan* plan, int mode, TcLayer* L) {
  const ConvArgs& a = plan->a;
  L->ok = false;
  L->dev_work = nullptr;
  L->n_real = mode == 0 ? a.co : a.ci;
  L->ck = mode == 0 ? a.ci : a.co;
  L->n_pad = rup(L->n_real, 16);
  L->ck_pad = rup(L->ck, 8);
  if (a.J > 64 || L->n_pad > 256) return;
  L->KC = tc_pick_kc(L->ck_pad, L->n_pad, a.K);
  L->GJ = env_int("HMVAE_TC_GROUP_COLS", 64) / L->n_pad;      // output joints per CTA (accumulator columns / n_pad)
  if (L->GJ < 1) L->GJ = 1;
  if (L->GJ > env_int("HMVAE_TC_GROUP_MAX", 4)) L->GJ = env_int("HMVAE_TC_GROUP_MAX", 4);
  if (L->GJ > a.J) L->GJ = a.J;
  L->groups = (a.J + L->GJ - 1) / L->GJ;
  // N-side joint -> K-side joints (fprop: neighbour list; dgrad: transpose, ascending)
  std::vector<std::vector<int>> lists(a.J);
  for (int j = 0; j < a.J; ++j)
    for (int m = plan->nb_off[j]; m < plan->nb_off[j + 1]; ++m) {
      if (mode == 0) lists[j].push_back(plan->nb_idx[m]);
      else lists[plan->nb_idx[m]].push_back(j);
    }
  // ---- dense joint groups for narrow layers.  A kind::tf32 MMA costs the same ~63 cycles for every N <= 128, and with few
  // channels per joint (6 or 12) the sparse formulation issues one N = 16..64 MMA per RUN of consecutive output joints that share
  // an input joint -- the last decoder level is then bound by the MMA COUNT (4050 per 128-row tile; tools/tc_phases.py: 21 of its
  // 32 us are issue time).  Here a group is GJ output joints packed at 8-channel granularity (N = GJ * n_pad ~ 96 columns), every
  // input joint in the union of the group's neighbour lists feeds ONE MMA over the whole group, and the weights of the (joint,
  // input joint) pairs that are masked are zeros in the packed copy (never written: the buffer is zero-initialised).  Chosen per
  // (layer, mode) when it cuts the MMA count by >= 1.4x (HMVAE_TC_DENSE: -1 auto, 0 never, 1 whenever the layer is narrow).
  bool dense = false;
  {
    const int want = env_int("HMVAE_TC_DENSE", -1);
    const int np8 = rup(L->n_real, 8);
    const int cols = env_int("HMVAE_TC_DENSE_COLS", 96);
    int gj = cols / np8;
    gj &= ~1;                                        // N = gj * np8 must be a multiple of 16
    if (want != 0 && np8 <= env_int("HMVAE_TC_DENSE_MAXNP", 16) && gj >= 2 && a.J > L->GJ) {
      long sparse_units = 0, dense_units = 0;
      for (int g = 0; g * L->GJ < a.J; ++g)
        for (int n = 0; n < a.J; ++n) {
          int prev = -2, runs = 0;
          for (int jl = 0; jl < L->GJ && g * L->GJ + jl < a.J; ++jl)
            for (int kn : lists[g * L->GJ + jl])
              if (kn == n) { runs += (jl != prev + 1); prev = jl; }
          sparse_units += runs;
        }
      for (int g = 0; g * gj < a.J; ++g)
        for (int n = 0; n < a.J; ++n) {
          bool any = false;
          for (int jl = 0; jl < gj && g * gj + jl < a.J; ++jl)
            for (int kn : lists[g * gj + jl]) any |= kn == n;
          dense_units += any;
        }
      dense = want == 1 || dense_units * 14 <= sparse_units * 10;
      if (dense) {
        L->n_pad = np8;
        L->GJ = gj;
        L->groups = (a.J + gj - 1) / gj;
        // one stage = all GJ slots: keep it around 48 KB
        L->KC = 8;
        const int kcs[3] = {32, 24, 16};
        for (int i = 2; i >= 0; --i)
          if (L->ck_pad % kcs[i] == 0 && (long)gj * a.K * (kcs[i] / 4) * np8 * 16 <= 48 * 1024) L->KC = kcs[i];
      }
    }
  }
  const int ncb = L->ck_pad / L->KC, qpb = L->KC / 4;
  L->slots.assign((size_t)L->groups * a.J, {});
  L->base.assign((size_t)L->groups * a.J, 0);
  std::vector<TcWorkG> host(L->groups);
  long off = 0;
  L->nbmax = 0;
  L->longest = 0;
  for (int g = 0; g < L->groups; ++g) {
    TcWorkG& w = host[g];
    memset(&w, 0, sizeof(w));
    int stages = 0;
    for (int n = 0; n < a.J; ++n) {
      std::vector<int>& sl = L->slots[(size_t)g * a.J + n];
      for (int jl = 0; jl < L->GJ && g * L->GJ + jl < a.J; ++jl)
        for (int kn : lists[g * L->GJ + jl])
          if (kn == n) sl.push_back(jl);
      if (dense && !sl.empty()) {                      // every joint of the group is a slot (absent pairs: zero weights)
        sl.clear();
        for (int jl = 0; jl < L->GJ; ++jl) sl.push_back(jl);
      }
      const int cnt = (int)sl.size();
      w.cnt[n] = cnt;
      w.base[n] = (int)off;
      L->base[(size_t)g * a.J + n] = off;
      if (cnt == 0) continue;
      if (cnt > L->nbmax) L->nbmax = cnt;
      stages += ncb;
      int nr = 0;
      for (int i = 0; i < cnt;) {
        int len = 1;
        while (i + len < cnt && sl[i + len] == sl[i] + len) ++len;
        w.run[n][nr++] = (unsigned)i | ((unsigned)sl[i] << 8) | ((unsigned)len << 16);
        i += len;
      }
      w.nruns[n] = nr;
      off += (long)ncb * a.K * qpb * cnt * L->n_pad;
    }
    if (stages > L->longest) L->longest = stages;
  }
  if (off * 16 >= (1L << 31)) return;
  L->packed_f4 = off;
  if (cudaMalloc(&L->dev_work, L->groups * sizeof(TcWorkG)) != cudaSuccess) return;
  if (cudaMemcpy(L->dev_work, host.data(), L->groups * sizeof(TcWorkG), cudaMemcpyHostToDevice) != cudaSuccess) return;
  L->ok = true;
}

static const TcLayers* tc_layers(const hmvae_conv_plan* plan) {
  auto it = plan->tc_tables.find(0);
  if (it != plan->tc_tables.end()) return reinterpret_cast<const TcLayers*>(it->second);
  TcLayers* LL = new TcLayers();
  tc_build_layer(plan, 0, &LL->m[0]);
  tc_build_layer(plan, 1, &LL->m[1]);
  LL->dev_pk = nullptr;
  const ConvArgs& a = plan->a;
  if (LL->m[0].ok && LL->m[1].ok) {
    std::vector<TcPk> pk(a.nnz);
    for (int j = 0; j < a.J; ++j)
      for (int m = plan->nb_off[j]; m < plan->nb_off[j + 1]; ++m) {
        const int n = plan->nb_idx[m];
        TcPk e;
        {
          const TcLayer& L = LL->m[0];
          const int g = j / L.GJ, jl = j % L.GJ;
          const std::vector<int>& sl = L.slots[(size_t)g * a.J + n];
          e.base_f = (int)L.base[(size_t)g * a.J + n];
          e.cnt_f = (int)sl.size();
          e.slot_f = 0;
          for (size_t i = 0; i < sl.size(); ++i) if (sl[i] == jl) e.slot_f = (int)i;
        }
        {
          const TcLayer& L = LL->m[1];
          const int g = n / L.GJ, nl = n % L.GJ;
          const std::vector<int>& sl = L.slots[(size_t)g * a.J + j];
          e.base_d = (int)L.base[(size_t)g * a.J + j];
          e.cnt_d = (int)sl.size();
          e.slot_d = 0;
          for (size_t i = 0; i < sl.size(); ++i) if (sl[i] == nl) e.slot_d = (int)i;
        }
        pk[m] = e;
      }
    if (cudaMalloc(&LL->dev_pk, pk.size() * sizeof(TcPk)) != cudaSuccess ||
        cudaMemcpy(LL->dev_pk, pk.data(), pk.size() * sizeof(TcPk), cudaMemcpyHostToDevice) != cudaSuccess)
      LL->m[0].ok = LL->m[1].ok = false;
  }
  plan->tc_tables[0] = LL;
  return LL;
}

void conv_tc_release(const hmvae_conv_plan* plan) {
  auto it = plan->tc_tables.find(0);
  if (it == plan->tc_tables.end()) return;
  TcLayers* LL = reinterpret_cast<TcLayers*>(it->second);
  for (int m = 0; m < 2; ++m)
    if (LL->m[m].dev_work) cudaFree(LL->m[m].dev_work);
  if (LL->dev_pk) cudaFree(LL->dev_pk);
  delete LL;
  plan->tc_tables.erase(it);
}

static bool tc_geometry_build(const hmvae_conv_plan* plan, int B, int T, int mode, TcArgs* out);

// geometry is a pure function of (plan, mode, B, T): memoised, so the per-call host cost is one map lookup
bool tc_geometry(const hmvae_conv_plan* plan, int B, int T, int mode, TcArgs* out) {
  static thread_local std::map<std::pair<unsigned long long, long>, std::pair<bool, TcArgs>> cache;
  const auto key = std::make_pair(plan->uid, ((long)mode << 48) | ((long)B << 20) | (long)T);
  auto it = cache.find(key);
  if (it == cache.end()) {
    TcArgs p;
    memset(&p, 0, sizeof(p));
    const bool ok = tc_geometry_build(plan, B, T, mode, &p);
    it = cache.emplace(key, std::make_pair(ok, p)).first;
  }
  if (it->second.first) *out = it->second.second;
  return it->second.first;
}

static bool tc_geometry_build(const hmvae_conv_plan* plan, int B, int T, int mode, TcArgs* out) {
  const ConvArgs& a = plan->a;
  const TcLayer& L = tc_layers(plan)->m[mode];
  if (!L.ok) return false;
  TcArgs p;
  memset(&p, 0, sizeof(p));
  p.a = a;
  p.mode = mode;
  p.B = B;
  p.T = T;
  p.T_out = conv_t_out(plan->d, T);
  const int Tq = T + 2 * a.p;
  p.Tt = (mode == 0) ? p.T_out : Tq;
  p.n_real = L.n_real; p.n_pad = L.n_pad; p.ck = L.ck; p.ck_pad = L.ck_pad; p.KC = L.KC; p.GJ = L.GJ; p.nbmax = L.nbmax;
  p.wtab = L.dev_work;
  p.groups = L.groups;
  p.dcols = L.GJ * L.n_pad;
  p.ntt = 1;
  p.U = (mode == 0) ? p.T_out : T;
  if (p.Tt < 1) return false;
  if (p.Tt > 128 && mode == 0) {
    // fprop with more than 128 output steps per sequence (stride 1): tile tt of a sequence owns the output steps
    // [tt*128, tt*128 + 128) and stages the 128 + K - 1 padded input rows they read (consecutive tiles overlap by K - 1 rows)
    if (a.s != 1) return false;
    p.ntt = (p.T_out + 127) / 128;
    p.U = 128;
    p.Bt = 1;
  } else if (p.Tt > 128) {
    // dgrad of long sequences (trajectory model: T = 128, K = 31 => 158 padded rows): tile over time, one sequence per tile
    int n = 2;
    for (; n <= 64; ++n) {
      const int U = (T + n - 1) / n;
      if (U + 2 * a.p <= 128 && U >= a.p + 1) break;
    }
    if (n > 64 || 2 * a.p + 1 > 128) return false;
    p.ntt = n;
    p.U = (T + n - 1) / n;
    p.Bt = 1;
  } else {
    p.Bt = 128 / p.Tt;
    if (p.Bt > B) p.Bt = B;
  }
  p.Tp2 = (Tq + 1) / 2;
  if (mode == 0) {
    if (a.s == 1) p.rows_alloc = (a.K - 1) * p.Bt + 128;
    else {
      const int need = (p.Tp2 + (a.K - 1) / 2) * p.Bt + 128;
      p.rows_alloc = 2 * p.Tp2 * p.Bt > need ? 2 * p.Tp2 * p.Bt : need;
    }
    const int fill = (p.ntt > 1) ? 128 + a.K - 1 : ((a.s == 1) ? Tq * p.Bt : 2 * p.Tp2 * p.Bt);
    if (fill > p.rows_alloc) p.rows_alloc = fill;
  } else {
    p.rows_alloc = (a.K - 1) * p.Bt + 128;
    const int fill = ((p.ntt > 1 ? 128 : Tq) + a.K - 1) * p.Bt;
    if (fill > p.rows_alloc) p.rows_alloc = fill;
  }
  p.rows_alloc = rup(p.rows_alloc, 8);
  if (p.rows_alloc * 16 >= (1 << 18)) return false;
  p.mtiles = (p.ntt > 1) ? B * p.ntt : (B + p.Bt - 1) / p.Bt;
  p.a_bytes = (p.KC / 4) * p.rows_alloc * 16;
  p.stage_bytes = rup(p.a_bytes + L.nbmax * a.K * (p.KC / 4) * p.n_pad * 16, 128);
  const int budget = env_int("HMVAE_TC_SMEM_KB", 208) * 1024;     // dynamic shared memory per CTA (stage ring)
  p.stages = budget / p.stage_bytes;
  {
    // Measured (profiles/r01_summary_v3.md): the K loop of a CTA is only 3-10 stages long at B=32, so a deep ring buys nothing,
    // while 2 stages (50-120 KB) let two CTAs -- of this kernel, or of the weight-gradient kernel running on the other stream --
    // share an SM: 1.197 -> 1.112 ms per step.
    int cap = env_int("HMVAE_TC_STAGES", 2);
    if (cap > TC_MAX_STAGES) cap = TC_MAX_STAGES;
    if (cap < 2) cap = 2;
    if (p.stages > cap) p.stages = cap;
  }
  if (p.stages < 2) return false;
  int cols = p.GJ * p.n_pad, pow2 = 32;
  while (pow2 < cols) pow2 <<= 1;
  if (pow2 > 512) return false;
  p.tmem_cols = pow2;
  // split-K so that the grid fills one wave of SMs (fixed per-CTA cost dominates beyond that); >= 3 stages per CTA
  const int ctas = p.mtiles * L.groups;
  // two CTAs fit an SM (2-stage rings): aim at 2 x SMs CTAs (measured: 148 / 296 / 444 / 592 -> 0.975 / 0.894 / 0.946 / 0.989 ms)
  int splits = env_int("HMVAE_TC_TARGET_CTAS", 2 * num_sms()) / ctas;
  const int min_len = env_int("HMVAE_TC_MIN_SPLIT_LEN", 3);
  if (splits > L.longest / min_len) splits = L.longest / min_len;
  if (splits < 1) splits = 1;
  if (splits > 64) splits = 64;
  p.split_len = (L.longest + splits - 1) / splits;
  p.splits = (L.longest + p.split_len - 1) / p.split_len;
  *out = p;
  return true;
}

bool conv_tc_supported(const hmvae_conv_plan* plan, int B, int T, int mode) {
  TcArgs p;
  return tc_geometry(plan, B, T, mode, &p);
}

long tc_stage_ws(const TcArgs& p) { return (long)p.mtiles * p.a.J * (p.ck_pad / p.KC) * p.a_bytes; }
long tc_part_ws(const TcArgs& p) {      // accumulator dump [splits][mtiles][groups][128][dcols]
  return (long)p.splits * p.mtiles * p.groups * 128 * p.dcols * 4;
}

long conv_tc_workspace_bytes(const hmvae_conv_plan* plan, int B, int T, int mode) {
  TcArgs p;
  if (!tc_geometry(plan, B, T, mode, &p)) return -1;
  return tc_stage_ws(p) + tc_part_ws(p);
}

bool conv_tc_sizes(const hmvae_conv_plan* plan, int B, int T, int mode, long* stage_bytes, long* dump_bytes) {
  TcArgs p;
  if (!tc_geometry(plan, B, T, mode, &p)) return false;
  *stage_bytes = tc_stage_ws(p);
  *dump_bytes = tc_part_ws(p);
  return true;
}

void conv_packed_sizes(const hmvae_conv_plan* plan, long* n_fprop, long* n_dgrad) {
  const TcLayers* LL = tc_layers(plan);
  *n_fprop = LL->m[0].ok ? LL->m[0].packed_f4 * 4 : 0;
  *n_dgrad = LL->m[1].ok ? LL->m[1].packed_f4 * 4 : 0;
}

int conv_pack(const hmvae_conv_plan* plan, const float* w, float* wp_f, float* wp_d, cudaStream_t st) {
  const ConvArgs& a = plan->a;
  const TcLayers* LL = tc_layers(plan);
  if (!LL->m[0].ok || !LL->m[1].ok || !LL->dev_pk) return fail_arg("conv_pack_weights: layer not supported by the tcgen05 path");
  if (!wp_f && !wp_d) return 0;
  if ((wp_f && !aligned16(wp_f)) || (wp_d && !aligned16(wp_d))) return fail_arg("conv_pack_weights: buffers must be 16-byte aligned");
  const size_t smem = (size_t)4 * a.ci * a.K * 4;
  if (smem > 48 * 1024) HMVAE_CUDA(cudaFuncSetAttribute(conv_pack_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  launch_pdl(conv_pack_kernel, dim3(a.nnz * ((a.co + 3) / 4)), dim3(128), smem, st, a, w, reinterpret_cast<float4*>(wp_f),
             reinterpret_cast<float4*>(wp_d), LL->dev_pk, LL->m[0].n_pad, LL->m[0].KC, LL->m[1].n_pad, LL->m[1].KC);
  return check_launch("conv_pack");
}

// The three phases of one conv, separately launchable (the stack-level path in ops.py replaces `stage` and `finish` of
// neighbouring layers by ONE link kernel, conv_link.cu).
int conv_tc_stage(const hmvae_conv_plan* plan, int mode, const float* src, const float* yact, int B, int T, void* stage_ws,
                  cudaStream_t st) {
  TcArgs p;
  if (!tc_geometry(plan, B, T, mode, &p)) return fail_arg("conv (tcgen05): unsupported geometry");
  if (!stage_ws || !aligned16(stage_ws)) return fail_arg("conv (tcgen05): staging workspace missing / misaligned");
  const int Tq = T + 2 * p.a.p;
  const long items = (long)p.mtiles * p.a.J * (p.ck_pad / 4) * p.Bt * (mode == 0 ? Tq : Tq + p.a.K - 1);
  long blocks = (items + 255) / 256, cap = (long)num_sms() * 8;
  launch_pdl(conv_tc_prep_kernel, dim3((int)(blocks < cap ? blocks : cap)), dim3(256), 0, st, p, src, yact,
             reinterpret_cast<float4*>(stage_ws));
  return check_launch("conv_tc_prep");
}

int conv_tc_run(const hmvae_conv_plan* plan, int mode, const float* wp, int B, int T, const void* stage_ws, void* dump_ws,
                cudaStream_t st) {
  TcArgs p;
  if (!tc_geometry(plan, B, T, mode, &p)) return fail_arg("conv (tcgen05): unsupported geometry");
  p.dbg = g_tc_dbg;
  if (!stage_ws || !dump_ws || !aligned16(stage_ws) || !aligned16(dump_ws) || !aligned16(wp))
    return fail_arg("conv (tcgen05): workspace / packed weights must be 16-byte aligned");
  const size_t smem = (size_t)p.stages * p.stage_bytes + 1024;
  // the limit is a per-function permission, sticky and process-wide (forward and autograd-engine threads both launch): raise it
  // once to the device maximum (opt-in limit minus the kernel's static shared memory) instead of per launch
  static std::atomic<long> smem_max{0};
  if (smem_max.load(std::memory_order_acquire) == 0) {
    int dev = 0, optin = 0;
    cudaFuncAttributes fa;
    HMVAE_CUDA(cudaGetDevice(&dev));
    HMVAE_CUDA(cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    HMVAE_CUDA(cudaFuncGetAttributes(&fa, conv_tc_kernel));
    const long lim = (long)optin - (long)fa.sharedSizeBytes;
    HMVAE_CUDA(cudaFuncSetAttribute(conv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lim));
    smem_max.store(lim, std::memory_order_release);
  }
  if ((long)smem > smem_max.load(std::memory_order_acquire)) return fail_arg("conv (tcgen05): stage ring exceeds the shared memory of an SM");
  dim3 grid(p.mtiles, p.groups, p.splits);
  launch_pdl<true>(conv_tc_kernel, grid, dim3(TC_THREADS), smem, st, p, reinterpret_cast<const unsigned char*>(stage_ws), wp,
                   (const float*)nullptr, reinterpret_cast<float*>(dump_ws));
  return check_launch(mode == 0 ? "conv_fprop_tc" : "conv_dgrad_tc");
}

int conv_tc_finish(const hmvae_conv_plan* plan, int mode, const void* dump_ws, const float* bias, float* dst, int B, int T,
                   cudaStream_t st) {
  TcArgs p;
  if (!tc_geometry(plan, B, T, mode, &p)) return fail_arg("conv (tcgen05): unsupported geometry");
  const long per = (long)p.B * p.a.J * p.n_real * (mode == 0 ? p.T_out : p.T);
  long blocks = (per + 255) / 256, cap = (long)num_sms() * 8;
  launch_pdl(conv_tc_finish_kernel, dim3((int)(blocks < cap ? blocks : cap)), dim3(256), 0, st, p, (const float*)dump_ws, bias, dst);
  return check_launch("conv_tc_finish");
}

int conv_tc_launch(const hmvae_conv_plan* plan, int mode, const float* src, const float* yact, const float* wp,
                   const float* bias, float* dst, int B, int T, void* workspace, long workspace_bytes, cudaStream_t st) {
  TcArgs p;
  if (!tc_geometry(plan, B, T, mode, &p)) return fail_arg("conv (tcgen05): unsupported geometry");
  const long need = tc_stage_ws(p) + tc_part_ws(p);
  if (!workspace || workspace_bytes < need) return fail_arg("conv (tcgen05): workspace too small");
  void* part = reinterpret_cast<unsigned char*>(workspace) + tc_stage_ws(p);
  int rc = conv_tc_stage(plan, mode, src, yact, B, T, workspace, st);
  if (rc) return rc;
  rc = conv_tc_run(plan, mode, wp, B, T, workspace, part, st);
  if (rc) return rc;
  return conv_tc_finish(plan, mode, part, bias, dst, B, T, st);
}

}  // namespace hmvae

// tools/tc_phases.py: device buffer of 8 x uint64 per CTA that receives globaltimer stamps of the next conv_tc launches
// (0 entry, 1 prologue done, 2 first stage landed, 3 last MMA issued, 4 accumulators complete, 5 dump stored, 6 exit); NULL = off.
extern "C" int hmvae_conv_tc_debug(void* buf) {
  hmvae::g_tc_dbg = reinterpret_cast<unsigned long long*>(buf);
  return 0;
}
