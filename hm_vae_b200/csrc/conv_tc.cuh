// Shared between the tcgen05 conv kernels (conv_tc.cu) and the inter-layer link kernel (conv_link.cu): geometry of one
// (plan, mode, B, T) launch, the accumulator-dump indexing and the staged-tile indexing.
#pragma once
#include "conv_common.cuh"

namespace hmvae {

struct TcArgs {
  ConvArgs a;
  int mode;            // 0 fprop, 1 dgrad
  int n_real, n_pad;   // N-side channels per joint (real, padded to 16)
  int ck, ck_pad;      // reduction channels per K-side joint (real, padded to 8)
  int KC;              // reduction channels per pipeline stage (multiple of 8, divides ck_pad)
  int Bt, Tt;          // sequences per tile, rows-per-sequence (M = Tt*Bt <= 128)
  int rows_alloc;      // rows per 16-byte chunk column of an activation tile
  int Tp2;             // fprop stride 2: rows per phase / Bt
  int ntt, U;          // dgrad time tiling (T + 2p > 128): tiles per sequence (1 = none), result time steps owned by a tile
  int GJ, nbmax, stages;
  int B, T, T_out, mtiles;
  int a_bytes, stage_bytes;
  int tmem_cols;
  int splits, split_len;   // split-K: gridDim.z CTAs per (tile, group), each handles split_len consecutive stages
  int groups, dcols;       // joint groups (gridDim.y); accumulator columns per group (= GJ * n_pad) in the dump
  const struct TcWorkG* wtab;   // [groups] host-built work tables (device memory, owned by the plan)
  unsigned long long* dbg;      // optional (tools/tc_phases.py): 8 globaltimer stamps per CTA
};
__device__ __forceinline__ unsigned long long gtimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
#define TC_STAMP(i) do { if (p.dbg) p.dbg[(((size_t)blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x) * 8 + (i)] = gtimer(); } while (0)


__device__ __forceinline__ uint32_t to_tf32(float x) {
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
  return u;
}

// dgrad time tiling: a tile is a window of 128 consecutive rows (padded input positions) of ONE sequence.  Tile tt owns the
// result steps u in [tt*U, (tt+1)*U); its window starts at 0 for the first tile, ends at Tq for the last one (so that the rows the
// reflect-padding fold needs are inside), and is centred on the owned rows otherwise.
__host__ __device__ __forceinline__ int tc_tile_t0(int tt, int ntt, int U, int pad, int Tq) {
  if (ntt <= 1 || tt == 0) return 0;
  if (tt == ntt - 1) return Tq - 128;
  int t0 = tt * U + pad - (128 - U) / 2;
  if (t0 < 0) t0 = 0;
  if (t0 > Tq - 128) t0 = Tq - 128;
  return t0;
}

__device__ __forceinline__ long tc_out_index(const ConvArgs& a, long b, int j, int o, int t, int T_out) {
  const int ch = j * a.ojs + a.oco + o;
  const long ctot = (long)a.J * a.ojs;
  return a.cl ? (b * T_out + t) * ctot + ch : (b * ctot + ch) * T_out + t;
}


// ---------------------------------------------------------------------------------------------- accumulator dump -> values
__device__ __forceinline__ float tc_dump_sum(const TcArgs& p, const float* __restrict__ dump, int mt, int g, int row, int col) {
  const size_t zstride = (size_t)p.mtiles * p.groups * 128 * p.dcols;
  const float* d = dump + (((size_t)mt * p.groups + g) * 128 + row) * p.dcols + col;
  float v = 0.f;
  for (int z = 0; z < p.splits; ++z) v += d[z * zstride];
  return v;
}

// fprop result before the activation: conv sum + bias of conv-output joint j, channel o, sequence b, step t
__device__ __forceinline__ float tc_fprop_value(const TcArgs& p, const float* __restrict__ dump, const float* __restrict__ bias,
                                                int b, int j, int o, int t) {
  int mt, row;
  if (p.ntt > 1) {                       // time-tiled fprop: tile tt of sequence b holds the output steps [tt*U, tt*U + 128)
    const int tt = t / p.U;
    mt = b * p.ntt + tt;
    row = t - tt * p.U;
  } else {
    mt = b / p.Bt;
    row = t * p.Bt + (b - mt * p.Bt);
  }
  float v = tc_dump_sum(p, dump, mt, j / p.GJ, row, (j % p.GJ) * p.n_pad + o);
  if (bias) v += bias[j * p.a.co + o];
  return v;
}

// dgrad result: gradient w.r.t. the (virtual) conv input of joint n, channel c, sequence b, step u (padding adjoint folded in)
__device__ __forceinline__ float tc_dgrad_value(const TcArgs& p, const float* __restrict__ dump, int b, int n, int c, int u) {
  const ConvArgs& a = p.a;
  int mt, bl, t0 = 0;
  if (p.ntt > 1) {
    const int tt = u / p.U;
    mt = b * p.ntt + tt;
    bl = 0;
    t0 = tc_tile_t0(tt, p.ntt, p.U, a.p, p.T + 2 * a.p);
  } else {
    mt = b / p.Bt;
    bl = b - mt * p.Bt;
  }
  const int g = n / p.GJ, col = (n % p.GJ) * p.n_pad + c;
  float v = tc_dump_sum(p, dump, mt, g, (u + a.p - t0) * p.Bt + bl, col);
  if (a.pad_mode == 1) {
    if (u >= 1 && u <= a.p) v += tc_dump_sum(p, dump, mt, g, (a.p - u - t0) * p.Bt + bl, col);
    if (u <= p.T - 2 && u >= p.T - 1 - a.p) v += tc_dump_sum(p, dump, mt, g, (a.p + 2 * (p.T - 1) - u - t0) * p.Bt + bl, col);
  }
  return v;
}


// geometry is a pure function of (plan, mode, B, T): memoised
bool tc_geometry(const hmvae_conv_plan* plan, int B, int T, int mode, TcArgs* out);
long tc_stage_ws(const TcArgs& p);      // bytes of the staged activation tiles
long tc_part_ws(const TcArgs& p);       // bytes of the accumulator dump [splits][mtiles][groups][128][dcols]

}  // namespace hmvae
