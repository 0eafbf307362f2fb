// Forward kinematics (fwd/bwd), rot6d -> rotmat (fwd/bwd), axis-angle -> rotmat.   sm_100a, CUDA cores, HBM-bound.
//
// Replaces fk_layer.py:47-93 (23 gather+bmm+slice-copy triplets per call, CopySlices chain in backward),
// my_tools.py:19-39 (~15 elementwise launches) and torchgeometry.angle_axis_to_rotation_matrix.
//
// Algorithmic bytes (fp32): FK fwd 48 B/joint-frame (36 in + 12 out), bwd 84 B/jf (12 dpos + 36 R in, 36 dR out);
// with 6D input 36 / 60 B/jf.  rot6d fwd 60, bwd 84 B per matrix.
#include "fk_device.cuh"

namespace hmvae {

// ------------------------------------------------------------------------------------------------ staging
// One warp stages `nvalid` rows of `rowf` floats (contiguous in global memory) into padded smem rows.
template <bool BULK>
__device__ __forceinline__ void stage_in(float* rows, int pitch, const float* g, int rowf, int nvalid, int lane,
                                         uint64_t* bar) {
  if (BULK) {
    if (lane < nvalid) bulk_g2s(rows + lane * pitch, g + (size_t)lane * rowf, rowf * 4, bar);
  } else {
    const int total = nvalid * rowf;
    for (int e = lane; e < total; e += 32) rows[(e / rowf) * pitch + (e % rowf)] = g[e];
  }
}

template <bool BULK>
__device__ __forceinline__ void stage_out(const float* rows, int pitch, float* g, int rowf, int nvalid, int lane) {
  if (BULK) {
    fence_proxy_async();
    __syncwarp();
    if (lane < nvalid) bulk_s2g(g + (size_t)lane * rowf, rows + lane * pitch, rowf * 4);
    bulk_commit();
  } else {
    __syncwarp();
    const int total = nvalid * rowf;
    for (int e = lane; e < total; e += 32) g[e] = rows[(e / rowf) * pitch + (e % rowf)];
  }
}

__host__ __device__ constexpr int pad_pitch(int rowf) {
  // 16-byte aligned pitch whose float4 accesses are bank-conflict free per quarter warp:
  // pitch/4 must be odd  (8 lanes * pitch/4 distinct 16-byte bank groups mod 8)
  int p = (rowf + 3) / 4;
  if (p % 2 == 0) p += 1;
  return p * 4;
}

// ------------------------------------------------------------------------------------------------ FK forward
// smem per warp: in rows [32][pitch_in], (positions rows [32][pitch_p]), out rows [32][pitch_p], (R out rows)
template <class Tree, bool IN6, bool PERFRAME, bool ROUT, bool BULK>
__global__ void __launch_bounds__(32) fk_fwd_kernel(const float* __restrict__ rot, const float* __restrict__ offsets,
                                                    const float* __restrict__ positions, float* __restrict__ pos,
                                                    float* __restrict__ rout, long n, TreeTable tab) {
  extern __shared__ __align__(16) float smem[];
  __shared__ uint64_t bar;
  __shared__ TreeTable stab;
  const int lane = threadIdx.x;
  constexpr int JM = Tree::JMAX;
  Tree tr;
  if constexpr (!Tree::kStatic) {
    if (lane == 0) stab = tab;
    __syncwarp();
    tr.t = &stab;
  }
  const int J = tr.joints();
  const int rin = (IN6 ? 6 : 9) * J, pitch_in = pad_pitch(rin);
  const int rp = 3 * J, pitch_p = pad_pitch(rp);
  const int pitch_r = pad_pitch(9 * J);
  float* s_in = smem;
  float* s_off = s_in + 32 * pitch_in;                       // PERFRAME: [32][pitch_p]; else [rp] shared offsets
  float* s_out = s_off + (PERFRAME ? 32 * pitch_p : ((rp + 3) / 4) * 4);
  float* s_rout = s_out + 32 * pitch_p;

  if (BULK && lane == 0) {
    mbar_init(&bar, 1);
    mbar_fence_init();
  }
  if (!PERFRAME)
    for (int e = lane; e < rp; e += 32) s_off[e] = offsets[e];
  __syncwarp();

  uint32_t parity = 0;
  const long ntiles = (n + 31) / 32;
  for (long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const long f0 = tile * 32;
    const int nvalid = (int)((n - f0) < 32 ? (n - f0) : 32);
    if (BULK) {
      if (lane == 0) mbar_arrive_expect_tx(&bar, (uint32_t)nvalid * (rin + (PERFRAME ? rp : 0)) * 4);
      __syncwarp();
    }
    stage_in<BULK>(s_in, pitch_in, rot + f0 * rin, rin, nvalid, lane, &bar);
    if (PERFRAME) stage_in<BULK>(s_off, pitch_p, positions + f0 * rp, rp, nvalid, lane, &bar);
    if (BULK) {
      mbar_wait(&bar, parity);
      parity ^= 1;
    } else {
      __syncwarp();
    }

    if (lane < nvalid) {
      const float* row = s_in + lane * pitch_in;
      const float* off = PERFRAME ? (s_off + lane * pitch_p) : s_off;
      float* orow = s_out + lane * pitch_p;
      float* rrow = s_rout + lane * pitch_r;
      float Rg[JM][9], pg[JM][3];
      RowWriter<3> pw;
      RowWriter<9> rw;
#pragma unroll
      for (int i = 0; i < JM; ++i) {
        if (i < J) {
          float R[9];
          if (IN6) {
            float a6[6];
            row_load<6>(row, 6 * i, a6);
            rot6d_fwd(a6, R);
            if (ROUT) rw.put(rrow, i, R, true, J);
          } else {
            row_load<9>(row, 9 * i, R);
          }
          float o[3];
          if (PERFRAME) row_load<3>(off, 3 * i, o);
          else { o[0] = off[3 * i]; o[1] = off[3 * i + 1]; o[2] = off[3 * i + 2]; }
          if (i == 0) {
#pragma unroll
            for (int k = 0; k < 9; ++k) Rg[0][k] = R[k];
            pg[0][0] = o[0]; pg[0][1] = o[1]; pg[0][2] = o[2];
          } else {
            const int p = tr.parent(i);
#pragma unroll
            for (int a = 0; a < 3; ++a)
              pg[i][a] = Rg[p][a * 3 + 0] * o[0] + Rg[p][a * 3 + 1] * o[1] + Rg[p][a * 3 + 2] * o[2] + pg[p][a];
            if (!Tree::kStatic || !tr.leaf(i)) mat_mul(Rg[p], R, Rg[i]);
          }
          pw.put(orow, i, pg[i], true, J);
        }
      }
    }
    stage_out<BULK>(s_out, pitch_p, pos + f0 * rp, rp, nvalid, lane);
    if (ROUT) stage_out<BULK>(s_rout, pitch_r, rout + f0 * 9 * J, 9 * J, nvalid, lane);
    if (BULK) bulk_wait_read_all();   // smem rows are reused by the next tile
    __syncwarp();
  }
  if (BULK) bulk_wait_all();
}

// ------------------------------------------------------------------------------------------------ FK backward
// smem per warp: in rows (overwritten with d(in) and stored), dpos rows, (positions rows), Rg slots [nslots*9][32]
template <class Tree, bool IN6, bool PERFRAME, bool BULK>
__global__ void __launch_bounds__(32) fk_bwd_kernel(const float* __restrict__ rot, const float* __restrict__ offsets,
                                                    const float* __restrict__ positions,
                                                    const float* __restrict__ dpos, float* __restrict__ drot, long n,
                                                    TreeTable tab) {
  extern __shared__ __align__(16) float smem[];
  __shared__ uint64_t bar;
  __shared__ TreeTable stab;
  const int lane = threadIdx.x;
  constexpr int JM = Tree::JMAX;
  Tree tr;
  int nslots = Smpl24Tree::kSlots;
  if constexpr (!Tree::kStatic) {
    if (lane == 0) stab = tab;
    __syncwarp();
    tr.t = &stab;
    nslots = stab.nslots;
  }
  const int J = tr.joints();
  constexpr int RD = IN6 ? 6 : 9;
  const int rin = RD * J, pitch_in = pad_pitch(rin);
  const int rp = 3 * J, pitch_p = pad_pitch(rp);
  float* s_in = smem;
  float* s_dp = s_in + 32 * pitch_in;
  float* s_off = s_dp + 32 * pitch_p;
  float* s_rg = s_off + (PERFRAME ? 32 * pitch_p : ((rp + 3) / 4) * 4);   // [(slot*9+k)][32]

  if (BULK && lane == 0) {
    mbar_init(&bar, 1);
    mbar_fence_init();
  }
  if (!PERFRAME)
    for (int e = lane; e < rp; e += 32) s_off[e] = offsets[e];
  __syncwarp();

  uint32_t parity = 0;
  const long ntiles = (n + 31) / 32;
  for (long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const long f0 = tile * 32;
    const int nvalid = (int)((n - f0) < 32 ? (n - f0) : 32);
    if (BULK) {
      if (lane == 0) mbar_arrive_expect_tx(&bar, (uint32_t)nvalid * (rin + rp + (PERFRAME ? rp : 0)) * 4);
      __syncwarp();
    }
    stage_in<BULK>(s_in, pitch_in, rot + f0 * rin, rin, nvalid, lane, &bar);
    stage_in<BULK>(s_dp, pitch_p, dpos + f0 * rp, rp, nvalid, lane, &bar);
    if (PERFRAME) stage_in<BULK>(s_off, pitch_p, positions + f0 * rp, rp, nvalid, lane, &bar);
    if (BULK) {
      mbar_wait(&bar, parity);
      parity ^= 1;
    } else {
      __syncwarp();
    }

    if (lane < nvalid) {
      float* row = s_in + lane * pitch_in;
      const float* drow = s_dp + lane * pitch_p;
      const float* off = PERFRAME ? (s_off + lane * pitch_p) : s_off;
      // ---- pass 1: recompute global rotations, keep the ones the reverse pass needs
      {
        float Rg[JM][9];
#pragma unroll
        for (int i = 0; i < JM; ++i) {
          if (i < J && (!Tree::kStatic || !tr.leaf(i))) {
            float R[9];
            if (IN6) {
              float a6[6];
              row_load<6>(row, 6 * i, a6);
              rot6d_fwd(a6, R);
            } else {
              row_load<9>(row, 9 * i, R);
            }
            if (i == 0) {
#pragma unroll
              for (int k = 0; k < 9; ++k) Rg[0][k] = R[k];
            } else {
              mat_mul(Rg[tr.parent(i)], R, Rg[i]);
            }
            const int s = tr.slot(i);
            if (s >= 0) {
#pragma unroll
              for (int k = 0; k < 9; ++k) s_rg[(s * 9 + k) * 32 + lane] = Rg[i][k];
            }
          }
        }
      }
      // ---- pass 2: reverse accumulation
      float gR[JM][9], gp[JM][3];
#pragma unroll
      for (int i = 0; i < JM; ++i) {
#pragma unroll
        for (int k = 0; k < 9; ++k) gR[i][k] = 0.f;
        gp[i][0] = gp[i][1] = gp[i][2] = 0.f;
      }
      RowWriter<RD> dw;
#pragma unroll
      for (int i = JM - 1; i >= 0; --i) {
        if (i < J) {
          float out[RD];
          float dR[9];
          if (i == 0) {
#pragma unroll
            for (int k = 0; k < 9; ++k) dR[k] = gR[0][k];
          } else {
            const int p = tr.parent(i);
            float g[3], o[3];
            row_load<3>(drow, 3 * i, g);
            if (PERFRAME) row_load<3>(off, 3 * i, o);
            else { o[0] = off[3 * i]; o[1] = off[3 * i + 1]; o[2] = off[3 * i + 2]; }
#pragma unroll
            for (int a = 0; a < 3; ++a) {
              g[a] += gp[i][a];
              gp[p][a] += g[a];
#pragma unroll
              for (int b = 0; b < 3; ++b) gR[p][a * 3 + b] += g[a] * o[b];
            }
            const bool leaf = tr.leaf(i);
            if (!leaf) {
              float R[9];
              if (IN6) {
                float a6[6];
                row_load<6>(row, 6 * i, a6);
                rot6d_fwd(a6, R);
              } else {
                row_load<9>(row, 9 * i, R);
              }
              // gR[p] += gR[i] * R^T
#pragma unroll
              for (int a = 0; a < 3; ++a)
#pragma unroll
                for (int b = 0; b < 3; ++b)
                  gR[p][a * 3 + b] += gR[i][a * 3 + 0] * R[b * 3 + 0] + gR[i][a * 3 + 1] * R[b * 3 + 1] + gR[i][a * 3 + 2] * R[b * 3 + 2];
              // dR_i = Rg_p^T * gR[i]
              const int s = tr.slot(p);
              float P[9];
#pragma unroll
              for (int k = 0; k < 9; ++k) P[k] = s_rg[(s * 9 + k) * 32 + lane];
#pragma unroll
              for (int a = 0; a < 3; ++a)
#pragma unroll
                for (int b = 0; b < 3; ++b)
                  dR[a * 3 + b] = P[0 * 3 + a] * gR[i][0 * 3 + b] + P[1 * 3 + a] * gR[i][1 * 3 + b] + P[2 * 3 + a] * gR[i][2 * 3 + b];
            } else {
#pragma unroll
              for (int k = 0; k < 9; ++k) dR[k] = 0.f;
            }
          }
          if (IN6) {
            float a6[6];
            row_load<6>(row, 6 * i, a6);
            rot6d_bwd(a6, dR, out);
          } else {
#pragma unroll
            for (int k = 0; k < 9; ++k) out[k] = dR[k];
          }
          dw.put(row, i, out, false, J);   // in place: joints >= 4*(i/4) are fully consumed when the group flushes
        }
      }
    }
    stage_out<BULK>(s_in, pitch_in, drot + f0 * rin, rin, nvalid, lane);
    if (BULK) bulk_wait_read_all();
    __syncwarp();
  }
  if (BULK) bulk_wait_all();
}

// ------------------------------------------------------------------------------------------------ FK backward, row-split
// The lane-per-frame kernel above is a ~3 400-instruction dependent chain per 32-frame tile with 54 KB of staging per warp:
// four warps per SM, each issuing every ~7 cycles (ncu, profiles/r01_fk_details.txt) => 35 % of the HBM roofline.  This variant
// splits every frame over THREE lanes.  All quantities of the chain decompose by the row r of the global rotation:
//     Rg_i[r,:] = Rg_p[r,:] R_i                        (forward recomputation)
//     S_i[r]    = dpos_i[r] + sum_children S_c[r]       (accumulated position gradient)
//     H_p[r,:] += S_i[r] off_i^T + H_i[r,:] R_i^T      (gradient w.r.t. the global rotation of the parent)
// so lane (f, r) walks the whole tree for frame f with one third of the arithmetic and one third of the live registers; only
//     dR_i = Rg_p^T H_i                                 (row t of dR_i = sum_u Rg_p[u,t] H_i[u,:])
// mixes rows: the three lanes of a frame exchange H_i (6 shuffles) and one column element of Rg_p each (2 shuffles) and every
// lane writes one row of dR_i in place of R_i.  A warp owns 10 consecutive frames (lanes 30, 31 shadow frame 9), i.e. 11.8 KB of
// staging, so 16 warps per SM are resident and the bulk copies of one warp hide under the arithmetic of the others.
constexpr int FKR_FRAMES = 10;
constexpr int FKR_WARPS = 4;

template <class Tree>
__global__ void __launch_bounds__(32 * FKR_WARPS, 4) fk_bwd_rows_kernel(const float* __restrict__ rot, const float* __restrict__ offsets,
                                                                         const float* __restrict__ dpos, float* __restrict__ drot,
                                                                         long n, TreeTable tab) {
  extern __shared__ __align__(16) float smem[];
  __shared__ uint64_t bars[FKR_WARPS];
  __shared__ TreeTable stab;
  __shared__ float s_off[3 * FK_MAX_J];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int JM = Tree::JMAX;
  Tree tr;
  if constexpr (!Tree::kStatic) {
    if (threadIdx.x == 0) stab = tab;
    tr.t = &stab;
  }
  if (lane == 0) {
    mbar_init(&bars[warp], 1);
    mbar_fence_init();
  }
  for (int e = threadIdx.x; e < 3 * tab.J; e += blockDim.x) s_off[e] = offsets[e];
  __syncthreads();
  const int J = tr.joints();
  const int rin = 9 * J, pitch_in = pad_pitch(rin);
  const int rp = 3 * J, pitch_p = pad_pitch(rp);
  float* s_in = smem + (size_t)warp * FKR_FRAMES * (pitch_in + pitch_p);
  float* s_dp = s_in + FKR_FRAMES * pitch_in;
  uint64_t* bar = &bars[warp];

  const int f = lane < 3 * FKR_FRAMES ? lane / 3 : FKR_FRAMES - 1;      // frame inside the tile
  const int r = lane < 3 * FKR_FRAMES ? lane - 3 * f : lane - 3 * FKR_FRAMES;   // row of the global rotation handled by this lane
  const int base = 3 * f;
  const int src1 = base + (r + 1) % 3, src2 = base + (r + 2) % 3;       // the other two lanes of the frame
  uint32_t parity = 0;
  const long ntiles = (n + FKR_FRAMES - 1) / FKR_FRAMES;
  for (long tile = (long)blockIdx.x * FKR_WARPS + warp; tile < ntiles; tile += (long)gridDim.x * FKR_WARPS) {
    const long f0 = tile * FKR_FRAMES;
    const int nvalid = (int)((n - f0) < FKR_FRAMES ? (n - f0) : FKR_FRAMES);
    if (lane == 0) mbar_arrive_expect_tx(bar, (uint32_t)nvalid * (rin + rp) * 4);
    __syncwarp();
    if (lane < nvalid) {
      bulk_g2s(s_in + lane * pitch_in, rot + (f0 + lane) * rin, rin * 4, bar);
      bulk_g2s(s_dp + lane * pitch_p, dpos + (f0 + lane) * rp, rp * 4, bar);
    }
    mbar_wait(bar, parity);
    parity ^= 1;

    float* row = s_in + f * pitch_in;
    const float* drow = s_dp + f * pitch_p;
    const bool writer = lane < 3 * FKR_FRAMES && f < nvalid;
    // ---- pass 1: row r of every global rotation that has children
    float Rg[JM][3];
#pragma unroll
    for (int i = 0; i < JM; ++i) {
      if (i < J && (!Tree::kStatic || !tr.leaf(i))) {
        float R[9];
        row_load<9>(row, 9 * i, R);
        if (i == 0) {
          Rg[0][0] = r == 0 ? R[0] : (r == 1 ? R[3] : R[6]);
          Rg[0][1] = r == 0 ? R[1] : (r == 1 ? R[4] : R[7]);
          Rg[0][2] = r == 0 ? R[2] : (r == 1 ? R[5] : R[8]);
        } else {
          const int p = tr.parent(i);
#pragma unroll
          for (int b = 0; b < 3; ++b) Rg[i][b] = Rg[p][0] * R[b] + Rg[p][1] * R[3 + b] + Rg[p][2] * R[6 + b];
        }
      }
    }
    // ---- pass 2: reverse accumulation (row r of H, component r of S)
    float H[JM][3], S[JM];
#pragma unroll
    for (int i = 0; i < JM; ++i) {
      H[i][0] = H[i][1] = H[i][2] = 0.f;
      S[i] = 0.f;
    }
#pragma unroll
    for (int i = JM - 1; i >= 0; --i) {
      if (i < J) {
        float d[3] = {0.f, 0.f, 0.f};
        if (i == 0) {
          d[0] = H[0][0]; d[1] = H[0][1]; d[2] = H[0][2];
        } else {
          const int p = tr.parent(i);
          const float g = drow[3 * i + r] + S[i];
          S[p] += g;
#pragma unroll
          for (int b = 0; b < 3; ++b) H[p][b] += g * s_off[3 * i + b];
          if (!tr.leaf(i)) {
            float R[9];
            row_load<9>(row, 9 * i, R);
#pragma unroll
            for (int b = 0; b < 3; ++b) H[p][b] += H[i][0] * R[3 * b] + H[i][1] * R[3 * b + 1] + H[i][2] * R[3 * b + 2];
            // row r of dR_i = sum_u Rg_p[u][r] * H_i[u][:]
            // what the OTHER lanes need from me: lane t = (r + 2) % 3 reads me as its src1 and wants Rg_p[r][t]; lane (r + 1) % 3
            // reads me as its src2 and wants Rg_p[r][(r + 1) % 3]
            const float own = r == 0 ? Rg[p][0] : (r == 1 ? Rg[p][1] : Rg[p][2]);
            const float for_src1_reader = r == 0 ? Rg[p][2] : (r == 1 ? Rg[p][0] : Rg[p][1]);      // element (r + 2) % 3
            const float for_src2_reader = r == 0 ? Rg[p][1] : (r == 1 ? Rg[p][2] : Rg[p][0]);      // element (r + 1) % 3
            const float c1 = __shfl_sync(0xffffffffu, for_src1_reader, src1);
            const float c2 = __shfl_sync(0xffffffffu, for_src2_reader, src2);
#pragma unroll
            for (int b = 0; b < 3; ++b) {
              const float h1 = __shfl_sync(0xffffffffu, H[i][b], src1);
              const float h2 = __shfl_sync(0xffffffffu, H[i][b], src2);
              d[b] = own * H[i][b] + c1 * h1 + c2 * h2;
            }
          }
        }
        __syncwarp();      // every lane of the frame has read R_i: its slot now receives dR_i
        if (writer) {
          row[9 * i + 3 * r] = d[0];
          row[9 * i + 3 * r + 1] = d[1];
          row[9 * i + 3 * r + 2] = d[2];
        }
      }
    }
    fence_proxy_async();
    __syncwarp();
    if (lane < nvalid) bulk_s2g(drot + (f0 + lane) * rin, s_in + lane * pitch_in, rin * 4);
    bulk_commit();
    bulk_wait_read_all();      // the rows are reused by the next tile
    __syncwarp();
  }
  bulk_wait_all();
}

// ------------------------------------------------------------------------------------------------ rot6d / aa
constexpr int ROT_TPB = 256;

__global__ void __launch_bounds__(ROT_TPB) rot6d_fwd_kernel(const float* __restrict__ x6, float* __restrict__ R, long m) {
  __shared__ __align__(16) float s_in[ROT_TPB * 6];
  __shared__ __align__(16) float s_out[ROT_TPB * 9 + 4];
  for (long base = (long)blockIdx.x * ROT_TPB; base < m; base += (long)gridDim.x * ROT_TPB) {
    const int cnt = (int)((m - base) < ROT_TPB ? (m - base) : ROT_TPB);
    const float* g = x6 + base * 6;
    if (cnt == ROT_TPB) {
      const float4* g4 = reinterpret_cast<const float4*>(g);
      for (int e = threadIdx.x; e < ROT_TPB * 6 / 4; e += ROT_TPB) reinterpret_cast<float4*>(s_in)[e] = g4[e];
    } else {
      for (int e = threadIdx.x; e < cnt * 6; e += ROT_TPB) s_in[e] = g[e];
    }
    __syncthreads();
    if (threadIdx.x < cnt) {
      float a6[6], Rm[9];
#pragma unroll
      for (int k = 0; k < 6; ++k) a6[k] = s_in[threadIdx.x * 6 + k];
      rot6d_fwd(a6, Rm);
#pragma unroll
      for (int k = 0; k < 9; ++k) s_out[threadIdx.x * 9 + k] = Rm[k];   // stride 9: conflict-free
    }
    __syncthreads();
    float* o = R + base * 9;
    if (cnt == ROT_TPB) {
      float4* o4 = reinterpret_cast<float4*>(o);
      for (int e = threadIdx.x; e < ROT_TPB * 9 / 4; e += ROT_TPB) o4[e] = reinterpret_cast<const float4*>(s_out)[e];
    } else {
      for (int e = threadIdx.x; e < cnt * 9; e += ROT_TPB) o[e] = s_out[e];
    }
    __syncthreads();
  }
}

// Warp-tiled, no block barriers: one warp owns 128 consecutive matrices per iteration.  Every lane first issues its 15
// float4 loads (6 of x6, 9 of dR: 240 bytes in flight per lane), parks them in the warp's shared tile, computes 4 matrices
// (matrix q*32 + lane: stride-9 / stride-6 shared reads) and the 6 float4 of dx6 per lane leave coalesced.  The block-wide
// load -> sync -> compute -> sync -> store version reached 53 % of the HBM roofline; its phases could not overlap.
constexpr int ROTB_WARPS = 4;
constexpr int ROTB_TILE = 128;

__global__ void __launch_bounds__(32 * ROTB_WARPS) rot6d_bwd_kernel(const float* __restrict__ x6, const float* __restrict__ dR,
                                                                    float* __restrict__ dx6, long m) {
  __shared__ __align__(16) float s_x[ROTB_WARPS][ROTB_TILE * 6];
  __shared__ __align__(16) float s_g[ROTB_WARPS][ROTB_TILE * 9];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* sx = s_x[w];
  float* sg = s_g[w];
  const long ntiles = (m + ROTB_TILE - 1) / ROTB_TILE;
  const long wid = (long)blockIdx.x * ROTB_WARPS + w, nw = (long)gridDim.x * ROTB_WARPS;
  for (long tile = wid; tile < ntiles; tile += nw) {
    const long base = tile * ROTB_TILE;
    const int cnt = (int)((m - base) < ROTB_TILE ? (m - base) : ROTB_TILE);
    const float* gx = x6 + base * 6;
    const float* gg = dR + base * 9;
    if (cnt == ROTB_TILE) {
      float4 vx[6], vg[9];
#pragma unroll
      for (int q = 0; q < 6; ++q) vx[q] = __ldcs(reinterpret_cast<const float4*>(gx) + q * 32 + lane);
#pragma unroll
      for (int q = 0; q < 9; ++q) vg[q] = __ldcs(reinterpret_cast<const float4*>(gg) + q * 32 + lane);
#pragma unroll
      for (int q = 0; q < 6; ++q) reinterpret_cast<float4*>(sx)[q * 32 + lane] = vx[q];
#pragma unroll
      for (int q = 0; q < 9; ++q) reinterpret_cast<float4*>(sg)[q * 32 + lane] = vg[q];
    } else {
      for (int e = lane; e < cnt * 6; e += 32) sx[e] = gx[e];
      for (int e = lane; e < cnt * 9; e += 32) sg[e] = gg[e];
    }
    __syncwarp();
    float g6[4][6];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int idx = q * 32 + lane;
      if (idx < cnt) {
        float a6[6], G[9];
#pragma unroll
        for (int k = 0; k < 6; ++k) a6[k] = sx[idx * 6 + k];
#pragma unroll
        for (int k = 0; k < 9; ++k) G[k] = sg[idx * 9 + k];
        rot6d_bwd(a6, G, g6[q]);
      }
    }
    __syncwarp();
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int idx = q * 32 + lane;
      if (idx < cnt) {
#pragma unroll
        for (int k = 0; k < 6; ++k) sx[idx * 6 + k] = g6[q][k];
      }
    }
    __syncwarp();
    float* o = dx6 + base * 6;
    if (cnt == ROTB_TILE) {
#pragma unroll
      for (int q = 0; q < 6; ++q) __stcs(reinterpret_cast<float4*>(o) + q * 32 + lane, reinterpret_cast<const float4*>(sx)[q * 32 + lane]);
    } else {
      for (int e = lane; e < cnt * 6; e += 32) o[e] = sx[e];
    }
    __syncwarp();
  }
}

// torchgeometry semantics (see oracle/hmvae_ref.py::angle_axis_to_rotation_matrix)
__global__ void aa2rot_kernel(const float* __restrict__ aa, float* __restrict__ out, long m) {
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += (long)gridDim.x * blockDim.x) {
    const float rx = aa[i * 3], ry = aa[i * 3 + 1], rz = aa[i * 3 + 2];
    const float t2 = rx * rx + ry * ry + rz * rz;
    float r[9];
    if (t2 > 1e-6f) {
      const float th = sqrtf(t2);
      const float inv = 1.f / (th + 1e-6f);
      const float wx = rx * inv, wy = ry * inv, wz = rz * inv;
      float s, c;
      sincosf(th, &s, &c);
      const float k = 1.f - c;
      r[0] = c + wx * wx * k;       r[1] = wx * wy * k - wz * s;  r[2] = wy * s + wx * wz * k;
      r[3] = wz * s + wx * wy * k;  r[4] = c + wy * wy * k;       r[5] = -wx * s + wy * wz * k;
      r[6] = -wy * s + wx * wz * k; r[7] = wx * s + wy * wz * k;  r[8] = c + wz * wz * k;
    } else {
      r[0] = 1.f; r[1] = -rz; r[2] = ry;
      r[3] = rz;  r[4] = 1.f; r[5] = -rx;
      r[6] = -ry; r[7] = rx;  r[8] = 1.f;
    }
    float4* o = reinterpret_cast<float4*>(out + i * 16);
    o[0] = make_float4(r[0], r[1], r[2], 0.f);
    o[1] = make_float4(r[3], r[4], r[5], 0.f);
    o[2] = make_float4(r[6], r[7], r[8], 0.f);
    o[3] = make_float4(0.f, 0.f, 0.f, 1.f);
  }
}

// ------------------------------------------------------------------------------------------------ host side
static int build_table(const int* parents, int J, TreeTable* t, bool* is_smpl) {
  if (J < 1 || J > FK_MAX_J) return fail_arg("fk: joints must be in [1, 32]");
  t->J = J;
  bool smpl = (J == 24);
  for (int i = 0; i < FK_MAX_J; ++i) { t->parent[i] = 0; t->slot[i] = -1; t->leaf[i] = 1; }
  for (int i = 1; i < J; ++i) {
    if (parents[i] < 0 || parents[i] >= i) return fail_arg("fk: parents[i] must satisfy 0 <= parents[i] < i");
    t->parent[i] = (signed char)parents[i];
    t->leaf[parents[i]] = 0;
    if (smpl && parents[i] != Smpl24Tree::parent_of(i)) smpl = false;
  }
  if (J == 1) t->leaf[0] = 0;
  int s = 0;
  for (int q = 0; q < J; ++q) {
    bool need = false;
    for (int c = 1; c < J; ++c)
      if (t->parent[c] == q && !t->leaf[c]) need = true;
    if (need) t->slot[q] = (signed char)s++;
  }
  t->nslots = s;
  *is_smpl = smpl;
  return 0;
}

template <class K>
static int set_smem(K kernel, size_t bytes) {
  HMVAE_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  return 0;
}

static int fk_grid(long n, size_t smem_bytes) {
  long tiles = (n + 31) / 32;
  int per_sm = (int)((227 * 1024) / (smem_bytes + 1024));
  if (per_sm < 1) per_sm = 1;
  if (per_sm > 16) per_sm = 16;
  long cap = (long)num_sms() * per_sm;
  return (int)(tiles < cap ? tiles : cap);
}

template <class Tree, bool IN6, bool PERFRAME, bool ROUT, bool BULK>
static int launch_fk_fwd(const float* rot, const float* offsets, const float* positions, float* pos, float* rout,
                         long n, const TreeTable& tab, cudaStream_t st) {
  const int J = tab.J;
  size_t fl = 32 * (size_t)pad_pitch((IN6 ? 6 : 9) * J) + (PERFRAME ? 32 * pad_pitch(3 * J) : ((3 * J + 3) / 4) * 4) +
              32 * (size_t)pad_pitch(3 * J) + (ROUT ? 32 * (size_t)pad_pitch(9 * J) : 0);
  size_t bytes = fl * 4;
  auto k = fk_fwd_kernel<Tree, IN6, PERFRAME, ROUT, BULK>;
  int rc = set_smem(k, bytes);
  if (rc) return rc;
  k<<<fk_grid(n, bytes), 32, bytes, st>>>(rot, offsets, positions, pos, rout, n, tab);
  return check_launch("fk_fwd");
}

template <class Tree, bool IN6, bool PERFRAME, bool BULK>
static int launch_fk_bwd(const float* rot, const float* offsets, const float* positions, const float* dpos,
                         float* drot, long n, const TreeTable& tab, cudaStream_t st) {
  const int J = tab.J;
  size_t fl = 32 * (size_t)pad_pitch((IN6 ? 6 : 9) * J) + 32 * (size_t)pad_pitch(3 * J) +
              (PERFRAME ? 32 * pad_pitch(3 * J) : ((3 * J + 3) / 4) * 4) + (size_t)tab.nslots * 9 * 32;
  size_t bytes = fl * 4;
  auto k = fk_bwd_kernel<Tree, IN6, PERFRAME, BULK>;
  int rc = set_smem(k, bytes);
  if (rc) return rc;
  k<<<fk_grid(n, bytes), 32, bytes, st>>>(rot, offsets, positions, dpos, drot, n, tab);
  return check_launch("fk_bwd");
}

template <class Tree>
static int launch_fk_bwd_rows(const float* rot, const float* offsets, const float* dpos, float* drot, long n, const TreeTable& tab,
                              cudaStream_t st) {
  const int J = tab.J;
  const size_t bytes = (size_t)FKR_WARPS * FKR_FRAMES * (pad_pitch(9 * J) + pad_pitch(3 * J)) * 4;
  auto k = fk_bwd_rows_kernel<Tree>;
  int rc = set_smem(k, bytes);
  if (rc) return rc;
  const long tiles = (n + FKR_FRAMES - 1) / FKR_FRAMES;
  long blocks = (tiles + FKR_WARPS - 1) / FKR_WARPS;
  const long cap = (long)num_sms() * 4;
  if (blocks > cap) blocks = cap;
  k<<<(int)blocks, 32 * FKR_WARPS, bytes, st>>>(rot, offsets, dpos, drot, n, tab);
  return check_launch("fk_bwd");
}

}  // namespace hmvae

using namespace hmvae;

extern "C" int hmvae_fk_fwd(const float* rot, int rot_dim, const float* offsets, const float* positions,
                            const int* parents, int joints, long n, float* pos, float* rotmat_out, void* stream) {
  if (n <= 0) return 0;
  if (!rot || !pos || !parents || (!offsets && !positions)) return fail_arg("fk_fwd: null pointer");
  if (rot_dim != 9 && rot_dim != 6) return fail_arg("fk_fwd: rot_dim must be 9 (3x3) or 6");
  if (rotmat_out && rot_dim != 6) return fail_arg("fk_fwd: rotmat_out only with 6D input");
  if (n <= 0) return 0;
  TreeTable tab;
  bool smpl;
  int rc = build_table(parents, joints, &tab, &smpl);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  const bool al = aligned16(rot) && aligned16(pos) && (!positions || aligned16(positions)) && (!rotmat_out || aligned16(rotmat_out));
  const bool in6 = rot_dim == 6, pf = positions != nullptr, ro = rotmat_out != nullptr;
  if (smpl && al) {
#define FWD_S(I6, PF, RO) if (in6 == I6 && pf == PF && ro == RO) return launch_fk_fwd<Smpl24Tree, I6, PF, RO, true>(rot, offsets, positions, pos, rotmat_out, n, tab, st);
    FWD_S(false, false, false) FWD_S(false, true, false) FWD_S(true, false, false) FWD_S(true, false, true)
    FWD_S(true, true, false) FWD_S(true, true, true)
#undef FWD_S
  }
#define FWD_G(I6, PF, RO) if (in6 == I6 && pf == PF && ro == RO) return launch_fk_fwd<RuntimeTree, I6, PF, RO, false>(rot, offsets, positions, pos, rotmat_out, n, tab, st);
  FWD_G(false, false, false) FWD_G(false, true, false) FWD_G(true, false, false) FWD_G(true, false, true)
  FWD_G(true, true, false) FWD_G(true, true, true)
#undef FWD_G
  return fail_arg("fk_fwd: unsupported combination");
}

extern "C" int hmvae_fk_bwd(const float* rot, int rot_dim, const float* offsets, const float* positions,
                            const int* parents, int joints, long n, const float* dpos, const float* drotmat_extra,
                            float* drot, void* stream) {
  if (n <= 0) return 0;
  if (!rot || !dpos || !drot || !parents || (!offsets && !positions)) return fail_arg("fk_bwd: null pointer");
  if (rot_dim != 9 && rot_dim != 6) return fail_arg("fk_bwd: rot_dim must be 9 (3x3) or 6");
  if (drotmat_extra) return fail_arg("fk_bwd: drotmat_extra is not supported (add it to drot on the caller side)");
  if (n <= 0) return 0;
  TreeTable tab;
  bool smpl;
  int rc = build_table(parents, joints, &tab, &smpl);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  const bool al = aligned16(rot) && aligned16(dpos) && aligned16(drot) && (!positions || aligned16(positions));
  const bool in6 = rot_dim == 6, pf = positions != nullptr;
  // rotation-matrix input with the layer's own offsets (the training / benchmark case): three lanes per frame
  // (bulk copies of whole frame rows: 9*J*4 and 3*J*4 bytes must be multiples of 16)
  if (!in6 && !pf && al && joints % 4 == 0 && env_int("HMVAE_FK_BWD_ROWS", 1)) {
    if (smpl) return launch_fk_bwd_rows<Smpl24Tree>(rot, offsets, dpos, drot, n, tab, st);
    return launch_fk_bwd_rows<RuntimeTree>(rot, offsets, dpos, drot, n, tab, st);
  }
  if (smpl && al) {
#define BWD_S(I6, PF) if (in6 == I6 && pf == PF) return launch_fk_bwd<Smpl24Tree, I6, PF, true>(rot, offsets, positions, dpos, drot, n, tab, st);
    BWD_S(false, false) BWD_S(false, true) BWD_S(true, false) BWD_S(true, true)
#undef BWD_S
  }
#define BWD_G(I6, PF) if (in6 == I6 && pf == PF) return launch_fk_bwd<RuntimeTree, I6, PF, false>(rot, offsets, positions, dpos, drot, n, tab, st);
  BWD_G(false, false) BWD_G(false, true) BWD_G(true, false) BWD_G(true, true)
#undef BWD_G
  return fail_arg("fk_bwd: unsupported combination");
}

extern "C" int hmvae_rot6d_fwd(const float* x6, float* rotmat, long m, void* stream) {
  if (m <= 0) return 0;
  if (!x6 || !rotmat) return fail_arg("rot6d_fwd: null pointer");
  if (!aligned16(x6) || !aligned16(rotmat)) return fail_arg("rot6d_fwd: pointers must be 16-byte aligned");
  if (m <= 0) return 0;
  long blocks = (m + ROT_TPB - 1) / ROT_TPB;
  long cap = (long)num_sms() * 8;
  rot6d_fwd_kernel<<<(int)(blocks < cap ? blocks : cap), ROT_TPB, 0, (cudaStream_t)stream>>>(x6, rotmat, m);
  return check_launch("rot6d_fwd");
}

extern "C" int hmvae_rot6d_bwd(const float* x6, const float* drotmat, float* dx6, long m, void* stream) {
  if (m <= 0) return 0;
  if (!x6 || !drotmat || !dx6) return fail_arg("rot6d_bwd: null pointer");
  if (!aligned16(x6) || !aligned16(drotmat) || !aligned16(dx6)) return fail_arg("rot6d_bwd: pointers must be 16-byte aligned");
  if (m <= 0) return 0;
  long blocks = (m + ROTB_TILE * ROTB_WARPS - 1) / (ROTB_TILE * ROTB_WARPS);
  long cap = (long)num_sms() * 7;
  rot6d_bwd_kernel<<<(int)(blocks < cap ? blocks : cap), 32 * ROTB_WARPS, 0, (cudaStream_t)stream>>>(x6, drotmat, dx6, m);
  return check_launch("rot6d_bwd");
}

extern "C" int hmvae_aa2rot_fwd(const float* aa, float* out44, long m, void* stream) {
  if (m <= 0) return 0;
  if (!aa || !out44) return fail_arg("aa2rot_fwd: null pointer");
  if (!aligned16(out44)) return fail_arg("aa2rot_fwd: output must be 16-byte aligned");
  if (m <= 0) return 0;
  long blocks = (m + 255) / 256;
  long cap = (long)num_sms() * 8;
  aa2rot_kernel<<<(int)(blocks < cap ? blocks : cap), 256, 0, (cudaStream_t)stream>>>(aa, out44, m);
  return check_launch("aa2rot_fwd");
}
