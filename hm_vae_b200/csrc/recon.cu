// Fused reconstruction loss: GT FK, rot6d->R, FK on the prediction, the three MSE sums AND d(total)/d(x6_pred)
// in one kernel.  Replaces seq_two_hier_sa_vae.py:343 (FK on GT), :441-468 (transpose, rot6d->R, FK),
// :395-397 + :430-434 (3x MSE) and the autograd replay of all of them (~400 launches in the reference).
// R_pred, pos_pred, pos_gt and every intermediate gradient stay on chip.
//
// One warp per 32 consecutive frames, one lane per frame.  The decoder output is read directly in its NCW layout
// (lane = time step => coalesced), the BTC ground-truth rows are staged with bulk (TMA) copies.
// Algorithmic bytes per frame (J=24): 144*4 (x6) + 144*4 (gt 6d) + 216*4 (gt R) in, 144*4 (dx6) out = 2592 B.
#include "fk_device.cuh"

namespace hmvae {

__host__ __device__ constexpr int rc_pitch(int rowf) {
  int p = (rowf + 3) / 4;
  if (p % 2 == 0) p += 1;
  return p * 4;
}

template <class Tree, bool NCW, bool BULK>
__global__ void __launch_bounds__(32) recon_kernel(const float* __restrict__ x6p, const float* __restrict__ gt6,
                                                   const float* __restrict__ gtR, const float* __restrict__ offsets, int B,
                                                   int T, float s6, float srot, float spos, float* __restrict__ losses,
                                                   float* __restrict__ dx6, float* __restrict__ pos_out,
                                                   float* __restrict__ gtpos_out, TreeTable tab) {
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
  const int r6 = 6 * J, r9 = 9 * J, r3 = 3 * J;
  const int p6 = rc_pitch(r6), p9 = rc_pitch(r9);
  float* s_g6 = smem;                       // [32][p6]   gt 6d rows
  float* s_gR = s_g6 + 32 * p6;             // [32][p9]   gt rotmat rows
  float* s_x6 = s_gR + 32 * p9;             // NCW: [r6][32] cache ; BTC: [32][p6] rows (overwritten with dx6)
  float* s_dp = s_x6 + (NCW ? r6 * 32 : 32 * p6);   // [r3][32]  gt pos, then spos*(pred-gt)
  float* s_rg = s_dp + r3 * 32;             // [nslots*9][32]
  float* s_off = s_rg + nslots * 9 * 32;    // [r3]

  if (BULK && lane == 0) {
    mbar_init(&bar, 1);
    mbar_fence_init();
  }
  for (int e = lane; e < r3; e += 32) s_off[e] = offsets[e];
  __syncwarp();

  const long n = (long)B * T;
  const long ntiles = (n + 31) / 32;
  uint32_t parity = 0;
  float l6 = 0.f, lrot = 0.f, lpos = 0.f;
  for (long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const long f0 = tile * 32;
    const int nvalid = (int)((n - f0) < 32 ? (n - f0) : 32);
    const long f = f0 + lane;
    const bool valid = lane < nvalid;
    if (BULK) {
      if (lane == 0) mbar_arrive_expect_tx(&bar, (uint32_t)nvalid * (r6 + r9 + (NCW ? 0 : r6)) * 4);
      __syncwarp();
      if (valid) {
        bulk_g2s(s_g6 + lane * p6, gt6 + f * r6, r6 * 4, &bar);
        bulk_g2s(s_gR + lane * p9, gtR + f * r9, r9 * 4, &bar);
        if (!NCW) bulk_g2s(s_x6 + lane * p6, x6p + f * r6, r6 * 4, &bar);
      }
    } else {
      for (int e = lane; e < nvalid * r6; e += 32) s_g6[(e / r6) * p6 + e % r6] = gt6[f0 * r6 + e];
      for (int e = lane; e < nvalid * r9; e += 32) s_gR[(e / r9) * p9 + e % r9] = gtR[f0 * r9 + e];
      if (!NCW)
        for (int e = lane; e < nvalid * r6; e += 32) s_x6[(e / r6) * p6 + e % r6] = x6p[f0 * r6 + e];
    }
    // NCW prediction: lane-coalesced loads straight into the [e][lane] cache while the bulk copies fly
    const long bq = valid ? f / T : 0;
    const int tq = valid ? (int)(f % T) : 0;
    const float* xbase = x6p + (bq * r6) * (long)T + tq;
    if (NCW && valid) {
      for (int e = 0; e < r6; ++e) s_x6[e * 32 + lane] = xbase[(long)e * T];
    }
    if (BULK) {
      mbar_wait(&bar, parity);
      parity ^= 1;
    } else {
      __syncwarp();
    }

    if (valid) {
      const float* row6 = s_g6 + lane * p6;
      const float* rowR = s_gR + lane * p9;
      float* rowx = s_x6 + lane * p6;   // BTC only
      // ---------------- pass A1: FK on the ground truth -> s_dp
      {
        float Rg[JM][9], pg[JM][3];
#pragma unroll
        for (int i = 0; i < JM; ++i) {
          if (i < J) {
            float R[9];
            row_load<9>(rowR, 9 * i, R);
            const float o0 = s_off[3 * i], o1 = s_off[3 * i + 1], o2 = s_off[3 * i + 2];
            if (i == 0) {
#pragma unroll
              for (int k = 0; k < 9; ++k) Rg[0][k] = R[k];
              pg[0][0] = o0; pg[0][1] = o1; pg[0][2] = o2;
            } else {
              const int p = tr.parent(i);
#pragma unroll
              for (int a = 0; a < 3; ++a)
                pg[i][a] = Rg[p][a * 3] * o0 + Rg[p][a * 3 + 1] * o1 + Rg[p][a * 3 + 2] * o2 + pg[p][a];
              if (!Tree::kStatic || !tr.leaf(i)) mat_mul(Rg[p], R, Rg[i]);
            }
#pragma unroll
            for (int a = 0; a < 3; ++a) s_dp[(3 * i + a) * 32 + lane] = pg[i][a];
            if (gtpos_out) {
#pragma unroll
              for (int a = 0; a < 3; ++a) gtpos_out[f * r3 + 3 * i + a] = pg[i][a];
            }
          }
        }
      }
      // ---------------- pass A2: prediction chain, the three squared-error sums, dpos
      {
        float Rg[JM][9], pg[JM][3];
#pragma unroll
        for (int i = 0; i < JM; ++i) {
          if (i < J) {
            float a6[6], g6v[6], R[9], Rt[9];
            if (NCW) {
#pragma unroll
              for (int k = 0; k < 6; ++k) a6[k] = s_x6[(6 * i + k) * 32 + lane];
            } else {
              row_load<6>(rowx, 6 * i, a6);
            }
            row_load<6>(row6, 6 * i, g6v);
            row_load<9>(rowR, 9 * i, Rt);
            rot6d_fwd(a6, R);
#pragma unroll
            for (int k = 0; k < 6; ++k) { const float d = a6[k] - g6v[k]; l6 += d * d; }
#pragma unroll
            for (int k = 0; k < 9; ++k) { const float d = R[k] - Rt[k]; lrot += d * d; }
            const float o0 = s_off[3 * i], o1 = s_off[3 * i + 1], o2 = s_off[3 * i + 2];
            if (i == 0) {
#pragma unroll
              for (int k = 0; k < 9; ++k) Rg[0][k] = R[k];
              pg[0][0] = o0; pg[0][1] = o1; pg[0][2] = o2;
            } else {
              const int p = tr.parent(i);
#pragma unroll
              for (int a = 0; a < 3; ++a)
                pg[i][a] = Rg[p][a * 3] * o0 + Rg[p][a * 3 + 1] * o1 + Rg[p][a * 3 + 2] * o2 + pg[p][a];
              if (!Tree::kStatic || !tr.leaf(i)) mat_mul(Rg[p], R, Rg[i]);
            }
            const int s = tr.slot(i);
            if (s >= 0) {
#pragma unroll
              for (int k = 0; k < 9; ++k) s_rg[(s * 9 + k) * 32 + lane] = Rg[i][k];
            }
#pragma unroll
            for (int a = 0; a < 3; ++a) {
              const float d = pg[i][a] - s_dp[(3 * i + a) * 32 + lane];
              lpos += d * d;
              s_dp[(3 * i + a) * 32 + lane] = spos * d;
            }
            if (pos_out) {
#pragma unroll
              for (int a = 0; a < 3; ++a) pos_out[f * r3 + 3 * i + a] = pg[i][a];
            }
          }
        }
      }
      // ---------------- pass B: reverse accumulation -> dx6
      if (dx6) {
        float gR[JM][9], gp[JM][3];
#pragma unroll
        for (int i = 0; i < JM; ++i) {
#pragma unroll
          for (int k = 0; k < 9; ++k) gR[i][k] = 0.f;
          gp[i][0] = gp[i][1] = gp[i][2] = 0.f;
        }
        RowWriter<6> dw;
        float* obase = dx6 + (bq * r6) * (long)T + tq;
#pragma unroll
        for (int i = JM - 1; i >= 0; --i) {
          if (i < J) {
            float a6[6], g6v[6], R[9], Rt[9], dR[9];
            if (NCW) {
#pragma unroll
              for (int k = 0; k < 6; ++k) a6[k] = s_x6[(6 * i + k) * 32 + lane];
            } else {
              row_load<6>(rowx, 6 * i, a6);
            }
            row_load<6>(row6, 6 * i, g6v);
            row_load<9>(rowR, 9 * i, Rt);
            rot6d_fwd(a6, R);
            if (i == 0) {
#pragma unroll
              for (int k = 0; k < 9; ++k) dR[k] = gR[0][k];
            } else {
              const int p = tr.parent(i);
              float g[3];
#pragma unroll
              for (int a = 0; a < 3; ++a) {
                g[a] = s_dp[(3 * i + a) * 32 + lane] + gp[i][a];
                gp[p][a] += g[a];
#pragma unroll
                for (int b = 0; b < 3; ++b) gR[p][a * 3 + b] += g[a] * s_off[3 * i + b];
              }
              if (!tr.leaf(i)) {
#pragma unroll
                for (int a = 0; a < 3; ++a)
#pragma unroll
                  for (int b = 0; b < 3; ++b)
                    gR[p][a * 3 + b] += gR[i][a * 3] * R[b * 3] + gR[i][a * 3 + 1] * R[b * 3 + 1] + gR[i][a * 3 + 2] * R[b * 3 + 2];
                const int s = tr.slot(p);
                float P[9];
#pragma unroll
                for (int k = 0; k < 9; ++k) P[k] = s_rg[(s * 9 + k) * 32 + lane];
#pragma unroll
                for (int a = 0; a < 3; ++a)
#pragma unroll
                  for (int b = 0; b < 3; ++b)
                    dR[a * 3 + b] = P[a] * gR[i][b] + P[3 + a] * gR[i][3 + b] + P[6 + a] * gR[i][6 + b];
              } else {
#pragma unroll
                for (int k = 0; k < 9; ++k) dR[k] = 0.f;
              }
            }
#pragma unroll
            for (int k = 0; k < 9; ++k) dR[k] += srot * (R[k] - Rt[k]);
            float out[6];
            rot6d_bwd(a6, dR, out);
#pragma unroll
            for (int k = 0; k < 6; ++k) out[k] += s6 * (a6[k] - g6v[k]);
            if (NCW) {
#pragma unroll
              for (int k = 0; k < 6; ++k) obase[(long)(6 * i + k) * T] = out[k];
            } else {
              dw.put(rowx, i, out, false, J);
            }
          }
        }
      }
    }
    if (!NCW && dx6) {
      if (BULK) {
        fence_proxy_async();
        __syncwarp();
        if (valid) bulk_s2g(dx6 + f * r6, s_x6 + lane * p6, r6 * 4);
        bulk_commit();
        bulk_wait_read_all();
      } else {
        __syncwarp();
        for (int e = lane; e < nvalid * r6; e += 32) dx6[f0 * r6 + e] = s_x6[(e / r6) * p6 + e % r6];
      }
    }
    __syncwarp();
  }
  if (BULK) bulk_wait_all();
  l6 = warp_sum(l6);
  lrot = warp_sum(lrot);
  lpos = warp_sum(lpos);
  if (lane == 0) {
    atomicAdd(losses + 0, l6);
    atomicAdd(losses + 1, lrot);
    atomicAdd(losses + 2, lpos);
  }
}

template <class Tree, bool NCW, bool BULK>
static int launch_recon(const float* x6p, const float* gt6, const float* gtR, const float* offsets, int B, int T, float s6,
                        float srot, float spos, float* losses, float* dx6, float* pos_out, float* gtpos_out,
                        const TreeTable& tab, cudaStream_t st) {
  const int J = tab.J;
  const int r6 = 6 * J, r9 = 9 * J, r3 = 3 * J;
  size_t fl = 32 * (size_t)rc_pitch(r6) + 32 * (size_t)rc_pitch(r9) + (NCW ? (size_t)r6 * 32 : 32 * (size_t)rc_pitch(r6)) +
              (size_t)r3 * 32 + (size_t)tab.nslots * 9 * 32 + r3 + 4;
  size_t bytes = fl * 4;
  auto k = recon_kernel<Tree, NCW, BULK>;
  HMVAE_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  const long tiles = ((long)B * T + 31) / 32;
  int per_sm = (int)((227 * 1024) / (bytes + 1024));
  if (per_sm < 1) per_sm = 1;
  long cap = (long)num_sms() * per_sm;
  k<<<(int)(tiles < cap ? tiles : cap), 32, bytes, st>>>(x6p, gt6, gtR, offsets, B, T, s6, srot, spos, losses, dx6, pos_out,
                                                         gtpos_out, tab);
  return check_launch("recon_fwdbwd");
}

}  // namespace hmvae

using namespace hmvae;

extern "C" int hmvae_recon_fwdbwd(const float* x6_pred, int ncw, const float* gt_6d, const float* gt_rotmat,
                                  const float* offsets, const int* parents, int joints, int batch, int t, float s6,
                                  float srot, float spos, float* losses, float* dx6, float* pos_pred_out,
                                  float* gt_pos_out, void* stream) {
  if (!x6_pred || !gt_6d || !gt_rotmat || !offsets || !parents || !losses) return fail_arg("recon_fwdbwd: null pointer");
  if (batch <= 0 || t <= 0) return 0;
  if (joints < 1 || joints > FK_MAX_J) return fail_arg("recon_fwdbwd: joints must be in [1, 32]");
  TreeTable tab;
  tab.J = joints;
  bool smpl = joints == 24;
  for (int i = 0; i < FK_MAX_J; ++i) { tab.parent[i] = 0; tab.slot[i] = -1; tab.leaf[i] = 1; }
  for (int i = 1; i < joints; ++i) {
    if (parents[i] < 0 || parents[i] >= i) return fail_arg("recon_fwdbwd: parents[i] must satisfy 0 <= parents[i] < i");
    tab.parent[i] = (signed char)parents[i];
    tab.leaf[parents[i]] = 0;
    if (smpl && parents[i] != Smpl24Tree::parent_of(i)) smpl = false;
  }
  if (joints == 1) tab.leaf[0] = 0;
  int s = 0;
  for (int q = 0; q < joints; ++q) {
    bool need = false;
    for (int c = 1; c < joints; ++c)
      if (tab.parent[c] == q && !tab.leaf[c]) need = true;
    if (need) tab.slot[q] = (signed char)s++;
  }
  tab.nslots = s;
  cudaStream_t st = (cudaStream_t)stream;
  const bool al = aligned16(gt_6d) && aligned16(gt_rotmat) && (ncw || (aligned16(x6_pred) && (!dx6 || aligned16(dx6))));
  if (smpl && al) {
    if (ncw) return launch_recon<Smpl24Tree, true, true>(x6_pred, gt_6d, gt_rotmat, offsets, batch, t, s6, srot, spos, losses, dx6, pos_pred_out, gt_pos_out, tab, st);
    return launch_recon<Smpl24Tree, false, true>(x6_pred, gt_6d, gt_rotmat, offsets, batch, t, s6, srot, spos, losses, dx6, pos_pred_out, gt_pos_out, tab, st);
  }
  if (ncw) return launch_recon<RuntimeTree, true, false>(x6_pred, gt_6d, gt_rotmat, offsets, batch, t, s6, srot, spos, losses, dx6, pos_pred_out, gt_pos_out, tab, st);
  return launch_recon<RuntimeTree, false, false>(x6_pred, gt_6d, gt_rotmat, offsets, batch, t, s6, srot, spos, losses, dx6, pos_pred_out, gt_pos_out, tab, st);
}
