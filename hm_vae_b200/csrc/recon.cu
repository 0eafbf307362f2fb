// Fused reconstruction loss: GT FK, rot6d->R, FK on the prediction, the three MSE sums AND d(total)/d(x6_pred)
// in one kernel.  Replaces seq_two_hier_sa_vae.py:343 (FK on GT), :441-468 (transpose, rot6d->R, FK),
// :395-397 + :430-434 (3x MSE) and the autograd replay of all of them (~400 launches in the reference).
// R_pred, pos_pred, pos_gt and every intermediate gradient stay on chip.
//
// Algorithmic bytes per frame (J=24): 144*4 (x6) + 144*4 (gt 6d) + 216*4 (gt R) in, 144*4 (dx6) out = 2592 B.
#include <string.h>

#include "fk_device.cuh"

namespace hmvae {

// ---------------------------------------------------------------------------------------------------------------------
// Mapping: one WARP per frame, one LANE per joint, RC_FR consecutive frames per CTA.  A training step has only B*T = 2048
// frames, so a lane-per-frame walk of the 24-joint tree (3 sequential FK passes) is a ~80 us latency chain on 64 warps;
// here every joint of every frame works in parallel and the tree is resolved by
//   * forward : each lane runs the Horner chain of its own ancestors (<= depth matvecs), reading the local rotations of
//               the frame from shared memory:  pos_i = off_0 + R_0 (off_c1 + R_c1 ( ... + R_p(i) off_i)),
//               and the matrix chain Rg_p(i) = R_0 ... R_p(i) it needs for the backward pass;
//   * backward: a level-synchronous bottom-up sweep (one __syncwarp per tree level):
//               S_i = g_i + sum_c S_c,   G_i = sum_c (S_c (x) off_c + G_c R_c^T)   over the children c of i,
//               dL/dR_i = Rg_p(i)^T G_i  (fk_layer.py:47-93 differentiated; identical to the reverse accumulation).
// The decoder output is read in its NCW layout through a [channel][frame] shared tile (8 consecutive time steps = one
// 32-byte sector per channel) and dx6 leaves through the same tile.
struct ParTree {
  int J, maxdepth;
  signed char parent[FK_MAX_J];
  signed char depth[FK_MAX_J];
  signed char child_off[FK_MAX_J + 1];
  signed char child_idx[FK_MAX_J];
};

constexpr int RC_FR = 8;          // frames (warps) per CTA
constexpr int RC_TP = RC_FR + 1;  // padded tile pitch

__device__ __forceinline__ void matvec3(const float* M, const float* v, float* o) {
#pragma unroll
  for (int a = 0; a < 3; ++a) o[a] = M[a * 3] * v[0] + M[a * 3 + 1] * v[1] + M[a * 3 + 2] * v[2];
}

template <bool NCW>
__global__ void __launch_bounds__(32 * RC_FR) recon_par_kernel(const float* __restrict__ x6p, const float* __restrict__ gt6,
                                                               const float* __restrict__ gtR,
                                                               const float* __restrict__ offsets, int B, int T, float s6,
                                                               float srot, float spos, float* __restrict__ losses,
                                                               float* __restrict__ dx6, float* __restrict__ pos_out,
                                                               float* __restrict__ gtpos_out, const float* __restrict__ mask,
                                                               float* __restrict__ rot_out, ParTree tr) {
  pdl_trigger();
  pdl_wait();
  __shared__ float s_tile[NCW ? 6 * FK_MAX_J * RC_TP : 1];
  __shared__ float s_R[RC_FR][FK_MAX_J][9], s_Rt[RC_FR][FK_MAX_J][9], s_G[RC_FR][FK_MAX_J][9], s_S[RC_FR][FK_MAX_J][3];
  __shared__ float s_off[FK_MAX_J * 3];
  __shared__ float s_red[RC_FR][3];
  const int tid = threadIdx.x, w = tid >> 5, i = tid & 31;
  const int J = tr.J;
  const int r6 = 6 * J, r9 = 9 * J, r3 = 3 * J;
  const long n = (long)B * T;
  const long f0 = (long)blockIdx.x * RC_FR;
  const long f = f0 + w;
  const bool act = f < n && i < J;

  for (int e = tid; e < r3; e += 32 * RC_FR) s_off[e] = offsets[e];
  if (NCW) {
    for (int e = tid; e < r6 * RC_FR; e += 32 * RC_FR) {
      const int c = e / RC_FR, tl = e % RC_FR;
      const long ff = f0 + tl;
      if (ff < n) s_tile[c * RC_TP + tl] = x6p[((ff / T) * r6 + c) * (long)T + (ff % T)];
    }
  }
  __syncthreads();

  float a6[6], g6v[6], R[9], Rt[9], M[9], G[9];
  float l6 = 0.f, lrot = 0.f, lpos = 0.f;
  // per (frame, joint) weight of the squared errors: l2_masked_criterion (seq_two_hier_sa_vae.py:717-735) multiplies the
  // element-wise squared error by mask[b, t, joint] before the mean over ALL elements; no mask = plain l2_criterion
  const float mk = (mask && act) ? mask[f * J + i] : 1.f;
#pragma unroll
  for (int k = 0; k < 9; ++k) { R[k] = 0.f; Rt[k] = 0.f; M[k] = 0.f; G[k] = 0.f; }
#pragma unroll
  for (int k = 0; k < 6; ++k) { a6[k] = 0.f; g6v[k] = 0.f; }
  if (act) {
#pragma unroll
    for (int k = 0; k < 6; ++k) {
      a6[k] = NCW ? s_tile[(6 * i + k) * RC_TP + w] : x6p[f * r6 + 6 * i + k];
      g6v[k] = gt6[f * r6 + 6 * i + k];
    }
#pragma unroll
    for (int k = 0; k < 9; ++k) Rt[k] = gtR[f * r9 + 9 * i + k];
    rot6d_fwd(a6, R);
#pragma unroll
    for (int k = 0; k < 6; ++k) { const float d = a6[k] - g6v[k]; l6 += mk * d * d; }
#pragma unroll
    for (int k = 0; k < 9; ++k) {
      const float d = R[k] - Rt[k];
      lrot += mk * d * d;
      if (rot_out) rot_out[f * r9 + 9 * i + k] = R[k];
      s_R[w][i][k] = R[k];
      s_Rt[w][i][k] = Rt[k];
    }
  }
  __syncwarp();

  // ---- forward: Horner chains over the ancestors (prediction and ground truth together)
  float gpos[3] = {0.f, 0.f, 0.f};
  if (act) {
    float v[3] = {0.f, 0.f, 0.f}, vg[3] = {0.f, 0.f, 0.f}, t3[3];
    if (i > 0) {
      v[0] = vg[0] = s_off[3 * i]; v[1] = vg[1] = s_off[3 * i + 1]; v[2] = vg[2] = s_off[3 * i + 2];
      int a = tr.parent[i];
#pragma unroll
      for (int k = 0; k < 9; ++k) M[k] = s_R[w][a][k];
      matvec3(M, v, t3); v[0] = t3[0]; v[1] = t3[1]; v[2] = t3[2];
      matvec3(s_Rt[w][a], vg, t3); vg[0] = t3[0]; vg[1] = t3[1]; vg[2] = t3[2];
      while (a != 0) {
        v[0] += s_off[3 * a]; v[1] += s_off[3 * a + 1]; v[2] += s_off[3 * a + 2];
        vg[0] += s_off[3 * a]; vg[1] += s_off[3 * a + 1]; vg[2] += s_off[3 * a + 2];
        a = tr.parent[a];
        float Ra[9], M2[9];
#pragma unroll
        for (int k = 0; k < 9; ++k) Ra[k] = s_R[w][a][k];
        matvec3(Ra, v, t3); v[0] = t3[0]; v[1] = t3[1]; v[2] = t3[2];
        matvec3(s_Rt[w][a], vg, t3); vg[0] = t3[0]; vg[1] = t3[1]; vg[2] = t3[2];
        mat_mul(Ra, M, M2);
#pragma unroll
        for (int k = 0; k < 9; ++k) M[k] = M2[k];
      }
    }
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      const float pp = s_off[a] + v[a], pg = s_off[a] + vg[a];
      const float d = pp - pg;
      lpos += mk * d * d;
      gpos[a] = spos * mk * d;
      s_S[w][i][a] = gpos[a];
      if (pos_out) pos_out[f * r3 + 3 * i + a] = pp;
      if (gtpos_out) gtpos_out[f * r3 + 3 * i + a] = pg;
    }
  }
  __syncwarp();

  // ---- backward: bottom-up sweep, one tree level per round
  if (dx6) {
    const int myd = (i < J) ? tr.depth[i] : -1;
    const int c0 = (i < J) ? tr.child_off[i] : 0, c1 = (i < J) ? tr.child_off[i + 1] : 0;
    for (int lvl = tr.maxdepth - 1; lvl >= 0; --lvl) {
      if (act && myd == lvl && c1 > c0) {
        float S[3] = {gpos[0], gpos[1], gpos[2]};
        for (int ci = c0; ci < c1; ++ci) {
          const int c = tr.child_idx[ci];
          const float Sc[3] = {s_S[w][c][0], s_S[w][c][1], s_S[w][c][2]};
#pragma unroll
          for (int a = 0; a < 3; ++a) {
            S[a] += Sc[a];
#pragma unroll
            for (int b = 0; b < 3; ++b) G[a * 3 + b] += Sc[a] * s_off[3 * c + b];
          }
          if (tr.child_off[c + 1] > tr.child_off[c]) {      // G_c R_c^T
            float Gc[9], Rc[9];
#pragma unroll
            for (int k = 0; k < 9; ++k) { Gc[k] = s_G[w][c][k]; Rc[k] = s_R[w][c][k]; }
#pragma unroll
            for (int a = 0; a < 3; ++a)
#pragma unroll
              for (int b = 0; b < 3; ++b)
                G[a * 3 + b] += Gc[a * 3] * Rc[b * 3] + Gc[a * 3 + 1] * Rc[b * 3 + 1] + Gc[a * 3 + 2] * Rc[b * 3 + 2];
          }
        }
#pragma unroll
        for (int a = 0; a < 3; ++a) s_S[w][i][a] = S[a];
#pragma unroll
        for (int k = 0; k < 9; ++k) s_G[w][i][k] = G[k];
      }
      __syncwarp();
    }
    if (act) {
      float dR[9], out[6];
      if (i == 0) {
#pragma unroll
        for (int k = 0; k < 9; ++k) dR[k] = G[k];
      } else {
#pragma unroll
        for (int a = 0; a < 3; ++a)
#pragma unroll
          for (int b = 0; b < 3; ++b) dR[a * 3 + b] = M[a] * G[b] + M[3 + a] * G[3 + b] + M[6 + a] * G[6 + b];
      }
#pragma unroll
      for (int k = 0; k < 9; ++k) dR[k] += srot * mk * (R[k] - Rt[k]);
      rot6d_bwd(a6, dR, out);
#pragma unroll
      for (int k = 0; k < 6; ++k) {
        const float o = out[k] + s6 * mk * (a6[k] - g6v[k]);
        if (NCW) s_tile[(6 * i + k) * RC_TP + w] = o;
        else dx6[f * r6 + 6 * i + k] = o;
      }
    }
  }
  // ---- loss sums: warp -> CTA -> one atomic per CTA and loss
  l6 = warp_sum(l6);
  lrot = warp_sum(lrot);
  lpos = warp_sum(lpos);
  if (i == 0) { s_red[w][0] = l6; s_red[w][1] = lrot; s_red[w][2] = lpos; }
  __syncthreads();
  if (tid < 3) {
    float s = 0.f;
#pragma unroll
    for (int q = 0; q < RC_FR; ++q) s += s_red[q][tid];
    atomicAdd(losses + tid, s);
  }
  if (NCW && dx6) {
    for (int e = tid; e < r6 * RC_FR; e += 32 * RC_FR) {
      const int c = e / RC_FR, tl = e % RC_FR;
      const long ff = f0 + tl;
      if (ff < n) dx6[((ff / T) * r6 + c) * (long)T + (ff % T)] = s_tile[c * RC_TP + tl];
    }
  }
}

template <bool NCW>
static int launch_recon_par(const float* x6p, const float* gt6, const float* gtR, const float* offsets, int B, int T, float s6,
                            float srot, float spos, float* losses, float* dx6, float* pos_out, float* gtpos_out,
                            const float* mask, float* rot_out, const ParTree& tr, cudaStream_t st) {
  const long tiles = ((long)B * T + RC_FR - 1) / RC_FR;
  if (tiles > 0x7fffffffL) return fail_arg("recon_fwdbwd: too many frames");
  launch_pdl(recon_par_kernel<NCW>, dim3((int)tiles), dim3(32 * RC_FR), 0, st, x6p, gt6, gtR, offsets, B, T, s6, srot, spos, losses, dx6, pos_out, gtpos_out, mask, rot_out, tr);
  return check_launch("recon_fwdbwd");
}

}  // namespace hmvae

using namespace hmvae;

extern "C" int hmvae_recon_masked_fwdbwd(const float* x6_pred, int ncw, const float* gt_6d, const float* gt_rotmat,
                                         const float* mask, const float* offsets, const int* parents, int joints, int batch,
                                         int t, float s6, float srot, float spos, float* losses, float* dx6,
                                         float* pos_pred_out, float* gt_pos_out, float* rot_pred_out, void* stream) {
  if (!x6_pred || !gt_6d || !gt_rotmat || !offsets || !parents || !losses) return fail_arg("recon_fwdbwd: null pointer");
  if (batch <= 0 || t <= 0) return 0;
  if (joints < 1 || joints > FK_MAX_J) return fail_arg("recon_fwdbwd: joints must be in [1, 32]");
  ParTree tr;
  memset(&tr, 0, sizeof(tr));
  tr.J = joints;
  int nchild[FK_MAX_J] = {0};
  for (int i = 1; i < joints; ++i) {
    if (parents[i] < 0 || parents[i] >= i) return fail_arg("recon_fwdbwd: parents[i] must satisfy 0 <= parents[i] < i");
    tr.parent[i] = (signed char)parents[i];
    tr.depth[i] = (signed char)(tr.depth[parents[i]] + 1);
    if (tr.depth[i] > tr.maxdepth) tr.maxdepth = tr.depth[i];
    ++nchild[parents[i]];
  }
  for (int i = 0; i < joints; ++i) tr.child_off[i + 1] = (signed char)(tr.child_off[i] + nchild[i]);
  int fill[FK_MAX_J] = {0};
  for (int i = 1; i < joints; ++i) {               // children in ascending order: fixed summation order
    const int p = parents[i];
    tr.child_idx[tr.child_off[p] + fill[p]++] = (signed char)i;
  }
  cudaStream_t st = (cudaStream_t)stream;
  if (ncw) return launch_recon_par<true>(x6_pred, gt_6d, gt_rotmat, offsets, batch, t, s6, srot, spos, losses, dx6, pos_pred_out, gt_pos_out, mask, rot_pred_out, tr, st);
  return launch_recon_par<false>(x6_pred, gt_6d, gt_rotmat, offsets, batch, t, s6, srot, spos, losses, dx6, pos_pred_out, gt_pos_out, mask, rot_pred_out, tr, st);
}

extern "C" int hmvae_recon_fwdbwd(const float* x6_pred, int ncw, const float* gt_6d, const float* gt_rotmat,
                                  const float* offsets, const int* parents, int joints, int batch, int t, float s6,
                                  float srot, float spos, float* losses, float* dx6, float* pos_pred_out,
                                  float* gt_pos_out, void* stream) {
  return hmvae_recon_masked_fwdbwd(x6_pred, ncw, gt_6d, gt_rotmat, nullptr, offsets, parents, joints, batch, t, s6, srot, spos,
                                   losses, dx6, pos_pred_out, gt_pos_out, nullptr, stream);
}
