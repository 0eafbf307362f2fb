// On-device batch assembly: the step right before the hot path.   sm_100a, CUDA cores, HBM-bound.
//
// Replaces the arithmetic of MotionSeqData.__getitem__ (utils_motion_vae.py:140-187) and rand_rotation_matrix (:17-57), which
// the reference runs per sequence on one DataLoader worker: slicing of the T x 579 feature rows into the seven training tensors,
// standardisation with the AMASS mean / std, the random root-rotation augmentation and the re-derivation of the 6D
// representation from the rotated matrices.  The random draws themselves (crop offset, fps factor, the three uniform numbers)
// stay on the host: they are index / seed logic, not arithmetic.
//
// Column layout (:146-158): [0,144) 6D | [144,360) rotation matrices | [360,432) FK joint positions | [432,504) linear velocity
// | [504,576) angular velocity | [576,579) root velocity.   Algorithmic bytes per frame: 579*4 in, 651*4 out.
// Standardisation is done in float64 and rounded once, exactly like numpy's (float32 - float64) / float64 -> .float().
#include "common.cuh"

namespace hmvae {

constexpr int BA_DIM = 579;

// utils_motion_vae.py:17-57, float64 like numpy; M = (V V^T - I) R
__global__ void rand_rotation_kernel(const double* __restrict__ rnd, double deflection, float* __restrict__ rot, long n) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double PI = 3.141592653589793238462643383279502884;
  const double theta = rnd[3 * i] * 2.0 * deflection * PI;
  const double phi = rnd[3 * i + 1] * 2.0 * PI;
  const double z = rnd[3 * i + 2] * 2.0 * deflection;
  const double r = sqrt(z);
  const double V[3] = {sin(phi) * r, cos(phi) * r, sqrt(2.0 - z)};
  const double st = sin(theta), ct = cos(theta);
  const double R[3][3] = {{ct, st, 0.0}, {-st, ct, 0.0}, {0.0, 0.0, 1.0}};
#pragma unroll
  for (int a = 0; a < 3; ++a)
#pragma unroll
    for (int b = 0; b < 3; ++b) {
      double acc = 0.0;
#pragma unroll
      for (int m = 0; m < 3; ++m) acc += (V[a] * V[m] - (a == m ? 1.0 : 0.0)) * R[m][b];
      rot[9 * i + 3 * a + b] = (float)acc;
    }
}

// one thread per (frame, input column); consecutive threads -> consecutive columns of one frame (coalesced both ways)
__global__ void __launch_bounds__(256) batch_assemble_kernel(const float* __restrict__ raw, const float* __restrict__ root_rot,
                                                             const double* __restrict__ mean, const double* __restrict__ stdv,
                                                             long frames, int T, float* __restrict__ rot6d,
                                                             float* __restrict__ rotmat, float* __restrict__ rot_pos,
                                                             float* __restrict__ joint_pos, float* __restrict__ linear_v,
                                                             float* __restrict__ angular_v, float* __restrict__ root_v) {
  const long total = frames * BA_DIM;
  for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x) {
    const long f = e / BA_DIM;
    const int c = (int)(e - f * BA_DIM);
    const float* row = raw + f * BA_DIM;
    const float* M = root_rot ? root_rot + (f / T) * 9 : nullptr;       // one rotation per sequence
    auto standardise = [&](float x) { return (float)(((double)x - mean[c]) / stdv[c]); };
    // element (r, cc) of the (augmented) rotation matrix of joint j
    auto rot_elem = [&](int j, int r, int cc) {
      if (M == nullptr || j != 0) return row[144 + 9 * j + 3 * r + cc];
      return M[3 * r] * row[144 + cc] + M[3 * r + 1] * row[144 + 3 + cc] + M[3 * r + 2] * row[144 + 6 + cc];
    };
    if (c < 144) {
      if (rot6d) {
        float v = row[c];
        if (M != nullptr) {                       // :179-185: 6D = columns 0 and 1 of every (rotated) matrix
          const int j = c / 6, k = c - 6 * j;
          v = rot_elem(j, k % 3, k / 3);
        }
        rot6d[f * 144 + c] = v;
      }
    } else if (c < 360) {
      if (rotmat) {
        const int q = c - 144, j = q / 9, k = q - 9 * j;
        rotmat[f * 216 + q] = rot_elem(j, k / 3, k % 3);
      }
    } else if (c < 432) {
      if (rot_pos) rot_pos[f * 72 + (c - 360)] = row[c];               // raw FK positions (never rotated by the reference)
      if (joint_pos) joint_pos[f * 72 + (c - 360)] = standardise(row[c]);
    } else if (c < 504) {
      if (linear_v) linear_v[f * 72 + (c - 432)] = standardise(row[c]);
    } else if (c < 576) {
      if (angular_v) angular_v[f * 72 + (c - 504)] = standardise(row[c]);
    } else if (root_v) {
      const int r = c - 576;
      float v = row[c];
      if (M != nullptr) v = M[3 * r] * row[576] + M[3 * r + 1] * row[577] + M[3 * r + 2] * row[578];    // :171-175
      root_v[f * 3 + r] = standardise(v);
    }
  }
}

}  // namespace hmvae

using namespace hmvae;

extern "C" int hmvae_rand_rotation(const double* randnums, double deflection, float* rot, long n, void* stream) {
  if (n <= 0) return 0;
  if (!randnums || !rot) return fail_arg("rand_rotation: null pointer");
  rand_rotation_kernel<<<(unsigned)((n + 127) / 128), 128, 0, (cudaStream_t)stream>>>(randnums, deflection, rot, n);
  return check_launch("rand_rotation");
}

extern "C" int hmvae_batch_assemble(const float* raw, const float* root_rot, const double* mean, const double* stdv, int batch,
                                    int t, float* rot6d, float* rotmat, float* rot_pos, float* joint_pos, float* linear_v,
                                    float* angular_v, float* root_v, void* stream) {
  if (batch <= 0 || t <= 0) return 0;
  if (!raw || !mean || !stdv) return fail_arg("batch_assemble: null pointer");
  const long frames = (long)batch * t;
  long blocks = (frames * BA_DIM + 255) / 256;
  const long cap = (long)num_sms() * 16;
  if (blocks > cap) blocks = cap;
  batch_assemble_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(raw, root_rot, mean, stdv, frames, t, rot6d, rotmat,
                                                                          rot_pos, joint_pos, linear_v, angular_v, root_v);
  return check_launch("batch_assemble");
}
