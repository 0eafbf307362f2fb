// Device-side building blocks for forward kinematics and the 6D <-> rotation-matrix transform.
//
// Mapping (DESIGN.md "FK"): one warp owns a group of 32 consecutive frames, one lane per frame.  Frame rows are
// staged in shared memory with 1-D bulk (TMA) copies; each lane walks the kinematic tree for its frame with the
// parent indices resolved at compile time (SMPL-24 instantiation: every array index below is a constant after
// unrolling, so global rotations live in registers) or from a table (generic instantiation: the same code, the
// per-joint arrays become lane-interleaved local memory).
//
// Semantics restated from fk_layer.py:47-93 and my_tools.py:6-39 (see oracle/hmvae_ref.py).
#pragma once
#include "common.cuh"

namespace hmvae {

constexpr int FK_MAX_J = 32;

// ---------------------------------------------------------------- trees
struct Smpl24Tree {
  static constexpr int JMAX = 24;
  static constexpr bool kStatic = true;
  __host__ __device__ static constexpr int parent_of(int i) {
    constexpr int P[24] = {0, 0, 0, 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 9, 9, 12, 13, 14, 16, 17, 18, 19, 20, 21};
    return P[i];
  }
  __host__ __device__ static constexpr bool is_leaf(int i) {
    bool leaf = true;
    for (int c = 1; c < 24; ++c)
      if (parent_of(c) == i) leaf = false;
    return leaf;
  }
  // Rg of joint i must be kept for the backward pass iff i is the parent of a non-leaf joint
  __host__ __device__ static constexpr int slot_of(int i) {
    int s = 0;
    for (int q = 0; q < 24; ++q) {
      bool need = false;
      for (int c = 1; c < 24; ++c)
        if (parent_of(c) == q && !is_leaf(c)) need = true;
      if (q == i) return need ? s : -1;
      if (need) ++s;
    }
    return -1;
  }
  static constexpr int kSlots = 14;
  __device__ __forceinline__ int joints() const { return 24; }
  __device__ __forceinline__ int parent(int i) const { return parent_of(i); }
  __device__ __forceinline__ bool leaf(int i) const { return is_leaf(i); }
  __device__ __forceinline__ int slot(int i) const { return slot_of(i); }
};

struct TreeTable {  // passed by value as a kernel parameter
  int J;
  int nslots;
  signed char parent[FK_MAX_J];
  signed char slot[FK_MAX_J];
  unsigned char leaf[FK_MAX_J];
};

struct RuntimeTree {
  static constexpr int JMAX = FK_MAX_J;
  static constexpr bool kStatic = false;
  const TreeTable* t;
  __device__ __forceinline__ int joints() const { return t->J; }
  __device__ __forceinline__ int parent(int i) const { return t->parent[i]; }
  __device__ __forceinline__ bool leaf(int i) const { return t->leaf[i] != 0; }
  __device__ __forceinline__ int slot(int i) const { return t->slot[i]; }
};

// ---------------------------------------------------------------- small algebra
__device__ __forceinline__ void mat_mul(const float* A, const float* B, float* C) {  // C = A*B (row-major 3x3)
#pragma unroll
  for (int a = 0; a < 3; ++a)
#pragma unroll
    for (int b = 0; b < 3; ++b) C[a * 3 + b] = A[a * 3 + 0] * B[0 * 3 + b] + A[a * 3 + 1] * B[1 * 3 + b] + A[a * 3 + 2] * B[2 * 3 + b];
}

// my_tools.py:19-39.  R = [x | y | z] as columns; normalize(v) = v / max(|v|, 1e-6).
__device__ __forceinline__ void rot6d_fwd(const float* a6, float* R) {
  const float eps = 1e-6f;
  float ax = a6[0], ay = a6[1], az = a6[2], bx = a6[3], by = a6[4], bz = a6[5];
  float na = sqrtf(ax * ax + ay * ay + az * az);
  float ia = 1.f / fmaxf(na, eps);
  float xx = ax * ia, xy = ay * ia, xz = az * ia;
  float zx = xy * bz - xz * by, zy = xz * bx - xx * bz, zz = xx * by - xy * bx;
  float nz = sqrtf(zx * zx + zy * zy + zz * zz);
  float iz = 1.f / fmaxf(nz, eps);
  zx *= iz; zy *= iz; zz *= iz;
  float yx = zy * xz - zz * xy, yy = zz * xx - zx * xz, yz = zx * xy - zy * xx;
  R[0] = xx; R[1] = yx; R[2] = zx;
  R[3] = xy; R[4] = yy; R[5] = zy;
  R[6] = xz; R[7] = yz; R[8] = zz;
}

// backward of rot6d_fwd: G = dL/dR (row-major) -> g6 = dL/da6
__device__ __forceinline__ void rot6d_bwd(const float* a6, const float* G, float* g6) {
  const float eps = 1e-6f;
  float ax = a6[0], ay = a6[1], az = a6[2], bx = a6[3], by = a6[4], bz = a6[5];
  float na = sqrtf(ax * ax + ay * ay + az * az);
  float da = fmaxf(na, eps);
  float ia = 1.f / da;
  float xx = ax * ia, xy = ay * ia, xz = az * ia;
  float z0x = xy * bz - xz * by, z0y = xz * bx - xx * bz, z0z = xx * by - xy * bx;
  float nz = sqrtf(z0x * z0x + z0y * z0y + z0z * z0z);
  float dz = fmaxf(nz, eps);
  float iz = 1.f / dz;
  float zx = z0x * iz, zy = z0y * iz, zz = z0z * iz;
  float gxx = G[0], gxy = G[3], gxz = G[6];
  float gyx = G[1], gyy = G[4], gyz = G[7];
  float gzx = G[2], gzy = G[5], gzz = G[8];
  // y = z x x :  gz += x x gy ; gx += gy x z
  gzx += xy * gyz - xz * gyy; gzy += xz * gyx - xx * gyz; gzz += xx * gyy - xy * gyx;
  gxx += gyy * zz - gyz * zy; gxy += gyz * zx - gyx * zz; gxz += gyx * zy - gyy * zx;
  // z = z0 / max(|z0|, eps)
  float g0x, g0y, g0z;
  if (nz > eps) {
    float d = zx * gzx + zy * gzy + zz * gzz;
    g0x = (gzx - zx * d) * iz; g0y = (gzy - zy * d) * iz; g0z = (gzz - zz * d) * iz;
  } else {
    g0x = gzx * iz; g0y = gzy * iz; g0z = gzz * iz;
  }
  // z0 = x x b :  gx += b x g0 ; gb = g0 x x
  gxx += by * g0z - bz * g0y; gxy += bz * g0x - bx * g0z; gxz += bx * g0y - by * g0x;
  g6[3] = g0y * xz - g0z * xy; g6[4] = g0z * xx - g0x * xz; g6[5] = g0x * xy - g0y * xx;
  // x = a / max(|a|, eps)
  if (na > eps) {
    float d = xx * gxx + xy * gxy + xz * gxz;
    g6[0] = (gxx - xx * d) * ia; g6[1] = (gxy - xy * d) * ia; g6[2] = (gxz - xz * d) * ia;
  } else {
    g6[0] = gxx * ia; g6[1] = gxy * ia; g6[2] = gxz * ia;
  }
}

// ---------------------------------------------------------------- smem row access (16-byte vector loads)
// Reads N consecutive floats starting at element e0 (a compile-time constant after unrolling) of a 16-byte aligned row.
template <int N>
__device__ __forceinline__ void row_load(const float* row, int e0, float* out) {
  const float4* r4 = reinterpret_cast<const float4*>(row);
  const int q0 = e0 >> 2, sh = e0 & 3;
  constexpr int NQ = (N + 3 + 3) / 4;  // worst case chunks
  float buf[NQ * 4];
#pragma unroll
  for (int q = 0; q < NQ; ++q) {
    if (q * 4 < sh + N) {
      float4 v = r4[q0 + q];
      buf[q * 4 + 0] = v.x; buf[q * 4 + 1] = v.y; buf[q * 4 + 2] = v.z; buf[q * 4 + 3] = v.w;
    }
  }
#pragma unroll
  for (int k = 0; k < N; ++k) out[k] = buf[sh + k];
}

// Buffers STRIDE floats per joint and flushes 16-byte chunks to an smem row once GROUP joints are complete.
// Works for ascending (fwd) or descending (bwd) joint order because flush happens on whole aligned groups.
template <int STRIDE>
struct RowWriter {
  static constexpr int GROUP = 4;  // 4 joints * STRIDE floats is always a multiple of 4 floats
  float buf[GROUP * STRIDE];
  __device__ __forceinline__ void put(float* row, int i, const float* v, bool ascending, int nj) {
#pragma unroll
    for (int k = 0; k < STRIDE; ++k) buf[(i % GROUP) * STRIDE + k] = v[k];
    bool done = ascending ? ((i % GROUP) == GROUP - 1 || i == nj - 1) : ((i % GROUP) == 0);
    if (done) {
      const int g0 = (i / GROUP) * GROUP;
      float4* w4 = reinterpret_cast<float4*>(row + g0 * STRIDE);
      const int cnt = ((nj - g0 < GROUP ? nj - g0 : GROUP) * STRIDE + 3) / 4;
#pragma unroll
      for (int q = 0; q < GROUP * STRIDE / 4; ++q)
        if (q < cnt) w4[q] = make_float4(buf[q * 4], buf[q * 4 + 1], buf[q * 4 + 2], buf[q * 4 + 3]);
    }
  }
};

}  // namespace hmvae
