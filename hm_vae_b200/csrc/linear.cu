// Latent heads (nn.Linear, seq_two_hier_sa_vae.py:132-136, 159-164, 225-229, 267): tiny fp32 GEMMs
// ([B*E, 384] x [384 -> 24/48] and [B*E, 12/24] -> 384).  One generic smem-tiled CUDA-core kernel with arbitrary strides
// covers y = x W^T + b, dx = dy W, dW = dy^T x; a column-sum kernel gives db.  fp32 exact (these layers are 0.1 % of the FLOPs).
#include "common.cuh"

namespace hmvae {

constexpr int LT = 32;   // tile

// C[m, n] = sum_k A(m, k) * B(k, n) (+ bias[n]);  A(m,k) = a[m*am + k*ak], B(k,n) = b[k*bk + n*bn], C row-major [M, N]
__global__ void __launch_bounds__(256) small_gemm_kernel(const float* __restrict__ a, long am, long ak,
                                                         const float* __restrict__ b, long bk, long bn,
                                                         const float* __restrict__ bias, float* __restrict__ c, int M, int N,
                                                         int K) {
  __shared__ float As[LT][LT + 1], Bs[LT][LT + 1];
  const int m0 = blockIdx.y * LT, n0 = blockIdx.x * LT;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;     // 8 x 32
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int k0 = 0; k0 < K; k0 += LT) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int r = ty + i * 8;
      // A tile: fast thread index along the unit-stride dimension
      {
        const int mm = (ak == 1) ? m0 + r : m0 + tx, kk = (ak == 1) ? k0 + tx : k0 + r;
        const float v = (mm < M && kk < K) ? a[mm * am + kk * ak] : 0.f;
        if (ak == 1) As[r][tx] = v; else As[tx][r] = v;
      }
      {
        const int kk = (bn == 1) ? k0 + r : k0 + tx, nn = (bn == 1) ? n0 + tx : n0 + r;
        const float v = (kk < K && nn < N) ? b[kk * bk + nn * bn] : 0.f;
        if (bn == 1) Bs[r][tx] = v; else Bs[tx][r] = v;
      }
    }
    __syncthreads();
#pragma unroll 8
    for (int kk = 0; kk < LT; ++kk) {
      const float bv = Bs[kk][tx];
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[i] += As[ty + i * 8][kk] * bv;
    }
    __syncthreads();
  }
  const int n = n0 + tx;
  if (n < N) {
    const float bb = bias ? bias[n] : 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int m = m0 + ty + i * 8;
      if (m < M) c[(long)m * N + n] = acc[i] + bb;
    }
  }
}

// out[n] = sum_m x[m, n]   (x row-major [M, N]); one warp per 32 columns chunk, fixed order
__global__ void __launch_bounds__(256) colsum_kernel(const float* __restrict__ x, float* __restrict__ out, int M, int N) {
  __shared__ float red[8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int n = blockIdx.x * 32 + tx;
  float acc = 0.f;
  if (n < N)
    for (int m = ty; m < M; m += 8) acc += x[(long)m * N + n];
  red[ty][tx] = acc;
  __syncthreads();
  if (ty == 0 && n < N) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += red[i][tx];
    out[n] = s;
  }
}

static int gemm(const float* a, long am, long ak, const float* b, long bk, long bn, const float* bias, float* c, int M, int N,
                int K, cudaStream_t st, const char* what) {
  dim3 grid((N + LT - 1) / LT, (M + LT - 1) / LT);
  small_gemm_kernel<<<grid, 256, 0, st>>>(a, am, ak, b, bk, bn, bias, c, M, N, K);
  return check_launch(what);
}

}  // namespace hmvae

using namespace hmvae;

extern "C" int hmvae_linear_fwd(const float* x, const float* w, const float* bias, float* y, int rows, int in_f, int out_f,
                                void* stream) {
  if (!x || !w || !y) return fail_arg("linear_fwd: null pointer");
  if (rows <= 0 || in_f <= 0 || out_f <= 0) return 0;
  // y[r, o] = sum_i x[r, i] * w[o, i]
  return gemm(x, in_f, 1, w, 1, in_f, bias, y, rows, out_f, in_f, (cudaStream_t)stream, "linear_fwd");
}

extern "C" int hmvae_linear_bwd(const float* x, const float* w, const float* dy, float* dx, float* dw, float* db, int rows,
                                int in_f, int out_f, void* stream) {
  if (!dy || (dx && !w) || (dw && !x)) return fail_arg("linear_bwd: null pointer");
  if (rows <= 0 || in_f <= 0 || out_f <= 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  int rc = 0;
  if (dx) rc = gemm(dy, out_f, 1, w, in_f, 1, nullptr, dx, rows, in_f, out_f, st, "linear_bwd(dx)");        // dx[r,i] = sum_o dy[r,o] w[o,i]
  if (!rc && dw) rc = gemm(dy, 1, out_f, x, in_f, 1, nullptr, dw, out_f, in_f, rows, st, "linear_bwd(dw)");  // dw[o,i] = sum_r dy[r,o] x[r,i]
  if (!rc && db) {
    colsum_kernel<<<(out_f + 31) / 32, 256, 0, st>>>(dy, db, rows, out_f);
    rc = check_launch("linear_bwd(db)");
  }
  return rc;
}
