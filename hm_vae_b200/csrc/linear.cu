// Latent heads (nn.Linear, seq_two_hier_sa_vae.py:132-136, 159-164, 225-229, 267): tiny fp32 GEMMs
// ([B*E, 384] x [384 -> 24/48] and [B*E, 12/24] -> 384).  One generic smem-tiled CUDA-core kernel with arbitrary strides
// covers y = x W^T + b, dx = dy W, dW = dy^T x; a column-sum kernel gives db.  fp32 exact (these layers are 0.1 % of the FLOPs).
#include "common.cuh"

namespace hmvae {

// y[r, o] = sum_i x[r, i] * w[o, i] + bias[o].  The whole weight (<= 96 KB) and RB rows of x are staged in shared memory with all
// loads in flight at once (these GEMMs are latency-, not throughput-bound); odd pitch => conflict-free.
constexpr int LIN_RB = 8;
__global__ void __launch_bounds__(256) linear_nt_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                        const float* __restrict__ bias, float* __restrict__ y, int R, int I,
                                                        int O) {
  extern __shared__ float sm[];
  const int pitch = I | 1;
  float* ws = sm;                    // [O][pitch]
  float* xs = sm + (size_t)O * pitch;   // [RB][I]
  const int r0 = blockIdx.x * LIN_RB;
  for (int e = threadIdx.x; e < O * I; e += 256) ws[(e / I) * pitch + e % I] = w[e];
  for (int e = threadIdx.x; e < LIN_RB * I; e += 256) {
    const int r = r0 + e / I;
    xs[e] = r < R ? x[(long)r * I + e % I] : 0.f;
  }
  __syncthreads();
  for (int e = threadIdx.x; e < LIN_RB * O; e += 256) {
    const int rl = e / O, o = e % O;
    if (r0 + rl >= R) continue;
    const float* xr = xs + rl * I;
    const float* wr = ws + o * pitch;
    float acc0 = 0.f, acc1 = 0.f;
    int i = 0;
    for (; i + 1 < I; i += 2) { acc0 += xr[i] * wr[i]; acc1 += xr[i + 1] * wr[i + 1]; }
    if (i < I) acc0 += xr[i] * wr[i];
    y[(long)(r0 + rl) * O + o] = acc0 + acc1 + (bias ? bias[o] : 0.f);
  }
}

// C[m, n] = sum_k A(m, k) * B(k, n),  A(m,k) = a[m*am + k*ak], B(k,n) = b[k*bk + n*bn], C row-major [M, N].
// One thread per output; the caller picks which of (m, n) runs fastest across threads so that the big operand is read
// coalesced and the other one is a warp broadcast.  Independent loads, unrolled => high memory-level parallelism.
__global__ void __launch_bounds__(256) small_gemm_kernel(const float* __restrict__ a, long am, long ak,
                                                         const float* __restrict__ b, long bk, long bn, float* __restrict__ c,
                                                         int M, int N, int K, int n_fastest) {
  const long e = (long)blockIdx.x * 256 + threadIdx.x;
  if (e >= (long)M * N) return;
  const int m = n_fastest ? (int)(e / N) : (int)(e % M);
  const int n = n_fastest ? (int)(e % N) : (int)(e / M);
  const float* ap = a + m * am;
  const float* bp = b + n * bn;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  int k = 0;
  for (; k + 3 < K; k += 4) {
#pragma unroll
    for (int u = 0; u < 4; ++u) acc[u] += ap[(k + u) * ak] * bp[(k + u) * bk];
  }
  for (; k < K; ++k) acc[0] += ap[k * ak] * bp[k * bk];
  c[(long)m * N + n] = (acc[0] + acc[1]) + (acc[2] + acc[3]);
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

static int gemm(const float* a, long am, long ak, const float* b, long bk, long bn, float* c, int M, int N, int K, int n_fastest,
                cudaStream_t st, const char* what) {
  const long total = (long)M * N;
  small_gemm_kernel<<<(int)((total + 255) / 256), 256, 0, st>>>(a, am, ak, b, bk, bn, c, M, N, K, n_fastest);
  return check_launch(what);
}

}  // namespace hmvae

using namespace hmvae;

extern "C" int hmvae_linear_fwd(const float* x, const float* w, const float* bias, float* y, int rows, int in_f, int out_f,
                                void* stream) {
  if (!x || !w || !y) return fail_arg("linear_fwd: null pointer");
  if (rows <= 0 || in_f <= 0 || out_f <= 0) return 0;
  const size_t smem = ((size_t)out_f * (in_f | 1) + (size_t)LIN_RB * in_f) * 4;
  if (smem > 200 * 1024) return fail_arg("linear_fwd: weight does not fit shared memory (latent heads only)");
  if (smem > 48 * 1024) HMVAE_CUDA(cudaFuncSetAttribute(linear_nt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  linear_nt_kernel<<<(rows + LIN_RB - 1) / LIN_RB, 256, smem, (cudaStream_t)stream>>>(x, w, bias, y, rows, in_f, out_f);
  return check_launch("linear_fwd");
}

extern "C" int hmvae_linear_bwd(const float* x, const float* w, const float* dy, float* dx, float* dw, float* db, int rows,
                                int in_f, int out_f, void* stream) {
  if (!dy || (dx && !w) || (dw && !x)) return fail_arg("linear_bwd: null pointer");
  if (rows <= 0 || in_f <= 0 || out_f <= 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  int rc = 0;
  // dx[r,i] = sum_o dy[r,o] w[o,i]   : threads run over i (w coalesced, dy broadcast)
  if (dx) rc = gemm(dy, out_f, 1, w, in_f, 1, dx, rows, in_f, out_f, 1, st, "linear_bwd(dx)");
  // dw[o,i] = sum_r dy[r,o] x[r,i]   : threads run over the wider of (o, i) so that the wider operand is read coalesced
  if (!rc && dw) rc = gemm(dy, 1, out_f, x, in_f, 1, dw, out_f, in_f, rows, in_f >= out_f ? 1 : 0, st, "linear_bwd(dw)");
  if (!rc && db) {
    colsum_kernel<<<(out_f + 31) / 32, 256, 0, st>>>(dy, db, rows, out_f);
    rc = check_launch("linear_bwd(db)");
  }
  return rc;
}
