// Latent heads (nn.Linear, seq_two_hier_sa_vae.py:132-136, 159-164, 225-229, 267): tiny fp32 GEMMs
// ([B*E, 384] x [384 -> 24/48] and [B*E, 12/24] -> 384).  One generic smem-tiled CUDA-core kernel with arbitrary strides
// covers y = x W^T + b, dx = dy W, dW = dy^T x; a column-sum kernel gives db.  fp32 exact (these layers are 0.1 % of the FLOPs).
#include "common.cuh"

namespace hmvae {

// ---- y[r, o] = sum_i x[r, i] * w[o, i] + bias[o]      (both operands contiguous along the reduction)
// Long reduction (encoder heads, I = 384): one warp per row keeps the row in registers (NT_XCH values per lane) and walks the
// outputs; the NT_XCH weight loads of an output are independent and issued together (the first version loaded one weight per
// loop iteration: a chain of L2 latencies, 12 us at B=32 and 80 us at B=512).
constexpr int NT_XCH = 12;
__global__ void __launch_bounds__(128) linear_nt_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                        const float* __restrict__ bias, float* __restrict__ y, int R, int I,
                                                        int O) {
  pdl_trigger();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const long r = (long)blockIdx.x * 4 + (threadIdx.x >> 5);
  if (r >= R) return;
  const float* xr = x + r * I;
  for (int o = 0; o < O; ++o) {
    const float* wr = w + (long)o * I;
    float acc = 0.f;
    for (int i0 = 0; i0 < I; i0 += 32 * NT_XCH) {
      float xv[NT_XCH], wv[NT_XCH];
#pragma unroll
      for (int k = 0; k < NT_XCH; ++k) {
        const int i = i0 + lane + 32 * k;
        xv[k] = i < I ? xr[i] : 0.f;              // L1-resident after the first output
        wv[k] = i < I ? wr[i] : 0.f;
      }
#pragma unroll
      for (int k = 0; k < NT_XCH; ++k) acc = fmaf(xv[k], wv[k], acc);
    }
    acc = warp_sum(acc);
    if (lane == 0) y[r * O + o] = acc + (bias ? bias[o] : 0.f);
  }
}

// Short reduction, wide output (decoder heads, I = 12 / 24, O = 384): a CTA stages NS_ROWS input rows in shared memory, every thread
// owns output columns o = tid, tid + 128, ..: it reads its weight row once and produces the column for all staged rows.
constexpr int NS_ROWS = 8;
__global__ void __launch_bounds__(128) linear_nt_short_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                              const float* __restrict__ bias, float* __restrict__ y, int R, int I,
                                                              int O) {
  extern __shared__ float xs[];                 // [NS_ROWS][I]
  pdl_trigger();
  pdl_wait();
  const long r0 = (long)blockIdx.x * NS_ROWS;
  const int nr = (R - r0 < NS_ROWS) ? (int)(R - r0) : NS_ROWS;
  for (int e = threadIdx.x; e < NS_ROWS * I; e += 128) xs[e] = (e / I) < nr ? x[r0 * I + e] : 0.f;
  __syncthreads();
  for (int o = threadIdx.x; o < O; o += 128) {
    const float* wr = w + (long)o * I;
    float acc[NS_ROWS];
    const float b = bias ? bias[o] : 0.f;
#pragma unroll
    for (int q = 0; q < NS_ROWS; ++q) acc[q] = b;
#pragma unroll 8
    for (int i = 0; i < I; ++i) {
      const float wv = wr[i];
#pragma unroll
      for (int q = 0; q < NS_ROWS; ++q) acc[q] = fmaf(wv, xs[q * I + i], acc[q]);
    }
#pragma unroll
    for (int q = 0; q < NS_ROWS; ++q)
      if (q < nr) y[(r0 + q) * O + o] = acc[q];
  }
}

// ---- dx[r, i] = sum_o dy[r, o] * w[o, i]      (short reduction, wide contiguous output): one thread per output
__global__ void __launch_bounds__(256) linear_nn_kernel(const float* __restrict__ dy, const float* __restrict__ w,
                                                        float* __restrict__ dx, int R, int I, int O) {
  pdl_trigger();
  pdl_wait();
  const long e = (long)blockIdx.x * 256 + threadIdx.x;
  if (e >= (long)R * I) return;
  const int r = (int)(e / I), i = (int)(e % I);
  const float* g = dy + (long)r * O;
  float acc0 = 0.f, acc1 = 0.f;
  int o = 0;
  for (; o + 1 < O; o += 2) { acc0 += g[o] * w[(long)o * I + i]; acc1 += g[o + 1] * w[(long)(o + 1) * I + i]; }
  if (o < O) acc0 += g[o] * w[(long)o * I + i];
  dx[e] = acc0 + acc1;
}

// ---- same product when the reduction is LONG and the output narrow (decoder heads: O = 384, I = 12 / 24): one warp per
// output, lanes stride over o (dy coalesced, the 18 KB weight matrix stays in L1), warp reduction.  The one-thread-per-output
// kernel above ran 384 dependent iterations on 21 CTAs (35 us).
__global__ void __launch_bounds__(256) linear_nn_wide_kernel(const float* __restrict__ dy, const float* __restrict__ w,
                                                             float* __restrict__ dx, int R, int I, int O) {
  pdl_trigger();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const long wid = (long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (wid >= (long)R * I) return;
  const int r = (int)(wid / I), i = (int)(wid % I);
  const float* g = dy + (long)r * O;
  float acc = 0.f;
  for (int o = lane; o < O; o += 32) acc += g[o] * w[(long)o * I + i];
  acc = warp_sum(acc);
  if (lane == 0) dx[wid] = acc;
}

// ---- C[a, b] = sum_r P[r, a] * Q[r, b]      (long reduction over rows; Q wide, P narrow)
// CTA = 8 a x 32 b outputs; the 8 warps split the rows, partial sums are combined through shared memory in a fixed order.
// out index = transpose ? b*A + a : a*Bn + b.
__global__ void __launch_bounds__(256) linear_tn_kernel(const float* __restrict__ P, const float* __restrict__ Q,
                                                        float* __restrict__ out, int R, int A, int Bn, int transpose) {
  pdl_trigger();
  pdl_wait();
  __shared__ float red[8][8][33];
  const int lane = threadIdx.x & 31, ks = threadIdx.x >> 5;
  const int b = blockIdx.x * 32 + lane, a0 = blockIdx.y * 8;
  float acc[8];
#pragma unroll
  for (int u = 0; u < 8; ++u) acc[u] = 0.f;
  if (b < Bn) {
    for (int r = ks; r < R; r += 8) {
      const float q = Q[(long)r * Bn + b];
      const float* pr = P + (long)r * A + a0;
#pragma unroll
      for (int u = 0; u < 8; ++u)
        if (a0 + u < A) acc[u] += pr[u] * q;
    }
  }
#pragma unroll
  for (int u = 0; u < 8; ++u) red[ks][u][lane] = acc[u];
  __syncthreads();
  // thread (ks = u, lane) finalises output (a0 + u, b)
  const int u = ks;
  if (b < Bn && a0 + u < A) {
    float v = 0.f;
#pragma unroll
    for (int k2 = 0; k2 < 8; ++k2) v += red[k2][u][lane];
    out[transpose ? (long)b * A + a0 + u : (long)(a0 + u) * Bn + b] = v;
  }
}

// out[n] = sum_m x[m, n]   (x row-major [M, N]); one warp per 32 columns chunk, fixed order
__global__ void __launch_bounds__(256) colsum_kernel(const float* __restrict__ x, float* __restrict__ out, int M, int N) {
  pdl_trigger();
  pdl_wait();
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

}  // namespace hmvae

using namespace hmvae;

extern "C" int hmvae_linear_fwd(const float* x, const float* w, const float* bias, float* y, int rows, int in_f, int out_f,
                                void* stream) {
  if (!x || !w || !y) return fail_arg("linear_fwd: null pointer");
  if (rows <= 0 || in_f <= 0 || out_f <= 0) return 0;
  if (in_f <= 64 && out_f >= 64)
    launch_pdl(linear_nt_short_kernel, dim3((rows + NS_ROWS - 1) / NS_ROWS), dim3(128), (size_t)NS_ROWS * in_f * sizeof(float),
               (cudaStream_t)stream, x, w, bias, y, rows, in_f, out_f);
  else
    launch_pdl(linear_nt_kernel, dim3((rows + 3) / 4), dim3(128), 0, (cudaStream_t)stream, x, w, bias, y, rows, in_f, out_f);
  return check_launch("linear_fwd");
}

extern "C" int hmvae_linear_bwd(const float* x, const float* w, const float* dy, float* dx, float* dw, float* db, int rows,
                                int in_f, int out_f, void* stream) {
  if (!dy || (dx && !w) || (dw && !x)) return fail_arg("linear_bwd: null pointer");
  if (rows <= 0 || in_f <= 0 || out_f <= 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  if (dx) {
    const long total = (long)rows * in_f;
    if (out_f >= 128 && total <= (1L << 20))
      launch_pdl(linear_nn_wide_kernel, dim3((int)((total + 7) / 8)), dim3(256), 0, st, dy, w, dx, rows, in_f, out_f);
    else
      launch_pdl(linear_nn_kernel, dim3((int)((total + 255) / 256)), dim3(256), 0, st, dy, w, dx, rows, in_f, out_f);
    int rc = check_launch("linear_bwd(dx)");
    if (rc) return rc;
  }
  if (dw) {
    // dw[o, i] = sum_r dy[r, o] x[r, i]: the wider of the two operands is read coalesced
    if (in_f >= out_f) {
      dim3 grid((in_f + 31) / 32, (out_f + 7) / 8);
      launch_pdl(linear_tn_kernel, dim3(grid), dim3(256), 0, st, dy, x, dw, rows, out_f, in_f, 0);
    } else {
      dim3 grid((out_f + 31) / 32, (in_f + 7) / 8);
      launch_pdl(linear_tn_kernel, dim3(grid), dim3(256), 0, st, x, dy, dw, rows, in_f, out_f, 1);
    }
    int rc = check_launch("linear_bwd(dw)");
    if (rc) return rc;
  }
  if (db) {
    launch_pdl(colsum_kernel, dim3((out_f + 31) / 32), dim3(256), 0, st, dy, db, rows, out_f);
    int rc = check_launch("linear_bwd(db)");
    if (rc) return rc;
  }
  return 0;
}
