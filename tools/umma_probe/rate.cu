// Microbenchmark: sustained issue/execute rate of tcgen05.mma kind::tf32 (M=128, K=8) from shared-memory operands.
//   layout 0: no-swizzle K-major (the conv kernels' layout), layout 1: SWIZZLE_128B K-major
//   issue  0: `if (tid == 0)` divergent single-thread loop with 64-bit descriptor adds (conv_tc / conv_wgrad_tc today)
//   issue  1: whole warp converged, descriptors warp-uniform, MMA predicated by elect.sync
// Prints cycles per MMA per CTA (one CTA per SM, all SMs busy).  Results are timing only (operands are zeros).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t mkdesc(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46) |
         ((uint64_t)layout << 61);
}
__device__ __forceinline__ void mma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n" ::"r"(d), "l"(a),
               "l"(b), "r"(idesc), "r"(acc)
               : "memory");
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(pred));
  return pred != 0;
}
template <int ISSUE>
__global__ void __launch_bounds__(128, 1) rate(int N, int layout, int nmma, int ntaps, int step_a, long long* out) {
  extern __shared__ __align__(1024) unsigned char sm[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tb;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 48 * 1024; i += blockDim.x) ((float*)sm)[i] = 0.f;
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(&tb)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tmem = tb;
  const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
  // A: 128 rows (+ tap shifts), B: N rows, both 64 KB apart
  const uint32_t a_base = s32(sm), b_base = s32(sm + 96 * 1024);
  uint64_t ad0, bd0;
  if (layout == 0) { ad0 = mkdesc(a_base, 256 * 16, 128, 0); bd0 = mkdesc(b_base, (uint32_t)N * 16, 128, 0); }
  else { ad0 = mkdesc(a_base, 16, 1024, 2); bd0 = mkdesc(b_base, 16, 1024, 2); }
  long long t0 = 0, t1 = 0;
  if (ISSUE == 0) {
    if (tid == 0) {
      t0 = clock64();
      for (int i = 0; i < nmma; i += ntaps) {
        uint64_t ad = ad0, bd = bd0;
        for (int t = 0; t < ntaps; ++t) {
          mma(tmem, ad, bd, idesc, 1u);
          ad += step_a;
          bd += 2;
        }
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(&bar)) : "memory");
    }
  } else {
    if (warp == 0) {
      t0 = clock64();
      for (int i = 0; i < nmma; i += ntaps) {
        uint64_t ad = ad0, bd = bd0;
        for (int t = 0; t < ntaps; ++t) {
          if (elect_one()) mma(tmem, ad, bd, idesc, 1u);
          ad += step_a;
          bd += 2;
        }
      }
      if (elect_one())
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(&bar)) : "memory");
      __syncwarp();
    }
  }
  asm volatile("{\n.reg .pred p;\nW:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n@p bra DN;\nbra W;\nDN:\n}\n" ::"r"(s32(&bar)) : "memory");
  t1 = clock64();
  if (tid == 0) out[blockIdx.x] = t1 - t0;
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}
int main() {
  long long* out;
  cudaMalloc(&out, 148 * 8);
  const int nmma = 3840, ntaps = 15;
  const size_t smem = 200 * 1024;
  cudaFuncSetAttribute(rate<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaFuncSetAttribute(rate<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const int Ns[6] = {16, 32, 64, 96, 128, 256};
  for (int layout = 0; layout < 2; ++layout)
    for (int issue = 0; issue < 2; ++issue)
      for (int grid : {1, 148})
        for (int ni = 0; ni < 6; ++ni) {
          const int N = Ns[ni];
          const int step_a = layout == 0 ? 2 : 64;   // tap shift: +2 rows (no swizzle) / +8 rows = 1024 B (swizzle 128B)
          for (int rep = 0; rep < 2; ++rep) {
            if (issue == 0) rate<0><<<grid, 128, smem>>>(N, layout, nmma, ntaps, step_a, out);
            else rate<1><<<grid, 128, smem>>>(N, layout, nmma, ntaps, step_a, out);
          }
          cudaError_t e = cudaDeviceSynchronize();
          long long h[148];
          cudaMemcpy(h, out, grid * 8, cudaMemcpyDeviceToHost);
          long long mx = 0, mn = 1LL << 60;
          for (int i = 0; i < grid; ++i) { if (h[i] > mx) mx = h[i]; if (h[i] < mn) mn = h[i]; }
          printf("layout %d issue %d grid %3d N %3d: %s  cycles/MMA min %.1f max %.1f  (floor 128*N/256 = %d)\n", layout, issue, grid, N,
                 cudaGetErrorString(e), (double)mn / nmma, (double)mx / nmma, 128 * N / 256);
        }
  return 0;
}
