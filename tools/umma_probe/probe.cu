// Probe: one tcgen05.mma kind::tf32 with MN-major A and/or B (no swizzle), D[128 x N] in TMEM, checked on the host.
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>
#include <vector>
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46);
}
// mode bit0: A MN-major, bit1: B MN-major.   KS k-steps of 8.
__global__ void probe(const float* A, const float* B, float* D, int N, int KS, int mode, int lboA, int sboA, int lboB, int sboB,
                      int abytes, int bbytes, int advA, int advB) {
  extern __shared__ __align__(1024) unsigned char sm[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tb;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < abytes / 4; i += blockDim.x) ((float*)sm)[i] = A[i];
  for (int i = tid; i < bbytes / 4; i += blockDim.x) ((float*)(sm + abytes))[i] = B[i];
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(&tb)), "r"(64));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tmem = tb;
  if (tid == 0) {
    uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
    if (mode & 1) idesc |= 1u << 15;
    if (mode & 2) idesc |= 1u << 16;
    uint64_t ad = desc(s32(sm), lboA, sboA), bd = desc(s32(sm + abytes), lboB, sboB);
    for (int ks = 0; ks < KS; ++ks) {
      uint32_t acc = ks > 0;
      asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem),
                   "l"(ad), "l"(bd), "r"(idesc), "r"(acc)
                   : "memory");
      // advance along K: K-major: 2 chunks of 16B = 2*LBO ; MN-major: 8 rows = one LBO group
      ad += advA;
      bd += advB;
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(&bar)) : "memory");
  }
  // everyone waits
  asm volatile("{\n.reg .pred p;\nW:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n@p bra DN;\nbra W;\nDN:\n}\n" ::"r"(s32(&bar)) : "memory");
  asm volatile("tcgen05.fence::after_thread_sync;");
  if (warp < 4) {
    for (int c = 0; c < N; c += 16) {
      uint32_t r[16];
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                   : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                     "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                   : "r"(tmem + ((uint32_t)(warp * 32) << 16) + c));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      for (int i = 0; i < 16; ++i) D[(warp * 32 + lane) * N + c + i] = __uint_as_float(r[i]);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(64));
}

int main() {
  const int M = 128, N = 32, K = 32, KS = K / 8;
  std::vector<float> a(M * K), b(N * K), ref(M * N, 0.f);
  for (auto& v : a) v = (float)((rand() % 17) - 8) / 8.f;     // exactly representable in tf32
  for (auto& v : b) v = (float)((rand() % 17) - 8) / 8.f;
  for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) { float s = 0; for (int k = 0; k < K; ++k) s += a[m * K + k] * b[n * K + k]; ref[m * N + n] = s; }
  for (int mode = 0; mode < 4; ++mode) {
    // A smem image.  K-major: [kchunk (K/4)][row m][4 k]  LBO = M*16, SBO = 128.   MN-major: [mchunk (M/4)][k row][4 m]  SBO = K*16, LBO = 128
    std::vector<float> as(M * K), bs(N * K);
    int lboA, sboA, lboB, sboB;
    if (!(mode & 1)) { for (int m = 0; m < M; ++m) for (int k = 0; k < K; ++k) as[((k / 4) * M + m) * 4 + k % 4] = a[m * K + k]; lboA = M * 16; sboA = 128; }
    else { for (int m = 0; m < M; ++m) for (int k = 0; k < K; ++k) as[((m / 4) * K + k) * 4 + m % 4] = a[m * K + k]; sboA = K * 16; lboA = 128; }
    if (!(mode & 2)) { for (int n = 0; n < N; ++n) for (int k = 0; k < K; ++k) bs[((k / 4) * N + n) * 4 + k % 4] = b[n * K + k]; lboB = N * 16; sboB = 128; }
    else { for (int n = 0; n < N; ++n) for (int k = 0; k < K; ++k) bs[((n / 4) * K + k) * 4 + n % 4] = b[n * K + k]; sboB = K * 16; lboB = 128; }
    const int advA = (mode & 1) ? 8 : 2 * M, advB = (mode & 2) ? 8 : 2 * N;   // 16-byte units per 8 reduction elements
    float *dA, *dB, *dD;
    cudaMalloc(&dA, as.size() * 4); cudaMalloc(&dB, bs.size() * 4); cudaMalloc(&dD, M * N * 4);
    cudaMemcpy(dA, as.data(), as.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, bs.data(), bs.size() * 4, cudaMemcpyHostToDevice);
    cudaMemset(dD, 0, M * N * 4);
    probe<<<1, 128, as.size() * 4 + bs.size() * 4 + 1024>>>(dA, dB, dD, N, KS, mode, lboA, sboA, lboB, sboB, (int)as.size() * 4, (int)bs.size() * 4, advA, advB);
    cudaError_t e = cudaDeviceSynchronize();
    std::vector<float> d(M * N);
    cudaMemcpy(d.data(), dD, M * N * 4, cudaMemcpyDeviceToHost);
    double err = 0, nz = 0;
    for (int i = 0; i < M * N; ++i) { err += fabs(d[i] - ref[i]); nz += d[i] != 0; }
    printf("mode %d (A %s, B %s): cuda=%s sum|err|=%g nonzero=%g d[0..3]=%g %g %g %g ref=%g %g %g %g\n", mode, (mode & 1) ? "MN" : "K",
           (mode & 2) ? "MN" : "K", cudaGetErrorString(e), err, nz, d[0], d[1], d[2], d[3], ref[0], ref[1], ref[2], ref[3]);
    // swapped LBO/SBO interpretation for the MN-major operands
    if (mode) {
      int la = lboA, sa = sboA, lb = lboB, sb = sboB;
      if (mode & 1) { la = sboA; sa = lboA; }
      if (mode & 2) { lb = sboB; sb = lboB; }
      cudaMemset(dD, 0, M * N * 4);
      probe<<<1, 128, as.size() * 4 + bs.size() * 4 + 1024>>>(dA, dB, dD, N, KS, mode, la, sa, lb, sb, (int)as.size() * 4, (int)bs.size() * 4, advA, advB);
      e = cudaDeviceSynchronize();
      cudaMemcpy(d.data(), dD, M * N * 4, cudaMemcpyDeviceToHost);
      err = 0;
      for (int i = 0; i < M * N; ++i) err += fabs(d[i] - ref[i]);
      printf("   swapped LBO/SBO fields: cuda=%s sum|err|=%g\n", cudaGetErrorString(e), err);
    }
    cudaFree(dA); cudaFree(dB); cudaFree(dD);
  }
  return 0;
}
