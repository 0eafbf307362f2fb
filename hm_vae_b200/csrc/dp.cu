// Data-parallel optimiser step fused with its collective, over NVLink peer memory.   sm_100a.
//
// Replaces  [NCCL all-reduce of 52.9 MB of gradients] -> [Adam over 52.9 MB x (p, g, m, v)]  (DataParallel's gather +
// torch.optim.Adam in the reference, train_motion_vae.py:49-53 / trainer_motion_vae.py:29-31, 92-93) by ONE kernel per rank:
//
//   rank r owns the r-th share of the (live) parameter elements.  For each owned element it
//     1. reads the gradient from EVERY rank's gradient arena (peer loads through NVLink; fixed order 0..W-1 => every element
//        is reduced exactly once, by one rank, deterministically)                               -- the reduce-scatter,
//     2. applies torch.optim.Adam (L2 decay added to the gradient) with its LOCAL slice of m and v (optimiser state and
//        optimiser arithmetic are sharded W ways)                                                -- the optimiser,
//     3. stores the new parameter value into EVERY rank's parameter arena (peer stores)          -- the all-gather.
//
//   Cross-rank ordering uses two flag barriers in peer memory (monotonic epoch numbers, release/acquire at system scope):
//     entry: "my gradients are final"  -- rank q's stream order guarantees that when its kernel starts;
//     exit : "I have finished reading your gradients and writing your parameters" -- the LAST CTA of each rank signals and
//            waits, so a rank's kernel (hence its next forward pass / its next backward pass overwriting the gradient arena)
//            cannot complete before all its peers are done with its memory.
//   No NCCL call, no host synchronisation: the kernel is CUDA-graph capturable (the epoch lives in device memory).
//   A waiter that sees no signal for HMVAE_DP_TIMEOUT_S seconds (default 120) raises state[2] and its CTA SKIPS the update: no
//   stale or half-written peer gradient is ever applied, the GPU never hangs, and the host raises at its next health check
//   (Trainer._check_health: every host sync, save(), every 500 steps).
#include <string.h>

#include "common.cuh"

namespace hmvae {


struct DpArgs {
  int world, rank;
  const float* grad[HMVAE_DP_MAX_WORLD];
  float* param[HMVAE_DP_MAX_WORLD];
  unsigned int* flags[HMVAE_DP_MAX_WORLD];      // [2 * world] per rank: entry flags, exit flags (indexed by the SIGNALLING rank)
  long beg[HMVAE_DP_MAX_RANGES], end[HMVAE_DP_MAX_RANGES];   // owned element ranges (multiples of 4, 16-byte aligned)
  int nranges;
  const float* mc_grad;     // NVSwitch multicast mappings of the two arenas (NULL: unicast peer loads / stores)
  float* mc_param;
  unsigned long long timeout_ns;
};

// NVLS: one load returns the sum over all ranks' copies (reduced inside the switch), one store lands in every rank's copy
__device__ __forceinline__ float4 multimem_ld_reduce_add(const float* p) {
  float4 v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(p)
               : "memory");
  return v;
}
__device__ __forceinline__ void multimem_st(float* p, const float4& v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

__device__ __forceinline__ void st_release_sys(unsigned int* p, unsigned int v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
// waits until *flag >= epoch (wrap-safe signed difference); false on timeout
__device__ __forceinline__ bool wait_flag(const unsigned int* flag, unsigned int epoch, unsigned long long timeout_ns) {
  const unsigned long long t0 = globaltimer_ns();
  unsigned int spins = 0;
  while ((int)(ld_acquire_sys(flag) - epoch) < 0) {
    if ((++spins & 1023u) == 0 && globaltimer_ns() - t0 > timeout_ns) return false;
    __nanosleep(64);
  }
  return true;
}

template <bool TWO>      // TWO: two float4 per thread and iteration (multi-rank: more peer loads in flight; costs registers)
__global__ void __launch_bounds__(256) dp_adam_kernel(DpArgs A, float* __restrict__ m, float* __restrict__ v,
                                                      const float* __restrict__ dyn2, float omb1, float beta2, float omb2, float eps,
                                                      float wd, float gscale, unsigned int* __restrict__ state) {
  pdl_trigger();
  pdl_wait();
  // state[0] = epoch of the last completed call, state[1] = CTAs done in this call, state[2] = timeout flag
  __shared__ unsigned int s_epoch, s_skip;
  const int W = A.world, R = A.rank;
  if (threadIdx.x == 0) {
    const unsigned int epoch = state[0] + 1u;
    unsigned int skip = 0;
    if (W > 1) {
      if (blockIdx.x == 0)
        for (int q = 0; q < W; ++q) st_release_sys(A.flags[q] + R, epoch);                  // my gradients are final
      for (int q = 0; q < W; ++q)
        if (!wait_flag(A.flags[R] + q, epoch, A.timeout_ns)) {
          atomicExch(state + 2, 1u);
          skip = 1;
        }
    }
    s_epoch = epoch;
    s_skip = skip;
  }
  __syncthreads();
  const int nranges = s_skip ? 0 : A.nranges;      // a peer is missing: apply nothing (the host raises at its next health check)
  const float lr_over_bc1 = dyn2[0], inv_sqrt_bc2 = dyn2[1];
  const long tid = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const long nthreads = (long)gridDim.x * blockDim.x;
  auto upd = [&](float& pp, float gg, float& mm, float& vv) {
    gg = gg * gscale + wd * pp;
    mm = mm + (gg - mm) * omb1;
    vv = vv * beta2 + omb2 * gg * gg;
    const float denom = sqrtf(vv) * inv_sqrt_bc2 + eps;
    pp = pp - lr_over_bc1 * (mm / denom);
  };
  for (int r = 0; r < nranges; ++r) {
    const long b4 = A.beg[r] >> 2, e4 = A.end[r] >> 2;
    if (A.mc_grad != nullptr) {
      // NVSwitch path: the reduce-scatter is one multimem.ld_reduce per element, the all-gather one multimem.st: every rank
      // moves N/W elements each way instead of N(W-1)/W.
      for (long i = b4 + tid; i < e4; i += 2 * nthreads) {           // two in-switch reductions in flight per thread
        const long i2 = i + nthreads;
        const bool two = i2 < e4;
        const float4 G = multimem_ld_reduce_add(A.mc_grad + 4 * i);
        float4 G2 = G;
        if (two) G2 = multimem_ld_reduce_add(A.mc_grad + 4 * i2);
        float4 P = reinterpret_cast<const float4*>(A.param[R])[i];
        float4 M = reinterpret_cast<float4*>(m)[i], V = reinterpret_cast<float4*>(v)[i];
        float4 P2 = P, M2 = M, V2 = V;
        if (two) {
          P2 = reinterpret_cast<const float4*>(A.param[R])[i2];
          M2 = reinterpret_cast<float4*>(m)[i2];
          V2 = reinterpret_cast<float4*>(v)[i2];
        }
        upd(P.x, G.x, M.x, V.x); upd(P.y, G.y, M.y, V.y); upd(P.z, G.z, M.z, V.z); upd(P.w, G.w, M.w, V.w);
        reinterpret_cast<float4*>(m)[i] = M;
        reinterpret_cast<float4*>(v)[i] = V;
        multimem_st(A.mc_param + 4 * i, P);
        if (two) {
          upd(P2.x, G2.x, M2.x, V2.x); upd(P2.y, G2.y, M2.y, V2.y); upd(P2.z, G2.z, M2.z, V2.z); upd(P2.w, G2.w, M2.w, V2.w);
          reinterpret_cast<float4*>(m)[i2] = M2;
          reinterpret_cast<float4*>(v)[i2] = V2;
          multimem_st(A.mc_param + 4 * i2, P2);
        }
      }
      continue;
    }
    if constexpr (!TWO) {
      for (long i = b4 + tid; i < e4; i += nthreads) {
        float4 G = __ldcg(reinterpret_cast<const float4*>(A.grad[0]) + i);
        for (int q = 1; q < W; ++q) {
          const float4 g2 = __ldcg(reinterpret_cast<const float4*>(A.grad[q]) + i);
          G.x += g2.x; G.y += g2.y; G.z += g2.z; G.w += g2.w;
        }
        float4 P = reinterpret_cast<const float4*>(A.param[R])[i];
        float4 M = reinterpret_cast<float4*>(m)[i], V = reinterpret_cast<float4*>(v)[i];
        upd(P.x, G.x, M.x, V.x); upd(P.y, G.y, M.y, V.y); upd(P.z, G.z, M.z, V.z); upd(P.w, G.w, M.w, V.w);
        reinterpret_cast<float4*>(m)[i] = M;
        reinterpret_cast<float4*>(v)[i] = V;
        for (int q = 0; q < W; ++q) reinterpret_cast<float4*>(A.param[q])[i] = P;
      }
      continue;
    }
    // two float4 per thread and iteration: 2 * W peer loads in flight before the first add (NVLink latency ~2-3 us)
    for (long i = b4 + tid; i < e4; i += (TWO ? 2 : 1) * nthreads) {
      const long i2 = i + nthreads;
      const bool two = TWO && i2 < e4;
      float4 Ga[HMVAE_DP_MAX_WORLD], Gb[HMVAE_DP_MAX_WORLD];
#pragma unroll
      for (int q = 0; q < HMVAE_DP_MAX_WORLD; ++q) {
        if (q < W) {
          Ga[q] = __ldcg(reinterpret_cast<const float4*>(A.grad[q]) + i);
          if (two) Gb[q] = __ldcg(reinterpret_cast<const float4*>(A.grad[q]) + i2);
        }
      }
      float4 P = reinterpret_cast<const float4*>(A.param[R])[i];
      float4 M = reinterpret_cast<float4*>(m)[i], V = reinterpret_cast<float4*>(v)[i];
      float4 P2 = P, M2 = M, V2 = V;
      if (two) {
        P2 = reinterpret_cast<const float4*>(A.param[R])[i2];
        M2 = reinterpret_cast<float4*>(m)[i2];
        V2 = reinterpret_cast<float4*>(v)[i2];
      }
      float4 G = Ga[0], G2 = Gb[0];
#pragma unroll
      for (int q = 1; q < HMVAE_DP_MAX_WORLD; ++q) {
        if (q < W) {                               // fixed order 0..W-1: deterministic
          G.x += Ga[q].x; G.y += Ga[q].y; G.z += Ga[q].z; G.w += Ga[q].w;
          if (two) { G2.x += Gb[q].x; G2.y += Gb[q].y; G2.z += Gb[q].z; G2.w += Gb[q].w; }
        }
      }
      upd(P.x, G.x, M.x, V.x); upd(P.y, G.y, M.y, V.y); upd(P.z, G.z, M.z, V.z); upd(P.w, G.w, M.w, V.w);
      reinterpret_cast<float4*>(m)[i] = M;
      reinterpret_cast<float4*>(v)[i] = V;
      for (int q = 0; q < W; ++q) reinterpret_cast<float4*>(A.param[q])[i] = P;
      if (two) {
        upd(P2.x, G2.x, M2.x, V2.x); upd(P2.y, G2.y, M2.y, V2.y); upd(P2.z, G2.z, M2.z, V2.z); upd(P2.w, G2.w, M2.w, V2.w);
        reinterpret_cast<float4*>(m)[i2] = M2;
        reinterpret_cast<float4*>(v)[i2] = V2;
        for (int q = 0; q < W; ++q) reinterpret_cast<float4*>(A.param[q])[i2] = P2;
      }
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence_system();
    const unsigned int done = atomicAdd(state + 1, 1u);
    if (done == gridDim.x - 1) {                       // last CTA of this rank
      __threadfence_system();
      const unsigned int epoch = s_epoch;
      if (W > 1) {
        for (int q = 0; q < W; ++q) st_release_sys(A.flags[q] + W + R, epoch);               // done with your memory
        for (int q = 0; q < W; ++q)
          if (!wait_flag(A.flags[R] + W + q, epoch, A.timeout_ns)) atomicExch(state + 2, 1u);
      }
      state[1] = 0;
      state[0] = epoch;
      __threadfence();
    }
  }
}

}  // namespace hmvae

using namespace hmvae;

extern "C" int hmvae_dp_adam_step(const hmvae_dp_peers* peers, float* m, float* v, const long* ranges, int nranges,
                                  const float* dyn2, double beta1, double beta2, float eps, float weight_decay,
                                  float grad_scale, unsigned int* state, int max_ctas, void* stream) {
  if (!peers || !m || !v || !dyn2 || !state || (nranges > 0 && !ranges)) return fail_arg("dp_adam_step: null pointer");
  if (peers->world < 1 || peers->world > HMVAE_DP_MAX_WORLD || peers->rank < 0 || peers->rank >= peers->world)
    return fail_arg("dp_adam_step: bad world / rank");
  if (nranges < 0 || nranges > HMVAE_DP_MAX_RANGES) return fail_arg("dp_adam_step: too many ranges");
  const float omb1 = (float)(1.0 - beta1), omb2 = (float)(1.0 - beta2);      // 1 - beta rounded ONCE from double, like torch
  DpArgs A;
  A.world = peers->world;
  A.rank = peers->rank;
  for (int q = 0; q < HMVAE_DP_MAX_WORLD; ++q) {
    const bool on = q < peers->world;
    A.grad[q] = on ? peers->grad[q] : nullptr;
    A.param[q] = on ? peers->param[q] : nullptr;
    A.flags[q] = on ? peers->flags[q] : nullptr;
    if (on && (!A.grad[q] || !A.param[q] || (peers->world > 1 && !A.flags[q]))) return fail_arg("dp_adam_step: null peer pointer");
    if (on && (!aligned16(A.grad[q]) || !aligned16(A.param[q]))) return fail_arg("dp_adam_step: arenas must be 16-byte aligned");
  }
  if (!aligned16(m) || !aligned16(v)) return fail_arg("dp_adam_step: m / v must be 16-byte aligned");
  A.mc_grad = peers->world > 1 ? peers->mc_grad : nullptr;
  A.mc_param = peers->world > 1 ? peers->mc_param : nullptr;
  if ((A.mc_grad == nullptr) != (A.mc_param == nullptr)) return fail_arg("dp_adam_step: both or neither multicast pointer");
  if (A.mc_grad && (!aligned16(A.mc_grad) || !aligned16(A.mc_param))) return fail_arg("dp_adam_step: multicast pointers must be 16-byte aligned");
  {
    int secs = env_int("HMVAE_DP_TIMEOUT_S", 120);
    if (secs < 1) secs = 1;
    A.timeout_ns = (unsigned long long)secs * 1000000000ull;
  }
  long total = 0;
  A.nranges = nranges;
  for (int r = 0; r < nranges; ++r) {
    A.beg[r] = ranges[2 * r];
    A.end[r] = ranges[2 * r + 1];
    if (A.beg[r] < 0 || A.end[r] < A.beg[r] || (A.beg[r] & 3) || (A.end[r] & 3)) return fail_arg("dp_adam_step: ranges must be multiples of 4");
    total += A.end[r] - A.beg[r];
  }
  // every rank must launch (the flag barriers pair up) even if it owns nothing; all CTAs are resident (<= 4 per SM)
  long blocks = (total / 4 + 255) / 256;
  const long cap = (long)num_sms() * 4;
  if (blocks > cap) blocks = cap;
  if (max_ctas > 0 && blocks > max_ctas) blocks = max_ctas;      // a call that runs under other kernels leaves them room
  if (blocks < 1) blocks = 1;
  if (A.world > 1 && A.mc_grad == nullptr)
    launch_pdl(dp_adam_kernel<true>, dim3((unsigned)blocks), dim3(256), 0, (cudaStream_t)stream, A, m, v, dyn2, omb1, (float)beta2, omb2, eps, weight_decay, grad_scale, state);
  else
    launch_pdl(dp_adam_kernel<false>, dim3((unsigned)blocks), dim3(256), 0, (cudaStream_t)stream, A, m, v, dyn2, omb1, (float)beta2, omb2, eps, weight_decay, grad_scale, state);
  return check_launch("dp_adam_step");
}

// ---------------------------------------------------------------- peer memory plumbing (CUDA IPC), used when
// torch.distributed._symmetric_memory is not usable in the container
extern "C" int hmvae_ipc_alloc(long bytes, void** ptr) {
  if (!ptr || bytes <= 0) return fail_arg("ipc_alloc: bad arguments");
  HMVAE_CUDA(cudaMalloc(ptr, (size_t)bytes));
  HMVAE_CUDA(cudaMemset(*ptr, 0, (size_t)bytes));
  HMVAE_CUDA(cudaDeviceSynchronize());
  return 0;
}
extern "C" int hmvae_ipc_free(void* ptr) {
  if (ptr) HMVAE_CUDA(cudaFree(ptr));
  return 0;
}
extern "C" int hmvae_ipc_get_handle(void* ptr, unsigned char* handle64) {
  if (!ptr || !handle64) return fail_arg("ipc_get_handle: null pointer");
  cudaIpcMemHandle_t h;
  HMVAE_CUDA(cudaIpcGetMemHandle(&h, ptr));
  static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t is 64 bytes");
  memcpy(handle64, &h, 64);
  return 0;
}
extern "C" int hmvae_ipc_open_handle(const unsigned char* handle64, void** ptr) {
  if (!handle64 || !ptr) return fail_arg("ipc_open_handle: null pointer");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  HMVAE_CUDA(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return 0;
}
extern "C" int hmvae_ipc_close_handle(void* ptr) {
  if (ptr) HMVAE_CUDA(cudaIpcCloseMemHandle(ptr));
  return 0;
}
