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
  const int2* units;        // device table of owned work units {first float4, number of float4 (<= 32)}; NULL: the ranges above
  long nunits;
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

// Work is handed out in UNITS of up to 32 consecutive float4 (one per lane): either cut on the fly from the element ranges of
// the kernel arguments, or read from a device table (the mask-aware path: the table lists only the parameter elements that can
// ever be non-zero, so the always-masked blocks of the skeleton-conv weights -- 22 % of the arena at len64 -- are never streamed
// by the reduce-scatter, the optimiser or the all-gather).  U units per warp and iteration = U peer loads in flight per thread
// (NVLink latency ~2-3 us: a call that runs on few CTAs under the backward pass needs the deeper variant).
template <int U>
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
  const bool table = A.units != nullptr;
  // a peer is missing: apply nothing (the host raises at its next health check)
  const int nsets = s_skip ? 0 : (table ? 1 : A.nranges);
  const float lr_over_bc1 = dyn2[0], inv_sqrt_bc2 = dyn2[1];
  const int lane = threadIdx.x & 31;
  const long gwarp = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long nwarps = ((long)gridDim.x * blockDim.x) >> 5;
  const bool mc = A.mc_grad != nullptr;
  auto upd = [&](float& pp, float gg, float& mm, float& vv) {
    gg = gg * gscale + wd * pp;
    mm = mm + (gg - mm) * omb1;
    vv = vv * beta2 + omb2 * gg * gg;
    const float denom = sqrtf(vv) * inv_sqrt_bc2 + eps;
    pp = pp - lr_over_bc1 * (mm / denom);
  };
  const float4* __restrict__ prm = reinterpret_cast<const float4*>(A.param[R]);
  float4* __restrict__ m4 = reinterpret_cast<float4*>(m);
  float4* __restrict__ v4 = reinterpret_cast<float4*>(v);
  for (int r = 0; r < nsets; ++r) {
    const long b4 = table ? 0 : (A.beg[r] >> 2), e4 = table ? 0 : (A.end[r] >> 2);
    const long nu = table ? A.nunits : ((e4 - b4 + 31) >> 5);
    for (long u0 = gwarp; u0 < nu; u0 += (long)U * nwarps) {
      long idx[U];
      bool ok[U];
#pragma unroll
      for (int k = 0; k < U; ++k) {
        const long u = u0 + (long)k * nwarps;
        ok[k] = u < nu;
        idx[k] = 0;
        if (ok[k]) {
          if (table) {
            const int2 e = __ldg(A.units + u);
            idx[k] = (long)e.x + lane;
            ok[k] = lane < e.y;
          } else {
            idx[k] = b4 + (u << 5) + lane;
            ok[k] = idx[k] < e4;
          }
        }
      }
      float4 G[U], P[U], M[U], V[U];
      if (mc) {
        // NVSwitch path: the reduce-scatter is one multimem.ld_reduce per element, the all-gather one multimem.st: every rank
        // moves N/W elements each way instead of N(W-1)/W.
#pragma unroll
        for (int k = 0; k < U; ++k)
          if (ok[k]) G[k] = multimem_ld_reduce_add(A.mc_grad + 4 * idx[k]);
      } else {
#pragma unroll
        for (int k = 0; k < U; ++k)
          if (ok[k]) G[k] = __ldcg(reinterpret_cast<const float4*>(A.grad[0]) + idx[k]);
      }
#pragma unroll
      for (int k = 0; k < U; ++k)
        if (ok[k]) {
          P[k] = prm[idx[k]];
          M[k] = m4[idx[k]];
          V[k] = v4[idx[k]];
        }
      if (!mc) {
        for (int q = 1; q < W; ++q) {                  // fixed order 0..W-1: deterministic; U peer loads in flight per step
          float4 t[U];
#pragma unroll
          for (int k = 0; k < U; ++k)
            if (ok[k]) t[k] = __ldcg(reinterpret_cast<const float4*>(A.grad[q]) + idx[k]);
#pragma unroll
          for (int k = 0; k < U; ++k)
            if (ok[k]) { G[k].x += t[k].x; G[k].y += t[k].y; G[k].z += t[k].z; G[k].w += t[k].w; }
        }
      }
#pragma unroll
      for (int k = 0; k < U; ++k)
        if (ok[k]) {
          upd(P[k].x, G[k].x, M[k].x, V[k].x); upd(P[k].y, G[k].y, M[k].y, V[k].y);
          upd(P[k].z, G[k].z, M[k].z, V[k].z); upd(P[k].w, G[k].w, M[k].w, V[k].w);
          m4[idx[k]] = M[k];
          v4[idx[k]] = V[k];
          if (mc) {
            multimem_st(A.mc_param + 4 * idx[k], P[k]);
          } else {
            for (int q = 0; q < W; ++q) reinterpret_cast<float4*>(A.param[q])[idx[k]] = P[k];
          }
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

static int dp_launch(const hmvae_dp_peers* peers, float* m, float* v, const long* ranges, int nranges, const int* units,
                     long nunits, const float* dyn2, double beta1, double beta2, float eps, float weight_decay, float grad_scale,
                     unsigned int* state, int max_ctas, int in_flight, void* stream) {
  if (!peers || !m || !v || !dyn2 || !state) return fail_arg("dp_adam_step: null pointer");
  if (peers->world < 1 || peers->world > HMVAE_DP_MAX_WORLD || peers->rank < 0 || peers->rank >= peers->world)
    return fail_arg("dp_adam_step: bad world / rank");
  const float omb1 = (float)(1.0 - beta1), omb2 = (float)(1.0 - beta2);      // 1 - beta rounded ONCE from double, like torch
  DpArgs A;
  memset(&A, 0, sizeof(A));
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
  long work_units = 0;            // 32-float4 units of this call
  if (units) {
    if (nunits < 0 || (reinterpret_cast<uintptr_t>(units) & 7)) return fail_arg("dp_adam_step_units: bad unit table");
    A.units = nunits > 0 ? reinterpret_cast<const int2*>(units) : nullptr;
    A.nunits = nunits;
    A.nranges = 0;
    work_units = nunits;
  } else {
    if (nranges < 0 || nranges > HMVAE_DP_MAX_RANGES) return fail_arg("dp_adam_step: too many ranges");
    A.nranges = nranges;
    for (int r = 0; r < nranges; ++r) {
      A.beg[r] = ranges[2 * r];
      A.end[r] = ranges[2 * r + 1];
      if (A.beg[r] < 0 || A.end[r] < A.beg[r] || (A.beg[r] & 3) || (A.end[r] & 3)) return fail_arg("dp_adam_step: ranges must be multiples of 4");
      work_units += ((A.end[r] - A.beg[r]) / 4 + 31) / 32;
    }
  }
  // every rank must launch (the flag barriers pair up) even if it owns nothing; all CTAs are resident (<= 4 per SM)
  long blocks = (work_units + 7) / 8;
  const long cap = (long)num_sms() * 4;
  if (blocks > cap) blocks = cap;
  if (max_ctas > 0 && blocks > max_ctas) blocks = max_ctas;      // a call that runs under other kernels leaves them room
  if (blocks < 1) blocks = 1;
  // loads in flight per thread: 1 on one rank (local memory), 2 across ranks, more on request (CTA-capped calls)
  int U = in_flight > 0 ? in_flight : env_int(A.world > 1 ? "HMVAE_DP_IN_FLIGHT" : "HMVAE_DP_IN_FLIGHT_LOCAL", A.world > 1 ? 2 : 1);
  const dim3 g((unsigned)blocks), b(256);
  cudaStream_t st = (cudaStream_t)stream;
  if (U >= 8)
    launch_pdl(dp_adam_kernel<8>, g, b, 0, st, A, m, v, dyn2, omb1, (float)beta2, omb2, eps, weight_decay, grad_scale, state);
  else if (U >= 4)
    launch_pdl(dp_adam_kernel<4>, g, b, 0, st, A, m, v, dyn2, omb1, (float)beta2, omb2, eps, weight_decay, grad_scale, state);
  else if (U >= 2)
    launch_pdl(dp_adam_kernel<2>, g, b, 0, st, A, m, v, dyn2, omb1, (float)beta2, omb2, eps, weight_decay, grad_scale, state);
  else
    launch_pdl(dp_adam_kernel<1>, g, b, 0, st, A, m, v, dyn2, omb1, (float)beta2, omb2, eps, weight_decay, grad_scale, state);
  return check_launch("dp_adam_step");
}

extern "C" int hmvae_dp_adam_step(const hmvae_dp_peers* peers, float* m, float* v, const long* ranges, int nranges,
                                  const float* dyn2, double beta1, double beta2, float eps, float weight_decay,
                                  float grad_scale, unsigned int* state, int max_ctas, void* stream) {
  if (nranges > 0 && !ranges) return fail_arg("dp_adam_step: null pointer");
  return dp_launch(peers, m, v, ranges, nranges, nullptr, 0, dyn2, beta1, beta2, eps, weight_decay, grad_scale, state, max_ctas, 0,
                   stream);
}

extern "C" int hmvae_dp_adam_step_units(const hmvae_dp_peers* peers, float* m, float* v, const int* units, long nunits,
                                        const float* dyn2, double beta1, double beta2, float eps, float weight_decay,
                                        float grad_scale, unsigned int* state, int max_ctas, int loads_in_flight, void* stream) {
  if (!units) return fail_arg("dp_adam_step_units: null unit table");
  return dp_launch(peers, m, v, nullptr, 0, units, nunits, dyn2, beta1, beta2, eps, weight_decay, grad_scale, state, max_ctas,
                   loads_in_flight, stream);
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
