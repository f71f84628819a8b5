// The path's one data-path collective as ONE kernel over NVLink peer memory: the all-reduce of the per-exit
// statistics sums when the S samples are sharded over the GPUs of a box (the reference is single-device,
// its `np.average(all_output_probs, axis=0)` over ALL passes, results_analyzer.py:247-248, is what the sum restores).
//
// Every rank owns a workspace in symmetric memory (same layout on every rank, mapped into every peer's address space by
// torch.distributed._symmetric_memory):   slot[n_pad] floats | ready[8] u32 | done[8] u32
// Kernel (grid <= SM count, all CTAs resident):
//   0. copy the local sums into the local slot                               (system-scope fence)
//   1. CTA 0: tell every peer "my slot holds epoch e" (st.release.sys into THEIR ready[rank]) and wait until all of
//      them said so in MY ready[]; then release the other CTAs through a local flag
//   2. every thread sums element i of slot_0 .. slot_{W-1} IN RANK ORDER (peer loads over NVLink) and writes the total
//      into the local sums - all ranks add the same numbers in the same order, so the totals are bit-identical everywhere
//   3. last CTA: tell every peer "I have read your slot" and wait for all of them: nobody overwrites a slot that is
//      still being read (the next step's copy happens behind this kernel in stream order)
// Payload: 84 KB per rank at C5 - this is a latency problem (one NVLink round trip per barrier), not a bandwidth one:
// NCCL's all-reduce takes 38 us inside the step's CUDA graph.
// A spin that exceeds `timeout_ns` sets state->err and carries on (garbage result, loud host error) - a missing peer must
// not hang the GPU.
#include "common.cuh"

namespace bnn {

struct PeerState {        // local device memory, zero-initialised once
  uint32_t epoch;         // all-reduces completed so far (written by the last CTA of each one)
  uint32_t go;            // == epoch: the ready barrier has been passed (CTA 0 -> the other CTAs)
  uint32_t arrive[2];     // CTAs done with phase 0 / phase 2
  uint32_t err;           // != 0: a wait timed out
};

struct PeerBases {
  char* base[8];
};

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ uint32_t ld_acquire_gpu(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_gpu(uint32_t* p, uint32_t v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ float4 ld_sys_f4(const float* p) {
  float4 v;
  asm volatile("ld.relaxed.sys.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float ld_sys_f(const float* p) {
  float v;
  asm volatile("ld.relaxed.sys.global.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ uint64_t global_ns() {
  uint64_t t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// spin until *p == want (acquire at `sys` or `gpu` scope); false on timeout
template <bool SYS>
__device__ __forceinline__ bool spin_until(const uint32_t* p, uint32_t want, uint64_t timeout_ns, PeerState* st) {
  const uint64_t t0 = global_ns();
  while ((SYS ? ld_acquire_sys(p) : ld_acquire_gpu(p)) != want) {
    if (global_ns() - t0 > timeout_ns) {
      st->err = 1;
      return false;
    }
  }
  return true;
}

__global__ void __launch_bounds__(256) peer_allreduce_kernel(float* __restrict__ flat, int64_t n, PeerBases pb, int world,
                                                             int rank, int64_t flags_off, PeerState* st,
                                                             uint64_t timeout_ns) {
  // this call's epoch: every CTA reads the counter before the last CTA of THIS launch (phase 3, behind a grid-wide
  // arrival) advances it; the previous launch finished in stream order
  const uint32_t e = ld_acquire_gpu(&st->epoch) + 1u;
  float* my_slot = reinterpret_cast<float*>(pb.base[rank]);
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, nthr = (int64_t)gridDim.x * blockDim.x;
  const int64_t n4 = n >> 2;
  // ---- 0: local sums -> local slot -------------------------------------------------------------------------------------
  for (int64_t i = tid; i < n4; i += nthr)
    reinterpret_cast<float4*>(my_slot)[i] = reinterpret_cast<const float4*>(flat)[i];
  for (int64_t i = (n4 << 2) + tid; i < n; i += nthr) my_slot[i] = flat[i];
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) atomicAdd(&st->arrive[0], 1u);
  // ---- 1: ready barrier over the ranks ---------------------------------------------------------------------------------
  if (blockIdx.x == 0) {
    if (threadIdx.x == 0) {
      spin_until<false>(&st->arrive[0], gridDim.x, timeout_ns, st);
      st->arrive[0] = 0;
      __threadfence_system();
    }
    __syncthreads();
    if ((int)threadIdx.x < world) {
      const int q = threadIdx.x;
      uint32_t* ready_q = reinterpret_cast<uint32_t*>(pb.base[q] + flags_off);           // ready[8] of rank q
      st_release_sys(ready_q + rank, e);
      const uint32_t* ready_me = reinterpret_cast<const uint32_t*>(pb.base[rank] + flags_off);
      spin_until<true>(ready_me + q, e, timeout_ns, st);
    }
    __syncthreads();
    if (threadIdx.x == 0) st_release_gpu(&st->go, e);
  } else {
    if (threadIdx.x == 0) spin_until<false>(&st->go, e, timeout_ns, st);
    __syncthreads();
  }
  // ---- 2: rank-ordered sum of every slot -> local sums --------------------------------------------------------------------
  for (int64_t i = tid; i < n4; i += nthr) {
    float4 acc = ld_sys_f4(reinterpret_cast<const float*>(pb.base[0]) + 4 * i);
    for (int r = 1; r < world; ++r) {
      const float4 v = ld_sys_f4(reinterpret_cast<const float*>(pb.base[r]) + 4 * i);
      acc.x += v.x;
      acc.y += v.y;
      acc.z += v.z;
      acc.w += v.w;
    }
    reinterpret_cast<float4*>(flat)[i] = acc;
  }
  for (int64_t i = (n4 << 2) + tid; i < n; i += nthr) {
    float acc = ld_sys_f(reinterpret_cast<const float*>(pb.base[0]) + i);
    for (int r = 1; r < world; ++r) acc += ld_sys_f(reinterpret_cast<const float*>(pb.base[r]) + i);
    flat[i] = acc;
  }
  // ---- 3: done barrier (the last CTA of this rank speaks for it) -----------------------------------------------------------
  __syncthreads();
  __shared__ uint32_t is_last;
  if (threadIdx.x == 0) {
    __threadfence();
    is_last = atomicAdd(&st->arrive[1], 1u) == gridDim.x - 1 ? 1u : 0u;
  }
  __syncthreads();
  if (is_last) {
    if (threadIdx.x == 0) st->arrive[1] = 0;
    if ((int)threadIdx.x < world) {
      const int q = threadIdx.x;
      uint32_t* done_q = reinterpret_cast<uint32_t*>(pb.base[q] + flags_off) + 8;
      st_release_sys(done_q + rank, e);
      const uint32_t* done_me = reinterpret_cast<const uint32_t*>(pb.base[rank] + flags_off) + 8;
      spin_until<true>(done_me + q, e, timeout_ns, st);
    }
    __syncthreads();
    if (threadIdx.x == 0) st_release_gpu(&st->epoch, e);
  }
}

}  // namespace bnn

using namespace bnn;

extern "C" int bnn_peer_allreduce(float* flat, int64_t n, void* const* peer_base, int world, int rank,
                                  int64_t flags_offset_bytes, void* state, void* stream) {
  if (int rc = check_device()) return rc;
  BNN_REQUIRE(flat && peer_base && state && n > 0, "bnn_peer_allreduce: null pointer / empty payload");
  BNN_REQUIRE(world >= 1 && world <= 8 && rank >= 0 && rank < world, "bnn_peer_allreduce: world=%d rank=%d (1..8 ranks)", world,
              rank);
  BNN_REQUIRE(flags_offset_bytes >= n * (int64_t)sizeof(float) && flags_offset_bytes % 16 == 0,
              "bnn_peer_allreduce: the flags (offset %lld) must lie behind the %lld-float slot, 16-byte aligned",
              (long long)flags_offset_bytes, (long long)n);
  BNN_REQUIRE((reinterpret_cast<uintptr_t>(flat) & 15) == 0, "bnn_peer_allreduce: sums must be 16-byte aligned");
  PeerBases pb{};
  for (int r = 0; r < world; ++r) {
    BNN_REQUIRE(peer_base[r] != nullptr, "bnn_peer_allreduce: peer %d has no mapped workspace", r);
    pb.base[r] = static_cast<char*>(peer_base[r]);
  }
  cudaStream_t st = (cudaStream_t)stream;
  int sms = 0, dev = 0;
  BNN_CUDA_OK(cudaGetDevice(&dev));
  BNN_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  int64_t want = (n / 4 + 255) / 256;
  int grid = (int)(want < 1 ? 1 : (want > sms ? sms : want));      // all CTAs resident: they wait for one another
  peer_allreduce_kernel<<<grid, 256, 0, st>>>(flat, n, pb, world, rank, flags_offset_bytes, static_cast<PeerState*>(state),
                                              2000000000ull /* 2 s */);
  BNN_LAUNCH_OK();
  return BNN_OK;
}

// err flag of the last all-reduces (device -> host, synchronises the stream): 0 = fine
extern "C" int bnn_peer_allreduce_status(const void* state, void* stream) {
  uint32_t host[5] = {0, 0, 0, 0, 0};
  BNN_CUDA_OK(cudaMemcpyAsync(host, state, sizeof(host), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
  BNN_CUDA_OK(cudaStreamSynchronize((cudaStream_t)stream));
  if (host[4] != 0) {
    set_error("bnn_peer_allreduce: a peer did not answer within 2 s (epoch %u)", host[0]);
    return BNN_E_CUDA;
  }
  return BNN_OK;
}
