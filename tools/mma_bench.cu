// Microbenchmark: cycles per tcgen05.mma (kind::f16, bf16 operands in shared memory, fp32 accumulator in TMEM)
// as a function of N, the swizzle width of the operand rows, and cta_group (1 CTA: M = 128; pair: M = 256).
// Every SM (or pair) issues a chain of MMAs that cycle through a few shared-memory operand buffers; the
// operands are never written (content is irrelevant to the timing), so this is the ISSUE/FETCH/MATH rate of
// the instruction without TMA traffic beside it.
// Build:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I include -o tools/mma_bench tools/mma_bench.cu
// Run (B200):  tools/mma_bench
#include <cstdio>
#include <vector>

#include "../interpreting_video_features_b200/csrc/conv_common.cuh"

using namespace ivf_tc;

void ivf_set_error(const char*, ...) {}

template <int NCTA>
__global__ void __launch_bounds__(256) mma_bench_kernel(int n, int rowb, int layout, int iters, int nbuf, int ksteps,
                                                        int nacc, int nissue, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t done_bar;
  __shared__ long long t_issue[8], t_done[8];
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rank = NCTA == 2 ? (int)cluster_ctarank() : 0;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t a_bytes = 128u * rowb;
  const uint32_t b_bytes = (((uint32_t)(n / NCTA) * rowb) + 1023u) & ~1023u;
  if (threadIdx.x == 0) {
    mbar_init(&done_bar, nissue);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    __syncwarp();
    if constexpr (NCTA == 2) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512u) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512u) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  // zero the operands (denormal/NaN patterns must not matter, but keep it clean)
  for (uint32_t i = threadIdx.x; i < (uint32_t)nbuf * (a_bytes + b_bytes) / 16; i += blockDim.x)
    reinterpret_cast<uint4*>(smem_raw + (base - smem_u32(smem_raw)))[i] = make_uint4(0, 0, 0, 0);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if constexpr (NCTA == 2) cluster_sync_all();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_slot;
  if (warp >= 1 && warp <= nissue && rank == 0) {
    const bool leader = elect_one();
    const uint32_t idesc = make_idesc_bf16(128 * NCTA, n);
    const uint32_t desc_hi = smem_desc_hi(8u * rowb, (uint32_t)layout);
    long long t0 = clock64();
    int buf = 0, acc = 0;
    for (int it = 0; it < iters; ++it) {
      // rotate over nacc accumulators (no back-to-back dependency); each issuing warp has its own set
      const uint32_t d = tmem + (uint32_t)(((warp - 1) * nacc + acc) * n);
      const uint32_t a_lo = smem_desc_lo(base + buf * a_bytes);
      const uint32_t b_lo = smem_desc_lo(base + nbuf * a_bytes + buf * b_bytes);
      if (leader) {
        for (int k = 0; k < ksteps; ++k) {
          if constexpr (NCTA == 2) umma_bf16_lo_pair(d, a_lo + 2u * k, b_lo + 2u * k, desc_hi, idesc, 1u);
          else umma_bf16_lo(d, a_lo + 2u * k, b_lo + 2u * k, desc_hi, idesc, 1u);
        }
      }
      __syncwarp();
      if (++buf == nbuf) buf = 0;
      if (++acc == nacc) acc = 0;
    }
    long long t1 = clock64();
    if (leader) {
      if constexpr (NCTA == 2) umma_commit_pair(&done_bar); else umma_commit(&done_bar);
    }
    __syncwarp();
    mbar_wait(&done_bar, 0);
    long long t2 = clock64();
    if (lane == 0) {
      t_issue[warp - 1] = t1 - t0;
      t_done[warp - 1] = t2 - t0;
    }
  } else if (NCTA == 2 && warp == 1 && rank == 1) {
    mbar_wait(&done_bar, 0);  // the commit is multicast to both CTAs
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x == 0 && rank == 0) {
    long long a = 0, b = 0;
    for (int w = 0; w < nissue; ++w) {
      a = max(a, t_issue[w]);
      b = max(b, t_done[w]);
    }
    out[(blockIdx.x / NCTA) * 2] = a;
    out[(blockIdx.x / NCTA) * 2 + 1] = b;
  }
  if constexpr (NCTA == 2) cluster_sync_all();
  if (warp == 0) {
    __syncwarp();
    if constexpr (NCTA == 2)
      asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
    else
      asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
  }
}

template <int NCTA>
double run(int n, int kch, int iters, int nbuf, int sms, int nacc, int nissue) {
  const int rowb = kch * 2, layout = kch == 64 ? 2 : 4, ksteps = kch / 16;
  const size_t a_bytes = 128 * rowb, b_bytes = ((size_t)(n / NCTA) * rowb + 1023) & ~(size_t)1023;
  const size_t smem = nbuf * (a_bytes + b_bytes) + 1024;
  cudaFuncSetAttribute(mma_bench_kernel<NCTA>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  long long* out;
  cudaMalloc(&out, sizeof(long long) * 2 * sms);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(sms);
  cfg.blockDim = dim3(256);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = NCTA;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  double best = 1e30;
  for (int rep = 0; rep < 3; ++rep) {
    cudaError_t e = cudaLaunchKernelEx(&cfg, mma_bench_kernel<NCTA>, n, rowb, layout, iters, nbuf, ksteps, nacc, nissue, out);
    if (e != cudaSuccess || (e = cudaDeviceSynchronize()) != cudaSuccess) {
      printf("launch failed: %s\n", cudaGetErrorString(e));
      return -1;
    }
    std::vector<long long> h(2 * (sms / NCTA));
    cudaMemcpy(h.data(), out, h.size() * sizeof(long long), cudaMemcpyDeviceToHost);
    double worst = 0;
    for (size_t i = 0; i < h.size() / 2; ++i) worst = std::max(worst, (double)h[2 * i + 1]);
    best = std::min(best, worst / ((double)iters * ksteps * nissue));
  }
  cudaFree(out);
  return best;
}

int main() {
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, 0);
  const int sms = prop.multiProcessorCount;
  printf("%s, %d SMs: cycles per tcgen05.mma K16 (slowest SM, best of 3), math = 128*N*16 / (4096 MAC/clk/SM) = N/2\n", prop.name, sms);
  const int iters = 4000;
  for (int kch : {64, 32}) {
    for (int nissue : {1, 2, 3, 4, 6}) {
      printf("rows of %3d B (SWIZZLE_%dB), %d issuing warp(s), accumulators in rotation\n", kch * 2, kch * 2, nissue);
      printf("   N :   1-CTA M=128  |  pair M=256   (math per SM)   cycles per MMA, all issuers together\n");
      for (int n : {32, 64, 96, 128, 256}) {
        const int nacc = std::max(1, std::min(2, 512 / (n * nissue)));
        if (n * nissue > 512) continue;
        double c1 = run<1>(n, kch, iters, 4, sms, nacc, nissue);
        double c2 = run<2>(n, kch, iters, 4, sms - sms % 2, nacc, nissue);
        printf(" %3d :   %7.1f      |   %7.1f      (%5.1f)   [%d acc per issuer]\n", n, c1, c2, 128.0 * n * 16 / 4096.0, nacc);
      }
    }
  }
  return 0;
}
