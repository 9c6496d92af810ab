// Developer microbenchmark (not part of the product library): what does one tcgen05.mma.kind::f16 cost on this B200?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I vq_seg_b200/csrc -I include -o gpurun_out/mma_rate scripts/dev/mma_rate.cu
// One elected thread per CTA pair issues `n` back-to-back M=256 x N=256 x K=16 (cta_group::2) MMAs on resident
// shared-memory operands; optional: 8 warps per CTA read the other accumulator with tcgen05.ld all the while.
// Reports SM cycles per MMA and the SM clock the kernel actually ran at (clock64 / globaltimer).
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include "tc_common.cuh"

using namespace vqseg;

__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ unsigned long long gtimer() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }

// mode bits: 1 readers, 2 commit after every 4 MMAs, 4 mbarrier wait + fence before every 4 MMAs, 8 a loader thread
// streams 16 KiB bulk copies (L2-resident source) into a 3-stage ring all the while
template <int CTAS>
__global__ void __launch_bounds__(352, 1) mma_rate_kernel(int n_mma, int mode, int random_data, long long* out, const unsigned char* src) {
  const int with_readers = mode & 1;
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const uint32_t sbase = smem_u32(smem);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = CTAS == 2 ? cluster_ctarank() : 0;
  // A: 4 chunks of 16 KiB, B: 4 chunks of 16 KiB (128 rows x 64 fp16, SWIZZLE_128B K-major)
  uint32_t* w = reinterpret_cast<uint32_t*>(smem);
  for (int i = threadIdx.x; i < 8 * 16384 / 4; i += blockDim.x) {
    uint32_t h = (uint32_t)i * 2654435761u + blockIdx.x * 40503u;
    h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
    // two fp16 values in [-2, 2): exponent 0x3c..0x3f
    uint32_t v = (h & 0x83ff83ffu) | 0x3c003c00u;
    w[i] = random_data ? v : 0u;
  }
  constexpr int kCtl = 11 * 16384;
  const uint32_t bar = sbase + kCtl, bar_ready = bar + 8, bar_dummy = bar + 16, bar_ld = bar + 24;   // bar_ld[3]
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + kCtl + 64);
  volatile int* stop = reinterpret_cast<volatile int*>(smem + kCtl + 128);
  if (threadIdx.x == 0) {
    mbar_init(bar, 1); mbar_init(bar_ready, 1); mbar_init(bar_dummy, 1);
    for (int s = 0; s < 3; ++s) mbar_init(bar_ld + 8 * s, 1);
    *stop = 0; asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 8) {
    if (CTAS == 2) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32((const void*)tmem_slot)), "r"(512u));
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32((const void*)tmem_slot)), "r"(512u));
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  if (CTAS == 2) cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  long long c0 = 0, c1 = 0; unsigned long long g0 = 0, g1 = 0;
  if (warp == 8) {
    if (rank == 0 && elect_one()) {
      const uint32_t idesc = make_idesc_f16(CTAS == 2 ? 256 : 128, 256);
      // both accumulators get written once before anybody reads them
      tc_mma_f16_2cta(tmem_base, make_desc(sbase), make_desc(sbase + 4 * 16384), idesc, 0u);
      tc_mma_f16_2cta(tmem_base + 256, make_desc(sbase), make_desc(sbase + 4 * 16384), idesc, 0u);
      tc_commit_2cta(bar_ready);
      mbar_wait(bar_ready, 0);
      g0 = gtimer(); c0 = clock64();
      for (int i = 0; i < n_mma; ++i) {
        if ((i & 3) == 0 && (mode & 4)) { mbar_wait(bar_ready, 0); tc_fence_after(); }
        const int ch = (i >> 2) & 3, k = i & 3;
        const uint64_t ad = make_desc(sbase + ch * 16384) + (uint64_t)(2 * k);
        const uint64_t bd = make_desc(sbase + (4 + ch) * 16384) + (uint64_t)(2 * k);
        if (CTAS == 2) tc_mma_f16_2cta(tmem_base, ad, bd, idesc, i ? 1u : 0u);
        else tc_mma_f16(tmem_base, ad, bd, idesc, i ? 1u : 0u);
        if ((i & 3) == 3 && (mode & 2)) tc_commit_2cta(bar_dummy);
      }
      if (CTAS == 2) tc_commit_2cta(bar); else tc_commit(bar);
      mbar_wait(bar, 0);
      c1 = clock64(); g1 = gtimer();
      *stop = 1;
      if (CTAS == 2) { asm volatile("st.shared::cluster.u32 [%0], %1;" ::"r"(mapa_u32(sbase + kCtl + 128, 1)), "r"(1) : "memory"); }
      out[(blockIdx.x / CTAS) * 4 + 0] = c1 - c0;
      out[(blockIdx.x / CTAS) * 4 + 1] = (long long)(g1 - g0);
    }
  } else if (warp == 9) {
    if ((mode & 8) && lane == 0) {
      long long n = 0;
      for (int s = 0; s < 3; ++s) {
        mbar_arrive_expect_tx(bar_ld + 8 * s, 16384);
        bulk_g2s(sbase + (8 + s) * 16384, src + (size_t)((blockIdx.x * 7 + s) & 255) * 16384, 16384, bar_ld + 8 * s);
      }
      while (!*stop) {
        const int s = (int)(n % 3);
        mbar_wait(bar_ld + 8 * s, (uint32_t)(n / 3) & 1);
        mbar_arrive_expect_tx(bar_ld + 8 * s, 16384);
        bulk_g2s(sbase + (8 + s) * 16384, src + (size_t)((blockIdx.x * 7 + n) & 255) * 16384, 16384, bar_ld + 8 * s);
        ++n;
      }
      for (int k = 0; k < 3; ++k, ++n) mbar_wait(bar_ld + 8 * (int)(n % 3), (uint32_t)(n / 3) & 1);   // drain
      if (rank == 0) out[(blockIdx.x / CTAS) * 4 + 3] = n;
    }
  } else if (with_readers) {
    mbar_wait(bar_ready, 0);
    tc_fence_after();
    // 8 warps read the OTHER accumulator (columns 256..511) in a loop: does the drain slow the MMAs down?
    const uint32_t lane_addr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + 256 + (warp >> 2) * 128;
    uint32_t acc = 0;
    long long reads = 0;
    while (!*stop) {
      for (int c = 0; c < 128; c += 32) { uint32_t v[32]; tmem_ld32(lane_addr + c, v); acc ^= v[0] ^ v[31]; }
      ++reads;
    }
    if (acc == 0x12345678u) out[1 << 20] = acc;
    if (lane == 0 && warp == 0 && rank == 0) out[(blockIdx.x / CTAS) * 4 + 2] = reads;
  }
  tc_fence_before();
  __syncthreads();
  if (CTAS == 2) cluster_sync_all();
  if (warp == 8) {
    if (CTAS == 2) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u));
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u));
  }
}

static void run(int ctas, int groups, int n_mma, int readers, int random_data) {
  long long* out;
  static unsigned char* src = nullptr;
  if (!src) { cudaMalloc(&src, 256 * 16384); cudaMemset(src, 0x3c, 256 * 16384); }
  cudaMalloc(&out, ((1 << 20) + 8) * sizeof(long long));
  cudaMemset(out, 0, ((1 << 20) + 8) * sizeof(long long));
  const int smem = 11 * 16384 + 2048;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int rep = 0; rep < 2; ++rep) {
    cudaEventRecord(e0);
    if (ctas == 2) {
      cudaFuncSetAttribute(mma_rate_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(2 * groups); cfg.blockDim = dim3(352); cfg.dynamicSmemBytes = smem;
      cudaLaunchAttribute at; at.id = cudaLaunchAttributeClusterDimension; at.val.clusterDim.x = 2; at.val.clusterDim.y = 1; at.val.clusterDim.z = 1;
      cfg.attrs = &at; cfg.numAttrs = 1;
      cudaLaunchKernelEx(&cfg, mma_rate_kernel<2>, n_mma, readers, random_data, out, (const unsigned char*)src);
    } else {
      cudaFuncSetAttribute(mma_rate_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
      mma_rate_kernel<1><<<groups, 352, smem>>>(n_mma, readers, random_data, out, src);
    }
    cudaEventRecord(e1);
    cudaError_t err = cudaDeviceSynchronize();
    if (err != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(err)); exit(1); }
  }
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  std::vector<long long> h(groups * 4);
  cudaMemcpy(h.data(), out, groups * 4 * sizeof(long long), cudaMemcpyDeviceToHost);
  double cyc = 0, ns = 0, rd = 0, ld = 0;
  for (int g = 0; g < groups; ++g) { cyc += h[g * 4]; ns += h[g * 4 + 1]; rd += h[g * 4 + 2]; ld += h[g * 4 + 3]; }
  cyc /= groups; ns /= groups; rd /= groups; ld /= groups;
  const double flop = 2.0 * (ctas == 2 ? 256 : 128) * 256 * 16 * n_mma * groups;
  printf("cta_group::%d  groups %3d  n_mma %6d  readers %d  data %s : %7.1f cycles/MMA  SM clock %6.0f MHz  kernel %8.3f ms  %7.1f TFLOP/s"
         "  (tmem drains %.0f, bulk copies %.0f = %.1f B/clk/SM)\n", ctas, groups, n_mma, readers, random_data ? "random" : "zeros ",
         cyc / n_mma, cyc / ns * 1e3, ms, flop / (ns * 1e-9) / 1e12, rd, ld, ld * 16384.0 / cyc);
  cudaFree(out);
}

int main(int argc, char** argv) {
  run(2, 1, 4096, 0, 1);
  for (int mode : {0, 2, 4, 6, 8, 14}) run(2, 74, 32768, mode, 1);
  return 0;
}
