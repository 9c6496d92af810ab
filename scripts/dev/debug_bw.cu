// Developer micro-benchmark (not part of the product path): achievable global->register bandwidth of the
// producer-style access patterns under the same conditions as the tcgen05 kernels (8 loading warps per SM,
// large shared-memory carve-out, data resident in L2 or not).  Exposed as vqseg_debug_load_bandwidth.
#include "common.cuh"
#include "tc_common.cuh"

namespace vqseg {

template <int PATTERN, int DEPTH>
__global__ void __launch_bounds__(256) load_bw_kernel(const float* __restrict__ x, long long n_floats,
                                                         long long row_stride, float* __restrict__ sink) {
  extern __shared__ unsigned char pad_smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // each CTA walks "tiles" of 128 px x 64 rows like the producers do: tile t -> rows r0.., px p0..
  const long long px_per_row = row_stride;                 // floats per row
  const long long tiles_per_row = px_per_row / 128;
  const long long n_rows = n_floats / row_stride;
  constexpr int kGroups = (PATTERN == 1) ? (DEPTH + 3) / 4 : (DEPTH + 7) / 8;     // 64-row groups per iteration
  const long long n_tiles = tiles_per_row * (n_rows / (64 * kGroups));
  float acc = 0.f;
  for (long long t = blockIdx.x; t < n_tiles; t += gridDim.x) {
    const long long rt = t / tiles_per_row, pt = t % tiles_per_row;
    const float* base = x + rt * 64 * kGroups * row_stride + pt * 128;
    if (PATTERN == 0) {            // LDG.128: warp -> 16 px, lane = quad*8+g8, 8 rows per thread: 8 x 64 B per instr
      const int g8 = lane & 7, quad = lane >> 3;
      const float* p = base + (long long)(8 * g8) * row_stride + 16 * warp + 4 * quad;
      float4 v[DEPTH];
#pragma unroll
      for (int j = 0; j < DEPTH; ++j)
        asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v[j].x), "=f"(v[j].y), "=f"(v[j].z), "=f"(v[j].w)
                     : "l"(p + (long long)((j & 7) + 64 * (j >> 3)) * row_stride));
#pragma unroll
      for (int j = 0; j < DEPTH; ++j) acc += v[j].x + v[j].y + v[j].z + v[j].w;
    } else if (PATTERN == 1) {     // LDG.256: warps 0-3 rows 0..31? -> warp pair covers 32 px: 8 x 128 B per instr
      const int g8 = lane & 7, o4 = lane >> 3;
      const int pw = warp & 3, half = warp >> 2;        // half selects rows 0-31 / 32-63 via j offset
      const float* p = base + (long long)(8 * g8 % 32 + 32 * half) * row_stride + 32 * pw + 8 * o4;
      float v[DEPTH][8];
#pragma unroll
      for (int j = 0; j < DEPTH; ++j)
        asm volatile("ld.global.nc.L1::no_allocate.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=f"(v[j][0]), "=f"(v[j][1]), "=f"(v[j][2]), "=f"(v[j][3]), "=f"(v[j][4]), "=f"(v[j][5]), "=f"(v[j][6]), "=f"(v[j][7])
                     : "l"(p + (long long)((j & 3) + 64 * (j >> 2)) * row_stride));
#pragma unroll
      for (int j = 0; j < DEPTH; ++j)
#pragma unroll
        for (int e = 0; e < 8; ++e) acc += v[j][e];
    } else {                       // LDG.128 fully contiguous 512 B per warp instruction
      const float* p = base + (long long)(8 * warp) * row_stride + 4 * lane;
      float4 v[DEPTH];
#pragma unroll
      for (int j = 0; j < DEPTH; ++j)
        asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v[j].x), "=f"(v[j].y), "=f"(v[j].z), "=f"(v[j].w)
                     : "l"(p + (long long)((j & 7) + 64 * (j >> 3)) * row_stride));
#pragma unroll
      for (int j = 0; j < DEPTH; ++j) acc += v[j].x + v[j].y + v[j].z + v[j].w;
    }
  }
  if (acc == 12345.678f) sink[0] = acc;
}

// Pattern 3: the same tiles, fetched by the bulk-copy engine (cp.async.bulk, 512 B per copy = one 128-pixel row
// segment, 16 segments = one 8 KiB stage) into a ring of STAGES stages; 8 consumer warps read each stage back
// from shared memory.  One issuing warp, so the bytes in flight are set by the ring, not by registers.
template <int STAGES, int VARIANT>
__global__ void __launch_bounds__(384) load_bw_bulk_kernel(const float* __restrict__ x, long long n_floats,
                                                           long long row_stride, float* __restrict__ sink) {
  extern __shared__ __align__(128) unsigned char bulk_smem[];
  const uint32_t base = smem_u32(bulk_smem);
  const uint32_t bar_full = base + STAGES * 8192, bar_empty = bar_full + 8 * STAGES;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(bar_full + 8 * s, VARIANT == 2 ? 4 : 1); mbar_init(bar_empty + 8 * s, 8); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const long long tiles_per_row = row_stride / 128;
  const long long n_rows = n_floats / row_stride;
  const long long n_tiles = tiles_per_row * (n_rows / 64);
  float acc = 0.f;
  long long q = 0;
  if (warp >= 8) {
    const int lw = warp - 8;                     // loader warp 0..3
    if (VARIANT != 2 && lw != 0) return;
    for (long long t = blockIdx.x; t < n_tiles; t += gridDim.x) {
      const long long rt = t / tiles_per_row, pt = t % tiles_per_row;
      const float* tb = x + rt * 64 * row_stride + pt * 128;
      for (int sub = 0; sub < 4; ++sub, ++q) {
        const int s = (int)(q % STAGES);
        const uint32_t use = (uint32_t)(q / STAGES);
        mbar_wait(bar_empty + 8 * s, (use & 1) ^ 1);
        if (VARIANT == 0) {
          if (lane == 0) mbar_arrive_expect_tx(bar_full + 8 * s, 8192);
          __syncwarp();
          if (lane < 16) bulk_g2s(base + s * 8192 + lane * 512, tb + (long long)(sub * 16 + lane) * row_stride, 512, bar_full + 8 * s);
        } else if (VARIANT == 1) {               // linear: stage q of this CTA = 8 KiB contiguous
          if (lane == 0) {
            mbar_arrive_expect_tx(bar_full + 8 * s, 8192);
            bulk_g2s(base + s * 8192, x + (t * 4 + sub) * 2048, 8192, bar_full + 8 * s);
          }
        } else {                                 // four loader warps, 4 segments each (barrier count 4)
          if (lane == 0) mbar_arrive_expect_tx(bar_full + 8 * s, 2048);
          __syncwarp();
          if (lane < 4) bulk_g2s(base + s * 8192 + (lw * 4 + lane) * 512, tb + (long long)(sub * 16 + lw * 4 + lane) * row_stride, 512, bar_full + 8 * s);
        }
      }
    }
  } else {
    for (long long t = blockIdx.x; t < n_tiles; t += gridDim.x) {
      for (int sub = 0; sub < 4; ++sub, ++q) {
        const int s = (int)(q % STAGES);
        const uint32_t use = (uint32_t)(q / STAGES);
        mbar_wait(bar_full + 8 * s, use & 1);
        const float4* st = reinterpret_cast<const float4*>(bulk_smem + s * 8192);
        const float4 v0 = st[threadIdx.x], v1 = st[threadIdx.x + 256];
        acc += v0.x + v0.y + v0.z + v0.w + v1.x + v1.y + v1.z + v1.w;
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_empty + 8 * s);
      }
    }
  }
  if (acc == 12345.678f) sink[0] = acc;
}
template <int STAGES, int VARIANT>
static int launch_bw_bulk(const float* x, long long n, long long rs, float* sink, cudaStream_t st) {
  const int smem = STAGES * 8192 + 16 * STAGES + 128;
  const int pad = 200 * 1024 > smem ? 200 * 1024 : smem;      // one CTA per SM, like the filter kernel
  cudaFuncSetAttribute(load_bw_bulk_kernel<STAGES, VARIANT>, cudaFuncAttributeMaxDynamicSharedMemorySize, pad);
  load_bw_bulk_kernel<STAGES, VARIANT><<<num_sms(), 384, pad, st>>>(x, n, rs, sink);
  VQSEG_LAUNCH_CHECK();
  return 0;
}

static int g_bw_blocks_per_sm = 1;
template <int P, int D>
static int launch_bw(const float* x, long long n, long long rs, float* sink, cudaStream_t st) {
  const int smem = g_bw_blocks_per_sm == 1 ? 200 * 1024 : 0;
  cudaFuncSetAttribute(load_bw_kernel<P, D>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  load_bw_kernel<P, D><<<num_sms() * g_bw_blocks_per_sm, 256, smem, st>>>(x, n, rs, sink);
  VQSEG_LAUNCH_CHECK();
  return 0;
}

}  // namespace vqseg

namespace vqseg {
__global__ void __launch_bounds__(640, 1) null_kernel(int* p) { if (p && threadIdx.x == 9999) p[0] = 1; }
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(640, 1) null_cluster_kernel(int* p) { if (p && threadIdx.x == 9999) p[0] = 1; }
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(640, 1) null_cluster_tmem_kernel(int* p) {
  __shared__ uint32_t slot;
  if ((threadIdx.x >> 5) == 0) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&slot)), "r"(512u));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
  }
  __syncthreads();
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  if ((threadIdx.x >> 5) == 0) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(slot), "r"(512u));
  if (p && threadIdx.x == 9999) p[0] = 1;
}
}  // namespace vqseg

// kind 0: 148 x 640 threads, no smem; 1: + 228 KB dynamic smem; 2: + cluster of 2; 3: + TMEM alloc/dealloc + cluster sync
extern "C" int vqseg_debug_null_launch(int kind, void* stream) {
  using namespace vqseg;
  cudaStream_t st = (cudaStream_t)stream;
  const int smem = 226 * 1024;
  if (kind == 0) null_kernel<<<num_sms(), 640, 0, st>>>(nullptr);
  else if (kind == 1) { cudaFuncSetAttribute(null_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); null_kernel<<<num_sms(), 640, smem, st>>>(nullptr); }
  else if (kind == 2) { cudaFuncSetAttribute(null_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); null_cluster_kernel<<<num_sms(), 640, smem, st>>>(nullptr); }
  else { cudaFuncSetAttribute(null_cluster_tmem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem - 1024); null_cluster_tmem_kernel<<<num_sms(), 640, smem - 1024, st>>>(nullptr); }
  VQSEG_LAUNCH_CHECK();
  return 0;
}

extern "C" int vqseg_debug_load_bandwidth(const float* x, int64_t n_floats, int64_t row_stride, int pattern, int depth,
                                          float* sink, void* stream) {
  using namespace vqseg;
  cudaStream_t st = (cudaStream_t)stream;
  g_bw_blocks_per_sm = pattern >= 10 ? pattern / 10 : 1;     // pattern 42 -> 4 blocks per SM, pattern 2
  pattern %= 10;
  if (pattern == 3 && depth == 4) return launch_bw_bulk<4, 0>(x, n_floats, row_stride, sink, st);
  if (pattern == 3 && depth == 16) return launch_bw_bulk<16, 0>(x, n_floats, row_stride, sink, st);
  if (pattern == 4 && depth == 4) return launch_bw_bulk<4, 1>(x, n_floats, row_stride, sink, st);
  if (pattern == 4 && depth == 16) return launch_bw_bulk<16, 1>(x, n_floats, row_stride, sink, st);
  if (pattern == 5 && depth == 4) return launch_bw_bulk<4, 2>(x, n_floats, row_stride, sink, st);
  if (pattern == 5 && depth == 16) return launch_bw_bulk<16, 2>(x, n_floats, row_stride, sink, st);
  if (pattern == 0 && depth == 8) return launch_bw<0, 8>(x, n_floats, row_stride, sink, st);
  if (pattern == 0 && depth == 24) return launch_bw<0, 24>(x, n_floats, row_stride, sink, st);
  if (pattern == 1 && depth == 12) return launch_bw<1, 12>(x, n_floats, row_stride, sink, st);
  if (pattern == 2 && depth == 24) return launch_bw<2, 24>(x, n_floats, row_stride, sink, st);
  if (pattern == 1 && depth == 4) return launch_bw<1, 4>(x, n_floats, row_stride, sink, st);
  if (pattern == 2 && depth == 8) return launch_bw<2, 8>(x, n_floats, row_stride, sink, st);
  return VQSEG_EINVAL;
}
