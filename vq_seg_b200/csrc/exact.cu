// Exact fp32 scorer: the arithmetic of ATen's CPU `_euclidean_dist` + argmin, restated for the GPU.
//
// Reference call sites: torch.cdist + torch.argmin (vector_quantizer/vq_img.py:167-168) and
// -torch.cdist + torch.argmax in kmeans (vq_img.py:39-41).
//
// d(x, e_k) = sqrt(max(c, 0)) where c is the fp32 result of the (D+2)-term augmented dot product
//   [-2x_0 .. -2x_{D-1}, |x|^2, 1] . [e_k0 .. e_k,D-1, 1, |e_k|^2]
// evaluated as MKL's sgemm evaluates it on CPU: one FMA chain per output in increasing term order,
// restarted every `kblock` terms with the partial results added in order (probed: kblock = inf for
// D+2 <= 384, 384-blocks beyond).  Because scaling by -2 commutes with fp32
// rounding, the chain over (-2 x_j) e_j equals -2 * chain over x_j e_j bit for bit.
//
// One warp owns one row: lane t loads x[t::32] (and is accumulator t of the torch-order |x|^2), then
// every lane runs the chain for its own candidate code(s).  The same kernel is (a) the brute-force
// exact path over all K codes, (b) the rescoring pass over the short-list written by the tcgen05
// filter (assign_tc.cu).  Ties: the lowest code index wins, like torch.argmin.
#include "common.cuh"
#include "kernels.cuh"

namespace vqseg {

// ---- |e_k|^2 in torch order, one warp per code -------------------------------------------------
__global__ void enorm_kernel(const float* __restrict__ E, int K, int D, int K_pad,
                             float* __restrict__ enorm, BlobHeader* hdr) {
  int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= K_pad) return;
  if (warp >= K) { if (lane == 0) enorm[warp] = 3.0e38f; return; }
  const float* row = E + (long long)warp * D;
  float s = torch_order_sumsq_warp([&](long long j) { float v = row[j]; return __fmul_rn(v, v); }, D, lane);
  float amax = 0.f;
  for (int j = lane; j < D; j += 32) amax = fmaxf(amax, fabsf(row[j]));
#pragma unroll
  for (int o = 16; o; o >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
  if (lane == 0) {
    enorm[warp] = s;
    if (hdr) {
      atomicMax(&hdr->max_enorm_bits, __float_as_uint(s));
      atomicMax(&hdr->max_abs_bits, __float_as_uint(amax));
    }
  }
}

// one augmented chain, generic block size. xs: shared x row, e: global code row
template <bool LDG>
__device__ __forceinline__ float chain_dist2(const float* __restrict__ xs, const float* __restrict__ e,
                                             int D, float xnorm, float enorm, int kb, bool vec4) {
  auto ld = [](const float* p) { return LDG ? __ldg(p) : *p; };
  const int L = D + 2;
  if (kb <= 0 || kb > L) kb = L;
  float c = 0.f;
  bool first = true;
  for (int blk = 0; blk < L; blk += kb) {
    int end = min(blk + kb, L);
    int dend = min(end, D);
    float t = 0.f;
    int j = blk;
    if (vec4) {
      for (; j < dend && (j & 3); ++j) t = __fmaf_rn(xs[j], ld(e + j), t);
      for (; j + 4 <= dend; j += 4) {
        float4 ev = __ldg(reinterpret_cast<const float4*>(e + j));
        float4 xv = *reinterpret_cast<const float4*>(xs + j);
        t = __fmaf_rn(xv.x, ev.x, t); t = __fmaf_rn(xv.y, ev.y, t);
        t = __fmaf_rn(xv.z, ev.z, t); t = __fmaf_rn(xv.w, ev.w, t);
      }
    }
    for (; j < dend; ++j) t = __fmaf_rn(xs[j], ld(e + j), t);
    float s = -2.f * t;                                   // exact
    if (end > D) {
      if (blk <= D) s = __fadd_rn(s, xnorm);              // term D   : |x|^2 * 1
      if (end > D + 1) s = __fadd_rn(s, enorm);           // term D+1 : 1 * |e|^2
    }
    c = first ? s : __fadd_rn(c, s);
    first = false;
  }
  return c;
}

// chain over smem-staged operands (both 16-byte aligned): loads are hoisted 8 terms ahead of the FMA chain
__device__ __forceinline__ float chain_dist2_smem(const float* __restrict__ xs, const float* __restrict__ es,
                                                  int D, float xnorm, float enorm, int kb) {
  const int L = D + 2;
  if (kb <= 0 || kb > L) kb = L;
  float c = 0.f;
  bool first = true;
  for (int blk = 0; blk < L; blk += kb) {
    const int end = min(blk + kb, L);
    const int dend = min(end, D);
    float t = 0.f;
    int j = blk;
    for (; j < dend && (j & 7); ++j) t = __fmaf_rn(xs[j], es[j], t);
    if (j + 8 <= dend) {
      // the next 8 terms are fetched from smem while the current 8 dependent FMAs retire; two register sets
      // alternate (16 terms per trip) so no register moves sit between the FMAs -- several warps share a
      // scheduler here and the loop is issue-bound
#define VQSEG_LD8(X0, X1, E0, E1, at)                                                     \
      X0 = *reinterpret_cast<const float4*>(xs + (at)); X1 = *reinterpret_cast<const float4*>(xs + (at) + 4); \
      E0 = *reinterpret_cast<const float4*>(es + (at)); E1 = *reinterpret_cast<const float4*>(es + (at) + 4)
#define VQSEG_FMA8(X0, X1, E0, E1)                                                        \
      t = __fmaf_rn(X0.x, E0.x, t); t = __fmaf_rn(X0.y, E0.y, t); t = __fmaf_rn(X0.z, E0.z, t); t = __fmaf_rn(X0.w, E0.w, t); \
      t = __fmaf_rn(X1.x, E1.x, t); t = __fmaf_rn(X1.y, E1.y, t); t = __fmaf_rn(X1.z, E1.z, t); t = __fmaf_rn(X1.w, E1.w, t)
      float4 ax0, ax1, ae0, ae1, bx0, bx1, be0, be1;
      VQSEG_LD8(ax0, ax1, ae0, ae1, j);                       // A holds terms [j, j + 8)
      while (j + 24 <= dend) {
        VQSEG_LD8(bx0, bx1, be0, be1, j + 8);
        VQSEG_FMA8(ax0, ax1, ae0, ae1);
        VQSEG_LD8(ax0, ax1, ae0, ae1, j + 16);
        VQSEG_FMA8(bx0, bx1, be0, be1);
        j += 16;
      }
      if (j + 16 <= dend) {
        VQSEG_LD8(bx0, bx1, be0, be1, j + 8);
        VQSEG_FMA8(ax0, ax1, ae0, ae1);
        VQSEG_FMA8(bx0, bx1, be0, be1);
        j += 16;
      } else {
        VQSEG_FMA8(ax0, ax1, ae0, ae1);
        j += 8;
      }
#undef VQSEG_LD8
#undef VQSEG_FMA8
    }
    for (; j < dend; ++j) t = __fmaf_rn(xs[j], es[j], t);
    float s = -2.f * t;
    if (end > D) {
      if (blk <= D) s = __fadd_rn(s, xnorm);
      if (end > D + 1) s = __fadd_rn(s, enorm);
    }
    c = first ? s : __fadd_rn(c, s);
    first = false;
  }
  return c;
}

// four independent chains per lane (codes k0 + 32*q), used by the all-codes path for ILP
__device__ __forceinline__ void chain_dist2_x4(const float* __restrict__ xs, const float* __restrict__ E,
                                               int D, int K, int k0, float xnorm,
                                               const float* __restrict__ enorm, int kb, float out[4]) {
  const int L = D + 2;
  if (kb <= 0 || kb > L) kb = L;
  const float* e[4];
  bool ok[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) { int k = k0 + 32 * q; ok[q] = k < K; e[q] = E + (long long)(ok[q] ? k : 0) * D; }
  float c[4] = {0.f, 0.f, 0.f, 0.f};
  bool first = true;
  for (int blk = 0; blk < L; blk += kb) {
    int end = min(blk + kb, L);
    int dend = min(end, D);
    float t[4] = {0.f, 0.f, 0.f, 0.f};
    int j = blk;
    for (; j < dend && (j & 3); ++j) {
      float xv = xs[j];
#pragma unroll
      for (int q = 0; q < 4; ++q) t[q] = __fmaf_rn(xv, __ldg(e[q] + j), t[q]);
    }
    for (; j + 4 <= dend; j += 4) {
      float4 xv = *reinterpret_cast<const float4*>(xs + j);
      float4 ev[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) ev[q] = __ldg(reinterpret_cast<const float4*>(e[q] + j));
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        t[q] = __fmaf_rn(xv.x, ev[q].x, t[q]); t[q] = __fmaf_rn(xv.y, ev[q].y, t[q]);
        t[q] = __fmaf_rn(xv.z, ev[q].z, t[q]); t[q] = __fmaf_rn(xv.w, ev[q].w, t[q]);
      }
    }
    for (; j < dend; ++j) {
      float xv = xs[j];
#pragma unroll
      for (int q = 0; q < 4; ++q) t[q] = __fmaf_rn(xv, __ldg(e[q] + j), t[q]);
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      float s = -2.f * t[q];
      if (end > D) {
        if (blk <= D) s = __fadd_rn(s, xnorm);
        if (end > D + 1) s = __fadd_rn(s, enorm[ok[q] ? k0 + 32 * q : 0]);
      }
      c[q] = first ? s : __fadd_rn(c[q], s);
    }
    first = false;
  }
#pragma unroll
  for (int q = 0; q < 4; ++q) out[q] = c[q];
}

__device__ __forceinline__ void lexmin(float& d, int& k, float d2, int k2) {
  if (d2 < d || (d2 == d && k2 < k)) { d = d2; k = k2; }
}


constexpr int kExactWarps = 8;

__global__ void __launch_bounds__(kExactWarps * 32) exact_score_kernel(ExactArgs a) {
  extern __shared__ __align__(16) float smem_x[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int D = (int)a.x.D;
  const int xs_stride = (D + 3) & ~3;
  const bool stage_e = a.stage_e != 0;                   // smem also holds cand_cap code rows per warp
  float* xs = smem_x + (size_t)wib * (xs_stride + (stage_e ? a.cand_cap * (xs_stride + 4) : 0));
  const bool vec4 = (D % 4 == 0) && ((reinterpret_cast<uintptr_t>(a.E) & 15) == 0);
  auto gtime = [] { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return (long long)t; };
  if (a.trace && threadIdx.x == 0) atomicMin((unsigned long long*)a.trace, (unsigned long long)gtime());
  long long c0 = clock64(), c1 = c0, c2 = c0, c3 = c0, c4 = c0;
  const long long n_rows = a.x.n_rows();
  const long long w_first = (long long)blockIdx.x * kExactWarps + wib;
  // the work counter and this warp's first row id are independent loads (the list has n_rows slots)
  int n_spec = (a.work_rows && w_first < n_rows) ? __ldg(a.work_rows + w_first) : 0;
  long long n_work = a.work_rows ? (long long)__ldg(a.work_count) : n_rows;
  if (a.work_rows && n_work > n_rows) n_work = n_rows;
  for (long long w = w_first; w < n_work; w += (long long)gridDim.x * kExactWarps) {
    const long long n = a.work_rows ? (w == w_first ? n_spec : a.work_rows[w]) : w;
    const float* xr = a.x.row(n);
    // wave 2 of dependent loads, all issued before the first use: candidate count, candidate ids (lane c
    // holds candidate c) and the row itself (8 strided loads per lane)
    c1 = clock64();
    int cnt = a.cand_cnt ? __ldg(a.cand_cnt + n) : -1;
    int my_k = 0;
    if (a.cand_idx && lane < a.cand_cap) my_k = __ldg(a.cand_idx + n * a.cand_cap + lane);
    __syncwarp();
    for (int j0 = 0; j0 < D; j0 += 256) {
      float t[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {      // unconditional (clamped) loads: a predicated load gets fused with its predicated store and serialises
        const int j = min(j0 + lane + 32 * u, D - 1);
        t[u] = __ldg(xr + (long long)j * a.x.sD);
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) { const int j = j0 + lane + 32 * u; if (j < D) xs[j] = t[u]; }
    }
    float best = __int_as_float(0x7f800000);   // +inf
    int best_k = 0x7fffffff;
    const bool listed = cnt >= 0 && cnt <= a.cand_cap;
    __syncwarp(); c2 = clock64();
    my_k = (listed && lane < cnt && my_k >= 0 && my_k < a.K) ? my_k : 0;
    if (listed && stage_e) {
      // wave 3: candidate code rows (coalesced, into smem; row stride es_stride = D rounded to 4, + 4:
      // 16-byte aligned rows, and the <= 8 lanes chaining different rows read disjoint bank quads) and
      // each candidate's |e|^2
      float* es = xs + xs_stride;
      const int es_stride = xs_stride + 4;
      const float my_en = lane < cnt ? __ldg(a.enorm + my_k) : 0.f;
      float xnorm;
      if (vec4 && D <= 256 && cnt <= 4) {
        // common case: <= 4 candidates of <= 256 dims: their rows travel in registers while |x|^2 is reduced
        float4 ev[4][2];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const float* er = a.E + (long long)__shfl_sync(0xffffffffu, my_k, c) * D;
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int j = 4 * lane + 128 * h;
            ev[c][h] = (c < cnt && j < D) ? __ldg(reinterpret_cast<const float4*>(er + j)) : make_float4(0.f, 0.f, 0.f, 0.f);
          }
        }
        __syncwarp();
        c3 = clock64();
        xnorm = torch_order_sumsq_warp([&](long long j) { float v = xs[j]; return __fmul_rn(v, v); }, D, lane);
#pragma unroll
        for (int c = 0; c < 4; ++c)
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int j = 4 * lane + 128 * h;
            if (c < cnt && j < D) *reinterpret_cast<float4*>(es + c * es_stride + j) = ev[c][h];
          }
        __syncwarp();
        c4 = clock64();
      } else {
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          if (c < cnt) {
            const float* er = a.E + (long long)__shfl_sync(0xffffffffu, my_k, c) * D;
            if (vec4) {
              for (int j = 4 * lane; j < D; j += 128)
                *reinterpret_cast<float4*>(es + c * es_stride + j) = __ldg(reinterpret_cast<const float4*>(er + j));
            } else {
              for (int j = lane; j < D; j += 32) es[c * es_stride + j] = __ldg(er + j);
            }
          }
        }
        __syncwarp();
        c3 = clock64();
        xnorm = torch_order_sumsq_warp([&](long long j) { float v = xs[j]; return __fmul_rn(v, v); }, D, lane);
        c4 = clock64();
      }
      if (lane < cnt) {
        float c2 = chain_dist2_smem(xs, es + lane * es_stride, D, xnorm, my_en, a.kblock);
        lexmin(best, best_k, __fsqrt_rn(fmaxf(c2, 0.f)), my_k);
      }
    } else if (listed) {
      __syncwarp();
      const float xnorm = torch_order_sumsq_warp([&](long long j) { float v = xs[j]; return __fmul_rn(v, v); }, D, lane);
      if (lane < cnt) {
        float c2 = chain_dist2<true>(xs, a.E + (long long)my_k * D, D, xnorm, a.enorm[my_k], a.kblock, vec4);
        lexmin(best, best_k, __fsqrt_rn(fmaxf(c2, 0.f)), my_k);
      }
    } else {
      __syncwarp();
      const float xnorm = torch_order_sumsq_warp([&](long long j) { float v = xs[j]; return __fmul_rn(v, v); }, D, lane);
      for (int k0 = lane; k0 < a.K; k0 += 128) {
        float c4[4];
        if (vec4) {
          chain_dist2_x4(xs, a.E, D, a.K, k0, xnorm, a.enorm, a.kblock, c4);
        } else {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            int k = k0 + 32 * q;
            c4[q] = k < a.K ? chain_dist2<true>(xs, a.E + (long long)k * D, D, xnorm, a.enorm[k], a.kblock, false) : 0.f;
          }
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          int k = k0 + 32 * q;
          if (k < a.K) lexmin(best, best_k, __fsqrt_rn(fmaxf(c4[q], 0.f)), k);
        }
      }
    }
    // NaN distances never win above; torch.argmin would return the first NaN -- documented divergence.
#pragma unroll
    for (int o = 16; o; o >>= 1) {
      float d2 = __shfl_xor_sync(0xffffffffu, best, o);
      int k2 = __shfl_xor_sync(0xffffffffu, best_k, o);
      lexmin(best, best_k, d2, k2);
    }
    if (a.trace && lane == 0) {
      long long c5 = clock64();
      atomicAdd((unsigned long long*)a.trace + 2, (unsigned long long)(c1 - c0));   // until row id known
      atomicAdd((unsigned long long*)a.trace + 3, (unsigned long long)(c2 - c1));   // wave 2 (cnt, ids, x row)
      atomicAdd((unsigned long long*)a.trace + 4, (unsigned long long)(c3 - c2));   // wave 3 (code rows)
      atomicAdd((unsigned long long*)a.trace + 5, (unsigned long long)(c4 - c3));   // |x|^2
      atomicAdd((unsigned long long*)a.trace + 6, (unsigned long long)(c5 - c4));   // chains + argmin
      atomicAdd((unsigned long long*)a.trace + 7, 1ull);
    }
    if (lane == 0) {
      if (best_k == 0x7fffffff) best_k = 0;
      if (a.idx_out) a.idx_out[n] = (long long)best_k + a.code_base;
      if (a.counts_out) atomicAdd(a.counts_out + best_k, 1ull);
      if (a.key_out)
        a.key_out[n] = ((unsigned long long)__float_as_uint(best) << 32) | (unsigned long long)(uint32_t)(best_k + a.code_base);
    }
  }
  if (a.trace && threadIdx.x == 0) atomicMax((unsigned long long*)a.trace + 1, (unsigned long long)gtime());
  if (a.usage_out) {
    // all counts are final once every block is through: the last one (ticket) reduces them -- saves a launch
    __shared__ int s_last;
    __shared__ int s_zero[kExactWarps];
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = (atomicAdd(a.done_blocks, 1) == (int)gridDim.x - 1);
    __syncthreads();
    if (s_last) {
      __threadfence();
      int z = 0;
      for (int k = threadIdx.x; k < a.K; k += blockDim.x) z += (__ldcg((const unsigned long long*)a.counts_out + k) == 0ull);
#pragma unroll
      for (int o = 16; o; o >>= 1) z += __shfl_xor_sync(0xffffffffu, z, o);
      if (lane == 0) s_zero[wib] = z;
      __syncthreads();
      if (threadIdx.x == 0) {
        int t = 0;
        for (int w2 = 0; w2 < kExactWarps; ++w2) t += s_zero[w2];
        *a.usage_out = __fmul_rn(100.f, __fdiv_rn((float)t, (float)a.K));
      }
    }
  }
}

int launch_enorm(const float* E, int K, int D, int K_pad, float* enorm, BlobHeader* hdr, cudaStream_t st) {
  int warps = K_pad;
  int threads = 256, blocks = (warps * 32 + threads - 1) / threads;
  enorm_kernel<<<blocks, threads, 0, st>>>(E, K, D, K_pad, enorm, hdr);
  VQSEG_LAUNCH_CHECK();
  return 0;
}

int launch_exact(const ExactArgs& a_in, long long max_work, cudaStream_t st) {
  if (max_work <= 0) return 0;
  ExactArgs a = a_in;
  const int D = (int)a.x.D;
  const size_t xs_bytes = (size_t)((D + 3) & ~3) * sizeof(float);
  const size_t es_bytes = (size_t)a.cand_cap * (((D + 3) & ~3) + 4) * sizeof(float);
  a.stage_e = (a.cand_idx != nullptr && kExactWarps * (xs_bytes + es_bytes) <= 100 * 1024) ? 1 : 0;
  size_t smem = (size_t)kExactWarps * (xs_bytes + (a.stage_e ? es_bytes : 0));
  if (smem > 200 * 1024) return VQSEG_EUNSUPPORTED;
  static size_t configured = 0;
  if (smem > 48 * 1024 && smem > configured) {
    cudaError_t e = cudaFuncSetAttribute(exact_score_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    configured = smem;
  }
  long long blocks = (max_work + kExactWarps - 1) / kExactWarps;
  long long cap = (long long)num_sms() * (a.stage_e ? 3 : 8);
  if (blocks > cap) blocks = cap;
  exact_score_kernel<<<(unsigned)blocks, kExactWarps * 32, smem, st>>>(a);
  VQSEG_LAUNCH_CHECK();
  return 0;
}

}  // namespace vqseg
