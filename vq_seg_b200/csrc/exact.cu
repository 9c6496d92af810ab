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
#include "codebook_prep.cuh"

namespace vqseg {

// ---- |e_k|^2 in torch order (+ max |e|, row fingerprints), one warp per code ---------------------------------
__global__ void enorm_kernel(const float* __restrict__ E, int K, int D, int K_pad,
                             float* __restrict__ enorm, BlobHeader* hdr, unsigned long long* __restrict__ hash) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  prep_enorm(E, K, D, K_pad, enorm, hdr, hash, warp, (int)((gridDim.x * blockDim.x) >> 5), lane);
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

// ---- brute force: every row against every code (VQSEG_ALGO_EXACT; the validator of the tensor-core path) ----------
__global__ void __launch_bounds__(kExactWarps * 32) exact_score_kernel(ExactArgs a) {
  extern __shared__ __align__(16) float smem_x[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int D = (int)a.x.D;
  const int xs_stride = (D + 3) & ~3;
  float* xs = smem_x + (size_t)wib * xs_stride;
  const bool vec4 = (D % 4 == 0) && ((reinterpret_cast<uintptr_t>(a.E) & 15) == 0);
  const long long n_rows = a.x.n_rows();
  for (long long n = (long long)blockIdx.x * kExactWarps + wib; n < n_rows; n += (long long)gridDim.x * kExactWarps) {
    const float* xr = a.x.row(n);
    __syncwarp();
    for (int j = lane; j < D; j += 32) xs[j] = __ldg(xr + (long long)j * a.x.sD);
    __syncwarp();
    const float xnorm = torch_order_sumsq_warp([&](long long j) { float v = xs[j]; return __fmul_rn(v, v); }, D, lane);
    float best = __int_as_float(0x7f800000);   // +inf
    int best_k = 0x7fffffff;
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
    // NaN distances never win above; torch.argmin would return the first NaN -- documented divergence.
#pragma unroll
    for (int o = 16; o; o >>= 1) {
      float d2 = __shfl_xor_sync(0xffffffffu, best, o);
      int k2 = __shfl_xor_sync(0xffffffffu, best_k, o);
      lexmin(best, best_k, d2, k2);
    }
    if (lane == 0) {
      if (best_k == 0x7fffffff) best_k = 0;
      if (a.idx_out) a.idx_out[n] = (long long)best_k + a.code_base;
      if (a.counts_out) atomicAdd(a.counts_out + best_k, 1ull);
      if (a.key_out)
        a.key_out[n] = ((unsigned long long)__float_as_uint(best) << 32) | (unsigned long long)(uint32_t)(best_k + a.code_base);
    }
  }
}

// ---- rescoring pass over the filter's work records ------------------------------------------------------------------
// Two undecided rows per warp, one per half-warp: lane hl of a half runs the exact chain of candidate hl.  With one
// row per warp only 2-5 lanes of 32 did chain work and six warps per scheduler made the 258-term dependent chains
// issue-bound (3.9 k cycles instead of ~1.1 k, round-1 trace); with two rows per warp, 8 warps per SM cover 2368 rows
// in one wave.  Dependent memory round trips per row: {work counter, record} -> {x row, candidate code rows, |e|^2}.
constexpr int kRsStage = 4;          // candidates per row staged in shared memory (more: read straight from L2)

__global__ void __launch_bounds__(kExactWarps * 32) rescore_kernel(ExactArgs a, int stage_cap, long long rec_cap) {
  extern __shared__ __align__(16) float smem_x[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const int hw = lane >> 4, hl = lane & 15;
  const int D = (int)a.x.D;
  const int xs_stride = (D + 3) & ~3, es_stride = xs_stride + 4;
  const int row_floats = xs_stride + stage_cap * es_stride;
  float* xs_w = smem_x + (size_t)(2 * wib) * row_floats;          // this warp's two rows
  float* xs = xs_w + (size_t)hw * row_floats;                     // this half-warp's row
  float* es = xs + xs_stride;
  const bool vec4 = (D % 4 == 0) && ((reinterpret_cast<uintptr_t>(a.E) & 15) == 0);
  auto gtime = [] { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return (long long)t; };
  if (a.trace && threadIdx.x == 0) atomicMin((unsigned long long*)a.trace, (unsigned long long)gtime());
  const long long w_first = ((long long)blockIdx.x * nwarps + wib) * 2;
  // the work counter and this warp's first two records are independent loads (the list has rec_cap slots)
  int rec_v = 0;
  if (w_first + hw < rec_cap && hl < 12) rec_v = __ldg(reinterpret_cast<const int*>(a.work + w_first + hw) + hl);
  long long n_work = (long long)__ldg(a.work_count);
  if (n_work > rec_cap) n_work = rec_cap;
  for (long long w = w_first; w < n_work; w += (long long)gridDim.x * nwarps * 2) {
    if (w != w_first) {
      rec_v = 0;
      if (w + hw < n_work && hl < 12) rec_v = __ldg(reinterpret_cast<const int*>(a.work + w + hw) + hl);
    }
    const bool valid = w + hw < n_work;
    const int row = __shfl_sync(0xffffffffu, rec_v, hw * 16);
    int cnt = __shfl_sync(0xffffffffu, rec_v, hw * 16 + 1);
    int my_k = __shfl_sync(0xffffffffu, rec_v, hw * 16 + 4 + (hl & 7));
    if (!valid) cnt = 0;
    const bool listed = cnt <= kWorkCandCap;
    const bool mine = valid && listed && hl < cnt;
    my_k = (mine && my_k >= 0 && my_k < a.K) ? my_k : 0;
    // ---- second wave of loads, all in flight together: the row, the staged candidate rows, the candidates' |e|^2
    const float* xr = a.x.row(valid ? row : 0);
    const float my_en = mine ? __ldg(a.enorm + my_k) : 0.f;
    __syncwarp();
    if (valid) {
      if (a.x.sD == 1 && vec4 && ((reinterpret_cast<uintptr_t>(xr) & 15) == 0)) {
        for (int j = 4 * hl; j < D; j += 64) *reinterpret_cast<float4*>(xs + j) = __ldg(reinterpret_cast<const float4*>(xr + j));
      } else {
        for (int j0 = 0; j0 < D; j0 += 128) {
          float t[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) { const int j = min(j0 + hl + 16 * u, D - 1); t[u] = __ldg(xr + (long long)j * a.x.sD); }
#pragma unroll
          for (int u = 0; u < 8; ++u) { const int j = j0 + hl + 16 * u; if (j < D) xs[j] = t[u]; }
        }
      }
    }
    const int n_staged = listed ? min(cnt, stage_cap) : 0;
    const int n_staged_w = max(n_staged, __shfl_xor_sync(0xffffffffu, n_staged, 16));      // warp-uniform trip count
    for (int c = 0; c < n_staged_w; ++c) {
      const float* er = a.E + (long long)__shfl_sync(0xffffffffu, my_k, hw * 16 + c, 32) * D;
      if (c >= n_staged) continue;
      if (vec4) {
        for (int j = 4 * hl; j < D; j += 64) *reinterpret_cast<float4*>(es + c * es_stride + j) = __ldg(reinterpret_cast<const float4*>(er + j));
      } else {
        for (int j = hl; j < D; j += 16) es[c * es_stride + j] = __ldg(er + j);
      }
    }
    __syncwarp();
    // ---- |x|^2 in ATen's order: the reduction is warp-wide (lane t = accumulator t), one row after the other
    const float xn0 = torch_order_sumsq_warp([&](long long j) { float v = xs_w[j]; return __fmul_rn(v, v); }, D, lane);
    const float xn1 = torch_order_sumsq_warp([&](long long j) { float v = xs_w[row_floats + j]; return __fmul_rn(v, v); }, D, lane);
    const float xnorm = hw ? xn1 : xn0;
    float best = __int_as_float(0x7f800000);   // +inf
    int best_k = 0x7fffffff;
    if (mine) {
      const float c2 = hl < n_staged ? chain_dist2_smem(xs, es + hl * es_stride, D, xnorm, my_en, a.kblock)
                                     : chain_dist2<true>(xs, a.E + (long long)my_k * D, D, xnorm, my_en, a.kblock, vec4);
      lexmin(best, best_k, __fsqrt_rn(fmaxf(c2, 0.f)), my_k);
    } else if (valid && !listed) {
      // the short-list overflowed (or the filter deferred the row): every code, 16 lanes striding over K
      for (int k = hl; k < a.K; k += 16) {
        const float c2 = chain_dist2<true>(xs, a.E + (long long)k * D, D, xnorm, a.enorm[k], a.kblock, vec4);
        lexmin(best, best_k, __fsqrt_rn(fmaxf(c2, 0.f)), k);
      }
    }
    // NaN distances never win above; torch.argmin would return the first NaN -- documented divergence.
#pragma unroll
    for (int o = 8; o; o >>= 1) {
      float d2 = __shfl_xor_sync(0xffffffffu, best, o);
      int k2 = __shfl_xor_sync(0xffffffffu, best_k, o);
      lexmin(best, best_k, d2, k2);
    }
    if (hl == 0 && valid) {
      if (best_k == 0x7fffffff) best_k = 0;
      if (a.idx_out) a.idx_out[row] = (long long)best_k + a.code_base;
      if (a.counts_out) atomicAdd(a.counts_out + best_k, 1ull);
      if (a.key_out)
        a.key_out[row] = ((unsigned long long)__float_as_uint(best) << 32) | (unsigned long long)(uint32_t)(best_k + a.code_base);
    }
    __syncwarp();
  }
  if (a.trace && threadIdx.x == 0) atomicMax((unsigned long long*)a.trace + 1, (unsigned long long)gtime());
}

int launch_enorm(const float* E, int K, int D, int K_pad, float* enorm, BlobHeader* hdr, unsigned long long* hash,
                 cudaStream_t st) {
  int warps = K_pad;
  int threads = 256, blocks = (warps * 32 + threads - 1) / threads;
  const int cap = num_sms() * 8;
  if (blocks > cap) blocks = cap;
  enorm_kernel<<<blocks, threads, 0, st>>>(E, K, D, K_pad, enorm, hdr, hash);
  VQSEG_LAUNCH_CHECK();
  return 0;
}

int launch_exact(const ExactArgs& a, long long max_work, cudaStream_t st) {
  if (max_work <= 0) return 0;
  const int D = (int)a.x.D;
  const size_t xs_bytes = (size_t)((D + 3) & ~3) * sizeof(float);
  if (!a.work) {                                             // brute force: one row per warp
    const size_t smem = (size_t)kExactWarps * xs_bytes;
    if (smem > 200 * 1024) return VQSEG_EUNSUPPORTED;
    static size_t configured[kMaxDevices] = {0};
    if (int rc = ensure_dynamic_smem(exact_score_kernel, smem, configured)) return rc;
    long long blocks = (max_work + kExactWarps - 1) / kExactWarps;
    const long long cap = (long long)num_sms() * 8;
    if (blocks > cap) blocks = cap;
    exact_score_kernel<<<(unsigned)blocks, kExactWarps * 32, smem, st>>>(a);
    VQSEG_LAUNCH_CHECK();
    return 0;
  }
  // rescoring pass: rows of (x + staged candidates) per warp pair; as many warps per block as ~96 KB allow
  const size_t es_bytes = xs_bytes + 16;
  int stage_cap = kRsStage;
  while (stage_cap > 0 && 2 * (xs_bytes + stage_cap * es_bytes) > 200 * 1024) --stage_cap;
  const size_t row_bytes = xs_bytes + stage_cap * es_bytes;
  if (2 * row_bytes > 200 * 1024) return VQSEG_EUNSUPPORTED;
  int nwarps = (int)((96 * 1024) / (2 * row_bytes));
  nwarps = nwarps < 1 ? 1 : (nwarps > kExactWarps ? kExactWarps : nwarps);
  const size_t smem = (size_t)nwarps * 2 * row_bytes;
  static size_t configured[kMaxDevices] = {0};
  if (int rc = ensure_dynamic_smem(rescore_kernel, smem, configured)) return rc;
  long long blocks = (max_work + 2 * nwarps - 1) / (2 * nwarps);
  const long long cap = (long long)num_sms() * 2;
  if (blocks > cap) blocks = cap;
  rescore_kernel<<<(unsigned)blocks, nwarps * 32, smem, st>>>(a, stage_cap, max_work);
  VQSEG_LAUNCH_CHECK();
  return 0;
}

}  // namespace vqseg
