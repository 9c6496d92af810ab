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
#include "exact_chain.cuh"

namespace vqseg {

// ---- |e_k|^2 in torch order (+ max |e|, row fingerprints), one warp per code ---------------------------------
__global__ void enorm_kernel(const float* __restrict__ E, int K, int D, int K_pad,
                             float* __restrict__ enorm, BlobHeader* hdr, unsigned long long* __restrict__ hash) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  prep_enorm(E, K, D, K_pad, enorm, hdr, hash, warp, (int)((gridDim.x * blockDim.x) >> 5), lane);
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
    const bool ip = a.ip != 0;
    const float xnorm = ip ? 0.f : torch_order_sumsq_warp([&](long long j) { float v = xs[j]; return __fmul_rn(v, v); }, D, lane);
    float best = __int_as_float(0x7f800000);   // +inf
    int best_k = 0x7fffffff;
    for (int k0 = lane; k0 < a.K; k0 += 128) {
      float c4[4];
      if (vec4) {
        chain_dist2_x4(xs, a.E, D, a.K, k0, xnorm, a.enorm, a.kblock, c4, ip);
      } else {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          int k = k0 + 32 * q;
          c4[q] = k < a.K ? chain_dist2<true>(xs, a.E + (long long)k * D, D, xnorm, a.enorm[k], a.kblock, false, ip) : 0.f;
        }
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        int k = k0 + 32 * q;
        if (k < a.K) lexmin(best, best_k, score_key(c4[q], ip), k);
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

// ---- rescoring pass over the filter's work records (body: exact_chain.cuh, shared with the fused tail of assign_tc3)
__global__ void __launch_bounds__(kExactWarps * 32) rescore_kernel(ExactArgs a, int stage_cap, long long rec_cap) {
  extern __shared__ __align__(16) float smem_x[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const int hw = lane >> 4, hl = lane & 15;
  const int D = (int)a.x.D;
  const int xs_stride = (D + 3) & ~3, es_stride = xs_stride + 4;
  const int row_floats = xs_stride + stage_cap * es_stride;
  float* xs_w = smem_x + (size_t)(2 * wib) * row_floats;          // this warp's two rows
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
    float best; int best_k, row;
    rescore_two_rows(a.x, a.E, a.K, a.enorm, a.kblock, rec_v, valid, xs_w, row_floats, stage_cap, lane, best, best_k, row,
                     a.ip != 0);
    if (hl == 0 && valid) {
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
  while (stage_cap > 1 && 2 * (xs_bytes + stage_cap * es_bytes) > 200 * 1024) --stage_cap;
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
