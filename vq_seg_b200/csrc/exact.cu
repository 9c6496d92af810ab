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
  pdl_trigger();                                  // the overflow kernel behind may become resident (it waits for us)
  if (blockIdx.x == 0 && a.ovf_keys) {            // minima / tickets of the overflow kernel's split mode start from scratch
    for (int i = threadIdx.x; i < kOvfSplitCap; i += blockDim.x) { a.ovf_keys[i] = ~0ull; a.ovf_tickets[i] = 0; }
  }
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
                     a.ip != 0, a.ovf_rows, a.ovf_count);
    if (hl == 0 && valid && row >= 0) {
      if (a.idx_out) a.idx_out[row] = (long long)best_k + a.code_base;
      if (a.counts_out) atomicAdd(a.counts_out + best_k, 1ull);
      if (a.key_out)
        a.key_out[row] = ((unsigned long long)__float_as_uint(best) << 32) | (unsigned long long)(uint32_t)(best_k + a.code_base);
    }
    __syncwarp();
  }
  if (a.trace && threadIdx.x == 0) atomicMax((unsigned long long*)a.trace + 1, (unsigned long long)gtime());
}

// ---- rows the rescoring pass deferred (short-list overflow: more than kWorkCandCap codes within the filter's error
// bound of the row minimum, or the filter could not bound the row at all): every code is scored, one BLOCK per row.
// Warp w owns the 32-code slices w, w + 8, ...: lane = code; the slice's code rows arrive as coalesced
// [32 codes][32 dims] tiles through a padded shared-memory transpose (the next tile is in registers while the current
// one feeds the chains), so each lane runs its code's chain -- the same terms in the same order and the same K-blocking
// as chain_dist2 -- from conflict-free shared memory.
constexpr int kOvfWarps = 8;
constexpr int kOvfRows = 4;        // rows a block scores against one pass over the codebook when MANY rows overflow
__global__ void __launch_bounds__(kOvfWarps * 32, 2) overflow_rows_kernel(ExactArgs a, int rows_per_block) {
  extern __shared__ __align__(16) float smem_x[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int D = (int)a.x.D, K = a.K;
  const bool ip = a.ip != 0;
  const int Dp = (D + 3) & ~3;
  float* xs = smem_x;                                                   // [rows_per_block][D rounded to 4]
  float* tile = smem_x + (size_t)rows_per_block * Dp + (size_t)wib * (32 * 33);      // this warp's [32 codes][33]
  __shared__ float s_best[kOvfRows][kOvfWarps];
  __shared__ int s_bestk[kOvfRows][kOvfWarps];
  __shared__ float s_xn[kOvfRows];
  __shared__ int s_rows[kOvfRows];
  pdl_trigger();                                  // the gather behind may become resident ...
  pdl_wait();                                     // ... and we need the rescoring pass's overflow list
  const int n_ovf = *reinterpret_cast<volatile const int*>(a.ovf_count);
  const int L = ip ? D : D + 2;
  int kb = a.kblock;
  if (kb <= 0 || kb > L) kb = L;
  const int n_slices = (K + 31) / 32;
  // few rows, many codes (ONE overflow row at K = 65536 kept a single SM busy for 2-5 ms): the row's code slices are
  // split over `split` blocks that meet in a packed (ordered score bits, code) atomicMin; the block that draws the
  // row's last ticket writes the result
  int split = 1;
  if (a.ovf_keys && n_ovf > 0 && n_ovf <= kOvfSplitCap && 2 * n_ovf <= (int)gridDim.x)
    split = min((int)gridDim.x / n_ovf, (n_slices + kOvfWarps - 1) / kOvfWarps);
  if (split < 1) split = 1;
  if (split == 1 && rows_per_block == kOvfRows && n_ovf >= 2 * (int)gridDim.x) {
    // MANY rows overflow (a collapsed codebook: every row): one pass over the codebook per row re-reads K x D floats
    // from L2 for each of them (17 GB on the config-2 map, 4.1 ms).  Here a block scores kOvfRows rows against each
    // code tile: every lane keeps one chain per row, same terms in the same order as below.
    const int n_groups = (n_ovf + kOvfRows - 1) / kOvfRows;
    for (int g = blockIdx.x; g < n_groups; g += gridDim.x) {
      __syncthreads();                                                  // the previous group's xs / s_best are done with
      if (threadIdx.x < kOvfRows) s_rows[threadIdx.x] = g * kOvfRows + (int)threadIdx.x < n_ovf ? a.ovf_rows[g * kOvfRows + threadIdx.x] : -1;
      __syncthreads();
#pragma unroll
      for (int r = 0; r < kOvfRows; ++r) {
        const int row = s_rows[r];
        const float* xr = a.x.row(row < 0 ? 0 : row);
        for (int j = threadIdx.x; j < Dp; j += blockDim.x) xs[r * Dp + j] = (row >= 0 && j < D) ? __ldg(xr + (long long)j * a.x.sD) : 0.f;
      }
      __syncthreads();
      if (wib < kOvfRows && !ip) {
        const float* xw = xs + wib * Dp;
        const float xn = torch_order_sumsq_warp([&](long long j) { float v = xw[j]; return __fmul_rn(v, v); }, D, lane);
        if (lane == 0) s_xn[wib] = xn;
      }
      __syncthreads();
      float best[kOvfRows];
      int best_k[kOvfRows];
#pragma unroll
      for (int r = 0; r < kOvfRows; ++r) { best[r] = __int_as_float(0x7f800000); best_k[r] = 0x7fffffff; }
      for (int s = wib; s < n_slices; s += kOvfWarps) {
        const int k = s * 32 + lane;
        float c[kOvfRows];
        bool first = true;
        for (int blk = 0; blk < L; blk += kb) {
          const int end = min(blk + kb, L), dend = min(end, D);
          float t[kOvfRows];
#pragma unroll
          for (int r = 0; r < kOvfRows; ++r) t[r] = 0.f;
          for (int d0 = blk; d0 < dend; d0 += 32) {
            // (no register double-buffering here: four chains per lane keep the warp busy four times as long per
            // tile, and the sixteen warps of an SM cover each other's tile loads)
            __syncwarp();
            {
              float nxt[32];                                            // tile rows = codes, lane = dim d0 + lane
#pragma unroll
              for (int cc = 0; cc < 32; ++cc) {
                const int kc = s * 32 + cc;
                nxt[cc] = (kc < K && d0 + lane < dend) ? __ldg(a.E + (long long)kc * D + d0 + lane) : 0.f;
              }
#pragma unroll
              for (int cc = 0; cc < 32; ++cc) tile[cc * 33 + lane] = nxt[cc];
            }
            __syncwarp();
            const int w = min(32, dend - d0);
            const float* tr = tile + lane * 33;
            if (w == 32 && (d0 & 3) == 0) {
#pragma unroll
              for (int j = 0; j < 32; j += 4) {
                const float e0 = tr[j], e1 = tr[j + 1], e2 = tr[j + 2], e3 = tr[j + 3];
#pragma unroll
                for (int r = 0; r < kOvfRows; ++r) {
                  const float4 xv = *reinterpret_cast<const float4*>(xs + r * Dp + d0 + j);
                  t[r] = __fmaf_rn(xv.x, e0, t[r]); t[r] = __fmaf_rn(xv.y, e1, t[r]);
                  t[r] = __fmaf_rn(xv.z, e2, t[r]); t[r] = __fmaf_rn(xv.w, e3, t[r]);
                }
              }
            } else {
              for (int j = 0; j < w; ++j) {
                const float e = tr[j];
#pragma unroll
                for (int r = 0; r < kOvfRows; ++r) t[r] = __fmaf_rn(xs[r * Dp + d0 + j], e, t[r]);
              }
            }
          }
#pragma unroll
          for (int r = 0; r < kOvfRows; ++r) {
            float sc = ip ? -t[r] : -2.f * t[r];                         // exact
            if (end > D) {
              if (blk <= D) sc = __fadd_rn(sc, ip ? 0.f : s_xn[r]);      // term D   : |x|^2 * 1
              if (end > D + 1) sc = __fadd_rn(sc, k < K ? a.enorm[k] : 0.f); // term D+1 : 1 * |e|^2
            }
            c[r] = first ? sc : __fadd_rn(c[r], sc);
          }
          first = false;
        }
#pragma unroll
        for (int r = 0; r < kOvfRows; ++r)
          if (k < K) lexmin(best[r], best_k[r], score_key(c[r], ip), k);
      }
#pragma unroll
      for (int r = 0; r < kOvfRows; ++r) {
#pragma unroll
        for (int o = 16; o; o >>= 1) {
          const float d2 = __shfl_xor_sync(0xffffffffu, best[r], o);
          const int k2 = __shfl_xor_sync(0xffffffffu, best_k[r], o);
          lexmin(best[r], best_k[r], d2, k2);
        }
        if (lane == 0) { s_best[r][wib] = best[r]; s_bestk[r][wib] = best_k[r]; }
      }
      __syncthreads();
      if (threadIdx.x < kOvfRows && s_rows[threadIdx.x] >= 0) {
        const int r = threadIdx.x, row = s_rows[r];
        float b = s_best[r][0];
        int bk = s_bestk[r][0];
        for (int w2 = 1; w2 < kOvfWarps; ++w2) lexmin(b, bk, s_best[r][w2], s_bestk[r][w2]);
        if (bk == 0x7fffffff) bk = 0;
        if (a.idx_out) a.idx_out[row] = (long long)bk + a.code_base;
        if (a.counts_out) atomicAdd(a.counts_out + bk, 1ull);
        if (a.key_out)
          a.key_out[row] = ((unsigned long long)__float_as_uint(b) << 32) | (unsigned long long)(uint32_t)(bk + a.code_base);
      }
    }
    return;
  }
  for (int w = blockIdx.x; w < n_ovf * split; w += gridDim.x) {
    const int r = w / split, part = w - r * split;
    const int row = a.ovf_rows[r];
    const float* xr = a.x.row(row);
    __syncthreads();                                                    // the previous row's xs / s_best are done with
    for (int j = threadIdx.x; j < D; j += blockDim.x) xs[j] = __ldg(xr + (long long)j * a.x.sD);
    __syncthreads();
    if (wib == 0 && !ip) {
      const float xn = torch_order_sumsq_warp([&](long long j) { float v = xs[j]; return __fmul_rn(v, v); }, D, lane);
      if (lane == 0) s_xn[0] = xn;
    }
    __syncthreads();
    const float xnorm = ip ? 0.f : s_xn[0];
    float best = __int_as_float(0x7f800000);
    int best_k = 0x7fffffff;
    for (int s = part * kOvfWarps + wib; s < n_slices; s += split * kOvfWarps) {
      const int k = s * 32 + lane;
      float c = 0.f;
      bool first = true;
      for (int blk = 0; blk < L; blk += kb) {
        const int end = min(blk + kb, L), dend = min(end, D);
        float t = 0.f;
        float nxt[32];
        auto fetch = [&](int d0) {                                      // tile rows = codes, lane = dim d0 + lane
#pragma unroll
          for (int cc = 0; cc < 32; ++cc) {
            const int kc = s * 32 + cc;
            nxt[cc] = (kc < K && d0 + lane < dend) ? __ldg(a.E + (long long)kc * D + d0 + lane) : 0.f;
          }
        };
        if (blk < dend) fetch(blk);
        for (int d0 = blk; d0 < dend; d0 += 32) {
          __syncwarp();
#pragma unroll
          for (int cc = 0; cc < 32; ++cc) tile[cc * 33 + lane] = nxt[cc];
          __syncwarp();
          if (d0 + 32 < dend) fetch(d0 + 32);
          const int w = min(32, dend - d0);
          const float* tr = tile + lane * 33;
          if (w == 32) {
#pragma unroll
            for (int j = 0; j < 32; ++j) t = __fmaf_rn(xs[d0 + j], tr[j], t);
          } else {
            for (int j = 0; j < w; ++j) t = __fmaf_rn(xs[d0 + j], tr[j], t);
          }
        }
        float sc = ip ? -t : -2.f * t;                                   // exact
        if (end > D) {
          if (blk <= D) sc = __fadd_rn(sc, xnorm);                       // term D   : |x|^2 * 1
          if (end > D + 1) sc = __fadd_rn(sc, k < K ? a.enorm[k] : 0.f); // term D+1 : 1 * |e|^2
        }
        c = first ? sc : __fadd_rn(c, sc);
        first = false;
      }
      if (k < K) lexmin(best, best_k, score_key(c, ip), k);
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
      const float d2 = __shfl_xor_sync(0xffffffffu, best, o);
      const int k2 = __shfl_xor_sync(0xffffffffu, best_k, o);
      lexmin(best, best_k, d2, k2);
    }
    if (lane == 0) { s_best[0][wib] = best; s_bestk[0][wib] = best_k; }
    __syncthreads();
    if (threadIdx.x == 0) {
      for (int w2 = 1; w2 < kOvfWarps; ++w2) lexmin(best, best_k, s_best[0][w2], s_bestk[0][w2]);
      if (split > 1) {
        // order-preserving bits of the score (-0 counts as +0, like lexmin's ==), the code in the low word: the
        // 64-bit minimum is lexmin over all blocks of the row
        if (best == 0.f) best = 0.f;
        uint32_t u = __float_as_uint(best);
        u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
        atomicMin(a.ovf_keys + r, ((unsigned long long)u << 32) | (unsigned long long)(uint32_t)best_k);
        __threadfence();
        if (atomicAdd(a.ovf_tickets + r, 1) != split - 1) continue;      // (thread 0 only: the block's other threads wait at the next row's barrier)
        __threadfence();
        const unsigned long long key = atomicMin(a.ovf_keys + r, ~0ull);
        u = (uint32_t)(key >> 32);
        best = __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
        best_k = (int)(uint32_t)key;
      }
      if (best_k == 0x7fffffff) best_k = 0;
      if (a.idx_out) a.idx_out[row] = (long long)best_k + a.code_base;
      if (a.counts_out) atomicAdd(a.counts_out + best_k, 1ull);
      if (a.key_out)
        a.key_out[row] = ((unsigned long long)__float_as_uint(best) << 32) | (unsigned long long)(uint32_t)(best_k + a.code_base);
    }
  }
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
  // rescoring pass: rows of (x + staged candidates) per warp pair; as many warps per block as ~96 KB allow.
  // (Tried: staging two candidates per round beyond D = 256 to double the resident warps -- 16 per SM instead of 8 at
  // D = 512: slower, 290 against 250 us per million config-4 rows; rows with three or more candidates pay a second
  // round of dependent loads.)
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
  if (a.ovf_rows) {
    // (kOvfRows rows per block when their staging leaves room for two blocks per SM)
    const size_t tiles_bytes = (size_t)kOvfWarps * 32 * 33 * sizeof(float);
    const int rows_per_block = kOvfRows * xs_bytes + tiles_bytes <= 100 * 1024 ? kOvfRows : 1;
    const size_t osmem = rows_per_block * xs_bytes + tiles_bytes;
    if (osmem > 200 * 1024) return VQSEG_EUNSUPPORTED;
    static size_t oconf[kMaxDevices] = {0};
    if (int rc = ensure_dynamic_smem(overflow_rows_kernel, osmem, oconf)) return rc;
    long long ob = max_work < 2ll * num_sms() ? max_work : 2ll * num_sms();
    cudaError_t le = launch_dependent(overflow_rows_kernel, dim3((unsigned)ob), dim3(kOvfWarps * 32), osmem, st, pdl_enabled(), a, rows_per_block);
    if (le != cudaSuccess) return (int)le;
    VQSEG_LAUNCH_CHECK();
  }
  return 0;
}

}  // namespace vqseg
