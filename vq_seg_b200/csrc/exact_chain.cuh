// The exact fp32 scorer's arithmetic as device functions: ATen's CPU `_euclidean_dist` chain (see exact.cu) and the
// rescoring of two undecided rows by one warp.
//
// Inner-product mode (`ip`, the cosine codebook: einsum('n d, e d -> n e') + argmax, vq_img.py:104-107, and
// `samples @ means^T` in kmeans, :37): the same MKL sgemm chain over the D plain terms (no augmentation), returned
// NEGATED -- negation commutes with fp32 rounding, so "lowest index among the minimal -<x, e>" is torch.argmax's
// "first index among the maximal <x, e>".  Shared by exact.cu (brute force + standalone rescoring kernel) and by
// the fused tail of assign_tc3.cu.
#pragma once
#include "common.cuh"
#include "kernels.cuh"

namespace vqseg {

// one augmented chain, generic block size. xs: shared x row, e: global code row
template <bool LDG>
static __device__ __noinline__ float chain_dist2(const float* __restrict__ xs, const float* __restrict__ e,
                                             int D, float xnorm, float enorm, int kb, bool vec4, bool ip = false) {
  auto ld = [](const float* p) { return LDG ? __ldg(p) : *p; };
  const int L = ip ? D : D + 2;
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
    float s = ip ? -t : -2.f * t;                         // exact
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
static __device__ __noinline__ float chain_dist2_smem(const float* __restrict__ xs, const float* __restrict__ es,
                                                  int D, float xnorm, float enorm, int kb, bool ip = false) {
  const int L = ip ? D : D + 2;
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
    float s = ip ? -t : -2.f * t;
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
                                               const float* __restrict__ enorm, int kb, float out[4], bool ip = false) {
  const int L = ip ? D : D + 2;
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
      float s = ip ? -t[q] : -2.f * t[q];
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

// what the argmin runs over: the distance sqrt(max(c, 0)) (cdist), or the negated inner product itself
__device__ __forceinline__ float score_key(float c, bool ip) { return ip ? c : __fsqrt_rn(fmaxf(c, 0.f)); }

__device__ __forceinline__ void lexmin(float& d, int& k, float d2, int k2) {
  if (d2 < d || (d2 == d && k2 < k)) { d = d2; k = k2; }
}



// |x|^2 of a shared-memory row in ATen's summation order; one copy of the (long) reduction code per kernel
static __device__ __noinline__ float torch_sumsq_smem(const float* xs, int D, int lane) {
  return torch_order_sumsq_warp([&](long long j) { float v = xs[j]; return __fmul_rn(v, v); }, D, lane);
}

constexpr int kRsStage = 4;          // candidates per row staged in shared memory (more: read straight from L2)

// Two undecided rows per warp, one per half-warp: lane hl of a half runs the exact chain of candidate hl.  With one
// row per warp only 2-5 lanes of 32 did chain work and six warps per scheduler made the 258-term dependent chains
// issue-bound (3.9 k cycles instead of ~1.1 k, round-1 trace).  Dependent memory round trips per row:
// record -> {x row, first group of candidate code rows, |e|^2} -> chains (-> next group of `stage_cap` candidates).
// Every candidate's code row is staged in shared memory before its chain runs: a chain fed by dependent L2 loads
// costs ~300 cycles per 4 terms (a row with one unstaged candidate took 19 k cycles).
// rec_v: lane 16 * hw + i holds int i of the half's WorkRec (i < 12); valid: per half-warp.  xs_w: the warp's shared
// scratch, two rows of row_floats = xs_stride + stage_cap * (xs_stride + 4) floats, 16-byte aligned; stage_cap >= 1.
// Returns on every lane of a half: the winning distance, code and the row id.  All 32 lanes must call.
__device__ __forceinline__ void rescore_two_rows(const Rows& x, const float* __restrict__ E, int K,
                                                 const float* __restrict__ enorm, int kblock, int rec_v, bool valid,
                                                 float* xs_w, int row_floats, int stage_cap, int lane,
                                                 float& best, int& best_k, int& row, bool ip = false,
                                                 int* ovf_rows = nullptr, int* ovf_count = nullptr) {
  const int hw = lane >> 4, hl = lane & 15;
  const int D = (int)x.D;
  const int xs_stride = (D + 3) & ~3, es_stride = xs_stride + 4;
  float* xs = xs_w + (size_t)hw * row_floats;                     // this half-warp's row
  float* es = xs + xs_stride;
  const bool vec4 = (D % 4 == 0) && ((reinterpret_cast<uintptr_t>(E) & 15) == 0);
  row = __shfl_sync(0xffffffffu, rec_v, hw * 16);
  int cnt = __shfl_sync(0xffffffffu, rec_v, hw * 16 + 1);
  int my_k = __shfl_sync(0xffffffffu, rec_v, hw * 16 + 4 + (hl & 7));
  if (!valid) cnt = 0;
  const bool listed = cnt <= kWorkCandCap;
  const bool mine = valid && listed && hl < cnt;
  my_k = (mine && my_k >= 0 && my_k < K) ? my_k : 0;
  const float* xr = x.row(valid ? row : 0);
  const float my_en = mine ? __ldg(enorm + my_k) : 0.f;
  const int n_list = listed ? cnt : 0;
  const int n_list_w = max(n_list, __shfl_xor_sync(0xffffffffu, n_list, 16));      // warp-uniform trip count
  best = __int_as_float(0x7f800000);   // +inf
  best_k = 0x7fffffff;
  float xnorm = 0.f;
  __syncwarp();
  for (int g0 = 0; g0 < n_list_w || g0 == 0; g0 += stage_cap) {
    // ---- stage candidates [g0, g0 + stage_cap) of each half (and, the first time round, the row itself)
    if (vec4 && D <= 512 && stage_cap <= 4) {
      // common case, per 256-dim half: the row (16 strided words per lane) and up to four code rows (4 float4 each,
      // two at a time in registers) are all requested before the first shared-memory store.  (Extending this path from
      // D <= 256 to D <= 512 left config 4's rescoring at 2.8 ms: the stall samples ncu shows on the stores behind
      // the loads are the first batch's latency, not a lack of batching.)
      const bool contig = x.sD == 1 && ((reinterpret_cast<uintptr_t>(xr) & 15) == 0);
#pragma unroll 1
      for (int dh = 0; dh < D; dh += 256) {
        float t[16];
        if (g0 == 0) {
#pragma unroll
          for (int u = 0; u < 16; ++u) {
            const int j = dh + (contig ? (4 * hl + 64 * (u >> 2) + (u & 3)) : (hl + 16 * u));
            t[u] = (valid && j < D) ? __ldg(xr + (long long)j * x.sD) : 0.f;
          }
        }
#pragma unroll 1
        for (int c0 = 0; c0 < stage_cap; c0 += 2) {
          float4 ev[2][4];
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            const int ci = g0 + c0 + c;                                     // index in the half's short-list
            const bool on = c0 + c < stage_cap && ci < n_list;
            const float* er = E + (long long)__shfl_sync(0xffffffffu, my_k, hw * 16 + (ci & 7), 32) * D;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int j = dh + 4 * hl + 64 * i;
              ev[c][i] = (on && j < D) ? __ldg(reinterpret_cast<const float4*>(er + j)) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
          }
          if (g0 == 0 && c0 == 0 && valid) {
#pragma unroll
            for (int u = 0; u < 16; ++u) {
              const int j = dh + (contig ? (4 * hl + 64 * (u >> 2) + (u & 3)) : (hl + 16 * u));
              if (j < D) xs[j] = t[u];
            }
          }
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            const bool on = c0 + c < stage_cap && g0 + c0 + c < n_list;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int j = dh + 4 * hl + 64 * i;
              if (on && j < D) *reinterpret_cast<float4*>(es + (c0 + c) * es_stride + j) = ev[c][i];
            }
          }
        }
      }
    } else {
      if (g0 == 0 && valid) {
        if (x.sD == 1 && vec4 && ((reinterpret_cast<uintptr_t>(xr) & 15) == 0)) {
          for (int j = 4 * hl; j < D; j += 64) *reinterpret_cast<float4*>(xs + j) = __ldg(reinterpret_cast<const float4*>(xr + j));
        } else {
          for (int j0 = 0; j0 < D; j0 += 128) {
            float t[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) { const int j = min(j0 + hl + 16 * u, D - 1); t[u] = __ldg(xr + (long long)j * x.sD); }
#pragma unroll
            for (int u = 0; u < 8; ++u) { const int j = j0 + hl + 16 * u; if (j < D) xs[j] = t[u]; }
          }
        }
      }
      for (int c = 0; c < stage_cap; ++c) {
        const int ci = g0 + c;
        const float* er = E + (long long)__shfl_sync(0xffffffffu, my_k, hw * 16 + (ci & 7), 32) * D;
        if (ci >= n_list) continue;
        if (vec4) {
          for (int j = 4 * hl; j < D; j += 64) *reinterpret_cast<float4*>(es + c * es_stride + j) = __ldg(reinterpret_cast<const float4*>(er + j));
        } else {
          for (int j = hl; j < D; j += 16) es[c * es_stride + j] = __ldg(er + j);
        }
      }
    }
    __syncwarp();
    if (g0 == 0 && !ip) {
      // |x|^2 in ATen's order: the reduction is warp-wide (lane t = accumulator t), one row after the other
      const float xn0 = torch_sumsq_smem(xs_w, D, lane);
      const float xn1 = torch_sumsq_smem(xs_w + row_floats, D, lane);
      xnorm = hw ? xn1 : xn0;
    }
    if (mine && hl >= g0 && hl < g0 + stage_cap) {
      const float c2 = chain_dist2_smem(xs, es + (hl - g0) * es_stride, D, xnorm, my_en, kblock, ip);
      lexmin(best, best_k, score_key(c2, ip), my_k);
    }
    __syncwarp();
  }
  if (valid && !listed && ovf_rows) {
    // the short-list overflowed (or the filter deferred the row): every code must be scored.  One half-warp doing
    // that alone is a 250-500 us latency chain (K = 512, D = 512-1024: a single such row on a real feature map made
    // the whole kernel 20x slower), so the row goes to overflow_rows_kernel, which spreads it over a block
    if (hl == 0) ovf_rows[atomicAdd(ovf_count, 1)] = row;
    row = -1;                                              // (no early return: the shuffles below are warp-wide)
  } else if (valid && !listed) {
    for (int k = hl; k < K; k += 16) {
      const float c2 = chain_dist2<true>(xs, E + (long long)k * D, D, xnorm, enorm[k], kblock, vec4, ip);
      lexmin(best, best_k, score_key(c2, ip), k);
    }
  }
  // NaN distances never win above; torch.argmin would return the first NaN -- documented divergence.
#pragma unroll
  for (int o = 8; o; o >>= 1) {
    float d2 = __shfl_xor_sync(0xffffffffu, best, o);
    int k2 = __shfl_xor_sync(0xffffffffu, best_k, o);
    lexmin(best, best_k, d2, k2);
  }
  if (best_k == 0x7fffffff) best_k = 0;
}

}  // namespace vqseg
