// Fused distance GEMM + argmin on the 5th-gen tensor cores (tcgen05 / TMEM), sm_100a only.
//
// Replaces torch.cdist + torch.argmin (vector_quantizer/vq_img.py:167-168, and :39-41 in kmeans)
// without ever materialising the N x K distance matrix.
//
// The tensor cores evaluate the ranking score  s_k = |e_k|^2 - 2 x.e_k  with fp16 operands and
// fp32 accumulation; that is a FILTER, not the answer.  Every code whose approximate score lies
// within a proven error bound of the row minimum is kept in a per-row short-list, and rows with
// more than one survivor are re-scored by exact.cu with the reference's own fp32 arithmetic.  The
// result therefore equals the exact path bit for bit (DESIGN.md §"why the filter is safe").
//
// Structure (one CTA per SM, persistent over 128-row tiles; unit = 128 rows x 256 codes):
//   warps 0-7  A producers : x (fp32, any strides) -> registers (two chunks of prefetch) -> fp16 ->
//                            SWIZZLE_128B K-major smem tile (128 rows x 64 dims), plus sum x^2 per row
//   warp  8    B loader    : cp.async.bulk (TMA bulk engine) of the pre-packed fp16 (-2 s E) tiles
//   warp  9    MMA issuer  : tcgen05.mma.cta_group::1.kind::f16, M=128 N=256 K=16, accumulators in
//                            TMEM, pre-loaded with s |e_k|^2 so the MMA yields the (scaled) score itself
//   warps 10-13 epilogue   : tcgen05.ld -> running row min -> short-list -> tcgen05.st re-init
// Pipelines: smem full/empty mbarriers (4 stages), TMEM full/empty (2 x 256 columns).
#include "tc_common.cuh"

namespace vqseg {

constexpr int kTileM = 128;           // rows (latent vectors) per tile = UMMA M = TMEM lanes
constexpr int kUnitN = 256;           // codes per unit = UMMA N
constexpr int kStages = 4;
constexpr int kAStageBytes = kTileM * kDChunk * 2;        // 16 KiB
constexpr int kBStageBytes = kUnitN * kDChunk * 2;        // 32 KiB
constexpr int kProducerWarps = 8;
constexpr int kEpiWarp0 = 10;
constexpr int kTcThreads = 14 * 32;
constexpr int kCandCap = 8;           // short-list entries kept per row before falling back to "all codes"
constexpr int kXsqBufs = 8;
constexpr uint32_t kIdesc = make_idesc_f16(kTileM, kUnitN);
constexpr int kEnormSmem = 2048;      // codes whose scaled norms are staged in smem (else read from L2)

struct TcSmem {
  // dynamic smem, 1024-aligned base
  static constexpr int off_a = 0;
  static constexpr int off_b = off_a + kStages * kAStageBytes;
  static constexpr int off_cand = off_b + kStages * kBStageBytes;            // [128][kCandCap] code indices
  static constexpr int off_xsq = off_cand + kTileM * kCandCap * 4;           // [kXsqBufs][128] float2 {|x|^2, |fp16(x)-x|^2}
  static constexpr int off_enorm = off_xsq + kXsqBufs * kTileM * 8;          // [kEnormSmem] float: s*|e|^2
  static constexpr int off_bar = off_enorm + kEnormSmem * 4;                 // mbarriers
  static constexpr int off_tmem = off_bar + 8 * (2 * kStages + 4);
  static constexpr int total = off_tmem + 16 + 1024;   // + slack for the runtime 1024-B alignment
};

struct TcArgs {
  Rows x;
  const unsigned char* blob;      // prepared codebook
  long long n_rows;
  int n_tiles, n_cc, n_dc;        // tiles of 128 rows, code chunks of 256, dim chunks of 64
  int K;
  float tau;                      // relative error bound of one fp16 x fp16 score vs the exact fp32 scorer
  // outputs
  long long* idx_out; unsigned long long* counts_out; long long code_base;
  int force_rescore;              // 1 -> every row goes to the exact pass (sharded mode needs exact distances)
  int* cand_idx; int* cand_cnt; int* work_rows; int* work_count;
  long long* trace;               // optional (dev tool): [cta][role][256] clock64 stamps
};

#define VQ_TRACE(role, slot) do { if (a.trace && lane == 0 && (slot) < 256) \
    a.trace[((long long)blockIdx.x * 4 + (role)) * 256 + (slot)] = clock64(); } while (0)

// ---- codebook packing -----------------------------------------------------------------------------
// image tile (cb, dc): 128 codes x 64 dims of fp16(-2 * s * e), SWIZZLE_128B K-major:
//   byte = row*128 + ((col/8) ^ (row & 7))*16 + (col % 8)*2
__global__ void __launch_bounds__(256) pack_codebook_kernel(const float* __restrict__ E, int K, int D,
                                                            unsigned char* __restrict__ blob) {
  BlobHeader* hdr = reinterpret_cast<BlobHeader*>(blob);
  const int K_pad = hdr->K_pad, D_pad = hdr->D_pad;
  // power-of-two prescale: max |s e| in [8, 16)
  const uint32_t mbits = hdr->max_abs_bits;
  int ex = (int)((mbits >> 23) & 0xff) - 127;
  if (mbits == 0) ex = 3;
  int se = 3 - ex;
  se = se < -100 ? -100 : (se > 100 ? 100 : se);
  const float s = __uint_as_float((uint32_t)(127 + se) << 23);
  float* enorm_s = reinterpret_cast<float*>(blob + hdr->off_enorm) + K_pad;      // scaled copy after the exact one
  const float* enorm = reinterpret_cast<const float*>(blob + hdr->off_enorm);
  __half* img = reinterpret_cast<__half*>(blob + hdr->off_image);
  const int n_dc = D_pad / kDChunk;
  const long long total = (long long)K_pad * (D_pad / 8);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int k = (int)(i / (D_pad / 8)), d8 = (int)(i % (D_pad / 8)) * 8;
    __align__(16) __half h[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float v = (k < K && d8 + j < D) ? E[(long long)k * D + d8 + j] : 0.f;
      h[j] = __float2half_rn(-2.f * s * v);
    }
    const int cb = k / kCodeBlock, row = k % kCodeBlock, dc = d8 / kDChunk, c8 = (d8 % kDChunk) / 8;
    unsigned char* tile = reinterpret_cast<unsigned char*>(img) + ((long long)cb * n_dc + dc) * kTileBytes;
    *reinterpret_cast<uint4*>(tile + row * 128 + ((c8 ^ (row & 7)) * 16)) = *reinterpret_cast<const uint4*>(h);
  }
  // |e|^2 limbs for the augmented K step: s|e_k|^2 = c * (h1 + h2 + h3), c a power of two putting the
  // largest norm in [2^13, 2^14) so every limb is a normal/subnormal fp16 with residual <= 2^-24 c
  const float men = __uint_as_float(hdr->max_enorm_bits) * s;
  int ce = (int)((__float_as_uint(men) >> 23) & 0xff) - 127 - 13;
  if (men == 0.f) ce = 0;
  ce = ce < -14 ? -14 : (ce > 15 ? 15 : ce);
  const float c = __uint_as_float((uint32_t)(127 + ce) << 23);
  const float cinv = __uint_as_float((uint32_t)(127 - ce) << 23);
  unsigned char* augbase = blob + hdr->off_aug;
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < K_pad; k += gridDim.x * blockDim.x) {
    enorm_s[k] = k < K ? enorm[k] * s : 3.0e38f;
    float v = k < K ? enorm[k] * s * cinv : 60000.f;
    __half h1, h2, h3;
    if (k < K) {
      if (!(v <= 60000.f)) { atomicOr(&hdr->flags, 1u); v = 60000.f; }
      h1 = __float2half_rn(v);
      const float r1 = v - __half2float(h1);
      h2 = __float2half_rn(r1);
      h3 = __float2half_rn(r1 - __half2float(h2));
    } else {
      h1 = h2 = h3 = __float2half_rn(60000.f);          // pad codes can never come near the row minimum
    }
    const int cb = k / kCodeBlock, row = k % kCodeBlock;
    // SWIZZLE_NONE K-major core matrices: 8 rows x 16 B; byte = (row/8)*256 + khalf*128 + (row%8)*16
    unsigned char* t = augbase + (long long)cb * 4096 + (row >> 3) * 256 + (row & 7) * 16;
    __align__(16) __half lo[8] = {h1, h2, h3, __float2half_rn(0.f), __float2half_rn(0.f), __float2half_rn(0.f),
                                  __float2half_rn(0.f), __float2half_rn(0.f)};
    *reinterpret_cast<uint4*>(t) = *reinterpret_cast<const uint4*>(lo);
    *reinterpret_cast<uint4*>(t + 128) = make_uint4(0u, 0u, 0u, 0u);
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    hdr->scale = s;
    hdr->max_enorm = __uint_as_float(hdr->max_enorm_bits);
    hdr->aug_c = c;
  }
}

// rounding error of the fp16 codebook operand, per code, exact: one warp per code -> header max
__global__ void __launch_bounds__(256) codebook_rounding_error_kernel(const float* __restrict__ E, int K, int D,
                                                                      unsigned char* __restrict__ blob) {
  BlobHeader* hdr = reinterpret_cast<BlobHeader*>(blob);
  const int k = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (k >= K) return;
  const float s = hdr->scale;
  float acc = 0.f;
  for (int d = lane; d < D; d += 32) {
    const float v = -2.f * s * E[(long long)k * D + d];
    const float df = __half2float(__float2half_rn(v)) - v;
    acc = fmaf(df, df, acc);
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) atomicMax(&hdr->max_de2_bits, __float_as_uint(acc * 1.0001f));
}

// ---- the kernel -----------------------------------------------------------------------------------
// MODE 0: pixel-contiguous float4 loads (NCHW), 1: dim-contiguous float4 (packed rows), 2: scalar, any strides
template <int MODE>
__global__ void __launch_bounds__(kTcThreads, 1) assign_tc_kernel(TcArgs a) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // SWIZZLE_128B needs 1024-B tiles
  const uint32_t sbase = smem_u32(smem);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const BlobHeader* hdr = reinterpret_cast<const BlobHeader*>(a.blob);
  const float scale = hdr->scale;

  const uint32_t bar_full = sbase + TcSmem::off_bar;                 // [kStages]
  const uint32_t bar_empty = bar_full + 8 * kStages;                 // [kStages]
  const uint32_t bar_tfull = bar_empty + 8 * kStages;                // [2]
  const uint32_t bar_tempty = bar_tfull + 16;                        // [2]
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + TcSmem::off_tmem);
  float* xsq = reinterpret_cast<float*>(smem + TcSmem::off_xsq);

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) { mbar_init(bar_full + 8 * s, kProducerWarps + 1); mbar_init(bar_empty + 8 * s, 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(bar_tfull + 8 * b, 1); mbar_init(bar_tempty + 8 * b, 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 9) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32((const void*)tmem_slot)), "r"(512u));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int my_tiles = a.n_tiles > (int)blockIdx.x ? (a.n_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
  const int my_units = my_tiles * a.n_cc;

  if (warp < kProducerWarps) {
    // ================= A producers =================
    // lane = quad*8 + g8: dims 8*g8 .. 8*g8+7 of the 64-dim chunk, rows 16*warp + 4*quad + i (i < 4).
    // The 8 lanes of one STS.128 phase share their rows and differ in g8, so the swizzled 16-byte
    // chunks (g8 ^ (row & 7)) land in 8 distinct bank groups: conflict-free with static indexing.
    // Loads run one chunk ahead of the convert/store (2 register buffers) to keep HBM busy.
    const int g8 = lane & 7, quad = lane >> 3;
    const int r0 = 16 * warp + 4 * quad;
    const int total = my_tiles * a.n_cc * a.n_dc;
    const int per_tile = a.n_cc * a.n_dc;
    const int D = (int)a.x.D;
    constexpr int mode = MODE;

    struct LoadState { int it, t, rem, dc; const float* rp[4]; bool rv[4]; };
    LoadState ls;
    ls.it = 0; ls.t = -1; ls.rem = 0; ls.dc = 0;
    auto load_chunk = [&](float (&v)[4][8]) {
      if (ls.it >= total) return;
      if (ls.rem == 0) {           // first chunk of a new tile: decode the 4 rows of this lane
        ls.t += 1; ls.rem = per_tile; ls.dc = 0;
        const long long n0 = ((long long)blockIdx.x + (long long)ls.t * gridDim.x) * kTileM + r0;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          ls.rv[i] = n0 + i < a.n_rows;
          ls.rp[i] = a.x.row(ls.rv[i] ? n0 + i : 0);
        }
      }
      const int d0 = ls.dc * kDChunk + 8 * g8;
      if (mode == 0) {
        const float* p0 = ls.rp[0] + (long long)d0 * a.x.sD;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float4 f = make_float4(0.f, 0.f, 0.f, 0.f);
          if (ls.rv[0] && d0 + j < D) f = ldg_stream_f4(p0 + (long long)j * a.x.sD);
          v[0][j] = f.x; v[1][j] = f.y; v[2][j] = f.z; v[3][j] = f.w;
        }
      } else if (mode == 1) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          float4 f0 = make_float4(0.f, 0.f, 0.f, 0.f), f1 = f0;
          if (ls.rv[i] && d0 + 4 <= D) f0 = ldg_stream_f4(ls.rp[i] + d0);
          if (ls.rv[i] && d0 + 8 <= D) f1 = ldg_stream_f4(ls.rp[i] + d0 + 4);
          v[i][0] = f0.x; v[i][1] = f0.y; v[i][2] = f0.z; v[i][3] = f0.w;
          v[i][4] = f1.x; v[i][5] = f1.y; v[i][6] = f1.z; v[i][7] = f1.w;
        }
      } else {
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 8; ++j)
            v[i][j] = (ls.rv[i] && d0 + j < D) ? ldg_stream_f1(ls.rp[i] + (long long)(d0 + j) * a.x.sD) : 0.f;
      }
      ls.it += 1; ls.rem -= 1;
      ls.dc = (ls.dc + 1 == a.n_dc) ? 0 : ls.dc + 1;
    };

    int pit = 0, p_in_tile = 0, p_tile = 0;       // chunk being converted, its position in the tile
    float ss[4] = {0.f, 0.f, 0.f, 0.f}, sd[4] = {0.f, 0.f, 0.f, 0.f};
    auto store_chunk = [&](float (&v)[4][8]) {
      if (pit >= total) return;
      const int s = pit & (kStages - 1);
      const uint32_t ph = ((uint32_t)pit / kStages) & 1;
      const bool first_pass = p_in_tile < a.n_dc;           // code chunk 0: accumulate |x|^2
      if (warp == 0) VQ_TRACE(0, 2 * pit);
      mbar_wait(bar_empty + 8 * s, ph ^ 1);
      if (warp == 0) VQ_TRACE(0, 2 * pit + 1);
      unsigned char* at = smem + TcSmem::off_a + s * kAStageBytes;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int r = r0 + i;
        uint32_t pk[4];
#pragma unroll
        for (int j = 0; j < 8; j += 2) {
          __half2 h = __floats2half2_rn(v[i][j], v[i][j + 1]);
          if (first_pass) {
            const float2 hb = __half22float2(h);
            const float d0 = hb.x - v[i][j], d1 = hb.y - v[i][j + 1];
            ss[i] = fmaf(v[i][j], v[i][j], ss[i]); ss[i] = fmaf(v[i][j + 1], v[i][j + 1], ss[i]);
            sd[i] = fmaf(d0, d0, sd[i]); sd[i] = fmaf(d1, d1, sd[i]);
          }
          pk[j >> 1] = *reinterpret_cast<uint32_t*>(&h);
        }
        *reinterpret_cast<uint4*>(at + r * 128 + ((g8 ^ (r & 7)) * 16)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
      }
      if (p_in_tile == a.n_dc - 1) {                         // row norms complete: publish before the arrive
        float2* xs = reinterpret_cast<float2*>(xsq) + (p_tile & (kXsqBufs - 1)) * kTileM;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          float v2 = ss[i], e2 = sd[i];
          v2 += __shfl_xor_sync(0xffffffffu, v2, 1); e2 += __shfl_xor_sync(0xffffffffu, e2, 1);
          v2 += __shfl_xor_sync(0xffffffffu, v2, 2); e2 += __shfl_xor_sync(0xffffffffu, e2, 2);
          v2 += __shfl_xor_sync(0xffffffffu, v2, 4); e2 += __shfl_xor_sync(0xffffffffu, e2, 4);
          // |x|^2 and |fp16(x) - x|^2 (inf when an element overflows fp16: the row then goes to the exact pass)
          if (g8 == 0) xs[r0 + i] = make_float2(v2, e2);
          ss[i] = 0.f; sd[i] = 0.f;
        }
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive_relaxed(bar_full + 8 * s);
      pit += 1;
      if (++p_in_tile == per_tile) { p_in_tile = 0; p_tile += 1; }
    };

    float va[4][8], vb[4][8];
    load_chunk(va);
    while (pit < total) {
      load_chunk(vb); store_chunk(va);
      load_chunk(va); store_chunk(vb);
    }
  } else if (warp == 8) {
    // ================= B loader (bulk async copies) =================
    if (lane == 0) {
      const unsigned char* img = a.blob + hdr->off_image;
      uint32_t it = 0;
      for (int t = 0; t < my_tiles; ++t)
        for (int cc = 0; cc < a.n_cc; ++cc)
          for (int dc = 0; dc < a.n_dc; ++dc, ++it) {
            const int s = it % kStages;
            const uint32_t ph = (it / kStages) & 1;
            VQ_TRACE(3, 2 * it);
            mbar_wait(bar_empty + 8 * s, ph ^ 1);
            VQ_TRACE(3, 2 * it + 1);
            const uint32_t dst = sbase + TcSmem::off_b + s * kBStageBytes;
            mbar_arrive_expect_tx(bar_full + 8 * s, kBStageBytes);
            bulk_g2s(dst, img + ((long long)(2 * cc) * a.n_dc + dc) * kTileBytes, kTileBytes, bar_full + 8 * s);
            bulk_g2s(dst + kTileBytes, img + ((long long)(2 * cc + 1) * a.n_dc + dc) * kTileBytes, kTileBytes, bar_full + 8 * s);
          }
    }
  } else if (warp == 9) {
    // ================= MMA issuer =================
    uint32_t it = 0;
    for (int u = 0; u < my_units; ++u) {
      const int buf = u & 1;
      const uint32_t use = (uint32_t)(u >> 1);
      VQ_TRACE(1, 128 + 2 * u);
      mbar_wait(bar_tempty + 8 * buf, use & 1);          // epilogue drained + re-initialised this buffer
      VQ_TRACE(1, 128 + 2 * u + 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + buf * kUnitN;
      for (int dc = 0; dc < a.n_dc; ++dc, ++it) {
        const int s = it % kStages;
        const uint32_t ph = (it / kStages) & 1;
        VQ_TRACE(1, 2 * it);
        mbar_wait(bar_full + 8 * s, ph);
        VQ_TRACE(1, 2 * it + 1);
        tc_fence_after();
        if (lane == 0) {
          const uint64_t ad = make_desc(sbase + TcSmem::off_a + s * kAStageBytes);
          const uint64_t bd = make_desc(sbase + TcSmem::off_b + s * kBStageBytes);
#pragma unroll
          for (int k = 0; k < kDChunk / 16; ++k)          // +32 B per K=16 step inside the 128-B swizzle row
            tc_mma_f16(d_tmem, ad + (uint64_t)(2 * k), bd + (uint64_t)(2 * k), kIdesc, 1u);
          tc_commit(bar_empty + 8 * s);                   // smem stage free when these MMAs retire
          if (dc == a.n_dc - 1) tc_commit(bar_tfull + 8 * buf);
        }
        __syncwarp();
      }
    }
  } else {
    // ================= epilogue =================
    const int quarter = warp & 3;                         // TMEM lane quarter this warp may access
    const int r = quarter * 32 + lane;                    // row within the tile == TMEM lane
    const uint32_t lane_addr = tmem_base + ((uint32_t)(quarter * 32) << 16);
    const float* enorm_g = reinterpret_cast<const float*>(a.blob + hdr->off_enorm) + hdr->K_pad;
    const float* enorm_sm = reinterpret_cast<const float*>(smem + TcSmem::off_enorm);
    const bool en_smem = hdr->K_pad <= kEnormSmem;
    int* cand = reinterpret_cast<int*>(smem + TcSmem::off_cand) + r * kCandCap;
    const float emax = sqrtf(hdr->max_enorm) * 1.0001f;
    const float e_s = emax * scale;
    const float de_max = sqrtf(__uint_as_float(hdr->max_de2_bits)) * 1.0001f;

    if (en_smem) {                                        // s*|e_k|^2 for every code, staged once per CTA
      float* dst = reinterpret_cast<float*>(smem + TcSmem::off_enorm);
      for (int i = (warp - kEpiWarp0) * 32 + lane; i < hdr->K_pad; i += 128) dst[i] = __ldg(enorm_g + i);
      asm volatile("bar.sync 1, 128;" ::: "memory");      // epilogue warps only
    }

    auto init_buf = [&](int buf, int cc) {
      const float* en = (en_smem ? enorm_sm : enorm_g) + cc * kUnitN;
#pragma unroll 1
      for (int c = 0; c < kUnitN; c += 32) {
        uint32_t vals[32];
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          float4 e4 = *reinterpret_cast<const float4*>(en + c + j);
          vals[j] = __float_as_uint(e4.x); vals[j + 1] = __float_as_uint(e4.y);
          vals[j + 2] = __float_as_uint(e4.z); vals[j + 3] = __float_as_uint(e4.w);
        }
        tmem_st32(lane_addr + buf * kUnitN + c, vals);
      }
      tmem_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_tempty + 8 * buf);
    };
    init_buf(0, 0);
    init_buf(1, 1 % a.n_cc);

    int u = 0;
    for (int t = 0; t < my_tiles; ++t) {
      const long long n = ((long long)blockIdx.x + (long long)t * gridDim.x) * kTileM + r;
      float m_run = __int_as_float(0x7f800000);
      float slack = 0.f;
      int cnt = 0;
      bool overflow = false;
      for (int cc = 0; cc < a.n_cc; ++cc, ++u) {
        const int buf = u & 1;
        if (quarter == 0) VQ_TRACE(2, 4 * u);
        mbar_wait(bar_tfull + 8 * buf, (uint32_t)(u >> 1) & 1);
        tc_fence_after();
        if (quarter == 0) VQ_TRACE(2, 4 * u + 1);
        if (cc == 0) {
          // all A chunks of this tile's first unit are in: the row's sum of x^2 has been published
          const float2 nr = reinterpret_cast<const float2*>(xsq)[(t & (kXsqBufs - 1)) * kTileM + r];
          const float xn = sqrtf(nr.x) * 1.0001f, dn = sqrtf(nr.y) * 1.0001f;
          const float sum = xn + emax;
          // |approx - exact| <= |dx| |e^| + |x| |de|  (Cauchy-Schwarz on the ACTUAL operand rounding errors,
          // both measured exactly: dx by the producers, de by codebook_rounding_error_kernel), two-sided, plus
          // the fp32 accumulation error of the tensor core and of the exact scorer's chain (scaled domain)
          slack = 2.002f * (dn * 2.002f * e_s + xn * de_max) + scale * (float)(a.x.D + 8) * 2.4e-7f * sum * sum;
          if (!(slack < 3.0e38f)) overflow = true;                                   // fp16 overflow in this row
        }
        const uint32_t tb = lane_addr + buf * kUnitN;
        // pass A: minimum of this unit's 256 scores
        float m_u = __int_as_float(0x7f800000);
#pragma unroll 1
        for (int c = 0; c < kUnitN; c += 32) {
          uint32_t v[32];
          tmem_ld32(tb + c, v);
          float m0 = m_u, m1 = __uint_as_float(v[1]);
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            m0 = fminf(m0, fminf(__uint_as_float(v[j]), __uint_as_float(v[j + 2])));
            m1 = fminf(m1, fminf(__uint_as_float(v[j + 1]), __uint_as_float(v[j + 3])));
          }
          m_u = fminf(m0, m1);
        }
        // every earlier short-list entry scored >= the old minimum: if that is now out of range, drop them all
        const float m_new = fminf(m_run, m_u);
        if (m_run > m_new + slack) cnt = 0;
        m_run = m_new;
        const float thr = m_run + slack;
        if (quarter == 0) VQ_TRACE(2, 4 * u + 2);
        // pass B: every score within the bound joins the short-list (indices only).  tcgen05.ld is
        // warp-collective (.sync.aligned): every lane runs the loop; the bit mask keeps the hot path
        // branch-free, the append loop runs only for lanes that found something.
#pragma unroll 1
        for (int c = 0; c < kUnitN; c += 32) {
          uint32_t v[32];
          tmem_ld32(tb + c, v);
          // bit (31 - j) of ~mk <=> v[j] <= thr: one FADD + one funnel shift per score (the sign of
          // thr - v[j] is shifted in), two independent chains
          uint32_t mka = 0u, mkb = 0u;
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const uint32_t da = __float_as_uint(thr - __uint_as_float(v[j]));
            const uint32_t db = __float_as_uint(thr - __uint_as_float(v[j + 16]));
            mka = __funnelshift_l(da, mka, 1);
            mkb = __funnelshift_l(db, mkb, 1);
          }
          uint32_t mk = ~((mka << 16) | (mkb & 0xffffu));
          if (overflow) mk = 0u;
          while (mk) {
            const int j = __clz(mk);
            mk &= ~(0x80000000u >> j);
            if (cnt < kCandCap) cand[cnt++] = cc * kUnitN + c + j;
            else { overflow = true; mk = 0u; }
          }
        }
        if (quarter == 0) VQ_TRACE(2, 4 * u + 3);
        init_buf(buf, (u + 2) % a.n_cc);                  // hand the buffer back, pre-loaded for unit u+2
      }
      // ---- tile done: resolve rows ----
      const bool in_range = n < a.n_rows;
      const int last = cnt > 0 ? cand[cnt - 1] : 0;
      const bool unique = !overflow && cnt == 1 && !a.force_rescore && last < a.K;
      if (in_range && unique) {
        a.idx_out[n] = (long long)last + a.code_base;
        if (a.counts_out) atomicAdd(a.counts_out + last, 1ull);
      }
      const bool flagged = in_range && !unique;
      const uint32_t fm = __ballot_sync(0xffffffffu, flagged);
      if (fm) {                                            // one atomic per warp on the work counter
        int base = 0;
        if (lane == 0) base = atomicAdd(a.work_count, __popc(fm));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (flagged) {
          int nc = 0;
          if (!overflow)
            for (int e = 0; e < cnt; ++e) {
              int k = cand[e];
              if (k < a.K) a.cand_idx[n * kCandCap + nc++] = k;
            }
          a.cand_cnt[n] = (overflow || nc == 0) ? kCandCap + 1 : nc;     // > cap => exact pass scans all codes
          a.work_rows[base + __popc(fm & ((1u << lane) - 1))] = (int)n;
        }
      }
    }
  }

  // ---- teardown ----
  if ((warp & 3) == 0 || warp == 9 || warp == 8) VQ_TRACE(warp < 8 ? 0 : (warp == 9 ? 1 : (warp == 8 ? 3 : 2)), 255);
  tc_fence_before();
  __syncthreads();
  if (warp == 9) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u));
  }
}

int launch_pack(const float* E, int K, int D, unsigned char* blob, cudaStream_t st) {
  int blocks = num_sms() * 2;
  pack_codebook_kernel<<<blocks, 256, 0, st>>>(E, K, D, blob);
  VQSEG_LAUNCH_CHECK();
  codebook_rounding_error_kernel<<<(K * 32 + 255) / 256, 256, 0, st>>>(E, K, D, blob);
  VQSEG_LAUNCH_CHECK();
  return 0;
}

template <int MODE>
static int launch_mode(const TcArgs& a, int grid, cudaStream_t st) {
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(assign_tc_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, TcSmem::total);
    if (e != cudaSuccess) return (int)e;
    configured = true;
  }
  assign_tc_kernel<MODE><<<grid, kTcThreads, TcSmem::total, st>>>(a);
  VQSEG_LAUNCH_CHECK();
  return 0;
}

int launch_assign_tc(const TcArgs& a, cudaStream_t st) {
  int grid = a.n_tiles < num_sms() ? a.n_tiles : num_sms();
  if (grid <= 0) return 0;
  const bool al16 = (reinterpret_cast<uintptr_t>(a.x.ptr) & 15) == 0 && (a.x.sB & 3) == 0;
  if (al16 && a.x.sP == 1 && (a.x.P & 3) == 0 && (a.x.sD & 3) == 0) return launch_mode<0>(a, grid, st);
  if (al16 && a.x.sD == 1 && (a.x.D & 3) == 0 && (a.x.sP & 3) == 0) return launch_mode<1>(a, grid, st);
  return launch_mode<2>(a, grid, st);
}

}  // namespace vqseg
