// C ABI entry points for codebook preparation and nearest-code assignment (see include/vqseg.h).
#include "common.cuh"
#include "kernels.cuh"
#include <string.h>

namespace vqseg {
constexpr int kCandCapHost = 8;

__global__ void blob_init_kernel(BlobHeader* h, int K, int D, int K_pad, int D_pad, unsigned long long off_enorm,
                                 unsigned long long off_image, unsigned long long off_aug) {
  h->off_aug = off_aug; h->aug_c = 1.f; h->flags = 0u; h->max_de2_bits = 0u;
  h->magic = kBlobMagic; h->K = K; h->D = D; h->K_pad = K_pad; h->D_pad = D_pad;
  h->scale = 1.f; h->max_enorm = 0.f; h->max_enorm_bits = 0u; h->max_abs_bits = 0u;
  h->off_enorm = off_enorm; h->off_image = off_image;
}

extern "C" int vqseg_internal_gather_ticket(const float* x, int64_t B, int64_t P, int64_t D, int64_t sB, int64_t sP, int64_t sD,
                                            const float* E, int64_t K, const int64_t* idx,
                                            float* q_out, int64_t qB, int64_t qP, int64_t qD, float* loss_out, int mode,
                                            void* ws, size_t ws_bytes, void* stream, int* ticket);

// one launch instead of three memset nodes at the head of the fused forward: per-code counts, the loss scalar and
// the assignment's {work counter, block ticket}
__global__ void forward_zero_kernel(unsigned long long* counts, int K, float* loss, int* work2) {
  for (int k = threadIdx.x; k < K; k += blockDim.x) counts[k] = 0ull;
  if (threadIdx.x == 0) { if (loss) *loss = 0.f; work2[0] = 0; work2[1] = 0; work2[2] = 0; }   // [2]: gather's loss ticket
}

// MKL's sgemm K-blocking as probed on the reference CPU path: one chain up to
// 384 terms, two halves up to 768, 384-blocks beyond.
static int auto_kblock(long long D) {
  long long L = D + 2;
  if (L <= 384) return 0;
  if (L <= 768) return (int)((L + 1) / 2);
  return 384;
}
}  // namespace vqseg

using namespace vqseg;

static int g_timing = 0;
static long long* g_trace = nullptr;
static int g_force_tc1 = 0;
static cudaEvent_t g_ev[4] = {nullptr, nullptr, nullptr, nullptr};
static int g_ev_valid[2] = {0, 0};
static void ev_record(int i, cudaStream_t st) {
  if (!g_timing || !g_ev[i]) return;
  if (cudaEventRecord(g_ev[i], st) != cudaSuccess) (void)cudaGetLastError();   // never poison the launch checks
}

extern "C" {

int vqseg_version(void) { return VQSEG_VERSION; }

void vqseg_debug_set_trace(void* dev_buf) { g_trace = (long long*)dev_buf; }
void vqseg_debug_force_streaming_kernel(int on) { g_force_tc1 = on; }

void vqseg_set_kernel_timing(int enable) {
  g_timing = enable; g_ev_valid[0] = g_ev_valid[1] = 0;
  if (enable)                               // created here, outside any stream capture
    for (int i = 0; i < 4; ++i)
      if (!g_ev[i] && cudaEventCreate(&g_ev[i]) != cudaSuccess) { (void)cudaGetLastError(); g_ev[i] = nullptr; }
}

float vqseg_get_kernel_timing_ms(int which) {
  if (which < 0 || which > 1 || !g_ev_valid[which]) return -1.f;
  float ms = -1.f;
  if (!g_ev[2 * which] || !g_ev[2 * which + 1]) return -1.f;
  if (cudaEventSynchronize(g_ev[2 * which + 1]) != cudaSuccess) { (void)cudaGetLastError(); return -1.f; }
  if (cudaEventElapsedTime(&ms, g_ev[2 * which], g_ev[2 * which + 1]) != cudaSuccess) { (void)cudaGetLastError(); return -1.f; }
  return ms;
}

const char* vqseg_error_string(int code) {
  switch (code) {
    case 0: return "ok";
    case VQSEG_EINVAL: return "vqseg: invalid argument";
    case VQSEG_EWORKSPACE: return "vqseg: workspace too small";
    case VQSEG_EUNSUPPORTED: return "vqseg: shape not supported by the requested algorithm";
    case VQSEG_EARCH: return "vqseg: device is not sm_100 (B200)";
  }
  if (code > 0) return cudaGetErrorString((cudaError_t)code);
  return "vqseg: unknown error";
}

static void blob_geometry(int64_t K, int64_t D, long long* K_pad, long long* D_pad, size_t* off_enorm, size_t* off_image,
                          size_t* total, size_t* off_aug = nullptr) {
  *K_pad = round_up(K, 256);
  *D_pad = round_up(D, kDChunk);
  *off_enorm = 1024;
  *off_image = *off_enorm + (size_t)round_up(2 * *K_pad * sizeof(float), 1024);
  size_t aug = *off_image + (size_t)(*K_pad / kCodeBlock) * (size_t)(*D_pad / kDChunk) * kTileBytes;
  if (off_aug) *off_aug = aug;
  *total = aug + (size_t)(*K_pad / kCodeBlock) * 4096;
}

size_t vqseg_codebook_blob_bytes(int64_t K, int64_t D) {
  if (K <= 0 || D <= 0) return 0;
  long long kp, dp; size_t oe, oi, tot;
  blob_geometry(K, D, &kp, &dp, &oe, &oi, &tot);
  return tot;
}

int vqseg_codebook_prepare_f32(const float* E, int64_t K, int64_t D, void* blob, size_t blob_bytes, void* stream) {
  if (!E || !blob || K <= 0 || D <= 0 || K >= (1ll << 30) || D >= (1ll << 20)) return VQSEG_EINVAL;
  long long kp, dp; size_t oe, oi, tot, oa;
  blob_geometry(K, D, &kp, &dp, &oe, &oi, &tot, &oa);
  if (blob_bytes < tot) return VQSEG_EWORKSPACE;
  if ((reinterpret_cast<uintptr_t>(blob) & 1023) != 0) return VQSEG_EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
  unsigned char* b = (unsigned char*)blob;
  blob_init_kernel<<<1, 1, 0, st>>>((BlobHeader*)b, (int)K, (int)D, (int)kp, (int)dp, oe, oi, oa);
  VQSEG_LAUNCH_CHECK();
  int rc = launch_enorm(E, (int)K, (int)D, (int)kp, (float*)(b + oe), (BlobHeader*)b, st);
  if (rc) return rc;
  return launch_pack(E, (int)K, (int)D, b, st);
}

size_t vqseg_assign_workspace_bytes(int64_t n_rows, int64_t D, int64_t K, int algo) {
  (void)D; (void)algo;
  size_t b = 256;                                              // work counter
  b += (size_t)round_up(n_rows * sizeof(int), 256);            // work_rows
  b += (size_t)round_up(n_rows * sizeof(int), 256);            // cand_cnt
  b += (size_t)round_up(n_rows * kCandCapHost * sizeof(int), 256);   // cand_idx
  b += (size_t)round_up(round_up(K, 256) * sizeof(float), 256);      // enorm when no blob is given
  return b;
}

static int assign_internal(const float* x, int64_t B, int64_t P, int64_t D, int64_t sB, int64_t sP, int64_t sD,
                           const float* E, int64_t K, const void* blob,
                           int64_t* idx_out, int64_t* counts_out, uint64_t* best_key_out,
                           int64_t code_base, int kblock, int algo, void* ws, size_t ws_bytes, void* stream,
                           float* usage_out, float* zero_loss = nullptr, bool zero_outputs = false);

int vqseg_assign_f32(const float* x, int64_t B, int64_t P, int64_t D, int64_t sB, int64_t sP, int64_t sD,
                     const float* E, int64_t K, const void* blob,
                     int64_t* idx_out, int64_t* counts_out, uint64_t* best_key_out,
                     int64_t code_base, int kblock, int algo, void* ws, size_t ws_bytes, void* stream) {
  return assign_internal(x, B, P, D, sB, sP, sD, E, K, blob, idx_out, counts_out, best_key_out, code_base, kblock, algo,
                         ws, ws_bytes, stream, nullptr);
}

static int assign_internal(const float* x, int64_t B, int64_t P, int64_t D, int64_t sB, int64_t sP, int64_t sD,
                           const float* E, int64_t K, const void* blob,
                           int64_t* idx_out, int64_t* counts_out, uint64_t* best_key_out,
                           int64_t code_base, int kblock, int algo, void* ws, size_t ws_bytes, void* stream,
                           float* usage_out, float* zero_loss, bool zero_outputs) {
  if (!x || !E || B < 0 || P < 0 || D <= 0 || K <= 0) return VQSEG_EINVAL;
  if (!idx_out && !best_key_out) return VQSEG_EINVAL;
  const long long n_rows = B * P;
  if (n_rows == 0) return 0;
  if (n_rows >= (1ll << 31) || K >= (1ll << 31) - 1 || D >= (1ll << 20)) return VQSEG_EUNSUPPORTED;
  if (!ws || ws_bytes < vqseg_assign_workspace_bytes(n_rows, D, K, algo)) return VQSEG_EWORKSPACE;
  int rc = check_arch();
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  if (kblock == 0) kblock = auto_kblock(D);

  char* p = (char*)ws;
  int* work_count = (int*)p;   p += 256;
  int* work_rows = (int*)p;    p += round_up(n_rows * sizeof(int), 256);
  int* cand_cnt = (int*)p;     p += round_up(n_rows * sizeof(int), 256);
  int* cand_idx = (int*)p;     p += round_up(n_rows * kCandCapHost * sizeof(int), 256);
  float* enorm_ws = (float*)p;

  const BlobHeader* hdr = (const BlobHeader*)blob;
  const float* enorm = nullptr;
  long long kp = 0, dp = 0;
  size_t oe = 0, oi = 0, tot = 0, oa = 0;
  if (blob) {
    blob_geometry(K, D, &kp, &dp, &oe, &oi, &tot, &oa);
    enorm = (const float*)((const char*)blob + oe);
  } else {
    rc = launch_enorm(E, (int)K, (int)D, (int)K, enorm_ws, nullptr, st);
    if (rc) return rc;
    enorm = enorm_ws;
  }
  (void)hdr;

  bool use_tc = false;
  if (algo == VQSEG_ALGO_TC) {
    if (!blob) return VQSEG_EINVAL;
    use_tc = true;
  } else if (algo == VQSEG_ALGO_AUTO) {
    use_tc = blob != nullptr && n_rows >= 64 && K >= 32;
  } else if (algo != VQSEG_ALGO_EXACT) {
    return VQSEG_EINVAL;
  }

  Rows xr{x, B, P, D, sB, sP, sD};
  ExactArgs ea;
  memset(&ea, 0, sizeof(ea));
  ea.x = xr; ea.E = E; ea.K = (int)K; ea.enorm = enorm; ea.kblock = kblock;
  ea.idx_out = (long long*)idx_out; ea.counts_out = (unsigned long long*)counts_out;
  ea.key_out = (unsigned long long*)best_key_out; ea.code_base = code_base;
  ea.done_blocks = work_count + 1;
  ea.usage_out = counts_out ? usage_out : nullptr;
  if (zero_outputs) {                                            // fused forward: counts + loss + work counter + ticket
    forward_zero_kernel<<<1, 256, 0, st>>>((unsigned long long*)counts_out, (int)K, zero_loss, work_count);
    VQSEG_LAUNCH_CHECK();
  } else {
    cudaError_t e0 = cudaMemsetAsync(work_count, 0, 2 * sizeof(int), st);      // work counter + block ticket
    if (e0 != cudaSuccess) return (int)e0;
  }

  if (!use_tc) return launch_exact(ea, n_rows, st);

  const float tau = 0.00390625f * 1.015625f;   // 2^-8 (two fp16 roundings per operand pair, both sides) + margin
  const int n_cc = (int)(kp / 256), n_dc = (int)(dp / kDChunk);
  const bool force = (best_key_out != nullptr || idx_out == nullptr);
  ev_record(0, st);
  if (tc2_supported(n_cc, n_dc) && g_force_tc1 == 0) {
    Tc2Args t2;
    memset(&t2, 0, sizeof(t2));
    t2.x = xr; t2.blob = (const unsigned char*)blob; t2.n_rows = n_rows;
    t2.n_ptiles = (int)((n_rows + 255) / 256); t2.n_cc = n_cc; t2.n_dc = n_dc;
    t2.K = (int)K; t2.K_pad = (int)kp; t2.off_image = oi; t2.off_aug = oa; t2.off_enorm = oe;
    t2.tau = tau;
    t2.idx_out = (long long*)idx_out; t2.counts_out = (unsigned long long*)counts_out; t2.code_base = code_base;
    t2.force_rescore = force ? 1 : 0;
    t2.cand_idx = cand_idx; t2.cand_cnt = cand_cnt; t2.work_rows = work_rows; t2.work_count = work_count;
    t2.trace = g_trace;
    rc = launch_assign_tc2(t2, st);
  } else {
    TcArgs ta;
    memset(&ta, 0, sizeof(ta));
    ta.x = xr; ta.blob = (const unsigned char*)blob; ta.n_rows = n_rows;
    ta.n_tiles = (int)((n_rows + 127) / 128); ta.n_cc = n_cc; ta.n_dc = n_dc;
    ta.K = (int)K;
    ta.tau = tau;
    ta.idx_out = (long long*)idx_out; ta.counts_out = (unsigned long long*)counts_out; ta.code_base = code_base;
    ta.force_rescore = force ? 1 : 0;
    ta.cand_idx = cand_idx; ta.cand_cnt = cand_cnt; ta.work_rows = work_rows; ta.work_count = work_count;
    ta.trace = g_trace;
    rc = launch_assign_tc(ta, st);
  }
  ev_record(1, st);
  if (rc) return rc;
  if (g_timing) g_ev_valid[0] = 1;
  ea.work_rows = work_rows; ea.work_count = work_count;
  ea.cand_idx = cand_idx; ea.cand_cnt = cand_cnt; ea.cand_cap = kCandCapHost;
  ea.trace = g_trace ? g_trace + 148 * 4 * 256 : nullptr;      // dev tool: 8 int64 after the filter's trace area
  ev_record(2, st);
  rc = launch_exact(ea, n_rows, st);
  ev_record(3, st);
  if (g_timing) g_ev_valid[1] = 1;
  return rc;
}

size_t vqseg_forward_workspace_bytes(int64_t n_rows, int64_t D, int64_t K) {
  return (size_t)round_up(vqseg_assign_workspace_bytes(n_rows, D, K, 0), 256) + vqseg_gather_workspace_bytes(n_rows, D);
}

int vqseg_vq_forward_f32(const float* x, int64_t B, int64_t P, int64_t D, int64_t sB, int64_t sP, int64_t sD,
                         const float* E, int64_t K, const void* blob,
                         int64_t* idx_out, int64_t* counts_out, float* usage_out,
                         float* q_out, int64_t qB, int64_t qP, int64_t qD, float* loss_out,
                         int mode, int algo, int kblock, void* ws, size_t ws_bytes, void* stream) {
  if (!counts_out || K <= 0 || B < 0 || P < 0) return VQSEG_EINVAL;
  if (B * P != 0 && (!idx_out || !q_out)) return VQSEG_EINVAL;   // (empty tensors have null data pointers)
  const size_t wa = (size_t)round_up(vqseg_assign_workspace_bytes(B * P, D, K, algo), 256);
  if (B * P != 0 && (!ws || ws_bytes < wa + vqseg_gather_workspace_bytes(B * P, D))) return VQSEG_EWORKSPACE;
  if (B * P == 0) {                                              // nothing to launch: outputs of an empty batch
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e = cudaMemsetAsync(counts_out, 0, (size_t)K * sizeof(int64_t), st);
    if (e == cudaSuccess && loss_out) e = cudaMemsetAsync(loss_out, 0, sizeof(float), st);
    if (e != cudaSuccess) return (int)e;
    return usage_out ? vqseg_code_usage(counts_out, K, usage_out, stream) : 0;     // no code is used: 100 %
  }
  int rc = assign_internal(x, B, P, D, sB, sP, sD, E, K, blob, idx_out, counts_out, nullptr, 0, kblock, algo, ws, wa, stream,
                           usage_out, loss_out, true);   // one zeroing launch; usage reduced by the exact pass's last block
  if (rc) return rc;
  // the first 256 bytes of the assignment workspace hold {work counter, block ticket, loss ticket}, all zeroed above
  return vqseg_internal_gather_ticket(x, B, P, D, sB, sP, sD, E, K, idx_out, q_out, qB, qP, qD, loss_out, mode,
                                      (char*)ws + wa, ws_bytes - wa, stream, (int*)ws + 2);
}

}  // extern "C"
