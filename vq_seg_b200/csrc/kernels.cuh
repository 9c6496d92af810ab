// Launch interface between the C ABI layer (api.cu) and the kernel translation units: argument structs and
// launchers, defined once here.
#pragma once
#include "common.cuh"

namespace vqseg {

// ---- exact.cu: exact fp32 scorer (brute force over all codes, or the rescoring pass over the filter's short-lists)
struct ExactArgs {
  Rows x;
  const float* E; int K;
  const float* enorm;
  int kblock;
  // candidate mode (null -> all rows x all codes)
  const int* work_rows; const int* work_count;        // flagged row ids, device counter
  const int* cand_idx; const int* cand_cnt; int cand_cap;   // per row: up to cand_cap codes; cnt > cap => all codes
  long long* idx_out; unsigned long long* counts_out; unsigned long long* key_out; long long code_base;
  int stage_e;
  long long* trace;     // dev tool: [0] min start ns, [1] max end ns, [2..] per-phase clock sums
  int* done_blocks;     // ticket counter (zeroed by the host) for the fused usage epilogue
  float* usage_out;     // nullable: the last block to finish writes 100 * (#counts == 0) / K   (vq_img.py:174-175)
};
int launch_enorm(const float* E, int K, int D, int K_pad, float* enorm, BlobHeader* hdr, cudaStream_t st);
int launch_exact(const ExactArgs& a, long long max_work, cudaStream_t st);

// ---- assign_tc.cu: single-CTA streaming tcgen05 filter, and the codebook packer
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
int launch_pack(const float* E, int K, int D, unsigned char* blob, cudaStream_t st);
int launch_assign_tc(const TcArgs& a, cudaStream_t st);

// ---- assign_tc2.cu: CTA-pair, codebook-resident tcgen05 filter
struct Tc2Args {
  Rows x;
  const unsigned char* blob;
  long long n_rows;
  int n_ptiles, n_cc, n_dc;       // pair tiles of 256 rows, code chunks of 256, dim chunks of 64
  int K, K_pad;
  unsigned long long off_image, off_aug, off_enorm;
  float tau;
  long long* idx_out; unsigned long long* counts_out; long long code_base;
  int force_rescore;
  int* cand_idx; int* cand_cnt; int* work_rows; int* work_count;
  long long* trace;
};
bool tc2_supported(int n_cc, int n_dc);
int launch_assign_tc2(const Tc2Args& a, cudaStream_t st);

}  // namespace vqseg
