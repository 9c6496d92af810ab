// Launch interface between the C ABI layer (api.cu) and the kernel translation units: argument structs and
// launchers, defined once here.
#pragma once
#include "common.cuh"

namespace vqseg {

// One undecided row handed from a tensor-core filter to the rescoring pass: the row id and the short-list of codes
// whose approximate score lies within the proven error bound of the row minimum.  cnt > kWorkCandCap means "the
// short-list overflowed: score every code".  48 bytes, 16-byte aligned: one coalesced fetch per record.
constexpr int kWorkCandCap = 8;
struct alignas(16) WorkRec { int row; int cnt; int pad0; int pad1; int cand[kWorkCandCap]; };

// ---- exact.cu: exact fp32 scorer (brute force over all codes, or the rescoring pass over the filter's short-lists)
constexpr int kOvfSplitCap = 512;   // overflow rows up to which each row's codes are split over several blocks
struct ExactArgs {
  Rows x;
  const float* E; int K;
  const float* enorm;
  int kblock;
  int ip;               // 1: inner-product mode (cosine codebook): argmax <x, e>, plain D-term chain (exact_chain.cuh)
  // candidate mode (null -> all rows x all codes)
  const WorkRec* work; const int* work_count;         // undecided rows + device counter
  int* ovf_rows; int* ovf_count;                      // rows whose short-list overflowed: deferred to overflow_rows_kernel
  unsigned long long* ovf_keys; int* ovf_tickets;     // kOvfSplitCap packed (score, code) minima + block tickets: few overflow rows are split over blocks
  long long* idx_out; unsigned long long* counts_out; unsigned long long* key_out; long long code_base;
  long long* trace;     // dev tool: [0] min start ns, [1] max end ns
};
int launch_enorm(const float* E, int K, int D, int K_pad, float* enorm, BlobHeader* hdr, unsigned long long* hash,
                 cudaStream_t st);
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
  float slack_t2;                 // coefficient of the |x|^2-magnitude roundings in the filter's slack (common.cuh filter_slack)
  WorkRec* work; int* work_count;
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
  float slack_t2;                 // coefficient of the |x|^2-magnitude roundings in the filter's slack (common.cuh filter_slack)
  WorkRec* work; int* work_count;
  long long* trace;
};
bool tc2_supported(int n_cc, int n_dc);
int launch_assign_tc2(const Tc2Args& a, cudaStream_t st);

// ---- assign_tc3.cu: TMA-fed CTA-pair, codebook-resident tcgen05 filter (pixel-contiguous / NCHW inputs)
struct Tc3Args {
  long long B, P, D;              // logical (B, P, D) view; memory is pixel-contiguous (sP == 1)
  long long n_rows;
  const unsigned char* blob;
  int n_tiles, tiles_per_image;   // tiles of 128 pixels, never straddling an image
  int n_ptiles, n_cc, n_dc;       // pair tiles (two tiles), code chunks of 256, dim chunks of 64
  int K, K_pad;
  unsigned long long off_image, off_aug, off_enorm;
  long long* idx_out; unsigned long long* counts_out; long long code_base;
  int force_rescore;
  float slack_t2;                 // coefficient of the |x|^2-magnitude roundings in the filter's slack (common.cuh filter_slack)
  WorkRec* work; int* work_count;
  long long* trace;
};
bool tc3_supported(const Rows& x, int n_cc, int n_dc);
int launch_assign_tc3(const Rows& x, const Tc3Args& a, cudaStream_t st);

// ---- assign_tc4.cu: TMA-fed CTA-pair STREAMING tcgen05 filter (any K <= 65536, D_pad <= 512; NCHW maps or packed rows)
struct Tc4Args {
  long long B, P, D;              // logical (B, P, D) view
  long long n_rows;
  const unsigned char* blob;
  int n_tiles, tiles_per_image;   // tiles of 128 rows (NCHW: 128 pixels of one image)
  int n_ptiles, n_cc, n_dc;       // pair tiles (two tiles), code chunks of 256, dim chunks of 64 PER WORK ITEM
  int n_slices, n_dc_total;       // split-D mode: dim slices per row (1 = off), dim chunks of the whole row
  float* part_scores;             // split-D mode: [n_rows][K_pad] partial scores, zeroed by the prologue (else null)
  float* part_norms;              //               [n_rows][2] {|x|^2, |fp16(x) - x|^2}
  const unsigned char* samp_img;  // MODE 2 (prepared samples): fp16 A tiles [row tile][dim chunk] x 16 KiB
  const float2* samp_norms;       //                            [padded rows] {|x|^2, |fp16(x) - x|^2}
  int K, K_pad;
  unsigned long long off_image, off_aug, off_enorm;
  long long* idx_out; unsigned long long* counts_out; long long code_base;
  int force_rescore;
  float slack_t2;                 // coefficient of the |x|^2-magnitude roundings in the filter's slack (common.cuh filter_slack)
  WorkRec* work; int* work_count;
  long long* trace;
};
// split-D mode's second step: per row the short-list over the summed scores (same rule as the filters' epilogues)
struct ShortlistArgs {
  const float* scores; const float* norms; long long n_rows; int K, K_pad, D;
  const unsigned char* blob;
  long long* idx_out; unsigned long long* counts_out; long long code_base; int force_rescore; float slack_t2;
  WorkRec* work; int* work_count;
};
int launch_shortlist(const ShortlistArgs& a, cudaStream_t st);
int tc4_layout(const Rows& x, long long K_pad, int n_dc);   // 0: unsupported, 1: NCHW maps, 2: packed rows
int launch_assign_tc4(const Rows& x, const Tc4Args& a, int layout, cudaStream_t st);    // layout 3: prepared samples
int launch_samples_prepare(const Rows& x, long long rows_padded, int D_pad, unsigned char* img, float2* norms, cudaStream_t st);

}  // namespace vqseg
