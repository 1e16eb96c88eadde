// Argument structs + launchers of the selection kernels (see select_finalize.cu).
#pragma once
#include "common.cuh"

namespace rl {

struct SelectArgs {
  const float* dump;          // [B, n_sample_rows]
  const int32_t* row_chunk;   // [n_rows]
  const float* eps;           // [B]
  float* thr;                 // [B] in/out
  Cand* cand;                 // [B, cap]
  int32_t* cand_cnt;          // [B]
  int32_t* ghist;             // [B, kHistBins]
  float* hist_inv_w;          // [B] out: 1 / histogram bin width chosen from the sample
  int64_t n_sample_rows, n_rows;
  int32_t S, cap, mode_sql, sel_k, reuse_thr;
};

struct FinalizeArgs {
  const float* E;
  const int32_t* row_chunk;
  const float* Q;             // [B, d] float32
  const double* q_sq;         // [B]
  const float* eps;           // [B]
  const Cand* cand;
  Cand* cand_rw;              // same list, writable: the streaming survivor path rescores in place
  const int32_t* cand_cnt;
  float* thr_out;             // [B] cut usable as the next emission threshold
  float* hit_sim;             // [B, H]
  int64_t* hit_chunk;         // [B, H]
  int32_t* hit_count;         // [B]
  int32_t* status;            // [B]
  int32_t* n_surv;            // [B]
  Header* header;
  int64_t ld, chunk_base, n_sample_rows;
  int32_t d, metric, cap, mode_sql, sel_k, H, launches, S, algo, e_f16, counted_unfiltered;
};

struct MergeArgs {
  const float* hit_sim;       // [R, B, H]
  const int64_t* hit_chunk;   // [R, B, H]
  const int32_t* hit_count;   // [R, B]
  float* out_sim;             // [B, k]
  int64_t* out_chunk;         // [B, k]
  int32_t* out_count;         // [B]
  int32_t R, B, H, num_hits, k;
  int32_t win;                // set by launch_merge: power of two >= R * H (or >= num_hits when prefiltering)
  int32_t prefilter;          // set by launch_merge: R * H exceeds the window, select the num_hits best from global memory first
  int64_t sim_rs, chunk_rs, count_rs;  // element strides between the ranks' lists (0 = contiguous [R, B, H] / [R, B])
};

constexpr int kFinalizeScratch = 20480;  // histogram (16 KB) / flags + positions (20 KB)

int launch_query_prep(const float* Q, int B, int d, int metric, int algo, const float* row_stats, double* q_sq,
                      float* q_inv, float* eps, cudaStream_t stream);
int launch_sim_floor_to_thr(const float* sim_floor, const double* q_sq, const float* eps, int metric, int bound, int B,
                            float* thr, cudaStream_t stream);
int launch_select(const SelectArgs& a, int B, cudaStream_t stream);
int launch_finalize(const FinalizeArgs& f, int B, cudaStream_t stream);
int launch_merge(const MergeArgs& m, cudaStream_t stream);

}  // namespace rl
