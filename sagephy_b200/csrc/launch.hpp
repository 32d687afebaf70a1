// Host-callable launchers of the kernels (one translation unit per kernel family keeps nvcc compile times parallel).
#pragma once
#include <cuda_runtime.h>
#include "kernels_args.cuh"

namespace sgp {
void launch_find_general(bool replay, bool domain, const FindArgs& a, int blocks, cudaStream_t st);
void launch_build_general(bool replay, bool domain, const BuildArgs& a, int blocks, cudaStream_t st);
void launch_bd_find(const BdArgs& a, int blocks, int blocks_per_sm, cudaStream_t st);      // blocks is sized for 4 per SM; fewer resident blocks leave room for a concurrent stream
void launch_bd_build(const BdArgs& a, int blocks, cudaStream_t st);
int launch_label_prune(LabelArgs a, int max_pool, int sm_count, cudaStream_t st);          // returns the number of kernels launched
int label_smem_max_pool();                                                                // largest tree the shared-memory label kernel takes (larger ones: in place, with LabelArgs::work)
int launch_emit_sizes(const EmitArgs& a, cudaStream_t st);                                // finishes the emission records and measures every stream; returns the number of kernels launched
int launch_emit_write(const EmitArgs& a, cudaStream_t st, cudaStream_t* side, int n_side, cudaEvent_t fork, cudaEvent_t* join);   // returns the number of kernels launched; side streams (optional) run the independent write kernels next to each other
void launch_relax_rates(bool replay, const RelaxArgs& a, int first_attempt, int only_flagged, cudaStream_t st);
void launch_relax_apply(const RelaxArgs& a, long long total_slots, cudaStream_t st);
void launch_relax_apply_per_rep(const RelaxArgs& a, cudaStream_t st);
void launch_ingest_count(const IngestArgs& g, cudaStream_t st);
void launch_ingest_fill(const IngestArgs& g, cudaStream_t st);
void launch_relax_emit(bool write, const RelaxArgs& a, cudaStream_t st);
}  // namespace sgp
