// engine_internal.h — engine internals shared with the tree-search translation unit (mcts.cu).  Not part of the ABI.
#pragma once
#include <cuda_runtime.h>

#include "quad_math.cuh"

int ipp_internal_step_params(const ipp_engine *e, ipp::StepParams *out);
cudaStream_t ipp_internal_stream(const ipp_engine *e);
void ipp_internal_count_launches(ipp_engine *e, int n);
int ipp_internal_fail(ipp_engine *e, int code, const char *msg);
int ipp_internal_layout(const ipp_engine *e);
int ipp_internal_shuffled_prior(ipp_engine *e, float *scale, int n, int gp_mode, float p0, float p1, uint64_t seed);  // fields.cu
