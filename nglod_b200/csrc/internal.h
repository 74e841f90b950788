// nglod_b200 -- internal (non-ABI) launchers shared between translation units.
#pragma once
#include "common.cuh"
#include "sparse_core.cuh"

int nglod_launch_sdf_forward_tc(const NetDev& nd, const float* x, long long n, float* out, cudaStream_t st);
// backward / fused training step of ONE head on the tcgen05 tensor cores (sdf_backward_tc.cu).  nd.num_lods == 1: gather
// from nd.grids[0] (the prefix-summed grid, or a one-level net), scatter into gd.grids[0] (or the private copies gd.priv);
// nd.num_lods > 1: per-LOD gather / scatter; sp != null: sparse octree model (nd ignored, the decoder is sp->sn.dec).
int nglod_launch_sdf_backward_tc(const NetDev& nd, const GradDev& gd, const float* x, long long n, const float* grad_out,
                                 const float* gt, float loss_scale, float* loss_out, bool fused_loss, cudaStream_t st,
                                 const SparseBwd* sp = nullptr);
