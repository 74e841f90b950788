// nglod_b200 -- internal (non-ABI) launchers shared between translation units.
#pragma once
#include "common.cuh"

int nglod_launch_sdf_forward_tc(const NetDev& nd, const float* x, long long n, float* out, cudaStream_t st);
// backward / fused training step of ONE single-grid head on the tcgen05 tensor cores (sdf_backward_tc.cu); gd.grids[0] is
// the gradient of nd.grids[0] (may be null)
int nglod_launch_sdf_backward_tc(const NetDev& nd, const GradDev& gd, const float* x, long long n, const float* grad_out,
                                 const float* gt, float loss_scale, float* loss_out, bool fused_loss, cudaStream_t st);
