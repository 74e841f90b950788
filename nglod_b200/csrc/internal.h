// nglod_b200 -- internal (non-ABI) launchers shared between translation units.
#pragma once
#include "common.cuh"

int nglod_launch_sdf_forward_tc(const NetDev& nd, const float* x, long long n, float* out, cudaStream_t st);
