// TEST INFRASTRUCTURE -- python binding for the reference's OWN spc_raytrace (compiled, unmodified, from
// /root/reference/sol-renderer/include/spc/spc/spc_raytrace_cuda.cpp:141-199 + spc_raytrace_cuda_kernel.cu by
// oracle/build_ref.py).  The reference builds these files into a C++ library without a Python module; this file adds
// only the module definition.
#include <torch/extension.h>

torch::Tensor spc_raytrace(torch::Tensor octree, torch::Tensor points, torch::Tensor pyramid, torch::Tensor Org,
                           torch::Tensor Dir, unsigned int targetLevel);

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
    m.def("spc_raytrace", &spc_raytrace, "reference spc_raytrace (sol-renderer/include/spc/spc)");
}
