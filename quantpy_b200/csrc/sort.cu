// Keys-only ascending sort of the bootstrap distances (the `dist.sort()` of quantpy/tomography/interval.py:610).
// Plumbing, not a hot kernel: CUB's device radix sort on the caller's stream with the library's scratch pool
// (torch.sort also returns the permutation, which costs three times as much for 1e5 doubles).
#include <cub/device/device_radix_sort.cuh>

#include "../../include/quantpy_b200.h"
#include "common.cuh"

using namespace qpb;

extern "C" int qpb_sort_f64(long long n, const double* in, double* out, void* stream) {
    QPB_REQUIRE(n >= 0, "negative length");
    if (n == 0) return QPB_OK;
    QPB_REQUIRE(in && out && in != out, "sort needs distinct input and output buffers");
    QPB_REQUIRE(n < (1ll << 31), "too many keys");
    cudaStream_t st = (cudaStream_t)stream;
    size_t bytes = 0;
    QPB_CUDA(cub::DeviceRadixSort::SortKeys(nullptr, bytes, in, out, (int)n, 0, 64, st));
    void* tmp = scratch(st, 5, bytes);
    if (!tmp) return QPB_ERR_NOMEM;
    QPB_CUDA(cub::DeviceRadixSort::SortKeys(tmp, bytes, in, out, (int)n, 0, 64, st));
    return check_cuda(cudaGetLastError(), "cub::DeviceRadixSort::SortKeys");  // library kernels: not counted as ours
}
