// uint8 pixels -> the fp32 images the training step reads: modules/datasets.py:28 / :57
//   x_data = (np.array(train_x).astype(float) - 127.5) / 127.5      (float64)
// followed by torch.FloatTensor(x_data[idx]) (:42, :64: round to fp32).  The reference does this once on the host and
// then moves 4 bytes per pixel to the GPU every step; here the dataset's native bytes cross PCIe (1 byte per pixel) and
// the identical fp64 -> fp32 arithmetic runs on the device.  HBM-bound: 1 B read + 4 B written per pixel, 16 pixels per
// thread per iteration (one 16-byte load, four 16-byte stores).
#include "common.cuh"

namespace cdg {
namespace {

__device__ __forceinline__ float px(unsigned v) {
    return (float)__ddiv_rn(__dsub_rn((double)v, 127.5), 127.5);
}

__global__ void __launch_bounds__(256) pixels_to_float_kernel(const uint8_t* __restrict__ in, float* __restrict__ out, int64_t n) {
    __shared__ float lut[256];
    lut[threadIdx.x] = px(threadIdx.x);
    __syncthreads();
    const int64_t nv = n >> 4;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nv; i += stride) {
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(in) + i);
        const unsigned w[4] = {v.x, v.y, v.z, v.w};
        float4* o = reinterpret_cast<float4*>(out) + i * 4;
#pragma unroll
        for (int j = 0; j < 4; ++j)
            o[j] = make_float4(lut[w[j] & 255u], lut[(w[j] >> 8) & 255u], lut[(w[j] >> 16) & 255u], lut[w[j] >> 24]);
    }
    for (int64_t i = (nv << 4) + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += stride) out[i] = lut[in[i]];
}

__global__ void pixels_to_float_unaligned_kernel(const uint8_t* __restrict__ in, float* __restrict__ out, int64_t n) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) out[i] = px(in[i]);
}

}  // namespace
}  // namespace cdg

using namespace cdg;

extern "C" int cdg_pixels_to_float(const uint8_t* pixels, int64_t n, float* out, void* stream) {
    CDG_REQUIRE(n >= 0, "pixels_to_float: negative count");
    if (n == 0) return CDG_OK;
    CDG_REQUIRE(pixels && out, "pixels_to_float: null pointer");
    const bool aligned = ((reinterpret_cast<uintptr_t>(pixels) | reinterpret_cast<uintptr_t>(out)) & 15) == 0;
    const unsigned blocks = (unsigned)imax64(1, imin64(((n >> 4) + 255) / 256 + 1, (int64_t)kNumSMs * 8));
    if (aligned) pixels_to_float_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(pixels, out, n);
    else pixels_to_float_unaligned_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(pixels, out, n);
    CDG_CHECK_LAUNCH();
    return CDG_OK;
}
