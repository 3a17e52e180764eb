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

// Batch assembly from a device-resident uint8 dataset: out[r, :] = convert(images[idx[r], :]).  One pass does what
// DataLoader(shuffle=True) + collate + datasets.py:28 do on the host (index_select of the sampler's permutation, then the
// byte -> fp32 conversion): idx[r] reads (row_bytes, 16-byte aligned) -> 4 x row_bytes written, coalesced on both sides.
__global__ void __launch_bounds__(256) pixels_gather_kernel(const uint8_t* __restrict__ images, const int64_t* __restrict__ idx,
                                                            float* __restrict__ out, int64_t rows, int64_t row_bytes) {
    __shared__ float lut[256];
    lut[threadIdx.x] = px(threadIdx.x);
    __syncthreads();
    const int64_t nv = row_bytes >> 4;                     // 16-byte units per image
    const int64_t total = rows * nv;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / nv, u = i - r * nv;
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(images + idx[r] * row_bytes) + u);
        const unsigned w[4] = {v.x, v.y, v.z, v.w};
        float4* o = reinterpret_cast<float4*>(out + r * row_bytes) + u * 4;
#pragma unroll
        for (int j = 0; j < 4; ++j)
            o[j] = make_float4(lut[w[j] & 255u], lut[(w[j] >> 8) & 255u], lut[(w[j] >> 16) & 255u], lut[w[j] >> 24]);
    }
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

extern "C" int cdg_pixels_gather_to_float(const uint8_t* images, int64_t n_images, int64_t row_bytes, const int64_t* idx,
                                          int64_t rows, float* out, void* stream) {
    CDG_REQUIRE(rows >= 0 && n_images >= 0 && row_bytes > 0, "pixels_gather_to_float: bad extents");
    if (rows == 0) return CDG_OK;
    CDG_REQUIRE(images && idx && out, "pixels_gather_to_float: null pointer");
    CDG_REQUIRE(row_bytes % 16 == 0 && ((reinterpret_cast<uintptr_t>(images) | reinterpret_cast<uintptr_t>(out)) & 15) == 0,
                "pixels_gather_to_float: images must be whole 16-byte units (row_bytes %% 16 == 0, 16-byte aligned buffers)");
    const int64_t total = rows * (row_bytes >> 4);
    const unsigned blocks = (unsigned)imax64(1, imin64((total + 255) / 256, (int64_t)kNumSMs * 8));
    pixels_gather_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(images, idx, out, rows, row_bytes);
    CDG_CHECK_LAUNCH();
    return CDG_OK;
}
