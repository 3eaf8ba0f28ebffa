// volume.cu — slices of a volume resident in HBM, as the stack / orthoplane loops of scripts/pdl_inference3d.py:110-176
// take them (data/volume_dataset.py:37-53 -> array_utils.take :6-23): n consecutive slices along one axis, gathered into
// a contiguous (n, A, B) batch.  Along axis 0 a slice is contiguous (one copy); along axis 1 its rows are (row copies);
// along axis 2 every element of it lies a whole row apart — a naive gather reads one byte per 32-byte sector — so n
// neighbouring slices are taken together: a 32 x 32 tile of (row, slice) pairs is read along the slices (contiguous in
// memory) and written along the rows of each slice, through shared memory.
#include <algorithm>
#include "common.cuh"

namespace emp {

template <typename T>
__global__ void __launch_bounds__(256)
take_axis1_kernel(const T* __restrict__ vol, int D, int H, int W, int i0, int n, T* __restrict__ out)
{
    // out[s][z][x] = vol[z][i0 + s][x]
    const size_t total = (size_t)n * D * W;
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
        const int x = (int)(e % W);
        const size_t r = e / W;
        const int z = (int)(r % D), s = (int)(r / D);
        out[e] = vol[((size_t)z * H + (i0 + s)) * W + x];
    }
}

template <typename T>
__global__ void __launch_bounds__(256)
take_axis2_kernel(const T* __restrict__ vol, int D, int H, int W, int i0, int n, T* __restrict__ out)
{
    // out[s][z][y] = vol[z][y][i0 + s]; block = one z, 32 rows y, 32 slices s
    __shared__ T tile[32][33];
    const int z = blockIdx.z;
    const int y0 = blockIdx.y * 32, s0 = blockIdx.x * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;                // 32 x 8
    for (int r = ty; r < 32; r += 8) {
        const int y = y0 + r, s = s0 + tx;
        if (y < H && s < n) tile[r][tx] = vol[((size_t)z * H + y) * W + i0 + s];
    }
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {
        const int s = s0 + r, y = y0 + tx;
        if (s < n && y < H) out[((size_t)s * D + z) * H + y] = tile[tx][r];
    }
}

}  // namespace emp

using namespace emp;

EMP_API int emp_take_slices(const void* vol, int elem_bytes, int D, int H, int W, int axis, int i0, int n, void* out, void* stream)
{
    EMP_REQUIRE(vol && out, EMP_ERR_INVALID, "null pointer");
    EMP_REQUIRE(elem_bytes == 1 || elem_bytes == 4, EMP_ERR_INVALID, "elements must be 1 or 4 bytes");
    EMP_REQUIRE(D > 0 && H > 0 && W > 0 && axis >= 0 && axis <= 2 && n >= 1 && i0 >= 0, EMP_ERR_INVALID, "bad shape / axis");
    const int len = axis == 0 ? D : axis == 1 ? H : W;
    EMP_REQUIRE(i0 + n <= len, EMP_ERR_INVALID, "slices %d .. %d exceed the axis (%d)", i0, i0 + n, len);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (axis == 0) {
        EMP_CUDA_CHECK(cudaMemcpyAsync(out, static_cast<const char*>(vol) + (size_t)i0 * H * W * elem_bytes,
                                       (size_t)n * H * W * elem_bytes, cudaMemcpyDeviceToDevice, st));
        return EMP_OK;
    }
    if (axis == 1) {
        const size_t total = (size_t)n * D * W;
        const unsigned grid = (unsigned)std::min<size_t>((total + 255) / 256, (size_t)device_sm_count() * 32);
        if (elem_bytes == 1) take_axis1_kernel<unsigned char><<<grid, 256, 0, st>>>(static_cast<const unsigned char*>(vol), D, H, W, i0, n, static_cast<unsigned char*>(out));
        else take_axis1_kernel<unsigned><<<grid, 256, 0, st>>>(static_cast<const unsigned*>(vol), D, H, W, i0, n, static_cast<unsigned*>(out));
    } else {
        EMP_REQUIRE(D <= 65535 && (H + 31) / 32 <= 65535, EMP_ERR_INVALID, "volume too large for one launch");
        const dim3 grid((n + 31) / 32, (H + 31) / 32, D);
        if (elem_bytes == 1) take_axis2_kernel<unsigned char><<<grid, 256, 0, st>>>(static_cast<const unsigned char*>(vol), D, H, W, i0, n, static_cast<unsigned char*>(out));
        else take_axis2_kernel<unsigned><<<grid, 256, 0, st>>>(static_cast<const unsigned*>(vol), D, H, W, i0, n, static_cast<unsigned*>(out));
    }
    EMP_CUDA_CHECK(cudaGetLastError());
    return EMP_OK;
}
