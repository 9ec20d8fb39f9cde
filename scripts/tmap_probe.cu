// tmap_probe.cu -- does the driver accept the 5-D tensor map of the UMMA row operand (non-monotonic strides)?
// nvcc -arch=sm_100a scripts/tmap_probe.cu -o scripts/tmap_probe -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
int main()
{
    cudaFree(0);
    const int nch = 211; const size_t rb = (size_t)nch * 128; const int rows = 4096;
    void* d = nullptr; cudaMalloc(&d, rb * rows);
    CUtensorMap m;
    // dims innermost first: 16-byte unit (4 x u32), u8 (8 rows), part (2), octet (rows/8), quad (4 * nch)
    cuuint64_t dims[5] = {4, 8, 2, (cuuint64_t)rows / 8, (cuuint64_t)4 * nch};
    cuuint64_t strides[4] = {rb, 16, 8 * rb, 32};
    cuuint32_t box[5] = {4, 8, 2, 4, 4};
    cuuint32_t es[5] = {1, 1, 1, 1, 1};
    CUresult r = cuTensorMapEncodeTiled(&m, CU_TENSOR_MAP_DATA_TYPE_UINT32, 5, d, dims, strides, box, es,
                                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                        CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    const char* s = nullptr; cuGetErrorString(r, &s);
    printf("5D unsorted strides: %d %s\n", (int)r, s ? s : "");
    cuuint32_t box2[5] = {4, 8, 2, 2, 4};
    r = cuTensorMapEncodeTiled(&m, CU_TENSOR_MAP_DATA_TYPE_UINT32, 5, d, dims, strides, box2, es,
                               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                               CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    cuGetErrorString(r, &s);
    printf("5D half box: %d %s\n", (int)r, s ? s : "");
    return 0;
}
