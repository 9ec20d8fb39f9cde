// Micro-benchmarks that size the tensor-core CCF contraction:
//  (1) mma.sync.m16n8k16 bf16 rate versus resident warps per SM (register-resident operands);
//  (2) L2 -> SM bandwidth of coalesced 128-bit loads over an L2-resident buffer, versus resident
//      warps and loads in flight per thread.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/tc_probe scripts/tc_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ void mma_bf16(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2])
{
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

template <int NACC>
__global__ void mma_kernel(float* out, int iters, uint32_t seed)
{
    float c[NACC][4];
#pragma unroll
    for (int i = 0; i < NACC; ++i) { c[i][0] = c[i][1] = c[i][2] = c[i][3] = 0.f; }
    uint32_t a[4] = {seed + threadIdx.x, seed * 3, seed ^ 0x3f803f80u, 0x3f803f80u};
    uint32_t b[2] = {0x3f803f80u, seed};
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) mma_bf16(c[i], a, b);
        a[0] += 1; b[1] += 1;
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    if (s == 123.456f) out[0] = s;
}

// each warp streams 512-byte lines: lane reads 16 B; U independent loads in flight per thread
template <int U>
__global__ void l2_kernel(const uint4* __restrict__ buf, size_t nvec, int iters, uint32_t* out)
{
    const size_t warp = (size_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const size_t nwarps = (size_t)gridDim.x * (blockDim.x >> 5);
    const int lane = threadIdx.x & 31;
    uint32_t acc = 0;
    size_t pos = warp * 32 * U;
    for (int it = 0; it < iters; ++it) {
        uint4 v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            size_t i = pos + (size_t)u * 32 + lane;
            if (i >= nvec) i -= nvec * (i / nvec);
            v[u] = __ldg(buf + i);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) acc += v[u].x ^ v[u].y ^ v[u].z ^ v[u].w;
        pos += nwarps * 32 * U;
        if (pos >= nvec) pos -= nvec * (pos / nvec);
    }
    if (acc == 0x12345678u) out[0] = acc;
}

int main()
{
    float* d; cudaMalloc(&d, 64);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
    const int nsm = prop.multiProcessorCount;
    printf("device %s, %d SMs, L2 %d MB\n", prop.name, nsm, prop.l2CacheSize >> 20);
    {
        const int iters = 4000;
        for (int wps = 4; wps <= 32; wps *= 2) {
            float best = 1e30f;
            for (int rep = 0; rep < 3; ++rep) {
                cudaEventRecord(e0);
                mma_kernel<12><<<nsm, wps * 32>>>(d, iters, 17u + rep);
                cudaEventRecord(e1); cudaEventSynchronize(e1);
                float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
            }
            const double macs = 2048.0 * 12 * iters * wps * nsm;
            printf("mma.sync m16n8k16 bf16, %2d warps/SM: %.1f dense TFLOP/s, %.0f MAC/clk/SM @1.965GHz, %.2f clk per MMA per SM\n",
                   wps, 2 * macs / best / 1e9, macs / (best * 1e-3) / nsm / 1.965e9,
                   (best * 1e-3) * 1.965e9 / (12.0 * iters * wps));
        }
    }
    {
        const size_t bytes = (size_t)48 << 20;     // L2-resident
        const size_t nvec = bytes / 16;
        uint4* buf; cudaMalloc(&buf, bytes); cudaMemset(buf, 1, bytes);
        uint32_t* o; cudaMalloc(&o, 64);
        for (int wps = 8; wps <= 64; wps *= 2)
            for (int u = 2; u <= 8; u *= 2) {
                const int thr = wps >= 32 ? 1024 : wps * 32, cps = wps >= 32 ? wps / 32 : 1;
                const int iters = 2048 / u;
                float best = 1e30f;
                for (int rep = 0; rep < 4; ++rep) {
                    cudaEventRecord(e0);
                    if (u == 2) l2_kernel<2><<<nsm * cps, thr>>>(buf, nvec, iters, o);
                    else if (u == 4) l2_kernel<4><<<nsm * cps, thr>>>(buf, nvec, iters, o);
                    else l2_kernel<8><<<nsm * cps, thr>>>(buf, nvec, iters, o);
                    cudaEventRecord(e1); cudaEventSynchronize(e1);
                    float ms; cudaEventElapsedTime(&ms, e0, e1); if (rep > 0 && ms < best) best = ms;
                }
                const double b = (double)nsm * cps * thr * 16.0 * u * iters;
                printf("L2 read 48MB, %2d warps/SM, %d x LDG.128 in flight: %.2f TB/s (%.1f B/clk/SM @1.965GHz)\n",
                       wps, u, b / best / 1e9, b / (best * 1e-3) / nsm / 1.965e9);
            }
    }
    cudaError_t e = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(e));
    return e != cudaSuccess;
}
