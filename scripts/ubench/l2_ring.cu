// Does an intermediate of Y_MIB MiB written by one kernel stay in the 126 MB L2 for the next kernel while that
// kernel streams 16x as many bytes out with st.global.cs (the map)?  Run under
//   ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum ./l2_ring [Y_MIB]
// and compare kernel k_consume's DRAM reads with Y_MIB.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>

__global__ void k_produce(float4 *y, long long n4)
{
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x)
        y[i] = make_float4((float)i, 1.f, 2.f, 3.f);
}
__global__ void k_consume(const float4 *y, long long n4, float4 *map)
{
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        const float4 v = y[i];
#pragma unroll
        for (int k = 0; k < 16; k++) __stcs(map + k * n4 + i, make_float4(v.x + k, v.y, v.z, v.w));      // whole lines per warp store
    }
}
int main(int argc, char **argv)
{
    const long long y_mib = argc > 1 ? atoll(argv[1]) : 32;
    const long long n4 = y_mib * (1 << 20) / 16;
    float4 *y, *map;
    cudaMalloc(&y, n4 * 16);
    cudaMalloc(&map, n4 * 16 * 16);
    for (int rep = 0; rep < 3; rep++) {
        k_produce<<<148 * 4, 256>>>(y, n4);
        k_consume<<<148 * 4, 256>>>(y, n4, map);
    }
    cudaError_t e = cudaDeviceSynchronize();
    printf("Y = %lld MiB, map = %lld MiB per round: %s\n", y_mib, y_mib * 16, cudaGetErrorString(e));
    return 0;
}
