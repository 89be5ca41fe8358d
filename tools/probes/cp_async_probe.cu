// Probe: does cp.async with the `ignore-src` predicate write zeros to shared memory (as the PTX ISA says) without touching the
// source?  Prints "IGNORE_SRC_ZEROS 1" if so.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -o cp_async_probe cp_async_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
__global__ void probe(const float* g, float* out) {
    __shared__ __align__(16) float s[32 * 4];
    for (int q = 0; q < 4; ++q) s[threadIdx.x * 4 + q] = 7.0f;
    __syncthreads();
    const unsigned sa = (unsigned)__cvta_generic_to_shared(s + threadIdx.x * 4);
    const unsigned valid = threadIdx.x & 1;      // odd lanes copy, even lanes ignore the source
    // even lanes get a wild pointer: it must not be dereferenced
    const float* src = valid ? g + threadIdx.x * 4 : reinterpret_cast<const float*>(0x10);
    asm volatile("{ .reg .pred p; setp.eq.u32 p, %2, 0; cp.async.cg.shared.global [%0], [%1], 16, p; }" ::"r"(sa), "l"(src), "r"(valid) : "memory");
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    for (int q = 0; q < 4; ++q) out[threadIdx.x * 4 + q] = s[threadIdx.x * 4 + q];
}
int main() {
    float *g, *o, h[128], r[128];
    for (int i = 0; i < 128; ++i) h[i] = 100.0f + i;
    cudaMalloc(&g, sizeof h); cudaMalloc(&o, sizeof h);
    cudaMemcpy(g, h, sizeof h, cudaMemcpyHostToDevice);
    probe<<<1, 32>>>(g, o);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("IGNORE_SRC_ZEROS 0 (%s)\n", cudaGetErrorString(e)); return 0; }
    cudaMemcpy(r, o, sizeof r, cudaMemcpyDeviceToHost);
    int ok = 1;
    for (int t = 0; t < 32; ++t)
        for (int q = 0; q < 4; ++q) {
            const float want = (t & 1) ? h[t * 4 + q] : 0.0f;
            if (r[t * 4 + q] != want) ok = 0;
        }
    printf("IGNORE_SRC_ZEROS %d (lane 0: %g, lane 1: %g)\n", ok, r[0], r[4]);
    return 0;
}
