// Feasibility probe (diagnostics): can a 512-thread kernel with ~200 KB of dynamic shared memory be launched as clusters of 12 / 16 CTAs
// (non-portable size) on this GPU, how many such clusters are co-resident, and what does a DSMEM gather of 88 KB per CTA cost?
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(512, 1) k(float* out, int iters) {
    extern __shared__ uint8_t smem[];
    unsigned rank, n;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(n));
    float4* mine = reinterpret_cast<float4*>(smem);
    for (int i = threadIdx.x; i < 96 * 1024 / 16; i += 512) mine[i] = make_float4(rank, i, 1.f, 2.f);
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
    unsigned long long t0, t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0)::"memory");
    float4 acc = make_float4(0, 0, 0, 0);
    for (int it = 0; it < iters; ++it) {
        // each thread reads one 16-byte chunk (row = tid/4, chunk = tid%4 of slab `rank`) from every CTA of the cluster
        // all remote loads in flight before the first add (non-volatile asm: the compiler may schedule them back to back)
        float4 v[16];
#pragma unroll
        for (unsigned r = 0; r < 16; ++r) {
            if (r < n) {
                const uint32_t local = (uint32_t)__cvta_generic_to_shared(smem) + ((threadIdx.x >> 2) * 128 + ((rank & 1) * 4 + (threadIdx.x & 3)) * 16) +
                                       (rank >> 1) * 16384;
                uint32_t remote;
                asm("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local), "r"(r));
                asm("ld.shared::cluster.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v[r].x), "=f"(v[r].y), "=f"(v[r].z), "=f"(v[r].w) : "r"(remote + it * 0));
            }
        }
#pragma unroll
        for (unsigned r = 0; r < 16; ++r)
            if (r < n) acc.x += v[r].x, acc.y += v[r].y, acc.z += v[r].z, acc.w += v[r].w;
    }
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1)::"memory");
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
    if (threadIdx.x == 0) out[blockIdx.x * 2] = acc.x + acc.y, out[blockIdx.x * 2 + 1] = (float)(t1 - t0);
}
int main() {
    float* d;
    cudaMalloc(&d, 4096);
    for (int cl : {3, 8, 12, 16}) {
        const int smem = 200 * 1024;
        cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(cl * 3), cfg.blockDim = dim3(512), cfg.dynamicSmemBytes = smem;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = cl, at[0].val.clusterDim.y = 1, at[0].val.clusterDim.z = 1;
        cfg.attrs = at, cfg.numAttrs = 1;
        int nclusters = -1;
        cudaError_t eo = cudaOccupancyMaxActiveClusters(&nclusters, k, &cfg);
        cudaError_t el = cudaLaunchKernelEx(&cfg, k, d, 1);
        cudaError_t es = cudaDeviceSynchronize();
        float h[96];
        cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
        printf("cluster %2d: attr %s, max active clusters %d (%s), launch %s, sync %s, gather of %d x 8 KB per CTA: %.2f us\n", cl, cudaGetErrorName(e),
               nclusters, cudaGetErrorName(eo), cudaGetErrorName(el), cudaGetErrorName(es), cl, h[1] * 1e-3);
        cudaGetLastError();
    }
    return 0;
}
