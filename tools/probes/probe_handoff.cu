// Probe (diagnostics): what does the hand-over between two dependent kernels cost on this GPU?
//   (a) programmatic dependent launch: consumer CTAs wait in griddepcontrol.wait (returns when the producer grid has completed and flushed)
//   (b) flag hand-over: every producer CTA stores its 32 KB tile, __threadfence(), atomicAdd on a counter; consumer CTAs (launched with PDL
//       so that they are resident early, but never calling griddepcontrol.wait) poll the counter with ld.acquire and then read the tile
// Reported: consumer "data ready" time minus the producer's last "stores issued" time (%globaltimer, max / min over CTAs), median of
// the iterations.  The consumer also checks the data it reads.
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <vector>
#include <cuda_runtime.h>
__device__ __forceinline__ unsigned long long gtime() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)::"memory");
    return t;
}
constexpr int kTileF4 = 32 * 1024 / 16;
__global__ void __launch_bounds__(512) producer(float4* data, unsigned* counter, unsigned long long* t_prod, int epoch, int spin) {
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    float x = (float)epoch;
    for (int i = 0; i < spin; ++i) x = x * 1.0000001f + 1e-9f;  // some work so that the consumer is resident before we finish
    float4* mine = data + (size_t)blockIdx.x * kTileF4;
    for (int i = threadIdx.x; i < kTileF4; i += 512) mine[i] = make_float4((float)epoch, (float)i, x, 0.f);
    __syncthreads();
    if (threadIdx.x == 0) {
        t_prod[blockIdx.x] = gtime();
        if (counter) {
            __threadfence();
            atomicAdd(counter, 1u);
        }
    }
}
__global__ void __launch_bounds__(512) consumer(const float4* data, unsigned* counter, unsigned target, unsigned long long* t_cons, int* bad,
                                                int epoch, int n_prod) {
    if (counter) {
        if (threadIdx.x == 0) {
            unsigned v;
            do {
                asm volatile("ld.acquire.gpu.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
            } while (v < target);
        }
        __syncthreads();
    } else {
        asm volatile("griddepcontrol.wait;" ::: "memory");
    }
    if (threadIdx.x == 0) t_cons[blockIdx.x] = gtime();
    // read a tile of another producer CTA and check it
    const float4* src = data + (size_t)((blockIdx.x * 7 + 3) % n_prod) * kTileF4;
    int wrong = 0;
    for (int i = threadIdx.x; i < kTileF4; i += 512) {
        const float4 v = counter ? __ldcg(src + i) : src[i];
        wrong |= v.x != (float)epoch || v.y != (float)i;
    }
    if (wrong) atomicAdd(bad, 1);
}
template <typename... KArgs, typename... Args>
static void launch_pdl(void (*k)(KArgs...), int grid, cudaStream_t s, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid), cfg.blockDim = dim3(512), cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization, at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at, cfg.numAttrs = 1;
    cudaLaunchKernelEx(&cfg, k, KArgs(args)...);
}
int main() {
    const int n_prod = 36, n_cons = 27, iters = 200;
    float4* data;
    unsigned* counter;
    unsigned long long *tp, *tc;
    int* bad;
    cudaMalloc(&data, (size_t)n_prod * kTileF4 * 16), cudaMalloc(&counter, 4), cudaMalloc(&tp, 8 * n_prod), cudaMalloc(&tc, 8 * n_cons), cudaMalloc(&bad, 4);
    cudaMemset(counter, 0, 4), cudaMemset(bad, 0, 4);
    cudaStream_t s;
    cudaStreamCreate(&s);
    std::vector<unsigned long long> hp(n_prod), hc(n_cons);
    for (int mode = 0; mode < 2; ++mode) {
        std::vector<double> first, last;
        for (int it = 0; it < iters; ++it) {
            const int epoch = mode * iters + it + 1;
            unsigned* c = mode ? counter : nullptr;
            launch_pdl(producer, n_prod, s, data, c, tp, epoch, 2000);
            launch_pdl(consumer, n_cons, s, (const float4*)data, c, (unsigned)((it + 1) * n_prod), tc, bad, epoch, n_prod);
            cudaStreamSynchronize(s);
            cudaMemcpy(hp.data(), tp, 8 * n_prod, cudaMemcpyDeviceToHost), cudaMemcpy(hc.data(), tc, 8 * n_cons, cudaMemcpyDeviceToHost);
            const unsigned long long p = *std::max_element(hp.begin(), hp.end());
            first.push_back((double)*std::min_element(hc.begin(), hc.end()) - (double)p), last.push_back((double)*std::max_element(hc.begin(), hc.end()) - (double)p);
        }
        std::sort(first.begin(), first.end()), std::sort(last.begin(), last.end());
        int hb = 0;
        cudaMemcpy(&hb, bad, 4, cudaMemcpyDeviceToHost);
        printf("%s: consumer ready after the last producer CTA issued its stores: first CTA %.0f ns, last CTA %.0f ns (median of %d); bad reads %d; %s\n",
               mode ? "flag hand-over (fence + atomic, ld.acquire poll)" : "griddepcontrol.wait (PDL)", first[iters / 2], last[iters / 2], iters, hb,
               cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
