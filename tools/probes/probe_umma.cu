// Probe (diagnostics): how fast does one SM execute tcgen05.mma (kind::f16, bf16 operands, M = 128, K = 16, cta_group::1) as a function of
//   N (64 / 128 / 192 / 256), where A comes from (shared memory, 128B-swizzled K-major / tensor memory), and whether consecutive MMAs
//   accumulate into the same TMEM columns or rotate over 2 / 4 accumulators?
// One CTA per SM (grid = number of SMs given on the command line, default 1), one elected lane issues `count` MMAs back to back and
// waits for the commit; reported: clocks per MMA from the difference of a long and a short run (fixed latencies cancel), and the floor
// N / 2 clocks.  Operand contents are irrelevant (zeros).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I gstreamer_vit_tracker_b200/csrc -I include tools/probes/probe_umma.cu -o tools/probes/probe_umma
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "tc_common.cuh"

using namespace vt::tc;

struct Cfg {
    int N, a_tmem, n_acc, pair;  // pair: alternate (N, N / 2) as the bf16x3 K-step does
};

__global__ void __launch_bounds__(128, 1) probe(const Cfg c, const int count, long long* out) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    for (int i = threadIdx.x; i < (96 * 1024) / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
    if (threadIdx.x == 0) mbar_init(&bar, 1), fence_barrier_init();
    if (threadIdx.x < 32) tmem_alloc(&tmem_base_s, 512), tmem_relinquish();
    fence_proxy_async_smem();
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem = tmem_base_s;
    if (threadIdx.x < 32) {
        // A: 4 k-blocks of [128 rows][128 B] (16 KB each) at 0; B: 4 k-blocks of [256 rows][128 B] (32 KB each)... only 2 fit: B walks 2
        const uint32_t a_lo = umma_desc_lo(smem_u32(smem)), b_lo = umma_desc_lo(smem_u32(smem + 32 * 1024));
        const uint32_t idesc = umma_idesc_bf16(128, c.N), idesc_h = umma_idesc_bf16(128, c.N / 2);
        // accumulators at columns 0, N, 2N, ... (n_acc * N <= 384); TMEM A operand at column 448
        long long t0 = 0, t1 = 0;
        if (elect_one_sync()) {
            t0 = clock64();
            for (int i = 0; i < count; ++i) {
                const int k = i & 3, kb = (i >> 2) & 1;
                const uint32_t acc = tmem + (uint32_t)((i % c.n_acc) * c.N);
                const uint64_t dA = umma_desc_from_lo(a_lo + kb * (16384 >> 4) + 2 * k), dB = umma_desc_from_lo(b_lo + kb * (32768 >> 4) + 2 * k);
                const bool half = c.pair && (i & 1);
                if (c.a_tmem) umma_bf16_ta(acc, tmem + 448 + 8 * k, dB, half ? idesc_h : idesc, 1);
                else umma_bf16(acc, dA, dB, half ? idesc_h : idesc, 1);
            }
            umma_commit(&bar);
        }
        __syncwarp();
        mbar_wait(&bar, 0);
        t1 = clock64();
        if (elect_one_sync() && blockIdx.x == 0) out[0] = t1 - t0;
        __syncwarp();
    }
    tcgen05_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc(tmem, 512);
}

int main(int argc, char** argv) {
    const int grid = argc > 1 ? atoi(argv[1]) : 1;
    long long* d_out;
    cudaMalloc(&d_out, 8);
    const size_t smem = 97 * 1024 + 1024;
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const Cfg cfgs[] = {{64, 0, 1, 0},  {128, 0, 1, 0}, {192, 0, 1, 0}, {256, 0, 1, 0}, {64, 0, 2, 0},  {64, 0, 4, 0}, {128, 0, 2, 0}, {128, 0, 1, 1},
                        {64, 1, 1, 0},  {128, 1, 1, 0}, {192, 1, 1, 0}, {256, 1, 1, 0}, {64, 1, 2, 0},  {64, 1, 4, 0}, {128, 1, 1, 1}};
    printf("grid %d CTA(s)\n%5s %6s %5s %5s | %10s %10s | %8s\n", grid, "N", "A", "accs", "pair", "clk/MMA", "floor", "ratio");
    for (const Cfg& c : cfgs) {
        long long t[2] = {0, 0};
        const int counts[2] = {64, 576};
        for (int r = 0; r < 2; ++r) {
            for (int rep = 0; rep < 3; ++rep) {  // last repetition counts (warm)
                probe<<<grid, 128, smem>>>(c, counts[r], d_out);
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) {
                    printf("CUDA error: %s\n", cudaGetErrorString(e));
                    return 1;
                }
                cudaMemcpy(&t[r], d_out, 8, cudaMemcpyDeviceToHost);
            }
        }
        const double per = (double)(t[1] - t[0]) / (counts[1] - counts[0]);
        const double floor_clk = c.pair ? 0.5 * (c.N / 2.0 + c.N / 4.0) : c.N / 2.0;
        printf("%5d %6s %5d %5d | %10.1f %10.1f | %8.2f\n", c.N, c.a_tmem ? "tmem" : "smem", c.n_acc, c.pair, per, floor_clk, per / floor_clk);
    }
    return 0;
}
