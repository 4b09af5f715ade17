// gemm_as.cu — the A-stationary throughput form of the K <= 192 contractions (QKV, FC1) for many rows (cfg4: 16 targets = 5120 rows;
// cfg5: stream groups), where the one-tile-per-CTA kernel of gemm_tc.cu spends most of a CTA's life in prologue, first-TMA latency and
// epilogue with the tensor pipe idle, and re-reads the activation tile once per column tile.
//
//   C[M, N] = epilogue( A[M, K] * W[N, K]^T + bias ),   K = 64 * num_kb <= 192
//
// One CTA (18 warps) owns one 128-row tile of A and a contiguous range of 64-column chunks of N:
//   * the A tile (all of K, bf16 hi + lo: <= 96 KB) is loaded ONCE and stays in shared memory;
//   * warp 0 streams the weight chunks through a ring of k-block stages ([W_hi; W_lo] of one 64 x 64 block = 16 KB; the first
//     stages are in flight before griddepcontrol.wait — weights never depend on the preceding kernel);
//   * warp 1 issues the UMMAs of chunk i into TMEM accumulator i % 2 (bf16x3: A_hi x [W_hi; W_lo] as one N = 128 UMMA + A_lo x W_hi,
//     as in gemm_tc.cu), so the main loop of chunk i + 1 runs under the epilogue of chunk i;
//   * warps 2..17 are the epilogue: tcgen05.ld -> release the accumulator -> bias / GELU -> bf16 (hi, lo) split -> staging tile in
//     shared memory (double buffered) -> fully coalesced 16-byte stores (same staging and copy-out as gemm_tc.cu; Q / K / V^T scatter).
// L2 -> SM traffic per row tile: the A tile once per CTA + every weight chunk once, instead of (A tile + chunk) per 128 x 64 tile.
// The sums are the same as in gemm_tc.cu, in the same order (k ascending; hi*hi + lo*hi in one accumulator half, hi*lo in the
// other, added in the epilogue): results are bit-identical to the one-tile form.
#include <stdlib.h>
#include <string.h>

#include "gemm_epi.cuh"

namespace vt {

constexpr int kAsStages = 4;           // weight k-block stages in flight
constexpr int kAsMaxKb = 3;            // K <= 192
constexpr int kAsThreads = 64 + kTcThreads;  // TMA warp, MMA warp, 16 epilogue warps
constexpr int kAsChunk = 64;           // columns per accumulator / epilogue step

template <int NSPLIT>
struct AsSmem {
    static constexpr int kParts = NSPLIT == 3 ? 2 : 1;
    static constexpr int kABytes = kAsMaxKb * kParts * kTileABytes;      // 96 / 48 KB
    static constexpr int kStageBytes = kParts * kAsChunk * kTcBK * 2;    // 16 / 8 KB
    static constexpr int kTileOBytes = kTcBM * kAsChunk * 2;             // 16 KB: one bf16 output tile
    static constexpr int kOffB = kABytes;
    static constexpr int kOffO = kOffB + kAsStages * kStageBytes;        // 2 buffers x (hi, lo)
    static constexpr int kTotal = kOffO + 2 * kParts * kTileOBytes + 1024;
};

template <int NSPLIT>
__global__ void __launch_bounds__(kAsThreads, 1) gemm_as_kernel(const __grid_constant__ TcMaps mp, const TcGemmArgs a, const int chunks_per_cta) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t a_bar[kAsMaxKb], full_bar[kAsStages], empty_bar[kAsStages], acc_full[2], acc_empty[2];
    __shared__ uint32_t tmem_base_s;
    __shared__ unsigned long long* trace_slot;
    using SM = AsSmem<NSPLIT>;
    constexpr bool kLo = NSPLIT == 3, kF16 = NSPLIT == 2;
    constexpr int kParts = SM::kParts, CPT = kAsChunk / kTcColGroups;  // 16 accumulator columns per epilogue thread
    constexpr uint32_t kAccCols = kLo ? 2 * kAsChunk : kAsChunk;      // bf16x3 keeps hi*lo in a second column half
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int m0 = blockIdx.y * kTcBM;
    const int num_kb = a.K / kTcBK, n_chunks = a.N / kAsChunk;
    const int c_begin = blockIdx.x * chunks_per_cta;
    const int c_end = c_begin + chunks_per_cta < n_chunks ? c_begin + chunks_per_cta : n_chunks;
    const int my_chunks = c_end - c_begin, n_items = my_chunks * num_kb;  // item = (chunk, k-block) in issue order
    bool ok = true;
    TraceRec tr;
    tr.begin(&trace_slot, a.trace, a.trace_id);

    // ---- prologue: independent of the preceding kernel
    if (tid == 0) {
        tma_prefetch_desc(&mp.Ahi), tma_prefetch_desc(&mp.Bhi);
        if (kLo) tma_prefetch_desc(&mp.Alo), tma_prefetch_desc(&mp.Blo);
        for (int i = 0; i < kAsMaxKb; ++i) mbar_init(&a_bar[i], 1);
        for (int s = 0; s < kAsStages; ++s) mbar_init(&full_bar[s], 1), mbar_init(&empty_bar[s], 1);
        for (int b = 0; b < 2; ++b) mbar_init(&acc_full[b], 1), mbar_init(&acc_empty[b], kTcThreads / 32);
        fence_barrier_init();
        const int npre = n_items < kAsStages ? n_items : kAsStages;
        for (int it = 0; it < npre; ++it) {
            const int c = c_begin + it / num_kb, kb = it % num_kb;
            uint8_t* sb = smem + SM::kOffB + it * SM::kStageBytes;
            mbar_arrive_expect_tx(&full_bar[it], SM::kStageBytes);
            tma_load_2d(sb, &mp.Bhi, &full_bar[it], kb * kTcBK, c * kAsChunk);
            if (kLo) tma_load_2d(sb + kAsChunk * 128, &mp.Blo, &full_bar[it], kb * kTcBK, c * kAsChunk);
        }
    }
    if (warp == 1) {
        tmem_alloc(&tmem_base_s, 2 * kAccCols);
        tmem_relinquish();
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem = tmem_base_s;

    pdl_wait();
    if (tid == 0) tr.mark(2);
    pdl_launch_dependents();

    if (warp == 0) {
        if (lane == 0) {  // ---- TMA producer: the activation tile once, then the rest of the weight stream
            for (int kb = 0; kb < num_kb; ++kb) {
                uint8_t* sa = smem + kb * kParts * kTileABytes;
                mbar_arrive_expect_tx(&a_bar[kb], kParts * kTileABytes);
                tma_load_2d(sa, &mp.Ahi, &a_bar[kb], kb * kTcBK, m0);
                if (kLo) tma_load_2d(sa + kTileABytes, &mp.Alo, &a_bar[kb], kb * kTcBK, m0);
            }
            for (int it = kAsStages; it < n_items; ++it) {
                const int s = it % kAsStages, c = c_begin + it / num_kb, kb = it % num_kb;
                ok &= mbar_wait(&empty_bar[s], ((it / kAsStages) - 1) & 1);
                uint8_t* sb = smem + SM::kOffB + s * SM::kStageBytes;
                mbar_arrive_expect_tx(&full_bar[s], SM::kStageBytes);
                tma_load_2d(sb, &mp.Bhi, &full_bar[s], kb * kTcBK, c * kAsChunk);
                if (kLo) tma_load_2d(sb + kAsChunk * 128, &mp.Blo, &full_bar[s], kb * kTcBK, c * kAsChunk);
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        if (lane == 0) {  // ---- MMA issuer
            constexpr uint32_t idesc = umma_idesc_h<kF16>(kTcBM, kAsChunk), idesc2n = umma_idesc_h<kF16>(kTcBM, 2 * kAsChunk);
            int it = 0;
            for (int i = 0; i < my_chunks; ++i) {
                const int buf = i & 1;
                if (i >= 2) {  // the epilogue of chunk i - 2 has read this accumulator
                    ok &= mbar_wait(&acc_empty[buf], ((i >> 1) - 1) & 1);
                    tcgen05_fence_after();
                }
                const uint32_t acc = tmem + buf * kAccCols;
                for (int kb = 0; kb < num_kb; ++kb, ++it) {
                    const int s = it % kAsStages;
                    if (i == 0) ok &= mbar_wait(&a_bar[kb], 0);
                    ok &= mbar_wait(&full_bar[s], (it / kAsStages) & 1);
                    tcgen05_fence_after();
                    const uint32_t sa = smem_u32(smem + kb * kParts * kTileABytes);
                    const uint32_t sb = smem_u32(smem + SM::kOffB + s * SM::kStageBytes);
#pragma unroll
                    for (int k = 0; k < kTcBK / 16; ++k) {
                        const uint32_t koff = k * 32;
                        const uint64_t dAhi = umma_desc_sw128(sa + koff), dBhi = umma_desc_sw128(sb + koff);
                        if (kLo) {
                            umma_bf16(acc, dAhi, dBhi, idesc2n, (kb | k) != 0);
                            umma_bf16(acc, umma_desc_sw128(sa + kTileABytes + koff), dBhi, idesc, 1);
                        } else {
                            umma_bf16(acc, dAhi, dBhi, idesc, (kb | k) != 0);
                        }
                    }
                    umma_commit(&empty_bar[s]);
                }
                umma_commit(&acc_full[buf]);
            }
        }
        __syncwarp();
    } else {
        // ---- epilogue warps: thread (row, g) owns 16 accumulator columns of its row in every chunk
        const int e = tid - 64, ew = e >> 5;
        const int quarter = warp & 3, row = quarter * 32 + lane, g = ew >> 2;
        const TileRows tr_rows(m0, a.period, a.batch_off);
        const uint32_t lane_base = tmem + ((uint32_t)(quarter * 32) << 16);
        const int Dm = a.N / 3;
        for (int i = 0; i < my_chunks; ++i) {
            const int buf = i & 1, n0 = (c_begin + i) * kAsChunk, nc = n0 + g * CPT;
            float bias_v[CPT];
            if (a.bias) {
#pragma unroll
                for (int j = 0; j < CPT; j += 4) {
                    const float4 b4 = __ldg(reinterpret_cast<const float4*>(a.bias + nc + j));
                    bias_v[j] = b4.x, bias_v[j + 1] = b4.y, bias_v[j + 2] = b4.z, bias_v[j + 3] = b4.w;
                }
            } else {
#pragma unroll
                for (int j = 0; j < CPT; ++j) bias_v[j] = 0.f;
            }
            ok &= mbar_wait(&acc_full[buf], (i >> 1) & 1);
            tcgen05_fence_after();
            if (e == 0 && i == 0) tr.mark(6);
            float v[CPT];
            tmem_ld_cols(lane_base + buf * kAccCols + g * CPT, v);
            if (kLo) {
                float hl[CPT];
                tmem_ld_cols(lane_base + buf * kAccCols + kAsChunk + g * CPT, hl);
#pragma unroll
                for (int j = 0; j < CPT; ++j) v[j] += hl[j];
            }
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_empty[buf]);  // accumulator i % 2 may be overwritten by chunk i + 2
            if (e == 0 && i == 0) tr.mark(4);
#pragma unroll
            for (int j = 0; j < CPT; ++j) v[j] += bias_v[j];
            if (a.gelu) {
#pragma unroll
                for (int j = 0; j < CPT; ++j) v[j] = gelu_erf(v[j]);
            }
            uint8_t* t_hi = smem + SM::kOffO + buf * kParts * SM::kTileOBytes;
            uint8_t* t_lo = t_hi + SM::kTileOBytes;
            int o_which = 0, o_h = 0;
            if (a.o_mode == 2) o_which = n0 / Dm, o_h = (n0 - o_which * Dm) / 64;  // this chunk is one head of Q, K or V
            if (o_which < 2) {
                stage_split<CPT, kF16>(v, t_hi, t_lo, row, g, kLo);
            } else {  // V^T: two unswizzled [64 d][64 tokens] sub-tiles (see gemm_tc.cu)
                const int sub = (row >> 6) * (kAsChunk * kHalfRows) + (g * CPT) * kHalfRows + (row & 63);
                unsigned short* uh = reinterpret_cast<unsigned short*>(t_hi) + sub;
                unsigned short* ul = reinterpret_cast<unsigned short*>(t_lo) + sub;
#pragma unroll
                for (int j = 0; j < CPT; j += 2) {
                    uint32_t h2, l2;
                    split2_h<kF16>(v[j], v[j + 1], h2, l2);
                    uh[j * kHalfRows] = (unsigned short)(h2 & 0xffffu), uh[(j + 1) * kHalfRows] = (unsigned short)(h2 >> 16);
                    if (kLo) ul[j * kHalfRows] = (unsigned short)(l2 & 0xffffu), ul[(j + 1) * kHalfRows] = (unsigned short)(l2 >> 16);
                }
            }
            // all 512 epilogue threads have staged chunk i (and, having passed this barrier for chunk i - 1 after copying chunk i - 2 out,
            // nobody still reads the buffer chunk i + ... writes next: two staging buffers, one barrier per chunk)
            if (e == 0 && i == 0) tr.mark(5);
            asm volatile("bar.sync 1, %0;" ::"n"(kTcThreads) : "memory");
            if (a.o_mode == 1) {
                tile_to_global<128>(t_hi, a.o[0], (int64_t)n0 * 2, tr_rows, a.o_row_off, 0, 0, e);
                if (kLo) tile_to_global<128>(t_lo, a.o[1], (int64_t)n0 * 2, tr_rows, a.o_row_off, 0, 0, e);
            } else if (o_which < 2) {  // Q / K: [B][heads][320][64]
                tile_to_global<128>(t_hi, a.o[2 * o_which], 0, tr_rows, 0, o_h, 0, e);
                if (kLo) tile_to_global<128>(t_lo, a.o[2 * o_which + 1], 0, tr_rows, 0, o_h, 0, e);
            } else {                   // V^T: [B][heads][64][320]
                vt_tile_to_global<kAsChunk>(t_hi, a.o[4], tr_rows, o_h, 0, e);
                if (kLo) vt_tile_to_global<kAsChunk>(t_lo, a.o[5], tr_rows, o_h, 0, e);
            }
            if (e == 0 && i == 0) tr.mark(7);  // chunk 0 copied out
        }
    }
    if (!ok && a.err) atomicExch(a.err, 1);
    tcgen05_fence_before();
    __syncthreads();
    if (tid == 0) tr.mark(3);
    if (warp == 1) tmem_dealloc(tmem, 2 * kAccCols);
}

cudaError_t tc_gemm_as_setup() {
    cudaError_t e = cudaFuncSetAttribute(gemm_as_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, AsSmem<1>::kTotal);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(gemm_as_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, AsSmem<2>::kTotal);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(gemm_as_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, AsSmem<3>::kTotal);
    return e;
}

// The plan must be a plain 64-column-tile plan (tc_plan_init with bn = 64) whose epilogue is a bf16 (hi, lo) tile store or the QKV scatter.
bool tc_gemm_as_supported(const TcGemmPlan& p) {
    const TcGemmArgs& a = p.args;
    return p.bn == 64 && a.K % kTcBK == 0 && a.K / kTcBK <= kAsMaxKb && a.N % kAsChunk == 0 && !a.conv_feat && !a.kb_per_split && !a.chain_n &&
           !a.ln_g && !a.c_on && !a.residual && !a.pos && !a.relu && (a.o_mode == 1 || a.o_mode == 2) && (a.o_mode != 2 || (a.N / 3) % 64 == 0);
}

cudaError_t tc_gemm_as_launch(const TcGemmPlan& p, int M, int nsplit, cudaStream_t s, bool pdl, int sm_count) {
    if (M <= 0) return cudaSuccess;
    if (!tc_gemm_as_supported(p)) return cudaErrorInvalidValue;
    TcGemmArgs a = p.args;
    a.M = M;
    a.chain_slices = 0, a.dup_hl = 0, a.dup_ln = 0, a.mcast = 0;
    const int row_tiles = (M + kTcBM - 1) / kTcBM, n_chunks = a.N / kAsChunk;
    // as many CTAs per row tile as fit in one wave, chunks dealt out evenly (the makespan is the largest share)
    int smax = sm_count / row_tiles;
    if (smax < 1) smax = 1;
    if (smax > n_chunks) smax = n_chunks;
    const int per_cta = (n_chunks + smax - 1) / smax;
    dim3 grid((n_chunks + per_cta - 1) / per_cta, row_tiles, 1);
    if (nsplit == 3) return launch_ex(gemm_as_kernel<3>, grid, dim3(kAsThreads), AsSmem<3>::kTotal, s, pdl, 1, p.maps, a, per_cta);
    if (nsplit == 2) return launch_ex(gemm_as_kernel<2>, grid, dim3(kAsThreads), AsSmem<2>::kTotal, s, pdl, 1, p.maps, a, per_cta);
    return launch_ex(gemm_as_kernel<1>, grid, dim3(kAsThreads), AsSmem<1>::kTotal, s, pdl, 1, p.maps, a, per_cta);
}

}  // namespace vt
