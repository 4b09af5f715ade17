// gemm_tc.cu — the dense contractions of the ViT on 5th-generation tensor cores.
//
//   C[M,N] = epilogue( A[M,K] * W[N,K]^T + bias )      A, W: bf16 (optionally split hi+lo), accumulate fp32 in TMEM
//
// One CTA (512 threads) computes one 128 x 64 output tile:
//   warp 0 (one elected lane)  TMA producer: cp.async.bulk.tensor 128B-swizzled boxes of A and W into a 4-stage smem ring,
//                              completion counted on mbarriers (expect_tx);
//   warp 1 (one elected lane)  issues tcgen05.mma (UMMA 128x64x16, kind::f16) from smem descriptors into a TMEM accumulator,
//                              tcgen05.commit releases each stage and finally signals the accumulator;
//   all 16 warps               epilogue: four threads per accumulator row (tcgen05.ld.x16: warp w reads lane quarter w % 4, columns
//                              16 (w / 4) ..) -> bias / GELU / ReLU / pos-embed / residual -> fp32 store and/or bf16 (hi, lo) stores
//                              that feed the next GEMM, optionally LayerNorm of the full row through a cluster exchange.  The
//                              epilogue is ALU / latency bound, which is why it is spread over 16 warps instead of 4.
//
// Precision: VT_GEMM_TCGEN05_BF16 issues A_hi*W_hi only; VT_GEMM_TCGEN05_BF16X3 adds A_hi*W_lo + A_lo*W_hi into the same
// accumulator (error ~2^-17 relative per product), which is what the 1e-3 score / exact-box parity needs; the extra tensor
// FLOPs are free at these sizes (the step is launch/latency bound).
// The 3x3 head convolution runs through the same kernel: its A operand is gathered by TMA from the [B,16,16,D] token grid with
// shifted (possibly negative) coordinates; out-of-bounds elements are zero-filled by the TMA unit = zero padding of the conv.
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "tc_common.cuh"
#include "vt_internal.h"

namespace vt {

using namespace tc;

constexpr int kTcBM = 128, kTcBN = 64, kTcBK = 64, kTcStages = 4;
constexpr int kTileABytes = kTcBM * kTcBK * 2;  // 16 KB, one precision part
constexpr int kTileBBytes = kTcBN * kTcBK * 2;  //  8 KB

template <int NSPLIT>
struct TcSmem {
    static constexpr int kParts = NSPLIT == 3 ? 2 : 1;
    static constexpr int kStageBytes = kParts * (kTileABytes + kTileBBytes);
    static constexpr int kTotal = kTcStages * kStageBytes + 1024;  // + alignment slack
};

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.f + erff(x * 0.70710678118654752440f)); }

// LayerNorm fused into the epilogue of the GEMMs that produce the residual stream (patch-embed, proj, FC2): the N / 64 CTAs
// that hold the column tiles of one 128-row tile form a thread-block cluster; thread t of every CTA owns row t, computes the
// (sum, M2) of its 64 columns from registers, pushes the pair into every peer's shared memory (st.shared::cluster), and after
// one cluster barrier combines the partials in rank order (Chan's parallel variance) — every CTA gets bit-identical
// statistics — normalises its own 64 columns and stores the bf16 (hi, lo) A operand of the next GEMM.
constexpr int kMaxLnCluster = 8;
constexpr int kTcThreads = 512;                       // 16 warps: warp w reads TMEM lane quarter w % 4, column group w / 4
constexpr int kTcColGroups = kTcThreads / kTcBM;      // 4 threads per accumulator row
constexpr int kTcColsPerThread = kTcBN / kTcColGroups;  // 16 columns each

__device__ __forceinline__ void split_store16(const float (&v)[16], __nv_bfloat16* hi_dst, __nv_bfloat16* lo_dst) {
    uint32_t hi[8], lo[8];
#pragma unroll
    for (int j = 0; j < 16; j += 2) {
        __nv_bfloat16 h0, l0, h1, l1;
        split_bf16(v[j], h0, l0), split_bf16(v[j + 1], h1, l1);
        hi[j >> 1] = pack_bf16(h0, h1), lo[j >> 1] = pack_bf16(l0, l1);
    }
    uint4* oh = reinterpret_cast<uint4*>(hi_dst);
    uint4* ol = reinterpret_cast<uint4*>(lo_dst);
    oh[0] = make_uint4(hi[0], hi[1], hi[2], hi[3]), oh[1] = make_uint4(hi[4], hi[5], hi[6], hi[7]);
    ol[0] = make_uint4(lo[0], lo[1], lo[2], lo[3]), ol[1] = make_uint4(lo[4], lo[5], lo[6], lo[7]);
}

template <int NSPLIT>
__global__ void __launch_bounds__(kTcThreads, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap mAhi, const __grid_constant__ CUtensorMap mAlo, const __grid_constant__ CUtensorMap mBhi,
               const __grid_constant__ CUtensorMap mBlo, const TcGemmArgs a) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t full_bar[kTcStages], empty_bar[kTcStages], accum_bar;
    __shared__ uint32_t tmem_base_s;
    __shared__ float2 ln_part[kMaxLnCluster][kTcColGroups][kTcBM];
    __shared__ unsigned long long* trace_slot;
    using SM = TcSmem<NSPLIT>;
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int row = (warp & 3) * 32 + lane;   // accumulator row inside the tile = TMEM lane
    const int g = warp >> 2;                  // column group: columns 16 g .. 16 g + 15 of the tile
    const int n0 = blockIdx.x * kTcBN, m0 = blockIdx.y * kTcBM;
    const int num_kb = a.K / kTcBK;
    const int npre = num_kb < kTcStages ? num_kb : kTcStages;
    bool ok = true;
    TraceRec tr;
    tr.begin(&trace_slot, a.trace, a.trace_id);

    // ---- prologue: independent of the preceding kernel, overlaps its tail under PDL
    if (tid == 0) {
        tma_prefetch_desc(&mAhi), tma_prefetch_desc(&mBhi);
        if (NSPLIT == 3) tma_prefetch_desc(&mAlo), tma_prefetch_desc(&mBlo);
        for (int s = 0; s < kTcStages; ++s) mbar_init(&full_bar[s], 1), mbar_init(&empty_bar[s], 1);
        mbar_init(&accum_bar, 1);
        fence_barrier_init();
        // the weights never depend on the preceding kernel: start streaming them right away
        for (int kb = 0; kb < npre; ++kb) {
            uint8_t* sb = smem + kb * SM::kStageBytes + SM::kParts * kTileABytes;
            mbar_arrive_expect_tx(&full_bar[kb], SM::kStageBytes);
            tma_load_2d(sb, &mBhi, &full_bar[kb], kb * kTcBK, n0);
            if (NSPLIT == 3) tma_load_2d(sb + kTileBBytes, &mBlo, &full_bar[kb], kb * kTcBK, n0);
        }
    }
    if (warp == 1) {
        tmem_alloc(&tmem_base_s, kTcBN);  // 64 fp32 accumulator columns
        tmem_relinquish();
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem = tmem_base_s;
    if (a.ln_g) cluster_arrive_release();  // "this CTA is running": peers may address its shared memory after the matching wait

    pdl_wait();               // the activations (A, residual) are complete and visible from here on
    if (tid == 0) tr.mark(2);
    pdl_launch_dependents();  // let the next kernel of the chain run its prologue under our main loop

    if (warp == 0) {
        if (lane == 0) {  // ---- TMA producer
            for (int kb = 0; kb < num_kb; ++kb) {
                const int s = kb % kTcStages;
                uint8_t* st = smem + s * SM::kStageBytes;
                if (kb >= kTcStages) {
                    ok &= mbar_wait(&empty_bar[s], ((kb / kTcStages) - 1) & 1);
                    mbar_arrive_expect_tx(&full_bar[s], SM::kStageBytes);
                    uint8_t* sb = st + SM::kParts * kTileABytes;
                    tma_load_2d(sb, &mBhi, &full_bar[s], kb * kTcBK, n0);
                    if (NSPLIT == 3) tma_load_2d(sb + kTileBBytes, &mBlo, &full_bar[s], kb * kTcBK, n0);
                }
                if (a.conv_feat) {  // 3x3 conv: K index = tap * feat + d; A rows are the 16x16 grid of one target
                    const int chunks = a.conv_feat / kTcBK, tap = kb / chunks, d0 = (kb % chunks) * kTcBK;
                    const int b = m0 / kNTx, y0 = (m0 % kNTx) / kMap;
                    tma_load_4d(st, &mAhi, &full_bar[s], d0, tap % 3 - 1, y0 + tap / 3 - 1, b);
                    if (NSPLIT == 3) tma_load_4d(st + kTileABytes, &mAlo, &full_bar[s], d0, tap % 3 - 1, y0 + tap / 3 - 1, b);
                } else {
                    tma_load_2d(st, &mAhi, &full_bar[s], kb * kTcBK, m0);
                    if (NSPLIT == 3) tma_load_2d(st + kTileABytes, &mAlo, &full_bar[s], kb * kTcBK, m0);
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        if (lane == 0) {  // ---- MMA issuer
            constexpr uint32_t idesc = umma_idesc_bf16(kTcBM, kTcBN);
            for (int kb = 0; kb < num_kb; ++kb) {
                const int s = kb % kTcStages;
                ok &= mbar_wait(&full_bar[s], (kb / kTcStages) & 1);
                tcgen05_fence_after();
                const uint32_t sa = smem_u32(smem + s * SM::kStageBytes);
                const uint32_t sb = sa + SM::kParts * kTileABytes;
#pragma unroll
                for (int k = 0; k < kTcBK / 16; ++k) {
                    const uint32_t koff = k * 32;  // 16 bf16 = 32 bytes inside the 128-byte swizzle row
                    const uint64_t dAhi = umma_desc_sw128(sa + koff), dBhi = umma_desc_sw128(sb + koff);
                    umma_bf16(tmem, dAhi, dBhi, idesc, (kb | k) != 0);
                    if (NSPLIT == 3) {
                        const uint64_t dAlo = umma_desc_sw128(sa + kTileABytes + koff), dBlo = umma_desc_sw128(sb + kTileBBytes + koff);
                        umma_bf16(tmem, dAhi, dBlo, idesc, 1);
                        umma_bf16(tmem, dAlo, dBhi, idesc, 1);
                    }
                }
                umma_commit(&empty_bar[s]);  // stage reusable once these MMAs have read it
            }
            umma_commit(&accum_bar);
        }
        __syncwarp();
    }

    // ---- epilogue: thread (row, g) owns 16 accumulator columns of its row; 16 warps keep the ALU / store work short
    const int m = m0 + row, nc = n0 + g * kTcColsPerThread;
    const bool row_ok = m < a.M;
    const int64_t crow = a.C ? ((int64_t)((m / a.c_rows_in) * a.c_rows_stride + a.c_row_off + (m % a.c_rows_in))) * a.ldc : 0;
    float v[kTcColsPerThread];
    if (a.residual && row_ok) {  // the residual rows do not depend on the accumulator: fetch them while the MMAs run
        const float4* cp = reinterpret_cast<const float4*>(a.C + crow + nc);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float4 r4 = cp[j];
            v[4 * j] = r4.x, v[4 * j + 1] = r4.y, v[4 * j + 2] = r4.z, v[4 * j + 3] = r4.w;
        }
    } else {
#pragma unroll
        for (int j = 0; j < kTcColsPerThread; ++j) v[j] = 0.f;
    }
    ok &= mbar_wait(&accum_bar, 0);
    tcgen05_fence_after();
    if (tid == 0) tr.mark(6);
    {
        float acc[kTcColsPerThread];
        tmem_ld_32x16(tmem + ((uint32_t)((warp & 3) * 32) << 16) + g * kTcColsPerThread, acc);
        if (tid == 0) tr.mark(4);
        if (a.bias) {
#pragma unroll
            for (int j = 0; j < 16; j += 4) {
                const float4 b4 = __ldg(reinterpret_cast<const float4*>(a.bias + nc + j));
                acc[j] += b4.x, acc[j + 1] += b4.y, acc[j + 2] += b4.z, acc[j + 3] += b4.w;
            }
        }
        if (a.gelu) {
#pragma unroll
            for (int j = 0; j < 16; ++j) acc[j] = gelu_erf(acc[j]);
        }
        if (a.relu) {
#pragma unroll
            for (int j = 0; j < 16; ++j) acc[j] = fmaxf(acc[j], 0.f);
        }
        if (a.pos && row_ok) {
            const float* pp = a.pos + (int64_t)(m % a.pos_rows) * a.N + nc;
#pragma unroll
            for (int j = 0; j < 16; j += 4) {
                const float4 p4 = __ldg(reinterpret_cast<const float4*>(pp + j));
                acc[j] += p4.x, acc[j + 1] += p4.y, acc[j + 2] += p4.z, acc[j + 3] += p4.w;
            }
        }
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = acc[j] + v[j];  // + residual (0 when off): same order as the unfused path
    }
    if (tid == 0) tr.mark(5);
    for (int rep = 0; rep < (a.trace_id >= 100 ? 2 : 1); ++rep) {  // diagnostics: trace ids >= 100 run the store block twice
    if (row_ok) {
        if (a.C) {
            float4* cp = reinterpret_cast<float4*>(a.C + crow + nc);
#pragma unroll
            for (int j = 0; j < 4; ++j) cp[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        }
        if (a.qkv_heads) {
            const int Dm = a.N / 3, which = n0 / Dm, hh = (n0 % Dm) / kTcBN;
            const int64_t bh = (int64_t)(m / kNTok) * a.qkv_heads + hh;
            const int tok = m % kNTok, d0 = g * kTcColsPerThread;
            if (which < 2) {  // Q / K: [bh][tok][64]
                const int64_t o = (bh * kNTok + tok) * kTcBN + d0;
                split_store16(v, (which == 0 ? a.Qhi : a.Khi) + o, (which == 0 ? a.Qlo : a.Klo) + o);
            } else {  // V^T: [bh][d][tok]; a warp's 32 rows are 32 consecutive tokens -> 64-byte coalesced stores per d
                __nv_bfloat16* vh = a.Vthi + (bh * kTcBN + d0) * kNTok + tok;
                __nv_bfloat16* vl = a.Vtlo + (bh * kTcBN + d0) * kNTok + tok;
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    __nv_bfloat16 h, l;
                    split_bf16(v[j], h, l);
                    vh[(int64_t)j * kNTok] = h, vl[(int64_t)j * kNTok] = l;
                }
            }
        }
        if (a.Ohi) split_store16(v, a.Ohi + (int64_t)m * a.ldo + nc, a.Olo + (int64_t)m * a.ldo + nc);
    }
    if (tid == 0) tr.mark(rep ? 4 : 7);
    }
    if (a.ln_g) {  // ---- fused LayerNorm over the full row (N columns = cluster of N / 64 CTAs x 4 column groups)
        const uint32_t nct = cluster_nctarank(), me = cluster_ctarank();
        float s = 0.f;
#pragma unroll
        for (int j = 0; j < 16; ++j) s += v[j];
        const float mu = s * (1.f / 16.f);
        float m2 = 0.f;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const float d = v[j] - mu;
            m2 = fmaf(d, d, m2);
        }
        const uint32_t mine = smem_u32(&ln_part[me][g][row]);
        cluster_wait_acquire();  // every CTA of the cluster has started (arrive in the prologue)
        for (uint32_t r = 0; r < nct; ++r) st_shared_cluster_f2(cluster_map_shared(mine, r), s, m2);
        cluster_arrive_release();
        cluster_wait_acquire();
        float tot = 0.f;
        for (uint32_t r = 0; r < nct; ++r)
#pragma unroll
            for (int q = 0; q < kTcColGroups; ++q) tot += ln_part[r][q][row].x;
        const float mean = tot / (float)a.N;
        float M2 = 0.f;
        for (uint32_t r = 0; r < nct; ++r)
#pragma unroll
            for (int q = 0; q < kTcColGroups; ++q) {
                const float2 p = ln_part[r][q][row];
                const float d = p.x * (1.f / 16.f) - mean;
                M2 += p.y + 16.f * d * d;
            }
        const float rstd = 1.f / sqrtf(M2 / (float)a.N + 1e-6f);
        const int mi = m % a.ln_rows_in;
        if (row_ok && mi >= a.ln_skip) {
            const int64_t lrow = (int64_t)(m / a.ln_rows_in) * a.ln_rows_stride + a.ln_row_off + mi;
            float y[16];
#pragma unroll
            for (int j = 0; j < 16; j += 4) {
                const float4 g4 = __ldg(reinterpret_cast<const float4*>(a.ln_g + nc + j)), b4 = __ldg(reinterpret_cast<const float4*>(a.ln_b + nc + j));
                y[j] = (v[j] - mean) * rstd * g4.x + b4.x, y[j + 1] = (v[j + 1] - mean) * rstd * g4.y + b4.y;
                y[j + 2] = (v[j + 2] - mean) * rstd * g4.z + b4.z, y[j + 3] = (v[j + 3] - mean) * rstd * g4.w + b4.w;
            }
            split_store16(y, a.ln_hi + lrow * a.N + nc, a.ln_lo + lrow * a.N + nc);
        }
    }
    if (!ok && a.err) atomicExch(a.err, 1);
    tcgen05_fence_before();
    __syncthreads();
    if (tid == 0) tr.mark(3);
    if (warp == 1) tmem_dealloc(tmem, kTcBN);
}

// ---- host side ------------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

// bf16 tensor [dims...] (dims[0] innermost, contiguous), box per dim, 128-byte swizzle, zero OOB fill
bool tc_make_map(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes, const uint32_t* box) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) {
        set_error("cuTensorMapEncodeTiled is not available from the driver");
        return false;
    }
    cuuint64_t gd[5], gs[5];
    cuuint32_t bx[5], es[5];
    for (int i = 0; i < rank; ++i) gd[i] = dims[i], bx[i] = box[i], es[i] = 1;
    for (int i = 0; i + 1 < rank; ++i) gs[i] = strides_bytes[i];
    CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed with CUresult %d (rank %d, dims %llu x %llu)", (int)r, rank, (unsigned long long)dims[0],
                  (unsigned long long)(rank > 1 ? dims[1] : 1));
        return false;
    }
    return true;
}

bool tc_make_map_2d(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint32_t box_rows) {
    const uint64_t dims[2] = {cols, rows}, strides[1] = {cols * 2};
    const uint32_t box[2] = {(uint32_t)kTcBK, box_rows};
    return tc_make_map(out, base, 2, dims, strides, box);
}

bool tc_plan_init(TcGemmPlan* p, const __nv_bfloat16* Ahi, const __nv_bfloat16* Alo, uint64_t a_rows, const __nv_bfloat16* Whi,
                  const __nv_bfloat16* Wlo, int N, int K, int conv_feat, int conv_batch) {
    memset(p, 0, sizeof(*p));
    if (N % kTcBN || K % kTcBK || (conv_feat && conv_feat % kTcBK)) {
        set_error("tcgen05 GEMM needs N %% 64 == 0 and K %% 64 == 0 (N=%d K=%d)", N, K);
        return false;
    }
    bool ok = true;
    if (conv_feat) {  // A = token grid [batch][16][16][feat]
        const uint64_t dims[4] = {(uint64_t)conv_feat, kMap, kMap, (uint64_t)conv_batch};
        const uint64_t strides[3] = {(uint64_t)conv_feat * 2, (uint64_t)conv_feat * 2 * kMap, (uint64_t)conv_feat * 2 * kMap * kMap};
        const uint32_t box[4] = {(uint32_t)kTcBK, kMap, kTcBM / kMap, 1};
        ok &= tc_make_map(&p->mAhi, Ahi, 4, dims, strides, box);
        ok &= tc_make_map(&p->mAlo, Alo ? Alo : Ahi, 4, dims, strides, box);
    } else {
        ok &= tc_make_map_2d(&p->mAhi, Ahi, a_rows, K, kTcBM);
        ok &= tc_make_map_2d(&p->mAlo, Alo ? Alo : Ahi, a_rows, K, kTcBM);
    }
    ok &= tc_make_map_2d(&p->mBhi, Whi, N, K, kTcBN);
    ok &= tc_make_map_2d(&p->mBlo, Wlo ? Wlo : Whi, N, K, kTcBN);
    p->args.N = N, p->args.K = K, p->args.conv_feat = conv_feat;
    p->args.c_rows_in = 1 << 30, p->args.pos_rows = 1, p->args.ln_rows_in = 1 << 30;
    return ok;
}

cudaError_t tc_gemm_setup() {
    cudaError_t e = cudaFuncSetAttribute(gemm_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, TcSmem<1>::kTotal);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(gemm_tc_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, TcSmem<3>::kTotal);
}

cudaError_t tc_gemm_launch(const TcGemmPlan& p, int M, int nsplit, cudaStream_t s, bool pdl) {
    if (M <= 0) return cudaSuccess;
    TcGemmArgs a = p.args;
    a.M = M;
    const dim3 grid(a.N / kTcBN, (M + kTcBM - 1) / kTcBM);
    int cluster_x = 1;
    if (a.ln_g) {
        cluster_x = a.N / kTcBN;
        if (cluster_x > kMaxLnCluster) return cudaErrorInvalidValue;
    }
    static const bool dup = getenv("VT_B200_DUP") != nullptr;  // diagnostics: launch twice to compare cold / warm instruction fetch
    if (dup && nsplit == 3) launch_ex(gemm_tc_kernel<3>, grid, dim3(kTcThreads), TcSmem<3>::kTotal, s, pdl, cluster_x, p.mAhi, p.mAlo, p.mBhi, p.mBlo, a);
    if (nsplit == 3) return launch_ex(gemm_tc_kernel<3>, grid, dim3(kTcThreads), TcSmem<3>::kTotal, s, pdl, cluster_x, p.mAhi, p.mAlo, p.mBhi, p.mBlo, a);
    return launch_ex(gemm_tc_kernel<1>, grid, dim3(kTcThreads), TcSmem<1>::kTotal, s, pdl, cluster_x, p.mAhi, p.mAlo, p.mBhi, p.mBlo, a);
}

// fp32 -> bf16 (hi, lo) split of a dense buffer
__global__ void split_bf16_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        __nv_bfloat16 h, l;
        split_bf16(x[i], h, l);
        hi[i] = h, lo[i] = l;
    }
}
cudaError_t launch_split_bf16(const float* x, __nv_bfloat16* hi, __nv_bfloat16* lo, size_t n, cudaStream_t s) {
    if (!n) return cudaSuccess;
    size_t blocks = (n + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    split_bf16_kernel<<<(unsigned)blocks, 256, 0, s>>>(x, hi, lo, n);
    return cudaGetLastError();
}

}  // namespace vt

// ---- diagnostics entry: one GEMM on host data through the tensor-core kernel (unit tests) -----------------------------------
extern "C" vt_status vt_debug_gemm(int32_t device, int32_t M, int32_t N, int32_t K, const float* A, const float* W, const float* bias,
                                   int32_t nsplit, int32_t gelu, float* C_out, int32_t* err_out) {
    using namespace vt;
    if (!A || !W || !C_out || M <= 0 || (nsplit != 1 && nsplit != 3)) return VT_ERR_INVALID;
    VT_CUDA(cudaSetDevice(device));
    VT_CUDA(tc_gemm_setup());
    float *dA = nullptr, *dW = nullptr, *dB = nullptr, *dC = nullptr;
    __nv_bfloat16 *Ahi = nullptr, *Alo = nullptr, *Whi = nullptr, *Wlo = nullptr;
    int* dErr = nullptr;
    const size_t na = (size_t)M * K, nw = (size_t)N * K;
    VT_CUDA(cudaMalloc(&dA, na * 4)); VT_CUDA(cudaMalloc(&dW, nw * 4)); VT_CUDA(cudaMalloc(&dC, (size_t)M * N * 4));
    VT_CUDA(cudaMalloc(&Ahi, na * 2)); VT_CUDA(cudaMalloc(&Alo, na * 2)); VT_CUDA(cudaMalloc(&Whi, nw * 2)); VT_CUDA(cudaMalloc(&Wlo, nw * 2));
    VT_CUDA(cudaMalloc(&dErr, 4)); VT_CUDA(cudaMemset(dErr, 0, 4));
    VT_CUDA(cudaMemcpy(dA, A, na * 4, cudaMemcpyHostToDevice)); VT_CUDA(cudaMemcpy(dW, W, nw * 4, cudaMemcpyHostToDevice));
    if (bias) {
        VT_CUDA(cudaMalloc(&dB, (size_t)N * 4));
        VT_CUDA(cudaMemcpy(dB, bias, (size_t)N * 4, cudaMemcpyHostToDevice));
    }
    VT_CUDA(cudaMemset(dC, 0, (size_t)M * N * 4));
    VT_CUDA(launch_split_bf16(dA, Ahi, Alo, na, 0)); VT_CUDA(launch_split_bf16(dW, Whi, Wlo, nw, 0));
    TcGemmPlan plan;
    vt_status st = VT_OK;
    if (!tc_plan_init(&plan, Ahi, Alo, M, Whi, Wlo, N, K, 0, 0)) st = VT_ERR_CUDA;
    if (st == VT_OK) {
        plan.args.bias = dB, plan.args.C = dC, plan.args.ldc = N, plan.args.gelu = gelu, plan.args.err = dErr;
        cudaError_t e = tc_gemm_launch(plan, M, nsplit, 0, false);
        if (e == cudaSuccess) e = cudaDeviceSynchronize();
        if (e != cudaSuccess) {
            set_error("tcgen05 GEMM failed: %s", cudaGetErrorString(e));
            st = VT_ERR_CUDA;
        }
    }
    if (st == VT_OK) {
        VT_CUDA(cudaMemcpy(C_out, dC, (size_t)M * N * 4, cudaMemcpyDeviceToHost));
        int herr = 0;
        VT_CUDA(cudaMemcpy(&herr, dErr, 4, cudaMemcpyDeviceToHost));
        if (err_out) *err_out = herr;
    }
    cudaFree(dA), cudaFree(dW), cudaFree(dB), cudaFree(dC), cudaFree(Ahi), cudaFree(Alo), cudaFree(Whi), cudaFree(Wlo), cudaFree(dErr);
    return st;
}
