// gemm_tc.cu — the dense contractions of the ViT on 5th-generation tensor cores.
//
//   C[M,N] = epilogue( A[M,K] * W[N,K]^T + bias )      A, W: bf16 (optionally split hi+lo), accumulate fp32 in TMEM
//
// One CTA (128 threads) computes one 128 x 64 output tile:
//   warp 0 (one elected lane)  TMA producer: cp.async.bulk.tensor 128B-swizzled boxes of A and W into a 4-stage smem ring,
//                              completion counted on mbarriers (expect_tx);
//   warp 1 (one elected lane)  issues tcgen05.mma (UMMA 128x64x16, kind::f16) from smem descriptors into a TMEM accumulator,
//                              tcgen05.commit releases each stage and finally signals the accumulator;
//   all 4 warps                epilogue: tcgen05.ld (32 lanes x 32 columns per warp) -> bias / GELU / ReLU / pos-embed /
//                              residual -> fp32 store and/or bf16 (hi, lo) stores that feed the next GEMM.
//
// Precision: VT_GEMM_TCGEN05_BF16 issues A_hi*W_hi only; VT_GEMM_TCGEN05_BF16X3 adds A_hi*W_lo + A_lo*W_hi into the same
// accumulator (error ~2^-17 relative per product), which is what the 1e-3 score / exact-box parity needs; the extra tensor
// FLOPs are free at these sizes (the step is launch/latency bound).
// The 3x3 head convolution runs through the same kernel: its A operand is gathered by TMA from the [B,16,16,D] token grid with
// shifted (possibly negative) coordinates; out-of-bounds elements are zero-filled by the TMA unit = zero padding of the conv.
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

template <int NSPLIT>
__global__ void __launch_bounds__(128, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap mAhi, const __grid_constant__ CUtensorMap mAlo, const __grid_constant__ CUtensorMap mBhi,
               const __grid_constant__ CUtensorMap mBlo, const TcGemmArgs a) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t full_bar[kTcStages], empty_bar[kTcStages], accum_bar;
    __shared__ uint32_t tmem_base_s;
    using SM = TcSmem<NSPLIT>;
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int n0 = blockIdx.x * kTcBN, m0 = blockIdx.y * kTcBM;
    const int num_kb = a.K / kTcBK;
    bool ok = true;

    if (tid == 0) {
        tma_prefetch_desc(&mAhi), tma_prefetch_desc(&mBhi);
        if (NSPLIT == 3) tma_prefetch_desc(&mAlo), tma_prefetch_desc(&mBlo);
        for (int s = 0; s < kTcStages; ++s) mbar_init(&full_bar[s], 1), mbar_init(&empty_bar[s], 1);
        mbar_init(&accum_bar, 1);
        fence_barrier_init();
    }
    if (warp == 1) {
        tmem_alloc(&tmem_base_s, kTcBN);  // 64 fp32 accumulator columns
        tmem_relinquish();
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem = tmem_base_s;

    if (warp == 0) {
        if (lane == 0) {  // ---- TMA producer
            for (int kb = 0; kb < num_kb; ++kb) {
                const int s = kb % kTcStages;
                if (kb >= kTcStages) ok &= mbar_wait(&empty_bar[s], ((kb / kTcStages) - 1) & 1);
                uint8_t* st = smem + s * SM::kStageBytes;
                mbar_arrive_expect_tx(&full_bar[s], SM::kStageBytes);
                if (a.conv_feat) {  // 3x3 conv: K index = tap * feat + d; A rows are the 16x16 grid of one target
                    const int chunks = a.conv_feat / kTcBK, tap = kb / chunks, d0 = (kb % chunks) * kTcBK;
                    const int b = m0 / kNTx, y0 = (m0 % kNTx) / kMap;
                    tma_load_4d(st, &mAhi, &full_bar[s], d0, tap % 3 - 1, y0 + tap / 3 - 1, b);
                    if (NSPLIT == 3) tma_load_4d(st + kTileABytes, &mAlo, &full_bar[s], d0, tap % 3 - 1, y0 + tap / 3 - 1, b);
                } else {
                    tma_load_2d(st, &mAhi, &full_bar[s], kb * kTcBK, m0);
                    if (NSPLIT == 3) tma_load_2d(st + kTileABytes, &mAlo, &full_bar[s], kb * kTcBK, m0);
                }
                uint8_t* sb = st + SM::kParts * kTileABytes;
                tma_load_2d(sb, &mBhi, &full_bar[s], kb * kTcBK, n0);
                if (NSPLIT == 3) tma_load_2d(sb + kTileBBytes, &mBlo, &full_bar[s], kb * kTcBK, n0);
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

    // ---- epilogue: thread t owns accumulator row t (TMEM lane t)
    ok &= mbar_wait(&accum_bar, 0);
    tcgen05_fence_after();
    const int m = m0 + tid;
    const bool row_ok = m < a.M;
    const int64_t crow = a.C ? ((int64_t)((m / a.c_rows_in) * a.c_rows_stride + a.c_row_off + (m % a.c_rows_in))) * a.ldc : 0;
#pragma unroll
    for (int c0 = 0; c0 < kTcBN; c0 += 32) {
        float v[32];
        tmem_ld_32x32(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
        if (row_ok) {
            const int n = n0 + c0;
            if (a.bias) {
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    const float4 b4 = __ldg(reinterpret_cast<const float4*>(a.bias + n + j));
                    v[j] += b4.x, v[j + 1] += b4.y, v[j + 2] += b4.z, v[j + 3] += b4.w;
                }
            }
            if (a.gelu) {
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = gelu_erf(v[j]);
            }
            if (a.relu) {
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
            }
            if (a.pos) {
                const float* pp = a.pos + (int64_t)(m % a.pos_rows) * a.N + n;
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    const float4 p4 = __ldg(reinterpret_cast<const float4*>(pp + j));
                    v[j] += p4.x, v[j + 1] += p4.y, v[j + 2] += p4.z, v[j + 3] += p4.w;
                }
            }
            if (a.C) {
                float* cp = a.C + crow + n;
                if (a.residual) {
#pragma unroll
                    for (int j = 0; j < 32; j += 4) {
                        const float4 r4 = *reinterpret_cast<const float4*>(cp + j);
                        v[j] += r4.x, v[j + 1] += r4.y, v[j + 2] += r4.z, v[j + 3] += r4.w;
                    }
                }
#pragma unroll
                for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(cp + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
            }
            if (a.qkv_heads) {
                uint32_t hi[16], lo[16];
#pragma unroll
                for (int j = 0; j < 32; j += 2) {
                    __nv_bfloat16 h0, l0, h1, l1;
                    split_bf16(v[j], h0, l0), split_bf16(v[j + 1], h1, l1);
                    hi[j >> 1] = pack_bf16(h0, h1), lo[j >> 1] = pack_bf16(l0, l1);
                }
                const int Dm = a.N / 3, which = n0 / Dm, hh = (n0 % Dm) / kTcBN;
                const int64_t bh = (int64_t)(m / kNTok) * a.qkv_heads + hh;
                const int tok = m % kNTok;
                if (which < 2) {  // Q / K: [bh][tok][64], this thread's 32 columns are contiguous
                    __nv_bfloat16* dh_ = (which == 0 ? a.Qhi : a.Khi) + (bh * kNTok + tok) * kTcBN + c0;
                    __nv_bfloat16* dl_ = (which == 0 ? a.Qlo : a.Klo) + (bh * kNTok + tok) * kTcBN + c0;
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        reinterpret_cast<uint4*>(dh_)[j] = make_uint4(hi[4 * j], hi[4 * j + 1], hi[4 * j + 2], hi[4 * j + 3]);
                        reinterpret_cast<uint4*>(dl_)[j] = make_uint4(lo[4 * j], lo[4 * j + 1], lo[4 * j + 2], lo[4 * j + 3]);
                    }
                } else {  // V^T: [bh][d][tok]; a warp's 32 rows are 32 consecutive tokens -> 64-byte coalesced stores per d
                    unsigned short* vh = reinterpret_cast<unsigned short*>(a.Vthi) + (bh * kTcBN + c0) * kNTok + tok;
                    unsigned short* vl = reinterpret_cast<unsigned short*>(a.Vtlo) + (bh * kTcBN + c0) * kNTok + tok;
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        vh[(int64_t)j * kNTok] = (unsigned short)((j & 1) ? (hi[j >> 1] >> 16) : (hi[j >> 1] & 0xffffu));
                        vl[(int64_t)j * kNTok] = (unsigned short)((j & 1) ? (lo[j >> 1] >> 16) : (lo[j >> 1] & 0xffffu));
                    }
                }
            }
            if (a.Ohi) {
                uint32_t hi[16], lo[16];
#pragma unroll
                for (int j = 0; j < 32; j += 2) {
                    __nv_bfloat16 h0, l0, h1, l1;
                    split_bf16(v[j], h0, l0), split_bf16(v[j + 1], h1, l1);
                    hi[j >> 1] = pack_bf16(h0, h1), lo[j >> 1] = pack_bf16(l0, l1);
                }
                uint4* oh = reinterpret_cast<uint4*>(a.Ohi + (int64_t)m * a.ldo + n);
                uint4* ol = reinterpret_cast<uint4*>(a.Olo + (int64_t)m * a.ldo + n);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    oh[j] = make_uint4(hi[4 * j], hi[4 * j + 1], hi[4 * j + 2], hi[4 * j + 3]);
                    ol[j] = make_uint4(lo[4 * j], lo[4 * j + 1], lo[4 * j + 2], lo[4 * j + 3]);
                }
            }
        }
    }
    if (!ok && a.err) atomicExch(a.err, 1);
    tcgen05_fence_before();
    __syncthreads();
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
    p->args.c_rows_in = 1 << 30, p->args.pos_rows = 1;
    return ok;
}

cudaError_t tc_gemm_setup() {
    cudaError_t e = cudaFuncSetAttribute(gemm_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, TcSmem<1>::kTotal);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(gemm_tc_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, TcSmem<3>::kTotal);
}

cudaError_t tc_gemm_launch(const TcGemmPlan& p, int M, int nsplit, cudaStream_t s) {
    if (M <= 0) return cudaSuccess;
    TcGemmArgs a = p.args;
    a.M = M;
    dim3 grid(a.N / kTcBN, (M + kTcBM - 1) / kTcBM);
    if (nsplit == 3)
        gemm_tc_kernel<3><<<grid, 128, TcSmem<3>::kTotal, s>>>(p.mAhi, p.mAlo, p.mBhi, p.mBlo, a);
    else
        gemm_tc_kernel<1><<<grid, 128, TcSmem<1>::kTotal, s>>>(p.mAhi, p.mAlo, p.mBhi, p.mBlo, a);
    return cudaGetLastError();
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
        cudaError_t e = tc_gemm_launch(plan, M, nsplit, 0);
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
