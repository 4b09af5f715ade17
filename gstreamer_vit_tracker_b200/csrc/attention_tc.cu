// attention_tc.cu — joint attention over the 320-token sequence on tcgen05 tensor cores (head_dim 64).
//
// One CTA (512 threads) per (128-query tile, head, target):
//   TMA        Q tile [128 x 64], K [320 x 64] and V^T [64 x 320] (bf16 hi/lo parts, 128B swizzle) -> shared memory
//   tcgen05    S[128 x 320] = Q K^T  -> TMEM columns 0..319 (two UMMAs per K-step: N = 256 and N = 64)
//   softmax    FOUR threads per query row (warp w reads TMEM lane quarter w % 4; column group g = w / 4): the exp work is what
//              bounds this kernel, so it is spread over all 16 warps.  Per 64-key chunk thread (row, g) owns 16 keys:
//              tcgen05.ld.x16 -> ex2(s * k - max * k) -> bf16 (hi, lo) split -> 2 x 16 B stores into the 128B-swizzled K-major
//              P tile (double buffered); row max / row sum partials are combined through shared memory in a fixed order
//   tcgen05    O[128 x 64] += P_chunk V_chunk -> TMEM columns 320..383, overlapped with the next chunk's softmax
//   epilogue   O / sum -> bf16 (hi, lo) rows of the proj GEMM's A operand (16 columns per thread).
// The P buffers alias the Q/K staging area once S is complete, which keeps the CTA at ~192 KB of shared memory.
#include "tc_common.cuh"
#include "vt_internal.h"

namespace vt {

using namespace tc;

constexpr int kDh = 64;
constexpr int kQTile = 128;
constexpr int kKeyChunk = 64;
constexpr int kNChunks = kNTok / kKeyChunk;       // 5
constexpr int kQBytes = kQTile * kDh * 2;         // 16 KB
constexpr int kKBytes = kNTok * kDh * 2;          // 40 KB
constexpr int kVBytes = kDh * kNTok * 2;          // 40 KB (5 blocks of [64 x 64])
constexpr int kPBytes = kQTile * kKeyChunk * 2;   // 16 KB
constexpr uint32_t kTmemCols = 512;
constexpr uint32_t kColO = 320;
constexpr int kAttThreads = 512;
constexpr int kColGroups = kAttThreads / kQTile;  // 4 threads per query row
constexpr int kKeysPerThread = kKeyChunk / kColGroups;  // 16 keys of every chunk

template <int NSPLIT>
struct AttSmem {
    static constexpr int kParts = NSPLIT == 3 ? 2 : 1;
    static constexpr int kQK = kParts * (kQBytes + kKBytes);
    static constexpr int kP = 2 * kParts * kPBytes;
    static constexpr int kRegion1 = kQK > kP ? kQK : kP;
    static constexpr int kV = kParts * kVBytes;
    static constexpr int kTotal = kRegion1 + kV + 1024;
};

template <int NSPLIT>
__global__ void __launch_bounds__(kAttThreads, 1)
attention_tc_kernel(const __grid_constant__ CUtensorMap mQhi, const __grid_constant__ CUtensorMap mQlo, const __grid_constant__ CUtensorMap mKhi,
                    const __grid_constant__ CUtensorMap mKlo, const __grid_constant__ CUtensorMap mVhi, const __grid_constant__ CUtensorMap mVlo,
                    __nv_bfloat16* __restrict__ out_hi, __nv_bfloat16* __restrict__ out_lo, int D, int heads, int* err,
                    unsigned long long* trace) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bar_qk, bar_v, bar_s, bar_p[2], bar_o;
    __shared__ uint32_t tmem_base_s;
    __shared__ float red[kColGroups][kQTile];  // row-max partials, then row-sum partials
    using SM = AttSmem<NSPLIT>;
    constexpr int P = SM::kParts;
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sQ = smem;                       // [P][128 x 128B]
    uint8_t* sK = smem + P * kQBytes;         // [P][320 x 128B]
    uint8_t* sP = smem;                       // [2 bufs][P][128 x 128B], aliases Q/K after S is complete
    uint8_t* sV = smem + SM::kRegion1;        // [P][5 blocks][64 x 128B]

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int row = (warp & 3) * 32 + lane;   // query row inside the tile = TMEM lane
    const int g = warp >> 2;                  // column group
    const int q0 = blockIdx.x * kQTile, h = blockIdx.y, b = blockIdx.z;
    const int bh = b * heads + h;
    const bool q_ok = q0 + row < kNTok;       // uniform per warp (320 = 2 * 128 + 64)
    bool ok = true;
    __shared__ unsigned long long* trace_slot;
    TraceRec tr;
    tr.begin(&trace_slot, trace, 10);

    if (tid == 0) {
        tma_prefetch_desc(&mQhi), tma_prefetch_desc(&mKhi), tma_prefetch_desc(&mVhi);
        if (P == 2) tma_prefetch_desc(&mQlo), tma_prefetch_desc(&mKlo), tma_prefetch_desc(&mVlo);
        mbar_init(&bar_qk, 1), mbar_init(&bar_v, 1), mbar_init(&bar_s, 1), mbar_init(&bar_p[0], 1), mbar_init(&bar_p[1], 1), mbar_init(&bar_o, 1);
        fence_barrier_init();
    }
    if (warp == 1) {
        tmem_alloc(&tmem_base_s, kTmemCols);
        tmem_relinquish();
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem = tmem_base_s;

    pdl_wait();               // Q, K, V^T come from the QKV GEMM
    if (tid == 0) tr.mark(2);
    pdl_launch_dependents();

    if (tid == 0) {  // ---- TMA: Q + K on one barrier, V^T on another (only needed after the softmax of the first chunk)
        mbar_arrive_expect_tx(&bar_qk, P * (kQBytes + kKBytes));
        tma_load_3d(sQ, &mQhi, &bar_qk, 0, q0, bh);
        if (P == 2) tma_load_3d(sQ + kQBytes, &mQlo, &bar_qk, 0, q0, bh);
        for (int c = 0; c < kNChunks; ++c) {
            tma_load_3d(sK + c * kKeyChunk * 128, &mKhi, &bar_qk, 0, c * kKeyChunk, bh);
            if (P == 2) tma_load_3d(sK + kKBytes + c * kKeyChunk * 128, &mKlo, &bar_qk, 0, c * kKeyChunk, bh);
        }
        mbar_arrive_expect_tx(&bar_v, P * kVBytes);
        for (int c = 0; c < kNChunks; ++c) {
            tma_load_3d(sV + c * (kDh * 128), &mVhi, &bar_v, c * kKeyChunk, 0, bh);
            if (P == 2) tma_load_3d(sV + kVBytes + c * (kDh * 128), &mVlo, &bar_v, c * kKeyChunk, 0, bh);
        }
    }
    if (tid == 32) {  // ---- S = Q K^T
        ok &= mbar_wait(&bar_qk, 0);
        tcgen05_fence_after();
        const uint32_t aQ = smem_u32(sQ), aK = smem_u32(sK);
        constexpr uint32_t idesc256 = umma_idesc_bf16(kQTile, 256), idesc64 = umma_idesc_bf16(kQTile, 64);
#pragma unroll
        for (int k = 0; k < kDh / 16; ++k) {
            const uint32_t koff = k * 32;
            const uint64_t qh = umma_desc_sw128(aQ + koff), kh0 = umma_desc_sw128(aK + koff), kh1 = umma_desc_sw128(aK + 256 * 128 + koff);
            umma_bf16(tmem, qh, kh0, idesc256, k != 0);
            umma_bf16(tmem + 256, qh, kh1, idesc64, k != 0);
            if (NSPLIT == 3) {
                const uint64_t ql = umma_desc_sw128(aQ + kQBytes + koff);
                const uint64_t kl0 = umma_desc_sw128(aK + kKBytes + koff), kl1 = umma_desc_sw128(aK + kKBytes + 256 * 128 + koff);
                umma_bf16(tmem, qh, kl0, idesc256, 1);
                umma_bf16(tmem + 256, qh, kl1, idesc64, 1);
                umma_bf16(tmem, ql, kh0, idesc256, 1);
                umma_bf16(tmem + 256, ql, kh1, idesc64, 1);
            }
        }
        umma_commit(&bar_s);
    }
    __syncwarp();

    // ---- softmax: thread (row, g) owns keys 64 c + 16 g .. + 15 of every chunk c
    ok &= mbar_wait(&bar_s, 0);
    tcgen05_fence_after();
    if (tid == 0) tr.mark(4);
    const uint32_t lane_addr = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    const float k2 = 1.4426950408889634f / sqrtf((float)kDh);  // scale * log2(e)
    float mx = -INFINITY;
    if (q_ok) {
#pragma unroll
        for (int c = 0; c < kNChunks; ++c) {
            float v[16];
            tmem_ld_32x16(lane_addr + c * kKeyChunk + g * kKeysPerThread, v);
#pragma unroll
            for (int j = 0; j < 16; ++j) mx = fmaxf(mx, v[j]);
        }
    }
    red[g][row] = mx;
    __syncthreads();
    mx = fmaxf(fmaxf(red[0][row], red[1][row]), fmaxf(red[2][row], red[3][row]));
    if (tid == 0) tr.mark(5);
    const float mk = mx * k2;
    float sum = 0.f;
    for (int c = 0; c < kNChunks; ++c) {
        const int buf = c & 1;
        if (c >= 2) ok &= mbar_wait(&bar_p[buf], ((c >> 1) - 1) & 1);  // the UMMAs that read this buffer are done
        if (q_ok) {
            float v[16];
            tmem_ld_32x16(lane_addr + c * kKeyChunk + g * kKeysPerThread, v);
            uint32_t hi[8], lo[8];
#pragma unroll
            for (int j = 0; j < 16; j += 2) {
                const float e0 = ex2_approx(fmaf(v[j], k2, -mk)), e1 = ex2_approx(fmaf(v[j + 1], k2, -mk));
                sum += e0;
                sum += e1;
                __nv_bfloat16 h0, l0, h1, l1;
                split_bf16(e0, h0, l0), split_bf16(e1, h1, l1);
                hi[j >> 1] = pack_bf16(h0, h1), lo[j >> 1] = pack_bf16(l0, l1);
            }
            uint8_t* pb = sP + buf * (P * kPBytes) + row * 128;
#pragma unroll
            for (int j = 0; j < 2; ++j) {  // 16-byte chunk index within the 128-byte row, XOR-swizzled with the row (Swizzle<3,4,3>)
                const int off = (((2 * g + j) ^ (row & 7)) << 4);
                *reinterpret_cast<uint4*>(pb + off) = make_uint4(hi[4 * j], hi[4 * j + 1], hi[4 * j + 2], hi[4 * j + 3]);
                if (P == 2) *reinterpret_cast<uint4*>(pb + kPBytes + off) = make_uint4(lo[4 * j], lo[4 * j + 1], lo[4 * j + 2], lo[4 * j + 3]);
            }
        }
        fence_proxy_async_smem();  // generic-proxy writes -> visible to the tensor core (async proxy)
        tcgen05_fence_before();
        __syncthreads();
        if (tid == 0 && c == 0) tr.mark(6);
        if (tid == 32) {  // ---- O += P_c V_c
            tcgen05_fence_after();
            if (c == 0) ok &= mbar_wait(&bar_v, 0);
            const uint32_t aP = smem_u32(sP + buf * (P * kPBytes)), aV = smem_u32(sV + c * (kDh * 128));
            constexpr uint32_t idesc = umma_idesc_bf16(kQTile, kDh);
#pragma unroll
            for (int k = 0; k < kKeyChunk / 16; ++k) {
                const uint32_t koff = k * 32;
                const uint64_t ph = umma_desc_sw128(aP + koff), vh = umma_desc_sw128(aV + koff);
                umma_bf16(tmem + kColO, ph, vh, idesc, (c | k) != 0);
                if (NSPLIT == 3) {
                    umma_bf16(tmem + kColO, ph, umma_desc_sw128(aV + kVBytes + koff), idesc, 1);
                    umma_bf16(tmem + kColO, umma_desc_sw128(aP + kPBytes + koff), vh, idesc, 1);
                }
            }
            umma_commit(&bar_p[buf]);
            if (c == kNChunks - 1) umma_commit(&bar_o);
        }
        __syncwarp();
    }
    if (tid == 0) tr.mark(7);
    // the row-max partials were consumed before the first chunk barrier: reuse the array for the row sums
    red[g][row] = sum;
    __syncthreads();
    sum = (red[0][row] + red[1][row]) + (red[2][row] + red[3][row]);

    // ---- epilogue: O / sum -> bf16 split rows [token][h*64 + d], 16 columns per thread
    ok &= mbar_wait(&bar_o, 0);
    tcgen05_fence_after();
    if (q_ok) {
        const float inv = 1.f / sum;
        float v[16];
        tmem_ld_32x16(lane_addr + kColO + g * 16, v);
        uint32_t hi[8], lo[8];
#pragma unroll
        for (int j = 0; j < 16; j += 2) {
            __nv_bfloat16 h0, l0, h1, l1;
            split_bf16(v[j] * inv, h0, l0), split_bf16(v[j + 1] * inv, h1, l1);
            hi[j >> 1] = pack_bf16(h0, h1), lo[j >> 1] = pack_bf16(l0, l1);
        }
        const int64_t idx = ((int64_t)b * kNTok + q0 + row) * D + h * kDh + g * 16;
        uint4* oh = reinterpret_cast<uint4*>(out_hi + idx);
        uint4* ol = reinterpret_cast<uint4*>(out_lo + idx);
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            oh[j] = make_uint4(hi[4 * j], hi[4 * j + 1], hi[4 * j + 2], hi[4 * j + 3]);
            ol[j] = make_uint4(lo[4 * j], lo[4 * j + 1], lo[4 * j + 2], lo[4 * j + 3]);
        }
    }
    if (!ok && err) atomicExch(err, 2);
    tcgen05_fence_before();
    __syncthreads();
    if (tid == 0) tr.mark(3);
    if (warp == 1) tmem_dealloc(tmem, kTmemCols);
}

bool tc_make_map(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes, const uint32_t* box);

bool tc_attention_plan_init(TcAttentionPlan* p, const __nv_bfloat16* Qhi, const __nv_bfloat16* Qlo, const __nv_bfloat16* Khi, const __nv_bfloat16* Klo,
                            const __nv_bfloat16* Vthi, const __nv_bfloat16* Vtlo, int batch_heads) {
    bool ok = true;
    {   // Q, K: [batch*heads][320][64], box {64, rows, 1}
        const uint64_t dims[3] = {kDh, kNTok, (uint64_t)batch_heads}, strides[2] = {kDh * 2, (uint64_t)kDh * 2 * kNTok};
        const uint32_t boxq[3] = {kDh, kQTile, 1}, boxk[3] = {kDh, kKeyChunk, 1};
        ok &= tc_make_map(&p->mQhi, Qhi, 3, dims, strides, boxq) && tc_make_map(&p->mQlo, Qlo, 3, dims, strides, boxq);
        ok &= tc_make_map(&p->mKhi, Khi, 3, dims, strides, boxk) && tc_make_map(&p->mKlo, Klo, 3, dims, strides, boxk);
    }
    {   // V^T: [batch*heads][64][320], box {64 keys, 64 rows, 1}
        const uint64_t dims[3] = {kNTok, kDh, (uint64_t)batch_heads}, strides[2] = {kNTok * 2, (uint64_t)kNTok * 2 * kDh};
        const uint32_t box[3] = {kKeyChunk, kDh, 1};
        ok &= tc_make_map(&p->mVhi, Vthi, 3, dims, strides, box) && tc_make_map(&p->mVlo, Vtlo, 3, dims, strides, box);
    }
    return ok;
}

cudaError_t tc_attention_setup() {
    cudaError_t e = cudaFuncSetAttribute(attention_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, AttSmem<1>::kTotal);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(attention_tc_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, AttSmem<3>::kTotal);
}

cudaError_t tc_attention_launch(const TcAttentionPlan& p, __nv_bfloat16* out_hi, __nv_bfloat16* out_lo, int B, int D, int heads, int nsplit,
                                int* err, cudaStream_t s, bool pdl, unsigned long long* trace) {
    if (B <= 0) return cudaSuccess;
    const dim3 grid((kNTok + kQTile - 1) / kQTile, heads, B);
    if (nsplit == 3)
        return launch_ex(attention_tc_kernel<3>, grid, dim3(kAttThreads), AttSmem<3>::kTotal, s, pdl, 1, p.mQhi, p.mQlo, p.mKhi, p.mKlo, p.mVhi,
                         p.mVlo, out_hi, out_lo, D, heads, err, trace);
    return launch_ex(attention_tc_kernel<1>, grid, dim3(kAttThreads), AttSmem<1>::kTotal, s, pdl, 1, p.mQhi, p.mQlo, p.mKhi, p.mKlo, p.mVhi, p.mVlo,
                     out_hi, out_lo, D, heads, err, trace);
}

}  // namespace vt
