// tracker.cu — vt_tracker: device memory, per-handle CUDA stream, CUDA-graph replay of the per-frame
// kernel chain, and the VitTrack / convert / overlay / timing entry points of include/vt_tracker.h.
//
// Per frame (update/submit):   H2D frame (pinned, cudaMemcpyAsync on the handle's stream)
//   -> K2 crop+convert+resize+normalise (search window of every active target, rect_last read on device)
//   -> template-token gather -> patch-embed GEMM -> depth x [LN+QKV, attention, proj+res, LN+FC1+GELU, FC2+res]
//   -> final LN -> 3x3 head conv (im2col GEMM) -> K8 decode (updates rect_last on device)
//   -> optional K9 box overlay -> results published into the pinned host block by the last kernel (+ touched rows for pageable frames).
// rect_last never leaves the device between frames, so consecutive frames can be enqueued without a
// host round trip.
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <sys/stat.h>

#include <algorithm>
#include <chrono>
#include <atomic>
#include <map>
#include <memory>
#include <mutex>

#include "tc_common.cuh"
#include "vt_internal.h"

namespace vt {

static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

struct BlockW {
    const float *ln1_g, *ln1_b, *qkv_w, *qkv_b, *proj_w, *proj_b, *ln2_g, *ln2_b, *fc1_w, *fc1_b, *fc2_w, *fc2_b;
};

// One device copy of a weight file (fp32 master + bf16 hi / lo split) shared by every handle of that device: with many streams per GPU
// the 45 MB stay L2 resident once instead of once per stream.  Keyed by device, path, size and mtime; freed with the last handle.
struct WeightSet {
    int device = 0;
    float* d_weights = nullptr;
    __nv_bfloat16 *w_hi = nullptr, *w_lo = nullptr;
    __nv_bfloat16* w_f16 = nullptr;  // fp16 copy for handles in VT_GEMM_TCGEN05_FP16 mode (made on first use)
    size_t n = 0;
    int32_t hdr[7] = {0, 0, 0, 0, 0, 0, 0};
    std::mutex split_mutex;
    ~WeightSet() {
        cudaSetDevice(device);
        if (d_weights) cudaFree(d_weights);
        if (w_hi) cudaFree(w_hi);
        if (w_lo) cudaFree(w_lo);
        if (w_f16) cudaFree(w_f16);
    }
};
static std::mutex g_weight_mutex;
// Live tracker handles per device.  With one or two streams on a GPU most SMs idle during a frame, and the "spread" GEMM forms trade
// them for latency (tiles replicated so that each replica stores a share of the epilogue output); with more streams SM time is the
// budget and the plain forms are used.  Read at every frame: the graph variant follows the handle count.
constexpr int kMaxDevices = 64, kSpreadMaxHandles = 2, kUnchainTargets = 8;
static std::atomic<int> g_live_handles[kMaxDevices];
static std::map<std::string, std::weak_ptr<WeightSet>> g_weight_cache;

enum { EV_START = 0, EV_H2D, EV_PRE, EV_VIT, EV_DEC, EV_OVL, EV_END, EV_COUNT };

}  // namespace vt

using namespace vt;

struct vt_tracker {
    vt_config cfg;
    int D = 0, depth = 0, heads = 0, hidden = 0, head_ch = 0;
    int W = 0, H = 0, fmt = 0, maxT = 1;
    size_t frame_bytes = 0;
    float threshold = 0.2f;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[EV_COUNT] = {};
    bool ev_valid = false;

    // weights (one device allocation, shared between the handles of a device: WeightSet)
    std::shared_ptr<WeightSet> wset;
    float* d_weights = nullptr;
    size_t n_weights = 0;
    const float *patch_w, *patch_b, *pos_z, *pos_x, *lnf_g, *lnf_b, *h1_w, *h1_b, *h2_w, *h2_b;
    std::vector<BlockW> blk;
    float *d_lut = nullptr, *d_hann = nullptr;

    // frame + state
    uint8_t* d_frame = nullptr;
    uint8_t* d_frames[2] = {nullptr, nullptr};  // double buffer: d_frame points at the one most recently filled (queue slot s uses [s])
    cudaStream_t copy_stream = nullptr;          // uploads of pipelined host frames overlap the in-flight frame's kernels
    cudaEvent_t ev_up[2] = {nullptr, nullptr};
    uint8_t* d_rgb = nullptr;  // lazily allocated, vt_convert_nv12_rgb
    uint8_t *d_fmt_in = nullptr, *d_fmt_out = nullptr;  // lazily grown scratch of the format entry points (YUY2, resize)
    size_t fmt_in_cap = 0, fmt_out_cap = 0;
    uint8_t* h_stage = nullptr;  // pinned staging for non-pinned callers
    size_t h_stage_bytes = 0;
    int frame_valid = 1;
    TargetState* d_state = nullptr;
    int32_t* d_slots = nullptr;
    // one device block [DeviceResult x maxT][u64 stamps x ST_COUNT][int tc_err, pad] written into the pinned host block of the frame's queue slot by the frame's last kernel
    DeviceResult *d_res = nullptr, *h_res = nullptr;
    size_t res_block_bytes = 0;
    unsigned long long *d_stamps = nullptr, *h_stamps = nullptr;
    int* h_tc_err = nullptr;
    const uint8_t** d_frame_slot = nullptr;  // device cell: address of the frame the step reads (d_frame, or the caller's device frame)
    uint8_t** d_host_slot = nullptr;   // device cell: address of the caller's pinned host frame for the zero-copy overlay mirror (or null)
    uint32_t** d_hblk_slot = nullptr;  // device cell: address of the pinned host result block of the frame's queue slot
    bool inflight_mirrored = false;
    float* d_maps = nullptr;
    OverlayCmdDev *d_cmds = nullptr, *h_cmds = nullptr;
    std::vector<int> active;          // slot indices, ascending
    std::vector<vt_bbox> rect_mirror; // host mirror of rect_last (valid after wait)
    std::vector<int> inited;

    // activations
    float *patches_x = nullptr, *patches_z = nullptr, *Zemb = nullptr, *X = nullptr, *QKV = nullptr, *ATT = nullptr, *HID = nullptr,
          *Yf = nullptr, *H1 = nullptr, *d_dbg = nullptr;
    int debug_capture = 0;

    // tensor-core mode (gemm_mode != VT_GEMM_FP32_SIMT): bf16 (hi, lo) copies of the weights and of every GEMM A operand
    int nsplit = 0;  // 0 = fp32 SIMT, 1 = bf16, 2 = fp16 (single pass), 3 = bf16x3
    bool f16 = false;  // nsplit == 2: the operand "hi" buffers hold fp16 values, the "lo" buffers are unused (kernels get null)
    __nv_bfloat16 *w_hi = nullptr, *w_lo = nullptr;
    __nv_bfloat16 *px_hi = nullptr, *px_lo = nullptr, *pz_hi = nullptr, *pz_lo = nullptr, *ln_hi = nullptr, *ln_lo = nullptr, *att_hi = nullptr,
                  *att_lo = nullptr, *hid_hi = nullptr, *hid_lo = nullptr, *yf_hi = nullptr, *yf_lo = nullptr;
    __nv_bfloat16 *q_hi = nullptr, *q_lo = nullptr, *k_hi = nullptr, *k_lo = nullptr, *vt_hi = nullptr, *vt_lo = nullptr;
    __nv_bfloat16 *zln_hi = nullptr, *zln_lo = nullptr;  // LN1 (block 0) of the template tokens, [B][64][D], computed at init
    bool fuse_ln = false;       // LayerNorm fused into the producing GEMM's epilogue (cluster of D / 64 CTAs)
    bool chain_mlp = false;     // FC2 partial products computed inside the FC1 kernel + reduce_ln_kernel (no hidden round trip)
    float* Pbuf = nullptr;      // [hidden / 64][B][320][D] fp32 partial FC2 results (also the 4 split-K partials of the patch embed)
    bool split_k = false;       // patch embed and 3x3 head conv as split-K partial GEMMs + reduce kernels
    float *Phead = nullptr, *d_cand = nullptr;  // [9 taps][B][256][head_ch] conv partials; [B][16][8] row candidates of the decode
    unsigned* d_counters = nullptr;
    bool pdl = true;            // programmatic dependent launch along the kernel chain
    bool spread_ok = true;      // latency-mode GEMM forms allowed (VT_B200_NO_SPREAD disables)
    int unchain_n = kUnchainTargets;  // active targets from which the MLP runs unchained (VT_B200_UNCHAIN_N overrides)
    bool counted = false;       // this handle is included in g_live_handles
    bool tc_attention = false;  // head_dim == 64
    TcAttentionPlan plan_att;
    int* d_tc_err = nullptr;
    unsigned long long* d_trace = nullptr;  // VT_B200_TRACE=1: device timeline of the chain (vt_tracker_debug_trace)
    TcGemmPlan plan_patch_x, plan_patch_z, plan_head;
    struct BlockPlans {
        TcGemmPlan qkv, proj, fc1, fc2;
        TcAttentionPlan att;  // plan_att + this block's W_proj maps (chained form)
    };
    bool att_chain_ok = false;  // the chained attention form is available (VT_B200_NO_ATT_CHAIN disables)
    std::vector<BlockPlans> plans;

    std::map<int, cudaGraphExec_t> graphs;
    int kernels_per_frame = 0;
    uint64_t kernel_launches = 0, frames = 0, h2d_bytes = 0, d2h_bytes = 0;

    // in-flight frames: a queue of depth kQueue.  rect_last lives on the device, so frame t+1 can be enqueued before the results of
    // frame t have been read back; every slot has its own pinned result block and completion event.
    static constexpr int kQueue = 2;
    struct Slot {
        uint8_t* frame = nullptr;  // caller's host frame (null for device-resident frames)
        size_t len = 0;
        bool mirrored = false, pageable = false;
        std::chrono::steady_clock::time_point t_submit;
    } q[kQueue];
    int q_head = 0, q_count = 0;
    DeviceResult* h_blk[kQueue] = {nullptr, nullptr};
    cudaEvent_t q_done[kQueue] = {nullptr, nullptr};
    bool in_flight = false;          // q_count > 0
    uint8_t* inflight_frame = nullptr;
    std::chrono::steady_clock::time_point t_submit;

    // VT_B200_HOSTPROF=1: host-side wall time of the submit / wait phases, printed at destroy (diagnostics)
    bool hostprof = false;
    double hp[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    uint64_t hp_n = 0;

    TimingStats stats;
    Ring<float> r_h2d, r_pre, r_vit, r_dec, r_ovl, r_d2h, r_tot;
    float last[7] = {0, 0, 0, 0, 0, 0, 0};
};

namespace vt {

static bool is_pinned(const void* p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return a.type == cudaMemoryTypeHost;
}

// Header of a VTW1 file: magic, shape, and the float count the shape implies — checked against the file size BEFORE anything is
// allocated (the file comes from outside: negative or absurd header fields must not turn into a huge allocation or an overflow).
static vt_status read_weight_header(const char* path, FILE** f_out, int32_t hdr[7], size_t* n_out, struct stat* sb_out) {
    struct stat sb;
    if (!path || stat(path, &sb) != 0 || !S_ISREG(sb.st_mode)) {
        set_error("cannot open weight file %s", path ? path : "(null)");
        return VT_ERR_WEIGHTS;
    }
    FILE* f = fopen(path, "rb");
    if (!f) {
        set_error("cannot open weight file %s", path);
        return VT_ERR_WEIGHTS;
    }
    char magic[4];
    if (fread(magic, 1, 4, f) != 4 || memcmp(magic, "VTW1", 4) != 0 || fread(hdr, 4, 7, f) != 7) {
        fclose(f);
        set_error("%s is not a VTW1 weight file", path);
        return VT_ERR_WEIGHTS;
    }
    const long long Dl = hdr[0], L = hdr[1], heads = hdr[2], Hl = hdr[3], Cl = hdr[4];
    if (Dl <= 0 || Dl > 1024 || Dl % 32 || L <= 0 || L > 64 || Hl <= 0 || Hl > 8192 || Hl % 32 || Cl <= 0 || Cl > 1024 || Cl % 32 || heads <= 0 ||
        Dl % heads || (Dl / heads != 16 && Dl / heads != 32 && Dl / heads != 64)) {
        fclose(f);
        set_error("unsupported model shape D=%d depth=%d heads=%d hidden=%d head_ch=%d", hdr[0], hdr[1], hdr[2], hdr[3], hdr[4]);
        return VT_ERR_WEIGHTS;
    }
    const size_t D = (size_t)Dl, H = (size_t)Hl, C = (size_t)Cl;
    const size_t n = D * kPatchK + D + kNTz * D + kNTx * D + (size_t)L * (4 * D + 3 * D * D + 3 * D + D * D + D + H * D + H + D * H + D) + 2 * D +
                     C * D * 9 + C + 5 * C + 5;
    if ((unsigned long long)sb.st_size != 32ull + 4ull * n) {
        fclose(f);
        set_error("weight file %s: %lld bytes, the header implies %llu", path, (long long)sb.st_size, 32ull + 4ull * n);
        return VT_ERR_WEIGHTS;
    }
    *n_out = n;
    if (sb_out) *sb_out = sb;
    if (f_out) *f_out = f;
    else fclose(f);
    return VT_OK;
}

static vt_status load_weights(vt_tracker* t, const char* path) {
    struct stat sb;
    if (!path || stat(path, &sb) != 0) {
        set_error("cannot open weight file %s", path ? path : "(null)");
        return VT_ERR_WEIGHTS;
    }
    char key[1200];
    snprintf(key, sizeof(key), "%d|%s|%lld|%lld", t->cfg.device, path, (long long)sb.st_size, (long long)sb.st_mtime);
    std::lock_guard<std::mutex> lock(g_weight_mutex);
    std::shared_ptr<WeightSet> ws = g_weight_cache[key].lock();
    if (!ws) {
        FILE* f = nullptr;
        int32_t hdr[7];
        size_t n = 0;
        vt_status hs = read_weight_header(path, &f, hdr, &n, nullptr);
        if (hs != VT_OK) return hs;
        const size_t D = hdr[0], C = hdr[4];
        std::vector<float> host(n);
        const size_t got = fread(host.data(), sizeof(float), n, f);
        fclose(f);
        if (got != n) {
            set_error("weight file %s is truncated (%zu of %zu floats)", path, got, n);
            return VT_ERR_WEIGHTS;
        }
        // head conv weight [C, D, 3, 3] -> [C, tap, D] so that the im2col K axis is tap-major
        const size_t h1_off = n - (5 + 5 * C + C + C * D * 9);
        {
            std::vector<float> re(C * D * 9);
            for (size_t c = 0; c < C; ++c)
                for (size_t d = 0; d < D; ++d)
                    for (size_t tap = 0; tap < 9; ++tap) re[(c * 9 + tap) * D + d] = host[h1_off + (c * D + d) * 9 + tap];
            std::copy(re.begin(), re.end(), host.begin() + h1_off);
        }
        ws = std::make_shared<WeightSet>();
        ws->device = t->cfg.device, ws->n = n;
        memcpy(ws->hdr, hdr, sizeof(hdr));
        VT_CUDA(cudaMalloc(&ws->d_weights, n * sizeof(float)));
        VT_CUDA(cudaMemcpy(ws->d_weights, host.data(), n * sizeof(float), cudaMemcpyHostToDevice));
        g_weight_cache[key] = ws;
    }
    t->wset = ws;
    t->D = ws->hdr[0], t->depth = ws->hdr[1], t->heads = ws->hdr[2], t->hidden = ws->hdr[3], t->head_ch = ws->hdr[4];
    const size_t D = t->D, H = t->hidden, C = t->head_ch;
    t->n_weights = ws->n;
    t->d_weights = ws->d_weights;
    const float* p = t->d_weights;
    auto take = [&](size_t cnt) {
        const float* r = p;
        p += cnt;
        return r;
    };
    t->patch_w = take(D * kPatchK), t->patch_b = take(D), t->pos_z = take(kNTz * D), t->pos_x = take(kNTx * D);
    t->blk.resize(t->depth);
    for (auto& b : t->blk) {
        b.ln1_g = take(D), b.ln1_b = take(D), b.qkv_w = take(3 * D * D), b.qkv_b = take(3 * D);
        b.proj_w = take(D * D), b.proj_b = take(D), b.ln2_g = take(D), b.ln2_b = take(D);
        b.fc1_w = take(H * D), b.fc1_b = take(H), b.fc2_w = take(D * H), b.fc2_b = take(D);
    }
    t->lnf_g = take(D), t->lnf_b = take(D), t->h1_w = take(C * D * 9), t->h1_b = take(C), t->h2_w = take(5 * C), t->h2_b = take(5);
    return VT_OK;
}

// template tokens (fixed since init) -> rows 0..63 of every target's sequence: residual stream X and, on the fused-LN
// tensor-core path, their block-0 LN1 as the bf16 split A operand of the first QKV GEMM.
// Independent of the crop kernel ahead of it: launched as its programmatic dependent, it does its copies WHILE the crop runs and only
// then waits for it, so that the patch GEMM behind (whose dependency wait covers this kernel only) still starts after the crop.
__global__ void gather_template_kernel(float* __restrict__ X, const float* __restrict__ Zemb, const int32_t* __restrict__ slots, int D,
                                       uint32_t* __restrict__ ln_hi, uint32_t* __restrict__ ln_lo, const uint32_t* __restrict__ zln_hi,
                                       const uint32_t* __restrict__ zln_lo, unsigned long long* stamp) {
    tc::pdl_launch_dependents();
    const int bi = blockIdx.y;
    const int n = kNTz * D;
    const size_t so = (size_t)slots[bi] * n, xo = (size_t)bi * kNTok * D;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        X[xo + i] = Zemb[so + i];
        if (ln_hi && i < n / 2) {
            ln_hi[xo / 2 + i] = zln_hi[so / 2 + i];
            if (ln_lo) ln_lo[xo / 2 + i] = zln_lo[so / 2 + i];
        }
    }
    tc::pdl_wait();  // the crop kernel has completed: end of the preprocess stage
    if (stamp && threadIdx.x == 0 && blockIdx.x == 0 && blockIdx.y == 0) *stamp = device_time_ns();
}

static GemmArgs gemm_args(const float* A, int64_t lda, const float* W, const float* bias, float* C, int64_t ldc, int M, int N, int K) {
    GemmArgs g;
    memset(&g, 0, sizeof(g));
    g.A = A, g.lda = lda, g.W = W, g.bias = bias, g.C = C, g.ldc = ldc, g.M = M, g.N = N, g.K = K;
    g.a_rows_in = g.c_rows_in = 1 << 30;
    return g;
}

#define VT_LAUNCH(call)                                                                   \
    do {                                                                                  \
        cudaError_t e_ = (call);                                                          \
        if (e_ != cudaSuccess) {                                                          \
            set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
            return VT_ERR_CUDA;                                                           \
        }                                                                                 \
        ++launches;                                                                       \
    } while (0)

// Enqueues crop -> ViT -> decode (-> box overlay) for the n active targets on t->stream.
static vt_status enqueue_forward(vt_tracker* t, int n, int& launches, bool record_events, bool capturing, bool spread) {
    const int D = t->D, Hd = t->hidden, C = t->head_ch;
    (void)record_events, (void)capturing;  // stage times come from device stamps (ST_*), not from event nodes
    cudaStream_t s = t->stream;
    FrameDesc fd{t->d_frame, t->W, t->H, t->fmt, t->frame_valid, t->d_frame_slot};
    auto LO = [&](__nv_bfloat16* p) { return t->f16 ? nullptr : p; };  // fp16 mode: "hi only" (see vt_internal.h: operand_bits)
    VT_LAUNCH(launch_crop_resize_norm(fd, t->d_state, t->d_slots, n, 4, kSearch, t->d_lut, t->patches_x, (size_t)kNTx * kPatchK, t->px_hi,
                                      LO(t->px_lo), s, t->d_stamps + ST_PRE));
    {
        dim3 grid((kNTz * D + 255) / 256, n);
        const bool f = t->fuse_ln;
        VT_LAUNCH(launch_ex(gather_template_kernel, grid, dim3(256), 0, s, t->pdl && !t->debug_capture, 1, t->X, (const float*)t->Zemb,
                            (const int32_t*)t->d_slots, D, f ? (uint32_t*)t->ln_hi : nullptr, (uint32_t*)LO(t->ln_lo), (const uint32_t*)t->zln_hi,
                            (const uint32_t*)t->zln_lo, t->d_stamps + ST_VIT));
    }
    const int M = n * kNTok;
    if (t->nsplit == 0) {
        // ---------------- fp32 CUDA-core path ----------------
        {
            GemmArgs g = gemm_args(t->patches_x, kPatchK, t->patch_w, t->patch_b, t->X, D, n * kNTx, D, kPatchK);
            g.pos = t->pos_x;
            g.c_rows_in = kNTx, g.c_rows_stride = kNTok, g.c_row_off = kNTz;
            VT_LAUNCH(launch_gemm_simt(g, s));
        }
        if (t->debug_capture) VT_CUDA(cudaMemcpyAsync(t->d_dbg, t->X, sizeof(float) * M * D, cudaMemcpyDeviceToDevice, s));
        for (int l = 0; l < t->depth; ++l) {
            const BlockW& b = t->blk[l];
            {
                GemmArgs g = gemm_args(t->X, D, b.qkv_w, b.qkv_b, t->QKV, 3 * D, M, 3 * D, D);
                g.ln_g = b.ln1_g, g.ln_b = b.ln1_b;
                VT_LAUNCH(launch_gemm_simt(g, s));
            }
            VT_LAUNCH(launch_attention(t->QKV, t->ATT, nullptr, nullptr, n, D, t->heads, s));
            {
                GemmArgs g = gemm_args(t->ATT, D, b.proj_w, b.proj_b, t->X, D, M, D, D);
                g.residual = 1;
                VT_LAUNCH(launch_gemm_simt(g, s));
            }
            {
                GemmArgs g = gemm_args(t->X, D, b.fc1_w, b.fc1_b, t->HID, Hd, M, Hd, D);
                g.ln_g = b.ln2_g, g.ln_b = b.ln2_b, g.gelu = 1;
                VT_LAUNCH(launch_gemm_simt(g, s));
            }
            {
                GemmArgs g = gemm_args(t->HID, Hd, b.fc2_w, b.fc2_b, t->X, D, M, D, Hd);
                g.residual = 1;
                VT_LAUNCH(launch_gemm_simt(g, s));
            }
            if (t->debug_capture)
                VT_CUDA(cudaMemcpyAsync(t->d_dbg + (size_t)(l + 1) * t->maxT * kNTok * D, t->X, sizeof(float) * M * D, cudaMemcpyDeviceToDevice, s));
        }
        VT_LAUNCH(launch_layernorm(t->X, D, t->lnf_g, t->lnf_b, t->Yf, D, n * kNTx, D, kNTx, kNTok, kNTz, s));
        {
            GemmArgs g = gemm_args(t->Yf, D, t->h1_w, t->h1_b, t->H1, C, n * kNTx, C, 9 * D);
            g.relu = 1, g.im2col_feat = D;
            g.a_rows_in = kNTx, g.a_rows_stride = kNTx, g.a_row_off = 0;
            VT_LAUNCH(launch_gemm_simt(g, s));
        }
    } else {
        // ---------------- tensor-core path: tcgen05 GEMMs fed by TMA, bf16 (x3 split) operands, fp32 TMEM accumulators ----------------
        const int ns = t->nsplit;
        const bool pdl = t->pdl && !t->debug_capture, fuse = t->fuse_ln;
        VT_LAUNCH(tc_gemm_launch(t->plan_patch_x, n * kNTx, ns, s, pdl));  // fused: + LN1 of block 0 for the search rows
        if (t->split_k) {  // X[64.., :] = pos_x + patch_b + sum of the 4 K-slices; LN1 of block 0
            ReduceLnArgs r{};
            r.P = t->Pbuf, r.np = 4, r.p_stride = (int64_t)t->maxT * kNTx * D, r.bias = t->patch_b, r.add = t->pos_x, r.add_period = kNTx;
            r.X = t->X, r.M = n * kNTx, r.D = D, r.period = kNTx, r.x_rows = kNTok, r.x_row_off = kNTz;
            r.ln_g = t->blk[0].ln1_g, r.ln_b = t->blk[0].ln1_b, r.ln_hi = t->ln_hi, r.ln_lo = LO(t->ln_lo), r.ln_rows = kNTok, r.ln_row_off = kNTz;
            VT_LAUNCH(launch_reduce_ln(r, s, pdl));
        }
        if (t->debug_capture) VT_CUDA(cudaMemcpyAsync(t->d_dbg, t->X, sizeof(float) * M * D, cudaMemcpyDeviceToDevice, s));
        for (int l = 0; l < t->depth; ++l) {
            const BlockW& b = t->blk[l];
            const vt_tracker::BlockPlans& p = t->plans[l];
            if (!fuse) VT_LAUNCH(launch_layernorm_split(t->X, D, b.ln1_g, b.ln1_b, t->ln_hi, LO(t->ln_lo), M, D, 1 << 30, 0, 0, s, pdl));
            VT_LAUNCH(tc_gemm_launch(p.qkv, M, ns, s, pdl, spread));
            // Latency mode: the proj GEMM is folded into the attention kernel (per-head partial products, D / 64 replicas per tile) and
            // reduce_ln adds the heads + bias + residual and applies LN2 — one kernel and one dependency edge less per block.
            const bool att_chain = spread && t->att_chain_ok && n * t->heads * 3 * (D / kAttChainW) <= kSpreadCtas;
            if (t->tc_attention)
                VT_LAUNCH(tc_attention_launch(att_chain ? p.att : t->plan_att, n, t->heads, ns, t->d_tc_err, s, pdl, t->d_trace,
                                              att_chain ? VT_ATT_CHAIN : (spread ? VT_ATT_DUP : VT_ATT_PLAIN)));
            else
                VT_LAUNCH(launch_attention(t->QKV, nullptr, t->att_hi, LO(t->att_lo), n, D, t->heads, s));
            if (att_chain) {
                ReduceLnArgs r{};
                r.P = t->Pbuf, r.np = t->heads, r.p_stride = (int64_t)t->maxT * kNTok * D, r.bias = b.proj_b, r.add = t->X, r.add_period = 0;
                r.X = t->X, r.M = M, r.D = D, r.period = kNTok, r.x_rows = kNTok, r.x_row_off = 0;
                r.ln_g = b.ln2_g, r.ln_b = b.ln2_b, r.ln_hi = t->ln_hi, r.ln_lo = LO(t->ln_lo), r.ln_rows = kNTok, r.ln_row_off = 0;
                VT_LAUNCH(launch_reduce_ln(r, s, pdl));
            } else {
                VT_LAUNCH(tc_gemm_launch(p.proj, M, ns, s, pdl && t->tc_attention, spread));  // fused: + LN2
            }
            if (!fuse) VT_LAUNCH(launch_layernorm_split(t->X, D, b.ln2_g, b.ln2_b, t->ln_hi, LO(t->ln_lo), M, D, 1 << 30, 0, 0, s, pdl));
            // Chained form (FC2 partial products inside the FC1 kernel, summed by reduce_ln): shortest critical path for a few targets.
            // From kUnchainTargets targets on the 12 fp32 partial planes per row tile cost more than the hidden round trip
            // (cfg4, 16 targets: ViT stage 905 -> 819 us unchained), so FC1 writes the hidden tile and FC2 runs as its own GEMM.
            const bool chain = t->chain_mlp && n < t->unchain_n;
            if (chain) {
                VT_LAUNCH(tc_gemm_launch(p.fc1, M, ns, s, pdl, spread));
            } else {
                TcGemmPlan fc1 = p.fc1;
                fc1.args.chain_n = 0, fc1.args.o_mode = 1;
                VT_LAUNCH(tc_gemm_launch(fc1, M, ns, s, pdl));
            }
            if (chain) {  // X += fc2_b + sum of the partials; LN1 of the next block / the final LN of the search rows
                const bool last = l + 1 == t->depth;
                ReduceLnArgs r{};
                r.P = t->Pbuf, r.np = Hd / 64, r.p_stride = (int64_t)t->maxT * kNTok * D, r.bias = b.fc2_b, r.add = t->X, r.add_period = 0;
                r.X = t->X, r.M = M, r.D = D, r.period = kNTok, r.x_rows = kNTok, r.x_row_off = 0;
                r.ln_g = last ? t->lnf_g : t->blk[l + 1].ln1_g, r.ln_b = last ? t->lnf_b : t->blk[l + 1].ln1_b;
                r.ln_hi = last ? t->yf_hi : t->ln_hi, r.ln_lo = LO(last ? t->yf_lo : t->ln_lo);
                r.ln_rows = last ? kNTx : kNTok, r.ln_row_off = last ? -kNTz : 0;
                VT_LAUNCH(launch_reduce_ln(r, s, pdl));
            } else {
                VT_LAUNCH(tc_gemm_launch(p.fc2, M, ns, s, pdl));  // fused: + LN1 of the next block / the final LN of the search rows
            }
            if (t->debug_capture)
                VT_CUDA(cudaMemcpyAsync(t->d_dbg + (size_t)(l + 1) * t->maxT * kNTok * D, t->X, sizeof(float) * M * D, cudaMemcpyDeviceToDevice, s));
        }
        if (!fuse) VT_LAUNCH(launch_layernorm_split(t->X, D, t->lnf_g, t->lnf_b, t->yf_hi, LO(t->yf_lo), n * kNTx, D, kNTx, kNTok, kNTz, s, pdl));
        VT_LAUNCH(tc_gemm_launch(t->plan_head, n * kNTx, ns, s, pdl));
    }
    if (t->nsplit && t->split_k)
        VT_LAUNCH(launch_head_decode(t->Phead, 9, (int64_t)t->maxT * kNTx * C, C, t->h1_b, t->h2_w, t->h2_b, t->d_hann, t->d_state, t->d_slots, n,
                                     t->threshold, t->d_res, t->d_maps, t->d_cand, t->d_counters, t->d_stamps, s, t->pdl && !t->debug_capture));
    else
        VT_LAUNCH(launch_decode(t->H1, C, t->h2_w, t->h2_b, t->d_hann, t->d_state, t->d_slots, n, t->threshold, t->d_res, t->d_maps, t->d_stamps, s));
    if (t->cfg.box_overlay)
        VT_LAUNCH(launch_box_overlay(t->d_frame, t->frame_bytes, t->W, t->H, overlay_format(t->fmt), t->d_res, t->d_slots, n, t->cfg.overlay_gate,
                                     t->d_host_slot, t->d_stamps + ST_OVL_END, s, (uint8_t* const*)t->d_frame_slot, t->pdl && !t->debug_capture,
                                     t->d_res, t->d_hblk_slot, t->res_block_bytes));  // ... and publishes the result block
    else  // results, stage stamps and the error flag -> the pinned host block of this frame's queue slot
        VT_LAUNCH(launch_publish(t->d_res, t->d_hblk_slot, t->res_block_bytes, s, t->pdl && !t->debug_capture));
    return VT_OK;
}

static vt_status run_forward(vt_tracker* t) {
    const int n = (int)t->active.size();
    if (n == 0) return VT_OK;
    int launches = 0;
    const int dev = t->cfg.device;
    const bool spread = t->spread_ok && dev >= 0 && dev < kMaxDevices && g_live_handles[dev].load(std::memory_order_relaxed) <= kSpreadMaxHandles;
    if (!t->cfg.use_cuda_graph || t->debug_capture) {
        vt_status st = enqueue_forward(t, n, launches, true, false, spread);
        t->kernel_launches += launches;
        t->kernels_per_frame = launches;
        return st;
    }
    const int key = (n * 2 + (t->frame_valid ? 1 : 0)) * 2 + (spread ? 1 : 0);  // frame_valid is a kernel parameter baked into the graph
    auto it = t->graphs.find(key);
    if (it == t->graphs.end()) {
        cudaGraph_t graph = nullptr;
        VT_CUDA(cudaStreamBeginCapture(t->stream, cudaStreamCaptureModeThreadLocal));
        vt_status st = enqueue_forward(t, n, launches, true, true, spread);
        cudaError_t e = cudaStreamEndCapture(t->stream, &graph);
        if (st != VT_OK) {
            if (graph) cudaGraphDestroy(graph);
            return st;
        }
        if (e != cudaSuccess) {
            set_error("cudaStreamEndCapture failed: %s", cudaGetErrorString(e));
            return VT_ERR_CUDA;
        }
        cudaGraphExec_t exec = nullptr;
        VT_CUDA(cudaGraphInstantiate(&exec, graph, 0));
        cudaGraphDestroy(graph);
        it = t->graphs.emplace(key, exec).first;
        t->kernels_per_frame = launches;
    }
    VT_CUDA(cudaGraphLaunch(it->second, t->stream));
    t->kernel_launches += t->kernels_per_frame;
    return VT_OK;
}

static vt_status sync_slots(vt_tracker* t) {
    if (!t->active.empty())
        VT_CUDA(cudaMemcpyAsync(t->d_slots, t->active.data(), sizeof(int32_t) * t->active.size(), cudaMemcpyHostToDevice, t->stream));
    VT_CUDA(cudaStreamSynchronize(t->stream));  // t->active is pageable: make the copy complete before it can change
    return VT_OK;
}

// Search window of a target in frame coordinates (App. A.1 with factor 4), clipped to the frame and grown to even coordinates
// (NV12 chroma pairs).  Returns false when the window misses the frame.
static bool search_window(const vt_tracker* t, const vt_bbox& r, int& x0, int& y0, int& x1, int& y1) {
    if (r.width <= 0 || r.height <= 0) return false;
    const int c = (int)ceil(sqrt((double)(int)((long long)r.width * r.height)) * 4.0);
    const int wx = r.x + (r.width - c) / 2, wy = r.y + (r.height - c) / 2;
    x0 = std::max(wx, 0) & ~1, y0 = std::max(wy, 0) & ~1;
    x1 = std::min((std::min(wx + c, t->W) + 1) & ~1, t->W), y1 = std::min((std::min(wy + c, t->H) + 1) & ~1, t->H);
    return x1 > x0 && y1 > y0;
}

// host -> device frame upload (pinned: direct async; pageable: staged through the handle's pinned buffer).
// cfg.upload_window: only the search windows of the active targets travel (PCIe is the end-to-end roofline):
// the fused crop kernel reads nothing else.  rect_mirror is exact whenever no frame is in flight.
static vt_status upload_frame(vt_tracker* t, const uint8_t* frame, size_t len, bool allow_window = false, bool device_src = false,
                              cudaStream_t stream = nullptr) {
    if (!stream) stream = t->stream;
    size_t n = std::min(len, t->frame_bytes);
    t->frame_valid = (t->fmt == VT_FMT_NV12) ? (len >= (size_t)t->W * t->H * 3 / 2) : (len >= t->frame_bytes);
    if (!t->frame_valid && t->fmt == VT_FMT_NV12) n = 0;  // src/nv12_convert.rs:48-50 -> black image
    if (n == 0) return VT_OK;
    const bool pinned = device_src || is_pinned(frame);
    const cudaMemcpyKind kind = device_src ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
    if (allow_window && t->cfg.upload_window && pinned && t->frame_valid && len >= t->frame_bytes && !t->active.empty() && (t->W % 2 == 0) &&
        (t->H % 2 == 0 || t->fmt != VT_FMT_NV12)) {
        struct Win { int x0, y0, x1, y1; };
        std::vector<Win> wins;
        size_t bytes = 0;
        const size_t bpp_num = t->fmt == VT_FMT_NV12 ? 3 : (t->fmt == VT_FMT_GRAY8 ? 2 : 6);  // bytes per pixel x 2
        for (int s : t->active) {
            Win w;
            if (!search_window(t, t->rect_mirror[s], w.x0, w.y0, w.x1, w.y1)) continue;  // the crop kernel flags it; nothing to read
            wins.push_back(w);
            bytes += (size_t)(w.x1 - w.x0) * (w.y1 - w.y0) * bpp_num / 2;
        }
        if (bytes * 2 <= t->frame_bytes) {
            for (const Win& w : wins) {
                const size_t cols = (size_t)(w.x1 - w.x0), rows = (size_t)(w.y1 - w.y0);
                if (t->fmt == VT_FMT_GRAY8) {
                    const size_t W = (size_t)t->W, o = (size_t)w.y0 * W + w.x0;
                    VT_CUDA(cudaMemcpy2DAsync(t->d_frame + o, W, frame + o, W, cols, rows, kind, stream));
                } else if (t->fmt == VT_FMT_NV12) {
                    const size_t W = (size_t)t->W, yo = (size_t)w.y0 * W + w.x0, uvo = W * t->H + (size_t)(w.y0 / 2) * W + w.x0;
                    VT_CUDA(cudaMemcpy2DAsync(t->d_frame + yo, W, frame + yo, W, cols, rows, kind, stream));
                    VT_CUDA(cudaMemcpy2DAsync(t->d_frame + uvo, W, frame + uvo, W, cols, rows / 2, kind, stream));
                } else {
                    const size_t pitch = (size_t)t->W * 3, o = (size_t)w.y0 * pitch + (size_t)w.x0 * 3;
                    VT_CUDA(cudaMemcpy2DAsync(t->d_frame + o, pitch, frame + o, pitch, cols * 3, rows, kind, stream));
                }
            }
            if (!device_src) t->h2d_bytes += bytes;
            return VT_OK;
        }
    }
    if (!device_src) t->h2d_bytes += n;
    if (pinned) {
        VT_CUDA(cudaMemcpyAsync(t->d_frame, frame, n, kind, stream));
    } else {
        memcpy(t->h_stage, frame, n);
        VT_CUDA(cudaMemcpyAsync(t->d_frame, t->h_stage, n, cudaMemcpyHostToDevice, stream));
    }
    return VT_OK;
}

static void fill_results(vt_tracker* t, vt_result* results) {
    for (int s = 0; s < t->maxT; ++s) {
        vt_result r;
        memset(&r, 0, sizeof(r));
        if (!t->inited[s]) {
            r.status = VT_ERR_NOT_INIT;
        } else {
            const DeviceResult& d = t->h_res[s];
            r.success = d.success, r.score = d.score, r.status = d.status;
            r.bbox = vt_bbox{d.bbox[0], d.bbox[1], d.bbox[2], d.bbox[3]};
            if (d.status == VT_OK && d.success) t->rect_mirror[s] = r.bbox;
        }
        if (results) results[s] = r;
    }
}

// Rows touched by a rect / crosshair, following the reference's clamping exactly
// (src/nv12_convert.rs:181-212,224-241; src/drawing_rgb.rs:55-73).  Returns false if nothing is drawn.
static bool rect_rows(int fmt, long long H, int y, int h, int th, long long& r0, long long& r1) {
    if (H <= 0) return false;
    if (format_is_luma(fmt)) {
        const long long y1 = std::max(y, 0);
        const long long sum = (long long)(int32_t)((uint32_t)y + (uint32_t)h);
        const long long y2 = sum < 0 ? H - 1 : std::min(sum, H - 1);  // negative i32 -> huge usize -> clamped
        const long long t = std::max(th, 1);
        r0 = std::min(y1, std::max(0LL, y2 - t + 1));
        r1 = std::max(y2, std::min(y1 + t - 1, H - 1));
    } else {  // rows y+t, y+rh-1-t (t < thickness) and y..y+rh-1, each bounds-checked per pixel
        const long long t = std::max(th, 1);
        r0 = std::min<long long>(y, (long long)y + h - t), r1 = std::max<long long>((long long)y + h - 1, (long long)y + t - 1);
        if (r1 < 0 || r0 > H - 1) return false;
    }
    r0 = std::max(0LL, std::min(r0, H - 1)), r1 = std::max(0LL, std::min(r1, H - 1));
    return r1 >= r0;
}
static bool cross_rows(long long H, int cy, int size, long long& r0, long long& r1) {
    const long long c = std::max(cy, 0), s = std::max(size, 0);
    r0 = std::max(0LL, std::min(c - s, H - 1)), r1 = std::max(0LL, std::min(c + s, H - 1));
    return r1 >= r0;
}

// rows of the frame touched by the box overlay of the current results, merged
static void box_rows(const vt_tracker* t, std::vector<std::pair<int, int>>& spans) {
    for (int s : t->active) {
        const DeviceResult& d = t->h_res[s];
        if (d.status != VT_OK || !d.success || !(d.score > t->cfg.overlay_gate)) continue;
        long long r0, r1;
        if (rect_rows(t->fmt, t->H, d.bbox[1], d.bbox[3], 3, r0, r1)) spans.emplace_back((int)r0, (int)r1);
        if (cross_rows(t->H, d.bbox[1] + d.bbox[3] / 2, 15, r0, r1)) spans.emplace_back((int)r0, (int)r1);
    }
}

static void merge_spans(std::vector<std::pair<int, int>>& spans) {
    std::sort(spans.begin(), spans.end());
    std::vector<std::pair<int, int>> out;
    for (auto& sp : spans) {
        if (!out.empty() && sp.first <= out.back().second + 8) out.back().second = std::max(out.back().second, sp.second);
        else out.push_back(sp);
    }
    spans.swap(out);
}

// device -> host copy of whole rows [r0, r1] of the drawable plane (Y plane for NV12, the image for RGB24)
static vt_status download_rows(vt_tracker* t, uint8_t* frame, size_t len, const std::vector<std::pair<int, int>>& spans, bool* staged) {
    const size_t pitch = (size_t)t->W * (format_is_luma(t->fmt) ? 1 : 3);
    const bool pinned = is_pinned(frame);
    *staged = !pinned;
    for (auto& sp : spans) {
        const size_t off = (size_t)sp.first * pitch;
        size_t n = (size_t)(sp.second - sp.first + 1) * pitch;
        if (off >= len) continue;
        n = std::min(n, len - off);
        t->d2h_bytes += n;
        VT_CUDA(cudaMemcpyAsync((pinned ? frame : t->h_stage) + off, t->d_frame + off, n, cudaMemcpyDeviceToHost, t->stream));
    }
    return VT_OK;
}
static void unstage_rows(vt_tracker* t, uint8_t* frame, size_t len, const std::vector<std::pair<int, int>>& spans) {
    const size_t pitch = (size_t)t->W * (format_is_luma(t->fmt) ? 1 : 3);
    for (auto& sp : spans) {
        const size_t off = (size_t)sp.first * pitch;
        size_t n = (size_t)(sp.second - sp.first + 1) * pitch;
        if (off >= len) continue;
        memcpy(frame + off, t->h_stage + off, std::min(n, len - off));
    }
}

static void collect_timing(vt_tracker* t) {
    // stage boundaries stamped on the device (ns): submit, crop start, ViT start, decode start, decode end, overlay end
    const unsigned long long* st = t->h_stamps;
    auto span = [&](int a, int b) { return st[b] > st[a] && st[a] ? (float)((double)(st[b] - st[a]) * 1e-6) : 0.f; };
    const int last_dev = t->cfg.box_overlay ? ST_OVL_END : ST_DEC_END;
    const float wall = (float)std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t->t_submit).count();
    float ms[7];
    ms[0] = span(ST_SUBMIT, ST_PRE), ms[1] = span(ST_PRE, ST_VIT), ms[2] = span(ST_VIT, ST_DEC), ms[3] = span(ST_DEC, ST_DEC_END);
    ms[4] = t->cfg.box_overlay ? span(ST_DEC_END, ST_OVL_END) : 0.f;
    const float dev = span(ST_SUBMIT, last_dev);
    ms[5] = wall > dev ? wall - dev : 0.f;  // results (and overlay rows) back in host memory + completion latency, host clock
    ms[6] = wall;
    if (t->active.empty()) ms[1] = ms[2] = ms[3] = ms[4] = 0.f;
    memcpy(t->last, ms, sizeof(ms));
    t->r_h2d.push(ms[0]), t->r_pre.push(ms[1]), t->r_vit.push(ms[2]), t->r_dec.push(ms[3]), t->r_ovl.push(ms[4]), t->r_d2h.push(ms[5]),
        t->r_tot.push(ms[6]);
}

static inline double now_us() {
    return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

static void bind_slot(vt_tracker* t, int slot) {
    t->h_res = t->h_blk[slot];
    t->h_stamps = reinterpret_cast<unsigned long long*>(t->h_res + t->maxT);
    t->h_tc_err = reinterpret_cast<int*>(t->h_stamps + ST_COUNT);
}

static vt_status submit_common(vt_tracker* t, uint8_t* frame, const uint8_t* d_src, size_t len) {
    if (t->q_count >= vt_tracker::kQueue) {
        set_error("%d frames are already in flight on this handle", t->q_count);
        return VT_ERR_INVALID;
    }
    const bool pinned = d_src || is_pinned(frame);
    if (t->q_count > 0 && (!pinned || t->q[t->q_head].pageable)) {
        set_error("pageable host frames are staged through one buffer and cannot be pipelined: wait() first or use pinned frames");
        return VT_ERR_INVALID;
    }
    const double hp0 = t->hostprof ? now_us() : 0;
    const int slot = (t->q_head + t->q_count) % vt_tracker::kQueue;
    vt_tracker::Slot& q = t->q[slot];
    q.t_submit = std::chrono::steady_clock::now();
    q.frame = d_src ? nullptr : frame, q.len = len, q.pageable = !pinned;
    // zero-copy overlay mirror: only when the caller's frame is pinned (device-mapped under UVA)
    q.mirrored = t->cfg.box_overlay && frame && !d_src && len >= t->frame_bytes && pinned;
    // device-resident frame: track (and draw the box) straight in the caller's device memory — no device->device copy
    const bool in_place = d_src && len >= t->frame_bytes;
    if (!in_place) t->d_frame = t->d_frames[slot];  // the slot's own frame buffer: the other one may still be read by the frame in flight
    if (!in_place && !d_src && t->q_count > 0) {
        // pipelined host frame: upload on the copy stream while the frame in flight computes (whole frame: the host mirror of
        // rect_last lags by one frame); the main stream picks it up through an event
        vt_status st = upload_frame(t, frame, len, false, false, t->copy_stream);
        if (st != VT_OK) return st;
        VT_CUDA(cudaEventRecord(t->ev_up[slot], t->copy_stream));
        VT_CUDA(cudaStreamWaitEvent(t->stream, t->ev_up[slot], 0));
    }
    VT_CUDA(launch_stamp(t->d_stamps + ST_SUBMIT, t->d_frame_slot, in_place ? d_src : t->d_frame, t->d_host_slot, q.mirrored ? frame : nullptr,
                         t->d_hblk_slot, reinterpret_cast<uint32_t*>(t->h_blk[slot]), t->stream));
    ++t->kernel_launches;
    if (in_place) {
        t->frame_valid = 1;
    } else if (d_src) {  // short device frame: device->device copy of what there is (NV12: black frame, src/nv12_convert.rs:48-50)
        vt_status st = upload_frame(t, d_src, len, false, true);
        if (st != VT_OK) return st;
    } else if (t->q_count == 0) {
        vt_status st = upload_frame(t, frame, len, true);  // the host mirror of rect_last is exact: the search windows suffice
        if (st != VT_OK) return st;
    }
    const double hp1 = t->hostprof ? now_us() : 0;
    vt_status st = run_forward(t);
    if (st != VT_OK) return st;
    const double hp2 = t->hostprof ? now_us() : 0;
    // the result block reaches t->h_blk[slot] through publish_kernel, the last kernel of run_forward (no target active: nothing ran)
    if (t->active.empty()) VT_CUDA(cudaMemcpyAsync(t->h_blk[slot], t->d_res, t->res_block_bytes, cudaMemcpyDeviceToHost, t->stream));
    VT_CUDA(cudaEventRecord(t->q_done[slot], t->stream));
    t->d2h_bytes += t->res_block_bytes;
    if (t->hostprof) t->hp[0] += hp1 - hp0, t->hp[1] += hp2 - hp1, t->hp[2] += now_us() - hp2;
    ++t->q_count;
    t->in_flight = true;
    return VT_OK;
}

static vt_status wait_common(vt_tracker* t, vt_result* results) {
    if (t->q_count == 0) {
        set_error("no frame in flight");
        return VT_ERR_INVALID;
    }
    const int slot = t->q_head;
    const vt_tracker::Slot q = t->q[slot];
    t->q_head = (t->q_head + 1) % vt_tracker::kQueue;
    --t->q_count;
    t->in_flight = t->q_count > 0;
    t->t_submit = q.t_submit, t->inflight_frame = q.frame, t->inflight_mirrored = q.mirrored;
    const size_t len = q.len;
    const double hp0 = t->hostprof ? now_us() : 0;
    VT_CUDA(cudaEventSynchronize(t->q_done[slot]));
    bind_slot(t, slot);
    const double hp1 = t->hostprof ? now_us() : 0;
    if (*t->h_tc_err) {  // travels with the results; reset on the device for the next frame
        cudaMemsetAsync(t->d_tc_err, 0, sizeof(int), t->stream);
        *t->h_tc_err = 0;
        set_error("tcgen05 path: a bounded mbarrier wait expired (pipeline protocol error)");
        return VT_ERR_CUDA;
    }
    fill_results(t, results);
    const double hp2 = t->hostprof ? now_us() : 0;
    if (t->cfg.box_overlay && t->inflight_frame && t->inflight_mirrored) {
        // the overlay kernel wrote the box pixels straight into the caller's pinned frame: count them as device->host traffic
        for (int sl : t->active) {
            const DeviceResult& d = t->h_res[sl];
            if (d.status == VT_OK && d.success && d.score > t->cfg.overlay_gate) t->d2h_bytes += 6ull * (size_t)(std::max(d.bbox[2], 0) + std::max(d.bbox[3], 0)) + 62;
        }
    } else if (t->cfg.box_overlay && t->inflight_frame) {  // pageable frame (never pipelined): copy the touched rows back
        std::vector<std::pair<int, int>> spans;
        box_rows(t, spans);
        merge_spans(spans);
        if (!spans.empty()) {
            bool staged = false;
            vt_status st = download_rows(t, t->inflight_frame, len, spans, &staged);
            if (st != VT_OK) return st;
            VT_CUDA(cudaStreamSynchronize(t->stream));
            if (staged) unstage_rows(t, t->inflight_frame, len, spans);
        }
    }
    const double hp3 = t->hostprof ? now_us() : 0;
    collect_timing(t);
    ++t->frames;
    if (t->hostprof) t->hp[3] += hp1 - hp0, t->hp[4] += hp2 - hp1, t->hp[5] += hp3 - hp2, t->hp[6] += now_us() - hp3, ++t->hp_n;
    return VT_OK;
}

}  // namespace vt

// ==================================================================================================
// C ABI
// ==================================================================================================
extern "C" {

int32_t vt_abi_version(void) { return VT_ABI_VERSION; }
const char* vt_last_error(void) { return vt::g_err; }

void vt_config_default(vt_config* c) {
    if (!c) return;
    memset(c, 0, sizeof(*c));
    c->struct_size = sizeof(vt_config);
    c->format = VT_FMT_NV12;
    c->width = 1920, c->height = 1080;  // src/pipeline.rs:26-27
    c->max_targets = 1;
    c->score_threshold = 0.20f;
    c->gemm_mode = VT_GEMM_TCGEN05_BF16X3;  // the parity-safe tensor-core path; VT_GEMM_FP32_SIMT is the numerically anchoring fallback
    c->use_cuda_graph = 1;
    c->box_overlay = 0;
    c->overlay_gate = 0.25f;  // src/tracker_context.rs:93,122
}

vt_status vt_weights_probe(const char* path, int32_t shape_out[5]) {
    int32_t hdr[7];
    size_t n = 0;
    vt_status st = read_weight_header(path, nullptr, hdr, &n, nullptr);
    if (st == VT_OK && shape_out) memcpy(shape_out, hdr, 5 * sizeof(int32_t));
    return st;
}

vt_status vt_alloc_pinned(size_t bytes, void** out) {
    if (!out) return VT_ERR_INVALID;
    VT_CUDA(cudaHostAlloc(out, bytes, cudaHostAllocDefault));
    return VT_OK;
}
void vt_free_pinned(void* p) {
    if (p) cudaFreeHost(p);
}

void vt_tracker_destroy(vt_tracker* t) {
    if (!t) return;
    if (t->hostprof && t->hp_n) {
        const double n = (double)t->hp_n;
        fprintf(stderr, "[vt hostprof] frames %llu  submit: upload %.1f  graph launch %.1f  result copy+event %.1f | wait: stream sync %.1f  "
                        "err check+results %.1f  overlay rows %.1f  timing %.1f (us / frame)\n",
                (unsigned long long)t->hp_n, t->hp[0] / n, t->hp[1] / n, t->hp[2] / n, t->hp[3] / n, t->hp[4] / n, t->hp[5] / n, t->hp[6] / n);
    }
    cudaSetDevice(t->cfg.device);
    if (t->counted) g_live_handles[t->cfg.device].fetch_sub(1);
    if (t->stream) cudaStreamSynchronize(t->stream);
    for (auto& kv : t->graphs) cudaGraphExecDestroy(kv.second);
    for (auto& e : t->ev)
        if (e) cudaEventDestroy(e);
    if (t->d_fmt_in) cudaFree(t->d_fmt_in);
    if (t->d_fmt_out) cudaFree(t->d_fmt_out);
    if (t->copy_stream) cudaStreamSynchronize(t->copy_stream), cudaStreamDestroy(t->copy_stream);
    for (auto& e : t->ev_up)
        if (e) cudaEventDestroy(e);
    void* dev[] = {t->d_lut, t->d_hann, t->d_frames[0], t->d_frames[1], t->d_rgb, t->d_state, t->d_slots, t->d_res, t->d_maps, t->d_cmds,
                   t->patches_x, t->patches_z, t->Zemb, t->X, t->QKV, t->ATT, t->HID, t->Yf, t->H1, t->d_dbg,
                   t->px_hi, t->px_lo, t->pz_hi, t->pz_lo, t->ln_hi, t->ln_lo, t->att_hi, t->att_lo,
                   t->hid_hi, t->hid_lo, t->yf_hi, t->yf_lo, t->q_hi, t->q_lo, t->k_hi, t->k_lo, t->vt_hi, t->vt_lo, t->zln_hi, t->zln_lo, t->d_trace, t->Pbuf, t->Phead, t->d_cand, t->d_counters};
    for (void* p : dev)
        if (p) cudaFree(p);
    if (t->h_stage) cudaFreeHost(t->h_stage);
    for (int i = 0; i < vt_tracker::kQueue; ++i) {
        if (t->h_blk[i]) cudaFreeHost(t->h_blk[i]);
        if (t->q_done[i]) cudaEventDestroy(t->q_done[i]);
    }
    if (t->d_host_slot) cudaFree(t->d_host_slot);
    if (t->d_hblk_slot) cudaFree(t->d_hblk_slot);
    if (t->d_frame_slot) cudaFree(t->d_frame_slot);
    if (t->h_cmds) cudaFreeHost(t->h_cmds);
    if (t->stream) cudaStreamDestroy(t->stream);
    delete t;
}

vt_status vt_tracker_create(const vt_config* cfg, vt_tracker** out) {
    if (!cfg || !out || !cfg->weights_path || cfg->width <= 0 || cfg->height <= 0 || cfg->max_targets <= 0 || cfg->max_targets > 64 ||
        (cfg->format != VT_FMT_NV12 && cfg->format != VT_FMT_RGB24 && cfg->format != VT_FMT_GRAY8)) {
        set_error("vt_tracker_create: invalid configuration");
        return VT_ERR_INVALID;
    }
    if (cfg->gemm_mode < VT_GEMM_FP32_SIMT || cfg->gemm_mode > VT_GEMM_TCGEN05_FP16) {
        set_error("vt_tracker_create: unknown gemm_mode %d", cfg->gemm_mode);
        return VT_ERR_INVALID;
    }
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) {
        cudaGetLastError();
        set_error("no CUDA device: libvittrack_b200 has no CPU fallback");
        return VT_ERR_CUDA;
    }
    VT_CUDA(cudaSetDevice(cfg->device));
    vt_tracker* t = new vt_tracker();
    t->cfg = *cfg;
    t->cfg.weights_path = nullptr;
    t->W = cfg->width, t->H = cfg->height, t->fmt = cfg->format, t->maxT = cfg->max_targets;
    t->frame_bytes = cfg->format == VT_FMT_NV12    ? (size_t)t->W * t->H + (size_t)((t->H + 1) / 2) * t->W + (t->W & 1)
                     : cfg->format == VT_FMT_GRAY8 ? (size_t)t->W * t->H
                                                   : (size_t)t->W * t->H * 3;
    t->threshold = cfg->score_threshold > 0.f ? cfg->score_threshold : 0.20f;
    t->debug_capture = cfg->debug_capture;
    t->hostprof = getenv("VT_B200_HOSTPROF") != nullptr;
    auto fail = [&](vt_status st) {
        vt_tracker_destroy(t);
        return st;
    };
    vt_status st = load_weights(t, cfg->weights_path);
    if (st != VT_OK) return fail(st);
#define VT_TRY(call)                                                                                             \
    do {                                                                                                         \
        cudaError_t e_ = (call);                                                                                 \
        if (e_ != cudaSuccess) {                                                                                 \
            set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__);             \
            return fail(VT_ERR_CUDA);                                                                            \
        }                                                                                                        \
    } while (0)
    VT_TRY(cudaStreamCreateWithFlags(&t->stream, cudaStreamNonBlocking));
    for (auto& e : t->ev) VT_TRY(cudaEventCreate(&e));
    const size_t D = t->D, Hd = t->hidden, C = t->head_ch, B = t->maxT;
    // A.4 normalisation LUT: (v/255 - mean_c)/std_c in double, rounded once to fp32 (channels in memory order)
    {
        const double mean[3] = {0.485, 0.456, 0.406}, stdv[3] = {0.229, 0.224, 0.225};
        float lut[768];
        for (int c = 0; c < 3; ++c)
            for (int v = 0; v < 256; ++v) lut[c * 256 + v] = (float)(((double)v / 255.0 - mean[c]) / stdv[c]);
        VT_TRY(cudaMalloc(&t->d_lut, sizeof(lut)));
        VT_TRY(cudaMemcpy(t->d_lut, lut, sizeof(lut), cudaMemcpyHostToDevice));
        // A.5 hann window, fp32 exactly as OpenCV builds it
        float h1[16], hann[256];
        for (int i = 0; i < 16; ++i) h1[i] = 0.5f * (1.f - cosf((float)(2 * M_PI / 17) * (float)(i + 1)));
        for (int y = 0; y < 16; ++y)
            for (int x = 0; x < 16; ++x) hann[y * 16 + x] = h1[y] * h1[x];
        VT_TRY(cudaMalloc(&t->d_hann, sizeof(hann)));
        VT_TRY(cudaMemcpy(t->d_hann, hann, sizeof(hann), cudaMemcpyHostToDevice));
    }
    for (int i = 0; i < 2; ++i) {
        VT_TRY(cudaMalloc(&t->d_frames[i], t->frame_bytes + 256));
        VT_TRY(cudaMemset(t->d_frames[i], 0, t->frame_bytes + 256));
        VT_TRY(cudaEventCreateWithFlags(&t->ev_up[i], cudaEventDisableTiming));
    }
    t->d_frame = t->d_frames[0];
    VT_TRY(cudaStreamCreateWithFlags(&t->copy_stream, cudaStreamNonBlocking));
    t->h_stage_bytes = t->frame_bytes + 256;
    VT_TRY(cudaHostAlloc(&t->h_stage, t->h_stage_bytes, cudaHostAllocDefault));
    VT_TRY(cudaMalloc(&t->d_state, sizeof(TargetState) * B));
    VT_TRY(cudaMemset(t->d_state, 0, sizeof(TargetState) * B));
    VT_TRY(cudaMalloc(&t->d_slots, sizeof(int32_t) * B));
    t->res_block_bytes = sizeof(DeviceResult) * B + sizeof(unsigned long long) * ST_COUNT + 2 * sizeof(int);
    VT_TRY(cudaMalloc(&t->d_res, t->res_block_bytes));
    VT_TRY(cudaMemset(t->d_res, 0, t->res_block_bytes));
    for (int i = 0; i < vt_tracker::kQueue; ++i) {
        VT_TRY(cudaHostAlloc(&t->h_blk[i], t->res_block_bytes, cudaHostAllocDefault));
        memset(t->h_blk[i], 0, t->res_block_bytes);
        VT_TRY(cudaEventCreateWithFlags(&t->q_done[i], cudaEventDisableTiming));
    }
    t->d_stamps = reinterpret_cast<unsigned long long*>(t->d_res + B);
    t->d_tc_err = reinterpret_cast<int*>(t->d_stamps + ST_COUNT);
    bind_slot(t, 0);
    VT_TRY(cudaMalloc(&t->d_frame_slot, sizeof(uint8_t*)));
    VT_TRY(cudaMemcpy(t->d_frame_slot, &t->d_frame, sizeof(uint8_t*), cudaMemcpyHostToDevice));
    VT_TRY(cudaMalloc(&t->d_host_slot, sizeof(uint8_t*)));
    VT_TRY(cudaMemset(t->d_host_slot, 0, sizeof(uint8_t*)));
    VT_TRY(cudaMalloc(&t->d_hblk_slot, sizeof(uint32_t*)));
    VT_TRY(cudaMemset(t->d_hblk_slot, 0, sizeof(uint32_t*)));
    VT_TRY(cudaMalloc(&t->d_maps, sizeof(float) * 1280 * B));
    VT_TRY(cudaMemset(t->d_maps, 0, sizeof(float) * 1280 * B));
    VT_TRY(cudaMalloc(&t->d_cmds, sizeof(OverlayCmdDev) * kMaxCmds));
    VT_TRY(cudaHostAlloc(&t->h_cmds, sizeof(OverlayCmdDev) * kMaxCmds, cudaHostAllocDefault));
    VT_TRY(cudaMalloc(&t->patches_x, sizeof(float) * B * kNTx * kPatchK));
    VT_TRY(cudaMalloc(&t->patches_z, sizeof(float) * kNTz * kPatchK));
    VT_TRY(cudaMalloc(&t->Zemb, sizeof(float) * B * kNTz * D));
    VT_TRY(cudaMalloc(&t->X, sizeof(float) * B * kNTok * D));
    VT_TRY(cudaMalloc(&t->QKV, sizeof(float) * B * kNTok * 3 * D));
    VT_TRY(cudaMalloc(&t->ATT, sizeof(float) * B * kNTok * D));
    VT_TRY(cudaMalloc(&t->HID, sizeof(float) * B * kNTok * Hd));
    VT_TRY(cudaMalloc(&t->Yf, sizeof(float) * B * kNTx * D));
    VT_TRY(cudaMalloc(&t->H1, sizeof(float) * B * kNTx * C));
    VT_TRY(cudaMemset(t->patches_x, 0, sizeof(float) * B * kNTx * kPatchK));
    VT_TRY(cudaMemset(t->patches_z, 0, sizeof(float) * kNTz * kPatchK));
    if (t->debug_capture) VT_TRY(cudaMalloc(&t->d_dbg, sizeof(float) * (size_t)(t->depth + 1) * B * kNTok * D));
    t->nsplit = cfg->gemm_mode == VT_GEMM_TCGEN05_BF16X3 ? 3 : (cfg->gemm_mode == VT_GEMM_TCGEN05_BF16 ? 1 : (cfg->gemm_mode == VT_GEMM_TCGEN05_FP16 ? 2 : 0));
    t->f16 = t->nsplit == 2;
    if (t->nsplit) {
        if (D % 64 || Hd % 64 || C % 64) {
            set_error("the tcgen05 path needs D, hidden and head_ch to be multiples of 64 (D=%zu hidden=%zu head_ch=%zu)", D, Hd, C);
            return fail(VT_ERR_WEIGHTS);
        }
        VT_TRY(tc_gemm_setup());
        t->fuse_ln = D / 64 <= 8 && !getenv("VT_B200_NO_FUSE_LN");
        t->pdl = !getenv("VT_B200_NO_PDL");
        t->spread_ok = !getenv("VT_B200_NO_SPREAD");
        if (const char* e = getenv("VT_B200_UNCHAIN_N")) t->unchain_n = atoi(e);
        t->chain_mlp = t->fuse_ln && D <= 192 && !getenv("VT_B200_NO_CHAIN");
        if (t->chain_mlp) VT_TRY(cudaMalloc(&t->Pbuf, sizeof(float) * (Hd / 64) * B * kNTok * D));
        t->att_chain_ok = t->chain_mlp && t->fuse_ln && D / t->heads == 64 && (int)(Hd / 64) >= t->heads && t->nsplit && !getenv("VT_B200_NO_ATT_CHAIN");
        t->split_k = t->chain_mlp && Hd / 64 >= 4 && (C == 64 || C == 128) && !getenv("VT_B200_NO_SPLITK");
        if (t->split_k) {
            VT_TRY(cudaMalloc(&t->Phead, sizeof(float) * 9 * B * kNTx * C));
            VT_TRY(cudaMalloc(&t->d_cand, sizeof(float) * B * 16 * 8));
            VT_TRY(cudaMalloc(&t->d_counters, sizeof(unsigned) * B));
            VT_TRY(cudaMemset(t->d_counters, 0, sizeof(unsigned) * B));
        }
        {   // bf16 (hi, lo) split of the shared weights: done once per WeightSet
            std::lock_guard<std::mutex> lock(t->wset->split_mutex);
            if (!t->wset->w_hi) {
                const size_t nw = t->n_weights;
                VT_TRY(cudaMalloc(&t->wset->w_hi, nw * 2));
                VT_TRY(cudaMalloc(&t->wset->w_lo, nw * 2));
                VT_TRY(launch_split_bf16(t->d_weights, t->wset->w_hi, t->wset->w_lo, nw, t->stream));
                VT_TRY(cudaStreamSynchronize(t->stream));
            }
            t->w_hi = t->wset->w_hi, t->w_lo = t->wset->w_lo;
            if (t->f16) {  // single-pass fp16 operands: an fp16 copy of the weights stands in for the hi part
                if (!t->wset->w_f16) {
                    VT_TRY(cudaMalloc(&t->wset->w_f16, t->n_weights * 2));
                    VT_TRY(launch_split_bf16(t->d_weights, t->wset->w_f16, nullptr, t->n_weights, t->stream));
                    VT_TRY(cudaStreamSynchronize(t->stream));
                }
                t->w_hi = t->wset->w_f16;
            }
        }
        auto balloc = [&](__nv_bfloat16** hi, __nv_bfloat16** lo, size_t n) -> cudaError_t {
            cudaError_t e = cudaMalloc(hi, n * 2);
            if (e == cudaSuccess) e = cudaMalloc(lo, n * 2);
            if (e == cudaSuccess) e = cudaMemset(*hi, 0, n * 2);
            if (e == cudaSuccess) e = cudaMemset(*lo, 0, n * 2);
            return e;
        };
        VT_TRY(balloc(&t->px_hi, &t->px_lo, B * kNTx * kPatchK));
        VT_TRY(balloc(&t->pz_hi, &t->pz_lo, (size_t)128 * kPatchK));  // one 128-row tile; rows 64..127 stay zero
        VT_TRY(balloc(&t->ln_hi, &t->ln_lo, B * kNTok * D));
        VT_TRY(balloc(&t->zln_hi, &t->zln_lo, B * kNTz * D));
        VT_TRY(balloc(&t->att_hi, &t->att_lo, B * kNTok * D));
        VT_TRY(balloc(&t->hid_hi, &t->hid_lo, B * kNTok * Hd));
        VT_TRY(balloc(&t->yf_hi, &t->yf_lo, B * kNTx * D));
        t->tc_attention = (D / t->heads == 64);
        if (t->tc_attention) {
            const size_t nq = B * t->heads * kNTok * 64;
            VT_TRY(balloc(&t->q_hi, &t->q_lo, nq));
            VT_TRY(balloc(&t->k_hi, &t->k_lo, nq));
            VT_TRY(balloc(&t->vt_hi, &t->vt_lo, nq));
            VT_TRY(tc_attention_setup());
            if (!tc_attention_plan_init(&t->plan_att, t->q_hi, t->q_lo, t->k_hi, t->k_lo, t->vt_hi, t->vt_lo, (int)(B * t->heads), t->att_hi, t->att_lo, (int)D, (int)B))
                return fail(VT_ERR_CUDA);
        }
        auto whi = [&](const float* w) { return t->w_hi + (w - t->d_weights); };
        auto wlo = [&](const float* w) { return t->w_lo + (w - t->d_weights); };
        bool ok = true;
        const uint64_t rows = B * kNTok;
        // Column-tile width of QKV / proj / patch / head.  128x32 tiles (VT_B200_TILE32=1: twice the CTAs, half the epilogue per CTA) were
        // measured and bring nothing: the accumulator is ready at the same 2.05 us (the 96 KB A tile per CTA bounds it, not the UMMAs) and
        // the 6-CTA LayerNorm cluster of proj is slower than the 3-CTA one (profiles/r1d_final.md).  FC1 needs the 64-column tile anyway.
        const int bn_lat = getenv("VT_B200_TILE32") ? 32 : 64;
        // outputs of the GEMM epilogues (TcOut: dense [planes][targets][heads][rows][cols]) and the flat residual TMA source
        CUtensorMap mXres;
        ok &= tc_resid_map(&mXres, t->X, B * kNTok, D);
        const int Bi = (int)B, Di = (int)D;
        const TcOut oX = tc_out(t->X, 4, Di, kNTok, 1, Bi, 1);
        const TcOut oLn[2] = {tc_out(t->ln_hi, 2, Di, kNTok, 1, Bi, 1), tc_out(t->ln_lo, 2, Di, kNTok, 1, Bi, 1)};
        const TcOut oYf[2] = {tc_out(t->yf_hi, 2, Di, kNTx, 1, Bi, 1), tc_out(t->yf_lo, 2, Di, kNTx, 1, Bi, 1)};
        const TcOut oHid[2] = {tc_out(t->hid_hi, 2, (int)Hd, kNTok, 1, Bi, 1), tc_out(t->hid_lo, 2, (int)Hd, kNTok, 1, Bi, 1)};
        const TcOut oH1 = tc_out(t->H1, 4, (int)C, kNTx, 1, Bi, 1), oZ = tc_out(t->Zemb, 4, Di, kNTz, 1, Bi, 1);
        const TcOut oQKV = tc_out(t->QKV, 4, 3 * Di, kNTok, 1, Bi, 1);
        TcOut oQ[6] = {};
        if (t->tc_attention) {
            oQ[0] = tc_out(t->q_hi, 2, 64, kNTok, t->heads, Bi, 1), oQ[1] = tc_out(t->q_lo, 2, 64, kNTok, t->heads, Bi, 1);
            oQ[2] = tc_out(t->k_hi, 2, 64, kNTok, t->heads, Bi, 1), oQ[3] = tc_out(t->k_lo, 2, 64, kNTok, t->heads, Bi, 1);
            oQ[4] = tc_out_vt(t->vt_hi, kNTok, t->heads, Bi), oQ[5] = tc_out_vt(t->vt_lo, kNTok, t->heads, Bi);
        }
        // patch embed (search): A = patches [B*256, 768] -> X rows 64.. of every target, + pos_x
        ok &= tc_plan_init(&t->plan_patch_x, t->px_hi, t->px_lo, B * kNTx, whi(t->patch_w), wlo(t->patch_w), (int)D, kPatchK, 0, 0, bn_lat);
        {
            TcGemmArgs& a = t->plan_patch_x.args;
            a.bias = t->patch_b, a.pos = t->pos_x, a.pos_rows = kNTx;
            a.period = kNTx, a.c_on = 1, a.c_row_off = kNTz, a.c = oX;
            if (t->split_k) {  // 4 x K = 192 slices -> fp32 partials [4][B][256][D]; bias, pos, LN1 happen in reduce_ln_kernel
                a.bias = nullptr, a.pos = nullptr, a.c_row_off = 0, a.kb_per_split = kPatchK / 64 / 4;
                a.c = tc_out(t->Pbuf, 4, Di, kNTx, 1, Bi, 4);
            } else if (t->fuse_ln) {  // LN1 of block 0 for the search rows, straight into the first QKV GEMM's A operand
                a.ln_g = t->blk[0].ln1_g, a.ln_b = t->blk[0].ln1_b, a.ln_row_off = kNTz;
                a.ln_out[0] = oLn[0], a.ln_out[1] = oLn[1];
            }
        }
        // patch embed (template, at init): one 128-row tile whose rows 64.. are clipped; batch_off = the target slot, set per call
        ok &= tc_plan_init(&t->plan_patch_z, t->pz_hi, t->pz_lo, 128, whi(t->patch_w), wlo(t->patch_w), (int)D, kPatchK, 0, 0);
        {
            TcGemmArgs& a = t->plan_patch_z.args;
            a.bias = t->patch_b, a.pos = t->pos_z, a.pos_rows = kNTz;
            a.period = 128, a.c_on = 1, a.c = oZ;
        }
        t->plans.resize(t->depth);
        for (int l = 0; l < t->depth && ok; ++l) {
            const BlockW& b = t->blk[l];
            vt_tracker::BlockPlans& p = t->plans[l];
            ok &= tc_plan_init(&p.qkv, t->ln_hi, t->ln_lo, rows, whi(b.qkv_w), wlo(b.qkv_w), (int)(3 * D), (int)D, 0, 0, bn_lat);
            p.qkv.args.bias = b.qkv_b, p.qkv.args.period = kNTok;
            if (t->tc_attention) {
                p.qkv.args.o_mode = 2;
                for (int i = 0; i < 6; ++i) p.qkv.args.o[i] = oQ[i];
            } else {
                p.qkv.args.c_on = 1, p.qkv.args.c = oQKV;
            }
            ok &= tc_plan_init(&p.proj, t->att_hi, t->att_lo, rows, whi(b.proj_w), wlo(b.proj_w), (int)D, (int)D, 0, 0,
                               D / bn_lat <= 8 ? bn_lat : 64);
            p.proj.args.bias = b.proj_b, p.proj.args.period = kNTok, p.proj.args.residual = 1, p.proj.args.c_on = 1;
            p.proj.maps.R = mXres, p.proj.args.c = oX;
            if (t->fuse_ln) p.proj.args.ln_g = b.ln2_g, p.proj.args.ln_b = b.ln2_b, p.proj.args.ln_out[0] = oLn[0], p.proj.args.ln_out[1] = oLn[1];
            ok &= tc_plan_init(&p.fc1, t->ln_hi, t->ln_lo, rows, whi(b.fc1_w), wlo(b.fc1_w), (int)Hd, (int)D, 0, 0);
            p.fc1.args.bias = b.fc1_b, p.fc1.args.gelu = 1, p.fc1.args.period = kNTok, p.fc1.args.o_mode = 1;
            p.fc1.args.o[0] = oHid[0], p.fc1.args.o[1] = oHid[1];
            if (t->chain_mlp) ok &= tc_plan_chain(&p.fc1, whi(b.fc2_w), wlo(b.fc2_w), (int)D, t->Pbuf, kNTok, B);
            if (t->att_chain_ok) {
                p.att = t->plan_att;
                ok &= tc_attention_plan_chain(&p.att, whi(b.proj_w), wlo(b.proj_w), t->Pbuf, (int64_t)B * kNTok * D);
            }
            ok &= tc_plan_init(&p.fc2, t->hid_hi, t->hid_lo, rows, whi(b.fc2_w), wlo(b.fc2_w), (int)D, (int)Hd, 0, 0);
            p.fc2.args.bias = b.fc2_b, p.fc2.args.period = kNTok, p.fc2.args.residual = 1, p.fc2.args.c_on = 1;
            p.fc2.maps.R = mXres, p.fc2.args.c = oX;
            if (t->fuse_ln) {
                TcGemmArgs& a = p.fc2.args;
                if (l + 1 < t->depth) {
                    a.ln_g = t->blk[l + 1].ln1_g, a.ln_b = t->blk[l + 1].ln1_b, a.ln_out[0] = oLn[0], a.ln_out[1] = oLn[1];
                } else {  // final LN, search rows only -> the head conv's [B,16,16,D] grid (the template rows fall outside and are clipped)
                    a.ln_g = t->lnf_g, a.ln_b = t->lnf_b, a.ln_row_off = -kNTz, a.ln_out[0] = oYf[0], a.ln_out[1] = oYf[1];
                }
            }
        }
        // 3x3 head conv: A gathered by TMA from the [B,16,16,D] final-LN grid (zero fill = zero padding), weights [C][tap][D]
        ok &= tc_plan_init(&t->plan_head, t->yf_hi, t->yf_lo, 0, whi(t->h1_w), wlo(t->h1_w), (int)C, (int)(9 * D), (int)D, (int)B, bn_lat);
        t->plan_head.args.bias = t->h1_b, t->plan_head.args.relu = 1, t->plan_head.args.period = kNTx, t->plan_head.args.c_on = 1;
        t->plan_head.args.c = oH1;
        if (t->split_k) {  // one tap per slice -> fp32 partials [9][B][256][C]; bias, ReLU, 1x1 conv and decode in head_decode_kernel
            t->plan_head.args.bias = nullptr, t->plan_head.args.relu = 0, t->plan_head.args.kb_per_split = (int)(D / 64);
            t->plan_head.args.c = tc_out(t->Phead, 4, (int)C, kNTx, 1, Bi, 9);
        }
        if (!ok) return fail(VT_ERR_CUDA);
        for (TcGemmPlan* p : {&t->plan_patch_x, &t->plan_patch_z, &t->plan_head}) p->args.err = t->d_tc_err;
        for (auto& p : t->plans) p.qkv.args.err = p.proj.args.err = p.fc1.args.err = p.fc2.args.err = t->d_tc_err;
        if (getenv("VT_B200_TRACE")) {
            VT_TRY(cudaMalloc(&t->d_trace, tc::kTraceWords * 8));
            VT_TRY(cudaMemset(t->d_trace, 0, tc::kTraceWords * 8));
            t->plan_patch_x.args.trace = t->plan_head.args.trace = t->d_trace;
            t->plan_patch_x.args.trace_id = 1, t->plan_head.args.trace_id = 6;
            for (auto& p : t->plans) {
                p.qkv.args.trace = p.proj.args.trace = p.fc1.args.trace = p.fc2.args.trace = t->d_trace;
                p.qkv.args.trace_id = 2, p.proj.args.trace_id = 3, p.fc1.args.trace_id = 4, p.fc2.args.trace_id = 5;
                if (getenv("VT_B200_TRACE_LASTX")) p.qkv.args.trace_id |= 0x100, p.fc1.args.trace_id |= 0x100;  // trace the last column tile
            }
        }
    }
    t->rect_mirror.assign(B, vt_bbox{0, 0, 0, 0});
    t->inited.assign(B, 0);
    VT_TRY(cudaStreamSynchronize(t->stream));
#undef VT_TRY
    if (t->cfg.device >= 0 && t->cfg.device < kMaxDevices) g_live_handles[t->cfg.device].fetch_add(1), t->counted = true;
    *out = t;
    return VT_OK;
}

vt_status vt_tracker_init(vt_tracker* t, int32_t target, const uint8_t* frame, size_t len, vt_bbox box) {
    if (!t || !frame || target < 0 || target >= t->maxT) {
        set_error("vt_tracker_init: invalid argument");
        return VT_ERR_INVALID;
    }
    if (t->in_flight) {
        set_error("vt_tracker_init: a frame is in flight");
        return VT_ERR_INVALID;
    }
    VT_CUDA(cudaSetDevice(t->cfg.device));
    // the template window must intersect the frame (cv2 raises an ROI assertion otherwise, App. A.1)
    {
        if (box.width <= 0 || box.height <= 0) {
            set_error("vt_tracker_init: empty box");
            return VT_ERR_CROP_OUTSIDE;
        }
        const int c = (int)ceil(sqrt((double)((long long)box.width * box.height)) * 2.0);
        const int x1 = box.x + (box.width - c) / 2, y1 = box.y + (box.height - c) / 2;
        const int pl = std::max(0, -x1), pt = std::max(0, -y1), pr = std::max(x1 + c - t->W, 0), pb = std::max(y1 + c - t->H, 0);
        if (c - pl - pr <= 0 || c - pt - pb <= 0) {
            set_error("vt_tracker_init: template window lies outside the frame");
            return VT_ERR_CROP_OUTSIDE;
        }
    }
    vt_status st = upload_frame(t, frame, len);
    if (st != VT_OK) return st;
    TargetState hs;
    memset(&hs, 0, sizeof(hs));
    hs.rect[0] = box.x, hs.rect[1] = box.y, hs.rect[2] = box.width, hs.rect[3] = box.height, hs.active = 1;
    int32_t slot = target;
    VT_CUDA(cudaMemcpyAsync(t->d_state + target, &hs, sizeof(hs), cudaMemcpyHostToDevice, t->stream));
    // d_slots is reused as a one-element list for the template pass, then restored
    VT_CUDA(cudaMemcpyAsync(t->d_slots, &slot, sizeof(slot), cudaMemcpyHostToDevice, t->stream));
    FrameDesc fd{t->d_frame, t->W, t->H, t->fmt, t->frame_valid, nullptr};
    int launches = 0;
    VT_LAUNCH(launch_crop_resize_norm(fd, t->d_state, t->d_slots, 1, 2, kTemplate, t->d_lut, t->patches_z, (size_t)kNTz * kPatchK, t->pz_hi,
                                      t->f16 ? nullptr : t->pz_lo, t->stream));
    if (t->nsplit == 0) {
        GemmArgs g = gemm_args(t->patches_z, kPatchK, t->patch_w, t->patch_b, t->Zemb + (size_t)target * kNTz * t->D, t->D, kNTz, t->D, kPatchK);
        g.pos = t->pos_z;
        g.c_rows_in = kNTz, g.c_rows_stride = kNTz, g.c_row_off = 0;
        VT_LAUNCH(launch_gemm_simt(g, t->stream));
    } else {
        TcGemmPlan p = t->plan_patch_z;
        p.args.batch_off = target;
        VT_LAUNCH(tc_gemm_launch(p, kNTz, t->nsplit, t->stream, false));
        VT_LAUNCH(launch_layernorm_split(t->Zemb + (size_t)target * kNTz * t->D, t->D, t->blk[0].ln1_g, t->blk[0].ln1_b, t->zln_hi + (size_t)target * kNTz * t->D,
                                         t->f16 ? nullptr : t->zln_lo + (size_t)target * kNTz * t->D, kNTz, t->D, 1 << 30, 0, 0, t->stream, false));
    }
    t->kernel_launches += launches;
    VT_CUDA(cudaStreamSynchronize(t->stream));
    if (!t->inited[target]) {
        t->inited[target] = 1;
        t->active.push_back(target);
        std::sort(t->active.begin(), t->active.end());
    }
    t->rect_mirror[target] = box;
    return sync_slots(t);
}

vt_status vt_tracker_drop(vt_tracker* t, int32_t target) {
    if (!t || target < 0 || target >= t->maxT || t->in_flight) return VT_ERR_INVALID;
    VT_CUDA(cudaSetDevice(t->cfg.device));
    if (t->inited[target]) {
        t->inited[target] = 0;
        t->active.erase(std::remove(t->active.begin(), t->active.end(), target), t->active.end());
        VT_CUDA(cudaMemsetAsync(t->d_state + target, 0, sizeof(TargetState), t->stream));
    }
    return sync_slots(t);
}

vt_status vt_tracker_submit(vt_tracker* t, uint8_t* frame, size_t len) {
    if (!t || !frame) return VT_ERR_INVALID;
    VT_CUDA(cudaSetDevice(t->cfg.device));
    return submit_common(t, frame, nullptr, len);
}

vt_status vt_tracker_submit_device(vt_tracker* t, uint8_t* d_frame, size_t len) {
    if (!t || !d_frame) return VT_ERR_INVALID;
    VT_CUDA(cudaSetDevice(t->cfg.device));
    return submit_common(t, nullptr, d_frame, len);
}

vt_status vt_tracker_wait(vt_tracker* t, vt_result* results) {
    if (!t) return VT_ERR_INVALID;
    VT_CUDA(cudaSetDevice(t->cfg.device));
    return wait_common(t, results);
}

vt_status vt_tracker_update(vt_tracker* t, uint8_t* frame, size_t len, vt_result* results) {
    if (!t || !frame) return VT_ERR_INVALID;
    VT_CUDA(cudaSetDevice(t->cfg.device));
    if (t->q_count) {
        set_error("vt_tracker_update: frames are in flight (submit/wait); drain them first");
        return VT_ERR_INVALID;
    }
    vt_status st = submit_common(t, frame, nullptr, len);
    if (st != VT_OK) return st;
    return wait_common(t, results);
}

vt_status vt_tracker_update_device(vt_tracker* t, uint8_t* d_frame, size_t len, vt_result* results) {
    if (!t || !d_frame) return VT_ERR_INVALID;
    VT_CUDA(cudaSetDevice(t->cfg.device));
    if (t->q_count) {
        set_error("vt_tracker_update_device: frames are in flight (submit/wait); drain them first");
        return VT_ERR_INVALID;
    }
    vt_status st = submit_common(t, nullptr, d_frame, len);
    if (st != VT_OK) return st;
    return wait_common(t, results);
}

vt_status vt_tracker_get_rect(vt_tracker* t, int32_t target, vt_bbox* out) {
    if (!t || !out || target < 0 || target >= t->maxT) return VT_ERR_INVALID;
    if (!t->inited[target]) return VT_ERR_NOT_INIT;
    *out = t->rect_mirror[target];
    return VT_OK;
}

vt_status vt_tracker_set_rect(vt_tracker* t, int32_t target, vt_bbox box) {
    if (!t || target < 0 || target >= t->maxT || t->in_flight) return VT_ERR_INVALID;
    if (!t->inited[target]) return VT_ERR_NOT_INIT;
    VT_CUDA(cudaSetDevice(t->cfg.device));
    int32_t r[4] = {box.x, box.y, box.width, box.height};
    VT_CUDA(cudaMemcpyAsync(&t->d_state[target].rect[0], r, sizeof(r), cudaMemcpyHostToDevice, t->stream));
    VT_CUDA(cudaStreamSynchronize(t->stream));
    t->rect_mirror[target] = box;
    return VT_OK;
}

int32_t vt_tracker_model_dim(const vt_tracker* t, int32_t which) {
    if (!t) return 0;
    switch (which) {
        case 0: return t->D;
        case 1: return t->depth;
        case 2: return t->heads;
        case 3: return t->hidden;
        default: return t->head_ch;
    }
}

// patch-major [tokens][768] -> planar CHW blob
static void patches_to_chw(const std::vector<float>& p, int size, float* chw) {
    const int nt = size / 16;
    for (int tok = 0; tok < nt * nt; ++tok)
        for (int c = 0; c < 3; ++c)
            for (int py = 0; py < 16; ++py)
                for (int px = 0; px < 16; ++px)
                    chw[(size_t)c * size * size + (size_t)((tok / nt) * 16 + py) * size + (tok % nt) * 16 + px] =
                        p[(size_t)tok * kPatchK + c * 256 + py * 16 + px];
}

vt_status vt_tracker_debug_read(vt_tracker* t, int32_t target, float* search_blob, float* template_blob, float* conf_win, float* size_map,
                                float* off_map, float* tokens) {
    if (!t || target < 0 || target >= t->maxT || t->in_flight) return VT_ERR_INVALID;
    if (!t->inited[target]) return VT_ERR_NOT_INIT;
    VT_CUDA(cudaSetDevice(t->cfg.device));
    VT_CUDA(cudaStreamSynchronize(t->stream));
    const int bi = (int)(std::find(t->active.begin(), t->active.end(), target) - t->active.begin());
    if (search_blob) {
        std::vector<float> p((size_t)kNTx * kPatchK);
        VT_CUDA(cudaMemcpy(p.data(), t->patches_x + (size_t)bi * kNTx * kPatchK, p.size() * sizeof(float), cudaMemcpyDeviceToHost));
        patches_to_chw(p, kSearch, search_blob);
    }
    if (template_blob) {  // patches_z holds the most recently initialised template
        std::vector<float> p((size_t)kNTz * kPatchK);
        VT_CUDA(cudaMemcpy(p.data(), t->patches_z, p.size() * sizeof(float), cudaMemcpyDeviceToHost));
        patches_to_chw(p, kTemplate, template_blob);
    }
    if (conf_win || size_map || off_map) {
        float m[1280];
        VT_CUDA(cudaMemcpy(m, t->d_maps + (size_t)target * 1280, sizeof(m), cudaMemcpyDeviceToHost));
        if (conf_win) memcpy(conf_win, m, 256 * sizeof(float));
        if (size_map) memcpy(size_map, m + 256, 512 * sizeof(float));
        if (off_map) memcpy(off_map, m + 768, 512 * sizeof(float));
    }
    if (tokens) {  // final-LN search-token features [256, D] placed at rows 64..319; template rows zero
        memset(tokens, 0, sizeof(float) * kNTok * t->D);
        if (t->nsplit) {  // the tensor-core path keeps only the bf16 split of the final LN: recompute the fp32 copy from X
            VT_CUDA(launch_layernorm(t->X, t->D, t->lnf_g, t->lnf_b, t->Yf, t->D, (int)t->active.size() * kNTx, t->D, kNTx, kNTok, kNTz, t->stream));
            VT_CUDA(cudaStreamSynchronize(t->stream));
        }
        VT_CUDA(cudaMemcpy(tokens + (size_t)kNTz * t->D, t->Yf + (size_t)bi * kNTx * t->D, sizeof(float) * kNTx * t->D, cudaMemcpyDeviceToHost));
    }
    return VT_OK;
}

// Device timeline of the kernels launched since the last call (VT_B200_TRACE=1 at create): out[8 i ..] = {kernel id, t_entry,
// t_after_pdl_wait, t_end, 4 kernel-specific marks} in ns; returns the number of records (<= max_records) through *n and resets the counter.
vt_status vt_tracker_debug_trace(vt_tracker* t, unsigned long long* out, int32_t max_records, int32_t* n) {
    if (!t || !out || !n || !t->d_trace) return VT_ERR_INVALID;
    VT_CUDA(cudaSetDevice(t->cfg.device));
    VT_CUDA(cudaStreamSynchronize(t->stream));
    unsigned long long cnt = 0;
    VT_CUDA(cudaMemcpy(&cnt, t->d_trace, 8, cudaMemcpyDeviceToHost));
    const int32_t k = (int32_t)std::min<unsigned long long>(std::min<unsigned long long>(cnt, 2048), (unsigned long long)std::max(max_records, 0));
    if (k > 0) VT_CUDA(cudaMemcpy(out, t->d_trace + 1, (size_t)k * 64, cudaMemcpyDeviceToHost));
    VT_CUDA(cudaMemset(t->d_trace, 0, 8));
    *n = k;
    return VT_OK;
}

// which: 0 embeddings, 1..depth block outputs (needs cfg.debug_capture = 1)
vt_status vt_tracker_debug_tokens(vt_tracker* t, int32_t target, int32_t which, float* out) {
    if (!t || !out || target < 0 || target >= t->maxT || !t->debug_capture || which < 0 || which > t->depth) return VT_ERR_INVALID;
    VT_CUDA(cudaSetDevice(t->cfg.device));
    VT_CUDA(cudaStreamSynchronize(t->stream));
    const int bi = (int)(std::find(t->active.begin(), t->active.end(), target) - t->active.begin());
    VT_CUDA(cudaMemcpy(out, t->d_dbg + ((size_t)which * t->maxT + bi) * kNTok * t->D, sizeof(float) * kNTok * t->D, cudaMemcpyDeviceToHost));
    return VT_OK;
}

// ---- NV12 -> RGB -------------------------------------------------------------------------------
vt_status vt_convert_nv12_rgb(vt_tracker* t, const uint8_t* nv12, size_t len, uint8_t* rgb_out) {
    if (!t || !nv12 || !rgb_out || t->in_flight) return VT_ERR_INVALID;
    VT_CUDA(cudaSetDevice(t->cfg.device));
    const size_t W = t->W, H = t->H, out_bytes = W * H * 3;
    if (len < W * H * 3 / 2) {  // src/nv12_convert.rs:48-50
        memset(rgb_out, 0, out_bytes);
        return VT_OK;
    }
    if (t->fmt != VT_FMT_NV12) {
        set_error("vt_convert_nv12_rgb: handle was created for RGB24 frames");
        return VT_ERR_INVALID;
    }
    if (!t->d_rgb) VT_CUDA(cudaMalloc(&t->d_rgb, out_bytes + 256));
    const size_t n = std::min(len, t->frame_bytes);
    const bool pin_in = is_pinned(nv12), pin_out = is_pinned(rgb_out);
    if (!pin_in) memcpy(t->h_stage, nv12, n);
    VT_CUDA(cudaMemcpyAsync(t->d_frame, pin_in ? nv12 : t->h_stage, n, cudaMemcpyHostToDevice, t->stream));
    cudaError_t e = launch_nv12_to_rgb(t->d_frame, t->frame_bytes, t->d_rgb, out_bytes, t->W, t->H, 1, t->stream);
    if (e != cudaSuccess) {
        set_error("nv12_to_rgb launch failed: %s", cudaGetErrorString(e));
        return VT_ERR_CUDA;
    }
    ++t->kernel_launches;
    if (pin_out) {
        VT_CUDA(cudaMemcpyAsync(rgb_out, t->d_rgb, out_bytes, cudaMemcpyDeviceToHost, t->stream));
        VT_CUDA(cudaStreamSynchronize(t->stream));
    } else {
        VT_CUDA(cudaStreamSynchronize(t->stream));
        VT_CUDA(cudaMemcpy(rgb_out, t->d_rgb, out_bytes, cudaMemcpyDeviceToHost));
    }
    return VT_OK;
}

vt_status vt_convert_nv12_rgb_device(vt_tracker* t, const uint8_t* d_nv12, size_t stride_in, uint8_t* d_rgb, size_t stride_out,
                                     int32_t n_frames) {
    if (!t || !d_nv12 || !d_rgb || n_frames <= 0) return VT_ERR_INVALID;
    VT_CUDA(cudaSetDevice(t->cfg.device));
    cudaError_t e = launch_nv12_to_rgb(d_nv12, stride_in, d_rgb, stride_out, t->W, t->H, n_frames, t->stream);
    if (e != cudaSuccess) {
        set_error("nv12_to_rgb launch failed: %s", cudaGetErrorString(e));
        return VT_ERR_CUDA;
    }
    ++t->kernel_launches;
    return VT_OK;
}

// ---- format steps either side of the RGB probe ---------------------------------------------------------------
static vt_status fmt_scratch(vt_tracker* t, size_t in_bytes, size_t out_bytes) {
    if (in_bytes > t->fmt_in_cap) {
        if (t->d_fmt_in) cudaFree(t->d_fmt_in);
        t->d_fmt_in = nullptr, t->fmt_in_cap = 0;
        VT_CUDA(cudaMalloc(&t->d_fmt_in, in_bytes + 256));
        t->fmt_in_cap = in_bytes;
    }
    if (out_bytes > t->fmt_out_cap) {
        if (t->d_fmt_out) cudaFree(t->d_fmt_out);
        t->d_fmt_out = nullptr, t->fmt_out_cap = 0;
        VT_CUDA(cudaMalloc(&t->d_fmt_out, out_bytes + 256));
        t->fmt_out_cap = out_bytes;
    }
    return VT_OK;
}

vt_status vt_convert_yuy2_rgb_device(vt_tracker* t, const uint8_t* d_yuy2, size_t stride_in, uint8_t* d_rgb, size_t stride_out, int32_t width,
                                     int32_t height, int32_t n_frames) {
    if (!t || !d_yuy2 || !d_rgb || width <= 0 || height <= 0 || n_frames <= 0) return VT_ERR_INVALID;
    VT_CUDA(cudaSetDevice(t->cfg.device));
    VT_CUDA(launch_yuy2_to_rgb(d_yuy2, stride_in, d_rgb, stride_out, width, height, n_frames, t->stream));
    ++t->kernel_launches;
    return VT_OK;
}

vt_status vt_convert_yuy2_rgb(vt_tracker* t, const uint8_t* yuy2, size_t len, int32_t width, int32_t height, uint8_t* rgb_out) {
    if (!t || !yuy2 || !rgb_out || width <= 0 || height <= 0 || t->in_flight) return VT_ERR_INVALID;
    VT_CUDA(cudaSetDevice(t->cfg.device));
    const size_t in_bytes = (((size_t)width * 2 + 3) & ~(size_t)3) * height, out_bytes = (size_t)width * height * 3;
    if (len < in_bytes) {  // short buffer -> black frame, as the NV12 path (src/nv12_convert.rs:48-50)
        memset(rgb_out, 0, out_bytes);
        return VT_OK;
    }
    vt_status st = fmt_scratch(t, in_bytes, out_bytes);
    if (st != VT_OK) return st;
    VT_CUDA(cudaMemcpyAsync(t->d_fmt_in, yuy2, in_bytes, cudaMemcpyHostToDevice, t->stream));
    VT_CUDA(launch_yuy2_to_rgb(t->d_fmt_in, in_bytes, t->d_fmt_out, out_bytes, width, height, 1, t->stream));
    ++t->kernel_launches;
    VT_CUDA(cudaMemcpyAsync(rgb_out, t->d_fmt_out, out_bytes, cudaMemcpyDeviceToHost, t->stream));
    VT_CUDA(cudaStreamSynchronize(t->stream));
    t->h2d_bytes += in_bytes, t->d2h_bytes += out_bytes;
    return VT_OK;
}

vt_status vt_resize_rgb_device(vt_tracker* t, const uint8_t* d_rgb, int32_t sw, int32_t sh, uint8_t* d_out, int32_t dw, int32_t dh) {
    if (!t || !d_rgb || !d_out || sw <= 0 || sh <= 0 || dw <= 0 || dh <= 0) return VT_ERR_INVALID;
    VT_CUDA(cudaSetDevice(t->cfg.device));
    VT_CUDA(launch_resize_rgb_linear(d_rgb, sw, sh, d_out, dw, dh, t->stream));
    ++t->kernel_launches;
    return VT_OK;
}

vt_status vt_resize_rgb(vt_tracker* t, const uint8_t* rgb, int32_t sw, int32_t sh, uint8_t* out, int32_t dw, int32_t dh) {
    if (!t || !rgb || !out || sw <= 0 || sh <= 0 || dw <= 0 || dh <= 0 || t->in_flight) return VT_ERR_INVALID;
    VT_CUDA(cudaSetDevice(t->cfg.device));
    const size_t in_bytes = (size_t)sw * sh * 3, out_bytes = (size_t)dw * dh * 3;
    vt_status st = fmt_scratch(t, in_bytes, out_bytes);
    if (st != VT_OK) return st;
    VT_CUDA(cudaMemcpyAsync(t->d_fmt_in, rgb, in_bytes, cudaMemcpyHostToDevice, t->stream));
    VT_CUDA(launch_resize_rgb_linear(t->d_fmt_in, sw, sh, t->d_fmt_out, dw, dh, t->stream));
    ++t->kernel_launches;
    VT_CUDA(cudaMemcpyAsync(out, t->d_fmt_out, out_bytes, cudaMemcpyDeviceToHost, t->stream));
    VT_CUDA(cudaStreamSynchronize(t->stream));
    t->h2d_bytes += in_bytes, t->d2h_bytes += out_bytes;
    return VT_OK;
}

// exposed for bench.py: stream handle so that device-side timing happens on the launching stream
void* vt_tracker_stream(vt_tracker* t) { return t ? (void*)t->stream : nullptr; }
vt_status vt_tracker_sync(vt_tracker* t) {
    if (!t) return VT_ERR_INVALID;
    VT_CUDA(cudaSetDevice(t->cfg.device));
    VT_CUDA(cudaStreamSynchronize(t->stream));
    return VT_OK;
}

// ---- overlay -------------------------------------------------------------------------------------
int vt_glyph_rows(int ch, uint8_t rows[7]);  // host_state.cpp

static vt_status build_cmds(vt_tracker* t, const vt_overlay_cmd* cmds, int n, std::vector<std::pair<int, int>>& spans) {
    if (n < 0 || n > kMaxCmds) {
        set_error("vt_overlay: at most %d commands", kMaxCmds);
        return VT_ERR_INVALID;
    }
    const long long H = t->H;
    for (int i = 0; i < n; ++i) {
        const vt_overlay_cmd& c = cmds[i];
        OverlayCmdDev& d = t->h_cmds[i];
        memset(&d, 0, sizeof(d));
        d.kind = c.kind, d.x = c.x, d.y = c.y, d.w = c.w, d.h = c.h, d.a = c.a, d.r = c.r, d.g = c.g, d.b = c.b;
        long long r0 = 0, r1 = -1;
        switch (c.kind) {
            case VT_OV_RECT:
                if (!rect_rows(t->fmt, H, c.y, c.h, c.a, r0, r1)) r0 = 0, r1 = -1;
                break;
            case VT_OV_CROSSHAIR:
                if (!cross_rows(H, c.y, c.a, r0, r1)) r0 = 0, r1 = -1;
                break;
            case VT_OV_TEXT: {
                size_t len = strnlen(c.text, sizeof(c.text));
                d.nchar = (uint8_t)len;
                for (size_t k = 0; k < len; ++k) {
                    uint8_t rows[7];
                    if (vt_glyph_rows((unsigned char)c.text[k], rows) == 0) {
                        d.known[k] = 1;
                        memcpy(d.glyph[k], rows, 7);
                    } else if (c.strict_glyphs) {
                        set_error("vt_overlay: no glyph for character 0x%02x", (unsigned char)c.text[k]);
                        return VT_ERR_GLYPH;
                    }
                }
                r0 = c.y, r1 = (long long)c.y + 7LL * std::max(c.a, 0);
                break;
            }
            case VT_OV_BACKGROUND:
                r0 = c.y, r1 = (long long)c.y + c.h;
                if (t->fmt == VT_FMT_RGB24 && (long long)c.y + c.h < 0) r0 = 0, r1 = H - 1;  // (y+bh) as usize wraps, src/drawing_rgb.rs:45
                break;
            case VT_OV_CURSOR: {
                const long long yc = std::max(0LL, std::min<long long>(c.y, H - 1));  // src/drawing.rs:7 clamps, the RGB path does not
                r0 = std::min<long long>(c.y, yc) - 26, r1 = std::max<long long>(c.y, yc) + 26;
                break;
            }
            case VT_OV_SELECTION: r0 = std::min(c.y, c.h), r1 = std::max(c.y, c.h); break;
            default: set_error("vt_overlay: unknown command kind %d", c.kind); return VT_ERR_INVALID;
        }
        if (r1 >= r0 && r1 >= 0 && r0 <= H - 1) spans.emplace_back((int)std::max(0LL, r0), (int)std::min(r1, H - 1));
    }
    return VT_OK;
}

static vt_status overlay_impl(vt_tracker* t, uint8_t* frame, size_t len, const vt_overlay_cmd* cmds, int32_t n, bool upload) {
    if (!t || !frame || (n > 0 && !cmds) || t->in_flight) return VT_ERR_INVALID;
    VT_CUDA(cudaSetDevice(t->cfg.device));
    const size_t need = format_is_luma(t->fmt) ? (size_t)t->W * t->H : 0;
    if (len < need) {
        set_error("vt_overlay: frame shorter than its Y plane");
        return VT_ERR_INVALID;
    }
    std::vector<std::pair<int, int>> spans;
    vt_status st = build_cmds(t, cmds, n, spans);
    if (st != VT_OK) return st;
    if (n == 0) return VT_OK;
    VT_CUDA(cudaEventRecord(t->ev[EV_DEC], t->stream));
    if (upload) {
        const size_t nb = std::min(len, t->frame_bytes);
        if (is_pinned(frame)) {
            VT_CUDA(cudaMemcpyAsync(t->d_frame, frame, nb, cudaMemcpyHostToDevice, t->stream));
        } else {
            memcpy(t->h_stage, frame, nb);
            VT_CUDA(cudaMemcpyAsync(t->d_frame, t->h_stage, nb, cudaMemcpyHostToDevice, t->stream));
        }
    }
    VT_CUDA(cudaMemcpyAsync(t->d_cmds, t->h_cmds, sizeof(OverlayCmdDev) * n, cudaMemcpyHostToDevice, t->stream));
    cudaError_t e = launch_overlay(t->d_frame, std::min(len, t->frame_bytes), t->W, t->H, overlay_format(t->fmt), t->d_cmds, n, t->stream);
    if (e != cudaSuccess) {
        set_error("overlay launch failed: %s", cudaGetErrorString(e));
        return VT_ERR_CUDA;
    }
    ++t->kernel_launches;
    VT_CUDA(cudaEventRecord(t->ev[EV_OVL], t->stream));
    merge_spans(spans);
    bool staged = false;
    st = download_rows(t, frame, len, spans, &staged);
    if (st != VT_OK) return st;
    VT_CUDA(cudaStreamSynchronize(t->stream));
    if (staged) unstage_rows(t, frame, len, spans);
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, t->ev[EV_DEC], t->ev[EV_OVL]) == cudaSuccess) t->last[4] = ms;
    else cudaGetLastError();
    return VT_OK;
}

vt_status vt_overlay(vt_tracker* t, uint8_t* frame, size_t len, const vt_overlay_cmd* cmds, int32_t n) {
    return overlay_impl(t, frame, len, cmds, n, true);
}
vt_status vt_overlay_current(vt_tracker* t, uint8_t* frame, size_t len, const vt_overlay_cmd* cmds, int32_t n) {
    if (t && t->cfg.upload_window) {
        set_error("vt_overlay_current needs the whole frame on the device: create the handle with upload_window = 0");
        return VT_ERR_INVALID;
    }
    return overlay_impl(t, frame, len, cmds, n, false);
}

// ---- timing --------------------------------------------------------------------------------------
vt_status vt_timing_get(vt_tracker* t, vt_timing* o) {
    if (!t || !o) return VT_ERR_INVALID;
    memset(o, 0, sizeof(*o));
    o->fps = t->stats.fps(), o->avg_conv_ms = t->stats.avg_conv_ms(), o->avg_track_ms = t->stats.avg_track_ms();
    o->h2d_ms = t->last[0], o->preprocess_ms = t->last[1], o->vit_ms = t->last[2], o->decode_ms = t->last[3], o->overlay_ms = t->last[4];
    o->d2h_ms = t->last[5], o->total_ms = t->last[6];
    o->avg_h2d_ms = (float)t->r_h2d.mean(), o->avg_preprocess_ms = (float)t->r_pre.mean(), o->avg_vit_ms = (float)t->r_vit.mean();
    o->avg_decode_ms = (float)t->r_dec.mean(), o->avg_overlay_ms = (float)t->r_ovl.mean(), o->avg_d2h_ms = (float)t->r_d2h.mean();
    o->avg_total_ms = (float)t->r_tot.mean();
    o->frames = t->frames, o->kernel_launches = t->kernel_launches;
    o->h2d_bytes = t->h2d_bytes, o->d2h_bytes = t->d2h_bytes;
    return VT_OK;
}
vt_status vt_timing_add_interval(vt_tracker* t, uint64_t us) {
    if (!t) return VT_ERR_INVALID;
    t->stats.add_interval(us);
    return VT_OK;
}
vt_status vt_timing_add_times(vt_tracker* t, uint64_t conv_us, uint64_t track_us) {
    if (!t) return VT_ERR_INVALID;
    t->stats.add_times(conv_us, track_us);
    return VT_OK;
}

}  // extern "C"
