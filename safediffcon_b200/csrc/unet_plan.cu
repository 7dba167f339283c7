// Whole-network executor behind include/safediffcon_b200_plan.h: the inference schedule of Unet2D.forward
// (/root/reference/1D/model/unet.py:382-426) as ONE C entry point.  The handle owns the packed tcgen05 weights, the FiLM table
// for all integer diffusion times and a first-fit layout of every activation inside the caller's workspace; a forward is
// ~120 launches of the kernels declared in safediffcon_b200_unet.h, no allocation, no synchronisation (graph capturable).
//
// Dataflow per level (FP16 mode):
//   ResnetBlock  conv3x3 (+GroupNorm+FiLM+SiLU in its epilogue on the 16x128 level, else a separate in-place sdc_gn_silu)
//                -> conv3x3 (+GroupNorm+SiLU+residual likewise); 1x1 res_conv when the channel count changes
//   LinearAttention  LayerNorm -> qkv 1x1 conv with the q-softmax in its epilogue -> context (softmax_n(k) v^T) -> context folded
//                into a per-sample output projection -> per-sample 1x1 conv -> LayerNorm + residual
//   Downsample2d as a 5-D TMA view (pixel unshuffle) feeding a 1x1 conv; Upsample2d as four 2x2 phase convolutions
//   last block: conv + GroupNorm + SiLU + residual + 1x1 head conv in ONE kernel, NCHW fp32 out.
#include "common.cuh"
#include "../../include/safediffcon_b200_unet.h"
#include "../../include/safediffcon_b200_plan.h"
#include <nvtx3/nvToolsExt.h>
#include <stdlib.h>
#include <map>
#include <string>
#include <vector>

using namespace sdc;

namespace {

constexpr int HID = 128;   // 4 heads x 32
enum { K1 = 0, K3 = 1, KUN = 2, KUP = 3 };

struct ParamDesc { std::string name; int64_t numel; };

struct ConvP {
    int w = -1, b = -1;          // parameter indices (b = -1: no bias)
    int cout = 0, cin = 0, kind = K1;
    bool up = false;             // additionally packed as a fused-upsample (kind 3) weight
    void* packed = nullptr;      // Wp[cout, taps*cin] operand precision
    void* packed_up = nullptr;   // Wp[4*cout, 4*cin]
    float* bias = nullptr;
    float* wt = nullptr;         // backward-data: Wt[cin, taps*cout] TF32 (sdc_pack_conv_weight_dgrad), packed when SDC_UNET_BACKWARD is set
};
struct BlockP {
    int mlp_w, mlp_b, g1w, g1b, g2w, g2b;
    ConvP c1, c2, res;
    bool has_res = false;
    int cin = 0, cout = 0, film_off = 0;
    float *g1[2] = {nullptr, nullptr}, *g2[2] = {nullptr, nullptr};
    std::string name;
};
struct AttnP {
    ConvP qkv, out;
    int g_in_i = -1, g_out_i = -1, dim = 0;
    bool full = false;
    float *g_in = nullptr, *g_out = nullptr, *out_w32 = nullptr;
    void* qkv_ln = nullptr;      // FP16, dim <= 256: qkv weight with the PreNorm gain folded in (sdc_pack_qkv_ln)
    float* qkv_wsum = nullptr;   // its row sums
    std::string name;
};
struct LevelP { BlockP b1, b2; AttnP attn; ConvP resample; bool resamples = false; };

struct ProfEntry { const char* name; cudaEvent_t e0, e1; double bytes, flops; float ms; };

// first-fit allocator over the caller's workspace; deterministic for a given schedule, so a dry run yields the exact size
struct Arena {
    uint8_t* base = nullptr;
    int64_t cap = 0, high = 0;
    bool dry = false;
    std::map<int64_t, int64_t> free_;   // offset -> size
    std::map<int64_t, int64_t> live_;
    void reset(void* b, int64_t c, bool d) {
        base = (uint8_t*)b; cap = c; dry = d; high = 0;
        free_.clear(); live_.clear();
        free_[0] = d ? (int64_t)1 << 60 : c;
    }
    void* alloc(int64_t bytes) {
        bytes = (bytes + 1023) / 1024 * 1024;   // TMA store / swizzle friendly, keeps every buffer 1 KB aligned
        for (auto it = free_.begin(); it != free_.end(); ++it) {
            if (it->second >= bytes) {
                const int64_t off = it->first, rest = it->second - bytes;
                free_.erase(it);
                if (rest > 0) free_[off + bytes] = rest;
                live_[off] = bytes;
                if (off + bytes > high) high = off + bytes;
                return base + off;
            }
        }
        return nullptr;
    }
    void release(void* p) {
        if (!p) return;
        const int64_t off = (uint8_t*)p - base;
        auto it = live_.find(off);
        if (it == live_.end()) return;
        int64_t size = it->second, o = off;
        live_.erase(it);
        auto nx = free_.lower_bound(o);
        if (nx != free_.end() && nx->first == o + size) { size += nx->second; nx = free_.erase(nx); }
        if (nx != free_.begin()) {
            auto pv = std::prev(nx);
            if (pv->first + pv->second == o) { o = pv->first; size += pv->second; free_.erase(pv); }
        }
        free_[o] = size;
    }
};

}  // namespace

struct sdc_unet {
    int dim = 0, channels = 0, out_dim = 0, prec = 0, table_T = 0, init_dim = 0;
    float theta = 10000.f;
    std::vector<int> mults;
    std::vector<ParamDesc> params;
    // architecture
    ConvP stem;            // 7x7 as a 1x1 GEMM over the (high | low) im2col operand
    int stem_kp = 0;
    int time_w1 = -1, time_b1 = -1, time_w2 = -1, time_b2 = -1;
    std::vector<LevelP> downs, ups;
    BlockP mid1, mid2, fin;
    AttnP mid_attn;
    int head_w_i = -1, head_b_i = -1;
    float *head_w = nullptr, *head_b = nullptr;
    int film_total = 0;
    // device storage (one cudaMalloc on the first pack)
    uint8_t* slab = nullptr;
    int64_t slab_bytes = 0;
    std::vector<std::pair<void**, int64_t>> slab_items;   // (where to put the pointer, bytes)
    float *film_w = nullptr, *film_b = nullptr, *table = nullptr, *t_arange = nullptr, *emb = nullptr, *th1 = nullptr, *th2 = nullptr;
    float *tw1 = nullptr, *tb1 = nullptr, *tw2 = nullptr, *tb2 = nullptr, *stem_rep = nullptr;
    float *film_a3 = nullptr, *film_w3 = nullptr;
    int table_rows = 0;
    bool packed = false;
    bool want_bwd = false;   // SDC_UNET_BACKWARD: also pack the data-gradient weights (second slab, +0.56 GB for dim 128)
    uint8_t* dslab = nullptr;
    int64_t dslab_bytes = 0;
    std::vector<std::pair<void**, int64_t>> dslab_items;
    float* stem_wt = nullptr;   // [stem_kp2, dim]
    int stem_kp2 = 0;
    bool bwd_packed = false;
    bool film_tc = true;     // FiLM table GEMM on tcgen05 (TF32, split operands) instead of the fp32 CUDA-core loop
    int fuse_ln = 0;         // 0: separate LayerNorm kernels; 1: both LayerNorms of LinearAttention in the projection epilogues; 2: PreNorm only
    bool fuse_gn = false;    // (value 1 and fuse_gn measured slower than the separate kernels on B200)
    // profile
    bool prof_on = false;
    std::vector<ProfEntry> prof;
    size_t prof_used = 0;
    mutable std::map<int64_t, int64_t> ws_cache;   // (B, H, W) -> workspace bytes
};

namespace {

int add_param(sdc_unet* n, const std::string& name, int64_t numel) {
    n->params.push_back({name, numel});
    return (int)n->params.size() - 1;
}
void want(sdc_unet* n, void** where, int64_t bytes) { n->slab_items.push_back({where, bytes}); }

ConvP make_conv(sdc_unet* n, const std::string& name, int cout, int cin, int kind, bool bias, bool up = false) {
    ConvP c;
    c.cout = cout; c.cin = cin; c.kind = kind; c.up = up;
    const int taps = kind == K3 ? 9 : 1;
    c.w = add_param(n, name + ".weight", (int64_t)cout * cin * taps);
    if (bias) c.b = add_param(n, name + ".bias", cout);
    const int64_t esz = n->prec == SDC_PREC_F16 ? 2 : 4;
    want(n, &c.packed, (int64_t)cout * cin * taps * esz);
    if (up) want(n, &c.packed_up, (int64_t)16 * cout * cin * esz);
    if (bias) want(n, (void**)&c.bias, (int64_t)cout * 4);
    n->dslab_items.push_back({(void**)&c.wt, (int64_t)cout * cin * taps * 4});
    return c;
}
// NOTE: make_conv registers `&c.packed` of a LOCAL; the callers below re-register after the struct has reached its final
// address (fix_conv), so the slab pointers land in the handle's own members.
void fix_conv(sdc_unet* n, ConvP& c, size_t& cursor) {
    n->dslab_items.back().first = (void**)&c.wt;   // (registered by the make_conv call just before)
    n->slab_items[cursor++].first = &c.packed;
    if (c.up) n->slab_items[cursor++].first = &c.packed_up;
    if (c.b >= 0) n->slab_items[cursor++].first = (void**)&c.bias;
}

void build_block(sdc_unet* n, BlockP& b, const std::string& name, int cin, int cout) {
    const int td = 4 * n->dim;
    b.name = name; b.cin = cin; b.cout = cout;
    b.mlp_w = add_param(n, name + ".mlp.1.weight", (int64_t)2 * cout * td);
    b.mlp_b = add_param(n, name + ".mlp.1.bias", 2 * cout);
    size_t cur = n->slab_items.size();
    b.c1 = make_conv(n, name + ".block1.proj", cout, cin, K3, true);
    fix_conv(n, b.c1, cur);
    b.g1w = add_param(n, name + ".block1.norm.weight", cout);
    b.g1b = add_param(n, name + ".block1.norm.bias", cout);
    cur = n->slab_items.size();
    b.c2 = make_conv(n, name + ".block2.proj", cout, cout, K3, true);
    fix_conv(n, b.c2, cur);
    b.g2w = add_param(n, name + ".block2.norm.weight", cout);
    b.g2b = add_param(n, name + ".block2.norm.bias", cout);
    b.has_res = cin != cout;
    if (b.has_res) {
        cur = n->slab_items.size();
        b.res = make_conv(n, name + ".res_conv", cout, cin, K1, true);
        fix_conv(n, b.res, cur);
    }
    for (int k = 0; k < 2; ++k) { want(n, (void**)&b.g1[k], cout * 4); want(n, (void**)&b.g2[k], cout * 4); }
    b.film_off = n->film_total;
    n->film_total += 2 * cout;
}

void build_attn(sdc_unet* n, AttnP& a, const std::string& name, int dim, bool full) {
    a.name = name; a.dim = dim; a.full = full;
    size_t cur = n->slab_items.size();
    a.qkv = make_conv(n, name + ".fn.fn.to_qkv", 3 * HID, dim, K1, false);
    fix_conv(n, a.qkv, cur);
    cur = n->slab_items.size();
    a.out = make_conv(n, name + (full ? ".fn.fn.to_out" : ".fn.fn.to_out.0"), dim, HID, K1, true);
    fix_conv(n, a.out, cur);
    if (!full) {
        a.g_out_i = add_param(n, name + ".fn.fn.to_out.1.g", dim);
        want(n, (void**)&a.g_out, dim * 4);
        want(n, (void**)&a.out_w32, (int64_t)dim * HID * 4);
    }
    a.g_in_i = add_param(n, name + ".fn.norm.g", dim);
    want(n, (void**)&a.g_in, dim * 4);
    if (!full && n->prec == SDC_PREC_F16 && dim <= 256) {
        want(n, &a.qkv_ln, (int64_t)3 * HID * dim * 2);
        want(n, (void**)&a.qkv_wsum, 3 * HID * 4);
    }
}

#define PLAN_CUDA(expr)                                                                      \
    do {                                                                                     \
        cudaError_t e_ = (expr);                                                             \
        if (e_ != cudaSuccess) {                                                             \
            set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_), __FILE__, __LINE__); \
            return SDC_ERR_CUDA;                                                             \
        }                                                                                    \
    } while (0)

__global__ void count_nonfinite_kernel(const float* __restrict__ x, int64_t n, uint32_t* __restrict__ counter) {
    int bad = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        bad += !isfinite(x[i]);
    bad = warp_sum_i(bad);
    if ((threadIdx.x & 31) == 0 && bad) atomicAdd(counter, (uint32_t)bad);
}

// ------------------------------------------------------------------------------------------------ forward context
struct Fwd {
    sdc_unet* n;
    Arena ar;
    bool dry;
    void* stream;
    int B, prec, rc = 0;
    bool f16;
    int64_t esz;
    const float* film = nullptr;   // row 0 of the FiLM rows in use
    int64_t E = 0;
    const int32_t* t_index = nullptr;
    double* stats = nullptr;
    uint8_t* slots = nullptr;
    int stat_i = 0;

    void* alloc(int64_t bytes) {
        void* p = ar.alloc(bytes);
        if (!p && !rc) { set_error("sdc_unet_forward: workspace too small (need more than %lld bytes)", (long long)ar.cap); rc = SDC_ERR_STATE; }
        return p;
    }
    void* opd(int64_t rows, int64_t c) { return alloc(rows * c * esz); }
    void* f32(int64_t rows, int64_t c) { return alloc(rows * c * 4); }
    void release(void* p) { ar.release(p); }

    void begin(const char* name, double bytes, double flops) {
        if (dry || !n->prof_on) return;
        if (n->prof_used == n->prof.size()) {
            ProfEntry e{};
            cudaEventCreate(&e.e0);
            cudaEventCreate(&e.e1);
            n->prof.push_back(e);
        }
        ProfEntry& e = n->prof[n->prof_used];
        e.name = name; e.bytes = bytes; e.flops = flops; e.ms = -1.f;
        cudaEventRecord(e.e0, as_stream(stream));
    }
    void end() {
        if (dry || !n->prof_on) return;
        cudaEventRecord(n->prof[n->prof_used].e1, as_stream(stream));
        ++n->prof_used;
    }
};

#define RUN(name, bytes, flops, call)       \
    do {                                    \
        if (!f.dry && !f.rc) {              \
            f.begin(name, bytes, flops);    \
            const int rc__ = (call);        \
            f.end();                        \
            if (rc__ > 0) f.rc = rc__;      \
        }                                   \
    } while (0)

struct Range {   // NVTX range per network block (visible in nsys / ncu --nvtx)
    bool on;
    Range(const Fwd& f, const std::string& name) : on(!f.dry) { if (on) nvtxRangePushA(name.c_str()); }
    ~Range() { if (on) nvtxRangePop(); }
};

void conv(Fwd& f, int kind, const void* a0, int c0, const void* a1, int c1, const ConvP& cw, const void* wp, const void* residual, void* out,
          double* st, int operand_out, int h, int w, double algo_k = -1.0, int prec_override = -1, int cout_override = -1,
          bool no_bias = false) {
    const int prec = prec_override >= 0 ? prec_override : f.prec;
    const int cout = cout_override >= 0 ? cout_override : cw.cout;
    const float* bias = no_bias ? nullptr : cw.bias;
    const int taps = kind == K3 ? 9 : (kind == K1 ? 1 : 4);
    const double rows = (double)f.B * h * w * (kind == KUP ? 4 : 1);
    const double k = algo_k > 0 ? algo_k : (double)taps * (c0 + c1);
    const double oesz = operand_out ? (double)f.esz : 4.0;
    const double esz = prec == SDC_PREC_F16 ? 2.0 : 4.0;
    const double bytes = (double)f.B * h * w * (kind == KUN ? 4 : 1) * (c0 + c1) * esz + rows * cout * (operand_out ? esz : 4.0) + (double)cout * k * esz +
                         (residual ? rows * cout * esz : 0.0);
    (void)oesz;
    const char* nm = kind == K3 ? "conv3x3" : (kind == K1 ? "conv1x1" : (kind == KUN ? "conv_unshuffle" : "conv_upsample"));
    int rc = -1;
    if (!f.dry && !f.rc) {
        f.begin(nm, bytes, 2.0 * rows * cout * k);
        if (kind == K3 && (w == 128 || w == 64) && cout <= 128)
            rc = sdc_conv3x3_row(prec, a0, c0, a1, c1, wp, bias, residual, out, st, operand_out, f.B, h, w, cout, f.stream);
        if (rc != 0)
            rc = sdc_conv_gemm(prec, kind, a0, c0, a1, c1, wp, bias, residual, out, st, operand_out, f.B, h, w, cout, f.stream);
        f.end();
        if (rc > 0) f.rc = rc;
    }
}

// EXPERIMENTAL (SDC_FUSE_GN=1): conv + GroupNorm in one kernel on the 16x128 level -- correct but slower on B200 (see conv_row.cu)
bool gn_fusable(const Fwd& f, const BlockP& p, int c0, int c1, int h, int w) {
    return f.n->fuse_gn && f.f16 && w == 128 && h % 4 == 0 && h <= 16 && p.cout <= 128 && p.cout % 32 == 0 && c0 % 64 == 0 && c1 % 64 == 0;
}

// ResnetBlock (unet.py:166-180) on one or two concatenated NHWC inputs -> operand [B*h*w, cout] (nullptr when `head_out`
// is given: the block is the network's last one and writes eps[B, out_dim, H, W] directly)
// rowstats (FP16 compact path only): the block's last GroupNorm apply also emits per-row LayerNorm statistics of its output
void* resnet(Fwd& f, const BlockP& p, const void* a0, int c0, const void* a1, int c1, int h, int w, float* head_out = nullptr,
             float* rowstats = nullptr) {
    sdc_unet* n = f.n;
    Range r(f, p.name);
    const int64_t M = (int64_t)f.B * h * w, HW = (int64_t)h * w;
    const int cout = p.cout;
    double* s1 = f.stats ? f.stats + (int64_t)f.stat_i * f.B * 2 : nullptr;
    double* s2 = f.stats ? f.stats + (int64_t)(f.stat_i + 1) * f.B * 2 : nullptr;
    uint8_t* n1 = f.slots ? f.slots + (int64_t)f.stat_i * f.B * SDC_GN_SLOT_BYTES : nullptr;
    uint8_t* n2 = f.slots ? f.slots + (int64_t)(f.stat_i + 1) * f.B * SDC_GN_SLOT_BYTES : nullptr;
    f.stat_i += 2;
    const float* ss = f.film + p.film_off;
    const double e = (double)f.esz;
    if (gn_fusable(f, p, c0, c1, h, w) && (!head_out || n->out_dim <= 4)) {
        void* h1 = f.opd(M, cout);
        RUN("conv3x3_gn", (double)M * (c0 + c1) * e + (double)M * cout * e, 2.0 * M * cout * 9.0 * (c0 + c1),
            sdc_conv3x3_row_gn(a0, c0, a1, c1, p.c1.packed, p.c1.bias, h1, s1, n1, p.g1[0], p.g1[1], ss, f.t_index, f.E, nullptr, f.B, h, w,
                               cout, f.stream));
        const void* res = a0;
        void* resbuf = nullptr;
        if (p.has_res) {
            resbuf = f.opd(M, cout);
            conv(f, K1, a0, c0, a1, c1, p.res, p.res.packed, nullptr, resbuf, nullptr, 1, h, w);
            res = resbuf;
        }
        void* out2 = nullptr;
        if (head_out) {
            if (!f.dry && !f.rc && cudaMemsetAsync(head_out, 0, (size_t)M * n->out_dim * 4, as_stream(f.stream)) != cudaSuccess) f.rc = SDC_ERR_CUDA;
            RUN("conv3x3_gn_head", (double)M * cout * 2.0 * e + (double)M * n->out_dim * 4.0, 2.0 * M * cout * (9.0 * cout + n->out_dim),
                sdc_conv3x3_row_gn_head(h1, cout, nullptr, 0, p.c2.packed, p.c2.bias, s2, n2, p.g2[0], p.g2[1], res, n->head_w, n->head_b,
                                        head_out, n->out_dim, f.B, h, w, cout, f.stream));
        } else {
            out2 = f.opd(M, cout);
            RUN("conv3x3_gn", (double)M * cout * 3.0 * e, 2.0 * M * cout * 9.0 * cout,
                sdc_conv3x3_row_gn(h1, cout, nullptr, 0, p.c2.packed, p.c2.bias, out2, s2, n2, p.g2[0], p.g2[1], nullptr, nullptr, 0, res, f.B,
                                   h, w, cout, f.stream));
        }
        f.release(h1);
        f.release(resbuf);
        return out2;
    }
    if (f.f16 && head_out && cout == 128 && n->out_dim <= 4) {
        // Last block of the network: fp32 norm inputs, and the second normalisation fused with the 1x1 head convolution
        // (sdc_gn_silu_head keeps the activation fp32 in registers).  The last rounding sites in front of the output dominate the
        // eps error (40 % of its variance is made here with fp16 intermediates): 7.4-8.3e-4 -> 6.5-7.4e-4 relative.
        void* raw = f.f32(M, cout);
        conv(f, K3, a0, c0, a1, c1, p.c1, p.c1.packed, nullptr, raw, s1, 0, h, w);
        void* h1 = f.opd(M, cout);
        RUN("gn_silu_f32", (double)M * cout * (4.0 + e), 0.0,
            sdc_gn_silu(f.prec, raw, 0, s1, p.g1[0], p.g1[1], ss, f.t_index, f.E, nullptr, 0, h1, f.B, (int)HW, cout, f.stream));
        conv(f, K3, h1, cout, nullptr, 0, p.c2, p.c2.packed, nullptr, raw, s2, 0, h, w);   // conv1's output is dead: reuse its buffer
        const void* res = a0;
        if (p.has_res) {
            conv(f, K1, a0, c0, a1, c1, p.res, p.res.packed, nullptr, h1, nullptr, 1, h, w);   // h1 is dead after conv2
            res = h1;
        }
        RUN("gn_silu_head", (double)M * cout * (4.0 + e) + (double)M * n->out_dim * 4.0, 2.0 * M * cout * n->out_dim,
            sdc_gn_silu_head((const float*)raw, s2, p.g2[0], p.g2[1], res, 1, n->head_w, n->head_b, head_out, f.B, (int)HW, cout, n->out_dim,
                             f.stream));
        f.release(raw);
        f.release(h1);
        return nullptr;
    }
    if (f.f16) {
        // compact intermediates: fp16 conv outputs (statistics from the fp32 accumulators), GroupNorm in place
        void* raw = f.opd(M, cout);
        conv(f, K3, a0, c0, a1, c1, p.c1, p.c1.packed, nullptr, raw, s1, 1, h, w);
        RUN("gn_silu", (double)M * cout * 2.0 * e, 0.0,
            sdc_gn_silu(f.prec, raw, 1, s1, p.g1[0], p.g1[1], ss, f.t_index, f.E, nullptr, 0, raw, f.B, (int)HW, cout, f.stream));
        void* raw2 = f.opd(M, cout);
        conv(f, K3, raw, cout, nullptr, 0, p.c2, p.c2.packed, nullptr, raw2, s2, 1, h, w);
        const void* res = a0;
        if (p.has_res) {
            conv(f, K1, a0, c0, a1, c1, p.res, p.res.packed, nullptr, raw, nullptr, 1, h, w);   // conv1's buffer is dead after conv2
            res = raw;
        }
        if (head_out) {
            // last block outside the fused path (other image sizes): GroupNorm apply, then the plain head convolution
            RUN("gn_silu", (double)M * cout * 3.0 * e, 0.0,
                sdc_gn_silu(f.prec, raw2, 1, s2, p.g2[0], p.g2[1], nullptr, nullptr, 0, res, 1, raw2, f.B, (int)HW, cout, f.stream));
            RUN("head_conv1", (double)M * cout * e + (double)M * n->out_dim * 4.0, 2.0 * M * cout * n->out_dim,
                sdc_head_conv1(f.prec, raw2, n->head_w, n->head_b, head_out, f.B, (int)HW, cout, n->out_dim, f.stream));
            f.release(raw);
            f.release(raw2);
            return nullptr;
        }
        if (rowstats)
            RUN("gn_silu", (double)M * cout * 3.0 * e, 0.0,
                sdc_gn_silu_rowstats(raw2, s2, p.g2[0], p.g2[1], nullptr, nullptr, 0, res, raw2, rowstats, f.B, (int)HW, cout, f.stream));
        else
            RUN("gn_silu", (double)M * cout * 3.0 * e, 0.0,
                sdc_gn_silu(f.prec, raw2, 1, s2, p.g2[0], p.g2[1], nullptr, nullptr, 0, res, 1, raw2, f.B, (int)HW, cout, f.stream));
        f.release(raw);
        return raw2;
    }
    // TF32: fp32 conv outputs, GroupNorm in place (the normalised tensor is TF32-rounded by the kernel)
    void* raw = f.f32(M, cout);
    conv(f, K3, a0, c0, a1, c1, p.c1, p.c1.packed, nullptr, raw, s1, 0, h, w);
    RUN("gn_silu", (double)M * cout * 8.0, 0.0,
        sdc_gn_silu(f.prec, raw, 0, s1, p.g1[0], p.g1[1], ss, f.t_index, f.E, nullptr, 0, raw, f.B, (int)HW, cout, f.stream));
    void* raw2 = f.f32(M, cout);
    conv(f, K3, raw, cout, nullptr, 0, p.c2, p.c2.packed, nullptr, raw2, s2, 0, h, w);
    const void* res = a0;
    int res_operand = 1;
    if (p.has_res) {
        conv(f, K1, a0, c0, a1, c1, p.res, p.res.packed, nullptr, raw, nullptr, 0, h, w);
        res = raw;
        res_operand = 0;
    }
    RUN("gn_silu", (double)M * cout * 12.0, 0.0,
        sdc_gn_silu(f.prec, raw2, 0, s2, p.g2[0], p.g2[1], nullptr, nullptr, 0, res, res_operand, raw2, f.B, (int)HW, cout, f.stream));
    f.release(raw);
    if (head_out) {
        RUN("head_conv1", (double)M * cout * 4.0 + (double)M * n->out_dim * 4.0, 2.0 * M * cout * n->out_dim,
            sdc_head_conv1(f.prec, raw2, n->head_w, n->head_b, head_out, f.B, (int)HW, cout, n->out_dim, f.stream));
        f.release(raw2);
        return nullptr;
    }
    return raw2;
}

// Residual(PreNorm(LinearAttention | Attention)) (unet.py:16-22,65-76,182-258); returns a new operand [B*h*w, c]
bool ln_fusable(const Fwd& f, const AttnP& p, int h, int w) {
    return f.n->fuse_ln && f.f16 && !p.full && p.qkv_ln && (h * w) % 128 == 0 && (p.dim == 128 || p.dim == 256) && !f.n->fuse_gn;
}

void* attention(Fwd& f, const AttnP& p, const void* xin, int c, int h, int w, const float* rowstats = nullptr) {
    Range r(f, p.name);
    const int64_t M = (int64_t)f.B * h * w;
    const int n = h * w;
    const double e = (double)f.esz;
    if (rowstats) {
        // FP16, <= 256 channels: both LayerNorms live in the convolution epilogues (see safediffcon_b200_unet.h)
        void* qs = f.opd(M, HID);
        void* kv = f.opd(M, 2 * HID);
        RUN("conv1x1_qkv", (double)M * (c + 3.0 * HID) * e + (double)M * 8.0, 2.0 * M * 3.0 * HID * c,
            sdc_conv1x1_qkv_ln(xin, c, p.qkv_ln, p.qkv_wsum, rowstats, qs, kv, f.B, h, w, HID, f.stream));
        void* ws = f.alloc(sdc_linear_attention_workspace(f.B));
        RUN("linattn_context", (double)M * 2.0 * HID * e, 2.0 * M * HID * 32.0,
            sdc_linear_attention_context(kv, (const uint8_t*)kv + HID * f.esz, 2 * HID, 1, ws, f.B, n, f.stream));
        f.release(kv);
        void* wf = f.opd((int64_t)f.B * c, HID);
        RUN("linattn_fold", (double)f.B * c * HID * e, 2.0 * f.B * c * HID * 32.0,
            sdc_linear_attention_fold(f.prec, ws, p.out_w32, wf, f.B, c, f.stream));
        if (f.n->fuse_ln == 2) {
            // PreNorm folded into the qkv projection only; output projection + LayerNorm + residual as separate kernels
            void* proj = f.opd(M, c);
            RUN("conv1x1_per_sample", (double)M * HID * e + (double)M * c * 2.0 + (double)f.B * c * HID * e, 2.0 * M * c * HID,
                sdc_conv1x1_per_sample(f.prec, qs, HID, wf, p.out.bias, proj, 1, f.B, h, w, c, f.stream));
            void* out = f.opd(M, c);
            RUN("layernorm", (double)M * c * (2.0 + 2.0 * e), 0.0, sdc_channel_layernorm(f.prec, proj, 1, p.g_out, xin, out, M, c, 1, f.stream));
            f.release(qs); f.release(ws); f.release(wf); f.release(proj);
            return out;
        }
        void* out = f.opd(M, c);
        RUN("conv1x1_per_sample_ln", (double)M * (HID + 2.0 * c) * e + (double)f.B * c * HID * e, 2.0 * M * c * HID,
            sdc_conv1x1_per_sample_ln(qs, HID, wf, p.out.bias, p.g_out, xin, out, f.B, h, w, c, f.stream));
        f.release(qs); f.release(ws); f.release(wf);
        return out;
    }
    void* xn = f.opd(M, c);
    RUN("layernorm", (double)M * c * 2.0 * e, 0.0, sdc_channel_layernorm(f.prec, xin, 1, p.g_in, nullptr, xn, M, c, 1, f.stream));
    if (!p.full && n % 128 == 0) {
        void* qs = f.opd(M, HID);
        void* kv = f.opd(M, 2 * HID);
        RUN("conv1x1_qkv", (double)M * (c + 3.0 * HID) * e, 2.0 * M * 3.0 * HID * c,
            sdc_conv1x1_qkv(f.prec, xn, c, p.qkv.packed, qs, kv, f.f16 ? 1 : 0, f.B, h, w, HID, f.stream));
        void* ws = f.alloc(sdc_linear_attention_workspace(f.B));
        RUN("linattn_context", (double)M * 2.0 * HID * e, 2.0 * M * HID * 32.0,
            sdc_linear_attention_context(kv, (const uint8_t*)kv + HID * f.esz, 2 * HID, f.f16 ? 1 : 0, ws, f.B, n, f.stream));
        void* wf = f.opd((int64_t)f.B * c, HID);
        RUN("linattn_fold", (double)f.B * c * HID * e, 2.0 * f.B * c * HID * 32.0,
            sdc_linear_attention_fold(f.prec, ws, p.out_w32, wf, f.B, c, f.stream));
        void* proj = f.f16 ? f.opd(M, c) : f.f32(M, c);
        RUN("conv1x1_per_sample", (double)M * HID * e + (double)M * c * (f.f16 ? 2.0 : 4.0) + (double)f.B * c * HID * e, 2.0 * M * c * HID,
            sdc_conv1x1_per_sample(f.prec, qs, HID, wf, p.out.bias, proj, f.f16 ? 1 : 0, f.B, h, w, c, f.stream));
        RUN("layernorm", (double)M * c * ((f.f16 ? 2.0 : 4.0) + 2.0 * e), 0.0,
            sdc_channel_layernorm(f.prec, proj, f.f16 ? 1 : 0, p.g_out, xin, xn, M, c, 1, f.stream));
        f.release(qs); f.release(kv); f.release(ws); f.release(wf); f.release(proj);
        return xn;
    }
    void* qkv = f.f32(M, 3 * HID);
    ConvP nb = p.qkv;
    conv(f, K1, xn, c, nullptr, 0, nb, p.qkv.packed, nullptr, qkv, nullptr, 0, h, w);
    void* att = f.opd(M, HID);
    if (p.full) {
        RUN("attention", (double)M * (3.0 * HID * 4.0 + HID * e), 4.0 * M * n * HID, sdc_attention(f.prec, (const float*)qkv, att, f.B, n, f.stream));
        conv(f, K1, att, HID, nullptr, 0, p.out, p.out.packed, xin, xn, nullptr, 1, h, w);
    } else {
        void* ws = f.alloc(sdc_linear_attention_workspace(f.B));
        RUN("linear_attention", (double)M * (3.0 * HID * 4.0 + HID * e), 4.0 * M * HID * 32.0,
            sdc_linear_attention(f.prec, (const float*)qkv, att, ws, f.B, n, f.stream));
        void* proj = f.f32(M, c);
        conv(f, K1, att, HID, nullptr, 0, p.out, p.out.packed, nullptr, proj, nullptr, 0, h, w);
        RUN("layernorm", (double)M * c * (4.0 + 2.0 * e), 0.0, sdc_channel_layernorm(f.prec, proj, 0, p.g_out, xin, xn, M, c, 1, f.stream));
        f.release(ws);
        f.release(proj);
    }
    f.release(qkv);
    f.release(att);
    return xn;
}

int run_forward(sdc_unet* n, Fwd& f, const float* x, float* eps, int H, int W, uint32_t* nonfinite) {
    const int B = f.B;
    const int n_gn = 2 * (int)(2 * n->downs.size() + 2 + 2 * n->ups.size() + 1);
    f.stats = (double*)f.alloc((int64_t)n_gn * B * 2 * sizeof(double));
    f.slots = f.f16 ? (uint8_t*)f.alloc((int64_t)n_gn * B * SDC_GN_SLOT_BYTES) : nullptr;
    if (f.rc) return f.rc;
    if (!f.dry) {
        PLAN_CUDA(cudaMemsetAsync(f.stats, 0, (size_t)n_gn * B * 2 * sizeof(double), as_stream(f.stream)));
        if (f.slots) PLAN_CUDA(cudaMemsetAsync(f.slots, 0xFF, (size_t)n_gn * B * SDC_GN_SLOT_BYTES, as_stream(f.stream)));
    }
    int c = n->init_dim, h = H, w = W;
    void* cur = f.opd((int64_t)B * H * W, c);
    {
        Range r(f, "init_conv");
        static const bool stem_tc = []() { const char* e = getenv("SDC_STEM_TC"); return !(e && e[0] == '0'); }();
        int rc = -1;
        if (f.f16 && stem_tc && W == 128 && n->init_dim == 128 && n->stem_kp == 320 && n->channels <= 3) {
            rc = 0;   // (eligibility mirrors sdc_stem_conv7_tc, so that the dry run lays out the same buffers)
            RUN("stem_conv7", (double)B * H * W * (n->channels * 4.0 + c * 2.0), 2.0 * B * H * W * c * n->channels * 49.0,
                sdc_stem_conv7_tc(x, n->stem.packed, n->stem.bias, cur, B, n->channels, H, W, c, n->stem_kp, f.stream));
        }
        if (rc != 0) {
            void* patches = f.opd((int64_t)B * H * W, n->stem_kp);
            RUN("stem_im2col", (double)B * H * W * (n->channels * 4.0 + n->stem_kp * (double)f.esz), 0.0,
                sdc_stem_im2col(f.prec, x, patches, B, n->channels, H, W, n->stem_kp, f.stream));
            conv(f, K1, patches, n->stem_kp, nullptr, 0, n->stem, n->stem.packed, nullptr, cur, nullptr, 1, H, W, n->channels * 49.0);
            f.release(patches);
        }
    }
    void* r0 = cur;
    const int r_c = c;
    std::vector<std::pair<void*, int>> skips;
    for (size_t li = 0; li < n->downs.size(); ++li) {
        const LevelP& L = n->downs[li];
        void* a = resnet(f, L.b1, cur, c, nullptr, 0, h, w);
        if (cur != r0) f.release(cur);
        skips.push_back({a, c});
        float* rs = (ln_fusable(f, L.attn, h, w) && !gn_fusable(f, L.b2, c, 0, h, w)) ? (float*)f.alloc((int64_t)B * h * w * 8) : nullptr;
        void* b = resnet(f, L.b2, a, c, nullptr, 0, h, w, nullptr, rs);
        void* at = attention(f, L.attn, b, c, h, w, rs);
        f.release(b);
        f.release(rs);
        skips.push_back({at, c});
        const int cout = L.resample.cout;
        Range r(f, "downs." + std::to_string(li) + ".3");
        if (L.resamples) { h /= 2; w /= 2; }
        void* nxt = f.opd((int64_t)B * h * w, cout);
        conv(f, L.resamples ? KUN : K3, at, c, nullptr, 0, L.resample, L.resample.packed, nullptr, nxt, nullptr, 1, h, w);
        cur = nxt;
        c = cout;
    }
    {
        void* a = resnet(f, n->mid1, cur, c, nullptr, 0, h, w);
        f.release(cur);
        void* at = attention(f, n->mid_attn, a, c, h, w);
        f.release(a);
        cur = resnet(f, n->mid2, at, c, nullptr, 0, h, w);
        f.release(at);
    }
    for (size_t li = 0; li < n->ups.size(); ++li) {
        const LevelP& L = n->ups[li];
        auto s = skips.back(); skips.pop_back();
        void* a = resnet(f, L.b1, cur, c, s.first, s.second, h, w);
        f.release(cur);
        f.release(s.first);
        c = L.b1.cout;
        s = skips.back(); skips.pop_back();
        float* rs = (ln_fusable(f, L.attn, h, w) && !gn_fusable(f, L.b2, c, s.second, h, w)) ? (float*)f.alloc((int64_t)B * h * w * 8) : nullptr;
        void* b = resnet(f, L.b2, a, c, s.first, s.second, h, w, nullptr, rs);
        f.release(a);
        f.release(s.first);
        void* at = attention(f, L.attn, b, c, h, w, rs);
        f.release(b);
        f.release(rs);
        const int cout = L.resample.cout;
        Range r(f, "ups." + std::to_string(li) + ".3");
        void* nxt;
        if (L.resamples && (w == 16 || w % 32 == 0)) {
            nxt = f.opd((int64_t)B * 4 * h * w, cout);
            conv(f, KUP, at, c, nullptr, 0, L.resample, L.resample.packed_up, nullptr, nxt, nullptr, 1, h, w);
            h *= 2; w *= 2;
        } else {
            void* in = at;
            void* upb = nullptr;
            if (L.resamples) {
                upb = f.opd((int64_t)B * 4 * h * w, c);
                RUN("upsample2x", (double)B * 5.0 * h * w * c * f.esz, 0.0, sdc_upsample2x(f.prec, at, upb, B, h, w, c, f.stream));
                h *= 2; w *= 2;
                in = upb;
            }
            nxt = f.opd((int64_t)B * h * w, cout);
            conv(f, K3, in, c, nullptr, 0, L.resample, L.resample.packed, nullptr, nxt, nullptr, 1, h, w);
            f.release(upb);
        }
        f.release(at);
        cur = nxt;
        c = cout;
    }
    if (h != H || w != W) { set_error("sdc_unet_forward: internal size mismatch"); return SDC_ERR_STATE; }
    resnet(f, n->fin, cur, c, r0, r_c, h, w, eps);
    f.release(cur);
    f.release(r0);
    if (nonfinite && !f.dry && !f.rc) {
        const int64_t ne = (int64_t)B * n->out_dim * H * W;
        f.begin("count_nonfinite", ne * 4.0, 0.0);
        count_nonfinite_kernel<<<296, 256, 0, as_stream(f.stream)>>>(eps, ne, nonfinite);
        g_launches.fetch_add(1);
        f.end();
        PLAN_CUDA(cudaGetLastError());
    }
    return f.rc;
}


// ================================================================================================ backward-data pass
// d<eps, g>/dx (VJP with respect to the denoiser input; the reference reaches it through autograd when a guidance callable
// differentiates eps_theta(x_t, t), /root/reference/1D/model/diffusion.py:254-262).  run_record = the forward schedule with every
// normalisation input kept (fp32 convolution outputs, separate attention kernels, unfused upsample); run_backward walks the
// records in reverse: GroupNorm / LayerNorm / attention backward kernels (csrc/unet_bwd.cu) and the SAME tcgen05 convolution
// kernels in TF32 with the transposed, tap-flipped weights of sdc_pack_conv_weight_dgrad.  Gradients are fp32 containers,
// TF32-rounded where they feed a data-gradient convolution.
struct Rec {
    int kind;                       // 0 resnet, 1 attn, 2 push, 3 down, 4 up
    const BlockP* blk = nullptr;
    const AttnP* att = nullptr;
    const ConvP* cv = nullptr;
    int c0 = 0, c1 = 0, c = 0, h = 0, w = 0;
    bool flag = false;              // down: unshuffle; up: upsample
    void *raw1 = nullptr, *raw2 = nullptr, *xin = nullptr, *qkv = nullptr, *proj = nullptr, *ws = nullptr;
    double *s1 = nullptr, *s2 = nullptr;
    const float* ss = nullptr;
};

void* resnet_rec(Fwd& f, const BlockP& p, const void* a0, int c0, const void* a1, int c1, int h, int w, std::vector<Rec>& tape) {
    Range r(f, p.name);
    const int64_t M = (int64_t)f.B * h * w, HW = (int64_t)h * w;
    const int cout = p.cout;
    double* s1 = f.stats + (int64_t)f.stat_i * f.B * 2;
    double* s2 = f.stats + (int64_t)(f.stat_i + 1) * f.B * 2;
    f.stat_i += 2;
    const float* ss = f.film + p.film_off;
    void* raw1 = f.f32(M, cout);
    conv(f, K3, a0, c0, a1, c1, p.c1, p.c1.packed, nullptr, raw1, s1, 0, h, w);
    void* h1 = f.opd(M, cout);
    RUN("gn_silu", (double)M * cout * (4.0 + f.esz), 0.0,
        sdc_gn_silu(f.prec, raw1, 0, s1, p.g1[0], p.g1[1], ss, f.t_index, f.E, nullptr, 0, h1, f.B, (int)HW, cout, f.stream));
    void* raw2 = f.f32(M, cout);
    conv(f, K3, h1, cout, nullptr, 0, p.c2, p.c2.packed, nullptr, raw2, s2, 0, h, w);
    f.release(h1);
    const void* res = a0;
    void* resbuf = nullptr;
    int res_operand = 1;
    if (p.has_res) {
        resbuf = f.f32(M, cout);
        conv(f, K1, a0, c0, a1, c1, p.res, p.res.packed, nullptr, resbuf, nullptr, 0, h, w);
        res = resbuf;
        res_operand = 0;
    }
    void* out = f.opd(M, cout);
    RUN("gn_silu", (double)M * cout * (8.0 + f.esz), 0.0,
        sdc_gn_silu(f.prec, raw2, 0, s2, p.g2[0], p.g2[1], nullptr, nullptr, 0, res, res_operand, out, f.B, (int)HW, cout, f.stream));
    f.release(resbuf);
    Rec rc;
    rc.kind = 0; rc.blk = &p; rc.c0 = c0; rc.c1 = c1; rc.h = h; rc.w = w; rc.raw1 = raw1; rc.raw2 = raw2; rc.s1 = s1; rc.s2 = s2; rc.ss = ss;
    tape.push_back(rc);
    return out;
}

void* attention_rec(Fwd& f, const AttnP& p, void* xin, int c, int h, int w, std::vector<Rec>& tape) {
    Range r(f, p.name);
    const int64_t M = (int64_t)f.B * h * w;
    const int n = h * w;
    void* xn = f.opd(M, c);
    RUN("layernorm", (double)M * c * 2.0 * f.esz, 0.0, sdc_channel_layernorm(f.prec, xin, 1, p.g_in, nullptr, xn, M, c, 1, f.stream));
    void* qkv = f.f32(M, 3 * HID);
    conv(f, K1, xn, c, nullptr, 0, p.qkv, p.qkv.packed, nullptr, qkv, nullptr, 0, h, w);
    f.release(xn);
    void* att = f.opd(M, HID);
    void* out = f.opd(M, c);
    Rec rc;
    rc.kind = 1; rc.att = &p; rc.c = c; rc.h = h; rc.w = w; rc.xin = xin; rc.qkv = qkv;
    if (p.full) {
        RUN("attention", (double)M * HID * 16.0, 4.0 * M * n * HID, sdc_attention(f.prec, (const float*)qkv, att, f.B, n, f.stream));
        conv(f, K1, att, HID, nullptr, 0, p.out, p.out.packed, xin, out, nullptr, 1, h, w);
    } else {
        void* ws = f.alloc(sdc_linear_attention_workspace(f.B));
        RUN("linear_attention", (double)M * HID * 16.0, 4.0 * M * HID * 32.0, sdc_linear_attention(f.prec, (const float*)qkv, att, ws, f.B, n, f.stream));
        void* proj = f.f32(M, c);
        conv(f, K1, att, HID, nullptr, 0, p.out, p.out.packed, nullptr, proj, nullptr, 0, h, w);
        RUN("layernorm", (double)M * c * (4.0 + 2.0 * f.esz), 0.0, sdc_channel_layernorm(f.prec, proj, 0, p.g_out, xin, out, M, c, 1, f.stream));
        rc.ws = ws; rc.proj = proj;
    }
    f.release(att);
    tape.push_back(rc);
    return out;
}

// data gradient of convolution `cw` restricted to its input channels [lo, hi): a convolution of g with rows [lo, hi) of Wt
void* dgrad(Fwd& f, int kind, const void* g, int cin_g, const ConvP& cw, int lo, int hi, const void* residual, int operand_out, int h, int w) {
    const int taps = kind == K3 ? 9 : 1;
    const float* wt = cw.wt + (int64_t)lo * taps * cin_g;
    void* out = f.f32((int64_t)f.B * h * w, hi - lo);
    conv(f, kind, g, cin_g, nullptr, 0, cw, wt, residual, out, nullptr, operand_out, h, w, -1.0, SDC_PREC_TF32, hi - lo, true);
    return out;
}

void* gn_bwd(Fwd& f, const void* dy, const void* raw, const double* stats, float* const* gb, const float* ss, double* sums, int hw, int c) {
    void* dx = f.f32((int64_t)f.B * hw, c);
    RUN("gn_silu_bwd", (double)f.B * hw * c * 16.0, 0.0,
        sdc_gn_silu_bwd((const float*)dy, (const float*)raw, stats, gb[0], gb[1], ss, ss ? f.t_index : nullptr, ss ? f.E : 0, sums, (float*)dx, f.B,
                        hw, c, f.stream));
    return dx;
}

int run_vjp(sdc_unet* n, Fwd& f, const float* x, const float* g_eps, float* eps, float* gx, int H, int W) {
    const int B = f.B;
    const int n_gn = 2 * (int)(2 * n->downs.size() + 2 + 2 * n->ups.size() + 1);
    f.stats = (double*)f.alloc((int64_t)n_gn * B * 2 * sizeof(double));
    double* sums = (double*)f.alloc((int64_t)B * 2 * sizeof(double));
    if (f.rc) return f.rc;
    if (!f.dry) PLAN_CUDA(cudaMemsetAsync(f.stats, 0, (size_t)n_gn * B * 2 * sizeof(double), as_stream(f.stream)));
    std::vector<Rec> tape;
    // ---------------- forward with records ----------------
    int c = n->init_dim, h = H, w = W;
    void* cur = f.opd((int64_t)B * H * W, c);
    {
        void* patches = f.opd((int64_t)B * H * W, n->stem_kp);
        RUN("stem_im2col", 0.0, 0.0, sdc_stem_im2col(f.prec, x, patches, B, n->channels, H, W, n->stem_kp, f.stream));
        conv(f, K1, patches, n->stem_kp, nullptr, 0, n->stem, n->stem.packed, nullptr, cur, nullptr, 1, H, W, n->channels * 49.0);
        f.release(patches);
    }
    void* r0 = cur;
    const int r_c = c;
    std::vector<std::pair<void*, int>> skips;
    Rec push; push.kind = 2;
    for (size_t li = 0; li < n->downs.size(); ++li) {
        const LevelP& L = n->downs[li];
        void* a = resnet_rec(f, L.b1, cur, c, nullptr, 0, h, w, tape);
        skips.push_back({a, c});
        tape.push_back(push);
        void* b = resnet_rec(f, L.b2, a, c, nullptr, 0, h, w, tape);
        void* at = attention_rec(f, L.attn, b, c, h, w, tape);
        skips.push_back({at, c});
        tape.push_back(push);
        const int cout = L.resample.cout;
        if (L.resamples) { h /= 2; w /= 2; }
        void* nxt = f.opd((int64_t)B * h * w, cout);
        conv(f, L.resamples ? KUN : K3, at, c, nullptr, 0, L.resample, L.resample.packed, nullptr, nxt, nullptr, 1, h, w);
        Rec d; d.kind = 3; d.cv = &L.resample; d.flag = L.resamples; d.c = c; d.h = h; d.w = w;
        tape.push_back(d);
        cur = nxt;
        c = cout;
    }
    cur = resnet_rec(f, n->mid1, cur, c, nullptr, 0, h, w, tape);
    cur = attention_rec(f, n->mid_attn, cur, c, h, w, tape);
    cur = resnet_rec(f, n->mid2, cur, c, nullptr, 0, h, w, tape);
    for (size_t li = 0; li < n->ups.size(); ++li) {
        const LevelP& L = n->ups[li];
        auto s = skips.back(); skips.pop_back();
        cur = resnet_rec(f, L.b1, cur, c, s.first, s.second, h, w, tape);
        c = L.b1.cout;
        s = skips.back(); skips.pop_back();
        cur = resnet_rec(f, L.b2, cur, c, s.first, s.second, h, w, tape);
        cur = attention_rec(f, L.attn, cur, c, h, w, tape);
        const int cout = L.resample.cout;
        void* in = cur;
        if (L.resamples) {
            void* upb = f.opd((int64_t)B * 4 * h * w, c);
            RUN("upsample2x", 0.0, 0.0, sdc_upsample2x(f.prec, cur, upb, B, h, w, c, f.stream));
            h *= 2; w *= 2;
            in = upb;
        }
        void* nxt = f.opd((int64_t)B * h * w, cout);
        conv(f, K3, in, c, nullptr, 0, L.resample, L.resample.packed, nullptr, nxt, nullptr, 1, h, w);
        Rec u; u.kind = 4; u.cv = &L.resample; u.flag = L.resamples; u.c = c; u.h = h; u.w = w;
        tape.push_back(u);
        cur = nxt;
        c = cout;
    }
    cur = resnet_rec(f, n->fin, cur, c, r0, r_c, h, w, tape);
    const int cfin = n->fin.cout;
    RUN("head_conv1", 0.0, 0.0, sdc_head_conv1(f.prec, cur, n->head_w, n->head_b, eps, B, H * W, cfin, n->out_dim, f.stream));
    // ---------------- backward walk ----------------
    auto resnet_bwd = [&](const Rec& r, void* g, void*& g0, void*& g1) {
        const BlockP& p = *r.blk;
        const int cout = p.cout, hw = r.h * r.w;
        void* d_raw2 = gn_bwd(f, g, r.raw2, r.s2, p.g2, nullptr, sums, hw, cout);
        f.release(r.raw2);
        void* d_h1 = dgrad(f, K3, d_raw2, cout, p.c2, 0, cout, nullptr, 0, r.h, r.w);
        f.release(d_raw2);
        void* d_raw1 = gn_bwd(f, d_h1, r.raw1, r.s1, p.g1, r.ss, sums, hw, cout);
        f.release(d_h1);
        f.release(r.raw1);
        void* outs[2] = {nullptr, nullptr};
        const int lohi[2][2] = {{0, r.c0}, {r.c0, r.c0 + r.c1}};
        for (int sgm = 0; sgm < 2; ++sgm) {
            const int lo = lohi[sgm][0], hi = lohi[sgm][1];
            if (hi == lo) continue;
            void* side = g;
            void* side_buf = nullptr;
            if (p.has_res) { side_buf = dgrad(f, K1, g, cout, p.res, lo, hi, nullptr, 0, r.h, r.w); side = side_buf; }
            outs[sgm] = dgrad(f, K3, d_raw1, cout, p.c1, lo, hi, side, 1, r.h, r.w);
            f.release(side_buf);
        }
        f.release(d_raw1);
        g0 = outs[0];
        g1 = outs[1];
    };
    auto attn_bwd = [&](const Rec& r, void* g) -> void* {
        const AttnP& p = *r.att;
        const int cc = r.c, nn = r.h * r.w;
        const int64_t M = (int64_t)B * nn;
        void* d_qkv = f.f32(M, 3 * HID);
        if (p.full) {
            void* d_att = dgrad(f, K1, g, cc, p.out, 0, HID, nullptr, 0, r.h, r.w);
            RUN("attention_bwd", 0.0, 0.0, sdc_attention_bwd((const float*)r.qkv, (const float*)d_att, (float*)d_qkv, B, nn, f.stream));
            f.release(d_att);
        } else {
            void* d_proj = f.f32(M, cc);
            RUN("layernorm_bwd", 0.0, 0.0,
                sdc_channel_layernorm_bwd((const float*)g, r.proj, 0, p.g_out, nullptr, (float*)d_proj, M, cc, 1, f.stream));
            void* d_att = dgrad(f, K1, d_proj, cc, p.out, 0, HID, nullptr, 0, r.h, r.w);
            f.release(d_proj);
            void* wsb = f.alloc(sdc_linear_attention_bwd_workspace(B));
            RUN("linear_attention_bwd", 0.0, 0.0,
                sdc_linear_attention_bwd((const float*)r.qkv, (const float*)d_att, r.ws, wsb, (float*)d_qkv, B, nn, f.stream));
            f.release(d_att); f.release(wsb); f.release(r.ws); f.release(r.proj);
        }
        f.release(r.qkv);
        void* d_xn = dgrad(f, K1, d_qkv, 3 * HID, p.qkv, 0, cc, nullptr, 0, r.h, r.w);
        f.release(d_qkv);
        void* d_x = f.f32(M, cc);
        RUN("layernorm_bwd", 0.0, 0.0,
            sdc_channel_layernorm_bwd((const float*)d_xn, r.xin, f.f16 ? 1 : 0, p.g_in, (const float*)g, (float*)d_x, M, cc, 1, f.stream));
        f.release(d_xn);
        return d_x;
    };
    auto add_into = [&](void* a, const void* b, int64_t count) {
        RUN("add_inplace", 0.0, 0.0, sdc_add_inplace((float*)a, (const float*)b, count, 1, f.stream));
    };
    int i = (int)tape.size() - 1;
    void* g = f.f32((int64_t)B * H * W, cfin);
    RUN("head_conv1_bwd", 0.0, 0.0, sdc_head_conv1_bwd(g_eps, n->head_w, (float*)g, B, H * W, cfin, n->out_dim, 1, f.stream));
    void *g_main = nullptr, *g_r = nullptr;
    resnet_bwd(tape[i], g, g_main, g_r);
    f.release(g);
    g = g_main;
    --i;
    std::vector<void*> skip_grads;
    int gc = n->ups.empty() ? cfin : tape[i].cv ? tape[i].cv->cout : cfin;   // channel count of g (tracked below)
    (void)gc;
    for (; i >= 0; --i) {
        const Rec& r = tape[i];
        if (r.kind == 4) {          // up
            const ConvP& cw = *r.cv;
            if (r.flag) {
                void* g_hi = dgrad(f, K3, g, cw.cout, cw, 0, r.c, nullptr, 0, r.h, r.w);
                f.release(g);
                g = f.f32((int64_t)B * (r.h / 2) * (r.w / 2), r.c);
                RUN("upsample2x_bwd", 0.0, 0.0, sdc_upsample2x_bwd((const float*)g_hi, (float*)g, B, r.h / 2, r.w / 2, r.c, 1, f.stream));
                f.release(g_hi);
            } else {
                void* g2 = dgrad(f, K3, g, cw.cout, cw, 0, r.c, nullptr, 1, r.h, r.w);
                f.release(g);
                g = g2;
            }
        } else if (r.kind == 1) {   // attention
            void* g2 = attn_bwd(r, g);
            f.release(g);
            g = g2;
        } else if (r.kind == 0) {   // resnet
            void *a = nullptr, *b = nullptr;
            resnet_bwd(r, g, a, b);
            f.release(g);
            g = a;
            if (b) skip_grads.push_back(b);
        } else if (r.kind == 3) {   // down: the tensor entering it was also pushed as a skip
            const ConvP& cw = *r.cv;
            void* sg = skip_grads.back(); skip_grads.pop_back();
            if (r.flag) {
                void* t = dgrad(f, K1, g, cw.cout, cw, 0, 4 * r.c, nullptr, 0, r.h, r.w);
                f.release(g);
                g = f.f32((int64_t)B * 4 * r.h * r.w, r.c);
                RUN("pixel_shuffle_bwd", 0.0, 0.0, sdc_pixel_shuffle_bwd((const float*)t, (const float*)sg, (float*)g, B, r.h, r.w, r.c, 1, f.stream));
                f.release(t);
            } else {
                void* g2 = dgrad(f, K3, g, cw.cout, cw, 0, r.c, sg, 1, r.h, r.w);
                f.release(g);
                g = g2;
            }
            f.release(sg);
        } else if (r.kind == 2) {   // push: consumed by the following "down" record, else add the skip gradient here
            if (!(i + 1 < (int)tape.size() && tape[i + 1].kind == 3)) {
                void* sg = skip_grads.back(); skip_grads.pop_back();
                const Rec& prev = tape[i - 1];   // the resnet whose output was pushed: its geometry gives the element count
                add_into(g, sg, (int64_t)B * prev.h * prev.w * prev.blk->cout);
                f.release(sg);
            }
        }
    }
    if (!skip_grads.empty()) { set_error("sdc_unet_backward_data: internal skip-gradient mismatch"); return SDC_ERR_STATE; }
    add_into(g, g_r, (int64_t)B * H * W * n->init_dim);   // the stem output also feeds final_res_block (r = x.clone(), unet.py:393)
    f.release(g_r);
    void* t = f.f32((int64_t)B * H * W, n->stem_kp2);
    {
        ConvP st = n->stem;
        if (!f.dry && !f.rc) {
            f.begin("conv1x1", 0.0, 0.0);
            const int rc2 = sdc_conv_gemm(SDC_PREC_TF32, K1, g, n->init_dim, nullptr, 0, n->stem_wt, nullptr, nullptr, t, nullptr, 0, B, H, W, n->stem_kp2, f.stream);
            f.end();
            if (rc2) f.rc = rc2;
        }
    }
    RUN("stem_col2im", 0.0, 0.0, sdc_stem_col2im((const float*)t, gx, B, n->channels, H, W, n->stem_kp2, f.stream));
    return f.rc;
}

}  // namespace

// ------------------------------------------------------------------------------------------------ C ABI
extern "C" int sdc_unet_create(sdc_unet** out, int dim, const int* dim_mults, int n_mults, int channels, int out_dim, int prec,
                               float theta, int table_timesteps) {
    SDC_REQUIRE(out && dim_mults && n_mults >= 1 && n_mults <= 8, "sdc_unet_create: bad arguments");
    SDC_REQUIRE(prec == SDC_PREC_F16 || prec == SDC_PREC_TF32, "sdc_unet_create: precision %d", prec);
    SDC_REQUIRE(dim > 0 && dim % (prec == SDC_PREC_F16 ? 64 : 32) == 0, "sdc_unet_create: dim %d must be a multiple of %d", dim,
                prec == SDC_PREC_F16 ? 64 : 32);
    SDC_REQUIRE(channels >= 1 && channels <= 8 && out_dim >= 1 && out_dim <= 8 && table_timesteps >= 1, "sdc_unet_create: channels / out_dim / table");
    sdc_unet* n = new sdc_unet();
    n->dim = dim; n->init_dim = dim; n->channels = channels; n->out_dim = out_dim; n->prec = prec; n->theta = theta; n->table_T = table_timesteps;
    n->mults.assign(dim_mults, dim_mults + n_mults);
    const int td = 4 * dim;
    // parameter order = Unet2D.named_parameters() (safediffcon_b200/unet.py; keys = the reference's state_dict keys)
    n->time_w1 = add_param(n, "time_mlp.1.weight", (int64_t)td * dim);
    n->time_b1 = add_param(n, "time_mlp.1.bias", td);
    n->time_w2 = add_param(n, "time_mlp.3.weight", (int64_t)td * td);
    n->time_b2 = add_param(n, "time_mlp.3.bias", td);
    // stem: W[c, Cin*49] repeated at columns 0 and kp/2 of a [c, kp] matrix (sdc_stem_im2col's high | low operand)
    const int k_stem = channels * 49, kh = (k_stem + 31) / 32 * 32;
    n->stem_kp = (2 * kh) % 64 == 0 ? 2 * kh : 2 * (kh + 32);
    n->stem.cout = dim; n->stem.cin = n->stem_kp; n->stem.kind = K1;
    n->stem.w = add_param(n, "init_conv.weight", (int64_t)dim * channels * 49);
    n->stem.b = add_param(n, "init_conv.bias", dim);
    const int64_t esz = prec == SDC_PREC_F16 ? 2 : 4;
    want(n, &n->stem.packed, (int64_t)dim * n->stem_kp * esz);
    want(n, (void**)&n->stem.bias, dim * 4);
    want(n, (void**)&n->stem_rep, (int64_t)dim * n->stem_kp * 4);
    n->stem_kp2 = (k_stem + 63) / 64 * 64;
    n->dslab_items.push_back({(void**)&n->stem_wt, (int64_t)n->stem_kp2 * dim * 4});
    std::vector<int> dims{dim};
    for (int m : n->mults) dims.push_back(dim * m);
    const int nl = n_mults;
    n->downs.resize(nl);
    n->ups.resize(nl);
    for (int i = 0; i < nl; ++i) {
        const int d_in = dims[i], d_out = dims[i + 1];
        const bool last = i == nl - 1;
        LevelP& L = n->downs[i];
        const std::string base = "downs." + std::to_string(i);
        build_block(n, L.b1, base + ".0", d_in, d_in);
        build_block(n, L.b2, base + ".1", d_in, d_in);
        build_attn(n, L.attn, base + ".2", d_in, false);
        size_t cur = n->slab_items.size();
        L.resamples = !last;
        L.resample = last ? make_conv(n, base + ".3", d_out, d_in, K3, true) : make_conv(n, base + ".3.1", d_out, 4 * d_in, KUN, true);
        fix_conv(n, L.resample, cur);
    }
    const int mid = dims.back();
    build_block(n, n->mid1, "mid_block1", mid, mid);
    build_attn(n, n->mid_attn, "mid_attn", mid, true);
    build_block(n, n->mid2, "mid_block2", mid, mid);
    for (int i = 0; i < nl; ++i) {
        const int d_in = dims[nl - 1 - i], d_out = dims[nl - i];
        const bool last = i == nl - 1;
        LevelP& L = n->ups[i];
        const std::string base = "ups." + std::to_string(i);
        build_block(n, L.b1, base + ".0", d_out + d_in, d_out);
        build_block(n, L.b2, base + ".1", d_out + d_in, d_out);
        build_attn(n, L.attn, base + ".2", d_out, false);
        size_t cur = n->slab_items.size();
        L.resamples = !last;
        L.resample = make_conv(n, base + (last ? ".3" : ".3.1"), d_in, d_out, K3, true, !last);
        fix_conv(n, L.resample, cur);
    }
    build_block(n, n->fin, "final_res_block", 2 * dim, dim);
    n->head_w_i = add_param(n, "final_conv.weight", (int64_t)out_dim * dim);
    n->head_b_i = add_param(n, "final_conv.bias", out_dim);
    want(n, (void**)&n->head_w, (int64_t)out_dim * dim * 4);
    want(n, (void**)&n->head_b, out_dim * 4);
    // time MLP + FiLM projections + table
    want(n, (void**)&n->tw1, (int64_t)td * dim * 4);
    want(n, (void**)&n->tb1, td * 4);
    want(n, (void**)&n->tw2, (int64_t)td * td * 4);
    want(n, (void**)&n->tb2, td * 4);
    want(n, (void**)&n->film_w, (int64_t)n->film_total * td * 4);
    want(n, (void**)&n->film_b, (int64_t)n->film_total * 4);
    // FiLM table on the tensor cores (TF32, operands split into high + low parts): table rows padded to a multiple of 128
    n->table_rows = (table_timesteps + 127) / 128 * 128;
    want(n, (void**)&n->table, (int64_t)n->table_rows * n->film_total * 4);
    want(n, (void**)&n->film_a3, (int64_t)n->table_rows * 3 * td * 4);
    want(n, (void**)&n->film_w3, (int64_t)n->film_total * 3 * td * 4);
    want(n, (void**)&n->t_arange, (int64_t)table_timesteps * 4);
    want(n, (void**)&n->emb, (int64_t)table_timesteps * dim * 4);
    want(n, (void**)&n->th1, (int64_t)table_timesteps * td * 4);
    want(n, (void**)&n->th2, (int64_t)table_timesteps * td * 4);
    int64_t total = 0;
    for (auto& it : n->slab_items) total += (it.second + 255) / 256 * 256;
    n->slab_bytes = total;
    total = 0;
    for (auto& it : n->dslab_items) total += (it.second + 255) / 256 * 256;
    n->dslab_bytes = total;
    *out = n;
    return SDC_OK;
}

extern "C" void sdc_unet_destroy(sdc_unet* n) {
    if (!n) return;
    if (n->slab) cudaFree(n->slab);
    if (n->dslab) cudaFree(n->dslab);
    for (auto& e : n->prof) { cudaEventDestroy(e.e0); cudaEventDestroy(e.e1); }
    delete n;
}

extern "C" int sdc_unet_param_count(const sdc_unet* n) { return n ? (int)n->params.size() : 0; }
extern "C" const char* sdc_unet_param_name(const sdc_unet* n, int i) {
    return (n && i >= 0 && i < (int)n->params.size()) ? n->params[i].name.c_str() : nullptr;
}
extern "C" int64_t sdc_unet_param_numel(const sdc_unet* n, int i) {
    return (n && i >= 0 && i < (int)n->params.size()) ? n->params[i].numel : -1;
}

namespace {
// out[r, :] = (hi | lo | hi) (w_mode = 0, activations) or (hi | hi | lo) (w_mode = 1, weights) of act(x[r, :]), hi = tf32(x),
// lo = tf32(x - hi): [a_hi | a_lo | a_hi] . [w_hi | w_hi | w_lo]^T = a_hi w_hi + a_lo w_hi + a_hi w_lo, i.e. the fp32 product to
// ~2^-21 on TF32 tensor cores.  Rows >= R are zero.  silu: exact expf (this runs once per weight version).
__global__ void split3_kernel(const float* __restrict__ x, float* __restrict__ out, int R, int rows_out, int K, int w_mode, int silu_in) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)rows_out * K) return;
    const int r = (int)(i / K), k = (int)(i - (int64_t)r * K);
    float v = 0.f;
    if (r < R) {
        v = x[(int64_t)r * K + k];
        if (silu_in) v = v / (1.0f + expf(-v));
    }
    const float hi = to_tf32(v), lo = to_tf32(v - hi);
    float* o = out + (int64_t)r * 3 * K + k;
    o[0] = hi;
    o[K] = w_mode ? hi : lo;
    o[2 * K] = w_mode ? lo : hi;
}

__global__ void arange_kernel(float* t, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) t[i] = (float)i;
}

int copy_param(sdc_unet* n, float* dst, const float* const* params, int idx, cudaStream_t st) {
    PLAN_CUDA(cudaMemcpyAsync(dst, params[idx], (size_t)n->params[idx].numel * 4, cudaMemcpyDeviceToDevice, st));
    return SDC_OK;
}
int pack_conv(sdc_unet* n, ConvP& c, const float* const* params, void* stream) {
    int rc = sdc_pack_conv_weight(n->prec, c.kind == KUN ? KUN : c.kind, params[c.w], c.packed, c.cout, c.cin, stream);
    if (rc) return rc;
    if (n->want_bwd && (rc = sdc_pack_conv_weight_dgrad(c.kind == K3 ? 1 : (c.kind == KUN ? 2 : 0), params[c.w], c.wt, c.cout, c.cin, stream))) return rc;
    if (c.up) { rc = sdc_pack_conv_weight(n->prec, KUP, params[c.w], c.packed_up, c.cout, c.cin, stream); if (rc) return rc; }
    if (c.b >= 0) return copy_param(n, c.bias, params, c.b, as_stream(stream));
    return SDC_OK;
}
int pack_block(sdc_unet* n, BlockP& b, const float* const* params, void* stream) {
    cudaStream_t st = as_stream(stream);
    int rc;
    if ((rc = pack_conv(n, b.c1, params, stream))) return rc;
    if ((rc = pack_conv(n, b.c2, params, stream))) return rc;
    if (b.has_res && (rc = pack_conv(n, b.res, params, stream))) return rc;
    if ((rc = copy_param(n, b.g1[0], params, b.g1w, st))) return rc;
    if ((rc = copy_param(n, b.g1[1], params, b.g1b, st))) return rc;
    if ((rc = copy_param(n, b.g2[0], params, b.g2w, st))) return rc;
    if ((rc = copy_param(n, b.g2[1], params, b.g2b, st))) return rc;
    const int td = 4 * n->dim;
    if ((rc = copy_param(n, n->film_w + (int64_t)b.film_off * td, params, b.mlp_w, st))) return rc;
    return copy_param(n, n->film_b + b.film_off, params, b.mlp_b, st);
}
int pack_attn(sdc_unet* n, AttnP& a, const float* const* params, void* stream) {
    cudaStream_t st = as_stream(stream);
    int rc;
    if ((rc = pack_conv(n, a.qkv, params, stream))) return rc;
    if ((rc = pack_conv(n, a.out, params, stream))) return rc;
    if ((rc = copy_param(n, a.g_in, params, a.g_in_i, st))) return rc;
    if (!a.full) {
        if ((rc = copy_param(n, a.g_out, params, a.g_out_i, st))) return rc;
        if ((rc = copy_param(n, a.out_w32, params, a.out.w, st))) return rc;
    }
    if (a.qkv_ln && (rc = sdc_pack_qkv_ln(params[a.qkv.w], params[a.g_in_i], a.qkv_ln, a.qkv_wsum, 3 * HID, a.dim, stream))) return rc;
    return SDC_OK;
}
}  // namespace

extern "C" int sdc_unet_pack_weights(sdc_unet* n, const float* const* params, int n_params, void* stream) {
    SDC_REQUIRE(n && params, "sdc_unet_pack_weights: null arguments");
    SDC_REQUIRE(n_params == (int)n->params.size(), "sdc_unet_pack_weights: expected %d parameter pointers, got %d", (int)n->params.size(), n_params);
    for (int i = 0; i < n_params; ++i) SDC_REQUIRE(params[i], "sdc_unet_pack_weights: parameter %s is null", n->params[i].name.c_str());
    cudaStream_t st = as_stream(stream);
    if (!n->slab) {
        PLAN_CUDA(cudaMalloc((void**)&n->slab, (size_t)n->slab_bytes));
        int64_t off = 0;
        for (auto& it : n->slab_items) { *it.first = n->slab + off; off += (it.second + 255) / 256 * 256; }
    }
    if (n->want_bwd && !n->dslab) {
        PLAN_CUDA(cudaMalloc((void**)&n->dslab, (size_t)n->dslab_bytes));
        int64_t off = 0;
        for (auto& it : n->dslab_items) { *it.first = n->dslab + off; off += (it.second + 255) / 256 * 256; }
    }
    int rc;
    if (n->want_bwd) {
        // stem data gradient: dY[M, c] x W[c, Cin*49] as a 1x1 convolution onto kp2 (zero padded) columns
        PLAN_CUDA(cudaMemsetAsync(n->stem_wt, 0, (size_t)n->stem_kp2 * n->dim * 4, st));
        if ((rc = sdc_pack_conv_weight_dgrad(0, params[n->stem.w], n->stem_wt, n->dim, n->channels * 49, stream))) return rc;
    }
    n->bwd_packed = n->want_bwd;
    // stem: replicate the 7x7 weight into the (high | low) column ranges, then pack as a 1x1 GEMM operand
    {
        const int c = n->dim, k = n->channels * 49, kp = n->stem_kp;
        PLAN_CUDA(cudaMemsetAsync(n->stem_rep, 0, (size_t)c * kp * 4, st));
        PLAN_CUDA(cudaMemcpy2DAsync(n->stem_rep, (size_t)kp * 4, params[n->stem.w], (size_t)k * 4, (size_t)k * 4, c, cudaMemcpyDeviceToDevice, st));
        PLAN_CUDA(cudaMemcpy2DAsync(n->stem_rep + kp / 2, (size_t)kp * 4, params[n->stem.w], (size_t)k * 4, (size_t)k * 4, c, cudaMemcpyDeviceToDevice, st));
        if ((rc = sdc_pack_conv_weight(n->prec, K1, n->stem_rep, n->stem.packed, c, kp, stream))) return rc;
        if ((rc = copy_param(n, n->stem.bias, params, n->stem.b, st))) return rc;
    }
    for (auto& L : n->downs) {
        if ((rc = pack_block(n, L.b1, params, stream))) return rc;
        if ((rc = pack_block(n, L.b2, params, stream))) return rc;
        if ((rc = pack_attn(n, L.attn, params, stream))) return rc;
        if ((rc = pack_conv(n, L.resample, params, stream))) return rc;
    }
    if ((rc = pack_block(n, n->mid1, params, stream))) return rc;
    if ((rc = pack_attn(n, n->mid_attn, params, stream))) return rc;
    if ((rc = pack_block(n, n->mid2, params, stream))) return rc;
    for (auto& L : n->ups) {
        if ((rc = pack_block(n, L.b1, params, stream))) return rc;
        if ((rc = pack_block(n, L.b2, params, stream))) return rc;
        if ((rc = pack_attn(n, L.attn, params, stream))) return rc;
        if ((rc = pack_conv(n, L.resample, params, stream))) return rc;
    }
    if ((rc = pack_block(n, n->fin, params, stream))) return rc;
    if ((rc = copy_param(n, n->head_w, params, n->head_w_i, st))) return rc;
    if ((rc = copy_param(n, n->head_b, params, n->head_b_i, st))) return rc;
    if ((rc = copy_param(n, n->tw1, params, n->time_w1, st))) return rc;
    if ((rc = copy_param(n, n->tb1, params, n->time_b1, st))) return rc;
    if ((rc = copy_param(n, n->tw2, params, n->time_w2, st))) return rc;
    if ((rc = copy_param(n, n->tb2, params, n->time_b2, st))) return rc;
    // FiLM table for every integer time: sinusoid -> Linear -> GELU -> Linear (time_mlp), then every block's SiLU -> Linear
    const int T = n->table_T, td = 4 * n->dim;
    arange_kernel<<<(T + 255) / 256, 256, 0, st>>>(n->t_arange, T);
    g_launches.fetch_add(1);
    PLAN_CUDA(cudaGetLastError());
    if ((rc = sdc_sinusoidal_embedding(n->t_arange, n->emb, T, n->dim, n->theta, stream))) return rc;
    if ((rc = sdc_linear_rows(n->emb, n->tw1, n->tb1, n->th1, T, n->dim, td, 0, stream))) return rc;
    if ((rc = sdc_linear_rows(n->th1, n->tw2, n->tb2, n->th2, T, td, td, 2, stream))) return rc;
    // every block's SiLU -> Linear for all T times at once: [T, 512] x [512, E] = 8.3 GFLOP for dim 128 -- 9 ms as an fp32 CUDA-core
    // loop, ~0.2 ms as ONE tcgen05 TF32 GEMM over split operands (paid on every optimiser / EMA step of a fine-tuning loop)
    if (n->film_tc && td % 32 == 0 && n->film_total % 32 == 0) {
        const int64_t na = (int64_t)n->table_rows * td, nw = (int64_t)n->film_total * td;
        split3_kernel<<<(unsigned)((na + 255) / 256), 256, 0, st>>>(n->th2, n->film_a3, T, n->table_rows, td, 0, 1);
        split3_kernel<<<(unsigned)((nw + 255) / 256), 256, 0, st>>>(n->film_w, n->film_w3, n->film_total, n->film_total, td, 1, 0);
        g_launches.fetch_add(2);
        PLAN_CUDA(cudaGetLastError());
        if ((rc = sdc_conv_gemm(SDC_PREC_TF32, K1, n->film_a3, 3 * td, nullptr, 0, n->film_w3, n->film_b, nullptr, n->table, nullptr, 0,
                                n->table_rows / 128, 1, 128, n->film_total, stream))) return rc;
    } else if ((rc = sdc_linear_rows(n->th2, n->film_w, n->film_b, n->table, T, td, n->film_total, 1, stream))) return rc;
    n->packed = true;
    return SDC_OK;
}

extern "C" int64_t sdc_unet_workspace_bytes(const sdc_unet* n, int B, int H, int W) {
    if (!n || B <= 0 || H <= 0 || W <= 0) return 0;
    const int64_t key = ((int64_t)B << 32) | ((int64_t)H << 16) | W;
    auto it = n->ws_cache.find(key);
    if (it != n->ws_cache.end()) return it->second;
    Fwd f{};
    f.n = const_cast<sdc_unet*>(n); f.dry = true; f.stream = nullptr; f.B = B; f.prec = n->prec; f.f16 = n->prec == SDC_PREC_F16;
    f.esz = f.f16 ? 2 : 4;
    f.ar.reset((void*)(uintptr_t)4096, 0, true);
    float dummy_film = 0.f;
    f.film = &dummy_film;
    if (run_forward(f.n, f, nullptr, (float*)(uintptr_t)4096, H, W, nullptr)) return 0;
    n->ws_cache[key] = f.ar.high;
    return f.ar.high;
}

extern "C" int sdc_unet_forward(sdc_unet* n, const float* x, const int32_t* t_index, int t_uniform, float* eps, int B, int H, int W,
                                void* workspace, int64_t workspace_bytes, uint32_t* nonfinite, void* stream) {
    SDC_REQUIRE(n && x && eps && workspace, "sdc_unet_forward: null arguments");
    if (!n->packed) { set_error("sdc_unet_forward: sdc_unet_pack_weights has not been called"); return SDC_ERR_STATE; }
    const int down = 1 << ((int)n->mults.size() - 1);
    SDC_REQUIRE(B > 0 && H % down == 0 && W % down == 0 && 128 % W == 0 && (H / down) * (W / down) % 32 == 0,
                "sdc_unet_forward: unsupported image size %d x %d (need W | 128 and %d | H, W)", H, W, down);
    SDC_REQUIRE(t_index || (t_uniform >= 0 && t_uniform < n->table_T), "sdc_unet_forward: diffusion time %d outside the FiLM table [0, %d)",
                t_uniform, n->table_T);
    SDC_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 1023) == 0, "sdc_unet_forward: workspace must be 1024-byte aligned");
    const int64_t need = sdc_unet_workspace_bytes(n, B, H, W);
    if (workspace_bytes < need) {
        set_error("sdc_unet_forward: workspace of %lld bytes, need %lld", (long long)workspace_bytes, (long long)need);
        return SDC_ERR_STATE;
    }
    Fwd f{};
    f.n = n; f.dry = false; f.stream = stream; f.B = B; f.prec = n->prec; f.f16 = n->prec == SDC_PREC_F16;
    f.esz = f.f16 ? 2 : 4;
    f.ar.reset(workspace, workspace_bytes, false);
    f.E = n->film_total;
    f.t_index = t_index;
    f.film = t_index ? n->table : n->table + (int64_t)t_uniform * n->film_total;
    if (n->prof_on) n->prof_used = 0;
    nvtxRangePushA("sdc_unet_forward");
    const int rc = run_forward(n, f, x, eps, H, W, nonfinite);
    nvtxRangePop();
    return rc;
}

static int64_t vjp_workspace(const sdc_unet* n, int B, int H, int W) {
    Fwd f{};
    f.n = const_cast<sdc_unet*>(n); f.dry = true; f.stream = nullptr; f.B = B; f.prec = n->prec; f.f16 = n->prec == SDC_PREC_F16;
    f.esz = f.f16 ? 2 : 4;
    f.ar.reset((void*)(uintptr_t)4096, 0, true);
    float dummy_film = 0.f;
    f.film = &dummy_film;
    if (run_vjp(f.n, f, nullptr, nullptr, (float*)(uintptr_t)4096, (float*)(uintptr_t)4096, H, W)) return 0;
    return f.ar.high;
}

extern "C" int64_t sdc_unet_backward_workspace_bytes(const sdc_unet* n, int B, int H, int W) {
    if (!n || B <= 0 || H <= 0 || W <= 0) return 0;
    return vjp_workspace(n, B, H, W);
}

extern "C" int sdc_unet_backward_data(sdc_unet* n, const float* x, const int32_t* t_index, int t_uniform, const float* grad_eps, float* eps,
                                      float* grad_x, int B, int H, int W, void* workspace, int64_t workspace_bytes, void* stream) {
    SDC_REQUIRE(n && x && grad_eps && eps && grad_x && workspace, "sdc_unet_backward_data: null arguments");
    if (!n->packed || !n->bwd_packed) {
        set_error("sdc_unet_backward_data: set SDC_UNET_BACKWARD and call sdc_unet_pack_weights first (the data-gradient weights are not packed)");
        return SDC_ERR_STATE;
    }
    const int down = 1 << ((int)n->mults.size() - 1);
    SDC_REQUIRE(B > 0 && H % down == 0 && W % down == 0 && 128 % W == 0 && (H / down) * (W / down) % 32 == 0,
                "sdc_unet_backward_data: unsupported image size %d x %d", H, W);
    SDC_REQUIRE(t_index || (t_uniform >= 0 && t_uniform < n->table_T), "sdc_unet_backward_data: diffusion time %d outside the FiLM table", t_uniform);
    SDC_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 1023) == 0, "sdc_unet_backward_data: workspace must be 1024-byte aligned");
    const int64_t need = vjp_workspace(n, B, H, W);
    if (workspace_bytes < need) {
        set_error("sdc_unet_backward_data: workspace of %lld bytes, need %lld", (long long)workspace_bytes, (long long)need);
        return SDC_ERR_STATE;
    }
    Fwd f{};
    f.n = n; f.dry = false; f.stream = stream; f.B = B; f.prec = n->prec; f.f16 = n->prec == SDC_PREC_F16;
    f.esz = f.f16 ? 2 : 4;
    f.ar.reset(workspace, workspace_bytes, false);
    f.E = n->film_total;
    f.t_index = t_index;
    f.film = t_index ? n->table : n->table + (int64_t)t_uniform * n->film_total;
    if (n->prof_on) n->prof_used = 0;
    nvtxRangePushA("sdc_unet_backward_data");
    const int rc = run_vjp(n, f, x, grad_eps, eps, grad_x, H, W);
    nvtxRangePop();
    return rc;
}

extern "C" int sdc_unet_film_table(const sdc_unet* n, float* out, int* rows, int* cols, void* stream) {
    SDC_REQUIRE(n, "sdc_unet_film_table: null handle");
    if (rows) *rows = n->table_T;
    if (cols) *cols = n->film_total;
    if (out) {
        SDC_REQUIRE(n->packed, "sdc_unet_film_table: weights not packed");
        PLAN_CUDA(cudaMemcpyAsync(out, n->table, (size_t)n->table_T * n->film_total * 4, cudaMemcpyDeviceToDevice, as_stream(stream)));
    }
    return SDC_OK;
}

extern "C" int sdc_unet_set_flag(sdc_unet* n, int flag, int value) {
    SDC_REQUIRE(n && flag >= SDC_UNET_FUSE_LN && flag <= SDC_UNET_BACKWARD, "sdc_unet_set_flag: unknown flag %d", flag);
    if (flag == SDC_UNET_BACKWARD) { n->want_bwd = value != 0; return SDC_OK; }   // takes effect at the next sdc_unet_pack_weights
    if (flag == SDC_UNET_FUSE_LN) n->fuse_ln = value < 0 ? 0 : (value > 2 ? 2 : value);
    else if (flag == SDC_UNET_FUSE_GN) n->fuse_gn = value != 0;
    else n->film_tc = value != 0;   // takes effect at the next sdc_unet_pack_weights
    n->ws_cache.clear();   // the activation layout depends on the schedule
    return SDC_OK;
}

extern "C" int sdc_unet_profile_enable(sdc_unet* n, int enable) {
    SDC_REQUIRE(n, "sdc_unet_profile_enable: null handle");
    n->prof_on = enable != 0;
    n->prof_used = 0;
    return SDC_OK;
}
extern "C" int sdc_unet_profile_count(const sdc_unet* n) { return n ? (int)n->prof_used : 0; }
extern "C" int sdc_unet_profile_entry(const sdc_unet* n, int i, const char** name, float* ms, double* bytes, double* flops) {
    SDC_REQUIRE(n && i >= 0 && i < (int)n->prof_used, "sdc_unet_profile_entry: index %d out of range", i);
    const ProfEntry& e = n->prof[i];
    float t = 0.f;
    PLAN_CUDA(cudaEventElapsedTime(&t, e.e0, e.e1));
    if (name) *name = e.name;
    if (ms) *ms = t;
    if (bytes) *bytes = e.bytes;
    if (flops) *flops = e.flops;
    return SDC_OK;
}
