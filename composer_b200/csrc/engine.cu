// The engine: parameter layout, workspace carving and the orchestration of the
// kernels into Transformer.call / loss / tape.gradient / Adam / cached decoding.
// Everything is enqueued on the caller's stream; see include/composer_b200.h
// for the contract and the reference lines each entry point replaces.
#include <atomic>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/composer_b200.h"
#include "attention.h"
#include "decode.h"
#include "elementwise.h"
#include "gemm.h"

namespace cb200 {

static thread_local char g_error[1024] = "";

static std::atomic<long long> g_launches{0};
void note_launch(long long n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
}

typedef __nv_bfloat16 bf16;

struct LayerOffsets {
    int64_t ln1_g, ln1_b, attn_w, attn_b, proj_w, proj_b, ln2_g, ln2_b, fc_w, fc_b, proj2_w, proj2_b;
};

struct TensorInfo {
    std::string name;
    int64_t offset;
    int rows, cols;
};

struct ParamLayout {
    int64_t wte, wpe, lnf_g, lnf_b, total;
    std::vector<LayerOffsets> layers;
    std::vector<TensorInfo> tensors;
};

static int build_layout(const cb200_config& c, ParamLayout& out) {
    CB200_REQUIRE(c.vocab_size > 0 && c.embedding_size > 0 && c.window_size > 0 && c.decoder_layers_count > 0 &&
                      c.attention_head_count > 0,
                  "invalid configuration");
    CB200_REQUIRE(c.embedding_size % c.attention_head_count == 0,
                  "embedding_size must be divisible by attention_head_count");   // transformer.py:255
    CB200_REQUIRE(c.embedding_size % 256 == 0 && c.embedding_size <= 1024,
                  "this build needs embedding_size to be a multiple of 256 and <= 1024 (got %d)", c.embedding_size);
    const int d = c.embedding_size / c.attention_head_count;
    CB200_REQUIRE(d == 16 || d == 32 || d == 64, "head size %d is not supported (16, 32 or 64)", d);
    CB200_REQUIRE(c.vocab_size <= 512, "vocabularies above 512 symbols are not supported by the fused logits kernel (got %d)",
                  c.vocab_size);
    const int64_t E = c.embedding_size, V = c.vocab_size, W = c.window_size, F = 4 * E;
    int64_t off = 0;
    out.tensors.clear();
    out.layers.clear();
    auto add = [&](const std::string& name, int rows, int cols) {
        const int64_t o = off;
        out.tensors.push_back(TensorInfo{name, o, rows, cols});
        off += static_cast<int64_t>(rows) * cols;
        off = (off + 3) & ~int64_t(3);   // keep every tensor 16-byte aligned
        return o;
    };
    out.wte = add("wte/weight", (int)V, (int)E);
    out.wpe = add("wpe/embeddings", (int)W, (int)E);
    for (int i = 1; i <= c.decoder_layers_count; ++i) {
        const std::string p = "h_" + std::to_string(i) + "/";
        LayerOffsets l;
        l.ln1_g = add(p + "ln_1/gamma", 1, (int)E);
        l.ln1_b = add(p + "ln_1/beta", 1, (int)E);
        l.attn_w = add(p + "attn/c_attn/weight", (int)E, (int)(3 * E));
        l.attn_b = add(p + "attn/c_attn/bias", 1, (int)(3 * E));
        l.proj_w = add(p + "attn/c_proj/weight", (int)E, (int)E);
        l.proj_b = add(p + "attn/c_proj/bias", 1, (int)E);
        l.ln2_g = add(p + "ln_2/gamma", 1, (int)E);
        l.ln2_b = add(p + "ln_2/beta", 1, (int)E);
        l.fc_w = add(p + "mlp/c_fc/weight", (int)E, (int)F);
        l.fc_b = add(p + "mlp/c_fc/bias", 1, (int)F);
        l.proj2_w = add(p + "mlp/c_proj/weight", (int)F, (int)E);
        l.proj2_b = add(p + "mlp/c_proj/bias", 1, (int)E);
        out.layers.push_back(l);
    }
    out.lnf_g = add("ln_f/gamma", 1, (int)E);
    out.lnf_b = add("ln_f/beta", 1, (int)E);
    out.total = off;
    return 0;
}

struct LayerBuffers {
    float *ln1_stats, *ln2_stats, *lse;
    bf16 *x1, *qkv, *att, *x2, *mln, *u, *gl, *x3;
};

struct Engine {
    cb200_config cfg;
    ParamLayout lay;
    int E, V, W, L, H, D, F, Vpad;
    // shadow arena offsets of the transposed copies
    std::vector<int64_t> shT_attn, shT_proj, shT_fc, shT_proj2;
    int64_t shT_wte, shadow_total;
    // bound memory
    float *params = nullptr, *grads = nullptr, *adam_m = nullptr, *adam_v = nullptr;
    bf16* shadow = nullptr;
    uint8_t* ws = nullptr;
    int64_t ws_bytes = 0;
    int max_B = 0, max_T = 0, training = 0;
    TransposeJob* jobs_dev = nullptr;
    int num_jobs = 0;
    // carved workspace
    bf16* h0 = nullptr;
    std::vector<LayerBuffers> lb;
    float* lnf_stats = nullptr;
    bf16 *hf = nullptr, *dlogits = nullptr;
    bf16 *bufP = nullptr, *bufQ = nullptr, *bufR = nullptr, *bufG = nullptr, *du = nullptr, *dqkv = nullptr;
    float *dq_acc = nullptr, *delta = nullptr;
    // state of the last training forward
    const int32_t* ids = nullptr;
    int B = 0, T = 0;
    DropoutParams drop_res{}, drop_attn{};
    bool have_forward = false;
    cudaStream_t decode_stream = nullptr;
};

static DropoutParams make_dropout(float rate, uint64_t seed, uint32_t step, bool enabled) {
    DropoutParams p;
    p.seed_lo = static_cast<uint32_t>(seed);
    p.seed_hi = static_cast<uint32_t>(seed >> 32);
    p.step = step;
    if (!enabled || rate <= 0.f) {
        p.threshold16 = 0;
        p.keep_scale = 1.f;
        p.rate = 0.f;
    } else {
        double t = static_cast<double>(rate) * 65536.0 + 0.5;
        if (t > 65535.0) t = 65535.0;
        p.threshold16 = static_cast<uint32_t>(t);
        p.keep_scale = 1.0f / (1.0f - rate);   // Keras Dropout: kept values scaled by 1 / (1 - rate)
        p.rate = rate;
    }
    return p;
}

struct Bump {
    uint8_t* base;
    int64_t off = 0;
    template <typename T>
    T* take(int64_t count) {
        off = (off + 255) & ~int64_t(255);
        T* p = base ? reinterpret_cast<T*>(base + off) : nullptr;
        off += count * static_cast<int64_t>(sizeof(T));
        return p;
    }
};

// Lays the workspace out; with base == nullptr only the size is computed.
static int64_t carve(Engine& e, uint8_t* base, int B, int T, int training) {
    Bump b{base};
    const int64_t M = static_cast<int64_t>(B) * T, E = e.E, F = e.F;
    e.jobs_dev = b.take<TransposeJob>(4 * e.L + 1);
    e.h0 = b.take<bf16>(M * E);
    e.lb.assign(e.L, LayerBuffers{});
    for (int l = 0; l < e.L; ++l) {
        if (!training && l > 0) {   // inference: every block reuses block 0's buffers
            e.lb[l] = e.lb[0];
            continue;
        }
        LayerBuffers& x = e.lb[l];
        x.ln1_stats = b.take<float>(2 * M);
        x.ln2_stats = b.take<float>(2 * M);
        x.lse = b.take<float>(static_cast<int64_t>(B) * e.H * T);
        x.x1 = b.take<bf16>(M * E);
        x.qkv = b.take<bf16>(M * 3 * E);
        x.att = b.take<bf16>(M * E);
        x.x2 = b.take<bf16>(M * E);
        x.mln = b.take<bf16>(M * E);
        x.u = b.take<bf16>(M * F);
        x.gl = b.take<bf16>(M * F);
        if (training) x.x3 = b.take<bf16>(M * E);
    }
    if (!training) {
        // block outputs alternate between two buffers so that a block never overwrites its own input
        bf16* ping = b.take<bf16>(M * E);
        bf16* pong = b.take<bf16>(M * E);
        for (int l = 0; l < e.L; ++l) e.lb[l].x3 = (l & 1) ? pong : ping;
    }
    e.lnf_stats = b.take<float>(2 * M);
    e.hf = b.take<bf16>(M * E);
    if (training) {
        e.dlogits = b.take<bf16>(M * e.Vpad);
        e.bufP = b.take<bf16>(M * E);
        e.bufQ = b.take<bf16>(M * E);
        e.bufR = b.take<bf16>(M * E);
        e.bufG = b.take<bf16>(M * E);
        e.du = b.take<bf16>(M * F);
        e.dqkv = b.take<bf16>(M * 3 * E);
        e.dq_acc = b.take<float>(M * E);
        e.delta = b.take<float>(static_cast<int64_t>(B) * e.H * T);
    } else {
        e.dlogits = e.bufP = e.bufQ = e.bufR = e.bufG = e.du = e.dqkv = nullptr;
        e.dq_acc = e.delta = nullptr;
    }
    return (b.off + 255) & ~int64_t(255);
}

static GemmDesc gemm_desc(GemmKind kind, int M, int N, int K, const bf16* A, int lda, const bf16* Bm, int ldb) {
    GemmDesc d;
    memset(&d, 0, sizeof(d));
    d.kind = kind; d.M = M; d.N = N; d.K = K; d.A = A; d.lda = lda; d.B = Bm; d.ldb = ldb;
    return d;
}

static int refresh_shadows(Engine& e, cudaStream_t s) {
    int rc = cast_bf16(e.params, e.shadow, static_cast<size_t>(e.lay.total), s);
    if (rc) return rc;
    return transpose_cast(e.params, e.shadow, e.jobs_dev, e.num_jobs, s);
}

// ---------------------------------------------------------------------------
// Forward (+ loss)
// ---------------------------------------------------------------------------
// Where the k, v rows of a forward pass go when it doubles as the prefill of a KV cache (transformer.py:419-432: the
// `presents` of a call without `past`).  layout 0: [L, 2, B, H, t_max, d_h] (the reference's present per layer, with
// room for t_max positions; per-step kernels); 1: the cluster kernel's [L, B, H, t_max, 2, d_h], swizzled pieces.
struct KvExport {
    bf16* cache;
    int t_max, layout;
    bool skip_head;     // the caller does not need logits: stop after the last block
};

static int forward(Engine& e, const int32_t* ids, const int32_t* labels, int B, int T, int training, uint64_t seed,
                   uint32_t step, float grad_scale, float* loss_sum, int32_t* correct, float* logits, cudaStream_t s,
                   const KvExport* kx = nullptr) {
    CB200_REQUIRE(e.params != nullptr && e.ws != nullptr, "engine is not bound");
    CB200_REQUIRE(B >= 1 && T >= 1 && static_cast<int64_t>(B) * T <= static_cast<int64_t>(e.max_B) * e.max_T,
                  "batch %d x %d exceeds the bound workspace (%d x %d)", B, T, e.max_B, e.max_T);
    CB200_REQUIRE(T <= e.W, "sequence length %d exceeds window_size %d (wpe has only window_size rows)", T, e.W);
    CB200_REQUIRE(!training || e.training, "engine was bound for inference only");
    const int M = B * T, E = e.E, F = e.F;
    const DropoutParams drop_res = make_dropout(e.cfg.residual_dropout_rate, seed, step, training != 0);
    const DropoutParams drop_attn = make_dropout(e.cfg.attention_dropout_rate, seed, step, training != 0);
    const float* P = e.params;
    const bf16* S = e.shadow;
    const float att_scale = e.cfg.scale_attention ? 1.0f / sqrtf(static_cast<float>(e.D)) : 1.0f;
    int rc;

    if ((rc = embed_fwd(ids, P + e.lay.wte, P + e.lay.wpe, e.h0, B, T, E, 0, e.V, drop_res, s))) return rc;
    const bf16* x_in = e.h0;
    for (int l = 0; l < e.L; ++l) {
        const LayerOffsets& o = e.lay.layers[l];
        LayerBuffers& x = e.lb[l];
        const uint32_t layer = static_cast<uint32_t>(l + 1);
        const bf16* x1 = x_in;
        if (e.cfg.use_layer_normalization) {
            if ((rc = layernorm_fwd(x_in, P + o.ln1_g, P + o.ln1_b, x.x1, x.ln1_stats, M, E, e.cfg.layer_normalization_epsilon, s))) return rc;
            x1 = x.x1;
        }
        {   // c_attn (transformer.py:416)
            GemmDesc d = gemm_desc(GEMM_BIAS, M, 3 * E, E, x1, E, S + e.shT_attn[l], E);
            d.bias = P + o.attn_b; d.out0 = x.qkv; d.ld_out0 = 3 * E;
            if ((rc = gemm_launch(d, s))) return rc;
        }
        if (kx != nullptr) {
            const int64_t layer_stride = 2ll * B * e.H * kx->t_max * e.D;
            bf16* cl = kx->cache + l * layer_stride;
            if (kx->layout == 1) rc = kv_export_mega(x.qkv, cl, B, T, e.H, e.D, kx->t_max, s);
            else rc = kv_export(x.qkv, cl, cl + layer_stride / 2, B, T, e.H, e.D, kx->t_max, s);
            if (rc) return rc;
        }
        if ((rc = attention_fwd(x.qkv, x.att, x.lse, B, T, e.H, e.D, att_scale, drop_attn, layer, s))) return rc;
        {   // attn c_proj + dropout + residual (transformer.py:443-444, 587)
            GemmDesc d = gemm_desc(GEMM_BIAS_DROP_RES, M, E, E, x.att, E, S + e.shT_proj[l], E);
            d.bias = P + o.proj_b; d.out0 = x.x2; d.ld_out0 = E; d.aux = x1; d.ld_aux = E;
            d.drop = drop_res; d.drop_site = SITE_ATTN_RESID; d.drop_layer = layer;
            if ((rc = gemm_launch(d, s))) return rc;
        }
        const bf16* mln = x.x2;
        if (e.cfg.use_layer_normalization) {
            if ((rc = layernorm_fwd(x.x2, P + o.ln2_g, P + o.ln2_b, x.mln, x.ln2_stats, M, E, e.cfg.layer_normalization_epsilon, s))) return rc;
            mln = x.mln;
        }
        {   // c_fc + gelu (transformer.py:504)
            GemmDesc d = gemm_desc(GEMM_BIAS_GELU, M, F, E, mln, E, S + e.shT_fc[l], E);
            d.bias = P + o.fc_b; d.out0 = x.u; d.ld_out0 = F; d.out1 = x.gl; d.ld_out1 = F;
            if ((rc = gemm_launch(d, s))) return rc;
        }
        {   // mlp c_proj + dropout + residual (transformer.py:505-506, 594)
            GemmDesc d = gemm_desc(GEMM_BIAS_DROP_RES, M, E, F, x.gl, F, S + e.shT_proj2[l], F);
            d.bias = P + o.proj2_b; d.out0 = x.x3; d.ld_out0 = E; d.aux = x.x2; d.ld_aux = E;
            d.drop = drop_res; d.drop_site = SITE_MLP; d.drop_layer = layer;
            if ((rc = gemm_launch(d, s))) return rc;
        }
        x_in = x.x3;
    }
    if (kx != nullptr && kx->skip_head) return 0;
    // ln_f (transformer.py:811) and the tied logits (:818) fused with the loss (:888, 918)
    if ((rc = layernorm_fwd(x_in, P + e.lay.lnf_g, P + e.lay.lnf_b, e.hf, e.lnf_stats, M, E, e.cfg.layer_normalization_epsilon, s))) return rc;
    {
        GemmDesc d = gemm_desc(GEMM_CE, M, e.V, E, e.hf, E, S + e.lay.wte, E);
        d.labels = labels;
        d.dlogits = (training && labels != nullptr) ? e.dlogits : nullptr;
        d.ld_dlogits = e.Vpad; d.grad_scale = grad_scale;
        d.loss_sum = labels ? loss_sum : nullptr; d.correct = labels ? correct : nullptr;
        d.outf = logits; d.ld_outf = e.V;
        if ((rc = gemm_launch(d, s))) return rc;
    }
    if (training) {
        e.ids = ids; e.B = B; e.T = T; e.drop_res = drop_res; e.drop_attn = drop_attn;
        e.have_forward = labels != nullptr;
    }
    return 0;
}

// ---------------------------------------------------------------------------
// Backward
// ---------------------------------------------------------------------------
static int backward_head(Engine& e, cudaStream_t s) {
    const int M = e.B * e.T, E = e.E;
    float* G = e.grads;
    int rc;
    {   // d hf = dlogits wte        (A [M, Vpad], B = wte^T padded [E, Vpad])
        GemmDesc d = gemm_desc(GEMM_BIAS, M, E, e.Vpad, e.dlogits, e.Vpad, e.shadow + e.shT_wte, e.Vpad);
        d.out0 = e.bufQ; d.ld_out0 = E;
        if ((rc = gemm_launch(d, s))) return rc;
    }
    {   // d wte += dlogits^T hf
        GemmDesc d = gemm_desc(GEMM_WGRAD, e.V, E, M, e.dlogits, e.Vpad, e.hf, E);
        d.outf = G + e.lay.wte; d.ld_outf = E;
        if ((rc = gemm_launch(d, s))) return rc;
    }
    const bf16* x_last = e.lb[e.L - 1].x3;
    // d x3 of the last block, and in the same pass what its MLP needs first: g = dropout_bwd(d x3), d b_proj2
    const LnBwdTail tail{e.bufG, G + e.lay.layers[e.L - 1].proj2_b, e.drop_res, SITE_MLP, static_cast<uint32_t>(e.L)};
    return layernorm_bwd_tail(e.bufQ, nullptr, x_last, e.lnf_stats, e.params + e.lay.lnf_g, nullptr, e.bufP,
                              G + e.lay.lnf_g, G + e.lay.lnf_b, M, E, tail, s);
}

__global__ void add3_kernel(const bf16* a, const bf16* b, const bf16* c, bf16* out, size_t n8) {
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n8;
         i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const uint4 ra = reinterpret_cast<const uint4*>(a)[i];
        const uint4 rb = b ? reinterpret_cast<const uint4*>(b)[i] : make_uint4(0, 0, 0, 0);
        const uint4 rc = c ? reinterpret_cast<const uint4*>(c)[i] : make_uint4(0, 0, 0, 0);
        const uint32_t wa[4] = {ra.x, ra.y, ra.z, ra.w}, wb[4] = {rb.x, rb.y, rb.z, rb.w}, wc[4] = {rc.x, rc.y, rc.z, rc.w};
        uint32_t o[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const float2 x = unpack_bf16(wa[e]), y = unpack_bf16(wb[e]), z = unpack_bf16(wc[e]);
            o[e] = pack_bf16(x.x + y.x + z.x, x.y + y.y + z.y);
        }
        reinterpret_cast<uint4*>(out)[i] = make_uint4(o[0], o[1], o[2], o[3]);
    }
}

static int add3(const bf16* a, const bf16* b, const bf16* c, bf16* out, size_t n, cudaStream_t s) {
    size_t n8 = n / 8;
    size_t blocks = (n8 + 255) / 256;
    if (blocks > 4096) blocks = 4096;
    add3_kernel<<<static_cast<int>(blocks), 256, 0, s>>>(a, b, c, out, n8);
    CB200_CUDA_OK(cudaGetLastError());
    note_launch(1);
    return 0;
}

// One decoder block; on entry bufP holds d(loss)/d(x3), on exit d(loss)/d(block input).
static int backward_layer(Engine& e, int l, cudaStream_t s) {
    const int M = e.B * e.T, E = e.E, F = e.F;
    const LayerOffsets& o = e.lay.layers[l];
    LayerBuffers& x = e.lb[l];
    const uint32_t layer = static_cast<uint32_t>(l + 1);
    float* G = e.grads;
    const float* P = e.params;
    const bf16* S = e.shadow;
    const bool ln = e.cfg.use_layer_normalization != 0;
    const bool dropping = e.drop_res.threshold16 != 0;
    const float att_scale = e.cfg.scale_attention ? 1.0f / sqrtf(static_cast<float>(e.D)) : 1.0f;
    const bf16* x_in = (l == 0) ? e.h0 : e.lb[l - 1].x3;
    const bf16* x1 = ln ? x.x1 : x_in;
    const bf16* mln = ln ? x.mln : x.x2;
    DropoutParams no_drop = e.drop_res;
    no_drop.threshold16 = 0;
    int rc;

    // ---- MLP ----
    // g = dropout_bwd(d x3); d b_proj2 += colsum(g): done by the kernel that produced d x3 (ln_f backward for the last
    // block, ln_1 backward of the block above otherwise).  Without LayerNorm in the blocks only the last block is
    // served that way (ln_f is unconditional, transformer.py:811); the others get their own pass here.
    if (!ln && l != e.L - 1 && (rc = bias_grad(e.bufP, e.bufG, G + o.proj2_b, M, E, e.drop_res, SITE_MLP, layer, s))) return rc;
    const bf16* g_mlp = dropping ? e.bufG : e.bufP;
    {   // du = (g W2^T) * gelu'(u)
        GemmDesc d = gemm_desc(GEMM_MUL_DGELU, M, F, E, g_mlp, E, S + o.proj2_w, E);
        d.out0 = e.du; d.ld_out0 = F; d.aux = x.u; d.ld_aux = F;
        if ((rc = gemm_launch(d, s))) return rc;
    }
    {   // d W2 += gelu(u)^T g
        GemmDesc d = gemm_desc(GEMM_WGRAD, F, E, M, x.gl, F, g_mlp, E);
        d.outf = G + o.proj2_w; d.ld_outf = E;
        if ((rc = gemm_launch(d, s))) return rc;
    }
    if ((rc = bias_grad(e.du, nullptr, G + o.fc_b, M, F, no_drop, 0, 0, s))) return rc;
    {   // d mln = du W1^T
        GemmDesc d = gemm_desc(GEMM_BIAS, M, E, F, e.du, F, S + o.fc_w, F);
        d.out0 = e.bufQ; d.ld_out0 = E;
        if ((rc = gemm_launch(d, s))) return rc;
    }
    {   // d W1 += mln^T du
        GemmDesc d = gemm_desc(GEMM_WGRAD, E, F, M, mln, E, e.du, F);
        d.outf = G + o.fc_w; d.ld_outf = F;
        if ((rc = gemm_launch(d, s))) return rc;
    }
    // d x2 = LN2_bwd(d mln) + d x3
    if (ln) {
        // ... and g = dropout_bwd(d x2), d b_proj for the attention projection in the same pass
        const LnBwdTail tail{e.bufG, G + o.proj_b, e.drop_res, SITE_ATTN_RESID, layer};
        if ((rc = layernorm_bwd_tail(e.bufQ, nullptr, x.x2, x.ln2_stats, P + o.ln2_g, e.bufP, e.bufR, G + o.ln2_g, G + o.ln2_b, M, E, tail, s))) return rc;
    } else {
        if ((rc = add3(e.bufQ, e.bufP, nullptr, e.bufR, static_cast<size_t>(M) * E, s))) return rc;
        if ((rc = bias_grad(e.bufR, e.bufG, G + o.proj_b, M, E, e.drop_res, SITE_ATTN_RESID, layer, s))) return rc;
    }
    // ---- attention ----
    const bf16* g_att = dropping ? e.bufG : e.bufR;
    {   // d att = g Wproj^T
        GemmDesc d = gemm_desc(GEMM_BIAS, M, E, E, g_att, E, S + o.proj_w, E);
        d.out0 = e.bufQ; d.ld_out0 = E;
        if ((rc = gemm_launch(d, s))) return rc;
    }
    {   // d Wproj += att^T g
        GemmDesc d = gemm_desc(GEMM_WGRAD, E, E, M, x.att, E, g_att, E);
        d.outf = G + o.proj_w; d.ld_outf = E;
        if ((rc = gemm_launch(d, s))) return rc;
    }
    if ((rc = attention_bwd(x.qkv, x.att, e.bufQ, x.lse, e.delta, e.dq_acc, e.dqkv, e.B, e.T, e.H, e.D, att_scale,
                            e.drop_attn, layer, s))) return rc;
    if ((rc = bias_grad(e.dqkv, nullptr, G + o.attn_b, M, 3 * E, no_drop, 0, 0, s))) return rc;
    {   // d x1 (attention path) = dqkv Wqkv^T
        GemmDesc d = gemm_desc(GEMM_BIAS, M, E, 3 * E, e.dqkv, 3 * E, S + o.attn_w, 3 * E);
        d.out0 = e.bufQ; d.ld_out0 = E;
        if ((rc = gemm_launch(d, s))) return rc;
    }
    {   // d Wqkv += x1^T dqkv
        GemmDesc d = gemm_desc(GEMM_WGRAD, E, 3 * E, M, x1, E, e.dqkv, 3 * E);
        d.outf = G + o.attn_w; d.ld_outf = 3 * E;
        if ((rc = gemm_launch(d, s))) return rc;
    }
    // d x_in = LN1_bwd(d x2 + d x1_attn): ln_1 overwrote the residual stream (transformer.py:583-587),
    // so the block input is reached only through ln_1.
    if (ln) {
        // (for l > 0 also the MLP tail of the block below: g = dropout_bwd(d x3), d b_proj2)
        LnBwdTail tail{};
        if (l > 0) tail = LnBwdTail{e.bufG, G + e.lay.layers[l - 1].proj2_b, e.drop_res, SITE_MLP, static_cast<uint32_t>(l)};
        if ((rc = layernorm_bwd_tail(e.bufR, e.bufQ, x_in, x.ln1_stats, P + o.ln1_g, nullptr, e.bufP, G + o.ln1_g, G + o.ln1_b, M, E, tail, s))) return rc;
    } else {
        if ((rc = add3(e.bufR, e.bufQ, nullptr, e.bufP, static_cast<size_t>(M) * E, s))) return rc;
    }
    return 0;
}

static int backward_embed(Engine& e, cudaStream_t s) {
    return embed_bwd(e.ids, e.bufP, e.grads + e.lay.wte, e.grads + e.lay.wpe, e.B, e.T, e.E, 0, e.V, e.drop_res, s);
}

static int backward(Engine& e, int stage, cudaStream_t s) {
    CB200_REQUIRE(e.have_forward, "cb200_backward needs a preceding training cb200_forward with labels");
    CB200_REQUIRE(e.grads != nullptr, "no gradient arena is bound");
    int rc;
    if (stage < 0) {
        if ((rc = backward_head(e, s))) return rc;
        for (int l = e.L - 1; l >= 0; --l)
            if ((rc = backward_layer(e, l, s))) return rc;
        return backward_embed(e, s);
    }
    if (stage == 0) return backward_head(e, s);
    if (stage <= e.L) return backward_layer(e, e.L - stage, s);
    if (stage == e.L + 1) return backward_embed(e, s);
    set_error("backward stage %d out of range", stage);
    return -1;
}

// ---------------------------------------------------------------------------
// Generation
// ---------------------------------------------------------------------------
static int g_decode_impl = 0;           // 0: persistent cluster kernel when the shape allows it, 1: per-step kernels + CUDA graph
static long long* g_decode_prof = nullptr;   // device buffer of 16 counters for the cluster kernel's phase profile
static int g_decode_cluster_size = 0;    // 0 = automatic, 4 or 8 CTAs per cluster
static int g_decode_max_clusters = 0;   // > 0 caps the clusters of the persistent kernel (tests)

static int g_decode_prefill = 1;        // 1: prompts go through one batched forward pass (KV export), 0: teacher-forced token by token

struct DecodeBuffers {
    int32_t *cur, *state, *all_ids, *forced, *prefix;
    float *logits, *uniforms;
    bf16 *x, *x1, *qkv, *att, *x2, *mln, *u, *gl, *y;
    float* stats;
    uint8_t* mega_stream;      // packed weight stream of the persistent cluster kernel
    int64_t mega_stream_bytes;
};

static int64_t carve_decode(const Engine& e, uint8_t* base, int B, int steps, DecodeBuffers& d) {
    Bump b{base};
    const int64_t E = e.E, F = e.F;
    d.state = b.take<int32_t>(8);
    d.cur = b.take<int32_t>(B);
    d.all_ids = b.take<int32_t>(static_cast<int64_t>(B) * steps);
    d.forced = b.take<int32_t>(static_cast<int64_t>(B) * steps);
    d.prefix = b.take<int32_t>(static_cast<int64_t>(B) * steps);
    d.uniforms = b.take<float>(static_cast<int64_t>(B) * steps);
    d.logits = b.take<float>(static_cast<int64_t>(B) * e.V);
    d.stats = b.take<float>(2 * B);
    d.x = b.take<bf16>(B * E);
    d.y = b.take<bf16>(B * E);
    d.x1 = b.take<bf16>(B * E);
    d.qkv = b.take<bf16>(B * 3 * E);
    d.att = b.take<bf16>(B * E);
    d.x2 = b.take<bf16>(B * E);
    d.mln = b.take<bf16>(B * E);
    d.u = b.take<bf16>(B * F);
    d.gl = b.take<bf16>(B * F);
    d.mega_stream_bytes = decode_mega_supported(e.E, e.H, e.D, e.V, e.L) ? decode_mega_stream_bytes(e.E, e.H, e.D, e.V, e.L) : 0;
    d.mega_stream = b.take<uint8_t>(d.mega_stream_bytes);
    return (b.off + 255) & ~int64_t(255);
}

// step0: the first step the decode loop runs (0, or prompt_len - 1 after a batched prefill of the tokens before it)
__global__ void decode_init_kernel(const int32_t* prompt, int B, int P, int steps, int step0, int32_t* cur,
                                   int32_t* forced, int32_t* prefix, int32_t* state) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < 8) state[i] = (i == 0 || i == 2) ? step0 : 0;      // [0] position, [1] ticket, [2] output column
    if (i < B) cur[i] = prompt[static_cast<size_t>(i) * P + step0];
    if (step0 > 0 && i < B * step0) prefix[i] = prompt[static_cast<size_t>(i / step0) * P + i % step0];
    if (i < B * steps) {
        const int b = i / steps, s = i % steps;
        forced[i] = (s + 1 < P) ? prompt[static_cast<size_t>(b) * P + s + 1] : -1;
    }
}

// One decoding step for all B sequences at position state[0] (device side).
static int decode_step(Engine& e, DecodeBuffers& d, bf16* cache, int t_max, int B, int steps, float temperature,
                       uint64_t seed, int64_t seq_base, cudaStream_t s) {
    const int E = e.E, F = e.F;
    const float* P = e.params;
    const bf16* S = e.shadow;
    const float att_scale = e.cfg.scale_attention ? 1.0f / sqrtf(static_cast<float>(e.D)) : 1.0f;
    const int64_t layer_stride = 2ll * B * e.H * t_max * e.D;
    int rc;
    if ((rc = decode_embed(d.cur, P + e.lay.wte, P + e.lay.wpe, d.x, d.state, B, E, e.V, s))) return rc;
    bf16* x_in = d.x;
    bf16* x_out = d.y;
    for (int l = 0; l < e.L; ++l) {
        const LayerOffsets& o = e.lay.layers[l];
        const bf16* x1 = x_in;
        if (e.cfg.use_layer_normalization) {
            if ((rc = layernorm_fwd(x_in, P + o.ln1_g, P + o.ln1_b, d.x1, d.stats, B, E, e.cfg.layer_normalization_epsilon, s))) return rc;
            x1 = d.x1;
        }
        // c_attn, in-place cache append + attention, c_proj + residual
        if ((rc = decode_linear(0, x1, E, S + e.shT_attn[l], P + o.attn_b, nullptr, 0, d.qkv, 3 * E, B, 3 * E, E, s))) return rc;
        bf16* kc = cache + l * layer_stride;
        bf16* vc = kc + layer_stride / 2;
        if ((rc = decode_attention(d.qkv, kc, vc, d.att, d.state, B, e.H, e.D, t_max, att_scale, s))) return rc;
        if ((rc = decode_linear(2, d.att, E, S + e.shT_proj[l], P + o.proj_b, x1, E, d.x2, E, B, E, E, s))) return rc;
        const bf16* mln = d.x2;
        if (e.cfg.use_layer_normalization) {
            if ((rc = layernorm_fwd(d.x2, P + o.ln2_g, P + o.ln2_b, d.mln, d.stats, B, E, e.cfg.layer_normalization_epsilon, s))) return rc;
            mln = d.mln;
        }
        // c_fc + gelu, mlp c_proj + residual
        if ((rc = decode_linear(1, mln, E, S + e.shT_fc[l], P + o.fc_b, nullptr, 0, d.gl, F, B, F, E, s))) return rc;
        if ((rc = decode_linear(2, d.gl, F, S + e.shT_proj2[l], P + o.proj2_b, d.x2, E, x_out, E, B, E, F, s))) return rc;
        bf16* t = x_in; x_in = x_out; x_out = t;
    }
    // ln_f + tied logits + sampling in one kernel
    return logits_sample(x_in, P + e.lay.lnf_g, P + e.lay.lnf_b, e.cfg.layer_normalization_epsilon, S + e.lay.wte, E, e.V,
                         temperature, seed, static_cast<int>(seq_base), d.all_ids, steps, d.cur, d.forced, steps,
                         d.state, d.state + 2, d.uniforms, d.logits, B, s);
}

static int generate_on(Engine& e, bf16* cache, int t_max, uint8_t* ws, int64_t ws_bytes, const int32_t* prompt, int B,
                       int P, int n_new, float temperature, uint64_t seed, int64_t seq_base, int32_t* out_ids,
                       float* uniforms_out, float* step_logits, cudaStream_t s) {
    const int steps = P - 1 + n_new;
    DecodeBuffers d;
    const int64_t need = carve_decode(e, nullptr, B, steps, d);
    CB200_REQUIRE(ws_bytes >= need, "decode workspace too small: %lld < %lld", (long long)ws_bytes, (long long)need);
    carve_decode(e, ws, B, steps, d);
    // Prompt tokens 0 .. P-2 only feed the cache: one batched forward pass (the training kernels, inference mode)
    // writes their k, v rows, and the decode loop starts at the last prompt token (transformer.py:735-770: a call
    // without `past` returns the presents of the whole prompt).  Needs an engine workspace bound for B x (P - 1).
    const bool prefill = g_decode_prefill != 0 && P > 1 && e.ws != nullptr &&
                         static_cast<int64_t>(B) * (P - 1) <= static_cast<int64_t>(e.max_B) * e.max_T;
    const int step0 = prefill ? P - 1 : 0;
    const int n_init = B * steps > 8 ? B * steps : 8;
    decode_init_kernel<<<(n_init + 255) / 256, 256, 0, s>>>(prompt, B, P, steps, step0, d.cur, d.forced, d.prefix, d.state);
    CB200_CUDA_OK(cudaGetLastError());
    note_launch(1);
    int rc;
    // (the cluster kernel copies whole 64-token chunks of the cache: t_max must be a multiple of the chunk)
    const bool mega = g_decode_impl == 0 && t_max % 64 == 0 && decode_mega_supported(e.E, e.H, e.D, e.V, e.L);
    if (mega)   // chunks are copied whole: positions that are not cached yet must read as zeros (V rows enter an MMA)
        CB200_CUDA_OK(cudaMemsetAsync(cache, 0, sizeof(bf16) * 2ull * B * e.H * t_max * e.D * e.L, s));
    if (prefill) {
        if (uniforms_out) CB200_CUDA_OK(cudaMemsetAsync(d.uniforms, 0, sizeof(float) * B * steps, s));
        const KvExport kx{cache, t_max, mega ? 1 : 0, true};
        if ((rc = forward(e, d.prefix, nullptr, B, P - 1, 0, 0, 0, 0.f, nullptr, nullptr, nullptr, s, &kx))) return rc;
    }
    if (mega) {
        // the whole generation (all steps, all layers) is one persistent cluster kernel
        MegaArgs m{};
        m.params = e.params; m.shadow = e.shadow; m.cache = cache;
        m.first = d.cur; m.forced = d.forced; m.out_ids = d.all_ids; m.uniforms = d.uniforms; m.logits_out = d.logits;
        m.prof = g_decode_prof;
        m.layer_stride = 2ll * B * e.H * t_max * e.D;
        m.B = B; m.E = e.E; m.H = e.H; m.F = e.F; m.V = e.V; m.L = e.L; m.t_max = t_max; m.steps = steps;
        m.step0 = step0;
        m.use_ln = e.cfg.use_layer_normalization ? 1 : 0;
        m.greedy = temperature <= 0.f ? 1 : 0;
        m.seq_base = static_cast<int>(seq_base);
        m.eps = e.cfg.layer_normalization_epsilon;
        m.scale_log2 = (e.cfg.scale_attention ? 1.0f / sqrtf(static_cast<float>(e.D)) : 1.0f) * 1.4426950408889634f;
        m.inv_temperature = m.greedy ? 1.f : 1.0f / temperature;
        m.seed_lo = static_cast<uint32_t>(seed); m.seed_hi = static_cast<uint32_t>(seed >> 32);
        m.wte = static_cast<uint32_t>(e.lay.wte); m.wpe = static_cast<uint32_t>(e.lay.wpe);
        m.lnf_g = static_cast<uint32_t>(e.lay.lnf_g); m.lnf_b = static_cast<uint32_t>(e.lay.lnf_b);
        m.wte_sh = static_cast<uint32_t>(e.lay.wte);
        for (int l = 0; l < e.L; ++l) {
            const LayerOffsets& o = e.lay.layers[l];
            MegaLayer& w = m.layers[l];
            w.ln1_g = static_cast<uint32_t>(o.ln1_g); w.ln1_b = static_cast<uint32_t>(o.ln1_b);
            w.ln2_g = static_cast<uint32_t>(o.ln2_g); w.ln2_b = static_cast<uint32_t>(o.ln2_b);
            w.attn_b = static_cast<uint32_t>(o.attn_b); w.proj_b = static_cast<uint32_t>(o.proj_b);
            w.fc_b = static_cast<uint32_t>(o.fc_b); w.proj2_b = static_cast<uint32_t>(o.proj2_b);
            w.attn_w = static_cast<uint32_t>(e.shT_attn[l]); w.proj_w = static_cast<uint32_t>(e.shT_proj[l]);
            w.fc_w = static_cast<uint32_t>(e.shT_fc[l]); w.proj2_w = static_cast<uint32_t>(e.shT_proj2[l]);
        }
        if ((rc = decode_mega(m, e.D, g_decode_max_clusters, g_decode_cluster_size, d.mega_stream, d.mega_stream_bytes, s))) return rc;
    } else {
    // step 0 runs eagerly (also configures kernel attributes outside of capture); the rest replays a graph
    if ((rc = decode_step(e, d, cache, t_max, B, steps, temperature, seed, seq_base, s))) return rc;
    if (steps - step0 > 1) {
        cudaGraph_t graph = nullptr;
        cudaGraphExec_t exec = nullptr;
        CB200_CUDA_OK(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
        const long long before = g_launches.load();
        rc = decode_step(e, d, cache, t_max, B, steps, temperature, seed, seq_base, s);
        const long long per_step = g_launches.load() - before;
        note_launch(per_step * (steps - step0 - 2));   // the captured step itself is replayed steps-step0-1 times
        cudaError_t ce = cudaStreamEndCapture(s, &graph);
        if (rc) { if (graph) cudaGraphDestroy(graph); return rc; }
        CB200_CUDA_OK(ce);
        CB200_CUDA_OK(cudaGraphInstantiate(&exec, graph, 0));
        for (int i = step0 + 1; i < steps; ++i) {
            cudaError_t le = cudaGraphLaunch(exec, s);
            if (le != cudaSuccess) {
                cudaGraphExecDestroy(exec); cudaGraphDestroy(graph);
                set_error("cudaGraphLaunch failed: %s", cudaGetErrorString(le));
                return -2;
            }
        }
        cudaError_t se = cudaStreamSynchronize(s);
        cudaGraphExecDestroy(exec);
        cudaGraphDestroy(graph);
        CB200_CUDA_OK(se);
    }
    }
    CB200_CUDA_OK(cudaMemcpy2DAsync(out_ids, n_new * sizeof(int32_t), d.all_ids + (P - 1), steps * sizeof(int32_t),
                                    n_new * sizeof(int32_t), B, cudaMemcpyDeviceToDevice, s));
    if (uniforms_out)
        CB200_CUDA_OK(cudaMemcpyAsync(uniforms_out, d.uniforms, sizeof(float) * B * steps, cudaMemcpyDeviceToDevice, s));
    if (step_logits)
        CB200_CUDA_OK(cudaMemcpyAsync(step_logits, d.logits, sizeof(float) * B * e.V, cudaMemcpyDeviceToDevice, s));
    CB200_CUDA_OK(cudaStreamSynchronize(s));
    return 0;
}

// The step graph is captured on a private stream (the caller's may be the legacy
// default stream, which cannot be captured); the caller's stream is drained first
// and the private stream is drained before returning, so ordering is preserved.
static int generate(Engine& e, bf16* cache, int t_max, uint8_t* ws, int64_t ws_bytes, const int32_t* prompt, int B,
                    int P, int n_new, float temperature, uint64_t seed, int64_t seq_base, int32_t* out_ids,
                    float* uniforms_out, float* step_logits, cudaStream_t user_stream) {
    CB200_REQUIRE(e.params != nullptr, "engine is not bound");
    CB200_REQUIRE(B >= 1 && P >= 1 && n_new >= 1, "generate needs B, prompt_len, n_new >= 1");
    const int steps = P - 1 + n_new;
    // positions 0 .. steps-1 are embedded; wpe has window_size rows (transformer.py:675-679, 770)
    CB200_REQUIRE(steps <= e.W, "prompt_len + length - 1 = %d positions exceed window_size %d", steps, e.W);
    CB200_REQUIRE(steps <= t_max, "KV cache too small: %d positions, t_max %d", steps, t_max);
    CB200_CUDA_OK(cudaStreamSynchronize(user_stream));
    if (e.decode_stream == nullptr)
        CB200_CUDA_OK(cudaStreamCreateWithFlags(&e.decode_stream, cudaStreamNonBlocking));
    return generate_on(e, cache, t_max, ws, ws_bytes, prompt, B, P, n_new, temperature, seed, seq_base, out_ids,
                       uniforms_out, step_logits, e.decode_stream);
}

// `Transformer.call(inputs, past=None)` for a KV cache the caller keeps: logits of every position (optional) and the
// presents of the prompt written into `cache` ([L, 2, B, H, t_max, d_h]: layer l's slice [2, B, H, :T, d_h] is the
// reference's `present`, transformer.py:430-432).
static int prefill(Engine& e, const int32_t* ids, int B, int T, bf16* cache, int t_max, float* logits, cudaStream_t s) {
    CB200_REQUIRE(T <= t_max, "KV cache too small: %d positions, t_max %d", T, t_max);
    const KvExport kx{cache, t_max, 0, logits == nullptr};
    return forward(e, ids, nullptr, B, T, 0, 0, 0, 0.f, nullptr, nullptr, logits, s, &kx);
}

__global__ void decode_step_init_kernel(const int32_t* ids, int B, int pos, int32_t* cur, int32_t* forced, int32_t* state) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < 8) state[i] = (i == 0) ? pos : 0;
    if (i < B) { cur[i] = ids[i]; forced[i] = -1; }
}

// `Transformer.call(inputs[:, -1:], past=presents)` (transformer.py:735-737, 423-426): one token per sequence at
// position `pos` = the past length, k, v appended to the cache in place, logits [B, V] of that position.
static int decode_one(Engine& e, bf16* cache, int t_max, uint8_t* ws, int64_t ws_bytes, const int32_t* ids, int B,
                      int pos, float* logits_out, cudaStream_t s) {
    CB200_REQUIRE(e.params != nullptr, "engine is not bound");
    CB200_REQUIRE(pos >= 0 && pos < t_max, "position %d outside the KV cache (t_max %d)", pos, t_max);
    CB200_REQUIRE(pos < e.W, "position %d exceeds window_size %d (wpe has only window_size rows)", pos, e.W);
    DecodeBuffers d;
    const int64_t need = carve_decode(e, nullptr, B, 1, d);
    CB200_REQUIRE(ws_bytes >= need, "decode workspace too small: %lld < %lld", (long long)ws_bytes, (long long)need);
    carve_decode(e, ws, B, 1, d);
    decode_step_init_kernel<<<(B + 255) / 256 > 0 ? (B + 255) / 256 : 1, 256, 0, s>>>(ids, B, pos, d.cur, d.forced, d.state);
    CB200_CUDA_OK(cudaGetLastError());
    note_launch(1);
    int rc = decode_step(e, d, cache, t_max, B, 1, 0.f, 0, 0, s);
    if (rc) return rc;
    if (logits_out)
        CB200_CUDA_OK(cudaMemcpyAsync(logits_out, d.logits, sizeof(float) * B * e.V, cudaMemcpyDeviceToDevice, s));
    return 0;
}

}  // namespace cb200

// ===========================================================================
// C ABI
// ===========================================================================
using namespace cb200;

extern "C" {

const char* cb200_last_error(void) { return g_error; }
int cb200_abi_version(void) { return 1; }
long long cb200_launch_count(void) { return g_launches.load(); }

int64_t cb200_param_elems(const cb200_config* cfg) {
    ParamLayout l;
    if (!cfg || build_layout(*cfg, l)) return -1;
    return l.total;
}

int cb200_param_tensor_count(const cb200_config* cfg) {
    ParamLayout l;
    if (!cfg || build_layout(*cfg, l)) return -1;
    return static_cast<int>(l.tensors.size());
}

int cb200_param_tensor_info(const cb200_config* cfg, int index, char* name, int name_capacity, int64_t* offset,
                            int32_t* rows, int32_t* cols) {
    ParamLayout l;
    if (!cfg) { set_error("null config"); return -1; }
    int rc = build_layout(*cfg, l);
    if (rc) return rc;
    CB200_REQUIRE(index >= 0 && index < (int)l.tensors.size(), "tensor index %d out of range", index);
    const TensorInfo& t = l.tensors[index];
    if (name && name_capacity > 0) snprintf(name, name_capacity, "%s", t.name.c_str());
    if (offset) *offset = t.offset;
    if (rows) *rows = t.rows;
    if (cols) *cols = t.cols;
    return 0;
}

int cb200_engine_create(const cb200_config* cfg, void** engine) {
    CB200_REQUIRE(cfg && engine, "null argument");
    Engine* e = new Engine();
    e->cfg = *cfg;
    int rc = build_layout(*cfg, e->lay);
    if (rc) { delete e; return rc; }
    e->E = cfg->embedding_size; e->V = cfg->vocab_size; e->W = cfg->window_size; e->L = cfg->decoder_layers_count;
    e->H = cfg->attention_head_count; e->D = e->E / e->H; e->F = 4 * e->E;
    e->Vpad = (e->V + 15) / 16 * 16;
    int64_t off = (e->lay.total + 127) & ~int64_t(127);
    auto take = [&](int64_t n) { int64_t o = off; off = (off + n + 127) & ~int64_t(127); return o; };
    for (int l = 0; l < e->L; ++l) {
        e->shT_attn.push_back(take(3ll * e->E * e->E));
        e->shT_proj.push_back(take(1ll * e->E * e->E));
        e->shT_fc.push_back(take(1ll * e->F * e->E));
        e->shT_proj2.push_back(take(1ll * e->E * e->F));
    }
    e->shT_wte = take(1ll * e->E * e->Vpad);
    e->shadow_total = off;
    *engine = e;
    return 0;
}

int cb200_engine_destroy(void* engine) {
    Engine* e = static_cast<Engine*>(engine);
    if (e && e->decode_stream) cudaStreamDestroy(e->decode_stream);
    delete e;
    return 0;
}

int64_t cb200_shadow_elems(void* engine) { return engine ? static_cast<Engine*>(engine)->shadow_total : -1; }

int64_t cb200_workspace_bytes(void* engine, int B, int T, int training) {
    if (!engine) return -1;
    Engine tmp = *static_cast<Engine*>(engine);
    return carve(tmp, nullptr, B, T, training);
}

int cb200_engine_bind(void* engine, float* params, float* grads, float* adam_m, float* adam_v, void* shadow,
                      void* workspace, int64_t workspace_bytes, int max_B, int max_T, int training) {
    CB200_REQUIRE(engine && params && shadow && workspace, "null argument");
    Engine& e = *static_cast<Engine*>(engine);
    CB200_REQUIRE(!training || (grads && adam_m && adam_v), "training needs gradient and Adam arenas");
    const int64_t need = carve(e, nullptr, max_B, max_T, training);
    CB200_REQUIRE(workspace_bytes >= need, "workspace too small: %lld < %lld bytes", (long long)workspace_bytes, (long long)need);
    e.params = params; e.grads = grads; e.adam_m = adam_m; e.adam_v = adam_v;
    e.shadow = static_cast<bf16*>(shadow);
    e.ws = static_cast<uint8_t*>(workspace); e.ws_bytes = workspace_bytes;
    e.max_B = max_B; e.max_T = max_T; e.training = training;
    e.have_forward = false;
    carve(e, e.ws, max_B, max_T, training);
    // transpose job table: [in, out] fp32 weights -> [out, in] bf16 shadows, and wte -> padded wte^T
    std::vector<TransposeJob> jobs;
    for (int l = 0; l < e.L; ++l) {
        const LayerOffsets& o = e.lay.layers[l];
        jobs.push_back(TransposeJob{(unsigned long long)o.attn_w, (unsigned long long)e.shT_attn[l], e.E, 3 * e.E, e.E, 0});
        jobs.push_back(TransposeJob{(unsigned long long)o.proj_w, (unsigned long long)e.shT_proj[l], e.E, e.E, e.E, 0});
        jobs.push_back(TransposeJob{(unsigned long long)o.fc_w, (unsigned long long)e.shT_fc[l], e.E, e.F, e.E, 0});
        jobs.push_back(TransposeJob{(unsigned long long)o.proj2_w, (unsigned long long)e.shT_proj2[l], e.F, e.E, e.F, 0});
    }
    jobs.push_back(TransposeJob{(unsigned long long)e.lay.wte, (unsigned long long)e.shT_wte, e.V, e.E, e.Vpad, 0});
    e.num_jobs = static_cast<int>(jobs.size());
    CB200_CUDA_OK(cudaMemcpy(e.jobs_dev, jobs.data(), jobs.size() * sizeof(TransposeJob), cudaMemcpyHostToDevice));
    CB200_CUDA_OK(cudaMemset(e.shadow + e.shT_wte, 0, sizeof(bf16) * e.E * e.Vpad));   // zero padding columns
    if (training) CB200_CUDA_OK(cudaMemset(e.dq_acc, 0, sizeof(float) * static_cast<size_t>(max_B) * max_T * e.E));
    return 0;
}

int cb200_refresh_shadows(void* engine, void* stream) {
    CB200_REQUIRE(engine, "null engine");
    Engine& e = *static_cast<Engine*>(engine);
    CB200_REQUIRE(e.params && e.shadow, "engine is not bound");
    return refresh_shadows(e, static_cast<cudaStream_t>(stream));
}

int cb200_forward(void* engine, const int32_t* ids, const int32_t* labels, int B, int T, int training, uint64_t seed,
                  uint32_t step, float grad_scale, float* loss_sum, int32_t* correct, float* logits, void* stream) {
    CB200_REQUIRE(engine && ids, "null argument");
    return forward(*static_cast<Engine*>(engine), ids, labels, B, T, training, seed, step, grad_scale, loss_sum,
                   correct, logits, static_cast<cudaStream_t>(stream));
}

int cb200_backward(void* engine, int stage, void* stream) {
    CB200_REQUIRE(engine, "null engine");
    return backward(*static_cast<Engine*>(engine), stage, static_cast<cudaStream_t>(stream));
}

int cb200_zero_grads(void* engine, void* stream) {
    CB200_REQUIRE(engine, "null engine");
    Engine& e = *static_cast<Engine*>(engine);
    CB200_REQUIRE(e.grads, "no gradient arena is bound");
    CB200_CUDA_OK(cudaMemsetAsync(e.grads, 0, sizeof(float) * e.lay.total, static_cast<cudaStream_t>(stream)));
    return 0;
}

int cb200_adam_step(void* engine, float lr, float b1, float b2, float eps, int64_t t, float grad_scale, void* stream) {
    CB200_REQUIRE(engine && t >= 1, "bad argument");
    Engine& e = *static_cast<Engine*>(engine);
    CB200_REQUIRE(e.grads && e.adam_m && e.adam_v, "no optimizer state is bound");
    const double lr_t = static_cast<double>(lr) * sqrt(1.0 - pow(static_cast<double>(b2), static_cast<double>(t))) /
                        (1.0 - pow(static_cast<double>(b1), static_cast<double>(t)));
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    int rc = adam_step(e.params, e.grads, e.adam_m, e.adam_v, e.shadow, static_cast<size_t>(e.lay.total),
                       static_cast<float>(lr_t), b1, b2, eps, grad_scale, s);
    if (rc) return rc;
    return transpose_cast(e.params, e.shadow, e.jobs_dev, e.num_jobs, s);
}

int64_t cb200_kv_cache_elems(void* engine, int B, int t_max) {
    if (!engine) return -1;
    Engine& e = *static_cast<Engine*>(engine);
    return 2ll * e.L * B * e.H * t_max * e.D;
}

int64_t cb200_decode_workspace_bytes(void* engine, int B) {
    if (!engine) return -1;
    Engine& e = *static_cast<Engine*>(engine);
    DecodeBuffers d;
    return carve_decode(e, nullptr, B, e.W, d);
}

int cb200_generate(void* engine, void* cache, int t_max, void* workspace, int64_t workspace_bytes,
                   const int32_t* prompt, int B, int prompt_len, int n_new, float temperature, uint64_t seed,
                   int64_t seq_index_base, int32_t* out_ids, float* uniforms_out, float* step_logits, void* stream) {
    CB200_REQUIRE(engine && cache && workspace && prompt && out_ids, "null argument");
    return generate(*static_cast<Engine*>(engine), static_cast<bf16*>(cache), t_max, static_cast<uint8_t*>(workspace),
                    workspace_bytes, prompt, B, prompt_len, n_new, temperature, seed, seq_index_base, out_ids,
                    uniforms_out, step_logits, static_cast<cudaStream_t>(stream));
}

int cb200_prefill(void* engine, const int32_t* ids, int B, int T, void* cache, int t_max, float* logits, void* stream) {
    CB200_REQUIRE(engine && ids && cache, "null argument");
    return prefill(*static_cast<Engine*>(engine), ids, B, T, static_cast<bf16*>(cache), t_max, logits,
                   static_cast<cudaStream_t>(stream));
}

int cb200_decode_step(void* engine, void* cache, int t_max, void* workspace, int64_t workspace_bytes, const int32_t* ids,
                      int B, int pos, float* logits, void* stream) {
    CB200_REQUIRE(engine && cache && workspace && ids, "null argument");
    return decode_one(*static_cast<Engine*>(engine), static_cast<bf16*>(cache), t_max, static_cast<uint8_t*>(workspace),
                      workspace_bytes, ids, B, pos, logits, static_cast<cudaStream_t>(stream));
}

int cb200_set_decode_prefill(int enabled) {
    g_decode_prefill = enabled ? 1 : 0;
    return 0;
}

// ---- single kernels ---------------------------------------------------------
int cb200_gemm(int kind, int M, int N, int K, const void* A, int lda, const void* B, int ldb, const float* bias,
               void* out0, int ld_out0, void* out1, int ld_out1, const void* aux, int ld_aux, float* outf, int ld_outf,
               float dropout_rate, uint64_t seed, uint32_t step, uint32_t site, uint32_t layer, void* stream) {
    CB200_REQUIRE(kind == 0 || kind == 1 || kind == 2 || kind == 3 || kind == 4 || kind == 6, "bad GEMM kind %d", kind);
    GemmDesc d = gemm_desc(static_cast<GemmKind>(kind), M, N, K, static_cast<const bf16*>(A), lda,
                           static_cast<const bf16*>(B), ldb);
    d.bias = bias; d.out0 = static_cast<bf16*>(out0); d.ld_out0 = ld_out0; d.out1 = static_cast<bf16*>(out1);
    d.ld_out1 = ld_out1; d.aux = static_cast<const bf16*>(aux); d.ld_aux = ld_aux; d.outf = outf; d.ld_outf = ld_outf;
    d.drop = make_dropout(dropout_rate, seed, step, dropout_rate > 0.f); d.drop_site = site; d.drop_layer = layer;
    return gemm_launch(d, static_cast<cudaStream_t>(stream));
}

int cb200_logits_ce(int M, int V, int E, const void* h, const void* wte, const int32_t* labels, void* dlogits,
                    int ld_dlogits, float grad_scale, float* loss_sum, int32_t* correct, float* logits, void* stream) {
    GemmDesc d = gemm_desc(GEMM_CE, M, V, E, static_cast<const bf16*>(h), E, static_cast<const bf16*>(wte), E);
    d.labels = labels; d.dlogits = static_cast<bf16*>(dlogits); d.ld_dlogits = ld_dlogits; d.grad_scale = grad_scale;
    d.loss_sum = loss_sum; d.correct = correct; d.outf = logits; d.ld_outf = V;
    return gemm_launch(d, static_cast<cudaStream_t>(stream));
}

int cb200_embed_fwd(const int32_t* ids, const float* wte, const float* wpe, void* out, int B, int T, int E, int pos0,
                    int vocab, float dropout_rate, uint64_t seed, uint32_t step, void* stream) {
    return embed_fwd(ids, wte, wpe, static_cast<bf16*>(out), B, T, E, pos0, vocab,
                     make_dropout(dropout_rate, seed, step, dropout_rate > 0.f), static_cast<cudaStream_t>(stream));
}

int cb200_embed_bwd(const int32_t* ids, const void* dh, float* dwte, float* dwpe, int B, int T, int E, int pos0,
                    int vocab, float dropout_rate, uint64_t seed, uint32_t step, void* stream) {
    return embed_bwd(ids, static_cast<const bf16*>(dh), dwte, dwpe, B, T, E, pos0, vocab,
                     make_dropout(dropout_rate, seed, step, dropout_rate > 0.f), static_cast<cudaStream_t>(stream));
}

int cb200_layernorm_fwd(const void* x, const float* gamma, const float* beta, void* y, float* stats, int rows, int E,
                        float eps, void* stream) {
    return layernorm_fwd(static_cast<const bf16*>(x), gamma, beta, static_cast<bf16*>(y), stats, rows, E, eps,
                         static_cast<cudaStream_t>(stream));
}

int cb200_layernorm_bwd(const void* dy_a, const void* dy_b, const void* x, const float* stats, const float* gamma,
                        const void* dres, void* dx, float* dgamma, float* dbeta, int rows, int E, void* stream) {
    return layernorm_bwd(static_cast<const bf16*>(dy_a), static_cast<const bf16*>(dy_b), static_cast<const bf16*>(x),
                         stats, gamma, static_cast<const bf16*>(dres), static_cast<bf16*>(dx), dgamma, dbeta, rows, E,
                         static_cast<cudaStream_t>(stream));
}

int cb200_bias_grad(const void* dy, void* g_out, float* dbias, int rows, int N, float dropout_rate, uint64_t seed,
                    uint32_t step, uint32_t site, uint32_t layer, void* stream) {
    return bias_grad(static_cast<const bf16*>(dy), static_cast<bf16*>(g_out), dbias, rows, N,
                     make_dropout(dropout_rate, seed, step, dropout_rate > 0.f), site, layer,
                     static_cast<cudaStream_t>(stream));
}

int cb200_attention_fwd(const void* qkv, void* out, float* lse, int B, int T, int H, int D, float scale,
                        float dropout_rate, uint64_t seed, uint32_t step, uint32_t layer, void* stream) {
    return attention_fwd(static_cast<const bf16*>(qkv), static_cast<bf16*>(out), lse, B, T, H, D, scale,
                         make_dropout(dropout_rate, seed, step, dropout_rate > 0.f), layer,
                         static_cast<cudaStream_t>(stream));
}

int cb200_attention_bwd(const void* qkv, const void* out, const void* dout, const float* lse, float* delta,
                        float* dq_acc, void* dqkv, int B, int T, int H, int D, float scale, float dropout_rate,
                        uint64_t seed, uint32_t step, uint32_t layer, void* stream) {
    return attention_bwd(static_cast<const bf16*>(qkv), static_cast<const bf16*>(out), static_cast<const bf16*>(dout),
                         lse, delta, dq_acc, static_cast<bf16*>(dqkv), B, T, H, D, scale,
                         make_dropout(dropout_rate, seed, step, dropout_rate > 0.f), layer,
                         static_cast<cudaStream_t>(stream));
}

int cb200_set_decode_impl(int impl, int max_clusters, int cluster_size) {
    CB200_REQUIRE(impl == 0 || impl == 1, "decode implementation must be 0 (persistent cluster kernel) or 1 (per-step kernels)");
    CB200_REQUIRE(max_clusters >= 0, "max_clusters must be >= 0");
    CB200_REQUIRE(cluster_size == 0 || cluster_size == 4 || cluster_size == 8, "cluster_size must be 0 (automatic), 4 or 8");
    g_decode_impl = impl;
    g_decode_max_clusters = max_clusters;
    g_decode_cluster_size = cluster_size;
    return 0;
}

int cb200_decode_cluster_capacity(void* engine, int cluster_size) {
    if (!engine) return -1;
    Engine& e = *static_cast<Engine*>(engine);
    if (cluster_size != 4 && cluster_size != 8) return 0;
    return decode_mega_capacity(e.E, e.H, e.V, e.D, e.L, cluster_size);
}

int cb200_set_decode_profile(void* counters) {
    g_decode_prof = static_cast<long long*>(counters);
    return 0;
}

int cb200_set_attention_fwd_impl(int impl) {
    CB200_REQUIRE((impl >= 0 && impl <= 4) || impl == 7 || (impl >= 100 && impl < 200),
                  "attention forward implementation must be 0 (tcgen05, P in TMEM), 1 (mma.sync), 2 (tcgen05, P in smem) "
                  "or 3 / 4 (tile-shape variants of 0)");
    attention_set_fwd_impl(impl);
    return 0;
}

int cb200_set_attention_trace(void* buffer) {
    attention_set_trace(static_cast<long long*>(buffer));
    return 0;
}

int cb200_attention_dropout_mask(uint8_t* mask, int B, int T, int H, float dropout_rate, uint64_t seed, uint32_t step,
                                 uint32_t layer, void* stream) {
    return attention_mask_export(mask, B, T, H, make_dropout(dropout_rate, seed, step, dropout_rate > 0.f), layer,
                                 static_cast<cudaStream_t>(stream));
}

__global__ void rowmajor_mask_kernel(uint8_t* mask, int rows, int cols, DropoutParams drop, uint32_t site, uint32_t layer) {
    const size_t n = static_cast<size_t>(rows) * (cols / 8);
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n;
         i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const uint32_t row = static_cast<uint32_t>(i / (cols / 8)), c8 = static_cast<uint32_t>(i % (cols / 8));
        const Philox4 r = drop_bits_rowmajor(drop, site, layer, row, c8);
        for (int e = 0; e < 8; ++e)
            mask[static_cast<size_t>(row) * cols + c8 * 8 + e] = (drop.threshold16 == 0 || drop_u16(r, e) >= drop.threshold16) ? 1 : 0;
    }
}

int cb200_rowmajor_dropout_mask(uint8_t* mask, int rows, int cols, float dropout_rate, uint64_t seed, uint32_t step,
                                uint32_t site, uint32_t layer, void* stream) {
    CB200_REQUIRE(cols % 8 == 0, "cols must be a multiple of 8");
    rowmajor_mask_kernel<<<1024, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        mask, rows, cols, make_dropout(dropout_rate, seed, step, dropout_rate > 0.f), site, layer);
    CB200_CUDA_OK(cudaGetLastError());
    note_launch(1);
    return 0;
}

int cb200_adam(float* p, const float* g, float* m, float* v, void* shadow, int64_t n, float lr_t, float beta_1,
               float beta_2, float epsilon, float grad_scale, void* stream) {
    return adam_step(p, g, m, v, static_cast<bf16*>(shadow), static_cast<size_t>(n), lr_t, beta_1, beta_2, epsilon,
                     grad_scale, static_cast<cudaStream_t>(stream));
}

int cb200_decode_attention(const void* qkv, void* kcache, void* vcache, void* out, const int32_t* pos, int B, int H,
                           int D, int t_max, float scale, void* stream) {
    return decode_attention(static_cast<const bf16*>(qkv), static_cast<bf16*>(kcache), static_cast<bf16*>(vcache),
                            static_cast<bf16*>(out), pos, B, H, D, t_max, scale, static_cast<cudaStream_t>(stream));
}

int cb200_decode_linear(int epilogue, const void* X, int ldx, const void* Wt, const float* bias, const void* res,
                        int ldres, void* Y, int ldy, int B, int N, int K, void* stream) {
    return decode_linear(epilogue, static_cast<const bf16*>(X), ldx, static_cast<const bf16*>(Wt), bias,
                         static_cast<const bf16*>(res), ldres, static_cast<bf16*>(Y), ldy, B, N, K,
                         static_cast<cudaStream_t>(stream));
}

}  // extern "C"
