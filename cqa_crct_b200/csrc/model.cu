// Whole-model inference forward scheduled by the library (crct_create / crct_bind_params / crct_workspace_bytes /
// crct_forward).  Host code only: it enqueues the operator entry points of this library (K1 GEMM, K2/K3 attention, K4 LayerNorm,
// K5/K6 embeddings, K7/K8 heads + loss, var-len row maps) in the order of `VisualDialogEncoder._run_forward` (cqa_crct_b200/
// encoder.py) in evaluation mode, i.e. the order of the reference's
//   BertModel.forward            CRCT/backbone/vilbert.py:1348-1441   (masks, embeddings, encoder, poolers)
//   BertEncoder.forward          :822-946                             (v / t / connection-layer schedule)
//   BertLayer / BertImageLayer   :474-485, :605-616
//   BertConnectionLayer          :774-788 (biOutput called with crossed arguments, :780)
//   BertPreTrainingHeads         :1048-1062, Regressor CRCT/backbone/regressor.py:36-42, losses vilbert.py:1586-1657
// with dropout off (inference).  Same kernels, same order, same operands: outputs are bit-identical to the Python host's.
// The handle owns host tables only; parameters, inputs, outputs and the workspace belong to the caller.
#include "common.cuh"
#include <string>
#include <unordered_map>
#include <utility>
#include <vector>

struct Bound {
    const float* w32;
    const void* w16;
    size_t numel;
};

struct crct_model_s {
    crct_config_t cfg;
    std::unordered_map<std::string, Bound> params;
    std::vector<std::pair<char, int>> schedule;      // ('v' | 't' | 'c', index): vilbert.py:852-939 flattened
};

namespace {

constexpr size_t ALIGN = 256;
inline size_t up(size_t n) { return (n + ALIGN - 1) / ALIGN * ALIGN; }

// Workspace plan: a fixed set of buffers sized for the batch, reused layer after layer (nothing is kept for a backward).
struct Plan {
    size_t total = 0;
    size_t take(size_t bytes) { const size_t o = total; total += up(bytes); return o; }
    // row maps
    size_t cu_t, src_t, cu_vq, src_vq, cu_v, src_v;
    // text stream
    size_t t[2], tres[2], qkv_t, ctx_t, z1_t, a_t, ares_t, h_t, z2_t;
    // visual stream
    size_t probs, gimg, vq, vqres, v[2], vres[2], qkv_v, ctx_v, z1_v, a_v, ares_v, h_v, z2_v;
    // heads (fp32)
    size_t hw0, hv0, pt, pv, pooled, pipe_t[3], pipe_v[3], prefusion, fus[4];
};

Plan make_plan(const crct_config_t& c, int B, int Bq, int T, int R, bool grouped) {
    Plan p;
    const size_t H = c.hidden_size, Hv = c.v_hidden_size, Hb = c.bi_hidden_size, I = c.intermediate_size, Iv = c.v_intermediate_size;
    const size_t F = c.v_feature_size, Mt = (size_t)B * T, Mvq = (size_t)Bq * R, Mv = (size_t)B * R;
    const size_t Wt = H > Hb ? H : Hb, Wv = Hv > Hb ? Hv : Hb;
    p.cu_t = p.take((B + 1) * 4); p.src_t = p.take(Mt * 4);
    p.cu_vq = p.take((Bq + 1) * 4); p.src_vq = p.take(Mvq * 4);
    p.cu_v = p.take((B + 1) * 4); p.src_v = p.take(Mv * 4);
    for (int i = 0; i < 2; ++i) { p.t[i] = p.take(Mt * H * 2); p.tres[i] = p.take(Mt * H * 4); }
    p.qkv_t = p.take(Mt * 3 * Wt * 2); p.ctx_t = p.take(Mt * Wt * 2); p.z1_t = p.take(Mt * H * 4);
    p.a_t = p.take(Mt * H * 2); p.ares_t = p.take(Mt * H * 4); p.h_t = p.take(Mt * I * 2); p.z2_t = p.take(Mt * H * 4);
    p.probs = p.take(Mvq * F * 2); p.gimg = p.take(Mvq * Hv * 2);
    p.vq = p.take(grouped ? Mvq * Hv * 2 : 0); p.vqres = p.take(grouped ? Mvq * Hv * 4 : 0);
    for (int i = 0; i < 2; ++i) { p.v[i] = p.take(Mv * Hv * 2); p.vres[i] = p.take(Mv * Hv * 4); }
    p.qkv_v = p.take(Mv * 3 * Wv * 2); p.ctx_v = p.take(Mv * Wv * 2); p.z1_v = p.take(Mv * Hv * 4);
    p.a_v = p.take(Mv * Hv * 2); p.ares_v = p.take(Mv * Hv * 4); p.h_v = p.take(Mv * Iv * 2); p.z2_v = p.take(Mv * Hv * 4);
    const size_t b = (size_t)B;
    p.hw0 = p.take(b * H * 4); p.hv0 = p.take(b * Hv * 4); p.pt = p.take(b * Hb * 4); p.pv = p.take(b * Hb * 4); p.pooled = p.take(b * Hb * 4);
    p.pipe_t[0] = p.take(b * H * 4); p.pipe_t[1] = p.take(b * 512 * 4); p.pipe_t[2] = p.take(b * 256 * 4);
    p.pipe_v[0] = p.take(b * Hv * 4); p.pipe_v[1] = p.take(b * 512 * 4); p.pipe_v[2] = p.take(b * 256 * 4);
    p.prefusion = p.take(b * 512 * 4);
    p.fus[0] = p.take(b * 512 * 4); p.fus[1] = p.take(b * 256 * 4); p.fus[2] = p.take(b * 256 * 4); p.fus[3] = p.take(b * 4);
    return p;
}

// One row layout (cqa_crct_b200.encoder._Rows, packed form)
struct Rows {
    int B, L;
    const int32_t* cu;      // [B+1]
    const int32_t* src;     // [rows]
    const int32_t* n;       // cu + B: the device-side row count
    int hint;               // expected row count (tile shapes only)
};

struct Runner {
    crct_model_s* m;
    crct_stream_t st;
    uint8_t* ws;
    int rc = CRCT_OK;

    template <typename T> T* at(size_t off) const { return reinterpret_cast<T*>(ws + off); }
    bool ok() const { return rc == CRCT_OK; }
    void run(int r) { if (rc == CRCT_OK && r != CRCT_OK) rc = r; }

    const Bound* find(const std::string& name) {
        auto it = m->params.find(name);
        if (it == m->params.end()) {
            if (rc == CRCT_OK) { crct_set_error("crct_forward: parameter '%s' is not bound", name.c_str()); rc = CRCT_ERR_ARG; }
            return nullptr;
        }
        return &it->second;
    }
    const float* P(const std::string& name) { const Bound* b = find(name); return b ? b->w32 : nullptr; }
    const void* W(const std::string& name) {
        const Bound* b = find(name);
        if (b && !b->w16 && rc == CRCT_OK) { crct_set_error("crct_forward: parameter '%s' has no bf16 copy bound", name.c_str()); rc = CRCT_ERR_ARG; }
        return b ? b->w16 : nullptr;
    }

    // D[M,N] = epilogue(x W^T + bias); x [M,K] bf16, W [N,K] bf16 (nn.Linear layout)
    void gemm(const void* x, const void* Wt, const float* bias, void* D, int M, int N, int K, int epi, const void* aux, const Rows& rw) {
        if (!ok()) return;
        crct_gemm_t g;
        memset(&g, 0, sizeof(g));
        g.A = x; g.B = Wt; g.D = D; g.bias = bias; g.aux = aux;
        g.M = M; g.N = N; g.K = K; g.lda = K; g.ldb = K; g.ldd = N; g.ldaux = aux ? N : 0;
        g.epilogue = epi;
        g.a_rows_dev = rw.n; g.rows_hint = rw.hint; g.drop_rows = (epi == CRCT_EPI_BIAS_RES_F32) ? rw.src : nullptr;
        run(crct_gemm_bf16(&g, st));
    }
    void ln(const float* z, const std::string& pre, void* y, float* y32, int rows, int H, const Rows& rw) {
        if (!ok()) return;
        const float* g = P(pre + ".weight");
        const float* b = P(pre + ".bias");
        if (!ok()) return;
        run(crct_layernorm_fwd(z, g, b, y, y32, nullptr, nullptr, rows, H, 1, rw.n, st));
    }
    void attn(const void* q, const void* k, const void* v, int ld, void* out, int ldo, int B, int nh, int dh, int Lq, int Lk,
              const Rows& rq, const Rows& rk) {
        if (!ok()) return;
        crct_attn_fwd_t a;
        memset(&a, 0, sizeof(a));
        a.q = q; a.k = k; a.v = v; a.ldq = a.ldk = a.ldv = ld; a.out = out; a.ldo = ldo;
        a.B = B; a.nh = nh; a.dh = dh; a.Lq = Lq; a.Lk = Lk; a.cu_q = rq.cu; a.cu_k = rk.cu;
        run(crct_attn_fwd(&a, st));
    }

    // intermediate + output blocks (vilbert.py:454-457,467-471): a -> y; z2 keeps the pre-LayerNorm sum (what the heads read)
    void ffn(const void* a, const float* ares, const std::string& pre_i, const std::string& pre_o, void* h, float* z2, void* y, float* yres,
             int M, int H, int I, const Rows& rw) {
        gemm(a, W(pre_i + ".dense.weight"), P(pre_i + ".dense.bias"), h, M, I, H, CRCT_EPI_BIAS_GELU, nullptr, rw);
        gemm(h, W(pre_o + ".dense.weight"), P(pre_o + ".dense.bias"), z2, M, H, I, CRCT_EPI_BIAS_RES_F32, ares, rw);
        ln(z2, pre_o + ".LayerNorm", y, yres, M, H, rw);
    }
    // dense + residual + LayerNorm (vilbert.py:424-428 / 555-559 / 749-756): (ctx, x residual) -> a
    void attn_out(const void* ctx, int Kc, const float* xres, const std::string& pre_dense, const std::string& pre_ln, float* z1, void* a,
                  float* ares, int M, int H, const Rows& rw) {
        gemm(ctx, W(pre_dense + ".weight"), P(pre_dense + ".bias"), z1, M, H, Kc, CRCT_EPI_BIAS_RES_F32, xres, rw);
        ln(z1, pre_ln, a, ares, M, H, rw);
    }
};

// every tensor crct_forward reads (the reference's live parameters; `w16`: the tensor is a tensor-core operand and needs its bf16 copy)
std::vector<std::pair<std::string, bool>> required_params(const crct_config_t& c) {
    std::vector<std::pair<std::string, bool>> r;
    auto lin = [&](const std::string& pre, bool w16) { r.emplace_back(pre + ".weight", w16); r.emplace_back(pre + ".bias", false); };
    auto lnp = [&](const std::string& pre) { r.emplace_back(pre + ".weight", false); r.emplace_back(pre + ".bias", false); };
    const std::string e = "bert.embeddings", ev = "bert.v_embeddings";
    r.emplace_back(e + ".word_embeddings.weight", false); r.emplace_back(e + ".position_embeddings.weight", false);
    r.emplace_back(e + ".plotqa_type_embeddings.weight", false); lin(e + ".txt_location_embeddings", false); lnp(e + ".LayerNorm");
    lin(ev + ".new_image_embeddings", true); lin(ev + ".new_loc_emb", false); r.emplace_back(ev + ".color_emb.weight", false); lnp(ev + ".LayerNorm");
    auto self_layer = [&](const std::string& pre) {
        for (const char* n : {"query", "key", "value"}) lin(pre + ".attention.self." + n, true);
        lin(pre + ".attention.output.dense", true); lnp(pre + ".attention.output.LayerNorm");
        lin(pre + ".intermediate.dense", true); lin(pre + ".output.dense", true); lnp(pre + ".output.LayerNorm");
    };
    for (int i = 0; i < c.num_hidden_layers; ++i) self_layer("bert.encoder.layer." + std::to_string(i));
    for (int i = 0; i < c.v_num_hidden_layers; ++i) self_layer("bert.encoder.v_layer." + std::to_string(i));
    for (int i = 0; i < c.num_connections; ++i) {
        const std::string pre = "bert.encoder.c_layer." + std::to_string(i);
        for (const char* n : {"query1", "key1", "value1", "query2", "key2", "value2"}) lin(pre + ".biattention." + n, true);
        lin(pre + ".biOutput.dense1", true); lnp(pre + ".biOutput.LayerNorm1"); lin(pre + ".biOutput.dense2", true); lnp(pre + ".biOutput.LayerNorm2");
        lin(pre + ".v_intermediate.dense", true); lin(pre + ".v_output.dense", true); lnp(pre + ".v_output.LayerNorm");
        lin(pre + ".t_intermediate.dense", true); lin(pre + ".t_output.dense", true); lnp(pre + ".t_output.LayerNorm");
    }
    lin("bert.t_pooler.dense", false); lin("bert.v_pooler.dense", false); lin("cls.bi_seq_relationship", false);
    for (const char* pipe : {"txt_pipe", "vis_pipe", "fusion"})
        for (int k = 0; k < 4; ++k) lin(std::string("regressor.") + pipe + "." + std::to_string(2 * k), false);
    return r;
}

int check_adjacent(crct_model_s* m, const std::string& pre, const char* const names[3]) {
    const Bound* w[3];
    const Bound* b[3];
    for (int i = 0; i < 3; ++i) {
        auto iw = m->params.find(pre + names[i] + ".weight");
        auto ib = m->params.find(pre + names[i] + ".bias");
        if (iw == m->params.end() || ib == m->params.end()) CRCT_FAIL(CRCT_ERR_ARG, "crct_bind_params: %s%s.{weight,bias} missing", pre.c_str(), names[i]);
        w[i] = &iw->second; b[i] = &ib->second;
        if (!w[i]->w16) CRCT_FAIL(CRCT_ERR_ARG, "crct_bind_params: %s%s.weight needs its bf16 copy", pre.c_str(), names[i]);
    }
    for (int i = 0; i < 2; ++i) {
        if (reinterpret_cast<const uint8_t*>(w[i + 1]->w16) != reinterpret_cast<const uint8_t*>(w[i]->w16) + w[i]->numel * 2 ||
            b[i + 1]->w32 != b[i]->w32 + b[i]->numel)
            CRCT_FAIL(CRCT_ERR_ARG, "crct_bind_params: %s{%s,%s,%s} must be adjacent in memory (fused projection)", pre.c_str(), names[0], names[1], names[2]);
    }
    return CRCT_OK;
}

}  // namespace

extern "C" CRCT_API int crct_create(const crct_config_t* c, crct_handle_t* out) {
    if (!c || !out) CRCT_FAIL(CRCT_ERR_ARG, "crct_create: null pointer");
    if (c->num_connections < 1 || c->num_connections > CRCT_MAX_CONNECTIONS) CRCT_FAIL(CRCT_ERR_ARG, "crct_create: 1 .. %d connection layers", CRCT_MAX_CONNECTIONS);
    const int sizes[3][2] = {{c->hidden_size, c->num_attention_heads}, {c->v_hidden_size, c->v_num_attention_heads}, {c->bi_hidden_size, c->bi_num_attention_heads}};
    for (auto& s : sizes) {
        if (s[0] <= 0 || s[1] <= 0 || s[0] % s[1]) CRCT_FAIL(CRCT_ERR_SHAPE, "crct_create: hidden size %d is not a multiple of %d heads", s[0], s[1]);   // vilbert.py:364-368
        const int dh = s[0] / s[1];
        if (dh != 32 && dh != 48 && dh != 64) CRCT_FAIL(CRCT_ERR_SHAPE, "crct_create: head width %d (kernels: 32, 48, 64)", dh);
        if (s[0] % 8 || s[0] > 1024) CRCT_FAIL(CRCT_ERR_SHAPE, "crct_create: row width %d (multiple of 8, <= 1024)", s[0]);
    }
    crct_model_s* m = new (std::nothrow) crct_model_s;
    if (!m) CRCT_FAIL(CRCT_ERR_ARG, "crct_create: out of host memory");
    m->cfg = *c;
    int v0 = 0, t0 = 0;                                                  // vilbert.py:852-939, as spec.ModelConfig.schedule
    for (int k = 0; k < c->num_connections; ++k) {
        const int v1 = c->v_biattention_id[k], t1 = c->t_biattention_id[k];
        if (v1 < v0 || t1 < t0 || v1 >= c->v_num_hidden_layers || t1 >= c->num_hidden_layers) {
            delete m;
            CRCT_FAIL(CRCT_ERR_ARG, "crct_create: biattention ids must ascend and stay below the layer counts");       // vilbert.py:192-194
        }
        for (int i = v0; i < v1; ++i) m->schedule.emplace_back('v', i);
        for (int i = t0; i < t1; ++i) m->schedule.emplace_back('t', i);
        m->schedule.emplace_back('c', k);
        v0 = v1; t0 = t1;
    }
    for (int i = v0; i < c->v_num_hidden_layers; ++i) m->schedule.emplace_back('v', i);
    for (int i = t0; i < c->num_hidden_layers; ++i) m->schedule.emplace_back('t', i);
    *out = m;
    return CRCT_OK;
}

extern "C" CRCT_API int crct_destroy(crct_handle_t h) {
    delete h;
    return CRCT_OK;
}

extern "C" CRCT_API int crct_bind_params(crct_handle_t h, const char* const* names, const float* const* w32, const void* const* w16,
                                         const size_t* numel, int n) {
    if (!h || !names || !w32 || !numel || n <= 0) CRCT_FAIL(CRCT_ERR_ARG, "crct_bind_params: null pointer");
    h->params.clear();
    for (int i = 0; i < n; ++i) {
        if (!names[i] || !w32[i]) CRCT_FAIL(CRCT_ERR_ARG, "crct_bind_params: entry %d has no name / fp32 pointer", i);
        if (((uintptr_t)w32[i] & 15) || (w16 && w16[i] && ((uintptr_t)w16[i] & 15)))
            CRCT_FAIL(CRCT_ERR_ARG, "crct_bind_params: '%s' is not 16-byte aligned", names[i]);
        h->params[names[i]] = Bound{w32[i], w16 ? w16[i] : nullptr, numel[i]};
    }
    for (const auto& need : required_params(h->cfg)) {
        auto it = h->params.find(need.first);
        if (it == h->params.end()) CRCT_FAIL(CRCT_ERR_ARG, "crct_bind_params: parameter '%s' is missing", need.first.c_str());
        if (need.second && !it->second.w16) CRCT_FAIL(CRCT_ERR_ARG, "crct_bind_params: '%s' needs its bf16 copy (tensor-core operand)", need.first.c_str());
    }
    static const char* const qkv[3] = {"query", "key", "value"};
    static const char* const qkv1[3] = {"query1", "key1", "value1"};
    static const char* const qkv2[3] = {"query2", "key2", "value2"};
    const crct_config_t& c = h->cfg;
    for (int i = 0; i < c.num_hidden_layers; ++i)
        if (int rc = check_adjacent(h, "bert.encoder.layer." + std::to_string(i) + ".attention.self.", qkv)) return rc;
    for (int i = 0; i < c.v_num_hidden_layers; ++i)
        if (int rc = check_adjacent(h, "bert.encoder.v_layer." + std::to_string(i) + ".attention.self.", qkv)) return rc;
    for (int i = 0; i < c.num_connections; ++i) {
        if (int rc = check_adjacent(h, "bert.encoder.c_layer." + std::to_string(i) + ".biattention.", qkv1)) return rc;
        if (int rc = check_adjacent(h, "bert.encoder.c_layer." + std::to_string(i) + ".biattention.", qkv2)) return rc;
    }
    return CRCT_OK;
}

extern "C" CRCT_API size_t crct_workspace_bytes(crct_handle_t h, int B, int Bq, int T, int R) {
    if (!h || B <= 0 || Bq <= 0 || T <= 0 || R <= 0) return 0;
    return make_plan(h->cfg, B, Bq, T, R, true).total + ALIGN;
}

extern "C" CRCT_API int crct_forward(crct_handle_t h, const crct_batch_t* bt, const crct_out_t* out, void* workspace, size_t workspace_bytes,
                                     crct_stream_t stream) {
    if (!h || !bt || !out || !workspace) CRCT_FAIL(CRCT_ERR_ARG, "crct_forward: null pointer");
    if (int rc = crct_device_check()) return rc;
    const crct_config_t& c = h->cfg;
    const int B = bt->B, Bq = bt->Bq, T = bt->T, R = bt->R;
    if (B <= 0 || Bq <= 0 || T <= 0 || R <= 0) CRCT_FAIL(CRCT_ERR_SHAPE, "crct_forward: empty batch");
    const bool grouped = bt->group != nullptr;
    if (!grouped && Bq != B) CRCT_FAIL(CRCT_ERR_ARG, "crct_forward: %d text rows but %d visual rows (pass `group` to share visual inputs between candidates)", B, Bq);
    if (!bt->tokens || !bt->segments || !bt->loc || !bt->attention_mask || !bt->image_feat || !bt->image_loc || !bt->image_target ||
        !bt->image_mask || !bt->R4)
        CRCT_FAIL(CRCT_ERR_ARG, "crct_forward: null input tensor");
    if (!out->logits || !out->reg_pred || !out->reg_loss || !out->reg_l1 || !out->reg_dist || !out->scalars)
        CRCT_FAIL(CRCT_ERR_ARG, "crct_forward: null output tensor");
    if (T > c.max_position_embeddings) CRCT_FAIL(CRCT_ERR_SHAPE, "crct_forward: T = %d exceeds max_position_embeddings = %d", T, c.max_position_embeddings);
    const Plan p = make_plan(c, B, Bq, T, R, grouped);
    uint8_t* ws = reinterpret_cast<uint8_t*>(((uintptr_t)workspace + ALIGN - 1) / ALIGN * ALIGN);
    if ((size_t)(ws - reinterpret_cast<uint8_t*>(workspace)) + p.total > workspace_bytes)
        CRCT_FAIL(CRCT_ERR_ARG, "crct_forward: workspace of %zu bytes, %zu needed (crct_workspace_bytes)", workspace_bytes, p.total + ALIGN);

    Runner r{h, stream, ws};
    const int H = c.hidden_size, Hv = c.v_hidden_size, Hb = c.bi_hidden_size, I = c.intermediate_size, Iv = c.v_intermediate_size, F = c.v_feature_size;
    const int Mt = B * T, Mvq = Bq * R, Mv = B * R;
    auto hint = [](float fill, int rows) { return fill > 0.f ? (int)(fill * (float)rows) : 0; };

    // ---- packed rows (csrc/varlen.cu): compact the valid tokens / regions once; no additive masks from here on
    int32_t* cu_t = r.at<int32_t>(p.cu_t);
    int32_t* cu_vq = r.at<int32_t>(p.cu_vq);
    int32_t* cu_v = r.at<int32_t>(p.cu_v);
    r.run(crct_row_map(bt->attention_mask, bt->attention_mask_kind, B, T, cu_t, r.at<int32_t>(p.src_t), stream));
    r.run(crct_row_map(bt->image_mask, bt->image_mask_kind, Bq, R, cu_vq, r.at<int32_t>(p.src_vq), stream));
    const Rows rt{B, T, cu_t, r.at<int32_t>(p.src_t), cu_t + B, hint(bt->text_fill, Mt)};
    const Rows rvq{Bq, R, cu_vq, r.at<int32_t>(p.src_vq), cu_vq + Bq, hint(bt->region_fill, Mvq)};
    Rows rv = rvq;
    if (grouped) {                       // candidate n shares the packed region rows of question group[n] (f3)
        r.run(crct_group_map(cu_vq, bt->group, B, cu_v, r.at<int32_t>(p.src_v), stream));
        rv = Rows{B, R, cu_v, r.at<int32_t>(p.src_v), cu_v + B, hint(bt->region_fill, Mv)};
    }

    // ---- embeddings (vilbert.py:1412-1413)
    int ti = 0, vi = 0;                  // which of the two hidden-state buffers holds the stream's current state
    {
        const std::string e = "bert.embeddings";
        crct_embed_text_t a;
        memset(&a, 0, sizeof(a));
        a.ids = bt->tokens; a.types = bt->segments; a.loc = bt->loc;
        a.word = r.P(e + ".word_embeddings.weight"); a.pos = r.P(e + ".position_embeddings.weight"); a.type = r.P(e + ".plotqa_type_embeddings.weight");
        a.w_loc = r.P(e + ".txt_location_embeddings.weight"); a.b_loc = r.P(e + ".txt_location_embeddings.bias");
        a.gamma = r.P(e + ".LayerNorm.weight"); a.beta = r.P(e + ".LayerNorm.bias");
        a.y = r.at<void>(p.t[0]); a.y32 = r.at<float>(p.tres[0]);
        a.B = B; a.T = T; a.H = H; a.max_pos = c.max_position_embeddings;
        a.src_row = rt.src; a.rows_dev = rt.n;
        if (r.ok()) r.run(crct_embed_text_fwd(&a, stream));
    }
    {
        const std::string e = "bert.v_embeddings";
        if (r.ok()) r.run(crct_softmax_rows(bt->image_feat, r.at<void>(p.probs), Mvq, F, rvq.src, rvq.n, stream));
        r.gemm(r.at<void>(p.probs), r.W(e + ".new_image_embeddings.weight"), r.P(e + ".new_image_embeddings.bias"), r.at<void>(p.gimg), Mvq, Hv, F,
               CRCT_EPI_BIAS, nullptr, rvq);
        crct_embed_vis_t a;
        memset(&a, 0, sizeof(a));
        a.g = r.at<void>(p.gimg); a.box = bt->image_loc; a.cls = bt->image_target;
        a.w_loc = r.P(e + ".new_loc_emb.weight"); a.b_loc = r.P(e + ".new_loc_emb.bias"); a.color = r.P(e + ".color_emb.weight");
        a.gamma = r.P(e + ".LayerNorm.weight"); a.beta = r.P(e + ".LayerNorm.bias");
        a.y = r.at<void>(grouped ? p.vq : p.v[0]); a.y32 = r.at<float>(grouped ? p.vqres : p.vres[0]);
        a.rows = Mvq; a.H = Hv; a.src_row = rvq.src; a.rows_dev = rvq.n;
        if (r.ok()) r.run(crct_embed_vis_fwd(&a, stream));
        if (grouped && r.ok()) {         // the visual embedding depends on the image only: computed once per question, fanned out here
            r.run(crct_gather_rows(r.at<void>(p.vq), rv.src, r.at<void>(p.v[0]), Mv, (long long)Hv * 2, rv.n, stream));
            r.run(crct_gather_rows(r.at<void>(p.vqres), rv.src, r.at<void>(p.vres[0]), Mv, (long long)Hv * 4, rv.n, stream));
        }
    }

    // ---- encoder (vilbert.py:852-939)
    static const char* const QKV[3] = {"query", "key", "value"};
    auto self_layer = [&](bool text, int idx) {
        const int Hs = text ? H : Hv, Is = text ? I : Iv, nh = text ? c.num_attention_heads : c.v_num_attention_heads, L = text ? T : R;
        const int M = text ? Mt : Mv;
        const Rows& rw = text ? rt : rv;
        int& cur = text ? ti : vi;
        const size_t* hs = text ? p.t : p.v;
        const size_t* hres = text ? p.tres : p.vres;
        const std::string pre = std::string(text ? "bert.encoder.layer." : "bert.encoder.v_layer.") + std::to_string(idx);
        uint8_t* qkv = r.at<uint8_t>(text ? p.qkv_t : p.qkv_v);
        void* ctx = r.at<void>(text ? p.ctx_t : p.ctx_v);
        r.gemm(r.at<void>(hs[cur]), r.W(pre + ".attention.self." + QKV[0] + ".weight"), r.P(pre + ".attention.self." + QKV[0] + ".bias"), qkv, M, 3 * Hs, Hs,
               CRCT_EPI_BIAS, nullptr, rw);
        r.attn(qkv, qkv + (size_t)Hs * 2, qkv + (size_t)2 * Hs * 2, 3 * Hs, ctx, Hs, rw.B, nh, Hs / nh, L, L, rw, rw);
        r.attn_out(ctx, Hs, r.at<float>(hres[cur]), pre + ".attention.output.dense", pre + ".attention.output.LayerNorm", r.at<float>(text ? p.z1_t : p.z1_v),
                   r.at<void>(text ? p.a_t : p.a_v), r.at<float>(text ? p.ares_t : p.ares_v), M, Hs, rw);
        r.ffn(r.at<void>(text ? p.a_t : p.a_v), r.at<float>(text ? p.ares_t : p.ares_v), pre + ".intermediate", pre + ".output",
              r.at<void>(text ? p.h_t : p.h_v), r.at<float>(text ? p.z2_t : p.z2_v), r.at<void>(hs[cur ^ 1]), r.at<float>(hres[cur ^ 1]), M, Hs, Is, rw);
        cur ^= 1;
    };
    auto co_layer = [&](int idx) {       // BertConnectionLayer (vilbert.py:774-788); stream 1 = visual, stream 2 = text
        const std::string pre = "bert.encoder.c_layer." + std::to_string(idx);
        const int nh = c.bi_num_attention_heads, dh = Hb / nh, ld = 3 * Hb;
        uint8_t* qkv1 = r.at<uint8_t>(p.qkv_v);
        uint8_t* qkv2 = r.at<uint8_t>(p.qkv_t);
        r.gemm(r.at<void>(p.v[vi]), r.W(pre + ".biattention.query1.weight"), r.P(pre + ".biattention.query1.bias"), qkv1, Mv, ld, Hv, CRCT_EPI_BIAS, nullptr, rv);
        r.gemm(r.at<void>(p.t[ti]), r.W(pre + ".biattention.query2.weight"), r.P(pre + ".biattention.query2.bias"), qkv2, Mt, ld, H, CRCT_EPI_BIAS, nullptr, rt);
        // biOutput is called with crossed arguments (vilbert.py:780): visual <- ctx2 via dense1 / LayerNorm1, text <- ctx1 via dense2 / LayerNorm2
        void* ctx2 = r.at<void>(p.ctx_v);        // visual queries over text keys / values
        r.attn(qkv1, qkv2 + (size_t)Hb * 2, qkv2 + (size_t)2 * Hb * 2, ld, ctx2, Hb, B, nh, dh, R, T, rv, rt);
        r.attn_out(ctx2, Hb, r.at<float>(p.vres[vi]), pre + ".biOutput.dense1", pre + ".biOutput.LayerNorm1", r.at<float>(p.z1_v), r.at<void>(p.a_v),
                   r.at<float>(p.ares_v), Mv, Hv, rv);
        r.ffn(r.at<void>(p.a_v), r.at<float>(p.ares_v), pre + ".v_intermediate", pre + ".v_output", r.at<void>(p.h_v), r.at<float>(p.z2_v),
              r.at<void>(p.v[vi ^ 1]), r.at<float>(p.vres[vi ^ 1]), Mv, Hv, Iv, rv);
        void* ctx1 = r.at<void>(p.ctx_t);        // text queries over visual keys / values
        r.attn(qkv2, qkv1 + (size_t)Hb * 2, qkv1 + (size_t)2 * Hb * 2, ld, ctx1, Hb, B, nh, dh, T, R, rt, rv);
        r.attn_out(ctx1, Hb, r.at<float>(p.tres[ti]), pre + ".biOutput.dense2", pre + ".biOutput.LayerNorm2", r.at<float>(p.z1_t), r.at<void>(p.a_t),
                   r.at<float>(p.ares_t), Mt, H, rt);
        r.ffn(r.at<void>(p.a_t), r.at<float>(p.ares_t), pre + ".t_intermediate", pre + ".t_output", r.at<void>(p.h_t), r.at<float>(p.z2_t),
              r.at<void>(p.t[ti ^ 1]), r.at<float>(p.tres[ti ^ 1]), Mt, H, I, rt);
        vi ^= 1; ti ^= 1;
    };
    std::string ln_t, ln_v;              // LayerNorm that follows the stream's last pre-LayerNorm sum (z2_t / z2_v)
    for (const auto& item : h->schedule) {
        if (!r.ok()) break;
        if (item.first == 't') { self_layer(true, item.second); ln_t = "bert.encoder.layer." + std::to_string(item.second) + ".output.LayerNorm"; }
        else if (item.first == 'v') { self_layer(false, item.second); ln_v = "bert.encoder.v_layer." + std::to_string(item.second) + ".output.LayerNorm"; }
        else {
            co_layer(item.second);
            ln_t = "bert.encoder.c_layer." + std::to_string(item.second) + ".t_output.LayerNorm";
            ln_v = "bert.encoder.c_layer." + std::to_string(item.second) + ".v_output.LayerNorm";
        }
    }
    if (!r.ok()) return r.rc;

    // ---- heads in fp32 (vilbert.py:955-976, 1048-1062; regressor.py:36-42; losses vilbert.py:1586-1657)
    // first token / region of every sample (row cu[b]): LayerNorm of the last pre-LayerNorm sum in fp32
    float* hw0 = r.at<float>(p.hw0);
    float* hv0 = r.at<float>(p.hv0);
    r.run(crct_layernorm_rows_f32(r.at<float>(p.z2_t), r.P(ln_t + ".weight"), r.P(ln_t + ".bias"), rt.cu, T, hw0, B, H, stream));
    if (r.ok()) r.run(crct_layernorm_rows_f32(r.at<float>(p.z2_v), r.P(ln_v + ".weight"), r.P(ln_v + ".bias"), rv.cu, R, hv0, B, Hv, stream));
    auto lin = [&](const float* x, int ldx, const std::string& name, int N, int K, int act, float* o, int ldc) {
        crct_linear_t a;
        memset(&a, 0, sizeof(a));
        a.A = x; a.sa_m = ldx; a.sa_k = 1; a.B = r.P(name + ".weight"); a.sb_k = 1; a.sb_n = K; a.C = o; a.ldc = ldc; a.bias = r.P(name + ".bias");
        a.M = B; a.N = N; a.K = K; a.act = act;
        return a;
    };
    float* prefusion = r.at<float>(p.prefusion);          // cat((hv, hw), -1), regressor.py:40
    const float* xt = hw0;
    const float* xv = hv0;
    int kt = H, kv = Hv;
    const int pipe_out[4][2] = {{H, Hv}, {512, 512}, {256, 256}, {256, 256}};
    std::vector<crct_linear_t> probs;
    probs.push_back(lin(hw0, H, "bert.t_pooler.dense", Hb, H, 1, r.at<float>(p.pt), Hb));
    probs.push_back(lin(hv0, Hv, "bert.v_pooler.dense", Hb, Hv, 1, r.at<float>(p.pv), Hb));
    for (int i = 0; i < 4 && r.ok(); ++i) {
        const bool last = i == 3;
        const std::string idx = std::to_string(2 * i);
        float* ov = last ? prefusion : r.at<float>(p.pipe_v[i]);
        float* ot = last ? prefusion + 256 : r.at<float>(p.pipe_t[i]);
        const int nv = pipe_out[i][1], nt = pipe_out[i][0];
        probs.push_back(lin(xv, kv, "regressor.vis_pipe." + idx, nv, kv, last ? 0 : 2, ov, last ? 512 : nv));
        probs.push_back(lin(xt, kt, "regressor.txt_pipe." + idx, nt, kt, last ? 0 : 2, ot, last ? 512 : nt));
        if (r.ok()) r.run(crct_linear_f32_batched(probs.data(), (int)probs.size(), stream));
        probs.clear();
        xv = ov; xt = ot; kv = nv; kt = nt;
    }
    if (r.ok()) r.run(crct_pool_mul_fwd(r.at<float>(p.pt), r.at<float>(p.pv), r.at<float>(p.pooled), B * Hb, 0.f, 0, nullptr, stream));
    probs.push_back(lin(r.at<float>(p.pooled), Hb, "cls.bi_seq_relationship", 2, Hb, 0, out->logits, 2));
    const int fus_out[4] = {512, 256, 256, 1};
    const float* xf = prefusion;
    int kf = 512;
    for (int i = 0; i < 4 && r.ok(); ++i) {
        probs.push_back(lin(xf, kf, "regressor.fusion." + std::to_string(2 * i), fus_out[i], kf, i == 3 ? 3 : 2, r.at<float>(p.fus[i]), fus_out[i]));
        if (r.ok()) r.run(crct_linear_f32_batched(probs.data(), (int)probs.size(), stream));
        probs.clear();
        xf = r.at<float>(p.fus[i]); kf = fus_out[i];
    }
    if (!r.ok()) return r.rc;
    crct_loss_t ls;
    memset(&ls, 0, sizeof(ls));
    ls.logits = out->logits; ls.reg = r.at<float>(p.fus[3]); ls.labels = nullptr; ls.R = bt->R4;
    ls.reg_pred = out->reg_pred; ls.reg_loss = out->reg_loss; ls.reg_l1 = out->reg_l1; ls.reg_dist = out->reg_dist; ls.scalars = out->scalars;
    ls.B = B; ls.l1 = c.l1; ls.zero_impossible = 0;      // loss kind 'L1' at evaluation (encoder_decorator.py:106)
    ls.unit_grads = 1; ls.tol_margin = c.tol_margin; ls.nsp_coeff = 1.f; ls.reg_coeff = 1.f;
    r.run(crct_hybrid_loss(&ls, stream));
    return r.rc;
}
