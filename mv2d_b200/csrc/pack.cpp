// mv2d_pack_weights / mv2d_pack_neck: one-time re-layout of the reference state_dict into the buffers the kernels read.
// Pure host code (no CUDA calls): the caller hands HOST fp32 tensors under their reference key names, gets a HOST arena
// image + a directory back and uploads the image with one copy; the weight structs are filled with device pointers
// computed from `device_base`.
//
// Derived matrices are formed in fp64 and rounded once to fp32:
//  * ca_q_w / ca_q_b  -- cross-attention query side with the key projection absorbed: per head h,
//    scale * Wk_h^T Wq_h ([256 x 256]) and scale * Wk_h^T bq_h, stacked to [2048,256] / [2048].  The key bias only adds a
//    per-(query, head) constant to the logits, which softmax cancels (utils/petr_transformer.py:503-508 -> torch
//    MultiheadAttention).
//  * ca_o_w / ca_o_b  -- output side with the value projection absorbed: per head Wo[:, 32h:32h+32] Wv_h stacked along K
//    to [256,2048], bias Wo bv + bo (probabilities sum to one).
//  * xa_* -- the plain per-role projections of the key-stationary form: 1/sqrt(32) folded into the query side, the
//    value bias moved behind the softmax.
//  * l0.sa_const -- layer 0's self-attention output (target = 0, cross_attention_head.py:32, so every value row is bv).
//  * the 3x3 convolutions are stored K-major with K ordered (ky, kx, c_in).
//  * every tensor-core operand is pre-split into TF32 hi + lo (or pre-rounded to TF32 for the single-pass PE MLPs).
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/mv2d_b200.h"

namespace mv2d { void set_error(const char* fmt, ...); }
#define MV2D_CHECK_ARG(cond, msg) do { if (!(cond)) { mv2d::set_error("%s", msg); return -1; } } while (0)

namespace {

constexpr int E = 256, HEADS = 8, HD = 32, FF = 2048;

inline float round_tf32_host(float x) {      // cvt.rna.tf32.f32: round to nearest, ties away from zero, on the bit pattern
    uint32_t b;
    memcpy(&b, &x, 4);
    b = (b + 0x1000u) & ~0x1FFFu;
    float y;
    memcpy(&y, &b, 4);
    return y;
}

struct Packer {
    const Mv2dNamedTensor* sd;
    int n_sd;
    float* arena;                 // host image, nullable = size query
    int64_t cap, used = 0;
    const char* dev;
    Mv2dPackedEntry* dir;
    int dir_cap, n_dir = 0;
    std::string err;

    const Mv2dNamedTensor* find(const std::string& key, int64_t numel) {
        for (int i = 0; i < n_sd; ++i) {
            const char* nm = sd[i].name;
            if (!nm) continue;
            if (!strncmp(nm, "roi_head.", 9)) nm += 9;
            if (key == nm) {
                if (sd[i].numel != numel || !sd[i].data) {
                    if (err.empty()) err = "pack_weights: '" + key + "' has " + std::to_string(sd[i].numel) + " elements, expected " + std::to_string(numel);
                    return nullptr;
                }
                return &sd[i];
            }
        }
        if (err.empty()) err = "pack_weights: missing key '" + key + "'";
        return nullptr;
    }
    // reserve [numel] floats under `name`; returns the host pointer to fill (scratch when only sizing) and the device pointer
    float* put(const std::string& name, int64_t numel, const float** dptr) {
        const int64_t off = used;
        used += (numel * 4 + 255) / 256 * 256;
        if (dptr) *dptr = reinterpret_cast<const float*>(dev + off);
        if (dir && n_dir < dir_cap) {
            Mv2dPackedEntry& e = dir[n_dir];
            memset(&e, 0, sizeof(e));
            strncpy(e.name, name.c_str(), sizeof(e.name) - 1);
            e.offset = off;
            e.numel = numel;
        }
        ++n_dir;
        if (!arena) return nullptr;
        if (used > cap) {
            if (err.empty()) err = "pack_weights: arena too small";
            return nullptr;
        }
        return arena + off / 4;
    }
    // directory entry for a sub-range of an existing buffer (no new storage)
    void alias(const std::string& name, int64_t offset, int64_t numel) {
        if (dir && n_dir < dir_cap) {
            Mv2dPackedEntry& e = dir[n_dir];
            memset(&e, 0, sizeof(e));
            strncpy(e.name, name.c_str(), sizeof(e.name) - 1);
            e.offset = offset;
            e.numel = numel;
        }
        ++n_dir;
    }
    void put_copy(const std::string& name, const float* src, int64_t numel, const float** dptr) {
        float* d = put(name, numel, dptr);
        if (d && src) memcpy(d, src, numel * 4);
    }
    void put_key(const std::string& name, const std::string& key, int64_t numel, const float** dptr) {
        const Mv2dNamedTensor* t = arena ? find(key, numel) : nullptr;
        put_copy(name, t ? t->data : nullptr, numel, dptr);
    }
    void put_rounded(const std::string& name, const std::string& key, int64_t numel, const float** dptr) {
        const Mv2dNamedTensor* t = arena ? find(key, numel) : nullptr;
        float* d = put(name, numel, dptr);
        if (d && t) for (int64_t i = 0; i < numel; ++i) d[i] = round_tf32_host(t->data[i]);
    }
    // w = hi + lo, both TF32-representable
    void put_split(const std::string& hi_name, const std::string& lo_name, const float* src, int64_t numel, const float** dhi, const float** dlo) {
        float* h = put(hi_name, numel, dhi);
        float* l = put(lo_name, numel, dlo);
        if (h && l && src)
            for (int64_t i = 0; i < numel; ++i) {
                const float hi = round_tf32_host(src[i]);
                h[i] = hi;
                l[i] = round_tf32_host(src[i] - hi);
            }
    }
};

// [co, ci, 3, 3] -> [co, (ky, kx, ci)]
void conv3x3_kmajor(const float* w, int co, int ci, std::vector<float>& out) {
    out.resize((size_t)co * 9 * ci);
    for (int o = 0; o < co; ++o)
        for (int c = 0; c < ci; ++c)
            for (int t = 0; t < 9; ++t) out[((size_t)o * 9 + t) * ci + c] = w[((size_t)o * ci + c) * 9 + t];
}

int finish(Packer& P, int* n_dir, int64_t* bytes) {
    if (n_dir) *n_dir = P.n_dir;
    if (bytes) *bytes = P.used;
    if (!P.err.empty()) {
        mv2d::set_error("%s", P.err.c_str());
        return -1;
    }
    if (P.dir && P.n_dir > P.dir_cap) {
        mv2d::set_error("pack_weights: directory too small");
        return -1;
    }
    return 0;
}

int pack_all(Packer& P, int L, int fold_first, Mv2dLayerWeights* layers, Mv2dBranchWeights* br) {
    const bool fill = P.arena != nullptr;
    const float* dp;
    const std::string pe = "position_encoding.", qg = "query_generator.", bh = "bbox_head.";
    // K1 position-embedding MLPs: single-pass TF32 operands
    P.put_rounded("w_pos0", pe + "position_encoder.0.weight", 1024 * 192, &dp); P.put_key("b_pos0", pe + "position_encoder.0.bias", 1024, &dp);
    P.put_rounded("w_pos2", pe + "position_encoder.2.weight", 256 * 1024, &dp); P.put_key("b_pos2", pe + "position_encoder.2.bias", 256, &dp);
    P.put_rounded("w_adapt0", pe + "adapt_pos3d.0.weight", 1024 * 384, &dp); P.put_key("b_adapt0", pe + "adapt_pos3d.0.bias", 1024, &dp);
    P.put_rounded("w_adapt2", pe + "adapt_pos3d.2.weight", 256 * 1024, &dp); P.put_key("b_adapt2", pe + "adapt_pos3d.2.bias", 256, &dp);
    P.put_rounded("w_se_reduce", pe + "fpe.conv_reduce.weight", 256 * 256, &dp); P.put_key("b_se_reduce", pe + "fpe.conv_reduce.bias", 256, &dp);
    P.put_rounded("w_se_expand", pe + "fpe.conv_expand.weight", 256 * 256, &dp); P.put_key("b_se_expand", pe + "fpe.conv_expand.bias", 256, &dp);
    // K3 query generator
    {
        std::vector<float> km;
        const Mv2dNamedTensor* t = fill ? P.find(qg + "shared_convs.0.conv.weight", (int64_t)E * E * 9) : nullptr;
        if (t) conv3x3_kmajor(t->data, E, E, km);
        P.put_split("w_conv", "w_conv_lo", t ? km.data() : nullptr, (int64_t)E * 9 * E, &dp, &dp);
        P.put_key("b_conv", qg + "shared_convs.0.conv.bias", E, &dp);
    }
    struct Fc { const char* name; std::string key; int out, in, in_pad; };
    const Fc fcs[] = {{"fc", qg + "shared_fcs.0", 1024, 256, 256},           {"enc0", qg + "extra_enc.0", 512, 1040, 1056},
                      {"enc2", qg + "extra_enc.2", 256, 512, 512},         {"center", qg + "fc_center", 3, 256, 256},
                      {"qe0", bh + "query_embedding.0", 256, 384, 384},     {"qe2", bh + "query_embedding.2", 256, 256, 256}};
    for (const Fc& f : fcs) {
        const std::string wn = std::string("w_") + f.name, bn = std::string("b_") + f.name;
        const Mv2dNamedTensor* t = fill ? P.find(f.key + ".weight", (int64_t)f.out * f.in) : nullptr;
        float* w = P.put(wn, (int64_t)f.out * f.in_pad, &dp);
        if (w && t)
            for (int o = 0; o < f.out; ++o) {      // extra_enc.0: K padded 1040 -> 1056 with zeros (16-byte rows for TMA)
                memcpy(w + (size_t)o * f.in_pad, t->data + (size_t)o * f.in, (size_t)f.in * 4);
                for (int k = f.in; k < f.in_pad; ++k) w[(size_t)o * f.in_pad + k] = 0.f;
            }
        P.put_key(bn, f.key + ".bias", f.out, &dp);
        if (strcmp(f.name, "center")) P.put_split(wn + "_hi", wn + "_lo", w, (int64_t)f.out * f.in_pad, &dp, &dp);   // 3xTF32 FC chain for batches
    }
    // K4 decoder layers
    // the plain K / V projections of all layers form ONE [L*2*256, 256] operand (rows (2 l + side) * 256 ..), hi and lo:
    // the persistent projection kernel (kvproj.cu) addresses every matrix through one tensor map
    const float *kv_hi_dev = nullptr, *kv_lo_dev = nullptr;
    const int64_t kv_hi_off = P.used;
    float* kv_hi = P.put("xa_kv_w", (int64_t)L * 2 * E * E, &kv_hi_dev);
    const int64_t kv_lo_off = P.used;
    float* kv_lo = P.put("xa_kv_w_lo", (int64_t)L * 2 * E * E, &kv_lo_dev);
    const double scale = 1.0 / std::sqrt((double)HD);
    std::vector<float> qw((size_t)HEADS * E * E), qb((size_t)HEADS * E), ow((size_t)E * HEADS * E), ob(E), xq((size_t)E * E), xqb(E), xob(E), sac(E);
    for (int l = 0; l < L; ++l) {
        Mv2dLayerWeights lw;
        memset(&lw, 0, sizeof(lw));
        const std::string p = bh + "transformer.decoder.layers." + std::to_string(l) + ".", n = "l" + std::to_string(l) + ".";
        const Mv2dNamedTensor *sin_w = nullptr, *sin_b = nullptr, *sout_w = nullptr, *sout_b = nullptr, *cin_w = nullptr, *cin_b = nullptr,
                              *cout_w = nullptr, *cout_b = nullptr, *f1 = nullptr, *f2 = nullptr;
        if (fill) {
            sin_w = P.find(p + "attentions.0.attn.in_proj_weight", 3 * E * E); sin_b = P.find(p + "attentions.0.attn.in_proj_bias", 3 * E);
            sout_w = P.find(p + "attentions.0.attn.out_proj.weight", E * E);   sout_b = P.find(p + "attentions.0.attn.out_proj.bias", E);
            cin_w = P.find(p + "attentions.1.attn.in_proj_weight", 3 * E * E); cin_b = P.find(p + "attentions.1.attn.in_proj_bias", 3 * E);
            cout_w = P.find(p + "attentions.1.attn.out_proj.weight", E * E);   cout_b = P.find(p + "attentions.1.attn.out_proj.bias", E);
            f1 = P.find(p + "ffns.0.layers.0.0.weight", (int64_t)FF * E);      f2 = P.find(p + "ffns.0.layers.1.weight", (int64_t)E * FF);
        }
        const bool ok = sin_w && sin_b && sout_w && sout_b && cin_w && cin_b && cout_w && cout_b && f1 && f2;
        if (ok) {
            const float *wq = cin_w->data, *wk = wq + E * E, *wv = wk + E * E, *bq = cin_b->data, *bv = bq + 2 * E, *wo = cout_w->data;
            for (int h = 0; h < HEADS; ++h)
                for (int i = 0; i < E; ++i) {          // key-input dim
                    double sb = 0;
                    for (int d = 0; d < HD; ++d) sb += (double)wk[(h * HD + d) * E + i] * bq[h * HD + d];
                    qb[h * E + i] = (float)(scale * sb);
                    for (int j = 0; j < E; ++j) {      // query-input dim
                        double s = 0;
                        for (int d = 0; d < HD; ++d) s += (double)wk[(h * HD + d) * E + i] * wq[(h * HD + d) * E + j];
                        qw[((size_t)h * E + i) * E + j] = (float)(scale * s);
                    }
                }
            for (int o = 0; o < E; ++o) {
                for (int h = 0; h < HEADS; ++h)
                    for (int j = 0; j < E; ++j) {      // memory dim
                        double s = 0;
                        for (int d = 0; d < HD; ++d) s += (double)wo[o * E + h * HD + d] * wv[(h * HD + d) * E + j];
                        ow[(size_t)o * HEADS * E + h * E + j] = (float)s;
                    }
                double s = cout_b->data[o];
                for (int k = 0; k < E; ++k) s += (double)wo[o * E + k] * bv[k];
                ob[o] = xob[o] = (float)s;
                xqb[o] = (float)(scale * bq[o]);
                double c = sout_b->data[o];
                const float* sbv = sin_b->data + 2 * E;
                for (int k = 0; k < E; ++k) c += (double)sout_w->data[o * E + k] * sbv[k];
                sac[o] = (float)c;
            }
            for (int i = 0; i < E * E; ++i) xq[i] = (float)(scale * wq[i]);
        }
        P.put_copy(n + "sa_in_w", ok ? sin_w->data : nullptr, 3 * E * E, &lw.sa_in_w);
        P.put_copy(n + "sa_in_b", ok ? sin_b->data : nullptr, 3 * E, &lw.sa_in_b);
        P.put_copy(n + "sa_out_w", ok ? sout_w->data : nullptr, E * E, &lw.sa_out_w);
        P.put_copy(n + "sa_out_b", ok ? sout_b->data : nullptr, E, &lw.sa_out_b);
        P.put_split(n + "ca_q_w", n + "ca_q_w_lo", ok ? qw.data() : nullptr, (int64_t)HEADS * E * E, &lw.ca_q_w, &lw.ca_q_w_lo);
        P.put_split(n + "ca_o_w", n + "ca_o_w_lo", ok ? ow.data() : nullptr, (int64_t)E * HEADS * E, &lw.ca_o_w, &lw.ca_o_w_lo);
        P.put_split(n + "ffn_w1", n + "ffn_w1_lo", ok ? f1->data : nullptr, (int64_t)FF * E, &lw.ffn_w1, &lw.ffn_w1_lo);
        P.put_split(n + "ffn_w2", n + "ffn_w2_lo", ok ? f2->data : nullptr, (int64_t)E * FF, &lw.ffn_w2, &lw.ffn_w2_lo);
        P.put_copy(n + "xa_q_w", ok ? xq.data() : nullptr, E * E, &lw.xa_q_w);
        P.put_copy(n + "xa_q_b", ok ? xqb.data() : nullptr, E, &lw.xa_q_b);
        for (int side = 0; side < 2; ++side) {
            const int64_t sub = (int64_t)(2 * l + side) * E * E;
            const float* src = ok ? cin_w->data + (1 + side) * E * E : nullptr;
            if (kv_hi && kv_lo && src)
                for (int i = 0; i < E * E; ++i) {
                    const float hi = round_tf32_host(src[i]);
                    kv_hi[sub + i] = hi;
                    kv_lo[sub + i] = round_tf32_host(src[i] - hi);
                }
            (side ? lw.xa_v_w : lw.xa_k_w) = kv_hi_dev + sub;
            (side ? lw.xa_v_w_lo : lw.xa_k_w_lo) = kv_lo_dev + sub;
            P.alias(n + (side ? "xa_v_w" : "xa_k_w"), kv_hi_off + sub * 4, E * E);
            P.alias(n + (side ? "xa_v_w_lo" : "xa_k_w_lo"), kv_lo_off + sub * 4, E * E);
        }
        P.put_split(n + "sa_in_w_hi", n + "sa_in_w_lo", ok ? sin_w->data : nullptr, 3 * E * E, &lw.sa_in_w_hi, &lw.sa_in_w_lo);
        P.put_split(n + "sa_out_w_hi", n + "sa_out_w_lo", ok ? sout_w->data : nullptr, E * E, &lw.sa_out_w_hi, &lw.sa_out_w_lo);
        P.put_split(n + "xa_q_w_hi", n + "xa_q_w_lo", ok ? xq.data() : nullptr, E * E, &lw.xa_q_w_hi, &lw.xa_q_w_lo);
        P.put_split(n + "xa_o_w_hi", n + "xa_o_w_lo", ok ? cout_w->data : nullptr, E * E, &lw.xa_o_w_hi, &lw.xa_o_w_lo);
        P.put_copy(n + "xa_k_raw", ok ? cin_w->data + E * E : nullptr, E * E, &lw.xa_k_raw);
        P.put_copy(n + "xa_v_raw", ok ? cin_w->data + 2 * E * E : nullptr, E * E, &lw.xa_v_raw);
        P.put_copy(n + "xa_o_w", ok ? cout_w->data : nullptr, E * E, &lw.xa_o_w);
        P.put_copy(n + "xa_o_b", ok ? xob.data() : nullptr, E, &lw.xa_o_b);
        P.put_copy(n + "ca_q_b", ok ? qb.data() : nullptr, HEADS * E, &lw.ca_q_b);
        P.put_copy(n + "ca_o_b", ok ? ob.data() : nullptr, E, &lw.ca_o_b);
        P.put_key(n + "ffn_b1", p + "ffns.0.layers.0.0.bias", FF, &lw.ffn_b1);
        P.put_key(n + "ffn_b2", p + "ffns.0.layers.1.bias", E, &lw.ffn_b2);
        for (int k = 0; k < 3; ++k) {
            P.put_key(n + "ln_g" + std::to_string(k), p + "norms." + std::to_string(k) + ".weight", E, &lw.ln_g[k]);
            P.put_key(n + "ln_b" + std::to_string(k), p + "norms." + std::to_string(k) + ".bias", E, &lw.ln_b[k]);
        }
        if (l == 0 && fold_first) P.put_copy("l0.sa_const", ok ? sac.data() : nullptr, E, &lw.sa_const);
        if (layers) layers[l] = lw;
    }
    // K5 branches, stacked over layers
    Mv2dBranchWeights b;
    memset(&b, 0, sizeof(b));
    struct Br { const char* field; const char* key; int numel; const float** slot; const float **hi, **lo; };
    const Br brs[] = {
        {"cls_w0", "cls_branches.%d.0.weight", E * E, &b.cls_w0, &b.cls_w0_hi, &b.cls_w0_lo}, {"cls_b0", "cls_branches.%d.0.bias", E, &b.cls_b0, nullptr, nullptr},
        {"cls_g0", "cls_branches.%d.1.weight", E, &b.cls_g0, nullptr, nullptr},             {"cls_be0", "cls_branches.%d.1.bias", E, &b.cls_be0, nullptr, nullptr},
        {"cls_w1", "cls_branches.%d.3.weight", E * E, &b.cls_w1, &b.cls_w1_hi, &b.cls_w1_lo}, {"cls_b1", "cls_branches.%d.3.bias", E, &b.cls_b1, nullptr, nullptr},
        {"cls_g1", "cls_branches.%d.4.weight", E, &b.cls_g1, nullptr, nullptr},             {"cls_be1", "cls_branches.%d.4.bias", E, &b.cls_be1, nullptr, nullptr},
        {"cls_w2", "cls_branches.%d.6.weight", 10 * E, &b.cls_w2, nullptr, nullptr},        {"cls_b2", "cls_branches.%d.6.bias", 10, &b.cls_b2, nullptr, nullptr},
        {"reg_w0", "reg_branches.%d.0.weight", E * E, &b.reg_w0, &b.reg_w0_hi, &b.reg_w0_lo}, {"reg_b0", "reg_branches.%d.0.bias", E, &b.reg_b0, nullptr, nullptr},
        {"reg_w1", "reg_branches.%d.2.weight", E * E, &b.reg_w1, &b.reg_w1_hi, &b.reg_w1_lo}, {"reg_b1", "reg_branches.%d.2.bias", E, &b.reg_b1, nullptr, nullptr},
        {"reg_w2", "reg_branches.%d.4.weight", 10 * E, &b.reg_w2, nullptr, nullptr},        {"reg_b2", "reg_branches.%d.4.bias", 10, &b.reg_b2, nullptr, nullptr}};
    for (const Br& r : brs) {
        float* d = P.put(std::string("br.") + r.field, (int64_t)L * r.numel, r.slot);
        if (d)
            for (int l = 0; l < L; ++l) {
                char key[96];
                snprintf(key, sizeof(key), r.key, l);
                const Mv2dNamedTensor* t = P.find(bh + key, r.numel);
                if (t) memcpy(d + (size_t)l * r.numel, t->data, (size_t)r.numel * 4);
            }
        if (r.hi) P.put_split(std::string("br.") + r.field + "_hi", std::string("br.") + r.field + "_lo", d, (int64_t)L * r.numel, r.hi, r.lo);
    }
    P.put_key("post_g", bh + "transformer.decoder.post_norm.weight", E, &b.post_g);
    P.put_key("post_b", bh + "transformer.decoder.post_norm.bias", E, &b.post_b);
    if (br) *br = b;
    // frequency table of the sine embeddings (pe.py:24-25, positional_encoding.py:78-80): 10000^(2*(i//2)/128) in fp32
    {
        // a caller may pass its own "dim_t" (the Python binding passes the table torch's fp32 pow produces, so the sine
        // embeddings see the reference's exact frequencies); otherwise correctly-rounded pow
        const Mv2dNamedTensor* given = nullptr;
        for (int i = 0; i < P.n_sd && P.arena; ++i)
            if (P.sd[i].name && !strcmp(P.sd[i].name, "dim_t") && P.sd[i].numel == 128 && P.sd[i].data) given = &P.sd[i];
        float* d = P.put("dim_t", 128, &dp);
        if (d)
            for (int i = 0; i < 128; ++i) d[i] = given ? given->data[i] : (float)std::pow(10000.0, (double)(float)((float)(2 * (i / 2)) / 128.f));
    }
    return 0;
}

}  // namespace

extern "C" {

MV2D_API int64_t mv2d_pack_weights_bytes(int num_layers, int fold_first_self_attn, int* n_entries) {
    if (num_layers < 1 || num_layers > MV2D_MAX_LAYERS) return -1;
    Packer P{nullptr, 0, nullptr, 0, 0, nullptr, nullptr, 0};
    pack_all(P, num_layers, fold_first_self_attn, nullptr, nullptr);
    if (n_entries) *n_entries = P.n_dir;
    return P.used;
}

MV2D_API int mv2d_pack_weights(const Mv2dNamedTensor* state_dict, int n_tensors, int num_layers, int fold_first_self_attn,
                               void* host_arena, int64_t arena_bytes, const void* device_base, Mv2dPackedEntry* dir, int dir_cap,
                               int* n_dir, Mv2dLayerWeights* layers, Mv2dBranchWeights* branches) {
    MV2D_CHECK_ARG(state_dict && n_tensors > 0 && host_arena && layers && branches, "pack_weights: null argument");
    MV2D_CHECK_ARG(num_layers >= 1 && num_layers <= MV2D_MAX_LAYERS, "pack_weights: bad num_layers");
    MV2D_CHECK_ARG(((uintptr_t)host_arena & 15) == 0 && ((uintptr_t)device_base & 255) == 0, "pack_weights: arena alignment (host 16, device 256)");
    Packer P{state_dict, n_tensors, (float*)host_arena, arena_bytes, 0, (const char*)device_base, dir, dir ? dir_cap : 0};
    pack_all(P, num_layers, fold_first_self_attn, layers, branches);
    return finish(P, n_dir, nullptr);
}

MV2D_API int64_t mv2d_pack_neck_bytes(int* n_entries) {
    if (n_entries) *n_entries = 6;
    const int64_t a = ((int64_t)E * E * 4 + 255) / 256 * 256, b = ((int64_t)E * 9 * E * 4 + 255) / 256 * 256, c = 256 * 4;
    return 2 * a + 2 * b + 2 * c;
}

MV2D_API int mv2d_pack_neck(const Mv2dNamedTensor* state_dict, int n_tensors, void* host_arena, int64_t arena_bytes,
                            const void* device_base, Mv2dPackedEntry* dir, int dir_cap, int* n_dir) {
    MV2D_CHECK_ARG(state_dict && n_tensors > 0 && host_arena, "pack_neck: null argument");
    MV2D_CHECK_ARG(((uintptr_t)host_arena & 15) == 0 && ((uintptr_t)device_base & 255) == 0, "pack_neck: arena alignment (host 16, device 256)");
    std::vector<Mv2dNamedTensor> sd(state_dict, state_dict + n_tensors);
    for (auto& t : sd)
        if (t.name && !strncmp(t.name, "neck.", 5)) t.name += 5;
    Packer P{sd.data(), n_tensors, (float*)host_arena, arena_bytes, 0, (const char*)device_base, dir, dir ? dir_cap : 0};
    const float* dp;
    const Mv2dNamedTensor* lat = P.find("lateral_convs.0.conv.weight", E * E);
    P.put_split("lat_w", "lat_w_lo", lat ? lat->data : nullptr, E * E, &dp, &dp);
    P.put_key("lat_b", "lateral_convs.0.conv.bias", E, &dp);
    std::vector<float> km;
    const Mv2dNamedTensor* fpn = P.find("fpn_convs.0.conv.weight", (int64_t)E * E * 9);
    if (fpn) conv3x3_kmajor(fpn->data, E, E, km);
    P.put_split("fpn_w", "fpn_w_lo", fpn ? km.data() : nullptr, (int64_t)E * 9 * E, &dp, &dp);
    P.put_key("fpn_b", "fpn_convs.0.conv.bias", E, &dp);
    return finish(P, n_dir, nullptr);
}

}  // extern "C"
