// kokoro.cu -- Kokoro-82M on B200: weight repacking into HBM and the batched two-pass forward.
//
// Replaces the reference's per-utterance GGML graph build + CPU compute
//   kokoro_model::assign_weight / post_load_assign   src/models/kokoro/model.cpp:310-427
//   build_kokoro_duration_graph                      src/models/kokoro/model.cpp:938-1047
//   build_kokoro_graph + set_inputs + run            src/models/kokoro/model.cpp:1141-1325
// with a fixed sequence of sm_100a kernels over a ragged batch of independent utterances:
// activations are channels-last [utterance][time][channel] padded to the longest utterance; the one-hot duration-mask
// matmuls become gathers; every F16-weight contraction runs in conv_gemm (tensor cores, fp32 accumulate); norms /
// AdaIN / snake / leaky-relu are fused single passes; the six bi-LSTMs are persistent cluster kernels.
#include "kokoro.h"
#include <algorithm>
#include <cmath>
#include <cstring>
#include <cstdlib>

namespace b2 {

// ------------------------------------------------------------------------------------------ arena
int Arena::reserve(size_t bytes) {
    if (bytes <= cap) { off = 0; return 0; }
    if (base) cudaFree(base);
    base = nullptr; cap = 0; off = 0;
    if (cudaMalloc(&base, bytes) != cudaSuccess) { cudaGetLastError(); set_error("workspace of %.1f GB does not fit in HBM", bytes / 1e9); return 1; }
    cap = bytes;
    return 0;
}
void * Arena::alloc(size_t bytes) {
    const size_t a = (off + 255) & ~(size_t) 255;
    if (a + bytes > cap) { set_error("workspace arena exhausted (%zu + %zu > %zu)", a, bytes, cap); return nullptr; }
    off = a + bytes;
    return base + a;
}
void Arena::release() { if (base) cudaFree(base); base = nullptr; cap = off = 0; }

// ------------------------------------------------------------------------------------------ weight hand-off
static inline float h2f(uint16_t h) { __half_raw r; r.x = h; return __half2float(__half(r)); }

int Kokoro::assign(const char * name, int type, int n_dims, const int64_t * ne, const void * data, size_t nbytes) {
    if (prepared) { set_error("assign_weight after prepare"); return 1; }
    std::string nm(name);
    if (nm.rfind("kokoro.", 0) == 0) nm = nm.substr(7);
    HostTensor t;
    int64_t n = 1;
    for (int i = n_dims - 1; i >= 0; i--) { t.shape.push_back(ne[i]); n *= ne[i]; }
    t.v.resize((size_t) n);
    if (type == 0) {
        if (nbytes < (size_t) n * 4) { set_error("tensor %s: short data", name); return 1; }
        memcpy(t.v.data(), data, (size_t) n * 4);
    } else if (type == 1) {
        if (nbytes < (size_t) n * 2) { set_error("tensor %s: short data", name); return 1; }
        const uint16_t * s = (const uint16_t *) data;
        for (int64_t i = 0; i < n; i++) t.v[(size_t) i] = h2f(s[i]);
        t.f16 = true;
    } else {
        set_error("tensor %s: ggml type %d not supported (F32/F16 only)", name, type);
        return 1;
    }
    host[nm] = std::move(t);
    return 0;
}

namespace {

struct Prep {
    Kokoro * m;
    bool ok = true;
    const HostTensor * get(const std::string & n) {
        auto it = m->host.find(n);
        if (it == m->host.end()) { set_error("missing tensor kokoro.%s", n.c_str()); ok = false; return nullptr; }
        return &it->second;
    }
    void * dev(const void * src, size_t bytes) {
        void * d = nullptr;
        if (cudaMalloc(&d, bytes) != cudaSuccess) { cudaGetLastError(); set_error("cudaMalloc(%zu) failed for weights", bytes); ok = false; return nullptr; }
        cudaMemcpy(d, src, bytes, cudaMemcpyHostToDevice);
        m->dev_allocs.push_back(d);
        m->weight_bytes += bytes;
        return d;
    }
    float * f32(const std::string & n) { auto t = get(n); return t ? (float *) dev(t->v.data(), t->v.size() * 4) : nullptr; }
    float * f32v(const std::vector<float> & v) { return (float *) dev(v.data(), v.size() * 4); }
    // [N][Cin][K] (numpy order) -> fp16 [Npad][K][CinPad]
    W16 w16_from(const std::vector<float> & src, int N, int Cin, int K) {
        W16 w; w.N = N; w.Cin = Cin; w.KW = K; w.CinPad = round_up(Cin, 64); w.Npad = N > 64 ? round_up(N, 128) : 64;
        std::vector<__half> h((size_t) w.Npad * K * w.CinPad, __float2half(0.f));
        for (int n = 0; n < N; n++)
            for (int ci = 0; ci < Cin; ci++)
                for (int k = 0; k < K; k++) h[((size_t) n * K + k) * w.CinPad + ci] = __float2half(src[((size_t) n * Cin + ci) * K + k]);
        w.w = (__half *) dev(h.data(), h.size() * 2);
        return w;
    }
    // rank: the tensor's rank in the GGUF file when it may arrive with fewer dimensions -- a caller that hands over ggml tensors reports ggml_n_dims(), which drops
    // outermost dimensions of size 1 (a [1][256][1] conv kernel arrives as [256][1]); missing leading dimensions are 1
    W16 w16(const std::string & n, int rank = 0) {
        auto t = get(n);
        if (!t) return W16();
        std::vector<int64_t> sh = t->shape;
        while ((int) sh.size() < rank) sh.insert(sh.begin(), 1);
        const int N = (int) sh[0], Cin = (int) sh[1], K = sh.size() > 2 ? (int) sh[2] : 1;
        return w16_from(t->v, N, Cin, K);
    }
    Lstm lstm(const std::string & base) {
        Lstm L;
        const char * wp[2] = {"weights", "reverse_weights"};
        const char * bp[2] = {"biases", "reverse_biases"};
        auto t0 = get(base + ".0.weights.0");
        if (!t0) return L;
        const int H = (int) t0->shape[0], In = (int) t0->shape[1];
        if (H != 256) { set_error("LSTM %s: hidden size %d != 256", base.c_str(), H); ok = false; return L; }
        std::vector<float> wih((size_t) 2 * 4 * H * In), bih((size_t) 2 * 4 * H), bhh((size_t) 2 * 4 * H);
        std::vector<__half> whh((size_t) 2 * 4 * H * H);
        for (int d = 0; d < 2; d++)
            for (int g = 0; g < 4; g++) {
                auto wi = get(base + ".0." + wp[d] + "." + std::to_string(2 * g));
                auto wh = get(base + ".0." + wp[d] + "." + std::to_string(2 * g + 1));
                auto bi = get(base + ".0." + bp[d] + "." + std::to_string(2 * g));
                auto bh = get(base + ".0." + bp[d] + "." + std::to_string(2 * g + 1));
                if (!wi || !wh || !bi || !bh) return L;
                for (int u = 0; u < H; u++) {
                    const size_t row = (size_t) d * 4 * H + (size_t) u * 4 + g;          // (dir, unit, gate) order for the projection GEMM
                    memcpy(&wih[row * In], &wi->v[(size_t) u * In], (size_t) In * 4);
                    bih[row] = bi->v[u];
                    bhh[(size_t) d * 4 * H + g * H + u] = bh->v[u];
                    for (int k = 0; k < H; k++) whh[((size_t) d * 4 * H + g * H + u) * H + k] = __float2half(wh->v[(size_t) u * H + k]);
                }
            }
        L.wih = w16_from(wih, 2 * 4 * H, In, 1);
        L.bih = f32v(bih);
        L.bhh = f32v(bhh);
        L.whh = (__half *) dev(whh.data(), whh.size() * 2);
        return L;
    }
    StyleSlot style(int kind, const std::string & gw, const std::string & gb, const std::string & bw, const std::string & bb) {
        StyleSlot s;
        auto a = get(gw), b = get(gb), c = get(bw), d = get(bb);
        if (!a || !b || !c || !d) return s;
        s.C = (int) a->shape[0];
        auto & W = m->sty_w_host[kind]; auto & Bv = m->sty_b_host[kind];
        s.goff = (int) Bv.size();
        W.insert(W.end(), a->v.begin(), a->v.end()); Bv.insert(Bv.end(), b->v.begin(), b->v.end());
        s.boff = (int) Bv.size();
        W.insert(W.end(), c->v.begin(), c->v.end()); Bv.insert(Bv.end(), d->v.begin(), d->v.end());
        return s;
    }
    AdaBlock ada(const std::string & base, int kind) {
        AdaBlock k;
        auto c1 = get(base + ".conv1_weight");
        if (!c1) return k;
        k.cout = (int) c1->shape[0]; k.cin = (int) c1->shape[1];
        k.conv1 = w16(base + ".conv1_weight"); k.b1 = f32(base + ".conv1_bias");
        k.conv2 = w16(base + ".conv2_weight"); k.b2 = f32(base + ".conv2_bias");
        k.n1 = style(kind, base + ".norm1_gamma_weight", base + ".norm1_gamma_bias", base + ".norm1_beta_weight", base + ".norm1_beta_bias");
        k.n2 = style(kind, base + ".norm2_gamma_weight", base + ".norm2_gamma_bias", base + ".norm2_beta_weight", base + ".norm2_beta_bias");
        if (m->host.count(base + ".pool_weight")) { k.pool = true; k.poolw = f32(base + ".pool_weight"); k.poolb = f32(base + ".pool_bias"); }
        if (m->host.count(base + ".conv1x1_weight")) { k.has1x1 = true; k.conv1x1 = w16(base + ".conv1x1_weight"); }
        if (k.pool && !k.has1x1) { set_error("%s: pool without conv1x1 is not supported", base.c_str()); ok = false; }
        return k;
    }
    GenResBlock gres(const std::string & base, const std::string & kvbase) {
        GenResBlock r;
        for (int i = 0; i < 3; i++) {
            const std::string p = base + "." + std::to_string(i) + ".";
            auto c = get(p + "convs1_weight");
            if (!c) return r;
            r.C = (int) c->shape[0]; r.K = (int) c->shape[2];
            r.c1[i] = w16(p + "convs1_weight"); r.b1[i] = f32(p + "convs1_bias");
            r.c2[i] = w16(p + "convs2_weight"); r.b2[i] = f32(p + "convs2_bias");
            r.a1[i] = f32(p + "alpha1"); r.a2[i] = f32(p + "alpha2");
            r.s1[i] = style(1, p + "gamma1_weight", p + "gamma1_bias", p + "beta1_weight", p + "beta1_bias");
            r.s2[i] = style(1, p + "gamma2_weight", p + "gamma2_bias", p + "beta2_weight", p + "beta2_bias");
            auto pk = m->kv.find(kvbase + "." + std::to_string(i) + ".padding");
            auto dk = m->kv.find(kvbase + "." + std::to_string(i) + ".dilation");
            if (pk == m->kv.end() || dk == m->kv.end()) {   // model.cpp:264-268 aborts the same way
                set_error("Could not find dilation and padding for generator residual block at key, '%s.%d'.", kvbase.c_str(), i); ok = false; return r;
            }
            r.pad[i] = (int) pk->second; r.dil[i] = (int) dk->second;
        }
        return r;
    }
};

}  // namespace

int Kokoro::prepare() {
    if (prepared) return 0;
    Prep P{this};
    auto kvget = [&](const std::string & k, uint32_t def) { auto it = kv.find(k); return it == kv.end() ? def : it->second; };
    recurrence = (int) kvget("kokoro.duration_predictor.albert.recurrence", 12);
    heads = (int) kvget("kokoro.duration_predictor.albert.attn_heads", 12);
    // ALBERT
    tok_embd = P.f32("albert.token_embd"); pos_embd = P.f32("albert.position_embd"); type_embd = P.f32("albert.token_type_embd");
    if (auto t = P.get("albert.token_embd")) n_vocab = (int) (t->v.size() / 128);        // rows of the two token tables: run_batch rejects ids beyond either
    if (auto t = P.get("albert.position_embd")) n_positions = (int) (t->v.size() / 128);
    in_nw = P.f32("albert.norm"); in_nb = P.f32("albert.norm_bias");
    embd_w = P.f32("albert.embd"); embd_b = P.f32("albert.embd_bias");
    {
        auto q = P.get("albert.layer.0.q"), k = P.get("albert.layer.0.k"), v = P.get("albert.layer.0.v");
        auto qb = P.get("albert.layer.0.q_bias"), kb = P.get("albert.layer.0.k_bias"), vb = P.get("albert.layer.0.v_bias");
        if (!q || !k || !v || !qb || !kb || !vb) return 1;
        std::vector<float> w(q->v); w.insert(w.end(), k->v.begin(), k->v.end()); w.insert(w.end(), v->v.begin(), v->v.end());
        std::vector<float> b(qb->v); b.insert(b.end(), kb->v.begin(), kb->v.end()); b.insert(b.end(), vb->v.begin(), vb->v.end());
        qkv = P.w16_from(w, 3 * (int) q->shape[0], (int) q->shape[1], 1);
        qkv_b = P.f32v(b);
    }
    o = P.w16("albert.layer.0.o"); o_b = P.f32("albert.layer.0.o_bias");
    ffn = P.w16("albert.layer.0.ffn"); ffn_b = P.f32("albert.layer.0.ffn_bias");
    ffn_out = P.w16("albert.layer.0.ffn_out"); ffn_out_b = P.f32("albert.layer.0.ffn_out_bias");
    attn_norm_w = P.f32("albert.layer.0.attn_norm"); attn_norm_b = P.f32("albert.layer.0.attn_norm_bias");
    ffn_norm_w = P.f32("albert.layer.0.ffn_norm"); ffn_norm_b = P.f32("albert.layer.0.ffn_norm_bias");
    // prosody predictor
    encode = P.w16("duration_predictor.encode"); encode_b = P.f32("duration_predictor.encode_bias");
    for (int i = 0; i < 3; i++) {
        dp_lstm[i] = P.lstm("duration_predictor.layers." + std::to_string(2 * i) + ".lstm");
        const std::string p = "duration_predictor.layers." + std::to_string(2 * i + 1) + ".";
        dp_ada[i] = P.style(0, p + "gamma_weight", p + "gamma_bias", p + "beta_weight", p + "beta_bias");
    }
    dur_lstm = P.lstm("duration_predictor.duration_lstm");
    shared_lstm = P.lstm("duration_predictor.shared_lstm");
    dur_proj = P.w16("duration_predictor.duration_proj"); dur_proj_b = P.f32("duration_predictor.duration_proj_bias");
    for (int i = 0; i < 3; i++) {
        f0_blocks[i] = P.ada("duration_predictor.f0_blocks." + std::to_string(i), 0);
        n_blocks[i] = P.ada("duration_predictor.n_blocks." + std::to_string(i), 0);
    }
    f0_proj = P.w16("duration_predictor.f0_proj_kernel", 3); f0_proj_b = P.f32("duration_predictor.f0_proj_bias");
    n_proj = P.w16("duration_predictor.n_proj_kernel", 3); n_proj_b = P.f32("duration_predictor.n_proj_bias");
    // text encoder
    if (auto t = P.get("text_encoder.embedding_weight")) {
        n_vocab = std::min(n_vocab, (int) (t->v.size() / 512));
        std::vector<__half> h(t->v.size());
        for (size_t i = 0; i < h.size(); i++) h[i] = __float2half(t->v[i]);
        text_embd = (__half *) P.dev(h.data(), h.size() * 2);
    }
    for (int i = 0; i < 3; i++) {
        const std::string p = "text_encoder.layers." + std::to_string(i) + ".";
        te_conv[i] = P.w16(p + "weight"); te_b[i] = P.f32(p + "bias"); te_gamma[i] = P.f32(p + "gamma"); te_beta[i] = P.f32(p + "beta");
    }
    text_lstm = P.lstm("text_encoder.lstm");
    // decoder
    if (auto t = P.get("decoder.f0_conv_weight")) memcpy(f0_conv_w, t->v.data(), 12);
    if (auto t = P.get("decoder.n_conv_weight")) memcpy(n_conv_w, t->v.data(), 12);
    if (auto t = P.get("decoder.f0_conv_bias")) f0_conv_b[0] = t->v[0];
    if (auto t = P.get("decoder.n_conv_bias")) n_conv_b[0] = t->v[0];
    asr_conv = P.w16("decoder.asr_conv_weight"); asr_conv_b = P.f32("decoder.asr_conv_bias");
    enc_block = P.ada("decoder.encoder_block", 1);
    for (int i = 0; i < 4; i++) dec_blocks[i] = P.ada("decoder.decoder_blocks." + std::to_string(i), 1);
    // generator
    const std::string g = "decoder.generator.", G = "kokoro.decoder.generator.";
    if (auto t = P.get(g + "m_source_weight")) { if (t->v.size() != 9) { set_error("m_source_weight must have 9 inputs"); return 1; } memcpy(m_src_w, t->v.data(), 36); }
    if (auto t = P.get(g + "m_source_bias")) m_src_b = t->v[0];
    for (int i = 0; i < 2; i++) {
        auto t = P.get(g + "ups." + std::to_string(i) + ".weight");
        if (!t) return 1;
        Up & u = ups[i];
        u.Cin = (int) t->shape[0]; u.Cout = (int) t->shape[1]; u.K = (int) t->shape[2];
        std::vector<float> r((size_t) u.K * u.Cin * u.Cout);
        for (int ci = 0; ci < u.Cin; ci++)
            for (int co = 0; co < u.Cout; co++)
                for (int k = 0; k < u.K; k++) r[((size_t) k * u.Cin + ci) * u.Cout + co] = t->v[((size_t) ci * u.Cout + co) * u.K + k];
        u.w = P.f32v(r); u.b = P.f32(g + "ups." + std::to_string(i) + ".bias");
        auto sk = kv.find(G + "up_convs." + std::to_string(i) + ".stride"), pk = kv.find(G + "up_convs." + std::to_string(i) + ".padding");
        if (sk == kv.end() || pk == kv.end()) { set_error("both padding and stride keys must be assigned in order to initialize a kokoro upsample block."); return 1; }
        u.stride = (int) sk->second; u.pad = (int) pk->second;
        if (u.K == 2 * u.stride && u.pad < u.stride && (3 * u.Cin) % 64 == 0 && getenv("B2TTS_NO_POLY_CONVT") == nullptr) {
            // ConvTranspose1d(K = 2s) == s interleaved 2-tap convolutions (one per output phase r = (o + p) mod s):
            //   out[q*s + r - p] = x[q] . W[:, :, r] + x[q-1] . W[:, :, r+s]
            // run as ONE tensor-core GEMM with N = s*Cout.  The F32 kernel semantics are kept by splitting both operands
            // into fp16 hi + lo parts (x = hi + lo, W = Whi + Wlo; hi*Whi + lo*Whi + hi*Wlo, error ~2^-22): K = 2 taps x 3*Cin.
            const int s = u.stride, C3 = 3 * u.Cin, N = s * u.Cout;
            std::vector<float> src((size_t) N * C3 * 2), brep((size_t) N);
            auto bt = P.get(g + "ups." + std::to_string(i) + ".bias");
            if (!bt) return 1;
            for (int r = 0; r < s; r++)
                for (int co = 0; co < u.Cout; co++) {
                    const size_t n = (size_t) r * u.Cout + co;
                    brep[n] = bt->v[co];
                    for (int ci = 0; ci < u.Cin; ci++)
                        for (int k = 0; k < 2; k++) {
                            const float wv = t->v[((size_t) ci * u.Cout + co) * u.K + (k == 0 ? r + s : r)];   // tap 0 reads x[q-1]
                            const float whi = __half2float(__float2half(wv)), wlo = wv - whi;
                            src[(n * C3 + ci) * 2 + k] = whi;                 // x hi * W hi
                            src[(n * C3 + u.Cin + ci) * 2 + k] = whi;         // x lo * W hi
                            src[(n * C3 + 2 * u.Cin + ci) * 2 + k] = wlo;     // x hi * W lo
                        }
                }
            u.w3 = P.w16_from(src, N, C3, 2);
            u.b_rep = P.f32v(brep);
            u.poly = true;
        }
        nconv[i].w = P.w16(g + "noise_blocks." + std::to_string(i) + ".conv_weight");
        nconv[i].b = P.f32(g + "noise_blocks." + std::to_string(i) + ".conv_bias");
        auto ns = kv.find(G + "noise_blocks." + std::to_string(i) + ".stride"), np = kv.find(G + "noise_blocks." + std::to_string(i) + ".padding");
        if (ns == kv.end() || np == kv.end()) { set_error("both padding and stride keys must be assigned in order to initialize a kokoro noise block."); return 1; }
        nconv[i].stride = (int) ns->second; nconv[i].pad = (int) np->second;
        if (nconv[i].stride > 1) {
            // out[t] = sum_k x[t*s + k - pad] w[k].  View the (channels-last, 64-half rows) input as rows of s frames, X'[r][j*64 + c] =
            // x[r*s + j][c] -- a pure reinterpretation when the per-utterance pitch is a multiple of s -- and shift the taps by
            // sh = (s - pad % s) % s:  out[t] = sum_m sum_{j,c} X'[t - pad' + m][j*64 + c] w'[m][j*64 + c],  w'[m][j] = w[m*s + j - sh]
            // (zero outside [0, K)), pad' = pad / s + (sh != 0).  A stride-1 conv with K' = ceil((K + sh) / s) taps: tcgen05-eligible.
            auto t = P.get(g + "noise_blocks." + std::to_string(i) + ".conv_weight");
            const int s = nconv[i].stride, N = nconv[i].w.N, Cin = nconv[i].w.Cin, K = nconv[i].w.KW, CX = nconv[i].w.CinPad;
            const int sh = (s - nconv[i].pad % s) % s, Kp = (K + sh + s - 1) / s, Cp = s * CX;
            std::vector<float> src((size_t) N * Cp * Kp, 0.f);
            for (int n = 0; n < N; n++)
                for (int m = 0; m < Kp; m++)
                    for (int j = 0; j < s; j++) {
                        const int k = m * s + j - sh;
                        if (k < 0 || k >= K) continue;
                        for (int c = 0; c < Cin; c++) src[((size_t) n * Cp + j * CX + c) * Kp + m] = t->v[((size_t) n * Cin + c) * K + k];
                    }
            nconv[i].wp = P.w16_from(src, N, Cp, Kp);
            nconv[i].wp.Cin = K * Cin / Kp;   // roofline accounting counts the conv's own K*Cin products, not the zero-padded polyphase ones
            nconv[i].padp = nconv[i].pad / s + (sh ? 1 : 0);
            nconv[i].poly = true;
        }
        nres[i] = P.gres(g + "noise_blocks." + std::to_string(i) + ".resblock", G + "noise_blocks." + std::to_string(i) + ".res_block");
    }
    for (int i = 0; i < 6; i++) res[i] = P.gres(g + "resblocks." + std::to_string(i), G + "res_blocks." + std::to_string(i));
    conv_post = P.w16(g + "conv_post_weight"); conv_post_b = P.f32(g + "conv_post_bias");
    post_pad = (int) kvget(G + "padding", 3);
    if (!P.ok) return 1;
    for (int k = 0; k < 2; k++) {
        sty_n[k] = (int) sty_b_host[k].size();
        sty_w[k] = P.f32v(sty_w_host[k]); sty_b[k] = P.f32v(sty_b_host[k]);
        sty_w_host[k].clear(); sty_w_host[k].shrink_to_fit();
    }
    for (auto & kvp : host)
        if (kvp.first.rfind("voice_tensors.", 0) == 0) { voice_names.push_back(kvp.first.substr(14)); voices_host[kvp.first.substr(14)] = kvp.second.v; }
    if (voice_names.empty()) { set_error("model has no voice tensors"); return 1; }
    if (!P.ok) return 1;
    host.clear();
    for (int i = 0; i < 4; i++) B2_CUDA(cudaEventCreate(&ev[i]));
    cudaDeviceSynchronize();               // legacy-stream uploads above vs kernels on the non-blocking ctx->stream
    prepared = true;
    return 0;
}

void Kokoro::free_all() {
    for (void * p : dev_allocs) cudaFree(p);
    dev_allocs.clear();
    a1.release(); a2.release();
    if (pcm_pinned) cudaFreeHost(pcm_pinned);
    if (lens_pinned) cudaFreeHost(lens_pinned);
    for (int i = 0; i < 4; i++) if (ev[i]) cudaEventDestroy(ev[i]);
}

// ------------------------------------------------------------------------------------------ forward
namespace {

struct Fwd {
    Kokoro * m; Ctx * ctx; Arena * ar; int B;
    bool fail = false;
    template <class T> T * al(size_t n) { T * p = (T *) ar->alloc(n * sizeof(T)); if (!p) fail = true; return p; }

    int tap(const char * name, const void * ptr, int64_t rows, int64_t cols, int64_t ld, int64_t padded) {
        auto ov = m->overrides.find(name);
        if (ov != m->overrides.end()) {
            if ((int64_t) ov->second.size() != rows * cols) { set_error("override '%s': have %zu floats, buffer is %lld x %lld", name, ov->second.size(), (long long) rows, (long long) cols); return 1; }
            if (m->taps_on) {   // keep what the kernels computed (the tap), then teacher-force the consumer's input
                float * keep = al<float>((size_t) rows * cols);
                if (fail) return 1;
                B2_CUDA(cudaMemcpy2DAsync(keep, cols * 4, ptr, ld * 4, cols * 4, rows, cudaMemcpyDeviceToDevice, ctx->stream));
                m->taps[name] = Tap{keep, rows, cols, cols, padded};
            }
            B2_CUDA(cudaMemcpy2DAsync((void *) ptr, ld * 4, ov->second.data(), cols * 4, cols * 4, rows, cudaMemcpyHostToDevice, ctx->stream));
            return 0;
        }
        if (m->taps_on) m->taps[name] = Tap{ptr, rows, cols, ld, padded};
        return 0;
    }

    int gemm(const __half * A, int lda, const W16 & W, const float * bias, int Lin, int Lout, const int * lenIn, const int * lenOut, int stride, int dil,
             int pad, float * outF, int ldo, int coff, __half * outH = nullptr, int ldoh = 0, int coffh = 0, const float * add1 = nullptr, int ldadd1 = 0,
             const float * add2 = nullptr, int ldadd2 = 0, float div = 0.f, int act = ACT_NONE, int Bn = -1, bool tailClean = false) {
        ConvGemmParams p;
        p.tailClean = tailClean;
        p.A = A; p.lda = lda; p.W = W.w; p.bias = bias; p.outF = outF; p.ldo = ldo; p.coff = coff; p.outH = outH; p.ldoh = ldoh; p.coffh = coffh;
        p.add1 = add1; p.ldadd1 = ldadd1; p.add2 = add2; p.ldadd2 = ldadd2; p.div = div; p.act = act;
        p.B = Bn < 0 ? B : Bn; p.LmaxIn = Lin; p.LmaxOut = Lout; p.lenIn = lenIn; p.lenOut = lenOut;
        p.N = W.N; p.Npad = W.Npad; p.KW = W.KW; p.CinPad = W.CinPad; p.CinTrue = W.Cin; p.stride = stride; p.dil = dil; p.pad = pad;
        return conv_gemm(ctx, p);
    }

    // conv_gemm whose fp32 result (ld == N) is InstanceNorm'ed next: the statistics come out of the tcgen05 epilogue
    // (per-tile partials + a tiny finalize); if the shape fell back to the mma.sync kernel, a separate pass computes them.
    int gemm_stats(const __half * A, int lda, const W16 & W, const float * bias, int L, const int * len, int dil, int pad, float * outF,
                   const float * add1, const float * add2, float div, float * part, double * sums_out, int Lin = 0, const int * lenInP = nullptr,
                   bool tailClean = true) {
        ConvGemmParams p;
        p.tailClean = tailClean;   // gen_resblock: the operand always comes straight from adain_apply
        p.A = A; p.lda = lda; p.W = W.w; p.bias = bias; p.outF = outF; p.ldo = W.N; p.add1 = add1; p.ldadd1 = W.N; p.add2 = add2; p.ldadd2 = W.N; p.div = div;
        p.B = B; p.LmaxIn = Lin ? Lin : L; p.LmaxOut = L; p.lenIn = lenInP ? lenInP : len; p.lenOut = len;
        p.N = W.N; p.Npad = W.Npad; p.KW = W.KW; p.CinPad = W.CinPad; p.CinTrue = W.Cin; p.stride = 1; p.dil = dil; p.pad = pad;
        p.statsPart = part;
        const uint64_t before = ctx->umma_launches;
        if (conv_gemm(ctx, p)) return 1;
        if (ctx->umma_launches != before) return stats_finalize(ctx, part, B, cdiv(L, conv_umma_tile_m(p)), W.N, sums_out);
        return inorm_stats(ctx, outF, W.N, W.N, B, L, len, sums_out);
    }

    // bi-LSTM over a16 [B][Lmax][CinPad] -> out fp32 (ldo/coff) (+ fp16 copy)
    int lstm(const Lstm & L, const __half * a16, int lda, int Lmax, const int * len, int maxLen, float * out, int ldo, int coff, __half * outH, int ldoh,
             int coffh) {
        float * xp = al<float>((size_t) B * Lmax * 2048);
        if (fail) return 1;
        if (gemm(a16, lda, L.wih, L.bih, Lmax, Lmax, len, len, 1, 1, 0, xp, 2048, 0)) return 1;
        LstmParams p;
        p.xp = xp; p.whh = L.whh; p.bhh = L.bhh; p.out = out; p.ldo = ldo; p.coff = coff; p.outH = outH; p.ldoh = ldoh; p.coffh = coffh;
        p.len = len; p.B = B; p.Lmax = Lmax; p.maxLen = maxLen;
        return bilstm(ctx, p);
    }

    // AdaIN residual block (model.cpp:88-134): x [B][Lin][cin] -> out [B][Lout][cout] at (ldo, coff)
    int ada_block(const AdaBlock & k, const float * gb, int ldgb, const float * x, int ldx, int Lin, const int * lenIn, int Lout, const int * lenOut,
                  float * out, int ldo, int coff) {
        const int cmax = std::max(k.cin, k.cout);
        double * sums = al<double>((size_t) B * cmax * 2);
        __half * a16 = al<__half>((size_t) B * Lout * k.conv1.CinPad);
        float * h = al<float>((size_t) B * Lout * k.cout);
        __half * h16 = al<__half>((size_t) B * Lout * k.conv2.CinPad);
        if (fail) return 1;
        if (inorm_stats(ctx, x, ldx, k.cin, B, Lin, lenIn, sums)) return 1;
        AdainParams ap;
        ap.x = x; ap.ldx = ldx; ap.C = k.cin; ap.B = B; ap.Lmax = Lin; ap.len = lenIn; ap.sums = sums; ap.gb = gb; ap.ldgb = ldgb;
        ap.goff = k.n1.goff; ap.boff = k.n1.boff; ap.act = NACT_LRELU02;
        if (k.pool) {
            if (Lout != 2 * Lin) { set_error("ada_block: pool expects Lout == 2*Lin"); return 1; }
            const int ldt = round_up(k.cin, 4);
            float * tmp = al<float>((size_t) B * Lin * ldt);
            if (fail) return 1;
            ap.outF = tmp; ap.ldof = ldt;
            if (adain_apply(ctx, ap)) return 1;
            if (pool_convt(ctx, tmp, ldt, k.cin, B, Lin, lenIn, k.poolw, k.poolb, a16, k.conv1.CinPad, k.conv1.CinPad)) return 1;
        } else {
            ap.outH = a16; ap.ldoh = k.conv1.CinPad; ap.Cpad = k.conv1.CinPad;
            if (adain_apply(ctx, ap)) return 1;
        }
        if (gemm(a16, k.conv1.CinPad, k.conv1, k.b1, Lout, Lout, lenOut, lenOut, 1, 1, 1, h, k.cout, 0, nullptr, 0, 0, nullptr, 0, nullptr, 0, 0.f, ACT_NONE, -1, !k.pool)) return 1;
        if (inorm_stats(ctx, h, k.cout, k.cout, B, Lout, lenOut, sums)) return 1;
        AdainParams ap2;
        ap2.x = h; ap2.ldx = k.cout; ap2.C = k.cout; ap2.B = B; ap2.Lmax = Lout; ap2.len = lenOut; ap2.sums = sums; ap2.gb = gb; ap2.ldgb = ldgb;
        ap2.goff = k.n2.goff; ap2.boff = k.n2.boff; ap2.act = NACT_LRELU02; ap2.outH = h16; ap2.ldoh = k.conv2.CinPad; ap2.Cpad = k.conv2.CinPad;
        if (adain_apply(ctx, ap2)) return 1;
        const float * sc = x; int ldsc = ldx;
        if (k.has1x1) {
            __half * sc16 = al<__half>((size_t) B * Lout * k.conv1x1.CinPad);
            float * scf = al<float>((size_t) B * Lout * k.cout);
            if (fail) return 1;
            if (cast_rows(ctx, x, ldx, k.cin, B, Lin, lenOut, Lout, k.pool ? 1 : 0, 1.0f, sc16, k.conv1x1.CinPad, k.conv1x1.CinPad)) return 1;
            if (gemm(sc16, k.conv1x1.CinPad, k.conv1x1, nullptr, Lout, Lout, lenOut, lenOut, 1, 1, 0, scf, k.cout, 0)) return 1;   // bias never applied: model.cpp:129
            sc = scf; ldsc = k.cout;
        }
        return gemm(h16, k.conv2.CinPad, k.conv2, k.b2, Lout, Lout, lenOut, lenOut, 1, 1, 1, out, ldo, coff, nullptr, 0, 0, sc, ldsc, nullptr, 0, sqrtf(2.0f), ACT_NONE, -1, true);
    }

    // generator residual block (model.cpp:136-165): x [B][L][C] -> out = (x_final [+ add2]) [/ div]
    // generator residual block (model.cpp:136-165): x [B][L][C] -> out = (x_final [+ add2]) [/ div].
    // sums_x: InstanceNorm statistics of x if the caller already has them (nullptr -> computed here);
    // sums_out: if non-null, receives the statistics of `out` (fused into the last conv's epilogue).
    int gen_resblock(const GenResBlock & r, const float * gb, int ldgb, const float * x, int L, const int * len, float * out, const float * add2, float div,
                     float * scratch[3], __half * a16, double * sums2[3], float * part, const double * sums_x, double * sums_out) {
        const float * inp = x;
        const int C = r.C, Cp = r.c1[0].CinPad;
        const double * s_in = sums_x;
        if (!s_in) { if (inorm_stats(ctx, inp, C, C, B, L, len, sums2[0])) return 1; s_in = sums2[0]; }
        for (int i = 0; i < 3; i++) {
            float * h = scratch[2];
            float * nxt = (i == 2) ? out : scratch[i & 1];
            AdainParams ap;
            ap.x = inp; ap.ldx = C; ap.C = C; ap.B = B; ap.Lmax = L; ap.len = len; ap.sums = s_in; ap.gb = gb; ap.ldgb = ldgb;
            ap.goff = r.s1[i].goff; ap.boff = r.s1[i].boff; ap.act = NACT_SNAKE; ap.alpha = r.a1[i]; ap.outH = a16; ap.ldoh = Cp; ap.Cpad = Cp;
            if (adain_apply(ctx, ap)) return 1;
            if (gemm_stats(a16, Cp, r.c1[i], r.b1[i], L, len, r.dil[i], r.pad[i], h, nullptr, nullptr, 0.f, part, sums2[1])) return 1;
            ap.x = h; ap.sums = sums2[1]; ap.goff = r.s2[i].goff; ap.boff = r.s2[i].boff; ap.alpha = r.a2[i];
            if (adain_apply(ctx, ap)) return 1;
            double * s_next = (i == 2) ? sums_out : sums2[2 - (i & 1) * 2];   // ping-pong between sums2[2] and sums2[0] (s_in may alias sums2[0] only at i == 0)
            if (s_next) {
                if (gemm_stats(a16, Cp, r.c2[i], r.b2[i], L, len, 1, r.pad[0], nxt, inp, (i == 2) ? add2 : nullptr, (i == 2) ? div : 0.f, part, s_next)) return 1;
            } else {
                if (gemm(a16, Cp, r.c2[i], r.b2[i], L, L, len, len, 1, 1, r.pad[0], nxt, C, 0, nullptr, 0, 0, inp, C, (i == 2) ? add2 : nullptr, C, (i == 2) ? div : 0.f, ACT_NONE, -1, true)) return 1;
            }
            inp = nxt; s_in = s_next;
        }
        return 0;
    }
};

}  // namespace

int Kokoro::run_batch(int B, const uint32_t * tokens, const int32_t * n_tokens, const char * voice, const uint64_t * noise_skip, const float ** pcm,
                      int64_t * n_samples, const float ** durations) {
    if (!prepared) { set_error("model not prepared"); return 1; }
    if (B <= 0) return 0;
    std::string vname = (voice && *voice) ? voice : "af_heart";
    auto vit = voices_host.find(vname);
    if (vit == voices_host.end()) { set_error("Failed to find Kokoro voice '%s' aborting.", vname.c_str()); return 1; }
    const std::vector<float> & vt = vit->second;
    const int n_vrows = (int) (vt.size() / 256);
    int Nmax = 0, ntot = 0;
    std::vector<int> ntok(B), tok_off(B);
    for (int b = 0; b < B; b++) {
        const int cap = std::min(std::min(512, n_positions), n_vrows + 2);
        if (n_tokens[b] < 3 || n_tokens[b] > cap) { set_error("utterance %d: n_tokens=%d out of range [3,%d]", b, n_tokens[b], cap); return 1; }
        for (int i = 0; i < n_tokens[b]; i++)                                      // the ids index albert.token_embd and text_encoder.embedding_weight on the device
            if (tokens[ntot + i] >= (uint32_t) n_vocab) { set_error("utterance %d: token %d is %u, the model's vocabulary has %d entries", b, i, tokens[ntot + i], n_vocab); return 1; }
        ntok[b] = n_tokens[b]; tok_off[b] = ntot; ntot += ntok[b]; Nmax = std::max(Nmax, ntok[b]);
    }
    taps.clear();
    cudaStream_t st = ctx->stream;
    B2_CUDA(cudaEventRecord(ev[0], st));

    // ================================================================ pass 1: durations (model.cpp:938-1047)
    const size_t rows1 = (size_t) B * Nmax;
    if (a1.reserve(rows1 * (2304 + 2048 * 2 + 768 * 6 + 640 * 4 + 2048 + 512 * 8 + 4096) * 4 + ((size_t) B * (sty_n[0] + sty_n[1]) + 65536) * 8 + (64 << 20))) return 1;
    Fwd F{this, ctx, &a1, B};
    int * d_tok = F.al<int>(ntot); int * d_tokoff = F.al<int>(B); int * d_ntok = F.al<int>(B);
    float * d_sty = F.al<float>((size_t) B * 256);
    unsigned long long * d_skip = F.al<unsigned long long>(B);
    if (F.fail) return 1;
    {
        std::vector<int> tk(ntot);
        for (int i = 0; i < ntot; i++) tk[i] = (int) tokens[i];
        std::vector<float> sty((size_t) B * 256);
        for (int b = 0; b < B; b++) {
            const float * row = &vt[(size_t) (ntok[b] - 3) * 256];                 // voice[n_tokens-3] (model.cpp:1013,1213)
            memcpy(&sty[(size_t) b * 128], row + 128, 512);                        // prosody style
            memcpy(&sty[(size_t) (B + b) * 128], row, 512);                        // decoder style
        }
        std::vector<unsigned long long> sk(B, 0ull);
        if (noise_skip) for (int b = 0; b < B; b++) sk[b] = noise_skip[b];
        B2_CUDA(cudaMemcpyAsync(d_tok, tk.data(), ntot * 4, cudaMemcpyHostToDevice, st));
        B2_CUDA(cudaMemcpyAsync(d_tokoff, tok_off.data(), B * 4, cudaMemcpyHostToDevice, st));
        B2_CUDA(cudaMemcpyAsync(d_ntok, ntok.data(), B * 4, cudaMemcpyHostToDevice, st));
        B2_CUDA(cudaMemcpyAsync(d_sty, sty.data(), sty.size() * 4, cudaMemcpyHostToDevice, st));
        B2_CUDA(cudaMemcpyAsync(d_skip, sk.data(), B * 8, cudaMemcpyHostToDevice, st));
        B2_CUDA(cudaStreamSynchronize(st));   // host staging vectors go out of scope
    }
    const float * styP = d_sty, * styD = d_sty + (size_t) B * 128;
    float * gbP = F.al<float>((size_t) B * sty_n[0]); float * gbD = F.al<float>((size_t) B * sty_n[1]);
    if (F.fail) return 1;
    if (linear_f32(ctx, styP, 128, sty_w[0], sty_b[0], B, 128, sty_n[0], gbP, sty_n[0])) return 1;
    if (linear_f32(ctx, styD, 128, sty_w[1], sty_b[1], B, 128, sty_n[1], gbD, sty_n[1])) return 1;

    float * e128 = F.al<float>(rows1 * 128);
    float * x = F.al<float>(rows1 * 768); float * y = F.al<float>(rows1 * 768); float * x2 = F.al<float>(rows1 * 768);
    __half * x16 = F.al<__half>(rows1 * 768); __half * x2_16 = F.al<__half>(rows1 * 768); __half * att16 = F.al<__half>(rows1 * 768);
    float * qkvb = F.al<float>(rows1 * 2304); __half * f16b = F.al<__half>(rows1 * 2048);
    if (F.fail) return 1;
    if (albert_embed(ctx, d_tok, d_tokoff, tok_embd, pos_embd, type_embd, in_nw, in_nb, B, Nmax, d_ntok, e128, 128)) return 1;
    if (linear_f32(ctx, e128, 128, embd_w, embd_b, (int) rows1, 128, 768, x, 768)) return 1;     // rows beyond len are garbage-in/garbage-out, never read
    if (F.tap("albert_embeddings", x, rows1, 768, 768, Nmax)) return 1;
    if (cast_rows(ctx, x, 768, 768, B, Nmax, d_ntok, Nmax, 0, 1.0f, x16, 768, 768)) return 1;
    const int hd = 768 / heads;
    for (int r = 0; r < recurrence; r++) {
        if (F.gemm(x16, 768, qkv, qkv_b, Nmax, Nmax, d_ntok, d_ntok, 1, 1, 0, qkvb, 2304, 0)) return 1;
        if (albert_attention(ctx, qkvb, B, Nmax, d_ntok, heads, hd, 0.125f, att16, 768)) return 1;
        if (F.gemm(att16, 768, o, o_b, Nmax, Nmax, d_ntok, d_ntok, 1, 1, 0, y, 768, 0, nullptr, 0, 0, x, 768)) return 1;
        RowNormParams rn;
        rn.x = y; rn.ldx = 768; rn.C = 768; rn.B = B; rn.Lmax = Nmax; rn.len = d_ntok; rn.eps = 1e-12f; rn.mode = LN_AFFINE;
        rn.w = ffn_norm_w; rn.bias = ffn_norm_b; rn.outF = x2; rn.ldof = 768; rn.outH = x2_16; rn.ldoh = 768;   // crossed names: model.cpp:765-770,994
        if (row_norm(ctx, rn)) return 1;
        if (F.gemm(x2_16, 768, ffn, ffn_b, Nmax, Nmax, d_ntok, d_ntok, 1, 1, 0, nullptr, 0, 0, f16b, 2048, 0, nullptr, 0, nullptr, 0, 0.f, ACT_GELU_F16LUT)) return 1;
        if (F.gemm(f16b, 2048, ffn_out, ffn_out_b, Nmax, Nmax, d_ntok, d_ntok, 1, 1, 0, y, 768, 0, nullptr, 0, 0, x2, 768)) return 1;
        rn.w = attn_norm_w; rn.bias = attn_norm_b; rn.outF = x; rn.outH = x16;
        if (row_norm(ctx, rn)) return 1;
    }
    if (F.tap("albert", x, rows1, 768, 768, Nmax)) return 1;

    float * cur = F.al<float>(rows1 * 640); __half * cur16 = F.al<__half>(rows1 * 640);
    float * lo = F.al<float>(rows1 * 512); __half * lo16 = F.al<__half>(rows1 * 512);
    float * logits = F.al<float>(rows1 * 64); float * d_lens = F.al<float>(rows1); int * d_T = F.al<int>(B);
    if (F.fail) return 1;
    // re-round the final ALBERT state for the encode matmul (x16 already holds it)
    if (F.gemm(x16, 768, encode, encode_b, Nmax, Nmax, d_ntok, d_ntok, 1, 1, 0, cur, 640, 0, cur16, 640, 0)) return 1;
    if (bcast_cols(ctx, styP, 128, 128, B, Nmax, d_ntok, cur, 640, 512, cur16, 640, 512)) return 1;
    for (int i = 0; i < 3; i++) {
        if (F.lstm(dp_lstm[i], cur16, 640, Nmax, d_ntok, Nmax, lo, 512, 0, nullptr, 0, 0)) return 1;
        RowNormParams rn;
        rn.x = lo; rn.ldx = 512; rn.C = 512; rn.B = B; rn.Lmax = Nmax; rn.len = d_ntok; rn.eps = 1e-5f; rn.mode = LN_ADA;
        rn.gb = gbP; rn.ldgb = sty_n[0]; rn.goff = dp_ada[i].goff; rn.boff = dp_ada[i].boff;
        rn.outF = cur; rn.ldof = 640; rn.outH = cur16; rn.ldoh = 640;
        if (row_norm(ctx, rn)) return 1;
    }
    if (F.tap("d", cur, rows1, 640, 640, Nmax)) return 1;
    if (overrides.count("d")) { if (cast_rows(ctx, cur, 640, 640, B, Nmax, d_ntok, Nmax, 0, 1.0f, cur16, 640, 640)) return 1; }
    if (F.lstm(dur_lstm, cur16, 640, Nmax, d_ntok, Nmax, lo, 512, 0, lo16, 512, 0)) return 1;
    if (F.gemm(lo16, 512, dur_proj, dur_proj_b, Nmax, Nmax, d_ntok, d_ntok, 1, 1, 0, logits, 64, 0)) return 1;
    if (F.tap("dur_logits", logits, rows1, dur_proj.N, 64, Nmax)) return 1;
    if (duration_tail(ctx, logits, 64, dur_proj.N, B, Nmax, d_ntok, d_lens)) return 1;
    if (F.tap("lens", d_lens, rows1, 1, 1, Nmax)) return 1;
    if (build_alignment(ctx, d_lens, B, Nmax, d_ntok, 0, nullptr, d_T)) return 1;
    B2_CUDA(cudaEventRecord(ev[1], st));

    // ---- host: total frames per utterance (model.cpp:1284-1287)
    if (lens_pinned_cap < 2 * (rows1 + B)) {
        if (lens_pinned) cudaFreeHost(lens_pinned);
        lens_pinned = nullptr; lens_pinned_cap = 0;
        B2_CUDA(cudaMallocHost(&lens_pinned, 2 * (rows1 + B) * 4));
        lens_pinned_cap = 2 * (rows1 + B);
    }
    std::vector<int> T(B);
    B2_CUDA(cudaMemcpyAsync(lens_pinned, d_lens, rows1 * 4, cudaMemcpyDeviceToHost, st));
    B2_CUDA(cudaMemcpyAsync(lens_pinned + rows1, d_T, B * 4, cudaMemcpyDeviceToHost, st));
    B2_CUDA(cudaStreamSynchronize(st));
    int Tmax = 0; size_t Ssum = 0;
    for (int b = 0; b < B; b++) { T[b] = ((int *) (lens_pinned + rows1))[b]; Tmax = std::max(Tmax, T[b]); Ssum += (size_t) T[b] * 600; }
    // compact per-utterance durations for the caller
    float * dur_out = lens_pinned + rows1 + B;
    for (int b = 0; b < B; b++) memcpy(dur_out + tok_off[b], lens_pinned + (size_t) b * Nmax, (size_t) ntok[b] * 4);
    if (durations) *durations = dur_out;
    if (chain_noise) {                                          // consecutive calls of one reference process: each utterance continues the uniform stream of the one before
        chain_skips.assign((size_t) B, 0ull);
        unsigned long long at = chain_noise_start;
        for (int b = 0; b < B; b++) { chain_skips[(size_t) b] = at; at += 9ull * 600ull * (unsigned long long) T[b]; }
        B2_CUDA(cudaMemcpyAsync(d_skip, chain_skips.data(), (size_t) B * 8, cudaMemcpyHostToDevice, st));      // (member vector: alive until the next synchronisation below)
    }

    // ================================================================ pass 2: generation (model.cpp:1141-1242)
    const int L1 = Tmax, L2 = 2 * Tmax, L4 = 120 * Tmax + 1, S = 600 * Tmax;
    // row pitches of the two generator stages.  The polyphase ConvTranspose GEMM writes [q][phase][Cout] rows, i.e. the stage
    // tensor shifted by `pad` rows inside a buffer of (Lin_pitch + 1) * stride rows per utterance; every buffer of a stage shares it.
    const int P0 = ups[0].poly ? (L2 + 1) * ups[0].stride : 20 * Tmax;
    const int P1 = ups[1].poly ? (P0 + 1) * ups[1].stride : L4;
    {
        const size_t big = (size_t) B * P1 * 128 * 4;     // one fp32 generator activation at full rate
        const size_t need = big * 14 + (size_t) B * L1 * (640 * 6 + 2048 * 4 + 1090 * 4 * 3 + 1024 * 8) + (size_t) B * S * 4 * 3 + (256 << 20);
        if (a2.reserve(need)) return 1;
    }
    Fwd Gf{this, ctx, &a2, B};
    std::vector<int> hl(7 * (size_t) B);
    for (int b = 0; b < B; b++) {
        hl[b] = T[b]; hl[B + b] = 2 * T[b]; hl[2 * B + b] = 20 * T[b]; hl[3 * B + b] = 120 * T[b] + 1; hl[4 * B + b] = 600 * T[b];
        hl[5 * B + b] = 2 * T[b] + 1; hl[6 * B + b] = 20 * T[b] + 1;      // polyphase GEMM rows per utterance (Lin + 1)
    }
    int * d_len = Gf.al<int>(7 * (size_t) B);
    int * d_idx = Gf.al<int>((size_t) B * L1);
    if (Gf.fail) return 1;
    B2_CUDA(cudaMemcpyAsync(d_len, hl.data(), hl.size() * 4, cudaMemcpyHostToDevice, st));
    B2_CUDA(cudaStreamSynchronize(st));
    const int * lT = d_len, * l2T = d_len + B, * l20 = d_len + 2 * B, * l120 = d_len + 3 * B, * lS = d_len + 4 * B;
    const int * lq[2] = {d_len + 5 * B, d_len + 6 * B};
    if (build_alignment(ctx, d_lens, B, Nmax, d_ntok, L1, d_idx, d_T)) return 1;

    // B0/B1: en = gather(d) -> shared bi-LSTM
    __half * en16 = Gf.al<__half>((size_t) B * L1 * 640); float * shared = Gf.al<float>((size_t) B * L1 * 512);
    if (Gf.fail) return 1;
    if (gather_rows(ctx, cur, 640, Nmax, d_idx, 640, B, L1, lT, nullptr, 0, en16, 640, 640)) return 1;
    if (Gf.lstm(shared_lstm, en16, 640, L1, lT, Tmax, shared, 512, 0, nullptr, 0, 0)) return 1;
    if (Gf.tap("shared", shared, (int64_t) B * L1, 512, 512, L1)) return 1;

    // B2: F0 / N curves
    float * f0 = Gf.al<float>((size_t) B * L2); float * nc = Gf.al<float>((size_t) B * L2);
    if (Gf.fail) return 1;
    for (int br = 0; br < 2; br++) {
        const AdaBlock * blk = br == 0 ? f0_blocks : n_blocks;
        const size_t mark = a2.off;
        float * y0 = Gf.al<float>((size_t) B * L1 * 512); float * y1 = Gf.al<float>((size_t) B * L2 * 256); float * y2 = Gf.al<float>((size_t) B * L2 * 256);
        __half * y16 = Gf.al<__half>((size_t) B * L2 * 256);
        if (Gf.fail) return 1;
        if (Gf.ada_block(blk[0], gbP, sty_n[0], shared, 512, L1, lT, L1, lT, y0, 512, 0)) return 1;
        if (Gf.ada_block(blk[1], gbP, sty_n[0], y0, 512, L1, lT, L2, l2T, y1, 256, 0)) return 1;
        if (Gf.ada_block(blk[2], gbP, sty_n[0], y1, 256, L2, l2T, L2, l2T, y2, 256, 0)) return 1;
        if (cast_rows(ctx, y2, 256, 256, B, L2, l2T, L2, 0, 1.0f, y16, 256, 256)) return 1;
        if (Gf.gemm(y16, 256, br == 0 ? f0_proj : n_proj, br == 0 ? f0_proj_b : n_proj_b, L2, L2, l2T, l2T, 1, 1, 0, br == 0 ? f0 : nc, 1, 0)) return 1;
        a2.off = mark;   // branch scratch is dead (stream order keeps this safe)
    }
    if (Gf.tap("f0", f0, (int64_t) B * L2, 1, 1, L2)) return 1;
    if (Gf.tap("n", nc, (int64_t) B * L2, 1, 1, L2)) return 1;

    // B3: text encoder (model.cpp:1196-1206)
    float * t_en = Gf.al<float>(rows1 * 512);
    {
        const size_t mark = a2.off;
        __half * e16 = Gf.al<__half>(rows1 * 512); float * cf = Gf.al<float>(rows1 * 512);
        if (Gf.fail) return 1;
        if (embed_rows_h(ctx, d_tok, d_tokoff, text_embd, 512, B, Nmax, d_ntok, e16, 512)) return 1;
        for (int i = 0; i < 3; i++) {
            if (Gf.gemm(e16, 512, te_conv[i], te_b[i], Nmax, Nmax, d_ntok, d_ntok, 1, 1, 2, cf, 512, 0)) return 1;
            RowNormParams rn;
            rn.x = cf; rn.ldx = 512; rn.C = 512; rn.B = B; rn.Lmax = Nmax; rn.len = d_ntok; rn.eps = 1e-5f; rn.mode = LN_AFFINE;
            rn.w = te_gamma[i]; rn.bias = te_beta[i]; rn.lrelu02 = 1; rn.outH = e16; rn.ldoh = 512;
            if (row_norm(ctx, rn)) return 1;
        }
        if (Gf.lstm(text_lstm, e16, 512, Nmax, d_ntok, Nmax, t_en, 512, 0, nullptr, 0, 0)) return 1;
        a2.off = mark;
    }
    if (Gf.tap("t_en", t_en, rows1, 512, 512, Nmax)) return 1;

    // B4: decoder (model.cpp:1215-1231)
    float * dec = Gf.al<float>((size_t) B * L2 * 512);
    {
        const size_t mark = a2.off;
        // concat buffers are padded to a row pitch that is a multiple of 4 floats (514 -> 516, 1090 -> 1092) so every kernel takes its
        // 16-byte path; the pad columns are never read as channels
        constexpr int LD514 = 516, LD1090 = 1092;
        float * x0 = Gf.al<float>((size_t) B * L1 * LD514); __half * asr16 = Gf.al<__half>((size_t) B * L1 * 512);
        float * side = Gf.al<float>((size_t) B * L1 * 66);
        float * xin[2] = {Gf.al<float>((size_t) B * L1 * LD1090), Gf.al<float>((size_t) B * L1 * LD1090)};
        if (Gf.fail) return 1;
        if (gather_rows(ctx, t_en, 512, Nmax, d_idx, 512, B, L1, lT, x0, LD514, asr16, 512, 512)) return 1;        // asr = t_en . mask
        if (curve_conv_s2(ctx, f0, L2, B, lT, L1, l2T, f0_conv_w, f0_conv_b, side, 66, 64)) return 1;
        if (curve_conv_s2(ctx, nc, L2, B, lT, L1, l2T, n_conv_w, n_conv_b, side, 66, 65)) return 1;
        if (Gf.gemm(asr16, 512, asr_conv, asr_conv_b, L1, L1, lT, lT, 1, 1, 0, side, 66, 0)) return 1;            // asr_res
        if (copy_cols(ctx, side, 66, 64, x0, LD514, 512, 2, B, L1, lT)) return 1;
        if (Gf.tap("dec_in", x0, (int64_t) B * L1, 514, LD514, L1)) return 1;
        if (copy_cols(ctx, side, 66, 0, xin[0], LD1090, 1024, 66, B, L1, lT)) return 1;
        if (copy_cols(ctx, side, 66, 0, xin[1], LD1090, 1024, 66, B, L1, lT)) return 1;
        if (Gf.ada_block(enc_block, gbD, sty_n[1], x0, LD514, L1, lT, L1, lT, xin[0], LD1090, 0)) return 1;
        for (int i = 0; i < 4; i++) {
            const size_t mk2 = a2.off;
            if (i < 3) { if (Gf.ada_block(dec_blocks[i], gbD, sty_n[1], xin[i & 1], LD1090, L1, lT, L1, lT, xin[(i + 1) & 1], LD1090, 0)) return 1; }
            else       { if (Gf.ada_block(dec_blocks[i], gbD, sty_n[1], xin[i & 1], LD1090, L1, lT, L2, l2T, dec, 512, 0)) return 1; }
            a2.off = mk2;
        }
        a2.off = mark;
    }
    if (Gf.tap("dec", dec, (int64_t) B * L2, 512, 512, L2)) return 1;

    // B5: harmonic source + STFT (model.cpp:173-206)
    const int Fmax = L4;
    const int hsp = nconv[0].w.CinPad;   // 22 channels padded for the GEMM operand
    // fp16 STFT operand: per-utterance pitch Fp = Fmax rounded up to the strided noise conv's stride (its polyphase view needs whole rows)
    int fp_mult = 1;
    for (int i = 0; i < 2; i++) if (nconv[i].poly) fp_mult = fp_mult * nconv[i].stride / std::__gcd(fp_mult, nconv[i].stride);
    const int Fp = round_up(Fmax, fp_mult);
    float * har = Gf.al<float>((size_t) B * S); __half * hs16 = Gf.al<__half>((size_t) B * Fp * hsp);
    float * phase = Gf.al<float>((size_t) B * 9 * L2);
    float * hsF = taps_on || overrides.count("har_spec") ? Gf.al<float>((size_t) B * Fmax * 22) : nullptr;
    if (Gf.fail) return 1;
    {
        SourceParams sp;
        sp.f0 = f0; sp.B = B; sp.L2max = L2; sp.len2 = l2T; sp.noise_skip = d_skip; memcpy(sp.w_src, m_src_w, 36); sp.b_src = m_src_b;
        sp.phase = phase; sp.har = har; sp.Smax = S;
        if (source_har(ctx, sp)) return 1;
        if (Gf.tap("har", har, B, S, S, S)) return 1;
        if (stft20(ctx, har, S, B, lS, Fmax, hs16, hsp, hsp, hsF, 22, Fp)) return 1;
        if (hsF) {
            if (Gf.tap("har_spec", hsF, (int64_t) B * Fmax, 22, 22, Fmax)) return 1;
            if (overrides.count("har_spec")) { if (cast_rows(ctx, hsF, 22, 22, B, Fmax, l120, Fp, 0, 1.0f, hs16, hsp, hsp)) return 1; }
        }
        if (fp_mult > 1 && zero_rows_past_end(ctx, hs16, hsp, hsp, B, Fp, l120)) return 1;   // the polyphase conv cannot mask rows per utterance
    }

    // B6/B7: generator stages (model.cpp:208-230)
    const float * gin = dec; int gin_L = L2; const int * gin_len = l2T;
    float * stage_out = nullptr;
    for (int i = 0; i < 2; i++) {
        const int Lo = i == 0 ? P0 : P1; const int * lo_len = i == 0 ? l20 : l120; const int C = ups[i].Cout;
        const int Cp = res[3 * i].c1[0].CinPad;
        float * ubuf = Gf.al<float>((size_t) B * Lo * C + 64 * C); float * xs = Gf.al<float>((size_t) B * Lo * C); float * curg = Gf.al<float>((size_t) B * Lo * C);
        float * u = ubuf;
        float * scr[3] = {Gf.al<float>((size_t) B * Lo * C), Gf.al<float>((size_t) B * Lo * C), Gf.al<float>((size_t) B * Lo * C)};
        float * acc[2] = {Gf.al<float>((size_t) B * Lo * C), Gf.al<float>((size_t) B * Lo * C)};
        __half * a16 = Gf.al<__half>((size_t) B * Lo * Cp);
        double * sums2[3] = {Gf.al<double>((size_t) B * C * 2), Gf.al<double>((size_t) B * C * 2), Gf.al<double>((size_t) B * C * 2)};
        double * sums_cur = Gf.al<double>((size_t) B * C * 2);
        float * part = Gf.al<float>((size_t) B * cdiv(Lo, 128) * C * 2);
        if (Gf.fail) return 1;
        if (ups[i].poly) {
            // lrelu(0.1) -> split fp16 hi/lo -> one tensor-core GEMM over all `stride` output phases (see Kokoro::prepare)
            const int Lq = gin_L + 1, s = ups[i].stride, refl = i == 1 ? 1 : 0;
            const size_t mk = a2.off;
            __half * a3 = Gf.al<__half>((size_t) B * Lq * 3 * ups[i].Cin);
            if (Gf.fail) return 1;
            if (split3_rows(ctx, gin, ups[i].Cin, ups[i].Cin, B, gin_L, gin_len, 0.1f, a3, Lq)) return 1;
            if (Gf.gemm(a3, 3 * ups[i].Cin, ups[i].w3, ups[i].b_rep, Lq, Lq, gin_len, lq[i], 1, 1, 1, ubuf, s * C, 0)) return 1;
            a2.off = mk;
            u = ubuf + (size_t) (ups[i].pad - refl) * C;          // out[o] = Y[o + pad]; with the 1-sample left reflect pad out'[o] = out[o-1]
            if (refl) B2_CUDA(cudaMemcpy2DAsync(u, (size_t) Lo * C * 4, u + 2 * (size_t) C, (size_t) Lo * C * 4, (size_t) C * 4, B, cudaMemcpyDeviceToDevice, st));   // out'[0] = out[1]
        } else {
            if (convt_cl(ctx, gin, ups[i].Cin, ups[i].Cin, B, gin_L, gin_len, ups[i].w, ups[i].b, ups[i].K, C, ups[i].stride, ups[i].pad, 0.1f, i == 1 ? 1 : 0, u, C, Lo, lo_len)) return 1;
        }
        // noise conv; where it is a stride-1 layer its epilogue also yields the InstanceNorm statistics the noise resblock starts with
        const bool xs_stats = nconv[i].stride == 1 && nconv[i].w.KW == 1;
        double * sums_xs = xs_stats ? Gf.al<double>((size_t) B * C * 2) : nullptr;
        if (Gf.fail) return 1;
        if (xs_stats) { if (Gf.gemm_stats(hs16, hsp, nconv[i].w, nconv[i].b, Lo, lo_len, 1, nconv[i].pad, xs, nullptr, nullptr, 0.f, part, sums_xs, Fp, l120, false)) return 1; }
        else if (nconv[i].poly && Fp % nconv[i].stride == 0) {
            if (Gf.gemm(hs16, nconv[i].stride * hsp, nconv[i].wp, nconv[i].b, Fp / nconv[i].stride, Lo, nullptr, lo_len, 1, 1, nconv[i].padp, xs, C, 0, nullptr, 0, 0, nullptr, 0,
                        nullptr, 0, 0.f, ACT_NONE, -1, true)) return 1;
        } else if (Gf.gemm(hs16, hsp, nconv[i].w, nconv[i].b, Fp, Lo, l120, lo_len, nconv[i].stride, 1, nconv[i].pad, xs, C, 0)) return 1;
        if (Gf.gen_resblock(nres[i], gbD, sty_n[1], xs, Lo, lo_len, curg, u, 0.f, scr, a16, sums2, part, sums_xs, sums_cur)) return 1;   // cur = up + x_source
        if (Gf.tap(i == 0 ? "gen_in0" : "gen_in1", curg, (int64_t) B * Lo, C, C, Lo)) return 1;
        for (int j = 0; j < 3; j++) {
            float * dst = acc[j & 1];
            if (Gf.gen_resblock(res[3 * i + j], gbD, sty_n[1], curg, Lo, lo_len, dst, j == 0 ? nullptr : acc[(j + 1) & 1], j == 2 ? 3.0f : 0.f, scr, a16, sums2, part,
                                overrides.count(i == 0 ? "gen_in0" : "gen_in1") ? nullptr : sums_cur, nullptr)) return 1;
        }
        stage_out = acc[0];   // j == 2 writes acc[0]
        if (Gf.tap(i == 0 ? "gen_out0" : "gen_out1", stage_out, (int64_t) B * Lo, C, C, Lo)) return 1;
        gin = stage_out; gin_L = Lo; gin_len = lo_len;
    }
    // B8: conv_post -> exp / sin -> iSTFT (model.cpp:232-241)
    __half * p16 = Gf.al<__half>((size_t) B * P1 * conv_post.CinPad); float * specph = Gf.al<float>((size_t) B * P1 * 22);
    float * pcm_d = Gf.al<float>((size_t) B * S);
    if (Gf.fail) return 1;
    if (cast_rows(ctx, stage_out, 128, 128, B, P1, l120, P1, 0, 0.01f, p16, conv_post.CinPad, conv_post.CinPad)) return 1;
    if (Gf.gemm(p16, conv_post.CinPad, conv_post, conv_post_b, P1, P1, l120, l120, 1, 1, post_pad, specph, 22, 0, nullptr, 0, 0, nullptr, 0, nullptr, 0, 0.f, ACT_EXP_SIN_11)) return 1;
    if (Gf.tap("spec", specph, (int64_t) B * P1, 22, 22, P1)) return 1;
    if (istft20(ctx, specph, 22, B, l120, P1, pcm_d, S)) return 1;
    if (Gf.tap("pcm", pcm_d, B, S, S, S)) return 1;
    B2_CUDA(cudaEventRecord(ev[2], st));

    // ---- D2H into the runner-owned pinned buffer (borrowed by the caller until the next call, like tts_response.data)
    // one copy of the padded [B][S] block when the padding is small (a copy per utterance costs ~10 us of launch overhead each);
    // ragged batches whose padding would add > 25 % bytes are copied utterance by utterance, packed
    last_pcm_dev = pcm_d; last_pcm_stride = S;
    if (keep_on_device) {                                       // the caller takes the PCM from device memory (NCCL gather): no host copy
        for (int b = 0; b < B; b++) { if (pcm) pcm[b] = nullptr; if (n_samples) n_samples[b] = (int64_t) T[b] * 600; }
        B2_CUDA(cudaEventRecord(ev[3], st));
        B2_CUDA(cudaStreamSynchronize(st));
        cudaEventElapsedTime(&timings[0], ev[0], ev[1]);
        cudaEventElapsedTime(&timings[1], ev[1], ev[2]);
        cudaEventElapsedTime(&timings[2], ev[0], ev[3]);
        return 0;
    }
    const bool one_copy = (size_t) B * S <= Ssum + Ssum / 4;
    const size_t need_pinned = one_copy ? (size_t) B * S : Ssum;
    if (pcm_pinned_cap < need_pinned) {
        if (pcm_pinned) cudaFreeHost(pcm_pinned);
        pcm_pinned = nullptr; pcm_pinned_cap = 0;
        B2_CUDA(cudaMallocHost(&pcm_pinned, std::max<size_t>(need_pinned, 1) * 4));
        pcm_pinned_cap = need_pinned;
    }
    if (one_copy) B2_CUDA(cudaMemcpyAsync(pcm_pinned, pcm_d, (size_t) B * S * 4, cudaMemcpyDeviceToHost, st));
    size_t offp = 0;
    for (int b = 0; b < B; b++) {
        const size_t n = (size_t) T[b] * 600;
        if (one_copy) offp = (size_t) b * S;
        else if (n) B2_CUDA(cudaMemcpyAsync(pcm_pinned + offp, pcm_d + (size_t) b * S, n * 4, cudaMemcpyDeviceToHost, st));
        if (pcm) pcm[b] = pcm_pinned + offp;
        if (n_samples) n_samples[b] = (int64_t) n;
        offp += n;
    }
    B2_CUDA(cudaEventRecord(ev[3], st));
    B2_CUDA(cudaStreamSynchronize(st));
    cudaEventElapsedTime(&timings[0], ev[0], ev[1]);
    cudaEventElapsedTime(&timings[1], ev[1], ev[2]);
    cudaEventElapsedTime(&timings[2], ev[0], ev[3]);
    return 0;
}

}  // namespace b2
