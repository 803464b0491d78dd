// oracle/ref_ops_driver.cpp -- TEST INFRASTRUCTURE (never linked into the product).
//
// Runs single operators of the UNMODIFIED reference (the ggml fork's patched ops and
// the util.cpp wrappers: reference ggml/include/ggml.h:935-971,1618-1651,1743-1793;
// src/util.cpp:86-137,203-217) on inputs read from raw float32 files, and writes the
// result as raw float32.  Used to generate the op-level known-answer vectors under
// tests/golden/ (script: tests/golden/make_golden.py) that pin oracle/kokoro_port.py
// and the CUDA kernels.
//
// usage: ops_ref <op> <out.f32> <args...>       (see the table in main())
#include "ggml.h"
#include "ggml-cpu.h"
#include "util.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

static ggml_context * g_ctx;

static std::vector<float> read_f32(const char * path) {
    FILE * f = fopen(path, "rb");
    if (!f) { fprintf(stderr, "cannot read %s\n", path); exit(2); }
    fseek(f, 0, SEEK_END); long n = ftell(f); fseek(f, 0, SEEK_SET);
    std::vector<float> v(n / 4);
    if (fread(v.data(), 4, v.size(), f) != v.size()) { fprintf(stderr, "short read %s\n", path); exit(2); }
    fclose(f);
    return v;
}

// new F32 (or F16 when f16=true) tensor with ggml ne order, filled from file
static ggml_tensor * load(const char * path, int64_t ne0, int64_t ne1 = 1, int64_t ne2 = 1, int64_t ne3 = 1, bool f16 = false) {
    std::vector<float> v = read_f32(path);
    if ((int64_t) v.size() != ne0 * ne1 * ne2 * ne3) {
        fprintf(stderr, "%s: have %zu floats, want %lld\n", path, v.size(), (long long) (ne0 * ne1 * ne2 * ne3)); exit(2);
    }
    ggml_tensor * t = ggml_new_tensor_4d(g_ctx, f16 ? GGML_TYPE_F16 : GGML_TYPE_F32, ne0, ne1, ne2, ne3);
    if (f16) ggml_fp32_to_fp16_row(v.data(), (ggml_fp16_t *) t->data, v.size());
    else memcpy(t->data, v.data(), v.size() * 4);
    return t;
}

static void run_and_write(ggml_tensor * out, const char * path, int threads) {
    out = ggml_cont(g_ctx, out);
    ggml_cgraph * gf = ggml_new_graph(g_ctx);
    ggml_build_forward_expand(gf, out);
    ggml_graph_compute_with_ctx(g_ctx, gf, threads);
    FILE * f = fopen(path, "wb");
    fwrite(out->data, 4, ggml_nelements(out), f);
    fclose(f);
    printf("OUT ne=[%lld,%lld,%lld,%lld]\n", (long long) out->ne[0], (long long) out->ne[1], (long long) out->ne[2], (long long) out->ne[3]);
}

#define A(i) argv[3 + (i)]
#define I(i) atoll(argv[3 + (i)])
#define F(i) atof(argv[3 + (i)])

int main(int argc, char ** argv) {
    if (argc < 4) { fprintf(stderr, "usage: %s op out.f32 args...\n", argv[0]); return 2; }
    std::string op = argv[1]; const char * out = argv[2];
    ggml_init_params ip = { (size_t) 2048 * 1024 * 1024, nullptr, false };
    g_ctx = ggml_init(ip);
    int threads = 4;
    int n = argc - 3;
    ggml_tensor * r = nullptr;
    if (op == "convt1d" && n == 11) {            // kernel K CoutG Cin data L s p op groups f16
        ggml_tensor * k = load(A(0), I(1), I(2), I(3), 1, I(10) != 0);
        ggml_tensor * x = load(A(4), I(5), I(3));
        r = ggml_conv_transpose_1d(g_ctx, k, x, I(6), I(7), 1, I(8), I(9));
    } else if (op == "conv1d" && n == 10) {      // kernel K Cin Cout data L s p d f16
        ggml_tensor * k = load(A(0), I(1), I(2), I(3), 1, I(9) != 0);
        ggml_tensor * x = load(A(4), I(5), I(2));
        r = ggml_conv_1d(g_ctx, k, x, I(6), I(7), I(8));
    } else if (op == "conv1d_dw" && n == 9) {    // kernel K C data L s p d f16   (kernel ne = [K,1,C])
        ggml_tensor * k = load(A(0), I(1), 1, I(2), 1, I(8) != 0);
        ggml_tensor * x = load(A(3), I(4), I(2));
        r = ggml_conv_1d_dw(g_ctx, k, x, I(5), I(6), I(7));
    } else if (op == "mul_mat" && n == 6) {      // W K N  X M f16     W ne=[K,N], X ne=[K,M] -> [N,M]
        ggml_tensor * w = load(A(0), I(1), I(2), 1, 1, I(5) != 0);
        ggml_tensor * x = load(A(3), I(1), I(4));
        r = ggml_mul_mat(g_ctx, w, x);
    } else if (op == "stft" && n == 6) {         // data L nfft hop abs_angle one_sided
        ggml_tensor * x = load(A(0), I(1));
        std::vector<float> w; hann_window(I(2), w);
        ggml_tensor * wt = ggml_new_tensor_1d(g_ctx, GGML_TYPE_F32, I(2)); memcpy(wt->data, w.data(), w.size() * 4);
        r = stft(g_ctx, x, wt, I(2), I(3), I(4) != 0, I(5) != 0);
    } else if (op == "istft" && n == 5) {        // data nbins frames nfft hop   (data ne = [nbins, frames, 1, 2], mag then phase)
        ggml_tensor * x = load(A(0), I(1), I(2), 1, 2);
        std::vector<float> w; hann_window(I(3), w);
        ggml_tensor * wt = ggml_new_tensor_1d(g_ctx, GGML_TYPE_F32, I(3)); memcpy(wt->data, w.data(), w.size() * 4);
        int64_t n_out = (I(2) - 1) * I(4);
        ggml_tensor * wss = ggml_new_tensor_1d(g_ctx, GGML_TYPE_F32, n_out);
        compute_window_squared_sum(I(3), I(4), n_out / I(4), (float *) wss->data, w.data());
        r = istft(g_ctx, x, wss, wt, I(3), I(4), true, true);
    } else if (op == "cumsum" && n == 3) {       // data L R
        r = ggml_cumsum(g_ctx, load(A(0), I(1), I(2)));
    } else if (op == "mod" && n == 3) {          // data N val
        r = ggml_mod(g_ctx, load(A(0), I(1)), F(2));
    } else if (op == "round" && n == 2) {
        r = ggml_round(g_ctx, load(A(0), I(1)));
    } else if (op == "reciprocal" && n == 2) {
        r = ggml_reciprocal(g_ctx, load(A(0), I(1)));
    } else if (op == "upscale_linear" && n == 4) { // data L R factor
        r = ggml_upscale_linear(g_ctx, load(A(0), I(1), I(2)), I(3));
    } else if (op == "upscale" && n == 4) {      // data L R newL
        ggml_tensor * x = load(A(0), I(1), I(2));
        r = ggml_upscale_ext(g_ctx, x, I(3), x->ne[1], 1, 1);
    } else if (op == "snake" && n == 4) {        // alpha C data L     (alpha ne = [1,C,1], data ne=[L,C])
        ggml_tensor * a = load(A(0), 1, I(1), 1);
        r = snake_1d(g_ctx, a, load(A(2), I(3), I(1)));
    } else if (op == "norm" && n == 4) {         // data L R eps
        r = ggml_norm(g_ctx, load(A(0), I(1), I(2)), F(3));
    } else if (op == "gelu" && n == 2) {
        r = ggml_gelu(g_ctx, load(A(0), I(1)));
    } else if (op == "sigmoid" && n == 2) {
        r = ggml_sigmoid(g_ctx, load(A(0), I(1)));
    } else if (op == "tanh" && n == 2) {
        r = ggml_tanh(g_ctx, load(A(0), I(1)));
    } else if (op == "softmax" && n == 4) {      // data L R scale
        r = ggml_soft_max_ext(g_ctx, load(A(0), I(1), I(2)), nullptr, F(3), 0.0f);
    } else if (op == "leaky_relu" && n == 3) {
        r = ggml_leaky_relu(g_ctx, load(A(0), I(1)), F(2), false);
    } else if (op == "uniform" && n == 1) {      // first N draws of the reference's static uniform engine (src/util.cpp:66-72)
        std::vector<float> v(I(0)); random_uniform_gen(I(0), v.data());
        FILE * f = fopen(out, "wb"); fwrite(v.data(), 4, v.size(), f); fclose(f); printf("OUT ne=[%lld,1,1,1]\n", (long long) v.size()); return 0;
    } else if (op == "wss" && n == 3) {          // nfft hop frames -> window squared sum
        std::vector<float> w; hann_window(I(0), w); std::vector<float> v(I(1) * I(2));
        compute_window_squared_sum(I(0), I(1), I(2), v.data(), w.data());
        FILE * f = fopen(out, "wb"); fwrite(v.data(), 4, v.size(), f); fclose(f); printf("OUT ne=[%lld,1,1,1]\n", (long long) v.size()); return 0;
    } else {
        fprintf(stderr, "bad op/arity: %s (%d args)\n", op.c_str(), n); return 2;
    }
    run_and_write(r, out, threads);
    return 0;
}
