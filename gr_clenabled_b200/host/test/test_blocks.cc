// Drives the gr::clenabled block classes the way the reference's CLI tools do
// (lib/test_clenabled.cc instantiates the _impl classes directly and calls the work
// functions, no scheduler): make(), work()/general_work(), known answers.
#include <clenabled/blocks.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <stdexcept>

using namespace gr::clenabled;
static int failures = 0;
#define CHECK(cond)                                                                     \
    do {                                                                                \
        if (!(cond)) {                                                                  \
            printf("FAIL %s:%d: %s\n", __FILE__, __LINE__, #cond);                      \
            failures++;                                                                 \
        }                                                                               \
    } while (0)

template <class E, class F>
bool throws(F f)
{
    try {
        f();
    } catch (const E &) {
        return true;
    } catch (...) {
        return false;
    }
    return false;
}

int main()
{
    const int GPU = OCLTYPE_GPU, FIRST = OCLDEVICESELECTOR_FIRST;
    // --- Multiply Const: (1.0, 0.5) * 2 = (2.0, 1.0)   (test_clenabled.cc:1193,1336,1351-1352)
    {
        auto blk = clMathConst::make(DTYPE_COMPLEX, GPU, FIRST, 0, 0, 2.0f, MATHOP_MULTIPLY);
        std::vector<gr_complex> in(8192, gr_complex(1.0f, 0.5f)), out(8192);
        gr_vector_const_void_star iv{in.data()};
        gr_vector_void_star ov{out.data()};
        CHECK(blk->work(8192, iv, ov) == 8192);
        bool ok = true;
        for (auto &v : out) ok = ok && v == gr_complex(2.0f, 1.0f);
        CHECK(ok);
        blk->set_k(3.0f);
        CHECK(blk->k() == 3.0f);
        CHECK(blk->output_multiple() > 0);
    }
    // --- Multiply (2 -> 1)
    {
        auto blk = clMathOp::make(DTYPE_COMPLEX, GPU, FIRST, 0, 0, MATHOP_MULTIPLY_CONJUGATE);
        std::vector<gr_complex> a(1000, gr_complex(1.0f, 2.0f)), b(1000, gr_complex(3.0f, -1.0f)), c(1000);
        gr_vector_const_void_star iv{a.data(), b.data()};
        gr_vector_void_star ov{c.data()};
        CHECK(blk->work(1000, iv, ov) == 1000);
        CHECK(c[0] == gr_complex(1.0f, 2.0f) * std::conj(gr_complex(3.0f, -1.0f)));
        CHECK(blk->work(0, iv, ov) == 0);
    }
    // --- FFT: the tone of test_clenabled.cc:835-851 is i*e^{-2 pi i n/N}: one bin, i*N at N-1
    {
        const int N = 8192, nvec = 3;
        auto blk = clFFT::make(N, CLFFT_FORWARD_DIR, std::vector<float>(), DTYPE_COMPLEX, GPU, FIRST, 0, 0, 0, 1, false);
        std::vector<gr_complex> in(N * nvec), out(N * nvec);
        for (int v = 0; v < nvec; v++)
            for (int i = 0; i < N; i++)
                in[v * N + i] = gr_complex((float)sin(2 * M_PI * i / N), (float)cos(2 * M_PI * i / N));
        gr_vector_const_void_star iv{in.data()};
        gr_vector_void_star ov{out.data()};
        CHECK(blk->work(nvec, iv, ov) == nvec);        // items are vectors
        double worst = 0;
        for (int v = 0; v < nvec; v++)
            for (int i = 0; i < N; i++) {
                gr_complex want = (i == N - 1) ? gr_complex(0.0f, (float)N) : gr_complex(0, 0);
                worst = std::max(worst, (double)std::abs(out[v * N + i] - want));
            }
        CHECK(worst < 1e-5 * N);
        CHECK(throws<std::runtime_error>([&] {
            clFFT::make(N, CLFFT_FORWARD_DIR, std::vector<float>(100, 1.0f), DTYPE_COMPLEX, GPU, FIRST, 0, 0);
        }));
    }
    // --- Filter: impulse response = taps, both kernels; decimation; tap swap
    for (int use_time = 0; use_time < 2; use_time++) {
        std::vector<float> taps(256);
        for (int i = 0; i < 256; i++) taps[i] = i / 1000.0f;            // test-clfilter.cc:98-100
        auto blk = clFilter::make(GPU, FIRST, 0, 0, 1, taps, 1, 0, use_time != 0);
        std::vector<gr_complex> in(8192, gr_complex(0, 0)), out(8192);
        in[0] = gr_complex(1.0f, 0.0f);
        gr_vector_const_void_star iv{in.data()};
        gr_vector_void_star ov{out.data()};
        CHECK(blk->work(8192, iv, ov) == 8192);
        double worst = 0;
        for (int i = 0; i < 8192; i++) worst = std::max(worst, (double)std::abs(out[i] - gr_complex(i < 256 ? taps[i] : 0.0f, 0.0f)));
        CHECK(worst < 1e-5);
        CHECK(blk->taps().size() == 256);
        auto dec = clFilter::make(GPU, FIRST, 0, 0, 4, taps, 1, 0, use_time != 0);
        CHECK(dec->decimation() == 4);
        CHECK(dec->work(2048, iv, ov) == 2048);        // consumes 8192, produces 2048
        CHECK(std::abs(out[1] - gr_complex(taps[4], 0)) < 1e-5);
    }
    // --- Polyphase channelizer: item accounting of general_work
    {
        const int M = 64, T = 128, buf = 4096;
        std::vector<float> taps(T, 1.0f / T);
        std::vector<int> map{5, 0, 63};
        auto blk = clPolyphaseChannelizer::make(GPU, FIRST, 0, 0, taps, buf, M, M, map);
        CHECK(blk->history() == (unsigned)T);
        CHECK(blk->output_multiple() == 3 * buf / M);
        std::vector<gr_complex> in(buf + T, gr_complex(1.0f, 0.0f)), out(3 * buf / M);
        gr_vector_int ni{(int)in.size()};
        gr_vector_const_void_star iv{in.data()};
        gr_vector_void_star ov{out.data()};
        CHECK(blk->general_work(3 * buf / M, ni, iv, ov) == 3 * buf / M);
        CHECK(blk->consumed_each() == buf);
        // DC input: all energy in channel 0 = sum(taps) * ... = 1 per arm * M arms / M
        CHECK(std::abs(out[1] - gr_complex(1.0f, 0.0f)) < 1e-4 && std::abs(out[0]) < 1e-4 && std::abs(out[2]) < 1e-4);
        CHECK(throws<std::invalid_argument>([&] { clPolyphaseChannelizer::make(GPU, FIRST, 0, 0, taps, 100, M, M, map); }));
    }
    // --- X-engine: two polarisations through the stream ports, PDU out
    {
        const int A = 4, F = 8, T = 64, npol = 2;
        auto blk = clXEngine::make(GPU, FIRST, 0, 0, false, DTYPE_BYTE, npol, A, 1, 0, F, T, {});
        std::vector<std::vector<signed char>> ports(A * npol, std::vector<signed char>((size_t)T * F * 2));
        unsigned seed = 1;
        for (auto &p : ports)
            for (auto &v : p) {
                seed = seed * 1664525u + 1013904223u;
                v = (signed char)((int)(seed >> 24) % 100 - 50);
            }
        gr_vector_const_void_star iv;
        for (auto &p : ports) iv.push_back(p.data());
        gr_vector_int ni(A * npol, T);
        gr_vector_void_star ov;
        CHECK(blk->general_work(40, ni, iv, ov) == 40);          // partial integration: no PDU yet
        CHECK(blk->published("xcorr").empty());
        gr_vector_const_void_star iv2;
        for (auto &p : ports) iv2.push_back(p.data() + 40 * F * 2);
        CHECK(blk->general_work(1000, ni, iv2, ov) == 24);       // only what completes the integration
        // the correlation and its read-back run asynchronously; the matrix is published by a later call or by stop()
        // (the reference publishes it when the next integration completes, lib/clXEngine_impl.cc:1062-1090)
        blk->stop();
        auto &msgs = blk->published("xcorr");
        CHECK(msgs.size() == 1);
        if (msgs.size() == 1) {
            CHECK(pmt::symbol_to_string(pmt::car(msgs[0])) == "triang_matrix");
            const auto &m = pmt::c32vector_elements(pmt::cdr(msgs[0]));
            CHECK((int)m.size() == F * (A * (A + 1) / 2) * npol * npol);
            // baseline (0,0), XX of channel 0 = sum |x|^2 / 127^2, imaginary part 0
            double p = 0;
            for (int t = 0; t < T; t++) {
                double re = ports[0][(size_t)t * F * 2], im = ports[0][(size_t)t * F * 2 + 1];
                p += re * re + im * im;
            }
            CHECK(std::abs(m[0].real() - p / (127.0 * 127.0)) < 1e-4 * p / (127.0 * 127.0) + 1e-6);
            CHECK(m[0].imag() == 0.0f);
        }
        CHECK(throws<std::out_of_range>([&] { clXEngine::make(GPU, FIRST, 0, 0, false, DTYPE_BYTE, 1, 1, 1, 0, F, T, {}); }));
        // file sink: raw cf32_le frames + JSON sidecar (lib/clXEngine_impl.cc:393-465)
        const char *base = "/tmp/clb200_xengine_test.bin";
        auto fblk = clXEngine::make(GPU, FIRST, 0, 0, false, DTYPE_BYTE, npol, A, 1, 0, F, T, {"a1", "a2", "a3", "a4"},
                                    true, base, 0, false, 1234, "obj", 1.0e9, 250e3);
        CHECK(fblk->general_work(T, ni, iv, ov) == T);
        CHECK(fblk->general_work(T, ni, iv, ov) == T);
        fblk->stop();
        FILE *fp = fopen(base, "rb");
        CHECK(fp != nullptr);
        if (fp) {
            fseek(fp, 0, SEEK_END);
            CHECK(ftell(fp) == 2L * F * (A * (A + 1) / 2) * npol * npol * (long)sizeof(gr_complex));
            fclose(fp);
        }
        FILE *js = fopen("/tmp/clb200_xengine_test.bin.json", "r");
        CHECK(js != nullptr);
        if (js) {
            char buf[2048];
            size_t n = fread(buf, 1, sizeof(buf) - 1, js);
            buf[n] = 0;
            fclose(js);
            std::string t(buf);
            CHECK(t.find("\"sync_timestamp\":1234") != std::string::npos);
            CHECK(t.find("\"num_baselines\":10") != std::string::npos);
            CHECK(t.find("\"antenna_names\":[\"a1\",\"a2\",\"a3\",\"a4\"]") != std::string::npos);
            CHECK(t.find("\"data_type\":\"cf32_le\"") != std::string::npos);
        }
    }
    // --- X-engine streaming: many small general_work() calls, results picked up on later calls; pipeline_integration
    //     sums integrations on the device; setDebug logs the kernel description through the block's logger
    {
        const int A = 3, F = 16, T = 32, NI = 6;
        auto blk = clXEngine::make(GPU, FIRST, 0, 0, true, DTYPE_BYTE, 1, A, 1, 0, F, T, {});
        CHECK(!blk->test_log_lines().empty() && blk->test_log_lines()[0].find("clXEngine 3 inputs") != std::string::npos);
        auto pip = clXEngine::make(GPU, FIRST, 0, 0, false, DTYPE_BYTE, 1, A, 1, 0, F, T, {}, false, "", 0, false, 0, "", 0.0,
                                   0.0, false, 3);
        std::vector<std::vector<signed char>> ports(A, std::vector<signed char>((size_t)NI * T * F * 2));
        unsigned seed = 7;
        for (auto &p : ports)
            for (auto &v : p) {
                seed = seed * 1664525u + 1013904223u;
                v = (signed char)((int)(seed >> 24) % 60 - 30);
            }
        gr_vector_void_star ov;
        gr_vector_int ni(A, 5);
        int pos = 0;
        while (pos < NI * T) {                                    // 5 time steps per call, never past an integration
            gr_vector_const_void_star iv;
            for (auto &p : ports) iv.push_back(p.data() + (size_t)pos * F * 2);
            const int want = std::min(5, T - pos % T);
            CHECK(blk->general_work(5, ni, iv, ov) == want);
            CHECK(pip->general_work(5, ni, iv, ov) == want);
            pos += want;
        }
        blk->stop();
        pip->stop();
        auto &msgs = blk->published("xcorr");
        auto &pm = pip->published("xcorr");
        CHECK((int)msgs.size() == NI);
        CHECK((int)pm.size() == NI / 3);
        const int nbl = A * (A + 1) / 2;
        for (int k = 0; k < (int)msgs.size(); k++) {
            const auto &m = pmt::c32vector_elements(pmt::cdr(msgs[k]));
            // baseline (1,0) of channel 2 by direct summation: V = sum_t x1 conj(x0) / 127^2
            double re = 0, im = 0;
            for (int t = 0; t < T; t++) {
                const size_t o = ((size_t)(k * T + t) * F + 2) * 2;
                const double ar = ports[1][o], ai = ports[1][o + 1], br = ports[0][o], bi = ports[0][o + 1];
                re += ar * br + ai * bi;
                im += ai * br - ar * bi;
            }
            const gr_complex v = m[(size_t)2 * nbl + 1];
            CHECK(std::abs(v.real() - re / 16129.0) < 1e-3 && std::abs(v.imag() - im / 16129.0) < 1e-3);
        }
        if ((int)msgs.size() == NI && (int)pm.size() == NI / 3)
            for (int g = 0; g < NI / 3; g++) {                    // a pipelined matrix = the sum of its three integrations
                const auto &s = pmt::c32vector_elements(pmt::cdr(pm[g]));
                double worst = 0;
                for (size_t i = 0; i < s.size(); i++) {
                    gr_complex acc(0, 0);
                    for (int j = 0; j < 3; j++) acc += pmt::c32vector_elements(pmt::cdr(msgs[3 * g + j]))[i];
                    worst = std::max(worst, (double)std::abs(acc - s[i]));
                }
                CHECK(worst < 1e-3);
            }
    }
    // --- clXEngine ATA synchroniser (lib/clXEngine_impl.cc:1158-1226): inputs whose first SNAP sequence tags differ
    //     are trimmed to the highest tag; once equal, "synctimestamp" is published and data flows
    {
        const int A = 2, F = 4, T = 16;
        CHECK(throws<std::out_of_range>([&] { clXEngine::make(GPU, FIRST, 0, 0, false, DTYPE_BYTE, 1, A, 1, 0, F, 24, {}, false, "", 0, true); }));
        auto blk = clXEngine::make(GPU, FIRST, 0, 0, false, DTYPE_BYTE, 1, A, 1, 0, F, T, {}, false, "", 0, true);
        CHECK(blk->output_multiple() == 16);
        std::vector<int8_t> a(T * F * 2, 1), b(T * F * 2, 1);
        gr_vector_const_void_star iv{a.data(), b.data()};
        gr_vector_void_star ov;
        gr_vector_int ni{T, T};
        blk->test_add_tag(0, 0, pmt::mp("seq"), pmt::from_uint64(1000));
        blk->test_add_tag(1, 0, pmt::mp("seq"), pmt::from_uint64(1016));
        CHECK(blk->general_work(T, ni, iv, ov) == 0);
        CHECK(blk->consumed().size() == 2 && blk->consumed()[0] == 16 && blk->consumed()[1] == 0);
        CHECK(blk->published("sync").empty());
        blk->test_set_read_offset(0, 16);                         // the scheduler advanced input 0 by 16 items
        blk->test_add_tag(0, 16, pmt::mp("seq"), pmt::from_uint64(1016));
        CHECK(blk->general_work(T, ni, iv, ov) == T);
        CHECK(blk->published("sync").size() == 1);
        if (!blk->published("sync").empty()) CHECK(pmt::to_uint64(pmt::cdr(blk->published("sync")[0])) == 1016);
        blk->stop();
        CHECK(blk->published("xcorr").size() == 1);
    }
    // --- clXCorrelate: a copy delayed by 5 samples -> corrective lag -5, published on port "corr"
    {
        const int L = 512, MS = 64;
        std::vector<float> ref(L), sig(L);
        for (int i = 0; i < L; i++) ref[i] = 0.1f + (float)((i * 2654435761u) >> 20 & 1023) / 1024.0f;
        for (int i = 0; i < L; i++) sig[i] = i >= 5 ? ref[i - 5] : 0.3f;
        auto blk = clXCorrelate::make(GPU, FIRST, 0, 0, false, 2, L, DTYPE_FLOAT, sizeof(float), MS, 1);
        gr_vector_const_void_star iv{ref.data(), sig.data()};
        gr_vector_void_star ov;
        CHECK(blk->output_multiple() == L);
        CHECK(blk->work(L - 1, iv, ov) == 0);
        CHECK(blk->work(L, iv, ov) == L);
        auto &msgs = blk->published("corr");
        CHECK(msgs.size() == 1);
        if (!msgs.empty()) {
            pmt::pmt_t meta = pmt::car(msgs[0]);
            auto lags = pmt::s32vector_elements(pmt::dict_ref(meta, pmt::mp("corrective_lags"), pmt::PMT_NIL));
            auto corr = pmt::f32vector_elements(pmt::dict_ref(meta, pmt::mp("corrvect"), pmt::PMT_NIL));
            CHECK(lags.size() == 1 && lags[0] == -5);
            CHECK(corr.size() == 1 && corr[0] > 0.99f);
        }
        CHECK(throws<std::invalid_argument>([&] { clXCorrelate::make(GPU, FIRST, 0, 0, false, 2, 511, DTYPE_FLOAT, 4, MS, 1); }));
        CHECK(throws<std::invalid_argument>([&] { clXCorrelate::make(GPU, FIRST, 0, 0, false, 2, L, DTYPE_FLOAT, 0, MS, 1); }));
    }
    // --- clxcorrelate_fft_vcf on time series: circular delay of 3 -> peak at N/2 - 3
    {
        const int N = 256;
        std::vector<gr_complex> a(N), b(N);
        for (int i = 0; i < N; i++) a[i] = gr_complex((float)((i * 2654435761u) >> 20 & 1023) / 512.0f - 1.0f, (float)((i * 40503u) & 255) / 128.0f - 1.0f);
        for (int i = 0; i < N; i++) b[i] = a[(i + N - 3) % N];
        auto blk = clxcorrelate_fft_vcf::make(N, 2, GPU, FIRST, 0, 0, 2);
        std::vector<float> out(N);
        gr_vector_const_void_star iv{a.data(), b.data()};
        gr_vector_void_star ov{out.data()};
        CHECK(blk->work(1, iv, ov) == 1);
        int best = 0;
        for (int i = 1; i < N; i++)
            if (out[i] > out[best]) best = i;
        CHECK(best == N / 2 - 3);
    }
    // --- clComplexFilter: an impulse returns the taps; clQuadratureDemod: constant rotation; clSignalSource
    {
        std::vector<gr_complex> taps{{1, 2}, {3, -1}, {0.5f, 0.25f}};
        auto blk = clComplexFilter::make(GPU, FIRST, 0, 0, 1, taps);
        std::vector<gr_complex> in(64, gr_complex(0, 0)), out(64);
        in[0] = gr_complex(1, 0);
        gr_vector_const_void_star iv{in.data()};
        gr_vector_void_star ov{out.data()};
        CHECK(blk->work(64, iv, ov) == 64);
        CHECK(out[0] == taps[0] && out[1] == taps[1] && out[2] == taps[2] && out[3] == gr_complex(0, 0));
        auto qd = clQuadratureDemod::make(2.0f, GPU, FIRST, 0, 0);
        std::vector<gr_complex> tone(1024);
        for (int i = 0; i < 1024; i++) tone[i] = std::polar(1.0f, 0.25f * (float)i);
        std::vector<float> ph(1024);
        gr_vector_const_void_star qi{tone.data()};
        gr_vector_void_star qo{ph.data()};
        CHECK(qd->work(1024, qi, qo) == 1024);
        CHECK(std::fabs(ph[10] - 0.5f) < 1e-4f && std::fabs(ph[1023] - 0.5f) < 1e-4f);
        auto src = clSignalSource::make(DTYPE_COMPLEX, GPU, FIRST, 0, 0, 1000.0, SIGSOURCE_COS, 10.0, 2.0f);
        std::vector<gr_complex> s(100);
        gr_vector_const_void_star none;
        gr_vector_void_star so{s.data()};
        CHECK(src->work(100, none, so) == 100);
        CHECK(std::abs(s[25] - std::polar(2.0f, (float)(2.0 * M_PI * 10.0 / 1000.0 * 25))) < 1e-5f);
    }
    if (failures == 0) printf("ALL OK\n");
    return failures == 0 ? 0 : 1;
}
