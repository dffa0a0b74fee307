// Thread-safety driver of the block layer (built with -fsanitize=thread by `make -C gr_clenabled_b200/host tsan`):
// what a flowgraph does concurrently -- a setter called from another thread than work() (clFilter::set_taps2,
// clMathConst::set_k; reference: d_setlock / d_mutex) and the X-engine's push / pickup path.
#include <clenabled/blocks.h>

#include <atomic>
#include <cmath>
#include <cstdio>
#include <thread>

using namespace gr::clenabled;

int main()
{
    const int GPU = OCLTYPE_GPU, FIRST = OCLDEVICESELECTOR_FIRST;
    int failures = 0;
    {   // clFilter: work() on one thread, set_taps2() on another; every call must see one consistent tap set
        const int K = 64, L = 2048;
        std::vector<float> ta(K, 1.0f / K), tb(K, 2.0f / K);
        auto blk = clFilter::make(GPU, FIRST, 0, 0, 1, ta);
        std::atomic<bool> stop{false};
        std::thread setter([&] {
            for (int i = 0; !stop.load(); i++) blk->set_taps2(i % 2 ? ta : tb);
        });
        std::vector<gr_complex> in(L, gr_complex(1.0f, 0.0f)), out(L);
        gr_vector_const_void_star iv{in.data()};
        gr_vector_void_star ov{out.data()};
        for (int it = 0; it < 200; it++) {
            if (blk->work(L, iv, ov) != L) failures++;
            const float g = out[L - 1].real();                  // DC gain of the active tap set: 1 or 2
            if (!(std::fabs(g - 1.0f) < 1e-4f || std::fabs(g - 2.0f) < 1e-4f)) failures++;
            for (int i = K; i < L; i++)
                if (std::fabs(out[i].real() - g) > 1e-4f) {
                    failures++;
                    break;
                }
        }
        stop = true;
        setter.join();
    }
    {   // clMathConst: set_k from another thread
        auto blk = clMathConst::make(DTYPE_COMPLEX, GPU, FIRST, 0, 0, 2.0f, MATHOP_MULTIPLY);
        std::atomic<bool> stop{false};
        std::thread setter([&] {
            for (int i = 0; !stop.load(); i++) blk->set_k(i % 2 ? 2.0f : 3.0f);
        });
        std::vector<gr_complex> in(8192, gr_complex(1.0f, 1.0f)), out(8192);
        gr_vector_const_void_star iv{in.data()};
        gr_vector_void_star ov{out.data()};
        for (int it = 0; it < 200; it++) {
            blk->work(8192, iv, ov);
            const float g = out[0].real();
            for (auto &v : out)
                if (v.real() != g) {
                    failures++;
                    break;
                }
        }
        stop = true;
        setter.join();
    }
    {   // clXEngine: the scheduler thread pushes and picks up; stop() drains from the main thread afterwards
        const int A = 4, F = 16, T = 32;
        auto blk = clXEngine::make(GPU, FIRST, 0, 0, false, DTYPE_BYTE, 1, A, 1, 0, F, T, {});
        std::vector<std::vector<signed char>> ports(A, std::vector<signed char>((size_t)T * F * 2, 3));
        std::thread sched([&] {
            gr_vector_const_void_star iv;
            for (auto &p : ports) iv.push_back(p.data());
            gr_vector_void_star ov;
            gr_vector_int ni(A, T);
            for (int it = 0; it < 50; it++) blk->general_work(T, ni, iv, ov);
        });
        sched.join();
        blk->stop();
        if (blk->published("xcorr").size() != 50) failures++;
    }
    printf("%s (%d failures)\n", failures ? "THREAD TEST FAILED" : "thread test ok", failures);
    return failures ? 1 : 0;
}
