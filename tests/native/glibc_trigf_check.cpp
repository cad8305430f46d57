// Checks csrc/glibc_trigf.cuh (the engine's restatement of glibc 2.39 sinf/cosf, host instantiation) against the libm this
// process links: every float bit pattern b with b % stride == phase (stride 1 = all 2^32), for both builds (FMA / SSE2).
// Prints one line per build: how many results differ from libm in any bit. Exactly one build must show 0 / 0: the one
// this machine's libm selected at load time. Exit code 0 if so.
//   g++ -O2 -std=c++17 -ffp-contract=off -mfma -pthread -o check glibc_trigf_check.cpp -lm ; ./check [stride] [threads]
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

#include "../../montecarlolocalisation_b200/csrc/glibc_trigf.cuh"

static inline uint32_t bits_of(float f) { uint32_t b; memcpy(&b, &f, 4); return b; }

int main(int argc, char** argv) {
    const uint64_t stride = argc > 1 ? strtoull(argv[1], nullptr, 10) : 61;
    const int threads = argc > 2 ? atoi(argv[2]) : (int)std::max(1u, std::thread::hardware_concurrency());
    std::atomic<uint64_t> bad[2][2];
    for (auto& a : bad) for (auto& b : a) b = 0;
    std::atomic<uint64_t> first_bad[2];
    first_bad[0] = first_bad[1] = ~0ull;
    std::vector<std::thread> pool;
    const uint64_t total = (1ull << 32);
    for (int t = 0; t < threads; t++) {
        pool.emplace_back([&, t] {
            uint64_t local[2][2] = {{0, 0}, {0, 0}};
            const uint64_t lo = total * t / threads, hi = total * (t + 1) / threads;
            for (uint64_t b = lo + (stride - lo % stride) % stride; b < hi; b += stride) {
                float y; const uint32_t bb = (uint32_t)b; memcpy(&y, &bb, 4);
                volatile float vy = y;                       // keep the compiler from folding the libm calls
                const uint32_t s = bits_of(sinf(vy)), c = bits_of(cosf(vy));
                const uint32_t s1 = bits_of(mcl::glibc_trig::sinf_as_glibc<true>(y)), c1 = bits_of(mcl::glibc_trig::cosf_as_glibc<true>(y));
                const uint32_t s0 = bits_of(mcl::glibc_trig::sinf_as_glibc<false>(y)), c0 = bits_of(mcl::glibc_trig::cosf_as_glibc<false>(y));
                if (s1 != s) { local[1][0]++; uint64_t e = first_bad[1]; while (b < e && !first_bad[1].compare_exchange_weak(e, b)) {} }
                if (c1 != c) local[1][1]++;
                if (s0 != s) { local[0][0]++; uint64_t e = first_bad[0]; while (b < e && !first_bad[0].compare_exchange_weak(e, b)) {} }
                if (c0 != c) local[0][1]++;
            }
            for (int v = 0; v < 2; v++) for (int f = 0; f < 2; f++) bad[v][f] += local[v][f];
        });
    }
    for (auto& th : pool) th.join();
    const uint64_t count = (total + stride - 1) / stride;
    int matching = -1;
    for (int v = 1; v >= 0; v--) {
        printf("build %s: %llu arguments, sinf mismatches %llu, cosf mismatches %llu", v ? "fma " : "sse2", (unsigned long long)count,
               (unsigned long long)bad[v][0].load(), (unsigned long long)bad[v][1].load());
        if (first_bad[v].load() != ~0ull) printf(" (first sinf mismatch at bits 0x%08llx)", (unsigned long long)first_bad[v].load());
        printf("\n");
        if (bad[v][0] == 0 && bad[v][1] == 0) matching = v;
    }
    if (matching < 0) { printf("neither build reproduces this libm\n"); return 1; }
    printf("host libm = %s build, reproduced bit for bit\n", matching ? "fma" : "sse2");
    return 0;
}
