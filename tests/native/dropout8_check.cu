// Host-side check (compiled by tests/test_host_cpu.py with nvcc): the branch-free dropout8 of csrc/common.cuh takes exactly the
// decisions of the element-wise crct_keep, for 32- and 64-bit element counters and every threshold.
#include "common.cuh"
#include <cstdio>
#include <random>
void crct_set_error(const char*, ...) {}
int main() {
    std::mt19937_64 rng(7);
    long bad = 0, n = 0;
    for (int t = 0; t < 200000; ++t) {
        uint64_t seed = rng();
        uint64_t e0 = (t % 3 == 0 ? (rng() & 0x3FFFFFFFFFull) : (rng() & 0x7FFFFFFFull)) & ~7ull;
        uint32_t thr = t % 5 == 0 ? 0xFFFFu : (uint32_t)(rng() % 65536);
        if (thr == 0) thr = 1;
        float f[8]; for (int j = 0; j < 8; ++j) f[j] = 1.f;
        dropout8(f, seed, e0, thr, 2.f);
        for (int j = 0; j < 8; ++j) { bool k = crct_keep(seed, e0 + j, thr); bad += (k ? 2.f : 0.f) != f[j]; ++n; }
    }
    printf("%ld mismatches of %ld\n", bad, n);
    return bad != 0;
}
