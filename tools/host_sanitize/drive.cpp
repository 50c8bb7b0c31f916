#include <cstdio>
#include <cstdint>
#include <random>
#include <vector>
#include <poppy_host.h>
int main() {
    const int W = 1920, H = 1080, N = 3000, F = 48;
    std::mt19937 rng(3);
    std::uniform_real_distribution<float> ux(2, W - 3), uy(2, H - 3), uj(-4, 4);
    std::vector<float> a(2 * N), b(2 * N), s(F);
    for (int i = 0; i < N; ++i) { a[2*i] = ux(rng); a[2*i+1] = uy(rng); b[2*i] = a[2*i] + uj(rng); b[2*i+1] = a[2*i+1] + uj(rng); }
    for (int f = 0; f < F; ++f) s[f] = (float)f / (F - 1);
    long long total = 0;
    for (int rep = 0; rep < 3; ++rep)
        for (int threads : {8, 3, 1}) {
            poppy_host_plan* plan = nullptr;
            if (poppy_host_plan_create(&plan, a.data(), b.data(), N, W, H, F, s.data(), 0, threads) != 0) { printf("failed: %s\n", poppy_host_last_error()); return 1; }
            const int32_t *tri, *off; int mx;
            poppy_host_plan_triangles(plan, &tri, &off, &mx);
            total += off[F];
            poppy_host_plan_destroy(plan);
        }
    printf("ok, %lld triangles\n", total);
}
