// scenegen.cpp - host-side synthetic input generator.
// Not on the rasterization path: it only produces the triangle soup both the CUDA backend and
// the CPU oracle are fed with.  std::mt19937_64 + std::uniform_real_distribution<double>(0,1)
// of libstdc++ is the generator the survey's known-answer counts were obtained with.
#include <cstdint>
#include <random>

extern "C" __attribute__((visibility("default")))
int trb_gen_soup_clip(uint64_t seed, uint64_t ntris, int width, int height, double r, int round_fp32,
                      double* clip12, float* pos9) {
    if ((!clip12 && !pos9) || width <= 0 || height <= 0) return -1;   // either output may be NULL
    std::mt19937_64 rng(seed);
    std::uniform_real_distribution<double> U(0.0, 1.0);
    const double W = width, H = height;
    for (uint64_t t = 0; t < ntris; ++t) {
        double cx = U(rng) * W;
        double cy = U(rng) * H;
        double z = 2.0 * U(rng) - 1.0;
        double px[3] = {cx - r, cx + r, cx};
        double py[3] = {cy - r, cy - r, cy + r};
        for (int v = 0; v < 3; ++v) {
            double x = px[v] / (W / 2.0) - 1.0;
            double y = py[v] / (H / 2.0) - 1.0;
            double zz = z;
            if (round_fp32) {
                x = (double)(float)x;
                y = (double)(float)y;
                zz = (double)(float)zz;
            }
            if (clip12) {
                double* c = clip12 + t * 12 + v * 4;
                c[0] = x; c[1] = y; c[2] = zz; c[3] = 1.0;
            }
            if (pos9) {
                float* p = pos9 + t * 9 + v * 3;
                p[0] = (float)x; p[1] = (float)y; p[2] = (float)zz;
            }
        }
    }
    return 0;
}
