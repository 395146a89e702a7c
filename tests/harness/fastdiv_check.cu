// tests/harness/fastdiv_check.cu - TEST ONLY.  Brute-force check on the GPU that exact.cuh's
// shared-divisor division div_rn(a, make_rcp(b)) equals the IEEE division a / b bit for bit.
// usage: fastdiv_check [millions of random pairs per class, default 2000]
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../tinyrenderder_b200/csrc/exact.cuh"
using namespace trbx;

__device__ __forceinline__ uint64_t mix(uint64_t x) {  // splitmix64
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
__device__ __forceinline__ double make_double(uint64_t mant, int exp, bool neg) {
    uint64_t b = (mant & 0x000FFFFFFFFFFFFFull) | ((uint64_t)(exp + 1023) << 52) | (neg ? 0x8000000000000000ull : 0);
    return bits_f64(b);
}
__device__ __forceinline__ void check(double a, double b, unsigned long long* bad, double* ex) {
    double want = a / b;
    double got = div_rn(a, make_rcp(b));
    if (f64_bits(want) != f64_bits(got) && !(want != want && got != got)) {
        unsigned long long k = atomicAdd(bad, 1ull);
        if (k < 4) { ex[3 * k] = a; ex[3 * k + 1] = b; ex[3 * k + 2] = got; }
    }
}
__global__ void k_check(uint64_t seed, uint64_t per_thread, int cls, unsigned long long* bad, double* ex) {
    uint64_t s = seed + (uint64_t)(blockIdx.x * blockDim.x + threadIdx.x) * 0x100000001B3ull;
    for (uint64_t i = 0; i < per_thread; ++i) {
        uint64_t r1 = mix(s), r2 = mix(s + 1), r3 = mix(s + 2);
        s += 3;
        double a, b;
        if (cls == 0) {          // random mantissas, exponents anywhere (exercises the fallback too)
            a = make_double(r1, (int)(r3 % 2040) - 1020, r3 >> 63);
            b = make_double(r2, (int)((r3 >> 16) % 2040) - 1020, (r3 >> 62) & 1);
        } else if (cls == 1) {   // random mantissas, moderate exponents (the window the kernels live in)
            a = make_double(r1, (int)(r3 % 120) - 60, r3 >> 63);
            b = make_double(r2, (int)((r3 >> 16) % 120) - 60, (r3 >> 62) & 1);
        } else if (cls == 2) {   // exact and almost exact quotients: a = q*b rounded, +- a few ulps
            double q = make_double(r1 & 0x000FFFFFFFF00000ull, (int)(r3 % 40) - 20, false);
            b = make_double(r2 & 0x000FFFFF00000000ull, (int)((r3 >> 16) % 40) - 20, (r3 >> 62) & 1);
            a = bits_f64(f64_bits(q * b) + (int64_t)((r3 >> 32) % 5) - 2);
        } else if (cls == 3) {   // divisors with extreme mantissas (all ones, 1.0, 1+ulp, ...)
            uint64_t m = (r3 & 1) ? 0x000FFFFFFFFFFFFFull - ((r3 >> 8) % 4) : ((r3 >> 8) % 4);
            b = make_double(m, (int)((r3 >> 16) % 80) - 40, (r3 >> 62) & 1);
            a = make_double((r3 & 2) ? r1 : (0x000FFFFFFFFFFFFFull - (r1 % 8)), (int)(r3 % 80) - 40, r3 >> 63);
        } else {                 // quotients next to a rounding midpoint: a = (q + half ulp) * b, perturbed
            double q = make_double(r1, 0, false);
            b = make_double(r2, (int)((r3 >> 16) % 20) - 10, (r3 >> 62) & 1);
            double qh = bits_f64(f64_bits(q)) ;
            double mid = qh * b;                       // rounded product; low bits perturbed below
            a = bits_f64(f64_bits(mid) + (int64_t)((r3 >> 32) % 9) - 4);
            if ((r3 >> 40) & 1) a = fma(0.5 * (bits_f64(f64_bits(q) + 1) - q), b, mid);  // ~ (q + ulp/2) * b
        }
        check(a, b, bad, ex);
        if (cls == 1) {          // numerators that are special for the guard: zeros, tiny, huge
            check(0.0, b, bad, ex);
            check(-0.0, b, bad, ex);
        }
    }
}
int main(int argc, char** argv) {
    uint64_t millions = argc > 1 ? strtoull(argv[1], 0, 10) : 2000;
    unsigned long long* bad;
    double* ex;
    cudaMallocManaged(&bad, 8);
    cudaMallocManaged(&ex, 12 * 8);
    unsigned long long total_bad = 0;
    const int blocks = 148 * 8, threads = 256;
    uint64_t per_thread = millions * 1000000ull / ((uint64_t)blocks * threads) + 1;
    for (int cls = 0; cls < 5; ++cls) {
        *bad = 0;
        k_check<<<blocks, threads>>>(0x1234567ull * (cls + 1), per_thread, cls, bad, ex);
        if (cudaDeviceSynchronize() != cudaSuccess) { printf("cuda error\n"); return 2; }
        printf("class %d: %llu pairs, %llu mismatches\n", cls, (unsigned long long)(per_thread * blocks * threads), *bad);
        for (unsigned long long k = 0; k < *bad && k < 4; ++k)
            printf("   a=%a b=%a got=%a want=%a\n", ex[3 * k], ex[3 * k + 1], ex[3 * k + 2], ex[3 * k] / ex[3 * k + 1]);
        total_bad += *bad;
    }
    printf("fastdiv_check: %s\n", total_bad ? "FAILED" : "ok");
    return total_bad ? 1 : 0;
}
