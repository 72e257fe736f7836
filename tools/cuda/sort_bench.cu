// Dev harness (GPU): the library's radix sort on depth-like keys — correctness against std::stable_sort, time per call,
// and (with -DSEGS_RS_PHASE_TIMING) the average duration of every phase of rs_pass_kernel.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -DSEGS_RS_PHASE_TIMING -I include -o build/sort_bench tools/cuda/sort_bench.cu
#include <algorithm>
#include <chrono>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <numeric>
#include <random>
#include <vector>
#ifdef SEGS_RS_PHASE_TIMING
__device__ unsigned long long g_rs_phase[8];
__device__ unsigned long long g_rs_cycles[8];
__device__ unsigned int g_rs_phase_n;
#endif
#include "../../segs_slam_b200/csrc/radix_sort.cu"

namespace segs {
void set_error(const char* fmt, ...) { va_list ap; va_start(ap, fmt); vfprintf(stderr, fmt, ap); va_end(ap); fputc('\n', stderr); }
void count_launch() {}
unsigned readback_event_flags() { return 0; }
cudaEvent_t readback_event() { return nullptr; }
}

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

int main(int argc, char** argv)
{
    const size_t n = argc > 1 ? strtoull(argv[1], nullptr, 10) : 1000000;
    const int passes = argc > 2 ? atoi(argv[2]) : 4;
    const int reps = argc > 3 ? atoi(argv[3]) : 50;
    std::mt19937 rng(5);
    std::uniform_real_distribution<float> depth(0.5f, 6.0f);
    std::vector<uint32_t> keys(n);
    for (size_t i = 0; i < n; ++i) {
        float d = depth(rng);
        uint32_t k; memcpy(&k, &d, 4);
        keys[i] = (rng() % 9 == 0) ? 0xFFFFFFFFu : (passes < 4 ? (k >> (32 - 8 * passes)) : k);
    }
    std::vector<uint32_t> order(n);
    std::iota(order.begin(), order.end(), 0u);
    std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return keys[a] < keys[b]; });
    uint32_t *ka, *kb, *va, *vb, *tmp;
    const size_t tw = segs::radix_sort_temp_words(n, passes);
    CK(cudaMalloc(&ka, n * 4)); CK(cudaMalloc(&kb, n * 4)); CK(cudaMalloc(&va, n * 4)); CK(cudaMalloc(&vb, n * 4)); CK(cudaMalloc(&tmp, tw * 4));
    cudaStream_t st; CK(cudaStreamCreate(&st));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    std::vector<float> ms(reps);
    for (int r = 0; r < reps + 5; ++r) {
        CK(cudaMemcpyAsync(ka, keys.data(), n * 4, cudaMemcpyHostToDevice, st));
#ifdef SEGS_RS_PHASE_TIMING
        if (r == 5) { unsigned long long z[8] = {0}; CK(cudaMemcpyToSymbolAsync(g_rs_cycles, z, sizeof(z), 0, cudaMemcpyHostToDevice, st)); unsigned zn = 0; CK(cudaMemcpyToSymbolAsync(g_rs_phase, z, sizeof(z), 0, cudaMemcpyHostToDevice, st)); CK(cudaMemcpyToSymbolAsync(g_rs_phase_n, &zn, 4, 0, cudaMemcpyHostToDevice, st)); }
#endif
        CK(cudaEventRecord(e0, st));
        if (segs::radix_sort_pairs(ka, kb, va, vb, n, 0, passes, true, tmp, st)) return 1;
        CK(cudaEventRecord(e1, st));
        CK(cudaStreamSynchronize(st));
        if (r >= 5) CK(cudaEventElapsedTime(&ms[r - 5], e0, e1));
    }
    std::vector<uint32_t> out(n);
    CK(cudaMemcpy(out.data(), (passes & 1) ? vb : va, n * 4, cudaMemcpyDeviceToHost));
    size_t bad = 0;
    for (size_t i = 0; i < n; ++i) bad += out[i] != order[i];
    std::sort(ms.begin(), ms.end());
    printf("n=%zu passes=%d mismatches=%zu  time us: min %.1f median %.1f max %.1f\n", n, passes, bad, ms[0] * 1e3, ms[reps / 2] * 1e3, ms[reps - 1] * 1e3);
#ifdef SEGS_RS_PHASE_TIMING
    unsigned long long ph[8]; unsigned pn = 0;
    CK(cudaMemcpyFromSymbol(ph, g_rs_phase, sizeof(ph))); CK(cudaMemcpyFromSymbol(&pn, g_rs_phase_n, 4));
    (void)pn;
    unsigned long long cy[8];
    CK(cudaMemcpyFromSymbol(cy, g_rs_cycles, sizeof(cy)));
    if (cy[7]) { printf("sort_tile phases, avg cycles per tile (init+load | early counts+publish | ranking | look-back | scans | reorder | store):"); for (int i = 0; i < 7; ++i) printf(" %.0f", double(cy[i]) / cy[7]); printf("\n"); }
    printf("stage end times (us after the first ticket):"); for (int i = 1; i < 8 && ph[i]; ++i) printf(" %.1f", double(ph[i] - ph[0]) * 1e-3); printf("\n");
#endif
    return bad != 0;
}
