// FP64 peak micro-benchmark for B200 (sm_100a): DFMA pipe vs DMMA (mma.sync m8n8k4 / m16n8k8 f64).
// Writes one JSON line; bench.py reads profiles/fp64_peak_r01.json as the FP64 roofline denominator.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/fp64_peak tools/fp64_peak.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

template <int ILP>
__global__ void __launch_bounds__(256) dfma_kernel(double* out, int iters, double a, double b) {
    double acc[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) acc[i] = threadIdx.x * 1e-9 + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) acc[i] = fma(acc[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += acc[i];
    if (s == 123.456) out[0] = s;
}

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int ILP>
__global__ void __launch_bounds__(256) dmma884_kernel(double* out, int iters, double a, double b) {
    double c0[ILP], c1[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) { c0[i] = i; c1[i] = -i; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) dmma884(c0[i], c1[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += c0[i] + c1[i];
    if (s == 123.456) out[0] = s;
}

__device__ __forceinline__ void dmma1688(double (&c)[4], const double (&a)[4], const double (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
                 : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}

template <int ILP>
__global__ void __launch_bounds__(256) dmma1688_kernel(double* out, int iters, double a, double b) {
    double c[ILP][4];
    double af[4] = {a, a + 1, a + 2, a + 3};
    double bf[2] = {b, b + 1};
#pragma unroll
    for (int i = 0; i < ILP; ++i) { c[i][0] = i; c[i][1] = -i; c[i][2] = 1; c[i][3] = 2; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) dmma1688(c[i], af, bf);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    if (s == 123.456) out[0] = s;
}

// mixed: DFMA and DMMA issued together (are the pipes shared?)
template <int ILP>
__global__ void __launch_bounds__(256) mixed_kernel(double* out, int iters, double a, double b) {
    double c0[ILP], c1[ILP], acc[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) { c0[i] = i; c1[i] = -i; acc[i] = i * 0.5; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) { dmma884(c0[i], c1[i], a, b); acc[i] = fma(acc[i], a, b); acc[i] = fma(acc[i], a, b);
            acc[i] = fma(acc[i], a, b); acc[i] = fma(acc[i], a, b); acc[i] = fma(acc[i], a, b); acc[i] = fma(acc[i], a, b);
            acc[i] = fma(acc[i], a, b); acc[i] = fma(acc[i], a, b); }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += c0[i] + c1[i] + acc[i];
    if (s == 123.456) out[0] = s;
}

template <typename F>
static double time_ms(F launch, int reps) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    launch(); launch(); launch();
    CK(cudaDeviceSynchronize());
    double best = 1e30;
    for (int r = 0; r < reps; ++r) {
        CK(cudaEventRecord(e0));
        launch();
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    return best;
}

int main() {
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    int sms = prop.multiProcessorCount;
    double* out; CK(cudaMalloc(&out, 8));
    const int iters = 4096;
    const int ctas = sms * 8;  // 8 CTAs x 256 thr = 64 warps / SM
    double res[8];
    const char* names[8] = {"dfma_ilp8", "dfma_ilp16", "dmma884_ilp8", "dmma884_ilp16", "dmma1688_ilp4", "dmma1688_ilp8", "mixed_ilp4", "dfma_long"};
    {
        double ms = time_ms([&] { dfma_kernel<8><<<ctas, 256>>>(out, iters, 1.0000001, 1e-9); }, 5);
        res[0] = 2.0 * 8 * iters * 256.0 * ctas / (ms * 1e-3) / 1e12;
    }
    {
        double ms = time_ms([&] { dfma_kernel<16><<<ctas, 256>>>(out, iters, 1.0000001, 1e-9); }, 5);
        res[1] = 2.0 * 16 * iters * 256.0 * ctas / (ms * 1e-3) / 1e12;
    }
    {
        double ms = time_ms([&] { dmma884_kernel<8><<<ctas, 256>>>(out, iters, 1.0000001, 1e-9); }, 5);
        res[2] = 2.0 * 256 * 8 * iters * 8.0 * ctas / (ms * 1e-3) / 1e12;  // 256 FMA per warp-instr, 8 warps
    }
    {
        double ms = time_ms([&] { dmma884_kernel<16><<<ctas, 256>>>(out, iters, 1.0000001, 1e-9); }, 5);
        res[3] = 2.0 * 256 * 16 * iters * 8.0 * ctas / (ms * 1e-3) / 1e12;
    }
    {
        double ms = time_ms([&] { dmma1688_kernel<4><<<ctas, 256>>>(out, iters, 1.0000001, 1e-9); }, 5);
        res[4] = 2.0 * 1024 * 4 * iters * 8.0 * ctas / (ms * 1e-3) / 1e12;
    }
    {
        double ms = time_ms([&] { dmma1688_kernel<8><<<ctas, 256>>>(out, iters, 1.0000001, 1e-9); }, 5);
        res[5] = 2.0 * 1024 * 8 * iters * 8.0 * ctas / (ms * 1e-3) / 1e12;
    }
    {
        double ms = time_ms([&] { mixed_kernel<4><<<ctas, 256>>>(out, iters, 1.0000001, 1e-9); }, 5);
        res[6] = 2.0 * (256 + 8 * 32) * 4 * iters * 8.0 * ctas / (ms * 1e-3) / 1e12;
    }
    {   // sustained: ~2 s of DFMA
        double ms = time_ms([&] { dfma_kernel<16><<<ctas * 64, 256>>>(out, iters, 1.0000001, 1e-9); }, 3);
        res[7] = 2.0 * 16 * iters * 256.0 * ctas * 64 / (ms * 1e-3) / 1e12;
    }
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"clock_khz\": %d", prop.name, sms, prop.clockRate);
    for (int i = 0; i < 8; ++i) printf(", \"%s_tflops\": %.3f", names[i], res[i]);
    printf("}\n");
    return 0;
}
