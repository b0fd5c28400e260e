// pipes.cu -- instruction-throughput probes for the non-tensor pipes this backend
// lives on (FP32 FMA, FP64 FMA/ADD, F2F f32->f64, MUFU rsqrt, double sqrt, LDS.128
// broadcast, shared atomics).  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipes pipes.cu
// Prints lane-ops/s per SM-clock so the DESIGN.md roofline figures are measured, not assumed.
#include <cstdio>
#include <cuda_runtime.h>

constexpr int ITERS = 4096;
constexpr int ILP = 8;

template <int OP>
__global__ void __launch_bounds__(256) probe(double *out, float seed)
{
    double d[ILP];
    float f[ILP];
    __shared__ float4 tile[256];
    __shared__ unsigned hist[256];
    tile[threadIdx.x] = make_float4(seed, seed, seed, seed);
    hist[threadIdx.x] = 0;
    __syncthreads();
#pragma unroll
    for (int k = 0; k < ILP; ++k) { d[k] = seed + k + threadIdx.x; f[k] = seed + k + threadIdx.x; }
    const double a = 1.0000001, b = 1e-9;
    const float af = 1.0000001f, bf = 1e-9f;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int k = 0; k < ILP; ++k) {
            if (OP == 0) f[k] = fmaf(f[k], af, bf);                       // FFMA
            if (OP == 1) d[k] = fma(d[k], a, b);                          // DFMA
            if (OP == 2) d[k] = d[k] + b;                                 // DADD
            if (OP == 3) d[k] = d[k] * a;                                 // DMUL
            if (OP == 4) { d[k] = d[k] + (double)f[k]; f[k] += bf; }      // F2F.F64.F32 + DADD + FADD
            if (OP == 5) f[k] = rsqrtf(f[k]) + af;                        // MUFU.RSQ + FADD
            if (OP == 6) d[k] = sqrt(d[k]) + a;                           // IEEE double sqrt + DADD
            if (OP == 7) { float4 t = tile[(it + k) & 255]; f[k] += t.x; }   // LDS.128 broadcast + FADD
            if (OP == 8) atomicAdd(&hist[(threadIdx.x * 7 + it + k) & 255], 1u);   // ATOMS spread
            if (OP == 9) f[k] = (float)(int)f[k] + bf;                    // F2I + I2F + FADD
            if (OP == 10) d[k] = round(d[k]) + b;                         // double round (half away)
            if (OP == 11 && (k & 1) == 0) {                               // FFMA2 (two FMAs per lane per instruction)
                unsigned long long x, y = ((unsigned long long)__float_as_uint(af) << 32) | __float_as_uint(af),
                                   z = ((unsigned long long)__float_as_uint(bf) << 32) | __float_as_uint(bf);
                x = ((unsigned long long)__float_as_uint(f[k + 1]) << 32) | __float_as_uint(f[k]);
                asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(x) : "l"(x), "l"(y), "l"(z));
                f[k] = __uint_as_float((unsigned)x); f[k + 1] = __uint_as_float((unsigned)(x >> 32));
            }
            if (OP == 12) { float r; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(f[k])); f[k] = r; }   // MUFU.SQRT chain
            if (OP == 13) { float2 t = reinterpret_cast<float2 *>(tile)[(it + k) & 511]; f[k] += t.x; }     // LDS.64 broadcast + FADD
        }
    }
    double s = 0;
#pragma unroll
    for (int k = 0; k < ILP; ++k) s += d[k] + f[k];
    if (s == 12345.678) out[0] = s + hist[3];
}

template <int OP>
void run(const char *name, double *out, int sms, double clock_ghz, double ops_per_iter)
{
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int blocks = sms * 8;
    probe<OP><<<blocks, 256>>>(out, 1.5f);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    probe<OP><<<blocks, 256>>>(out, 1.5f);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double total = (double)blocks * 256 * ITERS * ILP * ops_per_iter;
    printf("%-34s %8.3f ms  %9.2f Gop/s  %6.2f lane-ops/clk/SM (at %.3f GHz)\n", name, ms, total / ms / 1e6,
           total / (ms * 1e-3) / sms / (clock_ghz * 1e9), clock_ghz);
}

int main()
{
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    int clk_khz = 0;
    cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    const double ghz = clk_khz / 1e6;
    printf("%s  SMs=%d  clock=%.3f GHz\n", p.name, p.multiProcessorCount, ghz);
    double *out;
    cudaMalloc(&out, 8);
    const int sms = p.multiProcessorCount;
    run<0>("FFMA", out, sms, ghz, 1);
    run<1>("DFMA", out, sms, ghz, 1);
    run<2>("DADD", out, sms, ghz, 1);
    run<3>("DMUL", out, sms, ghz, 1);
    run<4>("F2F.F64.F32 + DADD + FADD", out, sms, ghz, 1);
    run<5>("MUFU.RSQ + FADD", out, sms, ghz, 1);
    run<6>("sqrt(double) + DADD", out, sms, ghz, 1);
    run<7>("LDS.128 broadcast + FADD", out, sms, ghz, 1);
    run<8>("ATOMS.ADD spread", out, sms, ghz, 1);
    run<9>("F2I + I2F + FADD", out, sms, ghz, 1);
    run<10>("round(double) + DADD", out, sms, ghz, 1);
    run<11>("FFMA2 (2 FMA per lane-instr)", out, sms, ghz, 1);
    run<12>("MUFU.SQRT", out, sms, ghz, 1);
    run<13>("LDS.64 broadcast + FADD", out, sms, ghz, 1);
    return 0;
}
