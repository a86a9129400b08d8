// Micro-benchmark: can 8 warps per SM stream FP32 [rows][128] arrays in the "thread = row" mapping with
// 256-bit global loads / stores (32 lanes -> 32 different 512-byte rows, one full 32-byte sector each)?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o rowstream rowstream.cu && ./rowstream
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ void ld256(const float* p, float* v) {
    asm volatile("ld.global.nc.L1::no_allocate.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]) : "l"(p));
}
__device__ __forceinline__ void st256(float* p, const float* v) {
    asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                 ::"l"(p), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7]) : "memory");
}

// MODE 0: read only (sum), 1: write only, 2: copy.  DEPTH loads of 32 bytes in flight per thread.
template <int MODE, int DEPTH>
__global__ void __launch_bounds__(256, 1) rowstream(const float* __restrict__ in, float* __restrict__ out, int64_t rows, float* sink) {
    const int64_t tiles = rows / 256;
    float acc = 0.f;
    for (int64_t t = blockIdx.x; t < tiles; t += gridDim.x) {
        const int64_t row = t * 256 + threadIdx.x;
        const float* src = in + row * 128;
        float* dst = out + row * 128;
#pragma unroll 1
        for (int c = 0; c < 128; c += 8 * DEPTH) {
            float v[DEPTH][8];
            if (MODE != 1) {
#pragma unroll
                for (int d = 0; d < DEPTH; ++d) ld256(src + c + 8 * d, v[d]);
            } else {
#pragma unroll
                for (int d = 0; d < DEPTH; ++d)
#pragma unroll
                    for (int j = 0; j < 8; ++j) v[d][j] = (float)(c + j);
            }
            if (MODE == 0) {
#pragma unroll
                for (int d = 0; d < DEPTH; ++d)
#pragma unroll
                    for (int j = 0; j < 8; ++j) acc += v[d][j];
            } else {
#pragma unroll
                for (int d = 0; d < DEPTH; ++d) st256(dst + c + 8 * d, v[d]);
            }
        }
    }
    if (MODE == 0 && acc == 123.456f) *sink = acc;
}

// coalesced reference: float4 per thread, consecutive threads consecutive addresses
template <int MODE>
__global__ void __launch_bounds__(256, 1) coalesced(const float4* __restrict__ in, float4* __restrict__ out, int64_t n4, float* sink) {
    float acc = 0.f;
    for (int64_t i = blockIdx.x * 256 * 4 + threadIdx.x; i < n4; i += (int64_t)gridDim.x * 256 * 4) {
        float4 v[4];
#pragma unroll
        for (int d = 0; d < 4; ++d) v[d] = MODE != 1 ? __ldg(in + i + d * 256) : make_float4(1, 2, 3, 4);
        if (MODE == 0) {
#pragma unroll
            for (int d = 0; d < 4; ++d) acc += v[d].x + v[d].y + v[d].z + v[d].w;
        } else {
#pragma unroll
            for (int d = 0; d < 4; ++d) out[i + d * 256] = v[d];
        }
    }
    if (MODE == 0 && acc == 123.456f) *sink = acc;
}

template <typename F>
float time_ms(F f) {
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    f();
    cudaEventRecord(a);
    for (int i = 0; i < 5; ++i) f();
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    return ms / 5;
}

int main() {
    const int64_t rows = 1 << 21;     // 1 GiB per array
    float *in, *out, *sink;
    cudaMalloc(&in, rows * 512); cudaMalloc(&out, rows * 512); cudaMalloc(&sink, 4);
    cudaMemset(in, 0, rows * 512);
    const double gb = rows * 512 / 1e9;
    const int g = 148;
#define RUN(name, expr, bytes) { float ms = time_ms([&] { expr; }); printf("%-34s %8.3f ms  %8.1f GB/s\n", name, ms, (bytes) / ms * 1e3); }
    RUN("row256 read  depth2", (rowstream<0, 2><<<g, 256>>>(in, out, rows, sink)), gb);
    RUN("row256 read  depth4", (rowstream<0, 4><<<g, 256>>>(in, out, rows, sink)), gb);
    RUN("row256 read  depth8", (rowstream<0, 8><<<g, 256>>>(in, out, rows, sink)), gb);
    RUN("row256 write depth2", (rowstream<1, 2><<<g, 256>>>(in, out, rows, sink)), gb);
    RUN("row256 write depth4", (rowstream<1, 4><<<g, 256>>>(in, out, rows, sink)), gb);
    RUN("row256 copy  depth2", (rowstream<2, 2><<<g, 256>>>(in, out, rows, sink)), 2 * gb);
    RUN("row256 copy  depth4", (rowstream<2, 4><<<g, 256>>>(in, out, rows, sink)), 2 * gb);
    RUN("row256 copy  depth8", (rowstream<2, 8><<<g, 256>>>(in, out, rows, sink)), 2 * gb);
    RUN("coalesced read  (256 thr/SM)", (coalesced<0><<<g, 256>>>((const float4*)in, (float4*)out, rows * 32, sink)), gb);
    RUN("coalesced write (256 thr/SM)", (coalesced<1><<<g, 256>>>((const float4*)in, (float4*)out, rows * 32, sink)), gb);
    RUN("coalesced copy  (256 thr/SM)", (coalesced<2><<<g, 256>>>((const float4*)in, (float4*)out, rows * 32, sink)), 2 * gb);
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
