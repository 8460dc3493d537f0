// Microbenchmark: what store patterns does a B200 SM sustain when 4096 independent warps each stream ~120 KB?
// Isolates the output path of k_env from its rule logic.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3
// -o store_patterns store_patterns.cu ; run on the GPU box.  Prints GB/s per variant (CUDA events, best of 5).
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

constexpr int kRowsPerWarp = 480;          // 480 rows x 240 B = 115 200 B per warp
constexpr int kVecPerWarp = kRowsPerWarp * 15;

__device__ __forceinline__ void st_default(float4* p, float4 v) { *p = v; }
__device__ __forceinline__ void st_cs(float4* p, float4 v) { __stcs(p, v); }

// variant 0/1: 32 lanes x 16 B, warp-contiguous chunk
template <bool CS>
__global__ void __launch_bounds__(128, 8) k_aligned(float4* out, int nwarps) {
    int w = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (w >= nwarps) return;
    float4* dst = out + (size_t)w * kVecPerWarp + lane;
    float4 v = make_float4(1.f, 0.f, 1.f, (float)w);
#pragma unroll 4
    for (int i = 0; i < kVecPerWarp / 32; i++, dst += 32) { if (CS) st_cs(dst, v); else st_default(dst, v); }
}
// variant 2/3: lanes 0..29, two 240-B rows per instruction
template <bool CS>
__global__ void __launch_bounds__(128, 8) k_rows30(float4* out, int nwarps) {
    int w = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (w >= nwarps || lane >= 30) return;
    float4* dst = out + (size_t)w * kVecPerWarp + lane;
    float4 v = make_float4(1.f, 0.f, 1.f, (float)w);
#pragma unroll 4
    for (int i = 0; i < kRowsPerWarp / 2; i++, dst += 30) { if (CS) st_cs(dst, v); else st_default(dst, v); }
}
// variant 4: rows30 + the LDS(row) -> LDS(lut) -> FMUL chain of k_env
template <bool CS>
__global__ void __launch_bounds__(128, 8) k_rows30_lut(float4* out, int nwarps) {
    __shared__ float4 rows[4][288];
    __shared__ float4 lut[8];
    int wib = threadIdx.x >> 5;
    int w = blockIdx.x * 4 + wib, lane = threadIdx.x & 31;
    if (threadIdx.x < 5) lut[threadIdx.x] = make_float4(threadIdx.x > 0, threadIdx.x > 1, threadIdx.x > 2, threadIdx.x > 3);
    for (int i = lane; i < 288; i += 32) rows[wib][i] = make_float4(__uint_as_float(0x43210123u * (i + 1)), __uint_as_float(0x01234321u + i), 1.f, 0.f);
    __syncthreads();
    if (w >= nwarps || lane >= 30) return;
    int first = lane / 15, k = lane % 15; bool hi = k >= 8; unsigned sh = 4 * (k & 7);
    float4* dst = out + (size_t)w * kVecPerWarp + lane;
#pragma unroll 4
    for (int r = first; r < kRowsPerWarp; r += 2, dst += 30) {
        float4 raw = rows[wib][r % 288];
        unsigned wd = hi ? __float_as_uint(raw.y) : __float_as_uint(raw.x);
        float4 q = lut[((wd >> sh) & 15u) % 5];
        float s = raw.z; q.x *= s; q.y *= s; q.z *= s; q.w *= s;
        if (CS) st_cs(dst, q); else st_default(dst, q);
    }
}
// variant 5: grid-strided like a fill kernel (every warp instruction 512 B, warps interleaved)
__global__ void __launch_bounds__(128, 8) k_gridstride(float4* out, size_t nvec) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
    float4 v = make_float4(1.f, 0.f, 1.f, 2.f);
#pragma unroll 4
    for (; i < nvec; i += stride) out[i] = v;
}
// variant 6: TMA bulk store of a 7 680-B tile per warp (single buffer)
__global__ void __launch_bounds__(128, 8) k_tma(float4* out, int nwarps) {
    extern __shared__ __align__(128) float4 tiles[];
    int wib = threadIdx.x >> 5, w = blockIdx.x * 4 + wib, lane = threadIdx.x & 31;
    if (w >= nwarps) return;
    float4* tile = tiles + wib * 480;
    char* dst = (char*)(out + (size_t)w * kVecPerWarp);
    for (int t = 0; t < kRowsPerWarp / 32; t++) {
        if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        __syncwarp();
        for (int k = 0; k < 15; k++) tile[lane * 15 + k] = make_float4(1.f, 0.f, (float)k, (float)t);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) {
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst + (size_t)t * 7680),
                         "r"((unsigned)__cvta_generic_to_shared(tile)), "r"(7680) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}

// variant 7: rows30_cs repeated K times by the same warps into K different buffers inside ONE launch (a K-step launch)
__global__ void __launch_bounds__(128, 7) k_rows30_steps(float4* out, int nwarps, int K) {
    int w = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (w >= nwarps || lane >= 30) return;
    float4 v = make_float4(1.f, 0.f, 1.f, (float)w);
    for (int s = 0; s < K; s++) {
        float4* dst = out + ((size_t)s * nwarps + w) * kVecPerWarp + lane;
#pragma unroll 4
        for (int i = 0; i < kRowsPerWarp / 2; i++, dst += 30) st_cs(dst, v);
    }
}

// variant 8: 256-bit stores (st.global.v8.f32, sm_100+): lanes 0..29 x 32 B = four 240-byte rows per warp instruction
__device__ __forceinline__ void st_v8(float* p, float4 a, float4 b) {
    asm volatile("st.global.cs.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "f"(a.x), "f"(a.y), "f"(a.z),
                 "f"(a.w), "f"(b.x), "f"(b.y), "f"(b.z), "f"(b.w) : "memory");
}
__global__ void __launch_bounds__(128, 8) k_rows30_v8(float4* out, int nwarps) {
    int w = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (w >= nwarps || lane >= 30) return;
    float* dst = (float*)(out + (size_t)w * kVecPerWarp) + lane * 8;
    float4 v = make_float4(1.f, 0.f, 1.f, (float)w);
#pragma unroll 4
    for (int i = 0; i < kRowsPerWarp / 4; i++, dst += 240) st_v8(dst, v, v);
}
__global__ void __launch_bounds__(128, 8) k_gridstride_v8(float4* out, size_t nvec8) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
    float4 v = make_float4(1.f, 0.f, 1.f, 2.f);
    for (; i < nvec8; i += stride) st_v8((float*)out + i * 8, v, v);
}

// variant 9: the warps of a CTA write ONE region at a time together (row pairs interleaved between the warps), region
// after region: the same bytes per CTA as rows30, but 1/WARPS as many concurrently open output streams
template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32) k_rows30_cta(float4* out, int nwarps) {
    const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane >= 30) return;
    float4 v = make_float4(1.f, 0.f, 1.f, (float)wib);
    for (int j = 0; j < WARPS; j++) {
        const int w = blockIdx.x * WARPS + j;
        if (w >= nwarps) return;
        float4* dst = out + (size_t)w * kVecPerWarp + wib * 30 + lane;
#pragma unroll 4
        for (int i = wib; i < kRowsPerWarp / 2; i += WARPS, dst += 30 * WARPS) st_cs(dst, v);
    }
}

// variant 10: "expand" as short-lived CTAs in address order: every thread writes 4 float4 of thermometer rows computed
// from 16-byte row descriptors (packed counts + scale) read through L1/L2 -- the output stage of an env-step as its own
// kernel.  CTA = 256 threads = 16 KB of output, then exit.
__global__ void __launch_bounds__(256) k_expand_oneshot(const float4* __restrict__ desc, float4* __restrict__ out, size_t nvec) {
    const size_t j0 = (size_t)blockIdx.x * 1024 + threadIdx.x;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const size_t j = j0 + 256 * k;
        if (j >= nvec) return;
        const unsigned row = (unsigned)(j / 15), rank = (unsigned)(j - (size_t)row * 15);
        const float4 d = __ldg(&desc[row]);
        const unsigned w = rank >= 8 ? __float_as_uint(d.y) : __float_as_uint(d.x);
        const unsigned c = (w >> (4 * (rank & 7))) & 15u;
        const float s = d.z;
        out[j] = make_float4(c > 0 ? s : 0.f, c > 1 ? s : 0.f, c > 2 ? s : 0.f, c > 3 ? s : 0.f);
    }
}

template <class F>
static float best_ms(F launch) {
    cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    float best = 1e9f;
    for (int i = 0; i < 6; i++) {
        CK(cudaEventRecord(a)); launch(); CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
        float ms; CK(cudaEventElapsedTime(&ms, a, b)); if (i > 0 && ms < best) best = ms;
    }
    CK(cudaGetLastError());
    return best;
}

int main(int argc, char** argv) {
    const int nwarps = argc > 1 ? atoi(argv[1]) : 4096, grid = nwarps / 4;
    const size_t nvec = (size_t)nwarps * kVecPerWarp, bytes = nvec * 16;
    float4* out; CK(cudaMalloc(&out, bytes + 4096));
    CK(cudaFuncSetAttribute(k_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * 7680));
    printf("{\"nwarps\": %d, \"bytes\": %zu", nwarps, bytes);
#define REPORT(name, call) { float ms = best_ms([&] { call; }); printf(", \"%s_GBs\": %.0f", name, bytes / ms / 1e6); }
    REPORT("aligned512", (k_aligned<false><<<grid, 128>>>(out, nwarps)));
    REPORT("aligned512_cs", (k_aligned<true><<<grid, 128>>>(out, nwarps)));
    REPORT("rows30", (k_rows30<false><<<grid, 128>>>(out, nwarps)));
    REPORT("rows30_cs", (k_rows30<true><<<grid, 128>>>(out, nwarps)));
    REPORT("rows30_lut", (k_rows30_lut<false><<<grid, 128>>>(out, nwarps)));
    REPORT("rows30_lut_cs", (k_rows30_lut<true><<<grid, 128>>>(out, nwarps)));
    REPORT("gridstride", (k_gridstride<<<grid, 128>>>(out, nvec)));
    REPORT("gridstride_148x8", (k_gridstride<<<148 * 8, 128>>>(out, nvec)));
    REPORT("tma_tile", (k_tma<<<grid, 128, 4 * 7680>>>(out, nwarps)));
    {
        const size_t nrows = nvec / 15;
        float4* desc; CK(cudaMalloc(&desc, nrows * 16));
        CK(cudaMemset(desc, 0x3f, nrows * 16));
        REPORT("expand_oneshot_16KB_ctas", (k_expand_oneshot<<<(unsigned)((nvec + 1023) / 1024), 256>>>(desc, out, nvec)));
        CK(cudaFree(desc));
    }
    REPORT("rows30_cta4_1024_streams", (k_rows30_cta<4><<<nwarps / 4, 128>>>(out, nwarps)));
    REPORT("rows30_cta8_512_streams", (k_rows30_cta<8><<<nwarps / 8, 256>>>(out, nwarps)));
    REPORT("rows30_cta16_256_streams", (k_rows30_cta<16><<<nwarps / 16, 512>>>(out, nwarps)));
    REPORT("rows30_v8_cs", (k_rows30_v8<<<grid, 128>>>(out, nwarps)));
    REPORT("gridstride_v8", (k_gridstride_v8<<<grid, 128>>>(out, nvec / 2)));
    REPORT("memset", CK(cudaMemsetAsync(out, 1, bytes)));
    {   // K steps: K launches back to back vs one launch that loops (both write K x bytes into K buffers)
        const int K = 8;
        float4* big; CK(cudaMalloc(&big, bytes * K + 4096));
        float ms = best_ms([&] { for (int s = 0; s < K; s++) k_rows30<true><<<grid, 128>>>(big + (size_t)s * nvec, nwarps); });
        printf(", \"rows30_cs_8_launches_GBs\": %.0f", bytes * K / ms / 1e6);
        ms = best_ms([&] { k_rows30_steps<<<grid, 128>>>(big, nwarps, K); });
        printf(", \"rows30_cs_one_launch_8_steps_GBs\": %.0f", bytes * K / ms / 1e6);
        ms = best_ms([&] { k_gridstride<<<grid, 128>>>(big, nvec * K); });
        printf(", \"gridstride_8x_bytes_GBs\": %.0f", bytes * K / ms / 1e6);
        ms = best_ms([&] { k_gridstride_v8<<<grid, 128>>>(big, nvec * K / 2); });
        printf(", \"gridstride_v8_8x_bytes_GBs\": %.0f", bytes * K / ms / 1e6);
        ms = best_ms([&] { CK(cudaMemsetAsync(big, 1, bytes * K)); });
        printf(", \"memset_8x_bytes_GBs\": %.0f", bytes * K / ms / 1e6);
        CK(cudaFree(big));
    }
    printf("}\n");
    return 0;
}
