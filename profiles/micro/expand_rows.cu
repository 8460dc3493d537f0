// Microbenchmark for a split env-step: could the thermometer rows be written by their OWN kernel (a pure row writer that
// reads 16-byte row descriptors -- packed counts + scale -- left in global memory by a rule kernel) faster than by the
// long-lived warps of k_env?  Variants: rows per warp 480 (k_env's shape: 4096 warps, one wave), 120, 60, 30 (short-lived
// warps in many waves), values from a PRNG (no loads, the upper bound) or from descriptors; plain and compressible memory.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o expand_rows expand_rows.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)
#define CU(x) do { CUresult e = (x); if (e != CUDA_SUCCESS) { const char* s; cuGetErrorString(e, &s); printf("driver error %s at %d\n", s, __LINE__); exit(1); } } while (0)

struct __align__(16) Desc { unsigned lo, hi; float s, pad; };

__global__ void k_fill_desc(Desc* d, size_t nrows, unsigned seed) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nrows) return;
    unsigned x = seed ^ (unsigned)(i * 2654435761u), lo = 0, hi = 0;
    for (int k = 0; k < 15; k++) {
        x = x * 1664525u + 1013904223u;
        const unsigned c = (x >> 24) % 9 < 5 ? 0u : ((x >> 16) & 3u) + 1u;      // mostly 0, else 1..4 (like card counts)
        if (k < 8) lo |= c << (4 * k); else hi |= c << (4 * (k - 8));
    }
    Desc v; v.lo = lo; v.hi = hi; v.s = (i % 9 >= 7) ? 0.4375f : 1.f; v.pad = 0.f;
    d[i] = v;
}

// ROWS rows per warp, WARPS warps per CTA; lanes 0..29 = (row parity, rank): two rows per warp store instruction
template <int ROWS, int WARPS, bool FROM_DESC>
__global__ void __launch_bounds__(WARPS * 32) k_rows(float4* __restrict__ out, const Desc* __restrict__ desc, int nwarps, unsigned seed) {
    __shared__ float4 lut[8];
    if (threadIdx.x < 5) lut[threadIdx.x] = make_float4(threadIdx.x > 0, threadIdx.x > 1, threadIdx.x > 2, threadIdx.x > 3);
    __syncthreads();
    const int w = blockIdx.x * WARPS + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (w >= nwarps || lane >= 30) return;
    const int k = lane % 15;
    const size_t row0 = (size_t)w * ROWS;
    float4* dst = out + row0 * 15 + lane;
    unsigned x = seed ^ (w * 2654435761u) ^ (k * 40503u);
#pragma unroll 4
    for (int r = lane / 15; r < ROWS; r += 2, dst += 30) {
        float4 q;
        if (FROM_DESC) {
            const Desc d = desc[row0 + r];
            const unsigned c = ((k < 8 ? d.lo : d.hi) >> (4 * (k & 7))) & 15u;
            q = lut[c];
            q.x *= d.s; q.y *= d.s; q.z *= d.s; q.w *= d.s;
        } else {
            x = x * 1664525u + 1013904223u;
            q = lut[(x >> 24) % 9 < 5 ? 0u : ((x >> 16) & 3u) + 1u];
        }
        __stcs(dst, q);
    }
}

template <class F> static float best_ms(F launch) {
    cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    float best = 1e9f;
    for (int i = 0; i < 6; i++) {
        CK(cudaEventRecord(a)); launch(); CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
        float ms; CK(cudaEventElapsedTime(&ms, a, b)); if (i > 0 && ms < best) best = ms;
    }
    CK(cudaGetLastError());
    return best;
}

static float4* alloc_vmm(size_t bytes, bool compressible) {
    CUmemAllocationProp prop = {};
    prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
    prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
    prop.location.id = 0;
    if (compressible) prop.allocFlags.compressionType = CU_MEM_ALLOCATION_COMP_GENERIC;
    size_t gran = 0;
    CU(cuMemGetAllocationGranularity(&gran, &prop, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED));
    const size_t size = (bytes + gran - 1) / gran * gran;
    CUmemGenericAllocationHandle h;
    CU(cuMemCreate(&h, size, &prop, 0));
    CUdeviceptr p;
    CU(cuMemAddressReserve(&p, size, 0, 0, 0));
    CU(cuMemMap(p, size, 0, h, 0));
    CUmemAccessDesc acc = {};
    acc.location = prop.location; acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
    CU(cuMemSetAccess(p, size, &acc, 1));
    return (float4*)p;
}

template <int ROWS, int WARPS>
static void run(const char* mem, float4* out, const Desc* desc, size_t nrows, size_t bytes) {
    const int nwarps = (int)(nrows / ROWS), grid = (nwarps + WARPS - 1) / WARPS;
    float ms = best_ms([&] { k_rows<ROWS, WARPS, false><<<grid, WARPS * 32>>>(out, desc, nwarps, 7u); });
    printf(", \"%s_rows%d_x%d_prng_GBs\": %.0f", mem, ROWS, WARPS, bytes / ms / 1e6);
    ms = best_ms([&] { k_rows<ROWS, WARPS, true><<<grid, WARPS * 32>>>(out, desc, nwarps, 7u); });
    printf(", \"%s_rows%d_x%d_desc_GBs\": %.0f", mem, ROWS, WARPS, bytes / ms / 1e6);
}

int main() {
    const size_t nrows = (size_t)4096 * 480, bytes = nrows * 240;            // one env-step's rows: 472 MB
    CK(cudaFree(0));
    int supported = 0; CUdevice dev; CU(cuDeviceGet(&dev, 0));
    CU(cuDeviceGetAttribute(&supported, CU_DEVICE_ATTRIBUTE_GENERIC_COMPRESSION_SUPPORTED, dev));
    printf("{\"rows\": %zu, \"bytes\": %zu, \"generic_compression_supported\": %d", nrows, bytes, supported);
    Desc* desc; CK(cudaMalloc(&desc, nrows * sizeof(Desc)));
    k_fill_desc<<<(unsigned)((nrows + 255) / 256), 256>>>(desc, nrows, 3u);
    CK(cudaDeviceSynchronize());
    float4* plain = alloc_vmm(bytes, false);
    float4* comp = supported ? alloc_vmm(bytes, true) : nullptr;
    struct { const char* name; float4* p; } bufs[] = {{"plain", plain}, {"compressible", comp}};
    for (auto& b : bufs) {
        if (!b.p) continue;
        run<480, 4>(b.name, b.p, desc, nrows, bytes);
        run<120, 4>(b.name, b.p, desc, nrows, bytes);
        run<60, 4>(b.name, b.p, desc, nrows, bytes);
        run<60, 1>(b.name, b.p, desc, nrows, bytes);
        run<30, 2>(b.name, b.p, desc, nrows, bytes);
    }
    printf("}\n");
    return 0;
}
