// Microbenchmark: do k_env's thermometer rows (floats that are 0 or 1, or 0 or a scale) stream faster into a
// COMPRESSIBLE allocation (cuMemCreate with CU_MEM_ALLOCATION_COMP_GENERIC: the L2 compresses lines on their way to HBM)?
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o compressible compressible.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)
#define CU(x) do { CUresult e = (x); if (e != CUDA_SUCCESS) { const char* s; cuGetErrorString(e, &s); printf("driver error %s at %d\n", s, __LINE__); exit(1); } } while (0)

constexpr int kRowsPerWarp = 480, kVecPerWarp = kRowsPerWarp * 15;

// the row writer of k_env: lanes 0..29 = (row parity, rank), values from a thermometer LUT, realistic nibble contents
__global__ void __launch_bounds__(128, 7) k_rows(float4* out, int nwarps, unsigned seed) {
    __shared__ float4 lut[8];
    if (threadIdx.x < 5) lut[threadIdx.x] = make_float4(threadIdx.x > 0, threadIdx.x > 1, threadIdx.x > 2, threadIdx.x > 3);
    __syncthreads();
    int w = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (w >= nwarps || lane >= 30) return;
    const int k = lane % 15;
    float4* dst = out + (size_t)w * kVecPerWarp + lane;
    unsigned x = seed ^ (w * 2654435761u) ^ (k * 40503u);
#pragma unroll 4
    for (int r = lane / 15; r < kRowsPerWarp; r += 2, dst += 30) {
        x = x * 1664525u + 1013904223u;
        const unsigned c = (x >> 24) % 9 < 5 ? 0u : ((x >> 16) & 3u) + 1u;      // mostly 0, else 1..4 (like card counts)
        __stcs(dst, lut[c]);
    }
}
__global__ void k_read(const float4* in, size_t nvec, float* sink) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
    float acc = 0.f;
    for (; i < nvec; i += stride) { float4 v = __ldcs(in + i); acc += v.x + v.y + v.z + v.w; }
    if (acc == -1.f) *sink = acc;
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

static float4* alloc_vmm(size_t bytes, bool compressible, size_t* mapped) {
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
    CUmemAllocationProp got = {};
    CU(cuMemGetAllocationPropertiesFromHandle(&got, h));
    printf(", \"%s_compression_type\": %d", compressible ? "comp" : "plain", (int)got.allocFlags.compressionType);
    CUdeviceptr p;
    CU(cuMemAddressReserve(&p, size, 0, 0, 0));
    CU(cuMemMap(p, size, 0, h, 0));
    CUmemAccessDesc acc = {};
    acc.location = prop.location; acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
    CU(cuMemSetAccess(p, size, &acc, 1));
    *mapped = size;
    return (float4*)p;
}

int main(int argc, char** argv) {
    const int nwarps = argc > 1 ? atoi(argv[1]) : 4096, grid = nwarps / 4;
    const size_t nvec = (size_t)nwarps * kVecPerWarp, bytes = nvec * 16;
    CK(cudaFree(0));
    int supported = 0; CUdevice dev; CU(cuDeviceGet(&dev, 0));
    CU(cuDeviceGetAttribute(&supported, CU_DEVICE_ATTRIBUTE_GENERIC_COMPRESSION_SUPPORTED, dev));
    printf("{\"nwarps\": %d, \"bytes\": %zu, \"generic_compression_supported\": %d", nwarps, bytes, supported);
    float* sink; CK(cudaMalloc(&sink, 4));
    float4* plain; CK(cudaMalloc(&plain, bytes));
    size_t m1 = 0, m2 = 0;
    float4* vmm_plain = alloc_vmm(bytes, false, &m1);
    float4* vmm_comp = supported ? alloc_vmm(bytes, true, &m2) : nullptr;
    struct { const char* name; float4* p; } bufs[] = {{"cudaMalloc", plain}, {"vmm_plain", vmm_plain}, {"vmm_compressible", vmm_comp}};
    for (auto& b : bufs) {
        if (!b.p) continue;
        float ms = best_ms([&] { k_rows<<<grid, 128>>>(b.p, nwarps, 7u); });
        printf(", \"%s_write_GBs\": %.0f", b.name, bytes / ms / 1e6);
        ms = best_ms([&] { for (int s = 0; s < 8; s++) k_rows<<<grid, 128>>>(b.p, nwarps, 7u + s); });
        printf(", \"%s_write_8_launches_GBs\": %.0f", b.name, 8 * bytes / ms / 1e6);
        ms = best_ms([&] { k_read<<<148 * 16, 256>>>(b.p, nvec, sink); });
        printf(", \"%s_read_GBs\": %.0f", b.name, bytes / ms / 1e6);
    }
    printf("}\n");
    return 0;
}
