// Microbenchmark: how fast can SM-issued stores fill HBM on a B200, against cudaMemsetAsync on the same buffer?
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o store_ceiling store_ceiling.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

enum { DEF = 0, CS = 1, CG = 2, WT = 3, NOALLOC = 4, EVF = 5 };
template <int KIND>
__device__ __forceinline__ void st(float4* p, float4 v, unsigned long long pol) {
    if (KIND == DEF) asm volatile("st.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
    if (KIND == CS) asm volatile("st.global.cs.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
    if (KIND == CG) asm volatile("st.global.cg.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
    if (KIND == WT) asm volatile("st.global.wt.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
    if (KIND == NOALLOC) asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
    if (KIND == EVF) asm volatile("st.global.L2::cache_hint.v4.f32 [%0], {%1,%2,%3,%4}, %5;" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "l"(pol) : "memory");
}
// every warp instruction = 512 contiguous bytes; all threads of the grid advance together
template <int KIND>
__global__ void __launch_bounds__(256) k_stride(float4* out, size_t nvec) {
    unsigned long long pol = 0;
    if (KIND == EVF) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
    float4 v = make_float4(1.f, 0.f, 1.f, (float)threadIdx.x);
#pragma unroll 8
    for (; i < nvec; i += stride) st<KIND>(out + i, v, pol);
}
// each CTA owns contiguous chunks of CH bytes (chunk-strided): fewer DRAM pages open at a time per CTA
template <int KIND>
__global__ void __launch_bounds__(256) k_chunk(float4* out, size_t nvec, int chunk_vec) {
    unsigned long long pol = 0;
    const size_t nchunks = nvec / chunk_vec;
    float4 v = make_float4(1.f, 0.f, 1.f, (float)threadIdx.x);
    for (size_t c = blockIdx.x; c < nchunks; c += gridDim.x) {
        float4* base = out + c * chunk_vec;
#pragma unroll 8
        for (int i = threadIdx.x; i < chunk_vec; i += 256) st<KIND>(base + i, v, pol);
    }
}
// TMA bulk stores from a shared-memory tile, CH bytes per instruction
__global__ void __launch_bounds__(128) k_bulk(float4* out, size_t nvec, int chunk_vec) {
    extern __shared__ __align__(128) float4 tile[];
    for (int i = threadIdx.x; i < chunk_vec; i += 128) tile[i] = make_float4(1.f, 0.f, 1.f, (float)i);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    const size_t nchunks = nvec / chunk_vec;
    if (threadIdx.x == 0) {
        const unsigned s = (unsigned)__cvta_generic_to_shared(tile);
        int inflight = 0;
        for (size_t c = blockIdx.x; c < nchunks; c += gridDim.x) {
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(out + c * chunk_vec), "r"(s), "r"(chunk_vec * 16) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            if (++inflight >= 8) { asm volatile("cp.async.bulk.wait_group.read 7;" ::: "memory"); }
        }
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
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
    const size_t bytes = (size_t)(argc > 1 ? atof(argv[1]) : 3.775e9) / 65536 * 65536, nvec = bytes / 16;
    float4* out; CK(cudaMalloc(&out, bytes));
    CK(cudaFuncSetAttribute(k_bulk, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
    printf("{\"bytes\": %zu", bytes);
#define REPORT(name, call) { float ms = best_ms([&] { call; }); printf(", \"%s\": %.0f", name, bytes / ms / 1e6); fflush(stdout); }
    REPORT("memset", CK(cudaMemsetAsync(out, 1, bytes)));
    REPORT("memsetD32", CK(cudaMemsetAsync(out, 0, bytes)));
    for (int g : {148, 296, 592, 1184, 2368, 4736}) {
        char nm[64]; snprintf(nm, sizeof nm, "stride_def_g%d", g);
        REPORT(nm, (k_stride<DEF><<<g, 256>>>(out, nvec)));
    }
    for (int g : {9472, 18944, 37888, 75776}) {
        char nm[64]; snprintf(nm, sizeof nm, "stride_def_g%d", g);
        REPORT(nm, (k_stride<DEF><<<g, 256>>>(out, nvec)));
    }
    REPORT("one_store_per_thread", (k_stride<DEF><<<(unsigned)((nvec + 255) / 256), 256>>>(out, nvec)));
    REPORT("one_store_per_thread_cs", (k_stride<CS><<<(unsigned)((nvec + 255) / 256), 256>>>(out, nvec)));
    REPORT("four_stores_per_thread", (k_stride<DEF><<<(unsigned)((nvec + 1023) / 1024), 256>>>(out, nvec)));
    REPORT("chunk4096_g_all", (k_chunk<DEF><<<(unsigned)(nvec / 256), 256>>>(out, nvec, 256)));
    REPORT("chunk16384_g_all", (k_chunk<DEF><<<(unsigned)(nvec / 1024), 256>>>(out, nvec, 1024)));
    REPORT("chunk65536_g_all", (k_chunk<DEF><<<(unsigned)(nvec / 4096), 256>>>(out, nvec, 4096)));
    REPORT("chunk131072_g_all", (k_chunk<DEF><<<(unsigned)(nvec / 8192), 256>>>(out, nvec, 8192)));
    REPORT("stride_cs_g1184", (k_stride<CS><<<1184, 256>>>(out, nvec)));
    REPORT("stride_cg_g1184", (k_stride<CG><<<1184, 256>>>(out, nvec)));
    REPORT("stride_wt_g1184", (k_stride<WT><<<1184, 256>>>(out, nvec)));
    REPORT("stride_noalloc_g1184", (k_stride<NOALLOC><<<1184, 256>>>(out, nvec)));
    REPORT("stride_evictfirst_g1184", (k_stride<EVF><<<1184, 256>>>(out, nvec)));
    for (int ch : {4096, 16384, 65536}) {
        char nm[64]; snprintf(nm, sizeof nm, "chunk%d_def_g1184", ch);
        REPORT(nm, (k_chunk<DEF><<<1184, 256>>>(out, nvec, ch / 16)));
    }
    for (int ch : {4096, 16384, 65536}) {
        char nm[64]; snprintf(nm, sizeof nm, "bulk%d_g592", ch);
        REPORT(nm, (k_bulk<<<592, 128, ch>>>(out, nvec, ch / 16)));
    }
    REPORT("bulk16384_g148", (k_bulk<<<148, 128, 16384>>>(out, nvec, 1024)));
    REPORT("bulk16384_g1184", (k_bulk<<<1184, 128, 16384>>>(out, nvec, 1024)));
    printf("}\n");
    return 0;
}
