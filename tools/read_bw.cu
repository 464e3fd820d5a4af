// read_bw.cu — practical read-only HBM bandwidth on this GPU, to put the scan kernel's
// achieved GB/s in context (the driver's MEASURED_PEAKS.json figure is a read+write copy).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o read_bw tools/read_bw.cu && ./read_bw
// Variants: (a) LDG.128 grid-stride sum, (b) cp.async.bulk (TMA) per-warp 12 KB tiles into a
// 2-stage smem ring with no math at all (the scan kernel's exact access pattern).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(1024) ldg_sum(const float4* __restrict__ p, size_t n4, float* out) {
    float acc = 0.f;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i + 3 * stride < n4; i += 4 * stride) {
        float4 a = __ldcs(p + i), b = __ldcs(p + i + stride), c = __ldcs(p + i + 2 * stride), d = __ldcs(p + i + 3 * stride);
        acc += a.x + b.y + c.z + d.w;
    }
    for (; i < n4; i += stride) acc += __ldcs(p + i).x;
    if (acc == 123.456f) *out = acc;
}

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int WARPS, int STAGES>
__global__ void __launch_bounds__(WARPS * 32, 1) tma_stream(const uint8_t* __restrict__ base, size_t n_tiles, int tile_bytes, float* out) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)WARPS * STAGES * tile_bytes);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x < WARPS * STAGES) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bars[threadIdx.x])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();
    const size_t gw = (size_t)blockIdx.x * WARPS + warp, W = (size_t)gridDim.x * WARPS;
    const uint32_t st0 = s32(smem) + warp * STAGES * tile_bytes, b0 = s32(&bars[warp * STAGES]);
    auto issue = [&](size_t t, int s) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b0 + s * 8), "r"(tile_bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(st0 + s * tile_bytes), "l"(base + t * tile_bytes), "r"(tile_bytes), "r"(b0 + s * 8) : "memory");
    };
    if (lane == 0) for (int s = 0; s < STAGES; ++s) if (gw + s * W < n_tiles) issue(gw + s * W, s);
    int stage = 0; uint32_t par = 0; float acc = 0.f;
    for (size_t t = gw; t < n_tiles; t += W) {
        uint32_t ok = 0;
        while (!ok) asm volatile("{.reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0,1,0,p;}" : "=r"(ok) : "r"(b0 + stage * 8), "r"(par) : "memory");
        acc += reinterpret_cast<const float*>(smem + ((size_t)warp * STAGES + stage) * tile_bytes)[lane];
        __syncwarp();
        if (lane == 0 && t + STAGES * W < n_tiles) issue(t + STAGES * W, stage);
        if (++stage == STAGES) { stage = 0; par ^= 1; }
    }
    if (acc == 123.456f) *out = acc;
}

int main() {
    const size_t bytes = 12288000000ull;
    uint8_t* d; float* out;
    cudaMalloc(&d, bytes); cudaMalloc(&out, 4); cudaMemset(d, 1, bytes);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    auto time = [&](auto fn, const char* name) {
        for (int i = 0; i < 3; ++i) fn();
        cudaEventRecord(e0);
        for (int i = 0; i < 20; ++i) fn();
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("%-44s %8.1f GB/s  (%.3f ms)  %s\n", name, bytes / (ms / 20 * 1e-3) / 1e9, ms / 20, cudaGetErrorString(cudaGetLastError()));
    };
    for (int mult : {2, 4, 8, 16})
        time([&] { ldg_sum<<<148 * mult, 1024 / (mult > 2 ? 2 : 1)>>>((const float4*)d, bytes / 16, out); }, mult == 2 ? "ldg.128 148x2 CTAs x1024" : mult == 4 ? "ldg.128 148x4 x512" : mult == 8 ? "ldg.128 148x8 x512" : "ldg.128 148x16 x512");
    const int tb = 12288;
    {
        auto k = tma_stream<8, 2>; int sm = 8 * 2 * tb + 256;
        cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, sm);
        time([&] { k<<<148, 256, sm>>>(d, bytes / tb, tb, out); }, "tma bulk 12KB tiles, 8 warps x 2 stages");
    }
    {
        auto k = tma_stream<4, 4>; int sm = 4 * 4 * tb + 256;
        cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, sm);
        time([&] { k<<<148, 128, sm>>>(d, bytes / tb, tb, out); }, "tma bulk 12KB tiles, 4 warps x 4 stages");
    }
    {
        auto k = tma_stream<16, 1>; int sm = 16 * 1 * tb + 256;
        cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, sm);
        time([&] { k<<<148, 512, sm>>>(d, bytes / tb, tb, out); }, "tma bulk 12KB tiles, 16 warps x 1 stage");
    }
    {
        auto k = tma_stream<8, 4>; int sm = 8 * 4 * 6144 + 256;
        cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, sm);
        time([&] { k<<<148, 256, sm>>>(d, bytes / 6144, 6144, out); }, "tma bulk 6KB tiles, 8 warps x 4 stages");
    }
    {
        auto k = tma_stream<4, 2>; int sm = 4 * 2 * 24576 + 256;
        cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, sm);
        time([&] { k<<<148, 128, sm>>>(d, bytes / 24576, 24576, out); }, "tma bulk 24KB tiles, 4 warps x 2 stages");
    }
    return 0;
}
