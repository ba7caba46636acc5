// Micro-benchmark: how fast can ONE SM stream K/V-like tiles from global memory into shared memory / registers?
//   mode 0: 1-D bulk async copies (cp.async.bulk, TMA engine), `nsub` tiles in flight per warp, issued by lane 0
//   mode 1: coalesced LDG.128 by the whole warp (4 rows of 128 B per instruction), unrolled `nsub` x 8 deep
// Each of 8 warps of a CTA walks its own region of a large buffer (footprint chosen to hit or miss L2).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench_kv tools/ubench_kv.cu ; run: ./ubench_kv
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c)); }
__device__ __forceinline__ void mbar_expect(uint64_t* b, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
    asm volatile("{\n.reg .pred p;\nW1:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D1;\nbra W1;\nD1:\n}\n" ::"r"(smem_u32(b)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk(void* dst, const void* src, uint32_t bytes, uint64_t* bar, uint64_t pol, int hint) {
    if (hint)
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(pol) : "memory");
    else
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// each warp: `ntiles` tiles of `bytes` bytes, tile i of warp w of CTA b at offset ((b * 8 + w) * ntiles + i) * stride
__global__ void __launch_bounds__(256, 1) k_bulk(const uint8_t* g, size_t stride, int bytes, int ntiles, int nsub, int hint, float* sink, long long* cyc) {
    extern __shared__ __align__(128) uint8_t sm[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm);                 // [8][8]
    uint8_t* buf = sm + 1024 + (size_t)warp * nsub * bytes;
    if (threadIdx.x == 0) {
        for (int i = 0; i < 64; ++i) mbar_init(&bars[i], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    uint64_t pol = 0;
    if (hint == 1) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    if (hint == 2) asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    const uint8_t* base = g + ((size_t)(blockIdx.x * 8 + warp) * ntiles) * stride;
    float acc = 0.f;
    const long long t0 = clock64();
    if (lane == 0)
        for (int i = 0; i < nsub && i < ntiles; ++i) { mbar_expect(&bars[warp * 8 + i], bytes); bulk(buf + i * bytes, base + (size_t)i * stride, bytes, &bars[warp * 8 + i], pol, hint); }
    for (int i = 0; i < ntiles; ++i) {
        const int s = i % nsub;
        mbar_wait(&bars[warp * 8 + s], (i / nsub) & 1);
        acc += reinterpret_cast<const float*>(buf + s * bytes)[lane];      // touch the tile
        __syncwarp();
        if (lane == 0 && i + nsub < ntiles) { mbar_expect(&bars[warp * 8 + s], bytes); bulk(buf + s * bytes, base + (size_t)(i + nsub) * stride, bytes, &bars[warp * 8 + s], pol, hint); }
    }
    const long long t1 = clock64();
    if (acc == 12345.678f) sink[0] = acc;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

__global__ void __launch_bounds__(256, 1) k_ldg(const uint8_t* g, size_t stride, int bytes, int ntiles, int unroll, float* sink, long long* cyc) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint8_t* base = g + ((size_t)(blockIdx.x * 8 + warp) * ntiles) * stride;
    float acc = 0.f;
    const long long t0 = clock64();
    const int per = bytes / 512;      // LDG.128 warp instructions per tile
    for (int i = 0; i < ntiles; ++i) {
        const float4* p = reinterpret_cast<const float4*>(base + (size_t)i * stride) + lane;
        if (unroll == 8) {
#pragma unroll 8
            for (int j = 0; j < per; ++j) { const float4 v = __ldcs(p + j * 32); acc += v.x + v.y + v.z + v.w; }
        } else {
#pragma unroll 16
            for (int j = 0; j < per; ++j) { const float4 v = __ldcs(p + j * 32); acc += v.x + v.y + v.z + v.w; }
        }
    }
    const long long t1 = clock64();
    if (acc == 12345.678f) sink[0] = acc;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

int main() {
    const size_t footprint_big = (size_t)3 << 30, footprint_small = (size_t)48 << 20;
    uint8_t* g;
    cudaMalloc(&g, footprint_big);
    cudaMemset(g, 0, footprint_big);
    float* sink; long long* cyc;
    cudaMalloc(&sink, 4); cudaMalloc(&cyc, 148 * 8);
    cudaFuncSetAttribute(k_bulk, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    long long h[148];
    printf("mode,ctas,bytes,inflight,hint,footprint,cycles_per_tile_per_warp,B_per_clk_per_SM\n");
    for (int ctas : {16, 148})
        for (int big = 0; big < 2; ++big)
            for (int bytes : {2048, 4096, 9216, 18432})
                for (int nsub : {1, 2, 4, 8}) {
                    if ((size_t)nsub * bytes * 8 > 190 * 1024) continue;
                    for (int hint : {0, 1, 2}) {
                        const size_t fp = big ? footprint_big : footprint_small;
                        int ntiles = (int)(fp / ((size_t)ctas * 8 * bytes));
                        if (ntiles > 512) ntiles = 512;
                        for (int rep = 0; rep < 2; ++rep)       // second pass: L2-warm for the small footprint
                            k_bulk<<<ctas, 256, 1024 + 8 * nsub * bytes>>>(g, bytes, bytes, ntiles, nsub, hint, sink, cyc);
                        cudaError_t e = cudaDeviceSynchronize();
                        if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
                        cudaMemcpy(h, cyc, ctas * 8, cudaMemcpyDeviceToHost);
                        double mx = 0; for (int i = 0; i < ctas; ++i) mx = h[i] > mx ? h[i] : mx;
                        printf("bulk,%d,%d,%d,%d,%s,%.0f,%.1f\n", ctas, bytes, nsub, hint, big ? "3GB" : "48MB", mx / ntiles, 8.0 * ntiles * bytes / mx);
                    }
                }
    for (int ctas : {16, 148})
        for (int big = 0; big < 2; ++big)
            for (int bytes : {4096, 18432})
                for (int unroll : {8, 16}) {
                    const size_t fp = big ? footprint_big : footprint_small;
                    int ntiles = (int)(fp / ((size_t)ctas * 8 * bytes));
                    if (ntiles > 512) ntiles = 512;
                    for (int rep = 0; rep < 2; ++rep) k_ldg<<<ctas, 256>>>(g, bytes, bytes, ntiles, unroll, sink, cyc);
                    cudaDeviceSynchronize();
                    cudaMemcpy(h, cyc, ctas * 8, cudaMemcpyDeviceToHost);
                    double mx = 0; for (int i = 0; i < ctas; ++i) mx = h[i] > mx ? h[i] : mx;
                    printf("ldg,%d,%d,%d,-,%s,%.0f,%.1f\n", ctas, bytes, unroll, big ? "3GB" : "48MB", mx / ntiles, 8.0 * ntiles * bytes / mx);
                }
    return 0;
}
