// Micro-benchmarks that size the counting design on B200: shared/global atomic and scatter rates.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
typedef uint32_t u32; typedef uint64_t u64; typedef unsigned long long ull;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA %s line %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)
__device__ __forceinline__ u64 mix(u64 x) { x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33; return x; }

__global__ void k_smem_add(u32* out, int iters, u32 bins) {
    extern __shared__ u32 h[];
    for (u32 i = threadIdx.x; i < bins; i += blockDim.x) h[i] = 0;
    __syncthreads();
    u64 s = mix(blockIdx.x * 1024 + threadIdx.x + 1);
    for (int i = 0; i < iters; ++i) { s = s * 6364136223846793005ULL + 1442695040888963407ULL; atomicAdd(&h[(u32)(s >> 40) % bins], 1u); }
    __syncthreads();
    if (threadIdx.x == 0) out[blockIdx.x] = h[0];
}
__global__ void k_smem_add_ret(u32* out, int iters, u32 bins) {
    extern __shared__ u32 h[];
    for (u32 i = threadIdx.x; i < bins; i += blockDim.x) h[i] = 0;
    __syncthreads();
    u64 s = mix(blockIdx.x * 1024 + threadIdx.x + 1);
    u32 acc = 0;
    for (int i = 0; i < iters; ++i) { s = s * 6364136223846793005ULL + 1442695040888963407ULL; acc += atomicAdd(&h[(u32)(s >> 40) & (bins - 1)], 1u); }
    __syncthreads();
    if (threadIdx.x == 0) out[blockIdx.x] = h[0] + acc;
    if (acc == 0xFFFFFFFF) out[1] = acc;
}
__global__ void k_smem_or_ret(u32* out, int iters, u32 words) {
    extern __shared__ u32 h[];
    for (u32 i = threadIdx.x; i < words; i += blockDim.x) h[i] = 0;
    __syncthreads();
    u64 s = mix(blockIdx.x * 1024 + threadIdx.x + 1);
    u32 acc = 0;
    for (int i = 0; i < iters; ++i) { s = s * 6364136223846793005ULL + 1442695040888963407ULL; acc += atomicOr(&h[(u32)(s >> 40) & (words - 1)], 1u << ((u32)(s >> 35) & 31)) & 1; }
    __syncthreads();
    if (threadIdx.x == 0) out[blockIdx.x] = h[0] + acc;
}
__global__ void k_smem_add_noret_pow2(u32* out, int iters, u32 bins) {
    extern __shared__ u32 h[];
    for (u32 i = threadIdx.x; i < bins; i += blockDim.x) h[i] = 0;
    __syncthreads();
    u64 s = mix(blockIdx.x * 1024 + threadIdx.x + 1);
    for (int i = 0; i < iters; ++i) { s = s * 6364136223846793005ULL + 1442695040888963407ULL; atomicAdd(&h[(u32)(s >> 40) & (bins - 1)], 1u); }
    __syncthreads();
    if (threadIdx.x == 0) out[blockIdx.x] = h[0];
}
__global__ void k_ballot_split(u32* out, int iters) {       // 7-bit multisplit rank by ballots
    u64 s = mix((u64)blockIdx.x * blockDim.x + threadIdx.x + 1);
    u32 acc = 0;
    const u32 lt = (1u << (threadIdx.x & 31)) - 1u;
    for (int i = 0; i < iters; ++i) {
        s = s * 6364136223846793005ULL + 1442695040888963407ULL;
        const u32 d = (u32)(s >> 57);
        u32 peers = 0xffffffffu;
#pragma unroll
        for (int b = 0; b < 7; ++b) { const u32 m = __ballot_sync(0xffffffffu, (d >> b) & 1u); peers &= ((d >> b) & 1u) ? m : ~m; }
        acc += __popc(peers & lt);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
__global__ void k_smem_cas(u32* out, int iters, u32 slots /*pow2*/) {
    extern __shared__ ull t[];
    u32* cnt = (u32*)(t + slots);
    for (u32 i = threadIdx.x; i < slots; i += blockDim.x) { t[i] = ~0ull; cnt[i] = 0; }
    __syncthreads();
    u64 s = mix(blockIdx.x * 1024 + threadIdx.x + 1);
    for (int i = 0; i < iters; ++i) {
        s = s * 6364136223846793005ULL + 1442695040888963407ULL;
        const ull key = (s >> 20) % (slots / 2);           // ~50% load, with repeats
        u32 p = (u32)mix(key) & (slots - 1);
        while (true) {
            ull cur = t[p];
            if (cur == key) break;
            if (cur == ~0ull) { cur = atomicCAS(&t[p], ~0ull, key); if (cur == ~0ull || cur == key) break; }
            p = (p + 1) & (slots - 1);
        }
        atomicAdd(&cnt[p], 1u);
    }
    __syncthreads();
    if (threadIdx.x == 0) out[blockIdx.x] = cnt[0];
}
__global__ void k_gmem_red(u32* table, u64 bins, int iters) {
    u64 s = mix((u64)blockIdx.x * blockDim.x + threadIdx.x + 1);
    for (int i = 0; i < iters; ++i) { s = s * 6364136223846793005ULL + 1442695040888963407ULL; atomicAdd(&table[(s >> 24) % bins], 1u); }
}
__global__ void k_gmem_store8(ull* table, u64 n, int iters) {
    u64 s = mix((u64)blockIdx.x * blockDim.x + threadIdx.x + 1);
    for (int i = 0; i < iters; ++i) { s = s * 6364136223846793005ULL + 1442695040888963407ULL; table[(s >> 24) % n] = s; }
}
// bucketed scatter: nb buckets, each thread appends to bucket cursor (smem cursors), consecutive keys of a
// bucket land in consecutive addresses -> models a single-level many-way partition write pattern
__global__ void k_gmem_bucket_store(ull* table, u64 bucket_cap, u32 nb, int iters) {
    extern __shared__ u32 cur[];
    for (u32 i = threadIdx.x; i < nb; i += blockDim.x) cur[i] = 0;
    __syncthreads();
    u64 s = mix((u64)blockIdx.x * blockDim.x + threadIdx.x + 1);
    const u64 cta_off = (u64)blockIdx.x * (bucket_cap / gridDim.x);
    for (int i = 0; i < iters; ++i) {
        s = s * 6364136223846793005ULL + 1442695040888963407ULL;
        const u32 b = (u32)(s >> 40) % nb;
        const u32 slot = atomicAdd(&cur[b], 1u);
        table[(u64)b * bucket_cap + cta_off + slot] = s;
    }
}
__global__ void k_gmem_cas(ull* table, u32* cnt, u64 slots /*pow2*/, int iters) {
    u64 s = mix((u64)blockIdx.x * blockDim.x + threadIdx.x + 1);
    for (int i = 0; i < iters; ++i) {
        s = s * 6364136223846793005ULL + 1442695040888963407ULL;
        const ull key = (s >> 16) % (slots / 2);
        u64 p = mix(key) & (slots - 1);
        while (true) {
            ull cur = table[p];
            if (cur == key) break;
            if (cur == ~0ull) { cur = atomicCAS(&table[p], ~0ull, key); if (cur == ~0ull || cur == key) break; }
            p = (p + 1) & (slots - 1);
        }
        atomicAdd(&cnt[p], 1u);
    }
}
__global__ void k_match(u32* out, int iters) {
    u64 s = mix((u64)blockIdx.x * blockDim.x + threadIdx.x + 1);
    u32 acc = 0;
    for (int i = 0; i < iters; ++i) { s = s * 6364136223846793005ULL + 1442695040888963407ULL; acc += __popc(__match_any_sync(0xffffffffu, (u32)(s >> 56))); }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
__global__ void k_copy(const uint4* in, uint4* out, u64 n) {
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x) out[i] = in[i];
}

template <class F> float timeit(F f, int reps = 3) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    f(); cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < reps; ++r) { cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b); float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms; }
    return best;
}
int main() {
    int sms = 148; cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0)); sms = p.multiProcessorCount;
    printf("device %s SMs %d\n", p.name, sms);
    u32* out; CK(cudaMalloc(&out, 1 << 24));
    const int iters = 2000;
    // shared-memory atomicAdd
    for (u32 bins : {64u, 1024u, 16384u, 32768u}) for (int threads : {256, 1024}) {
        int grid = sms * (threads == 256 ? 4 : 1);
        if (bins * 4 > 48 * 1024) cudaFuncSetAttribute(k_smem_add, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        float ms = timeit([&] { k_smem_add<<<grid, threads, bins * 4>>>(out, iters, bins); });
        printf("smem atomicAdd bins=%u threads=%d grid=%d: %.1f Gop/s\n", bins, threads, grid, (double)grid * threads * iters / ms / 1e6);
    }
    for (u32 bins : {128u, 4096u, 16384u}) for (int threads : {256, 1024}) {
        int grid = sms * (threads == 256 ? 4 : 1);
        float ms = timeit([&] { k_smem_add_ret<<<grid, threads, bins * 4>>>(out, iters, bins); });
        float ms2 = timeit([&] { k_smem_add_noret_pow2<<<grid, threads, bins * 4>>>(out, iters, bins); });
        float ms3 = timeit([&] { k_smem_or_ret<<<grid, threads, bins * 4>>>(out, iters, bins); });
        printf("smem atomics bins=%u threads=%d: add+return %.1f Gop/s, add no return %.1f Gop/s, or+return %.1f Gop/s\n", bins, threads,
               (double)grid * threads * iters / ms / 1e6, (double)grid * threads * iters / ms2 / 1e6, (double)grid * threads * iters / ms3 / 1e6);
    }
    {
        int grid = sms * 8, threads = 256;
        float ms = timeit([&] { k_ballot_split<<<grid, threads>>>(out, 1000); });
        printf("7-ballot multisplit rank: %.1f G lane-ops/s\n", (double)grid * threads * 1000 / ms / 1e6);
    }
    CK(cudaGetLastError());
    for (u32 slots : {4096u, 16384u}) for (int threads : {256, 1024}) {
        int grid = sms * (threads == 256 && slots == 4096 ? 4 : 1);
        size_t sm = (size_t)slots * 12;
        cudaFuncSetAttribute(k_smem_cas, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        float ms = timeit([&] { k_smem_cas<<<grid, threads, sm>>>(out, iters, slots); });
        printf("smem hash insert (CAS64+add) slots=%u threads=%d grid=%d: %.1f Gop/s\n", slots, threads, grid, (double)grid * threads * iters / ms / 1e6);
    }
    CK(cudaGetLastError());
    for (u64 mb : {16ull, 64ull, 256ull, 2048ull}) {
        u32* table; u64 bins = mb * 1024 * 1024 / 4; CK(cudaMalloc(&table, bins * 4)); CK(cudaMemset(table, 0, bins * 4));
        int grid = sms * 8, threads = 256;
        float ms = timeit([&] { k_gmem_red<<<grid, threads>>>(table, bins, 500); });
        printf("global RED u32 table=%lluMB: %.1f Gop/s\n", (ull)mb, (double)grid * threads * 500 / ms / 1e6);
        cudaFree(table);
    }
    {
        u64 n = 600ull * 1024 * 1024 / 8; ull* table; CK(cudaMalloc(&table, n * 8));
        int grid = sms * 8, threads = 256;
        float ms = timeit([&] { k_gmem_store8<<<grid, threads>>>(table, n, 500); });
        printf("global random 8B store into 600MB: %.1f Gop/s\n", (double)grid * threads * 500 / ms / 1e6);
        for (u32 nb : {256u, 1024u, 4096u, 8192u}) {
            int g2 = sms * 2, t2 = 1024, it2 = 64;       // per CTA: 64K keys
            u64 cap = n / nb;                              // slots per bucket
            float ms2 = timeit([&] { k_gmem_bucket_store<<<g2, t2, nb * 4>>>(table, cap, nb, it2); });
            printf("bucketed 8B store nb=%u (CTA-private runs, %.1f keys/bucket/CTA): %.1f Gop/s\n", nb, 65536.0 / nb, (double)g2 * t2 * it2 / ms2 / 1e6);
        }
        cudaFree(table);
    }
    for (u64 mb : {32ull, 1024ull}) {
        u64 slots = mb * 1024 * 1024 / 8; ull* table; u32* cnt; CK(cudaMalloc(&table, slots * 8)); CK(cudaMalloc(&cnt, slots * 4));
        int grid = sms * 8, threads = 256, it = 100;
        float ms = timeit([&] { cudaMemsetAsync(table, 0xff, slots * 8); cudaMemsetAsync(cnt, 0, slots * 4); k_gmem_cas<<<grid, threads>>>(table, cnt, slots, it); }, 2);
        float ms0 = timeit([&] { cudaMemsetAsync(table, 0xff, slots * 8); cudaMemsetAsync(cnt, 0, slots * 4); }, 2);
        printf("global hash insert (CAS64+RED) keys table=%lluMB: %.1f Gop/s (memset %.3f ms)\n", (ull)mb, (double)grid * threads * it / (ms - ms0) / 1e6, ms0);
        cudaFree(table); cudaFree(cnt);
    }
    {
        int grid = sms * 8, threads = 256;
        float ms = timeit([&] { k_match<<<grid, threads>>>(out, 1000); });
        printf("match_any: %.1f G lane-ops/s\n", (double)grid * threads * 1000 / ms / 1e6);
    }
    {
        u64 n = 1ull << 26; uint4 *a, *b; CK(cudaMalloc(&a, n * 16)); CK(cudaMalloc(&b, n * 16));
        float ms = timeit([&] { k_copy<<<sms * 16, 512>>>(a, b, n); });
        printf("copy 1 GiB: %.1f GB/s (read+write)\n", 2.0 * n * 16 / ms / 1e6);
    }
    return 0;
}
