// mercat2_b200 -- shared device/host helpers (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <stdexcept>

typedef uint8_t u8;
typedef uint16_t u16;
typedef uint32_t u32;
typedef uint64_t u64;
typedef int64_t i64;
typedef unsigned long long ull;

#define MC2_SEP 0xFFu   // record separator in the compacted symbol stream (inputs are 7-bit ASCII)

struct Mc2Error : std::runtime_error {
    int code;
    Mc2Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

#define CUDA_CHECK(expr)                                                                        \
    do {                                                                                        \
        cudaError_t _e = (expr);                                                                \
        if (_e != cudaSuccess) {                                                                \
            char _buf[512];                                                                     \
            snprintf(_buf, sizeof _buf, "CUDA error %s at %s:%d: %s", cudaGetErrorName(_e),     \
                     __FILE__, __LINE__, cudaGetErrorString(_e));                               \
            throw Mc2Error(-2, _buf);                                                           \
        }                                                                                       \
    } while (0)

static inline u64 div_up(u64 a, u64 b) { return (a + b - 1) / b; }

// Block barrier that first reconverges the warp.  Measured on B200 / CUDA 12.9: after a data-dependent loop
// (hash probing, byte compares) the lanes of a warp can reach BAR.SYNC at different times even though ptxas
// emitted BSSY/BSYNC.RECONVERGENT before it, and the barrier then releases other warps while the late lanes
// are still inserting -- counts were lost.  __syncwarp() (WARPSYNC.ALL) before the barrier fixes it; every
// barrier in this code base goes through this macro.
// (the warp barrier is inline asm so that the compiler cannot drop it as "redundant after BSYNC")
#define BLOCK_SYNC() do { asm volatile("bar.warp.sync 0xffffffff;" ::: "memory"); __syncthreads(); } while (0)

// ---------------------------------------------------------------------------------------------
// Block-wide exclusive scan of a u32 with an arbitrary associative operator.
// Op::combine(a, b) = "a followed by b".  All threads of the block must call.  NWARPS = blockDim/32.
// Returns the combination of all elements before this thread (identity for thread 0); *total (if
// given) receives the combination over the whole block.
// ---------------------------------------------------------------------------------------------
//
// NOTE on convergence: every block-/warp-collective helper here starts with __syncwarp().  Measured on
// B200 (sm_100a, CUDA 12.9): after a data-dependent loop with early exits the compiler's
// BSSY/BSYNC.RECONVERGENT pair did NOT leave the warp converged for the SHFL / BAR.RED that followed
// (shuffle-based scans returned garbage, __syncthreads_count under-counted); an explicit WARPSYNC
// fixes both.  For the same reason __syncthreads_count/_or are not used: see block_count().
template <class Op, int NWARPS>
__device__ __forceinline__ u32 block_exclusive_scan(u32 x, u32* smem /*NWARPS+1 words*/, u32* total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __syncwarp();
    u32 incl = x;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        u32 y = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl = Op::combine(y, incl);
    }
    u32 excl = __shfl_up_sync(0xffffffffu, incl, 1);
    if (lane == 0) excl = Op::identity();
    BLOCK_SYNC();                       // protect smem reuse between consecutive calls
    if (lane == 31) smem[warp] = incl;
    BLOCK_SYNC();
    if (warp == 0) {
        u32 w = lane < NWARPS ? smem[lane] : Op::identity();
        u32 wi = w;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            u32 y = __shfl_up_sync(0xffffffffu, wi, d);
            if (lane >= d) wi = Op::combine(y, wi);
        }
        u32 we = __shfl_up_sync(0xffffffffu, wi, 1);
        if (lane == 0) we = Op::identity();
        if (lane < NWARPS) smem[lane] = we;
        if (lane == NWARPS - 1) smem[NWARPS] = wi;
    }
    BLOCK_SYNC();
    u32 base = smem[warp];
    if (total) *total = smem[NWARPS];
    return Op::combine(base, excl);
}

struct OpAdd {
    __device__ static __forceinline__ u32 identity() { return 0u; }
    __device__ static __forceinline__ u32 combine(u32 a, u32 b) { return a + b; }
};

// 64-bit sum variant (two words through the same smem area: needs 2*(NWARPS+1) words)
template <int NWARPS>
__device__ __forceinline__ u64 block_exclusive_sum64(u64 x, u64* smem /*NWARPS+1*/, u64* total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __syncwarp();
    u64 incl = x;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        u64 y = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += y;
    }
    u64 excl = incl - x;
    BLOCK_SYNC();
    if (lane == 31) smem[warp] = incl;
    BLOCK_SYNC();
    if (warp == 0) {
        u64 w = lane < NWARPS ? smem[lane] : 0ull;
        u64 wi = w;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            u64 y = __shfl_up_sync(0xffffffffu, wi, d);
            if (lane >= d) wi += y;
        }
        if (lane < NWARPS) smem[lane] = wi - w;
        if (lane == NWARPS - 1) smem[NWARPS] = wi;
    }
    BLOCK_SYNC();
    u64 base = smem[warp];
    if (total) *total = smem[NWARPS];
    return base + excl;
}

// number of threads of the block whose predicate is true (all threads must call; <= 32 warps)
__device__ __forceinline__ u32 block_count(bool pred) {
    __shared__ u32 s_block_count[33];
    __syncwarp();
    const u32 b = __ballot_sync(0xffffffffu, pred);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    BLOCK_SYNC();
    if (lane == 0) s_block_count[warp] = __popc(b);
    BLOCK_SYNC();
    if (warp == 0) {
        const int nw = (blockDim.x + 31) >> 5;
        u32 v = lane < nw ? s_block_count[lane] : 0u;
#pragma unroll
        for (int d = 16; d; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
        if (lane == 0) s_block_count[32] = v;
    }
    BLOCK_SYNC();
    return s_block_count[32];
}

__device__ __forceinline__ uint4 ld_nc_16(const void* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
