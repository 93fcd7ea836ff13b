// Device-wide primitives written for this engine: exclusive scans, an LSD radix sort (8-bit digits,
// stable, keys-only or key+payload) and run-length / reduce-by-key kernels over sorted sequences.
// They are the sparse (large-k) counting path that replaces the reference's dict
// (lib/mercat2_kmers.py:56-60 "kmerlist[k] += 1") and its serial per-sample merge + sorted()
// (bin/mercat2.py:121-132).
#pragma once
#include "common.cuh"

#define SCAN_THREADS 256
#define SCAN_ITEMS 16
#define SCAN_TILE (SCAN_THREADS * SCAN_ITEMS)

// ---- generic exclusive scan: tile reduce -> single-CTA scan of tile sums -> tile downsweep -------
template <typename T>
__global__ void __launch_bounds__(SCAN_THREADS) scan_reduce_kernel(const T* __restrict__ in, u64 n, u64* __restrict__ tile_sum) {
    __shared__ u64 sm[SCAN_THREADS / 32 + 1];
    const u64 base = (u64)blockIdx.x * SCAN_TILE;
    u64 acc = 0;
#pragma unroll
    for (int j = 0; j < SCAN_ITEMS; ++j) {
        const u64 i = base + (u64)j * SCAN_THREADS + threadIdx.x;
        if (i < n) acc += (u64)in[i];
    }
    u64 total;
    block_exclusive_sum64<SCAN_THREADS / 32>(acc, sm, &total);
    if (threadIdx.x == 0) tile_sum[blockIdx.x] = total;
}

#ifndef SCAN1_THREADS
#define SCAN1_THREADS 1024
#endif
__global__ void __launch_bounds__(SCAN1_THREADS) scan_sums_kernel(u64* __restrict__ sums, u32 n, ull* total_out) {
    __shared__ u64 sm[SCAN1_THREADS / 32 + 1];
    const u32 per = (n + SCAN1_THREADS - 1) / SCAN1_THREADS;
    const u32 t0 = min(n, threadIdx.x * per), t1 = min(n, t0 + per);
    u64 acc = 0;
    for (u32 t = t0; t < t1; ++t) acc += sums[t];
    u64 total;
    u64 base = block_exclusive_sum64<SCAN1_THREADS / 32>(acc, sm, &total);
    for (u32 t = t0; t < t1; ++t) { const u64 v = sums[t]; sums[t] = base; base += v; }
    if (threadIdx.x == 0 && total_out) *total_out = total;
}

// thread t owns SCAN_ITEMS consecutive elements (blocked order) so that the scan is a plain
// thread-sum + block scan + thread rescan; in == out is allowed.
template <typename T, typename O>
__global__ void __launch_bounds__(SCAN_THREADS) scan_down_kernel(const T* in, O* out, u64 n, const u64* __restrict__ tile_base) {
    __shared__ u64 sm[SCAN_THREADS / 32 + 1];
    const u64 first = (u64)blockIdx.x * SCAN_TILE + (u64)threadIdx.x * SCAN_ITEMS;
    T v[SCAN_ITEMS];
    u64 acc = 0;
#pragma unroll
    for (int j = 0; j < SCAN_ITEMS; ++j) {
        v[j] = (first + j < n) ? in[first + j] : (T)0;
        acc += (u64)v[j];
    }
    u64 run = block_exclusive_sum64<SCAN_THREADS / 32>(acc, sm, nullptr) + tile_base[blockIdx.x];
#pragma unroll
    for (int j = 0; j < SCAN_ITEMS; ++j) {
        if (first + j < n) out[first + j] = (O)run;
        run += (u64)v[j];
    }
}

// small inputs (<= SCAN_SMALL_MAX elements): one CTA does the whole scan -- one launch instead of three
#define SCAN_SMALL_MAX 32768u
template <typename T, typename O>
__global__ void __launch_bounds__(SCAN1_THREADS) scan_small_kernel(const T* in, O* out, u32 n, ull* total_out) {
    __shared__ u64 sm[SCAN1_THREADS / 32 + 1];
    const u32 per = (n + SCAN1_THREADS - 1) / SCAN1_THREADS;           // <= 32 consecutive elements per thread
    const u32 t0 = min(n, threadIdx.x * per), t1 = min(n, t0 + per);
    T v[32];
    u64 acc = 0;
#pragma unroll
    for (u32 j = 0; j < 32; ++j) {
        v[j] = (j < per && t0 + j < t1) ? in[t0 + j] : (T)0;
        acc += (u64)v[j];
    }
    u64 total;
    u64 run = block_exclusive_sum64<SCAN1_THREADS / 32>(acc, sm, &total);
#pragma unroll
    for (u32 j = 0; j < 32; ++j) {
        if (j < per && t0 + j < t1) out[t0 + j] = (O)run;
        run += (u64)v[j];
    }
    if (threadIdx.x == 0 && total_out) *total_out = total;
}

// ---- radix sort -----------------------------------------------------------------------------------
#define RS_THREADS 256
#define RS_WARPS (RS_THREADS / 32)
#define RS_ITEMS 16
#define RS_TILE (RS_THREADS * RS_ITEMS)

// per-tile digit histogram, written digit-major: hist[d * ntiles + tile]
__global__ void __launch_bounds__(RS_THREADS) rs_hist_kernel(const u64* __restrict__ keys, u32 n, int shift, u32* __restrict__ hist, u32 ntiles) {
    __shared__ u32 h[256];
    h[threadIdx.x] = 0;
    BLOCK_SYNC();
    const u32 base = blockIdx.x * RS_TILE;
#pragma unroll
    for (int j = 0; j < RS_ITEMS; ++j) {
        const u32 i = base + j * RS_THREADS + threadIdx.x;
        if (i < n) atomicAdd(&h[(u32)(keys[i] >> shift) & 255u], 1u);
    }
    BLOCK_SYNC();
    hist[threadIdx.x * ntiles + blockIdx.x] = h[threadIdx.x];
}

struct NoVal {};

// stable scatter: warp w of the tile owns the contiguous segment [w*512, (w+1)*512) of the tile and
// walks it 32 keys at a time, so tile order is (warp, round, lane); ranks come from match.any peer
// groups plus per-warp running digit counters.
template <typename V, bool HAS_V>
__global__ void __launch_bounds__(RS_THREADS)
rs_scatter_kernel(const u64* __restrict__ kin, u64* __restrict__ kout, const V* __restrict__ vin, V* __restrict__ vout,
                  u32 n, int shift, const u32* __restrict__ offs /*scanned hist*/, u32 ntiles) {
    __shared__ u32 wcnt[RS_WARPS][256];
    __shared__ u32 gbase[256];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < RS_WARPS * 256; i += RS_THREADS) (&wcnt[0][0])[i] = 0;
    BLOCK_SYNC();
    const u32 seg = blockIdx.x * RS_TILE + warp * (32 * RS_ITEMS);
    u64 key[RS_ITEMS];
    u32 rank[RS_ITEMS];
    __syncwarp();
#pragma unroll
    for (int r = 0; r < RS_ITEMS; ++r) {
        const u32 i = seg + r * 32 + lane;
        const bool ok = i < n;
        key[r] = ok ? kin[i] : 0ull;
        const u32 d = ok ? ((u32)(key[r] >> shift) & 255u) : (256u + lane);
        const u32 peers = __match_any_sync(0xffffffffu, d);
        const u32 leader = __ffs(peers) - 1;
        u32 old = 0;
        if (ok && lane == (int)leader) { old = wcnt[warp][d]; wcnt[warp][d] = old + __popc(peers); }
        old = __shfl_sync(0xffffffffu, old, leader);
        rank[r] = old + __popc(peers & ((1u << lane) - 1u));
        __syncwarp();
    }
    BLOCK_SYNC();
    {   // exclusive scan over warps for digit = threadIdx.x; fold in the tile's global base
        const u32 d = threadIdx.x;
        u32 run = 0;
#pragma unroll
        for (int w2 = 0; w2 < RS_WARPS; ++w2) { const u32 c = wcnt[w2][d]; wcnt[w2][d] = run; run += c; }
        gbase[d] = offs[d * ntiles + blockIdx.x];
    }
    BLOCK_SYNC();
#pragma unroll
    for (int r = 0; r < RS_ITEMS; ++r) {
        const u32 i = seg + r * 32 + lane;
        if (i < n) {
            const u32 d = (u32)(key[r] >> shift) & 255u;
            const u32 pos = gbase[d] + wcnt[warp][d] + rank[r];
            kout[pos] = key[r];
            if (HAS_V) vout[pos] = vin[i];
        }
    }
}

// ---- run-length encode with threshold over a sorted sequence (unit weights) -----------------------
// Acc::eq(i, j): items i and j are equal.  A run is kept iff its length >= c; because the sequence is
// sorted, "run starting at i has length >= c"  <=>  eq(i, i + c - 1).
struct KeyEq {
    const u64* keys;
    __device__ __forceinline__ bool eq(u64 i, u64 j) const { return keys[i] == keys[j]; }
};

// window equality for the wide path: item i is the k-byte window src[pos[idx[i]] ..]
struct WindowEq {
    const u8* src;
    const u64* pos;
    const u32* idx;
    int k;
    __device__ __forceinline__ bool eq(u64 i, u64 j) const {
        const u8* a = src + pos[idx[i]];
        const u8* b = src + pos[idx[j]];
        for (int t = 0; t < k; ++t) if (a[t] != b[t]) return false;
        return true;
    }
};

#define RLE_THREADS 256
template <class Acc>
__device__ __forceinline__ bool rle_survivor_head(const Acc& acc, u64 i, u64 m, u64 c) {
    if (i + c - 1 >= m) return false;
    if (i > 0 && acc.eq(i - 1, i)) return false;
    return c <= 1 || acc.eq(i, i + c - 1);
}

template <class Acc>
__global__ void __launch_bounds__(RLE_THREADS) rle_count_kernel(Acc acc, u64 m, u64 c, u32* __restrict__ tile_cnt) {
    const u64 i = (u64)blockIdx.x * RLE_THREADS + threadIdx.x;
    const bool s = i < m && rle_survivor_head(acc, i, m, c);
    const u32 total = block_count(s);
    if (threadIdx.x == 0) tile_cnt[blockIdx.x] = total;
}

// writes, for survivor number u, start[u] = index of the run head and count[u] = run length
template <class Acc>
__global__ void __launch_bounds__(RLE_THREADS)
rle_write_kernel(Acc acc, u64 m, u64 c, const u64* __restrict__ tile_off, u64* __restrict__ start, u64* __restrict__ count) {
    __shared__ u32 sm[RLE_THREADS / 32 + 1];
    const u64 i = (u64)blockIdx.x * RLE_THREADS + threadIdx.x;
    const bool s = i < m && rle_survivor_head(acc, i, m, c);
    const u32 off = block_exclusive_scan<OpAdd, RLE_THREADS / 32>(s ? 1u : 0u, sm, nullptr);
    if (s) {
        // exponential + binary search for the end of the run (first j > i with !eq(i, j))
        u64 lo = i + (c > 1 ? c - 1 : 0);       // known equal
        u64 step = 1, hi = lo + 1;
        while (hi < m && acc.eq(i, hi)) { lo = hi; step <<= 1; hi = lo + step; }
        if (hi > m) hi = m;
        // invariant: eq(i, lo); (hi == m or !eq(i, hi)) -- find first unequal in (lo, hi]
        while (hi - lo > 1) {
            const u64 mid = lo + (hi - lo) / 2;
            if (acc.eq(i, mid)) lo = mid; else hi = mid;
        }
        const u64 u = tile_off[blockIdx.x] + off;
        start[u] = i;
        count[u] = hi - i;
    }
}

// ---- reduce-by-key with weights over sorted (key, weight) pairs -----------------------------------
// head flags -> segment starts; weight sums from an exclusive prefix of the weights.
template <class Acc>
__global__ void __launch_bounds__(RLE_THREADS) seg_count_kernel(Acc acc, u64 m, u32* __restrict__ tile_cnt) {
    const u64 i = (u64)blockIdx.x * RLE_THREADS + threadIdx.x;
    const bool h = i < m && (i == 0 || !acc.eq(i - 1, i));
    const u32 total = block_count(h);
    if (threadIdx.x == 0) tile_cnt[blockIdx.x] = total;
}
template <class Acc>
__global__ void __launch_bounds__(RLE_THREADS)
seg_write_kernel(Acc acc, u64 m, const u64* __restrict__ tile_off, u64* __restrict__ seg_start) {
    __shared__ u32 sm[RLE_THREADS / 32 + 1];
    const u64 i = (u64)blockIdx.x * RLE_THREADS + threadIdx.x;
    const bool h = i < m && (i == 0 || !acc.eq(i - 1, i));
    const u32 off = block_exclusive_scan<OpAdd, RLE_THREADS / 32>(h ? 1u : 0u, sm, nullptr);
    if (h) seg_start[tile_off[blockIdx.x] + off] = i;
}
// seg_sum[s] = wprefix[end] - wprefix[start]; flag survivors
__global__ void __launch_bounds__(RLE_THREADS)
seg_sum_count_kernel(const u64* __restrict__ seg_start, u64 nseg, u64 m, const u64* __restrict__ wprefix, u64 wtotal,
                     u64 c, u64* __restrict__ seg_sum, u32* __restrict__ tile_cnt) {
    const u64 s = (u64)blockIdx.x * RLE_THREADS + threadIdx.x;
    bool keep = false;
    if (s < nseg) {
        const u64 a = seg_start[s];
        const u64 hi = (s + 1 < nseg) ? wprefix[seg_start[s + 1]] : wtotal;
        const u64 sum = hi - wprefix[a];
        seg_sum[s] = sum;
        keep = sum >= c;
    }
    const u32 total = block_count(keep);
    if (threadIdx.x == 0) tile_cnt[blockIdx.x] = total;
}
__global__ void __launch_bounds__(RLE_THREADS)
seg_compact_kernel(const u64* __restrict__ seg_start, const u64* __restrict__ seg_sum, u64 nseg, u64 c,
                   const u64* __restrict__ tile_off, u64* __restrict__ start, u64* __restrict__ count) {
    __shared__ u32 sm[RLE_THREADS / 32 + 1];
    const u64 s = (u64)blockIdx.x * RLE_THREADS + threadIdx.x;
    const bool keep = s < nseg && seg_sum[s] >= c;
    const u32 off = block_exclusive_scan<OpAdd, RLE_THREADS / 32>(keep ? 1u : 0u, sm, nullptr);
    if (keep) {
        const u64 u = tile_off[blockIdx.x] + off;
        start[u] = seg_start[s];
        count[u] = seg_sum[s];
    }
}

// small gathers
// first and last key of every part (ends[2i], ends[2i+1]); parts are non-empty
__global__ void part_ends_kernel(const u64* const* __restrict__ ptrs, const u64* __restrict__ ns, u32 np, u64* __restrict__ ends) {
    const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < np) { ends[2 * i] = ptrs[i][0]; ends[2 * i + 1] = ptrs[i][ns[i] - 1]; }
}
__global__ void gather_u64_kernel(const u64* __restrict__ src, const u64* __restrict__ idx, u64 n, u64* __restrict__ dst) {
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = src[idx[i]];
}
__global__ void gather_u64_by_u32_kernel(const u64* __restrict__ src, const u32* __restrict__ idx, u64 n, u64* __restrict__ dst) {
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = src[idx[i]];
}
__global__ void iota_u32_kernel(u32* dst, u64 n) {
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = (u32)i;
}
__global__ void fill_u64_kernel(u64* dst, u64 n, u64 v) {
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = v;
}

// out[i] = number of keys (sorted ascending) smaller than q[i]
__global__ void lower_bound_kernel(const u64* __restrict__ keys, u64 n, const u64* __restrict__ q, u64 m, u64* __restrict__ out) {
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    const u64 x = q[i];
    u64 lo = 0, hi = n;
    while (lo < hi) {
        const u64 mid = (lo + hi) >> 1;
        if (keys[mid] < x) lo = mid + 1; else hi = mid;
    }
    out[i] = lo;
}

// concatenate ranges [src_off[r], src_off[r] + len) of src into dst at dst_off[r] (one CTA per 2048 output items)
__global__ void gather_ranges_kernel(const u64* __restrict__ src, const u64* __restrict__ src_off, const u64* __restrict__ dst_off,
                                     u32 nranges, u64 total, u64* __restrict__ dst) {
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (u64)gridDim.x * blockDim.x) {
        u32 lo = 0, hi = nranges;                              // last range with dst_off <= i
        while (hi - lo > 1) { const u32 mid = (lo + hi) >> 1; if (dst_off[mid] <= i) lo = mid; else hi = mid; }
        dst[i] = src[src_off[lo] + (i - dst_off[lo])];
    }
}
