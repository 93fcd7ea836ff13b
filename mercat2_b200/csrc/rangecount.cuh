// K3 -- large-k counting by ORDER-PRESERVING key-range partitioning + shared-memory tables.
//
// Replaces  kmerlist[k] += 1  over ~10^8..10^10 64-bit keys per chunk (lib/mercat2_kmers.py:56-60), the per-file
// count >= min_count  filter (:73-78) and -- because every level of the partition keeps key order -- the
// sorted(kmers.items())  of bin/mercat2.py:132: a chunk's surviving rows leave the counting kernel already sorted, so
// nothing is sorted afterwards.
//
// Sized from measurements on B200 (tools/microbench.cu, profiles/r01_microbench.txt): random shared-memory atomics
// run at ~2 T/s chip-wide and coalesced traffic at ~6.2 TB/s, while random 8-byte global stores reach only
// ~25-50 G/s.  So keys are partitioned with COALESCED writes in two levels (nb1 <= 384 buckets, then 128 sub-buckets
// each) until a sub-bucket (~3.5 k keys) fits an open-addressing table in shared memory; one CTA counts one
// sub-bucket entirely on chip, orders its survivors by a counting sort on the next key bits and writes them at the
// sub-bucket's own slot.  Chunks too large for two levels get a level-0 partition first (host_count.inl).
//
// The partition function works on the 32-bit PREFIX p of the left-aligned key (for 2-bit nucleotide keys: the first
// 16 symbols) and is monotone in the key:
//     b1 = lut[(p - base) >> sh]                 RP_LUT entries, built from a sampled prefix histogram so that
//                                                level-1 buckets hold equal shares of the sample (rp_plan_kernel)
//     b2 = ((p - l1[b1].x) * l1[b1].y) >> 32     linear inside the bucket: .y = 2^39 / (prefixes the bucket spans), so
//                                                the 128 sub-buckets tile the bucket's range exactly
// Offsets are exact (a histogram pass over all keys), so skew can never overflow memory; it only makes sub-buckets
// uneven.  A sub-bucket whose repeated keys do not fit the table is listed and redone by the sort path.
//
//   rp_sample_* prefix histogram of a strided sample               rp_plan   LUT + level-1 descriptors (one CTA)
//   *_hist      exact histogram over all nb1*128 sub-buckets       hc_scan   offsets, cursors, tile map (one CTA)
//   *_scatter1  keys grouped by level-1 bucket (8 B/key written)   hc_scatter2  grouped by sub-bucket (8 + 8 B/key)
//   rc_count    per sub-bucket: count in shared memory, filter, ordered emit (8 B/key read)
//   rc_offsets / rc_gather  survivor slots -> one dense sorted array
#pragma once
#include "common.cuh"
#include "extract.cuh"

#define HC_EMPTY 0xFFFFFFFFFFFFFFFFull
#define HC_NB2_LOG2 7
#define HC_NB2 (1u << HC_NB2_LOG2)
#define HC_MAX_NB1 384u
#define HC_TILE 4096u
#define HC_STAGE_SLOTS (HC_TILE + 2)                    // one tile + a spare slot for invalid keys (kept 16-byte aligned)
#define HC_SDST_OFFSET ((size_t)HC_STAGE_SLOTS * 8)
#define HC_SCATTER_SMEM16 (((size_t)HC_STAGE_SLOTS * 10 + 15) & ~(size_t)15) // staged keys (8 B) + 16-bit digits, padded to 16 B
#define HC_SCATTER_SMEM_LUT (HC_SCATTER_SMEM16 + (size_t)RP_LUT * 2)          // ... + the level-1 LUT behind them

#define RP_LUT_LOG2 13
#define RP_LUT (1u << RP_LUT_LOG2)

// One level of the partition as the kernels see it (all pointers: device memory).
struct RpView {
    const u16* lut;        // RP_LUT entries
    const uint2* l1;       // nb1 entries: .x = first prefix of the bucket, .y = scale of its linear sub-buckets (2^39 / span)
    u32 base, sh, nb1;
    u32 down, up;          // prefix of a right-aligned key of kb bits: (u32)(key >> down) << up   (kb >= 32: down = kb - 32)
    // When the sampled prefixes are spread evenly enough, rp_plan_kernel fills the LUT with the closed form
    // lut[i] = (i * mul) >> RP_MUL_SHIFT and sets *linear: the kernels then compute the bucket instead of looking it up
    // (one shared-memory load less per key in kernels that are bound by shared-memory traffic).
    const u32* linear;
    u32 mul;
};
#define RP_MUL_SHIFT 20
__device__ __forceinline__ u32 rp_prefix(u64 key, u32 down, u32 up) { return (u32)(key >> down) << up; }
__device__ __forceinline__ u32 rp_lut_index(u32 p, u32 base, u32 sh) { return min((p - base) >> sh, RP_LUT - 1u); }

// Shared-memory atomics issued from inside divergent probe loops go through inline PTX: the compiler otherwise
// rewrites atomicAdd(addr, 1) into a warp-aggregated VOTE + leader ATOMS + SHFL sequence.
__device__ __forceinline__ u32 smem_atom_inc(u32* addr) {
    u32 old;
    asm volatile("atom.shared.add.u32 %0, [%1], 1;" : "=r"(old) : "r"((u32)__cvta_generic_to_shared(addr)) : "memory");
    return old;
}
__device__ __forceinline__ u32 smem_atom_add(u32* addr, u32 v) {
    u32 old;
    asm volatile("atom.shared.add.u32 %0, [%1], %2;" : "=r"(old) : "r"((u32)__cvta_generic_to_shared(addr)), "r"(v) : "memory");
    return old;
}
__device__ __forceinline__ void smem_red_inc(u32* addr) {
    asm volatile("red.shared.add.u32 [%0], 1;" :: "r"((u32)__cvta_generic_to_shared(addr)) : "memory");
}
__device__ __forceinline__ u32 smem_atom_inc_if(u32 addr32, u32 pred) {
    u32 old = 0;
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %2, 0;\n\t@q atom.shared.add.u32 %0, [%1], 1;\n\t}" : "+r"(old) : "r"(addr32), "r"(pred) : "memory");
    return old;
}
__device__ __forceinline__ void smem_red_inc_if(u32 addr32, u32 pred) {
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %1, 0;\n\t@q red.shared.add.u32 [%0], 1;\n\t}" :: "r"(addr32), "r"(pred) : "memory");
}

// The partition tables in shared memory (hist and scatter kernels load them once per CTA).
struct RpShared {
    u16 lut[RP_LUT];
    uint2 l1[HC_MAX_NB1];
};
__device__ __forceinline__ void rp_load_shared(const RpView& r, RpShared& s, bool with_l1) {
    const uint4* src = reinterpret_cast<const uint4*>(r.lut);
    uint4* dst = reinterpret_cast<uint4*>(s.lut);
    for (u32 i = threadIdx.x; i < RP_LUT / 8; i += blockDim.x) dst[i] = src[i];
    if (with_l1)
        for (u32 i = threadIdx.x; i < r.nb1; i += blockDim.x) s.l1[i] = r.l1[i];
}
__device__ __forceinline__ u32 rp_b1_linear(const RpView& r, u32 p) { return min((rp_lut_index(p, r.base, r.sh) * r.mul) >> RP_MUL_SHIFT, r.nb1 - 1u); }
__device__ __forceinline__ u32 rp_b1(const RpShared& s, const RpView& r, u32 p) { return s.lut[rp_lut_index(p, r.base, r.sh)]; }
// (< 128 by construction of .y for every prefix of the bucket; the clamp only guards a caller that passed wrong bounds)
__device__ __forceinline__ u32 rp_b2(uint2 d, u32 p) { return min(__umulhi(p - d.x, d.y), HC_NB2 - 1u); }
// ordering digit inside a sub-bucket: the next RC_FINE_LOG2 bits of the same product (monotone in p within the sub-bucket)
__device__ __forceinline__ u32 rp_fine(uint2 d, u32 p, u32 bits) { return (u32)(((u64)(p - d.x) * d.y) >> (32 - bits)) & ((1u << bits) - 1u); }
__device__ __forceinline__ u32 rp_sub(const RpShared& s, const RpView& r, u32 p) {
    const u32 b1 = rp_b1(s, r, p);
    return b1 * HC_NB2 + rp_b2(s.l1[b1], p);
}

// ---- rp_plan: sampled prefix histogram -> LUT + level-1 descriptors (one CTA) ----------------------------------------
// shist[i]: sampled keys with LUT index i; nidx: number of LUT indices the range covers (<= RP_LUT).  Bucket of index i
// = floor(nb1 * (sample mass before i) / total): monotone, and no bucket exceeds its share by more than one index.
// With an empty sample the indices are spread evenly.  A bucket id that no index maps to stays empty (zero keys).
__global__ void __launch_bounds__(1024)
rp_plan_kernel(const u32* __restrict__ shist, u32 nb1, u32 base, u32 sh, u32 nidx, u32 mul, u16* __restrict__ lut, uint2* __restrict__ l1,
               u32* __restrict__ linear) {
    __shared__ u64 sm[1024 / 32 + 1];
    __shared__ u32 first[HC_MAX_NB1 + 1];
    __shared__ u32 mass[HC_MAX_NB1];
    __shared__ u32 s_max;
    constexpr u32 PER = RP_LUT / 1024;
    u32 h[PER];
    u64 acc = 0;
#pragma unroll
    for (u32 j = 0; j < PER; ++j) { h[j] = shist[threadIdx.x * PER + j]; acc += h[j]; }
    for (u32 i = threadIdx.x; i <= nb1; i += 1024) first[i] = 0xFFFFFFFFu;
    for (u32 i = threadIdx.x; i < nb1; i += 1024) mass[i] = 0;
    if (threadIdx.x == 0) s_max = 0;
    u64 total;
    u64 run = block_exclusive_sum64<32>(acc, sm, &total);
    // sample mass per bucket under the closed-form (evenly spaced) assignment: good enough if no bucket exceeds its share by 12 %
#pragma unroll
    for (u32 j = 0; j < PER; ++j) {
        const u32 idx = threadIdx.x * PER + j;
        if (idx < nidx && h[j]) atomicAdd(&mass[min(nb1 - 1, (idx * mul) >> RP_MUL_SHIFT)], h[j]);
    }
    BLOCK_SYNC();
    for (u32 i = threadIdx.x; i < nb1; i += 1024) atomicMax(&s_max, mass[i]);
    BLOCK_SYNC();
    const bool lin = total == 0 || (u64)s_max * nb1 * 100 <= total * 112;
    if (threadIdx.x == 0) *linear = lin ? 1u : 0u;
#pragma unroll
    for (u32 j = 0; j < PER; ++j) {
        const u32 idx = threadIdx.x * PER + j;
        u32 b;
        if (lin) b = min(nb1 - 1, (min(idx, RP_LUT - 1u) * mul) >> RP_MUL_SHIFT);
        else if (idx >= nidx) b = nb1 - 1;
        else b = (u32)min((u64)(nb1 - 1), (run * nb1) / total);
        lut[idx] = (u16)b;
        if (idx < nidx) atomicMin(&first[b], idx);
        run += h[j];
    }
    BLOCK_SYNC();
    if (threadIdx.x == 0) {                         // empty buckets inherit the start of the next one (span 0)
        first[nb1] = nidx;
        for (int b = (int)nb1 - 1; b >= 0; --b)
            if (first[b] == 0xFFFFFFFFu) first[b] = first[b + 1];
    }
    BLOCK_SYNC();
    for (u32 b = threadIdx.x; b < nb1; b += 1024) {
        const u64 lo = (u64)base + ((u64)first[b] << sh);
        const u64 span = (u64)(first[b + 1] - first[b]) << sh;             // prefix values the bucket covers
        // x in [0, span)  ->  (x * scale) >> 32 in [0, 128): scale = floor(2^39 / span)   (tiny spans: one prefix per sub-bucket)
        const u64 scale = span >= HC_NB2 ? min((u64)0xFFFFFFFFull, (u64)((1ull << (32 + HC_NB2_LOG2)) / span)) : 0xFFFFFFFFull;
        l1[b] = make_uint2((u32)min(lo, (u64)0xFFFFFFFFull), (u32)scale);
    }
}

// ---- key-array sources: sample, histogram -----------------------------------------------------------------------------
#define HK_HIST_THREADS 1024
__global__ void __launch_bounds__(HK_HIST_THREADS)
rp_sample_keys_kernel(const u64* __restrict__ keys, u64 n, u64 stride, u32 down, u32 up, u32 base, u32 sh, u32* __restrict__ shist) {
    __shared__ u32 hist[RP_LUT];
    for (u32 i = threadIdx.x; i < RP_LUT; i += HK_HIST_THREADS) hist[i] = 0;
    BLOCK_SYNC();
    for (u64 i = ((u64)blockIdx.x * HK_HIST_THREADS + threadIdx.x) * stride; i < n; i += (u64)gridDim.x * HK_HIST_THREADS * stride)
        atomicAdd(&hist[rp_lut_index(rp_prefix(keys[i], down, up), base, sh)], 1u);
    BLOCK_SYNC();
    for (u32 b = threadIdx.x; b < RP_LUT; b += HK_HIST_THREADS) {
        const u32 c = hist[b];
        if (c) atomicAdd(&shist[b], c);
    }
}

// FINE: histogram over all nb1 * 128 sub-buckets; otherwise over the nb1 level-1 buckets (level-0 partition)
template <bool FINE>
__global__ void __launch_bounds__(HK_HIST_THREADS)
hk_hist_kernel(const u64* __restrict__ keys, u64 n, RpView r, u32 nb, u32* __restrict__ ghist) {
    extern __shared__ __align__(16) u8 dyn[];
    RpShared& rs = *reinterpret_cast<RpShared*>(dyn);
    u32* hist = reinterpret_cast<u32*>(dyn + sizeof(RpShared));
    rp_load_shared(r, rs, FINE);
    for (u32 i = threadIdx.x; i < nb; i += HK_HIST_THREADS) hist[i] = 0;
    BLOCK_SYNC();
    const bool lin = *r.linear != 0;
    auto bucket = [&](u32 p) {
        const u32 b1 = lin ? rp_b1_linear(r, p) : rp_b1(rs, r, p);
        return FINE ? b1 * HC_NB2 + rp_b2(rs.l1[b1], p) : b1;
    };
    // four independent loads in flight per thread (one per iteration left the kernel latency-bound: 0.24 ms per 65 M keys)
    const u64 stride = (u64)gridDim.x * HK_HIST_THREADS;
    u64 i = (u64)blockIdx.x * HK_HIST_THREADS + threadIdx.x;
    for (; i + 3 * stride < n; i += 4 * stride) {
        u64 kk[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) kk[j] = keys[i + j * stride];
#pragma unroll
        for (int j = 0; j < 4; ++j) smem_red_inc(&hist[bucket(rp_prefix(kk[j], r.down, r.up))]);
    }
    for (; i < n; i += stride) smem_red_inc(&hist[bucket(rp_prefix(keys[i], r.down, r.up))]);
    BLOCK_SYNC();
    for (u32 b = threadIdx.x; b < nb; b += HK_HIST_THREADS) {
        const u32 c = hist[b];
        if (c) atomicAdd(&ghist[b], c);
    }
}

// ---- byte-symbol sources (general lane): sample, histogram ---------------------------------------------------------------
template <int ENC>
__global__ void __launch_bounds__(EX_THREADS)
rp_sample_sym_kernel(SymView v, int k, u64 tile_stride, u32 down, u32 up, u32 base, u32 sh, u32* __restrict__ shist) {
    __shared__ u32 hist[RP_LUT];
    __shared__ u64 s_code[EX_THREADS + EX_HALO];
    __shared__ u32 s_meta[EX_THREADS + EX_HALO];
    for (u32 i = threadIdx.x; i < RP_LUT; i += EX_THREADS) hist[i] = 0;
    BLOCK_SYNC();
    const int kb = k * EncTraits<ENC>::BITS;
    const u64 mask = kb >= 64 ? ~0ull : ((1ull << kb) - 1);
    const u64 ntiles = (v.n + EX_TILE - 1) / EX_TILE;
    for (u64 tile = (u64)blockIdx.x * tile_stride; tile < ntiles; tile += (u64)gridDim.x * tile_stride) {
        const u64 tile_start = tile * EX_TILE;
        TileCtx<ENC> ctx;
        tile_begin<ENC>(v, tile_start, k, s_code, s_meta, ctx);
        const u64 first = tile_start + 16ull * threadIdx.x;
        SampleTag st;
        st.init(v, first);
        tile_walk<ENC>(ctx, k, [&](int i, u64 code, bool fast) {
            if (fast && first + i < v.n) atomicAdd(&hist[rp_lut_index(rp_prefix((code & mask) | st.tag(v, first + i), down, up), base, sh)], 1u);
        });
        BLOCK_SYNC();
    }
    BLOCK_SYNC();
    for (u32 b = threadIdx.x; b < RP_LUT; b += EX_THREADS) {
        const u32 c = hist[b];
        if (c) atomicAdd(&shist[b], c);
    }
}

template <int ENC>
__global__ void __launch_bounds__(EX_THREADS)
hc_hist_kernel(SymView v, u64 s0, u64 s1, int k, RpView r, u32 nb, u32* __restrict__ ghist) {
    extern __shared__ __align__(16) u8 dyn[];
    RpShared& rs = *reinterpret_cast<RpShared*>(dyn);
    u32* hist = reinterpret_cast<u32*>(dyn + sizeof(RpShared));
    __shared__ u64 s_code[EX_THREADS + EX_HALO];
    __shared__ u32 s_meta[EX_THREADS + EX_HALO];
    rp_load_shared(r, rs, true);
    for (u32 i = threadIdx.x; i < nb; i += EX_THREADS) hist[i] = 0;
    BLOCK_SYNC();
    const int kb = k * EncTraits<ENC>::BITS;
    const u64 mask = kb >= 64 ? ~0ull : ((1ull << kb) - 1);
    const u64 ntiles = (s1 - s0 + EX_TILE - 1) / EX_TILE;
    for (u64 tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const u64 tile_start = s0 + tile * EX_TILE;
        TileCtx<ENC> ctx;
        tile_begin<ENC>(v, tile_start, k, s_code, s_meta, ctx);
        const u64 first = tile_start + 16ull * threadIdx.x;
        SampleTag st;
        st.init(v, first);
        tile_walk<ENC>(ctx, k, [&](int i, u64 code, bool fast) {
            if (fast && first + i < s1) atomicAdd(&hist[rp_sub(rs, r, rp_prefix((code & mask) | st.tag(v, first + i), r.down, r.up))], 1u);
        });
        BLOCK_SYNC();
    }
    BLOCK_SYNC();
    for (u32 b = threadIdx.x; b < nb; b += EX_THREADS) {
        const u32 n = hist[b];
        if (n) atomicAdd(&ghist[b], n);
    }
}

// ---- hc_scan (one CTA) ------------------------------------------------------------------------------------
// sub_base[b] (b <= nb): start of sub-bucket b in the level-2 array (== level-1 array position of its
// bucket when b % nb2 == 0); cur1[b1] / cur2[b]: running cursors for the two scatters;
// tile_pref[b1] (b1 <= nb1): first level-2-scatter tile of bucket b1.
__global__ void __launch_bounds__(1024)
hc_scan_kernel(const u32* __restrict__ ghist, u32 nb, u32 nb1, u32 nb2, u32* __restrict__ sub_base, u32* __restrict__ cur1,
               u32* __restrict__ cur2, u32* __restrict__ tile_pref, ull* __restrict__ total_out) {
    extern __shared__ __align__(16) u8 dyn_scan[];                 // nb words: the histogram, then its exclusive prefix
    u32* h = reinterpret_cast<u32*>(dyn_scan);
    __shared__ u64 sm[1024 / 32 + 1];
    __shared__ u32 s_l1[HC_MAX_NB1 + 1];
    for (u32 i = threadIdx.x; i < nb; i += 1024) h[i] = ghist[i];   // coalesced; the serial part below runs on shared memory
    BLOCK_SYNC();
    const u32 per = (nb + 1023) / 1024;
    const u32 t0 = min(nb, threadIdx.x * per), t1 = min(nb, t0 + per);
    u64 acc = 0;
    for (u32 t = t0; t < t1; ++t) acc += h[t];
    u64 total;
    u64 base = block_exclusive_sum64<32>(acc, sm, &total);
    for (u32 t = t0; t < t1; ++t) {
        const u32 c = h[t];
        h[t] = (u32)base;
        if (t % nb2 == 0) s_l1[t / nb2] = (u32)base;
        base += c;
    }
    if (threadIdx.x == 0) { sub_base[nb] = (u32)total; s_l1[nb1] = (u32)total; *total_out = total; }
    BLOCK_SYNC();
    for (u32 i = threadIdx.x; i < nb; i += 1024) { const u32 b = h[i]; sub_base[i] = b; cur2[i] = b; }
    for (u32 i = threadIdx.x; i < nb1; i += 1024) cur1[i] = s_l1[i];
    // tiles per level-1 bucket -> exclusive prefix
    u64 tl = 0;
    if (threadIdx.x < nb1) tl = (s_l1[threadIdx.x + 1] - s_l1[threadIdx.x] + HC_TILE - 1) / HC_TILE;
    u64 ttotal;
    const u64 tp = block_exclusive_sum64<32>(tl, sm, &ttotal);
    if (threadIdx.x < nb1) tile_pref[threadIdx.x] = (u32)tp;
    if (threadIdx.x == 0) tile_pref[nb1] = (u32)ttotal;
}

// ---- shared helper: group up to 16 keys per thread by a small digit and write coalesced runs -----------------
// cnt / loff / gbase: nd words each; stage / sdst: one slot per key of the tile.  Every thread calls; bit i of
// `valid` says key i exists.  cursors[d] is advanced atomically by the tile's count for digit d.  Keys are
// re-ordered through shared memory so that consecutive threads store consecutive addresses of one digit's run.
// `mine(i)` yields key i, `dig(i)` its digit (both may recompute instead of holding registers).
// `base64` (may be NULL): 64-bit start of every digit's region; the cursors are then relative to it (level-0 groups of
// chunks with more than 2^32 windows).
// FULL: all 16 keys of every thread are valid (interior tiles): no per-key tests.  Otherwise the rank atomic is
// predicated in PTX and invalid keys are staged into a spare slot, so the unrolled per-key code stays free of
// branches (every branch re-derives the shared-memory window base and brackets itself with BSSY/BSYNC).
template <bool FULL, class KeyFn, class DigitFn>
__device__ __forceinline__ void hc_group_and_write(KeyFn mine, DigitFn dig, u32 valid, u32 nd, u64* stage, u16* sdig,
                                                   u32* cnt, u32* loff, u32* gbase, u32* sm, u32* __restrict__ cursors,
                                                   u64* __restrict__ out, const u64* __restrict__ base64) {
    u32 rd[16];                             // (rank within (tile, digit)) << 16 | digit
    const u32 cnt32 = (u32)__cvta_generic_to_shared(cnt);
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const u32 d = dig(i);
        const u32 r = FULL ? atomicAdd(&cnt[d], 1u) : smem_atom_inc_if(cnt32 + 4u * d, (valid >> i) & 1u);
        rd[i] = (r << 16) | d;
    }
    BLOCK_SYNC();
    const u32 per = (nd + EX_THREADS - 1) / EX_THREADS;
    u32 acc = 0;
    for (u32 j = 0; j < per; ++j) { const u32 d = threadIdx.x * per + j; if (d < nd) acc += cnt[d]; }
    u32 total;
    u32 run = block_exclusive_scan<OpAdd, EX_WARPS>(acc, sm, &total);
    for (u32 j = 0; j < per; ++j) {
        const u32 d = threadIdx.x * per + j;
        if (d < nd) {
            const u32 c = cnt[d];
            loff[d] = run;
            gbase[d] = c ? atomicAdd(&cursors[d], c) : 0u;
            run += c;
        }
    }
    BLOCK_SYNC();
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const u32 d = rd[i] & 0xFFFFu;
        const u32 r = rd[i] >> 16;
        u32 pos = loff[d] + r;
        if (!FULL) pos = ((valid >> i) & 1u) ? pos : HC_TILE;
        stage[pos] = mine(i);
        sdig[pos] = (u16)d;
    }
    BLOCK_SYNC();
    for (u32 i = threadIdx.x; i < total; i += EX_THREADS) {
        const u32 d = sdig[i];
        const u64 at = (base64 ? base64[d] : 0ull) + gbase[d] + (i - loff[d]);
        out[at] = stage[i];
    }
}

// ---- hc_scatter1: symbols -> level-1 groups ----------------------------------------------------------------
template <int ENC>
__global__ void __launch_bounds__(EX_THREADS)
hc_scatter1_kernel(SymView v, u64 s0, u64 s1, int k, RpView r, u32* __restrict__ cur1, u64* __restrict__ keys1) {
    __shared__ u64 s_code[EX_THREADS + EX_HALO];
    __shared__ u32 s_meta[EX_THREADS + EX_HALO];
    extern __shared__ __align__(16) u8 dyn_sc[];                    // HC_SCATTER_SMEM16 + lut
    u64* stage = reinterpret_cast<u64*>(dyn_sc);
    u16* sdig = reinterpret_cast<u16*>(dyn_sc + HC_SDST_OFFSET);
    u16* s_lut = reinterpret_cast<u16*>(dyn_sc + HC_SCATTER_SMEM16);
    __shared__ u32 cnt[HC_MAX_NB1], loff[HC_MAX_NB1], gbase[HC_MAX_NB1];
    __shared__ u32 sm[EX_WARPS + 1];
    for (u32 i = threadIdx.x; i < r.nb1; i += EX_THREADS) cnt[i] = 0;
    for (u32 i = threadIdx.x; i < RP_LUT / 8; i += EX_THREADS) reinterpret_cast<uint4*>(s_lut)[i] = reinterpret_cast<const uint4*>(r.lut)[i];
    const int kb = k * EncTraits<ENC>::BITS;
    const u64 mask = kb >= 64 ? ~0ull : ((1ull << kb) - 1);
    const u64 tile_start = s0 + (u64)blockIdx.x * EX_TILE;
    TileCtx<ENC> ctx;
    tile_begin<ENC>(v, tile_start, k, s_code, s_meta, ctx);        // contains the barrier that publishes cnt = 0 and the LUT
    const u64 first = tile_start + 16ull * threadIdx.x;
    u64 mine[16];
    u32 valid = 0;
    SampleTag st;
    st.init(v, first);
    tile_walk<ENC>(ctx, k, [&](int i, u64 code, bool fast) {
        mine[i] = code & mask;
        if (fast && first + i < s1) { valid |= 1u << i; mine[i] |= st.tag(v, first + i); }
    });
    auto key = [&](int i) { return mine[i]; };
    auto dig = [&](int i) { return ((valid >> i) & 1u) ? (u32)s_lut[rp_lut_index(rp_prefix(mine[i], r.down, r.up), r.base, r.sh)] : 0u; };
    hc_group_and_write<false>(key, dig, valid, r.nb1, stage, sdig, cnt, loff, gbase, sm, cur1, keys1, nullptr);
}

// ---- hk_scatter1: key array -> level-1 groups (or level-0 groups with base64) -----------------------------------------
// Persistent CTAs (three per SM) walk the tiles: the 16 KB LUT is loaded once per CTA -- per 4096-key tile it was half as
// many bytes again as the keys themselves -- and not at all when the plan is the closed form.
__global__ void __launch_bounds__(EX_THREADS, 3)
hk_scatter1_kernel(const u64* __restrict__ keys, u64 n, RpView r, u32* __restrict__ cur1, u64* __restrict__ keys1,
                   const u64* __restrict__ base64) {
    extern __shared__ __align__(16) u8 dyn_sc[];
    u64* stage = reinterpret_cast<u64*>(dyn_sc);
    u16* sdig = reinterpret_cast<u16*>(dyn_sc + HC_SDST_OFFSET);
    u16* s_lut = reinterpret_cast<u16*>(dyn_sc + HC_SCATTER_SMEM16);
    __shared__ u32 cnt[HC_MAX_NB1], loff[HC_MAX_NB1], gbase[HC_MAX_NB1];
    __shared__ u32 sm[EX_WARPS + 1];
    const bool lin = *r.linear != 0;
    if (!lin)
        for (u32 i = threadIdx.x; i < RP_LUT / 8; i += EX_THREADS) reinterpret_cast<uint4*>(s_lut)[i] = reinterpret_cast<const uint4*>(r.lut)[i];
    const u64 ntiles = (n + HC_TILE - 1) / HC_TILE;
    for (u64 tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        for (u32 i = threadIdx.x; i < r.nb1; i += EX_THREADS) cnt[i] = 0;
        const u64 base = tile * HC_TILE;
        u64 mine[16];
        u32 valid = 0;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const u64 i = base + (u64)j * EX_THREADS + threadIdx.x;
            mine[j] = 0;
            if (i < n) { mine[j] = keys[i]; valid |= 1u << j; }
        }
        BLOCK_SYNC();
        auto key = [&](int i) { return mine[i]; };
        auto dig = [&](int i) {
            const u32 p = rp_prefix(mine[i], r.down, r.up);
            return ((valid >> i) & 1u) ? (lin ? rp_b1_linear(r, p) : (u32)s_lut[rp_lut_index(p, r.base, r.sh)]) : 0u;
        };
        if (base + HC_TILE <= n) hc_group_and_write<true>(key, dig, valid, r.nb1, stage, sdig, cnt, loff, gbase, sm, cur1, keys1, base64);
        else hc_group_and_write<false>(key, dig, valid, r.nb1, stage, sdig, cnt, loff, gbase, sm, cur1, keys1, base64);
        BLOCK_SYNC();                                          // the staging area and the counters are re-used by the next tile
    }
}

// ---- hc_scatter2: level-1 groups -> sub-buckets --------------------------------------------------------------
__global__ void __launch_bounds__(EX_THREADS, 3)
hc_scatter2_kernel(const u64* __restrict__ keys1, const u32* __restrict__ sub_base, const u32* __restrict__ tile_pref,
                   u32 nb, u32 nb2, RpView r, u32* __restrict__ cur2, u64* __restrict__ keys2) {
    extern __shared__ __align__(16) u8 dyn_sc[];                    // HC_SCATTER_SMEM16 bytes
    u64* stage = reinterpret_cast<u64*>(dyn_sc);
    u16* sdig = reinterpret_cast<u16*>(dyn_sc + HC_SDST_OFFSET);
    __shared__ u32 cnt[HC_NB2], loff[HC_NB2], gbase[HC_NB2];
    __shared__ u32 sm[EX_WARPS + 1];
    __shared__ u32 s_b1;
    const u32 nb1 = r.nb1;
    if (blockIdx.x >= tile_pref[nb1]) return;
    if (threadIdx.x == 0) {                      // last b1 with tile_pref[b1] <= blockIdx.x
        u32 lo = 0, hi = nb1;
        while (hi - lo > 1) { const u32 mid = (lo + hi) / 2; if (tile_pref[mid] <= blockIdx.x) lo = mid; else hi = mid; }
        s_b1 = lo;
    }
    for (u32 i = threadIdx.x; i < nb2; i += EX_THREADS) cnt[i] = 0;
    BLOCK_SYNC();
    const u32 b1 = s_b1;
    const uint2 d1 = r.l1[b1];
    const u32 lo = sub_base[b1 * nb2], hi = sub_base[min(nb, (b1 + 1) * nb2)];
    const u32 t_in = blockIdx.x - tile_pref[b1];
    const u32 base = lo + t_in * HC_TILE;
    u64 mine[16];
    u32 valid = 0;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        const u32 i = base + j * EX_THREADS + threadIdx.x;
        mine[j] = 0;
        if (i < hi) { mine[j] = keys1[i]; valid |= 1u << j; }
    }
    auto key = [&](int i) { return mine[i]; };
    auto dig = [&](int i) { return ((valid >> i) & 1u) ? rp_b2(d1, rp_prefix(mine[i], r.down, r.up)) : 0u; };
    if (base + HC_TILE <= hi) hc_group_and_write<true>(key, dig, valid, nb2, stage, sdig, cnt, loff, gbase, sm, cur2 + b1 * nb2, keys2, nullptr);
    else hc_group_and_write<false>(key, dig, valid, nb2, stage, sdig, cnt, loff, gbase, sm, cur2 + b1 * nb2, keys2, nullptr);
}

// ---- rc_count: persistent CTAs, one sub-bucket at a time ---------------------------------------------------------------
// MODE 0 (min_count == 1, or data in which most keys repeat): every key is inserted into an exact open-addressing
// table (claim by CAS, bump by reduction).
// MODE 1 (min_count >= 2, mostly singletons -- a metagenome chunk): pass 1a sets one bit per key in a 2^18-bit bitmap;
// only a key that finds its bit already set (a repeat, or a ~1 % false positive) is queued and then given a table
// slot (pass 1b, one key per lane).  If some key was flagged >= min_count - 1 times the bucket may hold a survivor and
// pass 2 recounts every key that has a slot exactly; otherwise nothing in it can reach min_count.
// Keys are held in registers (8 per thread = one ROUND of 4096 keys, the next bucket's first round is prefetched while
// the current one is processed); larger buckets take further rounds from global memory.
// Ordered emit: the survivors (distinct keys with count >= min_count) are ordered by a counting sort on the bucket's
// own key range (RC_FINE bins, then a rank among the few keys of a bin) and written as 16-byte rows (key, count) at
// slot sub_base[b] / min_count of `out` -- a bucket of n keys has at most n / min_count survivors, so slots never
// collide.  rows[b] = number of rows written.
#define RC_THREADS 512
#define RC_SLOTS 4096u
#define RC_LIMIT 3072u
#define RC_CLAIM_CAP (RC_LIMIT + RC_THREADS)
#define RC_BM_WORDS 8192u                        // 2^18 bits
#define RC_PREFETCH 8
#define RC_ROUND (RC_PREFETCH * RC_THREADS)
#define RC_FLIST 1024u                           // flagged keys queued per round before the dense insert step
#define RC_FINE_LOG2 10
#define RC_FINE (1u << RC_FINE_LOG2)             // ordering bins
#define RC_STAGE_ROWS RC_CLAIM_CAP
// region A: bitmap (pass 1) | staged survivors (keys 8 B + counts 4 B); region B: flagged-key queue | ordering bins
#define RC_A_BYTES ((size_t)RC_STAGE_ROWS * 12 > (size_t)RC_BM_WORDS * 4 ? (size_t)RC_STAGE_ROWS * 12 : (size_t)RC_BM_WORDS * 4)
#define RC_B_BYTES ((size_t)RC_FLIST * 8)
#define RC_SMEM (RC_A_BYTES + (size_t)RC_SLOTS * 12 + (size_t)RC_CLAIM_CAP * 2 + RC_B_BYTES)
static_assert(RC_A_BYTES % 16 == 0 && (RC_CLAIM_CAP * 2) % 8 == 0, "shared-memory regions stay aligned");
static_assert(RC_FINE * 4 <= RC_B_BYTES, "ordering bins fit the queue region");
static_assert(RC_FINE == 2 * RC_THREADS, "the bin scan takes two bins per thread");

__device__ __forceinline__ u32 rc_hash(ull key) { return (u32)key * 0x9E3779B1u + (u32)(key >> 32) * 0x85EBCA77u; }
__device__ __forceinline__ u32 rc_slot(u32 h) { return (h * 0xC2B2AE3Du) >> (32 - 12); }
static_assert(RC_SLOTS == 4096u, "rc_slot yields 12 bits");

// scal: [0] all-ones-key occurrences (flagged ones in MODE 1 pass 1), [1] distinct, [2] overflow, [3] need pass 2,
//       [4] exact all-ones-key count (pass 2), [5] queue length, [6] survivors
// Insert / bump: returns the key's count after this occurrence.
__device__ __forceinline__ u32 rc_insert(ull key, u32 h, ull* tkeys, u32* tcnt, u16* claimed, u32* scal) {
    if (*(volatile u32*)&scal[2]) return 0u;                 // table already overflowed: the bucket is redone by sorting
    u32 p = rc_slot(h);
    while (true) {
        ull cur = tkeys[p];
        if (cur == HC_EMPTY) {
            cur = atomicCAS(&tkeys[p], HC_EMPTY, key);
            if (cur == HC_EMPTY) {
                const u32 d = smem_atom_inc(&scal[1]);
                if (d < RC_CLAIM_CAP) claimed[d] = (u16)p;
                if (d >= RC_LIMIT) scal[2] = 1;
                cur = key;
            }
        }
        if (cur == key) return smem_atom_inc(&tcnt[p]) + 1u;
        p = (p + 1) & (RC_SLOTS - 1);
    }
}
__device__ __noinline__ u32 rc_flagged(ull key, u32 h, ull* tkeys, u32* tcnt, u16* claimed, u32* scal) {
    if (key == HC_EMPTY) return smem_atom_inc(&scal[0]) + 1u;
    return rc_insert(key, h, tkeys, tcnt, claimed, scal);
}
__device__ __forceinline__ void rc_pass1(ull key, u32* bm, ull* tkeys, u32* tcnt, u16* claimed, ull* flist, u32* scal, u32 need_at) {
    const u32 h = rc_hash(key);
    const u32 bit = 1u << ((h >> 14) & 31u);
    const u32 old = atomicOr(&bm[h >> 19], bit);
    if (old & bit) {
        const u32 q = smem_atom_inc(&scal[5]);
        if (q < RC_FLIST) flist[q] = key;
        else if (rc_flagged(key, h, tkeys, tcnt, claimed, scal) >= need_at) scal[3] = 1;
    }
}
__device__ __forceinline__ void rc_pass2(ull key, const ull* tkeys, u32* tcnt, u32* scal) {
    if (key == HC_EMPTY) { smem_red_inc(&scal[4]); return; }
    u32 p = rc_slot(rc_hash(key));
    while (true) {
        const ull cur = tkeys[p];
        if (cur == key) { smem_red_inc(&tcnt[p]); return; }
        if (cur == HC_EMPTY) return;
        p = (p + 1) & (RC_SLOTS - 1);
    }
}

struct RcRow { ull key, count; };

// MODE 0 insert: no claim log, no distinct counter (the emit pass walks the whole table); a probe sequence that does not
// end within RC_PROBE_LIMIT slots means the table is (nearly) full: the bucket is flagged and redone by sorting.
#define RC_PROBE_LIMIT 512u
__device__ __forceinline__ void rc_insert0(ull key, ull* tkeys, u32* tcnt, u32* scal) {
    u32 p = rc_slot(rc_hash(key));
#pragma unroll 1
    for (u32 probes = 0; probes < RC_PROBE_LIMIT; ++probes) {
        ull cur = tkeys[p];
        if (cur == HC_EMPTY) {
            cur = atomicCAS(&tkeys[p], HC_EMPTY, key);
            if (cur == HC_EMPTY) cur = key;
        }
        if (cur == key) { smem_red_inc(&tcnt[p]); return; }
        p = (p + 1) & (RC_SLOTS - 1);
    }
    scal[2] = 1;
}

template <int MODE>
__global__ void __launch_bounds__(RC_THREADS, 2)
rc_count_kernel(const u64* __restrict__ keys2, const u32* __restrict__ sub_base, u32 nb, u32 c, RpView r,
                RcRow* __restrict__ out, u32* __restrict__ rows, u32* __restrict__ ovf_list, u32* __restrict__ ovf_n,
                ull* __restrict__ flagged_total /*MODE 1: keys that found their bitmap bit set (repeats + ~1 % false positives);
                                                  MODE 0: distinct keys held by the tables (overflowed sub-buckets excluded)*/) {
    extern __shared__ __align__(16) u8 dyn[];
    u32* bm = reinterpret_cast<u32*>(dyn);                                               // region A
    ull* skey = reinterpret_cast<ull*>(dyn);
    u32* scnt = reinterpret_cast<u32*>(dyn + (size_t)RC_STAGE_ROWS * 8);
    ull* tkeys = reinterpret_cast<ull*>(dyn + RC_A_BYTES);
    u32* tcnt = reinterpret_cast<u32*>(dyn + RC_A_BYTES + (size_t)RC_SLOTS * 8);
    u16* claimed = reinterpret_cast<u16*>(dyn + RC_A_BYTES + (size_t)RC_SLOTS * 12);
    ull* flist = reinterpret_cast<ull*>(dyn + RC_A_BYTES + (size_t)RC_SLOTS * 12 + (size_t)RC_CLAIM_CAP * 2);   // region B
    u32* bins = reinterpret_cast<u32*>(flist);
    __shared__ u32 s_scal[2][8];
    __shared__ u32 s_warp[RC_THREADS / 32 + 1];
    for (u32 i = threadIdx.x; i < RC_SLOTS; i += RC_THREADS) { tkeys[i] = HC_EMPTY; tcnt[i] = 0; }
    if (MODE == 1) for (u32 i = threadIdx.x; i < RC_BM_WORDS; i += RC_THREADS) bm[i] = 0;
    if (MODE == 0) for (u32 i = threadIdx.x; i < RC_FINE; i += RC_THREADS) bins[i] = 0;
    if (threadIdx.x < 16) (&s_scal[0][0])[threadIdx.x] = 0;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const u32 need_at = c - 1;                                 // flagged occurrences at which a key may reach min_count
    ull knext[RC_PREFETCH];
    u32 b = blockIdx.x;
    u32 lo_n = 0, n_n = 0;
    if (b < nb) {
        lo_n = sub_base[b];
        n_n = sub_base[b + 1] - lo_n;
#pragma unroll
        for (int j = 0; j < RC_PREFETCH; ++j) {
            const u32 i = j * RC_THREADS + threadIdx.x;
            knext[j] = i < n_n ? keys2[lo_n + i] : 0ull;
        }
    }
    BLOCK_SYNC();
    u32 par = 0;
    ull cta_flagged = 0;
    for (; b < nb; b += gridDim.x, par ^= 1u) {
        u32* scal = s_scal[par];
        const u32 n = n_n, lo = lo_n;
        const uint2 d1 = __ldg(&r.l1[b >> HC_NB2_LOG2]);           // (used by the emit pass: fetched here so that its latency is hidden)
        ull kcur[RC_PREFETCH];
        const u32 nrounds = (n + RC_ROUND - 1) / RC_ROUND;         // block-uniform; 1 at the default sizing
        // The keys of the step after the current one are requested early and arrive while the current ones are counted.
        // MODE 1 (one round at the default sizing; its registers are spoken for): the first round of this CTA's next
        // bucket, requested here; a further round of a bucket loads its keys as it goes.  MODE 0 (duplicate-rich data may
        // use large buckets): every round hands over to the bucket's next round or to the next bucket's first round.
        if (MODE == 1) {
#pragma unroll
            for (int j = 0; j < RC_PREFETCH; ++j) kcur[j] = knext[j];
            const u32 bn = b + gridDim.x;
            if (bn < nb) {
                lo_n = sub_base[bn];
                n_n = sub_base[bn + 1] - lo_n;
#pragma unroll
                for (int j = 0; j < RC_PREFETCH; ++j) {
                    const u32 i = j * RC_THREADS + threadIdx.x;
                    knext[j] = i < n_n ? keys2[lo_n + i] : 0ull;
                }
            }
        }
        // ---- pass 1 ----
        for (u32 rd = 0; rd < max(nrounds, 1u); ++rd) {
            const u32 off = rd * RC_ROUND;
            if (MODE == 0) {
#pragma unroll
                for (int j = 0; j < RC_PREFETCH; ++j) kcur[j] = knext[j];
                if (rd + 1 < nrounds) {                          // (uniform)
#pragma unroll
                    for (int j = 0; j < RC_PREFETCH; ++j) {
                        const u32 i = off + RC_ROUND + j * RC_THREADS + threadIdx.x;
                        knext[j] = i < n ? keys2[lo + i] : 0ull;
                    }
                } else if (b + gridDim.x < nb) {
                    const u32 bn = b + gridDim.x;
                    lo_n = sub_base[bn];
                    n_n = sub_base[bn + 1] - lo_n;
#pragma unroll
                    for (int j = 0; j < RC_PREFETCH; ++j) {
                        const u32 i = j * RC_THREADS + threadIdx.x;
                        knext[j] = i < n_n ? keys2[lo_n + i] : 0ull;
                    }
                }
            }
#pragma unroll
            for (int j = 0; j < RC_PREFETCH; ++j) {
                const u32 i = off + j * RC_THREADS + threadIdx.x;
                if (i < n) {
                    const ull key = (MODE == 0 || rd == 0) ? kcur[j] : keys2[lo + i];
                    if (MODE == 1) rc_pass1(key, bm, tkeys, tcnt, claimed, flist, scal, need_at);
                    else if (key == HC_EMPTY) smem_red_inc(&scal[4]);
                    else rc_insert0(key, tkeys, tcnt, scal);
                }
            }
            if (MODE == 1) {
                BLOCK_SYNC();
                // pass 1b: the queued keys claim / bump their table slots, one key per lane
                const u32 nq = min(scal[5], RC_FLIST);
                for (u32 i = threadIdx.x; i < nq; i += RC_THREADS) {
                    const ull key = flist[i];
                    if (rc_flagged(key, rc_hash(key), tkeys, tcnt, claimed, scal) >= need_at) scal[3] = 1;
                }
                if (rd + 1 < nrounds) {                          // (uniform) another round re-uses the queue
                    BLOCK_SYNC();
                    if (threadIdx.x == 0) scal[5] = 0;
                    BLOCK_SYNC();
                }
            }
        }
        BLOCK_SYNC();
        bool ovf = scal[2] != 0;
        if (MODE == 1) cta_flagged += scal[5];
        const u32 nd = MODE == 1 ? min(scal[1], (u32)RC_CLAIM_CAP) : RC_SLOTS;      // table slots the emit pass visits
        bool exact = MODE == 0;
        if (MODE == 1) {
            const bool need2 = !ovf && scal[3];
            // exact counts are needed only if some key may reach min_count: clear the flagged-occurrence counters ...
            if (need2) {
                for (u32 i = threadIdx.x; i < nd; i += RC_THREADS) tcnt[claimed[i]] = 0;
                for (u32 i = threadIdx.x; i < RC_FINE; i += RC_THREADS) bins[i] = 0;      // (the queue of pass 1 lived here)
                BLOCK_SYNC();
                // ... and pass 2 counts every occurrence of the keys that have a slot
                for (u32 rd = 0; rd < nrounds; ++rd) {
                    const u32 off = rd * RC_ROUND;
#pragma unroll
                    for (int j = 0; j < RC_PREFETCH; ++j) {
                        const u32 i = off + j * RC_THREADS + threadIdx.x;
                        if (i < n) rc_pass2(rd == 0 ? kcur[j] : keys2[lo + i], tkeys, tcnt, scal);
                    }
                }
                BLOCK_SYNC();
                exact = true;
            }
        }
        // ---- ordered emit ----
        // slot visited by step i of the emit loops: MODE 1 walks the claim log, MODE 0 the whole table
        auto slot_of = [&](u32 i) { return MODE == 1 ? (u32)claimed[i] : i; };
        auto fine = [&](ull key) { return rp_fine(d1, rp_prefix(key, r.down, r.up), RC_FINE_LOG2); };
        const u32 n_empty = exact ? scal[4] : 0u;
        u32 nsurv = 0;
        if (exact && !ovf) {                                       // (block-uniform)
            // A: survivors per ordering bin (bins are zero here)
            u32 held = 0;                                          // MODE 0: distinct keys of the bucket (occupied slots)
            for (u32 i = threadIdx.x; i < nd; i += RC_THREADS) {
                const u32 p = slot_of(i);
                const u32 cnt = tcnt[p];
                if (MODE == 0) held += cnt ? 1u : 0u;
                if (cnt >= c) smem_red_inc(&bins[fine(tkeys[p])]);
            }
            if (MODE == 0) {
                held = __reduce_add_sync(0xffffffffu, held);
                if (lane == 0 && held) atomicAdd(&scal[1], held);
            }
            BLOCK_SYNC();
            if (MODE == 0) cta_flagged += scal[1];
            // exclusive scan of the bins (2 per thread), in place: bins[d] = first staging row of bin d; total = survivors
            const u32 c0 = bins[2 * threadIdx.x], c1 = bins[2 * threadIdx.x + 1];
            __syncwarp();
            u32 incl = c0 + c1;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const u32 y = __shfl_up_sync(0xffffffffu, incl, d);
                if (lane >= d) incl += y;
            }
            if (lane == 31) s_warp[warp] = incl;
            BLOCK_SYNC();
            if (warp == 0) {
                u32 w = lane < RC_THREADS / 32 ? s_warp[lane] : 0u;
                u32 wi = w;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const u32 y = __shfl_up_sync(0xffffffffu, wi, d);
                    if (lane >= d) wi += y;
                }
                if (lane < RC_THREADS / 32) s_warp[lane] = wi - w;
                if (lane == RC_THREADS / 32 - 1) s_warp[RC_THREADS / 32] = wi;
            }
            BLOCK_SYNC();
            nsurv = s_warp[RC_THREADS / 32];
            const u32 ex = s_warp[warp] + incl - (c0 + c1);
            bins[2 * threadIdx.x] = ex;
            bins[2 * threadIdx.x + 1] = ex + c0;
            if (nsurv > RC_STAGE_ROWS) { ovf = true; nsurv = 0; }    // (MODE 0, min_count 1, table almost full of distinct keys)
            BLOCK_SYNC();
        }
        if (ovf && threadIdx.x == 0) ovf_list[atomicAdd(ovf_n, 1u)] = b;
        if (nsurv) {                                               // (block-uniform)
            // B: place (after this pass bins[d] = END of bin d); the table slots are released here
            for (u32 i = threadIdx.x; i < nd; i += RC_THREADS) {
                const u32 p = slot_of(i);
                const u32 cnt = tcnt[p];
                if (MODE == 0 && cnt == 0) continue;              // (an empty slot of the table)
                const ull key = tkeys[p];
                tkeys[p] = HC_EMPTY;
                tcnt[p] = 0;
                if (cnt >= c) {
                    const u32 at = smem_atom_inc(&bins[fine(key)]);
                    skey[at] = key;
                    scnt[at] = cnt;
                }
            }
            BLOCK_SYNC();
            // C: rank inside the bin, write
            RcRow* dst = out + sub_base[b] / c;
            for (u32 i = threadIdx.x; i < nsurv; i += RC_THREADS) {
                const ull key = skey[i];
                const u32 d = fine(key);
                const u32 s0 = d ? bins[d - 1] : 0u, s1 = bins[d];
                u32 rank = 0;
                for (u32 j = s0; j < s1; ++j) rank += skey[j] < key ? 1u : 0u;
                RcRow row;
                row.key = key;
                row.count = scnt[i];
                dst[s0 + rank] = row;
            }
            if (n_empty >= c && threadIdx.x == 0) {               // the all-ones key is the largest key there is
                RcRow row;
                row.key = HC_EMPTY;
                row.count = n_empty;
                dst[nsurv] = row;
            }
            if (threadIdx.x == 0) rows[b] = nsurv + (n_empty >= c ? 1u : 0u);
        } else {
            // nothing to order: release the table slots
            for (u32 i = threadIdx.x; i < nd; i += RC_THREADS) {
                const u32 p = slot_of(i);
                if (MODE == 0 && tcnt[p] == 0) continue;
                tkeys[p] = HC_EMPTY;
                tcnt[p] = 0;
            }
            if (threadIdx.x == 0) {
                const bool one = exact && !ovf && n_empty >= c && n_empty > 0;
                if (one) {
                    RcRow row;
                    row.key = HC_EMPTY;
                    row.count = n_empty;
                    out[sub_base[b] / c] = row;
                }
                rows[b] = one ? 1u : 0u;
            }
        }
        BLOCK_SYNC();
        {   // region A: staging -> bitmap again (MODE 1, all of it: the staging area overlays it); region B: bins -> zero / queue
            if (MODE == 1) {
                uint4* bm4 = reinterpret_cast<uint4*>(bm);
                for (u32 i = threadIdx.x; i < RC_BM_WORDS / 4; i += RC_THREADS) bm4[i] = make_uint4(0, 0, 0, 0);
            } else {
                for (u32 i = threadIdx.x; i < RC_FINE; i += RC_THREADS) bins[i] = 0;
            }
            if (threadIdx.x < 8) s_scal[par ^ 1u][threadIdx.x] = 0;
        }
        BLOCK_SYNC();
    }
    if (threadIdx.x == 0 && flagged_total && cta_flagged) atomicAdd(flagged_total, cta_flagged);
}

// ---- survivor slots -> one dense array ----------------------------------------------------------------------------------
// rc_offsets (one CTA): off[b] = *base + rows before bucket b; *total = rows of this launch; *base += total when ADVANCE.
__global__ void __launch_bounds__(1024)
rc_offsets_kernel(const u32* __restrict__ rows, u32 nb, u64* __restrict__ off, ull* __restrict__ base_io, ull* __restrict__ total_out, int advance) {
    __shared__ u64 sm[1024 / 32 + 1];
    const u32 per = (nb + 1023) / 1024;
    const u32 t0 = min(nb, threadIdx.x * per), t1 = min(nb, t0 + per);
    u64 acc = 0;
    for (u32 t = t0; t < t1; ++t) acc += rows[t];
    u64 total;
    u64 run = block_exclusive_sum64<32>(acc, sm, &total) + (base_io ? (u64)*base_io : 0ull);
    for (u32 t = t0; t < t1; ++t) { off[t] = run; run += rows[t]; }
    BLOCK_SYNC();
    if (threadIdx.x == 0) {
        *total_out = total;
        if (base_io && advance) *base_io += total;
    }
}
// one warp per sub-bucket: rows out of the slot array into (a) separate key / count arrays or (b) a dense row array
__global__ void __launch_bounds__(256)
rc_gather_kernel(const RcRow* __restrict__ slots, const u32* __restrict__ sub_base, const u32* __restrict__ rows, const u64* __restrict__ off,
                 u32 nb, u32 c, u64* __restrict__ out_keys, u64* __restrict__ out_counts, RcRow* __restrict__ out_rows) {
    const u32 lane = threadIdx.x & 31;
    for (u32 b = blockIdx.x * 8 + (threadIdx.x >> 5); b < nb; b += gridDim.x * 8) {
        const u32 m = rows[b];
        if (!m) continue;
        const RcRow* src = slots + sub_base[b] / c;
        const u64 at = off[b];
        for (u32 i = lane; i < m; i += 32) {
            const RcRow row = src[i];
            if (out_rows) out_rows[at + i] = row;
            else { out_keys[at + i] = row.key; out_counts[at + i] = row.count; }
        }
    }
}
// Device-driven handling of overflowed sub-buckets when the host does not wait for every group of a very large chunk:
// their keys are appended to one array (counted by the sort path after the last group).  counters: [1] overflow keys,
// [3] keys of all groups, [4] overflowed sub-buckets.  Keys beyond `cap` are not copied (the host sees [1] > cap and
// redoes the chunk group by group).
__global__ void __launch_bounds__(256)
rc_overflow_collect_kernel(const u32* __restrict__ ovf_list, const u32* __restrict__ ovf_n, const u32* __restrict__ sub_base, const u64* __restrict__ keys2,
                           const ull* __restrict__ group_total, u64* __restrict__ ovf_keys, u64 cap, ull* __restrict__ counters) {
    __shared__ ull s_at;
    const u32 n_ovf = *ovf_n;
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        atomicAdd(&counters[3], *group_total);
        if (n_ovf) atomicAdd(&counters[4], (ull)n_ovf);
    }
    for (u32 i = blockIdx.x; i < n_ovf; i += gridDim.x) {
        const u32 b = ovf_list[i];
        const u32 lo = sub_base[b], n = sub_base[b + 1] - lo;
        if (threadIdx.x == 0) s_at = atomicAdd(&counters[1], (ull)n);
        BLOCK_SYNC();
        const ull at = s_at;
        if (at + n <= cap)
            for (u32 j = threadIdx.x; j < n; j += 256) ovf_keys[at + j] = keys2[lo + j];
        BLOCK_SYNC();
    }
}
__global__ void rc_split_rows_kernel(const RcRow* __restrict__ in, u64 n, u64* __restrict__ keys, u64* __restrict__ counts) {
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { const RcRow row = in[i]; keys[i] = row.key; counts[i] = row.count; }
}

// rows -> (key, 32-bit count) arrays; *too_big is set when a count needs more than 32 bits
__global__ void rc_split_rows32_kernel(const RcRow* __restrict__ in, u64 n, u64* __restrict__ keys, u32* __restrict__ counts, u32* __restrict__ too_big) {
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        const RcRow row = in[i];
        keys[i] = row.key;
        counts[i] = (u32)row.count;
        if (row.count >> 32) *too_big = 1u;
    }
}
__global__ void narrow_counts_kernel(const u64* __restrict__ counts, u64 n, u32* __restrict__ out, u32* __restrict__ too_big) {
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        const u64 c = counts[i];
        out[i] = (u32)c;
        if (c >> 32) *too_big = 1u;
    }
}

__global__ void rc_join_rows_kernel(const u64* __restrict__ keys, const u64* __restrict__ counts, u64 n, RcRow* __restrict__ out) {
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { RcRow row; row.key = keys[i]; row.count = counts[i]; out[i] = row; }
}

// ---- debug: every key of the level-1 / level-2 arrays must sit in the range of its own bucket -------------------
__global__ void hc_verify_kernel(const u64* __restrict__ keys, const u32* __restrict__ sub_base, u32 nb, u32 step /*1: sub-buckets, HC_NB2: level-1*/,
                                 u32 total, ull* __restrict__ bad_count, RpView r) {
    const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const u32 p = rp_prefix(keys[i], r.down, r.up);
    const u32 b1 = r.lut[rp_lut_index(p, r.base, r.sh)];
    const u32 b = b1 * HC_NB2 + rp_b2(r.l1[b1], p);
    const u32 lo_b = (b / step) * step, hi_b = min(nb, lo_b + step);
    if (i < sub_base[lo_b] || i >= sub_base[hi_b]) atomicAdd(bad_count, 1ull);
}
