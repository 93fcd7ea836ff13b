// K3 (hash variant) -- large-k counting by hash partitioning + shared-memory hash tables.
//
// Replaces  kmerlist[k] += 1  over ~10^8 mostly distinct 64-bit keys per chunk
// (lib/mercat2_kmers.py:56-60) followed by the per-file  count >= min_count  filter (:73-78).
//
// Sized from measurements on B200 (tools/microbench.cu, profiles/r01_microbench.txt): random
// shared-memory atomics run at ~1.5 T/s chip-wide and coalesced traffic at ~6.2 TB/s, while random
// 8-byte global stores reach only ~25-50 G/s and an LSD radix sort needs 8 passes.  So keys are
// hash-partitioned with COALESCED writes in two levels (nb1 <= 512 buckets, then 128 sub-buckets each)
// until a sub-bucket (~7k keys) fits a 16384-slot open-addressing table in shared memory; one CTA
// counts one sub-bucket entirely on chip and emits only the keys whose count reaches the threshold.
//
//   hc_hist     symbols -> histogram over all nb1*nb2 sub-buckets          (reads sym once)
//   hc_scan     exact sub-bucket / bucket offsets, cursors, tile map       (one CTA)
//   hc_scatter1 symbols -> keys grouped by level-1 bucket                  (reads sym, writes 8 B/key)
//   hc_scatter2 level-1 groups -> keys grouped by sub-bucket               (reads 8, writes 8 B/key)
//   hc_count    per sub-bucket: smem hash insert, threshold, emit          (reads 8 B/key)
//
// Offsets are exact (two passes over the symbols), so skewed inputs cannot overflow a bucket's memory;
// a sub-bucket with more distinct keys than the table holds is reported in an overflow list and redone
// by the sort path.  Output rows are unordered (sorted once per sample at the end).
#pragma once
#include "common.cuh"
#include "extract.cuh"

#define HC_SLOTS 16384u
#define HC_THREADS 1024
#define HC_LIMIT 12288u                 // distinct keys per table before a bucket is declared overflowed
#define HC_EMPTY 0xFFFFFFFFFFFFFFFFull
#define HC_NB2_LOG2 7
#define HC_NB2 (1u << HC_NB2_LOG2)
#define HC_MAX_NB1 400u
#define HC_TILE 4096u
#define HC_STAGE_SLOTS (HC_TILE + 2)                    // one tile + a spare slot for invalid keys (kept 16-byte aligned)
#define HC_SDST_OFFSET ((size_t)HC_STAGE_SLOTS * 8)
#define HC_SCATTER_SMEM ((size_t)HC_STAGE_SLOTS * 12)  // staged keys (8 B) + destination indices (4 B)
#define HC_SCATTER_SMEM16 ((size_t)HC_STAGE_SLOTS * 10) // staged keys (8 B) + 16-bit digits

#define HC_MULT1 0x9E3779B97F4A7C15ull           // bucket hash of a chunk's keys
#define HC_MULT2 0xC2B2AE3D27D4EB4Full           // independent bucket hash inside a level-0 group (very large chunks)
#define HC_MULT3 0xA0761D6478BD642Full           // level-0 hash of a key array that was itself selected by HC_MULT1 (keys received from other GPUs)
__device__ __forceinline__ u32 hc_bucket(u64 key, u32 nb, u64 mult = HC_MULT1) {
    const u32 h = (u32)((key * mult) >> 32);
    return __umulhi(h, nb);                               // uniform in [0, nb)
}
// Shared-memory atomics issued from inside divergent probe loops go through inline PTX: the compiler otherwise
// rewrites atomicAdd(addr, 1) into a warp-aggregated VOTE + leader ATOMS + SHFL sequence.
__device__ __forceinline__ u32 smem_atom_inc(u32* addr) {
    u32 old;
    asm volatile("atom.shared.add.u32 %0, [%1], 1;" : "=r"(old) : "r"((u32)__cvta_generic_to_shared(addr)) : "memory");
    return old;
}
__device__ __forceinline__ void smem_red_inc(u32* addr) {
    asm volatile("red.shared.add.u32 [%0], 1;" :: "r"((u32)__cvta_generic_to_shared(addr)) : "memory");
}

__device__ __forceinline__ u32 hc_slot(u64 key) {
    return (u32)((key * 0xD6E8FEB86659FD93ull) >> 43) & (HC_SLOTS - 1);
}

// ---- hc_hist -------------------------------------------------------------------------------------------
template <int ENC>
__global__ void __launch_bounds__(EX_THREADS)
hc_hist_kernel(SymView v, u64 s0, u64 s1, int k, u32 nb, u32* __restrict__ ghist) {
    extern __shared__ __align__(16) u8 dyn[];
    u32* hist = reinterpret_cast<u32*>(dyn);
    __shared__ u64 s_code[EX_THREADS + EX_HALO];
    __shared__ u32 s_meta[EX_THREADS + EX_HALO];
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
        tile_walk<ENC>(ctx, k, [&](int i, u64 code, bool fast) {
            if (fast && first + i < s1) atomicAdd(&hist[hc_bucket(code & mask, nb)], 1u);
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
// `valid` says mine[i] holds a key.  cursors[d] is advanced atomically by the tile's count for digit d.  Keys are
// re-ordered through shared memory so that consecutive threads store consecutive addresses of one digit's run.
// `mine` yields key i in the ranking phase, `again` in the staging phase: the same registers, or a re-load that
// lets the keys die in between (fewer live registers, more resident CTAs).
// `base64` (may be NULL): 64-bit start of every digit's region; the cursors are then relative to it (level-0 groups of
// chunks with more than 2^32 windows).
// FULL: all 16 keys of every thread are valid (interior tiles): no per-key tests.  Otherwise the rank atomic is
// predicated in PTX and invalid keys are staged into a spare slot, so the unrolled per-key code stays free of
// branches (every branch re-derives the shared-memory window base and brackets itself with BSSY/BSYNC).
__device__ __forceinline__ u32 smem_atom_inc_if(u32 addr32, u32 pred) {
    u32 old = 0;
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %2, 0;\n\t@q atom.shared.add.u32 %0, [%1], 1;\n\t}" : "+r"(old) : "r"(addr32), "r"(pred) : "memory");
    return old;
}

__device__ __forceinline__ void smem_red_inc_if(u32 addr32, u32 pred) {
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %1, 0;\n\t@q red.shared.add.u32 [%0], 1;\n\t}" :: "r"(addr32), "r"(pred) : "memory");
}

template <bool USE_DST, bool FULL, class KeyFn, class KeyFn2, class DigitFn>
__device__ __forceinline__ void hc_group_and_write3(KeyFn mine, KeyFn2 again, u32 valid, u32 nd, DigitFn dig, u64* stage, u32* sdst,
                                                    u32* cnt, u32* loff, u32* gbase, u32* sm, u32* __restrict__ cursors,
                                                    u64* __restrict__ out, const u64* __restrict__ base64, u32 spare /*stage slot for invalid keys*/) {
    u32 rd[16];                             // (rank within (tile, digit)) << 16 | digit
    const u32 cnt32 = (u32)__cvta_generic_to_shared(cnt);
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const u32 d = dig(mine(i));
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
        if (!FULL) pos = ((valid >> i) & 1u) ? pos : spare;
        stage[pos] = again(i);
        if (USE_DST) sdst[pos] = gbase[d] + r;
        else reinterpret_cast<u16*>(sdst)[pos] = (u16)d;
    }
    BLOCK_SYNC();
    if (USE_DST) {
        for (u32 i = threadIdx.x; i < total; i += EX_THREADS) out[sdst[i]] = stage[i];
    } else {
        for (u32 i = threadIdx.x; i < total; i += EX_THREADS) {
            const u32 d = reinterpret_cast<u16*>(sdst)[i];
            const u64 at = (base64 ? base64[d] : 0ull) + gbase[d] + (i - loff[d]);
            out[at] = stage[i];
        }
    }
}

template <bool USE_DST, class KeyFn, class KeyFn2, class DigitFn>
__device__ __forceinline__ void hc_group_and_write2(KeyFn mine, KeyFn2 again, u32 valid, u32 nd, DigitFn dig, u64* stage, u32* sdst,
                                                    u32* cnt, u32* loff, u32* gbase, u32* sm, u32* __restrict__ cursors,
                                                    u64* __restrict__ out, const u64* __restrict__ base64 = nullptr) {
    // (stage / sdst hold one spare slot behind the tile: HC_SPARE_SLOT)
    hc_group_and_write3<USE_DST, false>(mine, again, valid, nd, dig, stage, sdst, cnt, loff, gbase, sm, cursors, out, base64, HC_TILE);
}

template <bool USE_DST, class KeyFn, class DigitFn>
__device__ __forceinline__ void hc_group_and_write(KeyFn mine, u32 valid, u32 nd, DigitFn dig, u64* stage, u32* sdst,
                                                   u32* cnt, u32* loff, u32* gbase, u32* sm, u32* __restrict__ cursors,
                                                   u64* __restrict__ out) {
    hc_group_and_write2<USE_DST>(mine, mine, valid, nd, dig, stage, sdst, cnt, loff, gbase, sm, cursors, out);
}

// ---- hc_scatter1: symbols -> level-1 groups ----------------------------------------------------------------
template <int ENC>
__global__ void __launch_bounds__(EX_THREADS)
hc_scatter1_kernel(SymView v, u64 s0, u64 s1, int k, u32 nb, u32 nb1, u32* __restrict__ cur1, u64* __restrict__ keys1) {
    __shared__ u64 s_code[EX_THREADS + EX_HALO];
    __shared__ u32 s_meta[EX_THREADS + EX_HALO];
    extern __shared__ __align__(16) u8 dyn_sc[];                    // HC_SCATTER_SMEM bytes
    u64* stage = reinterpret_cast<u64*>(dyn_sc);
    u32* sdst = reinterpret_cast<u32*>(dyn_sc + HC_SDST_OFFSET);
    __shared__ u32 cnt[HC_MAX_NB1], loff[HC_MAX_NB1], gbase[HC_MAX_NB1];
    __shared__ u32 sm[EX_WARPS + 1];
    for (u32 i = threadIdx.x; i < nb1; i += EX_THREADS) cnt[i] = 0;
    const int kb = k * EncTraits<ENC>::BITS;
    const u64 mask = kb >= 64 ? ~0ull : ((1ull << kb) - 1);
    const u64 tile_start = s0 + (u64)blockIdx.x * EX_TILE;
    TileCtx<ENC> ctx;
    tile_begin<ENC>(v, tile_start, k, s_code, s_meta, ctx);        // contains the barrier that publishes cnt = 0
    const u64 first = tile_start + 16ull * threadIdx.x;
    u64 mine[16];
    u32 valid = 0;
    tile_walk<ENC>(ctx, k, [&](int i, u64 code, bool fast) {
        mine[i] = code & mask;
        if (fast && first + i < s1) valid |= 1u << i;
    });
    auto dig = [nb](u64 key) { return hc_bucket(key, nb) >> HC_NB2_LOG2; };      // nb == nb1 * HC_NB2
    hc_group_and_write<true>([&](int i) { return mine[i]; }, valid, nb1, dig, stage, sdst, cnt, loff, gbase, sm, cur1, keys1);
}

// ---- hc_scatter2: level-1 groups -> sub-buckets --------------------------------------------------------------
template <bool USE_DST, bool RELOAD>
__global__ void __launch_bounds__(EX_THREADS, RELOAD ? 5 : 3)
hc_scatter2_kernel(const u64* __restrict__ keys1, const u32* __restrict__ sub_base, const u32* __restrict__ tile_pref,
                   u32 nb, u32 nb1, u32 nb2, u32* __restrict__ cur2, u64* __restrict__ keys2, u64 mult) {
    extern __shared__ __align__(16) u8 dyn_sc[];                    // HC_SCATTER_SMEM bytes
    u64* stage = reinterpret_cast<u64*>(dyn_sc);
    u32* sdst = reinterpret_cast<u32*>(dyn_sc + HC_SDST_OFFSET);
    __shared__ u32 cnt[HC_NB2], loff[HC_NB2], gbase[HC_NB2];
    __shared__ u32 sm[EX_WARPS + 1];
    __shared__ u32 s_b1;
    if (blockIdx.x >= tile_pref[nb1]) return;
    if (threadIdx.x == 0) {                      // last b1 with tile_pref[b1] <= blockIdx.x
        u32 lo = 0, hi = nb1;
        while (hi - lo > 1) { const u32 mid = (lo + hi) / 2; if (tile_pref[mid] <= blockIdx.x) lo = mid; else hi = mid; }
        s_b1 = lo;
    }
    for (u32 i = threadIdx.x; i < nb2; i += EX_THREADS) cnt[i] = 0;
    BLOCK_SYNC();
    const u32 b1 = s_b1;
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
    auto dig = [nb, nb2, mult](u64 key) { return hc_bucket(key, nb, mult) & (nb2 - 1); };
    const bool full = base + HC_TILE <= hi;                  // (block-uniform) interior tile: every key slot is valid
    if (!RELOAD && full) {
        auto key = [&](int i) { return mine[i]; };
        hc_group_and_write3<USE_DST, true>(key, key, valid, nb2, dig, stage, sdst, cnt, loff, gbase, sm, cur2 + b1 * nb2, keys2, nullptr, HC_TILE);
    } else if (RELOAD) {
        // the staging phase reads the keys again (an L2 hit: the tile was read a few microseconds ago) instead of
        // carrying 32 registers across two barriers
        auto again = [&](int j) {
            u64 x;
            asm volatile("ld.global.cg.u64 %0, [%1];" : "=l"(x) : "l"(keys1 + base + j * EX_THREADS + threadIdx.x));
            return x;
        };
        auto first = [&](int i) { return mine[i]; };
        if (full) hc_group_and_write3<USE_DST, true>(first, again, valid, nb2, dig, stage, sdst, cnt, loff, gbase, sm, cur2 + b1 * nb2, keys2, nullptr, HC_TILE);
        else hc_group_and_write2<USE_DST>(first, again, valid, nb2, dig, stage, sdst, cnt, loff, gbase, sm, cur2 + b1 * nb2, keys2);
    } else {
        hc_group_and_write<USE_DST>([&](int i) { return mine[i]; }, valid, nb2, dig, stage, sdst, cnt, loff, gbase, sm, cur2 + b1 * nb2, keys2);
    }
}

// ---- key-array sources (level-0 groups of very large chunks): histogram and level-1 scatter over 64-bit keys ----
#define HK_HIST_THREADS 1024
__global__ void __launch_bounds__(HK_HIST_THREADS)
hk_hist_kernel(const u64* __restrict__ keys, u64 n, u32 nb, u64 mult, u32* __restrict__ ghist) {
    extern __shared__ __align__(16) u8 dyn[];
    u32* hist = reinterpret_cast<u32*>(dyn);
    for (u32 i = threadIdx.x; i < nb; i += HK_HIST_THREADS) hist[i] = 0;
    BLOCK_SYNC();
    for (u64 i = (u64)blockIdx.x * HK_HIST_THREADS + threadIdx.x; i < n; i += (u64)gridDim.x * HK_HIST_THREADS)
        atomicAdd(&hist[hc_bucket(keys[i], nb, mult)], 1u);
    BLOCK_SYNC();
    for (u32 b = threadIdx.x; b < nb; b += HK_HIST_THREADS) {
        const u32 c = hist[b];
        if (c) atomicAdd(&ghist[b], c);
    }
}

__global__ void __launch_bounds__(EX_THREADS, 3)
hk_scatter1_kernel(const u64* __restrict__ keys, u64 n, u32 nb, u32 nb1, u64 mult, u32* __restrict__ cur1, u64* __restrict__ keys1,
                   const u64* __restrict__ base64) {
    extern __shared__ __align__(16) u8 dyn_sc[];                    // HC_TILE * 10 bytes
    u64* stage = reinterpret_cast<u64*>(dyn_sc);
    u32* sdst = reinterpret_cast<u32*>(dyn_sc + HC_SDST_OFFSET);
    __shared__ u32 cnt[HC_MAX_NB1], loff[HC_MAX_NB1], gbase[HC_MAX_NB1];
    __shared__ u32 sm[EX_WARPS + 1];
    for (u32 i = threadIdx.x; i < nb1; i += EX_THREADS) cnt[i] = 0;
    BLOCK_SYNC();
    const u64 base = (u64)blockIdx.x * HC_TILE;
    u64 mine[16];
    u32 valid = 0;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        const u64 i = base + (u64)j * EX_THREADS + threadIdx.x;
        mine[j] = 0;
        if (i < n) { mine[j] = keys[i]; valid |= 1u << j; }
    }
    auto dig = [nb, mult](u64 key) { return hc_bucket(key, nb, mult) >> HC_NB2_LOG2; };
    auto key = [&](int i) { return mine[i]; };
    hc_group_and_write2<false>(key, key, valid, nb1, dig, stage, sdst, cnt, loff, gbase, sm, cur1, keys1, base64);
}

// ---- hc_count: persistent CTAs, one sub-bucket at a time ---------------------------------------------------------
// The table is initialised once per CTA; every claimed slot is logged in `claimed`, so thresholding and
// clean-up touch only the distinct keys of the bucket (not all 16384 slots).  The keys of the NEXT bucket
// are prefetched into registers while the current one is inserted.
#define HC_PREFETCH 8
#define HC_CLAIM_CAP (HC_LIMIT + HC_THREADS)
#define HC_COUNT_SMEM ((size_t)HC_SLOTS * 12 + (size_t)HC_CLAIM_CAP * 2)

__device__ __forceinline__ void hc_insert(ull key, ull* tkeys, u32* tcnt, u16* claimed, u32* s_distinct, u32* s_overflow) {
    u32 p = hc_slot(key);
    while (true) {
        ull cur = tkeys[p];
        if (cur == HC_EMPTY) {
            cur = atomicCAS(&tkeys[p], HC_EMPTY, key);
            if (cur == HC_EMPTY) {
                const u32 d = smem_atom_inc(s_distinct);
                if (d < HC_CLAIM_CAP) claimed[d] = (u16)p;
                if (d >= HC_LIMIT) *s_overflow = 1;
                cur = key;
            }
        }
        if (cur == key) { smem_red_inc(&tcnt[p]); return; }
        p = (p + 1) & (HC_SLOTS - 1);
    }
}

__global__ void __launch_bounds__(HC_THREADS, 1)
hc_count_kernel(const u64* __restrict__ keys2, const u32* __restrict__ sub_base, u32 nb, u64 c, u64* __restrict__ out_keys,
                u64* __restrict__ out_cnt, ull* __restrict__ out_n, u64 out_cap, u32* __restrict__ ovf_list, u32* __restrict__ ovf_n) {
    extern __shared__ __align__(16) u8 dyn[];
    ull* tkeys = reinterpret_cast<ull*>(dyn);                               // HC_SLOTS
    u32* tcnt = reinterpret_cast<u32*>(dyn + (size_t)HC_SLOTS * 8);           // HC_SLOTS
    u16* claimed = reinterpret_cast<u16*>(dyn + (size_t)HC_SLOTS * 12);       // HC_CLAIM_CAP
    __shared__ u32 s_empty, s_distinct, s_overflow;
    for (u32 i = threadIdx.x; i < HC_SLOTS; i += HC_THREADS) { tkeys[i] = HC_EMPTY; tcnt[i] = 0; }
    if (threadIdx.x == 0) { s_empty = 0; s_distinct = 0; s_overflow = 0; }
    const int lane = threadIdx.x & 31;
    ull knext[HC_PREFETCH];
    u32 b = blockIdx.x;
    u32 lo_n = 0, n_n = 0;
    if (b < nb) {
        lo_n = sub_base[b];
        n_n = sub_base[b + 1] - lo_n;
#pragma unroll
        for (int j = 0; j < HC_PREFETCH; ++j) {
            const u32 i = j * HC_THREADS + threadIdx.x;
            knext[j] = i < n_n ? keys2[lo_n + i] : 0ull;
        }
    }
    BLOCK_SYNC();
    for (; b < nb; b += gridDim.x) {
        const u32 n = n_n;
        ull kcur[HC_PREFETCH];
#pragma unroll
        for (int j = 0; j < HC_PREFETCH; ++j) kcur[j] = knext[j];
        const u32 bn = b + gridDim.x;                       // prefetch the next bucket of this CTA
        if (bn < nb) {
            lo_n = sub_base[bn];
            n_n = sub_base[bn + 1] - lo_n;
#pragma unroll
            for (int j = 0; j < HC_PREFETCH; ++j) {
                const u32 i = j * HC_THREADS + threadIdx.x;
                knext[j] = i < n_n ? keys2[lo_n + i] : 0ull;
            }
        }
        const bool big = n > HC_PREFETCH * HC_THREADS;              // see hc_count2_kernel
#pragma unroll
        for (int j = 0; j < HC_PREFETCH; ++j) {
            const u32 i = j * HC_THREADS + threadIdx.x;
            if (!big && i < n && !*(volatile u32*)&s_overflow) {
                if (kcur[j] == HC_EMPTY) atomicAdd(&s_empty, 1u);      // the all-ones key (T^32) is counted aside
                else hc_insert(kcur[j], tkeys, tcnt, claimed, &s_distinct, &s_overflow);
            }
        }
        BLOCK_SYNC();
        const u32 nd = min(s_distinct, (u32)HC_CLAIM_CAP);
        const bool ovf = big || s_overflow != 0;
        const u32 n_empty = big ? 0u : s_empty;
        if (ovf && threadIdx.x == 0) ovf_list[atomicAdd(ovf_n, 1u)] = b;
        // threshold + clean-up over the claimed slots (warp-aggregated output reservation)
        for (u32 i0 = 0; i0 < nd + (n_empty ? 1u : 0u); i0 += HC_THREADS) {
            const u32 i = i0 + threadIdx.x;
            ull key = 0;
            u32 cnt = 0;
            if (i < nd) {
                const u32 p = claimed[i];
                key = tkeys[p];
                cnt = tcnt[p];
                tkeys[p] = HC_EMPTY;
                tcnt[p] = 0;
            } else if (i == nd && n_empty) {
                key = HC_EMPTY;
                cnt = n_empty;
            }
            const bool keep = !ovf && cnt >= c && cnt > 0;
            __syncwarp();
            const u32 m = __ballot_sync(0xffffffffu, keep);
            if (m) {
                ull base = 0;
                const int leader = __ffs(m) - 1;
                if (lane == leader) base = atomicAdd(out_n, (ull)__popc(m));
                base = __shfl_sync(0xffffffffu, base, leader);
                if (keep) {
                    const u64 o = base + __popc(m & ((1u << lane) - 1u));
                    if (o < out_cap) { out_keys[o] = key; out_cnt[o] = cnt; }
                }
            }
        }
        BLOCK_SYNC();
        if (threadIdx.x == 0) { s_empty = 0; s_distinct = 0; s_overflow = 0; }
        BLOCK_SYNC();
    }
}

// ---- hc_count2: bitmap pre-filter (min_count >= 2) --------------------------------------------------------------------
// Most keys of a metagenome chunk occur once.  Pass 1 sets one bit per key in a 2^19-bit shared bitmap; only a key
// that finds its bit already set (a repeat, or a ~1 % false positive) is entered into the exact table.  Pass 2
// re-walks the keys (still in registers) and counts those present in the table.  Every key occurring >= 2 times is
// in the table (its 2nd occurrence sees the bit), its count is exact (all occurrences are counted in pass 2), keys
// occurring once can never reach min_count >= 2.  One uniform step per key per pass instead of a divergent probe
// loop over a half-full table.
#define HC2_THREADS 512
#define HC2_SLOTS 4096u
#define HC2_LIMIT 3072u
#define HC2_CLAIM_CAP (HC2_LIMIT + HC2_THREADS)
#define HC2_BM_WORDS 8192u                       // 2^18 bits
#define HC2_BM_BITS_LOG2 18
#define HC2_PREFETCH 8
#define HC2_FLIST 1024u                          // flagged keys queued per bucket before the dense insert step
#define HC2_SMEM ((size_t)HC2_BM_WORDS * 4 + (size_t)HC2_SLOTS * 12 + (size_t)HC2_CLAIM_CAP * 2 + (size_t)HC2_FLIST * 8)

// One 32-bit multiply-add hash per key: its top 18 bits address the bitmap; the table slot comes from a second
// multiply that only the (rare) repeat path pays for.  (The bucket index used the high bits of a different, 64-bit
// product, so these bits are independent of the bucket the key sits in.)
__device__ __forceinline__ u32 hc2_hash(ull key) { return (u32)key * 0x9E3779B1u + (u32)(key >> 32) * 0x85EBCA77u; }
__device__ __forceinline__ u32 hc2_slot(u32 h) { return (h * 0xC2B2AE3Du) >> (32 - 12); }
static_assert(HC2_SLOTS == 4096u, "hc2_slot yields 12 bits");
static_assert((HC2_CLAIM_CAP * 2) % 8 == 0, "the flagged-key queue follows the claim log and holds 8-byte keys");

// Pass 1, repeat path: the key found its bitmap bit already set (a repeat, or a ~1 % false positive) and is given a
// table slot.  Returns the number of times this key has now been seen with its bit set.  The true multiplicity of
// a key is that number or that number + 1 (its very first occurrence set the bit itself unless another key's bit
// collided), so a bucket in which no key reaches min_count - 1 flagged occurrences has no survivor and needs no
// exact second pass.  The all-ones key (T^32) doubles as the empty-slot marker and is counted aside.
__device__ __noinline__ u32 hc2_flagged(ull key, u32 h, ull* tkeys, u32* tcnt, u16* claimed, u32* scal) {
    if (*(volatile u32*)&scal[2]) return 0u;                 // table already overflowed: the bucket is redone by sorting
    if (key == HC_EMPTY) return smem_atom_inc(&scal[0]) + 1u;
    u32 p = hc2_slot(h);
    while (true) {
        ull cur = tkeys[p];
        if (cur == HC_EMPTY) {
            cur = atomicCAS(&tkeys[p], HC_EMPTY, key);
            if (cur == HC_EMPTY) {
                const u32 d = smem_atom_inc(&scal[1]);
                if (d < HC2_CLAIM_CAP) claimed[d] = (u16)p;
                if (d >= HC2_LIMIT) scal[2] = 1;
                cur = key;
            }
        }
        if (cur == key) return smem_atom_inc(&tcnt[p]) + 1u;
        p = (p + 1) & (HC2_SLOTS - 1);
    }
}
// Pass 1a: bitmap test-and-set.  A flagged key is queued (scal[5] = queue length) so that the probe loops run
// afterwards with one key per lane instead of one or two active lanes per warp; if the queue is full (heavily
// duplicated data) the key is inserted on the spot.
__device__ __forceinline__ void hc2_pass1(ull key, u32* bm, ull* tkeys, u32* tcnt, u16* claimed, ull* flist, u32* scal, u32 need_at) {
    const u32 h = hc2_hash(key);
    const u32 bit = 1u << ((h >> 14) & 31u);
    const u32 old = atomicOr(&bm[h >> 19], bit);
    if (old & bit) {
        const u32 q = smem_atom_inc(&scal[5]);
        if (q < HC2_FLIST) flist[q] = key;
        else if (hc2_flagged(key, h, tkeys, tcnt, claimed, scal) >= need_at) scal[3] = 1;
    }
}
__device__ __forceinline__ u32 hc2_pass2(ull key, const ull* tkeys, u32* tcnt, u32* scal) {
    if (key == HC_EMPTY) { smem_red_inc(&scal[4]); return 1u; }
    u32 p = hc2_slot(hc2_hash(key));
    while (true) {
        const ull cur = tkeys[p];
        if (cur == key) { smem_red_inc(&tcnt[p]); return 1u; }
        if (cur == HC_EMPTY) return 0u;
        p = (p + 1) & (HC2_SLOTS - 1);
    }
}

// two CTAs per SM (88 KB of shared memory each) so that one CTA's barriers hide behind the other's work
__global__ void __launch_bounds__(HC2_THREADS, 2)
hc_count2_kernel(const u64* __restrict__ keys2, const u32* __restrict__ sub_base, u32 nb, u64 c, u64* __restrict__ out_keys,
                 u64* __restrict__ out_cnt, ull* __restrict__ out_n, u64 out_cap, u32* __restrict__ ovf_list, u32* __restrict__ ovf_n,
                 ull* __restrict__ dbg, u32 exp_flags) {
    extern __shared__ __align__(16) u8 dyn[];
    u32* bm = reinterpret_cast<u32*>(dyn);                                              // HC2_BM_WORDS
    ull* tkeys = reinterpret_cast<ull*>(dyn + (size_t)HC2_BM_WORDS * 4);                 // HC2_SLOTS
    u32* tcnt = reinterpret_cast<u32*>(dyn + (size_t)HC2_BM_WORDS * 4 + (size_t)HC2_SLOTS * 8);
    u16* claimed = reinterpret_cast<u16*>(dyn + (size_t)HC2_BM_WORDS * 4 + (size_t)HC2_SLOTS * 12);
    ull* flist = reinterpret_cast<ull*>(dyn + (size_t)HC2_BM_WORDS * 4 + (size_t)HC2_SLOTS * 12 + (size_t)HC2_CLAIM_CAP * 2);
    // per parity: flagged empty-key occurrences, distinct, overflow, need-pass-2, exact empty-key count, queue length
    __shared__ u32 s_scal[2][8];
    for (u32 i = threadIdx.x; i < HC2_SLOTS; i += HC2_THREADS) { tkeys[i] = HC_EMPTY; tcnt[i] = 0; }
    for (u32 i = threadIdx.x; i < HC2_BM_WORDS; i += HC2_THREADS) bm[i] = 0;
    if (threadIdx.x < 16) (&s_scal[0][0])[threadIdx.x] = 0;
    const int lane = threadIdx.x & 31;
    const u32 need_at = (u32)min(c - 1, (u64)0xFFFFFFFFu);     // flagged occurrences at which a key may reach min_count
    ull knext[HC2_PREFETCH];
    u32 b = blockIdx.x;
    u32 lo_n = 0, n_n = 0;
    if (b < nb) {
        lo_n = sub_base[b];
        n_n = sub_base[b + 1] - lo_n;
#pragma unroll
        for (int j = 0; j < HC2_PREFETCH; ++j) {
            const u32 i = j * HC2_THREADS + threadIdx.x;
            knext[j] = i < n_n ? keys2[lo_n + i] : 0ull;
        }
    }
    BLOCK_SYNC();
    u32 par = 0;
    for (; b < nb; b += gridDim.x, par ^= 1u) {
        u32* scal = s_scal[par];
        const u32 n = n_n;
        ull kcur[HC2_PREFETCH];
#pragma unroll
        for (int j = 0; j < HC2_PREFETCH; ++j) kcur[j] = knext[j];
        const u32 bn = b + gridDim.x;
        if (bn < nb && !(exp_flags & 1u)) {
            lo_n = sub_base[bn];
            n_n = sub_base[bn + 1] - lo_n;
#pragma unroll
            for (int j = 0; j < HC2_PREFETCH; ++j) {
                const u32 i = j * HC2_THREADS + threadIdx.x;
                knext[j] = i < n_n ? keys2[lo_n + i] : 0ull;
            }
        }
        // A bucket that does not fit the prefetch registers (> 4096 keys: never at the default sizing, which targets
        // 3500) is handed to the sort fallback like an overflowed table.  [An earlier version walked the tail of such
        // buckets from global memory in both passes; on B200 that variant lost a few counts per million keys in a
        // timing-dependent way that neither warp syncs, block fences nor uniform trip counts removed -- see DESIGN.md.]
        const bool big = n > HC2_PREFETCH * HC2_THREADS;
        const u32 n1 = big ? 0u : n;
        // pass 1: bitmap test-and-set (ten instructions per key); repeats claim a table slot
#pragma unroll
        for (int j = 0; j < HC2_PREFETCH; ++j)
            if (j * HC2_THREADS + threadIdx.x < n1) hc2_pass1(kcur[j], bm, tkeys, tcnt, claimed, flist, scal, need_at);
        BLOCK_SYNC();
        // pass 1b: the queued keys claim / bump their table slots, one key per lane
        {
            const u32 nq = min(scal[5], HC2_FLIST);
            for (u32 i = threadIdx.x; i < nq; i += HC2_THREADS) {
                const ull key = flist[i];
                if (hc2_flagged(key, hc2_hash(key), tkeys, tcnt, claimed, scal) >= need_at) scal[3] = 1;
            }
        }
        BLOCK_SYNC();
        const bool ovf = big || scal[2] != 0;
        const u32 nd = min(scal[1], (u32)HC2_CLAIM_CAP);
        const bool need2 = !ovf && scal[3];
        // exact counts are needed only if some key may reach min_count: clear the flagged-occurrence counters ...
        if (need2)
            for (u32 i = threadIdx.x; i < nd; i += HC2_THREADS) tcnt[claimed[i]] = 0;
        if (need2) BLOCK_SYNC();                               // (uniform condition)
        // ... and pass 2 counts every occurrence of the keys that have a slot
        u32 hits = 0;
        if (need2) {
#pragma unroll
            for (int j = 0; j < HC2_PREFETCH; ++j) {
                const u32 i = j * HC2_THREADS + threadIdx.x;
                if (i < n) hits += hc2_pass2(kcur[j], tkeys, tcnt, scal);
            }

        }
        if (dbg && hits) atomicAdd(&dbg[0], (ull)hits);
        BLOCK_SYNC();
        const u32 n_empty = need2 ? scal[4] : 0u;
        if (ovf && threadIdx.x == 0) ovf_list[atomicAdd(ovf_n, 1u)] = b;
        for (u32 i0 = 0; i0 < nd + (n_empty ? 1u : 0u); i0 += HC2_THREADS) {
            const u32 i = i0 + threadIdx.x;
            ull key = 0;
            u32 cnt = 0;
            if (i < nd) {
                const u32 p = claimed[i];
                key = tkeys[p];
                cnt = tcnt[p];
                tkeys[p] = HC_EMPTY;
                tcnt[p] = 0;
                if (dbg && need2) { atomicAdd(&dbg[1], (ull)cnt); atomicAdd(&dbg[2], 1ull); if (key == HC_EMPTY) atomicAdd(&dbg[3], 1ull); }
            } else if (i == nd && n_empty) {
                key = HC_EMPTY;
                cnt = n_empty;
            }
            // (without pass 2 the counters hold flagged occurrences < min_count - 1: nothing survives, only clean-up)
            const bool keep = !ovf && (need2 || i >= nd) && cnt >= c && cnt > 0;
            __syncwarp();
            const u32 m = __ballot_sync(0xffffffffu, keep);
            if (m) {
                ull base = 0;
                const int leader = __ffs(m) - 1;
                if (lane == leader) base = atomicAdd(out_n, (ull)__popc(m));
                base = __shfl_sync(0xffffffffu, base, leader);
                if (keep) {
                    const u64 o = base + __popc(m & ((1u << lane) - 1u));
                    if (o < out_cap) { out_keys[o] = key; out_cnt[o] = cnt; }
                }
            }
        }
        {   // clear the bitmap (16 bytes per store); reset the other parity's scalars for the next bucket
            uint4* bm4 = reinterpret_cast<uint4*>(bm);
            for (u32 i = threadIdx.x; i < HC2_BM_WORDS / 4; i += HC2_THREADS) bm4[i] = make_uint4(0, 0, 0, 0);
            if (threadIdx.x < 8) s_scal[par ^ 1u][threadIdx.x] = 0;
        }
        BLOCK_SYNC();
    }
}

// ---- debug: every key of the level-1 / level-2 arrays must sit in the range of its own bucket -------------------
__global__ void hc_verify_kernel(const u64* __restrict__ keys, const u32* __restrict__ sub_base, u32 nb, u32 step /*1: sub-buckets, HC_NB2: level-1*/,
                                 u32 total, ull* __restrict__ bad_count, u64 mult) {
    const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const u32 b = hc_bucket(keys[i], nb, mult);
    const u32 lo_b = (b / step) * step, hi_b = min(nb, lo_b + step);
    if (i < sub_base[lo_b] || i >= sub_base[hi_b]) atomicAdd(bad_count, 1ull);
}

// ---- hc_count3: 16-bit counter pre-filter (experiment, engine option count_variant=3; slower than hc_count2) -------
// Every key first adds 1 to a 16-bit hashed counter with RED; after a barrier it reads the counter back with a plain
// load.  A key occurring m >= 2 times reads >= m >= 2 at EVERY occurrence, so all its occurrences enter the exact
// table and the table count is exact without a second pass; a key occurring once enters only on a counter collision
// (~5 %) and is dropped by the threshold.  Counters cannot wrap: a bucket holds at most 4096 keys.
// It was built on the guess that shared-memory atomics which return a value are much slower than reductions; the
// microbenchmark (tools/microbench.cu) shows both at ~2.2 T lane-ops/s, and with 128 KB of counters only one CTA
// fits per SM, so the extra barrier is not hidden: 0.50 ms vs 0.38 ms per 76.7 M keys when it was measured.
#define HC3_THREADS 1024
#define HC3_PREFETCH 4
#define HC3_CNT_LOG2 16                                   // 2^16 counters of 16 bits = 128 KB
#define HC3_CNT_WORDS (1u << (HC3_CNT_LOG2 - 1))
#define HC3_SLOTS 4096u
#define HC3_LIMIT 3072u
#define HC3_CLAIM_CAP (HC3_LIMIT + HC3_THREADS)
#define HC3_SMEM ((size_t)HC3_CNT_WORDS * 4 + (size_t)HC3_SLOTS * 12 + (size_t)HC3_CLAIM_CAP * 2)

__device__ __forceinline__ void smem_red_add(u32* addr, u32 v) {
    asm volatile("red.shared.add.u32 [%0], %1;" :: "r"((u32)__cvta_generic_to_shared(addr)), "r"(v) : "memory");
}

__global__ void __launch_bounds__(HC3_THREADS, 1)
hc_count3_kernel(const u64* __restrict__ keys2, const u32* __restrict__ sub_base, u32 nb, u64 c, u64* __restrict__ out_keys,
                 u64* __restrict__ out_cnt, ull* __restrict__ out_n, u64 out_cap, u32* __restrict__ ovf_list, u32* __restrict__ ovf_n) {
    extern __shared__ __align__(16) u8 dyn[];
    u32* cnt16 = reinterpret_cast<u32*>(dyn);                                                  // HC3_CNT_WORDS
    ull* tkeys = reinterpret_cast<ull*>(dyn + (size_t)HC3_CNT_WORDS * 4);                       // HC3_SLOTS
    u32* tcnt = reinterpret_cast<u32*>(dyn + (size_t)HC3_CNT_WORDS * 4 + (size_t)HC3_SLOTS * 8);
    u16* claimed = reinterpret_cast<u16*>(dyn + (size_t)HC3_CNT_WORDS * 4 + (size_t)HC3_SLOTS * 12);
    __shared__ u32 s_scal[2][4];                              // per parity: all-ones-key count, distinct, overflow
    for (u32 i = threadIdx.x; i < HC3_SLOTS; i += HC3_THREADS) { tkeys[i] = HC_EMPTY; tcnt[i] = 0; }
    for (u32 i = threadIdx.x; i < HC3_CNT_WORDS; i += HC3_THREADS) cnt16[i] = 0;
    if (threadIdx.x < 8) (&s_scal[0][0])[threadIdx.x] = 0;
    const int lane = threadIdx.x & 31;
    ull knext[HC3_PREFETCH];
    u32 b = blockIdx.x;
    u32 n_n = 0;
    if (b < nb) {
        const u32 lo_n = sub_base[b];
        n_n = sub_base[b + 1] - lo_n;
#pragma unroll
        for (int j = 0; j < HC3_PREFETCH; ++j) {
            const u32 i = j * HC3_THREADS + threadIdx.x;
            knext[j] = i < n_n ? keys2[lo_n + i] : 0ull;
        }
    }
    BLOCK_SYNC();
    u32 par = 0;
    for (; b < nb; b += gridDim.x, par ^= 1u) {
        u32* s_empty = &s_scal[par][0];
        u32* s_distinct = &s_scal[par][1];
        u32* s_overflow = &s_scal[par][2];
        const u32 n = n_n;
        const bool big = n > HC3_PREFETCH * HC3_THREADS;      // handed to the sort fallback (see hc_count2_kernel)
        ull kcur[HC3_PREFETCH];
#pragma unroll
        for (int j = 0; j < HC3_PREFETCH; ++j) kcur[j] = knext[j];
        const u32 bn = b + gridDim.x;
        if (bn < nb) {
            const u32 lo_n = sub_base[bn];
            n_n = sub_base[bn + 1] - lo_n;
#pragma unroll
            for (int j = 0; j < HC3_PREFETCH; ++j) {
                const u32 i = j * HC3_THREADS + threadIdx.x;
                knext[j] = i < n_n ? keys2[lo_n + i] : 0ull;
            }
        }
        // pass A: hashed 16-bit counters += 1 (reduction, nothing returned)
        u32 hh[HC3_PREFETCH];
        u32 live = 0;                                         // bit j: kcur[j] is a key of this bucket
#pragma unroll
        for (int j = 0; j < HC3_PREFETCH; ++j) {
            const u32 i = j * HC3_THREADS + threadIdx.x;
            hh[j] = 0;
            if (!big && i < n) {
                if (kcur[j] == HC_EMPTY) atomicAdd(s_empty, 1u);
                else {
                    const u64 prod = kcur[j] * 0xD6E8FEB86659FD93ull;
                    hh[j] = (u32)(prod >> 32);                  // top 16 bits: counter, bits 4..15: table slot
                    live |= 1u << j;
                    const u32 h = hh[j] >> 16;
                    smem_red_add(&cnt16[h >> 1], 1u << (16 * (h & 1)));
                }
            }
        }
        BLOCK_SYNC();
        // pass B: keys whose counter reached 2 are counted exactly in the table
#pragma unroll
        for (int j = 0; j < HC3_PREFETCH; ++j) {
            if (!((live >> j) & 1u)) continue;
            const u32 h = hh[j] >> 16;
            const u32 seen = (cnt16[h >> 1] >> (16 * (h & 1))) & 0xFFFFu;
            if (seen >= 2 && !*(volatile u32*)s_overflow) {
                const ull key = kcur[j];
                u32 p = (hh[j] >> 4) & (HC3_SLOTS - 1);
                while (true) {
                    ull cur = tkeys[p];
                    if (cur == HC_EMPTY) {
                        cur = atomicCAS(&tkeys[p], HC_EMPTY, key);
                        if (cur == HC_EMPTY) {
                            const u32 d = smem_atom_inc(s_distinct);
                            if (d < HC3_CLAIM_CAP) claimed[d] = (u16)p;
                            if (d >= HC3_LIMIT) *s_overflow = 1;
                            cur = key;
                        }
                    }
                    if (cur == key) { smem_red_inc(&tcnt[p]); break; }
                    p = (p + 1) & (HC3_SLOTS - 1);
                }
            }
        }
        BLOCK_SYNC();
        const bool ovf = big || *s_overflow != 0;
        const u32 nd = min(*s_distinct, (u32)HC3_CLAIM_CAP);
        const u32 n_empty = big ? 0u : *s_empty;
        if (ovf && threadIdx.x == 0) ovf_list[atomicAdd(ovf_n, 1u)] = b;
        for (u32 i0 = 0; i0 < nd + (n_empty ? 1u : 0u); i0 += HC3_THREADS) {
            const u32 i = i0 + threadIdx.x;
            ull key = 0;
            u32 cnt = 0;
            if (i < nd) {
                const u32 p = claimed[i];
                key = tkeys[p];
                cnt = tcnt[p];
                tkeys[p] = HC_EMPTY;
                tcnt[p] = 0;
            } else if (i == nd && n_empty) {
                key = HC_EMPTY;
                cnt = n_empty;
            }
            const bool keep = !ovf && cnt >= c && cnt > 0;
            __syncwarp();
            const u32 m = __ballot_sync(0xffffffffu, keep);
            if (m) {
                ull base = 0;
                const int leader = __ffs(m) - 1;
                if (lane == leader) base = atomicAdd(out_n, (ull)__popc(m));
                base = __shfl_sync(0xffffffffu, base, leader);
                if (keep) {
                    const u64 o = base + __popc(m & ((1u << lane) - 1u));
                    if (o < out_cap) { out_keys[o] = key; out_cnt[o] = cnt; }
                }
            }
        }
        // undo the counters this thread touched; reset the other parity's scalars for the next bucket
#pragma unroll
        for (int j = 0; j < HC3_PREFETCH; ++j)
            if ((live >> j) & 1u) cnt16[hh[j] >> 17] = 0;
        if (threadIdx.x < 4) s_scal[par ^ 1u][threadIdx.x] = 0;
        BLOCK_SYNC();
    }
}
