// Summing (key, count) rows by key on the range path -- the dict merge of bin/mercat2.py:121-127 (already filtered tables
// of a sample's pieces, of other ranks, of overflow fallbacks) without sorting.
//
// Round 1 merged such rows with an 8-pass LSD radix sort + reduce-by-key (two 16-byte row arrays crossing HBM eight
// times).  Here the rows run through the same order-preserving two-level partition as the keys of a chunk
// (rangecount.cuh: sampled LUT, exact histogram, two coalesced scatters -- each carrying the count beside the key), and a
// counting kernel sums the rows of one sub-bucket in a shared-memory table (64-bit sums) and emits them in key order.
// Every row crosses HBM three times instead of sixteen, and the result is born sorted.
#pragma once
#include "rangecount.cuh"

#define HKV_STAGE_V_OFFSET (HC_SCATTER_SMEM16)                                   // counts staged behind keys + digits
#define HKV_SCATTER_SMEM (HC_SCATTER_SMEM16 + (((size_t)HC_STAGE_SLOTS * 8 + 15) & ~(size_t)15))
#define HKV_SCATTER_SMEM_LUT (HKV_SCATTER_SMEM + (size_t)RP_LUT * 2)

// hc_group_and_write with a 64-bit payload per key
template <bool FULL, class KeyFn, class ValFn, class DigitFn>
__device__ __forceinline__ void hkv_group_and_write(KeyFn mine, ValFn val, DigitFn dig, u32 valid, u32 nd, u64* stage, u64* stagev, u16* sdig,
                                                    u32* cnt, u32* loff, u32* gbase, u32* sm, u32* __restrict__ cursors,
                                                    u64* __restrict__ out, u64* __restrict__ outv) {
    u32 rd[16];
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
        stagev[pos] = val(i);
        sdig[pos] = (u16)d;
    }
    BLOCK_SYNC();
    for (u32 i = threadIdx.x; i < total; i += EX_THREADS) {
        const u32 d = sdig[i];
        const u64 at = (u64)gbase[d] + (i - loff[d]);
        out[at] = stage[i];
        outv[at] = stagev[i];
    }
}

__global__ void __launch_bounds__(EX_THREADS, 2)
hkv_scatter1_kernel(const u64* __restrict__ keys, const u64* __restrict__ vals, u64 n, RpView r, u32* __restrict__ cur1,
                    u64* __restrict__ keys1, u64* __restrict__ vals1) {
    extern __shared__ __align__(16) u8 dyn_sc[];
    u64* stage = reinterpret_cast<u64*>(dyn_sc);
    u16* sdig = reinterpret_cast<u16*>(dyn_sc + HC_SDST_OFFSET);
    u64* stagev = reinterpret_cast<u64*>(dyn_sc + HKV_STAGE_V_OFFSET);
    u16* s_lut = reinterpret_cast<u16*>(dyn_sc + HKV_SCATTER_SMEM);
    __shared__ u32 cnt[HC_MAX_NB1], loff[HC_MAX_NB1], gbase[HC_MAX_NB1];
    __shared__ u32 sm[EX_WARPS + 1];
    for (u32 i = threadIdx.x; i < RP_LUT / 8; i += EX_THREADS) reinterpret_cast<uint4*>(s_lut)[i] = reinterpret_cast<const uint4*>(r.lut)[i];
    const u64 ntiles = (n + HC_TILE - 1) / HC_TILE;
    for (u64 tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        for (u32 i = threadIdx.x; i < r.nb1; i += EX_THREADS) cnt[i] = 0;
        const u64 base = tile * HC_TILE;
        u64 mine[16], mv[16];
        u32 valid = 0;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const u64 i = base + (u64)j * EX_THREADS + threadIdx.x;
            mine[j] = 0;
            mv[j] = 0;
            if (i < n) { mine[j] = keys[i]; mv[j] = vals[i]; valid |= 1u << j; }
        }
        BLOCK_SYNC();
        auto key = [&](int i) { return mine[i]; };
        auto val = [&](int i) { return mv[i]; };
        auto dig = [&](int i) { return ((valid >> i) & 1u) ? (u32)s_lut[rp_lut_index(rp_prefix(mine[i], r.down, r.up), r.base, r.sh)] : 0u; };
        if (base + HC_TILE <= n) hkv_group_and_write<true>(key, val, dig, valid, r.nb1, stage, stagev, sdig, cnt, loff, gbase, sm, cur1, keys1, vals1);
        else hkv_group_and_write<false>(key, val, dig, valid, r.nb1, stage, stagev, sdig, cnt, loff, gbase, sm, cur1, keys1, vals1);
        BLOCK_SYNC();
    }
}

__global__ void __launch_bounds__(EX_THREADS, 2)
hkv_scatter2_kernel(const u64* __restrict__ keys1, const u64* __restrict__ vals1, const u32* __restrict__ sub_base, const u32* __restrict__ tile_pref,
                    u32 nb, u32 nb2, RpView r, u32* __restrict__ cur2, u64* __restrict__ keys2, u64* __restrict__ vals2) {
    extern __shared__ __align__(16) u8 dyn_sc[];
    u64* stage = reinterpret_cast<u64*>(dyn_sc);
    u16* sdig = reinterpret_cast<u16*>(dyn_sc + HC_SDST_OFFSET);
    u64* stagev = reinterpret_cast<u64*>(dyn_sc + HKV_STAGE_V_OFFSET);
    __shared__ u32 cnt[HC_NB2], loff[HC_NB2], gbase[HC_NB2];
    __shared__ u32 sm[EX_WARPS + 1];
    __shared__ u32 s_b1;
    const u32 nb1 = r.nb1;
    if (blockIdx.x >= tile_pref[nb1]) return;
    if (threadIdx.x == 0) {
        u32 lo = 0, hi = nb1;
        while (hi - lo > 1) { const u32 mid = (lo + hi) / 2; if (tile_pref[mid] <= blockIdx.x) lo = mid; else hi = mid; }
        s_b1 = lo;
    }
    for (u32 i = threadIdx.x; i < nb2; i += EX_THREADS) cnt[i] = 0;
    BLOCK_SYNC();
    const u32 b1 = s_b1;
    const uint2 d1 = r.l1[b1];
    const u32 lo = sub_base[b1 * nb2], hi = sub_base[min(nb, (b1 + 1) * nb2)];
    const u32 base = lo + (blockIdx.x - tile_pref[b1]) * HC_TILE;
    u64 mine[16], mv[16];
    u32 valid = 0;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        const u32 i = base + j * EX_THREADS + threadIdx.x;
        mine[j] = 0;
        mv[j] = 0;
        if (i < hi) { mine[j] = keys1[i]; mv[j] = vals1[i]; valid |= 1u << j; }
    }
    auto key = [&](int i) { return mine[i]; };
    auto val = [&](int i) { return mv[i]; };
    auto dig = [&](int i) { return ((valid >> i) & 1u) ? rp_b2(d1, rp_prefix(mine[i], r.down, r.up)) : 0u; };
    if (base + HC_TILE <= hi) hkv_group_and_write<true>(key, val, dig, valid, nb2, stage, stagev, sdig, cnt, loff, gbase, sm, cur2 + b1 * nb2, keys2, vals2);
    else hkv_group_and_write<false>(key, val, dig, valid, nb2, stage, stagev, sdig, cnt, loff, gbase, sm, cur2 + b1 * nb2, keys2, vals2);
}

// ---- rc_merge: one sub-bucket of rows per CTA iteration: sum by key in shared memory, emit in key order ------------------
#define RM_THREADS 512
#define RM_SLOTS 4096u
#define RM_SMEM ((size_t)RM_SLOTS * 16 + (size_t)RM_SLOTS * 2 + (size_t)RC_FINE * 4)

__global__ void __launch_bounds__(RM_THREADS, 2)
rc_merge_kernel(const u64* __restrict__ keys2, const u64* __restrict__ vals2, const u32* __restrict__ sub_base, u32 nb, RpView r,
                RcRow* __restrict__ out, u32* __restrict__ rows, u32* __restrict__ ovf_n) {
    extern __shared__ __align__(16) u8 dyn[];
    ull* tkeys = reinterpret_cast<ull*>(dyn);
    ull* tsum = reinterpret_cast<ull*>(dyn + (size_t)RM_SLOTS * 8);
    u16* sidx = reinterpret_cast<u16*>(dyn + (size_t)RM_SLOTS * 16);                 // slots in output order
    u32* bins = reinterpret_cast<u32*>(dyn + (size_t)RM_SLOTS * 18);
    __shared__ u32 s_warp[RM_THREADS / 32 + 1];
    __shared__ u32 s_ovf;
    __shared__ ull s_empty;                                                         // summed count of the all-ones key
    for (u32 i = threadIdx.x; i < RM_SLOTS; i += RM_THREADS) { tkeys[i] = HC_EMPTY; tsum[i] = 0; }
    for (u32 i = threadIdx.x; i < RC_FINE; i += RM_THREADS) bins[i] = 0;
    if (threadIdx.x == 0) { s_ovf = 0; s_empty = 0; }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    BLOCK_SYNC();
    for (u32 b = blockIdx.x; b < nb; b += gridDim.x) {
        const u32 lo = sub_base[b], n = sub_base[b + 1] - lo;
        const uint2 d1 = __ldg(&r.l1[b >> HC_NB2_LOG2]);
        auto fine = [&](ull key) { return rp_fine(d1, rp_prefix(key, r.down, r.up), RC_FINE_LOG2); };
        // rounds of 4 rows per thread: the loads of a round are in flight together (one load at a time left the kernel
        // waiting on memory: 3 G rows/s)
        for (u32 i0 = 0; i0 < n; i0 += 4 * RM_THREADS) {
            ull kk[4], ww[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const u32 i = i0 + j * RM_THREADS + threadIdx.x;
                kk[j] = i < n ? keys2[lo + i] : 0ull;
                ww[j] = i < n ? vals2[lo + i] : 0ull;
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const u32 i = i0 + j * RM_THREADS + threadIdx.x;
                if (i >= n) continue;
                const ull key = kk[j], w = ww[j];
                if (key == HC_EMPTY) { atomicAdd(&s_empty, w); continue; }
                u32 p = rc_slot(rc_hash(key));
                u32 probes = 0;
#pragma unroll 1
                for (; probes < RC_PROBE_LIMIT; ++probes) {
                    ull cur = tkeys[p];
                    if (cur == HC_EMPTY) {
                        cur = atomicCAS(&tkeys[p], HC_EMPTY, key);
                        if (cur == HC_EMPTY) cur = key;
                    }
                    if (cur == key) { atomicAdd(&tsum[p], w); break; }
                    p = (p + 1) & (RM_SLOTS - 1);
                }
                if (probes == RC_PROBE_LIMIT) s_ovf = 1;
            }
        }
        BLOCK_SYNC();
        const bool ovf = s_ovf != 0;
        const ull n_empty = s_empty;
        u32 nsurv = 0;
        if (!ovf) {
            for (u32 p = threadIdx.x; p < RM_SLOTS; p += RM_THREADS)
                if (tkeys[p] != HC_EMPTY) smem_red_inc(&bins[fine(tkeys[p])]);
            BLOCK_SYNC();
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
                u32 w = lane < RM_THREADS / 32 ? s_warp[lane] : 0u;
                u32 wi = w;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const u32 y = __shfl_up_sync(0xffffffffu, wi, d);
                    if (lane >= d) wi += y;
                }
                if (lane < RM_THREADS / 32) s_warp[lane] = wi - w;
                if (lane == RM_THREADS / 32 - 1) s_warp[RM_THREADS / 32] = wi;
            }
            BLOCK_SYNC();
            nsurv = s_warp[RM_THREADS / 32];
            const u32 ex = s_warp[warp] + incl - (c0 + c1);
            bins[2 * threadIdx.x] = ex;
            bins[2 * threadIdx.x + 1] = ex + c0;
            BLOCK_SYNC();
            for (u32 p = threadIdx.x; p < RM_SLOTS; p += RM_THREADS)
                if (tkeys[p] != HC_EMPTY) sidx[smem_atom_inc(&bins[fine(tkeys[p])])] = (u16)p;
            BLOCK_SYNC();
            RcRow* dst = out + lo;                                                  // (a bucket of n rows has at most n distinct keys)
            for (u32 i = threadIdx.x; i < nsurv; i += RM_THREADS) {
                const u32 p = sidx[i];
                const ull key = tkeys[p];
                const u32 d = fine(key);
                const u32 s0 = d ? bins[d - 1] : 0u, s1 = bins[d];
                u32 rank = 0;
                for (u32 j = s0; j < s1; ++j) rank += tkeys[sidx[j]] < key ? 1u : 0u;
                RcRow row;
                row.key = key;
                row.count = tsum[p];
                dst[s0 + rank] = row;
            }
            if (n_empty && threadIdx.x == 0) {
                RcRow row;
                row.key = HC_EMPTY;
                row.count = n_empty;
                dst[nsurv] = row;
            }
        }
        if (threadIdx.x == 0) {
            rows[b] = ovf ? 0u : nsurv + (n_empty ? 1u : 0u);
            if (ovf) atomicAdd(ovf_n, 1u);
        }
        BLOCK_SYNC();
        for (u32 p = threadIdx.x; p < RM_SLOTS; p += RM_THREADS) { tkeys[p] = HC_EMPTY; tsum[p] = 0; }
        for (u32 i = threadIdx.x; i < RC_FINE; i += RM_THREADS) bins[i] = 0;
        if (threadIdx.x == 0) { s_ovf = 0; s_empty = 0; }
        BLOCK_SYNC();
    }
}
