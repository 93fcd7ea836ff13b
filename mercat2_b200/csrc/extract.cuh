// K2 / K3 -- rolling k-mer extraction over the compacted symbol stream and counting.
//
// Replaces the reference's window loop  for i in range(len(seq)-k+1): kmerlist[seq[i:i+k]] += 1
// (lib/mercat2_kmers.py:56-60, :65-69).  A k-mer is ANY k consecutive non-separator bytes of `sym`
// (raw substring, no canonicalisation, no alphabet check -- SURVEY.md section 0.1).  Windows whose k
// symbols all belong to the sample's fast alphabet are encoded order-preservingly in one 64-bit word
// (2 bits ACGT / 5 bits 'A'..'Z' / 8 bits any ASCII) and counted here; every other window (it
// contains e.g. 'N' or a lower-case letter) is an "exception window" handled bit-exactly by the wide
// path (wide.cuh) under its literal bytes.
//
// Thread t of a tile owns 16 consecutive symbols.  It publishes the packed code of its own 16
// symbols and the lengths of its trailing fast / separator-free runs to shared memory, so that its
// right neighbours can start their rolling code and validity counters without re-reading symbols
// (k-1 <= 32 predecessors for 2-bit codes = two neighbours; one neighbour for 5- and 8-bit codes).
#pragma once
#include "common.cuh"

#define EX_THREADS 256
#define EX_WARPS (EX_THREADS / 32)
#define EX_TILE (EX_THREADS * 16)
#define EX_HALO 8                    // halo threads left of a tile: 128 symbols >= k-1 for k <= 128

enum : int { ENC_NT2 = 0, ENC_AA5 = 1, ENC_BYTE = 2 };

template <int ENC> struct EncTraits;
template <> struct EncTraits<ENC_NT2> {
    static constexpr int BITS = 2;
    __device__ static __forceinline__ bool fast(u32 c) { return (c - 65u) < 26u && ((0x00080045u >> (c - 65u)) & 1u); }
    __device__ static __forceinline__ u32 digit(u32 c) { return ((c >> 1) ^ (c >> 2)) & 3u; }   // A,C,G,T -> 0,1,2,3
};
template <> struct EncTraits<ENC_AA5> {
    static constexpr int BITS = 5;
    __device__ static __forceinline__ bool fast(u32 c) { return (c - 65u) < 26u; }
    __device__ static __forceinline__ u32 digit(u32 c) { return c - 65u; }
};
template <> struct EncTraits<ENC_BYTE> {
    static constexpr int BITS = 8;
    __device__ static __forceinline__ bool fast(u32 c) { return c < 128u; }
    __device__ static __forceinline__ u32 digit(u32 c) { return c; }
};

struct SymView {
    const u8* sym;   // 16-byte aligned
    u64 n;           // symbols
    // batched samples (mc2_count_batch): sample j owns the symbols [sample_start[j], sample_start[j + 1]); a window's key
    // carries its sample in the bits above the k-mer code, so that one pass counts every (sample, k-mer) pair
    const u64* sample_start = nullptr;
    u32 n_samples = 0;
    u32 sample_shift = 0;
};

// sample of a symbol position, for positions visited in ascending order by one thread
struct SampleTag {
    u32 sid;
    u64 next;
    __device__ __forceinline__ void init(const SymView& v, u64 s) {
        sid = 0;
        next = ~0ull;
        if (!v.sample_start) return;
        u32 lo = 0, hi = v.n_samples;                       // last j with sample_start[j] <= s
        while (hi - lo > 1) { const u32 mid = (lo + hi) >> 1; if (v.sample_start[mid] <= s) lo = mid; else hi = mid; }
        sid = lo;
        next = lo + 1 < v.n_samples ? v.sample_start[lo + 1] : ~0ull;
    }
    __device__ __forceinline__ u64 tag(const SymView& v, u64 s) {
        while (s >= next) { ++sid; next = sid + 1 < v.n_samples ? v.sample_start[sid + 1] : ~0ull; }
        return (u64)sid << v.sample_shift;
    }
};

__device__ __forceinline__ void load_sym16(const SymView& v, i64 p0, u32 w[4]) {
    if (p0 >= 0 && (u64)p0 + 16 <= v.n) {
        const uint4 q = *reinterpret_cast<const uint4*>(v.sym + p0);
        w[0] = q.x; w[1] = q.y; w[2] = q.z; w[3] = q.w;
    } else {
        w[0] = w[1] = w[2] = w[3] = 0xFFFFFFFFu;            // separators outside the stream
        if (p0 + 16 > 0 && p0 < (i64)v.n) {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const i64 p = p0 + i;
                if (p >= 0 && p < (i64)v.n) {
                    const u32 c = v.sym[p];
                    w[i >> 2] = (w[i >> 2] & ~(0xFFu << (8 * (i & 3)))) | (c << (8 * (i & 3)));
                }
            }
        }
    }
}

// summary of 16 symbols: packed code of the 16 digits (low 64 bits) and the trailing run lengths
template <int ENC>
__device__ __forceinline__ void summarize_sym16(const u32 w[4], u64& code, u32& meta) {
    typedef EncTraits<ENC> E;
    u64 c64 = 0;
    u32 tf = 0, ts = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const u32 c = (w[i >> 2] >> (8 * (i & 3))) & 0xFFu;
        const bool f = E::fast(c);
        c64 = (c64 << E::BITS) | (f ? E::digit(c) : 0u);
        tf = f ? tf + 1 : 0;
        ts = (c != MC2_SEP) ? ts + 1 : 0;
    }
    code = c64;
    meta = tf | (ts << 8);
}

// Per-tile context: every thread gets its 16 symbols plus the rolling state inherited from the left.
template <int ENC>
struct TileCtx {
    u32 w[4];
    u64 code;        // rolling code before the thread's first symbol
    u32 run_fast;    // consecutive fast symbols ending just before the thread's first symbol (saturated)
    u32 run_nosep;   // same for non-separator symbols
};

// smem: s_code[EX_THREADS + EX_HALO], s_meta[EX_THREADS + EX_HALO]
template <int ENC>
__device__ __forceinline__ void tile_begin(const SymView& v, u64 tile_start, int k, u64* s_code, u32* s_meta, TileCtx<ENC>& ctx) {
    const int t = threadIdx.x;
    load_sym16(v, (i64)tile_start + 16 * t, ctx.w);
    u64 code; u32 meta;
    summarize_sym16<ENC>(ctx.w, code, meta);
    s_code[t + EX_HALO] = code;
    s_meta[t + EX_HALO] = meta;
    if (t < EX_HALO) {                             // the 128 symbols left of the tile
        u32 hw[4];
        load_sym16(v, (i64)tile_start - 16 * (EX_HALO - t), hw);
        u64 hc; u32 hm;
        summarize_sym16<ENC>(hw, hc, hm);
        s_code[t] = hc;
        s_meta[t] = hm;
    }
    BLOCK_SYNC();
    const u64 c1 = s_code[t + EX_HALO - 1], c2 = s_code[t + EX_HALO - 2];
    ctx.code = (EncTraits<ENC>::BITS == 2) ? ((c2 << 32) | (c1 & 0xFFFFFFFFull)) : c1;
    // run lengths ending just before this thread: walk left while whole 16-symbol groups qualify
    const int need = min(EX_HALO, (k + 14) / 16);          // ceil((k-1)/16) groups can matter
    u32 rf = 0, rs = 0;
    bool open_f = true, open_s = true;
    for (int j = 1; j <= need; ++j) {
        const u32 m = s_meta[t + EX_HALO - j];
        const u32 f = m & 0xFFu, n = m >> 8;
        if (open_f) { rf += f; open_f = f == 16; }
        if (open_s) { rs += n; open_s = n == 16; }
    }
    ctx.run_fast = rf;
    ctx.run_nosep = rs;
}

// Walk the thread's 16 symbols.  f(i, code, is_fast_window, is_exception_window) is called for every
// position i whose window (the k symbols ENDING at symbol i) is separator-free.
template <int ENC, class F>
__device__ __forceinline__ void tile_walk(TileCtx<ENC>& ctx, int k, F&& f) {
    typedef EncTraits<ENC> E;
    u64 code = ctx.code;
    u32 rf = ctx.run_fast, rs = ctx.run_nosep;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const u32 c = (ctx.w[i >> 2] >> (8 * (i & 3))) & 0xFFu;
        const bool fast = E::fast(c);
        code = (code << E::BITS) | (fast ? E::digit(c) : 0u);
        rf = fast ? min(rf + 1, 255u) : 0u;
        rs = (c != MC2_SEP) ? min(rs + 1, 255u) : 0u;
        if (rs >= (u32)k) f(i, code, rf >= (u32)k);
    }
}

// ---- dense tables ---------------------------------------------------------------------------------
// bin index of a fast window: NT2 -> the 2k-bit code itself (4^k bins); AA5 -> base-26 Horner value of
// the k 5-bit digits (26^k bins).  Both orders equal byte order of the k-mer text.
template <int ENC>
__device__ __forceinline__ u32 dense_index(u64 code, int k, u64 mask) {
    if (ENC == ENC_NT2) return (u32)(code & mask);
    u32 idx = 0;
    for (int j = k - 1; j >= 0; --j) idx = idx * 26u + (u32)((code >> (5 * j)) & 31u);
    return idx;
}

// persistent CTAs; shared-memory privatised histogram (nrep interleaved copies), flushed once per CTA
template <int ENC>
__global__ void __launch_bounds__(EX_THREADS)
dense_smem_kernel(SymView v, u64 s0, u64 s1, int k, u32 bins, u32 nrep, u32* __restrict__ table) {
    extern __shared__ __align__(16) u8 dyn[];
    u32* hist = reinterpret_cast<u32*>(dyn);                 // [nrep][bins]
    __shared__ u64 s_code[EX_THREADS + EX_HALO];
    __shared__ u32 s_meta[EX_THREADS + EX_HALO];
    for (u32 i = threadIdx.x; i < bins * nrep; i += EX_THREADS) hist[i] = 0;
    BLOCK_SYNC();
    const u64 mask = (2 * k >= 64) ? ~0ull : ((1ull << (2 * k)) - 1);
    u32* my = hist + (u32)((threadIdx.x >> 5) % nrep) * bins;
    const u64 ntiles = (s1 - s0 + EX_TILE - 1) / EX_TILE;
    for (u64 tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const u64 tile_start = s0 + tile * EX_TILE;
        TileCtx<ENC> ctx;
        tile_begin<ENC>(v, tile_start, k, s_code, s_meta, ctx);
        const u64 first = tile_start + 16ull * threadIdx.x;
        tile_walk<ENC>(ctx, k, [&](int i, u64 code, bool fast) {
            if (fast && first + i < s1) atomicAdd(&my[dense_index<ENC>(code, k, mask)], 1u);
        });
        BLOCK_SYNC();                                     // s_code/s_meta reuse
    }
    BLOCK_SYNC();
    for (u32 b = threadIdx.x; b < bins; b += EX_THREADS) {
        u32 sum = 0;
        for (u32 r = 0; r < nrep; ++r) sum += hist[r * bins + b];
        if (sum) atomicAdd(&table[b], sum);
    }
}

// larger tables: counts go straight to the (L2-resident) global table with reduction atomics
template <int ENC>
__global__ void __launch_bounds__(EX_THREADS)
dense_global_kernel(SymView v, u64 s0, u64 s1, int k, u32* __restrict__ table) {
    __shared__ u64 s_code[EX_THREADS + EX_HALO];
    __shared__ u32 s_meta[EX_THREADS + EX_HALO];
    const u64 mask = (2 * k >= 64) ? ~0ull : ((1ull << (2 * k)) - 1);
    const u64 tile_start = s0 + (u64)blockIdx.x * EX_TILE;
    TileCtx<ENC> ctx;
    tile_begin<ENC>(v, tile_start, k, s_code, s_meta, ctx);
    const u64 first = tile_start + 16ull * threadIdx.x;
    tile_walk<ENC>(ctx, k, [&](int i, u64 code, bool fast) {
        if (fast && first + i < s1) atomicAdd(&table[dense_index<ENC>(code, k, mask)], 1u);
    });
}

// sample[b] += chunk[b] >= c ? chunk[b] : 0 ; chunk[b] = 0     (lib/mercat2_kmers.py:73-78 then
// bin/mercat2.py:121-127: the filter is per chunk file, the sum per sample)
__global__ void dense_fold_kernel(u32* __restrict__ chunk, u64* __restrict__ sample, u32 bins, u64 c) {
    const u32 b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < bins) {
        const u32 n = chunk[b];
        if (n) { if (n >= c) sample[b] += n; chunk[b] = 0; }
    }
}
// multi-batch chunks: fold a u32 batch table into a u64 chunk table (no filter), then filter later
__global__ void dense_fold_batch_kernel(u32* __restrict__ batch, u64* __restrict__ chunk64, u32 bins) {
    const u32 b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < bins) { const u32 n = batch[b]; if (n) { chunk64[b] += n; batch[b] = 0; } }
}
__global__ void dense_fold64_kernel(u64* __restrict__ chunk64, u64* __restrict__ sample, u32 bins, u64 c) {
    const u32 b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < bins) { const u64 n = chunk64[b]; if (n) { if (n >= c) sample[b] += n; chunk64[b] = 0; } }
}

__global__ void __launch_bounds__(256) dense_nonzero_count_kernel(const u64* __restrict__ sample, u32 bins, u32* __restrict__ tile_cnt) {
    const u32 b = blockIdx.x * 256 + threadIdx.x;
    const u32 total = block_count(b < bins && sample[b] != 0);
    if (threadIdx.x == 0) tile_cnt[blockIdx.x] = total;
}
__global__ void __launch_bounds__(256)
dense_nonzero_write_kernel(const u64* __restrict__ sample, u32 bins, const u64* __restrict__ tile_off, u64* __restrict__ keys, u64* __restrict__ counts) {
    __shared__ u32 sm[256 / 32 + 1];
    const u32 b = blockIdx.x * 256 + threadIdx.x;
    const bool nz = b < bins && sample[b] != 0;
    const u32 off = block_exclusive_scan<OpAdd, 8>(nz ? 1u : 0u, sm, nullptr);
    if (nz) { const u64 u = tile_off[blockIdx.x] + off; keys[u] = b; counts[u] = sample[b]; }
}

// ---- sparse: materialise the 64-bit keys of all fast windows (compacted, CTA order arbitrary) -------
template <int ENC>
__global__ void __launch_bounds__(EX_THREADS)
extract_keys_kernel(SymView v, u64 s0, u64 s1, int k, u64* __restrict__ keys, ull* __restrict__ nkeys) {
    __shared__ u64 s_code[EX_THREADS + EX_HALO];
    __shared__ u32 s_meta[EX_THREADS + EX_HALO];
    __shared__ u64 s_keys[EX_TILE];
    __shared__ u32 sm[EX_WARPS + 1];
    __shared__ ull s_base;
    const int kb = k * EncTraits<ENC>::BITS;
    const u64 mask = kb >= 64 ? ~0ull : ((1ull << kb) - 1);
    const u64 tile_start = s0 + (u64)blockIdx.x * EX_TILE;
    TileCtx<ENC> ctx;
    tile_begin<ENC>(v, tile_start, k, s_code, s_meta, ctx);
    const u64 first = tile_start + 16ull * threadIdx.x;
    u64 mine[16];
    u32 valid = 0;
    tile_walk<ENC>(ctx, k, [&](int i, u64 code, bool fast) {
        mine[i] = code & mask;
        if (fast && first + i < s1) valid |= 1u << i;
    });
    u32 total;
    u32 off = block_exclusive_scan<OpAdd, EX_WARPS>((u32)__popc(valid), sm, &total);
#pragma unroll
    for (int i = 0; i < 16; ++i) if ((valid >> i) & 1u) s_keys[off++] = mine[i];
    if (threadIdx.x == 0) s_base = total ? atomicAdd(nkeys, (ull)total) : 0ull;
    BLOCK_SYNC();
    const u64 base = s_base;
    for (u32 i = threadIdx.x; i < total; i += EX_THREADS) keys[base + i] = s_keys[i];
}

// positions (window START index in sym) of windows selected by MODE:
//   MODE 0: exception windows of encoding ENC (separator-free but not all-fast)
//   MODE 1: every separator-free window (used when the whole sample runs on the wide path)
template <int ENC, int MODE>
__global__ void __launch_bounds__(EX_THREADS)
extract_positions_kernel(SymView v, u64 s0, u64 s1, int k, u64* __restrict__ pos, u64 cap, ull* __restrict__ npos) {
    __shared__ u64 s_code[EX_THREADS + EX_HALO];
    __shared__ u32 s_meta[EX_THREADS + EX_HALO];
    __shared__ u32 sm[EX_WARPS + 1];
    __shared__ ull s_base;
    const u64 tile_start = s0 + (u64)blockIdx.x * EX_TILE;
    TileCtx<ENC> ctx;
    tile_begin<ENC>(v, tile_start, k, s_code, s_meta, ctx);
    const u64 first = tile_start + 16ull * threadIdx.x;
    u32 sel = 0;
    tile_walk<ENC>(ctx, k, [&](int i, u64 code, bool fast) {
        if ((MODE == 1 || !fast) && first + i < s1) sel |= 1u << i;
    });
    u32 total;
    u32 off = block_exclusive_scan<OpAdd, EX_WARPS>((u32)__popc(sel), sm, &total);
    if (threadIdx.x == 0) s_base = total ? atomicAdd(npos, (ull)total) : 0ull;
    BLOCK_SYNC();
    const u64 base = s_base;
#pragma unroll
    for (int i = 0; i < 16; ++i)
        if ((sel >> i) & 1u) {
            const u64 u = base + off++;
            if (u < cap) pos[u] = first + i + 1 - (u64)k;
        }
}

