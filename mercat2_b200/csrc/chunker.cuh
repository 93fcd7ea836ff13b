// A2 -- virtual Chunker: the split points of lib/mercat2_Chunker.py:39-59 computed on device, without
// writing piece files.
//
// Reference rule (stream_delim): walking the text line by line (text mode, so "\r\n" has already
// become "\n"), at every line that CONTAINS '>' the size of the current piece is compared with
// chunksize; if size >= chunksize the line starts a new piece.  Piece size = bytes written so far =
// translated length, i.e. raw length minus one per "\r\n" pair.
//
// Device formulation: a "candidate" is the first '>' of a line; its line start `ls` is found by
// walking left from the '>' until a terminator (candidate) or another '>' (not the first: dropped),
// so no cross-tile state is needed.  T(ls) = ls - #{"\r\n" pairs before ls} is the translated
// offset; candidates come out ordered, T is monotonic, and the sequential rule becomes a chain of
// lower_bound searches done by one thread (one search per piece).
#pragma once
#include "common.cuh"

#define CH_THREADS 256
#define CH_ITEMS 16
#define CH_TILE (CH_THREADS * CH_ITEMS)

__device__ __forceinline__ bool ch_is_nl(u32 c) { return c == 10u || c == 13u; }

// returns true and the line start if text[p] == '>' is the first '>' of its line
__device__ __forceinline__ bool ch_candidate(const u8* __restrict__ text, u64 p, u64& ls) {
    u64 q = p;
    while (q > 0) {
        const u32 c = text[q - 1];
        if (ch_is_nl(c)) break;
        if (c == '>') return false;
        --q;
    }
    ls = q;
    return true;
}

__device__ __forceinline__ u32 ch_nz(u32 x) { return (((x & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | x) & 0x80808080u; }
__device__ __forceinline__ u32 ch_eq(u32 w, u32 c4) { return ~ch_nz(w ^ c4) & 0x80808080u; }
__device__ __forceinline__ u32 ch_movemask(u32 m) { return (((m >> 7) * 0x00204081u) >> 21) & 0xFu; }

// Each thread owns 16 bytes (aligned 16-byte loads; `text` itself may be misaligned).  Groups without
// '>' and '\r' -- nearly all of them -- cost two SWAR tests per word.
template <bool WRITE>
__global__ void __launch_bounds__(CH_THREADS)
chunk_candidates_kernel(const u8* __restrict__ text, u64 n, u32* __restrict__ tile_crlf, u32* __restrict__ tile_cand,
                        const u64* __restrict__ crlf_off, const u64* __restrict__ cand_off,
                        u64* __restrict__ cand_ls, u64* __restrict__ cand_t) {
    __shared__ u32 sm[CH_THREADS / 32 + 1];
    const u64 mis = (u64)(uintptr_t)text & 15ull;
    const u8* aligned = text - mis;
    const u64 v0 = (u64)blockIdx.x * CH_TILE + (u64)threadIdx.x * CH_ITEMS;      // virtual position of byte 0
    u32 w[4] = {0, 0, 0, 0};
    if (v0 >= mis && v0 + 16 <= mis + n) {
        const uint4 q = *reinterpret_cast<const uint4*>(aligned + v0);
        w[0] = q.x; w[1] = q.y; w[2] = q.z; w[3] = q.w;
    } else if (v0 + 16 > mis && v0 < mis + n) {
        for (int i = 0; i < 16; ++i) {
            const u64 v = v0 + i;
            if (v >= mis && v < mis + n) w[i >> 2] |= (u32)aligned[v] << (8 * (i & 3));
        }
    }
    u32 gt = 0, cr = 0;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        gt |= ch_movemask(ch_eq(w[q], 0x3E3E3E3Eu)) << (4 * q);
        cr |= ch_movemask(ch_eq(w[q], 0x0D0D0D0Du)) << (4 * q);
    }
    u32 ncrlf = 0, ncand = 0, candmask = 0, crlfmask = 0;
    u64 ls_arr[CH_ITEMS];
    u32 m = cr;
    while (m) {                                               // "\r\n" pairs (the '\n' may be in the next group)
        const int j = __ffs(m) - 1;
        m &= m - 1;
        const u64 p = v0 + j - mis;
        if (p + 1 < n && text[p + 1] == 10u) { ncrlf++; crlfmask |= 1u << j; }
    }
    m = gt;
    while (m) {
        const int j = __ffs(m) - 1;
        m &= m - 1;
        u64 ls;
        if (ch_candidate(text, v0 + j - mis, ls)) { ncand++; candmask |= 1u << j; ls_arr[j] = ls; }
    }
    if (!WRITE) {
        u32 tc, tk;
        block_exclusive_scan<OpAdd, CH_THREADS / 32>(ncrlf, sm, &tc);
        block_exclusive_scan<OpAdd, CH_THREADS / 32>(ncand, sm, &tk);
        if (threadIdx.x == 0) { tile_crlf[blockIdx.x] = tc; tile_cand[blockIdx.x] = tk; }
    } else {
        u64 crlf_before = crlf_off[blockIdx.x] + block_exclusive_scan<OpAdd, CH_THREADS / 32>(ncrlf, sm, nullptr);
        u64 u = cand_off[blockIdx.x] + block_exclusive_scan<OpAdd, CH_THREADS / 32>(ncand, sm, nullptr);
        u32 both = candmask | crlfmask;
        while (both) {
            const int j = __ffs(both) - 1;
            both &= both - 1;
            if ((candmask >> j) & 1u) {
                // no "\r\n" pair lies between the line start and the '>' (no terminators there)
                cand_ls[u] = ls_arr[j];
                cand_t[u] = ls_arr[j] - crlf_before;
                ++u;
            }
            if ((crlfmask >> j) & 1u) ++crlf_before;
        }
    }
}

// one thread: bounds[0] = 0; then repeatedly the first candidate whose piece size reached chunk_bytes
__global__ void chunk_select_kernel(const u64* __restrict__ cand_ls, const u64* __restrict__ cand_t, u64 ncand,
                                    u64 chunk_bytes, u64* __restrict__ bounds, u64 max_bounds, ull* __restrict__ nbounds) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    u64 nb = 0;
    bounds[nb++] = 0;
    u64 piece_t = 0;        // translated offset where the current piece starts
    u64 from = 0;
    while (from < ncand) {
        const u64 target = piece_t + chunk_bytes;
        u64 lo = from, hi = ncand;               // first j in [from, ncand) with cand_t[j] >= target
        while (lo < hi) {
            const u64 mid = lo + (hi - lo) / 2;
            if (cand_t[mid] >= target) hi = mid; else lo = mid + 1;
        }
        if (lo >= ncand) break;
        if (cand_ls[lo] != 0) {                  // a piece boundary at offset 0 is the file start itself
            if (nb < max_bounds) bounds[nb] = cand_ls[lo];
            nb++;
        }
        piece_t = cand_t[lo];
        from = lo + 1;
    }
    *nbounds = nb;
}

// ---- fast path for text without '\r' ---------------------------------------------------------------------------
// Without "\r\n" pairs the translated offset of the reference (text-mode read) equals the raw offset, so the
// boundary after b is simply the first line that starts at or after b + chunk_bytes and contains '>'.  One kernel
// checks that no '\r' exists; one CTA then walks the whole boundary chain with block-parallel searches (a few
// hundred bytes per boundary for read files) instead of classifying every byte of the text twice.
__global__ void __launch_bounds__(256) chunk_has_cr_kernel(const u8* __restrict__ text, u64 n, u32* __restrict__ flag) {
    const u64 mis = (16 - ((u64)(uintptr_t)text & 15ull)) & 15ull;          // bytes before the first aligned 16
    bool hit = false;
    const u64 gid = (u64)blockIdx.x * blockDim.x + threadIdx.x, gsz = (u64)gridDim.x * blockDim.x;
    if (gid < mis && gid < n) hit |= text[gid] == 13u;
    const u64 nvec = n > mis ? (n - mis) / 16 : 0;
    const uint4* v = reinterpret_cast<const uint4*>(text + mis);
    for (u64 i = gid; i < nvec; i += gsz) {
        const uint4 q = v[i];
        hit |= (ch_eq(q.x, 0x0D0D0D0Du) | ch_eq(q.y, 0x0D0D0D0Du) | ch_eq(q.z, 0x0D0D0D0Du) | ch_eq(q.w, 0x0D0D0D0Du)) != 0;
    }
    const u64 tail = mis + nvec * 16;
    if (tail + gid < n) hit |= text[tail + gid] == 13u;                      // fewer than 16 tail bytes
    if (hit) atomicOr(flag, 1u);
}

#define CH_NONE 0xFFFFFFFFFFFFFFFFull
// first position in [start, n) whose byte satisfies MODE (0: terminator, 1: '>'); all threads call
template <int MODE>
__device__ __forceinline__ u64 ch_find_first(const u8* __restrict__ text, u64 start, u64 n, ull* s_pos) {
    for (u64 base = start; base < n; base += 256 * 8) {
        if (threadIdx.x == 0) *s_pos = CH_NONE;
        BLOCK_SYNC();
        const u64 a = base + (u64)threadIdx.x * 8;
        u64 found = CH_NONE;
        for (int j = 7; j >= 0; --j) {
            const u64 p = a + j;
            if (p < n) {
                const u32 c = text[p];
                if (MODE == 0 ? ch_is_nl(c) : c == '>') found = p;
            }
        }
        if (found != CH_NONE) atomicMin(s_pos, (ull)found);
        BLOCK_SYNC();
        const u64 r = *s_pos;
        BLOCK_SYNC();
        if (r != CH_NONE) return r;
    }
    return CH_NONE;
}
// last terminator in [lo, hi), CH_NONE if there is none; all threads call
__device__ __forceinline__ u64 ch_find_last_nl(const u8* __restrict__ text, u64 lo, u64 hi, ull* s_pos) {
    u64 top = hi;
    while (top > lo) {
        const u64 base = top - lo >= 256 * 8 ? top - 256 * 8 : lo;
        if (threadIdx.x == 0) *s_pos = 0;
        BLOCK_SYNC();
        const u64 a = base + (u64)threadIdx.x * 8;
        u64 found = 0;                                   // position + 1, 0 = none
        for (int j = 0; j < 8; ++j) {
            const u64 p = a + j;
            if (p < top && ch_is_nl(text[p])) found = p + 1;
        }
        if (found) atomicMax(s_pos, (ull)found);
        BLOCK_SYNC();
        const u64 r = *s_pos;
        BLOCK_SYNC();
        if (r) return r - 1;
        top = base;
    }
    return CH_NONE;
}

__global__ void __launch_bounds__(256)
chunk_chain_kernel(const u8* __restrict__ text, u64 n, u64 chunk_bytes, const u32* __restrict__ has_cr, u64* __restrict__ bounds,
                   u64 max_bounds, ull* __restrict__ nbounds) {
    __shared__ ull s_pos;
    if (*has_cr) { if (threadIdx.x == 0) *nbounds = CH_NONE; return; }      // "\r" present: use the exact kernels
    u64 nb = 1, b = 0;
    if (threadIdx.x == 0) bounds[0] = 0;
    while (true) {
        const u64 p = b + chunk_bytes;
        if (p >= n) break;
        u64 ls = p;
        if (!ch_is_nl(text[p - 1])) {                                       // p is inside a line: go to the next line
            const u64 q = ch_find_first<0>(text, p, n, &s_pos);
            if (q == CH_NONE) break;
            ls = q + 1;
        }
        if (ls >= n) break;
        const u64 g = ch_find_first<1>(text, ls, n, &s_pos);                 // first '>' in an eligible line
        if (g == CH_NONE) break;
        const u64 r = ch_find_last_nl(text, ls, g, &s_pos);                  // start of the line that holds it
        const u64 bs = r == CH_NONE ? ls : r + 1;
        if (threadIdx.x == 0 && nb < max_bounds) bounds[nb] = bs;
        nb++;
        b = bs;
    }
    if (threadIdx.x == 0) *nbounds = nb;
}
