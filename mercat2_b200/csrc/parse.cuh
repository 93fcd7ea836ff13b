// K1 -- FASTA text -> compacted symbol stream, on device.
//
// Semantics restated from the reference parser, lib/mercat2_kmers.py:47-63:
//   * lines end at '\n', '\r\n' or a lone '\r' (text-mode universal newlines); because an empty
//     line contributes nothing, treating '\r' and '\n' each as a terminator is equivalent;
//   * each line is strip()-ed of ASCII whitespace {09-0D, 1C-1F, 20} at both ends (interior
//     whitespace is a symbol);
//   * a stripped line starting with '>' is a header: it closes the record (we emit one MC2_SEP);
//   * any other line contributes its bytes minus every '*' to the record's sequence text.
// The result is `sym`: all sequence symbols in file order, records separated by MC2_SEP bytes, so
// that "a k-mer" is any k consecutive non-SEP bytes of `sym`.
//
// Parallel formulation: per byte class, a 3-state line automaton (S0 = only whitespace seen on this
// line, H = header, Q = sequence) whose per-segment transition map is summarised by two fields
// (segment contains a terminator?, end state when entered in S0) -- an associative operator scanned
// at thread (16 B), block (4 KiB tile) and grid level; trailing-whitespace detection is the mirror
// image: a reverse scan of "class of the first non-whitespace byte to my right".
#pragma once
#include "common.cuh"

#define PARSE_THREADS 256
#define PARSE_WARPS (PARSE_THREADS / 32)
#define PARSE_TILE (PARSE_THREADS * 16)

enum : u32 { CL_X = 0, CL_W = 1, CL_NL = 2, CL_GT = 3, CL_STAR = 4, CL_BAD = 5 };
enum : u32 { ST_S0 = 0, ST_H = 1, ST_Q = 2 };
enum : u32 { R_NONE = 0, R_NL = 1, R_OTHER = 2 };

// parse statistics (device struct, zeroed before the count pass)
struct ParseStats {
    ull n_acgt;     // emitted symbols in {A,C,G,T}
    ull n_upper;    // emitted symbols in 'A'..'Z'
    ull n_ascii;    // all emitted non-separator symbols
    ull n_sep;      // separators emitted
    ull n_bad;      // bytes >= 0x80 that would become symbols (anywhere outside header lines: unsupported -> error)
    ull n_sym;      // total bytes of `sym` (set by the offsets scan)
};

__device__ __forceinline__ u32 byte_class(u32 c) {
    if (c >= 64u) return c >= 128u ? CL_BAD : CL_X;       // letters: the hot case
    const u64 bit = 1ull << c;
    const u64 NLM = (1ull << 10) | (1ull << 13);
    const u64 WM = (1ull << 9) | (1ull << 11) | (1ull << 12) | (0xFull << 28) | (1ull << 32);
    if (bit & NLM) return CL_NL;
    if (bit & WM) return CL_W;
    if (c == 62u) return CL_GT;
    if (c == 42u) return CL_STAR;
    return CL_X;
}

// forward summary word: bits 0-1 end state when entered in S0, bit 2 = segment has a terminator
struct FwdOp {
    __device__ static __forceinline__ u32 identity() { return ST_S0; }   // hn = 0, e = S0
    __device__ static __forceinline__ u32 combine(u32 a, u32 b) {
        const u32 ae = a & 3u, be = b & 3u, bhn = b & 4u;
        const u32 e = bhn ? be : (ae == ST_S0 ? be : ae);
        return e | ((a | b) & 4u);
    }
    __device__ static __forceinline__ u32 apply(u32 summary, u32 state_in) {
        const u32 e = summary & 3u;
        return (summary & 4u) ? e : (state_in == ST_S0 ? e : state_in);
    }
};

struct ParseTileView {
    const u8* aligned;   // 16-byte aligned base (<= text)
    u64 lo, hi;          // valid virtual positions are [lo, hi)
};

__device__ __forceinline__ ParseTileView make_view(const u8* text, u64 len) {
    ParseTileView v;
    const u64 mis = (u64)(uintptr_t)text & 15ull;
    v.aligned = text - mis;
    v.lo = mis;
    v.hi = mis + len;
    return v;
}

// Load this thread's 16 bytes (virtual position p0, multiple of 16); bytes outside [lo,hi) come back
// as '\n' (a terminator: no symbol, resets the line state, counts as end-of-file for strip()).
__device__ __forceinline__ void load16(const ParseTileView& v, u64 p0, u32 w[4]) {
    if (p0 >= v.lo && p0 + 16 <= v.hi) {
        uint4 q = ld_nc_16(v.aligned + p0);
        w[0] = q.x; w[1] = q.y; w[2] = q.z; w[3] = q.w;
    } else {
        w[0] = w[1] = w[2] = w[3] = 0x0A0A0A0Au;
        if (p0 + 16 > v.lo && p0 < v.hi) {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const u64 p = p0 + i;
                if (p >= v.lo && p < v.hi) {
                    const u32 c = v.aligned[p];
                    w[i >> 2] = (w[i >> 2] & ~(0xFFu << (8 * (i & 3)))) | (c << (8 * (i & 3)));
                }
            }
        }
    }
}

__device__ __forceinline__ u32 byte_of(const u32 w[4], int i) { return (w[i >> 2] >> (8 * (i & 3))) & 0xFFu; }

// classes of 16 bytes packed 4 bits each; also the thread's forward and reverse summaries
// (badmask: bit i = byte i is >= 0x80; such a byte acts like any other non-blank byte in the line automaton.  Inside a
// header line it is dropped with the rest of the header -- the reference decodes UTF-8 there and never looks at it
// (lib/mercat2_kmers.py:52) -- anywhere else it would become a symbol and the text is rejected.)
__device__ __forceinline__ u64 classify16(const u32 w[4], u32& fwd, u32& rev, u32& badmask) {
    u64 cls = 0;
    u32 e = ST_S0, hn = 0, r = R_NONE, bad = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        u32 c = byte_class(byte_of(w, i));
        if (c == CL_BAD) { bad |= 1u << i; c = CL_X; }
        cls |= (u64)c << (4 * i);
        if (c == CL_NL) { e = ST_S0; hn = 4u; }
        else if (e == ST_S0 && c != CL_W) e = (c == CL_GT) ? ST_H : ST_Q;
        if (r == R_NONE && c != CL_W) r = (c == CL_NL) ? R_NL : R_OTHER;
    }
    fwd = e | hn;
    rev = r;
    badmask = bad;
    return cls;
}

// class of the first non-whitespace byte to the right of this thread, within the block; R_NONE if
// there is none (caller substitutes the tile's follow class).  smem: PARSE_WARPS words.
__device__ __forceinline__ u32 block_follow_class(u32 rev, u32* smem) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __syncwarp();
    const u32 nn = __ballot_sync(0xffffffffu, rev != R_NONE);
    const u32 first_lane = nn ? (u32)(__ffs(nn) - 1) : 0u;
    const u32 warp_first = __shfl_sync(0xffffffffu, rev, first_lane);
    BLOCK_SYNC();
    if (lane == 0) smem[warp] = nn ? warp_first : R_NONE;
    BLOCK_SYNC();
    const u32 above = (lane == 31) ? 0u : (nn & (0xFFFFFFFEu << lane));
    const u32 src = above ? (u32)(__ffs(above) - 1) : 0u;
    u32 res = __shfl_sync(0xffffffffu, rev, src);
    if (!above) {
        res = R_NONE;
        for (int w2 = warp + 1; w2 < PARSE_WARPS; ++w2) {
            const u32 f = smem[w2];
            if (f != R_NONE) { res = f; break; }
        }
    }
    return res;
}

// ---- K1a: per-tile summaries ------------------------------------------------------------------
__global__ void __launch_bounds__(PARSE_THREADS)
parse_summarize_kernel(const u8* __restrict__ text, u64 len, u8* __restrict__ tile_fwd, u8* __restrict__ tile_rev,
                       ParseStats* stats) {
    __shared__ u32 sm[PARSE_WARPS + 1];
    __shared__ u32 sm2[PARSE_WARPS];
    const ParseTileView v = make_view(text, len);
    const u64 p0 = (u64)blockIdx.x * PARSE_TILE + (u64)threadIdx.x * 16;
    u32 w[4];
    load16(v, p0, w);
    u32 fwd, rev, nbad;
    classify16(w, fwd, rev, nbad);
    u32 total;
    block_exclusive_scan<FwdOp, PARSE_WARPS>(fwd, sm, &total);
    // first non-NONE over the block, in order
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __syncwarp();
    const u32 nn = __ballot_sync(0xffffffffu, rev != R_NONE);
    const u32 first_lane = nn ? (u32)(__ffs(nn) - 1) : 0u;
    const u32 wf = __shfl_sync(0xffffffffu, rev, first_lane);
    if (lane == 0) sm2[warp] = nn ? wf : R_NONE;
    BLOCK_SYNC();
    if (threadIdx.x == 0) {
        u32 r = R_NONE;
        for (int i = 0; i < PARSE_WARPS; ++i) if (sm2[i] != R_NONE) { r = sm2[i]; break; }
        tile_fwd[blockIdx.x] = (u8)total;
        tile_rev[blockIdx.x] = (u8)r;
    }
}

// ---- K1b: grid-level scans of the tile summaries (single CTA) ------------------------------------
// tile_fwd[t] <- line state entering tile t; tile_rev[t] <- follow class after tile t (R_NL at EOF)
#define SCAN1_THREADS 1024
__global__ void __launch_bounds__(SCAN1_THREADS)
parse_scan_tiles_kernel(u8* tile_fwd, u8* tile_rev, u32 ntiles) {
    __shared__ u32 sm[SCAN1_THREADS / 32 + 1];
    __shared__ u32 part[SCAN1_THREADS];
    const u32 per = (ntiles + SCAN1_THREADS - 1) / SCAN1_THREADS;
    const u32 t0 = min(ntiles, threadIdx.x * per), t1 = min(ntiles, t0 + per);
    // forward
    u32 acc = FwdOp::identity();
    for (u32 t = t0; t < t1; ++t) acc = FwdOp::combine(acc, tile_fwd[t]);
    u32 pre = block_exclusive_scan<FwdOp, SCAN1_THREADS / 32>(acc, sm, nullptr);
    u32 state = FwdOp::apply(pre, ST_S0);
    for (u32 t = t0; t < t1; ++t) {
        const u32 s = tile_fwd[t];
        tile_fwd[t] = (u8)state;
        state = FwdOp::apply(s, state);
    }
    // reverse: first non-NONE class in tiles > t
    u32 first = R_NONE;
    for (u32 t = t0; t < t1; ++t) if (tile_rev[t] != R_NONE) { first = tile_rev[t]; break; }
    part[threadIdx.x] = first;
    BLOCK_SYNC();
    u32 follow = R_NL;                       // end of text behaves like a terminator for strip()
    for (u32 j = threadIdx.x + 1; j < SCAN1_THREADS; ++j) if (part[j] != R_NONE) { follow = part[j]; break; }
    for (u32 t = t1; t > t0; --t) {
        const u32 r = tile_rev[t - 1];
        tile_rev[t - 1] = (u8)follow;
        if (r != R_NONE) follow = r;
    }
}

// ---- K1c / K1e: count or write the symbols of each tile ------------------------------------------
template <bool WRITE>
__global__ void __launch_bounds__(PARSE_THREADS)
parse_emit_kernel(const u8* __restrict__ text, u64 len, const u8* __restrict__ tile_in, const u8* __restrict__ tile_follow,
                  u32* __restrict__ tile_cnt, const u64* __restrict__ tile_off, u8* __restrict__ sym,
                  ParseStats* stats, int toupper) {
    __shared__ u32 sm[PARSE_WARPS + 1];
    __shared__ u32 sm2[PARSE_WARPS];
    __shared__ __align__(16) u8 stage[WRITE ? PARSE_TILE : 16];
    const ParseTileView v = make_view(text, len);
    const u64 p0 = (u64)blockIdx.x * PARSE_TILE + (u64)threadIdx.x * 16;
    u32 w[4];
    load16(v, p0, w);
    u32 fwd, rev, nbad;
    const u64 cls = classify16(w, fwd, rev, nbad);
    const u32 pre = block_exclusive_scan<FwdOp, PARSE_WARPS>(fwd, sm, nullptr);
    u32 state = FwdOp::apply(pre, tile_in[blockIdx.x]);
    u32 follow = block_follow_class(rev, sm2);
    if (follow == R_NONE) follow = tile_follow[blockIdx.x];

    // backward pass: bit i of trail = the first non-whitespace byte after byte i is a terminator
    u32 trail = 0, nxt = follow;
#pragma unroll
    for (int i = 15; i >= 0; --i) {
        const u32 c = (u32)(cls >> (4 * i)) & 15u;
        if (nxt == R_NL) trail |= 1u << i;
        if (c != CL_W) nxt = (c == CL_NL) ? R_NL : R_OTHER;
    }
    // forward pass: emit
    u64 out_lo = 0, out_hi = 0;
    u32 cnt = 0, n_acgt = 0, n_upper = 0, n_sep = 0, n_hi = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const u32 c = (u32)(cls >> (4 * i)) & 15u;
        u32 b = byte_of(w, i);
        bool emit = false;
        if (c == CL_NL) state = ST_S0;
        else if (state == ST_S0) {
            if (c == CL_GT) { state = ST_H; emit = true; b = MC2_SEP; }
            else if (c != CL_W) { state = ST_Q; emit = (c == CL_X); }
        } else if (state == ST_Q) {
            emit = (c == CL_X) || (c == CL_GT) || (c == CL_W && !((trail >> i) & 1u));
        }
        if (emit) {
            if ((nbad >> i) & 1u) n_hi++;                      // a byte >= 0x80 outside a header line
            if (b != MC2_SEP) {
                if (toupper && b >= 'a' && b <= 'z') b -= 32;
                n_upper += (b >= 'A' && b <= 'Z');
                n_acgt += (b == 'A' || b == 'C' || b == 'G' || b == 'T');
            } else n_sep++;
            if (cnt < 8) out_lo |= (u64)b << (8 * cnt); else out_hi |= (u64)b << (8 * (cnt - 8));
            cnt++;
        }
    }
    u32 total;
    const u32 off = block_exclusive_scan<OpAdd, PARSE_WARPS>(cnt, sm, &total);
    if (!WRITE) {
        if (threadIdx.x == 0) tile_cnt[blockIdx.x] = total;
        // block-reduce the statistics (reuse the scan for simplicity: totals only)
        u32 t_acgt, t_upper, t_sep;
        block_exclusive_scan<OpAdd, PARSE_WARPS>(n_acgt, sm, &t_acgt);
        block_exclusive_scan<OpAdd, PARSE_WARPS>(n_upper, sm, &t_upper);
        block_exclusive_scan<OpAdd, PARSE_WARPS>(n_sep, sm, &t_sep);
        if (n_hi) atomicAdd(&stats->n_bad, (ull)n_hi);
        if (threadIdx.x == 0 && total) {
            atomicAdd(&stats->n_acgt, (ull)t_acgt);
            atomicAdd(&stats->n_upper, (ull)t_upper);
            atomicAdd(&stats->n_sep, (ull)t_sep);
            atomicAdd(&stats->n_ascii, (ull)(total - t_sep));
        }
    } else {
#pragma unroll
        for (int i = 0; i < 16; ++i)
            if ((u32)i < cnt) stage[off + i] = (u8)((i < 8 ? out_lo >> (8 * i) : out_hi >> (8 * (i - 8))) & 0xFFu);
        BLOCK_SYNC();
        u8* dst = sym + tile_off[blockIdx.x];
        for (u32 i = threadIdx.x; i < total; i += PARSE_THREADS) dst[i] = stage[i];
    }
}

// ---- generic single-CTA exclusive scan of per-tile u32 counts into u64 offsets ------------------
__global__ void __launch_bounds__(SCAN1_THREADS)
scan_counts_kernel(const u32* __restrict__ cnt, u64* __restrict__ off, u32 n, ull* total_out) {
    __shared__ u64 sm[SCAN1_THREADS / 32 + 1];
    const u32 per = (n + SCAN1_THREADS - 1) / SCAN1_THREADS;
    const u32 t0 = min(n, threadIdx.x * per), t1 = min(n, t0 + per);
    u64 acc = 0;
    for (u32 t = t0; t < t1; ++t) acc += cnt[t];
    u64 total;
    u64 base = block_exclusive_sum64<SCAN1_THREADS / 32>(acc, sm, &total);
    for (u32 t = t0; t < t1; ++t) { off[t] = base; base += cnt[t]; }
    if (threadIdx.x == 0 && total_out) *total_out = total;
}

__global__ void __launch_bounds__(256) symbol_stats_kernel(const u8* __restrict__ sym, u64 n, ParseStats* stats) {
    ull acgt = 0, upper = 0, bad = 0;
    for (u64 i = (u64)blockIdx.x * 256 + threadIdx.x; i < n; i += (u64)gridDim.x * 256) {
        const u32 b = sym[i];
        upper += (b >= 'A' && b <= 'Z');
        acgt += (b == 'A' || b == 'C' || b == 'G' || b == 'T');
        bad += (b >= 128u);
    }
    __syncwarp();
    for (int d = 16; d; d >>= 1) {
        acgt += __shfl_xor_sync(0xffffffffu, acgt, d);
        upper += __shfl_xor_sync(0xffffffffu, upper, d);
        bad += __shfl_xor_sync(0xffffffffu, bad, d);
    }
    if ((threadIdx.x & 31) == 0) {
        if (acgt) atomicAdd(&stats->n_acgt, acgt);
        if (upper) atomicAdd(&stats->n_upper, upper);
        if (bad) atomicAdd(&stats->n_bad, bad);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) { stats->n_ascii = n; stats->n_sym = n; }
}
