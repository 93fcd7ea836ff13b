// K1f + packed extraction -- the nucleotide fast lane.
//
// Same semantics as parse.cuh (reference parser lib/mercat2_kmers.py:47-63) restricted to "simple" FASTA
// text: line ends '\n', '\r\n' or '\r', 7-bit ASCII, and outside header lines no other whitespace/control
// byte and no '*' (header lines may say anything: their bytes are dropped).  Any other
// byte -- including a byte >= 0x80 outside a header line -- raises the `complex` flag and the chunk is redone by the
// general parser, so results never depend on this lane.  What it buys: the byte-granular work is done once, with SWAR on 32-bit words
// (about 11 ALU ops per text byte instead of ~55), and everything downstream works on 2-bit packed
// symbols: a k-mer is two funnel shifts instead of a 16-step rolling loop.
//
// Output of K1f: `codes` (u32, 16 symbols per word, symbol i at bits 30 - 2(i%16): the FIRST symbol of a word is its
// most significant pair) and `bad` (u32, one bit per symbol, symbol i at bit i%32: record separator or non-ACGT
// symbol).  A window of k symbols is countable on the fast lane iff none of its bad bits is set.  With this layout
// two funnel shifts yield the window's 64-bit code with its first symbol most significant, i.e. the
// order-preserving key itself (numeric order == Python str order of the k-mer): the range partition of
// rangecount.cuh works on its top 32 bits and no key is ever converted.
#pragma once
#include "common.cuh"
#include "parse.cuh"
#include "rangecount.cuh"

#define FN_THREADS 256
#define FN_WARPS (FN_THREADS / 32)
#define FN_TILE (FN_THREADS * 16)
#define FN_LOOKBACK_ITERS 2048          // x 32 bytes: how far a tile looks left for its line start

struct FnStats {
    ull n_sym;      // symbols emitted (kept bytes + one separator per header line)
    ull packed;     // low 32 bits: kept bytes; high 32 bits: kept bytes outside {A,C,G,T}   (count pass, MODE 0)
    ull complex;    // != 0: text is not "simple" -> use the general parser
    ull packed2;    // same fields as `packed`, accumulated by the write pass
};

__device__ __forceinline__ u32 fn_nz(u32 x) { return (((x & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | x) & 0x80808080u; }
__device__ __forceinline__ u32 fn_eq(u32 w, u32 c4) { return ~fn_nz(w ^ c4) & 0x80808080u; }
__device__ __forceinline__ u32 fn_lt21(u32 w) { return ~(((w & 0x7F7F7F7Fu) + 0x5F5F5F5Fu) | w) & 0x80808080u; }
__device__ __forceinline__ u32 fn_movemask(u32 m) { return (((m >> 7) * 0x00204081u) >> 21) & 0xFu; }

struct FnMasks {
    u32 nl, gt, acgt;   // 16-bit masks over the thread's 16 bytes
    u32 codes;          // 2-bit code of every byte (garbage where the byte is not ACGT)
    u32 cx;             // 16-bit mask: whitespace/control bytes and '*' (not simple unless inside a header line)
    u32 hi;             // 16-bit mask: bytes >= 0x80 (simple only inside a header line, whose text is dropped)
};

template <bool CODES>
__device__ __forceinline__ FnMasks fn_classify(const u32 w[4]) {
    FnMasks m;
    m.nl = m.gt = m.acgt = m.codes = m.cx = m.hi = 0;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const u32 x = w[q];
        const u32 mnl = fn_eq(x, 0x0A0A0A0Au) | fn_eq(x, 0x0D0D0D0Du);
        const u32 mgt = fn_eq(x, 0x3E3E3E3Eu);
        m.cx |= fn_movemask((fn_lt21(x) & ~mnl) | fn_eq(x, 0x2A2A2A2Au)) << (4 * q);
        m.hi |= fn_movemask(x & 0x80808080u) << (4 * q);
        m.nl |= fn_movemask(mnl) << (4 * q);
        m.gt |= fn_movemask(mgt) << (4 * q);
        if (!CODES) continue;
        const u32 t = ((x >> 1) ^ (x >> 2)) & 0x03030303u;            // A,C,G,T -> 0,1,2,3 per byte
        u32 p = t | (t >> 6);
        p = (p | (p >> 12)) & 0xFFu;
        m.codes |= p << (8 * q);
        u32 sel = (t | (t >> 4)) & 0x00330033u;
        sel = (sel | (sel >> 8)) & 0x3333u;
        const u32 back = __byte_perm(0x54474341u, 0u, sel);            // "ACGT"[code] for the four bytes
        m.acgt |= (fn_movemask(fn_nz(back ^ x)) ^ 0xFu) << (4 * q);
    }
    return m;
}

// forward summary of 16 bytes under the simple rules (header <=> first byte of the line is '>')
__device__ __forceinline__ u32 fn_summary(const FnMasks& m) {
    if (m.nl == 0) return (m.gt & 1u) ? ST_H : ST_Q;
    if (m.nl >> 15) return ST_S0 | 4u;
    const int ls = 32 - __clz(m.nl);                  // first byte after the last terminator
    return (((m.gt >> ls) & 1u) ? ST_H : ST_Q) | 4u;
}

struct FnEmit {
    u32 emit;      // bytes that become symbols (kept bytes + header starts)
    u32 bad;       // symbols that stop a fast window (header start, or kept byte outside ACGT)
    u32 hs;        // header starts
    u32 keep;
    u32 cx;        // != 0: a byte the simple rules do not cover
};

__device__ __forceinline__ FnEmit fn_emit_masks(const FnMasks& m, u32 state) {
    const u32 open = ~m.nl;                                            // upper 16 bits are ones: carries run off the top
    const u32 ls = ((m.nl << 1) | (state == ST_S0 ? 1u : 0u)) & 0xFFFFu;
    const u32 hs = m.gt & ls;
    const u32 seed = hs | ((state == ST_H && !(m.nl & 1u)) ? 1u : 0u);
    const u32 hdr = (open & ~(open + seed)) & 0xFFFFu;                 // from each seed up to the next terminator
    FnEmit e;
    e.keep = ~hdr & ~m.nl & 0xFFFFu;
    e.hs = hs;
    e.emit = e.keep | hs;
    e.bad = hs | (e.keep & ~m.acgt);
    e.cx = (m.cx | m.hi) & ~hdr;
    return e;
}

// line state at the first byte of a tile: walk left to the previous terminator (warp 0)
__device__ __forceinline__ u32 fn_tile_state(const u8* text, u64 idx /*text index of the tile's first byte*/, bool& complex) {
    const int lane = threadIdx.x & 31;
    if (idx == 0) return ST_S0;
    u64 line_start = 0;
    bool found = false;
    for (int it = 0; it < FN_LOOKBACK_ITERS; ++it) {
        const u64 back = (u64)it * 32 + lane + 1;
        const bool in = back <= idx;
        const u32 c = in ? text[idx - back] : 10u;                    // the start of the text acts as a terminator
        __syncwarp();
        const u32 m = __ballot_sync(0xffffffffu, c == 10u || c == 13u);
        if (m) {
            const u64 b = (u64)it * 32 + (__ffs(m) - 1) + 1;
            line_start = b > idx ? 0 : idx - b + 1;
            found = true;
            break;
        }
    }
    if (!found) { complex = true; return ST_Q; }
    if (line_start == idx) return ST_S0;
    return text[line_start] == '>' ? ST_H : ST_Q;
}

// compress the bits of `field_src` selected by the runs of `mask` (BITS per position)
template <int BITS>
__device__ __forceinline__ u32 fn_compress(u32 src, u32 mask) {
    if (mask == 0xFFFFu) return src;
    u32 out = 0, o = 0, m = mask;
    while (m) {
        const int a = __ffs(m) - 1;
        const u32 run = m >> a;
        const int len = __ffs(~run) - 1;
        const u32 lm = len * BITS >= 32 ? 0xFFFFFFFFu : ((1u << (len * BITS)) - 1u);
        out |= ((src >> (a * BITS)) & lm) << (o * BITS);
        o += len;
        m &= ~(((1u << len) - 1u) << a);
    }
    return out;
}

// 16 two-bit symbols assembled low-pair-first -> the stored word (first symbol in the top pair)
__device__ __forceinline__ u32 fn_store_order(u32 x) {
    x = __brev(x);                                                 // pairs reversed AND swapped inside
    return ((x & 0x55555555u) << 1) | ((x >> 1) & 0x55555555u);
}

// ---- K1f: count pass and write pass ---------------------------------------------------------------------------
// MODE 0: count pass with alphabet statistics (first piece of a sample); MODE 1: write pass (also counts the kept
// non-ACGT bytes); MODE 2: light count pass (line structure only).
// Every thread owns FN_V consecutive 16-byte groups (64 bytes), so one pair of block scans (line state, symbol offset)
// serves a 16 KiB tile: the barriers, not the SWAR arithmetic, bounded the 4 KiB-tile version.
#define FN_V 4
#define FN_VTILE (FN_THREADS * 16 * FN_V)
#define FN_SC_WORDS (FN_VTILE / 16 + 4)
#define FN_SB_WORDS (FN_VTILE / 32 + 4)

template <int MODE>
__global__ void __launch_bounds__(FN_THREADS)
fn_parse_kernel(const u8* __restrict__ text, u64 len, u8* __restrict__ tile_state, u32* __restrict__ tile_cnt,
                const u64* __restrict__ tile_off, u32* __restrict__ codes, u32* __restrict__ bad, FnStats* stats) {
    constexpr bool WRITE = MODE == 1;
    __shared__ u32 sm[FN_WARPS + 1];
    __shared__ u32 s_state;
    __shared__ u32 s_c[WRITE ? FN_SC_WORDS : 1], s_b[WRITE ? FN_SB_WORDS : 1];
    const ParseTileView v = make_view(text, len);
    const u64 tile_v = (u64)blockIdx.x * FN_VTILE;
    const u64 p0 = tile_v + (u64)threadIdx.x * (16 * FN_V);
    FnMasks m[FN_V];
    u32 summary = FwdOp::identity();
#pragma unroll
    for (int g = 0; g < FN_V; ++g) {
        u32 w[4];
        load16(v, p0 + 16 * g, w);
        m[g] = fn_classify<MODE != 2>(w);
        summary = FwdOp::combine(summary, fn_summary(m[g]));
    }
    if (!WRITE) {
        if (threadIdx.x < 32) {
            bool cx = false;
            const u32 st = tile_v <= v.lo ? (u32)ST_S0 : fn_tile_state(text, tile_v - v.lo, cx);
            if (threadIdx.x == 0) {
                s_state = st;
                tile_state[blockIdx.x] = (u8)st;
                if (cx) atomicAdd(&stats->complex, 1ull);
            }
        }
    } else {
        if (threadIdx.x == 0) s_state = tile_state[blockIdx.x];
        for (u32 i = threadIdx.x; i < FN_SC_WORDS; i += FN_THREADS) s_c[i] = 0;
        for (u32 i = threadIdx.x; i < FN_SB_WORDS; i += FN_THREADS) s_b[i] = 0;
    }
    const u32 pre = block_exclusive_scan<FwdOp, FN_WARPS>(summary, sm, nullptr);   // barriers publish s_state
    u32 state = FwdOp::apply(pre, s_state);
    u32 emit[FN_V], bads[FN_V];
    u32 cnt = 0, kept = 0, slow = 0, cx = 0;
#pragma unroll
    for (int g = 0; g < FN_V; ++g) {
        const FnEmit e = fn_emit_masks(m[g], state);
        state = FwdOp::apply(fn_summary(m[g]), state);
        emit[g] = e.emit;
        bads[g] = e.bad;
        cnt += __popc(e.emit);
        kept += __popc(e.keep);
        slow += __popc(e.keep & ~m[g].acgt);
        cx |= e.cx;
    }
    u32 total;
    const u32 off = block_exclusive_scan<OpAdd, FN_WARPS>(cnt, sm, &total);
    if (!WRITE) {
        if (threadIdx.x == 0) tile_cnt[blockIdx.x] = total;
        // statistics: one 64-bit atomic per tile (kept | non-ACGT << 32); `complex` only when it happens
        if (MODE == 0) {
            u32 t_kept, t_slow;
            block_exclusive_scan<OpAdd, FN_WARPS>(kept, sm, &t_kept);
            block_exclusive_scan<OpAdd, FN_WARPS>(slow, sm, &t_slow);
            if (threadIdx.x == 0 && (t_kept | t_slow)) atomicAdd(&stats->packed, (ull)t_kept | ((ull)t_slow << 32));
        }
        if (cx) atomicAdd(&stats->complex, 1ull);
    } else {
        {   // kept bytes outside ACGT (their windows go to the wide path): rare, so only warps that see one report
            __syncwarp();
            if (__ballot_sync(0xffffffffu, slow != 0)) {
                u32 t = slow;
#pragma unroll
                for (int d = 16; d; d >>= 1) t += __shfl_xor_sync(0xffffffffu, t, d);
                if ((threadIdx.x & 31) == 0) atomicAdd(&stats->packed2, (ull)t << 32);
            }
        }
        if (total == 0) return;
        const u64 s_tile = tile_off[blockIdx.x];            // global symbol index of the tile's first symbol
        const u64 base_sym = s_tile & ~31ull;
        const u32 rel0 = (u32)(s_tile - base_sym);
        u32 rel = rel0 + off;
#pragma unroll
        for (int g = 0; g < FN_V; ++g) {
            const u32 c = __popc(emit[g]);
            if (c) {
                const u32 cf = fn_compress<2>(m[g].codes, emit[g]);
                const u32 bf = fn_compress<1>(bads[g], emit[g]);
                const u64 cv = (u64)(c == 16 ? cf : (cf & ((1u << (2 * c)) - 1u))) << (2 * (rel & 15));
                atomicOr(&s_c[rel >> 4], (u32)cv);
                if (cv >> 32) atomicOr(&s_c[(rel >> 4) + 1], (u32)(cv >> 32));
                const u64 bv = (u64)(bf & ((1u << c) - 1u)) << (rel & 31);
                atomicOr(&s_b[rel >> 5], (u32)bv);
                if (bv >> 32) atomicOr(&s_b[(rel >> 5) + 1], (u32)(bv >> 32));
                rel += c;
            }
        }
        BLOCK_SYNC();
        const u32 fw = rel0 >> 4, nw = (rel0 + total + 15) >> 4;
        u32* gc = codes + (base_sym >> 4);
        for (u32 i = fw + threadIdx.x; i < nw; i += FN_THREADS) {
            const u32 val = fn_store_order(s_c[i]);
            if (i == fw || i == nw - 1) { if (val) atomicOr(&gc[i], val); }
            else gc[i] = val;
        }
        const u32 nbw = (rel0 + total + 31) >> 5;
        u32* gb = bad + (base_sym >> 5);
        for (u32 i = threadIdx.x; i < nbw; i += FN_THREADS) {
            const u32 val = s_b[i];
            if (i == 0 || i == nbw - 1) { if (val) atomicOr(&gb[i], val); }
            else gb[i] = val;
        }
    }
}

// ---- packed extraction ---------------------------------------------------------------------------------------
struct PackedView {
    const u32* codes;
    const u32* bad;
    u64 n;             // symbols
};

// the three code words that hold the 16 windows starting in word g, and the mask of countable windows
struct FnWords { u32 w0, w1, w2; };
// top 32 bits of the left-aligned code of window j (its first 16 symbols); j must be a compile-time constant
__device__ __forceinline__ u32 fn_hi(const FnWords& w, int j) { return __funnelshift_l(w.w1, w.w0, 2 * j); }
// the window's key: k symbols, first symbol most significant, right-aligned in 2k bits
__device__ __forceinline__ u64 fn_key(const FnWords& w, int j, int rshift /*64 - 2k*/) {
    const u32 hi = __funnelshift_l(w.w1, w.w0, 2 * j);
    const u32 lo = __funnelshift_l(w.w2, w.w1, 2 * j);
    return (((u64)hi << 32) | lo) >> rshift;
}
__device__ __forceinline__ u32 fn_load_windows(const PackedView& pv, u64 g, int k, FnWords& w) {
    const u64 nwords = (pv.n + 15) >> 4;
    w.w0 = w.w1 = w.w2 = 0;
    if (g >= nwords) return 0;
    w.w0 = pv.codes[g];
    w.w1 = g + 1 < nwords ? pv.codes[g + 1] : 0u;
    w.w2 = g + 2 < nwords ? pv.codes[g + 2] : 0u;
    const u64 nbw = (pv.n + 31) >> 5;
    const u64 bi = g >> 1;
    u64 b = (u64)pv.bad[bi] | ((bi + 1 < nbw ? (u64)pv.bad[bi + 1] : 0xFFFFFFFFull) << 32);
    b >>= (g & 1) * 16;
    b |= 0xFFFF000000000000ull;                                  // only 48 bits are real
    const u64 first = g << 4;
    if (first + 48 > pv.n) b |= ~0ull << (pv.n - first);          // symbols past the end stop every window
    // OR of the k bad bits of every window: doubling smear
    u64 x = b;
    int cur = 1;
    while (cur * 2 <= k) { x |= x >> cur; cur *= 2; }
    if (cur < k) x |= x >> (k - cur);
    return ~(u32)x & 0xFFFFu;
}
// prefix mask: for k < 16 only the first k symbols of the top word belong to the key
__device__ __forceinline__ u32 fn_pmask(int k) { return k >= 16 ? 0xFFFFFFFFu : ~(0xFFFFFFFFu >> (2 * k)); }

#define FN_HIST_THREADS 1024
// prefix histogram of a strided sample of the packed stream (every `stride`-th code word) for rp_plan_kernel
__global__ void __launch_bounds__(FN_HIST_THREADS)
rp_sample_packed_kernel(PackedView pv, int k, u64 stride, u32 base, u32 sh, u32* __restrict__ shist) {
    __shared__ u32 hist[RP_LUT];
    for (u32 i = threadIdx.x; i < RP_LUT; i += FN_HIST_THREADS) hist[i] = 0;
    BLOCK_SYNC();
    const u32 hist32 = (u32)__cvta_generic_to_shared(hist);
    const u32 pm = fn_pmask(k);
    const u64 nwords = (pv.n + 15) >> 4;
    for (u64 g = ((u64)blockIdx.x * FN_HIST_THREADS + threadIdx.x) * stride; g < nwords; g += (u64)gridDim.x * FN_HIST_THREADS * stride) {
        FnWords w;
        const u32 valid = fn_load_windows(pv, g, k, w);
#pragma unroll
        for (int j = 0; j < 16; ++j)
            smem_red_inc_if(hist32 + 4u * rp_lut_index(fn_hi(w, j) & pm, base, sh), (valid >> j) & 1u);
    }
    BLOCK_SYNC();
    for (u32 b = threadIdx.x; b < RP_LUT; b += FN_HIST_THREADS) {
        const u32 c = hist[b];
        if (c) atomicAdd(&shist[b], c);
    }
}

// exact histogram over all sub-buckets (FINE) or over the level-1 buckets only (level-0 partition of a very large chunk)
template <bool FINE>
__global__ void __launch_bounds__(FN_HIST_THREADS)
fn_hist_kernel(PackedView pv, int k, RpView r, u32 nb, u32* __restrict__ ghist) {
    extern __shared__ __align__(16) u8 dyn[];
    RpShared& rs = *reinterpret_cast<RpShared*>(dyn);
    u32* hist = reinterpret_cast<u32*>(dyn + sizeof(RpShared));
    rp_load_shared(r, rs, FINE);
    for (u32 i = threadIdx.x; i < nb; i += FN_HIST_THREADS) hist[i] = 0;
    BLOCK_SYNC();
    const u32 hist32 = (u32)__cvta_generic_to_shared(hist);
    const u32 pm = fn_pmask(k);
    const u64 nwords = (pv.n + 15) >> 4;
    const bool lin = *r.linear != 0;
    for (u64 g = (u64)blockIdx.x * FN_HIST_THREADS + threadIdx.x; g < nwords; g += (u64)gridDim.x * FN_HIST_THREADS) {
        FnWords w;
        const u32 valid = fn_load_windows(pv, g, k, w);
#pragma unroll
        for (int j = 0; j < 16; ++j) {                     // predicated reduction: no branch per window
            const u32 p = fn_hi(w, j) & pm;
            const u32 b1 = lin ? rp_b1_linear(r, p) : rp_b1(rs, r, p);
            smem_red_inc_if(hist32 + 4u * (FINE ? b1 * HC_NB2 + rp_b2(rs.l1[b1], p) : b1), (valid >> j) & 1u);
        }
    }
    BLOCK_SYNC();
    for (u32 b = threadIdx.x; b < nb; b += FN_HIST_THREADS) {
        const u32 n = hist[b];
        if (n) atomicAdd(&ghist[b], n);
    }
}

// ---- K1f in ONE pass: classification, line state, symbol offsets by chained look-back, packed write -----------------
// The two-pass version reads the text twice (count, then write at scanned offsets).  Here every tile publishes its
// symbol count in a descriptor word (flag | value) and obtains its exclusive offset by walking back over its
// predecessors' descriptors until it meets one that already carries an inclusive prefix ("decoupled look-back").
// Tiles take their index from an atomic ticket, so a tile only ever waits for tiles that are already running or
// done.  A bounded spin turns any protocol failure into the `complex` flag (the chunk is then redone by the general
// parser) instead of a hang.  STATS: also the kept / non-ACGT totals of the count pass (first piece of a sample).
#define FN_DESC_AGG (1ull << 62)
#define FN_DESC_PREFIX (2ull << 62)
#define FN_DESC_MASK ((1ull << 62) - 1)
#define FN_SPIN_LIMIT (1u << 24)

template <bool STATS>
__global__ void __launch_bounds__(FN_THREADS)
fn_parse_single_kernel(const u8* __restrict__ text, u64 len, u32 ntiles, ull* __restrict__ desc, u32* __restrict__ ticket,
                       u32* __restrict__ codes, u32* __restrict__ bad, FnStats* stats) {
    __shared__ u32 sm[FN_WARPS + 1];
    __shared__ u32 s_state, s_tile_id;
    __shared__ ull s_off, s_red[FN_WARPS];
    __shared__ u32 s_c[260], s_b[132];
    if (threadIdx.x == 0) s_tile_id = atomicAdd(ticket, 1u);
    for (u32 i = threadIdx.x; i < 260; i += FN_THREADS) s_c[i] = 0;
    for (u32 i = threadIdx.x; i < 132; i += FN_THREADS) s_b[i] = 0;
    BLOCK_SYNC();
    const u32 tile = s_tile_id;
    const ParseTileView v = make_view(text, len);
    const u64 tile_v = (u64)tile * FN_TILE;
    const u64 p0 = tile_v + (u64)threadIdx.x * 16;
    u32 w[4];
    load16(v, p0, w);
    const FnMasks m = fn_classify<true>(w);
    if (threadIdx.x < 32) {
        bool cx = false;
        const u32 st = tile_v <= v.lo ? (u32)ST_S0 : fn_tile_state(text, tile_v - v.lo, cx);
        if (threadIdx.x == 0) {
            s_state = st;
            if (cx) atomicAdd(&stats->complex, 1ull);
        }
    }
    const u32 pre = block_exclusive_scan<FwdOp, FN_WARPS>(fn_summary(m), sm, nullptr);   // barriers publish s_state
    const u32 state = FwdOp::apply(pre, s_state);
    const FnEmit e = fn_emit_masks(m, state);
    const u32 cnt = __popc(e.emit);
    u32 total;
    const u32 off = block_exclusive_scan<OpAdd, FN_WARPS>(cnt, sm, &total);
    if (e.cx) atomicAdd(&stats->complex, 1ull);
    // ---- chained offsets ----
    if (threadIdx.x == 0) {
        ull excl = 0;
        if (tile == 0) {
            atomicExch(&desc[0], FN_DESC_PREFIX | (ull)total);
        } else {
            atomicExch(&desc[tile], FN_DESC_AGG | (ull)total);
            u32 spins = 0;
            for (u32 p = tile; p-- > 0;) {
                ull d;
                while (((d = *(volatile ull*)&desc[p]) >> 62) == 0) {
                    if (++spins > FN_SPIN_LIMIT) break;
                    __nanosleep(20);
                }
                if ((d >> 62) == 0) { atomicAdd(&stats->complex, 1ull); break; }     // give up: the host falls back
                excl += d & FN_DESC_MASK;
                if (d & FN_DESC_PREFIX) break;
            }
            atomicExch(&desc[tile], FN_DESC_PREFIX | (excl + total));
        }
        s_off = excl;
        if (tile == ntiles - 1) stats->n_sym = excl + total;
    }
    // ---- statistics ----
    {
        const u32 kept = __popc(e.keep), slow = __popc(e.keep & ~m.acgt);
        if (STATS) {
            ull t = (ull)kept | ((ull)slow << 32);
            __syncwarp();
#pragma unroll
            for (int d = 16; d; d >>= 1) t += __shfl_xor_sync(0xffffffffu, t, d);
            if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = t;
        } else {
            __syncwarp();
            if (__ballot_sync(0xffffffffu, slow != 0)) {                  // rare: only warps that see a non-ACGT byte report
                u32 t = slow;
#pragma unroll
                for (int d = 16; d; d >>= 1) t += __shfl_xor_sync(0xffffffffu, t, d);
                if ((threadIdx.x & 31) == 0) atomicAdd(&stats->packed2, (ull)t << 32);
            }
        }
    }
    BLOCK_SYNC();                                                          // s_off, s_red
    if (STATS && threadIdx.x == 0) {
        ull t = 0;
#pragma unroll
        for (int i = 0; i < FN_WARPS; ++i) t += s_red[i];
        if (t) {
            atomicAdd(&stats->packed, t);
            if (t >> 32) atomicAdd(&stats->packed2, t & 0xFFFFFFFF00000000ull);
        }
    }
    if (total == 0) return;
    const u64 s_tile = s_off;                                              // global symbol index of the tile's first symbol
    const u64 base_sym = s_tile & ~31ull;
    const u32 rel0 = (u32)(s_tile - base_sym);
    if (cnt) {
        const u32 rel = rel0 + off;
        const u32 cf = fn_compress<2>(m.codes, e.emit);
        const u32 bf = fn_compress<1>(e.bad, e.emit);
        const u64 cv = (u64)(cnt == 16 ? cf : (cf & ((1u << (2 * cnt)) - 1u))) << (2 * (rel & 15));
        atomicOr(&s_c[rel >> 4], (u32)cv);
        if (cv >> 32) atomicOr(&s_c[(rel >> 4) + 1], (u32)(cv >> 32));
        const u64 bv = (u64)(bf & ((1u << cnt) - 1u)) << (rel & 31);
        atomicOr(&s_b[rel >> 5], (u32)bv);
        if (bv >> 32) atomicOr(&s_b[(rel >> 5) + 1], (u32)(bv >> 32));
    }
    BLOCK_SYNC();
    const u32 fw = rel0 >> 4, nw = (rel0 + total + 15) >> 4;
    u32* gc = codes + (base_sym >> 4);
    for (u32 i = fw + threadIdx.x; i < nw; i += FN_THREADS) {
        const u32 val = fn_store_order(s_c[i]);
        if (i == fw || i == nw - 1) { if (val) atomicOr(&gc[i], val); }
        else gc[i] = val;
    }
    const u32 nbw = (rel0 + total + 31) >> 5;
    u32* gb = bad + (base_sym >> 5);
    for (u32 i = threadIdx.x; i < nbw; i += FN_THREADS) {
        const u32 val = s_b[i];
        if (i == 0 || i == nbw - 1) { if (val) atomicOr(&gb[i], val); }
        else gb[i] = val;
    }
}

// ---- dense tables straight from the packed stream (4^k bins, k <= 15) -----------------------------------------
// Index = the window's code with its first symbol most significant (the layout of the sample table) = the top 2k bits
// of the window's first code word.  SMEM: per-CTA histogram replicated per warp group, flushed once; otherwise
// reduction atomics on the L2-resident global table.
template <bool SMEM>
__global__ void __launch_bounds__(FN_HIST_THREADS)
fn_dense_kernel(PackedView pv, int k, u32 bins, u32 nrep, u32* __restrict__ table) {
    extern __shared__ __align__(16) u8 dyn[];
    u32* hist = reinterpret_cast<u32*>(dyn);                      // [nrep][bins] (SMEM only)
    if (SMEM) {
        for (u32 i = threadIdx.x; i < bins * nrep; i += FN_HIST_THREADS) hist[i] = 0;
        BLOCK_SYNC();
    }
    u32* my = SMEM ? hist + (u32)((threadIdx.x >> 5) % nrep) * bins : table;
    const u32 my32 = SMEM ? (u32)__cvta_generic_to_shared(my) : 0u;
    const int rs = 32 - 2 * k;
    const u64 nwords = (pv.n + 15) >> 4;
    for (u64 g = (u64)blockIdx.x * FN_HIST_THREADS + threadIdx.x; g < nwords; g += (u64)gridDim.x * FN_HIST_THREADS) {
        FnWords w;
        const u32 valid = fn_load_windows(pv, g, k, w);
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const u32 idx = fn_hi(w, j) >> rs;
            if (SMEM) smem_red_inc_if(my32 + 4u * idx, (valid >> j) & 1u);           // predicated: no branch per window
            else if ((valid >> j) & 1u) atomicAdd(&my[idx], 1u);
        }
    }
    if (SMEM) {
        BLOCK_SYNC();
        for (u32 b = threadIdx.x; b < bins; b += FN_HIST_THREADS) {
            u32 sum = 0;
            for (u32 r = 0; r < nrep; ++r) sum += hist[r * bins + b];
            if (sum) atomicAdd(&table[b], sum);
        }
    }
}

// packed stream -> keys grouped by level-1 bucket (or by level-0 group with 64-bit group bases); persistent CTAs, the
// LUT is loaded once per CTA and only when the plan is not the closed form (see hk_scatter1_kernel)
__global__ void __launch_bounds__(EX_THREADS)
fn_scatter1_kernel(PackedView pv, int k, RpView r, u32* __restrict__ cur1, u64* __restrict__ keys1, const u64* __restrict__ base64) {
    extern __shared__ __align__(16) u8 dyn_sc[];                    // HC_SCATTER_SMEM_LUT bytes
    u64* stage = reinterpret_cast<u64*>(dyn_sc);
    u16* sdig = reinterpret_cast<u16*>(dyn_sc + HC_SDST_OFFSET);
    u16* s_lut = reinterpret_cast<u16*>(dyn_sc + HC_SCATTER_SMEM16);
    __shared__ u32 cnt[HC_MAX_NB1], loff[HC_MAX_NB1], gbase[HC_MAX_NB1];
    __shared__ u32 sm[EX_WARPS + 1];
    const bool lin = *r.linear != 0;
    if (!lin)
        for (u32 i = threadIdx.x; i < RP_LUT / 8; i += EX_THREADS) reinterpret_cast<uint4*>(s_lut)[i] = reinterpret_cast<const uint4*>(r.lut)[i];
    const int rshift = 64 - 2 * k;
    const u32 pm = fn_pmask(k);
    const u64 nwords = (pv.n + 15) >> 4;
    const u64 ntiles = (nwords + EX_THREADS - 1) / EX_THREADS;
    for (u64 tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        for (u32 i = threadIdx.x; i < r.nb1; i += EX_THREADS) cnt[i] = 0;
        FnWords w;
        const u32 valid = fn_load_windows(pv, tile * EX_THREADS + threadIdx.x, k, w);
        BLOCK_SYNC();
        // keys and digits are recomputed (funnel shifts) wherever they are needed instead of living in 32 registers
        auto key = [&](int i) { return fn_key(w, i, rshift); };
        auto dig = [&](int i) {
            const u32 p = fn_hi(w, i) & pm;
            return lin ? rp_b1_linear(r, p) : (u32)s_lut[rp_lut_index(p, r.base, r.sh)];
        };
        hc_group_and_write<false>(key, dig, valid, r.nb1, stage, sdig, cnt, loff, gbase, sm, cur1, keys1, base64);
        BLOCK_SYNC();                                          // the staging area and the counters are re-used by the next tile
    }
}

// debug: order-independent checksum of all countable windows of the packed stream
__global__ void fn_checksum_kernel(PackedView pv, int k, ull* __restrict__ out /*[0]=sum, [1]=xor, [2]=count*/) {
    const int rshift = 64 - 2 * k;
    const u64 nwords = (pv.n + 15) >> 4;
    ull sum = 0, x = 0, cnt = 0;
    for (u64 g = (u64)blockIdx.x * blockDim.x + threadIdx.x; g < nwords; g += (u64)gridDim.x * blockDim.x) {
        FnWords w;
        const u32 valid = fn_load_windows(pv, g, k, w);
#pragma unroll
        for (int j = 0; j < 16; ++j)
            if ((valid >> j) & 1u) { const u64 key = fn_key(w, j, rshift); sum += key * 0x9E3779B97F4A7C15ull; x ^= key; cnt++; }
    }
    atomicAdd(&out[0], sum); atomicXor(&out[1], x); atomicAdd(&out[2], cnt);
}
__global__ void key_checksum_kernel(const u64* __restrict__ keys, u64 n, ull* __restrict__ out) {
    ull sum = 0, x = 0, cnt = 0;
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x) {
        sum += keys[i] * 0x9E3779B97F4A7C15ull; x ^= keys[i]; cnt++;
    }
    atomicAdd(&out[0], sum); atomicXor(&out[1], x); atomicAdd(&out[2], cnt);
}
