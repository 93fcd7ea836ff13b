// Row N3: the sample x k-mer table of merge_tsv / merge_tsv_T (reference lib/mercat2_report.py:98-194)
// from tables that already sit on the device, plus a device parser for per-sample TSV files so that the reference's
// file-based entry point keeps working.
//
//   union of the samples' k-mers  = sort + unique of all rows as k literal bytes (the wide path's LSD sort)
//   matrix[u][s]                  = count of k-mer u in sample s (binary search of every sample row in the union)
//   text                          = one thread per cell: "<k-mer>" (first cell of a row) "\t<count>" ("\n" after the last)
#pragma once
#include "common.cuh"
#include "tsv.cuh"

// ---- TSV text -> rows ---------------------------------------------------------------------------------------------
#define MG_THREADS 256
#define MG_TILE (MG_THREADS * 16)

__global__ void __launch_bounds__(MG_THREADS)
mg_newline_count_kernel(const u8* __restrict__ text, u64 n, u32* __restrict__ tile_cnt) {
    const u64 p0 = (u64)blockIdx.x * MG_TILE + (u64)threadIdx.x * 16;
    u32 c = 0;
    for (int i = 0; i < 16; ++i) c += (p0 + i < n && text[p0 + i] == '\n') ? 1u : 0u;
    __shared__ u32 sm[MG_THREADS / 32 + 1];
    u32 t;
    block_exclusive_scan<OpAdd, MG_THREADS / 32>(c, sm, &t);
    if (threadIdx.x == 0) tile_cnt[blockIdx.x] = t;
}

__global__ void __launch_bounds__(MG_THREADS)
mg_newline_write_kernel(const u8* __restrict__ text, u64 n, const u64* __restrict__ tile_off, u64* __restrict__ pos) {
    const u64 p0 = (u64)blockIdx.x * MG_TILE + (u64)threadIdx.x * 16;
    u32 c = 0;
    for (int i = 0; i < 16; ++i) c += (p0 + i < n && text[p0 + i] == '\n') ? 1u : 0u;
    __shared__ u32 sm[MG_THREADS / 32 + 1];
    u32 off = block_exclusive_scan<OpAdd, MG_THREADS / 32>(c, sm, nullptr);
    u64 at = tile_off[blockIdx.x] + off;
    for (int i = 0; i < 16; ++i)
        if (p0 + i < n && text[p0 + i] == '\n') pos[at++] = p0 + i;
}

// row j (j >= 1; row 0 is the header line) spans (nl[j-1], nl[j]): "<k bytes>\t<digits>"; bad[0] counts malformed rows
__global__ void mg_parse_rows_kernel(const u8* __restrict__ text, const u64* __restrict__ nl, u64 nrows, int k, u8* __restrict__ rows,
                                     u64* __restrict__ counts, ull* __restrict__ bad) {
    const u64 j = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= nrows) return;
    const u64 a = nl[j] + 1, b = nl[j + 1];                       // nl[0] ends the header line
    u64 e = b;
    if (e > a && text[e - 1] == '\r') --e;
    bool ok = e >= a + (u64)k + 2 && text[a + k] == '\t';
    u64 v = 0;
    if (ok) {
        for (u64 p = a + k + 1; p < e; ++p) {
            const u32 d = (u32)text[p] - '0';
            if (d > 9) { ok = false; break; }
            v = v * 10 + d;
        }
    }
    if (!ok) { atomicAdd(bad, 1ull); v = 0; }
    for (int i = 0; i < k; ++i) rows[j * k + i] = ok ? text[a + i] : (u8)0;
    counts[j] = v;
}

// ---- packed rows -> k literal bytes ----------------------------------------------------------------------------------
__global__ void mg_decode_rows_kernel(const u64* __restrict__ keys, u64 n, int k, int kind, u8* __restrict__ rows) {
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    u8 t[32];
    tsv_decode(keys[i], k, kind, t);
    for (int j = 0; j < k; ++j) rows[i * k + j] = t[j];
}

// ---- matrix fill: position of every sample row in the (sorted, unique) union -------------------------------------
__global__ void mg_fill_kernel(const u8* __restrict__ uni, u64 nu, int k, const u8* __restrict__ rows, const u64* __restrict__ counts,
                               u64 n, u64* __restrict__ mat, u32 nsamples, u32 sample) {
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const u8* row = rows + i * k;
    u64 lo = 0, hi = nu;
    while (lo < hi) {
        const u64 mid = (lo + hi) >> 1;
        if (tsv_cmp(uni + mid * k, row, k) < 0) lo = mid + 1; else hi = mid;
    }
    if (lo < nu) atomicAdd((ull*)&mat[lo * nsamples + sample], (ull)counts[i]);      // (a k-mer listed twice in one file adds up)
}

// ---- text: one thread per cell -----------------------------------------------------------------------------------
// cell (r, c): value vals[r * stride_r + c * stride_c]; a row's first cell is preceded by its label (label_len bytes at
// labels + r * label_len, may be 0) and its last cell is followed by '\n'.
__global__ void mg_cell_len_kernel(const u64* __restrict__ vals, u64 stride_r, u64 stride_c, u64 nrows, u64 ncols, u32 label_len,
                                   u32* __restrict__ len) {
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nrows * ncols) return;
    const u64 r = i / ncols, c = i % ncols;
    len[i] = 1u + tsv_digits(vals[r * stride_r + c * stride_c]) + (c == 0 ? label_len : 0u) + (c == ncols - 1 ? 1u : 0u);
}

__global__ void mg_cell_write_kernel(const u64* __restrict__ vals, u64 stride_r, u64 stride_c, u64 nrows, u64 ncols, u32 label_len,
                                     const u8* __restrict__ labels, const u64* __restrict__ off, u8* __restrict__ out) {
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nrows * ncols) return;
    const u64 r = i / ncols, c = i % ncols;
    u8* o = out + off[i];
    if (c == 0)
        for (u32 j = 0; j < label_len; ++j) *o++ = labels[r * label_len + j];
    *o++ = '\t';
    u64 v = vals[r * stride_r + c * stride_c];
    const u32 nd = tsv_digits(v);
    for (u32 j = nd; j-- > 0;) { o[j] = (u8)('0' + (u32)(v % 10)); v /= 10; }
    o += nd;
    if (c == ncols - 1) *o = '\n';
}

// ---- row N4: inputs of the alpha-diversity metrics and of the top-k-mer summary ---------------------------------------
// The reference hands every sample's count vector to scikit-bio (lib/mercat2_diversity.py:23-27: counts += [int(line.split()[1])])
// and picks the 5 k-mers with the largest mean count across samples for its summary plot (lib/mercat2_figures.py:50-65).
// Both need reductions over counts only.  spectrum[0] = observed k-mers, [1] = sum of counts, [2..3] = sum of squared
// counts (low, high 64 bits), [4] = largest count, [5 + i] = k-mers seen exactly i + 1 times (i < 10: the singleton /
// doubleton / rare-abundance classes of chao1, goods_coverage and ace).  The arithmetic of the metrics stays with skbio.
#define MG_SPECTRUM_WORDS 16
__global__ void __launch_bounds__(256) mg_spectrum_kernel(const u64* __restrict__ counts, u64 n, unsigned long long* __restrict__ spectrum) {
    __shared__ unsigned long long s[MG_SPECTRUM_WORDS];
    if (threadIdx.x < MG_SPECTRUM_WORDS) s[threadIdx.x] = 0;
    BLOCK_SYNC();
    unsigned long long sum = 0, sq_lo = 0, sq_hi = 0, mx = 0, cls[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0}, rows = 0;
    for (u64 i = (u64)blockIdx.x * 256 + threadIdx.x; i < n; i += (u64)gridDim.x * 256) {
        const unsigned long long c = counts[i];
        ++rows;
        sum += c;
        const unsigned long long lo = c * c, hi = __umul64hi(c, c);
        sq_lo += lo;
        sq_hi += hi + (sq_lo < lo ? 1ull : 0ull);
        mx = max(mx, c);
#pragma unroll
        for (int j = 0; j < 10; ++j) cls[j] += (c == (unsigned long long)(j + 1)) ? 1ull : 0ull;
    }
    atomicAdd(&s[0], rows);
    atomicAdd(&s[1], sum);
    const unsigned long long old = atomicAdd(&s[2], sq_lo);
    atomicAdd(&s[3], sq_hi + (old + sq_lo < old ? 1ull : 0ull));
    atomicMax(&s[4], mx);
#pragma unroll
    for (int j = 0; j < 10; ++j) if (cls[j]) atomicAdd(&s[5 + j], cls[j]);
    BLOCK_SYNC();
    if (threadIdx.x < MG_SPECTRUM_WORDS && threadIdx.x != 2 && threadIdx.x != 3 && threadIdx.x != 4 && s[threadIdx.x]) atomicAdd(&spectrum[threadIdx.x], s[threadIdx.x]);
    if (threadIdx.x == 4) atomicMax(&spectrum[4], s[4]);
    if (threadIdx.x == 2) {
        const unsigned long long o = atomicAdd(&spectrum[2], s[2]);
        atomicAdd(&spectrum[3], s[3] + (o + s[2] < o ? 1ull : 0ull));
    }
}
// sum of every matrix row (counts: rows x samples, row-major)
__global__ void mg_row_sums_kernel(const u64* __restrict__ counts, u64 rows, u32 samples, u64* __restrict__ sums) {
    const u64 r = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rows) return;
    u64 acc = 0;
    for (u32 j = 0; j < samples; ++j) acc += counts[r * samples + j];
    sums[r] = acc;
}
// The `top` rows with the largest sums, earlier rows winning ties (one CTA; top <= 32): every round takes the
// lexicographic maximum of (sum, ~row) over the rows not taken yet.
__global__ void __launch_bounds__(1024) mg_top_rows_kernel(const u64* __restrict__ sums, u64 rows, u32 top, u64* __restrict__ out) {
    __shared__ u64 s_sum[1024], s_row[1024];
    __shared__ u64 taken[32];
    for (u32 t = 0; t < top; ++t) {
        u64 best = 0, best_row = ~0ull;
        for (u64 r = threadIdx.x; r < rows; r += 1024) {
            bool skip = false;
            for (u32 q = 0; q < t; ++q) skip |= taken[q] == r;
            if (skip) continue;
            const u64 v = sums[r];
            if (best_row == ~0ull || v > best) { best = v; best_row = r; }       // (ascending r per thread: ties keep the earlier row)
        }
        s_sum[threadIdx.x] = best;
        s_row[threadIdx.x] = best_row;
        BLOCK_SYNC();
        for (u32 d = 512; d; d >>= 1) {
            if (threadIdx.x < d) {
                const u64 a = s_sum[threadIdx.x], ar = s_row[threadIdx.x], b = s_sum[threadIdx.x + d], br = s_row[threadIdx.x + d];
                const bool take_b = ar == ~0ull || (br != ~0ull && (b > a || (b == a && br < ar)));
                if (take_b) { s_sum[threadIdx.x] = b; s_row[threadIdx.x] = br; }
            }
            BLOCK_SYNC();
        }
        if (threadIdx.x == 0) { taken[t] = s_row[0]; out[t] = s_row[0]; }
        BLOCK_SYNC();
    }
}
