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
