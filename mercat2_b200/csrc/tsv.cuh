// K5 — TSV body formatted on the device.
//
// Replaces the row loop of run_mercat2 (reference bin/mercat2.py:125-134: `for kmer,count in sorted(...)` ->
// f"{kmer}\t{count}\n").  The sample table already sits on the device as two sorted row sets: "fast" rows (a k-mer
// packed into a 64-bit key whose numeric order is its text order) and "wide" rows (k literal bytes).  The two sets
// are disjoint, so the merged position of a row is its own index plus its rank in the other set; row lengths are
// scanned into byte offsets and every row is written in place.  The host only prepends the header line and writes
// the buffer out.
#pragma once
#include "common.cuh"

enum : int { TSV_NT2 = 0, TSV_AA5 = 1, TSV_BYTE = 2, TSV_DENSE_AA = 3 };

__device__ __forceinline__ void tsv_decode(u64 key, int k, int kind, u8* out) {
    if (kind == TSV_NT2) {
        for (int j = k - 1; j >= 0; --j) { out[j] = (u8)((0x54474341u >> (8 * (u32)(key & 3))) & 0xFFu); key >>= 2; }    // "ACGT"
    } else if (kind == TSV_AA5) {
        for (int j = k - 1; j >= 0; --j) { out[j] = (u8)('A' + (u32)(key & 31)); key >>= 5; }
    } else if (kind == TSV_BYTE) {
        for (int j = k - 1; j >= 0; --j) { out[j] = (u8)(key & 255); key >>= 8; }
    } else {
        for (int j = k - 1; j >= 0; --j) { out[j] = (u8)('A' + (u32)(key % 26)); key /= 26; }
    }
}

__device__ __forceinline__ int tsv_cmp(const u8* a, const u8* b, int k) {
    for (int j = 0; j < k; ++j) {
        const int d = (int)a[j] - (int)b[j];
        if (d) return d;
    }
    return 0;
}

__device__ __forceinline__ u32 tsv_digits(u64 v) {
    u32 n = 1;
    while (v >= 10) { v /= 10; ++n; }
    return n;
}

// Merged position and byte length of every row.  Threads [0, nf) handle fast rows, [nf, nf + nw) wide rows.
// pos[] is indexed by the row's own id (fast rows first), len[] by the merged position.
__global__ void tsv_place_kernel(const u64* __restrict__ fkeys, const u64* __restrict__ fcnt, u64 nf, const u8* __restrict__ wrows,
                                 const u64* __restrict__ wcnt, u64 nw, int k, int kind, u64* __restrict__ pos, u32* __restrict__ len) {
    const u64 t = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nf + nw) return;
    u8 text[32];
    u64 p, count;
    if (t < nf) {
        u64 lo = 0, hi = nw;                                  // wide rows smaller than this fast row
        if (nw) {
            tsv_decode(fkeys[t], k, kind, text);
            while (lo < hi) {
                const u64 mid = (lo + hi) >> 1;
                if (tsv_cmp(wrows + mid * k, text, k) < 0) lo = mid + 1; else hi = mid;
            }
        }
        p = t + lo;
        count = fcnt[t];
    } else {
        const u64 j = t - nf;
        const u8* row = wrows + j * k;
        u64 lo = 0, hi = nf;                                  // fast rows smaller than this wide row (fast keys have k <= 32)
        while (lo < hi) {
            const u64 mid = (lo + hi) >> 1;
            tsv_decode(fkeys[mid], k, kind, text);
            if (tsv_cmp(text, row, k) < 0) lo = mid + 1; else hi = mid;
        }
        p = j + lo;
        count = wcnt[j];
    }
    pos[t] = p;
    len[p] = (u32)k + 2u + tsv_digits(count);
}

// One thread per row: k-mer, tab, decimal count, newline at the row's byte offset.
__global__ void tsv_write_kernel(const u64* __restrict__ fkeys, const u64* __restrict__ fcnt, u64 nf, const u8* __restrict__ wrows,
                                 const u64* __restrict__ wcnt, u64 nw, int k, int kind, const u64* __restrict__ pos,
                                 const u64* __restrict__ off, u8* __restrict__ out) {
    const u64 t = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nf + nw) return;
    u8* o = out + off[pos[t]];
    u64 count;
    if (t < nf) {
        u8 text[32];
        tsv_decode(fkeys[t], k, kind, text);
        for (int j = 0; j < k; ++j) o[j] = text[j];
        count = fcnt[t];
    } else {
        const u8* row = wrows + (t - nf) * k;
        for (int j = 0; j < k; ++j) o[j] = row[j];
        count = wcnt[t - nf];
    }
    o += k;
    *o++ = '\t';
    const u32 nd = tsv_digits(count);
    for (u32 j = nd; j-- > 0;) { o[j] = (u8)('0' + (u32)(count % 10)); count /= 10; }
    o[nd] = '\n';
}
