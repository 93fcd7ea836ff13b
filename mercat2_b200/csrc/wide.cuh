// K7 -- the wide (exception) path: k-mers whose code does not fit the fast 64-bit encodings.
//
// The reference counts every window under its literal text (lib/mercat2_kmers.py:56-60: dict keys
// are str slices; 'N', lower case, digits, interior blanks are all ordinary characters).  Here such
// windows are identified by their START position in a byte buffer; they are ordered by an LSD radix
// sort over 8-byte big-endian limbs gathered on the fly (so numeric limb order == byte order ==
// Python str order for ASCII), then run-length encoded with the same threshold kernels as the fast
// path (radix.cuh, WindowEq).
#pragma once
#include "common.cuh"

// keys[i] = limb `limb` (bytes 8*limb .. 8*limb+7, zero padded past k) of window idx[i]
__global__ void wide_gather_limb_kernel(const u8* __restrict__ src, const u64* __restrict__ pos, const u32* __restrict__ idx,
                                        u64 m, int k, int limb, u64* __restrict__ keys) {
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    const u8* p = src + pos[idx[i]] + 8 * limb;
    const int nb = min(8, k - 8 * limb);
    u64 v = 0;
    for (int j = 0; j < 8; ++j) v = (v << 8) | (j < nb ? (u64)p[j] : 0ull);
    keys[i] = v;
}

// rows[r*k .. r*k+k) = bytes of the window that heads surviving run r
__global__ void wide_gather_rows_kernel(const u8* __restrict__ src, const u64* __restrict__ pos, const u32* __restrict__ idx,
                                        const u64* __restrict__ start, u64 nrows, int k, u8* __restrict__ rows) {
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nrows * (u64)k) return;
    const u64 r = i / k;
    const int j = (int)(i % k);
    rows[i] = src[pos[idx[start[r]]] + j];
}

__global__ void wide_row_positions_kernel(u64* pos, u64 n, int k) {
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) pos[i] = i * (u64)k;
}

// ---- literal rows that the sample's packed alphabet can express (mc2_sample_add_rows: tables counted elsewhere) --------
// A table merged into a sample arrives as k literal bytes per row.  A row whose k bytes all belong to the sample's fast
// alphabet is the SAME k-mer as a packed row the sample may already hold, so it must join the packed rows (or the dense
// table) to be summed with them; only true exceptions stay literal.  enc: 0 ACGT, 1 'A'..'Z', 2 any ASCII byte.
__device__ __forceinline__ bool row_encode(const u8* row, int k, int enc, bool dense_aa, u64& code) {
    u64 c = 0;
    for (int j = 0; j < k; ++j) {
        const u32 b = row[j];
        if (enc == 0) {
            if (!(b == 'A' || b == 'C' || b == 'G' || b == 'T')) return false;
            c = (c << 2) | (((b >> 1) ^ (b >> 2)) & 3u);
        } else if (enc == 1) {
            if (b < 'A' || b > 'Z') return false;
            c = dense_aa ? c * 26u + (b - 'A') : ((c << 5) | (b - 'A'));
        } else {
            if (b >= 128u) return false;
            c = (c << 8) | b;
        }
    }
    code = c;
    return true;
}
__global__ void rows_flag_kernel(const u8* __restrict__ rows, u64 n, int k, int enc, u32* __restrict__ flag) {
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    u64 code;
    flag[i] = row_encode(rows + i * (u64)k, k, enc, false, code) ? 1u : 0u;
}
// pos[i] = number of encodable rows before row i.  Encodable rows go to (fkeys, fcounts) -- or, with `dense`, straight
// into the per-sample dense table -- the others are compacted into (wrows, wcounts).
__global__ void rows_split_kernel(const u8* __restrict__ rows, const u64* __restrict__ counts, u64 n, int k, int enc, const u32* __restrict__ flag,
                                  const u64* __restrict__ pos, u64* __restrict__ fkeys, u64* __restrict__ fcounts, unsigned long long* __restrict__ dense,
                                  u8* __restrict__ wrows, u64* __restrict__ wcounts) {
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const u8* row = rows + i * (u64)k;
    if (flag[i]) {
        u64 code;
        row_encode(row, k, enc, dense != nullptr && enc == 1, code);
        if (dense) atomicAdd(&dense[code], (unsigned long long)counts[i]);
        else { fkeys[pos[i]] = code; fcounts[pos[i]] = counts[i]; }
    } else {
        const u64 o = i - pos[i];
        for (int j = 0; j < k; ++j) wrows[o * (u64)k + j] = row[j];
        wcounts[o] = counts[i];
    }
}
