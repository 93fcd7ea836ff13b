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
