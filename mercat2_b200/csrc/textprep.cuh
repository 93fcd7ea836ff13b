// Rows N1 / N2: the text transforms that run immediately before the hot path, on the device, so that the
// cleaned text is already in HBM when counting starts.
//
//   FASTQ -> FASTA   lib/mercat2_fasta.py:175-198 (`sed -n '1~4s/^@/>/p;2~4p'`): of every four '\n'-lines keep the first
//                    with its leading '@' turned into '>' (dropped when it does not start with '@') and the second.
//                    Bytes are copied as they are ('\r' stays part of a line, a last line without '\n' stays without).
#pragma once
#include "common.cuh"

// length each line contributes to the FASTA text; line j = (nl[j-1], nl[j]] (the '\n' included), the last line may
// end at n without one
__global__ void fq_line_len_kernel(const u8* __restrict__ text, u64 n, const u64* __restrict__ nl, u64 n_nl, u64 nlines, u32* __restrict__ len) {
    const u64 j = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= nlines) return;
    const u64 a = j ? nl[j - 1] + 1 : 0;
    const u64 b = j < n_nl ? nl[j] + 1 : n;
    u32 out = 0;
    if ((j & 3) == 0) out = (b > a && text[a] == '@') ? (u32)(b - a) : 0u;
    else if ((j & 3) == 1) out = (u32)(b - a);
    len[j] = out;
}
// one warp per line
__global__ void __launch_bounds__(256)
fq_copy_kernel(const u8* __restrict__ text, u64 n, const u64* __restrict__ nl, u64 n_nl, u64 nlines, const u32* __restrict__ len,
               const u64* __restrict__ off, u8* __restrict__ out) {
    const u32 lane = threadIdx.x & 31;
    const u64 j = (u64)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (j >= nlines) return;
    const u32 m = len[j];
    if (!m) return;
    const u64 a = j ? nl[j - 1] + 1 : 0;
    u8* dst = out + off[j];
    for (u32 i = lane; i < m; i += 32) {
        u8 c = text[a + i];
        if (i == 0 && (j & 3) == 0) c = '>';
        dst[i] = c;
    }
}
