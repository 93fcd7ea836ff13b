// K6 -- per-record protein metrics on device.
//
// Restates the numeric part of plot_sample_metrics (lib/mercat2_figures.py:157-183) and the three
// functions of lib/mercat2_metrics.py:
//   * record parsing: every line is strip()-ed, then rstrip('*')-ed (interior '*' stay); a line that
//     then starts with '>' opens a record named line[1:]; text before the first header is ignored;
//   * pI: ProMoST bisection (metrics.py:57-101) -- depends only on the first and last residue and on
//     the counts of D, E, C, Y, H, K, R;  * MW: sum of residue masses + water (metrics.py:158-163);
//   * hydropathy: Kyte-Doolittle sum (metrics.py:166-170).
// Three stages: line index (positions after each terminator), one thread per LINE (strip, classify,
// per-line partial sums), one thread per RECORD (sum its lines in order, bisection in fp64).
#pragma once
#include "common.cuh"

// pK used for the first residue: ionisable table column 2, else terminal table column 1
__device__ const double g_pk_first[26] = {
    /*A*/ 3.75, /*B*/ 3.57, /*C*/ 9.00, /*D*/ 4.57, /*E*/ 4.75, /*F*/ 3.98, /*G*/ 3.70, /*H*/ 6.89, /*I*/ 3.72,
    /*J*/ 3.73, /*K*/ 10.30, /*L*/ 3.73, /*M*/ 3.68, /*N*/ 3.64, /*O*/ 3.50, /*P*/ 3.40, /*Q*/ 3.57, /*R*/ 11.50,
    /*S*/ 3.61, /*T*/ 3.57, /*U*/ 5.60, /*V*/ 3.69, /*W*/ 3.78, /*X*/ 3.57, /*Y*/ 10.34, /*Z*/ 3.535};
// pK used for the last residue: ionisable table column 0, else terminal table column 0
__device__ const double g_pk_last[26] = {
    /*A*/ 7.58, /*B*/ 7.46, /*C*/ 8.00, /*D*/ 3.57, /*E*/ 4.15, /*F*/ 6.96, /*G*/ 7.50, /*H*/ 4.89, /*I*/ 7.48,
    /*J*/ 7.46, /*K*/ 10.00, /*L*/ 7.46, /*M*/ 6.98, /*N*/ 7.22, /*O*/ 7.00, /*P*/ 8.36, /*Q*/ 6.73, /*R*/ 11.50,
    /*S*/ 6.86, /*T*/ 7.02, /*U*/ 5.20, /*V*/ 7.44, /*W*/ 7.11, /*X*/ 7.26, /*Y*/ 9.34, /*Z*/ 6.96};
__device__ const double g_mass[26] = {
    /*A*/ 71.0788, /*B*/ 114.6686, /*C*/ 103.1388, /*D*/ 115.0886, /*E*/ 129.1155, /*F*/ 147.1766, /*G*/ 57.0519,
    /*H*/ 137.1411, /*I*/ 113.1594, /*J*/ 0.0, /*K*/ 128.1741, /*L*/ 113.1594, /*M*/ 131.1926, /*N*/ 114.1038,
    /*O*/ 237.3018, /*P*/ 97.1167, /*Q*/ 128.1307, /*R*/ 156.1875, /*S*/ 87.0782, /*T*/ 101.1051, /*U*/ 150.0388,
    /*V*/ 99.1326, /*W*/ 186.2132, /*X*/ 111.1138, /*Y*/ 163.176, /*Z*/ 128.7531};
__device__ const double g_hydro[26] = {
    /*A*/ 1.8, /*B*/ 0.0, /*C*/ 2.5, /*D*/ -3.5, /*E*/ -3.5, /*F*/ 2.8, /*G*/ -0.4, /*H*/ -3.2, /*I*/ 4.5, /*J*/ 0.0,
    /*K*/ -3.9, /*L*/ 3.8, /*M*/ 1.9, /*N*/ -3.5, /*O*/ 0.0, /*P*/ -1.6, /*Q*/ -3.5, /*R*/ -4.5, /*S*/ -0.8, /*T*/ -0.7,
    /*U*/ 0.0, /*V*/ 4.2, /*W*/ -0.9, /*X*/ 0.0, /*Y*/ -1.3, /*Z*/ 0.0};

__device__ __forceinline__ bool mt_is_nl(u32 c) { return c == 10u || c == 13u; }
__device__ __forceinline__ bool mt_is_ws(u32 c) { return (c >= 9u && c <= 13u) || (c >= 28u && c <= 32u); }

// line starts: position 0 and every position following a terminator
template <bool WRITE>
__global__ void __launch_bounds__(256)
mt_lines_kernel(const u8* __restrict__ text, u64 n, u32* __restrict__ tile_cnt, const u64* __restrict__ tile_off, u64* __restrict__ line_off) {
    __shared__ u32 sm[256 / 32 + 1];
    const u64 p = (u64)blockIdx.x * 256 + threadIdx.x;
    const bool is_start = p < n && (p == 0 || mt_is_nl(text[p - 1]));
    if (!WRITE) {
        const u32 total = block_count(is_start);
        if (threadIdx.x == 0) tile_cnt[blockIdx.x] = total;
    } else {
        const u32 off = block_exclusive_scan<OpAdd, 8>(is_start ? 1u : 0u, sm, nullptr);
        if (is_start) line_off[tile_off[blockIdx.x] + off] = p;
    }
}

struct LineStat {
    double mass, hydro;
    u64 hdr_off;       // header lines: offset of the text after '>'
    u32 hdr_len;
    u32 kept;          // sequence lines: symbols kept
    u32 cnt[7];        // D E C Y H K R
    u8 first, last, is_header, pad;
};

__global__ void mt_line_stats_kernel(const u8* __restrict__ text, u64 n, const u64* __restrict__ line_off, u64 nlines,
                                     LineStat* __restrict__ out) {
    const u64 l = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= nlines) return;
    const u64 s = line_off[l];
    u64 e = (l + 1 < nlines) ? line_off[l + 1] - 1 : n;          // exclusive end, terminator dropped
    if (l + 1 >= nlines && e > s && mt_is_nl(text[e - 1])) --e;  // last line may end with a terminator
    u64 a = s, b = e;
    while (a < b && mt_is_ws(text[a])) ++a;
    while (b > a && mt_is_ws(text[b - 1])) --b;
    while (b > a && text[b - 1] == '*') --b;
    LineStat st;
    st.mass = 0.0; st.hydro = 0.0; st.hdr_off = 0; st.hdr_len = 0; st.kept = 0;
    for (int i = 0; i < 7; ++i) st.cnt[i] = 0;
    st.first = 0; st.last = 0; st.is_header = 0; st.pad = 0;
    if (a < b && text[a] == '>') {
        st.is_header = 1;
        st.hdr_off = a + 1;
        st.hdr_len = (u32)(b - a - 1);
    } else if (a < b) {
        st.kept = (u32)(b - a);
        st.first = text[a];
        st.last = text[b - 1];
        double mass = 0.0, hydro = 0.0;
        u32 cD = 0, cE = 0, cC = 0, cY = 0, cH = 0, cK = 0, cR = 0;
        for (u64 p = a; p < b; ++p) {
            const u32 c = text[p];
            const u32 d = c - 65u;
            if (d < 26u) { mass += g_mass[d]; hydro += g_hydro[d]; }
            cD += c == 'D'; cE += c == 'E'; cC += c == 'C'; cY += c == 'Y'; cH += c == 'H'; cK += c == 'K'; cR += c == 'R';
        }
        st.mass = mass; st.hydro = hydro;
        st.cnt[0] = cD; st.cnt[1] = cE; st.cnt[2] = cC; st.cnt[3] = cY; st.cnt[4] = cH; st.cnt[5] = cK; st.cnt[6] = cR;
    }
    out[l] = st;
}

template <bool WRITE>
__global__ void __launch_bounds__(256)
mt_headers_kernel(const LineStat* __restrict__ ls, u64 nlines, u32* __restrict__ tile_cnt, const u64* __restrict__ tile_off,
                  u64* __restrict__ hdr_line) {
    __shared__ u32 sm[256 / 32 + 1];
    const u64 l = (u64)blockIdx.x * 256 + threadIdx.x;
    const bool h = l < nlines && ls[l].is_header;
    if (!WRITE) {
        const u32 total = block_count(h);
        if (threadIdx.x == 0) tile_cnt[blockIdx.x] = total;
    } else {
        const u32 off = block_exclusive_scan<OpAdd, 8>(h ? 1u : 0u, sm, nullptr);
        if (h) hdr_line[tile_off[blockIdx.x] + off] = l;
    }
}

struct RecordOut {
    u64 hdr_off;
    u64 length;
    double pi, mw, hydro;
    u32 hdr_len;
    u32 status;     // 0 ok, 1 last residue unknown, 2 first residue unknown, 255 empty sequence (skipped)
};

// length/first/last/counts/sums of one sequence -> MW, hydropathy and the ProMoST bisection
__device__ __forceinline__ void mt_finish_record(RecordOut& o, u64 length, u32 first, u32 last, const u64 cnt[7], double mass,
                                                 double hydro) {
    o.length = length;
    o.mw = mass + 18.01524;
    o.hydro = hydro;
    o.pi = 0.0;
    o.status = 0;
    if (!length) { o.status = 255; return; }
    const u32 fi = first - 65u, la = last - 65u;
    if (fi >= 26u) { o.status = 2; return; }
    if (la >= 26u) { o.status = 1; return; }
    const double pk_first = g_pk_first[fi], pk_last = g_pk_last[la];
    const double nD = (double)cnt[0], nE = (double)cnt[1], nC = (double)cnt[2], nY = (double)cnt[3];
    const double nH = (double)cnt[4], nK = (double)cnt[5], nR = (double)cnt[6];
    double ph = 6.51, lo = 0.0, hi = 14.0;
    for (int it = 0; it < 64; ++it) {
        double q = -1.0 / (1.0 + pow(10.0, pk_first - ph));
        q += -nD / (1.0 + pow(10.0, 4.07 - ph));
        q += -nE / (1.0 + pow(10.0, 4.45 - ph));
        q += -nC / (1.0 + pow(10.0, 8.28 - ph));
        q += -nY / (1.0 + pow(10.0, 9.84 - ph));
        q += nH / (1.0 + pow(10.0, ph - 6.08));
        q += 1.0 / (1.0 + pow(10.0, ph - pk_last));
        q += nK / (1.0 + pow(10.0, ph - 9.80));
        q += nR / (1.0 + pow(10.0, ph - 12.50));
        if (q < 0.0) { const double t = ph; ph = ph - (ph - lo) / 2.0; hi = t; }
        else { const double t = ph; ph = ph + (hi - ph) / 2.0; lo = t; }
        if (ph - lo < 0.01 && hi - ph < 0.01) break;
    }
    o.pi = ph;
}

__global__ void mt_records_kernel(const LineStat* __restrict__ ls, u64 nlines, const u64* __restrict__ hdr_line, u64 nrec,
                                  RecordOut* __restrict__ out) {
    const u64 r = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= nrec) return;
    const u64 h = hdr_line[r];
    const u64 end = (r + 1 < nrec) ? hdr_line[r + 1] : nlines;
    RecordOut o;
    o.hdr_off = ls[h].hdr_off;
    o.hdr_len = ls[h].hdr_len;
    u64 length = 0;
    double mass = 0.0, hydro = 0.0;
    u64 cnt[7] = {0, 0, 0, 0, 0, 0, 0};
    u32 first = 0, last = 0;
    for (u64 l = h + 1; l < end; ++l) {
        const LineStat s = ls[l];
        if (!s.kept) continue;
        if (!length) first = s.first;
        last = s.last;
        length += s.kept;
        mass += s.mass;
        hydro += s.hydro;
        for (int i = 0; i < 7; ++i) cnt[i] += s.cnt[i];
    }
    mt_finish_record(o, length, first, last, cnt, mass, hydro);
    out[r] = o;
}

// one thread per raw sequence (no FASTA parsing): the scalar entry points of lib/mercat2_metrics.py
__global__ void mt_sequences_kernel(const u8* __restrict__ seqs, const u64* __restrict__ offs, u64 nseq, RecordOut* __restrict__ out) {
    const u64 r = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= nseq) return;
    const u64 a = offs[r], b = offs[r + 1];
    RecordOut o;
    o.hdr_off = 0; o.hdr_len = 0;
    double mass = 0.0, hydro = 0.0;
    u64 cnt[7] = {0, 0, 0, 0, 0, 0, 0};
    for (u64 p = a; p < b; ++p) {
        const u32 c = seqs[p];
        const u32 d = c - 65u;
        if (d < 26u) { mass += g_mass[d]; hydro += g_hydro[d]; }
        cnt[0] += c == 'D'; cnt[1] += c == 'E'; cnt[2] += c == 'C'; cnt[3] += c == 'Y';
        cnt[4] += c == 'H'; cnt[5] += c == 'K'; cnt[6] += c == 'R';
    }
    mt_finish_record(o, b - a, b > a ? seqs[a] : 0u, b > a ? seqs[b - 1] : 0u, cnt, mass, hydro);
    out[r] = o;
}

