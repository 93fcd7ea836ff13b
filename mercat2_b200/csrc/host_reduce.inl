// Part of mc2.cu (textually included there, in this order): K1 host side (parse driver), path planning, reductions over sorted sequences.

// =====================================================================================================
// K1 parse
// =====================================================================================================
struct Parsed {
    DBuf<u8> sym;
    u64 nsym = 0;
    ParseStats stats;
};

static void parse_text(mc2_engine* e, const u8* dtext, u64 len, int toupper, Parsed& out, DBuf<u64>* tile_off_out = nullptr) {
    memset(&out.stats, 0, sizeof out.stats);
    out.nsym = 0;
    if (len == 0) return;
    const u64 mis = (u64)(uintptr_t)dtext & 15ull;
    const u64 ntiles = div_up(mis + len, PARSE_TILE);
    if (ntiles >= (1ull << 31)) throw Mc2Error(MC2_ERR_LIMIT, "parse: text larger than 8 TiB");
    DBuf<u8> tf(e, ntiles), tr(e, ntiles);
    DBuf<u32> cnt(e, ntiles);
    DBuf<u64> off(e, ntiles);
    DBuf<ParseStats> st(e, 1);
    st.zero();
    LAUNCH(e, parse_summarize_kernel, (unsigned)ntiles, PARSE_THREADS, 0, dtext, len, tf.p, tr.p, st.p);
    LAUNCH(e, parse_scan_tiles_kernel, 1, SCAN1_THREADS, 0, tf.p, tr.p, (u32)ntiles);
    LAUNCH(e, parse_emit_kernel<false>, (unsigned)ntiles, PARSE_THREADS, 0, dtext, len, tf.p, tr.p, cnt.p,
           (const u64*)nullptr, (u8*)nullptr, st.p, toupper);
    dev_exclusive_scan<u32, u64>(e, cnt.p, off.p, ntiles, &st.p->n_sym);
    out.stats = read_scalar<ParseStats>(e, st.p);
    if (out.stats.n_bad)
        throw Mc2Error(MC2_ERR_NON_ASCII, "input contains " + std::to_string(out.stats.n_bad) +
                                              " non-ASCII byte(s); only 7-bit ASCII FASTA text is supported");
    out.nsym = out.stats.n_sym;
    out.sym.alloc(e, (out.nsym + 64) & ~15ull);
    if (out.nsym)
        LAUNCH(e, parse_emit_kernel<true>, (unsigned)ntiles, PARSE_THREADS, 0, dtext, len, tf.p, tr.p, cnt.p,
               (const u64*)off.p, out.sym.p, st.p, toupper);
    if (tile_off_out) *tile_off_out = std::move(off);       // symbol index of every 4 KiB text tile's first symbol
}

// =====================================================================================================
// planning
// =====================================================================================================
static int enc_bits(int enc) { return enc == ENC_NT2 ? 2 : enc == ENC_AA5 ? 5 : 8; }

static void make_plan(mc2_engine* e, const ParseStats& st, int k, Plan& plan) {
    int enc;
    if (e->opt_force_enc >= 0) enc = e->opt_force_enc;
    else if (st.n_ascii == 0 || st.n_acgt * 10 >= st.n_ascii * 9) enc = ENC_NT2;
    else if (st.n_upper * 2 >= st.n_ascii) enc = ENC_AA5;
    else enc = ENC_BYTE;
    plan.enc = enc;
    const int kb = k * enc_bits(enc);
    int path = kb <= 64 ? PATH_SPARSE : PATH_WIDE;
    u64 bins = 0;
    if (path == PATH_SPARSE && enc != ENC_BYTE) {
        const u64 base = enc == ENC_NT2 ? 4 : 26;
        bins = 1;
        for (int i = 0; i < k && bins <= (1ull << 40); ++i) bins *= base;
        if (bins <= e->opt_dense_max_bins && bins < (1ull << 31)) path = PATH_DENSE;
        // a table much larger than the first chunk costs more to clear, fold and scan than the chunk costs to count
        // sparsely (26^5 bins = 143 MB of tables for a 1.6 M-residue proteome): the range path takes such samples
        if (path == PATH_DENSE && bins > e->opt_smem_max_bins && st.n_ascii * 4 < bins && e->opt_force_path != PATH_DENSE) path = PATH_SPARSE;
    }
    if (e->opt_force_path == PATH_SPARSE && kb <= 64) path = PATH_SPARSE;
    if (e->opt_force_path == PATH_WIDE) path = PATH_WIDE;
    if (e->opt_force_path == PATH_DENSE && bins && bins < (1ull << 31)) path = PATH_DENSE;
    plan.path = path;
    plan.bins = path == PATH_DENSE ? (u32)bins : 0;
    plan.smem = path == PATH_DENSE && bins <= e->opt_smem_max_bins && bins * 4 <= 200 * 1024;
    plan.nrep = 1;
    if (plan.smem) {
        u64 r = (48 * 1024) / (bins * 4);
        plan.nrep = (u32)std::min<u64>(EX_WARPS, std::max<u64>(1, r));
    }
}

// =====================================================================================================
// reductions over sorted sequences
// =====================================================================================================
// unit weights: survivors of a sorted sequence of m items; returns count and fills start/count buffers
template <class Acc>
static u64 rle_threshold(mc2_engine* e, Acc acc, u64 m, u64 c, DBuf<u64>& start, DBuf<u64>& count) {
    const u64 ntiles = div_up(m, RLE_THREADS);
    DBuf<u32> tc(e, ntiles);
    DBuf<u64> to(e, ntiles);
    LAUNCH(e, rle_count_kernel<Acc>, (unsigned)ntiles, RLE_THREADS, 0, acc, m, c, tc.p);
    const u64 ns = offsets_from_counts(e, tc.p, to.p, ntiles);
    start.alloc(e, ns);
    count.alloc(e, ns);
    if (ns) LAUNCH(e, rle_write_kernel<Acc>, (unsigned)ntiles, RLE_THREADS, 0, acc, m, c, (const u64*)to.p, start.p, count.p);
    return ns;
}

// weighted: sorted items with weights w_sorted[i]; survivors have weight sum >= c
template <class Acc>
static u64 seg_reduce(mc2_engine* e, Acc acc, u64 m, const u64* w_sorted, u64 c, DBuf<u64>& start, DBuf<u64>& count) {
    const u64 ntiles = div_up(m, RLE_THREADS);
    DBuf<u32> tc(e, ntiles);
    DBuf<u64> to(e, ntiles);
    LAUNCH(e, seg_count_kernel<Acc>, (unsigned)ntiles, RLE_THREADS, 0, acc, m, tc.p);
    const u64 nseg = offsets_from_counts(e, tc.p, to.p, ntiles);
    DBuf<u64> seg_start(e, nseg), seg_sum(e, nseg);
    LAUNCH(e, seg_write_kernel<Acc>, (unsigned)ntiles, RLE_THREADS, 0, acc, m, (const u64*)to.p, seg_start.p);
    DBuf<u64> wprefix(e, m);
    DBuf<ull> wtotal(e, 1);
    dev_exclusive_scan<u64, u64>(e, w_sorted, wprefix.p, m, wtotal.p);
    const u64 wt = (u64)read_scalar<ull>(e, wtotal.p);
    const u64 nt2 = div_up(nseg, RLE_THREADS);
    DBuf<u32> tc2(e, nt2);
    DBuf<u64> to2(e, nt2);
    LAUNCH(e, seg_sum_count_kernel, (unsigned)nt2, RLE_THREADS, 0, (const u64*)seg_start.p, nseg, m, (const u64*)wprefix.p, wt, c,
           seg_sum.p, tc2.p);
    const u64 ns = offsets_from_counts(e, tc2.p, to2.p, nt2);
    start.alloc(e, ns);
    count.alloc(e, ns);
    if (ns)
        LAUNCH(e, seg_compact_kernel, (unsigned)nt2, RLE_THREADS, 0, (const u64*)seg_start.p, (const u64*)seg_sum.p, nseg, c,
               (const u64*)to2.p, start.p, count.p);
    return ns;
}

// Parts that are each sorted and whose key ranges follow each other without overlap (the groups of a range-partitioned
// chunk, in group order): the merged table is their concatenation.  Returns false when the parts are not of that kind.
static bool concat_sorted_parts(mc2_engine* e, std::vector<FastPart>& parts, u64 M, FastPart& out) {
    std::vector<const u64*> ptrs;
    std::vector<u64> ns;
    for (auto& p : parts) {
        if (!p.n) continue;
        if (!p.sorted) return false;
        ptrs.push_back(p.keys.p);
        ns.push_back(p.n);
    }
    const u32 np = (u32)ptrs.size();
    DBuf<const u64*> dptr(e, np);
    DBuf<u64> dn(e, np), dends(e, 2ull * np);
    CUDA_CHECK(cudaMemcpyAsync(dptr.p, ptrs.data(), np * sizeof(u64*), cudaMemcpyHostToDevice, e->stream));
    CUDA_CHECK(cudaMemcpyAsync(dn.p, ns.data(), np * 8ull, cudaMemcpyHostToDevice, e->stream));
    LAUNCH(e, part_ends_kernel, (unsigned)div_up(np, 128), 128, 0, (const u64* const*)dptr.p, (const u64*)dn.p, np, dends.p);
    std::vector<u64> ends(2ull * np);
    d2h(e, ends.data(), (const u64*)dends.p, 2ull * np);
    for (u32 i = 1; i < np; ++i)
        if (ends[2 * i] <= ends[2 * i - 1]) return false;              // first key of part i must exceed the last key of part i-1
    out.n = M;
    out.sorted = true;
    out.keys.alloc(e, M);
    out.counts.alloc(e, M);
    u64 at = 0;
    for (auto& p : parts) {
        if (!p.n) continue;
        CUDA_CHECK(cudaMemcpyAsync(out.keys.p + at, p.keys.p, p.n * 8, cudaMemcpyDeviceToDevice, e->stream));
        CUDA_CHECK(cudaMemcpyAsync(out.counts.p + at, p.counts.p, p.n * 8, cudaMemcpyDeviceToDevice, e->stream));
        at += p.n;
        p.keys.release();
        p.counts.release();
        p.n = 0;
    }
    return true;
}

static bool merge_rows_range(mc2_engine* e, const u64* keys, const u64* counts, u64 M, int key_bits, FastPart& out);   // host_count.inl
static u64 merge_rows_capacity(const mc2_engine* e);                                                                     // host_count.inl
__global__ void lower_bound_kernel(const u64* __restrict__ keys, u64 n, const u64* __restrict__ q, u64 m, u64* __restrict__ out);

// More rows than one two-level pass takes, every part sorted (the filtered tables of the many pieces of a large sample):
// cut ALL parts at the same splitter keys (quantiles of the largest part), sum each key range on its own and put the
// results behind each other -- ranges are disjoint and ascending, so the concatenation is the sorted table.
static bool merge_sorted_parts_by_range(mc2_engine* e, std::vector<FastPart>& parts, u64 M, int key_bits, FastPart& out) {
    const u64 cap = merge_rows_capacity(e);
    std::vector<FastPart*> live;
    for (auto& p : parts) {
        if (!p.n) continue;
        if (!p.sorted) return false;
        live.push_back(&p);
    }
    const u32 S = (u32)std::min<u64>(4096, div_up(M, std::max<u64>(1, cap / 2)));         // ranges
    if (S < 2 || live.size() < 2) return false;
    FastPart* big = live[0];
    for (auto* p : live) if (p->n > big->n) big = p;
    std::vector<u64> pick(S - 1), split(S - 1);
    for (u32 j = 1; j < S; ++j) pick[j - 1] = (u64)(((unsigned __int128)big->n * j) / S);
    DBuf<u64> dpick(e, S - 1), dsplit(e, S - 1), dcut(e, S - 1);
    CUDA_CHECK(cudaMemcpyAsync(dpick.p, pick.data(), (S - 1) * 8ull, cudaMemcpyHostToDevice, e->stream));
    LAUNCH(e, gather_u64_kernel, (unsigned)div_up(S - 1, 256), 256, 0, (const u64*)big->keys.p, (const u64*)dpick.p, (u64)(S - 1), dsplit.p);
    std::vector<std::vector<u64>> cuts(live.size(), std::vector<u64>(S + 1, 0));
    for (size_t i = 0; i < live.size(); ++i) {
        LAUNCH(e, lower_bound_kernel, (unsigned)div_up(S - 1, 128), 128, 0, (const u64*)live[i]->keys.p, (u64)live[i]->n, (const u64*)dsplit.p, (u64)(S - 1), dcut.p);
        d2h(e, cuts[i].data() + 1, (const u64*)dcut.p, (u64)(S - 1));
        cuts[i][S] = live[i]->n;
        for (u32 j = 1; j <= S; ++j) cuts[i][j] = std::max(cuts[i][j], cuts[i][j - 1]);   // (equal splitters: keep the cuts monotone)
    }
    u64 biggest = 0;
    for (u32 j = 0; j < S; ++j) {
        u64 m = 0;
        for (size_t i = 0; i < live.size(); ++i) m += cuts[i][j + 1] - cuts[i][j];
        biggest = std::max(biggest, m);
    }
    if (biggest > cap) return false;                           // (parts too unlike each other: the caller sorts)
    std::vector<FastPart> done(S);
    DBuf<u64> k0(e, biggest), v0(e, biggest);
    u64 total = 0;
    for (u32 j = 0; j < S; ++j) {
        u64 m = 0;
        for (size_t i = 0; i < live.size(); ++i) {
            const u64 a = cuts[i][j], n = cuts[i][j + 1] - a;
            if (!n) continue;
            CUDA_CHECK(cudaMemcpyAsync(k0.p + m, live[i]->keys.p + a, n * 8, cudaMemcpyDeviceToDevice, e->stream));
            CUDA_CHECK(cudaMemcpyAsync(v0.p + m, live[i]->counts.p + a, n * 8, cudaMemcpyDeviceToDevice, e->stream));
            m += n;
        }
        if (!m) continue;
        if (!merge_rows_range(e, k0.p, v0.p, m, key_bits, done[j])) {
            // a range too small for the range pass (or one that overflowed a table): the sort-based sum of just this range
            DBuf<u64> k1(e, m), v1(e, m);
            const int r = radix_sort<u64, true>(e, k0.p, k1.p, v0.p, v1.p, m, 0, std::min(64, (key_bits + 7) & ~7));
            const u64* ks = r ? k1.p : k0.p;
            const u64* vs = r ? v1.p : v0.p;
            DBuf<u64> start, count;
            KeyEq acc{ks};
            const u64 ns = seg_reduce(e, acc, m, vs, 1, start, count);
            done[j].n = ns;
            done[j].keys.alloc(e, ns);
            done[j].counts = std::move(count);
            if (ns) LAUNCH(e, gather_u64_kernel, (unsigned)div_up(ns, 256), 256, 0, ks, (const u64*)start.p, ns, done[j].keys.p);
            CUDA_CHECK(cudaStreamSynchronize(e->stream));
        }
        total += done[j].n;
    }
    for (auto* p : live) { p->keys.release(); p->counts.release(); p->n = 0; }
    out.n = total;
    out.sorted = true;
    out.keys.alloc(e, total);
    out.counts.alloc(e, total);
    u64 at = 0;
    for (u32 j = 0; j < S; ++j) {
        if (!done[j].n) continue;
        CUDA_CHECK(cudaMemcpyAsync(out.keys.p + at, done[j].keys.p, done[j].n * 8, cudaMemcpyDeviceToDevice, e->stream));
        CUDA_CHECK(cudaMemcpyAsync(out.counts.p + at, done[j].counts.p, done[j].n * 8, cudaMemcpyDeviceToDevice, e->stream));
        at += done[j].n;
    }
    CUDA_CHECK(cudaStreamSynchronize(e->stream));
    return true;
}

// merge several (key, count) parts: concat, sort pairs, sum equal keys, keep sums >= c
static void reduce_fast_parts(mc2_engine* e, std::vector<FastPart>& parts, int key_bits, u64 c, FastPart& out) {
    u64 M = 0;
    for (auto& p : parts) M += p.n;
    out.n = 0;
    if (M == 0) return;
    if (c <= 1) {
        FastPart* only = nullptr;
        int nonempty = 0;
        for (auto& p : parts) if (p.n) { only = &p; nonempty++; }
        if (nonempty == 1 && only->sorted) { out = std::move(*only); return; }
        if (nonempty > 1 && concat_sorted_parts(e, parts, M, out)) return;
        if (nonempty > 1 && e->opt_row_merge == 1 && M > merge_rows_capacity(e) && merge_sorted_parts_by_range(e, parts, M, key_bits, out)) return;
    }
    DBuf<u64> k0(e, M), k1(e, M), v0(e, M), v1(e, M);
    u64 at = 0;
    for (auto& p : parts) {
        if (!p.n) continue;
        CUDA_CHECK(cudaMemcpyAsync(k0.p + at, p.keys.p, p.n * 8, cudaMemcpyDeviceToDevice, e->stream));
        CUDA_CHECK(cudaMemcpyAsync(v0.p + at, p.counts.p, p.n * 8, cudaMemcpyDeviceToDevice, e->stream));
        at += p.n;
    }
    // plain sums (the dict merge of already filtered tables) through the range partition: opt-in ("row_merge" = 1).
    // Measured on 16 filtered pieces, 350 M rows: 0.77-1.26 s against 0.63 s for the sort below.
    if (c <= 1 && e->opt_row_merge == 1 && merge_rows_range(e, k0.p, v0.p, M, key_bits, out)) return;
    const int r = radix_sort<u64, true>(e, k0.p, k1.p, v0.p, v1.p, M, 0, std::min(64, (key_bits + 7) & ~7));
    const u64* ks = r ? k1.p : k0.p;
    const u64* vs = r ? v1.p : v0.p;
    DBuf<u64> start, count;
    KeyEq acc{ks};
    const u64 ns = seg_reduce(e, acc, M, vs, c, start, count);
    out.n = ns;
    out.sorted = true;
    out.keys.alloc(e, ns);
    out.counts = std::move(count);
    if (ns) LAUNCH(e, gather_u64_kernel, (unsigned)div_up(ns, 256), 256, 0, ks, (const u64*)start.p, ns, out.keys.p);
}

// order m windows (k bytes each, at src + pos[i]) and reduce; weights == nullptr means unit weights
static void wide_reduce(mc2_engine* e, const u8* src, const u64* pos, const u64* weights, u64 m, int k, u64 c, WidePart& out) {
    out.n = 0;
    if (m == 0) return;
    if (m >= (1ull << 32)) throw Mc2Error(MC2_ERR_LIMIT, "wide path: more than 2^32-1 windows in one batch");
    DBuf<u32> i0(e, m), i1(e, m);
    DBuf<u64> k0(e, m), k1(e, m);
    LAUNCH(e, iota_u32_kernel, (unsigned)div_up(m, 256), 256, 0, i0.p, m);
    u32* idx[2] = {i0.p, i1.p};
    int cur = 0;
    const int L = (k + 7) / 8;
    for (int limb = L - 1; limb >= 0; --limb) {
        LAUNCH(e, wide_gather_limb_kernel, (unsigned)div_up(m, 256), 256, 0, src, pos, (const u32*)idx[cur], m, k, limb, k0.p);
        const int nb = std::min(8, k - 8 * limb);
        const int r = radix_sort<u32, true>(e, k0.p, k1.p, idx[cur], idx[1 - cur], m, 8 * (8 - nb), 64);
        if (r) cur ^= 1;     // an odd number of passes leaves the payload in the other buffer
    }
    const u32* sidx = idx[cur];
    WindowEq acc{src, pos, sidx, k};
    DBuf<u64> start, count;
    u64 ns;
    if (!weights) {
        ns = rle_threshold(e, acc, m, c, start, count);
    } else {
        DBuf<u64> ws(e, m);
        LAUNCH(e, gather_u64_by_u32_kernel, (unsigned)div_up(m, 256), 256, 0, weights, sidx, m, ws.p);
        ns = seg_reduce(e, acc, m, (const u64*)ws.p, c, start, count);
    }
    out.n = ns;
    out.rows.alloc(e, ns * (u64)k);
    out.counts = std::move(count);
    if (ns)
        LAUNCH(e, wide_gather_rows_kernel, (unsigned)div_up(ns * (u64)k, 256), 256, 0, src, pos, sidx, (const u64*)start.p, ns, k,
               out.rows.p);
}

static void reduce_wide_parts(mc2_engine* e, std::vector<WidePart>& parts, int k, u64 c, WidePart& out) {
    u64 M = 0;
    for (auto& p : parts) M += p.n;
    out.n = 0;
    if (M == 0) return;
    if (parts.size() == 1 && c <= 1 && parts[0].sorted) { out = std::move(parts[0]); return; }
    DBuf<u8> rows(e, M * (u64)k);
    DBuf<u64> w(e, M), pos(e, M);
    u64 at = 0;
    for (auto& p : parts) {
        if (!p.n) continue;
        CUDA_CHECK(cudaMemcpyAsync(rows.p + at * k, p.rows.p, p.n * (u64)k, cudaMemcpyDeviceToDevice, e->stream));
        CUDA_CHECK(cudaMemcpyAsync(w.p + at, p.counts.p, p.n * 8, cudaMemcpyDeviceToDevice, e->stream));
        at += p.n;
    }
    LAUNCH(e, wide_row_positions_kernel, (unsigned)div_up(M, 256), 256, 0, pos.p, M, k);
    wide_reduce(e, rows.p, pos.p, w.p, M, k, c, out);
}
