// mercat2_b200 engine: host orchestration + the C ABI of include/mercat2_b200.h.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a (see mercat2_b200/build.py).
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include <chrono>
#include "mercat2_b200.h"
#include "common.cuh"
#include "parse.cuh"
#include "extract.cuh"
#include "radix.cuh"
#include "wide.cuh"
#include "chunker.cuh"
#include "hashcount.cuh"
#include "fastnt.cuh"
#include "metrics.cuh"
#include "tsv.cuh"
#include "merge.cuh"

static thread_local std::string g_err;

// =====================================================================================================
// engine
// =====================================================================================================
struct mc2_engine {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t copy_stream = nullptr;
    int num_sms = 148;
    // options
    u64 opt_dense_max_bins = 1ull << 24;
    u64 opt_smem_max_bins = 32768;
    u64 opt_batch_symbols = 1ull << 28;
    int opt_force_path = 0;
    int opt_force_enc = -1;
    int opt_fast_nt = 1;                   // use the SWAR/packed nucleotide lane when the text is simple
    int opt_big_chunks = 1;                // chunks beyond one hash batch: level-0 key partition in HBM (0 = sort fallback)
    int opt_prefetch_pass = 1;             // run the next chunk's count pass behind the current chunk (one sync fewer per chunk)
    int opt_parse_single = 0;              // packed lane: 1 = one pass over the text with chained look-back (measured slower: 0.39 vs 0.28 ms per 100 MiB)
    u64 opt_file_piece = 32ull << 20;      // bytes per piece of the streaming file reader
    u64 opt_span_bytes = 1ull << 30;       // the packed lane parses a chunk in spans of about this many bytes
    int opt_count_variant = 2;             // min_count >= 2: 2 = bitmap pre-filter (faster as measured), 3 = 16-bit counter pre-filter
    int opt_scatter_variant = 0;           // bit0: stage destination indices, bit1: max shared-memory carveout
    int opt_sparse_algo = 0;               // 0 auto (hash tables when min_count >= 2), 1 radix sort, 2 hash tables
    u64 opt_hash_bucket_keys = 3500;       // target keys per shared-memory table
    // stats
    u64 launches = 0, h2d_bytes = 0, d2h_bytes = 0, chunks = 0, ovf_buckets = 0;
    double device_us = 0;
    // pinned scratch
    void* pin_small = nullptr;                 // 4 KiB for scalar readbacks
    const void* ride_dev = nullptr;            // a second small readback that rides on the next read_scalar's sync
    size_t ride_len = 0;
    bool ride_done = false;
    u8* pin_stage[2] = {nullptr, nullptr};     // H2D staging for pageable sources
    u8* file_pin[4] = {nullptr, nullptr, nullptr, nullptr};   // pinned pool of the streaming file reader (lazy)
    cudaEvent_t stage_ev[2] = {nullptr, nullptr};
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    static constexpr u64 STAGE_BYTES = 32ull << 20;
    // optional per-kernel timing (option "profile"): CUDA events around every launch on `stream`
    int profile = 0;
    struct ProfRec { const char* name; cudaEvent_t a, b; };
    std::vector<ProfRec> prof_pending;
    std::vector<cudaEvent_t> ev_pool;
    std::map<std::string, std::pair<u64, double>> prof_total;      // name -> (launches, microseconds)
    cudaEvent_t get_event() {
        if (!ev_pool.empty()) { cudaEvent_t ev = ev_pool.back(); ev_pool.pop_back(); return ev; }
        cudaEvent_t ev;
        CUDA_CHECK(cudaEventCreate(&ev));
        return ev;
    }
    void resolve_profile() {
        if (prof_pending.empty()) return;
        CUDA_CHECK(cudaStreamSynchronize(stream));
        for (auto& r : prof_pending) {
            float ms = 0;
            if (cudaEventElapsedTime(&ms, r.a, r.b) == cudaSuccess) {
                auto& t = prof_total[r.name];
                t.first++;
                t.second += ms * 1000.0;
            }
            ev_pool.push_back(r.a);
            ev_pool.push_back(r.b);
        }
        prof_pending.clear();
    }
};

#define LAUNCHN(e, name, kern, grid, block, smem, ...)                            \
    do {                                                                          \
        cudaEvent_t _pa = nullptr, _pb = nullptr;                                 \
        if ((e)->profile) {                                                       \
            _pa = (e)->get_event();                                               \
            _pb = (e)->get_event();                                               \
            cudaEventRecord(_pa, (e)->stream);                                    \
        }                                                                         \
        kern<<<(grid), (block), (smem), (e)->stream>>>(__VA_ARGS__);              \
        if ((e)->profile) {                                                       \
            cudaEventRecord(_pb, (e)->stream);                                    \
            (e)->prof_pending.push_back({name, _pa, _pb});                        \
            if ((e)->prof_pending.size() >= 4096) (e)->resolve_profile();         \
        }                                                                         \
        (e)->launches++;                                                          \
        CUDA_CHECK(cudaGetLastError());                                           \
    } while (0)
#define LAUNCH(e, kern, grid, block, smem, ...) LAUNCHN(e, #kern, kern, grid, block, smem, __VA_ARGS__)

template <typename T>
struct DBuf {
    mc2_engine* e = nullptr;
    T* p = nullptr;
    u64 n = 0;
    DBuf() {}
    DBuf(mc2_engine* e_, u64 n_) { alloc(e_, n_); }
    DBuf(const DBuf&) = delete;
    DBuf& operator=(const DBuf&) = delete;
    DBuf(DBuf&& o) noexcept : e(o.e), p(o.p), n(o.n) { o.p = nullptr; o.n = 0; }
    DBuf& operator=(DBuf&& o) noexcept {
        if (this != &o) { release(); e = o.e; p = o.p; n = o.n; o.p = nullptr; o.n = 0; }
        return *this;
    }
    void alloc(mc2_engine* e_, u64 n_) {
        release();
        e = e_;
        n = n_;
        u64 bytes = n * sizeof(T);
        if (bytes < 256) bytes = 256;
        CUDA_CHECK(cudaMallocAsync((void**)&p, bytes, e->stream));
    }
    void zero() { if (p) CUDA_CHECK(cudaMemsetAsync(p, 0, std::max<u64>(n * sizeof(T), 1), e->stream)); }
    void release() {
        if (p) cudaFreeAsync(p, e->stream);
        p = nullptr;
        n = 0;
    }
    ~DBuf() { release(); }
};

template <typename T>
static T read_scalar(mc2_engine* e, const T* dev) {
    CUDA_CHECK(cudaMemcpyAsync(e->pin_small, dev, sizeof(T), cudaMemcpyDeviceToHost, e->stream));
    if (e->ride_dev) {
        CUDA_CHECK(cudaMemcpyAsync((u8*)e->pin_small + 2048, e->ride_dev, e->ride_len, cudaMemcpyDeviceToHost, e->stream));
        e->ride_dev = nullptr;
        e->ride_done = true;
    }
    CUDA_CHECK(cudaStreamSynchronize(e->stream));
    e->d2h_bytes += sizeof(T);
    T v;
    memcpy(&v, e->pin_small, sizeof(T));
    return v;
}

// MC2_DEBUG_PHASES=1: wall time of coarse phases (synchronises the stream at every mark)
struct PhaseTimer {
    mc2_engine* e;
    bool on;
    std::chrono::steady_clock::time_point t0;
    explicit PhaseTimer(mc2_engine* e_) : e(e_), on(getenv("MC2_DEBUG_PHASES") != nullptr) { if (on) { cudaStreamSynchronize(e->stream); t0 = std::chrono::steady_clock::now(); } }
    void mark(const char* what) {
        if (!on) return;
        cudaStreamSynchronize(e->stream);
        const auto t1 = std::chrono::steady_clock::now();
        fprintf(stderr, "[phase] %-28s %9.3f ms\n", what, std::chrono::duration<double, std::milli>(t1 - t0).count());
        t0 = t1;
    }
};

template <typename T>
static void d2h(mc2_engine* e, T* host, const T* dev, u64 n) {
    if (!n) return;
    CUDA_CHECK(cudaMemcpyAsync(host, dev, n * sizeof(T), cudaMemcpyDeviceToHost, e->stream));
    CUDA_CHECK(cudaStreamSynchronize(e->stream));
    e->d2h_bytes += n * sizeof(T);
}

// ---- scans ---------------------------------------------------------------------------------------
template <typename T, typename O>
static void dev_exclusive_scan(mc2_engine* e, const T* in, O* out, u64 n, ull* total_dev) {
    if (n == 0) {
        if (total_dev) CUDA_CHECK(cudaMemsetAsync(total_dev, 0, sizeof(ull), e->stream));
        return;
    }
    const u64 ntiles = div_up(n, SCAN_TILE);
    DBuf<u64> sums(e, ntiles);
    LAUNCH(e, scan_reduce_kernel<T>, (unsigned)ntiles, SCAN_THREADS, 0, in, n, sums.p);
    LAUNCH(e, scan_sums_kernel, 1, SCAN1_THREADS, 0, sums.p, (u32)ntiles, total_dev);
    auto down = scan_down_kernel<T, O>;
    LAUNCHN(e, "scan_down_kernel", down, (unsigned)ntiles, SCAN_THREADS, 0, in, out, n, sums.p);
}

// ---- radix sort ----------------------------------------------------------------------------------
// Sorts n keys (and payloads) on bits [begin_bit, end_bit); returns which of the two buffers holds the
// result (0 -> k0/v0, 1 -> k1/v1).
template <typename V, bool HAS_V>
static int radix_sort(mc2_engine* e, u64* k0, u64* k1, V* v0, V* v1, u64 n, int begin_bit, int end_bit) {
    if (n >= (1ull << 32)) throw Mc2Error(MC2_ERR_LIMIT, "radix_sort: more than 2^32-1 keys in one batch");
    if (n <= 1 || end_bit <= begin_bit) return 0;
    const u32 ntiles = (u32)div_up(n, RS_TILE);
    DBuf<u32> hist(e, 256ull * ntiles);
    u64* kk[2] = {k0, k1};
    V* vv[2] = {v0, v1};
    int cur = 0;
    for (int shift = begin_bit; shift < end_bit; shift += 8) {
        LAUNCH(e, rs_hist_kernel, ntiles, RS_THREADS, 0, kk[cur], (u32)n, shift, hist.p, ntiles);
        dev_exclusive_scan<u32, u32>(e, hist.p, hist.p, 256ull * ntiles, nullptr);
        auto sc = rs_scatter_kernel<V, HAS_V>;
        LAUNCHN(e, "rs_scatter_kernel", sc, ntiles, RS_THREADS, 0, kk[cur], kk[1 - cur], vv[cur], vv[1 - cur], (u32)n, shift, hist.p, ntiles);
        cur ^= 1;
    }
    return cur;
}

// ---- compaction helper: tile counts -> offsets + total (synchronises) ---------------------------
static u64 offsets_from_counts(mc2_engine* e, const u32* tile_cnt, u64* tile_off, u64 ntiles) {
    DBuf<ull> total(e, 1);
    dev_exclusive_scan<u32, u64>(e, tile_cnt, tile_off, ntiles, total.p);
    return (u64)read_scalar<ull>(e, total.p);
}

// =====================================================================================================
// tables and sample accumulators
// =====================================================================================================
struct FastPart {          // unique (key, count) rows in the sample's fast encoding
    DBuf<u64> keys, counts;
    u64 n = 0;
    bool sorted = true;    // rows ordered by key (the hash path emits unordered rows)
};
struct WidePart {          // unique k-byte rows
    DBuf<u8> rows;
    DBuf<u64> counts;
    u64 n = 0;
    bool sorted = true;
};

enum : int { PATH_UNSET = 0, PATH_DENSE = 1, PATH_SPARSE = 2, PATH_WIDE = 3 };
enum : int { KEY_CODE = 0, KEY_DENSE_AA = 1 };   // how FastPart keys decode

struct Plan {
    int enc = -1;
    int path = PATH_UNSET;
    u32 bins = 0;
    bool smem = false;
    u32 nrep = 1;
};

// One span (< 4 GiB, starting at a line start) of simple FASTA text on its way to packed symbols.
struct FnSpan {
    const u8* text = nullptr;
    u64 len = 0, ntiles = 0, nsym = 0;
    DBuf<u8> tstate;
    DBuf<u32> tcnt;
    DBuf<u64> toff;
    DBuf<u32> codes, bad;
};

// The count pass of the NEXT chunk, enqueued ahead of the current chunk's final readback so that one host
// synchronisation serves both (its statistics land in the pinned scratch at offset 3072).
struct PrePass {
    bool valid = false;
    const u8* text = nullptr;
    u64 len = 0;
    FnSpan sp;
    DBuf<struct FnStats> st;
};

struct mc2_sample {
    mc2_engine* e = nullptr;
    int k = 0;
    u64 c = 1;
    Plan plan;
    DBuf<u64> dense_sample;
    DBuf<u32> dense_chunk;
    DBuf<u64> dense_chunk64;
    std::vector<FastPart> fast;
    std::vector<WidePart> wide;
    u64 n_chunks = 0;
    double bucket_scale = 1.0;             // shrinks when many hash buckets overflow (heavily duplicated keys)
    PrePass pre;
    const u8* next_text = nullptr;         // the chunk that follows the one being counted (resident text), for PrePass
    u64 next_len = 0;
};

struct mc2_table {
    mc2_engine* e = nullptr;
    int k = 0;
    int enc = ENC_NT2;
    int key_kind = KEY_CODE;
    FastPart fast;
    WidePart wide;
    // formatted TSV body kept between the size query and the copy-out of mc2_table_tsv
    DBuf<u8> tsv_body;
    u64 tsv_bytes = 0;
    bool tsv_ready = false;
    // host side (filled by ensure_host)
    bool on_host = false;
    std::vector<char> kmers;
    std::vector<u64> counts;
    u64 total = 0;
};

// =====================================================================================================
// K1 parse
// =====================================================================================================
struct Parsed {
    DBuf<u8> sym;
    u64 nsym = 0;
    ParseStats stats;
};

static void parse_text(mc2_engine* e, const u8* dtext, u64 len, int toupper, Parsed& out) {
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
    }
    DBuf<u64> k0(e, M), k1(e, M), v0(e, M), v1(e, M);
    u64 at = 0;
    for (auto& p : parts) {
        if (!p.n) continue;
        CUDA_CHECK(cudaMemcpyAsync(k0.p + at, p.keys.p, p.n * 8, cudaMemcpyDeviceToDevice, e->stream));
        CUDA_CHECK(cudaMemcpyAsync(v0.p + at, p.counts.p, p.n * 8, cudaMemcpyDeviceToDevice, e->stream));
        at += p.n;
    }
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

// =====================================================================================================
// per-chunk counting
// =====================================================================================================
template <int ENC>
static void dense_batch(mc2_engine* e, const Plan& plan, SymView v, u64 s0, u64 s1, int k, u32* table) {
    const u64 ntiles = div_up(s1 - s0, EX_TILE);
    if (plan.smem) {
        auto kern = dense_smem_kernel<ENC>;
        const size_t smem = (size_t)plan.bins * plan.nrep * 4;
        static thread_local bool attr_set[3] = {false, false, false};
        if (!attr_set[ENC]) {
            CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
            attr_set[ENC] = true;
        }
        int per_sm = 1;
        CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, EX_THREADS, smem));
        if (per_sm < 1) per_sm = 1;
        const u64 grid = std::min<u64>(ntiles, (u64)e->num_sms * per_sm);
        LAUNCHN(e, "dense_smem_kernel", kern, (unsigned)grid, EX_THREADS, smem, v, s0, s1, k, plan.bins, plan.nrep, table);
    } else {
        auto kern = dense_global_kernel<ENC>;
        LAUNCHN(e, "dense_global_kernel", kern, (unsigned)ntiles, EX_THREADS, 0, v, s0, s1, k, table);
    }
}

// sort + RLE of an already materialised key range (fallback for overflowed hash buckets)
static void count_key_range_sorted(mc2_engine* e, mc2_sample* s, const u64* keys, u64 m, int kb) {
    if (!m) return;
    DBuf<u64> k0(e, m), k1(e, m);
    CUDA_CHECK(cudaMemcpyAsync(k0.p, keys, m * 8, cudaMemcpyDeviceToDevice, e->stream));
    const int r = radix_sort<NoVal, false>(e, k0.p, k1.p, (NoVal*)nullptr, (NoVal*)nullptr, m, 0, std::min(64, (kb + 7) & ~7));
    const u64* ks = r ? k1.p : k0.p;
    KeyEq acc{ks};
    DBuf<u64> start, count;
    const u64 ns = rle_threshold(e, acc, m, s->c, start, count);
    if (!ns) return;
    FastPart part;
    part.n = ns;
    part.keys.alloc(e, ns);
    part.counts = std::move(count);
    LAUNCH(e, gather_u64_kernel, (unsigned)div_up(ns, 256), 256, 0, ks, (const u64*)start.p, ns, part.keys.p);
    s->fast.push_back(std::move(part));
}

// hash-partition + shared-memory tables (hashcount.cuh); the chunk must fit one batch.  Keys come either from
// the byte symbol stream `v` (encoding ENC) or, when `pv` is given, from the packed nucleotide stream.
static void prefetch_next_count_pass(mc2_engine* e, mc2_sample* s);

struct KeySpan {               // keys already extracted (one level-0 group of a very large chunk)
    const u64* keys;
    u64 n;
    bool stream_order;         // keys come from the packed lane (first symbol in the low bits)
};

template <int ENC>
static void sparse_chunk_hash(mc2_engine* e, mc2_sample* s, SymView v, const PackedView* pv = nullptr, const KeySpan* ks = nullptr) {
    const int k = s->k;
    const int kb = k * EncTraits<ENC>::BITS;
    const u64 cap = ks ? ks->n : pv ? pv->n : v.n;              // upper bound on the number of windows
    const u64 mult = ks ? HC_MULT2 : HC_MULT1;
    const bool stream_order = pv || (ks && ks->stream_order);
    // (`cap` of a symbol stream counts ~25 % more positions than windows; a key array is exact, so aim lower there to
    // keep the same head room below the 4096 keys a bucket may hold)
    const u64 bucket_keys = std::max<u64>(1, (u64)((double)e->opt_hash_bucket_keys * s->bucket_scale * (ks ? 0.8 : 1.0)));
    const u32 nb1 = (u32)std::min<u64>(HC_MAX_NB1, std::max<u64>(1, div_up(cap, bucket_keys * HC_NB2)));
    const u32 nb = nb1 * HC_NB2;
    DBuf<u32> ghist(e, nb), sub_base(e, nb + 1), cur1(e, nb1), cur2(e, nb), tile_pref(e, nb1 + 1), ovf_list(e, nb);
    struct Tail { ull total, out_n; u32 ovf_n, pad; };
    DBuf<Tail> tail(e, 1);
    ghist.zero();
    tail.zero();
    const size_t hist_smem = (size_t)nb * 4;
    if (ks) {
        static thread_local bool attr_set = false;
        if (!attr_set) {
            CUDA_CHECK(cudaFuncSetAttribute(hk_hist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 208 * 1024));
            CUDA_CHECK(cudaFuncSetAttribute(hk_scatter1_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)HC_SCATTER_SMEM));
            attr_set = true;
        }
        int per_sm = 1;
        CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, hk_hist_kernel, HK_HIST_THREADS, hist_smem));
        const u64 grid = std::min<u64>(div_up(cap, HK_HIST_THREADS * 8), (u64)e->num_sms * std::max(per_sm, 1));
        LAUNCH(e, hk_hist_kernel, (unsigned)std::max<u64>(grid, 1), HK_HIST_THREADS, hist_smem, ks->keys, ks->n, nb, mult, ghist.p);
    } else if (pv) {
        static thread_local bool attr_set = false;
        if (!attr_set) {
            CUDA_CHECK(cudaFuncSetAttribute(fn_hist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 208 * 1024));
            attr_set = true;
        }
        int per_sm = 1;
        CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn_hist_kernel, FN_HIST_THREADS, hist_smem));
        const u64 nwords = div_up(cap, 16);
        const u64 grid = std::min<u64>(div_up(nwords, FN_HIST_THREADS), (u64)e->num_sms * std::max(per_sm, 1));
        LAUNCH(e, fn_hist_kernel, (unsigned)grid, FN_HIST_THREADS, hist_smem, *pv, k, nb, ghist.p);
    } else {
        auto kern = hc_hist_kernel<ENC>;
        static thread_local bool attr_set[3] = {false, false, false};
        if (!attr_set[ENC]) {
            CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 208 * 1024));
            attr_set[ENC] = true;
        }
        int per_sm = 1;
        CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, EX_THREADS, hist_smem));
        const u64 grid = std::min<u64>(div_up(cap, EX_TILE), (u64)e->num_sms * std::max(per_sm, 1));
        LAUNCHN(e, "hc_hist_kernel", kern, (unsigned)grid, EX_THREADS, hist_smem, v, (u64)0, v.n, k, nb, ghist.p);
    }
    {
        static thread_local bool attr_set = false;
        if (!attr_set) {
            CUDA_CHECK(cudaFuncSetAttribute(hc_scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((size_t)HC_MAX_NB1 * HC_NB2 * 4)));
            attr_set = true;
        }
    }
    LAUNCH(e, hc_scan_kernel, 1, 1024, (size_t)nb * 4, (const u32*)ghist.p, nb, nb1, (u32)HC_NB2, sub_base.p, cur1.p, cur2.p, tile_pref.p, &tail.p->total);
    DBuf<u64> keys1(e, cap), keys2(e, cap);
    const bool dbg = getenv("MC2_DEBUG_HASH") != nullptr;
    if (dbg) {
        CUDA_CHECK(cudaMemsetAsync(keys1.p, 0xEE, cap * 8, e->stream));
        CUDA_CHECK(cudaMemsetAsync(keys2.p, 0xEE, cap * 8, e->stream));
    }
    const bool use_dst = (e->opt_scatter_variant & 1) && !ks;
    const size_t sc_smem = use_dst ? HC_SCATTER_SMEM : HC_SCATTER_SMEM16;
    {
        static thread_local int attr_variant = -1;
        if (attr_variant != e->opt_scatter_variant) {
            const int carve = (e->opt_scatter_variant & 2) ? 100 : -1;
            CUDA_CHECK(cudaFuncSetAttribute(fn_scatter1_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)HC_SCATTER_SMEM));
            CUDA_CHECK(cudaFuncSetAttribute(fn_scatter1_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)HC_SCATTER_SMEM));
            CUDA_CHECK(cudaFuncSetAttribute(hc_scatter2_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)HC_SCATTER_SMEM));
            CUDA_CHECK(cudaFuncSetAttribute(hc_scatter2_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)HC_SCATTER_SMEM));
            CUDA_CHECK(cudaFuncSetAttribute(hc_scatter2_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)HC_SCATTER_SMEM));
            CUDA_CHECK(cudaFuncSetAttribute(hc_scatter2_kernel<false, true>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
            const int carve1 = (e->opt_scatter_variant & 4) ? 85 : carve;
            CUDA_CHECK(cudaFuncSetAttribute(fn_scatter1_kernel<true>, cudaFuncAttributePreferredSharedMemoryCarveout, carve1));
            CUDA_CHECK(cudaFuncSetAttribute(fn_scatter1_kernel<false>, cudaFuncAttributePreferredSharedMemoryCarveout, carve1));
            CUDA_CHECK(cudaFuncSetAttribute(hc_scatter2_kernel<true, false>, cudaFuncAttributePreferredSharedMemoryCarveout, carve));
            CUDA_CHECK(cudaFuncSetAttribute(hc_scatter2_kernel<false, false>, cudaFuncAttributePreferredSharedMemoryCarveout, carve));
            attr_variant = e->opt_scatter_variant;
        }
    }
    if (ks) {
        LAUNCH(e, hk_scatter1_kernel, (unsigned)div_up(cap, HC_TILE), EX_THREADS, HC_SCATTER_SMEM16, ks->keys, ks->n, nb, nb1, mult, cur1.p, keys1.p,
               (const u64*)nullptr);
    } else if (pv) {
        if (use_dst) LAUNCH(e, fn_scatter1_kernel<true>, (unsigned)div_up(div_up(cap, 16), EX_THREADS), EX_THREADS, sc_smem, *pv, k, nb, nb1, cur1.p, keys1.p, (const u64*)nullptr);
        else LAUNCH(e, fn_scatter1_kernel<false>, (unsigned)div_up(div_up(cap, 16), EX_THREADS), EX_THREADS, sc_smem, *pv, k, nb, nb1, cur1.p, keys1.p, (const u64*)nullptr);
    } else {
        auto kern = hc_scatter1_kernel<ENC>;
        static thread_local bool attr_set[3] = {false, false, false};
        if (!attr_set[ENC]) {
            CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)HC_SCATTER_SMEM));
            attr_set[ENC] = true;
        }
        LAUNCHN(e, "hc_scatter1_kernel", kern, (unsigned)div_up(cap, EX_TILE), EX_THREADS, HC_SCATTER_SMEM, v, (u64)0, v.n, k, nb, nb1, cur1.p, keys1.p);
    }
    if (use_dst)
        LAUNCHN(e, "hc_scatter2_kernel", (hc_scatter2_kernel<true, false>), (unsigned)(div_up(cap, HC_TILE) + nb1), EX_THREADS, sc_smem, (const u64*)keys1.p, (const u32*)sub_base.p,
               (const u32*)tile_pref.p, nb, nb1, (u32)HC_NB2, cur2.p, keys2.p, mult);
    else if (e->opt_scatter_variant & 8)
        LAUNCHN(e, "hc_scatter2_kernel", (hc_scatter2_kernel<false, true>), (unsigned)(div_up(cap, HC_TILE) + nb1), EX_THREADS, sc_smem, (const u64*)keys1.p, (const u32*)sub_base.p,
               (const u32*)tile_pref.p, nb, nb1, (u32)HC_NB2, cur2.p, keys2.p, mult);
    else
        LAUNCHN(e, "hc_scatter2_kernel", (hc_scatter2_kernel<false, false>), (unsigned)(div_up(cap, HC_TILE) + nb1), EX_THREADS, sc_smem, (const u64*)keys1.p, (const u32*)sub_base.p,
               (const u32*)tile_pref.p, nb, nb1, (u32)HC_NB2, cur2.p, keys2.p, mult);
    if (dbg) {
        DBuf<ull> badc(e, 2);
        badc.zero();
        const u64 total = (u64)read_scalar<ull>(e, &tail.p->total);
        if (total) {
            LAUNCH(e, hc_verify_kernel, (unsigned)div_up(total, 256), 256, 0, (const u64*)keys1.p, (const u32*)sub_base.p, nb, (u32)HC_NB2, (u32)total, badc.p, mult);
            LAUNCH(e, hc_verify_kernel, (unsigned)div_up(total, 256), 256, 0, (const u64*)keys2.p, (const u32*)sub_base.p, nb, 1u, (u32)total, badc.p + 1, mult);
        }
        const ull b1 = read_scalar<ull>(e, badc.p), b2 = read_scalar<ull>(e, badc.p + 1);
        fprintf(stderr, "[hash] cap=%llu total=%llu nb1=%u nb=%u packed=%d misplaced level1=%llu level2=%llu\n", (ull)cap, (ull)total, nb1, nb,
                pv ? 1 : 0, b1, b2);
        DBuf<ull> cs(e, 9);
        cs.zero();
        if (pv) LAUNCH(e, fn_checksum_kernel, 256, 256, 0, *pv, k, cs.p);
        LAUNCH(e, key_checksum_kernel, 256, 256, 0, (const u64*)keys1.p, total, cs.p + 3);
        LAUNCH(e, key_checksum_kernel, 256, 256, 0, (const u64*)keys2.p, total, cs.p + 6);
        ull h[9];
        d2h(e, h, cs.p, 9);
        fprintf(stderr, "[hash] checksum stream (%llx %llx %llu) keys1 (%llx %llx %llu) keys2 (%llx %llx %llu)%s\n", h[0], h[1], h[2], h[3],
                h[4], h[5], h[6], h[7], h[8], (pv && (h[0] != h[6] || h[1] != h[7] || h[0] != h[3])) ? "  MISMATCH" : "");
    }
    const u64 out_cap = cap / s->c + 2;
    FastPart part;
    part.keys.alloc(e, out_cap);
    part.counts.alloc(e, out_cap);
    unsigned cgrid = (unsigned)std::min<u64>(nb, (u64)e->num_sms);
    if (s->c >= 2 && (e->opt_count_variant & 15) == 3) {
        static thread_local bool attr_set = false;
        if (!attr_set) {
            CUDA_CHECK(cudaFuncSetAttribute(hc_count3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)HC3_SMEM));
            attr_set = true;
        }
        LAUNCH(e, hc_count3_kernel, cgrid, HC3_THREADS, HC3_SMEM, (const u64*)keys2.p, (const u32*)sub_base.p, nb, s->c,
               part.keys.p, part.counts.p, &tail.p->out_n, out_cap, ovf_list.p, &tail.p->ovf_n);
    } else if (s->c >= 2) {                                        // (count_variant bits >= 4: timing experiments)
        cgrid = (unsigned)std::min<u64>(nb, 2ull * e->num_sms);
        static thread_local bool attr_set = false;
        if (!attr_set) {
            CUDA_CHECK(cudaFuncSetAttribute(hc_count2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)HC2_SMEM));
            CUDA_CHECK(cudaFuncSetAttribute(hc_count2_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
            attr_set = true;
        }
        LAUNCH(e, hc_count2_kernel, cgrid, HC2_THREADS, HC2_SMEM, (const u64*)keys2.p, (const u32*)sub_base.p, nb, s->c,
               part.keys.p, part.counts.p, &tail.p->out_n, out_cap, ovf_list.p, &tail.p->ovf_n, (ull*)nullptr,
               (u32)(e->opt_count_variant >> 4));
        if (dbg) {                                                // stress: repeat on the same keys, results must not vary
            DBuf<ull> dc(e, 8 + (u64)cgrid * 64);
            DBuf<Tail> t2(e, 1);
            DBuf<u64> ok2(e, out_cap), oc2(e, out_cap);
            for (int rep = 0; rep < 40; ++rep) {
                dc.zero();
                t2.zero();
                LAUNCH(e, hc_count2_kernel, cgrid, HC2_THREADS, HC2_SMEM, (const u64*)keys2.p, (const u32*)sub_base.p, nb, s->c,
                       ok2.p, oc2.p, &t2.p->out_n, out_cap, ovf_list.p, &t2.p->ovf_n, dc.p, 0u);
                ull h[4];
                d2h(e, h, dc.p, 4);
                if (cgrid == nb) {                                 // one bucket per CTA: barrier timestamps are meaningful
                    std::vector<ull> ts((u64)cgrid * 64);
                    d2h(e, ts.data(), dc.p + 8, (u64)cgrid * 64);
                    int leaks1 = 0, leaks2 = 0;
                    for (unsigned cta = 0; cta < cgrid; ++cta) {
                        ull max_end1 = 0, min_beg2 = ~0ull, max_end2 = 0, min_beg3 = ~0ull;
                        for (int w = 0; w < 16; ++w) {
                            const ull* q = &ts[((u64)cta * 16 + w) * 4];
                            max_end1 = std::max(max_end1, q[0]); min_beg2 = std::min(min_beg2, q[1]);
                            max_end2 = std::max(max_end2, q[2]); min_beg3 = std::min(min_beg3, q[3]);
                        }
                        if (min_beg2 < max_end1) leaks1++;
                        if (min_beg3 < max_end2) leaks2++;
                    }
                    if (leaks1 || leaks2) fprintf(stderr, "[hash] BARRIER LEAK rep %d: %d CTAs passed the pass1|pass2 barrier early, %d the pass2|emit barrier\n", rep, leaks1, leaks2);
                }
                const Tail tt = read_scalar<Tail>(e, t2.p);
                static ull first_out = 0, first_hits = 0, first_claims = 0;
                if (rep == 0) { first_out = tt.out_n; first_hits = h[0]; first_claims = h[2]; }
                if (h[0] != h[1] || tt.out_n != first_out || h[0] != first_hits || h[2] != first_claims || h[3])
                    fprintf(stderr, "[hash] STRESS rep %d: pass2 hits %llu, counts read back %llu, claims %llu (first %llu), empty-key slots %llu, survivors %llu (first %llu)\n",
                            rep, h[0], h[1], h[2], first_claims, h[3], (ull)tt.out_n, first_out);
            }
        }
    } else {
        static thread_local bool attr_set = false;
        if (!attr_set) {
            CUDA_CHECK(cudaFuncSetAttribute(hc_count_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)HC_COUNT_SMEM));
            attr_set = true;
        }
        LAUNCH(e, hc_count_kernel, cgrid, HC_THREADS, HC_COUNT_SMEM, (const u64*)keys2.p, (const u32*)sub_base.p, nb, s->c,
               part.keys.p, part.counts.p, &tail.p->out_n, out_cap, ovf_list.p, &tail.p->ovf_n);
    }
    if (pv && !ks) prefetch_next_count_pass(e, s);               // rides on the synchronisation below
    const Tail t = read_scalar<Tail>(e, tail.p);
    if (t.out_n > out_cap) throw Mc2Error(MC2_ERR_LIMIT, "hash path: survivor buffer overflow (internal error)");
    if (ks && getenv("MC2_DEBUG_PHASES"))
        fprintf(stderr, "[phase]   group: %llu keys, %u buckets, %llu survivors, %u overflowed buckets\n", (ull)cap, nb, (ull)t.out_n, t.ovf_n);
    if (dbg) {                                                    // the sort path on the same keys must agree
        mc2_sample tmp;
        tmp.e = e; tmp.k = k; tmp.c = s->c;
        count_key_range_sorted(e, &tmp, keys2.p, t.total, 64);
        u64 ref_rows = 0;
        for (auto& p : tmp.fast) ref_rows += p.n;
        fprintf(stderr, "[hash] survivors: tables %llu (+%u overflowed buckets) sort %llu%s\n", (ull)t.out_n, t.ovf_n, (ull)ref_rows,
                (!t.ovf_n && ref_rows != t.out_n) ? "  MISMATCH" : "");
    }
    if (t.out_n) {
        if (stream_order) LAUNCH(e, fn_canon_kernel, (unsigned)div_up(t.out_n, 256), 256, 0, part.keys.p, (u64)t.out_n, k);
        part.n = t.out_n;
        part.sorted = false;
        s->fast.push_back(std::move(part));
    }
    if (t.ovf_n) {
        // Buckets the tables could not take (more keys than the prefetch registers hold, or too many distinct
        // repeats): their keys are gathered into ONE array and counted by a single sort + run-length pass (buckets
        // hold disjoint key sets).  Heavily duplicated data makes bucket sizes spread (sigma ~ sqrt(size * copies)),
        // so the overflow rate also steers the bucket size of the sample's next chunks.
        std::vector<u32> ovf(t.ovf_n), base(nb + 1);
        d2h(e, ovf.data(), ovf_list.p, t.ovf_n);
        d2h(e, base.data(), sub_base.p, nb + 1);
        std::vector<u64> src_off(t.ovf_n), dst_off(t.ovf_n);
        u64 m = 0;
        for (u32 i = 0; i < t.ovf_n; ++i) {
            src_off[i] = base[ovf[i]];
            dst_off[i] = m;
            m += base[ovf[i] + 1] - base[ovf[i]];
        }
        if (m) {
            DBuf<u64> so(e, t.ovf_n), dof(e, t.ovf_n), gathered(e, m);
            CUDA_CHECK(cudaMemcpyAsync(so.p, src_off.data(), t.ovf_n * 8ull, cudaMemcpyHostToDevice, e->stream));
            CUDA_CHECK(cudaMemcpyAsync(dof.p, dst_off.data(), t.ovf_n * 8ull, cudaMemcpyHostToDevice, e->stream));
            LAUNCH(e, gather_ranges_kernel, (unsigned)std::min<u64>(div_up(m, 256), 65535), 256, 0, (const u64*)keys2.p, (const u64*)so.p,
                   (const u64*)dof.p, t.ovf_n, m, gathered.p);
            if (stream_order) LAUNCH(e, fn_canon_kernel, (unsigned)div_up(m, 256), 256, 0, gathered.p, m, k);
            CUDA_CHECK(cudaStreamSynchronize(e->stream));          // (host vectors were the sources of async copies)
            count_key_range_sorted(e, s, gathered.p, m, kb);
        }
        e->ovf_buckets += t.ovf_n;
        if ((u64)t.ovf_n * 200 > nb && s->bucket_scale > 0.3) s->bucket_scale *= 0.8;      // > 0.5 % of the buckets overflowed
    }
}

template <int ENC>
static void sparse_chunk(mc2_engine* e, mc2_sample* s, SymView v) {
    {
        const u64 hash_max = (u64)HC_MAX_NB1 * HC_NB2 * e->opt_hash_bucket_keys;
        const bool want_hash = e->opt_sparse_algo == 2 || (e->opt_sparse_algo == 0 && s->c >= 2);
        if (want_hash && v.n <= std::min<u64>(hash_max, e->opt_batch_symbols) && v.n < (1ull << 32)) {
            sparse_chunk_hash<ENC>(e, s, v);
            return;
        }
    }
    const int k = s->k;
    const int kb = k * EncTraits<ENC>::BITS;
    const u64 batch = std::max<u64>(EX_TILE, e->opt_batch_symbols / EX_TILE * EX_TILE);
    const u64 nb = div_up(v.n, batch);
    std::vector<FastPart> partial;
    for (u64 b = 0; b < nb; ++b) {
        const u64 s0 = b * batch, s1 = std::min(v.n, s0 + batch);
        const u64 cap = s1 - s0;
        DBuf<u64> k0(e, cap), k1(e, cap);
        DBuf<ull> nk(e, 1);
        nk.zero();
        auto kern = extract_keys_kernel<ENC>;
        LAUNCHN(e, "extract_keys_kernel", kern, (unsigned)div_up(cap, EX_TILE), EX_THREADS, 0, v, s0, s1, k, k0.p, nk.p);
        const u64 m = (u64)read_scalar<ull>(e, nk.p);
        if (!m) continue;
        const int r = radix_sort<NoVal, false>(e, k0.p, k1.p, (NoVal*)nullptr, (NoVal*)nullptr, m, 0, std::min(64, (kb + 7) & ~7));
        const u64* ks = r ? k1.p : k0.p;
        KeyEq acc{ks};
        DBuf<u64> start, count;
        const u64 ns = rle_threshold(e, acc, m, nb == 1 ? s->c : 1, start, count);
        if (!ns) continue;
        FastPart part;
        part.n = ns;
        part.keys.alloc(e, ns);
        part.counts = std::move(count);
        LAUNCH(e, gather_u64_kernel, (unsigned)div_up(ns, 256), 256, 0, ks, (const u64*)start.p, ns, part.keys.p);
        (nb == 1 ? s->fast : partial).push_back(std::move(part));
    }
    if (nb > 1 && !partial.empty()) {
        FastPart merged;
        // force the merge path even for a single partial so that the chunk threshold is applied
        if (partial.size() == 1) partial.emplace_back();
        reduce_fast_parts(e, partial, kb, s->c, merged);
        if (merged.n) s->fast.push_back(std::move(merged));
    }
}

template <int ENC>
static void dense_chunk(mc2_engine* e, mc2_sample* s, SymView v) {
    const Plan& plan = s->plan;
    const u64 batch = std::max<u64>(EX_TILE, std::min<u64>(e->opt_batch_symbols, 1ull << 31) / EX_TILE * EX_TILE);
    const u64 nb = div_up(v.n, batch);
    const unsigned fgrid = (unsigned)div_up(plan.bins, 256);
    for (u64 b = 0; b < nb; ++b) {
        const u64 s0 = b * batch, s1 = std::min(v.n, s0 + batch);
        dense_batch<ENC>(e, plan, v, s0, s1, s->k, s->dense_chunk.p);
        if (nb > 1) {
            if (!s->dense_chunk64.p) { s->dense_chunk64.alloc(e, plan.bins); s->dense_chunk64.zero(); }
            LAUNCH(e, dense_fold_batch_kernel, fgrid, 256, 0, s->dense_chunk.p, s->dense_chunk64.p, plan.bins);
        }
    }
    if (nb > 1) LAUNCH(e, dense_fold64_kernel, fgrid, 256, 0, s->dense_chunk64.p, s->dense_sample.p, plan.bins, s->c);
    else LAUNCH(e, dense_fold_kernel, fgrid, 256, 0, s->dense_chunk.p, s->dense_sample.p, plan.bins, s->c);
}

// exception / all-window positions -> wide reduce -> append
template <int ENC, int MODE>
static void wide_chunk(mc2_engine* e, mc2_sample* s, SymView v, u64 cap_hint) {
    const int k = s->k;
    const u64 batch = std::max<u64>(EX_TILE, e->opt_batch_symbols / EX_TILE * EX_TILE);
    const u64 nb = div_up(v.n, batch);
    std::vector<WidePart> partial;
    for (u64 b = 0; b < nb; ++b) {
        const u64 s0 = b * batch, s1 = std::min(v.n, s0 + batch);
        u64 cap = std::min<u64>(s1 - s0, cap_hint);
        DBuf<ull> np(e, 1);
        u64 m = 0;
        DBuf<u64> pos;
        for (int attempt = 0; attempt < 2; ++attempt) {
            pos.alloc(e, cap);
            np.zero();
            auto kern = extract_positions_kernel<ENC, MODE>;
            LAUNCHN(e, "extract_positions_kernel", kern, (unsigned)div_up(s1 - s0, EX_TILE), EX_THREADS, 0, v, s0, s1, k, pos.p, cap, np.p);
            m = (u64)read_scalar<ull>(e, np.p);
            if (m <= cap) break;
            cap = m;                       // hint was too small: rerun with the exact size
        }
        if (!m) continue;
        WidePart part;
        wide_reduce(e, v.sym, pos.p, nullptr, m, k, nb == 1 ? s->c : 1, part);
        if (part.n) (nb == 1 ? s->wide : partial).push_back(std::move(part));
    }
    if (nb > 1 && !partial.empty()) {
        WidePart merged;
        if (partial.size() == 1) partial.emplace_back();
        reduce_wide_parts(e, partial, k, s->c, merged);
        if (merged.n) s->wide.push_back(std::move(merged));
    }
}

// raw symbol stream (calculateKmerCount: no FASTA parsing): copy into an aligned buffer + statistics
static void adopt_symbols(mc2_engine* e, const u8* dsym, u64 len, Parsed& out) {
    memset(&out.stats, 0, sizeof out.stats);
    out.nsym = len;
    if (!len) return;
    out.sym.alloc(e, (len + 64) & ~15ull);
    CUDA_CHECK(cudaMemcpyAsync(out.sym.p, dsym, len, cudaMemcpyDeviceToDevice, e->stream));
    DBuf<ParseStats> st(e, 1);
    st.zero();
    LAUNCH(e, symbol_stats_kernel, (unsigned)std::min<u64>(div_up(len, 256 * 16), 4096), 256, 0, (const u8*)out.sym.p, len, st.p);
    out.stats = read_scalar<ParseStats>(e, st.p);
    if (out.stats.n_bad)
        throw Mc2Error(MC2_ERR_NON_ASCII, "sequence contains non-ASCII characters; only 7-bit ASCII is supported");
}

// The nucleotide fast lane (fastnt.cuh).  Returns false when the chunk must go through the general parser
// (text not simple, sample is not nucleotide / not on the hash path); *need_exceptions is set when the chunk
// holds non-ACGT symbols whose windows still have to be counted by the wide path.
static std::vector<u64> chunk_bounds(mc2_engine* e, const u8* dtext, u64 n, u64 chunk_bytes);


// count pass: tile line states + symbols per tile (+ alphabet statistics on the first piece of a sample)
static void fn_count_pass_launch(mc2_engine* e, FnSpan& sp, bool with_stats, DBuf<FnStats>& st) {
    const u64 mis = (u64)(uintptr_t)sp.text & 15ull;
    sp.ntiles = div_up(mis + sp.len, FN_VTILE);
    sp.tstate.alloc(e, sp.ntiles);
    sp.tcnt.alloc(e, sp.ntiles);
    sp.toff.alloc(e, sp.ntiles);
    if (with_stats)
        LAUNCH(e, fn_parse_kernel<0>, (unsigned)sp.ntiles, FN_THREADS, 0, sp.text, sp.len, sp.tstate.p, sp.tcnt.p, (const u64*)nullptr,
               (u32*)nullptr, (u32*)nullptr, st.p);
    else
        LAUNCH(e, fn_parse_kernel<2>, (unsigned)sp.ntiles, FN_THREADS, 0, sp.text, sp.len, sp.tstate.p, sp.tcnt.p, (const u64*)nullptr,
               (u32*)nullptr, (u32*)nullptr, st.p);
    dev_exclusive_scan<u32, u64>(e, sp.tcnt.p, sp.toff.p, sp.ntiles, &st.p->n_sym);
}
static FnStats fn_count_pass(mc2_engine* e, FnSpan& sp, bool with_stats, DBuf<FnStats>& st) {
    fn_count_pass_launch(e, sp, with_stats, st);
    const FnStats fs = read_scalar<FnStats>(e, st.p);
    sp.nsym = fs.n_sym;
    if (getenv("MC2_DEBUG_FAST"))
        fprintf(stderr, "[fast_nt] len=%llu n_sym=%llu kept=%llu non_acgt=%llu complex=%llu\n", (ull)sp.len, fs.n_sym,
                fs.packed & 0xFFFFFFFFull, fs.packed >> 32, fs.complex);
    return fs;
}

static void prefetch_next_count_pass(mc2_engine* e, mc2_sample* s) {
    if (!s->next_len || s->pre.valid || !e->opt_prefetch_pass) return;
    const u64 hash_max = std::min<u64>((u64)HC_MAX_NB1 * HC_NB2 * e->opt_hash_bucket_keys, e->opt_batch_symbols);
    const bool ok = e->opt_fast_nt && !e->opt_parse_single && s->k <= 32 && s->c >= 2 && e->opt_sparse_algo != 1 &&
                    s->plan.enc == ENC_NT2 && s->plan.path == PATH_SPARSE && e->opt_force_enc <= 0 && e->opt_force_path != PATH_WIDE &&
                    s->next_len <= std::min<u64>(e->opt_span_bytes, hash_max);
    if (!ok) { s->next_len = 0; return; }
    PrePass& p = s->pre;
    p.text = s->next_text;
    p.len = s->next_len;
    p.sp = FnSpan();
    p.sp.text = p.text;
    p.sp.len = p.len;
    p.st.alloc(e, 1);
    p.st.zero();
    fn_count_pass_launch(e, p.sp, false, p.st);
    CUDA_CHECK(cudaMemcpyAsync((u8*)e->pin_small + 3072, p.st.p, sizeof(FnStats), cudaMemcpyDeviceToHost, e->stream));
    p.valid = true;
    s->next_len = 0;
}

// count + write in one pass over the text (chained look-back for the symbol offsets, see fn_parse_single_kernel)
static FnStats fn_single_pass(mc2_engine* e, FnSpan& sp, bool with_stats, DBuf<FnStats>& st) {
    const u64 mis = (u64)(uintptr_t)sp.text & 15ull;
    sp.ntiles = div_up(mis + sp.len, FN_TILE);
    const u64 cap_sym = sp.len + 1;                              // every symbol comes from its own text byte
    sp.codes.alloc(e, div_up(cap_sym, 16) + 4);
    sp.bad.alloc(e, div_up(cap_sym, 32) + 4);
    sp.codes.zero();
    sp.bad.zero();
    DBuf<ull> desc(e, sp.ntiles);
    DBuf<u32> ticket(e, 1);
    desc.zero();
    ticket.zero();
    if (with_stats)
        LAUNCHN(e, "fn_parse_single_kernel<stats>", fn_parse_single_kernel<true>, (unsigned)sp.ntiles, FN_THREADS, 0, sp.text, sp.len,
                (u32)sp.ntiles, desc.p, ticket.p, sp.codes.p, sp.bad.p, st.p);
    else
        LAUNCHN(e, "fn_parse_single_kernel", fn_parse_single_kernel<false>, (unsigned)sp.ntiles, FN_THREADS, 0, sp.text, sp.len,
                (u32)sp.ntiles, desc.p, ticket.p, sp.codes.p, sp.bad.p, st.p);
    const FnStats fs = read_scalar<FnStats>(e, st.p);
    sp.nsym = fs.n_sym;
    if (getenv("MC2_DEBUG_FAST"))
        fprintf(stderr, "[fast_nt] single pass len=%llu n_sym=%llu kept=%llu non_acgt=%llu complex=%llu\n", (ull)sp.len, fs.n_sym,
                fs.packed & 0xFFFFFFFFull, fs.packed >> 32, fs.complex);
    return fs;
}

// write pass: 2-bit codes + bad bits (also counts the kept non-ACGT bytes into st->packed2)
static void fn_write_pass(mc2_engine* e, FnSpan& sp, DBuf<FnStats>& st) {
    sp.codes.alloc(e, div_up(sp.nsym, 16) + 4);
    sp.bad.alloc(e, div_up(sp.nsym, 32) + 4);
    sp.codes.zero();
    sp.bad.zero();
    LAUNCH(e, fn_parse_kernel<1>, (unsigned)sp.ntiles, FN_THREADS, 0, sp.text, sp.len, sp.tstate.p, sp.tcnt.p, (const u64*)sp.toff.p,
           sp.codes.p, sp.bad.p, st.p);
    sp.tstate.release();
    sp.tcnt.release();
    sp.toff.release();
}

static void fn_dense_span(mc2_engine* e, mc2_sample* s, const PackedView& pv) {
    const Plan& plan = s->plan;
    const u64 nwords = div_up(pv.n, 16);
    if (plan.smem) {
        const size_t smem = (size_t)plan.bins * plan.nrep * 4;
        static thread_local bool attr_set = false;
        if (!attr_set) {
            CUDA_CHECK(cudaFuncSetAttribute(fn_dense_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
            attr_set = true;
        }
        const unsigned grid = (unsigned)std::min<u64>(div_up(nwords, FN_HIST_THREADS), (u64)e->num_sms * (smem <= 96 * 1024 ? 2 : 1));
        LAUNCHN(e, "fn_dense_kernel<smem>", fn_dense_kernel<true>, grid, FN_HIST_THREADS, smem, pv, s->k, plan.bins, plan.nrep, s->dense_chunk.p);
    } else {
        const unsigned grid = (unsigned)std::min<u64>(div_up(nwords, FN_HIST_THREADS), (u64)e->num_sms * 2);
        LAUNCHN(e, "fn_dense_kernel<global>", fn_dense_kernel<false>, grid, FN_HIST_THREADS, 0, pv, s->k, plan.bins, 1u, s->dense_chunk.p);
    }
}

// A chunk with more windows than one hash batch holds (-s 0 on a large file): a level-0 partition of ALL its keys by an
// independent hash into groups that fit, written once to HBM (8 B per window -- sized for the 180 GB of a B200), then
// the usual two-level pipeline per group.  All occurrences of a key meet in one group, so the -c filter stays exact
// for the whole chunk.  Returns false (nothing counted) when the keys do not fit in free device memory.
// Level-0 partition of key sources into g0 groups by hash `mult`: keys0 (grouped, exact offsets in gbase[g0 + 1]).
// Sources: packed symbol streams (their windows) or one key array.  Returns false if the result does not fit in
// free device memory (`extra` = bytes the caller still needs afterwards) or a group would exceed `group_max` keys.
struct Level0 {
    DBuf<u64> keys0;
    std::vector<u64> gbase;
    u64 gmax = 0;
};
static bool level0_partition(mc2_engine* e, int k, const std::vector<PackedView>& pvs, const KeySpan* ks, u32 g0, u64 mult,
                             u64 group_max, u64 extra, Level0& out) {
    const u32 nb0 = g0 * HC_NB2;
    DBuf<u32> ghist(e, nb0);
    ghist.zero();
    static thread_local bool attr_set = false;
    if (!attr_set) {
        CUDA_CHECK(cudaFuncSetAttribute(fn_hist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 208 * 1024));
        CUDA_CHECK(cudaFuncSetAttribute(hk_hist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 208 * 1024));
        CUDA_CHECK(cudaFuncSetAttribute(fn_scatter1_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)HC_SCATTER_SMEM));
        CUDA_CHECK(cudaFuncSetAttribute(hk_scatter1_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)HC_SCATTER_SMEM));
        attr_set = true;
    }
    const size_t hist_smem = (size_t)nb0 * 4;
    PhaseTimer pt(e);
    if (ks) {
        int per_sm = 1;
        CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, hk_hist_kernel, HK_HIST_THREADS, hist_smem));
        const u64 grid = std::min<u64>(div_up(ks->n, HK_HIST_THREADS * 8), (u64)e->num_sms * std::max(per_sm, 1));
        if (ks->n) LAUNCH(e, hk_hist_kernel, (unsigned)std::max<u64>(grid, 1), HK_HIST_THREADS, hist_smem, ks->keys, ks->n, nb0, mult, ghist.p);
    } else {
        if (mult != HC_MULT1) throw Mc2Error(MC2_ERR_INVALID, "level-0 partition of packed streams uses the first hash (internal error)");
        int per_sm = 1;
        CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn_hist_kernel, FN_HIST_THREADS, hist_smem));
        for (auto& pv : pvs) {
            const u64 grid = std::min<u64>(div_up(div_up(pv.n, 16), FN_HIST_THREADS), (u64)e->num_sms * std::max(per_sm, 1));
            if (grid) LAUNCH(e, fn_hist_kernel, (unsigned)grid, FN_HIST_THREADS, hist_smem, pv, k, nb0, ghist.p);
        }
    }
    std::vector<u32> h(nb0);
    d2h(e, h.data(), (const u32*)ghist.p, nb0);
    pt.mark("level-0 histogram");
    out.gbase.assign(g0 + 1, 0);
    out.gmax = 0;
    for (u32 g = 0; g < g0; ++g) {
        u64 n = 0;
        for (u32 j = 0; j < HC_NB2; ++j) n += h[(u64)g * HC_NB2 + j];
        out.gbase[g + 1] = out.gbase[g] + n;
        out.gmax = std::max(out.gmax, n);
    }
    const u64 total = out.gbase[g0];
    if (total == 0) return true;
    if (out.gmax > group_max || out.gmax >= (1ull << 32)) return false;
    {   // the level-0 array plus what follows must fit (free memory + what the pool holds unused)
        size_t free_b = 0, total_b = 0;
        CUDA_CHECK(cudaMemGetInfo(&free_b, &total_b));
        cudaMemPool_t pool;
        CUDA_CHECK(cudaDeviceGetDefaultMemPool(&pool, e->device));
        unsigned long long reserved = 0, used = 0;
        CUDA_CHECK(cudaMemPoolGetAttribute(pool, cudaMemPoolAttrReservedMemCurrent, &reserved));
        CUDA_CHECK(cudaMemPoolGetAttribute(pool, cudaMemPoolAttrUsedMemCurrent, &used));
        const u64 avail = (u64)free_b + (reserved > used ? (u64)(reserved - used) : 0);
        if (total * 8 + extra + (1ull << 30) > avail) return false;
    }
    pt.mark("level-0 memory check");
    out.keys0.alloc(e, total);
    DBuf<u64> gbase_dev(e, g0 + 1);
    DBuf<u32> cur0(e, g0);
    cur0.zero();
    pt.mark("level-0 allocation");
    CUDA_CHECK(cudaMemcpyAsync(gbase_dev.p, out.gbase.data(), (g0 + 1) * 8, cudaMemcpyHostToDevice, e->stream));
    if (ks) {
        LAUNCH(e, hk_scatter1_kernel, (unsigned)div_up(ks->n, HC_TILE), EX_THREADS, HC_SCATTER_SMEM16, ks->keys, ks->n, nb0, g0, mult, cur0.p,
               out.keys0.p, (const u64*)gbase_dev.p);
    } else {
        for (auto& pv : pvs) {
            const u64 grid = div_up(div_up(pv.n, 16), EX_THREADS);
            if (grid) LAUNCH(e, fn_scatter1_kernel<false>, (unsigned)grid, EX_THREADS, HC_SCATTER_SMEM16, pv, k, nb0, g0, cur0.p, out.keys0.p,
                             (const u64*)gbase_dev.p);
        }
    }
    CUDA_CHECK(cudaStreamSynchronize(e->stream));                  // (host vector was the source of an async copy)
    pt.mark("level-0 scatter");
    return true;
}

static u32 level0_groups(u64 cap, u64 hash_max) {
    return (u32)std::min<u64>(HC_MAX_NB1, std::max<u64>(2, div_up(cap, std::max<u64>(1, hash_max / 2))));
}

// A chunk with more windows than one hash batch holds (-s 0 on a large file): a level-0 partition of ALL its keys by an
// independent hash into groups that fit, written once to HBM (8 B per window -- sized for the 180 GB of a B200), then
// the usual two-level pipeline per group.  All occurrences of a key meet in one group, so the -c filter stays exact
// for the whole chunk.  Returns false (nothing counted) when the keys do not fit in free device memory.
static bool sparse_chunk_hash_big(mc2_engine* e, mc2_sample* s, const std::vector<PackedView>& pvs, const KeySpan* ks, u64 hash_max) {
    u64 cap = ks ? ks->n : 0;
    for (auto& pv : pvs) cap += pv.n;
    const u32 g0 = level0_groups(cap, hash_max);
    if (div_up(cap, g0) > hash_max) return false;
    Level0 l0;
    if (!level0_partition(e, s->k, pvs, ks, g0, ks ? HC_MULT3 : HC_MULT1, hash_max, 3 * 8 * div_up(cap, g0) * 2, l0)) return false;
    PhaseTimer pt(e);
    for (u32 g = 0; g < g0; ++g) {
        const u64 n = l0.gbase[g + 1] - l0.gbase[g];
        if (!n) continue;
        KeySpan span{l0.keys0.p + l0.gbase[g], n, true};
        sparse_chunk_hash<ENC_NT2>(e, s, SymView{nullptr, 0}, nullptr, &span);
    }
    pt.mark("groups");
    return true;
}

// Spans of at most ~span_bytes, cut where the Chunker would cut (at a line containing '>'); only real header lines
// may separate spans (the Chunker also cuts at a '>' inside a sequence line), so that no window crosses a cut.
// cuts = span starts + len.  false: use the general path.
static bool fn_span_cuts(mc2_engine* e, const u8* dtext, u64 len, std::vector<u64>& cuts) {
    cuts.assign(1, 0);
    if (len > e->opt_span_bytes) {
        cuts = chunk_bounds(e, dtext, len, e->opt_span_bytes);
        if (cuts.empty() || cuts[0] != 0) return false;
    }
    for (size_t i = 1; i < cuts.size(); ++i) {
        u8 first = 0;
        CUDA_CHECK(cudaMemcpyAsync(&first, dtext + cuts[i], 1, cudaMemcpyDeviceToHost, e->stream));
        CUDA_CHECK(cudaStreamSynchronize(e->stream));
        if (first != '>') return false;
    }
    cuts.push_back(len);
    for (size_t i = 0; i + 1 < cuts.size(); ++i)
        if (cuts[i + 1] - cuts[i] >= (1ull << 32)) return false;      // one record of 4 GiB
    return true;
}

static bool count_chunk_fast_nt(mc2_engine* e, mc2_sample* s, const u8* dtext, u64 len, bool* need_exceptions) {
    *need_exceptions = false;
    if (!e->opt_fast_nt || s->k > 32 || len == 0) return false;
    if (e->opt_force_enc > 0 || e->opt_force_path == PATH_WIDE) return false;
    const u64 hash_max = std::min<u64>((u64)HC_MAX_NB1 * HC_NB2 * e->opt_hash_bucket_keys, e->opt_batch_symbols);
    // which plans the packed lane serves: 2-bit sparse keys through the hash tables (min_count >= 2), and dense 4^k tables
    auto served = [&](const Plan& pl) {
        if (pl.enc != ENC_NT2) return false;
        if (pl.path == PATH_DENSE) return s->k <= 15;
        if (pl.path != PATH_SPARSE || s->c < 2 || e->opt_sparse_algo == 1) return false;
        return len <= hash_max || e->opt_big_chunks != 0;
    };
    if (s->plan.path != PATH_UNSET && !served(s->plan)) return false;
    std::vector<u64> cuts;
    if (!fn_span_cuts(e, dtext, len, cuts)) return false;
    const size_t nspans = cuts.size() - 1;
    DBuf<FnStats> st;
    std::vector<FnSpan> spans(nspans);
    // the count pass of this chunk may already have run behind the previous chunk (its statistics arrived with that
    // chunk's final synchronisation)
    const bool prefetched = s->pre.valid && nspans == 1 && s->pre.text == dtext && s->pre.len == len && s->plan.path != PATH_UNSET;
    if (prefetched) {
        spans[0] = std::move(s->pre.sp);
        st = std::move(s->pre.st);
    } else {
        st.alloc(e, 1);
        st.zero();
    }
    if (s->pre.valid) {                                            // consumed or stale
        s->pre.valid = false;
        if (!prefetched) { s->pre.sp = FnSpan(); s->pre.st.release(); }
    }
    u64 nsym_total = 0;
    PhaseTimer pt(e);
    for (size_t i = 0; i < nspans; ++i) {
        FnSpan& sp = spans[i];
        const bool plan_known = s->plan.path != PATH_UNSET;
        const bool single = e->opt_parse_single != 0 && !prefetched;
        FnStats fs;
        if (prefetched) {
            memcpy(&fs, (const u8*)e->pin_small + 3072, sizeof fs);
            sp.nsym = fs.n_sym;
        } else {
            sp.text = dtext + cuts[i];
            sp.len = cuts[i + 1] - cuts[i];
            fs = single ? fn_single_pass(e, sp, !plan_known, st) : fn_count_pass(e, sp, !plan_known, st);
        }
        if (fs.complex) return false;
        if (!plan_known) {
            // kept / ACGT counts are exact from the statistics pass; later spans learn their non-ACGT count in the write pass
            const u64 n_kept = fs.packed & 0xFFFFFFFFull, n_acgt = n_kept - (fs.packed >> 32);
            if (n_kept == 0) {
                if (nspans == 1) return true;                              // headers only: nothing to count, plan stays open
                return false;
            }
            if (n_acgt * 10 < n_kept * 9) return false;                   // not nucleotide-like: let the general path decide
            ParseStats ps;
            memset(&ps, 0, sizeof ps);
            ps.n_acgt = n_acgt;
            ps.n_upper = n_acgt;
            ps.n_ascii = n_kept;
            make_plan(e, ps, s->k, s->plan);
            if (s->plan.path == PATH_DENSE) {
                s->dense_sample.alloc(e, s->plan.bins);
                s->dense_sample.zero();
                s->dense_chunk.alloc(e, s->plan.bins);
                s->dense_chunk.zero();
            }
            if (!served(s->plan)) return false;
        }
        if (sp.nsym && !single) fn_write_pass(e, sp, st);
        nsym_total += sp.nsym;
    }
    if (nsym_total == 0) return true;
    pt.mark("packed parse");
    if (s->plan.path == PATH_DENSE) {
        const Plan& plan = s->plan;
        const unsigned fgrid = (unsigned)div_up(plan.bins, 256);
        const bool wide_counts = nsym_total >= (1ull << 32);              // a u32 bin could wrap: fold span by span into 64 bits
        for (auto& sp : spans) {
            if (!sp.nsym) continue;
            fn_dense_span(e, s, PackedView{sp.codes.p, sp.bad.p, sp.nsym});
            if (wide_counts) {
                if (!s->dense_chunk64.p) { s->dense_chunk64.alloc(e, plan.bins); s->dense_chunk64.zero(); }
                LAUNCH(e, dense_fold_batch_kernel, fgrid, 256, 0, s->dense_chunk.p, s->dense_chunk64.p, plan.bins);
            }
        }
        if (wide_counts) LAUNCH(e, dense_fold64_kernel, fgrid, 256, 0, s->dense_chunk64.p, s->dense_sample.p, plan.bins, s->c);
        else LAUNCH(e, dense_fold_kernel, fgrid, 256, 0, s->dense_chunk.p, s->dense_sample.p, plan.bins, s->c);
        const FnStats fsd = read_scalar<FnStats>(e, st.p);
        *need_exceptions = (fsd.packed2 >> 32) != 0;
        return true;
    }
    if (nspans == 1 && nsym_total <= hash_max) {
        PackedView pv{spans[0].codes.p, spans[0].bad.p, spans[0].nsym};
        // the write pass's statistics come back with the hash path's own final readback (one sync fewer per chunk)
        e->ride_dev = st.p;
        e->ride_len = sizeof(FnStats);
        e->ride_done = false;
        try {
            sparse_chunk_hash<ENC_NT2>(e, s, SymView{nullptr, 0}, &pv);
        } catch (...) {
            e->ride_dev = nullptr;
            throw;
        }
        FnStats fs2;
        if (e->ride_done) memcpy(&fs2, (const u8*)e->pin_small + 2048, sizeof fs2);
        else { e->ride_dev = nullptr; fs2 = read_scalar<FnStats>(e, st.p); }
        *need_exceptions = (fs2.packed2 >> 32) != 0;
        return true;
    }
    if (!e->opt_big_chunks) return false;
    std::vector<PackedView> pvs;
    for (auto& sp : spans)
        if (sp.nsym) pvs.push_back(PackedView{sp.codes.p, sp.bad.p, sp.nsym});
    if (!sparse_chunk_hash_big(e, s, pvs, nullptr, hash_max)) return false;
    const FnStats fs3 = read_scalar<FnStats>(e, st.p);
    *need_exceptions = (fs3.packed2 >> 32) != 0;
    return true;
}

static void count_chunk(mc2_engine* e, mc2_sample* s, const u8* dtext, u64 len, bool raw_symbols = false) {
    s->n_chunks++;
    e->chunks++;
    bool exceptions_only = false;
    if (!raw_symbols) {
        bool need_exc = false;
        if (count_chunk_fast_nt(e, s, dtext, len, &need_exc)) {
            if (!need_exc) return;
            exceptions_only = true;            // fast windows are counted; the general parse below feeds the wide path
        }
    }
    Parsed ps;
    if (raw_symbols) adopt_symbols(e, dtext, len, ps);
    else parse_text(e, dtext, len, 0, ps);
    if (ps.nsym == 0) return;
    if (s->plan.path == PATH_UNSET) {
        if (ps.stats.n_ascii == 0) return;           // only separators so far: decide on a later chunk
        make_plan(e, ps.stats, s->k, s->plan);
        if (s->plan.path == PATH_DENSE) {
            s->dense_sample.alloc(e, s->plan.bins);
            s->dense_sample.zero();
            s->dense_chunk.alloc(e, s->plan.bins);
            s->dense_chunk.zero();
        }
    }
    const Plan& plan = s->plan;
    SymView v{ps.sym.p, ps.nsym};
    const u64 n_fast_syms = plan.enc == ENC_NT2 ? ps.stats.n_acgt : plan.enc == ENC_AA5 ? ps.stats.n_upper : ps.stats.n_ascii;
    const u64 n_slow_syms = ps.stats.n_ascii - n_fast_syms;
    if (plan.path == PATH_WIDE) {
        wide_chunk<ENC_BYTE, 1>(e, s, v, ~0ull);
        return;
    }
    if (exceptions_only) {
        // fast windows were already counted by the packed lane
    } else if (plan.path == PATH_DENSE) {
        if (plan.enc == ENC_NT2) dense_chunk<ENC_NT2>(e, s, v); else dense_chunk<ENC_AA5>(e, s, v);
    } else {
        if (plan.enc == ENC_NT2) sparse_chunk<ENC_NT2>(e, s, v);
        else if (plan.enc == ENC_AA5) sparse_chunk<ENC_AA5>(e, s, v);
        else sparse_chunk<ENC_BYTE>(e, s, v);
    }
    if (n_slow_syms) {
        const u64 hint = std::max<u64>(1024, n_slow_syms * (u64)s->k);
        if (plan.enc == ENC_NT2) wide_chunk<ENC_NT2, 0>(e, s, v, hint);
        else if (plan.enc == ENC_AA5) wide_chunk<ENC_AA5, 0>(e, s, v, hint);
    }
}

// =====================================================================================================
// text upload and chunk boundaries
// =====================================================================================================
static const u8* to_device(mc2_engine* e, const void* text, u64 nbytes, int space, DBuf<u8>& holder) {
    if (space == MC2_DEVICE || nbytes == 0) return (const u8*)text;
    holder.alloc(e, nbytes + 16);
    cudaPointerAttributes attr;
    bool pinned = false;
    if (cudaPointerGetAttributes(&attr, text) == cudaSuccess) pinned = attr.type == cudaMemoryTypeHost;
    else cudaGetLastError();
    if (pinned) {
        CUDA_CHECK(cudaMemcpyAsync(holder.p, text, nbytes, cudaMemcpyHostToDevice, e->stream));
    } else {
        // pageable source: double-buffered pinned staging so the CPU copy overlaps the DMA
        const u8* src = (const u8*)text;
        int slot = 0;
        for (u64 at = 0; at < nbytes; at += mc2_engine::STAGE_BYTES, slot ^= 1) {
            const u64 nb = std::min<u64>(mc2_engine::STAGE_BYTES, nbytes - at);
            CUDA_CHECK(cudaEventSynchronize(e->stage_ev[slot]));
            memcpy(e->pin_stage[slot], src + at, nb);
            CUDA_CHECK(cudaMemcpyAsync(holder.p + at, e->pin_stage[slot], nb, cudaMemcpyHostToDevice, e->stream));
            CUDA_CHECK(cudaEventRecord(e->stage_ev[slot], e->stream));
        }
    }
    e->h2d_bytes += nbytes;
    return holder.p;
}

static std::vector<u64> chunk_bounds(mc2_engine* e, const u8* dtext, u64 n, u64 chunk_bytes) {
    std::vector<u64> bounds(1, 0);
    if (chunk_bytes == 0 || n == 0) return bounds;
    {   // fast path: no '\r' anywhere -> raw offsets are the reference's translated offsets
        const u64 max_bounds = n / chunk_bytes + 2;
        DBuf<u64> db(e, max_bounds);
        DBuf<ull> nbd(e, 1);
        DBuf<u32> flag(e, 1);
        flag.zero();
        LAUNCH(e, chunk_has_cr_kernel, (unsigned)std::min<u64>(div_up(n, 256 * 16 * 4), (u64)e->num_sms * 8), 256, 0, dtext, n, flag.p);
        LAUNCH(e, chunk_chain_kernel, 1, 256, 0, dtext, n, chunk_bytes, (const u32*)flag.p, db.p, max_bounds, nbd.p);
        const u64 nbounds = (u64)read_scalar<ull>(e, nbd.p);
        if (nbounds != ~0ull) {
            if (nbounds > max_bounds) throw Mc2Error(MC2_ERR_LIMIT, "chunker: boundary buffer overflow");
            bounds.resize(nbounds);
            d2h(e, bounds.data(), db.p, nbounds);
            return bounds;
        }
    }
    const u64 ntiles = div_up(n + ((u64)(uintptr_t)dtext & 15ull), CH_TILE);
    DBuf<u32> tcr(e, ntiles), tca(e, ntiles);
    DBuf<u64> ocr(e, ntiles), oca(e, ntiles);
    LAUNCH(e, chunk_candidates_kernel<false>, (unsigned)ntiles, CH_THREADS, 0, dtext, n, tcr.p, tca.p, (const u64*)nullptr,
           (const u64*)nullptr, (u64*)nullptr, (u64*)nullptr);
    dev_exclusive_scan<u32, u64>(e, tcr.p, ocr.p, ntiles, nullptr);
    const u64 ncand = offsets_from_counts(e, tca.p, oca.p, ntiles);
    if (!ncand) return bounds;
    DBuf<u64> cls(e, ncand), ct(e, ncand);
    LAUNCH(e, chunk_candidates_kernel<true>, (unsigned)ntiles, CH_THREADS, 0, dtext, n, tcr.p, tca.p, (const u64*)ocr.p,
           (const u64*)oca.p, cls.p, ct.p);
    const u64 max_bounds = n / chunk_bytes + 2;
    DBuf<u64> db(e, max_bounds);
    DBuf<ull> nbd(e, 1);
    LAUNCH(e, chunk_select_kernel, 1, 32, 0, (const u64*)cls.p, (const u64*)ct.p, ncand, chunk_bytes, db.p, max_bounds, nbd.p);
    const u64 nbounds = (u64)read_scalar<ull>(e, nbd.p);
    if (nbounds > max_bounds) throw Mc2Error(MC2_ERR_LIMIT, "chunker: boundary buffer overflow");
    bounds.resize(nbounds);
    d2h(e, bounds.data(), db.p, nbounds);
    return bounds;
}

// Pipelined upload: the text goes to the device in pieces on the copy stream (straight from pinned memory, or
// through the two pinned staging buffers for a pageable source) while the compute stream already chunks and
// counts the pieces that have landed.
struct Uploader {
    mc2_engine* e;
    const u8* src;
    u8* dst;
    u64 n, piece, issued = 0;
    bool pinned = false;
    std::vector<cudaEvent_t> done;        // done[j]: piece j is on the device
    int slot = 0;
    Uploader(mc2_engine* e_, const void* text, u64 nbytes, u8* dst_) : e(e_), src((const u8*)text), dst(dst_), n(nbytes) {
        piece = mc2_engine::STAGE_BYTES;
        cudaPointerAttributes attr;
        if (cudaPointerGetAttributes(&attr, text) == cudaSuccess) pinned = attr.type == cudaMemoryTypeHost;
        else cudaGetLastError();
        done.resize(div_up(n, piece), nullptr);
    }
    ~Uploader() {
        for (auto ev : done) if (ev) e->ev_pool.push_back(ev);
    }
    // make sure everything below `upto` has been issued; the compute stream then waits for it on the device
    void need(u64 upto) {
        upto = std::min(upto, n);
        while (issued < upto) {
            const u64 j = issued / piece, len = std::min(piece, n - issued);
            if (pinned) {
                CUDA_CHECK(cudaMemcpyAsync(dst + issued, src + issued, len, cudaMemcpyHostToDevice, e->copy_stream));
            } else {
                CUDA_CHECK(cudaEventSynchronize(e->stage_ev[slot]));
                memcpy(e->pin_stage[slot], src + issued, len);
                CUDA_CHECK(cudaMemcpyAsync(dst + issued, e->pin_stage[slot], len, cudaMemcpyHostToDevice, e->copy_stream));
                CUDA_CHECK(cudaEventRecord(e->stage_ev[slot], e->copy_stream));
                slot ^= 1;
            }
            done[j] = e->get_event();
            CUDA_CHECK(cudaEventRecord(done[j], e->copy_stream));
            issued += len;
        }
        if (upto) CUDA_CHECK(cudaStreamWaitEvent(e->stream, done[(upto - 1) / piece], 0));
    }
    // with a pinned source all copies can be queued at once (they run in order on the copy stream)
    void issue_all() { if (pinned) { const u64 keep = issued; (void)keep; while (issued < n) need_issue_only(); } }
    void need_issue_only() {
        const u64 j = issued / piece, len = std::min(piece, n - issued);
        CUDA_CHECK(cudaMemcpyAsync(dst + issued, src + issued, len, cudaMemcpyHostToDevice, e->copy_stream));
        done[j] = e->get_event();
        CUDA_CHECK(cudaEventRecord(done[j], e->copy_stream));
        issued += len;
    }
};

static void sample_add(mc2_sample* s, const void* text, u64 nbytes, int space, u64 chunk_bytes, u64* n_chunks,
                       std::vector<u64>* bounds_out) {
    mc2_engine* e = s->e;
    std::vector<u64> bounds(1, 0);
    if (space == MC2_DEVICE || nbytes == 0 || chunk_bytes == 0 || nbytes <= 2 * mc2_engine::STAGE_BYTES) {
        // resident text (or a single piece): find all boundaries at once
        DBuf<u8> holder;
        const u8* d = to_device(e, text, nbytes, space, holder);
        bounds = chunk_bounds(e, d, nbytes, chunk_bytes);
        for (size_t i = 0; i < bounds.size(); ++i) {
            const u64 a = bounds[i], b = i + 1 < bounds.size() ? bounds[i + 1] : nbytes;
            if (i + 1 < bounds.size()) {
                s->next_text = d + b;
                s->next_len = (i + 2 < bounds.size() ? bounds[i + 2] : nbytes) - b;
            } else {
                s->next_len = 0;
            }
            count_chunk(e, s, d + a, b - a);
        }
        s->next_len = 0;
        if (s->pre.valid) { s->pre.valid = false; s->pre.sp = FnSpan(); s->pre.st.release(); }
    } else {
        // host text, chunked: overlap the upload with chunking + counting.  The boundary after `b` is the first
        // candidate line whose translated offset from b reaches chunk_bytes (lib/mercat2_Chunker.py:45-52); it is
        // searched in the window [b, b + chunk_bytes + margin) and the window grows until it is found.
        DBuf<u8> holder(e, nbytes + 16);
        CUDA_CHECK(cudaStreamSynchronize(e->stream));            // the buffer exists before the copy stream writes it
        Uploader up(e, text, nbytes, holder.p);
        up.issue_all();
        e->h2d_bytes += nbytes;
        const u64 margin = 4ull << 20;
        u64 b = 0;
        while (b < nbytes) {
            u64 win_end = std::min(nbytes, b + chunk_bytes + margin);
            u64 nb = nbytes;
            while (true) {
                up.need(win_end);
                const std::vector<u64> wb = chunk_bounds(e, holder.p + b, win_end - b, chunk_bytes);
                if (wb.size() >= 2) { nb = b + wb[1]; break; }
                if (win_end == nbytes) { nb = nbytes; break; }
                win_end = std::min(nbytes, win_end + chunk_bytes);
            }
            count_chunk(e, s, holder.p + b, nb - b);
            b = nb;
            if (b < nbytes) bounds.push_back(b);
        }
        CUDA_CHECK(cudaStreamSynchronize(e->stream));
    }
    if (n_chunks) *n_chunks = bounds.size();
    if (bounds_out) *bounds_out = bounds;
}

#include "filestream.inl"

static mc2_table* sample_finish(mc2_sample* s) {
    mc2_engine* e = s->e;
    PhaseTimer pt(e);
    std::unique_ptr<mc2_table> t(new mc2_table);
    t->e = e;
    t->k = s->k;
    t->enc = s->plan.enc < 0 ? ENC_NT2 : s->plan.enc;
    if (s->plan.path == PATH_DENSE) {
        const u32 bins = s->plan.bins;
        const u64 ntiles = div_up(bins, 256);
        DBuf<u32> tc(e, ntiles);
        DBuf<u64> to(e, ntiles);
        LAUNCH(e, dense_nonzero_count_kernel, (unsigned)ntiles, 256, 0, (const u64*)s->dense_sample.p, bins, tc.p);
        const u64 ns = offsets_from_counts(e, tc.p, to.p, ntiles);
        t->fast.n = ns;
        t->fast.keys.alloc(e, ns);
        t->fast.counts.alloc(e, ns);
        if (ns)
            LAUNCH(e, dense_nonzero_write_kernel, (unsigned)ntiles, 256, 0, (const u64*)s->dense_sample.p, bins, (const u64*)to.p,
                   t->fast.keys.p, t->fast.counts.p);
        t->key_kind = s->plan.enc == ENC_AA5 ? KEY_DENSE_AA : KEY_CODE;
    } else if (!s->fast.empty()) {
        reduce_fast_parts(e, s->fast, s->k * enc_bits(t->enc), 1, t->fast);
    }
    pt.mark("finish: packed rows");
    if (!s->wide.empty()) reduce_wide_parts(e, s->wide, s->k, 1, t->wide);
    CUDA_CHECK(cudaStreamSynchronize(e->stream));
    pt.mark("finish: literal rows");
    return t.release();
}

// =====================================================================================================
// export
// =====================================================================================================
static void decode_key(const mc2_table* t, u64 key, char* out) {
    const int k = t->k;
    if (t->key_kind == KEY_DENSE_AA) {
        for (int j = k - 1; j >= 0; --j) { out[j] = (char)('A' + key % 26); key /= 26; }
    } else if (t->enc == ENC_NT2) {
        for (int j = k - 1; j >= 0; --j) { out[j] = "ACGT"[key & 3]; key >>= 2; }
    } else if (t->enc == ENC_AA5) {
        for (int j = k - 1; j >= 0; --j) { out[j] = (char)('A' + (key & 31)); key >>= 5; }
    } else {
        for (int j = k - 1; j >= 0; --j) { out[j] = (char)(key & 255); key >>= 8; }
    }
}

static void ensure_host(mc2_table* t) {
    if (t->on_host) return;
    mc2_engine* e = t->e;
    const u64 nf = t->fast.n, nw = t->wide.n, k = t->k;
    std::vector<u64> fk(nf), fc(nf), wc(nw);
    std::vector<u8> wr(nw * k);
    d2h(e, fk.data(), t->fast.keys.p, nf);
    d2h(e, fc.data(), t->fast.counts.p, nf);
    d2h(e, wc.data(), t->wide.counts.p, nw);
    d2h(e, wr.data(), t->wide.rows.p, nw * k);
    t->kmers.resize((nf + nw) * k);
    t->counts.resize(nf + nw);
    std::vector<char> tmp(k + 1);
    u64 i = 0, j = 0, o = 0, total = 0;
    bool have = false;
    while (i < nf || j < nw) {
        bool take_fast;
        if (i < nf && !have) { decode_key(t, fk[i], tmp.data()); have = true; }
        if (i >= nf) take_fast = false;
        else if (j >= nw) take_fast = true;
        else take_fast = memcmp(tmp.data(), wr.data() + j * k, k) < 0;     // the two sets are disjoint
        if (take_fast) { memcpy(&t->kmers[o * k], tmp.data(), k); t->counts[o] = fc[i]; ++i; have = false; }
        else { memcpy(&t->kmers[o * k], wr.data() + j * k, k); t->counts[o] = wc[j]; ++j; }
        total += t->counts[o];
        ++o;
    }
    t->total = total;
    t->on_host = true;
}

// TSV body (every row, no header line) formatted on the device; returns its size in bytes.
static u64 tsv_body_device(mc2_table* t, DBuf<u8>& body) {
    mc2_engine* e = t->e;
    const u64 nf = t->fast.n, nw = t->wide.n, rows = nf + nw;
    if (!rows) return 0;
    if (nf && t->k > 32) throw Mc2Error(MC2_ERR_INVALID, "tsv: packed rows with k > 32 (internal error)");
    const int kind = t->key_kind == KEY_DENSE_AA ? TSV_DENSE_AA : t->enc == ENC_NT2 ? TSV_NT2 : t->enc == ENC_AA5 ? TSV_AA5 : TSV_BYTE;
    DBuf<u64> pos(e, rows), off(e, rows);
    DBuf<u32> len(e, rows);
    DBuf<ull> total(e, 1);
    const unsigned grid = (unsigned)div_up(rows, 256);
    LAUNCH(e, tsv_place_kernel, grid, 256, 0, (const u64*)t->fast.keys.p, (const u64*)t->fast.counts.p, nf, (const u8*)t->wide.rows.p,
           (const u64*)t->wide.counts.p, nw, t->k, kind, pos.p, len.p);
    dev_exclusive_scan<u32, u64>(e, len.p, off.p, rows, total.p);
    const u64 nbytes = (u64)read_scalar<ull>(e, total.p);
    body.alloc(e, nbytes);
    LAUNCH(e, tsv_write_kernel, grid, 256, 0, (const u64*)t->fast.keys.p, (const u64*)t->fast.counts.p, nf, (const u8*)t->wide.rows.p,
           (const u64*)t->wide.counts.p, nw, t->k, kind, (const u64*)pos.p, (const u64*)off.p, body.p);
    return nbytes;
}

// Device -> host in 32 MiB pieces through the engine's two pinned buffers: while piece j is handed to `sink`
// (memcpy or fwrite), piece j+1 is already crossing PCIe on the copy stream.
template <class Sink>
static void download_pipelined(mc2_engine* e, const u8* dev, u64 nbytes, Sink sink) {
    if (!nbytes) return;
    CUDA_CHECK(cudaStreamSynchronize(e->stream));                 // the producer kernels ran on the compute stream
    const u64 piece = mc2_engine::STAGE_BYTES;
    const u64 np = div_up(nbytes, piece);
    auto issue = [&](u64 j) {
        const u64 o = j * piece, m = std::min(piece, nbytes - o);
        CUDA_CHECK(cudaMemcpyAsync(e->pin_stage[j & 1], dev + o, m, cudaMemcpyDeviceToHost, e->copy_stream));
        CUDA_CHECK(cudaEventRecord(e->stage_ev[j & 1], e->copy_stream));
    };
    CUDA_CHECK(cudaEventSynchronize(e->stage_ev[0]));
    CUDA_CHECK(cudaEventSynchronize(e->stage_ev[1]));
    issue(0);
    for (u64 j = 0; j < np; ++j) {
        if (j + 1 < np) issue(j + 1);
        CUDA_CHECK(cudaEventSynchronize(e->stage_ev[j & 1]));
        const u64 o = j * piece, m = std::min(piece, nbytes - o);
        sink(e->pin_stage[j & 1], o, m);
    }
    e->d2h_bytes += nbytes;
}

static std::string tsv_header(const char* basename) { return std::string("k-mer\t") + basename + "_Count\n"; }

// =====================================================================================================
// C ABI
// =====================================================================================================
#define API_BEGIN try {
#define API_END                                                   \
    }                                                             \
    catch (const Mc2Error& err) { g_err = err.what(); return err.code; } \
    catch (const std::exception& err) { g_err = err.what(); return MC2_ERR_INVALID; } \
    return MC2_OK;

extern "C" {

const char* mc2_last_error(void) { return g_err.c_str(); }
const char* mc2_version(void) { return "mercat2_b200 0.1 (sm_100a)"; }

int mc2_engine_create(int device, mc2_engine** out) {
    API_BEGIN
    if (!out) throw Mc2Error(MC2_ERR_INVALID, "out is NULL");
    int ndev = 0;
    cudaError_t st = cudaGetDeviceCount(&ndev);
    if (st != cudaSuccess || ndev == 0)
        throw Mc2Error(MC2_ERR_CUDA, std::string("no CUDA device available (") + cudaGetErrorString(st) +
                                         "); mercat2_b200 has no CPU fallback");
    if (device < 0 || device >= ndev) throw Mc2Error(MC2_ERR_INVALID, "bad device index");
    CUDA_CHECK(cudaSetDevice(device));
    std::unique_ptr<mc2_engine> e(new mc2_engine);
    e->device = device;
    cudaDeviceProp prop;
    CUDA_CHECK(cudaGetDeviceProperties(&prop, device));
    e->num_sms = prop.multiProcessorCount;
    CUDA_CHECK(cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking));
    CUDA_CHECK(cudaStreamCreateWithFlags(&e->copy_stream, cudaStreamNonBlocking));
    CUDA_CHECK(cudaMallocHost(&e->pin_small, 4096));
    for (int i = 0; i < 2; ++i) {
        CUDA_CHECK(cudaMallocHost((void**)&e->pin_stage[i], mc2_engine::STAGE_BYTES));
        CUDA_CHECK(cudaEventCreateWithFlags(&e->stage_ev[i], cudaEventDisableTiming));
    }
    CUDA_CHECK(cudaEventCreate(&e->ev0));
    CUDA_CHECK(cudaEventCreate(&e->ev1));
    cudaMemPool_t pool;
    CUDA_CHECK(cudaDeviceGetDefaultMemPool(&pool, device));
    u64 thresh = ~0ull;                       // keep freed blocks cached in the pool
    CUDA_CHECK(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thresh));
    *out = e.release();
    API_END
}

void mc2_engine_destroy(mc2_engine* e) {
    if (!e) return;
    cudaSetDevice(e->device);
    cudaStreamSynchronize(e->stream);
    for (auto& r : e->prof_pending) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
    for (auto ev : e->ev_pool) cudaEventDestroy(ev);
    for (int i = 0; i < 2; ++i) {
        if (e->pin_stage[i]) cudaFreeHost(e->pin_stage[i]);
        if (e->file_pin[2 * i]) cudaFreeHost(e->file_pin[2 * i]);
        if (e->file_pin[2 * i + 1]) cudaFreeHost(e->file_pin[2 * i + 1]);
        if (e->stage_ev[i]) cudaEventDestroy(e->stage_ev[i]);
    }
    if (e->pin_small) cudaFreeHost(e->pin_small);
    if (e->ev0) cudaEventDestroy(e->ev0);
    if (e->ev1) cudaEventDestroy(e->ev1);
    if (e->stream) cudaStreamDestroy(e->stream);
    if (e->copy_stream) cudaStreamDestroy(e->copy_stream);
    delete e;
}

int mc2_engine_set_option(mc2_engine* e, const char* name, int64_t value) {
    API_BEGIN
    if (!e || !name) throw Mc2Error(MC2_ERR_INVALID, "NULL argument");
    const std::string n(name);
    if (n == "dense_max_bins") e->opt_dense_max_bins = (u64)value;
    else if (n == "smem_max_bins") e->opt_smem_max_bins = (u64)value;
    else if (n == "batch_symbols") e->opt_batch_symbols = (u64)std::max<int64_t>(value, EX_TILE);
    else if (n == "force_path") e->opt_force_path = (int)value;
    else if (n == "force_encoding") e->opt_force_enc = (int)value;
    else if (n == "sparse_algo") e->opt_sparse_algo = (int)value;
    else if (n == "fast_nt") e->opt_fast_nt = (int)value;
    else if (n == "scatter_variant") e->opt_scatter_variant = (int)value;
    else if (n == "count_variant") e->opt_count_variant = (int)value;
    else if (n == "big_chunks") e->opt_big_chunks = (int)value;
    else if (n == "parse_single") e->opt_parse_single = (int)value;
    else if (n == "prefetch_pass") e->opt_prefetch_pass = (int)value;
    else if (n == "file_piece_bytes") e->opt_file_piece = (u64)std::max<int64_t>(4096, value);
    else if (n == "span_bytes") e->opt_span_bytes = value < 4096 ? 4096 : (value > (3ull << 30) ? (3ull << 30) : (u64)value);
    else if (n == "hash_bucket_keys") e->opt_hash_bucket_keys = (u64)std::max<int64_t>(value, 16);
    else if (n == "profile") { e->resolve_profile(); e->profile = value ? 1 : 0; if (value == 2) e->prof_total.clear(); }
    else throw Mc2Error(MC2_ERR_INVALID, "unknown option " + n);
    API_END
}

int64_t mc2_engine_get_stat(mc2_engine* e, const char* name) {
    if (!e || !name) return -1;
    const std::string n(name);
    if (n == "launches") return (int64_t)e->launches;
    if (n == "h2d_bytes") return (int64_t)e->h2d_bytes;
    if (n == "d2h_bytes") return (int64_t)e->d2h_bytes;
    if (n == "chunks") return (int64_t)e->chunks;
    if (n == "overflow_buckets") return (int64_t)e->ovf_buckets;
    if (n == "device_us") return (int64_t)e->device_us;
    if (n == "num_sms") return e->num_sms;
    return -1;
}

int mc2_engine_profile(mc2_engine* e, char* buf, uint64_t cap, uint64_t* size) {
    API_BEGIN
    if (!e) throw Mc2Error(MC2_ERR_INVALID, "engine is NULL");
    CUDA_CHECK(cudaSetDevice(e->device));
    e->resolve_profile();
    std::string js = "{";
    bool first = true;
    for (auto& kv : e->prof_total) {
        char tmp[256];
        snprintf(tmp, sizeof tmp, "%s\"%s\": {\"launches\": %llu, \"us\": %.3f}", first ? "" : ", ", kv.first.c_str(),
                 (ull)kv.second.first, kv.second.second);
        js += tmp;
        first = false;
    }
    js += "}";
    if (size) *size = js.size();
    if (buf) {
        if (cap < js.size()) throw Mc2Error(MC2_ERR_INVALID, "buffer too small");
        memcpy(buf, js.data(), js.size());
    }
    API_END
}

static void check_count_args(mc2_engine* e, const void* text, u64 nbytes, int k) {
    if (!e) throw Mc2Error(MC2_ERR_INVALID, "engine is NULL");
    if (!text && nbytes) throw Mc2Error(MC2_ERR_INVALID, "text is NULL");
    if (k < 1) throw Mc2Error(MC2_ERR_INVALID, "k must be >= 1");
    if (k > 128) throw Mc2Error(MC2_ERR_LIMIT, "k > 128 is not supported");
}

int mc2_sample_begin(mc2_engine* e, int k, int64_t min_count, mc2_sample** out) {
    API_BEGIN
    check_count_args(e, "", 0, k);
    if (!out) throw Mc2Error(MC2_ERR_INVALID, "out is NULL");
    CUDA_CHECK(cudaSetDevice(e->device));
    mc2_sample* s = new mc2_sample;
    s->e = e;
    s->k = k;
    s->c = min_count < 1 ? 1 : (u64)min_count;
    *out = s;
    API_END
}

int mc2_sample_add_text(mc2_sample* s, const void* text, uint64_t nbytes, int space, uint64_t chunk_bytes, uint64_t* n_chunks) {
    API_BEGIN
    if (!s) throw Mc2Error(MC2_ERR_INVALID, "sample is NULL");
    check_count_args(s->e, text, nbytes, s->k);
    CUDA_CHECK(cudaSetDevice(s->e->device));
    u64 nc = 0;
    sample_add(s, text, nbytes, space, chunk_bytes, &nc, nullptr);
    if (n_chunks) *n_chunks = nc;
    API_END
}

int mc2_sample_add_file(mc2_sample* s, const char* path, int gunzip, uint64_t chunk_bytes, uint64_t* n_chunks, uint64_t* text_bytes) {
    API_BEGIN
    if (!s || !path) throw Mc2Error(MC2_ERR_INVALID, "NULL argument");
    CUDA_CHECK(cudaSetDevice(s->e->device));
    bool gz = gunzip > 0;
    if (gunzip < 0) {
        const size_t n = strlen(path);
        gz = n >= 3 && strcmp(path + n - 3, ".gz") == 0;
    }
    u64 nc = 0, nb = 0;
    sample_add_file(s, path, gz, chunk_bytes, &nc, &nb);
    if (n_chunks) *n_chunks = nc;
    if (text_bytes) *text_bytes = nb;
    API_END
}

int mc2_sample_add_rows(mc2_sample* s, const char* kmers, const uint64_t* counts, uint64_t rows) {
    API_BEGIN
    if (!s || (rows && (!kmers || !counts))) throw Mc2Error(MC2_ERR_INVALID, "NULL argument");
    mc2_engine* e = s->e;
    CUDA_CHECK(cudaSetDevice(e->device));
    if (rows) {
        WidePart part;
        part.n = rows;
        part.sorted = false;
        part.rows.alloc(e, rows * (u64)s->k);
        part.counts.alloc(e, rows);
        CUDA_CHECK(cudaMemcpyAsync(part.rows.p, kmers, rows * (u64)s->k, cudaMemcpyHostToDevice, e->stream));
        CUDA_CHECK(cudaMemcpyAsync(part.counts.p, counts, rows * 8, cudaMemcpyHostToDevice, e->stream));
        CUDA_CHECK(cudaStreamSynchronize(e->stream));
        e->h2d_bytes += rows * ((u64)s->k + 8);
        s->wide.push_back(std::move(part));
    }
    API_END
}

int mc2_table_info(const mc2_table* t, int* encoding, int* key_kind, uint64_t* packed_rows, uint64_t* wide_rows) {
    API_BEGIN
    if (!t) throw Mc2Error(MC2_ERR_INVALID, "NULL argument");
    if (encoding) *encoding = t->enc;
    if (key_kind) *key_kind = t->key_kind;
    if (packed_rows) *packed_rows = t->fast.n;
    if (wide_rows) *wide_rows = t->wide.n;
    API_END
}

int mc2_table_device_rows(mc2_table* t, const uint64_t** keys, const uint64_t** counts, uint64_t* rows) {
    API_BEGIN
    if (!t || !keys || !counts || !rows) throw Mc2Error(MC2_ERR_INVALID, "NULL argument");
    *keys = (const uint64_t*)t->fast.keys.p;
    *counts = (const uint64_t*)t->fast.counts.p;
    *rows = t->fast.n;
    API_END
}

int mc2_table_lower_bound(mc2_table* t, const uint64_t* splitters, uint64_t m, uint64_t* cuts) {
    API_BEGIN
    if (!t || (m && (!splitters || !cuts))) throw Mc2Error(MC2_ERR_INVALID, "NULL argument");
    if (!m) return MC2_OK;
    mc2_engine* e = t->e;
    CUDA_CHECK(cudaSetDevice(e->device));
    DBuf<u64> sp(e, m), out(e, m);
    CUDA_CHECK(cudaMemcpyAsync(sp.p, splitters, m * 8, cudaMemcpyHostToDevice, e->stream));
    LAUNCH(e, lower_bound_kernel, (unsigned)div_up(m, 128), 128, 0, (const u64*)t->fast.keys.p, (u64)t->fast.n, (const u64*)sp.p, m, out.p);
    d2h(e, (u64*)cuts, (const u64*)out.p, m);
    API_END
}

int mc2_table_export_wide(mc2_table* t, char* kmers, uint64_t* counts) {
    API_BEGIN
    if (!t) throw Mc2Error(MC2_ERR_INVALID, "NULL argument");
    if (t->wide.n && (!kmers || !counts)) throw Mc2Error(MC2_ERR_INVALID, "NULL buffer");
    CUDA_CHECK(cudaSetDevice(t->e->device));
    d2h(t->e, (u8*)kmers, (const u8*)t->wide.rows.p, t->wide.n * (u64)t->k);
    d2h(t->e, (u64*)counts, (const u64*)t->wide.counts.p, t->wide.n);
    API_END
}

int mc2_table_from_rows(mc2_engine* e, int k, int encoding, int key_kind, const uint64_t* keys, const uint64_t* counts,
                        uint64_t rows, int space, const char* wide_kmers, const uint64_t* wide_counts, uint64_t wide_rows,
                        mc2_table** out) {
    API_BEGIN
    if (!e || !out) throw Mc2Error(MC2_ERR_INVALID, "NULL argument");
    if (k < 1 || k > 128) throw Mc2Error(MC2_ERR_INVALID, "k out of range");
    if (encoding < ENC_NT2 || encoding > ENC_BYTE) throw Mc2Error(MC2_ERR_INVALID, "unknown encoding");
    if ((rows && (!keys || !counts)) || (wide_rows && (!wide_kmers || !wide_counts))) throw Mc2Error(MC2_ERR_INVALID, "NULL rows");
    if (rows && k * enc_bits(encoding) > 64) throw Mc2Error(MC2_ERR_INVALID, "packed rows need k * bits <= 64");
    CUDA_CHECK(cudaSetDevice(e->device));
    mc2_sample s;
    s.e = e;
    s.k = k;
    s.c = 1;
    s.plan.enc = encoding;
    s.plan.path = PATH_SPARSE;
    if (rows) {
        FastPart part;
        part.n = rows;
        part.sorted = false;
        part.keys.alloc(e, rows);
        part.counts.alloc(e, rows);
        const cudaMemcpyKind kind = space == MC2_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
        CUDA_CHECK(cudaMemcpyAsync(part.keys.p, keys, rows * 8, kind, e->stream));
        CUDA_CHECK(cudaMemcpyAsync(part.counts.p, counts, rows * 8, kind, e->stream));
        if (space != MC2_DEVICE) e->h2d_bytes += rows * 16;
        s.fast.push_back(std::move(part));
    }
    if (wide_rows) {
        WidePart part;
        part.n = wide_rows;
        part.sorted = false;
        part.rows.alloc(e, wide_rows * (u64)k);
        part.counts.alloc(e, wide_rows);
        CUDA_CHECK(cudaMemcpyAsync(part.rows.p, wide_kmers, wide_rows * (u64)k, cudaMemcpyHostToDevice, e->stream));
        CUDA_CHECK(cudaMemcpyAsync(part.counts.p, wide_counts, wide_rows * 8, cudaMemcpyHostToDevice, e->stream));
        e->h2d_bytes += wide_rows * ((u64)k + 8);
        s.wide.push_back(std::move(part));
    }
    CUDA_CHECK(cudaStreamSynchronize(e->stream));                 // the source buffers may be released by the caller
    mc2_table* t = sample_finish(&s);
    t->key_kind = key_kind == KEY_DENSE_AA ? KEY_DENSE_AA : KEY_CODE;
    *out = t;
    API_END
}

int mc2_table_tsv_body(mc2_table* t, char* buf, uint64_t cap, uint64_t* size) {
    API_BEGIN
    if (!t) throw Mc2Error(MC2_ERR_INVALID, "NULL argument");
    CUDA_CHECK(cudaSetDevice(t->e->device));
    if (!t->tsv_ready) {
        t->tsv_bytes = tsv_body_device(t, t->tsv_body);
        t->tsv_ready = true;
    }
    if (size) *size = t->tsv_bytes;
    if (buf) {
        if (cap < t->tsv_bytes) throw Mc2Error(MC2_ERR_INVALID, "buffer too small");
        download_pipelined(t->e, (const u8*)t->tsv_body.p, t->tsv_bytes, [&](const u8* src, u64 o, u64 m) { memcpy(buf + o, src, m); });
        t->tsv_body.release();
        t->tsv_ready = false;
    }
    API_END
}

int mc2_sample_dense(mc2_sample* s, uint64_t** table, uint64_t* bins, int* encoding) {
    API_BEGIN
    if (!s || !table || !bins) throw Mc2Error(MC2_ERR_INVALID, "NULL argument");
    CUDA_CHECK(cudaSetDevice(s->e->device));
    CUDA_CHECK(cudaStreamSynchronize(s->e->stream));
    const bool dense = s->plan.path == PATH_DENSE;
    *table = dense ? (uint64_t*)s->dense_sample.p : nullptr;
    *bins = dense ? s->plan.bins : 0;
    if (encoding) *encoding = s->plan.enc;
    API_END
}

int mc2_device_copy(mc2_engine* e, void* dst, const void* src, uint64_t nbytes) {
    API_BEGIN
    if (!e || (nbytes && (!dst || !src))) throw Mc2Error(MC2_ERR_INVALID, "NULL argument");
    CUDA_CHECK(cudaSetDevice(e->device));
    if (nbytes) CUDA_CHECK(cudaMemcpyAsync(dst, src, nbytes, cudaMemcpyDeviceToDevice, e->stream));
    CUDA_CHECK(cudaStreamSynchronize(e->stream));
    API_END
}

int mc2_sample_dense_plan(mc2_sample* s, int encoding) {
    API_BEGIN
    if (!s) throw Mc2Error(MC2_ERR_INVALID, "NULL argument");
    if (s->plan.path != PATH_UNSET) throw Mc2Error(MC2_ERR_INVALID, "the sample already has a plan");
    if (encoding != ENC_NT2 && encoding != ENC_AA5) throw Mc2Error(MC2_ERR_INVALID, "dense tables exist for encodings 0 and 1");
    mc2_engine* e = s->e;
    CUDA_CHECK(cudaSetDevice(e->device));
    ParseStats ps;
    memset(&ps, 0, sizeof ps);
    ps.n_ascii = ps.n_upper = 1;                                 // shape the statistics so that make_plan picks `encoding`
    ps.n_acgt = encoding == ENC_NT2 ? 1 : 0;
    make_plan(e, ps, s->k, s->plan);
    if (s->plan.path != PATH_DENSE || s->plan.enc != encoding) {
        s->plan = Plan();
        throw Mc2Error(MC2_ERR_INVALID, "k too large for a dense table of this encoding");
    }
    s->dense_sample.alloc(e, s->plan.bins);
    s->dense_sample.zero();
    s->dense_chunk.alloc(e, s->plan.bins);
    s->dense_chunk.zero();
    CUDA_CHECK(cudaStreamSynchronize(e->stream));
    API_END
}

struct mc2_keys {
    mc2_engine* e = nullptr;
    int k = 0;
    u32 groups = 0;
    Level0 l0;
    u64 exception_symbols = 0;
};

int mc2_partition_keys(mc2_engine* e, const void* text, uint64_t nbytes, int space, int k, uint32_t groups, mc2_keys** out) {
    API_BEGIN
    check_count_args(e, text, nbytes, k);
    if (!out) throw Mc2Error(MC2_ERR_INVALID, "out is NULL");
    if (k > 32) throw Mc2Error(MC2_ERR_LIMIT, "key partition needs k <= 32 (2-bit packed keys)");
    if (groups < 1 || groups > HC_MAX_NB1) throw Mc2Error(MC2_ERR_INVALID, "groups must be in [1, 400]");
    CUDA_CHECK(cudaSetDevice(e->device));
    std::unique_ptr<mc2_keys> ks(new mc2_keys);
    ks->e = e;
    ks->k = k;
    ks->groups = groups;
    ks->l0.gbase.assign(groups + 1, 0);
    if (nbytes) {
        DBuf<u8> holder;
        const u8* d = to_device(e, text, nbytes, space, holder);
        std::vector<u64> cuts;
        if (!fn_span_cuts(e, d, nbytes, cuts)) throw Mc2Error(MC2_ERR_LIMIT, "key partition: text cannot be cut into spans at header lines");
        DBuf<FnStats> st(e, 1);
        st.zero();
        std::vector<FnSpan> spans(cuts.size() - 1);
        std::vector<PackedView> pvs;
        for (size_t i = 0; i + 1 < cuts.size(); ++i) {
            FnSpan& sp = spans[i];
            sp.text = d + cuts[i];
            sp.len = cuts[i + 1] - cuts[i];
            const bool single = e->opt_parse_single != 0;
            const FnStats fs = single ? fn_single_pass(e, sp, false, st) : fn_count_pass(e, sp, false, st);
            if (fs.complex) throw Mc2Error(MC2_ERR_LIMIT, "key partition: text is not plain FASTA (whitespace, '*' or non-ASCII bytes in sequence lines)");
            if (!sp.nsym) continue;
            if (!single) fn_write_pass(e, sp, st);
            pvs.push_back(PackedView{sp.codes.p, sp.bad.p, sp.nsym});
        }
        if (!level0_partition(e, k, pvs, nullptr, groups, HC_MULT1, (1ull << 32) - 1, 0, ks->l0))
            throw Mc2Error(MC2_ERR_LIMIT, "key partition: the keys do not fit in free device memory");
        const FnStats fs2 = read_scalar<FnStats>(e, st.p);
        ks->exception_symbols = fs2.packed2 >> 32;
    }
    CUDA_CHECK(cudaStreamSynchronize(e->stream));
    *out = ks.release();
    API_END
}

int mc2_keys_info(mc2_keys* ks, const uint64_t** keys, uint64_t* sizes, uint64_t* total, uint64_t* exception_symbols) {
    API_BEGIN
    if (!ks) throw Mc2Error(MC2_ERR_INVALID, "NULL argument");
    if (keys) *keys = (const uint64_t*)ks->l0.keys0.p;
    if (sizes)
        for (u32 g = 0; g < ks->groups; ++g) sizes[g] = ks->l0.gbase[g + 1] - ks->l0.gbase[g];
    if (total) *total = ks->l0.gbase[ks->groups];
    if (exception_symbols) *exception_symbols = ks->exception_symbols;
    API_END
}

void mc2_keys_free(mc2_keys* ks) {
    if (!ks) return;
    cudaSetDevice(ks->e->device);
    delete ks;
}

int mc2_sample_add_keys(mc2_sample* s, const uint64_t* keys, uint64_t n, int space) {
    API_BEGIN
    if (!s || (n && !keys)) throw Mc2Error(MC2_ERR_INVALID, "NULL argument");
    mc2_engine* e = s->e;
    CUDA_CHECK(cudaSetDevice(e->device));
    if (s->k > 32) throw Mc2Error(MC2_ERR_LIMIT, "packed keys need k <= 32");
    if (s->plan.path == PATH_UNSET) {
        s->plan.enc = ENC_NT2;
        s->plan.path = PATH_SPARSE;
    } else if (s->plan.enc != ENC_NT2 || s->plan.path != PATH_SPARSE) {
        throw Mc2Error(MC2_ERR_INVALID, "the sample is not counting 2-bit packed keys");
    }
    s->n_chunks++;
    e->chunks++;
    if (n) {
        DBuf<u64> holder;
        const u64* d = keys;
        if (space != MC2_DEVICE) {
            holder.alloc(e, n);
            CUDA_CHECK(cudaMemcpyAsync(holder.p, keys, n * 8, cudaMemcpyHostToDevice, e->stream));
            e->h2d_bytes += n * 8;
            d = holder.p;
        }
        const u64 hash_max = std::min<u64>((u64)HC_MAX_NB1 * HC_NB2 * e->opt_hash_bucket_keys, e->opt_batch_symbols);
        KeySpan span{d, n, true};
        if (n <= hash_max) sparse_chunk_hash<ENC_NT2>(e, s, SymView{nullptr, 0}, nullptr, &span);
        else if (!sparse_chunk_hash_big(e, s, std::vector<PackedView>(), &span, hash_max))
            throw Mc2Error(MC2_ERR_LIMIT, "add_keys: the keys do not fit in free device memory");
        CUDA_CHECK(cudaStreamSynchronize(e->stream));
    }
    API_END
}

int mc2_count_exceptions(mc2_engine* e, const void* text, uint64_t nbytes, int space, int k, mc2_table** out) {
    API_BEGIN
    check_count_args(e, text, nbytes, k);
    if (!out) throw Mc2Error(MC2_ERR_INVALID, "out is NULL");
    CUDA_CHECK(cudaSetDevice(e->device));
    mc2_sample s;
    s.e = e;
    s.k = k;
    s.c = 1;
    s.plan.enc = ENC_NT2;
    s.plan.path = k <= 32 ? PATH_SPARSE : PATH_WIDE;
    if (nbytes) {
        DBuf<u8> holder;
        const u8* d = to_device(e, text, nbytes, space, holder);
        Parsed ps;
        parse_text(e, d, nbytes, 0, ps);
        SymView v{ps.sym.p, ps.nsym};
        const u64 n_slow = ps.stats.n_ascii - ps.stats.n_acgt;
        if (ps.nsym && k > 32) wide_chunk<ENC_BYTE, 1>(e, &s, v, ~0ull);
        else if (ps.nsym && n_slow) wide_chunk<ENC_NT2, 0>(e, &s, v, std::max<u64>(1024, n_slow * (u64)k));
    }
    *out = sample_finish(&s);
    API_END
}

// ---- merge_tsv (sample x k-mer matrix) ------------------------------------------------------------------------------
struct mc2_matrix {
    mc2_engine* e = nullptr;
    int k = 0;
    u32 samples = 0;
    u64 rows = 0;
    DBuf<u8> kmers;        // rows * k bytes, sorted
    DBuf<u64> counts;      // rows * samples, row-major
};

int mc2_table_from_tsv(mc2_engine* e, const void* text, uint64_t nbytes, int space, mc2_table** out) {
    API_BEGIN
    if (!e || !out || (nbytes && !text)) throw Mc2Error(MC2_ERR_INVALID, "NULL argument");
    CUDA_CHECK(cudaSetDevice(e->device));
    std::unique_ptr<mc2_table> t(new mc2_table);
    t->e = e;
    t->enc = ENC_BYTE;
    DBuf<u8> holder;
    const u8* d = to_device(e, text, nbytes, space, holder);
    // the text must end with a newline for the row scan: work on a copy with one appended when it does not
    DBuf<u8> padded;
    u64 n = nbytes;
    if (nbytes) {
        u8 last = 0;
        CUDA_CHECK(cudaMemcpyAsync(&last, d + nbytes - 1, 1, cudaMemcpyDeviceToHost, e->stream));
        CUDA_CHECK(cudaStreamSynchronize(e->stream));
        if (last != '\n') {
            padded.alloc(e, nbytes + 1);
            CUDA_CHECK(cudaMemcpyAsync(padded.p, d, nbytes, cudaMemcpyDeviceToDevice, e->stream));
            CUDA_CHECK(cudaMemsetAsync(padded.p + nbytes, '\n', 1, e->stream));
            d = padded.p;
            n = nbytes + 1;
        }
    }
    const u64 ntiles = div_up(std::max<u64>(n, 1), MG_TILE);
    DBuf<u32> tc(e, ntiles);
    DBuf<u64> to(e, ntiles);
    LAUNCH(e, mg_newline_count_kernel, (unsigned)ntiles, MG_THREADS, 0, d, n, tc.p);
    const u64 nlines = offsets_from_counts(e, tc.p, to.p, ntiles);
    if (nlines >= 2) {
        DBuf<u64> nl(e, nlines);
        LAUNCH(e, mg_newline_write_kernel, (unsigned)ntiles, MG_THREADS, 0, d, n, (const u64*)to.p, nl.p);
        // k = length of the first row's first column
        u64 ends[2];
        d2h(e, ends, (const u64*)nl.p, 2);
        const u64 a = ends[0] + 1, len1 = ends[1] - a;
        std::vector<u8> first(std::min<u64>(len1, 4096));
        d2h(e, first.data(), d + a, first.size());
        size_t tab = 0;
        while (tab < first.size() && first[tab] != '\t') ++tab;
        if (tab == 0 || tab >= first.size()) throw Mc2Error(MC2_ERR_INVALID, "TSV: the first row has no <k-mer>\\t<count>");
        if (tab > 128) throw Mc2Error(MC2_ERR_LIMIT, "TSV: k > 128");
        t->k = (int)tab;
        const u64 nrows = nlines - 1;
        t->wide.n = nrows;
        t->wide.sorted = false;
        t->wide.rows.alloc(e, nrows * (u64)t->k);
        t->wide.counts.alloc(e, nrows);
        DBuf<ull> bad(e, 1);
        bad.zero();
        LAUNCH(e, mg_parse_rows_kernel, (unsigned)div_up(nrows, 256), 256, 0, d, (const u64*)nl.p, nrows, t->k, t->wide.rows.p, t->wide.counts.p, bad.p);
        if (read_scalar<ull>(e, bad.p)) throw Mc2Error(MC2_ERR_INVALID, "TSV: a row is not <k-mer of the first row's length>\\t<decimal count>");
    }
    CUDA_CHECK(cudaStreamSynchronize(e->stream));
    *out = t.release();
    API_END
}

// all rows of a table as k literal bytes + counts (appended to rows/counts at row offset `at`)
static void table_text_rows(mc2_engine* e, mc2_table* t, u8* rows, u64* counts) {
    const int k = t->k;
    const u64 nf = t->fast.n, nw = t->wide.n;
    if (nf) {
        const int kind = t->key_kind == KEY_DENSE_AA ? TSV_DENSE_AA : t->enc == ENC_NT2 ? TSV_NT2 : t->enc == ENC_AA5 ? TSV_AA5 : TSV_BYTE;
        LAUNCH(e, mg_decode_rows_kernel, (unsigned)div_up(nf, 256), 256, 0, (const u64*)t->fast.keys.p, nf, k, kind, rows);
        CUDA_CHECK(cudaMemcpyAsync(counts, t->fast.counts.p, nf * 8, cudaMemcpyDeviceToDevice, e->stream));
    }
    if (nw) {
        CUDA_CHECK(cudaMemcpyAsync(rows + nf * (u64)k, t->wide.rows.p, nw * (u64)k, cudaMemcpyDeviceToDevice, e->stream));
        CUDA_CHECK(cudaMemcpyAsync(counts + nf, t->wide.counts.p, nw * 8, cudaMemcpyDeviceToDevice, e->stream));
    }
}

int mc2_merge_tables(mc2_engine* e, mc2_table* const* tables, uint32_t n, mc2_matrix** out) {
    API_BEGIN
    if (!e || !out || (n && !tables)) throw Mc2Error(MC2_ERR_INVALID, "NULL argument");
    CUDA_CHECK(cudaSetDevice(e->device));
    std::unique_ptr<mc2_matrix> m(new mc2_matrix);
    m->e = e;
    m->samples = n;
    int k = 0;
    u64 total = 0;
    for (u32 i = 0; i < n; ++i) {
        if (!tables[i]) throw Mc2Error(MC2_ERR_INVALID, "NULL table");
        const u64 r = tables[i]->fast.n + tables[i]->wide.n;
        if (!r) continue;
        if (k && tables[i]->k != k) throw Mc2Error(MC2_ERR_INVALID, "merge: the tables hold k-mers of different lengths");
        k = tables[i]->k;
        total += r;
    }
    m->k = k;
    if (total) {
        DBuf<u8> rows(e, total * (u64)k);
        DBuf<u64> counts(e, total), pos(e, total);
        std::vector<u64> at(n + 1, 0);
        for (u32 i = 0; i < n; ++i) {
            const u64 r = tables[i]->fast.n + tables[i]->wide.n;
            if (r) table_text_rows(e, tables[i], rows.p + at[i] * k, counts.p + at[i]);
            at[i + 1] = at[i] + r;
        }
        LAUNCH(e, wide_row_positions_kernel, (unsigned)div_up(total, 256), 256, 0, pos.p, total, k);
        WidePart uni;
        wide_reduce(e, rows.p, pos.p, nullptr, total, k, 1, uni);                 // sorted unique k-mers of all samples
        m->rows = uni.n;
        m->kmers = std::move(uni.rows);
        m->counts.alloc(e, uni.n * (u64)n);
        m->counts.zero();
        for (u32 i = 0; i < n; ++i) {
            const u64 r = at[i + 1] - at[i];
            if (r) LAUNCH(e, mg_fill_kernel, (unsigned)div_up(r, 256), 256, 0, (const u8*)m->kmers.p, m->rows, k, (const u8*)rows.p + at[i] * k,
                          (const u64*)counts.p + at[i], r, m->counts.p, n, i);
        }
    }
    CUDA_CHECK(cudaStreamSynchronize(e->stream));
    *out = m.release();
    API_END
}

uint64_t mc2_matrix_rows(const mc2_matrix* m) { return m ? m->rows : 0; }
int mc2_matrix_k(const mc2_matrix* m) { return m ? m->k : 0; }

int mc2_matrix_export(mc2_matrix* m, char* kmers, uint64_t* counts) {
    API_BEGIN
    if (!m) throw Mc2Error(MC2_ERR_INVALID, "NULL argument");
    if (m->rows && (!kmers || !counts)) throw Mc2Error(MC2_ERR_INVALID, "NULL buffer");
    CUDA_CHECK(cudaSetDevice(m->e->device));
    d2h(m->e, (u8*)kmers, (const u8*)m->kmers.p, m->rows * (u64)m->k);
    d2h(m->e, (u64*)counts, (const u64*)m->counts.p, m->rows * (u64)m->samples);
    API_END
}

// cells of `nrows` x `ncols` formatted on the device and streamed into f
static void matrix_emit(mc2_matrix* m, FILE* f, const u64* vals, u64 stride_r, u64 stride_c, u64 nrows, u64 ncols, u32 label_len, const u8* labels,
                        bool* ok) {
    mc2_engine* e = m->e;
    const u64 cells = nrows * ncols;
    if (!cells) return;
    DBuf<u32> len(e, cells);
    DBuf<u64> off(e, cells);
    DBuf<ull> total(e, 1);
    LAUNCH(e, mg_cell_len_kernel, (unsigned)div_up(cells, 256), 256, 0, vals, stride_r, stride_c, nrows, ncols, label_len, len.p);
    dev_exclusive_scan<u32, u64>(e, len.p, off.p, cells, total.p);
    const u64 nbytes = (u64)read_scalar<ull>(e, total.p);
    DBuf<u8> body(e, nbytes);
    LAUNCH(e, mg_cell_write_kernel, (unsigned)div_up(cells, 256), 256, 0, vals, stride_r, stride_c, nrows, ncols, label_len, labels,
           (const u64*)off.p, body.p);
    download_pipelined(e, (const u8*)body.p, nbytes, [&](const u8* src, u64, u64 n) { *ok = *ok && fwrite(src, 1, n, f) == n; });
}

int mc2_matrix_write_tsv(mc2_matrix* m, const char* path, const char* corner, const char* const* names, int transposed) {
    API_BEGIN
    if (!m || !path || !corner || (m->samples && !names)) throw Mc2Error(MC2_ERR_INVALID, "NULL argument");
    mc2_engine* e = m->e;
    CUDA_CHECK(cudaSetDevice(e->device));
    FILE* f = fopen(path, "wb");
    if (!f) throw Mc2Error(MC2_ERR_IO, std::string("cannot open ") + path);
    bool ok = true;
    std::string head = corner;
    if (!transposed) {
        // lib/mercat2_report.py:118: print(header, '\t'.join(names), sep='\t'); then one line per k-mer of the sorted union
        for (u32 s = 0; s < m->samples; ++s) { head += '\t'; head += names[s]; }
        head += '\n';
        ok = fwrite(head.data(), 1, head.size(), f) == head.size();
        matrix_emit(m, f, m->counts.p, m->samples, 1, m->rows, m->samples, (u32)m->k, m->kmers.p, &ok);
    } else {
        // lib/mercat2_report.py:176-193: 'sample' + every k-mer as a column (the reference's column order is the
        // iteration order of a Python set; here: sorted), then one line per sample
        std::vector<u8> km(m->rows * (u64)m->k);
        d2h(e, km.data(), (const u8*)m->kmers.p, km.size());
        for (u64 u = 0; u < m->rows; ++u) { head += '\t'; head.append((const char*)&km[u * m->k], m->k); }
        head += '\n';
        ok = fwrite(head.data(), 1, head.size(), f) == head.size();
        for (u32 s = 0; s < m->samples && ok; ++s) {
            ok = fwrite(names[s], 1, strlen(names[s]), f) == strlen(names[s]);
            if (m->rows) matrix_emit(m, f, m->counts.p + s, 0, m->samples, 1, m->rows, 0, nullptr, &ok);
            else ok = ok && fputc('\n', f) != EOF;
        }
    }
    ok = (fclose(f) == 0) && ok;
    if (!ok) throw Mc2Error(MC2_ERR_IO, std::string("short write to ") + path);
    API_END
}

void mc2_matrix_free(mc2_matrix* m) {
    if (!m) return;
    cudaSetDevice(m->e->device);
    delete m;
}

int mc2_sample_finish(mc2_sample* s, mc2_table** out) {
    API_BEGIN
    if (!s || !out) throw Mc2Error(MC2_ERR_INVALID, "NULL argument");
    std::unique_ptr<mc2_sample> guard(s);
    CUDA_CHECK(cudaSetDevice(s->e->device));
    *out = sample_finish(s);
    API_END
}

void mc2_sample_abort(mc2_sample* s) {
    if (!s) return;
    cudaSetDevice(s->e->device);
    delete s;
}

int mc2_count_sample(mc2_engine* e, const void* text, uint64_t nbytes, int space, int k, int64_t min_count,
                     uint64_t chunk_bytes, mc2_table** out, uint64_t* n_chunks, uint64_t* piece_offsets,
                     uint64_t piece_capacity) {
    API_BEGIN
    check_count_args(e, text, nbytes, k);
    if (!out) throw Mc2Error(MC2_ERR_INVALID, "out is NULL");
    CUDA_CHECK(cudaSetDevice(e->device));
    std::unique_ptr<mc2_sample> s(new mc2_sample);
    s->e = e;
    s->k = k;
    s->c = min_count < 1 ? 1 : (u64)min_count;
    CUDA_CHECK(cudaEventRecord(e->ev0, e->stream));
    u64 nc = 0;
    std::vector<u64> bounds;
    sample_add(s.get(), text, nbytes, space, chunk_bytes, &nc, &bounds);
    mc2_table* t = sample_finish(s.get());
    CUDA_CHECK(cudaEventRecord(e->ev1, e->stream));
    CUDA_CHECK(cudaEventSynchronize(e->ev1));
    float ms = 0;
    CUDA_CHECK(cudaEventElapsedTime(&ms, e->ev0, e->ev1));
    e->device_us = ms * 1000.0;
    e->resolve_profile();
    if (n_chunks) *n_chunks = nc;
    if (piece_offsets)
        for (u64 i = 0; i < bounds.size() && i < piece_capacity; ++i) piece_offsets[i] = bounds[i];
    *out = t;
    API_END
}

int mc2_count_text(mc2_engine* e, const void* text, uint64_t nbytes, int space, int k, int64_t min_count, mc2_table** out) {
    return mc2_count_sample(e, text, nbytes, space, k, min_count, 0, out, nullptr, nullptr, 0);
}

int mc2_count_symbols(mc2_engine* e, const void* symbols, uint64_t nbytes, int space, int k, int64_t min_count, mc2_table** out) {
    API_BEGIN
    check_count_args(e, symbols, nbytes, k);
    if (!out) throw Mc2Error(MC2_ERR_INVALID, "out is NULL");
    CUDA_CHECK(cudaSetDevice(e->device));
    std::unique_ptr<mc2_sample> s(new mc2_sample);
    s->e = e;
    s->k = k;
    s->c = min_count < 1 ? 1 : (u64)min_count;
    DBuf<u8> holder;
    const u8* d = to_device(e, symbols, nbytes, space, holder);
    count_chunk(e, s.get(), d, nbytes, true);
    *out = sample_finish(s.get());
    API_END
}

int mc2_chunk_offsets(mc2_engine* e, const void* text, uint64_t nbytes, int space, uint64_t chunk_bytes,
                      uint64_t* piece_offsets, uint64_t piece_capacity, uint64_t* n_pieces) {
    API_BEGIN
    if (!e || (!text && nbytes)) throw Mc2Error(MC2_ERR_INVALID, "NULL argument");
    CUDA_CHECK(cudaSetDevice(e->device));
    DBuf<u8> holder;
    const u8* d = to_device(e, text, nbytes, space, holder);
    const std::vector<u64> bounds = chunk_bounds(e, d, nbytes, chunk_bytes);
    CUDA_CHECK(cudaStreamSynchronize(e->stream));
    if (n_pieces) *n_pieces = bounds.size();
    if (piece_offsets)
        for (u64 i = 0; i < bounds.size() && i < piece_capacity; ++i) piece_offsets[i] = bounds[i];
    API_END
}

uint64_t mc2_table_rows(const mc2_table* t) { return t ? t->fast.n + t->wide.n : 0; }
int mc2_table_k(const mc2_table* t) { return t ? t->k : 0; }
uint64_t mc2_table_total(const mc2_table* t) {
    if (!t) return 0;
    try { ensure_host(const_cast<mc2_table*>(t)); } catch (...) { return 0; }
    return t->total;
}

int mc2_table_export(mc2_table* t, char* kmers, uint64_t* counts) {
    API_BEGIN
    if (!t) throw Mc2Error(MC2_ERR_INVALID, "table is NULL");
    CUDA_CHECK(cudaSetDevice(t->e->device));
    ensure_host(t);
    if (kmers && !t->kmers.empty()) memcpy(kmers, t->kmers.data(), t->kmers.size());
    if (counts && !t->counts.empty()) memcpy(counts, t->counts.data(), t->counts.size() * 8);
    API_END
}

int mc2_table_tsv(mc2_table* t, const char* basename, char* buf, uint64_t cap, uint64_t* size) {
    API_BEGIN
    if (!t || !basename) throw Mc2Error(MC2_ERR_INVALID, "NULL argument");
    CUDA_CHECK(cudaSetDevice(t->e->device));
    const std::string head = tsv_header(basename);
    if (!t->tsv_ready) {                                  // (a size query followed by the copy-out formats once)
        t->tsv_bytes = tsv_body_device(t, t->tsv_body);
        t->tsv_ready = true;
    }
    const u64 nbytes = t->tsv_bytes;
    if (size) *size = head.size() + nbytes;
    if (buf) {
        if (cap < head.size() + nbytes) throw Mc2Error(MC2_ERR_INVALID, "buffer too small");
        memcpy(buf, head.data(), head.size());
        u8* dst = (u8*)buf + head.size();
        download_pipelined(t->e, (const u8*)t->tsv_body.p, nbytes, [&](const u8* src, u64 o, u64 m) { memcpy(dst + o, src, m); });
        t->tsv_body.release();
        t->tsv_ready = false;
    }
    API_END
}
int mc2_table_write_tsv(mc2_table* t, const char* path, const char* basename) {
    try {
        if (!t || !path || !basename) throw Mc2Error(MC2_ERR_INVALID, "NULL argument");
        if (mc2_table_rows(t) == 0) return 1;
        CUDA_CHECK(cudaSetDevice(t->e->device));
        const std::string head = tsv_header(basename);
        DBuf<u8> body;
        const u64 nbytes = tsv_body_device(t, body);
        FILE* f = fopen(path, "wb");
        if (!f) throw Mc2Error(MC2_ERR_IO, std::string("cannot open ") + path);
        bool ok = fwrite(head.data(), 1, head.size(), f) == head.size();
        download_pipelined(t->e, (const u8*)body.p, nbytes, [&](const u8* src, u64, u64 m) { ok = ok && fwrite(src, 1, m, f) == m; });
        ok = (fclose(f) == 0) && ok;
        if (!ok) throw Mc2Error(MC2_ERR_IO, std::string("short write to ") + path);
    } catch (const Mc2Error& err) { g_err = err.what(); return err.code; }
    catch (const std::exception& err) { g_err = err.what(); return MC2_ERR_INVALID; }
    return MC2_OK;
}

void mc2_table_free(mc2_table* t) {
    if (!t) return;
    cudaSetDevice(t->e->device);
    delete t;
}

}  // extern "C"

#include "metrics_api.inl"
