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
#include "rangecount.cuh"
#include "rowmerge.cuh"
#include "fastnt.cuh"
#include "metrics.cuh"
#include "tsv.cuh"
#include "merge.cuh"
#include "textprep.cuh"

static thread_local std::string g_err;

// =====================================================================================================
// engine
// =====================================================================================================
struct RangeWork;
// Rows of a count delivered straight to caller-owned host memory (mc2_count_text_rows): the groups of a very large chunk
// stream their finished rows out on the copy stream while later groups are still being counted.
struct HostRowSink {
    bool active = false, failed = false, complete = false;
    void* rows = nullptr;                  // 16 bytes per row: key, count
    u64* keys64 = nullptr;                 // split delivery (mc2_count_text_rows_split): 8 + 4 bytes per row in two arrays
    u32* counts32 = nullptr;
    bool too_big = false;                  // split delivery: some count needs more than 32 bits
    u64 capacity = 0, delivered = 0;
};
struct mc2_engine {
    int device = 0;
    RangeWork* work = nullptr;             // device workspace of the range-partition path, grown on demand and reused by every chunk
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
    int opt_group_sync = 0;                // very large chunks: 1 = one host round trip per level-0 group (the min_count 1 path) even for min_count >= 2
    HostRowSink host_rows;
    unsigned long long* pin_groups = nullptr;   // pinned: per-group row counters of the streaming download
    u64 l0_fit_total = 0, l0_fit_extra = 0; // largest level-0 request that passed the memory check (see level0_partition)
    int opt_grid_waves = 0;                // persistent kernels of the range path: CTAs launched = this x the resident slots (see sparse_chunk_range)
    int opt_bucket_growth = 1;             // duplicate-rich data: size sub-buckets by the measured keys per distinct key (0 = fixed 1.25 x)
    int opt_row_merge = 0;                 // sums of (key, count) row sets: 0 = sort + segmented sum, 1 = range partition + shared-memory sums
    int opt_count_mode = -1;               // counting kernel: -1 auto, 0 every key into the table, 1 bitmap pre-filter (min_count >= 2 only)
    int opt_sparse_algo = 0;               // 0 auto (range partition + shared-memory tables), 1 radix sort, 2 range partition
    u64 opt_hash_bucket_keys = 3500;       // target keys per shared-memory table
    bool range_attrs_set = false;          // kernel attributes (dynamic shared memory opt-in) are per device: set once per engine
    bool dense_attrs_set[3] = {false, false, false};
    // stats
    u64 launches = 0, h2d_bytes = 0, d2h_bytes = 0, chunks = 0, ovf_buckets = 0, row_merges = 0;
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

// profile 1/2: CUDA events around every kernel; 3: only around the three long kernels of the sparse path (the events
// themselves cost ~6 % of a step when every launch carries a pair)
static inline bool prof_on(const mc2_engine* e, const char* name) {
    if (!e->profile) return false;
    if (e->profile != 3) return true;
    return !strncmp(name, "hc_scatter2", 11) || !strncmp(name, "fn_scatter1", 11) || !strncmp(name, "hk_scatter1", 11) || !strncmp(name, "rc_count", 8);
}

#define LAUNCHN(e, name, kern, grid, block, smem, ...)                            \
    do {                                                                          \
        cudaEvent_t _pa = nullptr, _pb = nullptr;                                 \
        if (prof_on((e), name)) {                                                 \
            _pa = (e)->get_event();                                               \
            _pb = (e)->get_event();                                               \
            cudaEventRecord(_pa, (e)->stream);                                    \
        }                                                                         \
        kern<<<(grid), (block), (smem), (e)->stream>>>(__VA_ARGS__);              \
        if (_pa) {                                                                \
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

// grow-only device buffer: get(n) returns room for n elements, reallocating (with head room) only when it must
template <typename T>
struct WBuf {
    DBuf<T> b;
    T* get(mc2_engine* e, u64 n) {
        if (!b.p || b.n < n) b.alloc(e, n + n / 16 + 64);
        return b.p;
    }
};
// Workspace of one range-partitioned chunk / group (rangecount.cuh).  Chunks of a sample and groups of a very large
// chunk follow each other on one stream and differ a little in size; allocating their buffers afresh each time made
// the stream-ordered pool grow and fragment (a freed 536 MB block does not serve a 540 MB request) and cost more host
// time than the kernels took.  One set of buffers, sized for the largest request so far, serves them all.
struct RangeWork {
    WBuf<u64> keys1, keys2, row_off;
    WBuf<u64> keys0;                       // level-0 key array of a very large chunk (8 B per window: tens of GB).  Kept between
                                           // calls -- re-creating it cost 17 ms when the pool still held the memory and 0.2-0.9 s
                                           // when it did not (e.g. after NCCL had sent from it) -- and released by mc2_engine_trim
    bool keys0_busy = false;               // an mc2_keys object is holding it
    WBuf<u32> ghist, sub_base, cur1, cur2, tile_pref, ovf_list, rows, shist;
    WBuf<u8> slots1;                       // RcRow slots when min_count == 1 (otherwise the slots overlay keys1)
    WBuf<u16> lut[2];                      // level-0 plan and group / chunk plan are alive at the same time
    WBuf<uint2> l1[2];
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
    if (n <= SCAN_SMALL_MAX) {
        auto small = scan_small_kernel<T, O>;
        LAUNCHN(e, "scan_small_kernel", small, 1, SCAN1_THREADS, 0, in, out, (u32)n, total_dev);
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
    bool added = false;    // rows handed in by mc2_sample_add_rows: may duplicate k-mers the packed rows hold (see fold_added_rows)
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
    DBuf<u32> codes;                        // 2-bit codes, followed in the same allocation by the bad bits (one memset)
    u32* bad = nullptr;
    void alloc_packed(mc2_engine* e, u64 nsym) {
        const u64 wc = (div_up(nsym, 16) + 4 + 3) & ~3ull, wb = div_up(nsym, 32) + 4;
        codes.alloc(e, wc + wb);
        codes.zero();
        bad = codes.p + wc;
    }
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
    double bucket_scale = 1.0;             // shrinks when many sub-buckets overflow their table
    double dup_ratio = 0;                  // duplicate-rich data: keys per distinct key, measured on an earlier chunk / group (0 = not known)
    bool dup_rich = false;                 // most keys repeat (seen on an earlier chunk / group): count without the bitmap pre-filter
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

#include "host_reduce.inl"
#include "host_count.inl"
#include "host_ingest.inl"
#include "host_export.inl"

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
    if (e->pin_groups) cudaFreeHost(e->pin_groups);
    if (e->ev0) cudaEventDestroy(e->ev0);
    if (e->ev1) cudaEventDestroy(e->ev1);
    delete e->work;
    if (e->stream) cudaStreamDestroy(e->stream);
    if (e->copy_stream) cudaStreamDestroy(e->copy_stream);
    delete e;
}

int mc2_engine_trim(mc2_engine* e) {
    API_BEGIN
    if (!e) throw Mc2Error(MC2_ERR_INVALID, "engine is NULL");
    CUDA_CHECK(cudaSetDevice(e->device));
    CUDA_CHECK(cudaStreamSynchronize(e->stream));
    if (e->work && e->work->keys0_busy) throw Mc2Error(MC2_ERR_INVALID, "a key partition (mc2_keys) still uses the workspace");
    delete e->work;
    e->work = nullptr;
    e->l0_fit_total = e->l0_fit_extra = 0;
    cudaMemPool_t pool;
    CUDA_CHECK(cudaDeviceGetDefaultMemPool(&pool, e->device));
    CUDA_CHECK(cudaStreamSynchronize(e->stream));
    CUDA_CHECK(cudaMemPoolTrimTo(pool, 0));
    API_END
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
    else if (n == "count_mode") e->opt_count_mode = (int)value;
    else if (n == "row_merge") e->opt_row_merge = (int)value;
    else if (n == "bucket_growth") e->opt_bucket_growth = (int)value;
    else if (n == "grid_waves") e->opt_grid_waves = (int)value;
    else if (n == "group_sync") e->opt_group_sync = (int)value;
    else if (n == "big_chunks") e->opt_big_chunks = (int)value;
    else if (n == "parse_single") e->opt_parse_single = (int)value;
    else if (n == "prefetch_pass") e->opt_prefetch_pass = (int)value;
    else if (n == "file_piece_bytes") e->opt_file_piece = (u64)std::max<int64_t>(4096, value);
    else if (n == "span_bytes") e->opt_span_bytes = value < 4096 ? 4096 : (value > (3ull << 30) ? (3ull << 30) : (u64)value);
    else if (n == "hash_bucket_keys") e->opt_hash_bucket_keys = (u64)std::max<int64_t>(value, 16);
    else if (n == "profile") {          // 0 off, 1 on, 2 on + totals cleared, 3 = like 2 but only the three long kernels carry events
        e->resolve_profile();
        e->profile = value == 3 ? 3 : (value ? 1 : 0);
        if (value >= 2) e->prof_total.clear();
    }
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
    if (n == "row_merges") return (int64_t)e->row_merges;       // (key, count) row sets summed on the range path
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
        part.added = true;
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

int mc2_table_export_packed(mc2_table* t, uint64_t* keys, uint64_t* counts, uint64_t capacity, uint64_t* rows) {
    API_BEGIN
    if (!t || !rows) throw Mc2Error(MC2_ERR_INVALID, "NULL argument");
    const u64 n = t->fast.n;
    if (n && (!keys || !counts)) throw Mc2Error(MC2_ERR_INVALID, "NULL buffer");
    if (n > capacity) throw Mc2Error(MC2_ERR_INVALID, "buffer too small");
    mc2_engine* e = t->e;
    CUDA_CHECK(cudaSetDevice(e->device));
    if (n) {
        CUDA_CHECK(cudaMemcpyAsync(keys, t->fast.keys.p, n * 8, cudaMemcpyDeviceToHost, e->stream));
        CUDA_CHECK(cudaMemcpyAsync(counts, t->fast.counts.p, n * 8, cudaMemcpyDeviceToHost, e->stream));
        CUDA_CHECK(cudaStreamSynchronize(e->stream));
        e->d2h_bytes += n * 16;
    }
    *rows = n;
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
    ~mc2_keys() { if (l0.borrowed && e && e->work) e->work->keys0_busy = false; }
    mc2_engine* e = nullptr;
    int k = 0;
    u32 groups = 0;
    // open state: the packed text and its sampled prefix histogram (mc2_keys_open), until mc2_keys_partition ran
    DBuf<u8> holder;
    std::vector<FnSpan> spans;
    std::vector<PackedView> pvs;
    DBuf<u32> shist;
    bool partitioned = false;
    Level0 l0;
    u64 exception_symbols = 0;
    u64 windows = 0;
};

int mc2_keys_open(mc2_engine* e, const void* text, uint64_t nbytes, int space, int k, mc2_keys** out) {
    API_BEGIN
    check_count_args(e, text, nbytes, k);
    if (!out) throw Mc2Error(MC2_ERR_INVALID, "out is NULL");
    if (k > 32) throw Mc2Error(MC2_ERR_LIMIT, "key partition needs k <= 32 (2-bit packed keys)");
    CUDA_CHECK(cudaSetDevice(e->device));
    range_kernel_attrs(e);
    std::unique_ptr<mc2_keys> ks(new mc2_keys);
    ks->e = e;
    ks->k = k;
    ks->shist.alloc(e, RP_LUT);
    ks->shist.zero();
    if (nbytes) {
        const u8* d = to_device(e, text, nbytes, space, ks->holder);
        std::vector<u64> cuts;
        if (!fn_span_cuts(e, d, nbytes, cuts)) throw Mc2Error(MC2_ERR_LIMIT, "key partition: text cannot be cut into spans at header lines");
        DBuf<FnStats> st(e, 1);
        st.zero();
        ks->spans.resize(cuts.size() - 1);
        for (size_t i = 0; i + 1 < cuts.size(); ++i) {
            FnSpan& sp = ks->spans[i];
            sp.text = d + cuts[i];
            sp.len = cuts[i + 1] - cuts[i];
            const bool single = e->opt_parse_single != 0;
            const FnStats fs = single ? fn_single_pass(e, sp, false, st) : fn_count_pass(e, sp, false, st);
            if (fs.complex) throw Mc2Error(MC2_ERR_LIMIT, "key partition: text is not plain FASTA (whitespace, '*' or non-ASCII bytes in sequence lines)");
            if (!sp.nsym) continue;
            if (!single) fn_write_pass(e, sp, st);
            ks->pvs.push_back(PackedView{sp.codes.p, sp.bad, sp.nsym});
            ks->windows += sp.nsym;
        }
        const FnStats fs2 = read_scalar<FnStats>(e, st.p);
        ks->exception_symbols = fs2.packed2 >> 32;
        ks->holder.release();                                    // the packed codes are all that is needed from here on
        // sampled prefix histogram over the whole key space (the caller may sum it over ranks before partitioning)
        const u64 stride = std::max<u64>(1, ks->windows / RP_SAMPLE_WINDOWS);
        for (auto& pv : ks->pvs) {
            const u64 nwords = div_up(pv.n, 16);
            const u64 grid = std::min<u64>(div_up(div_up(nwords, stride), FN_HIST_THREADS), (u64)e->num_sms * 2);
            LAUNCH(e, rp_sample_packed_kernel, (unsigned)std::max<u64>(grid, 1), FN_HIST_THREADS, 0, pv, k, stride, 0u, 32u - RP_LUT_LOG2, ks->shist.p);
        }
    }
    CUDA_CHECK(cudaStreamSynchronize(e->stream));
    *out = ks.release();
    API_END
}

int mc2_keys_sample(mc2_keys* ks, uint32_t** hist, uint64_t* entries) {
    API_BEGIN
    if (!ks || !hist || !entries) throw Mc2Error(MC2_ERR_INVALID, "NULL argument");
    *hist = ks->shist.p;
    *entries = RP_LUT;
    API_END
}

int mc2_keys_partition(mc2_keys* ks, uint32_t groups) {
    API_BEGIN
    if (!ks) throw Mc2Error(MC2_ERR_INVALID, "NULL argument");
    if (ks->partitioned) throw Mc2Error(MC2_ERR_INVALID, "the keys are already partitioned");
    if (groups < 1 || groups > HC_MAX_NB1) throw Mc2Error(MC2_ERR_INVALID, "groups must be in [1, 384]");
    mc2_engine* e = ks->e;
    CUDA_CHECK(cudaSetDevice(e->device));
    ks->groups = groups;
    ks->l0.gbase.assign(groups + 1, 0);
    ks->l0.bounds.assign(groups + 1, 1ull << 32);
    ks->l0.bounds[0] = 0;
    if (!level0_partition(e, ks->k, ks->pvs, nullptr, groups, (1ull << 32) - 1, 0, ks->l0, ks->shist.p, true))
        throw Mc2Error(MC2_ERR_LIMIT, "key partition: the keys do not fit in free device memory");
    ks->pvs.clear();
    ks->spans.clear();
    ks->partitioned = true;
    CUDA_CHECK(cudaStreamSynchronize(e->stream));
    API_END
}

int mc2_partition_keys(mc2_engine* e, const void* text, uint64_t nbytes, int space, int k, uint32_t groups, mc2_keys** out) {
    if (!out) { g_err = "out is NULL"; return MC2_ERR_INVALID; }
    if (groups < 1 || groups > HC_MAX_NB1) { g_err = "groups must be in [1, 384]"; return MC2_ERR_INVALID; }
    mc2_keys* ks = nullptr;
    int st = mc2_keys_open(e, text, nbytes, space, k, &ks);
    if (st < 0) return st;
    st = mc2_keys_partition(ks, groups);
    if (st < 0) { mc2_keys_free(ks); return st; }
    *out = ks;
    return MC2_OK;
}

int mc2_keys_info(mc2_keys* ks, const uint64_t** keys, uint64_t* sizes, uint64_t* total, uint64_t* exception_symbols, uint64_t* bounds) {
    API_BEGIN
    if (!ks) throw Mc2Error(MC2_ERR_INVALID, "NULL argument");
    if (!ks->partitioned) throw Mc2Error(MC2_ERR_INVALID, "mc2_keys_partition has not run");
    if (keys) *keys = (const uint64_t*)ks->l0.keys0;
    if (sizes)
        for (u32 g = 0; g < ks->groups; ++g) sizes[g] = ks->l0.gbase[g + 1] - ks->l0.gbase[g];
    if (bounds)
        for (u32 g = 0; g <= ks->groups; ++g) bounds[g] = ks->l0.bounds[g];
    if (total) *total = ks->l0.gbase[ks->groups];
    if (exception_symbols) *exception_symbols = ks->exception_symbols;
    API_END
}

void mc2_keys_free(mc2_keys* ks) {
    if (!ks) return;
    cudaSetDevice(ks->e->device);
    delete ks;
}

int mc2_sample_add_keys(mc2_sample* s, const uint64_t* keys, uint64_t n, int space, uint64_t prefix_lo, uint64_t prefix_hi) {
    API_BEGIN
    if (!s || (n && !keys)) throw Mc2Error(MC2_ERR_INVALID, "NULL argument");
    mc2_engine* e = s->e;
    CUDA_CHECK(cudaSetDevice(e->device));
    if (s->k > 32) throw Mc2Error(MC2_ERR_LIMIT, "packed keys need k <= 32");
    if (prefix_hi > (1ull << 32) || (prefix_hi && prefix_lo >= prefix_hi)) throw Mc2Error(MC2_ERR_INVALID, "bad prefix range");
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
        const u64 hash_max = range_batch_max(e, s);
        KeySpan span{d, n, prefix_hi ? prefix_lo : 0ull, prefix_hi ? prefix_hi : (1ull << 32)};
        if (n <= hash_max) sparse_chunk_range<ENC_NT2>(e, s, SymView{nullptr, 0}, nullptr, &span);
        else if (!sparse_chunk_big(e, s, std::vector<PackedView>(), &span, hash_max))
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

// merge_tsv exactly as the reference writes it (lib/mercat2_report.py:98-160), for tables still on the device.  The
// reference walks the per-sample TSVs with one cursor per file; the label of an output line is the smallest NEXT k-mer
// among the files that advanced on the previous line (all files at the start), a file prints its count and advances
// when its current k-mer is <= the label and prints 0 otherwise, and the walk ends when no file that advanced has a
// row left.  With identical k-mer sets this is the sorted union; with differing sets a pending smaller k-mer is
// printed under a larger label, labels repeat, and rows of files that never advance again are dropped -- downstream
// consumers of the reference's combined table see exactly that, so this entry point reproduces it byte for byte
// (the sorted union is mc2_merge_tables + mc2_matrix_write_tsv).  The cursor walk is inherently serial and runs on the
// host over the tables' exported rows; the device did the parse / count / sort that produced them.
int mc2_merge_tables_reference(mc2_engine* e, mc2_table* const* tables, uint32_t n, const char* path, const char* corner,
                               const char* const* names) {
    API_BEGIN
    if (!e || !path || !corner || (n && (!tables || !names))) throw Mc2Error(MC2_ERR_INVALID, "NULL argument");
    CUDA_CHECK(cudaSetDevice(e->device));
    for (u32 i = 0; i < n; ++i) {
        if (!tables[i]) throw Mc2Error(MC2_ERR_INVALID, "NULL table");
        ensure_host(tables[i]);
        if (tables[i]->counts.empty()) throw Mc2Error(MC2_ERR_INVALID, "merge: a table without rows (the reference fails on a TSV without rows)");
    }
    FILE* f = fopen(path, "wb");
    if (!f) throw Mc2Error(MC2_ERR_IO, std::string("cannot open ") + path);
    std::string buf = corner;
    for (u32 i = 0; i < n; ++i) { buf += '\t'; buf += names[i]; }
    buf += '\n';
    bool ok = true;
    struct Key { const char* p; size_t len; };
    auto less = [](const Key& a, const Key& b) {                   // Python str order of ASCII text
        const int d = memcmp(a.p, b.p, std::min(a.len, b.len));
        return d < 0 || (d == 0 && a.len < b.len);
    };
    auto key_of = [&](u32 i, u64 row) { return Key{tables[i]->kmers.data() + row * (u64)tables[i]->k, (size_t)tables[i]->k}; };
    std::vector<u64> cur(n, 0);
    bool have = false;
    Key label{nullptr, 0};
    for (u32 i = 0; i < n; ++i) {
        const Key k0 = key_of(i, 0);
        if (!have || less(k0, label)) { label = k0; have = true; }
    }
    char num[24];
    while (have) {
        buf.append(label.p, label.len);
        bool next_have = false;
        Key next{nullptr, 0};
        for (u32 i = 0; i < n; ++i) {
            const u64 rows = tables[i]->counts.size();
            if (cur[i] >= rows || less(label, key_of(i, cur[i]))) {
                buf += "\t0";
            } else {
                const int len = snprintf(num, sizeof num, "\t%llu", (ull)tables[i]->counts[cur[i]]);
                buf.append(num, len);
                if (++cur[i] < rows) {
                    const Key k1 = key_of(i, cur[i]);
                    if (!next_have || less(k1, next)) { next = k1; next_have = true; }
                }
            }
        }
        buf += '\n';
        if (buf.size() >= (8u << 20)) { ok = ok && fwrite(buf.data(), 1, buf.size(), f) == buf.size(); buf.clear(); }
        have = next_have;
        label = next;
    }
    ok = ok && fwrite(buf.data(), 1, buf.size(), f) == buf.size();
    ok = (fclose(f) == 0) && ok;
    if (!ok) throw Mc2Error(MC2_ERR_IO, std::string("short write to ") + path);
    API_END
}

int mc2_table_export_counts(mc2_table* t, uint64_t* counts, uint64_t* spectrum) {
    API_BEGIN
    if (!t) throw Mc2Error(MC2_ERR_INVALID, "NULL argument");
    mc2_engine* e = t->e;
    CUDA_CHECK(cudaSetDevice(e->device));
    const u64 nf = t->fast.n, nw = t->wide.n;
    if (counts) {
        ensure_host(t);                                        // rows in sorted k-mer order, packed and literal rows merged
        if (!t->counts.empty()) memcpy(counts, t->counts.data(), t->counts.size() * 8);
    }
    if (spectrum) {
        DBuf<ull> sp(e, MG_SPECTRUM_WORDS);
        sp.zero();
        if (nf) LAUNCH(e, mg_spectrum_kernel, (unsigned)std::min<u64>(div_up(nf, 256 * 8), (u64)e->num_sms * 4), 256, 0, (const u64*)t->fast.counts.p, nf, sp.p);
        if (nw) LAUNCH(e, mg_spectrum_kernel, (unsigned)std::min<u64>(div_up(nw, 256 * 8), (u64)e->num_sms * 4), 256, 0, (const u64*)t->wide.counts.p, nw, sp.p);
        d2h(e, (ull*)spectrum, (const ull*)sp.p, (u64)MG_SPECTRUM_WORDS);
    }
    API_END
}

int mc2_matrix_top_rows(mc2_matrix* m, uint32_t top, uint64_t* rows_out, uint32_t* found) {
    API_BEGIN
    if (!m || !rows_out || !found) throw Mc2Error(MC2_ERR_INVALID, "NULL argument");
    if (top > 32) throw Mc2Error(MC2_ERR_INVALID, "top must be <= 32");
    mc2_engine* e = m->e;
    CUDA_CHECK(cudaSetDevice(e->device));
    const u32 take = (u32)std::min<u64>(top, m->rows);
    *found = take;
    if (take) {
        DBuf<u64> sums(e, m->rows), out(e, take);
        LAUNCH(e, mg_row_sums_kernel, (unsigned)div_up(m->rows, 256), 256, 0, (const u64*)m->counts.p, m->rows, m->samples, sums.p);
        LAUNCH(e, mg_top_rows_kernel, 1, 1024, 0, (const u64*)sums.p, m->rows, take, out.p);
        d2h(e, (u64*)rows_out, (const u64*)out.p, (u64)take);
    }
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

// ---- text transforms ahead of the hot path (rows N1 / N2) ---------------------------------------------------------------
struct mc2_text {
    mc2_engine* e = nullptr;
    DBuf<u8> data;
    u64 nbytes = 0;
};

// '\n' positions of a device text (ascending); returns their number
static u64 newline_positions(mc2_engine* e, const u8* d, u64 n, DBuf<u64>& nl) {
    const u64 ntiles = div_up(std::max<u64>(n, 1), MG_TILE);
    DBuf<u32> tc(e, ntiles);
    DBuf<u64> to(e, ntiles);
    LAUNCH(e, mg_newline_count_kernel, (unsigned)ntiles, MG_THREADS, 0, d, n, tc.p);
    const u64 n_nl = offsets_from_counts(e, tc.p, to.p, ntiles);
    nl.alloc(e, std::max<u64>(n_nl, 1));
    if (n_nl) LAUNCH(e, mg_newline_write_kernel, (unsigned)ntiles, MG_THREADS, 0, d, n, (const u64*)to.p, nl.p);
    return n_nl;
}

int mc2_fastq_to_fasta(mc2_engine* e, const void* text, uint64_t nbytes, int space, mc2_text** out) {
    API_BEGIN
    if (!e || !out || (nbytes && !text)) throw Mc2Error(MC2_ERR_INVALID, "NULL argument");
    CUDA_CHECK(cudaSetDevice(e->device));
    std::unique_ptr<mc2_text> t(new mc2_text);
    t->e = e;
    if (nbytes) {
        DBuf<u8> holder;
        const u8* d = to_device(e, text, nbytes, space, holder);
        DBuf<u64> nl;
        const u64 n_nl = newline_positions(e, d, nbytes, nl);
        u8 last = 0;
        CUDA_CHECK(cudaMemcpyAsync(&last, d + nbytes - 1, 1, cudaMemcpyDeviceToHost, e->stream));
        CUDA_CHECK(cudaStreamSynchronize(e->stream));
        const u64 nlines = n_nl + (last != '\n' ? 1 : 0);
        DBuf<u32> len(e, nlines);
        DBuf<u64> off(e, nlines);
        DBuf<ull> total(e, 1);
        LAUNCH(e, fq_line_len_kernel, (unsigned)div_up(nlines, 256), 256, 0, d, (u64)nbytes, (const u64*)nl.p, n_nl, nlines, len.p);
        dev_exclusive_scan<u32, u64>(e, len.p, off.p, nlines, total.p);
        t->nbytes = (u64)read_scalar<ull>(e, total.p);
        t->data.alloc(e, t->nbytes + 16);
        if (t->nbytes)
            LAUNCH(e, fq_copy_kernel, (unsigned)div_up(nlines, 8), 256, 0, d, (u64)nbytes, (const u64*)nl.p, n_nl, nlines, (const u32*)len.p,
                   (const u64*)off.p, t->data.p);
        CUDA_CHECK(cudaStreamSynchronize(e->stream));
    }
    *out = t.release();
    API_END
}

int mc2_text_info(const mc2_text* t, const void** device_ptr, uint64_t* nbytes) {
    API_BEGIN
    if (!t) throw Mc2Error(MC2_ERR_INVALID, "NULL argument");
    if (device_ptr) *device_ptr = t->data.p;
    if (nbytes) *nbytes = t->nbytes;
    API_END
}

int mc2_text_export(mc2_text* t, void* host, uint64_t capacity) {
    API_BEGIN
    if (!t || (t->nbytes && !host)) throw Mc2Error(MC2_ERR_INVALID, "NULL argument");
    if (capacity < t->nbytes) throw Mc2Error(MC2_ERR_INVALID, "buffer too small");
    CUDA_CHECK(cudaSetDevice(t->e->device));
    d2h(t->e, (u8*)host, (const u8*)t->data.p, t->nbytes);
    API_END
}

void mc2_text_free(mc2_text* t) {
    if (!t) return;
    cudaSetDevice(t->e->device);
    delete t;
}

// ---- several small samples in ONE pass (BASELINE config 5: many proteomes, one table each) --------------------------
// Replaces the sample loop of bin/mercat2.py:411-448 (one countKmers task per file) for samples that are each ONE piece
// (smaller than the -s trigger): the texts are laid out behind each other (a header line between them, each starting
// on a 4 KiB parse tile), parsed once, and every window's key carries its sample index above the k-mer code, so a single
// run of the range path counts all (sample, k-mer) pairs, filters each pair with min_count -- exactly the per-file
// filter of lib/mercat2_kmers.py:73-78 -- and leaves the rows sorted by (sample, k-mer); the per-sample tables are
// slices of that array.  A 1.6 M-residue proteome alone cannot fill the GPU (its whole pass is launch latency);
// 64 of them together are one ordinary 100 MB piece.  Falls back to one pass per sample when the batch does not fit
// the scheme (windows outside the packed alphabet, k too large for the key, too many symbols).
static void count_one_text(mc2_engine* e, const u8* d, u64 nbytes, int k, u64 c, mc2_table** out) {
    mc2_sample s;
    s.e = e; s.k = k; s.c = c;
    if (nbytes) count_chunk(e, &s, d, nbytes);
    *out = sample_finish(&s);
}

__global__ void mask_keys_kernel(const u64* __restrict__ in, u64 n, u64 mask, u64* __restrict__ out) {
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = in[i] & mask;
}

int mc2_count_batch(mc2_engine* e, const void* const* texts, const uint64_t* nbytes, uint32_t n, int space, int k, int64_t min_count,
                    mc2_table** out) {
    API_BEGIN
    if (!e || !out || (n && (!texts || !nbytes))) throw Mc2Error(MC2_ERR_INVALID, "NULL argument");
    check_count_args(e, "", 0, k);
    CUDA_CHECK(cudaSetDevice(e->device));
    const u64 c = min_count < 1 ? 1 : (u64)min_count;
    for (u32 j = 0; j < n; ++j) out[j] = nullptr;
    if (n == 0) return MC2_OK;
    // layout: text j at off[j] (a multiple of the parse tile), then "\n>\n...\n" up to the next tile boundary: the header
    // line closes the last record of sample j and emits the separator that keeps windows inside their sample
    std::vector<u64> off(n + 1, 0);
    for (u32 j = 0; j < n; ++j) {
        if (!texts[j] && nbytes[j]) throw Mc2Error(MC2_ERR_INVALID, "NULL text");
        off[j + 1] = (off[j] + nbytes[j] + 3 + PARSE_TILE - 1) / PARSE_TILE * PARSE_TILE;
    }
    const u64 total = off[n];
    DBuf<u8> cat(e, total + 16);
    CUDA_CHECK(cudaMemsetAsync(cat.p, '\n', total, e->stream));
    for (u32 j = 0; j < n; ++j) {
        if (nbytes[j]) {
            CUDA_CHECK(cudaMemcpyAsync(cat.p + off[j], texts[j], nbytes[j], space == MC2_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, e->stream));
            if (space != MC2_DEVICE) e->h2d_bytes += nbytes[j];
        }
        CUDA_CHECK(cudaMemsetAsync(cat.p + off[j] + nbytes[j] + 1, '>', 1, e->stream));
    }
    auto fallback = [&]() {
        for (u32 j = 0; j < n; ++j)
            if (!out[j]) count_one_text(e, cat.p + off[j], nbytes[j], k, c, &out[j]);
    };
    try {
        Parsed ps;
        DBuf<u64> tile_off;
        parse_text(e, cat.p, total, 0, ps, &tile_off);
        e->chunks += n;
        Plan plan;
        if (ps.stats.n_ascii) make_plan(e, ps.stats, k, plan);
        int sbits = 0;
        while ((1u << sbits) < n) ++sbits;
        const int bits = plan.enc >= 0 ? enc_bits(plan.enc) : 8;
        const u64 n_fast = plan.enc == ENC_NT2 ? ps.stats.n_acgt : plan.enc == ENC_AA5 ? ps.stats.n_upper : ps.stats.n_ascii;
        mc2_sample probe;
        probe.e = e; probe.k = k; probe.c = c;
        const bool batchable = ps.stats.n_ascii > 0 && plan.enc >= 0 && plan.path != PATH_WIDE && k * bits + sbits <= 64 &&
                               n_fast == ps.stats.n_ascii && ps.nsym <= range_batch_max(e, &probe) && ps.nsym < (1ull << 32) &&
                               e->opt_sparse_algo != 1 && e->opt_force_path != PATH_WIDE;
        if (!batchable) {
            if (ps.stats.n_ascii == 0) {                    // nothing but headers: n empty tables
                for (u32 j = 0; j < n; ++j) { mc2_sample s0; s0.e = e; s0.k = k; s0.c = c; out[j] = sample_finish(&s0); }
                return MC2_OK;
            }
            fallback();
            return MC2_OK;
        }
        // symbol index where every sample starts = symbol offset of its first parse tile
        std::vector<u64> tiles(n);
        for (u32 j = 0; j < n; ++j) tiles[j] = off[j] / PARSE_TILE;
        DBuf<u64> dtiles(e, n), starts(e, n);
        CUDA_CHECK(cudaMemcpyAsync(dtiles.p, tiles.data(), n * 8ull, cudaMemcpyHostToDevice, e->stream));
        LAUNCH(e, gather_u64_kernel, (unsigned)div_up(n, 256), 256, 0, (const u64*)tile_off.p, (const u64*)dtiles.p, (u64)n, starts.p);
        mc2_sample s;
        s.e = e; s.k = k; s.c = c;
        s.plan = plan;
        s.plan.path = PATH_SPARSE;
        SymView v{ps.sym.p, ps.nsym, starts.p, n, (u32)(k * bits)};
        if (plan.enc == ENC_NT2) sparse_chunk_range<ENC_NT2>(e, &s, v);
        else if (plan.enc == ENC_AA5) sparse_chunk_range<ENC_AA5>(e, &s, v);
        else sparse_chunk_range<ENC_BYTE>(e, &s, v);
        FastPart all;
        if (!s.fast.empty()) reduce_fast_parts(e, s.fast, k * bits + sbits, 1, all);
        // cut the (sample, k-mer)-sorted rows at the sample boundaries
        std::vector<u64> cuts(n + 1, 0);
        cuts[n] = all.n;
        if (all.n && n > 1) {
            std::vector<u64> split(n - 1);
            for (u32 j = 1; j < n; ++j) split[j - 1] = (u64)j << (k * bits);
            DBuf<u64> sp(e, n - 1), lb(e, n - 1);
            CUDA_CHECK(cudaMemcpyAsync(sp.p, split.data(), (n - 1) * 8ull, cudaMemcpyHostToDevice, e->stream));
            LAUNCH(e, lower_bound_kernel, (unsigned)div_up(n - 1, 128), 128, 0, (const u64*)all.keys.p, (u64)all.n, (const u64*)sp.p, (u64)(n - 1), lb.p);
            d2h(e, cuts.data() + 1, (const u64*)lb.p, (u64)(n - 1));
        } else {
            CUDA_CHECK(cudaStreamSynchronize(e->stream));        // (host vectors were the sources of async copies)
        }
        const u64 kmask = k * bits >= 64 ? ~0ull : ((1ull << (k * bits)) - 1);
        for (u32 j = 0; j < n; ++j) {
            std::unique_ptr<mc2_table> t(new mc2_table);
            t->e = e;
            t->k = k;
            t->enc = plan.enc;
            const u64 rows = cuts[j + 1] - cuts[j];
            t->fast.n = rows;
            t->fast.keys.alloc(e, rows);
            t->fast.counts.alloc(e, rows);
            if (rows) {
                LAUNCH(e, mask_keys_kernel, (unsigned)div_up(rows, 256), 256, 0, (const u64*)all.keys.p + cuts[j], rows, kmask, t->fast.keys.p);
                CUDA_CHECK(cudaMemcpyAsync(t->fast.counts.p, all.counts.p + cuts[j], rows * 8, cudaMemcpyDeviceToDevice, e->stream));
            }
            out[j] = t.release();
        }
        CUDA_CHECK(cudaStreamSynchronize(e->stream));
    } catch (...) {
        for (u32 j = 0; j < n; ++j) { if (out[j]) { delete out[j]; out[j] = nullptr; } }
        throw;
    }
    API_END
}

// host_rows (16-byte rows) or host_keys + host_counts (split delivery, 12 bytes per row)
static void count_text_rows_impl(mc2_engine* e, const void* text, uint64_t nbytes, int space, int k, int64_t min_count, void* host_rows,
                                 uint64_t* host_keys, uint32_t* host_counts, uint64_t capacity, uint64_t* rows) {
    check_count_args(e, text, nbytes, k);
    const bool split = host_keys != nullptr || host_counts != nullptr;
    if (!rows || (capacity && (split ? (!host_keys || !host_counts) : !host_rows))) throw Mc2Error(MC2_ERR_INVALID, "NULL argument");
    CUDA_CHECK(cudaSetDevice(e->device));
    mc2_sample s;
    s.e = e;
    s.k = k;
    s.c = min_count < 1 ? 1 : (u64)min_count;
    CUDA_CHECK(cudaEventRecord(e->ev0, e->stream));
    e->host_rows = HostRowSink();
    e->host_rows.active = true;
    e->host_rows.rows = host_rows;
    e->host_rows.keys64 = split ? host_keys : nullptr;
    e->host_rows.counts32 = split ? host_counts : nullptr;
    e->host_rows.capacity = capacity;
    HostRowSink hs;
    try {
        DBuf<u8> holder;
        const u8* d = to_device(e, text, nbytes, space, holder);
        if (nbytes) count_chunk(e, &s, d, nbytes);
        hs = e->host_rows;
        e->host_rows = HostRowSink();
    } catch (...) {
        cudaStreamSynchronize(e->copy_stream);
        e->host_rows = HostRowSink();
        throw;
    }
    if (hs.complete && s.fast.empty() && s.wide.empty() && s.plan.path == PATH_SPARSE) {
        if (hs.too_big) throw Mc2Error(MC2_ERR_LIMIT, "a count does not fit 32 bits: use mc2_count_text_rows");
        *rows = hs.delivered;                                   // every row already sits in the caller's buffer
    } else {
        std::unique_ptr<mc2_table> t(sample_finish(&s));
        if (t->wide.n) throw Mc2Error(MC2_ERR_LIMIT, "the table holds literal-byte rows (k-mers outside the packed alphabet): use mc2_count_text");
        if (t->fast.n > capacity) throw Mc2Error(MC2_ERR_INVALID, "buffer too small");
        if (t->fast.n && !split) {
            DBuf<RcRow> aos(e, t->fast.n);
            LAUNCH(e, rc_join_rows_kernel, (unsigned)div_up(t->fast.n, 256), 256, 0, (const u64*)t->fast.keys.p, (const u64*)t->fast.counts.p, (u64)t->fast.n, aos.p);
            CUDA_CHECK(cudaMemcpyAsync(host_rows, aos.p, t->fast.n * sizeof(RcRow), cudaMemcpyDeviceToHost, e->stream));
            CUDA_CHECK(cudaStreamSynchronize(e->stream));
            e->d2h_bytes += t->fast.n * sizeof(RcRow);
        } else if (t->fast.n) {
            DBuf<u32> narrow(e, t->fast.n), flag(e, 1);
            flag.zero();
            LAUNCH(e, narrow_counts_kernel, (unsigned)div_up(t->fast.n, 256), 256, 0, (const u64*)t->fast.counts.p, (u64)t->fast.n, narrow.p, flag.p);
            CUDA_CHECK(cudaMemcpyAsync(host_keys, t->fast.keys.p, t->fast.n * 8, cudaMemcpyDeviceToHost, e->stream));
            CUDA_CHECK(cudaMemcpyAsync(host_counts, narrow.p, t->fast.n * 4, cudaMemcpyDeviceToHost, e->stream));
            const u32 too_big = read_scalar<u32>(e, flag.p);
            e->d2h_bytes += t->fast.n * 12;
            if (too_big) throw Mc2Error(MC2_ERR_LIMIT, "a count does not fit 32 bits: use mc2_count_text_rows");
        }
        *rows = t->fast.n;
    }
    CUDA_CHECK(cudaEventRecord(e->ev1, e->stream));
    CUDA_CHECK(cudaEventSynchronize(e->ev1));
    float ms = 0;
    CUDA_CHECK(cudaEventElapsedTime(&ms, e->ev0, e->ev1));
    e->device_us = ms * 1000.0;
    e->resolve_profile();
}

int mc2_count_text_rows(mc2_engine* e, const void* text, uint64_t nbytes, int space, int k, int64_t min_count, void* host_rows,
                        uint64_t capacity, uint64_t* rows) {
    API_BEGIN
    count_text_rows_impl(e, text, nbytes, space, k, min_count, host_rows, nullptr, nullptr, capacity, rows);
    API_END
}

int mc2_count_text_rows_split(mc2_engine* e, const void* text, uint64_t nbytes, int space, int k, int64_t min_count,
                              uint64_t* host_keys, uint32_t* host_counts, uint64_t capacity, uint64_t* rows) {
    API_BEGIN
    if (!host_keys || !host_counts) throw Mc2Error(MC2_ERR_INVALID, "NULL argument");
    count_text_rows_impl(e, text, nbytes, space, k, min_count, nullptr, host_keys, host_counts, capacity, rows);
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
