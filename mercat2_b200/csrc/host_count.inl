// Part of mc2.cu (textually included there, in this order): per-chunk counting: hash path, dense and literal-byte paths, packed nucleotide lane, level-0 partition.

// =====================================================================================================
// per-chunk counting
// =====================================================================================================
template <int ENC>
static void dense_batch(mc2_engine* e, const Plan& plan, SymView v, u64 s0, u64 s1, int k, u32* table) {
    const u64 ntiles = div_up(s1 - s0, EX_TILE);
    if (plan.smem) {
        auto kern = dense_smem_kernel<ENC>;
        const size_t smem = (size_t)plan.bins * plan.nrep * 4;
        if (!e->dense_attrs_set[ENC]) {          // (function attributes are per device: remembered per engine, not per thread)
            CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
            e->dense_attrs_set[ENC] = true;
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

// sort + RLE of an already materialised key range (fallback for overflowed sub-buckets)
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

static void prefetch_next_count_pass(mc2_engine* e, mc2_sample* s);

// Where the sorted rows of the groups of ONE very large chunk go when the host does not wait for each group
// (min_count >= 2): back to back, as 16-byte rows, into the part of the level-0 key array whose keys have already been
// consumed -- a group of n keys has at most n / 2 surviving rows, so the rows of groups 0..g (16 B each) always fit in
// front of group g + 1's keys (8 B each).  counters (device): [0] rows so far, [1] overflow keys, [2] flagged keys
// (bitmap mode), [3] keys of all groups, [4] overflowed sub-buckets.
struct GroupSink {
    RcRow* arena;
    ull* counters;
    u64* ovf_keys;
    u64 ovf_cap;
};

struct KeySpan {               // keys already extracted (one level-0 group of a very large chunk, or keys received from other ranks)
    const u64* keys;
    u64 n;
    u64 p_lo, p_hi;            // the keys' 32-bit prefixes lie in [p_lo, p_hi)   (p_hi <= 2^32)
};

// One level of the range partition on the device (rangecount.cuh): LUT + level-1 descriptors (workspace memory).
struct RpPlan {
    u16* lut = nullptr;
    uint2* l1 = nullptr;
    u32* linear = nullptr;
    u32 base = 0, sh = 0, nb1 = 1, nidx = 1, down = 0, up = 0, mul = 0;
    RpView view() const { return RpView{lut, l1, base, sh, nb1, down, up, linear, mul}; }
};
static RangeWork& range_work(mc2_engine* e) {
    if (!e->work) e->work = new RangeWork;
    return *e->work;
}
static void plan_geometry(RpPlan& pl, int kb, u32 nb1, u64 p_lo, u64 p_hi) {
    pl.nb1 = nb1;
    pl.down = kb >= 32 ? (u32)(kb - 32) : 0u;
    pl.up = kb >= 32 ? 0u : (u32)(32 - kb);
    pl.base = (u32)p_lo;
    const u64 span = std::max<u64>(1, p_hi - p_lo);
    pl.sh = 0;
    while (((span - 1) >> pl.sh) >= RP_LUT) ++pl.sh;
    pl.nidx = (u32)((span - 1) >> pl.sh) + 1;
    pl.mul = (u32)(((u64)nb1 << RP_MUL_SHIFT) / pl.nidx);           // closed-form LUT: (index * mul) >> RP_MUL_SHIFT  (< nb1)
}
static const u64 RP_SAMPLE_WINDOWS = 1ull << 21;     // keys whose prefixes shape a level's LUT

// set the attributes of every kernel that needs more than 48 KB of dynamic shared memory (once per device)
template <int ENC>
static void range_kernel_attrs_enc() {
    CUDA_CHECK(cudaFuncSetAttribute(hc_hist_kernel<ENC>, cudaFuncAttributeMaxDynamicSharedMemorySize, 216 * 1024));
    CUDA_CHECK(cudaFuncSetAttribute(hc_scatter1_kernel<ENC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)HC_SCATTER_SMEM_LUT));
}
static void range_kernel_attrs(mc2_engine* e) {
    if (e->range_attrs_set) return;
    range_kernel_attrs_enc<ENC_NT2>();
    range_kernel_attrs_enc<ENC_AA5>();
    range_kernel_attrs_enc<ENC_BYTE>();
    CUDA_CHECK(cudaFuncSetAttribute(hk_hist_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 216 * 1024));
    CUDA_CHECK(cudaFuncSetAttribute(hk_hist_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 216 * 1024));
    CUDA_CHECK(cudaFuncSetAttribute(fn_hist_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 216 * 1024));
    CUDA_CHECK(cudaFuncSetAttribute(fn_hist_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 216 * 1024));
    CUDA_CHECK(cudaFuncSetAttribute(hc_scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((size_t)HC_MAX_NB1 * HC_NB2 * 4)));
    CUDA_CHECK(cudaFuncSetAttribute(hk_scatter1_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)HC_SCATTER_SMEM_LUT));
    CUDA_CHECK(cudaFuncSetAttribute(fn_scatter1_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)HC_SCATTER_SMEM_LUT));
    CUDA_CHECK(cudaFuncSetAttribute(hc_scatter2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)HC_SCATTER_SMEM16));
    CUDA_CHECK(cudaFuncSetAttribute(rc_count_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)RC_SMEM));
    CUDA_CHECK(cudaFuncSetAttribute(rc_count_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)RC_SMEM));
    CUDA_CHECK(cudaFuncSetAttribute(rc_count_kernel<0>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    CUDA_CHECK(cudaFuncSetAttribute(rc_count_kernel<1>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    CUDA_CHECK(cudaFuncSetAttribute(fn_dense_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    CUDA_CHECK(cudaFuncSetAttribute(hkv_scatter1_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)HKV_SCATTER_SMEM_LUT));
    CUDA_CHECK(cudaFuncSetAttribute(hkv_scatter2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)HKV_SCATTER_SMEM));
    CUDA_CHECK(cudaFuncSetAttribute(rc_merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)RM_SMEM));
    e->range_attrs_set = true;
}

// sampled prefix histogram of the source (packed streams, a key array, or a byte-symbol stream) -> plan tables
template <int ENC>
static void build_plan(mc2_engine* e, int k, const std::vector<PackedView>& pvs, const KeySpan* ks, SymView v, u64 cap, RpPlan& pl,
                       int level, const u32* shist_given = nullptr) {
    RangeWork& w = range_work(e);
    pl.lut = w.lut[level].get(e, RP_LUT);
    pl.l1 = w.l1[level].get(e, HC_MAX_NB1 + 1);
    pl.linear = reinterpret_cast<u32*>(pl.l1 + HC_MAX_NB1);         // (one spare descriptor slot holds the flag)
    struct { u32* p; } shist{nullptr};
    const u32* sh_p = shist_given;
    if (!sh_p) {
        shist.p = w.shist.get(e, RP_LUT);
        CUDA_CHECK(cudaMemsetAsync(shist.p, 0, RP_LUT * 4, e->stream));
        if (ks) {
            const u64 stride = std::max<u64>(1, ks->n / (RP_SAMPLE_WINDOWS / 4));       // (a narrow key range: a quarter of the sample does)
            const u64 grid = std::min<u64>(div_up(div_up(ks->n, stride), HK_HIST_THREADS), (u64)e->num_sms * 2);
            if (ks->n) LAUNCH(e, rp_sample_keys_kernel, (unsigned)std::max<u64>(grid, 1), HK_HIST_THREADS, 0, ks->keys, ks->n, stride, pl.down, pl.up, pl.base, pl.sh, shist.p);
        } else if (!pvs.empty()) {
            const u64 stride = std::max<u64>(1, cap / RP_SAMPLE_WINDOWS);
            for (auto& pv : pvs) {
                const u64 nwords = div_up(pv.n, 16);
                const u64 grid = std::min<u64>(div_up(div_up(nwords, stride), FN_HIST_THREADS), (u64)e->num_sms * 2);
                if (nwords) LAUNCH(e, rp_sample_packed_kernel, (unsigned)std::max<u64>(grid, 1), FN_HIST_THREADS, 0, pv, k, stride, pl.base, pl.sh, shist.p);
            }
        } else {
            const u64 ntiles = div_up(v.n, EX_TILE);
            const u64 stride = std::max<u64>(1, cap / RP_SAMPLE_WINDOWS);
            const u64 grid = std::min<u64>(div_up(ntiles, stride), (u64)e->num_sms * 4);
            auto kern = rp_sample_sym_kernel<ENC>;
            if (ntiles) LAUNCHN(e, "rp_sample_sym_kernel", kern, (unsigned)std::max<u64>(grid, 1), EX_THREADS, 0, v, k, stride, pl.down, pl.up, pl.base, pl.sh, shist.p);
        }
        sh_p = shist.p;
    }
    LAUNCH(e, rp_plan_kernel, 1, 1024, 0, sh_p, pl.nb1, pl.base, pl.sh, pl.nidx, pl.mul, pl.lut, pl.l1, pl.linear);
}

// keys a sub-bucket is sized for.  min_count 1 keeps every distinct key in the table (smaller buckets); duplicate-rich
// data (most keys repeat: few distinct keys per bucket) takes 25 % more keys per bucket -- still one register round for
// most buckets (two rounds measured 9 % slower in the counting kernel), and a 2.1e8-key exchange round of the 8-GPU
// job then fits one two-level pass.
#define RC_DISTINCT_TARGET 1600.0
#define RC_NB1_SWEET 192u                 // level-1 bins up to which the scatters keep their speed (measured: 46 ms / 10^10 keys at <= 220 bins, 95 ms at 384)
static double bucket_fill(const mc2_engine* e, const mc2_sample* s) {
    if (s->c < 2) return 0.7;
    const bool direct = e->opt_count_mode >= 0 ? e->opt_count_mode == 0 : s->dup_rich;
    return direct && s->dup_rich ? 1.25 : 1.0;
}
// Every distinct key takes a table slot, however often it occurs: with r keys per distinct key (measured on an earlier
// chunk / group of the sample) a sub-bucket may hold r times the keys -- further rounds of the counting kernel -- for the
// same table load.  Used only where the default size would need more than RC_NB1_SWEET level-1 bins: larger sub-buckets cost
// the counting kernel ~25 % (measured), too many bins cost the scatters 2x.
static double bucket_fill_max(const mc2_engine* e, const mc2_sample* s) {
    const double base = bucket_fill(e, s);
    if (base < 1.25 || !(s->dup_ratio > 0) || !e->opt_bucket_growth) return base;
    return std::min(16384.0 / 3500.0, std::max(base, RC_DISTINCT_TARGET * s->dup_ratio / 3500.0));
}
// keys one two-level partition can take (beyond it: level-0 partition first)
static u64 range_batch_max(const mc2_engine* e, const mc2_sample* s) {
    return std::min<u64>((u64)((double)HC_MAX_NB1 * HC_NB2 * (double)e->opt_hash_bucket_keys * bucket_fill_max(e, s)), e->opt_batch_symbols);
}

// Range partition + shared-memory tables (rangecount.cuh); the chunk must fit one batch.  Keys come from the byte
// symbol stream `v` (encoding ENC), from the packed nucleotide stream `pv`, or from a key array `ks`.  The surviving
// rows are appended to the sample as ONE part that is already sorted by key.
template <int ENC>
static void sparse_chunk_range(mc2_engine* e, mc2_sample* s, SymView v, const PackedView* pv = nullptr, const KeySpan* ks = nullptr,
                               const GroupSink* sink = nullptr) {
    range_kernel_attrs(e);
    const int k = s->k;
    int sample_bits = 0;                                         // batched samples: the sample index rides above the k-mer code
    if (v.sample_start) while ((1u << sample_bits) < v.n_samples) ++sample_bits;
    const int kb = k * EncTraits<ENC>::BITS + sample_bits;
    // Persistent kernels divide their work statically over a grid that just fills the GPU.  When something else holds SMs
    // (the NCCL kernels of a key exchange in flight) part of such a grid runs as a second wave and the kernel takes
    // twice as long (measured: hk_hist 18 -> 35 ms, hk_scatter1 46 -> 81 ms per 10^10 keys); "grid_waves" = w launches w
    // times the CTAs, so that a CTA that starts late costs 1 / w of the kernel instead.
    // (0 = the default, 2: measured 310.2 against 312.1 ms per headline step at N = 1; the key exchange asks for 4)
    const u64 waves = e->opt_grid_waves > 0 ? (u64)std::min(16, e->opt_grid_waves) : 2;
    const u64 cap = ks ? ks->n : pv ? pv->n : v.n;              // upper bound on the number of windows
    if (cap == 0) return;
    const u32 c = (u32)std::min<u64>(s->c, 0xFFFFFFFFull);
    // MODE 0 keeps every distinct key of a sub-bucket in the table, MODE 1 only the repeated ones
    const int mode = c < 2 ? 0 : e->opt_count_mode >= 0 ? (e->opt_count_mode ? 1 : 0) : (s->dup_rich ? 0 : 1);
    // (`cap` of a symbol stream counts ~25 % more positions than windows; a key array is exact, so aim lower there to
    // keep the same head room below the keys a table may hold)
    const double fill = (ks ? 0.8 : 1.0) * bucket_fill(e, s);
    u64 bucket_keys = std::max<u64>(1, (u64)((double)e->opt_hash_bucket_keys * s->bucket_scale * fill));
    if (div_up(cap, bucket_keys * HC_NB2) > RC_NB1_SWEET) {
        const u64 most = (u64)((double)e->opt_hash_bucket_keys * s->bucket_scale * (ks ? 0.8 : 1.0) * bucket_fill_max(e, s));
        bucket_keys = std::max(bucket_keys, std::min(most, div_up(cap, (u64)RC_NB1_SWEET * HC_NB2)));
    }
    const u32 nb1 = (u32)std::min<u64>(HC_MAX_NB1, std::max<u64>(1, div_up(cap, bucket_keys * HC_NB2)));
    const u32 nb = nb1 * HC_NB2;
    RpPlan pl;
    plan_geometry(pl, kb, nb1, ks ? ks->p_lo : 0, ks ? ks->p_hi : (1ull << 32));
    std::vector<PackedView> pvs;
    if (pv) pvs.push_back(*pv);
    build_plan<ENC>(e, k, pvs, ks, v, cap, pl, 1);
    const RpView rv = pl.view();
    struct Tail { ull total, rows, flagged; u32 ovf_n, pad; };
    // (the bucket histogram and the result counters share one allocation: one memset per chunk instead of two)
    RangeWork& w = range_work(e);
    struct { u32* p; } ghist{w.ghist.get(e, nb + sizeof(Tail) / 4)}, sub_base{w.sub_base.get(e, nb + 1)}, cur1{w.cur1.get(e, nb1)}, cur2{w.cur2.get(e, nb)},
        tile_pref{w.tile_pref.get(e, nb1 + 1)}, ovf_list{w.ovf_list.get(e, nb)}, rows{w.rows.get(e, nb)};
    struct { u64* p; } row_off{w.row_off.get(e, nb)};
    struct { Tail* p; } tail{reinterpret_cast<Tail*>(ghist.p + nb)};       // nb is a multiple of 128: 8-byte aligned
    CUDA_CHECK(cudaMemsetAsync(ghist.p, 0, (size_t)nb * 4 + sizeof(Tail), e->stream));
    const size_t hist_smem = sizeof(RpShared) + (size_t)nb * 4;
    if (ks) {
        const u64 grid = std::min<u64>(div_up(cap, HK_HIST_THREADS * 8), (u64)e->num_sms * waves);
        LAUNCHN(e, "hk_hist_kernel", hk_hist_kernel<true>, (unsigned)std::max<u64>(grid, 1), HK_HIST_THREADS, hist_smem, ks->keys, ks->n, rv, nb, ghist.p);
    } else if (pv) {
        const u64 nwords = div_up(cap, 16);
        const u64 grid = std::min<u64>(div_up(nwords, FN_HIST_THREADS), (u64)e->num_sms);
        LAUNCHN(e, "fn_hist_kernel", fn_hist_kernel<true>, (unsigned)grid, FN_HIST_THREADS, hist_smem, *pv, k, rv, nb, ghist.p);
    } else {
        auto kern = hc_hist_kernel<ENC>;
        int per_sm = 1;
        CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, EX_THREADS, hist_smem));
        const u64 grid = std::min<u64>(div_up(cap, EX_TILE), (u64)e->num_sms * std::max(per_sm, 1));
        LAUNCHN(e, "hc_hist_kernel", kern, (unsigned)grid, EX_THREADS, hist_smem, v, (u64)0, v.n, k, rv, nb, ghist.p);
    }
    LAUNCH(e, hc_scan_kernel, 1, 1024, (size_t)nb * 4, (const u32*)ghist.p, nb, nb1, (u32)HC_NB2, sub_base.p, cur1.p, cur2.p, tile_pref.p, &tail.p->total);
    // keys1 also serves as the survivors' slot array (16-byte rows at slot sub_base[b] / c) once scatter 2 has read it:
    // for c >= 2 the slots of cap keys need 16 * (cap / c + 1) <= 8 * cap + 16 bytes
    struct { u64* p; } keys1{w.keys1.get(e, cap + 8)}, keys2{w.keys2.get(e, cap)};
    RcRow* slots = reinterpret_cast<RcRow*>(keys1.p);
    if (c < 2) slots = reinterpret_cast<RcRow*>(w.slots1.get(e, (cap + 2) * sizeof(RcRow)));
    const bool dbg = getenv("MC2_DEBUG_HASH") != nullptr;
    if (dbg) {
        CUDA_CHECK(cudaMemsetAsync(keys1.p, 0xEE, cap * 8, e->stream));
        CUDA_CHECK(cudaMemsetAsync(keys2.p, 0xEE, cap * 8, e->stream));
    }
    if (ks) {
        LAUNCH(e, hk_scatter1_kernel, (unsigned)std::min<u64>(div_up(cap, HC_TILE), (u64)e->num_sms * 3 * waves), EX_THREADS, HC_SCATTER_SMEM_LUT, ks->keys, ks->n, rv, cur1.p, keys1.p, (const u64*)nullptr);
    } else if (pv) {
        LAUNCH(e, fn_scatter1_kernel, (unsigned)std::min<u64>(div_up(div_up(cap, 16), EX_THREADS), (u64)e->num_sms * 3), EX_THREADS, HC_SCATTER_SMEM_LUT, *pv, k, rv, cur1.p, keys1.p, (const u64*)nullptr);
    } else {
        auto kern = hc_scatter1_kernel<ENC>;
        LAUNCHN(e, "hc_scatter1_kernel", kern, (unsigned)div_up(cap, EX_TILE), EX_THREADS, HC_SCATTER_SMEM_LUT, v, (u64)0, v.n, k, rv, cur1.p, keys1.p);
    }
    LAUNCH(e, hc_scatter2_kernel, (unsigned)(div_up(cap, HC_TILE) + nb1), EX_THREADS, HC_SCATTER_SMEM16, (const u64*)keys1.p, (const u32*)sub_base.p,
           (const u32*)tile_pref.p, nb, (u32)HC_NB2, rv, cur2.p, keys2.p);
    if (dbg) {
        DBuf<ull> badc(e, 2);
        badc.zero();
        const u64 total = (u64)read_scalar<ull>(e, &tail.p->total);
        if (total) {
            LAUNCH(e, hc_verify_kernel, (unsigned)div_up(total, 256), 256, 0, (const u64*)keys1.p, (const u32*)sub_base.p, nb, (u32)HC_NB2, (u32)total, badc.p, rv);
            LAUNCH(e, hc_verify_kernel, (unsigned)div_up(total, 256), 256, 0, (const u64*)keys2.p, (const u32*)sub_base.p, nb, 1u, (u32)total, badc.p + 1, rv);
        }
        const ull b1 = read_scalar<ull>(e, badc.p), b2 = read_scalar<ull>(e, badc.p + 1);
        fprintf(stderr, "[range] cap=%llu total=%llu nb1=%u nb=%u packed=%d mode=%d misplaced level1=%llu level2=%llu\n", (ull)cap, (ull)total, nb1, nb,
                pv ? 1 : 0, mode, b1, b2);
        DBuf<ull> cs(e, 9);
        cs.zero();
        if (pv) LAUNCH(e, fn_checksum_kernel, 256, 256, 0, *pv, k, cs.p);
        LAUNCH(e, key_checksum_kernel, 256, 256, 0, (const u64*)keys1.p, total, cs.p + 3);
        LAUNCH(e, key_checksum_kernel, 256, 256, 0, (const u64*)keys2.p, total, cs.p + 6);
        ull h[9];
        d2h(e, h, cs.p, 9);
        fprintf(stderr, "[range] checksum stream (%llx %llx %llu) keys1 (%llx %llx %llu) keys2 (%llx %llx %llu)%s\n", h[0], h[1], h[2], h[3],
                h[4], h[5], h[6], h[7], h[8], ((pv && (h[0] != h[6] || h[1] != h[7] || h[0] != h[3])) || b1 || b2) ? "  MISMATCH" : "");
    }
    const unsigned cgrid = (unsigned)std::min<u64>(nb, 2ull * e->num_sms * waves);
    ull* flagged_dev = sink ? sink->counters + (mode == 1 ? 2 : 5) : &tail.p->flagged;      // (MODE 0 reports distinct keys there)
    if (mode == 1)
        LAUNCHN(e, "rc_count_kernel<1>", rc_count_kernel<1>, cgrid, RC_THREADS, RC_SMEM, (const u64*)keys2.p, (const u32*)sub_base.p, nb, c, rv, slots, rows.p,
                ovf_list.p, &tail.p->ovf_n, flagged_dev);
    else
        LAUNCHN(e, "rc_count_kernel<0>", rc_count_kernel<0>, cgrid, RC_THREADS, RC_SMEM, (const u64*)keys2.p, (const u32*)sub_base.p, nb, c, rv, slots, rows.p,
                ovf_list.p, &tail.p->ovf_n, flagged_dev);
    if (sink) {
        // no host round trip: rows go behind the rows of the earlier groups (device counter), overflowed sub-buckets are
        // copied aside by the device, the statistics accumulate in the sink's counters
        LAUNCH(e, rc_offsets_kernel, 1, 1024, 0, (const u32*)rows.p, nb, row_off.p, sink->counters, &tail.p->rows, 1);
        LAUNCH(e, rc_gather_kernel, (unsigned)std::min<u64>(div_up(nb, 8), (u64)e->num_sms * 8), 256, 0, (const RcRow*)slots, (const u32*)sub_base.p,
               (const u32*)rows.p, (const u64*)row_off.p, nb, c, (u64*)nullptr, (u64*)nullptr, sink->arena);
        LAUNCH(e, rc_overflow_collect_kernel, 64, 256, 0, (const u32*)ovf_list.p, (const u32*)&tail.p->ovf_n, (const u32*)sub_base.p, (const u64*)keys2.p,
               (const ull*)&tail.p->total, sink->ovf_keys, sink->ovf_cap, sink->counters);
        return;
    }
    LAUNCH(e, rc_offsets_kernel, 1, 1024, 0, (const u32*)rows.p, nb, row_off.p, (ull*)nullptr, &tail.p->rows, 0);
    if (pv && !ks) prefetch_next_count_pass(e, s);               // rides on the synchronisation below
    const Tail t = read_scalar<Tail>(e, tail.p);
    if (t.rows > cap / c + 1) throw Mc2Error(MC2_ERR_LIMIT, "range path: more survivors than slots (internal error)");
    if (ks && getenv("MC2_DEBUG_PHASES"))
        fprintf(stderr, "[phase]   group: %llu keys, %u buckets, mode %d, %llu survivors, %u overflowed buckets\n", (ull)cap, nb, mode, (ull)t.rows, t.ovf_n);
    if (t.rows) {
        FastPart part;
        part.n = t.rows;
        part.sorted = true;
        part.keys.alloc(e, t.rows);
        part.counts.alloc(e, t.rows);
        LAUNCH(e, rc_gather_kernel, (unsigned)std::min<u64>(div_up(nb, 8), (u64)e->num_sms * 8), 256, 0, (const RcRow*)slots, (const u32*)sub_base.p,
               (const u32*)rows.p, (const u64*)row_off.p, nb, c, part.keys.p, part.counts.p, (RcRow*)nullptr);
        s->fast.push_back(std::move(part));
    }
    if (dbg) {                                                    // the sort path on the same keys must agree
        mc2_sample tmp;
        tmp.e = e; tmp.k = k; tmp.c = s->c;
        count_key_range_sorted(e, &tmp, keys2.p, t.total, 64);
        u64 ref_rows = 0;
        for (auto& p : tmp.fast) ref_rows += p.n;
        fprintf(stderr, "[range] survivors: tables %llu (+%u overflowed buckets) sort %llu%s\n", (ull)t.rows, t.ovf_n, (ull)ref_rows,
                (!t.ovf_n && ref_rows != t.rows) ? "  MISMATCH" : "");
    }
    // duplicate-rich data (at least half of the keys found their bitmap bit already set): the following chunks / groups of the sample
    // skip the bitmap pre-filter.  If that makes tables overflow, go back.
    if (c >= 2) {
        if (mode == 1 && t.flagged * 2 >= t.total && t.total >= 65536) s->dup_rich = true;      // half of the keys are repeats
        if (mode == 0 && (u64)t.ovf_n * 50 > nb) s->dup_rich = false;
        if (mode == 0 && t.flagged && t.total >= 65536) s->dup_ratio = (double)t.total / (double)t.flagged;
        if (mode == 0 && (u64)t.ovf_n * 200 > nb) s->dup_ratio = std::max(1.0, s->dup_ratio * 0.5);
    }
    if (t.ovf_n) {
        // Sub-buckets whose distinct (MODE 0) or repeated (MODE 1) keys do not fit the table: their keys are gathered
        // into ONE array and counted by a single sort + run-length pass (sub-buckets hold disjoint key sets).  The
        // overflow rate also steers the bucket size of the sample's next chunks.
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
            CUDA_CHECK(cudaStreamSynchronize(e->stream));          // (host vectors were the sources of async copies)
            count_key_range_sorted(e, s, gathered.p, m, kb);
        }
        e->ovf_buckets += t.ovf_n;
        if ((u64)t.ovf_n * 200 > nb && s->bucket_scale > 0.3) s->bucket_scale *= 0.8;      // > 0.5 % of the buckets overflowed
    }
}

// Sum M (key, count) rows by key through the range partition (rowmerge.cuh): `out` receives the summed rows sorted by
// key.  Returns false (nothing done) when the rows are too few to be worth it, too many for one two-level pass, or a
// sub-bucket held more distinct keys than its table (the caller then sorts).
static u64 merge_rows_capacity(const mc2_engine* e) {
    return std::min<u64>((u64)((double)HC_MAX_NB1 * HC_NB2 * (double)e->opt_hash_bucket_keys * 0.45), (1ull << 32) - 1);
}
static bool merge_rows_range(mc2_engine* e, const u64* keys, const u64* counts, u64 M, int key_bits, FastPart& out) {
    if (M < 32768 || M > merge_rows_capacity(e)) return false;
    e->row_merges++;
    range_kernel_attrs(e);
    const u64 bucket_keys = std::max<u64>(16, (u64)((double)e->opt_hash_bucket_keys * 0.45));     // every row of a bucket may be a distinct key, and sub-buckets are uneven
    const u32 nb1 = (u32)std::min<u64>(HC_MAX_NB1, std::max<u64>(1, div_up(M, bucket_keys * HC_NB2)));
    const u32 nb = nb1 * HC_NB2;
    RpPlan pl;
    plan_geometry(pl, key_bits, nb1, 0, 1ull << 32);
    KeySpan ks{keys, M, 0, 1ull << 32};
    build_plan<ENC_NT2>(e, 0, std::vector<PackedView>(), &ks, SymView{nullptr, 0}, M, pl, 1);
    const RpView rv = pl.view();
    RangeWork& w = range_work(e);
    struct Tail { ull total, rows; u32 ovf_n, pad; };
    u32* ghist = w.ghist.get(e, nb + sizeof(Tail) / 4 + 8);
    u32* sub_base = w.sub_base.get(e, nb + 1);
    u32* cur1 = w.cur1.get(e, nb1);
    u32* cur2 = w.cur2.get(e, nb);
    u32* tile_pref = w.tile_pref.get(e, nb1 + 1);
    u32* rows = w.rows.get(e, nb);
    u64* row_off = w.row_off.get(e, nb);
    Tail* tail = reinterpret_cast<Tail*>(ghist + nb);
    CUDA_CHECK(cudaMemsetAsync(ghist, 0, (size_t)nb * 4 + sizeof(Tail), e->stream));
    const size_t hist_smem = sizeof(RpShared) + (size_t)nb * 4;
    LAUNCHN(e, "hk_hist_kernel", hk_hist_kernel<true>, (unsigned)std::max<u64>(std::min<u64>(div_up(M, HK_HIST_THREADS * 8), (u64)e->num_sms), 1), HK_HIST_THREADS,
            hist_smem, keys, M, rv, nb, ghist);
    LAUNCH(e, hc_scan_kernel, 1, 1024, (size_t)nb * 4, (const u32*)ghist, nb, nb1, (u32)HC_NB2, sub_base, cur1, cur2, tile_pref, &tail->total);
    // keys1 / keys2 serve as key arrays, two fresh arrays carry the counts; the 16-byte row slots need 2 * M words
    u64* keys1 = w.keys1.get(e, M + 8);
    u64* keys2 = w.keys2.get(e, M + 8);
    DBuf<u64> vals1(e, M), vals2(e, M);
    DBuf<RcRow> slots(e, M + 2);
    LAUNCH(e, hkv_scatter1_kernel, (unsigned)std::min<u64>(div_up(M, HC_TILE), (u64)e->num_sms * 2), EX_THREADS, HKV_SCATTER_SMEM_LUT, keys, counts, M, rv, cur1,
           keys1, vals1.p);
    LAUNCH(e, hkv_scatter2_kernel, (unsigned)(div_up(M, HC_TILE) + nb1), EX_THREADS, HKV_SCATTER_SMEM, (const u64*)keys1, (const u64*)vals1.p,
           (const u32*)sub_base, (const u32*)tile_pref, nb, (u32)HC_NB2, rv, cur2, keys2, vals2.p);
    LAUNCH(e, rc_merge_kernel, (unsigned)std::min<u64>(nb, 2ull * e->num_sms), RM_THREADS, RM_SMEM, (const u64*)keys2, (const u64*)vals2.p, (const u32*)sub_base,
           nb, rv, slots.p, rows, &tail->ovf_n);
    LAUNCH(e, rc_offsets_kernel, 1, 1024, 0, (const u32*)rows, nb, row_off, (ull*)nullptr, &tail->rows, 0);
    const Tail t = read_scalar<Tail>(e, tail);
    if (t.ovf_n) return false;
    out.n = t.rows;
    out.sorted = true;
    out.keys.alloc(e, t.rows);
    out.counts.alloc(e, t.rows);
    if (t.rows)
        LAUNCH(e, rc_gather_kernel, (unsigned)std::min<u64>(div_up(nb, 8), (u64)e->num_sms * 8), 256, 0, (const RcRow*)slots.p, (const u32*)sub_base, (const u32*)rows,
               (const u64*)row_off, nb, 1u, out.keys.p, out.counts.p, (RcRow*)nullptr);
    return true;
}

template <int ENC>
static void sparse_chunk(mc2_engine* e, mc2_sample* s, SymView v) {
    {
        const bool want_range = e->opt_sparse_algo != 1;
        if (want_range && v.n <= range_batch_max(e, s) && v.n < (1ull << 32)) {
            sparse_chunk_range<ENC>(e, s, v);
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
    const u64 hash_max = range_batch_max(e, s);
    const bool ok = e->opt_fast_nt && !e->opt_parse_single && s->k <= 32 && e->opt_sparse_algo != 1 &&
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
    sp.alloc_packed(e, cap_sym);
    DBuf<ull> desc(e, sp.ntiles);
    DBuf<u32> ticket(e, 1);
    desc.zero();
    ticket.zero();
    if (with_stats)
        LAUNCHN(e, "fn_parse_single_kernel<stats>", fn_parse_single_kernel<true>, (unsigned)sp.ntiles, FN_THREADS, 0, sp.text, sp.len,
                (u32)sp.ntiles, desc.p, ticket.p, sp.codes.p, sp.bad, st.p);
    else
        LAUNCHN(e, "fn_parse_single_kernel", fn_parse_single_kernel<false>, (unsigned)sp.ntiles, FN_THREADS, 0, sp.text, sp.len,
                (u32)sp.ntiles, desc.p, ticket.p, sp.codes.p, sp.bad, st.p);
    const FnStats fs = read_scalar<FnStats>(e, st.p);
    sp.nsym = fs.n_sym;
    if (getenv("MC2_DEBUG_FAST"))
        fprintf(stderr, "[fast_nt] single pass len=%llu n_sym=%llu kept=%llu non_acgt=%llu complex=%llu\n", (ull)sp.len, fs.n_sym,
                fs.packed & 0xFFFFFFFFull, fs.packed >> 32, fs.complex);
    return fs;
}

// write pass: 2-bit codes + bad bits (also counts the kept non-ACGT bytes into st->packed2)
static void fn_write_pass(mc2_engine* e, FnSpan& sp, DBuf<FnStats>& st) {
    sp.alloc_packed(e, sp.nsym);
    LAUNCH(e, fn_parse_kernel<1>, (unsigned)sp.ntiles, FN_THREADS, 0, sp.text, sp.len, sp.tstate.p, sp.tcnt.p, (const u64*)sp.toff.p,
           sp.codes.p, sp.bad, st.p);
    sp.tstate.release();
    sp.tcnt.release();
    sp.toff.release();
}

static void fn_dense_span(mc2_engine* e, mc2_sample* s, const PackedView& pv) {
    const Plan& plan = s->plan;
    const u64 nwords = div_up(pv.n, 16);
    if (plan.smem) {
        const size_t smem = (size_t)plan.bins * plan.nrep * 4;
        range_kernel_attrs(e);
        const unsigned grid = (unsigned)std::min<u64>(div_up(nwords, FN_HIST_THREADS), (u64)e->num_sms * (smem <= 96 * 1024 ? 2 : 1));
        LAUNCHN(e, "fn_dense_kernel<smem>", fn_dense_kernel<true>, grid, FN_HIST_THREADS, smem, pv, s->k, plan.bins, plan.nrep, s->dense_chunk.p);
    } else {
        const unsigned grid = (unsigned)std::min<u64>(div_up(nwords, FN_HIST_THREADS), (u64)e->num_sms * 2);
        LAUNCHN(e, "fn_dense_kernel<global>", fn_dense_kernel<false>, grid, FN_HIST_THREADS, 0, pv, s->k, plan.bins, 1u, s->dense_chunk.p);
    }
}

// Level-0 partition of key sources into g0 groups of ascending, disjoint key ranges: keys0 (grouped, exact offsets in
// gbase[g0 + 1]; bounds[g] = first 32-bit prefix of group g, bounds[g0] = end of the partitioned range).  Sources:
// packed symbol streams (their windows) or one key array.  The group boundaries come from a sampled prefix histogram
// (`shist_given`: a histogram the caller already holds, e.g. summed over all ranks so that every rank cuts at the same
// keys).  Returns false if the result does not fit in free device memory (`extra` = bytes the caller still needs
// afterwards) or a group would exceed `group_max` keys.
struct Level0 {
    u64* keys0 = nullptr;                  // the engine's workspace array (borrowed) or, when that is taken, `own`
    DBuf<u64> own;
    bool borrowed = false;
    std::vector<u64> gbase;
    std::vector<u64> bounds;
    u64 gmax = 0;
};
static bool level0_partition(mc2_engine* e, int k, const std::vector<PackedView>& pvs, const KeySpan* ks, u32 g0, u64 group_max, u64 extra,
                             Level0& out, const u32* shist_given = nullptr, bool hold_workspace = false) {
    range_kernel_attrs(e);
    u64 cap = ks ? ks->n : 0;
    for (auto& pv : pvs) cap += pv.n;
    RpPlan pl;
    const u64 p_lo = ks ? ks->p_lo : 0, p_hi = ks ? ks->p_hi : (1ull << 32);
    plan_geometry(pl, 2 * k, g0, p_lo, p_hi);
    build_plan<ENC_NT2>(e, k, pvs, ks, SymView{nullptr, 0}, cap, pl, 0, shist_given);
    const RpView rv = pl.view();
    DBuf<u32> ghist(e, g0);
    ghist.zero();
    const size_t hist_smem = sizeof(RpShared) + (size_t)g0 * 4;
    PhaseTimer pt(e);
    if (ks) {
        const u64 grid = std::min<u64>(div_up(ks->n, HK_HIST_THREADS * 8), (u64)e->num_sms * 2);
        if (ks->n) LAUNCHN(e, "hk_hist_kernel<level0>", hk_hist_kernel<false>, (unsigned)std::max<u64>(grid, 1), HK_HIST_THREADS, hist_smem, ks->keys, ks->n, rv, g0, ghist.p);
    } else {
        for (auto& pv : pvs) {
            const u64 grid = std::min<u64>(div_up(div_up(pv.n, 16), FN_HIST_THREADS), (u64)e->num_sms * 2);
            if (grid) LAUNCHN(e, "fn_hist_kernel<level0>", fn_hist_kernel<false>, (unsigned)grid, FN_HIST_THREADS, hist_smem, pv, k, rv, g0, ghist.p);
        }
    }
    std::vector<u32> h(g0);
    std::vector<uint2> l1(g0);
    d2h(e, h.data(), (const u32*)ghist.p, g0);
    d2h(e, l1.data(), (const uint2*)pl.l1, g0);
    pt.mark("level-0 histogram");
    out.gbase.assign(g0 + 1, 0);
    out.bounds.assign(g0 + 1, p_hi);
    out.gmax = 0;
    for (u32 g = 0; g < g0; ++g) {
        out.gbase[g + 1] = out.gbase[g] + h[g];
        out.gmax = std::max<u64>(out.gmax, h[g]);
        out.bounds[g] = g == 0 ? p_lo : std::max<u64>(out.bounds[g - 1], std::min<u64>(l1[g].x, p_hi));
    }
    const u64 total = out.gbase[g0];
    if (total == 0) return true;
    if (out.gmax > group_max || out.gmax >= (1ull << 32)) return false;
    const RangeWork& wq = range_work(e);
    const bool fits_like_before = !wq.keys0_busy && wq.keys0.b.n >= total && total <= e->l0_fit_total && extra <= e->l0_fit_extra;
    if (!fits_like_before) {
        // the level-0 array plus what follows must fit (free memory + what the pool holds unused).  Asked only when this
        // call needs more than an earlier one that fitted in the workspace the engine still holds: cudaMemGetInfo was
        // seen to take ~55 ms every few calls while the pool has frees in flight.
        size_t free_b = 0, total_b = 0;
        CUDA_CHECK(cudaMemGetInfo(&free_b, &total_b));
        cudaMemPool_t pool;
        CUDA_CHECK(cudaDeviceGetDefaultMemPool(&pool, e->device));
        unsigned long long reserved = 0, used = 0;
        CUDA_CHECK(cudaMemPoolGetAttribute(pool, cudaMemPoolAttrReservedMemCurrent, &reserved));
        CUDA_CHECK(cudaMemPoolGetAttribute(pool, cudaMemPoolAttrUsedMemCurrent, &used));
        const u64 avail = (u64)free_b + (reserved > used ? (u64)(reserved - used) : 0);
        const RangeWork& w0 = range_work(e);
        const u64 have = w0.keys0_busy ? 0 : w0.keys0.b.n * 8;              // the workspace array is re-used (it is freed first if it must grow)
        if (total * 8 + extra + (1ull << 30) > avail + have) return false;
        e->l0_fit_total = std::max(e->l0_fit_total, total);
        e->l0_fit_extra = std::max(e->l0_fit_extra, extra);
    }
    pt.mark("level-0 memory check");
    {
        RangeWork& w = range_work(e);
        if (!w.keys0_busy) {
            out.keys0 = w.keys0.get(e, total);
            out.borrowed = hold_workspace;
            if (hold_workspace) w.keys0_busy = true;
        } else {
            out.own.alloc(e, total);
            out.keys0 = out.own.p;
        }
    }
    DBuf<u64> gbase_dev(e, g0 + 1);
    DBuf<u32> cur0(e, g0);
    cur0.zero();
    pt.mark("level-0 allocation");
    CUDA_CHECK(cudaMemcpyAsync(gbase_dev.p, out.gbase.data(), (g0 + 1) * 8, cudaMemcpyHostToDevice, e->stream));
    if (ks) {
        LAUNCH(e, hk_scatter1_kernel, (unsigned)std::min<u64>(div_up(ks->n, HC_TILE), (u64)e->num_sms * 3), EX_THREADS, HC_SCATTER_SMEM_LUT, ks->keys, ks->n, rv, cur0.p, out.keys0,
               (const u64*)gbase_dev.p);
    } else {
        for (auto& pv : pvs) {
            const u64 grid = std::min<u64>(div_up(div_up(pv.n, 16), EX_THREADS), (u64)e->num_sms * 3);
            if (grid) LAUNCH(e, fn_scatter1_kernel, (unsigned)grid, EX_THREADS, HC_SCATTER_SMEM_LUT, pv, k, rv, cur0.p, out.keys0, (const u64*)gbase_dev.p);
        }
    }
    CUDA_CHECK(cudaStreamSynchronize(e->stream));                  // (host vector was the source of an async copy)
    pt.mark("level-0 scatter");
    return true;
}

static u32 level0_groups(u64 cap, u64 hash_max) {
    return (u32)std::min<u64>(HC_MAX_NB1, std::max<u64>(2, div_up(cap, std::max<u64>(1, hash_max / 2))));
}

// A chunk with more windows than one batch holds (-s 0 on a large file): a level-0 partition of ALL its keys by key
// range into groups that fit, written once to HBM (8 B per window -- sized for the 180 GB of a B200), then the usual
// two-level pipeline per group.  All occurrences of a key meet in one group, so the -c filter stays exact for the
// whole chunk, and the groups' sorted rows follow each other in key order.  Returns false (nothing counted) when the
// keys do not fit in free device memory.
static bool sparse_chunk_big(mc2_engine* e, mc2_sample* s, const std::vector<PackedView>& pvs, const KeySpan* ks, u64 hash_max, bool allow_async = true) {
    u64 cap = ks ? ks->n : 0;
    for (auto& pv : pvs) cap += pv.n;
    const u32 g0 = level0_groups(cap, hash_max);
    if (div_up(cap, g0) > hash_max) return false;
    Level0 l0;
    if (!level0_partition(e, s->k, pvs, ks, g0, hash_max, 3 * 8 * div_up(cap, g0) * 2, l0)) return false;
    PhaseTimer pt(e);
    auto span_of = [&](u32 g) {
        return KeySpan{l0.keys0 + l0.gbase[g], l0.gbase[g + 1] - l0.gbase[g], l0.bounds[g], std::max<u64>(l0.bounds[g + 1], l0.bounds[g] + 1)};
    };
    const u64 total = l0.gbase[g0];
    if (s->c < 2 || !allow_async || e->opt_group_sync) {
        // min_count 1 (a group may emit as many 16-byte rows as it has 8-byte keys): one host round trip per group
        for (u32 g = 0; g < g0; ++g) {
            const KeySpan span = span_of(g);
            if (span.n) sparse_chunk_range<ENC_NT2>(e, s, SymView{nullptr, 0}, nullptr, &span);
        }
        pt.mark("groups");
        return true;
    }
    // min_count >= 2: the groups are enqueued back to back; their rows collect in the consumed front of the key array
    DBuf<ull> counters(e, 8);
    counters.zero();
    const u64 ovf_cap = std::max<u64>(l0.gmax, total / 32) + 4096;
    DBuf<u64> ovf_keys(e, ovf_cap);
    GroupSink sink{reinterpret_cast<RcRow*>(l0.keys0), counters.p, ovf_keys.p, ovf_cap};
    HostRowSink* hs = e->host_rows.active && !e->host_rows.failed ? &e->host_rows : nullptr;
    std::vector<cudaEvent_t> evs;
    u64 copied = 0;
    u32 next_copy = 0, launched = 0;
    ull* pin = nullptr;
    const u64 SPLIT_STAGE_ROWS = 16ull << 20;
    DBuf<u64> stage_keys;
    DBuf<u32> stage_counts, split_flag;
    if (hs && hs->keys64) {
        stage_keys.alloc(e, SPLIT_STAGE_ROWS);
        stage_counts.alloc(e, SPLIT_STAGE_ROWS);
        split_flag.alloc(e, 1);
        split_flag.zero();
    }
    if (hs) {
        if (!e->pin_groups) CUDA_CHECK(cudaMallocHost((void**)&e->pin_groups, (HC_MAX_NB1 + 8) * sizeof(ull)));
        pin = e->pin_groups;
    }
    auto drain = [&](bool all) {
        // hand the rows of finished groups to the copy stream (their count arrived with an earlier event)
        while (hs && next_copy < launched && (all || cudaEventQuery(evs[next_copy]) == cudaSuccess)) {
            if (all) CUDA_CHECK(cudaEventSynchronize(evs[next_copy]));
            const u64 upto = pin[next_copy];
            if (upto > hs->capacity) { hs->failed = true; hs = nullptr; break; }
            if (upto > copied) {
                CUDA_CHECK(cudaStreamWaitEvent(e->copy_stream, evs[next_copy], 0));
                if (!hs->keys64) {
                    CUDA_CHECK(cudaMemcpyAsync((RcRow*)hs->rows + copied, sink.arena + copied, (upto - copied) * sizeof(RcRow), cudaMemcpyDeviceToHost, e->copy_stream));
                    e->d2h_bytes += (upto - copied) * sizeof(RcRow);
                } else {
                    // split delivery: rows -> (key, 32-bit count) in a staging pair on the copy stream, 12 bytes per row leave
                    // (the split kernel and its two copies run in stream order, so one staging pair is enough)
                    for (u64 at = copied; at < upto; at += SPLIT_STAGE_ROWS) {
                        const u64 n = std::min<u64>(SPLIT_STAGE_ROWS, upto - at);
                        rc_split_rows32_kernel<<<(unsigned)div_up(n, 256), 256, 0, e->copy_stream>>>((const RcRow*)(sink.arena + at), n, stage_keys.p, stage_counts.p, split_flag.p);
                        CUDA_CHECK(cudaGetLastError());
                        e->launches++;
                        CUDA_CHECK(cudaMemcpyAsync(hs->keys64 + at, stage_keys.p, n * 8, cudaMemcpyDeviceToHost, e->copy_stream));
                        CUDA_CHECK(cudaMemcpyAsync(hs->counts32 + at, stage_counts.p, n * 4, cudaMemcpyDeviceToHost, e->copy_stream));
                    }
                    e->d2h_bytes += (upto - copied) * 12;
                }
                copied = upto;
            }
            ++next_copy;
        }
    };
    bool mode_known = s->dup_rich || e->opt_count_mode >= 0;
    bool ratio_known = s->dup_ratio > 0;
    for (u32 g = 0; g < g0; ++g) {
        const KeySpan span = span_of(g);
        if (!span.n) continue;
        const bool direct = e->opt_count_mode >= 0 ? e->opt_count_mode == 0 : s->dup_rich;
        sparse_chunk_range<ENC_NT2>(e, s, SymView{nullptr, 0}, nullptr, &span, &sink);
        if (!mode_known) {                                       // one look at the first group decides the counting mode of the rest
            ull c5[8];
            d2h(e, c5, (const ull*)counters.p, 8);
            if (c5[2] * 2 >= c5[3] && c5[3] >= 65536) s->dup_rich = true;
            mode_known = true;
        } else if (!ratio_known && direct && s->dup_rich) {      // ... and one at the first group counted without the pre-filter sizes the sub-buckets
            ull c5[8];
            d2h(e, c5, (const ull*)counters.p, 8);
            if (c5[5] && span.n >= 65536) s->dup_ratio = (double)span.n / (double)c5[5];
            ratio_known = true;
        }
        if (hs) {
            CUDA_CHECK(cudaMemcpyAsync(&pin[launched], counters.p, sizeof(ull), cudaMemcpyDeviceToHost, e->stream));
            cudaEvent_t ev = e->get_event();
            CUDA_CHECK(cudaEventRecord(ev, e->stream));
            evs.push_back(ev);
            ++launched;
            // stay at most three groups ahead of the device: the row counts of finished groups must reach the host while
            // later groups are still running, or nothing could be handed to the copy engine before the end
            if (launched >= 3) CUDA_CHECK(cudaEventSynchronize(evs[launched - 3]));
            drain(false);
        }
    }
    ull fin[8];
    d2h(e, fin, (const ull*)counters.p, 8);
    pt.mark("groups");
    drain(true);
    for (auto ev : evs) e->ev_pool.push_back(ev);
    const u64 R = fin[0], ovf_m = fin[1];
    if (ovf_m > ovf_cap) {
        // more overflowed keys than the side array holds (heavily skewed data): redo the chunk with one round trip per group
        CUDA_CHECK(cudaStreamSynchronize(e->copy_stream));
        if (e->host_rows.active) e->host_rows.failed = true;
        l0 = Level0();
        return sparse_chunk_big(e, s, pvs, ks, hash_max, false);
    }
    e->ovf_buckets += fin[4];
    if ((u64)fin[4] * 200 > div_up(total, std::max<u64>(1, e->opt_hash_bucket_keys)) && s->bucket_scale > 0.3) s->bucket_scale *= 0.8;
    if ((u64)fin[4] * 200 > div_up(total, (u64)(e->opt_hash_bucket_keys * bucket_fill_max(e, s)) + 1)) s->dup_ratio = std::max(1.0, s->dup_ratio * 0.5);
    const bool host_done = hs && !ovf_m;
    if (hs && ovf_m) e->host_rows.failed = true;               // rows of the sort path would have to be merged in: the caller falls back
    if (host_done) {
        CUDA_CHECK(cudaStreamSynchronize(e->copy_stream));
        if (hs->keys64 && read_scalar<u32>(e, split_flag.p)) e->host_rows.too_big = true;
        e->host_rows.delivered = R;
        e->host_rows.complete = true;
    } else if (R) {
        CUDA_CHECK(cudaStreamSynchronize(e->copy_stream));
        FastPart part;
        part.n = R;
        part.sorted = true;
        part.keys.alloc(e, R);
        part.counts.alloc(e, R);
        LAUNCH(e, rc_split_rows_kernel, (unsigned)div_up(R, 256), 256, 0, (const RcRow*)sink.arena, R, part.keys.p, part.counts.p);
        s->fast.push_back(std::move(part));
    }
    if (ovf_m) count_key_range_sorted(e, s, ovf_keys.p, ovf_m, 2 * s->k);
    CUDA_CHECK(cudaStreamSynchronize(e->stream));               // (the key array may be handed to the next chunk)
    pt.mark("rows");
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
    const u64 hash_max = range_batch_max(e, s);
    // which plans the packed lane serves: 2-bit sparse keys through the range partition, and dense 4^k tables
    auto served = [&](const Plan& pl) {
        if (pl.enc != ENC_NT2) return false;
        if (pl.path == PATH_DENSE) return s->k <= 15;
        if (pl.path != PATH_SPARSE || e->opt_sparse_algo == 1) return false;
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
            fn_dense_span(e, s, PackedView{sp.codes.p, sp.bad, sp.nsym});
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
        PackedView pv{spans[0].codes.p, spans[0].bad, spans[0].nsym};
        // the write pass's statistics come back with the hash path's own final readback (one sync fewer per chunk)
        e->ride_dev = st.p;
        e->ride_len = sizeof(FnStats);
        e->ride_done = false;
        try {
            sparse_chunk_range<ENC_NT2>(e, s, SymView{nullptr, 0}, &pv);
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
        if (sp.nsym) pvs.push_back(PackedView{sp.codes.p, sp.bad, sp.nsym});
    const FnStats fs3 = read_scalar<FnStats>(e, st.p);
    *need_exceptions = (fs3.packed2 >> 32) != 0;
    if (*need_exceptions) e->host_rows.failed = true;            // literal-byte rows will join the table: no streaming of packed rows
    if (!sparse_chunk_big(e, s, pvs, nullptr, hash_max)) { *need_exceptions = false; return false; }
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
