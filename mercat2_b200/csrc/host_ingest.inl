// Part of mc2.cu (textually included there, in this order): text upload, virtual Chunker boundaries, sample accumulation, streaming file reader.

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

// Literal rows added with mc2_sample_add_rows (tables counted elsewhere) that the sample's packed alphabet can express
// are the same k-mers as its packed rows / dense bins: move them there so that equal k-mers are summed.
static void fold_added_rows(mc2_sample* s) {
    mc2_engine* e = s->e;
    const int enc = s->plan.enc;
    if (enc < 0 || (s->plan.path != PATH_DENSE && s->plan.path != PATH_SPARSE) || s->k * enc_bits(enc) > 64) return;
    for (auto& part : s->wide) {
        if (!part.added || !part.n) continue;
        const u64 n = part.n;
        DBuf<u32> flag(e, n);
        DBuf<u64> pos(e, n);
        DBuf<ull> total(e, 1);
        LAUNCH(e, rows_flag_kernel, (unsigned)div_up(n, 256), 256, 0, (const u8*)part.rows.p, n, s->k, enc, flag.p);
        dev_exclusive_scan<u32, u64>(e, flag.p, pos.p, n, total.p);
        const u64 nf = (u64)read_scalar<ull>(e, total.p);
        if (!nf) continue;
        const bool dense = s->plan.path == PATH_DENSE;
        FastPart fp;
        if (!dense) {
            fp.n = nf;
            fp.sorted = false;
            fp.keys.alloc(e, nf);
            fp.counts.alloc(e, nf);
        }
        WidePart rest;
        rest.n = n - nf;
        rest.sorted = false;
        rest.rows.alloc(e, std::max<u64>(1, rest.n * (u64)s->k));
        rest.counts.alloc(e, std::max<u64>(1, rest.n));
        LAUNCH(e, rows_split_kernel, (unsigned)div_up(n, 256), 256, 0, (const u8*)part.rows.p, (const u64*)part.counts.p, n, s->k, enc,
               (const u32*)flag.p, (const u64*)pos.p, fp.keys.p, fp.counts.p, dense ? (unsigned long long*)s->dense_sample.p : nullptr,
               rest.rows.p, rest.counts.p);
        if (!dense) s->fast.push_back(std::move(fp));
        part = std::move(rest);
    }
}

static mc2_table* sample_finish(mc2_sample* s) {
    mc2_engine* e = s->e;
    PhaseTimer pt(e);
    std::unique_ptr<mc2_table> t(new mc2_table);
    t->e = e;
    t->k = s->k;
    t->enc = s->plan.enc < 0 ? ENC_NT2 : s->plan.enc;
    fold_added_rows(s);
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
