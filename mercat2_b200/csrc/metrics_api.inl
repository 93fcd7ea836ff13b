// C ABI for K6 (included at the end of mc2.cu)
struct mc2_metrics {
    std::vector<RecordOut> rec;     // empty sequences already dropped
};

extern "C" {

int mc2_protein_metrics(mc2_engine* e, const void* text, uint64_t nbytes, int space, mc2_metrics** out) {
    API_BEGIN
    if (!e || !out) throw Mc2Error(MC2_ERR_INVALID, "NULL argument");
    if (!text && nbytes) throw Mc2Error(MC2_ERR_INVALID, "text is NULL");
    CUDA_CHECK(cudaSetDevice(e->device));
    std::unique_ptr<mc2_metrics> m(new mc2_metrics);
    if (nbytes) {
        DBuf<u8> holder;
        const u8* d = to_device(e, text, nbytes, space, holder);
        const u64 nt = div_up(nbytes, 256);
        DBuf<u32> tc(e, nt);
        DBuf<u64> to(e, nt);
        LAUNCH(e, mt_lines_kernel<false>, (unsigned)nt, 256, 0, d, (u64)nbytes, tc.p, (const u64*)nullptr, (u64*)nullptr);
        const u64 nlines = offsets_from_counts(e, tc.p, to.p, nt);
        DBuf<u64> line_off(e, nlines);
        LAUNCH(e, mt_lines_kernel<true>, (unsigned)nt, 256, 0, d, (u64)nbytes, tc.p, (const u64*)to.p, line_off.p);
        DBuf<LineStat> ls(e, nlines);
        LAUNCH(e, mt_line_stats_kernel, (unsigned)div_up(nlines, 128), 128, 0, d, (u64)nbytes, (const u64*)line_off.p, nlines, ls.p);
        const u64 nt2 = div_up(nlines, 256);
        DBuf<u32> tc2(e, nt2);
        DBuf<u64> to2(e, nt2);
        LAUNCH(e, mt_headers_kernel<false>, (unsigned)nt2, 256, 0, (const LineStat*)ls.p, nlines, tc2.p, (const u64*)nullptr, (u64*)nullptr);
        const u64 nrec = offsets_from_counts(e, tc2.p, to2.p, nt2);
        if (nrec) {
            DBuf<u64> hdr(e, nrec);
            LAUNCH(e, mt_headers_kernel<true>, (unsigned)nt2, 256, 0, (const LineStat*)ls.p, nlines, tc2.p, (const u64*)to2.p, hdr.p);
            DBuf<RecordOut> ro(e, nrec);
            LAUNCH(e, mt_records_kernel, (unsigned)div_up(nrec, 128), 128, 0, (const LineStat*)ls.p, nlines, (const u64*)hdr.p, nrec, ro.p);
            std::vector<RecordOut> all(nrec);
            d2h(e, all.data(), ro.p, nrec);
            m->rec.reserve(nrec);
            for (auto& r : all) if (r.status != 255) m->rec.push_back(r);
        }
        CUDA_CHECK(cudaStreamSynchronize(e->stream));
    }
    *out = m.release();
    API_END
}

int mc2_sequence_metrics(mc2_engine* e, const void* seqs, const uint64_t* offsets, uint64_t nseq, mc2_metrics** out) {
    API_BEGIN
    if (!e || !out || !offsets) throw Mc2Error(MC2_ERR_INVALID, "NULL argument");
    CUDA_CHECK(cudaSetDevice(e->device));
    std::unique_ptr<mc2_metrics> m(new mc2_metrics);
    if (nseq) {
        const u64 total = offsets[nseq];
        DBuf<u8> dseq(e, total + 16);
        DBuf<u64> doff(e, nseq + 1);
        if (total) CUDA_CHECK(cudaMemcpyAsync(dseq.p, seqs, total, cudaMemcpyHostToDevice, e->stream));
        CUDA_CHECK(cudaMemcpyAsync(doff.p, offsets, (nseq + 1) * 8, cudaMemcpyHostToDevice, e->stream));
        e->h2d_bytes += total + (nseq + 1) * 8;
        DBuf<RecordOut> ro(e, nseq);
        LAUNCH(e, mt_sequences_kernel, (unsigned)div_up(nseq, 128), 128, 0, (const u8*)dseq.p, (const u64*)doff.p, (u64)nseq, ro.p);
        m->rec.resize(nseq);                      // keeps empty sequences (status 255): one row per input
        d2h(e, m->rec.data(), ro.p, nseq);
    }
    *out = m.release();
    API_END
}

uint64_t mc2_metrics_records(const mc2_metrics* m) { return m ? m->rec.size() : 0; }

int mc2_metrics_export(const mc2_metrics* m, uint64_t* header_off, uint32_t* header_len, uint64_t* length, double* pi,
                       double* mw, double* hydro, uint8_t* status) {
    API_BEGIN
    if (!m) throw Mc2Error(MC2_ERR_INVALID, "metrics is NULL");
    for (size_t i = 0; i < m->rec.size(); ++i) {
        const RecordOut& r = m->rec[i];
        if (header_off) header_off[i] = r.hdr_off;
        if (header_len) header_len[i] = r.hdr_len;
        if (length) length[i] = r.length;
        if (pi) pi[i] = r.pi;
        if (mw) mw[i] = r.mw;
        if (hydro) hydro[i] = r.hydro;
        if (status) status[i] = (uint8_t)r.status;
    }
    API_END
}

void mc2_metrics_free(mc2_metrics* m) { delete m; }

}  // extern "C"
