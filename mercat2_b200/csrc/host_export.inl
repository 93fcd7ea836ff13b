// Part of mc2.cu (textually included there, in this order): sample finish is above; export of tables: host arrays and TSV text.

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
