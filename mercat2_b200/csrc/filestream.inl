// Streaming file source for mc2_sample_add_file (included by mc2.cu).
//
// Replaces the reader side of lib/mercat2_kmers.py:47 (`gzip.open(file,'rt') if suffix=='.gz' else open(file)`) and
// of the Chunker (lib/mercat2_Chunker.py:33-37, which re-reads and re-writes the whole file): a reader thread fills
// a small pool of pinned 32 MiB buffers from the file (read(2), or zlib inflate for .gz); the engine thread sends
// every filled buffer to the device with cudaMemcpyAsync on the copy stream while the reader already fills the
// next one, and the compute stream chunks + counts the text that has arrived.  The device side keeps only a
// sliding window of the text (a few pieces), not the file.

#include <condition_variable>
#include <fcntl.h>
#include <mutex>
#include <sys/stat.h>
#include <thread>
#include <unistd.h>
#include <zlib.h>

struct FileSource {
    int fd = -1;
    gzFile gz = nullptr;
    std::string path;
    ~FileSource() {
        if (gz) gzclose(gz);
        else if (fd >= 0) close(fd);
    }
    void open_path(const char* p, bool gunzip) {
        path = p;
        fd = open(p, O_RDONLY);
        if (fd < 0) throw Mc2Error(MC2_ERR_IO, std::string("cannot open ") + p);
        if (gunzip) {
            gz = gzdopen(fd, "rb");
            if (!gz) { close(fd); fd = -1; throw Mc2Error(MC2_ERR_IO, std::string("cannot read ") + p + " as gzip"); }
            gzbuffer(gz, 1u << 20);
        }
    }
    // fills up to cap bytes; returns the number read (0 = end of file); throws on error
    size_t read_some(u8* dst, size_t cap) {
        size_t got = 0;
        while (got < cap) {
            long n;
            if (gz) {
                n = gzread(gz, dst + got, (unsigned)std::min<size_t>(cap - got, 1u << 30));
                if (n < 0) {
                    int err = 0;
                    const char* msg = gzerror(gz, &err);
                    throw Mc2Error(MC2_ERR_IO, path + ": " + (msg ? msg : "gzip error"));
                }
            } else {
                n = (long)::read(fd, dst + got, cap - got);
                if (n < 0) throw Mc2Error(MC2_ERR_IO, path + ": read error");
            }
            if (n == 0) {
                if (gz) {                       // a truncated stream ends without a read error: ask zlib (Python's gzip raises EOFError)
                    int err = Z_OK;
                    gzerror(gz, &err);
                    if (err != Z_OK && err != Z_STREAM_END) throw Mc2Error(MC2_ERR_IO, path + ": truncated or damaged gzip stream");
                }
                break;
            }
            got += (size_t)n;
        }
        return got;
    }
};

// Reader threads + pinned buffer pool.  Piece p (32 MiB of the file's text) goes to buffer p % NBUF; thread j fills
// the pieces p = j (mod T).  A plain file is read by T = 4 threads with pread (page-cache copies run in parallel); a
// gzip stream is sequential (T = 1).  No CUDA call happens on a reader thread.
struct PinnedReader {
    static constexpr int NBUF = 4;
    static constexpr u64 PIECE_MAX = 32ull << 20;             // size of a pinned buffer
    u64 PIECE;                                                // bytes per piece (engine option file_piece_bytes; tests shrink it)
    FileSource* src;
    int nthreads;
    u8* buf[NBUF] = {nullptr, nullptr, nullptr, nullptr};
    size_t len[NBUF] = {0, 0, 0, 0};
    int state[NBUF] = {0, 0, 0, 0};          // 0 free, 1 filled, 2 in flight to the device
    u64 next_piece = 0;                      // consumer side
    bool done = false;                       // the consumer has seen the last piece
    bool failed = false, stop = false;
    std::string error;
    int error_code = MC2_ERR_IO;
    std::mutex mu;
    std::condition_variable cv;
    std::vector<std::thread> th;

    PinnedReader(mc2_engine* e, FileSource* s) : PIECE(std::min<u64>(PIECE_MAX, std::max<u64>(4096, e->opt_file_piece))), src(s), nthreads(s->gz ? 1 : NBUF) {
        for (int i = 0; i < NBUF; ++i) {                         // the pinned pool lives with the engine (allocated once)
            if (!e->file_pin[i]) CUDA_CHECK(cudaMallocHost((void**)&e->file_pin[i], PIECE_MAX));
            buf[i] = e->file_pin[i];
        }
        for (int j = 0; j < nthreads; ++j) th.emplace_back([this, j] { run(j); });
    }
    ~PinnedReader() {
        {
            std::lock_guard<std::mutex> lk(mu);
            stop = true;
        }
        cv.notify_all();
        for (auto& t : th) if (t.joinable()) t.join();
    }
    size_t fill(u64 piece, u8* dst) {
        if (src->gz) return src->read_some(dst, (size_t)PIECE);
        size_t got = 0;
        while (got < PIECE) {
            const ssize_t n = pread(src->fd, dst + got, PIECE - got, (off_t)(piece * PIECE + got));
            if (n < 0) throw Mc2Error(MC2_ERR_IO, src->path + ": read error");
            if (n == 0) break;
            got += (size_t)n;
        }
        return got;
    }
    void run(int j) {
        try {
            for (u64 piece = (u64)j;; piece += (u64)nthreads) {
                const int i = (int)(piece % NBUF);
                {
                    std::unique_lock<std::mutex> lk(mu);
                    cv.wait(lk, [&] { return stop || state[i] == 0; });
                    if (stop) return;
                }
                const size_t n = fill(piece, buf[i]);
                {
                    std::lock_guard<std::mutex> lk(mu);
                    len[i] = n;
                    state[i] = 1;
                }
                cv.notify_all();
                if (n < PIECE) return;                           // the last piece (possibly empty)
            }
        } catch (const Mc2Error& err) {
            std::lock_guard<std::mutex> lk(mu);
            failed = true;
            error = err.what();
            error_code = err.code;
            cv.notify_all();
        } catch (const std::exception& err) {
            std::lock_guard<std::mutex> lk(mu);
            failed = true;
            error = err.what();
            cv.notify_all();
        }
    }
    int claim_locked() {
        const int i = (int)(next_piece % NBUF);
        if (len[i] < PIECE) done = true;                         // short piece: the file ends here
        if (len[i] == 0) { state[i] = 0; return -1; }
        state[i] = 2;
        ++next_piece;
        return i;
    }
    // next piece in file order (blocks); returns -1 at end of file
    int take() {
        std::unique_lock<std::mutex> lk(mu);
        if (done) return -1;
        const int i = (int)(next_piece % NBUF);
        cv.wait(lk, [&] { return failed || state[i] == 1; });
        if (failed) throw Mc2Error(error_code, error);
        return claim_locked();
    }
    // like take(), but returns -2 instead of blocking when the next piece is not ready yet
    int try_take() {
        std::lock_guard<std::mutex> lk(mu);
        if (failed) throw Mc2Error(error_code, error);
        if (done) return -1;
        if (state[next_piece % NBUF] != 1) return -2;
        return claim_locked();
    }
    void give_back(int i) {
        {
            std::lock_guard<std::mutex> lk(mu);
            state[i] = 0;
        }
        cv.notify_all();
    }
};

// Sliding window of the file's text on the device.
struct StreamText {
    mc2_engine* e;
    PinnedReader* rd;
    DBuf<u8> dev;
    u64 cap = 0, head = 0, filled = 0;        // valid text: dev[head, filled)
    bool eof = false;
    u64 total = 0;                            // bytes read from the source so far
    cudaEvent_t slot_ev[PinnedReader::NBUF] = {nullptr, nullptr, nullptr, nullptr};
    bool slot_busy[PinnedReader::NBUF] = {false, false, false, false};
    cudaEvent_t last_ev = nullptr;

    StreamText(mc2_engine* e_, PinnedReader* r, u64 initial) : e(e_), rd(r) {
        cap = std::max<u64>(initial, 4 * rd->PIECE);
        dev.alloc(e, cap + 16);
        for (auto& ev : slot_ev) CUDA_CHECK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
        CUDA_CHECK(cudaEventCreateWithFlags(&last_ev, cudaEventDisableTiming));
        CUDA_CHECK(cudaStreamSynchronize(e->stream));             // the buffer exists before the copy stream writes it
    }
    ~StreamText() {
        cudaStreamSynchronize(e->copy_stream);
        for (auto ev : slot_ev) if (ev) cudaEventDestroy(ev);
        if (last_ev) cudaEventDestroy(last_ev);
    }
    const u8* ptr() const { return dev.p + head; }
    u64 len() const { return filled - head; }
    void reclaim(bool wait) {
        for (int i = 0; i < PinnedReader::NBUF; ++i) {
            if (!slot_busy[i]) continue;
            if (wait) CUDA_CHECK(cudaEventSynchronize(slot_ev[i]));
            else if (cudaEventQuery(slot_ev[i]) != cudaSuccess) { cudaGetLastError(); continue; }
            slot_busy[i] = false;
            rd->give_back(i);
        }
    }
    // room for `more` bytes behind `filled`: compact the window to the front of the buffer, or grow the buffer
    void make_room(u64 more) {
        if (filled + more <= cap) return;
        CUDA_CHECK(cudaStreamSynchronize(e->stream));             // nothing reads the old window any more
        CUDA_CHECK(cudaStreamSynchronize(e->copy_stream));
        const u64 keep = filled - head;
        if (keep + more <= cap && head >= keep) {                  // non-overlapping move to the front
            if (keep) CUDA_CHECK(cudaMemcpyAsync(dev.p, dev.p + head, keep, cudaMemcpyDeviceToDevice, e->copy_stream));
        } else {
            const u64 ncap = std::max<u64>(2 * cap, keep + more + rd->PIECE);
            DBuf<u8> bigger(e, ncap + 16);
            CUDA_CHECK(cudaStreamSynchronize(e->stream));
            if (keep) CUDA_CHECK(cudaMemcpyAsync(bigger.p, dev.p + head, keep, cudaMemcpyDeviceToDevice, e->copy_stream));
            CUDA_CHECK(cudaStreamSynchronize(e->copy_stream));
            dev = std::move(bigger);
            cap = ncap;
        }
        CUDA_CHECK(cudaStreamSynchronize(e->copy_stream));
        head = 0;
        filled = keep;
    }
    // pull one more piece from the reader (false at end of file)
    bool pull() {
        if (eof) return false;
        reclaim(false);
        int i = rd->try_take();
        if (i == -2) {                       // nothing filled yet: hand every buffer back before blocking, or the
            reclaim(true);                   // reader could be waiting for one that we still hold
            i = rd->take();
        }
        if (i < 0) { eof = true; return false; }
        const u64 n = rd->len[i];
        make_room(n);
        CUDA_CHECK(cudaMemcpyAsync(dev.p + filled, rd->buf[i], n, cudaMemcpyHostToDevice, e->copy_stream));
        CUDA_CHECK(cudaEventRecord(slot_ev[i], e->copy_stream));
        slot_busy[i] = true;
        filled += n;
        total += n;
        e->h2d_bytes += n;
        return true;
    }
    // the window holds at least `want` bytes (or everything up to the end of the file); the compute stream waits
    // for the copies it is about to read
    void ensure(u64 want) {
        while (len() < want && pull()) {}
        CUDA_CHECK(cudaEventRecord(last_ev, e->copy_stream));
        CUDA_CHECK(cudaStreamWaitEvent(e->stream, last_ev, 0));
    }
    void consume(u64 n) { head += n; }
};

static void sample_add_file(mc2_sample* s, const char* path, bool gunzip, u64 chunk_bytes, u64* n_chunks, u64* text_bytes) {
    mc2_engine* e = s->e;
    FileSource src;
    src.open_path(path, gunzip);
    PinnedReader reader(e, &src);
    struct stat stt;
    u64 disk = 0;
    if (stat(path, &stt) == 0) disk = (u64)stt.st_size;
    const u64 margin = std::min<u64>(4ull << 20, reader.PIECE);
    // window: a chunked file needs a few pieces; an unchunked one its whole text (size known for plain files)
    const u64 initial = chunk_bytes ? 2 * (chunk_bytes + margin) + 2 * reader.PIECE : (gunzip ? disk * 2 : disk) + reader.PIECE;   // (a gzip text of unknown size: the window grows by doubling)
    StreamText st(e, &reader, initial);
    u64 pieces = 0;
    while (true) {
        u64 n;
        if (chunk_bytes) {
            u64 want = chunk_bytes + margin;
            while (true) {
                st.ensure(want);
                if (st.len() == 0) { n = 0; break; }
                const std::vector<u64> wb = chunk_bounds(e, st.ptr(), st.len(), chunk_bytes);
                if (wb.size() >= 2) { n = wb[1]; break; }
                if (st.eof) { n = st.len(); break; }
                want = st.len() + chunk_bytes;
            }
        } else {
            st.ensure(~0ull >> 1);
            n = st.len();
        }
        if (n == 0 && st.eof) {
            if (pieces == 0) { count_chunk(e, s, st.ptr(), 0); pieces = 1; }      // an empty file is one (empty) piece
            break;
        }
        count_chunk(e, s, st.ptr(), n);
        st.consume(n);
        ++pieces;
        if (st.eof && st.len() == 0) break;
    }
    CUDA_CHECK(cudaStreamSynchronize(e->stream));
    st.reclaim(true);
    if (n_chunks) *n_chunks = pieces;
    if (text_bytes) *text_bytes = st.total;
}
