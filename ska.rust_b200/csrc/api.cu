// api.cu -- the extern "C" surface of libskagpu.so (include/skagpu.h) and its host-side plumbing:
// device context, stage timers, host->device staging of sequence batches, handle lifetime.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>

#include "../../include/skagpu.h"
#include "common.cuh"

namespace ska {

// ---- per-device state.  A host thread is bound to one device (skagpu_set_device; threads that never call it use the device
// the process chose first): one process per GPU under torchrun, or one thread per GPU inside one process (`ska --gpus N`).
namespace {
struct DevCache {
    std::mutex mu;
    std::map<void*, size_t> live;                       // block -> capacity
    std::multimap<size_t, void*> free_blocks;           // capacity -> block
    size_t cached_bytes = 0;
};
struct PendingExport { DevBuf<uint8_t> t; cudaEvent_t done; };      // a transposed alignment on its way to the host (skagpu_export_transposed_async)
struct DeviceState {
    Context ctx;
    Arena arena;
    StageTimer timer;
    DevCache cache;
    std::vector<std::unique_ptr<PendingExport>> exports;
    cudaStream_t d2h_stream = nullptr;
    uint8_t* rb_host = nullptr;                         // mapped pinned staging of the small read-backs (read_back below)
    uint8_t* rb_dev = nullptr;
    size_t rb_used = 0;
    struct RbItem { void* dst; size_t off, bytes; };
    std::vector<RbItem> rb_pending;
    void* comm = nullptr;                               // multi-GPU communicator of the thread bound to this device (comm.cu)
};
constexpr int MAX_DEVICES = 64;
DeviceState* g_dev[MAX_DEVICES];                        // created on first use, never destroyed: buffers may outlive static teardown
std::mutex g_dev_mu;
int g_default_dev = -1;
thread_local int t_dev = -1;
thread_local DeviceState* t_state = nullptr;
DeviceState& state_of(int d) {
    if (d < 0 || d >= MAX_DEVICES) throw Error("bad device index " + std::to_string(d));
    std::lock_guard<std::mutex> lock(g_dev_mu);
    if (!g_dev[d]) { g_dev[d] = new DeviceState(); g_dev[d]->ctx.device = d; }
    return *g_dev[d];
}
DeviceState& dev_state() {
    if (t_state) return *t_state;
    int d = t_dev;
    if (d < 0) {
        std::lock_guard<std::mutex> lock(g_dev_mu);
        if (g_default_dev < 0) g_default_dev = 0;
        d = g_default_dev;
    }
    t_state = &state_of(d);
    return *t_state;
}
}  // namespace
void bind_thread_to_device(int device) {
    {
        std::lock_guard<std::mutex> lock(g_dev_mu);
        if (g_default_dev < 0) g_default_dev = device;
    }
    t_dev = device;
    t_state = &state_of(device);
}
void*& comm_slot() { return dev_state().comm; }
Context& ctx() { return dev_state().ctx; }
void Context::ensure() {
    static std::mutex mu;                                   // a host may create the context on a helper thread while it parses its input
    std::lock_guard<std::mutex> lock(mu);
    if (init) { cudaSetDevice(device); return; }
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0)
        throw Error(std::string("no CUDA device available (libskagpu has no CPU fallback): ") + cudaGetErrorString(e));
    SKA_CUDA(cudaSetDevice(device));
    cudaDeviceProp p;
    SKA_CUDA(cudaGetDeviceProperties(&p, device));
    if (p.major < 10) throw Error("libskagpu is built for sm_100a (B200) only; found compute capability " +
                                  std::to_string(p.major) + "." + std::to_string(p.minor));
    sm_count = p.multiProcessorCount;
    smem_optin = p.sharedMemPerBlockOptin;
    SKA_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
    SKA_CUDA(cudaStreamCreateWithFlags(&copy_stream, cudaStreamNonBlocking));
    // keep freed blocks in the pool: repeated builds should not pay cudaMalloc again
    cudaMemPool_t pool;
    SKA_CUDA(cudaDeviceGetDefaultMemPool(&pool, device));
    uint64_t thr = UINT64_MAX;
    SKA_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr));
    init = true;
}

// ---- DevBuf's allocator: cudaMalloc'ed blocks, cached per device
namespace {
DevCache& dev_cache() { return dev_state().cache; }
}  // namespace
void* dev_alloc(size_t bytes) {
    DevCache& dc = dev_cache();
    const size_t want = (bytes + ((size_t)1 << 21) - 1) & ~(((size_t)1 << 21) - 1);      // 2 MB granules
    std::lock_guard<std::mutex> lock(dc.mu);
    auto it = dc.free_blocks.lower_bound(want);
    if (it != dc.free_blocks.end() && it->first <= want + want / 4 + ((size_t)8 << 20)) {   // close enough in size: reuse
        void* p = it->second;
        dc.cached_bytes -= it->first;
        dc.live[p] = it->first;
        dc.free_blocks.erase(it);
        return p;
    }
    void* p = nullptr;
    cudaError_t e = cudaMalloc(&p, want);
    if (e != cudaSuccess) {                               // out of memory: give the cached blocks back and try once more
        (void)cudaGetLastError();
        cudaDeviceSynchronize();
        for (auto& fb : dc.free_blocks) cudaFree(fb.second);
        dc.free_blocks.clear(); dc.cached_bytes = 0;
        e = cudaMalloc(&p, want);
    }
    if (e != cudaSuccess) throw Error(std::string("cudaMalloc of ") + std::to_string(want >> 20) + " MB failed: " + cudaGetErrorString(e));
    dc.live[p] = want;
    return p;
}
static bool cache_take_back(DevCache& dc, void* p) {
    std::lock_guard<std::mutex> lock(dc.mu);
    auto it = dc.live.find(p);
    if (it == dc.live.end()) return false;
    dc.free_blocks.emplace(it->second, p);
    dc.cached_bytes += it->second;
    dc.live.erase(it);
    return true;
}
void dev_free(void* p) {
    if (!p) return;
    if (cache_take_back(dev_cache(), p)) return;
    for (int d = 0; d < MAX_DEVICES; d++) {               // a handle released by a thread bound to another device (ska --gpus N)
        DeviceState* st;
        { std::lock_guard<std::mutex> lock(g_dev_mu); st = g_dev[d]; }
        if (st && cache_take_back(st->cache, p)) return;
    }
}

// ---- small device -> host read-backs (bucket totals, flags, tile counts) that every build waits for.  As cudaMemcpyAsync
// they go through a copy engine and queue behind the bulk transfers of a pipelined host (the next batch's upload, the
// previous alignment's download): at 8 GPUs per host an 8-byte read-back waited 12 ms.  Here a few threads store the bytes
// into mapped pinned memory instead; read_back_sync() synchronises the stream and hands them to their destinations.
constexpr size_t RB_CAP = (size_t)1 << 20;
__global__ void read_back_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, size_t bytes, int words) {
    const size_t stride = (size_t)gridDim.x * blockDim.x, i0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (words) for (size_t i = i0; i < bytes / 4; i += stride) reinterpret_cast<uint32_t*>(dst)[i] = reinterpret_cast<const uint32_t*>(src)[i];
    else for (size_t i = i0; i < bytes; i += stride) dst[i] = src[i];
}
void read_back(void* h_dst, const void* d_src, size_t bytes) {
    if (!bytes) return;
    DeviceState& st = dev_state();
    Context& c = st.ctx;
    static const bool dma = getenv("SKA_READBACK_DMA") != nullptr;        // A/B: the copy-engine form
    const size_t need = (bytes + 15) & ~(size_t)15;
    if (!dma && !st.rb_host) {
        if (cudaHostAlloc((void**)&st.rb_host, RB_CAP, cudaHostAllocMapped) != cudaSuccess ||
            cudaHostGetDevicePointer((void**)&st.rb_dev, st.rb_host, 0) != cudaSuccess) { (void)cudaGetLastError(); st.rb_host = nullptr; }
    }
    if (dma || !st.rb_host || st.rb_used + need > RB_CAP) {
        SKA_CUDA(cudaMemcpyAsync(h_dst, d_src, bytes, cudaMemcpyDeviceToHost, c.stream));
        return;
    }
    const int words = ((reinterpret_cast<uintptr_t>(d_src) | bytes) & 3) == 0;
    const unsigned grid = (unsigned)std::min<size_t>(32, (bytes + 1023) / 1024);
    read_back_kernel<<<grid, 256, 0, c.stream>>>(static_cast<const uint8_t*>(d_src), st.rb_dev + st.rb_used, bytes, words);
    SKA_CUDA(cudaGetLastError());
    st.rb_pending.push_back({h_dst, st.rb_used, bytes});
    st.rb_used += need;
}
// memset as a kernel on the library stream: cudaMemsetAsync may be carried out by a copy engine, where it waits behind the
// bulk transfers of a pipelined host exactly like the read-backs above
__global__ void dev_fill_kernel(uint8_t* __restrict__ p, uint32_t word, size_t bytes, size_t head) {
    const size_t stride = (size_t)gridDim.x * blockDim.x, i0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t n16 = (bytes - head) / 16;
    uint4* q = reinterpret_cast<uint4*>(p + head);
    for (size_t i = i0; i < n16; i += stride) q[i] = make_uint4(word, word, word, word);
    const size_t tail0 = head + n16 * 16;
    for (size_t i = i0; i < head; i += stride) p[i] = (uint8_t)word;
    for (size_t i = tail0 + i0; i < bytes; i += stride) p[i] = (uint8_t)word;
}
void dev_fill(void* p, int byte, size_t bytes) {
    if (!bytes) return;
    Context& c = ctx();
    const uint32_t w = (uint32_t)(uint8_t)byte * 0x01010101u;
    size_t head = (16 - (reinterpret_cast<uintptr_t>(p) & 15)) & 15;
    if (head > bytes) head = bytes;
    const unsigned grid = (unsigned)std::min<size_t>((size_t)c.sm_count * 8, (bytes / 16 + 255) / 256 + 1);
    dev_fill_kernel<<<grid, 256, 0, c.stream>>>(static_cast<uint8_t*>(p), w, bytes, head);
    c.launches++;
    SKA_CUDA(cudaGetLastError());
}

void read_back_abandon() {                                  // an error unwound past a pending read-back: its destination is gone
    if (!t_state) return;
    t_state->rb_pending.clear();
    t_state->rb_used = 0;
}
void read_back_sync() {
    DeviceState& st = dev_state();
    SKA_CUDA(cudaStreamSynchronize(st.ctx.stream));
    for (auto& it : st.rb_pending) memcpy(it.dst, st.rb_host + it.off, it.bytes);
    st.rb_pending.clear();
    st.rb_used = 0;
}

Arena& arena() { return dev_state().arena; }
void Arena::reset() {
    if (slabs.size() > 1) {             // the previous build had to grow: merge into one slab of the total size
        size_t total = 0;
        for (auto& s : slabs) total += s.cap;
        SKA_CUDA(cudaStreamSynchronize(ctx().stream));
        release_all();
        Slab s{nullptr, total, 0};
        SKA_CUDA(cudaMalloc((void**)&s.p, total));
        slabs.push_back(s);
    }
    for (auto& s : slabs) s.used = 0;
}
void* Arena::take(size_t bytes) {
    bytes = (bytes + 255) & ~(size_t)255;
    for (auto& s : slabs)
        if (s.cap - s.used >= bytes) { void* r = s.p + s.used; s.used += bytes; return r; }
    Slab s{nullptr, std::max<size_t>(bytes, (size_t)64 << 20), bytes};
    SKA_CUDA(cudaMalloc((void**)&s.p, s.cap));
    slabs.push_back(s);
    return s.p;
}
void Arena::release_all() {
    for (auto& s : slabs) cudaFree(s.p);
    slabs.clear();
}

StageTimer& timer() { return dev_state().timer; }
void StageTimer::clear() {
    for (auto& s : stages) { cudaEventDestroy(s.beg); cudaEventDestroy(s.end); }
    stages.clear();
}
void StageTimer::begin(const char* name) {
    if (muted) return;
    Stage s; s.name = name; s.launches = 0;
    SKA_CUDA(cudaEventCreate(&s.beg)); SKA_CUDA(cudaEventCreate(&s.end));
    SKA_CUDA(cudaEventRecord(s.beg, ctx().stream));
    launch_mark = ctx().launches;
    stages.push_back(s);
}
void StageTimer::end() {
    if (muted || stages.empty()) return;
    Stage& s = stages.back();
    SKA_CUDA(cudaEventRecord(s.end, ctx().stream));
    s.launches = ctx().launches - launch_mark;
}

__global__ void set_start_bits_kernel(const uint64_t* __restrict__ starts, uint64_t n, uint32_t* __restrict__ bits) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint64_t p = starts[i] + FRONT_PAD;
    atomicOr(&bits[p >> 5], 1u << (p & 31));
}

// K1 also counts the records by the first partition digit (default on; SKA_K1_HIST=0 keeps the two kernels apart for
// A/B runs): the partition then never reads the records for its first histogram
static bool k1_hist_enabled() { const char* e = getenv("SKA_K1_HIST"); return !e || atoi(e) != 0; }

static void check_k(uint32_t k) {
    if (k < 5 || k > 63 || (k % 2) == 0) throw Error("Invalid k-mer length");      // ska_dict.rs:342-344
}

// skagpu_build on FASTA input: the extraction kernel runs on the tiles of a chunk of samples while the next chunk is
// still crossing PCIe (copies on their own stream, K1 on the library stream behind an event per chunk).  K1 is hidden
// behind the copies, so it also counts the first partition histogram here (see build_from_batch for the trade-off).
struct K1Overlap {
    uint32_t k; bool rc; uint8_t min_qual;
    uint64_t* rec; uint32_t* valid; unsigned long long* hist; int hist_shift, hist_bits;
    uint64_t tiles_done = 0;
};
constexpr uint64_t OVERLAP_CHUNK_BYTES = 8ull << 20;

// Host -> device staging of the parsed records of all samples.
static std::unique_ptr<SeqBatch> upload(const ska_input_t* in, size_t n, double proportion_reads, K1Overlap* ov = nullptr) {
    Context& c = ctx();
    c.ensure();
    if (n == 0) throw Error("no input samples");
    if (n > 65535) throw Error("at most 65535 samples per build (16-bit sample ids)");
    std::unique_ptr<SeqBatch> b(new SeqBatch());
    b->n_samples = n;
    b->h_sample_off.assign(n + 1, 0);
    bool any_qual = false, all_qual = true;
    uint64_t n_rec_total = 0;
    for (size_t s = 0; s < n; s++) {
        if (in[s].n_rec == 0 || !in[s].rec_off || !in[s].seq) throw Error("Invalid FASTA/Q record");   // ska_dict.rs:358-360
        if (in[s].n_rec_file1 > in[s].n_rec) throw Error("n_rec_file1 > n_rec");
        b->h_sample_off[s + 1] = b->h_sample_off[s] + (in[s].rec_off[in[s].n_rec] - in[s].rec_off[0]);
        any_qual |= in[s].qual != nullptr; all_qual &= in[s].qual != nullptr;
        n_rec_total += in[s].n_rec;
    }
    // FASTA and FASTQ samples may share a build: is_reads is a per-sample property (ska_dict.rs:356-366); a FASTA sample of
    // such a batch carries quality 0xFF ("no quality": every test passes, split_kmer.rs:66-71 `None => true`) and skips the count filter
    b->has_qual = any_qual;
    if (any_qual) { b->h_is_reads.resize(n); for (size_t s = 0; s < n; s++) b->h_is_reads[s] = in[s].qual != nullptr; }
    (void)all_qual;
    b->n_pos = b->h_sample_off[n];
    const size_t padded = padded_positions(b->n_pos);
    b->seq.alloc(padded);
    // pads are 'N': invalid bases, so no window can reach outside the batch
    dev_fill(b->seq.p, 'N', FRONT_PAD);
    dev_fill(b->seq.p + FRONT_PAD + b->n_pos, 'N', padded - FRONT_PAD - b->n_pos);
    if (b->has_qual) { b->qual.alloc(padded); dev_fill(b->qual.p, '!', padded); }

    size_t step = 1;
    if (proportion_reads > 0) step = (size_t)std::llround(1.0 / proportion_reads);      // ska_dict.rs:125-129
    if (step == 0) step = 1;
    std::vector<uint64_t> starts; starts.reserve(n_rec_total + 1);
    if (ov && (b->has_qual || step != 1)) ov = nullptr;            // overlap is for plain FASTA batches
    cudaEvent_t ev_prep = nullptr;
    std::vector<cudaEvent_t> chunk_events;
    if (ov) {
        // everything K1 reads besides the bases goes first, on the library stream: record starts, sample offsets
        for (size_t s = 0; s < n; s++) {
            const uint64_t base = b->h_sample_off[s], r0 = in[s].rec_off[0];
            for (uint64_t r = 0; r < in[s].n_rec; r++) starts.push_back(base + (in[s].rec_off[r] - r0));
        }
        starts.push_back(b->n_pos);
        b->start_bits.alloc(padded / 32 + 1);
        dev_fill(b->start_bits.p, 0, b->start_bits.bytes());
        DevBuf<uint64_t> d_starts(starts.size());
        SKA_CUDA(cudaMemcpyAsync(d_starts.p, starts.data(), starts.size() * 8, cudaMemcpyHostToDevice, c.stream));
        SKA_LAUNCH(set_start_bits_kernel, (unsigned)((starts.size() + 255) / 256), 256, 0, d_starts.p, (uint64_t)starts.size(),
                   b->start_bits.p);
        b->sample_off.alloc(n + 1);
        SKA_CUDA(cudaMemcpyAsync(b->sample_off.p, b->h_sample_off.data(), (n + 1) * 8, cudaMemcpyHostToDevice, c.stream));
        SKA_CUDA(cudaEventCreateWithFlags(&ev_prep, cudaEventDisableTiming));
        SKA_CUDA(cudaEventRecord(ev_prep, c.stream));
        SKA_CUDA(cudaStreamWaitEvent(c.copy_stream, ev_prep, 0));   // the sequence buffer was allocated in library-stream order
        uint64_t chunk_bytes = 0;
        for (size_t s = 0; s < n; s++) {
            const uint64_t base = b->h_sample_off[s], r0 = in[s].rec_off[0];
            const uint64_t len = b->h_sample_off[s + 1] - base;
            if (len) SKA_CUDA(cudaMemcpyAsync(b->seq.p + FRONT_PAD + base, in[s].seq + r0, len, cudaMemcpyHostToDevice, c.copy_stream));
            chunk_bytes += len;
            if (chunk_bytes >= OVERLAP_CHUNK_BYTES || s + 1 == n) {
                cudaEvent_t ev;
                SKA_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
                SKA_CUDA(cudaEventRecord(ev, c.copy_stream));
                SKA_CUDA(cudaStreamWaitEvent(c.stream, ev, 0));
                chunk_events.push_back(ev);
                const uint64_t ready = (s + 1 == n) ? ~0ULL : k1_tiles_ready(b->h_sample_off[s + 1], K1_ASCII_TAIL);
                if (ready > ov->tiles_done) {
                    launch_k1(*b, ov->k, ov->rc, false, ov->min_qual, ov->rec, nullptr, ov->valid, 0, ov->hist, ov->hist_shift, ov->hist_bits,
                              ov->tiles_done, ready);
                    ov->tiles_done = ready;
                }
                chunk_bytes = 0;
            }
        }
        SKA_CUDA(cudaStreamSynchronize(c.stream));                 // `starts` (pageable) must outlive its copy; K1 is done
        cudaEventDestroy(ev_prep);
        for (cudaEvent_t e : chunk_events) cudaEventDestroy(e);
        return b;
    }
    for (size_t s = 0; s < n; s++) {
        const uint64_t base = b->h_sample_off[s], r0 = in[s].rec_off[0];
        const uint64_t len = b->h_sample_off[s + 1] - base;
        if (step == 1) {
            if (len) {
                SKA_CUDA(cudaMemcpyAsync(b->seq.p + FRONT_PAD + base, in[s].seq + r0, len, cudaMemcpyHostToDevice, c.stream));
                if (b->has_qual) {
                    if (in[s].qual) SKA_CUDA(cudaMemcpyAsync(b->qual.p + FRONT_PAD + base, in[s].qual + r0, len, cudaMemcpyHostToDevice, c.stream));
                    else dev_fill(b->qual.p + FRONT_PAD + base, 0xFF, len);
                }
            }
        } else {
            // --proportion-reads: keep record i of each file iff i % step == 0 (ska_dict.rs:131-141);
            // skipped records stay in the position space as runs of 'N' (they yield no window)
            if (len) dev_fill(b->seq.p + FRONT_PAD + base, 'N', len);
            for (uint64_t r = 0; r < in[s].n_rec; r++) {
                uint64_t idx_in_file = r < in[s].n_rec_file1 ? r : r - in[s].n_rec_file1;
                if (idx_in_file % step) continue;
                uint64_t o = in[s].rec_off[r] - r0, l = in[s].rec_off[r + 1] - in[s].rec_off[r];
                if (!l) continue;
                SKA_CUDA(cudaMemcpyAsync(b->seq.p + FRONT_PAD + base + o, in[s].seq + r0 + o, l, cudaMemcpyHostToDevice, c.stream));
                if (b->has_qual) {
                    if (in[s].qual) SKA_CUDA(cudaMemcpyAsync(b->qual.p + FRONT_PAD + base + o, in[s].qual + r0 + o, l, cudaMemcpyHostToDevice, c.stream));
                    else dev_fill(b->qual.p + FRONT_PAD + base + o, 0xFF, l);
                }
            }
        }
        for (uint64_t r = 0; r < in[s].n_rec; r++) starts.push_back(base + (in[s].rec_off[r] - r0));
    }
    starts.push_back(b->n_pos);                                   // end sentinel: "a record starts after the last base"
    b->start_bits.alloc(padded / 32 + 1);
    dev_fill(b->start_bits.p, 0, b->start_bits.bytes());
    DevBuf<uint64_t> d_starts(starts.size());
    SKA_CUDA(cudaMemcpyAsync(d_starts.p, starts.data(), starts.size() * 8, cudaMemcpyHostToDevice, c.stream));
    SKA_LAUNCH(set_start_bits_kernel, (unsigned)((starts.size() + 255) / 256), 256, 0, d_starts.p, (uint64_t)starts.size(),
               b->start_bits.p);
    b->sample_off.alloc(n + 1);
    SKA_CUDA(cudaMemcpyAsync(b->sample_off.p, b->h_sample_off.data(), (n + 1) * 8, cudaMemcpyHostToDevice, c.stream));
    SKA_CUDA(cudaStreamSynchronize(c.stream));                   // `starts` (pageable) must outlive the copy
    return b;
}

// Host-packed FASTA batches (skagpu_build_packed / skagpu_seqbatch_upload_packed): 2-bit codes + invalid mask cross PCIe
// (0.375 B per base instead of 1).  Every sample starts on a 32-position boundary of the batch's position space (so that its
// bytes of both streams land on byte boundaries); the gaps are invalid positions, and record ENDS are marked like record
// starts, so that "a run of exactly k valid bases at its record's end yields nothing" (split_kmer.rs:89,121) still sees
// the end of a record that is followed by a gap.
static std::unique_ptr<SeqBatch> upload_packed(const ska_input_packed_t* in, size_t n, K1Overlap* ov, bool async = false) {
    Context& c = ctx();
    c.ensure();
    if (n == 0) throw Error("no input samples");
    if (n > 65535) throw Error("at most 65535 samples per build (16-bit sample ids)");
    std::unique_ptr<SeqBatch> b(new SeqBatch());
    b->n_samples = n; b->packed = true;
    b->h_sample_off.assign(n + 1, 0);
    uint64_t n_rec_total = 0;
    for (size_t s = 0; s < n; s++) {
        if (in[s].n_rec == 0 || !in[s].rec_off || !in[s].packed2 || !in[s].invalid) throw Error("Invalid FASTA/Q record");
        if (in[s].rec_off[0] != 0) throw Error("packed samples start at base 0 (rec_off[0] must be 0)");
        const uint64_t len = in[s].rec_off[in[s].n_rec];
        b->h_sample_off[s + 1] = ((b->h_sample_off[s] + len + 31) / 32) * 32;
        n_rec_total += in[s].n_rec;
    }
    b->n_pos = b->h_sample_off[n];
    const size_t padded = padded_positions(b->n_pos);
    b->seq.alloc(padded / 4); b->qual.alloc(padded / 8);
    dev_fill(b->seq.p, 0, padded / 4);
    dev_fill(b->qual.p, 0xFF, padded / 8);          // everything invalid until a sample says otherwise
    std::vector<uint64_t>& marks = b->h_marks;                      // lives with the batch: an asynchronous upload returns before it is copied
    marks.reserve(2 * n_rec_total + 1);
    for (size_t s = 0; s < n; s++) {
        const uint64_t base = b->h_sample_off[s];
        for (uint64_t r = 0; r < in[s].n_rec; r++) {
            if (in[s].rec_off[r + 1] < in[s].rec_off[r]) throw Error("record offsets must ascend");
            marks.push_back(base + in[s].rec_off[r]);
            marks.push_back(base + in[s].rec_off[r + 1]);
        }
    }
    marks.push_back(b->n_pos);
    b->start_bits.alloc(padded / 32 + 1);
    dev_fill(b->start_bits.p, 0, b->start_bits.bytes());
    DevBuf<uint64_t> d_marks(marks.size());
    SKA_CUDA(cudaMemcpyAsync(d_marks.p, marks.data(), marks.size() * 8, cudaMemcpyHostToDevice, c.stream));
    SKA_LAUNCH(set_start_bits_kernel, (unsigned)((marks.size() + 255) / 256), 256, 0, d_marks.p, (uint64_t)marks.size(), b->start_bits.p);
    b->sample_off.alloc(n + 1);
    SKA_CUDA(cudaMemcpyAsync(b->sample_off.p, b->h_sample_off.data(), (n + 1) * 8, cudaMemcpyHostToDevice, c.stream));
    cudaEvent_t ev_prep;
    SKA_CUDA(cudaEventCreateWithFlags(&ev_prep, cudaEventDisableTiming));
    SKA_CUDA(cudaEventRecord(ev_prep, c.stream));
    SKA_CUDA(cudaStreamWaitEvent(c.copy_stream, ev_prep, 0));       // the buffers were allocated and cleared in library-stream order
    std::vector<cudaEvent_t> chunk_events;
    uint64_t chunk_bytes = 0;
    // samples that lie back to back on both sides (one pinned block on the host, whole bytes on the device) travel as ONE copy:
    // PCIe moves 1 MB pieces at 46 GB/s (34 with traffic in the other direction), 8 MB pieces at 54 (49)  [profiles/microbench/pcie_mb.cu]
    struct Run { uint8_t* dst = nullptr; const uint8_t* src = nullptr; size_t bytes = 0; } run2, run1;
    auto flush = [&](Run& r) {
        if (r.bytes) SKA_CUDA(cudaMemcpyAsync(r.dst, r.src, r.bytes, cudaMemcpyHostToDevice, c.copy_stream));
        r = Run();
    };
    auto add = [&](Run& r, uint8_t* dst, const uint8_t* src, size_t bytes) {
        if (r.bytes && r.dst + r.bytes == dst && r.src + r.bytes == src) { r.bytes += bytes; return; }
        flush(r);
        r.dst = dst; r.src = src; r.bytes = bytes;
    };
    const uint64_t chunk_limit = ov ? OVERLAP_CHUNK_BYTES / 2 : (uint64_t)16 << 20;
    for (size_t s = 0; s < n; s++) {
        const uint64_t base = b->h_sample_off[s], len = in[s].rec_off[in[s].n_rec];
        if (len) {
            add(run2, b->seq.p + (FRONT_PAD + base) / 4, in[s].packed2, (len + 3) / 4);
            add(run1, b->qual.p + (FRONT_PAD + base) / 8, in[s].invalid, (len + 7) / 8);
        }
        chunk_bytes += (len + 3) / 4 + (len + 7) / 8;
        if (chunk_bytes >= chunk_limit || s + 1 == n) {
            flush(run2); flush(run1);
            if (async) { chunk_bytes = 0; continue; }
            cudaEvent_t ev;
            SKA_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
            SKA_CUDA(cudaEventRecord(ev, c.copy_stream));
            SKA_CUDA(cudaStreamWaitEvent(c.stream, ev, 0));
            chunk_events.push_back(ev);
            if (ov) {
                const uint64_t ready = (s + 1 == n) ? ~0ULL : k1_tiles_ready(b->h_sample_off[s + 1], K1_PACKED_TAIL);
                if (ready > ov->tiles_done) {
                    launch_k1(*b, ov->k, ov->rc, false, ov->min_qual, ov->rec, nullptr, ov->valid, 0, ov->hist, ov->hist_shift, ov->hist_bits,
                              ov->tiles_done, ready);
                    ov->tiles_done = ready;
                }
            }
            chunk_bytes = 0;
        }
    }
    cudaEventDestroy(ev_prep);
    if (async) {
        // the library stream must not wait here: builds already queued keep running while this batch crosses PCIe; the
        // build that consumes the batch waits for `ready`
        SKA_CUDA(cudaEventCreateWithFlags(&b->ready, cudaEventDisableTiming));
        SKA_CUDA(cudaEventRecord(b->ready, c.copy_stream));
        for (cudaEvent_t e : chunk_events) cudaEventDestroy(e);
        return b;
    }
    SKA_CUDA(cudaStreamSynchronize(c.stream));                     // K1 (or the last copy) is done
    for (cudaEvent_t e : chunk_events) cudaEventDestroy(e);
    return b;
}

// bucket grid of a build: narrow (k <= 31) or wide keys
static PartPlan plan_for(uint32_t k, uint64_t n_pos, uint64_t n_samples) {
    return make_part_plan(k, k > 31 ? merge_prefix_bits_wide(n_pos, k) : merge_prefix_bits(n_pos, k, n_samples), n_samples);
}
static void merge_records(uint64_t* rec, const SeqBatch& b, uint32_t k, bool rc, Array& a, const PartPlan& pl, const unsigned long long* hist) {
    if (k > 31) build_array_from_records_wide(rec, b.n_pos, b.sample_off.p, (uint32_t)b.n_samples, 0, b.n_samples, k, rc, a, pl.P, hist);
    else build_array_from_records(rec, b.n_pos, b.sample_off.p, (uint32_t)b.n_samples, 0, b.n_samples, k, rc, a, pl.P, hist);
}

// records (and K1's histogram) already extracted while the batch was uploaded: partition + merge
static std::unique_ptr<Array> build_from_extracted(const SeqBatch& b, uint32_t k, bool rc, const PartPlan& pl, const K1Overlap& ov) {
    Context& c = ctx();
    std::vector<uint32_t> h_valid(b.n_samples);
    read_back(h_valid.data(), ov.valid, b.n_samples * 4);
    std::unique_ptr<Array> a(new Array());
    merge_records(ov.rec, b, k, rc, *a, pl, ov.hist);
    read_back_sync();
    for (size_t s = 0; s < b.n_samples; s++)
        if (h_valid[s] == 0) throw Error("sample " + std::to_string(s) + " has no valid sequence");   // ska_dict.rs:374-376
    return a;
}

static std::unique_ptr<Array> build_from_batch(const SeqBatch& b, uint32_t k, bool rc, uint16_t min_count, uint8_t min_qual,
                                               int qual_filter) {
    Context& c = ctx();
    check_k(k);
    if (k > 31 && b.packed) throw Error("k > 31 (128-bit split k-mers) takes ASCII input (skagpu_build)");
    if (b.ready) SKA_CUDA(cudaStreamWaitEvent(c.stream, b.ready, 0));   // asynchronous upload: the bytes may still be crossing PCIe
    arena().reset();
    if (b.has_qual) {                                              // reads: quality masks + count filter (k1q_reads.cu)
        std::unique_ptr<Array> a(new Array());
        build_array_from_reads(b, k, rc, min_count, min_qual, qual_filter, *a);
        return a;
    }
    uint64_t* rec = arena().take_n<uint64_t>(b.n_pos * (k > 31 ? 2 : 1));
    uint32_t* valid = arena().take_n<uint32_t>(b.n_samples);
    dev_fill(valid, 0, b.n_samples * 4);
    // K1 also counts the records by the first partition digit, so that the partition never reads them for its
    // histogram.  Measured on 128 x 5 Mbp: the step gets 0.67 ms shorter (partition -0.91 ms, K1 +0.24 ms: 4096 shared-memory
    // atomics per tile at 3.6 SM-cycles per warp instruction, profiles/microbench/scatter_mb2.cu).
    const PartPlan pl = plan_for(k, b.n_pos, b.n_samples);
    unsigned long long* k1_hist = nullptr;
    if (k1_hist_enabled() && k <= 31) {                      // wide keys: the partition counts its own histogram (see launch_k1)
        k1_hist = arena().take_n<unsigned long long>((size_t)K1_HIST_COPIES << pl.P1);
        dev_fill(k1_hist, 0, ((size_t)K1_HIST_COPIES << pl.P1) * 8);
    }
    timer().begin("k1_extract");
    launch_k1(b, k, rc, false, min_qual, rec, nullptr, valid, 0, k1_hist, 4 + pl.key_bits - pl.P1, pl.P1);
    timer().end();
    std::vector<uint32_t> h_valid(b.n_samples);
    read_back(h_valid.data(), valid, b.n_samples * 4);
    std::unique_ptr<Array> a(new Array());
    merge_records(rec, b, k, rc, *a, pl, k1_hist);
    read_back_sync();
    for (size_t s = 0; s < b.n_samples; s++)
        if (h_valid[s] == 0) throw Error("sample " + std::to_string(s) + " has no valid sequence");   // ska_dict.rs:374-376
    return a;
}

}  // namespace ska

// =================================================================================================
using namespace ska;
static thread_local std::string g_err;
#define API_TRY try {
#define API_CATCH                                         \
    }                                                     \
    catch (std::exception & e) { g_err = e.what(); ska::read_back_abandon(); return 1; } \
    catch (...) { g_err = "unknown error"; ska::read_back_abandon(); return 1; }

struct ska_array { Array a; };
struct ska_seqbatch { std::unique_ptr<SeqBatch> b; };
struct ska_mapping {            // RefSka after map(): hits in reference order
    uint64_t n_samples = 0, n_ref_kmers = 0, n_pos = 0; uint32_t k = 0; int key_words = 1;
    std::vector<uint64_t> pos;          // window-start positions of the hits (record-concatenated reference coordinates)
    std::vector<uint64_t> records;      // K1 records of every reference position
    DevBuf<uint8_t> rows;               // n_hits x n_samples, complemented where the reference k-mer is on the reverse strand
};
struct ska_records {            // packed records of one rank, grouped by first-level region; arena memory, valid until the next build
    uint64_t* rec = nullptr; uint64_t n = 0;
    uint32_t k = 0; bool rc = false; PartPlan plan;
    std::vector<uint64_t> off;  // 2^P1 + 1 region offsets
};

extern "C" {

const char* skagpu_last_error(void) { return g_err.c_str(); }

int skagpu_set_device(int device) {
    API_TRY
    bind_thread_to_device(device);                        // one device per host thread (one process per GPU, or one thread per GPU)
    ctx().ensure();
    return 0;
    API_CATCH
}

int skagpu_seqbatch_upload(const ska_input_t* inputs, size_t n_samples, double proportion_reads, ska_seqbatch_t** out) {
    API_TRY
    *out = nullptr;
    std::unique_ptr<ska_seqbatch> h(new ska_seqbatch());
    h->b = upload(inputs, n_samples, proportion_reads);
    *out = h.release();
    return 0;
    API_CATCH
}

int skagpu_build_from_batch(const ska_seqbatch_t* batch, uint32_t k, int rc, uint16_t min_count, uint8_t min_qual,
                            int qual_filter, ska_array_t** out) {
    API_TRY
    *out = nullptr;
    ctx().ensure();
    timer().clear();
    std::unique_ptr<Array> a = build_from_batch(*batch->b, k, rc != 0, min_count, min_qual, qual_filter);
    std::unique_ptr<ska_array> h(new ska_array());
    h->a = std::move(*a);
    *out = h.release();
    return 0;
    API_CATCH
}

void skagpu_seqbatch_free(ska_seqbatch_t* batch) { delete batch; }

int skagpu_build(const ska_input_t* inputs, size_t n_samples, uint32_t k, int rc, uint16_t min_count, uint8_t min_qual,
                 int qual_filter, double proportion_reads, ska_array_t** out) {
    API_TRY
    *out = nullptr;
    check_k(k);
    ctx().ensure();
    timer().clear();
    bool fasta = n_samples > 0 && !(proportion_reads > 0);
    uint64_t n_pos = 0;
    for (size_t s = 0; fasta && s < n_samples; s++) {
        if (inputs[s].qual || !inputs[s].rec_off || !inputs[s].n_rec) fasta = false;
        else n_pos += inputs[s].rec_off[inputs[s].n_rec] - inputs[s].rec_off[0];
    }
    if (const char* e = getenv("SKA_H2D_OVERLAP")) if (atoi(e) == 0) fasta = false;     // tests / measurements: plain upload, then K1
    std::unique_ptr<Array> a;
    if (fasta && n_pos > 0) {
        // FASTA: K1 runs chunk by chunk behind the copies (stage "h2d+k1")
        arena().reset();
        const PartPlan pl = plan_for(k, n_pos, n_samples);
        K1Overlap ov;
        ov.k = k; ov.rc = rc != 0; ov.min_qual = min_qual;
        ov.rec = arena().take_n<uint64_t>(n_pos * (k > 31 ? 2 : 1));
        ov.valid = arena().take_n<uint32_t>(n_samples);
        ov.hist = k > 31 ? nullptr : arena().take_n<unsigned long long>((size_t)K1_HIST_COPIES << pl.P1);
        ov.hist_shift = 4 + pl.key_bits - pl.P1; ov.hist_bits = pl.P1;
        dev_fill(ov.valid, 0, n_samples * 4);
        if (ov.hist) dev_fill(ov.hist, 0, ((size_t)K1_HIST_COPIES << pl.P1) * 8);
        timer().begin("h2d+k1");
        std::unique_ptr<SeqBatch> b = upload(inputs, n_samples, proportion_reads, &ov);
        timer().end();
        a = build_from_extracted(*b, k, rc != 0, pl, ov);
    } else {
        timer().begin("h2d");
        std::unique_ptr<SeqBatch> b = upload(inputs, n_samples, proportion_reads);
        timer().end();
        a = build_from_batch(*b, k, rc != 0, min_count, min_qual, qual_filter);
    }
    std::unique_ptr<ska_array> h(new ska_array());
    h->a = std::move(*a);
    *out = h.release();
    return 0;
    API_CATCH
}

// K1 behind the copies of a host-packed batch, then partition + merge
static std::unique_ptr<Array> build_packed(const ska_input_packed_t* inputs, size_t n_samples, uint32_t k, bool rc) {
    check_k(k);
    if (k > 31) throw Error("k > 31 (128-bit split k-mers): use skagpu_build (ASCII input)");
    uint64_t n_pos = 0;
    for (size_t s = 0; s < n_samples; s++) {
        if (!inputs[s].rec_off || !inputs[s].n_rec) throw Error("Invalid FASTA/Q record");
        n_pos = ((n_pos + inputs[s].rec_off[inputs[s].n_rec] + 31) / 32) * 32;
    }
    if (n_pos == 0) throw Error("sample 0 has no valid sequence");
    arena().reset();
    const PartPlan pl = make_part_plan(k, merge_prefix_bits(n_pos, k, n_samples), n_samples);
    K1Overlap ov;
    ov.k = k; ov.rc = rc; ov.min_qual = 0;
    ov.rec = arena().take_n<uint64_t>(n_pos);
    ov.valid = arena().take_n<uint32_t>(n_samples);
    ov.hist = arena().take_n<unsigned long long>((size_t)K1_HIST_COPIES << pl.P1);
    ov.hist_shift = 4 + pl.key_bits - pl.P1; ov.hist_bits = pl.P1;
    dev_fill(ov.valid, 0, n_samples * 4);
    dev_fill(ov.hist, 0, ((size_t)K1_HIST_COPIES << pl.P1) * 8);
    timer().begin("h2d+k1");
    std::unique_ptr<SeqBatch> b = upload_packed(inputs, n_samples, &ov);
    timer().end();
    return build_from_extracted(*b, k, rc, pl, ov);
}

int skagpu_build_packed(const ska_input_packed_t* inputs, size_t n_samples, uint32_t k, int rc, ska_array_t** out) {
    API_TRY
    *out = nullptr;
    ctx().ensure();
    timer().clear();
    std::unique_ptr<Array> a = build_packed(inputs, n_samples, k, rc != 0);
    std::unique_ptr<ska_array> h(new ska_array());
    h->a = std::move(*a);
    *out = h.release();
    return 0;
    API_CATCH
}

int skagpu_seqbatch_upload_packed(const ska_input_packed_t* inputs, size_t n_samples, ska_seqbatch_t** out) {
    API_TRY
    *out = nullptr;
    std::unique_ptr<ska_seqbatch> h(new ska_seqbatch());
    h->b = upload_packed(inputs, n_samples, nullptr);
    *out = h.release();
    return 0;
    API_CATCH
}

int skagpu_seqbatch_upload_packed_async(const ska_input_packed_t* inputs, size_t n_samples, ska_seqbatch_t** out) {
    API_TRY
    *out = nullptr;
    std::unique_ptr<ska_seqbatch> h(new ska_seqbatch());
    h->b = upload_packed(inputs, n_samples, nullptr, true);
    *out = h.release();
    return 0;
    API_CATCH
}

// CPU helper of the host side (no device work): what a FASTA reader calls on every record it parses.
// positions first .. first + n - 1 of the sample; `first` need not be byte aligned (records are packed back to back)
static inline void pack_one(uint8_t ch, uint64_t p, uint8_t* packed2, uint8_t* invalid) {
    const uint8_t code = (ch >> 1) & 3;                              // bit_encoding.rs:34-36
    const int sh2 = 2 * (int)(p & 3), sh1 = (int)(p & 7);
    packed2[p >> 2] = (uint8_t)((packed2[p >> 2] & ~(3u << sh2)) | (code << sh2));
    if ((ch & 0xF) == 14) invalid[p >> 3] |= (uint8_t)(1u << sh1);     // bit_encoding.rs:52-54
    else invalid[p >> 3] &= (uint8_t)~(1u << sh1);
}
void skagpu_pack_bases(const uint8_t* ascii, uint64_t n, uint64_t first, uint8_t* packed2, uint8_t* invalid) {
    uint64_t i = 0;
    while (i < n && ((first + i) & 7)) { pack_one(ascii[i], first + i, packed2, invalid); i++; }
    // 8 bases -> 2 bytes of codes + 1 byte of mask, SWAR on one 64-bit load
    for (; i + 8 <= n; i += 8) {
        uint64_t w; memcpy(&w, ascii + i, 8);
        uint64_t c = (w >> 1) & 0x0303030303030303ULL;              // one 2-bit code per byte
        c = (c | (c >> 6)) & 0x000F000F000F000FULL;                  // two codes per 16 bits
        c = (c | (c >> 12)) & 0x000000FF000000FFULL;                 // four codes per 32 bits
        const uint64_t p = first + i;
        packed2[p >> 2] = (uint8_t)c; packed2[(p >> 2) + 1] = (uint8_t)(c >> 32);
        const uint64_t t = (w & 0x0F0F0F0F0F0F0F0FULL) ^ 0x0E0E0E0E0E0E0E0EULL;           // zero byte <=> invalid base
        uint64_t z = ~(t + 0x7F7F7F7F7F7F7F7FULL) & 0x8080808080808080ULL;               // (low nibbles only: no carry between bytes)
        z = (z >> 7) * 0x0102040810204080ULL;                        // gather the 8 flags into the top byte
        invalid[p >> 3] = (uint8_t)(z >> 56);
    }
    for (; i < n; i++) pack_one(ascii[i], first + i, packed2, invalid);
}

int skagpu_array_from_host(uint32_t k, int rc, uint32_t key_bits, uint64_t n_rows, uint64_t n_samples, const void* keys,
                           const uint8_t* variants, const uint64_t* counts, ska_array_t** out) {
    API_TRY
    *out = nullptr;
    Context& c = ctx();
    c.ensure();
    if (key_bits != 64 && key_bits != 128) throw Error("key_bits must be 64 or 128");
    std::unique_ptr<ska_array> h(new ska_array());
    Array& a = h->a;
    // k > 31 arrays are held with two-word keys whatever the file's integers fit in (the reference's loader tries u64 first,
    // lib.rs:635-661; here the width follows k so that every array of one k joins with every other)
    std::vector<uint64_t> widened;
    if (k > 31 && key_bits == 64) {
        widened.assign(2 * n_rows, 0);
        for (uint64_t i = 0; i < n_rows; i++) widened[2 * i] = static_cast<const uint64_t*>(keys)[i];
        keys = widened.data(); key_bits = 128;
    }
    a.k = k; a.rc = rc != 0; a.key_bits = key_bits; a.n_rows = a.n_keys = n_rows; a.n_samples = n_samples;
    const size_t kw = key_bits / 64;
    a.keys.alloc(n_rows * kw); a.variants.alloc(n_rows * n_samples); a.counts.alloc(n_rows);
    std::vector<uint32_t> c32(n_rows);
    for (uint64_t i = 0; i < n_rows; i++) c32[i] = (uint32_t)counts[i];
    if (n_rows) {
        SKA_CUDA(cudaMemcpyAsync(a.keys.p, keys, n_rows * kw * 8, cudaMemcpyHostToDevice, c.stream));
        if (n_samples) SKA_CUDA(cudaMemcpyAsync(a.variants.p, variants, n_rows * n_samples, cudaMemcpyHostToDevice, c.stream));
        SKA_CUDA(cudaMemcpyAsync(a.counts.p, c32.data(), n_rows * 4, cudaMemcpyHostToDevice, c.stream));
    }
    SKA_CUDA(cudaStreamSynchronize(c.stream));
    *out = h.release();
    return 0;
    API_CATCH
}

int skagpu_dims(const ska_array_t* h, uint64_t* n_rows, uint64_t* n_keys, uint64_t* n_samples, uint32_t* k, int* rc,
                uint32_t* key_bits) {
    API_TRY
    if (!h) throw Error("null array");
    if (n_rows) *n_rows = h->a.n_rows;
    if (n_keys) *n_keys = h->a.n_keys;
    if (n_samples) *n_samples = h->a.n_samples;
    if (k) *k = h->a.k;
    if (rc) *rc = h->a.rc;
    if (key_bits) *key_bits = h->a.key_bits;
    return 0;
    API_CATCH
}

int skagpu_filter(ska_array_t* h, uint64_t min_count, int filter_ambig_as_missing, int filter_type, int mask_ambig,
                  int ignore_const_gaps, int update_kmers, int32_t* removed) {
    API_TRY
    if (!h) throw Error("null array");
    ctx().ensure();
    int32_t r = filter_array(h->a, min_count, filter_ambig_as_missing != 0, filter_type, mask_ambig != 0,
                             ignore_const_gaps != 0, update_kmers != 0);
    if (removed) *removed = r;
    return 0;
    API_CATCH
}

int skagpu_distance(const ska_array_t* h, double constant, int filt_ambig, ska_dist_t* out) {
    API_TRY
    if (!h) throw Error("null array");
    ctx().ensure();
    static_assert(sizeof(ska_dist_t) == sizeof(DistOut), "layout");
    distance_array(h->a, constant, filt_ambig != 0, reinterpret_cast<DistOut*>(out));
    return 0;
    API_CATCH
}

int skagpu_export(const ska_array_t* h, void* keys, uint8_t* variants, uint64_t* counts) {
    API_TRY
    if (!h) throw Error("null array");
    Context& c = ctx();
    c.ensure();
    const Array& a = h->a;
    std::vector<uint32_t> c32;
    if (keys && a.n_keys) SKA_CUDA(cudaMemcpyAsync(keys, a.keys.p, a.n_keys * (a.key_bits / 64) * 8, cudaMemcpyDeviceToHost, c.stream));
    if (variants && a.n_rows * a.n_samples)
        SKA_CUDA(cudaMemcpyAsync(variants, a.variants.p, a.n_rows * a.n_samples, cudaMemcpyDeviceToHost, c.stream));
    if (counts && a.n_rows) {
        c32.resize(a.n_rows);
        SKA_CUDA(cudaMemcpyAsync(c32.data(), a.counts.p, a.n_rows * 4, cudaMemcpyDeviceToHost, c.stream));
    }
    SKA_CUDA(cudaStreamSynchronize(c.stream));
    if (counts) for (uint64_t i = 0; i < a.n_rows; i++) counts[i] = c32[i];
    return 0;
    API_CATCH
}

int skagpu_export_transposed(const ska_array_t* h, uint8_t* out) {
    API_TRY
    if (!h) throw Error("null array");
    Context& c = ctx();
    c.ensure();
    const Array& a = h->a;
    if (a.n_rows * a.n_samples == 0) return 0;
    DevBuf<uint8_t> t(a.n_rows * a.n_samples);
    transpose_array(a, t.p);
    if (out) SKA_CUDA(cudaMemcpyAsync(out, t.p, t.bytes(), cudaMemcpyDeviceToHost, c.stream));
    SKA_CUDA(cudaStreamSynchronize(c.stream));
    return 0;
    API_CATCH
}

// ---- pipelined export: the copy of one build's alignment to the host overlaps the next build's upload and kernels.
// The bytes are pushed by a few CTAs storing straight into the (mapped) pinned host buffer, on a high-priority stream of
// their own -- NOT by the copy engine: that engine serves its queue in order, and the few-byte read-backs of the next build
// (bucket counts, row totals) would sit behind the whole 400 MB transfer while the GPU idles (measured: 21.5 ms per step
// against 14.4 ms of kernels).
__global__ void __launch_bounds__(256)
push_to_host_kernel(const uint4* __restrict__ src, uint4* __restrict__ dst, size_t n16, const uint8_t* __restrict__ src_tail, uint8_t* __restrict__ dst_tail,
                    uint32_t n_tail) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += stride) dst[i] = __ldcs(src + i);
    if (blockIdx.x == 0 && threadIdx.x < n_tail) dst_tail[threadIdx.x] = src_tail[threadIdx.x];
}

#define g_exports (dev_state().exports)
#define g_d2h_stream (dev_state().d2h_stream)
static void reap_exports(bool all) {
    Context& c = ctx();
    for (size_t i = 0; i < g_exports.size();) {
        PendingExport& e = *g_exports[i];
        if (all) SKA_CUDA(cudaEventSynchronize(e.done));
        else if (cudaEventQuery(e.done) != cudaSuccess) { i++; continue; }
        SKA_CUDA(cudaStreamWaitEvent(c.stream, e.done, 0));        // the buffer goes back to the pool in library-stream order
        cudaEventDestroy(e.done);
        g_exports.erase(g_exports.begin() + i);
    }
}

int skagpu_export_transposed_async(const ska_array_t* h, uint8_t* out) {
    API_TRY
    if (!h || !out) throw Error("null argument");
    Context& c = ctx();
    c.ensure();
    if (!g_d2h_stream) {
        int lo = 0, hi = 0;
        SKA_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
        SKA_CUDA(cudaStreamCreateWithPriority(&g_d2h_stream, cudaStreamNonBlocking, hi));   // its few CTAs get in between the build's kernels
    }
    reap_exports(false);
    while (g_exports.size() >= 3) { SKA_CUDA(cudaEventSynchronize(g_exports.front()->done)); reap_exports(false); }
    const Array& a = h->a;
    if (a.n_rows * a.n_samples == 0) return 0;
    std::unique_ptr<PendingExport> e(new PendingExport());
    e->t.alloc(a.n_rows * a.n_samples);
    transpose_array(a, e->t.p);
    cudaEvent_t ready;
    SKA_CUDA(cudaEventCreateWithFlags(&ready, cudaEventDisableTiming));
    SKA_CUDA(cudaEventRecord(ready, c.stream));
    SKA_CUDA(cudaStreamWaitEvent(g_d2h_stream, ready, 0));
    // The copy engine carries the alignment (55 GB/s).  Before the builds' small read-backs left the copy engines (read_back
    // above) this transfer delayed them, and a 32-CTA kernel storing into mapped host memory took its place; that kernel moves
    // ~10 GB/s and its stalled stores slow every other kernel by 25-30 %, which at 8 GPUs (610 MB of alignment per rank and
    // step) made it the whole step.  SKA_EXPORT_PUSH_KERNEL=1 selects it for A/B runs.
    void* out_dev = nullptr;
    const bool mapped = getenv("SKA_EXPORT_PUSH_KERNEL") && cudaHostGetDevicePointer(&out_dev, out, 0) == cudaSuccess && out_dev &&
                        ((uintptr_t)out_dev & 15) == 0;
    if (mapped) {
        const size_t n16 = e->t.bytes() / 16;
        push_to_host_kernel<<<32, 256, 0, g_d2h_stream>>>(reinterpret_cast<const uint4*>(e->t.p), reinterpret_cast<uint4*>(out_dev), n16,
                                                          e->t.p + n16 * 16, (uint8_t*)out_dev + n16 * 16, (uint32_t)(e->t.bytes() - n16 * 16));
        c.launches++;
        SKA_CUDA(cudaGetLastError());
    } else {
        (void)cudaGetLastError();
        SKA_CUDA(cudaMemcpyAsync(out, e->t.p, e->t.bytes(), cudaMemcpyDeviceToHost, g_d2h_stream));
    }
    SKA_CUDA(cudaEventCreateWithFlags(&e->done, cudaEventDisableTiming));
    SKA_CUDA(cudaEventRecord(e->done, g_d2h_stream));
    cudaEventDestroy(ready);
    g_exports.push_back(std::move(e));
    return 0;
    API_CATCH
}

int skagpu_export_wait(void) {
    API_TRY
    ctx().ensure();
    reap_exports(true);
    return 0;
    API_CATCH
}

void skagpu_free(ska_array_t* h) { delete h; }

// ---- row-sharded (multi-GPU) build: extract + first partition pass | exchange (caller) | second pass + merge ----------
int skagpu_extract_records(const ska_seqbatch_t* batch, uint32_t k, int rc, uint32_t sample_id_base, uint64_t n_samples_total,
                           int prefix_bits, ska_records_t** out) {
    API_TRY
    *out = nullptr;
    Context& c = ctx();
    c.ensure();
    check_k(k);
    if (k > 31) throw Error("k > 31 (128-bit split k-mers): the record-level multi-GPU exchange is built for k <= 31 only");
    const SeqBatch& b = *batch->b;
    if (b.has_qual) throw Error("FASTQ input (count/bloom filter, K2q) is not yet built on the device");
    if (sample_id_base + b.n_samples > n_samples_total) throw Error("sample ids exceed n_samples_total");
    timer().clear();
    arena().reset();
    std::unique_ptr<ska_records> r(new ska_records());
    r->k = k; r->rc = rc != 0;
    r->plan = make_part_plan(k, prefix_bits, n_samples_total);
    uint64_t* rec = arena().take_n<uint64_t>(b.n_pos);
    uint32_t* valid = arena().take_n<uint32_t>(b.n_samples);
    dev_fill(valid, 0, b.n_samples * 4);
    timer().begin("k1_extract");
    launch_k1(b, k, rc != 0, false, 0, rec, nullptr, valid);
    timer().end();
    std::vector<uint32_t> h_valid(b.n_samples);
    SKA_CUDA(cudaMemcpyAsync(h_valid.data(), valid, b.n_samples * 4, cudaMemcpyDeviceToHost, c.stream));
    uint64_t *d_off = nullptr, *d_tiles = nullptr;
    partition_coarse(rec, b.n_pos, b.sample_off.p, (uint32_t)b.n_samples, sample_id_base, r->plan, &r->rec, &d_off, &d_tiles);
    r->off.resize(((size_t)1 << r->plan.P1) + 1);
    SKA_CUDA(cudaMemcpyAsync(r->off.data(), d_off, r->off.size() * 8, cudaMemcpyDeviceToHost, c.stream));
    SKA_CUDA(cudaStreamSynchronize(c.stream));
    r->n = r->off.back();
    for (size_t s = 0; s < b.n_samples; s++)
        if (h_valid[s] == 0) throw Error("sample " + std::to_string(sample_id_base + s) + " has no valid sequence");
    *out = r.release();
    return 0;
    API_CATCH
}

int skagpu_records_info(const ska_records_t* r, uint64_t* n, const void** rec_dev, int* region_bits,
                        uint64_t* region_offsets /* 2^region_bits + 1, may be NULL */) {
    API_TRY
    if (!r) throw Error("null records");
    if (n) *n = r->n;
    if (rec_dev) *rec_dev = r->rec;
    if (region_bits) *region_bits = r->plan.P1;
    if (region_offsets) memcpy(region_offsets, r->off.data(), r->off.size() * 8);
    return 0;
    API_CATCH
}

void skagpu_records_free(ska_records_t* r) { delete r; }

int skagpu_merge_records_dev(void* rec_dev, int n_runs, const uint64_t* run_offsets, uint32_t region_lo, uint32_t region_hi,
                             int prefix_bits, uint64_t n_samples_total, uint32_t k, int rc, ska_array_t** out) {
    API_TRY
    *out = nullptr;
    Context& c = ctx();
    c.ensure();
    check_k(k);
    if (k > 31) throw Error("k > 31 (128-bit split k-mers): the record-level multi-GPU exchange is built for k <= 31 only");
    const PartPlan pl = make_part_plan(k, prefix_bits, n_samples_total);
    if (pl.P != prefix_bits) throw Error("prefix_bits does not match the record plan (use skagpu_merge_prefix_bits)");
    if (region_hi < region_lo || region_hi > (1ull << pl.P1)) throw Error("bad region range");
    if (n_runs < 1 || !run_offsets) throw Error("row shard needs at least one run");
    const uint32_t nr = region_hi - region_lo;
    uint64_t n_rec = 0;
    for (int w = 0; w < n_runs; w++) n_rec = std::max(n_rec, run_offsets[(size_t)w * (nr + 1) + nr]);
    if (n_rec && !rec_dev) throw Error("null record buffer");
    // the caller's buffers may have been filled on another stream (NCCL): order after everything
    SKA_CUDA(cudaDeviceSynchronize());
    arena().reset();
    std::unique_ptr<ska_array> h(new ska_array());
    timer().clear();
    build_array_from_runs((uint64_t*)rec_dev, n_rec, n_runs, run_offsets, region_lo, region_hi, pl, n_samples_total, k, rc != 0, h->a);
    *out = h.release();
    return 0;
    API_CATCH
}

// ---- row-sharded build, partial-row exchange: local arrays | exchange of rows by key range (caller) | join on the key ----
int skagpu_array_dev(const ska_array_t* h, const void** keys_dev, const void** variants_dev, const void** counts_dev) {
    API_TRY
    if (!h) throw Error("null array");
    ctx().ensure();
    SKA_CUDA(cudaStreamSynchronize(ctx().stream));          // the caller reads these on its own stream
    if (keys_dev) *keys_dev = h->a.keys.p;
    if (variants_dev) *variants_dev = h->a.variants.p;
    if (counts_dev) *counts_dev = h->a.counts.p;
    return 0;
    API_CATCH
}

int skagpu_key_cumulative(const ska_array_t* h, int region_bits, uint64_t* cum) {
    API_TRY
    if (!h || !cum) throw Error("null argument");
    ctx().ensure();
    key_cumulative(h->a, region_bits, cum);
    return 0;
    API_CATCH
}

int skagpu_merge_partial_rows_dev(const void* keys_dev, const void* variants_dev, const void* counts_dev, int n_runs,
                                  const uint64_t* run_rows, const uint32_t* run_cols, const uint32_t* run_col_base,
                                  uint64_t n_samples_total, uint32_t k, int rc, ska_array_t** out) {
    API_TRY
    *out = nullptr;
    ctx().ensure();
    check_k(k);
    if (k > 31) throw Error("k > 31 (128-bit split k-mers): the partial-row multi-GPU join is built for k <= 31 only");
    if (n_runs < 1 || !run_rows || !run_cols || !run_col_base) throw Error("row shard needs at least one run");
    // the caller's buffers may have been filled on another stream (NCCL): order after everything
    SKA_CUDA(cudaDeviceSynchronize());
    arena().reset();                                        // the stage timer keeps the stages of the local build before it
    std::unique_ptr<ska_array> h(new ska_array());
    merge_partial_rows((const uint64_t*)keys_dev, (const uint8_t*)variants_dev, (const uint32_t*)counts_dev, n_runs, run_rows, run_cols,
                       run_col_base, n_samples_total, k, rc != 0, h->a);
    *out = h.release();
    return 0;
    API_CATCH
}

// ---- the library's own communicator (comm.cu)
int skagpu_comm_unique_id(uint8_t* id) {
    API_TRY
    if (!id) throw Error("null id");
    comm_unique_id(id);
    return 0;
    API_CATCH
}
int skagpu_comm_init(const uint8_t* id, int world, int rank) {
    API_TRY
    if (!id) throw Error("null id");
    comm_init(id, world, rank);
    return 0;
    API_CATCH
}
int skagpu_comm_destroy(void) {
    API_TRY
    comm_destroy();
    return 0;
    API_CATCH
}
int skagpu_exchange_rows(const ska_array_t* local, uint64_t sample_base, uint64_t n_samples_total, ska_array_t** slab, uint64_t* stats) {
    API_TRY
    if (!local || !slab) throw Error("null argument");
    *slab = nullptr;
    ctx().ensure();
    std::unique_ptr<ska_array> h(new ska_array());
    exchange_partial_rows(local->a, sample_base, n_samples_total, h->a, stats);
    *slab = h.release();
    return 0;
    API_CATCH
}
int skagpu_allreduce_u64(uint64_t* values, size_t n) {
    API_TRY
    if (!values && n) throw Error("null argument");
    ctx().ensure();
    allreduce_u64(values, n);
    return 0;
    API_CATCH
}
int skagpu_distance_sharded(const ska_array_t* slab, double constant_local, int filt_ambig, ska_dist_t* out) {
    API_TRY
    if (!slab || !out) throw Error("null argument");
    ctx().ensure();
    distance_array(slab->a, constant_local, filt_ambig != 0, reinterpret_cast<DistOut*>(out), true);
    return 0;
    API_CATCH
}
int skagpu_plan_owners(const uint64_t* region_rows, uint32_t n_regions, int world, uint64_t* cuts) {
    API_TRY
    if (!region_rows || !cuts || world < 1) throw Error("bad argument");
    plan_owners(region_rows, n_regions, world, cuts);
    return 0;
    API_CATCH
}

int skagpu_merge_prefix_bits(uint64_t n_records_total, uint32_t k, uint64_t n_samples_total) {
    try { return make_part_plan(k, merge_prefix_bits(n_records_total, k, n_samples_total), n_samples_total).P; }
    catch (std::exception& e) { g_err = e.what(); return -1; }
}

int skagpu_distance_counters(const ska_array_t* h, int filt_ambig, uint64_t* counters) {
    API_TRY
    if (!h) throw Error("null array");
    ctx().ensure();
    distance_counters(h->a, filt_ambig != 0, counters);
    return 0;
    API_CATCH
}

// ---- joins on the key: ska merge / delete / weed / map (setops.cu)
int skagpu_merge_arrays(ska_array_t* a, ska_array_t* b, ska_array_t** out) {
    API_TRY
    *out = nullptr;
    if (!a || !b) throw Error("null array");
    ctx().ensure();
    std::unique_ptr<ska_array> h(new ska_array());
    extend_arrays(a->a, b->a, h->a);
    *out = h.release();
    return 0;
    API_CATCH
}

int skagpu_delete_samples(ska_array_t* a, const uint32_t* sample_idx, size_t n_delete) {
    API_TRY
    if (!a) throw Error("null array");
    ctx().ensure();
    delete_columns(a->a, std::vector<uint32_t>(sample_idx, sample_idx + n_delete));
    return 0;
    API_CATCH
}

int skagpu_weed(ska_array_t* a, const ska_input_t* weed_fasta, int reverse, uint64_t* n_weed_kmers) {
    API_TRY
    if (!a || !weed_fasta) throw Error("null argument");
    ctx().ensure();
    if (weed_fasta->qual) throw Error("Cannot create reference from FASTQ files");          // ska_ref.rs:204-206
    std::unique_ptr<SeqBatch> b = upload(weed_fasta, 1, 0.0);
    const uint64_t n = weed_rows(a->a, *b, reverse != 0);
    if (n_weed_kmers) *n_weed_kmers = n;
    return 0;
    API_CATCH
}

int skagpu_map(ska_array_t* a, const ska_input_t* ref_fasta, ska_mapping_t** out) {
    API_TRY
    *out = nullptr;
    if (!a || !ref_fasta) throw Error("null argument");
    ctx().ensure();
    if (ref_fasta->qual) throw Error("Cannot create reference from FASTQ files");           // ska_ref.rs:204-206
    std::unique_ptr<SeqBatch> b = upload(ref_fasta, 1, 0.0);
    std::unique_ptr<ska_mapping> m(new ska_mapping());
    m->n_samples = a->a.n_samples; m->k = a->a.k; m->key_words = a->a.k > 31 ? 2 : 1; m->n_pos = b->n_pos;
    map_reference(a->a, *b, m->pos, m->rows, m->n_ref_kmers, &m->records);
    SKA_CUDA(cudaStreamSynchronize(ctx().stream));
    if (m->n_ref_kmers == 0) throw Error("reference has no valid sequence");                 // ska_ref.rs:253-255
    *out = m.release();
    return 0;
    API_CATCH
}

int skagpu_map_dims(const ska_mapping_t* m, uint64_t* n_mapped, uint64_t* n_ref_kmers, uint64_t* n_samples, uint64_t* n_positions) {
    API_TRY
    if (!m) throw Error("null mapping");
    if (n_mapped) *n_mapped = m->pos.size();
    if (n_ref_kmers) *n_ref_kmers = m->n_ref_kmers;
    if (n_samples) *n_samples = m->n_samples;
    if (n_positions) *n_positions = m->n_pos;
    return 0;
    API_CATCH
}

int skagpu_map_export(const ska_mapping_t* m, uint64_t* middle_positions, uint8_t* variants, uint64_t* ref_records) {
    API_TRY
    if (!m) throw Error("null mapping");
    Context& c = ctx();
    c.ensure();
    const uint64_t h = (m->k - 1) / 2;
    if (middle_positions) for (size_t i = 0; i < m->pos.size(); i++) middle_positions[i] = m->pos[i] + h;
    if (variants && m->rows.n) {
        SKA_CUDA(cudaMemcpyAsync(variants, m->rows.p, m->rows.bytes(), cudaMemcpyDeviceToHost, c.stream));
        SKA_CUDA(cudaStreamSynchronize(c.stream));
    }
    if (ref_records && !m->records.empty()) memcpy(ref_records, m->records.data(), m->records.size() * 8);
    return 0;
    API_CATCH
}

void skagpu_map_free(ska_mapping_t* m) { delete m; }

int skagpu_kmer_count_hist(const ska_input_t* reads, uint32_t k, int rc, uint64_t* hist) {
    API_TRY
    if (!reads || !hist) throw Error("null argument");
    check_k(k);
    ctx().ensure();
    timer().clear();
    ska_input_t plain = *reads;
    plain.qual = nullptr;                                  // quality scores are ignored by the counter (coverage.rs:101)
    std::unique_ptr<SeqBatch> b = upload(&plain, 1, 0.0);
    kmer_count_histogram(*b, k, rc != 0, hist);
    return 0;
    API_CATCH
}

int skagpu_array_checksum(const ska_array_t* h, uint32_t col_lo, uint32_t col_hi, uint64_t* out) {
    API_TRY
    if (!h || !out) throw Error("null argument");
    ctx().ensure();
    array_checksum(h->a, col_lo, col_hi, out);
    return 0;
    API_CATCH
}

void* skagpu_host_alloc(size_t bytes) {
    void* p = nullptr;
    try { ctx().ensure(); } catch (std::exception& e) { g_err = e.what(); return nullptr; }
    if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) { g_err = "cudaMallocHost failed"; return nullptr; }
    return p;
}
void skagpu_host_free(void* p) { if (p) cudaFreeHost(p); }

int skagpu_last_timings(const char** names, float* ms, uint64_t* launches, int max_stages) {
    int n = 0;                                      // stages of the same name (one per sample of a read set) are summed
    for (auto& s : timer().stages) {
        float t = 0;
        if (cudaEventSynchronize(s.end) != cudaSuccess || cudaEventElapsedTime(&t, s.beg, s.end) != cudaSuccess) t = -1.f;
        int j = 0;
        while (j < n && strcmp(names[j], s.name) != 0) j++;
        if (j == n) {
            if (n >= max_stages) continue;
            names[n] = s.name; ms[n] = 0.f; if (launches) launches[n] = 0;
            n++;
        }
        if (t < 0 || ms[j] < 0) ms[j] = -1.f; else ms[j] += t;
        if (launches) launches[j] += s.launches;
    }
    return n;
}

uint64_t skagpu_launch_count(void) { return ctx().launches; }

int skagpu_stream_timer(int op, float* ms) {
    API_TRY
    static thread_local cudaEvent_t e0 = nullptr, e1 = nullptr;      // per host thread: a thread drives one device
    Context& c = ctx();
    c.ensure();
    if (!e0) { SKA_CUDA(cudaEventCreate(&e0)); SKA_CUDA(cudaEventCreate(&e1)); }
    if (op == 0) {
        SKA_CUDA(cudaStreamSynchronize(c.stream));
        SKA_CUDA(cudaEventRecord(e0, c.stream));
    } else {
        SKA_CUDA(cudaEventRecord(e1, c.stream));
        SKA_CUDA(cudaEventSynchronize(e1));
        float t = 0;
        SKA_CUDA(cudaEventElapsedTime(&t, e0, e1));
        if (ms) *ms = t;
    }
    return 0;
    API_CATCH
}

int skagpu_k1_extract(const ska_seqbatch_t* batch, uint32_t k, int rc, uint8_t min_qual, int qual_filter, int iters,
                      uint64_t* rec_out, uint64_t* n_positions, float* mean_ms) {
    API_TRY
    Context& c = ctx();
    c.ensure();
    check_k(k);
    const SeqBatch& b = *batch->b;
    if (k > 31 && (b.packed || b.has_qual)) throw Error("k > 31: the raw extraction entry point takes ASCII assemblies");
    if (n_positions) *n_positions = b.n_pos;
    DevBuf<uint64_t> rec(b.n_pos * (k > 31 ? 2 : 1));
    // the kernel variant the build launches: FASTA batches also fill the first partition histogram
    const PartPlan pl = plan_for(k, b.n_pos, b.n_samples);
    DevBuf<unsigned long long> hist((size_t)K1_HIST_COPIES << pl.P1);
    dev_fill(hist.p, 0, hist.bytes());
    // the extraction kernel proper; inside a build the same kernel also fills the partition histogram (stage "k1_extract")
    unsigned long long* hp = (getenv("SKA_K1_ROOFLINE_WITH_HIST") && !b.has_qual) ? hist.p : nullptr;
    const int hshift = 4 + pl.key_bits - pl.P1;
    cudaEvent_t e0, e1;
    SKA_CUDA(cudaEventCreate(&e0)); SKA_CUDA(cudaEventCreate(&e1));
    if (iters < 1) iters = 1;
    launch_k1(b, k, rc != 0, qual_filter == SKA_QUAL_STRICT, min_qual, rec.p, nullptr, nullptr, 0, hp, hshift, pl.P1);   // warm-up
    SKA_CUDA(cudaEventRecord(e0, c.stream));
    for (int i = 0; i < iters; i++) launch_k1(b, k, rc != 0, qual_filter == SKA_QUAL_STRICT, min_qual, rec.p, nullptr, nullptr, 0, hp, hshift, pl.P1);
    SKA_CUDA(cudaEventRecord(e1, c.stream));
    SKA_CUDA(cudaEventSynchronize(e1));
    float ms = 0; SKA_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    if (mean_ms) *mean_ms = ms / iters;
    if (rec_out && b.n_pos) {
        SKA_CUDA(cudaMemcpyAsync(rec_out, rec.p, rec.bytes(), cudaMemcpyDeviceToHost, c.stream));
        SKA_CUDA(cudaStreamSynchronize(c.stream));
    }
    return 0;
    API_CATCH
}

}  // extern "C"
