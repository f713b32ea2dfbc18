// comm.cu -- the multi-GPU exchange behind the C ABI: the library owns the NCCL communicator, so a
// host without torch (the C++ `ska --gpus N`, a Rust binding) shards a build over the GPUs of a node.
//
// The reference is single-node shared memory; its build_and_merge (src/merge_ska_dict.rs:354-417) has one point where data
// of different samples meet: rows of the merged dictionary are keyed by split k-mer.  Every rank (a host thread or process
// bound to one GPU) builds the array of ITS samples; the ranks' cumulative key histograms over 2^12 key ranges are
// all-gathered and cut into `world` contiguous owner ranges of equal row count; ONE grouped exchange (keys, variant rows and
// counts of every peer inside a single ncclGroupStart/End, i.e. one fused send/recv kernel over NVLink) moves every partial
// row to its owner; the owner joins the runs on the key (rowmerge.cu).  The same plan, transport aside, as distributed.py's
// "rows" strategy, which the tests hold against the single-GPU array.
//
// NCCL is loaded with dlopen at skagpu_comm_init: libskagpu.so has no link-time dependency on it, a process that already
// holds a libnccl.so.2 (torch's) shares it, and single-GPU users never load it.
#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <cstring>
#include <mutex>
#include <vector>

#include "common.cuh"

namespace ska {

namespace {
struct NcclApi {
    void* lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};
NcclApi& nccl() {
    static NcclApi api;
    static std::mutex mu;
    std::lock_guard<std::mutex> lock(mu);
    if (api.lib) return api;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names) if ((api.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL))) break;
    if (!api.lib) throw Error(std::string("multi-GPU needs NCCL: ") + dlerror());
    auto sym = [&](const char* n) { void* p = dlsym(api.lib, n); if (!p) throw Error(std::string("NCCL symbol missing: ") + n); return p; };
    api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(sym("ncclGetUniqueId"));
    api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(sym("ncclCommInitRank"));
    api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(sym("ncclCommDestroy"));
    api.AllGather = reinterpret_cast<decltype(api.AllGather)>(sym("ncclAllGather"));
    api.AllReduce = reinterpret_cast<decltype(api.AllReduce)>(sym("ncclAllReduce"));
    api.Send = reinterpret_cast<decltype(api.Send)>(sym("ncclSend"));
    api.Recv = reinterpret_cast<decltype(api.Recv)>(sym("ncclRecv"));
    api.GroupStart = reinterpret_cast<decltype(api.GroupStart)>(sym("ncclGroupStart"));
    api.GroupEnd = reinterpret_cast<decltype(api.GroupEnd)>(sym("ncclGroupEnd"));
    api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(sym("ncclGetErrorString"));
    return api;
}
#define SKA_NCCL(call)                                                                                  \
    do {                                                                                                \
        ncclResult_t r_ = (call);                                                                       \
        if (r_ != ncclSuccess) throw ska::Error(std::string(#call) + " failed: " + nccl().GetErrorString(r_)); \
    } while (0)

struct Comm { ncclComm_t comm = nullptr; int world = 1, rank = 0; };
Comm& comm_of_thread() {
    Comm* c = static_cast<Comm*>(comm_slot());
    if (!c) throw Error("no communicator on this device: call skagpu_comm_init first");
    return *c;
}
}  // namespace

void comm_unique_id(uint8_t* id /* SKAGPU_COMM_ID_BYTES */) {
    static_assert(sizeof(ncclUniqueId) == 128, "id size");
    ncclUniqueId u;
    SKA_NCCL(nccl().GetUniqueId(&u));
    memcpy(id, &u, sizeof(u));
}

void comm_init(const uint8_t* id, int world, int rank) {
    if (world < 1 || rank < 0 || rank >= world) throw Error("bad communicator shape");
    ctx().ensure();
    if (comm_slot()) throw Error("this device already has a communicator");
    ncclUniqueId u;
    memcpy(&u, id, sizeof(u));
    Comm* c = new Comm();
    c->world = world; c->rank = rank;
    SKA_NCCL(nccl().CommInitRank(&c->comm, world, u, rank));
    comm_slot() = c;
}

void comm_destroy() {
    Comm* c = static_cast<Comm*>(comm_slot());
    if (!c) return;
    cudaStreamSynchronize(ctx().stream);
    nccl().CommDestroy(c->comm);
    delete c;
    comm_slot() = nullptr;
}

int comm_world() { return comm_of_thread().world; }
int comm_rank() { return comm_of_thread().rank; }

// Cut `n_regions` key ranges (ascending) into `world` contiguous owner ranges of near-equal row count.  counts: rows per
// region summed over the ranks.  -> cuts[world + 1] region indices.  The arithmetic of distributed.plan_owners.
void plan_owners(const uint64_t* counts, uint32_t n_regions, int world, uint64_t* cuts) {
    std::vector<uint64_t> cum(n_regions + 1, 0);
    for (uint32_t i = 0; i < n_regions; i++) cum[i + 1] = cum[i] + counts[i];
    const uint64_t total = cum[n_regions];
    cuts[0] = 0;
    for (int g = 1; g < world; g++) {
        const unsigned __int128 t = (unsigned __int128)total * (unsigned)g / (unsigned)world;
        const uint64_t target = (uint64_t)t;
        uint64_t b = (uint64_t)(std::lower_bound(cum.begin(), cum.end(), target) - cum.begin());   // searchsorted(side="left")
        b = std::min<uint64_t>(std::max<uint64_t>(b, cuts[g - 1]), n_regions);
        cuts[g] = b;
    }
    cuts[world] = n_regions;
}

void allreduce_u64(uint64_t* h_vals, size_t n) {
    Comm& cm = comm_of_thread();
    Context& c = ctx();
    if (cm.world == 1 || n == 0) return;
    DevBuf<uint64_t> d(n);
    SKA_CUDA(cudaMemcpyAsync(d.p, h_vals, n * 8, cudaMemcpyHostToDevice, c.stream));
    SKA_NCCL(nccl().AllReduce(d.p, d.p, n, ncclUint64, ncclSum, cm.comm, c.stream));
    SKA_CUDA(cudaMemcpyAsync(h_vals, d.p, n * 8, cudaMemcpyDeviceToHost, c.stream));
    SKA_CUDA(cudaStreamSynchronize(c.stream));
}

// local: this rank's array (its own samples: columns [sample_base, sample_base + local.n_samples) of the whole job).
// -> slab: the rows of this rank's key range with all n_samples_total columns.  Collective.
void exchange_partial_rows(const Array& local, uint64_t sample_base, uint64_t n_samples_total, Array& slab, uint64_t* stats /* [4] or null */) {
    Comm& cm = comm_of_thread();
    Context& c = ctx();
    const int W = cm.world, me = cm.rank;
    if (local.key_bits != 64) throw Error("k > 31 (128-bit split k-mers): the partial-row multi-GPU join is built for k <= 31 only");
    if (local.n_keys != local.n_rows) throw Error("array keys and rows are out of step (filtered without update_kmers)");
    const int region_bits = std::min(12, 2 * ((int)local.k - 1));
    const uint32_t R = 1u << region_bits;
    const size_t T = (size_t)R + 3;                                    // per rank: R + 1 cumulative counts, sample_base, n_local
    std::vector<uint64_t> mine(T), table((size_t)W * T);
    key_cumulative(local, region_bits, mine.data());
    mine[R + 1] = sample_base; mine[R + 2] = local.n_samples;
    {
        DevBuf<uint64_t> d_mine(T), d_all((size_t)W * T);
        SKA_CUDA(cudaMemcpyAsync(d_mine.p, mine.data(), T * 8, cudaMemcpyHostToDevice, c.stream));
        SKA_NCCL(nccl().AllGather(d_mine.p, d_all.p, T, ncclUint64, cm.comm, c.stream));
        read_back(table.data(), d_all.p, (size_t)W * T * 8);
        read_back_sync();
    }
    // owners: identical arithmetic on every rank
    std::vector<uint64_t> counts(R, 0), cuts(W + 1);
    for (int r = 0; r < W; r++)
        for (uint32_t i = 0; i < R; i++) counts[i] += table[r * T + i + 1] - table[r * T + i];
    plan_owners(counts.data(), R, W, cuts.data());
    std::vector<uint64_t> send_rows(W), recv_rows(W), send_first(W);
    std::vector<uint32_t> cols(W), col_base(W);
    uint64_t cols_sum = 0;
    for (int r = 0; r < W; r++) {
        send_first[r] = table[me * T + cuts[r]];
        send_rows[r] = table[me * T + cuts[r + 1]] - table[me * T + cuts[r]];
        recv_rows[r] = table[r * T + cuts[me + 1]] - table[r * T + cuts[me]];
        col_base[r] = (uint32_t)table[r * T + R + 1];
        cols[r] = (uint32_t)table[r * T + R + 2];
        cols_sum += cols[r];
    }
    if (cols_sum != n_samples_total) throw Error("the ranks' samples do not add up to n_samples_total");
    uint64_t n_recv = 0, recv_var_bytes = 0;
    for (int r = 0; r < W; r++) { n_recv += recv_rows[r]; recv_var_bytes += recv_rows[r] * cols[r]; }
    DevBuf<uint64_t> r_keys(std::max<uint64_t>(n_recv, 1));
    DevBuf<uint32_t> r_cnt(std::max<uint64_t>(n_recv, 1));
    DevBuf<uint8_t> r_var(std::max<uint64_t>(recv_var_bytes, 1));
    const uint64_t nl = local.n_samples;
    // one grouped exchange: three buffers to and from every peer; the own share is a device copy
    SKA_NCCL(nccl().GroupStart());
    uint64_t ro = 0, vo = 0;
    for (int r = 0; r < W; r++) {
        const uint64_t rr = recv_rows[r], sr = send_rows[r], sf = send_first[r];
        if (r == me) {
            if (rr != sr) throw Error("internal: own share differs between the send and the receive plan");
            if (rr) {
                SKA_CUDA(cudaMemcpyAsync(r_keys.p + ro, local.keys.p + sf, rr * 8, cudaMemcpyDeviceToDevice, c.stream));
                SKA_CUDA(cudaMemcpyAsync(r_cnt.p + ro, local.counts.p + sf, rr * 4, cudaMemcpyDeviceToDevice, c.stream));
                SKA_CUDA(cudaMemcpyAsync(r_var.p + vo, local.variants.p + sf * nl, rr * nl, cudaMemcpyDeviceToDevice, c.stream));
            }
        } else {
            if (sr) {
                SKA_NCCL(nccl().Send(local.keys.p + sf, sr * 8, ncclChar, r, cm.comm, c.stream));
                SKA_NCCL(nccl().Send(local.counts.p + sf, sr * 4, ncclChar, r, cm.comm, c.stream));
                SKA_NCCL(nccl().Send(local.variants.p + sf * nl, sr * nl, ncclChar, r, cm.comm, c.stream));
            }
            if (rr) {
                SKA_NCCL(nccl().Recv(r_keys.p + ro, rr * 8, ncclChar, r, cm.comm, c.stream));
                SKA_NCCL(nccl().Recv(r_cnt.p + ro, rr * 4, ncclChar, r, cm.comm, c.stream));
                SKA_NCCL(nccl().Recv(r_var.p + vo, rr * cols[r], ncclChar, r, cm.comm, c.stream));
            }
        }
        ro += rr; vo += rr * cols[r];
    }
    SKA_NCCL(nccl().GroupEnd());
    arena().reset();
    merge_partial_rows(r_keys.p, r_var.p, r_cnt.p, W, recv_rows.data(), cols.data(), col_base.data(), n_samples_total, local.k, local.rc, slab);
    SKA_CUDA(cudaStreamSynchronize(c.stream));                          // the receive buffers die here
    if (stats) {
        stats[0] = local.n_rows - send_rows[me];                        // rows sent to peers
        stats[1] = n_recv - recv_rows[me];                              // rows received from peers
        stats[2] = (local.n_rows - send_rows[me]) * (12 + nl);          // bytes sent
        stats[3] = slab.n_rows;
    }
}

}  // namespace ska
