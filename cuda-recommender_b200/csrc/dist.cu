// dist.cu — multi-GPU exchange: one process per GPU, NCCL over NVLink 5 / NVSwitch.
//
// The reference is single-GPU (cudaSetDevice(0), cuda_src/CCD_CUDA.cu:170, ALS_CUDA.cu:189); the
// sharding is the north star's: GPU r keeps CSR row block r and CSC column block r (nnz-balanced),
// every GPU keeps full-length factor vectors, and the block of u_t / v_t (CCD++) or W / H (ALS) a GPU
// has just solved is all-gathered in place.  Blocks are unequal (nnz-balanced, not row-balanced), so
// the gather is a grouped set of ncclBroadcast calls, one per owner, which NCCL fuses into one launch.
//
// NCCL is bound at run time (dlopen of libnccl.so.2) so the library loads, and single-GPU sessions
// run, on a box without NCCL, and so that inside a torch process the already-loaded NCCL is shared.
#include <dlfcn.h>
#include <stdlib.h>
#include <string.h>

#include "session.cuh"

namespace mf {
namespace {

typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
enum { ncclSuccess = 0 };
enum { ncclInt8 = 0, ncclFloat32 = 7 };  // ncclDataType_t (nccl.h): int8 0, uint8 1, int32 2, uint32 3, int64 4, uint64 5, half 6, float 7

struct Api {
    void* handle = nullptr;
    int (*GetUniqueId)(ncclUniqueId*) = nullptr;
    int (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    int (*CommDestroy)(ncclComm_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    int (*Broadcast)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
};

Api g_api;

int load_api() {
    if (g_api.handle) return MF_OK;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    void* h = nullptr;
    for (const char* n : names) {
        h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (h) break;
    }
    if (!h) {
        set_error("NCCL not found (dlopen libnccl.so.2): %s", dlerror());
        return MF_ERR_NCCL;
    }
    Api a;
    a.handle = h;
#define MF_SYM(field, name)                                               \
    *(void**)(&a.field) = dlsym(h, name);                                 \
    if (!a.field) { set_error("NCCL symbol %s missing", name); return MF_ERR_NCCL; }
    MF_SYM(GetUniqueId, "ncclGetUniqueId")
    MF_SYM(CommInitRank, "ncclCommInitRank")
    MF_SYM(CommDestroy, "ncclCommDestroy")
    MF_SYM(GroupStart, "ncclGroupStart")
    MF_SYM(GroupEnd, "ncclGroupEnd")
    MF_SYM(Broadcast, "ncclBroadcast")
    MF_SYM(AllGather, "ncclAllGather")
    MF_SYM(GetErrorString, "ncclGetErrorString")
#undef MF_SYM
    g_api = a;
    return MF_OK;
}

#define MF_NCCL(expr)                                                                             \
    do {                                                                                          \
        int _r = (expr);                                                                          \
        if (_r != ncclSuccess) {                                                                  \
            set_error("NCCL error %d at %s:%d: %s", _r, __FILE__, __LINE__, g_api.GetErrorString(_r)); \
            return MF_ERR_NCCL;                                                                   \
        }                                                                                         \
    } while (0)

}  // namespace

// Communicators are kept for the life of the process, keyed by the 128-byte unique id (+ rank, size, device): a host that
// opens one session after another with the same id — a training service, the benchmark's end-to-end leg — pays
// ncclCommInitRank (0.5-1 s) once.  mf_release_cached_memory() destroys them.
// The peer-to-peer state of a communicator is kept with it as well: the exported receive buffers / flag words and the
// CUDA IPC mappings of the peers' (21 cudaIpcOpenMemHandle calls at 8 ranks) cost 45-70 ms per session at 8 GPUs — more
// than the three outer iterations of an end-to-end call.  A later session of the same shape on the same communicator
// adopts them (the exchange epoch carries on where the previous session stopped, so stale receive words never match).
struct P2PState {
    int64_t ldm = 0, ldn = 0;
    bool in_use = false;
    std::vector<void*> opened;
    unsigned** d_peerFlags = nullptr;
    unsigned* flags = nullptr;
    unsigned long long *llW = nullptr, *llH = nullptr;
    unsigned long long **d_peerLLW = nullptr, **d_peerLLH = nullptr;
    unsigned epoch = 0;
};
struct CachedComm {
    char id[128];
    int rank, nranks, device;
    ncclComm_t comm;
    P2PState* p2p;  // nullptr: none cached
};
std::vector<CachedComm> g_comms;

struct Dist {
    ncclComm_t comm = nullptr;
    int rank = 0, nranks = 1;
    // peer-to-peer exchange over NVLink (CUDA IPC mappings of every peer's factor buffers and flag words)
    bool p2p = false;
    std::vector<void*> opened;      // mappings to close
    unsigned** d_peerFlags = nullptr;  // [nranks] device array: flag words of every rank
    unsigned* flags = nullptr;      // [nranks + 1] own flag words (flags[r] written by rank r) + block ticket
    unsigned long long* llW = nullptr;  // own low-latency receive buffers: one (value, epoch) word per factor entry
    unsigned long long* llH = nullptr;
    unsigned long long** d_peerLLW = nullptr;  // [nranks] device arrays of every rank's receive buffers
    unsigned long long** d_peerLLH = nullptr;
    unsigned epoch = 0;             // exchanges issued so far
    P2PState* cached = nullptr;     // the communicator's cached peer-to-peer state this session runs on (not owned)
};

int dist_unique_id(void* id128) {
    MF_REQUIRE(id128 != nullptr, "NULL argument");
    MF_TRY(load_api());
    ncclUniqueId id;
    MF_NCCL(g_api.GetUniqueId(&id));
    memcpy(id128, &id, sizeof(id));
    return MF_OK;
}

int dist_create(Dist** out, int rank, int nranks, const void* id128, int device) {
    MF_TRY(load_api());
    MF_CUDA(cudaSetDevice(device));
    ncclUniqueId id;
    memcpy(&id, id128, sizeof(id));
    Dist* d = new Dist();
    d->rank = rank;
    d->nranks = nranks;
    for (const CachedComm& c : g_comms)
        if (c.rank == rank && c.nranks == nranks && c.device == device && memcmp(c.id, id128, 128) == 0) d->comm = c.comm;
    if (!d->comm) {
        int r = g_api.CommInitRank(&d->comm, nranks, id, rank);
        if (r != ncclSuccess) {
            set_error("ncclCommInitRank failed (%d): %s", r, g_api.GetErrorString(r));
            delete d;
            return MF_ERR_NCCL;
        }
        CachedComm c;
        memcpy(c.id, id128, 128);
        c.rank = rank; c.nranks = nranks; c.device = device; c.comm = d->comm; c.p2p = nullptr;
        g_comms.push_back(c);
    }
    *out = d;
    return MF_OK;
}

static void p2p_free(P2PState* c, ncclComm_t comm, int rank, int nranks) {
    for (void* p : c->opened) cudaIpcCloseMemHandle(p);
    if (comm) {  // every rank has closed its imports before any rank frees its exports (see dist_destroy)
        char* d_b = nullptr;
        if (dev_alloc(&d_b, (size_t)nranks) == MF_OK) {
            if (g_api.AllGather(d_b + rank, d_b, 1, ncclInt8, comm, nullptr) == ncclSuccess) cudaStreamSynchronize(nullptr);
            dev_free(d_b);
        }
    }
    void* ptrs[] = {c->d_peerFlags, c->flags, c->llW, c->llH, c->d_peerLLW, c->d_peerLLH};
    for (void* p : ptrs)
        if (p) dev_free(p);
    delete c;
}

int dist_destroy(Dist* d) {
    if (!d) return MF_OK;
    if (d->cached) {  // the state stays with the communicator for the next session of this shape
        d->cached->epoch = d->epoch;
        d->cached->in_use = false;
        delete d;
        return MF_OK;
    }
    for (void* p : d->opened) cudaIpcCloseMemHandle(p);
    if (d->p2p && d->comm) {
        // the exporter must not free a buffer a peer still has mapped (CUDA IPC: undefined behaviour): every rank has closed
        // its imports above before any rank frees its exports below — a one-byte all-gather is the barrier (p2p is agreed
        // across the ranks, so they all come here)
        char* d_b = nullptr;
        if (dev_alloc(&d_b, (size_t)d->nranks) == MF_OK) {
            if (g_api.AllGather(d_b + d->rank, d_b, 1, ncclInt8, d->comm, nullptr) == ncclSuccess) cudaStreamSynchronize(nullptr);
            dev_free(d_b);
        }
    }
    if (d->d_peerFlags) dev_free(d->d_peerFlags);
    if (d->flags) dev_free(d->flags);
    if (d->llW) dev_free(d->llW);
    if (d->llH) dev_free(d->llH);
    if (d->d_peerLLW) dev_free(d->d_peerLLW);
    if (d->d_peerLLH) dev_free(d->d_peerLLH);
    delete d;  // the communicator stays in g_comms
    return MF_OK;
}

void dist_release_cached(int device) {
    for (size_t i = 0; i < g_comms.size();) {
        if (g_comms[i].device == device) {
            if (g_comms[i].p2p && !g_comms[i].p2p->in_use) { p2p_free(g_comms[i].p2p, g_comms[i].comm, g_comms[i].rank, g_comms[i].nranks); g_comms[i].p2p = nullptr; }
            if (g_api.CommDestroy) g_api.CommDestroy(g_comms[i].comm);
            g_comms.erase(g_comms.begin() + i);
        } else {
            ++i;
        }
    }
}

// Maps every peer's low-latency receive buffers and flag words into this process (CUDA IPC over NVLink / NVSwitch) so that
// the finalize can store a freshly solved block straight into the peers (ccd_kernels.cu).  The 64-byte IPC handles travel
// through one ncclAllGather, and whether the peer-to-peer path is used at all is AGREED across the ranks: every rank takes
// part in both collectives below whatever happened to it locally (MF_NO_P2P, no IPC support, a mapping that failed), and
// the path is switched on only when every rank succeeded — a rank that silently stayed on NCCL while its peers pushed
// LL words would hang the job.
int dist_setup_p2p(Dist* d, float* W, float* H, int64_t ldm, int64_t ldn, cudaStream_t st) {
    (void)W; (void)H;
    if (!d || d->nranks <= 1) return MF_OK;
    const int P = d->nranks;
    CachedComm* slot = nullptr;
    for (CachedComm& c : g_comms)
        if (c.comm == d->comm) slot = &c;
    {   // collective 0: can EVERY rank adopt the state cached with the communicator?  (same shape, not held by a live session)
        const bool have = slot && slot->p2p && !slot->p2p->in_use && slot->p2p->ldm == ldm && slot->p2p->ldn == ldn && getenv("MF_NO_P2P") == nullptr &&
                          getenv("MF_NO_P2P_CACHE") == nullptr;
        int* d_have = nullptr;
        MF_TRY(dev_alloc(&d_have, (size_t)P));
        const bool stale = !have && slot && slot->p2p && !slot->p2p->in_use;  // cached for another shape (or switched off)
        const int mine_have = have ? 1 : (stale ? 2 : 0);
        MF_CUDA(cudaMemcpyAsync(d_have + d->rank, &mine_have, sizeof(int), cudaMemcpyHostToDevice, st));
        MF_NCCL(g_api.AllGather(d_have + d->rank, d_have, sizeof(int), ncclInt8, d->comm, st));
        std::vector<int> all_have((size_t)P);
        MF_CUDA(cudaMemcpyAsync(all_have.data(), d_have, sizeof(int) * (size_t)P, cudaMemcpyDeviceToHost, st));
        MF_CUDA(cudaStreamSynchronize(st));
        dev_free(d_have);
        bool everyone_has = true, everyone_stale = true;
        for (int r = 0; r < P; ++r) {
            everyone_has = everyone_has && all_have[r] == 1;
            everyone_stale = everyone_stale && all_have[r] == 2;
        }
        if (everyone_has) {
            P2PState* c = slot->p2p;
            c->in_use = true;
            d->cached = c;
            d->flags = c->flags; d->llW = c->llW; d->llH = c->llH;
            d->d_peerFlags = c->d_peerFlags; d->d_peerLLW = c->d_peerLLW; d->d_peerLLH = c->d_peerLLH;
            d->epoch = c->epoch;
            d->p2p = true;
            return MF_OK;
        }
        if (have || stale) {
            // the cached state is of no use any more: every rank that has one drops it here.  When ALL ranks do (the usual
            // case: the next session has another shape) the exports are freed behind the communicator barrier, like in
            // dist_destroy; in a mixed situation there is no collective every rank would enter.
            p2p_free(slot->p2p, everyone_stale ? d->comm : nullptr, d->rank, P);
            slot->p2p = nullptr;
        }
    }
    struct Handles { cudaIpcMemHandle_t f, lw, lh; int ok; int pad[15]; };
    static_assert(sizeof(Handles) == 256, "three 64-byte handles + a flag");
    Handles mine;
    memset(&mine, 0, sizeof(mine));
    bool local_ok = getenv("MF_NO_P2P") == nullptr;
    if (local_ok) {
        local_ok = dev_alloc(&d->flags, (size_t)P + 1) == MF_OK && dev_alloc(&d->llW, (size_t)ldm) == MF_OK && dev_alloc(&d->llH, (size_t)ldn) == MF_OK &&
                   cudaMemsetAsync(d->flags, 0, sizeof(unsigned) * ((size_t)P + 1), st) == cudaSuccess &&
                   cudaMemsetAsync(d->llW, 0, sizeof(unsigned long long) * (size_t)ldm, st) == cudaSuccess &&
                   cudaMemsetAsync(d->llH, 0, sizeof(unsigned long long) * (size_t)ldn, st) == cudaSuccess &&
                   cudaIpcGetMemHandle(&mine.f, d->flags) == cudaSuccess && cudaIpcGetMemHandle(&mine.lw, d->llW) == cudaSuccess &&
                   cudaIpcGetMemHandle(&mine.lh, d->llH) == cudaSuccess;
        if (!local_ok) cudaGetLastError();
    }
    mine.ok = local_ok ? 1 : 0;
    // collective 1: everybody's handles and "I can export"
    Handles* d_all = nullptr;
    MF_TRY(dev_alloc(&d_all, (size_t)P));
    MF_CUDA(cudaMemcpyAsync(d_all + d->rank, &mine, sizeof(Handles), cudaMemcpyHostToDevice, st));
    MF_NCCL(g_api.AllGather(d_all + d->rank, d_all, sizeof(Handles), ncclInt8, d->comm, st));
    std::vector<Handles> all((size_t)P);
    MF_CUDA(cudaMemcpyAsync(all.data(), d_all, sizeof(Handles) * (size_t)P, cudaMemcpyDeviceToHost, st));
    MF_CUDA(cudaStreamSynchronize(st));
    bool everyone = true;
    for (int r = 0; r < P; ++r) everyone = everyone && all[r].ok != 0;
    std::vector<unsigned*> pf((size_t)P);
    std::vector<unsigned long long*> plw((size_t)P), plh((size_t)P);
    bool ok = everyone;
    for (int r = 0; r < P && ok; ++r) {
        if (r == d->rank) { pf[r] = d->flags; plw[r] = d->llW; plh[r] = d->llH; continue; }
        const cudaIpcMemHandle_t* hs[3] = {&all[r].f, &all[r].lw, &all[r].lh};
        void* m[3] = {nullptr, nullptr, nullptr};
        for (int q = 0; q < 3 && ok; ++q) {
            ok = cudaIpcOpenMemHandle(&m[q], *hs[q], cudaIpcMemLazyEnablePeerAccess) == cudaSuccess;
            if (ok) d->opened.push_back(m[q]);
        }
        pf[r] = (unsigned*)m[0]; plw[r] = (unsigned long long*)m[1]; plh[r] = (unsigned long long*)m[2];
    }
    if (!ok) cudaGetLastError();
    // collective 2: "I could map everybody" — the path is on only if that holds on every rank
    int* d_flag = reinterpret_cast<int*>(d_all);  // reuse: P ints fit in the handle table
    const int my_flag = ok ? 1 : 0;
    MF_CUDA(cudaMemcpyAsync(d_flag + d->rank, &my_flag, sizeof(int), cudaMemcpyHostToDevice, st));
    MF_NCCL(g_api.AllGather(d_flag + d->rank, d_flag, sizeof(int), ncclInt8, d->comm, st));
    std::vector<int> flags((size_t)P);
    MF_CUDA(cudaMemcpyAsync(flags.data(), d_flag, sizeof(int) * (size_t)P, cudaMemcpyDeviceToHost, st));
    MF_CUDA(cudaStreamSynchronize(st));
    dev_free(d_all);
    for (int r = 0; r < P; ++r) ok = ok && flags[r] != 0;
    if (!ok) {  // somebody cannot: everybody stays on the NCCL broadcast path
        for (void* p : d->opened) cudaIpcCloseMemHandle(p);
        d->opened.clear();
        return MF_OK;
    }
    MF_TRY(dev_alloc(&d->d_peerFlags, (size_t)P));
    MF_CUDA(cudaMemcpy(d->d_peerFlags, pf.data(), sizeof(unsigned*) * (size_t)P, cudaMemcpyHostToDevice));
    MF_TRY(dev_alloc(&d->d_peerLLW, (size_t)P));
    MF_TRY(dev_alloc(&d->d_peerLLH, (size_t)P));
    MF_CUDA(cudaMemcpy(d->d_peerLLW, plw.data(), sizeof(void*) * (size_t)P, cudaMemcpyHostToDevice));
    MF_CUDA(cudaMemcpy(d->d_peerLLH, plh.data(), sizeof(void*) * (size_t)P, cudaMemcpyHostToDevice));
    d->p2p = true;
    if (slot && !slot->p2p && getenv("MF_NO_P2P_CACHE") == nullptr) {  // keep it with the communicator (ownership moves to the cache)
        P2PState* c = new P2PState();
        c->ldm = ldm; c->ldn = ldn; c->in_use = true;
        c->opened.swap(d->opened);
        c->flags = d->flags; c->llW = d->llW; c->llH = d->llH;
        c->d_peerFlags = d->d_peerFlags; c->d_peerLLW = d->d_peerLLW; c->d_peerLLH = d->d_peerLLH;
        slot->p2p = c;
        d->cached = c;
    }
    return MF_OK;
}

bool dist_p2p(const Dist* d) { return d && d->p2p; }
int dist_rank(const Dist* d) { return d ? d->rank : 0; }
unsigned* const* dist_peer_flags(const Dist* d) { return d->d_peerFlags; }
unsigned* dist_flags(const Dist* d) { return d->flags; }
unsigned dist_next_epoch(Dist* d) { return ++d->epoch; }
unsigned dist_advance_epoch(Dist* d, unsigned n) { const unsigned e = d->epoch; d->epoch += n; return e; }
unsigned long long* const* dist_peer_ll(const Dist* d, bool h) { return h ? d->d_peerLLH : d->d_peerLLW; }
unsigned long long* dist_ll(const Dist* d, bool h) { return h ? d->llH : d->llW; }

int dist_allgather_blocks(Dist* d, float* vec, const std::vector<int64_t>& bound, int64_t unit, cudaStream_t st) {
    if (!d || d->nranks <= 1) return MF_OK;
    MF_NCCL(g_api.GroupStart());
    for (int r = 0; r < d->nranks; ++r) {
        const int64_t lo = bound[r] * unit, n = (bound[r + 1] - bound[r]) * unit;
        if (n <= 0) continue;
        MF_NCCL(g_api.Broadcast(vec + lo, vec + lo, (size_t)n, ncclFloat32, r, d->comm, st));
    }
    MF_NCCL(g_api.GroupEnd());
    return MF_OK;
}

}  // namespace mf
