/*
 * pgpu_exchange.cu -- EXCHANGE over NCCL: ship every particle whose cell belongs to another rank.
 *
 * Replaces: EXCHANGE (src/grid_and_partition.f90:525-622) = reverse scan + pack of non-local particles
 * (:562-583), COUNTING_SORT by destination (:482-518), MPI_ALLTOALL of counts (:590), MPI_ALLTOALLV of
 * MPI_PARTICLE_DATA_STRUCTURE (:603-605), append of the received particles (:610-612); and the
 * MPI_REDUCE + MPI_BCAST of RHS in DEPOSIT_CHARGE (src/fields.f90:5529-5535) as one ncclAllReduce.
 *
 * One process per GPU; the NCCL unique id is created by rank 0 and broadcast by the host program
 * (MPI_BCAST in PANTERA, torch.distributed in this repo's harness).  NCCL is dlopen'ed on first use so
 * that single-GPU runs have no NCCL dependency and so that the copy already loaded by the host process
 * (e.g. the one bundled with PyTorch) is the one used.
 *
 * Wire format: 88-byte records {X,Y,Z,VX,VY,VZ,EROT,EVIB,DTRIM,(IC,S_ID),ID}, grouped by destination.
 */
#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>

#include "pgpu_internal.cuh"

namespace {

struct NcclApi {
    ncclResult_t (*GetUniqueId)(ncclUniqueId*);
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int);
    ncclResult_t (*CommDestroy)(ncclComm_t);
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t);
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t);
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
    ncclResult_t (*GroupStart)();
    ncclResult_t (*GroupEnd)();
    const char* (*GetErrorString)(ncclResult_t);
    void* lib;
};
NcclApi g_nccl = {};

const char* load_nccl() {
    if (g_nccl.lib) return nullptr;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    void* h = nullptr;
    for (const char* n : names) {
        h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (h) break;
    }
    if (!h) return "libnccl.so.2 not found";
#define SYM(field, name)                                   \
    *(void**)(&g_nccl.field) = dlsym(h, name);             \
    if (!g_nccl.field) return "missing NCCL symbol " name;
    SYM(GetUniqueId, "ncclGetUniqueId")
    SYM(CommInitRank, "ncclCommInitRank")
    SYM(CommDestroy, "ncclCommDestroy")
    SYM(AllGather, "ncclAllGather")
    SYM(AllReduce, "ncclAllReduce")
    SYM(Send, "ncclSend")
    SYM(Recv, "ncclRecv")
    SYM(GroupStart, "ncclGroupStart")
    SYM(GroupEnd, "ncclGroupEnd")
    SYM(GetErrorString, "ncclGetErrorString")
#undef SYM
    g_nccl.lib = h;
    return nullptr;
}

#define PG_NCCL(ctx, call)                                                                                  \
    do {                                                                                                    \
        ncclResult_t r_ = (call);                                                                           \
        if (r_ != ncclSuccess) {                                                                            \
            snprintf((ctx)->err, sizeof((ctx)->err), "NCCL error %s at %s:%d", g_nccl.GetErrorString(r_),   \
                     __FILE__, __LINE__);                                                                   \
            return PGPU_ERR_NCCL;                                                                           \
        }                                                                                                   \
    } while (0)

#define XREC_DOUBLES 11 /* 88 bytes */

/* pass 1: emigrants per destination rank (COUNTING_SORT's histogram, ref: grid_and_partition.f90:497-500) */
__global__ void __launch_bounds__(256) xcount_kernel(const __grid_constant__ PgDev d, unsigned long long* __restrict__ xcount) {
    __shared__ unsigned int s[64];
    if (threadIdx.x < 64) s[threadIdx.x] = 0;
    __syncthreads();
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < (long long)d.status->np_cur) {
        int ic = d.p.pos[i].ic;
        if (ic >= 1) {
            int dest = __ldg(&d.cell_procs[ic - 1]);
            if (dest != d.rank) atomicAdd(&s[dest], 1u);
        }
    }
    __syncthreads();
    if (threadIdx.x < d.nranks && s[threadIdx.x]) atomicAdd(&xcount[threadIdx.x], (unsigned long long)s[threadIdx.x]);
}

/* pass 2: pack emigrants behind their destination's offset and tombstone them locally */
__global__ void __launch_bounds__(256) xpack_kernel(const __grid_constant__ PgDev d, double* __restrict__ sendbuf,
                                                    const unsigned long long* __restrict__ xoff,
                                                    unsigned long long* __restrict__ xcursor) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)d.status->np_cur) return;
    const PgPos P = d.p.pos[i];
    int ic = P.ic;
    if (ic < 1) return;
    int dest = __ldg(&d.cell_procs[ic - 1]);
    if (dest == d.rank) return;
    unsigned long long k = xoff[dest] + atomicAdd(&xcursor[dest], 1ULL);
    double* r = sendbuf + k * XREC_DOUBLES;
    const PgVel V = d.p.vel[i];
    const PgIe E = pg_ie_load(d.p.ie, i);
    r[0] = P.x; r[1] = P.y; r[2] = P.z;
    r[3] = V.vx; r[4] = V.vy; r[5] = V.vz;
    r[6] = E.erot; r[7] = E.evib;
    r[8] = d.p.dtrim ? d.p.dtrim[i] : d.dt;
    int2 cs = make_int2(ic, P.sid);
    r[9] = *reinterpret_cast<double*>(&cs);
    long long id = d.p.id ? d.p.id[i] : 0;
    r[10] = *reinterpret_cast<double*>(&id);
    d.p.pos[i].ic = PG_TOMBSTONE;
}

/* pass 2, fast path: the emigrants are the entries of the list the ADVECT epilogue wrote (pgpu_advect.cu
 * store_particle), so packing touches the emigrants only instead of sweeping the whole store */
__global__ void __launch_bounds__(256) xpack_list_kernel(const __grid_constant__ PgDev d, const unsigned int* __restrict__ list,
                                                         long long n, double* __restrict__ sendbuf,
                                                         const unsigned long long* __restrict__ xoff,
                                                         unsigned long long* __restrict__ xcursor) {
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const long long i = (long long)list[t];
    const PgPos P = d.p.pos[i];
    const int ic = P.ic;
    const int dest = __ldg(&d.cell_procs[ic - 1]);
    unsigned long long k = xoff[dest] + atomicAdd(&xcursor[dest], 1ULL);
    double* r = sendbuf + k * XREC_DOUBLES;
    const PgVel V = d.p.vel[i];
    const PgIe E = pg_ie_load(d.p.ie, i);
    r[0] = P.x; r[1] = P.y; r[2] = P.z;
    r[3] = V.vx; r[4] = V.vy; r[5] = V.vz;
    r[6] = E.erot; r[7] = E.evib;
    r[8] = d.p.dtrim ? d.p.dtrim[i] : d.dt;
    int2 cs = make_int2(ic, P.sid);
    r[9] = *reinterpret_cast<double*>(&cs);
    long long id = d.p.id ? d.p.id[i] : 0;
    r[10] = *reinterpret_cast<double*>(&id);
    d.p.pos[i].ic = PG_TOMBSTONE;
}

/* append the received records (ADD_PARTICLE_ARRAY loop, ref: grid_and_partition.f90:610-612).  with_rank: the
 * histogram of the ADVECT epilogue is still valid; the immigrant joins it, so
 * the counting sort that follows needs no recount. */
__global__ void __launch_bounds__(256) xunpack_kernel(const __grid_constant__ PgDev d, const double* __restrict__ recvbuf,
                                                      long long n, PgPlanes p, long long base, int with_rank) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double* r = recvbuf + i * XREC_DOUBLES;
    long long o = base + i;
    double t = r[9];
    int2 cs = *reinterpret_cast<int2*>(&t);
    PgPos P;
    P.x = r[0]; P.y = r[1]; P.z = r[2]; P.ic = cs.x; P.sid = cs.y;
    PgVel V;
    V.vx = r[3]; V.vy = r[4]; V.vz = r[5]; V.sid = cs.y; V.pad = 0;
    PgIe E;
    E.erot = r[6]; E.evib = r[7];
    p.pos[o] = P;
    p.vel[o] = V;
    pg_ie_store(p.ie, o, E.erot, E.evib, d.status);
    if (E.erot != 0. || E.evib != 0.) d.status->flag_ie = 1;
    double dtrim = r[8];
    if (p.dtrim) p.dtrim[o] = dtrim;
    else if (dtrim != d.dt) d.status->need_dtrim = 1;
    if (dtrim != d.dt) d.status->flag_dtrim = 1;
    t = r[10];
    if (p.id) p.id[o] = *reinterpret_cast<long long*>(&t);
    if (with_rank) atomicAdd(&d.cell_count[cs.x - 1], 1u);
}

/* 88-byte wire records -> 96-byte host records (TYPE PARTICLE_DATA_STRUCTURE, src/particle.f90:28-37) */
__global__ void __launch_bounds__(256) xrec_to_aos_kernel(const double* __restrict__ rec, long long n, unsigned char* __restrict__ aos,
                                                          int rank, long long id_base, int keep_id) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double* r = rec + i * XREC_DOUBLES;
    double* f = reinterpret_cast<double*>(aos + (size_t)i * 96);
    for (int k = 0; k < 9; k++) f[k] = r[k];
    f[9] = r[9]; /* (IC, S_ID) */
    double t = r[10];
    long long id = *reinterpret_cast<long long*>(&t);
    if (!keep_id) id = (long long)rank + ((id_base + i) << 8);
    *reinterpret_cast<long long*>(aos + (size_t)i * 96 + 80) = id;
    *reinterpret_cast<long long*>(aos + (size_t)i * 96 + 88) = 0;
}

/* ------------------------------------------------------------------------------------------
 * EXCHANGE without host round trips (the per-step path).  Every pair of ranks exchanges ONE fixed-capacity message per
 * step: record 0 is a header holding the number of particle records that follow, the capacity comes from the counts of
 * the previous step (known to both ends: what one sent is what the other received), so neither the counts nor the
 * sizes have to visit the host before ncclSend/ncclRecv can be queued.  ("fixed-capacity per-neighbour
 * send slots with the count in the header".)
 * ---------------------------------------------------------------------------------------- */
struct XPlan {
    double* msg[64];           /* first record (the header) of the message for / from rank r.  Sending: the slot of
                                  this rank in r's window (peer memory over NVLink) or a segment of the local send
                                  buffer (NCCL pair); receiving: the slot of r in this rank's window, or a segment of
                                  the receive buffer */
    unsigned long long* flag[64]; /* sending: r's arrival flag for this rank (window pairs), else null;
                                     receiving: this rank's arrival flag for r, else null */
    unsigned int cap[64];      /* particle records the message holds */
    unsigned char on[64];      /* a message is exchanged with rank r this step (never with oneself) */
    unsigned long long seq;    /* number of this exchange (arrival flags carry it) */
    int nr, rank;
};

/* Emigrant list -> messages, written where they are read: for a window pair the message slot is PEER MEMORY, so the
 * pack IS the send (no staging copy, no transfer kernel, and only the records that exist cross NVLink, not the slot's
 * capacity).  A warp takes 32 entries of the list, reserves their message slots with one atomic per destination, lays
 * the 88-byte records out in shared memory and writes each destination's run as ONE contiguous, lane-coalesced range
 * (8-byte stores of scattered records would cross the link as quarter-filled packets). */
__global__ void __launch_bounds__(256) xpack_fast_kernel(const __grid_constant__ PgDev d, const unsigned int* __restrict__ list,
                                                         const unsigned long long* __restrict__ xrow,
                                                         const __grid_constant__ XPlan plan, unsigned long long* __restrict__ xcursor) {
    __shared__ double stage[8][32 * XREC_DOUBLES];
    unsigned long long n = xrow[PG_XROW_LEN];
    if (n > (unsigned long long)d.emig_cap) n = (unsigned long long)d.emig_cap;
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    const int lane = threadIdx.x & 31;
    double* const st = stage[threadIdx.x >> 5];
    for (unsigned long long base = (unsigned long long)blockIdx.x * blockDim.x; base < n; base += stride) {
        const unsigned long long t = base + threadIdx.x;
        const bool on = t < n;
        double rec[XREC_DOUBLES];
        int dest = -1;
        if (on) {
            const long long i = (long long)list[t];
            const PgPos P = d.p.pos[i];
            dest = __ldg(&d.cell_procs[P.ic - 1]);
            const PgVel V = d.p.vel[i];
            const PgIe E = pg_ie_load(d.p.ie, i);
            rec[0] = P.x; rec[1] = P.y; rec[2] = P.z;
            rec[3] = V.vx; rec[4] = V.vy; rec[5] = V.vz;
            rec[6] = E.erot; rec[7] = E.evib;
            rec[8] = d.p.dtrim ? d.p.dtrim[i] : d.dt;
            int2 cs = make_int2(P.ic, P.sid);
            rec[9] = *reinterpret_cast<double*>(&cs);
            long long id = d.p.id ? d.p.id[i] : 0;
            rec[10] = *reinterpret_cast<double*>(&id);
            d.p.pos[i].ic = PG_TOMBSTONE;
        }
        unsigned int rem = __ballot_sync(0xffffffffu, on);
        while (rem) { /* one round per destination among the warp's entries (a neighbour or two) */
            const int lead = __ffs(rem) - 1;
            const int dg = __shfl_sync(0xffffffffu, dest, lead);
            const unsigned int grp = __ballot_sync(0xffffffffu, on && dest == dg);
            const int cnt = __popc(grp);
            unsigned long long k0 = 0;
            if (lane == lead) k0 = atomicAdd(&xcursor[dg], (unsigned long long)cnt);
            k0 = __shfl_sync(0xffffffffu, k0, lead);
            const unsigned long long cap = plan.on[dg] ? (unsigned long long)plan.cap[dg] : 0ULL;
            const int fit = (k0 >= cap) ? 0 : (int)((cap - k0 < (unsigned long long)cnt) ? cap - k0 : (unsigned long long)cnt);
            if (grp & (1u << lane)) {
                const int rk = __popc(grp & ((1u << lane) - 1u));
#pragma unroll
                for (int j = 0; j < XREC_DOUBLES; j++) st[rk * XREC_DOUBLES + j] = rec[j];
            }
            __syncwarp();
            if (fit < cnt && lane == lead) { /* does not fit: the step fails loudly at its read-back */
                atomicAdd(&d.status->x_overflow, (unsigned long long)(cnt - fit));
                pg_set_err(d, PG_DEVERR_XCAP);
            }
            if (fit > 0) {
                double* out = plan.msg[dg] + (size_t)(1 + k0) * XREC_DOUBLES;
                for (int j = lane; j < fit * XREC_DOUBLES; j += 32) out[j] = st[j];
            }
            __syncwarp();
            rem &= ~grp;
        }
    }
}

/* headers of the outgoing messages; also the per-destination counts for the host's next plan.  Window pairs: the
 * records were written into the peer's memory by the pack kernel, which has completed (stream order); the header
 * follows, then -- fenced behind it -- the arrival flag the peer's xwait_kernel is polling. */
__global__ void xheader_kernel(const __grid_constant__ PgDev d, const unsigned long long* __restrict__ xrow, const __grid_constant__ XPlan plan) {
    const int r = threadIdx.x;
    if (r == 0 && xrow[PG_XROW_OVF]) { /* the emigrant list itself overflowed: particles would be lost */
        atomicAdd(&d.status->x_overflow, xrow[PG_XROW_OVF]);
        pg_set_err(d, PG_DEVERR_XCAP);
    }
    if (r >= plan.nr) return;
    unsigned long long n = (r == plan.rank) ? 0ULL : xrow[r];
    d.status->x_sent[r] = n;
    if (!plan.on[r]) return; /* (an emigrant for a rank without a message has already failed the step in the pack) */
    if (n > plan.cap[r]) n = plan.cap[r];
    unsigned long long* h = reinterpret_cast<unsigned long long*>(plan.msg[r]);
    h[0] = n;
    if (plan.flag[r]) {
        __threadfence_system();
        *reinterpret_cast<volatile unsigned long long*>(plan.flag[r]) = plan.seq;
    }
}

/* window pairs: wait until every peer's message of this exchange has arrived (its flag carries the exchange number).
 * The peers' packs do not depend on this rank, so the wait is bounded by the skew between ranks; a peer that stopped
 * is reported after 30 s instead of hanging the device. */
__global__ void xwait_kernel(const __grid_constant__ PgDev d, const __grid_constant__ XPlan plan) {
    const int r = threadIdx.x;
    if (r >= plan.nr || !plan.on[r] || !plan.flag[r]) return;
    const volatile unsigned long long* f = reinterpret_cast<const volatile unsigned long long*>(plan.flag[r]);
    unsigned long long t0, t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    while (*f < plan.seq) {
        __nanosleep(200);
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
        if (t1 - t0 > 30000000000ULL) {
            pg_set_err(d, PG_DEVERR_XWAIT);
            break;
        }
    }
    __threadfence_system();
}

/* incoming messages -> append region; grid (capacity blocks, source rank).  The immigrants join the histogram of the
 * ADVECT epilogue, so the counting sort that follows needs no recount. */
__global__ void __launch_bounds__(256) xunpack_fast_kernel(const __grid_constant__ PgDev d, const __grid_constant__ XPlan plan, PgPlanes p) {
    const int src = blockIdx.y;
    if (!plan.on[src]) return;
    const unsigned long long* hdr = reinterpret_cast<const unsigned long long*>(plan.msg[src]);
    unsigned long long n = hdr[0];
    if (n > plan.cap[src]) n = plan.cap[src];
    unsigned long long base = d.status->np_cur;
    for (int q = 0; q < src; q++) {
        if (!plan.on[q]) continue;
        unsigned long long m = *reinterpret_cast<const unsigned long long*>(plan.msg[q]);
        base += (m > plan.cap[q]) ? plan.cap[q] : m;
    }
    for (unsigned long long k = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (unsigned long long)gridDim.x * blockDim.x) {
        const double* r = plan.msg[src] + (size_t)(1 + k) * XREC_DOUBLES;
        const long long o = (long long)(base + k);
        if (o >= d.capacity) { pg_set_err(d, PG_DEVERR_CAPACITY); continue; }
        double t = r[9];
        int2 cs = *reinterpret_cast<int2*>(&t);
        PgPos P;
        P.x = r[0]; P.y = r[1]; P.z = r[2]; P.ic = cs.x; P.sid = cs.y;
        PgVel V;
        V.vx = r[3]; V.vy = r[4]; V.vz = r[5]; V.sid = cs.y; V.pad = 0;
        PgIe E;
        E.erot = r[6]; E.evib = r[7];
        p.pos[o] = P;
        p.vel[o] = V;
        pg_ie_store(p.ie, o, E.erot, E.evib, d.status);
        if (E.erot != 0. || E.evib != 0.) d.status->flag_ie = 1;
        if (p.dtrim) p.dtrim[o] = r[8];
        t = r[10];
        if (p.id) p.id[o] = *reinterpret_cast<long long*>(&t);
        atomicAdd(&d.cell_count[cs.x - 1], 1u);
    }
}

/* after the unpack: the particle count moves on the device; per-source counts for the host's next plan */
__global__ void xfinish_kernel(const __grid_constant__ PgDev d, const __grid_constant__ XPlan plan) {
    if (threadIdx.x != 0) return;
    unsigned long long tot = 0;
    for (int q = 0; q < plan.nr; q++) {
        unsigned long long m = 0;
        if (plan.on[q]) {
            m = *reinterpret_cast<const unsigned long long*>(plan.msg[q]);
            d.status->x_recv[q] = m; /* (the uncapped count: the next plan must grow with it) */
            if (m > plan.cap[q]) m = plan.cap[q];
        } else {
            d.status->x_recv[q] = 0;
        }
        tot += m;
    }
    d.status->n_immigrants = tot;
    unsigned long long np = d.status->np_cur + tot;
    if (np > (unsigned long long)d.capacity) { pg_set_err(d, PG_DEVERR_CAPACITY); np = (unsigned long long)d.capacity; }
    d.status->np_cur = np;
}

__global__ void fill_kernel(double* p, long long n, double v) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

} /* namespace */

/*
 * EXCHANGE over peer memory.  Every rank owns a WINDOW: 64 arrival flags, then two parities of one message slot per
 * source rank.  The windows are cudaMalloc'ed once, published as CUDA IPC handles (gathered with ncclAllGather) and mapped
 * into every other rank of the node, so that the pack kernel of a sender writes its records straight into the
 * receiver's memory over NVLink and raises the receiver's flag; NCCL then carries only the pairs that could not be
 * mapped (another node, IPC refused, PANTERA_GPU_XCHG_P2P=0) or whose message outgrew the slot
 * (PANTERA_GPU_XCHG_WINDOW records per slot; default 2^21, at most 4 GiB per rank).
 * Two parities are enough: a sender re-uses a slot two exchanges later, and it cannot get there before the receiver
 * has unpacked it -- its own next exchange waits for the receiver's next message, which the receiver's stream sends
 * after that unpack.
 * Collective (every rank of the communicator calls it from pgpu_comm_init).  A pair is used only if BOTH ends mapped
 * each other (the outcomes are gathered too), so the two ends of a pair always agree on the transport.
 */
static int setup_windows(pgpu_ctx* c) {
    const int nr = c->comm_nranks, me = c->comm_rank;
    ncclComm_t comm = (ncclComm_t)c->nccl_comm;
    const char* e = getenv("PANTERA_GPU_XCHG_P2P");
    const bool want = !(e && *e == '0') && nr > 1;
    long long recs = 1LL << 21;
    if (const char* w = getenv("PANTERA_GPU_XCHG_WINDOW")) recs = std::max(1024LL, atoll(w));
    recs = std::min(recs, (4LL << 30) / (2LL * nr * 8 * XREC_DOUBLES) - 1);
    struct Blob { cudaIpcMemHandle_t h; long long recs; int ok; int pad; };
    static_assert(sizeof(Blob) <= 128, "blob");
    std::vector<unsigned char> mine(128, 0), all((size_t)nr * 128, 0);
    Blob* b = reinterpret_cast<Blob*>(mine.data());
    b->recs = recs;
    if (want) {
        const size_t bytes = 1024 + 2 * (size_t)nr * (size_t)(recs + 1) * XREC_DOUBLES * sizeof(double);
        if (cudaMalloc(&c->x_win, bytes) == cudaSuccess && cudaMemset(c->x_win, 0, 1024) == cudaSuccess &&
            cudaIpcGetMemHandle(&b->h, c->x_win) == cudaSuccess) {
            b->ok = 1;
        } else {
            cudaGetLastError();
            if (c->x_win) cudaFree(c->x_win);
            c->x_win = nullptr;
        }
    }
    unsigned char* d_io = nullptr;
    PG_CUDA(c, cudaMalloc((void**)&d_io, (size_t)(nr + 1) * 128));
    auto gather = [&]() -> int {
        PG_CUDA(c, cudaMemcpyAsync(d_io, mine.data(), 128, cudaMemcpyHostToDevice, c->stream));
        PG_NCCL(c, g_nccl.AllGather(d_io, d_io + 128, 128, ncclChar, comm, c->stream));
        PG_CUDA(c, cudaMemcpyAsync(all.data(), d_io + 128, (size_t)nr * 128, cudaMemcpyDeviceToHost, c->stream));
        PG_CUDA(c, cudaStreamSynchronize(c->stream));
        return PGPU_OK;
    };
    int rc = gather();
    if (rc) { cudaFree(d_io); return rc; }
    void* opened[64] = {};
    for (int r = 0; r < nr && b->ok; r++) {
        if (r == me) continue;
        const Blob* pb = reinterpret_cast<const Blob*>(all.data() + (size_t)r * 128);
        if (!pb->ok || pb->recs != recs) continue;
        void* q = nullptr;
        if (cudaIpcOpenMemHandle(&q, pb->h, cudaIpcMemLazyEnablePeerAccess) == cudaSuccess) opened[r] = q;
        else cudaGetLastError(); /* (another node, the same process, IPC not permitted: that pair stays on NCCL) */
    }
    /* second round: who mapped whom */
    memset(mine.data(), 0, 128);
    for (int r = 0; r < nr; r++) mine[(size_t)r] = opened[r] ? 1 : 0;
    rc = gather();
    cudaFree(d_io);
    if (rc) return rc;
    int npairs = 0;
    for (int r = 0; r < nr; r++) {
        if (r == me || !opened[r]) continue;
        if (all[(size_t)r * 128 + (size_t)me]) { c->x_peer_win[r] = opened[r]; npairs++; }
        else cudaIpcCloseMemHandle(opened[r]);
    }
    c->x_win_recs = recs;
    c->x_seq = 0;
    if (getenv("PANTERA_GPU_XCHG_TRACE"))
        fprintf(stderr, "[pantera_gpu rank %d] EXCHANGE windows: %d of %d peers mapped over CUDA IPC, %lld records per slot\n", me, npairs, nr - 1, recs);
    return PGPU_OK;
}

extern "C" int pgpu_comm_unique_id(char id128[128]) {
    const char* e = load_nccl();
    if (e) return PGPU_ERR_NCCL;
    ncclUniqueId id;
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    if (g_nccl.GetUniqueId(&id) != ncclSuccess) return PGPU_ERR_NCCL;
    memcpy(id128, &id, 128);
    return PGPU_OK;
}

extern "C" void pg_comm_destroy(pgpu_ctx* c);

extern "C" int pgpu_comm_init(pgpu_ctx* c, int nranks, int rank, const char id128[128]) {
    if (!c) return PGPU_ERR_INVALID;
    cudaSetDevice(c->device);
    if (nranks < 1 || nranks > 64 || rank < 0 || rank >= nranks) PG_FAIL(c, PGPU_ERR_INVALID, "comm_init: bad rank %d of %d (max 64 ranks)", rank, nranks);
    const char* e = load_nccl();
    if (e) PG_FAIL(c, PGPU_ERR_NCCL, "comm_init: %s", e);
    if (c->nccl_comm) { /* a second initialisation replaces the first (communicator, windows, counters) */
        pg_comm_destroy(c);
        if (c->d_xcount) { cudaFree(c->d_xcount); c->d_xcount = nullptr; }
        c->x_counts_valid = false;
    }
    ncclUniqueId id;
    memcpy(&id, id128, 128);
    ncclComm_t comm;
    PG_NCCL(c, g_nccl.CommInitRank(&comm, nranks, id, rank));
    c->nccl_comm = (void*)comm;
    c->comm_nranks = nranks;
    c->comm_rank = rank;
    PG_CUDA(c, cudaMalloc((void**)&c->d_xcount, sizeof(unsigned long long) * PG_XROW * (4 + 64)));
    c->emig_valid = false;
    return setup_windows(c);
}

extern "C" void pg_comm_destroy(pgpu_ctx* c) {
    for (int r = 0; r < 64; r++)
        if (c->x_peer_win[r]) { cudaIpcCloseMemHandle(c->x_peer_win[r]); c->x_peer_win[r] = nullptr; }
    if (c->x_win) { cudaFree(c->x_win); c->x_win = nullptr; }
    if (c->nccl_comm && g_nccl.lib) g_nccl.CommDestroy((ncclComm_t)c->nccl_comm);
    c->nccl_comm = nullptr;
}

int pg_allreduce_sum_f64(pgpu_ctx* c, double* buf, long long n) {
    if (!c->nccl_comm) PG_FAIL(c, PGPU_ERR_STATE, "nranks > 1 but pgpu_comm_init has not been called");
    PG_NCCL(c, g_nccl.AllReduce(buf, buf, (size_t)n, ncclDouble, ncclSum, (ncclComm_t)c->nccl_comm, c->stream));
    return PGPU_OK;
}

static int ensure_xbuf(pgpu_ctx* c, long long need) {
    if (need <= c->xbuf_capacity) return PGPU_OK;
    cudaFree(c->d_sendbuf); cudaFree(c->d_recvbuf);
    c->d_sendbuf = c->d_recvbuf = nullptr;
    long long cap = need + need / 2 + 1024;
    /* (96 bytes per slot: the buffers also stage host-format records for pgpu_exchange_pack) */
    PG_CUDA(c, cudaMalloc(&c->d_sendbuf, 96 * (size_t)cap));
    PG_CUDA(c, cudaMalloc(&c->d_recvbuf, 96 * (size_t)cap));
    c->xbuf_capacity = cap;
    return PGPU_OK;
}

extern "C" int pgpu_exchange_pack(pgpu_ctx* c, void* sendbuf_aos, int64_t capacity, int64_t* sendcount) {
    if (!c || !sendcount) return PGPU_ERR_INVALID;
    int rc;
    if ((rc = pg_prepare(c))) return rc;
    const int nr = c->cfg.nranks > 1 ? c->cfg.nranks : 1;
    for (int r = 0; r < nr; r++) sendcount[r] = 0;
    if (nr == 1 || c->np == 0) return PGPU_OK;
    if (!c->dev.cell_procs) PG_FAIL(c, PGPU_ERR_STATE, "exchange_pack: CELL_PROCS not set");
    if (c->cell_procs_max >= nr) PG_FAIL(c, PGPU_ERR_INVALID, "exchange_pack: CELL_PROCS holds rank %d but there are %d ranks", c->cell_procs_max, nr);
    if (!c->d_xcount) PG_CUDA(c, cudaMalloc((void**)&c->d_xcount, sizeof(unsigned long long) * PG_XROW * (4 + 64)));
    if ((rc = pg_reset_status(c))) return rc;
    if ((rc = pg_resolve_vel(c))) return rc;
    pg_refresh_dev(c);
    unsigned long long* xcount = c->d_xcount;
    unsigned long long* xoff = c->d_xcount + PG_XROW;
    unsigned long long* xcursor = c->d_xcount + 2 * PG_XROW;
    const unsigned int blocks = (unsigned int)((c->np + 255) / 256);
    PG_CUDA(c, cudaMemsetAsync(xcount, 0, sizeof(unsigned long long) * 3 * PG_XROW, c->stream));
    xcount_kernel<<<blocks, 256, 0, c->stream>>>(c->dev, xcount);
    unsigned long long h[PG_XROW];
    PG_CUDA(c, cudaMemcpyAsync(h, xcount, sizeof(h), cudaMemcpyDeviceToHost, c->stream));
    PG_CUDA(c, cudaStreamSynchronize(c->stream));
    unsigned long long soff[65] = {0};
    for (int r = 0; r < nr; r++) {
        sendcount[r] = (int64_t)h[r];
        soff[r + 1] = soff[r] + h[r];
    }
    const long long nsend = (long long)soff[nr];
    c->emig_valid = false;
    if (nsend == 0) return PGPU_OK;
    if (nsend > capacity) PG_FAIL(c, PGPU_ERR_INVALID, "exchange_pack: %lld emigrants do not fit the buffer (%lld records)", nsend, (long long)capacity);
    if (!sendbuf_aos) PG_FAIL(c, PGPU_ERR_INVALID, "exchange_pack: null buffer");
    if ((rc = ensure_xbuf(c, nsend))) return rc;
    PG_CUDA(c, cudaMemcpyAsync(xoff, soff, sizeof(unsigned long long) * 64, cudaMemcpyHostToDevice, c->stream));
    xpack_kernel<<<blocks, 256, 0, c->stream>>>(c->dev, (double*)c->d_sendbuf, xoff, xcursor);
    xrec_to_aos_kernel<<<(unsigned int)((nsend + 255) / 256), 256, 0, c->stream>>>((const double*)c->d_sendbuf, nsend, (unsigned char*)c->d_recvbuf,
                                                                                 c->cfg.rank, c->id_counter, c->cfg.track_id ? 1 : 0);
    PG_CUDA(c, cudaMemcpyAsync(sendbuf_aos, c->d_recvbuf, 96 * (size_t)nsend, cudaMemcpyDeviceToHost, c->stream));
    c->launches += 3;
    c->maybe_tombstones = true;
    c->sorted = false;
    c->hist_valid = false;
    return pg_finish(c, "pgpu_exchange_pack");
}

int pg_exchange_impl(pgpu_ctx* c) {
    const int nr = c->cfg.nranks;
    if (nr <= 1) return PGPU_OK;
    if (!c->nccl_comm) PG_FAIL(c, PGPU_ERR_STATE, "exchange: pgpu_comm_init has not been called");
    if (nr != c->comm_nranks || c->cfg.rank != c->comm_rank) PG_FAIL(c, PGPU_ERR_STATE, "exchange: config rank/nranks differ from the communicator");
    if (!c->dev.cell_procs) PG_FAIL(c, PGPU_ERR_STATE, "exchange: CELL_PROCS not set");
    if (c->cell_procs_max >= nr) PG_FAIL(c, PGPU_ERR_INVALID, "exchange: CELL_PROCS holds rank %d but there are %d ranks", c->cell_procs_max, nr);
    int rc;
    if ((rc = pg_resolve_vel(c))) return rc;
    pg_refresh_dev(c);
    ncclComm_t comm = (ncclComm_t)c->nccl_comm;
    unsigned long long* xcount = c->d_xcount;                 /* [PG_XROW] my counts per destination (+ list length, overflow) */
    unsigned long long* xoff = c->d_xcount + PG_XROW;         /* [PG_XROW] */
    unsigned long long* xcursor = c->d_xcount + 2 * PG_XROW;  /* [PG_XROW] */
    unsigned long long* xall = c->d_xcount + 4 * PG_XROW;     /* [nr][PG_XROW] everyone's rows */
    unsigned int blocks = (unsigned int)((c->np + 255) / 256);
    /* fast path: ADVECT has just recorded the emigrants (list + counts) and the histogram of the survivors is
     * valid; nothing has touched the store since.  Every rank takes the same branch (same call sequence). */
    bool fast = c->emig_valid && c->hist_valid;
    if (fast && c->x_counts_valid && !c->has_dtrim) {
        /* ---- the per-step path: no host round trip (see the XPlan kernels above) ---- */
        c->emig_valid = false;
        XPlan sp, rp;
        memset(&sp, 0, sizeof(sp));
        memset(&rp, 0, sizeof(rp));
        sp.nr = rp.nr = nr;
        sp.rank = rp.rank = c->cfg.rank;
        unsigned long long so = 0, ro = 0;
        /* Peers: the ranks that share a face with this one, plus any rank particles went to or came from last step.
         * (Both conditions are symmetric, so every Send has its Recv.)  A rank that is neither gets no message --
         * at N ranks in strips that is N-3 of them, each a launch-latency-bound transfer on the critical path -- and
         * an emigrant for it fails the step loudly (a particle crossed a whole rank's region in one step):
         * PANTERA_GPU_XCHG_ALLPAIRS=1 brings the all-pairs messages back. */
        if (c->x_allpairs < 0) {
            const char* e = getenv("PANTERA_GPU_XCHG_ALLPAIRS");
            c->x_allpairs = (e && *e == '1') ? 1 : 0;
        }
        bool peer[64];
        for (int r = 0; r < nr; r++)
            peer[r] = r != c->cfg.rank && (c->x_allpairs || !c->x_adjacent_known || (int)c->x_adjacent.size() != nr || c->x_adjacent[r] ||
                                           c->x_prev_send[r] > 0 || c->x_prev_recv[r] > 0);
        /* window pairs (peer memory) and NCCL pairs; both ends of a pair decide alike: the mapping was agreed on at
         * pgpu_comm_init and the capacities derive from counts both ends know */
        const unsigned long long seq = ++c->x_seq;
        sp.seq = rp.seq = seq;
        const size_t slot_d = (size_t)(c->x_win_recs + 1) * XREC_DOUBLES; /* doubles per message slot of a window */
        bool win[64];
        unsigned int soff[64] = {0}, roff[64] = {0};
        unsigned long long imm = 0; /* upper bound of the immigrants */
        for (int r = 0; r < nr; r++) {
            sp.on[r] = rp.on[r] = peer[r] ? 1 : 0;
            win[r] = false;
            if (!peer[r]) continue;
            sp.cap[r] = (unsigned int)std::min<double>(4.0e9, c->x_headroom * (double)c->x_prev_send[r] + 8192.0);
            rp.cap[r] = (unsigned int)std::min<double>(4.0e9, c->x_headroom * (double)c->x_prev_recv[r] + 8192.0);
            imm += rp.cap[r];
            win[r] = c->x_peer_win[r] && (long long)sp.cap[r] <= c->x_win_recs && (long long)rp.cap[r] <= c->x_win_recs;
            if (win[r]) continue;
            soff[r] = (unsigned int)so;
            roff[r] = (unsigned int)ro;
            so += (unsigned long long)sp.cap[r] + 1;
            ro += (unsigned long long)rp.cap[r] + 1;
        }
        if ((rc = ensure_xbuf(c, (long long)std::max(so, ro)))) return rc;
        /* (pg_advect_impl has made room for the immigrants before it recorded the emigrant list: a re-allocation
         *  here would drop that list) */
        if (c->np_ub + (long long)imm > c->capacity)
            PG_FAIL(c, PGPU_ERR_STATE, "exchange: %lld slots in use + %llu message slots exceed the capacity %lld (raise pgpu_config.capacity)",
                    c->np_ub, imm, c->capacity);
        pg_refresh_dev(c);
        for (int r = 0; r < nr; r++) {
            if (!peer[r]) continue;
            if (win[r]) {
                /* window layout: 128 doubles of flags, then [parity][source rank] message slots */
                const size_t par = (size_t)(seq & 1ULL) * (size_t)nr;
                sp.msg[r] = (double*)c->x_peer_win[r] + 128 + (par + (size_t)c->cfg.rank) * slot_d;
                sp.flag[r] = (unsigned long long*)c->x_peer_win[r] + c->cfg.rank;
                rp.msg[r] = (double*)c->x_win + 128 + (par + (size_t)r) * slot_d;
                rp.flag[r] = (unsigned long long*)c->x_win + r;
            } else {
                sp.msg[r] = (double*)c->d_sendbuf + (size_t)soff[r] * XREC_DOUBLES;
                rp.msg[r] = (double*)c->d_recvbuf + (size_t)roff[r] * XREC_DOUBLES;
            }
        }
        if (c->x_trace < 0) {
            const char* e = getenv("PANTERA_GPU_XCHG_TRACE");
            c->x_trace = (e && *e == '1') ? 1 : 0;
            if (c->x_trace) for (int i = 0; i < 4; i++) cudaEventCreate(&c->xev[i]);
        } else if (c->x_trace && c->x_traced >= 0) { /* the previous step's events have completed (pg_finish synchronised) */
            float ms;
            for (int i = 0; i < 3; i++)
                if (cudaEventElapsedTime(&ms, c->xev[i], c->xev[i + 1]) == cudaSuccess) c->x_ms[i] += ms;
            c->x_traced++;
        }
        if (c->x_trace) cudaEventRecord(c->xev[0], c->stream);
        c->dev.emig_list = c->d_emig_list; /* (emig_cap is read by the pack kernel) */
        PG_CUDA(c, cudaMemsetAsync(xcursor, 0, sizeof(unsigned long long) * PG_XROW, c->stream));
        xpack_fast_kernel<<<148 * 4, 256, 0, c->stream>>>(c->dev, c->d_emig_list, xcount, sp, xcursor);
        xheader_kernel<<<1, 64, 0, c->stream>>>(c->dev, xcount, sp);
        c->dev.emig_list = nullptr;
        if (c->x_trace) cudaEventRecord(c->xev[1], c->stream);
        bool any_nccl = false, any_win = false;
        for (int r = 0; r < nr; r++) {
            any_nccl |= peer[r] && !win[r];
            any_win |= peer[r] && win[r];
        }
        if (any_nccl) {
            PG_NCCL(c, g_nccl.GroupStart());
            for (int r = 0; r < nr; r++) {
                if (!peer[r] || win[r]) continue;
                PG_NCCL(c, g_nccl.Send(sp.msg[r], ((size_t)sp.cap[r] + 1) * XREC_DOUBLES, ncclDouble, r, comm, c->stream));
                PG_NCCL(c, g_nccl.Recv(rp.msg[r], ((size_t)rp.cap[r] + 1) * XREC_DOUBLES, ncclDouble, r, comm, c->stream));
            }
            PG_NCCL(c, g_nccl.GroupEnd());
        }
        if (any_win) xwait_kernel<<<1, 64, 0, c->stream>>>(c->dev, rp);
        if (c->x_trace) cudaEventRecord(c->xev[2], c->stream);
        unsigned int maxcap = 1;
        for (int r = 0; r < nr; r++) maxcap = std::max(maxcap, rp.cap[r]);
        PgPlanes p = c->planes[c->cur];
        p.dtrim = nullptr; /* every particle has DTRIM == DT right after ADVECT */
        dim3 grid(std::min<unsigned int>((maxcap + 255) / 256, 148u * 4u), (unsigned int)nr);
        xunpack_fast_kernel<<<grid, 256, 0, c->stream>>>(c->dev, rp, p);
        xfinish_kernel<<<1, 32, 0, c->stream>>>(c->dev, rp);
        if (c->x_trace) cudaEventRecord(c->xev[3], c->stream);
        c->launches += 4 + (any_win ? 1 : 0);
        c->np_ub += (long long)imm;
        c->np_dev_valid = true;
        c->x_plan_pending = true; /* pg_finish turns status->x_sent / x_recv into the next plan */
        c->maybe_tombstones = true;
        c->sorted = false;
        c->owned_clean = true;
        c->id_counter = std::max(c->id_counter, c->np_ub);
        PG_CUDA(c, cudaGetLastError());
        return PGPU_OK;
    }
    c->emig_valid = false;
    c->x_counts_valid = false;
    std::vector<unsigned long long> h((size_t)nr * PG_XROW);
    for (int attempt = 0; attempt < 2; attempt++) {
        PG_CUDA(c, cudaMemsetAsync(xoff, 0, sizeof(unsigned long long) * 2 * PG_XROW, c->stream));
        if (!fast) {
            PG_CUDA(c, cudaMemsetAsync(xcount, 0, sizeof(unsigned long long) * PG_XROW, c->stream));
            if (c->np > 0) {
                xcount_kernel<<<blocks, 256, 0, c->stream>>>(c->dev, xcount);
                c->launches++;
            }
        }
        /* MPI_ALLTOALL of the counts -> all-gather of the rows */
        PG_NCCL(c, g_nccl.AllGather(xcount, xall, PG_XROW, ncclUint64, comm, c->stream));
        PG_CUDA(c, cudaMemcpyAsync(h.data(), xall, sizeof(unsigned long long) * (size_t)nr * PG_XROW, cudaMemcpyDeviceToHost, c->stream));
        PG_CUDA(c, cudaStreamSynchronize(c->stream));
        if (!fast) break;
        /* a rank whose emigrants did not fit its list sends everybody (itself included) through the full sweep */
        bool overflow = false;
        for (int r = 0; r < nr; r++) overflow |= h[(size_t)r * PG_XROW + PG_XROW_OVF] != 0;
        if (!overflow) break;
        fast = false;
    }
    std::vector<unsigned long long> soff(65, 0), roff(65, 0);
    for (int r = 0; r < nr; r++) {
        soff[r + 1] = soff[r] + h[(size_t)c->cfg.rank * PG_XROW + r];
        roff[r + 1] = roff[r] + h[(size_t)r * PG_XROW + c->cfg.rank];
    }
    for (int r = 0; r < nr; r++) { /* the counts that size the next step's fixed-capacity messages */
        c->x_prev_send[r] = (long long)h[(size_t)c->cfg.rank * PG_XROW + r];
        c->x_prev_recv[r] = (long long)h[(size_t)r * PG_XROW + c->cfg.rank];
    }
    c->x_counts_valid = true;
    c->owned_clean = true;
    long long nsend = (long long)soff[nr], nrecv = (long long)roff[nr];
    long long need = std::max(nsend, nrecv);
    if ((rc = ensure_xbuf(c, need))) return rc;
    if (nsend > 0) {
        /* (packed before the store can be re-allocated below: the list holds slots of the current planes) */
        PG_CUDA(c, cudaMemcpyAsync(xoff, soff.data(), sizeof(unsigned long long) * 64, cudaMemcpyHostToDevice, c->stream));
        if (fast) {
            xpack_list_kernel<<<(unsigned int)((nsend + 255) / 256), 256, 0, c->stream>>>(c->dev, c->d_emig_list, nsend,
                                                                                          (double*)c->d_sendbuf, xoff, xcursor);
        } else {
            xpack_kernel<<<blocks, 256, 0, c->stream>>>(c->dev, (double*)c->d_sendbuf, xoff, xcursor);
        }
        c->launches++;
        c->maybe_tombstones = true;
    }
    if (!fast) c->hist_valid = false;
    if ((rc = pg_ensure_capacity(c, c->np + nrecv))) return rc; /* (a re-allocation drops the histogram) */
    pg_refresh_dev(c);
    /* MPI_ALLTOALLV -> one grouped send/recv to every peer with traffic */
    PG_NCCL(c, g_nccl.GroupStart());
    for (int r = 0; r < nr; r++) {
        if (r == c->cfg.rank) continue;
        unsigned long long ns = soff[r + 1] - soff[r], nv = roff[r + 1] - roff[r];
        if (ns) PG_NCCL(c, g_nccl.Send((const double*)c->d_sendbuf + soff[r] * XREC_DOUBLES, (size_t)ns * XREC_DOUBLES, ncclDouble, r, comm, c->stream));
        if (nv) PG_NCCL(c, g_nccl.Recv((double*)c->d_recvbuf + roff[r] * XREC_DOUBLES, (size_t)nv * XREC_DOUBLES, ncclDouble, r, comm, c->stream));
    }
    PG_NCCL(c, g_nccl.GroupEnd());
    if (nrecv > 0) {
        PgPlanes p = c->planes[c->cur];
        xunpack_kernel<<<(unsigned int)((nrecv + 255) / 256), 256, 0, c->stream>>>(c->dev, (const double*)c->d_recvbuf, nrecv, p, c->np,
                                                                                   c->hist_valid ? 1 : 0);
        c->launches++;
        PG_CUDA(c, cudaMemcpyAsync(&c->h_status->flag_dtrim, &c->d_status->flag_dtrim, 2 * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
        PG_CUDA(c, cudaStreamSynchronize(c->stream));
        if (c->h_status->flag_ie) c->ie_seen = true;
        if (c->h_status->flag_dtrim && !c->has_dtrim) {
            if (c->np > 0) fill_kernel<<<(unsigned int)((c->np + 255) / 256), 256, 0, c->stream>>>(p.dtrim, c->np, c->cfg.dt);
            c->has_dtrim = true;
        }
        c->np += nrecv;
        c->id_counter = std::max(c->id_counter, c->np);
    }
    c->sorted = false;
    c->np_ub = c->np;
    pg_set_np_cur(c); /* the count changed on the host: bring the device's copy along */
    pg_refresh_dev(c);
    PG_CUDA(c, cudaGetLastError());
    return PGPU_OK;
}

/* slots the next fast EXCHANGE may append (sum of the receive capacities); 0 when the slow path will run */
long long pg_exchange_recv_slots(const pgpu_ctx* c) {
    if (c->cfg.nranks <= 1 || !c->x_counts_valid) return 0;
    double ro = 0;
    for (int r = 0; r < c->cfg.nranks; r++)
        if (r != c->cfg.rank) ro += std::min<double>(4.0e9, c->x_headroom * (double)c->x_prev_recv[r] + 8192.0) + 1.0;
    return (long long)ro;
}
