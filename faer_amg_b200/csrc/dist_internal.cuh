// dist_internal.cuh -- structures shared by dist.cu (distributed cycle / PCG) and dist_setup.cu
// (distributed hierarchy construction).  Not part of the C ABI.
#pragma once
#include <nccl.h>

#include "mg_internal.cuh"
#include "p2p_dev.cuh"

namespace famg {

struct NcclApi {
    void *handle = nullptr;
    decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
    decltype(&ncclCommInitRank) CommInitRank = nullptr;
    decltype(&ncclCommDestroy) CommDestroy = nullptr;
    decltype(&ncclGetErrorString) GetErrorString = nullptr;
    decltype(&ncclGroupStart) GroupStart = nullptr;
    decltype(&ncclGroupEnd) GroupEnd = nullptr;
    decltype(&ncclSend) Send = nullptr;
    decltype(&ncclRecv) Recv = nullptr;
    decltype(&ncclAllReduce) AllReduce = nullptr;
    decltype(&ncclBroadcast) Broadcast = nullptr;
    decltype(&ncclAllGather) AllGather = nullptr;
};
extern NcclApi g_nccl;
famg_status nccl_load();

#define NCCL_TRY(expr)                                                                              \
    do {                                                                                            \
        ncclResult_t r__ = (expr);                                                                  \
        if (r__ != ncclSuccess) FAMG_FAIL(FAMG_ERR_COMM, "%s failed: %s", #expr, ::famg::g_nccl.GetErrorString(r__)); \
    } while (0)

}  // namespace famg

// One process per GPU: nlocal == 1 and `rank` is this process' rank.  The setup phase can also be
// run by ONE process that hosts all `nranks` virtual ranks on one device (nlocal == nranks, no NCCL):
// every exchange becomes a device-to-device copy.  That mode exists so that the distributed
// hierarchy construction is testable on a single GPU; it cannot run the distributed cycle (the
// peer-memory flag protocol needs concurrently running ranks).
struct famg_comm {
    famg_ctx *ctx = nullptr;
    int nranks = 1, rank = 0;
    int nlocal = 1;
    ncclComm_t comm = nullptr;
    cudaEvent_t ev_packed = nullptr, ev_halo = nullptr;      // fork / join of the communication stream (inside captures)
    cudaEvent_t ev_packed_e = nullptr, ev_halo_e = nullptr;  // the same for eager (uncaptured) work
    int vrank(int li) const { return nlocal > 1 ? li : rank; }
};

namespace famg {

// ---------------------------------------------------------------- halo plans (P2PPlanDev: p2p_dev.cuh)
// Small collectives over the same arenas: sum-all-reduce of <= 4 doubles (PCG dot products) and the
// all-gather of the restricted residual at the replicated-level transition.  Every rank stores its
// contribution into every peer's slot, publishes the epoch, waits for all peers, then reduces in
// rank order -- every rank forms bit-identical sums.
constexpr int P2P_AR_WIDTH = 4;
struct P2PCollDev {
    int nranks, rank;
    // all-reduce
    double *ar_rslots[2][P2P_MAX_NB + 1];            // peer p's slot array (parity), indexed [rank*W + w]
    unsigned long long *ar_rflag[P2P_MAX_NB + 1];    // my flag slot on peer p
    const unsigned long long *ar_lflag;              // my flag array (one per peer)
    const double *ar_lslots[2];
    unsigned long long *ar_epoch;
    // all-gather
    double *ag_rbuf[2][P2P_MAX_NB + 1];              // peer p's gather buffer (parity)
    unsigned long long *ag_rflag[P2P_MAX_NB + 1];
    const unsigned long long *ag_lflag;
    const double *ag_lbuf[2];
    unsigned long long *ag_epoch;
    unsigned int *ag_done;
    int *err;
};

struct HaloPlan {
    bool p2p = false;
    P2PPlanDev dev{};
    size_t arena_flags = 0, arena_recv[2] = {0, 0};  // byte offsets inside this rank's arena
    int nloc = 0, nghost = 0;
    std::vector<int> recv_cnt, recv_off, send_cnt, send_off;
    int total_send = 0;
    int *d_send_idx = nullptr;
    double *d_sendbuf = nullptr;
    bool any = false;
    // rows outside [push_lo, push_hi) hold every entry some peer needs (the send list): a producer computes them
    // first, so that the exchange of its output can start while the remaining rows are still being computed
    int push_lo = 0, push_hi = 0;
    std::vector<int> send_idx_host;  // the send list (local ids, grouped by peer)
    // fused push (set up with the peer-memory arenas): device copy of `dev`, and per boundary row -- rows [0, push_lo) then
    // [push_hi, nloc) -- the (neighbour, slot) its value goes to, or -1.  fusable: every sent row has exactly one destination.
    P2PPlanDev *d_dev = nullptr;
    int *d_push_map = nullptr;
    bool fusable = false;
};

// One (virtual) rank's share of a row-partitioned operator.  Columns are global indices until the
// halo plan is built (dmat_finalize); afterwards they are renumbered to [owned | ghost] in place --
// the order of the entries inside a row is never changed, so every row still accumulates in
// ascending *global* column order, like the undistributed operator.
struct DistOp {
    famg_csr *local = nullptr;
    HaloPlan halo;
    int ib = 0, ie = 0;             // rows [ib, ie) reference no ghost column
    bool local_cols = false;        // columns renumbered to [owned | ghost]
    std::vector<int> ghost_gid;     // ascending global ids of the ghost columns
    int *d_ghost_gid = nullptr;
};

struct DistLevel {
    int64_t r0 = 0, r1 = 0;  // owned rows of this level
    DistOp *A = nullptr, *R = nullptr, *P = nullptr;  // parts of retained famg_dmat handles.  R: rows of level l+1
                                                      // (owned), cols level l.  P: rows level l, cols level l+1
    double *d = nullptr;     // owned slice of the Diag smoother
    double *x = nullptr, *b = nullptr, *t = nullptr;  // work vectors with ghost tails
    int64_t ld = 0;
};

void halo_free(HaloPlan &h);
void distop_free(DistOp &o);

}  // namespace famg

// Row-partitioned sparse matrix: one DistOp per virtual rank hosted by this process.
struct famg_dmat {
    famg_comm *comm = nullptr;
    int64_t nrows = 0, ncols = 0;           // global shape
    std::vector<int64_t> rsplit, csplit;    // nranks + 1 each
    std::vector<famg::DistOp> part;         // nlocal
    bool finalized = false;                 // plans built (or columns declared replicated)
    bool replicated_cols = false;           // the consumer vector is replicated: global column ids, no halo
    std::atomic<int> refs{1};
};

struct famg_dist_mg {
    famg_comm *comm = nullptr;
    famg_mg *global = nullptr;  // replicated tail: levels tail_first.. of this multigrid
    int tail_first = 0;         // index inside `global` of the first replicated level
    bool owns_tail = false;
    int lrep = 0;               // number of distributed levels == index of the first replicated level
    std::vector<famg::DistLevel> lv;
    std::vector<famg_dmat *> keep;             // retained distributed operators (level-wise construction)
    std::vector<std::vector<int64_t>> splits;  // per level (0..lrep), nranks+1
    // transition buffers: gathered rhs / replicated solution of level lrep, and this rank's piece
    // of the restricted residual before the gather
    double *g_f = nullptr, *g_v = nullptr, *fc_loc = nullptr;
    // PCG work vectors (with ghost tail for p)
    double *pcg = nullptr; int64_t pcg_ld = 0;
    // peer-memory exchange state
    unsigned char *arena = nullptr; size_t arena_bytes = 0;
    std::vector<void *> peer_arena;  // IPC-mapped arenas of the other ranks
    int *d_p2p_err = nullptr;
    bool p2p = false;
    famg::P2PCollDev coll{};
    bool p2p_coll = false;
    // the distributed cycle (kernels + peer-memory exchanges) is captured into one CUDA graph per
    // (out, rhs) pair and replayed; disabled on the first failure
    std::map<std::tuple<const void *, const void *, const void *>, GraphEntry> graphs;  // (out, rhs, dot partials)
    bool use_graph = true;
    // producer-side halo exchange on the communication stream, overlapped with the rows nobody is waiting for
    // (peer-memory mode; FAMG_OVERLAP=0 keeps the exchange in front of every apply)
    bool overlap = true;
    int overlap_mode = 3;       // FAMG_OVERLAP: 0 exchange in front of every apply (round-1 order), 1 producer-side exchange after
                                // the producing kernel, 2 producer-side exchange running next to the producing kernel, 3 fused
                                // push: the producing kernel stores its boundary rows into the neighbours and publishes them
    int fuse_max_boundary_pct = 20;  // fused push only for plans whose boundary rows are at most this share of the slab (FAMG_FUSE_MAX_BOUNDARY_PCT)
    int reserve_ctas = 16;      // CTA slots the persistent producer leaves to the exchange kernel, which uses at most as many (FAMG_RESERVE_CTAS)
    unsigned *d_sig = nullptr;  // finished-boundary-chunk counter of the producer kernel in flight
    bool pending = false;       // an exchange is in flight on the communication stream
    cudaEvent_t pending_ev = nullptr;
    int split_min_rows = 16384; // smallest remainder worth a launch of its own (FAMG_OVERLAP_MIN_ROWS)
};

namespace famg {

// ---------------------------------------------------------------- setup-time exchange (dist_setup.cu)
struct XSeg { void *p = nullptr; size_t bytes = 0; };
// host metadata all-to-all: send[li][p * w + j] goes to virtual rank p; recv[li][p * w + j] came from p
famg_status xchg_meta(famg_comm *cm, int w, const std::vector<std::vector<int64_t>> &send, std::vector<std::vector<int64_t>> &recv);
// every virtual rank contributes w values; all[p * w + j] on return (identical on every rank)
famg_status xchg_allgather_meta(famg_comm *cm, int w, const std::vector<std::vector<int64_t>> &mine, std::vector<int64_t> &all);
// device data all-to-all on the context stream: send[li][p] -> recv[p'][..] of the peer; sizes must agree
famg_status xchg_data(famg_comm *cm, const std::vector<std::vector<XSeg>> &send, const std::vector<std::vector<XSeg>> &recv);

// builds the halo plan of every part (collective) and renumbers the columns in place
famg_status dmat_finalize(famg_dmat *m, bool replicated_cols);
void dmat_release(famg_dmat *m);
// fill the ghost tail of x_ext (one vector per virtual rank) -- setup-time exchange (pack + xchg_data)
famg_status setup_halo_exchange(famg_comm *cm, std::vector<DistOp *> &ops, std::vector<double *> &x_ext);

}  // namespace famg
