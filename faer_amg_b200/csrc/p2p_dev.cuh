// p2p_dev.cuh -- device-side view of a peer-memory halo plan and the system-scope synchronisation primitives,
// shared by the exchange kernels (dist.cu) and the SpMV-family kernels that push their boundary rows themselves (spmv.cu).
#pragma once
#include <cuda_runtime.h>

namespace famg {

// Peer-memory halo exchange (NVLink loads/stores instead of NCCL send/recv): every rank owns an
// "arena" exported with CUDA IPC; per plan it holds one flag per peer, an epoch counter and two
// receive buffers (parity = epoch & 1).  The producer of a vector stores this rank's boundary entries straight
// into its neighbours' receive buffers, fences, and publishes the epoch in their flag slots; the
// consumer spins on its own flag slots (acquire, with a 20 s timeout) and moves the ghosts into the
// vector's tail.  Epochs live in device memory, so the sequence replays inside CUDA graphs.
constexpr int P2P_MAX_NB = 8;
struct P2PPlanDev {
    int nnb;                                  // neighbours (peers we exchange flags with, both ways)
    int nghost;
    double *rdst[2][P2P_MAX_NB];              // where my entries go on neighbour nb (per parity)
    unsigned long long *rflag[P2P_MAX_NB];    // my slot in neighbour nb's flag array
    const unsigned long long *lflag[P2P_MAX_NB];  // neighbour nb's slot in my flag array
    int soff[P2P_MAX_NB], scnt[P2P_MAX_NB];   // my pack-list range for neighbour nb
    unsigned long long *epoch;                // this plan's exchange counter (device)
    unsigned int *done;                       // blocks of the pack kernel that have finished their stores
    int total_send;
    const double *lrecv[2];                   // my receive buffers
    int *err;                                 // set on spin timeout
};
constexpr int PUSH_SLOT_BITS = 28;            // push map entry: (neighbour << 28) | slot in that neighbour's receive buffer

#ifdef __CUDACC__
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
#endif

}  // namespace famg
