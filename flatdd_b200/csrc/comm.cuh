// comm.cuh — multi-GPU plumbing of the sharded state: one process per GPU.
//   * NCCL (loaded with dlopen so the library has no link-time dependency and shares the NCCL
//     instance of the hosting process, e.g. torch's) for the rendezvous, barriers in stream
//     order and the send/recv exchange path;
//   * CUDA IPC so that every rank maps the state buffers of all peers: the P2P exchange kernel
//     reads the partner's half shard straight over NVLink (NVSwitch: full bandwidth to any peer).
#pragma once

#include "kernels.cuh"

#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>

#include <stdexcept>
#include <string>

namespace fddb200 {

struct NcclApi {
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;

    static NcclApi& get() {
        static NcclApi api = load();
        return api;
    }

private:
    template <class F> static void sym(void* h, F& f, const char* name) {
        f = reinterpret_cast<F>(dlsym(h, name));
        if (f == nullptr) throw std::runtime_error(std::string("NCCL symbol missing: ") + name);
    }
    static NcclApi load() {
        NcclApi a;
        // RTLD_NOLOAD first: reuse the instance the process already has (same SONAME)
        a.handle = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);
        if (a.handle == nullptr) a.handle = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (a.handle == nullptr) a.handle = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (a.handle == nullptr) throw std::runtime_error(std::string("cannot load NCCL: ") + dlerror());
        sym(a.handle, a.GetUniqueId, "ncclGetUniqueId");
        sym(a.handle, a.CommInitRank, "ncclCommInitRank");
        sym(a.handle, a.CommDestroy, "ncclCommDestroy");
        sym(a.handle, a.GetErrorString, "ncclGetErrorString");
        sym(a.handle, a.AllReduce, "ncclAllReduce");
        sym(a.handle, a.AllGather, "ncclAllGather");
        sym(a.handle, a.Send, "ncclSend");
        sym(a.handle, a.Recv, "ncclRecv");
        sym(a.handle, a.GroupStart, "ncclGroupStart");
        sym(a.handle, a.GroupEnd, "ncclGroupEnd");
        return a;
    }
};

// ------------------------------------------------------------------------------------------------
// exchange kernels
// ------------------------------------------------------------------------------------------------
// SWAP(global physical bit, local physical bit `pl`) seen from one shard whose global bit has the
// value `myBit`: the half with bit pl == myBit stays, the other half is replaced by the
// partner's half with bit pl == myBit.  Out of place (mine -> out) so that the partner can read
// `mine` at the same time; half of the reads cross NVLink, 16-byte vectors, 4 in flight per thread.
template <int U>
__global__ void __launch_bounds__(256) exchange_p2p_kernel(const double2* __restrict__ mine, const double2* __restrict__ partner,
                                                           double2* __restrict__ out, uint64_t n, int pl, int myBit) {
    // h enumerates the half space (index without bit pl); every h yields one kept and one traded amplitude
    const uint64_t half = n >> 1;
    const uint64_t lowMask = (uint64_t{1} << pl) - 1;
    const uint64_t keepBit = static_cast<uint64_t>(myBit) << pl;
    const uint64_t flip = uint64_t{1} << pl;
    const uint64_t stride = static_cast<uint64_t>(gridDim.x) * blockDim.x;
    uint64_t h = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    for (; h + (U - 1) * stride < half; h += U * stride) {
        double2 far[U], near[U];
        uint64_t at[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const uint64_t x = h + u * stride;
            at[u] = ((x & ~lowMask) << 1) | (x & lowMask) | keepBit; // index with bit pl == myBit
            far[u] = ld_stream(partner + at[u]);                    // crosses NVLink
        }
#pragma unroll
        for (int u = 0; u < U; ++u) near[u] = ld_stream(mine + at[u]);
#pragma unroll
        for (int u = 0; u < U; ++u) {
            st_stream(out + at[u], near[u]);
            st_stream(out + (at[u] ^ flip), far[u]);
        }
    }
    for (; h < half; h += stride) {
        const uint64_t i = ((h & ~lowMask) << 1) | (h & lowMask) | keepBit;
        out[i] = mine[i];
        out[i ^ flip] = partner[i];
    }
}

// ---- the same exchange with its own cross-GPU ordering (no NCCL call on the data path) -----------------------------------
// Each rank owns two epoch words in peer-mapped memory: flags[0] = "everything I launched before exchange e has completed: my
// current buffer holds the state", flags[1] = "I have finished reading my partner's buffer in exchange e".  The kernel
//   1. publishes flags[0] = e (it runs after the rank's earlier launches, so their writes are complete) and waits for the
//      partner's flags[0] >= e before it touches the partner's buffer;
//   2. copies (same body as above);
//   3. its last CTA publishes flags[1] = e and does not exit before the partner's flags[1] >= e: the next launch of this
//      rank's stream overwrites the buffer the partner has been reading.
// Replaces two stream-ordered one-element ncclAllReduce "barriers" (about 55 us of the 164 us per exchange at 8 GPUs and
// 64 MiB messages).  Epochs only grow; every rank takes part in every exchange, so they agree on e.
// (st_release_sys / ld_acquire_sys: kernels.cuh)
template <int U>
__global__ void __launch_bounds__(256) exchange_p2p_flag_kernel(const double2* __restrict__ mine, const double2* __restrict__ partner,
                                                                double2* __restrict__ out, uint64_t n, int pl, int myBit, uint32_t* myFlags,
                                                                const uint32_t* partnerFlags, uint32_t epoch, unsigned int* ctaCounter) {
    if (threadIdx.x == 0) {
        if (blockIdx.x == 0) st_release_sys(myFlags, epoch);
        while (ld_acquire_sys(partnerFlags) < epoch) __nanosleep(64);
    }
    __syncthreads();
    const uint64_t half = n >> 1;
    const uint64_t lowMask = (uint64_t{1} << pl) - 1;
    const uint64_t keepBit = static_cast<uint64_t>(myBit) << pl;
    const uint64_t flip = uint64_t{1} << pl;
    const uint64_t stride = static_cast<uint64_t>(gridDim.x) * blockDim.x;
    uint64_t h = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    for (; h + (U - 1) * stride < half; h += U * stride) {
        double2 far[U], near[U];
        uint64_t at[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const uint64_t x = h + u * stride;
            at[u] = ((x & ~lowMask) << 1) | (x & lowMask) | keepBit;
            far[u] = ld_stream(partner + at[u]); // crosses NVLink
        }
#pragma unroll
        for (int u = 0; u < U; ++u) near[u] = ld_stream(mine + at[u]);
#pragma unroll
        for (int u = 0; u < U; ++u) {
            st_stream(out + at[u], near[u]);
            st_stream(out + (at[u] ^ flip), far[u]);
        }
    }
    for (; h < half; h += stride) {
        const uint64_t i = ((h & ~lowMask) << 1) | (h & lowMask) | keepBit;
        out[i] = mine[i];
        out[i ^ flip] = partner[i];
    }
    __syncthreads(); // every load of this CTA has returned (its value was stored)
    if (threadIdx.x == 0) {
        __threadfence();
        const unsigned int arrived = atomicAdd(ctaCounter, 1u);
        if (arrived == gridDim.x - 1) {
            *ctaCounter = 0; // for the next exchange (every CTA has arrived)
            st_release_sys(myFlags + 1, epoch);
            while (ld_acquire_sys(partnerFlags + 1) < epoch) __nanosleep(64);
        }
    }
}

// local SWAP of two physical bits (a < b), out of place; used by the NCCL path to bring `pl` to the top
__global__ void __launch_bounds__(256) swap_local_bits_kernel(const double2* __restrict__ in, double2* __restrict__ out, uint64_t n,
                                                              int a, int b) {
    const uint64_t stride = static_cast<uint64_t>(gridDim.x) * blockDim.x;
    for (uint64_t i = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
        const uint64_t ba = (i >> a) & 1ULL, bb = (i >> b) & 1ULL;
        const uint64_t j = (ba != bb) ? (i ^ ((uint64_t{1} << a) | (uint64_t{1} << b))) : i;
        out[i] = in[j];
    }
}

} // namespace fddb200
