// slab_comm.cu -- the collectives of the z-slab partition issued from inside the library: NCCL over NVLink / NVSwitch.
//
// The reference has no distributed mode; the partition follows SURVEY.md section 8e:
//   * neighbour exchange of whole planes along z for the lifting of every level (grouped ncclSend / ncclRecv),
//   * one all-reduce(MIN) of two packed int64 keys for the field extrema and for every layer's residual extrema,
//   * an all-gather of CUDA IPC handles, once per geometry, so that the symbol exchange into the global wavelet-space
//     order (slab_order.cu) can read the peers' buffers directly over NVLink.
// NCCL is loaded at run time (dlopen of libnccl.so.2: the copy already in the process -- e.g. PyTorch's -- or the
// system's), so the library itself has no link-time dependency on it and loads on machines without NCCL; a C, C++ or
// Fortran caller needs nothing but this library and an NCCL installation to run the multi-GPU mode:
//   rank 0: wrb_comm_unique_id(id)  ->  (broadcast the 128 bytes by any means)  ->  all ranks: wrb_set_comm(c, rank, n, id)
#include <dlfcn.h>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>
#include "wr_common.cuh"
#include "wr_kernels.h"
#include "slab_comm.h"

namespace wrb {

// the few NCCL declarations used (nccl.h: ncclUniqueId :38, data types :278-290, reduction ops :260-266)
struct NcclUniqueId { char internal[128]; };
typedef void* NcclComm;
enum { kNcclUint8 = 1, kNcclInt64 = 4, kNcclSum = 0, kNcclMin = 3 };

struct NcclApi {
    void* lib = nullptr;
    int (*GetUniqueId)(NcclUniqueId*) = nullptr;
    int (*CommInitRank)(NcclComm*, int, NcclUniqueId, int) = nullptr;
    int (*CommDestroy)(NcclComm) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    int (*Send)(const void*, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
    int (*Recv)(void*, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
    int (*AllGather)(const void*, void*, size_t, int, NcclComm, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    std::string err;
};

static NcclApi* nccl_api()
{
    static std::once_flag once;
    static NcclApi api;
    std::call_once(once, [] {
        const char* names[] = {"libnccl.so.2", "libnccl.so", nullptr};
        for (int i = 0; names[i] && !api.lib; i++) api.lib = dlopen(names[i], RTLD_NOW | RTLD_LOCAL);
        if (!api.lib) { api.err = std::string("cannot load NCCL: ") + dlerror(); return; }
        bool ok = true;
        auto sym = [&](const char* n) -> void* { void* p = dlsym(api.lib, n); if (!p) { ok = false; api.err = std::string("NCCL symbol missing: ") + n; } return p; };
        api.GetUniqueId = (int (*)(NcclUniqueId*))sym("ncclGetUniqueId");
        api.CommInitRank = (int (*)(NcclComm*, int, NcclUniqueId, int))sym("ncclCommInitRank");
        api.CommDestroy = (int (*)(NcclComm))sym("ncclCommDestroy");
        api.GroupStart = (int (*)())sym("ncclGroupStart");
        api.GroupEnd = (int (*)())sym("ncclGroupEnd");
        api.Send = (int (*)(const void*, size_t, int, int, NcclComm, cudaStream_t))sym("ncclSend");
        api.Recv = (int (*)(void*, size_t, int, int, NcclComm, cudaStream_t))sym("ncclRecv");
        api.AllReduce = (int (*)(const void*, void*, size_t, int, int, NcclComm, cudaStream_t))sym("ncclAllReduce");
        api.AllGather = (int (*)(const void*, void*, size_t, int, NcclComm, cudaStream_t))sym("ncclAllGather");
        api.GetErrorString = (const char* (*)(int))sym("ncclGetErrorString");
        if (!ok) { dlclose(api.lib); api.lib = nullptr; }
    });
    return &api;
}

struct SlabComm {
    NcclComm comm = nullptr;
    int rank = 0, nranks = 1;
    cudaStream_t* stream = nullptr;          // the codec's stream (read at every call: wrb_set_stream may change it)
    std::string err;
    unsigned long long halo_bytes = 0, halo_calls = 0, reduce_calls = 0;
};

static int nccl_fail(SlabComm* sc, int rc, const char* what)
{
    NcclApi* a = nccl_api();
    sc->err = std::string(what) + ": " + ((a->GetErrorString && rc > 0) ? a->GetErrorString(rc) : "error");
    return 1;
}

int slab_comm_unique_id(unsigned char id[128], std::string* err)
{
    NcclApi* a = nccl_api();
    if (!a->lib) { if (err) *err = a->err; return 1; }
    NcclUniqueId u;
    int rc = a->GetUniqueId(&u);
    if (rc != 0) { if (err) *err = std::string("ncclGetUniqueId: ") + a->GetErrorString(rc); return 1; }
    memcpy(id, u.internal, 128);
    return 0;
}

SlabComm* slab_comm_create(int rank, int nranks, const unsigned char id[128], cudaStream_t* stream, std::string* err)
{
    NcclApi* a = nccl_api();
    if (!a->lib) { if (err) *err = a->err; return nullptr; }
    SlabComm* sc = new SlabComm();
    sc->rank = rank; sc->nranks = nranks; sc->stream = stream;
    NcclUniqueId u;
    memcpy(u.internal, id, 128);
    int rc = a->CommInitRank(&sc->comm, nranks, u, rank);
    if (rc != 0) {
        if (err) *err = std::string("ncclCommInitRank: ") + a->GetErrorString(rc);
        delete sc;
        return nullptr;
    }
    return sc;
}

void slab_comm_destroy(SlabComm* sc)
{
    if (!sc) return;
    NcclApi* a = nccl_api();
    if (sc->comm && a->lib) a->CommDestroy(sc->comm);
    delete sc;
}

const char* slab_comm_error(const SlabComm* sc) { return sc ? sc->err.c_str() : ""; }
void slab_comm_counters(const SlabComm* sc, unsigned long long out[3])
{
    out[0] = sc ? sc->halo_bytes : 0; out[1] = sc ? sc->halo_calls : 0; out[2] = sc ? sc->reduce_calls : 0;
}

// HaloFn: the four transfers of one neighbour exchange in one NCCL group on the codec's stream
int slab_comm_halo(void* user, const void* send_down, const void* send_up, void* recv_lo, void* recv_hi,
                   unsigned long long down_bytes, unsigned long long up_bytes)
{
    SlabComm* sc = (SlabComm*)user;
    NcclApi* a = nccl_api();
    cudaStream_t s = *sc->stream;
    int rc = a->GroupStart();
    if (rc) return nccl_fail(sc, rc, "ncclGroupStart");
    if (sc->rank + 1 < sc->nranks) {
        if (up_bytes && (rc = a->Send(send_up, up_bytes, kNcclUint8, sc->rank + 1, sc->comm, s))) return nccl_fail(sc, rc, "ncclSend");
        if (down_bytes && (rc = a->Recv(recv_hi, down_bytes, kNcclUint8, sc->rank + 1, sc->comm, s))) return nccl_fail(sc, rc, "ncclRecv");
        sc->halo_bytes += down_bytes;
    }
    if (sc->rank > 0) {
        if (down_bytes && (rc = a->Send(send_down, down_bytes, kNcclUint8, sc->rank - 1, sc->comm, s))) return nccl_fail(sc, rc, "ncclSend");
        if (up_bytes && (rc = a->Recv(recv_lo, up_bytes, kNcclUint8, sc->rank - 1, sc->comm, s))) return nccl_fail(sc, rc, "ncclRecv");
        sc->halo_bytes += up_bytes;
    }
    rc = a->GroupEnd();
    if (rc) return nccl_fail(sc, rc, "ncclGroupEnd");
    sc->halo_calls++;
    return 0;
}

// ReduceFn: in-place all-reduce of |count| int64 values: MIN for count > 0, SUM for count < 0
int slab_comm_reduce(void* user, long long* d_buf, int count)
{
    SlabComm* sc = (SlabComm*)user;
    NcclApi* a = nccl_api();
    int rc = a->AllReduce(d_buf, d_buf, (size_t)(count < 0 ? -count : count), kNcclInt64, count < 0 ? kNcclSum : kNcclMin, sc->comm,
                          *sc->stream);
    if (rc) return nccl_fail(sc, rc, "ncclAllReduce");
    sc->reduce_calls++;
    return 0;
}

// all-gather of `bytes` per rank (device buffers; recv holds nranks * bytes)
int slab_comm_allgather(SlabComm* sc, const void* d_send, void* d_recv, size_t bytes)
{
    NcclApi* a = nccl_api();
    int rc = a->AllGather(d_send, d_recv, bytes, kNcclUint8, sc->comm, *sc->stream);
    if (rc) return nccl_fail(sc, rc, "ncclAllGather");
    return 0;
}

}  // namespace wrb
