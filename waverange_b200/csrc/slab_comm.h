// slab_comm.h -- NCCL transport of the z-slab partition (slab_comm.cu) and the global symbol order (slab_order.cu).
#pragma once
#include <cstddef>
#include <cstdint>
#include <string>
#include <cuda_runtime.h>
#include "wr_kernels.h"

namespace wrb {

struct SlabComm;
int slab_comm_unique_id(unsigned char id[128], std::string* err);
SlabComm* slab_comm_create(int rank, int nranks, const unsigned char id[128], cudaStream_t* stream, std::string* err);
void slab_comm_destroy(SlabComm* sc);
const char* slab_comm_error(const SlabComm* sc);
void slab_comm_counters(const SlabComm* sc, unsigned long long out[3]);   // halo bytes received, halo calls, reduce calls
int slab_comm_halo(void* user, const void* send_down, const void* send_up, void* recv_lo, void* recv_hi,
                   unsigned long long down_bytes, unsigned long long up_bytes);      // HaloFn, user = SlabComm*
int slab_comm_reduce(void* user, long long* d_buf, int count);                       // ReduceFn, user = SlabComm*
int slab_comm_allgather(SlabComm* sc, const void* d_send, void* d_recv, size_t bytes);

// ---- slab_order.cu: the global wavelet-space symbol order under a z-slab partition -------
constexpr int kMaxRanks = 16;
// Geometry shared by all ranks.  The global symbol sequence is the reference's array order of the GLOBAL coefficient
// array, j = x + nx*(y + ny*w) with w the wavelet-space plane (waveletcdf97_3d.c:128-135,256-263 de-interleave every
// level inside the shrinking box); rank r codes the chunks [cb[r], cb[r+1]) of it, i.e. symbols [j0[r], j0[r+1]).
struct OrderGeom {
    int nx, ny, nz, nzl, levels, nranks;
    unsigned long long chunk_len;
    unsigned long long ntot, nchunks;
    unsigned long long cb[kMaxRanks + 1];     // chunk range of every rank
    unsigned long long j0[kMaxRanks + 1];     // symbol range of every rank
};
OrderGeom make_order_geom(int nx, int ny, int nz, int nranks, int levels, unsigned long long chunk_len);
struct PeerPtrs { const uint8_t* p[kMaxRanks]; };
// encode side: my run of the global sequence, gathered from the ranks' local symbol planes (peer[r] + layer*peer_stride,
// local array order) into the coder's chunk-major padded layout, with the coder blocks' histograms
void gather_global_run(const OrderGeom& og, int rank, const PeerPtrs& peer, unsigned long long peer_stride, int nlayers,
                       const int* active, const ChunkGeom& g, uint8_t* sym, unsigned long long sym_layer_stride,
                       uint32_t* hist, unsigned long long hist_layer_stride, cudaStream_t s, cudaEvent_t after_gather = nullptr);
// decode side: my local symbol planes (array order, layer l at out + l*out_stride) gathered from the ranks' decoded
// runs (peer[r] + layer*peer_stride + (j - j0[r]))
void scatter_local_planes(const OrderGeom& og, int rank, const PeerPtrs& peer, unsigned long long peer_stride, int nlay,
                          uint8_t* out, unsigned long long out_stride, cudaStream_t s);
// host restatement of the index map (tests): global wavelet-space plane of local plane p of `rank` for an (x, y)
// position of region reg (1..levels: leaves the low box at that level; levels+1: coarsest approximation)
int order_global_plane(const OrderGeom& og, int rank, int p, int reg);

}  // namespace wrb
