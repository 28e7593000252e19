// wr_kernels.h -- internal launcher interface between codec.cu and the kernel files.
#pragma once
#include <mutex>
#include <cstdint>
#include <cuda_runtime.h>
#include "wr_common.cuh"

namespace wrb {

// ---- wavelet.cu ---------------------------------------------------------------------------
// pipe != null (and level 1 fused, nz/2 >= 64): the field arrives in four z-pieces on another stream, pipe->ev[i] marks
// piece i complete; level 1 then runs piece by piece behind the copy.  pipe->used tells the caller whether it did.
struct HostSource { cudaEvent_t ev[4]; int used; };
void wavelet_forward(const void* src, int src_is_f32, double* coef, double* tmp, double* lllA, double* lllB,
                     int nx, int ny, int nz, int levels, DevState* st, cudaStream_t s, HostSource* pipe = nullptr);
// sym != null: the detail coefficients are rebuilt from the symbol planes inside the z pass (coef is then only
// scratch); requires nz > 1 at every level, otherwise dequantise into coef first and pass sym = null
// sink != null (and every level fused): the last level runs in z-pieces and every finished piece of `out` is copied
// to sink->host on sink->copy while the next one is computed; sink->used tells the caller whether that happened
struct HostSink { void* host; cudaStream_t copy; cudaEvent_t ev[4]; int used; };
void wavelet_inverse(double* coef, double* tmp, double* lllA, double* lllB, void* out, int out_is_f32,
                     int nx, int ny, int nz, int levels, cudaStream_t s, const uint8_t* sym = nullptr,
                     unsigned long long layer_stride = 0, unsigned long long chunk_len = 0, unsigned long long pitch = 0,
                     int nlay = 0, const double* deps = nullptr, const double* minval = nullptr, HostSink* sink = nullptr,
                     double* zb = nullptr, size_t zb_bytes = 0);      // scratch of the two-pass levels (wavelet_inv2.cu)

void wavelet_xy_passes(const void* cur, int cur_is_f32, long long csy, long long csz, double* scratch, long long ay,
                       long long az, double* dst, long long dsy, long long dsz, int n0, int n1, int n2,
                       unsigned long long* in_min, unsigned long long* in_max, cudaStream_t s);
void wavelet_yx_inverse_passes(const double* src, double* scratch, long long ay, long long az, int n0, int n1, int n2,
                               void* out, int out_is_f32, long long osy, long long osz, cudaStream_t s);

// ---- wavelet_slab.cu ----------------------------------------------------------------------
// Collectives of the z-slab partition: NCCL issued from the library (slab_comm.cu, wrb_set_comm), or injected by the
// host as callbacks (wrb_set_slab: gloo in the CPU tests, an in-process emulation of several ranks on one GPU):
//   halo  : neighbour exchange along z.  Send `down_bytes` from send_down to rank-1 and receive as many into recv_hi
//           from rank+1; send `up_bytes` from send_up to rank+1 and receive as many into recv_lo from rank-1.  Nothing is
//           sent or received at the domain ends.  (Forward lifting: up = my last 4 planes, down = my first 3.)
//   reduce: in-place global MIN over `count` signed 64-bit integers (the codec packs min keys and
//           complemented max keys into one buffer so a single all-reduce(MIN) serves both); count < 0: global SUM
//           over -count values (the coded bytes of all ranks, for the container's seek-point budget)
// Both are enqueued on / ordered with the codec's stream and return 0 on success.
typedef int (*HaloFn)(void* user, const void* send_down, const void* send_up, void* recv_lo, void* recv_hi,
                      unsigned long long down_bytes, unsigned long long up_bytes);
typedef int (*ReduceFn)(void* user, long long* d_buf, int count);
struct SlabHooks { int rank = 0, nranks = 1; HaloFn halo = nullptr; ReduceFn reduce = nullptr; void* user = nullptr; };
// halo exchange of a buffer that holds `lo` halo planes, `nown` own planes and `hi` halo planes back to back
inline int halo_exchange_contiguous(const SlabHooks& hk, void* buf, size_t plane_bytes, int nown, int lo, int hi)
{
    char* b = (char*)buf;
    return hk.halo(hk.user, b + (size_t)lo * plane_bytes, b + (size_t)nown * plane_bytes, b,
                   b + (size_t)(lo + nown) * plane_bytes, (unsigned long long)hi * plane_bytes,
                   (unsigned long long)lo * plane_bytes);
}
int wavelet_slab_supported(int nx, int ny, int nz, int z0, int nzl, int levels);
// halo1: room for 7 planes of the level-1 input (nx*ny elements each, the field's own type): its z-neighbour planes
// arrive there, the slab itself is read in place
int wavelet_forward_slab(const void* src, int src_is_f32, double* coef, double* tmp, double* lllA, double* lllB, int nx,
                         int ny, int nz, int z0, int nzl, int levels, DevState* st, const SlabHooks& hk, cudaStream_t s,
                         void* halo1 = nullptr);
// sym != null (flat symbol planes of the rank-local array, layer l at sym + l*lstride): the band buffers are built
// straight from the symbols -- no dequantise pass, coef untouched; needs wavelet_inverse_slab_fused_ok()
int wavelet_inverse_slab(double* coef, double* tmp, double* lllA, double* lllB, double* ext, void* out, int out_is_f32,
                         int nx, int ny, int nz, int z0, int nzl, int levels, const SlabHooks& hk, cudaStream_t s,
                         const uint8_t* sym = nullptr, unsigned long long lstride = 0, int nlay = 0,
                         const double* deps = nullptr, const double* minval = nullptr, double* zb = nullptr, size_t zb_bytes = 0,
                         double* hb = nullptr);      // zb / hb (14 planes of nx*ny doubles): scratch of the two-pass levels
int wavelet_inverse_slab_fused_ok(int nx, int ny, int nz, int nzl, int levels);
int wavelet_inverse_slab_two_pass_ok(int nx, int ny, int nz, int nzl, int levels);   // no band buffers (`ext`) needed then

// ---- wavelet_fused.cu ---------------------------------------------------------------------
bool fused_forward_supported(int n0, int n1, int n2);
void fused_forward_level(const void* src, int src_is_f32, long long ssy, long long ssz, double* coef, long long ay,
                         long long az, double* lll, int n0, int n1, int n2, unsigned long long* in_min,
                         unsigned long long* in_max, unsigned long long* out_min, unsigned long long* out_max,
                         cudaStream_t s, int zoff = 0, int pair_lo = 0, int nl = -1, int hi_off = -1,
                         const void* halo_lo = nullptr, const void* halo_hi = nullptr);   // slab mode: halo planes apart from src

// ---- wavelet_inv_fused.cu -----------------------------------------------------------------
bool fused_inverse_supported(int n0, int n1, int n2);
// sym != null: coefficients are rebuilt from the FLAT symbol planes (layer l at sym + l*lstride); else read from coef
void fused_inverse_level(const double* coef, long long ay, long long az, const uint8_t* sym, unsigned long long lstride,
                         int nlay, const double* deps, const double* minval, const double* lll, void* dst,
                         int dst_is_f32, long long dsy, long long dsz, int n0, int n1, int n2, cudaStream_t s,
                         int seg_lo = -1, int seg_hi = -1);          // output pairs [seg_lo, seg_hi) only (default: all)

void fused_inverse_level_bands(const double* lowb, const double* highb, long long bsy, long long bsz, int halo, int pair_lo,
                               int nown, void* dst, int dst_is_f32, long long dsy, long long dsz, int n0, int n1, int n2g,
                               cudaStream_t s);

// ---- wavelet_inv2.cu ----------------------------------------------------------------------
// two-pass level (streaming z pass + TMA-staged y/x tile pass); same shapes as fused_inverse_supported()
bool inverse_two_pass_enabled();
size_t inverse_two_pass_scratch_bytes(int nx, int ny, int nz);
void inverse_level_two_pass(const double* coef, long long ay, long long az, const uint8_t* sym, unsigned long long lstride,
                            int nlay, const double* deps, const double* minval, const double* lll, void* dst,
                            int dst_is_f32, long long dsy, long long dsz, int n0, int n1, int n2, double* zb, size_t zb_bytes,
                            cudaStream_t s, int seg_lo = -1, int seg_hi = -1, const double* halo = nullptr, int n2_global = 0,
                            int own0 = 0);          // z-slab mode: see InvZArgs2 (wavelet_inv2.cu)

// ---- quant.cu -----------------------------------------------------------------------------
// Geometry of the chunked symbol container of one layer.
//   chunk c covers symbols [c*chunk_len, min(ntot,(c+1)*chunk_len)); each chunk is cut into coder
//   blocks of 60000 symbols exactly as the reference's range_encode() cuts a whole array
//   (wrappers.cpp:85-128: a chunk whose length is a multiple of 60000 ends with an EMPTY block).
//   Symbols are kept chunk-major with a padded pitch so that every chunk starts 16-byte aligned.
struct ChunkGeom {
    unsigned long long ntot;        // symbols per layer
    unsigned long long chunk_len;   // symbols per chunk (last chunk may be shorter)
    unsigned long long pitch;       // bytes between chunk starts in the symbol buffer (multiple of 16)
    unsigned int nchunks;
    unsigned int blocks_per_chunk;  // coder blocks in a full chunk (incl. a possible empty one)
    unsigned int nblocks;           // total coder blocks in the layer
    unsigned int nseek;             // seek points per chunk (decoder entry points inside a chunk); 0 = none
    unsigned int sub_len;           // symbols between seek points (multiple of 16); 0 when nseek == 0
};
// nseek_req seek points are granted only to single-block chunks (chunk_len < 60000); grids of 7 / 3 / 1 points are nested
ChunkGeom make_geom(unsigned long long ntot, unsigned long long chunk_len, unsigned int nseek_req);

void state_init(DevState* st, cudaStream_t s);
void state_prepare(DevState* st, double tolrel, cudaStream_t s);          // after field extrema are known
void layer_params(DevState* st, int layer, cudaStream_t s);                // before quantising `layer`
void quantise_layer(const double* coef, const ChunkGeom& g, int layer, DevState* st, uint8_t* sym,
                    uint32_t* hist, cudaStream_t s);
// local cutoff of encoding_wrap()'s mx*my*mz > 1 branch (see quant.cu)
struct LocalCutoff {
    int per_point;            // 1: transform off, every point gets its block's cutoff; 0: precmask = tolabs everywhere
    int nx, ny, nz, mx, my, mz;
    const double* cut;        // device, mx*my*mz
    double tolrel;            // min(cut)
};
void quantise_layer_masked(const double* coef, const ChunkGeom& g, int layer, DevState* st, uint8_t* sym, uint32_t* hist,
                           const LocalCutoff& lc, cudaStream_t s);
void dequantise(const uint8_t* sym, unsigned long long layer_stride, const ChunkGeom& g, int nlay,
                const double* deps, const double* minval, double* coef, cudaStream_t s);

// ---- rangecoder.cu ------------------------------------------------------------------------
// slot_pitch: bytes reserved per chunk in the scratch output (worst case 2 B/symbol + tables)
unsigned long long chunk_slot_pitch(const ChunkGeom& g);
void range_encode_chunks(const uint8_t* sym, unsigned long long sym_layer_stride, const uint32_t* hist,
                         unsigned long long hist_layer_stride, const ChunkGeom& g, int nlayers, const int* active,
                         uint8_t* slots, unsigned long long slot_pitch, unsigned long long* lens, uint32_t* seek,
                         cudaStream_t s);
// seek_auto: the container keeps as few of the g.nseek recorded seek points as the decoder needs and the size budget
// allows (rangecoder.cu: kSeekLaneTarget, kSeekBudget); the count goes into st->nseek_keep and the layer headers
// gtot != null (z-slab mode, global symbol order): [0] = coded bytes of ALL ranks, [1] = chunk streams of all ranks; the
// seek-point decision is then the one a single GPU makes for the whole field, and the same on every rank
void assemble_container(const uint8_t* slots, unsigned long long slot_pitch, const unsigned long long* lens,
                        const uint32_t* seek, const ChunkGeom& g, int chunked, int seek_auto, DevState* st, uint8_t* blob,
                        unsigned long long cap, unsigned long long* dst_off, cudaStream_t s,
                        const unsigned long long* gtot = nullptr);
// my coded bytes and chunk streams (active layers) -> out[0], out[1]
void sum_chunk_lens(const unsigned long long* lens, const ChunkGeom& g, const DevState* st, unsigned long long* out, cudaStream_t s);
void parse_container(const uint8_t* blob, const ChunkGeom& g, int chunked, int nlay, const unsigned long long* lay_off,
                     unsigned long long* offs, int* error, cudaStream_t s);
void range_decode_chunks(const uint8_t* blob, const unsigned long long* offs, const unsigned long long* lay_off,
                         const ChunkGeom& g, int nlay, uint8_t* sym, unsigned long long sym_layer_stride,
                         unsigned long long blob_len, int* error, cudaStream_t s);

// ---- codec.cu -----------------------------------------------------------------------------
void note_launch(int n);   // kernel-launch accounting (wrb_launch_count)

// "first use on this device?" for per-device one-time setup (cudaFuncSetAttribute is per device / context, a codec
// handle may live on any GPU, and several handles may be driven from different host threads)
struct DeviceOnce {
    std::mutex m;
    unsigned long long seen[2] = {0, 0};
    template <class F> void run(F f)
    {
        int dev = 0;
        const bool known = cudaGetDevice(&dev) == cudaSuccess && dev >= 0 && dev < 128;
        std::lock_guard<std::mutex> lock(m);                  // a second thread must not launch before the first has set it
        if (known && (seen[dev >> 6] >> (dev & 63) & 1ull)) return;
        f();
        if (known) seen[dev >> 6] |= 1ull << (dev & 63);
    }
};

}  // namespace wrb
