// codec.cu -- codec handle, device-side pipeline orchestration and the wrb_* C ABI.
//
// Pipeline (compress), all on one stream, no host round trip until the final header read-back:
//   state_init -> forward transform (field extrema + coefficient extrema fused) -> state_prepare
//   -> for l in 0..7 { layer_params(l); quantise(l) }   (layers after the last one exit at once)
//   -> range_encode (all layers, all chunks in one launch) -> container assembly -> D2H(state)
// which restates encoding_wrap() (reference src/core/wrappers.cpp:228-452).
// Decompress mirrors decoding_wrap() (:456-527).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <algorithm>
#include <functional>
#include <atomic>
#include <new>
#include <thread>
#include <vector>
#include <algorithm>
#include "wr_common.cuh"
#include "wr_kernels.h"
#include "slab_comm.h"
#include "../../include/waverange_b200.h"

namespace wrb {
static std::atomic<unsigned long long> g_launches{0};
void note_launch(int n) { g_launches.fetch_add((unsigned long long)n, std::memory_order_relaxed); }
}  // namespace wrb

using namespace wrb;

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t n)
    {
        if (n <= cap) return cudaSuccess;
        if (p) { cudaFree(p); p = nullptr; cap = 0; }
        size_t want = n + n / 16 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) { e = cudaMalloc(&p, n); want = n; }
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

struct wrb_codec {
    int device = 0;
    cudaStream_t stream = nullptr;
    int chunk_blocks = 1;
    int seek_points = -1;            // decoder entry points inside a chunk; -1: the encoder decides (assemble_container)
    std::string err;
    DevBuf coef, tmp, lllA, lllB, sym, hist, slots, lens, dstoff, seek, state, blob, field, offs, layoff, misc, ext, lcut, zring;
    int lc_mx = 0, lc_my = 0, lc_mz = 0;      // local cutoff grid (0: off), wrb_set_local_cutoff
    double lc_min = 0;
    SlabHooks hooks;                  // z-slab partition collectives (nranks == 1: none)
    DevState* h_state = nullptr;      // pinned
    unsigned long long* h_u64 = nullptr;   // pinned scratch (64 KiB)
    int timing = 0;
    cudaEvent_t ev[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
    float stage_ms[4] = {0, 0, 0, 0};
    // pinned staging ring for pageable host buffers (wrb_encode_host / wrb_decode_host), allocated on first use
    void* stage[2] = {nullptr, nullptr};
    cudaEvent_t stage_ev[2] = {nullptr, nullptr};
    // second stream + events: wrb_decode_host copies finished z-pieces of the field out while the rest is computed
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t piece_ev[4] = {nullptr, nullptr, nullptr, nullptr};
    // z-slab partition, global order: the exchange of layer l runs on this stream while layer l+1 is quantised
    cudaStream_t xchg_stream = nullptr;
    cudaEvent_t xchg_ev[2] = {nullptr, nullptr};
    HostSource* src_pipe = nullptr;   // set by wrb_encode_host for the duration of one call: the field arrives in z-pieces
    unsigned long long guess_misses = 0;      // encodes that had to be repeated with all 8 layers (layer_guess)
    // ---- z-slab partition: transport and the global symbol order (slab_comm.cu, slab_order.cu) ----
    SlabComm* comm = nullptr;                 // NCCL transport created by wrb_set_comm (else: callbacks of wrb_set_slab)
    int order_global = 0;                     // code the GLOBAL wavelet-space symbol order (needs the peers' windows)
    DevBuf xsym, xrun, halo1, ipcbuf;         // exchange windows (local symbol planes / my decoded run), level-1 halo planes
    size_t xsym_stride = 0, xrun_stride = 0;  // bytes between layers inside the windows
    int win_layers = 0;                       // layers the windows hold
    unsigned long long win_key[4] = {0, 0, 0, 0};   // geometry the peers' pointers belong to
    bool peers_ok = false;
    PeerPtrs peer_xsym{}, peer_xrun{};
    void* ipc_open[2][kMaxRanks] = {};        // mappings of the peers' windows (cudaIpcOpenMemHandle)
    wrb_codec* local_peers[kMaxRanks] = {};   // ranks emulated in one process (wrb_set_slab_peers)
    int n_local_peers = 0;
};

#define CK(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) {                                                                   \
            c->err = std::string(#call) + ": " + cudaGetErrorString(e_);                           \
            return (e_ == cudaErrorMemoryAllocation) ? WRB_E_NOMEM : WRB_E_CUDA;                   \
        }                                                                                          \
    } while (0)

static int fail(wrb_codec* c, int code, const char* msg) { c->err = msg; return code; }

// ---- host <-> device transfers of caller-owned buffers ---------------------------------------------------
// A program written for the reference hands encoding_wrap()/decoding_wrap() ordinary (pageable) arrays.  One
// cudaMemcpy of such a buffer runs at ~10 GB/s (the driver stages it on one thread; a fresh output array is also
// first-touched page by page): 123 / 237 ms for a 512^3 double field against 24 ms from pinned memory.  Pageable
// buffers therefore go through a ring of two pinned slabs: several host threads copy slab i+1 (touching the pages in
// parallel) while the DMA engine moves slab i.  Pinned or registered buffers are copied directly.
static constexpr size_t kStageBytes = 32ull << 20;

static bool host_ptr_is_pinned(const void* p)
{
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeHost || a.type == cudaMemoryTypeManaged;
}

static void parallel_memcpy(void* dst, const void* src, size_t n)
{
    static const unsigned hw = std::thread::hardware_concurrency();
    const unsigned want = hw >= 16 ? 8u : (hw >= 4 ? hw / 2 : 1u);
    const size_t per = 2ull << 20;                               // at least 2 MiB per thread
    unsigned nt = (unsigned)std::min<size_t>(want, (n + per - 1) / per);
    if (nt <= 1) { memcpy(dst, src, n); return; }
    std::vector<std::thread> th;
    th.reserve(nt - 1);
    const size_t slice = ((n / nt) + 4095) & ~(size_t)4095;
    for (unsigned t = 1; t < nt; t++) {
        const size_t o = slice * t;
        if (o >= n) break;
        const size_t len = std::min(slice, n - o);
        th.emplace_back([=]() { memcpy((char*)dst + o, (const char*)src + o, len); });
    }
    memcpy(dst, src, std::min(slice, n));
    for (auto& x : th) x.join();
}

static int ensure_stage(wrb_codec* c)
{
    for (int i = 0; i < 2; i++) {
        if (!c->stage[i]) CK(cudaHostAlloc(&c->stage[i], kStageBytes, cudaHostAllocDefault));
        if (!c->stage_ev[i]) CK(cudaEventCreateWithFlags(&c->stage_ev[i], cudaEventDisableTiming));
    }
    return 0;
}

// host -> device on the codec's stream; returns once the last slab is enqueued (pageable source: already consumed)
static int copy_to_device(wrb_codec* c, void* d_dst, const void* h_src, size_t n)
{
    if (n == 0) return 0;
    if (n < (4ull << 20) || host_ptr_is_pinned(h_src)) {
        CK(cudaMemcpyAsync(d_dst, h_src, n, cudaMemcpyHostToDevice, c->stream));
        return 0;
    }
    int rc = ensure_stage(c);
    if (rc) return rc;
    size_t off = 0;
    for (int i = 0; off < n; i++, off += kStageBytes) {
        const int b = i & 1;
        const size_t len = std::min(kStageBytes, n - off);
        CK(cudaEventSynchronize(c->stage_ev[b]));        // slab b has left for the device (also from an earlier call)
        parallel_memcpy(c->stage[b], (const char*)h_src + off, len);
        CK(cudaMemcpyAsync((char*)d_dst + off, c->stage[b], len, cudaMemcpyHostToDevice, c->stream));
        CK(cudaEventRecord(c->stage_ev[b], c->stream));
    }
    return 0;
}

// device -> host after everything enqueued on the codec's stream; returns when the data is in h_dst
static int copy_to_host(wrb_codec* c, void* h_dst, const void* d_src, size_t n)
{
    if (n == 0) return 0;
    if (n < (4ull << 20) || host_ptr_is_pinned(h_dst)) {
        CK(cudaMemcpyAsync(h_dst, d_src, n, cudaMemcpyDeviceToHost, c->stream));
        CK(cudaStreamSynchronize(c->stream));
        return 0;
    }
    int rc = ensure_stage(c);
    if (rc) return rc;
    const size_t nslab = (n + kStageBytes - 1) / kStageBytes;
    auto issue = [&](size_t i) -> cudaError_t {
        const size_t off = i * kStageBytes, len = std::min(kStageBytes, n - off);
        cudaError_t e = cudaMemcpyAsync(c->stage[i & 1], (const char*)d_src + off, len, cudaMemcpyDeviceToHost, c->stream);
        return e != cudaSuccess ? e : cudaEventRecord(c->stage_ev[i & 1], c->stream);
    };
    CK(issue(0));
    for (size_t i = 0; i < nslab; i++) {
        if (i + 1 < nslab) CK(issue(i + 1));                                     // the other slab fills meanwhile
        CK(cudaEventSynchronize(c->stage_ev[i & 1]));
        const size_t off = i * kStageBytes, len = std::min(kStageBytes, n - off);
        parallel_memcpy((char*)h_dst + off, c->stage[i & 1], len);
    }
    return 0;
}

static unsigned long long chunk_len_of(const wrb_codec* c)
{
    return c->chunk_blocks > 0 ? (unsigned long long)c->chunk_blocks * kBlock - 1ull : 0ull;
}

static inline int half_up(int n) { return (n + 1) / 2; }

extern "C" {

int wrb_create(wrb_codec** out, int device)
{
    if (!out) return WRB_E_ARG;
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) return WRB_E_CUDA;   // no CPU fallback
    if (device < 0 || device >= ndev) return WRB_E_ARG;
    wrb_codec* c = new (std::nothrow) wrb_codec();
    if (!c) return WRB_E_NOMEM;
    c->device = device;
    if (cudaSetDevice(device) != cudaSuccess || cudaMallocHost((void**)&c->h_state, sizeof(DevState)) != cudaSuccess ||
        cudaMallocHost((void**)&c->h_u64, 65536) != cudaSuccess) {
        delete c;
        return WRB_E_CUDA;
    }
    for (int i = 0; i < 5; i++) cudaEventCreate(&c->ev[i]);
    const char* env = getenv("WRB_CHUNK_BLOCKS");
    if (env && *env) { int v = atoi(env); if (v >= 0) c->chunk_blocks = v; }
    env = getenv("WRB_SEEK_POINTS");
    if (env && *env) { int v = atoi(env); if (v >= 0 && v <= 15) c->seek_points = v; }
    *out = c;
    return 0;
}

int wrb_trim(wrb_codec* c)
{
    if (!c) return WRB_E_ARG;
    cudaSetDevice(c->device);
    DevBuf* all[] = {&c->coef, &c->tmp, &c->lllA, &c->lllB, &c->sym, &c->hist, &c->slots, &c->lens,
                     &c->dstoff, &c->seek, &c->blob, &c->field, &c->offs, &c->layoff, &c->misc, &c->ext, &c->lcut, &c->zring, &c->halo1};
    for (DevBuf* b : all) b->release();
    c->lc_mx = c->lc_my = c->lc_mz = 0;                  // the local-cutoff grid went with lcut: off until set again
    return 0;
}

static void close_peer_windows(wrb_codec* c)
{
    for (int w = 0; w < 2; w++)
        for (int r = 0; r < kMaxRanks; r++)
            if (c->ipc_open[w][r]) { cudaIpcCloseMemHandle(c->ipc_open[w][r]); c->ipc_open[w][r] = nullptr; }
    c->peers_ok = false;
}

void wrb_destroy(wrb_codec* c)
{
    if (!c) return;
    wrb_trim(c);
    cudaSetDevice(c->device);
    close_peer_windows(c);
    c->xsym.release(); c->xrun.release(); c->ipcbuf.release();
    if (c->comm) { slab_comm_destroy(c->comm); c->comm = nullptr; }
    c->state.release();
    if (c->h_state) cudaFreeHost(c->h_state);
    if (c->h_u64) cudaFreeHost(c->h_u64);
    for (int i = 0; i < 5; i++) if (c->ev[i]) cudaEventDestroy(c->ev[i]);
    for (int i = 0; i < 2; i++) {
        if (c->stage[i]) cudaFreeHost(c->stage[i]);
        if (c->stage_ev[i]) cudaEventDestroy(c->stage_ev[i]);
    }
    for (int i = 0; i < 4; i++) if (c->piece_ev[i]) cudaEventDestroy(c->piece_ev[i]);
    if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
    for (auto& e : c->xchg_ev) if (e) cudaEventDestroy(e);
    if (c->xchg_stream) cudaStreamDestroy(c->xchg_stream);
    delete c;
}

const char* wrb_last_error(const wrb_codec* c) { return c ? c->err.c_str() : "null codec"; }
int wrb_set_stream(wrb_codec* c, void* s) { if (!c) return WRB_E_ARG; c->stream = (cudaStream_t)s; return 0; }
int wrb_set_chunk_blocks(wrb_codec* c, int b) { if (!c || b < 0) return WRB_E_ARG; c->chunk_blocks = b; return 0; }
int wrb_set_local_cutoff(wrb_codec* c, int mx, int my, int mz, const double* cutoffvec)
{
    if (!c) return WRB_E_ARG;
    if (!cutoffvec || mx < 1 || my < 1 || mz < 1 || (long long)mx * my * mz <= 1) { c->lc_mx = c->lc_my = c->lc_mz = 0; return 0; }
    const size_t m = (size_t)mx * my * mz;
    CK(cudaSetDevice(c->device));
    CK(c->lcut.ensure(m * 8));
    CK(cudaMemcpyAsync(c->lcut.p, cutoffvec, m * 8, cudaMemcpyHostToDevice, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    double mn = cutoffvec[0];
    for (size_t k = 1; k < m; k++) if (cutoffvec[k] < mn) mn = cutoffvec[k];     // wrappers.cpp:292-293
    c->lc_mx = mx; c->lc_my = my; c->lc_mz = mz; c->lc_min = mn;
    return 0;
}
int wrb_set_seek_points(wrb_codec* c, int n) { if (!c || n < -1 || n > 15) return WRB_E_ARG; c->seek_points = n; return 0; }
unsigned long long wrb_launch_count(const wrb_codec*) { return g_launches.load(); }
unsigned long long wrb_layer_guess_misses(const wrb_codec* c) { return c ? c->guess_misses : 0ull; }
int wrb_current_device(int* device)
{
    if (!device) return WRB_E_ARG;
    return cudaGetDevice(device) == cudaSuccess ? 0 : WRB_E_CUDA;
}
int wrb_set_timing(wrb_codec* c, int on) { if (!c) return WRB_E_ARG; c->timing = on; return 0; }
int wrb_last_stage_ms(const wrb_codec* c, float ms[4])
{
    if (!c || !ms) return WRB_E_ARG;
    for (int i = 0; i < 4; i++) ms[i] = c->stage_ms[i];
    return 0;
}

// reference wrappers.cpp:531-541
void wrb_setup(int nx, int ny, int nz, unsigned char* nlaymax, unsigned long* ntot_enc_max)
{
    unsigned long ntot = (unsigned long)nx * (unsigned long)ny * (unsigned long)nz;
    if (nlaymax) *nlaymax = (unsigned char)kNLayMax;
    if (ntot_enc_max) *ntot_enc_max = 1ul * kNLayMax * (ntot < 1024ul ? 1024ul : ntot);
}

}  // extern "C"

// ------------------------------------------------------------------------------------------
// small utility kernels
// ------------------------------------------------------------------------------------------
template <class T>
__global__ void fill_kernel(T* out, unsigned long long n, double v)
{
    for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < n;
         i += (unsigned long long)gridDim.x * blockDim.x)
        out[i] = (T)v;
}

// chunk-major padded symbols <-> array order
__global__ void unpad_symbols_kernel(const uint8_t* __restrict__ sym, unsigned long long layer_stride, ChunkGeom g,
                                     uint8_t* __restrict__ flat)
{
    const unsigned int c = blockIdx.x;
    const int l = blockIdx.z;
    const unsigned long long cstart = (unsigned long long)c * g.chunk_len;
    const unsigned long long clen = (g.ntot - cstart < g.chunk_len) ? g.ntot - cstart : g.chunk_len;
    const uint8_t* in = sym + (unsigned long long)l * layer_stride + (unsigned long long)c * g.pitch;
    uint8_t* out = flat + (unsigned long long)l * g.ntot + cstart;
    for (unsigned long long i = blockIdx.y * (unsigned long long)blockDim.x + threadIdx.x; i < clen;
         i += (unsigned long long)gridDim.y * blockDim.x)
        out[i] = in[i];
}

// array-order symbols -> chunk-major padded layout + per-coder-block histograms (one CTA per block)
__global__ void __launch_bounds__(256) pad_hist_kernel(const uint8_t* __restrict__ flat, ChunkGeom g,
                                                       uint8_t* __restrict__ sym, uint32_t* __restrict__ hist)
{
    __shared__ uint32_t s_hist[256];
    const int tid = threadIdx.x;
    s_hist[tid] = 0;
    __syncthreads();
    const unsigned int b = blockIdx.x;
    const unsigned int c = b / g.blocks_per_chunk, kb = b % g.blocks_per_chunk;
    const unsigned long long cstart = (unsigned long long)c * g.chunk_len;
    const unsigned long long clen = (g.ntot - cstart < g.chunk_len) ? g.ntot - cstart : g.chunk_len;
    const unsigned long long boff = (unsigned long long)kb * kBlock;
    const unsigned int bs = (clen - boff < kBlock) ? (unsigned int)(clen - boff) : kBlock;
    const uint8_t* in = flat + cstart + boff;
    uint8_t* out = sym + (unsigned long long)c * g.pitch + boff;
    for (unsigned int i = tid; i < bs; i += 256) {
        uint8_t q = in[i];
        out[i] = q;
        atomicAdd(&s_hist[q], 1u);
    }
    __syncthreads();
    hist[(unsigned long long)b * 256 + tid] = s_hist[tid];
}

// extrema keys <-> one signed-comparable buffer: buf[0] = min key, buf[1] = ~max key, both with the top
// bit flipped so that signed MIN orders them like the unsigned keys; max = ~min(~max)
__global__ void pack_keys_kernel(const unsigned long long* kmin, const unsigned long long* kmax, long long* buf)
{
    buf[0] = (long long)(*kmin ^ 0x8000000000000000ull);
    buf[1] = (long long)(~*kmax ^ 0x8000000000000000ull);
}
__global__ void unpack_keys_kernel(unsigned long long* kmin, unsigned long long* kmax, const long long* buf)
{
    *kmin = (unsigned long long)buf[0] ^ 0x8000000000000000ull;
    *kmax = ~((unsigned long long)buf[1] ^ 0x8000000000000000ull);
}

__global__ void set_nlay_kernel(DevState* st, int nlay) { st->nlay = nlay; st->error = 0; }

static int grid_for(unsigned long long n)
{
    unsigned long long b = (n + 256ull * 8 - 1) / (256ull * 8);
    if (b < 1) b = 1;
    if (b > 148ull * 16) b = 148ull * 16;
    return (int)b;
}

// ------------------------------------------------------------------------------------------
// buffer sizing
// ------------------------------------------------------------------------------------------
struct SlabGeom { int nz_global, z0; };   // slab mode: the nz passed around is the LOCAL plane count

// Does the transform of an (nx, ny, nz) field go through the one-pass-per-level kernels at every level?  Then neither
// direction needs the full-size scratch `tmp`, and decoding from symbols does not need the coefficient array either
// (same predicates as wavelet_forward / wavelet_inverse in wavelet.cu).
static bool all_levels_fused(int nx, int ny, int nz, int levels)
{
    if (levels <= 0 || (long long)nx * ny >= (1ll << 31) || getenv("WRB_NO_FUSED_INVERSE") != nullptr) return false;
    int n0 = nx, n1 = ny, n2 = nz;
    for (int k = 0; k < levels; k++) {
        if (!fused_forward_supported(n0, n1, n2) || !fused_inverse_supported(n0, n1, n2)) return false;
        n0 = half_up(n0); n1 = half_up(n1); n2 = half_up(n2);
    }
    return true;
}

// z-slab mode: does every level of the forward / inverse transform run through the one-pass kernels?  (same predicates
// as wavelet_forward_slab / wavelet_inverse_slab)  Then the slab needs no full-size scratch: the level-1 halo planes have
// a buffer of their own, and the inverse builds its band buffers straight from the symbols.
static bool slab_forward_all_fused(int nx, int ny, int nz, int nzl, int levels)
{
    int n0 = nx, n1 = ny, n2g = nz, n2l = nzl;
    for (int k = 0; k < levels; k++) {
        if (!fused_forward_supported(n0, n1, n2g) || n2l < 2) return false;
        n0 = half_up(n0); n1 = half_up(n1); n2g /= 2; n2l /= 2;
    }
    return levels > 0;
}

// scratch of the transform.  need_coef / need_tmp: see all_levels_fused().  slab != null: z-slab mode (nz is the local
// plane count): halo room in the low-pass scratch, the level-1 halo planes, the band buffers of the inverse.
struct SlabGeom;
static int ensure_transform_buffers(wrb_codec* c, int nx, int ny, int nz, bool slab = false, bool need_coef = true,
                                    bool need_tmp = true, bool need_zring = false, bool need_ext = true)
{
    if (need_zring && inverse_two_pass_enabled()) CK(c->zring.ensure(inverse_two_pass_scratch_bytes(nx, ny, nz)));
    const size_t ntot = (size_t)nx * ny * nz;
    const size_t m1 = (size_t)half_up(nx) * half_up(ny) * half_up(nz);
    const size_t m2 = (size_t)half_up(half_up(nx)) * half_up(half_up(ny)) * half_up(half_up(nz));
    if (need_coef) CK(c->coef.ensure(ntot * 8));
    if (need_tmp) CK(c->tmp.ensure((ntot + (slab ? 7ull * nx * ny : 0ull)) * 8));
    if (slab && need_ext) CK(c->ext.ensure(((size_t)nz + 8) * nx * ny * 8));
    if (slab) CK(c->halo1.ensure(14ull * nx * ny * 8 + 64));     // level-1 halo planes (forward) / boundary + halo planes (inverse)
    const size_t h1 = slab ? 7ull * half_up(nx) * half_up(ny) : 0, h2 = slab ? 7ull * half_up(half_up(nx)) * half_up(half_up(ny)) : 0;
    CK(c->lllA.ensure((m1 + h1) * 8 + 64));      // slab mode: room for 4 + 3 halo planes
    CK(c->lllB.ensure((m2 + h2) * 8 + 64));
    CK(c->state.ensure(sizeof(DevState)));
    return 0;
}

static int ensure_coder_buffers(wrb_codec* c, const ChunkGeom& g, int nlayers, bool need_slots, bool need_hist = true,
                                bool need_sym = true)
{
    if (need_sym) CK(c->sym.ensure((size_t)nlayers * (((size_t)g.nchunks * g.pitch + 15) & ~(size_t)15) + 64));
    if (need_hist) CK(c->hist.ensure((size_t)nlayers * g.nblocks * 256 * 4));
    if (need_slots) {
        CK(c->slots.ensure((size_t)nlayers * g.nchunks * chunk_slot_pitch(g)));
        CK(c->lens.ensure((size_t)nlayers * g.nchunks * 8));
        CK(c->dstoff.ensure((size_t)nlayers * g.nchunks * 8));
        CK(c->seek.ensure((size_t)nlayers * g.nchunks * (g.nseek ? g.nseek : 1) * 12 + 64));      // scratch: three words per recorded point
    }
    CK(c->offs.ensure((size_t)nlayers * g.nchunks * 8 + 64));
    CK(c->layoff.ensure(16 * 8));
    CK(c->misc.ensure(256));
    CK(c->state.ensure(sizeof(DevState)));
    return 0;
}

static void header_from_state(const DevState& s, int wtflag, wrb_header* hdr)
{
    memset(hdr, 0, sizeof(*hdr));
    hdr->tolabs = s.tolabs;
    hdr->midval = s.midval;
    hdr->halfspanval = s.halfspan;
    hdr->wlev = (unsigned char)(wtflag ? kWavLvl : 0);        // wrappers.cpp:241 (set even when trivial)
    hdr->nlay = (unsigned char)s.nlay;
    hdr->ntot_enc = (unsigned long)s.ntot_enc;
    for (int l = 0; l < s.nlay && l < kNLayMax; l++) {
        hdr->deps_vec[l] = s.deps[l];
        hdr->minval_vec[l] = s.minval[l];
        hdr->len_enc_vec[l] = (unsigned long)s.len_enc[l];
    }
}

// transform + all layers; leaves coefficients in c->coef, symbols (padded) in c->sym, histograms
// in c->hist and layer parameters in the device state
// global min / max of one extrema pair across the ranks of a slab partition
static int reduce_extrema(wrb_codec* c, unsigned long long* kmin, unsigned long long* kmax)
{
    long long* buf = (long long*)((char*)c->misc.p + 64);
    pack_keys_kernel<<<1, 1, 0, c->stream>>>(kmin, kmax, buf);
    if (c->hooks.reduce(c->hooks.user, buf, 2)) return 1;
    unpack_keys_kernel<<<1, 1, 0, c->stream>>>(kmin, kmax, buf);
    note_launch(2);
    return 0;
}

// How many layers to launch.  The layer loop ends on the device (layer l is the last one when its natural step
// deps_l falls below tolabs, wrappers.cpp:326-333) and the host does not wait for that, so it used to launch all 8
// layers, 3 to 5 of them no-ops costing ~11 us each at 512^3.  The count is bounded, though: the residual of a
// layer lies within half a step, so deps_{l+1} <= deps_l / 255, and deps_0 = (coefficient span) / 255 with the
// span at most G times the field's; tolabs >= tol * span / 3.5.  Layer l therefore ends the loop as soon as
// 255^(l+1) > 3.5 G / tol.  G = 1024 is generous for four levels of the 9/7 transform but NOT proven: the caller
// checks the `done` flag that comes back with the header and repeats the call with all 8 layers if it is clear.
static int layer_guess(double tolrel)
{
    if (!(tolrel > 0)) return kNLayMax;
    const double need = 3.5 * 1024.0 / tolrel;
    double p = 255.0;
    int n = 1;
    while (p <= need && n < kNLayMax) { p *= 255.0; n++; }
    return n;
}

// sym_out / sym_stride: where the layers' symbols go (default: the codec's chunk-major buffer c->sym)
static int run_transform_and_quantise(wrb_codec* c, const void* d_field, int dtype, int nx, int ny, int nz, int wtflag,
                                      double tolrel, const ChunkGeom& g, const SlabGeom* sg = nullptr, int nlayers = kNLayMax,
                                      uint8_t* sym_out = nullptr, unsigned long long sym_stride = 0,
                                      uint32_t* hist_out = nullptr, const std::function<int(int)>* layer_complete = nullptr)
{
    DevState* st = (DevState*)c->state.p;
    cudaStream_t s = c->stream;
    state_init(st, s);
    if (c->timing) cudaEventRecord(c->ev[0], s);
    const bool dist = sg != nullptr && c->hooks.nranks > 1;
    const bool local = c->lc_mx > 0;
    if (local) {
        if (sg != nullptr) return fail(c, WRB_E_ARG, "local cutoff is not available in z-slab mode");
        tolrel = c->lc_min;
    }
    const LocalCutoff lc{wtflag ? 0 : 1, nx, ny, nz, c->lc_mx, c->lc_my, c->lc_mz, (const double*)c->lcut.p, c->lc_min};
    if (sg != nullptr && wtflag) {
        int rc = wavelet_forward_slab(d_field, dtype == WRB_F32, (double*)c->coef.p, (double*)c->tmp.p, (double*)c->lllA.p,
                                      (double*)c->lllB.p, nx, ny, sg->nz_global, sg->z0, nz, kWavLvl, st, c->hooks, s, c->halo1.p);
        if (rc) return fail(c, WRB_E_CUDA, c->comm ? slab_comm_error(c->comm) : "halo exchange callback failed");
    } else {
        wavelet_forward(d_field, dtype == WRB_F32, (double*)c->coef.p, (double*)c->tmp.p, (double*)c->lllA.p,
                        (double*)c->lllB.p, nx, ny, nz, wtflag ? kWavLvl : 0, st, s, c->src_pipe);
    }
    if (dist && reduce_extrema(c, &st->fmin_key, &st->fmax_key)) return fail(c, WRB_E_CUDA, "reduce callback failed");
    state_prepare(st, tolrel, s);
    if (c->timing) cudaEventRecord(c->ev[1], s);
    const unsigned long long lstride = sym_out ? sym_stride : (unsigned long long)g.nchunks * g.pitch;
    uint8_t* const symp = sym_out ? sym_out : (uint8_t*)c->sym.p;
    const unsigned long long hstride = (unsigned long long)g.nblocks * 256;
    uint32_t* const histp = hist_out ? hist_out : (uint32_t*)c->hist.p;
    CK(cudaMemsetAsync(histp, 0, (size_t)nlayers * hstride * 4, s));          // the quantiser adds partial histograms
    for (int l = 0; l < nlayers; l++) {
        // global extrema of the coefficients (l == 0) / of the residual left by layer l-1 (wrappers.cpp:308-314)
        if (dist && reduce_extrema(c, &st->rmin_key[l], &st->rmax_key[l])) return fail(c, WRB_E_CUDA, "reduce callback failed");
        // the all-reduce needs every rank's residual extrema of layer l-1: once it has completed on this stream, layer
        // l-1 is quantised on ALL ranks
        if (dist && l > 0 && layer_complete != nullptr) { const int rc = (*layer_complete)(l - 1); if (rc) return rc; }
        layer_params(st, l, s);
        if (local)
            quantise_layer_masked((const double*)c->coef.p, g, l, st, symp + l * lstride, histp + l * hstride, lc, s);
        else
            quantise_layer((const double*)c->coef.p, g, l, st, symp + l * lstride, histp + l * hstride, s);
    }
    if (c->timing) cudaEventRecord(c->ev[2], s);
    CK(cudaGetLastError());
    return 0;
}

extern "C" {

}  // extern "C"

// ---- z-slab partition: the global symbol order --------------------------------------------------------------------
// a collective with no payload: every rank has executed everything enqueued before it on its stream when it completes
static int slab_barrier(wrb_codec* c)
{
    long long* buf = (long long*)((char*)c->misc.p + 96);
    CK(cudaMemsetAsync(buf, 0, 16, c->stream));
    if (c->hooks.reduce(c->hooks.user, buf, 2)) return fail(c, WRB_E_CUDA, c->comm ? slab_comm_error(c->comm) : "reduce callback failed");
    return 0;
}

static bool global_order_on(const wrb_codec* c, const SlabGeom* sg)
{
    return sg != nullptr && c->order_global && c->hooks.nranks > 1 && c->chunk_blocks > 0;
}

// The exchange windows of this rank -- its local symbol planes (xsym) and its decoded run of the global sequence (xrun),
// `nlayers` layers each -- and the peers' pointers to theirs: CUDA IPC handles all-gathered over NCCL, or the codecs of
// the other emulated ranks of this process.  Collective: every rank calls it with the same geometry.
static int slab_windows(wrb_codec* c, const OrderGeom& og, int nlayers)
{
    const int R = og.nranks, rank = c->hooks.rank;
    const size_t ntl = (size_t)og.nx * og.ny * og.nzl;
    const size_t xs = ((ntl + 15) & ~(size_t)15) + 16;
    size_t maxrun = 0;                                                   // chunk-major padded: every chunk 16-byte aligned
    for (int r = 0; r < R; r++) maxrun = std::max(maxrun, (size_t)(og.cb[r + 1] - og.cb[r]) * (((size_t)og.chunk_len + 15) & ~(size_t)15));
    const size_t rs = ((maxrun + 15) & ~(size_t)15) + 64;
    const unsigned long long key[4] = {og.ntot, (unsigned long long)R | ((unsigned long long)og.levels << 8) | ((unsigned long long)og.nx << 16),
                                       og.chunk_len, (unsigned long long)og.ny};
    const bool same = c->peers_ok && memcmp(key, c->win_key, sizeof(key)) == 0 && nlayers <= c->win_layers;
    if (same) return 0;
    if (R > kMaxRanks) return fail(c, WRB_E_ARG, "too many ranks");
    CK(cudaStreamSynchronize(c->stream));
    close_peer_windows(c);
    const int nl = std::max(nlayers, c->win_layers);
    CK(c->xsym.ensure((size_t)nl * xs + 64));
    CK(c->xrun.ensure((size_t)nl * rs + 64));
    c->xsym_stride = xs; c->xrun_stride = rs; c->win_layers = nl;
    memcpy(c->win_key, key, sizeof(key));
    CK(c->misc.ensure(256));
    if (c->comm != nullptr) {
        cudaIpcMemHandle_t mine[2];
        CK(cudaIpcGetMemHandle(&mine[0], c->xsym.p));
        CK(cudaIpcGetMemHandle(&mine[1], c->xrun.p));
        const size_t hb = sizeof(mine);                                  // 128 bytes
        CK(c->ipcbuf.ensure((size_t)(R + 1) * hb));
        CK(cudaMemcpyAsync(c->ipcbuf.p, mine, hb, cudaMemcpyHostToDevice, c->stream));
        if (slab_comm_allgather(c->comm, c->ipcbuf.p, (char*)c->ipcbuf.p + hb, hb)) return fail(c, WRB_E_CUDA, slab_comm_error(c->comm));
        std::vector<cudaIpcMemHandle_t> all(2 * (size_t)R);
        CK(cudaMemcpyAsync(all.data(), (char*)c->ipcbuf.p + hb, (size_t)R * hb, cudaMemcpyDeviceToHost, c->stream));
        CK(cudaStreamSynchronize(c->stream));
        for (int r = 0; r < R; r++) {
            if (r == rank) { c->peer_xsym.p[r] = (const uint8_t*)c->xsym.p; c->peer_xrun.p[r] = (const uint8_t*)c->xrun.p; continue; }
            for (int w = 0; w < 2; w++) {
                void* p = nullptr;
                CK(cudaIpcOpenMemHandle(&p, all[2 * r + w], cudaIpcMemLazyEnablePeerAccess));
                c->ipc_open[w][r] = p;
                (w == 0 ? c->peer_xsym : c->peer_xrun).p[r] = (const uint8_t*)p;
            }
        }
    } else if (c->n_local_peers == R) {
        int rc = slab_barrier(c);                                        // every emulated rank has allocated its windows
        if (rc) return rc;
        CK(cudaStreamSynchronize(c->stream));
        for (int r = 0; r < R; r++) {
            const wrb_codec* o = c->local_peers[r];
            if (!o || !o->xsym.p || !o->xrun.p || o->xsym_stride != xs || o->xrun_stride != rs) return fail(c, WRB_E_ARG, "peer codec has no matching exchange window");
            c->peer_xsym.p[r] = (const uint8_t*)o->xsym.p;
            c->peer_xrun.p[r] = (const uint8_t*)o->xrun.p;
        }
        rc = slab_barrier(c);                                            // ... and nobody reallocates before everyone has read
        if (rc) return rc;
    } else {
        return fail(c, WRB_E_ARG, "the global symbol order needs wrb_set_comm (NCCL) or wrb_set_slab_peers (one process)");
    }
    c->peers_ok = true;
    return 0;
}

static int encode_slab_global(wrb_codec* c, const void* d_field, int dtype, int nx, int ny, int nzl, int wtflag, double tolrel,
                              wrb_header* hdr, unsigned char* d_data_enc, unsigned long cap, const SlabGeom* sg, int nlayers)
{
    const int R = c->hooks.nranks, rank = c->hooks.rank;
    if (sg->z0 != rank * nzl || (long long)nzl * R != sg->nz_global) return fail(c, WRB_E_ARG, "the global symbol order needs equal slabs in rank order");
    const OrderGeom og = make_order_geom(nx, ny, sg->nz_global, R, wtflag ? kWavLvl : 0, chunk_len_of(c));
    if (og.nchunks < (unsigned long long)R) return fail(c, WRB_E_ARG, "fewer chunks than ranks");
    const unsigned long long ntl = (unsigned long long)nx * ny * nzl;
    const unsigned long long runlen = og.j0[rank + 1] - og.j0[rank];
    const ChunkGeom gq = make_geom(ntl, 0, 0);                          // the quantiser's view: the local array, flat, one "chunk"
    const ChunkGeom gr = make_geom(runlen, og.chunk_len, c->seek_points < 0 ? 7u : (unsigned)c->seek_points);   // my run of the global sequence
    int rc;
    const bool fused_fwd = !wtflag || slab_forward_all_fused(nx, ny, sg->nz_global, nzl, kWavLvl);
    if ((rc = ensure_transform_buffers(c, nx, ny, nzl, true, true, !fused_fwd, false, false))) return rc;
    if ((rc = ensure_coder_buffers(c, gr, nlayers, true, true, false))) return rc;         // the run lives in the exchange window
    // the coder blocks' histograms, then the quantiser's (unused) histograms of its local blocks
    const unsigned long long hstride = (unsigned long long)gr.nblocks * 256;
    CK(c->hist.ensure(((size_t)nlayers * hstride + (size_t)nlayers * gq.nblocks * 256) * 4));
    uint32_t* const qhist = (uint32_t*)c->hist.p + (size_t)nlayers * hstride;
    const OrderGeom ogw = og;
    if ((rc = slab_windows(c, ogw, nlayers))) return rc;
    DevState* st = (DevState*)c->state.p;
    cudaStream_t s = c->stream;
    // (the two windows serve both directions: `xsym` holds the rank's local symbol planes -- written by the quantiser
    //  here, by the exchange when decoding -- and `xrun` its run of the global sequence -- gathered here in the coder's
    //  chunk-major layout, written by the range decoder when decoding; no further symbol buffer exists in this mode)
    const unsigned long long lstride = c->xrun_stride;
    uint8_t* const runp = (uint8_t*)c->xrun.p;
    // My run of layer l comes straight out of the peers' windows (NVLink), lands in the coder's chunk-major layout and is
    // histogrammed per coder block: one gather + one histogram launch per layer.
    auto exchange_layer = [&](int l, cudaStream_t xs, cudaEvent_t mid) {
        PeerPtrs pl = c->peer_xsym;
        for (int r = 0; r < R; r++) pl.p[r] += (unsigned long long)l * c->xsym_stride;
        gather_global_run(og, rank, pl, c->xsym_stride, 1, st->active + l, gr, runp + (unsigned long long)l * lstride, lstride,
                          (uint32_t*)c->hist.p + (unsigned long long)l * hstride, hstride, xs, mid);
    };
    static const bool dbg = getenv("WRB_DEBUG_TIMING") != nullptr;         // development: split the exchange on stderr
    // WRB_SLAB_OVERLAP=1: layers 0 .. nlayers-2 are exchanged on a second stream while the next layer is quantised (the
    // all-reduce of the residual extrema that opens layer l+1 is also the proof that layer l is complete on every rank).
    // Measured on 4 B200 (512 x 512 x 2048, profiles/r2_bench_512_n4_overlap_ab.txt): no gain -- the stage takes 1.91 ms
    // instead of 1.76, the quantiser and the gather slow each other down by what the overlap saves -- so it is off.
    const char* eo = getenv("WRB_SLAB_OVERLAP");
    const bool overlap = !dbg && (eo && *eo == '1') && nlayers > 1;
    if (overlap) {
        if (!c->xchg_stream) CK(cudaStreamCreateWithFlags(&c->xchg_stream, cudaStreamNonBlocking));
        for (auto& e : c->xchg_ev) if (!e) CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        const std::function<int(int)> layer_complete = [&](int l) -> int {
            CK(cudaEventRecord(c->xchg_ev[0], s));
            CK(cudaStreamWaitEvent(c->xchg_stream, c->xchg_ev[0], 0));
            exchange_layer(l, c->xchg_stream, nullptr);
            return 0;
        };
        if ((rc = run_transform_and_quantise(c, d_field, dtype, nx, ny, nzl, wtflag, tolrel, gq, sg, nlayers, (uint8_t*)c->xsym.p,
                                             c->xsym_stride, qhist, &layer_complete))) {
            cudaStreamSynchronize(c->xchg_stream);
            return rc;
        }
        if ((rc = slab_barrier(c))) { cudaStreamSynchronize(c->xchg_stream); return rc; }     // the last layer is complete everywhere
        exchange_layer(nlayers - 1, s, nullptr);
        CK(cudaEventRecord(c->xchg_ev[1], c->xchg_stream));
        CK(cudaStreamWaitEvent(s, c->xchg_ev[1], 0));
    } else {
        if ((rc = run_transform_and_quantise(c, d_field, dtype, nx, ny, nzl, wtflag, tolrel, gq, sg, nlayers, (uint8_t*)c->xsym.p,
                                             c->xsym_stride, qhist))) return rc;
        cudaEvent_t de[4] = {nullptr, nullptr, nullptr, nullptr};
        if (dbg) { for (auto& e : de) cudaEventCreate(&e); cudaEventRecord(de[0], s); }
        if ((rc = slab_barrier(c))) return rc;                              // every rank's symbol planes are complete
        if (dbg) cudaEventRecord(de[1], s);
        gather_global_run(og, rank, c->peer_xsym, c->xsym_stride, nlayers, st->active, gr, runp, lstride,
                          (uint32_t*)c->hist.p, hstride, s, de[2]);
        if (dbg) {
            cudaEventRecord(de[3], s);
            cudaEventSynchronize(de[3]);
            float t[4];
            for (int i = 0; i < 3; i++) cudaEventElapsedTime(&t[i], de[i], de[i + 1]);
            cudaEventElapsedTime(&t[3], c->ev[1], de[0]);
            fprintf(stderr, "[wrb rank %d] encode: quantise %.3f ms, barrier %.3f ms, gather %.3f ms, block histograms %.3f ms\n", rank, t[3], t[0], t[1], t[2]);
            for (auto& e : de) cudaEventDestroy(e);
        }
    }
    if (c->timing) cudaEventRecord(c->ev[2], s);                        // "quantise" includes the exchange in this mode
    const unsigned long long sp = chunk_slot_pitch(gr);
    range_encode_chunks(runp, lstride, (const uint32_t*)c->hist.p, hstride, gr, nlayers, st->active,
                        (uint8_t*)c->slots.p, sp, (unsigned long long*)c->lens.p, (uint32_t*)c->seek.p, s);
    if (c->timing) cudaEventRecord(c->ev[3], s);
    // the seek-point budget is decided on the coded bytes of ALL ranks (one SUM all-reduce of two values): every rank
    // keeps the number of decoder entry points a single GPU would keep for the whole field -- a rank whose run is the
    // highly compressible fine-scale band would otherwise keep none and decode eight times slower than the others
    unsigned long long* gtot = (unsigned long long*)((char*)c->misc.p + 128);
    sum_chunk_lens((const unsigned long long*)c->lens.p, gr, st, gtot, s);
    if (c->hooks.reduce(c->hooks.user, (long long*)gtot, -2)) return fail(c, WRB_E_CUDA, c->comm ? slab_comm_error(c->comm) : "reduce callback failed");
    assemble_container((const uint8_t*)c->slots.p, sp, (const unsigned long long*)c->lens.p, (const uint32_t*)c->seek.p, gr,
                       1, c->seek_points < 0, st, d_data_enc, cap, (unsigned long long*)c->dstoff.p, s, gtot);
    if (c->timing) cudaEventRecord(c->ev[4], s);
    CK(cudaMemcpyAsync(c->h_state, st, sizeof(DevState), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    CK(cudaGetLastError());
    if (c->timing) for (int i = 0; i < 4; i++) cudaEventElapsedTime(&c->stage_ms[i], c->ev[i], c->ev[i + 1]);
    if (!c->h_state->trivial && !c->h_state->done && nlayers < kNLayMax) {     // same decision on every rank (global extrema)
        c->guess_misses++;
        return encode_slab_global(c, d_field, dtype, nx, ny, nzl, wtflag, tolrel, hdr, d_data_enc, cap, sg, kNLayMax);
    }
    header_from_state(*c->h_state, wtflag, hdr);
    if (c->h_state->error) return fail(c, WRB_E_OVERFLOW, "encoded data does not fit in data_enc");
    return 0;
}

static int encode_impl(wrb_codec* c, const void* d_field, int dtype, int nx, int ny, int nz, int wtflag, double tolrel,
                       wrb_header* hdr, unsigned char* d_data_enc, unsigned long cap, const SlabGeom* sg, int nlayers = 0)
{
    if (c && c->lc_mx > 0) tolrel = c->lc_min;                    // local cutoff: the minimum sets tolabs (wrappers.cpp:292-296)
    if (nlayers <= 0) {
        const char* e = getenv("WRB_LAYER_GUESS");            // tests: force a (wrong) guess to exercise the repeat below
        nlayers = (e && *e) ? atoi(e) : layer_guess(tolrel);
        if (nlayers < 1 || nlayers > kNLayMax) nlayers = kNLayMax;
    }
    if (!c || !d_field || !hdr || !d_data_enc || nx < 1 || ny < 1 || nz < 1 || (dtype != WRB_F64 && dtype != WRB_F32))
        return c ? fail(c, WRB_E_ARG, "bad argument") : WRB_E_ARG;
    CK(cudaSetDevice(c->device));
    if (global_order_on(c, sg) && c->lc_mx == 0)
        return encode_slab_global(c, d_field, dtype, nx, ny, nz, wtflag, tolrel, hdr, d_data_enc, cap, sg, nlayers);
    const unsigned long long ntot = (unsigned long long)nx * ny * nz;
    const ChunkGeom g = make_geom(ntot, chunk_len_of(c), c->seek_points < 0 ? 7u : (unsigned)c->seek_points);
    const int chunked = c->chunk_blocks > 0;
    int rc;
    const bool fused_all = sg == nullptr && all_levels_fused(nx, ny, nz, wtflag ? kWavLvl : 0);
    const bool slab_fused_fwd = sg != nullptr && (!wtflag || slab_forward_all_fused(nx, ny, sg->nz_global, nz, kWavLvl));
    if ((rc = ensure_transform_buffers(c, nx, ny, nz, sg != nullptr, true, sg != nullptr ? !slab_fused_fwd : !fused_all, false, false))) return rc;
    if ((rc = ensure_coder_buffers(c, g, nlayers, true))) return rc;      // symbols, histograms and coder scratch of the layers launched
    DevState* st = (DevState*)c->state.p;
    cudaStream_t s = c->stream;
    if ((rc = run_transform_and_quantise(c, d_field, dtype, nx, ny, nz, wtflag, tolrel, g, sg, nlayers))) return rc;
    const unsigned long long lstride = (unsigned long long)g.nchunks * g.pitch;
    const unsigned long long hstride = (unsigned long long)g.nblocks * 256;
    const unsigned long long sp = chunk_slot_pitch(g);
    range_encode_chunks((const uint8_t*)c->sym.p, lstride, (const uint32_t*)c->hist.p, hstride, g, nlayers, st->active,
                        (uint8_t*)c->slots.p, sp, (unsigned long long*)c->lens.p, (uint32_t*)c->seek.p, s);
    if (c->timing) cudaEventRecord(c->ev[3], s);
    assemble_container((const uint8_t*)c->slots.p, sp, (const unsigned long long*)c->lens.p, (const uint32_t*)c->seek.p, g,
                       chunked, c->seek_points < 0, st, d_data_enc, cap, (unsigned long long*)c->dstoff.p, s);
    if (c->timing) cudaEventRecord(c->ev[4], s);
    CK(cudaMemcpyAsync(c->h_state, st, sizeof(DevState), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    CK(cudaGetLastError());
    if (c->timing) for (int i = 0; i < 4; i++) cudaEventElapsedTime(&c->stage_ms[i], c->ev[i], c->ev[i + 1]);
    if (!c->h_state->trivial && !c->h_state->done && nlayers < kNLayMax) {     // the layer bound did not hold: all 8
        c->guess_misses++;
        return encode_impl(c, d_field, dtype, nx, ny, nz, wtflag, tolrel, hdr, d_data_enc, cap, sg, kNLayMax);
    }
    header_from_state(*c->h_state, wtflag, hdr);
    if (c->h_state->error) return fail(c, WRB_E_OVERFLOW, "encoded data does not fit in data_enc");
    return 0;
}

extern "C" {

int wrb_encode_device(wrb_codec* c, const void* d_field, int dtype, int nx, int ny, int nz, int wtflag, double tolrel,
                      wrb_header* hdr, unsigned char* d_data_enc, unsigned long cap)
{
    return encode_impl(c, d_field, dtype, nx, ny, nz, wtflag, tolrel, hdr, d_data_enc, cap, nullptr);
}

int wrb_set_slab(wrb_codec* c, int rank, int nranks, wrb_halo_fn halo, wrb_reduce_fn reduce, void* user)
{
    if (!c || nranks < 1 || rank < 0 || rank >= nranks || (nranks > 1 && (!halo || !reduce))) return c ? fail(c, WRB_E_ARG, "bad slab setup") : WRB_E_ARG;
    c->hooks.rank = rank; c->hooks.nranks = nranks; c->hooks.halo = halo; c->hooks.reduce = reduce; c->hooks.user = user;
    return 0;
}

int wrb_comm_unique_id(unsigned char id[128])
{
    if (!id) return WRB_E_ARG;
    std::string err;
    if (slab_comm_unique_id(id, &err)) { fprintf(stderr, "waverange_b200: %s\n", err.c_str()); return WRB_E_CUDA; }
    return 0;
}

int wrb_set_comm(wrb_codec* c, int rank, int nranks, const unsigned char id[128])
{
    if (!c || !id || nranks < 1 || nranks > kMaxRanks || rank < 0 || rank >= nranks) return c ? fail(c, WRB_E_ARG, "bad communicator setup") : WRB_E_ARG;
    CK(cudaSetDevice(c->device));
    if (c->comm) { close_peer_windows(c); slab_comm_destroy(c->comm); c->comm = nullptr; }
    std::string err;
    c->comm = slab_comm_create(rank, nranks, id, &c->stream, &err);
    if (!c->comm) return fail(c, WRB_E_CUDA, err.c_str());
    c->hooks.rank = rank; c->hooks.nranks = nranks; c->hooks.halo = slab_comm_halo; c->hooks.reduce = slab_comm_reduce; c->hooks.user = c->comm;
    c->order_global = 1;
    c->n_local_peers = 0;
    c->peers_ok = false;
    return 0;
}

int wrb_set_slab_peers(wrb_codec* c, wrb_codec* const* peers, int n)
{
    if (!c || n < 0 || n > kMaxRanks || (n > 0 && !peers)) return c ? fail(c, WRB_E_ARG, "bad peer list") : WRB_E_ARG;
    for (int r = 0; r < n; r++) c->local_peers[r] = peers[r];
    c->n_local_peers = n;
    c->peers_ok = false;
    c->order_global = n > 0 ? 1 : c->order_global;
    return 0;
}

int wrb_set_slab_order(wrb_codec* c, int global)
{
    if (!c) return WRB_E_ARG;
    c->order_global = global ? 1 : 0;
    return 0;
}

int wrb_comm_counters(const wrb_codec* c, unsigned long long out[3])
{
    if (!c || !out) return WRB_E_ARG;
    slab_comm_counters(c->comm, out);
    return 0;
}

// host restatement of the partition's index map, for tests: global wavelet-space plane of local plane p of `rank` for an
// (x, y) position of region reg, and the chunk range a rank codes
int wrb_slab_order_plane(int nx, int ny, int nz, int nranks, int levels, int rank, int p, int reg)
{
    if (nranks < 1 || nranks > kMaxRanks || nz % nranks != 0) return -1;
    const OrderGeom og = make_order_geom(nx, ny, nz, nranks, levels, 59999);
    return order_global_plane(og, rank, p, reg);
}

int wrb_slab_chunk_range(int nx, int ny, int nz, int nranks, unsigned long chunk_len, int rank, unsigned long* c0, unsigned long* c1)
{
    if (nranks < 1 || nranks > kMaxRanks || rank < 0 || rank >= nranks || !c0 || !c1) return WRB_E_ARG;
    const OrderGeom og = make_order_geom(nx, ny, nz, nranks, 4, chunk_len);
    *c0 = (unsigned long)og.cb[rank]; *c1 = (unsigned long)og.cb[rank + 1];
    return 0;
}

static int slab_check(wrb_codec* c, int nx, int ny, int nz, int z0, int nzl, int levels)
{
    if (nzl < 1 || z0 < 0 || z0 + nzl > nz) return fail(c, WRB_E_ARG, "slab outside the field");
    if (!wavelet_slab_supported(nx, ny, nz, z0, nzl, levels))
        return fail(c, WRB_E_ARG, "z-slab mode needs nz % 16 == 0 and slab start/size multiples of 32");
    return 0;
}

int wrb_encode_slab_device(wrb_codec* c, const void* d_field_slab, int dtype, int nx, int ny, int nz, int z0, int nzl,
                           int wtflag, double tolrel, wrb_header* hdr, unsigned char* d_data_enc, unsigned long cap)
{
    if (!c) return WRB_E_ARG;
    int rc = slab_check(c, nx, ny, nz, z0, nzl, wtflag ? kWavLvl : 0);
    if (rc) return rc;
    SlabGeom sg{nz, z0};
    return encode_impl(c, d_field_slab, dtype, nx, ny, nzl, wtflag, tolrel, hdr, d_data_enc, cap, &sg);
}

}  // extern "C"

static int quantise_impl(wrb_codec* c, const void* d_field, int dtype, int nx, int ny, int nz, int wtflag,
                         double tolrel, wrb_header* hdr, double* d_coef, unsigned char* d_sym, const SlabGeom* sg)
{
    if (!c || !d_field || nx < 1 || ny < 1 || nz < 1 || (dtype != WRB_F64 && dtype != WRB_F32))
        return c ? fail(c, WRB_E_ARG, "bad argument") : WRB_E_ARG;
    CK(cudaSetDevice(c->device));
    const unsigned long long ntot = (unsigned long long)nx * ny * nz;
    const ChunkGeom g = make_geom(ntot, chunk_len_of(c), c->seek_points < 0 ? 7u : (unsigned)c->seek_points);
    int rc;
    if ((rc = ensure_transform_buffers(c, nx, ny, nz, sg != nullptr))) return rc;
    if ((rc = ensure_coder_buffers(c, g, kNLayMax, false))) return rc;
    DevState* st = (DevState*)c->state.p;
    cudaStream_t s = c->stream;
    if ((rc = run_transform_and_quantise(c, d_field, dtype, nx, ny, nz, wtflag, tolrel, g, sg))) return rc;
    CK(cudaMemcpyAsync(c->h_state, st, sizeof(DevState), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    if (hdr) { header_from_state(*c->h_state, wtflag, hdr); hdr->ntot_enc = 0; }
    if (d_coef) CK(cudaMemcpyAsync(d_coef, c->coef.p, ntot * 8, cudaMemcpyDeviceToDevice, s));
    if (d_sym && c->h_state->nlay > 0) {
        unsigned long long per = (g.chunk_len + 255) / 256;
        unsigned int gy = (unsigned int)(per < 64 ? (per ? per : 1) : 64);
        dim3 grid(g.nchunks, gy, c->h_state->nlay);
        unpad_symbols_kernel<<<grid, 256, 0, s>>>((const uint8_t*)c->sym.p, (unsigned long long)g.nchunks * g.pitch, g, d_sym);
        note_launch(1);
    }
    CK(cudaStreamSynchronize(s));
    CK(cudaGetLastError());
    return 0;
}

extern "C" {

int wrb_quantise_device(wrb_codec* c, const void* d_field, int dtype, int nx, int ny, int nz, int wtflag,
                        double tolrel, wrb_header* hdr, double* d_coef, unsigned char* d_sym)
{
    return quantise_impl(c, d_field, dtype, nx, ny, nz, wtflag, tolrel, hdr, d_coef, d_sym, nullptr);
}

int wrb_quantise_slab_device(wrb_codec* c, const void* d_field_slab, int dtype, int nx, int ny, int nz, int z0, int nzl,
                             int wtflag, double tolrel, wrb_header* hdr, double* d_coef, unsigned char* d_sym)
{
    if (!c) return WRB_E_ARG;
    int rc = slab_check(c, nx, ny, nz, z0, nzl, wtflag ? kWavLvl : 0);
    if (rc) return rc;
    SlabGeom sg{nz, z0};
    return quantise_impl(c, d_field_slab, dtype, nx, ny, nzl, wtflag, tolrel, hdr, d_coef, d_sym, &sg);
}

}  // extern "C"

// d_sym_flat != null: the layers' symbols are given (nlay planes of ntot bytes, array order) instead of coded data --
// the parse and range-decoder stages are skipped (wrb_decode_symbols_device, wrb_decode_slab_symbols_device)
static int decode_impl(wrb_codec* c, void* d_out, int dtype, int nx, int ny, int nz, const wrb_header* hdr,
                       const unsigned char* d_data_enc, const SlabGeom* sg, const unsigned char* d_sym_flat = nullptr,
                       HostSink* sink = nullptr)
{
    if (!c || !d_out || !hdr || nx < 1 || ny < 1 || nz < 1 || (dtype != WRB_F64 && dtype != WRB_F32))
        return c ? fail(c, WRB_E_ARG, "bad argument") : WRB_E_ARG;
    CK(cudaSetDevice(c->device));
    cudaStream_t s = c->stream;
    const unsigned long long ntot = (unsigned long long)nx * ny * nz;
    if (hdr->ntot_enc == 0 && (d_sym_flat == nullptr || hdr->nlay == 0)) {      // wrappers.cpp:462-469
        if (dtype == WRB_F32) fill_kernel<float><<<grid_for(ntot), 256, 0, s>>>((float*)d_out, ntot, hdr->midval);
        else fill_kernel<double><<<grid_for(ntot), 256, 0, s>>>((double*)d_out, ntot, hdr->midval);
        note_launch(1);
        CK(cudaStreamSynchronize(s));
        return 0;
    }
    if ((!d_data_enc && !d_sym_flat) || hdr->nlay < 1 || hdr->nlay > kNLayMax || hdr->wlev > 16) return fail(c, WRB_E_ARG, "bad header");
    const int nlay = hdr->nlay;
    if (d_sym_flat != nullptr) {
        ChunkGeom g = make_geom(ntot, 0, 0);
        g.pitch = g.chunk_len;
        int rc;
        if ((rc = ensure_transform_buffers(c, nx, ny, nz, sg != nullptr, true, true, hdr->wlev > 0))) return rc;
        const unsigned long long lstride = ntot;
        if (c->timing) for (int i = 0; i < 4; i++) cudaEventRecord(c->ev[i], s);
        const bool slab_fused = sg != nullptr && hdr->wlev > 0 && wavelet_inverse_slab_fused_ok(nx, ny, sg->nz_global, nz, (int)hdr->wlev);
        const bool fuse = slab_fused || (sg == nullptr && hdr->wlev > 0 && nz >= (1 << hdr->wlev));
        if (!fuse) dequantise(d_sym_flat, lstride, g, nlay, hdr->deps_vec, hdr->minval_vec, (double*)c->coef.p, s);
        if (sg != nullptr && hdr->wlev > 0) {
            if (wavelet_inverse_slab((double*)c->coef.p, (double*)c->tmp.p, (double*)c->lllA.p, (double*)c->lllB.p, (double*)c->ext.p,
                                     d_out, dtype == WRB_F32, nx, ny, sg->nz_global, sg->z0, nz, (int)hdr->wlev, c->hooks, s,
                                     slab_fused ? d_sym_flat : nullptr, lstride, nlay, hdr->deps_vec, hdr->minval_vec,
                                     (double*)c->zring.p, c->zring.cap, (double*)c->halo1.p))
                return fail(c, WRB_E_CUDA, "halo exchange callback failed");
        } else if (fuse) {
            wavelet_inverse((double*)c->coef.p, (double*)c->tmp.p, (double*)c->lllA.p, (double*)c->lllB.p, d_out, dtype == WRB_F32,
                            nx, ny, nz, (int)hdr->wlev, s, d_sym_flat, lstride, g.chunk_len, g.pitch, nlay, hdr->deps_vec, hdr->minval_vec,
                            nullptr, (double*)c->zring.p, c->zring.cap);
        } else {
            wavelet_inverse((double*)c->coef.p, (double*)c->tmp.p, (double*)c->lllA.p, (double*)c->lllB.p, d_out, dtype == WRB_F32,
                            nx, ny, nz, (int)hdr->wlev, s, nullptr, 0, 0, 0, 0, nullptr, nullptr, nullptr, (double*)c->zring.p, c->zring.cap);
        }
        if (c->timing) cudaEventRecord(c->ev[4], s);
        CK(cudaStreamSynchronize(s));
        CK(cudaGetLastError());
        if (c->timing) for (int i = 0; i < 4; i++) cudaEventElapsedTime(&c->stage_ms[i], c->ev[i], c->ev[i + 1]);
        return 0;
    }
    // z-slab mode in the global symbol order: the blob holds this rank's run of whole chunks of the GLOBAL sequence
    const bool global = global_order_on(c, sg);
    OrderGeom og{};
    unsigned long long nsym_expected = ntot;
    if (global) {
        const int R = c->hooks.nranks, rank = c->hooks.rank;
        if (sg->z0 != rank * nz || (long long)nz * R != sg->nz_global) return fail(c, WRB_E_ARG, "the global symbol order needs equal slabs in rank order");
        og = make_order_geom(nx, ny, sg->nz_global, R, (int)hdr->wlev, chunk_len_of(c));
        nsym_expected = og.j0[rank + 1] - og.j0[rank];
    }
    unsigned long long* lay = c->h_u64;                       // layer offsets [nlay+1]
    lay[0] = 0;
    for (int l = 0; l < nlay; l++) lay[l + 1] = lay[l] + hdr->len_enc_vec[l];
    if (lay[nlay] != hdr->ntot_enc) return fail(c, WRB_E_FORMAT, "len_enc_vec does not sum to ntot_enc");
    // container geometry from the layers' headers: every layer is checked on the host before any kernel reads its
    // tables (a truncated or corrupt file must end in WRB_E_FORMAT, not in an out-of-bounds read on the device)
    unsigned char* peek = (unsigned char*)(c->h_u64 + 64);
    if (c->timing) cudaEventRecord(c->ev[0], s);
    size_t npeek[kNLayMax];
    for (int l = 0; l < nlay; l++) {
        npeek[l] = hdr->len_enc_vec[l] < 32 ? hdr->len_enc_vec[l] : 32;
        if (npeek[l] == 0) return fail(c, WRB_E_FORMAT, "empty layer");
        CK(cudaMemcpyAsync(peek + 32 * l, d_data_enc + lay[l], npeek[l], cudaMemcpyDeviceToHost, s));
    }
    CK(cudaStreamSynchronize(s));
    int chunked = 0;
    unsigned long long chunk_len = 0, nseek = 0;
    for (int l = 0; l < nlay; l++) {
        const unsigned char* pk = peek + 32 * l;
        const bool wrck = npeek[l] >= 32 && pk[0] == 'W' && pk[1] == 'R' && pk[2] == 'C' && pk[3] == 'K';
        if (l > 0 && (int)wrck != chunked) return fail(c, WRB_E_FORMAT, "layers mix container and stream layouts");
        if (wrck) {
            chunked = 1;
            unsigned long long cl = 0, nsym = 0, nch = 0, ver = 0, nsk = 0;
            for (int k = 0; k < 8; k++) { cl |= (unsigned long long)pk[8 + k] << (8 * k); nsym |= (unsigned long long)pk[16 + k] << (8 * k); }
            for (int k = 0; k < 4; k++) { ver |= (unsigned long long)pk[4 + k] << (8 * k); nch |= (unsigned long long)pk[24 + k] << (8 * k); nsk |= (unsigned long long)pk[28 + k] << (8 * k); }
            // version 3: 10-byte seek entries on nested grids; a version-2 container without seek points has the same layout
            if (!(ver == 3 || (ver == 2 && nsk == 0)) || nsym != nsym_expected || cl == 0 || cl > nsym_expected ||
                nch != (nsym_expected + cl - 1) / cl || nsk > 15)
                return fail(c, WRB_E_FORMAT, "chunk container header does not match the field size");
            if (l > 0 && (cl != chunk_len || nsk != nseek)) return fail(c, WRB_E_FORMAT, "layers disagree on the chunk geometry");
            // header + tables + the shortest possible streams must fit into the layer: this also bounds every
            // allocation derived from the chunk count by the size of the data actually handed in
            if (32ull + (4ull + 10ull * nsk + 8ull) * nch > hdr->len_enc_vec[l]) return fail(c, WRB_E_FORMAT, "chunk tables exceed the layer");
            chunk_len = cl; nseek = nsk;
        } else if (pk[0] != 0x00) {
            return fail(c, WRB_E_FORMAT, "layer is neither a WRCK container nor a reference stream");
        }
    }
    ChunkGeom g = make_geom(nsym_expected, chunk_len, (unsigned)nseek);
    if (g.nseek != nseek) return fail(c, WRB_E_FORMAT, "seek table does not match the chunk geometry");
    if (global && (!chunked || chunk_len != (og.chunk_len < nsym_expected ? og.chunk_len : nsym_expected)))
        return fail(c, WRB_E_FORMAT, "not a run of the global chunk sequence");
    g.pitch = g.chunk_len;        // decoded symbols are kept flat (array order): the inverse transform indexes them directly
    int rc;
    const bool fused_all = sg == nullptr && all_levels_fused(nx, ny, nz, (int)hdr->wlev) && nz >= (1 << hdr->wlev) &&
                           getenv("WRB_NO_FUSED_DEQUANT") == nullptr;
    const bool slab_fused_inv = sg != nullptr && hdr->wlev > 0 && wavelet_inverse_slab_fused_ok(nx, ny, sg->nz_global, nz, (int)hdr->wlev);
    const bool need_full = sg != nullptr ? !slab_fused_inv : !fused_all;
    const bool slab_two_pass = sg != nullptr && hdr->wlev > 0 && wavelet_inverse_slab_two_pass_ok(nx, ny, sg->nz_global, nz, (int)hdr->wlev);
    if ((rc = ensure_transform_buffers(c, nx, ny, nz, sg != nullptr, need_full, need_full, (sg == nullptr && hdr->wlev > 0) || slab_two_pass,
                                       !slab_two_pass))) return rc;
    if ((rc = ensure_coder_buffers(c, g, nlay, false, false, !global))) return rc;
    ChunkGeom gl = make_geom(ntot, 0, 0);                     // the local array, flat (global mode: after the exchange)
    gl.pitch = gl.chunk_len;
    const uint8_t* symp = (const uint8_t*)c->sym.p;           // where the inverse finds the symbol planes
    if (global) {
        if ((rc = slab_windows(c, og, nlay))) return rc;
        symp = (const uint8_t*)c->xsym.p;
        if ((rc = slab_barrier(c))) return rc;                 // the peers have finished reading my run of the previous call
    }
    int* d_err = (int*)c->misc.p;
    CK(cudaMemsetAsync(d_err, 0, sizeof(int), s));
    CK(cudaMemcpyAsync(c->layoff.p, lay, (nlay + 1) * 8, cudaMemcpyHostToDevice, s));
    parse_container(d_data_enc, g, chunked, nlay, (const unsigned long long*)c->layoff.p, (unsigned long long*)c->offs.p, d_err, s);
    if (c->timing) cudaEventRecord(c->ev[1], s);
    // layer planes of the decoded symbols start 16-byte aligned (the z pass of the inverse reads them as words)
    unsigned long long lstride = ((unsigned long long)g.nchunks * g.pitch + 15ull) & ~15ull;
    if (global) {
        // my run is decoded into the exchange window; then every rank gathers its local planes from the peers' runs
        static const bool dbg = getenv("WRB_DEBUG_TIMING") != nullptr;      // development: split this stage on stderr
        cudaEvent_t de[4] = {nullptr, nullptr, nullptr, nullptr};
        if (dbg) { for (auto& e : de) cudaEventCreate(&e); cudaEventRecord(de[0], s); }
        range_decode_chunks(d_data_enc, (const unsigned long long*)c->offs.p, (const unsigned long long*)c->layoff.p, g, nlay,
                            (uint8_t*)c->xrun.p, c->xrun_stride, (unsigned long long)hdr->ntot_enc, d_err, s);
        if (dbg) cudaEventRecord(de[1], s);
        if ((rc = slab_barrier(c))) return rc;
        if (dbg) cudaEventRecord(de[2], s);
        lstride = c->xsym_stride;
        scatter_local_planes(og, c->hooks.rank, c->peer_xrun, c->xrun_stride, nlay, (uint8_t*)c->xsym.p, lstride, s);
        if (dbg) {
            cudaEventRecord(de[3], s);
            cudaEventSynchronize(de[3]);
            float t[3];
            for (int i = 0; i < 3; i++) cudaEventElapsedTime(&t[i], de[i], de[i + 1]);
            fprintf(stderr, "[wrb rank %d] decode: range_decode %.3f ms, barrier %.3f ms, scatter %.3f ms (nseek %u)\n", c->hooks.rank, t[0], t[1], t[2], g.nseek);
            for (auto& e : de) cudaEventDestroy(e);
        }
        g = gl;
    } else {
        range_decode_chunks(d_data_enc, (const unsigned long long*)c->offs.p, (const unsigned long long*)c->layoff.p, g, nlay,
                            (uint8_t*)c->sym.p, lstride, (unsigned long long)hdr->ntot_enc, d_err, s);
    }
    if (c->timing) cudaEventRecord(c->ev[2], s);
    // The inverse z pass rebuilds the coefficients from the symbols itself; a separate dequantise pass is only
    // needed without a transform, for extent-1 z, and in slab mode (its band buffers are built from coef).
    const bool slab_fused = sg != nullptr && hdr->wlev > 0 && wavelet_inverse_slab_fused_ok(nx, ny, sg->nz_global, nz, (int)hdr->wlev);
    const bool fuse_deq = slab_fused || (sg == nullptr && hdr->wlev > 0 && nz >= (1 << hdr->wlev) && getenv("WRB_NO_FUSED_DEQUANT") == nullptr);
    if (!fuse_deq) dequantise(symp, lstride, g, nlay, hdr->deps_vec, hdr->minval_vec, (double*)c->coef.p, s);
    if (c->timing) cudaEventRecord(c->ev[3], s);
    if (sg != nullptr && hdr->wlev > 0) {
        if (wavelet_inverse_slab((double*)c->coef.p, (double*)c->tmp.p, (double*)c->lllA.p, (double*)c->lllB.p, (double*)c->ext.p,
                                 d_out, dtype == WRB_F32, nx, ny, sg->nz_global, sg->z0, nz, (int)hdr->wlev, c->hooks, s,
                                 slab_fused ? symp : nullptr, lstride, nlay, hdr->deps_vec, hdr->minval_vec,
                                 (double*)c->zring.p, c->zring.cap, (double*)c->halo1.p))
            return fail(c, WRB_E_CUDA, c->comm ? slab_comm_error(c->comm) : "halo exchange callback failed");
    } else {
        if (fuse_deq)
            wavelet_inverse((double*)c->coef.p, (double*)c->tmp.p, (double*)c->lllA.p, (double*)c->lllB.p, d_out, dtype == WRB_F32,
                            nx, ny, nz, (int)hdr->wlev, s, (const uint8_t*)c->sym.p, lstride, g.chunk_len, g.pitch, nlay,
                            hdr->deps_vec, hdr->minval_vec, sink, (double*)c->zring.p, c->zring.cap);
        else
            wavelet_inverse((double*)c->coef.p, (double*)c->tmp.p, (double*)c->lllA.p, (double*)c->lllB.p, d_out, dtype == WRB_F32,
                            nx, ny, nz, (int)hdr->wlev, s, nullptr, 0, 0, 0, 0, nullptr, nullptr, nullptr, (double*)c->zring.p, c->zring.cap);
    }
    if (c->timing) cudaEventRecord(c->ev[4], s);
    int* h_err = (int*)(c->h_u64 + 128);
    CK(cudaMemcpyAsync(h_err, d_err, sizeof(int), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    CK(cudaGetLastError());
    if (c->timing) for (int i = 0; i < 4; i++) cudaEventElapsedTime(&c->stage_ms[i], c->ev[i], c->ev[i + 1]);
    if (*h_err) return fail(c, WRB_E_FORMAT, "range decoder: malformed chunk stream");
    return 0;
}

extern "C" {

int wrb_decode_device(wrb_codec* c, void* d_out, int dtype, int nx, int ny, int nz, const wrb_header* hdr,
                      const unsigned char* d_data_enc)
{
    return decode_impl(c, d_out, dtype, nx, ny, nz, hdr, d_data_enc, nullptr);
}

int wrb_decode_slab_device(wrb_codec* c, void* d_out_slab, int dtype, int nx, int ny, int nz, int z0, int nzl,
                           const wrb_header* hdr, const unsigned char* d_data_enc)
{
    if (!c || !hdr) return WRB_E_ARG;
    int rc = slab_check(c, nx, ny, nz, z0, nzl, (int)hdr->wlev);
    if (rc) return rc;
    SlabGeom sg{nz, z0};
    return decode_impl(c, d_out_slab, dtype, nx, ny, nzl, hdr, d_data_enc, &sg);
}

int wrb_decode_symbols_device(wrb_codec* c, void* d_out, int dtype, int nx, int ny, int nz, const wrb_header* hdr,
                              const unsigned char* d_sym)
{
    if (!d_sym) return c ? fail(c, WRB_E_ARG, "bad argument") : WRB_E_ARG;
    return decode_impl(c, d_out, dtype, nx, ny, nz, hdr, nullptr, nullptr, d_sym);
}

int wrb_decode_slab_symbols_device(wrb_codec* c, void* d_out_slab, int dtype, int nx, int ny, int nz, int z0, int nzl,
                                   const wrb_header* hdr, const unsigned char* d_sym)
{
    if (!c || !hdr || !d_sym) return c ? fail(c, WRB_E_ARG, "bad argument") : WRB_E_ARG;
    int rc = slab_check(c, nx, ny, nz, z0, nzl, (int)hdr->wlev);
    if (rc) return rc;
    SlabGeom sg{nz, z0};
    return decode_impl(c, d_out_slab, dtype, nx, ny, nzl, hdr, nullptr, &sg, d_sym);
}

int wrb_encode_host(wrb_codec* c, const void* field, int dtype, int nx, int ny, int nz, int wtflag, double tolrel,
                    wrb_header* hdr, unsigned char* data_enc, unsigned long cap)
{
    if (!c || !field || !hdr || !data_enc || nx < 1 || ny < 1 || nz < 1 || (dtype != WRB_F64 && dtype != WRB_F32))
        return c ? fail(c, WRB_E_ARG, "bad argument") : WRB_E_ARG;
    CK(cudaSetDevice(c->device));
    const size_t ntot = (size_t)nx * ny * nz;
    const size_t esz = dtype == WRB_F32 ? 4 : 8;
    CK(c->field.ensure(ntot * esz));
    CK(c->blob.ensure((size_t)cap + 64));
    int rc;
    // pinned source and a level-1 box the one-pass kernel takes: the field is copied in four z-pieces on a second stream
    // and level 1 of the transform follows piece by piece (only the last piece's share of it stays exposed)
    if (wtflag && nz / 2 >= 64 && (nz & 1) == 0 && ntot * esz >= (32ull << 20) && fused_forward_supported(nx, ny, nz) &&
        host_ptr_is_pinned(field)) {
        if (!c->copy_stream) CK(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
        for (int i = 0; i < 4; i++) if (!c->piece_ev[i]) CK(cudaEventCreateWithFlags(&c->piece_ev[i], cudaEventDisableTiming));
        HostSource pipe{};
        const size_t plane = (size_t)nx * ny * esz;
        const int m2 = nz / 2;
        for (int i = 0; i < 4; i++) {
            const size_t z0 = 2 * (size_t)((long long)m2 * i / 4), z1 = 2 * (size_t)((long long)m2 * (i + 1) / 4);
            CK(cudaMemcpyAsync((char*)c->field.p + z0 * plane, (const char*)field + z0 * plane, (z1 - z0) * plane,
                               cudaMemcpyHostToDevice, c->copy_stream));
            CK(cudaEventRecord(c->piece_ev[i], c->copy_stream));
            pipe.ev[i] = c->piece_ev[i];
        }
        c->src_pipe = &pipe;
        rc = wrb_encode_device(c, c->field.p, dtype, nx, ny, nz, wtflag, tolrel, hdr, (unsigned char*)c->blob.p, cap);
        c->src_pipe = nullptr;
        // on failure (or when the transform could not follow the pieces) nothing of ours may still be reading the
        // caller's buffer when this call returns
        if (rc || !pipe.used) cudaStreamSynchronize(c->copy_stream);
    } else {
        rc = copy_to_device(c, c->field.p, field, ntot * esz);
        if (rc) return rc;
        rc = wrb_encode_device(c, c->field.p, dtype, nx, ny, nz, wtflag, tolrel, hdr, (unsigned char*)c->blob.p, cap);
    }
    if (rc) return rc;
    return copy_to_host(c, data_enc, c->blob.p, hdr->ntot_enc);
}

int wrb_decode_host(wrb_codec* c, void* field_out, int dtype, int nx, int ny, int nz, const wrb_header* hdr,
                    const unsigned char* data_enc)
{
    if (!c || !field_out || !hdr || nx < 1 || ny < 1 || nz < 1 || (dtype != WRB_F64 && dtype != WRB_F32))
        return c ? fail(c, WRB_E_ARG, "bad argument") : WRB_E_ARG;
    CK(cudaSetDevice(c->device));
    const size_t ntot = (size_t)nx * ny * nz;
    const size_t esz = dtype == WRB_F32 ? 4 : 8;
    CK(c->field.ensure(ntot * esz));
    CK(c->blob.ensure((size_t)hdr->ntot_enc + 64));
    if (hdr->ntot_enc) {
        if (!data_enc) return fail(c, WRB_E_ARG, "data_enc is null");
        const int rcc = copy_to_device(c, c->blob.p, data_enc, hdr->ntot_enc);
        if (rcc) return rcc;
        CK(cudaMemsetAsync((unsigned char*)c->blob.p + hdr->ntot_enc, 0, 64, c->stream));
    }
    // pinned destination: the last inverse level hands its z-pieces to a copy stream as they finish
    HostSink sink{};
    HostSink* use = nullptr;
    if (ntot * esz >= (32ull << 20) && host_ptr_is_pinned(field_out)) {
        if (!c->copy_stream) CK(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
        for (int i = 0; i < 4; i++) if (!c->piece_ev[i]) CK(cudaEventCreateWithFlags(&c->piece_ev[i], cudaEventDisableTiming));
        sink.host = field_out; sink.copy = c->copy_stream; sink.used = 0;
        for (int i = 0; i < 4; i++) sink.ev[i] = c->piece_ev[i];
        use = &sink;
    }
    int rc = decode_impl(c, c->field.p, dtype, nx, ny, nz, hdr, (const unsigned char*)c->blob.p, nullptr, nullptr, use);
    if (use && sink.used) {                     // also on failure: nothing of ours may still be writing the caller's buffer
        cudaError_t e = cudaStreamSynchronize(c->copy_stream);
        if (rc) return rc;
        CK(e);
        return 0;
    }
    if (rc) return rc;
    return copy_to_host(c, field_out, c->field.p, ntot * esz);
}

int wrb_wavelet3d_device(wrb_codec* c, double* d_x, int nx, int ny, int nz, int lvl)
{
    if (!c || !d_x || nx < 1 || ny < 1 || nz < 1) return c ? fail(c, WRB_E_ARG, "bad argument") : WRB_E_ARG;
    CK(cudaSetDevice(c->device));
    const size_t ntot = (size_t)nx * ny * nz;
    int rc;
    if ((rc = ensure_transform_buffers(c, nx, ny, nz, false, true, true, lvl < 0))) return rc;
    CK(c->field.ensure(ntot * 8));
    cudaStream_t s = c->stream;
    if (lvl == 0) return 0;
    if (lvl > 0) {
        DevState* st = (DevState*)c->state.p;
        state_init(st, s);
        CK(cudaMemcpyAsync(c->field.p, d_x, ntot * 8, cudaMemcpyDeviceToDevice, s));
        wavelet_forward(c->field.p, 0, d_x, (double*)c->tmp.p, (double*)c->lllA.p, (double*)c->lllB.p, nx, ny, nz, lvl, st, s);
    } else {
        wavelet_inverse(d_x, (double*)c->tmp.p, (double*)c->lllA.p, (double*)c->lllB.p, c->field.p, 0, nx, ny, nz, -lvl, s,
                        nullptr, 0, 0, 0, 0, nullptr, nullptr, nullptr, (double*)c->zring.p, c->zring.cap);
        CK(cudaMemcpyAsync(d_x, c->field.p, ntot * 8, cudaMemcpyDeviceToDevice, s));
    }
    CK(cudaStreamSynchronize(s));
    CK(cudaGetLastError());
    return 0;
}

int wrb_range_encode_device(wrb_codec* c, const unsigned char* d_sym, unsigned long n, unsigned long chunk_len,
                            unsigned char* d_out, unsigned long cap, unsigned long* lens, unsigned long* total)
{
    if (!c || !d_sym || !d_out || n < 1) return c ? fail(c, WRB_E_ARG, "bad argument") : WRB_E_ARG;
    CK(cudaSetDevice(c->device));
    const ChunkGeom g = make_geom(n, chunk_len, 0);
    int rc;
    if ((rc = ensure_coder_buffers(c, g, 1, true))) return rc;
    cudaStream_t s = c->stream;
    DevState* st = (DevState*)c->state.p;
    state_init(st, s);
    set_nlay_kernel<<<1, 1, 0, s>>>(st, 1);
    pad_hist_kernel<<<g.nblocks, 256, 0, s>>>(d_sym, g, (uint8_t*)c->sym.p, (uint32_t*)c->hist.p);
    note_launch(2);
    const unsigned long long sp = chunk_slot_pitch(g);
    range_encode_chunks((const uint8_t*)c->sym.p, 0, (const uint32_t*)c->hist.p, 0, g, 1, nullptr, (uint8_t*)c->slots.p, sp,
                        (unsigned long long*)c->lens.p, (uint32_t*)c->seek.p, s);
    assemble_container((const uint8_t*)c->slots.p, sp, (const unsigned long long*)c->lens.p, (const uint32_t*)c->seek.p, g, 0, 0,
                       st, d_out, cap, (unsigned long long*)c->dstoff.p, s);
    CK(cudaMemcpyAsync(c->h_state, st, sizeof(DevState), cudaMemcpyDeviceToHost, s));
    if (lens) {
        if ((size_t)g.nchunks * 8 <= 65536 - 2048) {
            CK(cudaMemcpyAsync(c->h_u64 + 256, c->lens.p, (size_t)g.nchunks * 8, cudaMemcpyDeviceToHost, s));
            CK(cudaStreamSynchronize(s));
            for (unsigned int i = 0; i < g.nchunks; i++) lens[i] = (unsigned long)c->h_u64[256 + i];
        } else {
            CK(cudaStreamSynchronize(s));
            CK(cudaMemcpy(lens, c->lens.p, (size_t)g.nchunks * 8, cudaMemcpyDeviceToHost));
        }
    }
    CK(cudaStreamSynchronize(s));
    CK(cudaGetLastError());
    if (total) *total = (unsigned long)c->h_state->ntot_enc;
    if (c->h_state->error) return fail(c, WRB_E_OVERFLOW, "encoded data does not fit");
    return 0;
}

int wrb_range_decode_device(wrb_codec* c, const unsigned char* d_in, const unsigned long* lens, unsigned long n,
                            unsigned long chunk_len, unsigned char* d_sym)
{
    if (!c || !d_in || !lens || !d_sym || n < 1) return c ? fail(c, WRB_E_ARG, "bad argument") : WRB_E_ARG;
    CK(cudaSetDevice(c->device));
    const ChunkGeom g = make_geom(n, chunk_len, 0);
    int rc;
    if ((rc = ensure_coder_buffers(c, g, 1, false))) return rc;
    cudaStream_t s = c->stream;
    unsigned long long* offs = (unsigned long long*)malloc((size_t)g.nchunks * 8);
    if (!offs) return fail(c, WRB_E_NOMEM, "host allocation failed");
    unsigned long long acc = 0;
    for (unsigned int i = 0; i < g.nchunks; i++) { offs[i] = acc; acc += lens[i]; }
    int* d_err = (int*)c->misc.p;
    cudaError_t e1 = cudaMemsetAsync(d_err, 0, sizeof(int), s);
    cudaError_t e2 = cudaMemcpyAsync(c->offs.p, offs, (size_t)g.nchunks * 8, cudaMemcpyHostToDevice, s);
    cudaError_t e3 = cudaStreamSynchronize(s);
    free(offs);
    CK(e1); CK(e2); CK(e3);
    range_decode_chunks(d_in, (const unsigned long long*)c->offs.p, nullptr, g, 1, (uint8_t*)c->sym.p, 0, acc, d_err, s);
    unsigned long long per = (g.chunk_len + 255) / 256;
    unsigned int gy = (unsigned int)(per < 64 ? (per ? per : 1) : 64);
    dim3 grid(g.nchunks, gy, 1);
    unpad_symbols_kernel<<<grid, 256, 0, s>>>((const uint8_t*)c->sym.p, 0, g, d_sym);
    note_launch(1);
    int* h_err = (int*)(c->h_u64 + 128);
    CK(cudaMemcpyAsync(h_err, d_err, sizeof(int), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    CK(cudaGetLastError());
    if (*h_err) return fail(c, WRB_E_FORMAT, "range decoder: malformed chunk stream");
    return 0;
}

// reference src/waveletcdf97_3d/waveletcdf97_3d.c:473-553.  The index is updated one direction
// at a time and re-tested against the current box before each direction, as the reference does.
void wrb_ind_p2w_3d(int lvlin, int n1, int n2, int n3, int i1, int i2, int i3, int* lvl, int* o1, int* o2, int* o3)
{
    int n[3] = {n1, n2, n3}, idx[3] = {i1, i2, i3};
    int level = 0, moved = 0;
    for (int k = 1; k <= lvlin; k++) {
        int m[3] = {half_up(n[0]), half_up(n[1]), half_up(n[2])};
        for (int d = 0; d < 3; d++) {
            if (n[d] <= 1) continue;
            if (idx[0] < n[0] && idx[1] < n[1] && idx[2] < n[2]) {
                idx[d] = (idx[d] & 1) ? idx[d] / 2 + m[d] : idx[d] / 2;
                moved = 1;
            }
        }
        for (int d = 0; d < 3; d++) n[d] = m[d];
        if (moved) level++;
    }
    if (lvl) *lvl = level;
    if (o1) *o1 = idx[0];
    if (o2) *o2 = idx[1];
    if (o3) *o3 = idx[2];
}

}  // extern "C"
