// rangecoder.cu -- chunk-parallel range coder.
//
// Replaces range_encode()/range_decode() (reference src/core/wrappers.cpp:68-149, :153-224) and
// the rngcod13 primitives they call (src/rangecod/rangecod.c:170-404).
//
// A range-coded stream is one serial dependency chain (32-bit low/range recurrence with an
// integer division per symbol, rangecod.c:217-229), so the parallel unit is the *chunk*: a
// contiguous run of chunk_len symbols of a layer, coded as an independent stream that is byte
// for byte what the reference's range_encode() produces for that sub-array (lead byte 00,
// per-60000-symbol block: 1-of-2 marker, 256 raw 16-bit counts, symbols against the static
// cumulative table; end marker; 5-byte flush).  One GPU thread owns one chunk (32 chunks per
// warp); per-lane frequency tables sit in shared memory laid out [symbol][lane] so the
// data-dependent lookups of a warp never bank-conflict.  All arithmetic is u32, bit-exact.
#include "wr_common.cuh"
#include "wr_kernels.h"

namespace wrb {

constexpr uint32_t kTop = 0x80000000u;      // Top_value    (rangecod.c:121)
constexpr uint32_t kBottom = 0x00800000u;   // Bottom_value (rangecod.c:129)
constexpr int kShiftBits = 23;              // SHIFT_BITS   (rangecod.c:127)
constexpr int kExtraBits = 7;               // EXTRA_BITS   (rangecod.c:128)

unsigned long long chunk_slot_pitch(const ChunkGeom& g)
{
    // worst case 2 B/symbol + 513 B per coder block + framing (wrappers.cpp:79 uses 2*BLOCKSIZE+1000)
    unsigned long long p = 2 * g.chunk_len + 1024ull * (g.blocks_per_chunk + 1);
    return (p + 15ull) & ~15ull;
}

// exact n / d for n <= 2^31 by multiply-shift: q = (n * mul) >> sh
struct Magic { uint32_t mul; uint32_t sh; };
__device__ __forceinline__ Magic make_magic(uint32_t d)
{
    uint32_t l = (d <= 1) ? 0 : 32 - __clz(d - 1);          // ceil(log2 d)
    unsigned long long p = 1ull << (31 + l);
    Magic m;
    m.mul = (uint32_t)((p + d - 1) / d);
    m.sh = 31 + l;
    return m;
}
__device__ __forceinline__ uint32_t div_magic(uint32_t n, Magic m)
{
    return (uint32_t)(((unsigned long long)n * m.mul) >> m.sh);
}

// 8-byte packing byte sink; base must be 8-byte aligned
struct Sink {
    uint8_t* base;
    unsigned long long acc;
    unsigned long long pos;
    __device__ __forceinline__ void put(uint32_t b)
    {
        acc |= (unsigned long long)(b & 0xFFu) << ((pos & 7ull) * 8);
        pos++;
        if ((pos & 7ull) == 0) { *reinterpret_cast<unsigned long long*>(base + pos - 8) = acc; acc = 0; }
    }
    __device__ __forceinline__ void flush()
    {
        unsigned long long r = pos & 7ull;
        for (unsigned long long k = 0; k < r; k++) base[pos - r + k] = (uint8_t)(acc >> (8 * k));
    }
};

struct Enc {
    uint32_t low, range, pending, count, held;
    Sink out;
};

// rangecod.c:182-207
__device__ __forceinline__ void enc_renorm(Enc& e)
{
    while (e.range <= kBottom) {
        if (e.low < (0xFFu << kShiftBits)) {
            e.out.put(e.held);
            for (; e.pending; e.pending--) e.out.put(0xFF);
            e.held = e.low >> kShiftBits;
        } else if (e.low & kTop) {
            e.out.put(e.held + 1);
            for (; e.pending; e.pending--) e.out.put(0x00);
            e.held = (e.low >> kShiftBits) & 0xFF;
        } else {
            e.pending++;
        }
        e.range <<= 8;
        e.low = (e.low << 8) & (kTop - 1);
        e.count++;
    }
}

// rangecod.c:217-229 with tot == 2 (block / end markers, wrappers.cpp:95,131)
__device__ __forceinline__ void enc_marker(Enc& e, uint32_t bit)
{
    enc_renorm(e);
    uint32_t r = e.range >> 1;
    if (bit) { e.low += r; e.range -= r; }   // sy=1, lt=1: lt+sy == tot
    else     { e.range = r; }                // sy=1, lt=0
}

// rangecod.c:231-245 as encode_short (rangecod.h:155)
__device__ __forceinline__ void enc_short(Enc& e, uint32_t v)
{
    enc_renorm(e);
    uint32_t r = e.range >> 16, t = r * v;
    e.low += t;
    if ((v + 1) >> 16) e.range -= t; else e.range = r;
}

// rangecod.c:254-276
__device__ __forceinline__ void enc_finish(Enc& e)
{
    enc_renorm(e);
    e.count += 5;
    uint32_t t = e.low >> kShiftBits;
    if (!((e.low & (kBottom - 1)) < ((e.count & 0xFFFFFFu) >> 1))) t += 1;
    if (t > 0xFF) { e.out.put(e.held + 1); for (; e.pending; e.pending--) e.out.put(0x00); }
    else          { e.out.put(e.held);     for (; e.pending; e.pending--) e.out.put(0xFF); }
    e.out.put(t);
    e.out.put(e.count >> 16);
    e.out.put(e.count >> 8);
    e.out.put(e.count);
    e.out.flush();
}

// grid (ceil(nchunks/32), layers), block 32: lane == chunk
__global__ void __launch_bounds__(32) range_encode_kernel(const uint8_t* __restrict__ sym,
                                                          unsigned long long sym_layer_stride,
                                                          const uint32_t* __restrict__ hist,
                                                          unsigned long long hist_layer_stride, ChunkGeom g,
                                                          const int* __restrict__ active, uint8_t* __restrict__ slots,
                                                          unsigned long long slot_pitch,
                                                          unsigned long long* __restrict__ lens)
{
    const int layer = blockIdx.y;
    if (active != nullptr && !active[layer]) return;
    __shared__ uint32_t tab[256 * 32];            // [symbol][lane] = cum << 16 | count
    const unsigned int lane = threadIdx.x;
    const unsigned int chunk = blockIdx.x * 32 + lane;
    if (chunk >= g.nchunks) return;
    const unsigned long long cstart = (unsigned long long)chunk * g.chunk_len;
    const unsigned long long clen = (g.ntot - cstart < g.chunk_len) ? g.ntot - cstart : g.chunk_len;
    const uint8_t* __restrict__ in = sym + (unsigned long long)layer * sym_layer_stride + (unsigned long long)chunk * g.pitch;
    const uint32_t* __restrict__ hrow = hist + (unsigned long long)layer * hist_layer_stride +
                                        (unsigned long long)chunk * g.blocks_per_chunk * 256;
    Enc e;
    e.low = 0; e.range = kTop; e.pending = 0; e.count = 0; e.held = 0;     // rangecod.c:170-176
    e.out.base = slots + ((unsigned long long)layer * g.nchunks + chunk) * slot_pitch;
    e.out.acc = 0; e.out.pos = 0;
    unsigned long long done = 0;
    for (;;) {                                                              // wrappers.cpp:85-128
        const uint32_t bs = (clen - done < kBlock) ? (uint32_t)(clen - done) : kBlock;
        enc_marker(e, 1);
        uint32_t cum = 0;
        for (int s = 0; s < 256; s++) {
            uint32_t c = hrow[s];
            enc_short(e, c);
            tab[s * 32 + lane] = (cum << 16) | c;
            cum += c;
        }
        const Magic mg = make_magic(bs);
        const uint4* __restrict__ p = reinterpret_cast<const uint4*>(in + done);
        for (uint32_t i = 0; i < bs; i += 16) {
            uint4 w = p[i >> 4];
            uint32_t ww[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
            for (int k = 0; k < 16; k++) {
                if (i + k < bs) {
                    uint32_t c = (ww[k >> 2] >> ((k & 3) * 8)) & 0xFFu;
                    uint32_t ent = tab[c * 32 + lane];
                    uint32_t sy = ent & 0xFFFFu, lt = ent >> 16;
                    enc_renorm(e);                                         // rangecod.c:217-229
                    uint32_t r = div_magic(e.range, mg), t = r * lt;
                    e.low += t;
                    if (lt + sy < bs) e.range = r * sy; else e.range -= t;
                }
            }
        }
        done += bs;
        hrow += 256;
        if (bs < kBlock) break;
    }
    enc_marker(e, 0);
    enc_finish(e);
    lens[(unsigned long long)layer * g.nchunks + chunk] = e.out.pos;
}

void range_encode_chunks(const uint8_t* sym, unsigned long long sym_layer_stride, const uint32_t* hist,
                         unsigned long long hist_layer_stride, const ChunkGeom& g, int nlayers, const int* active,
                         uint8_t* slots, unsigned long long slot_pitch, unsigned long long* lens, cudaStream_t s)
{
    dim3 grid((g.nchunks + 31) / 32, nlayers, 1);
    range_encode_kernel<<<grid, 32, 0, s>>>(sym, sym_layer_stride, hist, hist_layer_stride, g, active, slots, slot_pitch, lens);
    note_launch(1);
}

// ------------------------------------------------------------------------------------------
// container assembly: layer blob = [header 32 B][u32 len per chunk][chunk streams back to back]
// (chunked mode), or the bare stream (single-stream mode, identical to the reference's layer).
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void put_le(uint8_t* p, unsigned long long v, int nbytes)
{
    for (int k = 0; k < nbytes; k++) p[k] = (uint8_t)(v >> (8 * k));
}

__global__ void __launch_bounds__(1024) assemble_scan_kernel(const unsigned long long* __restrict__ lens, ChunkGeom g,
                                                             int chunked, DevState* st, uint8_t* __restrict__ blob,
                                                             unsigned long long cap,
                                                             unsigned long long* __restrict__ dst_off)
{
    __shared__ unsigned long long s_part[1024];
    __shared__ unsigned long long s_base;
    const int t = threadIdx.x;
    const unsigned int per = (g.nchunks + 1023) / 1024;
    const unsigned int c0 = t * per, c1 = (c0 + per < g.nchunks) ? c0 + per : g.nchunks;
    if (t == 0) s_base = 0;
    __syncthreads();
    const int nlay = st->nlay;
    for (int l = 0; l < nlay; l++) {
        const unsigned long long* ll = lens + (unsigned long long)l * g.nchunks;
        unsigned long long sum = 0;
        for (unsigned int c = c0; c < c1; c++) sum += ll[c];
        s_part[t] = sum;
        __syncthreads();
        // exclusive scan of 1024 partials (Hillis-Steele on shared memory)
        for (int o = 1; o < 1024; o <<= 1) {
            unsigned long long v = (t >= o) ? s_part[t - o] : 0;
            __syncthreads();
            s_part[t] += v;
            __syncthreads();
        }
        const unsigned long long total = s_part[1023];
        unsigned long long off = s_part[t] - sum;          // exclusive prefix of this thread's run
        const unsigned long long base = s_base;
        const unsigned long long hdr = chunked ? 32ull + 4ull * g.nchunks : 0ull;
        const bool fits = base + hdr + total <= cap;
        if (fits) {
            if (chunked && t == 0) {
                uint8_t* h = blob + base;
                h[0] = 'W'; h[1] = 'R'; h[2] = 'C'; h[3] = 'K';
                put_le(h + 4, 1, 4);
                put_le(h + 8, g.chunk_len, 8);
                put_le(h + 16, g.ntot, 8);
                put_le(h + 24, g.nchunks, 4);
                put_le(h + 28, 0, 4);
            }
            for (unsigned int c = c0; c < c1; c++) {
                dst_off[(unsigned long long)l * g.nchunks + c] = base + hdr + off;
                if (chunked) put_le(blob + base + 32 + 4ull * c, ll[c], 4);
                off += ll[c];
            }
        }
        __syncthreads();
        if (t == 0) {
            if (!fits) st->error = 1;                       // wrappers.cpp:422-426 (overflow)
            st->len_enc[l] = hdr + total;
            s_base = base + hdr + total;
        }
        __syncthreads();
    }
    if (t == 0) st->ntot_enc = s_base;
}

__global__ void __launch_bounds__(256) assemble_copy_kernel(const uint8_t* __restrict__ slots,
                                                            unsigned long long slot_pitch,
                                                            const unsigned long long* __restrict__ lens,
                                                            const unsigned long long* __restrict__ dst_off, ChunkGeom g,
                                                            const DevState* st, uint8_t* __restrict__ blob)
{
    const int l = blockIdx.y;
    if (l >= st->nlay || st->error) return;
    const unsigned long long id = (unsigned long long)l * g.nchunks + blockIdx.x;
    const uint8_t* __restrict__ src = slots + id * slot_pitch;
    uint8_t* __restrict__ dst = blob + dst_off[id];
    const unsigned long long n = lens[id];
    // align the destination, then move 16 bytes per thread with funnel-shifted aligned loads
    unsigned long long head = (16 - ((unsigned long long)dst & 15ull)) & 15ull;
    if (head > n) head = n;
    for (unsigned long long i = threadIdx.x; i < head; i += blockDim.x) dst[i] = src[i];
    const unsigned long long body = (n - head) & ~15ull;
    const uint8_t* sb = src + head;
    uint8_t* db = dst + head;
    const unsigned int sh = (unsigned int)((unsigned long long)sb & 3ull);
    const uint32_t* sw = reinterpret_cast<const uint32_t*>(sb - sh);
    for (unsigned long long i = (unsigned long long)threadIdx.x * 16; i < body; i += (unsigned long long)blockDim.x * 16) {
        const uint32_t* q = sw + (i >> 2);
        uint32_t a0 = q[0], a1 = q[1], a2 = q[2], a3 = q[3], a4 = sh ? q[4] : 0;
        uint4 o;
        o.x = __funnelshift_r(a0, a1, sh * 8);
        o.y = __funnelshift_r(a1, a2, sh * 8);
        o.z = __funnelshift_r(a2, a3, sh * 8);
        o.w = __funnelshift_r(a3, a4, sh * 8);
        *reinterpret_cast<uint4*>(db + i) = o;
    }
    for (unsigned long long i = head + body + threadIdx.x; i < n; i += blockDim.x) dst[i] = src[i];
}

void assemble_container(const uint8_t* slots, unsigned long long slot_pitch, const unsigned long long* lens,
                        const ChunkGeom& g, int chunked, DevState* st, uint8_t* blob, unsigned long long cap,
                        unsigned long long* dst_off, cudaStream_t s)
{
    assemble_scan_kernel<<<1, 1024, 0, s>>>(lens, g, chunked, st, blob, cap, dst_off);
    dim3 grid(g.nchunks, kNLayMax, 1);
    assemble_copy_kernel<<<grid, 256, 0, s>>>(slots, slot_pitch, lens, dst_off, g, st, blob);
    note_launch(2);
}

// ------------------------------------------------------------------------------------------
// decoder
// ------------------------------------------------------------------------------------------
struct Dec {
    uint32_t low, range, help, held;
    const uint8_t* in;
    unsigned long long pos;
};

// rangecod.c:293-300
__device__ __forceinline__ void dec_renorm(Dec& d)
{
    while (d.range <= kBottom) {
        d.low = (d.low << 8) | ((d.held << kExtraBits) & 0xFFu);
        d.held = d.in[d.pos++];
        d.low |= d.held >> (8 - kExtraBits);
        d.range <<= 8;
    }
}

// rangecod.c:321-331 + :362-366 (decode_short), :339-351 (decode_update)
__device__ __forceinline__ uint32_t dec_short(Dec& d)
{
    dec_renorm(d);
    d.help = d.range >> 16;
    uint32_t t = d.low / d.help;
    if (t >> 16) t = 0xFFFFu;
    uint32_t tmp = d.help * t;
    d.low -= tmp;
    if (t + 1 < (1u << 16)) d.range = d.help; else d.range -= tmp;
    return t;
}

// per-layer chunk offsets from the container tables: offs[l*nchunks + c] = byte offset of the
// chunk's stream inside the blob
__global__ void __launch_bounds__(1024) parse_container_kernel(const uint8_t* __restrict__ blob, ChunkGeom g,
                                                               int chunked, int nlay, const unsigned long long* lay_off,
                                                               unsigned long long* __restrict__ offs, int* error)
{
    __shared__ unsigned long long s_part[1024];
    const int t = threadIdx.x;
    const unsigned int per = (g.nchunks + 1023) / 1024;
    const unsigned int c0 = t * per, c1 = (c0 + per < g.nchunks) ? c0 + per : g.nchunks;
    for (int l = 0; l < nlay; l++) {
        const unsigned long long base = lay_off[l];
        if (!chunked) { if (t == 0) offs[l] = base; continue; }
        const uint8_t* tabp = blob + base + 32;
        unsigned long long sum = 0;
        for (unsigned int c = c0; c < c1; c++) {
            const uint8_t* q = tabp + 4ull * c;
            sum += (unsigned long long)q[0] | ((unsigned long long)q[1] << 8) | ((unsigned long long)q[2] << 16) |
                   ((unsigned long long)q[3] << 24);
        }
        s_part[t] = sum;
        __syncthreads();
        for (int o = 1; o < 1024; o <<= 1) {
            unsigned long long v = (t >= o) ? s_part[t - o] : 0;
            __syncthreads();
            s_part[t] += v;
            __syncthreads();
        }
        unsigned long long off = s_part[t] - sum;
        const unsigned long long total = s_part[1023];
        for (unsigned int c = c0; c < c1; c++) {
            const uint8_t* q = tabp + 4ull * c;
            unsigned long long len = (unsigned long long)q[0] | ((unsigned long long)q[1] << 8) |
                                     ((unsigned long long)q[2] << 16) | ((unsigned long long)q[3] << 24);
            offs[(unsigned long long)l * g.nchunks + c] = base + 32 + 4ull * g.nchunks + off;
            off += len;
        }
        if (t == 0 && base + 32 + 4ull * g.nchunks + total != lay_off[l + 1]) *error = 2;   // table / length mismatch
        __syncthreads();
    }
}

// grid (ceil(nchunks/32), layers), block 32: lane == chunk.   wrappers.cpp:153-224
__global__ void __launch_bounds__(32) range_decode_kernel(const uint8_t* __restrict__ blob,
                                                          const unsigned long long* __restrict__ offs, ChunkGeom g,
                                                          uint8_t* __restrict__ sym, unsigned long long sym_layer_stride,
                                                          int* error)
{
    __shared__ uint32_t cumt[257 * 32];           // [symbol][lane] exclusive cumulative counts
    const int layer = blockIdx.y;
    const unsigned int lane = threadIdx.x;
    const unsigned int chunk = blockIdx.x * 32 + lane;
    if (chunk >= g.nchunks) return;
    const unsigned long long cstart = (unsigned long long)chunk * g.chunk_len;
    const unsigned long long clen = (g.ntot - cstart < g.chunk_len) ? g.ntot - cstart : g.chunk_len;
    uint8_t* __restrict__ out = sym + (unsigned long long)layer * sym_layer_stride + (unsigned long long)chunk * g.pitch;
    Dec d;
    d.in = blob + offs[(unsigned long long)layer * g.nchunks + chunk];
    d.pos = 1;                                    // lead byte (rangecod.c:283)
    d.held = d.in[d.pos++];
    d.low = d.held >> (8 - kExtraBits);
    d.range = 1u << kExtraBits;
    d.help = 0;
    unsigned long long n = 0;
    bool bad = false;
    for (;;) {
        dec_renorm(d);                            // decode_culfreq(rc, 2)  (wrappers.cpp:174)
        d.help = d.range >> 1;
        uint32_t bit = d.low / d.help;
        if (bit >= 2) bit = 1;
        if (!bit) break;
        d.low -= d.help; d.range -= d.help;       // decode_update(rc,1,1,2)
        uint32_t acc = 0;
        for (int s = 0; s < 256; s++) {           // readcounts (rangecod.c:400-404) + prefix sums
            uint32_t c = dec_short(d);
            cumt[s * 32 + lane] = acc;
            acc += c;
        }
        cumt[256 * 32 + lane] = acc;
        const uint32_t bs = acc;
        if (n + bs > clen) { bad = true; break; }
        const Magic mg = make_magic(bs ? bs : 1);
        unsigned long long pack = 0;
        for (uint32_t i = 0; i < bs; i++) {
            dec_renorm(d);                        // decode_culfreq(rc, bs)
            d.help = div_magic(d.range, mg);
            uint32_t cf = d.low / d.help;
            if (cf >= bs) cf = bs - 1;
            uint32_t lo = 0, hi = 256;            // largest s with cum[s] <= cf  (wrappers.cpp:203-205)
#pragma unroll
            for (int it = 0; it < 8; it++) {
                uint32_t mid = (lo + hi) >> 1;
                if (cumt[mid * 32 + lane] <= cf) lo = mid; else hi = mid;
            }
            uint32_t lt = cumt[lo * 32 + lane], nx = cumt[(lo + 1) * 32 + lane];
            while (nx <= cf) { lo++; lt = nx; nx = cumt[(lo + 1) * 32 + lane]; }
            uint32_t sy = nx - lt;
            uint32_t tmp = d.help * lt;           // decode_update (rangecod.c:339-351)
            d.low -= tmp;
            if (lt + sy < bs) d.range = d.help * sy; else d.range -= tmp;
            pack |= (unsigned long long)lo << ((n & 7ull) * 8);
            n++;
            if ((n & 7ull) == 0) { *reinterpret_cast<unsigned long long*>(out + n - 8) = pack; pack = 0; }
        }
        if (n & 7ull) {                            // flush partial word (block ends are 8-aligned except the last)
            unsigned long long r = n & 7ull;
            for (unsigned long long k = 0; k < r; k++) out[n - r + k] = (uint8_t)(pack >> (8 * k));
        }
        if (bs < kBlock) {
            // the reference keeps reading markers; a well-formed chunk has its end marker next
        }
    }
    if (bad || n != clen) atomicExch(error, 3);
}

void range_decode_chunks(const uint8_t* blob, const unsigned long long* offs, const ChunkGeom& g, int nlay, uint8_t* sym,
                         unsigned long long sym_layer_stride, int* error, cudaStream_t s)
{
    dim3 grid((g.nchunks + 31) / 32, nlay, 1);
    range_decode_kernel<<<grid, 32, 0, s>>>(blob, offs, g, sym, sym_layer_stride, error);
    note_launch(1);
}

void parse_container(const uint8_t* blob, const ChunkGeom& g, int chunked, int nlay, const unsigned long long* lay_off,
                     unsigned long long* offs, int* error, cudaStream_t s)
{
    parse_container_kernel<<<1, 1024, 0, s>>>(blob, g, chunked, nlay, lay_off, offs, error);
    note_launch(1);
}

}  // namespace wrb
