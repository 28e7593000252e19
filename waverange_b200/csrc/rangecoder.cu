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
// cumulative table; end marker; 5-byte flush).  One GPU thread owns one chunk; per-lane
// frequency tables sit in shared memory laid out [symbol][lane] so the data-dependent lookups of
// a warp never bank-conflict.  All coder arithmetic is u32, bit-exact.
//
// Three things keep the serial chain short (a lone warp pays ~4 cycles per dependent instruction):
//  * k-step renormalisation: for tot <= 60000 the quotient r = range/tot is >= 139, so at most two
//    byte shifts are pending before a symbol; k is computed directly and applied with one shift.
//  * deferred carries: instead of the reference's buffer/bytes_to_follow bookkeeping
//    (rangecod.c:182-207) the encoder stores the raw 9-bit value low >> 23 of every shift; the
//    compaction kernel then forms byte j as B_j + carry(first later entry != 0xFF), which is what
//    that bookkeeping produces, in parallel over all bytes.
//  * seek points: the encoder records (low, range, position) at a few interior symbol indices of a
//    chunk.  From them and the final stream bytes W at that position a decoder state follows as
//    X = W - 2*low (mod 2^32), X being the decoder's (low << 1 | last bit read), so several lanes
//    decode one chunk concurrently.  The chunk stream itself is unchanged.
#include <cstdlib>
#include <type_traits>
#include "wr_common.cuh"
#include "wr_kernels.h"

namespace wrb {

constexpr uint32_t kTop = 0x80000000u;      // Top_value    (rangecod.c:121)
constexpr uint32_t kBottom = 0x00800000u;   // Bottom_value (rangecod.c:129)
constexpr int kShiftBits = 23;              // SHIFT_BITS   (rangecod.c:127)
constexpr uint32_t kTerm = 0x8000u;         // raw-entry flag: written by done_encoding, stops carry scans

unsigned long long chunk_slot_pitch(const ChunkGeom& g)
{
    // Raw 16-bit entries, one per output byte.  Upper bound of a block's output: a symbol of count c in a block of bs
    // symbols costs log2(range / (r * c)) bits with r = floor(range / bs) >= 2^23 / 60000 > 139, i.e. at most
    // log2(bs / c) + log2(1 + 1/139) < log2(bs / c) + 0.0104 bits, and sum log2(bs / c) <= 8 bs for 256 symbols: a block
    // is at most bs * 8.0104 / 8 bytes + its 512-byte count table + markers and flush.  (The reference sizes its buffer
    // 2 * BLOCKSIZE + 1000, wrappers.cpp:79; half of that is never reached.)
    unsigned long long p = 2 * (g.chunk_len + g.chunk_len / 512 + 640ull * g.blocks_per_chunk + 64ull);
    return (p + 15ull) & ~15ull;
}

// exact n / d for n <= 2^31, d >= 2: q = umulhi(n, mul) >> sh with l = ceil(log2 d),
// mul = ceil(2^(31+l) / d), sh = l - 1
struct Magic { uint32_t mul, sh; bool one; };
__device__ __forceinline__ Magic make_magic(uint32_t d)
{
    Magic m;
    m.one = d <= 1;
    const uint32_t l2 = m.one ? 1 : 32 - __clz(d - 1);
    m.mul = m.one ? 0 : (uint32_t)(((1ull << (31 + l2)) + d - 1) / d);
    m.sh = l2 - 1;
    return m;
}
__device__ __forceinline__ uint32_t div_magic(uint32_t n, const Magic& m)
{
    return m.one ? n : (__umulhi(n, m.mul) >> m.sh);
}

// ------------------------------------------------------------------------------------------
// encoder
// ------------------------------------------------------------------------------------------
struct Enc {
    uint32_t low, range, pos;    // pos = raw entries written (entry 0 is the lead byte)
    uint16_t* raw;
    unsigned long long acc;      // PACK: the entries of the current group of four (pos & ~3 ...), not yet stored
};

// PACK: raw entries leave in 8-byte stores of four.  Every lane writes its own chunk's scratch, so a 2-byte store per
// lane is 32 scattered sector writes per warp instruction; a layer of nearly incompressible symbols makes one per
// symbol and lane and the L2's write path, not the recurrence, sets the pace (512^3 f64 at 8 layers: 1.06 ms per layer
// against 0.63 at 3).  Packing costs ~8 instructions per symbol, which the one-warp-per-scheduler case has to spare only
// partly -- so it is a launch-time choice (range_encode_chunks).
template <bool PACK>
__device__ __forceinline__ void enc_emit(Enc& e, uint32_t v)
{
    if (!PACK) { e.raw[e.pos++] = (uint16_t)v; return; }
    const uint32_t slot = e.pos & 3u;
    e.acc |= (unsigned long long)v << (16u * slot);
    e.pos++;
    if (slot == 3u) { *reinterpret_cast<unsigned long long*>(e.raw + (e.pos - 4u)) = e.acc; e.acc = 0ull; }
}
template <bool PACK>
__device__ __forceinline__ void enc_flush_pack(Enc& e)
{
    if (!PACK) return;
    for (uint32_t i = e.pos & ~3u; i < e.pos; i++) e.raw[i] = (uint16_t)(e.acc >> (16u * (i & 3u)));
}

// generic renormalisation (markers, raw shorts, flush).  rangecod.c:182-207 minus the carry logic
template <bool PACK>
__device__ __forceinline__ void enc_renorm(Enc& e)
{
    while (e.range <= kBottom) {
        enc_emit<PACK>(e, e.low >> kShiftBits);
        e.range <<= 8;
        e.low = (e.low << 8) & (kTop - 1);
    }
}

// rangecod.c:217-229 with tot == 2 (block / end markers, wrappers.cpp:95,131)
template <bool PACK>
__device__ __forceinline__ void enc_marker(Enc& e, uint32_t bit)
{
    enc_renorm<PACK>(e);
    const uint32_t r = e.range >> 1;
    if (bit) { e.low += r; e.range -= r; }   // sy=1, lt=1: lt+sy == tot
    else     { e.range = r; }                // sy=1, lt=0
}

// rangecod.c:231-245 as encode_short (rangecod.h:155)
template <bool PACK>
__device__ __forceinline__ void enc_short(Enc& e, uint32_t v)
{
    enc_renorm<PACK>(e);
    const uint32_t r = e.range >> 16, t = r * v;
    e.low += t;
    if ((v + 1) >> 16) e.range -= t; else e.range = r;
}

// rangecod.c:254-276.  bytecount of the reference == number of shifts == pos - 1.
template <bool PACK>
__device__ __forceinline__ void enc_finish(Enc& e)
{
    enc_renorm<PACK>(e);
    const uint32_t count = (e.pos - 1) + 5;
    uint32_t t = e.low >> kShiftBits;
    if (!((e.low & (kBottom - 1)) < ((count & 0xFFFFFFu) >> 1))) t += 1;
    enc_emit<PACK>(e, t | kTerm);                              // carries like any entry (t > 0xFF)
    enc_emit<PACK>(e, ((count >> 16) & 0xFFu) | kTerm);
    enc_emit<PACK>(e, ((count >> 8) & 0xFFu) | kTerm);
    enc_emit<PACK>(e, (count & 0xFFu) | kTerm);
    enc_flush_pack<PACK>(e);
}

// Code one symbol (rangecod.c:217-229).  ent = cum << 16 | count; `last`: the symbol is the last
// one with a non-zero count, the only one with lt + sy == tot.  Branch-free: the (at most two) raw
// entries go out through predicated stores and the shifted state is chosen with selects.  ONE: the block
// holds a single symbol (tot == 1: no division); decided per block, outside the symbol loop, so that the
// choice costs no select on the range recurrence.
template <bool ONE, bool PACK>
__device__ __forceinline__ void enc_symbol(Enc& e, uint32_t ent, bool last, const Magic& mg)
{
    const uint32_t range = e.range, low = e.low;
    const bool k1 = range <= kBottom, k2 = range <= (kBottom >> 8);
    if (PACK) {
        const uint32_t e1 = low >> kShiftBits, e2 = (low >> (kShiftBits - 8)) & 0xFFu;
        const uint32_t vm = k2 ? (e1 | (e2 << 16)) : (k1 ? e1 : 0u);          // the entries of this symbol, first one low
        const uint32_t slot = e.pos & 3u;
        e.acc |= (unsigned long long)vm << (16u * slot);                      // slot 3 with two entries: the second falls off the top
        const uint32_t np = e.pos + (k1 ? 1u : 0u) + (k2 ? 1u : 0u);
        if ((np ^ e.pos) & 4u) {                                              // the group of four is complete
            *reinterpret_cast<unsigned long long*>(e.raw + (e.pos & ~3u)) = e.acc;
            e.acc = (slot == 3u && k2) ? (unsigned long long)e2 : 0ull;
        }
        e.pos = np;
    } else {
        uint16_t* w = e.raw + e.pos;
        // (plain `if (k1) w[0] = ...` makes the compiler branch around the stores: 38 instructions per symbol and a
        //  reconvergence point; the predicated stores are spelled out instead)
        asm volatile(
            "{\n\t"
            ".reg .pred p1, p2;\n\t"
            "setp.le.u32 p1, %1, 0x800000;\n\t"
            "setp.le.u32 p2, %1, 0x8000;\n\t"
            "@p1 st.global.u16 [%0], %2;\n\t"
            "@p2 st.global.u16 [%0+2], %3;\n\t"
            "}"
            :: "l"(w), "r"(range), "h"((uint16_t)(low >> kShiftBits)), "h"((uint16_t)((low >> (kShiftBits - 8)) & 0xFFu))
            : "memory");
        e.pos += (k1 ? 1u : 0u) + (k2 ? 1u : 0u);
    }
    const uint32_t rs = k2 ? (range << 16) : (k1 ? (range << 8) : range);
    const uint32_t ls = k2 ? ((low << 16) & (kTop - 1)) : (k1 ? ((low << 8) & (kTop - 1)) : low);   // carry bit survives k == 0
    const uint32_t r = ONE ? rs : (__umulhi(rs, mg.mul) >> mg.sh);               // exact range / bs
    const uint32_t t = r * (ent >> 16);
    e.low = ls + t;
    e.range = last ? rs - t : r * (ent & 0xFFFFu);
}

// grid (ceil(nchunks/32), layers), block 32: lane == chunk.
// COMPACT: the per-lane tables hold 16-bit cumulative counts only, two per word ([symbol pair][lane]: conflict-free like
// the full form) -- 16.5 KB per warp instead of 32 KB, so twice the warps fit an SM; an entry (cum, count) is then two
// loads and a subtraction, ~6 instructions more per symbol, all off the recurrence.  Measured slower even where the grid
// exceeds what is resident (1024^3 f64, 5 layers, 2797 warps: 16.3 vs 14.1 ms), so it is only a switch (WRB_ENC_TABLES).
template <bool COMPACT, bool PACK>
__global__ void __launch_bounds__(32) range_encode_kernel(const uint8_t* __restrict__ sym,
                                                          unsigned long long sym_layer_stride,
                                                          const uint32_t* __restrict__ hist,
                                                          unsigned long long hist_layer_stride, ChunkGeom g,
                                                          const int* __restrict__ active, uint8_t* __restrict__ slots,
                                                          unsigned long long slot_pitch,
                                                          unsigned long long* __restrict__ lens,
                                                          uint32_t* __restrict__ seek)
{
    const int layer = blockIdx.y;
    if (active != nullptr && !active[layer]) return;
    __shared__ uint32_t tab[(COMPACT ? 129 : 256) * 32];      // [symbol][lane] = cum << 16 | count; COMPACT: [symbol / 2][lane] = cum pair
    const unsigned int lane = threadIdx.x;
    const unsigned int chunk = blockIdx.x * 32 + lane;
    if (chunk >= g.nchunks) return;
    const unsigned long long id = (unsigned long long)layer * g.nchunks + chunk;
    const unsigned long long cstart = (unsigned long long)chunk * g.chunk_len;
    const unsigned long long clen = (g.ntot - cstart < g.chunk_len) ? g.ntot - cstart : g.chunk_len;
    const uint8_t* __restrict__ in = sym + (unsigned long long)layer * sym_layer_stride + (unsigned long long)chunk * g.pitch;
    const uint32_t* __restrict__ hrow = hist + (unsigned long long)layer * hist_layer_stride +
                                        (unsigned long long)chunk * g.blocks_per_chunk * 256;
    uint32_t* __restrict__ sk = seek + id * g.nseek * 3;
    Enc e;
    e.low = 0; e.range = kTop;                                             // rangecod.c:170-176
    e.raw = reinterpret_cast<uint16_t*>(slots + id * slot_pitch);
    e.acc = 0ull;
    if (!PACK) e.raw[0] = 0;                                               // lead byte (PACK: entry 0 of the first group)
    e.pos = 1;
    const uint32_t* __restrict__ tl = tab + lane;
    auto table_entry = [&](uint32_t c) -> uint32_t {              // cum << 16 | count of symbol c
        if (!COMPACT) return tl[c * 32];
        const uint32_t w0 = tl[(c >> 1) * 32], w1 = tl[((c + 1) >> 1) * 32];
        const uint32_t lo = (c & 1u) ? (w0 >> 16) : (w0 & 0xFFFFu);                  // cum[c]
        const uint32_t hi = (c & 1u) ? (w1 & 0xFFFFu) : (w0 >> 16);                  // cum[c + 1]
        return (lo << 16) | (hi - lo);
    };
    const uint32_t sub16 = g.sub_len >> 4;        // seek interval in 16-symbol groups (0: none)
    unsigned long long done = 0;
    for (;;) {                                                              // wrappers.cpp:85-128
        const uint32_t bs = (clen - done < kBlock) ? (uint32_t)(clen - done) : kBlock;
        enc_marker<PACK>(e, 1);
        uint32_t cum = 0, lastsym = 0;
        for (int s = 0; s < 256; s += 4) {
            const uint4 c4 = *reinterpret_cast<const uint4*>(hrow + s);
            const uint32_t cc[4] = {c4.x, c4.y, c4.z, c4.w};
            uint32_t pairw = 0;
#pragma unroll
            for (int k = 0; k < 4; k++) {
                enc_short<PACK>(e, cc[k]);
                if (COMPACT) {
                    if (k & 1) tab[((s + k) >> 1) * 32 + lane] = pairw | (cum << 16); else pairw = cum;
                } else {
                    tab[(s + k) * 32 + lane] = (cum << 16) | cc[k];
                }
                cum += cc[k];
                if (cc[k]) lastsym = s + k;
            }
        }
        if (COMPACT) tab[128 * 32 + lane] = cum;                             // cum[256] = block size (< 2^16)
        enc_renorm<PACK>(e);                      // up to three shifts may be pending after a raw short
        const Magic mg = make_magic(bs);
        const uint4* __restrict__ p = reinterpret_cast<const uint4*>(in + done);
        const uint32_t nfull = bs >> 4;
        auto code_block = [&](auto one_tag) {
        constexpr bool ONE = decltype(one_tag)::value;
        uint4 w = (bs > 0) ? p[0] : make_uint4(0, 0, 0, 0);
        uint32_t until_seek = sub16, nsk = 0;
        for (uint32_t i = 0; i < nfull; i++) {
            if (g.nseek) {                                                 // seek point before symbol 16*i
                if (until_seek == 0) {
                    if (nsk < g.nseek) { sk[nsk * 3 + 0] = e.low; sk[nsk * 3 + 1] = e.range; sk[nsk * 3 + 2] = e.pos; nsk++; }
                    until_seek = sub16;
                }
                until_seek--;
            }
            const uint4 wn = p[i + 1];                                     // next 16 symbols (pitch slack covers it)
            // every lane streams its own chunk: 32 sectors per load, DRAM latency well above one group's ~1400 cycles
            // when the symbols have left L2 -- pull the line eight groups ahead into L2 (prefetches never fault)
            if ((i & 7u) == 0) asm volatile("prefetch.global.L2 [%0];" :: "l"(p + i + 16));
            const uint32_t ww[4] = {w.x, w.y, w.z, w.w};
            uint32_t cs[16], ent[16];
#pragma unroll
            for (int k = 0; k < 16; k++) {
                cs[k] = (ww[k >> 2] >> ((k & 3) * 8)) & 0xFFu;
                ent[k] = table_entry(cs[k]);
            }
#pragma unroll
            for (int k = 0; k < 16; k++) enc_symbol<ONE, PACK>(e, ent[k], cs[k] == lastsym, mg);
            w = wn;
        }
        {
            const uint32_t ww[4] = {w.x, w.y, w.z, w.w};
            const uint32_t rem = bs & 15u;
            if (g.nseek && rem && until_seek == 0 && nsk < g.nseek) {     // seek point at the remainder's first symbol
                sk[nsk * 3 + 0] = e.low; sk[nsk * 3 + 1] = e.range; sk[nsk * 3 + 2] = e.pos; nsk++;
            }
            for (uint32_t k = 0; k < rem; k++) {
                const uint32_t c = (ww[k >> 2] >> ((k & 3) * 8)) & 0xFFu;
                enc_symbol<ONE, PACK>(e, table_entry(c), c == lastsym, mg);
            }
        }
        for (; g.nseek && nsk < g.nseek; nsk++) { sk[nsk * 3 + 0] = 0; sk[nsk * 3 + 1] = 0; sk[nsk * 3 + 2] = 0; }
        };
        if (mg.one) code_block(std::true_type{}); else code_block(std::false_type{});     // single-symbol blocks are rare
        done += bs;
        hrow += 256;
        if (bs < kBlock) break;
    }
    enc_marker<PACK>(e, 0);
    enc_finish<PACK>(e);
    lens[id] = (unsigned long long)e.pos;
}

void range_encode_chunks(const uint8_t* sym, unsigned long long sym_layer_stride, const uint32_t* hist,
                         unsigned long long hist_layer_stride, const ChunkGeom& g, int nlayers, const int* active,
                         uint8_t* slots, unsigned long long slot_pitch, unsigned long long* lens, uint32_t* seek,
                         cudaStream_t s)
{
    dim3 grid((g.nchunks + 31) / 32, nlayers, 1);
    const char* e = getenv("WRB_ENC_TABLES");                        // "full" / "compact": force a form (tests, A/B timing)
    // measured (1024^3 f64, 5 layers, 2797 warps): compact 16.3 ms, full 14.1 ms -- the extra instructions cost more than
    // the doubled residency buys, so the full form is the default at every size and the compact one stays a switch
    bool compact = false;
    if (e && *e == 'f') compact = false;
    if (e && *e == 'c') compact = true;
    const char* pe = getenv("WRB_ENC_PACK");                         // "0" / "1": force scattered 2-byte / packed 8-byte stores of the raw entries
    // measured at 512^3: packed stores make the loop ~135 instead of 62 cycles per symbol whatever the data (a second
    // dependency chain through the accumulator), scattered stores cost 1.88 ms at 3 layers, 3.3 at 5, 8.5 at 8: packing
    // pays only when most layers are nearly incompressible (4.1 ms at 8 layers)
    bool pack = nlayers >= 7;
    if (pe && *pe) pack = atoi(pe) != 0;
#define WRB_ENC_LAUNCH(C, P) range_encode_kernel<C, P><<<grid, 32, 0, s>>>(sym, sym_layer_stride, hist, hist_layer_stride, g, active, slots, slot_pitch, lens, seek)
    if (compact) { if (pack) WRB_ENC_LAUNCH(true, true); else WRB_ENC_LAUNCH(true, false); }
    else { if (pack) WRB_ENC_LAUNCH(false, true); else WRB_ENC_LAUNCH(false, false); }
#undef WRB_ENC_LAUNCH
    note_launch(1);
}

// ------------------------------------------------------------------------------------------
// container assembly.  Layer blob (chunked mode):
//   [ 0] "WRCK"  [ 4] u32 version = 2  [ 8] u64 chunk_len  [16] u64 symbols in layer
//   [24] u32 nchunks  [28] u32 nseek (seek points per chunk)
//   u32 len[nchunks]                      byte length of every chunk stream
//   u32 seek[nchunks][nseek][3]           (low, range, stream position) before symbol (j+1)*sub_len
//   chunk streams back to back
// Single-stream mode: the bare stream (identical to the reference's layer).
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void put_le(uint8_t* p, unsigned long long v, int nbytes)
{
    for (int k = 0; k < nbytes; k++) p[k] = (uint8_t)(v >> (8 * k));
}

constexpr unsigned int kSeekBytes = 10;       // u32 low, u32 range, u16 stream position relative to the previous point
constexpr unsigned int kContainerVersion = 3;

__host__ __device__ inline unsigned long long container_header_bytes(const ChunkGeom& g, int chunked, unsigned int nseek)
{
    return chunked ? 32ull + 4ull * g.nchunks + (unsigned long long)kSeekBytes * nseek * g.nchunks : 0ull;
}

// Seek points cost bytes (10 per point and chunk) and buy decoder lanes: the decoder keeps a chunk's tables in shared
// memory, so lanes per chunk are also what fills an SM (measured at 1024^3 f32, 53.7 k chunks: 0 / 1 / 3 / 7 points
// decode in 75.9 / 39.6 / 21.8 / 13.5 ms).  With seek_auto the container keeps as many points of the recorded grid as
// fit into kSeekBudget of the coded bytes for the chunk tables (length + seek entries), so that it stays within 1 % of
// the reference's layer streams whatever the data.
constexpr double kSeekBudget = 0.0085;

__global__ void __launch_bounds__(1024) sum_lens_kernel(const unsigned long long* __restrict__ lens, ChunkGeom g, const DevState* st,
                                                        unsigned long long* __restrict__ out)
{
    __shared__ unsigned long long s_part[1024];
    const int t = threadIdx.x;
    const int nlay = st->nlay;
    unsigned long long sum = 0;
    for (int l = 0; l < nlay; l++)
        for (unsigned int c = t; c < g.nchunks; c += 1024) sum += lens[(unsigned long long)l * g.nchunks + c];
    s_part[t] = sum;
    __syncthreads();
    for (int o = 512; o > 0; o >>= 1) { if (t < o) s_part[t] += s_part[t + o]; __syncthreads(); }
    if (t == 0) { out[0] = s_part[0]; out[1] = (unsigned long long)g.nchunks * (unsigned long long)(nlay > 0 ? nlay : 0); }
}

void sum_chunk_lens(const unsigned long long* lens, const ChunkGeom& g, const DevState* st, unsigned long long* out, cudaStream_t s)
{
    sum_lens_kernel<<<1, 1024, 0, s>>>(lens, g, st, out);
    note_launch(1);
}

__global__ void __launch_bounds__(1024) assemble_scan_kernel(const unsigned long long* __restrict__ lens, ChunkGeom g,
                                                             int chunked, int seek_auto, DevState* st,
                                                             uint8_t* __restrict__ blob, unsigned long long cap,
                                                             unsigned long long* __restrict__ dst_off,
                                                             const unsigned long long* __restrict__ gtot)
{
    __shared__ unsigned long long s_part[1024];
    __shared__ unsigned long long s_base;
    __shared__ unsigned int s_keep;
    const int t = threadIdx.x;
    const unsigned int per = (g.nchunks + 1023) / 1024;
    const unsigned int c0 = t * per, c1 = (c0 + per < g.nchunks) ? c0 + per : g.nchunks;
    const int nlay = st->nlay;
    if (t == 0) { s_base = 0; s_keep = chunked ? g.nseek : 0; }
    if (chunked && seek_auto && g.nseek > 0) {            // how many of the recorded seek points to keep (one answer for all layers)
        unsigned long long sum = 0;
        for (int l = 0; l < nlay; l++)
            for (unsigned int c = c0; c < c1; c++) sum += lens[(unsigned long long)l * g.nchunks + c];
        s_part[t] = sum;
        __syncthreads();
        for (int o = 512; o > 0; o >>= 1) { if (t < o) s_part[t] += s_part[t + o]; __syncthreads(); }
        if (t == 0) {
            const unsigned long long chunks = gtot ? gtot[1] : (unsigned long long)g.nchunks * (unsigned long long)(nlay > 0 ? nlay : 1);
            const double budget = kSeekBudget * (double)(gtot ? gtot[0] : s_part[0]);
            unsigned int keep = g.nseek;
            while (keep > 0 && (double)(chunks * (4ull + (unsigned long long)kSeekBytes * keep)) > budget) keep = (keep - 1) / 2;
            s_keep = keep;
        }
    }
    __syncthreads();
    const unsigned int keep = s_keep;
    if (t == 0) st->nseek_keep = (int)keep;
    for (int l = 0; l < nlay; l++) {
        const unsigned long long* ll = lens + (unsigned long long)l * g.nchunks;
        unsigned long long sum = 0;
        for (unsigned int c = c0; c < c1; c++) sum += ll[c];
        s_part[t] = sum;
        __syncthreads();
        for (int o = 1; o < 1024; o <<= 1) {          // inclusive scan of the 1024 partial sums
            unsigned long long v = (t >= o) ? s_part[t - o] : 0;
            __syncthreads();
            s_part[t] += v;
            __syncthreads();
        }
        const unsigned long long total = s_part[1023];
        unsigned long long off = s_part[t] - sum;
        const unsigned long long base = s_base;
        const unsigned long long hdr = container_header_bytes(g, chunked, keep);
        const bool fits = base + hdr + total <= cap;
        if (fits) {
            if (chunked && t == 0) {
                uint8_t* h = blob + base;
                h[0] = 'W'; h[1] = 'R'; h[2] = 'C'; h[3] = 'K';
                put_le(h + 4, kContainerVersion, 4);
                put_le(h + 8, g.chunk_len, 8);
                put_le(h + 16, g.ntot, 8);
                put_le(h + 24, g.nchunks, 4);
                put_le(h + 28, keep, 4);
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
            st->lay_off[l] = base;
            st->len_enc[l] = hdr + total;
            s_base = base + hdr + total;
        }
        __syncthreads();
    }
    if (t == 0) st->ntot_enc = s_base;
}

// carry of a raw entry: 1 if its 9/10-bit value exceeds 0xFF
__device__ __forceinline__ uint32_t raw_carry(uint32_t v) { return ((v & 0x7FFFu) > 0xFFu) ? 1u : 0u; }

// one CTA per (chunk, layer): resolve carries and move the stream to its place in the blob
__global__ void __launch_bounds__(256) assemble_copy_kernel(const uint8_t* __restrict__ slots,
                                                            unsigned long long slot_pitch,
                                                            const unsigned long long* __restrict__ lens,
                                                            const unsigned long long* __restrict__ dst_off,
                                                            const uint32_t* __restrict__ seek, ChunkGeom g, int chunked,
                                                            const DevState* st, uint8_t* __restrict__ blob)
{
    const int l = blockIdx.y;
    if (l >= st->nlay || st->error) return;
    const unsigned long long id = (unsigned long long)l * g.nchunks + blockIdx.x;
    const uint16_t* __restrict__ raw = reinterpret_cast<const uint16_t*>(slots + id * slot_pitch);
    uint8_t* __restrict__ dst = blob + dst_off[id];
    const uint32_t n = (uint32_t)lens[id];
    // eight entries per thread (one 16-byte load; slots are 16-byte aligned): the carry into entry i is decided by the
    // first later entry that is not a plain 0xFF, resolved backwards inside the group from the decision for entry 8
    for (uint32_t j = 8u * threadIdx.x; j < n; j += 8u * blockDim.x) {
        const uint4 q = *reinterpret_cast<const uint4*>(raw + j);
        const uint32_t w[4] = {q.x, q.y, q.z, q.w};
        const uint32_t m = (n - j < 8u) ? n - j : 8u;        // valid entries of this group
        uint32_t k = j + m, dec = 0;                          // decision carried into the group's last valid entry
        while (k < n) {
            const uint32_t u = raw[k];
            if ((u & kTerm) || (u & 0x7FFFu) != 0xFFu) { dec = raw_carry(u); break; }
            k++;
        }
        uint8_t ob[8];
#pragma unroll
        for (int i = 7; i >= 0; i--) {
            const uint32_t v = (w[i >> 1] >> ((i & 1) * 16)) & 0xFFFFu;
            if ((uint32_t)i < m) {
                ob[i] = (uint8_t)((v & 0x7FFFu) + dec);
                if ((v & kTerm) || (v & 0x7FFFu) != 0xFFu) dec = raw_carry(v);     // a plain 0xFF passes the decision on
            }
        }
#pragma unroll
        for (int i = 0; i < 8; i++) if ((uint32_t)i < m) dst[j + i] = ob[i];
    }
    const uint32_t keep = (uint32_t)st->nseek_keep;
    if (chunked && keep) {
        // every stride-th recorded point, 10 bytes each: low, range, stream position as the distance from the previous
        // kept point (< 2^16: a sub-range of <= 30016 symbols codes into < 16 bits per symbol plus the 513-byte table)
        const uint32_t stride = (g.nseek + 1) / (keep + 1);
        uint8_t* sdst = blob + st->lay_off[l] + 32 + 4ull * g.nchunks + (unsigned long long)kSeekBytes * keep * blockIdx.x;
        const uint32_t* ssrc = seek + id * g.nseek * 3;
        for (uint32_t t = threadIdx.x; t < keep; t += blockDim.x) {
            const uint32_t* e = ssrc + ((t + 1) * stride - 1) * 3;
            const uint32_t prev = t ? ssrc[(t * stride - 1) * 3 + 2] : 0u;
            put_le(sdst + kSeekBytes * t, e[0], 4);
            put_le(sdst + kSeekBytes * t + 4, e[1], 4);
            put_le(sdst + kSeekBytes * t + 8, e[2] >= prev ? e[2] - prev : 0u, 2);     // unused points (short last chunk) are zero
        }
    }
}

void assemble_container(const uint8_t* slots, unsigned long long slot_pitch, const unsigned long long* lens,
                        const uint32_t* seek, const ChunkGeom& g, int chunked, int seek_auto, DevState* st, uint8_t* blob,
                        unsigned long long cap, unsigned long long* dst_off, cudaStream_t s, const unsigned long long* gtot)
{
    assemble_scan_kernel<<<1, 1024, 0, s>>>(lens, g, chunked, seek_auto, st, blob, cap, dst_off, gtot);
    dim3 grid(g.nchunks, kNLayMax, 1);
    assemble_copy_kernel<<<grid, 256, 0, s>>>(slots, slot_pitch, lens, dst_off, seek, g, chunked, st, blob);
    note_launch(2);
}

// ------------------------------------------------------------------------------------------
// decoder
// ------------------------------------------------------------------------------------------
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
        const unsigned long long hdr = container_header_bytes(g, 1, g.nseek);
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
            offs[(unsigned long long)l * g.nchunks + c] = base + hdr + off;
            off += len;
        }
        if (t == 0 && base + hdr + total != lay_off[l + 1]) *error = 2;       // table / length mismatch
        __syncthreads();
    }
}

// floor(a / b) for a < 2^31, 0 < b, a / b < 2^17.  The float estimate is biased low (operands rounded
// towards a smaller quotient, reciprocal nudged two ulps down), so it is q or q - 1: one fix-up.
__device__ __forceinline__ uint32_t div_small_quot(uint32_t a, uint32_t b)
{
    float rb;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rb) : "f"(__uint2float_ru(b)));      // <= 1 ulp off
    rb = __int_as_float(__float_as_int(rb) - 2);                                  // now <= 1 / b
    uint32_t q = __float2uint_rz(__fmul_rz(__uint2float_rz(a), rb));              // q_true - 1 <= q <= q_true
    return q + ((a - q * b >= b) ? 1u : 0u);
}
// The same for b < 2^24 and a / b < 2^23 (the symbol loop: b = range / tot <= 2^31 / 256 for blocks of >= 256
// symbols, quotient < tot < 2^16): b converts exactly, so the round-up conversion (a slow-pipe instruction) is not
// needed, and floor() of the estimate is taken with the 2^23 trick (add with round-towards-zero leaves the integer
// part in the mantissa) instead of a float-to-int conversion -- two slow-pipe instructions fewer on the chain.
__device__ __forceinline__ uint32_t div_small_quot_fast(uint32_t a, uint32_t b)
{
    float rb;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rb) : "f"(__uint2float_rz(b)));      // exact operand, result <= 1 ulp off
    rb = __int_as_float(__float_as_int(rb) - 2);                                  // now <= 1 / b
    const float est = __fmul_rz(__uint2float_rz(a), rb);                          // <= a / b
    const uint32_t q = __float_as_uint(__fadd_rz(est, 8388608.0f)) & 0x7FFFFFu;   // floor(est); q_true - 1 <= q <= q_true
    return q + ((a - q * b >= b) ? 1u : 0u);
}

// Decoder state in "X form": X = (low << 1) | (lowest bit of the last byte read), i.e. the last four
// stream bytes read minus twice the coder's cumulative low; a renormalisation step (rangecod.c:293-300)
// is then X = X << 8 | next byte.
struct Dec {
    uint32_t X, range;
    uint32_t ip;                    // index of the next stream byte
    uint32_t lim;                   // last byte index that may be read: stream length + kDecSlack - 1
    const uint8_t* p;               // stream start
};
// A well-formed stream is never read more than a few bytes past its end (the reference's decoder does the same); a
// corrupt one must not walk out of the blob: reads are clamped to stream end + kDecSlack (inside the 64 bytes of slack
// every blob has) and a position beyond that flags the chunk as malformed.
constexpr uint32_t kDecSlack = 12;

__device__ __forceinline__ void dec_renorm(Dec& d)
{
    while (d.range <= kBottom) { d.X = (d.X << 8) | d.p[min(d.ip, d.lim)]; d.ip++; d.range <<= 8; }
}

// rangecod.c:321-331 + :362-366 (decode_short), :339-351 (decode_update)
__device__ __forceinline__ uint32_t dec_short(Dec& d)
{
    dec_renorm(d);
    const uint32_t help = d.range >> 16;
    uint32_t t = (d.X >> 1) / help;
    if (t >> 16) t = 0xFFFFu;
    const uint32_t tmp = help * t;
    d.X -= 2 * tmp;
    if (t + 1 < (1u << 16)) d.range = help; else d.range -= tmp;
    return t;
}

constexpr int kDecVariantDefault = 3;     // bit 0: lazy stream loads, bit 1: L1 prefetch of the stream two sectors ahead (see range_decode_kernel)
// Table footprint per chunk decides how many chunks an SM can hold (227 KB of shared memory), and with one or two lanes
// per chunk that is what bounds the decoder.  Two forms: compact (16-bit cumulative counts only -- a count is the
// difference of two neighbours -- and 128-wide LUT buckets: 1 KB per chunk) and fast (cum/count entry pairs, 64-wide
// buckets: 3 KB, three instructions per symbol fewer).
// The form is a launch-time choice (decode_uses_pair_tables): the fast form while the whole grid is resident with it,
// the compact form for one and two lanes per chunk and whenever shared memory would cap the resident blocks.
__host__ __device__ constexpr int dec_lut_shift(bool pair) { return pair ? 6 : 7; }
__host__ __device__ constexpr int dec_lut_size(bool pair) { return (kBlock >> dec_lut_shift(pair)) + 3; }   // 938 / 469 buckets + 2 beyond the last
__host__ __device__ constexpr int dec_table_bytes(bool pair)     // per chunk: symbol table + LUT (rounded to words)
{
    return (pair ? 257 * 8 : 258 * 2) + ((dec_lut_size(pair) + 3) & ~3);
}
static bool decode_uses_pair_tables(unsigned int nsub, unsigned long long blocks)
{
    const char* e = getenv("WRB_DEC_TABLES");                     // "pair" / "compact": force a form (tests, A/B timing)
    if (e && *e == 'p') return true;
    if (e && *e == 'c') return false;
    if (nsub < 4) return false;
    const unsigned long long per_block = (unsigned long long)dec_table_bytes(true) * (32 / nsub);
    unsigned long long resident = (227ull * 1024) / per_block;    // blocks of one warp per SM, by shared memory
    if (resident > 32) resident = 32;
    return blocks <= 148ull * resident;
}

// grid (ceil(nchunks/CPW), layers), block 32: lane == (chunk column, sub-chunk); NSUB lanes decode one
// chunk, CPW = 32/NSUB chunks per warp.   wrappers.cpp:153-224
// Symbol search: the reference builds a 60001-entry inverse table per block (wrappers.cpp:191-196);
// here a 1-byte LUT over cf >> 6 gives the first symbol whose interval meets the bucket, followed by
// a short forward scan over the packed (cum, count) table (zero-count symbols are skipped by the
// same scan, wrappers.cpp:205).  The lanes that share a chunk share its tables (column = chunk slot
// inside the warp): all of them decode the identical table and store identical values.
// The symbol loop is one dependency chain per lane; with one or two lanes per chunk the warps are few and the length
// of that chain counts, with eight lanes per chunk (~3 warps per scheduler) the number of instructions does:
//   * stream words come through the read-only path (ld.global.nc);
//   * renormalisation is branch-free (one funnel shift by 0/8/16);
//   * pair table: entry s holds (ent[s], ent[s+1]) so one 64-bit shared load serves both probes;
//   * symbols leave as 32-bit words (groups of four), bytes only for a ragged head/tail.
// LAZY (WRB_DEC_VARIANT bit 0, for A/B timing; both bit-identical): a new stream word is fetched (predicated) only
// by the lanes that crossed a word boundary, one word ahead of its use, instead of two 32-sector loads per symbol.
template <int NSUB, bool LAZY, bool PAIR, bool PF>
__global__ void __launch_bounds__(32) range_decode_kernel(const uint8_t* __restrict__ blob,
                                                          const unsigned long long* __restrict__ offs,
                                                          const unsigned long long* __restrict__ lay_off, ChunkGeom g,
                                                          uint8_t* __restrict__ sym, unsigned long long sym_layer_stride,
                                                          unsigned long long blob_len, int* error)
{
    constexpr unsigned int CPW = 32 / NSUB;
    constexpr bool kPair = PAIR, kLazy = LAZY;
    constexpr int kLutShift = dec_lut_shift(PAIR), kLutSize = dec_lut_size(PAIR);
    constexpr unsigned int TW = kPair ? 2 : 1;                            // words per table entry
    extern __shared__ __align__(16) uint32_t smem_dyn[];
    uint32_t* tab = smem_dyn;                                             // [symbol][column] = cum << 16 | count (+ sentinel row)
    uint16_t* cum16 = reinterpret_cast<uint16_t*>(smem_dyn);              // compact form: [symbol 0..257][column] = cum
    uint8_t* lut = kPair ? reinterpret_cast<uint8_t*>(smem_dyn + 257 * CPW * TW)      // [bucket][column]
                         : reinterpret_cast<uint8_t*>(smem_dyn) + 258 * CPW * 2;
    const int layer = blockIdx.y;
    const unsigned int lane = threadIdx.x;
    const unsigned int col = lane / NSUB, sub = lane % NSUB;
    const unsigned int chunk = blockIdx.x * CPW + col;
    if (chunk >= g.nchunks) return;
    const unsigned long long cstart = (unsigned long long)chunk * g.chunk_len;
    const unsigned long long clen = (g.ntot - cstart < g.chunk_len) ? g.ntot - cstart : g.chunk_len;
    if (sub > 0 && (unsigned long long)sub * g.sub_len >= clen) return;       // nothing for this lane
    uint8_t* __restrict__ outb = sym + (unsigned long long)layer * sym_layer_stride + (unsigned long long)chunk * g.pitch;
    Dec d;
    // byte range of this chunk's stream; everything read below stays inside [start, end + kDecSlack)
    const unsigned long long oidx = (unsigned long long)layer * g.nchunks + chunk;
    const unsigned long long cbeg = offs[oidx];
    const unsigned long long cend = (chunk + 1 < g.nchunks) ? offs[oidx + 1] : (lay_off != nullptr ? lay_off[layer + 1] : blob_len);
    if (cend > blob_len || cbeg + 8 > cend || cend - cbeg > 0x7FFFFFFFull) { atomicExch(error, 4); return; }   // shortest stream: 8 bytes
    const uint32_t slen = (uint32_t)(cend - cbeg);
    d.lim = slen + kDecSlack - 1;
    d.p = blob + cbeg;
    d.X = d.p[1];                                 // lead byte skipped (rangecod.c:283-288)
    d.ip = 2;
    d.range = 1u << 7;
    const uint32_t* tl = tab + col * TW;
    const uint8_t* ll = lut + col;
    unsigned long long n = 0;
    unsigned int nblocks = 0;
    bool bad = false;
    for (;;) {
        dec_renorm(d);                            // decode_culfreq(rc, 2)  (wrappers.cpp:174)
        uint32_t help = d.range >> 1;
        if ((d.X >> 1) < help) break;             // marker 0: no further block
        d.X -= 2 * help; d.range -= help;         // decode_update(rc,1,1,2)
        if (++nblocks > g.blocks_per_chunk) { bad = true; break; }
        uint32_t acc = 0, lastsym = 0;            // lastsym: last symbol with a non-zero count
        int nextb = 0;                            // first bucket not yet assigned
        for (int s = 0; s < 256; s++) {           // readcounts (rangecod.c:400-404) + prefix sums
            const uint32_t c = dec_short(d);
            if (kPair) {
                tab[(s * CPW + col) * TW] = (acc << 16) | c;
                if (s > 0) tab[((s - 1) * CPW + col) * TW + 1] = (acc << 16) | c;
            } else {
                cum16[s * CPW + col] = (uint16_t)acc;
            }
            if (c) {
                const int lastb = (int)((acc + c - 1) >> kLutShift);
                if (lastb < kLutSize) for (; nextb <= lastb; nextb++) lut[nextb * CPW + col] = (uint8_t)s;
                lastsym = (uint32_t)s;
            }
            acc += c;
            if (acc > kBlock) { bad = true; break; }
        }
        if (bad) break;
        for (; nextb < kLutSize; nextb++) lut[nextb * CPW + col] = (uint8_t)lastsym;      // buckets past the last symbol (search bound below)
        if (kPair) {                                     // entries past the last symbol: read, never chosen (s < lastsym guards)
            tab[(256 * CPW + col) * TW] = 0xFFFF0000u;
            tab[(255 * CPW + col) * TW + 1] = 0xFFFF0000u; tab[(256 * CPW + col) * TW + 1] = 0xFFFF0000u;
        } else {
            cum16[256 * CPW + col] = (uint16_t)acc;      // cum[256] = block size: count[255] = cum[256] - cum[255]
            cum16[257 * CPW + col] = (uint16_t)acc;
        }
        const uint32_t bs = acc;
        if (n + bs > clen) { bad = true; break; }
        uint32_t s0 = 0, s1 = bs;
        if (NSUB > 1) {                           // this lane's share of the (single) block
            if (bs != clen) { bad = true; break; }
            s0 = sub * g.sub_len;
            s1 = (s0 + g.sub_len < bs) ? s0 + g.sub_len : bs;
            if (sub > 0) {                        // restart from the seek point: X = W - 2*low
                const uint8_t* sp = blob + lay_off[layer] + 32 + 4ull * g.nchunks +
                                    (unsigned long long)kSeekBytes * ((unsigned long long)chunk * g.nseek + (sub - 1));
                const uint32_t slow = sp[0] | (sp[1] << 8) | (sp[2] << 16) | ((uint32_t)sp[3] << 24);
                const uint32_t srng = sp[4] | (sp[5] << 8) | (sp[6] << 16) | ((uint32_t)sp[7] << 24);
                uint32_t spos = 0;                // positions are stored as distances from the previous point
                for (unsigned int j = 0; j < sub; j++) {
                    const uint8_t* q = sp - (unsigned long long)kSeekBytes * j;
                    spos += q[8] | (q[9] << 8);
                }
                if (spos + 4 > slen) { bad = true; break; }           // the entry point must lie inside the stream
                const uint8_t* q = d.p + spos;
                const uint32_t W = ((uint32_t)q[0] << 24) | (q[1] << 16) | (q[2] << 8) | q[3];
                d.X = W - 2 * slow;
                d.range = srng;
                d.ip = spos + 4;
                if (srng == 0) { bad = true; break; }
            }
        }
        dec_renorm(d);                            // up to three shifts may be pending after a raw short
        const Magic mg = make_magic(bs);
        // stream window: win = next 4 stream bytes, big-endian, assembled from the two aligned words that
        // hold them.  The words for the NEXT symbol are loaded (L1 hits) as soon as this symbol's byte
        // consumption is known and are only touched at the top of the next iteration, so the loads
        // overlap the division / table look-ups; a prefetch runs one cache line ahead.
        const unsigned long long pa = (unsigned long long)d.p;
        const uint32_t* __restrict__ wbase = reinterpret_cast<const uint32_t*>(pa & ~3ull);
        const uint32_t boff = (uint32_t)(pa & 3ull);
        uint32_t a = boff + d.ip;
        const uint32_t wmax = (boff + d.lim) >> 2;            // last word that may be read (see kDecSlack)
        auto ldw = [&](uint32_t widx) -> uint32_t { return __ldg(wbase + min(widx, wmax)); };
        uint32_t w0 = ldw(a >> 2), w1 = ldw((a >> 2) + 1);
        uint32_t w2 = kLazy ? ldw((a >> 2) + 2) : 0u;
        uint32_t sel = 0x0123u + 0x1111u * (a & 3u);
        uint32_t X = d.X, range = d.range;
        // Symbols leave through 32-bit stores: a lane's run may start at any byte of the (flat) symbol buffer, so up to
        // three head symbols go out as bytes, then groups of four as one word, then the tail as bytes.  Stores are off
        // the dependency chain and the bytes a lane writes complete whole 32-byte sectors in L2 long before eviction.
        uint8_t* const op = outb + n + s0;
        const uint32_t cnt = s1 - s0;
        // two copies of the loop (a generic lambda instantiated for both values): the fast division needs
        // help < 2^24 in every lane, decided once per block
        auto symbol_loop = [&](auto fast_tag) {
        constexpr bool kFastDiv = decltype(fast_tag)::value;
        // One symbol (rangecod.c:309-319 decode_culfreq, wrappers.cpp:203-205 look-up, rangecod.c:339-351 decode_update).
        // The float quotient q is floor(V / help) or one less; it only picks the LUT bucket.  The symbol itself is found
        // with exact integer thresholds: s is the symbol iff help * cum[s] <= V < help * cum[s+1], which is cum[s] <= cf <
        // cum[s+1] for cf = floor(V / help) without ever forming cf (no fix-up on the chain).  The reference's clamp
        // cf <= tot - 1 (rangecod.c:318) only matters above the last symbol with a non-zero count: the search stops there.
        auto step = [&]() -> uint32_t {
            const uint32_t win = __byte_perm(w0, w1, sel);
            const bool k1 = range <= kBottom, k2 = range <= (kBottom >> 8);
            const uint32_t sh = k2 ? 16u : (k1 ? 8u : 0u);
            X = __funnelshift_l(win, X, sh);                                                    // renormalise
            range <<= sh;
            const uint32_t an = a + (sh >> 3);
            if (kLazy) {
                // at most 2 bytes per step: one word boundary.  The word fetched is the one AFTER the next: it is first
                // touched four stream bytes later, so even an L1 miss (one per 32-byte sector and lane, i.e. every few
                // symbols somewhere in the warp) is off the dependency chain.
                if ((an ^ a) & 4u) { w0 = w1; w1 = w2; w2 = ldw((an >> 2) + 2); }
                // every lane reads its own stream, so the first touch of a 32-byte sector is an L2 round trip for the whole
                // warp (r2 profile: 19 % of the loop's stall samples are long-scoreboard waits on these words): pull the
                // sector after next into L1 when a sector boundary is crossed
                if (PF && ((an ^ a) & 32u)) asm volatile("prefetch.global.L1 [%0];" :: "l"(wbase + min((an >> 2) + 16u, wmax)));
            } else {
                if ((an ^ a) & 0x80u) asm volatile("prefetch.global.L1 [%0];" :: "l"(wbase + min((an >> 2) + 64, wmax)));
                w0 = ldw(an >> 2);
                w1 = ldw((an >> 2) + 1);
            }
            a = an;
            sel = 0x0123u + 0x1111u * (a & 3u);
            const uint32_t V = X >> 1;
            uint32_t q;
            if (kFastDiv) {
                help = __umulhi(range, mg.mul) >> mg.sh;                 // decode_culfreq(rc, bs): bs >= 256 here
                float rb;
                asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rb) : "f"(__uint2float_rz(help)));      // help < 2^24 converts exactly
                rb = __int_as_float(__float_as_int(rb) - 2);                                    // now <= 1 / help
                const float est = __fmul_rz(__uint2float_rz(V), rb);                            // <= V / help
                q = __float_as_uint(__fadd_rz(est, 8388608.0f)) & 0x7FFFFFu;                    // floor(est): cf - 1 <= q <= cf
            } else {
                help = div_magic(range, mg);
                q = div_small_quot(V, help);
            }
            q = min(q, bs - 1);                   // V / help can exceed tot (the remainder range / tot leaves): last bucket
            uint32_t s = ll[(q >> kLutShift) * CPW];
            uint32_t lt, sy;                      // cum and count of the symbol found
            if (kPair) {
                const uint2 e = *reinterpret_cast<const uint2*>(tl + s * (CPW * 2));      // entries s and s + 1 in one load
                const bool adv = s < lastsym && help * (e.y >> 16) <= V;
                uint32_t ent = adv ? e.y : e.x;
                s += adv ? 1u : 0u;
                if (adv) {                        // rare: more than one step from the bucket's first symbol
                    const uint32_t nx = tl[(s + 1) * (CPW * TW)];
                    if (s < lastsym && help * (nx >> 16) <= V) {
                        // Third or later symbol of its bucket: where many symbols have small counts a bucket holds a dozen of
                        // them, and a linear scan costs every lane of the warp the longest scan among its 32 lanes at every
                        // symbol (measured: a run of such data decoded in 5.4 ms instead of 2.0).  Binary search for the
                        // largest symbol whose cumulative count does not exceed cf = V / help; it lies before the first
                        // symbol of the bucket after next (the estimate q is cf or cf - 1).  (Two linear steps before the
                        // search were measured too: 256^3 2.02 instead of 2.11 ms, but 512^3 at 8 layers 5.07 instead of 4.27.)
                        uint32_t lo = s + 1, hi = min(lastsym, (uint32_t)ll[((q >> kLutShift) + 2) * CPW]);
                        while (lo < hi) {
                            const uint32_t mid = (lo + hi + 1) >> 1;
                            if (help * (tl[mid * (CPW * TW)] >> 16) <= V) lo = mid; else hi = mid - 1;
                        }
                        s = lo;
                        ent = tl[s * (CPW * TW)];
                    }
                }
                // (testing help * (cum + count) of the chosen entry instead -- no load, branch rarely taken -- measured
                //  slower: 2.94 vs 2.69 ms at 8 lanes per chunk, 15.2 vs 13.0 ms at 2)
                lt = ent >> 16; sy = ent & 0xFFFFu;
            } else {
                const uint16_t* cl = cum16 + col + s * CPW;                               // cum[s], cum[s+1], cum[s+2]
                const uint32_t c0 = cl[0], c1 = cl[CPW], c2 = cl[2 * CPW];
                const bool adv = s < lastsym && help * c1 <= V;
                lt = adv ? c1 : c0;
                uint32_t nx = adv ? c2 : c1;
                s += adv ? 1u : 0u;
                if (adv && s < lastsym && help * nx <= V) {       // third or later symbol of its bucket: binary search (see above)
                    uint32_t lo = s + 1, hi = min(lastsym, (uint32_t)ll[((q >> kLutShift) + 2) * CPW]);
                    while (lo < hi) {
                        const uint32_t mid = (lo + hi + 1) >> 1;
                        if (help * (uint32_t)cum16[col + mid * CPW] <= V) lo = mid; else hi = mid - 1;
                    }
                    s = lo;
                    lt = cum16[col + s * CPW];
                    nx = cum16[col + (s + 1) * CPW];
                }
                sy = nx - lt;
            }
            const uint32_t tmp = help * lt;                   // decode_update
            X -= 2 * tmp;
            range = (s != lastsym) ? help * sy : range - tmp;
            return s;
        };
        uint32_t head = (4u - ((uint32_t)(unsigned long long)op & 3u)) & 3u;
        if (head > cnt) head = cnt;
        uint32_t i = 0;
#pragma unroll 1
        for (; i < head; i++) op[i] = (uint8_t)step();
        const uint32_t ngroups = (cnt - head) >> 2;
        uint32_t* __restrict__ o32 = reinterpret_cast<uint32_t*>(op + head);
#pragma unroll 1
        for (uint32_t gq = 0; gq < ngroups; gq++) {
            const uint32_t b0 = step(), b1 = step(), b2 = step(), b3 = step();
            o32[gq] = __byte_perm(__byte_perm(b0, b1, 0x0040), __byte_perm(b2, b3, 0x0040), 0x5410);
        }
        i = head + 4u * ngroups;
#pragma unroll 1
        for (; i < cnt; i++) op[i] = (uint8_t)step();
        };
        if (__all_sync(__activemask(), bs >= 256u)) symbol_loop(std::true_type{}); else symbol_loop(std::false_type{});
        d.X = X; d.range = range; d.ip = a - boff;
        if (d.ip > slen + 4) { bad = true; break; }           // consumed more bytes than the stream has: malformed
        n += bs;
        if (NSUB > 1) {
            if (s1 == bs) {                       // owner of the chunk's last symbol checks the end marker
                dec_renorm(d);
                if (!((d.X >> 1) < (d.range >> 1))) bad = true;
            }
            break;
        }
    }
    if (bad || n != clen) atomicExch(error, 3);
}

void range_decode_chunks(const uint8_t* blob, const unsigned long long* offs, const unsigned long long* lay_off,
                         const ChunkGeom& g, int nlay, uint8_t* sym, unsigned long long sym_layer_stride,
                         unsigned long long blob_len, int* error, cudaStream_t s)
{
    const unsigned int nsub = g.nseek + 1;        // make_geom grants 0, 1, 3 or 7 seek points
    const unsigned int cpw = 32 / nsub;
    dim3 grid((g.nchunks + cpw - 1) / cpw, nlay, 1);
    // lazy stream loads with one word of lookahead beat two 32-sector loads per symbol at every lane count
    // (512^3, 1 / 3 / 7 seek points: 12.4 vs 13.1, 5.01 vs 5.59, 2.67 vs 3.11 ms)
    const char* e = getenv("WRB_DEC_VARIANT");
    const int variant = (e && *e) ? atoi(e) : kDecVariantDefault;
    const bool lazy = variant & 1, pf = (variant & 2) != 0;
    const bool pair = decode_uses_pair_tables(nsub, (unsigned long long)grid.x * grid.y);
    const int smem = dec_table_bytes(pair) * (int)cpw;
#define WRB_DEC_LAUNCH(NS, V, P, F)                                                                                       \
    do {                                                                                                               \
        /* the fast form with one lane per chunk needs more than the 48 KB default; the attribute is per device */      \
        static DeviceOnce once;                                                                                        \
        once.run([] { cudaFuncSetAttribute(range_decode_kernel<NS, V, P, F>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024); }); \
        range_decode_kernel<NS, V, P, F><<<grid, 32, smem, s>>>(blob, offs, lay_off, g, sym, sym_layer_stride, blob_len, error); \
    } while (0)
#define WRB_DEC_VARIANTS(NS)                                                                                 \
    do {                                                                                                     \
        if (pair) { if (!lazy) WRB_DEC_LAUNCH(NS, false, true, false); else if (pf) WRB_DEC_LAUNCH(NS, true, true, true); else WRB_DEC_LAUNCH(NS, true, true, false); }        \
        else      { if (!lazy) WRB_DEC_LAUNCH(NS, false, false, false); else if (pf) WRB_DEC_LAUNCH(NS, true, false, true); else WRB_DEC_LAUNCH(NS, true, false, false); }      \
    } while (0)
    switch (nsub) {
    case 1: WRB_DEC_VARIANTS(1); break;
    case 2: WRB_DEC_VARIANTS(2); break;
    case 4: WRB_DEC_VARIANTS(4); break;
    default: WRB_DEC_VARIANTS(8); break;
    }
#undef WRB_DEC_VARIANTS
#undef WRB_DEC_LAUNCH
    note_launch(1);
}

void parse_container(const uint8_t* blob, const ChunkGeom& g, int chunked, int nlay, const unsigned long long* lay_off,
                     unsigned long long* offs, int* error, cudaStream_t s)
{
    parse_container_kernel<<<1, 1024, 0, s>>>(blob, g, chunked, nlay, lay_off, offs, error);
    note_launch(1);
}

}  // namespace wrb
