/*
 * waverange_b200.h -- C ABI of the B200-native WaveRange compress/decompress hot path.
 *
 * One shared library, libwaverange_b200.so, exports two groups of entry points:
 *
 *  (1) the reference's own library interface, unchanged, so that the reference front-ends
 *      (src/generic, src/flusi, src/mssg) and user Fortran/C++ codes link against it instead of
 *      libwaverange.{a,so}:  encoding_wrap / decoding_wrap / setup_wr and the Fortran twins
 *      (declared in include/waverange.h; reference src/core/wrappers.h:53,70,75,95,111,119);
 *
 *  (2) the wrb_* functions below: a codec handle that owns device scratch, device-pointer
 *      encode/decode (what the benchmarks time), host-pointer encode/decode (what group (1) is
 *      built on) and stage-level entry points used by the parity tests.
 *
 * Plain C types only; every pointer named d_* is a CUDA device pointer, every other pointer is
 * host memory.  All functions return 0 on success, a negative WRB_E_* code otherwise;
 * wrb_last_error() gives the text.  There is no CPU fallback: without a CUDA device every call
 * that needs one fails with WRB_E_CUDA.
 */
#ifndef WAVERANGE_B200_H
#define WAVERANGE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define WRB_NLAYMAX 8          /* reference src/core/defs.h:38  NLAYMAX   */
#define WRB_BLOCKSIZE 60000    /* reference src/core/defs.h:36  BLOCKSIZE */
#define WRB_WAV_LVL 4          /* reference src/core/defs.h:50  WAV_LVL   */

#define WRB_F64 0
#define WRB_F32 1

#define WRB_E_ARG (-1)
#define WRB_E_CUDA (-2)
#define WRB_E_OVERFLOW (-3)    /* encoded data does not fit (reference wrappers.cpp:422-426) */
#define WRB_E_FORMAT (-4)      /* malformed container / stream */
#define WRB_E_NOMEM (-5)

/* Coding metadata of one field: exactly the scalar/vector outputs of encoding_wrap()
 * (reference src/core/wrappers.h:35-52). */
typedef struct wrb_header {
    double tolabs, midval, halfspanval;
    unsigned char wlev, nlay;
    unsigned long ntot_enc;
    double deps_vec[WRB_NLAYMAX];
    double minval_vec[WRB_NLAYMAX];
    unsigned long len_enc_vec[WRB_NLAYMAX];
} wrb_header;

typedef struct wrb_codec wrb_codec;

/* ---- handle -------------------------------------------------------------------------------- */
int wrb_create(wrb_codec** out, int device);
void wrb_destroy(wrb_codec* c);
const char* wrb_last_error(const wrb_codec* c);
/* CUDA stream (cudaStream_t) all work of this handle is launched on; NULL = legacy default stream */
int wrb_set_stream(wrb_codec* c, void* cuda_stream);
/* Chunking of each layer's symbol sequence for the parallel range coder.
 *   blocks >= 1: chunk length = blocks*60000 - 1 symbols (default 1).  Each layer is stored as a
 *                "WRCK" container: 32-byte header, u32 byte length per chunk, seek table, then the
 *                chunk streams; every chunk stream is byte-identical to the reference's range_encode()
 *                (wrappers.cpp:68-149) of that symbol sub-array.
 *   blocks == 0: one stream per layer, byte-identical to the reference's encoding_wrap() output
 *                (readable by the stock wrdec); serial on the GPU, for interoperability only. */
int wrb_set_chunk_blocks(wrb_codec* c, int blocks);
/* Seek points per chunk: the encoder stores (low, range, stream position) at n interior symbol positions of
 * every single-block chunk, 10 bytes each, in the container's seek table; the decoder then runs n+1 lanes per
 * chunk.  Chunk streams are unaffected.  n = 0, 1, 3 or 7 (other values round down); n = -1 (default): the
 * encoder keeps as many points (7, 3, 1 or 0) as fit into 0.85 % of the coded bytes for the chunk tables, so
 * that the container stays within 1 % of the reference's layer streams whatever the data.
 * WRB_SEEK_POINTS in the environment sets the handle's initial value. */
int wrb_set_seek_points(wrb_codec* c, int n);
/* Local (spatially varying) cutoff: the mx*my*mz > 1 branch of encoding_wrap() (reference wrappers.cpp:343-379,
 * lcl_prec :55-64).  cutoffvec holds mx*my*mz relative tolerances, x fastest.  While set, wrb_encode_* ignore their
 * tolrel argument: tolabs comes from the minimum of cutoffvec (:292-299) and every layer codes symbol 0 where its
 * residual span is below the point's precision -- exactly what the reference does, including the fact that its
 * ind_p2w_3d() reports the full level for every point, so the per-point precision only takes effect with
 * wtflag == 0.  mx*my*mz <= 1 or cutoffvec == NULL switches it off.  Not available in z-slab mode. */
int wrb_set_local_cutoff(wrb_codec* c, int mx, int my, int mz, const double* cutoffvec);
/* Number of kernels launched by this handle since creation (bench.py's gpu_launches). */
unsigned long long wrb_launch_count(const wrb_codec* c);
/* Release cached device scratch (it is otherwise kept and grown on demand). */
int wrb_trim(wrb_codec* c);
/* The calling thread's current CUDA device (cudaGetDevice); the drop-in entry points of waverange.h create their
 * per-thread handle on it. */
int wrb_current_device(int* device);
/* The encoder launches only as many layer passes as the tolerance can need and repeats the call with all 8 when that
 * bound did not hold (the stream is the same either way, the call takes twice as long): how often that happened on
 * this handle since creation.  Expected: 0. */
unsigned long long wrb_layer_guess_misses(const wrb_codec* c);

/* ---- sizes: replaces setup_wr() (reference wrappers.cpp:531-541) ------------------------------ */
void wrb_setup(int nx, int ny, int nz, unsigned char* nlaymax, unsigned long* ntot_enc_max);

/* ---- device-resident path (inputs/outputs already in HBM) ----------------------------------- */
/* Compress: replaces encoding_wrap() (reference wrappers.cpp:228-452) for a field resident on the
 * device as f64 or f32 (f32 is widened on the fly exactly like gen_aux.cpp:305-309 does on the
 * host).  d_field is NOT modified.  d_data_enc (capacity cap bytes, >= wrb_setup's ntot_enc_max
 * recommended) receives the encoded layers back to back; hdr (host) receives the metadata.
 * The call returns after the stream has been synchronised. */
int wrb_encode_device(wrb_codec* c, const void* d_field, int dtype, int nx, int ny, int nz, int wtflag,
                      double tolrel, wrb_header* hdr, unsigned char* d_data_enc, unsigned long cap);
/* Decompress: replaces decoding_wrap() (reference wrappers.cpp:456-527).  d_data_enc holds
 * hdr->ntot_enc bytes followed by at least 32 readable bytes, and is 8-byte aligned.  d_field_out is f64 or f32. */
int wrb_decode_device(wrb_codec* c, void* d_field_out, int dtype, int nx, int ny, int nz, const wrb_header* hdr,
                      const unsigned char* d_data_enc);

/* ---- z-slab partition of one large field across the GPUs of a box -------- */
/* Rank r of n owns the planes [z0, z0+nzl) of an (nx, ny, nz) field (nz % 16 == 0, z0 and nzl multiples
 * of 32).  x/y lifting is slab-local; the z lifting of every level reads 4 planes from below and 3
 * from above (inverse: 2 + 2 per band) that come from the z-neighbours, and the field / coefficient /
 * per-layer residual extrema are reduced over all ranks (the codec packs them so that one MIN reduction
 * of two int64 values is enough).
 *
 * The collectives are issued by the library itself over NCCL (wrb_set_comm below: grouped ncclSend /
 * ncclRecv of whole planes, ncclAllReduce), so a C, C++ or Fortran caller needs nothing else; or they are
 * injected as callbacks (wrb_set_slab: gloo in the CPU tests, several emulated ranks on one GPU).
 *   halo  : neighbour exchange along z.  Send down_bytes from send_down to rank-1 and receive as many into
 *           recv_hi from rank+1; send up_bytes from send_up to rank+1 and receive as many into recv_lo from
 *           rank-1; nothing at the domain ends.  All pointers are device pointers.
 *   reduce: in-place global MIN over `count` signed 64-bit integers on the device (one all-reduce); a negative
 *           count asks for the global SUM over -count values.
 * Both must be ordered with the codec's stream. */
typedef int (*wrb_halo_fn)(void* user, const void* send_down, const void* send_up, void* recv_lo, void* recv_hi,
                           unsigned long long down_bytes, unsigned long long up_bytes);
typedef int (*wrb_reduce_fn)(void* user, long long* d_buf, int count);
int wrb_set_slab(wrb_codec* c, int rank, int nranks, wrb_halo_fn halo, wrb_reduce_fn reduce, void* user);
/* NCCL transport inside the library (libnccl.so.2 is loaded at run time; nothing to link).  Rank 0 obtains an id,
 * the 128 bytes reach the other ranks by any means (MPI_Bcast, a file, torch.distributed ...), and every rank -- one
 * process per GPU, its device current -- calls wrb_set_comm: ncclCommInitRank, halo exchange = grouped ncclSend /
 * ncclRecv, reductions = ncclAllReduce, all on the codec's stream.  It also switches on the GLOBAL symbol order:
 * before coding, the ranks exchange the 1-byte symbols so that rank r codes the chunks [r*nchunks/n, (r+1)*nchunks/n)
 * of the symbol sequence of the WHOLE field in the reference's order (wrappers.cpp:384-412 on the array that
 * waveletcdf97_3d.c:128-135,256-263 de-interleaves) -- every chunk stream is then byte for byte the one a single
 * GPU, and the reference's range_encode on that sub-array, produce.  The exchange reads the peers' symbol planes
 * directly over NVLink (CUDA IPC mappings, set up once per geometry through an ncclAllGather of the handles).
 * Each rank's data_enc is a self-contained WRCK container of its run of chunks; hdr->nlay / deps_vec / minval_vec
 * are identical on all ranks.  Requires equal slabs in rank order (z0 == rank * nzl). */
int wrb_comm_unique_id(unsigned char id[128]);
int wrb_set_comm(wrb_codec* c, int rank, int nranks, const unsigned char id[128]);
/* Several ranks emulated in ONE process on one device (tests): with the callbacks of wrb_set_slab, the codecs of all
 * ranks in rank order; switches on the global symbol order (the peers' windows are then plain pointers). */
int wrb_set_slab_peers(wrb_codec* c, wrb_codec* const* peers, int n);
/* 0: every rank codes its own coefficients in rank-local order (no symbol exchange); 1: the global order */
int wrb_set_slab_order(wrb_codec* c, int global);
/* NCCL transport counters since wrb_set_comm: halo bytes received, neighbour exchanges, all-reduces */
int wrb_comm_counters(const wrb_codec* c, unsigned long long out[3]);
/* The partition's index map (host arithmetic, for tests and tools): global wavelet-space plane of local plane p of
 * `rank` for an (x, y) position that leaves the low box at level `reg` (levels + 1: the coarsest box); and the chunks
 * [c0, c1) of the global sequence that `rank` codes. */
int wrb_slab_order_plane(int nx, int ny, int nz, int nranks, int levels, int rank, int p, int reg);
int wrb_slab_chunk_range(int nx, int ny, int nz, int nranks, unsigned long chunk_len, int rank, unsigned long* c0, unsigned long* c1);
int wrb_encode_slab_device(wrb_codec* c, const void* d_field_slab, int dtype, int nx, int ny, int nz, int z0, int nzl,
                           int wtflag, double tolrel, wrb_header* hdr, unsigned char* d_data_enc, unsigned long cap);
int wrb_decode_slab_device(wrb_codec* c, void* d_field_slab_out, int dtype, int nx, int ny, int nz, int z0, int nzl,
                           const wrb_header* hdr, const unsigned char* d_data_enc);
/* stage-level variant of wrb_quantise_device for a slab (rank-local coefficient / symbol order) */
int wrb_quantise_slab_device(wrb_codec* c, const void* d_field_slab, int dtype, int nx, int ny, int nz, int z0, int nzl,
                             int wtflag, double tolrel, wrb_header* hdr, double* d_coef, unsigned char* d_sym);

/* Dequantise + inverse transform from given symbols instead of coded data: d_sym holds hdr->nlay planes of ntot
 * (slab: nx*ny*nzl) bytes in array order, what wrb_quantise_device / wrb_quantise_slab_device return.  Together with
 * wrb_range_encode_device / wrb_range_decode_device these let a caller put its own exchange between the quantiser and
 * the coder -- waverange_b200/slab.py uses them to code a z-slab partitioned field in the GLOBAL symbol order, so that
 * the chunk streams are those of the single-GPU run. */
int wrb_decode_symbols_device(wrb_codec* c, void* d_field_out, int dtype, int nx, int ny, int nz, const wrb_header* hdr,
                              const unsigned char* d_sym);
int wrb_decode_slab_symbols_device(wrb_codec* c, void* d_field_slab_out, int dtype, int nx, int ny, int nz, int z0, int nzl,
                                   const wrb_header* hdr, const unsigned char* d_sym);

/* ---- host-buffer path (copies inside) ------------------------------------------------------- */
int wrb_encode_host(wrb_codec* c, const void* field, int dtype, int nx, int ny, int nz, int wtflag, double tolrel,
                    wrb_header* hdr, unsigned char* data_enc, unsigned long cap);
int wrb_decode_host(wrb_codec* c, void* field_out, int dtype, int nx, int ny, int nz, const wrb_header* hdr,
                    const unsigned char* data_enc);

/* ---- stage-level entry points (parity tests, profiling) -------------------------------------- */
/* In-place 3-D CDF 9/7 transform of a device f64 array, sign of lvl selects direction:
 * replaces waveletcdf97_3d() (reference src/waveletcdf97_3d/waveletcdf97_3d.h:39). */
int wrb_wavelet3d_device(wrb_codec* c, double* d_x, int nx, int ny, int nz, int lvl);
/* Transform + quantise only (no coder).  Outputs, all optional (NULL to skip):
 *   d_coef  ntot doubles: wavelet coefficients in array order;
 *   d_sym   WRB_NLAYMAX*ntot bytes, layer-major, array order: the symbols the reference holds in
 *           fld_q for each layer (wrappers.cpp:384-389);
 *   hdr     tolabs, midval, halfspanval, wlev, nlay, deps_vec, minval_vec (lengths are zero). */
int wrb_quantise_device(wrb_codec* c, const void* d_field, int dtype, int nx, int ny, int nz, int wtflag,
                        double tolrel, wrb_header* hdr, double* d_coef, unsigned char* d_sym);
/* Code n symbols as ceil(n/chunk_len) independent streams (chunk_len == 0: one stream).
 * d_out receives the streams back to back, lens (host, u64 per chunk) their byte lengths. */
int wrb_range_encode_device(wrb_codec* c, const unsigned char* d_sym, unsigned long n, unsigned long chunk_len,
                            unsigned char* d_out, unsigned long cap, unsigned long* lens, unsigned long* total);
/* Inverse of the above: streams back to back in d_in, lens from the encoder. */
int wrb_range_decode_device(wrb_codec* c, const unsigned char* d_in, const unsigned long* lens, unsigned long n,
                            unsigned long chunk_len, unsigned char* d_sym);
/* Physical -> wavelet-space index map: replaces ind_p2w_3d()
 * (reference src/waveletcdf97_3d/waveletcdf97_3d.c:473-553).  Host arithmetic. */
void wrb_ind_p2w_3d(int lvlin, int n1, int n2, int n3, int i1, int i2, int i3, int* lvl, int* o1, int* o2, int* o3);

/* Device time, in milliseconds, of the stages of the last wrb_encode_device / wrb_decode_device
 * call when timing is enabled with wrb_set_timing(c, 1) (CUDA events on the handle's stream).
 * encode: [0] transform, [1] quantise layers, [2] range coder, [3] container assembly
 * decode: [0] parse, [1] range decoder, [2] dequantise, [3] inverse transform */
int wrb_set_timing(wrb_codec* c, int on);
int wrb_last_stage_ms(const wrb_codec* c, float ms[4]);

#ifdef __cplusplus
}
#endif
#endif /* WAVERANGE_B200_H */
