/*
 * waverange_files.h -- the generic .wrh / .wrb file layer on top of the codec (C ABI).
 *
 * Replaces the file handling of the reference's generic front-end:
 *   encoder main          src/generic/gen_enc.cpp:56-656  (wrenc)
 *   decoder main          src/generic/gen_dec.cpp:53-268  (wrdec)
 *   raw field reader      src/generic/gen_aux.cpp:229-396 (read_field_gen: Fortran record marks, endian
 *                         flip, reversed index order, nh folded into z)
 *   raw field writer      src/generic/gen_aux.cpp:48-225  (write_field_gen)
 *   header record writer  src/generic/gen_aux.cpp:505-551 (write_header_gen_enc)
 *   header record reader  src/generic/gen_aux.cpp:554-644 (read_header_gen_enc)
 * File layouts are the reference's, byte for byte:
 *   .wrh  text: 6 preamble lines (gen_enc.cpp:512-518; the decoder skips 5 and reads nf at column 34,
 *         gen_dec.cpp:163-168), then per field " -----", the field index, a reminder line, nbytes, the
 *         8 hex bytes of the Fortran record length, nx ny nz nh idinv icomp, and for compressed fields
 *         tol_base tolabs midval halfspanval (19 significant digits) wlev nlay ntot_enc and the three
 *         vectors deps_vec, minval_vec, len_enc_vec;
 *   .wrb  the encoded layers of every compressed field back to back (ntot_enc bytes each), raw C-order
 *         values for uncompressed fields.
 * With chunk_blocks >= 1 (default) each layer in the .wrb is a WRCK chunk container (waverange_b200.h)
 * that only this library reads; with wrb_set_chunk_blocks(c, 0) the files are what the stock wrenc
 * writes and the stock wrdec reads them.  wrb_file_decode accepts both.
 *
 * Fields are compressed on the GPU through wrb_encode_host / wrb_decode_host: there is no CPU path.
 */
#ifndef WAVERANGE_FILES_H
#define WAVERANGE_FILES_H

#include "waverange_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

#define WRB_CODER_VERSION 31503   /* reference src/core/defs.h:34 CODER_VERSION */

/* file types (gen_enc.cpp:361): Fortran sequential with 4- or 8-byte record marks, or plain C order */
#define WRB_FILE_F77_4 0
#define WRB_FILE_F77_8 1
#define WRB_FILE_C 2

/* per-field parameters of the encoder (gen_enc.cpp:84-88, the `inmeta` blocks :203-255) */
typedef struct wrb_field_desc {
    int nbytes;          /* 4: single, 8: double */
    int nx, ny, nz, nh;  /* nh: higher (slowest) dimensions, folded into z (gen_enc.cpp:559) */
    int idinv;           /* 1: the file stores the indices in reverse order (gen_aux.cpp:329-373) */
    int icomp;           /* 0: store raw, 1: compress */
    double tol_base;     /* relative tolerance */
} wrb_field_desc;

/* one field record of a .wrh file */
typedef struct wrb_field_record {
    wrb_field_desc desc;
    unsigned char recl[8];   /* Fortran record length bytes as read from the input file */
    wrb_header hdr;          /* valid when desc.icomp != 0 */
} wrb_field_record;

/* ---- header file (no GPU needed) -------------------------------------------------------------- */
/* create/truncate the .wrh and write the preamble (gen_enc.cpp:508-519) */
int wrb_wrh_begin(const char* header_name, const char* encoded_name, int filetype, int endianflip, int nf);
/* append one field record (gen_aux.cpp:505-551) */
int wrb_wrh_append(const char* header_name, int idset, const wrb_field_record* rec);
/* read a whole .wrh: *nf receives the number of fields, up to max_recs records are stored.
 * Returns WRB_E_FORMAT on a malformed file (wrong field index: gen_aux.cpp:563-569). */
int wrb_wrh_read(const char* header_name, int* nf, wrb_field_record* recs, int max_recs);

/* ---- whole files --------------------------------------------------------------------------------- */
/* wrenc: compress the nf fields of in_name into encoded_name (+ header_name).  gen_enc.cpp:521-640
 * cutoff_all == NULL: every field is coded with its own tol_base.
 * cutoff_all != NULL: every field is coded with *cutoff_all while its own tol_base goes into the header -- what
 *   the reference's wrenc does: it fills the cutoff vector once, before the field loop, from the tolerance it
 *   parsed last (gen_enc.cpp:497-500), so with different per-field tolerances in `inmeta` or in the interactive
 *   dialogue all fields get the last one.  The wrenc front-end of this library passes that value. */
int wrb_file_encode(wrb_codec* c, const char* in_name, const char* encoded_name, const char* header_name, int filetype,
                    int endianflip, int nf, const wrb_field_desc* fields, const double* cutoff_all);
/* wrdec: reconstruct out_name from encoded_name + header_name.  gen_dec.cpp:146-262 */
int wrb_file_decode(wrb_codec* c, const char* encoded_name, const char* header_name, const char* out_name, int filetype,
                    int endianflip);
/* text of the last error of the file layer (per thread) */
const char* wrb_file_last_error(void);

#ifdef __cplusplus
}
#endif
#endif /* WAVERANGE_FILES_H */
