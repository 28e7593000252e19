/*
 * waverange_mssg.h -- the MSSG file layouts on top of the codec (C ABI).
 *
 * Replaces the file handling of the reference's MSSG front-end:
 *   encoder main            src/mssg/mssg_enc.cpp:56-615  (wrmssgenc)
 *   decoder main            src/mssg/mssg_dec.cpp:54-561  (wrmssgdec)
 *   restart namelist reader src/mssg/ctrl_aux.cpp:49-213  (read_control_file, PREFIX.nmlst)
 *   GrADS control reader    src/mssg/ctrl_aux.cpp:217-320 (read_control_file_grads, PREFIX.ctl)
 *   raw sub-domain I/O      src/mssg/ctrl_aux.cpp:324-472 (write_field_mssg, read_field_mssg)
 *   header record I/O       src/mssg/ctrl_aux.cpp:498-585 (write_header_mssg_enc, read_header_mssg_enc)
 *
 * Three file types (mssg_enc.cpp:239, 403-404):
 *   0  regular (GrADS) output: PREFIX.ctl names the data file and gives XDEF/YDEF/ZDEF/TDEF/UNDEF; every time
 *      instant is one field.  A field whose minimum lies below UNDEF·(1 ∓ 1e-4) is split into a two-valued mask
 *      (coded WITHOUT the wavelet transform at relative tolerance 0.126, record name "mask") and the field with
 *      the masked points padded by the mean of the others (mssg_enc.cpp:305-365).
 *   1  restart files PREFIX.p_0000 ... of all nprocx·nprocy sub-domains merged into one global field per record;
 *   2  one restart file PREFIX.p_<PROCID>, coded on its own.
 *      Types 1/2: PREFIX.nmlst gives the grid, the process grid and the record table; record 1 ("time") is not
 *      coded, its first 15 values go into the header as text (mssg_enc.cpp:483-502).
 * Outputs: PREFIX_h[PROCID]EXT (text header, layout of the reference byte for byte: 8 preamble lines, then per
 * record " -----", 1-based id, " Data set name = NAME", a reminder line, tolabs midval halfspanval with 19
 * significant digits, wlev nlay ntot_enc, and the three vectors) and PREFIX_f[PROCID]EXT (the encoded layers
 * of all records back to back).  With wrb_set_chunk_blocks(c, 0) both files are byte-identical to the stock
 * wrmssgenc's and the stock wrmssgdec reads them; by default each layer is a WRCK chunk container
 * (waverange_b200.h).  wrb_mssg_decode accepts both.
 *
 * Fields are compressed on the GPU through wrb_encode_host / wrb_decode_host: there is no CPU path.  The mask
 * split itself (a sequential sum whose order fixes the pad value's bits) is host work of the file layer.
 */
#ifndef WAVERANGE_MSSG_H
#define WAVERANGE_MSSG_H

#include "waverange_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

#define WRB_MSSG_NDSMAX 50          /* reference src/core/defs.h:52 NDSMAX */
#define WRB_MSSG_FILE_DIG 4         /* defs.h:54: digits of the .p_NNNN suffix */
#define WRB_MSSG_TIME_REC_LEN 15    /* defs.h:56 */
#define WRB_MSSG_MASK_TOLREL 0.126  /* defs.h:58 */
#define WRB_MSSG_MASK_THRESHOLD_ACC 1e-4 /* defs.h:60 */

#define WRB_MSSG_REGULAR 0
#define WRB_MSSG_RESTART_MERGED 1
#define WRB_MSSG_RESTART_DIVIDED 2

/* PREFIX.ctl (GrADS descriptor), ctrl_aux.cpp:217-320 */
typedef struct wrb_mssg_ctl {
    int nx, ny, nz, nt;
    double undef;
    char dset[256];
} wrb_mssg_ctl;

/* PREFIX.nmlst (restart namelist), ctrl_aux.cpp:49-213 */
typedef struct wrb_mssg_nmlst {
    int nx, ny, nz, nprocx, nprocy, ndset;
    char dset[WRB_MSSG_NDSMAX][256];   /* record names in file order (rec = 1 ... ndset) */
} wrb_mssg_nmlst;

/* ---- control and header files (no GPU needed) --------------------------------------------------- */
int wrb_mssg_read_ctl(const char* name, wrb_mssg_ctl* out);
int wrb_mssg_read_nmlst(const char* name, wrb_mssg_nmlst* out);
/* create/truncate the header and write its 8 preamble lines (mssg_enc.cpp:273-284 for type 0, :459-470 else) */
int wrb_mssg_header_begin(const char* header_name, const char* prefix, const char* ext, int filetype, int nbytes,
                          int endianflip, double tol_base);
/* the uncoded "time" record of a restart header (mssg_enc.cpp:472-486): id 1, name, 15 values */
int wrb_mssg_header_time(const char* header_name, const char* dsetname, const double* time_rec);
/* append one coded record; idset is 0-based, the file holds idset+1 (ctrl_aux.cpp:498-535) */
int wrb_mssg_header_append(const char* header_name, int idset, const char* dsetname, const wrb_header* hdr);
/* Read a whole header: skips the preamble (and, for types 1/2, reads the time record into time_rec[15]), then
 * reads records until the file ends.  ids[i] receives the 1-based id, names[i] the record name.
 * Returns WRB_E_FORMAT on a malformed file. */
int wrb_mssg_header_read(const char* header_name, int filetype, double* time_rec, int* nrec, int* ids,
                         char (*names)[256], wrb_header* hdrs, int max_recs);

/* ---- whole data sets -------------------------------------------------------------------------------- */
/* wrmssgenc: nbytes 4 or 8 = precision of the input files; procid: this sub-domain (type 2; names the file the
 * time record is read from for type 1, like the reference).  mssg_enc.cpp:236-602 */
int wrb_mssg_encode(wrb_codec* c, const char* prefix, const char* ext, int filetype, int nbytes, int endianflip,
                    double tol_base, int procid);
/* wrmssgdec: nbytes = precision of the files to write.  Copies the control file to OUT_PREFIX when the prefixes
 * differ.  mssg_dec.cpp:152-548 */
int wrb_mssg_decode(wrb_codec* c, const char* in_prefix, const char* ext, const char* out_prefix, int filetype,
                    int nbytes, int endianflip, int procid);
/* text of the last error of the MSSG layer (per thread) */
const char* wrb_mssg_last_error(void);

#ifdef __cplusplus
}
#endif
#endif /* WAVERANGE_MSSG_H */
