/*
 * waverange.h -- the reference's public library interface, re-declared for the B200 build.
 *
 * libwaverange_b200.so exports these six symbols with the reference's exact ABI, so a program
 * written against the reference (src/core/wrappers.h) links and runs unchanged:
 *
 *   encoding_wrap    reference src/core/wrappers.h:53   (impl wrappers.cpp:228-452)
 *   decoding_wrap    reference src/core/wrappers.h:70   (impl wrappers.cpp:456-527)
 *   setup_wr         reference src/core/wrappers.h:75   (impl wrappers.cpp:531-541)
 *   encoding_wrap_f  reference src/core/wrappers.h:95   (impl wrappers.cpp:545-563)
 *   decoding_wrap_f  reference src/core/wrappers.h:111  (impl wrappers.cpp:567-580)
 *   setup_wr_f       reference src/core/wrappers.h:119  (impl wrappers.cpp:584-594)
 *
 * The reference declares its by-reference parameters as C++ references; at the ABI level those
 * are pointers, which is how a C caller sees them below.
 *
 * Arrays are x-fastest ("Fortran order"), fld_1d holds nx*ny*nz doubles, data_enc must hold
 * ntot_enc_max bytes from setup_wr() on encode.  Differences from the reference, all at the
 * container level and documented in INTEGRATION.md:
 *   - by default each layer inside data_enc is a "WRCK" chunk container (many independent
 *     reference-format streams) instead of one stream; WRB_CHUNK_BLOCKS=0 in the environment
 *     selects the reference's single-stream layout;
 *   - fld_1d is left untouched by encoding_wrap (the reference overwrites it with the residual);
 *   - progress text goes to stdout only when WRB_VERBOSE=1.
 */
#ifndef WAVERANGE_COMPAT_H
#define WAVERANGE_COMPAT_H

#ifdef __cplusplus
extern "C" {
void encoding_wrap(int nx, int ny, int nz, double* fld_1d, int wtflag, int mx, int my, int mz, double* cutoffvec,
                   double& tolabs, double& midval, double& halfspanval, unsigned char& wlev, unsigned char& nlay,
                   unsigned long int& ntot_enc, double* deps_vec, double* minval_vec, unsigned long int* len_enc_vec,
                   unsigned char* data_enc);
void decoding_wrap(int nx, int ny, int nz, double* fld_1d, double& tolabs, double& midval, double& halfspanval,
                   unsigned char& wlev, unsigned char& nlay, unsigned long int& ntot_enc, double* deps_vec,
                   double* minval_vec, unsigned long int* len_enc_vec, unsigned char* data_enc);
void setup_wr(int nx, int ny, int nz, unsigned char& nlaymax, unsigned long int& ntot_enc_max);
void encoding_wrap_f(int* nx, int* ny, int* nz, double* fld, int* wtflag, double* tolrel, double& tolabs,
                     double& midval, double& halfspanval, unsigned char& wlev, unsigned char& nlay,
                     long int& ntot_enc, double* deps_vec, double* minval_vec, long int* len_enc_vec,
                     unsigned char* data_enc);
void decoding_wrap_f(int* nx, int* ny, int* nz, double* fld, double& midval, double& halfspanval,
                     unsigned char& wlev, unsigned char& nlay, long int& ntot_enc, double* deps_vec,
                     double* minval_vec, long int* len_enc_vec, unsigned char* data_enc);
void setup_wr_f(int* nx, int* ny, int* nz, int& nlaymax, long int& ntot_enc_max);
}
#else
void encoding_wrap(int nx, int ny, int nz, double* fld_1d, int wtflag, int mx, int my, int mz, double* cutoffvec,
                   double* tolabs, double* midval, double* halfspanval, unsigned char* wlev, unsigned char* nlay,
                   unsigned long int* ntot_enc, double* deps_vec, double* minval_vec, unsigned long int* len_enc_vec,
                   unsigned char* data_enc);
void decoding_wrap(int nx, int ny, int nz, double* fld_1d, double* tolabs, double* midval, double* halfspanval,
                   unsigned char* wlev, unsigned char* nlay, unsigned long int* ntot_enc, double* deps_vec,
                   double* minval_vec, unsigned long int* len_enc_vec, unsigned char* data_enc);
void setup_wr(int nx, int ny, int nz, unsigned char* nlaymax, unsigned long int* ntot_enc_max);
void encoding_wrap_f(int* nx, int* ny, int* nz, double* fld, int* wtflag, double* tolrel, double* tolabs,
                     double* midval, double* halfspanval, unsigned char* wlev, unsigned char* nlay, long int* ntot_enc,
                     double* deps_vec, double* minval_vec, long int* len_enc_vec, unsigned char* data_enc);
void decoding_wrap_f(int* nx, int* ny, int* nz, double* fld, double* midval, double* halfspanval,
                     unsigned char* wlev, unsigned char* nlay, long int* ntot_enc, double* deps_vec,
                     double* minval_vec, long int* len_enc_vec, unsigned char* data_enc);
void setup_wr_f(int* nx, int* ny, int* nz, int* nlaymax, long int* ntot_enc_max);
#endif
#endif /* WAVERANGE_COMPAT_H */
