// compat.cpp -- the reference's library entry points on top of the wrb_* C ABI.
//
// Same names, argument meaning and error behaviour as reference src/core/wrappers.cpp:
//   encoding_wrap :228-452   decoding_wrap :456-527   setup_wr :531-541
//   encoding_wrap_f :545-563 decoding_wrap_f :567-580 setup_wr_f :584-594
// A program linked against libwaverange_b200.so instead of libwaverange.so runs on the GPU:
// the host buffers are copied to the device, compressed / decompressed there, and copied back.
// There is no CPU path: without a usable CUDA device these functions throw.
#include <cstdio>
#include <cstdlib>
#include <exception>
#include <iostream>
#include <stdexcept>
#include <string>
#include "../../include/waverange_b200.h"
#include "../../include/waverange.h"

namespace {
// One codec handle per (host thread, device): the reference's entry points are re-entrant on distinct buffers
// (SURVEY.md section 8b "callable from one host thread per GPU/stream"), so two threads -- on one GPU or on two --
// must neither serialise on a lock nor share scratch buffers.  The device is the calling thread's current CUDA device
// (cudaSetDevice before the call, as with any CUDA library); WRB_DEVICE overrides it for programs that never touch CUDA.
struct ThreadCodecs {
    struct Entry { int dev; wrb_codec* c; };
    Entry e[16];
    int n = 0;
    ~ThreadCodecs() { for (int i = 0; i < n; i++) wrb_destroy(e[i].c); }
};
thread_local ThreadCodecs t_codecs;

wrb_codec* codec()
{
    int dev = 0;
    const char* env = getenv("WRB_DEVICE");
    if (env && *env) dev = atoi(env);
    else if (wrb_current_device(&dev) != 0) dev = 0;
    for (int i = 0; i < t_codecs.n; i++) if (t_codecs.e[i].dev == dev) return t_codecs.e[i].c;
    wrb_codec* c = nullptr;
    int rc = wrb_create(&c, dev);
    if (rc != 0 || !c) {
        fprintf(stderr, "waverange_b200: no usable CUDA device (wrb_create -> %d); there is no CPU fallback\n", rc);
        throw std::runtime_error("waverange_b200: CUDA device unavailable");
    }
    if (t_codecs.n < 16) { t_codecs.e[t_codecs.n].dev = dev; t_codecs.e[t_codecs.n].c = c; t_codecs.n++; }
    else { wrb_destroy(t_codecs.e[0].c); t_codecs.e[0].dev = dev; t_codecs.e[0].c = c; }
    return c;
}
bool verbose() { const char* e = getenv("WRB_VERBOSE"); return e && *e && *e != '0'; }
}  // namespace

extern "C" void encoding_wrap(int nx, int ny, int nz, double* fld_1d, int wtflag, int mx, int my, int mz,
                              double* cutoffvec, double& tolabs, double& midval, double& halfspanval,
                              unsigned char& wlev, unsigned char& nlay, unsigned long int& ntot_enc, double* deps_vec,
                              double* minval_vec, unsigned long int* len_enc_vec, unsigned char* data_enc)
{
    wrb_codec* c = codec();
    // minimum cutoff (wrappers.cpp:292-293); mx*my*mz > 1 selects the spatially varying branch (:343-379)
    unsigned int mtot = (unsigned int)(mx * my * mz);
    double tolrel = cutoffvec[0];
    for (unsigned int k = 1; k < mtot; k++) if (cutoffvec[k] < tolrel) tolrel = cutoffvec[k];
    struct LocalGuard {                      // the handle outlives the call: never leave the local cutoff set
        wrb_codec* c; bool on;
        ~LocalGuard() { if (on) wrb_set_local_cutoff(c, 0, 0, 0, nullptr); }
    } guard{c, mtot > 1};
    if (mtot > 1 && wrb_set_local_cutoff(c, mx, my, mz, cutoffvec) != 0)
        throw std::runtime_error(std::string("waverange_b200: ") + wrb_last_error(c));
    if (verbose()) std::cout << "Wavelet decomposition..." << std::endl << "Range encoding..." << std::endl;
    unsigned char nlaymax; unsigned long cap;
    wrb_setup(nx, ny, nz, &nlaymax, &cap);
    wrb_header h;
    int rc = wrb_encode_host(c, fld_1d, WRB_F64, nx, ny, nz, wtflag, tolrel, &h, data_enc, cap);
    if (rc == WRB_E_OVERFLOW) {                                  // wrappers.cpp:422-426
        std::cout << "Error: encoded array is too large. Use larger SAFETY_BUFFER_FACTOR" << std::endl;
        throw std::exception();
    }
    if (rc != 0) {
        fprintf(stderr, "waverange_b200: encoding_wrap failed: %s\n", wrb_last_error(c));
        throw std::runtime_error(std::string("waverange_b200: ") + wrb_last_error(c));
    }
    tolabs = h.tolabs; midval = h.midval; halfspanval = h.halfspanval;
    wlev = h.wlev; nlay = h.nlay; ntot_enc = h.ntot_enc;
    for (int l = 0; l < h.nlay; l++) {
        deps_vec[l] = h.deps_vec[l]; minval_vec[l] = h.minval_vec[l]; len_enc_vec[l] = h.len_enc_vec[l];
        if (verbose()) std::cout << "ilay=" << l << " deps=" << h.deps_vec[l] << " len_out_q=" << h.len_enc_vec[l] << std::endl;
    }
}

extern "C" void decoding_wrap(int nx, int ny, int nz, double* fld_1d, double& tolabs, double& midval,
                              double& halfspanval, unsigned char& wlev, unsigned char& nlay,
                              unsigned long int& ntot_enc, double* deps_vec, double* minval_vec,
                              unsigned long int* len_enc_vec, unsigned char* data_enc)
{
    wrb_codec* c = codec();
    wrb_header h{};
    h.tolabs = tolabs; h.midval = midval; h.halfspanval = halfspanval;
    h.wlev = wlev; h.nlay = nlay; h.ntot_enc = ntot_enc;
    for (int l = 0; l < nlay && l < WRB_NLAYMAX; l++) {
        h.deps_vec[l] = deps_vec[l]; h.minval_vec[l] = minval_vec[l]; h.len_enc_vec[l] = len_enc_vec[l];
    }
    if (verbose()) std::cout << "Range decoding..." << std::endl << "Wavelet reconstruction..." << std::endl;
    int rc = wrb_decode_host(c, fld_1d, WRB_F64, nx, ny, nz, &h, data_enc);
    if (rc == WRB_E_FORMAT) {                                    // wrappers.cpp:168-171
        fprintf(stderr, "could not successfully open input data\n");
        exit(1);
    }
    if (rc != 0) {
        fprintf(stderr, "waverange_b200: decoding_wrap failed: %s\n", wrb_last_error(c));
        throw std::runtime_error(std::string("waverange_b200: ") + wrb_last_error(c));
    }
}

extern "C" void setup_wr(int nx, int ny, int nz, unsigned char& nlaymax, unsigned long int& ntot_enc_max)
{
    wrb_setup(nx, ny, nz, &nlaymax, &ntot_enc_max);
}

extern "C" void encoding_wrap_f(int* nx, int* ny, int* nz, double* fld, int* wtflag, double* tolrel, double& tolabs,
                                double& midval, double& halfspanval, unsigned char& wlev, unsigned char& nlay,
                                long int& ntot_enc_sg, double* deps_vec, double* minval_vec, long int* len_enc_vec_sg,
                                unsigned char* data_enc)
{
    unsigned long int ntot_enc = 0;
    unsigned long int len_enc_vec[WRB_NLAYMAX] = {0};
    double cut[1] = {*tolrel};
    encoding_wrap(*nx, *ny, *nz, fld, *wtflag, 1, 1, 1, cut, tolabs, midval, halfspanval, wlev, nlay, ntot_enc,
                  deps_vec, minval_vec, len_enc_vec, data_enc);
    ntot_enc_sg = (long int)ntot_enc;
    for (int j = 0; j < WRB_NLAYMAX; j++) len_enc_vec_sg[j] = (long int)len_enc_vec[j];   // all 8 slots (:561-562)
}

extern "C" void decoding_wrap_f(int* nx, int* ny, int* nz, double* fld, double& midval, double& halfspanval,
                                unsigned char& wlev, unsigned char& nlay, long int& ntot_enc_sg, double* deps_vec,
                                double* minval_vec, long int* len_enc_vec_sg, unsigned char* data_enc)
{
    double tolabs = 0;
    unsigned long int ntot_enc = (unsigned long int)ntot_enc_sg;
    unsigned long int len_enc_vec[WRB_NLAYMAX];
    for (int j = 0; j < WRB_NLAYMAX; j++) len_enc_vec[j] = (unsigned long int)len_enc_vec_sg[j];
    decoding_wrap(*nx, *ny, *nz, fld, tolabs, midval, halfspanval, wlev, nlay, ntot_enc, deps_vec, minval_vec,
                  len_enc_vec, data_enc);
}

extern "C" void setup_wr_f(int* nx, int* ny, int* nz, int& nlaymax, long int& ntot_enc_max)
{
    unsigned char n; unsigned long m;
    wrb_setup(*nx, *ny, *nz, &n, &m);
    nlaymax = n;
    ntot_enc_max = (long int)m;
}
