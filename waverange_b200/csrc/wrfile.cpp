// wrfile.cpp -- generic .wrh / .wrb file layer (host C++; the codec calls go to the GPU).
//
// Behavioural mirror of the reference's generic front-end (see include/waverange_files.h for the
// file:line map).  Own structure: whole records are moved with single reads/writes and permuted in
// memory (the reference moves one value per stream call), single-precision files stay single
// precision on the way to the device (the widening the reference does on the host,
// gen_aux.cpp:305-309, happens in the transform kernel), and the header is parsed line by line.
#include <cerrno>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <limits>
#include <sstream>
#include <string>
#include <vector>

#include "../../include/waverange_files.h"

namespace {

thread_local std::string g_err;

int fail(int code, const std::string& msg)
{
    g_err = msg;
    return code;
}

bool verbose()
{
    const char* e = getenv("WRB_VERBOSE");
    return e && *e && *e != '0';
}

// number of significant digits the reference prints doubles with (gen_aux.cpp:532)
constexpr int kDigits = std::numeric_limits<long double>::digits10 + 1;

size_t mark_bytes(int filetype) { return filetype == WRB_FILE_F77_4 ? 4 : (filetype == WRB_FILE_F77_8 ? 8 : 0); }

void reverse_bytes(unsigned char* p, size_t n)
{
    for (size_t i = 0; i < n / 2; i++) { unsigned char t = p[i]; p[i] = p[n - 1 - i]; p[n - 1 - i] = t; }
}

void flip_elements(unsigned char* p, size_t count, int nbytes)
{
    if (nbytes == 4) {
        for (size_t i = 0; i < count; i++, p += 4) { unsigned char a = p[0], b = p[1]; p[0] = p[3]; p[1] = p[2]; p[2] = b; p[3] = a; }
    } else {
        for (size_t i = 0; i < count; i++, p += 8) reverse_bytes(p, 8);
    }
}

// file order with idinv: ix slowest ... ih fastest (gen_aux.cpp:329-373); array order: ix fastest
template <class T>
void permute(const T* src, T* dst, const wrb_field_desc& d, bool file_to_array)
{
    const size_t nx = d.nx, ny = d.ny, nz = d.nz, nh = d.nh;
    for (size_t ix = 0; ix < nx; ix++)
        for (size_t iy = 0; iy < ny; iy++)
            for (size_t iz = 0; iz < nz; iz++) {
                const size_t f0 = ((ix * ny + iy) * nz + iz) * nh;
                for (size_t ih = 0; ih < nh; ih++) {
                    const size_t a = ix + nx * (iy + ny * (iz + nz * ih));
                    if (file_to_array) dst[a] = src[f0 + ih]; else dst[f0 + ih] = src[a];
                }
            }
}

size_t field_points(const wrb_field_desc& d) { return (size_t)d.nx * d.ny * d.nz * d.nh; }

bool desc_ok(const wrb_field_desc& d)
{
    return (d.nbytes == 4 || d.nbytes == 8) && d.nx > 0 && d.ny > 0 && d.nz > 0 && d.nh > 0;
}

// One raw record of the input file -> values in array order, still in the file's precision.
// reference gen_aux.cpp:229-396
int read_raw_field(FILE* f, int filetype, int endianflip, const wrb_field_desc& d, unsigned char recl[8],
                   std::vector<unsigned char>& out)
{
    const size_t mb = mark_bytes(filetype), n = field_points(d), bytes = n * (size_t)d.nbytes;
    if (mb) {
        unsigned char m[8] = {0};
        if (fread(m, 1, mb, f) != mb) return fail(WRB_E_FORMAT, "cannot read the record length");
        if (endianflip) reverse_bytes(m, mb);
        memcpy(recl, m, mb);
    }
    out.resize(bytes);
    if (fread(out.data(), 1, bytes, f) != bytes) return fail(WRB_E_FORMAT, "input file is shorter than the field");
    if (endianflip) flip_elements(out.data(), n, d.nbytes);
    if (d.idinv) {
        std::vector<unsigned char> tmp(bytes);
        if (d.nbytes == 4) permute((const float*)out.data(), (float*)tmp.data(), d, true);
        else permute((const double*)out.data(), (double*)tmp.data(), d, true);
        out.swap(tmp);
    }
    if (mb) {
        unsigned char m[8];
        if (fread(m, 1, mb, f) != mb) return fail(WRB_E_FORMAT, "cannot read the closing record length");
    }
    return 0;
}

// values in array order -> one raw record of the output file.  reference gen_aux.cpp:48-225
int write_raw_field(FILE* f, int filetype, int endianflip, const wrb_field_desc& d, const unsigned char recl[8],
                    std::vector<unsigned char>& vals)
{
    const size_t mb = mark_bytes(filetype), n = field_points(d), bytes = n * (size_t)d.nbytes;
    unsigned char m[8];
    memcpy(m, recl, 8);
    if (endianflip && mb) reverse_bytes(m, mb);
    if (mb && fwrite(m, 1, mb, f) != mb) return fail(WRB_E_FORMAT, "write failed");
    if (d.idinv) {
        std::vector<unsigned char> tmp(bytes);
        if (d.nbytes == 4) permute((const float*)vals.data(), (float*)tmp.data(), d, false);
        else permute((const double*)vals.data(), (double*)tmp.data(), d, false);
        vals.swap(tmp);
    }
    if (endianflip) flip_elements(vals.data(), n, d.nbytes);
    if (fwrite(vals.data(), 1, bytes, f) != bytes) return fail(WRB_E_FORMAT, "write failed");
    if (mb && fwrite(m, 1, mb, f) != mb) return fail(WRB_E_FORMAT, "write failed");
    return 0;
}

struct LineReader {
    std::ifstream in;
    std::string line;
    bool next() { return (bool)std::getline(in, line); }
};

bool parse_long(const std::string& s, long long& v)
{
    errno = 0;
    char* e = nullptr;
    v = strtoll(s.c_str(), &e, 10);
    return e != s.c_str() && errno == 0;
}

bool parse_doubles(const std::string& s, double* v, int n)
{
    const char* p = s.c_str();
    for (int i = 0; i < n; i++) {
        char* e = nullptr;
        v[i] = strtod(p, &e);
        if (e == p) return false;
        p = e;
    }
    return true;
}

}  // namespace

extern "C" {

const char* wrb_file_last_error(void) { return g_err.c_str(); }

int wrb_wrh_begin(const char* header_name, const char* encoded_name, int filetype, int endianflip, int nf)
{
    if (!header_name || !encoded_name) return fail(WRB_E_ARG, "bad argument");
    std::ofstream h(header_name, std::ios::out | std::ios::trunc);
    if (!h) return fail(WRB_E_ARG, std::string("cannot create ") + header_name);
    h << " ===== Header file for compressed data =====" << "\n"
      << " Coder version: " << WRB_CODER_VERSION << "\n"
      << " Encoded data file name: " << encoded_name << "\n"
      << " File type (0: Fortran sequential w 4-byte recl; 1: Fortran sequential w 8-byte recl; 2: C/C++): " << filetype << "\n"
      << (endianflip ? " Converted big endian to little endian or vice versa" : " No endian conversion") << "\n"
      << " Number of fields in the file, nf: " << nf << "\n";
    return h ? 0 : fail(WRB_E_ARG, "header write failed");
}

int wrb_wrh_append(const char* header_name, int idset, const wrb_field_record* rec)
{
    if (!header_name || !rec) return fail(WRB_E_ARG, "bad argument");
    std::ofstream h(header_name, std::ios::out | std::ios::app);
    if (!h) return fail(WRB_E_ARG, std::string("cannot open ") + header_name);
    const wrb_field_desc& d = rec->desc;
    const wrb_header& c = rec->hdr;
    const bool coded = d.icomp != 0, nontrivial = coded && c.ntot_enc > 0;
    h << " -----\n" << idset << "\n";
    h << " nbytes; recl; nx; ny; nz; nh; idinv; icomp;";
    if (coded) h << " tol_base; tolabs; midval; halfspanval; wlev; nlay; ntot_enc;";
    // the reminder line mentions the vectors whenever ntot_enc > 0 -- for a raw field that is whatever the
    // caller left in hdr.ntot_enc (the reference tests a variable still holding the previous field's value,
    // gen_aux.cpp:518); readers skip this line
    if (c.ntot_enc > 0) h << " deps_vec(1:nlay); minval_vec(1:nlay); len_enc_vec(1:nlay)";
    h << "\n" << d.nbytes << "\n";
    h << std::hex;
    for (int j = 0; j < 8; j++) h << (unsigned)rec->recl[j] << " ";
    h << std::dec << "\n";
    h << d.nx << "\n" << d.ny << "\n" << d.nz << "\n" << d.nh << "\n" << d.idinv << "\n" << d.icomp << "\n";
    if (coded) {
        h << std::setprecision(kDigits);
        h << d.tol_base << "\n" << c.tolabs << "\n" << c.midval << "\n" << c.halfspanval << "\n";
        h << (unsigned)c.wlev << "\n" << (unsigned)c.nlay << "\n" << c.ntot_enc << "\n";
        if (nontrivial) {
            for (int j = 0; j < c.nlay; j++) h << c.deps_vec[j] << " ";
            h << "\n";
            for (int j = 0; j < c.nlay; j++) h << c.minval_vec[j] << " ";
            h << "\n";
            for (int j = 0; j < c.nlay; j++) h << c.len_enc_vec[j] << " ";
            h << "\n";
        }
    }
    return h ? 0 : fail(WRB_E_ARG, "header write failed");
}

int wrb_wrh_read(const char* header_name, int* nf_out, wrb_field_record* recs, int max_recs)
{
    if (!header_name || !nf_out) return fail(WRB_E_ARG, "bad argument");
    LineReader r;
    r.in.open(header_name);
    if (!r.in) return fail(WRB_E_ARG, std::string("cannot open ") + header_name);
    for (int j = 0; j < 6; j++)
        if (!r.next()) return fail(WRB_E_FORMAT, "header file is too short");
    long long v = 0;
    if (r.line.size() < 34 || !parse_long(r.line.substr(34), v) || v < 0) return fail(WRB_E_FORMAT, "cannot read the number of fields");
    const int nf = (int)v;
    *nf_out = nf;
    for (int it = 0; it < nf; it++) {
        wrb_field_record rec;
        memset(&rec, 0, sizeof(rec));
        auto need = [&](const char* what) -> bool {
            if (r.next()) return true;
            fail(WRB_E_FORMAT, std::string("header file ends inside field ") + std::to_string(it) + " (" + what + ")");
            return false;
        };
        auto need_int = [&](const char* what, int& dst) -> bool {
            long long x;
            if (!need(what)) return false;
            if (!parse_long(r.line, x)) { fail(WRB_E_FORMAT, std::string("bad value for ") + what); return false; }
            dst = (int)x;
            return true;
        };
        int idset = -1;
        if (!need("separator") || !need_int("field index", idset)) return WRB_E_FORMAT;
        if (idset != it) {
            std::ostringstream m;
            m << "Encoding header file read error: reading field " << it << ", found field " << idset;
            return fail(WRB_E_FORMAT, m.str());
        }
        if (!need("reminder") || !need_int("nbytes", rec.desc.nbytes) || !need("recl")) return WRB_E_FORMAT;
        {
            std::istringstream hs(r.line);
            hs >> std::hex;
            for (int j = 0; j < 8; j++) { unsigned b = 0; hs >> b; rec.recl[j] = (unsigned char)b; }
        }
        if (!need_int("nx", rec.desc.nx) || !need_int("ny", rec.desc.ny) || !need_int("nz", rec.desc.nz) ||
            !need_int("nh", rec.desc.nh) || !need_int("idinv", rec.desc.idinv) || !need_int("icomp", rec.desc.icomp))
            return WRB_E_FORMAT;
        if (rec.desc.icomp > 0) {
            double four[4];
            for (int j = 0; j < 4; j++) {
                if (!need("tolerances")) return WRB_E_FORMAT;
                if (!parse_doubles(r.line, &four[j], 1)) return fail(WRB_E_FORMAT, "bad floating-point value in the header");
            }
            rec.desc.tol_base = four[0]; rec.hdr.tolabs = four[1]; rec.hdr.midval = four[2]; rec.hdr.halfspanval = four[3];
            int wlev = 0, nlay = 0;
            if (!need_int("wlev", wlev) || !need_int("nlay", nlay) || !need("ntot_enc")) return WRB_E_FORMAT;
            long long ne = 0;
            if (!parse_long(r.line, ne) || ne < 0 || nlay < 0 || nlay > WRB_NLAYMAX) return fail(WRB_E_FORMAT, "bad layer count or size");
            rec.hdr.wlev = (unsigned char)wlev; rec.hdr.nlay = (unsigned char)nlay; rec.hdr.ntot_enc = (unsigned long)ne;
            if (ne > 0) {
                if (!need("deps_vec") || !parse_doubles(r.line, rec.hdr.deps_vec, nlay)) return fail(WRB_E_FORMAT, "bad deps_vec");
                if (!need("minval_vec") || !parse_doubles(r.line, rec.hdr.minval_vec, nlay)) return fail(WRB_E_FORMAT, "bad minval_vec");
                if (!need("len_enc_vec")) return WRB_E_FORMAT;
                const char* p = r.line.c_str();
                for (int j = 0; j < nlay; j++) {
                    char* e = nullptr;
                    rec.hdr.len_enc_vec[j] = strtoul(p, &e, 10);
                    if (e == p) return fail(WRB_E_FORMAT, "bad len_enc_vec");
                    p = e;
                }
            }
        }
        if (recs && it < max_recs) recs[it] = rec;
    }
    return 0;
}

int wrb_file_encode(wrb_codec* c, const char* in_name, const char* encoded_name, const char* header_name, int filetype,
                    int endianflip, int nf, const wrb_field_desc* fields, const double* cutoff_all)
{
    if (!c || !in_name || !encoded_name || !header_name || nf < 0 || (nf > 0 && !fields)) return fail(WRB_E_ARG, "bad argument");
    if (filetype < 0 || filetype > 2) return fail(WRB_E_ARG, "Error: unknown file type");
    for (int it = 0; it < nf; it++)
        if (!desc_ok(fields[it])) return fail(WRB_E_ARG, "Generic input nbytes must be equal to 4 or 8 and all extents positive");
    int rc = wrb_wrh_begin(header_name, encoded_name, filetype, endianflip, nf);
    if (rc) return rc;
    FILE* fin = fopen(in_name, "rb");
    if (!fin) return fail(WRB_E_ARG, std::string("Cannot read from ") + in_name);
    FILE* fout = fopen(encoded_name, "wb");
    if (!fout) { fclose(fin); return fail(WRB_E_ARG, std::string("cannot create ") + encoded_name); }
    unsigned char recl[8] = {0};                 // carried from field to field like the reference's (gen_enc.cpp:108-110)
    unsigned long last_ntot_enc = 0;             // likewise (only shows in the reminder line of raw fields)
    std::vector<unsigned char> raw, enc;
    for (int it = 0; it < nf && rc == 0; it++) {
        const wrb_field_desc& d = fields[it];
        wrb_field_record rec;
        memset(&rec, 0, sizeof(rec));
        rec.desc = d;
        if ((rc = read_raw_field(fin, filetype, endianflip, d, recl, raw))) break;
        memcpy(rec.recl, recl, 8);
        const size_t n = field_points(d);
        if (verbose()) std::cout << "Field number " << it << ": nx=" << d.nx << " ny=" << d.ny << " nz=" << d.nz << " nh=" << d.nh
                                 << (d.idinv ? " and reordering" : "") << ", " << d.nbytes << "-byte data" << std::endl;
        if (d.icomp) {
            unsigned char nlaymax;
            unsigned long cap;
            const long long nzh = (long long)d.nz * d.nh;
            if (nzh > 0x7fffffffll) { rc = fail(WRB_E_ARG, "nz*nh does not fit an int"); break; }
            wrb_setup(d.nx, d.ny, (int)nzh, &nlaymax, &cap);
            // the reference sizes its buffer with setup_wr's bound (gen_enc.cpp:592); coded fields are far
            // smaller than that, so start at a fraction and grow on overflow
            unsigned long want = (unsigned long)(n * (size_t)d.nbytes) + (1ul << 20);
            if (want > cap) want = cap;
            for (;;) {
                enc.resize(want);
                rc = wrb_encode_host(c, raw.data(), d.nbytes == 4 ? WRB_F32 : WRB_F64, d.nx, d.ny, (int)nzh, 1, cutoff_all ? *cutoff_all : d.tol_base,
                                     &rec.hdr, enc.data(), want);
                if (rc == WRB_E_OVERFLOW && want < cap) { want = cap; continue; }
                break;
            }
            if (rc) { fail(rc, std::string("encoding failed: ") + wrb_last_error(c)); break; }
            last_ntot_enc = rec.hdr.ntot_enc;
            if ((rc = wrb_wrh_append(header_name, it, &rec))) break;
            if (rec.hdr.ntot_enc > 0 && fwrite(enc.data(), 1, rec.hdr.ntot_enc, fout) != rec.hdr.ntot_enc)
                rc = fail(WRB_E_ARG, "write failed");
            if (verbose()) std::cout << "  tolabs=" << rec.hdr.tolabs << " nlay=" << (unsigned)rec.hdr.nlay << " ntot_enc=" << rec.hdr.ntot_enc << std::endl;
        } else {
            rec.hdr.ntot_enc = last_ntot_enc;
            if ((rc = wrb_wrh_append(header_name, it, &rec))) break;
            if (fwrite(raw.data(), 1, raw.size(), fout) != raw.size()) rc = fail(WRB_E_ARG, "write failed");   // gen_aux.cpp:419-468
        }
    }
    fclose(fin);
    if (fclose(fout) != 0 && rc == 0) rc = fail(WRB_E_ARG, "write failed");
    return rc;
}

int wrb_file_decode(wrb_codec* c, const char* encoded_name, const char* header_name, const char* out_name, int filetype,
                    int endianflip)
{
    if (!c || !encoded_name || !header_name || !out_name) return fail(WRB_E_ARG, "bad argument");
    if (filetype < 0 || filetype > 2) return fail(WRB_E_ARG, "Error: unknown file type");
    int nf = 0;
    int rc = wrb_wrh_read(header_name, &nf, nullptr, 0);
    if (rc) return rc;
    std::vector<wrb_field_record> recs((size_t)(nf > 0 ? nf : 1));
    if ((rc = wrb_wrh_read(header_name, &nf, recs.data(), nf))) return rc;
    FILE* fin = fopen(encoded_name, "rb");
    if (!fin) return fail(WRB_E_ARG, std::string("cannot open ") + encoded_name);
    FILE* fout = fopen(out_name, "wb");
    if (!fout) { fclose(fin); return fail(WRB_E_ARG, std::string("cannot create ") + out_name); }
    std::vector<unsigned char> vals, enc;
    for (int it = 0; it < nf && rc == 0; it++) {
        const wrb_field_record& rec = recs[it];
        const wrb_field_desc& d = rec.desc;
        if (!desc_ok(d)) { rc = fail(WRB_E_FORMAT, "bad field description in the header"); break; }
        const size_t n = field_points(d);
        vals.resize(n * (size_t)d.nbytes);
        if (d.icomp) {
            const long long nzh = (long long)d.nz * d.nh;
            if (nzh > 0x7fffffffll) { rc = fail(WRB_E_FORMAT, "nz*nh in the header does not fit an int"); break; }
            enc.resize(rec.hdr.ntot_enc + 64);
            if (rec.hdr.ntot_enc > 0 && fread(enc.data(), 1, rec.hdr.ntot_enc, fin) != rec.hdr.ntot_enc) {
                rc = fail(WRB_E_FORMAT, "encoded file is shorter than the header says");
                break;
            }
            // a trivial field (ntot_enc == 0) decodes to midval everywhere (gen_dec.cpp:200, wrappers.cpp:462-469)
            rc = wrb_decode_host(c, vals.data(), d.nbytes == 4 ? WRB_F32 : WRB_F64, d.nx, d.ny, (int)nzh, &rec.hdr, enc.data());
            if (rc) { fail(rc, std::string("decoding failed: ") + wrb_last_error(c)); break; }
        } else {
            if (fread(vals.data(), 1, vals.size(), fin) != vals.size()) { rc = fail(WRB_E_FORMAT, "encoded file is shorter than the raw field"); break; }
        }
        if (verbose()) std::cout << "Field number " << it << " reconstructed" << std::endl;
        rc = write_raw_field(fout, filetype, endianflip, d, rec.recl, vals);
    }
    fclose(fin);
    if (fclose(fout) != 0 && rc == 0) rc = fail(WRB_E_ARG, "write failed");
    return rc;
}

}  // extern "C"
