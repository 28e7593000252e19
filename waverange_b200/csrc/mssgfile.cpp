// mssgfile.cpp -- MSSG file layouts (host C++; the codec calls go to the GPU).
//
// Behavioural mirror of the reference's MSSG front-end (see include/waverange_mssg.h for the file:line
// map).  Own structure: the two control-file readers share one tokenizer that reports (token, mode) events,
// sub-domains are moved with one read/write per x-row block instead of one stream call per value,
// single-precision data stays single precision on the way to and from the device whenever no mask is
// involved (the widening the reference does on the host, ctrl_aux.cpp:452-455, happens in the transform
// kernel), and the header is parsed line by line into records before the decode loop walks them.
#include <cerrno>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <functional>
#include <iomanip>
#include <iostream>
#include <limits>
#include <sstream>
#include <string>
#include <vector>

#include "../../include/waverange_mssg.h"

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

constexpr int kDigits = std::numeric_limits<long double>::digits10 + 1;   // ctrl_aux.cpp:514

// ---- control files ------------------------------------------------------------------------------------
// Both readers of the reference are character state machines with three modes: skipping, collecting a
// name, collecting a value.  A token is reported when a separator follows at least one collected character;
// characters still pending at the end of the file are dropped (no separator follows them).
enum Mode { kSkip = 0, kName = 1, kValue = 2 };

struct Pair { std::string name, value; };

bool one_of(const std::string& s, std::initializer_list<const char*> names)
{
    for (const char* n : names) if (s == n) return true;
    return false;
}

// restart namelist (ctrl_aux.cpp:66-135): separators newline & blank ' , ; '=' switches to value mode and drops
// whatever was being collected (so "nx=4" without a blank before '=' is not seen, as in the reference); a name
// token is kept only if it is one of the known parameters, and only the next value token is paired with it.
int scan_nmlst(std::istream& in, std::vector<Pair>& out)
{
    Mode mode = kSkip;
    std::string tok;
    bool want_value = false;
    char ch;
    while (in.get(ch)) {
        if (ch == '=') { mode = kValue; tok.clear(); continue; }
        const bool sep = ch == '\n' || ch == '&' || ch == ' ' || ch == '\'' || ch == ',';
        if (!sep) {
            if (mode != kValue) mode = kName;
            if (tok.size() >= 255) return fail(WRB_E_FORMAT, "token longer than 255 characters in the namelist");
            tok.push_back(ch);
            continue;
        }
        if (tok.empty()) continue;
        if (mode == kName) {
            if (one_of(tok, {"nx", "ny", "nr", "npg", "i_over", "j_over", "nproc", "dim_size", "var", "rec"})) {
                if (want_value) out.back().name = tok;       // a second name before any value replaces the first
                else out.push_back({tok, ""});
                want_value = true;
            }
        } else if (mode == kValue && want_value) {
            out.back().value = tok;
            want_value = false;
        }
        mode = kSkip;
        tok.clear();
    }
    if (want_value) out.pop_back();                           // a name without a value never counts (npar not advanced)
    return 0;
}

// GrADS descriptor (ctrl_aux.cpp:236-292): separators newline ^ blank; a new line starts in name mode; a known
// name switches to value mode, the value switches to skipping the rest of the line.
void scan_ctl(std::istream& in, std::vector<Pair>& out)
{
    Mode mode = kName;
    std::string tok;
    bool open = false;
    char ch;
    while (in.get(ch)) {
        const bool sep = ch == '\n' || ch == '^' || ch == ' ';
        if (!sep) { if (tok.size() < 255) tok.push_back(ch); continue; }
        if (!tok.empty()) {
            if (mode == kName) {
                if (one_of(tok, {"DSET", "UNDEF", "XDEF", "YDEF", "ZDEF", "TDEF"})) {
                    if (open) out.back().name = tok; else out.push_back({tok, ""});
                    open = true;
                    mode = kValue;
                }
            } else if (mode == kValue) {
                out.back().value = tok;
                open = false;
                mode = kSkip;
            }
            tok.clear();
        }
        if (ch == '\n') mode = kName;
    }
    if (open) out.pop_back();
}

// ---- raw sub-domain I/O -----------------------------------------------------------------------------------
void flip_elements(unsigned char* p, size_t count, int nbytes)
{
    for (size_t i = 0; i < count; i++, p += nbytes)
        for (int a = 0, b = nbytes - 1; a < b; a++, b--) { unsigned char t = p[a]; p[a] = p[b]; p[b] = t; }
}

std::string proc_label(int id)
{
    std::ostringstream s;
    s << std::setw(WRB_MSSG_FILE_DIG) << std::setfill('0') << id;
    return s.str();
}

struct Box { int nx, ny, nz, nxloc, nyloc, ixst, iyst; };

// Record idset of a sub-domain file -> the (ixst, iyst) block of an nx·ny·nz array of nbytes-wide values.
// reference ctrl_aux.cpp:409-472 (which widens to double; here the file's precision is kept)
int read_block(const std::string& name, int endianflip, int nbytes, int idset, const Box& b, unsigned char* fld)
{
    FILE* f = fopen(name.c_str(), "rb");
    if (!f) return fail(WRB_E_ARG, "Cannot read from " + name);
    const size_t row = (size_t)b.nxloc * nbytes;
    const long long pos = (long long)idset * b.nz * b.nyloc * (long long)row;
    int rc = 0;
    if (fseeko(f, (off_t)pos, SEEK_SET) != 0) rc = fail(WRB_E_FORMAT, "Cannot read from " + name);
    for (int iz = 0; iz < b.nz && rc == 0; iz++)
        for (int iy = b.iyst; iy < b.iyst + b.nyloc; iy++) {
            unsigned char* dst = fld + ((size_t)b.ixst + (size_t)b.nx * ((size_t)iy + (size_t)b.ny * iz)) * nbytes;
            if (fread(dst, 1, row, f) != row) { rc = fail(WRB_E_FORMAT, "Cannot read from " + name); break; }
            if (endianflip) flip_elements(dst, (size_t)b.nxloc, nbytes);
        }
    fclose(f);
    return rc;
}

// The inverse: truncates the file for record 0, appends otherwise (ctrl_aux.cpp:324-405).
int write_block(const std::string& name, int endianflip, int nbytes, int idset, const Box& b, const unsigned char* fld)
{
    FILE* f = fopen(name.c_str(), idset == 0 ? "wb" : "ab");
    if (!f) return fail(WRB_E_ARG, "cannot write to " + name);
    const size_t row = (size_t)b.nxloc * nbytes;
    std::vector<unsigned char> line(row);
    int rc = 0;
    for (int iz = 0; iz < b.nz && rc == 0; iz++)
        for (int iy = b.iyst; iy < b.iyst + b.nyloc; iy++) {
            const unsigned char* src = fld + ((size_t)b.ixst + (size_t)b.nx * ((size_t)iy + (size_t)b.ny * iz)) * nbytes;
            memcpy(line.data(), src, row);
            if (endianflip) flip_elements(line.data(), (size_t)b.nxloc, nbytes);
            if (fwrite(line.data(), 1, row, f) != row) { rc = fail(WRB_E_ARG, "write failed: " + name); break; }
        }
    if (fclose(f) != 0 && rc == 0) rc = fail(WRB_E_ARG, "write failed: " + name);
    return rc;
}

int copy_file(const std::string& from, const std::string& to)
{
    std::ifstream in(from.c_str(), std::ios::binary);
    if (!in) return fail(WRB_E_ARG, "cannot open " + from);
    std::ofstream out(to.c_str(), std::ios::binary | std::ios::trunc);
    if (!out) return fail(WRB_E_ARG, "cannot create " + to);
    out << in.rdbuf();
    return 0;
}

int truncate_file(const std::string& name)
{
    FILE* f = fopen(name.c_str(), "wb");
    if (!f) return fail(WRB_E_ARG, "cannot create " + name);
    fclose(f);
    return 0;
}

int append_bytes(const std::string& name, const unsigned char* p, size_t n)
{
    FILE* f = fopen(name.c_str(), "ab");
    if (!f) return fail(WRB_E_ARG, "cannot open " + name);
    const bool ok = fwrite(p, 1, n, f) == n;
    if (fclose(f) != 0 || !ok) return fail(WRB_E_ARG, "write failed: " + name);
    return 0;
}

template <class T>
double minimum_of(const T* v, size_t n)
{
    T m = v[0];
    for (size_t i = 1; i < n; i++) m = v[i] < m ? v[i] : m;
    return (double)m;
}

// Encode one array through the GPU codec; the output buffer starts at the raw size and grows to setup_wr's
// bound on overflow (the reference allocates the bound up front, mssg_enc.cpp:368).
int encode_array(wrb_codec* c, const void* vals, int dtype, int nx, int ny, int nz, int wtflag, double tol,
                 wrb_header* hdr, std::vector<unsigned char>& enc)
{
    unsigned char nlaymax;
    unsigned long cap;
    wrb_setup(nx, ny, nz, &nlaymax, &cap);
    const size_t n = (size_t)nx * ny * nz;
    unsigned long want = (unsigned long)(n * (dtype == WRB_F32 ? 4 : 8)) + (1ul << 20);
    if (want > cap) want = cap;
    for (;;) {
        enc.resize(want);
        const int rc = wrb_encode_host(c, vals, dtype, nx, ny, nz, wtflag, tol, hdr, enc.data(), want);
        if (rc == WRB_E_OVERFLOW && want < cap) { want = cap; continue; }
        if (rc) return fail(rc, std::string("encoding failed: ") + wrb_last_error(c));
        return 0;
    }
}

int encode_and_store(wrb_codec* c, const void* vals, int dtype, int nx, int ny, int nz, int wtflag, double tol,
                     const std::string& header_name, const std::string& out_name, int idset, const char* dsetname,
                     std::vector<unsigned char>& enc)
{
    wrb_header hdr;
    memset(&hdr, 0, sizeof(hdr));
    int rc = encode_array(c, vals, dtype, nx, ny, nz, wtflag, tol, &hdr, enc);
    if (rc) return rc;
    if (verbose()) std::cout << "        tolabs=" << hdr.tolabs << " nlay=" << (unsigned)hdr.nlay << " ntot_enc=" << hdr.ntot_enc << std::endl;
    if ((rc = wrb_mssg_header_append(header_name.c_str(), idset, dsetname, &hdr))) return rc;
    if (hdr.ntot_enc > 0) rc = append_bytes(out_name, enc.data(), hdr.ntot_enc);
    return rc;
}

// Type 0, one time instant (mssg_enc.cpp:286-398).  vals holds the field in the file's precision.
template <class T>
int encode_regular_field(wrb_codec* c, std::vector<unsigned char>& raw, int dtype, const wrb_mssg_ctl& g, double tol,
                         const std::string& header_name, const std::string& out_name, int it, std::vector<unsigned char>& enc)
{
    const size_t n = (size_t)g.nx * g.ny * g.nz;
    T* v = reinterpret_cast<T*>(raw.data());
    const double minval = minimum_of(v, n);
    // slightly above the mask indicator value (mssg_enc.cpp:309)
    const double thresh = g.undef + fabs(g.undef) * WRB_MSSG_MASK_THRESHOLD_ACC;
    if (!(minval < thresh))
        return encode_and_store(c, v, dtype, g.nx, g.ny, g.nz, 1, tol, header_name, out_name, it, g.dset, enc);
    // Masked field: the pad value is the mean of the unmasked points, summed in array order in double like the
    // reference (the order decides its last bits); the padded field no longer fits single precision.
    double pad = 0;
    int count = 0;
    for (size_t j = 0; j < n; j++)
        if ((double)v[j] >= thresh) { pad += (double)v[j]; count++; }
    pad /= count;
    std::vector<double> fld(n);
    std::vector<T> mask(n);               // two values, minval and 0: both exact in the file's precision
    for (size_t j = 0; j < n; j++) {
        const bool masked = (double)v[j] < thresh;
        fld[j] = masked ? pad : (double)v[j];
        mask[j] = masked ? (T)minval : (T)0;
    }
    if (verbose()) std::cout << " Masking detected, padding with fld_pad=" << pad << ", mask min=" << minval << std::endl;
    int rc = encode_and_store(c, mask.data(), dtype, g.nx, g.ny, g.nz, 0, WRB_MSSG_MASK_TOLREL, header_name, out_name, it, "mask", enc);
    if (rc) return rc;
    return encode_and_store(c, fld.data(), WRB_F64, g.nx, g.ny, g.nz, 1, tol, header_name, out_name, it, g.dset, enc);
}

struct Record { int id; std::string name; wrb_header hdr; };

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

bool parse_ulongs(const std::string& s, unsigned long* v, int n)
{
    const char* p = s.c_str();
    for (int i = 0; i < n; i++) {
        char* e = nullptr;
        v[i] = strtoul(p, &e, 10);
        if (e == p) return false;
        p = e;
    }
    return true;
}

// All of a header file: preamble skipped, optional time record, then coded records until the end of the file.
int read_header_file(const std::string& name, int filetype, double* time_rec, std::vector<Record>& recs)
{
    std::ifstream in(name.c_str());
    if (!in) return fail(WRB_E_ARG, "cannot open " + name);
    std::string line;
    auto next = [&]() -> bool { return (bool)std::getline(in, line); };
    // type 0: 8 preamble lines (mssg_dec.cpp:210-211); types 1/2: those 8 + separator, id, name and the
    // caption of the time record (mssg_dec.cpp:413)
    const int skip = filetype == WRB_MSSG_REGULAR ? 8 : 12;
    for (int j = 0; j < skip; j++)
        if (!next()) return fail(WRB_E_FORMAT, "header file is too short");
    if (filetype != WRB_MSSG_REGULAR) {
        double t[WRB_MSSG_TIME_REC_LEN];
        if (!next() || !parse_doubles(line, t, WRB_MSSG_TIME_REC_LEN)) return fail(WRB_E_FORMAT, "cannot read the time record");
        if (time_rec) memcpy(time_rec, t, sizeof(t));
    }
    while (next()) {                                  // the " -----" separator (ctrl_aux.cpp:541-542)
        Record r;
        memset(&r.hdr, 0, sizeof(r.hdr));
        const std::string where = " in record " + std::to_string(recs.size() + 1) + " of " + name;
        if (!next()) return fail(WRB_E_FORMAT, "missing record id" + where);
        {
            char* e = nullptr;
            const long v = strtol(line.c_str(), &e, 10);
            if (e == line.c_str()) return fail(WRB_E_FORMAT, "bad record id" + where);
            r.id = (int)v;
        }
        if (!next()) return fail(WRB_E_FORMAT, "missing record name" + where);
        r.name = line.size() > 17 ? line.substr(17) : std::string();   // after " Data set name = " (ctrl_aux.cpp:557-558)
        if (!next()) return fail(WRB_E_FORMAT, "missing reminder line" + where);
        double three[3];
        for (int j = 0; j < 3; j++)
            if (!next() || !parse_doubles(line, &three[j], 1)) return fail(WRB_E_FORMAT, "bad floating-point value" + where);
        r.hdr.tolabs = three[0]; r.hdr.midval = three[1]; r.hdr.halfspanval = three[2];
        unsigned long ints[3];
        for (int j = 0; j < 3; j++)
            if (!next() || !parse_ulongs(line, &ints[j], 1)) return fail(WRB_E_FORMAT, "bad integer value" + where);
        if (ints[1] > WRB_NLAYMAX) return fail(WRB_E_FORMAT, "bad layer count" + where);
        r.hdr.wlev = (unsigned char)ints[0]; r.hdr.nlay = (unsigned char)ints[1]; r.hdr.ntot_enc = ints[2];
        if (r.hdr.ntot_enc > 0) {
            const int nl = r.hdr.nlay;
            if (!next() || !parse_doubles(line, r.hdr.deps_vec, nl)) return fail(WRB_E_FORMAT, "bad deps_vec" + where);
            if (!next() || !parse_doubles(line, r.hdr.minval_vec, nl)) return fail(WRB_E_FORMAT, "bad minval_vec" + where);
            if (!next() || !parse_ulongs(line, r.hdr.len_enc_vec, nl)) return fail(WRB_E_FORMAT, "bad len_enc_vec" + where);
        }
        recs.push_back(r);
    }
    return 0;
}

int id_mismatch(int want, int found)
{
    std::ostringstream m;   // ctrl_aux.cpp:547-553
    m << "Encoding header file does not match with the control file: idset+1 = " << want << " idset1 = " << found;
    return fail(WRB_E_FORMAT, m.str());
}

// next `n` bytes of the encoded file (+ pad the decoder may touch)
int read_encoded(FILE* f, unsigned long n, std::vector<unsigned char>& enc)
{
    enc.resize(n + 64);
    if (n > 0 && fread(enc.data(), 1, n, f) != n) return fail(WRB_E_FORMAT, "encoded file is shorter than the header says");
    return 0;
}

}  // namespace

extern "C" {

const char* wrb_mssg_last_error(void) { return g_err.c_str(); }

int wrb_mssg_read_nmlst(const char* name, wrb_mssg_nmlst* out)
{
    if (!name || !out) return fail(WRB_E_ARG, "bad argument");
    std::ifstream in(name);
    if (!in) return fail(WRB_E_ARG, "Unable to open namelist file");
    std::vector<Pair> p;
    int rc = scan_nmlst(in, p);
    if (rc) return rc;
    memset(out, 0, sizeof(*out));
    auto last = [&](const char* key, int& dst) -> bool {
        bool found = false;
        for (const Pair& q : p) if (q.name == key) { dst = atoi(q.value.c_str()); found = true; }
        return found;
    };
    int nproc = 0;
    bool have_grid = false;
    // regional runs give nx, ny; global (Yin-Yang) runs give npg and the overlaps, from which MSSG derives
    // nlg = 3·npg − 4, nx = nlg + 2·i_over, ny = 2·(npg + 2·j_over) (ctrl_aux.cpp:145-177).  When both
    // appear the one listed last wins, as the reference assigns in file order.
    for (const Pair& q : p) {
        if (q.name == "nx") {
            out->nx = atoi(q.value.c_str());
            last("ny", out->ny);
            have_grid = true;
        } else if (q.name == "npg") {
            const int npg = atoi(q.value.c_str());
            int i_over = 0, j_over = 0;
            for (const Pair& s : p) {          // the reference's if / else-if chain: a match on i_over hides j_over only for that entry
                if (s.name == "i_over") i_over = atoi(s.value.c_str());
                else if (s.name == "j_over") j_over = atoi(s.value.c_str());
            }
            out->nx = (3 * npg - 4) + i_over * 2;
            out->ny = (npg + j_over * 2) * 2;
            have_grid = true;
        }
    }
    const bool have_nz = last("nr", out->nz);
    const bool have_np = last("nproc", nproc);
    const bool have_px = last("dim_size", out->nprocx);
    if (!have_grid || !have_nz || !have_np || !have_px || out->nprocx <= 0)
        return fail(WRB_E_FORMAT, "namelist lacks nx/ny (or npg), nr, nproc or dim_size");
    out->nprocy = nproc / out->nprocx;
    // record table: from the first "var" on, the entries alternate var / rec (ctrl_aux.cpp:198-212)
    size_t i = 0;
    while (i < p.size() && p[i].name != "var") i++;
    if (i == p.size()) return fail(WRB_E_FORMAT, "namelist has no var records");
    for (; i < p.size(); i += 2) {
        if (i + 1 >= p.size()) return fail(WRB_E_FORMAT, "var without rec in the namelist");
        const int slot = atoi(p[i + 1].value.c_str()) - 1;
        if (slot < 0 || slot >= WRB_MSSG_NDSMAX) return fail(WRB_E_FORMAT, "record number out of range in the namelist");
        snprintf(out->dset[slot], sizeof(out->dset[slot]), "%s", p[i].value.c_str());
        out->ndset++;
    }
    return 0;
}

int wrb_mssg_read_ctl(const char* name, wrb_mssg_ctl* out)
{
    if (!name || !out) return fail(WRB_E_ARG, "bad argument");
    std::ifstream in(name);
    if (!in) return fail(WRB_E_ARG, "Unable to open namelist file");
    std::vector<Pair> p;
    scan_ctl(in, p);
    memset(out, 0, sizeof(*out));
    int seen = 0;
    for (const Pair& q : p) {                  // the last occurrence of a key wins (ctrl_aux.cpp:300-319)
        const char* v = q.value.c_str();
        if (q.name == "DSET") { snprintf(out->dset, sizeof(out->dset), "%s", v); seen |= 1; }
        else if (q.name == "UNDEF") { out->undef = atof(v); seen |= 2; }
        else if (q.name == "XDEF") { out->nx = atoi(v); seen |= 4; }
        else if (q.name == "YDEF") { out->ny = atoi(v); seen |= 8; }
        else if (q.name == "ZDEF") { out->nz = atoi(v); seen |= 16; }
        else if (q.name == "TDEF") { out->nt = atoi(v); seen |= 32; }
    }
    if ((seen & 61) != 61) return fail(WRB_E_FORMAT, "control file lacks DSET, XDEF, YDEF, ZDEF or TDEF");
    return 0;
}

int wrb_mssg_header_begin(const char* header_name, const char* prefix, const char* ext, int filetype, int nbytes,
                          int endianflip, double tol_base)
{
    if (!header_name || !prefix || !ext) return fail(WRB_E_ARG, "bad argument");
    std::ofstream h(header_name, std::ios::out | std::ios::trunc);
    if (!h) return fail(WRB_E_ARG, std::string("cannot create ") + header_name);
    const bool regular = filetype == WRB_MSSG_REGULAR;
    h << (regular ? " ===== Header file for compressed MSSG regular output data =====" : " ===== Header file for compressed MSSG restart data =====") << "\n"
      << " Coder version: " << 31503 << "\n"            // src/core/defs.h:34 CODER_VERSION
      << " File name prefix: " << prefix << "\n"
      << " Encoded file extension name: " << ext << "\n"
      << " File type (0: regular output; 1: backup merged; 2: backup separated): " << filetype << "\n"
      << " Input files contained " << nbytes << "-byte floating point data" << "\n"
      << (endianflip ? " Converted big endian to little endian or vice versa" : (regular ? " No endian conversion" : " Did not perform endian conversion")) << "\n"
      << " Base cutoff relative tolerance: " << tol_base << "\n";
    return h ? 0 : fail(WRB_E_ARG, "header write failed");
}

int wrb_mssg_header_time(const char* header_name, const char* dsetname, const double* time_rec)
{
    if (!header_name || !dsetname || !time_rec) return fail(WRB_E_ARG, "bad argument");
    std::ofstream h(header_name, std::ios::out | std::ios::app);
    if (!h) return fail(WRB_E_ARG, std::string("cannot open ") + header_name);
    h << " -----\n" << "1\n" << " Data set name = " << dsetname << "\n";
    h << " first " << WRB_MSSG_TIME_REC_LEN << " elements of time record\n";
    h << std::setprecision(kDigits);
    for (int j = 0; j < WRB_MSSG_TIME_REC_LEN; j++) h << time_rec[j] << " ";
    h << "\n";
    return h ? 0 : fail(WRB_E_ARG, "header write failed");
}

int wrb_mssg_header_append(const char* header_name, int idset, const char* dsetname, const wrb_header* hdr)
{
    if (!header_name || !dsetname || !hdr) return fail(WRB_E_ARG, "bad argument");
    std::ofstream h(header_name, std::ios::out | std::ios::app);
    if (!h) return fail(WRB_E_ARG, std::string("cannot open ") + header_name);
    const bool full = hdr->ntot_enc > 0;
    h << " -----\n" << idset + 1 << "\n" << " Data set name = " << dsetname << "\n";
    h << " tolabs; midval; halfspanval; wlev; nlay; ntot_enc;";
    if (full) h << " deps_vec(1:nlay); minval_vec(1:nlay); len_enc_vec(1:nlay)";
    h << "\n" << std::setprecision(kDigits);
    h << hdr->tolabs << "\n" << hdr->midval << "\n" << hdr->halfspanval << "\n";
    h << (unsigned)hdr->wlev << "\n" << (unsigned)hdr->nlay << "\n" << hdr->ntot_enc << "\n";
    if (full) {
        for (int j = 0; j < hdr->nlay; j++) h << hdr->deps_vec[j] << " ";
        h << "\n";
        for (int j = 0; j < hdr->nlay; j++) h << hdr->minval_vec[j] << " ";
        h << "\n";
        for (int j = 0; j < hdr->nlay; j++) h << hdr->len_enc_vec[j] << " ";
        h << "\n";
    }
    return h ? 0 : fail(WRB_E_ARG, "header write failed");
}

int wrb_mssg_header_read(const char* header_name, int filetype, double* time_rec, int* nrec, int* ids,
                         char (*names)[256], wrb_header* hdrs, int max_recs)
{
    if (!header_name || !nrec) return fail(WRB_E_ARG, "bad argument");
    if (filetype < 0 || filetype > 2) return fail(WRB_E_ARG, "Error: unknown file type");
    std::vector<Record> recs;
    const int rc = read_header_file(header_name, filetype, time_rec, recs);
    if (rc) return rc;
    *nrec = (int)recs.size();
    for (int i = 0; i < (int)recs.size() && i < max_recs; i++) {
        if (ids) ids[i] = recs[i].id;
        if (names) snprintf(names[i], 256, "%s", recs[i].name.c_str());
        if (hdrs) hdrs[i] = recs[i].hdr;
    }
    return 0;
}

int wrb_mssg_encode(wrb_codec* c, const char* prefix_c, const char* ext_c, int filetype, int nbytes, int endianflip,
                    double tol_base, int procid)
{
    if (!c || !prefix_c || !ext_c) return fail(WRB_E_ARG, "bad argument");
    if (filetype < 0 || filetype > 2) return fail(WRB_E_ARG, "Error: unknown file type");
    if (nbytes != 4 && nbytes != 8) return fail(WRB_E_ARG, "MSSG input nbytes must be equal to 4 or 8");
    const std::string prefix = prefix_c, ext = ext_c;
    const int dtype = nbytes == 4 ? WRB_F32 : WRB_F64;
    std::vector<unsigned char> raw, enc;
    int rc = 0;

    if (filetype == WRB_MSSG_REGULAR) {
        wrb_mssg_ctl g;
        if ((rc = wrb_mssg_read_ctl((prefix + ".ctl").c_str(), &g))) return rc;
        if (g.nx <= 0 || g.ny <= 0 || g.nz <= 0 || g.nt < 0) return fail(WRB_E_FORMAT, "bad extents in the control file");
        const std::string header_name = prefix + "_h" + ext, out_name = prefix + "_f" + ext;
        if (verbose()) std::cout << " dset=" << g.dset << " nx=" << g.nx << " ny=" << g.ny << " nz=" << g.nz << " nt=" << g.nt << " undef=" << g.undef << std::endl;
        if ((rc = wrb_mssg_header_begin(header_name.c_str(), prefix_c, ext_c, filetype, nbytes, endianflip, tol_base))) return rc;
        if ((rc = truncate_file(out_name))) return rc;
        const Box whole{g.nx, g.ny, g.nz, g.nx, g.ny, 0, 0};
        raw.resize((size_t)g.nx * g.ny * g.nz * nbytes);
        for (int it = 0; it < g.nt && rc == 0; it++) {
            if (verbose()) std::cout << "Field number it=" << it << std::endl;
            if ((rc = read_block(g.dset, endianflip, nbytes, it, whole, raw.data()))) break;
            rc = nbytes == 4 ? encode_regular_field<float>(c, raw, dtype, g, tol_base, header_name, out_name, it, enc)
                             : encode_regular_field<double>(c, raw, dtype, g, tol_base, header_name, out_name, it, enc);
        }
        return rc;
    }

    wrb_mssg_nmlst m;
    if ((rc = wrb_mssg_read_nmlst((prefix + ".nmlst").c_str(), &m))) return rc;
    if (m.nx <= 0 || m.ny <= 0 || m.nz <= 0 || m.nprocy <= 0 || m.ndset < 1) return fail(WRB_E_FORMAT, "bad extents in the namelist");
    const int nxloc = m.nx / m.nprocx, nyloc = m.ny / m.nprocy;
    if (nxloc <= 0 || nyloc <= 0 || (size_t)nxloc * nyloc * m.nz < WRB_MSSG_TIME_REC_LEN)
        return fail(WRB_E_FORMAT, "sub-domain smaller than the time record");
    const bool merged = filetype == WRB_MSSG_RESTART_MERGED;
    const std::string lbl = proc_label(procid);
    const std::string header_name = prefix + "_h" + (merged ? "" : lbl) + ext, out_name = prefix + "_f" + (merged ? "" : lbl) + ext;
    const std::string own_name = prefix + ".p_" + lbl;
    if ((rc = wrb_mssg_header_begin(header_name.c_str(), prefix_c, ext_c, filetype, nbytes, endianflip, tol_base))) return rc;
    const Box local{nxloc, nyloc, m.nz, nxloc, nyloc, 0, 0};
    {   // record 1 ("time") is stored as text: its first 15 values, from this process's file (mssg_enc.cpp:477-486)
        raw.resize((size_t)nxloc * nyloc * m.nz * nbytes);
        if ((rc = read_block(own_name, endianflip, nbytes, 0, local, raw.data()))) return rc;
        double t[WRB_MSSG_TIME_REC_LEN];
        for (int j = 0; j < WRB_MSSG_TIME_REC_LEN; j++)
            t[j] = nbytes == 4 ? (double)reinterpret_cast<const float*>(raw.data())[j] : reinterpret_cast<const double*>(raw.data())[j];
        if ((rc = wrb_mssg_header_time(header_name.c_str(), m.dset[0], t))) return rc;
    }
    if ((rc = truncate_file(out_name))) return rc;
    const int ex = merged ? m.nx : nxloc, ey = merged ? m.ny : nyloc;
    raw.resize((size_t)ex * ey * m.nz * nbytes);
    for (int idset = 1; idset < m.ndset && rc == 0; idset++) {
        if (merged) {
            // sub-domain files of all processes, placed at their offsets in the global array (mssg_enc.cpp:511-533).
            // Points the process grid does not cover (nx % nprocx) are zero here; the reference leaves them unset.
            memset(raw.data(), 0, raw.size());
            for (int py = 0; py < m.nprocy && rc == 0; py++)
                for (int px = 0; px < m.nprocx && rc == 0; px++) {
                    const Box b{m.nx, m.ny, m.nz, nxloc, nyloc, px * nxloc, py * nyloc};
                    rc = read_block(prefix + ".p_" + proc_label(px + m.nprocx * py), endianflip, nbytes, idset, b, raw.data());
                }
        } else {
            rc = read_block(own_name, endianflip, nbytes, idset, local, raw.data());
        }
        if (rc) break;
        if (verbose()) std::cout << " dset=" << m.dset[idset] << " nx=" << ex << " ny=" << ey << " nz=" << m.nz << std::endl;
        rc = encode_and_store(c, raw.data(), dtype, ex, ey, m.nz, 1, tol_base, header_name, out_name, idset, m.dset[idset], enc);
    }
    return rc;
}

int wrb_mssg_decode(wrb_codec* c, const char* in_prefix_c, const char* ext_c, const char* out_prefix_c, int filetype,
                    int nbytes, int endianflip, int procid)
{
    if (!c || !in_prefix_c || !ext_c || !out_prefix_c) return fail(WRB_E_ARG, "bad argument");
    if (filetype < 0 || filetype > 2) return fail(WRB_E_ARG, "Error: unknown file type");
    if (nbytes != 4 && nbytes != 8) return fail(WRB_E_ARG, "MSSG input nbytes must be equal to 4 or 8");
    const std::string in_prefix = in_prefix_c, ext = ext_c, out_prefix = out_prefix_c;
    const int dtype = nbytes == 4 ? WRB_F32 : WRB_F64;
    std::vector<unsigned char> vals, enc;
    std::vector<Record> recs;
    int rc = 0;

    if (filetype == WRB_MSSG_REGULAR) {
        wrb_mssg_ctl g;
        const std::string ctl = in_prefix + ".ctl";
        if ((rc = wrb_mssg_read_ctl(ctl.c_str(), &g))) return rc;
        if (g.nx <= 0 || g.ny <= 0 || g.nz <= 0 || g.nt < 0) return fail(WRB_E_FORMAT, "bad extents in the control file");
        if (in_prefix != out_prefix && (rc = copy_file(ctl, out_prefix + ".ctl"))) return rc;
        if ((rc = read_header_file(in_prefix + "_h" + ext, filetype, nullptr, recs))) return rc;
        const std::string in_name = in_prefix + "_f" + ext, out_name = out_prefix + ".grd";
        FILE* fin = fopen(in_name.c_str(), "rb");
        if (!fin) return fail(WRB_E_ARG, "cannot open " + in_name);
        const size_t n = (size_t)g.nx * g.ny * g.nz;
        const Box whole{g.nx, g.ny, g.nz, g.nx, g.ny, 0, 0};
        vals.resize(n * nbytes);
        std::vector<double> maskv;
        size_t k = 0;
        for (int it = 0; it < g.nt && rc == 0; it++) {
            if (k >= recs.size()) { rc = fail(WRB_E_FORMAT, "header file has fewer records than the control file says"); break; }
            if (recs[k].id != it + 1) { rc = id_mismatch(it + 1, recs[k].id); break; }
            // A record named "mask" precedes the field it belongs to, under the same id (mssg_dec.cpp:236-275).
            // The mask is decoded in double and thresholded against its own midval on the host.
            bool have_mask = false;
            std::vector<unsigned char> is_masked;
            if (recs[k].name == "mask") {
                have_mask = true;
                const wrb_header& mh = recs[k].hdr;
                is_masked.assign(n, 0);
                if (mh.ntot_enc > 0) {
                    if ((rc = read_encoded(fin, mh.ntot_enc, enc))) break;
                    maskv.resize(n);
                    rc = wrb_decode_host(c, maskv.data(), WRB_F64, g.nx, g.ny, g.nz, &mh, enc.data());
                    if (rc) { fail(rc, std::string("decoding failed: ") + wrb_last_error(c)); break; }
                    // The reference rebuilds a binary mask m = (decoded < mask midval ? undef : 0) and later replaces
                    // the field by m wherever m < mask midval (mssg_dec.cpp:260-262, 307-310): 1 -> undef, 2 -> 0
                    // (the latter only if 0 lies below the mask's midval, which a mask of negative undef never has).
                    const unsigned char below = g.undef < mh.midval ? 1 : 0, above = 0.0 < mh.midval ? 2 : 0;
                    for (size_t j = 0; j < n; j++) is_masked[j] = maskv[j] < mh.midval ? below : above;
                    k++;
                    if (k >= recs.size()) { rc = fail(WRB_E_FORMAT, "mask record without its field in the header"); break; }
                    if (recs[k].id != it + 1) { rc = id_mismatch(it + 1, recs[k].id); break; }
                }
                // a mask record without data is also taken as the field's record, like the reference does
            }
            const wrb_header& fh = recs[k].hdr;
            if ((rc = read_encoded(fin, fh.ntot_enc, enc))) break;
            rc = wrb_decode_host(c, vals.data(), dtype, g.nx, g.ny, g.nz, &fh, enc.data());
            if (rc) { fail(rc, std::string("decoding failed: ") + wrb_last_error(c)); break; }
            if (have_mask && recs[k].name != "mask") {
                if (nbytes == 4) { float* v = reinterpret_cast<float*>(vals.data()); const float u = (float)g.undef;
                    for (size_t j = 0; j < n; j++) if (is_masked[j]) v[j] = is_masked[j] == 1 ? u : 0.0f; }
                else { double* v = reinterpret_cast<double*>(vals.data());
                    for (size_t j = 0; j < n; j++) if (is_masked[j]) v[j] = is_masked[j] == 1 ? g.undef : 0.0; }
            }
            k++;
            if (verbose()) std::cout << "Field number it=" << it << " reconstructed" << std::endl;
            rc = write_block(out_name, endianflip, nbytes, it, whole, vals.data());
        }
        fclose(fin);
        return rc;
    }

    wrb_mssg_nmlst m;
    const std::string nml = in_prefix + ".nmlst";
    if ((rc = wrb_mssg_read_nmlst(nml.c_str(), &m))) return rc;
    if (m.nx <= 0 || m.ny <= 0 || m.nz <= 0 || m.nprocy <= 0 || m.ndset < 1) return fail(WRB_E_FORMAT, "bad extents in the namelist");
    const int nxloc = m.nx / m.nprocx, nyloc = m.ny / m.nprocy;
    if (nxloc <= 0 || nyloc <= 0 || (size_t)nxloc * nyloc * m.nz < WRB_MSSG_TIME_REC_LEN)
        return fail(WRB_E_FORMAT, "sub-domain smaller than the time record");
    const bool merged = filetype == WRB_MSSG_RESTART_MERGED;
    if (in_prefix != out_prefix && (rc = copy_file(nml, out_prefix + ".nmlst"))) return rc;
    const std::string lbl = proc_label(procid);
    const std::string header_name = in_prefix + "_h" + (merged ? "" : lbl) + ext, in_name = in_prefix + "_f" + (merged ? "" : lbl) + ext;
    double time_rec[WRB_MSSG_TIME_REC_LEN];
    if ((rc = read_header_file(header_name, filetype, time_rec, recs))) return rc;
    FILE* fin = fopen(in_name.c_str(), "rb");
    if (!fin) return fail(WRB_E_ARG, "cannot open " + in_name);
    const int ex = merged ? m.nx : nxloc, ey = merged ? m.ny : nyloc;
    const size_t n = (size_t)ex * ey * m.nz;
    vals.resize(n * nbytes);
    auto put = [&](size_t j, double v) {
        if (nbytes == 4) reinterpret_cast<float*>(vals.data())[j] = (float)v; else reinterpret_cast<double*>(vals.data())[j] = v;
    };
    auto get = [&](size_t j) -> double {
        return nbytes == 4 ? (double)reinterpret_cast<const float*>(vals.data())[j] : reinterpret_cast<const double*>(vals.data())[j];
    };
    for (int idset = 0; idset < m.ndset && rc == 0; idset++) {
        if (idset == 0) {
            // the time record: zeros but for its first 15 values, repeated at the start of every sub-domain of a
            // merged file (mssg_dec.cpp:406-437)
            memset(vals.data(), 0, vals.size());
            for (int j = 0; j < WRB_MSSG_TIME_REC_LEN; j++) put((size_t)j, time_rec[j]);
            if (merged)
                for (int py = 0; py < m.nprocy; py++)
                    for (int px = 0; px < m.nprocx; px++)
                        if (px + py > 0)
                            for (int ix = 0; ix < WRB_MSSG_TIME_REC_LEN; ix++) {
                                // copied element by element from the array itself, in this order, like the reference
                                // (matters only for rows shorter than the record, where the copies overlap)
                                const size_t j = (size_t)(ix + px * nxloc) + (size_t)m.nx * (size_t)(py * nyloc);
                                if (j < n) put(j, get((size_t)ix));
                            }
        } else {
            if ((size_t)(idset - 1) >= recs.size()) { rc = fail(WRB_E_FORMAT, "header file has fewer records than the namelist says"); break; }
            const Record& r = recs[(size_t)idset - 1];
            if (r.id != idset + 1) { rc = id_mismatch(idset + 1, r.id); break; }
            if ((rc = read_encoded(fin, r.hdr.ntot_enc, enc))) break;
            // ntot_enc == 0: every value equals midval (mssg_dec.cpp:489-493); wrb_decode_host does that
            rc = wrb_decode_host(c, vals.data(), dtype, ex, ey, m.nz, &r.hdr, enc.data());
            if (rc) { fail(rc, std::string("decoding failed: ") + wrb_last_error(c)); break; }
        }
        if (verbose()) std::cout << " dset=" << m.dset[idset] << " reconstructed" << std::endl;
        if (merged) {
            for (int py = 0; py < m.nprocy && rc == 0; py++)
                for (int px = 0; px < m.nprocx && rc == 0; px++) {
                    const Box b{m.nx, m.ny, m.nz, nxloc, nyloc, px * nxloc, py * nyloc};
                    rc = write_block(out_prefix + ".p_" + proc_label(px + m.nprocx * py), endianflip, nbytes, idset, b, vals.data());
                }
        } else {
            const Box local{nxloc, nyloc, m.nz, nxloc, nyloc, 0, 0};
            rc = write_block(out_prefix + ".p_" + lbl, endianflip, nbytes, idset, local, vals.data());
        }
    }
    fclose(fin);
    return rc;
}

}  // extern "C"
