// wrenc -- generic encoder front-end on the B200 codec.
//
// Command line, `inmeta` parameter file and interactive questions follow the reference's wrenc
// (src/generic/gen_enc.cpp:111-500) so that scripts written for it keep working:
//   wrenc INPUT_FILE ENCODED_FILE HEADER_FILE TYPE ENDIANFLIP NF PRECISION NX NY NZ TOLERANCE
//   TYPE       0: Fortran sequential w 4-byte recl; 1: Fortran sequential w 8-byte recl; 2: C/C++
//   ENDIANFLIP 0: no; 1: yes      NF number of fields      PRECISION 1: single; 2: double
// An `inmeta` file in the working directory takes precedence (namelist-like "&name = value" lines with a
// "%field = i" ... "/" block per field, or the old format: one value per line in the order of the
// questions); without arguments and without `inmeta` the parameters are asked for interactively.
// All work happens in wrb_file_encode() (libwaverange_b200.so); WRB_CHUNK_BLOCKS=0 writes files the stock
// wrdec can read, WRB_DEVICE selects the GPU.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>

#include "../../../include/waverange_files.h"

namespace {

struct Job {
    std::string in = "data.bin", out = "data.wrb", header = "data.wrh";
    int filetype = 0, flip = 0, nf = 1;
    double cutoff = 1e-16;      // the tolerance parsed last: the reference applies it to every field (gen_enc.cpp:497-500)
    std::vector<wrb_field_desc> fields;
};

std::string trimmed(const std::string& s)
{
    const char* ws = " \t\v\r\n";
    const size_t a = s.find_first_not_of(ws);
    if (a == std::string::npos) return "";
    return s.substr(a, s.find_last_not_of(ws) - a + 1);
}

template <class T> void take(const std::string& s, T& v) { if (!s.empty()) std::stringstream(s) >> v; }

wrb_field_desc default_field() { wrb_field_desc d{8, 16, 16, 16, 1, 0, 1, 1e-16}; return d; }

// "&name = value" -> (lower-case name, value); false when the line is not of that form
bool key_value(const std::string& line, std::string& key, std::string& val)
{
    const size_t eq = line.find('=');
    if (eq == std::string::npos || line.find('=', eq + 1) != std::string::npos) return false;
    key = trimmed(line.substr(0, eq));
    val = trimmed(line.substr(eq + 1));
    std::transform(key.begin(), key.end(), key.begin(), ::tolower);
    return true;
}

int from_inmeta(Job& job)
{
    std::ifstream f("inmeta");
    std::vector<std::string> lines;
    for (std::string s; std::getline(f, s);) lines.push_back(s);
    bool namelist = false;
    for (const std::string& s : lines) { const std::string t = trimmed(s); if (!t.empty() && t[0] == '&') namelist = true; }
    wrb_field_desc cur = default_field();
    int prec = 2;
    if (namelist) {
        // globals first (they may appear anywhere), then the field blocks in order
        for (const std::string& s : lines) {
            std::string k, v; const std::string t = trimmed(s);
            if (t.empty() || t[0] != '&') continue;
            if (!key_value(t, k, v)) { std::cout << "==== Error : malformed line in inmeta :" << t << " ====" << std::endl; return -1; }
            if (k == "&in_name") job.in = v.empty() ? job.in : v;
            else if (k == "&out_name") job.out = v.empty() ? job.out : v;
            else if (k == "&header_name") job.header = v.empty() ? job.header : v;
            else if (k == "&file_type") take(v, job.filetype);
            else if (k == "&endian_conversion") take(v, job.flip);
            else if (k == "&number_of_field") take(v, job.nf);
        }
        job.fields.assign((size_t)std::max(job.nf, 0), default_field());
        int id = -1, blocks = 0;
        for (const std::string& s : lines) {
            std::string k, v; const std::string t = trimmed(s);
            if (t.empty()) continue;
            if (t[0] == '%' && key_value(t, k, v) && k == "%field" && !v.empty()) { take(v, id); blocks++; }
            else if (t[0] == '&' && key_value(t, k, v)) {
                if (k == "&input_data_type") take(v, prec);
                else if (k == "&nx") take(v, cur.nx);
                else if (k == "&ny") take(v, cur.ny);
                else if (k == "&nz") take(v, cur.nz);
                else if (k == "&nh") take(v, cur.nh);
                else if (k == "&order") take(v, cur.idinv);
                else if (k == "&compress") take(v, cur.icomp);
                else if (k == "&tolerance") take(v, cur.tol_base);
            } else if (t[0] == '/') {       // end of a field block: values not given carry over from the previous block
                cur.nbytes = (prec == 1) ? 4 : 8;
                job.cutoff = cur.tol_base;
                if (id >= 0 && id < job.nf) job.fields[(size_t)id] = cur;
            }
        }
        if (blocks != job.nf) {
            std::cout << "==== Number of fields is " << job.nf << " but " << blocks << " field blocks were found in inmeta ====" << std::endl;
            return -1;
        }
    } else {                                 // old format: one value per line
        size_t p = 0;
        auto nxt = [&]() -> std::string { return p < lines.size() ? lines[p++] : std::string(); };
        std::string s;
        s = nxt(); if (!s.empty()) job.in = s;
        s = nxt(); if (!s.empty()) job.out = s;
        s = nxt(); if (!s.empty()) job.header = s;
        take(nxt(), job.filetype); take(nxt(), job.flip); take(nxt(), job.nf);
        for (int it = 0; it < job.nf; it++) {
            take(nxt(), prec); take(nxt(), cur.nx); take(nxt(), cur.ny); take(nxt(), cur.nz); take(nxt(), cur.nh);
            take(nxt(), cur.idinv); take(nxt(), cur.icomp); take(nxt(), cur.tol_base);
            cur.nbytes = (prec == 1) ? 4 : 8;
            job.cutoff = cur.tol_base;
            job.fields.push_back(cur);
        }
    }
    return 0;
}

std::string ask(const char* q)
{
    std::cout << q;
    std::string s;
    std::getline(std::cin, s);
    return s;
}

void interactive(Job& job)
{
    std::string s;
    s = ask("Enter input data file name [data.bin]: "); if (!s.empty()) job.in = s;
    s = ask("Enter encoded data file name [data.wrb]: "); if (!s.empty()) job.out = s;
    s = ask("Enter encoding header file name [data.wrh]: "); if (!s.empty()) job.header = s;
    take(ask("Enter file type (0: Fortran sequential w 4-byte recl; 1: Fortran sequential w 8-byte recl; 2: C/C++) [0]: "), job.filetype);
    take(ask("Enter endian conversion (0: do not perform; 1: inversion) [0]: "), job.flip);
    take(ask("Enter the number of fields in the file, nf [1]: "), job.nf);
    wrb_field_desc cur = default_field();
    int prec = 2;
    for (int it = 0; it < job.nf; it++) {
        std::cout << "Field number " << it << std::endl;
        take(ask("Enter input data type (1: float; 2: double) [2]: "), prec);
        cur.nbytes = (prec == 1) ? 4 : 8;
        take(ask("Enter the number of data points in the first dimension, nx [16]: "), cur.nx);
        take(ask("Enter the number of data points in the second dimension, ny [16]: "), cur.ny);
        take(ask("Enter the number of data points in the third dimension, nz [16]: "), cur.nz);
        take(ask("Enter the number of data points in the higher (slowest) dimensions, nh [1]: "), cur.nh);
        take(ask("Invert the order of the dimensions? (0: no; 1: yes) [0]: "), cur.idinv);
        take(ask("Enter compression flag (0: do not compress; 1: compress) [1]: "), cur.icomp);
        if (cur.icomp) take(ask("Enter base cutoff relative tolerance [1e-16]: "), cur.tol_base);
        wrb_field_desc d = cur;
        if (!d.icomp) d.tol_base = 0;
        job.cutoff = cur.tol_base;
        job.fields.push_back(d);
    }
}

}  // namespace

int main(int argc, char* argv[])
{
    Job job;
    if (std::ifstream("inmeta").good()) {
        std::cout << "==== inmeta exists. ====" << std::endl;
        if (from_inmeta(job)) return -1;
    } else {
        std::cout << "usage: ./wrenc INPUT_FILE ENCODED_FILE HEADER_FILE TYPE ENDIANFLIP NF PRECISION NX NY NZ TOLERANCE\n"
                     "where TYPE=(0: Fortran sequential w 4-byte recl; 1: Fortran sequential w 8-byte recl; 2: C/C++),\n"
                     "      ENDIANFLIP=(0:no; 1:yes), NF=(how many fields, e.g. 1), PRECISION=(1:single; 2:double),\n"
                     "      NX=(e.g. 16), NY=(e.g. 16), NZ=(e.g. 16) and TOLERANCE=(e.g. 1.0e-16)\n"
                     "interactive mode if not enough arguments are passed.\n";
        if (argc == 12) {
            job.in = argv[1]; job.out = argv[2]; job.header = argv[3];
            int prec = 2;
            wrb_field_desc d = default_field();
            take(argv[4], job.filetype); take(argv[5], job.flip); take(argv[6], job.nf); take(argv[7], prec);
            take(argv[8], d.nx); take(argv[9], d.ny); take(argv[10], d.nz); take(argv[11], d.tol_base);
            d.nbytes = (prec == 1) ? 4 : 8;
            job.cutoff = d.tol_base;
            job.fields.assign((size_t)std::max(job.nf, 0), d);
        } else {
            interactive(job);
        }
    }
    std::cout << "\n=== Compression parameters ===\n"
              << "Input data file name: " << job.in << "\nEncoded data file name: " << job.out
              << "\nEncoding header file name: " << job.header << "\nFile type: " << job.filetype
              << "\nNumber of fields in the file, nf: " << job.nf << std::endl;
    if (job.filetype < 0 || job.filetype > 2) { std::cout << "Error: unknown file type" << std::endl; return 0; }
    const char* dv = getenv("WRB_DEVICE");
    wrb_codec* c = nullptr;
    if (wrb_create(&c, dv ? atoi(dv) : 0)) { std::cerr << "wrenc: no CUDA device (there is no CPU path)" << std::endl; return 2; }
    const int rc = wrb_file_encode(c, job.in.c_str(), job.out.c_str(), job.header.c_str(), job.filetype, job.flip, job.nf, job.fields.data(), &job.cutoff);
    if (rc) std::cerr << "wrenc: " << wrb_file_last_error() << std::endl;
    wrb_destroy(c);
    std::cout << "=== End of compression ===\n";
    return rc ? 1 : 0;
}
