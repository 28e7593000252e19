// wrmssgenc -- MSSG encoder front-end on the B200 codec.
//
// Command line, `inmeta` parameter file and interactive questions follow the reference's wrmssgenc
// (src/mssg/mssg_enc.cpp:104-234) so that scripts written for it keep working:
//   wrmssgenc FILE_NAME_PREFIX ENCODED_NAME_EXT TYPE PRECISION ENDIANFLIP TOLERANCE PROCID
//   TYPE 0: regular output; 1: backup united; 2: backup divided     PRECISION 1: single; 2: double
// An `inmeta` file in the working directory takes precedence: "&name = value" lines (prefix_name, ext_name,
// file_type, input_data_type, endian_conversion, tolerance, id_of_proc), or the old format with one value per
// line in that order.  Without 7 arguments and without `inmeta` the parameters are asked for interactively.
// All work happens in wrb_mssg_encode() (libwaverange_b200.so); WRB_CHUNK_BLOCKS=0 writes files the stock
// wrmssgdec can read, WRB_DEVICE selects the GPU.
#include <algorithm>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>

#include "../../../include/waverange_mssg.h"

namespace {

struct Job {
    std::string prefix, ext = ".enc";
    std::string text[5];   // file type, precision, endian flip, tolerance, proc id -- as typed
};

std::string trimmed(const std::string& s)
{
    const char* ws = " \t\v\r\n";
    const size_t a = s.find_first_not_of(ws);
    if (a == std::string::npos) return "";
    return s.substr(a, s.find_last_not_of(ws) - a + 1);
}

std::string ask(const char* q)
{
    std::cout << q;
    std::string s;
    std::getline(std::cin, s);
    return s;
}

// 0: parsed, -1: malformed
int from_inmeta(std::ifstream& f, Job& job)
{
    static const char* keys[5] = {"&file_type", "&input_data_type", "&endian_conversion", "&tolerance", "&id_of_proc"};
    std::vector<std::string> lines;
    for (std::string s; std::getline(f, s);) lines.push_back(s);
    bool namelist = false;
    for (const std::string& raw : lines) {
        const std::string t = trimmed(raw);
        if (t.empty() || t[0] != '&') continue;       // anything else is a comment
        const size_t eq = t.find('=');
        if (eq == std::string::npos || eq + 1 >= t.size()) {
            std::cout << "==== Error 'value' is missing in a sentence :" << t << "====" << std::endl;
            return -1;
        }
        if (t.find('=', eq + 1) != std::string::npos) {
            std::cout << "==== Error : '=' exists twice in a sentence :" << t << "====" << std::endl;
            return -1;
        }
        namelist = true;
        std::string key = trimmed(t.substr(0, eq));
        const std::string val = trimmed(t.substr(eq + 1));
        std::transform(key.begin(), key.end(), key.begin(), ::tolower);
        if (key == "&prefix_name") job.prefix = val;
        else if (key == "&ext_name") job.ext = val;
        else for (int j = 0; j < 5; j++) if (key == keys[j]) job.text[j] = val;
    }
    if (!namelist) {
        std::cout << "==== read parameters from inmeta as old format. ====" << std::endl;
        auto line = [&](size_t i) { return i < lines.size() ? lines[i] : std::string(); };
        job.prefix = line(0);
        job.ext = line(1);
        for (int j = 0; j < 5; j++) job.text[j] = line(2 + (size_t)j);
    }
    return 0;
}

}  // namespace

int main(int argc, char* argv[])
{
    Job job;
    std::ifstream meta("inmeta");
    if (meta) {
        std::cout << "==== inmeta exists. ====" << std::endl;
        if (from_inmeta(meta, job)) return -1;
    } else {
        std::cout << "inmeta doesn't exists." << std::endl;
        std::cout << "usage: ./wrmssgenc FILE_NAME_PREFIX ENCODED_NAME_EXT TYPE PRECISION ENDIANFLIP TOLERANCE PROCID\n"
                     "where TYPE=(0: regular output; 1: backup united; 2: backup divided), PRECISION=(1:single; 2:double), ENDIANFLIP=(0:no; 1:yes), TOLERANCE=(e.g. 1.0e-16) and PROCID=(this proc id)\n"
                     "interactive mode if not enough arguments are passed.\n";
        if (argc == 8) {
            std::cout << "automatic mode.";
            job.prefix = argv[1];
            job.ext = argv[2];
            for (int j = 0; j < 5; j++) job.text[j] = argv[3 + j];
        } else {
            job.prefix = ask("Enter data file name prefix []: ");
            job.ext = ask("Enter encoded file extension name [.enc]: ");
            job.text[0] = ask("Enter file type (0: regular output; 1: backup merged; 2: backup separated) [0]: ");
            job.text[1] = ask("Enter input data type (1: float; 2: double) [2]: ");
            job.text[2] = ask("Enter endian conversion (0: do not perform; 1: inversion) [1]: ");
            job.text[3] = ask("Enter base cutoff relative tolerance [1e-16]: ");
            job.text[4] = ask("Enter id of this proc [0]: ");
        }
    }
    // defaults of the reference (mssg_enc.cpp:75-103); an empty or unparsable answer keeps them
    int filetype = 0, intype = 2, flip = 1, procid = 0;
    double tol = 1e-16;
    std::stringstream(job.text[0]) >> filetype;
    std::stringstream(job.text[1]) >> intype;
    std::stringstream(job.text[2]) >> flip;
    std::stringstream(job.text[3]) >> tol;
    std::stringstream(job.text[4]) >> procid;
    const int nbytes = intype == 1 ? 4 : 8;

    std::cout << "\n=== Compression parameters ===\nData file name prefix: " << job.prefix << "\nEncoded file extension name: " << job.ext
              << "\nFile type (0: regular output; 1: backup merged; 2: backup separated): " << filetype
              << "\nInput files contain " << nbytes << "-byte floating point data\n";
    if (flip) std::cout << "Convert big endian to little endian or vice versa\n";
    std::cout << "Base cutoff relative tolerance: " << tol << "\nThis proc id: " << procid << std::endl;
    if (filetype < 0 || filetype > 2) {
        std::cout << "Error: unknown file type" << std::endl;
        std::cout << "=== End of compression ===\n";
        return 0;
    }
    const char* dv = getenv("WRB_DEVICE");
    wrb_codec* c = nullptr;
    if (wrb_create(&c, dv ? atoi(dv) : 0)) { std::cerr << "wrmssgenc: no CUDA device (there is no CPU path)" << std::endl; return 2; }
    const int rc = wrb_mssg_encode(c, job.prefix.c_str(), job.ext.c_str(), filetype, nbytes, flip, tol, procid);
    if (rc) std::cerr << "wrmssgenc: " << wrb_mssg_last_error() << std::endl;
    wrb_destroy(c);
    std::cout << "=== End of compression ===\n";
    return rc ? 1 : 0;
}
