// wrdec -- generic decoder front-end on the B200 codec.
//
// Command line and interactive questions follow the reference's wrdec (src/generic/gen_dec.cpp:98-136):
//   wrdec ENCODED_FILE HEADER_FILE EXTRACTED_FILE TYPE ENDIANFLIP
// All work happens in wrb_file_decode() (libwaverange_b200.so).  Reads files written by this library's
// wrenc (chunk containers) and by the stock wrenc (one stream per layer).
#include <cstdlib>
#include <iostream>
#include <sstream>
#include <string>

#include "../../../include/waverange_files.h"

static std::string ask(const char* q)
{
    std::cout << q;
    std::string s;
    std::getline(std::cin, s);
    return s;
}

int main(int argc, char* argv[])
{
    std::string in = "data.wrb", header = "data.wrh", out = "datarec.bin";
    int filetype = 0, flip = 0;
    std::cout << "usage: ./wrdec ENCODED_FILE HEADER_FILE EXTRACTED_FILE TYPE ENDIANFLIP\n"
                 "where TYPE=(0: Fortran sequential w 4-byte recl; 1: Fortran sequential w 8-byte recl; 2: C/C++) and ENDIANFLIP=(0:no; 1:yes)\n"
                 "interactive mode if not enough arguments are passed.\n";
    if (argc == 6) {
        in = argv[1]; header = argv[2]; out = argv[3];
        std::stringstream(argv[4]) >> filetype;
        std::stringstream(argv[5]) >> flip;
    } else {
        std::string s;
        s = ask("Enter encoded data file name [data.wrb]: "); if (!s.empty()) in = s;
        s = ask("Enter encoding header file name [data.wrh]: "); if (!s.empty()) header = s;
        s = ask("Enter extracted (output) data file name [datarec.bin]: "); if (!s.empty()) out = s;
        s = ask("Enter file type (0: Fortran sequential w 4-byte recl; 1: Fortran sequential w 8-byte recl; 2: C/C++) [0]: ");
        if (!s.empty()) std::stringstream(s) >> filetype;
        s = ask("Enter endian conversion (0: do not perform; 1: inversion) [0]: ");
        if (!s.empty()) std::stringstream(s) >> flip;
    }
    std::cout << "\n=== Decoding parameters ===\nEncoded data file name " << in << "\nEncoding header file name " << header
              << "\nExtracted (output) data file name: " << out << "\nFile type: " << filetype << std::endl;
    if (filetype < 0 || filetype > 2) { std::cout << "Error: unknown file type" << std::endl; return 0; }
    const char* dv = getenv("WRB_DEVICE");
    wrb_codec* c = nullptr;
    if (wrb_create(&c, dv ? atoi(dv) : 0)) { std::cerr << "wrdec: no CUDA device (there is no CPU path)" << std::endl; return 2; }
    const int rc = wrb_file_decode(c, in.c_str(), header.c_str(), out.c_str(), filetype, flip);
    if (rc) std::cerr << "wrdec: " << wrb_file_last_error() << std::endl;
    wrb_destroy(c);
    std::cout << "=== End of decompression ===\n";
    return rc ? 1 : 0;
}
