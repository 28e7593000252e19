// wrmssgdec -- MSSG decoder front-end on the B200 codec.
//
// Command line and interactive questions follow the reference's wrmssgdec (src/mssg/mssg_dec.cpp:98-148):
//   wrmssgdec ENCODED_NAME_PREFIX ENCODED_NAME_EXT EXTRACTED_NAME_PREFIX TYPE PRECISION ENDIANFLIP PROCID
//   TYPE 0: regular output; 1: backup united; 2: backup divided     PRECISION 1: single; 2: double
// (the reference's decoder has no `inmeta` file; its sample `outmeta` is fed through standard input).
// All work happens in wrb_mssg_decode() (libwaverange_b200.so).  Reads files written by this library's
// wrmssgenc (chunk containers) and by the stock wrmssgenc (one stream per layer).
#include <cstdlib>
#include <iostream>
#include <sstream>
#include <string>

#include "../../../include/waverange_mssg.h"

static std::string ask(const char* q)
{
    std::cout << q;
    std::string s;
    std::getline(std::cin, s);
    return s;
}

int main(int argc, char* argv[])
{
    std::string in_prefix, ext, out_prefix, text[4];
    std::cout << "usage: ./wrmssgdec ENCODED_NAME_PREFIX ENCODED_NAME_EXT EXTRACTED_NAME_PREFIX TYPE PRECISION ENDIANFLIP PROCID\n"
                 "where TYPE=(0: regular output; 1: backup united; 2: backup divided), PRECISION=(1:single; 2:double), ENDIANFLIP=(0:no; 1:yes) and PROCID=(this proc id)\n"
                 "interactive mode if not enough arguments are passed.\n";
    if (argc == 8) {
        std::cout << "automatic mode.";
        in_prefix = argv[1]; ext = argv[2]; out_prefix = argv[3];
        for (int j = 0; j < 4; j++) text[j] = argv[4 + j];
    } else {
        in_prefix = ask("Enter encoded data file name prefix []: ");
        ext = ask("Enter encoded data file extension name [.enc]: ");
        out_prefix = ask("Enter extracted data file name prefix []: ");
        text[0] = ask("Enter file type (0: regular output; 1: backup merged; 2: backup separated) [0]: ");
        text[1] = ask("Enter extracted data type (1: float; 2: double) [2]: ");
        text[2] = ask("Enter endian conversion (0: do not perform; 1: inversion) [1]: ");
        text[3] = ask("Enter id of this proc [0]: ");
    }
    // the reference leaves the flip flag unset when nothing is typed (mssg_dec.cpp:75); the prompt promises 1
    int filetype = 0, outtype = 1, flip = 1, procid = 0;
    std::stringstream(text[0]) >> filetype;
    std::stringstream(text[1]) >> outtype;
    std::stringstream(text[2]) >> flip;
    std::stringstream(text[3]) >> procid;
    const int nbytes = outtype == 1 ? 4 : 8;     // mssg_dec.cpp:89,141: the default here is single precision

    std::cout << "\n=== Decoding parameters ===\nEncoded file name prefix: " << in_prefix << "\nEncoded file extension name: " << ext
              << "\nExtracted file name prefix: " << out_prefix
              << "\nFile type (0: regular output; 1: backup merged; 2: backup separated): " << filetype
              << "\nOutput files contain " << nbytes << "-byte floating point data\n";
    if (flip) std::cout << "Convert big endian to little endian or vice versa\n";
    std::cout << "This proc id: " << procid << std::endl;
    if (filetype < 0 || filetype > 2) {
        std::cout << "Error: unknown file type" << std::endl;
        std::cout << "=== End of decompression ===\n";
        return 0;
    }
    const char* dv = getenv("WRB_DEVICE");
    wrb_codec* c = nullptr;
    if (wrb_create(&c, dv ? atoi(dv) : 0)) { std::cerr << "wrmssgdec: no CUDA device (there is no CPU path)" << std::endl; return 2; }
    const int rc = wrb_mssg_decode(c, in_prefix.c_str(), ext.c_str(), out_prefix.c_str(), filetype, nbytes, flip, procid);
    if (rc) std::cerr << "wrmssgdec: " << wrb_mssg_last_error() << std::endl;
    wrb_destroy(c);
    std::cout << "=== End of decompression ===\n";
    return rc ? 1 : 0;
}
