/*
 * ref_shim.cpp -- TEST INFRASTRUCTURE ONLY.
 *
 * Compiled by oracle/build_oracle.py only where /root/reference exists.  It
 * pulls the reference's own, unmodified src/core/wrappers.cpp into this
 * translation unit (by path, nothing is copied into the repo) so that its two
 * file-static functions range_encode / range_decode (wrappers.cpp:68,153)
 * become callable; encoding_wrap / decoding_wrap / setup_wr are exported by
 * the included file itself.
 */
#ifndef WR_REF_WRAPPERS
#error "pass -DWR_REF_WRAPPERS='\"/root/reference/src/core/wrappers.cpp\"'"
#endif
#include WR_REF_WRAPPERS

extern "C" void ref_range_encode(unsigned char *sym, unsigned long n, unsigned char *out, unsigned long *len)
{
    unsigned long l = 0;
    range_encode(sym, n, out, l);      /* NB reads sym[n]: caller supplies one pad byte */
    *len = l;
}

extern "C" void ref_range_decode(unsigned char *in, unsigned long len, unsigned char *sym, unsigned long n)
{
    range_decode(in, len, sym, n);
}
