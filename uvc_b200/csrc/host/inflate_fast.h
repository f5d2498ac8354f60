// inflate_fast.h - raw DEFLATE decoder for whole BGZF members (see inflate_fast.cpp)
#ifndef UVC_INFLATE_FAST_H_INCLUDED
#define UVC_INFLATE_FAST_H_INCLUDED

#include <stddef.h>
#include <stdint.h>

extern "C" {
// Decodes the raw deflate stream in[0, in_len) into out[0, out_cap). Returns the number of bytes written, or -1 if the stream is invalid, does
// not end inside the input, or does not fit. The caller guarantees 16 readable bytes after the input (BGZF: the member's footer and the
// buffer's slack).
int64_t uvc_inflate_raw(const uint8_t *in, size_t in_len, uint8_t *out, size_t out_cap);
// Same, and when the fast decoder reports an error, zlib decides (its result is returned: -1 if it also fails).
int64_t uvc_inflate_member(const uint8_t *in, size_t in_len, uint8_t *out, size_t out_cap);
}

#endif
