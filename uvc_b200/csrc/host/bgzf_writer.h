// bgzf_writer.h - block-gzip output of the uvc1 host. Replaces the reference's bgzip_string / clearstring / bgzf_close use of htslib
// (main.cpp:99-130, 1541-1551, 1555, 1581): text is cut into blocks of at most 0xff00 bytes, each deflated independently (so blocks are
// compressed in parallel on the host threads) and framed as a BGZF member (SAM specification section 4.1); the file ends with the
// 28-byte empty end-of-file block.
#ifndef UVC_BGZF_WRITER_H_INCLUDED
#define UVC_BGZF_WRITER_H_INCLUDED

#include <stddef.h>
#include <string>

// Appends the BGZF members of `text` to `out`. level = zlib level (the reference uses 5). n_threads <= 1 compresses serially.
int uvchost_bgzf_compress(std::string & out, const char *text, size_t len, int level, int n_threads);
// The end-of-file marker block.
void uvchost_bgzf_eof(std::string & out);

#endif
