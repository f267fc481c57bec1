/* oracle/wally_config/config.h — TEST INFRASTRUCTURE.  Hand-written libwally config so that
 * external/libwally-core/src/amalgamation/combined.c compiles with plain gcc (no autotools in this image). */
#ifndef LIBWALLYCORE_CONFIG_H
#define LIBWALLYCORE_CONFIG_H
#define HAVE_EXPLICIT_BZERO 1
#define HAVE_UNALIGNED_ACCESS 1
#define HAVE_BYTESWAP_H 1
#define HAVE_LITTLE_ENDIAN 1
#define HAVE_MMAP 1
#define HAVE_POSIX_MEMALIGN 1
#define HAVE_SYS_MMAN_H 1
#define HAVE_INLINE_ASM 1
#include "ccan_config.h"
#endif
