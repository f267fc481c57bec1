/* oracle/ccan_config/config.h — TEST INFRASTRUCTURE.  Minimal hand-written CCAN config so that the
 * reference's ccan/ccan/crypto/sha256/sha256.c compiles with plain gcc (the reference generates
 * this file with ccan/tools/configurator; we only need the handful of feature macros sha256.c,
 * endian.h and compiler.h look at on x86-64/aarch64 little-endian Linux with gcc). */
#ifndef ORACLE_CCAN_CONFIG_H
#define ORACLE_CCAN_CONFIG_H
#ifndef _GNU_SOURCE
#define _GNU_SOURCE
#endif
#define HAVE_ATTRIBUTE_COLD 1
#define HAVE_ATTRIBUTE_CONST 1
#define HAVE_ATTRIBUTE_NORETURN 1
#define HAVE_ATTRIBUTE_PRINTF 1
#define HAVE_ATTRIBUTE_PURE 1
#define HAVE_ATTRIBUTE_UNUSED 1
#define HAVE_ATTRIBUTE_USED 1
#define HAVE_ATTRIBUTE_MAY_ALIAS 1
#define HAVE_ATTRIBUTE_DEPRECATED 1
#define HAVE_ATTRIBUTE_NONSTRING 0
#define HAVE_ATTRIBUTE_SENTINEL 1
#define HAVE_BUILTIN_CONSTANT_P 1
#define HAVE_BUILTIN_EXPECT 1
#define HAVE_WARN_UNUSED_RESULT 1
#define HAVE_BIG_ENDIAN 0
#define HAVE_LITTLE_ENDIAN 1
#define HAVE_BYTESWAP_H 1
#define HAVE_BSWAP_64 1
#define HAVE_TYPEOF 1
#define HAVE_STATEMENT_EXPR 1
#define HAVE_BUILTIN_TYPES_COMPATIBLE_P 1
#define HAVE_BUILTIN_CHOOSE_EXPR 1
#define HAVE_UNALIGNED_ACCESS 1
#endif
