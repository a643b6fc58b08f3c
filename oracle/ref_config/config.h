/* Hand-written stand-in for the reference's generated config.h (config.h.in:13-15,120,128),
 * used ONLY to compile reference sources from /root/reference into oracle/_ref. */
#pragma once
#define PROTO_BASE 0
#define MFSBLOCKSINCHUNK 1024
#define MFSBLOCKSIZE 65536
#define LIZARDFS_HAVE_STD_TO_STRING
#define LIZARDFS_HAVE_STD_STOULL
#define LIZARDFS_HAVE_THREAD_LOCAL
#define LIZARDFS_HAVE_CPU_CHECK
#define HAVE_CRCUTIL
#define ENABLE_CRC
