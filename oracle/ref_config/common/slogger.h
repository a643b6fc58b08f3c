/* Shadows the reference's common/slogger.h (which needs spdlog, absent in this image) so that
 * common/massert.h compiles.  Only used for the oracle/_ref build. */
#pragma once
#include <cstdio>
#include <syslog.h>
#define lzfs_pretty_syslog(prio, ...) do { fprintf(stderr, __VA_ARGS__); fputc('\n', stderr); } while (0)
#define lzfs_pretty_errlog(prio, ...) lzfs_pretty_syslog(prio, __VA_ARGS__)
#define lzfs_silent_syslog(prio, ...) ((void)0)
#define lzfs_silent_errlog(prio, ...) ((void)0)
