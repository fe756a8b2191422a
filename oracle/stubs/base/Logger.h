/* oracle/stubs — TEST INFRASTRUCTURE ONLY.  Stand-in for the reference's src/base/Logger.h (needs Loki): the Log* macros of
 * that header as no-ops, so that the reference's own VolMesh.cpp / CuttableMesh.cpp compile here unmodified.  Contains no
 * reference code. */
#ifndef FB_STUB_LOGGER_H
#define FB_STUB_LOGGER_H
#include <stdio.h>
#include <vector>
#include <string>
#include "base/String.h"
static inline void psLog(int, const char *, int, const char *, ...) {}
struct EventLogger { enum { etInfo, etError, etWarning }; };
#define LogInfo(m) ((void)0)
#define LogError(m) ((void)0)
#define LogWarning(m) ((void)0)
#define LogInfoArg1(m, a) ((void)0)
#define LogErrorArg1(m, a) ((void)0)
#define LogWarningArg1(m, a) ((void)0)
#define LogInfoArg2(m, a, b) ((void)0)
#define LogErrorArg2(m, a, b) ((void)0)
#define LogWarningArg2(m, a, b) ((void)0)
#define LogInfoArg3(m, a, b, c) ((void)0)
#define LogErrorArg3(m, a, b, c) ((void)0)
#define LogWarningArg3(m, a, b, c) ((void)0)
#define LogInfoArg4(m, a, b, c, d) ((void)0)
#define LogErrorArg4(m, a, b, c, d) ((void)0)
#define LogWarningArg4(m, a, b, c, d) ((void)0)
#endif
