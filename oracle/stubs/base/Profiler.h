/* oracle/stubs — stand-in for src/base/Profiler.h (needs TBB, Loki): profiling macros as no-ops. */
#ifndef FB_STUB_PROFILER_H
#define FB_STUB_PROFILER_H
#define ProfileAuto() ((void)0)
#define ProfileAutoArg(a) ((void)0)
#define ProfileStart() ((void)0)
#define ProfileEnd() ((void)0)
#define ProfileStartArg(a) ((void)0)
#define ProfileEndArg(a) ((void)0)
#endif
