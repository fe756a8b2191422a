/* oracle/stubs — placeholder for tbb/tick_count.h. */
#ifndef FB_STUB_TBB_TICK_H
#define FB_STUB_TBB_TICK_H
namespace tbb { struct tick_count { static tick_count now() { return tick_count(); } }; }
#endif
