/* oracle/stubs — placeholder for tbb/concurrent_queue.h. */
#ifndef FB_STUB_TBB_CQ_H
#define FB_STUB_TBB_CQ_H
namespace tbb {}
#endif
