/* oracle/stubs — minimal stand-in for tbb/task.h (TBB not installed): a base class for the reference's DBInsertionTask
 * declaration.  Never scheduled by the checkers. */
#ifndef FB_STUB_TBB_TASK_H
#define FB_STUB_TBB_TASK_H
namespace tbb { class task { public: virtual ~task() {} virtual task *execute() = 0; }; }
#endif
