/* oracle/stubs — stand-in for tbb/task_scheduler_init.h: Deformable::syncForceModel only asks it for a thread count
 * (src/deformable/Deformable.cpp:204), which the PCG configuration of the integrator ignores. */
#ifndef FB_STUB_TBB_TSI_H
#define FB_STUB_TBB_TSI_H
namespace tbb { struct task_scheduler_init { static int default_num_threads() { return 1; } }; }
#endif
