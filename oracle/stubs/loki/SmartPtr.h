/* oracle/stubs — placeholder for Loki's SmartPtr.h (the compiled reference files only use std::shared_ptr). */
#ifndef FB_STUB_LOKI_SMARTPTR_H
#define FB_STUB_LOKI_SMARTPTR_H
namespace Loki {}
#endif
