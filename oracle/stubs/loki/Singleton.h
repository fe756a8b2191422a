/* oracle/stubs — minimal stand-in for Loki's Singleton.h (library not installed): just enough for the reference HEADERS that
 * mention SingletonHolder to parse.  TEST INFRASTRUCTURE ONLY; no reference or Loki code. */
#ifndef FB_STUB_LOKI_SINGLETON_H
#define FB_STUB_LOKI_SINGLETON_H
namespace Loki {
template <class T> struct CreateUsingNew {};
template <class T> struct PhoenixSingleton {};
template <class T> struct DefaultLifetime {};
template <class T, template <class> class C = CreateUsingNew, template <class> class L = DefaultLifetime>
struct SingletonHolder { static T &Instance() { static T t; return t; } };
}
#endif
