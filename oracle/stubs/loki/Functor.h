/* oracle/stubs — minimal stand-in for Loki's Functor.h: a callable wrapper type that the reference headers can declare
 * members of.  Never invoked by the checkers. */
#ifndef FB_STUB_LOKI_FUNCTOR_H
#define FB_STUB_LOKI_FUNCTOR_H
#include <functional>
namespace Loki {
struct NullType {};
template <typename R = void, class TList = NullType> class Functor {
public:
  Functor() {}
  template <typename F> Functor(F) {}
  template <typename O, typename M> Functor(O, M) {}
  template <typename... A> R operator()(A...) const { return R(); }
};
}
#endif
