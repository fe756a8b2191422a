/* oracle/stubs — stand-in for src/graphics/SGEffect.h (pulls ShaderManager -> Loki, GL): an empty effect class, which is all
 * SGNode's members need. */
#ifndef FB_STUB_SGEFFECT_H
#define FB_STUB_SGEFFECT_H
#include <memory>
namespace Loki {}
namespace PS { namespace GL {} namespace SG {
class SGEffect { public: virtual ~SGEffect() {} virtual void bind() {} virtual void unbind() {} };
typedef std::shared_ptr<SGEffect> SmartPtrSGEffect;
} }
using namespace Loki;
using namespace PS::GL;
#endif
