/* oracle/stubs — stand-in for src/graphics/SGEffect.h (pulls ShaderManager -> Loki, GL): an empty effect class, which is all
 * SGNode's members need. */
#ifndef FB_STUB_SGEFFECT_H
#define FB_STUB_SGEFFECT_H
#include <memory>
namespace Loki {}
namespace PS { namespace GL {} namespace SG {
class SGEffect { public: SGEffect() {} template <typename S> explicit SGEffect(S *) {} virtual ~SGEffect() {} virtual void bind() {} virtual void unbind() {} };
typedef std::shared_ptr<SGEffect> SmartPtrSGEffect;
} }
/* TheShaderManager::Instance().get("phong") — looked up by constructors for drawing; returns no shader here */
namespace PS { namespace GL { struct GLShader {}; struct FbStubShaderManager { template <typename N> GLShader *get(N) const { return 0; } template <typename N> bool has(N) const { return false; } };
struct TheShaderManager { static FbStubShaderManager &Instance() { static FbStubShaderManager m; return m; } }; } }
using namespace Loki;
using namespace PS::GL;
#endif
