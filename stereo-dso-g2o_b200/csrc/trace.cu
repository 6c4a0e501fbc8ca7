// Epipolar search for immature points on the device (D1-D3, E3). State owner.
#include "ctx.h"
namespace sdso {
struct TraceState { int dummy = 0; };
int trace_create(sdso_ctx* ctx) { ctx->trace = new TraceState(); return SDSO_OK; }
void trace_destroy(sdso_ctx* ctx) { delete ctx->trace; ctx->trace = nullptr; }
}  // namespace sdso
