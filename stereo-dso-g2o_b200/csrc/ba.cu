// Windowed bundle adjustment on the device (B1-B12, E2). Filled in by ba_*.cu; this file owns the state.
#include "ctx.h"
namespace sdso {
struct BAState { int dummy = 0; };
int ba_create(sdso_ctx* ctx) { ctx->ba = new BAState(); return SDSO_OK; }
void ba_destroy(sdso_ctx* ctx) { delete ctx->ba; ctx->ba = nullptr; }
}  // namespace sdso
