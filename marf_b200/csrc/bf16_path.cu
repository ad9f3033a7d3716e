// tcgen05 (bf16 operands, fp32 accumulate in TMEM) implementation of the fused step.  Placeholder until the
// tensor-core path lands: precision=bf16 is refused loudly (there is no fallback to fp32).
#include "engine.cuh"

namespace marf {
struct Bf16State { int unused; };
int bf16_create(marf_handle* h) { return fail(h, MARF_ERR_UNSUPPORTED, "precision=bf16 is not built yet"); }
void bf16_destroy(marf_handle* h) { delete h->bf16; h->bf16 = nullptr; }
bool bf16_supported(const marf_handle*, const marf_step_io*, std::string* why) { if (why) *why = "not built"; return false; }
int bf16_step(marf_handle* h, const marf_step_io*, cudaStream_t) { return fail(h, MARF_ERR_UNSUPPORTED, "precision=bf16 is not built yet"); }
}  // namespace marf
