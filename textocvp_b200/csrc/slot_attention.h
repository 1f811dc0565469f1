// Shared constants of the corrector kernels (slot_attention.cu, slot_attention_tc.cu).
#pragma once
#include <stddef.h>
namespace tocvp {
constexpr int SA_S = 8;        // slots
constexpr int SA_D = 128;      // slot dim == feature dim (named configs)
constexpr int SA_CHUNKS = 4;   // location chunks per sequence (one CTA each in the streaming kernels)
constexpr int SA_GVEC = SA_S * SA_D + 2 * SA_S;   // per sequence: g[8][128] (scaled), sg[8], cb[8]
constexpr int SA_PART = SA_S * SA_D + 2 * SA_S;   // per (sequence, chunk): Uacc[8][128], A[8], Mw[8]
}  // namespace tocvp
