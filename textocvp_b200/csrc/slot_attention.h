// Shared constants of the corrector kernels (slot_attention.cu, slot_attention_tc.cu).
#pragma once
#include <stddef.h>
namespace tocvp {
constexpr int SA_S = 8;        // slots
constexpr int SA_D = 128;      // slot dim == feature dim (named configs)
constexpr int SA_CHUNKS = 4;   // location chunks per sequence (one CTA each in the streaming kernels)
constexpr int SA_GVEC = SA_S * SA_D + 2 * SA_S;   // per sequence: g[8][128] (scaled), sg[8], cb[8]
constexpr int SA_PART = SA_S * SA_D + 2 * SA_S;   // per (sequence, chunk): Uacc[8][128], A[8], Mw[8]
// Record size (floats) for S slots: S*128 vector entries + 2S scalars, rounded up to a multiple of 4 floats so that every
// record starts 16-byte aligned whatever S and the sequence index are (S odd gave 8-byte-aligned records and misaligned
// 128-bit accesses for odd sequence indices).  sa_part(8) == SA_PART == SA_GVEC.
__host__ __device__ constexpr int sa_part(int S) { return S * SA_D + ((2 * S + 3) & ~3); }
static_assert(sa_part(SA_S) == SA_PART, "record size of the 8-slot kernels");
}  // namespace tocvp
