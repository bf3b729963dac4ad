// Device side of the fp32 output format (see wire.cu): how the exact fix-up kernels store a recomputed
// distance into an fp32 band.
#pragma once
#include "frc_internal.h"

namespace frc {

// out[off] = float(d).  A value fp32 cannot carry to 2^-23 relative (underflow: |d| < 1.2e-38) is also
// appended to the band's exception list in mapped pinned host memory (rare: a system-scope atomic and two
// stores over PCIe).  0, NaN and infinities round-trip and are never exceptions.
__device__ __forceinline__ void store_fixed(float* __restrict__ out, uint32_t off, double d, int64_t first,
                                            const Exceptions& ex) {
  const float f = static_cast<float>(d);
  out[off] = f;
  if (fabs(static_cast<double>(f) - d) > fabs(d) * 0x1p-23) {
    const unsigned long long k = atomicAdd_system(ex.count, 1ULL);
    if (k < static_cast<unsigned long long>(ex.cap)) {
      ex.index[k] = first + off;
      ex.value[k] = d;
    }
  }
}

}  // namespace frc
