// lfx_sector_extra.cu — the sector kernel (lfx_sector.cuh) for one more convolution_padding, LFX_EXTRA_P, compiled as
// its own translation unit so that the paddings build in parallel (make -j). The two deployed paddings (5: compiled
// default, hyper_parameter.hpp:35; 2: launch YAML) live in lfx_api.cu; 1, 3, 4, 6, 7 and 8 are built from this file.
// Everything non-template in the shared headers has internal linkage, so the units do not collide.
#include <cuda_runtime.h>

#include <type_traits>

#include "lfx_sector.cuh"

#ifndef LFX_EXTRA_P
#error "compile with -DLFX_EXTRA_P=<padding>"
#endif
#define LFX_CAT2(a, b) a##b
#define LFX_CAT(a, b) LFX_CAT2(a, b)

namespace lfxk
{

template<bool DIAG>
static void fill(void (**out)(const SectorArgs))
{
  out[0] = k_extract_sectors<LFX_EXTRA_P, fast_k(0), DIAG, false>;
  out[1] = k_extract_sectors<LFX_EXTRA_P, fast_k(1), DIAG, false>;
  out[2] = k_extract_sectors<LFX_EXTRA_P, fast_k(2), DIAG, false>;
  out[3] = k_extract_sectors<LFX_EXTRA_P, fast_k(0), DIAG, true>;
  out[4] = k_extract_sectors<LFX_EXTRA_P, fast_k(1), DIAG, true>;
  out[5] = k_extract_sectors<LFX_EXTRA_P, fast_k(2), DIAG, true>;
}

// the six instantiations (three lane classes, regular and indexed) of this padding: see pick_sector_kernels, lfx_api.cu
void LFX_CAT(sector_kernels_p, LFX_EXTRA_P)(bool diag, void (**out)(const SectorArgs))
{
  if (diag) { fill<true>(out); } else { fill<false>(out); }
}

}  // namespace lfxk
