// lfx_synth_host.cpp -> libsynth.so: the synthetic-scan generator's host entry points alone (no CUDA), for processes
// that must not map the product library (bench.py --impl reference, oracle-only tests). Same code as liblfx.so's
// lfx_synth_named / lfx_synth_scan_host (lfx_synth.h), compiled by the same host compiler without FMA contraction.
#include "lfx_synth.h"

extern "C" {
int lfx_synth_named(const char * name, lfx_synth_spec * out) { return lfx_synth::named(name, out); }
int lfx_synth_scan_host(const lfx_synth_spec * spec, uint64_t frame, void * out, uint32_t * n_points_out)
{
  return lfx_synth::scan_host(spec, frame, out, n_points_out);
}
}
