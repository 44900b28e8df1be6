// launch_coop.hpp -- host-callable launcher of the cooperative (three lanes per filter) replay kernel of
// ekf_coop.cuh; instantiated in inst_coop.cu (one translation unit per est_bias x direct_orien_method x noise source).
#pragma once

#include <cuda_runtime.h>

#include "ekf_kernels.cuh"

namespace qekf {

// groups of 96 threads (32 filters) per CTA the build instantiates; 0 terminates the list
constexpr int COOP_GROUPS_DEFAULT = 4;
bool coop_groups_available(int groups, bool bench_variant);

// FP64 single-rate only.  Returns cudaErrorInvalidConfiguration when `groups` is not instantiated for this variant.
template <bool BIAS, bool DIRECT, bool SYNTH, bool PF>
cudaError_t launch_run_coop(const RunArgs<double> &a, int groups, cudaStream_t stream);

// the two-role kernel of ekf_duo.cuh (two warps per 32 filters), instantiated in inst_duo.cu
constexpr int DUO_GROUPS_DEFAULT = 6;
bool duo_groups_available(int groups, bool bench_variant);
template <bool BIAS, bool DIRECT, bool SYNTH, bool PF>
cudaError_t launch_run_duo(const RunArgs<double> &a, int groups, cudaStream_t stream);

}  // namespace qekf
