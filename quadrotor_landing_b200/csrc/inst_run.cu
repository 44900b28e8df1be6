// inst_run.cu -- one translation unit per (real type, est_bias, direct_orien_method, multirate_ekf); compiled
// with -DQ_T=double|float -DQ_BIAS=0|1 -DQ_DIRECT=0|1 -DQ_MR=0|1.  Holds the explicit-stream and the
// synthetic-noise instantiations of the fused multi-tick kernel, with and without per-filter parameters.
#include "launch.hpp"

namespace qekf {

template <typename T, bool BIAS, bool DIRECT, bool SYNTH, bool MR, bool PF>
cudaError_t launch_run(const RunArgs<T> &a, unsigned grid, size_t smem, cudaStream_t stream)
{
    auto kern = run_kernel<T, BIAS, DIRECT, SYNTH, MR, PF, BlockOf<T>::value>;
    cudaError_t e = prep_kernel(kern, smem);
    if (e != cudaSuccess) return e;
    kern<<<grid, BlockOf<T>::value, smem, stream>>>(a);
    return cudaGetLastError();
}

#define INST(S, PF_)                                                                                                  \
    template cudaError_t launch_run<Q_T, (Q_BIAS != 0), (Q_DIRECT != 0), S, (Q_MR != 0), PF_>(const RunArgs<Q_T> &,    \
                                                                                               unsigned, size_t, cudaStream_t);
INST(false, false)
INST(true, false)
INST(false, true)
INST(true, true)

}  // namespace qekf
