// inst_duo.cu -- instantiations of the two-role replay kernel (ekf_duo.cuh); compiled with -DQ_BIAS=0|1 -DQ_DIRECT=0|1
// -DQ_SYNTH=0|1 (with and without per-filter parameters each).  The variant the benchmark runs (15 states, direct
// model, in-kernel noise, launch-wide parameters) is built for every CTA size of the sweep, the others for the default.
#include "launch_coop.hpp"
#include "launch.hpp"
#include "ekf_duo.cuh"

namespace qekf {

template <bool BIAS, bool DIRECT, bool SYNTH, bool PF, int G>
static cudaError_t launch_g(const RunArgs<double> &a, cudaStream_t stream)
{
    constexpr int N = BIAS ? 15 : 9;
    auto kern = duo::run_kernel_duo<double, BIAS, DIRECT, SYNTH, PF, G>;
    const size_t smem = duo::duo_smem_bytes<N>(G, sizeof(double));
    cudaError_t e = prep_kernel(kern, smem);
    if (e != cudaSuccess) return e;
    const unsigned grid = (unsigned)((a.st.n + 32 * G - 1) / (32 * G));
    kern<<<grid, 64 * G, smem, stream>>>(a);
    return cudaGetLastError();
}

#define QB (Q_BIAS != 0)
#define QD (Q_DIRECT != 0)
#define QS (Q_SYNTH != 0)

template <> cudaError_t launch_run_duo<QB, QD, QS, false>(const RunArgs<double> &a, int groups, cudaStream_t stream)
{
    switch (groups) {
    case DUO_GROUPS_DEFAULT: return launch_g<QB, QD, QS, false, DUO_GROUPS_DEFAULT>(a, stream);
#if Q_BIAS && Q_DIRECT && Q_SYNTH
    case 4: return launch_g<QB, QD, QS, false, 4>(a, stream);
    case 5: return launch_g<QB, QD, QS, false, 5>(a, stream);
#endif
    default: return cudaErrorInvalidConfiguration;
    }
}
template <> cudaError_t launch_run_duo<QB, QD, QS, true>(const RunArgs<double> &a, int groups, cudaStream_t stream)
{
    if (groups != DUO_GROUPS_DEFAULT) return cudaErrorInvalidConfiguration;
    return launch_g<QB, QD, QS, true, DUO_GROUPS_DEFAULT>(a, stream);
}

#if Q_BIAS && Q_DIRECT && Q_SYNTH
bool duo_groups_available(int groups, bool bench_variant)
{
    if (groups == DUO_GROUPS_DEFAULT) return true;
    return bench_variant && (groups == 4 || groups == 5);
}
#endif

}  // namespace qekf
