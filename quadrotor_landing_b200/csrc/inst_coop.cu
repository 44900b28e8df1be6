// inst_coop.cu -- instantiations of the cooperative replay kernel; compiled with -DQ_BIAS=0|1 -DQ_DIRECT=0|1
// -DQ_SYNTH=0|1 (with and without per-filter parameters each).  The variant the benchmark runs (15 states, direct
// model, in-kernel noise, launch-wide parameters) is built for every CTA size of the sweep, the others for the default.
#include "launch_coop.hpp"
#include "launch.hpp"
#include "ekf_coop.cuh"

namespace qekf {

template <bool BIAS, bool DIRECT, bool SYNTH, bool PF, int G>
static cudaError_t launch_g(const RunArgs<double> &a, cudaStream_t stream)
{
    constexpr int NB = BIAS ? 5 : 3;
    auto kern = coop::run_kernel_coop<double, BIAS, DIRECT, SYNTH, PF, G>;
    const size_t smem = coop::coop_smem_bytes<NB>(G, sizeof(double));
    cudaError_t e = prep_kernel(kern, smem);
    if (e != cudaSuccess) return e;
    const unsigned grid = (unsigned)((a.st.n + 32 * G - 1) / (32 * G));
    kern<<<grid, 96 * G, smem, stream>>>(a);
    return cudaGetLastError();
}

#define QB (Q_BIAS != 0)
#define QD (Q_DIRECT != 0)
#define QS (Q_SYNTH != 0)

template <> cudaError_t launch_run_coop<QB, QD, QS, false>(const RunArgs<double> &a, int groups, cudaStream_t stream)
{
    switch (groups) {
    case COOP_GROUPS_DEFAULT: return launch_g<QB, QD, QS, false, COOP_GROUPS_DEFAULT>(a, stream);
#if Q_BIAS && Q_DIRECT && Q_SYNTH
    case 3: return launch_g<QB, QD, QS, false, 3>(a, stream);
    case 5: return launch_g<QB, QD, QS, false, 5>(a, stream);
    case 6: return launch_g<QB, QD, QS, false, 6>(a, stream);
#endif
    default: return cudaErrorInvalidConfiguration;
    }
}
template <> cudaError_t launch_run_coop<QB, QD, QS, true>(const RunArgs<double> &a, int groups, cudaStream_t stream)
{
    if (groups != COOP_GROUPS_DEFAULT) return cudaErrorInvalidConfiguration;
    return launch_g<QB, QD, QS, true, COOP_GROUPS_DEFAULT>(a, stream);
}

#if Q_BIAS && Q_DIRECT && Q_SYNTH
bool coop_groups_available(int groups, bool bench_variant)
{
    if (groups == COOP_GROUPS_DEFAULT) return true;
    return bench_variant && (groups == 3 || groups == 5 || groups == 6);
}
#endif

}  // namespace qekf
