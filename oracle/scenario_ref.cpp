// scenario_ref.cpp -- TEST / BASELINE INFRASTRUCTURE, not product: the host-only synthetic landing scenario
// (quadrotor_landing_b200/csrc/scenario.hpp, plain C++ with no CUDA in it) compiled on its own, so that the CPU arm of
// bench.py (`--impl reference`) gets the identical clean workload without loading the product library.
#include "../quadrotor_landing_b200/csrc/scenario.hpp"

extern "C" {

void scn_defaults(qekf_scenario_spec *s) { qekf::scenario::defaults(s); }

// `p` has the layout of qekf_params (= orc_params, oracle/ekf_oracle.h)
void scn_sizes(const qekf_params *p, const qekf_scenario_spec *s, int64_t *T, int64_t *M) { qekf::scenario::sizes(*p, *s, T, M); }

void scn_generate(const qekf_params *p, const qekf_scenario_spec *s, double *truth, double *imu, int32_t *tag_step,
                  double *tag_pose, double *tag_stamp)
{
    qekf::scenario::generate(*p, *s, truth, imu, tag_step, tag_pose, tag_stamp);
}

}
