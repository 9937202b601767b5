// br2_tables.cuh -- device-side view of the lowered assembly (include/br2cuda.h) and kernel parameters.
#pragma once

#include "br2cuda.h"

namespace br2 {

struct DevTables {
    int n_slots, n_rods, n_forcing, n_constraints, n_joints, n_actions;
    long long words;
    const int *slot_rod;
    const int *rod_i;
    const double *rod_d;
    const double *slot_c;
    const int *forcing_i;
    const double *forcing_d;
    const int *cons_i;
    const double *cons_d;
    const int *joint_i;
    const double *joint_d;
    const int *node_ptr, *node_items, *elem_ptr, *elem_items;
    const int *overwrite_src;
    // fast path (BR2_FAST_* in br2cuda.h)
    const int *fast_i;      // [n_slots][BR2_FAST_NI]
    const double *fast_c;   // [BR2_FC_COUNT][n_slots]
    const int *act_i;       // [n_act] action index
    const double *act_d;    // [n_act][3] torque per unit pressure at full ramp
    double act_inv_ramp;    // 1 / ramp_up_time shared by every actuation row
};

struct StepParams {
    DevTables t;
    double *state;
    const double *pressures;
    long long batch;
    long long n_steps;
    double t0, dt;
    int *status;        // [batch] 0 = ok, 1 = the fast kernel left its validity range on this instance
    int only_flagged;   // generic kernel: integrate only instances whose status is set, then clear it
    unsigned long long *replays;  // counts instance-launches redone by the exact path
    double uni[BR2_FC_COUNT];  // uniform copy of fast_c (all element slots identical), constant bank
};

}  // namespace br2
