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

// Thread-coarsened ("tile") kernel: every thread owns C consecutive slots of one rod (br2_tile.cuh).
#define BR2_TILE_MAX_RODS 12
#define BR2_MAX_CHUNKS 8
#define BR2_TILE_MAX_CLS 4   // classes of rods with equal element constants (rods of different length: one class per length)
struct TilePlan {
    int C;                                  // slots per thread
    int round_section;                      // ln c_r1 = ln c_r2 = 2 ln c_r3: one exponential per element and step
    int n_ctas;                             // CTAs per instance: 1, or one per group of ring-connected rods (a thread-block
                                            // cluster) when the assembly does not fit the registers of one SM
    int rod_cta[BR2_TILE_MAX_RODS];         // cluster rank of the CTA that integrates the rod
    int cta_threads[8];                     // threads in use per rank
    int xi_stride;                          // rows of `xi` per rank
    int warp_map[16];                       // cluster kernels: logical warp executed by each hardware warp (a permutation)
    int rod_t0[BR2_TILE_MAX_RODS + 1];      // first thread of each rod inside its CTA
    int n_cls;                              // classes of rods with equal element constants (1: the constant bank serves all)
    int rod_cls[BR2_TILE_MAX_RODS];         // class of each rod
    int cls_slot[BR2_TILE_MAX_CLS];         // a table slot that carries the class's constants (host side)
    int partner[BR2_TILE_MAX_RODS];         // rod whose elements are rod one of the joints this rod's elements own, or -1
    int owner[BR2_TILE_MAX_RODS];           // rod whose elements own the joints in which this rod's elements are rod one, or -1
    double jd[BR2_TILE_MAX_RODS][8];        // {dv2[3], offset/2} of the joints the rod owns, {dv1[3], offset/2} as rod one
    // "extra" joints: tip-to-base joints between segments (TipToTipStraightJoint).  Few per
    // assembly; evaluated by designated threads from data their elements publish; none for single-segment assemblies.
    int n_xe, n_xj;                         // elements that publish director / omega / radius / node sums; extra joints
    int n_nr, nf_items;                     // node rows per CTA (nodes that add extra-joint forces) and items per row
    int xb_count[8];                        // per rank: published elements of OTHER CTAs whose data the CTA consumes (x 152 bytes per step on its XB barrier)
    const int *xe_mask;                     // [n_xe] bit r: the CTA of rank r consumes the element's data (device)
    const int *xi;                          // [n_ctas][xi_stride][C + 1] per thread: packed info of each slot (published xe + 1
                                            // | overwrite source xe + 1 << 8 | node items << 16 | node row << 20), then the
                                            // extra joint component the thread evaluates (joint << 2 | component) or -1 (device).
                                            // A joint between two CTAs of a cluster is evaluated by BOTH (each for its own nodes)
    const int *xjt;                         // [n_ctas][n_xj][4] where in the CTA's node rows the joint's half force goes: rod
                                            // one's two nodes (+), rod two's two nodes (-); -1: not in this CTA (device)
    const int *xj_i;                        // [n_xj][12] rank1, row1, j1, rank2, row2, j2 of the joined elements; xe of Q1, Q2, radius one, radius two
    const double *xj_d;                     // [n_xj][8] k, nu, l1[3], l2[3]
};
// host-side staging of the per-thread bookkeeping before it is packed (C = slots per thread): joint evaluated by the
// thread | per slot: published xe, overwrite source xe, number of node items | per slot BR2_FAST_NODE_ITEMS items
#define BR2_TILE_XI_JOINT 0
#define BR2_TILE_XI_PUB(C, j) (1 + (j))
#define BR2_TILE_XI_OW(C, j) (1 + (C) + (j))
#define BR2_TILE_XI_NCNT(C, j) (1 + 2 * (C) + (j))
#define BR2_TILE_XI_NITEM(C, j, q) (1 + 3 * (C) + (j) * BR2_FAST_NODE_ITEMS + (q))
#define BR2_TILE_XI_N(C) (1 + 3 * (C) + (C) * BR2_FAST_NODE_ITEMS)

#define BR2_KD_COUNT 14  // [0..2] dt c_t g, [3..4] +-0.5 / rest Voronoi length, [5] -nu_j / 4, [6] k_j / 2, [7] -k_rep / 2, [8..10] dt / J, [11] dt / 2, [12..13] shear stiffness / rest length
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
    double kd[BR2_KD_COUNT];   // loop-invariant products of uni and dt for the tile kernel (br2_step fills them)
    TilePlan tile;
    // launch chunking (tile kernel's work queue; the exact-path replay resumes at the chunk that failed:
    // status = 1 + chunk).  One chunk = the whole launch for the other kernels.
    int chunk_steps, n_chunks;
    long long queue_group;                // instances per group of the tile kernel's work queue (see the kernel)
    double chunk_t0[BR2_MAX_CHUNKS];      // fp64-accumulated start time of each chunk
    // per-instance material parameters (tile kernels' kPerInst instantiations; br2_set_instance_params)
    const double *inst_uni;               // [batch or 1][n_cls][BR2_FC_COUNT] what `uni` holds, per instance and class of rods
    long long inst_stride;                // rows of inst_uni per instance: n_cls, or 0 when every instance reads the same table
                                          // (assemblies whose rods differ in their constants, e.g. segments of different length)
    const double *inst_act_d;             // [batch][n_act][3] actuation torque per unit pressure, per instance
    const double *inst_joint_scale;       // [batch] factor on the tip-to-base joints' stiffness
    int n_act;
    unsigned long long *fail_counter;     // instances flagged by the optimistic pass of this launch
    unsigned long long *work_counter;     // next item of the queue (zeroed before every launch)
    int *progress;                        // [batch] chunks of the instance already written back
};

}  // namespace br2
