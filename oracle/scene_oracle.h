/*
 * scene_oracle.h -- CPU oracle for the PVE-MCC environment step (lane_num = 12).
 *
 * TEST INFRASTRUCTURE ONLY.  This is a sequential restatement of the reference's
 * traffic_interaction_scene.py (step 1501-1539, scene_update 222-376, add_new_veh 378-433,
 * delete_vehicle 435-444, get_virtual_distance 733-803, get_p 1250-1290, get_state 1292-1338,
 * virtual_lane_search_closer 1340-1405, check_lock 1469-1499).  Only tests/,
 * __graft_entry__.smoke() and bench.py's CPU-baseline legs may load it; the product path
 * (pve_mcc_for_unsignalized_intersection_b200/) never does.
 *
 * Parity status: PINNED.  tests/test_oracle_golden.py replays the traces under tests/golden/,
 * which were produced by running the unmodified reference scene (tests/golden/make_golden.py).
 */
#ifndef SCENE_ORACLE_H
#define SCENE_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_NLANE 12
#define ORC_OBS_W 28
#define ORC_OBS_H 7
#define ORC_NN 6

/* Scalars of the reference constructor (TIS:21-23) plus geometry constants that the Python
 * host derives with the reference's own expressions (TIS:148-152, 182-186, 1251). */
typedef struct orc_params {
    double dis_ctl, lane_cw, dt, dt2, vm, vM, am, aM, v0, collision_thr;
    double lane_in;          /* lane_info[m][0] = dis_ctl - 6*cw                   TIS:149 */
    double lane_len[3];      /* lane_info[m][1]: left arc, straight, right arc     TIS:149-151 */
    double remove_p;         /* -dis_ctl + int((12+1)/2)*cw                        TIS:341 */
    double cita, alpha, beta, gama, gama2;                                      /* TIS:182-186 */
    double rot_cos[4], rot_sin[4];   /* np.cos / np.sin of 3.141593/2*k            TIS:1251,1287 */
} orc_params;

/* Flat state view: B environments, veh_cap dense vehicle slots each, (lane asc, j asc) order. */
typedef struct orc_state_view {
    int32_t *tick;               /* [B]      scene updates done so far                       */
    int32_t *lane_n;             /* [B][12]  vehicles per lane                               */
    int32_t *veh_rec;            /* [B][12]  next arrival row per lane                       */
    int32_t *head_lane, *head_j; /* [B][12]  identity of virtual_lane_4[d][0], -1 if empty   */
    int32_t *id_seq, *passed_veh;/* [B]                                                      */
    int64_t *passed_step_total;  /* [B]                                                      */
    double *p, *v, *a, *jerk_sum;            /* [B][veh_cap] */
    int32_t *collision, *step, *seq_in_lane, *uid;   /* [B][veh_cap] */
    uint8_t *flags;              /* [B][veh_cap] bit0 control, bit1 finish, bit2 lock        */
    int8_t *lock_a;              /* [B][veh_cap]                                             */
    double *row0;                /* [B][veh_cap][28] stored observation row 0                */
} orc_state_view;

/* Dense per-agent outputs; env b owns rows agent_offset[b] .. agent_offset[b+1]. */
typedef struct orc_outputs {
    int64_t *agent_offset;   /* [B+1] filled by orc_step                                     */
    int32_t *ids;            /* [A][2]  (lane, j) before removal                             */
    int32_t *uid;            /* [A]                                                          */
    double *obs;             /* [A][7][28]                                                   */
    double *reward;          /* [A]                                                          */
    int32_t *cpv;            /* [A]  collision count reported for the agent (TIS:339)        */
    uint8_t *status;         /* [A]  bit0 Done, bit1 scheduled for removal, bit2 finished    */
    double *jerk_sum;        /* [A]  value appended to `jerks` when bit2 is set              */
    int32_t *nn;             /* [A][6][2] closer_cars                                        */
    int32_t *collisions;     /* [B] */
    int32_t *lock;           /* [B] */
    int32_t *n_removed;      /* [B] */
    int32_t *q5_undefined;   /* [B] reward[-1] written with no reward yet (reference: IndexError) */
} orc_outputs;

typedef struct orc_scene orc_scene;

orc_scene *orc_create(int32_t n_envs, int32_t veh_cap, const orc_params *prm);
void orc_destroy(orc_scene *s);
/* arrive: [B][K][12] seconds; kvalid: [B][12] rows usable per lane (ascending prefix).
 * warmup != 0 repeats empty scene updates until a vehicle exists, like TIS:214-220. */
int32_t orc_reset(orc_scene *s, const double *arrive, const int32_t *kvalid, int32_t K, int32_t warmup);
int32_t orc_set_state(orc_scene *s, const orc_state_view *in);
int32_t orc_get_state(const orc_scene *s, orc_state_view *out);
/* number of agents the next orc_step will emit (= controlled vehicles now) */
int64_t orc_count_agents(const orc_scene *s);
/* one tick for every environment: step() for each vehicle, scene_update(), delete_vehicle() */
int32_t orc_step(orc_scene *s, const float *actions, orc_outputs *out, int32_t n_threads);
int32_t orc_overflow(const orc_scene *s);
/* out[B][veh_cap]: veh["control"] of every slot, 0 past the live count (main.py:399-405) */
void orc_control_mask(const orc_scene *s, uint8_t *out);

#ifdef __cplusplus
}
#endif
#endif
