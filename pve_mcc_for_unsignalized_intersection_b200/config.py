"""Scene constants, derived with the reference's own expressions.

Mirrors ``TrafficInteraction.__init__`` (traffic_interaction_scene.py:21-220, 12-lane branch
146-186).  Every derived constant is computed here, on the host, in Python floats -- the same
IEEE operations in the same order as the reference -- and handed to the CUDA library through
``pve_config`` (include/pve_mcc.h), so the device never re-derives geometry.
"""
import ctypes as C
import dataclasses
import math

import numpy as np

NLANE = 12
OBS_H, OBS_W, NNBR = 7, 28, 6

#: lane2lane of the 12-lane intersection (TIS:153-166)
LANE2LANE = [
    [10, 3, 9, 7], [10, 6, 3, 4], [],
    [1, 6, 0, 10], [1, 9, 6, 7], [],
    [4, 9, 3, 1], [4, 0, 9, 10], [],
    [7, 0, 6, 4], [7, 3, 0, 1], [],
]


class PveConfig(C.Structure):
    """ctypes image of ``pve_config``."""
    _fields_ = [
        ("n_envs", C.c_int32), ("veh_cap", C.c_int32), ("agent_cap", C.c_int32), ("threads", C.c_int32),
        ("out_cap", C.c_int64),
        ("dt", C.c_double), ("dt2", C.c_double),
        ("vm", C.c_double), ("vM", C.c_double), ("am", C.c_double), ("aM", C.c_double), ("v0", C.c_double),
        ("collision_thr", C.c_double), ("lane_in", C.c_double), ("lane_len", C.c_double * 3),
        ("remove_p", C.c_double), ("lane_cw", C.c_double),
        ("vd_a1", (C.c_double * 4) * 2), ("vd_a2", (C.c_double * 4) * 2), ("vd_b", (C.c_double * 4) * 2),
        ("rot_cos", C.c_double * 4), ("rot_sin", C.c_double * 4),
        ("zero_uncontrolled", C.c_int32), ("lane_num", C.c_int32),
        ("n4_T", (C.c_double * 7) * 4), ("n4_C", (C.c_double * 7) * 4), ("n4_C2", (C.c_double * 7) * 4),
        ("n4_rw", C.c_double * 3),
    ]


@dataclasses.dataclass
class SceneConfig:
    """Arguments of the reference constructor that reach the environment step.

    ``TrafficInteraction(arrive_time, dis_ctl, args, deltaT=0.1, vm=5, vM=13, am=-3, aM=3, v0=10,
    lane_cw=2.5, lane_num=12)`` (TIS:21-23); ``args.collision_thr`` (TIS:32); ``args.o_agent_num``
    must be 6 and ``lane_num`` 12 (the only geometry in scope).
    main.py uses ``vm=6`` for training (MAIN:230) and the default ``vm=5`` for testing (MAIN:394).
    """
    vm: float = 5
    collision_thr: float = 2
    dis_ctl: float = 150
    deltaT: float = 0.1
    vM: float = 13
    am: float = -3
    aM: float = 3
    v0: float = 10
    lane_cw: float = 2.5
    lane_num: int = 12
    o_agent_num: int = 6
    #: take the action of every uncontrolled vehicle as 0 like the reference driver does (MAIN:401-405)
    zero_uncontrolled_actions: bool = False

    def __post_init__(self):
        if self.lane_num not in (12, 8, 4):
            raise NotImplementedError("lane_num must be 12, 8 or 4 (lane_num=%r: the reference's T-junction branch dies in "
                                      "its constructor, direction_num is never set)" % self.lane_num)
        if self.o_agent_num != 6:
            raise NotImplementedError("o_agent_num must be 6 (28-wide observation rows)")

    # ---- derived geometry -----------------------------------------------------------------
    def lane_len(self):
        cw = self.lane_cw
        if self.lane_num == 4:
            return [3.1415 / 2 * 3 * cw, 4 * cw, 3.1415 / 2 * cw]                 # TIS:53-55
        if self.lane_num == 8:
            return [3.1415 / 2 * 5 * cw, 8 * cw, 3.1415 / 2 * cw]                 # TIS:101-103
        return [3.1415 / 2 * 7 * cw, 12 * cw, 3.1415 / 2 * cw]                    # TIS:149-151

    def lane_in(self):
        if self.lane_num == 4:
            return self.dis_ctl - 2 * self.lane_cw                               # TIS:53
        if self.lane_num == 8:
            return self.dis_ctl - 4 * self.lane_cw                               # TIS:101
        return self.dis_ctl - 6 * self.lane_cw                                   # TIS:149

    def four_lane_tables(self):
        """``lane_num=4``: ``get_virtual_distance`` (TIS:453-531) as ``(T, C)`` -- member iff ``p1 - T > 0``,
        ``vd = abs(p1 - T) + C`` -- indexed ``[ego route % 3][position in lane2lane[ego route]]``, and the three
        constants of the ``get_state`` rewrite (TIS:1304-1316); every expression keeps the reference's association."""
        cw = self.lane_cw
        alpha = math.atan((4 - math.sqrt(2)) / (4 + math.sqrt(2)))               # TIS:79
        alpha_ = math.atan((4 + math.sqrt(2)) / (4 - math.sqrt(2)))              # TIS:80
        beta = math.atan(2 / math.sqrt(5))                                       # TIS:81
        beta_ = math.atan(math.sqrt(5) / 2)                                      # TIS:82
        gama = math.atan(1 / 2 * math.sqrt(2))                                   # TIS:83
        T = [[4 * cw - 3 * cw * math.cos(gama), (1.5 * 3.1415) * cw * (alpha_ / (0.5 * 3.1415)),      # TIS:456, 478
              1.5 * 3.1415 * cw * beta / (0.5 * 3.1415), 1.5 * 3.1415 * cw * beta_ / (0.5 * 3.1415),  # TIS:468, 473
              3 * cw * math.cos(gama), 0.0, 0.0],                                                   # TIS:483, 488, 493
             [cw, 1.5 * 3.1415 * cw * gama / (0.5 * 3.1415),                                          # TIS:499, 504
              1.5 * 3.1415 * cw * (0.5 * 3.1415 - gama) / (0.5 * 3.1415), 3 * cw, 0.0, 0.0, 0.0],     # TIS:509, 514
             [0.0] * 7]
        Cc = [[3 * cw * (0.5 * 3.1415 - gama), (1.5 * 3.1415) * cw * (alpha / (0.5 * 3.1415)),        # TIS:458, 480
               1.5 * 3.1415 * cw * beta_ / (0.5 * 3.1415), 1.5 * 3.1415 * cw * beta / (0.5 * 3.1415),  # TIS:470, 475
               1.5 * 3.1415 * cw * (gama / (0.5 * 3.1415)), 0.0, 0.0],                               # TIS:485
              [3 * cw, 3 * cw * math.cos(gama), 4 * cw - 3 * cw * math.cos(gama), cw, 0.0, 0.0, 0.0],  # TIS:501-516
              [0.0] * 7]
        rw = [(alpha_ - alpha) * 3 * cw, alpha_ * 3 * cw, alpha * 3 * cw]         # TIS:1304, 1309, 1316
        return T, Cc, rw

    def eight_lane_tables(self):
        """``lane_num=8``: ``get_virtual_distance`` (TIS:537-660) as ``(T, C, C2)`` -- member iff ``p1 - T > 0``,
        ``vd = (abs(p1 - T) + C) - C2`` -- indexed ``[ego route % 4][position in lane2lane[ego route]]``; every expression
        keeps the reference's association (``C2`` exists for TIS:638: ``abs(d) + 8 cw - sqrt(24) cw``)."""
        cw = self.lane_cw
        s24 = math.sqrt(24)
        T = [[8 * cw - s24 * cw, math.atan(3 / 4) * 5 * cw, 4 * cw, math.atan(4 / 3) * 5 * cw, 4 * cw, s24 * cw, 0.0],      # TIS:542-576
             [3 * cw, 3 * cw, math.atan(3 / 4) * 5 * cw, math.atan(4 / 3) * 5 * cw, 5 * cw, 5 * cw, 0.0],                   # TIS:581-615
             [cw, cw, math.atan(1 / s24) * 5 * cw, math.atan(s24) * 5 * cw, 7 * cw, 7 * cw, 0.0],                           # TIS:621-655
             [0.0] * 7]                                                                                                    # TIS:658
        Cc = [[math.atan(s24) * 5 * cw, math.atan(4 / 3) * 5 * cw, math.atan(4 / 3) * 5 * cw, math.atan(3 / 4) * 5 * cw,
               math.atan(3 / 4) * 5 * cw, math.atan(1 / s24) * 5 * cw, 0.0],
              [7 * cw, 5 * cw, 4 * cw, 4 * cw, 3 * cw, cw, 0.0],
              [7 * cw, 5 * cw, s24 * cw, 8 * cw, 3 * cw, cw, 0.0],
              [0.0] * 7]
        C2 = [[0.0] * 7, [0.0] * 7, [0.0, 0.0, 0.0, s24 * cw, 0.0, 0.0, 0.0], [0.0] * 7]
        return T, Cc, C2

    def spawn_p(self, lane):
        m = lane % 3                                                             # TIS:393-394
        return sum([self.lane_in(), self.lane_len()[m]])                         # TIS:395

    def remove_p(self):
        return -self.dis_ctl + int((self.lane_num + 1) / 2) * self.lane_cw       # TIS:341-342 (lane_num 12: -135, 8: -140, 4: -145)

    def angles(self):
        cw = self.lane_cw
        cita = (2 * math.sqrt(10) - 6) * cw                                      # TIS:182
        alpha = math.atan((6 * cw + cita) / (3 * cw))                            # TIS:183
        beta = math.pi / 2 - alpha                                               # TIS:184
        gama = math.atan((math.sqrt(13) * cw) / (6 * cw))                        # TIS:185
        gama2 = math.pi / 2 - gama                                               # TIS:186
        return cita, alpha, beta, gama, gama2

    def virtual_distance_table(self):
        """``get_virtual_distance`` (TIS:733-803) as ``delta = (p1 - a1) + a2; vd = b + delta``.

        Index ``[m][k]``: ``m`` = ego movement (0 left, 1 straight), ``k`` = position of the other
        vehicle's lane in ``lane2lane[ego]``.  The products keep the reference's association,
        e.g. ``self.beta * 7 * self.lane_cw`` is ``(beta * 7) * cw``.
        """
        cw = self.lane_cw
        cita, alpha, beta, gama, gama2 = self.angles()
        a1 = [[0.0] * 4 for _ in range(2)]
        a2 = [[0.0] * 4 for _ in range(2)]
        b = [[0.0] * 4 for _ in range(2)]
        # straight ego, TIS:733-766
        a1[1][0], b[1][0] = 3 * cw, 9 * cw                                       # TIS:736, 740
        a1[1][1], b[1][1] = beta * 7 * cw, 6 * cw + cita                         # TIS:744, 749
        a1[1][2], b[1][2] = alpha * 7 * cw, 6 * cw - cita                        # TIS:753, 758
        a1[1][3], b[1][3] = 9 * cw, 3 * cw                                       # TIS:761, 765
        # left-turn ego, TIS:771-799
        a1[0][0], a2[0][0], b[0][0] = 6 * cw, cita, alpha * 7 * cw               # TIS:773, 777
        a1[0][1], b[0][1] = gama * 7 * cw, gama2 * 7 * cw                        # TIS:780, 784
        a1[0][2], b[0][2] = gama2 * 7 * cw, gama * 7 * cw                        # TIS:787, 791
        a1[0][3], a2[0][3], b[0][3] = 6 * cw, -cita, beta * 7 * cw               # TIS:794, 798
        return a1, a2, b

    def rotation(self):
        rot = [3.141593 / 2 * k for k in range(4)]                               # TIS:1251
        return [float(np.cos(r)) for r in rot], [float(np.sin(r)) for r in rot]  # TIS:1287-1288

    def to_native(self, n_envs, veh_cap, agent_cap, out_cap, threads=0):
        c = PveConfig()
        c.n_envs, c.veh_cap, c.agent_cap, c.threads, c.out_cap = n_envs, veh_cap, agent_cap, threads, out_cap
        c.dt, c.dt2 = self.deltaT, pow(self.deltaT, 2)                           # TIS:1529
        c.vm, c.vM, c.am, c.aM, c.v0 = self.vm, self.vM, self.am, self.aM, self.v0
        c.collision_thr = self.collision_thr
        c.lane_in = self.lane_in()
        for m, L in enumerate(self.lane_len()):
            c.lane_len[m] = L
        c.remove_p = self.remove_p()
        c.lane_cw = self.lane_cw
        a1, a2, b = self.virtual_distance_table()
        for m in range(2):
            for k in range(4):
                c.vd_a1[m][k], c.vd_a2[m][k], c.vd_b[m][k] = a1[m][k], a2[m][k], b[m][k]
        cs, sn = self.rotation()
        for k in range(4):
            c.rot_cos[k], c.rot_sin[k] = cs[k], sn[k]
        c.zero_uncontrolled = int(bool(self.zero_uncontrolled_actions))
        c.lane_num = int(self.lane_num)
        if self.lane_num == 4:
            T, Cc, rw = self.four_lane_tables()
            for r in range(3):
                for k in range(7):
                    c.n4_T[r][k], c.n4_C[r][k] = T[r][k], Cc[r][k]
            for k in range(3):
                c.n4_rw[k] = rw[k]
        elif self.lane_num == 8:
            T, Cc, C2 = self.eight_lane_tables()
            for r in range(4):
                for k in range(7):
                    c.n4_T[r][k], c.n4_C[r][k], c.n4_C2[r][k] = T[r][k], Cc[r][k], C2[r][k]
        return c
