"""B = 1 view of the batched scene with the members main.py uses.

The reference has no plugin boundary; its driver (main.py:225-311, 394-441, 552-575) talks to the
Python object ``TrafficInteraction`` directly:

    env = TrafficInteraction(arrive_time, 150, args, vm=6, lane_num=12)      # MAIN:230 / 394
    for lane in range(12):
        for ind, veh in enumerate(env.veh_info[lane]):                       # MAIN:234-241
            a = actor(veh["state"][0]) if veh["control"] else 0
            env.step(lane, ind, a)
    ids, state_next, reward, actions, collisions, estm, cpv, jerks, lock = env.scene_update()
    ... env.veh_info[i][j]["Done"], ["buffer"], ["count"] via ids ...        # MAIN:243-266
    env.delete_vehicle()                                                     # MAIN:311

This class offers exactly that surface on top of ``BatchedScene`` (one intersection on the GPU),
so a main.py-style loop runs unchanged.  It is glue for drop-in use and for parity tests that read
like the reference's driver; throughput work should use ``BatchedScene`` directly.  All compute
still happens in the CUDA kernels: there is no host-side simulation here.
"""
import numpy as np
import torch

from . import _native as N
from .config import NLANE, OBS_H, OBS_W, SceneConfig
from .scene import BatchedScene


class TrafficInteraction:
    def __init__(self, arrive_time, dis_ctl, args, deltaT=0.1, vm=5, vM=13, am=-3, aM=3, v0=10, diff_max=220,
                 lane_cw=2.5, loc_con=True, show_col=False, virtual_l=True, lane_num=12,
                 device="cuda:0", veh_cap=128, agent_cap=96, _library=None):
        if not loc_con:
            raise NotImplementedError("loc_con=False selects a dead branch of the reference (TIS:1524-1526)")
        cfg = SceneConfig(vm=vm, collision_thr=getattr(args, "collision_thr", 2), dis_ctl=dis_ctl, deltaT=deltaT,
                          vM=vM, am=am, aM=aM, v0=v0, lane_cw=lane_cw, lane_num=lane_num,
                          o_agent_num=getattr(args, "o_agent_num", 6))
        self.cfg = cfg
        self.lane_num, self.deltaT, self.vm, self.vM, self.am, self.aM, self.v0 = lane_num, deltaT, vm, vM, am, aM, v0
        self.lane_cw, self.dis_control, self.collision_thr = lane_cw, dis_ctl, cfg.collision_thr
        self.closer_veh_num, self.c_mode = cfg.o_agent_num, getattr(args, "c_mode", "closer")
        self.arrive_time = np.asarray(arrive_time, dtype=np.float64)
        self.scene = BatchedScene(1, cfg, veh_cap=veh_cap, agent_cap=agent_cap, device=device, _library=_library)
        self.scene.reset(self.arrive_time, warmup=True)              # TIS:214-220
        self._host = self.scene.make_host_outputs()
        self._act = torch.zeros(1, self.scene.veh_cap, dtype=torch.float32)
        if self.scene.device.type == "cuda":
            self._act = self._act.pin_memory()
        self._records = {}                  # uid -> vehicle dict (driver-owned keys persist)
        self.veh_info = [[] for _ in range(NLANE)]
        self.delete_veh = []
        self._pending = None
        self._refresh(self.scene.get_state(), rebuild=True)

    # ---- host mirror of the device state --------------------------------------------------------
    def _new_record(self, uid, lane):
        return {"intention": lane % 3, "buffer": [], "route": lane, "count": 0, "Done": False, "p": 0.0,
                "jerk": 0, "jerk_sum": 0.0, "lock_a": 0, "lock": False, "vir_header": [-1, -1], "vir_dis": 100,
                "v": self.v0, "a": 0, "action": 0, "closer_p": 150, "lane": lane, "header": False, "reward": 10,
                "dis_front": 50, "seq_in_lane": -1, "control": True, "state": np.zeros((OBS_H, OBS_W)),
                "step": 0, "collision": 0, "finish": False, "estm_collision": 0, "estm_arrive_time": 0.0,
                "id_info": [uid, 0]}

    def _refresh(self, st, rebuild):
        """Copy the device state into the per-vehicle dicts; ``rebuild`` re-creates the lane lists in
        device order (after removal), otherwise only newly arrived vehicles are appended."""
        lane_n = st["lane_n"][0]
        self.current_time = float(st["tick"][0]) * self.deltaT
        self.id_seq = int(st["id_seq"][0])
        self.passed_veh = int(st["passed_veh"][0])
        self.passed_veh_step_total = int(st["passed_step_total"][0])
        self.veh_rec = [int(x) for x in st["veh_rec"][0]]
        new_lists = [[] for _ in range(NLANE)]
        k = 0
        for lane in range(NLANE):
            for _ in range(int(lane_n[lane])):
                uid = int(st["uid"][0, k])
                rec = self._records.get(uid)
                fresh = rec is None
                if fresh:
                    rec = self._records[uid] = self._new_record(uid, lane)
                fl = int(st["flags"][0, k])
                rec.update(p=float(st["p"][0, k]), v=float(st["v"][0, k]), a=float(st["a"][0, k]),
                           jerk_sum=float(st["jerk_sum"][0, k]), collision=int(st["collision"][0, k]),
                           step=int(st["step"][0, k]), control=bool(fl & N.F_CONTROL), finish=bool(fl & N.F_FINISH),
                           lock=bool(fl & N.F_LOCK), lock_a=int(st["lock_a"][0, k]))
                new_lists[lane].append(rec)
                if fresh and not rebuild:
                    self.veh_info[lane].append(rec)           # TIS:396: arrivals join the lane's tail
                k += 1
        self._post_lists = new_lists
        if rebuild:
            self.veh_info = new_lists
            alive = {r["id_info"][0] for lst in new_lists for r in lst}
            for uid in [u for u in self._records if u not in alive]:
                del self._records[uid]
        self.veh_num = [len(x) for x in self.veh_info]

    # ---- the reference's members ----------------------------------------------------------------
    def step(self, i, j, eval_a):
        """TIS:1501-1539.  Actions are buffered; the kinematics run on the GPU inside scene_update()."""
        off = sum(len(self.veh_info[q]) for q in range(i))
        self._act[0, off + j] = float(eval_a)

    def scene_update(self):
        """TIS:222-376: returns ids, re_state, reward, actions, collisions, estm_collisions,
        collisions_per_veh, jerks, lock."""
        if self.scene.device.type == "cuda":
            n = self.scene.step_host(self._act, self._host, copy_obs=True)
            o = self._host
        else:
            o = self.scene.step(self._act)
            n = o.n_agents
        self._act.zero_()
        ids_t = o.ids[:n].numpy()
        obs = o.obs[:n].numpy().astype(np.float64)
        status = o.status[:n].numpy()
        ids = [[int(a), int(b)] for a, b in ids_t[:, 1:3]]
        re_state = [obs[k].copy() for k in range(n)]
        reward = [float(x) for x in o.reward[:n].numpy()]
        actions = [[float(row[2]) for row in obs[k]] for k in range(n)]                # TIS:290
        cpv = [[int(c), 0] for c in o.cpv[:n].numpy()]
        jerks = [float(x) for x, s in zip(o.jerk_sum[:n].numpy(), status) if s & N.ST_FINISHED]
        self.delete_veh = []
        for k in range(n):
            lane, j = ids[k]
            rec = self.veh_info[lane][j]
            rec["state"] = re_state[k]                                                 # TIS:288
            rec["count"] += 1                                                          # TIS:292
            rec["Done"] = bool(status[k] & N.ST_DONE)
            if status[k] & N.ST_REMOVED:
                self.delete_veh.append([lane, j])
        self._refresh(self.scene.get_state(), rebuild=False)
        # vehicles removed this tick that were not agents (past the exit) are found by difference
        alive = {id(r) for lst in self._post_lists for r in lst}
        for lane in range(NLANE):
            for j, rec in enumerate(self.veh_info[lane]):
                if id(rec) not in alive:
                    rec["Done"] = True
                    if [lane, j] not in self.delete_veh:
                        self.delete_veh.append([lane, j])
        return (ids, re_state, reward, actions, int(o.env_collisions[0]), 0, cpv, jerks, int(o.env_lock[0]))

    def delete_vehicle(self):
        """TIS:435-444.  The device already compacted the lanes; drop the flagged records here."""
        self.veh_info = self._post_lists
        alive = {r["id_info"][0] for lst in self.veh_info for r in lst}
        for uid in [u for u in self._records if u not in alive]:
            del self._records[uid]
        self.veh_num = [len(x) for x in self.veh_info]
        self.delete_veh = []
