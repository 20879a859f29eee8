import os, sys, time
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", "/root/repo"))
from collision_b200 import scenes
from collision_b200.solver import CollisionSolver3d, PROXIMITY, COLLISION
from collision_b200 import dist as D
local = int(os.environ["LOCAL_RANK"]); torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
sc = scenes.layered_cloth(8, 251)
s = CollisionSolver3d(device=local); CollisionSolver3d.set_params_from(sc.params); s.assembleFromInterface(sc, sc.dt)
st = D.DistributedSolver(s, mode=os.environ.get("MODE", "owner"))
x, xn = sc.x, sc.x_new()
T = {}
def tick(name, t0):
    torch.cuda.synchronize(); t = time.perf_counter(); T[name] = T.get(name, 0) + (t - t0); return t
orig_ex = st._exchange_owner if st.mode == "owner" else st._exchange
if os.environ.get("FINE") and st.mode == "owner":
    L = s.ctx.L
    for name in ("clsn_bucket_records", "clsn_import_records", "clsn_apply_stage", "clsn_export_records"):
        f = getattr(L, name)
        def mk(f, name):
            def w(*a):
                torch.cuda.synchronize(); t0 = time.perf_counter(); r = f(*a); tick(name, t0); return r
            return w
        setattr(L, name, mk(f, name))
    for name in ("all_gather_into_tensor", "all_to_all_single"):
        f = getattr(dist, name)
        def mk2(f, name):
            def w(*a, **k):
                torch.cuda.synchronize(); t0 = time.perf_counter(); r = f(*a, **k); tick(name, t0); return r
            return w
        setattr(D.dist, name, mk2(f, name))
import ctypes as C
def timed_detect(mode):
    t0 = time.perf_counter(); r = s.detect(mode); tick("detect", t0); return r
def timed_exchange():
    t0 = time.perf_counter(); r = orig_ex(); tick("exchange+apply", t0); return r
for it in range(6):
    if it == 3: T.clear()
    s.upload(x, xn)
    torch.cuda.synchronize(); dist.barrier(); t00 = time.perf_counter()
    s.avg_velocity()
    stt = timed_detect(PROXIMITY); n = timed_exchange()
    if st.mode != "owner":
        t0 = time.perf_counter(); s.apply(True); tick("apply", t0)
    coll, cd = True, 0
    while coll and cd < 5:
        stt = timed_detect(COLLISION); n = timed_exchange(); coll = n > 0; cd += 1
        if st.mode != "owner":
            t0 = time.perf_counter(); s.apply(True); tick("apply", t0)
    s.boundary(); s.final_position(); s.synchronize()
    tick("total", t00)
if dist.get_rank() == 0:
    print(st.mode, {k: round(1e3 * v / 3, 2) for k, v in T.items()})
dist.destroy_process_group()
