"""Secondary configs of BASELINE.json (not the bench line): steps/s and first-steps parity per scene."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import oracle_lib
from parity_util import compare_worlds, is_exact
from rapier_b200 import scenes
from rapier_b200.world import PhysicsWorld

CONFIGS = [("b3d_many_pyramids", scenes.many_pyramids, 300), ("b3d_many_pyramids_80x20", scenes.many_pyramids_label, 300),
           ("pyramid3_50", lambda: scenes.pyramid3(50), 100), ("b3d_joint_grid_100", lambda: scenes.joint_grid(100), 200),
           ("keva3_5", lambda: scenes.keva(5), 200),
           ("large_world_300", lambda: scenes.large_world(grid=300, spheres=100), 200),
           ("convex_polyhedron3", lambda: scenes.convex_polyhedra(25), 200)]   # examples3d/convex_polyhedron3.rs (625 round hulls)   # (1000 = the reference size; opt-in by name)
only = sys.argv[1:]
out = []


def large_world_protocol_perf():
    """BASELINE configs[4] as the reference runs it (b3d_large_world.rs): 10^6 static colliders, one sphere inserted
    every 5 steps up to 100, 600 steps; wall-clock steps/s INCLUDING the insertions (host synchronisation + upload)."""
    from incremental_cases import OracleSets
    grid, spheres = 1000, 100
    w = PhysicsWorld(scenes.large_world_floor(grid))
    w.reserve(spheres, grid * grid + spheres)
    o = OracleSets(scenes.large_world_floor(grid), threads=8)
    exact = True
    dropped = 0
    for step in range(30):   # parity prefix (5 insertions)
        if step > 0 and step % 5 == 0:
            for world in (w, o):
                bb, cb = scenes.large_world_sphere(dropped, grid, spheres)
                world.insert(bb, cb)
            dropped += 1
        w.step(); o.step()
    exact = is_exact(compare_worlds(w, o))
    w.physics_pipeline.synchronize()
    t0 = time.perf_counter()
    for step in range(30, 630):
        if dropped < spheres and step % 5 == 0:
            bb, cb = scenes.large_world_sphere(dropped, grid, spheres)
            w.insert(bb, cb)
            dropped += 1
        w.step(sync=False)
    w.physics_pipeline.synchronize()
    wall = time.perf_counter() - t0
    w.physics_pipeline.enable_profiling(True)
    w.step(100)
    c = w.counters()
    t1 = time.perf_counter(); o.step(20); cpu = 20 / (time.perf_counter() - t1)
    rec = dict(scene="large_world_1000_protocol", bodies=c["num_bodies"], colliders=c["num_colliders"], pairs=c["num_pairs"], manifolds=c["num_active_manifolds"],
               first30_steps_bit_exact=bool(exact), steps_per_s_with_insertions=600 / wall, steady_collide_ms=c["collision_detection_ms"], steady_solve_ms=c["solver_ms"],
               steady_steps_per_s=1000.0 / max(c["collision_detection_ms"] + c["solver_ms"], 1e-9), cpu_port_steps_per_s_8thr_after_30_steps=cpu)
    print(json.dumps(rec), flush=True)


if not only or "large_world_1000_protocol" in only:
    large_world_protocol_perf()
for name, make, steps in CONFIGS:
    if only and name not in only:
        continue
    scene = make()
    w = PhysicsWorld(scene)
    o = oracle_lib.OracleWorld(scene, threads=16)
    exact = True
    for i in range(3):
        w.step(); o.step()
        d = compare_worlds(w, o, tables=(i == 2))
        exact = exact and is_exact(d)
    w.step(30)
    w.physics_pipeline.enable_profiling(True)
    t0 = time.perf_counter()
    w.step(steps)
    wall = time.perf_counter() - t0
    c = w.counters()
    t1 = time.perf_counter(); o.step(5); cpu = 5 / (time.perf_counter() - t1)
    st = w.debug_read("state", np.int32)
    rec = dict(scene=name, bodies=c["num_bodies"], pairs=c["num_pairs"], manifolds=c["num_active_manifolds"], joints=c["num_joints"],
               items=int(st[7]), streamed=int(st[19]), resident=int(st[20]), colors=c["num_colors"], large_bodies=int(st[12]), first3_steps_bit_exact=bool(exact),
               steps_per_s=steps / wall, collide_ms=c["collision_detection_ms"], solve_ms=c["solver_ms"], cpu_port_steps_per_s_16thr=cpu)
    print(json.dumps(rec), flush=True)
    out.append(rec)
